/* b200rl -- C-ABI of the B200-native learner-update path.
 *
 * The reference (ymg1114/pytorch-distributed-reinforcement-learning) is pure Python
 * and has no FFI of its own: its seams are Python call
 * boundaries.  This header declares the native entry points that the Python shims
 * in `pytorch-distributed-reinforcement-learning_b200/b200rl/` bind with ctypes; each
 * entry point cites the reference code it replaces (paths relative to the reference
 * root).  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 unless stated otherwise
 *     (the reference stores every rollout field as fp32,
 *     agents/storage_module/shared_batch.py:15,28-46);
 *   - batch-major layout (B, T, width), exactly what
 *     agents/learner.py:197-233 hands to `learning()`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return 0 on success, non-zero on error; `b200rl_last_error()` gives the text.
 *     Nothing aborts the process (SURVEY.md 8b "Errors").
 *   - H (LSTM width) is 64, the reference's `hidden_size` (utils/parameters.json:10).
 */
#ifndef B200RL_H_
#define B200RL_H_

#ifdef __cplusplus
extern "C" {
#endif

#define B200RL_H 64
#define B200RL_MAX_HEADS 32   /* NH <= 32 head outputs per net */
#define B200RL_MAX_ACT 8      /* discrete A / continuous act_dim <= 8 */

int b200rl_version(void);
const char* b200rl_last_error(void);

/* ---- per-timestep recurrences --------------------------------------------------
 * compute_gae(deltas, gamma, lambda_)          agents/learner_module/compute_loss.py:7-19
 *   deltas (B, n, 1) -> gae (B, n, 1);  g[t] = d[t] + gamma*lambda*g[t+1]           */
int b200rl_compute_gae(const float* deltas, float* gae, int B, int n,
                       float gamma, float lambda_, void* stream);

/* td target + delta + GAE fused           agents/learner_module/ppo/learning.py:48-57
 *   rew, is_fir, value (B, T, 1) -> td_target (B, T-1, 1), gae (B, T-1, 1)           */
int b200rl_td_gae(const float* rew, const float* is_fir, const float* value,
                  float* td_target, float* gae, int B, int T,
                  float gamma, float lambda_, void* stream);

/* compute_v_trace(...)                        agents/learner_module/compute_loss.py:22-66
 *   5 inputs (B, T, 1) -> rho (B,T-1,1), advantages (B,T-1,1), values_target (B,T,1) */
int b200rl_compute_v_trace(const float* behav_logp, const float* target_logp,
                           const float* is_fir, const float* rew, const float* values,
                           float* rho, float* adv, float* vs, int B, int T,
                           float gamma, float rho_bar, float c_bar, void* stream);

/* kldivergence(logits_p, logits_q)            agents/learner_module/compute_loss.py:74-77
 *   (rows, A) x2 -> (rows)                                                            */
int b200rl_kldivergence(const float* logits_p, const float* logits_q, float* kl,
                        long long rows, int A, void* stream);

/* soft_update(critic, target, tau)            agents/learner_module/compute_loss.py:69-71
 *   target <- (1-tau) target + tau src over n flat floats (src may alias target)      */
int b200rl_soft_update(const float* src, float* target, long long n, float tau, void* stream);

/* ---- torso + LSTMCell-over-T + linear heads -------------------------------------
 * Replaces MlpLSTMBase.forward / MlpLSTMActor* / MlpLSTMCritic* forward and its
 * autograd backward                                  networks/models.py:63-100,135-322
 *
 * u      (B, T, Fin)  torso input (obs, or [obs|act] for the continuous critic)
 * hx0/cx0 (B, H)      LSTM state before step 0 (`hx[:,0]`, ppo/learning.py:37-40)
 * Wt (H, Fin), bt (H) torso Linear;  W_ih/W_hh (4H, H), b_ih/b_hh (4H) PyTorch layout
 * saved activations (private layout, only lstm_bwd / lstm_wgrad read them):
 *   gates (B, T, H, 4) post-activation i,f,g,o ; c (B, T, H) ; h (B, T, H)
 * `variant`: 0 = FFMA kernel with the input projection in the loop (the independently written reference the
 *            tensor-core kernels are cross-checked against); 3 = tcgen05 input projection + tcgen05 recurrence
 *            (two kernels; any Fin).  The Python layer uses b200rl_lstm_fwd_fused below whenever Fin <= 8.       */
int b200rl_lstm_fwd(const float* u, const float* hx0, const float* cx0,
                    const float* Wt, const float* bt,
                    const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh,
                    float* gates, float* c, float* h,
                    int B, int T, int Fin, int variant, void* stream);

/* The same forward AND the heads in one persistent tcgen05 kernel (csrc/fwd_fused.cu): torso, input projection, LSTM
 * recurrence (networks/models.py:67-75) and head_out = h . W_head^T + b_head (models.py:77-78) without the HBM round trip
 * of the input projection.  Requires Fin <= 8; heads are fused when head_part != NULL (scratch of B*T*2*NH floats,
 * NH <= 4) and then head_out (B,T,NH) is written; with head_part == NULL only gates / c / h are produced.            */
int b200rl_lstm_fwd_fused(const float* u, const float* hx0, const float* cx0,
                          const float* Wt, const float* bt, const float* W_ih, const float* W_hh,
                          const float* b_ih, const float* b_hh, const float* W_head, const float* b_head,
                          float* gates, float* c, float* h, float* head_part, float* head_out,
                          int B, int T, int Fin, int NH, void* stream);

/* head_out (rows, NH) = h (rows, H) . W_head^T (NH, H) + b_head                      */
int b200rl_head_fwd(const float* h, const float* W_head, const float* b_head,
                    float* head_out, long long rows, int NH, void* stream);

/* BPTT, serial part: consumes dhead (B,T,NH) = dLoss/dhead_out and OVERWRITES `gates`
 * in place with dLoss/d(gate pre-activations) (same (B,T,H,4) layout).
 * `variant` 0: FFMA kernel (W_ih, dx unused, may be NULL).
 * `variant` 1: tcgen05 kernel; also writes dx (B,T,H) = dLoss/d(torso output, before its
 *              ReLU mask) = dpre . W_ih, which lstm_wgrad variant 3 consumes.          */
int b200rl_lstm_bwd(const float* dhead, const float* W_head, const float* W_hh, const float* W_ih,
                    float* gates, const float* c, const float* cx0, float* dx,
                    int B, int T, int NH, int variant, void* stream);

/* The tcgen05 BPTT kernel with the small gradients fused (csrc/bwd_tc.cu): besides dpre (in place in `gates`) it accumulates
 * dW_head / db_head (h recomputed as o tanh(c); NH <= 4) into slab row blockIdx of `partials` (nparts x P, >= min(ceil(B/32),
 * SMs) rows).  dx != NULL: dx (B,T,H) is stored as by lstm_bwd variant 1 and lstm_wgrad runs with variant 5 (tcgen05 dW +
 * torso pass); dx == NULL (needs Fin == 4): dWt / dbt are accumulated here too (dx straight from the accumulator, ReLU mask
 * from the recomputed torso), dx never goes to HBM and lstm_wgrad runs with variant 4 (tcgen05 dW only).                 */
int b200rl_lstm_bwd_fused(const float* dhead, const float* W_head, const float* W_hh, const float* W_ih,
                          float* gates, const float* c, const float* cx0, const float* u, const float* Wt, const float* bt,
                          float* partials, int nparts, long long P, long long off_Wt, long long off_bt,
                          long long off_Whead, long long off_bhead, float* dx, int B, int T, int Fin, int NH, void* stream);

/* BPTT, parallel part: weight gradients as `nparts` partial slabs of `P` floats each
 * (partials[part*P + off_*]); also the torso-input gradient du (B,T,Fin) if non-NULL.
 * `variant` 0: FFMA; 1: tcgen05 dW + tcgen05 dx; 2: tcgen05 dW + FFMA rest;
 *           3: tcgen05 dW + torso gradients from `dx` (see lstm_bwd variant 1; du must be NULL, Fin <= 8);
 *           4: tcgen05 dW only (the head / torso gradients were accumulated by b200rl_lstm_bwd_fused);
 *           5: tcgen05 dW + torso pass over `dx` (the head gradients were accumulated by b200rl_lstm_bwd_fused). */
int b200rl_lstm_wgrad(const float* dpre, const float* h, const float* hx0, const float* u,
                      const float* dhead, const float* Wt, const float* bt, const float* W_ih,
                      float* partials, int nparts, long long P,
                      long long off_Wt, long long off_bt, long long off_Wih, long long off_Whh,
                      long long off_bih, long long off_bhh, long long off_Whead, long long off_bhead,
                      float* du, const float* dx, int B, int T, int Fin, int NH, int variant, void* stream);

/* grad[i] = sum_part partials[part*P + i]                                             */
int b200rl_reduce_partials(const float* partials, float* grad, int nparts, long long P, void* stream);

/* ---- loss heads (scan + loss + dLoss/dhead_out in one pass) ----------------------
 * scalars: device array of 8 doubles, ACCUMULATED into (caller zeroes):
 *   [0] sum policy term  [1] sum value term  [2] sum entropy  [3] sum ratio
 *   [4] min ratio (as +value) [5] max ratio  [6],[7] algorithm specific
 * PPO head                                    agents/learner_module/ppo/learning.py:43-81
 *   head_out (B,T,NH): discrete NH=1+A -> [value, z_0..z_{A-1}];
 *                      continuous NH=1+2A (A==1) -> [value, z_mu, z_std]
 *   optional outputs (may be NULL): logits_norm (B,T,A), logp (B,T), entropy (B,T),
 *   td_target (B,T-1), gae (B,T-1), ratio (B,T-1)                                     */
int b200rl_ppo_head(const float* head_out, const float* act, const float* rew,
                    const float* is_fir, const float* behav_logp,
                    float* dhead, double* scalars,
                    float* logits_norm, float* logp, float* entropy,
                    float* td_target, float* gae, float* ratio,
                    int B, int T, int A, int continuous,
                    float gamma, float lambda_, float eps_clip,
                    float c_policy, float c_value, float c_entropy, void* stream);

/* IMPALA head                              agents/learner_module/impala/learning.py:43-69
 *   optional outputs: logits_norm, logp, entropy, rho (B,T-1), adv (B,T-1), vs (B,T)  */
int b200rl_impala_head(const float* head_out, const float* act, const float* rew,
                       const float* is_fir, const float* behav_logp,
                       float* dhead, double* scalars,
                       float* logits_norm, float* logp, float* entropy,
                       float* rho, float* adv, float* vs,
                       int B, int T, int A,
                       float gamma, float rho_bar, float c_bar,
                       float c_policy, float c_value, float c_entropy, void* stream);

/* V-MPO head                  agents/learner_module/v_mpo/learning.py:44-99, agents/learner.py:320-348
 *   head_out (B,T,1+A) = [value, z]; duals = device {log_eta, log_alpha} (fp64, as in the reference).
 *   vmpo_gae:    td target, delta, GAE scan -> gae (B,T-1); value loss sum -> scalars[1]; dLoss/dvalue -> dhead[...,0],
 *                the policy columns of dhead are zeroed.
 *   vmpo_select: per timestep column of gae_all (Bg,n): thr[t] = k-th largest over the batch (torch.topk(gae, k, dim=0));
 *                colstat[t] = column max, colstat[n+t] = sum over the selected of exp((g - max)/(eta+1e-7)).
 *                Exactly k rows per column, as topk: rows tied AT thr[t] are taken in global row order up to
 *                tie_cut[t] (INT_MAX when all of them belong to the top k).
 *                gae_all is the rank's own matrix on one GPU, the all-gathered one when the batch is sharded.
 *   vmpo_grad:   psi = softmax over all selected; adds dLoss/dz (policy + alpha*KL term) to dhead; scalars[0] += L_pi,
 *                scalars[2] += sum KL, scalars[3] += #selected, scalars[6],[7] = softmax max / normaliser.
 *   vmpo_dual_step: L_eta, L_alpha, their gradients and the RMSprop step of log_eta/log_alpha (never clipped);
 *                losses_out[6] = {L_pi, L_V, L_eta, L_alpha, total, mean KL}.
 *   N_global = B_global*(T-1) (every mean of the update), n_selected = k*(T-1).                                   */
int b200rl_vmpo_gae(const float* head_out, const float* rew, const float* is_fir,
                    float* gae, float* dhead, double* scalars, int B, int T, int A,
                    float gamma, float lambda_, float c_value, double N_global, void* stream);
int b200rl_vmpo_select(const float* gae_all, int Bg, int n, int k, const double* duals,
                       float* thr, float* colstat, int* tie_cut, void* stream);
int b200rl_vmpo_grad(const float* head_out, const float* act, const float* behav_logits, const float* gae,
                     const float* thr, const float* colstat, const int* tie_cut, const double* duals,
                     float* dhead, double* scalars, int B, int T, int A, int row_offset /* rank * B when sharded */,
                     float c_policy, double N_global, void* stream);
int b200rl_vmpo_dual_step(double* duals, double* dual_sq, const double* scalars, double* losses_out,
                          double coef_eta, double coef_alpha, double lr, double rms_alpha, double rms_eps,
                          double N_global, double n_selected, double c_policy, double c_value, void* stream);

/* ---- optimizer step ---------------------------------------------------------------
 * clip_grad_norm_(max_norm) + RMSprop(lr, alpha, eps)
 *                    agents/learner_module/ppo/learning.py:93-106, agents/learner.py:70
 * grad is scaled in place by `grad_scale` first (1/world_size after an allreduce-sum),
 * then clipped; norm_out (1 float, may be NULL) receives the pre-clip L2 norm.        */
int b200rl_clip_rmsprop(float* param, float* grad, float* sq_avg, long long n,
                        float grad_scale, float max_norm, float lr, float alpha, float eps,
                        float* norm_out, void* stream);

/* clip_grad_norm_ + Adam (bias-corrected, step = 1-based)   agents/learner.py:360-367 */
int b200rl_clip_adam(float* param, float* grad, float* m, float* v, long long n,
                     float grad_scale, float max_norm, float lr, float beta1, float beta2,
                     float eps, int step, const float* mask /* 0/1 per entry or NULL */,
                     float* norm_out, void* stream);

/* ---- data-parallel step over NVLink peer memory (csrc/peer_kernels.cu) ---------------------------------------------
 * For a batch sharded over the GPUs of one box (the reference has no multi-GPU learner) this ONE kernel
 * replaces  dist.all_reduce(grad) ; clip_grad_norm_ ; optimizer.step()  (single-GPU tail:
 * agents/learner_module/ppo/learning.py:93-106, agents/learner.py:70,360-367): flag handshake with the peers, one-shot
 * all-reduce of the flat gradient by direct peer loads (rank order, bit-identical on every rank), global-norm clip,
 * RMSprop (optimizer 0: s1 = square average, a1 = alpha) or Adam (optimizer 1: s1 = m, s2 = v, a1/a2 = betas).
 *   peer_alloc / peer_open / peer_close / peer_free: a cudaMalloc'd, zeroed region and its 64-byte CUDA-IPC handle; the
 *     other ranks map it with peer_open (peer access is enabled lazily).
 *   peer_step: peer_grad[r] / peer_scal[r] = rank r's gradient (n floats, n % 4 == 0) and 8 fp64 loss sums for THIS step
 *     (the caller alternates two buffers by step parity), peer_flags[r] = base of rank r's flag array (world x 32 u32),
 *     my_flags = this rank's; gsync / part = local scratch (one u32, zeroed once; 64 doubles); epoch = 1, 2, 3, ... per
 *     launch; gsync_target = sum of `grid` over all launches so far including this one.  grad_out receives the reduced,
 *     scaled, clipped gradient, scal_out the reduced loss sums ([0..3] summed, [4] min, [5] max over ranks), *status
 *     becomes 1 if a peer never announced its gradient within the timeout (default 30 s) -- the kernel never hangs.   */
int b200rl_peer_alloc(long long bytes, void** ptr_out, unsigned char* handle64_out);
int b200rl_peer_open(const unsigned char* handle64, void** ptr_out);
int b200rl_peer_close(void* ptr);
int b200rl_peer_free(void* ptr);
int b200rl_peer_set_timeout_ms(long long ms);
int b200rl_peer_step(int optimizer, float* param, float* s1, float* s2, const float* mask, float* grad_out,
                     const void* const* peer_grad, const void* const* peer_scal, void* const* peer_flags,
                     void* my_flags, void* gsync, double* part, double* scal_out, float* norm_out, int* status,
                     long long n, int rank, int world, unsigned epoch, int grid, unsigned gsync_target,
                     float grad_scale, float max_norm, float lr, float a1, float a2, float eps, int adam_step,
                     void* stream);

/* ---- SAC heads (element-wise over (B,T); scalars as above: [0] policy, [1] value, [2] entropy term) ----------
 * discrete SAC                                     agents/learner_module/sac/learning.py:36-143
 *   z, q1, q2, tq1, tq2 (B,T,A); log_alpha device double; dz / dq1 / dq2 (B,T,A); td_out (B,T-1) or NULL          */
int b200rl_sac_actor_head(const float* z, const float* q1, const float* q2, const double* log_alpha,
                          float* dz, double* scalars, int B, int T, int A, void* stream);
int b200rl_sac_critic_head(const float* z, const float* q1, const float* q2, const float* tq1, const float* tq2,
                           const float* act, const float* rew, const float* is_fir, const double* log_alpha,
                           float* dq1, float* dq2, float* td_out, double* scalars,
                           int B, int T, int A, float gamma, void* stream);
/* continuous SAC                 agents/learner_module/sac_continuous/learning.py:36-131, networks/models.py:199-231
 *   head_out (B,T,2A) = [mu | raw log-std]; eps / act / logp (B,T,A); q* (B,T,1); du* (B,T,f+A) = dLoss/d[obs|act]   */
/* a = tanh(mu + sd*eps), log pi(a).  eps_in given: that noise (tests); null: N(0,1) from Philox4x32-10 under
 * rng_state = {seed, draw number, ticket} (device memory; the kernel advances the draw number, so a replayed CUDA graph
 * draws fresh noise), saved to eps_out for the actor gradient.  act_out / act_in have row stride act_ld floats: the action
 * is written straight into the critic's [obs | act] input (b200rl_sac_cont_pack fills the obs part of the three inputs
 * of one update and the stored action of the third) -- no torch.cat / torch.randn on the update path.                  */
int b200rl_sac_cont_sample(const float* head_out, const float* eps_in, float* eps_out, unsigned long long* rng_state,
                           float* act_out, int act_ld, float* logp_out, int B, int T, int A, void* stream);
int b200rl_sac_cont_pack(const float* obs, const float* act, float* u1, float* u2, float* u3,
                         int B, int T, int f, int A, void* stream);
int b200rl_sac_cont_minq_grad(const float* q1, const float* q2, float* dq1, float* dq2, double* scalars,
                              int B, int T, int A, void* stream);
int b200rl_sac_cont_actor_grad(const float* head_out, const float* eps, const float* act_in, int act_ld, const float* logp,
                               const float* du1, const float* du2, const double* log_alpha,
                               float* dhead, double* scalars, int B, int T, int A, int f, void* stream);
int b200rl_sac_cont_critic_head(const float* tq1, const float* tq2, const float* logp, const float* q1, const float* q2,
                                const float* rew, const float* is_fir, const double* log_alpha,
                                float* dq1, float* dq2, float* td_out, double* scalars,
                                int B, int T, int A, float gamma, void* stream);
/* temperature: state = {log_alpha, adam_m, adam_v} (fp64); loss = exp(log_alpha)*(scalars[2]/denominator + target_entropy);
 * out = {loss_alpha, alpha after the step}                                            agents/learner.py:363-367          */
int b200rl_sac_alpha_step(double* state, const double* scalars, double* out, double denominator,
                          double target_entropy, double lr, int step, void* stream);

/* ---- host -> device batch path (reference agents/learner.py:197-233 + utils/utils.py:101-103) ----------
 * rows of `width_bytes` at stride `src_pitch_bytes` in PAGE-LOCKED host memory -> contiguous device rows, one async 2-D DMA:
 * hx / cx = [:, 0] of the (B, T, H) shared arrays go to the device without a host-side gather.                          */
int b200rl_h2d_rows(void* dst, const void* src, long long width_bytes, long long rows, long long src_pitch_bytes, void* stream);

/* ---- rollout -> batch assembly, storage side (host code) -----------------------------------------------
 * RolloutAssembler.push + make_as_array (buffers/rollout_assembler.py:9-82) fused with LearnerStorage.make_batch
 * (agents/learner_storage.py:123-159): one copy per environment step into a slot-shaped staging block, eight memcpys per
 * finished trajectory into slot `*count` of the shared arrays, `*count += 1`.  Reference semantics kept: active
 * trajectories older than max_age_s are dropped at the next push; a step of an unknown episode continues the first-parked
 * unfinished trajectory (seam marked is_fir = 1); a finished trajectory is dropped when *count == capacity.
 * asm_push returns 1 (trajectory written), 2 (dropped: batch full), 0 (step stored), < 0 (error).
 * stats[6] = {active, parked, emitted, dropped_full, dropped_age, spliced}.                                           */
void* b200rl_asm_create(int seq_len, int obs_width, int action_space, int hidden, double max_age_s);
int b200rl_asm_destroy(void* h);
int b200rl_asm_bind(void* h, float* obs, float* act, float* rew, float* logits, float* log_prob, float* is_fir,
                    float* hx, float* cx, int* count, int capacity);
int b200rl_asm_push(void* h, long long id, int done, double now_s, const float* obs, float act, float rew,
                    const float* logits, float log_prob, float is_fir, const float* hx, const float* cx);
int b200rl_asm_stats(void* h, long long* stats);

#ifdef __cplusplus
}
#endif
#endif /* B200RL_H_ */
