// Cross-translation-unit kernel launchers.
#pragma once
namespace b200rl {
// wgrad_tc.cu: dW_ih, dW_hh, db_ih, db_hh partial slabs on tcgen05 (3xTF32 split)
int launch_wgrad_tc(const float* dpre, const float* h, const float* hx0, const float* u,
                    const float* Wt, const float* bt, float* partials, int nparts, long long P,
                    long long off_Wih, long long off_Whh, long long off_bih, long long off_bhh,
                    int B, int T, int Fin, void* stream);
// dx_tc.cu: torso gradients (dx GEMM on tcgen05 + ReLU mask + reductions) and head gradients
int launch_dx_tc(const float* dpre, const float* u, const float* W_ih, const float* Wt, const float* bt,
                 const float* h, const float* dhead, float* partials, int nparts, long long P,
                 long long off_Wt, long long off_bt, long long off_Whead, long long off_bhead,
                 float* du, int B, int T, int Fin, int NH, void* stream);
// xproj_tc.cu: xproj[b,t,p] = bias + ReLU(torso) . W_ih^T for all rows, gate-interleaved columns
int launch_xproj_tc(const float* u, const float* Wt, const float* bt, const float* W_ih, const float* b_ih, const float* b_hh,
                    float* xproj, int B, int T, int Fin, int sms, void* stream);
// fwd_tc.cu: recurrent LSTM forward with the per-step GEMM on tcgen05; xg = xproj (i,g,f,o) in, saved activations out
int launch_fwd_tc(float* xg, const float* hx0, const float* cx0, const float* W_hh, float* c, float* h,
                  int B, int T, int sms, void* stream);
// fwd_fused.cu: torso + input projection + recurrence (+ heads) in one persistent tcgen05 kernel
int launch_fwd_fused(const float* u, const float* hx0, const float* cx0, const float* Wt, const float* bt, const float* W_ih,
                     const float* W_hh, const float* b_ih, const float* b_hh, const float* W_head, const float* b_head,
                     float* gates, float* c, float* h, float* head_part, float* head_out,
                     int B, int T, int Fin, int NH, int sms, void* stream);
// fwd_fused.cu: the recurrent kernels (fwd_fused, bwd_tc) run 16-sequence tiles when 32-sequence tiles would leave more than half of the SMs idle
bool recurrent_tile16(int B, int sms);
// dx_tc.cu: torso gradients from a stored dx (written by bwd_tc.cu) + head gradients
int launch_torso_from_dx(const float* dx, const float* u, const float* Wt, const float* bt,
                         const float* h, const float* dhead, float* partials, int nparts, long long P,
                         long long off_Wt, long long off_bt, long long off_Whead, long long off_bhead,
                         int B, int T, int Fin, int NH, void* stream);
// bwd_tc.cu: serial BPTT with dh_rec and dx = dpre . W_ih as one tcgen05 accumulation per step
int launch_bwd_tc(const float* dhead, const float* W_head, const float* W_hh, const float* W_ih, float* gates,
                  const float* c, const float* cx0, float* dx, int B, int T, int NH, int sms, void* stream);
// bwd_tc.cu: the same with the head (dW_head, db_head) and torso (dWt, dbt) gradients accumulated in the kernel (one slab row per CTA)
int launch_bwd_tc_fused(const float* dhead, const float* W_head, const float* W_hh, const float* W_ih, float* gates,
                        const float* c, const float* cx0, float* dx, const float* u, const float* Wt, const float* bt,
                        float* partials, long long P, long long off_Wt, long long off_bt, long long off_Whead, long long off_bhead,
                        int B, int T, int Fin, int NH, int sms, void* stream);
}  // namespace b200rl
