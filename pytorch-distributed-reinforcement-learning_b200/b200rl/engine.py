"""The learner update as a fixed sequence of native kernel launches.

One `OnPolicyEngine.update(batch)` is what one epoch of the reference's `learning()`
body does (`agents/learner_module/{ppo,impala}/learning.py:36-106`): forward over the
(B, T) window, GAE / V-trace, loss, backward, clip_grad_norm_, RMSprop step -- with no
Python loop over T, no autograd graph and no host sync:

    fused forward (torso + input projection + recurrence + heads) -> {ppo,impala}_head (scan + loss + d/d heads)
      -> BPTT (+ head / torso gradients) -> weight gradients -> reduce_partials
      -> peer step (gradient exchange over NVLink peer memory + global-norm clip + RMSprop, ONE kernel at every world size)

Multi-GPU: sequences are independent, every loss is a mean over
B*(T-1) elements, so B shards evenly across ranks; each rank computes the gradient of
its shard mean, ONE exchange (sum) of the flat gradient (+ the 8 loss scalars riding
along) and a 1/world scale inside the step kernel reproduce the full-batch gradient; the
global-norm clip comes after the reduction.  Communicators without peer memory (gloo tests,
`B200RL_PEER=0`) fall back to an all-reduce + the single-CTA clip / optimizer kernels.
"""
import math
import os

import torch

from . import comm as _comm
from . import ops as _ops

H = _ops.H


def _make_comm(process_group, comm, device, ops):
    """`comm` wins; a torch.distributed group becomes a PeerGroup (NVLink peer kernels) on CUDA devices, a TorchComm otherwise."""
    if comm is not None or process_group is None:
        return comm
    if _comm.PeerGroup.available(device) and not ops.lib.emulated:
        return _comm.PeerGroup(process_group, device, ops.lib)
    return _comm.TorchComm(process_group)


def _channel(comm, n, device, ops):
    """The fused exchange + clip + optimizer kernel works on a `PeerChannel`: over the communicator's peer mappings when
    there is one, else (single rank) over a private local region -- the same kernel runs at every world size."""
    if comm is not None and hasattr(comm, "channel"):
        return comm.channel(n)
    if comm is None:
        nbytes = _comm.HEADER_BYTES + 2 * ((n * 4 + 64 + 255) // 256 * 256)
        region = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        ch = _comm.PeerChannel(n, 0, 1, device, ops.lib, local_regions=[region.data_ptr()])
        ch._keep = region
        return ch
    return None                     # TorchComm (gloo / NCCL fallback): all_reduce + single-CTA clip/optimizer kernels

_SCALAR_INIT = [0.0, 0.0, 0.0, 0.0, float("inf"), float("-inf"), 0.0, 0.0]


def _sm_count(device, ops):
    if ops.lib.emulated or device.type != "cuda":
        return 4
    return torch.cuda.get_device_properties(device).multi_processor_count


class OnPolicyEngine:
    """PPO / PPO-Continuous / IMPALA update for one `_FlatNet` actor-critic."""

    def __init__(self, net, algo, hyper, ops=None, process_group=None, comm=None):
        self.net, self.algo, self.hp = net, algo, dict(hyper)
        self.ops = ops or _ops.default()
        self.flat = net.flat_parameters()
        self.L = net._layout
        dev = self.flat.device
        self.device = dev
        self.comm = _make_comm(process_group, comm, dev, self.ops)
        self.world = 1 if self.comm is None else self.comm.world
        self.rank = 0 if self.comm is None else self.comm.rank
        self.chan = _channel(self.comm, self.L.P, dev, self.ops)
        self.grad = torch.zeros_like(self.flat)
        self.sq_avg = torch.zeros_like(self.flat)            # RMSprop state (learner.py:70)
        self.scalars = torch.zeros(8, dtype=torch.float64, device=dev)
        self._scalar_init = torch.tensor(_SCALAR_INIT, dtype=torch.float64, device=dev)
        self.grad_norm = torch.zeros(1, device=dev)
        self._shape = None
        self.n_launches = 0
        self.debug_outs = None                                # set to {} to capture intermediates
        self._snap_stream = None                              # side stream of snapshot_async
        self._snap_guard = None                               # event: the last snapshot's copies have read flat / scalars / grad_norm
        # weight-gradient kernel: 1 = tcgen05 (3xTF32 split) for dW/db, 0 = all-FFMA
        self.wgrad_variant = int(os.environ.get("B200RL_WGRAD_VARIANT", "1"))

    # ------------------------------------------------------------------ workspace
    def _workspace(self, B, T):
        if self._shape == (B, T):
            return
        dev, L = self.device, self.L
        self.ws = dict(gates=torch.empty(B, T, H, 4, device=dev), c=torch.empty(B, T, H, device=dev),
                       h=torch.empty(B, T, H, device=dev))
        self.head_out = torch.empty(B, T, L.NH, device=dev)
        self.dhead = torch.empty(B, T, L.NH, device=dev)
        nchunks = (B * T + 31) // 32
        # one partial-gradient slab per CTA of the weight-gradient stage; wgrad_tc runs two CTAs per SM
        self.nparts = max(1, min(2 * _sm_count(dev, self.ops), nchunks))
        self.partials = torch.zeros(self.nparts, L.P, device=dev)   # pad slots stay zero
        self._shape = (B, T)

    def _alloc_outs(self, B, T):
        d, A, dev = self.debug_outs, self.net.n_outputs, self.device
        names = {"PPO": ("td_target", "gae", "ratio"), "IMPALA": ("ratio", "advantages")}["IMPALA" if self.algo == "IMPALA" else "PPO"]
        d.clear()
        d["log_probs"] = torch.empty(B, T, 1, device=dev)
        d["entropy"] = torch.empty(B, T, 1, device=dev)
        if not self.net.continuous:
            d["logits"] = torch.empty(B, T, A, device=dev)
        for n in names:
            d[n] = torch.empty(B, T - 1, 1, device=dev)
        if self.algo == "IMPALA":
            d["values_target"] = torch.empty(B, T, 1, device=dev)
        return d

    # --------------------------------------------------------------------- update
    def update(self, batch):
        """batch: dict of device tensors obs (B,T,f), act (B,T,1), rew, log_prob, is_fir (B,T,1),
        hx, cx (B,H).  Asynchronous: enqueues kernels on the current stream."""
        ops, L, hp = self.ops, self.L, self.hp
        obs = batch["obs"]
        B, T = obs.shape[0], obs.shape[1]
        self._workspace(B, T)
        u = obs.reshape(B, T, -1)
        outs = self._alloc_outs(B, T) if self.debug_outs is not None else None

        # gradient / loss-sum buffers of this step: the peer channel's (other ranks read them over NVLink) or the local ones
        gbuf, sbuf = self.chan.begin() if self.chan is not None else (self.grad, self.scalars)
        if self.chan is None:
            self._wait_snapshot()                             # (`sbuf` IS self.scalars here)
        nfw = ops.net_fwd(L, self.flat, u, batch["hx"], batch["cx"], self.ws, self.head_out)
        sbuf.copy_(self._scalar_init)
        if self.algo == "IMPALA":
            ops.impala_head(self.head_out, batch, self.dhead, sbuf, hp, self.net.n_outputs, outs)
        else:
            ops.ppo_head(self.head_out, batch, self.dhead, sbuf, hp, self.net.n_outputs,
                         self.net.continuous, outs)
        ops.lstm_bwd(L, self.flat, self.dhead, self.ws, batch["cx"], u=u, partials=self.partials)
        ops.lstm_wgrad(L, self.flat, self.ws, batch["hx"], u, self.dhead, self.partials, variant=self.wgrad_variant)
        ops.reduce_partials(self.partials, gbuf)
        scale = 1.0 / self.world
        if self.chan is not None:
            self._wait_snapshot()                             # the step overwrites flat / scalars / grad_norm
            # ONE kernel: peer handshake + one-shot all-reduce over NVLink + global-norm clip + RMSprop (+ the loss sums)
            self.chan.step(0, self.flat, self.sq_avg, None, None, self.grad, self.scalars, self.grad_norm, scale,
                           hp["max_grad_norm"], hp["lr"], 0.99, 0.0, 1e-5, 0, ops._stream(self.device))
        else:
            if self.world > 1:
                self.comm.all_reduce(self.grad)                                 # the ONE gradient collective
            ops.clip_rmsprop(self.flat, self.grad, self.sq_avg, L.P, scale, hp["max_grad_norm"], hp["lr"],
                             0.99, 1e-5, self.grad_norm)
        self.n_launches += nfw + 4 + ({2: 1, 1: 2}.get(self.ws.get("small_grads_done")) or {0: 1, 1: 3, 2: 2}[self.wgrad_variant])   # + head, bwd, reduce, step ; wgrad stage

    # -------------------------------------------------------------------- scalars
    def _reduced_scalars(self):
        s = self.scalars
        if self.world > 1 and self.chan is None:      # (the peer step already reduced them)
            s = s.clone()
            mm = torch.stack([-s[4], s[5]])
            self.comm.all_reduce(mm, op="max")
            sums = s[:4].clone()
            self.comm.all_reduce(sums)
            s = torch.cat([sums, torch.stack([-mm[0], mm[1]]), s[6:]])
        return s

    def _losses_from(self, v, grad_norm):
        B, T = self._shape
        N = float(B * (T - 1) * self.world)
        hp = self.hp
        pol, val, ent = v[0] / N, v[1] / N, v[2] / N
        return {
            "policy-loss": pol, "value-loss": val, "policy-entropy": ent,
            "loss": hp["policy_loss_coef"] * pol + hp["value_loss_coef"] * val - hp["entropy_coef"] * ent,
            "ratio-min": v[4], "ratio-max": v[5], "ratio-avg": v[3] / N,
            "grad-norm": grad_norm,
        }

    def losses(self):
        """Host copy of the loss scalars of the LAST update (synchronises).  Keys follow
        the reference's `detached_losses` (ppo/learning.py:83-88); `ratio` is reduced to
        (min, max, mean) instead of shipping the (B,T-1,1) tensor to the host."""
        out = self._losses_from(self._reduced_scalars().cpu().tolist(), float(self.grad_norm.item()))
        if self.chan is not None:
            self.chan.check()
        return out

    # ---------------------------------------------------------- stream-ordered snapshot
    def _wait_snapshot(self):
        """Order the current stream behind the copies of the last snapshot (enqueued ~one update earlier: never a stall)."""
        if self._snap_guard is not None:
            torch.cuda.current_stream(self.device).wait_event(self._snap_guard)
            self._snap_guard = None

    def snapshot_async(self, with_losses=True):
        """Enqueue, behind the update just issued, the device->pinned-host copies of everything the host consumes per
        iteration (flat parameters for `pub_model`, loss scalars, gradient norm) and return a handle; nothing blocks.
        `snapshot_wait(handle)` later returns (host_flat, losses-or-None).  Two slots alternate, so a handle stays valid
        until the snapshot after the next one is taken.  The copies run on a SIDE stream ordered behind the update (three
        small D2H transfers cost ~25 us of stream time -- 2.5 % of a 1 ms update -- when they sit between two updates);
        the next optimizer step, the only writer of what they read, waits for them (`_wait_snapshot`)."""
        if not hasattr(self, "_snap"):
            pin = self.device.type == "cuda"
            self._snap = [dict(flat=torch.empty(self.flat.shape, dtype=torch.float32, pin_memory=pin),
                               sc=torch.empty(8, dtype=torch.float64, pin_memory=pin),
                               gn=torch.empty(1, dtype=torch.float32, pin_memory=pin)) for _ in range(2)]
            self._snap_i = 0
        self._snap_i ^= 1
        b = self._snap[self._snap_i]
        sc = self._reduced_scalars() if with_losses else None      # (fallback communicators reduce on the compute stream)
        if self.device.type != "cuda":
            b["flat"].copy_(self.flat.detach())
            if with_losses:
                b["sc"].copy_(sc)
                b["gn"].copy_(self.grad_norm)
            return (b, None, with_losses)
        cur = torch.cuda.current_stream(self.device)
        if self._snap_stream is None:
            self._snap_stream = torch.cuda.Stream(self.device)
        done = torch.cuda.Event()
        done.record(cur)
        side = self._snap_stream
        side.wait_event(done)
        with torch.cuda.stream(side):
            b["flat"].copy_(self.flat.detach(), non_blocking=True)
            if with_losses:
                if sc is not self.scalars:
                    sc.record_stream(side)
                b["sc"].copy_(sc, non_blocking=True)
                b["gn"].copy_(self.grad_norm, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
        self._snap_guard = ev
        return (b, ev, with_losses)

    def snapshot_ready(self, handle):
        """True once the snapshot's copies have landed (never blocks)."""
        return handle[1] is None or handle[1].query()

    def snapshot_wait(self, handle):
        b, ev, with_losses = handle
        if ev is not None:
            ev.synchronize()
        return b["flat"], (self._losses_from(b["sc"].tolist(), float(b["gn"][0])) if with_losses else None)


class VMPOEngine(OnPolicyEngine):
    """V-MPO update (reference `v_mpo/learning.py:44-124`, `learner.py:320-348`): same LSTM
    kernels, the V-MPO head (GAE -> per-timestep top-k over the batch -> global softmax ->
    policy / KL gradients) and the two fp64 dual variables with their own RMSprop state.
    `log_eta`, `log_alpha` live on the device as `self.duals` (fp64)."""

    def __init__(self, net, hyper, ops=None, process_group=None, comm=None):
        super().__init__(net, "V-MPO", hyper, ops=ops, process_group=process_group, comm=comm)
        dev = self.device
        init = math.log(self.hp["v_mpo_lagrange_multiplier_init"])
        self.duals = torch.full((2,), init, dtype=torch.float64, device=dev)      # log_eta, log_alpha
        self.dual_sq = torch.zeros(2, dtype=torch.float64, device=dev)
        self.losses_dev = torch.zeros(6, dtype=torch.float64, device=dev)
        self._zero8 = torch.zeros(8, dtype=torch.float64, device=dev)

    def _vm_workspace(self, B, T):
        if getattr(self, "_vm_shape", None) == (B, T):
            return
        dev = self.device
        self.gae = torch.empty(B, T - 1, device=dev)
        self.gae_all = self.gae if self.world == 1 else torch.empty(B * self.world, T - 1, device=dev)
        self.thr = torch.empty(T - 1, device=dev)
        self.tie_cut = torch.empty(T - 1, dtype=torch.int32, device=dev)
        self.colstat = torch.empty(2 * (T - 1), device=dev)
        self._vm_shape = (B, T)

    def update(self, batch, coef_alpha):
        """`coef_alpha`: this epoch's eps_alpha draw (`LearnerSingleVMPO.get_coef_alpha()`)."""
        ops, L, hp, lib = self.ops, self.L, self.hp, self.ops.lib
        obs = batch["obs"]
        B, T = obs.shape[0], obs.shape[1]
        A = self.net.n_outputs
        self._workspace(B, T)
        self._vm_workspace(B, T)
        u = obs.reshape(B, T, -1)
        st = ops._stream(self.device)
        Bg = B * self.world
        N = float(Bg * (T - 1))
        k = math.ceil(Bg / 2)
        from .native import ptr
        ops._chk(batch["act"], batch["rew"], batch["is_fir"], batch["logits"], batch["hx"], batch["cx"])

        self._wait_snapshot()
        gbuf, sbuf = self.chan.begin() if self.chan is not None else (self.grad, self.scalars)
        nfw = ops.net_fwd(L, self.flat, u, batch["hx"], batch["cx"], self.ws, self.head_out)
        sbuf.copy_(self._zero8)
        lib.call("b200rl_vmpo_gae", ptr(self.head_out), ptr(batch["rew"]), ptr(batch["is_fir"]), ptr(self.gae),
                 ptr(self.dhead), ptr(sbuf), B, T, A, hp["gamma"], hp["lmbda"], hp["value_loss_coef"], N, st)
        if self.world > 1:
            self.comm.all_gather(self.gae_all, self.gae)       # the extra V-MPO exchange: top-k is over the WHOLE batch
        lib.call("b200rl_vmpo_select", ptr(self.gae_all), Bg, T - 1, k, ptr(self.duals), ptr(self.thr), ptr(self.colstat),
                 ptr(self.tie_cut), st)
        lib.call("b200rl_vmpo_grad", ptr(self.head_out), ptr(batch["act"]), ptr(batch["logits"]), ptr(self.gae), ptr(self.thr),
                 ptr(self.colstat), ptr(self.tie_cut), ptr(self.duals), ptr(self.dhead), ptr(sbuf), B, T, A, self.rank * B,
                 hp["policy_loss_coef"], N, st)
        ops.lstm_bwd(L, self.flat, self.dhead, self.ws, batch["cx"], u=u, partials=self.partials)
        ops.lstm_wgrad(L, self.flat, self.ws, batch["hx"], u, self.dhead, self.partials, variant=self.wgrad_variant)
        ops.reduce_partials(self.partials, gbuf)
        # (the kernels already normalise by the GLOBAL N: plain sum over ranks, no 1/world)
        if self.chan is not None:
            self.chan.step(0, self.flat, self.sq_avg, None, None, self.grad, self.scalars, self.grad_norm, 1.0,
                           hp["max_grad_norm"], hp["lr"], 0.99, 0.0, 1e-5, 0, st)
        else:
            if self.world > 1:
                self.comm.all_reduce(self.grad)
                self.comm.all_reduce(self.scalars[:4])
            ops.clip_rmsprop(self.flat, self.grad, self.sq_avg, L.P, 1.0, hp["max_grad_norm"], hp["lr"], 0.99, 1e-5, self.grad_norm)
        lib.call("b200rl_vmpo_dual_step", ptr(self.duals), ptr(self.dual_sq), ptr(self.scalars), ptr(self.losses_dev),
                 hp["coef_eta"], float(coef_alpha), hp["lr"], 0.99, 1e-5, N, float(k * (T - 1)),
                 hp["policy_loss_coef"], hp["value_loss_coef"], st)
        self.n_launches += nfw + 7 + ({2: 1, 1: 2}.get(self.ws.get("small_grads_done")) or {0: 1, 1: 3, 2: 2}[self.wgrad_variant])

    def losses(self):
        v = self.losses_dev.cpu().tolist()
        d = self.duals.cpu().tolist()
        if self.chan is not None:
            self.chan.check()
        return {"policy-loss": v[0], "value-loss": v[1], "loss-temperature": v[2], "loss-alpha": v[3], "loss": v[4],
                "kl-mean": v[5], "log_eta": d[0], "log_alpha": d[1], "selected": float(self.scalars[3].item()),
                "grad-norm": float(self.grad_norm.item())}


class _GraphCache:
    """CUDA graphs of a launch sequence, keyed by the device addresses it reads and writes (launch-bound shapes: one SAC-
    Continuous update is ~70 tiny kernels, ~20 us of host time each through the C-ABI).  The first call with a key runs
    eagerly (lazy one-time setup -- workspaces, function attributes -- stays out of the capture), the second captures,
    later ones replay.  Anything whose arguments change from step to step (the optimizer steps: bias corrections, peer
    epochs) is launched outside.  `limit` bounds the number of graphs; further keys run eagerly."""

    def __init__(self, limit=12):
        self.graphs, self.seen, self.limit = {}, set(), limit
        self.replays = 0

    def run(self, key, fn):
        g = self.graphs.get(key)
        if g is not None:
            g.replay()
            self.replays += 1
            return
        if key not in self.seen or len(self.graphs) >= self.limit:
            self.seen.add(key)
            fn()
            return
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            fn()
        self.graphs[key] = g
        g.replay()
        self.replays += 1


class SACEngine:
    """SAC / SAC-Continuous update (reference `sac/learning.py:36-143`,
    `sac_continuous/learning.py:36-131`, optimizers `learner.py:351-367`): actor step, temperature
    step, twin-critic step and the Polyak update, built from the shared LSTM kernels plus the
    element-wise SAC heads.  By default the target critic ALIASES the critic exactly as the
    reference does (`target_critic = model.critic.to(device)` is the same object, learner.py:358;
    SURVEY.md A.6-8), so the "target" forward is the critic forward and `soft_update` is
    theta <- (1-tau) theta + tau theta; `separate_target=True` keeps a real copy instead."""

    def __init__(self, model, hyper, continuous, ops=None, separate_target=False, tau=0.005, process_group=None, comm=None):
        self.ops = ops or _ops.default()
        self.hp = dict(hyper)
        self.cont = continuous
        self.actor, self.twin = model.actor, model.critic
        self.q = [self.twin.q1, self.twin.q2]
        self.fa = self.actor.flat_parameters()
        self.fq = self.twin.flat_joint()
        self.La, self.Lq = self.actor._layout, self.q[0]._layout
        dev = self.fa.device
        self.device = dev
        # data parallel: the replay batch is sharded over ranks (equal shards); every loss is a mean over the batch, so the
        # actor / twin-critic gradients and the temperature statistics are summed over ranks and scaled by 1/world.  Two
        # DEPENDENT exchanges per update (the critic target uses the actor AFTER its step, sac_continuous/learning.py:50-71),
        # each one fused peer kernel (exchange + clip + Adam; the loss / temperature sums ride along)
        self.comm = _make_comm(process_group, comm, dev, self.ops)
        self.world = 1 if self.comm is None else self.comm.world
        self.chan_a = _channel(self.comm, self.La.P, dev, self.ops)
        self.chan_q = _channel(self.comm, 2 * self.Lq.P, dev, self.ops)
        self.sc_a = torch.zeros(8, dtype=torch.float64, device=dev)      # reduced sums of the actor phase / the critic phase
        self.sc_q = torch.zeros(8, dtype=torch.float64, device=dev)
        z = lambda t: torch.zeros_like(t)
        self.ga, self.ma, self.va = z(self.fa), z(self.fa), z(self.fa)
        self.gq, self.mq, self.vq = z(self.fq), z(self.fq), z(self.fq)
        self.separate_target = separate_target
        self.ftq = self.fq.clone() if separate_target else self.fq
        self.tau = tau
        m = self.q[0].grad_mask()
        self.qmask = None if m is None else torch.cat([m, self.q[1].grad_mask()])
        self.alpha_state = torch.tensor([math.log(self.hp["alpha"]), 0.0, 0.0], dtype=torch.float64, device=dev)
        self.alpha_out = torch.zeros(2, dtype=torch.float64, device=dev)
        self.scalars = torch.zeros(8, dtype=torch.float64, device=dev)
        self.norm_a, self.norm_q = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
        self.step = 0
        self._shape = None
        self.n_launches = 0
        self.wgrad_variant = int(os.environ.get("B200RL_WGRAD_VARIANT", "1"))
        self.debug = None
        self.noise_fn = None          # SAC-Continuous: callable (B, T, A) -> (eps1, eps2) replacing the two N(0,1) draws (tests)
        # Philox state of the in-kernel N(0,1) draws: {seed, draw number, ticket}; every rank draws its own stream
        seed = (int(torch.initial_seed()) + 0x9E3779B97F4A7C15 * (1 + (0 if self.comm is None else self.comm.rank))) & 0x7FFFFFFFFFFFFFFF
        self.rng = torch.tensor([seed, 0, 0, 0], dtype=torch.int64, device=dev)
        # CUDA graphs of the two launch-bound phases (B200RL_GRAPH=1 forces them on, 0 off; default: on for small batches)
        self.graph_mode = os.environ.get("B200RL_GRAPH", "auto")
        loopback = type(self.comm).__name__ == "_LoopbackComm"      # (test harness: ranks are threads of one process -- no concurrent captures)
        self.graphs = _GraphCache(limit=16) if (dev.type == "cuda" and self.graph_mode != "0" and not loopback) else None

    @property
    def log_alpha(self):
        return self.alpha_state[0]

    def _ws(self, B, T, Ad):
        if self._shape == (B, T, Ad):
            return
        dev = self.device
        mk = lambda: dict(gates=torch.empty(B, T, H, 4, device=dev), c=torch.empty(B, T, H, device=dev), h=torch.empty(B, T, H, device=dev))
        self.wsa, self.wsq = mk(), [mk(), mk()]
        nch = (B * T + 31) // 32
        self.nparts = max(1, min(_sm_count(dev, self.ops), nch))
        self.pa = torch.zeros(self.nparts, self.La.P, device=dev)
        self.pq = [torch.zeros(self.nparts, self.Lq.P, device=dev) for _ in range(2)]
        self.hoa = torch.empty(B, T, self.La.NH, device=dev)
        self.dhoa = torch.empty_like(self.hoa)
        NQ = self.Lq.NH
        self.Q = [torch.empty(B, T, NQ, device=dev) for _ in range(2)]
        self.TQ = [torch.empty(B, T, NQ, device=dev) for _ in range(2)] if (self.separate_target or self.cont) else self.Q
        self.dQ = [torch.empty(B, T, NQ, device=dev) for _ in range(2)]
        self.td = torch.empty(B, T - 1, device=dev)
        if self.cont:
            Fq = self.Lq.Fin
            self.ub = [torch.empty(B, T, Fq, device=dev) for _ in range(3)]     # [obs | action]: policy action (actor phase), fresh policy action, stored action
            self.eps = torch.empty(B, T, Ad, device=dev)
            self.logp = torch.empty(B, T, Ad, device=dev)
            self.du = [torch.empty(B, T, Fq, device=dev) for _ in range(2)]
            self.scratch = torch.zeros(self.nparts, self.Lq.P, device=dev)
        self._shape = (B, T, Ad)

    def _net_fwd(self, L, flat, off, u, hx, cx, ws, out):
        f = flat[off:off + L.P]
        self.ops.net_fwd(L, f, u, hx, cx, ws, out)

    def _net_bwd(self, L, flat, off, u, hx, cx, ws, dhead, partials, grad, du=None):
        f = flat[off:off + L.P]
        self.ops.lstm_bwd(L, f, dhead, ws, cx, u=None if du is not None else u, partials=None if du is not None else partials)
        self.ops.lstm_wgrad(L, f, ws, hx, u, dhead, partials, du=du, variant=self.wgrad_variant)
        if grad is not None:
            self.ops.reduce_partials(partials, grad)

    GRAPH_MAX_ROWS = 1 << 16          # B200RL_GRAPH=auto: phases of batches up to this many transitions run as CUDA graphs

    def update(self, batch, noise=None):
        ops, hp, lib = self.ops, self.hp, self.ops.lib
        from .native import ptr
        obs = batch["obs"]
        B, T = obs.shape[0], obs.shape[1]
        f = obs.reshape(B, T, -1).shape[-1]
        act = batch["act"]
        Ad = act.shape[-1] if self.cont else self.La.NH
        self._ws(B, T, Ad)
        ops._chk(act, batch["rew"], batch["is_fir"], batch["hx"], batch["cx"])
        st_of = lambda: ops._stream(self.device)          # (inside a capture the current stream is the capture stream)
        hx, cx = batch["hx"], batch["cx"]
        uo = obs.reshape(B, T, -1)
        La, Lq, Pq = self.La, self.Lq, self.Lq.P
        N = float(B * (T - 1))
        self.step += 1
        lr, mg = hp["lr"], hp["max_grad_norm"]
        # gradient / loss-sum buffers of this update: the peer channels' (read by the other ranks over NVLink) or local ones
        ga, sA = self.chan_a.begin() if self.chan_a is not None else (self.ga, self.scalars)
        gq, sQ = self.chan_q.begin() if self.chan_q is not None else (self.gq, self.scalars)
        A = Ad
        eps1 = eps2 = None
        if self.cont:
            if noise is None and self.noise_fn is not None:
                noise = self.noise_fn(B, T, A)
            if noise is not None:
                eps1, eps2 = noise
                ops._chk(eps1, eps2)
            Fq = f + A
            ua = [ptr(u) + 4 * f for u in self.ub]        # where the action columns of the three critic inputs start

        def actor_phase():
            st = st_of()
            sA.zero_()
            if sQ is not sA:
                sQ.zero_()
            if not self.cont:
                self._net_fwd(La, self.fa, 0, uo, hx, cx, self.wsa, self.hoa)
                for i in range(2):
                    self._net_fwd(Lq, self.fq, i * Pq, uo, hx, cx, self.wsq[i], self.Q[i])
                lib.call("b200rl_sac_actor_head", ptr(self.hoa), ptr(self.Q[0]), ptr(self.Q[1]), ptr(self.alpha_state),
                         ptr(self.dhoa), ptr(sA), B, T, A, st)
            else:
                lib.call("b200rl_sac_cont_pack", ptr(uo), ptr(act), ptr(self.ub[0]), ptr(self.ub[1]), ptr(self.ub[2]), B, T, f, A, st)
                self._net_fwd(La, self.fa, 0, uo, hx, cx, self.wsa, self.hoa)
                lib.call("b200rl_sac_cont_sample", ptr(self.hoa), ptr(eps1), ptr(self.eps), ptr(self.rng), ua[0], Fq, ptr(self.logp), B, T, A, st)
                for i in range(2):
                    self._net_fwd(Lq, self.fq, i * Pq, self.ub[0], hx, cx, self.wsq[i], self.Q[i])
                lib.call("b200rl_sac_cont_minq_grad", ptr(self.Q[0]), ptr(self.Q[1]), ptr(self.dQ[0]), ptr(self.dQ[1]), ptr(sA), B, T, A, st)
                for i in range(2):                                                      # dQ/da: critic backward to its input only
                    self._net_bwd(Lq, self.fq, i * Pq, self.ub[0], hx, cx, self.wsq[i], self.dQ[i], self.scratch, None, du=self.du[i])
                lib.call("b200rl_sac_cont_actor_grad", ptr(self.hoa), ptr(eps1 if eps1 is not None else self.eps), ua[0], Fq, ptr(self.logp),
                         ptr(self.du[0]), ptr(self.du[1]), ptr(self.alpha_state), ptr(self.dhoa), ptr(sA), B, T, A, f, st)
            self._net_bwd(La, self.fa, 0, uo, hx, cx, self.wsa, self.dhoa, self.pa, ga)

        def critic_phase():
            st = st_of()
            self._net_fwd(La, self.fa, 0, uo, hx, cx, self.wsa, self.hoa)            # updated actor (continuous: fresh noise)
            if not self.cont:
                if self.separate_target:
                    for i in range(2):
                        self._net_fwd(Lq, self.ftq, i * Pq, uo, hx, cx, self.wsa, self.TQ[i])
                lib.call("b200rl_sac_critic_head", ptr(self.hoa), ptr(self.Q[0]), ptr(self.Q[1]), ptr(self.TQ[0]), ptr(self.TQ[1]),
                         ptr(act), ptr(batch["rew"]), ptr(batch["is_fir"]), ptr(self.alpha_state), ptr(self.dQ[0]), ptr(self.dQ[1]),
                         ptr(self.td), ptr(sQ), B, T, A, hp["gamma"], st)
                uq = uo
            else:
                lib.call("b200rl_sac_cont_sample", ptr(self.hoa), ptr(eps2), None, ptr(self.rng), ua[1], Fq, ptr(self.logp), B, T, A, st)
                for i in range(2):
                    self._net_fwd(Lq, self.ftq, i * Pq, self.ub[1], hx, cx, self.wsq[i], self.TQ[i])
                for i in range(2):
                    self._net_fwd(Lq, self.fq, i * Pq, self.ub[2], hx, cx, self.wsq[i], self.Q[i])
                lib.call("b200rl_sac_cont_critic_head", ptr(self.TQ[0]), ptr(self.TQ[1]), ptr(self.logp), ptr(self.Q[0]), ptr(self.Q[1]),
                         ptr(batch["rew"]), ptr(batch["is_fir"]), ptr(self.alpha_state), ptr(self.dQ[0]), ptr(self.dQ[1]), ptr(self.td),
                         ptr(sQ), B, T, A, hp["gamma"], st)
                uq = self.ub[2]
            for i in range(2):
                self._net_bwd(Lq, self.fq, i * Pq, uq, hx, cx, self.wsq[i], self.dQ[i], self.pq[i], gq[i * Pq:(i + 1) * Pq])

        use_graph = self.graphs is not None and (self.graph_mode == "1" or B * T <= self.GRAPH_MAX_ROWS)
        if use_graph:
            key = tuple(t.data_ptr() for t in (obs, act, batch["rew"], batch["is_fir"], hx, cx, ga, sA, gq, sQ)) + \
                (B, T, 0 if eps1 is None else eps1.data_ptr(), 0 if eps2 is None else eps2.data_ptr())
            run = lambda tag, fn: self.graphs.run((tag,) + key, fn)
        else:
            run = lambda tag, fn: fn()

        st = st_of()
        run("actor", actor_phase)
        sc = self._actor_step(ga, sA, mg, lr, st)
        pol_den = N * A if self.cont else N
        lib.call("b200rl_sac_alpha_step", ptr(self.alpha_state), ptr(sc), ptr(self.alpha_out), pol_den * self.world, float(A), lr, self.step, st)
        run("critic", critic_phase)
        # twin-critic step: exchange #2 + joint clip + Adam
        if self.chan_q is not None:
            self.chan_q.step(1, self.fq, self.mq, self.vq, self.qmask, self.gq, self.sc_q, self.norm_q, 1.0 / self.world, mg, lr,
                             0.9, 0.999, 1e-8, self.step, st)
        else:
            if self.world > 1:
                self.comm.all_reduce(self.gq)
            ops.clip_adam(self.fq, self.gq, self.mq, self.vq, 2 * Pq, 1.0 / self.world, mg, lr, self.step, norm_out=self.norm_q, mask=self.qmask)
        ops.soft_update_flat(self.fq, self.ftq, self.tau)
        self._pol_den, self._N = pol_den * self.world, N * self.world
        self.n_launches += 1

    def _actor_step(self, ga, sA, mg, lr, st):
        """Exchange #1 + clip + Adam on the actor; returns the rank-summed loss / temperature statistics of the actor phase
        (what the temperature step reads)."""
        if self.chan_a is not None:
            self.chan_a.step(1, self.fa, self.ma, self.va, None, self.ga, self.sc_a, self.norm_a, 1.0 / self.world, mg, lr,
                             0.9, 0.999, 1e-8, self.step, st)
            return self.sc_a
        sc = sA
        if self.world > 1:
            self.comm.all_reduce(self.ga)
            sc = sA.clone()                      # (`self.scalars` keeps accumulating local sums for the critic loss)
            self.comm.all_reduce(sc)
        self.ops.clip_adam(self.fa, self.ga, self.ma, self.va, self.La.P, 1.0 / self.world, mg, lr, self.step, norm_out=self.norm_a)
        return sc

    def losses(self):
        if self.chan_a is not None:
            sa, sq = self.sc_a.cpu().tolist(), self.sc_q.cpu().tolist()
            s = [sa[0], sq[1]]
            self.chan_a.check()
            self.chan_q.check()
        else:
            sc = self.scalars
            if self.world > 1:
                sc = sc.clone()
                self.comm.all_reduce(sc)
            s = sc.cpu().tolist()
        ao = self.alpha_out.cpu().tolist()
        pol, val = s[0] / self._pol_den, s[1] / self._N
        return {"policy-loss": pol, "value-loss": val, "loss_alpha": ao[0], "alpha": ao[1],
                "loss": self.hp["policy_loss_coef"] * pol + self.hp["value_loss_coef"] * val,
                "log_alpha": float(self.alpha_state[0].item())}
