"""The learner update as a fixed sequence of native kernel launches.

One `OnPolicyEngine.update(batch)` is what one epoch of the reference's `learning()`
body does (`agents/learner_module/{ppo,impala}/learning.py:36-106`): forward over the
(B, T) window, GAE / V-trace, loss, backward, clip_grad_norm_, RMSprop step -- with no
Python loop over T, no autograd graph and no host sync:

    lstm_fwd -> head_fwd -> {ppo,impala}_head -> lstm_bwd -> lstm_wgrad
             -> reduce_partials -> [NCCL all-reduce of the flat gradient] -> clip_rmsprop

Multi-GPU: sequences are independent, every loss is a mean over
B*(T-1) elements, so B shards evenly across ranks; each rank computes the gradient of
its shard mean, ONE all-reduce(sum) of the flat gradient (+ the 8 loss scalars riding
in a second tiny buffer) and a 1/world scale inside the optimizer kernel reproduce the
full-batch gradient; the global-norm clip comes after the reduction.
"""
import math
import os

import torch

from . import ops as _ops

H = _ops.H

_SCALAR_INIT = [0.0, 0.0, 0.0, 0.0, float("inf"), float("-inf"), 0.0, 0.0]


def _sm_count(device, ops):
    if ops.lib.emulated or device.type != "cuda":
        return 4
    return torch.cuda.get_device_properties(device).multi_processor_count


class OnPolicyEngine:
    """PPO / PPO-Continuous / IMPALA update for one `_FlatNet` actor-critic."""

    def __init__(self, net, algo, hyper, ops=None, process_group=None):
        self.net, self.algo, self.hp = net, algo, dict(hyper)
        self.ops = ops or _ops.default()
        self.pg = process_group
        self.world = 1 if process_group is None else torch.distributed.get_world_size(process_group)
        self.flat = net.flat_parameters()
        self.L = net._layout
        dev = self.flat.device
        self.device = dev
        self.grad = torch.zeros_like(self.flat)
        self.sq_avg = torch.zeros_like(self.flat)            # RMSprop state (learner.py:70)
        self.scalars = torch.zeros(8, dtype=torch.float64, device=dev)
        self._scalar_init = torch.tensor(_SCALAR_INIT, dtype=torch.float64, device=dev)
        self.grad_norm = torch.zeros(1, device=dev)
        self._shape = None
        self.n_launches = 0
        self.debug_outs = None                                # set to {} to capture intermediates
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
        self.nparts = max(1, min(_sm_count(dev, self.ops), nchunks))
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

        ops.lstm_fwd(L, self.flat, u, batch["hx"], batch["cx"], self.ws)
        ops.head_fwd(L, self.flat, self.ws["h"], self.head_out)
        self.scalars.copy_(self._scalar_init)
        if self.algo == "IMPALA":
            ops.impala_head(self.head_out, batch, self.dhead, self.scalars, hp, self.net.n_outputs, outs)
        else:
            ops.ppo_head(self.head_out, batch, self.dhead, self.scalars, hp, self.net.n_outputs,
                         self.net.continuous, outs)
        ops.lstm_bwd(L, self.flat, self.dhead, self.ws, batch["cx"])
        ops.lstm_wgrad(L, self.flat, self.ws, batch["hx"], u, self.dhead, self.partials, variant=self.wgrad_variant)
        ops.reduce_partials(self.partials, self.grad)
        scale = 1.0
        if self.world > 1:
            torch.distributed.all_reduce(self.grad, group=self.pg)          # the ONE gradient collective
            scale = 1.0 / self.world
        ops.clip_rmsprop(self.flat, self.grad, self.sq_avg, L.P, scale, hp["max_grad_norm"], hp["lr"],
                         0.99, 1e-5, self.grad_norm)
        self.n_launches += 8 + {0: 0, 1: 2, 2: 1}[self.wgrad_variant]

    # -------------------------------------------------------------------- scalars
    def losses(self):
        """Host copy of the loss scalars of the LAST update (synchronises).  Keys follow
        the reference's `detached_losses` (ppo/learning.py:83-88); `ratio` is reduced to
        (min, max, mean) instead of shipping the (B,T-1,1) tensor to the host."""
        B, T = self._shape
        s = self.scalars
        if self.world > 1:
            s = s.clone()
            mm = torch.stack([-s[4], s[5]])
            torch.distributed.all_reduce(mm, op=torch.distributed.ReduceOp.MAX, group=self.pg)
            sums = s[:4].clone()
            torch.distributed.all_reduce(sums, group=self.pg)
            s = torch.cat([sums, torch.stack([-mm[0], mm[1]]), s[6:]])
        v = s.cpu().tolist()
        N = float(B * (T - 1) * self.world)
        hp = self.hp
        pol, val, ent = v[0] / N, v[1] / N, v[2] / N
        return {
            "policy-loss": pol, "value-loss": val, "policy-entropy": ent,
            "loss": hp["policy_loss_coef"] * pol + hp["value_loss_coef"] * val - hp["entropy_coef"] * ent,
            "ratio-min": v[4], "ratio-max": v[5], "ratio-avg": v[3] / N,
            "grad-norm": float(self.grad_norm.item()),
        }


class VMPOEngine(OnPolicyEngine):
    """V-MPO update (reference `v_mpo/learning.py:44-124`, `learner.py:320-348`): same LSTM
    kernels, the V-MPO head (GAE -> per-timestep top-k over the batch -> global softmax ->
    policy / KL gradients) and the two fp64 dual variables with their own RMSprop state.
    `log_eta`, `log_alpha` live on the device as `self.duals` (fp64)."""

    def __init__(self, net, hyper, ops=None, process_group=None):
        super().__init__(net, "V-MPO", hyper, ops=ops, process_group=process_group)
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

        ops.lstm_fwd(L, self.flat, u, batch["hx"], batch["cx"], self.ws)
        ops.head_fwd(L, self.flat, self.ws["h"], self.head_out)
        self.scalars.copy_(self._zero8)
        lib.call("b200rl_vmpo_gae", ptr(self.head_out), ptr(batch["rew"]), ptr(batch["is_fir"]), ptr(self.gae),
                 ptr(self.dhead), ptr(self.scalars), B, T, A, hp["gamma"], hp["lmbda"], hp["value_loss_coef"], N, st)
        if self.world > 1:
            torch.distributed.all_gather_into_tensor(self.gae_all, self.gae, group=self.pg)
        lib.call("b200rl_vmpo_select", ptr(self.gae_all), Bg, T - 1, k, ptr(self.duals), ptr(self.thr), ptr(self.colstat), st)
        lib.call("b200rl_vmpo_grad", ptr(self.head_out), ptr(batch["act"]), ptr(batch["logits"]), ptr(self.gae), ptr(self.thr),
                 ptr(self.colstat), ptr(self.duals), ptr(self.dhead), ptr(self.scalars), B, T, A, hp["policy_loss_coef"], N, st)
        ops.lstm_bwd(L, self.flat, self.dhead, self.ws, batch["cx"])
        ops.lstm_wgrad(L, self.flat, self.ws, batch["hx"], u, self.dhead, self.partials, variant=self.wgrad_variant)
        ops.reduce_partials(self.partials, self.grad)
        if self.world > 1:                      # kernels already normalise by the GLOBAL N: plain sum, no 1/world
            torch.distributed.all_reduce(self.grad, group=self.pg)
            torch.distributed.all_reduce(self.scalars[:4], group=self.pg)
        ops.clip_rmsprop(self.flat, self.grad, self.sq_avg, L.P, 1.0, hp["max_grad_norm"], hp["lr"], 0.99, 1e-5, self.grad_norm)
        lib.call("b200rl_vmpo_dual_step", ptr(self.duals), ptr(self.dual_sq), ptr(self.scalars), ptr(self.losses_dev),
                 hp["coef_eta"], float(coef_alpha), hp["lr"], 0.99, 1e-5, N, float(k * (T - 1)),
                 hp["policy_loss_coef"], hp["value_loss_coef"], st)
        self.n_launches += 12 + {0: 0, 1: 2, 2: 1}[self.wgrad_variant]

    def losses(self):
        v = self.losses_dev.cpu().tolist()
        d = self.duals.cpu().tolist()
        return {"policy-loss": v[0], "value-loss": v[1], "loss-temperature": v[2], "loss-alpha": v[3], "loss": v[4],
                "kl-mean": v[5], "log_eta": d[0], "log_alpha": d[1], "selected": float(self.scalars[3].item()),
                "grad-norm": float(self.grad_norm.item())}
