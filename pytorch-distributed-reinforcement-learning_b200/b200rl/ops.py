"""Tensor-level wrappers over the C-ABI (one Python call -> one native entry point).

torch is used here only for device memory and the current CUDA stream.
"""
import os

import torch

from . import native
from .native import ptr

H = 64
# kernel variants (env-overridable for A/B measurements and the cross-implementation tests):
# forward 5 = ONE persistent tcgen05 kernel for torso + input projection + recurrence + heads (csrc/fwd_fused.cu; nets with
# Fin > 8 or more than 4 head outputs fall back to 3), 3 = tcgen05 input projection (xproj_tc.cu) + tcgen05 recurrent kernel
# (fwd_tc.cu) + head_fwd, 0 = the independently written FFMA kernel (the reference the tensor-core paths are checked against)
FWD_VARIANT = int(os.environ.get("B200RL_FWD_VARIANT", "5"))
# backward 2 = tcgen05 BPTT kernel that also accumulates the head and torso gradients in registers (no dx in HBM, no head_grad /
# torso_grad passes; nets with Fin != 4 or more than 4 head outputs, or callers that need d/d(input), fall back to 1),
# 3 = as 2 but only the head gradients are fused (dx is stored, the torso pass stays),
# 1 = tcgen05 BPTT kernel (emits dx = dpre . W_ih, so the weight-gradient stage skips that GEMM), 0 = FFMA kernel
BWD_VARIANT = int(os.environ.get("B200RL_BWD_VARIANT", "2"))


class Ops:
    def __init__(self, lib=None, allow_cpu=False):
        self._lib = lib
        self.allow_cpu = allow_cpu            # test-only (host-emulated library)

    @property
    def lib(self):
        if self._lib is None:
            self._lib = native.get()
        return self._lib

    # ------------------------------------------------------------------ helpers
    def _chk(self, *ts):
        dev = None
        for t in ts:
            if t is None:
                continue
            if t.dtype not in (torch.float32, torch.float64):
                raise TypeError(f"b200rl: expected float tensor, got {t.dtype}")
            if not t.is_contiguous():
                raise ValueError("b200rl: tensor must be contiguous")
            if not t.is_cuda and not self.allow_cpu:
                raise native.NativeError("b200rl: tensors must live on a CUDA device (no CPU fallback)")
            dev = t.device if dev is None else dev
            if t.device != dev:
                raise ValueError("b200rl: tensors on different devices")
        return dev

    def _stream(self, dev):
        return native.stream_ptr(dev) if dev is not None and dev.type == "cuda" else 0

    @staticmethod
    def _c(t):
        return t if t.is_contiguous() else t.contiguous()

    # ------------------------------------------------------------- recurrences
    def compute_gae(self, deltas, gamma, lambda_):
        d = self._c(deltas.float())
        dev = self._chk(d)
        B, n = d.shape[0], d.shape[1]
        out = torch.empty_like(d)
        self.lib.call("b200rl_compute_gae", ptr(d), ptr(out), B, n, gamma, lambda_, self._stream(dev))
        return out

    def td_gae(self, rew, is_fir, value, gamma, lambda_):
        rew, is_fir, value = self._c(rew), self._c(is_fir), self._c(value)
        dev = self._chk(rew, is_fir, value)
        B, T = value.shape[0], value.shape[1]
        td = torch.empty(B, T - 1, 1, device=value.device, dtype=torch.float32)
        gae = torch.empty_like(td)
        self.lib.call("b200rl_td_gae", ptr(rew), ptr(is_fir), ptr(value), ptr(td), ptr(gae), B, T,
                      gamma, lambda_, self._stream(dev))
        return td, gae

    def compute_v_trace(self, behav_log_probs, target_log_probs, is_fir, rewards, values, gamma,
                        rho_bar=0.8, c_bar=1.0):
        a = [self._c(x) for x in (behav_log_probs, target_log_probs, is_fir, rewards, values)]
        dev = self._chk(*a)
        B, T = a[4].shape[0], a[4].shape[1]
        rho = torch.empty(B, T - 1, 1, device=a[4].device, dtype=torch.float32)
        adv = torch.empty_like(rho)
        vs = torch.empty_like(a[4])
        self.lib.call("b200rl_compute_v_trace", *[ptr(x) for x in a], ptr(rho), ptr(adv), ptr(vs), B, T,
                      gamma, rho_bar, c_bar, self._stream(dev))
        return rho, adv, vs

    def kldivergence(self, logits_p, logits_q):
        p, q = self._c(logits_p), self._c(logits_q)
        dev = self._chk(p, q)
        A = p.shape[-1]
        rows = p.numel() // A
        out = torch.empty(p.shape[:-1], device=p.device, dtype=torch.float32)
        self.lib.call("b200rl_kldivergence", ptr(p), ptr(q), ptr(out), rows, A, self._stream(dev))
        return out.squeeze()

    def soft_update_flat(self, src, target, tau=0.005):
        dev = self._chk(src, target)
        self.lib.call("b200rl_soft_update", ptr(src), ptr(target), src.numel(), tau, self._stream(dev))


_default = None


def default():
    global _default
    if _default is None:
        _default = Ops()
    return _default


# ===================================================================== LSTM network ops
class NetLayout:
    """Flat fp32 parameter layout of one torso+LSTM+heads network:
    [Wt (H,Fin) | bt (H) | W_ih (4H,H) | W_hh (4H,H) | b_ih (4H) | b_hh (4H) | W_head (NH,H) | b_head (NH)]
    padded to a multiple of 4 floats (16-byte rows for the gradient slabs)."""

    def __init__(self, Fin, NH):
        self.Fin, self.NH = Fin, NH
        o = 0
        self.off = {}
        for name, n in (("Wt", H * Fin), ("bt", H), ("W_ih", 4 * H * H), ("W_hh", 4 * H * H),
                        ("b_ih", 4 * H), ("b_hh", 4 * H), ("W_head", NH * H), ("b_head", NH)):
            self.off[name] = (o, n)
            o += n
        self.n = o
        self.P = (o + 3) // 4 * 4

    def view(self, flat, name):
        o, n = self.off[name]
        return flat[o:o + n]


def _lstm_fwd(self, L, flat, u, hx0, cx0, ws, variant=None):
    """ws: workspace dict with gates (B,T,H,4), c, h (B,T,H)."""
    dev = self._chk(flat, u, hx0, cx0, ws["gates"], ws["c"], ws["h"])
    B, T, Fin = u.shape
    assert Fin == L.Fin
    v = lambda k: ptr(flat) + 4 * L.off[k][0]
    variant = FWD_VARIANT if variant is None else variant
    if variant == 5:
        if Fin <= 8:
            self.lib.call("b200rl_lstm_fwd_fused", ptr(u), ptr(hx0), ptr(cx0), v("Wt"), v("bt"), v("W_ih"), v("W_hh"),
                          v("b_ih"), v("b_hh"), 0, 0, ptr(ws["gates"]), ptr(ws["c"]), ptr(ws["h"]), 0, 0, B, T, Fin, L.NH,
                          self._stream(dev))
            return
        variant = 3
    self.lib.call("b200rl_lstm_fwd", ptr(u), ptr(hx0), ptr(cx0), v("Wt"), v("bt"), v("W_ih"), v("W_hh"),
                  v("b_ih"), v("b_hh"), ptr(ws["gates"]), ptr(ws["c"]), ptr(ws["h"]), B, T, Fin, variant, self._stream(dev))


def _net_fwd(self, L, flat, u, hx0, cx0, ws, head_out, variant=None):
    """Network forward: saved activations into `ws` and head_out (B, T, NH) = h . W_head^T + b_head.  Variant 5 with
    Fin <= 8 and NH <= 4 is ONE recurrent kernel (its CTAs fold their own partial heads at the end); everything else is lstm_fwd + head_fwd.
    Returns the number of kernel launches."""
    variant = FWD_VARIANT if variant is None else variant
    B, T, Fin = u.shape
    if variant == 5 and Fin <= 8 and L.NH <= 4:
        dev = self._chk(flat, u, hx0, cx0, ws["gates"], ws["c"], ws["h"], head_out)
        hp = ws.get("head_part")
        if hp is None or hp.shape != (B, T, 2, L.NH):
            hp = ws["head_part"] = torch.empty(B, T, 2, L.NH, device=u.device)
        v = lambda k: ptr(flat) + 4 * L.off[k][0]
        self.lib.call("b200rl_lstm_fwd_fused", ptr(u), ptr(hx0), ptr(cx0), v("Wt"), v("bt"), v("W_ih"), v("W_hh"),
                      v("b_ih"), v("b_hh"), v("W_head"), v("b_head"), ptr(ws["gates"]), ptr(ws["c"]), ptr(ws["h"]),
                      ptr(hp), ptr(head_out), B, T, Fin, L.NH, self._stream(dev))
        return 1
    self.lstm_fwd(L, flat, u, hx0, cx0, ws, variant=variant)
    self.head_fwd(L, flat, ws["h"], head_out)
    return 2 if (variant == 5 or variant < 2) else 3


def _head_fwd(self, L, flat, h, head_out):
    dev = self._chk(flat, h, head_out)
    rows = h.numel() // H
    v = lambda k: ptr(flat) + 4 * L.off[k][0]
    self.lib.call("b200rl_head_fwd", ptr(h), v("W_head"), v("b_head"), ptr(head_out), rows, L.NH, self._stream(dev))


def _lstm_bwd(self, L, flat, dhead, ws, cx0, variant=None, u=None, partials=None):
    """`u`, `partials`: the network input and the partial-gradient slabs -- given by callers that do NOT need d/d(input), so
    that variant 2 may fuse the head / torso gradients (it then sets ws["small_grads_done"] for lstm_wgrad)."""
    dev = self._chk(flat, dhead, ws["gates"], ws["c"], cx0)
    B, T = dhead.shape[0], dhead.shape[1]
    variant = BWD_VARIANT if variant is None else variant
    v = lambda k: ptr(flat) + 4 * L.off[k][0]
    o = lambda k: L.off[k][0]
    ws["small_grads_done"] = 0
    if variant in (2, 3):
        if u is not None and partials is not None and L.NH <= 4 and (variant == 3 or L.Fin == 4):
            self._chk(u, partials)
            dx = None
            if variant == 3:
                dx = ws.get("dx")
                if dx is None or dx.shape[:2] != (B, T):
                    dx = ws["dx"] = torch.empty(B, T, H, device=dhead.device)
            ws["dx_fresh"] = variant == 3
            self.lib.call("b200rl_lstm_bwd_fused", ptr(dhead), v("W_head"), v("W_hh"), v("W_ih"), ptr(ws["gates"]), ptr(ws["c"]),
                          ptr(cx0), ptr(u), v("Wt"), v("bt"), ptr(partials), partials.shape[0], L.P, o("Wt"), o("bt"), o("W_head"),
                          o("b_head"), ptr(dx), B, T, L.Fin, L.NH, self._stream(dev))
            ws["small_grads_done"] = 2 if variant == 2 else 1          # 2: head + torso gradients done ; 1: head gradients done
            return
        variant = 1
    dx = None
    if variant == 1:
        dx = ws.get("dx")
        if dx is None or dx.shape[:2] != (B, T):
            dx = ws["dx"] = torch.empty(B, T, H, device=dhead.device)
    ws["dx_fresh"] = variant == 1                 # lstm_wgrad may take the torso gradients from ws["dx"]
    self.lib.call("b200rl_lstm_bwd", ptr(dhead), v("W_head"), v("W_hh"), v("W_ih"), ptr(ws["gates"]), ptr(ws["c"]), ptr(cx0),
                  ptr(dx), B, T, L.NH, variant, self._stream(dev))


def _lstm_wgrad(self, L, flat, ws, hx0, u, dhead, partials, du=None, variant=0):
    dev = self._chk(flat, ws["gates"], ws["h"], hx0, u, dhead, partials, du)
    B, T, Fin = u.shape
    nparts = partials.shape[0]
    assert partials.shape[1] == L.P
    dx = None
    if variant == 1 and ws.get("small_grads_done") == 2 and du is None:
        variant = 4
    elif variant == 1 and ws.get("small_grads_done") == 1 and ws.get("dx_fresh") and du is None and Fin <= 8:
        variant, dx = 5, ws["dx"]
    elif variant == 1 and ws.get("dx_fresh") and du is None and Fin <= 8:
        variant, dx = 3, ws["dx"]
    v = lambda k: ptr(flat) + 4 * L.off[k][0]
    o = lambda k: L.off[k][0]
    self.lib.call("b200rl_lstm_wgrad", ptr(ws["gates"]), ptr(ws["h"]), ptr(hx0), ptr(u), ptr(dhead),
                  v("Wt"), v("bt"), v("W_ih"), ptr(partials), nparts, L.P,
                  o("Wt"), o("bt"), o("W_ih"), o("W_hh"), o("b_ih"), o("b_hh"), o("W_head"), o("b_head"),
                  ptr(du), ptr(dx), B, T, Fin, L.NH, variant, self._stream(dev))


def _reduce_partials(self, partials, grad):
    dev = self._chk(partials, grad)
    self.lib.call("b200rl_reduce_partials", ptr(partials), ptr(grad), partials.shape[0], partials.shape[1], self._stream(dev))


def _ppo_head(self, head_out, batch, dhead, scalars, hyper, A, continuous, outs=None):
    o = outs or {}
    dev = self._chk(head_out, batch["act"], batch["rew"], batch["is_fir"], batch["log_prob"], dhead, scalars, *o.values())
    B, T = head_out.shape[0], head_out.shape[1]
    g = lambda k: ptr(o.get(k))
    self.lib.call("b200rl_ppo_head", ptr(head_out), ptr(batch["act"]), ptr(batch["rew"]), ptr(batch["is_fir"]),
                  ptr(batch["log_prob"]), ptr(dhead), ptr(scalars),
                  g("logits"), g("log_probs"), g("entropy"), g("td_target"), g("gae"), g("ratio"),
                  B, T, A, int(continuous), hyper["gamma"], hyper["lmbda"], hyper["eps_clip"],
                  hyper["policy_loss_coef"], hyper["value_loss_coef"], hyper["entropy_coef"], self._stream(dev))


def _impala_head(self, head_out, batch, dhead, scalars, hyper, A, outs=None, rho_bar=0.8, c_bar=1.0):
    o = outs or {}
    dev = self._chk(head_out, batch["act"], batch["rew"], batch["is_fir"], batch["log_prob"], dhead, scalars, *o.values())
    B, T = head_out.shape[0], head_out.shape[1]
    g = lambda k: ptr(o.get(k))
    self.lib.call("b200rl_impala_head", ptr(head_out), ptr(batch["act"]), ptr(batch["rew"]), ptr(batch["is_fir"]),
                  ptr(batch["log_prob"]), ptr(dhead), ptr(scalars),
                  g("logits"), g("log_probs"), g("entropy"), g("ratio"), g("advantages"), g("values_target"),
                  B, T, A, hyper["gamma"], rho_bar, c_bar,
                  hyper["policy_loss_coef"], hyper["value_loss_coef"], hyper["entropy_coef"], self._stream(dev))


def _clip_rmsprop(self, param, grad, sq, n, grad_scale, max_norm, lr, alpha=0.99, eps=1e-5, norm_out=None):
    dev = self._chk(param, grad, sq, norm_out)
    self.lib.call("b200rl_clip_rmsprop", ptr(param), ptr(grad), ptr(sq), n, grad_scale, max_norm, lr, alpha, eps,
                  ptr(norm_out), self._stream(dev))


def _clip_adam(self, param, grad, m, v, n, grad_scale, max_norm, lr, step, b1=0.9, b2=0.999, eps=1e-8, norm_out=None, mask=None):
    dev = self._chk(param, grad, m, v, norm_out, mask)
    self.lib.call("b200rl_clip_adam", ptr(param), ptr(grad), ptr(m), ptr(v), n, grad_scale, max_norm, lr, b1, b2, eps,
                  step, ptr(mask), ptr(norm_out), self._stream(dev))


for _n, _f in (("lstm_fwd", _lstm_fwd), ("net_fwd", _net_fwd), ("head_fwd", _head_fwd), ("lstm_bwd", _lstm_bwd), ("lstm_wgrad", _lstm_wgrad),
               ("reduce_partials", _reduce_partials), ("ppo_head", _ppo_head), ("impala_head", _impala_head),
               ("clip_rmsprop", _clip_rmsprop), ("clip_adam", _clip_adam)):
    setattr(Ops, _n, _f)
