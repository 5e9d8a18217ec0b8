"""ctypes binding of the C-ABI in include/b200rl.h.

The product library is `../lib/libb200rl.so` (built by `b200rl.build.build()` /
`__graft_entry__.build()`).  There is NO fallback: if the library is missing or a
tensor is not on a CUDA device, the call raises.
"""
import ctypes as C
import os

import torch

from .build import LIB

_P = C.c_void_p
_I = C.c_int
_L = C.c_longlong
_F = C.c_float

# name -> argtypes (restype is always int unless listed in _RESTYPE)
SIGNATURES = {
    "b200rl_version": [],
    "b200rl_last_error": [],
    "b200rl_compute_gae": [_P, _P, _I, _I, _F, _F, _P],
    "b200rl_td_gae": [_P, _P, _P, _P, _P, _I, _I, _F, _F, _P],
    "b200rl_compute_v_trace": [_P] * 8 + [_I, _I, _F, _F, _F, _P],
    "b200rl_kldivergence": [_P, _P, _P, _L, _I, _P],
    "b200rl_soft_update": [_P, _P, _L, _F, _P],
    "b200rl_lstm_fwd": [_P] * 12 + [_I, _I, _I, _I, _P],
    "b200rl_lstm_fwd_fused": [_P] * 16 + [_I, _I, _I, _I, _P],
    "b200rl_head_fwd": [_P, _P, _P, _P, _L, _I, _P],
    "b200rl_lstm_bwd": [_P] * 8 + [_I, _I, _I, _I, _P],
    "b200rl_lstm_bwd_fused": [_P] * 11 + [_I, _L, _L, _L, _L, _L, _P, _I, _I, _I, _I, _P],
    "b200rl_lstm_wgrad": [_P] * 8 + [_P, _I, _L] + [_L] * 8 + [_P, _P, _I, _I, _I, _I, _I, _P],
    "b200rl_reduce_partials": [_P, _P, _I, _L, _P],
    "b200rl_ppo_head": [_P] * 13 + [_I, _I, _I, _I] + [_F] * 6 + [_P],
    "b200rl_impala_head": [_P] * 13 + [_I, _I, _I] + [_F] * 6 + [_P],
    "b200rl_vmpo_gae": [_P] * 6 + [_I, _I, _I, _F, _F, _F, C.c_double, _P],
    "b200rl_vmpo_select": [_P, _I, _I, _I, _P, _P, _P, _P, _P],
    "b200rl_vmpo_grad": [_P] * 10 + [_I, _I, _I, _I, _F, C.c_double, _P],
    "b200rl_vmpo_dual_step": [_P] * 4 + [C.c_double] * 9 + [_P],
    "b200rl_clip_rmsprop": [_P, _P, _P, _L] + [_F] * 5 + [_P, _P],
    "b200rl_clip_adam": [_P, _P, _P, _P, _L] + [_F] * 6 + [_I, _P, _P, _P],
    "b200rl_asm_create": [_I, _I, _I, _I, C.c_double],
    "b200rl_asm_destroy": [_P],
    "b200rl_asm_bind": [_P] * 10 + [_I],
    "b200rl_asm_push": [_P, _L, _I, C.c_double, _P, _F, _F, _P, _F, _F, _P, _P],
    "b200rl_asm_stats": [_P, _P],
    "b200rl_h2d_rows": [_P, _P, _L, _L, _L, _P],
    "b200rl_peer_alloc": [_L, _P, _P],
    "b200rl_peer_open": [_P, _P],
    "b200rl_peer_close": [_P],
    "b200rl_peer_free": [_P],
    "b200rl_peer_set_timeout_ms": [_L],
    "b200rl_peer_step": [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, C.c_uint, _I, C.c_uint]
                        + [_F] * 6 + [_I, _P],
    "b200rl_sac_actor_head": [_P] * 6 + [_I, _I, _I, _P],
    "b200rl_sac_critic_head": [_P] * 13 + [_I, _I, _I, _F, _P],
    "b200rl_sac_cont_sample": [_P] * 5 + [_I, _P, _I, _I, _I, _P],
    "b200rl_sac_cont_pack": [_P] * 5 + [_I, _I, _I, _I, _P],
    "b200rl_sac_cont_minq_grad": [_P] * 5 + [_I, _I, _I, _P],
    "b200rl_sac_cont_actor_grad": [_P] * 3 + [_I] + [_P] * 6 + [_I, _I, _I, _I, _P],
    "b200rl_sac_cont_critic_head": [_P] * 12 + [_I, _I, _I, _F, _P],
    "b200rl_sac_alpha_step": [_P, _P, _P, C.c_double, C.c_double, C.c_double, _I, _P],
}
_RESTYPE = {"b200rl_last_error": C.c_char_p, "b200rl_asm_create": C.c_void_p}


class NativeError(RuntimeError):
    pass


class Lib:
    """A loaded b200rl shared library.  `emulated=True` is reserved for the test-only
    host-emulation build (tests/emul); the product never sets it."""

    def __init__(self, path, emulated=False):
        if not os.path.exists(path):
            raise NativeError(
                f"b200rl native library not found at {path}; build it with "
                "`python __graft_entry__.py` (nvcc, sm_100a). There is no CPU fallback.")
        self.path = path
        self.emulated = emulated
        self._dll = C.CDLL(path)
        self.missing = []
        for name, argtypes in SIGNATURES.items():
            try:
                fn = getattr(self._dll, name)
            except AttributeError:
                self.missing.append(name)
                continue
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, C.c_int)
            setattr(self, name, fn)

    def last_error(self):
        return self.b200rl_last_error().decode()

    def call(self, name, *args):
        rc = getattr(self, name)(*args)
        if rc != 0:
            raise NativeError(f"{name}: {self.last_error()}")


_lib = None


def get():
    """The product library (loads on first use; raises NativeError if absent)."""
    global _lib
    if _lib is None:
        path = os.environ.get("B200RL_LIB") or LIB        # (B200RL_LIB: A/B builds of the same sources, tools/build_variant.py)
        _lib = Lib(path)
        if _lib.missing:
            raise NativeError(f"{path} lacks symbols {_lib.missing}; rebuild")
    return _lib


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream
