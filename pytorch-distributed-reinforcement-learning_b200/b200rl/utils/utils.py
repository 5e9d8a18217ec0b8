"""Host utilities the learner path touches -- mirror of the reference's `utils/utils.py`
surface for that path: `Params` (the `parameters.json` surface, utils.py:24-27),
`ExecutionTimer` (the reference's own throughput metric, utils.py:167-188),
`make_gpu_batch` (the H2D step, utils.py:101-103), `to_torch`, the wire `Protocol`.
"""
import copy
import json
import os
import time
from collections import defaultdict, deque
from contextlib import contextmanager
from enum import Enum, auto
from types import SimpleNamespace

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def load_params(path=None):
    """`parameters.json` -> SimpleNamespace.  Like the reference, a `utils/parameters.json`
    under the current working directory wins; otherwise the packaged copy (same keys and
    defaults as the reference's file) is used."""
    cands = [path] if path else [os.path.join(os.getcwd(), "utils", "parameters.json"),
                                 os.path.join(_HERE, "parameters.json")]
    for p in cands:
        if p and os.path.exists(p):
            with open(p) as f:
                return SimpleNamespace(**json.load(f))
    raise FileNotFoundError("parameters.json")


Params = load_params()

DataFrameKeyword = ["obs_batch", "act_batch", "rew_batch", "logits_batch", "log_prob_batch",
                    "is_fir_batch", "hx_batch", "cx_batch", "batch_num"]

to_torch = lambda nparray: torch.from_numpy(nparray).type(torch.float32)
flatten = lambda obj: obj.numpy().reshape(-1).astype(np.float32)


def mul(shape_dim):
    v = 1
    for e in shape_dim:
        v *= e
    return v


class Protocol(Enum):
    Model = auto()
    Rollout = auto()
    Stat = auto()


def hyper_from_args(args):
    """The keys of `args` the update kernels consume (Config row)."""
    keys = ("gamma", "lmbda", "eps_clip", "policy_loss_coef", "value_loss_coef", "entropy_coef", "max_grad_norm",
            "lr", "alpha", "coef_eta", "coef_alpha_upper", "coef_alpha_below", "v_mpo_lagrange_multiplier_init",
            "K_epoch", "hidden_size")
    return {k: getattr(args, k) for k in keys if hasattr(args, k)}


class ExecutionTimer:
    """Rolling wall-clock timers per named block; `check_throughput=True` also records
    `num_transition / elapsed` (transition/sec), the reference's learner throughput metric."""

    def __init__(self, threshold=100, num_transition=None):
        self.num_transition = num_transition
        self.timer_dict = defaultdict(lambda: deque(maxlen=threshold))
        self.throughput_dict = defaultdict(lambda: deque(maxlen=threshold))

    @contextmanager
    def timer(self, code_block_name, check_throughput=False):
        t0 = time.time()
        yield
        dt = time.time() - t0
        self.timer_dict[code_block_name].append(dt)
        if check_throughput and isinstance(self.num_transition, (int, float, np.number)):
            self.throughput_dict[code_block_name].append(self.num_transition / (dt + 1e-6))


# ------------------------------------------------------------------ host -> device
# Reference batch tuple order (agents/learner.py:224-233):
#   (obs, act, rew, logits, log_prob, is_fir, hx, cx), each (B, T, width) fp32 on the host.
BATCH_FIELDS = ("obs", "act", "rew", "logits", "log_prob", "is_fir", "hx", "cx")


class HostStager:
    """Pinned staging + async H2D, shipping only what the update reads: `hx`/`cx` are cut to
    `[:, 0]` on the host (the reference copies all (B,T,H) of both -- 268 MB at (4096,128) -- and then
    reads `[:, 0]`, ppo/learning.py:26,37-40) and fields the algorithm never reads are skipped.

    `stager(*batch)` copies on the current stream.  `prefetch(*batch)` starts the same copies on a
    side stream into the other of two buffer sets, so the next batch crosses PCIe while the
    current update runs; the following `stager(*batch)` call for that same tuple just makes the
    compute stream wait on the copy event."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.pinned = [{}, {}]
        self.flip = 0
        self.bytes_last = 0
        self.skip = set()            # fields the algorithm never reads (not shipped)
        self._pending = None         # (id of host tuple, device tuple, event, nbytes)
        self._copy_stream = None

    def _pin(self, slot, name, t):
        buf = self.pinned[slot].get(name)
        if buf is None or buf.shape != t.shape:
            buf = torch.empty(t.shape, dtype=torch.float32, pin_memory=self.device.type == "cuda")
            self.pinned[slot][name] = buf
        buf.copy_(t)
        return buf

    def _ship(self, batch_args, slot):
        out, nbytes = [], 0
        for name, t in zip(BATCH_FIELDS, batch_args):
            if t is None or name in self.skip:
                out.append(None)
                continue
            if name in ("hx", "cx") and t.dim() == 3:
                t = t[:, 0]
            if t.device == self.device:
                out.append(t.contiguous())
                continue
            src = t if (t.is_pinned() and t.is_contiguous()) else self._pin(slot, name, t)
            out.append(src.to(self.device, non_blocking=True))
            nbytes += src.numel() * 4
        return tuple(out), nbytes

    def prefetch(self, *batch_args):
        if self.device.type != "cuda" or self._pending is not None:
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        self.flip ^= 1
        with torch.cuda.stream(self._copy_stream):
            dev, nbytes = self._ship(batch_args, self.flip)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._pending = (id(batch_args), batch_args, dev, ev, nbytes)

    def __call__(self, *batch_args):
        pend = self._pending
        if pend is not None and pend[1] is batch_args:
            self._pending = None
            torch.cuda.current_stream(self.device).wait_event(pend[3])
            for t in pend[2]:
                if t is not None:
                    t.record_stream(torch.cuda.current_stream(self.device))
            self.bytes_last = pend[4]
            return pend[2]
        self.flip ^= 1
        dev, self.bytes_last = self._ship(batch_args, self.flip)
        return dev


def make_gpu_batch(*args, device):
    """Reference-compatible signature (utils.py:101-103): every tensor to `device`."""
    return tuple(t.to(device) for t in args)
