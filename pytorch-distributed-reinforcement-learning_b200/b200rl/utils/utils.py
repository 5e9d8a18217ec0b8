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
    side stream into the other of two PERSISTENT device buffer sets, so the next batch crosses PCIe while the
    current update runs; the following `stager(*batch)` call for that same tuple just makes the
    compute stream wait on the copy event.  The buffers are allocated once: per-step
    `tensor.to(device)` allocations on the copy stream plus `record_stream` cost 0.4 ms per step at
    B=4096, T=128 (tools/e2e_probe.py).  A set is overwritten only after the compute stream has passed the
    point where the learner asked for the NEXT batch (event `_released`), i.e. after every kernel that read it."""

    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:      # reference-style 'cuda': tensors report cuda:<index>
            self.device = torch.device("cuda", torch.cuda.current_device())
        # buffer sets: 0 / 1 = the learner-side double buffer (prefetch / __call__), ("ring", k) = the feeder-side ring (ship)
        self.pinned = defaultdict(dict)
        self._h2d_done = {}              # per set: event after the last async H2D that READ this set's pinned staging buffers
        self.devbuf = defaultdict(dict)
        self._released = {}              # compute-stream events: set k may be overwritten once this has happened
        self.flip = 0
        self._in_use = None              # buffer set the learner is currently reading
        self.ring_size = 3
        self._ring_busy = set()          # ring slots shipped and not yet released by the learner
        self._ring_next = 0
        self._ring_lookup = {}           # data_ptr of a shipped batch's first tensor -> (slot, copy event, bytes)
        self.bytes_last = 0
        self.skip = set()            # fields the algorithm never reads (not shipped)
        self._pending = None         # (id of host tuple, host tuple, device tuple, event, nbytes)
        self._copy_stream = None
        self.prefetch_hits = 0       # stager(*batch) calls served by an earlier prefetch(*batch)

    def _pin(self, slot, name, t):
        buf = self.pinned[slot].get(name)
        if buf is None or buf.shape != t.shape:
            buf = torch.empty(t.shape, dtype=torch.float32, pin_memory=self.device.type == "cuda")
            self.pinned[slot][name] = buf
        buf.copy_(t)
        return buf

    def _wait_staging_free(self, slot):
        """The host is about to overwrite set `slot`'s pinned staging buffers: the async H2D that last read them must have
        finished (`_released` only orders the DEVICE buffers; without this a slow copy could ship a half-overwritten batch)."""
        ev = self._h2d_done.get(slot)
        if ev is not None:
            ev.synchronize()
            self._h2d_done[slot] = None

    def _dev(self, slot, name, shape):
        buf = self.devbuf[slot].get(name)
        if buf is None or buf.shape != shape:
            buf = torch.empty(shape, dtype=torch.float32, device=self.device)
            self.devbuf[slot][name] = buf
        return buf

    def _ship(self, batch_args, slot):
        out, nbytes = [], 0
        self._wait_staging_free(slot)
        staged = False
        for name, t in zip(BATCH_FIELDS, batch_args):
            if t is None or name in self.skip:
                out.append(None)
                continue
            if name in ("hx", "cx") and t.dim() == 3:
                t = t[:, 0]
            if t.device == self.device:
                out.append(t.contiguous())
                continue
            if t.is_pinned() and t.is_contiguous():
                src = t
            elif (self.device.type == "cuda" and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1] and t.is_pinned()):
                # rows of a page-locked array at a fixed stride (hx / cx = [:, 0] of the registered (B, T, H) shared arrays):
                # one 2-D DMA instead of gathering 4096 rows scattered over 134 MB on the host
                from .. import native
                dst = self._dev(slot, name, t.shape)
                native.get().call("b200rl_h2d_rows", dst.data_ptr(), t.data_ptr(), t.shape[1] * 4, t.shape[0], t.stride(0) * 4,
                                  torch.cuda.current_stream(self.device).cuda_stream)
                out.append(dst)
                nbytes += t.numel() * 4
                continue
            else:
                src, staged = self._pin(slot, name, t), True
            dst = self._dev(slot, name, src.shape)
            dst.copy_(src, non_blocking=True)
            out.append(dst)
            nbytes += src.numel() * 4
        if staged and self.device.type == "cuda":
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._h2d_done[slot] = ev
        return tuple(out), nbytes

    def _release_current(self):
        """The learner is done enqueuing work on the set it was reading: mark the point on the compute stream."""
        if self.device.type == "cuda" and self._in_use is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._released[self._in_use] = ev
            if isinstance(self._in_use, tuple):              # a ring slot: the feeder may ship into it again
                self._ring_busy.discard(self._in_use[1])

    def prefetch(self, *batch_args):
        if self.device.type != "cuda" or self._pending is not None:
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        slot = self.flip ^ 1
        if slot == self._in_use:         # (cannot happen with two sets and one batch in flight; be safe)
            return
        self.flip = slot
        with torch.cuda.stream(self._copy_stream):
            if self._released.get(slot) is not None:
                self._copy_stream.wait_event(self._released[slot])
            dev, nbytes = self._ship(batch_args, slot)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._pending = (id(batch_args), batch_args, dev, ev, nbytes, slot)

    # ------------------------------------------------------------------ feeder side: zero-copy shipping
    def ship(self, *batch_args):
        """Start the H2D of a host batch on the copy stream into the next free slot of a small ring of persistent device buffer
        sets and return (device tuple, copy event) -- or None when every slot still holds a batch the learner has not consumed
        (back-pressure).  Meant for the BATCHING coroutine: with the shared-memory arrays page-locked
        (`SMInterFace.register_pinned`) the contiguous fields are DMA'd straight out of the slots the storage process wrote --
        no defensive host copy, no staging copy -- and the coroutine resets the slot counter once the event has fired.  The
        tuple is later recognised by `stager(*tuple)`, which only makes the compute stream wait for the copy."""
        if self.device.type != "cuda":
            return None
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        k = None
        for d in range(self.ring_size):
            c = (self._ring_next + d) % self.ring_size
            if c not in self._ring_busy:
                k = c
                break
        if k is None:
            return None
        self._ring_next = (k + 1) % self.ring_size
        self._ring_busy.add(k)
        key = ("ring", k)
        with torch.cuda.stream(self._copy_stream):
            if self._released.get(key) is not None:
                self._copy_stream.wait_event(self._released[key])
            dev, nbytes = self._ship(batch_args, key)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        first = next(t for t in dev if t is not None)
        self._ring_lookup[first.data_ptr()] = (k, ev, nbytes)
        return dev, ev

    def __call__(self, *batch_args):
        self._release_current()
        first = next((t for t in batch_args if t is not None), None)
        if first is not None and first.is_cuda and first.data_ptr() in self._ring_lookup:      # shipped by the feeder
            k, ev, nbytes = self._ring_lookup[first.data_ptr()]
            torch.cuda.current_stream(self.device).wait_event(ev)
            self.bytes_last = nbytes
            self._in_use = ("ring", k)
            return batch_args
        pend = self._pending
        # (`*batch_args` re-packs the caller's tuple, so compare the tensors, not the tuple object)
        if pend is not None and len(pend[1]) == len(batch_args) and all(x is y for x, y in zip(pend[1], batch_args)):
            self._pending = None
            self.prefetch_hits += 1
            torch.cuda.current_stream(self.device).wait_event(pend[3])
            self.bytes_last = pend[4]
            self._in_use = pend[5]
            return pend[2]
        if pend is not None:             # a prefetch for a different tuple is in flight on the other set: let it land first
            torch.cuda.current_stream(self.device).wait_event(pend[3])
            self._pending = None
        self.flip ^= 1
        dev, self.bytes_last = self._ship(batch_args, self.flip)     # same stream as the kernels: ordered after every earlier reader
        self._in_use = self.flip
        return dev


def make_gpu_batch(*args, device):
    """Reference-compatible signature (utils.py:101-103): every tensor to `device`."""
    return tuple(t.to(device) for t in args)
