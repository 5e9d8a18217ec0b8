"""Rollout -> batch assembly on the native assembler (`csrc/assembler.cu`).

Mirror of the reference's `buffers/rollout_assembler.py:24-82` (`RolloutAssembler(args, ready_roll)`, `async push(data)`,
`async pop()`) fused with `LearnerStorage.make_batch` (`agents/learner_storage.py:123-159`): bound to the shared-memory
arrays (`bind_shared`), every environment step is copied ONCE, straight into its row of a slot-shaped staging block, and
a finished trajectory is eight memcpys into the next free slot of the (page-locked) arrays the learner DMA's from.
Without shared memory it behaves like the reference class: finished trajectories come out of `pop()` as dicts of
`(T, width)` tensors.  The reference's semantics are kept (0.5 s age bound on active trajectories, continuation of the
first-parked unfinished trajectory with an `is_fir = 1` seam, drop when the shared batch is full).
"""
import asyncio
import ctypes as C
import time

import numpy as np
import torch

from .. import native

FIELDS = ("obs", "act", "rew", "logits", "log_prob", "is_fir", "hx", "cx")


class RolloutAssembler:
    def __init__(self, args, ready_roll=None, lib=None, clock=time.monotonic, max_age=0.5):
        self.args = args
        self.seq_len = args.seq_len
        self.ready_roll = ready_roll
        self.clock = clock
        self.lib = lib or native.get()
        self._h = None
        self._max_age = max_age
        self._obs_w = None
        self._own = None              # queue mode: a private one-slot "shared" block
        self._count = None

    # ------------------------------------------------------------------ set-up
    def _create(self, obs_w):
        fn = self.lib.b200rl_asm_create
        fn.restype = C.c_void_p
        self._h = fn(int(self.seq_len), int(obs_w), int(self.args.action_space), int(self.args.hidden_size), float(self._max_age))
        if not self._h:
            raise native.NativeError("asm_create: " + self.lib.last_error())
        self._obs_w = obs_w

    def bind_shared(self, shm_ref, obs_shape):
        """Write finished trajectories straight into the learner's shared arrays (`setup_shared_memory` layout)."""
        obs_w = int(np.prod(obs_shape))
        self._create(obs_w)
        views = [np.frombuffer(shm_ref[f"{k}_batch"][0].get_obj(), dtype=np.float32) for k in FIELDS]
        capacity = views[1].shape[0] // self.seq_len
        counter = shm_ref["batch_index"]
        self._count = counter
        cptr = C.addressof(counter.get_obj())
        self.lib.call("b200rl_asm_bind", self._h, *[v.ctypes.data for v in views], cptr, capacity)
        self._views = views
        return self

    def _bind_private(self, obs_w):
        self._create(obs_w)
        T, A, H = self.seq_len, self.args.action_space, self.args.hidden_size
        widths = (obs_w, 1, 1, A, 1, 1, H, H)
        self._own = [np.zeros(T * w, dtype=np.float32) for w in widths]
        self._own_count = (C.c_int * 1)(0)
        self.lib.call("b200rl_asm_bind", self._h, *[v.ctypes.data for v in self._own], C.addressof(self._own_count), 1)
        self._widths = widths

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                self.lib.b200rl_asm_destroy(C.c_void_p(self._h))
            except Exception:  # noqa: BLE001
                pass
            self._h = None

    # -------------------------------------------------------------------- steps
    @staticmethod
    def _f32(x, n):
        a = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
        assert a.size == n, (a.size, n)
        return a

    def push_now(self, data, now=None):
        """One environment step (the dict a worker sends, `agents/worker.py:108-128`).  Returns 1 if it completed a trajectory
        that was written out, 2 if the trajectory had to be dropped (shared batch full), else 0."""
        for k in FIELDS + ("id", "done"):
            if k == "logits" and k not in data:
                continue
            assert k in data, k
        if self._h is None:
            self._bind_private(int(np.prod(np.shape(data["obs"]))))
        A, H = self.args.action_space, self.args.hidden_size
        obs = self._f32(data["obs"], self._obs_w)
        logits = self._f32(data["logits"], A) if "logits" in data else np.zeros(A, dtype=np.float32)
        hx, cx = self._f32(data["hx"], H), self._f32(data["cx"], H)
        sc = lambda k: float(np.asarray(data[k].detach().cpu() if isinstance(data[k], torch.Tensor) else data[k]).reshape(-1)[0])
        rc = self.lib.b200rl_asm_push(C.c_void_p(self._h), int(hash(data["id"]) if not isinstance(data["id"], int) else data["id"]),
                                      int(bool(data["done"])), float(self.clock() if now is None else now),
                                      obs.ctypes.data, sc("act"), sc("rew"), logits.ctypes.data, sc("log_prob"), sc("is_fir"),
                                      hx.ctypes.data, cx.ctypes.data)
        if rc < 0:
            raise native.NativeError("asm_push: " + self.lib.last_error())
        return rc

    async def push(self, data):
        rc = self.push_now(data)
        if rc == 1 and self._own is not None and self.ready_roll is not None:
            T = self.seq_len
            out = {k: torch.from_numpy(v.reshape(T, w).copy()) for k, v, w in zip(FIELDS, self._own, self._widths)}
            self._own_count[0] = 0
            await self.ready_roll.put(out)
        return rc

    async def pop(self):
        return await self.ready_roll.get()

    def stats(self):
        s = (C.c_longlong * 6)()
        self.lib.call("b200rl_asm_stats", self._h, s)
        return dict(zip(("active", "parked", "emitted", "dropped_full", "dropped_age", "spliced"), [int(x) for x in s]))
