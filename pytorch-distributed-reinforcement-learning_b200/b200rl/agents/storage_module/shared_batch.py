"""Shared-memory batch layout between LearnerStorage (writer) and Learner (reader).

Same contract as the reference's `agents/storage_module/shared_batch.py:19-107`: one flat
fp32 `multiprocessing.Array` per field, sized `seq_len * batch * width` with widths
obs=f, act=1, rew=1, logits=A, log_prob=1, is_fir=1, hx=H, cx=H, plus an int slot counter
under the key "batch_index"; numpy views are created on the reader side.
"""
import multiprocessing as mp

import numpy as np

from ...utils.utils import mul

FIELD_WIDTHS = lambda args, obs_shape: (
    ("obs_batch", mul(obs_shape)), ("act_batch", 1), ("rew_batch", 1), ("logits_batch", args.action_space),
    ("log_prob_batch", 1), ("is_fir_batch", 1), ("hx_batch", args.hidden_size), ("cx_batch", args.hidden_size))


def setup_shared_memory(args, obs_shape, batch_size):
    shm_ref = {}
    for name, width in FIELD_WIDTHS(args, obs_shape):
        shm_ref[name] = (mp.Array("f", args.seq_len * batch_size * width), np.float32)
    counter = mp.Value("i", 0)
    shm_ref["batch_index"] = counter
    return shm_ref


def reset_shared_on_policy_memory(args, obs_shape):
    return setup_shared_memory(args, obs_shape, batch_size=args.batch_size)


def reset_shared_buffer_memory(args, obs_shape):
    return setup_shared_memory(args, obs_shape, batch_size=args.buffer_size)


class SMInterFace:
    def __init__(self, shm_ref):
        self.shm_ref = shm_ref

    def get_shared_memory(self, name):
        arr, dtype = self.shm_ref[name]
        return np.frombuffer(buffer=arr.get_obj(), dtype=dtype, count=-1)

    def get_shared_memory_interface(self):
        for key in ("obs", "act", "rew", "logits", "log_prob", "is_fir", "hx", "cx"):
            setattr(self, f"sh_{key}_batch", self.get_shared_memory(f"{key}_batch"))
        self.sh_data_num = self.shm_ref.get("batch_index")
        self.reset_data_num()

    def reset_data_num(self):
        self.sh_data_num.value = 0

    # ---- page-locking the shared arrays: the learner's DMA engine reads rollouts straight out of the slots the storage
    # process wrote instead of a defensive copy + a staging copy per batch
    def register_pinned(self):
        """cudaHostRegister every field array (idempotent).  Returns True when all of them are page-locked."""
        import torch
        if getattr(self, "_registered", None):
            return True
        rt = torch.cuda.cudart()
        done = []
        for key in ("obs", "act", "rew", "logits", "log_prob", "is_fir", "hx", "cx"):
            a = getattr(self, f"sh_{key}_batch")
            err = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
            if int(err) != 0:
                self._register_error = f"cudaHostRegister({key}, {a.nbytes} B) -> {err!r}"
                for p in done:
                    rt.cudaHostUnregister(p)
                return False
            done.append(a.ctypes.data)
        self._registered = done
        return True

    def unregister_pinned(self):
        import torch
        for p in getattr(self, "_registered", None) or []:
            torch.cuda.cudart().cudaHostUnregister(p)
        self._registered = None
