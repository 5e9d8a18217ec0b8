"""Communicators for the data-parallel learner (one process per GPU, batch sharded over ranks).

`TorchComm`      torch.distributed (NCCL on GPUs, gloo in the CPU tests): plumbing and the fallback data plane.
`PeerGroup`      symmetric buffers mapped into every rank through CUDA IPC (NVLink / NVSwitch peer access) and the
                 fused {one-shot all-reduce -> global-norm clip -> optimizer} kernel of `csrc/peer_kernels.cu`: the
                 per-step gradient exchange has no library collective on the stream.
`LoopbackWorld`  test-only: W "ranks" as threads of ONE process on one GPU (own stream each).  The peer kernels then
                 handshake through flags in the same device memory, which lets a single-GPU box exercise the exact
                 kernel the 8-GPU job runs.
"""
import ctypes as C
import os
import threading

import torch

from . import native

FLAG_STRIDE = 32          # u32 per flag (one 128-byte line), must match kFlagStride
MAX_WORLD = 16
HEADER_BYTES = 8192       # [0, 2048) flags | [2048] grid-barrier counter | [4096, 4608) per-CTA partials | [5120] status
GRID = 32


class _Raw:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _view(ptr, n, dtype, device):
    """A tensor over raw memory at `ptr` (device memory through the CUDA array interface; host memory -- the test-only
    emulated library -- through a ctypes buffer)."""
    typestr, size = {torch.float32: ("<f4", 4), torch.float64: ("<f8", 8), torch.int32: ("<i4", 4), torch.uint8: ("|u1", 1)}[dtype]
    if torch.device(device).type != "cuda":
        return torch.frombuffer((C.c_ubyte * (n * size)).from_address(ptr), dtype=dtype)
    return torch.as_tensor(_Raw(ptr, n, typestr), device=device)


class TorchComm:
    def __init__(self, process_group):
        import torch.distributed as dist
        self.dist = dist
        self.pg = process_group
        self.world = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)

    def all_reduce(self, t, op="sum"):
        d = self.dist
        d.all_reduce(t, op={"sum": d.ReduceOp.SUM, "max": d.ReduceOp.MAX, "min": d.ReduceOp.MIN}[op], group=self.pg)

    def all_gather(self, out, inp):
        self.dist.all_gather_into_tensor(out, inp, group=self.pg)

    def exchange_bytes(self, b):
        out = [None] * self.world
        self.dist.all_gather_object(out, b, group=self.pg)
        return out

    def barrier(self):
        self.dist.barrier(group=self.pg)


class PeerChannel:
    """One exchanged flat buffer (n floats + 8 fp64 loss sums), double-buffered by step parity."""

    def __init__(self, n, rank, world, device, lib, exchange=None, local_regions=None):
        assert n % 4 == 0 and 1 <= world <= MAX_WORLD
        self.n, self.rank, self.world, self.device, self.lib = n, rank, world, device, lib
        self.buf_bytes = (n * 4 + 64 + 255) // 256 * 256
        self.bytes = HEADER_BYTES + 2 * self.buf_bytes
        self.epoch = 0
        self.gsync_target = 0
        self._opened = []
        if local_regions is not None:                      # loopback: every rank's region lives in this process
            self.base = local_regions[rank]
            bases = list(local_regions)
            self._own = False
        else:
            p, h = C.c_void_p(), (C.c_ubyte * 64)()
            with torch.cuda.device(device):
                lib.call("b200rl_peer_alloc", self.bytes, C.byref(p), h)
            self.base, self._own = p.value, True
            handles = exchange(bytes(h))
            bases = []
            for r, hb in enumerate(handles):
                if r == rank:
                    bases.append(self.base)
                    continue
                q = C.c_void_p()
                with torch.cuda.device(device):
                    lib.call("b200rl_peer_open", (C.c_ubyte * 64).from_buffer_copy(hb), C.byref(q))
                bases.append(q.value)
                self._opened.append(q.value)
        self.bases = bases
        data = lambda base, p: base + HEADER_BYTES + p * self.buf_bytes
        arr = lambda xs: (C.c_void_p * world)(*xs)
        self._grad_ptrs = [arr([data(b, p) for b in bases]) for p in (0, 1)]
        self._scal_ptrs = [arr([data(b, p) + n * 4 for b in bases]) for p in (0, 1)]
        self._flag_ptrs = arr(bases)
        self.grad_views = [_view(data(self.base, p), n, torch.float32, device) for p in (0, 1)]
        self.scal_views = [_view(data(self.base, p) + n * 4, 8, torch.float64, device) for p in (0, 1)]
        self.status = _view(self.base + 5120, 1, torch.int32, device)
        self.grid = GRID
        self.pre_step = None           # test-only hook (LoopbackWorld): called right before the kernel is launched

    def begin(self):
        """Start the next step: returns (gradient buffer, loss-sum buffer) this rank must fill before `step()`."""
        self.epoch += 1
        p = self.epoch & 1
        return self.grad_views[p], self.scal_views[p]

    def step(self, optimizer, param, s1, s2, mask, grad_out, scal_out, norm_out, grad_scale, max_norm, lr, a1, a2, eps,
             adam_step, stream):
        p = self.epoch & 1
        self.gsync_target += self.grid
        if self.pre_step is not None:
            self.pre_step()
        ptr = native.ptr
        self.lib.call("b200rl_peer_step", optimizer, ptr(param), ptr(s1), ptr(s2), ptr(mask), ptr(grad_out),
                      self._grad_ptrs[p], self._scal_ptrs[p], self._flag_ptrs, self.base, self.base + 2048, self.base + 4096,
                      ptr(scal_out), ptr(norm_out), self.base + 5120, self.n, self.rank, self.world, self.epoch, self.grid,
                      self.gsync_target & 0xffffffff, grad_scale, max_norm, lr, a1, a2, eps, adam_step, stream)

    def check(self):
        """Host check of the kernel's status word (synchronises): raises if a peer never announced its gradient."""
        s = int(self.status.item())
        if s != 0:
            raise native.NativeError(f"b200rl peer step: rank {self.rank} timed out waiting for a peer's gradient (status {s})")

    def close(self):
        for q in self._opened:
            self.lib.call("b200rl_peer_close", q)
        self._opened = []
        if self._own and self.base:
            self.lib.call("b200rl_peer_free", self.base)
            self.base = None


class PeerGroup(TorchComm):
    """torch.distributed for set-up and the rare collectives, peer channels for the per-step gradient exchange."""

    def __init__(self, process_group, device, lib=None):
        super().__init__(process_group)
        self.device = torch.device(device)
        self.lib = lib or native.get()
        self.channels = []

    @staticmethod
    def available(device):
        return torch.device(device).type == "cuda" and os.environ.get("B200RL_PEER", "1") != "0"

    def channel(self, n):
        ch = PeerChannel(n, self.rank, self.world, self.device, self.lib, exchange=self.exchange_bytes)
        self.barrier()                 # every rank has mapped every region before anyone starts writing flags
        self.channels.append(ch)
        return ch

    def close(self):
        self.barrier()
        for ch in self.channels:
            ch.close()
        self.channels = []


# ------------------------------------------------------------------------------------------------ test-only loopback
class LoopbackWorld:
    """W ranks as threads of one process on one device.  `comm(rank)` returns that rank's communicator; collectives meet at
    a threading.Barrier after synchronising each rank's stream; `channel(n)` hands out peer channels whose "peer"
    regions are plain device allocations of this process."""

    def __init__(self, world, device, lib=None):
        self.world, self.device = world, torch.device(device)
        self.lib = lib or native.get()
        self.bar = threading.Barrier(world)
        self.slots = [None] * world
        self.lock = threading.Lock()
        self._regions = {}

    def comm(self, rank):
        return _LoopbackComm(self, rank)

    def _region(self, key, nbytes):
        with self.lock:
            if key not in self._regions:
                self._regions[key] = [torch.zeros(nbytes, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
            return self._regions[key]


class _LoopbackComm:
    def __init__(self, w, rank):
        self.w, self.rank, self.world, self.device, self.lib = w, rank, w.world, w.device, w.lib
        self._nch = 0

    def _sync(self):
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()

    def _meet(self, t):
        self._sync()
        self.w.slots[self.rank] = t
        self.w.bar.wait()
        got = list(self.w.slots)
        self.w.bar.wait()
        return got

    def all_reduce(self, t, op="sum"):
        got = self._meet(t.clone())
        s = torch.stack(got)
        t.copy_({"sum": lambda: s.sum(0), "max": lambda: s.max(0).values, "min": lambda: s.min(0).values}[op]())

    def all_gather(self, out, inp):
        got = self._meet(inp.clone())
        out.copy_(torch.cat([g.reshape(-1) for g in got]).view_as(out))

    def barrier(self):
        self._sync()
        self.w.bar.wait()

    def channel(self, n):
        key = (self._nch, n)
        self._nch += 1
        proto_bytes = HEADER_BYTES + 2 * ((n * 4 + 64 + 255) // 256 * 256)
        regions = self.w._region(key, proto_bytes)
        self.barrier()
        ch = PeerChannel(n, self.rank, self.world, self.device, self.lib, local_regions=[r.data_ptr() for r in regions])
        # all "ranks" share ONE GPU here: keep the spinning peer kernels small enough that whole SMs stay free for the other
        # ranks' recurrent kernels (which need an SM to themselves)
        ch.grid = max(1, min(GRID, -(-(n // 4) // (256 * 4))))
        # ... and let every rank drain its stream and meet the others before the handshake kernels are launched: with all "ranks"
        # in one process their streams share the device's launch queues, and a handshake kernel that spins while another rank's
        # producer kernels are still queued behind it can wait for ever.  (One process per GPU has no such coupling.)
        ch.pre_step = self.barrier
        return ch

    def close(self):
        pass
