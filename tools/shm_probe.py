"""Where does the shared-memory e2e path spend its time?  H2D bandwidth from (a) torch pinned memory, (b) a multiprocessing.Array
page-locked with cudaHostRegister, (c) pageable memory; and the feeder-side calls of one batch (dev tool)."""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200"))
sys.path.insert(0, ROOT)
import torch


def bw(src, dst, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        dst.copy_(src, non_blocking=True)
    e1.record()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return src.numel() * 4 / ms / 1e6, host / n * 1e3


def main():
    dev = torch.device("cuda:0")
    n = 4096 * 128 * 4
    dst = torch.empty(n, device=dev)
    pinned = torch.randn(n).pin_memory()
    print("torch pinned      : %.1f GB/s (host call %.3f ms)" % bw(pinned, dst))
    arr = mp.Array("f", n)
    view = np.frombuffer(arr.get_obj(), dtype=np.float32)
    view[:] = 1.0
    t = torch.from_numpy(view)
    print("mp.Array pageable : %.1f GB/s (host call %.3f ms), is_pinned %s" % (*bw(t, dst), t.is_pinned()))
    err = torch.cuda.cudart().cudaHostRegister(view.ctypes.data, view.nbytes, 0)
    t = torch.from_numpy(view)
    print("cudaHostRegister ->", err, "is_pinned", t.is_pinned())
    print("mp.Array registered: %.1f GB/s (host call %.3f ms)" % bw(t, dst))
    big = mp.Array("f", 4096 * 128 * 64)
    bv = np.frombuffer(big.get_obj(), dtype=np.float32)
    bv[:] = 2.0
    torch.cuda.cudart().cudaHostRegister(bv.ctypes.data, bv.nbytes, 0)
    hx = torch.from_numpy(bv.reshape(4096, 128, 64)[:, 0])
    stage = torch.empty(4096, 64).pin_memory()
    t0 = time.perf_counter()
    for _ in range(10):
        stage.copy_(hx)
    print("strided [:,0] gather of (4096,128,64) into pinned staging: %.3f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
    pg = torch.randn(n)
    t0 = time.perf_counter()
    for _ in range(10):
        pinned.copy_(pg)
    print("19 MB... %d MB pageable -> pinned host memcpy: %.3f ms" % (n * 4 >> 20, (time.perf_counter() - t0) / 10 * 1e3))
    t0 = time.perf_counter()
    for _ in range(10):
        np.array(view, dtype=np.float32, copy=True)
    print("np.array(copy=True) of %d MB: %.3f ms" % (n * 4 >> 20, (time.perf_counter() - t0) / 10 * 1e3))
    print("torch threads", torch.get_num_threads(), "cpus", len(os.sched_getaffinity(0)))


if __name__ == "__main__":
    main()
