import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200")); sys.path.insert(0, ROOT)
import torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "torch threads", torch.get_num_threads(), flush=True)
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    if os.path.exists(p): print(p, open(p).read().strip(), flush=True)
print(open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0], flush=True)
import oracle
from oracle import models as OM
from b200rl.synth import make_rollouts
for nt in (8, 16, 32, 64):
    torch.set_num_threads(nt)
    torch.manual_seed(0)
    L = oracle.OracleLearner("PPO", OM.init_params("single", 4, 2), fused_cell=True)
    b = {k: v[:512].contiguous() for k, v in make_rollouts(512, 128, 4, 2).items()}
    L.update(b)
    t0 = time.perf_counter(); L.update(b); dt = time.perf_counter() - t0
    print("threads", nt, "B=512 T=128 update", round(dt, 3), "s ->", round(512 * 128 / dt), "steps/s", flush=True)
