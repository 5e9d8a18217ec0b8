"""cProfile of bench.py's shared-memory e2e leg (dev tool): where do the host milliseconds go?"""
import cProfile
import io
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200"))
import torch
import torch.distributed as dist

import bench
from b200rl.synth import make_rollouts

dev = torch.device("cuda:0")
timed = bench.Timed(dev, 1, dist)
batch = make_rollouts(bench.B, bench.T, bench.F, bench.A)
pr = cProfile.Profile()
pr.enable()
ms, n, info = bench.e2e_from_shared_memory(dev, 1, dist, timed, batch, 30)
pr.disable()
print("ms/step", ms / n, info)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:6000])
