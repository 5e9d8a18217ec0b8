"""Where the e2e step goes on the host side (dev tool): H2D bandwidth from pinned memory, CPU time to enqueue one update."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200"))
import torch

from b200rl.engine import OnPolicyEngine
from b200rl.networks import models as M
from b200rl.synth import make_rollouts

HYPER = dict(hidden_size=64, gamma=0.99, lmbda=0.95, eps_clip=0.1, policy_loss_coef=1.0, value_loss_coef=0.5,
             entropy_coef=0.00005, lr=0.0001, max_grad_norm=40)
dev = torch.device("cuda:0")
for mb in (2.1, 8.4, 18.9, 64):
    n = int(mb * 1e6 / 4)
    h = torch.empty(n).pin_memory()
    d = torch.empty(n, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"H2D pinned {mb:5.1f} MB: {ms:.3f} ms  {mb / ms:.1f} GB/s")
B, T = 4096, 128
model = M.MlpLSTMSingle(4, 2, T, 64).to(dev)
eng = OnPolicyEngine(model.actor, "PPO", HYPER)
batch = {k: v.to(dev) for k, v in make_rollouts(B, T, 4, 2).items()}
for _ in range(3):
    eng.update(batch)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    eng.update(batch)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue of one update: {(t1 - t0) / 20 * 1e3:.3f} ms CPU  (20 updates drained after {(t2 - t0) * 1e3:.1f} ms)")
t0 = time.perf_counter()
for _ in range(20):
    h = eng.snapshot_async()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"snapshot_async: {(t1 - t0) / 20 * 1e3:.3f} ms CPU")
t0 = time.perf_counter()
for _ in range(20):
    flat, L = eng.snapshot_wait(h)
    sd = model.actor.state_dict_from_flat(flat)
t1 = time.perf_counter()
print(f"snapshot_wait + state_dict_from_flat: {(t1 - t0) / 20 * 1e3:.3f} ms CPU")
