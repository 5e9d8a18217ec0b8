"""Per-kernel CUDA-event timings of one PPO update at a given shape (dev tool)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200"))
sys.path.insert(0, ROOT)
import torch

from b200rl import ops as O
from b200rl.engine import OnPolicyEngine
from b200rl.networks import models as M
from b200rl.synth import make_rollouts

HYPER = dict(hidden_size=64, gamma=0.99, lmbda=0.95, eps_clip=0.1, policy_loss_coef=1.0, value_loss_coef=0.5,
             entropy_coef=0.00005, lr=0.0001, max_grad_norm=40)


def main(B=4096, T=128, f=4, A=2, iters=20, algo="PPO"):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = M.MlpLSTMSingle(f, A, T, 64).to(dev)
    eng = OnPolicyEngine(model.actor, algo, HYPER)
    batch = {k: v.to(dev) for k, v in make_rollouts(B, T, f, A).items()}
    ops, L = eng.ops, eng.L
    for _ in range(3):
        eng.update(batch)
    torch.cuda.synchronize()
    u = batch["obs"]
    stages = [
        ("lstm_fwd", lambda: ops.lstm_fwd(L, eng.flat, u, batch["hx"], batch["cx"], eng.ws)),
        ("head_fwd", lambda: ops.head_fwd(L, eng.flat, eng.ws["h"], eng.head_out)),
        ("ppo_head", lambda: ops.ppo_head(eng.head_out, batch, eng.dhead, eng.scalars, HYPER, A, False, None)),
        ("lstm_bwd", lambda: ops.lstm_bwd(L, eng.flat, eng.dhead, eng.ws, batch["cx"])),
        ("lstm_wgrad", lambda: ops.lstm_wgrad(L, eng.flat, eng.ws, batch["hx"], u, eng.dhead, eng.partials, variant=eng.wgrad_variant)),
        ("reduce_partials", lambda: ops.reduce_partials(eng.partials, eng.grad)),
        ("clip_rmsprop", lambda: ops.clip_rmsprop(eng.flat, eng.grad, eng.sq_avg, L.P, 1.0, 40.0, 1e-4, 0.99, 1e-5, eng.grad_norm)),
    ]
    res = {}
    for name, fn in stages:
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / iters
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        eng.update(batch)
    e1.record()
    torch.cuda.synchronize()
    res["update_total"] = e0.elapsed_time(e1) / iters
    res["env_steps_per_s"] = B * T / (res["update_total"] * 1e-3)
    print(json.dumps({"B": B, "T": T, "algo": algo, "ms": res}))
    print(eng.losses())


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    main(*a)
