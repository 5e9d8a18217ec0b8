"""CUDA-event time of the weight-gradient stage exactly as the PPO engine launches it (dev tool).  argv: B T iters"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200"))
sys.path.insert(0, ROOT)
import torch

from b200rl.engine import OnPolicyEngine
from b200rl.networks import models as M
from b200rl.synth import make_rollouts
from tools.kernel_times import HYPER


def main(B=4096, T=128, iters=20):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = M.MlpLSTMSingle(4, 2, T, 64).to(dev)
    eng = OnPolicyEngine(model.actor, "PPO", HYPER)
    batch = {k: v.to(dev) for k, v in make_rollouts(B, T, 4, 2).items()}
    for _ in range(3):
        eng.update(batch)
    torch.cuda.synchronize()
    fn = lambda: eng.ops.lstm_wgrad(eng.L, eng.flat, eng.ws, batch["hx"], batch["obs"], eng.dhead, eng.partials, variant=eng.wgrad_variant)
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"B": B, "T": T, "lstm_wgrad_ms": e0.elapsed_time(e1) / iters, "fused_small_grads": eng.ws.get("small_grads_done")}))


if __name__ == "__main__":
    main(*[int(x) for x in sys.argv[1:]])
