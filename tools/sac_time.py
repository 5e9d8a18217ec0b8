"""SAC-Continuous update time at BASELINE config 5's per-GPU shape, eager launches vs CUDA-graph phases (dev tool)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200"))
sys.path.insert(0, ROOT)
import torch


def main(B=8192, T=5, f=17, A=6, iters=50):
    from b200rl.engine import SACEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    from tools.kernel_times import HYPER
    dev = torch.device("cuda:0")
    hyper = dict(HYPER, alpha=0.1)
    out = {}
    for mode in ("0", "1"):
        os.environ["B200RL_GRAPH"] = mode
        torch.manual_seed(0)
        model = M.MlpLSTMSeperateContinuous(f, A, T, 64, act_dim=A).to(dev)
        eng = SACEngine(model, hyper, True)
        bs = [{k: v.to(dev) for k, v in make_rollouts(B, T, f, A, continuous=True, batch_index=i, act_dim=A).items()} for i in range(2)]
        for i in range(12):
            eng.update(bs[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            eng.update(bs[i % 2])
        e1.record()
        torch.cuda.synchronize()
        out["graph" if mode == "1" else "eager"] = e0.elapsed_time(e1) / iters
    print(json.dumps({"B": B, "T": T, "obs": f, "act": A, "ms_per_update": out}))


if __name__ == "__main__":
    main(*[int(x) for x in sys.argv[1:]])
