"""Scan / loss-head / V-MPO / SAC kernels alone, at the named shape and at B*T >= 2^26: CUDA-event time and
achieved algorithmic GB/s vs the measured HBM peak.  Also the process ncu profiles for those kernels (dev tool).

    python tools/scan_bench.py            # JSON lines
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200"))
sys.path.insert(0, ROOT)
import torch

from b200rl import ops as O
from b200rl.engine import VMPOEngine, SACEngine
from b200rl.networks import models as M
from b200rl.synth import make_rollouts

HYPER = dict(hidden_size=64, gamma=0.99, lmbda=0.95, eps_clip=0.1, policy_loss_coef=1.0, value_loss_coef=0.5,
             entropy_coef=0.00005, lr=0.0001, max_grad_norm=40, alpha=0.2, v_mpo_lagrange_multiplier_init=5, coef_eta=0.01,
             coef_alpha_upper=0.01, coef_alpha_below=0.005, K_epoch=1)


ITERS = int(os.environ.get("SCAN_BENCH_ITERS", "0"))      # > 0: fixed small iteration count (under ncu)


def timeit(fn, iters=20):
    iters = ITERS or iters
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    ops = O.default()
    peak = 6501.2
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    g = torch.Generator(device="cuda").manual_seed(0)
    for B, T in ((4096, 128), (1 << 19, 128)):
        r = lambda *s: torch.randn(*s, device=dev, generator=g)
        rew, val, lp, lq = 0.1 * r(B, T, 1), r(B, T, 1), -r(B, T, 1).abs(), -r(B, T, 1).abs()
        fir = (torch.rand(B, T, 1, device=dev, generator=g) < 0.02).float()
        delta = r(B, T - 1, 1)
        n = B * T
        L2 = " (inputs 5x the 126 MB L2)" if n >= 1 << 26 else " (fits L2: latency / launch bound)"
        for name, fn, bytes_per in (
                ("gae_kernel", lambda: ops.compute_gae(delta, 0.99, 0.95), 8),
                ("td_gae_kernel", lambda: ops.td_gae(rew, fir, val, 0.99, 0.95), 20),
                ("vtrace_kernel", lambda: ops.compute_v_trace(lq, lp, fir, rew, val, 0.99), 32)):
            ms = timeit(fn)
            gbs = n * bytes_per / (ms * 1e-3) / 1e9
            print(json.dumps({"kernel": name, "B": B, "T": T, "elements": n, "ms": ms, "algorithmic_bytes_per_element": bytes_per,
                              "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak, "note": "includes the output allocation" + L2}))
        del rew, val, lp, lq, fir, delta
        torch.cuda.empty_cache()
    # one V-MPO update (config 4 shard and full) and one SAC-Continuous update (config 5 shard): the processes ncu profiles
    for B in (1024, 8192):
        torch.manual_seed(0)
        model = M.MlpLSTMSingle(4, 2, 128, 64).to(dev)
        eng = VMPOEngine(model.actor, HYPER)
        b = {k: v.to(dev) for k, v in make_rollouts(B, 128, 4, 2).items()}
        ms = timeit(lambda: eng.update(b, 0.0075), 10)
        print(json.dumps({"update": "V-MPO", "B": B, "T": 128, "ms": ms, "env_steps_per_s": B * 128 / (ms * 1e-3)}))
    model = M.MlpLSTMSeperateContinuous(17, 6, 5, 64, act_dim=6).to(dev)
    eng = SACEngine(model, HYPER, True)
    b = {k: v.to(dev) for k, v in make_rollouts(8192, 5, 17, 6, continuous=True, act_dim=6).items()}
    ms = timeit(lambda: eng.update(b), 10)
    print(json.dumps({"update": "SAC-Continuous", "B": 8192, "T": 5, "obs": 17, "act": 6, "ms": ms, "env_steps_per_s": 8192 * 5 / (ms * 1e-3)}))


if __name__ == "__main__":
    main()
