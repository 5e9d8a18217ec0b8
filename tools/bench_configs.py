"""Device-time of one learner update at every BASELINE.json config's per-GPU shape (dev tool; one JSON line each).

configs[0] PPO B=64 T=32 | configs[1] IMPALA B=4096 T=128 | configs[2] PPO-Continuous obs=2 act=1 B=2048 T=128 |
configs[3] V-MPO B=8192 over 8 GPUs -> 1024 per GPU, T=128 | configs[4] SAC-Continuous obs=17 act=6, 65536 transitions
over 8 GPUs -> 8192 per GPU = 1024 sequences x T=8.  Inputs resident in HBM, CUDA events, 3 warm-up + N timed updates."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200"))
sys.path.insert(0, ROOT)
import torch

from b200rl.engine import OnPolicyEngine, SACEngine, VMPOEngine
from b200rl.networks import models as M
from b200rl.synth import make_rollouts

# utils/parameters.json values
HYPER = dict(hidden_size=64, gamma=0.99, lmbda=0.95, eps_clip=0.1, policy_loss_coef=1.0, value_loss_coef=0.5,
             entropy_coef=0.00005, alpha=0.2, v_mpo_lagrange_multiplier_init=5, coef_eta=0.01, coef_alpha_upper=0.01,
             coef_alpha_below=0.005, K_epoch=1, lr=0.0001, max_grad_norm=40, reward_scale=0.1)


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    hp = dict(HYPER)
    rows = []

    def onpolicy(name, algo, B, T, f, A, cont=False):
        torch.manual_seed(0)
        model = (M.MlpLSTMSingleContinuous if cont else M.MlpLSTMSingle)(f, A, T, 64).to(dev)
        eng = OnPolicyEngine(model.actor, algo, hp)
        batch = {k: v.to(dev) for k, v in make_rollouts(B, T, f, A, continuous=cont).items()}
        ms = timed(lambda: eng.update(batch), 20)
        rows.append(dict(config=name, algo=algo, B=B, T=T, obs=f, act=A, ms_per_update=ms, env_steps_per_s=B * T / (ms * 1e-3)))

    onpolicy("configs[0] PPO CartPole batch=64 T=32", "PPO", 64, 32, 4, 2)
    onpolicy("north_star / bench.py PPO CartPole B=4096 T=128", "PPO", 4096, 128, 4, 2)
    onpolicy("configs[1] IMPALA CartPole B=4096 T=128", "IMPALA", 4096, 128, 4, 2)
    onpolicy("configs[2] PPO-Continuous obs=2 act=1 B=2048 T=128", "PPO-Continuous", 2048, 128, 2, 1, cont=True)

    torch.manual_seed(0)
    B, T = 1024, 128
    model = M.MlpLSTMSingle(4, 2, T, 64).to(dev)
    eng = VMPOEngine(model.actor, hp)
    batch = {k: v.to(dev) for k, v in make_rollouts(B, T, 4, 2).items()}
    ms = timed(lambda: eng.update(batch, 0.0075), 20)
    rows.append(dict(config="configs[3] V-MPO top-k=50% B=8192/8 GPUs = 1024 per GPU, T=128", algo="V-MPO", B=B, T=T, obs=4, act=2,
                     ms_per_update=ms, env_steps_per_s=B * T / (ms * 1e-3)))

    torch.manual_seed(0)
    B, T, f, A = 1024, 8, 17, 6
    model = M.MlpLSTMSeperateContinuous(f, A, T, 64, act_dim=A).to(dev)
    eng = SACEngine(model, hp, True)
    batch = {k: v.to(dev) for k, v in make_rollouts(B, T, f, A, continuous=True, act_dim=A).items()}
    g = torch.Generator(device=dev).manual_seed(1)
    noise = [torch.randn(B, T, A, generator=g, device=dev), torch.randn(B, T, A, generator=g, device=dev)]
    ms = timed(lambda: eng.update(batch, noise=noise), 20)
    rows.append(dict(config="configs[4] SAC-Continuous obs=17 act=6, 65536 transitions/8 GPUs = 1024 seq x T=8 per GPU", algo="SAC-Continuous",
                     B=B, T=T, obs=f, act=A, ms_per_update=ms, env_steps_per_s=B * T / (ms * 1e-3)))
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
