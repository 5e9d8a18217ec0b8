#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference in-process.

Run in the build container only (needs /root/reference):

    python -B tests/golden/make_golden.py

It imports `/root/reference` with the one stub the survey found necessary
(`blosc2`), drives the reference's verbatim `learning()` coroutines
through a fake `parent`, and writes small `.npz` files next to this script.  The
fixtures travel to the GPU box; the reference does not.  Nothing here is imported
by the product or by the tests -- the tests only read the `.npz` files.
"""
import asyncio
import contextlib
import copy
import io
import math
import os
import sys
import tempfile
import types
from functools import partial
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

sys.dont_write_bytecode = True
sys.modules["blosc2"] = types.ModuleType("blosc2")
os.chdir(REF)
sys.path.insert(0, REF)
sys.path.insert(1, os.path.join(REPO, "pytorch-distributed-reinforcement-learning_b200"))

import torch  # noqa: E402
from torch.optim import Adam, RMSprop  # noqa: E402
from torch.distributions import Uniform  # noqa: E402

from utils.utils import Params, ExecutionTimer, make_gpu_batch  # noqa: E402  (reference)
from agents.learner_module import compute_loss as ref_cl  # noqa: E402
from agents.learner_module.ppo.learning import learning as learning_ppo  # noqa: E402
from agents.learner_module.impala.learning import learning as learning_impala  # noqa: E402
from agents.learner_module.v_mpo.learning import learning as learning_vmpo  # noqa: E402
from agents.learner_module.sac.learning import learning as learning_sac  # noqa: E402
from agents.learner_module.sac_continuous.learning import learning as learning_sacc  # noqa: E402
from networks import models as ref_models  # noqa: E402

from b200rl.synth import make_rollouts, as_reference_tuple  # noqa: E402  (ours: seeded inputs only)


class StopAfter:
    def __init__(self, n):
        self.n, self.calls = n, 0

    def is_set(self):
        self.calls += 1
        return self.calls > self.n


def base_args(algo, B, T, A, tmp, overrides=None):
    a = copy.copy(Params)
    for k, v in (overrides or {}).items():         # hyper-parameter overrides of a case (stored as hyper/<key> in the fixture)
        assert hasattr(a, k), k
        setattr(a, k, v)
    a.algo, a.batch_size, a.seq_len, a.action_space = algo, B, T, A
    a.device = torch.device("cpu")
    a.model_dir = tmp
    a.loss_log_interval = 1
    a.model_save_interval = 10 ** 9          # still fires at idx 0 (ppo/learning.py:113)
    return a


def run_learning(fn, parent, n, T, B):
    timer = ExecutionTimer(num_transition=B * T)

    async def go():
        q = asyncio.Queue()
        for b in parent._batches:
            await q.put(b)
        parent.batch_queue = q
        parent.batch_buffer = q
        await fn(parent, timer)

    with contextlib.redirect_stdout(io.StringIO()):
        asyncio.run(go())


def sd_np(sd, prefix):
    return {f"{prefix}/{k}": v.detach().numpy().copy() for k, v in sd.items()}


def on_policy_case(name, algo, B, T, f, A, n_steps, continuous=False, seed=0, overrides=None):
    torch.manual_seed(seed)
    cls = ref_models.MlpLSTMSingleContinuous if continuous else ref_models.MlpLSTMSingle
    model = cls(f, A, T, Params.hidden_size)
    out = {"meta": np.array([B, T, f, A, Params.hidden_size, n_steps, int(continuous)])}
    for k, v in (overrides or {}).items():
        out[f"hyper/{k}"] = np.array(v, dtype=np.float64)
    out.update(sd_np(model.actor.state_dict(), "sd0"))
    batches = [make_rollouts(B, T, f, A, continuous=continuous, batch_index=i) for i in range(n_steps)]
    for i, b in enumerate(batches):
        for k, v in b.items():
            out[f"in/{i}/{k}"] = v.numpy()

    # direct calls on the initial weights (intermediates the learning loop does not expose)
    b0 = batches[0]
    with torch.no_grad():
        logits, logp, ent, value = model.actor(b0["obs"], (b0["hx"], b0["cx"]), b0["act"])
        out["fwd/logits"], out["fwd/log_probs"] = logits.numpy(), logp.numpy()
        out["fwd/entropy"], out["fwd/value"] = ent.numpy(), value.numpy()
        td = b0["rew"][:, :-1] + Params.gamma * (1 - b0["is_fir"][:, 1:]) * value[:, 1:]
        out["fwd/td_target"] = td.numpy()
        out["fwd/gae"] = ref_cl.compute_gae(td - value[:, :-1], Params.gamma, Params.lmbda).numpy()
        if not continuous:
            rho, adv, vs = ref_cl.compute_v_trace(b0["log_prob"], logp, b0["is_fir"], b0["rew"], value, Params.gamma)
            out["fwd/vt_rho"], out["fwd/vt_adv"], out["fwd/vt_vs"] = rho.numpy(), adv.numpy(), vs.numpy()
            out["fwd/kl"] = ref_cl.kldivergence(b0["logits"][:, :-1], logits[:, :-1]).numpy()

    with tempfile.TemporaryDirectory() as tmp:
        args = base_args(algo, B, T, A, tmp, overrides)
        parent = SimpleNamespace(args=args, model=model, stop_event=StopAfter(n_steps), heartbeat=None,
                                 to_gpu=partial(make_gpu_batch, device=args.device), device=args.device)
        parent._batches = [as_reference_tuple(b, T) for b in batches]
        parent.optimizer = RMSprop(model.parameters(), lr=args.lr, eps=1e-5)          # learner.py:70
        logs, pubs = [], []
        parent.pub_model = lambda sd: pubs.append({k: v.detach().clone() for k, v in sd.items()})
        parent.log_loss_tensorboard = lambda timer, loss, dl: logs.append((loss.detach().clone(), dict(dl)))
        fn = {"PPO": learning_ppo, "PPO-Continuous": learning_ppo, "IMPALA": learning_impala,
              "V-MPO": learning_vmpo}[algo]
        if algo == "V-MPO":                                                            # learner.py:320-348
            parent.log_eta = torch.tensor(np.log(args.v_mpo_lagrange_multiplier_init), requires_grad=True)
            parent.log_alpha = torch.tensor(np.log(args.v_mpo_lagrange_multiplier_init), requires_grad=True)
            parent.optimizer = RMSprop([{"params": model.parameters()}, {"params": parent.log_eta},
                                        {"params": parent.log_alpha}], lr=args.lr, eps=1e-5)
            torch.manual_seed(seed + 7)
            coefs = [Uniform(torch.tensor(args.coef_alpha_below).log(), torch.tensor(args.coef_alpha_upper).log()
                             ).sample().exp() for _ in range(n_steps)]
            out["coef_alpha"] = np.array([float(c) for c in coefs], dtype=np.float32)
            it = iter(coefs)
            parent.get_coef_alpha = lambda: next(it)
            duals = []
            orig_pub = parent.pub_model
            parent.pub_model = lambda sd: (orig_pub(sd), duals.append((parent.log_eta.item(), parent.log_alpha.item())))
        run_learning(fn, parent, n_steps, T, B)
        assert len(logs) == n_steps and len(pubs) == n_steps
        out["out/loss"] = np.array([float(l) for l, _ in logs], dtype=np.float64)
        for key in logs[0][1]:
            if key == "ratio":
                for i, (_, dl) in enumerate(logs):
                    out[f"out/ratio/{i}"] = dl["ratio"].numpy()
            else:
                out[f"out/{key}"] = np.array([float(dl[key]) for _, dl in logs], dtype=np.float64)
        for i, sd in enumerate(pubs):
            out[f"out/params/{i}"] = torch.cat([v.reshape(-1) for v in sd.values()]).numpy()
        if algo == "V-MPO":
            out["out/duals"] = np.array(duals, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", out["out/loss"])


def sac_case(name, continuous, B, T, f, A, n_steps, seed=0, overrides=None):
    torch.manual_seed(seed)
    cls = ref_models.MlpLSTMSeperateContinuous if continuous else ref_models.MlpLSTMSeperate
    model = cls(f, A, T, Params.hidden_size)
    out = {"meta": np.array([B, T, f, A, Params.hidden_size, n_steps, int(continuous)])}
    for k, v in (overrides or {}).items():
        out[f"hyper/{k}"] = np.array(v, dtype=np.float64)
    out.update(sd_np(model.actor.state_dict(), "actor0"))
    out.update(sd_np(model.critic.state_dict(), "critic0"))
    batches = [make_rollouts(B, T, f, A, continuous=continuous, batch_index=i, act_dim=1) for i in range(n_steps)]
    for i, b in enumerate(batches):
        for k, v in b.items():
            out[f"in/{i}/{k}"] = v.numpy()

    noise = []
    import torch.distributions.normal as tdn
    orig = tdn._standard_normal

    def rec(shape, dtype, device):
        e = orig(shape, dtype=dtype, device=device)
        noise.append(e.clone())
        return e

    with tempfile.TemporaryDirectory() as tmp:
        args = base_args("SAC-Continuous" if continuous else "SAC", B, T, A, tmp, overrides)
        dev = args.device
        parent = SimpleNamespace(args=args, model=model, stop_event=StopAfter(n_steps), heartbeat=None,
                                 to_gpu=partial(make_gpu_batch, device=dev), device=dev)
        parent._batches = [as_reference_tuple(b, T) for b in batches]
        # learner.py:351-367
        parent.actor = model.actor.to(dev)
        parent.critic = model.critic.to(dev)
        parent.target_critic = model.critic.to(dev)
        parent.target_critic.load_state_dict(parent.critic.state_dict())
        parent.actor_optimizer = Adam(parent.actor.parameters(), lr=args.lr)
        parent.critic_optimizer = Adam(parent.critic.parameters(), lr=args.lr)
        parent.target_entropy = args.action_space
        parent.log_alpha = torch.tensor(np.log(args.alpha), requires_grad=True)
        parent.log_alpha_optimizer = Adam([parent.log_alpha], lr=args.lr)
        logs, pubs, crit, la = [], [], [], []
        parent.pub_model = lambda sd: (pubs.append({k: v.detach().clone() for k, v in sd.items()}),
                                       crit.append({k: v.detach().clone() for k, v in parent.critic.state_dict().items()}),
                                       la.append(parent.log_alpha.item()))
        parent.log_loss_tensorboard = lambda timer, loss, dl: logs.append((loss.detach().clone(), dict(dl)))
        torch.manual_seed(seed + 11)
        tdn._standard_normal = rec
        try:
            run_learning(learning_sacc if continuous else learning_sac, parent, n_steps, T, B)
        finally:
            tdn._standard_normal = orig
        assert len(logs) == n_steps
        out["out/loss"] = np.array([float(l) for l, _ in logs], dtype=np.float64)
        for key in ("policy-loss", "value-loss", "loss_alpha", "alpha"):
            out[f"out/{key}"] = np.array([float(dl[key]) for _, dl in logs], dtype=np.float64)
        for i in range(n_steps):
            out[f"out/actor_params/{i}"] = torch.cat([v.reshape(-1) for v in pubs[i].values()]).numpy()
            out[f"out/critic_params/{i}"] = torch.cat([v.reshape(-1) for v in crit[i].values()]).numpy()
        out["out/log_alpha"] = np.array(la, dtype=np.float64)
        if continuous:
            assert len(noise) == 2 * n_steps
            for i, e in enumerate(noise):
                out[f"noise/{i // 2}/{i % 2}"] = e.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", out["out/loss"])


def init_case():
    """Weights the reference constructors produce under torch.manual_seed(0) (pins oracle.init_params)."""
    out = {}
    torch.manual_seed(0)
    out.update(sd_np(ref_models.MlpLSTMSingle(4, 2, 5, 64).actor.state_dict(), "single"))
    torch.manual_seed(0)
    out.update(sd_np(ref_models.MlpLSTMSingleContinuous(2, 1, 5, 64).actor.state_dict(), "single_continuous"))
    torch.manual_seed(0)
    m = ref_models.MlpLSTMSeperate(4, 2, 5, 64)
    out.update(sd_np(m.actor.state_dict(), "sep_actor"))
    out.update(sd_np(m.critic.state_dict(), "sep_critic"))
    torch.manual_seed(0)
    m = ref_models.MlpLSTMSeperateContinuous(2, 1, 5, 64)
    out.update(sd_np(m.actor.state_dict(), "sepc_actor"))
    out.update(sd_np(m.critic.state_dict(), "sepc_critic"))
    # keep only a strided sample of each tensor + its sum: enough to pin the RNG consumption order
    slim = {}
    for k, v in out.items():
        flat = v.reshape(-1)
        slim[k + "#shape"] = np.array(v.shape)
        slim[k + "#head"] = flat[:8].copy()
        slim[k + "#sum"] = np.array(flat.astype(np.float64).sum())
    np.savez_compressed(os.path.join(HERE, "init_seed0.npz"), **slim)
    print("init_seed0 ok")


def assembler_case():
    """A random stream of per-step worker records through the reference's RolloutAssembler + LearnerStorage.make_batch
    (buffers/rollout_assembler.py:24-82, agents/learner_storage.py:123-159) under a scripted clock: the stream and the shared
    arrays it leaves behind pin the native assembler (csrc/assembler.cu)."""
    import buffers.rollout_assembler as ra
    from agents.learner_storage import LearnerStorage
    T, f, A, Hh, cap = 6, 4, 2, 64, 5
    args = copy.copy(Params)
    args.seq_len, args.action_space, args.hidden_size = T, A, Hh
    clock = SimpleNamespace(t=0.0)
    ra.time = SimpleNamespace(monotonic=lambda: clock.t)            # the module calls time.monotonic()
    rng = np.random.RandomState(7)
    ready = asyncio.Queue()
    asm = ra.RolloutAssembler(args, ready)
    widths = dict(obs=f, act=1, rew=1, logits=A, log_prob=1, is_fir=1, hx=Hh, cx=Hh)
    store = SimpleNamespace(args=args, obs_shape=[f], sh_data_num=SimpleNamespace(value=0))
    for k, w in widths.items():
        setattr(store, f"sh_{k}_batch", np.zeros(T * cap * w, dtype=np.float32))
    rec = {k: [] for k in ("id", "done", "now", *widths)}
    counts = []
    live = {}                                     # episode id -> steps left
    next_id = 100

    async def run():
        nonlocal next_id
        for step in range(140):
            clock.t += float(rng.choice([0.01, 0.02, 0.05, 0.3], p=[0.5, 0.3, 0.15, 0.05]))
            if len(live) < 3 or rng.rand() < 0.1:
                live[next_id] = int(rng.randint(2, 11))
                next_id += 1
            eid = int(rng.choice(list(live)))
            live[eid] -= 1
            done = live[eid] == 0
            if done:
                del live[eid]
            data = {k: torch.from_numpy(rng.randn(w).astype(np.float32)) for k, w in widths.items()}
            data["is_fir"] = torch.FloatTensor([1.0 if rng.rand() < 0.1 else 0.0])
            for k in widths:
                rec[k].append(data[k].numpy().copy())
            rec["id"].append(eid); rec["done"].append(done); rec["now"].append(clock.t)
            data["id"], data["done"] = eid, torch.FloatTensor([1.0 if done else 0.0])     # (agents/worker.py:108-121)
            with contextlib.redirect_stdout(io.StringIO()):
                await asm.push(data)
            while not ready.empty():
                LearnerStorage.make_batch(store, ready.get_nowait())
            counts.append(store.sh_data_num.value)
            if step == 90:                        # the learner takes the batch: the counter goes back to zero once
                store.sh_data_num.value = 0
    asyncio.run(run())
    out = {"meta": np.array([T, f, A, Hh, cap, 90]), "id": np.array(rec["id"]), "done": np.array(rec["done"]), "now": np.array(rec["now"]),
           "count_after": np.array(counts)}
    for k in widths:
        out[f"in/{k}"] = np.stack(rec[k])
        out[f"shared/{k}"] = getattr(store, f"sh_{k}_batch")
    np.savez_compressed(os.path.join(HERE, "assembler_stream.npz"), **out)
    print("assembler_stream: slots written after each push", counts[-1], "max", max(counts))


if __name__ == "__main__" and "--assembler" in sys.argv:
    assembler_case()
    sys.exit(0)

if __name__ == "__main__" and "--branches" in sys.argv:
    # round 2: branches the first fixtures never reached -- an ACTIVE clip_grad_norm_ (coef < 1) under RMSprop and under Adam,
    # and K_epoch = 2 (two updates per batch)
    on_policy_case("ppo_clip", "PPO", 8, 6, 4, 2, 3, overrides={"max_grad_norm": 1e-3})
    on_policy_case("ppo_kepoch2", "PPO", 8, 6, 4, 2, 2, overrides={"K_epoch": 2})
    sac_case("sac_clip", False, 8, 5, 4, 2, 2, overrides={"max_grad_norm": 1e-2})
    sys.exit(0)

if __name__ == "__main__":
    on_policy_case("ppo_small", "PPO", 8, 6, 4, 2, 3)
    on_policy_case("ppo_cfg1", "PPO", 64, 32, 4, 2, 2)
    on_policy_case("ppo_repo_default", "PPO", 128, 5, 4, 2, 2)
    on_policy_case("impala_small", "IMPALA", 8, 6, 4, 2, 3)
    on_policy_case("impala_mid", "IMPALA", 32, 16, 4, 2, 2)
    on_policy_case("ppo_continuous_small", "PPO-Continuous", 8, 6, 2, 1, 3, continuous=True)
    on_policy_case("vmpo_small", "V-MPO", 8, 6, 4, 2, 3)
    on_policy_case("vmpo_odd", "V-MPO", 9, 5, 4, 2, 2)
    sac_case("sac_small", False, 8, 5, 4, 2, 3)
    sac_case("sac_continuous_small", True, 8, 5, 2, 1, 3)
    init_case()
