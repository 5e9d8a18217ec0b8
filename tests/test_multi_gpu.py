"""Data-parallel learner: B sharded over ranks, one all-reduce of the flat gradient (PPO / IMPALA),
all-gather of the advantage matrix + global top-k (V-MPO), actor / critic / temperature reductions (SAC).  world_size=2 with gloo on CPU drives the
host logic through the test-only emulated kernels; the same check runs over NCCL when >= 2 GPUs exist."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, PKG


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, algo, backend, B, T, out_path):
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle
    from b200rl import native, ops as O
    from b200rl.engine import OnPolicyEngine, VMPOEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    if backend == "gloo":
        dev = torch.device("cpu")
        lib = native.Lib(os.path.join(ROOT, "tests", "emul", "_build", "libb200rl_emul.so"), emulated=True)
        ops = O.Ops(lib, allow_cpu=True)
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    else:
        dev = torch.device("cuda", rank)
        torch.cuda.set_device(dev)
        ops = O.default()
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(0)
    model = M.MlpLSTMSingle(4, 2, T, 64).to(dev)
    full = make_rollouts(B, T, 4, 2)
    sh = B // world
    shard = {k: v[rank * sh:(rank + 1) * sh].contiguous().to(dev) for k, v in full.items()}
    if algo == "V-MPO":
        eng = VMPOEngine(model.actor, oracle.DEFAULT_HYPER, ops=ops, process_group=dist.group.WORLD)
        for _ in range(2):
            eng.update(shard, 0.0075)
    else:
        eng = OnPolicyEngine(model.actor, algo, oracle.DEFAULT_HYPER, ops=ops, process_group=dist.group.WORLD)
        for _ in range(2):
            eng.update(shard)
    L = eng.losses()
    params = torch.cat([v.reshape(-1) for v in model.actor.views_of(eng.flat).values()]).cpu()
    if rank == 0:
        np.savez(out_path, params=params.numpy(), loss=L["loss"])
    dist.barrier()
    dist.destroy_process_group()


def _reference_run(algo, B, T):
    import oracle
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    torch.manual_seed(0)
    model = M.MlpLSTMSingle(4, 2, T, 64)
    sd = {k: v.clone() for k, v in model.actor.state_dict().items()}
    orc = oracle.OracleLearner(algo, sd)
    full = make_rollouts(B, T, 4, 2)
    for _ in range(2):
        r = orc.update(full, coef_alpha=torch.tensor(0.0075)) if algo == "V-MPO" else orc.update(full)
    return r


def _run(algo, backend, B, T, tmp_path):
    out = str(tmp_path / f"dp_{algo}.npz")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, algo, backend, B, T, out), nprocs=2, join=True)
    got = np.load(out)
    want = _reference_run(algo, B, T)
    np.testing.assert_allclose(got["loss"], float(want["loss"]), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(got["params"], want["params"].numpy(), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("algo", ["PPO", "IMPALA", "V-MPO"])
def test_dp2_gloo_equals_full_batch(algo, tmp_path):
    from conftest import _emul_ops
    _emul_ops()                                   # builds the emulated library once, before spawning
    _run(algo, "gloo", 8, 6, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("algo", ["PPO", "V-MPO"])
def test_dp2_nccl_equals_full_batch(algo, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(algo, "nccl", 64, 16, tmp_path)


# ------------------------------------------------------------------------------------------------ SAC / SAC-Continuous
def _sac_setup(cont, B, T):
    from collections import OrderedDict
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    f, A = (5, 2) if cont else (4, 3)
    torch.manual_seed(3)
    model = (M.MlpLSTMSeperateContinuous(f, A, T, 64, act_dim=A) if cont else M.MlpLSTMSeperate(f, A, T, 64))
    pa = OrderedDict((k, v.clone()) for k, v in model.actor.state_dict().items())
    pc = OrderedDict((k, v.clone()) for k, v in model.critic.state_dict().items())
    batches = [make_rollouts(B, T, f, A, continuous=cont, batch_index=i, **({"act_dim": A} if cont else {})) for i in range(2)]
    g = torch.Generator().manual_seed(11)
    noise = [[torch.randn(B, T, A, generator=g), torch.randn(B, T, A, generator=g)] for _ in range(2)] if cont else [None, None]
    return model, pa, pc, batches, noise


def _sac_worker(rank, world, port, cont, B, T, out_path):
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle
    from b200rl import native, ops as O
    from b200rl.engine import SACEngine
    lib = native.Lib(os.path.join(ROOT, "tests", "emul", "_build", "libb200rl_emul.so"), emulated=True)
    ops = O.Ops(lib, allow_cpu=True)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    model, _, _, batches, noise = _sac_setup(cont, B, T)
    eng = SACEngine(model, oracle.DEFAULT_HYPER, cont, ops=ops, process_group=dist.group.WORLD)
    sh = B // world
    cut = lambda t: t[rank * sh:(rank + 1) * sh].contiguous()
    for b, nz in zip(batches, noise):
        eng.update({k: cut(v) for k, v in b.items()}, noise=None if nz is None else [cut(e) for e in nz])
    L = eng.losses()
    if rank == 0:
        np.savez(out_path, actor=torch.cat([v.reshape(-1) for v in model.actor.state_dict().values()]).numpy(),
                 critic=torch.cat([v.reshape(-1) for v in model.critic.state_dict().values()]).numpy(),
                 loss=L["loss"], log_alpha=L["log_alpha"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cont", [False, True])
def test_sac_dp2_gloo_equals_full_batch(cont, tmp_path):
    """Two ranks on half the replay batch each == one learner on the whole batch (actor, twin critics, temperature)."""
    import oracle
    from conftest import _emul_ops
    _emul_ops()
    B, T = 8, 5
    out = str(tmp_path / "dp_sac.npz")
    mp.spawn(_sac_worker, args=(2, _free_port(), cont, B, T, out), nprocs=2, join=True)
    got = np.load(out)
    _, pa, pc, batches, noise = _sac_setup(cont, B, T)
    orc = oracle.OracleLearner("SAC-Continuous" if cont else "SAC", {"actor": pa, "critic": pc})
    for b, nz in zip(batches, noise):
        r = orc.update(b, noise=nz) if cont else orc.update(b)
    np.testing.assert_allclose(got["loss"], float(r["loss"]), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(got["log_alpha"], float(r["log_alpha"]), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(got["actor"], r["actor_params"].numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(got["critic"], r["critic_params"].numpy(), rtol=1e-4, atol=1e-6)
