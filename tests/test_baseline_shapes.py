"""CUDA path vs the oracle AT the shapes BASELINE.json names (not only at fixture sizes): headline PPO 4096x128, config 2
IMPALA 4096x128, config 3 PPO-Continuous 2048x128 (obs 2, act 1), config 4 V-MPO 8192x128 (full batch on one GPU and as
eight 1024-sequence shards), config 5 SAC-Continuous obs 17 / act 6 at the survey's per-GPU shape (8192 sequences x T=5).
One oracle update costs 5-10 s of CPU at these sizes, so the oracle runs live in the test.  Tolerances: rtol 1e-4 on
advantages / losses / parameters (BASELINE.json), gradient element-wise at rtol 1e-3 + 2e-5 max|g| and norm-wise 2e-4
(split-K, the 296-slab partial reduction, fp64 loss sums over 5e5 terms and the 3xTF32 / bf16 split terms are all in play
here and nowhere at fixture sizes)."""
import math
import threading
from collections import OrderedDict

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
PARAMS = dict(rtol=1e-4, atol=1e-6)


def _close(a, b, rtol=1e-4, atol=1e-6, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(1.0, float(np.abs(b).max())) if b.size else 1.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol * scale, err_msg=what)


def _flat_sd_order(net, flat_like):
    return torch.cat([v.reshape(-1) for v in net.views_of(flat_like).values()]).cpu()


def _grad_check(got, want, what):
    got, want = got.double(), want.double()
    _close(got, want, rtol=1e-3, atol=2e-5 * float(want.abs().max()), what=what + " (element-wise)")
    rel = float((got - want).norm() / want.norm())
    assert rel < 2e-4, f"{what}: differs from the oracle's autograd gradient by {rel:.2e} in norm"
    return rel


ON_POLICY = [
    ("PPO", 4096, 128, 4, 2, False),              # headline / north_star shape
    ("IMPALA", 4096, 128, 4, 2, False),           # BASELINE config 2
    ("PPO-Continuous", 2048, 128, 2, 1, True),    # BASELINE config 3 (MountainCarContinuous: obs 2, act 1)
]


@pytest.mark.parametrize("algo,B,T,f,A,cont", ON_POLICY, ids=[c[0] for c in ON_POLICY])
def test_on_policy_update_matches_oracle_at_baseline_shape(algo, B, T, f, A, cont):
    from b200rl.engine import OnPolicyEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = (M.MlpLSTMSingleContinuous if cont else M.MlpLSTMSingle)(f, A, T, 64)
    sd = OrderedDict((k, v.clone()) for k, v in model.actor.state_dict().items())
    model = model.to(dev)
    eng = OnPolicyEngine(model.actor, algo, oracle.DEFAULT_HYPER)
    eng.debug_outs = {}
    orc = oracle.OracleLearner(algo, sd)
    for i in range(2):                                     # two updates: the second starts from updated weights + RMSprop state
        batch = make_rollouts(B, T, f, A, continuous=cont, batch_index=i)
        r = orc.update(batch)
        eng.update({k: v.to(dev) for k, v in batch.items()})
        L = eng.losses()
        d = eng.debug_outs
        if algo == "IMPALA":
            _close(d["advantages"].cpu(), r["advantages"], what="advantages")
            _close(d["values_target"].cpu(), r["values_target"], what="values_target")
        else:
            _close(d["gae"].cpu(), r["gae"], what="gae")
            _close(d["td_target"].cpu(), r["td_target"], what="td_target")
        _close(d["ratio"].cpu(), r["ratio"], what="ratio")
        _close(d["log_probs"].cpu(), r["log_probs"], what="log_probs")
        for k in ("policy-loss", "value-loss", "policy-entropy", "loss"):
            _close(L[k], float(r[k]), rtol=1e-4, atol=1e-7, what=f"{k}[{i}]")
        _close(L["grad-norm"], float(r["grad_norm"]), rtol=2e-4, atol=1e-9, what="grad-norm")
        _grad_check(_flat_sd_order(model.actor, eng.grad), r["grad"], f"{algo} gradient[{i}]")
        _close(_flat_sd_order(model.actor, eng.flat), r["params"], what=f"params[{i}]", **PARAMS)


def _vmpo_reference(B, T, n_updates, coefs):
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    torch.manual_seed(0)
    model = M.MlpLSTMSingle(4, 2, T, 64)
    sd = OrderedDict((k, v.clone()) for k, v in model.actor.state_dict().items())
    orc = oracle.OracleLearner("V-MPO", sd)
    outs = []
    for i in range(n_updates):
        outs.append(orc.update(make_rollouts(B, T, 4, 2, batch_index=i), coef_alpha=torch.tensor(coefs[i])))
    return sd, outs


_VMPO_CACHE = {}


def _vmpo_cfg4():
    if "r" not in _VMPO_CACHE:
        _VMPO_CACHE["r"] = _vmpo_reference(8192, 128, 2, [0.0075, 0.006])
    return _VMPO_CACHE["r"]


def test_vmpo_full_batch_matches_oracle_at_cfg4_shape():
    """BASELINE config 4: V-MPO, 8192 sequences x T=128, top-k = 50 % per timestep over the whole batch -- on one GPU."""
    from b200rl.engine import VMPOEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    dev = torch.device("cuda:0")
    B, T = 8192, 128
    sd, want = _vmpo_cfg4()
    model = M.MlpLSTMSingle(4, 2, T, 64)
    model.actor.load_state_dict(sd)
    model = model.to(dev)
    eng = VMPOEngine(model.actor, oracle.DEFAULT_HYPER)
    for i, ca in enumerate([0.0075, 0.006]):
        r = want[i]
        eng.update({k: v.to(dev) for k, v in make_rollouts(B, T, 4, 2, batch_index=i).items()}, ca)
        L = eng.losses()
        assert L["selected"] == (B // 2) * (T - 1)
        _close(eng.gae.cpu().unsqueeze(-1), r["gae"], what="gae")
        for k in ("policy-loss", "value-loss", "loss-temperature", "loss-alpha", "loss"):
            _close(L[k], float(r[k]), rtol=1e-4, atol=1e-7, what=f"{k}[{i}]")
        _grad_check(_flat_sd_order(model.actor, eng.grad), r["grad"], f"V-MPO gradient[{i}]")
        _close(_flat_sd_order(model.actor, eng.flat), r["params"], what=f"params[{i}]", **PARAMS)
        _close([L["log_eta"], L["log_alpha"]], [float(r["log_eta"]), float(r["log_alpha"])], rtol=1e-6, atol=1e-9, what="duals")


def test_vmpo_eight_shards_match_oracle_at_cfg4_shape():
    """The same update as eight 1024-sequence shards (the 8-GPU partition of config 4), each "rank" a thread with its own
    stream and model copy on this one GPU: all-gather of the advantage matrix, global per-timestep top-k, and the fused
    peer step (one-shot all-reduce + clip + RMSprop) must reproduce the full-batch oracle."""
    from b200rl import comm as CM, native
    from b200rl.engine import VMPOEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    dev = torch.device("cuda:0")
    native.get().call("b200rl_peer_set_timeout_ms", 60000)
    world, B, T = 8, 8192, 128
    sd, want = _vmpo_cfg4()
    coefs = [0.0075, 0.006]
    fulls = [make_rollouts(B, T, 4, 2, batch_index=i) for i in range(2)]
    W = CM.LoopbackWorld(world, dev)
    out, errs = [None] * world, []

    warm = threading.Lock()

    def rank_main(r):
        try:
            torch.cuda.set_device(dev)
            with torch.cuda.stream(torch.cuda.Stream(dev)):
                sh = B // world
                shards = [{k: v[r * sh:(r + 1) * sh].contiguous().to(dev) for k, v in fulls[i].items()} for i in range(2)]
                # (one process, one device: allocations / lazy kernel loads synchronise the device and would wait for a peer
                # kernel that spins -- see test_peer_step_loopback_equals_full_batch: solo warm-up first, then start together)
                with warm:
                    scratch = M.MlpLSTMSingle(4, 2, T, 64).to(dev)
                    VMPOEngine(scratch.actor, oracle.DEFAULT_HYPER).update(shards[0], coefs[0])
                    torch.cuda.current_stream().synchronize()
                    del scratch
                model = M.MlpLSTMSingle(4, 2, T, 64)
                model.actor.load_state_dict(sd)
                model = model.to(dev)
                eng = VMPOEngine(model.actor, oracle.DEFAULT_HYPER, comm=W.comm(r))
                eng._workspace(sh, T)
                eng._vm_workspace(sh, T)
                W.comm(r).barrier()
                for i in range(2):
                    eng.update(shards[i], coefs[i])
                L = eng.losses()
                out[r] = (_flat_sd_order(model.actor, eng.flat), L)
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))
            W.bar.abort()

    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    r = want[1]
    for k in range(world):
        assert torch.equal(out[k][0], out[0][0]), f"rank {k} diverged from rank 0"
    L = out[0][1]
    assert L["selected"] == (B // 2) * (T - 1)
    for k in ("policy-loss", "value-loss", "loss-temperature", "loss-alpha", "loss"):
        _close(L[k], float(r[k]), rtol=1e-4, atol=1e-7, what=k)
    _close(out[0][0], r["params"], what="parameters after two sharded updates", **PARAMS)
    _close([L["log_eta"], L["log_alpha"]], [float(r["log_eta"]), float(r["log_alpha"])], rtol=1e-6, atol=1e-9, what="duals")


def test_sac_continuous_matches_oracle_at_cfg5_shape():
    """BASELINE config 5: SAC-Continuous, obs 17, act 6 (MuJoCo shape), replay batch 65 536 sequences x T=5 over 8 GPUs =
    8192 sequences per GPU -- one GPU's shard here.  The action width is generalised beyond the reference's hard-coded 1
    (networks/models.py:290) exactly as SURVEY.md 0.1 / A.5 specify, including the `.sum(-1)` target quirk."""
    from b200rl.engine import SACEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    dev = torch.device("cuda:0")
    B, T, f, A = 8192, 5, 17, 6
    torch.manual_seed(5)
    model = M.MlpLSTMSeperateContinuous(f, A, T, 64, act_dim=A)
    pa = OrderedDict((k, v.clone()) for k, v in model.actor.state_dict().items())
    pc = OrderedDict((k, v.clone()) for k, v in model.critic.state_dict().items())
    model = model.to(dev)
    eng = SACEngine(model, oracle.DEFAULT_HYPER, True)
    orc = oracle.OracleLearner("SAC-Continuous", {"actor": pa, "critic": pc})
    g = torch.Generator().manual_seed(9)
    for i in range(2):
        batch = make_rollouts(B, T, f, A, continuous=True, batch_index=i, act_dim=A)
        noise = [torch.randn(B, T, A, generator=g), torch.randn(B, T, A, generator=g)]
        r = orc.update(batch, noise=noise)
        eng.update({k: v.to(dev) for k, v in batch.items()}, noise=[e.to(dev) for e in noise])
        L = eng.losses()
        for k in ("policy-loss", "value-loss", "loss_alpha", "alpha"):
            _close(L[k], float(r[k]), rtol=1e-4, atol=1e-7, what=f"{k}[{i}]")
        _close(float(eng.norm_a.item()), float(r["actor_grad_norm"]), rtol=2e-4, what="actor grad norm")
        _close(float(eng.norm_q.item()), float(r["critic_grad_norm"]), rtol=2e-4, what="critic grad norm")
        _close(torch.cat([v.reshape(-1) for v in model.actor.state_dict().values()]).cpu(), r["actor_params"], what="actor params", **PARAMS)
        _close(torch.cat([v.reshape(-1) for v in model.critic.state_dict().values()]).cpu(), r["critic_params"], what="critic params", **PARAMS)


def test_ppo_cfg1_shape_matches_oracle():
    """BASELINE config 1 (the reference's own CPU-runnable case): PPO, 64 sequences x T=32 -- also pinned to a
    reference-generated fixture in test_learner_step.py; here 10 consecutive updates against the oracle (SURVEY 8d:
    'parameters after 1 and after 10 updates')."""
    from b200rl.engine import OnPolicyEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    dev = torch.device("cuda:0")
    B, T, f, A = 64, 32, 4, 2
    torch.manual_seed(0)
    model = M.MlpLSTMSingle(f, A, T, 64)
    sd = OrderedDict((k, v.clone()) for k, v in model.actor.state_dict().items())
    model = model.to(dev)
    eng = OnPolicyEngine(model.actor, "PPO", oracle.DEFAULT_HYPER)
    orc = oracle.OracleLearner("PPO", sd)
    for i in range(10):
        batch = make_rollouts(B, T, f, A, batch_index=i)
        r = orc.update(batch)
        eng.update({k: v.to(dev) for k, v in batch.items()})
        if i in (0, 9):
            _close(eng.losses()["loss"], float(r["loss"]), rtol=1e-4, atol=1e-7, what=f"loss[{i}]")
            _close(_flat_sd_order(model.actor, eng.flat), r["params"], what=f"params after {i + 1} updates", **PARAMS)
