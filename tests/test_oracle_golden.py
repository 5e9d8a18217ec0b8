"""The oracle (oracle/) against fixtures produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden, golden_batch, golden_sd
import oracle
from oracle import models as OM

RT = dict(rtol=1e-4, atol=1e-6)


def _close(a, b, rtol=1e-4, atol=1e-6, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(1.0, float(np.abs(b).max())) if b.size else 1.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol * scale, err_msg=what)


@pytest.mark.parametrize("name,algo", [
    ("ppo_small", "PPO"), ("ppo_cfg1", "PPO"), ("ppo_repo_default", "PPO"),
    ("impala_small", "IMPALA"), ("impala_mid", "IMPALA"),
    ("ppo_continuous_small", "PPO-Continuous"),
])
def test_on_policy_updates(name, algo):
    g = load_golden(name)
    B, T, f, A, H, n, cont = [int(x) for x in g["meta"]]
    L = oracle.OracleLearner(algo, golden_sd(g, "sd0"))
    for i in range(n):
        r = L.update(golden_batch(g, i))
        if i == 0:
            _close(r["value"], g["fwd/value"], what="value")
            _close(r["log_probs"], g["fwd/log_probs"], what="log_probs")
            _close(r["entropy"], g["fwd/entropy"], what="entropy")
            _close(r["logits"], g["fwd/logits"], what="logits")
            if algo == "IMPALA":
                _close(r["advantages"], g["fwd/vt_adv"], what="adv")
                _close(r["values_target"], g["fwd/vt_vs"], what="vs")
                _close(r["ratio"], g["fwd/vt_rho"], what="rho")
            else:
                _close(r["gae"], g["fwd/gae"], what="gae")
                _close(r["td_target"], g["fwd/td_target"], what="td")
        for k in ("policy-loss", "value-loss", "policy-entropy"):
            _close(float(r[k]), g["out/" + k][i], rtol=1e-4, atol=1e-7, what=f"{k}[{i}]")
        _close(float(r["loss"]), g["out/loss"][i], rtol=1e-4, atol=1e-7, what=f"loss[{i}]")
        _close(r["ratio"], g[f"out/ratio/{i}"], what="ratio")
        _close(r["params"], g[f"out/params/{i}"], rtol=1e-4, atol=1e-6, what=f"params[{i}]")


@pytest.mark.parametrize("name", ["vmpo_small", "vmpo_odd"])
def test_vmpo_updates(name):
    g = load_golden(name)
    B, T, f, A, H, n, cont = [int(x) for x in g["meta"]]
    L = oracle.OracleLearner("V-MPO", golden_sd(g, "sd0"))
    for i in range(n):
        r = L.update(golden_batch(g, i), coef_alpha=torch.tensor(g["coef_alpha"][i]))
        if i == 0:
            _close(r["gae"], g["fwd/gae"], what="gae")
            _close(r["kl"], g["fwd/kl"], what="kl")
        for k in ("policy-loss", "value-loss", "loss-temperature", "loss-alpha"):
            _close(float(r[k]), g["out/" + k][i], rtol=1e-4, atol=1e-7, what=f"{k}[{i}]")
        _close(float(r["loss"]), g["out/loss"][i], rtol=1e-5, atol=1e-8, what="loss")
        _close(r["params"], g[f"out/params/{i}"], what=f"params[{i}]")
        _close([float(r["log_eta"]), float(r["log_alpha"])], g["out/duals"][i], rtol=1e-7, atol=1e-10, what="duals")


@pytest.mark.parametrize("name,algo", [("sac_small", "SAC"), ("sac_continuous_small", "SAC-Continuous")])
def test_sac_updates(name, algo):
    g = load_golden(name)
    B, T, f, A, H, n, cont = [int(x) for x in g["meta"]]
    L = oracle.OracleLearner(algo, {"actor": golden_sd(g, "actor0"), "critic": golden_sd(g, "critic0")})
    for i in range(n):
        noise = None
        if cont:
            noise = [torch.from_numpy(g[f"noise/{i}/0"]), torch.from_numpy(g[f"noise/{i}/1"])]
        r = L.update(golden_batch(g, i), noise=noise)
        for k in ("policy-loss", "value-loss", "loss_alpha", "alpha"):
            _close(float(r[k]), g["out/" + k][i], rtol=1e-4, atol=1e-7, what=f"{k}[{i}]")
        _close(float(r["loss"]), g["out/loss"][i], rtol=1e-4, atol=1e-7, what="loss")
        _close(r["actor_params"], g[f"out/actor_params/{i}"], what="actor params")
        _close(r["critic_params"], g[f"out/critic_params/{i}"], what="critic params")
        _close(float(r["log_alpha"]), g["out/log_alpha"][i], rtol=1e-7, atol=1e-10, what="log_alpha")


def test_init_params_consume_rng_like_reference():
    g = load_golden("init_seed0")

    def chk(prefix, params):
        for k, v in params.items():
            assert tuple(g[f"{prefix}/{k}#shape"]) == tuple(v.shape), (prefix, k)
            np.testing.assert_array_equal(g[f"{prefix}/{k}#head"], v.reshape(-1)[:8].numpy())
            np.testing.assert_allclose(g[f"{prefix}/{k}#sum"], v.double().sum().item(), rtol=1e-12)

    torch.manual_seed(0)
    chk("single", OM.init_params("single", 4, 2))
    torch.manual_seed(0)
    chk("single_continuous", OM.init_params("single_continuous", 2, 1))
    torch.manual_seed(0)
    chk("sep_actor", OM.init_params("actor", 4, 2))
    chk("sep_critic", OM.init_twin("critic", 4, 2))
    torch.manual_seed(0)
    chk("sepc_actor", OM.init_params("actor_continuous", 2, 1))
    chk("sepc_critic", OM.init_twin("critic_continuous", 2, 1))


def test_fused_cell_equals_explicit_cell():
    torch.manual_seed(3)
    p = OM.init_params("single", 4, 2)
    obs = torch.randn(5, 7, 4)
    hx, cx = 0.1 * torch.randn(5, 64), 0.1 * torch.randn(5, 64)
    act = torch.randint(0, 2, (5, 7, 1)).float()
    a = OM.actor_critic_forward(p, obs, hx, cx, act, fused_cell=False)
    b = OM.actor_critic_forward(p, obs, hx, cx, act, fused_cell=True)
    for x, y in zip(a, b):
        torch.testing.assert_close(x, y, rtol=1e-5, atol=1e-6)
