"""GAE / V-trace / KL / soft-update kernels against the oracle (reference
compute_loss.py restated) on seeded inputs, golden fixtures, and size-independent
properties at the BASELINE shapes."""
import numpy as np
import pytest
import torch

from conftest import load_golden
import oracle

TOL = dict(rtol=1e-4, atol=1e-6)


def _rand(B, T, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(rew=0.1 * r(B, T, 1), is_fir=(torch.rand(B, T, 1, generator=g) < 0.1).float(),
                value=r(B, T, 1), blp=-torch.rand(B, T, 1, generator=g) * 2, tlp=-torch.rand(B, T, 1, generator=g) * 2)


@pytest.mark.parametrize("B,n", [(1, 1), (3, 5), (8, 31), (5, 32), (4, 33), (7, 127), (2, 128), (3, 300), (0, 7), (4, 0), (6, 64), (5, 65), (3, 129), (4, 255), (2, 513)])
def test_compute_gae(backend, B, n):
    ops, dev = backend
    d = torch.randn(B, n, 1, generator=torch.Generator().manual_seed(B * 1000 + n))
    want = oracle.compute_gae(d, 0.99, 0.95) if B and n else torch.empty(B, n, 1)
    got = ops.compute_gae(d.to(dev), 0.99, 0.95).cpu()
    torch.testing.assert_close(got, want, **TOL)


@pytest.mark.parametrize("B,T", [(2, 2), (3, 6), (4, 32), (4, 33), (5, 34), (3, 128), (2, 129), (2, 257),
                                 (1, 4), (5, 8), (3, 132), (3, 256), (2, 260), (9, 512)])     # (T % 4 == 0: four timesteps per lane)
def test_td_gae(backend, B, T):
    ops, dev = backend
    x = _rand(B, T, seed=T)
    td = x["rew"][:, :-1] + 0.99 * (1 - x["is_fir"][:, 1:]) * x["value"][:, 1:]
    want = oracle.compute_gae(td - x["value"][:, :-1], 0.99, 0.95)
    gtd, ggae = ops.td_gae(x["rew"].to(dev), x["is_fir"].to(dev), x["value"].to(dev), 0.99, 0.95)
    torch.testing.assert_close(gtd.cpu(), td, **TOL)
    torch.testing.assert_close(ggae.cpu(), want, **TOL)


@pytest.mark.parametrize("B,T", [(1, 1), (2, 2), (3, 6), (4, 32), (4, 33), (5, 34), (3, 64), (3, 128), (2, 129), (2, 257),
                                 (1, 4), (5, 8), (3, 132), (3, 256), (2, 260), (9, 512)])     # (T % 4 == 0: four timesteps per lane)
def test_compute_v_trace(backend, B, T):
    ops, dev = backend
    x = _rand(B, T, seed=100 + T)
    want = oracle.compute_v_trace(x["blp"], x["tlp"], x["is_fir"], x["rew"], x["value"], 0.99)
    got = ops.compute_v_trace(*[x[k].to(dev) for k in ("blp", "tlp", "is_fir", "rew", "value")], 0.99)
    for g, w in zip(got, want):
        torch.testing.assert_close(g.cpu(), w, **TOL)


@pytest.mark.parametrize("name", ["ppo_small", "ppo_cfg1", "impala_small", "impala_mid"])
def test_scans_against_reference_golden(backend, name):
    ops, dev = backend
    g = load_golden(name)
    rew = torch.from_numpy(g["in/0/rew"]).to(dev)
    fir = torch.from_numpy(g["in/0/is_fir"]).to(dev)
    val = torch.from_numpy(g["fwd/value"]).to(dev)
    td, gae = ops.td_gae(rew, fir, val, 0.99, 0.95)
    np.testing.assert_allclose(td.cpu().numpy(), g["fwd/td_target"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gae.cpu().numpy(), g["fwd/gae"], rtol=1e-4, atol=1e-6)
    blp = torch.from_numpy(g["in/0/log_prob"]).to(dev)
    tlp = torch.from_numpy(g["fwd/log_probs"]).to(dev)
    rho, adv, vs = ops.compute_v_trace(blp, tlp, fir, rew, val, 0.99)
    np.testing.assert_allclose(rho.cpu().numpy(), g["fwd/vt_rho"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(adv.cpu().numpy(), g["fwd/vt_adv"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(vs.cpu().numpy(), g["fwd/vt_vs"], rtol=1e-4, atol=1e-6)
    kl = ops.kldivergence(torch.from_numpy(g["in/0/logits"][:, :-1].copy()).to(dev),
                          torch.from_numpy(g["fwd/logits"][:, :-1].copy()).to(dev))
    np.testing.assert_allclose(kl.cpu().numpy(), g["fwd/kl"], rtol=1e-4, atol=1e-6)


def test_soft_update(backend):
    ops, dev = backend
    s = torch.randn(1000).to(dev)
    t = torch.randn(1000).to(dev)
    want = (1 - 0.005) * t + 0.005 * s
    ops.soft_update_flat(s, t, 0.005)
    torch.testing.assert_close(t, want, rtol=1e-6, atol=1e-7)
    # aliasing (reference quirk: target critic IS the critic): theta <- (1-tau) theta + tau theta
    a = torch.randn(257).to(dev)
    want = (1 - 0.005) * a + 0.005 * a
    ops.soft_update_flat(a, a, 0.005)
    torch.testing.assert_close(a, want, rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
def test_scan_properties_at_baseline_shape():
    """B=4096, T=128 (north-star shape): properties that need no O(B*T) python loop."""
    from b200rl import ops as O
    ops, dev = O.default(), torch.device("cuda:0")
    B, T = 4096, 128
    x = {k: v.to(dev) for k, v in _rand(B, T, seed=7).items()}
    td, gae = ops.td_gae(x["rew"], x["is_fir"], x["value"], 0.99, 0.95)
    delta = td - x["value"][:, :-1]
    # recurrence identity g[t] - k g[t+1] = delta[t]
    k = 0.99 * 0.95
    lhs = gae.clone()
    lhs[:, :-1] -= k * gae[:, 1:]
    torch.testing.assert_close(lhs, delta, rtol=1e-4, atol=1e-5)
    # linearity of the scan
    g2 = ops.compute_gae(2.0 * delta, 0.99, 0.95)
    torch.testing.assert_close(g2, 2.0 * gae, rtol=1e-4, atol=1e-5)
    # on-policy V-trace (rho = 1 clipped to 0.8) satisfies its own recurrence
    rho, adv, vs = ops.compute_v_trace(x["blp"], x["blp"], x["is_fir"], x["rew"], x["value"], 0.99)
    assert torch.all(rho == 0.8)
    disc = 0.99 * (1 - x["is_fir"])
    acc = vs - x["value"]
    rhs = 0.8 * (x["rew"][:, :-1] + disc[:, 1:] * x["value"][:, 1:] - x["value"][:, :-1]) + disc[:, 1:] * acc[:, 1:]
    torch.testing.assert_close(acc[:, :-1], rhs, rtol=1e-4, atol=1e-5)
    assert torch.all(acc[:, -1] == 0)
    # and agrees with the oracle on a slice
    w = oracle.compute_v_trace(*[x[k][:64].cpu() for k in ("blp", "blp", "is_fir", "rew", "value")], 0.99)
    torch.testing.assert_close(adv[:64].cpu(), w[1], rtol=1e-4, atol=1e-5)
