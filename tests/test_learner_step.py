"""End-to-end learner update (CUDA kernels through the C-ABI) against the oracle and the
reference-generated golden fixtures: forward outputs, advantages, losses, gradients and
parameters after each optimizer step."""
import numpy as np
import pytest
import torch

from conftest import load_golden, golden_batch, golden_sd
import oracle

PARAMS = dict(rtol=1e-4, atol=1e-6)


def _model(g, algo, dev):
    from b200rl.networks import models as M
    B, T, f, A, H, n, cont = [int(x) for x in g["meta"]]
    cls = M.MlpLSTMSingleContinuous if cont else M.MlpLSTMSingle
    model = cls(f, A, T, H)
    model.actor.load_state_dict(golden_sd(g, "sd0"))
    return model.to(dev)


def _flat_sd_order(net, flat_like):
    return torch.cat([v.reshape(-1) for v in net.views_of(flat_like).values()]).cpu()


def _close(a, b, rtol=1e-4, atol=1e-6, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(1.0, float(np.abs(b).max())) if b.size else 1.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol * scale, err_msg=what)


@pytest.mark.parametrize("name", ["ppo_small", "impala_small", "ppo_continuous_small", "ppo_repo_default"])
def test_forward_matches_reference(backend, name):
    ops, dev = backend
    g = load_golden(name)
    model = _model(g, None, dev)
    b = golden_batch(g, 0, dev)
    logits, logp, ent, value = model.actor.forward(b["obs"], (b["hx"], b["cx"]), b["act"], ops=ops)
    _close(value.cpu(), g["fwd/value"], what="value")
    _close(logp.cpu(), g["fwd/log_probs"], what="log_probs")
    _close(ent.cpu(), g["fwd/entropy"], what="entropy")
    _close(logits.cpu(), g["fwd/logits"], what="logits")


@pytest.mark.parametrize("variant", [0, 3, 5])
def test_forward_kernel_variants_agree_with_reference(backend, variant, monkeypatch):
    """The forward kernel generations that stay selectable (B200RL_FWD_VARIANT), all parity-green: 0 = FFMA with the
    projection in the loop (the independent reference), 3 = tcgen05 projection + tcgen05 recurrence + head_fwd (also the
    fallback for wide inputs / many heads), 5 = torso + input projection + recurrence + heads in ONE persistent tcgen05
    kernel (default)."""
    from b200rl import ops as O
    ops, dev = backend
    monkeypatch.setattr(O, "FWD_VARIANT", variant)
    g = load_golden("ppo_small")
    model = _model(g, None, dev)
    b = golden_batch(g, 0, dev)
    logits, logp, ent, value = model.actor.forward(b["obs"], (b["hx"], b["cx"]), b["act"], ops=ops)
    _close(value.cpu(), g["fwd/value"], what="value")
    _close(logits.cpu(), g["fwd/logits"], what="logits")


@pytest.mark.parametrize("bwd_variant,wgrad_variant", [(0, 0), (0, 1), (1, 1), (2, 1), (3, 1)])
def test_backward_kernel_variants_agree_with_oracle(backend, bwd_variant, wgrad_variant, monkeypatch):
    """Backward generations stay selectable and parity-green: FFMA BPTT + FFMA weight gradients (0, 0), FFMA BPTT +
    tcgen05 dW / dx (0, 1), tcgen05 BPTT with the dx GEMM fused + tcgen05 dW + torso pass over dx (1, 1; default)."""
    from b200rl import ops as O
    from b200rl.engine import OnPolicyEngine
    ops, dev = backend
    monkeypatch.setattr(O, "BWD_VARIANT", bwd_variant)
    monkeypatch.setenv("B200RL_WGRAD_VARIANT", str(wgrad_variant))
    g = load_golden("ppo_small")
    model = _model(g, "PPO", dev)
    eng = OnPolicyEngine(model.actor, "PPO", oracle.DEFAULT_HYPER, ops=ops)
    orc = oracle.OracleLearner("PPO", golden_sd(g, "sd0"))
    for i in range(2):
        r = orc.update(golden_batch(g, i))
        eng.update(golden_batch(g, i, dev))
        _close(_flat_sd_order(model.actor, eng.grad), r["grad"], rtol=1e-3, atol=2e-5 * max(1.0, float(r["grad"].abs().max())), what="grad")
        _close(_flat_sd_order(model.actor, eng.flat), g[f"out/params/{i}"], what=f"params[{i}]", **PARAMS)


CASES = [("ppo_small", "PPO"), ("impala_small", "IMPALA"), ("ppo_continuous_small", "PPO-Continuous"),
         ("ppo_repo_default", "PPO"), ("impala_mid", "IMPALA"), ("ppo_cfg1", "PPO")]


@pytest.mark.parametrize("name,algo", CASES)
def test_update_matches_reference_and_oracle(backend, name, algo):
    from b200rl.engine import OnPolicyEngine
    ops, dev = backend
    g = load_golden(name)
    B, T, f, A, H, n, cont = [int(x) for x in g["meta"]]
    if ops.lib.emulated and B * T > 700:
        pytest.skip("too slow for the host emulator")
    model = _model(g, algo, dev)
    eng = OnPolicyEngine(model.actor, algo, oracle.DEFAULT_HYPER, ops=ops)
    eng.debug_outs = {}
    orc = oracle.OracleLearner(algo, golden_sd(g, "sd0"))
    for i in range(n):
        r = orc.update(golden_batch(g, i))
        eng.update(golden_batch(g, i, dev))
        L = eng.losses()
        # intermediates vs oracle
        d = eng.debug_outs
        if algo == "IMPALA":
            _close(d["advantages"].cpu(), r["advantages"], what="advantages")
            _close(d["values_target"].cpu(), r["values_target"], what="values_target")
        else:
            _close(d["gae"].cpu(), r["gae"], what="gae")
            _close(d["td_target"].cpu(), r["td_target"], what="td_target")
        _close(d["ratio"].cpu(), r["ratio"], what="ratio")
        # gradient (clipped == unclipped here: norm << 40) vs oracle autograd
        gn = float(r["grad_norm"])
        assert gn < 40
        _close(_flat_sd_order(model.actor, eng.grad), r["grad"], rtol=1e-3, atol=2e-5 * max(1.0, float(r["grad"].abs().max())), what="grad")
        _close(L["grad-norm"], gn, rtol=1e-4, atol=1e-7, what="grad-norm")
        # losses vs the reference's own numbers
        for k in ("policy-loss", "value-loss", "policy-entropy"):
            _close(L[k], g["out/" + k][i], rtol=1e-4, atol=1e-7, what=f"{k}[{i}]")
        _close(L["loss"], g["out/loss"][i], rtol=1e-4, atol=1e-7, what=f"loss[{i}]")
        ratio = g[f"out/ratio/{i}"]
        _close([L["ratio-min"], L["ratio-max"], L["ratio-avg"]], [ratio.min(), ratio.max(), ratio.mean()], what="ratio stats")
        # parameters after the optimizer step vs the reference
        _close(_flat_sd_order(model.actor, eng.flat), g[f"out/params/{i}"], what=f"params[{i}]", **PARAMS)


def test_ragged_batch_and_single_row(backend):
    """B not a multiple of the 32-row tile, B*T not a multiple of the 32-row wgrad chunk."""
    from b200rl.engine import OnPolicyEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    ops, dev = backend
    for (B, T) in [(1, 2), (5, 3), (33, 3)]:
        torch.manual_seed(B)
        model = M.MlpLSTMSingle(4, 2, T, 64)
        sd = {k: v.clone() for k, v in model.actor.state_dict().items()}
        model = model.to(dev)
        batch = make_rollouts(B, T, 4, 2)
        orc = oracle.OracleLearner("PPO", sd)
        r = orc.update(batch)
        eng = OnPolicyEngine(model.actor, "PPO", oracle.DEFAULT_HYPER, ops=ops)
        eng.update({k: v.to(dev) for k, v in batch.items()})
        L = eng.losses()
        _close(L["loss"], float(r["loss"]), rtol=1e-4, atol=1e-7, what=f"loss B={B} T={T}")
        _close(_flat_sd_order(model.actor, eng.grad), r["grad"], rtol=1e-3, atol=2e-5 * max(1.0, float(r["grad"].abs().max())), what="grad")
        _close(_flat_sd_order(model.actor, eng.flat), r["params"], what="params")


@pytest.mark.parametrize("name", ["vmpo_small", "vmpo_odd"])
def test_vmpo_update_matches_reference(backend, name):
    from b200rl.engine import VMPOEngine
    ops, dev = backend
    g = load_golden(name)
    B, T, f, A, H, n, cont = [int(x) for x in g["meta"]]
    model = _model(g, "V-MPO", dev)
    eng = VMPOEngine(model.actor, oracle.DEFAULT_HYPER, ops=ops)
    orc = oracle.OracleLearner("V-MPO", golden_sd(g, "sd0"))
    for i in range(n):
        ca = float(g["coef_alpha"][i])
        r = orc.update(golden_batch(g, i), coef_alpha=torch.tensor(g["coef_alpha"][i]))
        eng.update(golden_batch(g, i, dev), ca)
        L = eng.losses()
        assert L["selected"] == -(-B // 2) * (T - 1)                 # exactly ceil(B/2) per timestep
        _close(eng.gae.cpu().unsqueeze(-1), r["gae"], what="gae")
        _close(_flat_sd_order(model.actor, eng.grad), r["grad"], rtol=1e-3, atol=2e-5 * max(1.0, float(r["grad"].abs().max())), what="grad")
        for k in ("policy-loss", "value-loss", "loss-temperature", "loss-alpha"):
            _close(L[k], g["out/" + k][i], rtol=1e-4, atol=1e-7, what=f"{k}[{i}]")
        _close(L["loss"], g["out/loss"][i], rtol=1e-4, atol=1e-7, what="loss")
        _close(_flat_sd_order(model.actor, eng.flat), g[f"out/params/{i}"], what=f"params[{i}]", **PARAMS)
        _close([L["log_eta"], L["log_alpha"]], g["out/duals"][i], rtol=1e-6, atol=1e-9, what="duals")


@pytest.mark.gpu
def test_vmpo_selection_properties_at_cfg4_shard_shape():
    """B=1024 sequences per GPU, T=128 (BASELINE config 4's per-GPU shard): exactly ceil(B/2) advantages are
    selected in every timestep column, the threshold is the k-th largest, psi sums to one."""
    from b200rl.engine import VMPOEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    dev = torch.device("cuda:0")
    B, T = 1024, 128
    torch.manual_seed(0)
    model = M.MlpLSTMSingle(4, 2, T, 64).to(dev)
    eng = VMPOEngine(model.actor, oracle.DEFAULT_HYPER)
    batch = {k: v.to(dev) for k, v in make_rollouts(B, T, 4, 2).items()}
    eng.update(batch, 0.0075)
    L = eng.losses()
    assert L["selected"] == (B // 2) * (T - 1)
    kth = torch.topk(eng.gae, B // 2, dim=0).values[-1]
    torch.testing.assert_close(eng.thr, kth, rtol=0, atol=0)
    eta = 5.0 + 1e-7
    sel = eng.gae >= eng.thr[None, :]
    u = (eng.gae / eta).double()
    psi = torch.where(sel, (u - u[sel].max()).exp(), torch.zeros_like(u))
    psi = psi / psi.sum()
    assert abs(float(psi.sum()) - 1.0) < 1e-9
    # the engine's normaliser equals the direct one
    Z = float(eng.scalars[7].item())
    direct = float(torch.where(sel, (u - float(eng.scalars[6].item())).exp(), torch.zeros_like(u)).sum())
    assert abs(Z - direct) / direct < 1e-5


@pytest.mark.parametrize("name,cont", [("sac_small", False), ("sac_continuous_small", True)])
def test_sac_update_matches_reference(backend, name, cont):
    from b200rl.engine import SACEngine
    from b200rl.networks import models as M
    ops, dev = backend
    g = load_golden(name)
    B, T, f, A, H, n, _ = [int(x) for x in g["meta"]]
    model = (M.MlpLSTMSeperateContinuous if cont else M.MlpLSTMSeperate)(f, A, T, H)
    model.actor.load_state_dict(golden_sd(g, "actor0"))
    model.critic.load_state_dict(golden_sd(g, "critic0"))
    model = model.to(dev)
    eng = SACEngine(model, oracle.DEFAULT_HYPER, cont, ops=ops)
    for i in range(n):
        noise = [torch.from_numpy(g[f"noise/{i}/{j}"]).to(dev) for j in range(2)] if cont else None
        eng.update(golden_batch(g, i, dev), noise=noise)
        L = eng.losses()
        for k in ("policy-loss", "value-loss", "loss_alpha", "alpha"):
            _close(L[k], g["out/" + k][i], rtol=1e-4, atol=1e-7, what=f"{k}[{i}]")
        _close(L["loss"], g["out/loss"][i], rtol=1e-4, atol=1e-7, what="loss")
        _close(L["log_alpha"], g["out/log_alpha"][i], rtol=1e-6, atol=1e-9, what="log_alpha")
        pa = torch.cat([v.reshape(-1) for v in model.actor.state_dict().values()]).cpu()
        pc = torch.cat([v.reshape(-1) for v in model.critic.state_dict().values()]).cpu()
        _close(pa, g[f"out/actor_params/{i}"], what=f"actor params[{i}]", **PARAMS)
        _close(pc, g[f"out/critic_params/{i}"], what=f"critic params[{i}]", **PARAMS)


def test_sac_continuous_generalised_action_width(backend):
    """BASELINE config 5 shape (obs=17, act=6): outside the reference's domain (its critic hard-codes a
    1-wide action, networks/models.py:290); the oracle is the reference math with that single generalisation,
    including the `.sum(-1)` target quirk."""
    from collections import OrderedDict
    from b200rl.engine import SACEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    ops, dev = backend
    B, T, f, A = (16, 5, 17, 6)
    torch.manual_seed(5)
    model = M.MlpLSTMSeperateContinuous(f, A, T, 64, act_dim=A)
    pa = OrderedDict((k, v.clone()) for k, v in model.actor.state_dict().items())
    pc = OrderedDict((k, v.clone()) for k, v in model.critic.state_dict().items())
    model = model.to(dev)
    eng = SACEngine(model, oracle.DEFAULT_HYPER, True, ops=ops)
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
        _close(torch.cat([v.reshape(-1) for v in model.actor.state_dict().values()]).cpu(), r["actor_params"], what="actor params", **PARAMS)
        _close(torch.cat([v.reshape(-1) for v in model.critic.state_dict().values()]).cpu(), r["critic_params"], what="critic params", **PARAMS)


@pytest.mark.gpu
def test_full_size_tcgen05_path_agrees_with_ffma_path(monkeypatch):
    """BASELINE shape (B=4096, T=128, obs 4, A=2), where the oracle is too slow: the tcgen05 kernels (forward 3, backward 1
    + torso pass) and the independently written FFMA kernels (forward 0, backward 0, FFMA weight gradients) must produce the
    same loss scalars and the same flat gradient from the same weights and batch; and the update must not depend on the
    order of the sequences (sums over B are the only coupling)."""
    from b200rl import ops as O
    from b200rl.engine import OnPolicyEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    dev = torch.device("cuda:0")
    B, T = 4096, 128
    batch = {k: v.to(dev) for k, v in make_rollouts(B, T, 4, 2).items()}

    def run(fwd, bwd, wg, perm=None):
        monkeypatch.setattr(O, "FWD_VARIANT", fwd)
        monkeypatch.setattr(O, "BWD_VARIANT", bwd)
        monkeypatch.setenv("B200RL_WGRAD_VARIANT", str(wg))
        torch.manual_seed(0)
        model = M.MlpLSTMSingle(4, 2, T, 64).to(dev)
        eng = OnPolicyEngine(model.actor, "PPO", oracle.DEFAULT_HYPER)
        b = batch if perm is None else {k: v[perm].contiguous() for k, v in batch.items()}
        eng.update(b)
        L = eng.losses()
        return L, eng.grad.clone()

    L_tc, g_tc = run(3, 1, 1)
    L_ff, g_ff = run(0, 0, 0)
    for k in ("policy-loss", "value-loss", "policy-entropy", "loss", "ratio-avg"):
        _close(L_tc[k], L_ff[k], rtol=1e-4, atol=1e-7, what=k)
    _close(g_tc.cpu(), g_ff.cpu(), rtol=1e-3, atol=2e-5 * float(g_ff.abs().max()), what="flat gradient, tcgen05 vs FFMA")
    # split-precision check: the tcgen05 kernels' third term (bf16 / tf32 lo parts) carries ~2^-11 of every product; if it were
    # dropped or mis-laid the gradient would be off by ~1e-4 in norm, two orders above the reduction-order noise measured here
    rel = float((g_tc - g_ff).norm() / g_ff.norm())
    assert rel < 2e-5, f"tcgen05 vs FFMA gradient differ by {rel:.2e} in norm"
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).to(dev)
    L_p, g_p = run(3, 1, 1, perm)
    _close(L_p["loss"], L_tc["loss"], rtol=1e-5, atol=1e-8, what="loss under a permutation of the sequences")
    _close(g_p.cpu(), g_tc.cpu(), rtol=1e-3, atol=2e-5 * float(g_tc.abs().max()), what="gradient under a permutation of the sequences")


@pytest.mark.parametrize("tiles,T,f,A", [(5, 5, 4, 2), (5.2, 3, 4, 2), (1.25, 1, 4, 2), (2.2, 2, 3, 3), (9.4, 1, 4, 1)])
def test_fused_forward_many_tiles_short_and_odd_windows(backend, tiles, T, f, A):
    """The fused forward's head pipeline (producer warps read h_t from the shared-memory operand three steps behind their own
    x production; two extra mbarriers against the cell warps) across tile boundaries of a CTA, with an odd window (two
    consecutive steps share an operand set there -- an ordering that once deadlocked in the emulator), windows shorter than
    the pipeline depth, ragged last tiles; against the oracle's forward (`networks/models.py:63-87`)."""
    from b200rl.networks import models as M
    ops, dev = backend
    ctas = 4 if dev.type == "cpu" else torch.cuda.get_device_properties(dev).multi_processor_count
    B = int(round(tiles * ctas * 32))
    torch.manual_seed(3)
    model = M.MlpLSTMSingle(f, A, T, 64)
    sd = {k: v.clone() for k, v in model.actor.state_dict().items()}
    model = model.to(dev)
    net = model.actor
    L, flat = net._layout, net.flat_parameters()
    u, hx, cx = torch.randn(B, T, f), 0.5 * torch.randn(B, 64), 0.5 * torch.randn(B, 64)
    ws = dict(gates=torch.empty(B, T, 64, 4, device=dev), c=torch.empty(B, T, 64, device=dev), h=torch.empty(B, T, 64, device=dev))
    ho = torch.full((B, T, L.NH), float("nan"), device=dev)
    assert ops.net_fwd(L, flat, u.to(dev), hx.to(dev), cx.to(dev), ws, ho) == 1          # ONE launch
    from oracle import models as OM
    hs = OM.lstm_unroll(OM._torso(u, sd), hx, cx, sd)
    _close(ws["h"].cpu(), hs, rtol=1e-4, atol=2e-6, what="h")
    got = ho.cpu()
    assert torch.isfinite(got).all()
    # the kernel's own head order is the layout's; compare as sets of columns through the layout offsets
    W_head = flat[L.off["W_head"][0]:L.off["W_head"][0] + L.NH * 64].view(L.NH, 64).cpu()
    b_head = flat[L.off["b_head"][0]:L.off["b_head"][0] + L.NH].cpu()
    _close(got, hs @ W_head.t() + b_head, rtol=1e-4, atol=2e-6, what="head_out")
