"""Branches of the learner path the round-1 suite never reached: an ACTIVE clip_grad_norm_ (coef < 1) under RMSprop and
Adam, K_epoch > 1, the SAC `learning()` coroutines behind `LearnerSeperate`, the shared-memory batch path
(`setup_shared_memory` -> slot writes as `LearnerStorage.make_batch` does -> `sample_batch_from_sh_memory` -> update), the
pipelined loop's publish order, V-MPO ties at the top-k threshold, and the fused peer step kernel.  Fixtures
`ppo_clip`, `ppo_kepoch2`, `sac_clip` come from the UNMODIFIED reference (tests/golden/make_golden.py --branches)."""
import asyncio
import math

import numpy as np
import pytest
import torch

from conftest import load_golden, golden_batch, golden_sd
import oracle

PARAMS = dict(rtol=1e-4, atol=1e-6)
KEYS = ("obs", "act", "rew", "logits", "log_prob", "is_fir", "hx", "cx")


def _hyper(g):
    h = dict(oracle.DEFAULT_HYPER)
    for k in g:
        if k.startswith("hyper/"):
            h[k[6:]] = float(g[k]) if k != "hyper/K_epoch" else int(g[k])
    return h


def _flat_sd_order(net, flat_like):
    return torch.cat([v.reshape(-1) for v in net.views_of(flat_like).values()]).cpu()


def _close(a, b, rtol=1e-4, atol=1e-6, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(1.0, float(np.abs(b).max())) if b.size else 1.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol * scale, err_msg=what)


class StopAfter:
    def __init__(self, n):
        self.n, self.calls = n, 0

    def is_set(self):
        self.calls += 1
        return self.calls > self.n


class FlagStop:
    def __init__(self):
        self.flag = False

    def set(self):
        self.flag = True

    def is_set(self):
        return self.flag


# ------------------------------------------------------------------------------------------- clip-active, RMSprop
def test_clip_active_rmsprop_matches_reference(backend):
    """max_grad_norm = 1e-3 << ||g||: coef = max_norm / (norm + 1e-6) < 1 scales the gradient before RMSprop
    (ppo/learning.py:93-96)."""
    from b200rl.engine import OnPolicyEngine
    from b200rl.networks import models as M
    ops, dev = backend
    g = load_golden("ppo_clip")
    hp = _hyper(g)
    assert hp["max_grad_norm"] == 1e-3
    B, T, f, A, H, n, cont = [int(x) for x in g["meta"]]
    model = M.MlpLSTMSingle(f, A, T, H)
    model.actor.load_state_dict(golden_sd(g, "sd0"))
    model = model.to(dev)
    eng = OnPolicyEngine(model.actor, "PPO", hp, ops=ops)
    orc = oracle.OracleLearner("PPO", golden_sd(g, "sd0"), hyper=hp)
    for i in range(n):
        r = orc.update(golden_batch(g, i))
        eng.update(golden_batch(g, i, dev))
        L = eng.losses()
        gn = float(r["grad_norm"])
        assert gn > 5 * hp["max_grad_norm"], "the fixture must exercise the clipping branch"
        _close(L["grad-norm"], gn, rtol=1e-4, atol=1e-9, what="pre-clip norm")
        coef = hp["max_grad_norm"] / (gn + 1e-6)
        _close(_flat_sd_order(model.actor, eng.grad), r["grad"] * coef, rtol=1e-3, atol=2e-5 * float(r["grad"].abs().max()) * coef,
               what="clipped gradient")
        _close(L["loss"], g["out/loss"][i], rtol=1e-4, atol=1e-7, what=f"loss[{i}]")
        _close(_flat_sd_order(model.actor, eng.flat), g[f"out/params/{i}"], what=f"params[{i}] vs reference", **PARAMS)
        _close(_flat_sd_order(model.actor, eng.flat), r["params"], what=f"params[{i}] vs oracle", **PARAMS)


def test_clip_active_legacy_single_cta_kernel(backend):
    """The stand-alone clip_rmsprop entry point (C-ABI) on the same branch."""
    ops, dev = backend
    torch.manual_seed(0)
    n = 1000
    p, g = torch.randn(n), torch.randn(n)
    sq = torch.rand(n)
    want_norm = g.double().norm().float()
    coef = min(1.0, 0.5 / (float(want_norm) + 1e-6))
    gg = g * coef
    v = 0.99 * sq + 0.01 * gg * gg
    want = p - 1e-4 * gg / (v.sqrt() + 1e-5)
    pd, gd, sd, no = p.to(dev), g.to(dev), sq.to(dev), torch.zeros(1, device=dev)
    ops.clip_rmsprop(pd, gd, sd, n, 1.0, 0.5, 1e-4, 0.99, 1e-5, no)
    _close(pd.cpu(), want, rtol=1e-6, atol=1e-7)
    _close(gd.cpu(), gg, rtol=1e-6, atol=1e-7)
    _close(float(no.item()), float(want_norm), rtol=1e-6)


# ------------------------------------------------------------------------------------------------- clip-active, Adam
def test_clip_active_adam_sac_matches_reference(backend):
    from b200rl.engine import SACEngine
    from b200rl.networks import models as M
    ops, dev = backend
    g = load_golden("sac_clip")
    hp = _hyper(g)
    B, T, f, A, H, n, _ = [int(x) for x in g["meta"]]
    model = M.MlpLSTMSeperate(f, A, T, H)
    model.actor.load_state_dict(golden_sd(g, "actor0"))
    model.critic.load_state_dict(golden_sd(g, "critic0"))
    orc = oracle.OracleLearner("SAC", {"actor": golden_sd(g, "actor0"), "critic": golden_sd(g, "critic0")}, hyper=hp)
    model = model.to(dev)
    eng = SACEngine(model, hp, False, ops=ops)
    for i in range(n):
        r = orc.update(golden_batch(g, i))
        assert float(r["critic_grad_norm"]) > 2 * hp["max_grad_norm"], "the fixture must clip the critic gradient"
        eng.update(golden_batch(g, i, dev))
        L = eng.losses()
        _close(float(eng.norm_q.item()), float(r["critic_grad_norm"]), rtol=1e-4, what="critic pre-clip norm")
        _close(float(eng.norm_a.item()), float(r["actor_grad_norm"]), rtol=1e-4, what="actor pre-clip norm")
        for k in ("policy-loss", "value-loss", "loss_alpha", "alpha"):
            _close(L[k], g["out/" + k][i], rtol=1e-4, atol=1e-7, what=f"{k}[{i}]")
        pa = torch.cat([v.reshape(-1) for v in model.actor.state_dict().values()]).cpu()
        pc = torch.cat([v.reshape(-1) for v in model.critic.state_dict().values()]).cpu()
        _close(pa, g[f"out/actor_params/{i}"], what=f"actor params[{i}]", **PARAMS)
        _close(pc, g[f"out/critic_params/{i}"], what=f"critic params[{i}]", **PARAMS)


# ------------------------------------------------------------------------------------------------------- K_epoch = 2
def _learner_args(dev, algo, B, T, A, tmp_path, **over):
    from b200rl.utils.utils import load_params
    args = load_params()
    args.algo, args.batch_size, args.seq_len, args.action_space = algo, B, T, A
    args.device, args.result_dir, args.model_dir = dev, None, str(tmp_path)
    args.loss_log_interval = 1
    for k, v in over.items():
        setattr(args, k, v)
    return args


def test_k_epoch_2_through_learning_coroutine(backend, tmp_path):
    """`for _ in range(K_epoch)` around the update (ppo/learning.py:36): two optimizer steps per batch."""
    from b200rl.agents import learner as LM
    from b200rl.agents.learner_module.ppo.learning import learning as learn_ppo
    from b200rl.networks import models as M
    from b200rl.utils.utils import ExecutionTimer
    ops, dev = backend
    g = load_golden("ppo_kepoch2")
    B, T, f, A, H, n, cont = [int(x) for x in g["meta"]]
    assert int(g["hyper/K_epoch"]) == 2
    args = _learner_args(dev, "PPO", B, T, A, tmp_path, K_epoch=2)
    model = M.MlpLSTMSingle(f, A, T, H)
    model.actor.load_state_dict(golden_sd(g, "sd0"))
    learner = LM.LearnerSinglePPO(args, None, model, None, StopAfter(n), None, None, [f], ops=None if dev.type == "cuda" else ops)
    logs, pubs = [], []
    learner.log_loss_tensorboard = lambda timer, loss, d: logs.append((float(loss), dict(d)))
    learner.pub_model = lambda sd: pubs.append(sd)

    async def go():
        q = asyncio.Queue()
        for i in range(n):
            q.put_nowait(tuple(torch.from_numpy(g[f"in/{i}/{k}"]) for k in KEYS))
        learner.batch_queue = q
        await learn_ppo(learner, ExecutionTimer(num_transition=B * T))

    asyncio.run(go())
    assert len(pubs) == n and len(logs) == n
    for i in range(n):
        np.testing.assert_allclose(logs[i][0], g["out/loss"][i], rtol=1e-4, atol=1e-7)
        flat = torch.cat([v.reshape(-1) for v in pubs[i].values()]).numpy()
        np.testing.assert_allclose(flat, g[f"out/params/{i}"], rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------- SAC coroutines behind LearnerSeperate
@pytest.mark.parametrize("name,cont", [("sac_small", False), ("sac_continuous_small", True)])
def test_sac_learning_coroutine_matches_reference(backend, name, cont, tmp_path):
    """`LearnerSeperate` + `sac/learning.py` / `sac_continuous/learning.py` driven like the reference drives them
    (agents/learner.py:351-400): batches from `batch_buffer`, three optimizer steps, actor state dict published."""
    from b200rl.agents import learner as LM
    from b200rl.agents.learner_module.sac.learning import learning as learn_sac
    from b200rl.agents.learner_module.sac_continuous.learning import learning as learn_sacc
    from b200rl.networks import models as M
    from b200rl.utils.utils import ExecutionTimer
    ops, dev = backend
    g = load_golden(name)
    B, T, f, A, H, n, _ = [int(x) for x in g["meta"]]
    args = _learner_args(dev, "SAC-Continuous" if cont else "SAC", B, T, A, tmp_path, buffer_size=B)
    model = (M.MlpLSTMSeperateContinuous if cont else M.MlpLSTMSeperate)(f, A, T, H)
    model.actor.load_state_dict(golden_sd(g, "actor0"))
    model.critic.load_state_dict(golden_sd(g, "critic0"))
    learner = LM.LearnerSeperate(args, None, model, None, StopAfter(n), None, None, [f], ops=None if dev.type == "cuda" else ops)
    assert learner.target_critic is learner.critic                       # the reference's aliasing (learner.py:358)
    assert learner.target_entropy == A
    if cont:                                                              # the two rsample draws per update, in the reference's order
        it = iter(range(n))
        learner.engine.noise_fn = lambda B_, T_, A_: [torch.from_numpy(g[f"noise/{i}/{j}"]).to(dev) for i in [next(it)] for j in range(2)]
    logs, pubs, crit, la = [], [], [], []
    learner.log_loss_tensorboard = lambda timer, loss, d: logs.append((float(loss), {k: float(v) for k, v in d.items()}))

    def pub(sd):
        pubs.append(sd)
        crit.append(torch.cat([v.reshape(-1) for v in learner.critic.state_dict().values()]).cpu().numpy().copy())
        la.append(float(learner.log_alpha.item()))
    learner.pub_model = pub

    async def go():
        q = asyncio.Queue()
        for i in range(n):
            b = {k: torch.from_numpy(g[f"in/{i}/{k}"]) for k in KEYS}
            b["hx"] = b["hx"][:, None, :].expand(-1, T, -1).contiguous()
            b["cx"] = b["cx"][:, None, :].expand(-1, T, -1).contiguous()
            q.put_nowait(tuple(b[k] for k in KEYS))
        learner.batch_buffer = q
        await (learn_sacc if cont else learn_sac)(learner, ExecutionTimer(num_transition=B * T))

    asyncio.run(go())
    assert learner.idx == n and len(logs) == n and len(pubs) == n
    ref_keys = [k[7:] for k in g if k.startswith("actor0/")]
    for i in range(n):
        np.testing.assert_allclose(logs[i][0], g["out/loss"][i], rtol=1e-4, atol=1e-7)
        for k in ("policy-loss", "value-loss", "loss_alpha", "alpha"):
            np.testing.assert_allclose(logs[i][1][k], g["out/" + k][i], rtol=1e-4, atol=1e-7, err_msg=k)
        assert list(pubs[i].keys()) == ref_keys
        np.testing.assert_allclose(torch.cat([v.reshape(-1) for v in pubs[i].values()]).numpy(), g[f"out/actor_params/{i}"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(crit[i], g[f"out/critic_params/{i}"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(la[i], g["out/log_alpha"][i], rtol=1e-6, atol=1e-9)
    ck = torch.load(tmp_path / f"{args.algo}_0.pt", weights_only=False)     # sac/learning.py: torch.save(parent.actor, ...)
    assert list(ck.state_dict().keys()) == ref_keys


# ---------------------------------------------------------------------------------------- shared-memory round trip
def _write_slots(shm_ref, args, obs_shape, batch, T):
    """What `LearnerStorage.make_batch` does per finished rollout (agents/learner_storage.py:123-159): flatten each
    (T, width) field into slot `num` of the flat shared arrays, then bump the counter."""
    B = batch["obs"].shape[0]
    views = {k: np.frombuffer(shm_ref[f"{k}_batch"][0].get_obj(), dtype=np.float32) for k in KEYS}
    H = args.hidden_size
    widths = dict(obs=int(np.prod(obs_shape)), act=1, rew=1, logits=args.action_space, log_prob=1, is_fir=1, hx=H, cx=H)
    for num in range(B):
        for k in KEYS:
            w = widths[k]
            src = batch[k][num]
            if k in ("hx", "cx"):                                   # the reference stores the (T, H) trajectory of states
                src = src[None, :].expand(T, -1)
            views[k][num * T * w:(num + 1) * T * w] = src.reshape(-1).numpy()
        shm_ref["batch_index"].value += 1


@pytest.mark.parametrize("algo", ["PPO", "SAC"])
def test_shared_memory_round_trip(backend, algo, tmp_path):
    """setup_shared_memory -> slot writes -> is_sh_ready -> put_batch_to_*_q -> learning(): the whole seam
    `learning_chain_<algo>` (agents/learner.py:168-233,267-305) against the oracle fed the same rollouts."""
    from b200rl.agents import learner as LM
    from b200rl.agents.storage_module.shared_batch import reset_shared_on_policy_memory, reset_shared_buffer_memory
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    ops, dev = backend
    B, T, f, A, H = 6, 5, 4, 2, 64
    sac = algo == "SAC"
    args = _learner_args(dev, algo, B, T, A, tmp_path, buffer_size=B, model_dir=None)
    torch.manual_seed(3)
    model = (M.MlpLSTMSeperate if sac else M.MlpLSTMSingle)(f, A, T, H)
    if sac:
        sd = {"actor": {k: v.clone() for k, v in model.actor.state_dict().items()},
              "critic": {k: v.clone() for k, v in model.critic.state_dict().items()}}
    else:
        sd = {k: v.clone() for k, v in model.actor.state_dict().items()}
    shm = (reset_shared_buffer_memory if sac else reset_shared_on_policy_memory)(args, [f])
    stop = FlagStop()
    cls = LM.LearnerSeperate if sac else LM.LearnerSinglePPO
    learner = cls(args, None, model, shm, stop, None, None, [f], ops=None if dev.type == "cuda" else ops)
    batch = make_rollouts(B, T, f, A)
    pubs = []

    def pub(state):
        pubs.append(state)
        stop.set()                                     # one update, then both coroutines leave their loops
    learner.pub_model = pub
    np.random.seed(0)
    idx = np.random.randint(0, B, size=B)              # the off-policy sampler's draw (learner.py:183)
    np.random.seed(0)
    assert not learner.is_sh_ready()
    _write_slots(shm, args, [f], batch, T)
    assert learner.is_sh_ready()
    asyncio.run(learner.learning_chain_sac() if sac else learner.learning_chain_ppo())
    assert len(pubs) == 1
    if sac:
        sampled = {k: v[idx] for k, v in batch.items()}
        r = oracle.OracleLearner("SAC", sd).update(sampled)
        want = r["actor_params"]
    else:
        assert shm["batch_index"].value == 0           # on-policy: the counter is reset after the batch was taken
        r = oracle.OracleLearner("PPO", sd).update(batch)
        want = r["params"]
    got = torch.cat([v.reshape(-1) for v in pubs[0].values()])
    _close(got, want, what="published parameters after the update", **PARAMS)


# ------------------------------------------------------------------------------------ pipelined loop: publish order
@pytest.mark.gpu
def test_pipelined_loop_publishes_before_waiting_for_the_next_batch(tmp_path):
    """With a feeder slower than the learner (the normal case) model i must be published BEFORE the loop blocks on
    `queue.get()` for batch i+1 -- otherwise workers collect batch i+1 under model i-1."""
    from b200rl.agents import learner as LM
    from b200rl.agents.learner_module.ppo.learning import learning as learn_ppo
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts, as_reference_tuple
    from b200rl.utils.utils import ExecutionTimer
    dev = torch.device("cuda:0")
    B, T, f, A = 16, 8, 4, 2
    args = _learner_args(dev, "PPO", B, T, A, tmp_path, model_dir=None)
    torch.manual_seed(0)
    stop = FlagStop()
    learner = LM.LearnerSinglePPO(args, None, M.MlpLSTMSingle(f, A, T, 64), None, stop, None, None, [f])
    events = []

    def pub(sd):
        events.append("pub")
        if events.count("pub") == 3:
            stop.set()
    learner.pub_model = pub
    learner.log_loss_tensorboard = lambda *a: None

    class Q(asyncio.Queue):
        async def get(self):
            if not events or events[-1] != "get":      # (the loop re-polls an empty queue; count the wait once)
                events.append("get")
            return await super().get()

    async def feeder(q):
        for i in range(3):
            await asyncio.sleep(0.05)                  # batches arrive one at a time, slower than an update
            q.put_nowait(as_reference_tuple(make_rollouts(B, T, f, A, batch_index=i), T))

    async def go():
        q = Q()
        learner.batch_queue = q
        await asyncio.gather(learn_ppo(learner, ExecutionTimer(num_transition=B * T)), feeder(q))

    asyncio.run(go())
    # (after the third publish the loop polls the queue once more, sees the stop flag and leaves)
    assert events[:6] == ["get", "pub", "get", "pub", "get", "pub"] and events[6:] in ([], ["get"]), events


# ------------------------------------------------------------------------------------------------ V-MPO top-k ties
def test_vmpo_topk_ties_select_exactly_k(backend):
    """Duplicate sequences make advantages tie AT the k-th value: exactly ceil(B/2) rows per column must be selected
    (torch.topk semantics), lowest row index first; the softmax normaliser and the temperature loss use that count."""
    from b200rl import native
    ops, dev = backend
    Bg, n, k = 8, 3, 4
    gae = torch.tensor([[1.0, 5.0, 0.0], [2.0, 5.0, 0.0], [2.0, 5.0, 0.0], [2.0, 1.0, 0.0],
                        [2.0, 1.0, 0.0], [3.0, 5.0, 0.0], [0.5, 5.0, 0.0], [2.0, 0.0, 0.0]]).to(dev)
    duals = torch.tensor([math.log(5.0), math.log(5.0)], dtype=torch.float64, device=dev)
    thr = torch.empty(n, device=dev)
    col = torch.empty(2 * n, device=dev)
    cut = torch.empty(n, dtype=torch.int32, device=dev)
    ptr = native.ptr
    ops.lib.call("b200rl_vmpo_select", ptr(gae), Bg, n, k, ptr(duals), ptr(thr), ptr(col), ptr(cut), ops._stream(dev))
    assert thr.cpu().tolist() == [2.0, 5.0, 0.0]
    # column 0: 3.0 + three of the five 2.0 (rows 1, 2, 3) ; column 1: four of the five 5.0 (rows 0, 1, 2, 5) ; column 2: rows 0-3
    assert cut.cpu().tolist() == [3, 5, 3]
    g = gae.cpu()
    sel = (g > thr.cpu()) | ((g == thr.cpu()) & (torch.arange(Bg)[:, None] <= cut.cpu()[None, :]))
    assert sel.sum(0).tolist() == [k, k, k]
    inv = 1.0 / (5.0 + 1e-7)
    want = [(torch.exp((g[:, t][sel[:, t]] - g[:, t].max()) * inv)).double().sum().item() for t in range(n)]
    _close(col[n:].cpu(), want, rtol=1e-6)


# ---------------------------------------------------------------------------------- fused peer step, single process
@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 8])
def test_peer_step_loopback_equals_full_batch(world):
    """`world` ranks as threads of this process on ONE GPU (own stream, own model copy, own shard): the fused
    {flag handshake -> one-shot all-reduce by direct peer loads -> clip -> RMSprop} kernel -- the one the multi-GPU job
    runs, here with the "peers" in the same device memory -- must reproduce the full-batch oracle update, and all ranks
    must end with bit-identical parameters."""
    import threading
    from b200rl import comm as CM, native
    from b200rl.engine import OnPolicyEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    dev = torch.device("cuda:0")
    native.get().call("b200rl_peer_set_timeout_ms", 20000)
    B, T, f, A = 16 * world, 12, 4, 2
    full = make_rollouts(B, T, f, A)
    torch.manual_seed(0)
    proto = M.MlpLSTMSingle(f, A, T, 64)
    sd = {k: v.clone() for k, v in proto.actor.state_dict().items()}
    W = CM.LoopbackWorld(world, dev)
    out, errs = [None] * world, []

    warm = threading.Lock()

    def rank_main(r):
        try:
            torch.cuda.set_device(dev)
            with torch.cuda.stream(torch.cuda.Stream(dev)):
                sh = B // world
                shard = {k: v[r * sh:(r + 1) * sh].contiguous().to(dev) for k, v in full.items()}
                # One process, one device: a cudaMalloc or a lazy kernel load in rank A synchronises the DEVICE and would wait
                # for rank B's peer kernel, which spins until A announces -- so every rank first runs a solo update of the same
                # shapes (kernels loaded, allocator pool of this stream filled), then all ranks start together.  (A real job is
                # one process per GPU and has no such coupling.)
                with warm:
                    scratch = M.MlpLSTMSingle(f, A, T, 64).to(dev)
                    OnPolicyEngine(scratch.actor, "PPO", oracle.DEFAULT_HYPER).update(shard)
                    torch.cuda.current_stream().synchronize()
                    del scratch
                model = M.MlpLSTMSingle(f, A, T, 64)
                model.actor.load_state_dict(sd)
                model = model.to(dev)
                eng = OnPolicyEngine(model.actor, "PPO", oracle.DEFAULT_HYPER, comm=W.comm(r))
                eng._workspace(sh, T)
                W.comm(r).barrier()
                for _ in range(3):
                    eng.update(shard)
                L = eng.losses()
                out[r] = (_flat_sd_order(model.actor, eng.flat), L)
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))
            W.bar.abort()

    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    orc = oracle.OracleLearner("PPO", sd)
    for _ in range(3):
        want = orc.update(full)
    for r in range(world):
        assert torch.equal(out[r][0], out[0][0]), f"rank {r} diverged from rank 0"
        _close(out[r][1]["loss"], float(want["loss"]), rtol=1e-4, atol=1e-7, what="loss")
    _close(out[0][0], want["params"], what="parameters after 3 sharded updates", **PARAMS)


def test_peer_step_single_rank_equals_clip_rmsprop(backend):
    """world = 1: the multi-CTA fused kernel against the single-CTA clip_rmsprop entry point, clip active and not."""
    from b200rl import comm as CM
    ops, dev = backend
    n = 4096 + 8
    for max_norm in (40.0, 0.25):
        torch.manual_seed(1)
        p, g, sq = torch.randn(n).to(dev), torch.randn(n).to(dev), torch.rand(n).to(dev)
        p2, g2, sq2 = p.clone(), g.clone(), sq.clone()
        no, no2 = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
        ops.clip_rmsprop(p, g, sq, n, 0.5, max_norm, 1e-3, 0.99, 1e-5, no)
        region = torch.zeros(CM.HEADER_BYTES + 2 * ((n * 4 + 64 + 255) // 256 * 256), dtype=torch.uint8, device=dev)
        ch = CM.PeerChannel(n, 0, 1, dev, ops.lib, local_regions=[region.data_ptr()])
        gbuf, sbuf = ch.begin()
        gbuf.copy_(g2)
        sbuf.copy_(torch.arange(8, dtype=torch.float64))
        gout, sout = torch.zeros(n, device=dev), torch.zeros(8, dtype=torch.float64, device=dev)
        ch.step(0, p2, sq2, None, None, gout, sout, no2, 0.5, max_norm, 1e-3, 0.99, 0.0, 1e-5, 0, ops._stream(dev))
        ch.check()
        _close(p2.cpu(), p.cpu(), rtol=1e-6, atol=1e-7, what="param")
        _close(sq2.cpu(), sq.cpu(), rtol=1e-6, atol=1e-7, what="square average")
        _close(gout.cpu(), g.cpu(), rtol=1e-6, atol=1e-7, what="clipped gradient")
        _close(no2.cpu(), no.cpu(), rtol=1e-6, what="norm")
        assert sout.cpu().tolist() == list(range(8))


# ------------------------------------------------------------------------------------------- model publish (f3)
def test_publish_interval_and_async_mailbox(monkeypatch):
    """`pub_model` with publish_interval = k sends every k-th state dict; with publish_async the wire encoding runs on a worker
    thread behind a newest-wins mailbox: the last published model always arrives, frames decode to the reference's wire format
    (pickle(Protocol.Model), compressed pickle of the state dict -- utils/utils.py:244-245)."""
    import pickle
    import sys
    import threading
    import time
    import types
    from types import SimpleNamespace
    from b200rl.agents import learner as LM
    from b200rl.utils.utils import Protocol
    fake = types.ModuleType("blosc2")
    fake.compress = lambda b, clevel=1: b[::-1]               # (blosc2 is not installed in the test image: any invertible stand-in)
    monkeypatch.setitem(sys.modules, "blosc2", fake)

    class Sock:
        def __init__(self, delay=0.0):
            self.frames, self.delay = [], delay

        def send_multipart(self, parts):
            time.sleep(self.delay)
            self.frames.append(parts)

    def decode(parts):
        proto, payload = parts
        return pickle.loads(proto), pickle.loads(payload[::-1])

    lb = LM.LearnerBase.__new__(LM.LearnerBase)
    lb.args = SimpleNamespace(publish_interval=3, publish_async=False)
    lb.pub_socket = Sock()
    for i in range(7):
        lb.pub_model({"w": torch.full((2,), float(i))})
    assert [float(decode(f)[1]["w"][0]) for f in lb.pub_socket.frames] == [0.0, 3.0, 6.0]
    assert decode(lb.pub_socket.frames[0])[0] == Protocol.Model

    lb2 = LM.LearnerBase.__new__(LM.LearnerBase)
    lb2.args = SimpleNamespace(publish_interval=1, publish_async=True)
    lb2.pub_socket = Sock(delay=0.02)
    buf = torch.zeros(2)
    for i in range(20):
        buf.fill_(float(i))                                  # the caller reuses its snapshot buffer
        lb2.pub_model({"w": buf})
    lb2.pub_flush()
    got = [float(decode(f)[1]["w"][0]) for f in lb2.pub_socket.frames]
    assert got == sorted(got) and got[-1] == 19.0 and len(got) + lb2.pub_dropped == 20 and lb2.pub_dropped > 0
    lb2.pub_socket = None                                     # (__del__ closes real sockets only)
    lb.pub_socket = None


# --------------------------------------------------------------------------------------- batched inference (f4)
@pytest.mark.parametrize("cont", [False, True])
def test_act_batch_matches_per_environment_act(backend, cont):
    """`act_batch` (N environments, one fused-forward launch with T = 1) against the reference's per-environment `act`
    math (networks/models.py:37-56): new LSTM state, normalised log-probs, and the log-prob of the sampled action."""
    from b200rl.networks import models as M
    ops, dev = backend
    N, f, A = 37, (2 if cont else 4), (1 if cont else 3)
    torch.manual_seed(1)
    model = (M.MlpLSTMSingleContinuous if cont else M.MlpLSTMSingle)(f, A, 5, 64)
    sd = {k: v.clone() for k, v in model.actor.state_dict().items()}
    obs, hx, cx = torch.randn(N, f), 0.3 * torch.randn(N, 64), 0.3 * torch.randn(N, 64)
    # the reference's math, per environment, on the CPU
    x = torch.relu(obs @ sd["body.0.weight"].T + sd["body.0.bias"])
    g = x @ sd["lstmcell.weight_ih"].T + sd["lstmcell.bias_ih"] + hx @ sd["lstmcell.weight_hh"].T + sd["lstmcell.bias_hh"]
    i, fg, gg, o = g.chunk(4, 1)
    c1 = torch.sigmoid(fg) * cx + torch.sigmoid(i) * torch.tanh(gg)
    h1 = torch.sigmoid(o) * torch.tanh(c1)
    z = h1 @ sd["logits.weight"].T + sd["logits.bias"]
    net = model.to(dev).actor
    action, logits, logp, (h2, c2) = net.act_batch(obs.to(dev), (hx.to(dev), cx.to(dev)), ops=ops)
    _close(h2.cpu(), h1, rtol=1e-4, atol=1e-6, what="hx")
    _close(c2.cpu(), c1, rtol=1e-4, atol=1e-6, what="cx")
    if cont:
        zs = h1 @ sd["logits_std.weight"].T + sd["logits_std.bias"]
        want = torch.distributions.Normal(torch.tanh(z), torch.nn.functional.softplus(zs)).log_prob(action.cpu())
        assert action.shape == (N, 1) and float(logits.abs().max()) == 0.0
        _close(logp.cpu(), want, rtol=1e-4, atol=1e-5, what="log_prob of the sampled action")
    else:
        want = torch.distributions.Categorical(torch.softmax(z, -1)).logits
        assert action.shape == (N, 1) and logits.shape == (N, A)
        _close(logits.cpu(), want, rtol=1e-4, atol=1e-6, what="normalised log-probs")
        _close(logp.cpu(), want.gather(1, action.cpu().long()), rtol=1e-4, atol=1e-6, what="log_prob of the sampled action")


# ------------------------------------------------------------------------------------- rollout assembler (f2)
def test_native_assembler_matches_reference_stream():
    """The recorded stream of worker step records (tests/golden/make_golden.py --assembler: the reference's RolloutAssembler
    + LearnerStorage.make_batch under a scripted clock, with age expiry, parked episodes spliced with an is_fir seam, a
    full shared batch and a counter reset) replayed through the native assembler: same slot counter after every push,
    same bytes in the shared arrays."""
    from types import SimpleNamespace
    from b200rl.agents.storage_module.shared_batch import setup_shared_memory
    from b200rl.buffers import RolloutAssembler
    g = load_golden("assembler_stream")
    T, f, A, H, cap, reset_at = [int(x) for x in g["meta"]]
    args = SimpleNamespace(seq_len=T, action_space=A, hidden_size=H)
    shm = setup_shared_memory(args, [f], batch_size=cap)
    asm = RolloutAssembler(args).bind_shared(shm, [f])
    fields = ("obs", "act", "rew", "logits", "log_prob", "is_fir", "hx", "cx")
    for i in range(len(g["id"])):
        data = {k: torch.from_numpy(g[f"in/{k}"][i]) for k in fields}
        data["id"], data["done"] = int(g["id"][i]), torch.FloatTensor([float(g["done"][i])])
        asm.push_now(data, now=float(g["now"][i]))
        assert shm["batch_index"].value == int(g["count_after"][i]), f"slot counter after push {i}"
        if i == reset_at:
            shm["batch_index"].value = 0
    for k in fields:
        got = np.frombuffer(shm[f"{k}_batch"][0].get_obj(), dtype=np.float32)
        np.testing.assert_array_equal(got, g[f"shared/{k}"], err_msg=k)
    st = asm.stats()
    assert st["spliced"] > 0 and st["dropped_age"] > 0 and st["emitted"] >= cap, st


def test_native_assembler_queue_mode_and_learner_round_trip(backend, tmp_path):
    """Without shared memory the class behaves like the reference's (`pop()` yields dicts of (T, width) tensors); bound to the
    learner's shared arrays, T x B pushed steps become one batch the learner's own coroutines consume."""
    from types import SimpleNamespace
    from b200rl.buffers import RolloutAssembler
    ops, dev = backend
    T, f, A, H = 5, 4, 2, 64
    args = SimpleNamespace(seq_len=T, action_space=A, hidden_size=H)
    q = asyncio.Queue()
    asm = RolloutAssembler(args, q, clock=lambda: 0.0)
    rng = np.random.RandomState(0)
    steps = [{k: torch.from_numpy(rng.randn(w).astype(np.float32)) for k, w in
              dict(obs=f, act=1, rew=1, logits=A, log_prob=1, is_fir=1, hx=H, cx=H).items()} for _ in range(T)]

    async def go():
        for s in steps:
            await asm.push(dict(s, id=1, done=torch.FloatTensor([0.0])))
        return await asm.pop()
    out = asyncio.run(go())
    for k in ("obs", "logits", "hx"):
        np.testing.assert_array_equal(out[k].numpy(), np.stack([s[k].numpy() for s in steps]))
    assert out["obs"].shape == (T, f) and out["act"].shape == (T, 1)


# ---------------------------------------------------------------- SAC-Continuous: in-kernel noise, packed critic inputs, CUDA graphs
def _philox_normal_ref(seed, draw, n):
    """Philox4x32-10 + Box-Muller exactly as csrc/sac_kernels.cu states it (numpy, 64-bit integer arithmetic)."""
    i = np.arange(n, dtype=np.uint64)
    c = [i & np.uint64(0xFFFFFFFF), i >> np.uint64(32), np.full(n, draw & 0xFFFFFFFF, np.uint64), np.full(n, draw >> 32, np.uint64)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32)
    M0, M1, m32 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & m32, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & m32]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & m32, (k1 + np.uint64(0xBB67AE85)) & m32
    u0 = ((c[0] >> np.uint64(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    u1 = ((c[1] >> np.uint64(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    return np.sqrt(np.float32(-2.0) * np.log(u0)) * np.cos(np.float32(6.283185307179586) * u1)


def test_sac_continuous_in_kernel_noise_and_pack(backend):
    """b200rl_sac_cont_sample without injected noise draws N(0,1) from Philox4x32-10 (checked against a numpy restatement of the
    generator), advances the device-side draw counter once per launch, and writes the action into the strided [obs | act]
    buffer that b200rl_sac_cont_pack prepared (replacing torch.randn / torch.cat of sac_continuous/learning.py:47-72)."""
    from b200rl.native import ptr
    ops, dev = backend
    lib = ops.lib
    B, T, f, A = 24, 5, 7, 3
    g = torch.Generator().manual_seed(3)
    obs, act = torch.randn(B, T, f, generator=g).to(dev), torch.rand(B, T, A, generator=g).to(dev)
    ho = (0.3 * torch.randn(B, T, 2 * A, generator=g)).to(dev)
    ub = [torch.full((B, T, f + A), float("nan"), device=dev) for _ in range(3)]
    st = ops._stream(dev)
    lib.call("b200rl_sac_cont_pack", ptr(obs), ptr(act), ptr(ub[0]), ptr(ub[1]), ptr(ub[2]), B, T, f, A, st)
    for k in range(3):
        assert torch.equal(ub[k][..., :f], obs)
    assert torch.equal(ub[2][..., f:], act) and torch.isnan(ub[0][..., f:]).all()
    seed = 0x1234567887654321
    rng = torch.tensor([seed, 5, 0, 0], dtype=torch.int64, device=dev)
    eps = [torch.empty(B, T, A, device=dev) for _ in range(2)]
    logp = torch.empty(B, T, A, device=dev)
    for k in range(2):
        lib.call("b200rl_sac_cont_sample", ptr(ho), None, ptr(eps[k]), ptr(rng), ptr(ub[k]) + 4 * f, f + A, ptr(logp), B, T, A, st)
    assert rng.cpu().tolist() == [seed, 7, 0, 0]
    for k in range(2):
        ref = _philox_normal_ref(seed, 5 + k, B * T * A).reshape(B, T, A)
        np.testing.assert_allclose(eps[k].cpu().numpy(), ref, rtol=2e-5, atol=2e-6)
        mu, sd = ho[..., :A], ho[..., A:].clamp(-20, 2).exp()
        np.testing.assert_allclose(ub[k][..., f:].cpu().numpy(), torch.tanh(mu + sd * eps[k]).cpu().numpy(), rtol=1e-5, atol=1e-6)
        assert torch.equal(ub[k][..., :f], obs)
    # with injected noise the generator state is untouched
    lib.call("b200rl_sac_cont_sample", ptr(ho), ptr(eps[0]), None, None, ptr(ub[1]) + 4 * f, f + A, ptr(logp), B, T, A, st)
    assert torch.equal(ub[1][..., f:], ub[0][..., f:]) and rng.cpu().tolist() == [seed, 7, 0, 0]
    # statistics of a larger draw
    n = 1 << 16
    z = _philox_normal_ref(seed, 99, n)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02 and abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 0.02


@pytest.mark.gpu
def test_sac_phases_as_cuda_graphs_equal_eager(monkeypatch):
    """The two launch-bound phases of a SAC(-Continuous) update replayed from CUDA graphs give bit-identical parameters to the
    eager launches (same kernels, same order), including fresh in-kernel noise per replay (the draw counter lives on the device)."""
    from b200rl.engine import SACEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    dev = torch.device("cuda:0")
    for cont, (B, T, f, A) in ((True, (64, 5, 17, 6)), (False, (32, 6, 4, 2))):
        res = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("B200RL_GRAPH", mode)
            torch.manual_seed(4)
            model = ((M.MlpLSTMSeperateContinuous(f, A, T, 64, act_dim=A)) if cont else M.MlpLSTMSeperate(f, A, T, 64)).to(dev)
            eng = SACEngine(model, oracle.DEFAULT_HYPER, cont)
            batch = {k: v.to(dev) for k, v in make_rollouts(B, T, f, A, continuous=cont, act_dim=A if cont else 1).items()}
            for _ in range(5):
                eng.update(batch)
            L = eng.losses()
            res[mode] = (torch.cat([v.reshape(-1) for v in model.actor.state_dict().values()]).cpu(),
                         torch.cat([v.reshape(-1) for v in model.critic.state_dict().values()]).cpu(), L,
                         eng.rng.cpu().tolist(), 0 if eng.graphs is None else eng.graphs.replays)
        assert res["0"][4] == 0 and res["1"][4] >= 4, (res["0"][4], res["1"][4])
        assert torch.equal(res["0"][0], res["1"][0]) and torch.equal(res["0"][1], res["1"][1])
        assert res["0"][2] == res["1"][2]
        assert res["0"][3] == res["1"][3] and res["0"][3][1] == (10 if cont else 0)
