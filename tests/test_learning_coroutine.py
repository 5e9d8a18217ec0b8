"""The reference-facing seam: `async def learning(parent, timer)` driven exactly like the
reference drives it (a Learner object, an asyncio queue of host batch tuples, a stop
event), compared with what the reference's own coroutine produced (golden fixtures)."""
import asyncio
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import load_golden, golden_sd


class StopAfter:
    def __init__(self, n):
        self.n, self.calls = n, 0

    def is_set(self):
        self.calls += 1
        return self.calls > self.n


@pytest.mark.parametrize("name,algo", [("ppo_small", "PPO"), ("impala_small", "IMPALA"), ("ppo_continuous_small", "PPO-Continuous"),
                                       ("vmpo_small", "V-MPO")])
def test_learning_coroutine_matches_reference(backend, name, algo, tmp_path):
    from b200rl.agents import learner as LM
    from b200rl.agents.learner_module.ppo.learning import learning as learn_ppo
    from b200rl.agents.learner_module.impala.learning import learning as learn_impala
    from b200rl.agents.learner_module.v_mpo.learning import learning as learn_vmpo
    from b200rl.networks import models as M
    from b200rl.utils.utils import ExecutionTimer, load_params
    ops, dev = backend
    g = load_golden(name)
    B, T, f, A, H, n, cont = [int(x) for x in g["meta"]]
    args = load_params()
    args.algo, args.batch_size, args.seq_len, args.action_space = algo, B, T, A
    args.device, args.result_dir, args.model_dir = dev, None, str(tmp_path)
    args.loss_log_interval = 1
    model = (M.MlpLSTMSingleContinuous if cont else M.MlpLSTMSingle)(f, A, T, H)
    model.actor.load_state_dict(golden_sd(g, "sd0"))
    cls = {"IMPALA": LM.LearnerSingleIMPALA, "V-MPO": LM.LearnerSingleVMPO}.get(algo, LM.LearnerSinglePPO)
    learner = cls(args, None, model, None, StopAfter(n), None, None, [f], ops=None if dev.type == "cuda" else ops)
    logs, pubs = [], []
    if algo == "V-MPO":
        coefs = iter(torch.tensor(c) for c in g["coef_alpha"])
        learner.get_coef_alpha = lambda: next(coefs)
    learner.log_loss_tensorboard = lambda timer, loss, d: logs.append((float(loss), dict(d)))
    learner.pub_model = lambda sd: pubs.append(sd)
    keys = ("obs", "act", "rew", "logits", "log_prob", "is_fir", "hx", "cx")

    async def go():
        q = asyncio.Queue()
        for i in range(n):
            b = {k: torch.from_numpy(g[f"in/{i}/{k}"]) for k in keys}
            # the reference hands (B, T, H) hx/cx; the stager cuts them to [:, 0]
            b["hx"] = b["hx"][:, None, :].expand(-1, T, -1).contiguous()
            b["cx"] = b["cx"][:, None, :].expand(-1, T, -1).contiguous()
            q.put_nowait(tuple(b[k] for k in keys))
        learner.batch_queue = q
        fn = {"IMPALA": learn_impala, "V-MPO": learn_vmpo}.get(algo, learn_ppo)
        await fn(learner, ExecutionTimer(num_transition=B * T))

    asyncio.run(go())
    assert learner.idx == n and len(logs) == n and len(pubs) == n
    for i in range(n):
        np.testing.assert_allclose(logs[i][0], g["out/loss"][i], rtol=1e-4, atol=1e-7)
        keys_ = ("policy-loss", "value-loss", "loss-temperature", "loss-alpha") if algo == "V-MPO" else ("policy-loss", "value-loss", "policy-entropy")
        for k in keys_:
            np.testing.assert_allclose(float(logs[i][1][k]), g["out/" + k][i], rtol=1e-4, atol=1e-7)
        if algo != "V-MPO":
            r = g[f"out/ratio/{i}"]
            np.testing.assert_allclose(float(logs[i][1]["ratio"].mean()), r.mean(), rtol=1e-4)
        # published state dict: reference keys, order, shapes and values
        ref_keys = [k[4:] for k in g if k.startswith("sd0/")]
        assert list(pubs[i].keys()) == ref_keys
        flat = torch.cat([v.reshape(-1) for v in pubs[i].values()]).numpy()
        np.testing.assert_allclose(flat, g[f"out/params/{i}"], rtol=1e-4, atol=1e-6)
    # checkpoint written at idx 0 under the reference's file name, loadable, reference keys
    ck = torch.load(tmp_path / f"{algo}_0.pt", weights_only=False)
    assert list(ck.actor.state_dict().keys()) == ref_keys


@pytest.mark.gpu
def test_host_stager_prefetch_is_consumed_and_buffers_rotate():
    """`prefetch(*batch)` followed by `stager(*batch)` must hand out the prefetched device tensors (no second H2D), and a
    buffer set must not be overwritten while the previous batch may still be read."""
    from b200rl.utils.utils import HostStager
    dev = torch.device("cuda:0")
    st = HostStager(dev)
    B, T, f = 8, 4, 3
    g = torch.Generator().manual_seed(0)

    def host_batch():
        return (torch.randn(B, T, f, generator=g).pin_memory(), torch.randn(B, T, 1, generator=g).pin_memory(),
                torch.randn(B, T, 1, generator=g).pin_memory(), torch.randn(B, T, 2, generator=g).pin_memory(),
                torch.randn(B, T, 1, generator=g).pin_memory(), torch.randn(B, T, 1, generator=g).pin_memory(),
                torch.randn(B, T, 5, generator=g), torch.randn(B, T, 5, generator=g))       # hx, cx as (B, T, H), pageable

    batches = [host_batch() for _ in range(4)]
    d0 = st(*batches[0])                                  # first batch: synchronous path
    assert st.prefetch_hits == 0
    for i in range(1, 4):
        st.prefetch(*batches[i])
        keep = [t.clone() for t in d0 if t is not None]   # what the "update" of the previous batch reads
        d1 = st(*batches[i])
        assert st.prefetch_hits == i
        torch.cuda.synchronize()
        for a, b in zip(keep, [t for t in d0 if t is not None]):
            torch.testing.assert_close(a, b, rtol=0, atol=0)          # the previous set was not clobbered by the prefetch
        for j, (h, d) in enumerate(zip(batches[i], d1)):
            want = h[:, 0] if j >= 6 else h
            torch.testing.assert_close(d.cpu(), want, rtol=0, atol=0)
        d0 = d1
