"""Drive the UNMODIFIED reference's `learning()` coroutine from `oracle/_ref/` (staged by oracle/build_ref.py).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by bench.py's CPU legs and by tests, never by the product.
The harness is SURVEY.md 8c's: a fake `parent` (pre-filled asyncio queue, no-op publish, loss capture hook, a stop event
that trips after n iterations), the reference's own model classes, optimizer constructors as at agents/learner.py:70, and
the one stub the reference needs in this image (`blosc2`, imported by utils/utils.py:9 but never touched by the step).
"""
import asyncio
import contextlib
import copy
import io
import os
import sys
import time
import types
from functools import partial
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, ".staged"))


_mods = {}


def _load():
    """Import the reference packages from oracle/_ref.  They keep their own top-level names (`agents`, `networks`, `utils`,
    `buffers`): the reference pickles its model by class path `networks.models.*` (ppo/learning.py:113-116), so the modules
    must stay importable.  Nothing in this repo uses those top-level names (the product is namespaced under `b200rl`)."""
    if _mods:
        return _mods
    import torch  # noqa: F401
    saved_cwd = os.getcwd()
    sys.modules.setdefault("blosc2", types.ModuleType("blosc2"))
    os.chdir(REF)                                   # utils/utils.py:24-26 opens utils/parameters.json relative to the cwd
    if REF not in sys.path:
        sys.path.insert(0, REF)
    flag = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        from utils.utils import Params, ExecutionTimer, make_gpu_batch
        from agents.learner_module.ppo.learning import learning as learning_ppo
        from networks import models as ref_models
        _mods.update(Params=Params, ExecutionTimer=ExecutionTimer, make_gpu_batch=make_gpu_batch,
                     learning_ppo=learning_ppo, models=ref_models)
    finally:
        sys.dont_write_bytecode = flag
        os.chdir(saved_cwd)
    return _mods


class _StopAfter:
    def __init__(self, n):
        self.n, self.calls = n, 0

    def is_set(self):
        self.calls += 1
        return self.calls > self.n


class ReferencePPO:
    """The reference's PPO learner on CPU: its model class under `torch.manual_seed(seed)`, RMSprop as at
    agents/learner.py:70, its verbatim `learning()` (agents/learner_module/ppo/learning.py:13-126)."""

    def __init__(self, f, A, T, B, seed=0, hidden=64):
        import torch
        from torch.optim import RMSprop
        m = _load()
        torch.manual_seed(seed)
        self.model = m["models"].MlpLSTMSingle(f, A, T, hidden)
        a = copy.copy(m["Params"])
        a.algo, a.batch_size, a.seq_len, a.action_space = "PPO", B, T, A
        a.device = torch.device("cpu")
        a.model_dir = None
        a.loss_log_interval = 1
        a.model_save_interval = 10 ** 9
        self.args, self.T, self.B = a, T, B
        self.optimizer = RMSprop(self.model.parameters(), lr=a.lr, eps=1e-5)
        self.m = m

    def initial_state_dict(self):
        return {k: v.detach().clone() for k, v in self.model.actor.state_dict().items()}

    def run(self, batches):
        """`batches`: list of the 8-tuples `sample_batch_from_sh_memory` returns.  Runs one `learning()` iteration per
        batch; returns (per-iteration wall seconds, list of (loss, detached_losses) as floats)."""
        import torch
        m, args = self.m, self.args
        logs, times = [], []
        last = [time.perf_counter()]

        def pub(sd):
            now = time.perf_counter()
            times.append(now - last[0])
            last[0] = now

        parent = SimpleNamespace(args=args, model=self.model, stop_event=_StopAfter(len(batches)), heartbeat=None,
                                 to_gpu=partial(m["make_gpu_batch"], device=args.device), device=args.device,
                                 optimizer=self.optimizer, pub_model=pub,
                                 log_loss_tensorboard=lambda timer, loss, dl: logs.append(
                                     (float(loss.detach()), {k: float(v) if k != "ratio" else float(v.mean()) for k, v in dl.items()})))
        timer = m["ExecutionTimer"](num_transition=self.B * self.T)

        async def go():
            q = asyncio.Queue()
            for b in batches:
                q.put_nowait(b)
            parent.batch_queue = q
            last[0] = time.perf_counter()
            await m["learning_ppo"](parent, timer)

        # the checkpoint at idx 0 (ppo/learning.py:113) needs a directory: give it a throw-away one
        import tempfile
        with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
            args.model_dir = tmp
            asyncio.run(go())
        thr = list(timer.throughput_dict.get("learner-throughput", []))
        return times, logs, thr

    def flat_params(self):
        import torch
        return torch.cat([v.detach().reshape(-1) for v in self.model.actor.state_dict().values()])
