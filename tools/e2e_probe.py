"""Why is the e2e step longer than the device-resident step? (dev tool)  Runs bench.py's e2e leg under a few switches:
host batches vs device-resident batches, pipelined vs synchronous schedule, and reports ms/step plus a CPU timeline."""
import asyncio
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200"))
import torch
import bench as BN

from b200rl.agents.learner import LearnerSinglePPO
from b200rl.agents.learner_module.ppo.learning import learning as learning_ppo
from b200rl.networks.models import MlpLSTMSingle
from b200rl.synth import make_rollouts, as_reference_tuple
from b200rl.utils.utils import ExecutionTimer

B, T, F, A, H = BN.B, BN.T, BN.F, BN.A, BN.H
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = MlpLSTMSingle(F, A, T, H)
learner = LearnerSinglePPO(BN.hyper_args(dev), None, model, None, BN.StopAfter(0), None, None, [F])
host = [make_rollouts(B, T, F, A, batch_index=i) for i in range(4)]
pinned = [tuple(t.pin_memory() for t in as_reference_tuple(b, T)[:6]) + (b["hx"].pin_memory(), b["cx"].pin_memory()) for b in host]
resident = [tuple(t.to(dev) for t in p) for p in pinned]
learner.args.loss_log_interval = 1
learner.log_loss_tensorboard = lambda timer, loss, detached: None


def run(batches, n):
    learner.stop_event = BN.StopAfter(n)

    async def go():
        q = asyncio.Queue()
        for i in range(n):
            q.put_nowait(batches[i % 4])
        learner.batch_queue = q
        await learning_ppo(learner, ExecutionTimer(num_transition=B * T))
    asyncio.run(go())


for label, batches, pipe in (("host batches, pipelined", pinned, "1"), ("host batches, synchronous", pinned, "0"),
                             ("device-resident batches, pipelined", resident, "1"), ("device-resident batches, synchronous", resident, "0")):
    os.environ["B200RL_PIPELINE"] = pipe
    run(batches, 5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    run(batches, 50)
    e1.record()
    torch.cuda.synchronize()
    print(f"{label:40s}: {e0.elapsed_time(e1) / 50:.3f} ms/step (wall {1e3 * (time.perf_counter() - t0) / 50:.3f})")

# interference test: device-resident updates while a side stream copies 18.9 MB host->device per update
eng = learner.engine
side = torch.cuda.Stream(dev)
bd = dict(zip(("obs", "act", "rew", "logits", "log_prob", "is_fir", "hx", "cx"), resident[0]))
bd["hx"], bd["cx"] = bd["hx"].contiguous(), bd["cx"].contiguous()
dst = [torch.empty_like(t, device=dev) for t in pinned[0]]
for with_copy in (False, True):
    for _ in range(3):
        eng.update(bd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        eng.update(bd)
        if with_copy:
            with torch.cuda.stream(side):
                for d, s in zip(dst, pinned[i % 4]):
                    d.copy_(s, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"engine.update x50, concurrent H2D on a side stream = {with_copy}: {e0.elapsed_time(e1) / 50:.3f} ms/step")

# how long does the 18.9 MB H2D take while an update is running?
for busy in (False, True):
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for i in range(10):
        if busy:
            eng.update(bd); eng.update(bd)
        with torch.cuda.stream(side):
            c0.record(side)
            for d, s in zip(dst, pinned[i % 4]):
                d.copy_(s, non_blocking=True)
            c1.record(side)
        torch.cuda.synchronize()
        tot += c0.elapsed_time(c1)
    print(f"H2D of one batch (8 copies, 18.9 MB), GPU busy with updates = {busy}: {tot / 10:.3f} ms")

# CPU timeline of the pipelined loop with host batches
import collections
acc = collections.Counter()


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def w(*a, **k):
        t = time.perf_counter()
        r = fn(*a, **k)
        acc[label] += time.perf_counter() - t
        return r
    setattr(obj, name, w)


stg = learner.to_gpu
learner.to_gpu = type("S", (), {"__call__": lambda self, *a: stg(*a), "prefetch": lambda self, *a: stg.prefetch(*a), "bytes_last": 0})()
wrap(learner.to_gpu.__class__, "__call__", "to_gpu")
wrap(learner.to_gpu.__class__, "prefetch", "prefetch")
wrap(eng, "update", "update enqueue")
wrap(eng, "snapshot_async", "snapshot_async")
wrap(eng, "snapshot_wait", "snapshot_wait (blocks on the GPU)")
wrap(learner, "pub_model", "pub_model")
os.environ["B200RL_PIPELINE"] = "1"
run(pinned, 5)
acc.clear()
t0 = time.perf_counter()
run(pinned, 50)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(f"CPU timeline, host batches pipelined: {1e3 * tot / 50:.3f} ms/step wall; " + ", ".join(f"{k} {1e3 * v / 50:.3f}" for k, v in acc.items()))

# GPU timeline: when do the updates and the H2D copies actually run?
base = torch.cuda.Event(enable_timing=True)
marks = []
upd = eng.update


def upd_ev(batch):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = upd(batch)
    b.record()
    marks.append(("U", a, b))
    return r


eng.update = upd_ev
ship = stg._ship


def ship_ev(args, slot):
    s = torch.cuda.current_stream()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    r = ship(args, slot)
    b.record(s)
    marks.append(("H2D", a, b))
    return r


stg._ship = ship_ev
run(pinned, 5)
marks.clear()
torch.cuda.synchronize()
base.record()
run(pinned, 8)
torch.cuda.synchronize()
print("GPU timeline (ms from start): " + "  ".join(f"{k}[{base.elapsed_time(a):.2f}-{base.elapsed_time(b):.2f}]" for k, a, b in marks))
