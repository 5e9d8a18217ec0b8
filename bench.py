#!/usr/bin/env python
"""Learner-update benchmark (BASELINE.json metric: learner env-steps/s).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one learner update (forward over the (B,T) window, GAE, PPO loss, BPTT, clip_grad_norm_, RMSprop) over one
batch of synthetic rollouts.  Workload = the north-star shape: PPO, CartPole-shape rollouts, B=4096 sequences x T=128
steps, obs=4, A=2, per GPU (weak scaling: every rank owns its own 4096 sequences; the gradient exchange is fused with the
clip + RMSprop step in one kernel over NVLink peer memory -- no library collective on the stream).  Rollouts are
pre-generated (workers / manager / storage are out of the learner path, as the north_star asks) and resident in HBM for
`value`; the `e2e` leg drives the reference-facing `learning()` coroutine with HOST batches, so it pays the H2D of every
batch and a D2H read of the losses each step.

Extra keys (all measured in the same run): `roofline` (dominant kernel vs the measured HBM peak, plus tensor-pipe
utilisation of the 3xTF32 GEMMs and the end-to-end compulsory-bytes fraction), `kernels_ms`, `configs` (the other
BASELINE.json configs at their stated per-GPU shapes), `strong_scaling` (N > 1: global B = 4096 split over the ranks),
`dp_parity` (N > 1: sharded PPO / V-MPO / SAC-Continuous updates vs the full-batch oracle), `cpu_baseline` +
`parity_in_run` (N = 1: the reference's CPU learner on this host and the first GPU update checked against it).

`--impl reference` times the reference's own CPU learner on the host cores: the UNMODIFIED reference staged under
`oracle/_ref` (oracle/build_ref.py; `kind: "reference"`), else the torch-CPU oracle port (`kind: "port"`).
"""
import argparse
import asyncio
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

B, T, F, A, H = 4096, 128, 4, 2, 64
WORKLOAD = "ppo_cartpole_B4096_T128_obs4_A2"
METRIC = "learner_env_steps_per_sec"
UNIT = "env-steps/s"
NBUF = 4                       # distinct pre-generated batches cycled through


def config_dict(world):
    """Identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "algo": "PPO", "sequences_per_gpu": B, "seq_len": T, "obs": F, "actions": A,
            "hidden": H, "global_batch_sequences": B * world, "parallelism": f"dp{world}",
            "l2": "per-step working set (saved activations 0.8 GB) >> 126 MB L2; 4 input batches cycled"}


def product_hyper():
    """Hyper-parameters of the product's own parameters.json (the values the reference ships, utils/parameters.json:1-31)."""
    from b200rl.utils.utils import load_params, hyper_from_args
    return hyper_from_args(load_params(os.path.join(PKG, "b200rl", "utils", "parameters.json")))


def hyper_args(device, algo="PPO", b=B, t=T, a_=A):
    from b200rl.utils.utils import load_params
    a = load_params(os.path.join(PKG, "b200rl", "utils", "parameters.json"))
    a.algo, a.batch_size, a.seq_len, a.action_space = algo, b, t, a_
    a.device = device
    a.result_dir = None
    a.model_dir = None
    return a


def usable_cpus():
    """Host threads we can really use: min(affinity mask, cgroup CPU quota).  The GPU boxes expose 128 logical CPUs under
    a 16-CPU cgroup quota; 128 torch threads there run ~10x slower than 16."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per) + 0.999)))
    except Exception:  # noqa: BLE001
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, (q + per - 1) // per))
        except Exception:  # noqa: BLE001
            pass
    return max(1, n)


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._th = [], set(), None, threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, str(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._th = threading.Thread(target=self._run, daemon=True)
            self._th.start()

    def stop(self):
        self._stop.set()
        if self._th is not None:
            self._th.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------ CPU legs (the only places that execute oracle/)
def _ref_available():
    from oracle import ref_driver
    return ref_driver.available()


def cpu_learner(rows):
    """(kind, run(batches) -> (seconds per update list, losses list, flat params))."""
    from b200rl.synth import make_rollouts, as_reference_tuple
    if _ref_available():
        from oracle import ref_driver
        ref = ref_driver.ReferencePPO(F, A, T, rows)

        def run(idx):
            batches = [as_reference_tuple({k: v[:rows].contiguous() for k, v in make_rollouts(B, T, F, A, batch_index=i).items()}, T)
                       for i in idx]
            times, logs, _ = ref.run(batches)
            return times, [l[0] for l in logs], ref.flat_params()
        return "reference", run
    import oracle
    from oracle import models as OM
    torch.manual_seed(0)
    learner = oracle.OracleLearner("PPO", OM.init_params("single", F, A), fused_cell=True)

    def run(idx):
        times, losses, r = [], [], None
        for i in idx:
            b = {k: v[:rows].contiguous() for k, v in make_rollouts(B, T, F, A, batch_index=i).items()}
            t0 = time.perf_counter()
            r = learner.update(b)
            times.append(time.perf_counter() - t0)
            losses.append(float(r["loss"]))
        return times, losses, r["params"]
    return "port", run


def reference_arm(steps, warmup, rank, world):
    """The reference's CPU learner (all usable host threads) on the same workload.  Each step is one `learning()` iteration
    on `rows` of the 4096 sequences (T=128): all of them when the run fits a few minutes, else a bounded sample."""
    if rank != 0:
        return
    nthreads = usable_cpus()
    torch.set_num_threads(nthreads)
    budget_steps = 5.0e7                                    # ~4-5 min at the ~1.7e5 env-steps/s a 16-thread host sustains
    rows = int(max(64, min(B, budget_steps / max(1, steps + warmup) / T)))
    rows = 1 << (rows.bit_length() - 1)
    kind, run = cpu_learner(rows)
    run([i % 2 for i in range(warmup)])
    times, _, _ = run([i % 2 for i in range(steps)])
    dt = sum(times)
    val = rows * T * steps / dt
    what = "the UNMODIFIED reference's learning() from oracle/_ref" if kind == "reference" else "oracle port of the reference's learning()"
    sample = f"{what}: PPO update on {rows} of the {B} sequences (T={T}) per step, {steps} steps, {nthreads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": config_dict(world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_and_parity(dev):
    """Reference CPU learner timed on this host (bounded sample: 1 warm-up + up to 3 full-size updates), and the first GPU
    update checked against its first update on the same batch and initial weights."""
    from b200rl.engine import OnPolicyEngine
    from b200rl.networks.models import MlpLSTMSingle
    from b200rl.synth import make_rollouts
    import oracle
    nthreads = usable_cpus()
    old = torch.get_num_threads()
    torch.set_num_threads(nthreads)
    kind, run = cpu_learner(B)
    t0 = time.perf_counter()
    times, losses, params1 = run([0])                       # first update, batch 0: the parity reference (and the warm-up)
    params1 = params1.clone()
    n = 3 if time.perf_counter() - t0 < 12 else 1
    times, _, _ = run([1 + i % 2 for i in range(n)])
    dt = sum(times)
    torch.set_num_threads(old)
    base = {"value": B * T * n / dt, "unit": UNIT, "cores": nthreads, "kind": kind,
            "sample": f"{n} full PPO updates at B={B}, T={T} after 1 warm-up ({dt:.1f} s of CPU work), "
                      + ("UNMODIFIED reference learning() from oracle/_ref" if kind == "reference" else "oracle port")}
    # the same first update on the GPU, fresh weights (same seed -> bit-identical initial state dict, tests/test_oracle_golden.py)
    torch.manual_seed(0)
    model = MlpLSTMSingle(F, A, T, H).to(dev)
    eng = OnPolicyEngine(model.actor, "PPO", product_hyper())
    eng.update({k: v.to(dev) for k, v in make_rollouts(B, T, F, A, batch_index=0).items()})
    L = eng.losses()
    got = torch.cat([v.reshape(-1) for v in model.actor.views_of(eng.flat).values()]).cpu()
    loss_rel = abs(L["loss"] - losses[0]) / max(abs(losses[0]), 1e-12)
    perr = (got - params1).abs()
    tol = 1e-6 + 1e-4 * params1.abs()
    parity = {"against": kind, "what": "first PPO update at B=4096,T=128: total loss and parameters after the RMSprop step",
              "loss_gpu": L["loss"], "loss_cpu": losses[0], "loss_rel_err": loss_rel,
              "params_max_abs_err": float(perr.max()), "params_violations": int((perr > tol).sum()), "rtol": 1e-4,
              "pass": bool(loss_rel < 1e-4 and int((perr > tol).sum()) == 0)}
    return base, parity


# ----------------------------------------------------------------------- GPU arm
class StopAfter:
    def __init__(self, n):
        self.n, self.calls = n, 0

    def is_set(self):
        self.calls += 1
        return self.calls > self.n


def event_time(fn, iters, dev):
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / iters


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


def load_traffic(kernel):
    """dram bytes per launch from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(kernel)
    return None


class Timed:
    """K steps between CUDA events on the compute stream, barrier + synchronize on both sides, max over ranks."""

    def __init__(self, dev, world, dist):
        self.dev, self.world, self.dist = dev, world, dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def run(self, fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())


def e2e_from_shared_memory(dev, world, dist, timed, batch, n):
    """`n` learner iterations fed from real shared-memory arrays through the learner's own two coroutines."""
    import numpy as np
    from b200rl.agents.learner import LearnerSinglePPO
    from b200rl.agents.storage_module.shared_batch import reset_shared_on_policy_memory
    from b200rl.networks.models import MlpLSTMSingle
    args = hyper_args(dev)
    shm = reset_shared_on_policy_memory(args, [F])
    for k in ("obs", "act", "rew", "logits", "log_prob", "is_fir", "hx", "cx"):
        v = np.frombuffer(shm[f"{k}_batch"][0].get_obj(), dtype=np.float32)
        src = batch[k]
        if k in ("hx", "cx"):                                  # the storage process keeps the (T, H) trajectory of states per slot
            src = src[:, None, :].expand(-1, T, -1)
        v[:] = src.reshape(-1).numpy()

    class Stop:
        flag = False

        def is_set(self):
            return self.flag

    stop = Stop()
    torch.manual_seed(0)
    learner = LearnerSinglePPO(args, None, MlpLSTMSingle(F, A, T, H), shm, stop, None, None, [F])
    learner.args.loss_log_interval = 1
    learner.log_loss_tensorboard = lambda *a: None
    state = {"pubs": 0, "target": 0}

    def pub(sd):
        state["pubs"] += 1
        if state["pubs"] >= state["target"]:
            stop.flag = True
    learner.pub_model = pub

    async def storage():                                       # an always-ready storage process: the batch is "collected" again at once
        while not stop.flag:
            if learner.sh_data_num.value == 0:
                learner.sh_data_num.value = B
            await asyncio.sleep(0)

    def run(k):
        stop.flag, state["pubs"], state["target"] = False, 0, k
        learner.sh_data_num.value = B

        async def go():
            await asyncio.gather(learner.learning_chain_ppo(), storage())
        asyncio.run(go())

    run(3)
    timed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(n)
    e1.record()
    timed.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    info = {"zero_copy": bool(learner.zero_copy), "register_error": getattr(learner, "_register_error", None),
            "first_field_pinned": bool(learner.sample_batch_from_sh_memory("on-policy", copy=False)[0].is_pinned())}
    del learner
    return float(ms.item()), n, info


def dp_parity(dev, rank, world, pg):
    """Sharded updates over the real communicator (NCCL group + NVLink peer kernels) vs the full-batch oracle, small shapes:
    PPO (one fused exchange), V-MPO (all-gather of the advantages + global top-k), SAC-Continuous (two dependent exchanges)."""
    import oracle
    from collections import OrderedDict
    from b200rl.engine import OnPolicyEngine, VMPOEngine, SACEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    out = {}
    Bs, Ts = 16 * world, 16
    sh = Bs // world
    cut = lambda t: t[rank * sh:(rank + 1) * sh].contiguous().to(dev)
    for algo in ("PPO", "V-MPO"):
        torch.manual_seed(0)
        model = M.MlpLSTMSingle(4, 2, Ts, 64)
        sd = OrderedDict((k, v.clone()) for k, v in model.actor.state_dict().items())
        model = model.to(dev)
        full = make_rollouts(Bs, Ts, 4, 2)
        shard = {k: cut(v) for k, v in full.items()}
        if algo == "V-MPO":
            eng = VMPOEngine(model.actor, product_hyper(), process_group=pg)
            for _ in range(2):
                eng.update(shard, 0.0075)
        else:
            eng = OnPolicyEngine(model.actor, algo, product_hyper(), process_group=pg)
            for _ in range(2):
                eng.update(shard)
        L = eng.losses()
        params = torch.cat([v.reshape(-1) for v in model.actor.views_of(eng.flat).values()]).cpu()
        if rank == 0:
            orc = oracle.OracleLearner(algo, sd)
            for _ in range(2):
                r = orc.update(full, coef_alpha=torch.tensor(0.0075)) if algo == "V-MPO" else orc.update(full)
            perr = (params - r["params"]).abs()
            out[algo] = {"loss_rel_err": abs(L["loss"] - float(r["loss"])) / abs(float(r["loss"])),
                         "params_max_abs_err": float(perr.max()),
                         "pass": bool(abs(L["loss"] - float(r["loss"])) <= 1e-4 * abs(float(r["loss"])) + 1e-7
                                      and bool((perr <= 1e-6 + 1e-4 * r["params"].abs()).all())),
                         "fused_peer_step": eng.chan is not None and eng.chan.world == world}
    # SAC-Continuous
    f, Ad, Tq = 5, 2, 5
    torch.manual_seed(3)
    model = M.MlpLSTMSeperateContinuous(f, Ad, Tq, 64, act_dim=Ad)
    pa = OrderedDict((k, v.clone()) for k, v in model.actor.state_dict().items())
    pc = OrderedDict((k, v.clone()) for k, v in model.critic.state_dict().items())
    model = model.to(dev)
    eng = SACEngine(model, product_hyper(), True, process_group=pg)
    g = torch.Generator().manual_seed(11)
    batches = [make_rollouts(Bs, Tq, f, Ad, continuous=True, batch_index=i, act_dim=Ad) for i in range(2)]
    noise = [[torch.randn(Bs, Tq, Ad, generator=g), torch.randn(Bs, Tq, Ad, generator=g)] for _ in range(2)]
    for b, nz in zip(batches, noise):
        eng.update({k: cut(v) for k, v in b.items()}, noise=[cut(e) for e in nz])
    L = eng.losses()
    pa_got = torch.cat([v.reshape(-1) for v in model.actor.state_dict().values()]).cpu()
    pc_got = torch.cat([v.reshape(-1) for v in model.critic.state_dict().values()]).cpu()
    if rank == 0:
        orc = oracle.OracleLearner("SAC-Continuous", {"actor": pa, "critic": pc})
        for b, nz in zip(batches, noise):
            r = orc.update(b, noise=nz)
        ea, ec = (pa_got - r["actor_params"]).abs(), (pc_got - r["critic_params"]).abs()
        out["SAC-Continuous"] = {"loss_rel_err": abs(L["loss"] - float(r["loss"])) / abs(float(r["loss"])),
                                 "params_max_abs_err": float(max(ea.max(), ec.max())),
                                 "pass": bool(abs(L["loss"] - float(r["loss"])) <= 1e-4 * abs(float(r["loss"])) + 1e-7
                                              and bool((ea <= 1e-6 + 1e-4 * r["actor_params"].abs()).all())
                                              and bool((ec <= 1e-6 + 1e-4 * r["critic_params"].abs()).all())),
                                 "fused_peer_step": eng.chan_a is not None and eng.chan_a.world == world}
        out["shape"] = f"{Bs} sequences over {world} ranks, 2 updates each, vs the full-batch oracle (rtol 1e-4)"
    return out


def other_configs(dev, rank, world, pg, timed, steps=10, warmup=3):
    """The other BASELINE.json configs at their stated per-GPU shapes (cfg4 / cfg5 are 8-GPU configs: 1/8 of the global batch
    per rank, so at N=8 these ARE the stated configs; at other N the global batch is N/8 of it)."""
    from b200rl.engine import OnPolicyEngine, VMPOEngine, SACEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    res = {}

    def run(name, eng, batches, Bn, Tn, upd, steps=steps, warmup=warmup):
        ms = timed.run(lambda i: upd(eng, batches[i % len(batches)]), steps, warmup) / steps
        res[name] = {"ms_per_step": ms, "env_steps_per_sec": world * Bn * Tn / (ms * 1e-3), "sequences_per_gpu": Bn, "seq_len": Tn}
        del eng
        torch.cuda.empty_cache()

    torch.manual_seed(0)
    for name, algo, Bn, Tn, f, Aa, cont in (("cfg2_impala_B4096_T128", "IMPALA", 4096, 128, 4, 2, False),
                                             ("cfg3_ppo_continuous_B2048_T128", "PPO-Continuous", 2048, 128, 2, 1, True),
                                             ("cfg1_ppo_B64_T32", "PPO", 64, 32, 4, 2, False)):
        model = (M.MlpLSTMSingleContinuous if cont else M.MlpLSTMSingle)(f, Aa, Tn, 64).to(dev)
        eng = OnPolicyEngine(model.actor, algo, product_hyper(), process_group=pg)
        bs = [{k: v.to(dev) for k, v in make_rollouts(Bn, Tn, f, Aa, continuous=cont, batch_index=rank * 100 + i).items()} for i in range(2)]
        run(name, eng, bs, Bn, Tn, lambda e, b: e.update(b))
    model = M.MlpLSTMSingle(4, 2, 128, 64).to(dev)
    eng = VMPOEngine(model.actor, product_hyper(), process_group=pg)
    bs = [{k: v.to(dev) for k, v in make_rollouts(1024, 128, 4, 2, batch_index=rank * 100 + i).items()} for i in range(2)]
    run("cfg4_vmpo_1024_per_gpu_T128", eng, bs, 1024, 128, lambda e, b: e.update(b, 0.0075))
    model = M.MlpLSTMSeperateContinuous(17, 6, 5, 64, act_dim=6).to(dev)
    eng = SACEngine(model, product_hyper(), True, process_group=pg)
    eng_graphs = [eng.graphs]
    bs = [{k: v.to(dev) for k, v in make_rollouts(8192, 5, 17, 6, continuous=True, batch_index=rank * 100 + i, act_dim=6).items()} for i in range(2)]
    # (2 input batches x 2 gradient-buffer parities: each of the 4 launch sequences runs once eagerly and is captured once
    #  before it is replayed as a CUDA graph -- the warm-up covers that)
    run("cfg5_sac_continuous_8192_per_gpu_T5_obs17_act6", eng, bs, 8192, 5, lambda e, b: e.update(b), steps=40, warmup=12)
    res["cfg5_sac_continuous_8192_per_gpu_T5_obs17_act6"]["cuda_graph_replays"] = 0 if eng_graphs[0] is None else eng_graphs[0].replays
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip configs / strong scaling / dp parity (profiling runs)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        reference_arm(a.steps, max(a.warmup, 1), rank, world)
        return
    assert a.warmup >= 3, "timing rules: at least 3 warm-up steps"
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"

    import torch.distributed as dist
    from b200rl.agents.learner import LearnerSinglePPO
    from b200rl.agents.learner_module.ppo.learning import learning as learning_ppo
    from b200rl.networks.models import MlpLSTMSingle
    from b200rl.synth import make_rollouts, as_reference_tuple
    from b200rl.utils.utils import ExecutionTimer

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    timed = Timed(dev, world, dist)
    clocks = ClockSampler(local)                             # (NVML init happens HERE, not between the barrier and the timed region)

    out_extra = {}
    if world > 1 and not a.no_extras:
        out_extra["dp_parity"] = dp_parity(dev, rank, world, pg)
        if rank == 0:
            log("dp parity:", json.dumps(out_extra["dp_parity"]))

    torch.manual_seed(0)                                     # same initial weights on every rank
    model = MlpLSTMSingle(F, A, T, H)
    args = hyper_args(dev)
    learner = LearnerSinglePPO(args, None, model, None, StopAfter(0), None, None, [F], process_group=pg)
    eng = learner.engine
    host = [make_rollouts(B, T, F, A, batch_index=rank * 1000 + i) for i in range(NBUF)]
    dev_batches = [{k: v.to(dev) for k, v in b.items()} for b in host]

    # ---- value: K updates, rollouts resident in HBM
    for i in range(a.warmup):
        eng.update(dev_batches[i % NBUF])
    timed.barrier()
    timed.barrier()                                          # (second pass: every rank leaves within microseconds of the others)
    clocks.start()
    launches0 = eng.n_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        eng.update(dev_batches[i % NBUF])
    e1.record()
    timed.barrier()
    clk = clocks.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    launches = eng.n_launches - launches0
    value = world * B * T * a.steps / (total_ms * 1e-3)
    losses = eng.losses()
    if rank == 0:
        log(f"value leg done: {total_ms / a.steps:.3f} ms/step")

    # ---- e2e: the reference-facing learning() coroutine fed with HOST batches (pinned, as a storage process that writes into
    # a registered ring would provide them)
    n_e2e = min(a.steps, 200)
    pinned = [tuple(t.pin_memory() for t in as_reference_tuple(b, T)[:6]) + (b["hx"].pin_memory(), b["cx"].pin_memory())
              for b in host]
    learner.args.loss_log_interval = 1                        # D2H read of the losses every step
    logged = []
    learner.log_loss_tensorboard = lambda timer, loss, detached: logged.append(float(loss))

    def run_coroutine(n, batches):
        learner.stop_event = StopAfter(n)

        async def go():
            q = asyncio.Queue()
            for i in range(n):
                q.put_nowait(batches[i % NBUF])
            learner.batch_queue = q
            await learning_ppo(learner, ExecutionTimer(num_transition=B * T))
        asyncio.run(go())

    def e2e_leg(batches, n):
        run_coroutine(3, batches)
        timed.barrier()
        e0.record()
        run_coroutine(n, batches)
        e1.record()
        timed.barrier()
        m2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(m2, op=dist.ReduceOp.MAX)
        return float(m2.item())

    ms2 = e2e_leg(pinned, n_e2e)
    e2e_value = world * B * T * n_e2e / (ms2 * 1e-3)
    h2d = learner.to_gpu.bytes_last
    d2h = 8 * 8 + 4
    # the same from PAGEABLE host tensors (what `sample_batch_from_sh_memory`'s defensive copy hands over): one extra host
    # memcpy into the pinned staging set per step
    n_pg = min(n_e2e, 50)
    pageable = [tuple(t.clone() for t in p) for p in pinned]
    ms3 = e2e_leg(pageable, n_pg)
    # the whole seam from the storage side: rollouts sit in the shared-memory arrays `LearnerStorage.make_batch` fills
    # (agents/learner_storage.py:123-159); `learning_chain_ppo`'s two coroutines take them from there.  The arrays are
    # page-locked once (cudaHostRegister), every batch is DMA'd straight out of the slots, and the slot counter is handed back to
    # the (here emulated: always-ready) storage process as soon as the copy engine has read them.
    shm_ms, n_shm, shm_info = e2e_from_shared_memory(dev, world, dist, timed, host[0], min(n_e2e, 100))
    if rank == 0:
        log(f"e2e leg done: {ms2 / n_e2e:.3f} ms/step pinned, {ms3 / n_pg:.3f} ms/step pageable, {shm_ms / n_shm:.3f} ms/step from shared memory")

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic", "config": config_dict(world),
        "clocks": clk, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": n_e2e, "api": "b200rl.agents.learner_module.ppo.learning.learning(parent, timer)",
                "host_batches": "pinned",
                "pageable_value": world * B * T * n_pg / (ms3 * 1e-3), "pageable_steps": n_pg,
                "shared_memory_value": world * B * T * n_shm / (shm_ms * 1e-3), "shared_memory_steps": n_shm,
                "shared_memory_info": shm_info,
                "shared_memory_api": "LearnerSinglePPO.learning_chain_ppo(): put_batch_to_batch_q (page-locked shm slots -> DMA) + learning()"},
        "loss": losses["loss"],
        "exchange": ("fused peer step (one-shot NVLink all-reduce + clip + RMSprop, csrc/peer_kernels.cu)"
                     if eng.chan is not None else "NCCL all-reduce + single-CTA clip/RMSprop"),
    }
    out.update(out_extra)

    # ---- the other BASELINE configs and the strong-scaling variant (every rank takes part)
    if not a.no_extras:
        out["configs"] = other_configs(dev, rank, world, pg, timed)
        if world > 1:
            Bs = B // world
            sb = [{k: v[:Bs].contiguous() for k, v in b.items()} for b in dev_batches[:2]]
            ms_s = timed.run(lambda i: eng.update(sb[i % 2]), 20, 5) / 20
            eng.update(dev_batches[0])                        # (back to the full-size workspaces for the per-kernel timings below)
            timed.barrier()
            out["strong_scaling"] = {"global_batch_sequences": B, "sequences_per_gpu": Bs, "ms_per_step": ms_s,
                                     "value": B * T / (ms_s * 1e-3), "unit": UNIT,
                                     "note": "the recurrent kernels are bound by 128 dependent timesteps per sequence tile; shards that leave more "
                                             "than half of the SMs without a 32-sequence tile run 16-sequence tiles (half the cell phase per "
                                             "step); below one 16-sequence tile per SM the step time stops shrinking"}

    if rank == 0:
        # ---- per-kernel device times (alone, CUDA events) and the roofline of the dominant one
        ops, L, hp = eng.ops, eng.L, eng.hp
        bt = dev_batches[0]
        u = bt["obs"]
        gbuf, sbuf = (eng.chan.grad_views[0], eng.chan.scal_views[0]) if eng.chan is not None else (eng.grad, eng.scalars)
        # the kernels of the update, as the engine launches them (fused variants), each timed alone
        stages = {
            "lstm_fwd": lambda i: ops.net_fwd(L, eng.flat, u, bt["hx"], bt["cx"], eng.ws, eng.head_out),          # fused forward (torso + projection + recurrence + heads)
            "ppo_head": lambda i: ops.ppo_head(eng.head_out, bt, eng.dhead, sbuf, hp, A, False, None),
            "lstm_bwd": lambda i: ops.lstm_bwd(L, eng.flat, eng.dhead, eng.ws, bt["cx"], u=u, partials=eng.partials),   # BPTT + head / torso gradients
            "lstm_wgrad": lambda i: ops.lstm_wgrad(L, eng.flat, eng.ws, bt["hx"], u, eng.dhead, eng.partials, variant=eng.wgrad_variant),
            "reduce_partials": lambda i: ops.reduce_partials(eng.partials, gbuf),
        }
        if world == 1:      # (the fused step handshakes with the peers: it cannot be timed alone on one rank of many)
            scratch_p, scratch_s, scratch_g = eng.flat.clone(), eng.sq_avg.clone(), torch.empty_like(eng.grad)
            if eng.chan is not None:
                def opt_step(i):
                    eng.chan.begin()
                    eng.chan.step(0, scratch_p, scratch_s, None, None, scratch_g, eng.scalars, eng.grad_norm, 1.0, 40.0, 1e-4,
                                  0.99, 0.0, 1e-5, 0, ops._stream(dev))
                stages["clip_rmsprop"] = opt_step
            else:
                stages["clip_rmsprop"] = lambda i: ops.clip_rmsprop(scratch_p, eng.grad, scratch_s, L.P, 1.0, 40.0, 1e-4, 0.99, 1e-5, eng.grad_norm)
        for fn in stages.values():                          # run the chain once so every stage sees valid inputs
            fn(0)
        kms = {k: event_time(fn, 10, dev) for k, fn in stages.items()}
        bt_ = B * T
        fused_bwd = bool(eng.ws.get("small_grads_done"))
        algo_bytes = {
            # obs in; gates, c, h out; partial heads out + in, head_out out; initial state; parameters
            "lstm_fwd": 4 * (bt_ * (F + 6 * H + 4 * L.NH) + 2 * B * H + L.P),
            "ppo_head": 4 * (bt_ * (2 * L.NH + 4)),
            # gates + c + dhead (+ obs) in, dpre out (+ dx out when the torso gradient is not fused)
            "lstm_bwd": 4 * (bt_ * (4 * H + H + L.NH + 4 * H + (F if fused_bwd else H)) + B * H),
            # dpre, h (+ dx, dhead for the separate torso / head passes), u in; one slab row per CTA out
            "lstm_wgrad": 4 * (bt_ * (4 * H + H + F + (0 if fused_bwd else H + L.NH)) + eng.nparts * L.P),
            "reduce_partials": 4 * (eng.nparts * L.P + L.P),
            "clip_rmsprop": 4 * 6 * L.P,
        }
        # model FLOPs and what the tensor pipe really executes: every GEMM runs as three TF32 passes
        algo_flop = {"lstm_fwd": 2.0 * bt_ * (F * H + 2 * H * 4 * H + H * L.NH), "lstm_bwd": 2.0 * bt_ * (2 * 4 * H * H),
                     "lstm_wgrad": 2.0 * bt_ * (2 * 4 * H * H + H * F + H * L.NH)}
        dom = max(kms, key=kms.get)
        peak, peak_src, peaks = load_peaks()
        ach = algo_bytes[dom] / (kms[dom] * 1e-3) / 1e9
        tf32_peak = peaks.get("bf16_tflops_sustained", 1417.8) / 2.0
        gemm_ms = sum(kms[k] for k in algo_flop)
        issued = 3.0 * sum(algo_flop.values())
        out["roofline"] = {
            "kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": load_traffic(dom), "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes[dom],
            "limited_by": "latency: persistent recurrent kernel, 128 dependent timesteps per 32-sequence tile (80 small tcgen05 MMAs -> "
                          "accumulator hand-over -> TMEM load -> MUFU cell phase -> operand rewrite); neither the HBM nor the "
                          "tensor roofline is reached -- `frac` is the HBM-roofline fraction the metric asks for",
            "tensor_pipe": {"kernels": sorted(algo_flop), "model_tflops": sum(algo_flop.values()) / (gemm_ms * 1e-3) / 1e12,
                            "issued_tflops_3xtf32": issued / (gemm_ms * 1e-3) / 1e12, "peak_tflops_tf32": tf32_peak,
                            "frac": issued / (gemm_ms * 1e-3) / 1e12 / tf32_peak,
                            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (dense TF32 runs at half the bf16 rate)"},
            "end_to_end": {"compulsory_bytes_per_env_step": 36, "achieved_gbs": value / world * 36 / 1e9,
                           "frac_of_hbm_peak": value / world * 36 / 1e9 / peak,
                           "note": "SURVEY 8d: inputs read once per update (obs 16 B + act/rew/log_prob/is_fir 16 B + hx0,cx0 "
                                   "512 B / T); the step is LSTM-latency-bound, not bound by these bytes"},
        }
        out["kernels_ms"] = kms
        out["kernels_hbm_gbs"] = {k: algo_bytes[k] / (v * 1e-3) / 1e9 for k, v in kms.items()}
        out["kernels_hbm_frac"] = {k: algo_bytes[k] / (v * 1e-3) / 1e9 / peak for k, v in kms.items()}
        out["step_share"] = {k: v / sum(kms.values()) for k, v in kms.items()}
        if world == 1 and not a.no_cpu_baseline:
            log("timing the CPU reference ...")
            out["cpu_baseline"], out["parity_in_run"] = cpu_baseline_and_parity(dev)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
