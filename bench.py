#!/usr/bin/env python
"""Learner-update benchmark (BASELINE.json metric: learner env-steps/s).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one learner update (forward over the (B,T) window, GAE, PPO loss, BPTT,
clip_grad_norm_, RMSprop) over one batch of synthetic rollouts.  Workload = the
north-star shape: PPO, CartPole-shape rollouts, B=4096 sequences x T=128 steps, obs=4,
A=2, per GPU (weak scaling: every rank owns its own 4096 sequences; one NCCL all-reduce
of the 135 KB flat gradient per step).  Rollouts are pre-generated (workers / manager /
storage are out of the learner path, as the north_star asks) and resident in HBM for
`value`; the `e2e` leg drives the reference-facing `learning()` coroutine with pinned
HOST batches, so it pays the H2D of every batch and a D2H read of the losses each step.

`--impl reference` times the reference's own algorithm on the host CPUs: the torch-CPU
oracle port (`oracle/`, the restatement pinned to the reference's outputs -- the
reference itself is pure Python and is not present on the GPU box).
"""
import argparse
import asyncio
import json
import os
import statistics
import sys
import threading
import time
from types import SimpleNamespace

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


def hyper_args(device):
    from b200rl.utils.utils import load_params
    a = load_params(os.path.join(PKG, "b200rl", "utils", "parameters.json"))
    a.algo, a.batch_size, a.seq_len, a.action_space = "PPO", B, T, A
    a.device = device
    a.result_dir = None
    a.model_dir = None
    return a


def usable_cpus():
    """Host threads we can really use: min(affinity mask, cgroup CPU quota).  The GPU boxes
    expose 128 logical CPUs under a 16-CPU cgroup quota; 128 torch threads there run ~10x
    slower than 16."""
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


# ------------------------------------------------------------------ reference arm
def reference_arm(steps, warmup, rank, world):
    """The reference's CPU learner (oracle port, all host threads) on the same workload.
    Each step is a bounded sample: a PPO update on `rows` of the 4096 sequences (T=128)."""
    if rank != 0:
        return
    import oracle
    from oracle import models as OM
    from b200rl.synth import make_rollouts
    torch.set_num_threads(usable_cpus())
    budget_steps = 2.5e7                                    # 2-3 min at the ~1.2e5 env-steps/s the 16-thread GPU-box host sustains
    rows = int(max(64, min(B, budget_steps / max(1, steps + warmup) / T)))
    rows = 1 << (rows.bit_length() - 1)
    torch.manual_seed(0)
    params = OM.init_params("single", F, A)
    learner = oracle.OracleLearner("PPO", params, fused_cell=True)
    batches = [{k: v[:rows].contiguous() for k, v in make_rollouts(B, T, F, A, batch_index=i).items()} for i in range(2)]
    for i in range(warmup):
        learner.update(batches[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        learner.update(batches[i % 2])
    dt = time.perf_counter() - t0
    val = rows * T * steps / dt
    sample = f"PPO update on {rows} of the {B} sequences (T={T}) per step, {steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "algo": "PPO", "sequences_per_step": rows, "seq_len": T},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline():
    """Oracle port timed on this host (bounded sample: 1 warm-up + 3 full-size updates)."""
    import oracle
    from oracle import models as OM
    from b200rl.synth import make_rollouts
    nthreads = usable_cpus()
    old = torch.get_num_threads()
    torch.set_num_threads(nthreads)
    torch.manual_seed(0)
    learner = oracle.OracleLearner("PPO", OM.init_params("single", F, A), fused_cell=True)
    batch = make_rollouts(B, T, F, A)
    t0 = time.perf_counter()
    learner.update(batch)
    n = 3 if time.perf_counter() - t0 < 10 else 1
    t0 = time.perf_counter()
    for _ in range(n):
        learner.update(batch)
    dt = time.perf_counter() - t0
    torch.set_num_threads(old)
    return {"value": B * T * n / dt, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": f"{n} full PPO updates at B={B}, T={T} after 1 warm-up ({dt:.1f} s of CPU work)"}


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    torch.manual_seed(0)                                     # same initial weights on every rank
    model = MlpLSTMSingle(F, A, T, H)
    args = hyper_args(dev)
    learner = LearnerSinglePPO(args, None, model, None, StopAfter(0), None, None, [F], process_group=pg)
    eng = learner.engine
    host = [make_rollouts(B, T, F, A, batch_index=rank * 1000 + i) for i in range(NBUF)]
    dev_batches = [{k: v.to(dev) for k, v in b.items()} for b in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- value: K updates, rollouts resident in HBM
    for i in range(a.warmup):
        eng.update(dev_batches[i % NBUF])
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = eng.n_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        eng.update(dev_batches[i % NBUF])
    e1.record()
    barrier()
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

    # ---- e2e: the reference-facing learning() coroutine fed with pinned HOST batches
    n_e2e = min(a.steps, 200)
    pinned = [tuple(t.pin_memory() for t in as_reference_tuple(b, T)[:6]) + (b["hx"].pin_memory(), b["cx"].pin_memory())
              for b in host]
    learner.args.loss_log_interval = 1                        # D2H read of the losses every step
    logged = []
    learner.log_loss_tensorboard = lambda timer, loss, detached: logged.append(float(loss))

    def run_coroutine(n):
        learner.stop_event = StopAfter(n)

        async def go():
            q = asyncio.Queue()
            for i in range(n):
                q.put_nowait(pinned[i % NBUF])
            learner.batch_queue = q
            await learning_ppo(learner, ExecutionTimer(num_transition=B * T))
        asyncio.run(go())

    run_coroutine(3)
    barrier()
    e0.record()
    run_coroutine(n_e2e)
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * T * n_e2e / (float(ms2.item()) * 1e-3)
    h2d = learner.to_gpu.bytes_last
    if rank == 0:
        log(f"e2e leg done: {float(ms2.item()) / n_e2e:.3f} ms/step")
    d2h = 8 * 8 + 4

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "algo": "PPO", "sequences_per_gpu": B, "seq_len": T, "obs": F, "actions": A,
                   "hidden": H, "global_batch_sequences": B * world, "parallelism": f"dp{world}",
                   "l2": "per-step working set (saved activations 0.8 GB) >> 126 MB L2; 4 input batches cycled"},
        "clocks": clk, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": n_e2e, "api": "b200rl.agents.learner_module.ppo.learning.learning(parent, timer)"},
        "loss": losses["loss"],
    }

    if rank == 0:
        # ---- per-kernel device times (alone, CUDA events) and the roofline of the dominant one
        ops, L, hp = eng.ops, eng.L, eng.hp
        bt = dev_batches[0]
        u = bt["obs"]
        stages = {
            "lstm_fwd": lambda i: ops.lstm_fwd(L, eng.flat, u, bt["hx"], bt["cx"], eng.ws),
            "head_fwd": lambda i: ops.head_fwd(L, eng.flat, eng.ws["h"], eng.head_out),
            "ppo_head": lambda i: ops.ppo_head(eng.head_out, bt, eng.dhead, eng.scalars, hp, A, False, None),
            "lstm_bwd": lambda i: ops.lstm_bwd(L, eng.flat, eng.dhead, eng.ws, bt["cx"]),
            "lstm_wgrad": lambda i: ops.lstm_wgrad(L, eng.flat, eng.ws, bt["hx"], u, eng.dhead, eng.partials, variant=eng.wgrad_variant),
            "reduce_partials": lambda i: ops.reduce_partials(eng.partials, eng.grad),
            "clip_rmsprop": lambda i: ops.clip_rmsprop(eng.flat, eng.grad, eng.sq_avg, L.P, 1.0, 40.0, 1e-4, 0.99, 1e-5, eng.grad_norm),
        }
        # run the chain once so every stage sees valid inputs, then time each alone
        for fn in stages.values():
            fn(0)
        kms = {k: event_time(fn, 10, dev) for k, fn in stages.items()}
        bt_ = B * T
        algo_bytes = {
            "lstm_fwd": 4 * (bt_ * (F + 6 * H) + 2 * B * H + L.P),
            "head_fwd": 4 * (bt_ * (H + L.NH)),
            "ppo_head": 4 * (bt_ * (2 * L.NH + 4)),
            "lstm_bwd": 4 * (bt_ * (4 * H + H + L.NH + 4 * H + H) + B * H),          # + dx (B,T,H) written by the tcgen05 BPTT kernel
            "lstm_wgrad": 4 * (bt_ * (4 * H + H + H + F + L.NH) + eng.nparts * L.P),    # dpre, h, dx, u, dhead
            "reduce_partials": 4 * (eng.nparts * L.P + L.P),
            "clip_rmsprop": 4 * 6 * L.P,
        }
        algo_flop = {"lstm_fwd": 2.0 * bt_ * (F * H + 2 * H * 4 * H), "lstm_bwd": 2.0 * bt_ * (4 * H * H),
                     "lstm_wgrad": 2.0 * bt_ * (3 * 4 * H * H + H * F + H * L.NH)}
        dom = max(kms, key=kms.get)
        peak, peak_src, peaks = load_peaks()
        ach = algo_bytes[dom] / (kms[dom] * 1e-3) / 1e9
        sm_max = clk.get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        out["roofline"] = {
            "kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": load_traffic(dom), "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes[dom],
            "note": "persistent recurrent kernel (per-step tcgen05 GEMM + cell phase, 128 dependent steps): bound by the per-step "
                    "serial latency (MMA issue ~31-44 cycles per M128xN32xK8 tf32 MMA, MUFU-bound cell phase), not by HBM; "
                    "`fp32` = algorithmic flop rate vs the FP32 FFMA peak for scale (the GEMMs run as 3xTF32 on tcgen05); see DESIGN.md 4",
            "fp32": {"achieved_tflops": algo_flop.get(dom, 0) / (kms[dom] * 1e-3) / 1e12, "peak_tflops": fp32_peak,
                     "frac": algo_flop.get(dom, 0) / (kms[dom] * 1e-3) / 1e12 / fp32_peak},
        }
        out["kernels_ms"] = kms
        out["kernels_hbm_gbs"] = {k: algo_bytes[k] / (v * 1e-3) / 1e9 for k, v in kms.items()}
        out["step_share"] = {k: v / sum(kms.values()) for k, v in kms.items()}
        if world == 1 and not a.no_cpu_baseline:
            log("timing the CPU port ...")
            out["cpu_baseline"] = cpu_baseline()
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
