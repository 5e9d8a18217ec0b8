"""profiles/<tag>_ncu_summary.md + profiles/traffic.json from one tools/gpu_prof_r2.sh run (dev tool).
   python tools/make_profile_summary_r2.py r2j [bench_json]   (reads gpurun_out/*_<tag>.*)"""
import csv
import json
import shutil
import subprocess
import sys

tag = sys.argv[1]
bench_path = sys.argv[2] if len(sys.argv) > 2 else f"gpurun_out/{tag}_bench.json"
shutil.copy(bench_path, f"profiles/{tag}_bench.json")
shutil.copy(f"gpurun_out/launches_{tag}.csv", f"profiles/{tag}_launches_bench.csv")
shutil.copy(f"gpurun_out/scan_bench_{tag}.jsonl", f"profiles/{tag}_scan_bench.jsonl")
b = json.loads([l for l in open(f"profiles/{tag}_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
run = lambda *a: subprocess.run(["python", "tools/summarize_ncu.py", *a], capture_output=True, text=True).stdout
launch = run("launches", f"profiles/{tag}_launches_bench.csv")
raw = run("raw", f"gpurun_out/raw_{tag}.csv")
rows = [l.split("|") for l in launch.splitlines() if l.startswith("| ") and not l.startswith("| kernel") and "**sum**" not in l]
ms = {r[1].strip(): float(r[3]) for r in rows}
pick = lambda sub: [k for k in ms if sub in k]
stage = {"lstm_fwd": pick("lstm_fwd_fused"), "ppo_head": pick("ppo_head"), "lstm_bwd": pick("lstm_bwdt"),
         "lstm_wgrad": pick("wgrad_tc") + pick("torso_grad") + pick("head_grad"), "reduce_partials": pick("reduce_partials"),
         "clip_rmsprop": pick("peer_step")}
ntot = sum(ms.values())
km = b["kernels_ms"]
tot = sum(km.values())
out = [f"# {tag} -- PPO B=4096 T=128 obs=4 A=2, one B200\n", "## Launch list of the bench command\n",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:<our kernels> -s 18 -c 36 --csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras`",
       f"(raw: `profiles/{tag}_launches_bench.csv`; per-launch times under ncu are cold-cache and serialised -- compare SHARES).\n", launch,
       f"bench.py's own CUDA-event times of the same kernels, each timed alone (`profiles/{tag}_bench.json` `kernels_ms`):\n",
       "| stage | bench ms | bench share | ncu share |", "|---|---|---|---|"]
for k, v in km.items():
    n = sum(ms[x] for x in stage.get(k, []))
    out.append(f"| {k} (= {' + '.join(x.replace('void ', '') for x in stage.get(k, []))}) | {v:.3f} | {100 * v / tot:.1f} % | {100 * n / ntot:.1f} % |")
cb = b.get("cpu_baseline") or {}
out.append(f"\nbench: value {b['value']:.3e} env-steps/s, {b['ms_per_step']:.3f} ms/step, {b['gpu_launches'] // b['steps']} kernel launches per update; "
           f"e2e (host batches through `learning()`) {b['e2e']['value']:.3e}, from shared memory {b['e2e'].get('shared_memory_value', 0):.3e}; "
           + (f"CPU {cb.get('kind')} {cb['value']:.3e} on {cb['cores']} cores; " if cb else "")
           + f"parity_in_run {json.dumps(b.get('parity_in_run'))}; clocks {b['clocks']['sm_mhz']} MHz, throttle reasons {b['clocks']['reasons']}.\n")
out.append("## `ncu --set full` of every kernel of one PPO update\n")
out.append("`ncu --set full --clock-control none --import-source on -k regex:<our kernels> -s 18 -c 6 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras`")
out.append(f"(`gpurun_out/prof_{tag}.ncu-rep` is not committed; table = `ncu -i ... --page raw --csv` through `tools/summarize_ncu.py raw`).\n")
out.append(raw)

# ---- scan / V-MPO / SAC kernels
rows = list(csv.reader(open(f"gpurun_out/raw_scans_{tag}.csv")))
hdr, units = rows[0], rows[1]
col = {n: hdr.index(n) for n in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
                                   "launch__block_size", "smsp__inst_executed.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
tob = lambda v, u: float(v.replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}[u]
out.append("\n## Scan, V-MPO and SAC kernels (`SCAN_BENCH_ITERS=1 ncu --set full -k regex:'gae_kernel|td_gae|vtrace|vmpo_|sac_|...' python tools/scan_bench.py`)\n")
out.append("First launches of each kernel in `tools/scan_bench.py`: the scans at (4096, 128) then at B*T = 2^26, one V-MPO update at 1024 and 8192 sequences, one SAC-Continuous update (obs 17, act 6, 8192 x 5).\n")
out.append("| kernel | grid x block | duration [us] | DRAM read + write [MB] | DRAM GB/s | DRAM % of peak (ncu) | warp instructions |")
out.append("|---|---|---|---|---|---|---|")
per = {}
for r in rows[2:]:
    n = r[col["Kernel Name"]].split("(")[0].replace("b200rl::", "").replace("void ", "").strip()
    t = tob(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
    by = tob(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + tob(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    per.setdefault(n, []).append((t, by, r))
for n, lst in per.items():
    picks = [("first launch", lst[0])]
    big = max(lst, key=lambda x: x[0])
    if big is not lst[0] and big[0] > 2 * lst[0][0]:
        picks.append(("largest shape", big))
    for label, (t, by, r) in picks:
        out.append(f"| `{n}` ({label}) | {r[col['launch__grid_size']]} x {r[col['launch__block_size']]} | {t * 1e6:.1f} | {by / 1e6:.2f} | {by / t / 1e9:.0f} | "
                   f"{float(r[col['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']]):.1f} | {float(r[col['smsp__inst_executed.sum']]):.0f} |")
out.append("\nCUDA-event numbers of the same script without the profiler (`profiles/%s_scan_bench.jsonl`):\n" % tag)
for l in open(f"profiles/{tag}_scan_bench.jsonl"):
    d = json.loads(l)
    if "kernel" in d:
        out.append(f"* `{d['kernel']}` B={d['B']} T={d['T']}: {d['ms'] * 1e3:.1f} us, {d['achieved_gbs']:.0f} GB/s algorithmic = {d['frac_of_hbm_peak']:.2f} of the measured HBM peak ({d['note']})")
    else:
        out.append(f"* one {d['update']} update B={d['B']} T={d['T']}: {d['ms']:.3f} ms, {d['env_steps_per_s']:.3e} env-steps/s")
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")

rows = list(csv.reader(open(f"gpurun_out/raw_{tag}.csv")))
hdr, units = rows[0], rows[1]
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
tb = lambda v, u: float(v.replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u]
seen = {}
for r in rows[2:]:
    n = r[ki].split("(")[0].replace("b200rl::", "").strip()
    if n not in seen:
        seen[n] = tb(r[ri], units[ri]) + tb(r[wi], units[wi])
tr = {k: sum(seen[x] for x in v) for k, v in stage.items() if v and all(x in seen for x in v)}
tr["_source"] = f"profiles/{tag}_ncu_summary.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, summed over the kernels of a stage)"
json.dump(tr, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps({k: b[k] for k in ("value", "ms_per_step")}), b["e2e"]["value"], tr)
