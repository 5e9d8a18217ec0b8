"""profiles/<tag>_ncu_summary.md + profiles/traffic.json from one tools/gpu_prof.sh run (dev tool).
   python tools/make_profile_summary.py r1k "<free-text reading, markdown>"   (reads gpurun_out/*_<tag>.*)"""
import csv
import json
import shutil
import subprocess
import sys

tag = sys.argv[1]
reading = sys.argv[2] if len(sys.argv) > 2 else ""
shutil.copy(f"gpurun_out/bench_{tag}.json", f"profiles/{tag}_bench.json")
shutil.copy(f"gpurun_out/launches_{tag}.csv", f"profiles/{tag}_launches_bench.csv")
shutil.copy(f"gpurun_out/configs_{tag}.jsonl", f"profiles/{tag}_configs.jsonl")
b = json.loads(open(f"profiles/{tag}_bench.json").read().strip().splitlines()[-1])
km = b["kernels_ms"]
tot = sum(km.values())
run = lambda *a: subprocess.run(["python", "tools/summarize_ncu.py", *a], capture_output=True, text=True).stdout
launch = run("launches", f"profiles/{tag}_launches_bench.csv")
raw = run("raw", f"gpurun_out/raw_{tag}.csv")
rows = [l.split("|") for l in launch.splitlines() if l.startswith("| ") and not l.startswith("| kernel") and "**sum**" not in l]
ms = {r[1].strip(): float(r[3]) for r in rows}
pick = lambda sub: [k for k in ms if sub in k]
stage = {"lstm_fwd": pick("xproj") + pick("lstm_fwdt"), "head_fwd": pick("head_fwd"), "ppo_head": pick("ppo_head"), "lstm_bwd": pick("lstm_bwdt"),
         "lstm_wgrad": pick("wgrad_tc") + pick("torso_grad") + pick("head_grad"), "reduce_partials": pick("reduce_partials"), "clip_rmsprop": pick("clip_rmsprop")}
ntot = sum(ms.values())
out = [f"# {tag} -- PPO B=4096 T=128 obs=4 A=2, one B200\n", "## Launch list of the bench command\n",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 66 --csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline`",
       f"(raw: `profiles/{tag}_launches_bench.csv`; per-launch times under ncu are cold-cache and serialised -- compare SHARES).\n", launch,
       f"bench.py's own CUDA-event times of the same stages, each timed alone (`profiles/{tag}_bench.json` `kernels_ms`):\n",
       "| stage | bench ms | bench share | ncu share |", "|---|---|---|---|"]
for k, v in km.items():
    n = sum(ms[x] for x in stage[k])
    out.append(f"| {k} (= {' + '.join(x.replace('void ', '') for x in stage[k])}) | {v:.3f} | {100 * v / tot:.1f} % | {100 * n / ntot:.1f} % |")
cb = b.get("cpu_baseline") or {}
out.append(f"\nbench: value {b['value']:.3e} env-steps/s, {b['ms_per_step']:.3f} ms/step; e2e (host batches through `learning()`) {b['e2e']['value']:.3e}; "
           + (f"CPU port {cb['value']:.3e} on {cb['cores']} cores; " if cb else "") + f"clocks {b['clocks']['sm_mhz']} MHz, throttle reasons {b['clocks']['reasons']}.\n")
out.append("## `ncu --set full` of the main kernels\n")
out.append("`ncu --set full --clock-control none --import-source on -k regex:'lstm_|torso_grad|head_grad|head_fwd' -s 21 -c 14 python tools/kernel_times.py 4096 128 4 2 1`")
out.append(f"(`gpurun_out/prof_{tag}.ncu-rep`, not committed: 27 MB; table = `ncu -i ... --page raw --csv` through `tools/summarize_ncu.py raw`).\n")
out.append(raw)
if reading:
    out.append("\n" + reading + "\n")
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out))
rows = list(csv.reader(open(f"gpurun_out/raw_{tag}.csv")))
hdr, units = rows[0], rows[1]
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
tob = lambda v, u: float(v.replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u]
seen = {}
for r in rows[2:]:
    n = r[ki].split("(")[0].replace("b200rl::", "").strip()
    if n not in seen:
        seen[n] = tob(r[ri], units[ri]) + tob(r[wi], units[wi])
tr = {k: sum(seen[x] for x in v) for k, v in stage.items() if v and all(x in seen for x in v)}
tr["_source"] = f"profiles/{tag}_ncu_summary.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, summed over the kernels of a stage)"
json.dump(tr, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps({k: b[k] for k in ("value", "ms_per_step")}), b["e2e"]["value"], tr)
