"""Turn ncu CSV exports into the markdown tables kept under profiles/ (dev tool).

  python tools/summarize_ncu.py launches <launches.csv>          # --metrics gpu__time_duration.sum --csv log
  python tools/summarize_ncu.py raw <raw.csv>                    # ncu -i x.ncu-rep --page raw --csv
"""
import csv
import sys
from collections import OrderedDict

KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
]
STALLS = "smsp__average_warp_latency_issue_stalled_"  # ncu >= 2024; older: smsp__average_warps_issue_stalled_*


def short(name):
    n = name.split("(")[0]
    return n.replace("b200rl::", "").strip()


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        if r[ui] in ("ns", "nsecond"):
            v *= 1e-6
        elif r[ui] in ("us", "usecond"):
            v *= 1e-3
        agg.setdefault(short(r[ki]), []).append(v)
    tot = sum(sum(v) / len(v) for v in agg.values())
    print("| kernel | launches | avg ms (ncu) | share |\n|---|---|---|---|")
    for k, v in agg.items():
        a = sum(v) / len(v)
        print(f"| {k} | {len(v)} | {a:.3f} | {100 * a / tot:.1f} % |")
    print(f"| **sum** | | **{tot:.3f}** | |")


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    seen, body = set(), []
    for r in rows[2:]:                                # one column per kernel: the first captured launch
        if short(r[ki]) not in seen:
            seen.add(short(r[ki]))
            body.append(r)
    rows = rows[:2] + body
    kern = [short(r[ki]) for r in rows[2:]]
    print("| metric | " + " | ".join(kern) + " |\n|---|" + "---|" * len(kern))
    for m, label in KEEP:
        if m in hdr:
            i = hdr.index(m)
            print(f"| {label} [{units[i]}] | " + " | ".join(r[i] for r in rows[2:]) + " |")
    idx = [i for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith(".ratio") and "not_issued" not in h]
    print("\nTop warp-stall reasons (cycles per issued instruction):\n")
    for r, k in zip(rows[2:], kern):
        st = sorted(((float(r[i].replace(",", "")) if r[i] else 0.0, hdr[i].split("issue_stalled_")[1].replace(".ratio", "")) for i in idx), reverse=True)[:6]
        print(f"* {k}: " + ", ".join(f"{n} {v:.2f}" for v, n in st))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
