#!/usr/bin/env python
"""Per-kernel SASS evidence of the product library: counts of the tcgen05 / TMEM / mbarrier / bulk-copy mnemonics in every
kernel of lib/libb200rl.so (`cuobjdump -sass`), written to profiles/sass_summary.md.  Runs in the GPU-less build container.

    python tools/sass_summary.py [tag]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200", "lib", "libb200rl.so")
MNEMONICS = [("UTCHMMA", "tcgen05.mma (kind::tf32 / kind::f16)"), ("LDTM", "tcgen05.ld"), ("STTM", "tcgen05.st"),
             ("UTCBAR", "tcgen05.commit"), ("SYNCS", "mbarrier"), ("UTMALDG", "TMA load"), ("UTMASTG", "TMA store"),
             ("UBLKCP", "bulk copy"), ("LDGSTS", "cp.async"), ("USETMAXREG", "setmaxnreg"), ("MUFU", "MUFU"),
             ("FFMA", "FFMA"), ("LDG", "LDG"), ("STG", "STG"), ("LDS", "LDS"), ("STS", "STS"), ("STL", "local store (spill)"),
             ("LDL", "local load (spill)")]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("b200rl::", "").replace("void ", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["_n"] += 1
            for key, _ in MNEMONICS:
                if op.split(".")[0] == key or (key in ("LDG", "STG", "LDS", "STS", "LDL", "STL") and op.split(".")[0] == key):
                    cur[key] += 1
    cols = [k for k, _ in MNEMONICS]
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.md")
    with open(path, "w") as f:
        f.write(f"# SASS summary of lib/libb200rl.so ({tag})\n\n`python tools/sass_summary.py {tag}` = `cuobjdump -sass` of the product library, "
                "instruction counts per kernel (static counts, not executed instructions).\n\n")
        f.write("Legend: " + "; ".join(f"`{k}` = {d}" for k, d in MNEMONICS) + ".\n\n")
        f.write("| kernel | instr | " + " | ".join(cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
        tot = collections.Counter()
        for name, c in kernels.items():
            f.write(f"| `{name}` | {c['_n']} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |\n")
            tot.update(c)
        f.write(f"| **total** | {tot['_n']} | " + " | ".join(str(tot[k]) for k in cols) + " |\n")
        tck = [n for n, c in kernels.items() if c["UTCHMMA"]]
        f.write("\nKernels that issue tcgen05.mma: " + ", ".join(f"`{n}`" for n in tck) + ".\n")
        f.write("\nNo TMA (`UTMALDG` / `UTMASTG` / `UBLKCP`) anywhere: the streaming operands are staged by per-thread 16-byte loads and "
                "cp.async; see DESIGN.md for why (K-major tf32 operands need a transposition TMA cannot do without an MN-major descriptor).\n"
                if not (tot["UTMALDG"] or tot["UTMASTG"] or tot["UBLKCP"]) else "")
    print("wrote", path)


if __name__ == "__main__":
    main()
