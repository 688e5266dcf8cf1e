#!/usr/bin/env python
"""scripts/sass_summary.py -- what the built library's SASS says about each kernel of the path (no GPU needed).

    python scripts/sass_summary.py > profiles/r2_sass_summary.md        # table: instruction count + the mnemonics that matter
    python scripts/sass_summary.py --dump 'translate_u16_rows_kernel<false>' > profiles/r2_sass_translate_rows.txt

Runs `cuobjdump -sass` on librir_b200/libs/libsignal_processing_b200.so, demangles the kernel names and counts, per kernel,
the instructions that prove what the source claims: UTMALDG (TMA box loads), SYNCS (mbarrier), DFMA / DADD (the exact fp64
blend), FFMA2 / FADD2 (packed fp32), LDG / STG widths, ATOMS (shared-memory histogram), BAR (CTA-wide barriers).
`--dump` writes one kernel's listing without the encodings.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "librir_b200", "libs", "libsignal_processing_b200.so")
KEYS = ["UTMALDG", "UTMAPF", "SYNCS", "BAR.SYNC", "DFMA", "DADD", "DMUL", "FFMA2", "FADD2", "FMUL2", "FFMA", "IMAD.WIDE", "LDS.128", "LDS.64", "LDG 256", "LDG 128",
        "STG 256", "STG 128", "ATOMS", "ATOMG", "RED", "I2F", "F2I", "F2F", "LDGSTS", "STL", "LDL"]


def kernels():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            names[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", line)
        if m and cur:
            names[cur].append((m.group(1), m.group(2)))
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    out = {}
    for mangled, d in zip(names, dem):
        d = re.sub(r"^void ", "", d)
        d = re.sub(r"\(.*$", "", d).replace("rirb::", "").replace("unsigned short", "u16")
        out[d] = names[mangled]
    return out


def main():
    ks = kernels()
    if len(sys.argv) > 2 and sys.argv[1] == "--dump":
        ins = ks[sys.argv[2]]
        print(f"# {sys.argv[2]}: {len(ins)} SASS instructions (cuobjdump -sass, sm_100a, encodings stripped)")
        for addr, text in ins:
            print(f"/*{addr}*/  {text}")
        return
    print("# SASS of the kernels in libsignal_processing_b200.so (round 2, `scripts/sass_summary.py`, nvcc 12.9, sm_100a)\n")
    print("Counts are static instructions in the kernel image (not executed counts). Columns that are zero for every kernel are dropped.\n")
    rows = {}
    for name, ins in sorted(ks.items()):
        c = collections.Counter()
        for _, t in ins:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            op = t.split()[0] if t else ""
            if op.startswith(("LDG.", "STG.")) and (".256" in op or ".128" in op):  # width whatever the cache / scope modifiers
                c[op[:3] + (" 256" if ".256" in op else " 128")] += 1
                continue
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    c[k] += 1
                    break
        rows[name] = (len(ins), c)
    used = [k for k in KEYS if any(r[1][k] for r in rows.values())]
    print("| kernel | instr | " + " | ".join(used) + " |")
    print("|---|---|" + "---|" * len(used))
    for name, (n, c) in rows.items():
        print(f"| `{name}` | {n} | " + " | ".join(str(c[k]) if c[k] else "" for k in used) + " |")


if __name__ == "__main__":
    main()
