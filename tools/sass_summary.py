#!/usr/bin/env python
"""SASS evidence: per kernel, the counts of the instructions that show how it moves data and where it computes.

  python tools/sass_summary.py > profiles/r02_sass_summary.txt

Disassembles (cuobjdump -sass, no GPU needed) libspenso_b200.so and every cubin in spenso_b200/_jit_cache (the fused
network kernels and the specialised pairwise kernels build() pre-compiles) and counts, per kernel:
  LDG.E.*256 / LDG.E.*128 / LDG.E.64   global loads by width          STG.E.*          global stores
  LDGSTS                               cp.async (global -> shared)   UBLKCP / UTMALDG  TMA bulk / tensor copies
  LDS / STS                            shared-memory accesses        SHFL              warp shuffles
  DFMA / DMUL / DADD                   FP64 pipe                     DMMA              FP64 tensor pipe
  UTCMMA / LDTM                        tcgen05 (none expected: no f64 kind exists)
"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATTERNS = [("LDG.256", r"\bLDG\.E\.[A-Z0-9.]*256"), ("LDG.128", r"\bLDG\.E\.[A-Z0-9.]*128"), ("LDG.64", r"\bLDG\.E\.[A-Z0-9.]*64\b"),
            ("LDG.other", r"\bLDG\.E(?![A-Z0-9.]*(256|128|64))"), ("STG", r"\bSTG\.E"), ("LDGSTS", r"\bLDGSTS"),
            ("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"), ("LDS", r"\bLDS\b|\bLDS\."), ("STS", r"\bSTS\b|\bSTS\."),
            ("SHFL", r"\bSHFL"), ("DFMA", r"\bDFMA"), ("DMUL", r"\bDMUL"), ("DADD", r"\bDADD"), ("DMMA", r"\bDMMA"),
            ("UTCMMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("BAR", r"\bBAR\."), ("SYNCS", r"\bSYNCS")]


def kernels_of(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur and "/*" in line:
            body[cur].append(line)
    return body


def demangle(name):
    try:
        full = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        full = full.replace("(anonymous namespace)::", "")
        m = re.match(r"(?:void\s+)?([\w:]+)(<[^(]*>)?\(", full)      # kernel name + template arguments, no parameter list
        return (m.group(1) + (m.group(2) or "")) if m else full
    except Exception:
        return name


def main():
    files = [os.path.join(ROOT, "spenso_b200", "libspenso_b200.so")] + sorted(glob.glob(os.path.join(ROOT, "spenso_b200", "_jit_cache", "*.cubin")))
    print("# SASS instruction counts per kernel (cuobjdump -sass, sm_100a); columns with a zero total are omitted per row")
    for f in files:
        ks = kernels_of(f)
        for name, lines in sorted(ks.items()):
            text = "\n".join(lines)
            counts = [(lab, len(re.findall(pat, text))) for lab, pat in PATTERNS]
            shown = " ".join(f"{lab}={n}" for lab, n in counts if n)
            short = demangle(name)
            short = re.sub(r"spn::\(anonymous namespace\)::", "", short)
            print(f"{os.path.basename(f)[:24]:24s} {short[:70]:70s} instr={len(lines):5d}  {shown}")


if __name__ == "__main__":
    main()
