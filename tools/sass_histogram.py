#!/usr/bin/env python
"""SASS evidence that the contraction kernels are Blackwell-native: counts of the tcgen05 / TMEM / TMA mnemonics in
libmavlm.so, per kernel family (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA ->
UTMALDG/UTMASTG, legacy mma.sync -> HMMA).     python tools/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "memory-augmented-vlm_b200", "libmavlm.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTMAPF|HMMA|HGMMA|SYNCS|MUFU\.EX2|RED|ATOMG)\b")
fam = None
per = collections.defaultdict(collections.Counter)
total = collections.Counter()
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        fam = ("gemm_tc_kernel" if "gemm_tc_kernel" in name else "attn_tc_kernel" if "attn_tc_kernel" in name else
               "attn_pair_kernel" if "attn_pair_kernel" in name else "other kernels")
        continue
    m = pat.search(line)
    if m and fam:
        per[fam][m.group(1)] += 1
        total[m.group(1)] += 1
n_fn = len(re.findall(r"Function : ", sass))
print(f"cuobjdump -sass memory-augmented-vlm_b200/libmavlm.so   ({n_fn} sm_100a functions)")
cols = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU.EX2", "HMMA", "HGMMA"]
print(f"{'kernel family':<20}" + "".join(f"{c:>10}" for c in cols))
for f in ("gemm_tc_kernel", "attn_tc_kernel", "attn_pair_kernel", "other kernels"):
    print(f"{f:<20}" + "".join(f"{per[f][c]:>10}" for c in cols))
print(f"{'TOTAL':<20}" + "".join(f"{total[c]:>10}" for c in cols))
print("\nUTCHMMA = tcgen05.mma kind::f16 (tmem[...] operand form = TS mode), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st,")
print("UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (TMA), UBLKCP = cp.async.bulk (1-D bulk copy: LayerNorm row staging),")
print("SYNCS = mbarrier ops.  HMMA / HGMMA (mma.sync / wgmma): none.")
