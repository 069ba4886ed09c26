#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i`): headline metrics, aggregate stall reasons, the SASS lines with
the most stall samples, and the execution counts of the barrier waits / MMA / TMA instructions.
    python tools/ncu_report.py gpurun_out/x.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = page(rep, "raw")
    hdr, val = raw[0], raw[2]
    want = ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "launch__registers_per_thread",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__grid_size",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed")
    for i, h in enumerate(hdr):
        if h in want:
            print(f"{h} = {val[i]}")
    src = page(rep, "source")
    h2 = src[1]
    si, ci, ie = h2.index("# Samples"), h2.index("Source"), h2.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for k, r in enumerate(src[2:]):
        try:
            data.append((int(r[si]), k, r))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data) or 1
    agg = {}
    for n, k, r in data:
        for i in stall_cols:
            agg[h2[i]] = agg.get(h2[i], 0) + int(r[i])
    print("total samples", tot, sorted(agg.items(), key=lambda x: -x[1])[:8])
    for n, k, r in sorted(data, key=lambda x: -x[0])[:top]:
        st = sorted([(int(r[i]), h2[i]) for i in stall_cols], reverse=True)[:2]
        print(f"{k:5d} {n:7d} {100 * n / tot:5.1f}% exec={r[ie]:>9} {r[ci].strip()[:72]:72s} {st}")
    print("--- sync / async-unit instructions")
    for n, k, r in data:
        s = r[ci]
        if any(t in s for t in ("TRYWAIT", "UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "STTM", "SYNCS.ARRIVE", "BAR.SYNC", "MEMBAR", "FENCE")):
            print(f"{k:5d} {n:6d} {r[ie]:>9} {s.strip()[:100]}")


if __name__ == "__main__":
    main()
