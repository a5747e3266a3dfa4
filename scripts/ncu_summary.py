#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full --import-source on) into a small text file for profiles/.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt

Per captured launch: duration, DRAM bytes, L2 / L1 / SM throughput, occupancy, registers, then the warp-stall
breakdown and the hottest SASS instructions from the source page (B200_PROFILING.md, "Here, with no GPU").
"""
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
       "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
       "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic"]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# {rep}: ncu --set full --clock-control none --import-source on; per launch\n")
    for n, r in enumerate(rows[2:]):
        print(f"[{n}] {r[ki]}")
        for m in RAW:
            if m in hdr:
                i = hdr.index(m)
                print(f"    {m:70s} {r[i]:>14s} {units[i]}")
    # source page: one block per launch
    src = ncu(rep, "source")
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    print("\n# warp-stall sampling (source page): share of all samples by reason, then the hottest SASS instructions\n")
    for n, b in enumerate(blocks):
        h = b["hdr"]
        si = h.index("# Samples")
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        tot = sum(int(r[si] or 0) for r in b["rows"])
        print(f"[{n}] {b['name']}  samples={tot}")
        if tot == 0:
            continue
        agg = sorted(((sum(int(r[i] or 0) for r in b["rows"]), c) for i, c in stall_cols), reverse=True)
        print("    " + ", ".join(f"{c[6:]} {100.0 * v / tot:.1f}%" for v, c in agg[:7] if v))
        hot = sorted(b["rows"], key=lambda r: -int(r[si] or 0))[:top_n]
        for r in hot:
            why = max(stall_cols, key=lambda ic: int(r[ic[0]] or 0))[1][6:]
            print(f"    {100.0 * int(r[si]) / tot:5.1f}%  {why:14s} {r[1].strip()[:100]}")


if __name__ == "__main__":
    main()
