#!/usr/bin/env python
"""profiles/traffic.json[workload][kernel family] = dram__bytes_read.sum + dram__bytes_write.sum per launch, from an
ncu --set full capture of one step (scripts/gpu_capture.sh).  bench.py reports it as roofline.traffic.

    python scripts/traffic_from_ncu.py gpurun_out/r01v8_full_C2.ncu-rep C2 2 [out.json]
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, wl, H = sys.argv[1], sys.argv[2], int(sys.argv[3])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki, ri, wi, ti = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                              "gpu__time_duration.sum"))
    # a capture may hold several steps: keep the last complete one (seed_kernel opens a forward pass,
    # finalize_loss_kernel closes the backward pass)
    body = rows[2:]
    starts = [i for i, r in enumerate(body) if "seed_kernel" in r[ki]]
    ends = [i for i, r in enumerate(body) if "finalize_loss_kernel" in r[ki]]
    if starts and ends:
        e = ends[-1]
        b = max(i for i in starts if i < e)
        body = body[b:e + 1]
    fam = {}
    n_fwd_in = n_bwd_in = 0
    for r in body:
        name = r[ki]
        m = re.match(r"(?:void )?(\w+)<([^>]*)>", name)
        base, targs = (m.group(1), [t.strip() for t in m.group(2).split(",")]) if m else (name, [])
        flag = targs[1] if len(targs) > 1 else ""
        base = base.replace("_tc_kernel", "_kernel")          # tcgen05 versions share the family of the mma.sync ones
        if base == "agg_fwd_kernel":
            key = "agg_fwd_0" if flag in ("1", "true") else f"agg_fwd_{(n_fwd_in := n_fwd_in + 1)}"
        elif base == "agg_bwd_kernel":
            if flag in ("1", "true"):
                key = "agg_bwd_0"
            else:
                key = f"agg_bwd_{H - 1 - n_bwd_in}"
                n_bwd_in += 1
        elif base == "leaf_entity_kernel":
            key = "leaf_entity_bwd" if flag in ("1", "true") else "leaf_entity_fwd"
        else:
            key = base.replace("_kernel", "")
        if key in fam:
            continue                      # first launch of each family (one step)
        byts = float(r[ri]) * UNIT[units[ri]] + float(r[wi]) * UNIT[units[wi]]
        fam[key] = {"dram_bytes": byts, "ncu_us": float(r[ti])}
    path = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[wl] = {k: v["dram_bytes"] for k, v in fam.items()}
    data.setdefault("_ncu_us", {})[wl] = {k: v["ncu_us"] for k, v in fam.items()}
    data["_source"] = "ncu --set full --clock-control none, one launch per kernel family (cold caches: ncu flushes between replays)"
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data[wl], indent=1))


if __name__ == "__main__":
    main()
