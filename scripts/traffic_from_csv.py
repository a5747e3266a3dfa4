#!/usr/bin/env python
"""profiles/traffic.json[workload][kernel family] = dram__bytes_read.sum + dram__bytes_write.sum per launch, from the
CSV log of   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv
over a few bench steps (scripts/gpu_round2.sh).  The last complete step of the log is used (seed_kernel opens a forward
pass, finalize_loss_kernel closes the backward pass); a family launched several times in the step is summed.  bench.py
reports the entry of its dominant kernel as roofline.traffic.

    python scripts/traffic_from_csv.py gpurun_out/x_launches_C4.csv C4 [out.json]
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3,
        "second": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def family(name, state, H):
    m = re.match(r"(?:void )?(?:mvin::)?(\w+)<([^>]*)>", name)
    base, targs = (m.group(1), [t.strip() for t in m.group(2).split(",")]) if m else (re.sub(r"\(.*", "", name), [])
    base = base.replace("mvin::", "").replace("_tc_kernel", "_kernel")
    flag = targs[1] if len(targs) > 1 else ""
    leaf = flag in ("1", "true", "(bool)1")
    if base == "agg_fwd_kernel":
        if leaf:
            return "agg_fwd_0"
        state["f"] += 1
        return f"agg_fwd_{state['f']}"
    if base == "agg_bwd_kernel":
        if leaf:
            return "agg_bwd_0"
        state["b"] += 1
        return f"agg_bwd_{H - state['b']}"
    if base in ("leaf_entity_kernel", "leaf_entity_reg_kernel"):
        return "leaf_entity_bwd" if leaf else "leaf_entity_fwd"
    if base == "virt_group_kernel":
        return "group_bwd" if leaf else "group_fwd"
    if base in ("relq_kernel", "reldv_kernel", "reldrk_kernel"):          # "_tc" was stripped above
        return {"relq_kernel": "gemm_q", "reldv_kernel": "gemm_dv", "reldrk_kernel": "gemm_drk"}[base]
    if base == "virt_rows_kernel":
        return "virt_rows_bwd" if leaf else "virt_rows_fwd"
    if base == "agg_bwd_leaf_kernel":
        return "agg_bwd_leaf_tc"
    if base == "gemm_kernel":
        state["g"] += 1
        return f"gemm#{state['g']}"
    return base.replace("_kernel", "")


def main():
    path, wl = sys.argv[1], sys.argv[2]
    lines = [l for l in open(path, newline="") if not l.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    idx = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches = {}
    order = []
    for r in rows[1:]:
        if len(r) <= idx["Metric Value"]:
            continue
        lid = r[idx["ID"]]
        if lid not in launches:
            launches[lid] = {"name": r[idx["Kernel Name"]]}
            order.append(lid)
        val = float(r[idx["Metric Value"]].replace(",", "")) * UNIT.get(r[idx["Metric Unit"]], 1.0)
        launches[lid][r[idx["Metric Name"]]] = val
    seq = [launches[i] for i in order]
    starts = [i for i, l in enumerate(seq) if "seed_kernel" in l["name"]]
    ends = [i for i, l in enumerate(seq) if "finalize_loss_kernel" in l["name"]]
    if not starts or not ends:
        raise SystemExit("no complete step in the log")
    e = ends[-1]
    b = max(i for i in starts if i < e)
    step = seq[b:e + 1]
    # aggregator iterations of the step: one agg_bwd launch each, except that the entity-table mode (table.cuh) has no
    # launch for iteration 0
    H = sum(1 for l in step if "agg_bwd_kernel" in l["name"]) + (1 if any("table_bwd_kernel" in l["name"] for l in step) else 0)
    state = {"f": 0, "b": 0, "g": 0}
    fam, us = {}, {}
    for l in step:
        k = family(l["name"], state, H)
        fam[k] = fam.get(k, 0.0) + l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
        us[k] = us.get(k, 0.0) + l.get("gpu__time_duration.sum", 0.0)
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "traffic.json")
    data = json.load(open(out)) if os.path.exists(out) else {}
    data[wl] = fam
    data.setdefault("_ncu_us", {})[wl] = us
    data["_source"] = ("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none: "
                       "bytes per launch of every kernel of the last complete step of the log (ncu serialises kernels and "
                       "flushes caches between them, so these are cold-cache figures)")
    json.dump(data, open(out, "w"), indent=1, sort_keys=True)
    tot_us = sum(us.values())
    for k in sorted(fam, key=lambda k: -us[k]):
        print(f"{k:20s} {us[k]:10.1f} us {100 * us[k] / tot_us:5.1f} %  {fam[k] / 1e6:10.1f} MB  {fam[k] / max(us[k], 1e-9) / 1e3:8.1f} GB/s")


if __name__ == "__main__":
    main()
