#!/usr/bin/env python
"""Warp-stall samples of one kernel of an .ncu-rep aggregated per CUDA source line.

    python scripts/ncu_lines.py gpurun_out/x.ncu-rep 'agg_bwd_kernel<(int)32, (bool)1>' [top_n]

ncu's CSV source page is SASS-only; the SASS offset -> (file, line) map comes from `nvdisasm -g` on the cubin of the
in-tree libmvin_b200.so (the same binary that ran on the GPU box).
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mvin_b200", "lib", "libmvin_b200.so")


def line_map():
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    maps, cur, loc = {}, None, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur = maps.setdefault(m.group(1), {})
            loc = None
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            loc = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and cur is not None:
            cur[int(m.group(1), 16)] = loc
    return maps


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    blk = next(b for b in blocks if pat in b["name"])
    maps = line_map()
    dm = demangle(list(maps))
    key = lambda s: re.sub(r"\(int\)|\(bool\)|mvin::|void |\s", "", s).replace("true", "1").replace("false", "0")
    want = key(blk["name"])
    sym = next(s for s, d in dm.items() if key(d) == want)
    lm = maps[sym]
    h = blk["hdr"]
    si = h.index("# Samples")
    stall = [(i, c[6:]) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    base = int(blk["rows"][0][0], 16)
    per = {}
    tot = 0
    for r in blk["rows"]:
        n = int(r[si] or 0)
        tot += n
        loc = lm.get(int(r[0], 16) - base)
        d = per.setdefault(loc, {"n": 0})
        d["n"] += n
        for i, c in stall:
            d[c] = d.get(c, 0) + int(r[i] or 0)
    src = {}
    print(f"{blk['name']}: {tot} samples")
    for loc, d in sorted(per.items(), key=lambda kv: -kv[1]["n"])[:top_n]:
        text = ""
        if loc:
            f = os.path.join(ROOT, "mvin_b200", "csrc", loc[0])
            if os.path.exists(f):
                src.setdefault(f, open(f).read().splitlines())
                text = src[f][loc[1] - 1].strip()[:90]
        why = sorted(((v, k) for k, v in d.items() if k != "n"), reverse=True)[:2]
        print(f"{100.0 * d['n'] / max(tot, 1):5.1f}%  {str(loc):24s} {'/'.join(k for v, k in why if v):22s} {text}")


if __name__ == "__main__":
    main()
