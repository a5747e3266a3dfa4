"""Extract a known-answer fixture from the reference's own case-study logs (case_st/amazon-book_20core/*.log).

Those logs were written by the REAL TensorFlow reference (util.py:59-127 dumps what MVIN.eval_case_study returned with
trained weights): for every logged user-item pair the sampled relation ids of the two KG hops and the attention
weights `probs_normalized` of aggregator 0 at hop 0 (importance_list_0, [K]) and hop 1 (importance_list_1, [K, K])
(model.py:294,304,319-323; aggregators.py:139-146).  They are the only numerical outputs of the reference that ship
with it, so they pin the attention path of the oracle against TF itself (tests/test_case_study_logs.py).

    python tests/golden/make_case_study_fixture.py        # needs /root/reference; writes tests/golden/case_study_att.npz
"""
import glob
import os
import re

import numpy as np

LOG_DIR = "/root/reference/case_st/amazon-book_20core"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "case_study_att.npz")
N_PER_FILE = 120
K = 8


def parse(path, limit):
    rel0, att0, rel1, att1 = [], [], [], []
    with open(path, errors="replace") as f:
        lines = f.read().split("\n")
    i, n = 0, len(lines)
    while i < n and len(rel0) < limit:
        if not lines[i].startswith("user_indices = "):
            i += 1
            continue
        # integer block
        i += 1
        r0, r1 = None, []
        while i < n and "entity_relation_name" not in lines[i]:
            ln = lines[i]
            if ln.startswith("rela_index 0 = "):
                ids = [int(x) for x in ln.split("=", 1)[1].split(",")]
                if r0 is None:
                    r0 = ids
                else:
                    r1.append(ids)
            i += 1
        # name / attention block.  Two log formats: "... att = x" per neighbour (later runs), or an "importance = ..." line
        # per node (numpy print, possibly wrapped, at hop 0; comma separated at hop 1)
        a0, a1, cur, layer = [], [], [], 0
        i0, i1 = [], []
        while i < n and not lines[i].startswith("user_indices = ") and not lines[i].startswith("*" * 50):
            ln = lines[i]
            if "second_layer" in ln:
                layer = 1
            m = re.search(r"att = ([0-9.eE+-]+)\s*$", ln)
            if m:
                if layer == 0:
                    a0.append(float(m.group(1)))
                else:
                    cur.append(float(m.group(1)))
                    if len(cur) == K:
                        a1.append(cur)
                        cur = []
            elif ln.startswith("importance = "):
                txt = ln.split("=", 1)[1]
                while "[" in txt and "]" not in txt and i + 1 < n:
                    i += 1
                    txt += " " + lines[i]
                vals = [float(x) for x in re.split(r"[,\s\[\]]+", txt) if x]
                if layer == 0:
                    i0 = vals
                else:
                    i1.append(vals)
            i += 1
        if len(a0) != K:
            a0 = i0
        if len(a1) != K:
            a1 = i1
        if r0 is not None and len(r0) == K and len(r1) == K and len(a0) == K and len(a1) == K and all(len(x) == K for x in a1):
            rel0.append(r0); att0.append(a0); rel1.append(r1); att1.append(a1)
    return (np.asarray(rel0, np.int16), np.asarray(att0, np.float32), np.asarray(rel1, np.int16),
            np.asarray(att1, np.float32))


def main():
    out = {}
    names = []
    for path in sorted(glob.glob(os.path.join(LOG_DIR, "*.log"))):
        key = os.path.basename(path)[len("amazon-book_20core_p01_h2_n_mix1_nb_8_"):-len(".log")]
        r0, a0, r1, a1 = parse(path, N_PER_FILE)
        print(key, r0.shape, a1.shape)
        names.append(key)
        out[key + "__rel0"], out[key + "__att0"], out[key + "__rel1"], out[key + "__att1"] = r0, a0, r1, a1
    out["names"] = np.asarray(names)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
