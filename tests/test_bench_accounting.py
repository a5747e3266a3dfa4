"""bench.py's roofline accounting against SURVEY.md 8(d): algorithmic bytes per user-item pair of every BASELINE.json
configuration, and the split of that figure over the kernel families."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# (fwd bytes per pair, fwd + bwd bytes per pair)
SURVEY_TABLE = {"C1": (18_648, 52_696), "C2": (71_576, 207_256), "C3": (346_392, 1_019_160),
                "C4": (8_938_456, 26_273_752), "C5": (2_296_856, 6_820_888)}


def test_bytes_per_pair_match_the_survey_table():
    b = _bench()
    for key, (fwd, total) in SURVEY_TABLE.items():
        f, w = b.bytes_per_pair(b.WORKLOADS[key])
        assert (f, f + w) == (fwd, total), key


def test_kernel_split_covers_the_row_bytes_of_the_per_pair_figure():
    """Per-launch algorithmic bytes of each kernel family = rows the launch covers x 8(d)'s per-row figure (4d-byte row
    + ids forward, row re-read + gradient row backward).  An inner level h is a child level of aggregator iterations
    0 .. L - h, so over a step it is counted L - h + 1 times (activation buffers of later iterations), while the leaf
    level -- the bulk of 8(d)'s per-pair figure -- is counted once, in iteration 0."""
    b = _bench()
    for key in SURVEY_TABLE:
        w = b.WORKLOADS[key]
        d, L, K, B, p, m = w["dim"], w["h_hop"], w["K"], w["B"], w["p"], w["m"]
        per = b.kernel_bytes_per_step(w)
        assert set(per) == {"transform_fwd", "transform_bwd", "user_fwd", "ripple_bwd"} | \
            {f"agg_{d_}_{i}" for d_ in ("fwd", "bwd") for i in range(L)}
        rows = [K ** h for h in range(L + 1)]
        # iteration i covers child levels 1 .. L - i, so level h is read by iterations 0 .. L - h
        child = sum(rows[h] * (L - h + 1) for h in range(1, L + 1))
        want_fwd = sum(rows[h] * (4 * d + 4) for h in range(L)) + child * (4 * d + 8) + \
            (2 * p * m + 1) * 4 * d + 12 * p * m + 8
        want_bwd = sum(rows[h] * (8 * d + 4) for h in range(L)) + child * (8 * d + 8) + 4 * p * m * 4 * d + 12 * p * m
        got_fwd = (per["transform_fwd"] + per["user_fwd"] + sum(per[f"agg_fwd_{i}"] for i in range(L))) / B
        got_bwd = (per["transform_bwd"] + per["ripple_bwd"] + sum(per[f"agg_bwd_{i}"] for i in range(L))) / B
        assert got_fwd == want_fwd and got_bwd == want_bwd, key
        # the leaf gather (iteration 0, level L) carries the bulk of 8(d)'s figure at every configuration with L >= 2
        if L >= 2:
            assert per["agg_fwd_0"] / B >= 0.5 * SURVEY_TABLE[key][0]


def test_workloads_are_the_baseline_configs():
    b = _bench()
    shape = lambda k: tuple(b.WORKLOADS[k][f] for f in ("dim", "h_hop", "K", "B", "p", "m"))
    assert shape("C1") == (16, 1, 8, 1024, 2, 64)
    assert shape("C2") == (32, 2, 16, 4096, 2, 64)
    assert shape("C3") == (64, 2, 32, 8192, 2, 64)
    assert shape("C4") == (64, 3, 32, 16384, 1, 16)
    assert shape("C5") == (128, 2, 64, 8192, 2, 64) and b.WORKLOADS["C5"]["n_entity_per_gpu"] * 8 == 100_000_000
    assert b.METRIC == "user-item pairs/sec fwd+bwd" and b.UNIT == "pairs/s"
