"""Known-answer test against the reference's OWN numerical output: the attention weights the real TensorFlow model
logged in case_st/amazon-book_20core/*.log (fixture tests/golden/case_study_att.npz, extracted by
tests/golden/make_case_study_fixture.py).  18 logs = 18 trained models (ablations `all` and `no_kg_eh_uo`, three epochs,
three runs); per log 120 user-item pairs with the relation ids of hop 0 ([K]) and hop 1 ([K, K]) and the logged
probs_normalized of aggregator 0 at both hops (model.py:294,304,319-323; aggregators.py:139-146).

What they pin:
  * softmax over the K axis: every logged attention vector sums to 1;
  * the logits are relation-only after the softmax: ONE table of per-relation scalars s[r] per model reproduces every
    logged vector of BOTH hops, whatever the user / the node -- fitted on half of the records, it predicts the other
    half to print precision;
  * the oracle's aggregator (oracle/mvin_oracle.py::sum_aggregator_urh, the [user; relation; self] concat order and
    the no-bias logit of aggregators.py:121-139) returns exactly those probabilities for arbitrary user / self vectors
    once Rel . w_r equals the fitted table.
"""
import os

import numpy as np
import pytest
import torch

from oracle import mvin_oracle as orc

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "case_study_att.npz")


def _load():
    z = np.load(FIX)
    for name in z["names"]:
        name = str(name)
        yield name, z[name + "__rel0"].astype(np.int64), z[name + "__att0"].astype(np.float64), \
            z[name + "__rel1"].astype(np.int64), z[name + "__att1"].astype(np.float64)


def _fit_scores(rel, att, n_rel):
    """least squares for s (up to a constant): log att_k - log att_j = s[r_k] - s[r_j] within each attention vector."""
    rows, rhs = [], []
    for r, a in zip(rel.reshape(-1, rel.shape[-1]), att.reshape(-1, att.shape[-1])):
        la = np.log(a)
        for k in range(1, len(r)):
            if r[k] != r[0]:
                row = np.zeros(n_rel)
                row[r[k]], row[r[0]] = 1.0, -1.0
                rows.append(row)
                rhs.append(la[k] - la[0])
    A, b = np.asarray(rows), np.asarray(rhs)
    s, *_ = np.linalg.lstsq(A, b, rcond=None)
    seen = np.abs(A).sum(0) > 0
    return s, seen


CASES = list(_load())


def test_fixture_is_complete():
    assert len(CASES) == 18
    for name, r0, a0, r1, a1 in CASES:
        assert r0.shape == (120, 8) and a0.shape == (120, 8) and r1.shape == (120, 8, 8) and a1.shape == (120, 8, 8), name


@pytest.mark.parametrize("name,r0,a0,r1,a1", CASES, ids=[c[0] for c in CASES])
def test_logged_attention_is_a_relation_only_softmax(name, r0, a0, r1, a1):
    assert np.abs(a0.sum(-1) - 1).max() < 2e-6 and np.abs(a1.sum(-1) - 1).max() < 2e-6
    n_rel = int(max(r0.max(), r1.max())) + 1
    half = 60
    rel_fit = np.concatenate([r0[:half].reshape(-1, 8), r1[:half].reshape(-1, 8)])
    att_fit = np.concatenate([a0[:half].reshape(-1, 8), a1[:half].reshape(-1, 8)])
    s, seen = _fit_scores(rel_fit, att_fit, n_rel)
    rel_chk = np.concatenate([r0[half:].reshape(-1, 8), r1[half:].reshape(-1, 8)])
    att_chk = np.concatenate([a0[half:].reshape(-1, 8), a1[half:].reshape(-1, 8)])
    ok = seen[rel_chk].all(-1)                      # vectors whose relations all occurred in the fitted half
    assert ok.mean() > 0.9
    logit = s[rel_chk[ok]]
    pred = np.exp(logit - logit.max(-1, keepdims=True))
    pred /= pred.sum(-1, keepdims=True)
    err = np.abs(pred - att_chk[ok]) / np.maximum(att_chk[ok], 1e-6)
    assert err.max() < 2e-4, (name, err.max())      # the logs print 8 significant digits of float32 values

    # the oracle's aggregator, with Rel . w_r = s and RANDOM user / self vectors and user / self weight thirds
    d, B = 4, 16
    g = torch.Generator().manual_seed(0)
    Rel = torch.zeros(n_rel, d, dtype=torch.float64)
    Rel[:, 0] = torch.from_numpy(s)
    w = torch.randn(3 * d, 1, generator=g, dtype=torch.float64)
    w[d:2 * d, 0] = torch.tensor([1.0, 0.3, -0.7, 2.0], dtype=torch.float64)   # only component 0 of Rel is non-zero
    P = {"agg_urh_weights": w, "agg_weights": torch.randn(d, d, generator=g, dtype=torch.float64),
         "agg_bias": torch.zeros(d, dtype=torch.float64)}
    cfg = orc.OracleConfig(dim=d, neighbor_sample_size=8, h_hop=1, n_mix_hop=1, p_hop=1, n_memory=4)
    idx = np.nonzero(ok)[0][:B]
    rel_b = torch.from_numpy(rel_chk[idx])                                        # [B, K]
    _, probs = orc.sum_aggregator_urh(P, "agg", cfg, torch.randn(B, 1, d, generator=g, dtype=torch.float64),
                                      torch.randn(B, 1, 8, d, generator=g, dtype=torch.float64),
                                      Rel[rel_b].reshape(B, 1, 8, d), torch.randn(B, d, generator=g, dtype=torch.float64))
    got = probs.reshape(B, 8).numpy()
    err = np.abs(got - att_chk[idx]) / np.maximum(att_chk[idx], 1e-6)
    assert err.max() < 2e-4, (name, err.max())


@pytest.mark.gpu
@pytest.mark.parametrize("case", [0, 5, 9, 14])
def test_cuda_importance_reproduces_logged_attention(case):
    """The CUDA path (mvin_importance through MVIN.eval_case_study) on a KG whose sampled relation ids are the LOGGED
    ones, with Rel . w_r = the fitted per-relation scores: importance_list_0 / _1 equal the attention the TensorFlow
    reference logged (18 trained models; 4 of them here)."""
    from mvin_b200 import MVIN
    from tests.synth import make_args
    name, r0, a0, r1, a1 = CASES[case]
    n_rel = int(max(r0.max(), r1.max())) + 1
    s, seen = _fit_scores(np.concatenate([r0.reshape(-1, 8), r1.reshape(-1, 8)]),
                          np.concatenate([a0.reshape(-1, 8), a1.reshape(-1, 8)]), n_rel)
    N, K, d = 48, 8, 8
    # entity i < N: item i, children N + 8 i + k; their relation rows are the logged hop-0 / hop-1 ids
    n_entity = N + N * K
    adj_e = np.zeros((n_entity, K), dtype=np.int64)
    adj_r = np.zeros((n_entity, K), dtype=np.int64)
    for i in range(N):
        adj_e[i] = N + K * i + np.arange(K)
        adj_r[i] = r0[i]
        for k in range(K):
            adj_e[N + K * i + k] = (np.arange(K) * 7 + i) % n_entity
            adj_r[N + K * i + k] = r1[i, k]
    args = make_args(dim=d, neighbor_sample_size=K, h_hop=2, p_hop=1, n_memory=4, batch_size=N)
    model = MVIN(args, 5, n_entity, n_rel, adj_e, adj_r)
    P = model.named_parameters()
    rel_emb = np.zeros((n_rel, d), dtype=np.float32)
    rel_emb[:, 0] = s
    urh = np.random.RandomState(0).randn(*P["agg_0_0_urh_weights"].shape).astype(np.float32)
    urh.reshape(-1)[d:2 * d] = 0
    urh.reshape(-1)[d] = 1.0                                   # relation third: picks component 0 of Rel
    P["relation_emb_matrix"], P["agg_0_0_urh_weights"] = rel_emb, urh
    model.load_named_parameters(P)
    rng = np.random.RandomState(1)
    fd = {model.user_indices: rng.randint(0, 5, N), model.item_indices: np.arange(N, dtype=np.int64),
          model.labels: np.ones(N, np.float32), model.memories_h[0]: rng.randint(0, n_entity, (N, 4)).astype(np.int32),
          model.memories_r[0]: rng.randint(0, n_rel, (N, 4)).astype(np.int32),
          model.memories_t[0]: rng.randint(0, n_entity, (N, 4)).astype(np.int32)}
    out = model.eval_case_study(None, fd)
    assert np.array_equal(out[4][0], r0[:N]) and np.array_equal(out[4][1].reshape(N, K, K), r1[:N])   # bit-exact ids
    imp0, imp1 = out[5].reshape(N, K), out[6].reshape(N, K, K)
    assert (np.abs(imp0 - a0[:N]) / np.maximum(a0[:N], 1e-6)).max() < 3e-4
    assert (np.abs(imp1 - a1[:N]) / np.maximum(a1[:N], 1e-6)).max() < 3e-4
