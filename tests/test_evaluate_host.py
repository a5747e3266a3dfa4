"""Host-side logic of mvin_b200/evaluate.py (packing of (user, candidate) pairs into batches, record handling) with a
stub model on the CPU -- the scoring and metric kernels themselves are covered by the -m gpu parity tests."""
import numpy as np
import torch

from mvin_b200 import evaluate as E


class StubModel:
    """score(user, item) = a fixed function of the ids and of the user's first ripple head; records every batch."""
    device = torch.device("cpu")
    batch_size = 16

    def __init__(self, uts, bound):
        self.uts_np = uts
        self._uts = torch.from_numpy(uts) if bound else None
        self.batches = []

    def gather_feed(self, users):
        blk = self._uts[users]                                       # [B, p, 3, m]
        return tuple(blk[:, :, c, :].permute(1, 0, 2).contiguous() for c in range(3))

    def forward_device(self, users, items, mem_h, mem_r, mem_t, scores=None, scores_normalized=None):
        assert users.shape == items.shape and users.shape[0] <= self.batch_size
        assert mem_h.shape == (self.uts_np.shape[1], users.shape[0], self.uts_np.shape[3]) and mem_h.dtype == torch.int32
        assert np.array_equal(mem_t.numpy(), self.uts_np[users.numpy()][:, :, 2].transpose(1, 0, 2))
        self.batches.append(users.shape[0])
        val = ((items * 7919 + users * 104729 + mem_h[0, :, 0].long()) % 1000).float() / 1000.0
        scores_normalized.copy_(val)

    def topk_metrics_device(self, scores, relevant, n_cand, n_answers, k_list):
        self.captured = (scores.clone(), relevant.clone(), n_cand.clone(), n_answers.clone(), list(k_list))
        z = np.zeros((scores.shape[0], len(k_list)), dtype=np.float32)
        return z + 0.25, z + 0.5, z + 0.75

    def ctr_metrics_device(self, scores_n, labels):
        return float(scores_n.mean()), float(labels.mean()), float(scores_n.shape[0])


def make_uts(n_user=12, p=2, m=4, seed=0):
    return np.random.RandomState(seed).randint(0, 50, size=(n_user, p, 3, m)).astype(np.int32)


def expected_score(uts, u, it):
    return np.float32(((it * 7919 + u * 104729 + int(uts[u, 0, 0, 0])) % 1000) / 1000.0)


def test_topk_eval_packs_pairs_and_builds_the_metric_inputs():
    uts = make_uts()
    item_set = set(range(23))
    train_record = {u: {u % 5, 7, (u * 3) % 23} for u in range(12)}
    test_record = {1: {2, 3}, 4: {22}, 5: {0, 1, 9, 11}, 9: {8}}
    for bound in (False, True):
        model = StubModel(uts, bound)
        prec, rec, ndcg, a, b = E.topk_eval(None, None, uts, model, [0, 1, 4, 5, 7, 9], train_record, {}, test_record,
                                            item_set, [1, 5, 10], 16, mode="test")
        assert (a, b) == (None, None)
        assert prec == [0.25] * 3 and rec == [0.5] * 3 and ndcg == [0.75] * 3
        scores, relevant, n_cand, n_answers, k_list = model.captured
        users = [1, 4, 5, 9]                                         # users outside the test record are skipped
        cands = [list(item_set - train_record[u]) for u in users]
        assert n_cand.tolist() == [len(c) for c in cands] and n_answers.tolist() == [2, 1, 4, 1] and k_list == [1, 5, 10]
        total = sum(len(c) for c in cands)
        assert sum(model.batches) == total and max(model.batches) <= 16          # packed: no padded pairs scored
        assert len(model.batches) == -(-total // 16)
        for i, u in enumerate(users):
            for j, it in enumerate(cands[i]):
                assert scores[i, j].item() == expected_score(uts, u, it)
                assert bool(relevant[i, j]) == (it in test_record[u])
            assert not relevant[i, len(cands[i]):].any()


def test_topk_eval_eval_mode_and_empty_user_list():
    uts = make_uts()
    model = StubModel(uts, True)
    out = E.topk_eval(None, None, uts, model, [3], {3: set()}, {3: {1}}, {}, set(range(5)), [1, 2], 16, mode="eval")
    assert model.captured[3].tolist() == [1] and model.captured[2].tolist() == [5]
    out = E.topk_eval(None, None, uts, model, [3], {3: set()}, {3: {1}}, {}, set(range(5)), [1, 2], 16, mode="test")
    assert all(np.isnan(v) for v in out[0] + out[1] + out[2])


def test_batch_size_is_capped_by_the_model():
    uts = make_uts()
    model = StubModel(uts, True)
    E.score_candidates(model, [0, 1], [list(range(20)), list(range(15))], uts, batch_size=4096)
    assert model.batches == [16, 16, 3]


def test_ctr_eval_walks_full_batches_only():
    uts = make_uts()
    rng = np.random.RandomState(1)
    data = np.stack([rng.randint(0, 12, 50), rng.randint(0, 23, 50), rng.randint(0, 2, 50)], axis=1)
    for bound in (False, True):
        model = StubModel(uts, bound)
        auc_l, acc_l, f1_l, auc, acc, f1 = E.ctr_eval(None, None, None, model, data, uts, 16)
        assert model.batches == [16, 16, 16] and len(auc_l) == 3       # the 2-row tail is dropped (util.py:49)
        for b in range(3):
            rows = data[b * 16:(b + 1) * 16]
            want = np.mean([expected_score(uts, u, it) for u, it, _ in rows])
            assert abs(auc_l[b] - want) < 1e-6 and abs(acc_l[b] - rows[:, 2].mean()) < 1e-6
        assert abs(auc - np.mean(auc_l)) < 1e-9 and f1 == 16.0
    out = E.ctr_eval(None, None, None, StubModel(uts, True), data[:5], uts, 16)
    assert out[0] == [] and np.isnan(out[3])


def test_train_epoch_walks_full_batches_in_permuted_order():
    from mvin_b200 import loop

    class LoopStub(StubModel):
        n_shards = 1

        def __init__(self, uts):
            super().__init__(uts, True)
            self.seen, self.adam = [], 0

        def forward_device(self, users, items, mem_h, mem_r, mem_t, scores=None, scores_normalized=None):
            assert users.is_contiguous() and items.is_contiguous() and users.dtype == torch.int64
            assert np.array_equal(mem_r.numpy(), self.uts_np[users.numpy()][:, :, 1].transpose(1, 0, 2))
            self.cur = (users.clone(), items.clone())

        def backward_device(self, labels, losses):
            assert labels.dtype == torch.float32 and labels.is_contiguous() and losses.shape == (4,)
            self.seen.append(torch.stack([self.cur[0], self.cur[1], labels.long()], dim=1))
            losses.fill_(float(len(self.seen)))

        def adam_step_device(self):
            self.adam += 1

    uts = make_uts()
    rng = np.random.RandomState(3)
    data = np.stack([rng.randint(0, 12, 53), rng.randint(0, 23, 53), rng.randint(0, 2, 53)], axis=1)
    model = LoopStub(uts)
    d = loop.upload_interactions(model, data)
    out = loop.train_epoch(model, d, 16, shuffle=False)
    assert out.shape == (3, 4) and out[:, 0].tolist() == [1.0, 2.0, 3.0] and model.adam == 3
    assert np.array_equal(torch.cat(model.seen).numpy(), data[:48])               # in order, 5-row tail skipped
    model = LoopStub(uts)
    g = torch.Generator().manual_seed(5)
    loop.train_epoch(model, d, 16, shuffle=True, generator=g, apply_adam=False)
    got = torch.cat(model.seen).numpy()
    assert model.adam == 0 and got.shape == (48, 3) and not np.array_equal(got, data[:48])
    rows = {tuple(r) for r in data.tolist()}
    assert all(tuple(r) in rows for r in got.tolist())
    want = data[torch.randperm(53, generator=torch.Generator().manual_seed(5)).numpy()][:48]
    assert np.array_equal(got, want)
    assert loop.train_epoch(LoopStub(uts), d[:7], 16).shape == (0, 4)
    unbound = StubModel(uts, False)
    unbound.n_shards = 1
    try:
        loop.train_epoch(unbound, d, 16)
        raise AssertionError("expected RuntimeError")
    except RuntimeError:
        pass


def test_drivers_accept_the_reference_dict_of_ripple_sets():
    """The reference's user_triplet_set is a defaultdict(user -> int32 [p, 3, m]) in which users without history are
    absent (data_loader_user_set.py:396-402): the drivers index it per user (util.py:210-217) and the packer used by
    MVIN.bind_user_triplet_set zero-fills absent users instead of tripping over the defaultdict's []."""
    import collections
    from mvin_b200.model import pack_user_triplet_set
    uts = make_uts()
    as_dict = collections.defaultdict(list)
    for u in range(uts.shape[0]):
        if u != 3:                                                   # user 3 has no history: absent from the mapping
            as_dict[u] = uts[u]
    packed = pack_user_triplet_set(as_dict, uts.shape[0], 2, 4)
    assert packed.shape == uts.shape and packed.dtype == np.int32
    assert np.array_equal(np.delete(packed, 3, axis=0), np.delete(uts, 3, axis=0)) and not packed[3].any()
    assert 3 not in as_dict                                          # packing did not insert keys into the defaultdict
    assert np.array_equal(pack_user_triplet_set(uts, uts.shape[0], 2, 4), uts)
    # unbound model: the drivers stack the dict rows per batch, same scores as with the packed array
    rng = np.random.RandomState(1)
    users = rng.choice([u for u in range(12) if u != 3], 40)
    data = np.stack([users, rng.randint(0, 23, 40), rng.randint(0, 2, 40)], axis=1)
    res_arr = E.ctr_eval(None, None, None, StubModel(uts, False), data, uts, 16)
    res_map = E.ctr_eval(None, None, None, StubModel(uts, False), data, as_dict, 16)
    assert res_arr[0] == res_map[0] and res_arr[3] == res_map[3]
    item_set = set(range(23))
    train_record = {u: {u % 5, 7} for u in range(12)}
    test_record = {1: {2, 3}, 4: {22}}
    m1, m2 = StubModel(uts, False), StubModel(uts, False)
    E.topk_eval(None, None, uts, m1, [1, 4], train_record, {}, test_record, item_set, [1, 5], 16)
    E.topk_eval(None, None, as_dict, m2, [1, 4], train_record, {}, test_record, item_set, [1, 5], 16)
    assert torch.equal(m1.captured[0], m2.captured[0])


def test_rows_into_matches_numpy_for_every_row_flavour():
    """mvin_b200.model.rows_into: the list-of-rows feed of train.py:118-120 into the pinned staging buffer."""
    import numpy as np
    from mvin_b200.model import rows_into
    rng = np.random.RandomState(0)
    B, m = 257, 6
    base = rng.randint(0, 1 << 30, size=(B, 3, m)).astype(np.int32)
    rows = [base[b][1] for b in range(B)]                     # contiguous int32 views, as user_triplet_set[u][hop][k]
    want = np.stack(rows)
    for flavour in ("int32 views", "int64", "strided", "lists", "mixed tail"):
        if flavour == "int64":
            v = [r.astype(np.int64) for r in rows]
        elif flavour == "strided":
            wide = rng.randint(0, 9, size=(B, 2 * m)).astype(np.int32)
            wide[:, ::2] = want
            v = [wide[b, ::2] for b in range(B)]
        elif flavour == "lists":
            v = [r.tolist() for r in rows]
        elif flavour == "mixed tail":
            v = rows[:-1] + [rows[-1].astype(np.int64)]       # byte length differs -> NumPy path
        else:
            v = rows
        dst = np.full((B, m), -1, dtype=np.int32)
        rows_into(dst, v)
        assert np.array_equal(dst, want), flavour
