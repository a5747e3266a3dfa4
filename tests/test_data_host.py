"""Data-side rows on the CPU: the reference's own loader functions, run on a small seeded KG by
tests/golden/make_loader_golden.py (fixture tests/golden/loader_small.npz), against mvin_b200/data.py.

* construct_kg (data_loader_user_set.py:324-343) is deterministic: our undirected CSR must hold, entity by entity, the
  same (tail, relation) multiset.
* contruct_random_adj (:375-388) and _get_user_triplet_set (:407-441) are random: ONE checker states their rules; it is
  applied to the reference's stored draw (so the rules are the reference's, not ours), to our NumPy port, and -- in
  tests/test_data_device.py -- to the CUDA kernels.
"""
import os

import numpy as np

from mvin_b200 import data as D

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loader_small.npz")


def load():
    z = np.load(FIX)
    return {k: z[k] for k in z.files}


def neighbour_sets(z):
    ptr, pairs = z["kg_ptr"], z["kg_pairs"]
    return [sorted(map(tuple, pairs[ptr[e]:ptr[e + 1]].tolist())) for e in range(int(z["n_entity"]))]


def check_adjacency(adj_e, adj_r, nbrs, K):
    """Rules of contruct_random_adj: zero rows for entities absent from the KG; every (neighbour, relation) pair is an
    edge; K distinct edge slots when degree >= K (np.random.choice(replace=False))."""
    for e, lst in enumerate(nbrs):
        if not lst:
            assert not adj_e[e].any() and not adj_r[e].any()
            continue
        picks = list(zip(adj_e[e].tolist(), adj_r[e].tolist()))
        assert all(p in lst for p in picks), e
        if len(lst) >= K:
            # without replacement over edge SLOTS: no pair may be used more often than it occurs in the list
            for p in set(picks):
                assert picks.count(p) <= lst.count(p), (e, p)


def check_ripple_sets(uts, hist, nbrs, n_memory, n_neighbor=16):
    """Rules of _get_user_triplet_set: heads of hop 0 are history items, heads of hop h are tails of hop h-1; every
    (head, relation, tail) is an edge; distinct candidate slots when there are at least n_memory candidates."""
    for u in range(uts.shape[0]):
        for hop in range(uts.shape[1]):
            h, r, t = (uts[u, hop, i].tolist() for i in range(3))
            src = list(hist[u]) if hop == 0 else uts[u, hop - 1, 2].tolist()
            assert set(h) <= set(src), (u, hop)
            assert all((tt, rr) in nbrs[hh] for hh, rr, tt in zip(h, r, t)), (u, hop)
            total = sum(min(len(nbrs[e]), n_neighbor) for e in src)
            if total >= n_memory and len(set(src)) == len(src) and all(len(nbrs[e]) <= n_neighbor for e in src):
                # every candidate slot is a distinct (source position, edge): with unique sources and full neighbourhoods
                # no triple may repeat more often than its edge multiplicity
                trip = list(zip(h, r, t))
                for x in set(trip):
                    assert trip.count(x) <= nbrs[x[0]].count((x[2], x[1])), (u, hop, x)


def test_reference_draw_obeys_the_rules():
    z = load()
    nbrs = neighbour_sets(z)
    check_adjacency(z["adj_entity"], z["adj_relation"], nbrs, 8)
    hist = [z["hist_items"][z["hist_ptr"][u]:z["hist_ptr"][u + 1]] for u in range(30)]
    check_ripple_sets(z["user_triplet_set"], hist, nbrs, 16)
    assert z["user_triplet_set"].dtype == np.int32 and z["user_triplet_set"].shape == (30, 2, 3, 16)   # :402


def test_csr_matches_construct_kg():
    z = load()
    nbrs = neighbour_sets(z)
    indptr, nbr, rel = D.build_undirected_csr(z["kg_np"], int(z["n_entity"]))
    for e, lst in enumerate(nbrs):
        got = sorted(zip(nbr[indptr[e]:indptr[e + 1]].tolist(), rel[indptr[e]:indptr[e + 1]].tolist()))
        assert got == lst, e


def test_numpy_samplers_obey_the_same_rules():
    z = load()
    nbrs = neighbour_sets(z)
    indptr, nbr, rel = D.build_undirected_csr(z["kg_np"], int(z["n_entity"]))
    adj_e, adj_r = D.sample_adjacency(indptr, nbr, rel, 8, seed=3)
    check_adjacency(adj_e, adj_r, nbrs, 8)
    hist = {u: z["hist_items"][z["hist_ptr"][u]:z["hist_ptr"][u + 1]] for u in range(30)}
    uts = D.build_ripple_sets(indptr, nbr, rel, hist, 30, 2, 16, seed=3)
    check_ripple_sets(uts, [hist[u] for u in range(30)], nbrs, 16)
    feed = D.stacked_memories(uts, np.array([3, 3, 7]))
    assert feed[0].shape == (2, 3, 16) and np.array_equal(feed[0][:, 0], uts[3, :, 0])      # train.py:112-122 layout
