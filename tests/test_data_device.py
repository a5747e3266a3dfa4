"""GPU construction of the sampled adjacency against the rules of contruct_random_adj
(data_loader_user_set.py:375-388): the reference is unseeded, so parity is on the properties its sampler guarantees."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _kg(n_entity=500, n_rel=7, T=4000, seed=0):
    rng = np.random.RandomState(seed)
    h = rng.randint(0, n_entity - 20, T)             # the last 20 entities stay isolated
    t = rng.randint(0, n_entity - 20, T)
    hub = rng.rand(T) < 0.2
    t[hub] = 3                                       # entity 3 is a hub
    return np.stack([h, rng.randint(0, n_rel, T), t], 1).astype(np.int64)


@pytest.mark.parametrize("K", [4, 16, 64])
def test_sampled_adjacency_properties(K):
    from mvin_b200 import data as D
    n_entity = 500
    kg = _kg(n_entity)
    packed, adj_e, adj_r, (indptr, nbr, rel, edges) = D.sample_adjacency_device(kg, n_entity, K, "cuda", seed=7,
                                                                               want_edges=True)
    indptr, nbr, rel, edges = (x.cpu().numpy() for x in (indptr, nbr, rel, edges))
    adj_e, adj_r, packed = adj_e.cpu().numpy(), adj_r.cpu().numpy(), packed.cpu().numpy()
    ref_indptr, ref_nbr, ref_rel = D.build_undirected_csr(kg, n_entity)
    assert np.array_equal(indptr, ref_indptr)
    deg = np.diff(indptr)
    # the CSR rows hold the same (neighbour, relation) multisets as the host construction
    for e in (0, 3, 17, 250):
        a = sorted(zip(nbr[indptr[e]:indptr[e + 1]], rel[indptr[e]:indptr[e + 1]]))
        b = sorted(zip(ref_nbr[ref_indptr[e]:ref_indptr[e + 1]], ref_rel[ref_indptr[e]:ref_indptr[e + 1]]))
        assert a == b
    assert np.array_equal(packed[:, 0], adj_e) and np.array_equal(packed[:, 1], adj_r)
    for e in range(n_entity):
        if deg[e] == 0:
            assert not adj_e[e].any() and not adj_r[e].any() and (edges[e] == -1).all()     # :376-377 zero rows
            continue
        assert ((edges[e] >= indptr[e]) & (edges[e] < indptr[e + 1])).all()                 # true neighbours only
        assert np.array_equal(adj_e[e], nbr[edges[e]]) and np.array_equal(adj_r[e], rel[edges[e]])
        if deg[e] >= K:
            assert len(set(edges[e].tolist())) == K                                         # replace=False
    assert (deg >= K).any() and ((deg > 0) & (deg < K)).any() if K > 4 else True
    # reproducible for a seed, different for another
    again = D.sample_adjacency_device(kg, n_entity, K, "cuda", seed=7)[1].cpu().numpy()
    other = D.sample_adjacency_device(kg, n_entity, K, "cuda", seed=8)[1].cpu().numpy()
    assert np.array_equal(again, adj_e) and not np.array_equal(other, adj_e)


def test_sampled_adjacency_is_uniform_on_a_hub():
    """Every edge of a row with degree >= K is picked with probability K / deg, in every position with 1 / deg."""
    from mvin_b200 import data as D
    n_entity, K, trials = 500, 8, 400
    kg = _kg(n_entity)
    counts = None
    first = None
    for s in range(trials):
        _, _, _, (indptr, _, _, edges) = D.sample_adjacency_device(kg, n_entity, K, "cuda", seed=1000 + s, want_edges=True)
        lo, hi = int(indptr[3]), int(indptr[4])
        e = (edges[3] - lo).cpu().numpy()
        if counts is None:
            counts, first = np.zeros(hi - lo), np.zeros(hi - lo)
        counts[e] += 1
        first[e[0]] += 1
    deg = counts.size
    expect = trials * K / deg
    # chi-square with deg-1 dof: mean deg-1, sd sqrt(2(deg-1)); allow 5 sd
    chi2 = ((counts - expect) ** 2 / expect).sum()
    assert abs(chi2 - (deg - 1)) < 5 * np.sqrt(2 * (deg - 1)), (chi2, deg)
    assert first.max() <= 6 + 6 * trials / deg
