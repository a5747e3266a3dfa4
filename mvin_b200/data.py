"""Host-side producers of the hot path's inputs, in the reference's layouts (vectorised NumPy).

Mirrors src/model/MVIN/data_loader_user_set.py of the reference:
  build_undirected_csr   <- construct_kg            :324-343  (undirected KG; CSR instead of a dict of lists)
  sample_adjacency       <- contruct_random_adj     :375-388  (K neighbours per entity, without replacement when
                                                               degree >= K, else with; isolated entities keep zeros)
  build_ripple_sets      <- get_user_triplet_set    :392-441  (per user and hop: <= 16 edges per head, then m
                                                               sampled, with replacement only when fewer than m)
  get_feed_dict          <- train.py:112-122                  (same keys / shapes; one fancy-index instead of
                                                               3 p B Python row picks)
and a seeded synthetic generator of KGs / interaction data with a given dataset's shape (there are no datasets on
the GPU box).  Sampling is distribution-equivalent to the reference, not stream-equivalent (the reference itself is
unseeded).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

# dataset shapes (verified by loading the shipped .npy / .csv files)
# `pop` = (n_distinct_items, s, c): item popularity P(rank) ~ (rank + c)^-s over the items that actually occur,
# fitted to the top-1 / top-100 / top-1000 interaction shares of the shipped train_pd.csv (MovieLens-1M: 0.35 % /
# 17.8 % / 64.9 %; last-fm: 0.10 % / 4.3 % / 20 %; amazon-book: 0.10 % / 4.1 % / 22 %).
DATASET_SHAPES = {
    "MovieLens-1M": dict(n_entity=182_011, n_relation=12, n_triples=1_241_995, n_user=6_036, n_item=2_445,
                         n_interactions=753_772, pop=(2_445, 0.60, 20.0)),
    "last-fm_50core": dict(n_entity=106_389, n_relation=9, n_triples=464_567, n_user=23_554, n_item=48_092,
                           n_interactions=1_000_000, pop=(15_471, 0.40, 2.0)),
    "amazon-book_20core": dict(n_entity=113_487, n_relation=39, n_triples=2_557_746, n_user=70_585, n_item=24_915,
                               n_interactions=600_000, pop=(9_854, 0.30, 0.0)),
}


def synthetic_kg(n_entity: int, n_relation: int, n_triples: int, seed: int = 2020, skew: float = 0.8) -> np.ndarray:
    """int64 [n_triples, 3] (head, relation, tail).  Tails follow a power law P(rank) ~ rank^-skew; skew = 0.8
    reproduces the hub structure of the shipped MovieLens-1M KG (max degree ~3e4, the top entity holding ~1.5 % and
    the top 100 ~10 % of all sampled-adjacency slots; calibrated against data/MovieLens-1M/kg_final.npy).  Every
    entity appears at least once as a head so no adjacency row is empty (true of the three shipped KGs)."""
    rng = np.random.RandomState(seed)
    n_triples = max(n_triples, n_entity)
    heads = np.concatenate([np.arange(n_entity), rng.randint(0, n_entity, size=n_triples - n_entity)])
    cdf = np.cumsum((np.arange(n_entity) + 1.0) ** -skew)
    cdf /= cdf[-1]
    tails = np.minimum(np.searchsorted(cdf, rng.rand(n_triples)), n_entity - 1)
    perm = rng.permutation(n_entity)                # hubs are spread over the id space
    tails = perm[tails]
    rels = rng.randint(0, n_relation, size=n_triples)
    return np.stack([heads, rels, tails], axis=1).astype(np.int64)


def build_undirected_csr(kg_np: np.ndarray, n_entity: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(indptr [n_entity+1], nbr [2T], rel [2T]): for every triple (h, r, t) both h->(t, r) and t->(h, r)."""
    h, r, t = kg_np[:, 0], kg_np[:, 1], kg_np[:, 2]
    src = np.concatenate([h, t])
    dst = np.concatenate([t, h])
    rr = np.concatenate([r, r])
    order = np.argsort(src, kind="stable")
    counts = np.bincount(src, minlength=n_entity)
    indptr = np.zeros(n_entity + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    return indptr, dst[order].astype(np.int64), rr[order].astype(np.int64)


def sample_adjacency(indptr, nbr, rel, K: int, seed: int = 2020) -> Tuple[np.ndarray, np.ndarray]:
    """adj_entity, adj_relation int64 [n_entity, K]."""
    rng = np.random.RandomState(seed)
    n_entity = indptr.shape[0] - 1
    deg = np.diff(indptr)
    adj_e = np.zeros((n_entity, K), dtype=np.int64)
    adj_r = np.zeros((n_entity, K), dtype=np.int64)
    # degree < K: with replacement
    small = np.nonzero((deg > 0) & (deg < K))[0]
    if small.size:
        pick = (rng.rand(small.size, K) * deg[small, None]).astype(np.int64) + indptr[small, None]
        adj_e[small], adj_r[small] = nbr[pick], rel[pick]
    # degree >= K: without replacement = first K of a random permutation of the row
    big = np.nonzero(deg >= K)[0]
    if big.size:
        keys = rng.rand(nbr.shape[0])
        row_of = np.repeat(np.arange(n_entity), deg)
        order = np.lexsort((keys, row_of))          # within each row, a uniformly random order
        pick = indptr[big, None] + np.arange(K)[None, :]
        sel = order[pick]
        adj_e[big], adj_r[big] = nbr[sel], rel[sel]
    return adj_e, adj_r


def synthetic_interactions(n_user: int, n_item: int, n_interactions: int, seed: int = 2020,
                           pop=None) -> np.ndarray:
    """int64 [N, 3] (user, item, label in {0,1}), half positive (the reference's negative sampling ratio).  Item
    popularity follows the fitted law of DATASET_SHAPES[...]['pop'] over a random subset of the item ids (item ids
    are sparse in last-fm / amazon-book)."""
    rng = np.random.RandomState(seed)
    users = rng.randint(0, n_user, size=n_interactions)
    n_distinct, s, c = pop if pop is not None else (n_item, 0.6, 20.0)
    n_distinct = min(n_distinct, n_item)
    cdf = np.cumsum((np.arange(1, n_distinct + 1) + c) ** -s)
    cdf /= cdf[-1]
    ranks = np.minimum(np.searchsorted(cdf, rng.rand(n_interactions)), n_distinct - 1)
    items = rng.permutation(n_item)[:n_distinct][ranks]
    labels = rng.randint(0, 2, size=n_interactions)
    return np.stack([users, items, labels], axis=1).astype(np.int64)


def user_history(data: np.ndarray, n_user: int, n_item: int, seed: int = 2020) -> Dict[int, np.ndarray]:
    """user -> positive items (data_loader_user_set.py:get_user_record); users without one get a random item so
    every user has a ripple set (the reference drops such users instead, :77-97)."""
    pos = data[data[:, 2] == 1]
    order = np.argsort(pos[:, 0], kind="stable")
    pos = pos[order]
    bounds = np.searchsorted(pos[:, 0], np.arange(n_user + 1))
    rng = np.random.RandomState(seed)
    hist = {}
    for u in range(n_user):
        items = np.unique(pos[bounds[u]:bounds[u + 1], 1])
        hist[u] = items if items.size else rng.randint(0, n_item, size=1)
    return hist


def build_ripple_sets(indptr, nbr, rel, history: Dict[int, np.ndarray], n_user: int, p_hop: int, n_memory: int,
                      n_neighbor: int = 16, seed: int = 2020) -> np.ndarray:
    """user_triplet_set as one dense int32 array [n_user, max(1,p), 3, m] (rows h, r, t), the packed form of the
    reference's dict user -> int32 [p, 3, m] (data_loader_user_set.py:402)."""
    rng = np.random.RandomState(seed)
    P = max(1, p_hop)
    out = np.zeros((n_user, P, 3, n_memory), dtype=np.int32)
    for u in range(n_user):
        tails = history[u]
        prev = None
        for hop in range(P):
            heads = np.asarray(tails, dtype=np.int64)
            deg = indptr[heads + 1] - indptr[heads]
            take = np.minimum(deg, n_neighbor)
            tot = int(take.sum())
            if tot == 0:
                out[u, hop] = prev                  # copy the previous hop (:425-426)
                continue
            rep_head = np.repeat(heads, take)
            start = np.repeat(indptr[heads], take)
            rep_deg = np.repeat(deg, take)
            # <= n_neighbor edges per head: random offsets (with replacement inside a head when deg > n_neighbor)
            within = np.arange(tot) - np.repeat(np.cumsum(take) - take, take)
            off = np.where(rep_deg > n_neighbor, (rng.rand(tot) * rep_deg).astype(np.int64), within)
            eidx = start + off
            idx = rng.choice(tot, size=n_memory, replace=tot < n_memory)
            trip = np.stack([rep_head[idx], rel[eidx[idx]], nbr[eidx[idx]]]).astype(np.int32)
            out[u, hop] = trip
            prev = trip
            tails = trip[2]
    return out


def get_feed_dict(model, data: np.ndarray, user_triplet_set: np.ndarray, start: int, end: int):
    """train.py:112-122 with the packed ripple sets: same keys, values int32 [B, m] arrays."""
    users = data[start:end, 0]
    feed = {model.user_indices: users, model.item_indices: data[start:end, 1], model.labels: data[start:end, 2]}
    trip = user_triplet_set[users]                  # [B, P, 3, m]
    for i in range(trip.shape[1]):
        feed[model.memories_h[i]] = trip[:, i, 0]
        feed[model.memories_r[i]] = trip[:, i, 1]
        feed[model.memories_t[i]] = trip[:, i, 2]
    return feed


def stacked_memories(user_triplet_set: np.ndarray, users: np.ndarray):
    """The C-ABI feed form: mem_h, mem_r, mem_t int32 [P, B, m] (contiguous)."""
    trip = user_triplet_set[users]                  # [B, P, 3, m]
    t = np.ascontiguousarray(trip.transpose(2, 1, 0, 3))   # [3, P, B, m]
    return t[0], t[1], t[2]


def make_synthetic_dataset(name: str, K: int, p_hop: int, n_memory: int, seed: int = 2020, n_interactions=None):
    shp = DATASET_SHAPES[name]
    kg = synthetic_kg(shp["n_entity"], shp["n_relation"], shp["n_triples"], seed)
    indptr, nbr, rel = build_undirected_csr(kg, shp["n_entity"])
    adj_e, adj_r = sample_adjacency(indptr, nbr, rel, K, seed)
    n_int = n_interactions or min(shp["n_interactions"], 400_000)
    data = synthetic_interactions(shp["n_user"], shp["n_item"], n_int, seed, pop=shp.get("pop"))
    hist = user_history(data, shp["n_user"], shp["n_item"], seed)
    uts = build_ripple_sets(indptr, nbr, rel, hist, shp["n_user"], p_hop, n_memory, seed=seed)
    return dict(shape=shp, adj_entity=adj_e, adj_relation=adj_r, data=data, user_triplet_set=uts)


# ----------------------------------------------------------------------------------------------------------------
# BASELINE.json config C5 (synthetic KG, 100 M entities / 1 B edges, K = 64): far too large for the reference's
# host-side Python loops and int64 NumPy adjacency, so it is generated on the device, directly in the packed
# layout the kernels read: degree ~ clamp(Poisson(20), 1, 48) undirected neighbours drawn
# uniformly over the entities, relation uniform, then K slots sampled WITH replacement from the entity's own
# neighbour list (the reference's deg < K rule, data_loader_user_set.py:383-386).  Same seed -> same graph on
# every rank.
# ----------------------------------------------------------------------------------------------------------------
def synthetic_packed_adjacency_device(n_entity: int, n_relation: int, K: int, device, seed: int = 1234,
                                      mean_degree: float = 20.0, max_degree: int = 48, chunk: int = 1 << 21):
    import torch
    gen = torch.Generator(device=device).manual_seed(seed)
    adj = torch.empty((n_entity, 2, K), dtype=torch.int32, device=device)
    for start in range(0, n_entity, chunk):
        n = min(chunk, n_entity - start)
        deg = torch.poisson(torch.full((n,), mean_degree, device=device), generator=gen).clamp_(1, max_degree)
        pool_e = torch.randint(0, n_entity, (n, max_degree), dtype=torch.int32, device=device, generator=gen)
        pool_r = torch.randint(0, n_relation, (n, max_degree), dtype=torch.int32, device=device, generator=gen)
        sel = (torch.rand((n, K), device=device, generator=gen) * deg[:, None]).long().clamp_(max=max_degree - 1)
        adj[start:start + n, 0] = torch.gather(pool_e, 1, sel)
        adj[start:start + n, 1] = torch.gather(pool_r, 1, sel)
        del deg, pool_e, pool_r, sel
    return adj


def synthetic_ripple_sets_device(adj_packed, n_user: int, n_item: int, p_hop: int, n_memory: int, seed: int = 1234,
                                 chunk: int = 1 << 17):
    """user_triplet_set int32 [n_user, max(1,p), 3, m] on the device: hop-0 heads uniform over the items, relations
    and tails read from the sampled adjacency, hop h+1 heads = hop h tails (data_loader_user_set.py:407-441)."""
    import torch
    device = adj_packed.device
    K = adj_packed.shape[2]
    P = max(1, p_hop)
    gen = torch.Generator(device=device).manual_seed(seed + 1)
    out = torch.empty((n_user, P, 3, n_memory), dtype=torch.int32, device=device)
    for start in range(0, n_user, chunk):
        n = min(chunk, n_user - start)
        heads = torch.randint(0, n_item, (n, n_memory), device=device, generator=gen)
        for hop in range(P):
            slot = torch.randint(0, K, (n, n_memory), device=device, generator=gen)
            rec = adj_packed[heads.reshape(-1)]                                     # [n*m, 2, K]
            pick = slot.reshape(-1, 1)
            tails = torch.gather(rec[:, 0], 1, pick).reshape(n, n_memory)
            rels = torch.gather(rec[:, 1], 1, pick).reshape(n, n_memory)
            out[start:start + n, hop, 0] = heads.to(torch.int32)
            out[start:start + n, hop, 1] = rels
            out[start:start + n, hop, 2] = tails
            heads = tails.long()
            del rec
    return out


def stacked_memories_device(user_triplet_set, users):
    """Device version of stacked_memories: the feed of one batch gathered from the device-resident ripple sets
    (the host loop of train.py:112-122 replaced by one gather): mem_h, mem_r, mem_t int32 [P, B, m]."""
    trip = user_triplet_set[users]                       # [B, P, 3, m]
    t = trip.permute(2, 1, 0, 3).contiguous()            # [3, P, B, m]
    return t[0], t[1], t[2]


def sample_adjacency_device(kg_np, n_entity: int, K: int, device, seed: int = 2020, want_edges: bool = False):
    """GPU version of construct_kg + contruct_random_adj (data_loader_user_set.py:324-343, :375-388): the undirected
    CSR is built with device sorts, the per-entity sampling runs in mvin_sample_adjacency.  Returns
    (adj_packed int32 [n_entity, 2, K], adj_entity int64 [n_entity, K], adj_relation int64 [n_entity, K]) CUDA tensors
    (+ the CSR and the chosen edge slots when want_edges)."""
    import torch
    from . import _lib
    lib = _lib.load()
    kg = torch.as_tensor(np.ascontiguousarray(kg_np), device=device).long()
    src = torch.cat([kg[:, 0], kg[:, 2]])
    dst = torch.cat([kg[:, 2], kg[:, 0]])
    rr = torch.cat([kg[:, 1], kg[:, 1]])
    order = torch.argsort(src, stable=True)
    indptr = torch.zeros(n_entity + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(torch.bincount(src, minlength=n_entity), 0)
    nbr, rel = dst[order].int().contiguous(), rr[order].int().contiguous()
    packed = torch.zeros((n_entity, 2, K), dtype=torch.int32, device=device)
    adj_e = torch.zeros((n_entity, K), dtype=torch.int64, device=device)
    adj_r = torch.zeros((n_entity, K), dtype=torch.int64, device=device)
    edges = torch.empty((n_entity, K), dtype=torch.int64, device=device) if want_edges else None
    _lib.check(lib.mvin_sample_adjacency(indptr.data_ptr(), nbr.data_ptr(), rel.data_ptr(), n_entity, K, seed,
                                         packed.data_ptr(), adj_e.data_ptr(), adj_r.data_ptr(),
                                         edges.data_ptr() if want_edges else None,
                                         torch.cuda.current_stream(device).cuda_stream), "mvin_sample_adjacency")
    if want_edges:
        return packed, adj_e, adj_r, (indptr, nbr, rel, edges)
    return packed, adj_e, adj_r


def build_ripple_sets_device(csr, history: Dict[int, np.ndarray], n_user: int, p_hop: int, n_memory: int, device,
                             n_neighbor: int = 16, seed: int = 2020, want_slots: bool = False):
    """GPU version of get_user_triplet_set (data_loader_user_set.py:392-441): csr = (indptr int64, nbr int32, rel int32)
    CUDA tensors (sample_adjacency_device(..., want_edges=True)[3][:3]); history: user -> positive items.  Returns the
    packed ripple sets int32 [n_user, max(1,p), 3, m] as a CUDA tensor -- ready for MVIN.bind_user_triplet_set."""
    import torch
    from . import _lib
    lib = _lib.load()
    indptr, nbr, rel = csr
    lens = np.array([len(history[u]) for u in range(n_user)], dtype=np.int64)
    hist_ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)])).to(device)
    hist_items = torch.from_numpy(np.concatenate([np.asarray(history[u], dtype=np.int32) for u in range(n_user)])).to(device)
    P = max(1, p_hop)
    uts = torch.zeros((n_user, P, 3, n_memory), dtype=torch.int32, device=device)
    slots = torch.empty((n_user, P, n_memory), dtype=torch.int64, device=device) if want_slots else None
    _lib.check(lib.mvin_build_ripple_sets(indptr.data_ptr(), nbr.data_ptr(), rel.data_ptr(), hist_ptr.data_ptr(),
                                          hist_items.data_ptr(), n_user, p_hop, n_memory, n_neighbor, seed, uts.data_ptr(),
                                          slots.data_ptr() if want_slots else None,
                                          torch.cuda.current_stream(device).cuda_stream), "mvin_build_ripple_sets")
    return (uts, slots) if want_slots else uts
