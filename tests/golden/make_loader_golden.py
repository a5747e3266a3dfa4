"""Golden fixture for the data-side rows: the reference's OWN construct_kg / contruct_random_adj
/ _get_user_triplet_set (data_loader_user_set.py:324-343, :375-388, :407-441) run here, unmodified, on a small seeded KG.

construct_kg is deterministic: its adjacency lists are stored as they are.  The two samplers draw from numpy's / random's
global generators, so what is stored is one seeded draw of each -- the tests check that the draw satisfies exactly the
properties our own samplers are tested for (true neighbours only, without replacement iff degree >= K, ...), i.e. that
those properties are the reference's and not ours.

    python tests/golden/make_loader_golden.py      # needs /root/reference; writes tests/golden/loader_small.npz
"""
import os
import random
import sys
import types

import numpy as np

REF = "/root/reference/src/model/MVIN"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "loader_small.npz")


def main():
    sys.path.insert(0, REF)
    import data_loader_user_set as ref          # plain numpy / pandas module, no TensorFlow
    rng = np.random.RandomState(0)
    n_entity, n_rel, T = 120, 5, 500
    h = rng.randint(0, n_entity - 10, T)
    t = rng.randint(0, n_entity - 10, T)
    t[rng.rand(T) < 0.25] = 7                   # a hub
    kg_np = np.stack([h, rng.randint(0, n_rel, T), t], 1).astype(np.int64)
    args = types.SimpleNamespace(neighbor_sample_size=8, p_hop=2, n_memory=16)
    kg, enti, rela = ref.construct_kg(args, kg_np)
    np.random.seed(2020)
    random.seed(2020)
    adj_e, adj_r = ref.contruct_random_adj(args, kg, n_entity)
    ref.g_kg = kg
    history = {u: sorted(set(rng.randint(0, 40, rng.randint(1, 9)).tolist())) for u in range(30)}
    trip = []
    for u in range(30):
        _, ret, _ = ref._get_user_triplet_set(u, history[u], p_hop=2, n_memory=16, n_neighbor=16)
        trip.append(np.asarray(ret, dtype=np.int32))
    # adjacency lists flattened: for entity e the (tail, relation) pairs in the reference's order
    ptr = [0]
    pairs = []
    for e in range(n_entity):
        for (tail, r) in kg.get(e, []):
            pairs.append((tail, r))
        ptr.append(len(pairs))
    hist_ptr = np.cumsum([0] + [len(history[u]) for u in range(30)])
    np.savez_compressed(OUT, kg_np=kg_np, n_entity=n_entity, kg_ptr=np.asarray(ptr), kg_pairs=np.asarray(pairs, dtype=np.int64),
                        adj_entity=adj_e, adj_relation=adj_r, hist_ptr=hist_ptr,
                        hist_items=np.concatenate([np.asarray(history[u]) for u in range(30)]),
                        user_triplet_set=np.stack(trip))
    print("wrote", OUT, os.path.getsize(OUT), "bytes; triplet set", np.stack(trip).shape)


if __name__ == "__main__":
    main()
