"""Seeded synthetic MVIN problems (KG adjacency, ripple memories, batch, parameters) for parity tests."""
import types

import numpy as np
import torch

from oracle import mvin_oracle as orc


def make_args(**over):
    a = dict(dataset="synthetic", load_pretrain_emb=False, h_hop=2, batch_size=64, neighbor_sample_size=8, p_hop=2,
             dim=16, l2_weight=1e-4, l2_agg_weight=1e-6, kge_weight=1e-2, lr=5e-3, save_model_name="t", n_mix_hop=1,
             n_memory=16, update_item_emb="transform_matrix", h0_att="st_att_h_set", path=None, User_orient=1,
             User_orient_rela=1, User_orient_kg_eh=1, PS_O_ft=1, wide_deep=1, PS_only=0, HO_only=0)
    a.update(over)
    return types.SimpleNamespace(**a)


def make_problem(args, n_user=50, n_entity=400, n_relation=7, n_item=60, seed=0, regime="trained", hub_frac=0.0):
    """Returns dict(adj_entity, adj_relation, users, items, labels, mem_h, mem_r, mem_t, P).  `hub_frac` routes that
    share of all neighbour slots to entity 0 (a hub) to exercise scatter-add contention."""
    rng = np.random.RandomState(seed)
    B, K, m = args.batch_size, args.neighbor_sample_size, args.n_memory
    adj_entity = rng.randint(0, n_entity, size=(n_entity, K)).astype(np.int64)
    if hub_frac > 0:
        adj_entity[rng.rand(n_entity, K) < hub_frac] = 0
    adj_relation = rng.randint(0, n_relation, size=(n_entity, K)).astype(np.int64)
    users = rng.randint(0, n_user, size=B).astype(np.int64)
    items = rng.randint(0, n_item, size=B).astype(np.int64)
    labels = rng.randint(0, 2, size=B).astype(np.float32)
    n_mem = max(1, args.p_hop)
    mem_h = [rng.randint(0, n_entity, size=(B, m)).astype(np.int32) for _ in range(n_mem)]
    mem_r = [rng.randint(0, n_relation, size=(B, m)).astype(np.int32) for _ in range(n_mem)]
    mem_t = [rng.randint(0, n_entity, size=(B, m)).astype(np.int32) for _ in range(n_mem)]
    cfg = orc.OracleConfig.from_args(args)
    P = orc.init_params(cfg, n_user, n_entity, n_relation, seed=seed + 1, regime=regime)
    if regime == "xavier":
        for k in P:
            if k.startswith("agg_") and k.endswith("_bias"):
                P[k] = torch.full_like(P[k], 0.01)
    return dict(cfg=cfg, n_user=n_user, n_entity=n_entity, n_relation=n_relation, adj_entity=adj_entity,
                adj_relation=adj_relation, users=users, items=items, labels=labels, mem_h=mem_h, mem_r=mem_r,
                mem_t=mem_t, P=P)


def feed_dict(model, prob):
    fd = {model.user_indices: prob["users"], model.item_indices: prob["items"], model.labels: prob["labels"]}
    for i in range(len(prob["mem_h"])):
        fd[model.memories_h[i]] = prob["mem_h"][i]
        fd[model.memories_r[i]] = prob["mem_r"][i]
        fd[model.memories_t[i]] = prob["mem_t"][i]
    return fd
