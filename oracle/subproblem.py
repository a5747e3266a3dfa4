"""Spot checks of FULL-SIZE runs against the oracle (test infrastructure, like everything under oracle/).

The oracle materialises the reference's intermediates ([B, K^L, 3d] concats, [B, m, d, d] relation matrices), so it
cannot run a BASELINE-sized batch, and at C5 it cannot even hold the entity table.  The forward pass has no cross-pair
term (model.py:158: one score per (user, item) row), so a handful of pairs of a full-size batch can be checked
exactly: this module cuts the sub-problem those pairs touch out of the big one -- their multi-hop neighbourhoods
(model.py:243-256), ripple memories and the embedding rows behind them -- renumbers the entities / users compactly and
hands the result to `mvin_oracle.forward`.  Integer neighbour ids are compared bit for bit after mapping back.

Used by tests/ (full-size parity cases) and by bench.py's `parity_check` (the checker, never the thing measured).
"""
from __future__ import annotations

import copy
from typing import Callable, Dict, List, Sequence

import numpy as np
import torch

from . import mvin_oracle as orc


def extract(items: np.ndarray, users: np.ndarray, mem_h: Sequence[np.ndarray], mem_r: Sequence[np.ndarray],
            mem_t: Sequence[np.ndarray], L: int, adj_lookup: Callable, entity_rows: Callable, user_rows: Callable):
    """items / users int64 [n]; mem_x[hop] int32 [n, m]; adj_lookup(ids) -> (nbr [len, K], rel [len, K]) int64;
    entity_rows(ids) / user_rows(ids) -> float32 [len, d].  Returns a dict with the compact problem and `entities`
    (the ORIGINAL ids of every level, what get_neighbors must reproduce)."""
    n = items.shape[0]
    ents = [np.asarray(items, dtype=np.int64).reshape(n, 1)]
    rels, rows = [], {}
    for i in range(L):
        flat = ents[i].ravel()
        uniq_lv, inv = np.unique(flat, return_inverse=True)
        nbr_u, rel_u = adj_lookup(uniq_lv)
        for e, a, b in zip(uniq_lv, nbr_u, rel_u):
            rows[int(e)] = (a, b)
        ents.append(nbr_u[inv].reshape(n, -1))                      # child k of node j at j*K+k (model.py:251)
        rels.append(rel_u[inv].reshape(n, -1))
    touched = [e.ravel() for e in ents] + [np.asarray(m, dtype=np.int64).ravel() for m in list(mem_h) + list(mem_t)]
    uniq = np.unique(np.concatenate(touched))
    remap = lambda a: np.searchsorted(uniq, a)
    K = ents[1].shape[1] if L > 0 else 1
    adj_e = np.zeros((uniq.size, K), dtype=np.int64)
    adj_r = np.zeros((uniq.size, K), dtype=np.int64)
    for e, (a, b) in rows.items():
        adj_e[remap(e)] = remap(a)
        adj_r[remap(e)] = b
    u_uniq = np.unique(users)
    return dict(uniq=uniq, entities=ents, relations=rels, adj_entity=adj_e, adj_relation=adj_r,
                items=remap(np.asarray(items, dtype=np.int64)), users=np.searchsorted(u_uniq, users),
                mem_h=[remap(np.asarray(m, dtype=np.int64)).astype(np.int32) for m in mem_h],
                mem_r=[np.asarray(m, dtype=np.int32) for m in mem_r],
                mem_t=[remap(np.asarray(m, dtype=np.int64)).astype(np.int32) for m in mem_t],
                entity_emb=entity_rows(uniq), user_emb=user_rows(u_uniq))


def oracle_scores(sub: Dict, cfg, small_params: Dict[str, np.ndarray]):
    """Scores of the compact problem.  `small_params`: every parameter under the reference's variable names except the
    entity and user tables (taken from `sub`)."""
    cfg = copy.copy(cfg)
    cfg.batch_size = int(sub["items"].shape[0])
    P = {k: torch.as_tensor(np.asarray(v), dtype=torch.float32) for k, v in small_params.items()}
    P["entity_emb_matrix"] = torch.as_tensor(sub["entity_emb"], dtype=torch.float32)
    P["user_emb_matrix"] = torch.as_tensor(sub["user_emb"], dtype=torch.float32)
    with torch.no_grad():
        out = orc.forward(P, cfg, sub["adj_entity"], sub["adj_relation"], sub["users"], sub["items"], sub["mem_h"],
                          sub["mem_r"], sub["mem_t"])
    ents_back = [sub["uniq"][e] for e in out.entities]              # compact ids -> original ids
    return out.scores.numpy(), ents_back, out.relations


def check_model_pairs(model, cfg, users, items, mem_h, mem_r, mem_t, scores, idx) -> Dict:
    """Compare `scores` (float32 [B], what the CUDA path produced for the batch) at positions `idx` with the oracle on the
    extracted sub-problem, and the device's integer neighbour expansion of those items with the oracle's.
    users / items int64 [B], mem_x int32 [max(1,p), B, m] -- NumPy (host copies of the batch).  In a sharded
    multi-rank run every rank must call this (the entity-row lookup is a collective); only rank 0's result counts."""
    idx = np.asarray(idx, dtype=np.int64)
    K, L = model.n_neighbor, model.h_hop * model.n_mix_hop
    adj = model.adj_packed

    def adj_lookup(ids):
        rec = adj[torch.from_numpy(ids).to(adj.device)].cpu().numpy().astype(np.int64)     # [len, 2, K]
        return rec[:, 0, :], rec[:, 1, :]

    user_tab = model.params["user_emb"]
    sub = extract(items[idx], users[idx], [m[idx] for m in mem_h], [m[idx] for m in mem_r], [m[idx] for m in mem_t], L,
                  adj_lookup, model.entity_rows,
                  lambda ids: user_tab[torch.from_numpy(ids).to(user_tab.device)].cpu().numpy())
    want, ents, rels = oracle_scores(sub, cfg, model.named_parameters(tables=False))
    got = np.asarray(scores, dtype=np.float64)[idx]
    denom = np.maximum(np.abs(want), np.mean(np.abs(want)) + 1e-30)
    d_ents, d_rels = model.get_neighbors(items[idx])
    ids_ok = all(np.array_equal(a, b) for a, b in zip(d_ents, ents)) and \
        all(np.array_equal(a, b) for a, b in zip(d_rels, rels))
    return {"pairs": int(idx.size), "max_rel_err": float(np.max(np.abs(got - want) / denom)),
            "ids_bit_exact": bool(ids_ok), "tolerance": 1e-4,
            "ids_compared": int(sum(e.size for e in ents) + sum(r.size for r in rels))}
