"""CPU oracle for the MVIN hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, not the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(``mvin_b200``) never routes through it.

It restates, op for op and with the same materialised intermediates, the TensorFlow-1 graph built by the
reference (all citations relative to ``/root/reference/src/model/MVIN/``):

* ``get_neighbors``            <- model.py:243-256
* ``key_addressing``           <- model.py:161-240 (incl. ``soft_attention_h_set`` :162-197)
* ``aggregate_delta_whole``    <- model.py:259-324
* ``sum_aggregator_urh``       <- aggregators.py:98-152
* ``forward`` (graph wiring)   <- model.py:125-159
* ``loss_terms``               <- model.py:378-412
* ``adam_step_tf1``            <- model.py:414 (tf.train.AdamOptimizer defaults, TF1 epsilon placement)
* ``build_feed``               <- train.py:112-122

The arithmetic itself lives in TensorFlow 1.13.1 (requirements.yml:47-48), which is absent from
``/root/reference`` and cannot be installed here (no cp312 wheel, no network).  The TF ops used are restated
with their published semantics in PyTorch-CPU (fp32 by default to mirror TF, fp64 switch for error
budgeting); gradients come from autograd, like TF's autodiff.

Parity pin: the reference has no tests or golden vectors.  The oracle is pinned instead against the
reference's OWN ``model.py`` / ``aggregators.py`` executed in this container under a minimal TF1-API shim
(``tests/golden/tf1_shim.py``; generator ``tests/golden/make_golden.py``; fixtures ``tests/golden/*.npz``),
so every wiring convention (concat orders, reshapes, which tensors feed which op, the L2 bookkeeping quirks)
is checked against reference code, while op numerics (matmul/softmax/...) follow the TF documentation.
The only numerical output of the real TensorFlow model that ships with the reference -- the attention weights of 18
trained models in ``case_st/amazon-book_20core/*.log`` -- is a second pin (``tests/golden/case_study_att.npz``,
``tests/test_case_study_logs.py``): ``sum_aggregator_urh`` reproduces them to print precision.  What remains
unpinned against a TF binary: the ripple side, the dense maps, the loss and Adam ("parity unpinned" for those).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


# ----------------------------------------------------------------------------------------------------------
# configuration (mirrors the fields MVIN._parse_args reads, model.py:17-47, and the ablation switches)
# ----------------------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    dim: int
    neighbor_sample_size: int  # K
    h_hop: int                 # H
    n_mix_hop: int = 1         # M
    p_hop: int = 2             # p
    n_memory: int = 64         # m
    batch_size: int = 1024
    l2_weight: float = 1e-4
    l2_agg_weight: float = 1e-6
    lr: float = 5e-4
    # --ablation switches (parameter_ablation.py:4-12 are the 'all' values)
    User_orient: bool = True
    User_orient_rela: bool = True
    User_orient_kg_eh: bool = True
    PS_O_ft: bool = True
    wide_deep: bool = True
    PS_only: bool = False
    HO_only: bool = False

    @property
    def L(self) -> int:
        return self.n_mix_hop * self.h_hop

    @classmethod
    def from_args(cls, args) -> "OracleConfig":
        g = lambda k, dflt: getattr(args, k, dflt)
        return cls(dim=args.dim, neighbor_sample_size=args.neighbor_sample_size, h_hop=args.h_hop,
                   n_mix_hop=g("n_mix_hop", 1), p_hop=args.p_hop, n_memory=args.n_memory,
                   batch_size=args.batch_size, l2_weight=g("l2_weight", 1e-4),
                   l2_agg_weight=g("l2_agg_weight", 1e-6), lr=g("lr", 5e-4),
                   User_orient=bool(g("User_orient", 1)), User_orient_rela=bool(g("User_orient_rela", 1)),
                   User_orient_kg_eh=bool(g("User_orient_kg_eh", 1)), PS_O_ft=bool(g("PS_O_ft", 1)),
                   wide_deep=bool(g("wide_deep", 1)), PS_only=bool(g("PS_only", 0)),
                   HO_only=bool(g("HO_only", 0)))


# ----------------------------------------------------------------------------------------------------------
# parameters (model.py:72-122, aggregators.py:83-93) -- same names as the reference's attributes
# ----------------------------------------------------------------------------------------------------------
def param_shapes(cfg: OracleConfig, n_user: int, n_entity: int, n_relation: int) -> Dict[str, Tuple[int, ...]]:
    d, H, M, p = cfg.dim, cfg.h_hop, cfg.n_mix_hop, cfg.p_hop
    shapes: Dict[str, Tuple[int, ...]] = {
        "user_emb_matrix": (n_user, d),                       # model.py:72-74
        "entity_emb_matrix": (n_entity, d),                   # :76-78
        "relation_emb_matrix": (n_relation, d),               # :80-82
        "relation_emb_KGE_matrix": (n_relation, d, d),        # :84-86
    }
    for n in range(M):                                        # :91-98
        shapes[f"enti_transfer_matrix_{n}"] = (d * (H + 1), d)
        shapes[f"enti_transfer_bias_{n}"] = (d,)
    shapes["user_mlp_matrix"] = (d * (p + 1 if cfg.PS_O_ft else p), d)   # :100-104
    shapes["user_mlp_bias"] = (d,)                                       # :105-106
    for e in range(M * H + 1):                                # :107-116
        shapes[f"transfer_agg_matrix_{e}"] = (d, d)
        shapes[f"transfer_agg_bias_{e}"] = (d,)
    shapes["h_emb_item_mlp_matrix"] = (2 * d, 1)              # :118-120
    shapes["h_emb_item_mlp_bias"] = (1,)                      # :121-122
    if not cfg.PS_only:
        # aggregators are created in the order n (mix layer) outer, i (hop) inner: model.py:286-291
        for n in range(M):
            for i in range(H):
                shapes[f"agg_{i}_{n}_weights"] = (d, d)       # aggregators.py:83-85
                shapes[f"agg_{i}_{n}_bias"] = (d,)            # :86-87
                shapes[f"agg_{i}_{n}_urh_weights"] = (3 * d, 1)   # :89-91
                shapes[f"agg_{i}_{n}_urh_bias"] = (1,)        # :92-93 (created, never used)
    return shapes


def _xavier_uniform(shape, gen: torch.Generator, dtype) -> torch.Tensor:
    """tf.contrib.layers.xavier_initializer (uniform=True): limit = sqrt(6 / (fan_in + fan_out)) with
    TF's fan rule (fan_in = shape[-2] * receptive, fan_out = shape[-1] * receptive; 1-D: fan_in=fan_out=n)."""
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    elif len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rec = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rec, shape[-1] * rec
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1).mul_(limit).to(dtype)


def init_params(cfg: OracleConfig, n_user: int, n_entity: int, n_relation: int, seed: int = 0,
                regime: str = "xavier", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """regime 'xavier' = the reference's init (model.py:14-15; aggregator biases zero, aggregators.py:87,93);
    regime 'trained' = trained-scale values N(0, 1/sqrt(d)) for tables and N(0, 1/sqrt(fan_in)) for weights
    (SURVEY.md section 8(d)), which make the softmaxes non-uniform and the ReLUs mixed."""
    gen = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(cfg, n_user, n_entity, n_relation).items():
        if regime == "xavier":
            if name.startswith("agg_") and name.endswith("bias"):
                t = torch.zeros(shape, dtype=dtype)
            else:
                t = _xavier_uniform(shape, gen, dtype)
        elif regime == "trained":
            if name.endswith("_emb_matrix") or name == "relation_emb_KGE_matrix":
                std = 1.0 / math.sqrt(cfg.dim)
            elif len(shape) >= 2:
                std = 1.0 / math.sqrt(shape[0])
            else:
                std = 0.1
            t = (torch.randn(shape, generator=gen, dtype=torch.float64) * std).to(dtype)
        else:
            raise ValueError(regime)
        out[name] = t
    return out


# ----------------------------------------------------------------------------------------------------------
# A3  get_neighbors  (model.py:243-256)  -- integer, bit-exact
# ----------------------------------------------------------------------------------------------------------
def get_neighbors(adj_entity: np.ndarray, adj_relation: np.ndarray, seeds: np.ndarray, n_levels: int,
                  batch_size: int) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """entities[i] int64 [B, K^i] (i=0..n_levels), relations[i] int64 [B, K^(i+1)] (i<n_levels).
    tf.gather(adj, entities[i]) has shape [B, K^i, K]; the reshape at model.py:251 flattens it row-major, so
    child k of node j sits at flat position j*K + k."""
    seeds = np.asarray(seeds, dtype=np.int64)
    entities = [seeds.reshape(-1, 1)]                         # expand_dims, :245
    relations = []
    for i in range(n_levels):                                 # :250
        ne = adj_entity[entities[i]].reshape(batch_size, -1)  # :251
        nr = adj_relation[entities[i]].reshape(batch_size, -1)  # :252
        entities.append(ne.astype(np.int64))
        relations.append(nr.astype(np.int64))
    return entities, relations


# ----------------------------------------------------------------------------------------------------------
# A7  SumAggregator_urh_matrix  (aggregators.py:98-152)
# ----------------------------------------------------------------------------------------------------------
def sum_aggregator_urh(P, name: str, cfg: OracleConfig, self_vectors, neighbor_vectors, neighbor_relations,
                       user_embeddings):
    """self [B,n,d], neighbor_vectors [B,n,K,d], neighbor_relations [B,n,K,d], user [B,d]
    -> (relu((self + agg) W + b) [B,n,d], probs_normalized [B,n,K] or None)."""
    B, n, K, d = neighbor_vectors.shape
    if cfg.User_orient_rela:
        # _mix_neighbor_vectors_urh, aggregators.py:118-146
        user = user_embeddings.reshape(B, 1, 1, d).expand(B, n, K, d)            # :121-122 (tile)
        selfv = self_vectors.unsqueeze(2).expand(B, n, K, d)                     # :126-127
        urh = torch.cat([user, neighbor_relations, selfv], dim=-1)               # :130-131  order u, r, self
        logits = (urh.reshape(-1, 3 * d) @ P[f"{name}_urh_weights"]).reshape(B, n, K)   # :133-136 (no bias)
        probs = torch.softmax(logits, dim=-1)                                    # :139
        agg = (probs.unsqueeze(-1) * neighbor_vectors).mean(dim=2)               # :141-144  reduce_MEAN
    else:
        agg = neighbor_vectors.mean(dim=2)                                       # :148-152
        probs = None
    out = (self_vectors + agg).reshape(-1, d)                                    # :108 (dropout keep_prob=1, :109)
    out = out @ P[f"{name}_weights"] + P[f"{name}_bias"]                         # :110
    out = out.reshape(B, -1, d)                                                  # :113
    return torch.relu(out), probs                                                # :116


# ----------------------------------------------------------------------------------------------------------
# A5  _key_addressing  (model.py:161-240)
# ----------------------------------------------------------------------------------------------------------
def key_addressing(P, cfg: OracleConfig, user_indices, item_indices, h_emb_list, r_emb_list, t_emb_list):
    d, m = cfg.dim, cfg.n_memory
    o_list = []
    item_embeddings = P["entity_emb_matrix"][item_indices]                       # :199 (never updated)
    if cfg.PS_O_ft:                                                              # :204-206
        # soft_attention_h_set, :162-197
        key = P["user_emb_matrix"][user_indices]                                 # :163
        item = key.unsqueeze(1).expand(-1, h_emb_list[0].shape[1], -1)           # :165-168
        h_emb_item = torch.cat([h_emb_list[0], item], dim=2)                     # :171-174  order [h ; user]
        h_emb_item = h_emb_item.reshape(-1, 2 * d)                               # :179
        probs = (h_emb_item @ P["h_emb_item_mlp_matrix"]).squeeze(-1) + P["h_emb_item_mlp_bias"]   # :182
        probs = probs.reshape(-1, h_emb_list[0].shape[1])                        # :186
        probs_normalized = torch.softmax(probs, dim=-1)                          # :189
        user_h_set = (h_emb_list[0] * probs_normalized.unsqueeze(2)).sum(dim=1)  # :192-195
        o_list.append(user_h_set)
    for hop in range(cfg.p_hop):                                                 # :210
        h_expanded = h_emb_list[hop].unsqueeze(3)                                # :211  [B,m,d,1]
        Rh = torch.matmul(r_emb_list[hop], h_expanded).squeeze(3)                # :214  R.h (not R^T.h)
        v = item_embeddings.unsqueeze(2)                                         # :217  [B,d,1]
        probs = torch.matmul(Rh, v).squeeze(2)                                   # :220  [B,m]
        probs_normalized = torch.softmax(probs, dim=-1)                          # :223
        o = (t_emb_list[hop] * probs_normalized.unsqueeze(2)).sum(dim=1)         # :226-229
        o_list.append(o)
    o_cat = torch.cat(o_list, dim=-1)                                            # :232
    width = d * (cfg.p_hop + 1) if cfg.PS_O_ft else d * cfg.p_hop                # :233-236
    user_o = o_cat.reshape(-1, width) @ P["user_mlp_matrix"] + P["user_mlp_bias"]
    return user_o, [user_o]                                                      # :238-240


# ----------------------------------------------------------------------------------------------------------
# A6  aggregate_delta_whole  (model.py:259-324)
# ----------------------------------------------------------------------------------------------------------
def aggregate_delta_whole(P, cfg: OracleConfig, entities: Sequence[torch.Tensor],
                          relations: Sequence[torch.Tensor], transfer_o: List[torch.Tensor]):
    B, d, K, H, M = cfg.batch_size, cfg.dim, cfg.neighbor_sample_size, cfg.h_hop, cfg.n_mix_hop
    user_query = transfer_o[0]                                                   # :261 (captured before :273)
    entity_vectors = [P["entity_emb_matrix"][i] for i in entities]               # :267
    relation_vectors = [P["relation_emb_matrix"][i] for i in relations]          # :268
    if cfg.User_orient:                                                          # :270-283
        t_o = transfer_o[0].unsqueeze(1)                                         # :273  [B,1,d]
        for e_i in range(len(entity_vectors)):
            n_entities = entity_vectors[e_i] + t_o                               # :277 (tile == broadcast)
            n_entities = n_entities.reshape(-1, d) @ P[f"transfer_agg_matrix_{e_i}"] \
                + P[f"transfer_agg_bias_{e_i}"]                                  # :279 (no activation)
            entity_vectors[e_i] = n_entities.reshape(B, entity_vectors[e_i].shape[1], d)   # :281
    importance_list: List[Optional[torch.Tensor]] = []
    for n in range(M):                                                           # :286
        mix_hop_tmp = [entity_vectors]                                           # :287-288
        for i in range(H):                                                       # :289
            name = f"agg_{i}_{n}"                                                # :290
            nxt = []
            if i == 0:
                importance_list = []                                             # :294
            for hop in range(H * M - (H * n + i)):                               # :295
                shape = (B, entity_vectors[hop].shape[1], K, d)                  # :296-297
                vector, probs = sum_aggregator_urh(
                    P, name, cfg,
                    self_vectors=entity_vectors[hop],
                    neighbor_vectors=entity_vectors[hop + 1].reshape(shape),     # :300
                    neighbor_relations=relation_vectors[hop].reshape(shape),     # :301  raw relation emb, never updated
                    user_embeddings=user_query)                                  # :302
                if i == 0:
                    importance_list.append(probs)                                # :304
                nxt.append(vector)
            entity_vectors = nxt                                                 # :306
            mix_hop_tmp.append(entity_vectors)                                   # :307
        entity_vectors = []                                                      # :309
        for mip_hop in zip(*mix_hop_tmp):                                        # :310  zip truncates to shortest
            mip = torch.cat(mip_hop, dim=-1)                                     # :311  order [ev^0_j, ..., ev^H_j]
            mip = mip.reshape(-1, d * (H + 1)) @ P[f"enti_transfer_matrix_{n}"] \
                + P[f"enti_transfer_bias_{n}"]                                   # :312
            entity_vectors.append(mip.reshape(B, -1, d))                         # :313-314
            if len(entity_vectors) == (M - (n + 1)) * H + 1:                     # :315
                break
    res = entity_vectors[0].reshape(B, d)                                        # :317
    return res, importance_list


# ----------------------------------------------------------------------------------------------------------
# forward wiring (model.py:125-159) + loss (model.py:378-412)
# ----------------------------------------------------------------------------------------------------------
@dataclass
class OracleOutput:
    scores: torch.Tensor
    scores_normalized: torch.Tensor
    user_o: torch.Tensor
    item_embeddings: torch.Tensor
    entities: List[np.ndarray]
    relations: List[np.ndarray]
    importance_list: List[Optional[torch.Tensor]]
    base_loss: Optional[torch.Tensor] = None
    l2_loss: Optional[torch.Tensor] = None
    l2_agg_loss: Optional[torch.Tensor] = None
    loss: Optional[torch.Tensor] = None
    extras: dict = field(default_factory=dict)


def _l2(t: torch.Tensor) -> torch.Tensor:
    """tf.nn.l2_loss = sum(t**2) / 2."""
    return (t * t).sum() / 2


def forward(P: Dict[str, torch.Tensor], cfg: OracleConfig, adj_entity: np.ndarray, adj_relation: np.ndarray,
            user_indices, item_indices, memories_h, memories_r, memories_t, labels=None) -> OracleOutput:
    """Feed contract = model.py:49-64: user/item int64 [B], labels f32 [B], memories_{h,r,t}[hop] int32 [B,m]
    for hop < max(1, p_hop)."""
    if not cfg.wide_deep:
        raise NotImplementedError("wide_deep=False selects MVIN.aggregate (model.py:327-376), which is broken "
                                  "in the reference (aggregator returns a tuple, aggregators.py:116)")
    as_idx = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.long)
    user_indices, item_indices = as_idx(user_indices), as_idx(item_indices)
    n_mem_hops = max(1, cfg.p_hop)
    mem_h = [as_idx(memories_h[i]) for i in range(n_mem_hops)]
    mem_r = [as_idx(memories_r[i]) for i in range(n_mem_hops)]
    mem_t = [as_idx(memories_t[i]) for i in range(n_mem_hops)]

    h_emb_list = [P["entity_emb_matrix"][mem_h[i]] for i in range(n_mem_hops)]          # :130
    r_emb_list = [P["relation_emb_KGE_matrix"][mem_r[i]] for i in range(n_mem_hops)]    # :132  [B,m,d,d]
    t_emb_list = [P["entity_emb_matrix"][mem_t[i]] for i in range(n_mem_hops)]          # :134

    ents_np, rels_np = get_neighbors(adj_entity, adj_relation, item_indices.numpy(), cfg.L, cfg.batch_size)   # :137
    entities = [torch.as_tensor(e) for e in ents_np]
    relations = [torch.as_tensor(r) for r in rels_np]

    importance: List[Optional[torch.Tensor]] = []
    if cfg.PS_only:                                                              # :142-144
        user_o, transfer_o = key_addressing(P, cfg, user_indices, item_indices, h_emb_list, r_emb_list, t_emb_list)
        item_embeddings = P["entity_emb_matrix"][item_indices]
    elif cfg.HO_only:                                                            # :146-150
        user_o = P["user_emb_matrix"][user_indices]
        if cfg.User_orient_kg_eh:
            _, transfer_o = key_addressing(P, cfg, user_indices, item_indices, h_emb_list, r_emb_list, t_emb_list)
        else:
            transfer_o = [user_o]
        item_embeddings, importance = aggregate_delta_whole(P, cfg, entities, relations, transfer_o)
    else:                                                                        # :152-156
        user_o, transfer_o = key_addressing(P, cfg, user_indices, item_indices, h_emb_list, r_emb_list, t_emb_list)
        if not cfg.User_orient_kg_eh:
            transfer_o = [P["user_emb_matrix"][user_indices]]
        item_embeddings, importance = aggregate_delta_whole(P, cfg, entities, relations, transfer_o)

    scores = (user_o * item_embeddings).sum(dim=1)                               # :158
    out = OracleOutput(scores=scores, scores_normalized=torch.sigmoid(scores), user_o=user_o,
                       item_embeddings=item_embeddings, entities=ents_np, relations=rels_np,
                       importance_list=importance)
    if labels is None:
        return out

    labels_t = torch.as_tensor(np.asarray(labels), dtype=scores.dtype)
    # :379-380  sigmoid_cross_entropy_with_logits = max(x,0) - x*z + log(1+exp(-|x|))
    out.base_loss = torch.nn.functional.binary_cross_entropy_with_logits(scores, labels_t, reduction="mean")
    l2 = scores.new_zeros(())
    for hop in range(cfg.p_hop):                                                 # :383-386  un-normalised sums, no 1/2
        l2 = l2 + (h_emb_list[hop] * h_emb_list[hop]).sum()
        l2 = l2 + (t_emb_list[hop] * t_emb_list[hop]).sum()
        l2 = l2 + (r_emb_list[hop] * r_emb_list[hop]).sum()
    l2 = l2 + _l2(P["relation_emb_matrix"])                                      # :388
    l2_agg = _l2(P["user_emb_matrix"])                                           # :392
    if not cfg.PS_only:                                                          # :393-396
        for n in range(cfg.n_mix_hop):
            for i in range(cfg.h_hop):
                l2_agg = l2_agg + _l2(P[f"agg_{i}_{n}_weights"]) + _l2(P[f"agg_{i}_{n}_urh_weights"])
    for n in range(cfg.n_mix_hop):                                               # :400-401
        l2_agg = l2_agg + _l2(P[f"enti_transfer_matrix_{n}"]) + _l2(P[f"enti_transfer_bias_{n}"])
    if cfg.p_hop > 0:                                                            # :403-408
        l2 = l2 + _l2(P["user_mlp_matrix"]) + _l2(P["user_mlp_bias"])
        last = cfg.L                                                             # self.transform_matrix = last created (:109-116)
        l2 = l2 + _l2(P[f"transfer_agg_matrix_{last}"]) + _l2(P[f"transfer_agg_bias_{last}"])   # :405
        for n in range(cfg.h_hop + 1):                                           # :407-408 (range(h_hop+1), not L+1)
            l2 = l2 + _l2(P[f"transfer_agg_matrix_{n}"]) + _l2(P[f"transfer_agg_bias_{n}"])
    l2 = l2 + _l2(P["h_emb_item_mlp_matrix"]) + _l2(P["h_emb_item_mlp_bias"])    # :410
    out.l2_loss, out.l2_agg_loss = l2, l2_agg
    out.loss = out.base_loss + cfg.l2_weight * l2 + cfg.l2_agg_weight * l2_agg   # :412
    return out


def loss_and_grads(P: Dict[str, torch.Tensor], cfg: OracleConfig, adj_entity, adj_relation, user_indices,
                   item_indices, memories_h, memories_r, memories_t, labels):
    """Forward + loss + d loss / d every parameter (dense, zero where a parameter is unused), via autograd --
    the analogue of TF's autodiff behind model.py:414."""
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    out = forward(Pg, cfg, adj_entity, adj_relation, user_indices, item_indices, memories_h, memories_r,
                  memories_t, labels)
    names = list(Pg.keys())
    grads = torch.autograd.grad(out.loss, [Pg[k] for k in names], allow_unused=True)
    gdict = {k: (g if g is not None else torch.zeros_like(Pg[k])) for k, g in zip(names, grads)}
    return out, gdict


# ----------------------------------------------------------------------------------------------------------
# Adam with TF1 semantics (model.py:414; tf.train.AdamOptimizer defaults beta1=.9, beta2=.999, eps=1e-8)
# ----------------------------------------------------------------------------------------------------------
def adam_step_tf1(P, grads, state, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
    var -= lr_t * m / (sqrt(v) + eps).  TF1's sparse path (IndexedSlices from the embedding gathers)
    decays m and v on EVERY row and scatter-adds the de-duplicated gradient, i.e. it equals this dense update
    with a zero-filled gradient."""
    t = state.get("t", 0) + 1
    state["t"] = t
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    for k in P:
        g = grads[k]
        m = state.setdefault("m_" + k, torch.zeros_like(P[k]))
        v = state.setdefault("v_" + k, torch.zeros_like(P[k]))
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        P[k] = P[k] - lr_t * m / (v.sqrt() + eps)
    return P


# ----------------------------------------------------------------------------------------------------------
# A11  feed assembly (train.py:112-122)
# ----------------------------------------------------------------------------------------------------------
def build_feed(cfg: OracleConfig, data: np.ndarray, user_triplet_set, start: int, end: int):
    """data int64 [N,3] (user, item, label); user_triplet_set[user] int32 [max(1,p), 3, m]
    (data_loader_user_set.py:402).  Returns (user, item, labels, mem_h, mem_r, mem_t) with mem_x a list of
    [B, m] int32 arrays, one per hop."""
    users = data[start:end, 0]
    mem_h, mem_r, mem_t = [], [], []
    for i in range(max(1, cfg.p_hop)):
        mem_h.append(np.asarray([user_triplet_set[u][i][0] for u in users], dtype=np.int32))
        mem_r.append(np.asarray([user_triplet_set[u][i][1] for u in users], dtype=np.int32))
        mem_t.append(np.asarray([user_triplet_set[u][i][2] for u in users], dtype=np.int32))
    return users, data[start:end, 1], data[start:end, 2].astype(np.float32), mem_h, mem_r, mem_t
