"""Stage-by-stage CPU model of the FUSED decomposition the CUDA kernels implement.  TEST INFRASTRUCTURE ONLY.

The oracle (oracle/mvin_oracle.py) restates the reference's TF graph op for op.  The CUDA path computes the
same function through a different factorisation:

  * attention over the K neighbours depends on the relation id only:  p = softmax_k(s_i[rel_k]),
    s_i = relation_emb @ urh_weights_i[d:2d]          (the user / self thirds are constant along K and cancel)
  * the leaf-level user-oriented transform is hoisted out of the weighted mean (sum_k p_k = 1)
  * ripple logits  v^T R h  are evaluated as  (R^T v) . h  with Q[b,r] = RK[r]^T v_b computed once per pair
  * the user half of the h-set logit cancels in the softmax

and a hand-derived backward.  This file spells that factorisation out in torch fp64/fp32 with MANUAL gradients
(no autograd), so (a) the derivation is checked against the oracle's autograd on CPU, and (b) every
intermediate buffer of the CUDA workspace has a named CPU counterpart for debugging.
Supported: --ablation all, n_mix_hop = 1 (what include/mvin_b200.h supports).
"""
from __future__ import annotations

import numpy as np
import torch


def _softmax(x):
    return torch.softmax(x, dim=-1)


def fused_forward_backward(P, cfg, adj_entity, adj_relation, users, items, mem_h, mem_r, mem_t, labels=None,
                           hoist_all=False, table=False):
    """table: the ENTITY-TABLE form of aggregator iteration 0 (table.cuh), one step beyond hoist_all.  With the hoist at
    every level, the pre-activation of iteration 0 at a level-h node with entity e of pair b is
        pre = (E[e] + u_b) M1_h + (S_e + u_b) M2_h + c_h = A_h[e] + C_h[b]
        M1_h = W_t[h] W_a0,  M2_h = W_t[h+1] W_a0 / K,  c_h = (b_t[h] + b_t[h+1] / K) W_a0 + b_a0
    i.e. a per-ENTITY table A_h = E M1_h + Se M2_h + c_h plus a per-PAIR vector C_h = u (M1_h + M2_h): iteration 0 has no
    per-row dense map at all, V[1][h][row] = relu(A_h[ent] + C_h[pair]) is a table gather, and the deepest level is never
    materialised (iteration 1 gathers it from A_{H-1} through the adjacency record).  Backward: the pre-activation
    gradients are summed per entity (dA_h) and per pair (dCs_h); everything below them is dense algebra on
    [n_entity, d] / [B, d] matrices and a d x d parameter chain.

    hoist_all: the ROW-LOCAL form of aggregator iteration 0 (level_tcb.cuh): the hoist that removes the leaf rows
    (sum_k p_k = 1, attention relation-only) holds at EVERY level of iteration 0, because the children of a level-h node
    with entity e are exactly adj[e] and their transformed rows are (E[n_k] + u) W_t[h+1] + b_t[h+1]:
        agg_0[h] = ((S_e + u) W_t[h+1] + b_t[h+1]) / K,   S_e = sum_k p_k(e) E[adj[e][k]]   (one vector per ENTITY)
    so iteration 0 needs no parent-child traffic at all, and its backward sends the children's share through the
    per-entity gradient GSe[e] (then dE[adj[e][k]] += p_k GSe[e], dp_k = GSe[e] . E[adj[e][k]], once per entity)."""
    d, K, H, p, m = cfg.dim, cfg.neighbor_sample_size, cfg.h_hop, cfg.p_hop, cfg.n_memory
    assert cfg.n_mix_hop == 1
    L = H
    items = torch.as_tensor(np.asarray(items), dtype=torch.long)
    B = items.shape[0]
    adjE = torch.as_tensor(adj_entity, dtype=torch.long)
    adjR = torch.as_tensor(adj_relation, dtype=torch.long)
    E, Rel, RK = P["entity_emb_matrix"], P["relation_emb_matrix"], P["relation_emb_KGE_matrix"]
    dt = E.dtype
    mh = [torch.as_tensor(np.asarray(x), dtype=torch.long) for x in mem_h]
    mr = [torch.as_tensor(np.asarray(x), dtype=torch.long) for x in mem_r]
    mt = [torch.as_tensor(np.asarray(x), dtype=torch.long) for x in mem_t]
    W = {}  # workspace: every named buffer below has a CUDA counterpart

    # ---------------- get_neighbors ----------------
    ent = [items.reshape(B, 1)]
    rel = []
    for h in range(L):
        ent.append(adjE[ent[h]].reshape(B, -1))
        rel.append(adjR[ent[h]].reshape(B, -1))

    # ---------------- ripple / key addressing ----------------
    v = E[items]                                              # Vbuf [B,d]
    Q = torch.einsum("bi,rij->brj", v, RK)                    # [B,n_rel,d]
    w_hi = P["h_emb_item_mlp_matrix"][:d, 0]
    h0 = E[mh[0]]
    a = _softmax(h0 @ w_hi)                                   # h-set probs [B,m]
    o_list = [(a.unsqueeze(-1) * h0).sum(1)]
    probs = []
    for hop in range(p):
        hrow = E[mh[hop]]
        key = torch.gather(Q, 1, mr[hop].unsqueeze(-1).expand(B, m, d))
        pr = _softmax((key * hrow).sum(-1))
        probs.append(pr)
        o_list.append((pr.unsqueeze(-1) * E[mt[hop]]).sum(1))
    O = torch.cat(o_list, -1)
    u = O @ P["user_mlp_matrix"] + P["user_mlp_bias"]         # user_o [B,d]
    W.update(Q=Q, hset_probs=a, probs=probs, O=O, u=u)

    # ---------------- KG side ----------------
    s = [Rel @ P[f"agg_{i}_0_urh_weights"][d:2 * d, 0] for i in range(H)]       # [H][n_rel]
    Wt = [P[f"transfer_agg_matrix_{e}"] for e in range(L + 1)]
    bt = [P[f"transfer_agg_bias_{e}"] for e in range(L + 1)]
    Wa = [P[f"agg_{i}_0_weights"] for i in range(H)]
    ba = [P[f"agg_{i}_0_bias"] for i in range(H)]

    def att(i, h):
        return _softmax(s[i][rel[h]].reshape(B, K ** h, K))

    XU = [E[ent[h]] + u.unsqueeze(1) for h in range(L)]       # [B,K^h,d]
    T = [XU[h] @ Wt[h] + bt[h] for h in range(L)]
    # V[i][h]: i = 0 is T; Y[i][h] = self + agg of aggregator i at hop h
    V = [dict() for _ in range(H + 1)]
    Y = [dict() for _ in range(H)]
    for h in range(L):
        V[0][h] = T[h]
    p0 = att(0, L - 1)
    S = (p0.unsqueeze(-1) * E[ent[L]].reshape(B, K ** (L - 1), K, d)).sum(2)
    SU = S + u.unsqueeze(1)
    Z = SU @ Wt[L] + bt[L]
    if table:
        hoist_all = False
        p_ent = _softmax(s[0][adjR])
        Se = (p_ent.unsqueeze(-1) * E[adjE]).sum(1)
        M1 = [Wt[h] @ Wa[0] for h in range(H)]
        M2 = [Wt[h + 1] @ Wa[0] / K for h in range(H)]
        cst = [(bt[h] + bt[h + 1] / K) @ Wa[0] + ba[0] for h in range(H)]
        A = [E @ M1[h] + Se @ M2[h] + cst[h] for h in range(H)]                # per-entity tables [n_entity, d]
        Cp = [u @ (M1[h] + M2[h]) for h in range(H)]                           # per-pair [B, d]
    if hoist_all:
        p_ent = _softmax(s[0][adjR])                          # [n_entity, K] attention of aggregator 0 per ENTITY
        Se = (p_ent.unsqueeze(-1) * E[adjE]).sum(1)           # [n_entity, d]
        SUh = [Se[ent[h]] + u.unsqueeze(1) for h in range(L)]
    for i in range(H):
        for h in range(L - i):
            if i == 0 and table:
                V[1][h] = torch.relu(A[h][ent[h]] + Cp[h].unsqueeze(1))
                continue
            if i == 0 and hoist_all:
                agg = (SUh[h] @ Wt[h + 1] + bt[h + 1]) / K
            elif i == 0 and h == L - 1:
                agg = Z / K
            else:
                agg = (att(i, h).unsqueeze(-1) * V[i][h + 1].reshape(B, K ** h, K, d)).sum(2) / K
            Y[i][h] = V[i][h] + agg
            V[i + 1][h] = torch.relu(Y[i][h] @ Wa[i] + ba[i])
    cat = torch.cat([V[i][0].reshape(B, d) for i in range(H + 1)], -1)
    item = cat @ P["enti_transfer_matrix_0"] + P["enti_transfer_bias_0"]
    scores = (u * item).sum(-1)
    W.update(ent=ent, rel=rel, s=s, XU=XU, T=T, SU=SU, Y=Y, V=V, item=item, scores=scores,
             imp=[att(0, h) for h in range(min(2, L))])
    if labels is None:
        return W, None

    # ---------------- loss ----------------
    z = torch.as_tensor(np.asarray(labels), dtype=dt)
    base = (torch.clamp(scores, min=0) - scores * z + torch.log1p(torch.exp(-scores.abs()))).mean()
    l2w, l2a = cfg.l2_weight, cfg.l2_agg_weight
    hs = lambda t: (t * t).sum() / 2
    l2 = sum((E[mh[k]] ** 2).sum() + (E[mt[k]] ** 2).sum() + (RK[mr[k]] ** 2).sum() for k in range(p))
    l2 = l2 + hs(Rel)
    dense_l2 = {"user_mlp_matrix": 1, "user_mlp_bias": 1, "h_emb_item_mlp_matrix": 1, "h_emb_item_mlp_bias": 1}
    if p > 0:
        for e in list(range(H + 1)) + [L]:                    # model.py:405-408: index L counted twice
            dense_l2[f"transfer_agg_matrix_{e}"] = dense_l2.get(f"transfer_agg_matrix_{e}", 0) + 1
            dense_l2[f"transfer_agg_bias_{e}"] = dense_l2.get(f"transfer_agg_bias_{e}", 0) + 1
    else:
        dense_l2.pop("user_mlp_matrix"), dense_l2.pop("user_mlp_bias")
    dense_agg = {"user_emb_matrix": 1, "enti_transfer_matrix_0": 1, "enti_transfer_bias_0": 1}
    for i in range(H):
        dense_agg[f"agg_{i}_0_weights"] = 1
        dense_agg[f"agg_{i}_0_urh_weights"] = 1
    l2 = l2 + sum(c * hs(P[k]) for k, c in dense_l2.items())
    l2agg = sum(c * hs(P[k]) for k, c in dense_agg.items())
    loss = base + l2w * l2 + l2a * l2agg
    W.update(loss=loss, base_loss=base, l2_loss=l2, l2_agg_loss=l2agg)

    # ---------------- backward (manual) ----------------
    G = {k: torch.zeros_like(t) for k, t in P.items()}
    for k, c in dense_l2.items():
        G[k] += l2w * c * P[k]
    for k, c in dense_agg.items():
        G[k] += l2a * c * P[k]
    G["relation_emb_matrix"] += l2w * Rel
    dE = G["entity_emb_matrix"]

    gsc = (torch.sigmoid(scores) - z) / B
    ditem = gsc.unsqueeze(-1) * u
    du = gsc.unsqueeze(-1) * item
    # mix
    G["enti_transfer_matrix_0"] += cat.T @ ditem
    G["enti_transfer_bias_0"] += ditem.sum(0)
    dcat = ditem @ P["enti_transfer_matrix_0"].T
    dV = [dict() for _ in range(H + 1)]
    for i in range(H + 1):
        dV[i][0] = dcat[:, i * d:(i + 1) * d].reshape(B, 1, d).clone()
    ds = [torch.zeros_like(s[i]) for i in range(H)]
    for i in reversed(range(H)):
        if i == 0 and table:
            break
        for h in range(L - i):                                # ascending hop: child grads land before self grads
            gout = dV[i + 1][h]
            gz = gout * (V[i + 1][h] > 0).to(dt)
            G[f"agg_{i}_0_weights"] += Y[i][h].reshape(-1, d).T @ gz.reshape(-1, d)
            G[f"agg_{i}_0_bias"] += gz.reshape(-1, d).sum(0)
            gs = gz @ Wa[i].T
            dV[i][h] = dV[i][h] + gs if h in dV[i] else gs.clone()
            grow = gs / K                                     # [B,K^h,d]
            pk = att(i, h)
            if i == 0 and hoist_all:
                G[f"transfer_agg_matrix_{h + 1}"] += SUh[h].reshape(-1, d).T @ grow.reshape(-1, d)
                G[f"transfer_agg_bias_{h + 1}"] += grow.reshape(-1, d).sum(0)
                gsu = grow @ Wt[h + 1].T
                du = du + gsu.sum(1)
                if h == 0:
                    GSe = torch.zeros_like(Se)
                GSe.index_put_((ent[h].reshape(-1),), gsu.reshape(-1, d), accumulate=True)
                continue                                      # no dchild, no per-row softmax gradient
            if i == 0 and h == L - 1:
                G[f"transfer_agg_matrix_{L}"] += SU.reshape(-1, d).T @ grow.reshape(-1, d)
                G[f"transfer_agg_bias_{L}"] += grow.reshape(-1, d).sum(0)
                gsu = grow @ Wt[L].T
                du = du + gsu.sum(1)
                child = E[ent[L]].reshape(B, K ** (L - 1), K, d)
                dp = (gsu.unsqueeze(2) * child).sum(-1)
                dE.index_put_((ent[L].reshape(-1),),
                              (pk.unsqueeze(-1) * gsu.unsqueeze(2)).reshape(-1, d), accumulate=True)
            else:
                child = V[i][h + 1].reshape(B, K ** h, K, d)
                dp = (grow.unsqueeze(2) * child).sum(-1)
                dchild = (pk.unsqueeze(-1) * grow.unsqueeze(2)).reshape(B, K ** (h + 1), d)
                assert (h + 1) not in dV[i]
                dV[i][h + 1] = dchild
            dlogit = pk * (dp - (pk * dp).sum(-1, keepdim=True))
            ds[i].index_put_((rel[h].reshape(-1),), dlogit.reshape(-1), accumulate=True)
    if table:
        GSe = torch.zeros_like(Se)
        dWa0, dba0 = G["agg_0_0_weights"], G["agg_0_0_bias"]
        for h in range(H):
            dpre = dV[1][h] * (V[1][h] > 0).to(dt)                             # [B, K^h, d]
            dA = torch.zeros_like(Se).index_put_((ent[h].reshape(-1),), dpre.reshape(-1, d), accumulate=True)
            dCs = dpre.sum(1)                                                  # [B, d]
            dE += dA @ M1[h].T
            GSe += dA @ M2[h].T
            dM1 = E.T @ dA + u.T @ dCs
            dM2 = Se.T @ dA + u.T @ dCs
            dc = dCs.sum(0)
            du = du + dCs @ (M1[h] + M2[h]).T
            # d x d parameter chain
            G[f"transfer_agg_matrix_{h}"] += dM1 @ Wa[0].T
            G[f"transfer_agg_matrix_{h + 1}"] += dM2 @ Wa[0].T / K
            dWa0 += Wt[h].T @ dM1 + Wt[h + 1].T @ dM2 / K + torch.outer(bt[h] + bt[h + 1] / K, dc)
            G[f"transfer_agg_bias_{h}"] += dc @ Wa[0].T
            G[f"transfer_agg_bias_{h + 1}"] += dc @ Wa[0].T / K
            dba0 += dc
    if hoist_all or table:
        # per-entity backward of S_e (leaf_entity_kernel<BWD>): every entity that occurred at some level < L
        child = E[adjE]                                       # [n_entity, K, d]
        dp = (GSe.unsqueeze(1) * child).sum(-1)
        dE.index_put_((adjE.reshape(-1),), (p_ent.unsqueeze(-1) * GSe.unsqueeze(1)).reshape(-1, d), accumulate=True)
        dlogit = p_ent * (dp - (p_ent * dp).sum(-1, keepdim=True))
        ds[0].index_put_((adjR.reshape(-1),), dlogit.reshape(-1), accumulate=True)
    for i in range(H):
        wr = P[f"agg_{i}_0_urh_weights"][d:2 * d, 0]
        G["relation_emb_matrix"] += ds[i].unsqueeze(-1) * wr
        G[f"agg_{i}_0_urh_weights"][d:2 * d, 0] += ds[i] @ Rel
    # user-oriented transform, levels 0..L-1
    for h in range(1 if table else L):
        dT = dV[0][h]
        G[f"transfer_agg_matrix_{h}"] += XU[h].reshape(-1, d).T @ dT.reshape(-1, d)
        G[f"transfer_agg_bias_{h}"] += dT.reshape(-1, d).sum(0)
        gx = dT @ Wt[h].T
        du = du + gx.sum(1)
        dE.index_put_((ent[h].reshape(-1),), gx.reshape(-1, d), accumulate=True)
    # user_o = O W_user + b
    G["user_mlp_matrix"] += O.T @ du
    G["user_mlp_bias"] += du.sum(0)
    dO = du @ P["user_mlp_matrix"].T
    # h-set
    go = dO[:, :d]
    dprob = (go.unsqueeze(1) * h0).sum(-1)
    dl = a * (dprob - (a * dprob).sum(-1, keepdim=True))
    dE.index_put_((mh[0].reshape(-1),), (a.unsqueeze(-1) * go.unsqueeze(1) + dl.unsqueeze(-1) * w_hi).reshape(-1, d),
                  accumulate=True)
    G["h_emb_item_mlp_matrix"][:d, 0] += (dl.unsqueeze(-1) * h0).sum((0, 1))
    dQ = torch.zeros_like(Q)
    cnt = torch.zeros(RK.shape[0], dtype=dt)
    for hop in range(p):
        go = dO[:, (hop + 1) * d:(hop + 2) * d]
        hrow, trow = E[mh[hop]], E[mt[hop]]
        pr = probs[hop]
        key = torch.gather(Q, 1, mr[hop].unsqueeze(-1).expand(B, m, d))
        dprob = (go.unsqueeze(1) * trow).sum(-1)
        dl = pr * (dprob - (pr * dprob).sum(-1, keepdim=True))
        dE.index_put_((mt[hop].reshape(-1),), (pr.unsqueeze(-1) * go.unsqueeze(1) + 2 * l2w * trow).reshape(-1, d),
                      accumulate=True)
        dE.index_put_((mh[hop].reshape(-1),), (dl.unsqueeze(-1) * key + 2 * l2w * hrow).reshape(-1, d),
                      accumulate=True)
        dQ.scatter_add_(1, mr[hop].unsqueeze(-1).expand(B, m, d), dl.unsqueeze(-1) * hrow)
        cnt += torch.bincount(mr[hop].reshape(-1), minlength=RK.shape[0]).to(dt)
    G["relation_emb_KGE_matrix"] += torch.einsum("bi,brj->rij", v, dQ) + 2 * l2w * cnt.reshape(-1, 1, 1) * RK
    dE.index_put_((items,), torch.einsum("brj,rij->bi", dQ, RK), accumulate=True)
    W.update(dV=dV, du=du, dO=dO, dQ=dQ, ds=ds)
    return W, G
