"""CPU twin of the n_mix_hop > 1 orchestration of mvin_b200/csrc/steps.cuh (forward_mix_impl / backward_mix_impl).

Every kernel launch of the CUDA path is one function here with the kernel's contract (level.cuh: agg_fwd_kernel,
agg_bwd_kernel, transform_fwd_kernel, transform_bwd_kernel; gemm.cuh for the mix layers; misc.cuh: sum_rows_kernel),
written with explicit backward formulas -- no autograd -- and the host loops, buffer names and index formulas are the
ones of the C++ (MixGeom, mix_in, mix_slot, the gradient sources of sum_grads).  tests/test_mix_twin.py checks it
against the oracle's autograd, so the buffer routing of the CUDA orchestration is validated without a GPU.
Test infrastructure only."""
import numpy as np
import torch


class MixGeom:
    def __init__(self, Hm, M):
        self.Hm, self.M, self.Lt = Hm, M, Hm * M

    def last_level(self, g):            # aggregator g maps levels 0 .. Lt-g-1
        return self.Lt - g - 1

    def mix_levels(self, n):            # mix block n maps levels 0 .. this
        return self.Lt - (n + 1) * self.Hm


def mix_in(buf, q, g, lv):
    n, i = divmod(g, q.Hm)
    if i > 0:
        return buf["V"][g][lv]
    return buf["V"][0][lv] if n == 0 else buf["X"][n][lv]


def mix_slot(buf, q, n, k, lv):
    if k > 0:
        return buf["V"][n * q.Hm + k][lv]
    return buf["V"][0][lv] if n == 0 else buf["X"][n][lv]


def sum_grads(*srcs):
    v = [s for s in srcs if s is not None]
    assert v, "gradient without a producer"
    out = v[0]
    for s in v[1:]:
        out = out + s
    return out


def _effective(W, u_kg, uo):
    """User_orient = 0 (steps.cuh): identity transforms, zero biases, zero user vector."""
    if uo:
        return W, u_kg
    W = dict(W)
    d = W["E"].shape[1]
    W["Wt"] = torch.eye(d, dtype=W["E"].dtype).expand(W["Wt"].shape[0], d, d).clone()
    W["bt"] = torch.zeros_like(W["bt"])
    return W, torch.zeros_like(u_kg)


def kg_forward(W, q, ents, adj_entity, adj_relation, u_kg, K, uo=True, uniform=False):
    """W: dict E, Rel, Wt [Lt+1,d,d], bt [Lt+1,d], Wa [Lt,d,d], ba [Lt,d], urh [Lt,3d], mix_w [M,(Hm+1)d,d], mix_b [M,d].
    ents[lv]: int64 [B K^lv] flat.  uo = User_orient, uniform = not User_orient_rela (the UNIFORM kernels of level.cuh:
    weights 1 / K, no further 1 / K).  Returns (item, buf)."""
    W, u_kg = _effective(W, u_kg, uo)
    invK = 1.0 if uniform else 1.0 / K
    Hm, M, Lt = q.Hm, q.M, q.Lt
    d = W["E"].shape[1]
    B = ents[0].shape[0]
    pair = [torch.arange(B).repeat_interleave(K ** lv) for lv in range(Lt)]
    buf = dict(V=[[None] * Lt for _ in range(Lt + 1)], Y=[[None] * Lt for _ in range(Lt)], X=[[None] * (Lt + 1) for _ in range(M)],
               P=[[None] * Lt for _ in range(Lt)], SU=None, pair=pair)
    s = torch.stack([W["Rel"] @ W["urh"][g][d:2 * d] for g in range(Lt)])        # rel_scores_kernel
    buf["s"] = s
    for lv in range(Lt):                                                         # transform_fwd_kernel
        buf["V"][0][lv] = (W["E"][ents[lv]] + u_kg[pair[lv]]) @ W["Wt"][lv] + W["bt"][lv]
    for n in range(M):
        for i in range(Hm):
            g = n * Hm + i
            for lv in range(q.last_level(g), -1, -1):                            # agg_fwd_kernel, one level
                e = ents[lv]
                p = torch.softmax(s[g][adj_relation[e]], dim=1)                  # [rows, K]
                if uniform:
                    p = torch.full_like(p, 1.0 / K)
                leaf = g == 0 and lv == Lt - 1
                self_v = mix_in(buf, q, g, lv)
                if leaf:
                    S = (p[:, :, None] * W["E"][adj_entity[e]]).sum(1)
                    buf["SU"] = S + u_kg[pair[lv]]
                    agg = (buf["SU"] @ W["Wt"][Lt] + W["bt"][Lt]) * invK
                else:
                    child = mix_in(buf, q, g, lv + 1).view(e.shape[0], K, d)
                    agg = (p[:, :, None] * child).sum(1) * invK
                buf["P"][g][lv] = p
                buf["Y"][g][lv] = self_v + agg
                buf["V"][g + 1][lv] = torch.relu(buf["Y"][g][lv] @ W["Wa"][g] + W["ba"][g])
        for lv in range(q.mix_levels(n) + 1):                                    # mix layer n: one GEMM per slot
            out = None
            for k in range(Hm + 1):
                part = mix_slot(buf, q, n, k, lv) @ W["mix_w"][n][k * d:(k + 1) * d]
                out = part + W["mix_b"][n] if k == 0 else out + part
            if n + 1 < M:
                buf["X"][n + 1][lv] = out
            else:
                buf["item"] = out
    return buf["item"], buf


def kg_backward(W, q, ents, adj_entity, adj_relation, u_kg, K, buf, ditem, uo=True, uniform=False):
    """Returns gradients dict: E, Rel, Wt, bt, Wa, ba, urh, mix_w, mix_b, u."""
    W, u_kg = _effective(W, u_kg, uo)
    invK = 1.0 if uniform else 1.0 / K
    Hm, M, Lt = q.Hm, q.M, q.Lt
    d = W["E"].shape[1]
    B = ents[0].shape[0]
    pair = buf["pair"]
    G = {k: torch.zeros_like(v) for k, v in W.items()}
    G["u"] = torch.zeros_like(u_kg)
    ds = torch.zeros_like(buf["s"])
    DS = [[None] * Lt for _ in range(Lt)]
    DC = [[None] * (Lt + 1) for _ in range(Lt + 1)]
    DM = [[None] * (Lt + 1) for _ in range(M)]
    gx_next = [None] * (Lt + 1)
    for n in range(M - 1, -1, -1):
        for lv in range(q.mix_levels(n) + 1):                                    # mix layer n backward
            dout = gx_next[lv] if n + 1 < M else ditem
            DM[n][lv] = [dout @ W["mix_w"][n][k * d:(k + 1) * d].T for k in range(Hm + 1)]      # gemm_mix_bwd
            for k in range(Hm + 1):                                              # dw_mix
                G["mix_w"][n][k * d:(k + 1) * d] += mix_slot(buf, q, n, k, lv).T @ dout
            G["mix_b"][n] += dout.sum(0)
        for i in range(Hm - 1, -1, -1):
            g = n * Hm + i
            for lv in range(q.last_level(g) + 1):
                s_mix = DM[n][lv][i + 1] if lv <= q.mix_levels(n) else None
                s_self = DS[g + 1][lv] if (i + 1 < Hm and lv <= q.last_level(g + 1)) else None
                s_par = DC[g + 1][lv] if (i + 1 < Hm and lv >= 1) else None
                gsum = sum_grads(s_mix, s_self, s_par)
                # agg_bwd_kernel, one level
                e = ents[lv]
                p = buf["P"][g][lv]
                leaf = g == 0 and lv == Lt - 1
                gz = gsum * (buf["V"][g + 1][lv] > 0)
                G["Wa"][g] += buf["Y"][g][lv].T @ gz
                G["ba"][g] += gz.sum(0)
                gs = gz @ W["Wa"][g].T
                DS[g][lv] = gs
                grow = gs * invK
                if leaf:
                    G["Wt"][Lt] += buf["SU"].T @ grow
                    G["bt"][Lt] += grow.sum(0)
                    gsu = grow @ W["Wt"][Lt].T
                    G["u"].index_add_(0, pair[lv], gsu)
                    nb = adj_entity[e]
                    G["E"].index_add_(0, nb.reshape(-1), (p[:, :, None] * gsu[:, None, :]).reshape(-1, d))
                    dp = (gsu[:, None, :] * W["E"][nb]).sum(-1)
                else:
                    child = mix_in(buf, q, g, lv + 1).view(e.shape[0], K, d)
                    DC[g][lv + 1] = (p[:, :, None] * grow[:, None, :]).reshape(-1, d)
                    dp = (grow[:, None, :] * child).sum(-1)
                if not uniform:
                    dlogit = p * (dp - (p * dp).sum(1, keepdim=True))
                    ds[g].index_add_(0, adj_relation[e].reshape(-1), dlogit.reshape(-1))
        g0 = n * Hm
        n_in = Lt - 1 if n == 0 else Lt - g0
        for lv in range(n_in + 1):
            s_mix = DM[n][lv][0] if lv <= q.mix_levels(n) else None
            s_self = DS[g0][lv] if lv <= q.last_level(g0) else None
            s_par = DC[g0][lv] if lv >= 1 else None
            gx_next[lv] = sum_grads(s_mix, s_self, s_par)
    for lv in range(Lt):                                                         # transform_bwd_kernel
        dT = gx_next[lv]
        x = W["E"][ents[lv]] + u_kg[pair[lv]]
        G["Wt"][lv] += x.T @ dT
        G["bt"][lv] += dT.sum(0)
        dx = dT @ W["Wt"][lv].T
        G["E"].index_add_(0, ents[lv], dx)
        G["u"].index_add_(0, pair[lv], dx)
    for g in range(Lt):                                                          # rel_scores_bwd_kernel
        w = W["urh"][g][d:2 * d]
        G["Rel"] += ds[g][:, None] * w[None, :]
        G["urh"][g][d:2 * d] += ds[g] @ W["Rel"]
    if not uo:                                                                   # accumulated in scratch on the CUDA path
        G["Wt"].zero_(); G["bt"].zero_(); G["u"].zero_()
    return G


def pack_weights(P, cfg):
    """Oracle parameter dict -> the stacked fields of include/mvin_b200.h (model.py: _name_map of mvin_b200/model.py)."""
    Hm, M = cfg.h_hop, cfg.n_mix_hop
    Lt = Hm * M
    W = dict(E=P["entity_emb_matrix"], Rel=P["relation_emb_matrix"],
             Wt=torch.stack([P[f"transfer_agg_matrix_{e}"] for e in range(Lt + 1)]),
             bt=torch.stack([P[f"transfer_agg_bias_{e}"] for e in range(Lt + 1)]),
             Wa=torch.stack([P[f"agg_{g % Hm}_{g // Hm}_weights"] for g in range(Lt)]),
             ba=torch.stack([P[f"agg_{g % Hm}_{g // Hm}_bias"] for g in range(Lt)]),
             urh=torch.stack([P[f"agg_{g % Hm}_{g // Hm}_urh_weights"].reshape(-1) for g in range(Lt)]),
             mix_w=torch.stack([P[f"enti_transfer_matrix_{n}"] for n in range(M)]),
             mix_b=torch.stack([P[f"enti_transfer_bias_{n}"] for n in range(M)]))
    return {k: v.detach().clone().double() for k, v in W.items()}


def unpack_grads(G, cfg):
    """Back to the oracle's names."""
    Hm, M = cfg.h_hop, cfg.n_mix_hop
    Lt = Hm * M
    out = {"entity_emb_matrix": G["E"], "relation_emb_matrix": G["Rel"]}
    for e in range(Lt + 1):
        out[f"transfer_agg_matrix_{e}"] = G["Wt"][e]
        out[f"transfer_agg_bias_{e}"] = G["bt"][e]
    for g in range(Lt):
        i, n = g % Hm, g // Hm
        out[f"agg_{i}_{n}_weights"] = G["Wa"][g]
        out[f"agg_{i}_{n}_bias"] = G["ba"][g]
        out[f"agg_{i}_{n}_urh_weights"] = G["urh"][g].reshape(-1, 1)
    for n in range(M):
        out[f"enti_transfer_matrix_{n}"] = G["mix_w"][n]
        out[f"enti_transfer_bias_{n}"] = G["mix_b"][n]
    return out
