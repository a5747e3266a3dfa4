"""Generate golden vectors by executing the UNMODIFIED reference model under the TF1 shim.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py [--out DIR] [case ...]

It imports ``/root/reference/src/model/MVIN/model.py`` (and its ``aggregators.py``) with
``tests/golden/tf1_shim.py`` registered as ``tensorflow``, overwrites the model's variables with seeded values,
feeds seeded inputs through ``MVIN.get_scores / eval_case_study / train`` (model.py:416-444), and stores
inputs, parameters and outputs in ``tests/golden/<case>.npz``.  The oracle and the CUDA path are tested against
these files.  Nothing here is imported by the product.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src/model/MVIN"
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import tf1_shim  # noqa: E402

from oracle.mvin_oracle import OracleConfig, init_params  # noqa: E402  (only for seeded parameter values)

N_USER, N_ENTITY, N_REL, N_ITEM = 12, 40, 5, 10

CASES = {
    # name: (cfg overrides, regime)
    "h1_m1_p2": (dict(h_hop=1, n_mix_hop=1, p_hop=2), "trained"),
    "h2_m1_p2": (dict(h_hop=2, n_mix_hop=1, p_hop=2), "trained"),
    "h2_m1_p2_xavier": (dict(h_hop=2, n_mix_hop=1, p_hop=2), "xavier"),
    "h3_m1_p1": (dict(h_hop=3, n_mix_hop=1, p_hop=1), "trained"),
    "h2_m2_p2": (dict(h_hop=2, n_mix_hop=2, p_hop=2), "trained"),
    "h1_m2_p1": (dict(h_hop=1, n_mix_hop=2, p_hop=1), "trained"),
    "h2_m1_p2_no_kg_eh_uo": (dict(h_hop=2, n_mix_hop=1, p_hop=2, User_orient_kg_eh=0), "trained"),
    "h2_m1_p2_no_uo": (dict(h_hop=2, n_mix_hop=1, p_hop=2, User_orient=0, User_orient_kg_eh=0), "trained"),
    "h2_m1_p2_no_uor": (dict(h_hop=2, n_mix_hop=1, p_hop=2, User_orient_rela=0), "trained"),
    "h2_m1_p2_no_ps_o_ft": (dict(h_hop=2, n_mix_hop=1, p_hop=2, PS_O_ft=0), "trained"),
    "h2_m1_p2_ps_only": (dict(h_hop=2, n_mix_hop=1, p_hop=2, PS_only=1), "trained"),
    "h2_m1_p2_ho_only": (dict(h_hop=2, n_mix_hop=1, p_hop=2, HO_only=1, User_orient_kg_eh=0), "trained"),
    "h2_m1_p2_ho_only_kg_eh": (dict(h_hop=2, n_mix_hop=1, p_hop=2, HO_only=1, User_orient_kg_eh=1), "trained"),
    # combinations parameter_ablation.py defines, and deeper mixing: they pin the ORACLE for the settings the CUDA path
    # is checked against it (tests/test_variants_gpu.py::test_ablation_settings_vs_oracle, ::test_mix_blocks_vs_oracle)
    "h1_m3_p1": (dict(h_hop=1, n_mix_hop=3, p_hop=1), "trained"),
    "h2_m1_p2_no_uo_ho_only": (dict(h_hop=2, n_mix_hop=1, p_hop=2, User_orient=0, User_orient_kg_eh=0, HO_only=1), "trained"),
    "h2_m1_p2_no_uor_ho_only": (dict(h_hop=2, n_mix_hop=1, p_hop=2, User_orient_rela=0, User_orient_kg_eh=0, HO_only=1),
                                "trained"),
    "h2_m1_p2_no_uor_no_kg_eh_uo": (dict(h_hop=2, n_mix_hop=1, p_hop=2, User_orient_rela=0, User_orient_kg_eh=0), "trained"),
    "h1_m2_p2_no_uor_no_ps_o_ft": (dict(h_hop=1, n_mix_hop=2, p_hop=2, User_orient_rela=0, PS_O_ft=0), "trained"),
}


def make_args(over):
    a = dict(dataset="synthetic", load_pretrain_emb=False, h_hop=2, batch_size=6, neighbor_sample_size=4,
             p_hop=2, dim=8, l2_weight=1e-2, l2_agg_weight=1e-3, kge_weight=1e-2, lr=5e-3,
             save_model_name="g", n_mix_hop=1, n_memory=6, update_item_emb="transform_matrix",
             h0_att="st_att_h_set", path=None, User_orient=1, User_orient_rela=1, User_orient_kg_eh=1,
             PS_O_ft=1, wide_deep=1, PS_only=0, HO_only=0)
    a.update(over)
    return types.SimpleNamespace(**a)


def oracle_name_to_var(model, args):
    H, M = args.h_hop, args.n_mix_hop
    mp = {"user_emb_matrix": model.user_emb_matrix, "entity_emb_matrix": model.entity_emb_matrix,
          "relation_emb_matrix": model.relation_emb_matrix,
          "relation_emb_KGE_matrix": model.relation_emb_KGE_matrix,
          "user_mlp_matrix": model.user_mlp_matrix, "user_mlp_bias": model.user_mlp_bias,
          "h_emb_item_mlp_matrix": model.h_emb_item_mlp_matrix, "h_emb_item_mlp_bias": model.h_emb_item_mlp_bias}
    for n in range(M):
        mp[f"enti_transfer_matrix_{n}"] = model.enti_transfer_matrix_list[n]
        mp[f"enti_transfer_bias_{n}"] = model.enti_transfer_bias_list[n]
    for e in range(M * H + 1):
        mp[f"transfer_agg_matrix_{e}"] = model.transfer_matrix_list[e]
        mp[f"transfer_agg_bias_{e}"] = model.transfer_matrix_bias[e]
    if hasattr(model, "aggregators"):
        for n in range(M):
            for i in range(H):
                ag = model.aggregators[n * H + i]
                mp[f"agg_{i}_{n}_weights"] = ag.weights
                mp[f"agg_{i}_{n}_bias"] = ag.bias
                mp[f"agg_{i}_{n}_urh_weights"] = ag.urh_weights
                mp[f"agg_{i}_{n}_urh_bias"] = ag.urh_bias
    return mp


def run_case(name, over, regime, MVIN, out_dir=HERE):
    args = make_args(over)
    B, K, m = args.batch_size, args.neighbor_sample_size, args.n_memory
    rng = np.random.RandomState(sum(ord(c) for c in name))       # deterministic across interpreter runs
    adj_entity = rng.randint(0, N_ENTITY, size=(N_ENTITY, K)).astype(np.int64)
    adj_relation = rng.randint(0, N_REL, size=(N_ENTITY, K)).astype(np.int64)
    users = rng.randint(0, N_USER, size=B).astype(np.int64)
    items = rng.randint(0, N_ITEM, size=B).astype(np.int64)
    items[1] = items[0]                                          # duplicate item -> exercises grad accumulation
    labels = rng.randint(0, 2, size=B).astype(np.float32)
    if labels.min() == labels.max():                              # AUC needs both classes (model.py:419-426)
        labels[0] = 1.0 - labels[0]
    n_mem = max(1, args.p_hop)
    mem_h = [rng.randint(0, N_ENTITY, size=(B, m)).astype(np.int32) for _ in range(n_mem)]
    mem_r = [rng.randint(0, N_REL, size=(B, m)).astype(np.int32) for _ in range(n_mem)]
    mem_t = [rng.randint(0, N_ENTITY, size=(B, m)).astype(np.int32) for _ in range(n_mem)]

    tf1_shim.reset_default_graph()
    tf1_shim.DEFAULT_BATCH = B
    with contextlib.redirect_stdout(io.StringIO()):             # the reference prints while building the graph
        model = MVIN(args, N_USER, N_ENTITY, N_REL, adj_entity, adj_relation)

    cfg = OracleConfig.from_args(args)
    P = init_params(cfg, N_USER, N_ENTITY, N_REL, seed=7, regime=regime)
    if regime == "xavier":   # give the (zero-init) aggregator biases a value so their wiring is exercised
        for k in P:
            if k.startswith("agg_") and k.endswith("_bias"):
                P[k] = torch.full_like(P[k], 0.01)
    name2var = oracle_name_to_var(model, args)
    assert set(name2var) == set(P), (sorted(set(P) ^ set(name2var)))
    assert len(name2var) == len(tf1_shim.global_variables())
    for k, var in name2var.items():
        assert tuple(var.value.shape) == tuple(P[k].shape), (k, var.value.shape, P[k].shape)
        var.value = P[k].clone()

    feed = {model.user_indices: users, model.item_indices: items, model.labels: labels}
    for i in range(n_mem):
        feed[model.memories_h[i]] = [row for row in mem_h[i]]    # lists of int32 rows, as train.py:118-120 builds
        feed[model.memories_r[i]] = [row for row in mem_r[i]]
        feed[model.memories_t[i]] = [row for row in mem_t[i]]

    sess = tf1_shim.Session()
    out = {}
    it, sn = model.get_scores(sess, feed)                        # model.py:443-444
    out["scores_normalized"] = sn
    out["scores"], out["loss"], out["base_loss"], out["l2_loss"], out["l2_agg_loss"] = sess.run(
        [model.scores, model.loss, model.base_loss, model.l2_loss, model.l2_agg_loss], feed)
    for i, e in enumerate(sess.run(model.entities_data, feed)):
        out[f"entities_{i}"] = e
    for i, r in enumerate(sess.run(model.relations_data, feed)):
        out[f"relations_{i}"] = r
    if hasattr(model, "importance_list"):
        for i, imp in enumerate(model.importance_list):
            if imp is not None:
                out[f"importance_{i}"] = sess.run(imp, feed)
    auc, acc, f1 = model.eval(sess, feed)                        # model.py:419-426
    out["eval_auc_acc_f1"] = np.array([auc, acc, f1], dtype=np.float64)

    # two optimiser steps (model.py:416-417); the gradient of the first is recorded
    _, loss0 = model.train(sess, feed)
    assert abs(float(loss0) - float(out["loss"])) < 1e-6
    for k, var in name2var.items():
        out["grad__" + k] = model.optimizer.opt.last_grads[var.name].numpy()
    _, loss1 = model.train(sess, feed)
    out["loss_step1"] = np.float32(loss1)
    for k, var in name2var.items():
        out["after2__" + k] = var.value.detach().numpy().copy()

    save = dict(out)
    save["cfg_json"] = np.array(json.dumps(vars(args)))
    save["regime"] = np.array(regime)
    save["adj_entity"], save["adj_relation"] = adj_entity, adj_relation
    save["users"], save["items"], save["labels"] = users, items, labels
    for i in range(n_mem):
        save[f"mem_h_{i}"], save[f"mem_r_{i}"], save[f"mem_t_{i}"] = mem_h[i], mem_r[i], mem_t[i]
    for k in P:
        save["param__" + k] = P[k].numpy()
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **save)
    print(f"{name:28s} loss={float(out['loss']):.6f} scores[:3]={out['scores'][:3]}")


def main(argv=None):
    """python tests/golden/make_golden.py [--out DIR] [case ...]   (default: every case, into tests/golden/)"""
    argv = list(sys.argv[1:] if argv is None else argv)
    out_dir = HERE
    if argv[:1] == ["--out"]:
        out_dir, argv = argv[1], argv[2:]
    names = argv or list(CASES)
    tf1_shim.install()
    sys.path.insert(0, REF)
    import importlib
    model_mod = importlib.import_module("model")                 # the reference's model.py, unmodified
    assert os.path.realpath(model_mod.__file__).startswith("/root/reference/"), model_mod.__file__
    for name in names:
        over, regime = CASES[name]
        run_case(name, over, regime, model_mod.MVIN, out_dir)


if __name__ == "__main__":
    main()
