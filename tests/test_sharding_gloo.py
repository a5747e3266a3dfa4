"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic in mvin_b200/sharding.py:

  * the row-sharded entity table round-trips (scatter on every rank, all-gather, reassemble);
  * the data-parallel protocol of mvin_set_batch_scale is exact: with the base loss divided by the GLOBAL batch
    and the dense L2 terms divided by the world size on every rank, a SUM all-reduce of losses and gradients over
    the ranks equals the oracle's loss / gradients on the concatenated batch (model.py:378-412).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvin_b200 import sharding
from oracle import mvin_oracle as orc
from tests.synth import make_args, make_problem

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        ret[rank] = fn(rank)
    finally:
        dist.destroy_process_group()


def _run(fn):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, fn, ret)) for r in range(WORLD)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    return dict(ret)


def _table_roundtrip(rank):
    full = torch.Generator().manual_seed(7)
    full = torch.rand((101, 8), generator=full)                     # 101 rows: not divisible by the world size
    mine = sharding.scatter_table(full, WORLD, rank)
    assert mine.shape[0] == sharding.shard_rows(101, WORLD)
    for l in range(mine.shape[0]):
        e = l * WORLD + rank
        if e < 101:
            assert sharding.owner_of(e, WORLD) == rank and sharding.local_row(e, WORLD) == l
            assert torch.equal(mine[l], full[e])
    parts = [torch.empty_like(mine) for _ in range(WORLD)]
    dist.all_gather(parts, mine)
    back = sharding.gather_table(parts, 101)
    return bool(torch.equal(back, full))


def test_entity_table_shards_roundtrip():
    assert all(_run(_table_roundtrip).values())


def _dp_protocol(rank, over=()):
    args = make_args(**{**dict(batch_size=16, dim=8, neighbor_sample_size=4, h_hop=2, p_hop=2, n_memory=8), **dict(over)})
    prob = make_problem(args, n_user=20, n_entity=60, n_relation=5, n_item=12, seed=3)
    cfg, P = prob["cfg"], prob["P"]
    B = args.batch_size
    sl = sharding.split_batch(B, rank, WORLD)
    local_cfg = orc.OracleConfig(**{**cfg.__dict__, "batch_size": B // WORLD})
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    mh = [m[sl] for m in prob["mem_h"]]
    mr = [m[sl] for m in prob["mem_r"]]
    mt = [m[sl] for m in prob["mem_t"]]
    out = orc.forward(Pg, local_cfg, prob["adj_entity"], prob["adj_relation"], prob["users"][sl], prob["items"][sl],
                      mh, mr, mt, prob["labels"][sl])
    # batch-gathered L2 terms (model.py:383-386) partition over pairs; everything else in l2 / l2_agg is dense
    E, RK = Pg["entity_emb_matrix"], Pg["relation_emb_KGE_matrix"]
    idx = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.long)
    batch_l2 = sum((E[idx(mh[h])] ** 2).sum() + (E[idx(mt[h])] ** 2).sum() + (RK[idx(mr[h])] ** 2).sum()
                   for h in range(cfg.p_hop))
    dense_l2 = out.l2_loss - batch_l2
    scale = 1.0 / WORLD
    local = (out.base_loss * (B // WORLD) / B + cfg.l2_weight * (batch_l2 + scale * dense_l2)
             + cfg.l2_agg_weight * scale * out.l2_agg_loss)
    names = list(Pg)
    grads = torch.autograd.grad(local, [Pg[k] for k in names], allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(Pg[k]) for k, g in zip(names, grads)]
    total = sharding.allreduce_flat(grads, None, extra=local.detach().reshape(1))
    ref_out, ref_grads = orc.loss_and_grads(P, cfg, prob["adj_entity"], prob["adj_relation"], prob["users"],
                                            prob["items"], prob["mem_h"], prob["mem_r"], prob["mem_t"], prob["labels"])
    ok = abs(float(total[0]) - float(ref_out.loss)) < 1e-5 * max(1.0, abs(float(ref_out.loss)))
    for k, g in zip(names, grads):
        ref = ref_grads[k]
        ok = ok and float((g - ref).abs().max()) <= 1e-5 * max(float(ref.abs().max()), 1e-8) + 1e-9
    return bool(ok)


def test_data_parallel_loss_scaling_protocol_matches_single_device():
    assert all(_run(_dp_protocol).values())


@pytest.mark.parametrize("over", [(("HO_only", 1), ("User_orient_kg_eh", 0)),          # U[user] carries real gradients
                                  (("h_hop", 1), ("n_mix_hop", 2), ("User_orient_rela", 0)),
                                  (("PS_only", 1), ("PS_O_ft", 0))])
def test_data_parallel_protocol_holds_for_the_model_variants(over):
    """The same split (base loss / global batch, dense L2 / world, batch L2 per pair) is exact for the other
    parameter_ablation.py settings and for several mix blocks -- with HO_only the user table's gradient is no longer the
    dense L2 term alone, which is why MVIN.allreduce_grads then exchanges the whole bucket."""
    import functools
    assert all(_run(functools.partial(_dp_protocol, over=over)).values())


def test_split_batch_rejects_ragged():
    with pytest.raises(ValueError):
        sharding.split_batch(10, 0, 4)
    assert sharding.split_batch(8, 1, 2) == slice(4, 8)
