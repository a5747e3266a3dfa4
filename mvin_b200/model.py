"""Python face of the drop-in: the reference's `MVIN` class API on top of libmvin_b200.so.

Mirrors src/model/MVIN/model.py of the reference:
  ctor            MVIN(args, n_user, n_entity, n_relation, adj_entity, adj_relation)      model.py:7
  feed keys       model.user_indices / item_indices / labels, model.memories_{h,r,t}[hop]  model.py:49-64
  train           -> (None, loss), one Adam step                                           model.py:416-417
  eval            -> (auc, acc, f1)                                                        model.py:419-426
  eval_case_study -> 7-tuple                                                               model.py:428-441
  get_scores      -> (item_indices, scores_normalized)                                     model.py:443-444
  save_pretrain_emb_fuc                                                                    model.py:66-67

`sess` arguments are accepted and ignored (there is no TF session).  torch tensors are used only as device
buffers whose `.data_ptr()` goes through the C ABI.  There is no CPU fallback: without a CUDA device or the
built library every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib, sharding
from ._lib import Config, Params, PARAM_FIELDS, check


class _Placeholder:
    """Stand-in for tf.placeholder: only used as a feed-dict key (train.py:113-120)."""

    def __init__(self, name, dtype, shape):
        self.name, self.dtype, self.shape = name, dtype, shape

    def __repr__(self):
        return f"<placeholder {self.name} {self.dtype} {self.shape}>"


def _flag(args, name, default):
    return bool(getattr(args, name, default))


def flags_from_args(args) -> int:
    bits = [("User_orient", 1), ("User_orient_rela", 1), ("User_orient_kg_eh", 1), ("PS_O_ft", 1), ("wide_deep", 1),
            ("PS_only", 0), ("HO_only", 0)]
    f = 0
    for i, (name, dflt) in enumerate(bits):
        if _flag(args, name, dflt):
            f |= 1 << i
    return f


def pack_user_triplet_set(user_triplet_set, n_user, n_hops, n_memory) -> np.ndarray:
    """The reference's ripple sets are a defaultdict(user -> int32 [p, 3, m]) (data_loader_user_set.py:396-402; users
    without history are simply absent and the defaultdict would hand back []).  Returns the packed int32
    [n_user, p, 3, m] array the device path binds; absent users get an all-zero block (they never occur in a batch:
    every user of the rating file has history).  Arrays pass through."""
    from collections.abc import Mapping
    if isinstance(user_triplet_set, Mapping):
        out = np.zeros((n_user, n_hops, 3, n_memory), dtype=np.int32)
        for u in list(user_triplet_set.keys()):
            blk = np.asarray(user_triplet_set[u], dtype=np.int32)
            if blk.size == 0:
                continue
            if not 0 <= int(u) < n_user:
                raise ValueError(f"user id {u} outside 0..{n_user - 1}")
            out[int(u)] = blk.reshape(n_hops, 3, n_memory)
        return out
    return np.ascontiguousarray(np.asarray(user_triplet_set), dtype=np.int32)


def rows_into(dst: np.ndarray, rows) -> None:
    """train.py:118-120 hands the ripple memories over as a Python list of B int32 [m] rows.  Writes them into `dst`
    (int32 [B, m], C-contiguous, typically pinned staging).  The rows are contiguous buffers, so one bytes.join is a
    single pass of memcpys in C -- about 5x faster than np.concatenate / np.stack over B small arrays (B = 16 384:
    0.9 ms against 4.5 / 13 ms); anything else (other dtypes, strided rows, lists of lists) takes the NumPy path, which
    casts."""
    flat = dst.reshape(-1)
    first = rows[0] if len(rows) else None
    if isinstance(first, np.ndarray) and first.dtype == np.int32:
        try:
            raw = b"".join(rows)
        except (TypeError, BufferError, ValueError):
            raw = None
        if raw is not None and len(raw) == flat.nbytes and all(getattr(r, "dtype", None) == np.int32 for r in rows[:: max(1, len(rows) // 8)]):
            flat[:] = np.frombuffer(raw, dtype=np.int32)
            return
    np.concatenate([np.asarray(r).reshape(-1) for r in rows], out=flat, casting="unsafe")


class MVIN(object):
    def __init__(self, args, n_user, n_entity, n_relation, adj_entity, adj_relation, device=None, seed: int = 1,
                 entity_shards: int = 1, process_group=None, leaf_exchange=None):
        """Reference signature (model.py:7) plus keyword-only extensions:
        device / seed      where the tables live, seed of the Xavier init;
        entity_shards = G  row-shard the entity table (and its gradient / Adam state) over G shards, entity e in
                           shard e % G at row e // G (include/mvin_b200.h: mvin_bind_entity_shards);
        process_group      a torch.distributed group of G one-GPU ranks of one box: every rank owns shard `rank`
                           and maps the peers' shards through CUDA IPC (NVLink peer loads / reductions).  Without a
                           group all G shards live on this device ("virtual shards", used by the tests).
        leaf_exchange      row-sharded table only: owner-side partial reduction of the deepest level (exchange.cuh;
                           include/mvin_b200.h: mvin_xchg_*) -- the leaf rows are reduced by the rank that owns them and
                           one d-vector per (parent node, owner) crosses NVLink instead of every raw row.  Default: on
                           with a process group (env MVIN_B200_XCHG=0 turns it off), off for virtual shards.
        `adj_entity` may also be a packed device tensor int32 [n_entity, 2, K] (ids then relation ids) with
        adj_relation=None, for graphs too large for the reference's host-side int64 arrays."""
        if not torch.cuda.is_available():
            raise RuntimeError("mvin_b200.MVIN needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.n_shards = int(entity_shards)
        self.group = process_group
        if self.n_shards < 1 or (self.n_shards & (self.n_shards - 1)) or self.n_shards > 16:
            raise ValueError("entity_shards must be a power of two in 1..16")
        if self.group is not None:
            import torch.distributed as dist
            if dist.get_world_size(self.group) != self.n_shards:
                raise ValueError("entity_shards must equal the size of process_group")
            self.rank = dist.get_rank(self.group)
        else:
            self.rank = 0
        if leaf_exchange is None:
            leaf_exchange = self.group is not None and os.environ.get("MVIN_B200_XCHG", "1") != "0"
        self.leaf_exchange = bool(leaf_exchange) and self.n_shards > 1
        self._xchg = None
        self._parse_args(args, adj_entity, adj_relation)
        self.n_user, self.n_entity, self.n_relation = int(n_user), int(n_entity), int(n_relation)
        self._build_inputs()
        with torch.cuda.device(self.device):
            self._create_handle()
            self._build_model(seed)
            self._build_train()
        if getattr(args, "load_pretrain_emb", False):
            self.load_pretrain_emb_fuc()

    # ------------------------------------------------------------------ model.py:17-47
    def _parse_args(self, args, adj_entity, adj_relation):
        self.args = args
        self.dataset = getattr(args, "dataset", None)
        self.load_pretrain_emb = getattr(args, "load_pretrain_emb", False)
        self.h_hop = int(args.h_hop)
        self.batch_size = int(args.batch_size)
        self.n_neighbor = int(args.neighbor_sample_size)
        self.p_hop = int(args.p_hop)
        self.dim = int(args.dim)
        self.l2_weight = float(args.l2_weight)
        self.l2_agg_weight = float(args.l2_agg_weight)
        self.kge_weight = getattr(args, "kge_weight", None)
        self.lr = float(args.lr)
        self.save_model_name = getattr(args, "save_model_name", "mvin")
        self.n_mix_hop = int(getattr(args, "n_mix_hop", 1))
        self.n_memory = int(args.n_memory)
        self.path = getattr(args, "path", None)
        self.flags = flags_from_args(args)
        if isinstance(adj_entity, torch.Tensor) and adj_relation is None:
            if (adj_entity.dtype != torch.int32 or adj_entity.ndim != 3 or adj_entity.shape[1] != 2
                    or adj_entity.shape[2] != self.n_neighbor or not adj_entity.is_cuda or not adj_entity.is_contiguous()):
                raise ValueError(f"packed adjacency must be a contiguous CUDA int32 [n_entity, 2, {self.n_neighbor}] tensor")
            self.adj_entity = self.adj_relation = None
            self._adj_packed_in = adj_entity
            return
        self._adj_packed_in = None
        adj_entity = np.ascontiguousarray(adj_entity, dtype=np.int64)
        adj_relation = np.ascontiguousarray(adj_relation, dtype=np.int64)
        if adj_entity.shape != adj_relation.shape or adj_entity.ndim != 2 or adj_entity.shape[1] != self.n_neighbor:
            raise ValueError(f"adj_entity / adj_relation must be int64 [n_entity, {self.n_neighbor}]")
        self.adj_entity, self.adj_relation = adj_entity, adj_relation

    # ------------------------------------------------------------------ model.py:49-64
    def _build_inputs(self):
        self.user_indices = _Placeholder("user_indices", np.int64, [None])
        self.item_indices = _Placeholder("item_indices", np.int64, [None])
        self.labels = _Placeholder("labels", np.float32, [None])
        self.memories_h, self.memories_r, self.memories_t = [], [], []
        for hop in range(max(1, self.p_hop)):
            self.memories_h.append(_Placeholder(f"memories_h_{hop}", np.int32, [None, self.n_memory]))
            self.memories_r.append(_Placeholder(f"memories_r_{hop}", np.int32, [None, self.n_memory]))
            self.memories_t.append(_Placeholder(f"memories_t_{hop}", np.int32, [None, self.n_memory]))

    def _create_handle(self):
        cfg = Config(dim=self.dim, neighbor_sample_size=self.n_neighbor, h_hop=self.h_hop, n_mix_hop=self.n_mix_hop,
                     p_hop=self.p_hop, n_memory=self.n_memory, n_user=self.n_user, n_entity=self.n_entity,
                     n_relation=self.n_relation, max_batch=self.batch_size, l2_weight=self.l2_weight,
                     l2_agg_weight=self.l2_agg_weight, flags=self.flags)
        self._handle = C.c_void_p()
        check(self.lib.mvin_create(C.byref(cfg), C.byref(self._handle)), "mvin_create")

    # ------------------------------------------------------------------ model.py:69-122, aggregators.py:83-93
    def param_shapes(self) -> Dict[str, tuple]:
        """Field shapes of include/mvin_b200.h `mvin_params_t`.  With M = n_mix_hop mix blocks of h_hop iterations there
        are L = h_hop M aggregators (aggregator (i, n) of model.py:290 at index n h_hop + i), L + 1 transfer matrices and
        M mix layers (stacked on a leading axis when M > 1)."""
        d, H, M, p = self.dim, self.h_hop, self.n_mix_hop, self.p_hop
        L = H * M
        n_ent_rows = self.n_entity if self.n_shards == 1 else self.n_local_rows
        mix_w, mix_b = (((H + 1) * d, d), (d,)) if M == 1 else ((M, (H + 1) * d, d), (M, d))
        return {"user_emb": (self.n_user, d), "entity_emb": (n_ent_rows, d), "relation_emb": (self.n_relation, d),
                "relation_kge": (self.n_relation, d, d), "mix_w": mix_w, "mix_b": mix_b,
                "user_mlp_w": ((p + (1 if self.flags & 0x08 else 0)) * d, d), "user_mlp_b": (d,),   # model.py:100-104
                "transfer_w": (L + 1, d, d),
                "transfer_b": (L + 1, d), "h_item_w": (2 * d,), "h_item_b": (1,), "agg_w": (L, d, d),
                "agg_b": (L, d), "agg_urh_w": (L, 3 * d), "agg_urh_b": (L,)}

    @staticmethod
    def _xavier(shape, gen, fan=None):
        """tf.contrib.layers.xavier_initializer (model.py:14-15): uniform(+-sqrt(6 / (fan_in + fan_out)))."""
        if fan is None:
            if len(shape) == 1:
                fan = (shape[0], shape[0])
            elif len(shape) == 2:
                fan = (shape[0], shape[1])
            else:
                rec = int(np.prod(shape[:-2]))
                fan = (shape[-2] * rec, shape[-1] * rec)
        limit = math.sqrt(6.0 / (fan[0] + fan[1]))
        return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * limit

    def _build_model(self, seed):
        d, H, p = self.dim, self.h_hop, self.p_hop
        gen = torch.Generator().manual_seed(seed)
        G = self.n_shards
        self.n_local_rows = (self.n_entity + G - 1) // G
        shapes = self.param_shapes()
        host = {}
        big = {}
        for name, shape in shapes.items():
            if name in ("entity_emb", "user_emb") and (G > 1 or int(np.prod(shape)) > (1 << 26)):
                # large or sharded tables are drawn on the device; fan-in/out are those of the FULL table
                rows_full = self.n_entity if name == "entity_emb" else self.n_user
                limit = math.sqrt(6.0 / (rows_full + d))
                dgen = torch.Generator(device=self.device).manual_seed(seed * 1000003 + 17 * self.rank + len(big))
                n_copies = G if (name == "entity_emb" and self.group is None) else 1
                t = torch.empty((n_copies,) + tuple(shape), dtype=torch.float32, device=self.device)
                t.uniform_(-limit, limit, generator=dgen)
                big[name] = t
                continue
            if name in ("agg_b", "agg_urh_b"):                     # aggregators.py:87,93 zero-init
                host[name] = torch.zeros(shape)
            elif name in ("transfer_w", "agg_w"):                  # stacks of [d, d] matrices
                host[name] = torch.stack([self._xavier((d, d), gen) for _ in range(shape[0])])
            elif name in ("transfer_b",):
                host[name] = torch.stack([self._xavier((d,), gen) for _ in range(shape[0])])
            elif name == "agg_urh_w":                              # [3d, 1] each
                host[name] = torch.stack([self._xavier((3 * d,), gen, fan=(3 * d, 1)) for _ in range(shape[0])])
            elif name == "h_item_w":                               # [2d, 1]
                host[name] = self._xavier((2 * d,), gen, fan=(2 * d, 1))
            elif name in ("mix_w", "mix_b") and self.n_mix_hop > 1:   # one [(H+1) d, d] layer / [d] bias per mix block
                host[name] = torch.stack([self._xavier(shape[1:], gen) for _ in range(shape[0])])
            else:
                host[name] = self._xavier(shape, gen)
        self.params = {k: v.to(self.device).contiguous() for k, v in host.items()}
        self._entity_all = self._entity_grad_all = None
        for name, t in big.items():
            if name == "entity_emb" and G > 1:
                self._entity_all = t                               # [G or 1, n_local_rows, d]; [0] is the local shard
                self.params[name] = t[0]
            else:
                self.params[name] = t[0]
        self.params = {k: self.params[k] for k in shapes}          # header field order
        # Every gradient tensor is a view into ONE flat buffer (segments padded to 64 floats, so each view keeps the
        # 16-byte alignment the C ABI asks for): the data-parallel gradient exchange is then a single in-place
        # all-reduce of `grad_flat`, no gather / scatter copies.  The row-sharded entity gradient stays outside (peers
        # reduce into it directly).
        names = [k for k in shapes if not (G > 1 and k == "entity_emb")]
        pad = lambda n: (n + 63) // 64 * 64
        self.grad_flat = torch.zeros(sum(pad(self.params[k].numel()) for k in names) + 64, dtype=torch.float32,
                                     device=self.device)
        self.grads, off = {}, 0
        for k in names:
            n = self.params[k].numel()
            self.grads[k] = self.grad_flat[off:off + n].view(self.params[k].shape)
            off += pad(n)
        self._user_grad_end = pad(self.params["user_emb"].numel())   # user_emb is the first segment
        self.loss_slot = self.grad_flat[off:off + 4]                 # 4 loss scalars ride in the same bucket
        if G > 1:
            self._entity_grad_all = torch.zeros_like(self._entity_all)
            self.grads["entity_emb"] = self._entity_grad_all[0]
        self.grads = {k: self.grads[k] for k in shapes}
        self.adam_m = {k: torch.zeros_like(v) for k, v in self.params.items()}
        self.adam_v = {k: torch.zeros_like(v) for k, v in self.params.items()}
        self._p_struct = self._make_struct(self.params)
        self._g_struct = self._make_struct(self.grads)
        self._m_struct = self._make_struct(self.adam_m)
        self._v_struct = self._make_struct(self.adam_v)
        check(self.lib.mvin_bind_params(self._handle, C.byref(self._p_struct)), "mvin_bind_params")
        check(self.lib.mvin_bind_grads(self._handle, C.byref(self._g_struct)), "mvin_bind_grads")
        if G > 1:
            self._bind_shards()
        # packed adjacency
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if self._adj_packed_in is not None:
            if self._adj_packed_in.shape[0] != self.n_entity:
                raise ValueError("packed adjacency has the wrong number of rows")
            self.adj_packed = self._adj_packed_in
        else:
            adj_e = torch.from_numpy(self.adj_entity).to(self.device)
            adj_r = torch.from_numpy(self.adj_relation).to(self.device)
            self.adj_packed = torch.empty((self.n_entity, 2, self.n_neighbor), dtype=torch.int32, device=self.device)
            check(self.lib.mvin_pack_adjacency(adj_e.data_ptr(), adj_r.data_ptr(), self.n_entity, self.n_neighbor,
                                               self.adj_packed.data_ptr(), stream), "mvin_pack_adjacency")
            torch.cuda.synchronize(self.device)
            del adj_e, adj_r
        check(self.lib.mvin_bind_adjacency(self._handle, self.adj_packed.data_ptr()), "mvin_bind_adjacency")
        torch.cuda.synchronize(self.device)
        self._workspace: Optional[torch.Tensor] = None
        self._workspace_B = 0
        self._staging: Optional[torch.Tensor] = None
        self._dev_feed = None

    def _bind_shards(self):
        """Shard table for mvin_bind_entity_shards: local pointers (virtual shards) or the peers' shards opened through
        CUDA IPC after an all-gather of the 64-byte handles (one process per GPU)."""
        G, rows, d = self.n_shards, self.n_local_rows, self.dim
        stride = rows * d * 4
        if self.group is None:
            e_ptrs = [self._entity_all.data_ptr() + g * stride for g in range(G)]
            g_ptrs = [self._entity_grad_all.data_ptr() + g * stride for g in range(G)]
        else:
            import torch.distributed as dist
            mine = []
            for t in (self._entity_all, self._entity_grad_all):
                hbuf = C.create_string_buffer(64)
                off = C.c_int64()
                check(self.lib.mvin_ipc_export(t.data_ptr(), hbuf, C.byref(off)), "mvin_ipc_export")
                mine.append((hbuf.raw, int(off.value)))
            everyone = [None] * G
            dist.all_gather_object(everyone, mine, group=self.group)
            e_ptrs, g_ptrs = [], []
            for r, (he, hg) in enumerate(everyone):
                if r == self.rank:
                    e_ptrs.append(self._entity_all.data_ptr())
                    g_ptrs.append(self._entity_grad_all.data_ptr())
                    continue
                for (raw, off), out in ((he, e_ptrs), (hg, g_ptrs)):
                    ptr = C.c_void_p()
                    check(self.lib.mvin_ipc_open(raw, off, C.byref(ptr)), "mvin_ipc_open")
                    out.append(ptr.value)
            dist.barrier(group=self.group)
        arr_e = (C.c_void_p * G)(*e_ptrs)
        arr_g = (C.c_void_p * G)(*g_ptrs)
        check(self.lib.mvin_bind_entity_shards(self._handle, G, arr_e, arr_g), "mvin_bind_entity_shards")
        self._shard_ptrs = (e_ptrs, g_ptrs)
        if self.group is not None:
            check(self.lib.mvin_set_batch_scale(self._handle, self.batch_size * G, 1.0 / G), "mvin_set_batch_scale")

    def _fence(self):
        """Stream-ordered barrier over the ranks of the group: a tiny NCCL all-reduce enqueued behind everything this
        rank has enqueued so far; what any rank enqueues after it runs only once every rank's earlier work is complete.
        No host synchronisation."""
        if self.group is not None:
            import torch.distributed as dist
            if getattr(self, "_fence_buf", None) is None:
                self._fence_buf = torch.zeros(1, dtype=torch.float32, device=self.device)
            dist.all_reduce(self._fence_buf, group=self.group)

    def begin_step(self):
        """Sharded mode: zero this rank's entity-gradient shard(s) and fence the ranks (peers scatter into it during
        the backward pass and read the entity shard during the forward pass).  No-op otherwise."""
        if self.n_shards == 1:
            return
        self._entity_grad_all.zero_()
        self._fence()                                # the zero fill has landed before any peer may scatter into it

    def end_step(self):
        """Sharded mode: fence the ranks after the backward pass (every peer's contribution has landed).  Not needed
        after allreduce_replicated(), which is itself a collective enqueued behind the backward pass of every rank."""
        self._fence()

    # ------------------------------------------------------------------ owner-side partial reduction (exchange.cuh)
    def _bind_exchange(self, B):
        """Exchange buffers for batches of B pairs: rows = B K^(H-1) leaf-level parent nodes per source rank.  One
        process per GPU: every rank allocates its receive / send buffers and maps the peers' through CUDA IPC (a
        collective: all ranks call it with the same B)."""
        G, d = self.n_shards, self.dim
        rows = B * self.n_neighbor ** (self.h_hop - 1)
        n_src = G if self.group is not None else 1
        x = dict(B=B, rows=rows,
                 part=torch.zeros((G, rows, d), dtype=torch.float32, device=self.device),
                 gsu=torch.zeros((rows, d), dtype=torch.float32, device=self.device),
                 dot=torch.zeros((G, rows), dtype=torch.float32, device=self.device),
                 ids=torch.zeros((n_src, rows), dtype=torch.int32, device=self.device),
                 ids_mine=torch.zeros((rows,), dtype=torch.int32, device=self.device))
        ptrs = {k: [x[k].data_ptr()] for k in ("part", "gsu", "dot")}
        if self.group is not None:
            import torch.distributed as dist
            mine = {}
            for k in ("part", "gsu", "dot"):
                hbuf = C.create_string_buffer(64)
                off = C.c_int64()
                check(self.lib.mvin_ipc_export(x[k].data_ptr(), hbuf, C.byref(off)), "mvin_ipc_export")
                mine[k] = (hbuf.raw, int(off.value))
            everyone = [None] * G
            dist.all_gather_object(everyone, mine, group=self.group)
            for k in ("part", "gsu", "dot"):
                ptrs[k] = []
                for r, theirs in enumerate(everyone):
                    if r == self.rank:
                        ptrs[k].append(x[k].data_ptr())
                        continue
                    ptr = C.c_void_p()
                    check(self.lib.mvin_ipc_open(theirs[k][0], theirs[k][1], C.byref(ptr)), "mvin_ipc_open")
                    ptrs[k].append(ptr.value)
            dist.barrier(group=self.group)
        arr = {k: (C.c_void_p * n_src)(*ptrs[k]) for k in ptrs}
        check(self.lib.mvin_xchg_bind(self._handle, n_src, self.rank if self.group is not None else 0, rows,
                                      x["ids"].data_ptr(), arr["part"], arr["gsu"], arr["dot"]), "mvin_xchg_bind")
        x["ptrs"] = ptrs
        self._xchg = x

    def _exchange_forward(self, items, B, ws):
        """Phases in front of mvin_forward: my leaf-level parent ids -> every owner; owners reduce the rows they hold."""
        if self._xchg is None or self._xchg["B"] != B:
            self._bind_exchange(B)
        x = self._xchg
        st = self._stream()
        if self.group is not None:
            import torch.distributed as dist
            check(self.lib.mvin_xchg_expand(self._handle, items.data_ptr(), B, x["ids_mine"].data_ptr(), ws.data_ptr(), st),
                  "mvin_xchg_expand")
            dist.all_gather_into_tensor(x["ids"].view(-1), x["ids_mine"], group=self.group)   # the index routing
            check(self.lib.mvin_xchg_owner_forward(self._handle, self.rank, ws.data_ptr(), st), "mvin_xchg_owner_forward")
            self._fence()                            # every owner's partials have landed in my receive buffer
        else:
            check(self.lib.mvin_xchg_expand(self._handle, items.data_ptr(), B, x["ids"].data_ptr(), ws.data_ptr(), st),
                  "mvin_xchg_expand")
            for g in range(self.n_shards):
                check(self.lib.mvin_xchg_owner_forward(self._handle, g, ws.data_ptr(), st), "mvin_xchg_owner_forward")

    def _exchange_backward(self, ws):
        """Phases behind mvin_backward: owners scatter-add into their own shards and return the softmax partials."""
        st = self._stream()
        if self.group is not None:
            self._fence()                            # every rank's gsu rows are complete
            check(self.lib.mvin_xchg_owner_backward(self._handle, self.rank, ws.data_ptr(), st), "mvin_xchg_owner_backward")
            self._fence()                            # every owner's partial dots have landed
        else:
            for g in range(self.n_shards):
                check(self.lib.mvin_xchg_owner_backward(self._handle, g, ws.data_ptr(), st), "mvin_xchg_owner_backward")
        check(self.lib.mvin_xchg_finish_backward(self._handle, ws.data_ptr(), st), "mvin_xchg_finish_backward")

    @staticmethod
    def _make_struct(tensors) -> Params:
        s = Params()
        for f in PARAM_FIELDS:
            setattr(s, f, tensors[f].data_ptr())
        return s

    def _build_train(self):
        self.step = 0
        self._losses_host = torch.zeros(4, dtype=torch.float32).pin_memory()
        self._losses_dev = torch.zeros(4, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ buffers
    def _ensure_workspace(self, B):
        if self._workspace is None or self._workspace_B != B:
            nbytes = self.lib.mvin_workspace_bytes(self._handle, B)
            if nbytes == 0:
                raise _lib.MvinError("mvin_workspace_bytes returned 0")
            # the layout is recomputed from B by every call of the library, so a buffer that is large enough is kept
            # (a smaller batch after a full one must not allocate a second multi-GB workspace)
            if self._workspace is None or self._workspace.numel() < nbytes:
                self._workspace = None
                self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            fbytes = self.lib.mvin_feed_bytes(self._handle, B)
            if self._staging is None or self._staging.numel() < fbytes:
                self._staging = torch.empty(fbytes, dtype=torch.uint8, device=self.device)
            self._workspace_B = B
        return self._workspace

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------ feed handling (train.py:112-122)
    def _pinned_feed(self, B, slot):
        """Page-locked host staging for the feed of one batch (two slots, used alternately by the prefetching train()):
        the feed-dict values are converted straight into these buffers, so the H2D copy of the C ABI is a true
        asynchronous DMA instead of a pageable copy."""
        if getattr(self, "_pin", None) is None or self._pin_B != B:
            n_mem = max(1, self.p_hop)
            mk = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
            self._pin = [(mk((B,), torch.int64), mk((B,), torch.int64), mk((B,), torch.float32),
                          mk((n_mem, B, self.n_memory), torch.int32), mk((n_mem, B, self.n_memory), torch.int32),
                          mk((n_mem, B, self.n_memory), torch.int32)) for _ in range(2)]
            self._pin_np = [tuple(t.numpy() for t in slot_t) for slot_t in self._pin]
            self._pin_B = B
        return self._pin_np[slot]

    def _host_feed(self, feed_dict, slot=0):
        """train.py:112-122 hands over NumPy slices for user / item / label and, per hop, either a Python list of B
        int32 [m] rows or a [B, m] array; everything is written into pinned staging buffers (views returned)."""
        items_in = np.asarray(feed_dict[self.item_indices])
        B = items_in.shape[0]
        if B > self.batch_size:
            raise ValueError(f"batch of {B} exceeds args.batch_size = {self.batch_size} (static in the reference graph)")
        users, items, labels, mem_h, mem_r, mem_t = self._pinned_feed(B, slot)
        items[...] = items_in
        users[...] = np.asarray(feed_dict[self.user_indices])
        lab = feed_dict.get(self.labels)
        if lab is None:
            labels[...] = 0
        else:
            labels[...] = np.asarray(lab)
        n_mem = max(1, self.p_hop)
        for keys, dst in ((self.memories_h, mem_h), (self.memories_r, mem_r), (self.memories_t, mem_t)):
            for hop in range(n_mem):
                v = feed_dict[keys[hop]]
                if isinstance(v, np.ndarray):
                    dst[hop] = v.reshape(B, self.n_memory)
                else:
                    rows_into(dst[hop], v)                        # list of B rows (train.py:118-120)
        return users, items, labels, mem_h, mem_r, mem_t

    def _device_feed(self, feed_dict):
        users, items, labels, mem_h, mem_r, mem_t = self._host_feed(feed_dict)
        to = lambda a: torch.from_numpy(a).to(self.device, non_blocking=False)
        self._dev_feed = tuple(to(a) for a in (users, items, labels, mem_h, mem_r, mem_t))   # keep alive for backward
        return (users.copy(), items.copy(), labels.copy()) + self._dev_feed   # copies: the staging is re-used

    # ------------------------------------------------------------------ device-resident entry points
    def forward_device(self, users, items, mem_h, mem_r, mem_t, scores=None, scores_normalized=None):
        """users/items int64 [B], mem_* int32 [max(1,p), B, m] -- torch CUDA tensors.  Enqueues the forward pass."""
        B = items.shape[0]
        ws = self._ensure_workspace(B)
        self._keepalive = (users, items, mem_h, mem_r, mem_t)     # the library reads them again in backward
        if self.leaf_exchange:
            self._exchange_forward(items, B, ws)
        check(self.lib.mvin_forward(self._handle, users.data_ptr(), items.data_ptr(), mem_h.data_ptr(), mem_r.data_ptr(),
                                    mem_t.data_ptr(), B, scores.data_ptr() if scores is not None else None,
                                    scores_normalized.data_ptr() if scores_normalized is not None else None,
                                    ws.data_ptr(), self._stream()), "mvin_forward")

    def backward_device(self, labels, losses=None):
        """labels float32 [B] CUDA tensor; fills self.grads and `losses` (4 floats: loss, base, l2, l2_agg)."""
        losses = self._losses_dev if losses is None else losses
        check(self.lib.mvin_backward(self._handle, labels.data_ptr(), labels.shape[0], losses.data_ptr(),
                                     self._workspace.data_ptr(), self._stream()), "mvin_backward")
        if self.leaf_exchange:
            self._exchange_backward(self._workspace)
        return losses

    def adam_step_device(self):
        self.step += 1
        check(self.lib.mvin_adam_step(self._handle, C.byref(self._m_struct), C.byref(self._v_struct), self.lr, 0.9,
                                      0.999, 1e-8, self.step, self._stream()), "mvin_adam_step")

    def train_step_host(self, users, items, labels, mem_h, mem_r, mem_t, apply_adam=True):
        """One fwd+bwd(+Adam) step from HOST numpy / pinned buffers through mvin_train_step_host (feed H2D copy and
        loss D2H inside).  Returns the 4 losses as a numpy array."""
        B = items.shape[0]
        ws = self._ensure_workspace(B)
        ptr = lambda a: a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr()
        if self.n_shards > 1:
            if apply_adam and self.group is None:
                raise NotImplementedError("virtual shards (entity_shards > 1 without a process group) are fwd/bwd only")
            self.begin_step()
        if self.leaf_exchange:
            # the exchange phases sit between library calls, so the step is driven from here: feed H2D, forward,
            # backward (with their owner phases), losses D2H
            to = lambda a: (torch.from_numpy(a) if isinstance(a, np.ndarray) else a).to(self.device, non_blocking=True)
            d = [to(a) for a in (users, items, labels, mem_h, mem_r, mem_t)]
            self.forward_device(d[0], d[1], d[3], d[4], d[5])
            self.backward_device(d[2], self._losses_dev)
            losses = self._losses_dev.cpu().numpy().copy()
            if self.group is not None:
                losses = self.allreduce_replicated(losses)
                if apply_adam:
                    self.adam_step_device()
            return losses
        fused_adam = apply_adam and self.group is None
        if fused_adam:
            self.step += 1
        check(self.lib.mvin_train_step_host(
            self._handle, ptr(users), ptr(items), ptr(labels), ptr(mem_h), ptr(mem_r), ptr(mem_t), B,
            self._staging.data_ptr(), ws.data_ptr(), C.byref(self._m_struct) if fused_adam else None,
            C.byref(self._v_struct) if fused_adam else None, self.lr, max(self.step, 1), self._losses_host.data_ptr(),
            self._stream()), "mvin_train_step_host")
        losses = self._losses_host.numpy().copy()
        if self.group is not None:
            losses = self.allreduce_replicated(losses)
            self.end_step()
            if apply_adam:
                self.adam_step_device()
        return losses

    # ------------------------------------------------------------------ double-buffered input pipeline
    def prefetch_feed(self, users, items, labels, mem_h, mem_r, mem_t):
        """Start the H2D copy of a coming batch (host numpy / pinned torch buffers, shapes of train_step_host) on the
        library's copy stream; it overlaps the step that is running.  At most two batches may be pending; each
        train_step_prefetched() consumes the oldest.  Loop shape: prefetch(0); for i: prefetch(i + 1); step()."""
        if self.n_shards > 1:
            raise NotImplementedError("prefetch pipeline: single-table configurations only")
        B = items.shape[0]
        self._ensure_workspace(B)
        if getattr(self, "_staging2", None) is None or self._staging2[0].numel() != self._staging.numel():
            self._staging2 = [self._staging, torch.empty_like(self._staging)]
            self._stage_next, self._pending = 0, []
        if len(self._pending) >= 2:
            raise RuntimeError("two batches are already pending: call train_step_prefetched() first")
        ptr = lambda a: a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr()
        slot = self._stage_next
        self._stage_next ^= 1
        buf = self._staging2[slot]
        self._pending.append((slot, B, (users, items, labels, mem_h, mem_r, mem_t)))   # keeps the host buffers alive
        check(self.lib.mvin_feed_prefetch(self._handle, ptr(users), ptr(items), ptr(labels), ptr(mem_h), ptr(mem_r),
                                          ptr(mem_t), B, buf.data_ptr(), slot, self._stream()), "mvin_feed_prefetch")

    def train_step_prefetched(self, apply_adam=True):
        """forward + backward (+ Adam) on the oldest pending batch of prefetch_feed(); returns the 4 losses."""
        slot, B, _ = self._pending.pop(0)
        if apply_adam:
            self.step += 1
        check(self.lib.mvin_train_step_prefetched(
            self._handle, B, self._staging2[slot].data_ptr(), slot, self._workspace.data_ptr(),
            C.byref(self._m_struct) if apply_adam else None, C.byref(self._v_struct) if apply_adam else None, self.lr,
            max(self.step, 1), self._losses_host.data_ptr(), self._stream()), "mvin_train_step_prefetched")
        return self._losses_host.numpy().copy()

    # ------------------------------------------------------------------ device-resident feed
    def bind_user_triplet_set(self, user_triplet_set):
        """Upload the packed ripple sets once: int32 [n_user, max(1,p), 3, n_memory] (data_loader_user_set.py:402 stacks
        one [p, 3, m] block per user).  After this, train_users / gather_feed replace get_feed_dict (train.py:112-122)."""
        uts = pack_user_triplet_set(user_triplet_set, self.n_user, max(1, self.p_hop), self.n_memory)
        want = (max(1, self.p_hop), 3, self.n_memory)
        if uts.ndim != 4 or tuple(uts.shape[1:]) != want:
            raise ValueError(f"user_triplet_set must be [n_user, {want[0]}, 3, {want[2]}], got {uts.shape}")
        self._uts = torch.from_numpy(uts).to(self.device)
        check(self.lib.mvin_bind_user_triplets(self._handle, self._uts.data_ptr()), "mvin_bind_user_triplets")

    def gather_feed(self, users):
        """users: int64 CUDA tensor [B] -> (mem_h, mem_r, mem_t) int32 CUDA tensors [max(1,p), B, m]."""
        B = users.shape[0]
        shape = (max(1, self.p_hop), B, self.n_memory)
        out = [torch.empty(shape, dtype=torch.int32, device=self.device) for _ in range(3)]
        check(self.lib.mvin_gather_feed(self._handle, users.data_ptr(), B, out[0].data_ptr(), out[1].data_ptr(),
                                        out[2].data_ptr(), self._stream()), "mvin_gather_feed")
        return tuple(out)

    def train_users(self, users, items, labels, apply_adam=True):
        """train(sess, get_feed_dict(...)) of the reference loop (train.py:62-64) with the feed assembled on the
        device: only user / item / label cross the bus.  HOST numpy or pinned torch buffers; returns the 4 losses."""
        if self.n_shards > 1:
            raise NotImplementedError("device-resident feed path: single-table configurations only")
        B = items.shape[0]
        ws = self._ensure_workspace(B)
        ptr = lambda a: a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr()
        if apply_adam:
            self.step += 1
        check(self.lib.mvin_train_step_users_host(
            self._handle, ptr(users), ptr(items), ptr(labels), B, self._staging.data_ptr(), ws.data_ptr(),
            C.byref(self._m_struct) if apply_adam else None, C.byref(self._v_struct) if apply_adam else None, self.lr,
            max(self.step, 1), self._losses_host.data_ptr(), self._stream()), "mvin_train_step_users_host")
        return self._losses_host.numpy().copy()

    def allreduce_replicated(self, losses=None):
        """One-process-per-GPU ranks: SUM all-reduce (NCCL) of the gradients of every replicated parameter (all but
        the sharded entity table, whose contributions already landed in the owners' shards through peer
        reductions, and the user table, which only carries its dense L2 term under --ablation all) and of the
        4 loss scalars."""
        import torch.distributed as dist
        if losses is not None:
            self.loss_slot.copy_(torch.as_tensor(losses, dtype=torch.float32))
        if self._user_grad_is_l2_only():
            # --ablation all: the user table only carries its dense L2 term (identical on every rank)
            dist.all_reduce(self.grad_flat[self._user_grad_end:], group=self.group)   # one bucket, in place
            self.grads["user_emb"].mul_(float(self.n_shards))
        else:
            # User_orient_kg_eh = 0: the KG side scatters into the user table, so it is exchanged with the rest
            dist.all_reduce(self.grad_flat, group=self.group)
        return self.loss_slot.cpu().numpy()

    def allreduce_grads(self, group=None):
        """Data-parallel replicas (tables replicated): ONE in-place SUM all-reduce of every gradient and
        of the 4 loss scalars (`loss_slot`, where backward_device(labels, model.loss_slot) put them).  With
        set_batch_scale(global batch, 1 / world) the result equals the single-device gradient on the concatenated
        batch (tests/test_sharding_gloo.py checks the protocol)."""
        import torch.distributed as dist
        if self._user_grad_is_l2_only():
            # --ablation all: nothing but the dense L2 term reaches the user table, and that term is
            # the same on every rank (l2 x U / world under set_batch_scale), so its sum is a local multiply -- the bucket
            # that crosses NVLink shrinks by the user table (C4: 48 -> 30 MB)
            dist.all_reduce(self.grad_flat[self._user_grad_end:], group=group)
            world = dist.get_world_size(group)
            if world > 1:
                self.grads["user_emb"].mul_(float(world))
        else:
            dist.all_reduce(self.grad_flat, group=group)

    def _user_grad_is_l2_only(self) -> bool:
        """U[user] is read only when User_orient_kg_eh = 0 (model.py:152-156) or HO_only = 1 (model.py:146-150)."""
        return bool(self.flags & 0x04) and not (self.flags & 0x40)

    def set_batch_scale(self, global_batch: int, dense_l2_scale: float):
        check(self.lib.mvin_set_batch_scale(self._handle, int(global_batch), float(dense_l2_scale)), "mvin_set_batch_scale")

    # ------------------------------------------------------------------ reference API (model.py:416-444)
    def train(self, sess, feed_dict):
        users, items, labels, mem_h, mem_r, mem_t = self._host_feed(feed_dict)
        losses = self.train_step_host(users, items, labels, mem_h, mem_r, mem_t, apply_adam=True)
        return None, float(losses[0])

    def loss_and_grads(self, feed_dict):
        """Forward + backward without the optimiser step: returns (losses[4], scores) and leaves self.grads filled."""
        users, items, labels, mem_h, mem_r, mem_t = self._host_feed(feed_dict)
        losses = self.train_step_host(users, items, labels, mem_h, mem_r, mem_t, apply_adam=False)
        return losses

    def _scores(self, feed_dict):
        _, items, labels, d_users, d_items, d_labels, d_mh, d_mr, d_mt = self._device_feed(feed_dict)
        B = items.shape[0]
        scores = torch.empty(B, dtype=torch.float32, device=self.device)
        scores_n = torch.empty(B, dtype=torch.float32, device=self.device)
        self.forward_device(d_users, d_items, d_mh, d_mr, d_mt, scores, scores_n)
        torch.cuda.synchronize(self.device)
        return items, labels, scores.cpu().numpy(), scores_n.cpu().numpy()

    def eval(self, sess, feed_dict):
        """(auc, acc, f1) of one batch (model.py:419-426) -- computed on the device (mvin_ctr_metrics: exact pair
        counts, ties 1/2, threshold 0.5), three floats come back instead of B scores for sklearn."""
        _, items, labels, d_users, d_items, d_labels, d_mh, d_mr, d_mt = self._device_feed(feed_dict)
        B = items.shape[0]
        scores_n = torch.empty(B, dtype=torch.float32, device=self.device)
        self.forward_device(d_users, d_items, d_mh, d_mr, d_mt, None, scores_n)
        return self.ctr_metrics_device(scores_n, d_labels)

    def topk_metrics_device(self, scores, relevant, n_cand, n_answers, k_list):
        """Per-user precision@k / recall@k / ndcg@k of topk_eval (util.py:183-197) on the device.  scores float32
        [n_users, max_cand], relevant uint8 [n_users, max_cand], n_cand / n_answers int32 [n_users] -- CUDA tensors;
        returns three numpy arrays [n_users, len(k_list)] (the reference then takes np.mean over users)."""
        n_users, max_cand = scores.shape
        nk = len(k_list)
        out = [torch.empty((n_users, nk), dtype=torch.float32, device=self.device) for _ in range(3)]
        ks = (C.c_int32 * nk)(*[int(k) for k in k_list])
        check(self.lib.mvin_topk_metrics(scores.data_ptr(), relevant.data_ptr(), n_cand.data_ptr(), n_answers.data_ptr(),
                                         n_users, max_cand, C.cast(ks, C.c_void_p), nk, out[0].data_ptr(), out[1].data_ptr(),
                                         out[2].data_ptr(), self._stream()), "mvin_topk_metrics")
        return tuple(o.cpu().numpy() for o in out)

    def ctr_metrics_device(self, scores_normalized, labels):
        """scores_normalized, labels: float32 CUDA tensors [B] -> (auc, acc, f1) python floats."""
        out = torch.empty(3, dtype=torch.float32, device=self.device)
        scratch = torch.empty(5, dtype=torch.int64, device=self.device)
        check(self.lib.mvin_ctr_metrics(self._handle, scores_normalized.data_ptr(), labels.data_ptr(),
                                        scores_normalized.shape[0], out.data_ptr(), scratch.data_ptr(), self._stream()),
              "mvin_ctr_metrics")
        auc, acc, f1 = out.cpu().tolist()
        return auc, acc, f1

    def get_scores(self, sess, feed_dict):
        items, _, _, scores_n = self._scores(feed_dict)
        return [items, scores_n]

    def get_raw_scores(self, feed_dict):
        _, _, scores, _ = self._scores(feed_dict)
        return scores

    def get_neighbors(self, items) -> (List[np.ndarray], List[np.ndarray]):
        """model.py:243-256 on the device; returns (entities[0..L], relations[0..L-1]) as int64 numpy arrays."""
        items = np.ascontiguousarray(items, dtype=np.int64)
        B, K, L = items.shape[0], self.n_neighbor, self.h_hop * self.n_mix_hop
        d_items = torch.from_numpy(items).to(self.device)
        ents = [torch.empty((B, K ** i), dtype=torch.int64, device=self.device) for i in range(L + 1)]
        rels = [torch.empty((B, K ** (i + 1)), dtype=torch.int64, device=self.device) for i in range(L)]
        e_ptrs = (C.c_void_p * (L + 1))(*[t.data_ptr() for t in ents])
        r_ptrs = (C.c_void_p * max(L, 1))(*[t.data_ptr() for t in rels])
        check(self.lib.mvin_get_neighbors(self._handle, d_items.data_ptr(), B, L, e_ptrs, r_ptrs, self._stream()),
              "mvin_get_neighbors")
        torch.cuda.synchronize(self.device)
        return [t.cpu().numpy() for t in ents], [t.cpu().numpy() for t in rels]

    def eval_case_study(self, sess, feed_dict):
        if self.flags & 0x20:
            # the reference never creates importance_list_0 / _1 under PS_only (model.py:142-144): its eval_case_study raises
            raise AttributeError("PS_only has no aggregators: importance_list_0 does not exist (model.py:142-144)")
        if not self.flags & 0x02:
            # User_orient_rela = 0: the aggregators return probs_normalized = None (aggregators.py:104-106) and the
            # reference's sess.run on importance_list_0 = None raises
            raise AttributeError("User_orient_rela = 0: there is no attention to report (aggregators.py:104-106)")
        users = np.asarray(feed_dict[self.user_indices])
        items, labels, _, _ = self._scores(feed_dict)
        B, K = items.shape[0], self.n_neighbor
        imp0 = torch.empty((B, 1, K), dtype=torch.float32, device=self.device)
        imp1 = torch.empty((B, K, K), dtype=torch.float32, device=self.device) if self.h_hop > 1 else None
        check(self.lib.mvin_importance(self._handle, imp0.data_ptr(), imp1.data_ptr() if imp1 is not None else None,
                                       self._workspace.data_ptr(), self._stream()), "mvin_importance")
        entities_data, relations_data = self.get_neighbors(items)
        torch.cuda.synchronize(self.device)
        return (users, labels, items, entities_data, relations_data, imp0.cpu().numpy(),
                imp1.cpu().numpy() if imp1 is not None else None)

    # ------------------------------------------------------------------ checkpoint of the four STWS tables
    def _emb_path(self):
        base = getattr(getattr(self.args, "path", None), "emb", None)
        if base is None:
            raise ValueError("args.path.emb is not set")
        return f"{base}_sw_para_{self.save_model_name}_parameter.npz"

    def save_pretrain_emb_fuc(self, sess=None, saver=None):
        """model.py:66-67 / train.py:43-54: persists the user, entity, relation and relation-KGE tables."""
        path = self._emb_path()
        tables = {k: self.params[k].cpu().numpy() for k in ("user_emb", "relation_emb", "relation_kge")}
        # row-sharded entity table: the checkpoint holds the FULL table (train.py:43-54 saves entity_emb_matrix whole);
        # every rank takes part in the gather, rank 0 alone writes
        tables["entity_emb"] = (self._gather_entity_table(self._entity_all) if self.n_shards > 1
                                else self.params["entity_emb"].cpu().numpy())
        if self.rank == 0:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            np.savez(path, **tables)
        if self.group is not None:
            import torch.distributed as dist
            dist.barrier(group=self.group)

    def load_pretrain_emb_fuc(self):
        z = np.load(self._emb_path())
        for k in ("user_emb", "relation_emb", "relation_kge"):
            self.params[k].copy_(torch.from_numpy(z[k]))
        full = torch.from_numpy(z["entity_emb"])
        if self.n_shards > 1:
            self._scatter_entity_table(full, self._entity_all)
        else:
            self.params["entity_emb"].copy_(full)
        torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------ test / interop helpers
    # oracle (reference attribute) name -> (field, index or None)
    def _name_map(self):
        H, M = self.h_hop, self.n_mix_hop
        mp = {"user_emb_matrix": ("user_emb", None), "entity_emb_matrix": ("entity_emb", None),
              "relation_emb_matrix": ("relation_emb", None), "relation_emb_KGE_matrix": ("relation_kge", None),
              "user_mlp_matrix": ("user_mlp_w", None), "user_mlp_bias": ("user_mlp_b", None),
              "h_emb_item_mlp_matrix": ("h_item_w", None), "h_emb_item_mlp_bias": ("h_item_b", None)}
        for n in range(M):                                         # model.py:91-98
            mp[f"enti_transfer_matrix_{n}"] = ("mix_w", None if M == 1 else n)
            mp[f"enti_transfer_bias_{n}"] = ("mix_b", None if M == 1 else n)
        for e in range(H * M + 1):                                 # model.py:107-116
            mp[f"transfer_agg_matrix_{e}"] = ("transfer_w", e)
            mp[f"transfer_agg_bias_{e}"] = ("transfer_b", e)
        if self.flags & 0x20:                                      # PS_only: the reference creates no aggregators
            return mp
        for n in range(M):                                         # created n outer, i inner (model.py:286-291)
            for i in range(H):
                g = n * H + i
                mp[f"agg_{i}_{n}_weights"] = ("agg_w", g)
                mp[f"agg_{i}_{n}_bias"] = ("agg_b", g)
                mp[f"agg_{i}_{n}_urh_weights"] = ("agg_urh_w", g)
                mp[f"agg_{i}_{n}_urh_bias"] = ("agg_urh_b", g)
        return mp

    def load_named_parameters(self, named: Dict[str, np.ndarray]):
        """Set parameters from a dict keyed by the reference's variable names (see oracle.param_shapes)."""
        for name, (field, idx) in self._name_map().items():
            src = torch.as_tensor(np.asarray(named[name]), dtype=torch.float32)
            if field == "entity_emb" and self.n_shards > 1:
                self._scatter_entity_table(src, self._entity_all)
                continue
            dst = self.params[field] if idx is None else self.params[field][idx]
            dst.copy_(src.reshape(dst.shape))
        torch.cuda.synchronize(self.device)

    def _scatter_entity_table(self, full, shards):
        """full [n_entity, d] (host) -> shards[g or 0][e // G] for e % G == g (all shards, or this rank's one)."""
        G = self.n_shards
        for g in (range(G) if self.group is None else [self.rank]):
            shards[g if self.group is None else 0].copy_(sharding.scatter_table(full, G, g))

    def _gather_entity_table(self, shards) -> np.ndarray:
        """Inverse of _scatter_entity_table; with a process group every rank receives the full table."""
        G = self.n_shards
        if self.group is None:
            parts = [shards[g] for g in range(G)]
        else:
            import torch.distributed as dist
            parts = [torch.empty_like(shards[0]) for _ in range(G)]
            dist.all_gather(parts, shards[0].contiguous(), group=self.group)
        return sharding.gather_table(parts, self.n_entity).numpy().copy()

    def entity_rows(self, ids) -> np.ndarray:
        """Rows `ids` (int64 NumPy) of the entity table as float32 [len, d] on the host, whatever its layout: one table,
        virtual shards, or shards spread over the ranks of the process group -- then a COLLECTIVE: every rank calls it,
        rank 0's ids are used and every rank gets the rows."""
        G, dev = self.n_shards, self.device
        ids_t = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64)).to(dev)
        if G == 1:
            return self.params["entity_emb"][ids_t].cpu().numpy()
        if self.group is None:
            return self._entity_all[ids_t % G, ids_t // G].cpu().numpy()
        import torch.distributed as dist
        src = dist.get_global_rank(self.group, 0)
        n = torch.tensor([ids_t.numel()], dtype=torch.int64, device=dev)
        dist.broadcast(n, src=src, group=self.group)
        if self.rank != 0:
            ids_t = torch.empty(int(n.item()), dtype=torch.int64, device=dev)
        dist.broadcast(ids_t, src=src, group=self.group)
        out = torch.zeros((ids_t.numel(), self.dim), dtype=torch.float32, device=dev)
        mine = (ids_t % G) == self.rank
        out[mine] = self._entity_all[0][ids_t[mine] // G]
        dist.all_reduce(out, group=self.group)
        return out.cpu().numpy()

    def _export(self, tensors, tables=True) -> Dict[str, np.ndarray]:
        from_shapes = {"h_emb_item_mlp_matrix": (2 * self.dim, 1), "h_emb_item_mlp_bias": (1,)}
        out = {}
        for name, (field, idx) in self._name_map().items():
            if not tables and field in ("entity_emb", "user_emb"):
                continue
            if field == "entity_emb" and self.n_shards > 1:
                out[name] = self._gather_entity_table(self._entity_all if tensors is self.params
                                                      else self._entity_grad_all)
                continue
            t = tensors[field] if idx is None else tensors[field][idx]
            a = t.detach().cpu().numpy().copy()
            if name.endswith("urh_weights"):
                a = a.reshape(3 * self.dim, 1)
            elif name.endswith("urh_bias"):
                a = a.reshape(1)
            elif name in from_shapes:
                a = a.reshape(from_shapes[name])
            out[name] = a
        return out

    def named_parameters(self, tables=True):
        """Parameters under the reference's variable names; tables=False leaves out the entity and user tables."""
        return self._export(self.params, tables)

    def named_gradients(self):
        return self._export(self.grads)

    def launch_count(self) -> int:
        return int(self.lib.mvin_launch_count(self._handle))

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self.lib.mvin_destroy(self._handle)
                self._handle = None
        except Exception:
            pass
