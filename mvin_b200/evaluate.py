"""Evaluation drivers of the reference's training loop with the scoring and the metrics on the device.

Mirrors the call shapes of `util.py` of the reference (`ctr_eval` :44-56, `topk_eval` :137-205) so that `train.py`
can switch its import; the per-batch `model.eval` / `model.get_scores` calls of the reference become one packed
scoring pass over all (user, candidate) pairs and one `mvin_topk_metrics` launch.  SURVEY.md 8(f) rank 4.

Scoring is per pair (no cross-pair term in the forward), so packing several users into one batch gives the scores the
reference gets from its one-user-per-batch loop; the reference's padding of the last batch with a repeated item only
re-inserts the same score into its dict and is dropped here.
"""
from typing import Dict, Iterable, List, Sequence, Set

import numpy as np
import torch


def ctr_eval(args, user_path, sess, model, data, user_triplet_set, batch_size):
    """util.py:44-56: (auc_list, acc_list, f1_list, mean auc, mean acc, mean f1) over the full batches of `data`.
    With ripple sets bound on the device (`model.bind_user_triplet_set`) only user / item / label rows are uploaded;
    `user_triplet_set` is then unused (kept for the reference's signature)."""
    data = np.asarray(data)
    n_full = data.shape[0] // batch_size
    auc_list, acc_list, f1_list = [], [], []
    if n_full == 0:
        return auc_list, acc_list, f1_list, float("nan"), float("nan"), float("nan")
    dev = model.device
    d_data = torch.from_numpy(np.ascontiguousarray(data[:n_full * batch_size, :3], dtype=np.int64)).to(dev)
    scores_n = torch.empty(batch_size, dtype=torch.float32, device=dev)
    for b in range(n_full):
        rows = d_data[b * batch_size:(b + 1) * batch_size]
        users, items = rows[:, 0].contiguous(), rows[:, 1].contiguous()
        labels = rows[:, 2].to(torch.float32)
        mh, mr, mt = _memories(model, users, user_triplet_set)
        model.forward_device(users, items, mh, mr, mt, None, scores_n)
        auc, acc, f1 = model.ctr_metrics_device(scores_n, labels)
        auc_list.append(auc)
        acc_list.append(acc)
        f1_list.append(f1)
    return auc_list, acc_list, f1_list, float(np.mean(auc_list)), float(np.mean(acc_list)), float(np.mean(f1_list))


def _memories(model, d_users, user_triplet_set):
    """Ripple sets of a batch as device tensors [max(1,p), B, m]: gathered on the device when the packed table is
    bound, else stacked on the host the way get_feed_dict does (util.py:207-218) and uploaded."""
    if getattr(model, "_uts", None) is not None:
        return model.gather_feed(d_users)
    from collections.abc import Mapping
    users = d_users.cpu().numpy()
    if isinstance(user_triplet_set, Mapping):
        # the reference's defaultdict(user -> int32 [p, 3, m]) (data_loader_user_set.py:396-402), util.py:210-217
        blk = np.stack([np.asarray(user_triplet_set[int(u)], dtype=np.int32) for u in users])
    else:
        blk = np.asarray(user_triplet_set)[users]                      # [B, p, 3, m]
    stack = lambda c: torch.from_numpy(np.ascontiguousarray(blk[:, :, c, :].transpose(1, 0, 2), dtype=np.int32)).to(
        model.device)
    return stack(0), stack(1), stack(2)


def score_candidates(model, users: Sequence[int], candidates: List[Sequence[int]], user_triplet_set=None,
                     batch_size: int = 4096):
    """normalized scores of every (users[i], candidates[i][j]) pair: float32 CUDA tensor [n_users, max_cand]
    (entries past len(candidates[i]) are padding), scored in packed batches of `batch_size` pairs (at most the
    model's args.batch_size, which sizes the library's workspace)."""
    batch_size = min(int(batch_size), model.batch_size)
    n_users = len(users)
    n_cand = np.array([len(c) for c in candidates], dtype=np.int64)
    max_cand = int(n_cand.max()) if n_users else 0
    item_mat = np.zeros((n_users, max_cand), dtype=np.int64)
    for i, c in enumerate(candidates):
        item_mat[i, :len(c)] = np.asarray(c, dtype=np.int64)
    valid = np.arange(max_cand)[None, :] < n_cand[:, None]
    flat_pos = np.flatnonzero(valid.ravel())                           # user-major order of the real pairs
    dev = model.device
    d_users = torch.from_numpy(np.repeat(np.asarray(users, dtype=np.int64), n_cand)).to(dev)
    d_items = torch.from_numpy(item_mat.ravel()[flat_pos]).to(dev)
    flat_scores = torch.zeros(d_users.shape[0], dtype=torch.float32, device=dev)
    aligned = lambda t: t if t.data_ptr() % 16 == 0 else t.clone()    # the C ABI asks for 16-byte aligned pointers
    out = torch.empty(batch_size, dtype=torch.float32, device=dev)
    for s in range(0, d_users.shape[0], batch_size):
        u, it = aligned(d_users[s:s + batch_size]), aligned(d_items[s:s + batch_size])
        mh, mr, mt = _memories(model, u, user_triplet_set)
        model.forward_device(u, it, mh, mr, mt, None, out[:u.shape[0]])
        flat_scores[s:s + u.shape[0]] = out[:u.shape[0]]
    scores = torch.zeros(n_users * max_cand, dtype=torch.float32, device=dev)
    scores[torch.from_numpy(flat_pos).to(dev)] = flat_scores
    return scores.view(n_users, max_cand), item_mat, n_cand


def topk_eval(sess, args, user_triplet_set, model, user_list: Iterable[int], train_record: Dict[int, Set[int]],
              eval_record: Dict[int, Set[int]], test_record: Dict[int, Set[int]], item_set: Set[int], k_list,
              batch_size, mode="test"):
    """util.py:137-205: (precision, recall, ndcg, None, None), each a list over k_list of the mean over the users of
    `user_list` present in the eval / test record.  Candidates of a user = item_set - train_record[user], in the order
    the reference iterates them (ties between equal scores keep that order, as sorted() does there); ndcg's hit list
    is cut at the last k of k_list (:190-195)."""
    ref_user = eval_record if mode == "eval" else test_record
    users = [u for u in user_list if u in ref_user]
    if not users:
        nan = [float("nan")] * len(k_list)
        return nan, list(nan), list(nan), None, None
    candidates = [list(item_set - train_record[u]) for u in users]
    scores, item_mat, n_cand = score_candidates(model, users, candidates, user_triplet_set, batch_size)
    relevant = np.zeros(item_mat.shape, dtype=np.uint8)
    for i, u in enumerate(users):
        relevant[i, :n_cand[i]] = np.isin(item_mat[i, :n_cand[i]], np.fromiter(ref_user[u], dtype=np.int64))
    dev = model.device
    n_answers = np.array([len(ref_user[u]) for u in users], dtype=np.int32)
    prec, rec, ndcg = model.topk_metrics_device(scores, torch.from_numpy(relevant).to(dev),
                                                torch.from_numpy(n_cand.astype(np.int32)).to(dev),
                                                torch.from_numpy(n_answers).to(dev), list(k_list))
    mean = lambda a: [float(np.mean(a[:, q])) for q in range(len(k_list))]
    return mean(prec), mean(rec), mean(ndcg), None, None
