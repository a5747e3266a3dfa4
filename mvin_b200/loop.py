"""The reference's epoch loop (train.py:56-64) with shuffling, batching and feed assembly on the device.

    np.random.shuffle(train_data)
    while start + batch_size <= len(train_data):        # the incomplete tail is skipped
        _, loss = model.train(sess, get_feed_dict(..., start, start + batch_size))

SURVEY.md 8(f) rank 2: the interaction table is uploaded once, permuted with a device generator, and every step reads
its (user, item, label) rows and gathers its ripple sets on the GPU (`MVIN.bind_user_triplet_set` first); nothing crosses
the bus inside the loop and the losses stay on the device until the caller reads them.
"""
import numpy as np
import torch


def upload_interactions(model, data):
    """`data`: int [N, 3] (user, item, label) as train.py holds it -> int64 CUDA tensor [N, 3]."""
    return torch.from_numpy(np.ascontiguousarray(np.asarray(data)[:, :3], dtype=np.int64)).to(model.device)


def train_epoch(model, d_data, batch_size=None, shuffle=True, generator=None, apply_adam=True):
    """One epoch over the device-resident interactions `d_data` (upload_interactions).  Returns the per-step losses as
    a float32 CUDA tensor [n_steps, 4] (loss, base, l2, l2_agg -- model.py:379-412); `.cpu()` it once per epoch."""
    if getattr(model, "_uts", None) is None:
        raise RuntimeError("train_epoch needs the ripple sets on the device: call model.bind_user_triplet_set first")
    if model.n_shards > 1:
        raise NotImplementedError("device-resident epoch loop: single-table configurations only")
    B = int(batch_size or model.batch_size)
    n_steps = d_data.shape[0] // B
    losses = torch.zeros((n_steps, 4), dtype=torch.float32, device=model.device)
    if n_steps == 0:
        return losses
    if shuffle:
        perm = torch.randperm(d_data.shape[0], device=d_data.device, generator=generator)
        d_data = d_data[perm]
    cols = d_data[:n_steps * B].t().contiguous()                       # [3, n_steps * B]: each batch slice contiguous
    labels = cols[2].to(torch.float32)
    aligned = lambda t: t if t.data_ptr() % 16 == 0 else t.clone()    # the C ABI asks for 16-byte aligned pointers
    for s in range(n_steps):
        lo, hi = s * B, (s + 1) * B
        users, items, lab = aligned(cols[0, lo:hi]), aligned(cols[1, lo:hi]), aligned(labels[lo:hi])
        mem_h, mem_r, mem_t = model.gather_feed(users)
        model.forward_device(users, items, mem_h, mem_r, mem_t)
        model.backward_device(lab, losses[s])
        if apply_adam:
            model.adam_step_device()
    return losses
