"""Pair sharding across GPUs: one process per GPU, pairs round-robin, ONE all-gather of records.

The reference is single-process / single-GPU (``nn.DataParallel(model, [gpu])``,
/root/reference/test.py:105) and processes pairs in a serial loop with no cross-pair state
(test.py:132-146), so the path shards embarrassingly: pair i -> rank i mod world.  The only collective
is the final gather of a fixed 32-float record per pair (4x4 pose, 3 timings, 4 counters, rte, rre,
success, pair id + padding).  Works with ``nccl`` (GPU tensors) and ``gloo`` (CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

RECORD = 32  # floats per pair


def shard_indices(n_pairs: int, rank: int, world: int):
    """Round-robin assignment (BASELINE.json config 4): pair i -> rank i mod world."""
    return list(range(rank, n_pairs, world))


def pack_record(pair_id, pose, times, num_inliers, num_mutual, num_inlier_ind, scales_used, rte=np.nan, rre=np.nan, success=0.0):
    r = np.zeros(RECORD, dtype=np.float32)
    r[:16] = np.asarray(pose, dtype=np.float64).reshape(16)
    r[16:19] = np.asarray(times, dtype=np.float64)[:3]
    r[19:23] = [num_inliers, num_mutual, num_inlier_ind, scales_used]
    r[23], r[24], r[25], r[26] = rte, rre, success, pair_id
    return r


def unpack_record(r):
    r = np.asarray(r)
    return dict(pair_id=int(round(float(r[26]))), pose=r[:16].reshape(4, 4).astype(np.float64), times=r[16:19].tolist(),
                num_inliers=int(r[19]), num_mutual=int(r[20]), num_inlier_ind=int(r[21]), scales_used=int(r[22]),
                rte=float(r[23]), rre=float(r[24]), success=bool(r[25] > 0.5))


def gather_records(local: np.ndarray, n_pairs: int, device=None):
    """local: [n_local, RECORD] records of this rank -> [n_pairs, RECORD] on every rank, ordered by pair id.
    Uses a single all_gather of equally padded blocks (ranks may own ceil or floor(n_pairs/world) pairs)."""
    local = np.asarray(local, dtype=np.float32).reshape(-1, RECORD)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = local
    else:
        world = dist.get_world_size()
        per = (n_pairs + world - 1) // world
        blk = torch.full((per, RECORD), float("nan"), dtype=torch.float32)
        blk[: local.shape[0]] = torch.from_numpy(local)
        if device is not None:
            blk = blk.to(device)
        bufs = [torch.empty_like(blk) for _ in range(world)]
        dist.all_gather(bufs, blk)
        out = torch.cat(bufs).cpu().numpy()
        out = out[~np.isnan(out[:, 26])]
    order = np.argsort(out[:, 26], kind="stable")
    return out[order]
