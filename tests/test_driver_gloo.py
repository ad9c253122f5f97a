"""Multi-process host logic on CPU: pair sharding + the single all-gather of result records (gloo, world 2)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_pairs, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bufferx_b200.driver import gather_records, pack_record, shard_indices, unpack_record
    mine = shard_indices(n_pairs, rank, world)
    recs = []
    for i in mine:
        pose = np.eye(4)
        pose[0, 3] = i
        recs.append(pack_record(i, pose, [0.1 * i, 0.2, 0.3], i + 1, i + 2, i + 3, 3, rte=0.5 * i, rre=1.0, success=i % 2))
    allr = gather_records(np.stack(recs) if recs else np.zeros((0, 32), np.float32), n_pairs)
    assert allr.shape == (n_pairs, 32)
    for i in range(n_pairs):
        r = unpack_record(allr[i])
        assert r["pair_id"] == i and r["pose"][0, 3] == i and r["num_inliers"] == i + 1 and r["success"] == bool(i % 2)
    if rank == 0:
        np.save(out, allr)
    dist.destroy_process_group()


def test_pair_sharding_and_gather_world2(tmp_path):
    n_pairs = 7                                   # odd: ranks own 4 and 3 pairs, blocks are padded
    out = str(tmp_path / "all.npy")
    mp.spawn(_worker, args=(2, 29517, n_pairs, out), nprocs=2, join=True)
    allr = np.load(out)
    assert (allr[:, 26] == np.arange(n_pairs)).all()


def test_shard_indices_round_robin():
    sys.path.insert(0, ROOT)
    from bufferx_b200.driver import shard_indices
    assert shard_indices(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((shard_indices(10, r, 4) for r in range(4)), [])) == list(range(10))
