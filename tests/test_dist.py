"""Host-side logic of the multi-GPU path on CPU: image sharding and the single gather of
detection records, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_frames_covers_everything_once():
    from masklab_b200 import dist as mdist
    for total, world in ((1024, 8), (1000, 8), (5, 4), (0, 2), (33, 1), (7, 8)):
        seen = []
        for r in range(world):
            sh = mdist.shard_frames(total, world, r)
            assert sh.count <= sh.padded
            seen += list(range(sh.start, sh.start + sh.count))
        assert seen == list(range(total))
    assert mdist.chunks(70) == [(0, 32), (32, 32), (64, 6)] and mdist.chunks(0) == []
    with pytest.raises(ValueError):
        mdist.shard_frames(4, 2, 2)


def _worker(rank, world, port, total, K, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from masklab_b200 import dist as mdist
    sh = mdist.shard_frames(total, world, rank)
    # every frame's records are a function of its global index, so the gathered result is checkable
    det = torch.full((sh.count, K, 6), -1.0)
    counts = torch.zeros((sh.count,), dtype=torch.int32)
    for i in range(sh.count):
        g = sh.start + i
        n = g % (K + 1)
        counts[i] = n
        det[i, :n] = float(g)
    det, counts = mdist.pad_shard(det, counts, sh.padded)
    all_det, all_counts = mdist.gather_detections(det, counts, total_frames=total)
    np.save(os.path.join(out_dir, f"det{rank}.npy"), all_det.numpy())
    np.save(os.path.join(out_dir, f"cnt{rank}.npy"), all_counts.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [6, 5])
def test_gather_detections_world2_gloo(tmp_path, total):
    world, K = 2, 4
    mp.spawn(_worker, args=(world, _free_port(), total, K, str(tmp_path)), nprocs=world, join=True)
    d0, d1 = np.load(tmp_path / "det0.npy"), np.load(tmp_path / "det1.npy")
    c0, c1 = np.load(tmp_path / "cnt0.npy"), np.load(tmp_path / "cnt1.npy")
    assert np.array_equal(d0, d1) and np.array_equal(c0, c1)
    assert d0.shape == (total, K, 6)
    for g in range(total):
        n = g % (K + 1)
        assert c0[g] == n
        assert np.all(d0[g, :n] == g) and np.all(d0[g, n:] == -1)


def test_single_process_gather_is_identity():
    from masklab_b200 import dist as mdist
    det, counts = torch.zeros((3, 2, 6)), torch.tensor([1, 0, 2], dtype=torch.int32)
    d, c = mdist.gather_detections(det, counts)
    assert d is det and c is counts
