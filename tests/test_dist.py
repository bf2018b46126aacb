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


def _record_worker(rank, world, port, B, K, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from masklab_b200 import dist as mdist
    _, words = mdist.record_words(B, K)
    rec = torch.zeros((words,), dtype=torch.int32)
    det, counts = mdist.record_views(rec, B, K)
    det.fill_(-1.0)
    for i in range(B):
        g = rank * B + i
        n = g % (K + 1)
        counts[i] = n
        det[i, :n] = float(g) + 0.25
    out = torch.full((world, words), 7, dtype=torch.int32)
    gathered = mdist.gather_records(rec, out=out)              # ONE collective
    ok = mdist.verify_gather(gathered, rec)
    # a corrupted slice must be caught on every rank
    bad = gathered.clone()
    bad[(rank + 1) % world, 3] ^= 1
    caught = not mdist.verify_gather(bad, rec)
    all_det, all_counts = mdist.unpack_records(gathered, B, K)
    np.save(os.path.join(out_dir, f"rdet{rank}.npy"), all_det.numpy())
    np.save(os.path.join(out_dir, f"rcnt{rank}.npy"), all_counts.numpy())
    np.save(os.path.join(out_dir, f"rok{rank}.npy"), np.array([ok, caught]))
    dist.destroy_process_group()


@pytest.mark.parametrize("B,K", [(3, 5), (2, 4)])
def test_packed_record_gather_world2_gloo(tmp_path, B, K):
    """The packed record (det + counts in one buffer) through ONE all_gather_into_tensor, verified the way
    bench.py verifies the NCCL gather on the GPUs (own slice bit-equal + exchanged 64-bit checksums)."""
    world = 2
    mp.spawn(_record_worker, args=(world, _free_port(), B, K, str(tmp_path)), nprocs=world, join=True)
    d0, d1 = np.load(tmp_path / "rdet0.npy"), np.load(tmp_path / "rdet1.npy")
    c0, c1 = np.load(tmp_path / "rcnt0.npy"), np.load(tmp_path / "rcnt1.npy")
    assert np.array_equal(d0, d1) and np.array_equal(c0, c1) and d0.shape == (world * B, K, 6)
    for g in range(world * B):
        n = g % (K + 1)
        assert c0[g] == n and np.all(d0[g, :n] == g + 0.25) and np.all(d0[g, n:] == -1)
    for r in range(world):
        ok, caught = np.load(tmp_path / f"rok{r}.npy")
        assert ok and caught


def test_record_views_alias_one_buffer():
    from masklab_b200 import dist as mdist
    for B, K in ((32, 100), (1, 1), (5, 7), (32, 1000)):
        dw, words = mdist.record_words(B, K)
        assert dw % 4 == 0 and words % 4 == 0 and dw >= B * K * 6 and words >= dw + B
        rec = torch.zeros((words,), dtype=torch.int32)
        det, counts = mdist.record_views(rec, B, K)
        assert det.shape == (B, K, 6) and counts.shape == (B,)
        assert det.data_ptr() == rec.data_ptr() and counts.data_ptr() == rec.data_ptr() + 4 * dw
        det.fill_(1.5)
        counts.fill_(3)
        assert int((rec != 0).sum()) == B * K * 6 + B
        d, c = mdist.unpack_records(mdist.gather_records(rec), B, K)
        assert torch.equal(d, det) and torch.equal(c, counts)
    with pytest.raises(ValueError):
        mdist.record_views(torch.zeros((5,), dtype=torch.int32), 2, 2)


def test_batch_wide_threshold_is_per_call():
    """CropAndPadMask's row filter is reduced over the whole det_outs tensor of ONE call
    (engine/layers/misc.py:366-369).  Sharding by image therefore reproduces the reference run on each
    shard, which differs from the unsharded call when a shard has no confidence above 50 - documented in
    masklab_b200/dist.py; shown here with the oracle."""
    from oracle import masklab_oracle as mo
    rng = np.random.default_rng(3)
    det = np.zeros((2, 2, 6), np.int32)
    det[0] = [[20, 20, 10, 10, 0, 80], [40, 30, 12, 8, 1, 30]]     # image 0: max conf 80
    det[1] = [[25, 25, 10, 10, 0, 40], [45, 35, 12, 8, 1, 20]]     # image 1: nothing above 50
    masks = (rng.random((2, 2, 28, 28)) > 0.5).astype(np.int32)
    whole = mo.crop_and_pad_mask((64, 64), det, masks)
    shard0 = mo.crop_and_pad_mask((64, 64), det[:1], masks[:1])
    shard1 = mo.crop_and_pad_mask((64, 64), det[1:], masks[1:])
    assert np.array_equal(whole[0], shard0[0])                     # threshold 50 either way
    assert not whole[1].any()                                      # unsharded: image 1's rows are filtered (< 50)
    assert shard1[0, 0].any() and shard1[0, 1].any()               # sharded: threshold -100, both rows pasted
