"""The multi-GPU path on real GPUs: two processes, one GPU each, NCCL.  Every rank runs
the fused pipeline on its own image shard, the packed detection records are gathered with ONE
all_gather_into_tensor, and what arrives is checked three ways: bit-equal to what each rank sent (own slice +
exchanged checksums, masklab_b200.dist.verify_gather), equal to the unsharded run of the same frames on one
GPU, and equal to the C restatement of the reference path frame by frame.  Skipped on a one-GPU box (the
gloo tests of tests/test_dist.py cover the host logic there)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import synth

pytestmark = pytest.mark.gpu

KW = dict(min_confidence=0.05, nms_iou_threshold=0.4, post_iou_threshold=0.65, nms_max_output_size=40)
SHAPE = dict(H=128, W=256, C=4, Cf=16, B=3)              # B frames per rank


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(total):
    cfgp = synth.prior_config()
    N = synth.num_anchors(cfgp, SHAPE["H"], SHAPE["W"])
    loc, cls = synth.head_tensors(total, N, SHAPE["C"], mu=-3.5, seed=901)
    cls[1] = 0                                          # a frame without detections
    fmaps = synth.fpn_maps(total, SHAPE["H"], SHAPE["W"], SHAPE["Cf"], seed=902)
    return cfgp, loc, cls, fmaps


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import masklab_b200 as ml
    from masklab_b200 import dist as mdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    B = SHAPE["B"]
    cfgp, loc, cls, fmaps = _inputs(world * B)
    sh = mdist.shard_frames(world * B, world, rank)
    sl = slice(sh.start, sh.start + sh.count)
    pipe = ml.PostProcessPipeline(cfgp, (SHAPE["H"], SHAPE["W"]), (SHAPE["H"], SHAPE["W"]), SHAPE["C"], SHAPE["Cf"], B,
                                  ml.DetectionConfig(max_k=2, base_size=36, **KW), device=rank, private_context=True)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    pipe.detect_and_align(d(loc[sl]), d(cls[sl]), [d(f[sl]) for f in fmaps])
    gathered = mdist.gather_records(pipe.record)                    # ONE collective, straight from the NMS output
    ok = mdist.verify_gather(gathered, pipe.record)
    det, counts = mdist.unpack_records(gathered, B, pipe.K)
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"det{rank}.npy"), det.cpu().numpy())
    np.save(os.path.join(out_dir, f"cnt{rank}.npy"), counts.cpu().numpy())
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([ok]))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_nccl_gather_of_packed_records_two_gpus(tmp_path):
    import masklab_b200 as ml
    from oracle import c_oracle as co
    world, B = 2, SHAPE["B"]
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    d0, d1 = np.load(tmp_path / "det0.npy"), np.load(tmp_path / "det1.npy")
    c0, c1 = np.load(tmp_path / "cnt0.npy"), np.load(tmp_path / "cnt1.npy")
    assert np.load(tmp_path / "ok0.npy")[0] and np.load(tmp_path / "ok1.npy")[0]
    assert np.array_equal(d0, d1) and np.array_equal(c0, c1) and d0.shape == (world * B, KW["nms_max_output_size"], 6)
    # the C restatement of the reference path, frame by frame
    cfgp, loc, cls, fmaps = _inputs(world * B)
    prior = co.prior_layer(cfgp, SHAPE["H"], SHAPE["W"])
    for g in range(world * B):
        boxes = co.restore_boxes(loc[g:g + 1], prior)
        want = co.detection_proposal(cls[g:g + 1], boxes, **KW)[0]
        n = int((want[:, 4] != -1).sum())
        assert c0[g] == n and np.array_equal(d0[g, :n], want[:n]) and np.all(d0[g, n:] == -1)
    assert c0[1] == 0
    # and the unsharded run of the same six frames on one GPU (detection mixes no images)
    pipe = ml.PostProcessPipeline(cfgp, (SHAPE["H"], SHAPE["W"]), (SHAPE["H"], SHAPE["W"]), SHAPE["C"], SHAPE["Cf"],
                                  world * B, ml.DetectionConfig(max_k=2, base_size=36, **KW), private_context=True)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    rois = pipe.detect_and_align(d(loc), d(cls), [d(f) for f in fmaps])
    assert np.array_equal(rois.det.cpu().numpy(), d0) and np.array_equal(rois.counts.cpu().numpy(), c0)
