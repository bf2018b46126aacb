"""Multi-GPU plumbing: frames are independent, so a stream of frames is sharded by image
across the ranks (one process per GPU) and the only collective is ONE all_gather of the
fixed-capacity detection records at the end.  Nothing in a2-a14 mixes images
(every map_fn / partition of the reference is keyed by image id), so no data-path collective
exists; masks and RoI crops stay on the GPU that produced them.

torch.distributed is the transport (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist

MAX_BATCH = 32          # per call, engine/layers/misc.py:275


@dataclass
class Shard:
    start: int          # first global frame index of this rank
    count: int          # frames owned by this rank
    padded: int         # frames per rank after padding (equal on all ranks)


def shard_frames(total_frames, world_size, rank):
    """Contiguous image shards, ceil(total/world) frames per rank; trailing ranks may own fewer
    (their tail is padded with empty frames so that all_gather sees equal shapes)."""
    if total_frames < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad sharding arguments")
    per = -(-total_frames // world_size) if total_frames else 0
    start = min(rank * per, total_frames)
    count = max(0, min(per, total_frames - start))
    return Shard(start, count, per)


def chunks(count, max_batch=MAX_BATCH):
    """Split a shard into calls of at most 32 frames (the reference's MoldBatch limit)."""
    out, s = [], 0
    while s < count:
        n = min(max_batch, count - s)
        out.append((s, n))
        s += n
    return out


def gather_detections(det, counts, total_frames=None, group=None):
    """det [B_local,K,6] (+ -1 padding), counts [B_local] on every rank -> (det [B_total,K,6],
    counts [B_total]) on every rank, in global frame order.  B_local must be equal on all ranks
    (pad with `pad_shard`).  One all_gather each; no reduction."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        out_d, out_c = det, counts
    else:
        world = dist.get_world_size(group)
        ds = [torch.empty_like(det) for _ in range(world)]
        cs = [torch.empty_like(counts) for _ in range(world)]
        dist.all_gather(ds, det.contiguous(), group=group)
        dist.all_gather(cs, counts.contiguous(), group=group)
        out_d, out_c = torch.cat(ds, 0), torch.cat(cs, 0)
    if total_frames is not None:
        out_d, out_c = out_d[:total_frames], out_c[:total_frames]
    return out_d, out_c


def pad_shard(det, counts, padded):
    """Pad a rank's records to `padded` frames with empty (-1 / 0) entries."""
    b = det.shape[0]
    if b == padded:
        return det, counts
    pd = torch.full((padded - b,) + tuple(det.shape[1:]), -1, dtype=det.dtype, device=det.device)
    pc = torch.zeros((padded - b,), dtype=counts.dtype, device=counts.device)
    return torch.cat([det, pd], 0), torch.cat([counts, pc], 0)
