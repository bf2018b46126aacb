"""Multi-GPU plumbing: frames are independent, so a stream of frames is sharded by image
across the ranks (one process per GPU) and the only collective is ONE all_gather of the
fixed-capacity detection records at the end.  Every map_fn / partition of
a2-a12 is keyed by image id, so no data-path collective exists; masks and RoI crops stay
on the GPU that produced them.

Two reductions of the reference ARE batch-wide and are evaluated per CALL (per local batch
of <= 32 frames) here, exactly as the reference evaluates them per session.run:
CropAndPadMask's row filter `50 if max(conf over the whole det_outs tensor) > 50 else -100`
(engine/layers/misc.py:366-369) and CrackToInstance's crack box / M' = M + 1 decision
(misc.py:546-590).  Sharding a batch across ranks, or chunking a shard into calls of 32,
therefore gives what the reference gives when it is fed those same calls - not what it
would give for one unsharded call (tests/test_dist.py::test_batch_wide_threshold_is_per_call
shows the difference on a batch where one shard has no confidence above 50).  The serving
graph runs batch 1 (retinamasklab.py:601), where the two coincide.

The records travel as ONE contiguous int32 buffer per rank - det [B,K,6] (float bits) followed
by counts [B] - which the cross-class NMS kernel writes in place (PostProcessPipeline.record),
so the gather is a single all_gather_into_tensor with no staging copy.

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


# ---- packed detection records: one buffer, one collective ---------------------------------
def record_words(batch, keep):
    """int32 words of a packed record: det [B,K,6] padded to a multiple of 4 words (so that the
    counts that follow stay 16-byte aligned), then counts [B]."""
    det_words = (batch * keep * 6 + 3) // 4 * 4
    return det_words, det_words + (batch + 3) // 4 * 4


def record_views(record, batch, keep):
    """(det f32 [B,K,6], counts i32 [B]) views into a packed record (no copy)."""
    det_words, total = record_words(batch, keep)
    if record.dtype != torch.int32 or record.numel() != total:
        raise ValueError(f"record must be int32 [{total}], got {record.dtype} [{record.numel()}]")
    det = record[:batch * keep * 6].view(torch.float32).view(batch, keep, 6)
    return det, record[det_words:det_words + batch]


def gather_records(record, out=None, group=None):
    """record int32 [W] on every rank -> [world, W] on every rank: ONE all_gather_into_tensor."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return record.unsqueeze(0) if out is None else out.copy_(record.unsqueeze(0))
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world, record.numel()), dtype=record.dtype, device=record.device)
    dist.all_gather_into_tensor(out.view(-1), record.view(-1), group=group)     # flat: gloo insists on 1-D
    return out


def unpack_records(gathered, batch, keep, total_frames=None):
    """[world, W] -> (det [world*B,K,6] f32, counts [world*B] i32) in global frame order."""
    det_words, _ = record_words(batch, keep)
    world = gathered.shape[0]
    det = gathered[:, :batch * keep * 6].contiguous().view(torch.float32).view(world * batch, keep, 6)
    counts = gathered[:, det_words:det_words + batch].contiguous().view(world * batch)
    if total_frames is not None:
        det, counts = det[:total_frames], counts[:total_frames]
    return det, counts


def record_checksum(record):
    """64-bit position-weighted checksum of an int32 buffer (a permutation or a dropped word changes it)."""
    w = record.reshape(-1).to(torch.int64) & 0xffffffff
    idx = torch.arange(1, w.numel() + 1, dtype=torch.int64, device=w.device)
    return (w * (idx % 65521 + 1)).sum() & 0x7fffffffffffffff


def verify_gather(gathered, record, group=None):
    """True on every rank iff what the collective delivered is what the ranks sent: this rank's slice
    equals its own record bit for bit, and the checksum of every slice equals the checksum its owner
    computed locally (one tiny all_gather of 64-bit sums)."""
    world = gathered.shape[0]
    rank = dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0
    ok = bool(torch.equal(gathered[rank], record))
    mine = record_checksum(record).reshape(1)
    if world > 1:
        sums = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine, group=group)
    else:
        sums = [mine]
    for r in range(world):
        ok = ok and int(record_checksum(gathered[r])) == int(sums[r])
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=record.device)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(int(flag) == 1)
