"""Pair-queue sharding for multi-GPU RANSAC verification.

Image pairs are independent units (the reference fans them out over a thread pool,
matching/feature_matcher.cc:117-126). One process per GPU; the pair table is split into contiguous blocks balanced by
cumulative correspondence count (the cost of a pair is proportional to its number of correspondences times its
iteration count); every rank verifies its block with thb_ransac_*_batch; one all_gather of the fixed-size result
records assembles the full table on every rank. No collective sits on the data path of the kernels.
Tracks (per-track BA, triangulation) shard the same way: shard_tracks.
"""
import numpy as np

from . import capi


def partition_by_work(pair_offset, world):
    """Contiguous [lo, hi) pair ranges, one per rank, with near-equal correspondence counts. Every pair is covered
    exactly once; ranks may be empty when there are fewer pairs than ranks."""
    pair_offset = np.asarray(pair_offset, dtype=np.int64)
    num_pairs = len(pair_offset) - 1
    total = int(pair_offset[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        # first pair whose END offset reaches the target; keep the sequence non-decreasing
        idx = int(np.searchsorted(pair_offset[1:], target, side="left")) + 1 if total > 0 else num_pairs * r // world
        bounds.append(min(max(idx, bounds[-1]), num_pairs))
    bounds.append(num_pairs)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def block_cyclic_indices(num_units, rank, world, block=8):
    """Over-decomposed static schedule: the table is cut into blocks of `block` consecutive units and block b goes to
    rank b % world. A pair's cost (its RANSAC iteration count) is unknown before it runs and differs by 100x between
    pairs; dealing many small blocks round-robin balances the ranks statistically without any communication, and the
    per-GPU atomic pair counter of k_ransac balances the SMs inside a rank. Returns the rank's global unit indices,
    ascending."""
    idx = np.arange(num_units, dtype=np.int64)
    return idx[(idx // block) % world == rank]


def mask_word_offsets(pair_offset):
    """word_offset [num_pairs + 1] for thb_pack_inlier_masks: ceil(n_p / 32) 32-bit words per pair."""
    n = np.diff(np.asarray(pair_offset, dtype=np.int64))
    out = np.zeros(len(n) + 1, np.int64)
    out[1:] = np.cumsum((n + 31) // 32)
    return out


def unpack_mask_words(words, n):
    """Inverse of thb_pack_inlier_masks for one pair: `n` flags from ceil(n / 32) little-endian words."""
    bits = np.unpackbits(np.ascontiguousarray(words, dtype=np.uint32).view(np.uint8), bitorder="little")
    return bits[:n]


def shard_batch(batch, rank, world):
    """The rank's block of a HostPairBatch (a new HostPairBatch with re-based offsets) and its [lo, hi) range."""
    lo, hi = partition_by_work(batch.pair_offset, world)[rank]
    o0, o1 = int(batch.pair_offset[lo]), int(batch.pair_offset[hi])
    sub = capi.HostPairBatch.__new__(capi.HostPairBatch)
    sub.width = batch.width
    sub.num_pairs = hi - lo
    sub.pair_offset = (batch.pair_offset[lo: hi + 1] - o0).astype(np.int64)
    sub.corr = np.ascontiguousarray(batch.corr[o0:o1])
    sub.seed = np.ascontiguousarray(batch.seed[lo:hi])
    return sub, lo, hi


def shard_tracks(prob, rank, world):
    """The rank's block of tracks of a HostBaProblem for thb_ba_tracks_batch / thb_triangulate_midpoint_batch (SURVEY 8e,
    "per-track BA / triangulation: tracks independent given fixed cameras - same scheme as pairs"): contiguous point range
    [lo, hi) balanced by cumulative observation count, the observations of those points with obs_pt re-based, and ALL
    cameras / intrinsics (read-only, replicated on every rank). Returns (HostBaProblem, lo, hi)."""
    obs_pt = prob.a["obs_pt"]
    counts = np.bincount(obs_pt, minlength=prob.num_points)
    offset = np.zeros(prob.num_points + 1, np.int64)
    offset[1:] = np.cumsum(counts)
    lo, hi = partition_by_work(offset, world)[rank]
    keep = (obs_pt >= lo) & (obs_pt < hi)
    a = dict(prob.a)
    a["pts"] = prob.a["pts"][lo:hi]
    a["pt_const"] = None if prob.a["pt_const"] is None else prob.a["pt_const"][lo:hi]
    for k in ("obs_cam", "obs_xy", "obs_sqrt_info"):
        a[k] = None if prob.a[k] is None else prob.a[k][keep]
    a["obs_pt"] = obs_pt[keep] - lo
    return capi.HostBaProblem(a), lo, hi


def all_gather_padded(local, counts, group=None):
    """all_gather of per-rank 1-D tensors of different lengths (counts[r] elements on rank r): one collective on buffers
    padded to the longest. Returns the list of per-rank tensors, trimmed."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    pad = int(max(counts))
    buf = torch.zeros(pad, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local.reshape(-1)
    out = torch.empty(world * pad, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return [out[r * pad: r * pad + int(counts[r])] for r in range(world)]


def assemble_block_cyclic(per_rank, num_units, world, block=8, width=1):
    """Global table (unit-major, `width` elements per unit) from the per-rank pieces of a block_cyclic_indices schedule."""
    import torch
    full = torch.empty(num_units * width, dtype=per_rank[0].dtype, device=per_rank[0].device)
    view = full.view(num_units, width)
    for r in range(world):
        idx = torch.from_numpy(block_cyclic_indices(num_units, r, world, block)).to(full.device)
        view[idx] = per_rank[r].view(-1, width)
    return full


def all_gather_results(local_records, ranges, group=None, device=None, dtype=None):
    """all_gather of per-unit result records (structured array of `dtype`, default capi.RELPOSE_DTYPE, or a uint8 torch
    tensor on the rank's device) into the full table, ordered by global pair / track index. Works on NCCL (GPU) and gloo
    (CPU tests)."""
    import torch
    import torch.distributed as dist
    dtype = capi.RELPOSE_DTYPE if dtype is None else np.dtype(dtype)
    rec = dtype.itemsize
    world = dist.get_world_size(group)
    counts = [hi - lo for lo, hi in ranges]
    if isinstance(local_records, np.ndarray):
        local = torch.from_numpy(local_records.view(np.uint8).reshape(-1).copy())
        if device is not None:
            local = local.to(device)
    else:
        local = local_records.reshape(-1)
    assert local.numel() == counts[dist.get_rank(group)] * rec
    pad = max(counts) * rec
    buf = torch.zeros(pad, dtype=torch.uint8, device=local.device)
    buf[: local.numel()] = local
    out = [torch.zeros(pad, dtype=torch.uint8, device=local.device) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    full = torch.cat([out[r][: counts[r] * rec] for r in range(world)])
    return np.frombuffer(full.cpu().numpy().tobytes(), dtype)
