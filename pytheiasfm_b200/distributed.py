"""One-process-per-GPU front end of the pair-verification path: the pair table of a view graph is dealt block-cyclically
over the ranks of a torch.distributed group, every rank runs its blocks through the C-ABI on its own GPU, and ONE all_gather
of [result records | bit-packed inlier masks] leaves the whole verified table on every rank. No other data-path collective.

    import torch.distributed as dist, pytheiasfm_b200.distributed as ptd
    dist.init_process_group("nccl"); torch.cuda.set_device(local_rank)
    results, masks = ptd.estimate_relative_poses(batch, params)      # batch: capi.HostPairBatch of ALL pairs, same on every rank

`run_local` lets the CPU tests (gloo) replace the device call; the product path always calls libtheia_b200."""
import ctypes as C

import numpy as np

from . import capi, sharding

PAIR_BLOCK = 8


def _device_runner(entry):
    def run(sub, params):
        lib = capi.load_library()
        res = np.zeros(sub.num_pairs, capi.RELPOSE_DTYPE)
        mask = np.zeros(int(sub.pair_offset[-1]), np.uint8)
        if sub.num_pairs:
            b = sub.struct()
            capi.check(getattr(lib, entry)(C.byref(b), C.byref(params), res.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p), None))
        return res, mask
    return run


def pack_masks(mask, pair_offset):
    """Host-side twin of thb_pack_inlier_masks: 32 flags per little-endian word, pairs padded to whole words."""
    woff = sharding.mask_word_offsets(pair_offset)
    words = np.zeros(int(woff[-1]), np.uint32)
    for p in range(len(pair_offset) - 1):
        m = mask[pair_offset[p]:pair_offset[p + 1]]
        bits = np.zeros(int(woff[p + 1] - woff[p]) * 32, np.uint8)
        bits[: len(m)] = m != 0
        words[woff[p]:woff[p + 1]] = np.packbits(bits, bitorder="little").view(np.uint32)
    return words, woff


def run_pairs_sharded(batch, params, run_local, group=None, device=None, block=PAIR_BLOCK):
    """batch: capi.HostPairBatch with the pairs of the whole view graph (identical on every rank). Returns (records
    [num_pairs] in pair order, list of per-pair uint8 inlier masks)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    idx = sharding.block_cyclic_indices(batch.num_pairs, rank, world, block)
    sub = capi.HostPairBatch([batch.corr[batch.pair_offset[i]:batch.pair_offset[i + 1]] for i in idx], batch.seed[idx], width=batch.width)
    res, mask = run_local(sub, params)
    words, _ = pack_masks(mask, sub.pair_offset)
    rec = capi.RELPOSE_DTYPE.itemsize
    payload = np.concatenate([res.view(np.uint8).reshape(-1), words.view(np.uint8)])
    sizes = np.diff(batch.pair_offset)
    counts = []
    for r in range(world):  # every rank knows every rank's payload size from the pair table: no size exchange
        ridx = sharding.block_cyclic_indices(batch.num_pairs, r, world, block)
        counts.append(len(ridx) * rec + 4 * int(((sizes[ridx] + 31) // 32).sum()))
    assert counts[rank] == payload.size
    local = torch.from_numpy(payload.copy())
    if device is not None:
        local = local.to(device)
    parts = sharding.all_gather_padded(local, counts, group) if world > 1 else [local]
    records = np.zeros(batch.num_pairs, capi.RELPOSE_DTYPE)
    masks = [None] * batch.num_pairs
    for r in range(world):
        ridx = sharding.block_cyclic_indices(batch.num_pairs, r, world, block)
        raw = parts[r].cpu().numpy()
        records[ridx] = np.frombuffer(raw[: len(ridx) * rec].tobytes(), capi.RELPOSE_DTYPE)
        w = np.frombuffer(raw[len(ridx) * rec:].tobytes(), np.uint32)
        o = 0
        for i in ridx:
            nw = (int(sizes[i]) + 31) // 32
            masks[i] = sharding.unpack_mask_words(w[o:o + nw], int(sizes[i]))
            o += nw
    return records, masks


def estimate_relative_poses(batch, params, group=None):
    """EstimateRelativePose for every pair of the view graph over the ranks of `group` (estimate_relative_pose.cc:159-172)."""
    import torch
    return run_pairs_sharded(batch, params, _device_runner("thb_ransac_relpose_batch"), group, torch.device("cuda", torch.cuda.current_device()))


def estimate_homographies(batch, params, group=None):
    """EstimateHomography for every pair (estimate_homography.cc:122-136) over the ranks of `group`."""
    import torch
    return run_pairs_sharded(batch, params, _device_runner("thb_ransac_homography_batch"), group, torch.device("cuda", torch.cuda.current_device()))
