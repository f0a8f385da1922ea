"""Multi-process (gloo, world_size 2 and 3, CPU) tests of the pair-queue sharding used for multi-GPU verification."""
import os
import socket

import numpy as np
import pytest

from pytheiasfm_b200 import capi, sharding, synthetic


def test_partition_covers_every_pair_once_and_balances_work():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 4, 8):
        for _ in range(20):
            sizes = rng.integers(0, 3000, size=int(rng.integers(0, 40)))
            off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
            parts = sharding.partition_by_work(off, world)
            assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == len(sizes)
            assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            assert all(lo <= hi for lo, hi in parts)
            if len(sizes) >= 4 * world and sizes.sum() > 0:
                work = [int(off[hi] - off[lo]) for lo, hi in parts]
                assert max(work) - min(work) <= 2 * sizes.max()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        corrs = [rng.uniform(-1, 1, (int(n), 4)) for n in rng.integers(5, 400, size=23)]
        batch = capi.HostPairBatch(corrs, 100 + np.arange(23))
        ranges = sharding.partition_by_work(batch.pair_offset, world)
        sub, lo, hi = sharding.shard_batch(batch, rank, world)
        assert (lo, hi) == ranges[rank] and sub.num_pairs == hi - lo
        np.testing.assert_array_equal(sub.seed, batch.seed[lo:hi])
        np.testing.assert_array_equal(sub.corr, batch.corr[batch.pair_offset[lo]: batch.pair_offset[hi]])
        # stand-in for the device results of this rank's block: records tagged with the global pair index
        local = np.zeros(hi - lo, capi.RELPOSE_DTYPE)
        local["success"] = 1
        local["num_iterations"] = np.arange(lo, hi)
        local["num_input_data_points"] = np.diff(sub.pair_offset)
        full = sharding.all_gather_results(local, ranges)
        assert len(full) == 23
        np.testing.assert_array_equal(full["num_iterations"], np.arange(23))
        np.testing.assert_array_equal(full["num_input_data_points"], np.diff(batch.pair_offset))
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shard_and_gather_over_gloo(world):
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_track_shards_cover_every_track_and_observation_once():
    from pytheiasfm_b200 import synthetic
    prob, _ = synthetic.make_ba_problem(9, 301, 4, seed=3)
    prob.a["pt_const"][::7] = 1
    for world in (1, 2, 3, 8):
        seen_pts = np.zeros(prob.num_points, int); n_obs = 0; work = []
        for rank in range(world):
            sub, lo, hi = sharding.shard_tracks(prob, rank, world)
            seen_pts[lo:hi] += 1
            n_obs += sub.num_observations
            work.append(sub.num_observations)
            assert sub.num_points == hi - lo and sub.num_cameras == prob.num_cameras
            np.testing.assert_array_equal(sub.a["pts"], prob.a["pts"][lo:hi])
            np.testing.assert_array_equal(sub.a["pt_const"], prob.a["pt_const"][lo:hi])
            keep = (prob.a["obs_pt"] >= lo) & (prob.a["obs_pt"] < hi)
            np.testing.assert_array_equal(sub.a["obs_pt"] + lo, prob.a["obs_pt"][keep])
            np.testing.assert_array_equal(sub.a["obs_xy"], prob.a["obs_xy"][keep])
        assert (seen_pts == 1).all() and n_obs == prob.num_observations
        assert max(work) - min(work) <= 8          # balanced to within a couple of tracks


def _track_worker(rank, world, port, ret):
    import torch.distributed as dist
    from pytheiasfm_b200 import synthetic
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prob, _ = synthetic.make_ba_problem(6, 57, 3, seed=4)
        counts = np.bincount(prob.a["obs_pt"], minlength=57)
        offset = np.concatenate([[0], np.cumsum(counts)])
        ranges = sharding.partition_by_work(offset, world)
        sub, lo, hi = sharding.shard_tracks(prob, rank, world)
        assert (lo, hi) == ranges[rank]
        # stand-in for this rank's thb_ba_tracks_batch results: records tagged with the global track index
        local = np.zeros(hi - lo, capi.TRACK_BA_DTYPE)
        local["num_iterations"] = np.arange(lo, hi)
        local["final_cost"] = sub.a["pts"][:, 0]
        full = sharding.all_gather_results(local, ranges, dtype=capi.TRACK_BA_DTYPE)
        np.testing.assert_array_equal(full["num_iterations"], np.arange(57))
        np.testing.assert_array_equal(full["final_cost"], prob.a["pts"][:, 0])
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_track_shard_and_gather_over_gloo():
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_track_worker, args=(2, port, ret), nprocs=2, join=True)
    assert all(ret.get(r) for r in range(2))


def _worker_pairs(rank, world, port, ret):
    import torch.distributed as dist
    from oracle import oracle_py
    from pytheiasfm_b200 import distributed as ptd, synthetic
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch, _ = synthetic.make_pair_batch(21, n=150, seed=9, base_seed=500)
        params = synthetic.c4_params(oracle_py.ransac_default_params())

        def run_local(sub, p):  # the CPU stand-in of the device call: the oracle on this rank's blocks
            rc, res, mask = oracle_py.ransac_relpose_batch(sub, p)
            assert rc == 0
            return res, mask
        records, masks = ptd.run_pairs_sharded(batch, params, run_local)
        rc, full, full_mask = oracle_py.ransac_relpose_batch(batch, params)
        assert records.tobytes() == full.tobytes()
        for i in range(batch.num_pairs):
            np.testing.assert_array_equal(masks[i], full_mask[batch.pair_offset[i]:batch.pair_offset[i + 1]])
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_pair_verification_front_end_over_gloo(world):
    """pytheiasfm_b200.distributed.run_pairs_sharded: block-cyclic deal, local runs, one all_gather of records + bit-packed
    masks; every rank ends with the table the single-process call gives (the oracle stands in for the device call on CPU)."""
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_pairs, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
