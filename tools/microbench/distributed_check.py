"""torchrun --nproc-per-node N tools/microbench/distributed_check.py: the public sharded front end on NCCL against the single-GPU call."""
import ctypes as C, os, sys
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from pytheiasfm_b200 import capi, synthetic, distributed as ptd
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
lib = capi.load_library()
batch, _ = synthetic.make_pair_batch(203, n=400, seed=4, base_seed=321)
params = synthetic.c4_params(capi.ThbRansacParams())
records, masks = ptd.estimate_relative_poses(batch, params)
res = np.zeros(batch.num_pairs, capi.RELPOSE_DTYPE); mask = np.zeros(int(batch.pair_offset[-1]), np.uint8)
b = batch.struct()
capi.check(lib.thb_ransac_relpose_batch(C.byref(b), C.byref(params), res.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p), None))
ok = records.tobytes() == res.tobytes() and all(np.array_equal(masks[i], mask[batch.pair_offset[i]:batch.pair_offset[i + 1]]) for i in range(batch.num_pairs))
print("rank %d of %d: sharded table identical to the single-GPU call: %s" % (rank, dist.get_world_size(), ok))
dist.destroy_process_group()
assert ok
