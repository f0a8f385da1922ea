"""Times thb_ransac_relpose_batch for small batches (device-resident inputs) in the mode THB_RANSAC_MODE selects. GPU box only."""
import ctypes as C, os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from pytheiasfm_b200 import capi, synthetic
lib = capi.load_library()
params = capi.ThbRansacParams(); lib.thb_ransac_default_params(C.byref(params)); params = synthetic.c4_params(params)
for npairs in (1, 4, 16, 64, 148, 444, 1000):
    batch, _ = synthetic.make_pair_batch(npairs, n=2000, seed=3, base_seed=99)
    off = torch.from_numpy(batch.pair_offset).cuda(); corr = torch.from_numpy(batch.corr).cuda(); seed = torch.from_numpy(batch.seed).cuda()
    res = torch.zeros(npairs * capi.RELPOSE_DTYPE.itemsize, dtype=torch.uint8, device="cuda"); mask = torch.zeros(int(batch.pair_offset[-1]), dtype=torch.uint8, device="cuda")
    b = capi.ThbPairBatch(npairs, capi.THB_MEM_DEVICE, off.data_ptr(), corr.data_ptr(), seed.data_ptr())
    def run():
        capi.check(lib.thb_ransac_relpose_batch(C.byref(b), C.byref(params), C.c_void_p(res.data_ptr()), C.c_void_p(mask.data_ptr()), None))
    for _ in range(3): run()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    reps = 10 if npairs <= 148 else 4
    for _ in range(reps): run()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    print("%s pairs=%4d  %.3f ms per call  %.0f pairs/s" % (os.environ.get("THB_RANSAC_MODE", "default"), npairs, dt * 1e3, npairs / dt))
