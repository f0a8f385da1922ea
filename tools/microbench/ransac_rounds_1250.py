"""1250 C4-shaped pairs (the per-rank share at 8 GPUs) in both schedules, with the per-phase times of the round schedule. GPU box only."""
import ctypes as C, os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from pytheiasfm_b200 import capi, synthetic
lib = capi.load_library()
params = capi.ThbRansacParams(); lib.thb_ransac_default_params(C.byref(params)); params = synthetic.c4_params(params)
npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
batch, _ = synthetic.make_pair_batch(npairs, n=2000, seed=3, base_seed=99)
off = torch.from_numpy(batch.pair_offset).cuda(); corr = torch.from_numpy(batch.corr).cuda(); seed = torch.from_numpy(batch.seed).cuda()
res = torch.zeros(npairs * capi.RELPOSE_DTYPE.itemsize, dtype=torch.uint8, device="cuda"); mask = torch.zeros(int(batch.pair_offset[-1]), dtype=torch.uint8, device="cuda")
b = capi.ThbPairBatch(npairs, capi.THB_MEM_DEVICE, off.data_ptr(), corr.data_ptr(), seed.data_ptr())
def run():
    capi.check(lib.thb_ransac_relpose_batch(C.byref(b), C.byref(params), C.c_void_p(res.data_ptr()), C.c_void_p(mask.data_ptr()), None))
for mode in ("rounds", "fused"):
    os.environ["THB_RANSAC_MODE"] = mode
    os.environ.pop("THB_TV_TIMING", None)
    for _ in range(3): run()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): run()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print("%s pairs=%d  %.3f ms per call  %.0f pairs/s" % (mode, npairs, dt * 1e3, npairs / dt), flush=True)
    if mode == "rounds":
        os.environ["THB_TV_TIMING"] = "1"; run(); torch.cuda.synchronize()
