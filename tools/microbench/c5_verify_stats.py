"""Statistics of the C5 verification batch (what the two-view BA stage spends its time on). GPU box only."""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np
from pytheiasfm_b200 import capi, synthetic
lib = capi.load_library()
scene = synthetic.config_c5()
pairs, intr = scene["pairs"], scene["intrinsics"]
o = capi.ThbTwoViewOptions(); lib.thb_two_view_default_options(C.byref(o))
info = np.zeros(pairs.num_pairs, capi.TWO_VIEW_INFO_DTYPE); mask = np.zeros(int(pairs.pair_offset[-1]), np.uint8)
b = pairs.struct()
mi = np.ascontiguousarray(intr)
for rep in range(2):
    t0 = time.perf_counter()
    capi.check(lib.thb_verify_two_view_matches_batch(C.byref(b), mi.ctypes.data_as(C.c_void_p), mi.ctypes.data_as(C.c_void_p), C.byref(o),
                                                     info.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p), None))
    print("call %.1f ms" % (1e3 * (time.perf_counter() - t0)))
it = info["ba_iterations"]
print("success %.3f  ba_iterations mean %.1f max %d  hist(>=0)" % (info["success"].mean(), it.mean(), it.max()), np.bincount(np.clip(it, 0, 100))[:101].nonzero()[0].tolist())
print("counts per iteration value:", {int(k): int(v) for k, v in zip(*np.unique(it, return_counts=True))})
print("triangulated mean %.0f  verified mean %.0f  ransac iterations mean %.0f max %d" % (info["num_triangulated"].mean(), info["num_verified_matches"].mean(), info["num_ransac_iterations"].mean(), info["num_ransac_iterations"].max()))
