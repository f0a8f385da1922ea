timeout 300 python - <<'PY'
import ctypes as C, numpy as np, sys, time
sys.path.insert(0, ".")
from pytheiasfm_b200 import capi, synthetic
lib = capi.load_library()
batch, i1, i2, gts = synthetic.make_two_view_batch(512, n=1886, seed=3)
o = capi.ThbTwoViewOptions(); lib.thb_two_view_default_options(C.byref(o))
info = np.zeros(batch.num_pairs, capi.TWO_VIEW_INFO_DTYPE); mask = np.zeros(int(batch.pair_offset[-1]), np.uint8)
b = batch.struct()
for rep in range(3):
    t0 = time.perf_counter()
    capi.check(lib.thb_verify_two_view_matches_batch(C.byref(b), i1.ctypes.data_as(C.c_void_p), i2.ctypes.data_as(C.c_void_p), C.byref(o), info.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p), None))
    print("call %.1f ms" % (1e3 * (time.perf_counter() - t0)))
it = info["ba_iterations"]
print("ba_iterations: mean %.1f max %d hist" % (it.mean(), it.max()), np.bincount(it)[:30], "triangulated mean", info["num_triangulated"].mean(), "verified", info["num_verified_matches"].mean(), "success", info["success"].mean())
PY
