"""Times K4 (dense FP64 Cholesky factor + solve) alone on a device-built SPD system. GPU box only."""
import ctypes as C
import sys
sys.path.insert(0, ".")
from pytheiasfm_b200 import capi
lib = capi.load_library()
for n in [int(a) for a in sys.argv[1:]] or [6000]:
    ms, res = C.c_double(0.0), C.c_double(0.0)
    capi.check(lib.thb_dense_spd_time(n, 3, C.byref(ms), C.byref(res), None))   # warm-up
    capi.check(lib.thb_dense_spd_time(n, 10, C.byref(ms), C.byref(res), None))
    print("n=%d K4 %.3f ms  %.2f TFLOP/s (n^3/3)  residual %.2e" % (n, ms.value, n ** 3 / 3 / ms.value / 1e9, res.value))
