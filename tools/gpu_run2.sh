timeout 300 python -m pytest tests/test_ba_gpu.py -m gpu -x -q -k "k4 or c2_full or c1_full" 2>&1 | tail -3
timeout 120 python tools/microbench/k4_time.py 1000 2500 6000
THB_K4_PROF=1 timeout 120 python - <<'PY' 2>&1 | grep K4PROF | tail -8
import ctypes as C, sys
sys.path.insert(0, ".")
from pytheiasfm_b200 import capi
lib = capi.load_library()
ms, res = C.c_double(0.0), C.c_double(0.0)
capi.check(lib.thb_dense_spd_time(1000, 1, C.byref(ms), C.byref(res), None))
PY
