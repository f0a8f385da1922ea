timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12
cat > /tmp/lo.py <<'PY'
import sys, ctypes as C; sys.path.insert(0,'.')
import numpy as np
from pytheiasfm_b200 import synthetic, capi
lib=capi.load_library()
batch,_=synthetic.make_pair_batch_indexed(range(6), n=400, seed=9)
p=synthetic.c4_params(capi.ThbRansacParams()); p.use_lo=1; p.lo_start_iterations=0
res=np.zeros(6,capi.RELPOSE_DTYPE); mask=np.zeros(int(batch.pair_offset[-1]),np.uint8)
b=batch.struct()
capi.check(lib.thb_ransac_relpose_batch(C.byref(b),C.byref(p),res.ctypes.data_as(C.c_void_p),mask.ctypes.data_as(C.c_void_p),None))
print(res["num_iterations"], res["num_lo_iterations"])
PY
timeout 900 compute-sanitizer --tool racecheck python /tmp/lo.py 2>&1 | grep -E "RACECHECK|Race|hazard|^\[" | cut -c1-200 | head -12
timeout 900 compute-sanitizer --tool memcheck python /tmp/lo.py 2>&1 | grep -E "ERROR SUMMARY|Invalid|^\[" | head
