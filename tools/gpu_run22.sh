THB_TV_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-adapter 2>&1 | grep "rounds<\|two-view" | tail -3
timeout 300 python - <<'PY'
import ctypes as C, numpy as np, sys
sys.path.insert(0, ".")
from pytheiasfm_b200 import capi, synthetic
lib = capi.load_library()
scene = synthetic.config_c5()
print(type(scene), [type(x) for x in scene] if isinstance(scene, tuple) else "")
PY
