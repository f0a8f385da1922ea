set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py -x -q -k "pcg or invalid_arguments" 2>&1 | tail -30
timeout 300 python - <<'PY'
import ctypes as C, numpy as np
from pytheiasfm_b200 import capi, synthetic
lib = capi.load_library()
prob, _ = synthetic.config_c2()
for solver in (0, 1):
    o = capi.default_options(lib); o.linear_solver = solver
    for rep in range(2):
        p = prob.copy(); ps = p.struct(); s = capi.ThbBaSummary()
        rc = lib.thb_ba_solve(C.byref(ps), C.byref(o), C.byref(s), None)
        d = s.as_dict()
        print(solver, rc, d["num_iterations"], d["num_linear_solves"], d["num_linear_solver_iterations"], "ms_solve/solve %.3f" % (d["ms_solve"] / max(1, d["num_linear_solves"])),
              "normal %.3f" % (d["ms_normal"] / max(1, d["num_linear_solves"])), "solve_time %.1f ms" % (1e3 * d["solve_time_in_seconds"]), ["%.6e" % c for c in d["iter_cost"]])
PY
