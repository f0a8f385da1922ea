"""C3 (BASELINE configs[2]) timing: complete solves with intrinsics refinement on the GPU vs the oracle. GPU box only."""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np
from pytheiasfm_b200 import capi, synthetic
from oracle import oracle_py
lib = capi.load_library()
prob, _ = synthetic.config_c3()
prob.a["intr"][:, 0] *= 1.02
for rep in range(3):
    p = prob.copy(); s = capi.ThbBaSummary(); o = capi.default_options(lib); o.max_num_iterations = 50
    ps = p.struct(); t0 = time.time()
    capi.check(lib.thb_ba_solve(C.byref(ps), C.byref(o), C.byref(s), None)); dt = time.time() - t0
    d = s.as_dict()
    print("gpu: %d iterations in %.1f ms wall (%.1f it/s), setup %.2f ms, phases/it: jac %.3f normal %.3f solve %.3f update %.3f ms, cost %.6g -> %.6g" % (
        d["num_iterations"], 1e3 * dt, d["num_iterations"] / dt, 1e3 * d["setup_time_in_seconds"], d["ms_jacobian"] / d["num_iterations"],
        d["ms_normal"] / d["num_iterations"], d["ms_solve"] / d["num_iterations"], d["ms_update"] / d["num_iterations"], d["initial_cost"], d["final_cost"]))
if "--oracle" in sys.argv:
    p = prob.copy(); o = oracle_py.default_options(); o.max_num_iterations = 3
    t0 = time.time(); r = oracle_py.ba_solve(p, o); dt = time.time() - t0
    print("oracle: %d iterations in %.1f s (%.2f it/s)" % (r["num_iterations"], dt, r["num_iterations"] / dt))
