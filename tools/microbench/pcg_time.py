"""C2 solve with THB_SOLVER_SCHUR_PCG (used under ncu to capture k_pcg). GPU box only."""
import ctypes as C, sys
sys.path.insert(0, ".")
from pytheiasfm_b200 import capi, synthetic
lib = capi.load_library()
prob, _ = synthetic.config_c2()
o = capi.default_options(lib); o.linear_solver = capi.SOLVER_SCHUR_PCG
for rep in range(2):
    p = prob.copy(); ps = p.struct(); s = capi.ThbBaSummary()
    capi.check(lib.thb_ba_solve(C.byref(ps), C.byref(o), C.byref(s), None))
    d = s.as_dict()
    print("PCG solve: %d LM iterations, %d CG iterations, %.3f ms per linear solve" % (d["num_iterations"], d["num_linear_solver_iterations"], d["ms_solve"] / d["num_linear_solves"]))
