"""A/B timing of the two K1 kernels on C2 (thb_ba_time_jacobian: CUDA events on the launch stream, L2 read-flush between
launches). Usage: python tools/microbench/k1_ab.py [mode ...]  (spec = gather | shared)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytheiasfm_b200 import capi, synthetic

lib = capi.load_library()
# THB_K1_AB_MODEL=<THB_MODEL_* id>, THB_K1_AB_LOSS=<THB_LOSS_* id>, THB_K1_AB_EUCLID=1: C2-shaped problem on another instantiation
MODEL = int(os.environ.get("THB_K1_AB_MODEL", capi.MODEL_PINHOLE)); LOSS = int(os.environ.get("THB_K1_AB_LOSS", capi.LOSS_TRIVIAL))
if MODEL == capi.MODEL_PINHOLE:
    prob, _ = synthetic.config_c2(scale=float(os.environ.get("THB_K1_AB_SCALE", "1.0")))
else:
    prob, _ = synthetic.make_ba_problem(1000, 100000, 10, models=(MODEL,), seed=2, num_rings=10, ring_radius=24.0, box=(10.0, 10.0, 3.0))
dev = {k: (None if v is None else torch.from_numpy(v).cuda()) for k, v in prob.a.items()}
pd = prob.struct()
pd.memory_space = capi.THB_MEM_DEVICE
for k, v in dev.items():
    setattr(pd, k, None if v is None else v.data_ptr())
sptr = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for spec in (sys.argv[1:] or ["gather", "shared", "gather", "shared"]):
    mode = spec
    os.environ["THB_K1_MODE"] = mode
    sess = C.c_void_p()
    o = capi.default_options(lib)
    o.use_inner_iterations = 0
    o.loss_function_type = LOSS
    o.use_homogeneous_point_parametrization = 0 if os.environ.get("THB_K1_AB_EUCLID") else 1
    capi.check(lib.thb_ba_create(C.byref(pd), C.byref(o), sptr, C.byref(sess)))
    ms = C.c_double(0.0)
    for rep in range(2):
        capi.check(lib.thb_ba_time_jacobian(sess, 20, 1, C.byref(ms)))
    capi.check(lib.thb_ba_finish(sess, None))
    ab = prob.num_observations * 200 + prob.num_cameras * 104 + prob.num_points * 32
    print("K1 model %d loss %d obs %d %-8s %.2f us  %.0f GB/s (%.1f MB algorithmic)" % (MODEL, LOSS, prob.num_observations, spec, ms.value * 1e3, ab / (ms.value * 1e-3) / 1e9, ab / 1e6), flush=True)
