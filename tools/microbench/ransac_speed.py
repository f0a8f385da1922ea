import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo')
import torch
from pytheiasfm_b200 import capi, synthetic
lib = capi.load_library()
t=time.time(); batch, gts = synthetic.make_pair_batch(2048, n=2000, seed=21); print("gen", time.time()-t)
params = synthetic.c4_params(capi.ThbRansacParams())
d_off = torch.from_numpy(batch.pair_offset).cuda(); d_corr = torch.from_numpy(batch.corr).cuda()
d_seed = torch.from_numpy(batch.seed.astype(np.int64)).cuda().to(torch.int32)
d_res = torch.zeros(batch.num_pairs * capi.RELPOSE_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
d_mask = torch.zeros(int(batch.pair_offset[-1]), dtype=torch.uint8, device="cuda")
b = capi.ThbPairBatch(); b.num_pairs = batch.num_pairs; b.memory_space = capi.THB_MEM_DEVICE
b.pair_offset = d_off.data_ptr(); b.corr = d_corr.data_ptr(); b.seed = d_seed.data_ptr()
st = torch.cuda.current_stream()
for rep in range(3):
    torch.cuda.synchronize(); t = time.time()
    capi.check(lib.thb_ransac_relpose_batch(C.byref(b), C.byref(params), C.c_void_p(d_res.data_ptr()), C.c_void_p(d_mask.data_ptr()), C.c_void_p(st.cuda_stream)))
    torch.cuda.synchronize(); dt = time.time() - t
    print("pairs/s", batch.num_pairs / dt, "ms", dt * 1e3)
res = np.frombuffer(d_res.cpu().numpy().tobytes(), capi.RELPOSE_DTYPE)
print("iters mean", res["num_iterations"].mean(), "max", res["num_iterations"].max(), "inliers mean", res["num_inliers"].mean())
