"""pt.sfm.BundleAdjustReconstruction on the 1M-observation C2 scene through the pybind adapter (bench.py::adapter_e2e alone). GPU box only.
THB_ADAPTER_PROF=1 prints the adapter's own gather / solve / scatter stamps."""
import json, sys
sys.path.insert(0, ".")
import bench
from pytheiasfm_b200 import synthetic
prob, _ = synthetic.config_c2()
print(json.dumps(bench.adapter_e2e(prob, 3)))
