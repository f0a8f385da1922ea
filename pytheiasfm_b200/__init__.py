"""theia-b200: B200-native bundle adjustment and RANSAC verification behind the pyTheia API.

Product path = libtheia_b200.so (hand-written sm_100a CUDA behind the C-ABI of include/theia_b200.h) and, above it,
the C++/pybind11 adapter `_pt` that mirrors the reference's `pt.sfm`, `pt.solvers` and `pt.matching` names:

    import pytheiasfm_b200 as pt
    pt.sfm.BundleAdjustReconstruction(opts, recon); pt.sfm.EstimateRelativePose(params, pt.sfm.RansacType.RANSAC, corrs)

No CPU fallback exists: see capi.load_library().
"""
from . import capi  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name in ("sfm", "solvers", "matching"):
        try:
            from . import _pt
        except ImportError as e:  # not built: fail loudly, there is nothing to fall back to
            raise capi.LibraryNotBuilt("the pybind11 adapter pytheiasfm_b200/_pt*.so is missing: run "
                                       "`python -c 'import __graft_entry__ as g; g.build()'` (%s)" % e)
        return getattr(_pt, name)
    raise AttributeError(name)
