"""ospray_b200 -- B200-native `cuda` device for OSPRay.

The product is three shared libraries under ospray_b200/lib (built by `make` / __graft_entry__.build()):
libospray.so (OSPRay C API), libospray_module_cuda.so (the device), libospcuda.so (sm_100a kernels).
This Python package is the host-side mirror used by tests and bench.py: a ctypes binding of the
OSPRay C API (capi), the ospray_testing scene builders restated on top of it (testing), the
ospTestSuite image metric (imgdiff) and the one-process-per-GPU tile gather (multigpu)."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.environ.get("OSPRAY_B200_LIB_DIR") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")


def lib_path(name):
    p = os.path.join(LIB_DIR, name)
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: build the native libraries first (make, or __graft_entry__.build()). "
            "There is no Python/CPU fallback for the cuda device.")
    return p
