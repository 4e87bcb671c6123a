import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The shared library is a build artefact (git-ignored): compile it once if a fresh checkout lacks it.
    (nvcc cross-compiles sm_100a without a GPU; on the GPU box the prebuilt .so travels with the snapshot.)"""
    so = os.path.join(ROOT, "neighbourlists.jl_b200", "libnlcuda.so")
    if not os.path.exists(so):
        subprocess.check_call(["bash", os.path.join(ROOT, "neighbourlists.jl_b200", "csrc", "build.sh")])
