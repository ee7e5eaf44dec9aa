import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libmctsb200_emu.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle_binding import Oracle, build

    build()
    return Oracle()


@pytest.fixture(scope="session")
def oracle_libm():
    from oracle.oracle_binding import Oracle, build

    build()
    return Oracle(libm=True)


@pytest.fixture(scope="session")
def emu_lib_path():
    """Host emulation build of the library sources (tests only; see tests/emu/cuda_shim.h)."""
    if os.environ.get("MCTSB200_EMU_LIB"):  # another build of the emulation (scripts/asan_emu.sh: the AddressSanitizer one)
        return os.environ["MCTSB200_EMU_LIB"]
    res = subprocess.run(["make", "-C", EMU_DIR], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    return EMU_LIB


@pytest.fixture()
def emu_ctx(emu_lib_path):
    import mcts_jl_b200 as m

    ctx = m.Context(0, lib_path=emu_lib_path)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def gpu_ctx():
    """The shipped CUDA library on cuda:0. Fails (not skips) if the extension is missing on a GPU box."""
    import mcts_jl_b200 as m

    ctx = m.Context(0)
    yield ctx
    ctx.close()
