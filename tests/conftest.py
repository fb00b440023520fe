import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # tests/test_ram_shard_gpu.py runs several "ranks" of the multi-GPU step on ONE device: a rank's barrier kernel spins
    # while the next rank's kernels start, so no kernel may need (device-synchronising) lazy module loading by then
    os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
    # ... and no two active streams may share a hardware queue (a spinning barrier at the head of a queue would block the
    # kernels it waits for): 8 in-process ranks x 2 streams need more than the default 8 connections.  A failed in-process
    # barrier should cost seconds, not the production 20 s.
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    os.environ.setdefault("RSG_BARRIER_TIMEOUT_MS", "4000")
    if os.environ.get("RSG_EMU") == "1":
        # development aid for a GPU-less container: run the RAM `-m gpu` tests against the kernels
        # compiled for the host-CPU CUDA emulator (tests/emu/).  Test infrastructure only.
        use_emulator()


def use_emulator():
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    from ramscb_b200 import host
    host.LIB_PATH = build_emu.build()
    host._lib = None
    os.environ["RSG_SCB_NO_CLUSTER"] = "1"     # thread-block clusters are not emulated (tests/emu/cooperative_groups.h)


@pytest.fixture(scope="session")
def oracle_built():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def default_grids():
    from ramscb_b200 import grids
    return grids.build_grids()
