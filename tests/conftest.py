import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu() -> bool:
    try:
        from hast_b200 import capi
        return capi.load_library().hast_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build what is missing (cheap no-op when the tree is already built)."""
    targets = {"lib": ROOT / "hast_b200/lib/libhast_b200.so", "tools": ROOT / "hast_b200/lib/libhast_tools.so",
               "host": ROOT / "bin/classify"}
    for tgt, path in targets.items():
        if not path.exists():
            subprocess.run(["make", "-C", str(ROOT), tgt], check=False, capture_output=True)
    if not (ROOT / "oracle/liboracle.so").exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "port"], check=False, capture_output=True)
    yield


@pytest.fixture(scope="session", params=[3, 1, 0, 2, 4], ids=["prefilter_mini", "prefilter", "direct", "prefilter_tma", "prefilter_mini_tma"])
def engine(request):
    """One context per fused-kernel variant: 3 = Bloom pre-filter addressed by the k-mer's minimizer + exact table (default;
    k = 21/25/31; any other k, 17 included, runs as 1), 1 = pre-filter addressed by a hash of the k-mer, 2 = the same with TMA-staged read
    bytes, 4 = 3 with TMA-staged read bytes, 0 = direct table probe per position.  All must be bit-identical to the oracle."""
    from hast_b200.capi import Engine
    e = Engine(0)
    e.set_option("kernel", request.param)
    yield e
    e.close()
