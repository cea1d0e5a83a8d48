import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # Build what is missing (seconds): the product library and the C oracle. The reference library in oracle/_ref is
    # only (re)built where /root/reference exists; elsewhere the prebuilt copy that travelled with the snapshot is used.
    if not os.path.exists(os.path.join(ROOT, "rendu_b200", "librendu_b200.so")):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "rendu_b200", "csrc")])
    if not os.path.exists(os.path.join(ROOT, "oracle", "librendu_oracle.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    ref_libs = [os.path.join(ROOT, "oracle", "_ref", n) for n in ("librendu_ref.so", "librendu_refvis.so")]
    if not all(os.path.exists(p) for p in ref_libs) and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])


def has_gpu():
    try:
        import rendu_b200
        return rendu_b200.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
