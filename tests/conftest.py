import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): build what is missing, the way
    __graft_entry__.build() does -- the CUDA library (nvcc cross-compiles without a GPU), the C restatement
    and, where /root/reference exists, the reference compiled from where it lies."""
    import subprocess
    need = [os.path.join(ROOT, "rtl_fm_player_b200", "libfmb.so"), os.path.join(ROOT, "rtl_fm_player_b200", "libfmsynth.so"),
            os.path.join(ROOT, "oracle", "libfm_oracle.so")]
    if os.path.exists("/root/reference/src/rtl_fm_player.c"):
        need.append(os.path.join(ROOT, "oracle", "_ref", "libfmref.so"))
    if all(os.path.exists(f) for f in need):
        return
    for cmd in (["make", "-C", os.path.join(ROOT, "rtl_fm_player_b200", "csrc"), "-j4"],
                ["make", "-C", os.path.join(ROOT, "oracle"), "port", "ref"]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            pytest.exit("building the test artefacts failed: " + " ".join(cmd) + "\n" + r.stdout[-3000:] + r.stderr[-3000:],
                        returncode=3)


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
