"""pytest configuration: registers the ``gpu`` marker and puts the repo root on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")


def pytest_sessionstart(session):
    """The native artefacts are git-ignored: on a fresh checkout build them once (nvcc cross-compiles sm_100a without a
    GPU).  The product itself never builds or falls back -- tf_seq2seq_losses_b200._lib raises when the library is
    missing -- this is test infrastructure, the same thing ``__graft_entry__.build()`` does."""
    import subprocess
    lib = os.path.join(ROOT, "tf_seq2seq_losses_b200", "libctc_b200.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tf_seq2seq_losses_b200", "csrc"), "-j", str(os.cpu_count() or 4)],
                       check=True, stdout=subprocess.DEVNULL)
    if not os.path.exists(os.path.join(ROOT, "oracle", "libctc_oracle.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible, so a plain ``pytest tests`` stays green."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
