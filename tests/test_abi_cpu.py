"""CPU-only checks of the C-ABI shared library and the host layer: the library loads without a GPU, exports every
symbol include/ctc_b200.h declares, validates descriptors, and the Python face refuses to run without CUDA (there is
no CPU fallback).  No kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

from tf_seq2seq_losses_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ctc_b200.h")).read()
    return sorted(set(re.findall(r"\b(ctcb200_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"libctc_b200.so does not export {name}"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert lib.ctcb200_version() == 100


def test_error_strings_and_descriptor_validation():
    lib = _lib.load()
    assert lib.ctcb200_strerror(0) == b"ok"
    for code in range(-6, 0):
        assert lib.ctcb200_strerror(code) not in (b"ok", b"unknown error")
    good = _lib.Desc(4, 10, 64, 3, 0, _lib.CLASSIC, 4, 0)
    assert lib.ctcb200_workspace_bytes(ctypes.byref(good), _lib.WS_LOSS_GRAD) > 0
    assert lib.ctcb200_workspace_bytes(ctypes.byref(good), _lib.WS_HESSIAN) > lib.ctcb200_workspace_bytes(
        ctypes.byref(good), _lib.WS_LOSS_GRAD)
    for bad in (_lib.Desc(-1, 10, 8, 3, 0, 0, 4, 0),      # negative batch
                _lib.Desc(4, 10, 8, 3, 8, 0, 4, 0),       # blank out of range (V = 8)
                _lib.Desc(4, 10, 8, 3, 0, 2, 4, 0),       # unknown variant
                _lib.Desc(4, 10, 8, 3, 0, 0, 4, 1 << 20),  # unknown flag
                _lib.Desc(4, 1200, 8, 1100, 0, 0, 1101, 0),  # more than 1024 label states
                _lib.Desc(4, 10, 40000, 3, 0, 0, 4, 0)):   # more than 32768 tokens
        assert lib.ctcb200_workspace_bytes(ctypes.byref(bad), _lib.WS_LOSS_GRAD) == 0
    # null pointers and a too-small workspace are reported, not dereferenced
    vp = ctypes.c_void_p
    assert lib.ctcb200_loss_grad(ctypes.byref(good), None, None, None, None, None, None, None, None, None, 0, None) == -1
    buf = (ctypes.c_char * 512)()
    p = ctypes.cast(buf, vp)
    aligned = vp((p.value + 255) & ~255)
    assert lib.ctcb200_loss_grad(ctypes.byref(good), aligned, aligned, aligned, aligned, None, aligned, aligned, None,
                                 vp(aligned.value + 4), 1 << 30, None) == -6    # misaligned workspace
    assert lib.ctcb200_loss_grad(ctypes.byref(good), aligned, aligned, aligned, aligned, None, aligned, aligned, None,
                                 aligned, 16, None) == -3                        # workspace too small
    # stage names: one fused launch by default, the three staged kernels on request
    assert lib.ctcb200_stage_names(ctypes.byref(good)) == b"kf_fused"
    staged = _lib.Desc(4, 10, 29, 3, 0, _lib.CLASSIC, 4, _lib.FORCE_STAGED)
    assert lib.ctcb200_stage_names(ctypes.byref(staged)) == b"k1_softmax_gather,k2_recursion,k3_grad"
    assert lib.ctcb200_launches_per_call(ctypes.byref(good)) == 1 and lib.ctcb200_launches_per_call(ctypes.byref(staged)) == 3
    narrow = _lib.Desc(4, 10, 29, 3, 0, _lib.CLASSIC, 4, 0)        # character-sized vocabulary: staged unless forced
    assert lib.ctcb200_launches_per_call(ctypes.byref(narrow)) == 3
    narrow.flags = _lib.FORCE_FUSED
    assert lib.ctcb200_stage_names(ctypes.byref(narrow)) == b"kf_fused"


def test_workspace_classes_and_kernel_choice():
    """Workspace sizing rules of include/ctc_b200.h (no kernel is launched): the training call's class is never larger
    than the general one and shrinks when the fused kernel serves the shape; ctcb200_stage_names reports the device path
    (fused from V >= 64, or from 48 utterances on for narrower vocabularies, unless forced)."""
    lib = _lib.load()
    ws = lambda d, w: lib.ctcb200_workspace_bytes(ctypes.byref(d), w)
    names = lambda d: lib.ctcb200_stage_names(ctypes.byref(d)).decode()
    north_star = _lib.Desc(256, 1000, 1024, 200, 0, _lib.SIMPLIFIED, 201, 0)
    assert names(north_star) == "kf_fused" and lib.ctcb200_launches_per_call(ctypes.byref(north_star)) == 1
    assert ws(north_star, _lib.WS_LOSS_GRAD_LOGITS) * 2 < ws(north_star, _lib.WS_LOSS_GRAD)
    assert ws(north_star, _lib.WS_HVP_LOGITS) > ws(north_star, _lib.WS_HESSIAN) > ws(north_star, _lib.WS_STATES)
    assert ws(north_star, _lib.WS_DECODE) == 2 * 256 * 1000 * 4          # arg-max token + its logit per row
    assert ws(north_star, 6) == 0 and ws(north_star, -1) == 0
    staged = _lib.Desc(256, 1000, 1024, 200, 0, _lib.SIMPLIFIED, 201, _lib.FORCE_STAGED)
    assert names(staged) == "k1_softmax_gather,k2_recursion,k3_grad"
    assert ws(staged, _lib.WS_LOSS_GRAD_LOGITS) == ws(staged, _lib.WS_LOSS_GRAD)
    small_char = _lib.Desc(32, 500, 29, 100, 0, _lib.CLASSIC, 101, 0)        # BASELINE configs[1]
    big_char = _lib.Desc(256, 255, 32, 255, 0, _lib.CLASSIC, 127, 0)         # the reference's tests/benchmark.py shape
    assert names(small_char).startswith("k1_") and names(big_char) == "kf_fused"
    assert names(_lib.Desc(32, 500, 29, 100, 0, _lib.CLASSIC, 101, _lib.FORCE_FUSED)) == "kf_fused"
    # beyond the fused kernel's 512 label states the staged kernels serve the call (up to 1024), bf16 rows do not exist
    wide = _lib.Desc(8, 1100, 1024, 1023, 0, _lib.CLASSIC, 1024, 0)
    assert names(wide) == "k1_softmax_gather,k2_recursion,k3_grad" and ws(wide, _lib.WS_LOSS_GRAD_LOGITS) > 0
    assert names(_lib.Desc(8, 1100, 1024, 1023, 0, _lib.CLASSIC, 1024, _lib.FORCE_FUSED)).startswith("k1_")
    assert names(_lib.Desc(8, 600, 1024, 512, 0, _lib.CLASSIC, 512, 0)) == "kf_fused"
    time_major = _lib.Desc(256, 1000, 1024, 200, 0, _lib.SIMPLIFIED, 201, _lib.TIME_MAJOR)
    assert ws(time_major, _lib.WS_LOSS_GRAD_LOGITS) == ws(north_star, _lib.WS_LOSS_GRAD_LOGITS)      # scratch keeps its layout
    assert ws(_lib.Desc(256, 1000, 1024, 200, 0, _lib.SIMPLIFIED, 201, 64), _lib.WS_LOSS_GRAD) == 0     # unknown flag bit
    assert ws(_lib.Desc(256, 1000, 1024, 200, 0, _lib.SIMPLIFIED, 201, _lib.GRAD_BF16), _lib.WS_LOSS_GRAD) == 0   # needs LOGITS_BF16
    bf16 = _lib.Desc(256, 1000, 1024, 200, 0, _lib.SIMPLIFIED, 201, _lib.LOGITS_BF16 | _lib.GRAD_BF16)
    assert ws(bf16, _lib.WS_LOSS_GRAD_LOGITS) == ws(north_star, _lib.WS_LOSS_GRAD_LOGITS) and names(bf16) == "kf_fused"
    # bf16 rows exist in the fused kernel only: a narrow vocabulary in a small batch takes it as well
    assert names(_lib.Desc(32, 500, 32, 100, 0, _lib.CLASSIC, 101, _lib.LOGITS_BF16)) == "kf_fused"
    # the small-batch (split) plan does not change the workspace class sizes' ordering
    slice32 = _lib.Desc(32, 1000, 1024, 200, 0, _lib.SIMPLIFIED, 201, 0)
    assert names(slice32) == "kf_fused" and ws(slice32, _lib.WS_LOSS_GRAD_LOGITS) * 2 < ws(slice32, _lib.WS_LOSS_GRAD)
    full_sweep = _lib.Desc(2048, 1600, 5000, 400, 0, _lib.CLASSIC, 401, 0)   # BASELINE configs[4] on one GPU
    assert names(full_sweep) == "kf_fused"
    assert ws(full_sweep, _lib.WS_LOSS_GRAD_LOGITS) + 2 * 2048 * 1600 * 5000 * 4 < 180e9      # fits one B200


def test_header_is_plain_c_and_a_c_host_links(tmp_path):
    """include/ctc_b200.h is the whole contract: it compiles as C99 with -pedantic (no CUDA, torch or C++ in the
    signatures) and a C program linked against libctc_b200.so calls the sizing / diagnostic entry points."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = os.path.join(ROOT, "tests", "abi_smoke.c")
    libdir = os.path.join(ROOT, "tf_seq2seq_losses_b200")
    _lib.load()                                                # raises when the library has not been built
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                    "-I", os.path.join(ROOT, "include"), src], check=True)
    exe = str(tmp_path / "abi_smoke")
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", libdir,
                    "-l:libctc_b200.so", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == 232682496 and int(out[1]) > int(out[0])     # 0.23 GB of scratch for the north-star call


def test_python_face_has_no_cpu_fallback():
    import tf_seq2seq_losses_b200 as pkg
    logits = torch.zeros((1, 4, 3))
    with pytest.raises(_lib.CtcB200Error):
        pkg.classic_ctc_loss(torch.ones((1, 2), dtype=torch.int32), logits, torch.tensor([2]), torch.tensor([4]), 0)
    with pytest.raises(_lib.CtcB200Error):
        pkg.simple_ctc_loss(torch.ones((1, 2), dtype=torch.int32), logits, torch.tensor([2]), torch.tensor([4]), 0)
    with pytest.raises(AssertionError):       # rank / dtype checks of tf_seq2seq_losses/base_loss.py:129-138
        pkg.classic_ctc_loss(torch.ones((1, 2), dtype=torch.int32), logits.double(), torch.tensor([2]), torch.tensor([4]), 0)
    assert pkg.simple_ctc_loss is pkg.simplified_ctc_loss


def test_product_package_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "tf_seq2seq_losses_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, name)).read()
                assert "import oracle" not in text and "from oracle" not in text and "ctc_oracle" not in text, name
