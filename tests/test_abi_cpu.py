"""CPU checks of the drop-in boundary: the C-ABI library builds/loads and exports exactly what include/tmp_b200.h
declares (no compute calls without a GPU), and the ctypes table matches the header."""
import ctypes
import os
import re

from medical_tri_modal_pilot_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    with open(os.path.join(ROOT, "include", "tmp_b200.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    decls = re.findall(r"\b(?:int|const char\*)\s+(tmp_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S)
    return {name: [a.strip() for a in args.split(",")] if args.strip() != "void" else [] for name, args in decls}


def test_library_builds_and_exports_every_declared_symbol():
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    funcs = header_functions()
    assert len(funcs) >= 19
    for name in funcs:
        assert hasattr(lib, name), f"{name} declared in include/tmp_b200.h but not exported"
    lib.tmp_abi_version.restype = ctypes.c_int
    assert lib.tmp_abi_version() == 5


def test_ctypes_table_matches_header():
    funcs = header_functions()
    assert set(funcs) == set(_lib.SIGNATURES), set(funcs) ^ set(_lib.SIGNATURES)
    for name, args in funcs.items():
        assert len(args) == len(_lib.SIGNATURES[name]), (name, len(args), len(_lib.SIGNATURES[name]))


def test_argument_errors_are_reported_not_crashed():
    """Argument validation runs before any CUDA call, so it can be exercised on the CPU box."""
    lib = _lib.load()
    rc = lib.tmp_gemm_bias_act_fwd(None, 0, 0, None, 0, 0, 1, 1, 1, 1.0, None, 0, None, 0, 0, None, 0, 0, 0.0, 0, 0, None,
                                   None, 0, None, 0, None, None, 0, None)
    assert rc < 0 and "null operand" in _lib.last_error()
    rc = lib.tmp_mma_attn_fwd(1, None, 1, 1, 3, 1, 8, 1, 128, 1, None)
    assert rc < 0 and "H==4" in _lib.last_error()
    rc = lib.tmp_adamw_step(1, 1, 1, 1, 6, 0.1, 0.9, 0.999, 1e-8, 0.0, 1, None)
    assert rc < 0
    rc = lib.tmp_adamw_step_dev(1, 1, 1, 1, 8, None, 0.9, 0.999, 1e-8, 0.0, None, 1, None)
    assert rc < 0


def test_no_cpu_fallback():
    import pytest
    import torch
    from medical_tri_modal_pilot_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.umse_embed(torch.zeros(4, 3), [torch.zeros(256)] * 4, [torch.zeros(256)] * 4, torch.zeros(20, 256))


def test_reference_copy_recipe():
    """oracle/_ref (the reference's own hot-path files, copied unmodified by oracle/build_ref.py; git-ignored) is present
    after __graft_entry__.build() and its files match the recorded sha256."""
    import pytest
    from oracle import build_ref
    if not build_ref.available():
        build_ref.build()
    if not build_ref.available():
        pytest.skip("reference not mounted and no oracle/_ref copy")
    assert build_ref.verify()
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run(["git", "-C", root, "check-ignore", "oracle/_ref/MANIFEST.json"], capture_output=True, text=True)
    assert out.returncode == 0, "oracle/_ref must stay out of the history"
