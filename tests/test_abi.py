"""CPU: the C-ABI shared library loads without a GPU and exports exactly what include/b200vqa.h declares."""
import os
import re
import subprocess

import pytest

import common  # noqa: F401
from explainable_spatial_vqa_b200 import _native as nat

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "b200vqa.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"B200VQA_API[^;(]*?\b(b200vqa_\w+)\s*\(", text)))


def test_header_declares_the_hot_path_entry_points():
    syms = declared_symbols()
    for must in ("b200vqa_create", "b200vqa_destroy", "b200vqa_iqap_forward", "b200vqa_iqap_decode",
                 "b200vqa_iqap_forward_host", "b200vqa_fa_project_images", "b200vqa_fa_step", "b200vqa_fa_forward",
                 "b200vqa_fa_run_chain", "b200vqa_last_error", "b200vqa_workspace_bytes"):
        assert must in syms


def test_library_exports_every_declared_symbol_and_nothing_else():
    assert os.path.exists(nat.LIB_PATH), "run __graft_entry__.build() first"
    out = subprocess.run(["nm", "-D", "--defined-only", nat.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted({line.split()[-1] for line in out.splitlines() if " T " in line})
    assert exported == declared_symbols()


def test_ctypes_signatures_cover_the_header():
    assert sorted(nat.SIGNATURES) == declared_symbols()
    lib = nat.lib()
    for name in nat.SIGNATURES:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.b200vqa_version()
    assert lib.b200vqa_profile_num_tags() > 0


def test_library_contains_only_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", nat.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_null_and_bad_arguments_are_reported_not_crashed():
    lib = nat.lib()
    import ctypes as C
    assert lib.b200vqa_create(None, 0, None) < 0
    out = C.c_void_p(0)
    assert lib.b200vqa_create(None, 0, C.byref(out)) == -1
    assert b"NULL" in lib.b200vqa_last_error()
    d = nat.ModelDesc()
    d.kind = 0
    d.d_model = 512
    assert lib.b200vqa_create(C.byref(d), 0, C.byref(out)) == -5  # unsupported shape, checked before any CUDA call
    assert lib.b200vqa_workspace_bytes(None, 4) == 0
    assert lib.b200vqa_launch_count(None) == 0
    assert lib.b200vqa_iqap_forward(None, None, None, 1, 27, None, None, None, None, None, None) == -1


def test_no_cpu_fallback_in_the_product_path():
    import torch
    m = common.seeded_iqap()
    with pytest.raises(nat.NativeError):
        m(torch.zeros(1, 196, 1024), torch.zeros(1, 46, dtype=torch.long))
    f = common.seeded_fa()
    with pytest.raises(nat.NativeError):
        f(torch.zeros(1, 1024, 14, 14), torch.zeros(1, 3, dtype=torch.long), torch.zeros(1, 4, dtype=torch.long))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or execute it."""
    pkg = os.path.join(REPO, "explainable-spatial-vqa_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")) or f == "Makefile":
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "executor_oracle" not in text and "oracle/" not in text, f


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", "/nonexistent/libb200vqa.so")
    with pytest.raises(nat.NativeError, match="no CPU"):
        nat.lib()
