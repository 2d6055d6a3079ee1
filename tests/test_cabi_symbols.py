"""The C-ABI library loads without a GPU and exports every symbol include/b200_lora.h declares, with the
argument counts the ctypes binding (ai_toolkit_b200/cabi.py) assumes.  No compute calls here."""
import ctypes
import os
import re

import pytest

from ai_toolkit_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_prototypes():
    text = open(os.path.join(ROOT, "include", "b200_lora.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef struct b200_gemm_desc \{.*?\} b200_gemm_desc;", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\*)\s+(b200_\w+)\s*\(([^;{]*)\)\s*;", text):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("void", "") else len([a for a in args.split(",") if a.strip()])
        protos[name] = n
    return protos


def test_library_exists_and_loads():
    lib = cabi.load_library()
    assert lib.b200_version() >= 100


def test_every_declared_symbol_is_exported_and_bound():
    protos = _header_prototypes()
    assert len(protos) >= 20
    lib = ctypes.CDLL(cabi.LIB_PATH)
    for name, nargs in protos.items():
        assert hasattr(lib, name), f"{name} declared in include/b200_lora.h but not exported"
        assert name in cabi.SIGNATURES, f"{name} has no ctypes signature in cabi.py"
        assert len(cabi.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} args, cabi.py binds {len(cabi.SIGNATURES[name][1])}"
    for name in cabi.SIGNATURES:
        assert name in protos, f"{name} bound in cabi.py but not declared in the header"


def test_gemm_desc_layout_matches_header():
    text = open(os.path.join(ROOT, "include", "b200_lora.h")).read()
    body = re.search(r"typedef struct b200_gemm_desc \{(.*?)\} b200_gemm_desc;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.findall(r"(\w+)\s*$", part.strip())[0])
    assert names == [f[0] for f in cabi.GemmDesc._fields_]


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("device present")
    lib = cabi.load_library()
    h = ctypes.c_void_p()
    rc = lib.b200_ctx_create(ctypes.byref(h), 0)
    assert rc != 0 and h.value is None
    assert "no" in cabi.last_error().lower()
