"""ABI of the drop-in boundary: include/pixelforge.h must agree with the reference header on every enum value,
public struct layout and function prototype (golden extracted from /root/reference/src/pixelforge.h by
tools/gen_abi_golden.py), and the built libraries must export every declared symbol."""
import json
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_abi_golden as G   # noqa: E402


@pytest.fixture(scope="module")
def golden_abi():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "abi_golden.json")))


def test_enum_values_and_struct_layouts(golden_abi):
    names = [n for n in golden_abi["values"] if not n.startswith(("sizeof_", "offsetof_"))]
    ours = G.probe(os.path.join(ROOT, "include"), names)
    assert ours == golden_abi["values"]


def test_prototypes_match_reference(golden_abi):
    ours = G.prototypes(open(os.path.join(ROOT, "include", "pixelforge.h")).read())
    assert len(golden_abi["prototypes"]) == 128
    assert set(ours) == set(golden_abi["prototypes"])
    for name, sig in golden_abi["prototypes"].items():
        assert ours[name] == sig, name


def _exported(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    return {l.split()[-1] for l in out.splitlines() if l.strip()}


@pytest.mark.parametrize("lib", ["pixelforge_b200/lib/libpixelforge.so", "oracle/_build/libpixelforge_oracle.so"])
def test_library_exports_every_declared_symbol(lib, golden_abi, built_libraries):
    path = os.path.join(ROOT, lib)
    if not os.path.exists(path):
        pytest.skip(f"{lib} not built (no nvcc here)")
    syms = _exported(path)
    from pixelforge_b200.binding import PFCU_SYMBOLS, PFX_SYMBOLS
    declared = set(golden_abi["prototypes"]) | {"pfRecti", "pfRectiv"} | set(PFCU_SYMBOLS) | set(PFX_SYMBOLS)
    for hdr in ("pfcu.h", "pfx.h"):           # every PFCU_API / PF_API declaration of our own headers too
        text = open(os.path.join(ROOT, "include", hdr)).read()
        declared |= set(re.findall(r"\b(pfcu_[a-z_0-9]+|pfx[A-Z]\w+)\s*\(", text))
    missing = sorted(declared - syms)
    assert not missing, f"{lib} does not export {missing}"


def test_product_library_loads_without_gpu(built_libraries):
    """The C-ABI library must load and fail LOUDLY (no CPU fallback) when no device is present."""
    path = os.path.join(ROOT, "pixelforge_b200/lib/libpixelforge.so")
    if not os.path.exists(path):
        pytest.skip("product not built (no nvcc here)")
    import ctypes
    lib = ctypes.CDLL(path)
    lib.pfcu_backend_name.restype = ctypes.c_char_p
    assert lib.pfcu_backend_name() == b"cuda-sm_100a"
    import torch
    if not torch.cuda.is_available():
        lib.pfcu_last_error.restype = ctypes.c_char_p
        assert lib.pfcu_init(-1) == 1                      # PFCU_ERR_NO_DEVICE
        lib.pfCreateContext.restype = ctypes.c_void_p
        buf = ctypes.create_string_buffer(64 * 64 * 4)
        assert lib.pfCreateContext(buf, 64, 64, 7, 0) is None    # PF_RGBA, PF_UNSIGNED_BYTE -> NULL, no fallback
