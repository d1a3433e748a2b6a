"""The table formulas used by the CUDA rcp_x86()/rsqrt_x86() (host twin: pfh_rcp_from_table /
pfh_rsqrt_from_table in pf_x86approx.c) must reproduce this CPU's RCPPS/RSQRTPS on every input class."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib(built_libraries):
    from checkers import load_oracle_pfcu
    return load_oracle_pfcu()


def test_tables_reproduce_hardware(lib):
    rcp, rb, rsq, sb = lib.harvest_tables()
    assert 8 <= rb <= 23 and 8 <= sb <= 23
    rng = np.random.default_rng(0)
    bits = np.concatenate([
        rng.integers(0, 2**32, 100000, dtype=np.uint64).astype(np.uint32),
        np.array([0, 0x80000000, 0x7F800000, 0xFF800000, 0x7FC00000, 0x00000001, 0x007FFFFF, 0x00800000, 0x7F7FFFFF,
                  0x7E800000, 0x7EFFFFFF, 0x7F000000, 0x3F800000, 0xBF800000, 0x3F7FFFFF, 0x3F800001], dtype=np.uint32),
        (np.arange(1, 255, dtype=np.uint32) << 23) | 0x2AAAAA,
    ])
    xs = bits.view(np.float32)
    L = lib.lib
    bad = []
    for x in xs.tolist():
        for hw, tab, t, k in ((L.pfh_hw_rcp, L.pfh_rcp_from_table, rcp, rb), (L.pfh_hw_rsqrt, L.pfh_rsqrt_from_table, rsq, sb)):
            a, b = hw(x), tab(t, k, x)
            ua = np.array([a], dtype=np.float32).view(np.uint32)[0]; ub = np.array([b], dtype=np.float32).view(np.uint32)[0]
            if ua != ub and not (a != a and b != b):
                bad.append((x, hex(ua), hex(ub)))
    assert not bad, bad[:5]
