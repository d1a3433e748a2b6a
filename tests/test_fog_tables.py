"""Host logic behind pfFogProcess on the device: in the PF_EXP / PF_EXP2 modes the fog alpha is
(PFubyte)((1 - expf(-density * (depth - start))) * alpha) with the HOST libm's expf / exp2f (context.c:2331-2338), which is
not correctly rounded, so the device cannot recompute it; the front end tabulates the steps of that function by bisection
over float bit patterns (pfcu_fog.thresholds) and the device counts thresholds <= depth.  The table must reproduce this
libm exactly: on random depths of the fog range and right at every step."""
import ctypes as C
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PF_FOG_MODE, PF_FOG_START, PF_FOG_END, PF_FOG_COLOR = 0, 2, 3, 4


@pytest.fixture(scope="module")
def front_end(oracle_scenes):
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libpixelforge_oracle.so"))
    lib.pfxFogTableCheck.restype = C.c_int
    lib.pfxFogTableCheck.argtypes = [C.c_uint]
    lib.pfFogf.argtypes = [C.c_int, C.c_float]
    lib.pfFogiv.argtypes = [C.c_int, C.POINTER(C.c_int)]
    return lib


@pytest.mark.parametrize("mode", [1, 2], ids=["exp", "exp2"])
@pytest.mark.parametrize("start,end", [(0.0, 1.0), (1.2, 3.0), (0.93, 0.985), (-2.0, 5.0), (10.0, 1000.0), (1e-3, 2e-3)])
@pytest.mark.parametrize("alpha", [255, 200, 1])
def test_fog_steps_reproduce_libm(front_end, oracle_scenes, mode, start, end, alpha):
    with oracle_scenes.open("gears", 64, 48) as sc:
        sc.make_current(0)
        front_end.pfFogf(PF_FOG_START, start); front_end.pfFogf(PF_FOG_END, end)
        col = (C.c_int * 4)(10, 20, 30, alpha); front_end.pfFogiv(PF_FOG_COLOR, col)
        m = (C.c_int * 1)(mode); front_end.pfFogiv(PF_FOG_MODE, m)        # the unchecked setter reaches the exponential modes
        assert front_end.pfxFogTableCheck(100000) == 0


def test_linear_mode_needs_no_table(front_end, oracle_scenes):
    with oracle_scenes.open("gears", 64, 48) as sc:
        sc.make_current(0)
        m = (C.c_int * 1)(0); front_end.pfFogiv(PF_FOG_MODE, m)
        assert front_end.pfxFogTableCheck(10) == -2
