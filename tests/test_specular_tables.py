"""Host logic behind Gouraud lighting on the device: the specular tables that replace powf (include/pfcu.h,
PFCU_POW_TABLE_SIZE; pf_pipeline.c pow_table_index).  powf is not correctly rounded, so the device cannot recompute
the host libm's value; the front end tabulates the step function x -> (PFubyte)(255 * powf(x, s)) by bisection over
float bit patterns.  The table must reproduce this libm's powf exactly: on random inputs and right at every step."""
import ctypes as C
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def front_end(oracle_scenes):
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libpixelforge_oracle.so"))
    lib.pfxSpecularTableCheck.restype = C.c_int
    lib.pfxSpecularTableCheck.argtypes = [C.c_float, C.c_uint]
    return lib


@pytest.mark.parametrize("shininess", [1.0, 1.5, 5.0, 16.0, 32.0, 64.0, 100.0, 128.0, 777.25, 1024.0])
def test_table_reproduces_libm_powf(front_end, oracle_scenes, shininess):
    with oracle_scenes.open("gears", 64, 48) as sc:
        sc.make_current(0)
        assert front_end.pfxSpecularTableCheck(shininess, 200000) == 0


@pytest.mark.parametrize("shininess", [0.0, 0.5, 0.999, 1025.0, float("inf"), float("nan"), -3.0])
def test_untabulated_shininess_falls_back_to_host_lighting(front_end, oracle_scenes, shininess):
    with oracle_scenes.open("gears", 64, 48) as sc:
        sc.make_current(0)
        assert front_end.pfxSpecularTableCheck(shininess, 10) == -1
