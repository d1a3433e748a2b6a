"""Random walks over the API (scene "fuzz"): the product on cuda:0 against the C oracle, colour and depth bit for bit.
usage: python tools/fuzz_gpu.py [first_seed] [last_seed] [ops]      (the CPU twin, oracle against the live reference, is
tests/test_oracle_parity.py::test_fuzz_oracle_matches_live_reference).
State at the end of round 2: written after the GPU budget was spent; the one 10-second attempt ended in pfscene_open returning
NULL before any frame was drawn (message cut off by the log tail) - to be run and looked at first thing next round."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from pixelforge_b200 import load_product_scenes
from checkers import load_oracle_scenes

lo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
hi = int(sys.argv[2]) if len(sys.argv) > 2 else 24
ops = int(sys.argv[3]) if len(sys.argv) > 3 else 150
prod, orc = load_product_scenes(), load_oracle_scenes()
bad = []
for seed in range(lo, hi):
    kw = dict(variant=0, seed=seed, size=ops)
    try:
        cp, dp, rp = prod.render("fuzz", 256, 192, **kw)
    except RuntimeError as e:       # say why: the library keeps the last error text
        import ctypes
        lib = ctypes.CDLL(os.path.join(ROOT, "pixelforge_b200", "lib", "libpixelforge.so"))
        lib.pfcu_last_error.restype = ctypes.c_char_p
        print("product failed:", e, "| pfcu_last_error:", lib.pfcu_last_error().decode(), flush=True)
        raise
    co, do, ro = orc.render("fuzz", 256, 192, **kw)
    dc, dd = int((cp != co).sum()), int((dp.view(np.uint32) != do.view(np.uint32)).sum())
    if dc or dd or rp.pixels_shaded != ro.pixels_shaded:
        bad.append((seed, dc, dd, rp.pixels_shaded, ro.pixels_shaded))
print("fuzz seeds", lo, hi, "ops", ops, "mismatches:", bad)
sys.exit(1 if bad else 0)
