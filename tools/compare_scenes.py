"""Compare two scene-runner builds on a list of scenes (developer tool; not part of the product)."""
import argparse, sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelforge_b200 import load_product_scenes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from checkers import load_oracle_scenes, load_reference_scenes

CASES = [
    ("gears", 800, 600, {}),
    ("gears", 800, 600, {"first_frame": 7}),
    ("textured", 640, 360, {"size": 64, "variant": 0}),
    ("textured", 640, 360, {"size": 64, "variant": 2 | 8}),
    ("textured", 640, 360, {"size": 64, "variant": 4 | 32}),
    ("phong", 640, 360, {"size": 96}),
    ("overdraw", 512, 256, {"size": 8}),
    ("overdraw", 512, 256, {"size": 8, "variant": 1}),
    ("batch", 256, 256, {"size": 3}),
]

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--a", default="product"); ap.add_argument("--b", default="ref")
    ap.add_argument("--bilinear", action="store_true")
    args = ap.parse_args()
    def load(k):
        return {"product": load_product_scenes, "oracle": load_oracle_scenes,
                "ref": lambda: load_reference_scenes(False), "ref_bfix": lambda: load_reference_scenes(True)}[k]()
    A, B = load(args.a), load(args.b)
    print("A:", A.backend, " B:", B.backend)
    bad = 0
    cases = list(CASES)
    if args.bilinear:
        cases = [("textured", 640, 360, {"size": 64, "variant": 1}), ("textured", 640, 360, {"size": 64, "variant": 1 | 4 | 8}),
                 ("overdraw", 512, 256, {"size": 4, "variant": 3})]
    for name, w, h, kw in cases:
        ca, da, ra = A.render(name, w, h, **kw)
        cb, db, rb = B.render(name, w, h, **kw)
        nc = int((ca != cb).sum()); nd = int((da.view(np.uint32) != db.view(np.uint32)).sum())
        mx = int(np.abs(ca.view(np.uint8).astype(int) - cb.view(np.uint8).astype(int)).max())
        print(f"{name:9s} {w}x{h} {kw}: colour mismatches {nc} (max channel diff {mx}), depth mismatches {nd}, "
              f"shadedA {ra.pixels_shaded} trisA {ra.triangles_submitted}/{ra.triangles_rasterised} msA {ra.ms_median:.2f} msB {rb.ms_median:.2f}")
        bad += (nc > 0) + (nd > 0)
    print("MISMATCHING CASES:", bad)
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
