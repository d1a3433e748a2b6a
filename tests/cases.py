"""Parity cases shared by the CPU (oracle) and GPU (product) test modules and by tools/gen_golden.py.

Each case is (id, scene, width, height, kwargs, needs_bilinear_fix).  Sizes are chosen so that the
scalar oracle finishes each case in well under a second; BASELINE.json's full sizes are covered by
property tests in test_gpu_parity.py.
"""


def micro_variant(blend=None, depth=None, flat=0, tex=0, bil=0, wrap=0, cull_off=0, mode=0, rgb=0, persp=0,
                  phong=0, fbo=0, spot=0, bgra=0):
    v = 0
    if blend is not None:
        v |= 8 | blend
    if depth is not None:
        v |= 128 | (depth << 4)
    v |= (flat << 8 | tex << 9 | bil << 10 | wrap << 11 | cull_off << 13 | mode << 14 | rgb << 17 | persp << 18
          | phong << 19 | fbo << 20 | spot << 21 | bgra << 22)
    return v


def _micro(desc, seed, size=40, ref_bfix=False, **kw):
    return (f"micro-{desc}", "micro", 160, 120, dict(variant=micro_variant(**kw), seed=seed, size=size), ref_bfix)


CASES = [
    # BASELINE.json configs at reduced size (same state, same code path)
    ("c1-gears-f0", "gears", 800, 600, dict(), False),
    ("c1-gears-f7", "gears", 800, 600, dict(first_frame=7), False),
    ("c2-textured-nearest-repeat", "textured", 640, 360, dict(size=64, variant=0), False),
    ("c2-textured-nearest-mirror-rgb8", "textured", 640, 360, dict(size=64, variant=2 | 8), False),
    ("c2-textured-nearest-clamp-arrays", "textured", 640, 360, dict(size=64, variant=4 | 32), False),
    ("c2-textured-bilinear-repeat", "textured", 640, 360, dict(size=64, variant=1), True),
    ("c2-textured-bilinear-clamp-rgb8", "textured", 640, 360, dict(size=64, variant=1 | 4 | 8), True),
    # close-up camera inside the torus: heavy near-plane / frustum clipping; vertex arrays -> device vertex stage
    ("c2-textured-closeup-clipped-arrays", "textured", 640, 360, dict(size=96, variant=32 | 64), False),
    ("c2-textured-closeup-clipped-bilinear-arrays", "textured", 640, 360, dict(size=96, variant=1 | 32 | 64), True),
    ("c2-textured-closeup-clipped-immediate", "textured", 640, 360, dict(size=96, variant=64), False),
    ("c3-phong", "phong", 640, 360, dict(size=96), False),
    ("c3-phong-arrays", "phong", 320, 200, dict(size=48, variant=32), False),
    ("c4-overdraw-add", "overdraw", 512, 256, dict(size=8), False),
    ("c4-overdraw-alpha-depth", "overdraw", 512, 256, dict(size=8, variant=1), False),
    ("c4-overdraw-alpha-depth-bilinear", "overdraw", 256, 128, dict(size=4, variant=3), True),
    ("c5-batch", "batch", 256, 256, dict(size=3), False),
]
# every blend mode / depth function / draw mode / wrap mode / shade mode of the hot path
CASES += [_micro(f"blend{b}", 1, blend=b, cull_off=1) for b in range(8)]
CASES += [_micro(f"depth{d}", 3, depth=d, cull_off=1) for d in range(6)]
CASES += [_micro(f"mode{m}-cull{c}", 5 + m, mode=m, cull_off=1 - c, depth=2) for m in range(6) for c in (0, 1)]
CASES += [_micro(f"flat{f}", 1, flat=f, cull_off=1, blend=1) for f in (0, 1)]
CASES += [_micro(f"tex-wrap{w}-rgb{r}", 11, tex=1, wrap=w, rgb=r, cull_off=1, blend=1) for w in range(3) for r in (0, 1)]
CASES += [_micro(f"tex-persp-wrap{w}-rgb{r}", 12, tex=1, wrap=w, rgb=r, cull_off=1, persp=1, depth=2) for w in range(3) for r in (0, 1)]
CASES += [_micro(f"bilinear-wrap{w}", 13, ref_bfix=True, tex=1, bil=1, wrap=w, cull_off=1) for w in range(3)]
CASES += [_micro(f"bilinear-persp-wrap{w}", 14, ref_bfix=True, tex=1, bil=1, wrap=w, cull_off=1, persp=1, depth=2) for w in range(3)]
CASES += [_micro(f"phong-spot{s}", 15, phong=1, persp=1, cull_off=1, depth=2, spot=s) for s in (0, 1)]
CASES += [_micro(f"phong-tex-spot{s}", 16, phong=1, persp=1, cull_off=1, depth=3, spot=s, tex=1, wrap=1) for s in (0, 1)]
CASES += [_micro(f"phong-2d-spot{s}", 17, phong=1, cull_off=1, spot=s, blend=2) for s in (0, 1)]
CASES += [_micro("fbo", 18, fbo=1, tex=1, cull_off=1, depth=2), _micro("fbo-persp", 19, fbo=1, tex=1, cull_off=1, depth=2, persp=1, blend=1)]
CASES += [_micro(f"random{s}", s, size=60, blend=s % 8, depth=s % 6, tex=s & 1, wrap=s % 3, cull_off=(s >> 1) & 1,
                 mode=s % 6, persp=(s >> 2) & 1, flat=(s >> 3) & 1) for s in range(20, 40)]



# breadth of the public API around the triangle path (scene "api", see scenes.c for the variant bits):
# viewport offsets, texture matrix, Gouraud with spot/attenuation/back materials/colour material/normalize,
# colour arrays, pfRect*, pfDrawPixels + zoom, fog, post-processing, pfReadPixels, pfClearDepth, aux buffer
def _api(desc, variant, seed=1, ref_bfix=False):
    return (f"api-{desc}", "api", 200, 150, dict(variant=variant, seed=seed), ref_bfix)


_B = lambda *bits: sum(1 << b for b in bits)
CASES += [_api("plain", 0), _api("viewport", _B(0)), _api("texmatrix", _B(1)), _api("gouraud", _B(2)),
          _api("gouraud-backmat", _B(2, 3)), _api("gouraud-backmat-colormat", _B(2, 3, 12)),
          _api("gouraud-normalize", _B(2, 11)), _api("arrays-colorptr", _B(13)),
          _api("gouraud-arrays-normalize-colormat", _B(2, 11, 12, 13)), _api("cullfront", _B(10)),
          _api("rects", _B(4)), _api("rects-viewport", _B(0, 4)), _api("drawpixels", _B(5)),
          _api("drawpixels-viewport-blend", _B(0, 5, 17)), _api("readpixels", _B(8)),
          _api("fog-linear", _B(6)), _api("fog-exp-opaque", _B(6, 15, 16)), _api("fog-cleardepth", _B(6, 9)),
          _api("postprocess", _B(7)), _api("swapbuffers", _B(14, 7)),
          _api("everything", 0x3ffff & ~_B(13), seed=2), _api("everything-bilinear", 0x7ffff & ~_B(13), seed=3, ref_bfix=True)]

CASE_IDS = [c[0] for c in CASES]
assert len(set(CASE_IDS)) == len(CASE_IDS)


# points, lines and PF_POINT / PF_LINE polygon modes (scene "prims"): the reference's scalar rasterisers
# (lines.c, points.c) - scalar blend / depth tables, thick lines, frustum-clipped 3D lines
def _prims(desc, blend=None, depth=None, persp=0, thick=0, seed=1, size=48):
    v = (persp << 8) | (thick << 9)
    if blend is not None:
        v |= 1 | (blend << 1)
    if depth is not None:
        v |= 16 | (depth << 5)
    return (f"prims-{desc}", "prims", 200, 150, dict(variant=v, seed=seed, size=size), False)


CASES += [_prims("plain"), _prims("thick", thick=1, seed=2), _prims("persp", persp=1, seed=3), _prims("persp-depth-less", persp=1, depth=2, seed=4)]
CASES += [_prims(f"blend{b}-thick", blend=b, thick=1, seed=5 + b) for b in range(8)]
CASES += [_prims(f"depth{d}-thick", depth=d, thick=1, seed=20 + d) for d in range(6)]
CASES += [_prims("persp-blend1-depth3-points", blend=1, depth=3, persp=1, thick=1, seed=30, size=90)]

CASE_IDS = [c[0] for c in CASES]
