"""Parity cases shared by the CPU (oracle) and GPU (product) test modules and by tools/gen_golden.py.

Each case is (id, scene, width, height, kwargs, needs_bilinear_fix).  Sizes are chosen so that the
scalar oracle finishes each case in well under a second; BASELINE.json's full sizes are covered by
property tests in test_gpu_parity.py.
"""


TARGET_RGBA, TARGET_BGRA, TARGET_RGB, TARGET_BGR = 0, 1, 2, 3     # variant bits 24-25 of every single-context scene


def micro_variant(blend=None, depth=None, flat=0, tex=0, bil=0, wrap=0, cull_off=0, mode=0, rgb=0, persp=0,
                  phong=0, fbo=0, spot=0, bgra=0, bgr=0, target=0):
    v = 0
    if blend is not None:
        v |= 8 | blend
    if depth is not None:
        v |= 128 | (depth << 4)
    v |= (flat << 8 | tex << 9 | bil << 10 | wrap << 11 | cull_off << 13 | mode << 14 | rgb << 17 | persp << 18
          | phong << 19 | fbo << 20 | spot << 21 | bgra << 22 | bgr << 23 | target << 24)
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
    # the application rewrites the vertex array in place between frames (static-geometry mirrors must follow, pfx.h)
    ("c2-textured-arrays-rewritten-f2", "textured", 640, 360, dict(size=64, variant=32 | 128, first_frame=0, frames=3), False),
    ("c3-phong", "phong", 640, 360, dict(size=96), False),
    ("c3-phong-arrays", "phong", 320, 200, dict(size=48, variant=32), False),
    ("c4-overdraw-add", "overdraw", 512, 256, dict(size=8), False),
    ("c4-overdraw-alpha-depth", "overdraw", 512, 256, dict(size=8, variant=1), False),
    ("c4-overdraw-alpha-depth-bilinear", "overdraw", 256, 128, dict(size=4, variant=3), True),
    ("c5-batch", "batch", 256, 256, dict(size=3), False),
    # the overdraw scene off its most specialised fragment program (bench extras ns4k_*): tinted corners, CLAMP_TO_EDGE,
    # RGB8 texture, two fragment states in one batch
    ("c4-overdraw-alpha-depth-tinted", "overdraw", 512, 256, dict(size=8, variant=1 | 4), False),
    ("c4-overdraw-alpha-depth-clamp", "overdraw", 512, 256, dict(size=8, variant=1 | 8), False),
    ("c4-overdraw-alpha-depth-rgb8", "overdraw", 512, 256, dict(size=8, variant=1 | 16), False),
    ("c4-overdraw-alpha-depth-two-state", "overdraw", 512, 256, dict(size=8, variant=1 | 32), False),
    ("c4-overdraw-add-screen-tinted-clamp-rgb8", "overdraw", 512, 256, dict(size=8, variant=4 | 8 | 16 | 32), False),
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



# render-list replay in the states that decide how a device-resident list is replayed - both face
# passes / front-face culling (one assembled form per face mode), with and without PF_COLOR_MATERIAL (pfColor feeds the
# material and the vertices take the current colour, or the recorded colours are used), per-pixel Phong, a texture
# matrix at replay (not expressible on the device: immediate-mode replay), a list recorded again between frames
CASES += [("c5-batch-cull-off", "batch", 256, 256, dict(size=2, variant=1), False),
          ("c5-batch-cull-front-nocolormat", "batch", 256, 256, dict(size=2, variant=2 | 16), False),
          ("c5-batch-recorded-colours", "batch", 256, 256, dict(size=2, variant=2 | 4), False),
          ("c5-batch-recorded-colours-colormat", "batch", 256, 256, dict(size=2, variant=4), False),
          ("c5-batch-phong-cull-off", "batch", 200, 160, dict(size=2, variant=1 | 32), False),
          ("c5-batch-texmatrix", "batch", 256, 256, dict(size=2, variant=8), False),
          ("c5-batch-rerecorded-f2", "batch", 256, 256, dict(size=2, variant=64, first_frame=0, frames=3), False)]

# BGRA8 / BGR8 textures and BGRA8 / RGB8 / BGR8 render targets against the LIVE reference.  The
# reference's BGRA8 getter and setter hand the first pixel of every group of four (counted from the triangle's xMin) to
# the whole group (Q19); these cases pin that behaviour - texel replication with per-pixel bilinear weights, the
# leader's own mask deciding its uv, blending against the leader's pixel, stores of the leader's fragment - under
# every blend mode, with depth testing, perspective, Phong, flat shading and an FBO of the same layout sampled back.
CASES += [_micro(f"bgra-tex-wrap{w}", 41 + w, tex=1, bgra=1, wrap=w, cull_off=1, blend=1) for w in range(3)]
CASES += [_micro(f"bgra-tex-persp-depth-wrap{w}", 44 + w, tex=1, bgra=1, wrap=w, cull_off=1, persp=1, depth=2) for w in range(3)]
CASES += [_micro(f"bgra-tex-bilinear-wrap{w}", 47 + w, ref_bfix=True, tex=1, bgra=1, bil=1, wrap=w, cull_off=1, depth=3) for w in range(3)]
CASES += [_micro("bgra-tex-phong", 50, tex=1, bgra=1, phong=1, persp=1, cull_off=1, depth=2, spot=1)]
CASES += [_micro(f"bgr-tex-wrap{w}", 51 + w, tex=1, bgr=1, wrap=w, cull_off=1, blend=1) for w in range(3)]
CASES += [_micro("bgr-tex-bilinear-persp", 54, ref_bfix=True, tex=1, bgr=1, bil=1, cull_off=1, persp=1, depth=2)]
CASES += [_micro(f"target-bgra-blend{b}", 60 + b, target=TARGET_BGRA, blend=b, cull_off=1) for b in range(8)]
CASES += [_micro(f"target-bgra-depth{d}", 68 + d, target=TARGET_BGRA, depth=d, cull_off=1, flat=d & 1) for d in range(6)]
CASES += [_micro("target-bgra-tex-persp", 74, target=TARGET_BGRA, tex=1, wrap=1, cull_off=1, persp=1, depth=2, blend=1),
          _micro("target-bgra-bgra-tex", 75, target=TARGET_BGRA, tex=1, bgra=1, cull_off=1, blend=1, depth=3),
          _micro("target-bgra-bilinear", 76, ref_bfix=True, target=TARGET_BGRA, tex=1, bil=1, cull_off=1, blend=2),
          _micro("target-bgra-phong", 77, target=TARGET_BGRA, phong=1, persp=1, cull_off=1, depth=2),
          _micro("target-bgra-fbo", 78, target=TARGET_BGRA, fbo=1, tex=1, cull_off=1, depth=2),
          _micro("target-bgra-fbo-persp-blend", 79, target=TARGET_BGRA, fbo=1, tex=1, cull_off=1, depth=2, persp=1, blend=1)]
CASES += [_micro(f"target-bgra-mode{m}", 80 + m, target=TARGET_BGRA, mode=m, cull_off=m & 1, depth=2, blend=1) for m in range(6)]
CASES += [_micro(f"target-rgb-blend{b}", 90 + b, target=TARGET_RGB, blend=b, cull_off=1, tex=b & 1) for b in range(8)]
CASES += [_micro(f"target-bgr-blend{b}", 100 + b, target=TARGET_BGR, blend=b, cull_off=1, depth=b % 6) for b in (0, 1, 2, 5)]
CASES += [_micro("target-rgb-fbo", 110, target=TARGET_RGB, fbo=1, tex=1, cull_off=1, depth=2, blend=1),
          _micro("target-bgr-phong-persp", 111, target=TARGET_BGR, phong=1, persp=1, cull_off=1, depth=2)]


# breadth of the public API around the triangle path (scene "api", see scenes.c for the variant bits):
# viewport offsets, texture matrix, Gouraud with spot/attenuation/back materials/colour material/normalize,
# colour arrays, pfRect*, pfDrawPixels + zoom, fog, post-processing, pfReadPixels, pfClearDepth, aux buffer
def _api(desc, variant, seed=1, ref_bfix=False, target=0):
    return (f"api-{desc}", "api", 200, 150, dict(variant=variant | target << 24, seed=seed), ref_bfix)


_B = lambda *bits: sum(1 << b for b in bits)
CASES += [_api("plain", 0), _api("viewport", _B(0)), _api("texmatrix", _B(1)), _api("gouraud", _B(2)),
          _api("gouraud-backmat", _B(2, 3)), _api("gouraud-backmat-colormat", _B(2, 3, 12)),
          _api("gouraud-normalize", _B(2, 11)), _api("arrays-colorptr", _B(13)),
          _api("gouraud-arrays-normalize-colormat", _B(2, 11, 12, 13)), _api("cullfront", _B(10)),
          _api("rects", _B(4)), _api("rects-viewport", _B(0, 4)), _api("drawpixels", _B(5)),
          _api("drawpixels-viewport-blend", _B(0, 5, 17)), _api("readpixels", _B(8)),
          _api("fog-linear", _B(6)), _api("fog-exp-opaque", _B(6, 15, 16)), _api("fog-cleardepth", _B(6, 9)),
          _api("fog-exp", _B(6, 19)), _api("fog-exp2-opaque", _B(6, 16, 20), seed=7), _api("fog-invalid-mode", _B(6, 19, 20), seed=8),
          _api("fog-exp-blend-viewport", _B(0, 6, 17, 19), seed=9), _api("pixel-layouts", _B(21), seed=10), _api("pixel-layouts-viewport", _B(0, 21), seed=11),
          _api("pixel-formats-all", _B(22), seed=15), _api("pixel-formats-all-gouraud-viewport", _B(0, 2, 22), seed=16),
          _api("postprocess", _B(7)), _api("swapbuffers", _B(14, 7)),
          _api("everything", 0x3ffff & ~_B(13), seed=2), _api("everything-bilinear", 0x7ffff & ~_B(13), seed=3, ref_bfix=True)]
# the same API breadth on the other target layouts (scalar getters / setters: rects, draw pixels, fog, post-processing,
# read pixels, swap buffers) and points / lines below
CASES += [_api("everything-target-bgra", 0x3ffff & ~_B(13), seed=4, target=TARGET_BGRA),
          _api("everything-target-rgb", 0x3ffff & ~_B(13), seed=5, target=TARGET_RGB),
          _api("gouraud-target-bgr", _B(2, 3), seed=6, target=TARGET_BGR),
          _api("fog-exp-layouts-target-bgra", _B(4, 5, 6, 19, 21), seed=12, target=TARGET_BGRA),
          _api("fog-exp2-layouts-target-rgb", _B(4, 5, 6, 20, 21), seed=13, target=TARGET_RGB),
          _api("fog-layouts-target-bgr", _B(4, 5, 6, 8, 21), seed=14, target=TARGET_BGR),
          _api("pixel-formats-all-target-bgra", _B(22), seed=17, target=TARGET_BGRA)]

CASE_IDS = [c[0] for c in CASES]
assert len(set(CASE_IDS)) == len(CASE_IDS)


# points, lines and PF_POINT / PF_LINE polygon modes (scene "prims"): the reference's scalar rasterisers
# (lines.c, points.c) - scalar blend / depth tables, thick lines, frustum-clipped 3D lines
def _prims(desc, blend=None, depth=None, persp=0, thick=0, seed=1, size=48, target=0):
    v = (persp << 8) | (thick << 9) | (target << 24)
    if blend is not None:
        v |= 1 | (blend << 1)
    if depth is not None:
        v |= 16 | (depth << 5)
    return (f"prims-{desc}", "prims", 200, 150, dict(variant=v, seed=seed, size=size), False)


CASES += [_prims("plain"), _prims("thick", thick=1, seed=2), _prims("persp", persp=1, seed=3), _prims("persp-depth-less", persp=1, depth=2, seed=4)]
CASES += [_prims(f"blend{b}-thick", blend=b, thick=1, seed=5 + b) for b in range(8)]
CASES += [_prims(f"depth{d}-thick", depth=d, thick=1, seed=20 + d) for d in range(6)]
CASES += [_prims("persp-blend1-depth3-points", blend=1, depth=3, persp=1, thick=1, seed=30, size=90)]
CASES += [_prims("target-bgra-blend1-thick", blend=1, thick=1, seed=31, target=TARGET_BGRA),
          _prims("target-rgb-blend2-depth2", blend=2, depth=2, seed=32, target=TARGET_RGB),
          _prims("target-bgr-persp-blend0", blend=0, persp=1, seed=33, target=TARGET_BGR)]

CASE_IDS = [c[0] for c in CASES]

# textures in every (format, type) pair besides the four 8-bit layouts (scene "texfmt": the micro scene sampling a 53x29
# texture of that pair; size = format * 16 + type): the reference's SIMD texel getters (pixel.h:2249-3040), including
# what they do with out-of-range half / float components
_PF_FORMATS = dict(red=0, green=1, blue=2, alpha=3, lum=4, luma=5, rgb=6, rgba=7, bgr=8, bgra=9)
_PF_TYPES = dict(ubyte=0, s565=2, s5551=3, s4444=4, half=9, float=10)


def texture_pairs():
    out = []
    for f, fv in _PF_FORMATS.items():
        comps = 1 if fv <= 4 else 2 if fv == 5 else 3 if fv in (6, 8) else 4
        for t, tv in _PF_TYPES.items():
            if (tv == 2 and comps != 3) or (tv in (3, 4) and comps != 4) or (tv == 0 and comps >= 3):
                continue
            out.append((f"{f}-{t}", fv * 16 + tv))
    return out


def _texfmt(name, code, k):
    nearest = micro_variant(tex=1, wrap=k % 3, blend=(None, 1, 5, 2)[k % 4], depth=(None, 1)[k % 2], persp=(k // 2) % 2, cull_off=k % 2)
    bilinear = micro_variant(tex=1, bil=1, wrap=(k + 1) % 3, blend=(1, None, 0)[k % 3], persp=k % 2, cull_off=1)
    return [(f"texfmt-{name}-nearest", "texfmt", 160, 120, dict(variant=nearest, seed=40 + k, size=code), False),
            (f"texfmt-{name}-bilinear", "texfmt", 160, 120, dict(variant=bilinear, seed=80 + k, size=code), True)]


for _k, (_name, _code) in enumerate(texture_pairs()):
    CASES += _texfmt(_name, _code, _k)

CASE_IDS = [c[0] for c in CASES]
assert len(set(CASE_IDS)) == len(CASE_IDS)

# the corners of the public API no other scene touches (scene "conform"): every pfGet*v getter over every PFgettable /
# PFstate / invalid name, sticky error codes, all pfColor* / pfVertex* / pfRasterPos* / pfRect* argument variants,
# pfFogfv, the framebuffer pixel accessors, pfClearFramebuffer, pfGetTexturePixels - what the getters return is drawn
# into the frame, so the colour / depth comparison covers it
def _conform(desc, variant, seed, target=0):
    return (f"conform-{desc}", "conform", 200, 224, dict(variant=variant | (target << 24), seed=seed), False)


CASES += [_conform("plain", 0, 1), _conform("blend-depth", 1, 2), _conform("target-bgra", 1, 3, TARGET_BGRA),
          _conform("target-rgb", 0, 4, TARGET_RGB), _conform("target-bgr", 1, 5, TARGET_BGR)]

CASE_IDS = [c[0] for c in CASES]
assert len(set(CASE_IDS)) == len(CASE_IDS)

# the call sequences of the reference's example programs (scene "examples": examples/common.h helpers, raylib_2D / _3D /
# _Framebuffer / _Points / _ModelWires / _TextureMatrix / _Texture2D / _FirstPerson) and vertex arrays in every component
# and index type pfDrawElements / pfDrawArrays accept
_EXAMPLES = {0: "2d-initial-state", 1: "3d-cube", 2: "framebuffer-drawpixels", 3: "points", 4: "wires-ushort-indices",
             5: "texture-matrix-ground", 6: "texture2d-sprites-luma", 8: "firstperson-spotlight", 9: "arrays-all-types",
             10: "loose-begin-end"}


def _example(which, frame, target=0):
    tname = {0: "", 1: "-target-bgra", 2: "-target-rgb", 3: "-target-bgr"}[target]
    return (f"examples-{_EXAMPLES[which]}-f{frame}{tname}", "examples", 320, 240,
            dict(variant=which | (target << 24), seed=1 + which, first_frame=frame), False)


CASES += [_example(k, f) for k in _EXAMPLES for f in (0, 3)]
CASES += [_example(2, 1, TARGET_BGRA), _example(6, 1, TARGET_BGRA), _example(9, 1, TARGET_RGB), _example(8, 1, TARGET_BGR)]

CASE_IDS = [c[0] for c in CASES]
assert len(set(CASE_IDS)) == len(CASE_IDS)

# a batch whose states differ in their texel layout (one of the other layouts + RGBA8: no one-program kernel, the run-time
# sampler of the tile rasteriser), and render lists that sample a 5-6-5 / BGRA8 texture (replayed through the ordinary path)
CASES += [("texfmt-rgba-s4444-plus-rgba8-blend", "texfmt", 160, 120, dict(variant=micro_variant(tex=1, blend=1) | (1 << 27), seed=3, size=7 * 16 + 4), False),
          ("texfmt-lum-half-plus-rgba8-persp-depth", "texfmt", 160, 120, dict(variant=micro_variant(tex=1, wrap=2, persp=1, depth=1) | (1 << 27), seed=4, size=4 * 16 + 9), False),
          ("texfmt-rgb-s565-plus-rgba8-bilinear", "texfmt", 160, 120, dict(variant=micro_variant(tex=1, bil=1, blend=1) | (1 << 27), seed=5, size=6 * 16 + 2), True),
          ("c5-batch-s565-textures", "batch", 256, 256, dict(size=2, variant=128), False),
          ("c5-batch-bgra8-textures", "batch", 256, 256, dict(size=2, variant=256), False),
          ("c5-batch-s565-textures-cull-off-recorded-colours", "batch", 256, 256, dict(size=2, variant=128 | 1 | 4), False)]

CASE_IDS = [c[0] for c in CASES]
assert len(set(CASE_IDS)) == len(CASE_IDS)

# all eight lights at once (Gouraud on the device through the specular threshold tables, and per-fragment Phong), and the
# three matrix stacks driven into overflow / underflow with nested transforms around textured quads and a 3D pass
CASES += [("examples-eight-lights-gouraud-f0", "examples", 320, 240, dict(variant=11, seed=2, first_frame=0), False),
          ("examples-eight-lights-gouraud-f2", "examples", 320, 240, dict(variant=11, seed=2, first_frame=2), False),
          ("examples-eight-lights-phong-f0", "examples", 320, 240, dict(variant=11 | 16, seed=2, first_frame=0), False),
          ("examples-eight-lights-phong-f2-target-bgra", "examples", 320, 240, dict(variant=11 | 16 | (TARGET_BGRA << 24), seed=2, first_frame=2), False),
          ("examples-matrix-stacks-f0", "examples", 320, 240, dict(variant=12, seed=2, first_frame=0), False),
          ("examples-matrix-stacks-f2", "examples", 320, 240, dict(variant=12, seed=2, first_frame=2), False)]

CASE_IDS = [c[0] for c in CASES]
assert len(set(CASE_IDS)) == len(CASE_IDS)
