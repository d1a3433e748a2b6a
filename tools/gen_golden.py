"""Generate tests/golden/scenes_golden.json from the UNMODIFIED reference built by oracle/build_ref.sh.

Run in the container that has /root/reference:   python tools/gen_golden.py
For every parity case in tests/cases.py the reference renders the scene (OMP_NUM_THREADS=1; the
OpenMP build is deterministic across thread counts except for its row-end store race)
and we record SHA-256 of the colour and depth buffers plus a few counts.  Bilinear cases use the
reference rebuilt with the one-token fix of SURVEY Q7 (upstream bilinear reads an uninitialised
vector).  The hashes depend on the host's libm and RCPPS/RSQRTPS tables; the file records a
fingerprint of those tables so that tests on a different CPU skip the hash comparison.
"""
import hashlib, json, os, sys
os.environ["OMP_NUM_THREADS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
sys.path.insert(0, os.path.join(ROOT, "tests"))
from checkers import load_reference_scenes, load_oracle_pfcu
from cases import CASES


def table_fingerprint():
    lib = load_oracle_pfcu()
    rcp, rb, rsq, sb = lib.harvest_tables()
    a = np.ctypeslib.as_array(rcp, shape=(1 << rb,)).tobytes() + np.ctypeslib.as_array(rsq, shape=(2 << sb,)).tobytes()
    return {"rcp_bits": rb, "rsqrt_bits": sb, "sha256": hashlib.sha256(a).hexdigest()}


def main():
    ref, bfix = load_reference_scenes(False), load_reference_scenes(True)
    out = {"host_tables": table_fingerprint(), "cases": {}}
    for cid, scene, w, h, kw, needs_fix in CASES:
        color, depth, _ = (bfix if needs_fix else ref).render(scene, w, h, **kw)
        out["cases"][cid] = {
            "color_sha256": hashlib.sha256(color.tobytes()).hexdigest(),
            "depth_sha256": hashlib.sha256(depth.tobytes()).hexdigest(),
            "nonzero_rgb": int(((color & 0xFFFFFF) != 0).sum()),
            "depth_written": int((depth != np.finfo(np.float32).max).sum()),
            "reference": "bilinear-fix" if needs_fix else "verbatim",
        }
        print(cid, out["cases"][cid]["color_sha256"][:12], out["cases"][cid]["nonzero_rgb"])
    # a few small full-resolution fixtures for debugging mismatches
    small = {}
    for cid, scene, w, h, kw, needs_fix in CASES:
        if cid in ("micro-blend1", "micro-depth2", "micro-tex-persp-wrap0-rgb0", "micro-phong-spot1", "micro-bilinear-wrap0"):
            color, depth, _ = (bfix if needs_fix else ref).render(scene, w, h, **kw)
            small[cid + ".color"] = color; small[cid + ".depth"] = depth
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "micro_fixtures.npz"), **small)
    with open(os.path.join(ROOT, "tests", "golden", "scenes_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
