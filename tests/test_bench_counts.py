"""bench.py's reference arm has no shaded-pixel counters (the reference library exposes none), so it uses
tests/golden/workload_counts.json (measured by the product's device counters).  Pin that table against the
oracle wherever the scalar oracle finishes in seconds, and against closed forms for the overdraw scenes."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def counts():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "workload_counts.json")))


@pytest.mark.parametrize("name", ["c1_gears_800x600", "c2_textured_1080p"])
def test_counts_match_oracle(name, counts, oracle_scenes):
    import sys
    sys.path.insert(0, ROOT)
    from bench import WORKLOADS
    wl = WORKLOADS[name]
    _, _, r = oracle_scenes.render(wl["scene"], wl["w"], wl["h"], variant=wl["variant"], size=wl["size"], want_depth=False)
    for k in ("pixels_shaded", "pixels_depth_failed", "triangles_submitted", "triangles_rasterised"):
        assert getattr(r, k) == counts[name][k], k


def test_overdraw_closed_form(counts, oracle_scenes):
    """Each full-screen quad layer shades (W-1)*H pixels plus the pixels on the shared diagonal (hit by both
    triangles, Q4); the count per layer is measured on a small surface with the same aspect logic and the
    committed 8K/4K numbers must be exactly layers * per-layer."""
    for name, w, h in (("c4_overdraw_8k", 7680, 4320), ("ns_textured_blend_4k", 3840, 2160), ("ns4k_tinted", 3840, 2160), ("ns4k_clamp", 3840, 2160),
                       ("ns4k_rgb8", 3840, 2160), ("ns4k_two_state", 3840, 2160), ("ns4k_bilinear", 3840, 2160)):
        per_layer = counts[name]["pixels_shaded"] // 64
        assert per_layer * 64 == counts[name]["pixels_shaded"]
        assert (w - 1) * h <= per_layer <= (w - 1) * h + w + h
    _, _, r = oracle_scenes.render("overdraw", 96, 54, size=3, want_depth=False)
    assert r.pixels_shaded % 3 == 0 and (96 - 1) * 54 <= r.pixels_shaded // 3 <= 95 * 54 + 96 + 54


def test_bench_gate_plumbing_on_cpu(oracle_scenes, ref_scenes, tmp_path):
    """bench.py's reference-equivalence gate end to end, without a GPU: the CPU arm dumps the frames it rendered
    (subprocess, as in the bench) and parity_against_dump compares a render of the same workload - here by the oracle
    build standing in for the product - colour and depth, bit for bit.  A corrupted dump must be reported."""
    import subprocess, sys
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    dump = str(tmp_path / "ref.npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--as-baseline", "--workload", "c1_gears_800x600",
                        "--baseline-frames", "1", "--dump", dump], capture_output=True, text=True, timeout=300)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["config"] == bench.bench_config("c1_gears_800x600", 1)
    p = bench.parity_against_dump(oracle_scenes, "c1_gears_800x600", dump)
    assert p["differing_px"] == 0 and p["differing_depth"] == 0 and p["pixels_compared"] == 800 * 600
    z = np.load(dump)
    c = z["color"].copy(); c[0, 300, 400] ^= 1
    bad = str(tmp_path / "bad.npz")
    np.savez(bad, color=c, depth=z["depth"], frames_rendered=z["frames_rendered"])
    assert bench.parity_against_dump(oracle_scenes, "c1_gears_800x600", bad)["differing_px"] == 1
