"""Run on the GPU box: shaded-pixel / triangle counts per bench workload from the product's device counters.
Writes gpurun_out/workload_counts.json (copy to tests/golden/)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS
from pixelforge_b200 import load_product_scenes
p = load_product_scenes(); out = {}
for name, wl in WORKLOADS.items():
    _, _, r = p.render(wl["scene"], wl["w"], wl["h"], variant=wl["variant"], size=wl["size"], want_depth=False)
    out[name] = {"pixels_shaded": r.pixels_shaded, "pixels_depth_failed": r.pixels_depth_failed,
                 "triangles_submitted": r.triangles_submitted, "triangles_rasterised": r.triangles_rasterised}
    print(name, out[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "workload_counts.json"), "w"), indent=1, sort_keys=True)
