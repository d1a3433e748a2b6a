"""Device timeline of end-to-end frames (PF_CUDA_TIMING=1 prints it from the library): python tools/e2e_timeline.py [workload ...]"""
import os, sys, time
os.environ["PF_CUDA_TIMING"] = "1"
os.environ.setdefault("PFSCENE_STATIC_ARRAYS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pixelforge_b200 import load_product_scenes
scenes = load_product_scenes()
for name in sys.argv[1:] or ["c3_phong_4k", "c2_textured_1080p"]:
    wl = bench.WORKLOADS[name]
    with scenes.open(wl["scene"], wl["w"], wl["h"], variant=wl["variant"], size=wl["size"], explicit_sync=1) as sc:
        for i in range(6):
            sc.frame(0); sc.finish()
        sys.stderr.flush()
        print("==", name, "(3 frames)", file=sys.stderr, flush=True)
        for i in range(3):
            t0 = time.perf_counter(); sc.frame(0); t1 = time.perf_counter(); sc.finish(); t2 = time.perf_counter()
            print("   frame() %.3f ms  finish() %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), file=sys.stderr, flush=True)
