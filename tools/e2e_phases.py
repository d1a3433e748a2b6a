"""Wall-clock split of the end-to-end step of a workload: API calls (host vertex stage + enqueue) vs finish (wait + D2H)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pixelforge_b200 import load_product_scenes
scenes = load_product_scenes()
for name in sys.argv[1:] or ["c5_batch_512", "c2_textured_1080p", "c1_gears_800x600", "c3_phong_4k"]:
    wl = bench.WORKLOADS[name]
    with scenes.open(wl["scene"], wl["w"], wl["h"], variant=wl["variant"], size=wl["size"], explicit_sync=1) as sc:
        for i in range(5):
            sc.frame(0); sc.finish()
        tf = tn = 0.0; N = 20
        for i in range(N):
            t0 = time.perf_counter(); sc.frame(0); t1 = time.perf_counter(); sc.finish(); t2 = time.perf_counter()
            tf += t1 - t0; tn += t2 - t1
        print("%-22s frame() %.3f ms   finish() %.3f ms" % (name, tf / N * 1e3, tn / N * 1e3), flush=True)
