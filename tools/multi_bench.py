"""Single-process multi-device mode (PF_CUDA_DEVICES) on one bench workload through the PUBLIC API: wall-clock per frame of
    render       pfClear + draw calls + pfxFlush, all devices idle again (no presentation)
    present      ... + the finished tiles gathered on device 0 (what pfReadPixels of one pixel forces)
    e2e          ... + read-back of the whole frame into the caller's buffer (pfxFinish)
Run:  PF_CUDA_DEVICES=0,1,2,3 python tools/multi_bench.py c4_overdraw_8k [steps]     -> one JSON line"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from pixelforge_b200 import load_product_scenes, load_pfcu
from pixelforge_b200.binding import Counters

name = sys.argv[1] if len(sys.argv) > 1 else "c4_overdraw_8k"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
wl = bench.WORKLOADS[name]
scenes = load_product_scenes(); pfcu = load_pfcu("product"); L = pfcu.lib
L.pfcu_surface_read_pixels.restype = C.c_int
L.pfcu_surface_read_pixels.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
devs = os.environ.get("PF_CUDA_DEVICES", "0")
out = {"workload": name, "devices": devs, "n_devices": len(devs.split(",")), "steps": steps, "timing": "wall clock around the public API calls, all devices idle before and after"}
one = np.zeros(1, np.uint32)
with scenes.open(wl["scene"], wl["w"], wl["h"], variant=wl["variant"], size=wl["size"], explicit_sync=1) as sc:
    L.pfxEnableQueuedReadback(0)
    surf = L.pfxGetSurfaceHandle()

    def render():
        sc.frame(0); L.pfxFlush(); L.pfcu_finish()

    def present():
        sc.frame(0); L.pfxFlush()
        pfcu.check(L.pfcu_surface_read_pixels(surf, 8, 0, 1, 1, 1, 7 * 16 + 0, one.ctypes.data), "read_pixels")      # PFCU_PIX(PF_RGBA, PF_UNSIGNED_BYTE)
        L.pfcu_finish()

    def e2e():
        sc.frame(0); sc.finish()

    for key, fn in (("render", render), ("present", present), ("e2e", e2e)):
        for _ in range(2):
            fn()
        L.pfxResetCounters()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        ms = (time.perf_counter() - t0) / steps * 1e3
        k = Counters(); L.pfcu_get_counters(k)
        out[key + "_ms"] = ms
        out[key + "_gpix_per_s"] = k.pixels_shaded / steps / (ms * 1e-3) / 1e9
    out["shaded_px_per_step"] = k.pixels_shaded / steps
    c, _ = sc.read_index(0, want_depth=False)
    import hashlib
    flat = c.copy().reshape(-1); flat[:8] = 0            # pfClear never clears pixels 0..7 (Q12): they depend on the frame count
    out["frame_sha256_16"] = hashlib.sha256(flat.tobytes()).hexdigest()[:16]
print(json.dumps(out))
