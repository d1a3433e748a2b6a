#!/usr/bin/env python
"""bench.py - headline benchmark of pixelforge-b200 (contract: see the task's bench.py section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one frame of the workload: pfClear + the scene's draw calls (the per-fragment hot path
over one batch of synthetic triangles).  Headline workload = BASELINE.json configs[1]: a textured
model at 1920x1080 with bilinear filtering, alpha blending and depth test.

Legs of the default (ours) arm, all printed in ONE JSON line by rank 0:
  value      shaded Gpix/s with the frame's triangle stream already resident in HBM: CUDA events on
             the launching stream around {clear, setup, bin, raster}; L2 flushed between iterations.
  e2e        the same metric through the public pixelforge.h API with HOST buffers: the application's API calls,
             H2D of the step's inputs from pinned host memory (assembled-triangle batches; vertex / index arrays
             that the scene allocated with pfxHostAlloc), device vertex stage + rasterisation, D2H of the
             framebuffer into the caller's (page-locked) buffer - wall clock around pfClear..pfxFinish.
  roofline   k_raster (the dominant kernel): algorithmic bytes / its CUDA-event time.
  cpu_baseline  the reference's own OpenMP+AVX2 code (oracle/_ref, compiled from /root/reference)
             on this box's host cores, bounded sample, in a subprocess.
  extra      the other BASELINE.json configs (C1, C3, C4, the 4K textured+blended target scene, C5).
`--impl reference` times the UNMODIFIED reference library on the same workload (rank 0 only).
With --gpus N > 1 (torchrun) every rank renders its own context (independent contexts, one per GPU,
no data-path collective: weak scaling); the C4 overdraw scene is additionally run screen-tile split
with an NCCL gather of the tiles to rank 0 and reported under extra.tile_split.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# workload table: scene, size, variant, resolution, algorithmic bytes per shaded pixel (SURVEY 8-d:
# 4[depth test] + 4 (z write) + 4[blend] + 4 (colour write) + 4*taps[texture])
WORKLOADS = {
    "c2_textured_1080p": dict(scene="textured", w=1920, h=1080, size=256, variant=1 | 32 | 64, bytes_px=32,
                              desc="C2: textured torus (65,536 quads*2 faces), 1920x1080, bilinear REPEAT, alpha blend, depth LESS, vertex arrays"),
    "c1_gears_800x600": dict(scene="gears", w=800, h=600, size=0, variant=0, bytes_px=12,
                             desc="C1: Gears, 800x600, immediate mode, Gouraud, depth LESS"),
    "c3_phong_4k": dict(scene="phong", w=3840, h=2160, size=708, variant=32, bytes_px=12,
                        desc="C3: 1,002,528-triangle height field, 3840x2160, per-pixel Blinn-Phong, depth LESS"),
    "c4_overdraw_8k": dict(scene="overdraw", w=7680, h=4320, size=64, variant=0, bytes_px=16,
                           desc="C4: 64 layers of additive-blended textured full-screen quads, 7680x4320, nearest REPEAT"),
    "ns_textured_blend_4k": dict(scene="overdraw", w=3840, h=2160, size=64, variant=1, bytes_px=20,
                                 desc="north-star scene: 64 layers textured + alpha-blended + depth-tested quads, 3840x2160, nearest"),
    # the north-star scene OFF its most specialised fragment program: same 64 layers at 3840x2160, alpha blend + depth test
    "ns4k_tinted": dict(scene="overdraw", w=3840, h=2160, size=64, variant=1 | 4, bytes_px=20,
                        desc="north-star scene with a different vertex colour per corner (smooth-shaded tint times texel)"),
    "ns4k_clamp": dict(scene="overdraw", w=3840, h=2160, size=64, variant=1 | 8, bytes_px=20,
                       desc="north-star scene with CLAMP_TO_EDGE wrapping"),
    "ns4k_rgb8": dict(scene="overdraw", w=3840, h=2160, size=64, variant=1 | 16, bytes_px=20,
                      desc="north-star scene with an RGB8 (3 bytes per texel) texture"),
    "ns4k_two_state": dict(scene="overdraw", w=3840, h=2160, size=64, variant=1 | 32, bytes_px=20,
                           desc="north-star scene with two fragment states in one batch (even layers alpha-blend, odd layers add)"),
    "ns4k_bilinear": dict(scene="overdraw", w=3840, h=2160, size=64, variant=1 | 2, bytes_px=32,
                          desc="north-star scene with bilinear filtering (4 taps)"),
    "c5_batch_512": dict(scene="batch", w=512, h=512, size=32, variant=0, bytes_px=16,
                         desc="C5: independent 512x512 contexts (render-list replay, textured + Gouraud lit, depth), 32 per GPU"),
}
HEADLINE = "c2_textured_1080p"
METRIC = "shaded_gpix_per_s"
UNIT = "Gpix/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---- clocks sampling ------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- the reference arm / cpu baseline ---------------------------------------------------------------

def shaded_pixels_table():
    """Shaded pixels / triangles per frame of each workload (frame 0), measured by the product's device
    counters and asserted equal to the oracle's in tests/test_bench_counts.py.  The reference library
    has no counters, so its arm uses this table (identical coverage is the parity gate)."""
    with open(os.path.join(ROOT, "tests", "golden", "workload_counts.json")) as f:
        return json.load(f)


def is_bilinear(wl):
    return (wl["scene"] == "textured" and bool(wl["variant"] & 1)) or (wl["scene"] == "overdraw" and bool(wl["variant"] & 2))


def gate_size(wl):
    """Scene size of the BOUNDED sample both the CPU arm and the reference-equivalence gate use: every context of C5,
    full frames of C1-C3, a 4-layer slice of the 64-layer overdraw scenes (the work per layer is identical)."""
    return min(wl["size"], 4) if wl["scene"] == "overdraw" else wl["size"]


def bench_config(wl_name, n_gpus):
    """The `config` object of the JSON line - the same in both arms (the driver compares them)."""
    wl = WORKLOADS[wl_name]
    counts = shaded_pixels_table().get(wl_name, {})
    return {"workload": wl_name, "description": wl["desc"], "width": wl["w"], "height": wl["h"],
            "triangles_per_step": counts.get("triangles_submitted"), "shaded_px_per_step": counts.get("pixels_shaded"),
            "parallelism": f"independent contexts x{n_gpus}",
            "l2_policy": "GPU arm: 256 MiB buffer written between timed iterations (untimed); CPU arm: n/a",
            "reference_library": "bilinear workloads: reference rebuilt with the one-token fix of its uninitialised-vector bug "
                                 "(src/internal/color.h:141, SURVEY Q7) in BOTH arms; everything else: unmodified reference"}


def run_reference(args, wl_name, bounded_frames=None, dump=None):
    """Times the reference's own OpenMP+AVX2 implementation (oracle/_ref, compiled from /root/reference) on frame 0 of
    the workload, all host threads.  With `dump`, the colour and depth of every context after the last frame are
    written there (np.savez): the GPU arm renders the same frame and compares (the reference-equivalence gate)."""
    import numpy as np
    # all host threads (torchrun exports OMP_NUM_THREADS=1 to its children; only rank 0 runs this arm)
    os.environ["OMP_NUM_THREADS"] = os.environ.get("PF_REF_THREADS", str(os.cpu_count()))
    os.environ.setdefault("OMP_WAIT_POLICY", "active")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from checkers import load_reference_scenes          # CPU arm only: the reference itself, never on the product path
    wl = WORKLOADS[wl_name]
    bfix = is_bilinear(wl)
    lib = load_reference_scenes(bfix)
    steps = bounded_frames if bounded_frames else args.steps
    warm = 1 if bounded_frames else args.warmup
    size = gate_size(wl)
    # bounded samples of the big workloads so that the CPU arm ends within minutes (work scales linearly)
    if wl["scene"] in ("overdraw", "phong", "batch") and not bounded_frames:
        steps, warm = min(steps, 3), min(warm, 1)
    sample = f"{steps} full frames (frame 0) after {warm} warm-up"
    if size != wl["size"]:
        sample += f", {size} of {wl['size']} layers"
    times = []
    with lib.open(wl["scene"], wl["w"], wl["h"], variant=wl["variant"], size=size) as sc:
        for i in range(warm + steps):
            t0 = time.perf_counter()
            sc.frame(0); sc.finish()
            if i >= warm:
                times.append((time.perf_counter() - t0) * 1e3)
        if dump:
            n_ctx = size if wl["scene"] == "batch" else 1
            frames = [sc.read_index(i, want_depth=True) for i in range(n_ctx)]
            # frames_rendered: pfClear never clears pixels 0..7, so with blending they depend on how many
            # frames were drawn into the buffer - the product renders the same number before it is compared
            np.savez(dump, color=np.stack([f[0] for f in frames]), depth=np.stack([f[1] for f in frames]), frames_rendered=warm + steps)
    counts = shaded_pixels_table().get(wl_name, {})
    scale = size / wl["size"] if wl["scene"] == "overdraw" else 1
    px = counts.get("pixels_shaded", 0) * scale
    tris = counts.get("triangles_submitted", 0) * scale
    ms = sum(times) / max(len(times), 1)
    return {"ms_per_step": ms, "gpix": px / (ms * 1e-3) / 1e9 if ms > 0 else 0.0, "mtri": tris / (ms * 1e-3) / 1e6 if ms > 0 else 0.0,
            "ms_median": sorted(times)[len(times) // 2] if times else 0.0, "cores": int(os.environ["OMP_NUM_THREADS"]), "sample": sample,
            "kind": "reference", "library": "oracle/_ref/" + ("libpf_ref_bfix.so (reference + the one-token Q7 bilinear fix)" if bfix else "libpf_ref.so (unmodified reference)")}


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl_name = args.workload
    try:
        r = run_reference(args, wl_name, bounded_frames=args.baseline_frames if args.as_baseline else None, dump=args.dump)
    except FileNotFoundError as e:
        emit_json({"impl": "reference", "unavailable": f"oracle/_ref not built: {e}"})
        return 0
    line = {
        "impl": "reference", "metric": METRIC, "value": r["gpix"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/i32/f32", "data": "synthetic",
        "config": bench_config(wl_name, args.gpus),
        "mtri_per_s": r["mtri"],
        "cpu_baseline": {"value": r["gpix"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": r["sample"], "library": r["library"]},
        "e2e": {"value": r["gpix"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)
    return 0


def parity_against_dump(scenes, wl_name, dump):
    """The reference-equivalence gate: render frame 0 of the workload at full size through the public
    API of the product and compare colour and depth, bit for bit, with what the reference rendered in this very run."""
    import numpy as np
    wl = WORKLOADS[wl_name]
    size = gate_size(wl)
    z = np.load(dump)
    ref_c, ref_d = z["color"], z["depth"]
    n_ctx = size if wl["scene"] == "batch" else 1
    dpx = dz = 0
    with scenes.open(wl["scene"], wl["w"], wl["h"], variant=wl["variant"], size=size, explicit_sync=1) as sc:
        for _ in range(int(z["frames_rendered"])):
            sc.frame(0); sc.finish()
        for i in range(n_ctx):
            c, d = sc.read_index(i, want_depth=True)
            dpx += int((c != ref_c[i]).sum())
            dz += int((d.view(np.uint32) != ref_d[i].view(np.uint32)).sum())
    out = {"differing_px": dpx, "differing_depth": dz, "pixels_compared": int(ref_c.size), "contexts": n_ctx, "frames_rendered": int(z["frames_rendered"]),
           "resolution": [wl["w"], wl["h"]]}
    if size != wl["size"]:
        out["layers"] = f"{size} of {wl['size']}"
    return out


# ---- our arm ---------------------------------------------------------------------------------------

def uses_vertex_arrays(wl):
    return wl["scene"] in ("textured", "phong") and bool(wl["variant"] & 32)


def measure_e2e_static(wl_name, steps, warmup, torch, scenes, pfcu):
    """End to end once more for the workloads that draw from vertex arrays, with the arrays declared static (pfxHostStatic:
    the application promises to announce changes, the library keeps the arrays in device memory): what a program with
    static meshes pays per frame - state, launches and the frame's read-back, no geometry over PCIe."""
    from pixelforge_b200.binding import Counters
    wl = WORKLOADS[wl_name]; L = pfcu.lib
    os.environ["PFSCENE_STATIC_ARRAYS"] = "1"
    try:
        with scenes.open(wl["scene"], wl["w"], wl["h"], variant=wl["variant"], size=wl["size"], explicit_sync=1) as sc:
            for i in range(max(warmup, 2)):
                sc.frame(0); sc.finish()
            L.pfxResetCounters()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(steps):
                sc.frame(0); sc.finish()
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / steps * 1e3
            k = Counters(); L.pfcu_get_counters(k)
            return {"ms_per_step": ms, "h2d_bytes_per_step": int(k.bytes_h2d / steps), "d2h_bytes_per_step": int(k.bytes_d2h / steps),
                    "gpix_per_s": k.pixels_shaded / steps / (ms * 1e-3) / 1e9, "mtri_per_s": k.triangles_submitted / steps / (ms * 1e-3) / 1e6,
                    "note": "vertex / index arrays declared static (pfxHostStatic): resident in HBM after the first frame"}
    finally:
        os.environ["PFSCENE_STATIC_ARRAYS"] = "0"


def measure_workload(wl_name, steps, warmup, torch, scenes, pfcu, stream, flush_buf, want_e2e=True, tile_owner=None):
    """Returns a dict with device-resident (`value`) and end-to-end numbers for one workload."""
    from pixelforge_b200.binding import Counters, Profile, STATE_DTYPE, TRIANGLE_DTYPE
    wl = WORKLOADS[wl_name]
    L = pfcu.lib
    out = {"workload": wl_name}
    if want_e2e and uses_vertex_arrays(wl):
        out["e2e_static"] = measure_e2e_static(wl_name, steps, warmup, torch, scenes, pfcu)
    with scenes.open(wl["scene"], wl["w"], wl["h"], variant=wl["variant"], size=wl["size"], explicit_sync=1) as sc:
        L.pfcu_set_stream(stream.cuda_stream)
        n_ctx = wl["size"] if wl["scene"] == "batch" else 1

        # ---- end to end through the public API (host buffers) ----
        if want_e2e:
            for i in range(max(warmup, 1)):
                sc.frame(0); sc.finish()
            L.pfxResetCounters()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(steps):
                sc.frame(0); sc.finish()
            torch.cuda.synchronize()
            e2e_s = (time.perf_counter() - t0) / steps
            k = Counters(); L.pfcu_get_counters(k)
            out.update(e2e_ms=e2e_s * 1e3, px_per_step=k.pixels_shaded / steps, tris_per_step=k.triangles_submitted / steps,
                       tris_rasterised_per_step=k.triangles_rasterised / steps, zfail_per_step=k.pixels_depth_failed / steps,
                       h2d_bytes=int(k.bytes_h2d / steps), d2h_bytes=int(k.bytes_d2h / steps))

        # ---- device-resident replay ----
        batches = []
        if wl["scene"] == "batch":
            # C5: the render lists ARE the device-resident input (assembled once, kept in HBM); a step is the public-API
            # replay of every context (a few KB of per-call tables cross PCIe) submitted as one multi-surface job list,
            # without the read-back of the 32 framebuffers that the end-to-end leg adds
            L.pfxEnableQueuedReadback(0)

            def step():
                sc.frame(0)
                L.pfxFlush()
        else:
            # capture one frame's triangle stream, keep it in HBM
            L.pfxCaptureBegin()
            sc.frame(0)
            states, tris = pfcu.capture_end()
            surf = L.pfxGetSurfaceHandle()
            if tile_owner:
                L.pfcu_surface_set_tile_owner(surf, tile_owner[0], tile_owner[1])
            b = L.pfcu_batch_upload(states.ctypes.data, len(states), tris.ctypes.data, len(tris))
            if not b:
                raise RuntimeError("pfcu_batch_upload failed: " + pfcu.error())
            batches.append((surf, b, len(tris)))
            sc.finish()
            clear_rgba, clear_z = 0xFF000000, 3.4028234663852886e38

            def step():
                for surf, b, _ in batches:
                    L.pfcu_surface_clear_ref(surf, 1, clear_rgba, 1, clear_z)
                    L.pfcu_batch_submit(surf, b)

        for i in range(max(warmup, 3)):
            step()
        L.pfcu_finish()
        L.pfxResetCounters()

        def timed_loop():
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            with torch.cuda.stream(stream):
                for i in range(steps):
                    flush_buf.fill_(i & 0xFF)          # L2 flush: write a buffer larger than L2 (untimed)
                    ev[i][0].record(stream)
                    L.pfcu_fence()                      # surfaces may live on other internal streams (lanes)
                    step()
                    L.pfcu_fence()
                    ev[i][1].record(stream)
            stream.synchronize()
            return sum(a.elapsed_time(b) for a, b in ev) / steps

        dev_ms = timed_loop()                           # the `value` leg: K steps, nothing but the step between the events
        k = Counters(); L.pfcu_get_counters(k)
        # the same K steps once more with the library's stage events on (CUDA events on the launching stream around the
        # front-end kernels and around the raster kernel): the per-kernel times of the roofline object
        L.pfcu_profile_enable(1)
        prof = Profile(); L.pfcu_profile_read(prof)
        prof_ms = timed_loop()
        L.pfcu_profile_read(prof)
        L.pfcu_profile_enable(0)
        out["dev_ms_with_stage_events"] = prof_ms
        out.update(dev_ms=dev_ms, dev_px_per_step=k.pixels_shaded / steps, dev_tris_per_step=k.triangles_submitted / steps,
                   raster_ms=prof.raster_ms / steps, frontend_ms=prof.frontend_ms / steps,
                   raster_launches_per_step=prof.raster_launches / steps, launches_per_step=k.kernel_launches / steps)
        for surf, b, _ in batches:
            L.pfcu_batch_destroy(b)
        L.pfxEnableQueuedReadback(1)
        sc.finish()
    return out


def single_process_scaling(world):
    """tools/multi_bench.py on 1 and on `world` devices (PF_CUDA_DEVICES), wall clock through the public API."""
    out = {"note": "one process, PF_CUDA_DEVICES=0..N-1, unmodified pixelforge.h calls; wall clock around the API calls; frame hash equal to one device's"}
    for name in ("c4_overdraw_8k", "ns_textured_blend_4k"):
        res = {}
        for n in (1, world):
            env = dict(os.environ); env["PF_CUDA_DEVICES"] = ",".join(str(i) for i in range(n)); env.pop("PF_CUDA_DEVICE", None)
            for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS"):
                env.pop(k, None)
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "multi_bench.py"), name, "5"], capture_output=True, text=True, timeout=300, env=env)
                res[n] = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception as e:
                res[n] = {"error": repr(e)}
        one, many = res.get(1, {}), res.get(world, {})
        o = {"one_device": one, f"{world}_devices": many}
        if "render_ms" in one and "render_ms" in many:
            o["scaling"] = "strong"
            for key in ("render", "present", "e2e"):
                o[f"{key}_speedup"] = one[f"{key}_ms"] / many[f"{key}_ms"]
                o[f"{key}_efficiency"] = one[f"{key}_ms"] / many[f"{key}_ms"] / world
            o["identical_frames"] = one.get("frame_sha256_16") == many.get("frame_sha256_16")
        out[name] = o
    return out


def ours_main(args):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("PF_CUDA_DEVICE", str(local))
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pixelforge_b200 import load_product_scenes, load_pfcu
    scenes = load_product_scenes(); pfcu = load_pfcu("product")
    stream = torch.cuda.Stream()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # 256 MiB > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t.item())

    wl_name = args.workload; wl = WORKLOADS[wl_name]
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    m = measure_workload(wl_name, args.steps, args.warmup, torch, scenes, pfcu, stream, flush_buf)
    barrier()
    clocks = None

    dev_ms = max_over_ranks(m["dev_ms"]); e2e_ms = max_over_ranks(m["e2e_ms"])
    px_all = sum_over_ranks(m["dev_px_per_step"]); tris_all = sum_over_ranks(m["dev_tris_per_step"])
    value = px_all / (dev_ms * 1e-3) / 1e9
    e2e_value = sum_over_ranks(m["px_per_step"]) / (e2e_ms * 1e-3) / 1e9
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    traffic = None; issue_active = None
    try:    # dram__bytes_read+write of the raster kernel per launch (and its issue utilisation: the north star's
            # evidence for setup/issue-bound scenes), from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            tj = json.load(f).get(wl_name, {})
            traffic = tj.get("traffic_bytes_per_launch"); issue_active = tj.get("sm_issue_active_pct")
    except Exception:
        pass
    alg_bytes = m["dev_px_per_step"] * wl["bytes_px"]
    achieved = alg_bytes / (m["raster_ms"] * 1e-3) / 1e9 if m["raster_ms"] > 0 else 0.0

    extra = {}
    if rank == 0 or world > 1:
        names = [] if args.no_extra else [n for n in WORKLOADS if n != wl_name]
        for n in names:
            try:
                steps = max(2, min(args.steps, 5)) if (n in ("c3_phong_4k", "c4_overdraw_8k", "c5_batch_512") or WORKLOADS[n]["scene"] == "overdraw") else args.steps
                x = measure_workload(n, steps, 3, torch, scenes, pfcu, stream, flush_buf)
                w2 = WORKLOADS[n]
                a2 = x["dev_px_per_step"] * w2["bytes_px"] / (x["raster_ms"] * 1e-3) / 1e9 if x["raster_ms"] > 0 else 0.0
                extra[n] = {"desc": w2["desc"], "gpix_per_s": x["dev_px_per_step"] / (x["dev_ms"] * 1e-3) / 1e9,
                            "mtri_per_s": x["dev_tris_per_step"] / (x["dev_ms"] * 1e-3) / 1e6, "ms_per_step": x["dev_ms"],
                            "e2e_gpix_per_s": x["px_per_step"] / (x["e2e_ms"] * 1e-3) / 1e9, "e2e_mtri_per_s": x["tris_per_step"] / (x["e2e_ms"] * 1e-3) / 1e6,
                            "e2e_ms_per_step": x["e2e_ms"], "shaded_px_per_step": x["dev_px_per_step"], "triangles_per_step": x["dev_tris_per_step"],
                            "roofline": {"bound": "hbm", "achieved": a2, "peak": peak, "unit": "GB/s", "frac": a2 / peak, "bytes_per_px": w2["bytes_px"],
                                         "raster_ms": x["raster_ms"], "frontend_ms": x["frontend_ms"]}}
                if "e2e_static" in x:
                    extra[n]["e2e_static_geometry"] = x["e2e_static"]
            except Exception as e:   # an extra must never take the headline down
                extra[n] = {"error": repr(e)}
        if rank == 0:
            clocks = sampler.stop()         # sampled every 100 ms over the headline's and the extras' timed regions
        if world > 1 and not args.no_extra:
            try:
                from pixelforge_b200.multigpu import tile_split_benchmark
                extra["tile_split_c4"] = tile_split_benchmark(torch, dist, scenes, pfcu, stream, WORKLOADS["c4_overdraw_8k"], rank, world, steps=3)
            except Exception as e:
                extra["tile_split_c4"] = {"error": repr(e)}

    # single-process multi-device mode of the library (PF_CUDA_DEVICES): rank 0 drives all N GPUs of the box through the
    # PUBLIC API in a fresh process while the other ranks wait - the C4 screen-tile split and the 4K scene, strong scaling
    # against the same tool on one device
    if world > 1 and not args.no_extra:
        import datetime
        barrier()
        # the other ranks wait on the rendezvous store, on the CPU: a NCCL barrier would keep a kernel spinning on their
        # GPUs, which the driver then time-slices with rank 0's work on the same devices
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            try:
                extra["single_process_multi_device"] = single_process_scaling(world)
            finally:
                store.set("pf_single_process_done", "1")
        else:
            store.wait(["pf_single_process_done"], datetime.timedelta(seconds=1200))
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    if clocks is None:
        clocks = sampler.stop()

    parity = {}

    def cpu_run(name, frames):
        """CPU baseline of one workload (the reference itself, in a subprocess, all host threads) and, from the very
        frames it rendered, the reference-equivalence gate of the product."""
        import tempfile
        env = dict(os.environ); env.pop("OMP_NUM_THREADS", None)      # torchrun pins it to 1; the baseline uses every core
        tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") else None
        with tempfile.TemporaryDirectory(dir=tmpdir) as d:
            dump = os.path.join(d, "ref.npz")
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--as-baseline", "--workload", name,
                                "--baseline-frames", str(frames), "--dump", dump], capture_output=True, text=True, timeout=900, env=env)
            j = json.loads(r.stdout.strip().splitlines()[-1])
            c = j.get("cpu_baseline") or {"unavailable": j.get("unavailable")}
            if "value" in c:
                c["ms_per_frame_of_sample"] = j["ms_per_step"]; c["mtri_per_s"] = j.get("mtri_per_s")
                parity[name] = parity_against_dump(scenes, name, dump)
        return c

    if not args.no_cpu_baseline and not args.no_extra:
        for n in extra:
            if n in WORKLOADS and "error" not in extra[n]:
                try:
                    extra[n]["cpu_baseline"] = cpu_run(n, 2)
                except Exception as e:
                    extra[n]["cpu_baseline"] = {"unavailable": repr(e)}

    cpu = None
    if not args.no_cpu_baseline:
        try:
            cpu = cpu_run(wl_name, 20)        # the headline's CPU sample: 20 frames, a few seconds of all host cores
            if "value" in (cpu or {}):
                cpu["ms_per_frame"] = cpu["ms_per_frame_of_sample"]
        except Exception as e:
            cpu = {"unavailable": repr(e)}
    gate_failed = [n for n, p in parity.items() if p["differing_px"] or p["differing_depth"]]

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32/f32", "data": "synthetic",
        "config": bench_config(wl_name, world),
        "measured_per_step": {"triangles": tris_all / world, "shaded_px": px_all / world},
        "mtri_per_s": tris_all / (dev_ms * 1e-3) / 1e6,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": m["h2d_bytes"], "d2h_bytes_per_step": m["d2h_bytes"],
                "ms_per_step": e2e_ms, "mtri_per_s": sum_over_ranks(m["tris_per_step"]) / (e2e_ms * 1e-3) / 1e6 if world == 1 else None,
                "static_geometry": m.get("e2e_static")},
        "gpu_launches": int(round(m["launches_per_step"] * args.steps)),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "k_raster_frag" if wl["scene"] in ("textured", "phong", "gears", "batch") else "k_raster",
                     "sm_issue_active_pct_ncu": issue_active, "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_shaded_px": wl["bytes_px"],
                     "kernel_ms": m["raster_ms"], "frontend_kernels_ms": m["frontend_ms"], "peak_source": peak_src},
        "cpu_baseline": cpu, "clocks": clocks,
        # the reference-equivalence gate: every workload's frame 0 at full size, product vs the frames the CPU arm rendered
        "parity": parity if parity else "not run (--no-cpu-baseline: no reference frames to compare with)",
        "extra": extra,
    }
    if gate_failed:
        # no number is reported for a build whose pixels differ from the reference's
        line["value"] = None; line["e2e"]["value"] = None
        line["error"] = "reference-equivalence gate failed for " + ", ".join(gate_failed)
    emit_json(line)
    if world > 1:
        dist.destroy_process_group()
    return 1 if gate_failed else 0


_REAL_STDOUT = None


def emit_json(obj):
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version
    banner, for one) was redirected to stderr at start-up."""
    data = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=list(WORKLOADS))
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--as-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--baseline-frames", type=int, default=3, help=argparse.SUPPRESS)
    ap.add_argument("--dump", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    return reference_main(args) if args.impl == "reference" else ours_main(args)


if __name__ == "__main__":
    sys.exit(main())
