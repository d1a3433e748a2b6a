"""Multi-GPU plumbing for the screen-tile split (BASELINE.json config C4).

One process per GPU (torch.distributed).  Every rank receives the same ordered triangle batch and
rasterises only the 64x64 tiles it owns (owner(tile) = tile % world, enforced inside k_raster), so the
per-pixel result is byte-identical to a single-GPU render.  The only exchange step is the gather of
the finished tiles to the presenting rank: each rank packs its tiles into a contiguous staging
buffer (k_pack_tiles), the buffers are gathered with NCCL over NVLink (or gloo in CPU tests, where
the "device" is the oracle build), and the presenter unpacks them into its surface.

The peer-memory alternative (connect_present_peer + pfcu_surface_push_tiles): the presenter exports its surface over
CUDA IPC and every other rank stores its own tiles straight into it over NVLink with one copy kernel; no staging
buffers, no collective on the data path, only a barrier before the presenter reads.
"""
import numpy as np


def owned_tiles(width, height, rank, world, tile=64):
    nt = ((width + tile - 1) // tile) * ((height + tile - 1) // tile)
    return nt // world + (1 if (nt % world) > rank else 0)


def gather_tiles(torch, dist, pfcu, surface, width, height, rank, world, with_depth=False, device="cuda", dst=0):
    """Pack this rank's tiles, gather all ranks' tiles on `dst`, unpack there.  Returns bytes moved to dst."""
    L = pfcu.lib
    per_tile = 64 * 64 * 4 * (2 if with_depth else 1)
    max_tiles = owned_tiles(width, height, 0, world)            # rank 0 owns the most
    # no initialising kernel on torch's stream: the pack kernel runs on the surface's pfcu lane, which is a different
    # stream unless the surface happens to live on lane 0, and would race with it (the tail of the buffer of a rank that
    # owns fewer tiles is never read)
    staging = torch.empty(max_tiles * per_tile, dtype=torch.uint8, device=device)
    if device == "cuda":
        torch.cuda.current_stream().synchronize()
    pfcu.check(L.pfcu_surface_pack_tiles(surface, rank, world, int(with_depth), staging.data_ptr()), "pack_tiles")
    if device == "cuda":
        # pack ran on the pfcu stream; make the collective wait for it
        pfcu.check(L.pfcu_finish(), "finish")
    if world == 1:
        return 0
    if rank == dst:
        bufs = [torch.empty_like(staging) for _ in range(world)]
        dist.gather(staging, bufs, dst=dst)
        if device == "cuda":
            # the unpack kernels run on the surface's pfcu lane, which is not the stream the collective ran on
            torch.cuda.current_stream().synchronize()
        moved = 0
        for r in range(world):
            if r == dst:
                continue
            pfcu.check(L.pfcu_surface_unpack_tiles(surface, r, world, int(with_depth), bufs[r].data_ptr()), "unpack_tiles")
            moved += owned_tiles(width, height, r, world) * per_tile
        if device == "cuda":
            pfcu.check(L.pfcu_finish(), "finish")
        return moved
    dist.gather(staging, None, dst=dst)
    return 0


def connect_present_peer(dist, pfcu, surface, rank, world, with_depth=False, dst=0):
    """Present over peer memory (include/pfcu.h): rank `dst` exports its surface with CUDA IPC, every other rank maps
    it; pfcu_surface_push_tiles then stores a rank's tiles into the presenter's surface over NVLink."""
    import ctypes as C
    L = pfcu.lib
    hc, hd = (C.c_ubyte * 64)(), (C.c_ubyte * 64)()
    payload = [None]
    if rank == dst:
        pfcu.check(L.pfcu_surface_ipc_handles(surface, hc, hd if with_depth else None), "ipc_handles")
        payload = [(bytes(hc), bytes(hd) if with_depth else None)]
    dist.broadcast_object_list(payload, src=dst)
    if rank != dst:
        c, d = payload[0]
        hc = (C.c_ubyte * 64).from_buffer_copy(c)
        hd = (C.c_ubyte * 64).from_buffer_copy(d) if d else None
        pfcu.check(L.pfcu_surface_set_present_peer(surface, hc, hd), "set_present_peer")


def tile_split_benchmark(torch, dist, scenes, pfcu, stream, wl, rank, world, steps=3):
    """Strong-scaling run of one big surface split by screen tiles across `world` GPUs."""
    from .binding import Counters
    L = pfcu.lib
    with scenes.open(wl["scene"], wl["w"], wl["h"], variant=wl["variant"], size=wl["size"], explicit_sync=1) as sc:
        L.pfcu_set_stream(stream.cuda_stream)
        L.pfxCaptureBegin()
        sc.frame(0)
        states, tris = pfcu.capture_end()
        surf = L.pfxGetSurfaceHandle()
        L.pfcu_surface_set_tile_owner(surf, rank, world)
        b = L.pfcu_batch_upload(states.ctypes.data, len(states), tris.ctypes.data, len(tris))
        times, gather_ms = [], []
        L.pfxResetCounters()
        for i in range(steps + 1):
            torch.cuda.synchronize(); dist.barrier()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            with torch.cuda.stream(stream):
                e0.record(stream)
                L.pfcu_fence()
                L.pfcu_surface_clear_ref(surf, 1, 0xFF000000, 1, 3.4028234663852886e38)
                L.pfcu_batch_submit(surf, b)
                L.pfcu_fence()
                e1.record(stream)
                gather_tiles(torch, dist, pfcu, surf, wl["w"], wl["h"], rank, world)
                e2.record(stream)
            torch.cuda.synchronize()
            if i > 0:
                times.append(e0.elapsed_time(e1)); gather_ms.append(e1.elapsed_time(e2))
        k = Counters(); L.pfcu_get_counters(k)
        # the same frame presented over peer memory: after rendering, every other rank stores its tiles straight
        # into rank 0's surface (one copy kernel over NVLink); all that is left is a barrier
        fused, fused_ok = [], None
        # verification on the presenting rank: the gathered image and, below, the peer-presented image against a full
        # single-GPU render of the same frame (pfClear never clears pixels 0..7, Q12: with additive layers they depend on
        # the number of frames drawn so far and are left out)
        def grab():
            c = np.zeros((wl["h"], wl["w"]), np.uint32)
            pfcu.check(L.pfcu_finish(), "finish")
            pfcu.check(L.pfcu_surface_download(surf, c.ctypes.data, None, 0, wl["h"]), "download")
            c.reshape(-1)[:8] = 0
            return c
        truth = gathered_diff = None
        if rank == 0:
            gathered = grab()
            L.pfcu_surface_set_tile_owner(surf, 0, 1)
            L.pfcu_surface_clear_ref(surf, 1, 0xFF000000, 1, 3.4028234663852886e38)
            L.pfcu_batch_submit(surf, b)
            truth = grab()
            L.pfcu_surface_set_tile_owner(surf, rank, world)
            gathered_diff = int((gathered != truth).sum())
            del gathered
        try:
            connect_present_peer(dist, pfcu, surf, rank, world)
            ok = 1.0
        except Exception as ex:     # no peer access between these devices
            ok, fused_ok = 0.0, repr(ex)
        okt = torch.tensor([ok], dtype=torch.float64, device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)          # every rank takes the same branch: collectives follow
        if float(okt[0]) > 0.5:
            for i in range(steps + 1):
                torch.cuda.synchronize(); dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    e0.record(stream)
                    L.pfcu_fence()
                    L.pfcu_surface_clear_ref(surf, 1, 0xFF000000, 1, 3.4028234663852886e38)
                # the presenter's clear must have finished before any peer stores a tile into its surface
                torch.cuda.synchronize(); dist.barrier()
                with torch.cuda.stream(stream):
                    L.pfcu_batch_submit(surf, b)
                    if rank != 0:
                        L.pfcu_surface_push_tiles(surf, rank, world, 0)
                    L.pfcu_fence()
                    e1.record(stream)
                torch.cuda.synchronize(); dist.barrier()
                if i > 0:
                    fused.append(e0.elapsed_time(e1))
            if rank == 0:
                fused_ok = int((grab() != truth).sum())
        elif fused_ok is None:
            fused_ok = "another rank could not map the presenter's surface"
        L.pfcu_surface_clear_present(surf)
        tf = torch.tensor([sum(fused) / len(fused) if fused else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        t = torch.tensor([sum(times) / len(times), sum(gather_ms) / len(gather_ms)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        px = torch.tensor([k.pixels_shaded / (steps + 1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(px, op=dist.ReduceOp.SUM)
        L.pfcu_surface_set_tile_owner(surf, 0, 1)
        L.pfcu_batch_destroy(b)
        sc.finish()
    render_ms, gat_ms, fused_ms = float(t[0]), float(t[1]), float(tf[0])
    out = {"desc": wl["desc"] + f", screen-tile split over {world} GPUs; present to rank 0 by NCCL gather and by peer-memory stores (render + push, clear barrier included)", "scaling": "strong",
           "render_ms": render_ms, "gather_ms": gat_ms, "shaded_px": float(px[0]),
           "gpix_per_s_render": float(px[0]) / (render_ms * 1e-3) / 1e9,
           "gpix_per_s_with_gather": float(px[0]) / ((render_ms + gat_ms) * 1e-3) / 1e9,
           "gathered_pixels_differing_from_single_gpu": gathered_diff}
    if fused_ms > 0:
        out.update(peer_present_ms=fused_ms, gpix_per_s_peer_present=float(px[0]) / (fused_ms * 1e-3) / 1e9,
                   peer_present_pixels_differing_from_single_gpu=fused_ok)
    else:
        out.update(peer_present_unavailable=str(fused_ok))
    return out
