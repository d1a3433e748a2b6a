import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
os.environ["PF_CUDA_DEVICE"] = str(local)
import torch, torch.distributed as dist
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from pixelforge_b200 import load_product_scenes, load_pfcu
from pixelforge_b200.multigpu import gather_tiles, connect_present_peer
scenes = load_product_scenes(); pfcu = load_pfcu("product"); L = pfcu.lib
W, H = int(sys.argv[1]), int(sys.argv[2])
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)       # lane 0 and NCCL on one stream, as bench.py does
with scenes.open("overdraw", W, H, variant=0, size=8, explicit_sync=1) as sc:
    L.pfcu_set_stream(stream.cuda_stream)
    L.pfxCaptureBegin(); sc.frame(0); states, tris = pfcu.capture_end()
    surf = L.pfxGetSurfaceHandle()
    L.pfcu_surface_set_tile_owner(surf, rank, world)
    b = L.pfcu_batch_upload(states.ctypes.data, len(states), tris.ctypes.data, len(tris))
    L.pfcu_surface_clear_ref(surf, 1, 0xFF000000, 1, 3.4028234663852886e38)
    L.pfcu_batch_submit(surf, b)
    gather_tiles(torch, dist, pfcu, surf, W, H, rank, world)
    ref = np.zeros((H, W), np.uint32)
    if rank == 0: pfcu.check(L.pfcu_surface_download(surf, ref.ctypes.data, None, 0, H))
    connect_present_peer(dist, pfcu, surf, rank, world)
    for i in range(4):
        torch.cuda.synchronize(); dist.barrier()
        with torch.cuda.stream(stream):
            L.pfcu_fence()
            L.pfcu_surface_clear_ref(surf, 1, 0xFF000000, 1, 3.4028234663852886e38)
        torch.cuda.synchronize(); dist.barrier()
        with torch.cuda.stream(stream):
            L.pfcu_batch_submit(surf, b)
            if rank != 0: L.pfcu_surface_push_tiles(surf, rank, world, 0)
            L.pfcu_fence()
        torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        c = np.zeros((H, W), np.uint32); pfcu.check(L.pfcu_surface_download(surf, c.ctypes.data, None, 0, H))
        c.reshape(-1)[:8] = 0; ref.reshape(-1)[:8] = 0      # never cleared (Q12), so they differ by the frame count
        d = c != ref
        print("differing pixels", int(d.sum()), "of", W * H)
        if d.any():
            ys, xs = np.nonzero(d)
            tiles = sorted(set(zip((ys // 64).tolist(), (xs // 64).tolist())))
            print("tiles with differences (ty,tx):", tiles[:20], "count", len(tiles))
            print("owner of those tiles:", sorted(set(((tx + ty * ((W + 63) // 64)) % world) for ty, tx in tiles)))
            print("sample", hex(c[ys[0], xs[0]]), hex(ref[ys[0], xs[0]]), ys[0], xs[0])
    L.pfcu_surface_clear_present(surf)
    L.pfcu_surface_set_tile_owner(surf, 0, 1)
    L.pfcu_batch_destroy(b); sc.finish()
dist.barrier(); dist.destroy_process_group()
