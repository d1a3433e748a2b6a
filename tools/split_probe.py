"""One rank's share of a tile-split surface on ONE GPU: device time of the C4 / 4K scene for (rank 0, world N), to see how
the per-rank time scales without any multi-GPU effects.  PF_CUDA_SLICES forces the slice height."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from pixelforge_b200 import load_product_scenes, load_pfcu
scenes = load_product_scenes(); pfcu = load_pfcu("product")
stream = torch.cuda.Stream(); flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
name = sys.argv[1] if len(sys.argv) > 1 else "c4_overdraw_8k"
for world in (1, 2, 4, 8):
    m = bench.measure_workload(name, 5, 3, torch, scenes, pfcu, stream, flush, want_e2e=False, tile_owner=(0, world) if world > 1 else None)
    print(name, "world", world, "dev_ms %.3f raster_ms %.3f front_ms %.3f  ideal %.3f" % (m["dev_ms"], m["raster_ms"], m["frontend_ms"], 0), flush=True)
