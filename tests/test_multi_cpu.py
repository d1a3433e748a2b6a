"""N>1 host logic on CPU: two gloo ranks run the screen-tile split with the oracle build as the
"device", gather the owned tiles to rank 0 with pixelforge_b200.multigpu.gather_tiles (the code bench.py runs
over NCCL), and rank 0 checks the reassembled surface against a single-rank render, byte for byte."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _stream(w, h, n, seed):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_parity import random_stream
    return random_stream(np.random.default_rng(seed), w, h, n, 1 | 2 | 16, big=True, n_states=3)


def _prims(w, h):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_parity import random_prims
    return random_prims(np.random.default_rng(77), w, h, 200, 10)


def _worker(rank, world, port, w, h, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from checkers import load_oracle_pfcu
    from pixelforge_b200.multigpu import gather_tiles
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = load_oracle_pfcu(); lib.init(); L = lib.lib
    states, tris = _stream(w, h, 300, 5)
    s = L.pfcu_surface_create(w, h)
    L.pfcu_surface_fill(s, 1, 0xFF102030, 1, np.finfo(np.float32).max)
    L.pfcu_surface_set_tile_owner(s, rank, world)
    lib.check(L.pfcu_submit(s, states.ctypes.data, len(states), tris.ctypes.data, len(tris)))
    prims = _prims(w, h)
    lib.check(L.pfcu_submit_prims(s, prims.ctypes.data, len(prims)))
    moved = gather_tiles(torch, dist, lib, s, w, h, rank, world, with_depth=True, device="cpu")
    if rank == 0:
        c = np.zeros((h, w), np.uint32); d = np.zeros((h, w), np.float32)
        L.pfcu_surface_download(s, c.ctypes.data, d.ctypes.data, 0, h)
        np.savez(out_path, color=c, depth=d, moved=moved)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tile_split_gather_gloo(world, built_libraries, tmp_path):
    import torch.multiprocessing as mp
    from checkers import load_oracle_pfcu
    w, h = 333, 200
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), w, h, out), nprocs=world, join=True)
    got = np.load(out)
    lib = load_oracle_pfcu(); lib.init()
    states, tris = _stream(w, h, 300, 5)
    full_c, full_d = lib.render_stream(w, h, states, tris, color0=np.full((h, w), 0xFF102030, np.uint32), prims=_prims(w, h))
    assert np.array_equal(got["color"], full_c)
    assert np.array_equal(got["depth"].view(np.uint32), full_d.view(np.uint32))
    tiles = ((w + 63) // 64) * ((h + 63) // 64)
    assert int(got["moved"]) == (tiles - (tiles // world + (1 if tiles % world else 0))) * 64 * 64 * 8


def test_owned_tile_accounting(built_libraries):
    from checkers import load_oracle_pfcu
    from pixelforge_b200.multigpu import owned_tiles
    lib = load_oracle_pfcu(); L = lib.lib
    for (w, h) in ((7680, 4320), (800, 600), (65, 1)):
        s = L.pfcu_surface_create(w, h)
        for world in (1, 2, 4, 8):
            total = sum(owned_tiles(w, h, r, world) for r in range(world))
            assert total == ((w + 63) // 64) * ((h + 63) // 64)
            for r in range(world):
                assert L.pfcu_surface_owned_bytes(s, r, world, 1) == owned_tiles(w, h, r, world) * 64 * 64 * 8
        L.pfcu_surface_destroy(s)
