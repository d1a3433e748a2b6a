"""pixelforge_b200 - B200-native drop-in for PixelForge's per-fragment triangle path.

The product is the C library ``pixelforge_b200/lib/libpixelforge.so`` (C99 front end exporting the
unchanged ``pixelforge.h`` API + hand-written sm_100a kernels behind the ``pfcu`` C-ABI).  This
Python package only holds thin ctypes bindings used by the tests and ``bench.py``; it never renders
anything itself and it raises loudly when the CUDA library is missing or no GPU is present.
"""
from .binding import (  # noqa: F401
    SceneLib, SceneCfg, SceneResult, load_product_scenes, load_pfcu, PfcuLib, LIB_DIR, REPO_ROOT, ProductUnavailable,
)
