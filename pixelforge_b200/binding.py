"""ctypes bindings: the scene runner and the pfcu C-ABI (include/pfcu.h) of the PRODUCT.

SceneLib / PfcuLib are generic typed wrappers over a shared library path; this package only ever points them at
pixelforge_b200/lib/ (libpfscenes_cuda.so -> libpixelforge.so, CUDA sm_100a).  The checkers (the scalar C oracle and
the unmodified reference under oracle/) are loaded by tests/checkers.py, never from here.
"""
import ctypes as C
import os

import numpy as np

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(REPO_ROOT, "pixelforge_b200", "lib")


class ProductUnavailable(RuntimeError):
    """libpixelforge.so is missing or no CUDA device is usable - there is no CPU fallback."""


class SceneCfg(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("frames", C.c_int), ("warmup", C.c_int),
                ("variant", C.c_int), ("size", C.c_int), ("seed", C.c_int), ("explicit_sync", C.c_int),
                ("first_frame", C.c_int)]


class SceneResult(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_min", C.c_double), ("ms_median", C.c_double),
                ("triangles_submitted", C.c_ulonglong), ("triangles_rasterised", C.c_ulonglong),
                ("pixels_shaded", C.c_ulonglong), ("pixels_depth_failed", C.c_ulonglong),
                ("kernel_launches", C.c_ulonglong), ("api_triangles", C.c_ulonglong)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SceneLib:
    """One build of the scene runner."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        self.lib.pfscene_render.restype = C.c_int
        self.lib.pfscene_render.argtypes = [C.c_char_p, C.POINTER(SceneCfg), C.c_void_p, C.c_void_p, C.POINTER(SceneResult)]
        self.lib.pfscene_backend.restype = C.c_char_p
        self.lib.pfscene_open.restype = C.c_void_p
        self.lib.pfscene_open.argtypes = [C.c_char_p, C.POINTER(SceneCfg)]
        self.lib.pfscene_frame.restype = None
        self.lib.pfscene_frame.argtypes = [C.c_void_p, C.c_int]
        self.lib.pfscene_finish.restype = None
        self.lib.pfscene_finish.argtypes = [C.c_void_p]
        self.lib.pfscene_make_current.restype = None
        self.lib.pfscene_make_current.argtypes = [C.c_void_p, C.c_int]
        self.lib.pfscene_read.restype = None
        self.lib.pfscene_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.pfscene_close.restype = None
        self.lib.pfscene_close.argtypes = [C.c_void_p]

    @property
    def backend(self):
        return self.lib.pfscene_backend().decode()

    def render(self, name, width, height, frames=1, warmup=0, variant=0, size=0, seed=1,
               explicit_sync=1, first_frame=0, want_depth=True):
        cfg = SceneCfg(width, height, frames, warmup, variant, size, seed, explicit_sync, first_frame)
        color = np.zeros((height, width), dtype=np.uint32)
        depth = np.zeros((height, width), dtype=np.float32) if want_depth else None
        res = SceneResult()
        rc = self.lib.pfscene_render(name.encode(), C.byref(cfg), color.ctypes.data,
                                     depth.ctypes.data if want_depth else None, C.byref(res))
        if rc != 0:
            raise RuntimeError(f"pfscene_render({name}) failed with {rc} on backend {self.path}")
        return color, depth, res


    def open(self, name, width, height, variant=0, size=0, seed=1, explicit_sync=1):
        """Keep a scene (context, textures, meshes) alive across frames; see Scene."""
        return Scene(self, name, SceneCfg(width, height, 1, 0, variant, size, seed, explicit_sync, 0))


class Scene:
    def __init__(self, lib, name, cfg):
        self.lib, self.cfg, self.name = lib, cfg, name
        self.handle = lib.lib.pfscene_open(name.encode(), C.byref(cfg))
        if not self.handle:
            raise RuntimeError(f"pfscene_open({name}) failed on {lib.path}")

    def frame(self, index=0):
        self.lib.lib.pfscene_frame(self.handle, index)

    def finish(self):
        self.lib.lib.pfscene_finish(self.handle)

    def make_current(self, index=0):
        self.lib.lib.pfscene_make_current(self.handle, index)

    def read_context(self, index):
        """Colour buffer of context `index` of a "batch" scene, exactly as the application sees it."""
        color = np.zeros((self.cfg.height, self.cfg.width), np.uint32)
        self.lib.lib.pfscene_read_context.restype = C.c_int
        self.lib.lib.pfscene_read_context.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        if not self.lib.lib.pfscene_read_context(self.handle, index, color.ctypes.data):
            raise IndexError(index)
        return color

    def read_index(self, index, want_depth=True):
        """(colour u32[h,w], depth f32[h,w] or None) of context `index`, read through the public API."""
        color = np.zeros((self.cfg.height, self.cfg.width), np.uint32)
        depth = np.zeros((self.cfg.height, self.cfg.width), np.float32) if want_depth else None
        f = self.lib.lib.pfscene_read_index
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        if not f(self.handle, index, color.ctypes.data, depth.ctypes.data if want_depth else None):
            raise IndexError(index)
        return color, depth

    def read(self, want_depth=False):
        color = np.zeros((self.cfg.height, self.cfg.width), dtype=np.uint32)
        depth = np.zeros((self.cfg.height, self.cfg.width), dtype=np.float32) if want_depth else None
        self.lib.lib.pfscene_read(self.handle, color.ctypes.data, depth.ctypes.data if want_depth else None)
        return color, depth

    def close(self):
        if self.handle:
            self.lib.lib.pfscene_close(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def load_product_scenes():
    path = os.path.join(LIB_DIR, "libpfscenes_cuda.so")
    if not os.path.exists(path) or not os.path.exists(os.path.join(LIB_DIR, "libpixelforge.so")):
        raise ProductUnavailable(f"{path} not built - run `make lib scenes` (or __graft_entry__.build())")
    return SceneLib(path)


# ---- pfcu C-ABI ------------------------------------------------------------------------------------

VERTEX_DTYPE = np.dtype([("sx", "<f4"), ("sy", "<f4"), ("zinv", "<f4"), ("u", "<f4"), ("v", "<f4"),
                         ("px", "<f4"), ("py", "<f4"), ("pz", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                         ("rgba", "<u4")])
TRIANGLE_DTYPE = np.dtype([("v", VERTEX_DTYPE, (3,)), ("state", "<u4"), ("face", "u1"), ("is3d", "u1"), ("pad", "<u2")])
PRIM_DTYPE = np.dtype([("x1", "<f4"), ("y1", "<f4"), ("x2", "<f4"), ("y2", "<f4"), ("z1", "<f4"), ("z2", "<f4"), ("c1", "<u4"), ("c2", "<u4"),
                       ("size", "<f4"), ("kind", "u1"), ("flags", "u1"), ("blend_mode", "u1"), ("depth_func", "u1")])      # pfcu_prim, 40 B
LIGHT_DTYPE = np.dtype([("position", "<f4", (3,)), ("direction", "<f4", (3,)), ("inner_cutoff", "<f4"),
                        ("outer_cutoff", "<f4"), ("att_constant", "<f4"), ("att_linear", "<f4"),
                        ("att_quadratic", "<f4"), ("ambient", "<u4"), ("diffuse", "<u4"), ("specular", "<u4")])
MATERIAL_DTYPE = np.dtype([("ambient", "<u4"), ("diffuse", "<u4"), ("specular", "<u4"), ("emission", "<u4"),
                           ("shininess", "<f4")])
STATE_DTYPE = np.dtype([("flags", "<u4"), ("blend_mode", "u1"), ("depth_func", "u1"), ("tex_filter", "u1"),
                        ("tex_wrap", "u1"), ("vp_min", "<i4", (2,)), ("vp_max", "<i4", (2,)), ("texture", "<u8"),
                        ("n_lights", "<u4"), ("lights", LIGHT_DTYPE, (8,)), ("material", MATERIAL_DTYPE, (2,)),
                        ("view_pos", "<f4", (3,)), ("pad", "<u4")], align=True)

ST_BLEND, ST_DEPTH_TEST, ST_TEXTURE, ST_PHONG, ST_SMOOTH = 1, 2, 4, 8, 16
TEX_RGBA8, TEX_BGRA8, TEX_RGB8, TEX_BGR8 = 0, 1, 2, 3


class Profile(C.Structure):
    _fields_ = [("raster_ms", C.c_double), ("frontend_ms", C.c_double), ("raster_launches", C.c_uint64)]


class Counters(C.Structure):
    _fields_ = [("triangles_submitted", C.c_uint64), ("triangles_rasterised", C.c_uint64),
                ("pixels_shaded", C.c_uint64), ("pixels_depth_failed", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("bytes_h2d", C.c_uint64), ("bytes_d2h", C.c_uint64)]


PFCU_SYMBOLS = [
    "pfcu_init", "pfcu_shutdown", "pfcu_last_error", "pfcu_backend_name", "pfcu_set_stream", "pfcu_get_stream",
    "pfcu_host_alloc", "pfcu_host_free", "pfcu_host_wait", "pfcu_host_register", "pfcu_host_unregister", "pfcu_set_approx_tables",
    "pfcu_surface_create", "pfcu_surface_wrap", "pfcu_surface_destroy", "pfcu_surface_width", "pfcu_surface_height",
    "pfcu_surface_color_ptr", "pfcu_surface_depth_ptr", "pfcu_surface_upload", "pfcu_surface_download",
    "pfcu_surface_fill", "pfcu_surface_clear_ref", "pfcu_surface_set_tile_owner", "pfcu_surface_owned_bytes",
    "pfcu_surface_pack_tiles", "pfcu_surface_unpack_tiles",
    "pfcu_texture_create", "pfcu_texture_from_surface", "pfcu_texture_update", "pfcu_texture_destroy",
    "pfcu_submit", "pfcu_batch_upload", "pfcu_batch_submit", "pfcu_batch_destroy",
    "pfcu_fence", "pfcu_finish", "pfcu_get_counters", "pfcu_reset_counters", "pfcu_profile_enable", "pfcu_profile_read",
    "pfcu_set_raster_path", "pfcu_submit_raw", "pfcu_submit_prims", "pfcu_surface_download_async", "pfcu_surface_wait", "pfcu_surface_ipc_handles", "pfcu_surface_set_present_peer",
    "pfcu_surface_set_present_surface", "pfcu_surface_clear_present", "pfcu_surface_push_tiles",
    "pfcu_surface_create_format", "pfcu_surface_format",
    "pfcu_list_create", "pfcu_list_destroy", "pfcu_list_size", "pfcu_list_job_supported", "pfcu_submit_list_jobs",
    "pfcu_host_set_static", "pfcu_host_modified", "pfcu_host_is_static", "pfcu_surface_rect", "pfcu_surface_fog", "pfcu_surface_draw_pixels", "pfcu_surface_read_pixels",
]

PFX_SYMBOLS = ["pfxSetSyncMode", "pfxFlush", "pfxFinish", "pfxGetCounters", "pfxResetCounters", "pfxSetTileOwner",
               "pfxGetDeviceColor", "pfxGetDeviceDepth", "pfxReadDepth", "pfxCaptureBegin", "pfxCaptureEnd",
               "pfxGetSurfaceHandle", "pfxBackendName", "pfxEnableDeviceVertexStage", "pfxSpecularTableCheck", "pfxHostAlloc", "pfxHostFree", "pfxTextureDirty", "pfxEnableQueuedReadback", "pfxHostStatic", "pfxHostModified", "pfxFogTableCheck"]


class Fog(C.Structure):          # pfcu_fog
    _fields_ = [("start", C.c_float), ("end", C.c_float), ("inv_len", C.c_float), ("density", C.c_float), ("rgba", C.c_uint32),
                ("mode", C.c_uint32), ("thresholds", C.c_void_p), ("n_thresholds", C.c_uint32)]


class Pixels(C.Structure):       # pfcu_pixels
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_int), ("xs", C.c_int32), ("ys", C.c_int32),
                ("xmin", C.c_int32), ("ymin", C.c_int32), ("xmax", C.c_int32), ("ymax", C.c_int32), ("inv_xlen", C.c_float), ("inv_ylen", C.c_float),
                ("z", C.c_float), ("flags", C.c_uint32), ("blend_mode", C.c_uint8), ("depth_func", C.c_uint8), ("pad", C.c_uint16)]


def pix_code(pf_format, pf_type):
    """PFCU_PIX(PFpixelformat, PFdatatype) of include/pfcu.h."""
    return pf_format * 16 + pf_type


class PfcuLib:
    """Typed access to a library exporting the pfcu C-ABI (the product or the oracle build)."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        L = self.lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        vp, u32, sz = C.c_void_p, C.c_uint32, C.c_size_t
        sig = {
            "pfcu_init": (C.c_int, [C.c_int]), "pfcu_shutdown": (None, []),
            "pfcu_last_error": (C.c_char_p, []), "pfcu_backend_name": (C.c_char_p, []),
            "pfcu_set_stream": (C.c_int, [vp]), "pfcu_get_stream": (vp, []),
            "pfcu_host_alloc": (vp, [sz]), "pfcu_host_free": (None, [vp]), "pfcu_host_wait": (C.c_int, [vp]),
            "pfcu_set_approx_tables": (C.c_int, [vp, C.c_int, vp, C.c_int]),
            "pfcu_surface_create": (vp, [u32, u32]), "pfcu_surface_create_format": (vp, [u32, u32, C.c_int]),
            "pfcu_surface_format": (C.c_int, [vp]), "pfcu_surface_wrap": (vp, [vp, vp, u32, u32]),
            "pfcu_surface_destroy": (None, [vp]), "pfcu_surface_width": (u32, [vp]), "pfcu_surface_height": (u32, [vp]),
            "pfcu_surface_color_ptr": (vp, [vp]), "pfcu_surface_depth_ptr": (vp, [vp]),
            "pfcu_surface_upload": (C.c_int, [vp, vp, vp, u32, u32]), "pfcu_surface_download": (C.c_int, [vp, vp, vp, u32, u32]),
            "pfcu_surface_fill": (C.c_int, [vp, C.c_int, u32, C.c_int, C.c_float]),
            "pfcu_surface_clear_ref": (C.c_int, [vp, C.c_int, u32, C.c_int, C.c_float]),
            "pfcu_surface_set_tile_owner": (C.c_int, [vp, u32, u32]),
            "pfcu_surface_owned_bytes": (sz, [vp, u32, u32, C.c_int]),
            "pfcu_surface_pack_tiles": (C.c_int, [vp, u32, u32, C.c_int, vp]),
            "pfcu_surface_unpack_tiles": (C.c_int, [vp, u32, u32, C.c_int, vp]),
            "pfcu_texture_create": (vp, [vp, u32, u32, C.c_int]), "pfcu_texture_from_surface": (vp, [vp]),
            "pfcu_texture_update": (C.c_int, [vp, vp]), "pfcu_texture_destroy": (None, [vp]),
            "pfcu_surface_ipc_handles": (C.c_int, [vp, vp, vp]), "pfcu_surface_set_present_peer": (C.c_int, [vp, vp, vp]),
            "pfcu_surface_set_present_surface": (C.c_int, [vp, vp]), "pfcu_surface_clear_present": (C.c_int, [vp]),
            "pfcu_surface_push_tiles": (C.c_int, [vp, u32, u32, C.c_int]),
            "pfcu_submit": (C.c_int, [vp, vp, u32, vp, u32]), "pfcu_submit_prims": (C.c_int, [vp, vp, u32]), "pfcu_batch_upload": (vp, [vp, u32, vp, u32]),
            "pfcu_batch_submit": (C.c_int, [vp, vp]), "pfcu_batch_destroy": (None, [vp]),
            "pfcu_surface_rect": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, u32]),
            "pfcu_surface_fog": (C.c_int, [vp, C.POINTER(Fog)]), "pfcu_surface_draw_pixels": (C.c_int, [vp, C.POINTER(Pixels)]),
            "pfcu_surface_read_pixels": (C.c_int, [vp, u32, u32, u32, u32, u32, C.c_int, vp]),
            "pfcu_fence": (C.c_int, []), "pfcu_finish": (C.c_int, []), "pfcu_get_counters": (C.c_int, [C.POINTER(Counters)]),
            "pfcu_reset_counters": (None, []),
            "pfcu_profile_enable": (None, [C.c_int]), "pfcu_set_raster_path": (None, [C.c_int]), "pfcu_profile_read": (C.c_int, [C.POINTER(Profile)]),
            # pfx extensions of the front end (operate on the calling thread's current context)
            "pfxFlush": (None, []), "pfxFinish": (None, []), "pfxSetTileOwner": (None, [u32, u32]),
            "pfxGetDeviceColor": (vp, []), "pfxGetDeviceDepth": (vp, []), "pfxGetSurfaceHandle": (vp, []),
            "pfxCaptureBegin": (None, []),
            "pfxCaptureEnd": (None, [C.POINTER(vp), C.POINTER(u32), C.POINTER(vp), C.POINTER(u32)]),
            "pfxResetCounters": (None, []), "pfxEnableQueuedReadback": (None, [C.c_ubyte]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        # host-side helpers exported by the front end (pf_x86approx.c)
        if hasattr(L, "pfh_harvest_tables"):
            L.pfh_harvest_tables.restype = C.c_int
            L.pfh_harvest_tables.argtypes = [C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_int),
                                             C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_int)]
            for n in ("pfh_rcp_from_table", "pfh_rsqrt_from_table"):
                f = getattr(L, n)
                f.restype, f.argtypes = C.c_float, [C.POINTER(C.c_uint32), C.c_int, C.c_float]
            for n in ("pfh_hw_rcp", "pfh_hw_rsqrt"):
                f = getattr(L, n)
                f.restype, f.argtypes = C.c_float, [C.c_float]

    @property
    def backend(self):
        return self.lib.pfcu_backend_name().decode()

    def error(self):
        return self.lib.pfcu_last_error().decode()

    def check(self, rc, what="pfcu call"):
        if rc != 0:
            raise RuntimeError(f"{what} failed with code {rc}: {self.error()}")

    def harvest_tables(self):
        rcp, rsq = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
        rb, sb = C.c_int(), C.c_int()
        if not self.lib.pfh_harvest_tables(C.byref(rcp), C.byref(rb), C.byref(rsq), C.byref(sb)):
            raise RuntimeError("could not characterise RCPPS/RSQRTPS on this CPU")
        return rcp, rb.value, rsq, sb.value

    def init(self, device=-1):
        rc = self.lib.pfcu_init(device)
        if rc != 0:
            raise ProductUnavailable(f"pfcu_init failed ({rc}): {self.error()}")
        rcp, rb, rsq, sb = self.harvest_tables()
        self.check(self.lib.pfcu_set_approx_tables(rcp, rb, rsq, sb), "pfcu_set_approx_tables")
        self._tables = (rcp, rb, rsq, sb)

    def capture_end(self):
        """-> (states ndarray[STATE_DTYPE], tris ndarray[TRIANGLE_DTYPE]) copied out of the front end."""
        ps, pt, ns, nt = C.c_void_p(), C.c_void_p(), C.c_uint32(), C.c_uint32()
        self.lib.pfxCaptureEnd(C.byref(ps), C.byref(ns), C.byref(pt), C.byref(nt))
        states = np.frombuffer((C.c_char * (ns.value * STATE_DTYPE.itemsize)).from_address(ps.value), dtype=STATE_DTYPE).copy() if ns.value else np.zeros(0, STATE_DTYPE)
        tris = np.frombuffer((C.c_char * (nt.value * TRIANGLE_DTYPE.itemsize)).from_address(pt.value), dtype=TRIANGLE_DTYPE).copy() if nt.value else np.zeros(0, TRIANGLE_DTYPE)
        return states, tris

    def render_stream(self, width, height, states, tris, color0=None, depth0=None, clear=None, tile_owner=None, prims=None, fmt=TEX_RGBA8):
        """Rasterise a triangle stream (then, optionally, a stream of points / lines) into a fresh surface of colour
        layout `fmt`; returns (color, depth f32[h,w]) with color u32[h,w] for the 4-byte layouts and u8[h,w,3] for
        RGB8 / BGR8 (the caller's layout, as pfcu_surface_download delivers it)."""
        L = self.lib
        s = L.pfcu_surface_create_format(width, height, fmt)
        if not s:
            raise RuntimeError("pfcu_surface_create failed: " + self.error())
        cshape, cdtype = ((height, width), np.uint32) if fmt in (TEX_RGBA8, TEX_BGRA8) else ((height, width, 3), np.uint8)
        try:
            color = np.ascontiguousarray(color0 if color0 is not None else np.zeros(cshape, cdtype))
            assert color.shape == cshape and color.dtype == cdtype
            depth = np.ascontiguousarray(depth0 if depth0 is not None else np.full((height, width), np.finfo(np.float32).max, np.float32))
            self.check(L.pfcu_surface_upload(s, color.ctypes.data, depth.ctypes.data, 0, height), "upload")
            if clear is not None:
                self.check(L.pfcu_surface_clear_ref(s, 1, clear[0], 1, clear[1]), "clear")
            if tile_owner is not None:
                self.check(L.pfcu_surface_set_tile_owner(s, tile_owner[0], tile_owner[1]), "tile owner")
            states = np.ascontiguousarray(states)
            tris = np.ascontiguousarray(tris)
            assert states.dtype == STATE_DTYPE and tris.dtype == TRIANGLE_DTYPE
            self.check(L.pfcu_submit(s, states.ctypes.data, len(states), tris.ctypes.data, len(tris)), "pfcu_submit")
            if prims is not None and len(prims):
                prims = np.ascontiguousarray(prims)
                assert prims.dtype == PRIM_DTYPE and prims.dtype.itemsize == 40
                self.check(L.pfcu_submit_prims(s, prims.ctypes.data, len(prims)), "pfcu_submit_prims")
            self.check(L.pfcu_finish(), "finish")
            out_c = np.zeros(cshape, cdtype)
            out_d = np.zeros((height, width), np.float32)
            self.check(L.pfcu_surface_download(s, out_c.ctypes.data, out_d.ctypes.data, 0, height), "download")
            return out_c, out_d
        finally:
            L.pfcu_surface_destroy(s)


def load_pfcu(which="product"):
    if which != "product":
        raise ValueError(f"{which!r}: this package only loads the product; the checkers live in tests/checkers.py")
    path = os.path.join(LIB_DIR, "libpixelforge.so")
    if not os.path.exists(path):
        raise ProductUnavailable(f"{path} not built - run `make lib` (or __graft_entry__.build())")
    return PfcuLib(path)
