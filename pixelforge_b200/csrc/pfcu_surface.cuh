/* pfcu_surface.cuh - kernels: surface fill / clear / tile packing.
 * Part of the single translation unit pfcu.cu (included there, in order; not a stand-alone header). */

/* ------------------------------------------------------------------------------------------------ */
/* kernels: surface utilities                                                                       */
/* ------------------------------------------------------------------------------------------------ */

__global__ void k_fill(uint32_t *color, float *depth, size_t first, size_t n, int do_color, uint32_t rgba, int do_depth, float z)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < first + n; i += stride) {
        if (do_color) color[i] = rgba;
        if (do_depth) depth[i] = z;
    }
}

/* vectorised body of a fill: [first4*4, (first4+n4)*4) */
__global__ void k_fill4(uint4 *color, float4 *depth, size_t first4, size_t n4, int do_color, uint32_t rgba, int do_depth, float z)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint4 cv = make_uint4(rgba, rgba, rgba, rgba);
    const float4 dv = make_float4(z, z, z, z);
    for (size_t i = first4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < first4 + n4; i += stride) {
        if (do_color) color[i] = cv;
        if (do_depth) depth[i] = dv;
    }
}

/* tail of the reference's pfClear: pixels [aligned, size) copy pixel 0 (context.c:710-713) */
__global__ void k_clear_tail(uint32_t *color, float *depth, unsigned aligned, unsigned size, int do_color, int do_depth)
{
    const unsigned i = aligned + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < size) { if (do_color) color[i] = color[0]; if (do_depth) depth[i] = depth[0]; }
}

/* pfClear of a tile-split surface: a rank clears only the tiles it rasterises (the others are never read on this device).
 * Same rule as the whole-surface form for surfaces whose size is a multiple of 8 pixels: pixels 0..7 keep their value. */
__global__ void __launch_bounds__(256)
k_clear_tiles(uint32_t *__restrict__ color, float *__restrict__ depth, int W, int H, int tilesX, unsigned nTiles, unsigned rank, unsigned world,
              int do_color, uint32_t rgba, int do_depth, float z)
{
    const unsigned tile = rank + blockIdx.x * world;
    if (tile >= nTiles) return;
    const int X0 = (tile % tilesX) * TILE, Y0 = (tile / tilesX) * TILE;
    if ((W & 3) == 0 && X0 + TILE <= W && tile != 0) {
        const uint4 cv = make_uint4(rgba, rgba, rgba, rgba); const float4 dv = make_float4(z, z, z, z);
        for (int k = threadIdx.x; k < TILE * 16; k += 256) {
            const int r = k >> 4, c4 = (k & 15) << 2;
            if (Y0 + r >= H) break;
            const size_t gi = (size_t)(Y0 + r) * W + X0 + c4;
            if (do_color) *reinterpret_cast<uint4 *>(color + gi) = cv;
            if (do_depth) *reinterpret_cast<float4 *>(depth + gi) = dv;
        }
    } else {
        for (int k = threadIdx.x; k < TILE_PIX; k += 256) {
            const int x = X0 + (k & (TILE - 1)), y = Y0 + (k >> 6);
            if (x >= W || y >= H) continue;
            const size_t gi = (size_t)y * W + x;
            if (gi < 8) continue;
            if (do_color) color[gi] = rgba;
            if (do_depth) depth[gi] = z;
        }
    }
}

/* Layout conversion between the caller's colour format (staging, 3 or 4 bytes per pixel) and the canonical RGBA8 the
 * device works in: to_native = 0: staging -> canonical (upload), 1: canonical -> staging (download).  The scalar
 * getters / setters of the reference (pixel.h:233-360,576-710): BGRA8 swaps R and B, RGB8 / BGR8 drop alpha and read it
 * back as 255. */
__global__ void __launch_bounds__(256)
k_surface_convert(uint32_t *__restrict__ canon, unsigned char *__restrict__ native, size_t n, int fmt, int to_native)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (fmt == PFCU_TEX_BGRA8) {
            uint32_t *nat = reinterpret_cast<uint32_t *>(native);
            if (to_native) nat[i] = __byte_perm(canon[i], 0, 0x3012); else canon[i] = __byte_perm(nat[i], 0, 0x3012);
        } else {
            unsigned char *q = native + 3 * i;
            const int r = fmt == PFCU_TEX_RGB8 ? 0 : 2, b = 2 - r;
            if (to_native) { const uint32_t c = canon[i]; q[r] = (unsigned char)c; q[1] = (unsigned char)(c >> 8); q[b] = (unsigned char)(c >> 16); }
            else canon[i] = (uint32_t)q[r] | ((uint32_t)q[1] << 8) | ((uint32_t)q[b] << 16) | 0xff000000u;
        }
    }
}

__global__ void k_pack_tiles(uint32_t *color, float *depth, int W, int H, int tilesX, unsigned nTiles,
                             unsigned rank, unsigned world, int with_depth, uint32_t *staging, int unpack)
{
    const unsigned tile = rank + blockIdx.x * world;
    if (tile >= nTiles) return;
    const int X0 = (tile % tilesX) * TILE, Y0 = (tile / tilesX) * TILE;
    uint32_t *sc = staging + (size_t)blockIdx.x * TILE_PIX * (with_depth ? 2 : 1);
    uint32_t *sd = sc + TILE_PIX;
    for (int k = threadIdx.x; k < TILE_PIX; k += blockDim.x) {
        const int x = X0 + (k & (TILE - 1)), y = Y0 + (k >> 6);
        if (x >= W || y >= H) continue;
        const size_t gi = (size_t)y * W + x;
        if (unpack) { color[gi] = sc[k]; if (with_depth) depth[gi] = __uint_as_float(sd[k]); }
        else { sc[k] = color[gi]; if (with_depth) sd[k] = __float_as_uint(depth[gi]); }
    }
}

/* owned tiles -> the presenting rank's surface (peer memory) or the caller's page-locked buffer (multi-device read-back),
 * 128 bits per access when the row layout allows.  A CTA takes tiles first_owned + blockIdx.x, + gridDim.x, ... of the
 * `count` given: one tile per CTA towards a peer over NVLink; a few dozen looping CTAs towards the host - PCIe stores
 * are slow, and a CTA waiting for them holds an SM slot that the rasterisation of the next band wants. */
__global__ void __launch_bounds__(256)
k_push_tiles(const uint32_t *__restrict__ color, const float *__restrict__ depth, uint32_t *__restrict__ peer_color, float *__restrict__ peer_depth,
             int W, int H, int tilesX, unsigned nTiles, unsigned rank, unsigned world, unsigned first_owned, unsigned count)
{
    for (unsigned k0 = blockIdx.x; k0 < count; k0 += gridDim.x) {
        const unsigned tile = rank + (first_owned + k0) * world;
        if (tile >= nTiles) return;
        const int X0 = (tile % tilesX) * TILE, Y0 = (tile / tilesX) * TILE;
        if ((W & 3) == 0 && X0 + TILE <= W) {
            for (int k = threadIdx.x; k < TILE * 16; k += 256) {
                const int r = k >> 4, c4 = (k & 15) << 2;
                if (Y0 + r >= H) break;
                const size_t gi = (size_t)(Y0 + r) * W + X0 + c4;
                __stcs(reinterpret_cast<uint4 *>(peer_color + gi), __ldcs(reinterpret_cast<const uint4 *>(color + gi)));
                if (peer_depth) __stcs(reinterpret_cast<float4 *>(peer_depth + gi), __ldcs(reinterpret_cast<const float4 *>(depth + gi)));
            }
        } else {
            for (int k = threadIdx.x; k < TILE_PIX; k += 256) {
                const int x = X0 + (k & (TILE - 1)), y = Y0 + (k >> 6);
                if (x >= W || y >= H) continue;
                const size_t gi = (size_t)y * W + x;
                peer_color[gi] = color[gi];
                if (peer_depth) peer_depth[gi] = depth[gi];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* kernels: full-surface operations (pfRect*, pfFogProcess, pfDrawPixels,          */
/* pfReadPixels).  Scalar blend / depth tables of the reference: pfp_blend / pfp_depth (pf_prims.h). */
/* ------------------------------------------------------------------------------------------------ */

/* (PFsizei)float as x86-64 gcc compiles it: CVTTSS2SI to a 64-bit register, low 32 bits kept; out of range -> 0 */
__device__ __forceinline__ uint32_t f2u_x86(float f)
{
    if (!(f < 9.2233720368547758e18f && f >= -9.2233720368547758e18f)) return 0u;
    return (uint32_t)(unsigned long long)__float2ll_rz(f);
}

/* pfRectf, context.c:1972-1976 */
__global__ void __launch_bounds__(256)
k_rect(uint32_t *__restrict__ color, uint32_t W, uint32_t npix, int x1, int y1, uint32_t cols, uint32_t rows, uint32_t rgba)
{
    const size_t n = (size_t)cols * rows, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t r = (uint32_t)(i / cols), c = (uint32_t)(i - (size_t)r * cols);
        const uint32_t o = (uint32_t)(y1 + (int)r) * W + (uint32_t)(x1 + (int)c);
        if (o < npix) color[o] = rgba;
    }
}

struct FogArgs { float start, end, inv_len; uint32_t rgba, mode, n_thr, alpha_or; };

/* pfFogProcess, context.c:2318-2342.  thr: the host-tabulated steps of (PFubyte)(t * alpha) for the exponential modes. */
__global__ void __launch_bounds__(256)
k_fog(uint32_t *__restrict__ color, const float *__restrict__ depth, size_t npix, FogArgs a, const float *__restrict__ thr)
{
    __shared__ float s_thr[256];
    for (unsigned k = threadIdx.x; k < 256; k += blockDim.x) s_thr[k] = k < a.n_thr ? thr[k] : 0.0f;
    __syncthreads();
    const uint32_t alpha = a.rgba >> 24, rgb = a.rgba & 0x00ffffffu;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const float d = depth[i];
        if (d >= a.end) {
            color[i] = (alpha == 255u ? a.rgba : pfp_blend(1, a.rgba, color[i])) | a.alpha_or;
        } else if (d > a.start) {
            uint32_t fa;
            if (a.mode == 0u) {
                const float t = __fmul_rn(__fsub_rn(d, a.start), a.inv_len);
                fa = (uint32_t)pfv_cvttss2si(__fmul_rn(t, __uint2float_rn(alpha))) & 255u;
            } else {
                /* number of thresholds <= d (they ascend): binary search */
                unsigned lo = 0, hi = a.n_thr;
                while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (s_thr[mid] <= d) lo = mid + 1; else hi = mid; }
                fa = lo;
            }
            color[i] = pfp_blend(1, rgb | (fa << 24), color[i]) | a.alpha_or;
        }
    }
}

/* source texels / destination pixels of pfDrawPixels / pfReadPixels: any (format, type) pair of the reference, pf_pixfmt.h */
struct PixArgs {
    const unsigned char *src; uint32_t sw, sh; int fmt;
    int xs, ys, xmin, ymin; uint32_t cols, rows;
    float inv_xlen, inv_ylen, z;
    uint32_t flags, blend_mode, depth_func, alpha_or;
};

/* one destination pixel of pfDrawPixels (context.c:2052-2075) */
__device__ __forceinline__ void draw_pixel(uint32_t *__restrict__ color, float *__restrict__ depth, uint32_t o, int x, int y, const PixArgs &a)
{
    if (!(a.flags & PFCU_ST_DEPTH_TEST) || pfp_depth((int)a.depth_func, a.z, depth[o])) {
        const float v = __fmul_rn(__int2float_rn(y - a.ys), a.inv_ylen), u = __fmul_rn(__int2float_rn(x - a.xs), a.inv_xlen);
        const uint32_t so = f2u_x86(__fmul_rn(v, __uint2float_rn(a.sh - 1u))) * a.sw + f2u_x86(__fmul_rn(u, __uint2float_rn(a.sw - 1u)));
        const uint32_t c = so < a.sw * a.sh ? pfx_get(a.src, so, a.fmt) : 0u;      /* upstream reads past the image there */
        depth[o] = a.z;
        color[o] = ((a.flags & PFCU_ST_BLEND) ? pfp_blend((int)a.blend_mode, c, color[o]) : c) | a.alpha_or;
    }
}

/* One thread per pixel of the rectangle.  When the rectangle spans columns 0 .. W (vpMax one past the right edge, SURVEY
 * Q20), pixel (y, W) and pixel (y+1, 0) share the address (y+1)*W and the reference's row-major loop applies them in
 * that order: the thread of (y+1, 0) then applies both, in order, and the thread of (y, W) stands down. */
__global__ void __launch_bounds__(256)
k_draw_pixels(uint32_t *__restrict__ color, float *__restrict__ depth, uint32_t W, uint32_t npix, PixArgs a)
{
    const bool conflict = a.xmin == 0 && a.cols == W + 1u;
    const size_t n = (size_t)a.cols * a.rows, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t r = (uint32_t)(i / a.cols), c = (uint32_t)(i - (size_t)r * a.cols);
        const int y = a.ymin + (int)r, x = a.xmin + (int)c;
        const uint32_t o = (uint32_t)y * W + (uint32_t)x;
        if (o >= npix) continue;
        if (conflict) {
            if (c == W && r + 1u < a.rows) continue;
            if (c == 0u && r > 0u) draw_pixel(color, depth, o, (int)W, y - 1, a);
        }
        draw_pixel(color, depth, o, x, y, a);
    }
}

/* pfReadPixels, context.c:2380-2394: region -> compact staging in the caller's (format, type) layout (scalar setters, pf_pixfmt.h) */
__global__ void __launch_bounds__(256)
k_read_pixels(const uint32_t *__restrict__ color, uint32_t W, uint32_t x0, uint32_t y0, uint32_t cols, uint32_t rows, int fmt, unsigned char *__restrict__ out)
{
    const size_t n = (size_t)cols * rows, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t r = (uint32_t)(i / cols), c = (uint32_t)(i - (size_t)r * cols);
        pfx_set(out, i, fmt, color[(size_t)(y0 + r) * W + x0 + c]);
    }
}
