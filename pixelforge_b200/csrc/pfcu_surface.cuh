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

/* owned tiles -> the presenting rank's surface (peer memory), 128 bits per access when the row layout allows */
__global__ void __launch_bounds__(256)
k_push_tiles(const uint32_t *__restrict__ color, const float *__restrict__ depth, uint32_t *__restrict__ peer_color, float *__restrict__ peer_depth,
             int W, int H, int tilesX, unsigned nTiles, unsigned rank, unsigned world)
{
    const unsigned tile = rank + blockIdx.x * world;
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
