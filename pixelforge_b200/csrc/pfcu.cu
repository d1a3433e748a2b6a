/*
 * pfcu.cu - sm_100a implementation of the pfcu C-ABI (include/pfcu.h): the per-fragment triangle
 * path of PixelForge on a B200.
 *
 * Pipeline per submitted batch (each surface's work stays on one stream "lane", order preserving):
 *   k_vertex_* (optional) device vertex stage for large vertex-array draws: pf_vstage.h compiled as device
 *             code (transform, Phong prologue, clipping, projection), count -> scan -> emit keeps order;
 *   k_setup   one thread per triangle: integer snap, signed area / face cull, bbox, int32 edge
 *             constants, 1/sum (reference: triangles.c:294-349); writes bbox[], TriSetup[], TriData[];
 *   k_bin_*   coarse binning (256x256 px bins) with warp-ballot compaction; per-bin triangle lists
 *             keep submission order (count -> scan -> ordered fill);
 *   k_raster  one CTA per 64x64 screen tile (or a 64x32 / 64x16 slice of it): colour + depth tile staged in
 *             shared memory, the bin's list is filtered against the tile (bbox + edge-function reject) by
 *             ballot compaction into a shared queue, and every warp walks the queue IN ORDER over the
 *             8x4-pixel blocks it owns (fixed pixel ownership => blending/depth order equals
 *             submission order without atomics).  Coverage, depth, colour interpolation,
 *             texturing, per-fragment Blinn-Phong and blending restate the reference's AVX2 lane
 *             arithmetic bit for bit (triangles.c:400-529, color.h, sampler.h, blend.h, depth.h,
 *             lighting.c:148-258, simd.h cephes log/exp); RCPPS/RSQRTPS come from host-harvested
 *             tables.  Tile load/store is 128-bit vectorised and coalesced.
 * No tensor cores: nothing on this path is a dense contraction.  Compile with -fmad=false.
 */
#include "pfcu.h"
#include "pf_vstage.h"
#include "pf_prims.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <math.h>

#include <vector>
#include <mutex>

/* ------------------------------------------------------------------------------------------------ */
/* configuration                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

#define TILE        64              /* screen tile edge in pixels                                    */
#define TILE_PIX    (TILE * TILE)
/* A bin is 2^k x 2^k pixels, chosen per batch: 256 (4x4 tiles) for batches of large triangles, where a
 * triangle would otherwise be listed in thousands of bins; 64 (= one tile) for batches of many small ones,
 * where every slice CTA of the rasteriser would otherwise scan a 16 times longer list. */
#define BIN_SHIFT_COARSE 8
#define BIN_SHIFT_FINE   6
#define RASTER_THREADS 256
#define QUEUE_CAP   1024            /* triangle indices buffered per tile between raster passes      */
#define SETUP_THREADS 256
#define BIN_BATCH   1024            /* triangles per binning CTA                                     */
#define MAX_BINS    12000           /* bin counters live in dynamic shared memory (48 KB); 7680x4320 in 64 px bins = 8160 */

#define TF_VALID    1u              /* survived cull and has a non-empty bbox on the surface         */
#define TF_SAFE     2u              /* int32 edge functions cannot wrap inside the bbox              */

struct __align__(16) TriSetup {     /* 48 B */
    int   w1R, w2R, w3R; float invSum;
    int   w1X, w1Y, w2X, w2Y;
    int   w3X, w3Y; unsigned flags; unsigned pad;
};

struct __align__(16) TriData {      /* 160 B, grouped by what each fragment program reads */
    float z1, z2, z3; unsigned meta;            /* meta: state | face<<24 | is3d<<25               */
    unsigned c1, c2, c3, pad0;
    float u1, u2, u3, pad1;
    float v1, v2, v3, pad2;
    float px[4], py[4], pz[4];
    float nx[4], ny[4], nz[4];
};

struct DevLight { float pos[3], dir[3], inner, outer, attc, attl, attq; unsigned ambient, diffuse, specular; };
struct DevMaterial { unsigned ambient, diffuse, specular, emission; float shininess; };

struct __align__(16) DevState {
    unsigned flags;
    unsigned char blend_mode, depth_func, tex_filter, tex_wrap;
    int vp_min[2], vp_max[2];
    const unsigned char *tex; unsigned tw, th; int tfmt;
    unsigned n_lights;
    float tex_fw, tex_fh, tex_tx, tex_ty;       /* (float)tw, (float)th, 1/tw, 1/th: bilinear constants (sampler.h:290-300) */
    DevLight lights[8];
    DevMaterial material[2];
    float view_pos[3];
};

static_assert(offsetof(DevState, tex_fw) % 16 == 0, "tex_fw..tex_ty are fetched as one float4");

struct pfcu_surface {
    uint32_t w, h; uint32_t *color; float *depth; bool owned;
    uint32_t rank, world; uint32_t tiles_x, tiles_y;
    int lane; cudaEvent_t done; bool has_done;      /* last work enqueued on this surface */
};
struct pfcu_texture { uint32_t w, h; int fmt; unsigned char *pixels; bool owned; pfcu_surface *alias; };
struct pfcu_batch {
    DevState *states; uint32_t n_states; pfcu_triangle *tris; uint32_t n_tris; unsigned feature_mask; int single_prog;
    std::vector<pfcu_surface *> deps;
};

/* ------------------------------------------------------------------------------------------------ */
/* runtime state                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

struct PinnedBlock { void *p; size_t bytes; cudaEvent_t done; bool pending; };

/* A lane = one CUDA stream plus the scratch buffers of the batches in flight on it.  Surfaces are spread
 * round-robin over the lanes so that independent contexts (BASELINE config C5) overlap on the GPU; work on
 * one surface always stays on its lane, which keeps it ordered. */
struct Lane {
    cudaStream_t stream = nullptr; bool own_stream = false;
    pfcu_triangle *d_tris = nullptr; size_t cap_tris = 0;
    DevState *d_states = nullptr; size_t cap_states = 0;
    int4 *d_bbox = nullptr; TriSetup *d_setup = nullptr; TriData *d_data = nullptr; size_t cap_setup = 0;
    unsigned *d_bin_counts = nullptr; size_t cap_bin_counts = 0;     /* [batches][bins] */
    uint2 *d_bin_list = nullptr; size_t cap_bin_list = 0;       /* {triangle, bbox relative to the bin} */
    unsigned *d_bin_start = nullptr;                                  /* [MAX_BINS+2] starts, then totals */
    unsigned char *d_varrays = nullptr; size_t cap_varrays = 0;        /* vertex arrays of the current draw */
    unsigned *d_vcounts = nullptr; size_t cap_vcounts = 0;
    /* pinned staging for pageable sources */
    void *h_stage = nullptr; size_t cap_stage = 0; cudaEvent_t stage_done = nullptr;
    DevState *h_states = nullptr; size_t cap_hstates = 0; cudaEvent_t states_done = nullptr;
    cudaEvent_t fence = nullptr;
    /* raw-triangle batches: their vertex stage is counted on a side stream so that the one host wait (for the
       output triangle count) does not wait for the rasterisation queued on the lane */
    cudaStream_t vstream = nullptr; cudaEvent_t raw_done = nullptr, vready = nullptr;
    unsigned char *d_raw = nullptr; size_t cap_raw = 0;
    unsigned *h_total = nullptr;                                        /* pinned: {last offset, last count} */
    unsigned *d_total = nullptr;                                        /* output count of a sync-free raw batch */
};

#define MAX_LANES 8

struct Runtime {
    bool ok = false; int device = 0; int sms = 148;
    Lane lanes[MAX_LANES]; int n_lanes = 1; unsigned next_lane = 0;
    Lane *cur = nullptr;                                              /* lane of the surface being worked on */
    unsigned long long *d_counters = nullptr;                         /* rasterised, shaded, depth-failed */
    uint32_t *d_rcp = nullptr, *d_rsq = nullptr; int rcp_bits = 0, rsq_bits = 0;
    std::vector<PinnedBlock> pinned;
    uint64_t submitted = 0, launches = 0, bytes_h2d = 0, bytes_d2h = 0;
    bool profiling = false;
    int raster_path = PFCU_RASTER_AUTO;
    std::vector<cudaEvent_t> prof_events;       /* triples: before setup, before raster, after raster */
    std::vector<cudaEvent_t> prof_pool;
    std::vector<pfcu_surface *> deps;           /* surfaces sampled as textures by the states being submitted */
    std::recursive_mutex mu;                    /* the C-ABI is serialised: contexts on several threads share one runtime */
    char err[512] = { 0 };
};

static Runtime g;
#define LN (*g.cur)
#define API_LOCK std::lock_guard<std::recursive_mutex> api_lock_(g.mu)

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf(g.err, sizeof g.err, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    return PFCU_ERR_CUDA; } } while (0)

#define CKP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf(g.err, sizeof g.err, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    return nullptr; } } while (0)

__constant__ const uint32_t *c_rcp_tab;
__constant__ const uint32_t *c_rsq_tab;
__constant__ int c_rcp_shift, c_rsq_shift, c_rsq_bits;

/* ------------------------------------------------------------------------------------------------ */
/* device: x86 lane semantics                                                                       */
/* ------------------------------------------------------------------------------------------------ */

#define FM(a, b) __fmul_rn((a), (b))
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))
#define FD(a, b) __fdiv_rn((a), (b))

/* MINPS / MAXPS: second operand when either is NaN */
__device__ __forceinline__ float min_x86(float a, float b) { return (a < b) ? a : b; }
__device__ __forceinline__ float max_x86(float a, float b) { return (a > b) ? a : b; }
__device__ __forceinline__ float clamp_x86(float x, float lo, float hi) { return min_x86(max_x86(x, lo), hi); }

/* CVTPS2DQ (round to nearest even; 0x80000000 when out of range / NaN) */
__device__ __forceinline__ int cvt_rne_x86(float x)
{
    int r = __float2int_rn(x);
    return (fabsf(x) < 2147483648.0f) ? r : INT_MIN;
}
__device__ __forceinline__ int cvt_trunc_x86(float x)
{
    int r = __float2int_rz(x);
    return (fabsf(x) < 2147483648.0f) ? r : INT_MIN;
}

/* RCPPS via the host-harvested table (simd.h:1217-1225; see host/pf_x86approx.c) */
__device__ __forceinline__ float rcp_x86(float x)
{
    const unsigned u = __float_as_uint(x), s = u & 0x80000000u, e = (u >> 23) & 255u, m = u & 0x7fffffu;
    const unsigned tv = __ldg(c_rcp_tab + (m >> c_rcp_shift));
    const int ex = (int)(tv >> 23) + 127 - (int)e;
    unsigned r = s | ((unsigned)ex << 23) | (tv & 0x7fffffu);
    if (ex <= 0) r = s;
    if (e == 0u) r = s | 0x7f800000u;
    if (e == 255u) r = m ? (u | 0x00400000u) : s;
    return __uint_as_float(r);
}

/* RSQRTPS (simd.h:1237-1245) */
__device__ __forceinline__ float rsqrt_x86(float x)
{
    const unsigned u = __float_as_uint(x), s = u & 0x80000000u, e = (u >> 23) & 255u, m = u & 0x7fffffu;
    const unsigned odd = (e & 1u) ^ 1u;
    const int half = ((int)e - 127 - (int)odd) / 2;
    const unsigned tv = __ldg(c_rsq_tab + ((odd << c_rsq_bits) | (m >> c_rsq_shift)));
    unsigned r = ((unsigned)((int)(tv >> 23) - half) << 23) | (tv & 0x7fffffu);
    if (e == 255u) r = 0u;
    if (s) r = 0xffc00000u;
    if (e == 0u) r = s | 0x7f800000u;
    if (e == 255u && m) r = u | 0x00400000u;
    return __uint_as_float(r);
}

/* _mm256_log_ps (simd.h:183-252) */
__device__ __forceinline__ float log_cephes(float x)
{
    const bool invalid = (x <= 0.0f);
    x = max_x86(x, __uint_as_float(0x00800000u));
    int imm0 = (int)(__float_as_uint(x) >> 23);
    x = __uint_as_float((__float_as_uint(x) & ~0x7f800000u) | 0x3f000000u);
    imm0 -= 0x7f;
    float e = __int2float_rn(imm0);
    e = FA(e, 1.0f);
    const bool lt = (x < 0.707106781186547524f);
    float tmp = lt ? x : 0.0f;
    x = FS(x, 1.0f);
    e = FS(e, lt ? 1.0f : 0.0f);
    x = FA(x, tmp);
    const float z = FM(x, x);
    float y = 7.0376836292E-2f;
    y = FM(y, x); y = FA(y, -1.1514610310E-1f);
    y = FM(y, x); y = FA(y, 1.1676998740E-1f);
    y = FM(y, x); y = FA(y, -1.2420140846E-1f);
    y = FM(y, x); y = FA(y, 1.4249322787E-1f);
    y = FM(y, x); y = FA(y, -1.6668057665E-1f);
    y = FM(y, x); y = FA(y, 2.0000714765E-1f);
    y = FM(y, x); y = FA(y, -2.4999993993E-1f);
    y = FM(y, x); y = FA(y, 3.3333331174E-1f);
    y = FM(y, x);
    y = FM(y, z);
    tmp = FM(e, -2.12194440e-4f);
    y = FA(y, tmp);
    tmp = FM(z, 0.5f);
    y = FS(y, tmp);
    tmp = FM(e, 0.693359375f);
    x = FA(x, y);
    x = FA(x, tmp);
    return invalid ? __uint_as_float(0xffffffffu) : x;
}

/* _mm256_exp_ps (simd.h:254-304) */
__device__ __forceinline__ float exp_cephes(float x)
{
    x = min_x86(x, 88.3762626647949f);
    x = max_x86(x, -88.3762626647949f);
    float fx = FM(x, 1.44269504088896341f);
    fx = FA(fx, 0.5f);
    float tmp = floorf(fx);
    const float mask = (tmp > fx) ? 1.0f : 0.0f;
    fx = FS(tmp, mask);
    tmp = FM(fx, 0.693359375f);
    float z = FM(fx, -2.12194440e-4f);
    x = FS(x, tmp);
    x = FS(x, z);
    z = FM(x, x);
    float y = 1.9875691500E-4f;
    y = FM(y, x); y = FA(y, 1.3981999507E-3f);
    y = FM(y, x); y = FA(y, 8.3334519073E-3f);
    y = FM(y, x); y = FA(y, 4.1665795894E-2f);
    y = FM(y, x); y = FA(y, 1.6666665459E-1f);
    y = FM(y, x); y = FA(y, 5.0000001201E-1f);
    y = FM(y, z);
    y = FA(y, x);
    y = FA(y, 1.0f);
    int imm0 = cvt_trunc_x86(fx);
    imm0 = (int)((unsigned)imm0 + 0x7fu);
    return FM(y, __uint_as_float((unsigned)imm0 << 23));
}

/* ------------------------------------------------------------------------------------------------ */
/* device: colour arithmetic                                                                        */
/* ------------------------------------------------------------------------------------------------ */

#define CHN(c, i) ((int)(((c) >> (8 * (i))) & 255u))
#define INV255 (1.0f / 255.0f)

__device__ __forceinline__ unsigned pack4(int r, int g, int b, int a)   /* OR of shifted lanes, no masking (color.h:112-122) */
{
    return (unsigned)r | ((unsigned)g << 8) | ((unsigned)b << 16) | ((unsigned)a << 24);
}

__device__ __forceinline__ unsigned quant(float v)                      /* color.h:124-135 */
{
    /* clamp_x86 maps NaN to 0 (MAXPS returns its second operand), so the product is always in [0, 255] and
       CVTPS2DQ's out-of-range result cannot occur */
    return (unsigned)__float2int_rn(FM(clamp_x86(v, 0.0f, 1.0f), 255.0f));
}

/* (texel * frag) >> 8 per channel (blend.h:199-212) */
__device__ __forceinline__ unsigned mul_color(unsigned a, unsigned b)
{
    return pack4((CHN(a, 0) * CHN(b, 0)) >> 8, (CHN(a, 1) * CHN(b, 1)) >> 8,
                 (CHN(a, 2) * CHN(b, 2)) >> 8, (CHN(a, 3) * CHN(b, 3)) >> 8);
}

__device__ __forceinline__ unsigned color_lerp(unsigned a, unsigned b, float t)   /* color.h:137-144 (Q7 fixed) */
{
    unsigned p = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float A = FM(__int2float_rn(CHN(a, i)), INV255), B = FM(__int2float_rn(CHN(b, i)), INV255);
        p |= quant(FA(A, FM(t, FS(B, A)))) << (8 * i);
    }
    return p;
}

__device__ __forceinline__ unsigned blend_px(int mode, unsigned s, unsigned d)   /* blend.h:137-274 */
{
    int o[4];
    switch (mode) {
    case 0:
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = (CHN(s, i) + CHN(d, i)) >> 1;
        break;
    case 1: {
        const int alpha = CHN(s, 3) + 1, inv = 256 - alpha;
#pragma unroll
        for (int i = 0; i < 3; i++) o[i] = (CHN(s, i) * alpha + CHN(d, i) * inv) >> 8;
        o[3] = (255 * alpha + CHN(d, 3) * inv) >> 8;
    } break;
    case 2:
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = min(CHN(s, i) + CHN(d, i), 255);
        break;
    case 3:                                 /* "subtractive" adds (Q6) and may carry into the next channel */
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = max(CHN(s, i) + CHN(d, i), 0);
        break;
    case 4:
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = (CHN(s, i) * CHN(d, i)) >> 8;
        break;
    case 5:
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = min(((CHN(d, i) * (255 - CHN(s, i))) >> 8) + CHN(s, i), 255);
        break;
    case 6:
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = max(CHN(s, i), CHN(d, i));
        break;
    default:
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = min(CHN(s, i), CHN(d, i));
        break;
    }
    return pack4(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ bool depth_pass(int func, float z, float zb)   /* depth.h:80-114; NOTEQUAL == EQUAL (Q5) */
{
    switch (func) {
    case 0: case 1: return z == zb;
    case 2: return z < zb;
    case 3: return z <= zb;
    case 4: return z > zb;
    default: return z >= zb;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* device: texturing (sampler.h:202-410)                                                            */
/* ------------------------------------------------------------------------------------------------ */

struct TexRegs {            /* texture state kept in registers while a warp stays in one state */
    const unsigned char *base; unsigned tw, th, total; float wm1, hm1; int fmt, wrap, filter;
};

__device__ __forceinline__ int tex_coord(int wrap, float t, float sm1)
{
    if (wrap == 0) {                    /* REPEAT: |RNE((t - trunc t) * (size-1))| */
        const float f = FM(FS(t, truncf(t)), sm1);
        return cvt_rne_x86(fabsf(f));   /* == |RNE(f)|: RNE is symmetric, 0x80000000 stays 0x80000000 */
    } else if (wrap == 1) {             /* MIRRORED_REPEAT */
        const float a = fabsf(t);
        float m = FS(a, FM(floorf(FD(a, 2.0f)), 2.0f));
        const float r = FS(1.0f, FS(m, 1.0f));
        if (m > 1.0f) m = r;
        return cvt_rne_x86(FA(FM(m, sm1), 0.5f));
    } else {                            /* CLAMP_TO_EDGE */
        return cvt_rne_x86(FA(FM(clamp_x86(t, 0.0f, 1.0f), sm1), 0.5f));
    }
}

__device__ __forceinline__ unsigned tex_fetch(const TexRegs &t, int x, int y)
{
    const int off = (int)((unsigned)y * t.tw + (unsigned)x);
    /* the reference reads out of bounds here (CLAMP/MIRROR round v*(h-1)+0.5 up to row h); defined as
       "memory after the texture reads as zero": RGBA 0, and alpha 255 for the 3-byte formats */
    if ((unsigned)off >= t.total) return (t.fmt >= PFCU_TEX_RGB8) ? 0xff000000u : 0u;
    if (t.fmt == PFCU_TEX_RGBA8) return __ldg((const unsigned *)t.base + off);
    if (t.fmt == PFCU_TEX_BGRA8) { const unsigned r = __ldg((const unsigned *)t.base + off); return __byte_perm(r, 0, 0x3012); }
    const unsigned char *p = t.base + 3 * (size_t)off;
    const unsigned b0 = __ldg(p), b1 = __ldg(p + 1), b2 = __ldg(p + 2);
    return (t.fmt == PFCU_TEX_RGB8) ? (b0 | (b1 << 8) | (b2 << 16) | 0xff000000u) : (b2 | (b1 << 8) | (b0 << 16) | 0xff000000u);
}

__device__ __forceinline__ unsigned tex_sample(const TexRegs &t, const DevState *st, float u, float v)
{
    const int x0 = tex_coord(t.wrap, u, t.wm1), y0 = tex_coord(t.wrap, v, t.hm1);
    if (t.filter == 0) return tex_fetch(t, x0, y0);
    const float4 k = __ldg(reinterpret_cast<const float4 *>(&st->tex_fw));       /* fw, fh, 1/fw, 1/fh */
    const float fw = k.x, fh = k.y, tx = k.z, ty = k.w;
    const int x1 = tex_coord(t.wrap, FA(u, tx), t.wm1), y1 = tex_coord(t.wrap, FA(v, ty), t.hm1);
    const float fx = clamp_x86(FS(FM(u, fw), __int2float_rn(x0)), 0.0f, 1.0f);
    const float fy = clamp_x86(FS(FM(v, fh), __int2float_rn(y0)), 0.0f, 1.0f);
    const unsigned c00 = tex_fetch(t, x0, y0), c10 = tex_fetch(t, x1, y0);
    const unsigned c01 = tex_fetch(t, x0, y1), c11 = tex_fetch(t, x1, y1);
    return color_lerp(color_lerp(c00, c10, fx), color_lerp(c01, c11, fx), fy);
}

/* ------------------------------------------------------------------------------------------------ */
/* device: per-fragment Blinn-Phong (lighting.c:148-258)                                            */
/* ------------------------------------------------------------------------------------------------ */

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    return FA(FA(FM(ax, bx), FM(ay, by)), FM(az, bz));
}

__device__ __noinline__ unsigned phong(unsigned frag, const DevState *st, int face,
                                       float Px, float Py, float Pz, float Nx, float Ny, float Nz)
{
    const DevMaterial *m = &st->material[face];
    float D[3], A[3], S[3], acc[3] = { 0.0f, 0.0f, 0.0f };
#pragma unroll
    for (int i = 0; i < 3; i++) {
        D[i] = FM(__int2float_rn(CHN(frag, i)), INV255);
        A[i] = FM(FM(__int2float_rn(CHN(m->ambient, i)), INV255), D[i]);
        S[i] = FM(__int2float_rn(CHN(m->specular, i)), INV255);
    }
    float Vx = FS(st->view_pos[0], Px), Vy = FS(st->view_pos[1], Py), Vz = FS(st->view_pos[2], Pz);
    {
        const float inv = rsqrt_x86(max_x86(dot3(Vx, Vy, Vz, Vx, Vy, Vz), 1e-5f));
        Vx = FM(Vx, inv); Vy = FM(Vy, inv); Vz = FM(Vz, inv);
    }
    const float shininess = m->shininess;
    for (unsigned li = 0; li < st->n_lights; li++) {
        const DevLight *l = &st->lights[li];
        float Lx = FS(l->pos[0], Px), Ly = FS(l->pos[1], Py), Lz = FS(l->pos[2], Pz);
        {
            const float inv = rsqrt_x86(max_x86(dot3(Lx, Ly, Lz, Lx, Ly, Lz), 1e-5f));
            Lx = FM(Lx, inv); Ly = FM(Ly, inv); Lz = FM(Lz, inv);
        }
        const float diff = max_x86(dot3(Nx, Ny, Nz, Lx, Ly, Lz), 0.0f);
        float Hx = FA(Lx, Vx), Hy = FA(Ly, Vy), Hz = FA(Lz, Vz);
        {
            const float inv = rsqrt_x86(dot3(Hx, Hy, Hz, Hx, Hy, Hz));      /* no epsilon here */
            Hx = FM(Hx, inv); Hy = FM(Hy, inv); Hz = FM(Hz, inv);
        }
        float spec = max_x86(dot3(Nx, Ny, Nz, Hx, Hy, Hz), 0.0f);
        spec = exp_cephes(FM(log_cephes(spec), shininess));                 /* pow(0) -> e^88 (Q9) */
        float inten = 1.0f; bool spot = false;
        if (l->inner < 3.14159265358979323846f) {
            spot = true;
            const float theta = dot3(Lx, Ly, Lz, FS(0.0f, l->dir[0]), FS(0.0f, l->dir[1]), FS(0.0f, l->dir[2]));
            inten = clamp_x86(FD(FS(theta, l->outer), FS(l->inner, l->outer)), 0.0f, 1.0f);
        }
        float att = 1.0f; bool atten = false;
        if (l->attl != 0.0f || l->attq != 0.0f) {
            atten = true;
            const float d0 = FS(l->pos[0], Px), d1 = FS(l->pos[1], Py);     /* y twice, z dropped (Q10) */
            const float dsq = FA(FM(d0, d0), FA(FM(d1, d1), FM(d1, d1)));
            const float dist = __fsqrt_rn(dsq);
            att = rcp_x86(FA(l->attc, FA(FM(l->attl, dist), FM(l->attq, dsq))));
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
            float amb = FM(FM(__int2float_rn(CHN(l->ambient, i)), INV255), A[i]);
            float dif = FM(FM(FM(__int2float_rn(CHN(l->diffuse, i)), INV255), diff), D[i]);
            float spc = FM(FM(FM(__int2float_rn(CHN(l->specular, i)), INV255), spec), S[i]);
            if (spot) { dif = FM(dif, inten); spc = FM(spc, inten); }
            if (atten) { amb = FM(amb, att); dif = FM(dif, att); spc = FM(spc, att); }
            acc[i] = FA(acc[i], amb); acc[i] = FA(acc[i], dif); acc[i] = FA(acc[i], spc);
        }
    }
    return quant(acc[0]) | (quant(acc[1]) << 8) | (quant(acc[2]) << 16);   /* alpha = 0 (Q9) */
}

/* ------------------------------------------------------------------------------------------------ */
/* kernels: setup                                                                                   */
/* ------------------------------------------------------------------------------------------------ */

__device__ __forceinline__ int to_int_x86(float f) { return cvt_trunc_x86(f); }   /* (PFint)f == CVTTSS2SI */
__device__ __forceinline__ int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
__device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__device__ __forceinline__ int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__device__ __forceinline__ long long labs64(long long v) { return v < 0 ? -v : v; }

/* setup of triangle i; returns its bbox (empty = (1,1,0,0)) and whether it counts as rasterised */
__device__ __forceinline__ int4 setup_one(const pfcu_triangle *__restrict__ tris, const DevState *__restrict__ states, unsigned i,
                                          int surfW, int surfH, int4 *__restrict__ bbox, TriSetup *__restrict__ setup,
                                          TriData *__restrict__ data, bool *rasterised)
{
    bool valid = false;
    int4 out_box = make_int4(1, 1, 0, 0);
    {
        const pfcu_triangle *t = tris + i;
        const pfcu_vertex *v1 = &t->v[0], *v2 = &t->v[1], *v3 = &t->v[2];
        const int face = t->face, is3d = t->is3d;
        const DevState *st = states + t->state;

        const int x1 = to_int_x86(v1->sx), y1 = to_int_x86(v1->sy);
        const int x2 = to_int_x86(v2->sx), y2 = to_int_x86(v2->sy);
        const int x3 = to_int_x86(v3->sx), y3 = to_int_x86(v3->sy);

        /* signed area in wrapping int32, compared as float like the reference (triangles.c:303-308) */
        const float area = __int2float_rn(wsub(wmul(wsub(x2, x1), wsub(y3, y1)), wmul(wsub(x3, x1), wsub(y2, y1))));
        const bool culled = (face == 0 && area >= 0.0f) || (face == 1 && area <= 0.0f);

        int xMin = min(x1, min(x2, x3)), yMin = min(y1, min(y2, y3));
        int xMax = max(x1, max(x2, x3)), yMax = max(y1, max(y2, y3));
        if (!is3d) {
            xMin = min(max(xMin, st->vp_min[0]), st->vp_max[0]); yMin = min(max(yMin, st->vp_min[1]), st->vp_max[1]);
            xMax = min(max(xMax, st->vp_min[0]), st->vp_max[0]); yMax = min(max(yMax, st->vp_min[1]), st->vp_max[1]);
        }
        int w1X = wsub(y3, y2), w1Y = wsub(x2, x3);
        int w2X = wsub(y1, y3), w2Y = wsub(x3, x1);
        int w3X = wsub(y2, y1), w3Y = wsub(x1, x2);
        if (face == 1) { w1X = wsub(0, w1X); w1Y = wsub(0, w1Y); w2X = wsub(0, w2X); w2Y = wsub(0, w2Y); w3X = wsub(0, w3X); w3Y = wsub(0, w3Y); }
        const int w1R = wadd(wmul(wsub(xMin, x2), w1X), wmul(w1Y, wsub(yMin, y2)));
        const int w2R = wadd(wmul(wsub(xMin, x3), w2X), wmul(w2Y, wsub(yMin, y3)));
        const int w3R = wadd(wmul(wsub(xMin, x1), w3X), wmul(w3Y, wsub(yMin, y1)));
        const float invSum = FD(1.0f, __int2float_rn(wadd(wadd(w1R, w2R), w3R)));

        /* can any edge function leave int32 inside the bbox?  (exact 64-bit bound) */
        const long long bw = (long long)xMax - xMin, bh = (long long)yMax - yMin;
        const long long r1 = ((long long)xMin - x2) * w1X + (long long)w1Y * ((long long)yMin - y2);
        const long long r2 = ((long long)xMin - x3) * w2X + (long long)w2Y * ((long long)yMin - y3);
        const long long r3 = ((long long)xMin - x1) * w3X + (long long)w3Y * ((long long)yMin - y1);
        const long long lim = 0x7fffffffLL;
        const bool coords_ok = labs64(x1) < (1 << 24) && labs64(y1) < (1 << 24) && labs64(x2) < (1 << 24) &&
                               labs64(y2) < (1 << 24) && labs64(x3) < (1 << 24) && labs64(y3) < (1 << 24);
        const bool safe = coords_ok &&
            labs64(r1) + labs64(w1X) * bw + labs64(w1Y) * bh < lim &&
            labs64(r2) + labs64(w2X) * bw + labs64(w2Y) * bh < lim &&
            labs64(r3) + labs64(w3X) * bw + labs64(w3Y) * bh < lim;

        /* clip the visited rectangle to the surface: x in [xMin, xMax-1], y in [yMin, yMax] */
        const bool nonempty = !culled && xMin < xMax && yMin <= yMax && xMax > 0 && yMax >= 0 && xMin < surfW && yMin < surfH;
        valid = nonempty;

        if (valid) out_box = make_int4(xMin, yMin, xMax, yMax);
        bbox[i] = out_box;
        TriSetup s;
        s.w1R = w1R; s.w2R = w2R; s.w3R = w3R; s.invSum = invSum;
        s.w1X = w1X; s.w1Y = w1Y; s.w2X = w2X; s.w2Y = w2Y; s.w3X = w3X; s.w3Y = w3Y;
        s.flags = (valid ? TF_VALID : 0u) | (safe ? TF_SAFE : 0u); s.pad = 0;
        setup[i] = s;
        if (valid) {
            TriData d;
            d.z1 = v1->zinv; d.z2 = v2->zinv; d.z3 = v3->zinv;
            d.meta = (t->state & 0xffffffu) | ((unsigned)face << 24) | ((unsigned)(is3d ? 1 : 0) << 25);
            d.c1 = v1->rgba; d.c2 = v2->rgba; d.c3 = v3->rgba; d.pad0 = 0;
            d.u1 = v1->u; d.u2 = v2->u; d.u3 = v3->u; d.pad1 = 0;
            d.v1 = v1->v; d.v2 = v2->v; d.v3 = v3->v; d.pad2 = 0;
            d.px[0] = v1->px; d.px[1] = v2->px; d.px[2] = v3->px; d.px[3] = 0;
            d.py[0] = v1->py; d.py[1] = v2->py; d.py[2] = v3->py; d.py[3] = 0;
            d.pz[0] = v1->pz; d.pz[1] = v2->pz; d.pz[2] = v3->pz; d.pz[3] = 0;
            d.nx[0] = v1->nx; d.nx[1] = v2->nx; d.nx[2] = v3->nx; d.nx[3] = 0;
            d.ny[0] = v1->ny; d.ny[1] = v2->ny; d.ny[2] = v3->ny; d.ny[3] = 0;
            d.nz[0] = v1->nz; d.nz[1] = v2->nz; d.nz[2] = v3->nz; d.nz[3] = 0;
            data[i] = d;
        }
        /* "rasterised" = survives the face / zero-area test */
        valid = !culled;
    }
    *rasterised = valid;
    return out_box;
}

__global__ void __launch_bounds__(SETUP_THREADS)
k_setup(const pfcu_triangle *__restrict__ tris, const DevState *__restrict__ states, unsigned n,
        int surfW, int surfH, int4 *__restrict__ bbox, TriSetup *__restrict__ setup, TriData *__restrict__ data,
        unsigned long long *__restrict__ counters)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    if (i < n) setup_one(tris, states, i, surfW, surfH, bbox, setup, data, &valid);
    const unsigned b = __ballot_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(counters + 0, (unsigned long long)__popc(b));
}

/* ------------------------------------------------------------------------------------------------ */
/* kernels: order-preserving coarse binning                                                         */
/* ------------------------------------------------------------------------------------------------ */

/* pass 1: counts[batch][bin] = number of triangles of this batch whose bbox touches the bin */
__global__ void __launch_bounds__(256)
k_bin_count(const int4 *__restrict__ bbox, unsigned n, int binsX, int binsY, int bshift, unsigned *__restrict__ counts)
{
    extern __shared__ unsigned s_cnt[];
    const int nb = binsX * binsY;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) s_cnt[k] = 0;
    __syncthreads();
    const unsigned base = blockIdx.x * BIN_BATCH;
    for (unsigned k = threadIdx.x; k < BIN_BATCH; k += blockDim.x) {
        const unsigned i = base + k;
        if (i >= n) break;
        const int4 b = __ldg(bbox + i);
        if (b.x >= b.z) continue;
        const int bx0 = max(b.x, 0) >> bshift, bx1 = min((b.z - 1) >> bshift, binsX - 1);
        const int by0 = max(b.y, 0) >> bshift, by1 = min(b.w >> bshift, binsY - 1);
        for (int by = by0; by <= by1; by++)
            for (int bx = bx0; bx <= bx1; bx++) atomicAdd(&s_cnt[by * binsX + bx], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nb; k += blockDim.x) counts[(size_t)blockIdx.x * nb + k] = s_cnt[k];
}

/* pass 2: per bin, exclusive scan of counts over batches (in place) + bin totals.  A CTA of 32 warps owns 32
 * consecutive bins (one 128-byte row segment per batch); warp w owns a contiguous range of batches: it sums its
 * range, the 32 partial sums are scanned across warps, and it walks its range again writing the prefixes. */
__global__ void __launch_bounds__(1024)
k_bin_scan(unsigned *__restrict__ counts, int nBatches, int nb, unsigned *__restrict__ totals)
{
    __shared__ unsigned s_part[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bin = blockIdx.x * 32 + lane;
    const bool live = bin < nb;
    const int per = (nBatches + 31) / 32;
    const int k0 = warp * per, k1 = min(k0 + per, nBatches);
    unsigned sum = 0;
    if (live) {
#pragma unroll 8
        for (int k = k0; k < k1; k++) sum += counts[(size_t)k * nb + bin];
    }
    s_part[warp][lane] = sum;
    __syncthreads();
    unsigned run = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 32; w++) { const unsigned c = s_part[w][lane]; if (w < warp) run += c; total += c; }
    if (live) {
        for (int k = k0; k < k1; k++) {
            unsigned *pc = counts + (size_t)k * nb + bin;
            const unsigned v = *pc; *pc = run; run += v;
        }
        if (warp == 0) totals[bin] = total;
    }
}

/* pass 3: bin start offsets (exclusive scan over the bin totals); single CTA of 1024 threads */
__global__ void __launch_bounds__(1024)
k_bin_starts(const unsigned *__restrict__ totals, int nb, unsigned *__restrict__ starts)
{
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int k = base + threadIdx.x;
        const unsigned v = (k < nb) ? totals[k] : 0u;
        unsigned x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            s_warp[lane] = w;                   /* inclusive over warps */
        }
        __syncthreads();
        const unsigned carry = s_carry, woff = warp ? s_warp[warp - 1] : 0u;
        if (k < nb) starts[k] = carry + woff + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + woff + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) starts[nb] = s_carry;
}

/* A bin-list entry carries the triangle's visited rectangle [x0, x1] x [y0, y1] (inclusive) clipped to the bin and
 * relative to the bin's origin, 8 bits per coordinate (bins are at most 256 pixels wide): the rasteriser's
 * queue filter then needs no dependent load. */
__device__ __forceinline__ unsigned bin_rel_bbox(const int4 b, int bx, int by, int bshift)
{
    const int ox = bx << bshift, oy = by << bshift, hi = (1 << bshift) - 1;
    const int x0 = min(max(b.x - ox, 0), hi), x1 = min(max(b.z - 1 - ox, 0), hi);
    const int y0 = min(max(b.y - oy, 0), hi), y1 = min(max(b.w - oy, 0), hi);
    return (unsigned)x0 | ((unsigned)y0 << 8) | ((unsigned)x1 << 16) | ((unsigned)y1 << 24);
}

/* pass 4: ordered fill.  The CTA walks its triangles 256 at a time.  Every bin COLUMN belongs to one warp
 * (bx & 7); each warp visits, in triangle order, the triangles whose bin rectangle has a column of its own and
 * appends them to those bins.  A bin is therefore written by one warp only, in submission order, with no
 * CTA barrier inside a group and no dependence on how the 256 triangles are spread over the screen. */
__global__ void __launch_bounds__(256)
k_bin_fill(const int4 *__restrict__ bbox, unsigned n, int binsX, int binsY, int bshift,
           const unsigned *__restrict__ offsets /* scanned counts */, const unsigned *__restrict__ starts,
           uint2 *__restrict__ list)
{
    extern __shared__ unsigned s_mem[];
    unsigned *s_pos = s_mem;                    /* [nb] running write position of this batch per bin */
    __shared__ int4 s_rect[256];
    __shared__ int4 s_bbox[256];
    const int nb = binsX * binsY;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) s_pos[k] = starts[k] + offsets[(size_t)blockIdx.x * nb + k];
    const unsigned base = blockIdx.x * BIN_BATCH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned k0 = 0; k0 < BIN_BATCH && base + k0 < n; k0 += 256) {
        const unsigned i = base + k0 + threadIdx.x;
        int4 r = make_int4(1, 1, 0, 0);
        int4 b = make_int4(1, 1, 0, 0);
        if (i < n) {
            b = __ldg(bbox + i);
            if (b.x < b.z) {
                r.x = max(b.x, 0) >> bshift; r.z = min((b.z - 1) >> bshift, binsX - 1);
                r.y = max(b.y, 0) >> bshift; r.w = min(b.w >> bshift, binsY - 1);
            }
        }
        __syncthreads();                        /* previous group done with s_rect (and s_pos initialised) */
        s_rect[threadIdx.x] = r;
        s_bbox[threadIdx.x] = b;
        __syncthreads();
        for (int g8 = 0; g8 < 8; g8++) {
            const int4 q = s_rect[g8 * 32 + lane];
            const int4 qb = s_bbox[g8 * 32 + lane];
            /* first column of the rectangle that this warp owns */
            const int first = q.x + ((warp - q.x) & 7);
            const bool mine = q.x <= q.z && q.y <= q.w && first <= q.z;
            const bool one = mine && first + 8 > q.z;                    /* exactly one owned column */
            unsigned mask = __ballot_sync(0xffffffffu, mine);
            const unsigned single = __ballot_sync(0xffffffffu, one);
            const unsigned my_idx = base + k0 + (unsigned)(g8 * 32 + lane);
            while (mask) {
                const int j = __ffs(mask) - 1;
                if ((single >> j) & 1u) {
                    /* a run of consecutive one-column triangles: bin row by bin row (a bin has one row, so all of
                       its entries are ranked in the same step), ranked per bin with one match */
                    const unsigned multi = mask & ~single;
                    const unsigned run = multi ? (mask & ((1u << (__ffs(multi) - 1)) - 1u)) : mask;
                    const bool in_run = (run >> lane) & 1u;
                    const int ylo = __reduce_min_sync(0xffffffffu, in_run ? q.y : INT_MAX);
                    const int yhi = __reduce_max_sync(0xffffffffu, in_run ? q.w : INT_MIN);
                    for (int by = ylo; by <= yhi; by++) {
                        const bool act = in_run && q.y <= by && by <= q.w;
                        const unsigned am = __ballot_sync(0xffffffffu, act);
                        if (act) {
                            const int bin = by * binsX + first;
                            const unsigned peers = __match_any_sync(am, bin);
                            const unsigned pos = s_pos[bin] + __popc(peers & ((1u << lane) - 1u));
                            list[pos] = make_uint2(my_idx, bin_rel_bbox(qb, first, by, bshift));
                            __syncwarp(peers);
                            if ((peers >> lane) == 1u) s_pos[bin] = pos + 1;   /* highest lane of the group */
                        }
                        __syncwarp();
                    }
                    mask &= ~run;
                } else {
                    mask &= mask - 1u;
                    const int4 t = s_rect[g8 * 32 + j];
                    const int4 tb = s_bbox[g8 * 32 + j];
                    const int f0 = t.x + ((warp - t.x) & 7);
                    const int ncols = ((t.z - f0) >> 3) + 1, rows = t.w - t.y + 1;
                    const unsigned idx = base + k0 + (unsigned)(g8 * 32 + j);
                    for (int e = lane; e < ncols * rows; e += 32) {
                        const int cy = e / ncols, cx = e - cy * ncols;
                        const int bin = (t.y + cy) * binsX + f0 + (cx << 3);
                        const unsigned pos = s_pos[bin];
                        list[pos] = make_uint2(idx, bin_rel_bbox(tb, f0 + (cx << 3), t.y + cy, bshift));
                        s_pos[bin] = pos + 1;
                    }
                }
                __syncwarp();
            }
        }
    }
}

/* Batches of at most 1024 triangles (a Gears frame, one context of a many-context batch): setup, bin count,
 * bin starts and the ordered fill in ONE single-CTA kernel instead of five launches; thread = triangle, warp w
 * owns the bin columns bx & 31 == w (see k_bin_fill). */
#define FRONT_SMALL_MAX 1024
#define FRONT_SMALL_CHUNKS 10           /* with a device-side count (raw batches after clipping): up to 10 x 1024 */
__global__ void __launch_bounds__(1024)
k_front_small(const pfcu_triangle *__restrict__ tris, const DevState *__restrict__ states, unsigned n_host, const unsigned *__restrict__ d_n,
              int surfW, int surfH,
              int4 *__restrict__ bbox, TriSetup *__restrict__ setup, TriData *__restrict__ data, unsigned long long *__restrict__ counters,
              int binsX, int binsY, int bshift, unsigned *__restrict__ starts, uint2 *__restrict__ list)
{
    extern __shared__ unsigned s_mem[];
    unsigned *s_pos = s_mem;                    /* [nb] counts, then running write positions */
    __shared__ int4 s_rect[FRONT_SMALL_MAX];
    __shared__ int4 s_bbox[FRONT_SMALL_MAX];
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    const unsigned n = d_n ? min(*d_n, (unsigned)(FRONT_SMALL_MAX * FRONT_SMALL_CHUNKS)) : n_host;
    const int nb = binsX * binsY;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < nb; k += 1024) s_pos[k] = 0;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();

    /* pass 1: setup + bin counts */
    for (unsigned base = 0; base < n; base += FRONT_SMALL_MAX) {
        const unsigned i = base + threadIdx.x;
        bool rasterised = false;
        int4 b = make_int4(1, 1, 0, 0);
        if (i < n) b = setup_one(tris, states, i, surfW, surfH, bbox, setup, data, &rasterised);
        if (b.x < b.z) {
            const int rx0 = max(b.x, 0) >> bshift, rx1 = min((b.z - 1) >> bshift, binsX - 1);
            const int ry0 = max(b.y, 0) >> bshift, ry1 = min(b.w >> bshift, binsY - 1);
            for (int by = ry0; by <= ry1; by++)
                for (int bx = rx0; bx <= rx1; bx++) atomicAdd(&s_pos[by * binsX + bx], 1u);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, rasterised);
        if (lane == 0 && bal) atomicAdd(counters + 0, (unsigned long long)__popc(bal));
    }
    __syncthreads();

    /* exclusive scan of the bin counts -> starts[] (global, for the rasteriser) and s_pos */
    for (int base = 0; base < nb; base += 1024) {
        const int k = base + threadIdx.x;
        const unsigned v = (k < nb) ? s_pos[k] : 0u;
        unsigned x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            s_warp[lane] = w;
        }
        __syncthreads();
        const unsigned carry = s_carry, woff = warp ? s_warp[warp - 1] : 0u;
        if (k < nb) { const unsigned e = carry + woff + x - v; s_pos[k] = e; starts[k] = e; }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + woff + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) starts[nb] = s_carry;

    /* pass 2: ordered fill, 1024 triangles at a time: every warp walks them 32 at a time and appends those with a
       bin column of its own */
    for (unsigned base = 0; base < n; base += FRONT_SMALL_MAX) {
        const unsigned i = base + threadIdx.x;
        int4 b = make_int4(1, 1, 0, 0), r = make_int4(1, 1, 0, 0);
        if (i < n) b = bbox[i];                 /* written by this very thread in pass 1 */
        if (b.x < b.z) {
            r.x = max(b.x, 0) >> bshift; r.z = min((b.z - 1) >> bshift, binsX - 1);
            r.y = max(b.y, 0) >> bshift; r.w = min(b.w >> bshift, binsY - 1);
        }
        __syncthreads();
        s_rect[threadIdx.x] = r; s_bbox[threadIdx.x] = b;
        __syncthreads();
        const unsigned groups = (min(n - base, (unsigned)FRONT_SMALL_MAX) + 31u) / 32u;
        for (unsigned g8 = 0; g8 < groups; g8++) {
            const int4 q = s_rect[g8 * 32 + lane];
            const int4 qb = s_bbox[g8 * 32 + lane];
            const int first = q.x + ((warp - q.x) & 31);
            const bool mine = q.x <= q.z && q.y <= q.w && first <= q.z;
            const bool one = mine && first + 32 > q.z;
            unsigned mask = __ballot_sync(0xffffffffu, mine);
            const unsigned single = __ballot_sync(0xffffffffu, one);
            const unsigned my_idx = base + g8 * 32 + (unsigned)lane;
            while (mask) {
                const int j = __ffs(mask) - 1;
                if ((single >> j) & 1u) {
                    const unsigned multi = mask & ~single;
                    const unsigned run = multi ? (mask & ((1u << (__ffs(multi) - 1)) - 1u)) : mask;
                    const bool in_run = (run >> lane) & 1u;
                    const int ylo = __reduce_min_sync(0xffffffffu, in_run ? q.y : INT_MAX);
                    const int yhi = __reduce_max_sync(0xffffffffu, in_run ? q.w : INT_MIN);
                    for (int by = ylo; by <= yhi; by++) {
                        const bool act = in_run && q.y <= by && by <= q.w;
                        const unsigned am = __ballot_sync(0xffffffffu, act);
                        if (act) {
                            const int bin = by * binsX + first;
                            const unsigned peers = __match_any_sync(am, bin);
                            const unsigned pos = s_pos[bin] + __popc(peers & ((1u << lane) - 1u));
                            list[pos] = make_uint2(my_idx, bin_rel_bbox(qb, first, by, bshift));
                            __syncwarp(peers);
                            if ((peers >> lane) == 1u) s_pos[bin] = pos + 1;
                        }
                        __syncwarp();
                    }
                    mask &= ~run;
                } else {
                    mask &= mask - 1u;
                    const int4 t = s_rect[g8 * 32 + j];
                    const int4 tb = s_bbox[g8 * 32 + j];
                    const int f0 = t.x + ((warp - t.x) & 31);
                    const int ncols = ((t.z - f0) >> 5) + 1, rows = t.w - t.y + 1;
                    const unsigned idx = base + g8 * 32 + (unsigned)j;
                    for (int e = lane; e < ncols * rows; e += 32) {
                        const int cy = e / ncols, cx = e - cy * ncols;
                        const int bin = (t.y + cy) * binsX + f0 + (cx << 5);
                        const unsigned pos = s_pos[bin];
                        list[pos] = make_uint2(idx, bin_rel_bbox(tb, f0 + (cx << 5), t.y + cy, bshift));
                        s_pos[bin] = pos + 1;
                    }
                }
                __syncwarp();
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* tile rasteriser: parameters, shared-tile addressing, packed colour arithmetic                    */
/* ------------------------------------------------------------------------------------------------ */

struct RasterParams {
    const int4 *bbox; const TriSetup *setup; const TriData *data; const DevState *states;
    const uint2 *bin_list; const unsigned *bin_starts; int binsX; int bin_tshift;   /* a bin is 2^bin_tshift tiles wide */
    uint32_t *color; float *depth; int W, H; int tilesX, tilesY;
    unsigned rank, world; unsigned nTiles;
    unsigned long long *counters;
};

/* swizzled tile address: rows are 64 words; XOR-ing bits 3..4 of x with (y & 3) makes both the
 * 8x4-block access of the shading loop and the 128-bit row access of load/store conflict-free */
__device__ __forceinline__ int tile_addr(int lx, int ly) { return ly * TILE + (lx ^ ((ly & 3) << 3)); }

/* ---- packed colour arithmetic ------------------------------------------------------------------
 * A colour is carried as two words with one channel per 16-bit lane: rb = r | b<<16, ga = g | a<<16.
 * Every per-channel formula of the reference keeps its intermediate below 2^16 (proofs inline), so
 * both lanes are computed by one 32-bit instruction with no cross-lane carry. */
struct Px2 { unsigned rb, ga; };

__device__ __forceinline__ Px2 px_split(unsigned c) { Px2 p; p.rb = c & 0x00ff00ffu; p.ga = (c >> 8) & 0x00ff00ffu; return p; }
/* the reference packs by OR-ing channel<<8i WITHOUT masking (color.h:112-122); lanes here may hold up
 * to 9 bits (blend "subtractive"), and OR-ing rb with ga<<8 reproduces exactly that carry-over */
__device__ __forceinline__ unsigned px_join(Px2 p) { return p.rb | (p.ga << 8); }

/* pfiColorBarySmooth_simd (color.h:153-181): ((u1*c1 + u2*c2 + u3*c3) * 257) >> 16 per channel.
 * u1+u2+u3 <= 256 for covered pixels, so a lane's sum x <= 65280; (x*257)>>16 == (x + (x>>8)) >> 8. */
__device__ __forceinline__ unsigned smooth_lanes(unsigned a, unsigned b, unsigned c, int u1, int u2, int u3)
{
    unsigned x = (unsigned)u1 * a + (unsigned)u2 * b + (unsigned)u3 * c;
    x = x + ((x >> 8) & 0x00ff00ffu);
    return (x >> 8) & 0x00ff00ffu;
}

/* (texel * frag) >> 8 per channel (blend.h:199-212).  dp2a multiplies one 16-bit lane of the fragment by
 * one byte of the texel without extracting the byte first (the other 16-bit lane is zero). */
__device__ __forceinline__ Px2 px_mul(unsigned texel, Px2 f)
{
    const unsigned r = __dp2a_lo(f.rb & 0xffffu, texel, 0u);          /* fr * texel.byte0 */
    const unsigned g = __dp2a_lo(f.ga << 16, texel, 0u);              /* fg * texel.byte1 */
    const unsigned b = __dp2a_hi(f.rb >> 16, texel, 0u);              /* fb * texel.byte2 */
    const unsigned a = __dp2a_hi(f.ga & 0xffff0000u, texel, 0u);      /* fa * texel.byte3 */
    Px2 o;
    o.rb = __byte_perm(r, b, 0x7531);      /* byte1 of each product; bytes 3 are zero */
    o.ga = __byte_perm(g, a, 0x7531);
    return o;
}

__device__ __noinline__ Px2 blend_slow(int mode, Px2 s, unsigned dst)
{
    const unsigned c = blend_px(mode, px_join(s) , dst);     /* only reached with lanes <= 255 */
    return px_split(c);
}

/* blend.h:137-274 on packed lanes */
__device__ __forceinline__ Px2 px_blend(int mode, Px2 s, unsigned dst)
{
    const Px2 d = px_split(dst);
    Px2 o;
    if (mode == 1) {                        /* ALPHA: (s*a + d*(256-a)) >> 8, a = s.a + 1; sums <= 255*256 */
        const unsigned alpha = (s.ga >> 16) + 1u, inv = 256u - alpha;
        o.rb = ((s.rb * alpha + d.rb * inv) >> 8) & 0x00ff00ffu;
        o.ga = ((((s.ga & 0xffffu) | 0x00ff0000u) * alpha + d.ga * inv) >> 8) & 0x00ff00ffu;
    } else if (mode == 2) {                 /* ADD: min(s + d, 255) */
        const unsigned rb = s.rb + d.rb, ga = s.ga + d.ga;
        o.rb = __vminu2(rb, 0x00ff00ffu); o.ga = __vminu2(ga, 0x00ff00ffu);
    } else if (mode == 0) {                 /* AVERAGE */
        o.rb = ((s.rb + d.rb) >> 1) & 0x00ff00ffu; o.ga = ((s.ga + d.ga) >> 1) & 0x00ff00ffu;
    } else if (mode == 3) {                 /* "SUB" adds without an upper clamp (Q6): lanes reach 510 */
        o.rb = s.rb + d.rb; o.ga = s.ga + d.ga;
    } else if (mode == 6) {
        o.rb = __vmaxu2(s.rb, d.rb); o.ga = __vmaxu2(s.ga, d.ga);
    } else if (mode == 7) {
        o.rb = __vminu2(s.rb, d.rb); o.ga = __vminu2(s.ga, d.ga);
    } else o = blend_slow(mode, s, dst);    /* MUL, SCREEN */
    return o;
}

/* depth.h:80-114 as a 3-bit mask over {less, equal, greater}; false on NaN like the ordered compares */
__device__ __forceinline__ unsigned depth_mask(int func)
{
    return (0x643122u >> (4 * func)) & 7u;   /* nibbles, low first: EQ 2, NEQ 2 (Q5), LT 1, LE 3, GT 4, GE 6 */
}

/* The compare selected by the warp-uniform zmask as three predicated compares (a switch or an if-chain
 * over the function both compile to a jump table inside the block loop). */
__device__ __forceinline__ bool depth_pass_mask(float z, float zb, unsigned zmask)
{
    unsigned r;
    asm("{\n\t.reg .pred pl, pe, pg;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %3, 1;\n\tsetp.ne.u32 pl, t, 0;\n\t"
        "and.b32 t, %3, 2;\n\tsetp.ne.u32 pe, t, 0;\n\t"
        "and.b32 t, %3, 4;\n\tsetp.ne.u32 pg, t, 0;\n\t"
        "setp.lt.and.f32 pl, %1, %2, pl;\n\tsetp.eq.and.f32 pe, %1, %2, pe;\n\tsetp.gt.and.f32 pg, %1, %2, pg;\n\t"
        "or.pred pl, pl, pe;\n\tor.pred pl, pl, pg;\n\tselp.u32 %0, 1, 0, pl;\n\t}"
        : "=r"(r) : "f"(z), "f"(zb), "r"(zmask));
    return r != 0u;
}

/* shared-memory access through 32-bit window addresses computed once per CTA (the compiler otherwise
 * rebuilds the cluster-window base of every __shared__ array at each access) */
#define SM_COLOR 0          /* byte offsets inside the CTA's shared block */
#define SM_DEPTH 16384
#define SM_RCP   32768
__device__ __forceinline__ unsigned lds_u32(unsigned addr) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ unsigned lds_color(unsigned addr) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ float lds_depth(unsigned addr) { float v; asm volatile("ld.shared.f32 %0, [%1+16384];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts_color(unsigned addr, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_depth(unsigned addr, float v) { asm volatile("st.shared.f32 [%0+16384], %1;" :: "r"(addr), "f"(v) : "memory"); }

/* RCPPS from the shared-memory copy of the table (fast path: normal input, normal result) */
__device__ __forceinline__ float rcp_fast(unsigned tab_addr, int shift, float x)
{
    const unsigned u = __float_as_uint(x), E = u & 0x7f800000u;
    if (E - 0x00800000u >= 0x7e000000u) return rcp_x86(x);              /* zero/denormal/huge/inf/NaN */
    unsigned tv; asm volatile("ld.shared.u32 %0, [%1+32768];" : "=r"(tv) : "r"(tab_addr + (((u & 0x007fffffu) >> shift) << 2)));
    return __uint_as_float((tv + 0x3f800000u - E) | (u & 0x80000000u));
}

/* ------------------------------------------------------------------------------------------------ */
/* kernel: tile rasteriser                                                                          */
/* ------------------------------------------------------------------------------------------------ */

#define RCP_SMEM_BITS 11

struct TileCtx {
    int X0, Y0, X1, Y1;                 /* tile rectangle on the surface, inclusive               */
    unsigned sm_base;                   /* shared-window byte address of the CTA's block (opaque)  */
    int rcp_shift; bool rcp_shared;
    int lx8, ly4, warp;
    unsigned lane_rel;                  /* byte offset of this lane's pixel in block (0,0), XOR term folded in (opaque) */
    unsigned shaded, covered;
    const TriData *data;
};

/* One triangle over the 8x4 blocks this warp owns.  TEXM: 0 no texture, 1 nearest+REPEAT+RGBA8,
 * 2 any sampler.  BLENDM: 0 off, 1 ALPHA, 2 ADD, 3 any mode.  Everything is computed for all 32 lanes
 * (no divergent regions); only the final stores are predicated by the coverage/depth mask.
 * BIG: the launch is a batch of large triangles in ONE state program with the RCPPS table in shared memory
 * (launch_pipeline checks both), so some per-block early-outs and run-time checks are dropped. */
template <int TEXM, int BLENDM, bool PHONG, int NW, bool BIG>
__device__ __forceinline__ void shade_tri(TileCtx &t, const unsigned ti, const int4 b, const TriSetup &s, const uint4 a0, const uint4 a1,
                                          const DevState *st, const unsigned flags, const unsigned zmask, const int blend_mode, const TexRegs &tex)
{
    const int cx0 = max(b.x, t.X0) - t.X0, cx1 = min(b.z - 1, t.X1) - t.X0;     /* tile-local, inclusive */
    const int cy0 = max(b.y, t.Y0) - t.Y0, cy1 = min(b.w, t.Y1) - t.Y0;
    const int bx0 = cx0 >> 3, bx1 = cx1 >> 3, by0 = cy0 >> 2, by1 = cy1 >> 2;
    const unsigned xspan = (unsigned)(cx1 - cx0), yspan = (unsigned)(cy1 - cy0);
    const float z1 = __uint_as_float(a0.x), z2 = __uint_as_float(a0.y), z3 = __uint_as_float(a0.z);
    const unsigned meta = a0.w;
    const bool is3d = (meta >> 25) & 1u;
    const bool smooth = (flags & PFCU_ST_SMOOTH) != 0;
    const bool ztest = zmask != 8u;
    const unsigned c1rb = a1.x & 0x00ff00ffu, c1ga = (a1.x >> 8) & 0x00ff00ffu;
    const unsigned c2rb = a1.y & 0x00ff00ffu, c2ga = (a1.y >> 8) & 0x00ff00ffu;
    const unsigned c3rb = a1.z & 0x00ff00ffu, c3ga = (a1.z >> 8) & 0x00ff00ffu;
    const bool same_color = (a1.x == a1.y) && (a1.y == a1.z);
    /* untinted (white / grey, alpha included) smooth-shaded textured triangles: the interpolated colour is one
       scalar, see the grey_tex branches below */
    const bool grey_tex = BIG && !PHONG && TEXM != 0 && same_color && smooth && a1.x == (a1.x & 0xffu) * 0x01010101u;
    float tu1 = 0, tu2 = 0, tu3 = 0, tv1 = 0, tv2 = 0, tv3 = 0;
    const bool texturing = TEXM != 0 && (!PHONG || (flags & PFCU_ST_TEXTURE));     /* the Phong variant checks at run time */
    const bool blending = BLENDM != 0 && (!PHONG || (flags & PFCU_ST_BLEND));
    if (texturing) {
        const uint4 a2 = __ldg(reinterpret_cast<const uint4 *>(t.data + ti) + 2);
        const uint4 a3 = __ldg(reinterpret_cast<const uint4 *>(t.data + ti) + 3);
        tu1 = __uint_as_float(a2.x); tu2 = __uint_as_float(a2.y); tu3 = __uint_as_float(a2.z);
        tv1 = __uint_as_float(a3.x); tv2 = __uint_as_float(a3.y); tv3 = __uint_as_float(a3.z);
    }
    /* edge values at this lane's pixel of block (0,0) */
    const int dx0 = t.X0 + t.lx8 - b.x, dy0 = t.Y0 + t.ly4 - b.y;
    const int e1 = wadd(wadd(s.w1R, wmul(dy0, s.w1Y)), wmul(dx0, s.w1X));
    const int e2 = wadd(wadd(s.w2R, wmul(dy0, s.w2Y)), wmul(dx0, s.w2X));
    const int e3 = wadd(wadd(s.w3R, wmul(dy0, s.w3Y)), wmul(dx0, s.w3X));
    const int rxc = t.lx8 - cx0, ryc = t.ly4 - cy0;            /* lane offset from the clipped bbox corner */

    /* block ownership: 8 warps -> warp w owns block (bx,by) iff (bx + 3*by) & 7 == w;
       16 warps -> additionally even block rows belong to warps 0..7, odd rows to warps 8..15 */
    for (int by = (NW == 16) ? by0 + ((by0 ^ (t.warp >> 3)) & 1) : by0; by <= by1; by += (NW == 16) ? 2 : 1) {
        const int bx = ((t.warp & 7) - 3 * by) & 7;
        /* skipping blocks left/right of the bbox early pays for small triangles only; the per-lane
           x-range test below rejects them anyway */
        if (!BIG && (bx < bx0 || bx > bx1)) continue;
        const int bx8 = bx << 3, by4 = by << 2;
        /* byte address of tile_addr(bx8 + lx8, by4 + ly4) */
        const unsigned sa = t.sm_base + ((unsigned)by << 10) + (t.lane_rel ^ ((unsigned)bx8 << 2));
        const int w1 = wadd(wmul(bx8, s.w1X), wadd(wmul(by4, s.w1Y), e1));
        const int w2 = wadd(wmul(bx8, s.w2X), wadd(wmul(by4, s.w2Y), e2));
        const int w3 = wadd(wmul(bx8, s.w3X), wadd(wmul(by4, s.w3Y), e3));
        bool m = ((w1 | w2 | w3) > 0) && (unsigned)(bx8 + rxc) <= xspan && (unsigned)(by4 + ryc) <= yspan;
        if (!__any_sync(0xffffffffu, m)) continue;
        /* depth-failed = covered - shaded, taken at the end; a predicated add (the compiler turns the C
           form into a three-instruction select when a branch follows) */
        asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(t.covered) : "r"((unsigned)m));

        const float W1 = FM(__int2float_rn(w1), s.invSum);
        const float W2 = FM(__int2float_rn(w2), s.invSum);
        const float W3 = FM(__int2float_rn(w3), s.invSum);
        const float zsum = FA(FA(FM(z1, W1), FM(z2, W2)), FM(z3, W3));
        const float z = (BIG || t.rcp_shared) ? rcp_fast(t.sm_base, t.rcp_shift, zsum) : rcp_x86(zsum);
        if (ztest) {
            const float zb = lds_depth(sa);
            const bool pass = depth_pass_mask(z, zb, zmask);
            m = m && pass;
            if (!__any_sync(0xffffffffu, m)) continue;
        }

        /* colour (color.h:153-203) */
        Px2 frag;
        unsigned kgrey = 0;
        if (smooth) {
            const int u1 = __float2int_rn(FM(W1, 255.0f)), u2 = __float2int_rn(FM(W2, 255.0f)), u3 = __float2int_rn(FM(W3, 255.0f));
            if (grey_tex) {                         /* all four channels equal: one scalar instead of two packed words */
                const unsigned x = (unsigned)(u1 + u2 + u3) * (a1.x & 0xffu);
                kgrey = (x + (x >> 8)) >> 8;        /* x <= 65280, so this is ((x*257)>>16) <= 255 */
                frag.rb = frag.ga = 0;
            } else if (same_color) {                       /* warp-uniform: (u1+u2+u3)*c has the same lanes as u1*c+u2*c+u3*c */
                const unsigned us = (unsigned)(u1 + u2 + u3);
                unsigned x = us * c1rb, y = us * c1ga;
                x = x + ((x >> 8) & 0x00ff00ffu); y = y + ((y >> 8) & 0x00ff00ffu);
                frag.rb = (x >> 8) & 0x00ff00ffu; frag.ga = (y >> 8) & 0x00ff00ffu;
            } else {
                frag.rb = smooth_lanes(c1rb, c2rb, c3rb, u1, u2, u3);
                frag.ga = smooth_lanes(c1ga, c2ga, c3ga, u1, u2, u3);
            }
        } else {
            const float mx = max_x86(W1, max_x86(W2, W3));
            frag = px_split(((mx == W1) ? a1.x : 0u) | ((mx == W2) ? a1.y : 0u) | ((mx == W3) ? a1.z : 0u));
        }

        if (texturing) {
            float u = FA(FA(FM(tu1, W1), FM(tu2, W2)), FM(tu3, W3));
            float v = FA(FA(FM(tv1, W1), FM(tv2, W2)), FM(tv3, W3));
            if (is3d) { u = FM(u, z); v = FM(v, z); }
            /* masked-off lanes: the reference samples (0,0) for them (triangles.c:510) only to stay inside
               the texture; here every fetch is bounds-checked and their result is never stored */
            unsigned texel;
            if (TEXM == 1) {
                /* |RNE(x)| == RNE(|x|) (round-to-nearest-even is symmetric; out-of-range and NaN give
                   0x80000000 either way), and |x| is a free source modifier */
                const float fu = FM(FS(u, truncf(u)), tex.wm1), fv = FM(FS(v, truncf(v)), tex.hm1);
                const int xi = cvt_rne_x86(fabsf(fu)), yi = cvt_rne_x86(fabsf(fv));
                const unsigned off = (unsigned)yi * tex.tw + (unsigned)xi;
                texel = 0u;
                if (off < tex.total) texel = __ldg((const unsigned *)tex.base + off);
            } else texel = tex_sample(tex, st, u, v);
            if (grey_tex) {                         /* (texel_c * k) >> 8 on packed lanes: products stay below 2^16 */
                frag.rb = (((texel & 0x00ff00ffu) * kgrey) >> 8) & 0x00ff00ffu;
                frag.ga = ((((texel >> 8) & 0x00ff00ffu) * kgrey) >> 8) & 0x00ff00ffu;
            } else frag = px_mul(texel, frag);
        }

        if (PHONG) {
            if (flags & PFCU_ST_PHONG) {
                const float4 *a = reinterpret_cast<const float4 *>(t.data + ti) + 4;
                const float4 px = __ldg(a), py = __ldg(a + 1), pz = __ldg(a + 2);
                const float4 nx = __ldg(a + 3), ny = __ldg(a + 4), nz = __ldg(a + 5);
                const float Nx = FA(FA(FM(nx.x, W1), FM(nx.y, W2)), FM(nx.z, W3));
                const float Ny = FA(FA(FM(ny.x, W1), FM(ny.y, W2)), FM(ny.z, W3));
                const float Nz = FA(FA(FM(nz.x, W1), FM(nz.y, W2)), FM(nz.z, W3));
                const float Px = FA(FA(FM(px.x, W1), FM(px.y, W2)), FM(px.z, W3));
                const float Py = FA(FA(FM(py.x, W1), FM(py.y, W2)), FM(py.z, W3));
                const float Pz = FA(FA(FM(pz.x, W1), FM(pz.y, W2)), FM(pz.z, W3));
                frag = px_split(phong(px_join(frag), st, (meta >> 24) & 1u, Px, Py, Pz, Nx, Ny, Nz));
            }
        }

        if (blending) {
            const unsigned dst = lds_color(sa);
            frag = px_blend(BLENDM == 3 ? blend_mode : BLENDM, frag, dst);
        }
        if (m) {
            sts_color(sa, px_join(frag));
            sts_depth(sa, z);                       /* written even with the depth test off (Q11) */
            t.shaded++;
        }
    }
}

/* FIXED_PROG >= 0: the whole batch runs one state program (texm*4 + blendm), known at launch; only that
 * variant is instantiated, which lets the register allocator fit 4 CTAs per SM.  -1: per-triangle dispatch. */
template <bool HAS_PHONG, int NW, int FIXED_PROG, int TH>
__global__ void __launch_bounds__(NW * 32, FIXED_PROG >= 0 ? 4 : (NW == 16 ? (HAS_PHONG ? 1 : 2) : (HAS_PHONG ? 2 : 3)))
k_raster(const RasterParams p)
{
    constexpr int NT = NW * 32;
    __shared__ __align__(16) unsigned s_mem[2 * TILE_PIX + (1 << RCP_SMEM_BITS)];   /* colour | depth | RCP table */
    unsigned *const s_color = s_mem;
    float *const s_depth = reinterpret_cast<float *>(s_mem + TILE_PIX);
    unsigned *const s_rcp = s_mem + 2 * TILE_PIX;
    __shared__ unsigned s_queue[QUEUE_CAP];
    __shared__ unsigned short s_qmask[QUEUE_CAP];
    __shared__ unsigned s_wcount[NW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    /* a CTA handles a 64 x TH slice of a 64x64 tile (TH = 32 halves the work quantum when the grid would
       otherwise be only a few waves deep); ownership for the multi-GPU split stays per 64x64 tile */
    constexpr int SUB = TILE / TH;
    const unsigned tile = (p.world > 1) ? (p.rank + (blockIdx.x / SUB) * p.world) : (blockIdx.x / SUB);
    if (tile >= p.nTiles) return;
    const int tx = tile % p.tilesX, ty = tile / p.tilesX;
    TileCtx t;
    t.X0 = tx * TILE; t.Y0 = ty * TILE + (int)(blockIdx.x % SUB) * TH;
    if (t.Y0 >= p.H) return;
    t.X1 = min(t.X0 + TILE, p.W) - 1; t.Y1 = min(t.Y0 + TH, p.H) - 1;
    const int X0 = t.X0, Y0 = t.Y0, X1 = t.X1, Y1 = t.Y1;
    const bool full_tile = (X0 + TILE <= p.W) && (Y0 + TH <= p.H) && ((p.W & 3) == 0);

    const int bin = (ty >> p.bin_tshift) * p.binsX + (tx >> p.bin_tshift);
    const unsigned lbeg = p.bin_starts[bin], lend = p.bin_starts[bin + 1];
    if (lbeg == lend) return;

    /* RCPPS table: shared copy when it has <= 2^11 entries (every CPU we met), else the global one */
    t.rcp_shift = c_rcp_shift;
    t.rcp_shared = t.rcp_shift >= 23 - RCP_SMEM_BITS;
    if (t.rcp_shared) for (int k = tid; k < (1 << (23 - t.rcp_shift)); k += NT) s_rcp[k] = c_rcp_tab[k];
    {   /* one opaque register holds the shared-window address; offsets are immediates in the ld/st */
        unsigned base = (unsigned)__cvta_generic_to_shared(s_mem);
        asm volatile("mov.u32 %0, %1;" : "=r"(t.sm_base) : "r"(base));
    }
    t.lx8 = lane & 7; t.ly4 = lane >> 3;
    asm volatile("mov.u32 %0, %1;" : "=r"(t.warp) : "r"(warp));      /* opaque: not re-derived from %tid in the block loop */
    {   /* tile_addr(bx*8 + lx8, by*4 + ly4)*4 == by*1024 + (lane_rel ^ (bx << 5)): the swizzle (ly4 << 5) and the
           pixel offset occupy disjoint bits.  Opaque so that it stays in a register instead of being rebuilt
           from %tid in every block iteration. */
        unsigned rel = (unsigned)(t.ly4 * (TILE * 4 + 32) + t.lx8 * 4);
        asm volatile("mov.u32 %0, %1;" : "=r"(t.lane_rel) : "r"(rel));
    }
    t.shaded = 0; t.covered = 0; t.data = p.data;

    bool loaded = false;

    for (unsigned base = lbeg; base < lend; ) {
        /* ---- fill the queue: ordered compaction of the bin list against this tile ---- */
        unsigned qn = 0;
        while (base < lend && qn + NT <= QUEUE_CAP) {
            const unsigned k = base + tid;
            bool hit = false; unsigned ti = 0, wmask = 0;
            if (k < lend) {
                ti = __ldg(&p.bin_list[k].x);
                const int4 b = __ldg(p.bbox + ti);
                hit = b.x <= X1 && b.z - 1 >= X0 && b.y <= Y1 && b.w >= Y0 && b.x < b.z;
                if (hit) {
                    const int rx0 = max(b.x, X0), rx1 = min(b.z - 1, X1), ry0 = max(b.y, Y0), ry1 = min(b.w, Y1);
                    /* edge-function reject of the whole tile (only when int32 cannot wrap) */
                    const TriSetup s = p.setup[ti];
                    if (s.flags & TF_SAFE) {
                        const int ax0 = rx0 - b.x, ax1 = rx1 - b.x, ay0 = ry0 - b.y, ay1 = ry1 - b.y;
                        const int m1 = s.w1R + (s.w1X > 0 ? ax1 : ax0) * s.w1X + (s.w1Y > 0 ? ay1 : ay0) * s.w1Y;
                        const int m2 = s.w2R + (s.w2X > 0 ? ax1 : ax0) * s.w2X + (s.w2Y > 0 ? ay1 : ay0) * s.w2Y;
                        const int m3 = s.w3R + (s.w3X > 0 ? ax1 : ax0) * s.w3X + (s.w3Y > 0 ? ay1 : ay0) * s.w3Y;
                        if ((m1 | m2 | m3) < 0) hit = false;
                    }
                    /* which warps own an 8x4 block inside the clipped bbox?  (see shade_tri) */
                    const int bx0 = (rx0 - X0) >> 3, nbx = ((rx1 - X0) >> 3) - bx0 + 1;
                    const int by0 = (ry0 - Y0) >> 2, nby = ((ry1 - Y0) >> 2) - by0 + 1;
                    const unsigned run = nbx >= 8 ? 0xffu : ((1u << nbx) - 1u);
                    for (int j = 0; j < min(nby, 8); j++) {
                        const int sh = (bx0 + 3 * (by0 + j)) & 7;
                        const unsigned bits = ((run << sh) | (run >> (8 - sh))) & 0xffu;
                        wmask |= (NW == 16 && ((by0 + j) & 1)) ? (bits << 8) : bits;
                    }
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_wcount[warp] = __popc(bal);
            __syncthreads();
            unsigned woff = 0, total = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) { const unsigned c = s_wcount[w]; if (w < warp) woff += c; total += c; }
            if (hit) {
                const unsigned pos = qn + woff + __popc(bal & ((1u << lane) - 1u)); s_queue[pos] = ti; s_qmask[pos] = (unsigned short)wmask;
                /* pull the triangle's attribute block towards L1 now: the warps that shade it later would
                   otherwise each pay a dependent L2 round trip per queue entry */
                asm volatile("prefetch.global.L1 [%0];" :: "l"(p.data + ti));
            }
            qn += total;
            base += NT;
            __syncthreads();
        }
        if (qn == 0) continue;

        /* ---- lazy tile load: 128-bit coalesced rows into the swizzled shared tile ---- */
        if (!loaded) {
            loaded = true;
            if (full_tile) {
                for (int r = tid >> 4; r < TH; r += NT / 16) {
                    const int c4 = (tid & 15) << 2;
                    const size_t gi = (size_t)(Y0 + r) * p.W + X0 + c4;
                    const uint4 cv = __ldcs(reinterpret_cast<const uint4 *>(p.color + gi));
                    const float4 dv = __ldcs(reinterpret_cast<const float4 *>(p.depth + gi));
                    const int sa = tile_addr(c4, r);
                    *reinterpret_cast<uint4 *>(s_color + sa) = cv;
                    *reinterpret_cast<float4 *>(s_depth + sa) = dv;
                }
            } else {
                for (int k = tid; k < TILE * TH; k += NT) {
                    const int lx = k & (TILE - 1), ly = k >> 6;
                    if (X0 + lx <= X1 && Y0 + ly <= Y1) {
                        const size_t gi = (size_t)(Y0 + ly) * p.W + X0 + lx;
                        s_color[tile_addr(lx, ly)] = p.color[gi];
                        s_depth[tile_addr(lx, ly)] = p.depth[gi];
                    }
                }
            }
            __syncthreads();
        }

        /* ---- every warp walks the queue in order over the 8x4 blocks it owns ---- */
        unsigned cur_state = 0xffffffffu;
        const DevState *st = nullptr;
        unsigned flags = 0, zmask = 8u; int blend_mode = 0, prog = 0;
        TexRegs tex; tex.base = nullptr; tex.tw = tex.th = tex.total = 0; tex.wm1 = tex.hm1 = 0.0f; tex.fmt = tex.wrap = tex.filter = 0;
        for (unsigned q0 = 0; q0 < qn; q0 += 32) {
            const unsigned mk = (q0 + lane < qn) ? s_qmask[q0 + lane] : 0u;
            unsigned rel = __ballot_sync(0xffffffffu, (mk >> warp) & 1u);
            while (rel) {
                const int j = __ffs(rel) - 1; rel &= rel - 1u;
                const unsigned ti = s_queue[q0 + j];
                if (rel) {                          /* software prefetch of this warp's next entry */
                    const unsigned tn = s_queue[q0 + __ffs(rel) - 1];
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(p.bbox + tn));
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(p.setup + tn));
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(p.data + tn));
                }
                const int4 b = __ldg(p.bbox + ti);
                const TriSetup s = p.setup[ti];
                const uint4 a0 = __ldg(reinterpret_cast<const uint4 *>(p.data + ti));
                const uint4 a1 = __ldg(reinterpret_cast<const uint4 *>(p.data + ti) + 1);
                if ((a0.w & 0xffffffu) != cur_state) {
                    cur_state = a0.w & 0xffffffu;
                    st = p.states + cur_state;
                    flags = st->flags; blend_mode = st->blend_mode;
                    zmask = (flags & PFCU_ST_DEPTH_TEST) ? depth_mask(st->depth_func) : 8u;   /* 8: no test */
                    int texm = 0;
                    if (flags & PFCU_ST_TEXTURE) {
                        tex.base = st->tex; tex.tw = st->tw; tex.th = st->th; tex.total = st->tw * st->th;
                        tex.wm1 = __uint2float_rn(st->tw - 1u); tex.hm1 = __uint2float_rn(st->th - 1u);
                        tex.fmt = st->tfmt; tex.wrap = st->tex_wrap; tex.filter = st->tex_filter;
                        texm = (tex.fmt == PFCU_TEX_RGBA8 && tex.wrap == 0 && tex.filter == 0) ? 1 : 2;
                    }
                    const int blendm = !(flags & PFCU_ST_BLEND) ? 0 : (blend_mode == 1 ? 1 : (blend_mode == 2 ? 2 : 3));
                    prog = texm * 4 + blendm;
                    if (HAS_PHONG && (flags & PFCU_ST_PHONG)) prog = 12;
                }
                if (FIXED_PROG >= 0) {
                    shade_tri<FIXED_PROG / 4, FIXED_PROG % 4, false, NW, true>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex);
                    continue;
                }
                switch (prog) {
                case 0:  shade_tri<0, 0, false, NW, false>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex); break;
                case 1:  shade_tri<0, 1, false, NW, false>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex); break;
                case 2:  shade_tri<0, 2, false, NW, false>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex); break;
                case 3:  shade_tri<0, 3, false, NW, false>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex); break;
                case 4:  shade_tri<1, 0, false, NW, false>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex); break;
                case 5:  shade_tri<1, 1, false, NW, false>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex); break;
                case 6:  shade_tri<1, 2, false, NW, false>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex); break;
                case 7:  shade_tri<1, 3, false, NW, false>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex); break;
                case 8:  shade_tri<2, 0, false, NW, false>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex); break;
                case 9:  shade_tri<2, 1, false, NW, false>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex); break;
                case 10: shade_tri<2, 2, false, NW, false>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex); break;
                case 11: shade_tri<2, 3, false, NW, false>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex); break;
                default: if (HAS_PHONG) shade_tri<2, 3, true, NW, false>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex); break;
                }
            }
        }
        __syncthreads();
    }

    /* ---- write the tile back ---- */
    if (loaded) {
        if (full_tile) {
            for (int r = tid >> 4; r < TH; r += NT / 16) {
                const int c4 = (tid & 15) << 2;
                const size_t gi = (size_t)(Y0 + r) * p.W + X0 + c4;
                const int sa = tile_addr(c4, r);
                __stcs(reinterpret_cast<uint4 *>(p.color + gi), *reinterpret_cast<const uint4 *>(s_color + sa));
                __stcs(reinterpret_cast<float4 *>(p.depth + gi), *reinterpret_cast<const float4 *>(s_depth + sa));
            }
        } else {
            for (int k = tid; k < TILE * TH; k += NT) {
                const int lx = k & (TILE - 1), ly = k >> 6;
                if (X0 + lx <= X1 && Y0 + ly <= Y1) {
                    const size_t gi = (size_t)(Y0 + ly) * p.W + X0 + lx;
                    p.color[gi] = s_color[tile_addr(lx, ly)];
                    p.depth[gi] = s_depth[tile_addr(lx, ly)];
                }
            }
        }
    }
    /* counters: warp reduce, one atomic per warp */
    unsigned shaded = t.shaded, zfailed = t.covered - t.shaded;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        shaded += __shfl_down_sync(0xffffffffu, shaded, o);
        zfailed += __shfl_down_sync(0xffffffffu, zfailed, o);
    }
    if (lane == 0) {
        if (shaded) atomicAdd(p.counters + 1, (unsigned long long)shaded);
        if (zfailed) atomicAdd(p.counters + 2, (unsigned long long)zfailed);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* kernel: fragment-compacting tile rasteriser (batches of many small triangles)                    */
/* ------------------------------------------------------------------------------------------------ */
/*
 * k_raster walks ONE triangle at a time per warp over 8x4-pixel blocks; with triangles of a dozen pixels
 * most lanes of a block are uncovered and the per-triangle prologue is paid by every warp the triangle
 * touches.  k_raster_frag turns the work around: a CTA owns a 64 x (NW) slice cut into 8x8-pixel REGIONS, one
 * region per warp (fixed pixel ownership, so submission order per pixel is kept without atomics).  Each warp
 *   1. gathers, in order, up to 32 queued triangles that touch its region (lane = triangle),
 *   2. every lane loads its triangle's constants (independent loads, one memory round trip per group), moves the
 *      edge functions to the region's origin, stages everything in shared memory (FRAG_NF fields), clips the
 *      bbox to the region (n = candidate pixels; 0 when an edge function is negative over the whole clipped
 *      rectangle) and a warp scan of n lays all candidates of the 32 triangles out as one ordered fragment stream,
 *   3. the stream is consumed 32 fragments at a time (lane = fragment, usually of several triangles): owner
 *      lookup by binary search over the scan, triangle constants from the staging area, then exactly the same
 *      coverage / depth / colour / texture / Phong / blend arithmetic as shade_tri,
 *   4. fragments of one chunk that hit the same pixel (shared edges and vertices are drawn by every
 *      triangle that owns them, Q4) are ranked by __match_any_sync and written in rank order.
 * The region tiles live in shared memory as [region][8][8] with a stride of 72 words so that the 128-bit row
 * load/store of the slice is bank-conflict free.  The slice and the RCPPS table arrive by cp.async while the
 * queue is filtered; the filter reads only the packed bin-list entries (no dependent loads).
 * Launched as 64x8 slices, 8 warps: 4 CTAs per SM (64 registers), 3 with Phong (80 registers).
 */
#define FRAG_RSTRIDE 72

struct FragCtx {
    unsigned col_base;                  /* shared-window byte address of this warp's colour region     */
    unsigned rcp_base;                  /* ... of the shared RCPPS table                                */
    int rcp_shift; bool rcp_shared;
    int RX0, RY0, RX1, RY1;             /* the region on the surface, inclusive                         */
    unsigned shaded, covered;
    unsigned tri_base;                  /* shared-window byte address of this warp's triangle staging   */
};

/* Per-warp staging of a group's triangle constants: field F of the triangle held by lane l is the 16-byte slot
 * [F][l], so lanes that fetch different triangles hit different banks and lanes on the same triangle broadcast.
 *   F0 E1 E2 E3 invSum      (edge functions at the region's pixel (0,0), wrapping int32)
 *   F1 w1X w1Y w2X w2Y      F2 w3X w3Y z1 z2      F3 z3 meta c1 c2      F4 c3 u1 u2 u3      F5 v1 v2 v3 -
 *   Phong only: F6 px1..3 py1   F7 py2 py3 pz1 pz2   F8 pz3 nx1..3   F9 ny1..3 nz1   F10 nz2 nz3 - -          */
#define FRAG_NF        6
#define FRAG_NF_PHONG  11
__device__ __forceinline__ uint4 lds_tri(unsigned base, int field, int j)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + (unsigned)((field * 32 + j) << 4)));
    return v;
}
__device__ __forceinline__ void sts_tri(unsigned base, int field, int j, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(base + (unsigned)((field * 32 + j) << 4)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int OFF> __device__ __forceinline__ float lds_f32_off(unsigned addr) { float v; asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF)); return v; }
template <int OFF> __device__ __forceinline__ void sts_f32_off(unsigned addr, float v) { asm volatile("st.shared.f32 [%0+%1], %2;" :: "r"(addr), "n"(OFF), "f"(v) : "memory"); }

__device__ __forceinline__ float rcp_tab(const FragCtx &t, float x)
{
    const unsigned u = __float_as_uint(x), E = u & 0x7f800000u;
    if (!t.rcp_shared || E - 0x00800000u >= 0x7e000000u) return rcp_x86(x);
    const unsigned tv = lds_u32(t.rcp_base + (((u & 0x007fffffu) >> t.rcp_shift) << 2));
    return __uint_as_float((tv + 0x3f800000u - E) | (u & 0x80000000u));
}

/* One group of <= 32 triangles in one state (lane l holds triangle ti with nn candidate pixels in this warp's
 * region; nn == 0 for lanes outside the group).  pk = cx0 | cy0<<4 | cw<<8 | ceil(1024/cw)<<12 describes the
 * clipped rectangle (region-local).  lo is a lane that is known to hold a valid triangle. */
template <int TEXM, int BLENDM, bool PHONG, int NW>
__device__ __forceinline__ void frag_run(FragCtx &t, const unsigned nn, const unsigned pk, const int lo,
                                         const DevState *st, const unsigned flags, const unsigned zmask, const int blend_mode, const TexRegs &tex)
{
    constexpr int DEPTH_OFF = NW * FRAG_RSTRIDE * 4;
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    unsigned I = nn;                                                /* inclusive scan of the candidate counts */
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(FULL, I, o); if ((int)lane >= o) I += y; }
    const unsigned total = __shfl_sync(FULL, I, 31);
    const unsigned Ex = I - nn;
    const bool smooth = (flags & PFCU_ST_SMOOTH) != 0;
    const bool ztest = zmask != 8u;
    const bool texturing = TEXM != 0 && (!PHONG || (flags & PFCU_ST_TEXTURE));
    const bool blending = BLENDM != 0 && (!PHONG || (flags & PFCU_ST_BLEND));

    for (unsigned base = 0; base < total; base += 32) {
        const unsigned f = base + lane;
        const bool valid = f < total;
        unsigned pos = 0;                                           /* owner = number of lanes whose scan value is <= f */
#pragma unroll
        for (int step = 16; step; step >>= 1) { const unsigned v = __shfl_sync(FULL, I, pos + step - 1); if (v <= f) pos += step; }
        const int j = valid ? (int)pos : lo;
        const unsigned Ej = __shfl_sync(FULL, Ex, j), pkj = __shfl_sync(FULL, pk, j);
        const unsigned r = valid ? f - Ej : 0u;
        const unsigned cw = (pkj >> 8) & 15u;
        const unsigned ry = (r * (pkj >> 12)) >> 10, rx = r - ry * cw;          /* r / cw, r % cw (r < 64, cw <= 8: exact) */
        const int px = (int)((pkj & 15u) + rx), py = (int)(((pkj >> 4) & 15u) + ry);

        const uint4 f0 = lds_tri(t.tri_base, 0, j), f1 = lds_tri(t.tri_base, 1, j), f2 = lds_tri(t.tri_base, 2, j);
        const int w1 = wadd(wadd((int)f0.x, wmul(py, (int)f1.y)), wmul(px, (int)f1.x));
        const int w2 = wadd(wadd((int)f0.y, wmul(py, (int)f1.w)), wmul(px, (int)f1.z));
        const int w3 = wadd(wadd((int)f0.z, wmul(py, (int)f2.y)), wmul(px, (int)f2.x));
        bool m = valid && ((w1 | w2 | w3) > 0);
        if (!__any_sync(FULL, m)) continue;
        t.covered += m ? 1u : 0u;

        const uint4 f3 = lds_tri(t.tri_base, 3, j);
        const unsigned meta = f3.y;
        const float invSum = __uint_as_float(f0.w);
        const float W1 = FM(__int2float_rn(w1), invSum);
        const float W2 = FM(__int2float_rn(w2), invSum);
        const float W3 = FM(__int2float_rn(w3), invSum);
        const float zsum = FA(FA(FM(__uint_as_float(f2.z), W1), FM(__uint_as_float(f2.w), W2)), FM(__uint_as_float(f3.x), W3));
        const float z = rcp_tab(t, zsum);

        /* same-pixel fragments of this chunk (different triangles) must be applied in triangle order */
        const unsigned sa = t.col_base + (unsigned)(((py << 3) + px) << 2);
        const unsigned peers = __match_any_sync(FULL, m ? sa : (0x80000000u | lane));
        const unsigned rank = __popc(peers & lt);
        const unsigned nr = __reduce_max_sync(FULL, m ? rank : 0u);
        if (ztest && nr == 0u) {                        /* no conflicts: test before shading, like the reference's early mask */
            const float zb = lds_f32_off<DEPTH_OFF>(sa);
            m = m && depth_pass_mask(z, zb, zmask);
            if (!__any_sync(FULL, m)) continue;
        }

        /* colour (color.h:153-203) */
        const uint4 f4 = lds_tri(t.tri_base, 4, j);
        const unsigned c1 = f3.z, c2 = f3.w, c3 = f4.x;
        Px2 frag;
        if (smooth) {
            const int u1 = __float2int_rn(FM(W1, 255.0f)), u2 = __float2int_rn(FM(W2, 255.0f)), u3 = __float2int_rn(FM(W3, 255.0f));
            frag.rb = smooth_lanes(c1 & 0x00ff00ffu, c2 & 0x00ff00ffu, c3 & 0x00ff00ffu, u1, u2, u3);
            frag.ga = smooth_lanes((c1 >> 8) & 0x00ff00ffu, (c2 >> 8) & 0x00ff00ffu, (c3 >> 8) & 0x00ff00ffu, u1, u2, u3);
        } else {
            const float mx = max_x86(W1, max_x86(W2, W3));
            frag = px_split(((mx == W1) ? c1 : 0u) | ((mx == W2) ? c2 : 0u) | ((mx == W3) ? c3 : 0u));
        }

        if (texturing) {
            const uint4 f5 = lds_tri(t.tri_base, 5, j);
            float u = FA(FA(FM(__uint_as_float(f4.y), W1), FM(__uint_as_float(f4.z), W2)), FM(__uint_as_float(f4.w), W3));
            float v = FA(FA(FM(__uint_as_float(f5.x), W1), FM(__uint_as_float(f5.y), W2)), FM(__uint_as_float(f5.z), W3));
            if ((meta >> 25) & 1u) { u = FM(u, z); v = FM(v, z); }
            unsigned texel;
            if (TEXM == 1) {
                const float fu = FM(FS(u, truncf(u)), tex.wm1), fv = FM(FS(v, truncf(v)), tex.hm1);
                const int xi = cvt_rne_x86(fabsf(fu)), yi = cvt_rne_x86(fabsf(fv));
                const unsigned off = (unsigned)yi * tex.tw + (unsigned)xi;
                texel = 0u;
                if (off < tex.total) texel = __ldg((const unsigned *)tex.base + off);
            } else texel = tex_sample(tex, st, u, v);
            frag = px_mul(texel, frag);
        }

        if (PHONG) {
            if (flags & PFCU_ST_PHONG) {
#define UF(x) __uint_as_float(x)
                const uint4 g6 = lds_tri(t.tri_base, 6, j), g7 = lds_tri(t.tri_base, 7, j), g8 = lds_tri(t.tri_base, 8, j);
                const uint4 g9 = lds_tri(t.tri_base, 9, j), g10 = lds_tri(t.tri_base, 10, j);
                const float Qx = FA(FA(FM(UF(g6.x), W1), FM(UF(g6.y), W2)), FM(UF(g6.z), W3));
                const float Qy = FA(FA(FM(UF(g6.w), W1), FM(UF(g7.x), W2)), FM(UF(g7.y), W3));
                const float Qz = FA(FA(FM(UF(g7.z), W1), FM(UF(g7.w), W2)), FM(UF(g8.x), W3));
                const float Nx = FA(FA(FM(UF(g8.y), W1), FM(UF(g8.z), W2)), FM(UF(g8.w), W3));
                const float Ny = FA(FA(FM(UF(g9.x), W1), FM(UF(g9.y), W2)), FM(UF(g9.z), W3));
                const float Nz = FA(FA(FM(UF(g9.w), W1), FM(UF(g10.x), W2)), FM(UF(g10.y), W3));
#undef UF
                frag = px_split(phong(px_join(frag), st, (meta >> 24) & 1u, Qx, Qy, Qz, Nx, Ny, Nz));
            }
        }

        /* ordered read-modify-write: round k applies the k-th fragment of every pixel */
        for (unsigned k = 0; k <= nr; k++) {
            if (m && rank == k) {
                bool ok = true;
                if (ztest && nr != 0u) ok = depth_pass_mask(z, lds_f32_off<DEPTH_OFF>(sa), zmask);
                if (ok) {
                    Px2 o = frag;
                    if (blending) o = px_blend(BLENDM == 3 ? blend_mode : BLENDM, frag, lds_color(sa));
                    sts_color(sa, px_join(o));
                    sts_f32_off<DEPTH_OFF>(sa, z);          /* written even with the depth test off (Q11) */
                    t.shaded++;
                }
            }
            if (nr != 0u) __syncwarp();
        }
    }
}

template <bool HAS_PHONG, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
k_raster_frag(const RasterParams p)
{
    constexpr int NT = NW * 32, TH = NW, SUB = TILE / TH;
    __shared__ __align__(16) unsigned s_tile[2 * NW * FRAG_RSTRIDE];     /* colour regions, then depth regions */
    unsigned *const s_col = s_tile;
    float *const s_dep = reinterpret_cast<float *>(s_tile + NW * FRAG_RSTRIDE);
    __shared__ unsigned s_rcp[1 << RCP_SMEM_BITS];
    __shared__ unsigned s_queue[QUEUE_CAP];
    __shared__ unsigned short s_qmask[QUEUE_CAP];
    __shared__ unsigned s_wcount[NW];
    __shared__ unsigned s_group[NW][32];
    extern __shared__ __align__(16) uint4 s_tri[];          /* [NW][NF][32] triangle staging, see FRAG_NF */
    constexpr int NF = HAS_PHONG ? FRAG_NF_PHONG : FRAG_NF;
    static_assert(NW == 8 || NW == 16, "one 8x8 region per warp, 8 regions per row");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tile = (p.world > 1) ? (p.rank + (blockIdx.x / SUB) * p.world) : (blockIdx.x / SUB);
    if (tile >= p.nTiles) return;
    const int tx = tile % p.tilesX, ty = tile / p.tilesX;
    const int X0 = tx * TILE, Y0 = ty * TILE + (int)(blockIdx.x % SUB) * TH;
    if (Y0 >= p.H) return;
    const int X1 = min(X0 + TILE, p.W) - 1, Y1 = min(Y0 + TH, p.H) - 1;
    const bool full_tile = (X0 + TILE <= p.W) && (Y0 + TH <= p.H) && ((p.W & 3) == 0);

    const int bin = (ty >> p.bin_tshift) * p.binsX + (tx >> p.bin_tshift);
    const unsigned lbeg = p.bin_starts[bin], lend = p.bin_starts[bin + 1];
    if (lbeg == lend) return;

    FragCtx t;
    t.rcp_shift = c_rcp_shift;
    t.rcp_shared = t.rcp_shift >= 23 - RCP_SMEM_BITS;
    /* the RCPPS table and (full slices) the colour/depth slice arrive asynchronously while the queue is filled */
    if (t.rcp_shared) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(s_rcp);
        for (int k = tid; k < (1 << (21 - t.rcp_shift)); k += NT)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst + (unsigned)(k << 4)), "l"(c_rcp_tab + 4 * k) : "memory");
    }
    if (full_tile) {
        const unsigned dcol = (unsigned)__cvta_generic_to_shared(s_col);
        for (int k = tid; k < TH * 16; k += NT) {
            const int r = k >> 4, c4 = (k & 15) << 2;
            const size_t gi = (size_t)(Y0 + r) * p.W + X0 + c4;
            const unsigned sa = (unsigned)((((r >> 3) * 8 + (c4 >> 3)) * FRAG_RSTRIDE + (r & 7) * 8 + (c4 & 7)) << 2);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dcol + sa), "l"(p.color + gi) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dcol + sa + NW * FRAG_RSTRIDE * 4), "l"(p.depth + gi) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    t.col_base = (unsigned)__cvta_generic_to_shared(s_col + warp * FRAG_RSTRIDE);
    t.rcp_base = (unsigned)__cvta_generic_to_shared(s_rcp);
    t.RX0 = X0 + (warp & 7) * 8; t.RY0 = Y0 + (warp >> 3) * 8;
    t.RX1 = min(t.RX0 + 7, X1); t.RY1 = min(t.RY0 + 7, Y1);
    t.shaded = 0; t.covered = 0;
    t.tri_base = (unsigned)__cvta_generic_to_shared(s_tri + warp * NF * 32);

    bool loaded = false;
    /* the slice relative to its bin, as the bin-list entries store their rectangles */
    const int bx0 = X0 - (((tx >> p.bin_tshift) << p.bin_tshift) * TILE), by0 = Y0 - (((ty >> p.bin_tshift) << p.bin_tshift) * TILE);
    const int bx1 = bx0 + (X1 - X0), by1 = by0 + (Y1 - Y0);

    for (unsigned base = lbeg; base < lend; ) {
        /* ---- fill the queue: ordered compaction of the bin list against this slice ---- */
        unsigned qn = 0;
        while (base < lend && qn + NT <= QUEUE_CAP) {
            const unsigned k = base + tid;
            bool hit = false; unsigned ti = 0, wmask = 0;
            if (k < lend) {
                const uint2 e = __ldg(p.bin_list + k);
                ti = e.x;
                const int ex0 = (int)(e.y & 255u), ey0 = (int)((e.y >> 8) & 255u), ex1 = (int)((e.y >> 16) & 255u), ey1 = (int)(e.y >> 24);
                hit = ex0 <= bx1 && ex1 >= bx0 && ey0 <= by1 && ey1 >= by0;
                if (hit) {
                    /* regions (= warps) touched by the clipped rectangle; the edge-function reject is done per
                       region when the group is staged */
                    const int gx0 = (max(ex0, bx0) - bx0) >> 3, gx1 = (min(ex1, bx1) - bx0) >> 3;
                    const int gy0 = (max(ey0, by0) - by0) >> 3, gy1 = (min(ey1, by1) - by0) >> 3;
                    const unsigned run = ((2u << gx1) - 1u) & ~((1u << gx0) - 1u);
                    wmask = (gy0 == 0 ? run : 0u) | ((NW == 16 && gy1 == 1) ? (run << 8) : 0u);
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_wcount[warp] = __popc(bal);
            __syncthreads();
            unsigned woff = 0, total = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) { const unsigned c = s_wcount[w]; if (w < warp) woff += c; total += c; }
            if (hit) {
                const unsigned pos = qn + woff + __popc(bal & ((1u << lane) - 1u)); s_queue[pos] = ti; s_qmask[pos] = (unsigned short)wmask;
                asm volatile("prefetch.global.L2 [%0];" :: "l"(p.data + ti));
            }
            qn += total;
            base += NT;
            __syncthreads();
        }
        if (qn == 0) continue;

        /* ---- the slice: cp.async issued at kernel start (full slices), or a bounds-checked load now ---- */
        if (!loaded) {
            loaded = true;
            if (!full_tile) {
                for (int k = tid; k < TILE * TH; k += NT) {
                    const int lx = k & (TILE - 1), ly = k >> 6;
                    if (X0 + lx <= X1 && Y0 + ly <= Y1) {
                        const size_t gi = (size_t)(Y0 + ly) * p.W + X0 + lx;
                        const int sa = ((ly >> 3) * 8 + (lx >> 3)) * FRAG_RSTRIDE + (ly & 7) * 8 + (lx & 7);
                        s_col[sa] = p.color[gi];
                        s_dep[sa] = p.depth[gi];
                    }
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
        }

        /* ---- every warp gathers the queue entries of its region, 32 at a time, and runs them ---- */
        unsigned cur_state = 0xffffffffu;
        const DevState *st = nullptr;
        unsigned flags = 0, zmask = 8u; int blend_mode = 0, prog = 0;
        TexRegs tex; tex.base = nullptr; tex.tw = tex.th = tex.total = 0; tex.wm1 = tex.hm1 = 0.0f; tex.fmt = tex.wrap = tex.filter = 0;
        unsigned cnt = 0;
        const unsigned ltm = (1u << lane) - 1u;
        for (unsigned q0 = 0; q0 < qn; q0 += 32) {
            bool mine = false; unsigned qti = 0;
            if (q0 + lane < qn) { mine = (s_qmask[q0 + lane] >> warp) & 1u; qti = s_queue[q0 + lane]; }
            unsigned rel = __ballot_sync(0xffffffffu, mine);
            const bool last = q0 + 32 >= qn;
            do {
                if (rel) {
                    const unsigned slot = cnt + __popc(rel & ltm);
                    const bool take = mine && slot < 32u;
                    if (take) { s_group[warp][slot] = qti; mine = false; }
                    const unsigned tk = __ballot_sync(0xffffffffu, take);
                    cnt += __popc(tk); rel &= ~tk;
                }
                if (cnt == 32u || (last && rel == 0u && cnt)) {
                    /* ---- run one group ---- */
                    __syncwarp();
                    const bool have = (unsigned)lane < cnt;
                    const unsigned ti = have ? s_group[warp][lane] : 0u;
                    __syncwarp();
                    unsigned state = 0xffffffffu, nn0 = 0, pk = 0;
                    if (have) {
                        /* stage this triangle's constants (every load is independent: one memory round trip per group) */
                        const int4 b = __ldg(p.bbox + ti);
                        const uint4 s0 = __ldg(reinterpret_cast<const uint4 *>(p.setup + ti));
                        const uint4 s1 = __ldg(reinterpret_cast<const uint4 *>(p.setup + ti) + 1);
                        const uint4 s2 = __ldg(reinterpret_cast<const uint4 *>(p.setup + ti) + 2);
                        const uint4 *da = reinterpret_cast<const uint4 *>(p.data + ti);
                        const uint4 a0 = __ldg(da), a1 = __ldg(da + 1), a2 = __ldg(da + 2), a3 = __ldg(da + 3);
                        state = a0.w & 0xffffffu;
                        const int ox = wsub(t.RX0, b.x), oy = wsub(t.RY0, b.y);
                        const unsigned E1 = (unsigned)wadd(wadd((int)s0.x, wmul(oy, (int)s1.y)), wmul(ox, (int)s1.x));
                        const unsigned E2 = (unsigned)wadd(wadd((int)s0.y, wmul(oy, (int)s1.w)), wmul(ox, (int)s1.z));
                        const unsigned E3 = (unsigned)wadd(wadd((int)s0.z, wmul(oy, (int)s2.y)), wmul(ox, (int)s2.x));
                        sts_tri(t.tri_base, 0, lane, make_uint4(E1, E2, E3, s0.w));
                        sts_tri(t.tri_base, 1, lane, s1);
                        sts_tri(t.tri_base, 2, lane, make_uint4(s2.x, s2.y, a0.x, a0.y));
                        sts_tri(t.tri_base, 3, lane, make_uint4(a0.z, a0.w, a1.x, a1.y));
                        sts_tri(t.tri_base, 4, lane, make_uint4(a1.z, a2.x, a2.y, a2.z));
                        sts_tri(t.tri_base, 5, lane, make_uint4(a3.x, a3.y, a3.z, 0u));
                        if (HAS_PHONG) {
                            const uint4 qx = __ldg(da + 4), qy = __ldg(da + 5), qz = __ldg(da + 6);
                            const uint4 nx = __ldg(da + 7), ny = __ldg(da + 8), nz = __ldg(da + 9);
                            sts_tri(t.tri_base, 6, lane, make_uint4(qx.x, qx.y, qx.z, qy.x));
                            sts_tri(t.tri_base, 7, lane, make_uint4(qy.y, qy.z, qz.x, qz.y));
                            sts_tri(t.tri_base, 8, lane, make_uint4(qz.z, nx.x, nx.y, nx.z));
                            sts_tri(t.tri_base, 9, lane, make_uint4(ny.x, ny.y, ny.z, nz.x));
                            sts_tri(t.tri_base, 10, lane, make_uint4(nz.y, nz.z, 0u, 0u));
                        }
                        const int cx0 = max(b.x, t.RX0) - t.RX0, cx1 = min(b.z - 1, t.RX1) - t.RX0;
                        const int cy0 = max(b.y, t.RY0) - t.RY0, cy1 = min(b.w, t.RY1) - t.RY0;
                        const int cw = cx1 - cx0 + 1, ch = cy1 - cy0 + 1;
                        if (cw > 0 && ch > 0) {
                            nn0 = (unsigned)(cw * ch);
                            pk = (unsigned)cx0 | ((unsigned)cy0 << 4) | ((unsigned)cw << 8) | (((1024u + (unsigned)cw - 1u) / (unsigned)cw) << 12);
                            if (s2.z & TF_SAFE) {       /* an edge function negative over the whole clipped rectangle: nothing to shade */
                                /* evaluated mod 2^32 from the region origin; the corner itself lies inside the bbox, where
                                   TF_SAFE guarantees the true value fits */
                                const int m1 = wadd((int)E1, wadd(wmul(((int)s1.x > 0) ? cx1 : cx0, (int)s1.x), wmul(((int)s1.y > 0) ? cy1 : cy0, (int)s1.y)));
                                const int m2 = wadd((int)E2, wadd(wmul(((int)s1.z > 0) ? cx1 : cx0, (int)s1.z), wmul(((int)s1.w > 0) ? cy1 : cy0, (int)s1.w)));
                                const int m3 = wadd((int)E3, wadd(wmul(((int)s2.x > 0) ? cx1 : cx0, (int)s2.x), wmul(((int)s2.y > 0) ? cy1 : cy0, (int)s2.y)));
                                if ((m1 | m2 | m3) < 0) nn0 = 0;
                            }
                        }
                    }
                    __syncwarp();
                    unsigned lo = 0;
                    while (lo < cnt) {
                        const unsigned sid = __shfl_sync(0xffffffffu, state, (int)lo);
                        const unsigned diff = __ballot_sync(0xffffffffu, have && (unsigned)lane >= lo && state != sid);
                        const unsigned hi = diff ? (unsigned)(__ffs(diff) - 1) : cnt;
                        const unsigned nn = ((unsigned)lane >= lo && (unsigned)lane < hi) ? nn0 : 0u;
                        if (sid != cur_state) {
                            cur_state = sid;
                            st = p.states + cur_state;
                            flags = st->flags; blend_mode = st->blend_mode;
                            zmask = (flags & PFCU_ST_DEPTH_TEST) ? depth_mask(st->depth_func) : 8u;
                            int texm = 0;
                            if (flags & PFCU_ST_TEXTURE) {
                                tex.base = st->tex; tex.tw = st->tw; tex.th = st->th; tex.total = st->tw * st->th;
                                tex.wm1 = __uint2float_rn(st->tw - 1u); tex.hm1 = __uint2float_rn(st->th - 1u);
                                tex.fmt = st->tfmt; tex.wrap = st->tex_wrap; tex.filter = st->tex_filter;
                                texm = (tex.fmt == PFCU_TEX_RGBA8 && tex.wrap == 0 && tex.filter == 0) ? 1 : 2;
                            }
                            const int blendm = !(flags & PFCU_ST_BLEND) ? 0 : (blend_mode == 1 ? 1 : (blend_mode == 2 ? 2 : 3));
                            prog = texm * 4 + blendm;
                            if (HAS_PHONG && (flags & PFCU_ST_PHONG)) prog = 12;
                        }
                        switch (prog) {
                        case 0:  frag_run<0, 0, false, NW>(t, nn, pk, (int)lo, st, flags, zmask, blend_mode, tex); break;
                        case 1:  frag_run<0, 1, false, NW>(t, nn, pk, (int)lo, st, flags, zmask, blend_mode, tex); break;
                        case 2:  frag_run<0, 2, false, NW>(t, nn, pk, (int)lo, st, flags, zmask, blend_mode, tex); break;
                        case 3:  frag_run<0, 3, false, NW>(t, nn, pk, (int)lo, st, flags, zmask, blend_mode, tex); break;
                        case 4:  frag_run<1, 0, false, NW>(t, nn, pk, (int)lo, st, flags, zmask, blend_mode, tex); break;
                        case 5:  frag_run<1, 1, false, NW>(t, nn, pk, (int)lo, st, flags, zmask, blend_mode, tex); break;
                        case 6:  frag_run<1, 2, false, NW>(t, nn, pk, (int)lo, st, flags, zmask, blend_mode, tex); break;
                        case 7:  frag_run<1, 3, false, NW>(t, nn, pk, (int)lo, st, flags, zmask, blend_mode, tex); break;
                        case 8:  frag_run<2, 0, false, NW>(t, nn, pk, (int)lo, st, flags, zmask, blend_mode, tex); break;
                        case 9:  frag_run<2, 1, false, NW>(t, nn, pk, (int)lo, st, flags, zmask, blend_mode, tex); break;
                        case 10: frag_run<2, 2, false, NW>(t, nn, pk, (int)lo, st, flags, zmask, blend_mode, tex); break;
                        case 11: frag_run<2, 3, false, NW>(t, nn, pk, (int)lo, st, flags, zmask, blend_mode, tex); break;
                        default: if (HAS_PHONG) frag_run<2, 3, true, NW>(t, nn, pk, (int)lo, st, flags, zmask, blend_mode, tex); break;
                        }
                        lo = hi;
                    }
                    cnt = 0;
                }
            } while (rel);
        }
        __syncthreads();
    }

    asm volatile("cp.async.wait_all;" ::: "memory");        /* nothing may be in flight when the CTA retires */
    /* ---- write the slice back ---- */
    if (loaded) {
        if (full_tile) {
            for (int k = tid; k < TH * 16; k += NT) {
                const int r = k >> 4, c4 = (k & 15) << 2;
                const size_t gi = (size_t)(Y0 + r) * p.W + X0 + c4;
                const int sa = ((r >> 3) * 8 + (c4 >> 3)) * FRAG_RSTRIDE + (r & 7) * 8 + (c4 & 7);
                __stcs(reinterpret_cast<uint4 *>(p.color + gi), *reinterpret_cast<const uint4 *>(s_col + sa));
                __stcs(reinterpret_cast<float4 *>(p.depth + gi), *reinterpret_cast<const float4 *>(s_dep + sa));
            }
        } else {
            for (int k = tid; k < TILE * TH; k += NT) {
                const int lx = k & (TILE - 1), ly = k >> 6;
                if (X0 + lx <= X1 && Y0 + ly <= Y1) {
                    const size_t gi = (size_t)(Y0 + ly) * p.W + X0 + lx;
                    const int sa = ((ly >> 3) * 8 + (lx >> 3)) * FRAG_RSTRIDE + (ly & 7) * 8 + (lx & 7);
                    p.color[gi] = s_col[sa];
                    p.depth[gi] = s_dep[sa];
                }
            }
        }
    }
    unsigned shaded = t.shaded, zfailed = t.covered - t.shaded;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        shaded += __shfl_down_sync(0xffffffffu, shaded, o);
        zfailed += __shfl_down_sync(0xffffffffu, zfailed, o);
    }
    if (lane == 0) {
        if (shaded) atomicAdd(p.counters + 1, (unsigned long long)shaded);
        if (zfailed) atomicAdd(p.counters + 2, (unsigned long long)zfailed);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* ------------------------------------------------------------------------------------------------ */
/* kernels: device vertex stage (pf_vstage.h compiled as device code)                               */
/* ------------------------------------------------------------------------------------------------ */

struct VtxArgs {
    const float *pos; int pos_size; const float *nrm; const float *uv; const unsigned char *col; int col_size;
    const void *idx; int idx_bytes; unsigned first, n_tri; unsigned cur_color; int n_faces; int face[2];
    unsigned state;
};

__device__ __forceinline__ unsigned vtx_index(const VtxArgs &a, unsigned k)
{
    if (!a.idx) return a.first + k;
    if (a.idx_bytes == 4) return __ldg((const unsigned *)a.idx + k);
    if (a.idx_bytes == 2) return __ldg((const unsigned short *)a.idx + k);
    return __ldg((const unsigned char *)a.idx + k);
}

/* vertex fetch with the reference's defaults for absent arrays (context.c:1253-1395) */
__device__ __forceinline__ void vtx_load(const VtxArgs &a, unsigned vi, pfv_vertex *v)
{
    v->position[0] = 0.0f; v->position[1] = 0.0f; v->position[2] = 0.0f; v->position[3] = 1.0f;
    for (int k = 0; k < a.pos_size; k++) v->position[k] = __ldg(a.pos + (size_t)vi * a.pos_size + k);
    for (int k = 0; k < 3; k++) v->normal[k] = a.nrm ? __ldg(a.nrm + (size_t)vi * 3 + k) : 0.0f;
    for (int k = 0; k < 2; k++) v->texcoord[k] = a.uv ? __ldg(a.uv + (size_t)vi * 2 + k) : 0.0f;
    unsigned c = a.cur_color;
    if (a.col) {
        c = 0xffffffffu;
        for (int k = 0; k < a.col_size; k++) c = (c & ~(255u << (8 * k))) | ((unsigned)__ldg(a.col + (size_t)vi * a.col_size + k) << (8 * k));
    }
    v->color = c;
    v->screen[0] = 0.0f; v->screen[1] = 0.0f;
}

/* runs the whole vertex stage for item (triangle, face pass); returns the number of output triangles */
__device__ __forceinline__ int vtx_process(const VtxArgs &a, const pfv_params &vp, unsigned item, pfv_vertex *poly, int *is3d, int *face_out)
{
    const unsigned tri = item / (unsigned)a.n_faces;
    const int face = a.face[item % (unsigned)a.n_faces];
    *face_out = face;
    for (int k = 0; k < 3; k++) {
        vtx_load(a, vtx_index(a, tri * 3u + k), &poly[k]);
        if (vp.lighting) pfv_prologue(&vp, face, &poly[k]);
    }
    int n = 3;
    *is3d = pfv_project_and_clip(&vp, poly, &n);
    return n >= 3 ? n - 2 : 0;
}

__global__ void __launch_bounds__(128)
k_vertex_count(const VtxArgs a, const pfv_params vp, unsigned n_items, unsigned *__restrict__ counts)
{
    const unsigned item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_items) return;
    pfv_vertex poly[PFV_MAX_POLY];
    int is3d, face;
    counts[item] = (unsigned)vtx_process(a, vp, item, poly, &is3d, &face);
}

__global__ void __launch_bounds__(128)
k_vertex_emit(const VtxArgs a, const pfv_params vp, unsigned n_items, const unsigned *__restrict__ offsets, pfcu_triangle *__restrict__ out)
{
    const unsigned item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_items) return;
    pfv_vertex poly[PFV_MAX_POLY];
    int is3d, face;
    const int n = vtx_process(a, vp, item, poly, &is3d, &face);
    pfcu_triangle *dst = out + offsets[item];
    for (int i = 0; i < n; i++) pfv_emit(dst + i, &poly[0], &poly[i + 1], &poly[i + 2], a.state, face, is3d);
}

/* ---- raw triangles (immediate mode, render lists): the whole per-triangle prologue on the device ---- */
struct RawArgs { const pfcu_rawtri *tris; const pfcu_vparams_lit *vp; const float *pow_tables; unsigned n; };

__device__ __forceinline__ int raw_process(const RawArgs &a, unsigned i, pfv_vertex *poly, int *is3d, int *face_out, unsigned *state)
{
    const pfcu_rawtri *t = a.tris + i;
    const pfcu_vparams_lit *e = a.vp + t->vparams;
    const int face = t->face;
    *face_out = face; *state = t->state;
    for (int k = 0; k < 3; k++) {
        const pfcu_rawvertex *r = &t->v[k];
        pfv_vertex *v = &poly[k];
        for (int j = 0; j < 4; j++) v->position[j] = r->pos[j];
        for (int j = 0; j < 3; j++) v->normal[j] = r->normal[j];
        v->texcoord[0] = r->uv[0]; v->texcoord[1] = r->uv[1];
        v->color = r->rgba;
        v->screen[0] = 0.0f; v->screen[1] = 0.0f;
        for (int j = 0; j < 4; j++) v->homogeneous[j] = 0.0f;
        if (e->base.lighting) pfv_prologue_lit(e, a.pow_tables, face, v);
    }
    int n = 3;
    *is3d = pfv_project_and_clip(&e->base, poly, &n);
    return n >= 3 ? n - 2 : 0;
}

__global__ void __launch_bounds__(128)
k_raw_count(const RawArgs a, unsigned *__restrict__ counts)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    pfv_vertex poly[PFV_MAX_POLY];
    int is3d, face; unsigned state;
    counts[i] = (unsigned)raw_process(a, i, poly, &is3d, &face, &state);
}

__global__ void __launch_bounds__(128)
k_raw_emit(const RawArgs a, const unsigned *__restrict__ offsets, pfcu_triangle *__restrict__ out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    pfv_vertex poly[PFV_MAX_POLY];
    int is3d, face; unsigned state;
    const int n = raw_process(a, i, poly, &is3d, &face, &state);
    pfcu_triangle *dst = out + offsets[i];
    for (int k = 0; k < n; k++) pfv_emit(dst + k, &poly[0], &poly[k + 1], &poly[k + 2], state, face, is3d);
}

/* ---- points and lines (pf_prims.h) ------------------------------------------------------------------
 * One CTA per 64x64 tile; every CTA walks ALL primitives in submission order and applies the fragments that
 * fall into its tile (threads = steps of one plain line / cells of one point), with a barrier between plain
 * lines.  Order per pixel = submission order; no inter-CTA communication.  Primitives whose rectangle cannot
 * touch the tile are skipped (only when every x of the line is inside the surface, because out-of-range columns
 * wrap into the neighbouring rows like upstream). */
struct PrimParams { const pfcu_prim *prims; unsigned n; uint32_t *color; float *depth; unsigned W, H; int tilesX; unsigned rank, world, nTiles; };

__device__ __forceinline__ void prim_pixel(const PrimParams &p, const pfcu_prim &pr, int X0, int Y0, uint32_t off, float z, uint32_t color, bool test)
{
    if (off >= p.W * p.H) return;
    const int x = (int)(off % p.W), y = (int)(off / p.W);
    if (x < X0 || x >= X0 + TILE || y < Y0 || y >= Y0 + TILE) return;
    if (test && !pfp_depth(pr.depth_func, z, p.depth[off])) return;
    p.color[off] = (pr.flags & PFCU_ST_BLEND) ? pfp_blend(pr.blend_mode, color, p.color[off]) : color;
    p.depth[off] = z;
}

__global__ void __launch_bounds__(256)
k_prims(const PrimParams p)
{
    const unsigned tile = (p.world > 1) ? (p.rank + blockIdx.x * p.world) : blockIdx.x;
    if (tile >= p.nTiles) return;
    const int X0 = (int)(tile % (unsigned)p.tilesX) * TILE, Y0 = (int)(tile / (unsigned)p.tilesX) * TILE;
    for (unsigned i = 0; i < p.n; i++) {
        const pfcu_prim pr = p.prims[i];
        const bool ztest = (pr.flags & PFCU_ST_DEPTH_TEST) != 0;
        if (pr.kind == PFP_KIND_POINT) {
            const int cx = PFV_F2I(pr.x1), cy = PFV_F2I(pr.y1);
            if (pr.size <= 1.0f) {
                if (threadIdx.x == 0) prim_pixel(p, pr, X0, Y0, (uint32_t)cy * p.W + (uint32_t)cx, pr.z1, pr.c1, ztest);
            } else {
                const float r = __fmul_rn(pr.size, 0.5f), r2 = __fmul_rn(r, r);
                const int R = PFV_F2I(r);
                if (R >= 0 && R < 16384 && !(cx + R < X0 || cx - R >= X0 + TILE || cy + R < Y0 || cy - R >= Y0 + TILE)) {
                    const int side = 2 * R + 1;
                    for (int c = threadIdx.x; c < side * side; c += 256) {
                        const int y = c / side - R, x = c % side - R;
                        if (__int2float_rn(y * y + x * x) <= r2) {
                            const uint32_t px = (uint32_t)(cx + x), py = (uint32_t)(cy + y);
                            if (px < p.W && py < p.H) prim_pixel(p, pr, X0, Y0, py * p.W + px, pr.z1, pr.c1, ztest);
                        }
                    }
                }
            }
            __syncthreads();
            continue;
        }
        int axis;
        const unsigned nsub = pfp_thick_count(pr.x1, pr.y1, pr.x2, pr.y2, pr.size, &axis);
        const bool thick = pr.size > 1.5f;
        /* conservative reject: all columns inside the surface (no wrapping) and the rectangle, widened by the
           thickness, misses the tile */
        {
            const int x1 = PFV_F2I(pr.x1), y1 = PFV_F2I(pr.y1), x2 = PFV_F2I(pr.x2), y2 = PFV_F2I(pr.y2);
            const int wd = (int)(nsub >> 1) + 1;
            const int xa = min(x1, x2) - wd, xb = max(x1, x2) + wd, ya = min(y1, y2) - wd, yb = max(y1, y2) + wd;
            if (xa >= 0 && xb < (int)p.W && (xb < X0 || xa >= X0 + TILE || yb < Y0 || ya >= Y0 + TILE)) continue;
        }
        for (unsigned sub = 0; sub < nsub; sub++) {
            const float sh = pfp_thick_shift(sub);
            pfp_line L;
            pfp_line_setup(&L, axis ? pr.x1 : __fadd_rn(pr.x1, sh), axis ? __fadd_rn(pr.y1, sh) : pr.y1,
                           axis ? pr.x2 : __fadd_rn(pr.x2, sh), axis ? __fadd_rn(pr.y2, sh) : pr.y2);
            const bool test = ztest || (thick && sub == 0);
            const unsigned steps = pfp_line_steps(&L);
            for (unsigned k = threadIdx.x; k < steps; k += 256) {
                float t;
                const uint32_t off = pfp_line_step(&L, k, p.W, &t);
                prim_pixel(p, pr, X0, Y0, off, __fadd_rn(pr.z1, __fmul_rn(t, __fsub_rn(pr.z2, pr.z1))), pfp_color_lerp(pr.c1, pr.c2, t), test);
            }
            __syncthreads();
        }
    }
}

/* Raw batches of at most 1024 triangles: count, scan and emission in ONE single-CTA kernel; the number of output
 * triangles (at most 10 per input after clipping) stays on the device: *d_total feeds k_front_small, so the host
 * never waits. */
__global__ void __launch_bounds__(1024)
k_raw_small(const RawArgs a, pfcu_triangle *__restrict__ out, unsigned *__restrict__ d_total, unsigned long long *__restrict__ counters)
{
    __shared__ unsigned s_warp[32];
    const unsigned i = threadIdx.x, lane = i & 31u, warp = i >> 5;
    pfv_vertex poly[PFV_MAX_POLY];
    int is3d = 0, face = 0, n = 0; unsigned state = 0;
    if (i < a.n) n = raw_process(a, i, poly, &is3d, &face, &state);
    unsigned x = (unsigned)n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, o); if ((int)lane >= o) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, w, o); if ((int)lane >= o) w += y; }
        s_warp[lane] = w;
    }
    __syncthreads();
    const unsigned off = (warp ? s_warp[warp - 1] : 0u) + x - (unsigned)n;
    for (int k = 0; k < n; k++) pfv_emit(out + off + k, &poly[0], &poly[k + 1], &poly[k + 2], state, face, is3d);
    if (i == 1023) { *d_total = off + (unsigned)n; atomicAdd(counters + 3, (unsigned long long)(off + (unsigned)n)); }
}

/* exclusive scan of up to 1024 items per CTA; sums[blockIdx] = CTA total */
__global__ void __launch_bounds__(256)
k_scan_block(const unsigned *__restrict__ in, unsigned *__restrict__ out, unsigned n, unsigned *__restrict__ sums)
{
    __shared__ unsigned s_warp[8];
    const unsigned base = blockIdx.x * 1024u + threadIdx.x * 4u;
    unsigned v[4], t = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = (base + k < n) ? in[base + k] : 0u; t += v[k]; }
    unsigned x = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    unsigned woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { const unsigned c = s_warp[w]; if (w < (int)(threadIdx.x >> 5)) woff += c; total += c; }
    unsigned run = woff + x - t;
#pragma unroll
    for (int k = 0; k < 4; k++) { if (base + k < n) out[base + k] = run; run += v[k]; }
    if (threadIdx.x == 0 && sums) sums[blockIdx.x] = total;
}

__global__ void k_scan_add(unsigned *__restrict__ data, unsigned n, const unsigned *__restrict__ block_offsets)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] += block_offsets[i / 1024u];
}

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

/* ------------------------------------------------------------------------------------------------ */
/* host: runtime                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

template <typename T> static int grow(T **p, size_t *cap, size_t need)
{
    if (need <= *cap) return PFCU_OK;
    size_t ncap = *cap ? *cap : 1024;
    while (ncap < need) ncap *= 2;
    CK(cudaStreamSynchronize(LN.stream));
    cudaFree(*p); *p = nullptr; *cap = 0;
    if (cudaMalloc(p, ncap * sizeof(T)) != cudaSuccess) { snprintf(g.err, sizeof g.err, "out of device memory growing scratch to %zu elements", ncap); return PFCU_ERR_OOM; }
    *cap = ncap;
    return PFCU_OK;
}

extern "C" {

const char *pfcu_last_error(void) { return g.err; }
const char *pfcu_backend_name(void) { return "cuda-sm_100a"; }

int pfcu_init(int device)
{
    API_LOCK;
    if (g.ok) return PFCU_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        snprintf(g.err, sizeof g.err, "no CUDA device: %s", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return PFCU_ERR_NO_DEVICE;
    }
    if (device < 0) {
        const char *env = getenv("PF_CUDA_DEVICE");
        if (!env) env = getenv("LOCAL_RANK");
        device = env ? atoi(env) : 0;
        if (device < 0 || device >= count) device = 0;
    }
    if (device >= count) { snprintf(g.err, sizeof g.err, "device %d out of range (%d devices)", device, count); return PFCU_ERR_NO_DEVICE; }
    CK(cudaSetDevice(device));
    g.device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    g.sms = prop.multiProcessorCount;
    if (prop.major < 10) {
        snprintf(g.err, sizeof g.err, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return PFCU_ERR_NO_DEVICE;
    }
    {
        const char *env = getenv("PF_CUDA_LANES");
        g.n_lanes = env ? atoi(env) : 4;
        if (g.n_lanes < 1) g.n_lanes = 1;
        if (g.n_lanes > MAX_LANES) g.n_lanes = MAX_LANES;
    }
    for (int i = 0; i < g.n_lanes; i++) {
        g.cur = &g.lanes[i];
        CK(cudaStreamCreateWithFlags(&LN.stream, cudaStreamNonBlocking));
        LN.own_stream = true;
        CK(cudaMalloc(&LN.d_bin_start, (MAX_BINS + 2) * 2 * sizeof(unsigned)));
        CK(cudaEventCreateWithFlags(&LN.stage_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&LN.states_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&LN.fence, cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&LN.vstream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&LN.raw_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&LN.vready, cudaEventDisableTiming));
        CK(cudaHostAlloc(&LN.h_total, 2 * sizeof(unsigned), cudaHostAllocDefault));
        CK(cudaMalloc(&LN.d_total, 64));
    }
    g.cur = &g.lanes[0];
    CK(cudaMalloc(&g.d_counters, 4 * sizeof(unsigned long long)));
    CK(cudaMemset(g.d_counters, 0, 4 * sizeof(unsigned long long)));
    g.ok = true;
    return PFCU_OK;
}

static void sync_all_lanes(void) { for (int i = 0; i < g.n_lanes; i++) cudaStreamSynchronize(g.lanes[i].stream); }
static void use_lane(const pfcu_surface *s) { g.cur = &g.lanes[s ? s->lane % g.n_lanes : 0]; }

void pfcu_shutdown(void)
{
    API_LOCK;
    if (!g.ok) return;
    sync_all_lanes();
    for (int i = 0; i < g.n_lanes; i++) {
        g.cur = &g.lanes[i];
        cudaFree(LN.d_tris); cudaFree(LN.d_states); cudaFree(LN.d_bbox); cudaFree(LN.d_setup); cudaFree(LN.d_data);
        cudaFree(LN.d_bin_counts); cudaFree(LN.d_bin_list); cudaFree(LN.d_bin_start); cudaFree(LN.d_varrays); cudaFree(LN.d_vcounts);
        if (LN.h_stage) cudaFreeHost(LN.h_stage);
        if (LN.h_states) cudaFreeHost(LN.h_states);
        if (LN.h_total) cudaFreeHost(LN.h_total);
        cudaFree(LN.d_raw); cudaFree(LN.d_total);
        if (LN.vstream) cudaStreamDestroy(LN.vstream);
        if (LN.own_stream) cudaStreamDestroy(LN.stream);
        g.lanes[i] = Lane();
    }
    cudaFree(g.d_counters); cudaFree(g.d_rcp); cudaFree(g.d_rsq);
    g.d_counters = nullptr; g.d_rcp = nullptr; g.d_rsq = nullptr;
    g.pinned.clear(); g.cur = nullptr; g.ok = false;
}

/* Order everything enqueued so far on every lane before everything enqueued afterwards on every lane, without
 * blocking the host: used by callers that bracket multi-surface work with events on lane 0's stream. */
int pfcu_fence(void)
{
    API_LOCK;
    if (!g.ok) return PFCU_ERR_NO_DEVICE;
    for (int i = 1; i < g.n_lanes; i++) {
        CK(cudaEventRecord(g.lanes[i].fence, g.lanes[i].stream));
        CK(cudaStreamWaitEvent(g.lanes[0].stream, g.lanes[i].fence, 0));
    }
    CK(cudaEventRecord(g.lanes[0].fence, g.lanes[0].stream));
    for (int i = 1; i < g.n_lanes; i++) CK(cudaStreamWaitEvent(g.lanes[i].stream, g.lanes[0].fence, 0));
    return PFCU_OK;
}

int pfcu_set_stream(void *cuda_stream)
{
    API_LOCK;
    if (!g.ok) return PFCU_ERR_NO_DEVICE;
    sync_all_lanes();
    Lane &l0 = g.lanes[0];
    if (l0.own_stream) { cudaStreamDestroy(l0.stream); l0.own_stream = false; }
    l0.stream = (cudaStream_t)cuda_stream;       /* lane 0 adopts the caller's stream; see pfcu_fence() */
    return PFCU_OK;
}

void *pfcu_get_stream(void) { return g.ok ? (void *)g.lanes[0].stream : nullptr; }

void *pfcu_host_alloc(size_t bytes)
{
    if (!g.ok && pfcu_init(-1) != PFCU_OK) return nullptr;
    PinnedBlock b; b.bytes = bytes; b.pending = false;
    if (cudaHostAlloc(&b.p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming) != cudaSuccess) { cudaFreeHost(b.p); return nullptr; }
    API_LOCK;
    g.pinned.push_back(b);
    return b.p;
}

static PinnedBlock *find_pinned(const void *p)
{
    for (auto &b : g.pinned) if ((const char *)p >= (const char *)b.p && (const char *)p < (const char *)b.p + b.bytes) return &b;
    return nullptr;
}

void pfcu_host_free(void *p)
{
    API_LOCK;
    for (size_t i = 0; i < g.pinned.size(); i++) if (g.pinned[i].p == p) {
        if (g.pinned[i].pending) cudaEventSynchronize(g.pinned[i].done);
        cudaEventDestroy(g.pinned[i].done); cudaFreeHost(p);
        g.pinned.erase(g.pinned.begin() + i);
        return;
    }
}

int pfcu_host_register(void *p, size_t bytes)
{
    API_LOCK;
    if (!g.ok || !p || bytes == 0) return PFCU_ERR_INVALID;
    /* page-aligned sub-range; the partial first/last pages stay pageable (cudaMemcpy handles mixed ranges) */
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return PFCU_ERR_CUDA; }
    return PFCU_OK;
}

void pfcu_host_unregister(void *p)
{
    API_LOCK;
    if (!g.ok || !p) return;
    sync_all_lanes();
    if (cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();
}

int pfcu_host_wait(const void *p)
{
    API_LOCK;
    PinnedBlock *b = find_pinned(p);
    if (b && b->pending) { CK(cudaEventSynchronize(b->done)); b->pending = false; }
    return PFCU_OK;
}

int pfcu_set_approx_tables(const uint32_t *rcp, int rcp_bits, const uint32_t *rsqrt, int rsqrt_bits)
{
    API_LOCK;
    if (!g.ok) return PFCU_ERR_NO_DEVICE;
    if (rcp_bits < 1 || rcp_bits > 23 || rsqrt_bits < 1 || rsqrt_bits > 23) return PFCU_ERR_INVALID;
    sync_all_lanes();
    cudaFree(g.d_rcp); cudaFree(g.d_rsq);
    CK(cudaMalloc(&g.d_rcp, sizeof(uint32_t) << rcp_bits));
    CK(cudaMalloc(&g.d_rsq, sizeof(uint32_t) << (rsqrt_bits + 1)));
    CK(cudaMemcpy(g.d_rcp, rcp, sizeof(uint32_t) << rcp_bits, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g.d_rsq, rsqrt, sizeof(uint32_t) << (rsqrt_bits + 1), cudaMemcpyHostToDevice));
    const int rshift = 23 - rcp_bits, sshift = 23 - rsqrt_bits;
    CK(cudaMemcpyToSymbol(c_rcp_tab, &g.d_rcp, sizeof(void *)));
    CK(cudaMemcpyToSymbol(c_rsq_tab, &g.d_rsq, sizeof(void *)));
    CK(cudaMemcpyToSymbol(c_rcp_shift, &rshift, sizeof(int)));
    CK(cudaMemcpyToSymbol(c_rsq_shift, &sshift, sizeof(int)));
    CK(cudaMemcpyToSymbol(c_rsq_bits, &rsqrt_bits, sizeof(int)));
    g.rcp_bits = rcp_bits; g.rsq_bits = rsqrt_bits;
    return PFCU_OK;
}

/* ---- surfaces ---- */

static void surface_dims(pfcu_surface *s) { s->tiles_x = (s->w + TILE - 1) / TILE; s->tiles_y = (s->h + TILE - 1) / TILE; }

pfcu_surface *pfcu_surface_create(uint32_t w, uint32_t h)
{
    API_LOCK;
    if (!g.ok || w == 0 || h == 0) { snprintf(g.err, sizeof g.err, "surface_create: runtime not initialised or empty surface"); return nullptr; }
    pfcu_surface *s = (pfcu_surface *)calloc(1, sizeof *s);
    if (!s) return nullptr;
    s->w = w; s->h = h; s->owned = true; s->world = 1;
    s->lane = (int)(g.next_lane++ % (unsigned)g.n_lanes);
    cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming);
    use_lane(s);
    surface_dims(s);
    const size_t bytes = ((size_t)w * h + 64) * 4;
    if (cudaMalloc(&s->color, bytes) != cudaSuccess || cudaMalloc(&s->depth, bytes) != cudaSuccess) {
        snprintf(g.err, sizeof g.err, "surface_create: out of device memory (%ux%u)", w, h);
        cudaFree(s->color); free(s); return nullptr;
    }
    cudaMemsetAsync(s->color, 0, bytes, LN.stream);
    cudaMemsetAsync(s->depth, 0, bytes, LN.stream);
    return s;
}

pfcu_surface *pfcu_surface_wrap(void *dev_color, void *dev_depth, uint32_t w, uint32_t h)
{
    API_LOCK;
    if (!g.ok || !dev_color || !dev_depth) return nullptr;
    pfcu_surface *s = (pfcu_surface *)calloc(1, sizeof *s);
    if (!s) return nullptr;
    s->w = w; s->h = h; s->color = (uint32_t *)dev_color; s->depth = (float *)dev_depth; s->owned = false; s->world = 1;
    s->lane = 0;                                  /* caller-owned memory: stay on the caller-visible stream */
    cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming);
    surface_dims(s);
    return s;
}

void pfcu_surface_destroy(pfcu_surface *s)
{
    API_LOCK;
    if (!s) return;
    if (g.ok) sync_all_lanes();
    if (s->owned) { cudaFree(s->color); cudaFree(s->depth); }
    if (s->done) cudaEventDestroy(s->done);
    free(s);
}

uint32_t pfcu_surface_width(const pfcu_surface *s) { return s->w; }
uint32_t pfcu_surface_height(const pfcu_surface *s) { return s->h; }
void *pfcu_surface_color_ptr(const pfcu_surface *s) { return s->color; }
void *pfcu_surface_depth_ptr(const pfcu_surface *s) { return s->depth; }

static void mark_done(pfcu_surface *s) { if (cudaEventRecord(s->done, LN.stream) == cudaSuccess) s->has_done = true; }

int pfcu_surface_upload(pfcu_surface *s, const void *hc, const float *hd, uint32_t y0, uint32_t rows)
{
    API_LOCK;
    if (y0 > s->h || rows > s->h - y0) return PFCU_ERR_INVALID;
    use_lane(s);
    const size_t off = (size_t)y0 * s->w, n = (size_t)rows * s->w * 4;
    if (hc) { CK(cudaMemcpyAsync(s->color + off, (const uint32_t *)hc + off, n, cudaMemcpyHostToDevice, LN.stream)); g.bytes_h2d += n; }
    if (hd) { CK(cudaMemcpyAsync(s->depth + off, hd + off, n, cudaMemcpyHostToDevice, LN.stream)); g.bytes_h2d += n; }
    /* pageable sources are staged by the driver before the call returns; pinned ones are not */
    CK(cudaStreamSynchronize(LN.stream));
    return PFCU_OK;
}

int pfcu_surface_download(pfcu_surface *s, void *hc, float *hd, uint32_t y0, uint32_t rows)
{
    API_LOCK;
    if (y0 > s->h || rows > s->h - y0) return PFCU_ERR_INVALID;
    use_lane(s);
    const size_t off = (size_t)y0 * s->w, n = (size_t)rows * s->w * 4;
    if (hc) { CK(cudaMemcpyAsync((uint32_t *)hc + off, s->color + off, n, cudaMemcpyDeviceToHost, LN.stream)); g.bytes_d2h += n; }
    if (hd) { CK(cudaMemcpyAsync(hd + off, s->depth + off, n, cudaMemcpyDeviceToHost, LN.stream)); g.bytes_d2h += n; }
    CK(cudaStreamSynchronize(LN.stream));
    return PFCU_OK;
}

static int fill_range(pfcu_surface *s, size_t first, size_t n, int dc, uint32_t rgba, int dd, float z)
{
    if (n == 0) return PFCU_OK;
    /* head (to 4-pixel alignment), vector body, tail */
    size_t head = (4 - (first & 3)) & 3; if (head > n) head = n;
    const size_t body4 = (n - head) / 4, tail = n - head - body4 * 4;
    const int blocks = g.sms * 8;
    if (head) { k_fill<<<1, 32, 0, LN.stream>>>(s->color, s->depth, first, head, dc, rgba, dd, z); g.launches++; }
    if (body4) {
        k_fill4<<<blocks, 256, 0, LN.stream>>>((uint4 *)s->color, (float4 *)s->depth, (first + head) / 4, body4, dc, rgba, dd, z);
        g.launches++;
    }
    if (tail) { k_fill<<<1, 32, 0, LN.stream>>>(s->color, s->depth, first + head + body4 * 4, tail, dc, rgba, dd, z); g.launches++; }
    CK(cudaGetLastError());
    return PFCU_OK;
}

int pfcu_surface_fill(pfcu_surface *s, int dc, uint32_t rgba, int dd, float z)
{
    API_LOCK;
    use_lane(s);
    const int rc = fill_range(s, 0, (size_t)s->w * s->h, dc, rgba, dd, z);
    mark_done(s);
    return rc;
}

int pfcu_surface_clear_ref(pfcu_surface *s, int dc, uint32_t rgba, int dd, float z)
{
    API_LOCK;
    use_lane(s);
    const unsigned size = s->w * s->h, aligned = size - (size % 8u);
    if (aligned > 8) { int rc = fill_range(s, 8, aligned - 8, dc, rgba, dd, z); if (rc) return rc; }
    if (aligned < size) { k_clear_tail<<<1, 32, 0, LN.stream>>>(s->color, s->depth, aligned, size, dc, dd); g.launches++; }
    CK(cudaGetLastError());
    mark_done(s);
    return PFCU_OK;
}

int pfcu_surface_set_tile_owner(pfcu_surface *s, uint32_t rank, uint32_t world)
{
    if (world == 0) world = 1;
    if (rank >= world) return PFCU_ERR_INVALID;
    s->rank = rank; s->world = world;
    return PFCU_OK;
}

static uint32_t owned_tiles(const pfcu_surface *s, uint32_t rank, uint32_t world)
{
    const uint32_t nt = s->tiles_x * s->tiles_y;
    if (world <= 1) return nt;
    return nt / world + ((nt % world) > rank ? 1u : 0u);
}

size_t pfcu_surface_owned_bytes(const pfcu_surface *s, uint32_t rank, uint32_t world, int with_depth)
{
    return (size_t)owned_tiles(s, rank, world) * TILE_PIX * 4u * (with_depth ? 2u : 1u);
}

static int pack_unpack(pfcu_surface *s, uint32_t rank, uint32_t world, int with_depth, void *staging, int unpack)
{
    API_LOCK;
    if (world == 0) world = 1;
    use_lane(s);
    const uint32_t n = owned_tiles(s, rank, world);
    if (n == 0) return PFCU_OK;
    k_pack_tiles<<<n, 256, 0, LN.stream>>>(s->color, s->depth, (int)s->w, (int)s->h, (int)s->tiles_x, s->tiles_x * s->tiles_y,
                                          rank, world, with_depth, (uint32_t *)staging, unpack);
    g.launches++;
    CK(cudaGetLastError());
    return PFCU_OK;
}

int pfcu_surface_pack_tiles(pfcu_surface *s, uint32_t r, uint32_t w, int wd, void *st) { return pack_unpack(s, r, w, wd, st, 0); }
int pfcu_surface_unpack_tiles(pfcu_surface *s, uint32_t r, uint32_t w, int wd, const void *st) { return pack_unpack(s, r, w, wd, (void *)st, 1); }

/* ---- textures ---- */

static size_t tex_bytes(uint32_t w, uint32_t h, int fmt) { return (size_t)w * h * ((fmt == PFCU_TEX_RGBA8 || fmt == PFCU_TEX_BGRA8) ? 4u : 3u); }

pfcu_texture *pfcu_texture_create(const void *host_pixels, uint32_t w, uint32_t h, int fmt)
{
    API_LOCK;
    if (!g.ok || fmt < PFCU_TEX_RGBA8 || fmt > PFCU_TEX_BGR8 || w == 0 || h == 0) return nullptr;
    pfcu_texture *t = (pfcu_texture *)calloc(1, sizeof *t);
    if (!t) return nullptr;
    t->w = w; t->h = h; t->fmt = fmt; t->owned = true;
    g.cur = &g.lanes[0];
    const size_t bytes = tex_bytes(w, h, fmt);
    if (cudaMalloc(&t->pixels, bytes + 16) != cudaSuccess) { snprintf(g.err, sizeof g.err, "texture_create: out of device memory"); free(t); return nullptr; }
    cudaMemsetAsync(t->pixels, 0, bytes + 16, LN.stream);
    if (host_pixels && pfcu_texture_update(t, host_pixels) != PFCU_OK) { cudaFree(t->pixels); free(t); return nullptr; }
    return t;
}

pfcu_texture *pfcu_texture_from_surface(pfcu_surface *s)
{
    API_LOCK;
    pfcu_texture *t = (pfcu_texture *)calloc(1, sizeof *t);
    if (!t) return nullptr;
    t->w = s->w; t->h = s->h; t->fmt = PFCU_TEX_RGBA8; t->pixels = (unsigned char *)s->color; t->owned = false; t->alias = s;
    return t;
}

int pfcu_texture_update(pfcu_texture *t, const void *host_pixels)
{
    API_LOCK;
    if (!t || !t->owned || !host_pixels) return PFCU_ERR_INVALID;
    sync_all_lanes();                           /* nobody may still be sampling the old texels */
    g.cur = &g.lanes[0];
    CK(cudaMemcpyAsync(t->pixels, host_pixels, tex_bytes(t->w, t->h, t->fmt), cudaMemcpyHostToDevice, LN.stream));
    g.bytes_h2d += tex_bytes(t->w, t->h, t->fmt);
    CK(cudaStreamSynchronize(LN.stream));
    return PFCU_OK;
}

void pfcu_texture_destroy(pfcu_texture *t)
{
    API_LOCK;
    if (!t) return;
    if (g.ok) sync_all_lanes();
    if (t->owned) cudaFree(t->pixels);
    free(t);
}

/* ---- the hot path ---- */

/* state program of a DevState, same numbering as k_raster's per-triangle dispatch */
static int state_program(const DevState *d)
{
    if (d->flags & PFCU_ST_PHONG) return 12;
    int texm = 0;
    if (d->flags & PFCU_ST_TEXTURE) texm = (d->tfmt == PFCU_TEX_RGBA8 && d->tex_wrap == 0 && d->tex_filter == 0) ? 1 : 2;
    const int blendm = !(d->flags & PFCU_ST_BLEND) ? 0 : (d->blend_mode == 1 ? 1 : (d->blend_mode == 2 ? 2 : 3));
    return texm * 4 + blendm;
}

static int g_last_single_prog = -1;      /* set by convert_states: the common program of all states, or -1 */

static unsigned convert_states(const pfcu_state *in, uint32_t n, DevState *out)
{
    unsigned mask = 0;
    g.deps.clear();
    for (uint32_t i = 0; i < n; i++) {
        const pfcu_state *s = in + i; DevState *d = out + i;
        memset(d, 0, sizeof *d);
        d->flags = s->flags;
        if (!(s->texture) ) d->flags &= ~PFCU_ST_TEXTURE;
        if (s->n_lights == 0) d->flags &= ~PFCU_ST_PHONG;
        d->blend_mode = s->blend_mode; d->depth_func = s->depth_func; d->tex_filter = s->tex_filter; d->tex_wrap = s->tex_wrap;
        d->vp_min[0] = s->vp_min[0]; d->vp_min[1] = s->vp_min[1]; d->vp_max[0] = s->vp_max[0]; d->vp_max[1] = s->vp_max[1];
        if (d->flags & PFCU_ST_TEXTURE) {
            d->tex = s->texture->pixels; d->tw = s->texture->w; d->th = s->texture->h; d->tfmt = s->texture->fmt;
            d->tex_fw = (float)d->tw; d->tex_fh = (float)d->th;
            { volatile float one = 1.0f; d->tex_tx = one / d->tex_fw; d->tex_ty = one / d->tex_fh; }   /* IEEE single division, as DIVPS */
            if (s->texture->alias) g.deps.push_back(s->texture->alias);      /* render-to-texture: order across lanes */
        }
        d->n_lights = s->n_lights > 8 ? 8 : s->n_lights;
        for (unsigned l = 0; l < d->n_lights; l++) {
            const pfcu_light *a = &s->lights[l]; DevLight *b = &d->lights[l];
            memcpy(b->pos, a->position, 12); memcpy(b->dir, a->direction, 12);
            b->inner = a->inner_cutoff; b->outer = a->outer_cutoff;
            b->attc = a->att_constant; b->attl = a->att_linear; b->attq = a->att_quadratic;
            b->ambient = a->ambient; b->diffuse = a->diffuse; b->specular = a->specular;
        }
        for (int f = 0; f < 2; f++) {
            d->material[f].ambient = s->material[f].ambient; d->material[f].diffuse = s->material[f].diffuse;
            d->material[f].specular = s->material[f].specular; d->material[f].emission = s->material[f].emission;
            d->material[f].shininess = s->material[f].shininess;
        }
        memcpy(d->view_pos, s->view_pos, 12);
        mask |= d->flags;
        const int prog = state_program(d);
        if (i == 0) g_last_single_prog = prog; else if (g_last_single_prog != prog) g_last_single_prog = -1;
    }
    return mask;
}

/* n: number of triangles, or - when d_n is given - an upper bound of the count the device holds in *d_n (raw batches
 * after clipping); n_est then stands in for n where only the batch shape matters. */
static int launch_pipeline(pfcu_surface *s, const pfcu_triangle *d_tris, const DevState *d_states, uint32_t n, unsigned feature_mask, int single_prog = -1,
                           const unsigned *d_n = nullptr, uint32_t n_est = 0, int bshift_forced = 0)
{
    if (n == 0) return PFCU_OK;
    if (!d_n) n_est = n;
    if (!g.d_rcp) { snprintf(g.err, sizeof g.err, "pfcu_set_approx_tables() has not been called"); return PFCU_ERR_INVALID; }
    int rc;
    /* surfaces sampled as textures that live on another lane: wait for their last write */
    for (pfcu_surface *dep : g.deps)
        if (dep != s && dep->lane != s->lane && dep->has_done) CK(cudaStreamWaitEvent(LN.stream, dep->done, 0));
    if (n > LN.cap_setup) {
        size_t c1 = LN.cap_setup, c2 = LN.cap_setup, c3 = LN.cap_setup;
        if ((rc = grow(&LN.d_bbox, &c1, n))) return rc;
        if ((rc = grow(&LN.d_setup, &c2, n))) return rc;
        if ((rc = grow(&LN.d_data, &c3, n))) return rc;
        LN.cap_setup = c1;
    }
    /* many small triangles per tile: fine bins (one per tile) and the fragment-compacting rasteriser;
       few large ones: coarse bins and the triangle-per-warp-step rasteriser */
    const unsigned nTilesAll = s->tiles_x * s->tiles_y;
    const bool small_tris = (size_t)n_est > (size_t)4 * nTilesAll;
    static const int force_bshift = getenv("PF_CUDA_BIN_SHIFT") ? atoi(getenv("PF_CUDA_BIN_SHIFT")) : 0;
    int bshift = bshift_forced ? bshift_forced : (force_bshift >= 6 ? force_bshift : (small_tris ? BIN_SHIFT_FINE : BIN_SHIFT_COARSE));
    int binsX, binsY, nb;
    for (;; bshift++) {
        binsX = (int)((s->w + (1u << bshift) - 1) >> bshift); binsY = (int)((s->h + (1u << bshift) - 1) >> bshift);
        nb = binsX * binsY;
        if (nb <= MAX_BINS) break;
    }
    if (bshift > 8) { snprintf(g.err, sizeof g.err, "surface too large for the binner (%u x %u)", s->w, s->h); return PFCU_ERR_INVALID; }
    const unsigned nBatches = (n + BIN_BATCH - 1) / BIN_BATCH;
    if ((rc = grow(&LN.d_bin_counts, &LN.cap_bin_counts, (size_t)nBatches * nb))) return rc;

    cudaEvent_t pe[3] = { nullptr, nullptr, nullptr };
    if (g.profiling) {
        for (int i = 0; i < 3; i++) {
            if (!g.prof_pool.empty()) { pe[i] = g.prof_pool.back(); g.prof_pool.pop_back(); }
            else CK(cudaEventCreate(&pe[i]));
        }
        CK(cudaEventRecord(pe[0], LN.stream));
    }
    if (d_n && !(nb <= 3072 && n <= FRONT_SMALL_MAX * FRONT_SMALL_CHUNKS)) { snprintf(g.err, sizeof g.err, "internal: device-side count outside the single-CTA front end"); return PFCU_ERR_INVALID; }
    if (d_n || (n <= FRONT_SMALL_MAX && nb <= 3072)) {      /* 3072 bin counters + the rectangles fit the 48 KB of static + dynamic shared memory */
        /* small batch: the (triangle, bin) overlap count is bounded by n * nb, no read-back needed */
        if ((rc = grow(&LN.d_bin_list, &LN.cap_bin_list, (size_t)n * nb))) return rc;
        k_front_small<<<1, 1024, nb * sizeof(unsigned), LN.stream>>>(d_tris, d_states, n, d_n, (int)s->w, (int)s->h, LN.d_bbox, LN.d_setup, LN.d_data,
                                                                      g.d_counters, binsX, binsY, bshift, LN.d_bin_start, LN.d_bin_list);
        g.launches += 1;
    } else {
        k_setup<<<(n + SETUP_THREADS - 1) / SETUP_THREADS, SETUP_THREADS, 0, LN.stream>>>(
            d_tris, d_states, n, (int)s->w, (int)s->h, LN.d_bbox, LN.d_setup, LN.d_data, g.d_counters);
        k_bin_count<<<nBatches, 256, nb * sizeof(unsigned), LN.stream>>>(LN.d_bbox, n, binsX, binsY, bshift, LN.d_bin_counts);
        unsigned *d_totals = LN.d_bin_start + (MAX_BINS + 2);
        k_bin_scan<<<(nb + 31) / 32, 1024, 0, LN.stream>>>(LN.d_bin_counts, (int)nBatches, nb, d_totals);
        k_bin_starts<<<1, 1024, 0, LN.stream>>>(d_totals, nb, LN.d_bin_start);
        /* Per-bin lists hold (triangle, bin) overlaps.  The exact total is only known on the device;
           n*nb bounds it.  Small cases are sized by the bound, large ones read the total back. */
        {
            const size_t bound = (size_t)n * (size_t)nb;
            size_t want = bound <= ((size_t)n * 4 > 65536 ? (size_t)n * 4 : 65536) ? bound : 0;
            if (!want && bound <= LN.cap_bin_list) want = bound;
            if (!want) {
                unsigned total = 0;
                CK(cudaMemcpyAsync(&total, LN.d_bin_start + nb, sizeof(unsigned), cudaMemcpyDeviceToHost, LN.stream));
                CK(cudaStreamSynchronize(LN.stream));
                want = total;
            }
            if ((rc = grow(&LN.d_bin_list, &LN.cap_bin_list, want ? want : 1))) return rc;
        }
        k_bin_fill<<<nBatches, 256, nb * sizeof(unsigned), LN.stream>>>(LN.d_bbox, n, binsX, binsY, bshift, LN.d_bin_counts, LN.d_bin_start, LN.d_bin_list);
        g.launches += 5;
    }
    RasterParams p;
    p.bbox = LN.d_bbox; p.setup = LN.d_setup; p.data = LN.d_data; p.states = d_states;
    p.bin_list = LN.d_bin_list; p.bin_starts = LN.d_bin_start; p.binsX = binsX; p.bin_tshift = bshift - 6;
    p.color = s->color; p.depth = s->depth; p.W = (int)s->w; p.H = (int)s->h;
    p.tilesX = (int)s->tiles_x; p.tilesY = (int)s->tiles_y;
    p.rank = s->rank; p.world = s->world ? s->world : 1; p.nTiles = s->tiles_x * s->tiles_y;
    p.counters = g.d_counters;
    const unsigned grid = owned_tiles(s, p.rank, p.world);
    if (g.profiling) CK(cudaEventRecord(pe[1], LN.stream));
    if (grid) {
        /* many small triangles per tile: 16 warps per tile halve the serial work of the busiest tiles;
           few large ones: 8 warps with more registers each issue faster */
        const bool ph = (feature_mask & PFCU_ST_PHONG) != 0;
        if (g.rcp_bits > RCP_SMEM_BITS) single_prog = -1;      /* the fixed-program kernels assume the shared RCPPS table */
        /* half-height slices when the 64x64 grid would be only a few waves deep with a ragged last wave */
        const int per_sm = (single_prog == 5 || single_prog == 6) ? 4 : 3;
        const double waves = (double)grid / ((double)g.sms * per_sm);
        const bool half = !small_tris && !ph && waves < 8.0 && (ceil(waves) / waves) > 1.06 && (ceil(2 * waves) / (2 * waves)) < (ceil(waves) / waves);
        /* small triangles: slices of 64x16 (64x32 with Phong) shorten the serial work of the busiest tiles and
           even out the SMs (C2: 0.51 -> 0.30 ms); PF_CUDA_SLICE=64|32|16 overrides for experiments */
        static const int force_slice = getenv("PF_CUDA_SLICE") ? atoi(getenv("PF_CUDA_SLICE")) : 0;
        static const bool env_once = [] {
            const char *e = getenv("PF_CUDA_FRAG");
            if (e) g.raster_path = atoi(e) == 0 ? PFCU_RASTER_TILES : (atoi(e) == 2 ? PFCU_RASTER_FRAGMENTS : PFCU_RASTER_AUTO);
            return true; }();
        (void)env_once;
        const bool use_frag = g.raster_path == PFCU_RASTER_FRAGMENTS || (g.raster_path == PFCU_RASTER_AUTO && small_tris && !force_slice);
        if (use_frag) {
            /* 64x8 slices of eight 8x8 regions, 8 warps; 4 CTAs per SM (Phong: 3, 80 registers) */
            static const bool attr_once = [] {
                cudaFuncSetAttribute(k_raster_frag<true, 8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * FRAG_NF_PHONG * 512);
                cudaFuncSetAttribute(k_raster_frag<false, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * FRAG_NF * 512);
                return true; }();
            (void)attr_once;
            if (ph) k_raster_frag<true, 8, 3><<<grid * 8, 256, 8 * FRAG_NF_PHONG * 512, LN.stream>>>(p);
            else    k_raster_frag<false, 8, 4><<<grid * 8, 256, 8 * FRAG_NF * 512, LN.stream>>>(p);
        }
        else if (small_tris) {
            const int th = force_slice ? force_slice : (ph ? 32 : 16);
            if (ph) { if (th <= 32) k_raster<true, 16, -1, 32><<<grid * 2, 512, 0, LN.stream>>>(p); else k_raster<true, 16, -1, 64><<<grid, 512, 0, LN.stream>>>(p); }
            else if (th <= 16) k_raster<false, 16, -1, 16><<<grid * 4, 512, 0, LN.stream>>>(p);
            else if (th <= 32) k_raster<false, 16, -1, 32><<<grid * 2, 512, 0, LN.stream>>>(p);
            else               k_raster<false, 16, -1, 64><<<grid, 512, 0, LN.stream>>>(p);
        }
        else if (ph)          k_raster<true, 8, -1, 64><<<grid, 256, 0, LN.stream>>>(p);
        else if (single_prog == 5) { if (half) k_raster<false, 8, 5, 32><<<grid * 2, 256, 0, LN.stream>>>(p); else k_raster<false, 8, 5, 64><<<grid, 256, 0, LN.stream>>>(p); }   /* nearest REPEAT RGBA8 texture + ALPHA blend */
        else if (single_prog == 6) { if (half) k_raster<false, 8, 6, 32><<<grid * 2, 256, 0, LN.stream>>>(p); else k_raster<false, 8, 6, 64><<<grid, 256, 0, LN.stream>>>(p); }   /* nearest REPEAT RGBA8 texture + ADD blend   */
        else { if (half) k_raster<false, 8, -1, 32><<<grid * 2, 256, 0, LN.stream>>>(p); else k_raster<false, 8, -1, 64><<<grid, 256, 0, LN.stream>>>(p); }
        g.launches++;
    }
    if (g.profiling) { CK(cudaEventRecord(pe[2], LN.stream)); for (int i = 0; i < 3; i++) g.prof_events.push_back(pe[i]); }
    CK(cudaGetLastError());
    mark_done(s);
    /* write-after-read: a sampled surface on another lane must not be overwritten before this batch read it */
    for (pfcu_surface *dep : g.deps)
        if (dep != s && dep->lane != s->lane) CK(cudaStreamWaitEvent(g.lanes[dep->lane % g.n_lanes].stream, s->done, 0));
    g.deps.clear();
    if (!d_n) g.submitted += n;              /* otherwise counted on the device (counters[3]) */
    return PFCU_OK;
}

/* Can a raw batch of n_raw triangles run without the host learning its output count?  Picks the bin size. */
static int sync_free_bshift(const pfcu_surface *s, uint32_t n_raw)
{
    static const int off = getenv("PF_CUDA_RAW_SYNC") ? atoi(getenv("PF_CUDA_RAW_SYNC")) : 0;
    if (off || n_raw > FRONT_SMALL_MAX) return 0;
    const size_t bound = (size_t)n_raw * FRONT_SMALL_CHUNKS;
    const bool small_tris = (size_t)n_raw > (size_t)4 * s->tiles_x * s->tiles_y;
    for (int bshift = small_tris ? BIN_SHIFT_FINE : BIN_SHIFT_COARSE; bshift <= BIN_SHIFT_COARSE; bshift += 2) {
        const size_t nb = (size_t)((s->w + (1u << bshift) - 1) >> bshift) * ((s->h + (1u << bshift) - 1) >> bshift);
        if (nb <= 3072 && bound * nb <= ((size_t)2 << 20)) return bshift;      /* bin list <= 16 MB */
    }
    return 0;
}

/* ---- device vertex stage ---- */

static int scan_exclusive(const unsigned *d_in, unsigned *d_out, unsigned n, unsigned *d_tmp /* >= n/1024 + n/1048576 + 4 */, cudaStream_t st = nullptr)
{
    if (!st) st = LN.stream;
    const unsigned nb = (n + 1023u) / 1024u;
    k_scan_block<<<nb, 256, 0, st>>>(d_in, d_out, n, d_tmp);
    g.launches++;
    if (nb > 1) {
        unsigned *d_tmp2 = d_tmp + nb;
        int rc = scan_exclusive(d_tmp, d_tmp, nb, d_tmp2, st);
        if (rc) return rc;
        k_scan_add<<<(n + 255u) / 256u, 256, 0, st>>>(d_out, n, d_tmp);
        g.launches++;
    }
    CK(cudaGetLastError());
    return PFCU_OK;
}

unsigned pfcu_capabilities(void) { return PFCU_CAP_DEVICE_VERTEX | PFCU_CAP_RAW_TRIANGLES; }

int pfcu_draw_triangles(pfcu_surface *s, const pfcu_state *state, const pfcu_vparams *vp, const pfcu_draw *d, uint32_t *n_out)
{
    API_LOCK;
    if (n_out) *n_out = 0;
    if (!g.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || !state || !vp || !d || !d->positions || d->pos_size < 2 || d->pos_size > 4 || d->n_faces < 1 || d->n_faces > 2) return PFCU_ERR_INVALID;
    const unsigned n_tri = d->count / 3u;
    if (n_tri == 0) return PFCU_OK;
    use_lane(s);
    const unsigned n_items = n_tri * d->n_faces;
    int rc;
    /* arrays -> device (pageable sources are staged by the driver; ordered on the stream) */
    const size_t nv = d->n_vertices;
    const size_t b_pos = nv * d->pos_size * 4, b_nrm = d->normals ? nv * 12 : 0, b_uv = d->texcoords ? nv * 8 : 0;
    const size_t b_col = d->colors ? nv * d->color_size : 0, b_idx = d->indices ? (size_t)d->count * d->index_bytes : 0;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t total_bytes = al(b_pos) + al(b_nrm) + al(b_uv) + al(b_col) + al(b_idx);
    g.bytes_h2d += b_pos + b_nrm + b_uv + b_col + b_idx + sizeof(DevState);
    if ((rc = grow(&LN.d_varrays, &LN.cap_varrays, total_bytes))) return rc;
    unsigned char *p = LN.d_varrays;
    VtxArgs a; memset(&a, 0, sizeof a);
    a.pos = (const float *)p; CK(cudaMemcpyAsync(p, d->positions, b_pos, cudaMemcpyHostToDevice, LN.stream)); p += al(b_pos);
    if (b_nrm) { a.nrm = (const float *)p; CK(cudaMemcpyAsync(p, d->normals, b_nrm, cudaMemcpyHostToDevice, LN.stream)); p += al(b_nrm); }
    if (b_uv) { a.uv = (const float *)p; CK(cudaMemcpyAsync(p, d->texcoords, b_uv, cudaMemcpyHostToDevice, LN.stream)); p += al(b_uv); }
    if (b_col) { a.col = p; CK(cudaMemcpyAsync(p, d->colors, b_col, cudaMemcpyHostToDevice, LN.stream)); p += al(b_col); }
    if (b_idx) { a.idx = p; CK(cudaMemcpyAsync(p, d->indices, b_idx, cudaMemcpyHostToDevice, LN.stream)); p += al(b_idx); }
    a.pos_size = (int)d->pos_size; a.col_size = (int)d->color_size; a.idx_bytes = (int)d->index_bytes;
    a.first = d->first; a.n_tri = n_tri; a.cur_color = d->current_color; a.n_faces = (int)d->n_faces;
    a.face[0] = d->faces[0]; a.face[1] = d->faces[1]; a.state = 0;

    /* pass A: output triangles per (triangle, face) item; scan; total */
    if ((rc = grow(&LN.d_vcounts, &LN.cap_vcounts, (size_t)n_items * 2 + n_items / 512 + 64))) return rc;
    unsigned *d_counts = LN.d_vcounts, *d_offsets = LN.d_vcounts + n_items, *d_tmp = LN.d_vcounts + 2 * (size_t)n_items;
    k_vertex_count<<<(n_items + 127u) / 128u, 128, 0, LN.stream>>>(a, *vp, n_items, d_counts);
    g.launches++;
    if ((rc = scan_exclusive(d_counts, d_offsets, n_items, d_tmp))) return rc;
    unsigned last[2] = { 0, 0 };
    CK(cudaMemcpyAsync(&last[0], d_offsets + (n_items - 1), 4, cudaMemcpyDeviceToHost, LN.stream));
    CK(cudaMemcpyAsync(&last[1], d_counts + (n_items - 1), 4, cudaMemcpyDeviceToHost, LN.stream));
    CK(cudaStreamSynchronize(LN.stream));
    const unsigned total = last[0] + last[1];
    if (n_out) *n_out = total;
    if (total == 0) return PFCU_OK;

    /* pass B: emit in order, then the usual setup -> bin -> raster pipeline */
    if ((rc = grow(&LN.d_tris, &LN.cap_tris, total))) return rc;
    if ((rc = grow(&LN.d_states, &LN.cap_states, 1))) return rc;
    DevState hs;
    const unsigned mask = convert_states(state, 1, &hs);
    CK(cudaMemcpyAsync(LN.d_states, &hs, sizeof hs, cudaMemcpyHostToDevice, LN.stream));
    k_vertex_emit<<<(n_items + 127u) / 128u, 128, 0, LN.stream>>>(a, *vp, n_items, d_offsets, LN.d_tris);
    g.launches++;
    CK(cudaEventRecord(LN.raw_done, LN.stream));        /* d_vcounts is shared with the raw-triangle path's side stream */
    CK(cudaGetLastError());
    return launch_pipeline(s, LN.d_tris, LN.d_states, total, mask, g_last_single_prog);
}

int pfcu_submit_raw(pfcu_surface *s, const pfcu_state *states, uint32_t n_states, const pfcu_vparams_lit *vparams, uint32_t n_vparams,
                    const float *pow_tables, uint32_t n_pow_tables, const pfcu_rawtri *tris, uint32_t n_tris, uint32_t *n_out)
{
    API_LOCK;
    if (n_out) *n_out = 0;
    if (!g.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || (n_tris && (!states || !tris || !vparams || n_states == 0 || n_vparams == 0))) return PFCU_ERR_INVALID;
    if (n_tris == 0) return PFCU_OK;
    use_lane(s);
    int rc;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_tris = (size_t)n_tris * sizeof(pfcu_rawtri), b_vp = (size_t)n_vparams * sizeof(pfcu_vparams_lit);
    const size_t b_pow = (size_t)n_pow_tables * PFCU_POW_TABLE_SIZE * sizeof(float);
    if (const int bshift = sync_free_bshift(s, n_tris)) {
        /* small batch: one fused count + scan + emit kernel, output count kept on the device, no host wait;
           everything on the surface's lane */
        const size_t need = al(b_tris) + al(b_vp) + al(b_pow);
        if ((rc = grow(&LN.d_raw, &LN.cap_raw, need))) return rc;
        const uint32_t bound = n_tris * FRONT_SMALL_CHUNKS;
        if ((rc = grow(&LN.d_tris, &LN.cap_tris, bound))) return rc;
        if ((rc = grow(&LN.d_states, &LN.cap_states, n_states))) return rc;
        unsigned char *q = LN.d_raw;
        RawArgs ra; ra.n = n_tris;
        ra.tris = (const pfcu_rawtri *)q;
        CK(cudaMemcpyAsync(q, tris, b_tris, cudaMemcpyHostToDevice, LN.stream));
        if (PinnedBlock *pb = find_pinned(tris)) { CK(cudaEventRecord(pb->done, LN.stream)); pb->pending = true; }
        q += al(b_tris);
        ra.vp = (const pfcu_vparams_lit *)q; CK(cudaMemcpyAsync(q, vparams, b_vp, cudaMemcpyHostToDevice, LN.stream)); q += al(b_vp);
        ra.pow_tables = (const float *)q; if (b_pow) CK(cudaMemcpyAsync(q, pow_tables, b_pow, cudaMemcpyHostToDevice, LN.stream));
        g.bytes_h2d += b_tris + b_vp + b_pow + n_states * sizeof(DevState);
        if (n_states > LN.cap_hstates) {
            CK(cudaEventSynchronize(LN.states_done));
            if (LN.h_states) cudaFreeHost(LN.h_states);
            size_t c = LN.cap_hstates ? LN.cap_hstates : 64; while (c < n_states) c *= 2;
            CK(cudaHostAlloc(&LN.h_states, c * sizeof(DevState), cudaHostAllocDefault));
            LN.cap_hstates = c;
        } else CK(cudaEventSynchronize(LN.states_done));
        const unsigned mask = convert_states(states, n_states, LN.h_states);
        CK(cudaMemcpyAsync(LN.d_states, LN.h_states, n_states * sizeof(DevState), cudaMemcpyHostToDevice, LN.stream));
        CK(cudaEventRecord(LN.states_done, LN.stream));
        unsigned *d_total = LN.d_total;
        k_raw_small<<<1, 1024, 0, LN.stream>>>(ra, LN.d_tris, d_total, g.d_counters);
        g.launches++;
        CK(cudaEventRecord(LN.raw_done, LN.stream));
        CK(cudaGetLastError());
        if (n_out) *n_out = n_tris;             /* the exact count stays on the device (pfcu_get_counters has it) */
        return launch_pipeline(s, LN.d_tris, LN.d_states, bound, mask, g_last_single_prog, d_total, n_tris, bshift);
    }
    /* the previous raw batch of this lane must have been emitted before its inputs are overwritten */
    CK(cudaStreamWaitEvent(LN.vstream, LN.raw_done, 0));
    if (al(b_tris) + al(b_vp) + al(b_pow) > LN.cap_raw) { CK(cudaEventSynchronize(LN.raw_done)); }
    if ((rc = grow(&LN.d_raw, &LN.cap_raw, al(b_tris) + al(b_vp) + al(b_pow)))) return rc;
    if ((rc = grow(&LN.d_vcounts, &LN.cap_vcounts, (size_t)n_tris * 2 + n_tris / 512 + 64))) return rc;
    unsigned char *p = LN.d_raw;
    RawArgs a; a.n = n_tris;
    a.tris = (const pfcu_rawtri *)p;
    CK(cudaMemcpyAsync(p, tris, b_tris, cudaMemcpyHostToDevice, LN.vstream));
    if (PinnedBlock *pb = find_pinned(tris)) { CK(cudaEventRecord(pb->done, LN.vstream)); pb->pending = true; }
    p += al(b_tris);
    a.vp = (const pfcu_vparams_lit *)p; CK(cudaMemcpyAsync(p, vparams, b_vp, cudaMemcpyHostToDevice, LN.vstream)); p += al(b_vp);
    a.pow_tables = (const float *)p; if (b_pow) CK(cudaMemcpyAsync(p, pow_tables, b_pow, cudaMemcpyHostToDevice, LN.vstream));
    g.bytes_h2d += b_tris + b_vp + b_pow + n_states * sizeof(DevState);

    unsigned *d_counts = LN.d_vcounts, *d_offsets = LN.d_vcounts + n_tris, *d_tmp = LN.d_vcounts + 2 * (size_t)n_tris;
    k_raw_count<<<(n_tris + 127u) / 128u, 128, 0, LN.vstream>>>(a, d_counts);
    g.launches++;
    if ((rc = scan_exclusive(d_counts, d_offsets, n_tris, d_tmp, LN.vstream))) return rc;
    CK(cudaMemcpyAsync(&LN.h_total[0], d_offsets + (n_tris - 1), 4, cudaMemcpyDeviceToHost, LN.vstream));
    CK(cudaMemcpyAsync(&LN.h_total[1], d_counts + (n_tris - 1), 4, cudaMemcpyDeviceToHost, LN.vstream));
    CK(cudaEventRecord(LN.vready, LN.vstream));
    CK(cudaStreamSynchronize(LN.vstream));
    const unsigned total = LN.h_total[0] + LN.h_total[1];
    if (n_out) *n_out = total;
    if (total == 0) { CK(cudaEventRecord(LN.raw_done, LN.vstream)); return PFCU_OK; }

    /* states + emission + the usual pipeline on the surface's lane */
    if ((rc = grow(&LN.d_tris, &LN.cap_tris, total))) return rc;
    if ((rc = grow(&LN.d_states, &LN.cap_states, n_states))) return rc;
    if (n_states > LN.cap_hstates) {
        CK(cudaEventSynchronize(LN.states_done));
        if (LN.h_states) cudaFreeHost(LN.h_states);
        size_t c = LN.cap_hstates ? LN.cap_hstates : 64; while (c < n_states) c *= 2;
        CK(cudaHostAlloc(&LN.h_states, c * sizeof(DevState), cudaHostAllocDefault));
        LN.cap_hstates = c;
    } else CK(cudaEventSynchronize(LN.states_done));
    const unsigned mask = convert_states(states, n_states, LN.h_states);
    CK(cudaMemcpyAsync(LN.d_states, LN.h_states, n_states * sizeof(DevState), cudaMemcpyHostToDevice, LN.stream));
    CK(cudaEventRecord(LN.states_done, LN.stream));
    CK(cudaStreamWaitEvent(LN.stream, LN.vready, 0));
    k_raw_emit<<<(n_tris + 127u) / 128u, 128, 0, LN.stream>>>(a, d_offsets, LN.d_tris);
    g.launches++;
    CK(cudaEventRecord(LN.raw_done, LN.stream));
    CK(cudaGetLastError());
    return launch_pipeline(s, LN.d_tris, LN.d_states, total, mask, g_last_single_prog);
}

int pfcu_submit_prims(pfcu_surface *s, const pfcu_prim *prims, uint32_t n)
{
    API_LOCK;
    if (!g.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || (n && !prims)) return PFCU_ERR_INVALID;
    if (n == 0) return PFCU_OK;
    use_lane(s);
    int rc;
    const size_t bytes = (size_t)n * sizeof(pfcu_prim);
    if ((rc = grow(&LN.d_varrays, &LN.cap_varrays, bytes))) return rc;
    CK(cudaMemcpyAsync(LN.d_varrays, prims, bytes, cudaMemcpyHostToDevice, LN.stream));
    g.bytes_h2d += bytes;
    PrimParams p;
    p.prims = (const pfcu_prim *)LN.d_varrays; p.n = n; p.color = s->color; p.depth = s->depth; p.W = s->w; p.H = s->h;
    p.tilesX = (int)s->tiles_x; p.rank = s->rank; p.world = s->world ? s->world : 1; p.nTiles = s->tiles_x * s->tiles_y;
    const unsigned grid = owned_tiles(s, p.rank, p.world);
    if (grid) { k_prims<<<grid, 256, 0, LN.stream>>>(p); g.launches++; }
    CK(cudaGetLastError());
    mark_done(s);
    return PFCU_OK;
}

int pfcu_submit(pfcu_surface *s, const pfcu_state *states, uint32_t n_states, const pfcu_triangle *tris, uint32_t n_tris)
{
    API_LOCK;
    if (!g.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || (n_tris && (!states || !tris || n_states == 0))) return PFCU_ERR_INVALID;
    if (n_tris == 0) return PFCU_OK;
    use_lane(s);
    int rc;
    if ((rc = grow(&LN.d_tris, &LN.cap_tris, n_tris))) return rc;
    if ((rc = grow(&LN.d_states, &LN.cap_states, n_states))) return rc;

    /* states: convert handles to device pointers in pinned staging */
    if (n_states > LN.cap_hstates) {
        CK(cudaEventSynchronize(LN.states_done));
        if (LN.h_states) cudaFreeHost(LN.h_states);
        size_t c = LN.cap_hstates ? LN.cap_hstates : 64; while (c < n_states) c *= 2;
        CK(cudaHostAlloc(&LN.h_states, c * sizeof(DevState), cudaHostAllocDefault));
        LN.cap_hstates = c;
    } else CK(cudaEventSynchronize(LN.states_done));
    const unsigned mask = convert_states(states, n_states, LN.h_states);
    CK(cudaMemcpyAsync(LN.d_states, LN.h_states, n_states * sizeof(DevState), cudaMemcpyHostToDevice, LN.stream));
    CK(cudaEventRecord(LN.states_done, LN.stream));
    g.bytes_h2d += n_states * sizeof(DevState) + (size_t)n_tris * sizeof(pfcu_triangle);

    /* triangles: one DMA from pinned memory, or staged through our own pinned buffer */
    const size_t bytes = (size_t)n_tris * sizeof(pfcu_triangle);
    PinnedBlock *pb = find_pinned(tris);
    if (pb) {
        CK(cudaMemcpyAsync(LN.d_tris, tris, bytes, cudaMemcpyHostToDevice, LN.stream));
        CK(cudaEventRecord(pb->done, LN.stream));
        pb->pending = true;
    } else {
        if (bytes > LN.cap_stage) {
            CK(cudaEventSynchronize(LN.stage_done));
            if (LN.h_stage) cudaFreeHost(LN.h_stage);
            size_t c = LN.cap_stage ? LN.cap_stage : (1u << 20); while (c < bytes) c *= 2;
            CK(cudaHostAlloc(&LN.h_stage, c, cudaHostAllocDefault));
            LN.cap_stage = c;
        } else CK(cudaEventSynchronize(LN.stage_done));
        memcpy(LN.h_stage, tris, bytes);
        CK(cudaMemcpyAsync(LN.d_tris, LN.h_stage, bytes, cudaMemcpyHostToDevice, LN.stream));
        CK(cudaEventRecord(LN.stage_done, LN.stream));
    }
    return launch_pipeline(s, LN.d_tris, LN.d_states, n_tris, mask, g_last_single_prog);
}

pfcu_batch *pfcu_batch_upload(const pfcu_state *states, uint32_t n_states, const pfcu_triangle *tris, uint32_t n_tris)
{
    API_LOCK;
    if (!g.ok || !states || !tris || n_states == 0 || n_tris == 0) return nullptr;
    pfcu_batch *b = (pfcu_batch *)calloc(1, sizeof *b);
    if (!b) return nullptr;
    g.cur = &g.lanes[0];
    std::vector<DevState> tmp(n_states);
    b->feature_mask = convert_states(states, n_states, tmp.data());
    b->single_prog = g_last_single_prog;
    new (&b->deps) std::vector<pfcu_surface *>(g.deps);
    b->n_states = n_states; b->n_tris = n_tris;
    CKP(cudaMalloc(&b->states, n_states * sizeof(DevState)));
    CKP(cudaMalloc(&b->tris, (size_t)n_tris * sizeof(pfcu_triangle)));
    CKP(cudaMemcpyAsync(b->states, tmp.data(), n_states * sizeof(DevState), cudaMemcpyHostToDevice, LN.stream));
    CKP(cudaMemcpyAsync(b->tris, tris, (size_t)n_tris * sizeof(pfcu_triangle), cudaMemcpyHostToDevice, LN.stream));
    CKP(cudaStreamSynchronize(LN.stream));
    return b;
}

int pfcu_batch_submit(pfcu_surface *s, pfcu_batch *b)
{
    API_LOCK;
    if (!g.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || !b) return PFCU_ERR_INVALID;
    use_lane(s);
    g.deps.clear();
    for (pfcu_surface *dep : b->deps) g.deps.push_back(dep);
    return launch_pipeline(s, b->tris, b->states, b->n_tris, b->feature_mask, b->single_prog);
}

void pfcu_batch_destroy(pfcu_batch *b)
{
    API_LOCK;
    if (!b) return;
    if (g.ok) sync_all_lanes();
    cudaFree(b->states); cudaFree(b->tris); b->deps.~vector(); free(b);
}

void pfcu_profile_enable(int on) { g.profiling = on != 0; }
void pfcu_set_raster_path(int path) { API_LOCK; g.raster_path = (path == PFCU_RASTER_TILES || path == PFCU_RASTER_FRAGMENTS) ? path : PFCU_RASTER_AUTO; }

int pfcu_profile_read(pfcu_profile *out)
{
    API_LOCK;
    if (!g.ok) return PFCU_ERR_NO_DEVICE;
    sync_all_lanes();
    out->raster_ms = 0; out->frontend_ms = 0; out->raster_launches = 0;
    for (size_t i = 0; i + 2 < g.prof_events.size(); i += 3) {
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, g.prof_events[i], g.prof_events[i + 1]));
        CK(cudaEventElapsedTime(&b, g.prof_events[i + 1], g.prof_events[i + 2]));
        out->frontend_ms += a; out->raster_ms += b; out->raster_launches++;
    }
    for (auto e : g.prof_events) g.prof_pool.push_back(e);
    g.prof_events.clear();
    return PFCU_OK;
}

int pfcu_finish(void)
{
    API_LOCK;
    if (!g.ok) return PFCU_ERR_NO_DEVICE;
    for (int i = 0; i < g.n_lanes; i++) CK(cudaStreamSynchronize(g.lanes[i].stream));
    return PFCU_OK;
}

int pfcu_get_counters(pfcu_counters *out)
{
    API_LOCK;
    if (!g.ok) return PFCU_ERR_NO_DEVICE;
    unsigned long long h[4] = { 0, 0, 0, 0 };
    sync_all_lanes();
    CK(cudaMemcpy(h, g.d_counters, sizeof h, cudaMemcpyDeviceToHost));
    out->triangles_submitted = g.submitted + h[3];
    out->triangles_rasterised = h[0];
    out->pixels_shaded = h[1];
    out->pixels_depth_failed = h[2];
    out->kernel_launches = g.launches;
    out->bytes_h2d = g.bytes_h2d; out->bytes_d2h = g.bytes_d2h;
    return PFCU_OK;
}

void pfcu_reset_counters(void)
{
    API_LOCK;
    if (!g.ok) return;
    sync_all_lanes();
    cudaMemset(g.d_counters, 0, 4 * sizeof(unsigned long long));
    g.submitted = 0; g.launches = 0; g.bytes_h2d = 0; g.bytes_d2h = 0;
}

} /* extern "C" */
