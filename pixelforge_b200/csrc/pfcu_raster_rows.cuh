/* pfcu_raster_rows.cuh - k_raster_rows: row-ordered rasteriser for render targets other than RGBA8 and for BGRA8 textures.
 * Part of the single translation unit pfcu.cu (included there, in order; not a stand-alone header). */

/* ------------------------------------------------------------------------------------------------ */
/* kernel: row-ordered rasteriser (the reference's 4-pixel "leader" behaviour)          */
/* ------------------------------------------------------------------------------------------------ */
/*
 * The reference walks a triangle's bounding box in groups of 8 pixels starting at its xMin (triangles.c:404-445).  Its
 * BGRA8 pixel getter and setter swap bytes with _mm256_shuffle_epi8 and the mask {2,1,0,3} repeated (pixel.h:2069-2078,
 * 2915-2920, simd.h:563-583); that instruction indexes bytes inside each 128-bit half, so all four dwords of a half
 * receive the swapped FIRST dword.  Consequences, which are the reference's actual output and therefore the spec:
 *   - BGRA8 texture: the texel of pixels xMin+4k+1..3 is the texel fetched for pixel xMin+4k (each bilinear tap too;
 *     the filter weights stay per pixel).  That "leader" fetches at uv = 0 when ITS OWN mask (coverage, x < xMax,
 *     depth test) is false (triangles.c:510).
 *   - BGRA8 render target: the blend destination of the four pixels is the leader's pixel, and all four receive the
 *     leader's final fragment (written where their own mask is set).  The leader's fragment is computed whether it is
 *     covered or not, with whatever its barycentric weights give.
 * A pixel's result thus depends on the state (colour, depth) of a pixel up to three columns to its left at the moment
 * the triangle is drawn, which the tile-owning rasterisers cannot know across a tile boundary.  This kernel gives every
 * surface ROW to one warp instead: a CTA takes ROWS_NW consecutive rows, filters the batch's triangles against its
 * band in submission order (ballot compaction into a shared queue), and each warp walks the queue over its row, 32
 * consecutive pixels from xMin per step, lane l = pixel xMin + 32k + l, so a pixel's leader is lane l & ~3 of the same
 * step and is read with a shuffle.  Every lane runs the whole fragment program (as the AVX2 lanes do), in the plain
 * per-channel form of pfcu_device_math.cuh, so uncovered leaders produce exactly the reference's values.
 * RGB8 / BGR8 render targets have no such quirk (their setters are scalar, pixel.h:1590-1688) but store no alpha and
 * read it back as 255; they take this kernel too (colour is kept as canonical RGBA8 with alpha forced to 255).
 * Not a throughput path: none of BASELINE.json's configurations uses these formats.  Surface memory is accessed
 * straight from L2 (ld/st .cg: another lane of the warp wrote the pixel for the previous triangle).
 */
#define ROWS_NW     4
#define ROWS_QUEUE  1024

struct RowsParams {
    const int4 *bbox; const TriSetup *setup; const TriData *data; const DevState *states;
    unsigned n; const unsigned *d_n;
    uint32_t *color; float *depth; int W, H;
    int fb_fmt;                             /* PFCU_TEX_* of the render target */
    unsigned long long *counters;
};

/* pfiColorBarySmooth_simd (color.h:153-181) for ANY weights: wrapping 32-bit products, logical shift, channels OR-ed
 * together without masking */
__device__ __forceinline__ unsigned color_smooth_any(unsigned c1, unsigned c2, unsigned c3, float W1, float W2, float W3)
{
    const unsigned u1 = (unsigned)cvt_rne_x86(FM(W1, 255.0f)), u2 = (unsigned)cvt_rne_x86(FM(W2, 255.0f)), u3 = (unsigned)cvt_rne_x86(FM(W3, 255.0f));
    unsigned p = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const unsigned s = u1 * (unsigned)CHN(c1, i) + u2 * (unsigned)CHN(c2, i) + u3 * (unsigned)CHN(c3, i);
        p |= ((s * 257u) >> 16) << (8 * i);
    }
    return p;
}

/* texel fetch through the reference's getter: `leader` >= 0 selects the BGRA8 getter's behaviour */
__device__ __forceinline__ unsigned tex_fetch_q(const TexRegs &t, int x, int y, int leader)
{
    unsigned v = tex_fetch<true>(t, x, y);
    if (leader >= 0) v = __shfl_sync(0xffffffffu, v, leader);
    return v;
}

__device__ __forceinline__ unsigned tex_sample_q(const TexRegs &t, const DevState *st, float u, float v, int leader)
{
    const int x0 = tex_coord(t.wrap, u, t.wm1), y0 = tex_coord(t.wrap, v, t.hm1);
    if (t.filter == 0) return tex_fetch_q(t, x0, y0, leader);
    const float4 k = __ldg(reinterpret_cast<const float4 *>(&st->tex_fw));
    const int x1 = tex_coord(t.wrap, FA(u, k.z), t.wm1), y1 = tex_coord(t.wrap, FA(v, k.w), t.hm1);
    const float fx = clamp_x86(FS(FM(u, k.x), __int2float_rn(x0)), 0.0f, 1.0f);
    const float fy = clamp_x86(FS(FM(v, k.y), __int2float_rn(y0)), 0.0f, 1.0f);
    const unsigned c00 = tex_fetch_q(t, x0, y0, leader), c10 = tex_fetch_q(t, x1, y0, leader);
    const unsigned c01 = tex_fetch_q(t, x0, y1, leader), c11 = tex_fetch_q(t, x1, y1, leader);
    return color_lerp(color_lerp(c00, c10, fx), color_lerp(c01, c11, fx), fy);
}

template <bool HAS_PHONG>
__global__ void __launch_bounds__(ROWS_NW * 32)
k_raster_rows(const RowsParams p)
{
    constexpr int NT = ROWS_NW * 32;
    __shared__ unsigned s_queue[ROWS_QUEUE];
    __shared__ unsigned s_wcount[ROWS_NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Y0 = (int)blockIdx.x * ROWS_NW, Y1 = min(Y0 + ROWS_NW, p.H) - 1;
    const int row = Y0 + warp;
    const unsigned n = p.d_n ? min(*p.d_n, (unsigned)(FRONT_SMALL_MAX * FRONT_SMALL_CHUNKS)) : p.n;
    const int leader = lane & ~3;
    const bool fb_leader = p.fb_fmt == PFCU_TEX_BGRA8;
    const unsigned alpha_or = (p.fb_fmt >= PFCU_TEX_RGB8) ? 0xff000000u : 0u;
    unsigned shaded = 0, zfailed = 0;

    unsigned cur_state = 0xffffffffu;
    const DevState *st = nullptr;
    unsigned flags = 0; int blend_mode = 0, depth_func = 0;
    TexRegs tex; tex.base = nullptr; tex.tw = tex.th = tex.total = 0; tex.wm1 = tex.hm1 = 0.0f; tex.fmt = tex.wrap = tex.filter = 0;
    bool tex_leader = false;

    for (unsigned base = 0; base < n; ) {
        /* ---- ordered compaction of the batch against this band of rows ---- */
        unsigned qn = 0;
        while (base < n && qn + NT <= ROWS_QUEUE) {
            const unsigned i = base + tid;
            bool hit = false;
            if (i < n) {
                const int4 b = __ldg(p.bbox + i);
                hit = b.x < b.z && b.y <= Y1 && b.w >= Y0;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_wcount[warp] = __popc(bal);
            __syncthreads();
            unsigned woff = 0, total = 0;
#pragma unroll
            for (int w = 0; w < ROWS_NW; w++) { const unsigned c = s_wcount[w]; if (w < warp) woff += c; total += c; }
            if (hit) s_queue[qn + woff + __popc(bal & ((1u << lane) - 1u))] = i;
            qn += total;
            base += NT;
            __syncthreads();
        }

        /* ---- every warp walks the queue over its own row ---- */
        if (row <= Y1) {
            for (unsigned q = 0; q < qn; q++) {
                const unsigned ti = s_queue[q];
                const int4 b = __ldg(p.bbox + ti);
                if (row < b.y || row > b.w) continue;
                const TriSetup s = p.setup[ti];
                const uint4 *da = reinterpret_cast<const uint4 *>(p.data + ti);
                const uint4 a0 = __ldg(da), a1 = __ldg(da + 1), a2 = __ldg(da + 2), a3 = __ldg(da + 3);
                const unsigned meta = a0.w;
                if ((meta & 0xffffffu) != cur_state) {
                    cur_state = meta & 0xffffffu;
                    st = p.states + cur_state;
                    flags = st->flags; blend_mode = st->blend_mode; depth_func = st->depth_func;
                    if (flags & PFCU_ST_TEXTURE) {
                        tex.base = st->tex; tex.tw = st->tw; tex.th = st->th; tex.total = st->tw * st->th;
                        tex.wm1 = __uint2float_rn(st->tw - 1u); tex.hm1 = __uint2float_rn(st->th - 1u);
                        tex.fmt = st->tfmt; tex.wrap = st->tex_wrap; tex.filter = st->tex_filter;
                        tex_leader = st->tex_leader != 0;
                    }
                }
                const bool ztest = (flags & PFCU_ST_DEPTH_TEST) != 0, smooth = (flags & PFCU_ST_SMOOTH) != 0;
                const bool texturing = (flags & PFCU_ST_TEXTURE) != 0, blending = (flags & PFCU_ST_BLEND) != 0;
                const bool phong_on = HAS_PHONG && (flags & PFCU_ST_PHONG) != 0;
                const bool is3d = (meta >> 25) & 1u;
                const int yrel = wsub(row, b.y);
                /* first 32-pixel step that reaches the surface; steps stay aligned to xMin (so do the groups of four) */
                const long long skip = b.x < 0 ? (-(long long)b.x) & ~31LL : 0LL;
                const long long xlast = min((long long)b.z, (long long)p.W - 1);            /* the reference starts groups while x <= xMax */
                for (long long xs = (long long)b.x + skip; xs <= xlast; xs += 32) {
                    const unsigned xrel = (unsigned)(xs - (long long)b.x) + (unsigned)lane;
                    const long long xl = xs + lane;
                    const bool ins = xl >= 0 && xl < (long long)p.W;
                    const int w1 = wadd(wadd(s.w1R, wmul(yrel, s.w1Y)), wmul((int)xrel, s.w1X));
                    const int w2 = wadd(wadd(s.w2R, wmul(yrel, s.w2Y)), wmul((int)xrel, s.w2X));
                    const int w3 = wadd(wadd(s.w3R, wmul(yrel, s.w3Y)), wmul((int)xrel, s.w3X));
                    const bool cov = ((w1 | w2 | w3) > 0) && xl < (long long)b.z && ins;
                    const float W1 = FM(__int2float_rn(w1), s.invSum), W2 = FM(__int2float_rn(w2), s.invSum), W3 = FM(__int2float_rn(w3), s.invSum);
                    const float z = rcp_x86(FA(FA(FM(__uint_as_float(a0.x), W1), FM(__uint_as_float(a0.y), W2)), FM(__uint_as_float(a0.z), W3)));
                    const size_t idx = (size_t)row * (size_t)p.W + (size_t)(ins ? xl : 0);
                    const float zb = ins ? __ldcg(p.depth + idx) : 0.0f;
                    const bool m = cov && (!ztest || depth_pass(depth_func, z, zb));

                    unsigned frag = smooth ? color_smooth_any(a1.x, a1.y, a1.z, W1, W2, W3) : 0u;
                    if (!smooth) {
                        const float mx = max_x86(W1, max_x86(W2, W3));
                        frag = ((mx == W1) ? a1.x : 0u) | ((mx == W2) ? a1.y : 0u) | ((mx == W3) ? a1.z : 0u);
                    }
                    if (texturing) {
                        float u = FA(FA(FM(__uint_as_float(a2.x), W1), FM(__uint_as_float(a2.y), W2)), FM(__uint_as_float(a2.z), W3));
                        float v = FA(FA(FM(__uint_as_float(a3.x), W1), FM(__uint_as_float(a3.y), W2)), FM(__uint_as_float(a3.z), W3));
                        if (is3d) { u = FM(u, z); v = FM(v, z); }
                        if (!m) { u = 0.0f; v = 0.0f; }
                        frag = mul_color(tex_sample_q(tex, st, u, v, tex_leader ? leader : -1), frag);
                    }
                    if (HAS_PHONG) {
                        if (phong_on) {
                            const float4 *a = reinterpret_cast<const float4 *>(p.data + ti) + 4;
                            const float4 px = __ldg(a), py = __ldg(a + 1), pz = __ldg(a + 2);
                            const float4 nx = __ldg(a + 3), ny = __ldg(a + 4), nz = __ldg(a + 5);
                            const float Nx = FA(FA(FM(nx.x, W1), FM(nx.y, W2)), FM(nx.z, W3));
                            const float Ny = FA(FA(FM(ny.x, W1), FM(ny.y, W2)), FM(ny.z, W3));
                            const float Nz = FA(FA(FM(nz.x, W1), FM(nz.y, W2)), FM(nz.z, W3));
                            const float Px = FA(FA(FM(px.x, W1), FM(px.y, W2)), FM(px.z, W3));
                            const float Py = FA(FA(FM(py.x, W1), FM(py.y, W2)), FM(py.z, W3));
                            const float Pz = FA(FA(FM(pz.x, W1), FM(pz.y, W2)), FM(pz.z, W3));
                            frag = phong(frag, st, (meta >> 24) & 1u, Px, Py, Pz, Nx, Ny, Nz);
                        }
                    }
                    if (blending) {
                        unsigned dst = ins ? __ldcg(p.color + idx) : 0u;
                        if (fb_leader) dst = __shfl_sync(0xffffffffu, dst, leader);
                        frag = blend_px(blend_mode, frag, dst);
                    }
                    if (fb_leader) frag = __shfl_sync(0xffffffffu, frag, leader);
                    if (m) {
                        __stcg(p.color + idx, frag | alpha_or);
                        __stcg(p.depth + idx, z);                   /* written even with the depth test off (Q11) */
                        shaded++;
                    } else if (cov) zfailed++;
                    __syncwarp();           /* the next step / triangle maps these pixels to other lanes */
                }
            }
        }
        __syncthreads();
    }
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
