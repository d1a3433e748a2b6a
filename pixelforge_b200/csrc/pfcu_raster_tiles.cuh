/* pfcu_raster_tiles.cuh - k_raster: tile rasteriser for batches of large triangles (one triangle per warp step over 8x4 blocks).
 * Part of the single translation unit pfcu.cu (included there, in order; not a stand-alone header). */

/* ------------------------------------------------------------------------------------------------ */
/* tile rasteriser: parameters, shared-tile addressing, packed colour arithmetic                    */
/* ------------------------------------------------------------------------------------------------ */

struct RasterParams {
    const int4 *bbox; const TriSetup *setup; const TriData *data; const DevState *states;
    const uint2 *bin_list; const unsigned *bin_starts; int binsX; int bsx, bsy;      /* a bin is 2^bsx x 2^bsy pixels (k_raster: bsy == bsx >= 6) */
    uint32_t *color; float *depth; int W, H; int tilesX, tilesY;
    unsigned rank, world; unsigned nTiles;
    unsigned long long *counters;
    /* bin_starts[nb] is the real number of list entries; when it exceeds list_cap the lists were not written (the host
       sizes them without waiting for the total) and every CTA filters all n triangles of the batch instead */
    int nb; unsigned list_cap, n;
    /* first tile of this launch: the rasterisation of a large surface goes out as a few launches over bands of tile rows
       (on separate streams, they overlap) so that the read-back of a band can start when ITS launch is done */
    unsigned tile_base;
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
#ifndef FILTER_REFINE
#define FILTER_REFINE 0      /* per-warp edge-function test in the queue filter: measured, costs 1 % on the 4K / 8K scenes and gains nothing (the per-row vote in shade_tri already rejects cheaply) */
#endif

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
 * 2 any sampler, 3 nearest with any wrap mode / texel layout, 4 bilinear with any wrap mode / texel layout (3 and 4
 * exist so that a batch in ONE such program gets a kernel without the other filter's code and registers).
 * BLENDM: 0 off, 1 ALPHA, 2 ADD, 3 any mode.  Everything is computed for all 32 lanes
 * (no divergent regions); only the final stores are predicated by the coverage/depth mask.
 * BIG: the launch is a batch of large triangles in ONE state program with the RCPPS table in shared memory
 * (launch_pipeline checks both), so some per-block early-outs and run-time checks are dropped. */
template <int TEXM, int BLENDM, bool PHONG, int NW, bool BIG, bool GREY = false>
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
       scalar, see the grey_tex branches below.  A template parameter (the caller tests the triangle with grey_triangle()):
       as a run-time flag the compiler evaluates both colour paths for every fragment and selects. */
    constexpr bool grey_tex = GREY;
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

    /* block ownership: 8 warps -> warp w owns block (bx,by) iff (bx + 3*by) & 7 == w (one block per block row);
       16 warps -> additionally even block rows belong to warps 0..7, odd rows to warps 8..15.
       Which block rows can this warp's block be covered in?  Lane r answers for block row r, all rows at once: the block
       must meet the clipped bbox and - when the int32 edge functions cannot wrap inside the bbox (TF_SAFE) - every edge
       function must reach a non-negative value on the block's part of the bbox (its maximum lies on the corner picked
       by the signs of the edge's steps).  Conservative, so the exact per-pixel test below decides; what it buys is that
       the half of a large triangle's bounding box that lies outside the triangle costs one vote per triangle instead
       of an edge evaluation per block.  Triangles of a few block rows skip the vote and walk their rows. */
    unsigned rows;
    if (by1 - by0 >= 3) {
        const int r = (int)(threadIdx.x & 31u);
        const int rbx = ((t.warp & 7) - 3 * r) & 7;
        const int xlo = max(rbx << 3, cx0), xhi = min((rbx << 3) + 7, cx1), ylo = max(r << 2, cy0), yhi = min((r << 2) + 3, cy1);
        bool ok = xlo <= xhi && ylo <= yhi && (NW != 16 || (((r ^ (t.warp >> 3)) & 1) == 0));
        if (s.flags & TF_SAFE) {
            const int ox = t.X0 - b.x, oy = t.Y0 - b.y;
            const int ax0 = xlo + ox, ax1 = xhi + ox, ay0 = ylo + oy, ay1 = yhi + oy;
            const int m1 = s.w1R + (s.w1X > 0 ? ax1 : ax0) * s.w1X + (s.w1Y > 0 ? ay1 : ay0) * s.w1Y;
            const int m2 = s.w2R + (s.w2X > 0 ? ax1 : ax0) * s.w2X + (s.w2Y > 0 ? ay1 : ay0) * s.w2Y;
            const int m3 = s.w3R + (s.w3X > 0 ? ax1 : ax0) * s.w3X + (s.w3Y > 0 ? ay1 : ay0) * s.w3Y;
            ok = ok && (m1 | m2 | m3) >= 0;
        }
        rows = __ballot_sync(0xffffffffu, ok);
    } else {
        rows = ((2u << by1) - 1u) & ~((1u << by0) - 1u);
        if (NW == 16) rows &= ((t.warp >> 3) & 1) ? 0xaaaaaaaau : 0x55555555u;
    }
    while (rows) {
        const int by = __ffs(rows) - 1; rows &= rows - 1u;
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
            } else if (TEXM == 3) texel = tex_fetch<false>(tex, tex_coord(tex.wrap, u, tex.wm1), tex_coord(tex.wrap, v, tex.hm1));
            else if (TEXM == 4) texel = tex_sample_bilinear<false>(tex, st, u, v);
            else texel = tex_sample<true>(tex, st, u, v);        /* run-time sampler: every texel layout */
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

/* smooth-shaded triangle whose three vertex colours are one grey (all four channels equal): see shade_tri<GREY> */
__device__ __forceinline__ bool grey_triangle(const uint4 a1, unsigned flags)
{
    return (flags & PFCU_ST_SMOOTH) && a1.x == a1.y && a1.y == a1.z && a1.x == (a1.x & 0xffu) * 0x01010101u;
}

/* FIXED_PROG >= 0: the whole batch runs one state program (texm*4 + blendm, texm in {0, 1, 3, 4}: state_program() in
 * pfcu.cu), known at launch; only that variant is instantiated, which lets the register allocator fit 4 CTAs per SM
 * (bilinear: 3).  -1: per-triangle dispatch over the programs texm in {0, 1, 2}. */
template <bool HAS_PHONG, int NW, int FIXED_PROG, int TH>
__global__ void __launch_bounds__(NW * 32, FIXED_PROG >= 0 ? (FIXED_PROG / 4 == 4 ? 3 : 4) : (NW == 16 ? (HAS_PHONG ? 1 : 2) : (HAS_PHONG ? 2 : 3)))
k_raster(const RasterParams p)
{
    pdl_wait();         /* launched as a programmatic dependent of the binning kernels; before ANY exit, so that the grid cannot complete early */
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
    const unsigned tile = (p.world > 1) ? (p.rank + (p.tile_base + blockIdx.x / SUB) * p.world) : (p.tile_base + blockIdx.x / SUB);
    if (tile >= p.nTiles) return;
    const int tx = tile % p.tilesX, ty = tile / p.tilesX;
    TileCtx t;
    t.X0 = tx * TILE; t.Y0 = ty * TILE + (int)(blockIdx.x % SUB) * TH;
    if (t.Y0 >= p.H) return;
    t.X1 = min(t.X0 + TILE, p.W) - 1; t.Y1 = min(t.Y0 + TH, p.H) - 1;
    const int X0 = t.X0, Y0 = t.Y0, X1 = t.X1, Y1 = t.Y1;
    const bool full_tile = (X0 + TILE <= p.W) && (Y0 + TH <= p.H) && ((p.W & 3) == 0);

    const int bin = ((ty * TILE) >> p.bsy) * p.binsX + ((tx * TILE) >> p.bsx);
    const bool overflow = p.bin_starts[p.nb] > p.list_cap;
    const unsigned lbeg = overflow ? 0u : p.bin_starts[bin], lend = overflow ? p.n : p.bin_starts[bin + 1];
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
                ti = overflow ? k : __ldg(&p.bin_list[k].x);
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
                    if (FILTER_REFINE && hit && (s.flags & TF_SAFE) && nbx * nby >= 4) {
                        /* a triangle of many blocks: test every block of the clipped bbox with the edge functions (their
                           maxima over the block's part of the bbox, as for the whole tile above) and list only the warps
                           that own a block the triangle can cover - a warp otherwise pays the triangle's whole prologue
                           to find out that, say, the other half of a screen-filling quad is not its business */
                        const int sx1 = s.w1X > 0, sx2 = s.w2X > 0, sx3 = s.w3X > 0, sy1 = s.w1Y > 0, sy2 = s.w2Y > 0, sy3 = s.w3Y > 0;
                        for (int r = by0; r < by0 + nby; r++) {
                            const int ay0 = max(Y0 + (r << 2), ry0) - b.y, ay1 = min(Y0 + (r << 2) + 3, ry1) - b.y;
                            const int r1 = s.w1R + (sy1 ? ay1 : ay0) * s.w1Y, r2 = s.w2R + (sy2 ? ay1 : ay0) * s.w2Y, r3 = s.w3R + (sy3 ? ay1 : ay0) * s.w3Y;
                            unsigned bits = 0;
                            for (int c = bx0; c < bx0 + nbx; c++) {
                                const int ax0 = max(X0 + (c << 3), rx0) - b.x, ax1 = min(X0 + (c << 3) + 7, rx1) - b.x;
                                const int m1 = r1 + (sx1 ? ax1 : ax0) * s.w1X, m2 = r2 + (sx2 ? ax1 : ax0) * s.w2X, m3 = r3 + (sx3 ? ax1 : ax0) * s.w3X;
                                if ((m1 | m2 | m3) >= 0) bits |= 1u << ((c + 3 * r) & 7);
                            }
                            wmask |= (NW == 16 && (r & 1)) ? (bits << 8) : bits;
                        }
                        if (!wmask) hit = false;
                    } else {
                        const unsigned run = nbx >= 8 ? 0xffu : ((1u << nbx) - 1u);
                        for (int j = 0; j < min(nby, 8); j++) {
                            const int sh = (bx0 + 3 * (by0 + j)) & 7;
                            const unsigned bits = ((run << sh) | (run >> (8 - sh))) & 0xffu;
                            wmask |= (NW == 16 && ((by0 + j) & 1)) ? (bits << 8) : bits;
                        }
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
                    if (FIXED_PROG / 4 != 0 && grey_triangle(a1, flags))
                        shade_tri<FIXED_PROG / 4, FIXED_PROG % 4, false, NW, true, true>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex);
                    else
                        shade_tri<FIXED_PROG / 4, FIXED_PROG % 4, false, NW, true, false>(t, ti, b, s, a0, a1, st, flags, zmask, blend_mode, tex);
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
