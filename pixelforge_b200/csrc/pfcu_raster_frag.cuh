/* pfcu_raster_frag.cuh - k_raster_frag: fragment-compacting rasteriser for batches of many small triangles.
 * Part of the single translation unit pfcu.cu (included there, in order; not a stand-alone header). */

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
#define RSQ_SMEM_BITS 10        /* RSQRTPS tables of up to 2 x 2^10 entries are copied to shared memory by the Phong kernel */

struct FragCtx {
    unsigned col_base;                  /* shared-window byte address of this warp's colour region     */
    unsigned rcp_base;                  /* ... of the shared RCPPS table                                */
    int rcp_shift; bool rcp_shared;     /* rcp_shared: table copy in shared memory at rcp_base                  */
    bool rcp_global;                    /* else, when the table has <= 2^11 entries: fast path through L1         */
    int RX0, RY0, RX1, RY1;             /* the region on the surface, inclusive                         */
    unsigned shaded, covered;
    unsigned tri_base;                  /* shared-window byte address of this warp's triangle staging   */
    unsigned ring_base;                 /* ... of this warp's ring of 64 covered fragments (16 bits each) */
    unsigned rsq_base;                  /* ... of the shared RSQRTPS table (Phong kernel), or 0 */
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

__device__ __forceinline__ uint2 lds_tri2(unsigned base, int field, int j)     /* the first half of a staged field */
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(base + (unsigned)((field * 32 + j) << 4)));
    return v;
}
__device__ __forceinline__ void sts_u16(unsigned addr, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ unsigned lds_u16(unsigned addr) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr)); return v; }

template <int OFF> __device__ __forceinline__ float lds_f32_off(unsigned addr) { float v; asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF)); return v; }
template <int OFF> __device__ __forceinline__ void sts_f32_off(unsigned addr, float v) { asm volatile("st.shared.f32 [%0+%1], %2;" :: "r"(addr), "n"(OFF), "f"(v) : "memory"); }

__device__ __forceinline__ float rcp_tab(const FragCtx &t, float x)
{
    const unsigned u = __float_as_uint(x), E = u & 0x7f800000u;
    if (!(t.rcp_shared || t.rcp_global) || E - 0x00800000u >= 0x7e000000u) return rcp_x86(x);
    const unsigned idx = (u & 0x007fffffu) >> t.rcp_shift;
    const unsigned tv = t.rcp_shared ? lds_u32(t.rcp_base + (idx << 2)) : __ldg(c_rcp_tab + idx);
    return __uint_as_float((tv + 0x3f800000u - E) | (u & 0x80000000u));
}

/* One group of <= 32 triangles in one state (lane l holds triangle ti with nn candidate pixels in this warp's
 * region; nn == 0 for lanes outside the group).  pk = cx0 | cy0<<4 | cw<<8 | ceil(1024/cw)<<12 describes the
 * clipped rectangle (region-local).
 *
 * Two interleaved phases over the ordered candidate stream (the clipped bounding rectangles of the group laid end to
 * end):
 *   coverage   32 candidates at a time (lane = candidate): owner lookup by binary search over the scan, pixel from the
 *              packed rectangle, the three integer edge functions - nothing else.  The covered ones are appended, in
 *              order, to the warp's ring of 64 packed fragments {triangle lane, x, y} (ballot compaction);
 *   shading    as soon as the ring holds 32 fragments (or the stream ends) they are shaded together, every lane a
 *              COVERED fragment: depth, colour, texture, Phong, ordered read-modify-write.
 * Only about a third (C2) to a half (C3) of the candidates of a small triangle are covered; with the fragment program
 * running on compacted fragments the warp executes it that many times less often.  Same-pixel fragments of one
 * shading step (shared edges and vertices are drawn by every triangle that owns them, Q4) are ranked by
 * __match_any_sync and written in rank order, as before. */
template <int TEXM, int BLENDM, bool PHONG, int DEPTH_OFF>
__device__ __forceinline__ void frag_run(FragCtx &t, const unsigned nn, const unsigned pk,
                                         const DevState *st, const unsigned flags, const unsigned zmask, const int blend_mode, const TexRegs &tex)
{
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

    unsigned base = 0, rc = 0, rh = 0;          /* next candidate; fragments waiting in the ring; ring head (all warp-uniform) */
    for (;;) {
        /* ---- coverage: fill the ring until a full shading step is there ---- */
        while (rc < 32u && base < total) {
            const unsigned f = base + lane;
            const bool valid = f < total;
            unsigned pos = 0;                                       /* owner = number of lanes whose scan value is <= f */
#pragma unroll
            for (int step = 16; step; step >>= 1) { const unsigned v = __shfl_sync(FULL, I, pos + step - 1); if (v <= f) pos += step; }
            const int j = (int)(pos & 31u);                         /* f >= total gives 32: any staged lane will do, the result is dropped */
            const unsigned Ej = __shfl_sync(FULL, Ex, j), pkj = __shfl_sync(FULL, pk, j);
            const unsigned r = valid ? f - Ej : 0u;
            const unsigned cw = (pkj >> 8) & 15u;
            const unsigned ry = (r * (pkj >> 12)) >> 10, rx = r - ry * cw;          /* r / cw, r % cw (r < 64, cw <= 8: exact) */
            const unsigned px = (pkj & 15u) + rx, py = ((pkj >> 4) & 15u) + ry;
            const uint4 f0 = lds_tri(t.tri_base, 0, j), f1 = lds_tri(t.tri_base, 1, j);
            const uint2 f2 = lds_tri2(t.tri_base, 2, j);
            const int w1 = wadd(wadd((int)f0.x, wmul((int)py, (int)f1.y)), wmul((int)px, (int)f1.x));
            const int w2 = wadd(wadd((int)f0.y, wmul((int)py, (int)f1.w)), wmul((int)px, (int)f1.z));
            const int w3 = wadd(wadd((int)f0.z, wmul((int)py, (int)f2.y)), wmul((int)px, (int)f2.x));
            const bool c = valid && ((w1 | w2 | w3) > 0);
            const unsigned bal = __ballot_sync(FULL, c);
            if (c) sts_u16(t.ring_base + (((rh + rc + __popc(bal & lt)) & 63u) << 1), (unsigned)j | (px << 5) | (py << 8));
            rc += __popc(bal);
            base += 32u;
        }
        if (rc == 0u) break;                                        /* stream exhausted, ring empty */
        __syncwarp();                                               /* ring writes -> ring reads by other lanes */
        const unsigned ns = min(rc, 32u);
        bool m = lane < ns;
        const unsigned ent = lds_u16(t.ring_base + (((rh + (m ? lane : 0u)) & 63u) << 1));
        rh = (rh + ns) & 63u; rc -= ns;
        const int j = (int)(ent & 31u), px = (int)((ent >> 5) & 7u), py = (int)(ent >> 8);
        t.covered += m ? 1u : 0u;

        const uint4 f0 = lds_tri(t.tri_base, 0, j), f1 = lds_tri(t.tri_base, 1, j), f2 = lds_tri(t.tri_base, 2, j);
        const int w1 = wadd(wadd((int)f0.x, wmul(py, (int)f1.y)), wmul(px, (int)f1.x));
        const int w2 = wadd(wadd((int)f0.y, wmul(py, (int)f1.w)), wmul(px, (int)f1.z));
        const int w3 = wadd(wadd((int)f0.z, wmul(py, (int)f2.y)), wmul(px, (int)f2.x));
        const uint4 f3 = lds_tri(t.tri_base, 3, j);
        const unsigned meta = f3.y;
        const float invSum = __uint_as_float(f0.w);
        const float W1 = FM(__int2float_rn(w1), invSum);
        const float W2 = FM(__int2float_rn(w2), invSum);
        const float W3 = FM(__int2float_rn(w3), invSum);
        const float zsum = FA(FA(FM(__uint_as_float(f2.z), W1), FM(__uint_as_float(f2.w), W2)), FM(__uint_as_float(f3.x), W3));
        const float z = rcp_tab(t, zsum);

        /* same-pixel fragments of this step (different triangles) must be applied in triangle order */
        const unsigned sa = t.col_base + (unsigned)(((py << 3) + px) << 2);
        const unsigned peers = __match_any_sync(FULL, m ? sa : (0x80000000u | lane));
        const unsigned rank = __popc(peers & lt);
        const unsigned nr = __reduce_max_sync(FULL, m ? rank : 0u);
        if (ztest && nr == 0u) {                        /* no conflicts: test before shading, like the reference's early mask */
            /* idle lanes do not read: their pixel may be another lane's fragment, written below */
            float zb = 0.0f;
            if (m) zb = lds_f32_off<DEPTH_OFF>(sa);
            m = m && depth_pass_mask(z, zb, zmask);
            if (!__any_sync(FULL, m)) { __syncwarp(); continue; }       /* the depth reads above precede later steps' stores */
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
            /* depth-failed lanes fetch nothing: their texels would be thrown away */
            unsigned texel = 0u;
            if (m) {
                if (TEXM == 1) {
                    const float fu = FM(FS(u, truncf(u)), tex.wm1), fv = FM(FS(v, truncf(v)), tex.hm1);
                    const int xi = cvt_rne_x86(fabsf(fu)), yi = cvt_rne_x86(fabsf(fv));
                    const unsigned off = (unsigned)yi * tex.tw + (unsigned)xi;
                    if (off < tex.total) texel = __ldg((const unsigned *)tex.base + off);
                } else texel = tex_sample<false>(tex, st, u, v);
            }
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
                frag = px_split(phong(px_join(frag), st, (meta >> 24) & 1u, Qx, Qy, Qz, Nx, Ny, Nz, t.rsq_base));
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
            /* orders this round's stores before the next round's (and the next step's) loads and stores of the same
               pixel by OTHER lanes: lanes are not pinned to pixels here, so program order alone does not cover it */
            __syncwarp();
        }
    }
}

/* per-warp shading state that persists across groups: the state snapshot in force */
struct GroupState {
    unsigned cur_state; const DevState *st; unsigned flags, zmask; int blend_mode, prog; TexRegs tex;
    __device__ __forceinline__ void reset()
    {
        cur_state = 0xffffffffu; st = nullptr; flags = 0; zmask = 8u; blend_mode = 0; prog = 0;
        tex.base = nullptr; tex.tw = tex.th = tex.total = 0; tex.wm1 = tex.hm1 = 0.0f; tex.fmt = tex.wrap = tex.filter = 0;
    }
};

/* One group: lane l < cnt holds triangle ti.  Stages the constants, clips to the region, splits the group into runs
 * of one state and shades them (frag_run). */
template <bool HAS_PHONG, int DEPTH_OFF>
__device__ __forceinline__ void frag_group(FragCtx &t, GroupState &G, const int4 *__restrict__ bbox_, const TriSetup *__restrict__ setup_,
                                           const TriData *__restrict__ data_, const DevState *__restrict__ states_,
                                           const unsigned ti, const bool have, const unsigned cnt)
{
    const int lane = threadIdx.x & 31;
    unsigned &cur_state = G.cur_state; const DevState *&st = G.st; unsigned &flags = G.flags, &zmask = G.zmask;
    int &blend_mode = G.blend_mode, &prog = G.prog; TexRegs &tex = G.tex;
    unsigned state = 0xffffffffu, nn0 = 0, pk = 0;
    if (have) {
        /* stage this triangle's constants (every load is independent: one memory round trip per group) */
        const int4 b = __ldg(bbox_ + ti);
        const uint4 s0 = __ldg(reinterpret_cast<const uint4 *>(setup_ + ti));
        const uint4 s1 = __ldg(reinterpret_cast<const uint4 *>(setup_ + ti) + 1);
        const uint4 s2 = __ldg(reinterpret_cast<const uint4 *>(setup_ + ti) + 2);
        const uint4 *da = reinterpret_cast<const uint4 *>(data_ + ti);
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
            st = states_ + cur_state;
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
        case 0:  frag_run<0, 0, false, DEPTH_OFF>(t, nn, pk, st, flags, zmask, blend_mode, tex); break;
        case 1:  frag_run<0, 1, false, DEPTH_OFF>(t, nn, pk, st, flags, zmask, blend_mode, tex); break;
        case 2:  frag_run<0, 2, false, DEPTH_OFF>(t, nn, pk, st, flags, zmask, blend_mode, tex); break;
        case 3:  frag_run<0, 3, false, DEPTH_OFF>(t, nn, pk, st, flags, zmask, blend_mode, tex); break;
        case 4:  frag_run<1, 0, false, DEPTH_OFF>(t, nn, pk, st, flags, zmask, blend_mode, tex); break;
        case 5:  frag_run<1, 1, false, DEPTH_OFF>(t, nn, pk, st, flags, zmask, blend_mode, tex); break;
        case 6:  frag_run<1, 2, false, DEPTH_OFF>(t, nn, pk, st, flags, zmask, blend_mode, tex); break;
        case 7:  frag_run<1, 3, false, DEPTH_OFF>(t, nn, pk, st, flags, zmask, blend_mode, tex); break;
        case 8:  frag_run<2, 0, false, DEPTH_OFF>(t, nn, pk, st, flags, zmask, blend_mode, tex); break;
        case 9:  frag_run<2, 1, false, DEPTH_OFF>(t, nn, pk, st, flags, zmask, blend_mode, tex); break;
        case 10: frag_run<2, 2, false, DEPTH_OFF>(t, nn, pk, st, flags, zmask, blend_mode, tex); break;
        case 11: frag_run<2, 3, false, DEPTH_OFF>(t, nn, pk, st, flags, zmask, blend_mode, tex); break;
        default: if (HAS_PHONG) frag_run<2, 3, true, DEPTH_OFF>(t, nn, pk, st, flags, zmask, blend_mode, tex); break;
        }
        lo = hi;
    }
}

template <bool HAS_PHONG, int NW>
__device__ __forceinline__ void raster_frag_body(const RasterParams &p)
{
    pdl_wait();         /* launched as a programmatic dependent of the binning kernels; before ANY exit, so that the grid cannot complete early */
    constexpr int NT = NW * 32, TH = NW, SUB = TILE / TH;
    __shared__ __align__(16) unsigned s_tile[2 * NW * FRAG_RSTRIDE];     /* colour regions, then depth regions */
    unsigned *const s_col = s_tile;
    float *const s_dep = reinterpret_cast<float *>(s_tile + NW * FRAG_RSTRIDE);
    __shared__ unsigned s_rcp[1 << RCP_SMEM_BITS];
    __shared__ unsigned s_queue[QUEUE_CAP];
    __shared__ unsigned short s_qmask[QUEUE_CAP];
    __shared__ unsigned s_wcount[NW];
    __shared__ unsigned s_group[NW][32];
    __shared__ unsigned short s_ring[NW][64];
    __shared__ unsigned s_rsq[HAS_PHONG ? (2 << RSQ_SMEM_BITS) : 1];
    extern __shared__ __align__(16) uint4 s_tri[];          /* [NW][NF][32] triangle staging, see FRAG_NF */
    constexpr int NF = HAS_PHONG ? FRAG_NF_PHONG : FRAG_NF;
    static_assert(NW == 8 || NW == 16, "one 8x8 region per warp, 8 regions per row");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tile = (p.world > 1) ? (p.rank + (p.tile_base + blockIdx.x / SUB) * p.world) : (p.tile_base + blockIdx.x / SUB);
    if (tile >= p.nTiles) return;
    const int tx = tile % p.tilesX, ty = tile / p.tilesX;
    const int X0 = tx * TILE, Y0 = ty * TILE + (int)(blockIdx.x % SUB) * TH;
    if (Y0 >= p.H) return;
    const int X1 = min(X0 + TILE, p.W) - 1, Y1 = min(Y0 + TH, p.H) - 1;
    const bool full_tile = (X0 + TILE <= p.W) && (Y0 + TH <= p.H) && ((p.W & 3) == 0);

    const int bin = (Y0 >> p.bsy) * p.binsX + (X0 >> p.bsx);
    const bool overflow = p.bin_starts[p.nb] > p.list_cap;      /* lists not written: filter the whole batch (see RasterParams) */
    const unsigned lbeg = overflow ? 0u : p.bin_starts[bin], lend = overflow ? p.n : p.bin_starts[bin + 1];
    if (lbeg == lend) return;

    FragCtx t;
    t.rcp_shift = c_rcp_shift;
    t.rcp_shared = t.rcp_shift >= 23 - RCP_SMEM_BITS; t.rcp_global = false;
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
    t.rsq_base = 0u;
    if (HAS_PHONG && c_rsq_bits <= RSQ_SMEM_BITS) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(s_rsq);
        for (int k = tid; k < (2 << c_rsq_bits) / 4; k += NT)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst + (unsigned)(k << 4)), "l"(c_rsq_tab + 4 * k) : "memory");
        t.rsq_base = dst;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    t.col_base = (unsigned)__cvta_generic_to_shared(s_col + warp * FRAG_RSTRIDE);
    t.rcp_base = (unsigned)__cvta_generic_to_shared(s_rcp);
    t.RX0 = X0 + (warp & 7) * 8; t.RY0 = Y0 + (warp >> 3) * 8;
    t.RX1 = min(t.RX0 + 7, X1); t.RY1 = min(t.RY0 + 7, Y1);
    t.shaded = 0; t.covered = 0;
    t.tri_base = (unsigned)__cvta_generic_to_shared(s_tri + warp * NF * 32);
    t.ring_base = (unsigned)__cvta_generic_to_shared(&s_ring[warp][0]);

    bool loaded = false;
    /* the slice relative to its bin, as the bin-list entries store their rectangles */
    const int bx0 = X0 & ((1 << p.bsx) - 1), by0 = Y0 & ((1 << p.bsy) - 1);
    const int bx1 = bx0 + (X1 - X0), by1 = by0 + (Y1 - Y0);

    for (unsigned base = lbeg; base < lend; ) {
        /* ---- fill the queue: ordered compaction of the bin list against this slice ---- */
        unsigned qn = 0;
        while (base < lend && qn + NT <= QUEUE_CAP) {
            const unsigned k = base + tid;
            bool hit = false; unsigned ti = 0, wmask = 0;
            if (k < lend) {
                uint2 e;
                if (!overflow) e = __ldg(p.bin_list + k);
                else {                                  /* the entry k_bin_fill would have written for this bin, or a miss */
                    const int4 b = __ldg(p.bbox + k);
                    const int ox = (X0 >> p.bsx) << p.bsx, oy = (Y0 >> p.bsy) << p.bsy;
                    const bool in_bin = b.x < b.z && b.x < ox + (1 << p.bsx) && b.z - 1 >= ox && b.y < oy + (1 << p.bsy) && b.w >= oy;
                    e = in_bin ? make_uint2(k, bin_rel_bbox(b, X0 >> p.bsx, Y0 >> p.bsy, p.bsx, p.bsy)) : make_uint2(k, 0x0000ffffu);   /* x0 = y0 = 255 > x1 = y1 = 0: never hits */
                }
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
        GroupState G; G.reset();
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
                    frag_group<HAS_PHONG, NW * FRAG_RSTRIDE * 4>(t, G, p.bbox, p.setup, p.data, p.states, ti, have, cnt);
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

template <bool HAS_PHONG, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
k_raster_frag(const RasterParams p)
{
    raster_frag_body<HAS_PHONG, NW>(p);
}
