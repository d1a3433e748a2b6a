/* pfcu_vertex_prims.cuh - kernels: device vertex stage (vertex arrays, raw triangles), points and lines, scans.
 * Part of the single translation unit pfcu.cu (included there, in order; not a stand-alone header). */

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

/* largest 32-bit index of an index buffer (how many vertices a pfDrawElements call references): scanning 3 M
 * indices costs the host ~1 ms even with AVX2, the device a few microseconds once they are uploaded anyway */
__global__ void __launch_bounds__(256)
k_index_max(const unsigned *__restrict__ idx, unsigned n, unsigned *__restrict__ out)
{
    unsigned m = 0;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u) m = max(m, __ldg(idx + i));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31u) == 0 && m) atomicMax(out, m);
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

/* the reference's per-triangle prologue + clipping for one unprocessed triangle in prologue environment e; the vertices
   carry colour `rgba` instead of their own when override is set (render-list replay under PF_COLOR_MATERIAL) */
__device__ __forceinline__ int raw_process_tri(const pfcu_rawtri *t, const pfcu_vparams_lit *e, const float *pow_tables, bool override, unsigned rgba,
                                               pfv_vertex *poly, int *is3d, int *face_out)
{
    const int face = t->face;
    *face_out = face;
    for (int k = 0; k < 3; k++) {
        const pfcu_rawvertex *r = &t->v[k];
        pfv_vertex *v = &poly[k];
        for (int j = 0; j < 4; j++) v->position[j] = r->pos[j];
        for (int j = 0; j < 3; j++) v->normal[j] = r->normal[j];
        v->texcoord[0] = r->uv[0]; v->texcoord[1] = r->uv[1];
        v->color = override ? rgba : r->rgba;
        v->screen[0] = 0.0f; v->screen[1] = 0.0f;
        for (int j = 0; j < 4; j++) v->homogeneous[j] = 0.0f;
        if (e->base.lighting) pfv_prologue_lit(e, pow_tables, face, v);
    }
    int n = 3;
    *is3d = pfv_project_and_clip(&e->base, poly, &n);
    return n >= 3 ? n - 2 : 0;
}

__device__ __forceinline__ int raw_process(const RawArgs &a, unsigned i, pfv_vertex *poly, int *is3d, int *face_out, unsigned *state)
{
    const pfcu_rawtri *t = a.tris + i;
    *state = t->state;
    return raw_process_tri(t, a.vp + t->vparams, a.pow_tables, false, 0u, poly, is3d, face_out);
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
struct PrimParams { const pfcu_prim *prims; unsigned n; uint32_t *color; float *depth; unsigned W, H; int tilesX; unsigned rank, world, nTiles;
                    unsigned alpha_or; /* 0xff000000 for RGB8 / BGR8 targets: no stored alpha, reads back as 255 */ };

__device__ __forceinline__ void prim_pixel(const PrimParams &p, const pfcu_prim &pr, int X0, int Y0, uint32_t off, float z, uint32_t color, bool test)
{
    if (off >= p.W * p.H) return;
    const int x = (int)(off % p.W), y = (int)(off / p.W);
    if (x < X0 || x >= X0 + TILE || y < Y0 || y >= Y0 + TILE) return;
    if (test && !pfp_depth(pr.depth_func, z, p.depth[off])) return;
    p.color[off] = ((pr.flags & PFCU_ST_BLEND) ? pfp_blend(pr.blend_mode, color, p.color[off]) : color) | p.alpha_or;
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

/* Raw batches of at most 1024 triangles: count, scan and emission in ONE launch of up to 8 CTAs of 128 threads.
 * Every triangle runs the vertex stage once; the output offset of a CTA is the running total its predecessor
 * publishes (a chained scan: flags[b] = launch sequence number << 32 | triangles emitted by logical blocks 0..b).
 * CUDA promises no dispatch order between the CTAs of a grid, so the LOGICAL block index is a ticket drawn with an
 * atomic when the CTA starts running (flags[15], as in decoupled look-back scans): whoever holds ticket b-1 is
 * resident and never waits for b, so the chain cannot deadlock however the hardware schedules the grid.  The CTA
 * with the last ticket resets the counter for the next launch on this lane (every other ticket is drawn by then).
 * The total stays on the device: *d_total feeds k_front_small, so the host never waits for it. */
__global__ void __launch_bounds__(128)
k_raw_chain(const RawArgs a, pfcu_triangle *__restrict__ out, unsigned *__restrict__ d_total,
            unsigned long long *__restrict__ flags, unsigned seq)
{
    __shared__ unsigned s_warp[4];
    __shared__ unsigned s_prev, s_bid;
    pdl_trigger(); pdl_wait();
    if (threadIdx.x == 0) s_bid = (unsigned)atomicAdd(flags + 15, 1ull);
    __syncthreads();
    const unsigned bid = s_bid;
    const unsigned i = bid * 128u + threadIdx.x, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    pfv_vertex poly[PFV_MAX_POLY];
    int is3d = 0, face = 0, n = 0; unsigned state = 0;
    if (i < a.n) n = raw_process(a, i, poly, &is3d, &face, &state);
    unsigned x = (unsigned)n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, o); if ((int)lane >= o) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    unsigned woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) { const unsigned c = s_warp[w]; if (w < (int)warp) woff += c; total += c; }
    if (threadIdx.x == 0) {
        unsigned prev = 0;
        if (bid > 0) {
            const volatile unsigned long long *f = flags + (bid - 1);
            unsigned long long v;
            do { v = *f; } while ((unsigned)(v >> 32) != seq);
            prev = (unsigned)v;
        }
        __threadfence();
        *((volatile unsigned long long *)(flags + bid)) = ((unsigned long long)seq << 32) | (prev + total);
        s_prev = prev;
        if (bid == gridDim.x - 1) { *d_total = prev + total; flags[15] = 0ull; }
    }
    __syncthreads();
    const unsigned off = s_prev + woff + x - (unsigned)n;
    for (int k = 0; k < n; k++) pfv_emit(out + off + k, &poly[0], &poly[k + 1], &poly[k + 2], state, face, is3d);
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

/* total of a count -> exclusive-scan pair, written straight into page-locked host memory (mapped into the device's
   address space): the host waits for the stream instead of issuing two small copies */
__global__ void k_scan_total(const unsigned *__restrict__ last_offset, const unsigned *__restrict__ last_count, unsigned *__restrict__ host_total)
{
    host_total[0] = *last_offset; host_total[1] = *last_count;
}

__global__ void k_scan_add(unsigned *__restrict__ data, unsigned n, const unsigned *__restrict__ block_offsets)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] += block_offsets[i / 1024u];
}
