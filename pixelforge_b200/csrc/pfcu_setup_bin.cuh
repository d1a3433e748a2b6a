/* pfcu_setup_bin.cuh - kernels: triangle setup, order-preserving binning, fused small front end.
 * Part of the single translation unit pfcu.cu (included there, in order; not a stand-alone header). */

/* ------------------------------------------------------------------------------------------------ */
/* kernels: setup                                                                                   */
/* ------------------------------------------------------------------------------------------------ */

/* Programmatic dependent launch (sm_90+): the kernels of one batch are launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel's CTAs are scheduled as soon as every CTA of
 * its predecessor has started (pdl_trigger) and only wait, in pdl_wait, for the predecessor's completion and memory
 * flush - the launch latency between the short front-end kernels overlaps their execution.  Every kernel calls
 * pdl_wait before it touches anything an earlier kernel wrote (completion is transitive along the stream); both are
 * no-ops for an ordinary launch. */
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ int to_int_x86(float f) { return cvt_trunc_x86(f); }   /* (PFint)f == CVTTSS2SI */
__device__ __forceinline__ int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
__device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__device__ __forceinline__ int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__device__ __forceinline__ long long labs64(long long v) { return v < 0 ? -v : v; }

/* setup of triangle i; returns its bbox (empty = (1,1,0,0)) and whether it counts as rasterised */
__device__ __forceinline__ int4 setup_one(const pfcu_triangle *t /* global or shared */, const DevState *__restrict__ states, unsigned i,
                                          int surfW, int surfH, int4 *__restrict__ bbox, TriSetup *__restrict__ setup,
                                          TriData *__restrict__ data, bool *rasterised)
{
    bool valid = false;
    int4 out_box = make_int4(1, 1, 0, 0);
    {
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

/* ---- bulk asynchronous copies (the TMA unit's 1-D form: cp.async.bulk + mbarrier, SASS UBLKCP) ------------------ */
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}

/* k_setup: HBM-bound (152 B in, 224 B out per triangle).  A persistent CTA walks chunks of SETUP_THREADS triangles with
 * the input double-buffered: one elected thread has the TMA unit fetch the NEXT chunk (cp.async.bulk, completion on an
 * mbarrier) while the CTA computes the current one into shared staging, and the three output arrays leave as bulk stores
 * as well - whole 128-byte lines in both directions, a handful of instructions per chunk instead of a load loop and
 * strided 16-byte stores per thread (round 2: 94 -> see profiles/ us on the 1 M-triangle mesh). */
#define SETUP_SMEM_BYTES (2 * SETUP_THREADS * (int)sizeof(pfcu_triangle) + SETUP_THREADS * (16 + (int)sizeof(TriSetup) + (int)sizeof(TriData)) + 32)
__global__ void __launch_bounds__(SETUP_THREADS)
k_setup(const pfcu_triangle *__restrict__ tris, const DevState *__restrict__ states, unsigned n,
        int surfW, int surfH, int4 *__restrict__ bbox, TriSetup *__restrict__ setup, TriData *__restrict__ data,
        unsigned long long *__restrict__ counters)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    constexpr unsigned CHUNK_IN = SETUP_THREADS * (unsigned)sizeof(pfcu_triangle);
    static_assert(CHUNK_IN % 128 == 0 && sizeof(TriSetup) % 16 == 0 && sizeof(TriData) % 16 == 0, "bulk copies move multiples of 16 bytes");
    unsigned char *s_in = s_dyn;                                                            /* [2][CHUNK_IN] */
    int4 *s_bbox = reinterpret_cast<int4 *>(s_dyn + 2 * CHUNK_IN);
    TriSetup *s_setup = reinterpret_cast<TriSetup *>(s_bbox + SETUP_THREADS);
    TriData *s_data = reinterpret_cast<TriData *>(s_setup + SETUP_THREADS);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(s_data + SETUP_THREADS);  /* [2] */
    const unsigned tid = threadIdx.x;
    const unsigned nChunks = (n + SETUP_THREADS - 1) / SETUP_THREADS;
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    pdl_trigger(); pdl_wait();
    auto issue = [&](unsigned c, unsigned buf) {          /* thread 0: fetch chunk c into buffer buf */
        const unsigned here = min((unsigned)SETUP_THREADS, n - c * SETUP_THREADS);
        const unsigned bytes = (here * (unsigned)sizeof(pfcu_triangle)) & ~15u;             /* 152 bytes: a multiple of 8, not of 16 */
        mbar_expect_tx(&bar[buf], bytes);
        bulk_load(s_in + buf * CHUNK_IN, reinterpret_cast<const unsigned char *>(tris) + (size_t)c * CHUNK_IN, bytes, &bar[buf]);
    };
    unsigned c = blockIdx.x;
    if (tid == 0 && c < nChunks) issue(c, 0);
    unsigned counted = 0;
    for (unsigned it = 0; c < nChunks; c += gridDim.x, it++) {
        const unsigned buf = it & 1u, base = c * SETUP_THREADS;
        const unsigned here = min((unsigned)SETUP_THREADS, n - base);
        if (tid == 0) {
            if (c + gridDim.x < nChunks) issue(c + gridDim.x, buf ^ 1u);        /* the other buffer was consumed an iteration ago */
            if (it) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      /* the last chunk's stores have read the staging */
        }
        mbar_wait(&bar[buf], (it >> 1) & 1u);
        const pfcu_triangle *in = reinterpret_cast<const pfcu_triangle *>(s_in + buf * CHUNK_IN);
        if (tid == here - 1u && ((here * (unsigned)sizeof(pfcu_triangle)) & 15u))       /* the odd 8 bytes end the last record: its owner fetches them */
            reinterpret_cast<uint2 *>(s_in + buf * CHUNK_IN)[here * 19u - 1u] =
                __ldcs(reinterpret_cast<const uint2 *>(reinterpret_cast<const unsigned char *>(tris) + (size_t)c * CHUNK_IN) + (here * 19u - 1u));
        __syncthreads();                                                        /* staging is free (thread 0 waited for the stores above) */
        bool valid = false;
        if (tid < here) setup_one(in + tid, states, tid, surfW, surfH, s_bbox, s_setup, s_data, &valid);
        const unsigned b = __ballot_sync(0xffffffffu, valid);
        if ((tid & 31u) == 0u) counted += __popc(b);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");            /* generic-proxy writes -> the bulk stores' reads */
        __syncthreads();
        if (tid == 0) {
            bulk_store(bbox + base, s_bbox, here * 16u);
            bulk_store(setup + base, s_setup, here * (unsigned)sizeof(TriSetup));
            bulk_store(data + base, s_data, here * (unsigned)sizeof(TriData));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     /* shared memory must outlive the stores */
    if ((tid & 31u) == 0u && counted) atomicAdd(counters + 0, (unsigned long long)counted);
}

/* ------------------------------------------------------------------------------------------------ */
/* kernels: order-preserving coarse binning                                                         */
/* ------------------------------------------------------------------------------------------------ */

/* pass 1: counts[batch][bin] = number of triangles of this batch whose bbox touches the bin */
__global__ void __launch_bounds__(256)
k_bin_count(const int4 *__restrict__ bbox, unsigned n, unsigned batch, int binsX, int binsY, int bshift, int bshy, unsigned *__restrict__ counts)
{
    extern __shared__ unsigned s_cnt[];
    const int nb = binsX * binsY;
    pdl_trigger();
    for (int k = threadIdx.x; k < nb; k += blockDim.x) s_cnt[k] = 0;
    __syncthreads();
    pdl_wait();
    const unsigned base = blockIdx.x * batch;
    for (unsigned k = threadIdx.x; k < batch; k += blockDim.x) {
        const unsigned i = base + k;
        if (i >= n) break;
        const int4 b = __ldg(bbox + i);
        if (b.x >= b.z) continue;
        const int bx0 = max(b.x, 0) >> bshift, bx1 = min((b.z - 1) >> bshift, binsX - 1);
        const int by0 = max(b.y, 0) >> bshy, by1 = min(b.w >> bshy, binsY - 1);
        for (int by = by0; by <= by1; by++)
            for (int bx = bx0; bx <= bx1; bx++) atomicAdd(&s_cnt[by * binsX + bx], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nb; k += blockDim.x) counts[(size_t)blockIdx.x * nb + k] = s_cnt[k];
}

/* exclusive scan over the bin totals -> bin start offsets, starts[nb] = total; one CTA of 1024 threads */
__device__ __forceinline__ void bin_starts_scan(const unsigned *totals, int nb, unsigned *__restrict__ starts)
{
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int k = base + threadIdx.x;
        const unsigned v = (k < nb) ? __ldcg(totals + k) : 0u;      /* written by other CTAs of this launch: read from L2 */
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

/* pass 2: per bin, exclusive scan of counts over batches (in place) + bin totals.  A CTA of 32 warps owns 32
 * consecutive bins (one 128-byte row segment per batch); warp w owns a contiguous range of batches: it sums its
 * range, the 32 partial sums are scanned across warps, and it walks its range again writing the prefixes.
 * pass 3 rides along: the CTA that finishes last (a ticket drawn with an atomic after its totals are out) turns the bin
 * totals into the bin start offsets - one launch less on the critical path of every batch. */
__global__ void __launch_bounds__(1024)
k_bin_scan(unsigned *__restrict__ counts, int nBatches, int nb, unsigned *__restrict__ totals, unsigned *__restrict__ starts, unsigned *__restrict__ ticket)
{
    __shared__ unsigned s_part[32][33];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bin = blockIdx.x * 32 + lane;
    const bool live = bin < nb;
    const int per = (nBatches + 31) / 32;
    const int k0 = warp * per, k1 = min(k0 + per, nBatches);
    pdl_trigger(); pdl_wait();
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
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) *ticket = 0u;         /* for the next launch on this lane */
    __threadfence();
    bin_starts_scan(totals, nb, starts);
}

/* A bin-list entry carries the triangle's visited rectangle [x0, x1] x [y0, y1] (inclusive) clipped to the bin and
 * relative to the bin's origin, 8 bits per coordinate (bins are at most 256 pixels wide): the rasteriser's
 * queue filter then needs no dependent load.  Bins are 2^bshift pixels wide and 2^bshy pixels high: batches of many
 * small triangles get bins as flat as the rasteriser's 64x8 slices, so that a slice reads (nearly) only its own list. */
__device__ __forceinline__ unsigned bin_rel_bbox(const int4 b, int bx, int by, int bshift, int bshy)
{
    const int ox = bx << bshift, oy = by << bshy, hi = (1 << bshift) - 1, hiy = (1 << bshy) - 1;
    const int x0 = min(max(b.x - ox, 0), hi), x1 = min(max(b.z - 1 - ox, 0), hi);
    const int y0 = min(max(b.y - oy, 0), hiy), y1 = min(max(b.w - oy, 0), hiy);
    return (unsigned)x0 | ((unsigned)y0 << 8) | ((unsigned)x1 << 16) | ((unsigned)y1 << 24);
}

/* pass 4: ordered fill.  The CTA walks its triangles 256 at a time.  Every bin COLUMN belongs to one warp
 * (bx & 7); each warp visits, in triangle order, the triangles whose bin rectangle has a column of its own and
 * appends them to those bins.  A bin is therefore written by one warp only, in submission order, with no
 * CTA barrier inside a group and no dependence on how the 256 triangles are spread over the screen. */
__global__ void __launch_bounds__(256)
k_bin_fill(const int4 *__restrict__ bbox, unsigned n, unsigned batch, int binsX, int binsY, int bshift, int bshy,
           const unsigned *__restrict__ offsets /* scanned counts */, const unsigned *__restrict__ starts,
           uint2 *__restrict__ list, unsigned list_cap)
{
    pdl_trigger(); pdl_wait();
    if (starts[binsX * binsY] > list_cap) return;       /* the lists do not fit: the rasterisers filter the batch themselves */
    extern __shared__ unsigned s_mem[];
    unsigned *s_pos = s_mem;                    /* [nb] running write position of this batch per bin */
    __shared__ int4 s_rect[256];
    __shared__ int4 s_bbox[256];
    const int nb = binsX * binsY;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) s_pos[k] = starts[k] + offsets[(size_t)blockIdx.x * nb + k];
    const unsigned base = blockIdx.x * batch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned k0 = 0; k0 < batch && base + k0 < n; k0 += 256) {
        const unsigned i = base + k0 + threadIdx.x;
        int4 r = make_int4(1, 1, 0, 0);
        int4 b = make_int4(1, 1, 0, 0);
        if (i < n) {
            b = __ldg(bbox + i);
            if (b.x < b.z) {
                r.x = max(b.x, 0) >> bshift; r.z = min((b.z - 1) >> bshift, binsX - 1);
                r.y = max(b.y, 0) >> bshy; r.w = min(b.w >> bshy, binsY - 1);
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
                            list[pos] = make_uint2(my_idx, bin_rel_bbox(qb, first, by, bshift, bshy));
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
                        list[pos] = make_uint2(idx, bin_rel_bbox(tb, f0 + (cx << 3), t.y + cy, bshift, bshy));
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
__device__ __forceinline__ void
front_small_body(const pfcu_triangle *__restrict__ tris, const DevState *__restrict__ states, unsigned n_host, const unsigned *__restrict__ d_n,
                 int surfW, int surfH,
                 int4 *__restrict__ bbox, TriSetup *__restrict__ setup, TriData *__restrict__ data, unsigned long long *__restrict__ counters,
                 int binsX, int binsY, int bshift, int bshy, unsigned *__restrict__ starts, uint2 *__restrict__ list)
{
    extern __shared__ unsigned s_mem[];
    unsigned *s_pos = s_mem;                    /* [nb] counts, then running write positions */
    __shared__ int4 s_rect[FRONT_SMALL_MAX];
    __shared__ int4 s_bbox[FRONT_SMALL_MAX];
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    pdl_trigger(); pdl_wait();
    const unsigned n = d_n ? min(*d_n, (unsigned)(FRONT_SMALL_MAX * FRONT_SMALL_CHUNKS)) : n_host;
    if (d_n && threadIdx.x == 0) atomicAdd(counters + 3, (unsigned long long)n);        /* "submitted", counted where the count is known */
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
        if (i < n) b = setup_one(tris + i, states, i, surfW, surfH, bbox, setup, data, &rasterised);
        if (b.x < b.z) {
            const int rx0 = max(b.x, 0) >> bshift, rx1 = min((b.z - 1) >> bshift, binsX - 1);
            const int ry0 = max(b.y, 0) >> bshy, ry1 = min(b.w >> bshy, binsY - 1);
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
            r.y = max(b.y, 0) >> bshy; r.w = min(b.w >> bshy, binsY - 1);
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
                            list[pos] = make_uint2(my_idx, bin_rel_bbox(qb, first, by, bshift, bshy));
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
                        list[pos] = make_uint2(idx, bin_rel_bbox(tb, f0 + (cx << 5), t.y + cy, bshift, bshy));
                        s_pos[bin] = pos + 1;
                    }
                }
                __syncwarp();
            }
        }
    }
}

__global__ void __launch_bounds__(1024)
k_front_small(const pfcu_triangle *__restrict__ tris, const DevState *__restrict__ states, unsigned n_host, const unsigned *__restrict__ d_n,
              int surfW, int surfH,
              int4 *__restrict__ bbox, TriSetup *__restrict__ setup, TriData *__restrict__ data, unsigned long long *__restrict__ counters,
              int binsX, int binsY, int bshift, int bshy, unsigned *__restrict__ starts, uint2 *__restrict__ list)
{
    front_small_body(tris, states, n_host, d_n, surfW, surfH, bbox, setup, data, counters, binsX, binsY, bshift, bshy, starts, list);
}
