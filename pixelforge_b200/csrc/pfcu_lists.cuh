/* pfcu_lists.cuh - kernels: device-resident render lists replayed for many surfaces per launch (pfcu_submit_list_jobs).
 * Part of the single translation unit pfcu.cu (included there, in order; not a stand-alone header). */

/* ------------------------------------------------------------------------------------------------ */
/* kernels: render-list jobs (grid y = job = one surface's pending replays)                         */
/* ------------------------------------------------------------------------------------------------ */

struct DevListSeg { const pfcu_rawtri *tris; unsigned first_tri, first_call; };

/* Everything one job's four kernels read, in device memory (uploaded with the job's tables in one copy). */
struct DevJob {
    DevListSeg seg[PFCU_LIST_JOB_MAX_SEGMENTS]; unsigned n_seg, n_raw;
    const pfcu_list_call *calls; const pfcu_vparams_lit *vp; const float *pow_tables; const DevState *states;
    uint32_t *color; float *depth; unsigned W, H;
    unsigned clear, clear_rgba; float clear_z;
    pfcu_triangle *d_tris; unsigned *d_total; unsigned long long *chain;         /* vertex-stage output, its count, chained-scan flags */
    int binsX, binsY, bshift;
    RasterParams rp;                                                            /* setup / binning outputs and the rasteriser's view of the job */
};

/* pfClear of every job's surface with the reference's SIMD behaviour (context.c:696-713, Q12): pixels [8, size - size%8)
 * receive the value, pixels 0..7 are left alone, the tail copies pixel 0. */
__global__ void __launch_bounds__(256)
k_jobs_clear(const DevJob *__restrict__ jobs)
{
    const DevJob &j = jobs[blockIdx.y];
    if (!j.clear) return;
    const unsigned size = j.W * j.H, aligned = size - (size % 8u);
    const unsigned stride = gridDim.x * 256u;
    const uint32_t c0 = j.color[0]; const float z0 = j.depth[0];      /* never written by this kernel */
    if ((aligned & 3u) == 0 && aligned > 8u) {
        const uint4 cv = make_uint4(j.clear_rgba, j.clear_rgba, j.clear_rgba, j.clear_rgba);
        const float4 zv = make_float4(j.clear_z, j.clear_z, j.clear_z, j.clear_z);
        for (unsigned k = 2u + blockIdx.x * 256u + threadIdx.x; k < aligned / 4u; k += stride) {
            reinterpret_cast<uint4 *>(j.color)[k] = cv; reinterpret_cast<float4 *>(j.depth)[k] = zv;
        }
    } else for (unsigned k = 8u + blockIdx.x * 256u + threadIdx.x; k < aligned; k += stride) { j.color[k] = j.clear_rgba; j.depth[k] = j.clear_z; }
    if (blockIdx.x == 0) for (unsigned k = aligned + threadIdx.x; k < size; k += 256u) { j.color[k] = c0; j.depth[k] = z0; }
}

/* Vertex stage of a job: k_raw_chain over the concatenation of the job's list segments.  Triangle i lives in the segment
 * whose [first_tri, next first_tri) holds i; its recorded call index selects the pfcu_list_call of this replay. */
__global__ void __launch_bounds__(128)
k_list_chain(const DevJob *__restrict__ jobs, unsigned seq)
{
    __shared__ unsigned s_warp[4];
    __shared__ unsigned s_prev, s_bid;
    const DevJob &job = jobs[blockIdx.y];
    unsigned long long *flags = job.chain;
    pdl_trigger(); pdl_wait();
    const unsigned nblocks = (job.n_raw + 127u) / 128u;
    if (blockIdx.x >= nblocks) {                              /* uniform per CTA; tickets are drawn by the first nblocks CTAs only */
        if (nblocks == 0 && blockIdx.x == 0 && threadIdx.x == 0) *job.d_total = 0u;     /* a job that only clears */
        return;
    }
    if (threadIdx.x == 0) s_bid = (unsigned)atomicAdd(flags + 15, 1ull);
    __syncthreads();
    const unsigned bid = s_bid;
    const unsigned i = bid * 128u + threadIdx.x, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    pfv_vertex poly[PFV_MAX_POLY];
    int is3d = 0, face = 0, n = 0; unsigned state = 0;
    if (i < job.n_raw) {
        unsigned sgi = 0;
        for (unsigned k = 1; k < job.n_seg; k++) if (job.seg[k].first_tri <= i) sgi = k;
        const DevListSeg sg = job.seg[sgi];
        const pfcu_rawtri *t = sg.tris + (i - sg.first_tri);
        const pfcu_list_call cl = job.calls[sg.first_call + t->state];
        state = cl.state;
        n = raw_process_tri(t, job.vp + cl.vparams, job.pow_tables, cl.override_color != 0u, cl.rgba, poly, &is3d, &face);
    }
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
        if (bid == nblocks - 1) { *job.d_total = prev + total; flags[15] = 0ull; }
    }
    __syncthreads();
    const unsigned off = s_prev + woff + x - (unsigned)n;
    for (int k = 0; k < n; k++) pfv_emit(job.d_tris + off + k, &poly[0], &poly[k + 1], &poly[k + 2], state, face, is3d);
}

__global__ void __launch_bounds__(1024)
k_front_small_jobs(const DevJob *__restrict__ jobs, unsigned long long *__restrict__ counters)
{
    const DevJob &j = jobs[blockIdx.y];
    front_small_body(j.d_tris, j.states, 0u, j.d_total, (int)j.W, (int)j.H, const_cast<int4 *>(j.rp.bbox), const_cast<TriSetup *>(j.rp.setup),
                     const_cast<TriData *>(j.rp.data), counters, j.binsX, j.binsY, j.bshift, j.bshift,
                     const_cast<unsigned *>(j.rp.bin_starts), const_cast<uint2 *>(j.rp.bin_list));
}

template <bool HAS_PHONG, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
k_raster_frag_jobs(const DevJob *__restrict__ jobs)
{
    __shared__ RasterParams s_p;                               /* the kernel-parameter copy the single-surface kernel gets for free */
    {
        const unsigned *src = reinterpret_cast<const unsigned *>(&jobs[blockIdx.y].rp);
        unsigned *dst = reinterpret_cast<unsigned *>(&s_p);
        pdl_wait();
        for (unsigned k = threadIdx.x; k < sizeof(RasterParams) / 4u; k += NW * 32) dst[k] = src[k];
    }
    __syncthreads();
    raster_frag_body<HAS_PHONG, NW>(s_p);
}
