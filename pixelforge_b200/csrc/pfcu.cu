/*
 * pfcu.cu - sm_100a implementation of the pfcu C-ABI (include/pfcu.h): the per-fragment triangle
 * path of PixelForge on a B200.  One translation unit; the kernels live in the .cuh files included below:
 *   pfcu_device_math.cuh    x86 lane semantics (MINPS/MAXPS, CVTPS2DQ, RCPPS/RSQRTPS tables, cephes log/exp),
 *                           colour math, texture sampling, per-fragment Blinn-Phong
 *   pfcu_setup_bin.cuh      k_setup, k_bin_count/scan/starts/fill, k_front_small
 *   pfcu_raster_tiles.cuh   k_raster      (batches of large triangles)
 *   pfcu_raster_frag.cuh    k_raster_frag (batches of many small triangles)
 *   pfcu_vertex_prims.cuh   k_vertex_*, k_raw_*, k_prims, k_scan_*
 *   pfcu_surface.cuh        k_fill*, k_clear_tail, k_pack_tiles
 * and this file holds the runtime (stream lanes, buffers, pinned staging) and the C-ABI entry points.
 *
 * Pipeline per submitted batch (each surface's work stays on one stream "lane", order preserving):
 *   vertex stage (optional)  k_vertex_* for vertex-array draws, k_raw_* / k_raw_chain for immediate mode and
 *             render lists: pf_vstage.h compiled as device code (normal transform, material multiply, Gouraud
 *             lighting, clipping, projection); count -> scan -> emit keeps order;
 *   k_setup   one thread per triangle: integer snap, signed area / face cull, bbox, int32 edge
 *             constants, 1/sum (reference: triangles.c:294-349); writes bbox[], TriSetup[], TriData[];
 *   k_bin_*   order-preserving binning (256 px bins for large triangles, 64 px bins for many small ones):
 *             count -> scan -> ordered fill of {triangle, bin-relative rectangle} entries
 *             (k_front_small fuses setup and binning for batches of <= 1024 triangles);
 *   k_raster / k_raster_frag   colour + depth staged in shared memory, the bin's list filtered into a shared
 *             queue in order, fixed pixel ownership per warp (=> blending/depth order equals submission
 *             order without atomics).  Coverage, depth, colour interpolation, texturing, per-fragment
 *             Blinn-Phong and blending restate the reference's AVX2 lane arithmetic bit for bit
 *             (triangles.c:400-529, color.h, sampler.h, blend.h, depth.h, lighting.c:148-258, simd.h cephes
 *             log/exp); RCPPS/RSQRTPS come from host-harvested tables.  Load/store is 128-bit and coalesced.
 *   k_prims   points and lines (pf_prims.h), in submission order per tile.
 * No tensor cores: nothing on this path is a dense contraction.  Compile with -fmad=false.
 */
#include "pfcu.h"
#include "pf_vstage.h"
#include "pf_prims.h"
#include "pf_pixfmt.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <math.h>
#include <time.h>

#include <vector>
#include <mutex>
#include <thread>
#include <condition_variable>
#include <functional>
#include <atomic>

/* ------------------------------------------------------------------------------------------------ */
/* configuration                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

#define MAX_DEVS    8               /* devices of the single-process multi-GPU mode (PF_CUDA_DEVICES), see "multi-device mode" below */
#define TILE        64              /* screen tile edge in pixels                                    */
#define TILE_PIX    (TILE * TILE)
/* A bin is 2^k x 2^k pixels, chosen per batch: 256 (4x4 tiles) for batches of large triangles, where a
 * triangle would otherwise be listed in thousands of bins; 64 (= one tile) for batches of many small ones,
 * where every slice CTA of the rasteriser would otherwise scan a 16 times longer list. */
#define BIN_SHIFT_COARSE 8
#define BIN_SHIFT_FINE   6
#define RASTER_THREADS 256
#define QUEUE_CAP   1024            /* triangle indices buffered per tile between raster passes      */
#define SETUP_THREADS 128
#define BIN_BATCH   1024            /* triangles per binning CTA (256 for mid-sized batches, see launch_pipeline) */
#define MAX_BINS    10240           /* bin counters live in dynamic shared memory: 40 KB + the 8 KB of static rectangles of k_bin_fill = the 48 KB
                                       a kernel gets without opting in; 7680x4320 in 64 px bins = 8160.  Larger surfaces take coarser bins. */

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

/* *_f: the colour channels as the per-fragment Blinn-Phong uses them, float(c) * (1/255) (lighting.c:155-175), computed once
   per state on the host (same IEEE single multiply) instead of once per fragment */
struct DevLight { float pos[3], dir[3], inner, outer, attc, attl, attq; unsigned ambient, diffuse, specular; float amb_f[3], dif_f[3], spc_f[3]; };
struct DevMaterial { unsigned ambient, diffuse, specular, emission; float shininess; float amb_f[3], spc_f[3]; };

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
    unsigned tex_leader;                        /* texels come through the reference's BGRA8 getter (pfcu_raster_rows.cuh) */
};

static_assert(offsetof(DevState, tex_fw) % 16 == 0, "tex_fw..tex_ty are fetched as one float4");

#define MAX_BANDS 6
#define PCIE_GBS 50.0        /* device -> host copies of page-locked memory as measured on this pool (8.3 MB in 0.16 ms) */
#define PUSH_HOST_CTAS 48u        /* CTAs of a tile store towards host memory (see k_push_tiles) */
struct pfcu_surface {
    uint32_t w, h; uint32_t *color; float *depth; bool owned;
    int fmt;                                        /* PFCU_TEX_*: the caller's layout; the device holds canonical RGBA8 */
    unsigned char *conv; size_t conv_bytes;         /* device staging for layout conversion at upload / download */
    uint32_t rank, world; uint32_t tiles_x, tiles_y;
    int lane; cudaEvent_t done; bool has_done;      /* last work enqueued on this surface */
    uint32_t *peer_color; float *peer_depth;        /* present target (peer memory or another local surface), or nullptr */
    void *ipc_color, *ipc_depth;                    /* mappings opened with cudaIpcOpenMemHandle (to close) */
    bool aliased;                                   /* a texture aliases the colour buffer (render to texture) */
    /* the last operation on the surface was a rasterisation in bands of tile rows: band b covers rows [band_y[b], band_y[b+1])
       and band_evt[b] fires when it is complete (see launch_raster_bands / pfcu_surface_download_async) */
    bool bands_valid; int n_bands; uint32_t band_y[MAX_BANDS + 1]; cudaEvent_t band_evt[MAX_BANDS];
    uint32_t band_owned[MAX_BANDS + 1];             /* tile split: band b = this rank's owned tiles [band_owned[b], band_owned[b+1]) */
    cudaEvent_t t_front, t_raster; bool t_pending; bool raster_longer;   /* banded frames time their rasterisation: see launch_pipeline */
    cudaEvent_t push_evt[MAX_BANDS]; bool pushed_in_bands; bool peer_is_host;      /* peer_is_host: peer_color / peer_depth address the caller's page-locked buffer */      /* pfcu_surface_push_tiles of a banded surface: band b's tiles have been stored */
    /* bands pay off only when a read-back follows the batch: batches rasterised since the last read-back, and how many
       there were between the two read-backs before - the batch predicted to be a frame's last one goes out in bands */
    unsigned n_since_read, n_per_read;
    /* multi-device mode: rep[d] is this surface on device d (rep[0] == the handle the caller holds, nullptr everywhere
       on replicas); split: every device rasterises only its own tiles and a read-back gathers them on device 0;
       full_evt (device 0): the last operation that wrote tiles of other devices too; pushed (replicas): the last store of
       this device's tiles into device 0's surface */
    pfcu_surface *rep[MAX_DEVS]; bool split; cudaEvent_t full_evt, pushed; bool has_full, peer_bands;
};
struct pfcu_texture { uint32_t w, h; int fmt; unsigned char *pixels; bool owned; pfcu_surface *alias; bool leader; pfcu_texture *rep[MAX_DEVS]; };
struct pfcu_list { pfcu_rawtri *tris; uint32_t n; };
struct pfcu_batch {
    DevState *states; uint32_t n_states; pfcu_triangle *tris; uint32_t n_tris; unsigned feature_mask; int single_prog; bool leader_tex, pix_tex;
    std::vector<pfcu_surface *> deps;
    pfcu_batch *rep[MAX_DEVS];
};

/* ------------------------------------------------------------------------------------------------ */
/* runtime state                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

struct PinnedBlock {
    /* done / pending / d_copy / dev_gen are per device (index = Runtime::index): every device copies from the block
       with its own stream and keeps its own mirror */
    void *p; size_t bytes; cudaEvent_t done[MAX_DEVS]; bool pending[MAX_DEVS];
    /* static geometry (pfcu_host_set_static): the block is mirrored in device memory and draws read the mirror; gen counts
       the application's modifications, dev_gen is what the mirror holds */
    bool is_static; unsigned gen, dev_gen[MAX_DEVS]; unsigned char *d_copy[MAX_DEVS];
    const void *imax_ptr; uint32_t imax_count, imax_value; unsigned imax_gen;     /* largest index of the last index range scanned */
};

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
    bool touched = false;               /* work enqueued on this lane since the last pfcu_fence()            */
    bool need_fence_wait = false;       /* the next use of this lane must first wait for lane 0's fence event */
    /* raw-triangle batches: their vertex stage is counted on a side stream so that the one host wait (for the
       output triangle count) does not wait for the rasterisation queued on the lane */
    cudaStream_t vstream = nullptr; cudaEvent_t raw_done = nullptr, vready = nullptr;
    unsigned char *d_raw = nullptr; size_t cap_raw = 0;
    unsigned *h_total = nullptr;                                        /* pinned: {last offset, last count} */
    unsigned *d_total = nullptr;                                        /* output count of a sync-free raw batch */
    unsigned long long *d_chain = nullptr; unsigned chain_seq = 0;      /* chained-scan flags of k_raw_chain */
    unsigned char *d_idx = nullptr; size_t cap_idx = 0;                 /* index buffer of a draw whose vertex count the device finds */
    /* bin-list sizing without a host wait: the total of the last large batch comes back asynchronously and only steers
       the capacity of later batches; a batch that does not fit its list is still rendered correctly (see launch_pipeline) */
    unsigned *h_list_total = nullptr; cudaEvent_t list_evt = nullptr; bool list_pending = false;
    size_t list_hint = 0; uint32_t list_hint_n = 0;
};

#define MAX_LANES 8

struct Runtime {
    int index = 0;                              /* 0: the primary runtime; 1..: the workers of the multi-device mode */
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
    /* render-list jobs (pfcu_submit_list_jobs): per-slot scratch, the packed upload and its pinned staging */
    struct JobSlot {
        pfcu_triangle *d_tris = nullptr; int4 *bbox = nullptr; TriSetup *setup = nullptr; TriData *data = nullptr; size_t cap_tris = 0;
        uint2 *bin_list = nullptr; size_t cap_list = 0; unsigned *bin_start = nullptr; unsigned *d_total = nullptr; unsigned long long *chain = nullptr;
    };
    std::vector<JobSlot> job_slots;
    unsigned char *d_jobs = nullptr, *h_jobs = nullptr; size_t cap_jobs = 0; cudaEvent_t jobs_copied = nullptr, jobs_done = nullptr; unsigned jobs_seq = 0;
    cudaStream_t band_streams[MAX_BANDS] = {}; cudaStream_t copy_stream = nullptr;
    cudaEvent_t front_evt = nullptr;
    bool frag_attr_set = false, setup_attr_set = false;
    std::recursive_mutex mu;                    /* the C-ABI is serialised: contexts on several threads share one runtime */
    char err[512] = { 0 };
};

/* One runtime per device.  Application threads use the primary one; the worker threads of the multi-device mode
 * (one per additional GPU) point t_rt at theirs, so that every entry point below runs unchanged on either. */
static Runtime g_main;
static thread_local Runtime *t_rt = nullptr;
#define RT (*(t_rt ? t_rt : &g_main))
#define LN (*RT.cur)

/* ---- multi-device mode (one process, several GPUs: PF_CUDA_DEVICES=0,1,...) ------------------------------------------
 * north_star / SURVEY 8-e "large framebuffers are screen-tile split": with more than one device listed, every surface,
 * texture and resident batch exists once per device (the handle the caller holds is device 0's and links to the
 * others), every submission is replayed on all devices - one worker thread per additional GPU, each with its own
 * Runtime - and a device rasterises only the 64x64 tiles it owns (tile % n == device, as in the multi-process split).
 * Full-surface operations run everywhere; a read-back first has every other device store its tiles into device 0's
 * surface over NVLink (peer access, k_push_tiles).  Surfaces that are sampled as textures (framebuffer objects), small
 * surfaces and non-RGBA8 targets are rendered in full by every device instead, which keeps them usable everywhere.
 * Results are byte-identical to one GPU.  Application-visible behaviour does not change; no entry point is added. */
struct Multi {
    int n = 1; int dev[MAX_DEVS] = { 0 };
    Runtime *rt[MAX_DEVS] = { &g_main };
    std::thread th[MAX_DEVS];
    std::mutex m; std::condition_variable cv_job, cv_done;
    const std::function<int(int)> *job = nullptr; std::atomic<unsigned> seq{0}; std::atomic<int> pending{0}; int rc[MAX_DEVS] = { 0 };
    std::atomic<bool> quit{false};
    size_t split_min_pixels = (size_t)1 << 20;
};
/* never destroyed: worker threads may still wait on its condition variable when the process exits without pfcu_shutdown */
static Multi *const mg_ptr = new Multi;
#define mg (*mg_ptr)
static thread_local bool t_dispatching = false;     /* this thread is inside multi_run_all: entry points it calls act on one device */
#define MULTI_HERE() (mg.n > 1 && !t_rt && !t_dispatching)

/* Calls come in bursts (the submissions of a frame): after a job a worker spins for a short while before it goes to
   sleep on the condition variable, and the dispatching thread spins for the workers likewise - a wake-up through the
   kernel costs tens of microseconds per call and device, which a 1 ms frame on 8 GPUs would feel. */
#define MULTI_SPIN_US 200.0
static double multi_now_us(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; }

static void multi_worker(int d)
{
    t_rt = mg.rt[d];
    unsigned seen = 0;
    for (;;) {
        const double t0 = multi_now_us();
        while (mg.seq.load(std::memory_order_acquire) == seen && !mg.quit.load(std::memory_order_relaxed) && multi_now_us() - t0 < MULTI_SPIN_US) { }
        if (mg.seq.load(std::memory_order_acquire) == seen && !mg.quit.load()) {
            std::unique_lock<std::mutex> lk(mg.m);
            mg.cv_job.wait(lk, [&] { return mg.quit.load() || mg.seq.load() != seen; });
        }
        if (mg.quit.load()) return;
        seen = mg.seq.load(std::memory_order_acquire);
        const std::function<int(int)> *job = mg.job;
        mg.rc[d] = (*job)(d);
        if (mg.pending.fetch_sub(1, std::memory_order_acq_rel) == 1) { std::lock_guard<std::mutex> lk(mg.m); mg.cv_done.notify_one(); }
    }
}

static int multi_run_all(const std::function<int(int)> &fn)
{
    if (t_rt) return fn(RT.index);                       /* a worker acts on its own device */
    if (mg.n <= 1 || t_dispatching) return fn(0);
    {
        std::lock_guard<std::mutex> lk(mg.m);
        mg.job = &fn; mg.pending.store(mg.n - 1, std::memory_order_relaxed); mg.seq.fetch_add(1, std::memory_order_release);
    }
    mg.cv_job.notify_all();
    t_dispatching = true;
    const int rc0 = fn(0);
    t_dispatching = false;
    {
        const double t0 = multi_now_us();
        while (mg.pending.load(std::memory_order_acquire) != 0 && multi_now_us() - t0 < MULTI_SPIN_US) { }
        if (mg.pending.load(std::memory_order_acquire) != 0) {
            std::unique_lock<std::mutex> lk(mg.m);
            mg.cv_done.wait(lk, [&] { return mg.pending.load() == 0; });
        }
    }
    if (rc0) return rc0;
    for (int d = 1; d < mg.n; d++) if (mg.rc[d]) {
        /* pfcu_last_error() reads the primary runtime: bring the worker's message over */
        char msg[sizeof g_main.err];
        snprintf(msg, sizeof msg, "device %d: %s", mg.dev[d], mg.rt[d]->err);
        memcpy(g_main.err, msg, sizeof msg);
        return mg.rc[d];
    }
    return PFCU_OK;
}

/* Device-side counters.  The workers of the multi-device mode repeat the setup of every triangle, and the rendering of
   surfaces that are not split: they count those into a sink, so that the sums over the devices stay what one GPU counts. */
static unsigned long long *setup_counters(void) { return RT.index == 0 ? RT.d_counters : RT.d_counters + 4; }
static unsigned long long *raster_counters(const pfcu_surface *s) { return (RT.index == 0 || s->world > 1) ? RT.d_counters : RT.d_counters + 4; }

/* the states of a submission as device d sees them: texture handles replaced by that device's replicas */
static std::vector<pfcu_state> states_for_device(const pfcu_state *states, uint32_t n, int d)
{
    std::vector<pfcu_state> v(states, states + n);
    if (d) for (auto &st : v) if (st.texture && st.texture->rep[d]) st.texture = st.texture->rep[d];
    return v;
}

/* shut the workers' runtimes down and end their threads */
static void multi_stop(void)
{
    if (mg.n <= 1) return;
    multi_run_all([](int d) -> int { if (d) pfcu_shutdown(); return (int)PFCU_OK; });
    { std::lock_guard<std::mutex> lk(mg.m); mg.quit.store(true); }
    mg.cv_job.notify_all();
    for (int d = 1; d < mg.n; d++) { if (mg.th[d].joinable()) mg.th[d].join(); delete mg.rt[d]; mg.rt[d] = nullptr; }
    mg.n = 1; mg.quit.store(false);
}
/* Every locked entry point also makes the runtime's device current on the calling thread: the C-ABI is shared by
 * contexts on several host threads, and a thread that has not called pfcu_init starts with device 0 current
 * (wrong allocations and "invalid resource handle" launches when PF_CUDA_DEVICE / LOCAL_RANK picked another one). */
struct ApiGuard {
    std::lock_guard<std::recursive_mutex> lk;
    ApiGuard() : lk(RT.mu)
    {
        static thread_local int t_device = -1;
        if (RT.ok && t_device != RT.device) { if (cudaSetDevice(RT.device) == cudaSuccess) t_device = RT.device; }
    }
};
#define API_LOCK ApiGuard api_lock_

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf(RT.err, sizeof RT.err, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    return PFCU_ERR_CUDA; } } while (0)

#define CKP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf(RT.err, sizeof RT.err, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    return nullptr; } } while (0)

__constant__ const uint32_t *c_rcp_tab;
__constant__ const uint32_t *c_rsq_tab;
__constant__ int c_rcp_shift, c_rsq_shift, c_rsq_bits;

#include "pfcu_device_math.cuh"

#include "pfcu_setup_bin.cuh"

#include "pfcu_raster_tiles.cuh"

#include "pfcu_raster_frag.cuh"

#include "pfcu_raster_rows.cuh"

#include "pfcu_vertex_prims.cuh"

#include "pfcu_surface.cuh"

#include "pfcu_lists.cuh"

/* ------------------------------------------------------------------------------------------------ */
/* host: runtime                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

/* Launch as a programmatic dependent of the previous kernel in the stream (see pdl_wait in pfcu_setup_bin.cuh). */
static const bool g_use_pdl = !(getenv("PF_CUDA_PDL") && atoi(getenv("PF_CUDA_PDL")) == 0);
template <typename... KArgs, typename... Args>
static cudaError_t launch_dep(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

/* the single-program instantiations of the big-triangle rasteriser, full tiles and half-height slices */
template <int PROG> static cudaError_t launch_fixed(int slices, unsigned grid, cudaStream_t st, const RasterParams &p)
{
    return slices == 2 ? launch_dep(k_raster<false, 8, PROG, 32>, dim3(grid * 2), dim3(256), (size_t)0, st, p)
                       : launch_dep(k_raster<false, 8, PROG, 64>, dim3(grid), dim3(256), (size_t)0, st, p);
}

template <typename T> static int grow(T **p, size_t *cap, size_t need)
{
    if (need <= *cap) return PFCU_OK;
    size_t ncap = *cap ? *cap : 1024;
    while (ncap < need) ncap *= 2;
    CK(cudaStreamSynchronize(LN.stream));
    cudaFree(*p); *p = nullptr; *cap = 0;
    if (cudaMalloc(p, ncap * sizeof(T)) != cudaSuccess) { snprintf(RT.err, sizeof RT.err, "out of device memory growing scratch to %zu elements", ncap); return PFCU_ERR_OOM; }
    *cap = ncap;
    return PFCU_OK;
}

extern "C" {

const char *pfcu_last_error(void) { return RT.err; }
const char *pfcu_backend_name(void) { return "cuda-sm_100a"; }

int pfcu_init(int device)
{
    API_LOCK;
    if (RT.ok) return PFCU_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        snprintf(RT.err, sizeof RT.err, "no CUDA device: %s", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return PFCU_ERR_NO_DEVICE;
    }
    /* PF_CUDA_DEVICES=a,b,c...: the multi-device mode; the first device listed is the primary one */
    int listed[MAX_DEVS], n_listed = 0;
    if (!t_rt && getenv("PF_CUDA_DEVICES")) {
        const char *e = getenv("PF_CUDA_DEVICES");
        while (*e && n_listed < MAX_DEVS) {
            char *end; const long v = strtol(e, &end, 10);
            if (end == e) break;
            bool dup = false; for (int k = 0; k < n_listed; k++) dup |= listed[k] == (int)v;
            if (v >= 0 && v < count && !dup) listed[n_listed++] = (int)v;
            e = (*end == ',') ? end + 1 : end;
            if (*end && *end != ',') break;
        }
        if (n_listed && device < 0) device = listed[0];
    }
    if (device < 0) {
        const char *env = getenv("PF_CUDA_DEVICE");
        if (!env) env = getenv("LOCAL_RANK");
        device = env ? atoi(env) : 0;
        if (device < 0 || device >= count) device = 0;
    }
    if (device >= count) { snprintf(RT.err, sizeof RT.err, "device %d out of range (%d devices)", device, count); return PFCU_ERR_NO_DEVICE; }
    CK(cudaSetDevice(device));
    RT.device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    RT.sms = prop.multiProcessorCount;
    if (prop.major < 10) {
        snprintf(RT.err, sizeof RT.err, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return PFCU_ERR_NO_DEVICE;
    }
    {
        const char *env = getenv("PF_CUDA_LANES");
        RT.n_lanes = env ? atoi(env) : 4;
        if (RT.n_lanes < 1) RT.n_lanes = 1;
        if (RT.n_lanes > MAX_LANES) RT.n_lanes = MAX_LANES;
    }
    for (int i = 0; i < RT.n_lanes; i++) {
        RT.cur = &RT.lanes[i];
        CK(cudaStreamCreateWithFlags(&LN.stream, cudaStreamNonBlocking));
        LN.own_stream = true;
        CK(cudaMalloc(&LN.d_bin_start, ((MAX_BINS + 2) * 2 + 4) * sizeof(unsigned)));      /* starts | totals | ticket of k_bin_scan */
        CK(cudaMemset(LN.d_bin_start, 0, ((MAX_BINS + 2) * 2 + 4) * sizeof(unsigned)));
        CK(cudaEventCreateWithFlags(&LN.stage_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&LN.states_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&LN.fence, cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&LN.vstream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&LN.raw_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&LN.vready, cudaEventDisableTiming));
        CK(cudaHostAlloc(&LN.h_total, 2 * sizeof(unsigned), cudaHostAllocDefault));
        CK(cudaHostAlloc(&LN.h_list_total, sizeof(unsigned), cudaHostAllocDefault));
        CK(cudaEventCreateWithFlags(&LN.list_evt, cudaEventDisableTiming));
        CK(cudaMalloc(&LN.d_total, 64));
        CK(cudaMalloc(&LN.d_chain, 16 * sizeof(unsigned long long)));
        CK(cudaMemset(LN.d_chain, 0, 16 * sizeof(unsigned long long)));
    }
    RT.cur = &RT.lanes[0];
    {   /* band streams in descending priority: the block scheduler then drains band 0 first, band 1 next ... - launched
           with equal priority the bands would share the SMs evenly and all finish together at the end */
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));          /* lo: least (numerically largest), hi: greatest */
        for (int b = 0; b < MAX_BANDS; b++) {
            int pr = hi + b; if (pr > lo) pr = lo;
            CK(cudaStreamCreateWithPriority(&RT.band_streams[b], cudaStreamNonBlocking, pr));
        }
        /* the copy stream also runs kernels (the tile stores of a finished band to another device): above every band,
           or they would wait for the SMs until the whole surface is rasterised */
        CK(cudaStreamCreateWithPriority(&RT.copy_stream, cudaStreamNonBlocking, hi));
    }
    CK(cudaEventCreateWithFlags(&RT.front_evt, cudaEventDisableTiming));
    CK(cudaMalloc(&RT.d_counters, 8 * sizeof(unsigned long long)));          /* [0..3] the counters, [4..7] a sink (see counters_for) */
    CK(cudaMemset(RT.d_counters, 0, 8 * sizeof(unsigned long long)));
    RT.ok = true;
    if (!t_rt && n_listed > 1 && listed[0] == device) {
        /* start one worker per additional device; each initialises its own runtime and opens peer access to the primary
           (it will store its tiles into the primary's surfaces) */
        if (getenv("PF_CUDA_SPLIT_MIN_PIXELS")) mg.split_min_pixels = (size_t)strtoull(getenv("PF_CUDA_SPLIT_MIN_PIXELS"), nullptr, 10);
        mg.n = n_listed;
        for (int d = 0; d < n_listed; d++) mg.dev[d] = listed[d];
        for (int d = 1; d < n_listed; d++) { mg.rt[d] = new Runtime(); mg.rt[d]->index = d; mg.th[d] = std::thread(multi_worker, d); }
        const int rc = multi_run_all([](int d) -> int {
            if (d == 0) return (int)PFCU_OK;
            int rc = pfcu_init(mg.dev[d]);
            if (rc) return rc;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, mg.dev[d], mg.dev[0]) != cudaSuccess || !can) { snprintf(g_main.err, sizeof g_main.err, "device %d cannot access device %d as a peer", mg.dev[d], mg.dev[0]); return (int)PFCU_ERR_NO_DEVICE; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(mg.dev[0], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { snprintf(g_main.err, sizeof g_main.err, "cudaDeviceEnablePeerAccess(%d -> %d): %s", mg.dev[d], mg.dev[0], cudaGetErrorString(e)); return (int)PFCU_ERR_CUDA; }
            cudaGetLastError();
            return (int)PFCU_OK;
        });
        if (rc) {
            fprintf(stderr, "pixelforge-b200: PF_CUDA_DEVICES: %s - continuing on device %d alone\n", g_main.err, device);
            multi_stop();
        }
    }
    return PFCU_OK;
}

/* band and copy streams only ever hold work that a lane stream waits for (an event per band / per read-back), except
   the read-backs themselves: pfcu_surface_wait covers those, and so does this */
static void sync_all_lanes(void) { for (int i = 0; i < RT.n_lanes; i++) cudaStreamSynchronize(RT.lanes[i].stream); if (RT.copy_stream) cudaStreamSynchronize(RT.copy_stream); }
static void use_lane(const pfcu_surface *s)
{
    RT.cur = &RT.lanes[s ? s->lane % RT.n_lanes : 0];
    if (LN.need_fence_wait) {           /* deferred half of pfcu_fence(): later work on this lane comes after lane 0's fence */
        if (RT.cur != &RT.lanes[0]) cudaStreamWaitEvent(LN.stream, RT.lanes[0].fence, 0);
        LN.need_fence_wait = false;
    }
    LN.touched = true;
}

void pfcu_shutdown(void)
{
    API_LOCK;
    if (!RT.ok) return;
    if (!t_rt) multi_stop();
    sync_all_lanes();
    for (int i = 0; i < RT.n_lanes; i++) {
        RT.cur = &RT.lanes[i];
        cudaFree(LN.d_tris); cudaFree(LN.d_states); cudaFree(LN.d_bbox); cudaFree(LN.d_setup); cudaFree(LN.d_data);
        cudaFree(LN.d_bin_counts); cudaFree(LN.d_bin_list); cudaFree(LN.d_bin_start); cudaFree(LN.d_varrays); cudaFree(LN.d_vcounts);
        if (LN.h_stage) cudaFreeHost(LN.h_stage);
        if (LN.h_states) cudaFreeHost(LN.h_states);
        if (LN.h_total) cudaFreeHost(LN.h_total);
        if (LN.h_list_total) cudaFreeHost(LN.h_list_total);
        if (LN.list_evt) cudaEventDestroy(LN.list_evt);
        cudaFree(LN.d_raw); cudaFree(LN.d_total); cudaFree(LN.d_chain); cudaFree(LN.d_idx);
        for (cudaEvent_t e : { LN.stage_done, LN.states_done, LN.fence, LN.raw_done, LN.vready }) if (e) cudaEventDestroy(e);
        if (LN.vstream) cudaStreamDestroy(LN.vstream);
        if (LN.own_stream) cudaStreamDestroy(LN.stream);
        RT.lanes[i] = Lane();
    }
    for (int b = 0; b < MAX_BANDS; b++) if (RT.band_streams[b]) { cudaStreamDestroy(RT.band_streams[b]); RT.band_streams[b] = nullptr; }
    if (RT.copy_stream) { cudaStreamDestroy(RT.copy_stream); RT.copy_stream = nullptr; }
    if (RT.front_evt) { cudaEventDestroy(RT.front_evt); RT.front_evt = nullptr; }
    cudaFree(RT.d_counters); cudaFree(RT.d_rcp); cudaFree(RT.d_rsq);
    RT.d_counters = nullptr; RT.d_rcp = nullptr; RT.d_rsq = nullptr;
    for (auto &S : RT.job_slots) { cudaFree(S.d_tris); cudaFree(S.bbox); cudaFree(S.setup); cudaFree(S.data); cudaFree(S.bin_list); cudaFree(S.bin_start); cudaFree(S.d_total); cudaFree(S.chain); }
    RT.job_slots.clear();
    cudaFree(RT.d_jobs); if (RT.h_jobs) cudaFreeHost(RT.h_jobs);
    RT.d_jobs = nullptr; RT.h_jobs = nullptr; RT.cap_jobs = 0;
    if (RT.jobs_copied) { cudaEventDestroy(RT.jobs_copied); cudaEventDestroy(RT.jobs_done); RT.jobs_copied = RT.jobs_done = nullptr; }
    /* blocks handed out by pfcu_host_alloc and never returned: released here, their pointers die with the runtime */
    if (RT.index == 0) for (auto &b : RT.pinned) { if (b.done[0]) cudaEventDestroy(b.done[0]); cudaFreeHost(b.p); cudaFree(b.d_copy[0]); }
    for (auto e : RT.prof_events) cudaEventDestroy(e);
    for (auto e : RT.prof_pool) cudaEventDestroy(e);
    RT.prof_events.clear(); RT.prof_pool.clear();
    RT.pinned.clear(); RT.cur = nullptr; RT.ok = false;
}

/* Order everything enqueued so far on every lane before everything enqueued afterwards on every lane, without
 * blocking the host: used by callers that bracket multi-surface work with events on lane 0's stream. */
int pfcu_fence(void)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (MULTI_HERE()) return multi_run_all([](int) -> int { return pfcu_fence(); });
    /* lanes that were not used since the last fence have nothing new to order before lane 0; their wait for lane 0 is
       deferred to their next use (use_lane) - a single-surface workload pays for one event, not for every lane */
    for (int i = 1; i < RT.n_lanes; i++) {
        if (!RT.lanes[i].touched) continue;
        CK(cudaEventRecord(RT.lanes[i].fence, RT.lanes[i].stream));
        CK(cudaStreamWaitEvent(RT.lanes[0].stream, RT.lanes[i].fence, 0));
        RT.lanes[i].touched = false;
    }
    CK(cudaEventRecord(RT.lanes[0].fence, RT.lanes[0].stream));
    for (int i = 1; i < RT.n_lanes; i++) RT.lanes[i].need_fence_wait = true;
    return PFCU_OK;
}

int pfcu_set_stream(void *cuda_stream)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    sync_all_lanes();
    Lane &l0 = RT.lanes[0];
    if (l0.own_stream) { cudaStreamDestroy(l0.stream); l0.own_stream = false; }
    l0.stream = (cudaStream_t)cuda_stream;       /* lane 0 adopts the caller's stream; see pfcu_fence() */
    return PFCU_OK;
}

void *pfcu_get_stream(void) { return RT.ok ? (void *)RT.lanes[0].stream : nullptr; }

void *pfcu_host_alloc(size_t bytes)
{
    if (!RT.ok && pfcu_init(-1) != PFCU_OK) return nullptr;
    API_LOCK;
    PinnedBlock b; memset(&b, 0, sizeof b); b.bytes = bytes;
    for (int d = 0; d < MAX_DEVS; d++) b.dev_gen[d] = ~0u;
    /* portable: page-locked for every device of the multi-device mode */
    if (cudaHostAlloc(&b.p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    g_main.pinned.push_back(b);
    return b.p;
}

/* The registry of page-locked blocks lives in the primary runtime.  Worker threads only look blocks up, and only
 * while the dispatching thread holds the primary's lock (multi_run), so the vector does not change under them; each
 * device touches its own slots of a block. */
static PinnedBlock *find_pinned(const void *p)
{
    for (auto &b : g_main.pinned) if ((const char *)p >= (const char *)b.p && (const char *)p < (const char *)b.p + b.bytes) return &b;
    return nullptr;
}

/* the copy just enqueued on `st` reads from block b: the host may overwrite the block once this device is done with it */
static int pinned_in_flight(PinnedBlock *b, cudaStream_t st)
{
    const int d = RT.index;
    if (!b->done[d]) CK(cudaEventCreateWithFlags(&b->done[d], cudaEventDisableTiming));
    CK(cudaEventRecord(b->done[d], st));
    b->pending[d] = true;
    return PFCU_OK;
}

/* this device's event and mirror of a block (runs on the device's own thread) */
static int release_block_on_device(PinnedBlock *b)
{
    const int d = RT.index;
    if (b->pending[d] && b->done[d]) cudaEventSynchronize(b->done[d]);
    b->pending[d] = false;
    if (b->d_copy[d]) { sync_all_lanes(); cudaFree(b->d_copy[d]); b->d_copy[d] = nullptr; b->dev_gen[d] = ~0u; }
    return PFCU_OK;
}

void pfcu_host_free(void *p)
{
    API_LOCK;
    for (size_t i = 0; i < g_main.pinned.size(); i++) if (g_main.pinned[i].p == p) {
        PinnedBlock *b = &g_main.pinned[i];
        multi_run_all([b](int) -> int { release_block_on_device(b); if (b->done[RT.index]) { cudaEventDestroy(b->done[RT.index]); b->done[RT.index] = nullptr; } return (int)PFCU_OK; });
        cudaFreeHost(p);
        g_main.pinned.erase(g_main.pinned.begin() + i);
        return;
    }
}

int pfcu_host_set_static(void *p, int on)
{
    API_LOCK;
    PinnedBlock *b = find_pinned(p);
    if (!b) return PFCU_ERR_INVALID;
    b->is_static = on != 0;
    if (!on) multi_run_all([b](int) -> int { return release_block_on_device(b); });
    return PFCU_OK;
}

int pfcu_host_is_static(const void *p)
{
    API_LOCK;
    PinnedBlock *b = find_pinned(p);
    return b && b->is_static;
}

int pfcu_host_modified(void *p)
{
    API_LOCK;
    PinnedBlock *b = find_pinned(p);
    if (!b) return PFCU_ERR_INVALID;
    b->gen++;
    return PFCU_OK;
}

/* Device address of host array `p` (bytes long) when it lies in a static block: the block's device mirror, brought up to
 * date first if the application modified the block since the last upload.  nullptr: not static (the caller copies). */
static const unsigned char *static_mirror(const void *p, size_t bytes, PinnedBlock **blk)
{
    PinnedBlock *b = find_pinned(p);
    if (blk) *blk = b;
    if (!b || !b->is_static || (const char *)p + bytes > (const char *)b->p + b->bytes) return nullptr;
    const int d = RT.index;
    if (!b->d_copy[d] && cudaMalloc(&b->d_copy[d], b->bytes + 16) != cudaSuccess) { cudaGetLastError(); b->d_copy[d] = nullptr; return nullptr; }
    if (b->dev_gen[d] != b->gen) {
        /* every lane may be reading the old mirror; the upload is rare (once per modification), so it simply waits */
        sync_all_lanes();
        if (cudaMemcpy(b->d_copy[d], b->p, b->bytes, cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        RT.bytes_h2d += b->bytes;
        b->dev_gen[d] = b->gen;
    }
    return b->d_copy[d] + ((const char *)p - (const char *)b->p);
}

int pfcu_host_register(void *p, size_t bytes)
{
    API_LOCK;
    if (!RT.ok || !p || bytes == 0) return PFCU_ERR_INVALID;
    /* page-aligned sub-range; the partial first/last pages stay pageable (cudaMemcpy handles mixed ranges) */
    cudaError_t e = cudaHostRegister(p, bytes, mg.n > 1 ? (cudaHostRegisterPortable | cudaHostRegisterMapped) : cudaHostRegisterDefault);     /* multi-device: every device stores its tiles straight into the buffer */
    if (e != cudaSuccess) { cudaGetLastError(); return PFCU_ERR_CUDA; }
    return PFCU_OK;
}

void pfcu_host_unregister(void *p)
{
    API_LOCK;
    if (!RT.ok || !p) return;
    sync_all_lanes();
    if (cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();
}

int pfcu_host_wait(const void *p)
{
    API_LOCK;
    PinnedBlock *b = find_pinned(p);
    if (b) for (int d = 0; d < MAX_DEVS; d++)
        if (b->pending[d]) { CK(cudaEventSynchronize(b->done[d])); b->pending[d] = false; }       /* waiting on another device's event is fine */
    return PFCU_OK;
}

int pfcu_set_approx_tables(const uint32_t *rcp, int rcp_bits, const uint32_t *rsqrt, int rsqrt_bits)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (rcp_bits < 1 || rcp_bits > 23 || rsqrt_bits < 1 || rsqrt_bits > 23) return PFCU_ERR_INVALID;
    if (MULTI_HERE()) return multi_run_all([&](int) -> int { return pfcu_set_approx_tables(rcp, rcp_bits, rsqrt, rsqrt_bits); });
    sync_all_lanes();
    cudaFree(RT.d_rcp); cudaFree(RT.d_rsq);
    CK(cudaMalloc(&RT.d_rcp, sizeof(uint32_t) << rcp_bits));
    CK(cudaMalloc(&RT.d_rsq, sizeof(uint32_t) << (rsqrt_bits + 1)));
    CK(cudaMemcpy(RT.d_rcp, rcp, sizeof(uint32_t) << rcp_bits, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(RT.d_rsq, rsqrt, sizeof(uint32_t) << (rsqrt_bits + 1), cudaMemcpyHostToDevice));
    const int rshift = 23 - rcp_bits, sshift = 23 - rsqrt_bits;
    CK(cudaMemcpyToSymbol(c_rcp_tab, &RT.d_rcp, sizeof(void *)));
    CK(cudaMemcpyToSymbol(c_rsq_tab, &RT.d_rsq, sizeof(void *)));
    CK(cudaMemcpyToSymbol(c_rcp_shift, &rshift, sizeof(int)));
    CK(cudaMemcpyToSymbol(c_rsq_shift, &sshift, sizeof(int)));
    CK(cudaMemcpyToSymbol(c_rsq_bits, &rsqrt_bits, sizeof(int)));
    RT.rcp_bits = rcp_bits; RT.rsq_bits = rsqrt_bits;
    return PFCU_OK;
}

/* ---- surfaces ---- */

static void surface_dims(pfcu_surface *s) { s->tiles_x = (s->w + TILE - 1) / TILE; s->tiles_y = (s->h + TILE - 1) / TILE; }

static size_t fmt_bytes(int fmt) { return (fmt == PFCU_TEX_RGBA8 || fmt == PFCU_TEX_BGRA8) ? 4u : 3u; }

pfcu_surface *pfcu_surface_create(uint32_t w, uint32_t h) { return pfcu_surface_create_format(w, h, PFCU_TEX_RGBA8); }
int pfcu_surface_format(const pfcu_surface *s) { return s ? s->fmt : -1; }

pfcu_surface *pfcu_surface_create_format(uint32_t w, uint32_t h, int fmt)
{
    API_LOCK;
    if (MULTI_HERE()) {
        pfcu_surface *r[MAX_DEVS] = { nullptr };
        const int rc = multi_run_all([&](int d) -> int { r[d] = pfcu_surface_create_format(w, h, fmt); return r[d] ? PFCU_OK : PFCU_ERR_OOM; });
        if (rc) { multi_run_all([&](int d) -> int { if (r[d]) pfcu_surface_destroy(r[d]); return (int)PFCU_OK; }); return nullptr; }
        pfcu_surface *s = r[0];
        /* large RGBA8 targets are split by tiles; everything else is rendered in full on every device */
        s->split = fmt == PFCU_TEX_RGBA8 && (size_t)w * h >= mg.split_min_pixels && s->tiles_x * s->tiles_y >= (unsigned)mg.n;
        for (int d = 0; d < mg.n; d++) {
            s->rep[d] = r[d];
            r[d]->rank = s->split ? (uint32_t)d : 0u; r[d]->world = s->split ? (uint32_t)mg.n : 1u;
            if (d) { r[d]->peer_color = s->color; r[d]->peer_depth = s->depth; }      /* peer addresses: valid on every device (UVA + peer access) */
        }
        return s;
    }
    if (!RT.ok || w == 0 || h == 0) { snprintf(RT.err, sizeof RT.err, "surface_create: runtime not initialised or empty surface"); return nullptr; }
    if (fmt < PFCU_TEX_RGBA8 || fmt > PFCU_TEX_BGR8) { snprintf(RT.err, sizeof RT.err, "surface_create: unsupported colour format %d", fmt); return nullptr; }
    pfcu_surface *s = (pfcu_surface *)calloc(1, sizeof *s);
    if (!s) return nullptr;
    s->w = w; s->h = h; s->owned = true; s->world = 1; s->fmt = fmt;
    s->lane = (int)(RT.next_lane++ % (unsigned)RT.n_lanes);
    cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming);
    if (mg.n > 1) { cudaEventCreateWithFlags(&s->full_evt, cudaEventDisableTiming); cudaEventCreateWithFlags(&s->pushed, cudaEventDisableTiming); }
    use_lane(s);
    surface_dims(s);
    const size_t bytes = ((size_t)w * h + 64) * 4;
    if (cudaMalloc(&s->color, bytes) != cudaSuccess || cudaMalloc(&s->depth, bytes) != cudaSuccess) {
        snprintf(RT.err, sizeof RT.err, "surface_create: out of device memory (%ux%u)", w, h);
        cudaFree(s->color); free(s); return nullptr;
    }
    cudaMemsetAsync(s->color, 0, bytes, LN.stream);
    cudaMemsetAsync(s->depth, 0, bytes, LN.stream);
    if (fmt != PFCU_TEX_RGBA8) {
        s->conv_bytes = (size_t)w * h * fmt_bytes(fmt) + 16;
        if (cudaMalloc(&s->conv, s->conv_bytes) != cudaSuccess) {
            snprintf(RT.err, sizeof RT.err, "surface_create: out of device memory (%ux%u staging)", w, h);
            cudaFree(s->color); cudaFree(s->depth); free(s); return nullptr;
        }
    }
    return s;
}

pfcu_surface *pfcu_surface_wrap(void *dev_color, void *dev_depth, uint32_t w, uint32_t h)
{
    API_LOCK;
    if (!RT.ok || !dev_color || !dev_depth) return nullptr;
    pfcu_surface *s = (pfcu_surface *)calloc(1, sizeof *s);
    if (!s) return nullptr;
    s->w = w; s->h = h; s->color = (uint32_t *)dev_color; s->depth = (float *)dev_depth; s->owned = false; s->world = 1; s->fmt = PFCU_TEX_RGBA8;
    s->lane = 0;                                  /* caller-owned memory: stay on the caller-visible stream */
    cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming);
    surface_dims(s);
    return s;
}

static uint32_t owned_tiles(const pfcu_surface *s, uint32_t rank, uint32_t world);
static void mark_done(pfcu_surface *s);

static void close_present(pfcu_surface *s)
{
    if (s->ipc_color) cudaIpcCloseMemHandle(s->ipc_color);
    if (s->ipc_depth) cudaIpcCloseMemHandle(s->ipc_depth);
    s->ipc_color = s->ipc_depth = nullptr; s->peer_color = nullptr; s->peer_depth = nullptr;
}

int pfcu_surface_ipc_handles(pfcu_surface *s, void *color_handle, void *depth_handle)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || !color_handle) return PFCU_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle buffers are 64 bytes");
    CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)color_handle, s->color));
    if (depth_handle) CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)depth_handle, s->depth));
    return PFCU_OK;
}

int pfcu_surface_set_present_peer(pfcu_surface *s, const void *color_handle, const void *depth_handle)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || !color_handle) return PFCU_ERR_INVALID;
    sync_all_lanes();
    close_present(s);
    cudaIpcMemHandle_t h;
    memcpy(&h, color_handle, sizeof h);
    CK(cudaIpcOpenMemHandle(&s->ipc_color, h, cudaIpcMemLazyEnablePeerAccess));
    s->peer_color = (uint32_t *)s->ipc_color;
    if (depth_handle) {
        memcpy(&h, depth_handle, sizeof h);
        CK(cudaIpcOpenMemHandle(&s->ipc_depth, h, cudaIpcMemLazyEnablePeerAccess));
        s->peer_depth = (float *)s->ipc_depth;
    }
    return PFCU_OK;
}

int pfcu_surface_set_present_surface(pfcu_surface *s, pfcu_surface *target)
{
    API_LOCK;
    if (!s || !target || target->w != s->w || target->h != s->h || target == s) return PFCU_ERR_INVALID;
    if (RT.ok) sync_all_lanes();
    close_present(s);
    s->peer_color = target->color; s->peer_depth = target->depth;
    return PFCU_OK;
}

int pfcu_surface_clear_present(pfcu_surface *s)
{
    API_LOCK;
    if (!s) return PFCU_ERR_INVALID;
    if (RT.ok) sync_all_lanes();
    close_present(s);
    return PFCU_OK;
}

int pfcu_surface_push_tiles(pfcu_surface *s, uint32_t rank, uint32_t world, int with_depth)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || !s->peer_color || (with_depth && !s->peer_depth)) return PFCU_ERR_INVALID;
    if (world == 0) world = 1;
    use_lane(s);
    const uint32_t n = owned_tiles(s, rank, world);
    if (n == 0) return PFCU_OK;
    if (s->bands_valid && s->band_owned[s->n_bands] == n) {
        /* the surface was last written by a banded rasterisation: band b's tiles leave on the copy stream as soon as band b
           is done, while the later bands are still being rasterised; push_evt[b] tells the receiving side */
        for (int b = 0; b < s->n_bands; b++) {
            const uint32_t o0 = s->band_owned[b], o1 = s->band_owned[b + 1];
            CK(cudaStreamWaitEvent(RT.copy_stream, s->band_evt[b], 0));
            if (o1 > o0) {
                const unsigned cnt = o1 - o0, ctas = s->peer_is_host && cnt > PUSH_HOST_CTAS ? PUSH_HOST_CTAS : cnt;
                k_push_tiles<<<ctas, 256, 0, RT.copy_stream>>>(s->color, s->depth, s->peer_color, with_depth ? s->peer_depth : nullptr, (int)s->w, (int)s->h,
                                                              (int)s->tiles_x, s->tiles_x * s->tiles_y, rank, world, o0, cnt);
                RT.launches++;
            }
            if (!s->push_evt[b]) CK(cudaEventCreateWithFlags(&s->push_evt[b], cudaEventDisableTiming));
            CK(cudaEventRecord(s->push_evt[b], RT.copy_stream));
        }
        CK(cudaGetLastError());
        const bool keep = s->bands_valid;
        if (cudaEventRecord(s->done, RT.copy_stream) == cudaSuccess) s->has_done = true;
        CK(cudaStreamWaitEvent(LN.stream, s->done, 0));         /* later work on the surface must not overtake the stores' reads */
        s->bands_valid = keep; s->pushed_in_bands = true;
        return PFCU_OK;
    }
    s->pushed_in_bands = false;
    k_push_tiles<<<(s->peer_is_host && n > PUSH_HOST_CTAS) ? PUSH_HOST_CTAS : n, 256, 0, LN.stream>>>(s->color, s->depth, s->peer_color, with_depth ? s->peer_depth : nullptr, (int)s->w, (int)s->h,
                                          (int)s->tiles_x, s->tiles_x * s->tiles_y, rank, world, 0u, n);
    RT.launches++;
    CK(cudaGetLastError());
    mark_done(s);
    return PFCU_OK;
}

void pfcu_surface_destroy(pfcu_surface *s)
{
    API_LOCK;
    if (!s) return;
    if (MULTI_HERE() && s->rep[1]) {
        pfcu_surface *r[MAX_DEVS]; memcpy(r, s->rep, sizeof r);
        for (int d = 1; d < mg.n; d++) { r[d]->peer_color = nullptr; r[d]->peer_depth = nullptr; }      /* not IPC mappings: nothing to close */
        multi_run_all([&](int d) -> int { pfcu_surface_destroy(r[d]); return (int)PFCU_OK; });
        return;
    }
    if (RT.ok) sync_all_lanes();
    close_present(s);
    if (s->owned) { cudaFree(s->color); cudaFree(s->depth); }
    cudaFree(s->conv);
    if (s->done) cudaEventDestroy(s->done);
    for (int b = 0; b < MAX_BANDS; b++) if (s->band_evt[b]) cudaEventDestroy(s->band_evt[b]);
    for (int b = 0; b < MAX_BANDS; b++) if (s->push_evt[b]) cudaEventDestroy(s->push_evt[b]);
    if (s->full_evt) cudaEventDestroy(s->full_evt);
    if (s->pushed) cudaEventDestroy(s->pushed);
    if (s->t_front) cudaEventDestroy(s->t_front);
    if (s->t_raster) cudaEventDestroy(s->t_raster);
    free(s);
}

uint32_t pfcu_surface_width(const pfcu_surface *s) { return s->w; }
uint32_t pfcu_surface_height(const pfcu_surface *s) { return s->h; }
void *pfcu_surface_color_ptr(const pfcu_surface *s) { return s->color; }
void *pfcu_surface_depth_ptr(const pfcu_surface *s) { return s->depth; }

/* end of an operation on the surface (recorded on its lane); whatever it was, the band events of an earlier
   rasterisation no longer describe the surface's last write */
static void mark_done(pfcu_surface *s) { s->bands_valid = false; if (cudaEventRecord(s->done, LN.stream) == cudaSuccess) s->has_done = true; }

/* PF_CUDA_TIMING=1: host-side microseconds of the phases of pfcu_draw_triangles, band decisions of the multi-device read-back, on stderr (development aid) */
static const bool g_timing = getenv("PF_CUDA_TIMING") && atoi(getenv("PF_CUDA_TIMING")) != 0;
/* PF_CUDA_TIMING=1: device timeline of the last banded frame (pipeline start, front end done, band b rasterised, band b on the
   host), printed by pfcu_surface_wait; debugging aid, never on in measurements */
struct Timeline { cudaEvent_t start = nullptr, front = nullptr, band[MAX_BANDS] = {}, copied[MAX_BANDS] = {}; int n = 0; bool armed = false, copies = false; double host_start = 0; };
static Timeline g_tl;
static void tl_event(cudaEvent_t *e, cudaStream_t st) { if (!*e) cudaEventCreate(e); cudaEventRecord(*e, st); }
static double now_us(void);
static void tl_print(void)
{
    if (g_tl.armed && g_tl.copies) {
        float f = 0; cudaEventElapsedTime(&f, g_tl.start, g_tl.front);
        fprintf(stderr, "timeline: host enqueue -> wait returned %.3f ms; device: front end %.3f", (now_us() - g_tl.host_start) * 1e-3, f);
        for (int b = 0; b < g_tl.n; b++) {
            float r = 0, c = 0; cudaEventElapsedTime(&r, g_tl.start, g_tl.band[b]); cudaEventElapsedTime(&c, g_tl.start, g_tl.copied[b]);
            fprintf(stderr, " | band %d raster %.3f host %.3f", b, r, c);
        }
        fprintf(stderr, " ms\n");
        g_tl.armed = false;
    }
}
static double now_us(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; }

/* ---- multi-device helpers (see "multi-device mode" at the top) ---- */
#define MULTI_SURF(s) (MULTI_HERE() && (s) && (s)->rep[1])

/* op(d, replica) on every device; device 0 then notes that its surface was written outside its own tiles */
static int multi_surface_op(pfcu_surface *s, const std::function<int(int, pfcu_surface *)> &op)
{
    return multi_run_all([&](int d) -> int {
        const int rc = op(d, s->rep[d]);
        if (d == 0 && s->full_evt) { use_lane(s); if (cudaEventRecord(s->full_evt, LN.stream) == cudaSuccess) s->has_full = true; }
        return rc;
    });
}

/* Before device 0 reads a split surface back (or reads pixels out of it): every other device stores its own tiles into
 * device 0's buffers over NVLink, behind device 0's last full-surface write; device 0's lane waits for the stores. */
static int multi_gather(pfcu_surface *s, int with_depth)
{
    if (!s->split) return PFCU_OK;
    const int rc = multi_run_all([&](int d) -> int {
        if (d == 0) return (int)PFCU_OK;
        pfcu_surface *r = s->rep[d];
        use_lane(r);
        r->n_per_read = r->n_since_read; r->n_since_read = 0;       /* the replicas band their last batch like device 0 does */
        if (s->has_full) { CK(cudaStreamWaitEvent(LN.stream, s->full_evt, 0)); CK(cudaStreamWaitEvent(RT.copy_stream, s->full_evt, 0)); }
        const int rc = pfcu_surface_push_tiles(r, (uint32_t)d, (uint32_t)mg.n, with_depth);
        use_lane(r);
        CK(cudaEventRecord(r->pushed, LN.stream));
        return rc;
    });
    if (rc) return rc;
    use_lane(s);
    /* device 0 waits for the stores: band by band on its copy stream when everyone worked in the same bands (the read-back
       that follows then streams band b out while band b+1 is still being rasterised and stored), else on its lane */
    bool per_band = s->bands_valid;
    for (int d = 1; d < mg.n && per_band; d++) per_band = s->rep[d]->pushed_in_bands && s->rep[d]->n_bands == s->n_bands;
    if (g_timing) fprintf(stderr, "multi_gather: per_band %d (device 0 bands_valid %d, n_bands %d; device 1 pushed_in_bands %d n_bands %d)\n", (int)per_band, (int)s->bands_valid, s->n_bands, (int)s->rep[1]->pushed_in_bands, s->rep[1]->n_bands);
    s->peer_bands = per_band;       /* pfcu_surface_download_async: band b also waits for every device's push_evt[b] (in the
                                       copy stream, right before band b's copy - not all bands' waits ahead of the first copy) */
    if (per_band) {
        for (int d = 1; d < mg.n; d++) CK(cudaStreamWaitEvent(LN.stream, s->rep[d]->pushed, 0));
        return PFCU_OK;
    }
    for (int d = 1; d < mg.n; d++) CK(cudaStreamWaitEvent(LN.stream, s->rep[d]->pushed, 0));
    return PFCU_OK;
}

int pfcu_surface_upload(pfcu_surface *s, const void *hc, const float *hd, uint32_t y0, uint32_t rows)
{
    API_LOCK;
    if (MULTI_SURF(s)) return multi_surface_op(s, [&](int, pfcu_surface *r) -> int { return pfcu_surface_upload(r, hc, hd, y0, rows); });
    if (y0 > s->h || rows > s->h - y0) return PFCU_ERR_INVALID;
    use_lane(s);
    s->bands_valid = false;
    const size_t off = (size_t)y0 * s->w, n = (size_t)rows * s->w * 4;
    if (hc && s->fmt != PFCU_TEX_RGBA8 && rows) {
        /* the caller's layout -> staging -> canonical RGBA8 */
        const size_t bpp = fmt_bytes(s->fmt), nb = (size_t)rows * s->w * bpp;
        CK(cudaMemcpyAsync(s->conv + off * bpp, (const unsigned char *)hc + off * bpp, nb, cudaMemcpyHostToDevice, LN.stream));
        k_surface_convert<<<RT.sms * 4, 256, 0, LN.stream>>>(s->color + off, s->conv + off * bpp, (size_t)rows * s->w, s->fmt, 0);
        RT.launches++; RT.bytes_h2d += nb;
        CK(cudaGetLastError());
        hc = nullptr;
    }
    if (hc) { CK(cudaMemcpyAsync(s->color + off, (const uint32_t *)hc + off, n, cudaMemcpyHostToDevice, LN.stream)); RT.bytes_h2d += n; }
    if (hd) { CK(cudaMemcpyAsync(s->depth + off, hd + off, n, cudaMemcpyHostToDevice, LN.stream)); RT.bytes_h2d += n; }
    /* pageable sources are staged by the driver before the call returns; pinned ones are not */
    CK(cudaStreamSynchronize(LN.stream));
    return PFCU_OK;
}

int pfcu_surface_download_async(pfcu_surface *s, void *hc, float *hd, uint32_t y0, uint32_t rows)
{
    API_LOCK;
    if (MULTI_SURF(s)) {
        /* A whole-surface read-back of a split surface into page-locked memory that the devices can address: every device
           stores ITS tiles straight into the caller's buffer (the same kernel that stores them into device 0's surface,
           band by band behind the rasterisation) - N PCIe links instead of one, no gather.  Anything else: gather the
           tiles on device 0 first, then copy from there. */
        void *dc0 = nullptr, *dd0 = nullptr;
        /* from 4 devices on: with 2, one link still carries the frame faster by DMA behind the gather than two links
           carry it by stores (measured, 8K: 5.3 vs 5.9 ms); PF_CUDA_DIRECT_PRESENT=0|1 overrides */
        static const int env_direct = getenv("PF_CUDA_DIRECT_PRESENT") ? atoi(getenv("PF_CUDA_DIRECT_PRESENT")) : -1;
        bool direct = s->split && s->fmt == PFCU_TEX_RGBA8 && hc && y0 == 0 && rows == s->h && (env_direct < 0 ? mg.n >= 4 : env_direct != 0);
        if (direct && (cudaHostGetDevicePointer(&dc0, hc, 0) != cudaSuccess || (hd && cudaHostGetDevicePointer(&dd0, hd, 0) != cudaSuccess))) { cudaGetLastError(); direct = false; }
        if (direct) {
            int rc = multi_run_all([&](int d) -> int {
                pfcu_surface *r = s->rep[d];
                void *dc = nullptr, *dd = nullptr;
                if (cudaHostGetDevicePointer(&dc, hc, 0) != cudaSuccess || (hd && cudaHostGetDevicePointer(&dd, hd, 0) != cudaSuccess)) { cudaGetLastError(); return (int)PFCU_ERR_CUDA; }
                r->n_per_read = r->n_since_read; r->n_since_read = 0;
                uint32_t *const kc = r->peer_color; float *const kd = r->peer_depth;
                r->peer_color = (uint32_t *)dc; r->peer_depth = (float *)dd; r->peer_is_host = true;
                const int rc = pfcu_surface_push_tiles(r, (uint32_t)d, (uint32_t)mg.n, hd != nullptr);
                r->peer_color = kc; r->peer_depth = kd; r->peer_is_host = false;
                use_lane(r);
                CK(cudaEventRecord(r->pushed, LN.stream));
                RT.bytes_d2h += pfcu_surface_owned_bytes(r, (uint32_t)d, (uint32_t)mg.n, hd != nullptr);
                return rc;
            });
            if (rc == PFCU_OK) {
                use_lane(s);
                for (int d = 1; d < mg.n; d++) CK(cudaStreamWaitEvent(LN.stream, s->rep[d]->pushed, 0));
                const bool keep = s->bands_valid;
                mark_done(s);
                s->bands_valid = keep;
                return PFCU_OK;
            }
            cudaGetLastError();         /* fall through to the gather */
        }
        const int rc = multi_gather(s, hd != nullptr); if (rc) return rc;
    }
    if (y0 > s->h || rows > s->h - y0) return PFCU_ERR_INVALID;
    use_lane(s);
    const size_t off = (size_t)y0 * s->w, n = (size_t)rows * s->w * 4;
    if (hc && s->fmt != PFCU_TEX_RGBA8 && rows) {
        /* canonical RGBA8 -> staging in the caller's layout -> host */
        const size_t bpp = fmt_bytes(s->fmt), nb = (size_t)rows * s->w * bpp;
        k_surface_convert<<<RT.sms * 4, 256, 0, LN.stream>>>(s->color + off, s->conv + off * bpp, (size_t)rows * s->w, s->fmt, 1);
        RT.launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync((unsigned char *)hc + off * bpp, s->conv + off * bpp, nb, cudaMemcpyDeviceToHost, LN.stream)); RT.bytes_d2h += nb;
        hc = nullptr;
    }
    if (hc && rows) { s->n_per_read = s->n_since_read; s->n_since_read = 0; }
    if (s->bands_valid && s->fmt == PFCU_TEX_RGBA8 && (hc || hd) && rows) {
        /* the surface was last written by a banded rasterisation: band b's rows go out on the copy stream as soon as band
           b is done, while later bands are still being rasterised; the lane then waits for the copies (later work on
           the surface must not overtake them) */
        for (int b = 0; b < s->n_bands; b++) {
            const uint32_t r0 = s->band_y[b] > y0 ? s->band_y[b] : y0, r1 = s->band_y[b + 1] < y0 + rows ? s->band_y[b + 1] : y0 + rows;
            if (r0 >= r1) continue;
            const size_t o = (size_t)r0 * s->w, nb = (size_t)(r1 - r0) * s->w * 4;
            CK(cudaStreamWaitEvent(RT.copy_stream, s->band_evt[b], 0));
            if (s->peer_bands) for (int d = 1; d < mg.n; d++) CK(cudaStreamWaitEvent(RT.copy_stream, s->rep[d]->push_evt[b], 0));
            if (hc) { CK(cudaMemcpyAsync((uint32_t *)hc + o, s->color + o, nb, cudaMemcpyDeviceToHost, RT.copy_stream)); RT.bytes_d2h += nb; }
            if (hd) { CK(cudaMemcpyAsync(hd + o, s->depth + o, nb, cudaMemcpyDeviceToHost, RT.copy_stream)); RT.bytes_d2h += nb; }
            if (g_timing && g_tl.armed) { tl_event(&g_tl.copied[b], RT.copy_stream); g_tl.copies = true; }
        }
        if (cudaEventRecord(s->done, RT.copy_stream) == cudaSuccess) s->has_done = true;
        CK(cudaStreamWaitEvent(LN.stream, s->done, 0));
        return PFCU_OK;
    }
    if (hc) { CK(cudaMemcpyAsync((uint32_t *)hc + off, s->color + off, n, cudaMemcpyDeviceToHost, LN.stream)); RT.bytes_d2h += n; }
    if (hd) { CK(cudaMemcpyAsync(hd + off, s->depth + off, n, cudaMemcpyDeviceToHost, LN.stream)); RT.bytes_d2h += n; }
    mark_done(s);
    return PFCU_OK;
}

int pfcu_surface_wait(pfcu_surface *s)
{
    API_LOCK;
    if (!s) return PFCU_ERR_INVALID;
    if (s->has_done) CK(cudaEventSynchronize(s->done));
    if (g_timing) tl_print();
    return PFCU_OK;
}

int pfcu_surface_download(pfcu_surface *s, void *hc, float *hd, uint32_t y0, uint32_t rows)
{
    API_LOCK;
    int rc = pfcu_surface_download_async(s, hc, hd, y0, rows);
    if (rc) return rc;
    CK(cudaStreamSynchronize(LN.stream));
    if (g_timing) tl_print();
    return PFCU_OK;
}

static int fill_range(pfcu_surface *s, size_t first, size_t n, int dc, uint32_t rgba, int dd, float z)
{
    if (n == 0) return PFCU_OK;
    /* head (to 4-pixel alignment), vector body, tail */
    size_t head = (4 - (first & 3)) & 3; if (head > n) head = n;
    const size_t body4 = (n - head) / 4, tail = n - head - body4 * 4;
    const int blocks = RT.sms * 8;
    if (head) { k_fill<<<1, 32, 0, LN.stream>>>(s->color, s->depth, first, head, dc, rgba, dd, z); RT.launches++; }
    if (body4) {
        k_fill4<<<blocks, 256, 0, LN.stream>>>((uint4 *)s->color, (float4 *)s->depth, (first + head) / 4, body4, dc, rgba, dd, z);
        RT.launches++;
    }
    if (tail) { k_fill<<<1, 32, 0, LN.stream>>>(s->color, s->depth, first + head + body4 * 4, tail, dc, rgba, dd, z); RT.launches++; }
    CK(cudaGetLastError());
    return PFCU_OK;
}

int pfcu_surface_fill(pfcu_surface *s, int dc, uint32_t rgba, int dd, float z)
{
    API_LOCK;
    if (MULTI_SURF(s)) return multi_surface_op(s, [&](int, pfcu_surface *r) -> int { return pfcu_surface_fill(r, dc, rgba, dd, z); });
    use_lane(s);
    if (s->fmt >= PFCU_TEX_RGB8) rgba |= 0xff000000u;      /* 3-byte targets store no alpha and read it back as 255 */
    const int rc = fill_range(s, 0, (size_t)s->w * s->h, dc, rgba, dd, z);
    mark_done(s);
    return rc;
}

int pfcu_surface_clear_ref(pfcu_surface *s, int dc, uint32_t rgba, int dd, float z)
{
    API_LOCK;
    if (MULTI_SURF(s)) return multi_surface_op(s, [&](int, pfcu_surface *r) -> int { return pfcu_surface_clear_ref(r, dc, rgba, dd, z); });
    use_lane(s);
    if (s->fmt >= PFCU_TEX_RGB8) rgba |= 0xff000000u;
    const unsigned size = s->w * s->h, aligned = size - (size % 8u);
    if (s->world > 1 && aligned == size && s->fmt == PFCU_TEX_RGBA8) {
        /* tile split: only the tiles this rank rasterises (an eighth of an 8K clear instead of all of it on every GPU) */
        const uint32_t n = owned_tiles(s, s->rank, s->world);
        if (n) { k_clear_tiles<<<n, 256, 0, LN.stream>>>(s->color, s->depth, (int)s->w, (int)s->h, (int)s->tiles_x, s->tiles_x * s->tiles_y, s->rank, s->world, dc, rgba, dd, z); RT.launches++; }
        CK(cudaGetLastError());
        mark_done(s);
        return PFCU_OK;
    }
    if (aligned > 8) { int rc = fill_range(s, 8, aligned - 8, dc, rgba, dd, z); if (rc) return rc; }
    if (aligned < size) { k_clear_tail<<<1, 32, 0, LN.stream>>>(s->color, s->depth, aligned, size, dc, dd); RT.launches++; }
    CK(cudaGetLastError());
    mark_done(s);
    return PFCU_OK;
}

/* ---- full-surface operations ---- */

int pfcu_surface_rect(pfcu_surface *s, int32_t x1, int32_t y1, int32_t x2, int32_t y2, uint32_t rgba)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (!s) return PFCU_ERR_INVALID;
    if (x2 < x1 || y2 < y1) return PFCU_OK;
    if (MULTI_SURF(s)) return multi_surface_op(s, [&](int, pfcu_surface *r) -> int { return pfcu_surface_rect(r, x1, y1, x2, y2, rgba); });
    use_lane(s);
    if (s->fmt >= PFCU_TEX_RGB8) rgba |= 0xff000000u;
    const uint32_t cols = (uint32_t)(x2 - x1) + 1u, rows = (uint32_t)(y2 - y1) + 1u;
    const size_t n = (size_t)cols * rows;
    const unsigned blocks = (unsigned)((n + 255) / 256 < (size_t)RT.sms * 8 ? (n + 255) / 256 : (size_t)RT.sms * 8);
    k_rect<<<blocks, 256, 0, LN.stream>>>(s->color, s->w, s->w * s->h, x1, y1, cols, rows, rgba);
    RT.launches++;
    CK(cudaGetLastError());
    mark_done(s);
    return PFCU_OK;
}

int pfcu_surface_fog(pfcu_surface *s, const pfcu_fog *f)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || !f || f->mode > 3u || f->n_thresholds > 255u || (f->mode != 0u && f->n_thresholds && !f->thresholds)) return PFCU_ERR_INVALID;
    if (MULTI_SURF(s)) return multi_surface_op(s, [&](int, pfcu_surface *r) -> int { return pfcu_surface_fog(r, f); });
    use_lane(s);
    int rc;
    const float *d_thr = nullptr;
    if ((f->mode == 1u || f->mode == 2u) && f->n_thresholds) {
        if ((rc = grow(&LN.d_varrays, &LN.cap_varrays, 256 * sizeof(float)))) return rc;
        /* pageable source: staged by the driver before the call returns */
        CK(cudaMemcpyAsync(LN.d_varrays, f->thresholds, f->n_thresholds * sizeof(float), cudaMemcpyHostToDevice, LN.stream));
        RT.bytes_h2d += f->n_thresholds * sizeof(float);
        d_thr = (const float *)LN.d_varrays;
    }
    FogArgs a;
    a.start = f->start; a.end = f->end; a.inv_len = f->inv_len; a.rgba = f->rgba; a.mode = f->mode;
    a.n_thr = (f->mode == 1u || f->mode == 2u) ? f->n_thresholds : 0u;       /* mode 3: t = 0 for every pixel in range */ a.alpha_or = s->fmt >= PFCU_TEX_RGB8 ? 0xff000000u : 0u;
    k_fog<<<RT.sms * 8, 256, 0, LN.stream>>>(s->color, s->depth, (size_t)s->w * s->h, a, d_thr);
    RT.launches++;
    CK(cudaGetLastError());
    mark_done(s);
    return PFCU_OK;
}

int pfcu_surface_draw_pixels(pfcu_surface *s, const pfcu_pixels *d)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || !d || !d->pixels || d->width == 0 || d->height == 0 || pfx_bytes(d->format) == 0) return PFCU_ERR_INVALID;
    if (d->xmax < d->xmin || d->ymax < d->ymin) return PFCU_OK;
    if (MULTI_SURF(s)) return multi_surface_op(s, [&](int, pfcu_surface *r) -> int { return pfcu_surface_draw_pixels(r, d); });
    use_lane(s);
    int rc;
    const size_t bytes = (size_t)d->width * d->height * (size_t)pfx_bytes(d->format);
    if ((rc = grow(&LN.d_varrays, &LN.cap_varrays, bytes + 16))) return rc;
    CK(cudaMemcpyAsync(LN.d_varrays, d->pixels, bytes, cudaMemcpyHostToDevice, LN.stream));
    /* the caller may reuse its image as soon as pfDrawPixels returns: page-locked sources are still being read */
    if (find_pinned(d->pixels)) CK(cudaStreamSynchronize(LN.stream));
    RT.bytes_h2d += bytes;
    PixArgs a;
    a.src = LN.d_varrays; a.sw = d->width; a.sh = d->height; a.fmt = d->format;
    a.xs = d->xs; a.ys = d->ys; a.xmin = d->xmin; a.ymin = d->ymin;
    a.cols = (uint32_t)(d->xmax - d->xmin) + 1u; a.rows = (uint32_t)(d->ymax - d->ymin) + 1u;
    a.inv_xlen = d->inv_xlen; a.inv_ylen = d->inv_ylen; a.z = d->z;
    a.flags = d->flags; a.blend_mode = d->blend_mode; a.depth_func = d->depth_func;
    a.alpha_or = s->fmt >= PFCU_TEX_RGB8 ? 0xff000000u : 0u;
    const size_t n = (size_t)a.cols * a.rows;
    const unsigned blocks = (unsigned)((n + 255) / 256 < (size_t)RT.sms * 8 ? (n + 255) / 256 : (size_t)RT.sms * 8);
    k_draw_pixels<<<blocks, 256, 0, LN.stream>>>(s->color, s->depth, s->w, s->w * s->h, a);
    RT.launches++;
    CK(cudaGetLastError());
    mark_done(s);
    return PFCU_OK;
}

int pfcu_surface_read_pixels(pfcu_surface *s, uint32_t x0, uint32_t y0, uint32_t cols, uint32_t rows, uint32_t dst_width, int format, void *host_pixels)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || !host_pixels || pfx_bytes(format) == 0) return PFCU_ERR_INVALID;
    if (cols == 0 || rows == 0) return PFCU_OK;
    if (x0 >= s->w || y0 >= s->h || cols > s->w - x0 || rows > s->h - y0 || cols > dst_width) return PFCU_ERR_INVALID;
    if (MULTI_SURF(s)) { const int rc = multi_gather(s, 0); if (rc) return rc; }
    use_lane(s);
    int rc;
    const size_t bpp = (size_t)pfx_bytes(format), n = (size_t)cols * rows;
    if ((rc = grow(&LN.d_varrays, &LN.cap_varrays, n * bpp + 16))) return rc;
    const unsigned blocks = (unsigned)((n + 255) / 256 < (size_t)RT.sms * 8 ? (n + 255) / 256 : (size_t)RT.sms * 8);
    k_read_pixels<<<blocks, 256, 0, LN.stream>>>(s->color, s->w, x0, y0, cols, rows, format, LN.d_varrays);
    RT.launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpy2DAsync(host_pixels, (size_t)dst_width * bpp, LN.d_varrays, (size_t)cols * bpp, (size_t)cols * bpp, rows, cudaMemcpyDeviceToHost, LN.stream));
    RT.bytes_d2h += n * bpp;
    CK(cudaStreamSynchronize(LN.stream));
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
    if (unpack) s->bands_valid = false;
    const uint32_t n = owned_tiles(s, rank, world);
    if (n == 0) return PFCU_OK;
    k_pack_tiles<<<n, 256, 0, LN.stream>>>(s->color, s->depth, (int)s->w, (int)s->h, (int)s->tiles_x, s->tiles_x * s->tiles_y,
                                          rank, world, with_depth, (uint32_t *)staging, unpack);
    RT.launches++;
    CK(cudaGetLastError());
    return PFCU_OK;
}

int pfcu_surface_pack_tiles(pfcu_surface *s, uint32_t r, uint32_t w, int wd, void *st) { return pack_unpack(s, r, w, wd, st, 0); }
int pfcu_surface_unpack_tiles(pfcu_surface *s, uint32_t r, uint32_t w, int wd, const void *st) { return pack_unpack(s, r, w, wd, (void *)st, 1); }

/* ---- textures ---- */

static bool tex_fmt_ok(int fmt) { return (fmt >= PFCU_TEX_RGBA8 && fmt <= PFCU_TEX_BGR8) || (fmt >= PFCU_TEX_PIX && pfx_bytes(fmt - PFCU_TEX_PIX) != 0); }
static size_t tex_bytes(uint32_t w, uint32_t h, int fmt)
{
    if (fmt >= PFCU_TEX_PIX) return (size_t)w * h * (size_t)pfx_bytes(fmt - PFCU_TEX_PIX);
    return (size_t)w * h * ((fmt == PFCU_TEX_RGBA8 || fmt == PFCU_TEX_BGRA8) ? 4u : 3u);
}

pfcu_texture *pfcu_texture_create(const void *host_pixels, uint32_t w, uint32_t h, int fmt)
{
    API_LOCK;
    if (MULTI_HERE()) {
        pfcu_texture *r[MAX_DEVS] = { nullptr };
        const int rc = multi_run_all([&](int d) -> int { r[d] = pfcu_texture_create(host_pixels, w, h, fmt); return r[d] ? PFCU_OK : PFCU_ERR_OOM; });
        if (rc) { multi_run_all([&](int d) -> int { if (r[d]) pfcu_texture_destroy(r[d]); return (int)PFCU_OK; }); return nullptr; }
        for (int d = 0; d < mg.n; d++) r[0]->rep[d] = r[d];
        return r[0];
    }
    if (!RT.ok || !tex_fmt_ok(fmt) || w == 0 || h == 0) return nullptr;
    pfcu_texture *t = (pfcu_texture *)calloc(1, sizeof *t);
    if (!t) return nullptr;
    t->w = w; t->h = h; t->fmt = fmt; t->owned = true; t->leader = (fmt == PFCU_TEX_BGRA8);
    RT.cur = &RT.lanes[0];
    const size_t bytes = tex_bytes(w, h, fmt);
    if (cudaMalloc(&t->pixels, bytes + 16) != cudaSuccess) { snprintf(RT.err, sizeof RT.err, "texture_create: out of device memory"); free(t); return nullptr; }
    cudaMemsetAsync(t->pixels, 0, bytes + 16, LN.stream);
    if (host_pixels && pfcu_texture_update(t, host_pixels) != PFCU_OK) { cudaFree(t->pixels); free(t); return nullptr; }
    return t;
}

pfcu_texture *pfcu_texture_from_surface(pfcu_surface *s)
{
    API_LOCK;
    if (MULTI_SURF(s)) {
        /* a surface that is sampled must be whole on every device: from now on each device renders all of it */
        if (s->split) {
            /* keep what was rendered so far: gather the tiles on device 0, then hand every other device the whole surface */
            int rc = multi_gather(s, 1);
            if (rc) return nullptr;
            use_lane(s);
            if (cudaEventRecord(s->full_evt, LN.stream) != cudaSuccess) return nullptr;
            s->has_full = true;
            const size_t bytes = (size_t)s->w * s->h * 4;
            rc = multi_run_all([&](int d) -> int {
                if (d == 0) return (int)PFCU_OK;
                pfcu_surface *r = s->rep[d];
                use_lane(r);
                CK(cudaStreamWaitEvent(LN.stream, s->full_evt, 0));
                CK(cudaMemcpyPeerAsync(r->color, mg.dev[d], s->color, mg.dev[0], bytes, LN.stream));
                CK(cudaMemcpyPeerAsync(r->depth, mg.dev[d], s->depth, mg.dev[0], bytes, LN.stream));
                mark_done(r);
                CK(cudaEventRecord(r->pushed, LN.stream));
                return (int)PFCU_OK;
            });
            if (rc) return nullptr;
            use_lane(s);
            for (int d = 1; d < mg.n; d++) if (cudaStreamWaitEvent(LN.stream, s->rep[d]->pushed, 0) != cudaSuccess) return nullptr;   /* device 0 must not overwrite what is being copied */
            s->split = false;
            for (int d = 0; d < mg.n; d++) { s->rep[d]->rank = 0; s->rep[d]->world = 1; }
        }
        pfcu_texture *r[MAX_DEVS] = { nullptr };
        const int rc = multi_run_all([&](int d) -> int { r[d] = pfcu_texture_from_surface(s->rep[d]); return r[d] ? PFCU_OK : PFCU_ERR_OOM; });
        if (rc) return nullptr;
        for (int d = 0; d < mg.n; d++) r[0]->rep[d] = r[d];
        return r[0];
    }
    pfcu_texture *t = (pfcu_texture *)calloc(1, sizeof *t);
    if (!t) return nullptr;
    /* the surface is held as canonical RGBA8 whatever the caller's layout; a BGRA8 target sampled as a texture still goes
       through the reference's BGRA8 getter, whose effect beyond the byte order is the leader replication */
    t->w = s->w; t->h = s->h; t->fmt = PFCU_TEX_RGBA8; t->pixels = (unsigned char *)s->color; t->owned = false; t->alias = s;
    t->leader = (s->fmt == PFCU_TEX_BGRA8);
    s->aliased = true;
    return t;
}

int pfcu_texture_update(pfcu_texture *t, const void *host_pixels)
{
    API_LOCK;
    if (!t || !t->owned || !host_pixels) return PFCU_ERR_INVALID;
    if (MULTI_HERE() && t->rep[1]) return multi_run_all([&](int d) -> int { return pfcu_texture_update(t->rep[d], host_pixels); });
    sync_all_lanes();                           /* nobody may still be sampling the old texels */
    RT.cur = &RT.lanes[0];
    CK(cudaMemcpyAsync(t->pixels, host_pixels, tex_bytes(t->w, t->h, t->fmt), cudaMemcpyHostToDevice, LN.stream));
    RT.bytes_h2d += tex_bytes(t->w, t->h, t->fmt);
    CK(cudaStreamSynchronize(LN.stream));
    return PFCU_OK;
}

void pfcu_texture_destroy(pfcu_texture *t)
{
    API_LOCK;
    if (!t) return;
    if (MULTI_HERE() && t->rep[1]) {
        pfcu_texture *r[MAX_DEVS]; memcpy(r, t->rep, sizeof r);
        multi_run_all([&](int d) -> int { pfcu_texture_destroy(r[d]); return (int)PFCU_OK; });
        return;
    }
    if (RT.ok) sync_all_lanes();
    if (t->owned) cudaFree(t->pixels);
    free(t);
}

/* ---- the hot path ---- */

/* state program of a DevState: texm * 4 + blendm with texm 0 none, 1 nearest + REPEAT + RGBA8, 3 nearest (any wrap mode /
   texel layout), 4 bilinear - the FIXED_PROG numbering of k_raster (pfcu_raster_tiles.cuh); per-fragment Phong: PROG_PHONG */
#define PROG_PHONG 20
static int state_program(const DevState *d)
{
    if (d->flags & PFCU_ST_PHONG) return PROG_PHONG;
    if ((d->flags & PFCU_ST_TEXTURE) && d->tfmt >= PFCU_TEX_PIX) return -1;       /* the other texel layouts: run-time sampler only */
    int texm = 0;
    if (d->flags & PFCU_ST_TEXTURE) texm = d->tex_filter != 0 ? 4 : ((d->tfmt == PFCU_TEX_RGBA8 && d->tex_wrap == 0) ? 1 : 3);
    const int blendm = !(d->flags & PFCU_ST_BLEND) ? 0 : (d->blend_mode == 1 ? 1 : (d->blend_mode == 2 ? 2 : 3));
    return texm * 4 + blendm;
}


/* what convert_states found, for the launch_pipeline call that follows it in the same entry point; per thread: in the
   multi-device mode every device's worker thread converts and launches on its own */
static thread_local int g_last_single_prog = -1;      /* the common program of all states, or -1 */
static thread_local bool g_last_leader_tex = false;   /* some state samples through the BGRA8 getter */
static thread_local bool g_last_pix_tex = false;      /* some state samples a texture beyond the four 8-bit layouts */

static unsigned convert_states(const pfcu_state *in, uint32_t n, DevState *out)
{
    unsigned mask = 0;
    RT.deps.clear();
    g_last_leader_tex = false; g_last_pix_tex = false;
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
            d->tex_leader = s->texture->leader ? 1u : 0u;
            if (s->texture->leader) g_last_leader_tex = true;
            if (s->texture->fmt >= PFCU_TEX_PIX) g_last_pix_tex = true;
            d->tex_fw = (float)d->tw; d->tex_fh = (float)d->th;
            { volatile float one = 1.0f; d->tex_tx = one / d->tex_fw; d->tex_ty = one / d->tex_fh; }   /* IEEE single division, as DIVPS */
            if (s->texture->alias) RT.deps.push_back(s->texture->alias);      /* render-to-texture: order across lanes */
        }
        d->n_lights = s->n_lights > 8 ? 8 : s->n_lights;
        for (unsigned l = 0; l < d->n_lights; l++) {
            const pfcu_light *a = &s->lights[l]; DevLight *b = &d->lights[l];
            memcpy(b->pos, a->position, 12); memcpy(b->dir, a->direction, 12);
            b->inner = a->inner_cutoff; b->outer = a->outer_cutoff;
            b->attc = a->att_constant; b->attl = a->att_linear; b->attq = a->att_quadratic;
            b->ambient = a->ambient; b->diffuse = a->diffuse; b->specular = a->specular;
            for (int c = 0; c < 3; c++) {
                volatile float k = 1.0f / 255.0f;       /* one rounded multiply per channel, as the reference's lanes */
                b->amb_f[c] = (float)((a->ambient >> (8 * c)) & 255u) * k;
                b->dif_f[c] = (float)((a->diffuse >> (8 * c)) & 255u) * k;
                b->spc_f[c] = (float)((a->specular >> (8 * c)) & 255u) * k;
            }
        }
        for (int f = 0; f < 2; f++) {
            d->material[f].ambient = s->material[f].ambient; d->material[f].diffuse = s->material[f].diffuse;
            d->material[f].specular = s->material[f].specular; d->material[f].emission = s->material[f].emission;
            d->material[f].shininess = s->material[f].shininess;
            for (int c = 0; c < 3; c++) {
                volatile float k = 1.0f / 255.0f;
                d->material[f].amb_f[c] = (float)((s->material[f].ambient >> (8 * c)) & 255u) * k;
                d->material[f].spc_f[c] = (float)((s->material[f].specular >> (8 * c)) & 255u) * k;
            }
        }
        memcpy(d->view_pos, s->view_pos, 12);
        mask |= d->flags;
        /* the program every state of the batch runs, or their join where one exists: blend modes that differ become
           "any mode" (run-time switch, blending on in all of them), nearest samplers that differ become "nearest, any wrap
           mode / layout" - e.RT. layers that alternate between alpha and additive blending still get a one-program kernel */
        const int prog = state_program(d);
        if (i == 0 || prog < 0) g_last_single_prog = prog;
        else if (g_last_single_prog != prog && g_last_single_prog >= 0) {
            const int a = g_last_single_prog, b = prog;
            int texm = -1, blendm = -1;
            if (a < PROG_PHONG && b < PROG_PHONG) {
                const int ta = a / 4, tb = b / 4, ba = a % 4, bb = b % 4;
                texm = ta == tb ? ta : (((ta == 1 || ta == 3) && (tb == 1 || tb == 3)) ? 3 : -1);
                blendm = ba == bb ? ba : ((ba != 0 && bb != 0) ? 3 : -1);
            }
            g_last_single_prog = (texm >= 0 && blendm >= 0) ? texm * 4 + blendm : -1;
        }
    }
    return mask;
}

/* n: number of triangles, or - when d_n is given - an upper bound of the count the device holds in *d_n (raw batches
 * after clipping); n_est then stands in for n where only the batch shape matters. */
static int launch_pipeline(pfcu_surface *s, const pfcu_triangle *d_tris, const DevState *d_states, uint32_t n, unsigned feature_mask, int single_prog = -1,
                           const unsigned *d_n = nullptr, uint32_t n_est = 0, int bshift_forced = 0)
{
    if (n == 0) return PFCU_OK;
    s->bands_valid = false;
    if (!d_n) n_est = n;
    if (g_timing && !g_tl.armed) { use_lane(s); tl_event(&g_tl.start, LN.stream); g_tl.host_start = now_us(); }
    /* render targets other than RGBA8 and BGRA8 textures: the row-ordered rasteriser (pfcu_raster_rows.cuh) */
    const bool rows_path = s->fmt != PFCU_TEX_RGBA8 || g_last_leader_tex;
    /* textures beyond the four 8-bit layouts: the tile rasteriser's run-time sampler (or the row-ordered one) has their getters */
    const bool pix_tex = g_last_pix_tex;
    g_last_leader_tex = false; g_last_pix_tex = false;
    if (rows_path && s->world > 1) {
        snprintf(RT.err, sizeof RT.err, "the screen-tile split needs an RGBA8 target and no BGRA8 textures (a pixel depends on its row neighbours there)");
        return PFCU_ERR_INVALID;
    }
    if (!RT.d_rcp) { snprintf(RT.err, sizeof RT.err, "pfcu_set_approx_tables() has not been called"); return PFCU_ERR_INVALID; }
    int rc;
    /* surfaces sampled as textures that live on another lane: wait for their last write */
    for (pfcu_surface *dep : RT.deps)
        if (dep != s && dep->lane != s->lane && dep->has_done) CK(cudaStreamWaitEvent(LN.stream, dep->done, 0));
    if (n > LN.cap_setup) {
        size_t c1 = LN.cap_setup, c2 = LN.cap_setup, c3 = LN.cap_setup;
        /* any failure leaves the three arrays with unknown capacities: forget them all, so that the next batch grows
           every one of them again instead of writing through a pointer that grow() has already freed */
        if ((rc = grow(&LN.d_bbox, &c1, n)) || (rc = grow(&LN.d_setup, &c2, n)) || (rc = grow(&LN.d_data, &c3, n))) { LN.cap_setup = 0; return rc; }
        LN.cap_setup = c1;
    }
    /* many small triangles per tile: fine bins (one per tile) and the fragment-compacting rasteriser;
       few large ones: coarse bins and the triangle-per-warp-step rasteriser */
    const unsigned nTilesAll = s->tiles_x * s->tiles_y;
    const bool small_tris = (size_t)n_est > (size_t)4 * nTilesAll;
    static const int force_bshift = getenv("PF_CUDA_BIN_SHIFT") ? atoi(getenv("PF_CUDA_BIN_SHIFT")) : 0;
    int bshift = bshift_forced ? bshift_forced : (force_bshift >= 6 ? force_bshift : (small_tris ? BIN_SHIFT_FINE : BIN_SHIFT_COARSE));
    static const bool env_once = [] {
        const char *e = getenv("PF_CUDA_FRAG");
        if (e) RT.raster_path = atoi(e) == 0 ? PFCU_RASTER_TILES : (atoi(e) == 2 ? PFCU_RASTER_FRAGMENTS : PFCU_RASTER_AUTO);
        return true; }();
    (void)env_once;
    static const int force_slice = getenv("PF_CUDA_SLICE") ? atoi(getenv("PF_CUDA_SLICE")) : 0;
    const bool use_frag = !rows_path && !pix_tex && (RT.raster_path == PFCU_RASTER_FRAGMENTS || (RT.raster_path == PFCU_RASTER_AUTO && small_tris && !force_slice));
    /* The fragment rasteriser works on 64x8 slices, and with square 64x64 bins the eight slices of a tile each filter the
       whole tile's list.  Flatter bins (64 x 2^bshy) shorten that, but every triangle then lands in more bins and the
       ordered fill pays for it: measured (PF_CUDA_BIN_ROWS sweep, profiles/README.md) 64x16 bins are a small net gain
       only for very large batches (the 1 M-triangle mesh: raster 0.475 -> 0.441 ms, binning +0.024 ms) and a loss below. */
    static const int force_bshy = getenv("PF_CUDA_BIN_ROWS") ? atoi(getenv("PF_CUDA_BIN_ROWS")) : 0;
    int bshy = bshift;
    if (use_frag && !bshift_forced && bshift == BIN_SHIFT_FINE)
        bshy = (force_bshy >= 3 && force_bshy <= 6) ? force_bshy : (n_est >= (1u << 19) ? 4 : BIN_SHIFT_FINE);
    int binsX, binsY, nb;
    for (;;) {
        binsX = (int)((s->w + (1u << bshift) - 1) >> bshift); binsY = (int)((s->h + (1u << bshy) - 1) >> bshy);
        nb = binsX * binsY;
        if (nb <= MAX_BINS) break;
        if (bshy < bshift) bshy++; else { bshift++; bshy++; }
    }
    if (bshift > 8) { snprintf(RT.err, sizeof RT.err, "surface too large for the binner (%u x %u)", s->w, s->h); return PFCU_ERR_INVALID; }
    /* one row of per-bin counters per binning CTA: 256 triangles per CTA give the order-preserving fill four times the
       CTAs (its per-CTA work is a serial chain) as long as the counter matrix stays small */
    unsigned bin_batch = ((size_t)((n + 255u) / 256u) * nb <= ((size_t)1 << 20)) ? 256u : BIN_BATCH;
    while ((size_t)((n + bin_batch - 1) / bin_batch) * nb > ((size_t)4 << 20) && bin_batch < 8192u) bin_batch *= 2;    /* keep the counter matrix below 16 MB */
    const unsigned nBatches = (n + bin_batch - 1) / bin_batch;
    if ((rc = grow(&LN.d_bin_counts, &LN.cap_bin_counts, (size_t)nBatches * nb))) return rc;

    size_t list_cap = ~(size_t)0;          /* capacity the kernels check the real total against (small batches: sized by the bound) */
    cudaEvent_t pe[3] = { nullptr, nullptr, nullptr };
    if (RT.profiling) {
        for (int i = 0; i < 3; i++) {
            if (!RT.prof_pool.empty()) { pe[i] = RT.prof_pool.back(); RT.prof_pool.pop_back(); }
            else CK(cudaEventCreate(&pe[i]));
        }
        CK(cudaEventRecord(pe[0], LN.stream));
    }
    if (d_n && !(nb <= 3072 && n <= FRONT_SMALL_MAX * FRONT_SMALL_CHUNKS)) { snprintf(RT.err, sizeof RT.err, "internal: device-side count outside the single-CTA front end"); return PFCU_ERR_INVALID; }
    bool list_total_wanted = false;
    if (d_n || (n <= FRONT_SMALL_MAX && nb <= 3072)) {      /* 3072 bin counters + the rectangles fit the 48 KB of static + dynamic shared memory */
        /* small batch: the (triangle, bin) overlap count is bounded by n * nb, no read-back needed */
        if ((rc = grow(&LN.d_bin_list, &LN.cap_bin_list, (size_t)n * nb))) return rc;
        CK(launch_dep(k_front_small, dim3(1), dim3(1024), nb * sizeof(unsigned), LN.stream, d_tris, d_states, n, d_n, (int)s->w, (int)s->h, LN.d_bbox, LN.d_setup, LN.d_data,
                      setup_counters(), binsX, binsY, bshift, bshy, LN.d_bin_start, LN.d_bin_list));
        RT.launches += 1;
    } else {
        if (!rows_path) {
            /* Per-bin lists hold (triangle, bin) overlaps; their exact total is only known on the device (starts[nb]) and
               n * nb merely bounds it.  The host does NOT wait for it: the list is sized from the bound when that is small,
               else generously (4 entries per triangle, or what earlier batches of this lane needed, scaled), and the kernels
               compare the real total with the capacity themselves - on overflow k_bin_fill writes nothing and the
               rasterisers filter the whole batch against their tile instead of reading a list (slow, correct, and it
               happens once: the total is copied back behind the batch and raises the capacity of the next ones). */
            if (LN.list_pending && cudaEventQuery(LN.list_evt) == cudaSuccess) {
                LN.list_pending = false;
                if ((size_t)*LN.h_list_total > LN.list_hint) { LN.list_hint = *LN.h_list_total; }
            }
            const size_t bound = (size_t)n * (size_t)nb;
            size_t want = (size_t)n * 4 > 65536 ? (size_t)n * 4 : 65536;
            if (LN.list_hint_n) {
                const size_t scaled = (size_t)((double)LN.list_hint * 1.25 * ((double)n / (double)LN.list_hint_n > 1.0 ? (double)n / (double)LN.list_hint_n : 1.0)) + 1024;
                if (scaled > want) want = scaled;
            }
            if (want > bound) want = bound;
            if (want < LN.cap_bin_list) want = LN.cap_bin_list < bound ? LN.cap_bin_list : bound;
            if ((rc = grow(&LN.d_bin_list, &LN.cap_bin_list, want ? want : 1))) return rc;
            list_cap = LN.cap_bin_list;
            list_total_wanted = true;
        }
        /* the kernels of the batch back to back, each a programmatic dependent of the one before */
        if (!RT.setup_attr_set) { cudaFuncSetAttribute(k_setup, cudaFuncAttributeMaxDynamicSharedMemorySize, SETUP_SMEM_BYTES); RT.setup_attr_set = true; }
        const unsigned setup_chunks = (n + SETUP_THREADS - 1) / SETUP_THREADS, setup_ctas = setup_chunks < (unsigned)RT.sms * 3u ? setup_chunks : (unsigned)RT.sms * 3u;
        CK(launch_dep(k_setup, dim3(setup_ctas), dim3(SETUP_THREADS), (size_t)SETUP_SMEM_BYTES, LN.stream,
                      d_tris, d_states, n, (int)s->w, (int)s->h, LN.d_bbox, LN.d_setup, LN.d_data, setup_counters()));
        RT.launches += 1;
        if (!rows_path) {
            unsigned *d_totals = LN.d_bin_start + (MAX_BINS + 2), *d_ticket = LN.d_bin_start + (MAX_BINS + 2) * 2;
            CK(launch_dep(k_bin_count, dim3(nBatches), dim3(256), nb * sizeof(unsigned), LN.stream, (const int4 *)LN.d_bbox, n, bin_batch, binsX, binsY, bshift, bshy, LN.d_bin_counts));
            CK(launch_dep(k_bin_scan, dim3((nb + 31) / 32), dim3(1024), 0, LN.stream, LN.d_bin_counts, (int)nBatches, nb, d_totals, LN.d_bin_start, d_ticket));
            CK(launch_dep(k_bin_fill, dim3(nBatches), dim3(256), nb * sizeof(unsigned), LN.stream, (const int4 *)LN.d_bbox, n, bin_batch, binsX, binsY, bshift, bshy,
                          (const unsigned *)LN.d_bin_counts, (const unsigned *)LN.d_bin_start, LN.d_bin_list, list_cap > 0xffffffffu ? 0xffffffffu : (unsigned)list_cap));
            RT.launches += 3;
        }
    }
    if (rows_path) {
        RowsParams rp;
        rp.bbox = LN.d_bbox; rp.setup = LN.d_setup; rp.data = LN.d_data; rp.states = d_states; rp.n = n; rp.d_n = d_n;
        rp.color = s->color; rp.depth = s->depth; rp.W = (int)s->w; rp.H = (int)s->h; rp.fb_fmt = s->fmt; rp.counters = raster_counters(s);
        if (RT.profiling) CK(cudaEventRecord(pe[1], LN.stream));
        const unsigned bands = (s->h + ROWS_NW - 1) / ROWS_NW;
        if (feature_mask & PFCU_ST_PHONG) k_raster_rows<true><<<bands, ROWS_NW * 32, 0, LN.stream>>>(rp);
        else                              k_raster_rows<false><<<bands, ROWS_NW * 32, 0, LN.stream>>>(rp);
        RT.launches++;
        if (RT.profiling) { CK(cudaEventRecord(pe[2], LN.stream)); for (int i = 0; i < 3; i++) RT.prof_events.push_back(pe[i]); }
        CK(cudaGetLastError());
        mark_done(s);
        for (pfcu_surface *dep : RT.deps)
            if (dep != s && dep->lane != s->lane) CK(cudaStreamWaitEvent(RT.lanes[dep->lane % RT.n_lanes].stream, s->done, 0));
        RT.deps.clear();
        if (!d_n) RT.submitted += n;
        return PFCU_OK;
    }
    RasterParams p;
    p.bbox = LN.d_bbox; p.setup = LN.d_setup; p.data = LN.d_data; p.states = d_states;
    p.bin_list = LN.d_bin_list; p.bin_starts = LN.d_bin_start; p.binsX = binsX; p.bsx = bshift; p.bsy = bshy;
    p.nb = nb; p.list_cap = list_cap > 0xffffffffu ? 0xffffffffu : (unsigned)list_cap; p.n = d_n ? 0u : n;
    p.color = s->color; p.depth = s->depth; p.W = (int)s->w; p.H = (int)s->h;
    p.tilesX = (int)s->tiles_x; p.tilesY = (int)s->tiles_y;
    p.rank = s->rank; p.world = s->world ? s->world : 1; p.nTiles = s->tiles_x * s->tiles_y;
    p.counters = raster_counters(s);
    const unsigned grid = owned_tiles(s, p.rank, p.world);
    if (RT.profiling) CK(cudaEventRecord(pe[1], LN.stream));
    p.tile_base = 0;
    /* the rasteriser over `grid` tiles starting at p.tile_base, on stream st_ */
    const unsigned grid_all = grid;
    auto launch_raster = [&](unsigned grid, cudaStream_t st_) -> int {
        /* many small triangles per tile: 16 warps per tile halve the serial work of the busiest tiles;
           few large ones: 8 warps with more registers each issue faster */
        const bool ph = (feature_mask & PFCU_ST_PHONG) != 0;
        if (RT.rcp_bits > RCP_SMEM_BITS) single_prog = -1;      /* the fixed-program kernels assume the shared RCPPS table */
        /* half-height slices when the 64x64 grid would be only a few waves deep with a ragged last wave */
        const bool fixed = single_prog >= 0 && single_prog < PROG_PHONG && !small_tris && !use_frag;
        const int per_sm = (fixed && single_prog / 4 != 4) ? 4 : 3;
        const double waves = (double)grid_all / ((double)RT.sms * per_sm);      /* bands run side by side: the whole surface counts */
        /* slices per tile: 64x64, 64x32 or 64x16 - whichever leaves the least of the last wave empty (a tile-split surface
           on 8 GPUs: 1020 tiles = 1.7 waves of 64x64 CTAs, 6.9 waves of 64x16 slices) */
        int slices = 1;
        if (!small_tris && !ph && waves < 8.0 && (ceil(waves) / waves) > 1.06 && (ceil(2 * waves) / (2 * waves)) < (ceil(waves) / waves)) slices = 2;
        /* 64x16 slices were tried for the 1.7-wave grids of an 8-GPU split: the per-slice cost (every warp pays each
           triangle's prologue for four block rows instead of sixteen) is +17 % against +4 % for 64x32, more than the
           emptier last wave costs (measured on one rank's share of the 8K scene: 1.274 / 1.250 / 1.351 ms) */
        static const int env_slices = getenv("PF_CUDA_SLICES") ? atoi(getenv("PF_CUDA_SLICES")) : 0;       /* 1 or 2: experiments */
        if (env_slices == 1 || env_slices == 2) slices = (small_tris || ph) ? 1 : env_slices;
        const bool half = slices == 2;
        if (use_frag) {
            /* 64x8 slices of eight 8x8 regions, 8 warps; 4 CTAs per SM (Phong: 3, 80 registers) */
            if (!RT.frag_attr_set) {            /* function attributes are per device: once per runtime, not per process */
                cudaFuncSetAttribute(k_raster_frag<true, 8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * FRAG_NF_PHONG * 512);
                cudaFuncSetAttribute(k_raster_frag<false, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * FRAG_NF * 512);
                RT.frag_attr_set = true;
            }
            if (ph) CK(launch_dep(k_raster_frag<true, 8, 3>, dim3(grid * 8), dim3(256), (size_t)(8 * FRAG_NF_PHONG * 512), st_, p));
            else    CK(launch_dep(k_raster_frag<false, 8, 4>, dim3(grid * 8), dim3(256), (size_t)(8 * FRAG_NF * 512), st_, p));
        }
        else if (small_tris) {
            const int th = force_slice ? force_slice : (ph ? 32 : 16);
            if (ph) { if (th <= 32) CK(launch_dep(k_raster<true, 16, -1, 32>, dim3(grid * 2), dim3(512), (size_t)(0), st_, p)); else CK(launch_dep(k_raster<true, 16, -1, 64>, dim3(grid), dim3(512), (size_t)(0), st_, p)); }
            else if (th <= 16) CK(launch_dep(k_raster<false, 16, -1, 16>, dim3(grid * 4), dim3(512), (size_t)(0), st_, p));
            else if (th <= 32) CK(launch_dep(k_raster<false, 16, -1, 32>, dim3(grid * 2), dim3(512), (size_t)(0), st_, p));
            else               CK(launch_dep(k_raster<false, 16, -1, 64>, dim3(grid), dim3(512), (size_t)(0), st_, p));
        }
        else if (ph)          CK(launch_dep(k_raster<true, 8, -1, 64>, dim3(grid), dim3(256), (size_t)(0), st_, p));
        else if (fixed) {
            /* one state program in the whole batch: the kernel that holds only that fragment program */
            switch (single_prog) {
#define FIXED_CASE(P) case P: CK(launch_fixed<P>(slices, grid, st_, p)); break;
            FIXED_CASE(0) FIXED_CASE(1) FIXED_CASE(2) FIXED_CASE(3) FIXED_CASE(4) FIXED_CASE(5) FIXED_CASE(6) FIXED_CASE(7)
            FIXED_CASE(12) FIXED_CASE(13) FIXED_CASE(14) FIXED_CASE(15) FIXED_CASE(16) FIXED_CASE(17) FIXED_CASE(18) FIXED_CASE(19)
#undef FIXED_CASE
            default: snprintf(RT.err, sizeof RT.err, "internal: no kernel for state program %d", single_prog); return PFCU_ERR_INVALID;
            }
        }
        else { if (half) CK(launch_dep(k_raster<false, 8, -1, 32>, dim3(grid * 2), dim3(256), (size_t)(0), st_, p)); else CK(launch_dep(k_raster<false, 8, -1, 64>, dim3(grid), dim3(256), (size_t)(0), st_, p)); }
        RT.launches++;
        return PFCU_OK;
    };
    bool banded = false;
    /* Large RGBA8 surfaces: a few launches over bands of tile rows, each on its own stream (they overlap on the GPU like
       one launch), with an event per band - a read-back that follows (pfcu_surface_download_async) copies band b as
       soon as band b is done, while the later bands are still being rasterised. */
    static const int env_bands = getenv("PF_CUDA_BANDS") ? atoi(getenv("PF_CUDA_BANDS")) : -1;
    int n_bands = 1;
    /* from 1 Mpixel (a band costs four more runtime calls on the launch path; measured on the 1080p scene: 2 bands gain
       nothing end to end, 4 bands 0.02 ms; 4K scenes 0.3 - 0.5 ms, 8K 1.9 ms) */
    s->n_since_read++;
    const bool read_back_expected = s->n_per_read != 0 && s->n_since_read == s->n_per_read;
    /* one GPU: six bands of unequal height (profiles/r02_e2e_timeline.txt).  Where the read-back is the longer leg, the frame
       ends one transfer of the whole surface after the FIRST band is rasterised: heights grow from the first band on
       (weights 1, g, g^2, ...; g below the copy / raster time ratio, so the copy engine is never left waiting).  Where
       rasterisation is the longer leg, the frame ends one transfer of the LAST band after it is rasterised: small at both
       ends (1, g, g^2, g^2, g, 1).  Which leg is longer is taken from the surface's previous banded frame: its
       rasterisation is timed with two events, the transfer is its bytes at PCIE_GBS.  Tile split (several devices): four
       equal bands, as measured there. */
    const bool grow_bands = p.world <= 1;
    if (grow_bands && s->t_pending) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, s->t_front, s->t_raster) == cudaSuccess) {
            s->raster_longer = (double)ms * 1e-3 > (double)s->w * s->h * 4.0 / (PCIE_GBS * 1e9);
            s->t_pending = false;
        } else cudaGetLastError();              /* not finished yet: keep the last answer */
    }
    static const double band_growth = getenv("PF_CUDA_BAND_GROWTH") ? atof(getenv("PF_CUDA_BAND_GROWTH")) : 0.0;
    if (grid && s->fmt == PFCU_TEX_RGBA8 && (size_t)s->w * s->h >= ((size_t)1 << 20) && s->tiles_y >= 8 && read_back_expected) n_bands = grow_bands ? MAX_BANDS : 4;
    if (env_bands >= 1) n_bands = env_bands > MAX_BANDS ? MAX_BANDS : env_bands;
    if (!grid || (unsigned)n_bands > s->tiles_y) n_bands = 1;
    /* tile split: this rank's tiles t = rank + k * world, k = 0 .. grid-1; the ones in tile rows < R are k < first_owned(R) */
    auto first_owned = [&](unsigned R) -> unsigned {
        const unsigned t = R * s->tiles_x;
        if (p.world <= 1) return t;
        return t <= p.rank ? 0u : (t - p.rank + p.world - 1u) / p.world;
    };
    if (n_bands > 1) {
        for (int b = 0; b < n_bands; b++) if (!s->band_evt[b]) CK(cudaEventCreateWithFlags(&s->band_evt[b], cudaEventDisableTiming));
        CK(cudaEventRecord(RT.front_evt, LN.stream));
        const bool time_it = grow_bands && !s->t_pending;
        if (time_it) { if (!s->t_front) { CK(cudaEventCreate(&s->t_front)); CK(cudaEventCreate(&s->t_raster)); } CK(cudaEventRecord(s->t_front, LN.stream)); }
        if (g_timing) { tl_event(&g_tl.front, LN.stream); g_tl.n = n_bands; g_tl.armed = true; g_tl.copies = false; }
        unsigned row0 = 0;
        for (int b = 0; b < n_bands; b++) {
            /* the first bands are the smaller ones: their copies start early, the last band's copy is what remains exposed */
            unsigned row1 = (b == n_bands - 1) ? s->tiles_y : (unsigned)(((uint64_t)s->tiles_y * (unsigned)(b + 1)) / (unsigned)n_bands);
            const double growth = band_growth > 0.0 ? band_growth : (s->raster_longer ? 1.5 : 1.35);
            if (grow_bands && growth > 1.0 && b != n_bands - 1) {
                double cum = 0.0, all = 0.0, w = 1.0;
                for (int k = 0; k < n_bands; k++) {
                    const int e = (s->raster_longer && n_bands - 1 - k < k) ? n_bands - 1 - k : k;     /* from the first band, or from the nearer end */
                    w = 1.0; for (int j = 0; j < e; j++) w *= growth;
                    if (k <= b) cum += w;
                    all += w;
                }
                row1 = (unsigned)((double)s->tiles_y * cum / all + 0.5);
                if (row1 <= row0) row1 = row0 + 1;
                const unsigned most = s->tiles_y - (unsigned)(n_bands - 1 - b);      /* every later band keeps a row */
                if (row1 > most) row1 = most;
            }
            cudaStream_t bs = RT.band_streams[b];
            CK(cudaStreamWaitEvent(bs, RT.front_evt, 0));
            const unsigned o0 = first_owned(row0), o1 = b == n_bands - 1 ? grid : first_owned(row1);
            p.tile_base = o0;
            if (o1 > o0 && (rc = launch_raster(o1 - o0, bs))) return rc;
            CK(cudaEventRecord(s->band_evt[b], bs));
            if (g_timing) tl_event(&g_tl.band[b], bs);
            CK(cudaStreamWaitEvent(LN.stream, s->band_evt[b], 0));
            s->band_y[b] = row0 * TILE; s->band_owned[b] = o0;
            row0 = row1;
        }
        if (time_it) { CK(cudaEventRecord(s->t_raster, LN.stream)); s->t_pending = true; }      /* the lane has waited for every band */
        s->band_y[n_bands] = s->h; s->band_owned[n_bands] = grid;
        s->n_bands = n_bands; banded = true;
    } else if (grid) {
        if ((rc = launch_raster(grid, LN.stream))) return rc;
    }
    if (RT.profiling) { CK(cudaEventRecord(pe[2], LN.stream)); for (int i = 0; i < 3; i++) RT.prof_events.push_back(pe[i]); }
    CK(cudaGetLastError());
    if (list_total_wanted) {            /* behind the batch, so that its kernels stay adjacent in the stream */
        CK(cudaMemcpyAsync(LN.h_list_total, LN.d_bin_start + nb, sizeof(unsigned), cudaMemcpyDeviceToHost, LN.stream));
        CK(cudaEventRecord(LN.list_evt, LN.stream));
        LN.list_pending = true; LN.list_hint_n = n > LN.list_hint_n ? n : LN.list_hint_n;
    }
    mark_done(s);
    s->bands_valid = banded;
    /* write-after-read: a sampled surface on another lane must not be overwritten before this batch read it */
    for (pfcu_surface *dep : RT.deps)
        if (dep != s && dep->lane != s->lane) CK(cudaStreamWaitEvent(RT.lanes[dep->lane % RT.n_lanes].stream, s->done, 0));
    RT.deps.clear();
    if (!d_n) RT.submitted += n;              /* otherwise counted on the device (counters[3]) */
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
    RT.launches++;
    if (nb > 1) {
        unsigned *d_tmp2 = d_tmp + nb;
        int rc = scan_exclusive(d_tmp, d_tmp, nb, d_tmp2, st);
        if (rc) return rc;
        k_scan_add<<<(n + 255u) / 256u, 256, 0, st>>>(d_out, n, d_tmp);
        RT.launches++;
    }
    CK(cudaGetLastError());
    return PFCU_OK;
}


/* device-resident lists run many small surfaces per launch on ONE device; the multi-device mode replays lists through
   pfcu_submit_raw instead, which every device executes */
unsigned pfcu_capabilities(void) { return PFCU_CAP_DEVICE_VERTEX | PFCU_CAP_RAW_TRIANGLES | (mg.n > 1 ? 0u : PFCU_CAP_LISTS); }

int pfcu_draw_triangles(pfcu_surface *s, const pfcu_state *state, const pfcu_vparams *vp, const pfcu_draw *d, uint32_t *n_out)
{
    API_LOCK;
    if (n_out) *n_out = 0;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || !state || !vp || !d || !d->positions || d->pos_size < 2 || d->pos_size > 4 || d->n_faces < 1 || d->n_faces > 2) return PFCU_ERR_INVALID;
    const unsigned n_tri = d->count / 3u;
    if (n_tri == 0) return PFCU_OK;
    if (MULTI_SURF(s)) return multi_run_all([&](int k) -> int {
        const std::vector<pfcu_state> st = states_for_device(state, 1, k);
        return pfcu_draw_triangles(s->rep[k], st.data(), vp, d, k == 0 ? n_out : nullptr); });
    use_lane(s);
    const unsigned n_items = n_tri * d->n_faces;
    int rc;
    const double t_start = g_timing ? now_us() : 0.0; double t_up = 0, t_cnt = 0;
    /* arrays -> device (pageable sources are staged by the driver; ordered on the stream); arrays in a static block
       (pfcu_host_set_static) are read from the block's device mirror instead and cross PCIe only when modified */
    size_t nv = d->n_vertices;
    const unsigned char *d_indices_ready = nullptr;
    if (nv == 0 && d->indices) {
        /* n_vertices unknown: the device finds the largest (32-bit) index */
        if (d->index_bytes != 4) return PFCU_ERR_INVALID;
        const size_t bi = (size_t)d->count * 4;
        PinnedBlock *ib = nullptr;
        const unsigned char *mir = static_mirror(d->indices, bi, &ib);
        if (mir && ib->imax_ptr == d->indices && ib->imax_count == d->count && ib->imax_gen == ib->gen) {
            nv = (size_t)ib->imax_value + 1;        /* scanned before, block unchanged since */
            d_indices_ready = mir;
        } else {
            if (mir) d_indices_ready = mir;
            else {
                if ((rc = grow(&LN.d_idx, &LN.cap_idx, bi))) return rc;
                CK(cudaMemcpyAsync(LN.d_idx, d->indices, bi, cudaMemcpyHostToDevice, LN.stream));
                RT.bytes_h2d += bi;
                d_indices_ready = LN.d_idx;
            }
            CK(cudaMemsetAsync(LN.d_total, 0, 4, LN.stream));
            k_index_max<<<RT.sms * 4, 256, 0, LN.stream>>>((const unsigned *)d_indices_ready, d->count, LN.d_total);
            RT.launches++;
            CK(cudaMemcpyAsync(&LN.h_total[0], LN.d_total, 4, cudaMemcpyDeviceToHost, LN.stream));
            CK(cudaStreamSynchronize(LN.stream));
            nv = (size_t)LN.h_total[0] + 1;
            if (mir) { ib->imax_ptr = d->indices; ib->imax_count = d->count; ib->imax_value = LN.h_total[0]; ib->imax_gen = ib->gen; }
        }
        if (nv > 0x7fffffffu) return PFCU_ERR_INVALID;
    }
    const size_t b_pos = nv * d->pos_size * 4, b_nrm = d->normals ? nv * 12 : 0, b_uv = d->texcoords ? nv * 8 : 0;
    const size_t b_col = d->colors ? nv * d->color_size : 0, b_idx = (d->indices && !d_indices_ready) ? (size_t)d->count * d->index_bytes : 0;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const unsigned char *m_pos = static_mirror(d->positions, b_pos, nullptr);
    const unsigned char *m_nrm = b_nrm ? static_mirror(d->normals, b_nrm, nullptr) : nullptr;
    const unsigned char *m_uv = b_uv ? static_mirror(d->texcoords, b_uv, nullptr) : nullptr;
    const unsigned char *m_col = b_col ? static_mirror(d->colors, b_col, nullptr) : nullptr;
    const unsigned char *m_idx = b_idx ? static_mirror(d->indices, b_idx, nullptr) : nullptr;
    const size_t total_bytes = (m_pos ? 0 : al(b_pos)) + (m_nrm ? 0 : al(b_nrm)) + (m_uv ? 0 : al(b_uv)) + (m_col ? 0 : al(b_col)) + (m_idx ? 0 : al(b_idx));
    RT.bytes_h2d += (m_pos ? 0 : b_pos) + (m_nrm ? 0 : b_nrm) + (m_uv ? 0 : b_uv) + (m_col ? 0 : b_col) + (m_idx ? 0 : b_idx) + sizeof(DevState);
    if ((rc = grow(&LN.d_varrays, &LN.cap_varrays, total_bytes ? total_bytes : 256))) return rc;
    unsigned char *p = LN.d_varrays;
    VtxArgs a; memset(&a, 0, sizeof a);
    if (m_pos) a.pos = (const float *)m_pos; else { a.pos = (const float *)p; CK(cudaMemcpyAsync(p, d->positions, b_pos, cudaMemcpyHostToDevice, LN.stream)); p += al(b_pos); }
    if (b_nrm) { if (m_nrm) a.nrm = (const float *)m_nrm; else { a.nrm = (const float *)p; CK(cudaMemcpyAsync(p, d->normals, b_nrm, cudaMemcpyHostToDevice, LN.stream)); p += al(b_nrm); } }
    if (b_uv) { if (m_uv) a.uv = (const float *)m_uv; else { a.uv = (const float *)p; CK(cudaMemcpyAsync(p, d->texcoords, b_uv, cudaMemcpyHostToDevice, LN.stream)); p += al(b_uv); } }
    if (b_col) { if (m_col) a.col = m_col; else { a.col = p; CK(cudaMemcpyAsync(p, d->colors, b_col, cudaMemcpyHostToDevice, LN.stream)); p += al(b_col); } }
    if (b_idx) { if (m_idx) a.idx = m_idx; else { a.idx = p; CK(cudaMemcpyAsync(p, d->indices, b_idx, cudaMemcpyHostToDevice, LN.stream)); p += al(b_idx); } }
    if (d_indices_ready) a.idx = d_indices_ready;
    a.pos_size = (int)d->pos_size; a.col_size = (int)d->color_size; a.idx_bytes = (int)d->index_bytes;
    a.first = d->first; a.n_tri = n_tri; a.cur_color = d->current_color; a.n_faces = (int)d->n_faces;
    a.face[0] = d->faces[0]; a.face[1] = d->faces[1]; a.state = 0;

    if (g_timing) t_up = now_us();
    /* pass A: output triangles per (triangle, face) item; scan; total */
    if ((rc = grow(&LN.d_vcounts, &LN.cap_vcounts, (size_t)n_items * 2 + n_items / 512 + 64))) return rc;
    unsigned *d_counts = LN.d_vcounts, *d_offsets = LN.d_vcounts + n_items, *d_tmp = LN.d_vcounts + 2 * (size_t)n_items;
    k_vertex_count<<<(n_items + 127u) / 128u, 128, 0, LN.stream>>>(a, *vp, n_items, d_counts);
    RT.launches++;
    if ((rc = scan_exclusive(d_counts, d_offsets, n_items, d_tmp))) return rc;
    k_scan_total<<<1, 1, 0, LN.stream>>>(d_offsets + (n_items - 1), d_counts + (n_items - 1), LN.h_total);
    RT.launches++;
    CK(cudaStreamSynchronize(LN.stream));
    const unsigned total = LN.h_total[0] + LN.h_total[1];
    if (g_timing) t_cnt = now_us();
    if (n_out) *n_out = total;
    if (total == 0) return PFCU_OK;

    /* pass B: emit in order, then the usual setup -> bin -> raster pipeline */
    if ((rc = grow(&LN.d_tris, &LN.cap_tris, total))) return rc;
    if ((rc = grow(&LN.d_states, &LN.cap_states, 1))) return rc;
    DevState hs;
    const unsigned mask = convert_states(state, 1, &hs);
    CK(cudaMemcpyAsync(LN.d_states, &hs, sizeof hs, cudaMemcpyHostToDevice, LN.stream));
    k_vertex_emit<<<(n_items + 127u) / 128u, 128, 0, LN.stream>>>(a, *vp, n_items, d_offsets, LN.d_tris);
    RT.launches++;
    CK(cudaEventRecord(LN.raw_done, LN.stream));        /* d_vcounts is shared with the raw-triangle path's side stream */
    CK(cudaGetLastError());
    rc = launch_pipeline(s, LN.d_tris, LN.d_states, total, mask, g_last_single_prog);
    if (g_timing) fprintf(stderr, "pfcu_draw_triangles: arrays %.1f us, count pass + wait %.1f us, emit + pipeline launches %.1f us\n", t_up - t_start, t_cnt - t_up, now_us() - t_cnt);
    return rc;
}

int pfcu_submit_raw(pfcu_surface *s, const pfcu_state *states, uint32_t n_states, const pfcu_vparams_lit *vparams, uint32_t n_vparams,
                    const float *pow_tables, uint32_t n_pow_tables, const pfcu_rawtri *tris, uint32_t n_tris, uint32_t *n_out)
{
    API_LOCK;
    if (n_out) *n_out = 0;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || (n_tris && (!states || !tris || !vparams || n_states == 0 || n_vparams == 0))) return PFCU_ERR_INVALID;
    if (n_tris == 0) return PFCU_OK;
    if (MULTI_SURF(s)) return multi_run_all([&](int k) -> int {
        const std::vector<pfcu_state> st = states_for_device(states, n_states, k);
        return pfcu_submit_raw(s->rep[k], st.data(), n_states, vparams, n_vparams, pow_tables, n_pow_tables, tris, n_tris, k == 0 ? n_out : nullptr); });
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
        if (PinnedBlock *pb = find_pinned(tris)) { if ((rc = pinned_in_flight(pb, LN.stream))) return rc; }
        q += al(b_tris);
        ra.vp = (const pfcu_vparams_lit *)q; CK(cudaMemcpyAsync(q, vparams, b_vp, cudaMemcpyHostToDevice, LN.stream)); q += al(b_vp);
        ra.pow_tables = (const float *)q; if (b_pow) CK(cudaMemcpyAsync(q, pow_tables, b_pow, cudaMemcpyHostToDevice, LN.stream));
        RT.bytes_h2d += b_tris + b_vp + b_pow + n_states * sizeof(DevState);
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
        if (++LN.chain_seq == 0) ++LN.chain_seq;         /* 0 is what the zero-initialised flags hold */
        CK(launch_dep(k_raw_chain, dim3((n_tris + 127u) / 128u), dim3(128), 0, LN.stream, ra, LN.d_tris, d_total, LN.d_chain, LN.chain_seq));
        RT.launches++;
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
    if (PinnedBlock *pb = find_pinned(tris)) { if ((rc = pinned_in_flight(pb, LN.vstream))) return rc; }
    p += al(b_tris);
    a.vp = (const pfcu_vparams_lit *)p; CK(cudaMemcpyAsync(p, vparams, b_vp, cudaMemcpyHostToDevice, LN.vstream)); p += al(b_vp);
    a.pow_tables = (const float *)p; if (b_pow) CK(cudaMemcpyAsync(p, pow_tables, b_pow, cudaMemcpyHostToDevice, LN.vstream));
    RT.bytes_h2d += b_tris + b_vp + b_pow + n_states * sizeof(DevState);

    unsigned *d_counts = LN.d_vcounts, *d_offsets = LN.d_vcounts + n_tris, *d_tmp = LN.d_vcounts + 2 * (size_t)n_tris;
    k_raw_count<<<(n_tris + 127u) / 128u, 128, 0, LN.vstream>>>(a, d_counts);
    RT.launches++;
    if ((rc = scan_exclusive(d_counts, d_offsets, n_tris, d_tmp, LN.vstream))) return rc;
    k_scan_total<<<1, 1, 0, LN.vstream>>>(d_offsets + (n_tris - 1), d_counts + (n_tris - 1), LN.h_total);
    RT.launches++;
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
    RT.launches++;
    CK(cudaEventRecord(LN.raw_done, LN.stream));
    CK(cudaGetLastError());
    return launch_pipeline(s, LN.d_tris, LN.d_states, total, mask, g_last_single_prog);
}

/* ---- device-resident render lists ---- */

pfcu_list *pfcu_list_create(const pfcu_rawtri *tris, uint32_t n_tris)
{
    API_LOCK;
    if (!RT.ok || !tris || n_tris == 0) return nullptr;
    pfcu_list *l = (pfcu_list *)calloc(1, sizeof *l);
    if (!l) return nullptr;
    l->n = n_tris;
    RT.cur = &RT.lanes[0];
    if (cudaMalloc(&l->tris, (size_t)n_tris * sizeof(pfcu_rawtri)) != cudaSuccess) { snprintf(RT.err, sizeof RT.err, "list_create: out of device memory"); free(l); return nullptr; }
    if (cudaMemcpyAsync(l->tris, tris, (size_t)n_tris * sizeof(pfcu_rawtri), cudaMemcpyHostToDevice, LN.stream) != cudaSuccess ||
        cudaStreamSynchronize(LN.stream) != cudaSuccess) { cudaFree(l->tris); free(l); return nullptr; }
    RT.bytes_h2d += (size_t)n_tris * sizeof(pfcu_rawtri);
    return l;
}

void pfcu_list_destroy(pfcu_list *l)
{
    API_LOCK;
    if (!l) return;
    if (RT.ok) sync_all_lanes();
    cudaFree(l->tris); free(l);
}

uint32_t pfcu_list_size(const pfcu_list *l) { return l ? l->n : 0u; }

int pfcu_list_job_supported(const pfcu_surface *s, uint32_t n_tris, uint32_t n_segments)
{
    if (!RT.ok || mg.n > 1 || !s || n_tris == 0 || n_tris > PFCU_LIST_JOB_MAX_TRIS || n_segments > PFCU_LIST_JOB_MAX_SEGMENTS) return 0;
    if (s->fmt != PFCU_TEX_RGBA8 || s->world > 1) return 0;
    return sync_free_bshift(s, n_tris) != 0;
}

int pfcu_submit_list_jobs(const pfcu_list_job *jobs, uint32_t n_jobs)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (n_jobs == 0) return PFCU_OK;
    if (!jobs) return PFCU_ERR_INVALID;
    if (!RT.d_rcp) { snprintf(RT.err, sizeof RT.err, "pfcu_set_approx_tables() has not been called"); return PFCU_ERR_INVALID; }
    Lane &L0 = RT.lanes[0];
    RT.cur = &L0;
    if (L0.need_fence_wait) L0.need_fence_wait = false;
    L0.touched = true;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    /* ---- sizes, validation ---- */
    size_t bytes = al(sizeof(DevJob) * (size_t)n_jobs);
    unsigned feature = 0, max_slices = 0; size_t max_nb = 0;
    std::vector<unsigned> n_raw(n_jobs);
    for (uint32_t j = 0; j < n_jobs; j++) {
        const pfcu_list_job &J = jobs[j];
        if (!J.surface || !J.states || !J.vparams || !J.calls || (J.n_segments && !J.segments) || J.n_states == 0 || J.n_vparams == 0) return PFCU_ERR_INVALID;
        unsigned n = 0;
        for (uint32_t k = 0; k < J.n_segments; k++) { if (!J.segments[k].list) return PFCU_ERR_INVALID; n += J.segments[k].list->n; }
        n_raw[j] = n;
        if (n && !pfcu_list_job_supported(J.surface, n, J.n_segments)) { snprintf(RT.err, sizeof RT.err, "list job %u is outside the limits of pfcu_submit_list_jobs", j); return PFCU_ERR_INVALID; }
        if (!n && (J.surface->fmt != PFCU_TEX_RGBA8)) return PFCU_ERR_INVALID;
        bytes += al(sizeof(DevState) * J.n_states) + al(sizeof(pfcu_vparams_lit) * J.n_vparams) + al(sizeof(pfcu_list_call) * J.n_calls)
               + al(sizeof(float) * PFCU_POW_TABLE_SIZE * J.n_pow_tables);
    }
    /* ---- slots and the packed upload ---- */
    if (!RT.jobs_copied) { CK(cudaEventCreateWithFlags(&RT.jobs_copied, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&RT.jobs_done, cudaEventDisableTiming)); }
    if (bytes > RT.cap_jobs) {
        CK(cudaStreamSynchronize(L0.stream));
        cudaFree(RT.d_jobs); if (RT.h_jobs) cudaFreeHost(RT.h_jobs);
        RT.d_jobs = nullptr; RT.h_jobs = nullptr; RT.cap_jobs = 0;
        size_t c = (size_t)1 << 16; while (c < bytes) c *= 2;
        CK(cudaMalloc(&RT.d_jobs, c)); CK(cudaHostAlloc(&RT.h_jobs, c, cudaHostAllocDefault));
        RT.cap_jobs = c;
    } else CK(cudaEventSynchronize(RT.jobs_copied));          /* the previous submission has left the staging buffer */
    if (RT.job_slots.size() < n_jobs) RT.job_slots.resize(n_jobs);
    DevJob *hj = reinterpret_cast<DevJob *>(RT.h_jobs);
    size_t off = al(sizeof(DevJob) * (size_t)n_jobs);
    for (uint32_t j = 0; j < n_jobs; j++) {
        const pfcu_list_job &J = jobs[j];
        pfcu_surface *s = J.surface;
        DevJob &D = hj[j];
        memset(&D, 0, sizeof D);
        Runtime::JobSlot &S = RT.job_slots[j];
        const unsigned n = n_raw[j], bound = n * FRONT_SMALL_CHUNKS;
        const int bshift = n ? sync_free_bshift(s, n) : BIN_SHIFT_COARSE;
        const int binsX = (int)((s->w + (1u << bshift) - 1) >> bshift), binsY = (int)((s->h + (1u << bshift) - 1) >> bshift), nb = binsX * binsY;
        if (bound > S.cap_tris || (size_t)bound * nb > S.cap_list || !S.bin_start) {
            CK(cudaStreamSynchronize(L0.stream));
            if (bound > S.cap_tris) {
                cudaFree(S.d_tris); cudaFree(S.bbox); cudaFree(S.setup); cudaFree(S.data); S.cap_tris = 0;
                size_t c = 1024; while (c < bound) c *= 2;
                if (cudaMalloc(&S.d_tris, c * sizeof(pfcu_triangle)) != cudaSuccess || cudaMalloc(&S.bbox, c * sizeof(int4)) != cudaSuccess ||
                    cudaMalloc(&S.setup, c * sizeof(TriSetup)) != cudaSuccess || cudaMalloc(&S.data, c * sizeof(TriData)) != cudaSuccess) {
                    snprintf(RT.err, sizeof RT.err, "out of device memory for list job scratch"); return PFCU_ERR_OOM;
                }
                S.cap_tris = c;
            }
            if ((size_t)bound * nb > S.cap_list) {
                cudaFree(S.bin_list); S.cap_list = 0;
                size_t c = 4096; while (c < (size_t)bound * nb) c *= 2;
                if (cudaMalloc(&S.bin_list, c * sizeof(uint2)) != cudaSuccess) { snprintf(RT.err, sizeof RT.err, "out of device memory for list job bins"); return PFCU_ERR_OOM; }
                S.cap_list = c;
            }
            if (!S.bin_start) {
                CK(cudaMalloc(&S.bin_start, (3072 + 2) * sizeof(unsigned)));
                CK(cudaMalloc(&S.d_total, 64));
                CK(cudaMalloc(&S.chain, 16 * sizeof(unsigned long long)));
                CK(cudaMemset(S.chain, 0, 16 * sizeof(unsigned long long)));
                CK(cudaMemset(S.d_total, 0, 64));
            }
        }
        /* tables */
        unsigned char *hb = RT.h_jobs, *db = RT.d_jobs;
        DevState *hs = reinterpret_cast<DevState *>(hb + off); D.states = reinterpret_cast<const DevState *>(db + off);
        const unsigned mask = convert_states(J.states, J.n_states, hs);
        off += al(sizeof(DevState) * J.n_states);
        if (g_last_leader_tex || g_last_pix_tex) {
            g_last_leader_tex = false; g_last_pix_tex = false;
            snprintf(RT.err, sizeof RT.err, "list jobs sample RGBA8 / RGB8 / BGR8 textures only"); return PFCU_ERR_INVALID;
        }
        for (pfcu_surface *dep : RT.deps) if (dep->has_done && dep->lane != 0) CK(cudaStreamWaitEvent(L0.stream, dep->done, 0));
        RT.deps.clear();
        feature |= mask;
        memcpy(hb + off, J.vparams, sizeof(pfcu_vparams_lit) * J.n_vparams); D.vp = reinterpret_cast<const pfcu_vparams_lit *>(db + off);
        off += al(sizeof(pfcu_vparams_lit) * J.n_vparams);
        memcpy(hb + off, J.calls, sizeof(pfcu_list_call) * J.n_calls); D.calls = reinterpret_cast<const pfcu_list_call *>(db + off);
        off += al(sizeof(pfcu_list_call) * J.n_calls);
        if (J.n_pow_tables) memcpy(hb + off, J.pow_tables, sizeof(float) * PFCU_POW_TABLE_SIZE * J.n_pow_tables);
        D.pow_tables = reinterpret_cast<const float *>(db + off);
        off += al(sizeof(float) * PFCU_POW_TABLE_SIZE * J.n_pow_tables);
        unsigned first = 0;
        for (uint32_t k = 0; k < J.n_segments; k++) {
            D.seg[k].tris = J.segments[k].list->tris; D.seg[k].first_tri = first; D.seg[k].first_call = J.segments[k].first_call;
            first += J.segments[k].list->n;
        }
        D.n_seg = J.n_segments; D.n_raw = n;
        D.color = s->color; D.depth = s->depth; D.W = s->w; D.H = s->h;
        D.clear = J.clear ? 1u : 0u; D.clear_rgba = J.clear_rgba; D.clear_z = J.clear_depth;
        D.d_tris = S.d_tris; D.d_total = S.d_total; D.chain = S.chain;
        D.binsX = binsX; D.binsY = binsY; D.bshift = bshift;
        RasterParams &p = D.rp;
        p.bbox = S.bbox; p.setup = S.setup; p.data = S.data; p.states = D.states;
        p.bin_list = S.bin_list; p.bin_starts = S.bin_start; p.binsX = binsX; p.bsx = bshift; p.bsy = bshift;
        p.color = s->color; p.depth = s->depth; p.W = (int)s->w; p.H = (int)s->h; p.tilesX = (int)s->tiles_x; p.tilesY = (int)s->tiles_y;
        p.rank = 0; p.world = 1; p.nTiles = n ? s->tiles_x * s->tiles_y : 0u;      /* a job without triangles rasterises nothing */
        p.counters = RT.d_counters; p.nb = nb; p.list_cap = 0xffffffffu; p.n = 0;
        if (n) { if (s->tiles_x * s->tiles_y * 8u > max_slices) max_slices = s->tiles_x * s->tiles_y * 8u; if ((size_t)nb > max_nb) max_nb = (size_t)nb; }
        RT.bytes_h2d += sizeof(DevState) * J.n_states + sizeof(pfcu_vparams_lit) * J.n_vparams + sizeof(pfcu_list_call) * J.n_calls + sizeof(DevJob);
    }
    /* ---- order against what is queued on the surfaces' own lanes, upload, four launches ---- */
    bool lane_used[MAX_LANES] = { false };
    for (uint32_t j = 0; j < n_jobs; j++) lane_used[jobs[j].surface->lane % RT.n_lanes] = true;
    for (int l = 1; l < RT.n_lanes; l++) if (lane_used[l]) {
        Lane &Ln = RT.lanes[l];
        if (Ln.need_fence_wait) { cudaStreamWaitEvent(Ln.stream, RT.lanes[0].fence, 0); Ln.need_fence_wait = false; }
        CK(cudaEventRecord(Ln.vready, Ln.stream));
        CK(cudaStreamWaitEvent(L0.stream, Ln.vready, 0));
    }
    CK(cudaMemcpyAsync(RT.d_jobs, RT.h_jobs, off, cudaMemcpyHostToDevice, L0.stream));
    CK(cudaEventRecord(RT.jobs_copied, L0.stream));
    const DevJob *dj = reinterpret_cast<const DevJob *>(RT.d_jobs);
    bool any_clear = false; for (uint32_t j = 0; j < n_jobs; j++) any_clear |= jobs[j].clear != 0;
    cudaEvent_t pe[3] = { nullptr, nullptr, nullptr };
    if (RT.profiling) {
        for (int i = 0; i < 3; i++) {
            if (!RT.prof_pool.empty()) { pe[i] = RT.prof_pool.back(); RT.prof_pool.pop_back(); }
            else CK(cudaEventCreate(&pe[i]));
        }
    }
    if (any_clear) { k_jobs_clear<<<dim3(32, n_jobs), 256, 0, L0.stream>>>(dj); RT.launches++; }
    if (RT.profiling) CK(cudaEventRecord(pe[0], L0.stream));
    if (max_slices) {
        if (++RT.jobs_seq == 0) ++RT.jobs_seq;
        CK(launch_dep(k_list_chain, dim3(PFCU_LIST_JOB_MAX_TRIS / 128u, n_jobs), dim3(128), 0, L0.stream, dj, RT.jobs_seq));
        CK(launch_dep(k_front_small_jobs, dim3(1, n_jobs), dim3(1024), max_nb * sizeof(unsigned), L0.stream, dj, RT.d_counters));
        static const bool attr_once = [] {
            cudaFuncSetAttribute(k_raster_frag_jobs<true, 8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * FRAG_NF_PHONG * 512);
            cudaFuncSetAttribute(k_raster_frag_jobs<false, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * FRAG_NF * 512);
            return true; }();
        (void)attr_once;
        if (RT.profiling) CK(cudaEventRecord(pe[1], L0.stream));
        if (feature & PFCU_ST_PHONG) CK(launch_dep(k_raster_frag_jobs<true, 8, 3>, dim3(max_slices, n_jobs), dim3(256), (size_t)(8 * FRAG_NF_PHONG * 512), L0.stream, dj));
        else                         CK(launch_dep(k_raster_frag_jobs<false, 8, 4>, dim3(max_slices, n_jobs), dim3(256), (size_t)(8 * FRAG_NF * 512), L0.stream, dj));
        RT.launches += 3;
    } else if (RT.profiling) CK(cudaEventRecord(pe[1], L0.stream));
    if (RT.profiling) { CK(cudaEventRecord(pe[2], L0.stream)); for (int i = 0; i < 3; i++) RT.prof_events.push_back(pe[i]); }
    CK(cudaGetLastError());
    CK(cudaEventRecord(RT.jobs_done, L0.stream));
    for (int l = 1; l < RT.n_lanes; l++) if (lane_used[l]) { CK(cudaStreamWaitEvent(RT.lanes[l].stream, RT.jobs_done, 0)); RT.lanes[l].touched = true; }
    for (uint32_t j = 0; j < n_jobs; j++) {
        pfcu_surface *s = jobs[j].surface;
        s->bands_valid = false;
        if (s->aliased) { if (cudaEventRecord(s->done, L0.stream) == cudaSuccess) s->has_done = true; }     /* someone may sample it from another lane */
        else s->has_done = false;          /* its lane already waits for this submission (above) */
    }
    return PFCU_OK;
}

int pfcu_submit_prims(pfcu_surface *s, const pfcu_prim *prims, uint32_t n)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || (n && !prims)) return PFCU_ERR_INVALID;
    if (n == 0) return PFCU_OK;
    if (MULTI_SURF(s)) return multi_run_all([&](int d) -> int { return pfcu_submit_prims(s->rep[d], prims, n); });
    use_lane(s);
    int rc;
    const size_t bytes = (size_t)n * sizeof(pfcu_prim);
    if ((rc = grow(&LN.d_varrays, &LN.cap_varrays, bytes))) return rc;
    CK(cudaMemcpyAsync(LN.d_varrays, prims, bytes, cudaMemcpyHostToDevice, LN.stream));
    RT.bytes_h2d += bytes;
    PrimParams p;
    p.prims = (const pfcu_prim *)LN.d_varrays; p.n = n; p.color = s->color; p.depth = s->depth; p.W = s->w; p.H = s->h;
    p.tilesX = (int)s->tiles_x; p.rank = s->rank; p.world = s->world ? s->world : 1; p.nTiles = s->tiles_x * s->tiles_y;
    p.alpha_or = s->fmt >= PFCU_TEX_RGB8 ? 0xff000000u : 0u;
    const unsigned grid = owned_tiles(s, p.rank, p.world);
    if (grid) { k_prims<<<grid, 256, 0, LN.stream>>>(p); RT.launches++; }
    CK(cudaGetLastError());
    mark_done(s);
    return PFCU_OK;
}

int pfcu_submit(pfcu_surface *s, const pfcu_state *states, uint32_t n_states, const pfcu_triangle *tris, uint32_t n_tris)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || (n_tris && (!states || !tris || n_states == 0))) return PFCU_ERR_INVALID;
    if (n_tris == 0) return PFCU_OK;
    if (MULTI_SURF(s)) return multi_run_all([&](int k) -> int {
        const std::vector<pfcu_state> st = states_for_device(states, n_states, k);
        return pfcu_submit(s->rep[k], st.data(), n_states, tris, n_tris); });
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
    RT.bytes_h2d += n_states * sizeof(DevState) + (size_t)n_tris * sizeof(pfcu_triangle);

    /* triangles: one DMA from pinned memory, or staged through our own pinned buffer */
    const size_t bytes = (size_t)n_tris * sizeof(pfcu_triangle);
    PinnedBlock *pb = find_pinned(tris);
    if (pb) {
        CK(cudaMemcpyAsync(LN.d_tris, tris, bytes, cudaMemcpyHostToDevice, LN.stream));
        if ((rc = pinned_in_flight(pb, LN.stream))) return rc;
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
    if (!RT.ok || !states || !tris || n_states == 0 || n_tris == 0) return nullptr;
    if (MULTI_HERE()) {
        pfcu_batch *r[MAX_DEVS] = { nullptr };
        const int rc = multi_run_all([&](int k) -> int {
            const std::vector<pfcu_state> st = states_for_device(states, n_states, k);
            r[k] = pfcu_batch_upload(st.data(), n_states, tris, n_tris); return r[k] ? PFCU_OK : PFCU_ERR_OOM; });
        if (rc) { multi_run_all([&](int k) -> int { if (r[k]) pfcu_batch_destroy(r[k]); return (int)PFCU_OK; }); return nullptr; }
        for (int k = 0; k < mg.n; k++) r[0]->rep[k] = r[k];
        return r[0];
    }
    pfcu_batch *b = (pfcu_batch *)calloc(1, sizeof *b);
    if (!b) return nullptr;
    RT.cur = &RT.lanes[0];
    std::vector<DevState> tmp(n_states);
    b->feature_mask = convert_states(states, n_states, tmp.data());
    b->single_prog = g_last_single_prog;
    b->leader_tex = g_last_leader_tex; g_last_leader_tex = false;
    b->pix_tex = g_last_pix_tex; g_last_pix_tex = false;
    new (&b->deps) std::vector<pfcu_surface *>(RT.deps);
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
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (!s || !b) return PFCU_ERR_INVALID;
    if (MULTI_SURF(s) && b->rep[1]) return multi_run_all([&](int k) -> int { return pfcu_batch_submit(s->rep[k], b->rep[k]); });
    use_lane(s);
    RT.deps.clear();
    for (pfcu_surface *dep : b->deps) RT.deps.push_back(dep);
    g_last_leader_tex = b->leader_tex; g_last_pix_tex = b->pix_tex;
    return launch_pipeline(s, b->tris, b->states, b->n_tris, b->feature_mask, b->single_prog);
}

void pfcu_batch_destroy(pfcu_batch *b)
{
    API_LOCK;
    if (!b) return;
    if (MULTI_HERE() && b->rep[1]) {
        pfcu_batch *r[MAX_DEVS]; memcpy(r, b->rep, sizeof r);
        multi_run_all([&](int k) -> int { pfcu_batch_destroy(r[k]); return (int)PFCU_OK; });
        return;
    }
    if (RT.ok) sync_all_lanes();
    cudaFree(b->states); cudaFree(b->tris); b->deps.~vector(); free(b);
}

void pfcu_profile_enable(int on) { RT.profiling = on != 0; }
void pfcu_set_raster_path(int path) { API_LOCK; if (MULTI_HERE()) { multi_run_all([path](int) -> int { pfcu_set_raster_path(path); return (int)PFCU_OK; }); return; } RT.raster_path = (path == PFCU_RASTER_TILES || path == PFCU_RASTER_FRAGMENTS) ? path : PFCU_RASTER_AUTO; }

int pfcu_profile_read(pfcu_profile *out)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    sync_all_lanes();
    out->raster_ms = 0; out->frontend_ms = 0; out->raster_launches = 0;
    for (size_t i = 0; i + 2 < RT.prof_events.size(); i += 3) {
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, RT.prof_events[i], RT.prof_events[i + 1]));
        CK(cudaEventElapsedTime(&b, RT.prof_events[i + 1], RT.prof_events[i + 2]));
        out->frontend_ms += a; out->raster_ms += b; out->raster_launches++;
    }
    for (auto e : RT.prof_events) RT.prof_pool.push_back(e);
    RT.prof_events.clear();
    return PFCU_OK;
}

int pfcu_finish(void)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (MULTI_HERE()) return multi_run_all([](int) -> int { return pfcu_finish(); });
    for (int i = 0; i < RT.n_lanes; i++) CK(cudaStreamSynchronize(RT.lanes[i].stream));
    return PFCU_OK;
}

int pfcu_get_counters(pfcu_counters *out)
{
    API_LOCK;
    if (!RT.ok) return PFCU_ERR_NO_DEVICE;
    if (MULTI_HERE()) {
        /* triangles are counted by device 0 (every device sets all of them up); pixels, launches and bytes are sums */
        pfcu_counters c[MAX_DEVS];
        const int rc = multi_run_all([&](int d) -> int { return pfcu_get_counters(&c[d]); });
        if (rc) return rc;
        *out = c[0];
        for (int d = 1; d < mg.n; d++) {
            out->pixels_shaded += c[d].pixels_shaded; out->pixels_depth_failed += c[d].pixels_depth_failed;
            out->kernel_launches += c[d].kernel_launches; out->bytes_h2d += c[d].bytes_h2d; out->bytes_d2h += c[d].bytes_d2h;
        }
        return PFCU_OK;
    }
    unsigned long long h[4] = { 0, 0, 0, 0 };
    sync_all_lanes();
    CK(cudaMemcpy(h, RT.d_counters, sizeof h, cudaMemcpyDeviceToHost));
    out->triangles_submitted = RT.submitted + h[3];
    out->triangles_rasterised = h[0];
    out->pixels_shaded = h[1];
    out->pixels_depth_failed = h[2];
    out->kernel_launches = RT.launches;
    out->bytes_h2d = RT.bytes_h2d; out->bytes_d2h = RT.bytes_d2h;
    return PFCU_OK;
}

void pfcu_reset_counters(void)
{
    API_LOCK;
    if (!RT.ok) return;
    if (MULTI_HERE()) { multi_run_all([](int) -> int { pfcu_reset_counters(); return (int)PFCU_OK; }); return; }
    sync_all_lanes();
    cudaMemset(RT.d_counters, 0, 4 * sizeof(unsigned long long));
    RT.submitted = 0; RT.launches = 0; RT.bytes_h2d = 0; RT.bytes_d2h = 0;
}

} /* extern "C" */
