/*
 * pfcu.h - the thin C-ABI between the C99 host state machine (pixelforge.h front end) and the
 * hand-written sm_100a rasteriser.  Plain pointers and sizes only: no C++, no torch types.
 *
 * What it replaces in the reference (Bigfoot71/PixelForge):
 *   - the internal seam  pfiProcessRasterize_TRIANGLE* -> static Rasterize_Triangle(face, is3D, v1, v2, v3, viewPos)
 *     (src/internal/primitives/primitives.h:31-33, src/internal/primitives/triangles.c:53-55,287-558),
 *     whose implicit inputs are read from the global context (triangles.c:373-396);
 *   - the surface side effects of pfClear (src/context.c:680-787) and the framebuffer / texture
 *     storage of src/framebuffer.c:29-64 and src/texture.c:30-69.
 *
 * Instead of one synchronous call per triangle the host appends screen-space triangles (the output
 * of the reference's Process_ProjectAndClipTriangle, triangles.c:246-280) to an ordered batch, tags
 * each with the index of a state snapshot, and hands the batch to pfcu_submit().  Submission order
 * is preserved per pixel, so blending and depth results equal the reference's sequential execution.
 *
 * Two implementations of this ABI exist:
 *   - pixelforge_b200/csrc/pfcu.cu          the product (CUDA, sm_100a) -> libpfcu.so / libpixelforge.so
 *   - oracle/pfcu_oracle.c                  TEST ONLY scalar C restatement of the reference algorithm
 */
#ifndef PFCU_H
#define PFCU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef PFCU_API
#  define PFCU_API __attribute__((visibility("default")))
#endif

/* ---- status codes --------------------------------------------------------------------------- */
enum {
    PFCU_OK            = 0,
    PFCU_ERR_NO_DEVICE = 1,   /* no CUDA device / driver: the product never falls back to the CPU */
    PFCU_ERR_OOM       = 2,
    PFCU_ERR_INVALID   = 3,
    PFCU_ERR_CUDA      = 4
};

/* ---- state snapshot flags (what selects the per-fragment program, triangles.c:390-396) ------ */
enum {
    PFCU_ST_BLEND      = 1u << 0,   /* PF_BLEND enabled                                  */
    PFCU_ST_DEPTH_TEST = 1u << 1,   /* PF_DEPTH_TEST enabled                             */
    PFCU_ST_TEXTURE    = 1u << 2,   /* PF_TEXTURE_2D enabled and a texture is bound      */
    PFCU_ST_PHONG      = 1u << 3,   /* PF_LIGHTING && lightModel == PF_PHONG && lights   */
    PFCU_ST_SMOOTH     = 1u << 4    /* shade model PF_SMOOTH (else PF_FLAT)              */
};

/* texel formats understood by the sampler (reference getters: src/internal/pixel.h:2604-2632,2909-2920) */
enum {
    PFCU_TEX_RGBA8 = 0,   /* PF_RGBA / PF_UNSIGNED_BYTE */
    PFCU_TEX_BGRA8 = 1,   /* PF_BGRA / PF_UNSIGNED_BYTE */
    PFCU_TEX_RGB8  = 2,   /* PF_RGB  / PF_UNSIGNED_BYTE  (3 bytes per texel, alpha reads 255) */
    PFCU_TEX_BGR8  = 3    /* PF_BGR  / PF_UNSIGNED_BYTE */
};
/* every other (PFpixelformat, PFdatatype) pair the reference has a SIMD texel getter for (pixel.h:2249-3040: single
   channels, luminance(-alpha), 5-6-5 / 5-5-5-1 / 4-4-4-4, half, float): PFCU_TEX_PIX + PFCU_PIX(format, type).
   Textures only; surfaces take the four codes above. */
#define PFCU_TEX_PIX 256

typedef struct pfcu_surface pfcu_surface;   /* colour RGBA8 + depth f32, row-major [y*W+x], in HBM   */
typedef struct pfcu_texture pfcu_texture;   /* texel array in HBM                                    */
typedef struct pfcu_batch   pfcu_batch;     /* triangles + states already resident in HBM            */

/* One projected vertex: the fields of PFIvertex (src/internal/context/context.h:203-210) that
 * Rasterize_Triangle reads.  48 bytes. */
typedef struct {
    float    sx, sy;        /* screen[0..1], +0.5 biased (internal/context/context.c:63-64)          */
    float    zinv;          /* homogeneous[2]: 1/z_clip for 3D, z_clip for "2D" (triangles.c:257-267) */
    float    u, v;          /* texcoord, pre-multiplied by zinv for 3D (triangles.c:269)             */
    float    px, py, pz;    /* object-space position (Phong only)                                    */
    float    nx, ny, nz;    /* normal (Phong only)                                                   */
    uint32_t rgba;          /* PFcolor as a little-endian dword: r | g<<8 | b<<16 | a<<24            */
} pfcu_vertex;

/* One Rasterize_Triangle invocation.  152 bytes. */
typedef struct {
    pfcu_vertex v[3];
    uint32_t    state;      /* index into the batch's state table                                    */
    uint8_t     face;       /* PF_FRONT (0) or PF_BACK (1): the faceToRender argument                */
    uint8_t     is3d;       /* the is3D argument (perspective uv, no viewport clamp of the bbox)     */
    uint16_t    pad;
} pfcu_triangle;

typedef struct {            /* PFIlight (internal/context/context.h:216-228) minus the list link     */
    float    position[3];
    float    direction[3];
    float    inner_cutoff, outer_cutoff;
    float    att_constant, att_linear, att_quadratic;
    uint32_t ambient, diffuse, specular;    /* PFcolor dwords                                         */
} pfcu_light;

typedef struct {            /* PFImaterial (internal/context/context.h:233-239)                       */
    uint32_t ambient, diffuse, specular, emission;
    float    shininess;
} pfcu_material;

/* Everything Rasterize_Triangle reads from G_currentCtx (triangles.c:318-321,373-396,520). */
typedef struct {
    uint32_t            flags;          /* PFCU_ST_*                                                  */
    uint8_t             blend_mode;     /* PFblendmode                                                */
    uint8_t             depth_func;     /* PFdepthmode                                                */
    uint8_t             tex_filter;     /* PFtexturefilter                                            */
    uint8_t             tex_wrap;       /* PFtexturewrap                                              */
    int32_t             vp_min[2];      /* ctx->vpMin, bbox clamp for "2D" triangles                  */
    int32_t             vp_max[2];      /* ctx->vpMax                                                 */
    const pfcu_texture *texture;        /* NULL when PFCU_ST_TEXTURE is clear                         */
    uint32_t            n_lights;       /* active lights in enable order (Phong only)                 */
    pfcu_light          lights[8];
    pfcu_material       material[2];    /* faceMaterial[PF_FRONT], faceMaterial[PF_BACK]              */
    float               view_pos[3];    /* translation row of inverse(matView) (triangles.c:84-86)    */
    uint32_t            pad;
} pfcu_state;

typedef struct {
    uint64_t triangles_submitted;   /* Rasterize_Triangle invocations                                 */
    uint64_t triangles_rasterised;  /* ... that survive the face / zero-area test (triangles.c:303)   */
    uint64_t pixels_shaded;         /* fragments whose final mask lane is set: colour+depth written   */
    uint64_t pixels_depth_failed;   /* covered fragments rejected by the depth test                   */
    uint64_t kernel_launches;       /* kernels launched by this library                               */
    uint64_t bytes_h2d;             /* host -> device bytes copied (batches, vertex arrays, surfaces, textures) */
    uint64_t bytes_d2h;             /* device -> host bytes copied (surface downloads)                */
} pfcu_counters;

/* ---- runtime -------------------------------------------------------------------------------- */

/* Multi-device mode.  With PF_CUDA_DEVICES=a,b,c... in the environment at pfcu_init() the library drives every listed GPU of
 * the box from the one process: every surface, texture and resident batch created afterwards exists once per device, every
 * call below that takes such a handle is replayed on all devices (the handle the caller holds is the first device's), a
 * device rasterises only the 64x64 tiles it owns of the large RGBA8 surfaces (tile % n == device) and read-backs collect the
 * tiles (pfcu_surface_download*, pfcu_surface_read_pixels).  Images, depth buffers and counters are what one GPU produces.
 * Not available in that mode: pfcu_surface_wrap, the explicit tile-owner / pack / IPC-present calls of the multi-PROCESS
 * split (they keep acting on the first device only) and device-resident list jobs (pfcu_list_job_supported returns 0). */

/* Bind to CUDA device `device` (-1: $PF_CUDA_DEVICE, else $LOCAL_RANK, else 0) and create the stream.
 * Idempotent.  Returns PFCU_ERR_NO_DEVICE when no GPU is usable. */
PFCU_API int  pfcu_init(int device);
PFCU_API void pfcu_shutdown(void);
PFCU_API const char *pfcu_last_error(void);
PFCU_API const char *pfcu_backend_name(void);          /* "cuda-sm_100a" or "oracle-c" (tests)      */

/* Use an externally owned CUDA stream (cudaStream_t passed as void*) for all later work. */
PFCU_API int  pfcu_set_stream(void *cuda_stream);
PFCU_API void *pfcu_get_stream(void);

/* Page-locked host memory for batches: pfcu_submit() from such a block is a single asynchronous
 * DMA with no staging copy.  pfcu_host_wait(p) blocks until the last submit that read from the
 * block starting at p has finished reading it (so the host may overwrite it). */
PFCU_API void *pfcu_host_alloc(size_t bytes);
PFCU_API void  pfcu_host_free(void *p);
PFCU_API int   pfcu_host_wait(const void *p);
/* Page-lock caller-owned memory in place (the application's target buffer) so that surface downloads are
 * direct DMA instead of staged copies.  Best effort: returns non-zero when the range cannot be pinned. */
PFCU_API int   pfcu_host_register(void *p, size_t bytes);
PFCU_API void  pfcu_host_unregister(void *p);

/* Static geometry: a block from pfcu_host_alloc whose content does not change from draw to draw is mirrored in device
 * memory at the first pfcu_draw_triangles that reads an array out of it; later draws read the mirror and nothing crosses
 * PCIe.  pfcu_host_modified(p) tells the library that the application rewrote (part of) the block: the next draw uploads
 * it again.  Both return PFCU_ERR_INVALID for memory that is not a pfcu_host_alloc block. */
PFCU_API int   pfcu_host_set_static(void *p, int on);
PFCU_API int   pfcu_host_modified(void *p);
PFCU_API int   pfcu_host_is_static(const void *p);      /* 1 when p lies in a block declared static */

/* Tables that reproduce the host's RCPPS / RSQRTPS (reference: src/internal/simd.h:1217-1245).
 * rcp[i], i = top `rcp_bits` mantissa bits: float bits of rcp(1.m);  rsqrt[(odd<<rsqrt_bits)|i]:
 * float bits of rsqrt(1.m * 2^odd).  Harvested by the host library from the CPU it runs on. */
PFCU_API int  pfcu_set_approx_tables(const uint32_t *rcp, int rcp_bits, const uint32_t *rsqrt, int rsqrt_bits);

/* ---- surfaces (framebuffer.c:29-64, context.c:126-138) -------------------------------------- */
PFCU_API pfcu_surface *pfcu_surface_create(uint32_t width, uint32_t height);           /* PF_RGBA / PF_UNSIGNED_BYTE */
/* Render targets in the other 8-bit formats (PFCU_TEX_BGRA8 / RGB8 / BGR8).  On the device every
 * surface is held in ONE canonical layout - colour as RGBA8 dwords (alpha 255 for the 3-byte formats, which is what
 * their getters return, pixel.h:576-588,2604-2632), depth f32 - and pfcu_surface_upload / download convert from / to
 * the caller's layout (3 or 4 bytes per pixel, row y at byte offset y * width * bytes).  What the format changes in
 * the triangle path is the reference's own SIMD behaviour: the BGRA8 setter and getter shuffle bytes
 * with a mask that addresses only the first dword of each 128-bit half (pixel.h:2069-2078,2915-2920, simd.h:563-583),
 * so the four pixels x0+4k .. x0+4k+3 of a triangle's row (x0 = its bbox xMin) all receive the blended fragment of
 * the first one, and blend against that pixel's colour.  Batches that target a non-RGBA8 surface, or sample a BGRA8
 * texture (same shuffle in the texel getter), are rasterised by a row-ordered kernel that reproduces this exactly. */
PFCU_API pfcu_surface *pfcu_surface_create_format(uint32_t width, uint32_t height, int pfcu_tex_format);
PFCU_API int pfcu_surface_format(const pfcu_surface *s);
/* Wrap caller-owned device memory (e.g. a torch tensor): colour u32[w*h], depth f32[w*h]. */
PFCU_API pfcu_surface *pfcu_surface_wrap(void *dev_color, void *dev_depth, uint32_t width, uint32_t height);
PFCU_API void     pfcu_surface_destroy(pfcu_surface *s);
PFCU_API uint32_t pfcu_surface_width(const pfcu_surface *s);
PFCU_API uint32_t pfcu_surface_height(const pfcu_surface *s);
PFCU_API void    *pfcu_surface_color_ptr(const pfcu_surface *s);    /* device pointers */
PFCU_API void    *pfcu_surface_depth_ptr(const pfcu_surface *s);
/* Host <-> device copies of whole rows [y0, y0+rows); either pointer may be NULL.  Host pointers
 * address the full surface (row-major, row y at offset y*W).  Asynchronous on the stream for
 * uploads; downloads return after the data is on the host. */
PFCU_API int pfcu_surface_upload(pfcu_surface *s, const void *host_color, const float *host_depth, uint32_t y0, uint32_t rows);
PFCU_API int pfcu_surface_download(pfcu_surface *s, void *host_color, float *host_depth, uint32_t y0, uint32_t rows);
/* The same copies enqueued behind the surface's pending work WITHOUT waiting (meant for page-locked host memory,
 * see pfcu_host_register); pfcu_surface_wait() returns once everything enqueued for the surface has completed.
 * The front end uses the pair to read a context back while the next contexts are still being drawn. */
PFCU_API int pfcu_surface_download_async(pfcu_surface *s, void *host_color, float *host_depth, uint32_t y0, uint32_t rows);
PFCU_API int pfcu_surface_wait(pfcu_surface *s);
/* Plain fill of every pixel (pfClearFramebuffer, framebuffer.c:89-102). */
PFCU_API int pfcu_surface_fill(pfcu_surface *s, int do_color, uint32_t rgba, int do_depth, float depth);
/* pfClear with the reference's exact SIMD behaviour (context.c:696-713): pixels
 * [8, size - size%8) receive the value, pixels 0..7 are left alone, the tail copies pixel 0. */
PFCU_API int pfcu_surface_clear_ref(pfcu_surface *s, int do_color, uint32_t rgba, int do_depth, float depth);
/* Multi-GPU screen-tile split: this process only rasterises tiles whose owner == rank
 * (owner(tile) = (tile_x + tile_y * tiles_x) % world).  world <= 1 disables the split. */
PFCU_API int pfcu_surface_set_tile_owner(pfcu_surface *s, uint32_t rank, uint32_t world);
/* Pack this rank's tiles into / unpack rank r's tiles from a contiguous device staging buffer
 * (colour then depth per tile), for the NCCL gather to the presenting rank. */
PFCU_API size_t pfcu_surface_owned_bytes(const pfcu_surface *s, uint32_t rank, uint32_t world, int with_depth);
PFCU_API int pfcu_surface_pack_tiles(pfcu_surface *s, uint32_t rank, uint32_t world, int with_depth, void *dev_staging);
PFCU_API int pfcu_surface_unpack_tiles(pfcu_surface *s, uint32_t rank, uint32_t world, int with_depth, const void *dev_staging);

/* Present over peer memory (NVLink / NVSwitch), the alternative to pack -> NCCL gather -> unpack: the presenting
 * rank exports its surface (CUDA IPC), every other rank maps it, and pfcu_surface_push_tiles() stores the rank's own
 * tiles straight into the presenter's surface, 128 bits at a time - no staging buffer, no collective on the data path
 * (the ranks only need a barrier before the presenter reads).  A variant that issued these peer stores from inside
 * the rasterisers' tile write-back (overlapping the transfer with the shading of the remaining tiles) cost the tuned
 * k_raster 16 % on the single-GPU scenes and was dropped; the push kernel moves a rank's share of an 8K surface in
 * well under 0.1 ms.  handle buffers: 64 bytes each (cudaIpcMemHandle_t).  pfcu_surface_set_present_surface() is
 * the same with a surface of THIS process as the target (single-GPU tests, several surfaces on one device). */
PFCU_API int pfcu_surface_ipc_handles(pfcu_surface *s, void *color_handle, void *depth_handle);
PFCU_API int pfcu_surface_set_present_peer(pfcu_surface *s, const void *color_handle, const void *depth_handle);
PFCU_API int pfcu_surface_set_present_surface(pfcu_surface *s, pfcu_surface *target);
PFCU_API int pfcu_surface_clear_present(pfcu_surface *s);
PFCU_API int pfcu_surface_push_tiles(pfcu_surface *s, uint32_t rank, uint32_t world, int with_depth);

/* ---- full-surface operations on the device (SURVEY 8-f "next" row 3) -------------------------------------------
 * pfRect*, pfDrawPixels, pfFogProcess and pfReadPixels of the reference are loops over the framebuffer with the SCALAR
 * getters / setters, blend table and depth table (context.c:1938-1977, 1988-2084, 2275-2344, 2349-2395).  The front end
 * does what the reference does before its loop (projection of the corner / raster position, clamping to the viewport)
 * and the device runs the loop on the surface, ordered after everything submitted before - no read-back, no host
 * mirror.  Pixels are addressed as y * W + x like upstream (a column one past the right edge, which a viewport smaller
 * than the buffer allows - SURVEY Q20 - lands in the next row); addresses outside the buffer are dropped. */

/* pfRectf: every pixel of the inclusive rectangle takes `rgba` (context.c:1972-1976). */
PFCU_API int pfcu_surface_rect(pfcu_surface *s, int32_t x1, int32_t y1, int32_t x2, int32_t y2, uint32_t rgba);

/* pfFogProcess (context.c:2310-2342).  depth >= end: the fog colour (blended with PF_BLEND_ALPHA unless its alpha is
 * 255); start < depth < end: the fog colour with alpha (PFubyte)(t * alpha) blended over the pixel, t = (depth - start)
 * * inv_len (PF_LINEAR), 1 - expf(-density * (depth - start)) (PF_EXP) or 1 - exp2f(...) (PF_EXP2).  expf / exp2f are
 * the HOST libm's and are not correctly rounded, so the device cannot recompute them: for the two exponential modes the
 * front end tabulates, by bisection over float bit patterns with its own libm, thresholds[k-1] = the smallest depth in
 * (start, end) whose fog alpha is >= k (k = 1 .. n_thresholds) and the device counts thresholds <= depth - the same
 * device as the Gouraud specular tables (PFCU_POW_TABLE_SIZE above).  The oracle ignores the table and calls libm. */
typedef struct {
    float    start, end, inv_len;   /* inv_len = 1 / (end - start), the front end's IEEE single division        */
    float    density;
    uint32_t rgba;                  /* fog colour, alpha included                                                */
    uint32_t mode;                  /* PFfogmode: 0 linear, 1 exp, 2 exp2; 3: any other value (t = 0 upstream)   */
    const float *thresholds;        /* host pointer, exponential modes only                                      */
    uint32_t n_thresholds;          /* <= 255                                                                    */
} pfcu_fog;
PFCU_API int pfcu_surface_fog(pfcu_surface *s, const pfcu_fog *fog);

/* pfDrawPixels (context.c:2026-2077): nearest-neighbour copy of a host image into the inclusive rectangle
 * [xmin, xmax] x [ymin, ymax] (already clamped to the viewport), source texel ((PFsizei)(v * (height - 1))) * width +
 * (PFsizei)(u * (width - 1)) with u = (x - xs) * inv_xlen, v = (y - ys) * inv_ylen; depth test against the constant
 * `z` with the scalar table (depth.h:28-78; NOTEQUAL really is "not equal" there), z written where it passes, colour
 * through the scalar blend table (blend.h:29-130).  `format` is PFCU_PIX(PFpixelformat, PFdatatype): any of the 38 pairs the
 * reference has a getter for (pixel.h:764-812); 8-bit, 5-6-5 / 5-5-5-1 / 4-4-4-4, half and float components, single
 * channels and luminance (pf_pixfmt.h restates the getters and setters). */
#define PFCU_PIX(format, type) ((int)(format) * 16 + (int)(type))
typedef struct {
    const void *pixels; uint32_t width, height; int format;
    int32_t  xs, ys;                /* the raster position on the screen                                         */
    int32_t  xmin, ymin, xmax, ymax;
    float    inv_xlen, inv_ylen;    /* 1 / (width * zoom_x), 1 / (height * zoom_y)                               */
    float    z;
    uint32_t flags;                 /* PFCU_ST_BLEND | PFCU_ST_DEPTH_TEST                                        */
    uint8_t  blend_mode, depth_func; uint16_t pad;
} pfcu_pixels;
PFCU_API int pfcu_surface_draw_pixels(pfcu_surface *s, const pfcu_pixels *d);

/* pfReadPixels (context.c:2349-2395): the surface's pixels [x0, x0+cols) x [y0, y0+rows) converted to `format`
 * (PFCU_PIX(format, type), any pair the reference has a setter for) on the device and copied to host_pixels, pixel (x, y) at index (y - y0) * dst_width + (x - x0); nothing
 * else of the destination is touched.  Returns after the data is on the host. */
PFCU_API int pfcu_surface_read_pixels(pfcu_surface *s, uint32_t x0, uint32_t y0, uint32_t cols, uint32_t rows,
                                      uint32_t dst_width, int format, void *host_pixels);

/* ---- textures (texture.c:30-69) -------------------------------------------------------------- */
PFCU_API pfcu_texture *pfcu_texture_create(const void *host_pixels, uint32_t width, uint32_t height, int pfcu_tex_format);
PFCU_API pfcu_texture *pfcu_texture_from_surface(pfcu_surface *s);   /* render-to-texture alias, RGBA8 */
PFCU_API int  pfcu_texture_update(pfcu_texture *t, const void *host_pixels);
PFCU_API void pfcu_texture_destroy(pfcu_texture *t);

/* ---- optional device vertex stage (SURVEY 8-f "next" row 1) ---------------------------------- */
/* Parameters of the per-vertex stage: what the reference latches at pfBegin (context.c:96-111) and reads in
 * triangles.c:62-116,246-280 / internal/context/context.c:51-65. */
typedef struct {
    float    mvp[16];           /* model-view-projection, row-vector convention (v' = v * M)             */
    float    normal_mat[16];
    int32_t  vp_pos[2];
    uint32_t vp_dim[2];         /* viewport size - 1                                                     */
    uint32_t lighting;          /* PF_LIGHTING with active lights: normal transform + colour * diffuse   */
    uint32_t diffuse[2];        /* faceMaterial[PF_FRONT/PF_BACK].diffuse                                */
} pfcu_vparams;

/* One pfDrawElements / pfDrawArrays call in PF_TRIANGLES mode over tightly packed float arrays
 * (context.c:1231-1575).  Host pointers; copied to the device by the call. */
typedef struct {
    const float   *positions;  uint32_t pos_size;      /* 2..4 floats per vertex                          */
    const float   *normals;                            /* 3 floats per vertex, or NULL (normal = 0)       */
    const float   *texcoords;                          /* 2 floats per vertex, or NULL (texcoord = 0)     */
    const uint8_t *colors;     uint32_t color_size;    /* 3 or 4 ubytes per vertex, or NULL               */
    uint32_t       n_vertices;                         /* vertices referenced (max index + 1); 0 with 32-bit indices: the device finds it */
    const void    *indices;    uint32_t index_bytes;   /* 1, 2 or 4; NULL = sequential from `first`       */
    uint32_t       first, count;                       /* number of indices / vertices to draw            */
    uint32_t       current_color;                      /* colour when there is no colour array            */
    uint32_t       n_faces;    uint8_t faces[2];       /* faceToRender passes per triangle, in order      */
    uint16_t       pad;
} pfcu_draw;

enum { PFCU_CAP_DEVICE_VERTEX = 1u, PFCU_CAP_RAW_TRIANGLES = 2u, PFCU_CAP_LISTS = 4u };
PFCU_API unsigned pfcu_capabilities(void);
/* Vertex stage + rasterisation of one draw call, entirely on the device; ordered after everything
 * submitted before it.  `state` is the single state snapshot in force.  *n_out receives the number of
 * Rasterize_Triangle-equivalent triangles produced (after clipping). */
PFCU_API int pfcu_draw_triangles(pfcu_surface *s, const pfcu_state *state, const pfcu_vparams *vp,
                                 const pfcu_draw *draw, uint32_t *n_out);

/* ---- device vertex stage for immediate mode and render lists ------------------------------------ */
/* The front end assembles triangles (draw modes, face passes: context.c:41-74, internal/context/context.c:94-244)
 * and ships them UNPROCESSED: object-space vertices as pfVertex* latched them.  The device runs the reference's
 * whole per-triangle prologue: normal transform, colour * material diffuse, Gouraud vertex lighting
 * (lighting.c:23-144), MVP transform, clipping, perspective preparation and viewport mapping (triangles.c:62-116,
 * 157-280), then the usual setup -> bin -> raster pipeline, in submission order. */
typedef struct { float pos[4]; float normal[3]; float uv[2]; uint32_t rgba; } pfcu_rawvertex;      /* 40 B */
typedef struct {
    pfcu_rawvertex v[3];
    uint32_t state;             /* index into states[]                                                    */
    uint32_t vparams;           /* index into vparams[]                                                   */
    uint8_t  face;              /* faceToRender of this pass                                              */
    uint8_t  pad[3];
} pfcu_rawtri;                  /* 132 B */
/* Everything the per-triangle prologue reads from G_currentCtx besides the fragment state. */
typedef struct {
    pfcu_vparams  base;
    uint32_t      gouraud;      /* PF_LIGHTING with active lights and lightingMode == PF_GOURAUD          */
    uint32_t      n_lights;     /* active lights in list order                                            */
    float         view_z[3];    /* matView[8..10]: the sign of N . view_z selects the material (triangles.c:95-101) */
    float         view_pos[3];  /* translation row of inverse(matView)                                    */
    pfcu_light    lights[8];
    pfcu_material material[2];
    uint32_t      pow_table[2]; /* per face material: index of its specular table in pow_tables[]         */
} pfcu_vparams_lit;
/* Gouraud specular = (PFubyte)(255 * powf(max(N.H, 0), shininess)) with the HOST libm's powf (lighting.c:117).
 * powf is not correctly rounded, so the device cannot recompute it; the front end tabulates, per shininess value,
 * the 256 smallest inputs x_k with (int)(255 * powf(x_k, shininess)) >= k, k = 1..256 (found by bisection over
 * float bit patterns with the host's own powf), and the device counts thresholds <= x. */
#define PFCU_POW_TABLE_SIZE 256
PFCU_API int pfcu_submit_raw(pfcu_surface *s, const pfcu_state *states, uint32_t n_states,
                             const pfcu_vparams_lit *vparams, uint32_t n_vparams,
                             const float *pow_tables, uint32_t n_pow_tables,
                             const pfcu_rawtri *tris, uint32_t n_tris, uint32_t *n_out);

/* ---- device-resident render lists, many surfaces per launch (SURVEY 8-f "next" row 2) ------------------------
 * The reference replays a render list by re-issuing pfColor / pfTexCoord / pfNormal / pfVertex for every recorded
 * vertex (renderlist.c:71-97, internal/context/context.h:271-294).  Here the front end assembles a list's triangles
 * ONCE (draw modes, face passes) and leaves them in HBM as unprocessed pfcu_rawtri records whose `state` field holds
 * the index of the recorded pfBegin..pfEnd call they came from (`vparams` is unused).  A replay then ships only what
 * the reference reads from the context at replay time - per recorded call one pfcu_list_call naming the fragment state
 * and the prologue environment in force (matrices, lights, the call's materials), both deduplicated - a few KB.
 * pfcu_submit_list_jobs takes the pending replays of MANY surfaces (the contexts of a batch server, BASELINE config
 * C5) and runs them with four launches in all: pfClear of every surface, vertex stage (k_list_chain), setup + binning
 * (k_front_small) and rasterisation (k_raster_frag), the surface being the grid's y coordinate. */
typedef struct pfcu_list pfcu_list;
PFCU_API pfcu_list *pfcu_list_create(const pfcu_rawtri *tris, uint32_t n_tris);
PFCU_API void       pfcu_list_destroy(pfcu_list *l);
PFCU_API uint32_t   pfcu_list_size(const pfcu_list *l);
typedef struct {
    uint32_t state, vparams;    /* indices into the job's states[] / vparams[]                                          */
    uint32_t override_color;    /* PF_COLOR_MATERIAL at replay: pfColor feeds the material and every vertex carries ... */
    uint32_t rgba;              /* ... the context's current colour instead of the recorded one (context.c:1687-1717)   */
} pfcu_list_call;
typedef struct { const pfcu_list *list; uint32_t first_call; uint32_t pad; } pfcu_list_segment;   /* calls[first_call + tri.state] */
typedef struct {
    pfcu_surface *surface;
    uint32_t clear; uint32_t clear_rgba; float clear_depth;     /* clear != 0: pfcu_surface_clear_ref(both buffers) first */
    const pfcu_state *states;           uint32_t n_states;
    const pfcu_vparams_lit *vparams;    uint32_t n_vparams;
    const float *pow_tables;            uint32_t n_pow_tables;
    const pfcu_list_call *calls;        uint32_t n_calls;
    const pfcu_list_segment *segments;  uint32_t n_segments;    /* replayed in this order */
} pfcu_list_job;
#define PFCU_LIST_JOB_MAX_TRIS     1024u    /* assembled triangles per job (before clipping)                          */
#define PFCU_LIST_JOB_MAX_SEGMENTS 16u
/* Can a job of n_tris list triangles on this surface take this path?  (RGBA8 target, no tile split, size limits; a
 * front end that gets 0 replays the list through pfcu_submit_raw or on the host instead.) */
PFCU_API int pfcu_list_job_supported(const pfcu_surface *s, uint32_t n_tris, uint32_t n_segments);
PFCU_API int pfcu_submit_list_jobs(const pfcu_list_job *jobs, uint32_t n_jobs);

/* ---- points and lines (SURVEY 8-f "next" row 3) -------------------------------------------------- */
/* The front end transforms and clips (lines.c:137-281, points.c:62-83) and submits screen-space primitives;
 * the device walks them in submission order (one CTA per 64x64 tile, every CTA visits all primitives and keeps
 * the pixels of its tile, so overlapping primitives blend and depth-test in order).  Arithmetic: pf_prims.h. */
typedef struct {
    float    x1, y1, x2, y2;    /* screen coordinates (a point uses x1, y1)                               */
    float    z1, z2;            /* homogeneous z of the endpoints (not inverted: lines.c:301-302)         */
    uint32_t c1, c2;            /* PFcolor dwords                                                          */
    float    size;              /* ctx->lineWidth or ctx->pointSize                                        */
    uint8_t  kind;              /* 0 point, 1 line                                                          */
    uint8_t  flags;             /* PFCU_ST_BLEND | PFCU_ST_DEPTH_TEST                                       */
    uint8_t  blend_mode, depth_func;
} pfcu_prim;                    /* 40 B */
PFCU_API int pfcu_submit_prims(pfcu_surface *s, const pfcu_prim *prims, uint32_t n_prims);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* Rasterise `n_tris` triangles, in order, into `s`.  Host pointers; the call copies them to the
 * device (pinned staging + cudaMemcpyAsync) and launches setup -> bin -> tile raster.  Asynchronous. */
PFCU_API int pfcu_submit(pfcu_surface *s, const pfcu_state *states, uint32_t n_states,
                         const pfcu_triangle *tris, uint32_t n_tris);
/* Same, split in two so that a batch can stay resident in HBM and be replayed (benchmarks, lists). */
PFCU_API pfcu_batch *pfcu_batch_upload(const pfcu_state *states, uint32_t n_states,
                                       const pfcu_triangle *tris, uint32_t n_tris);
PFCU_API int  pfcu_batch_submit(pfcu_surface *s, pfcu_batch *b);
PFCU_API void pfcu_batch_destroy(pfcu_batch *b);

/* Optional device-side timing of the pipeline stages (CUDA events on the launching stream). */
typedef struct {
    double   raster_ms;         /* sum of k_raster launch durations since the last read            */
    double   frontend_ms;       /* sum of setup + binning durations                                 */
    uint64_t raster_launches;
} pfcu_profile;
PFCU_API void pfcu_profile_enable(int on);
/* Which tile rasteriser a batch runs on: PFCU_RASTER_AUTO (default; by triangles per tile), PFCU_RASTER_TILES
 * (k_raster: one triangle per warp step over 8x4 blocks, for large triangles) or PFCU_RASTER_FRAGMENTS
 * (k_raster_frag: fragment compaction, for many small triangles).  A performance knob only: both produce
 * identical pixels (tests run every parity case under both).  $PF_CUDA_FRAG=0|1|2 sets the start-up value
 * (0 tiles, 1 auto, 2 fragments). */
#define PFCU_RASTER_AUTO      0
#define PFCU_RASTER_TILES     1
#define PFCU_RASTER_FRAGMENTS 2
PFCU_API void pfcu_set_raster_path(int path);
PFCU_API int  pfcu_profile_read(pfcu_profile *out);    /* implies pfcu_finish(); resets the sums      */

/* Work on different surfaces may run on different internal streams ("lanes", $PF_CUDA_LANES, default 4) so
 * that independent contexts overlap.  pfcu_fence() orders everything enqueued so far, on every lane, before
 * everything enqueued later, without blocking the host; callers that time multi-surface work with events
 * on pfcu_get_stream() call it after recording the start event and before recording the end event. */
PFCU_API int  pfcu_fence(void);
PFCU_API int  pfcu_finish(void);                       /* wait for everything queued so far          */
PFCU_API int  pfcu_get_counters(pfcu_counters *out);   /* implies pfcu_finish()                      */
PFCU_API void pfcu_reset_counters(void);

#ifdef __cplusplus
}
#endif
#endif /* PFCU_H */
