/*
 * pf_internal.h - private state of the C99 front end (not installed).
 *
 * The front end keeps the OpenGL-1.x style state machine of the reference
 * (src/internal/context/context.h:322-401) but, instead of rasterising inside pfVertex*, it runs the
 * per-vertex stage on the host and appends screen-space triangles to an ordered batch that the
 * pfcu C-ABI (include/pfcu.h) consumes.
 */
#ifndef PF_INTERNAL_H
#define PF_INTERNAL_H

#include "pixelforge.h"
#include "pfcu.h"
#include "../pf_vstage.h"

#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* constants that change results (reference: src/internal/config.h:25-55) */
#define PFH_PROJECTION_STACK   2
#define PFH_MODELVIEW_STACK    8
#define PFH_TEXTURE_STACK      4
#define PFH_MAX_LIGHTS         8
#define PFH_MAX_POLY_VERTS     12
#define PFH_CLIP_EPSILON       1e-5f
#define PFH_PI                 3.14159265358979323846
#define PFH_DEG2RAD            (PFH_PI / 180.0)

typedef float pf_mat4[16];

/* texture object behind a PFtexture handle (reference: struct PFItex, context.h:160-178) */
typedef struct pf_tex {
    void           *pixels;         /* caller-owned host texels (or calloc'ed by pfGenFramebuffer)   */
    PFsizei         w, h;
    PFpixelformat   format;
    PFdatatype      type;
    PFtexturewrap   wrap;
    PFtexturefilter filter;
    pfcu_texture   *dev;            /* device copy, created on first use in a draw                   */
    struct pf_surf *surf;           /* non-NULL when this texture is a render target                 */
} pf_tex;

/* a render target: colour texture + depth, with its device surface and host-mirror bookkeeping */
typedef struct pf_surf {
    pfcu_surface *dev;
    pf_tex       *tex;              /* colour; tex->pixels is the host mirror                        */
    PFfloat      *zhost;            /* host depth mirror (NULL: none kept)                           */
    int           z_public;         /* zbuffer pointer is visible to the user (PFframebuffer)        */
    int           host_newer;       /* host mirror modified since the last upload                    */
    int           dev_newer;        /* device modified since the last download                       */
    PFuint        dirty_y0, dirty_y1; /* rows [y0, y1) touched on the device since last download     */
    int           readback_queued;  /* the dirty rows are already on their way to the (page-locked) mirror   */
    pfcu_texture *as_texture;       /* alias used when the colour buffer is sampled                  */
    void         *pinned_color;     /* host mirror ranges page-locked in place (NULL: not pinned)    */
    void         *pinned_depth;
    struct pf_surf *next;           /* registry link (lookup from a PFframebuffer copy)              */
} pf_surf;

typedef pfv_vertex pf_vertex;         /* shared with the device vertex stage (pf_vstage.h); colour as a dword */

typedef struct {
    float   position[3], direction[3];
    float   innerCutOff, outerCutOff;
    float   attConstant, attLinear, attQuadratic;
    PFcolor ambient, diffuse, specular;
    int     next;                   /* index of next active light, -1 = end                         */
} pf_light;

typedef struct { PFcolor ambient, diffuse, specular, emission; float shininess; } pf_material;

typedef struct { const void *buffer; PFsizei stride; PFint size; PFdatatype type; } pf_attrib;

typedef struct { PFfogmode mode; float density, start, end; PFcolor color; } pf_fog;

typedef struct {
    float *data; size_t size, cap, elem;    /* elem = floats (or dwords) per entry */
} pf_fvec;

typedef struct {
    pf_material material[2];
    pf_fvec positions, texcoords, normals, colors;
    PFtexture texture;
    PFdrawmode mode;
} pf_call;

/* A render list and its device-resident forms: the list's triangles assembled once per face mode
 * (which faces a primitive is drawn for depends on the cull state at REPLAY time: 0 front only, 1 back only, 2 both,
 * internal/context/context.c:94-244) and kept in HBM as unprocessed triangles (pfcu_list). */
typedef struct {
    pf_call *calls; size_t size, cap;
    pfcu_list *dev[3]; uint32_t dev_tris[3]; int dev_state[3];     /* dev_state: 0 not built, 1 built, -1 cannot be built */
    int colors_uniform;                                            /* every call carries one colour (checked when first compiled) */
    int analysed;
} pf_list;

typedef struct {
    pf_material material[2];
    float texcoord[2], normal[3];
    PFcolor color;
    PFtexture texture;
    PFuint state;
} pf_backup;

typedef struct pf_ctx {
    /* targets */
    pf_surf       *main_surf;
    PFframebuffer  mainFramebuffer;        /* {texture handle, NULL}                                 */
    PFframebuffer *bindedFramebuffer;
    pf_surf       *cur_surf;               /* surface current draws go to                            */
    void          *auxFramebuffer;
    pf_tex        *currentTexture;

    /* viewport */
    PFint   vpPos[2]; PFsizei vpDim[2]; PFint vpMin[2], vpMax[2];

    /* immediate mode */
    pf_vertex  vertexBuffer[6];
    PFsizei    vertexCounter;
    float      currentNormal[3], currentTexcoord[2];
    PFcolor    currentColor;
    PFdrawmode currentDrawMode;
    pf_attrib  apos, anrm, acol, atex;

    /* fixed-function state */
    PFcolor clearColor; float clearDepth, pointSize, lineWidth;
    float rasterPos[4], pixelZoom[2];
    PFpolygonmode polygonMode[2];
    pf_material   material[2];
    PFface cmFace; PFenum cmMode;
    pf_light lights[PFH_MAX_LIGHTS]; int activeHead;
    pf_fog fog;
    PFblendmode blendMode; PFdepthmode depthMode;
    PFshademode shadingMode; PFlightmode lightingMode; PFface cullFace;
    PFerrcode errCode; PFuint state;

    /* matrices */
    pf_mat4 matProjection, matTexture, matModel, matView, matMVP, matNormal;
    pf_mat4 stackProjection[PFH_PROJECTION_STACK], stackModelview[PFH_MODELVIEW_STACK], stackTexture[PFH_TEXTURE_STACK];
    PFsizei nProjection, nModelview, nTexture;
    PFmatrixmode matrixMode; float *currentMatrix; int modelMatrixUsed;
    float viewPos[3]; int viewPosValid;

    /* render lists */
    pf_list  *recording;
    pf_backup backup;
    int       replaying;

    /* batching */
    pfcu_triangle *tris[2];     /* pinned double buffer */
    int            cur_buf;
    uint32_t       n_tris, tri_cap;
    pfcu_state    *states; uint32_t n_states, state_cap;
    int            state_dirty;
    uint64_t       tris_emitted;
    /* per-triangle prologue environments of the pending batch (pfcu_vparams_lit) and the raw-triangle mode:
       a batch holds either processed pfcu_triangle records or unprocessed pfcu_rawtri records (same buffer) */
    pfcu_vparams_lit *vparams; uint32_t n_vparams, vparams_cap;
    uint32_t       vp_epoch, vp_epoch_built;    /* bumped by every call that changes what the prologue reads */
    pf_material    vp_material[2];              /* materials of the newest entry (may change per vertex with PF_COLOR_MATERIAL) */
    int            batch_raw;
    pfcu_prim     *prims; uint32_t n_prims, prims_cap;      /* pending points / lines (never together with triangles) */
    /* pending replays of device-resident render lists (never together with triangles either): segments in replay order,
       one pfcu_list_call per recorded call of each; states[] / vparams[] above hold what the calls refer to.  A pfClear
       issued in explicit sync mode waits here too, so that clear + replays of many contexts go out as ONE submission
       (pfcu_submit_list_jobs); any other drawing issues it at once. */
    pfcu_list_segment *segs; uint32_t n_segs, segs_cap;
    pfcu_list_call    *lcalls; uint32_t n_lcalls, lcalls_cap;
    uint32_t       list_tris;
    int            clear_pending; uint32_t clear_rgba; float clear_z;
    int            registered;                               /* in the calling thread's table of contexts with pending list work */
    /* list compilation: pfh_process_primitive appends assembled, unprocessed triangles here instead of batching them */
    int            compiling; uint32_t cmp_call; pfcu_rawtri *cmp_tris; uint32_t cmp_n, cmp_cap;
    float         *pow_tables; float *pow_shininess; uint32_t n_pow, pow_cap;   /* specular tables by shininess */
    /* optional capture of the submitted stream (pfxCaptureBegin/End) */
    int            device_vertex;   /* large vertex-array draws run the vertex stage on the GPU (default on) */
    int            capturing;
    pfcu_triangle *cap_tris; size_t cap_ntris, cap_tris_cap;
    pfcu_state    *cap_states; size_t cap_nstates, cap_states_cap;
} pf_ctx;

extern PF_CTX_DECL pf_ctx *pf_cur;

/* pf_pipeline.c */
void pfh_process_primitive(pf_ctx *c);                 /* vertexBuffer full -> triangles -> batch       */
void pfh_flush(pf_ctx *c);                             /* submit the pending batch (asynchronous)       */
pf_surf *pfh_surf_first(void);                          /* registry of live surfaces                    */
pf_surf *pfh_surf_next(pf_surf *s);
void pfh_sync_surface(pf_ctx *c, pf_surf *s);          /* flush + bring the host mirror up to date      */
void pfh_queue_readback(pf_ctx *c, pf_surf *s);        /* explicit mode: start the read-back without waiting */
void pfh_upload_if_needed(pf_ctx *c, pf_surf *s);      /* host mirror -> device when host is newer      */
void pfh_end_of_draw(pf_ctx *c);                       /* PF_CUDA_SYNC=end policy hook                  */
int  pfh_sync_mode_explicit(void);
void pfh_set_sync_mode(int explicit_mode);
void pfh_update_matrices(pf_ctx *c, int with_normal);
void pfh_vstage_params(const pf_ctx *c, pfv_params *p);
void pfh_update_view_pos(pf_ctx *c);
#define PFH_VP_TOUCH(c) ((c)->vp_epoch++)
void pfh_snapshot_state(pf_ctx *c, pfcu_state *st);
int  pfh_call_list_device(pf_ctx *c, pf_list *l);      /* 1 = the replay was queued as device-resident segments */
void pfh_lists_flush_all(pf_ctx *c);                   /* submit the pending list work of every context of this thread */
int  pfh_lists_pending(void);
void pfh_issue_pending_clear(pf_ctx *c);
int  pfh_clear_deferrable(pf_ctx *c);
void pfh_register_pending(pf_ctx *c);
void pfh_list_release_device(pf_list *l);
PFsizei pfh_verts_per_primitive(PFdrawmode m);
void pfh_carry_over(pf_ctx *c);
int  pfh_device_draw(pf_ctx *c, PFsizei count, PFint first, int indexed, PFdatatype itype, const void *indices,
                     int useNrm, int useTex, int useCol);    /* 1 = drawn on the device vertex path */   /* the state the next primitive would use */

/* pf_objects.c */
pf_surf *pfh_surf_create(pf_tex *tex, PFfloat *zhost, int z_public);
void     pfh_surf_destroy(pf_surf *s);
pf_surf *pfh_surf_lookup(PFtexture tex);
int      pfh_tex_format_code(PFpixelformat f, PFdatatype t);   /* PFCU_TEX_* or -1 */
int      pfh_texture_code(PFpixelformat f, PFdatatype t);      /* ... or PFCU_TEX_PIX + pair code for the other texel layouts */
PFsizei  pfh_pixel_bytes(PFpixelformat f, PFdatatype t);
PFcolor  pfh_pixel_get(const pf_tex *t, size_t i);            /* host mirror access (RGBA8 family)     */
void     pfh_pixel_set(pf_tex *t, size_t i, PFcolor c);
int      pfh_runtime_init(void);                               /* pfcu_init + approx tables, once      */

/* pf_x86approx.c */
int pfh_harvest_rcp(uint32_t **table, int *bits);
int pfh_harvest_rsqrt(uint32_t **table, int *bits);

#endif
