/*
 * pf_pipeline.c - primitive assembly, the per-vertex stage and batching.
 *
 * Follows, in behaviour (not in code), the caller side of the reference's hot path:
 *   draw-mode dispatch          src/internal/context/context.c:94-244
 *   fan / strip splitting       src/internal/primitives/triangles.c:118-152
 *   per-triangle prologue       triangles.c:62-116  (normal transform, material multiply, Gouraud)
 *   Gouraud vertex lighting     src/internal/lighting/lighting.c:23-144 (integer Blinn-Phong)
 *   clip W / clip XYZ / project triangles.c:157-280, internal/context/context.c:51-90
 * The output of that stage - what the reference passes to Rasterize_Triangle - is appended to an
 * ordered batch of pfcu_triangle records and submitted through the pfcu C-ABI.
 */
#include "pf_internal.h"
#include "pf_math.h"
#include "../pf_prims.h"
#include "pfx.h"

#include <stdio.h>
#include <math.h>

/* ---- sync policy ---------------------------------------------------------------------------- */

static int g_sync_explicit = -1;

int pfh_sync_mode_explicit(void)
{
    if (g_sync_explicit < 0) {
        const char *e = getenv("PF_CUDA_SYNC");
        g_sync_explicit = (e && (strcmp(e, "explicit") == 0 || strcmp(e, "EXPLICIT") == 0)) ? 1 : 0;
    }
    return g_sync_explicit;
}

void pfh_set_sync_mode(int explicit_mode) { g_sync_explicit = explicit_mode ? 1 : 0; }

/* ---- matrices latched at pfBegin (context.c:96-111) ----------------------------------------- */

void pfh_update_matrices(pf_ctx *c, int with_normal)
{
    PFH_VP_TOUCH(c);
    if (c->modelMatrixUsed) {
        m4_mul(c->matMVP, c->matModel, c->matView);
        m4_mul(c->matMVP, c->matMVP, c->matProjection);
        if (with_normal && (c->state & PF_LIGHTING)) {
            m4_invert(c->matNormal, c->matModel);
            m4_transpose(c->matNormal, c->matNormal);
        }
    } else {
        m4_mul(c->matMVP, c->matView, c->matProjection);
        if (with_normal && (c->state & PF_LIGHTING)) m4_identity(c->matNormal);
    }
}

/* ---- state snapshots ------------------------------------------------------------------------- */

static uint32_t color_dword(PFcolor c) { uint32_t u; memcpy(&u, &c, 4); return u; }

static void fill_material(pfcu_material *d, const pf_material *s)
{
    d->ambient = color_dword(s->ambient);   d->diffuse = color_dword(s->diffuse);
    d->specular = color_dword(s->specular); d->emission = color_dword(s->emission);
    d->shininess = s->shininess;
}

static int lights_active(const pf_ctx *c) { return c->activeHead >= 0; }

/* Build the snapshot of everything the fragment stage reads (triangles.c:373-396).  Fields that
 * the selected fragment program does not read are zeroed so that equal programs compare equal. */
static void build_state(pf_ctx *c, pfcu_state *st)
{
    memset(st, 0, sizeof *st);
    pf_tex *tex = c->currentTexture;
    int texturing = (c->state & PF_TEXTURE_2D) && tex;
    int phong = (c->state & PF_LIGHTING) && c->lightingMode == PF_PHONG && lights_active(c);

    if (c->state & PF_BLEND)      { st->flags |= PFCU_ST_BLEND; st->blend_mode = (uint8_t)c->blendMode; }
    if (c->state & PF_DEPTH_TEST) { st->flags |= PFCU_ST_DEPTH_TEST; st->depth_func = (uint8_t)c->depthMode; }
    if (c->shadingMode == PF_SMOOTH) st->flags |= PFCU_ST_SMOOTH;
    st->vp_min[0] = c->vpMin[0]; st->vp_min[1] = c->vpMin[1];
    st->vp_max[0] = c->vpMax[0]; st->vp_max[1] = c->vpMax[1];

    if (texturing) {
        pfcu_texture *dev = NULL;
        if (tex->surf) {                       /* sampling a render target: alias its colour buffer */
            if (tex->surf == c->cur_surf) {
                /* feedback loop (undefined in any API); sample the state before this batch */
            }
            pfh_upload_if_needed(c, tex->surf);
            if (!tex->surf->as_texture) tex->surf->as_texture = pfcu_texture_from_surface(tex->surf->dev);
            dev = tex->surf->as_texture;
        } else {
            if (!tex->dev) {
                int code = pfh_texture_code(tex->format, tex->type);
                if (code >= 0 && tex->pixels) tex->dev = pfcu_texture_create(tex->pixels, tex->w, tex->h, code);
                if (!tex->dev) {
                    fprintf(stderr, "pixelforge-b200: texture format %d/%d is not supported by the CUDA sampler\n",
                            (int)tex->format, (int)tex->type);
                    c->errCode = PF_INVALID_OPERATION;
                }
            }
            dev = tex->dev;
        }
        if (dev) {
            st->flags |= PFCU_ST_TEXTURE;
            st->texture = dev;
            st->tex_filter = (uint8_t)tex->filter;
            st->tex_wrap = (uint8_t)tex->wrap;
        }
    }
    if (phong) {
        st->flags |= PFCU_ST_PHONG;
        uint32_t n = 0;
        for (int i = c->activeHead; i >= 0 && n < 8; i = c->lights[i].next) {
            const pf_light *l = &c->lights[i];
            pfcu_light *d = &st->lights[n++];
            memcpy(d->position, l->position, 12); memcpy(d->direction, l->direction, 12);
            d->inner_cutoff = l->innerCutOff; d->outer_cutoff = l->outerCutOff;
            d->att_constant = l->attConstant; d->att_linear = l->attLinear; d->att_quadratic = l->attQuadratic;
            d->ambient = color_dword(l->ambient); d->diffuse = color_dword(l->diffuse); d->specular = color_dword(l->specular);
        }
        st->n_lights = n;
        fill_material(&st->material[0], &c->material[0]);
        fill_material(&st->material[1], &c->material[1]);
        memcpy(st->view_pos, c->viewPos, 12);
    }
}

void pfh_snapshot_state(pf_ctx *c, pfcu_state *st) { build_state(c, st); }

static uint32_t current_state_index(pf_ctx *c)
{
    if (!c->state_dirty && c->n_states > 0) return c->n_states - 1;
    pfcu_state st;
    build_state(c, &st);
    c->state_dirty = 0;
    if (c->n_states > 0 && memcmp(&c->states[c->n_states - 1], &st, sizeof st) == 0) return c->n_states - 1;
    if (c->n_states == c->state_cap) {
        uint32_t ncap = c->state_cap ? c->state_cap * 2 : 16;
        pfcu_state *p = (pfcu_state *)realloc(c->states, (size_t)ncap * sizeof *p);
        if (!p) { c->errCode = PF_ERROR_OUT_OF_MEMORY; return 0; }
        c->states = p; c->state_cap = ncap;
    }
    c->states[c->n_states] = st;
    return c->n_states++;
}

/* ---- batch ----------------------------------------------------------------------------------- */

#define PFH_BATCH_TRIS (1u << 18)       /* 262,144 triangles = 38 MiB per pinned buffer */

/* Pinned double buffer for the triangle batch.  It starts small (a context that draws a few hundred
 * triangles per frame should not hold 80 MB of page-locked memory - BASELINE config C5 runs 256 contexts)
 * and grows 4x each time a batch fills up, to PFH_BATCH_TRIS at most. */
#define PFH_BATCH_TRIS_MIN 4096u

static uint32_t batch_limit(void)
{
    static uint32_t lim = 0;
    if (!lim) {
        lim = PFH_BATCH_TRIS;
        const char *e = getenv("PF_CUDA_BATCH_TRIS");
        if (e && atoi(e) > 0) lim = (uint32_t)atoi(e);
    }
    return lim;
}

static int alloc_batch(pf_ctx *c, uint32_t cap)
{
    pfcu_triangle *nb[2];
    for (int i = 0; i < 2; i++) {
        nb[i] = (pfcu_triangle *)pfcu_host_alloc((size_t)cap * sizeof(pfcu_triangle));
        if (!nb[i]) { if (i) pfcu_host_free(nb[0]); c->errCode = PF_ERROR_OUT_OF_MEMORY; return 0; }
    }
    for (int i = 0; i < 2; i++) {
        if (c->tris[i]) { pfcu_host_wait(c->tris[i]); pfcu_host_free(c->tris[i]); }
        c->tris[i] = nb[i];
    }
    c->tri_cap = cap; c->cur_buf = 0; c->n_tris = 0;
    return 1;
}

static int ensure_batch(pf_ctx *c)
{
    if (c->n_prims || c->n_segs) pfh_flush(c);  /* points / lines / list replays submitted before this triangle come first */
    if (c->clear_pending) pfh_issue_pending_clear(c);
    if (c->tris[0]) return 1;
    uint32_t cap = PFH_BATCH_TRIS_MIN;
    if (cap > batch_limit()) cap = batch_limit();
    return alloc_batch(c, cap);
}

void pfh_upload_if_needed(pf_ctx *c, pf_surf *s)
{
    (void)c;
    if (s && s->host_newer) {
        pfcu_surface_upload(s->dev, s->tex->pixels, s->zhost, 0, s->tex->h);
        s->host_newer = 0;
    }
}

static void capture_append(pf_ctx *c)
{
    if (c->cap_ntris + c->n_tris > c->cap_tris_cap) {
        size_t nc = c->cap_tris_cap ? c->cap_tris_cap * 2 : 4096;
        while (nc < c->cap_ntris + c->n_tris) nc *= 2;
        pfcu_triangle *p = (pfcu_triangle *)realloc(c->cap_tris, nc * sizeof *p);
        if (!p) { c->errCode = PF_ERROR_OUT_OF_MEMORY; return; }
        c->cap_tris = p; c->cap_tris_cap = nc;
    }
    if (c->cap_nstates + c->n_states > c->cap_states_cap) {
        size_t nc = c->cap_states_cap ? c->cap_states_cap * 2 : 16;
        while (nc < c->cap_nstates + c->n_states) nc *= 2;
        pfcu_state *p = (pfcu_state *)realloc(c->cap_states, nc * sizeof *p);
        if (!p) { c->errCode = PF_ERROR_OUT_OF_MEMORY; return; }
        c->cap_states = p; c->cap_states_cap = nc;
    }
    memcpy(c->cap_states + c->cap_nstates, c->states, c->n_states * sizeof(pfcu_state));
    const pfcu_triangle *src = c->tris[c->cur_buf];
    pfcu_triangle *dst = c->cap_tris + c->cap_ntris;
    for (uint32_t i = 0; i < c->n_tris; i++) { dst[i] = src[i]; dst[i].state += (uint32_t)c->cap_nstates; }
    c->cap_ntris += c->n_tris; c->cap_nstates += c->n_states;
}

static void flush_prims(pf_ctx *c)
{
    pf_surf *s = c->cur_surf;
    pfh_upload_if_needed(c, s);
    int rc = pfcu_submit_prims(s->dev, c->prims, c->n_prims);
    if (rc != PFCU_OK) {
        fprintf(stderr, "pixelforge-b200: pfcu_submit_prims failed (%d): %s\n", rc, pfcu_last_error());
        c->errCode = (rc == PFCU_ERR_OOM) ? PF_ERROR_OUT_OF_MEMORY : PF_INVALID_OPERATION;
    }
    s->dev_newer = 1; s->dirty_y0 = 0; s->dirty_y1 = s->tex->h; s->readback_queued = 0;
    c->n_prims = 0;
}

void pfh_flush(pf_ctx *c)
{
    if (!c) return;
    /* replays of device-resident lists (and a pfClear waiting with them) of this and every other context of the thread */
    if (c->n_segs || pfh_lists_pending() > (c->registered ? 1 : 0)) pfh_lists_flush_all(c);
    else if (c->clear_pending) pfh_issue_pending_clear(c);
    if (c && c->n_prims) flush_prims(c);
    if (!c || c->n_tris == 0) { if (c) { c->n_states = 0; c->state_dirty = 1; c->n_vparams = 0; } return; }
    pf_surf *s = c->cur_surf;
    if (c->capturing) capture_append(c);
    pfh_upload_if_needed(c, s);
    int rc;
    if (c->batch_raw) {
        uint32_t produced = 0;
        rc = pfcu_submit_raw(s->dev, c->states, c->n_states, c->vparams, c->n_vparams, c->pow_tables, c->n_pow,
                             (const pfcu_rawtri *)c->tris[c->cur_buf], c->n_tris, &produced);
        c->tris_emitted += produced;
    } else rc = pfcu_submit(s->dev, c->states, c->n_states, c->tris[c->cur_buf], c->n_tris);
    c->n_vparams = 0;
    if (rc != PFCU_OK) {
        fprintf(stderr, "pixelforge-b200: pfcu_submit failed (%d): %s\n", rc, pfcu_last_error());
        c->errCode = (rc == PFCU_ERR_OOM) ? PF_ERROR_OUT_OF_MEMORY : PF_INVALID_OPERATION;
    }
    s->dev_newer = 1; s->readback_queued = 0;
    c->cur_buf ^= 1;
    pfcu_host_wait(c->tris[c->cur_buf]);    /* the other buffer may still be in flight */
    c->n_tris = 0;
    c->n_states = 0;
    c->state_dirty = 1;
}

/* PF_CUDA_SYNC=explicit, leaving a context for another one (a batch of contexts drawn in turn, then presented):
 * queue the read-back of what this context drew right behind its kernels, so that it overlaps with the drawing of
 * the next contexts instead of being paid, one blocking copy after the other, when they are presented.  Only for
 * page-locked mirrors (the copy is then a plain DMA); drawing to the surface again simply invalidates it. */
static int g_queue_readback = 1;
void pfxEnableQueuedReadback(PFboolean on) { g_queue_readback = on ? 1 : 0; }

void pfh_queue_readback(pf_ctx *c, pf_surf *s)
{
    if (!g_queue_readback || !s || !s->dev_newer || s->readback_queued || !s->pinned_color) return;
    if (s->zhost && !s->pinned_depth) return;
    PFuint y0 = s->dirty_y0, y1 = s->dirty_y1;
    if (y1 > s->tex->h) y1 = s->tex->h;
    if (y0 >= y1) return;
    (void)c;
    if (pfcu_surface_download_async(s->dev, s->tex->pixels, s->zhost, y0, y1 - y0) == PFCU_OK) s->readback_queued = 1;
}

void pfh_sync_surface(pf_ctx *c, pf_surf *s)
{
    if (c && c->cur_surf == s) pfh_flush(c);
    if (pfh_lists_pending()) pfh_lists_flush_all(c);       /* another context of this thread may hold work for s */
    if (!s) return;
    if (s->dev_newer) {
        PFuint y0 = s->dirty_y0, y1 = s->dirty_y1;
        if (y1 > s->tex->h) y1 = s->tex->h;
        if (s->readback_queued) pfcu_surface_wait(s->dev);
        else if (y0 < y1) pfcu_surface_download(s->dev, s->tex->pixels, s->zhost, y0, y1 - y0);
        else pfcu_finish();
        s->readback_queued = 0;
        s->dev_newer = 0;
        s->dirty_y0 = s->tex->h; s->dirty_y1 = 0;
    }
}

void pfh_end_of_draw(pf_ctx *c)
{
    if (pfh_sync_mode_explicit() || c->replaying || c->recording) return;
    pfh_sync_surface(c, c->cur_surf);
}

static inline void emit_triangle(pf_ctx *c, int face, int is3d, const pf_vertex *a, const pf_vertex *b, const pf_vertex *d)
{
    if (!ensure_batch(c)) return;
    uint32_t sidx = current_state_index(c);
    pfcu_triangle *t = &c->tris[c->cur_buf][c->n_tris];
    pfv_emit(t, a, b, d, sidx, face, is3d);

    /* dirty rows for the next host-mirror refresh */
    float ymin = a->screen[1], ymax = a->screen[1];
    if (b->screen[1] < ymin) ymin = b->screen[1];
    if (d->screen[1] < ymin) ymin = d->screen[1];
    if (b->screen[1] > ymax) ymax = b->screen[1];
    if (d->screen[1] > ymax) ymax = d->screen[1];
    pf_surf *s = c->cur_surf;
    float h = (float)s->tex->h;
    if (!(ymin > 0.0f)) ymin = 0.0f;            /* also catches NaN */
    if (!(ymax < h)) ymax = h - 1.0f;
    if (ymin < h && ymax >= 0.0f) {
        PFuint y0 = (PFuint)ymin, y1 = (PFuint)ymax + 1u;
        if (y0 < s->dirty_y0) s->dirty_y0 = y0;
        if (y1 > s->dirty_y1) s->dirty_y1 = y1;
    }
    c->tris_emitted++;
    if (++c->n_tris == c->tri_cap) {
        pfh_flush(c);
        if (c->tri_cap < batch_limit()) {           /* the batch filled up: this context draws a lot, grow */
            uint32_t cap = c->tri_cap * 4u;
            alloc_batch(c, cap > batch_limit() ? batch_limit() : cap);
        }
    }
}

/* ---- Gouraud vertex lighting: pfv_light_vertex in pf_vstage.h (shared with the device) --------------- */

/* Specular tables for the device (pfcu.h, PFCU_POW_TABLE_SIZE): T[k-1] = the smallest float x >= 0 with
 * (int)(255 * powf(x, shininess)) >= k, found by bisection over float bit patterns with THIS libm's powf.
 * The device sees x = max(N.H, 0) of two normalised vectors, i.e. x <= 1 + a few ulp; the table is exact up to the
 * input where the result would reach 257 (x > 1.00001 even for shininess 1024), beyond which (PFubyte) wraps again.
 * Valid when that function is non-decreasing in x, which holds comfortably for shininess >= 1 (neighbouring
 * floats move the result by shininess * 2^-24 relative, far above powf's error) and is spot-checked below.
 * Returns the table index, or -1 when the shininess cannot be tabulated (then the host lights the vertices). */
static int spec_of(float x, float shin) { return (int)(255 * powf(fmaxf(x, 0.0f), shin)); }

static int pow_table_index(pf_ctx *c, float shininess)
{
    uint32_t key; memcpy(&key, &shininess, 4);
    for (uint32_t i = 0; i < c->n_pow; i++) { uint32_t k; memcpy(&k, &c->pow_shininess[i], 4); if (k == key) return (int)i; }
    if (!(shininess >= 1.0f && shininess <= 1024.0f)) return -1;
    if (c->n_pow == c->pow_cap) {
        uint32_t nc = c->pow_cap ? c->pow_cap * 2 : 4;
        float *t = (float *)realloc(c->pow_tables, (size_t)nc * PFCU_POW_TABLE_SIZE * sizeof(float));
        if (!t) return -1;
        c->pow_tables = t;
        float *sh = (float *)realloc(c->pow_shininess, (size_t)nc * sizeof(float));
        if (!sh) return -1;
        c->pow_shininess = sh; c->pow_cap = nc;
    }
    float *T = c->pow_tables + (size_t)c->n_pow * PFCU_POW_TABLE_SIZE;
    for (int k = 1; k <= PFCU_POW_TABLE_SIZE; k++) {
        uint32_t lo = 0u, hi;                           /* bit patterns of non-negative floats order like the floats */
        if (spec_of(0.0f, shininess) >= k) { T[k - 1] = 0.0f; continue; }
        {   /* upper end 1.01: 255 * 1.01^s is >= 256 and below 2^31 for every s in [1, 1024] */
            float top = 1.01f; memcpy(&hi, &top, 4);
            if (spec_of(top, shininess) < k) return -1;
        }
        while (hi - lo > 1u) {
            const uint32_t mid = lo + (hi - lo) / 2u;
            float x; memcpy(&x, &mid, 4);
            if (spec_of(x, shininess) >= k) hi = mid; else lo = mid;
        }
        memcpy(&T[k - 1], &hi, 4);
    }
    /* spot check of the step-function model against the real thing */
    uint32_t rs = 0x9e3779b9u ^ key;
    for (int i = 0; i < 4096; i++) {
        rs = rs * 1664525u + 1013904223u;
        const float x = (float)(rs >> 8) * (1.0f / 16777216.0f) * ((i & 15) ? 1.0f : 1.000001f);
        int lo = 0, hi = PFCU_POW_TABLE_SIZE;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (T[mid] <= x) lo = mid + 1; else hi = mid; }
        if ((uint8_t)lo != (uint8_t)spec_of(x, shininess)) return -1;
    }
    for (int k = 1; k < PFCU_POW_TABLE_SIZE; k++) if (T[k] < T[k - 1]) return -1;
    c->pow_shininess[c->n_pow] = shininess;
    return (int)c->n_pow++;
}

/* ---- parameters of the shared vertex stage (pf_vstage.h) ------------------------------------------ */

void pfh_vstage_params(const pf_ctx *c, pfv_params *p)
{
    memcpy(p->mvp, c->matMVP, sizeof p->mvp);
    memcpy(p->normal_mat, c->matNormal, sizeof p->normal_mat);
    p->vp_pos[0] = c->vpPos[0]; p->vp_pos[1] = c->vpPos[1];
    p->vp_dim[0] = c->vpDim[0]; p->vp_dim[1] = c->vpDim[1];
    p->lighting = ((c->state & PF_LIGHTING) && lights_active(c)) ? 1u : 0u;
    p->diffuse[0] = color_dword(c->material[0].diffuse);
    p->diffuse[1] = color_dword(c->material[1].diffuse);
}

int pfxSpecularTableCheck(PFfloat shininess, PFuint samples)
{
    pf_ctx *c = pf_cur;
    if (!c) return -2;
    const int ti = pow_table_index(c, shininess);
    if (ti < 0) return -1;
    const float *T = c->pow_tables + (size_t)ti * PFCU_POW_TABLE_SIZE;
    int bad = 0;
    uint32_t rs = 0x1234567u;
    for (PFuint i = 0; i < samples + 3u * PFCU_POW_TABLE_SIZE; i++) {
        float x;
        if (i < samples) { rs = rs * 1664525u + 1013904223u; x = (float)(rs >> 8) * (1.0f / 16777216.0f) * 1.000001f; }
        else {              /* the threshold itself and its two neighbours */
            uint32_t u; memcpy(&u, &T[(i - samples) / 3u], 4);
            const uint32_t k = (i - samples) % 3u;
            u = k == 0 ? (u ? u - 1u : u) : (k == 2 ? u + 1u : u);
            memcpy(&x, &u, 4);
        }
        int lo = 0, hi = PFCU_POW_TABLE_SIZE;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (T[mid] <= x) lo = mid + 1; else hi = mid; }
        if ((uint8_t)lo != (uint8_t)spec_of(x, shininess)) bad++;
    }
    return bad;
}

static void fill_vparams_lit(pf_ctx *c, pfcu_vparams_lit *e)
{
    memset(e, 0, sizeof *e);
    pfh_vstage_params(c, &e->base);
    e->gouraud = (e->base.lighting && c->lightingMode == PF_GOURAUD) ? 1u : 0u;
    if (e->base.lighting) {
        pfh_update_view_pos(c);
        memcpy(e->view_z, c->matView + 8, 12);
        memcpy(e->view_pos, c->viewPos, 12);
        uint32_t n = 0;
        for (int i = c->activeHead; i >= 0 && n < 8; i = c->lights[i].next) {
            const pf_light *l = &c->lights[i];
            pfcu_light *d = &e->lights[n++];
            memcpy(d->position, l->position, 12); memcpy(d->direction, l->direction, 12);
            d->inner_cutoff = l->innerCutOff; d->outer_cutoff = l->outerCutOff;
            d->att_constant = l->attConstant; d->att_linear = l->attLinear; d->att_quadratic = l->attQuadratic;
            d->ambient = color_dword(l->ambient); d->diffuse = color_dword(l->diffuse); d->specular = color_dword(l->specular);
        }
        e->n_lights = n;
        fill_material(&e->material[0], &c->material[0]);
        fill_material(&e->material[1], &c->material[1]);
        if (e->gouraud)
            for (int f = 0; f < 2; f++) {
                const int t = pow_table_index(c, e->material[f].shininess);
                e->pow_table[f] = t < 0 ? 0xffffffffu : (uint32_t)t;
            }
    }
}

/* The prologue environment of the next triangle: the newest table entry when nothing it reads has changed
 * (API calls bump vp_epoch; materials, which PF_COLOR_MATERIAL changes per vertex, are compared by value). */
static const pfcu_vparams_lit *current_vparams(pf_ctx *c, uint32_t *index)
{
    if (c->n_vparams > 0 && c->vp_epoch_built == c->vp_epoch &&
        memcmp(c->vp_material, c->material, sizeof c->vp_material) == 0) {
        *index = c->n_vparams - 1;
        return &c->vparams[c->n_vparams - 1];
    }
    pfcu_vparams_lit e;
    fill_vparams_lit(c, &e);                 /* may bump vp_epoch (view position) - record the epoch afterwards */
    c->vp_epoch_built = c->vp_epoch;
    memcpy(c->vp_material, c->material, sizeof c->vp_material);
    if (c->n_vparams > 0 && memcmp(&c->vparams[c->n_vparams - 1], &e, sizeof e) == 0) { *index = c->n_vparams - 1; return &c->vparams[*index]; }
    if (c->n_vparams == c->vparams_cap) {
        uint32_t nc = c->vparams_cap ? c->vparams_cap * 2 : 8;
        pfcu_vparams_lit *p = (pfcu_vparams_lit *)realloc(c->vparams, (size_t)nc * sizeof *p);
        if (!p) { c->errCode = PF_ERROR_OUT_OF_MEMORY; *index = 0; return c->n_vparams ? &c->vparams[0] : NULL; }
        c->vparams = p; c->vparams_cap = nc;
    }
    c->vparams[c->n_vparams] = e;
    *index = c->n_vparams++;
    return &c->vparams[*index];
}

void pfh_update_view_pos(pf_ctx *c)
{
    if (c->viewPosValid) return;
    pf_mat4 inv;
    m4_invert(inv, c->matView);
    c->viewPos[0] = inv[12]; c->viewPos[1] = inv[13]; c->viewPos[2] = inv[14];
    c->viewPosValid = 1;
    PFH_VP_TOUCH(c);
    if (c->lightingMode == PF_PHONG) c->state_dirty = 1;
}

/* ---- large vertex-array draws: the vertex stage runs on the device ---------------------------------
 * Eligible: PF_TRIANGLES over tightly packed PF_FLOAT positions (+ optional float normals / texcoords,
 * ubyte colours), filled polygons, no Gouraud lighting (its powf/sqrtf come from the host libm and stay
 * on the host).  Same functions (pf_vstage.h), same order of operations, same triangle order. */
#define PFH_DEVICE_DRAW_MIN_TRIS 1024u

/* Largest index of an index buffer = how many vertices the draw references.  The plain loop costs ~1 ns per index
 * (3 ms for the 1 M-triangle mesh, more than the GPU spends on the whole frame), so 32-bit indices take an AVX2
 * path when the CPU has it. */
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
__attribute__((target("avx2"))) static uint32_t max_index_u32_avx2(const uint32_t *p, size_t n)
{
    __m256i m0 = _mm256_setzero_si256(), m1 = m0, m2 = m0, m3 = m0;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        m0 = _mm256_max_epu32(m0, _mm256_loadu_si256((const __m256i *)(p + i)));
        m1 = _mm256_max_epu32(m1, _mm256_loadu_si256((const __m256i *)(p + i + 8)));
        m2 = _mm256_max_epu32(m2, _mm256_loadu_si256((const __m256i *)(p + i + 16)));
        m3 = _mm256_max_epu32(m3, _mm256_loadu_si256((const __m256i *)(p + i + 24)));
    }
    m0 = _mm256_max_epu32(_mm256_max_epu32(m0, m1), _mm256_max_epu32(m2, m3));
    uint32_t t[8]; _mm256_storeu_si256((__m256i *)t, m0);
    uint32_t mx = 0;
    for (int k = 0; k < 8; k++) if (t[k] > mx) mx = t[k];
    for (; i < n; i++) if (p[i] > mx) mx = p[i];
    return mx;
}
#endif
static uint32_t max_index_u32(const uint32_t *p, size_t n)
{
#if defined(__x86_64__) && defined(__GNUC__)
    if (__builtin_cpu_supports("avx2")) return max_index_u32_avx2(p, n);
#endif
    uint32_t mx = 0;
    for (size_t i = 0; i < n; i++) if (p[i] > mx) mx = p[i];
    return mx;
}

int pfh_device_draw(pf_ctx *c, PFsizei count, PFint first, int indexed, PFdatatype itype, const void *indices,
                    int useNrm, int useTex, int useCol)
{
    static int caps = -1;
    if (caps < 0) caps = (int)pfcu_capabilities();
    if (!(caps & PFCU_CAP_DEVICE_VERTEX) || !c->device_vertex || c->recording || c->capturing) return 0;
    if (count / 3u < PFH_DEVICE_DRAW_MIN_TRIS) return 0;
    if (c->apos.type != PF_FLOAT || !c->apos.buffer) return 0;
    if (useNrm && c->anrm.type != PF_FLOAT) return 0;
    if (useTex && c->atex.type != PF_FLOAT) return 0;
    if (useCol && c->acol.type != PF_UNSIGNED_BYTE) return 0;
    const int lighting = (c->state & PF_LIGHTING) && lights_active(c);
    if (lighting && c->lightingMode == PF_GOURAUD) return 0;

    pfcu_draw d; memset(&d, 0, sizeof d);
    if (c->state & PF_CULL_FACE) { d.n_faces = 1; d.faces[0] = (uint8_t)!c->cullFace; }
    else { d.n_faces = 2; d.faces[0] = PF_FRONT; d.faces[1] = PF_BACK; }
    for (uint32_t f = 0; f < d.n_faces; f++) if (c->polygonMode[d.faces[f]] != PF_FILL) return 0;

    /* vertices referenced by the draw */
    size_t nverts; int unknown = 0;
    if (indexed) {
        size_t mx = 0;
        switch (itype) {
        case PF_UNSIGNED_BYTE:  { const PFubyte *p = (const PFubyte *)indices; for (PFsizei i = 0; i < count; i++) if (p[i] > mx) mx = p[i]; d.index_bytes = 1; } break;
        case PF_UNSIGNED_SHORT: { const PFushort *p = (const PFushort *)indices; for (PFsizei i = 0; i < count; i++) if (p[i] > mx) mx = p[i]; d.index_bytes = 2; } break;
        default:                /* >= 1 M indices: the device scans them once they are uploaded (pfcu_draw.n_vertices == 0); below that the
                                   extra round trip costs more than the AVX2 scan */
                                /* indices in a block declared static: the device scans them once per modification */
                                if (count >= (1u << 20) || pfcu_host_is_static(indices)) unknown = 1; else mx = max_index_u32((const uint32_t *)indices, count);
                                d.index_bytes = 4; break;
        }
        nverts = unknown ? 0 : mx + 1; d.indices = indices; d.first = 0;
    } else { nverts = (size_t)first + count; d.indices = NULL; d.first = (uint32_t)first; }
    if (nverts > 0x7fffffffu) return 0;

    d.positions = (const float *)c->apos.buffer; d.pos_size = (uint32_t)c->apos.size;
    d.normals = useNrm ? (const float *)c->anrm.buffer : NULL;
    d.texcoords = useTex ? (const float *)c->atex.buffer : NULL;
    d.colors = useCol ? (const uint8_t *)c->acol.buffer : NULL; d.color_size = useCol ? (uint32_t)c->acol.size : 0;
    d.n_vertices = (uint32_t)nverts; d.count = count - count % 3u;
    d.current_color = color_dword(c->currentColor);

    pfh_flush(c);                               /* everything submitted before this draw comes first */
    pfh_update_matrices(c, 1);                  /* what pfBegin latches */
    pfv_params vp; pfh_vstage_params(c, &vp);
    if (vp.lighting) pfh_update_view_pos(c);
    pfcu_state st; build_state(c, &st);
    pf_surf *s = c->cur_surf;
    pfh_upload_if_needed(c, s);
    uint32_t produced = 0;
    int rc = pfcu_draw_triangles(s->dev, &st, &vp, &d, &produced);
    if (rc != PFCU_OK) {
        fprintf(stderr, "pixelforge-b200: pfcu_draw_triangles failed (%d): %s\n", rc, pfcu_last_error());
        c->errCode = (rc == PFCU_ERR_OOM) ? PF_ERROR_OUT_OF_MEMORY : PF_INVALID_OPERATION;
    }
    s->dev_newer = 1; s->dirty_y0 = 0; s->dirty_y1 = s->tex->h; s->readback_queued = 0;
    c->tris_emitted += produced;
    c->state_dirty = 1;
    c->currentDrawMode = PF_TRIANGLES; c->vertexCounter = 0;
    return 1;
}

/* ---- one triangle through the vertex stage (triangles.c:62-116) ------------------------------ */

/* Can this triangle go to the device unprocessed?  (pfcu_submit_raw; the oracle build has no such path) */
static int raw_path(pf_ctx *c, const pfcu_vparams_lit *e)
{
    static int caps = -1;
    if (caps < 0) caps = (int)pfcu_capabilities();
    if (!(caps & PFCU_CAP_RAW_TRIANGLES) || !c->device_vertex || c->capturing) return 0;
    if (e->gouraud && (e->pow_table[0] == 0xffffffffu || e->pow_table[1] == 0xffffffffu)) return 0;   /* shininess not tabulable */
    return 1;
}

static void process_triangle(pf_ctx *c, int face, pf_vertex poly[PFH_MAX_POLY_VERTS])
{
    if (c->compiling) {         /* render-list compilation: keep the assembled triangle, unprocessed, tagged with its call */
        if (c->cmp_n == c->cmp_cap) {
            uint32_t nc = c->cmp_cap ? c->cmp_cap * 2 : 1024;
            pfcu_rawtri *q = (pfcu_rawtri *)realloc(c->cmp_tris, (size_t)nc * sizeof *q);
            if (!q) { c->errCode = PF_ERROR_OUT_OF_MEMORY; return; }
            c->cmp_tris = q; c->cmp_cap = nc;
        }
        pfcu_rawtri *t = &c->cmp_tris[c->cmp_n++];
        memset(t, 0, sizeof *t);
        for (int i = 0; i < 3; i++) {
            const pf_vertex *v = &poly[i];
            memcpy(t->v[i].pos, v->position, 16); memcpy(t->v[i].normal, v->normal, 12);
            memcpy(t->v[i].uv, v->texcoord, 8); t->v[i].rgba = v->color;
        }
        t->state = c->cmp_call; t->face = (uint8_t)face;
        return;
    }
    uint32_t vpi = 0;
    const pfcu_vparams_lit *e = current_vparams(c, &vpi);
    if (!e) return;

    if (raw_path(c, e)) {
        if (!c->batch_raw && c->n_tris) {       /* a batch is all raw or all processed: close the processed one */
            pfcu_vparams_lit keep = *e;
            pfh_flush(c);
            c->vparams[0] = keep; c->n_vparams = 1; vpi = 0;
            c->vp_epoch_built = c->vp_epoch; memcpy(c->vp_material, c->material, sizeof c->vp_material);
        }
        c->batch_raw = 1;
        if (!ensure_batch(c)) return;
        const uint32_t sidx = current_state_index(c);
        pfcu_rawtri *t = (pfcu_rawtri *)c->tris[c->cur_buf] + c->n_tris;
        for (int i = 0; i < 3; i++) {
            const pf_vertex *v = &poly[i];
            memcpy(t->v[i].pos, v->position, 16); memcpy(t->v[i].normal, v->normal, 12);
            memcpy(t->v[i].uv, v->texcoord, 8); t->v[i].rgba = v->color;
        }
        t->state = sidx; t->vparams = vpi; t->face = (uint8_t)face; t->pad[0] = t->pad[1] = t->pad[2] = 0;
        pf_surf *s = c->cur_surf;               /* screen coordinates are not known here: everything is dirty */
        s->dirty_y0 = 0; s->dirty_y1 = s->tex->h;
        if (++c->n_tris == c->tri_cap) {
            pfh_flush(c);
            if (c->tri_cap < batch_limit()) {
                uint32_t cap = c->tri_cap * 4u;
                alloc_batch(c, cap > batch_limit() ? batch_limit() : cap);
            }
        }
        return;
    }
    if (c->batch_raw && c->n_tris) {
        pfcu_vparams_lit keep = *e;
        pfh_flush(c);
        c->vparams[0] = keep; c->n_vparams = 1; vpi = 0;
        c->vp_epoch_built = c->vp_epoch; memcpy(c->vp_material, c->material, sizeof c->vp_material);
        e = &c->vparams[0];
    }
    c->batch_raw = 0;

    int n = 3;
    if (e->base.lighting)
        for (int i = 0; i < 3; i++) pfv_prologue_lit(e, NULL, face, &poly[i]);
    int is3d = pfv_project_and_clip(&e->base, poly, &n);
    if (n < 3) return;
    for (int i = 0; i < n - 2; i++) emit_triangle(c, face, is3d, &poly[0], &poly[i + 1], &poly[i + 2]);
}

static void tri_list(pf_ctx *c, int face)
{
    pf_vertex p[PFH_MAX_POLY_VERTS];
    memcpy(p, c->vertexBuffer, 3 * sizeof(pf_vertex));
    process_triangle(c, face, p);
}

static void tri_fan(pf_ctx *c, int face, int count)
{
    for (int i = 0; i < count; i++) {
        pf_vertex p[PFH_MAX_POLY_VERTS];
        p[0] = c->vertexBuffer[0]; p[1] = c->vertexBuffer[i + 1]; p[2] = c->vertexBuffer[i + 2];
        process_triangle(c, face, p);
    }
}

static void tri_strip(pf_ctx *c, int face, int count)
{
    for (int i = 0; i < count; i++) {
        pf_vertex p[PFH_MAX_POLY_VERTS];
        if ((i & 1) == 0) { p[0] = c->vertexBuffer[i]; p[1] = c->vertexBuffer[i + 1]; p[2] = c->vertexBuffer[i + 2]; }
        else { p[0] = c->vertexBuffer[i + 2]; p[1] = c->vertexBuffer[i + 1]; p[2] = c->vertexBuffer[i]; }
        process_triangle(c, face, p);
    }
}

/* ---- points and lines: transform and clip on the host (few primitives), rasterise on the device ------------
 * points.c:62-83, lines.c:137-281.  No lighting, no texturing: the vertex colour is used as is. */

static void emit_prim(pf_ctx *c, const pfcu_prim *p)
{
    if (c->n_tris || c->n_segs) pfh_flush(c);   /* triangles submitted before this primitive come first */
    if (c->clear_pending) pfh_issue_pending_clear(c);
    if (c->n_prims == c->prims_cap) {
        uint32_t nc = c->prims_cap ? c->prims_cap * 2 : 256;
        pfcu_prim *q = (pfcu_prim *)realloc(c->prims, (size_t)nc * sizeof *q);
        if (!q) { c->errCode = PF_ERROR_OUT_OF_MEMORY; return; }
        c->prims = q; c->prims_cap = nc;
    }
    c->prims[c->n_prims++] = *p;
    c->tris_emitted++;
    if (c->n_prims >= 65536u) pfh_flush(c);
}

static void prim_state(const pf_ctx *c, pfcu_prim *p)
{
    p->flags = 0;
    if (c->state & PF_BLEND) { p->flags |= PFCU_ST_BLEND; }
    if (c->state & PF_DEPTH_TEST) { p->flags |= PFCU_ST_DEPTH_TEST; }
    p->blend_mode = (uint8_t)c->blendMode; p->depth_func = (uint8_t)c->depthMode;
}

static void process_point(pf_ctx *c, pf_vertex *v)
{
    pfv_params vp; pfh_vstage_params(c, &vp);
    pfv_transform(&vp, v);
    float *h = v->homogeneous;
    if (h[3] != 1.0f) {
        for (int i = 0; i < 3; i++) if (h[i] < -h[3] || h[i] > h[3]) return;
        const float iw = 1.0f / h[3];
        h[0] *= iw; h[1] *= iw;
    }
    pfv_to_screen(&vp, v);
    if (!(v->screen[0] >= c->vpMin[0] && v->screen[1] >= c->vpMin[1] && v->screen[0] <= c->vpMax[0] && v->screen[1] <= c->vpMax[1])) return;
    pfcu_prim p; memset(&p, 0, sizeof p);
    p.kind = PFP_KIND_POINT; p.x1 = v->screen[0]; p.y1 = v->screen[1]; p.z1 = h[2]; p.c1 = v->color; p.size = c->pointSize;
    prim_state(c, &p);
    emit_prim(c, &p);
}

static int clip_code_2d(const float *scr, PFint xMin, PFint yMin, PFint xMax, PFint yMax)
{
    int code = 0;
    if (scr[0] < xMin) code |= 1;
    if (scr[0] > xMax) code |= 2;
    if (scr[1] < yMin) code |= 8;
    if (scr[1] > yMax) code |= 4;
    return code;
}

static int clip_line_2d(const pf_ctx *c, pf_vertex *v1, pf_vertex *v2)      /* lines.c:137-189 (Cohen-Sutherland, its quirks kept) */
{
    const PFint xMin = c->vpMin[0], yMin = c->vpMin[1], xMax = c->vpMax[0], yMax = c->vpMax[1];
    float m = 0;
    if (v1->screen[0] != v2->screen[0]) m = (v2->screen[1] - v1->screen[1]) / (v2->screen[0] - v1->screen[0]);
    for (;;) {
        int code0 = clip_code_2d(v1->screen, xMin, yMin, xMax, yMax);
        int code1 = clip_code_2d(v2->screen, xMin, yMin, xMax, yMax);
        if ((code0 | code1) == 0) return 1;
        if (code0 & code1) return 0;
        if (code0 == 0) { int ct = code0; code0 = code1; code1 = ct; pf_vertex vt = *v1; *v1 = *v2; *v2 = vt; }
        if (code0 & 1)      { v1->screen[1] += (c->vpMin[0] - v1->screen[0]) * m; v1->screen[0] = (float)c->vpMin[0]; }
        else if (code0 & 2) { v1->screen[1] += (c->vpMax[0] - v1->screen[0]) * m; v1->screen[0] = (float)c->vpMax[0]; }
        else if (code0 & 4) { if (m) v1->screen[0] += (c->vpMin[1] - v1->screen[1]) / m; v1->screen[1] = (float)c->vpMin[1]; }
        else if (code0 & 8) { if (m) v1->screen[0] += (c->vpMax[1] - v1->screen[1]) / m; v1->screen[1] = (float)c->vpMax[1]; }
    }
}

static int clip_coord_3d(float q, float p, float *t1, float *t2)             /* lines.c:115-135 */
{
    if (fabsf(p) < PFH_CLIP_EPSILON) return !(q < -PFH_CLIP_EPSILON);
    const float r = q / p;
    if (p < 0) { if (r > *t2) return 0; if (r > *t1) *t1 = r; }
    else       { if (r < *t1) return 0; if (r < *t2) *t2 = r; }
    return 1;
}

static int clip_line_3d(pf_vertex *v1, pf_vertex *v2)                        /* lines.c:191-227 (Liang-Barsky in clip space) */
{
    float t1 = 0, t2 = 1, d[4];
    float *a = v1->homogeneous, *b = v2->homogeneous;
    for (int i = 0; i < 4; i++) d[i] = b[i] - a[i];
    for (int ax = 0; ax < 3; ax++) {
        if (!clip_coord_3d(a[3] - a[ax], -d[3] + d[ax], &t1, &t2)) return 0;
        if (!clip_coord_3d(a[3] + a[ax], -d[3] - d[ax], &t1, &t2)) return 0;
    }
    if (t2 < 1) for (int i = 0; i < 4; i++) b[i] = a[i] + d[i] * t2;
    if (t1 > 0) for (int i = 0; i < 4; i++) a[i] = a[i] + d[i] * t1;
    return 1;
}

/* returns 0 when the line is clipped away (lines.c:229-268) */
static int process_line(pf_ctx *c, const pf_vertex *in1, const pf_vertex *in2)
{
    pf_vertex l[2] = { *in1, *in2 };
    pfv_params vp; pfh_vstage_params(c, &vp);
    pfv_transform(&vp, &l[0]); pfv_transform(&vp, &l[1]);
    if (l[0].homogeneous[3] == 1.0f && l[1].homogeneous[3] == 1.0f) {
        pfv_to_screen(&vp, &l[0]); pfv_to_screen(&vp, &l[1]);
        if (!clip_line_2d(c, &l[0], &l[1])) return 0;
    } else {
        if (!clip_line_3d(&l[0], &l[1])) return 0;
        for (int i = 0; i < 2; i++) {
            const float iw = 1.0f / l[i].homogeneous[3];
            l[i].homogeneous[0] *= iw; l[i].homogeneous[1] *= iw;
        }
        pfv_to_screen(&vp, &l[0]); pfv_to_screen(&vp, &l[1]);
    }
    pfcu_prim p; memset(&p, 0, sizeof p);
    p.kind = PFP_KIND_LINE;
    p.x1 = l[0].screen[0]; p.y1 = l[0].screen[1]; p.x2 = l[1].screen[0]; p.y2 = l[1].screen[1];
    p.z1 = l[0].homogeneous[2]; p.z2 = l[1].homogeneous[2]; p.c1 = l[0].color; p.c2 = l[1].color; p.size = c->lineWidth;
    prim_state(c, &p);
    emit_prim(c, &p);
    return 1;
}

static void poly_points(pf_ctx *c, int n) { for (int i = 0; i < n; i++) { pf_vertex v = c->vertexBuffer[i]; process_point(c, &v); } }
/* the outline of a triangle / quad in PF_LINE polygon mode stops at the first edge that is clipped away (lines.c:93) */
static void poly_lines(pf_ctx *c, int n) { for (int i = 0; i < n; i++) if (!process_line(c, &c->vertexBuffer[i], &c->vertexBuffer[(i + 1) % n])) return; }

/* draw-mode dispatch (internal/context/context.c:94-244) */
void pfh_process_primitive(pf_ctx *c)
{
    int culled = (c->state & PF_CULL_FACE) != 0;
    int one = culled ? !c->cullFace : PF_FRONT_AND_BACK;     /* face to render */

    switch (c->currentDrawMode) {
    case PF_POINTS: { pf_vertex v = c->vertexBuffer[0]; process_point(c, &v); } break;
    case PF_LINES:  process_line(c, &c->vertexBuffer[0], &c->vertexBuffer[1]); break;
    case PF_TRIANGLES:
    case PF_QUADS: {
        int quad = c->currentDrawMode == PF_QUADS;
        int f0 = (one == PF_FRONT_AND_BACK) ? 0 : one, f1 = (one == PF_FRONT_AND_BACK) ? 1 : one;
        for (int f = f0; f <= f1; f++) {
            if (c->polygonMode[f] == PF_POINT) { poly_points(c, quad ? 4 : 3); continue; }
            if (c->polygonMode[f] == PF_LINE) { poly_lines(c, quad ? 4 : 3); continue; }
            if (quad) tri_fan(c, f, 2); else tri_list(c, f);
        }
    } break;
    case PF_TRIANGLE_FAN:
    case PF_QUAD_FAN: {
        int cnt = c->currentDrawMode == PF_QUAD_FAN ? 4 : 2;
        if (one == PF_FRONT_AND_BACK) { tri_fan(c, PF_FRONT, cnt); tri_fan(c, PF_BACK, cnt); }
        else tri_fan(c, one, cnt);
    } break;
    case PF_TRIANGLE_STRIP:
    case PF_QUAD_STRIP: {
        int cnt = c->currentDrawMode == PF_QUAD_STRIP ? 4 : 2;
        if (one == PF_FRONT_AND_BACK) { tri_strip(c, PF_FRONT, cnt); tri_strip(c, PF_BACK, cnt); }
        else tri_strip(c, one, cnt);
    } break;
    }
}

/* ---- device-resident render lists ---------------------------------------------------------
 * The reference replays a list by re-issuing pfColor4ubv / pfTexCoordfv / pfNormal3fv / pfVertex4fv for every recorded
 * vertex (renderlist.c:71-97).  What those calls compute from the context at replay time is small: the face passes
 * (cull state), the prologue environment (matrices latched by pfBegin, lights, the call's materials - which pfColor
 * feeds when PF_COLOR_MATERIAL is on) and the fragment state; the vertices themselves are the recorded ones as long as
 * the texture matrix is the identity and PF_NORMALIZE is off.  So a list is assembled into triangles ONCE per face mode
 * (with the very code that assembles immediate-mode primitives) and left on the device; a replay queues a reference to
 * it plus one pfcu_list_call per recorded call.  Whatever does not fit (points / lines, polygon modes, a texture
 * matrix, PF_NORMALIZE, per-vertex colours under PF_COLOR_MATERIAL, untabulable shininess, more than 1024 triangles)
 * is replayed through the immediate-mode path as before. */

static PF_CTX_DECL pf_ctx *g_pending[256];
static PF_CTX_DECL int g_npending = 0;

int pfh_lists_pending(void) { return g_npending; }

void pfh_register_pending(pf_ctx *c)
{
    if (c->registered) return;
    if (g_npending == 256) pfh_lists_flush_all(c);
    g_pending[g_npending++] = c; c->registered = 1;
}

int pfh_clear_deferrable(pf_ctx *c)
{
    static int caps = -1;
    if (caps < 0) caps = (int)pfcu_capabilities();
    return (caps & PFCU_CAP_LISTS) && pfh_sync_mode_explicit() && c->device_vertex && !c->capturing && !c->recording && !c->replaying &&
           c->cur_surf->tex->format == PF_RGBA && c->cur_surf->tex->type == PF_UNSIGNED_BYTE;
}

void pfh_issue_pending_clear(pf_ctx *c)
{
    if (!c->clear_pending) return;
    c->clear_pending = 0;
    pfcu_surface_clear_ref(c->cur_surf->dev, 1, c->clear_rgba, 1, c->clear_z);
}

void pfh_list_release_device(pf_list *l)
{
    for (int v = 0; v < 3; v++) { if (l->dev[v]) pfcu_list_destroy(l->dev[v]); l->dev[v] = NULL; l->dev_state[v] = 0; l->dev_tris[v] = 0; }
    l->analysed = 0;
}

void pfh_lists_flush_all(pf_ctx *cur)
{
    if (cur && !cur->registered && (cur->n_segs || cur->clear_pending)) pfh_register_pending(cur);
    if (g_npending == 0) return;
    pfcu_list_job jobs[256];
    int n = 0;
    for (int i = 0; i < g_npending; i++) {
        pf_ctx *x = g_pending[i];
        if (!(x->n_segs || x->clear_pending)) continue;
        pfcu_list_job *j = &jobs[n++];
        memset(j, 0, sizeof *j);
        j->surface = x->cur_surf->dev;
        j->clear = (uint32_t)x->clear_pending; j->clear_rgba = x->clear_rgba; j->clear_depth = x->clear_z;
        pfcu_state dummy_state; pfcu_vparams_lit dummy_vp;
        if (x->n_states == 0) { memset(&dummy_state, 0, sizeof dummy_state); }
        j->states = x->n_states ? x->states : NULL; j->n_states = x->n_states;
        j->vparams = x->vparams; j->n_vparams = x->n_vparams;
        j->pow_tables = x->pow_tables; j->n_pow_tables = x->n_pow;
        j->calls = x->lcalls; j->n_calls = x->n_lcalls;
        j->segments = x->segs; j->n_segments = x->n_segs;
        (void)dummy_vp;
    }
    /* clear-only jobs carry no tables: give them an empty state / environment so that the C-ABI's checks hold */
    static pfcu_state empty_state; static pfcu_vparams_lit empty_vp; static pfcu_list_call empty_call;
    for (int i = 0; i < n; i++) {
        if (!jobs[i].states) { jobs[i].states = &empty_state; jobs[i].n_states = 1; }
        if (!jobs[i].vparams || jobs[i].n_vparams == 0) { jobs[i].vparams = &empty_vp; jobs[i].n_vparams = 1; }
        if (!jobs[i].calls) { jobs[i].calls = &empty_call; jobs[i].n_calls = 1; }
    }
    if (n) {
        int rc = pfcu_submit_list_jobs(jobs, (uint32_t)n);
        if (rc != PFCU_OK) {
            fprintf(stderr, "pixelforge-b200: pfcu_submit_list_jobs failed (%d): %s\n", rc, pfcu_last_error());
            if (cur) cur->errCode = (rc == PFCU_ERR_OOM) ? PF_ERROR_OUT_OF_MEMORY : PF_INVALID_OPERATION;
        }
    }
    const int np = g_npending;
    g_npending = 0;
    for (int i = 0; i < np; i++) {
        pf_ctx *x = g_pending[i];
        const int had = x->n_segs || x->clear_pending;
        x->registered = 0;
        x->n_segs = 0; x->n_lcalls = 0; x->list_tris = 0; x->clear_pending = 0;
        if (x->n_tris == 0) { x->n_states = 0; x->n_vparams = 0; x->state_dirty = 1; }
        if (had) {
            pf_surf *s = x->cur_surf;
            s->dev_newer = 1; s->readback_queued = 0; s->dirty_y0 = 0; s->dirty_y1 = s->tex->h;
            if (x != cur && pfh_sync_mode_explicit()) pfh_queue_readback(x, s);      /* what leaving the context used to start */
        }
    }
}

static int mat_is_identity(const float *m)
{
    for (int i = 0; i < 16; i++) if (m[i] != ((i % 5 == 0) ? 1.0f : 0.0f)) return 0;
    return 1;
}

/* Assemble the list's triangles for the face mode the context is in and put them on the device. */
static void list_compile(pf_ctx *c, pf_list *l, int variant)
{
    pf_vertex saved[6]; memcpy(saved, c->vertexBuffer, sizeof saved);
    const PFsizei saved_count = c->vertexCounter; const PFdrawmode saved_mode = c->currentDrawMode;
    c->compiling = 1; c->cmp_n = 0;
    for (size_t ci = 0; ci < l->size; ci++) {
        const pf_call *k = &l->calls[ci];
        c->cmp_call = (uint32_t)ci;
        c->currentDrawMode = k->mode; c->vertexCounter = 0;
        const PFsizei per = pfh_verts_per_primitive(k->mode);
        for (size_t j = 0; j < k->positions.size; j++) {
            pf_vertex *vx = &c->vertexBuffer[c->vertexCounter++];
            memset(vx, 0, sizeof *vx);
            memcpy(vx->position, k->positions.data + 4 * j, 16);
            memcpy(vx->normal, k->normals.data + 3 * j, 12);
            memcpy(vx->texcoord, k->texcoords.data + 2 * j, 8);
            memcpy(&vx->color, k->colors.data + j, 4);
            if (c->vertexCounter == per) { pfh_process_primitive(c); pfh_carry_over(c); }
        }
    }
    c->compiling = 0;
    memcpy(c->vertexBuffer, saved, sizeof saved); c->vertexCounter = saved_count; c->currentDrawMode = saved_mode;
    l->dev_state[variant] = -1; l->dev_tris[variant] = c->cmp_n;
    if (c->cmp_n == 0 || c->cmp_n > PFCU_LIST_JOB_MAX_TRIS) return;
    l->dev[variant] = pfcu_list_create(c->cmp_tris, c->cmp_n);
    if (l->dev[variant]) l->dev_state[variant] = 1;
}

int pfh_call_list_device(pf_ctx *c, pf_list *l)
{
    static int caps = -1;
    if (caps < 0) caps = (int)pfcu_capabilities();
    if (!(caps & PFCU_CAP_LISTS) || !c->device_vertex || c->capturing || c->recording || c->replaying || l->size == 0) return 0;
    if ((c->state & PF_NORMALIZE) || !mat_is_identity(c->matTexture)) return 0;
    const int culled = (c->state & PF_CULL_FACE) != 0;
    const int variant = culled ? (c->cullFace == PF_BACK ? 0 : 1) : 2;      /* faces drawn: front only, back only, both */
    if ((variant != 1 && c->polygonMode[0] != PF_FILL) || (variant != 0 && c->polygonMode[1] != PF_FILL)) return 0;
    if (!l->analysed) {
        l->analysed = 1; l->colors_uniform = 1;
        for (size_t ci = 0; ci < l->size; ci++) {
            const pf_call *k = &l->calls[ci];
            if (k->mode < PF_TRIANGLES) { for (int v = 0; v < 3; v++) l->dev_state[v] = -1; break; }      /* points / lines: host replay */
            for (size_t j = 1; j < k->colors.size; j++) if (memcmp(k->colors.data + j, k->colors.data, 4) != 0) { l->colors_uniform = 0; break; }
        }
    }
    if ((c->state & PF_COLOR_MATERIAL) && !l->colors_uniform) return 0;
    for (size_t ci = 0; ci < l->size; ci++) {
        const pf_tex *t = (const pf_tex *)l->calls[ci].texture;
        if (t && t->surf && t->surf == c->cur_surf) return 0;               /* sampling its own target: let the ordinary path sort it out */
        if (t && !t->surf) {                                                /* list jobs sample RGBA8 / RGB8 / BGR8 texels only */
            const int code = pfh_texture_code(t->format, t->type);
            if (code == PFCU_TEX_BGRA8 || code >= PFCU_TEX_PIX || code < 0) return 0;
        }
    }
    if (l->dev_state[variant] == 0) list_compile(c, l, variant);
    if (l->dev_state[variant] != 1) return 0;

    pf_surf *s = c->cur_surf;
    if (c->n_tris || c->n_prims) pfh_flush(c);                              /* earlier drawing comes first */
    if (!pfcu_list_job_supported(s->dev, c->list_tris + l->dev_tris[variant], c->n_segs + 1)) {
        if (c->n_segs) pfh_flush(c);
        if (!pfcu_list_job_supported(s->dev, l->dev_tris[variant], 1)) return 0;
    }
    pfh_upload_if_needed(c, s);

    /* what pfCallList does to the context around and inside the replay (renderlist.c:71-97), minus the vertices */
    pf_backup keep = c->backup;
    pf_material m0[2]; memcpy(m0, c->material, sizeof m0);
    float tc0[2], n0[3]; memcpy(tc0, c->currentTexcoord, 8); memcpy(n0, c->currentNormal, 12);
    const PFcolor col0 = c->currentColor; pf_tex *tex0 = c->currentTexture; const PFuint state0 = c->state;
    const uint32_t first_call = c->n_lcalls, states0 = c->n_states, vparams0 = c->n_vparams;
    int ok = 1;
    c->replaying++;
    for (size_t ci = 0; ci < l->size && ok; ci++) {
        const pf_call *k = &l->calls[ci];
        memcpy(c->material, k->material, sizeof c->material);
        c->state_dirty = 1;
        pfBindTexture(k->texture);
        pfh_update_matrices(c, 1);                                          /* pfBegin of a triangle mode */
        c->currentDrawMode = k->mode; c->vertexCounter = 0;
        uint32_t override = 0, rgba = 0;
        if (k->colors.size) {
            PFcolor kc; memcpy(&kc, k->colors.data, 4);
            if (c->state & PF_COLOR_MATERIAL) { pfColor(kc); override = 1; memcpy(&rgba, &c->currentColor, 4); }
            else c->currentColor = kc;
        }
        uint32_t vpi = 0;
        const pfcu_vparams_lit *e = current_vparams(c, &vpi);               /* first: it brings the view position up to date, which a Phong state snapshots */
        if (!e || !raw_path(c, e)) { ok = 0; break; }
        const uint32_t sidx = current_state_index(c);
        if (c->n_lcalls == c->lcalls_cap) {
            uint32_t nc = c->lcalls_cap ? c->lcalls_cap * 2 : 32;
            pfcu_list_call *q = (pfcu_list_call *)realloc(c->lcalls, (size_t)nc * sizeof *q);
            if (!q) { c->errCode = PF_ERROR_OUT_OF_MEMORY; ok = 0; break; }
            c->lcalls = q; c->lcalls_cap = nc;
        }
        pfcu_list_call *lc = &c->lcalls[c->n_lcalls++];
        lc->state = sidx; lc->vparams = vpi; lc->override_color = override; lc->rgba = rgba;
        /* the last recorded vertex leaves its attributes behind (pfTexCoordfv / pfNormal3fv of the replay loop) */
        if (k->positions.size) {
            memcpy(c->currentTexcoord, k->texcoords.data + 2 * (k->positions.size - 1), 8);
            memcpy(c->currentNormal, k->normals.data + 3 * (k->positions.size - 1), 12);
        }
    }
    c->replaying--;
    /* backup_restore of pfCallList: materials, current attributes, texture and enable bits come back */
    memcpy(c->material, m0, sizeof m0); memcpy(c->currentTexcoord, tc0, 8); memcpy(c->currentNormal, n0, 12);
    c->currentColor = col0; c->currentTexture = tex0; c->state = state0; c->state_dirty = 1; c->backup = keep;
    PFH_VP_TOUCH(c);
    if (!ok) {                                                              /* nothing was queued: the tables shrink back, the caller replays on the host */
        c->n_lcalls = first_call;
        if (c->n_segs == 0 && c->n_tris == 0) { c->n_states = states0; c->n_vparams = vparams0; }
        return 0;
    }
    if (c->n_segs == c->segs_cap) {
        uint32_t nc = c->segs_cap ? c->segs_cap * 2 : 16;
        pfcu_list_segment *q = (pfcu_list_segment *)realloc(c->segs, (size_t)nc * sizeof *q);
        if (!q) { c->errCode = PF_ERROR_OUT_OF_MEMORY; c->n_lcalls = first_call; return 0; }
        c->segs = q; c->segs_cap = nc;
    }
    c->segs[c->n_segs].list = l->dev[variant]; c->segs[c->n_segs].first_call = first_call; c->segs[c->n_segs].pad = 0;
    c->n_segs++; c->list_tris += l->dev_tris[variant];
    c->tris_emitted += l->dev_tris[variant];
    s->dirty_y0 = 0; s->dirty_y1 = s->tex->h;
    pfh_register_pending(c);
    return 1;
}
