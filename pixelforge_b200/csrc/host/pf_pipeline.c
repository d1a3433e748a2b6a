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
                int code = pfh_tex_format_code(tex->format, tex->type);
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

void pfh_flush(pf_ctx *c)
{
    if (!c || c->n_tris == 0) { if (c) { c->n_states = 0; c->state_dirty = 1; } return; }
    pf_surf *s = c->cur_surf;
    if (c->capturing) capture_append(c);
    pfh_upload_if_needed(c, s);
    int rc = pfcu_submit(s->dev, c->states, c->n_states, c->tris[c->cur_buf], c->n_tris);
    if (rc != PFCU_OK) {
        fprintf(stderr, "pixelforge-b200: pfcu_submit failed (%d): %s\n", rc, pfcu_last_error());
        c->errCode = (rc == PFCU_ERR_OOM) ? PF_ERROR_OUT_OF_MEMORY : PF_INVALID_OPERATION;
    }
    s->dev_newer = 1;
    c->cur_buf ^= 1;
    pfcu_host_wait(c->tris[c->cur_buf]);    /* the other buffer may still be in flight */
    c->n_tris = 0;
    c->n_states = 0;
    c->state_dirty = 1;
}

void pfh_sync_surface(pf_ctx *c, pf_surf *s)
{
    if (c && c->cur_surf == s) pfh_flush(c);
    if (!s) return;
    if (s->dev_newer) {
        PFuint y0 = s->dirty_y0, y1 = s->dirty_y1;
        if (y1 > s->tex->h) y1 = s->tex->h;
        if (y0 < y1) pfcu_surface_download(s->dev, s->tex->pixels, s->zhost, y0, y1 - y0);
        else pfcu_finish();
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

/* ---- Gouraud vertex lighting: integer Blinn-Phong (lighting.c:23-144) ------------------------ */

static inline PFubyte min255(int n) { return (PFubyte)(n | ((255 - n) >> 31)); }

static PFcolor light_vertex(const pf_ctx *c, const pf_material *m, PFcolor diffuse,
                            const float *viewPos, const float *P, const float *N)
{
    PFubyte R = m->emission.r, G = m->emission.g, B = m->emission.b;
    PFubyte aR = (PFubyte)((m->ambient.r * diffuse.r) / 255);
    PFubyte aG = (PFubyte)((m->ambient.g * diffuse.g) / 255);
    PFubyte aB = (PFubyte)((m->ambient.b * diffuse.b) / 255);

    float V[3], vl2 = 0.0f;
    for (int i = 0; i < 3; i++) { V[i] = viewPos[i] - P[i]; vl2 += V[i] * V[i]; }
    { float il = 1.0f / sqrtf(vl2); for (int i = 0; i < 3; i++) V[i] = V[i] * il; }

    float shininess = m->shininess;
    PFcolor spc = m->specular;

    for (int li = c->activeHead; li >= 0; li = c->lights[li].next) {
        const pf_light *l = &c->lights[li];
        PFubyte lR = 0, lG = 0, lB = 0;
        float L[3] = { l->position[0] - P[0], l->position[1] - P[1], l->position[2] - P[2] };
        float d2 = L[0] * L[0] + L[1] * L[1] + L[2] * L[2];
        float dist = 0.0f;
        if (d2 != 0.0f) {
            dist = sqrtf(d2);
            float il = 1.0f / dist;
            L[0] *= il; L[1] *= il; L[2] *= il;
        }
        PFubyte intensity = 255;
        int skip = 0;
        if (l->innerCutOff < (float)PFH_PI) {
            float nd[3] = { -l->direction[0], -l->direction[1], -l->direction[2] };
            float theta = v3_dot(L, nd);
            float eps = l->innerCutOff - l->outerCutOff;
            int iv = (int)(255 * (theta - l->outerCutOff) / eps);
            intensity = (PFubyte)(iv < 0 ? 0 : (iv > 255 ? 255 : iv));
            if (intensity == 0) skip = 1;
        }
        PFubyte attenuation = 255;
        if (!skip && (l->attLinear || l->attQuadratic)) {
            attenuation = (PFubyte)(255 / (l->attConstant + l->attLinear * dist + l->attQuadratic * d2));
            if (attenuation == 0) skip = 1;
        }
        if (!skip) {
            PFubyte factor = (PFubyte)((intensity * attenuation) / 255);
            int di = (int)(255 * v3_dot(N, L));
            PFubyte diff = (PFubyte)(di > 0 ? di : 0);
            lR = min255(lR + (diffuse.r * l->diffuse.r * diff) / (255 * 255));
            lG = min255(lG + (diffuse.g * l->diffuse.g * diff) / (255 * 255));
            lB = min255(lB + (diffuse.b * l->diffuse.b * diff) / (255 * 255));

            float H[3] = { L[0] + V[0], L[1] + V[1], L[2] + V[2] };
            v3_normalize(H, H);
            PFubyte spec = (PFubyte)(255 * powf(fmaxf(v3_dot(N, H), 0.0f), shininess));
            lR = min255(lR + (spc.r * l->specular.r * spec) / (255 * 255));
            lG = min255(lG + (spc.g * l->specular.g * spec) / (255 * 255));
            lB = min255(lB + (spc.b * l->specular.b * spec) / (255 * 255));

            lR = (PFubyte)((lR * factor) / 255);
            lG = (PFubyte)((lG * factor) / 255);
            lB = (PFubyte)((lB * factor) / 255);
        }
        R = min255(R + lR + (aR * l->ambient.r) / 255);
        G = min255(G + lG + (aG * l->ambient.g) / 255);
        B = min255(B + lB + (aB * l->ambient.b) / 255);
    }
    PFcolor out = { R, G, B, diffuse.a };
    return out;
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

void pfh_update_view_pos(pf_ctx *c)
{
    if (c->viewPosValid) return;
    pf_mat4 inv;
    m4_invert(inv, c->matView);
    c->viewPos[0] = inv[12]; c->viewPos[1] = inv[13]; c->viewPos[2] = inv[14];
    c->viewPosValid = 1;
    if (c->lightingMode == PF_PHONG) c->state_dirty = 1;
}

/* ---- large vertex-array draws: the vertex stage runs on the device ---------------------------------
 * Eligible: PF_TRIANGLES over tightly packed PF_FLOAT positions (+ optional float normals / texcoords,
 * ubyte colours), filled polygons, no Gouraud lighting (its powf/sqrtf come from the host libm and stay
 * on the host).  Same functions (pf_vstage.h), same order of operations, same triangle order. */
#define PFH_DEVICE_DRAW_MIN_TRIS 1024u

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
    size_t nverts;
    if (indexed) {
        size_t mx = 0;
        switch (itype) {
        case PF_UNSIGNED_BYTE:  { const PFubyte *p = (const PFubyte *)indices; for (PFsizei i = 0; i < count; i++) if (p[i] > mx) mx = p[i]; d.index_bytes = 1; } break;
        case PF_UNSIGNED_SHORT: { const PFushort *p = (const PFushort *)indices; for (PFsizei i = 0; i < count; i++) if (p[i] > mx) mx = p[i]; d.index_bytes = 2; } break;
        default:                { const PFuint *p = (const PFuint *)indices; for (PFsizei i = 0; i < count; i++) if (p[i] > mx) mx = p[i]; d.index_bytes = 4; } break;
        }
        nverts = mx + 1; d.indices = indices; d.first = 0;
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
    s->dev_newer = 1; s->dirty_y0 = 0; s->dirty_y1 = s->tex->h;
    c->tris_emitted += produced;
    c->state_dirty = 1;
    c->currentDrawMode = PF_TRIANGLES; c->vertexCounter = 0;
    return 1;
}

/* ---- one triangle through the vertex stage (triangles.c:62-116) ------------------------------ */

static void process_triangle(pf_ctx *c, int face, pf_vertex poly[PFH_MAX_POLY_VERTS])
{
    pfv_params vp;
    pfh_vstage_params(c, &vp);
    int n = 3;

    if (vp.lighting) {
        pfh_update_view_pos(c);
        for (int i = 0; i < 3; i++) {
            pf_vertex *v = &poly[i];
            pfv_prologue(&vp, face, v);
            if (c->lightingMode == PF_GOURAUD) {
                float ndv = v3_dot(v->normal, c->matView + 8);
                PFcolor in, out; memcpy(&in, &v->color, 4);
                out = light_vertex(c, &c->material[(ndv < 0) ? PF_FRONT : PF_BACK], in, c->viewPos, v->position, v->normal);
                memcpy(&v->color, &out, 4);
            }
        }
    }

    int is3d = pfv_project_and_clip(&vp, poly, &n);
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

static void unsupported_primitive(pf_ctx *c, const char *what)
{
    static int warned = 0;
    if (!warned) {
        fprintf(stderr, "pixelforge-b200: %s are outside the CUDA triangle path (SURVEY.md 8-f NEXT-3); ignored\n", what);
        warned = 1;
    }
    c->errCode = PF_INVALID_OPERATION;
}

/* draw-mode dispatch (internal/context/context.c:94-244) */
void pfh_process_primitive(pf_ctx *c)
{
    int culled = (c->state & PF_CULL_FACE) != 0;
    int one = culled ? !c->cullFace : PF_FRONT_AND_BACK;     /* face to render */

    switch (c->currentDrawMode) {
    case PF_POINTS: unsupported_primitive(c, "PF_POINTS"); break;
    case PF_LINES:  unsupported_primitive(c, "PF_LINES"); break;
    case PF_TRIANGLES:
    case PF_QUADS: {
        int quad = c->currentDrawMode == PF_QUADS;
        int f0 = (one == PF_FRONT_AND_BACK) ? 0 : one, f1 = (one == PF_FRONT_AND_BACK) ? 1 : one;
        for (int f = f0; f <= f1; f++) {
            if (c->polygonMode[f] != PF_FILL) { unsupported_primitive(c, "PF_POINT / PF_LINE polygon modes"); continue; }
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
