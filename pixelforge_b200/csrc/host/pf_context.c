/*
 * pf_context.c - the pixelforge.h state machine (context, matrices, lights, materials, immediate
 * mode, vertex arrays, clears and the host-side full-surface operations).
 *
 * API semantics follow the reference's src/context.c and src/getter.c (line references at each
 * block); the implementation is a batching front end: nothing is rasterised here, completed
 * primitives go to pf_pipeline.c, which feeds the pfcu C-ABI.
 */
#include "pf_internal.h"
#include "pf_math.h"
#include "pfx.h"
#include "../pf_pixfmt.h"

#include <float.h>
#include <math.h>
#include <stdio.h>

PF_CTX_DECL pf_ctx *pf_cur = NULL;

void pfh_list_begin(pf_ctx *c, PFdrawmode mode);
void pfh_list_vertex(pf_ctx *c, const PFfloat *v);

#define CTX pf_ctx *c = pf_cur

static const PFcolor WHITE = { 255, 255, 255, 255 };

/* ---- context life cycle (context.c:115-351) --------------------------------------------------- */

PFcontext pfCreateContext(void *targetBuffer, PFsizei width, PFsizei height, PFpixelformat format, PFdatatype type)
{
    pf_ctx *c = (pf_ctx *)PF_CALLOC(1, sizeof *c);
    if (!c) return NULL;

    pf_ctx *saved = pf_cur;
    pf_cur = NULL;                          /* pfGenTexture must not report into another context */
    pf_tex *tex = (pf_tex *)pfGenTexture(targetBuffer, width, height, format, type);
    pf_cur = saved;
    if (!tex) { PF_FREE(c); return NULL; }
    c->main_surf = pfh_surf_create(tex, NULL, 0);
    if (!c->main_surf) { PF_FREE(tex); PF_FREE(c); return NULL; }
    c->mainFramebuffer.texture = tex;
    c->mainFramebuffer.zbuffer = NULL;      /* the main depth buffer lives on the device */
    c->cur_surf = c->main_surf;

    c->vpPos[0] = c->vpMin[0] = 0; c->vpPos[1] = c->vpMin[1] = 0;
    c->vpDim[0] = width - 1;  c->vpMax[0] = (PFint)(width - 1);
    c->vpDim[1] = height - 1; c->vpMax[1] = (PFint)(height - 1);

    c->currentDrawMode = PF_POINTS;
    c->blendMode = PF_BLEND_ALPHA; c->depthMode = PF_LESS;
    c->clearColor.a = 255; c->clearDepth = FLT_MAX;
    c->pointSize = 1.0f; c->lineWidth = 1.0f;
    c->polygonMode[0] = c->polygonMode[1] = PF_FILL;
    c->currentColor = WHITE;
    c->rasterPos[3] = 1.0f; c->pixelZoom[0] = c->pixelZoom[1] = 1.0f;

    for (int i = 0; i < PFH_MAX_LIGHTS; i++) {
        pf_light *l = &c->lights[i];
        l->innerCutOff = l->outerCutOff = (float)PFH_PI;
        l->attConstant = 1.0f;
        l->ambient = (PFcolor){ 51, 51, 51, 255 };
        l->diffuse = WHITE; l->specular = WHITE;
        l->next = -1;
    }
    c->activeHead = -1;
    c->fog.mode = PF_LINEAR; c->fog.density = 1.0f; c->fog.start = 0.0f; c->fog.end = 1.0f;

    for (int f = 0; f < 2; f++) {
        c->material[f].ambient = c->material[f].diffuse = c->material[f].specular = WHITE;
        c->material[f].emission = (PFcolor){ 0, 0, 0, 255 };
        c->material[f].shininess = 64.0f;
    }
    c->cmFace = PF_FRONT_AND_BACK; c->cmMode = PF_AMBIENT_AND_DIFFUSE;

    c->matrixMode = PF_MODELVIEW; c->currentMatrix = c->matView;
    m4_ortho(c->matProjection, -1.0f, 1.0f, -1.0f, 1.0f, -1.0f, 1.0f);
    m4_identity(c->matTexture); m4_identity(c->matNormal); m4_identity(c->matModel); m4_identity(c->matView);

    c->state |= PF_CULL_FACE;
    c->shadingMode = PF_SMOOTH; c->lightingMode = PF_GOURAUD; c->cullFace = PF_BACK;
    c->errCode = PF_NO_ERROR;
    c->state_dirty = 1;
    { const char *e = getenv("PF_CUDA_DEVICE_VERTEX"); c->device_vertex = !(e && e[0] == '0'); }
    return c;
}

void pfDeleteContext(PFcontext ctx)
{
    pf_ctx *c = (pf_ctx *)ctx;
    if (!c) return;
    pfh_sync_surface(c, c->main_surf);
    if (c->cur_surf != c->main_surf) pfh_sync_surface(c, c->cur_surf);
    pfcu_finish();
    for (int i = 0; i < 2; i++) if (c->tris[i]) pfcu_host_free(c->tris[i]);
    free(c->states); free(c->cap_tris); free(c->cap_states);
    free(c->vparams); free(c->pow_tables); free(c->pow_shininess); free(c->prims);
    free(c->segs); free(c->lcalls); free(c->cmp_tris);
    pf_tex *tex = (pf_tex *)c->mainFramebuffer.texture;
    pfh_surf_destroy(c->main_surf);
    PF_FREE(tex);
    if (pf_cur == c) pf_cur = NULL;
    PF_FREE(c);
}

void pfSetMainBuffer(void *targetBuffer, PFsizei width, PFsizei height, PFpixelformat format, PFdatatype type)
{
    CTX;
    if (!targetBuffer || width == 0 || height == 0) { c->errCode = PF_INVALID_VALUE; return; }
    pf_tex *old = (pf_tex *)c->mainFramebuffer.texture;
    /* The reference never touches the outgoing buffer here, and applications commonly free or realloc it before this
       call (window resize): submit what is pending, wait for a read-back that may already be in flight into it, and
       drop the rest of the device -> host debt instead of paying it into memory that may be gone. */
    if (c->cur_surf == c->main_surf) pfh_flush(c);
    if (c->main_surf->readback_queued) pfcu_surface_wait(c->main_surf->dev);
    c->main_surf->readback_queued = 0; c->main_surf->dev_newer = 0;
    c->main_surf->dirty_y0 = old->h; c->main_surf->dirty_y1 = 0;
    if (old->w == width && old->h == height && old->format == format && old->type == type) {
        if (c->main_surf->pinned_color) { pfcu_host_unregister(c->main_surf->pinned_color); c->main_surf->pinned_color = NULL; }
        old->pixels = targetBuffer;             /* same geometry: retarget the host mirror, keep depth */
        c->main_surf->host_newer = 1;
        pfcu_surface_upload(c->main_surf->dev, targetBuffer, NULL, 0, height);
        c->main_surf->host_newer = 0;
    } else {
        pf_ctx *saved = pf_cur; pf_cur = NULL;
        pf_tex *tex = (pf_tex *)pfGenTexture(targetBuffer, width, height, format, type);
        pf_cur = saved;
        pf_surf *ns = tex ? pfh_surf_create(tex, NULL, 0) : NULL;
        if (!ns) { if (tex) PF_FREE(tex); c->errCode = PF_ERROR_OUT_OF_MEMORY; return; }
        /* new areas of the depth buffer start at clearDepth (context.c:288-305); the old depth
           content is not carried over to the resized device surface */
        pfcu_surface_fill(ns->dev, 0, 0, 1, c->clearDepth);
        int was_cur = (c->cur_surf == c->main_surf);
        pfh_surf_destroy(c->main_surf);
        PF_FREE(old);
        c->main_surf = ns; c->mainFramebuffer.texture = tex;
        if (was_cur) c->cur_surf = ns;
    }
    c->auxFramebuffer = NULL;
}

void pfSetAuxBuffer(void *auxFramebuffer) { pf_cur->auxFramebuffer = auxFramebuffer; }

void pfSwapBuffers(void)
{
    CTX;
    if (!c->auxFramebuffer) { c->errCode = PF_INVALID_OPERATION; return; }
    pf_surf *s = c->cur_surf;
    pfh_sync_surface(c, s);                 /* finished frame -> old front buffer */
    if (s->pinned_color) { pfcu_host_unregister(s->pinned_color); s->pinned_color = NULL; }
    void *tmp = s->tex->pixels;
    s->tex->pixels = c->auxFramebuffer;
    c->auxFramebuffer = tmp;
    s->host_newer = 1;                      /* drawing continues on top of the other buffer's content */
}

PFcontext pfGetCurrentContext(void) { return pf_cur; }

void pfMakeCurrent(PFcontext ctx)
{
    if (pf_cur && pf_cur != ctx) {
        if (pfh_sync_mode_explicit()) {
            /* a context whose pending work is a clear and / or replays of device-resident lists keeps it: the work of
               all such contexts of this thread goes out together (one multi-surface submission) at the next flush */
            if (pf_cur->n_tris == 0 && pf_cur->n_prims == 0 && (pf_cur->n_segs || pf_cur->clear_pending)) {
                pfh_register_pending(pf_cur);
                /* ... in chunks: the first contexts of a long round are rendered and on their way back over PCIe (the
                   read-backs are what a many-context frame waits for) while the application still draws the next ones */
                static int chunk = 0;
                if (!chunk) { const char *e = getenv("PF_CUDA_LIST_CHUNK"); chunk = e && atoi(e) > 0 ? atoi(e) : 32; }      /* measured on C5: below ~16 the per-submission cost outweighs the earlier start */
                if (pfh_lists_pending() >= chunk) pfh_lists_flush_all(NULL);
            }
            else { pfh_flush(pf_cur); pfh_queue_readback(pf_cur, pf_cur->cur_surf); }
        } else pfh_sync_surface(pf_cur, pf_cur->cur_surf);
    }
    pf_cur = (pf_ctx *)ctx;
}

PFboolean pfIsEnabled(PFstate state) { return (pf_cur->state & state) != 0; }

static void retarget(pf_ctx *c, pf_surf *s)
{
    if (!s) s = c->main_surf;
    if (s != c->cur_surf) { pfh_flush(c); c->cur_surf = s; }
}

static pf_surf *bound_surface(pf_ctx *c)
{
    if (!c->bindedFramebuffer) return c->main_surf;
    pf_surf *s = pfh_surf_lookup(c->bindedFramebuffer->texture);
    return s ? s : c->main_surf;
}

void pfEnable(PFstate state)
{
    CTX;
    c->state |= state; c->state_dirty = 1; PFH_VP_TOUCH(c);
    if (state & PF_FRAMEBUFFER) retarget(c, bound_surface(c));
}

void pfDisable(PFstate state)
{
    CTX;
    c->state &= ~state; c->state_dirty = 1; PFH_VP_TOUCH(c);
    if (state & PF_FRAMEBUFFER) retarget(c, c->main_surf);
}

PFerrcode pfGetError(void) { CTX; PFerrcode e = c->errCode; c->errCode = PF_NO_ERROR; return e; }

/* ---- matrices (context.c:395-548) -------------------------------------------------------------- */

static void matrix_touched(pf_ctx *c) { c->viewPosValid = 0; PFH_VP_TOUCH(c); }

void pfMatrixMode(PFmatrixmode mode)
{
    CTX;
    switch (mode) {
    case PF_PROJECTION: c->currentMatrix = c->matProjection; break;
    case PF_MODELVIEW:  c->currentMatrix = c->modelMatrixUsed ? c->matModel : c->matView; break;
    case PF_TEXTURE:    c->currentMatrix = c->matTexture; break;
    default: c->errCode = PF_INVALID_ENUM; return;
    }
    c->matrixMode = mode;
}

void pfPushMatrix(void)
{
    CTX;
    switch (c->matrixMode) {
    case PF_PROJECTION:
        if (c->nProjection >= PFH_PROJECTION_STACK) { c->errCode = PF_STACK_OVERFLOW; return; }
        m4_copy(c->stackProjection[c->nProjection++], c->matProjection);
        break;
    case PF_MODELVIEW:
        if (c->nModelview >= PFH_MODELVIEW_STACK) { c->errCode = PF_STACK_OVERFLOW; return; }
        if (c->modelMatrixUsed) m4_copy(c->stackModelview[c->nModelview++], c->matModel);
        else { c->currentMatrix = c->matModel; c->modelMatrixUsed = 1; }   /* first push splits model from view */
        break;
    case PF_TEXTURE:
        if (c->nTexture >= PFH_TEXTURE_STACK) { c->errCode = PF_STACK_OVERFLOW; return; }
        m4_copy(c->stackTexture[c->nTexture++], c->matTexture);
        break;
    }
}

void pfPopMatrix(void)
{
    CTX;
    matrix_touched(c);
    switch (c->matrixMode) {
    case PF_PROJECTION:
        if (c->nProjection == 0) { c->errCode = PF_STACK_UNDERFLOW; return; }
        m4_copy(c->matProjection, c->stackProjection[--c->nProjection]);
        break;
    case PF_MODELVIEW:
        if (c->nModelview == 0) {
            if (!c->modelMatrixUsed) { c->errCode = PF_STACK_UNDERFLOW; return; }
            m4_identity(c->matModel);
            c->modelMatrixUsed = 0;
            c->currentMatrix = c->matView;
        } else m4_copy(c->matModel, c->stackModelview[--c->nModelview]);
        break;
    case PF_TEXTURE:
        if (c->nTexture == 0) { c->errCode = PF_STACK_UNDERFLOW; return; }
        m4_copy(c->matTexture, c->stackTexture[--c->nTexture]);
        break;
    }
}

void pfLoadIdentity(void) { CTX; matrix_touched(c); m4_identity(c->currentMatrix); }

void pfTranslatef(PFfloat x, PFfloat y, PFfloat z)
{
    CTX; pf_mat4 t; matrix_touched(c);
    m4_translate(t, x, y, z);
    m4_mul(c->currentMatrix, t, c->currentMatrix);      /* pre-multiplied, as the reference */
}

void pfRotatef(PFfloat angle, PFfloat x, PFfloat y, PFfloat z)
{
    CTX; pf_mat4 r; matrix_touched(c);
    m4_rotate(r, x, y, z, (float)(angle * PFH_DEG2RAD));
    m4_mul(c->currentMatrix, r, c->currentMatrix);
}

void pfScalef(PFfloat x, PFfloat y, PFfloat z)
{
    CTX; pf_mat4 s; matrix_touched(c);
    m4_scale(s, x, y, z);
    m4_mul(c->currentMatrix, s, c->currentMatrix);
}

void pfMultMatrixf(const PFfloat *mat) { CTX; matrix_touched(c); m4_mul(c->currentMatrix, c->currentMatrix, mat); }

void pfFrustum(PFfloat l, PFfloat r, PFfloat b, PFfloat t, PFfloat n, PFfloat f)
{
    CTX; pf_mat4 m; matrix_touched(c);
    m4_frustum(m, l, r, b, t, n, f);
    m4_mul(c->currentMatrix, c->currentMatrix, m);
}

void pfOrtho(PFfloat l, PFfloat r, PFfloat b, PFfloat t, PFfloat n, PFfloat f)
{
    CTX; pf_mat4 m; matrix_touched(c);
    m4_ortho(m, l, r, b, t, n, f);
    m4_mul(c->currentMatrix, c->currentMatrix, m);
}

/* ---- render state (context.c:553-797) ----------------------------------------------------------- */

void pfViewport(PFint x, PFint y, PFsizei width, PFsizei height)
{ if (pf_cur) PFH_VP_TOUCH(pf_cur);
    CTX;
    if (x <= -(PFint)width || y <= -(PFint)height) { c->errCode = PF_INVALID_OPERATION; return; }
    const pf_tex *mt = (const pf_tex *)c->mainFramebuffer.texture;
    c->vpPos[0] = x; c->vpPos[1] = y;
    c->vpDim[0] = width - 1; c->vpDim[1] = height - 1;
    c->vpMin[0] = PF_MAX(x, 0); c->vpMin[1] = PF_MAX(y, 0);
    /* unsigned arithmetic, as in the reference (PFsizei operands promote the sum) */
    c->vpMax[0] = (PFint)PF_MIN((PFsizei)x + width, mt->w - 1);
    c->vpMax[1] = (PFint)PF_MIN((PFsizei)y + height, mt->h - 1);
    c->state_dirty = 1;
}

void pfPolygonMode(PFface face, PFpolygonmode mode)
{
    CTX;
    if (mode > PF_FILL) { c->errCode = PF_INVALID_ENUM; return; }
    switch (face) {
    case PF_FRONT: c->polygonMode[0] = mode; break;
    case PF_BACK:  c->polygonMode[1] = mode; break;
    case PF_FRONT_AND_BACK: c->polygonMode[0] = c->polygonMode[1] = mode; break;
    default: c->errCode = PF_INVALID_ENUM; break;
    }
}

void pfShadeModel(PFshademode mode) { CTX; c->shadingMode = mode; c->state_dirty = 1; }
void pfLightModel(PFlightmode mode) { if (pf_cur) PFH_VP_TOUCH(pf_cur); CTX; c->lightingMode = mode; c->state_dirty = 1; }
void pfLineWidth(PFfloat w) { CTX; if (w <= 0.0f) { c->errCode = PF_INVALID_VALUE; return; } c->lineWidth = w; }
void pfPointSize(PFfloat s) { CTX; if (s <= 0.0f) { c->errCode = PF_INVALID_VALUE; return; } c->pointSize = s; }
void pfCullFace(PFface face) { CTX; if (face > PF_BACK) { c->errCode = PF_INVALID_ENUM; return; } c->cullFace = face; }

void pfBlendFunc(PFblendmode mode)
{
    CTX;
    if (mode > PF_BLEND_DARKEN) { c->errCode = PF_INVALID_ENUM; return; }
    c->blendMode = mode; c->state_dirty = 1;
}

void pfDepthFunc(PFdepthmode mode)
{
    CTX;
    if (mode > PF_GEQUAL) { c->errCode = PF_INVALID_ENUM; return; }
    c->depthMode = mode; c->state_dirty = 1;
}

void pfBindFramebuffer(PFframebuffer *framebuffer)
{
    CTX;
    c->bindedFramebuffer = framebuffer;
    if (c->state & PF_FRAMEBUFFER) retarget(c, bound_surface(c));
}

void pfBindTexture(PFtexture texture)
{
    CTX;
    pf_tex *t = (pf_tex *)texture;
    if (t != c->currentTexture) {
        /* sampling a surface that still has triangles pending on it: submit them first */
        if (t && t->surf && t->surf == c->cur_surf) pfh_flush(c);
        c->currentTexture = t; c->state_dirty = 1;
    }
}

void pfClear(PFclearflag flag)
{
    CTX;
    if (!flag) return;
    pf_surf *s = c->cur_surf;
    if (c->n_tris || c->n_prims || c->n_segs) pfh_flush(c);
    pfh_upload_if_needed(c, s);
    uint32_t rgba; memcpy(&rgba, &c->clearColor, 4);
    /* the reference's SIMD build clears BOTH buffers whenever either bit is set and never touches
       pixels 0..7 (context.c:696-713); pfcu_surface_clear_ref reproduces that */
    if (pfh_clear_deferrable(c)) { c->clear_pending = 1; c->clear_rgba = rgba; c->clear_z = c->clearDepth; pfh_register_pending(c); }
    else { c->clear_pending = 0; pfcu_surface_clear_ref(s->dev, 1, rgba, 1, c->clearDepth); }
    s->dev_newer = 1; s->dirty_y0 = 0; s->dirty_y1 = s->tex->h; s->readback_queued = 0;
    pfh_end_of_draw(c);
}

void pfClearDepth(PFfloat depth) { pf_cur->clearDepth = depth; }
void pfClearColor(PFubyte r, PFubyte g, PFubyte b, PFubyte a) { pf_cur->clearColor = (PFcolor){ r, g, b, a }; }

/* ---- lights and materials (context.c:802-1156) -------------------------------------------------- */

void pfEnableLight(PFsizei light)
{ if (pf_cur) PFH_VP_TOUCH(pf_cur);
    CTX;
    if (light >= PFH_MAX_LIGHTS) { c->errCode = PF_INVALID_VALUE; return; }
    int *link = &c->activeHead;
    while (*link >= 0) {
        if (*link == (int)light) { c->errCode = PF_INVALID_OPERATION; return; }   /* already on */
        link = &c->lights[*link].next;
    }
    c->lights[light].next = -1;
    *link = (int)light;                     /* appended: evaluation order = enable order */
    c->state_dirty = 1;
}

void pfDisableLight(PFsizei light)
{ if (pf_cur) PFH_VP_TOUCH(pf_cur);
    CTX;
    if (light >= PFH_MAX_LIGHTS) { c->errCode = PF_INVALID_VALUE; return; }
    for (int *link = &c->activeHead; *link >= 0; link = &c->lights[*link].next) {
        if (*link == (int)light) {
            *link = c->lights[light].next;
            c->lights[light].next = -1;
            c->state_dirty = 1;
            return;
        }
    }
    c->errCode = PF_INVALID_OPERATION;
}

PFboolean pfIsEnabledLight(PFsizei light)
{
    CTX;
    if (light >= PFH_MAX_LIGHTS) { c->errCode = PF_INVALID_VALUE; return PF_FALSE; }
    for (int i = c->activeHead; i >= 0; i = c->lights[i].next) if (i == (int)light) return PF_TRUE;
    return PF_FALSE;
}

static int cutoff_ok(float v) { return (v >= 0 && v <= 90) || v == 180; }

void pfLightf(PFsizei light, PFenum param, PFfloat value)
{ if (pf_cur) PFH_VP_TOUCH(pf_cur);
    CTX;
    if (light >= PFH_MAX_LIGHTS) { c->errCode = PF_STACK_OVERFLOW; return; }
    pf_light *l = &c->lights[light];
    c->state_dirty = 1;
    switch (param) {
    case PF_SPOT_INNER_CUTOFF: if (cutoff_ok(value)) l->innerCutOff = cosf((float)(value * PFH_DEG2RAD)); else c->errCode = PF_INVALID_VALUE; break;
    case PF_SPOT_OUTER_CUTOFF: if (cutoff_ok(value)) l->outerCutOff = cosf((float)(value * PFH_DEG2RAD)); else c->errCode = PF_INVALID_VALUE; break;
    case PF_CONSTANT_ATTENUATION:  l->attConstant = value; break;
    case PF_LINEAR_ATTENUATION:    l->attLinear = value; break;
    case PF_QUADRATIC_ATTENUATION: l->attQuadratic = value; break;
    default: c->errCode = PF_INVALID_ENUM; break;
    }
}

static PFcolor color_from_f3(const PFfloat *v)
{
    PFcolor k = { (PFubyte)(v[0] * 255.0f), (PFubyte)(v[1] * 255.0f), (PFubyte)(v[2] * 255.0f), 255 };
    return k;
}

void pfLightfv(PFsizei light, PFenum param, const void *value)
{ if (pf_cur) PFH_VP_TOUCH(pf_cur);
    CTX;
    if (light >= PFH_MAX_LIGHTS) { c->errCode = PF_STACK_OVERFLOW; return; }
    pf_light *l = &c->lights[light];
    const PFfloat *f = (const PFfloat *)value;
    c->state_dirty = 1;
    switch (param) {
    case PF_POSITION:       memcpy(l->position, value, 12); break;
    case PF_SPOT_DIRECTION: memcpy(l->direction, value, 12); break;
    case PF_SPOT_INNER_CUTOFF: case PF_SPOT_OUTER_CUTOFF: case PF_CONSTANT_ATTENUATION:
    case PF_LINEAR_ATTENUATION: case PF_QUADRATIC_ATTENUATION: pfLightf(light, param, f[0]); break;
    case PF_AMBIENT:  l->ambient = color_from_f3(f); break;
    case PF_DIFFUSE:  l->diffuse = color_from_f3(f); break;
    case PF_SPECULAR: l->specular = color_from_f3(f); break;
    default: c->errCode = PF_INVALID_ENUM; break;
    }
}

static int material_targets(pf_ctx *c, PFface face, pf_material **a, pf_material **b)
{
    switch (face) {
    case PF_FRONT: *a = *b = &c->material[0]; return 1;
    case PF_BACK:  *a = *b = &c->material[1]; return 1;
    case PF_FRONT_AND_BACK: *a = &c->material[0]; *b = &c->material[1]; return 1;
    default: c->errCode = PF_INVALID_ENUM; return 0;
    }
}

static void material_set(pf_ctx *c, PFface face, PFenum param, PFcolor k, float shininess)
{
    pf_material *a, *b;
    if (!material_targets(c, face, &a, &b)) return;
    c->state_dirty = 1;
    switch (param) {
    case PF_AMBIENT:  a->ambient = b->ambient = k; break;
    case PF_DIFFUSE:  a->diffuse = b->diffuse = k; break;
    case PF_SPECULAR: a->specular = b->specular = k; break;
    case PF_EMISSION: a->emission = b->emission = k; break;
    case PF_SHININESS: a->shininess = b->shininess = shininess; break;
    case PF_AMBIENT_AND_DIFFUSE: a->ambient = b->ambient = k; a->diffuse = b->diffuse = k; break;
    default: c->errCode = PF_INVALID_ENUM; break;
    }
}

void pfMaterialf(PFface face, PFenum param, PFfloat value)
{
    PFfloat v3[3] = { value, value, value };
    material_set(pf_cur, face, param, color_from_f3(v3), value);
}

void pfMaterialfv(PFface face, PFenum param, const void *value)
{
    const PFfloat *f = (const PFfloat *)value;
    if (param == PF_SHININESS) material_set(pf_cur, face, param, WHITE, f[0]);
    else material_set(pf_cur, face, param, color_from_f3(f), 0.0f);
}

void pfColorMaterial(PFface face, PFenum mode)
{
    CTX;
    if (face > PF_FRONT_AND_BACK) { c->errCode = PF_INVALID_ENUM; return; }
    if (mode < PF_AMBIENT_AND_DIFFUSE || mode > PF_EMISSION) { c->errCode = PF_INVALID_ENUM; return; }
    c->cmFace = face; c->cmMode = mode;
}

/* ---- immediate mode (context.c:41-74,1580-1913) ------------------------------------------------ */

PFsizei pfh_verts_per_primitive(PFdrawmode m)
{
    switch (m) {
    case PF_POINTS: return 1;  case PF_LINES: return 2;  case PF_TRIANGLES: return 3;
    case PF_TRIANGLE_FAN: case PF_TRIANGLE_STRIP: case PF_QUADS: return 4;
    case PF_QUAD_FAN: case PF_QUAD_STRIP: return 6;
    }
    return 0;
}

/* what survives into the next primitive of a fan/strip (Q16: only v[3], resp. v[4..5]) */
void pfh_carry_over(pf_ctx *c)
{
    switch (c->currentDrawMode) {
    case PF_TRIANGLE_FAN: case PF_TRIANGLE_STRIP:
        c->vertexCounter = 1; c->vertexBuffer[0] = c->vertexBuffer[3]; break;
    case PF_QUAD_FAN: case PF_QUAD_STRIP:
        c->vertexCounter = 2; c->vertexBuffer[0] = c->vertexBuffer[4]; c->vertexBuffer[1] = c->vertexBuffer[5]; break;
    default: c->vertexCounter = 0; break;
    }
}

void pfBegin(PFdrawmode mode)
{
    CTX;
    if (c->recording) { pfh_list_begin(c, mode); return; }
    if (mode > PF_QUAD_STRIP) { c->errCode = PF_INVALID_ENUM; return; }
    pfh_update_matrices(c, !(mode == PF_POINTS || mode == PF_LINES));
    c->currentDrawMode = mode;
    c->vertexCounter = 0;
}

void pfEnd(void)
{
    CTX;
    c->vertexCounter = 0;
    if (!c->recording) pfh_end_of_draw(c);
}

void pfVertex4fv(const PFfloat *v)
{
    CTX;
    if (c->recording) { pfh_list_vertex(c, v); return; }
    pf_vertex *vx = &c->vertexBuffer[c->vertexCounter++];
    memcpy(vx->position, v, 16);
    memcpy(vx->normal, c->currentNormal, 12);
    memcpy(vx->texcoord, c->currentTexcoord, 8);
    memcpy(&vx->color, &c->currentColor, 4);
    if (c->vertexCounter == pfh_verts_per_primitive(c->currentDrawMode)) {
        pfh_process_primitive(c);
        pfh_carry_over(c);
    }
}

void pfVertex2i(PFint x, PFint y) { PFfloat v[4] = { (PFfloat)x, (PFfloat)y, 0.0f, 1.0f }; pfVertex4fv(v); }
void pfVertex2f(PFfloat x, PFfloat y) { PFfloat v[4] = { x, y, 0.0f, 1.0f }; pfVertex4fv(v); }
void pfVertex2fv(const PFfloat *p) { PFfloat v[4] = { p[0], p[1], 0.0f, 1.0f }; pfVertex4fv(v); }
void pfVertex3i(PFint x, PFint y, PFint z) { PFfloat v[4] = { (PFfloat)x, (PFfloat)y, (PFfloat)z, 1.0f }; pfVertex4fv(v); }
void pfVertex3f(PFfloat x, PFfloat y, PFfloat z) { PFfloat v[4] = { x, y, z, 1.0f }; pfVertex4fv(v); }
void pfVertex3fv(const PFfloat *p) { PFfloat v[4] = { p[0], p[1], p[2], 1.0f }; pfVertex4fv(v); }
void pfVertex4i(PFint x, PFint y, PFint z, PFint w) { PFfloat v[4] = { (PFfloat)x, (PFfloat)y, (PFfloat)z, (PFfloat)w }; pfVertex4fv(v); }
void pfVertex4f(PFfloat x, PFfloat y, PFfloat z, PFfloat w) { PFfloat v[4] = { x, y, z, w }; pfVertex4fv(v); }

void pfColor(PFcolor color)
{
    CTX;
    if (!(c->state & PF_COLOR_MATERIAL)) { c->currentColor = color; return; }
    /* colour tracking: the material follows, the current colour does not change (context.c:1687-1717) */
    pf_material *a = &c->material[0], *b = &c->material[1];
    if (c->cmFace == PF_FRONT) b = a; else if (c->cmFace == PF_BACK) a = b;
    switch (c->cmMode) {
    case PF_AMBIENT_AND_DIFFUSE: a->ambient = b->ambient = color; a->diffuse = b->diffuse = color; break;
    case PF_AMBIENT:  a->ambient = b->ambient = color; break;
    case PF_DIFFUSE:  a->diffuse = b->diffuse = color; break;
    case PF_SPECULAR: a->specular = b->specular = color; break;
    case PF_EMISSION: a->emission = b->emission = color; break;
    }
    if (c->lightingMode == PF_PHONG) c->state_dirty = 1;
}

void pfColor1ui(PFuint color) { PFcolor k; memcpy(&k, &color, 4); pfColor(k); }
void pfColor3ub(PFubyte r, PFubyte g, PFubyte b) { pfColor((PFcolor){ r, g, b, 255 }); }
void pfColor3ubv(const PFubyte *v) { pfColor((PFcolor){ v[0], v[1], v[2], 255 }); }
void pfColor3us(PFushort r, PFushort g, PFushort b) { pfColor((PFcolor){ (PFubyte)(r >> 8), (PFubyte)(g >> 8), (PFubyte)(b >> 8), 255 }); }
void pfColor3usv(const PFushort *v) { pfColor3us(v[0], v[1], v[2]); }
void pfColor3ui(PFuint r, PFuint g, PFuint b) { pfColor((PFcolor){ (PFubyte)(r >> 24), (PFubyte)(g >> 24), (PFubyte)(b >> 24), 255 }); }
void pfColor3uiv(const PFuint *v) { pfColor3ui(v[0], v[1], v[2]); }
void pfColor3f(PFfloat r, PFfloat g, PFfloat b) { pfColor((PFcolor){ (PFubyte)(r * 255), (PFubyte)(g * 255), (PFubyte)(b * 255), 255 }); }
void pfColor3fv(const PFfloat *v) { pfColor3f(v[0], v[1], v[2]); }
void pfColor4ub(PFubyte r, PFubyte g, PFubyte b, PFubyte a) { pfColor((PFcolor){ r, g, b, a }); }
void pfColor4ubv(const PFubyte *v) { pfColor((PFcolor){ v[0], v[1], v[2], v[3] }); }
void pfColor4us(PFushort r, PFushort g, PFushort b, PFushort a) { pfColor((PFcolor){ (PFubyte)(r >> 8), (PFubyte)(g >> 8), (PFubyte)(b >> 8), (PFubyte)(a >> 8) }); }
void pfColor4usv(const PFushort *v) { pfColor4us(v[0], v[1], v[2], v[3]); }
void pfColor4ui(PFuint r, PFuint g, PFuint b, PFuint a) { pfColor((PFcolor){ (PFubyte)(r >> 24), (PFubyte)(g >> 24), (PFubyte)(b >> 24), (PFubyte)(a >> 24) }); }
void pfColor4uiv(const PFuint *v) { pfColor4ui(v[0], v[1], v[2], v[3]); }
void pfColor4f(PFfloat r, PFfloat g, PFfloat b, PFfloat a) { pfColor((PFcolor){ (PFubyte)(r * 255), (PFubyte)(g * 255), (PFubyte)(b * 255), (PFubyte)(a * 255) }); }
void pfColor4fv(const PFfloat *v) { pfColor4f(v[0], v[1], v[2], v[3]); }

void pfTexCoord2f(PFfloat u, PFfloat v)
{
    CTX;
    c->currentTexcoord[0] = u; c->currentTexcoord[1] = v;
    v2_transform(c->currentTexcoord, c->currentTexcoord, c->matTexture);
}

void pfTexCoordfv(const PFfloat *v) { pfTexCoord2f(v[0], v[1]); }

void pfNormal3f(PFfloat x, PFfloat y, PFfloat z)
{
    CTX;
    c->currentNormal[0] = x; c->currentNormal[1] = y; c->currentNormal[2] = z;
    if (c->state & PF_NORMALIZE) v3_normalize(c->currentNormal, c->currentNormal);
}

void pfNormal3fv(const PFfloat *v) { pfNormal3f(v[0], v[1], v[2]); }

/* ---- vertex arrays (context.c:1161-1575); `stride` is stored but ignored, as upstream (Q16) ---- */

void pfVertexPointer(PFint size, PFenum type, PFsizei stride, const void *pointer)
{
    CTX;
    if (size < 2 || size > 4) { c->errCode = PF_INVALID_VALUE; return; }
    if (!(type == PF_SHORT || type == PF_INT || type == PF_FLOAT || type == PF_DOUBLE)) { c->errCode = PF_INVALID_ENUM; return; }
    c->apos = (pf_attrib){ pointer, stride, size, (PFdatatype)type };
}

void pfNormalPointer(PFenum type, PFsizei stride, const void *pointer)
{
    CTX;
    if (!(type == PF_FLOAT || type == PF_DOUBLE)) { c->errCode = PF_INVALID_ENUM; return; }
    c->anrm = (pf_attrib){ pointer, stride, 3, (PFdatatype)type };
}

void pfTexCoordPointer(PFenum type, PFsizei stride, const void *pointer)
{
    CTX;
    if (!(type == PF_FLOAT || type == PF_DOUBLE)) { c->errCode = PF_INVALID_ENUM; return; }
    c->atex = (pf_attrib){ pointer, stride, 2, (PFdatatype)type };
}

void pfColorPointer(PFint size, PFenum type, PFsizei stride, const void *pointer)
{
    CTX;
    if (size < 3 || size > 4) { c->errCode = PF_INVALID_VALUE; return; }
    if (!(type == PF_UNSIGNED_BYTE || type == PF_UNSIGNED_SHORT || type == PF_UNSIGNED_INT || type == PF_FLOAT || type == PF_DOUBLE)) { c->errCode = PF_INVALID_ENUM; return; }
    c->acol = (pf_attrib){ pointer, stride, size, (PFdatatype)type };
}

static int fetch_float(const pf_attrib *a, size_t idx, float *out, int n)
{
    if (a->type == PF_FLOAT) { memcpy(out, (const PFfloat *)a->buffer + idx * (size_t)n, (size_t)n * sizeof(float)); return 1; }
    for (int k = 0; k < n; k++) {
        size_t e = idx * (size_t)n + (size_t)k;
        switch (a->type) {
        case PF_SHORT:  out[k] = (float)((const PFshort *)a->buffer)[e]; break;
        case PF_INT:    out[k] = (float)((const PFint *)a->buffer)[e]; break;
        case PF_FLOAT:  out[k] = ((const PFfloat *)a->buffer)[e]; break;
        case PF_DOUBLE: out[k] = (float)((const PFdouble *)a->buffer)[e]; break;
        default: return 0;
        }
    }
    return 1;
}

static int fetch_color(const pf_attrib *a, size_t idx, PFcolor *out)
{
    PFubyte *o = (PFubyte *)out;
    memset(o, 0xFF, 4);
    for (int k = 0; k < a->size; k++) {
        size_t e = idx * (size_t)a->size + (size_t)k;
        switch (a->type) {
        case PF_UNSIGNED_BYTE:  o[k] = ((const PFubyte *)a->buffer)[e]; break;
        case PF_UNSIGNED_SHORT: o[k] = (PFubyte)(((const PFushort *)a->buffer)[e] >> 8); break;
        case PF_UNSIGNED_INT:   o[k] = (PFubyte)(((const PFuint *)a->buffer)[e] >> 24); break;
        case PF_FLOAT:          o[k] = (PFubyte)(((const PFfloat *)a->buffer)[e] * 255); break;
        case PF_DOUBLE:         o[k] = (PFubyte)(((const PFdouble *)a->buffer)[e] * 255); break;
        default: return 0;
        }
    }
    return 1;
}

static void draw_indexed(PFdrawmode mode, PFsizei count, PFint first, int indexed, PFdatatype itype, const void *indices)
{
    CTX;
    if (!(c->state & PF_VERTEX_ARRAY)) { c->errCode = PF_INVALID_OPERATION; return; }
    int useTex = (c->state & PF_TEXTURE_COORD_ARRAY) && c->atex.buffer;
    int useNrm = (c->state & PF_NORMAL_ARRAY) && c->anrm.buffer;
    int useCol = (c->state & PF_COLOR_ARRAY) && c->acol.buffer;
    PFsizei per = pfh_verts_per_primitive(mode);

    if (mode == PF_TRIANGLES && pfh_device_draw(c, count, first, indexed, itype, indices, useNrm, useTex, useCol)) {
        pfh_end_of_draw(c);
        return;
    }
    for (PFsizei i = 0; i < per; i++) {
        memset(&c->vertexBuffer[i], 0, sizeof(pf_vertex));
        memcpy(&c->vertexBuffer[i].color, &c->currentColor, 4);
    }
    pfBegin(mode);
    for (PFsizei i = 0; i < count; i++) {
        pf_vertex *vx = &c->vertexBuffer[c->vertexCounter++];
        size_t j;
        if (indexed) {
            switch (itype) {
            case PF_UNSIGNED_BYTE:  j = ((const PFubyte *)indices)[i]; break;
            case PF_UNSIGNED_SHORT: j = ((const PFushort *)indices)[i]; break;
            default:                j = ((const PFuint *)indices)[i]; break;
            }
        } else j = (size_t)first + i;
        vx->position[0] = vx->position[1] = vx->position[2] = 0.0f; vx->position[3] = 1.0f;
        if (!fetch_float(&c->apos, j, vx->position, c->apos.size)) { c->errCode = PF_INVALID_ENUM; if (indexed) return; }
        if (useNrm && !fetch_float(&c->anrm, j, vx->normal, 3)) { c->errCode = PF_INVALID_ENUM; if (indexed) return; }
        if (useTex && !fetch_float(&c->atex, j, vx->texcoord, 2)) { c->errCode = PF_INVALID_ENUM; if (indexed) return; }
        if (useCol && !fetch_color(&c->acol, j, (PFcolor *)&vx->color)) { c->errCode = PF_INVALID_ENUM; if (indexed) return; }
        if (c->vertexCounter == per) {
            pfh_process_primitive(c);
            pfh_carry_over(c);
        }
    }
    pfEnd();
}

void pfDrawElements(PFdrawmode mode, PFsizei count, PFdatatype type, const void *indices)
{
    if (!(type == PF_UNSIGNED_BYTE || type == PF_UNSIGNED_SHORT || type == PF_UNSIGNED_INT)) { pf_cur->errCode = PF_INVALID_ENUM; return; }
    draw_indexed(mode, count, 0, 1, type, indices);
}

void pfDrawArrays(PFdrawmode mode, PFint first, PFsizei count) { draw_indexed(mode, count, first, 0, PF_UNSIGNED_INT, NULL); }

/* ---- full-surface operations (context.c:1918-2441) ----------------------------------------------
 * pfRect*, pfDrawPixels, pfFogProcess and pfReadPixels run on the device surface: the
 * front end does what the reference does before its pixel loop and pfcu_surface_* runs the loop, ordered
 * behind the batches submitted before it.  Only pfPostProcess - a HOST callback per pixel - needs the
 * mirror: it synchronises, calls the function for every pixel and marks the mirror as newer. */

static float *host_depth(pf_ctx *c, pf_surf *s, int *temp)
{
    *temp = 0;
    pfh_sync_surface(c, s);
    if (s->zhost) return s->zhost;
    float *z = (float *)malloc((size_t)s->tex->w * s->tex->h * sizeof(float));
    if (!z) { c->errCode = PF_ERROR_OUT_OF_MEMORY; return NULL; }
    pfcu_surface_download(s->dev, NULL, z, 0, s->tex->h);
    *temp = 1;
    return z;
}

static void host_depth_done(pf_surf *s, float *z, int temp, int modified)
{
    if (temp) {
        if (modified) pfcu_surface_upload(s->dev, NULL, z, 0, s->tex->h);
        pfcu_finish();
        free(z);
    }
}

/* everything queued for the bound surface goes first; a mirror the application wrote to is uploaded */
static pf_surf *surface_op_begin(pf_ctx *c)
{
    pf_surf *s = c->cur_surf;
    pfh_flush(c);
    pfh_upload_if_needed(c, s);
    return s;
}

static void surface_op_end(pf_ctx *c, pf_surf *s, PFint y0, PFint y1, int rc)       /* rows [y0, y1] were written */
{
    if (rc != PFCU_OK) {
        fprintf(stderr, "pixelforge-b200: surface operation failed (%d): %s\n", rc, pfcu_last_error());
        c->errCode = (rc == PFCU_ERR_OOM) ? PF_ERROR_OUT_OF_MEMORY : PF_INVALID_OPERATION;
    }
    if (y0 < 0) y0 = 0;
    if (y1 >= (PFint)s->tex->h) y1 = (PFint)s->tex->h - 1;
    if (y0 <= y1) {
        if (!s->dev_newer) { s->dirty_y0 = (PFuint)y0; s->dirty_y1 = (PFuint)y1 + 1; }
        else { if ((PFuint)y0 < s->dirty_y0) s->dirty_y0 = (PFuint)y0; if ((PFuint)y1 + 1 > s->dirty_y1) s->dirty_y1 = (PFuint)y1 + 1; }
        s->dev_newer = 1; s->readback_queued = 0;
    }
    pfh_end_of_draw(c);
}

void pfRectf(PFfloat x1, PFfloat y1, PFfloat x2, PFfloat y2)
{
    CTX;
    pfh_update_matrices(c, 0);
    PFfloat a[4] = { x1, y1, 0.0f, 1.0f }, b[4] = { x2, y2, 0.0f, 1.0f };
    v4_transform(a, a, c->matMVP); v4_transform(b, b, c->matMVP);
    PFint ix1 = (PFint)(c->vpPos[0] + (a[0] + 1.0f) * 0.5f * c->vpDim[0]);
    PFint iy1 = (PFint)(c->vpPos[1] + (1.0f - a[1]) * 0.5f * c->vpDim[1]);
    PFint ix2 = (PFint)(c->vpPos[0] + (b[0] + 1.0f) * 0.5f * c->vpDim[0]);
    PFint iy2 = (PFint)(c->vpPos[1] + (1.0f - b[1]) * 0.5f * c->vpDim[1]);
    if (ix2 < ix1) { PFint t = ix1; ix1 = ix2; ix2 = t; }
    if (iy2 < iy1) { PFint t = iy1; iy1 = iy2; iy2 = t; }
    ix1 = PF_CLAMP(ix1, c->vpMin[0], c->vpMax[0]); iy1 = PF_CLAMP(iy1, c->vpMin[1], c->vpMax[1]);
    ix2 = PF_CLAMP(ix2, c->vpMin[0], c->vpMax[0]); iy2 = PF_CLAMP(iy2, c->vpMin[1], c->vpMax[1]);
    pf_surf *s = surface_op_begin(c);
    uint32_t rgba; memcpy(&rgba, &c->currentColor, 4);
    surface_op_end(c, s, iy1, iy2 + 1, pfcu_surface_rect(s->dev, ix1, iy1, ix2, iy2, rgba));
}

void pfRects(PFshort x1, PFshort y1, PFshort x2, PFshort y2) { pfRectf((PFfloat)x1, (PFfloat)y1, (PFfloat)x2, (PFfloat)y2); }
void pfRectsv(const PFshort *v1, const PFshort *v2) { pfRectf((PFfloat)v1[0], (PFfloat)v1[1], (PFfloat)v2[0], (PFfloat)v2[1]); }
void pfRectfv(const PFfloat *v1, const PFfloat *v2) { pfRectf(v1[0], v1[1], v2[0], v2[1]); }
/* exported by the reference library although its header never declares them (context.c:1928-1936) */
PF_API void pfRecti(PFint x1, PFint y1, PFint x2, PFint y2) { pfRectf((PFfloat)x1, (PFfloat)y1, (PFfloat)x2, (PFfloat)y2); }
PF_API void pfRectiv(const PFint *v1, const PFint *v2) { pfRectf((PFfloat)v1[0], (PFfloat)v1[1], (PFfloat)v2[0], (PFfloat)v2[1]); }

void pfDrawPixels(PFsizei width, PFsizei height, PFpixelformat format, PFdatatype type, const void *pixels)
{
    CTX;
    if (width == 0 || height == 0) { c->errCode = PF_INVALID_VALUE; return; }
    if (format > PF_BGRA || type > PF_DOUBLE || pfx_bytes(PFCU_PIX(format, type)) == 0) { c->errCode = PF_INVALID_ENUM; return; }
    pfh_update_matrices(c, 0);
    PFfloat rp[4]; memcpy(rp, c->rasterPos, 16);
    v4_transform(rp, rp, c->matMVP);
    pfcu_pixels d; memset(&d, 0, sizeof d);
    d.pixels = pixels; d.width = width; d.height = height; d.format = PFCU_PIX(format, type);
    d.xs = (PFint)(c->vpPos[0] + (rp[0] + 1.0f) * 0.5f * c->vpDim[0]);
    d.ys = (PFint)(c->vpPos[1] + (1.0f - rp[1]) * 0.5f * c->vpDim[1]);
    d.z = rp[2];
    d.xmin = PF_CLAMP(d.xs, c->vpMin[0], c->vpMax[0]); d.ymin = PF_CLAMP(d.ys, c->vpMin[1], c->vpMax[1]);
    d.xmax = (PFint)PF_CLAMP(d.xs + width * c->pixelZoom[0], (PFfloat)c->vpMin[0], (PFfloat)c->vpMax[0]);
    d.ymax = (PFint)PF_CLAMP(d.ys + height * c->pixelZoom[1], (PFfloat)c->vpMin[1], (PFfloat)c->vpMax[1]);
    d.inv_xlen = 1.0f / (PFfloat)(width * c->pixelZoom[0]); d.inv_ylen = 1.0f / (PFfloat)(height * c->pixelZoom[1]);
    d.flags = ((c->state & PF_DEPTH_TEST) ? PFCU_ST_DEPTH_TEST : 0u) | ((c->state & PF_BLEND) ? PFCU_ST_BLEND : 0u);
    d.blend_mode = (uint8_t)c->blendMode; d.depth_func = (uint8_t)c->depthMode;
    /* the source may be the mirror of one of our surfaces (the aux buffer after pfSwapBuffers, a framebuffer's
       texture): bring it up to date before it is uploaded */
    for (pf_surf *o = pfh_surf_first(); o; o = pfh_surf_next(o))
        if (o->tex && o->tex->pixels == pixels) pfh_sync_surface(c, o);
    pf_surf *s = surface_op_begin(c);
    surface_op_end(c, s, d.ymin, d.ymax + 1, pfcu_surface_draw_pixels(s->dev, &d));
}

void pfPixelZoom(PFfloat xf, PFfloat yf) { pf_cur->pixelZoom[0] = xf; pf_cur->pixelZoom[1] = yf; }
static void raster_pos(PFfloat x, PFfloat y, PFfloat z, PFfloat w) { PFfloat *r = pf_cur->rasterPos; r[0] = x; r[1] = y; r[2] = z; r[3] = w; }
void pfRasterPos2i(PFint x, PFint y) { raster_pos((PFfloat)x, (PFfloat)y, 0.0f, 1.0f); }
void pfRasterPos2f(PFfloat x, PFfloat y) { raster_pos(x, y, 0.0f, 1.0f); }
void pfRasterPos2fv(const PFfloat *v) { raster_pos(v[0], v[1], 0.0f, 1.0f); }
void pfRasterPos3i(PFint x, PFint y, PFint z) { raster_pos((PFfloat)x, (PFfloat)y, (PFfloat)z, 1.0f); }
void pfRasterPos3f(PFfloat x, PFfloat y, PFfloat z) { raster_pos(x, y, z, 1.0f); }
void pfRasterPos3fv(const PFfloat *v) { raster_pos(v[0], v[1], v[2], 1.0f); }
void pfRasterPos4i(PFint x, PFint y, PFint z, PFint w) { raster_pos((PFfloat)x, (PFfloat)y, (PFfloat)z, (PFfloat)w); }
void pfRasterPos4f(PFfloat x, PFfloat y, PFfloat z, PFfloat w) { raster_pos(x, y, z, w); }
void pfRasterPos4fv(const PFfloat *v) { raster_pos(v[0], v[1], v[2], v[3]); }

/* fog parameter setters keep the reference's quirks (PF_FOG_DENSITY writes fog.mode, context.c:2161-2273) */
void pfFogi(PFfogparam pname, PFint param)
{
    CTX;
    switch (pname) {
    case PF_FOG_MODE: if (param >= PF_LINEAR && param <= PF_EXP2) c->fog.mode = (PFfogmode)param; else c->errCode = PF_INVALID_VALUE; break;
    case PF_FOG_DENSITY: if (param == 0 || param == 1) c->fog.mode = (PFfogmode)param; else c->errCode = PF_INVALID_VALUE; break;
    case PF_FOG_START: c->fog.start = (PFfloat)param; break;
    case PF_FOG_END: c->fog.end = (PFfloat)param; break;
    default: c->errCode = PF_INVALID_ENUM; break;
    }
}

void pfFogf(PFfogparam pname, PFfloat param)
{
    CTX;
    switch (pname) {
    case PF_FOG_DENSITY: if (param >= 0 && param <= 1) c->fog.mode = (PFfogmode)param; else c->errCode = PF_INVALID_VALUE; break;
    case PF_FOG_START: c->fog.start = param; break;
    case PF_FOG_END: c->fog.end = param; break;
    default: c->errCode = PF_INVALID_ENUM; break;
    }
}

void pfFogiv(PFfogparam pname, PFint *param)
{
    CTX;
    switch (pname) {
    case PF_FOG_MODE: case PF_FOG_DENSITY: c->fog.mode = (PFfogmode)*param; break;
    case PF_FOG_START: c->fog.start = (PFfloat)*param; break;
    case PF_FOG_END: c->fog.end = (PFfloat)*param; break;
    case PF_FOG_COLOR: c->fog.color = (PFcolor){ (PFubyte)param[0], (PFubyte)param[1], (PFubyte)param[2], (PFubyte)param[3] }; break;
    default: c->errCode = PF_INVALID_ENUM; break;
    }
}

void pfFogfv(PFfogparam pname, PFfloat *param)
{
    CTX;
    switch (pname) {
    case PF_FOG_DENSITY: c->fog.mode = (PFfogmode)*param; break;
    case PF_FOG_START: c->fog.start = *param; break;
    case PF_FOG_END: c->fog.end = *param; break;
    case PF_FOG_COLOR: c->fog.color = (PFcolor){ (PFubyte)(255 * param[0]), (PFubyte)(255 * param[1]), (PFubyte)(255 * param[2]), (PFubyte)(255 * param[3]) }; break;
    default: c->errCode = PF_INVALID_ENUM; break;
    }
}

/* The fog alpha of the exponential modes as the reference computes it (context.c:2331-2338), with the host's libm. */
static PFubyte fog_alpha_host(const pf_ctx *c, PFfloat depth)
{
    const PFfloat start = c->fog.start, density = c->fog.density;
    const PFubyte alpha = c->fog.color.a;
    PFfloat t = (c->fog.mode == PF_EXP) ? 1.0f - expf(-density * (depth - start)) : 1.0f - exp2f(-density * (depth - start));
    return (PFubyte)(t * alpha);
}

/* floats in their numeric order as integers (for bisection over bit patterns) */
static int32_t fog_ord(float f) { int32_t i; memcpy(&i, &f, 4); return i < 0 ? (int32_t)(0x80000000u - (uint32_t)i) : i; }
static float fog_unord(int32_t o) { int32_t i = o < 0 ? (int32_t)(0x80000000u - (uint32_t)o) : o; float f; memcpy(&f, &i, 4); return f; }

/* thresholds[k-1] = the smallest depth in (start, end) whose fog alpha is >= k.  Returns their number, or -1 when the
 * host's function turns out not to be monotonic over the interval (then a table cannot stand in for it). */
static int fog_thresholds(const pf_ctx *c, float *thr)
{
    const int32_t lo0 = fog_ord(c->fog.start) + 1, hi0 = fog_ord(c->fog.end) - 1;     /* the open interval */
    if (lo0 > hi0) return 0;
    const int top = fog_alpha_host(c, fog_unord(hi0));
    int32_t prev = lo0;
    int n = 0;
    for (int k = 1; k <= top; k++) {
        int32_t lo = prev, hi = hi0;                  /* alpha(hi) >= k; find the first position with alpha >= k */
        if (fog_alpha_host(c, fog_unord(lo)) >= k) hi = lo;
        while (lo < hi) {
            const int32_t mid = lo + (int32_t)(((int64_t)hi - lo) >> 1);
            if (fog_alpha_host(c, fog_unord(mid)) >= k) hi = mid; else lo = mid + 1;
        }
        thr[n++] = fog_unord(hi);
        prev = hi;
    }
    /* spot check: between consecutive thresholds the function must hold the step's value */
    uint32_t rs = 12345u;
    for (int k = 0; k <= n; k++) {
        const int32_t a = k ? fog_ord(thr[k - 1]) : lo0, b = k < n ? fog_ord(thr[k]) - 1 : hi0;
        if (a > b) continue;
        for (int j = 0; j < 24; j++) {
            rs = rs * 1664525u + 1013904223u;
            const int32_t p = j == 0 ? a : (j == 1 ? b : a + (int32_t)(((uint64_t)(rs >> 1) * (uint64_t)((int64_t)b - a + 1)) >> 31));
            if (fog_alpha_host(c, fog_unord(p)) != k) return -1;
        }
    }
    return n;
}

/* Diagnostics for the fog alpha steps the device uses instead of expf / exp2f (pfcu_fog.thresholds): tabulates the
 * current context's fog state and compares "number of thresholds <= depth" with the host function on `samples`
 * pseudo-random depths in (start, end) plus both neighbours of every threshold.  Returns the number of mismatches,
 * -1 when the host function is not monotonic (pfFogProcess then refuses), -2 without a context or in PF_LINEAR mode. */
int pfxFogTableCheck(PFuint samples)
{
    pf_ctx *c = pf_cur;
    if (!c || c->fog.mode == PF_LINEAR || (uint32_t)c->fog.mode > (uint32_t)PF_EXP2) return -2;
    float thr[256];
    const int n = fog_thresholds(c, thr);
    if (n < 0) return -1;
    const int32_t lo0 = fog_ord(c->fog.start) + 1, hi0 = fog_ord(c->fog.end) - 1;
    if (lo0 > hi0) return 0;
    int bad = 0;
    uint32_t rs = 2463534242u;
    for (PFuint i = 0; i < samples + 2u * (PFuint)n; i++) {
        int32_t p;
        if (i < samples) { rs = rs * 1664525u + 1013904223u; p = lo0 + (int32_t)(((uint64_t)(rs >> 1) * (uint64_t)((int64_t)hi0 - lo0 + 1)) >> 31); }
        else { const PFuint k = (i - samples) >> 1; p = fog_ord(thr[k]) - (int32_t)((i - samples) & 1u); if (p < lo0) continue; }
        const float d = fog_unord(p);
        int cnt = 0; while (cnt < n && thr[cnt] <= d) cnt++;
        if (cnt != (int)fog_alpha_host(c, d)) bad++;
    }
    return bad;
}

void pfFogProcess(void)
{
    CTX;
    pfcu_fog f; memset(&f, 0, sizeof f);
    float thr[256];
    f.start = c->fog.start; f.end = c->fog.end; f.inv_len = 1 / (c->fog.end - c->fog.start); f.density = c->fog.density;
    memcpy(&f.rgba, &c->fog.color, 4);
    f.mode = (uint32_t)c->fog.mode;
    if (f.mode > (uint32_t)PF_EXP2) f.mode = 3u;            /* no case of the reference's switch matches: t stays 0 */
    else if (c->fog.mode != PF_LINEAR) {
        const int n = fog_thresholds(c, thr);
        if (n < 0) {
            fprintf(stderr, "pixelforge-b200: pfFogProcess: the host's expf/exp2f is not monotonic over the fog range; cannot tabulate it for the device\n");
            c->errCode = PF_INVALID_OPERATION; return;
        }
        f.thresholds = thr; f.n_thresholds = (uint32_t)n;
    }
    pf_surf *s = surface_op_begin(c);
    surface_op_end(c, s, 0, (PFint)s->tex->h - 1, pfcu_surface_fog(s->dev, &f));
}

void pfReadPixels(PFint x, PFint y, PFsizei width, PFsizei height, PFpixelformat format, PFdatatype type, void *pixels)
{
    CTX;
    if (format > PF_BGRA || type > PF_DOUBLE || pfx_bytes(PFCU_PIX(format, type)) == 0) { c->errCode = PF_INVALID_ENUM; return; }
    pf_surf *s = surface_op_begin(c);
    PFint W = (PFint)s->tex->w, H = (PFint)s->tex->h;
    PFsizei xMin = (PFsizei)PF_CLAMP(x, 0, W - 1), yMin = (PFsizei)PF_CLAMP(y, 0, H - 1);
    PFsizei xMax = (PFsizei)PF_CLAMP(x + (PFint)width, 0, W), yMax = (PFsizei)PF_CLAMP(y + (PFint)height, 0, H);
    if (xMax <= xMin || yMax <= yMin) return;
    /* destination index (ySrc - yMin) * width + (xSrc - xMin) (context.c:2383-2386); a region wider than `width`
       (x < 0 moves xMin to 0) would make rows overlap upstream - the columns past `width` are dropped here */
    PFsizei cols = xMax - xMin; if (cols > width) cols = width;
    int rc = pfcu_surface_read_pixels(s->dev, xMin, yMin, cols, yMax - yMin, width, PFCU_PIX(format, type), pixels);
    if (rc != PFCU_OK) { fprintf(stderr, "pixelforge-b200: pfReadPixels failed (%d): %s\n", rc, pfcu_last_error()); c->errCode = PF_INVALID_OPERATION; }
}

void pfPostProcess(PFpostprocessfunc fn)
{
    CTX;
    pf_surf *s = c->cur_surf;
    int temp; float *z = host_depth(c, s, &temp);
    if (!z) return;
    PFint W = (PFint)s->tex->w, H = (PFint)s->tex->h;
    for (PFint y = 0; y < H; y++)
        for (PFint x = 0; x < W; x++) {
            size_t o = (size_t)y * (size_t)W + (size_t)x;
            pfh_pixel_set(s->tex, o, fn(x, y, z[o], pfh_pixel_get(s->tex, o)));
        }
    s->host_newer = 1;
    host_depth_done(s, z, temp, 0);
}

/* ---- getters (getter.c) -------------------------------------------------------------------------- */

void pfGetBooleanv(PFenum pname, PFboolean *params)
{
    CTX;
    switch (pname) {
    case PF_TEXTURE_2D: case PF_FRAMEBUFFER: case PF_BLEND: case PF_DEPTH_TEST: case PF_CULL_FACE: case PF_NORMALIZE:
    case PF_LIGHTING: case PF_COLOR_MATERIAL: case PF_VERTEX_ARRAY: case PF_NORMAL_ARRAY: case PF_COLOR_ARRAY:
    case PF_TEXTURE_COORD_ARRAY: *params = (c->state & pname) != 0; break;
    default: c->errCode = PF_INVALID_ENUM; break;
    }
}

void pfGetIntegerv(PFenum pname, PFint *p)
{
    CTX;
    switch (pname) {
    case PF_VIEWPORT: p[0] = c->vpPos[0]; p[1] = c->vpPos[1]; p[2] = (PFint)c->vpDim[0] + 1; p[3] = (PFint)c->vpDim[1] + 1; break;
    case PF_COLOR_CLEAR_VALUE: p[0] = c->clearColor.r; p[1] = c->clearColor.g; p[2] = c->clearColor.b; p[3] = c->clearColor.a; break;
    case PF_CULL_FACE_MODE: *p = (PFint)c->cullFace; break;
    case PF_CURRENT_COLOR: p[0] = c->currentColor.r; p[1] = c->currentColor.g; p[2] = c->currentColor.b; p[3] = c->currentColor.a; break;
    case PF_CURRENT_RASTER_POSITION: p[0] = (PFint)c->rasterPos[0]; p[1] = (PFint)c->rasterPos[1]; break;
    case PF_POLYGON_MODE: p[0] = (PFint)c->polygonMode[0]; p[1] = (PFint)c->polygonMode[1]; break;
    case PF_MATRIX_MODE: *p = (PFint)c->matrixMode; break;
    case PF_MAX_PROJECTION_STACK_DEPTH: *p = PFH_PROJECTION_STACK; break;
    case PF_MAX_MODELVIEW_STACK_DEPTH: *p = PFH_MODELVIEW_STACK; break;
    case PF_MAX_TEXTURE_STACK_DEPTH: *p = PFH_TEXTURE_STACK; break;
    case PF_SHADE_MODEL: *p = (PFint)c->shadingMode; break;
    case PF_MAX_LIGHTS: *p = PFH_MAX_LIGHTS; break;
    case PF_VERTEX_ARRAY_SIZE: *p = c->apos.size; break;
    case PF_VERTEX_ARRAY_STRIDE: *p = (PFint)c->apos.stride; break;
    case PF_VERTEX_ARRAY_TYPE: *p = (PFint)c->apos.type; break;
    case PF_NORMAL_ARRAY_STRIDE: *p = (PFint)c->anrm.stride; break;
    case PF_NORMAL_ARRAY_TYPE: *p = (PFint)c->anrm.type; break;
    case PF_TEXTURE_COORD_ARRAY_STRIDE: *p = (PFint)c->atex.stride; break;
    case PF_TEXTURE_COORD_ARRAY_TYPE: *p = (PFint)c->atex.type; break;
    case PF_COLOR_ARRAY_SIZE: *p = c->acol.size; break;
    case PF_COLOR_ARRAY_STRIDE: *p = (PFint)c->acol.stride; break;
    case PF_COLOR_ARRAY_TYPE: *p = (PFint)c->acol.type; break;
    default: c->errCode = PF_INVALID_ENUM; break;
    }
}

static int get_real(pf_ctx *c, PFenum pname, double *p)      /* returns the number of values written */
{
    const double k = PF_INV_255;
    switch (pname) {
    case PF_COLOR_CLEAR_VALUE: p[0] = c->clearColor.r * k; p[1] = c->clearColor.g * k; p[2] = c->clearColor.b * k; p[3] = c->clearColor.a * k; return 4;
    case PF_DEPTH_CLEAR_VALUE: p[0] = c->clearDepth; return 1;
    case PF_CURRENT_COLOR: p[0] = c->currentColor.r * k; p[1] = c->currentColor.g * k; p[2] = c->currentColor.b * k; p[3] = c->currentColor.a * k; return 4;
    case PF_CURRENT_NORMAL: for (int i = 0; i < 3; i++) p[i] = c->currentNormal[i]; return 3;
    case PF_CURRENT_TEXTURE_COORDS: p[0] = c->currentTexcoord[0]; p[1] = c->currentTexcoord[1]; return 2;
    case PF_CURRENT_RASTER_POSITION: p[0] = c->rasterPos[0]; p[1] = c->rasterPos[1]; return 2;
    case PF_POINT_SIZE: p[0] = c->pointSize; return 1;
    case PF_LINE_WIDTH: p[0] = c->lineWidth; return 1;
    case PF_PROJECTION_MATRIX: for (int i = 0; i < 16; i++) p[i] = c->matProjection[i]; return 16;
    case PF_MODELVIEW_MATRIX: { pf_mat4 mv; m4_mul(mv, c->matModel, c->matView); for (int i = 0; i < 16; i++) p[i] = mv[i]; return 16; }
    case PF_TEXTURE_MATRIX: for (int i = 0; i < 16; i++) p[i] = c->matTexture[i]; return 16;
    case PF_ZOOM_X: p[0] = c->pixelZoom[0]; return 1;
    case PF_ZOOM_Y: p[0] = c->pixelZoom[1]; return 1;
    default: c->errCode = PF_INVALID_ENUM; return 0;
    }
}

void pfGetFloatv(PFenum pname, PFfloat *params)
{
    CTX; double t[16]; int n = get_real(c, pname, t);
    if (pname == PF_COLOR_CLEAR_VALUE || pname == PF_CURRENT_COLOR) {      /* float multiply in the reference */
        const PFcolor k = pname == PF_CURRENT_COLOR ? c->currentColor : c->clearColor;
        params[0] = k.r * (PFfloat)PF_INV_255; params[1] = k.g * (PFfloat)PF_INV_255;
        params[2] = k.b * (PFfloat)PF_INV_255; params[3] = k.a * (PFfloat)PF_INV_255;
        return;
    }
    for (int i = 0; i < n; i++) params[i] = (PFfloat)t[i];
}

void pfGetDoublev(PFenum pname, PFdouble *params)
{
    CTX; double t[16]; int n = get_real(c, pname, t);
    for (int i = 0; i < n; i++) params[i] = t[i];
}

void pfGetPointerv(PFenum pname, const void **params)
{
    CTX;
    switch (pname) {
    case PF_TEXTURE_2D: *params = c->currentTexture; break;
    case PF_FRAMEBUFFER: *params = c->bindedFramebuffer; break;
    case PF_BLEND_FUNC: *params = &c->blendMode; break;      /* the reference returns its function-pointer slot */
    case PF_DEPTH_FUNC: *params = &c->depthMode; break;
    default: c->errCode = PF_INVALID_ENUM; break;
    }
}

/* ---- pfx extensions -------------------------------------------------------------------------------- */

void pfxSetSyncMode(PFboolean explicitSync) { pfh_set_sync_mode(explicitSync ? 1 : 0); }
void pfxFlush(void) { if (pf_cur) pfh_flush(pf_cur); }
void pfxFinish(void) { if (pf_cur) pfh_sync_surface(pf_cur, pf_cur->cur_surf); else pfcu_finish(); }

void pfxGetCounters(PFXcounters *out)
{
    pfcu_counters k; memset(&k, 0, sizeof k);
    if (pf_cur) pfh_flush(pf_cur);
    pfcu_get_counters(&k);
    out->triangles_submitted = k.triangles_submitted; out->triangles_rasterised = k.triangles_rasterised;
    out->pixels_shaded = k.pixels_shaded; out->pixels_depth_failed = k.pixels_depth_failed;
    out->kernel_launches = k.kernel_launches; out->bytes_h2d = k.bytes_h2d; out->bytes_d2h = k.bytes_d2h;
}

void pfxResetCounters(void) { if (pf_cur) pfh_flush(pf_cur); pfcu_finish(); pfcu_reset_counters(); }

void pfxSetTileOwner(PFuint rank, PFuint world)
{
    CTX; if (!c) return;
    pfh_flush(c);
    pfcu_surface_set_tile_owner(c->cur_surf->dev, rank, world);
}

void *pfxGetDeviceColor(void) { return pf_cur ? pfcu_surface_color_ptr(pf_cur->cur_surf->dev) : NULL; }
void *pfxGetDeviceDepth(void) { return pf_cur ? pfcu_surface_depth_ptr(pf_cur->cur_surf->dev) : NULL; }

void pfxReadDepth(PFfloat *out)
{
    CTX; if (!c) return;
    pfh_flush(c);
    pfh_upload_if_needed(c, c->cur_surf);
    pfcu_surface_download(c->cur_surf->dev, NULL, out, 0, c->cur_surf->tex->h);
}

void pfxTextureDirty(PFtexture texture)
{
    pf_tex *t = (pf_tex *)texture;
    if (!t || t->surf || !t->dev || !t->pixels) return;        /* render targets are tracked; never-used textures upload at first use */
    if (pf_cur) pfh_flush(pf_cur);                              /* draws issued so far sample the old texels */
    pfcu_texture_update(t->dev, t->pixels);
}

void *pfxHostAlloc(size_t bytes) { return pfh_runtime_init() ? pfcu_host_alloc(bytes) : NULL; }
void pfxHostFree(void *p) { if (p) { pfcu_finish(); pfcu_host_free(p); } }
void pfxHostStatic(void *p, PFboolean isStatic) { if (p && pfh_runtime_init()) pfcu_host_set_static(p, isStatic ? 1 : 0); }
void pfxHostModified(void *p) { if (p && pfh_runtime_init()) pfcu_host_modified(p); }
void pfxEnableDeviceVertexStage(PFboolean on) { if (pf_cur) pf_cur->device_vertex = on ? 1 : 0; }

void pfxCaptureBegin(void)
{
    CTX; if (!c) return;
    pfh_flush(c);
    c->capturing = 1; c->cap_ntris = 0; c->cap_nstates = 0;
}

void pfxCaptureEnd(const void **states, PFuint *nStates, const void **triangles, PFuint *nTriangles)
{
    CTX; if (!c) return;
    pfh_flush(c);
    c->capturing = 0;
    if (states) *states = c->cap_states;
    if (nStates) *nStates = (PFuint)c->cap_nstates;
    if (triangles) *triangles = c->cap_tris;
    if (nTriangles) *nTriangles = (PFuint)c->cap_ntris;
}

void *pfxGetSurfaceHandle(void) { return pf_cur ? pf_cur->cur_surf->dev : NULL; }

const char *pfxBackendName(void) { return pfcu_backend_name(); }
