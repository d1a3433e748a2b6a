/*
 * pf_objects.c - texture objects, framebuffer objects, render lists and the device runtime.
 *
 * API behaviour follows the reference (src/texture.c, src/framebuffer.c, src/renderlist.c); the
 * storage model differs: every render target owns a device surface (pfcu_surface) that is
 * authoritative, and the caller-visible host arrays are mirrors refreshed at sync points.
 */
#include "pf_internal.h"
#include "../pf_pixfmt.h"

#include <float.h>
#include <stdio.h>

/* ---- device runtime -------------------------------------------------------------------------- */

static int g_runtime_state = 0;     /* 0 untried, 1 ok, -1 failed */

int pfh_runtime_init(void)
{
    if (g_runtime_state) return g_runtime_state > 0;
    int rc = pfcu_init(-1);
    if (rc != PFCU_OK) {
        fprintf(stderr, "pixelforge-b200: no usable CUDA device (%s). This library has no CPU fallback.\n", pfcu_last_error());
        g_runtime_state = -1;
        return 0;
    }
    uint32_t *rcp = NULL, *rsq = NULL; int rb = 0, sb = 0;
    if (!pfh_harvest_rcp(&rcp, &rb) || !pfh_harvest_rsqrt(&rsq, &sb) ||
        pfcu_set_approx_tables(rcp, rb, rsq, sb) != PFCU_OK) {
        fprintf(stderr, "pixelforge-b200: could not reproduce this CPU's RCPPS/RSQRTPS on the device\n");
        free(rcp); free(rsq);
        g_runtime_state = -1;
        return 0;
    }
    free(rcp); free(rsq);
    g_runtime_state = 1;
    return 1;
}

/* ---- pixel helpers (host mirror only; reference tables: src/internal/pixel.h:765-859) -------- */

PFsizei pfh_pixel_bytes(PFpixelformat f, PFdatatype t)
{
    PFsizei comps = 0, sz = 0;
    switch (f) {
    case PF_RED: case PF_GREEN: case PF_BLUE: case PF_ALPHA: case PF_LUMINANCE: comps = 1; break;
    case PF_LUMINANCE_ALPHA: comps = 2; break;
    case PF_RGB: case PF_BGR: comps = 3; break;
    case PF_RGBA: case PF_BGRA: comps = 4; break;
    }
    switch (t) {
    case PF_UNSIGNED_BYTE: case PF_BYTE: sz = 1; break;
    case PF_UNSIGNED_SHORT: case PF_SHORT: case PF_UNSIGNED_SHORT_5_6_5: case PF_UNSIGNED_SHORT_5_5_5_1:
    case PF_UNSIGNED_SHORT_4_4_4_4: case PF_HALF_FLOAT: sz = 2; break;
    case PF_UNSIGNED_INT: case PF_INT: case PF_FLOAT: sz = 4; break;
    case PF_DOUBLE: sz = 8; break;
    }
    return comps * sz;
}

int pfh_tex_format_code(PFpixelformat f, PFdatatype t)
{
    if (t != PF_UNSIGNED_BYTE) return -1;
    switch (f) {
    case PF_RGBA: return PFCU_TEX_RGBA8;
    case PF_BGRA: return PFCU_TEX_BGRA8;
    case PF_RGB:  return PFCU_TEX_RGB8;
    case PF_BGR:  return PFCU_TEX_BGR8;
    default: return -1;
    }
}

/* texture format of the C-ABI: the four 8-bit layouts keep their codes (and fast paths), every other (format, type) pair
   the reference has a getter for is PFCU_TEX_PIX + PFCU_PIX(format, type); -1: no such getter upstream */
int pfh_texture_code(PFpixelformat f, PFdatatype t)
{
    const int legacy = pfh_tex_format_code(f, t);
    if (legacy >= 0) return legacy;
    if ((int)f < 0 || f > PF_BGRA || (int)t < 0 || t > PF_DOUBLE) return -1;
    return pfx_bytes(PFCU_PIX(f, t)) ? PFCU_TEX_PIX + PFCU_PIX(f, t) : -1;
}

PFcolor pfh_pixel_get(const pf_tex *t, size_t i)
{
    const PFubyte *p = (const PFubyte *)t->pixels;
    PFcolor c = { 0, 0, 0, 255 };
    switch (pfh_tex_format_code(t->format, t->type)) {
    case PFCU_TEX_RGBA8: c.r = p[4 * i]; c.g = p[4 * i + 1]; c.b = p[4 * i + 2]; c.a = p[4 * i + 3]; break;
    case PFCU_TEX_BGRA8: c.b = p[4 * i]; c.g = p[4 * i + 1]; c.r = p[4 * i + 2]; c.a = p[4 * i + 3]; break;
    case PFCU_TEX_RGB8:  c.r = p[3 * i]; c.g = p[3 * i + 1]; c.b = p[3 * i + 2]; break;
    case PFCU_TEX_BGR8:  c.b = p[3 * i]; c.g = p[3 * i + 1]; c.r = p[3 * i + 2]; break;
    default:
        if (pfx_bytes(PFCU_PIX(t->format, t->type))) { const uint32_t v = pfx_get(t->pixels, i, PFCU_PIX(t->format, t->type)); memcpy(&c, &v, 4); }
        break;
    }
    return c;
}

void pfh_pixel_set(pf_tex *t, size_t i, PFcolor c)
{
    PFubyte *p = (PFubyte *)t->pixels;
    switch (pfh_tex_format_code(t->format, t->type)) {
    case PFCU_TEX_RGBA8: p[4 * i] = c.r; p[4 * i + 1] = c.g; p[4 * i + 2] = c.b; p[4 * i + 3] = c.a; break;
    case PFCU_TEX_BGRA8: p[4 * i] = c.b; p[4 * i + 1] = c.g; p[4 * i + 2] = c.r; p[4 * i + 3] = c.a; break;
    case PFCU_TEX_RGB8:  p[3 * i] = c.r; p[3 * i + 1] = c.g; p[3 * i + 2] = c.b; break;
    case PFCU_TEX_BGR8:  p[3 * i] = c.b; p[3 * i + 1] = c.g; p[3 * i + 2] = c.r; break;
    default:
        if (pfx_bytes(PFCU_PIX(t->format, t->type))) { uint32_t v; memcpy(&v, &c, 4); pfx_set(t->pixels, i, PFCU_PIX(t->format, t->type), v); }
        break;
    }
}

/* ---- surfaces -------------------------------------------------------------------------------- */

static pf_surf *g_surfs = NULL;     /* registry: PFframebuffer is a by-value struct, we find the
                                       surface again through its texture handle                  */

pf_surf *pfh_surf_first(void) { return g_surfs; }
pf_surf *pfh_surf_next(pf_surf *s) { return s ? s->next : NULL; }

pf_surf *pfh_surf_create(pf_tex *tex, PFfloat *zhost, int z_public)
{
    if (!pfh_runtime_init()) return NULL;
    const int code = pfh_tex_format_code(tex->format, tex->type);
    if (code < 0) {
        fprintf(stderr, "pixelforge-b200: render targets must be PF_RGBA, PF_BGRA, PF_RGB or PF_BGR with PF_UNSIGNED_BYTE (got %d/%d); "
                        "the other framebuffer formats are outside the CUDA path (SURVEY.md 8-f NEXT-4)\n",
                (int)tex->format, (int)tex->type);
        return NULL;
    }
    pf_surf *s = (pf_surf *)calloc(1, sizeof *s);
    if (!s) return NULL;
    s->dev = pfcu_surface_create_format(tex->w, tex->h, code);
    if (!s->dev) { free(s); return NULL; }
    s->tex = tex; s->zhost = zhost; s->z_public = z_public;
    s->dirty_y0 = tex->h; s->dirty_y1 = 0;
    tex->surf = s;
    /* device depth starts at FLT_MAX (context.c:136-138, framebuffer.c:57-59); colour = host content */
    pfcu_surface_fill(s->dev, 0, 0, 1, FLT_MAX);
    if (tex->pixels) pfcu_surface_upload(s->dev, tex->pixels, NULL, 0, tex->h);
    /* large mirrors: pin in place so that read-backs are direct DMA (PF_CUDA_PIN_HOST=0 disables) */
    {
        const char *e = getenv("PF_CUDA_PIN_HOST");
        const size_t px = (size_t)tex->w * tex->h, bytes = px * pfh_pixel_bytes(tex->format, tex->type);
        if (!(e && e[0] == '0') && bytes >= ((size_t)1 << 20)) {
            if (tex->pixels && pfcu_host_register(tex->pixels, bytes) == PFCU_OK) s->pinned_color = tex->pixels;
            if (zhost && pfcu_host_register(zhost, px * 4) == PFCU_OK) s->pinned_depth = zhost;
        }
    }
    s->next = g_surfs; g_surfs = s;
    return s;
}

void pfh_surf_destroy(pf_surf *s)
{
    if (!s) return;
    for (pf_surf **p = &g_surfs; *p; p = &(*p)->next) if (*p == s) { *p = s->next; break; }
    pfcu_finish();
    if (s->pinned_color) pfcu_host_unregister(s->pinned_color);
    if (s->pinned_depth) pfcu_host_unregister(s->pinned_depth);
    if (s->as_texture) pfcu_texture_destroy(s->as_texture);
    pfcu_surface_destroy(s->dev);
    if (s->tex) s->tex->surf = NULL;
    free(s);
}

pf_surf *pfh_surf_lookup(PFtexture tex)
{
    pf_tex *t = (pf_tex *)tex;
    return t ? t->surf : NULL;
}

/* ---- textures (texture.c:30-123) -------------------------------------------------------------- */

static int format_valid(PFpixelformat f, PFdatatype t) { return f <= PF_BGRA && t <= PF_DOUBLE; }

PFtexture pfGenTexture(void *pixels, PFsizei width, PFsizei height, PFpixelformat format, PFdatatype type)
{
    if (!format_valid(format, type)) { if (pf_cur) pf_cur->errCode = PF_INVALID_ENUM; return NULL; }
    pf_tex *t = (pf_tex *)PF_CALLOC(1, sizeof *t);
    if (!t) { if (pf_cur) pf_cur->errCode = PF_ERROR_OUT_OF_MEMORY; return NULL; }
    t->pixels = pixels; t->w = width; t->h = height; t->format = format; t->type = type;
    t->wrap = PF_REPEAT; t->filter = PF_NEAREST;
    return t;
}

void pfDeleteTexture(PFtexture *texture, PFboolean freeBuffer)
{
    pf_tex *t = (pf_tex *)*texture;
    if (!t) return;
    if (pf_cur) { pfh_flush(pf_cur); if (pf_cur->currentTexture == t) { pf_cur->currentTexture = NULL; pf_cur->state_dirty = 1; } }
    pfcu_finish();
    if (t->dev) pfcu_texture_destroy(t->dev);
    if (t->surf) pfh_surf_destroy(t->surf);
    if (freeBuffer && t->pixels) PF_FREE(t->pixels);
    PF_FREE(t);
    *texture = NULL;
}

PFboolean pfIsValidTexture(const PFtexture texture)
{
    const pf_tex *t = (const pf_tex *)texture;
    return t && t->pixels && t->w > 0 && t->h > 0;
}

void pfTextureParameter(PFtexture texture, PFtexturewrap wrapMode, PFtexturefilter filterMode)
{
    pf_tex *t = (pf_tex *)texture;
    if (wrapMode > PF_CLAMP_TO_EDGE || filterMode > PF_BILINEAR) { if (pf_cur) pf_cur->errCode = PF_INVALID_ENUM; return; }
    t->wrap = wrapMode; t->filter = filterMode;
    if (pf_cur) pf_cur->state_dirty = 1;
}

void *pfGetTexturePixels(const PFtexture texture, PFsizei *width, PFsizei *height, PFpixelformat *format, PFdatatype *type)
{
    pf_tex *t = (pf_tex *)texture;
    if (!t) return NULL;
    if (width) *width = t->w;
    if (height) *height = t->h;
    if (format) *format = t->format;
    if (type) *type = t->type;
    if (t->surf) pfh_sync_surface(pf_cur, t->surf);     /* caller is about to read the pixels */
    return t->pixels;
}

/* ---- framebuffer objects (framebuffer.c:29-148) ----------------------------------------------- */

PFframebuffer pfGenFramebuffer(PFsizei width, PFsizei height, PFpixelformat format, PFdatatype type)
{
    PFframebuffer fb = { NULL, NULL };
    PFsizei size = width * height;
    void *pixels = PF_CALLOC(size, pfh_pixel_bytes(format, type));
    if (!pixels) return fb;
    PFtexture tex = pfGenTexture(pixels, width, height, format, type);
    if (!tex) { if (pf_cur) pf_cur->errCode = PF_INVALID_ENUM; PF_FREE(pixels); return fb; }
    PFfloat *z = (PFfloat *)PF_MALLOC(size * sizeof(PFfloat));
    if (!z) { if (pf_cur) pf_cur->errCode = PF_ERROR_OUT_OF_MEMORY; pfDeleteTexture(&tex, PF_TRUE); return fb; }
    for (PFsizei i = 0; i < size; i++) z[i] = FLT_MAX;
    if (!pfh_surf_create((pf_tex *)tex, z, 1)) {
        if (pf_cur) pf_cur->errCode = PF_INVALID_OPERATION;
        PF_FREE(z); pfDeleteTexture(&tex, PF_TRUE); return fb;
    }
    fb.texture = tex; fb.zbuffer = z;
    return fb;
}

void pfDeleteFramebuffer(PFframebuffer *framebuffer)
{
    if (!framebuffer) return;
    if (pf_cur && framebuffer->texture && pf_cur->cur_surf == pfh_surf_lookup(framebuffer->texture)) {
        pfh_flush(pf_cur);
        pf_cur->cur_surf = pf_cur->main_surf;
        if (pf_cur->bindedFramebuffer == framebuffer) pf_cur->bindedFramebuffer = NULL;
    }
    if (framebuffer->texture) pfDeleteTexture(&framebuffer->texture, PF_TRUE);
    if (framebuffer->zbuffer) PF_FREE(framebuffer->zbuffer);
    framebuffer->texture = NULL; framebuffer->zbuffer = NULL;
}

PFboolean pfIsValidFramebuffer(PFframebuffer *framebuffer)
{
    const pf_tex *t = (const pf_tex *)framebuffer->texture;
    return t && framebuffer->zbuffer && t->w > 0 && t->h > 0 && pfIsValidTexture(framebuffer->texture);
}

void pfClearFramebuffer(PFframebuffer *framebuffer, PFcolor color, PFfloat depth)
{
    pf_surf *s = pfh_surf_lookup(framebuffer->texture);
    if (!s) return;
    if (pf_cur && pf_cur->cur_surf == s) pfh_flush(pf_cur);
    uint32_t rgba; memcpy(&rgba, &color, 4);
    s->host_newer = 0;                       /* every pixel is overwritten */
    pfcu_surface_fill(s->dev, 1, rgba, 1, depth);
    s->dev_newer = 1; s->dirty_y0 = 0; s->dirty_y1 = s->tex->h; s->readback_queued = 0;
    if (pf_cur && !pfh_sync_mode_explicit()) pfh_sync_surface(pf_cur, s);
}

PFcolor pfGetFramebufferPixel(const PFframebuffer *framebuffer, PFsizei x, PFsizei y)
{
    pf_tex *t = (pf_tex *)framebuffer->texture;
    if (t->surf) pfh_sync_surface(pf_cur, t->surf);
    return pfh_pixel_get(t, (size_t)y * t->w + x);
}

PFfloat pfGetFramebufferDepth(const PFframebuffer *framebuffer, PFsizei x, PFsizei y)
{
    pf_tex *t = (pf_tex *)framebuffer->texture;
    if (t->surf) pfh_sync_surface(pf_cur, t->surf);
    return framebuffer->zbuffer[(size_t)y * t->w + x];
}

static int depth_cmp(PFdepthmode m, float s, float d)     /* scalar table, depth.h:28-73 */
{
    switch (m) {
    case PF_EQUAL: return s == d;   case PF_NOTEQUAL: return s != d;
    case PF_LESS: return s < d;     case PF_LEQUAL: return s <= d;
    case PF_GREATER: return s > d;  default: return s >= d;
    }
}

static void host_write_begin(pf_surf *s) { if (s) pfh_sync_surface(pf_cur, s); }
static void host_write_end(pf_surf *s) { if (s) s->host_newer = 1; }

void pfSetFramebufferPixelDepthTest(PFframebuffer *framebuffer, PFsizei x, PFsizei y, PFfloat z, PFcolor color, PFdepthmode depthMode)
{
    if (depthMode > PF_GEQUAL) { if (pf_cur) pf_cur->errCode = PF_INVALID_ENUM; return; }
    pf_tex *t = (pf_tex *)framebuffer->texture;
    host_write_begin(t->surf);
    size_t off = (size_t)y * t->w + x;
    if (depth_cmp(depthMode, z, framebuffer->zbuffer[off])) {
        pfh_pixel_set(t, off, color);
        framebuffer->zbuffer[off] = z;
        host_write_end(t->surf);
    }
}

void pfSetFramebufferPixelDepth(PFframebuffer *framebuffer, PFsizei x, PFsizei y, PFfloat z, PFcolor color)
{
    pf_tex *t = (pf_tex *)framebuffer->texture;
    host_write_begin(t->surf);
    size_t off = (size_t)y * t->w + x;
    pfh_pixel_set(t, off, color);
    framebuffer->zbuffer[off] = z;
    host_write_end(t->surf);
}

void pfSetFramebufferPixel(PFframebuffer *framebuffer, PFsizei x, PFsizei y, PFcolor color)
{
    pf_tex *t = (pf_tex *)framebuffer->texture;
    host_write_begin(t->surf);
    pfh_pixel_set(t, (size_t)y * t->w + x, color);
    host_write_end(t->surf);
}

/* ---- render lists (renderlist.c:14-97) --------------------------------------------------------- */

static void fvec_push(pf_fvec *v, const void *src)
{
    if (v->size == v->cap) {
        size_t nc = v->cap ? v->cap * 2 : 8;
        float *p = (float *)realloc(v->data, nc * v->elem * sizeof(float));
        if (!p) return;
        v->data = p; v->cap = nc;
    }
    memcpy(v->data + v->size * v->elem, src, v->elem * sizeof(float));
    v->size++;
}

static void call_free(pf_call *c)
{
    free(c->positions.data); free(c->texcoords.data); free(c->normals.data); free(c->colors.data);
    memset(c, 0, sizeof *c);
}

static void backup_make(pf_ctx *c)       /* internal/context/context.c:25-35 */
{
    pf_backup *b = &c->backup;
    memcpy(b->material, c->material, sizeof b->material);
    memcpy(b->texcoord, c->currentTexcoord, sizeof b->texcoord);
    memcpy(b->normal, c->currentNormal, sizeof b->normal);
    b->color = c->currentColor; b->texture = c->currentTexture; b->state = c->state;
}

static void backup_restore(pf_ctx *c)    /* internal/context/context.c:37-47 */
{
    pf_backup *b = &c->backup;
    memcpy(c->material, b->material, sizeof b->material);
    memcpy(c->currentTexcoord, b->texcoord, sizeof b->texcoord);
    memcpy(c->currentNormal, b->normal, sizeof b->normal);
    c->currentColor = b->color; c->currentTexture = (pf_tex *)b->texture; c->state = b->state;
    c->state_dirty = 1;
}

PFrenderlist pfGenList(void)
{
    if (!pf_cur) return NULL;
    pf_list *l = (pf_list *)PF_CALLOC(1, sizeof *l);
    if (!l) pf_cur->errCode = PF_ERROR_OUT_OF_MEMORY;
    return l;
}

void pfDeleteList(PFrenderlist *renderList)
{
    if (!renderList) return;
    pf_list *l = (pf_list *)*renderList;
    if (l) {
        if (pf_cur) pfh_flush(pf_cur);
        if (pfh_lists_pending()) pfh_lists_flush_all(pf_cur);      /* queued replays refer to the device copies */
        pfh_list_release_device(l);
        for (size_t i = 0; i < l->size; i++) call_free(&l->calls[i]);
        free(l->calls);
        PF_FREE(l);
    }
    *renderList = NULL;
}

void pfNewList(PFrenderlist renderList)
{
    pf_ctx *c = pf_cur;
    if (!renderList) { c->errCode = PF_INVALID_VALUE; return; }
    pf_list *l = (pf_list *)renderList;
    pfh_flush(c);
    if (pfh_lists_pending()) pfh_lists_flush_all(c);
    pfh_list_release_device(l);
    for (size_t i = 0; i < l->size; i++) call_free(&l->calls[i]);
    l->size = 0;
    c->recording = l;
    backup_make(c);
}

void pfEndList(void)
{
    pf_ctx *c = pf_cur;
    if (!c->recording) c->errCode = PF_INVALID_OPERATION;
    c->recording = NULL;
    backup_restore(c);
}

/* used by pfBegin / pfVertex4fv while recording (context.c:1590-1602,1675-1684) */
void pfh_list_begin(pf_ctx *c, PFdrawmode mode)
{
    pf_list *l = c->recording;
    if (l->size == l->cap) {
        size_t nc = l->cap ? l->cap * 2 : 4;
        pf_call *p = (pf_call *)realloc(l->calls, nc * sizeof *p);
        if (!p) { c->errCode = PF_ERROR_OUT_OF_MEMORY; return; }
        l->calls = p; l->cap = nc;
    }
    pf_call *k = &l->calls[l->size++];
    memset(k, 0, sizeof *k);
    k->positions.elem = 4; k->texcoords.elem = 2; k->normals.elem = 3; k->colors.elem = 1;
    memcpy(k->material, c->material, sizeof k->material);
    k->texture = c->currentTexture; k->mode = mode;
}

void pfh_list_vertex(pf_ctx *c, const PFfloat *v)
{
    pf_list *l = c->recording;
    if (l->size == 0) return;
    pf_call *k = &l->calls[l->size - 1];
    fvec_push(&k->positions, v);
    fvec_push(&k->texcoords, c->currentTexcoord);
    fvec_push(&k->normals, c->currentNormal);
    fvec_push(&k->colors, &c->currentColor);
}

void pfCallList(const PFrenderlist renderList)
{
    pf_ctx *c = pf_cur;
    pf_list *l = (pf_list *)renderList;
    if (pfh_call_list_device(c, l)) { pfh_end_of_draw(c); return; }     /* replayed from its device-resident form */
    backup_make(c);
    int outer = c->replaying;
    c->replaying = 1;
    for (size_t i = 0; i < l->size; i++) {
        const pf_call *k = &l->calls[i];
        memcpy(c->material, k->material, sizeof c->material);
        c->state_dirty = 1;
        pfBindTexture(k->texture);
        pfBegin(k->mode);
        for (size_t j = 0; j < k->positions.size; j++) {
            pfColor4ubv((const PFubyte *)(k->colors.data + j));
            pfTexCoordfv(k->texcoords.data + 2 * j);
            pfNormal3fv(k->normals.data + 3 * j);
            pfVertex4fv(k->positions.data + 4 * j);
        }
        pfEnd();
    }
    c->replaying = outer;
    backup_restore(c);
    pfh_end_of_draw(c);
}
