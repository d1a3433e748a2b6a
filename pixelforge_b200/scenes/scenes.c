/*
 * scenes.c - synthetic scenes for parity tests and benchmarks, written against pixelforge.h ONLY,
 * so the very same translation unit is compiled three times:
 *   - against the product            (libpfscenes_cuda.so,   -DPFSCENE_HAVE_PFX)
 *   - against front end + C oracle   (libpfscenes_oracle.so, -DPFSCENE_HAVE_PFX)   tests only
 *   - against the unmodified reference library and header (libpfscenes_ref*.so)    tests / CPU baseline
 * Scene definitions follow BASELINE.json configs C1..C5 and SURVEY.md 8-d; all inputs are
 * generated from an LCG (s <- 1664525 s + 1013904223), no files are read.
 */
#include "pixelforge.h"
#ifdef PFSCENE_HAVE_PFX
#include "pfx.h"
#endif

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define SCN_PI 3.14159265358979323846

#if defined(__GNUC__)
#define SCN_API __attribute__((visibility("default")))
#else
#define SCN_API
#endif

typedef struct {
    int width, height;
    int frames;         /* timed frames                                                           */
    int warmup;         /* untimed frames before them                                             */
    int variant;        /* scene specific bit field                                               */
    int size;           /* scene specific size (grid resolution, layers, contexts ...)            */
    int seed;
    int explicit_sync;  /* product only: 1 = PF_CUDA_SYNC=explicit + pfxFinish per frame          */
    int first_frame;    /* animation frame index of the first rendered frame                      */
} pfscene_cfg;

typedef struct {
    double ms_total, ms_min, ms_median;
    unsigned long long triangles_submitted, triangles_rasterised, pixels_shaded, pixels_depth_failed, kernel_launches;
    unsigned long long api_triangles;   /* triangles the scene asked for (per frame)              */
} pfscene_result;

/* ---- helpers ---------------------------------------------------------------------------------- */

static uint32_t lcg_state;
static uint32_t lcg(void) { lcg_state = 1664525u * lcg_state + 1013904223u; return lcg_state; }
static float lcgf(void) { return (float)(lcg() >> 8) * (1.0f / 16777216.0f); }

static double now_ms(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void finish(void)
{
#ifdef PFSCENE_HAVE_PFX
    pfxFinish();
#endif
}

static float *g_depth_dst; static int g_depth_w;
static PFcolor grab_depth(PFint x, PFint y, PFfloat depth, PFcolor c) { g_depth_dst[(size_t)y * g_depth_w + x] = depth; return c; }

static void read_depth(float *dst, int w)
{
    if (!dst) return;
#ifdef PFSCENE_HAVE_PFX
    (void)w; pfxReadDepth(dst);
#else
    g_depth_dst = dst; g_depth_w = w; pfPostProcess(grab_depth);
#endif
}

static void cam_perspective(double fovy_deg, double aspect, double zn, double zf)
{
    double top = zn * tan(fovy_deg * 0.5 * SCN_PI / 180.0), right = top * aspect;
    pfMatrixMode(PF_PROJECTION); pfLoadIdentity();
    pfFrustum((PFfloat)-right, (PFfloat)right, (PFfloat)-top, (PFfloat)top, (PFfloat)zn, (PFfloat)zf);
    pfMatrixMode(PF_MODELVIEW); pfLoadIdentity();
}

static void cam_lookat(const float eye[3], const float at[3])
{
    float f[3] = { eye[0] - at[0], eye[1] - at[1], eye[2] - at[2] };
    float l = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]); f[0] /= l; f[1] /= l; f[2] /= l;
    float up[3] = { 0, 1, 0 };
    float r[3] = { up[1] * f[2] - up[2] * f[1], up[2] * f[0] - up[0] * f[2], up[0] * f[1] - up[1] * f[0] };
    l = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]); r[0] /= l; r[1] /= l; r[2] /= l;
    float u[3] = { f[1] * r[2] - f[2] * r[1], f[2] * r[0] - f[0] * r[2], f[0] * r[1] - f[1] * r[0] };
    float m[16] = { r[0], u[0], f[0], 0, r[1], u[1], f[1], 0, r[2], u[2], f[2], 0,
                    -(r[0] * eye[0] + r[1] * eye[1] + r[2] * eye[2]), -(u[0] * eye[0] + u[1] * eye[1] + u[2] * eye[2]),
                    -(f[0] * eye[0] + f[1] * eye[1] + f[2] * eye[2]), 1 };
    pfMatrixMode(PF_MODELVIEW); pfLoadIdentity(); pfMultMatrixf(m);
}

static void ortho2d(int w, int h)
{
    pfViewport(0, 0, (PFsizei)w, (PFsizei)h);
    pfMatrixMode(PF_PROJECTION); pfLoadIdentity();
    pfOrtho(0.0f, (PFfloat)w, (PFfloat)h, 0.0f, 0.0f, 1.0f);
    pfMatrixMode(PF_MODELVIEW); pfLoadIdentity();
}

static uint8_t *make_texture(int w, int h, int comps, uint32_t seed, int lo, int hi, int alo, int ahi)
{
    /* Two zeroed guard rows: with CLAMP_TO_EDGE / MIRRORED_REPEAT the reference rounds
       v*(h-1)+0.5 half-to-even and can address row h (sampler.h:220-255), i.e. it reads past the
       texture.  The guard rows make that read deterministic (0) for the reference; the product and
       the oracle return 0 for any out-of-range texel index. */
    uint8_t *t = (uint8_t *)calloc((size_t)w * (h + 2) * comps + 16, 1);
    lcg_state = seed;
    for (size_t i = 0; i < (size_t)w * h; i++) {
        uint32_t r = lcg();
        for (int c = 0; c < comps; c++) {
            int v = (int)((r >> (8 * c)) & 255u);
            if (c < 3) v = lo + v % (hi - lo + 1); else v = alo + v % (ahi - alo + 1);
            t[i * comps + c] = (uint8_t)v;
        }
    }
    return t;
}

/* A texture in any (format, type) pair the reference has texel getters for ("texfmt" scene).  Byte and packed types take
   random bits; half and float components lie in [0, 1] with a few above 1 and below 0 mixed in (the reference's getters
   OR the converted channels together without masking, pixel.h:2652-3040: those texels smear into their neighbours). */
static uint8_t *make_texture_pair(int w, int h, PFpixelformat f, PFdatatype t, uint32_t seed, int *texel_bytes)
{
    const int comps = f <= PF_LUMINANCE ? 1 : (f == PF_LUMINANCE_ALPHA ? 2 : ((f == PF_RGB || f == PF_BGR) ? 3 : 4));
    const int bytes = t == PF_UNSIGNED_BYTE ? comps : (t == PF_HALF_FLOAT ? 2 * comps : (t == PF_FLOAT ? 4 * comps : 2));
    uint8_t *px = (uint8_t *)calloc((size_t)w * (h + 2) * bytes + 16, 1);
    lcg_state = seed;
    const size_t n = (size_t)w * h;
    if (t == PF_HALF_FLOAT) {
        uint16_t *p = (uint16_t *)px;
        for (size_t i = 0; i < n * comps; i++) {
            uint32_t r = lcg() >> 8;
            uint16_t v = (uint16_t)(r % 0x3C01u);                     /* 0 .. 1.0 */
            if ((r >> 16) % 61u == 0) v = 0x3E00;                     /* 1.5 */
            else if ((r >> 16) % 67u == 0) v |= 0x8000u;              /* negative */
            p[i] = v;
        }
    } else if (t == PF_FLOAT) {
        float *p = (float *)px;
        for (size_t i = 0; i < n * comps; i++) {
            uint32_t r = lcg() >> 8;
            float v = (float)(r & 0xFFFFu) / 65535.0f;
            if ((r >> 16) % 61u == 0) v += 1.0f;
            else if ((r >> 16) % 67u == 0) v = -v;
            p[i] = v;
        }
    } else {
        for (size_t i = 0; i < n * bytes; i++) px[i] = (uint8_t)(lcg() >> 24);
    }
    *texel_bytes = bytes;
    return px;
}

/* ---- "conform": the corners of the public API no other scene touches ------------------------------ */
/* Every pfGet*v getter over every PFgettable and PFstate (plus invalid names), pfIsEnabled / pfIsEnabledLight / pfGetError,
   all pfColor* / pfVertex* / pfRasterPos* / pfRect* argument variants, pfFogfv, the framebuffer pixel accessors
   (pfSetFramebufferPixel / ...Depth / ...DepthTest, pfGetFramebufferPixel / ...Depth, pfClearFramebuffer, pfIsValid*),
   pfGetTexturePixels, pfGetCurrentContext.  What the getters return is written word by word into a 128x64 framebuffer
   object through pfSetFramebufferPixel and drawn onto the target with pfDrawPixels, so that the ordinary colour / depth
   comparison against the reference covers it.  variant bit 0: blending + depth test while the variants draw. */
/* exported by the reference library (context.c:1928-1936) but missing from its header */
extern void pfRecti(PFint x1, PFint y1, PFint x2, PFint y2);
extern void pfRectiv(const PFint *v1, const PFint *v2);
typedef struct { PFframebuffer *fb; uint32_t n; } conf_log;
static void conf_word(conf_log *l, uint32_t wd)
{
    if (l->n >= 128u * 60u) return;
    PFcolor c; memcpy(&c, &wd, 4);
    pfSetFramebufferPixel(l->fb, (PFsizei)(l->n % 128u), (PFsizei)(l->n / 128u), c);
    l->n++;
}
static void conf_words(conf_log *l, const void *p, int nwords)
{
    const uint32_t *u = (const uint32_t *)p;
    for (int i = 0; i < nwords; i++) conf_word(l, u[i]);
}
static void conf_getters(conf_log *l, PFenum name)
{
    PFint iv[16]; PFfloat fv[16]; PFdouble dv[16]; PFboolean bv[4];
    memset(iv, 0, sizeof iv); memset(fv, 0, sizeof fv); memset(dv, 0, sizeof dv); memset(bv, 0, sizeof bv);
    pfGetIntegerv(name, iv); conf_word(l, (uint32_t)pfGetError()); conf_words(l, iv, 4);
    pfGetFloatv(name, fv);   conf_word(l, (uint32_t)pfGetError()); conf_words(l, fv, 16);
    pfGetDoublev(name, dv);  conf_word(l, (uint32_t)pfGetError()); conf_words(l, dv, 32);
    pfGetBooleanv(name, bv); conf_word(l, (uint32_t)pfGetError()); conf_word(l, (uint32_t)(bv[0] != 0));
}

static void conform_scene(const pfscene_cfg *cfg, PFtexture tex, PFframebuffer *fbo, uint8_t *aux, uint8_t *target, PFpixelformat tfmt)
{
    const int v = cfg->variant, w = cfg->width, h = cfg->height;
    static float varr[12]; static uint8_t carr[16];
    conf_log log = { fbo, 0 };
    lcg_state = (uint32_t)cfg->seed * 22695477u + 1u;
    pfClearFramebuffer(fbo, (PFcolor){ 1, 2, 3, 4 }, 0.5f);
    pfClearColor(12, 34, 56, 255);
    pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));

    /* 1. state with a distinct value everywhere, then every getter */
    pfViewport(3, 5, (PFsizei)(w - 10), (PFsizei)(h - 9));
    pfClearDepth(0.75f);
    pfEnable(PF_CULL_FACE); pfCullFace(PF_FRONT);
    pfColor4us(0x1234, 0x5678, 0x9abc, 0xdef0);
    pfNormal3f(0.25f, -0.5f, 0.75f);
    pfTexCoord2f(0.125f, 0.625f);
    pfRasterPos4f(7.5f, 9.25f, 0.5f, 2.0f);
    pfBlendFunc(PF_BLEND_SCREEN); pfDepthFunc(PF_GEQUAL);
    pfPolygonMode(PF_FRONT, PF_LINE); pfPolygonMode(PF_BACK, PF_POINT);
    pfPointSize(3.5f); pfLineWidth(2.25f);
    pfShadeModel(PF_FLAT);
    pfMatrixMode(PF_PROJECTION); pfLoadIdentity(); pfFrustum(-1.0f, 1.5f, -0.75f, 0.5f, 1.0f, 40.0f);
    pfMatrixMode(PF_TEXTURE); pfLoadIdentity(); pfScalef(2.0f, 0.5f, 1.0f); pfTranslatef(0.25f, 0.5f, 0.0f);
    pfMatrixMode(PF_MODELVIEW); pfLoadIdentity(); pfTranslatef(1.0f, -2.0f, -5.0f); pfRotatef(33.0f, 0.3f, 0.5f, 0.7f);
    pfPushMatrix(); pfScalef(1.5f, 0.5f, 2.0f);
    pfVertexPointer(3, PF_FLOAT, 12, varr); pfNormalPointer(PF_FLOAT, 12, varr); pfTexCoordPointer(PF_FLOAT, 8, varr);
    pfColorPointer(4, PF_UNSIGNED_BYTE, 4, carr);
    pfEnable(PF_VERTEX_ARRAY); pfEnable(PF_COLOR_ARRAY);
    pfPixelZoom(1.5f, -2.0f);
    pfEnable(PF_TEXTURE_2D); pfBindTexture(tex); pfEnable(PF_NORMALIZE); pfEnable(PF_COLOR_MATERIAL);
    pfEnableLight(PF_LIGHT3);
    conf_word(&log, (uint32_t)pfGetError());
    for (int name = PF_VIEWPORT - 1; name <= PF_ZOOM_Y + 1; name++) conf_getters(&log, (PFenum)name);
    for (int bit = 0; bit < 13; bit++) {
        PFboolean b = 0;
        pfGetBooleanv((PFenum)(1 << bit), &b); conf_word(&log, (uint32_t)pfGetError()); conf_word(&log, (uint32_t)(b != 0));
        conf_word(&log, (uint32_t)(pfIsEnabled((PFstate)(1 << bit)) != 0));
    }
    for (int li = 0; li < 8; li++) conf_word(&log, (uint32_t)(pfIsEnabledLight((PFsizei)li) != 0));
    {
        const void *ptr = NULL;
        pfGetPointerv(PF_TEXTURE_2D, &ptr);  conf_word(&log, (uint32_t)(ptr == (const void *)tex));
        ptr = (const void *)&log; pfGetPointerv(PF_FRAMEBUFFER, &ptr); conf_word(&log, (uint32_t)(ptr == NULL));
        pfGetPointerv(PF_BLEND_FUNC, &ptr);  conf_word(&log, (uint32_t)(ptr != NULL)); conf_word(&log, (uint32_t)pfGetError());
        pfGetPointerv(PF_DEPTH_FUNC, &ptr);  conf_word(&log, (uint32_t)(ptr != NULL)); conf_word(&log, (uint32_t)pfGetError());
        pfGetPointerv(PF_VIEWPORT, &ptr);    conf_word(&log, (uint32_t)pfGetError());
        PFsizei tw = 0, th = 0; PFpixelformat tf = PF_RED; PFdatatype tt = PF_FLOAT;
        void *px = pfGetTexturePixels(tex, &tw, &th, &tf, &tt);
        conf_word(&log, (uint32_t)tw); conf_word(&log, (uint32_t)th); conf_word(&log, (uint32_t)tf); conf_word(&log, (uint32_t)tt);
        conf_words(&log, px, 4);
        conf_word(&log, (uint32_t)(pfIsValidTexture(tex) != 0)); conf_word(&log, (uint32_t)(pfIsValidFramebuffer(fbo) != 0));
        PFframebuffer none = { NULL, NULL };
        conf_word(&log, (uint32_t)(pfIsValidFramebuffer(&none) != 0));
        conf_word(&log, (uint32_t)(pfGetCurrentContext() != NULL));
    }
    /* errors are sticky until fetched, and fetching clears them */
    pfBlendFunc((PFblendmode)77); pfDepthFunc((PFdepthmode)99); pfCullFace((PFface)17);
    conf_word(&log, (uint32_t)pfGetError()); conf_word(&log, (uint32_t)pfGetError());
    pfPopMatrix(); pfPopMatrix(); pfPopMatrix();
    conf_word(&log, (uint32_t)pfGetError());
    pfLightf((PFsizei)99, PF_SHININESS, 1.0f); conf_word(&log, (uint32_t)pfGetError());
    pfMaterialf(PF_FRONT, PF_POSITION, 1.0f); conf_word(&log, (uint32_t)pfGetError());
    {
        PFfloat fc[4] = { 0.2f, 0.4f, 0.6f, 0.8f }, fs = 1.5f, fe = 4.0f, fd = 1.0f;
        pfFogfv(PF_FOG_COLOR, fc); pfFogfv(PF_FOG_START, &fs); pfFogfv(PF_FOG_END, &fe); pfFogfv(PF_FOG_DENSITY, &fd);
        pfFogfv((PFfogparam)55, &fd); conf_word(&log, (uint32_t)pfGetError());
    }

    /* 2. back to a plain 2D state; one small rectangle / quad per argument variant */
    pfDisable(PF_VERTEX_ARRAY); pfDisable(PF_COLOR_ARRAY);
    pfDisable(PF_TEXTURE_2D); pfDisable(PF_NORMALIZE); pfDisable(PF_COLOR_MATERIAL); pfDisableLight(PF_LIGHT3);
    pfPolygonMode(PF_FRONT_AND_BACK, PF_FILL); pfShadeModel(PF_SMOOTH); pfPointSize(1.0f); pfLineWidth(1.0f);
    pfMatrixMode(PF_TEXTURE); pfLoadIdentity();
    ortho2d(w, h);
    pfDisable(PF_CULL_FACE); pfCullFace(PF_BACK);
    if (v & 1) { pfEnable(PF_BLEND); pfBlendFunc(PF_BLEND_ALPHA); pfEnable(PF_DEPTH_TEST); pfDepthFunc(PF_LEQUAL); }
    else { pfDisable(PF_BLEND); pfDisable(PF_DEPTH_TEST); pfBlendFunc(PF_BLEND_ALPHA); pfDepthFunc(PF_LESS); }
    for (int k = 0; k < 17; k++) {
        const PFubyte ub[4] = { (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(128 + (lcg() >> 25)) };
        const PFushort us[4] = { (PFushort)(lcg() >> 16), (PFushort)(lcg() >> 16), (PFushort)(lcg() >> 16), (PFushort)(0x8000u | (lcg() >> 17)) };
        const PFuint ui[4] = { lcg(), lcg(), lcg(), 0x80000000u | lcg() };
        const PFfloat fl[4] = { lcgf(), lcgf(), lcgf(), 0.5f + 0.5f * lcgf() };
        switch (k) {
        case 0: pfColor1ui(ui[0] | 0xc0000000u); break;        case 1: pfColor3ub(ub[0], ub[1], ub[2]); break;
        case 2: pfColor3ubv(ub); break;                         case 3: pfColor3us(us[0], us[1], us[2]); break;
        case 4: pfColor3usv(us); break;                         case 5: pfColor3ui(ui[0], ui[1], ui[2]); break;
        case 6: pfColor3uiv(ui); break;                         case 7: pfColor3f(fl[0], fl[1], fl[2]); break;
        case 8: pfColor3fv(fl); break;                          case 9: pfColor4ub(ub[0], ub[1], ub[2], ub[3]); break;
        case 10: pfColor4ubv(ub); break;                        case 11: pfColor4us(us[0], us[1], us[2], us[3]); break;
        case 12: pfColor4usv(us); break;                        case 13: pfColor4ui(ui[0], ui[1], ui[2], ui[3]); break;
        case 14: pfColor4uiv(ui); break;                        case 15: pfColor4f(fl[0], fl[1], fl[2], fl[3]); break;
        default: pfColor4fv(fl); break;
        }
        PFint ci[4]; memset(ci, 0, sizeof ci); pfGetIntegerv(PF_CURRENT_COLOR, ci); conf_words(&log, ci, 4);
        const int x0 = 6 + (k % 6) * 24, y0 = 70 + (k / 6) * 14;
        const PFshort s1[2] = { (PFshort)x0, (PFshort)y0 }, s2[2] = { (PFshort)(x0 + 18), (PFshort)(y0 + 10) };
        const PFint i1[2] = { x0, y0 }, i2[2] = { x0 + 18, y0 + 10 };
        const PFfloat f1[2] = { (PFfloat)x0 + 0.25f, (PFfloat)y0 + 0.5f }, f2[2] = { (PFfloat)x0 + 18.5f, (PFfloat)y0 + 10.25f };
        switch (k % 6) {
        case 0: pfRects(s1[0], s1[1], s2[0], s2[1]); break;     case 1: pfRectsv(s1, s2); break;
        case 2: pfRecti(i1[0], i1[1], i2[0], i2[1]); break;     case 3: pfRectiv(i1, i2); break;
        case 4: pfRectf(f1[0], f1[1], f2[0], f2[1]); break;     default: pfRectfv(f1, f2); break;
        }
    }
    /* vertex argument variants: nine overlapping quads (w = 1 or 2: the 4-component forms divide) */
    for (int k = 0; k < 9; k++) {
        const int x0 = 10 + k * 16, y0 = 118;
        pfColor4ub((PFubyte)(40 + 20 * k), (PFubyte)(250 - 25 * k), (PFubyte)(90 + 10 * k), 200);
        pfBegin(PF_QUADS);
        for (int c = 0; c < 4; c++) {
            const int xi = x0 + ((c == 1 || c == 2) ? 22 : 0), yi = y0 + ((c >= 2) ? 20 : 0);
            const PFfloat f4[4] = { (PFfloat)xi + 0.5f, (PFfloat)yi + 0.25f, -0.5f, 1.0f };
            switch (k) {
            case 0: pfVertex2i(xi, yi); break;                  case 1: pfVertex2f(f4[0], f4[1]); break;
            case 2: pfVertex2fv(f4); break;                     case 3: pfVertex3i(xi, yi, 0); break;
            case 4: pfVertex3f(f4[0], f4[1], f4[2]); break;     case 5: pfVertex3fv(f4); break;
            case 6: pfVertex4i(xi, yi, 0, 1); break;            case 7: pfVertex4f(f4[0], f4[1], f4[2], f4[3]); break;
            default: pfVertex4fv(f4); break;
            }
        }
        pfEnd();
    }
    /* raster position variants, each followed by a 6x5 pfDrawPixels */
    for (size_t i = 0; i < 6u * 5u * 4u; i++) aux[i] = (uint8_t)(lcg() >> 24);
    for (int k = 0; k < 9; k++) {
        const int xi = 8 + k * 15, yi = 40;
        const PFfloat f4[4] = { (PFfloat)xi + 0.75f, (PFfloat)yi + 0.5f, 0.25f, 1.0f };
        switch (k) {
        case 0: pfRasterPos2i(xi, yi); break;                   case 1: pfRasterPos2f(f4[0], f4[1]); break;
        case 2: pfRasterPos2fv(f4); break;                      case 3: pfRasterPos3i(xi, yi, 0); break;
        case 4: pfRasterPos3f(f4[0], f4[1], f4[2]); break;      case 5: pfRasterPos3fv(f4); break;
        case 6: pfRasterPos4i(xi, yi, 0, 1); break;             case 7: pfRasterPos4f(f4[0], f4[1], f4[2], f4[3]); break;
        default: pfRasterPos4fv(f4); break;
        }
        PFfloat rp[4] = { 0, 0, 0, 0 }; pfGetFloatv(PF_CURRENT_RASTER_POSITION, rp); conf_words(&log, rp, 4);
        pfPixelZoom(1.0f, 1.0f);
        pfDrawPixels(6, 5, PF_RGBA, PF_UNSIGNED_BYTE, aux);
    }

    /* 3. framebuffer pixel accessors on the object's last rows (the log never reaches row 60) */
    for (int k = 0; k < 24; k++) {
        const PFsizei x = (PFsizei)(5 + 4 * k), y = (PFsizei)(60 + (k & 3));
        const PFcolor c = { (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24) };
        const PFfloat z = 0.25f + 0.5f * lcgf();          /* around the 0.5 the object was cleared to */
        if (k % 3 == 0) pfSetFramebufferPixelDepth(fbo, x, y, z, c);
        else if (k % 3 == 1) pfSetFramebufferPixelDepthTest(fbo, x, y, z, c, (PFdepthmode)(k % 6));
        else pfSetFramebufferPixel(fbo, x, y, c);
        const PFcolor g = pfGetFramebufferPixel(fbo, x, y); const PFfloat gz = pfGetFramebufferDepth(fbo, x, y);
        conf_words(&log, &g, 1); conf_words(&log, &gz, 1);
    }
    pfSetFramebufferPixelDepthTest(fbo, 1, 61, 0.1f, (PFcolor){ 9, 9, 9, 9 }, (PFdepthmode)42);
    conf_word(&log, (uint32_t)pfGetError());

    /* 3b. double buffering with pfSetMainBuffer (same geometry: the reference keeps the depth buffer and adopts the new
       pixels as they are; a different size would run into upstream's realloc with an element count, context.c:284) */
    {
        uint8_t *second = aux + 4096;
        for (size_t i = 0; i < (size_t)w * h * 4; i++) second[i] = (uint8_t)(lcg() >> 24);
        finish();                                  /* explicit-sync mode: the outgoing buffer is brought up to date at sync points only */
        pfSetMainBuffer(second, (PFsizei)w, (PFsizei)h, tfmt, PF_UNSIGNED_BYTE);
        conf_word(&log, (uint32_t)pfGetError());
        pfEnable(PF_BLEND); pfBlendFunc(PF_BLEND_ADD);
        pfColor4ub(40, 80, 20, 90); pfRecti(30, 30, 90, 60);
        pfDisable(PF_BLEND);
        PFcolor rd[16 * 8]; memset(rd, 0, sizeof rd);
        pfReadPixels(40, 40, 16, 8, PF_RGBA, PF_UNSIGNED_BYTE, rd); conf_words(&log, rd, 16 * 8);
        pfReadPixels(100, 100, 8, 2, PF_RGBA, PF_UNSIGNED_BYTE, rd); conf_words(&log, rd, 8 * 2);
        finish();
        pfSetMainBuffer(target, (PFsizei)w, (PFsizei)h, tfmt, PF_UNSIGNED_BYTE);
        pfSetMainBuffer(NULL, (PFsizei)w, (PFsizei)h, tfmt, PF_UNSIGNED_BYTE);
        conf_word(&log, (uint32_t)pfGetError());
        if (v & 1) { pfEnable(PF_BLEND); pfBlendFunc(PF_BLEND_ALPHA); }
        pfColor4ub(200, 100, 50, 160); pfRecti(150, 150, 190, 200);
    }

    /* 4. the log and the object's depth onto the target */
    {
        PFsizei tw = 0, th = 0; PFpixelformat tf = PF_RED; PFdatatype tt = PF_FLOAT;
        void *px = pfGetTexturePixels(fbo->texture, &tw, &th, &tf, &tt);
        pfDisable(PF_BLEND); pfDisable(PF_DEPTH_TEST);
        pfRasterPos2i(16, 150); pfPixelZoom(1.0f, 1.0f);
        pfDrawPixels(tw, th, tf, tt, px);
        for (int k = 0; k < 24; k++) {
            const PFsizei x = (PFsizei)(5 + 4 * k), y = (PFsizei)(60 + (k & 3));
            const PFfloat gz = pfGetFramebufferDepth(fbo, x, y);
            PFcolor c; memcpy(&c, &gz, 4);
            pfColor(c); pfRecti(20 + 5 * k, 8, 24 + 5 * k, 12);
        }
    }
    pfColor4ub(255, 255, 255, 255);
}

/* ---- C1: gears (call sequence of the reference's Gears demo, examples/SDL2/SDL2_Gears.c:4-194) --- */

static unsigned long long g_api_tris;

static void ring_vertex(double r, double a, double z) { pfVertex3f((PFfloat)(r * cos(a)), (PFfloat)(r * sin(a)), (PFfloat)z); }

static void gear(double inner, double outer, double width, int teeth, double tooth_depth)
{
    const double r0 = inner, r1 = outer - tooth_depth / 2.0, r2 = outer + tooth_depth / 2.0;
    const double da = 2.0 * SCN_PI / teeth / 4.0, hz = width * 0.5;
    const double step = 2.0 * SCN_PI / teeth;

    pfShadeModel(PF_FLAT);
    for (int side = 0; side < 2; side++) {                  /* front (+z) then back (-z) */
        const double z = side ? -hz : hz;
        pfNormal3f(0.0f, 0.0f, side ? -1.0f : 1.0f);
        pfBegin(PF_QUAD_STRIP);                             /* the face disc */
        for (int i = 0; i <= teeth; i++) {
            const double a = i * step;
            if (!side) { ring_vertex(r0, a, z); ring_vertex(r1, a, z); ring_vertex(r0, a, z); ring_vertex(r1, a + 3 * da, z); }
            else       { ring_vertex(r1, a, z); ring_vertex(r0, a, z); ring_vertex(r1, a + 3 * da, z); ring_vertex(r0, a, z); }
        }
        pfEnd();
        pfBegin(PF_QUADS);                                  /* the tooth faces */
        for (int i = 0; i < teeth; i++) {
            const double a = i * step;
            if (!side) { ring_vertex(r1, a, z); ring_vertex(r2, a + da, z); ring_vertex(r2, a + 2 * da, z); ring_vertex(r1, a + 3 * da, z); }
            else       { ring_vertex(r1, a + 3 * da, z); ring_vertex(r2, a + 2 * da, z); ring_vertex(r2, a + da, z); ring_vertex(r1, a, z); }
        }
        pfEnd();
        g_api_tris += (unsigned long long)(teeth + 1) * 2 + (unsigned long long)teeth * 2;
    }
    pfBegin(PF_QUAD_STRIP);                                 /* outward faces of the teeth */
    for (int i = 0; i < teeth; i++) {
        const double a = i * step;
        ring_vertex(r1, a, hz); ring_vertex(r1, a, -hz);
        double u = r2 * cos(a + da) - r1 * cos(a), v = r2 * sin(a + da) - r1 * sin(a), len = sqrt(u * u + v * v);
        u /= len; v /= len;
        pfNormal3f((PFfloat)v, (PFfloat)-u, 0.0f);
        ring_vertex(r2, a + da, hz); ring_vertex(r2, a + da, -hz);
        pfNormal3f((PFfloat)cos(a), (PFfloat)sin(a), 0.0f);
        ring_vertex(r2, a + 2 * da, hz); ring_vertex(r2, a + 2 * da, -hz);
        u = r1 * cos(a + 3 * da) - r2 * cos(a + 2 * da); v = r1 * sin(a + 3 * da) - r2 * sin(a + 2 * da);
        pfNormal3f((PFfloat)v, (PFfloat)-u, 0.0f);
        ring_vertex(r1, a + 3 * da, hz); ring_vertex(r1, a + 3 * da, -hz);
        pfNormal3f((PFfloat)cos(a), (PFfloat)sin(a), 0.0f);
    }
    ring_vertex(r1, 0.0, hz); ring_vertex(r1, 0.0, -hz);
    pfEnd();
    g_api_tris += (unsigned long long)teeth * 8;

    pfShadeModel(PF_SMOOTH);
    pfBegin(PF_QUAD_STRIP);                                 /* bore */
    for (int i = 0; i <= teeth; i++) {
        const double a = i * step;
        pfNormal3f((PFfloat)-cos(a), (PFfloat)-sin(a), 0.0f);
        ring_vertex(r0, a, -hz); ring_vertex(r0, a, hz);
    }
    pfEnd();
    g_api_tris += (unsigned long long)teeth * 2;
}

static void gears_setup(int w, int h)
{
    float pos[3] = { 5.0f, 5.0f, 10.0f }, dir[3];
    float l = sqrtf(150.0f);
    dir[0] = -pos[0] / l; dir[1] = -pos[1] / l; dir[2] = -pos[2] / l;
    pfLightfv(PF_LIGHT0, PF_POSITION, pos);
    pfLightfv(PF_LIGHT0, PF_SPOT_DIRECTION, dir);
    pfEnable(PF_CULL_FACE); pfEnable(PF_LIGHTING); pfEnableLight(PF_LIGHT0); pfEnable(PF_DEPTH_TEST);
    float aspect = (float)h / (float)w;
    pfViewport(0, 0, (PFsizei)w, (PFsizei)h);
    pfMatrixMode(PF_PROJECTION); pfLoadIdentity();
    pfFrustum(-1.0f, 1.0f, -aspect, aspect, 5.0f, 60.0f);
    pfMatrixMode(PF_MODELVIEW); pfLoadIdentity();
    pfTranslatef(0.0f, 0.0f, -40.0f);
}

static void gears_frame(float angle)
{
    pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
    pfEnable(PF_COLOR_MATERIAL);
    pfColorMaterial(PF_FRONT_AND_BACK, PF_AMBIENT_AND_DIFFUSE);
    pfPushMatrix();
    pfRotatef(20.0f, 1.0f, 0.0f, 0.0f); pfRotatef(30.0f, 0.0f, 1.0f, 0.0f); pfRotatef(0.0f, 0.0f, 0.0f, 1.0f);
    pfPushMatrix(); pfTranslatef(-3.0f, -2.0f, 0.0f); pfRotatef(angle, 0.0f, 0.0f, 1.0f);
    pfColor3ub(255, 0, 0); gear(1.0, 4.0, 1.0, 20, 0.7); pfPopMatrix();
    pfPushMatrix(); pfTranslatef(3.1f, -2.0f, 0.0f); pfRotatef(-2.0f * angle - 9.0f, 0.0f, 0.0f, 1.0f);
    pfColor3ub(0, 255, 0); gear(0.5, 2.0, 2.0, 10, 0.7); pfPopMatrix();
    pfPushMatrix(); pfTranslatef(-3.1f, 4.2f, 0.0f); pfRotatef(-2.0f * angle - 25.0f, 0.0f, 0.0f, 1.0f);
    pfColor3ub(0, 0, 255); gear(1.3, 2.0, 0.5, 10, 0.7); pfPopMatrix();
    pfPopMatrix();
    pfDisable(PF_COLOR_MATERIAL);
}

/* ---- C2: textured torus, bilinear/nearest, repeat/clamp/mirror, alpha blend + depth ------------- */
/* variant: bit0 bilinear, bits1-2 wrap mode, bit3 RGB8 texture instead of RGBA8, bit4 no blending */

typedef struct { float *pos, *nrm, *uv; uint32_t *idx; int nverts, nidx; } mesh_t;

/* Vertex and index arrays of the big meshes: page-locked when the library offers it (pfx.h; bench.py's end-to-end leg
 * copies its inputs from pinned host memory), plain malloc for the reference. */
#ifdef PFSCENE_HAVE_PFX
/* PFSCENE_STATIC_ARRAYS=1: the meshes never change after they are built, and the scene says so (pfxHostStatic): the
   library then keeps them in device memory instead of copying them at every draw (bench.py reports both ways) */
static void *mesh_alloc(size_t bytes)
{
    void *p = pfxHostAlloc(bytes);
    const char *e = getenv("PFSCENE_STATIC_ARRAYS");
    if (p && e && e[0] == '1') pfxHostStatic(p, PF_TRUE);
    return p;
}
static void mesh_free(void *p) { pfxHostFree(p); }
#else
static void *mesh_alloc(size_t bytes) { return malloc(bytes); }
static void mesh_free(void *p) { free(p); }
#endif

static mesh_t make_torus(int nu, int nv, float R, float r, float uvscale)
{
    mesh_t m; m.nverts = (nu + 1) * (nv + 1); m.nidx = nu * nv * 6;
    m.pos = (float *)mesh_alloc(sizeof(float) * 3 * m.nverts); m.nrm = (float *)mesh_alloc(sizeof(float) * 3 * m.nverts);
    m.uv = (float *)mesh_alloc(sizeof(float) * 2 * m.nverts); m.idx = (uint32_t *)mesh_alloc(sizeof(uint32_t) * m.nidx);
    for (int i = 0; i <= nu; i++) for (int j = 0; j <= nv; j++) {
        double a = 2.0 * SCN_PI * i / nu, b = 2.0 * SCN_PI * j / nv;
        int k = i * (nv + 1) + j;
        m.pos[3 * k] = (float)((R + r * cos(b)) * cos(a)); m.pos[3 * k + 1] = (float)(r * sin(b)); m.pos[3 * k + 2] = (float)((R + r * cos(b)) * sin(a));
        m.nrm[3 * k] = (float)(cos(b) * cos(a)); m.nrm[3 * k + 1] = (float)sin(b); m.nrm[3 * k + 2] = (float)(cos(b) * sin(a));
        m.uv[2 * k] = uvscale * (float)i / nu * 4.0f - 0.5f * (uvscale - 1.0f); m.uv[2 * k + 1] = uvscale * (float)j / nv - 0.5f * (uvscale - 1.0f);
    }
    int n = 0;
    for (int i = 0; i < nu; i++) for (int j = 0; j < nv; j++) {
        uint32_t a = (uint32_t)(i * (nv + 1) + j), b = a + 1, c = a + (uint32_t)(nv + 1), d = c + 1;
        m.idx[n++] = a; m.idx[n++] = b; m.idx[n++] = c; m.idx[n++] = b; m.idx[n++] = d; m.idx[n++] = c;
    }
    return m;
}

static void free_mesh(mesh_t *m) { mesh_free(m->pos); mesh_free(m->nrm); mesh_free(m->uv); mesh_free(m->idx); }

static void draw_mesh_immediate(const mesh_t *m)
{
    pfBegin(PF_TRIANGLES);
    for (int k = 0; k < m->nidx; k++) {
        uint32_t i = m->idx[k];
        pfNormal3fv(m->nrm + 3 * i); pfTexCoordfv(m->uv + 2 * i); pfVertex3fv(m->pos + 3 * i);
    }
    pfEnd();
    g_api_tris += (unsigned long long)m->nidx / 3;
}

static void draw_mesh_arrays(const mesh_t *m)
{
    pfEnable(PF_VERTEX_ARRAY); pfEnable(PF_NORMAL_ARRAY); pfEnable(PF_TEXTURE_COORD_ARRAY);
    pfVertexPointer(3, PF_FLOAT, 0, m->pos); pfNormalPointer(PF_FLOAT, 0, m->nrm); pfTexCoordPointer(PF_FLOAT, 0, m->uv);
    pfDrawElements(PF_TRIANGLES, (PFsizei)m->nidx, PF_UNSIGNED_INT, m->idx);
    pfDisable(PF_VERTEX_ARRAY); pfDisable(PF_NORMAL_ARRAY); pfDisable(PF_TEXTURE_COORD_ARRAY);
    g_api_tris += (unsigned long long)m->nidx / 3;
}

/* ---- C3: height field, per-pixel Blinn-Phong ----------------------------------------------------- */

static mesh_t make_heightfield(int n)
{
    mesh_t m; m.nverts = (n + 1) * (n + 1); m.nidx = n * n * 6;
    m.pos = (float *)mesh_alloc(sizeof(float) * 3 * m.nverts); m.nrm = (float *)mesh_alloc(sizeof(float) * 3 * m.nverts);
    m.uv = (float *)mesh_alloc(sizeof(float) * 2 * m.nverts); m.idx = (uint32_t *)mesh_alloc(sizeof(uint32_t) * m.nidx);
    for (int j = 0; j <= n; j++) for (int i = 0; i <= n; i++) {
        double x = -2.0 + 4.0 * i / n, y = -1.2 + 2.4 * j / n;
        double z = 0.3 * sin(3 * x) * cos(3 * y);
        double dzdx = 0.9 * cos(3 * x) * cos(3 * y), dzdy = -0.9 * sin(3 * x) * sin(3 * y);
        double l = sqrt(dzdx * dzdx + dzdy * dzdy + 1.0);
        int k = j * (n + 1) + i;
        m.pos[3 * k] = (float)x; m.pos[3 * k + 1] = (float)y; m.pos[3 * k + 2] = (float)z;
        m.nrm[3 * k] = (float)(-dzdx / l); m.nrm[3 * k + 1] = (float)(-dzdy / l); m.nrm[3 * k + 2] = (float)(1.0 / l);
        m.uv[2 * k] = (float)i / n; m.uv[2 * k + 1] = (float)j / n;
    }
    int c = 0;
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
        uint32_t a = (uint32_t)(j * (n + 1) + i), b = a + 1, d = a + (uint32_t)(n + 1), e = d + 1;
        m.idx[c++] = a; m.idx[c++] = b; m.idx[c++] = e; m.idx[c++] = a; m.idx[c++] = e; m.idx[c++] = d;
    }
    return m;
}

/* ---- C4: overdraw -------------------------------------------------------------------------------- */

static void draw_textured_quad(PFtexture tex, float x, float y, float w, float h, float urep, float vrep, int tinted)
{
    pfBindTexture(tex);
    pfBegin(PF_QUADS);
    if (tinted) pfColor4ub(255, 200, 150, 220);
    pfTexCoord2f(0.0f, 0.0f); pfVertex2f(x, y);
    if (tinted) pfColor4ub(200, 255, 180, 255);
    pfTexCoord2f(0.0f, vrep); pfVertex2f(x, y + h);
    if (tinted) pfColor4ub(160, 210, 255, 240);
    pfTexCoord2f(urep, vrep); pfVertex2f(x + w, y + h);
    if (tinted) pfColor4ub(255, 255, 200, 230);
    pfTexCoord2f(urep, 0.0f); pfVertex2f(x + w, y);
    pfEnd();
    pfBindTexture(0);
    g_api_tris += 2;
}

/* ---- micro scenes for parity ---------------------------------------------------------------------- */

static void random_vertex(int w, int h, int with_uv, float zlo, float zhi)
{
    PFcolor c = { (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(64 + ((lcg() >> 24) % 192)) };
    pfColor(c);
    if (with_uv) pfTexCoord2f(lcgf() * 6.0f - 3.0f, lcgf() * 6.0f - 3.0f);
    pfNormal3f(lcgf() - 0.5f, lcgf() - 0.5f, lcgf() + 0.2f);
    pfVertex3f(lcgf() * (w * 1.2f) - 0.1f * w, lcgf() * (h * 1.2f) - 0.1f * h, zlo + lcgf() * (zhi - zlo));
}

/* variant bits: 0-2 blend mode (with bit3: blending on), 4-6 depth func (bit7: depth test on),
   8 flat shading, 9 texture on, 10 bilinear, 11-12 wrap, 13 cull off, 14-16 draw mode selector,
   17 RGB8 texture, 18 perspective camera instead of 2D ortho, 19 Phong lights, 20 render into an FBO
   and composite it, 21 spotlight+attenuation, 22 BGRA texture, 23 BGR8 texture, 24-25 target layout (see pfscene_open) */
static void micro_scene(const pfscene_cfg *cfg, PFtexture tex)
{
    const int v = cfg->variant, w = cfg->width, h = cfg->height;
    pfClearColor(10, 20, 30, 255);
    pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
    if (v & (1 << 18)) {
        cam_perspective(60.0, (double)w / h, 0.1, 100.0);
        pfViewport(0, 0, (PFsizei)w, (PFsizei)h);
        float eye[3] = { 0.3f, 0.4f, 3.0f }, at[3] = { 0, 0, 0 };
        cam_lookat(eye, at);
    } else ortho2d(w, h);
    if (v & 8) { pfEnable(PF_BLEND); pfBlendFunc((PFblendmode)(v & 7)); } else pfDisable(PF_BLEND);
    if (v & 128) { pfEnable(PF_DEPTH_TEST); pfDepthFunc((PFdepthmode)((v >> 4) & 7) > PF_GEQUAL ? PF_LESS : (PFdepthmode)((v >> 4) & 7)); } else pfDisable(PF_DEPTH_TEST);
    pfShadeModel((v & 256) ? PF_FLAT : PF_SMOOTH);
    if (v & 512) {
        pfEnable(PF_TEXTURE_2D);
        pfTextureParameter(tex, (PFtexturewrap)(((v >> 11) & 3) % 3), (v & 1024) ? PF_BILINEAR : PF_NEAREST);
        pfBindTexture(tex);
    } else pfDisable(PF_TEXTURE_2D);
    if (v & (1 << 13)) pfDisable(PF_CULL_FACE); else pfEnable(PF_CULL_FACE);
    if (v & (1 << 19)) {
        float lp[3] = { 1.5f, 2.0f, 2.5f }, lp2[3] = { -2.0f, 0.5f, 1.0f }, ld[3] = { 0.5f, -0.3f, -0.8f };
        float amb[3] = { 0.8f, 0.3f, 0.2f }, spec[3] = { 1.0f, 1.0f, 1.0f };
        pfEnable(PF_LIGHTING); pfLightModel(PF_PHONG);
        pfLightfv(PF_LIGHT0, PF_POSITION, lp); pfEnableLight(PF_LIGHT0);
        pfLightfv(PF_LIGHT1, PF_POSITION, lp2);
        if (v & (1 << 21)) {
            pfLightfv(PF_LIGHT1, PF_SPOT_DIRECTION, ld);
            pfLightf(PF_LIGHT1, PF_SPOT_INNER_CUTOFF, 25.0f); pfLightf(PF_LIGHT1, PF_SPOT_OUTER_CUTOFF, 40.0f);
            pfLightf(PF_LIGHT1, PF_LINEAR_ATTENUATION, 0.05f); pfLightf(PF_LIGHT1, PF_QUADRATIC_ATTENUATION, 0.02f);
        }
        pfEnableLight(PF_LIGHT1);
        pfMaterialfv(PF_FRONT_AND_BACK, PF_AMBIENT_AND_DIFFUSE, amb); pfMaterialfv(PF_FRONT_AND_BACK, PF_SPECULAR, spec);
        pfMaterialf(PF_FRONT_AND_BACK, PF_SHININESS, 32.0f);
    }
    lcg_state = (uint32_t)cfg->seed * 2654435761u + 12345u;
    const int n = cfg->size > 0 ? cfg->size : 64;
    static const PFdrawmode modes[6] = { PF_TRIANGLES, PF_QUADS, PF_TRIANGLE_FAN, PF_TRIANGLE_STRIP, PF_QUAD_FAN, PF_QUAD_STRIP };
    const PFdrawmode mode = modes[((v >> 14) & 7) % 6];
    const int persp = (v >> 18) & 1;
    pfBegin(mode);
    for (int i = 0; i < n * 3; i++) {
        if (persp) {
            PFcolor c = { (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(64 + ((lcg() >> 24) % 192)) };
            pfColor(c);
            pfTexCoord2f(lcgf() * 4.0f - 2.0f, lcgf() * 4.0f - 2.0f);
            pfNormal3f(lcgf() - 0.5f, lcgf() - 0.5f, lcgf() + 0.2f);
            pfVertex3f(lcgf() * 5.0f - 2.5f, lcgf() * 4.0f - 2.0f, lcgf() * 6.0f - 2.0f);     /* some cross the near plane */
        } else random_vertex(w, h, 1, -0.9f, -0.1f);
    }
    pfEnd();
    pfDisable(PF_LIGHTING); pfDisableLight(PF_LIGHT0); pfDisableLight(PF_LIGHT1); pfLightModel(PF_GOURAUD);
    pfBindTexture(0);
}


/* ---- "api": breadth of the public API around the triangle path ------------------------------------ */
/* variant bits: 0 viewport sub-rectangle, 1 texture matrix, 2 Gouraud lighting with two lights (one spot
   with attenuation), 3 separate back material + no culling, 4 pfRect*, 5 pfDrawPixels with zoom,
   6 fog, 7 pfPostProcess, 8 pfReadPixels -> pfDrawPixels, 9 pfClearDepth(0.9), 10 cull front faces,
   11 PF_NORMALIZE with unnormalised normals, 12 colour material (front, diffuse), 13 vertex arrays with a
   colour pointer (pfDrawArrays, quads), 14 aux buffer + pfSwapBuffers, 15 fog mode request (overwritten by the density
   call, as upstream), 16 opaque fog colour, 17 blend on, 18 bilinear, 19-20 fog mode set AFTER the density call
   (1 PF_EXP, 2 PF_EXP2, 3 an invalid mode), 21 pfReadPixels / pfDrawPixels round trip through the BGRA8, RGB8 and BGR8
   layouts with a region that leaves the surface, 22 the same round trip through every (format, type) pair the
   reference has a getter and a setter for (38 pairs: single channels, luminance, 5-6-5 / 5-5-5-1 / 4-4-4-4, half, float) */

static PFcolor api_postprocess(PFint x, PFint y, PFfloat depth, PFcolor c)
{
    PFcolor o = c;
    if (((x >> 3) ^ (y >> 3)) & 1) { o.r = (PFubyte)(255 - c.r); o.b = (PFubyte)((c.b + c.g) >> 1); }
    if (depth < 0.5f) o.g = (PFubyte)(c.g >> 1);
    return o;
}

static void api_scene(const pfscene_cfg *cfg, PFtexture tex, uint8_t *aux)
{
    const int v = cfg->variant, w = cfg->width, h = cfg->height;
    lcg_state = (uint32_t)cfg->seed * 747796405u + 2891336453u;
    if (v & (1 << 14)) pfSetAuxBuffer(aux);
    pfClearColor(40, 30, 20, 255);
    pfClearDepth((v & (1 << 9)) ? 0.9f : 3.4028234663852886e38f);
    pfViewport(0, 0, (PFsizei)w, (PFsizei)h);
    pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
    cam_perspective(55.0, (double)w / h, 0.1, 50.0);
    if (v & 1) pfViewport(w / 8, h / 8, (PFsizei)(w * 3 / 4), (PFsizei)(h * 3 / 4));
    float eye[3] = { 0.5f, 0.8f, 4.0f }, at[3] = { 0, 0, 0 };
    cam_lookat(eye, at);
    pfEnable(PF_DEPTH_TEST); pfDepthFunc(PF_LESS);
    pfEnable(PF_TEXTURE_2D); pfBindTexture(tex);
    if (v & (1 << 17)) { pfEnable(PF_BLEND); pfBlendFunc(PF_BLEND_ALPHA); } else pfDisable(PF_BLEND);
    if (v & 2) {
        pfMatrixMode(PF_TEXTURE); pfLoadIdentity();
        pfTranslatef(0.25f, 0.1f, 0.0f); pfScalef(2.0f, 0.5f, 1.0f); pfRotatef(30.0f, 0.0f, 0.0f, 1.0f);
        pfMatrixMode(PF_MODELVIEW);
    }
    if (v & 8) pfDisable(PF_CULL_FACE); else { pfEnable(PF_CULL_FACE); pfCullFace((v & (1 << 10)) ? PF_FRONT : PF_BACK); }
    if (v & 4) {
        float p0[3] = { 1.0f, 2.0f, 3.0f }, p1[3] = { -2.0f, 1.0f, 2.0f }, d1[3] = { 0.6f, -0.3f, -0.7f };
        float dif[3] = { 0.9f, 0.8f, 0.6f }, spec[3] = { 0.5f, 0.6f, 0.7f }, amb[3] = { 0.15f, 0.1f, 0.2f };
        float mdif[3] = { 0.7f, 0.9f, 0.5f }, mspec[3] = { 0.9f, 0.9f, 0.9f }, memi[3] = { 0.05f, 0.0f, 0.1f }, bdif[3] = { 0.2f, 0.3f, 0.9f };
        pfEnable(PF_LIGHTING); pfLightModel(PF_GOURAUD);
        pfLightfv(PF_LIGHT0, PF_POSITION, p0); pfLightfv(PF_LIGHT0, PF_DIFFUSE, dif); pfLightfv(PF_LIGHT0, PF_AMBIENT, amb);
        pfLightf(PF_LIGHT0, PF_LINEAR_ATTENUATION, 0.08f); pfLightf(PF_LIGHT0, PF_CONSTANT_ATTENUATION, 0.9f);
        pfEnableLight(PF_LIGHT0);
        pfLightfv(PF_LIGHT2, PF_POSITION, p1); pfLightfv(PF_LIGHT2, PF_SPOT_DIRECTION, d1); pfLightfv(PF_LIGHT2, PF_SPECULAR, spec);
        pfLightf(PF_LIGHT2, PF_SPOT_INNER_CUTOFF, 20.0f); pfLightf(PF_LIGHT2, PF_SPOT_OUTER_CUTOFF, 35.0f);
        pfLightf(PF_LIGHT2, PF_QUADRATIC_ATTENUATION, 0.03f);
        pfEnableLight(PF_LIGHT2);
        pfMaterialfv(PF_FRONT, PF_DIFFUSE, mdif); pfMaterialfv(PF_FRONT, PF_SPECULAR, mspec); pfMaterialfv(PF_FRONT, PF_EMISSION, memi);
        pfMaterialf(PF_FRONT, PF_SHININESS, 12.0f);
        if (v & 8) { pfMaterialfv(PF_BACK, PF_AMBIENT_AND_DIFFUSE, bdif); pfMaterialf(PF_BACK, PF_SHININESS, 40.0f); }
        if (v & (1 << 12)) { pfEnable(PF_COLOR_MATERIAL); pfColorMaterial(PF_FRONT, PF_DIFFUSE); }
    }
    if (v & (1 << 11)) pfEnable(PF_NORMALIZE);
    const int n = cfg->size > 0 ? cfg->size : 48;
    if (v & (1 << 13)) {
        static float pos[4 * 64 * 3], nrm[4 * 64 * 3], uv[4 * 64 * 2]; static PFubyte col[4 * 64 * 4];
        const int q = n > 64 ? 64 : n;
        for (int i = 0; i < q; i++) {
            float cx = lcgf() * 4.0f - 2.0f, cy = lcgf() * 3.0f - 1.5f, cz = lcgf() * 3.0f - 1.5f, sz = 0.2f + lcgf() * 0.8f;
            static const float ox[4] = { -1, -1, 1, 1 }, oy[4] = { 1, -1, -1, 1 };
            for (int k = 0; k < 4; k++) {
                int j = i * 4 + k;
                pos[3 * j] = cx + ox[k] * sz; pos[3 * j + 1] = cy + oy[k] * sz; pos[3 * j + 2] = cz + 0.3f * ox[k] * oy[k];
                nrm[3 * j] = 0.3f * ox[k]; nrm[3 * j + 1] = 0.2f * oy[k]; nrm[3 * j + 2] = (v & (1 << 11)) ? 2.5f : 0.93f;
                uv[2 * j] = 0.5f * (ox[k] + 1.0f) * 1.5f; uv[2 * j + 1] = 0.5f * (oy[k] + 1.0f) * 1.5f;
                for (int c = 0; c < 4; c++) col[4 * j + c] = (PFubyte)(c == 3 ? 128 + (lcg() >> 25) : lcg() >> 24);
            }
        }
        pfEnable(PF_VERTEX_ARRAY); pfEnable(PF_NORMAL_ARRAY); pfEnable(PF_TEXTURE_COORD_ARRAY); pfEnable(PF_COLOR_ARRAY);
        pfVertexPointer(3, PF_FLOAT, 0, pos); pfNormalPointer(PF_FLOAT, 0, nrm); pfTexCoordPointer(PF_FLOAT, 0, uv);
        pfColorPointer(4, PF_UNSIGNED_BYTE, 0, col);
        pfDrawArrays(PF_QUADS, 0, (PFsizei)(q * 4));
        pfDisable(PF_VERTEX_ARRAY); pfDisable(PF_NORMAL_ARRAY); pfDisable(PF_TEXTURE_COORD_ARRAY); pfDisable(PF_COLOR_ARRAY);
    } else {
        pfBegin(PF_TRIANGLES);
        for (int i = 0; i < n * 3; i++) {
            PFcolor c = { (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(128 + (lcg() >> 25)) };
            pfColor(c);
            pfTexCoord2f(lcgf() * 3.0f - 1.0f, lcgf() * 3.0f - 1.0f);
            float nz = (v & (1 << 11)) ? 2.5f : 0.93f;
            pfNormal3f(0.6f * (lcgf() - 0.5f), 0.6f * (lcgf() - 0.5f), nz);
            pfVertex3f(lcgf() * 5.0f - 2.5f, lcgf() * 4.0f - 2.0f, lcgf() * 5.0f - 2.0f);
        }
        pfEnd();
    }
    pfDisable(PF_LIGHTING); pfDisableLight(PF_LIGHT0); pfDisableLight(PF_LIGHT2); pfDisable(PF_COLOR_MATERIAL); pfDisable(PF_NORMALIZE);
    pfMatrixMode(PF_TEXTURE); pfLoadIdentity(); pfMatrixMode(PF_MODELVIEW);
    pfBindTexture(0); pfDisable(PF_TEXTURE_2D);

    if (v & (1 << 4)) {
        ortho2d(w, h);
        pfColor4ub(250, 200, 10, 255); pfRectf(10.5f, 12.25f, 0.4f * w, 0.3f * h);
        pfColor4ub(10, 200, 250, 255); pfRects((PFshort)(w - 30), (PFshort)(h - 20), (PFshort)(w + 50), (PFshort)(h / 2));
        float a[2] = { 0.5f * w, 0.6f * h }, b[2] = { 0.45f * w, 0.9f * h };
        pfColor4ub(200, 20, 220, 255); pfRectfv(a, b);
    }
    static uint32_t sprite[24 * 16];
    if (v & ((1 << 5) | (1 << 8))) {
        ortho2d(w, h);
        for (int i = 0; i < 24 * 16; i++) sprite[i] = lcg() | ((i & 4) ? 0xFF000000u : 0x60000000u);
    }
    if (v & (1 << 5)) {
        pfEnable(PF_BLEND); pfBlendFunc(PF_BLEND_ALPHA);
        pfPixelZoom(2.5f, 1.75f); pfRasterPos3f(0.3f * w, 0.35f * h, -0.25f);
        pfDrawPixels(24, 16, PF_RGBA, PF_UNSIGNED_BYTE, sprite);
        pfDisable(PF_DEPTH_TEST); pfDisable(PF_BLEND);
        pfPixelZoom(1.0f, 1.0f); pfRasterPos2i(w - 12, h - 9);                   /* clipped by the viewport */
        pfDrawPixels(24, 16, PF_RGBA, PF_UNSIGNED_BYTE, sprite);
        pfEnable(PF_DEPTH_TEST);
    }
    if (v & (1 << 8)) {
        static uint32_t grab[40 * 30];
        memset(grab, 0, sizeof grab);
        pfReadPixels(w / 3, h / 3, 40, 30, PF_RGBA, PF_UNSIGNED_BYTE, grab);
        pfDisable(PF_DEPTH_TEST); pfDisable(PF_BLEND);
        pfPixelZoom(1.0f, 1.0f); pfRasterPos2f(4.0f, (float)h - 40.0f);
        pfDrawPixels(40, 30, PF_RGBA, PF_UNSIGNED_BYTE, grab);
        pfEnable(PF_DEPTH_TEST);
    }
    if (v & (1 << 21)) {
        /* read-back conversions and draw-pixels sources in the other 8-bit layouts; the first region starts left of and
           above the surface (clamped by pfReadPixels), the last one hangs over the right / bottom edge */
        static uint8_t conv[3][48 * 36 * 4];
        static const PFpixelformat fm[3] = { PF_BGRA, PF_RGB, PF_BGR };
        ortho2d(w, h);
        pfDisable(PF_BLEND);
        for (int k = 0; k < 3; k++) {
            memset(conv[k], 0x5a, sizeof conv[k]);
            pfReadPixels(k == 0 ? -7 : w / 4 + 31 * k, k == 0 ? -5 : (k == 2 ? h - 20 : h / 5), 48, 36, fm[k], PF_UNSIGNED_BYTE, conv[k]);
            if (k == 1) { pfEnable(PF_DEPTH_TEST); pfDepthFunc(PF_NOTEQUAL); } else pfDisable(PF_DEPTH_TEST);
            if (k == 2) { pfEnable(PF_BLEND); pfBlendFunc(PF_BLEND_SUB); }
            pfPixelZoom(k == 1 ? 1.5f : 1.0f, k == 2 ? 0.75f : 1.0f); pfRasterPos3f(6.0f + 52.0f * k, (float)h - 60.0f, -0.5f);
            pfDrawPixels(48, 36, fm[k], PF_UNSIGNED_BYTE, conv[k]);
        }
        pfDisable(PF_BLEND); pfEnable(PF_DEPTH_TEST); pfDepthFunc(PF_LESS); pfPixelZoom(1.0f, 1.0f);
    }
    if (v & (1 << 22)) {
        static const PFpixelformat fmts[10] = { PF_RED, PF_GREEN, PF_BLUE, PF_ALPHA, PF_LUMINANCE, PF_LUMINANCE_ALPHA, PF_RGB, PF_RGBA, PF_BGR, PF_BGRA };
        static const PFdatatype types[6] = { PF_UNSIGNED_BYTE, PF_UNSIGNED_SHORT_5_6_5, PF_UNSIGNED_SHORT_5_5_5_1, PF_UNSIGNED_SHORT_4_4_4_4, PF_HALF_FLOAT, PF_FLOAT };
        static uint8_t conv[20 * 14 * 16];
        ortho2d(w, h);
        pfDisable(PF_DEPTH_TEST);
        int k = 0;
        for (int f = 0; f < 10; f++) for (int t = 0; t < 6; t++) {
            const int comps = f < 5 ? 1 : (f == 5 ? 2 : ((fmts[f] == PF_RGB || fmts[f] == PF_BGR) ? 3 : 4));
            if ((t == 1 && comps != 3) || ((t == 2 || t == 3) && comps != 4)) continue;      /* pairs without getter / setter upstream */
            memset(conv, 0, sizeof conv);
            pfReadPixels(w / 5 + 3 * k, h / 4 + 2 * (k % 7), 20, 14, fmts[f], types[t], conv);
            if (k & 1) { pfEnable(PF_BLEND); pfBlendFunc((PFblendmode)(k % 8)); } else pfDisable(PF_BLEND);
            pfPixelZoom(1.0f, 1.0f); pfRasterPos2f(2.0f + 22.0f * (float)(k % 9), 2.0f + 16.0f * (float)(k / 9));
            pfDrawPixels(20, 14, fmts[f], types[t], conv);
            k++;
        }
        pfDisable(PF_BLEND); pfEnable(PF_DEPTH_TEST);
    }
    if (v & (1 << 6)) {
        PFint fc[4] = { 180, 190, 220, (v & (1 << 16)) ? 255 : 200 };
        pfFogi(PF_FOG_MODE, (PFint)(PF_LINEAR + ((v >> 15) & 1) * 1));
        pfFogf(PF_FOG_DENSITY, 0.8f); pfFogf(PF_FOG_START, 0.93f); pfFogf(PF_FOG_END, 0.985f);
        pfFogiv(PF_FOG_COLOR, fc);
        if ((v >> 19) & 3) {
            /* the exponential modes: reachable only through a call that does not range-check (context.c:2212-2226) */
            PFint m = ((v >> 19) & 3) == 3 ? 7 : (PFint)(PF_LINEAR + ((v >> 19) & 3));
            pfFogiv(PF_FOG_MODE, &m);
            pfFogf(PF_FOG_START, 1.2f); pfFogf(PF_FOG_END, 3.0f);      /* the depth range of the scene's triangles */
        }
        pfFogProcess();
    }
    if (v & (1 << 7)) pfPostProcess(api_postprocess);
    if (v & (1 << 14)) {
        /* present: the finished frame moves to the aux buffer, drawing continues in the other one */
        pfSwapBuffers();
        pfClearColor(1, 2, 3, 255); pfClear(PF_COLOR_BUFFER_BIT);
        ortho2d(w, h);
        pfDisable(PF_DEPTH_TEST);
        pfColor4ub(255, 255, 255, 255);
        pfBegin(PF_TRIANGLES); pfVertex2f(5.0f, 5.0f); pfVertex2f(5.0f, 0.8f * h); pfVertex2f(0.7f * w, 0.5f * h); pfEnd();
        pfSwapBuffers();
        /* thumbnail of what was drawn into the other buffer, so that it is part of the compared output */
        pfDisable(PF_BLEND);
        pfPixelZoom(0.25f, 0.25f); pfRasterPos2f(0.7f * w, 4.0f);
        pfDrawPixels((PFsizei)w, (PFsizei)h, PF_RGBA, PF_UNSIGNED_BYTE, aux);
        pfPixelZoom(1.0f, 1.0f);
    }
    pfDisable(PF_BLEND);
}

/* ---- "examples": the call sequences of the reference's example programs --------------------------- */
/* Restated from what the programs under examples/ do (examples/common.h:42-296 helpers, raylib/raylib_2D.c, _3D.c,
   _Framebuffer.c, _Points.c, _ModelWires.c, _TextureMatrix.c, _Texture2D.c, _FirstPerson.c, raylib_common.h:170-280), without
   their window system: variant & 15 picks the program, first_frame is its clock.  9 = vertex arrays in every component
   type the API accepts (pfDrawElements / pfDrawArrays). */
static void ex_begin3d(int w, int h, double fovy)
{
    pfMatrixMode(PF_PROJECTION); pfPushMatrix(); pfLoadIdentity();
    const float aspect = (float)w / (float)h, top = 0.01f * tanf((float)(fovy * 0.5 * SCN_PI / 180.0)), right = top * aspect;
    pfFrustum(-right, right, -top, top, 0.01f, 1000.0f);
    pfMatrixMode(PF_MODELVIEW); pfLoadIdentity();
    pfEnable(PF_DEPTH_TEST);
}
static void ex_end3d(void)
{
    pfMatrixMode(PF_PROJECTION); pfPopMatrix();
    pfMatrixMode(PF_MODELVIEW); pfLoadIdentity();
    pfDisable(PF_DEPTH_TEST);
}
static void ex_camera(float px, float py, float pz, float tx, float ty, float tz)
{
    const float eye[3] = { px, py, pz }, at[3] = { tx, ty, tz };
    cam_lookat(eye, at);
}
static void ex_cube(float size)
{
    const float h = size * 0.5f;
    static const float face[6][4][3] = {
        { { -1, -1, 1 }, { 1, -1, 1 }, { 1, 1, 1 }, { -1, 1, 1 } },     { { 1, -1, -1 }, { -1, -1, -1 }, { -1, 1, -1 }, { 1, 1, -1 } },
        { { -1, -1, -1 }, { -1, -1, 1 }, { -1, 1, 1 }, { -1, 1, -1 } }, { { 1, -1, 1 }, { 1, -1, -1 }, { 1, 1, -1 }, { 1, 1, 1 } },
        { { -1, 1, 1 }, { 1, 1, 1 }, { 1, 1, -1 }, { -1, 1, -1 } },     { { 1, -1, 1 }, { -1, -1, 1 }, { -1, -1, -1 }, { 1, -1, -1 } } };
    pfBegin(PF_QUADS);
    for (int f = 0; f < 6; f++) {
        pfColor3f(f < 2 ? 1.0f : 0.0f, (f >> 1) == 1 ? 1.0f : 0.0f, f >= 4 ? 1.0f : 0.0f);
        for (int k = 0; k < 4; k++) pfVertex3f(face[f][k][0] * h, face[f][k][1] * h, face[f][k][2] * h);
    }
    pfEnd();
}
static void ex_grid(int slices, float spacing)
{
    const int hs = slices / 2;
    pfBegin(PF_LINES);
    for (int i = -hs; i <= hs; i++) {
        if (i == 0) pfColor3f(0.5f, 0.5f, 0.5f); else pfColor3f(0.75f, 0.75f, 0.75f);
        pfVertex3f((float)i * spacing, 0.0f, (float)-hs * spacing); pfVertex3f((float)i * spacing, 0.0f, (float)hs * spacing);
        pfVertex3f((float)-hs * spacing, 0.0f, (float)i * spacing); pfVertex3f((float)hs * spacing, 0.0f, (float)i * spacing);
    }
    pfEnd();
}
static void ex_rotated_sprite(PFtexture tex, float x, float y, float wd, float ht, float ox, float oy, float deg)
{
    const float a = deg * (float)(SCN_PI / 180.0), c = cosf(a), sn = sinf(a), hw = wd * 0.5f, hh = ht * 0.5f;
    const float cx[4] = { -hw, -hw, hw, hw }, cy[4] = { -hh, hh, hh, -hh }, tu[4] = { 0, 0, 1, 1 }, tv[4] = { 0, 1, 1, 0 };
    pfBindTexture(tex);
    pfBegin(PF_QUADS);
    for (int k = 0; k < 4; k++) {
        pfTexCoord2f(tu[k], tv[k]);
        pfVertex2f(x - ox + (cx[k] * c - cy[k] * sn) + hw, y - oy + (cx[k] * sn + cy[k] * c) + hh);
    }
    pfEnd();
    pfBindTexture(0);
}

static void examples_scene(const pfscene_cfg *cfg, PFtexture tex, PFframebuffer *fbo, uint8_t *aux, const mesh_t *mesh, int frame)
{
    const int which = cfg->variant & 15, w = cfg->width, h = cfg->height;
    const float timer = 0.35f * (float)frame + 0.2f;
    lcg_state = (uint32_t)cfg->seed * 69069u + 7u;
    switch (which) {
    case 0:     /* 2D: a triangle in normalised device coordinates, nothing but the context's initial state */
        pfBegin(PF_TRIANGLES);
        pfColor3f(1.0f, 0.0f, 0.0f); pfVertex2f(-0.5f, -0.5f);
        pfColor3f(0.0f, 1.0f, 0.0f); pfVertex2f(0.5f, -0.5f);
        pfColor3f(0.0f, 0.0f, 1.0f); pfVertex2f(0.0f, 0.5f);
        pfEnd();
        break;
    case 1:     /* 3D: orbiting camera around the coloured cube */
        ortho2d(w, h);
        pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
        ex_begin3d(w, h, 60.0);
        ex_camera(2.0f * cosf(timer), 1.5f, 2.0f * sinf(timer), 0, 0, 0);
        ex_cube(1.0f);
        ex_end3d();
        break;
    case 2:     /* Framebuffer: the cube into an object of the target's size, then pfDrawPixels of it at half size */
        ortho2d(w, h);
        pfBindFramebuffer(fbo);
        pfEnable(PF_TEXTURE_2D);
        pfEnable(PF_FRAMEBUFFER);
        pfClearColor(255, 255, 255, 255);
        pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
        ex_begin3d(w, h, 60.0);
        ex_camera(2.0f * cosf(timer), 1.5f, 2.0f * sinf(timer), 0, 0, 0);
        ex_cube(1.0f);
        ex_end3d();
        pfDisable(PF_FRAMEBUFFER);
        pfClearColor(0, 0, 0, 255);
        pfClear(PF_COLOR_BUFFER_BIT);
        pfPixelZoom(0.5f, 0.5f);
        pfRasterPos2f((float)(w - w / 2) / 2.0f, (float)(h - h / 2) / 2.0f);
        {
            PFpixelformat format; PFdatatype type;
            const void *pixels = pfGetTexturePixels(fbo->texture, NULL, NULL, &format, &type);
            pfDrawPixels((PFsizei)w, (PFsizei)h, format, type, pixels);
        }
        pfPixelZoom(1.0f, 1.0f);
        pfDisable(PF_TEXTURE_2D);
        break;
    case 3:     /* Points: a lattice of sized points */
        ortho2d(w, h);
        pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
        pfPointSize(sinf(2.0f * timer) * 2.0f + 3.0f);
        ex_begin3d(w, h, 60.0);
        ex_camera(5.0f * cosf(timer), 3.0f, 5.0f * sinf(timer), 0, 0, 0);
        pfBegin(PF_POINTS);
        for (float x = -2.0f; x <= 2.0f; x += 0.5f)
            for (float y = -2.0f; y <= 2.0f; y += 0.5f)
                for (float z = -2.0f; z <= 2.0f; z += 0.5f) {
                    pfColor3f((x + 2.0f) / 4.0f, (y + 2.0f) / 4.0f, (y + 2.0f) / 4.0f);
                    pfVertex3f(x, y, z);
                }
        pfEnd();
        ex_end3d();
        pfPointSize(1.0f);
        break;
    case 4:     /* ModelWires: grid + a model drawn with front faces as lines (16-bit indices, as raylib meshes have) */
    case 8: {   /* FirstPerson: the model textured and lit by a spotlight carried by the camera */
        ortho2d(w, h);
        pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
        ex_begin3d(w, h, 60.0);
        static uint16_t idx16[65536];
        const int nidx = mesh->nidx < 65536 ? mesh->nidx : 65535 / 3 * 3;
        for (int k = 0; k < nidx; k++) idx16[k] = (uint16_t)mesh->idx[k];
        if (which == 4) {
            ex_camera(25.0f, 25.0f, 25.0f, 0, 10.0f, 0);
            ex_grid(10, 10.0f);
            pfPolygonMode(PF_FRONT, PF_LINE);
        } else {
            const float cam[3] = { 30.0f * cosf(timer), 14.0f, 30.0f * sinf(timer) }, dir[3] = { -cam[0], 10.0f - cam[1], -cam[2] };
            const float amb[3] = { 0.2f, 0.2f, 0.3f };
            ex_camera(cam[0], cam[1], cam[2], 0, 10.0f, 0);
            pfLightf(PF_LIGHT0, PF_SPOT_INNER_CUTOFF, 17.5f); pfLightf(PF_LIGHT0, PF_SPOT_OUTER_CUTOFF, 32.5f);
            pfLightfv(PF_LIGHT0, PF_AMBIENT, amb);
            pfLightf(0, PF_LINEAR_ATTENUATION, 0.009f); pfLightf(0, PF_QUADRATIC_ATTENUATION, 0.0032f);
            pfLightfv(PF_LIGHT0, PF_POSITION, cam); pfLightfv(PF_LIGHT0, PF_SPOT_DIRECTION, dir);
            pfEnable(PF_LIGHTING); pfEnableLight(PF_LIGHT0);
            pfEnable(PF_TEXTURE_2D); pfBindTexture(tex);
        }
        pfPushMatrix();
        pfTranslatef(0.0f, 10.0f, 0.0f); pfScalef(0.9f, 0.9f, 0.9f);
        pfColor4ub(255, 255, 255, 255);
        pfEnable(PF_VERTEX_ARRAY); pfVertexPointer(3, PF_FLOAT, 0, mesh->pos);
        pfEnable(PF_NORMAL_ARRAY); pfNormalPointer(PF_FLOAT, 0, mesh->nrm);
        pfEnable(PF_TEXTURE_COORD_ARRAY); pfTexCoordPointer(PF_FLOAT, 0, mesh->uv);
        pfDrawElements(PF_TRIANGLES, (PFsizei)nidx, PF_UNSIGNED_SHORT, idx16);
        pfDisable(PF_VERTEX_ARRAY); pfDisable(PF_NORMAL_ARRAY); pfDisable(PF_TEXTURE_COORD_ARRAY);
        pfPopMatrix();
        if (which == 4) pfPolygonMode(PF_FRONT, PF_FILL);
        else { pfBindTexture(0); pfDisable(PF_TEXTURE_2D); pfDisable(PF_LIGHTING); pfDisableLight(PF_LIGHT0); }
        ex_end3d();
        break; }
    case 5:     /* TextureMatrix: a ground plane far larger than the frustum, its texture scrolled by the texture matrix */
        ortho2d(w, h);
        pfEnable(PF_TEXTURE_2D);
        pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
        ex_begin3d(w, h, 60.0);
        ex_camera(-2.0f, 1.5f + 0.1f * (float)frame, -2.0f, 0, 0, 0);
        pfMatrixMode(PF_TEXTURE); pfLoadIdentity();
        pfTranslatef(0.25f * timer, 0.25f * timer, 0.0f);
        pfMatrixMode(PF_MODELVIEW);
        pfBindTexture(tex);
        pfColor4ub(255, 255, 255, 255);
        pfBegin(PF_QUADS);
        pfTexCoord2f(0, 0); pfVertex3f(-1000, 0, -1000);
        pfTexCoord2f(0, 200); pfVertex3f(-1000, 0, 1000);
        pfTexCoord2f(200, 200); pfVertex3f(1000, 0, 1000);
        pfTexCoord2f(200, 0); pfVertex3f(1000, 0, -1000);
        pfEnd();
        pfBindTexture(NULL);
        ex_end3d();
        pfMatrixMode(PF_TEXTURE); pfLoadIdentity(); pfMatrixMode(PF_MODELVIEW);
        pfDisable(PF_TEXTURE_2D);
        break;
    case 6:     /* Texture2D: two states in one pfEnable, a LUMINANCE_ALPHA background stretched by pfDrawPixels, rotated sprites */
        ortho2d(w, h);
        pfEnable((PFstate)(PF_TEXTURE_2D | PF_BLEND));
        pfClear(PF_COLOR_BUFFER_BIT);
        for (size_t i = 0; i < 40u * 20u * 2u; i++) aux[i] = (uint8_t)(lcg() >> 24);
        pfRasterPos2i(0, h - h / 2);
        pfPixelZoom((float)w / 40.0f, (float)(h / 2) / 20.0f);
        pfDrawPixels(40, 20, PF_LUMINANCE_ALPHA, PF_UNSIGNED_BYTE, aux);
        pfPixelZoom(1.0f, 1.0f);
        for (int k = 0; k < 12; k++)
            ex_rotated_sprite(tex, 20.0f + lcgf() * (float)(w - 100), 20.0f + lcgf() * (float)(h - 100), 64, 64, 32, 32, 360.0f * lcgf() + 20.0f * timer);
        pfDisable((PFstate)(PF_TEXTURE_2D | PF_BLEND));
        break;
    case 12: {  /* matrix stacks: overflow and underflow of all three, nested transforms, rotation about a non-unit axis */
        ortho2d(w, h);
        pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
        PFerrcode errs[8]; int ne = 0;
        static const PFmatrixmode stacks[3] = { PF_PROJECTION, PF_MODELVIEW, PF_TEXTURE };
        for (int m = 0; m < 3; m++) {
            pfMatrixMode(stacks[m]);
            int pushed = 0;
            for (int k = 0; k < 70; k++) { pfPushMatrix(); if (pfGetError() != PF_NO_ERROR) break; pushed++; if (stacks[m] == PF_MODELVIEW) pfTranslatef(0.5f, 0.25f, 0.0f); }
            errs[ne++] = (PFerrcode)pushed;
            int popped = 0;
            for (int k = 0; k < 80; k++) { pfPopMatrix(); if (pfGetError() != PF_NO_ERROR) break; popped++; }
            errs[ne++] = (PFerrcode)popped;
        }
        pfMatrixMode(PF_MODELVIEW); pfLoadIdentity();
        pfEnable(PF_TEXTURE_2D); pfBindTexture(tex);
        for (int k = 0; k < 6; k++) {
            pfPushMatrix();
            pfTranslatef(50.0f + 44.0f * (float)k, 60.0f + 20.0f * (float)(k & 1), 0.0f);
            pfRotatef(17.0f * (float)k + 5.0f * timer, 0.0f, 0.0f, 2.5f);              /* axis length 2.5 */
            pfScalef(1.0f + 0.1f * (float)k, 0.8f, 1.0f);
            pfPushMatrix();
            const float shear[16] = { 1, 0.2f, 0, 0, 0.3f, 1, 0, 0, 0, 0, 1, 0, 2.0f, -3.0f, 0, 1 };
            pfMultMatrixf(shear);
            pfMatrixMode(PF_TEXTURE); pfPushMatrix(); pfRotatef(10.0f * (float)k, 0, 0, 1); pfScalef(2.0f, 2.0f, 1.0f); pfMatrixMode(PF_MODELVIEW);
            pfColor4ub((PFubyte)(255 - 30 * k), (PFubyte)(100 + 25 * k), 200, 255);
            pfBegin(PF_QUADS);
            pfTexCoord2f(0, 0); pfVertex2f(-18, -14); pfTexCoord2f(0, 1); pfVertex2f(-18, 14);
            pfTexCoord2f(1, 1); pfVertex2f(18, 14); pfTexCoord2f(1, 0); pfVertex2f(18, -14);
            pfEnd();
            pfMatrixMode(PF_TEXTURE); pfPopMatrix(); pfMatrixMode(PF_MODELVIEW);
            pfPopMatrix();
            pfColor4ub(40, (PFubyte)(40 * k), 90, 255); pfRectf(-4.0f, 18.0f, 4.0f, 26.0f);      /* under the outer transform only */
            pfPopMatrix();
        }
        pfBindTexture(0); pfDisable(PF_TEXTURE_2D);
        /* a projection pushed inside a frame, 3D under it, then back to the 2D one */
        ex_begin3d(w, h, 45.0);
        ex_camera(3.0f, 2.0f, 4.0f, 0, 0, 0);
        pfPushMatrix(); pfRotatef(30.0f + 10.0f * timer, 1.0f, 1.0f, 0.0f); ex_cube(1.5f); pfPopMatrix();
        ex_end3d();
        for (int k = 0; k < ne; k++) { pfColor4ub((PFubyte)errs[k], (PFubyte)(errs[k] >> 8), (PFubyte)k, 255); pfRecti(4 * k, 0, 4 * k + 4, 4); }
        pfColor4ub(255, 255, 255, 255);
        break; }
    case 11: {  /* all eight lights at once (Gouraud; variant bit 4: per-fragment Phong), spots and attenuation among them */
        ortho2d(w, h);
        pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
        ex_begin3d(w, h, 60.0);
        ex_camera(28.0f * cosf(timer), 16.0f, 28.0f * sinf(timer), 0, 0, 0);
        pfEnable(PF_LIGHTING); pfLightModel((cfg->variant & 16) ? PF_PHONG : PF_GOURAUD);
        pfEnable(PF_CULL_FACE); pfCullFace(PF_BACK);
        for (int li = 0; li < 8; li++) {
            const float a = (float)li * 0.785398f;
            const float pos[3] = { 22.0f * cosf(a), 6.0f + 3.0f * (float)(li & 3), 22.0f * sinf(a) }, dir[3] = { -pos[0], -pos[1], -pos[2] };
            const float dif[3] = { 0.25f + 0.09f * (float)li, 0.9f - 0.1f * (float)li, 0.3f + 0.05f * (float)((li * 3) & 7) };
            const float spc[3] = { 0.6f, 0.5f + 0.05f * (float)li, 0.4f }, amb[3] = { 0.02f * (float)li, 0.03f, 0.04f };
            pfLightfv((PFsizei)li, PF_POSITION, pos); pfLightfv((PFsizei)li, PF_DIFFUSE, dif);
            pfLightfv((PFsizei)li, PF_SPECULAR, spc); pfLightfv((PFsizei)li, PF_AMBIENT, amb);
            if (li & 1) { pfLightfv((PFsizei)li, PF_SPOT_DIRECTION, dir); pfLightf((PFsizei)li, PF_SPOT_INNER_CUTOFF, 20.0f + 2.0f * (float)li); pfLightf((PFsizei)li, PF_SPOT_OUTER_CUTOFF, 35.0f + 2.0f * (float)li); }
            if (li & 2) { pfLightf((PFsizei)li, PF_LINEAR_ATTENUATION, 0.01f * (float)li); pfLightf((PFsizei)li, PF_QUADRATIC_ATTENUATION, 0.001f * (float)li); }
            pfEnableLight((PFsizei)li);
        }
        {
            const float mdif[3] = { 0.8f, 0.7f, 0.6f }, mspec[3] = { 0.9f, 0.9f, 0.8f }, mamb[3] = { 0.3f, 0.3f, 0.4f };
            pfMaterialfv(PF_FRONT_AND_BACK, PF_DIFFUSE, mdif); pfMaterialfv(PF_FRONT_AND_BACK, PF_SPECULAR, mspec);
            pfMaterialfv(PF_FRONT_AND_BACK, PF_AMBIENT, mamb); pfMaterialf(PF_FRONT_AND_BACK, PF_SHININESS, 24.0f);
        }
        pfColor4ub(255, 255, 255, 255);
        draw_mesh_arrays(mesh);
        for (int li = 0; li < 8; li++) pfDisableLight((PFsizei)li);
        pfDisable(PF_LIGHTING); pfLightModel(PF_GOURAUD); pfDisable(PF_CULL_FACE);
        ex_end3d();
        break; }
    case 10: {  /* pfBegin / pfEnd used loosely: what the reference does with it (context.c:1580-1608, 1658-1685) */
        ortho2d(w, h);
        pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
        pfColor4ub(250, 60, 60, 255);
        pfBegin(PF_TRIANGLES); pfVertex2f(10, 10); pfVertex2f(10, 90); pfEnd();                 /* incomplete: dropped */
        pfBegin(PF_QUADS); pfVertex2f(20, 20); pfVertex2f(20, 80);
        pfBegin(PF_TRIANGLES);                                                                  /* a second pfBegin starts over */
        pfColor4ub(60, 250, 60, 255); pfVertex2f(30, 20); pfVertex2f(30, 100); pfVertex2f(120, 60);
        pfEnd();
        pfColor4ub(60, 60, 250, 255);                                                           /* vertices after pfEnd: still assembled */
        pfVertex2f(130, 20); pfVertex2f(130, 100); pfVertex2f(220, 60);
        pfTranslatef(40.0f, 100.0f, 0.0f);                                                      /* ... with the matrices of the last pfBegin */
        pfColor4ub(250, 250, 60, 255);
        pfVertex2f(130, 20); pfVertex2f(130, 100); pfVertex2f(220, 60);
        pfBegin((PFdrawmode)77);                                                                /* refused: mode and counter stay */
        const PFerrcode e1 = pfGetError();
        pfColor4ub(60, 250, 250, 255);
        pfVertex2f(10, 120); pfVertex2f(10, 200); pfVertex2f(100, 160);
        pfEnd(); pfEnd();
        pfBegin(PF_TRIANGLE_STRIP);
        pfColor4ub(250, 60, 250, 255);
        pfVertex2f(230, 120); pfVertex2f(230, 200); pfVertex2f(260, 120); pfVertex2f(260, 200); pfVertex2f(290, 120);
        pfEnd();
        pfVertex2f(290, 200); pfVertex2f(310, 120);                                             /* the strip after its pfEnd */
        pfVertex2f(310, 200); pfVertex2f(315, 120);
        pfLoadIdentity();
        pfColor4ub((PFubyte)e1, 7, 9, 255); pfRecti(0, 0, 8, 4);
        pfColor4ub(255, 255, 255, 255);
        break; }
    default: {  /* 9: vertex arrays in every component type; strides are ignored upstream (context.c:1282-1370 index j*size+k) */
        ortho2d(w, h);
        pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
        static PFshort ps[64 * 2]; static PFint pi[64 * 3]; static PFfloat pf4[64 * 4]; static PFdouble pd[64 * 3];
        static PFfloat nf[64 * 3]; static PFdouble nd[64 * 3]; static PFfloat tf[64 * 2]; static PFdouble td[64 * 2];
        static PFubyte cub[64 * 4]; static PFushort cus[64 * 3]; static PFuint cui[64 * 4]; static PFfloat cf[64 * 3]; static PFdouble cd[64 * 4];
        static PFubyte i8[96]; static PFushort i16[96]; static PFuint i32[96];
        pfEnable(PF_TEXTURE_2D); pfBindTexture(tex);
        pfEnable(PF_BLEND); pfBlendFunc(PF_BLEND_ALPHA);
        for (int pass = 0; pass < 10; pass++) {
            const int ox = 10 + (pass % 5) * 60, oy = 10 + (pass / 5) * 90;
            for (int v = 0; v < 64; v++) {
                const float x = (float)ox + lcgf() * 50.0f, y = (float)oy + lcgf() * 80.0f;
                ps[2 * v] = (PFshort)x; ps[2 * v + 1] = (PFshort)y;
                pi[3 * v] = (PFint)x; pi[3 * v + 1] = (PFint)y; pi[3 * v + 2] = 0;
                pf4[4 * v] = x; pf4[4 * v + 1] = y; pf4[4 * v + 2] = -0.25f; pf4[4 * v + 3] = 1.0f;
                pd[3 * v] = x; pd[3 * v + 1] = y; pd[3 * v + 2] = -0.5;
                for (int k = 0; k < 3; k++) { nf[3 * v + k] = lcgf() - 0.5f; nd[3 * v + k] = nf[3 * v + k]; }
                for (int k = 0; k < 2; k++) { tf[2 * v + k] = lcgf() * 2.0f; td[2 * v + k] = tf[2 * v + k]; }
                for (int k = 0; k < 4; k++) {
                    const uint32_t r = lcg();
                    cub[4 * v + k] = (PFubyte)(r >> 24); cui[4 * v + k] = r; cd[4 * v + k] = (double)(r >> 8) / 16777216.0;
                    if (k < 3) { cus[3 * v + k] = (PFushort)(r >> 16); cf[3 * v + k] = (float)(r >> 8) / 16777216.0f; }
                }
            }
            for (int k = 0; k < 96; k++) { const uint32_t j = lcg() >> 26; i8[k] = (PFubyte)j; i16[k] = (PFushort)j; i32[k] = j; }
            pfEnable(PF_VERTEX_ARRAY);
            switch (pass % 4) {
            case 0: pfVertexPointer(2, PF_SHORT, 0, ps); break;     case 1: pfVertexPointer(3, PF_INT, 64, pi); break;
            case 2: pfVertexPointer(4, PF_FLOAT, 0, pf4); break;    default: pfVertexPointer(3, PF_DOUBLE, 8, pd); break;
            }
            if (pass & 1) { pfEnable(PF_NORMAL_ARRAY); if (pass & 2) pfNormalPointer(PF_DOUBLE, 0, nd); else pfNormalPointer(PF_FLOAT, 4, nf); }
            if (pass % 3) { pfEnable(PF_TEXTURE_COORD_ARRAY); if (pass & 4) pfTexCoordPointer(PF_DOUBLE, 0, td); else pfTexCoordPointer(PF_FLOAT, 0, tf); }
            pfEnable(PF_COLOR_ARRAY);
            switch (pass % 5) {
            case 0: pfColorPointer(4, PF_UNSIGNED_BYTE, 0, cub); break;    case 1: pfColorPointer(3, PF_UNSIGNED_SHORT, 0, cus); break;
            case 2: pfColorPointer(4, PF_UNSIGNED_INT, 16, cui); break;    case 3: pfColorPointer(3, PF_FLOAT, 0, cf); break;
            default: pfColorPointer(4, PF_DOUBLE, 0, cd); break;
            }
            static const PFdrawmode modes[5] = { PF_TRIANGLES, PF_QUADS, PF_TRIANGLE_STRIP, PF_TRIANGLE_FAN, PF_QUAD_STRIP };
            const PFdrawmode mode = modes[pass % 5];
            if (pass % 3 == 0) pfDrawElements(mode, 96, PF_UNSIGNED_BYTE, i8);
            else if (pass % 3 == 1) pfDrawElements(mode, 96, PF_UNSIGNED_SHORT, i16);
            else pfDrawElements(mode, 96, PF_UNSIGNED_INT, i32);
            pfDrawArrays(pass & 1 ? PF_TRIANGLES : PF_QUADS, 4 + pass, 24);
            pfDisable(PF_VERTEX_ARRAY); pfDisable(PF_NORMAL_ARRAY); pfDisable(PF_TEXTURE_COORD_ARRAY); pfDisable(PF_COLOR_ARRAY);
        }
        /* what the error paths leave behind goes into the corner of the frame */
        pfDrawElements(PF_TRIANGLES, 3, PF_UNSIGNED_INT, i32);       /* arrays disabled */
        PFerrcode e1 = pfGetError();
        pfEnable(PF_VERTEX_ARRAY);
        pfDrawElements(PF_TRIANGLES, 3, PF_FLOAT, i32);              /* not an index type */
        PFerrcode e2 = pfGetError();
        pfDisable(PF_VERTEX_ARRAY);
        pfDrawArrays(PF_TRIANGLES, 0, 3);
        PFerrcode e3 = pfGetError();
        pfDisable(PF_BLEND); pfBindTexture(0); pfDisable(PF_TEXTURE_2D);
        pfColor4ub((PFubyte)e1, (PFubyte)e2, (PFubyte)e3, 255); pfRecti(0, 0, 8, 4);
        pfColor4ub(255, 255, 255, 255);
        break; }
    }
}

/* ---- "fuzz": a random walk over the API --------------------------------------------------------- */
/* `size` operations drawn from the LCG (seed): state toggles, blend / depth / cull / shade / light model / polygon mode,
   matrix operations on all three stacks, 2D and perspective projections, lights and materials, texture parameters, immediate
   primitives of every draw mode, rectangles, render lists recorded and replayed on the spot, clears, viewports, pfDrawPixels,
   vertex arrays, a framebuffer object rendered into and sampled, fog.
   Everything the reference does not bounds-check stays inside: 2D coordinates keep 14 pixels from the border, lines and
   points are one pixel wide, PF_POINT / PF_LINE polygon modes and points / lines are 2D only. */
static void fuzz_scene(const pfscene_cfg *cfg, PFtexture tex, uint8_t *aux, PFframebuffer *fbo)
{
    const int nops = cfg->size > 0 ? cfg->size : 120;
    lcg_state = (uint32_t)cfg->seed * 2246822519u + 374761393u;
    /* two contexts: the scene's own and a 128x96 one created for the frame; the walk switches between them, each keeps its own
       state (and the walk its own notes about that state); lists, the texture and the framebuffer object serve both */
    struct { int w, h, persp, tex_enabled, tex_bound, mv_far, clean; } T[2];
    int cur = 0, list_persp[4] = { 0, 0, 0, 0 };
    PFcontext ctxs[2] = { pfGetCurrentContext(), NULL };
    uint8_t *second = (uint8_t *)calloc(128u * 96u * 4u + 64u, 1);
    ctxs[1] = pfCreateContext(second, 128, 96, PF_RGBA, PF_UNSIGNED_BYTE);
    T[0].w = cfg->width; T[0].h = cfg->height; T[1].w = 128; T[1].h = 96;
#define w (T[cur].w)
#define h (T[cur].h)
#define persp (T[cur].persp)
#define tex_enabled (T[cur].tex_enabled)
#define tex_bound (T[cur].tex_bound)
#define mv_far (T[cur].mv_far)          /* the modelview holds the camera, rotations and shrinking scales only */
#define clean (T[cur].clean)            /* 2D, identity modelview, full viewport: only then points, lines, PF_POINT / PF_LINE modes and pfDrawPixels (upstream checks no bounds there) */
    /* perspective + texture: geometry stays farther than 2 * near from the eye - where clip-space z crosses 0 the reference's
       perspective-corrected texcoords become inf / NaN and its gathers leave the texture by gigabytes (it crashes) */
#define FUZZ_FAR() (tex_enabled && tex_bound)
#define FUZZ_DIRTY() do { clean = 0; pfPolygonMode(PF_FRONT_AND_BACK, PF_FILL); } while (0)
    PFrenderlist lists[4] = { NULL, NULL, NULL, NULL };
    for (cur = 1; cur >= 0; cur--) {
        pfMakeCurrent(ctxs[cur]);
        persp = 0; tex_enabled = 0; tex_bound = 0; mv_far = 1; clean = 1;
        pfClearColor((PFubyte)(16 + 40 * cur), 24, 40, 255); pfClearDepth(3.4028234663852886e38f);
        pfViewport(0, 0, (PFsizei)w, (PFsizei)h);
        pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
        ortho2d(w, h);
        pfPolygonMode(PF_FRONT_AND_BACK, PF_FILL); pfLineWidth(1.0f); pfPointSize(1.0f);
    }
    cur = 0;
    for (int op = 0; op < nops; op++) {
        const uint32_t r = lcg() >> 8;
        if (getenv("PFSCENE_FUZZ_TRACE")) fprintf(stderr, "fuzz op %d kind %u r %06x persp %d\n", op, r % 19u, r, persp);
        switch (r % 19u) {
        case 0: case 1: {
            static const PFstate st[7] = { PF_BLEND, PF_DEPTH_TEST, PF_CULL_FACE, PF_TEXTURE_2D, PF_LIGHTING, PF_NORMALIZE, PF_COLOR_MATERIAL };
            const PFstate which = st[(r >> 5) % 7u];
            if ((r >> 9) & 1u) pfEnable(which); else pfDisable(which);
            if (which == PF_TEXTURE_2D) tex_enabled = (int)((r >> 9) & 1u);
            break; }
        case 2:
            pfBlendFunc((PFblendmode)((r >> 5) % 8u)); pfDepthFunc((PFdepthmode)((r >> 9) % 6u));
            pfCullFace((r >> 13) & 1u ? PF_FRONT : PF_BACK);
            break;
        case 3:
            pfShadeModel((r >> 5) & 1u ? PF_FLAT : PF_SMOOTH); pfLightModel((r >> 6) & 1u ? PF_PHONG : PF_GOURAUD);
            if (clean) { pfPolygonMode((r >> 7) & 1u ? PF_FRONT : PF_BACK, (PFpolygonmode)((r >> 8) % 3u)); }
            break;
        case 4: {
            pfMatrixMode(PF_MODELVIEW);
            const uint32_t k = (r >> 5) % 6u;
            if (k == 0) pfPushMatrix(); else if (k == 1) { pfPopMatrix(); FUZZ_DIRTY(); mv_far = 0; }
            else if (k == 2) { FUZZ_DIRTY(); if (persp) { const float tx = lcgf() - 0.5f, ty = lcgf() - 0.5f, tz = lcgf() - 0.5f; if (!FUZZ_FAR()) { pfTranslatef(tx, ty, tz); mv_far = 0; } } else pfTranslatef(6.0f * lcgf() - 3.0f, 6.0f * lcgf() - 3.0f, 0.0f); }
            else if (k == 3) { if (persp) pfRotatef(40.0f * lcgf(), lcgf(), lcgf(), lcgf() + 0.1f); }
            else if (k == 4) { if (persp) pfScalef(0.8f + 0.2f * lcgf(), 0.8f + 0.2f * lcgf(), 0.8f + 0.2f * lcgf()); }
            else { pfLoadIdentity(); if (persp) FUZZ_DIRTY(); }
            (void)pfGetError();
            break; }
        case 5:
            persp = (int)((r >> 5) & 1u);
            pfPolygonMode(PF_FRONT_AND_BACK, PF_FILL); pfViewport(0, 0, (PFsizei)w, (PFsizei)h);
            /* empty the modelview stack first: while a push is outstanding pfLoadIdentity resets the MODEL matrix only and the
               view matrix of an earlier camera would stay (context.c:434-441) */
            pfMatrixMode(PF_MODELVIEW); for (int k = 0; k < 40; k++) pfPopMatrix();
            (void)pfGetError();
            clean = !persp; mv_far = 1;
            if (persp) {
                cam_perspective(50.0 + 20.0 * lcgf(), (double)w / h, 0.1, 60.0);
                const float eye[3] = { 2.0f * lcgf() - 1.0f, 2.0f * lcgf() - 1.0f, 3.0f + 2.0f * lcgf() }, at[3] = { 0, 0, 0 };
                cam_lookat(eye, at);
            } else ortho2d(w, h);
            break;
        case 6: {
            const PFcolor c = { (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(96 + ((lcg() >> 24) % 160u)) };
            pfColor(c); pfNormal3f(lcgf() - 0.5f, lcgf() - 0.5f, lcgf() + 0.1f); pfTexCoord2f(3.0f * lcgf() - 1.0f, 3.0f * lcgf() - 1.0f);
            break; }
        case 7: case 8: case 9: case 10: {
            static const PFdrawmode modes[8] = { PF_TRIANGLES, PF_QUADS, PF_TRIANGLE_FAN, PF_TRIANGLE_STRIP, PF_QUAD_FAN, PF_QUAD_STRIP, PF_POINTS, PF_LINES };
            const PFdrawmode mode = modes[(r >> 5) % (clean ? 8u : 6u)];
            const int nv = 3 + (int)((r >> 9) % 10u);
            if (persp && FUZZ_FAR() && !mv_far) { pfBindTexture(NULL); tex_bound = 0; }
            pfBegin(mode);
            for (int v = 0; v < nv; v++) {
                if ((lcg() >> 30) == 0) pfColor4ub((PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(128 + (lcg() >> 25)));
                pfTexCoord2f(3.0f * lcgf() - 1.0f, 3.0f * lcgf() - 1.0f);
                pfNormal3f(lcgf() - 0.5f, lcgf() - 0.5f, lcgf() + 0.2f);
                if (persp) { const float x = lcgf() - 0.5f, y = lcgf() - 0.5f, z = lcgf() - 0.4f; if (FUZZ_FAR()) pfVertex3f(2.4f * x, 2.0f * y, 2.0f * z); else pfVertex3f(5.0f * x, 4.0f * y, 5.0f * z); }
                else pfVertex3f(20.0f + lcgf() * (float)(w - 40), 20.0f + lcgf() * (float)(h - 40), -0.9f * lcgf());
            }
            pfEnd();
            break; }
        case 11:
            if (!persp) { const float x = 20.0f + lcgf() * (float)(w - 90), y = 20.0f + lcgf() * (float)(h - 70); pfRectf(x, y, x + 10.0f + 40.0f * lcgf(), y + 10.0f + 30.0f * lcgf()); }
            break;
        case 12: {
            const PFsizei li = (PFsizei)((r >> 5) % 4u);
            const float pos[3] = { 6.0f * lcgf() - 3.0f, 6.0f * lcgf() - 3.0f, 1.0f + 4.0f * lcgf() }, dir[3] = { -pos[0], -pos[1], -pos[2] };
            const float col[3] = { lcgf(), lcgf(), lcgf() };
            switch ((r >> 8) % 6u) {
            case 0: pfEnableLight(li); break;                        case 1: pfDisableLight(li); break;
            case 2: pfLightfv(li, PF_POSITION, pos); break;           case 3: pfLightfv(li, (r >> 12) & 1u ? PF_DIFFUSE : PF_SPECULAR, col); break;
            case 4: pfLightfv(li, PF_SPOT_DIRECTION, dir); pfLightf(li, PF_SPOT_INNER_CUTOFF, 15.0f + 20.0f * lcgf()); pfLightf(li, PF_SPOT_OUTER_CUTOFF, 40.0f + 20.0f * lcgf()); break;
            default: pfLightf(li, PF_LINEAR_ATTENUATION, 0.1f * lcgf()); pfLightf(li, PF_QUADRATIC_ATTENUATION, 0.05f * lcgf()); break;
            }
            break; }
        case 13: {
            const float col[3] = { lcgf(), lcgf(), lcgf() };
            static const PFface faces[3] = { PF_FRONT, PF_BACK, PF_FRONT_AND_BACK };
            static const PFenum what[5] = { PF_AMBIENT, PF_DIFFUSE, PF_SPECULAR, PF_EMISSION, PF_AMBIENT_AND_DIFFUSE };
            const PFface f = faces[(r >> 5) % 3u];
            if ((r >> 8) & 1u) pfMaterialfv(f, what[(r >> 9) % 5u], col); else pfMaterialf(f, PF_SHININESS, 2.0f + 60.0f * lcgf());
            if ((r >> 12) % 5u == 0) pfColorMaterial(f, what[(r >> 15) % 5u]);
            break; }
        case 14:
            pfTextureParameter(tex, (PFtexturewrap)((r >> 5) % 3u), (r >> 8) & 1u ? PF_BILINEAR : PF_NEAREST);
            pfBindTexture((r >> 9) % 4u ? tex : NULL); tex_bound = (r >> 9) % 4u != 0;
            if ((r >> 12) % 4u == 0) { pfMatrixMode(PF_TEXTURE); pfLoadIdentity(); if ((r >> 14) & 1u) { pfScalef(1.0f + lcgf(), 1.0f + lcgf(), 1.0f); pfTranslatef(lcgf(), lcgf(), 0.0f); } pfMatrixMode(PF_MODELVIEW); }
            break;
        case 15: {
            const int li = (int)((r >> 5) % 4u);
            if (!lists[li]) lists[li] = pfGenList();
            if (persp && !mv_far) break;             /* replays bind the textures their calls were recorded with */
            if ((r >> 8) & 1u) {
                list_persp[li] = persp;
                static const PFdrawmode lmodes[3] = { PF_TRIANGLES, PF_QUADS, PF_TRIANGLE_STRIP };
                pfNewList(lists[li]);
                pfBegin(lmodes[(r >> 9) % 3u]);
                const int nv = 3 + (int)((r >> 12) % 6u);
                for (int v = 0; v < nv; v++) {
                    pfTexCoord2f(2.0f * lcgf(), 2.0f * lcgf()); pfNormal3f(lcgf() - 0.5f, lcgf() - 0.5f, 1.0f);
                    if (persp) pfVertex3f(2.4f * lcgf() - 1.2f, 2.0f * lcgf() - 1.0f, 2.0f * lcgf() - 1.0f);
                    else pfVertex2f(24.0f + lcgf() * (float)(w - 48), 24.0f + lcgf() * (float)(h - 48));
                }
                pfEnd();
                pfEndList();
            }
            if (list_persp[li] == persp) pfCallList(lists[li]);
            break; }
        case 16:
            if ((r >> 5) % 6u == 0) pfClear((PFclearflag)(((r >> 9) & 1u ? PF_COLOR_BUFFER_BIT : 0) | ((r >> 10) & 1u ? PF_DEPTH_BUFFER_BIT : 0)));
            else if ((r >> 5) % 6u == 3) { cur ^= 1; pfMakeCurrent(ctxs[cur]); }
            else if ((r >> 5) % 6u == 1) {
                const int vx = (int)((r >> 9) % 24u), vy = (int)((r >> 14) % 24u);
                pfViewport(vx, vy, (PFsizei)(w - 2 * vx - 8), (PFsizei)(h - 2 * vy - 8)); FUZZ_DIRTY();
            }
            break;
        case 17:
            if (clean) {
                for (size_t i = 0; i < 12u * 9u * 4u; i++) aux[i] = (uint8_t)(lcg() >> 24);
                pfRasterPos2i(20 + (int)((r >> 5) % (uint32_t)(w - 80)), 20 + (int)((r >> 14) % (uint32_t)(h - 70)));
                pfPixelZoom(1.0f + 0.5f * (float)((r >> 3) & 3u), 1.0f + 0.5f * (float)((r >> 1) & 3u));
                pfDrawPixels(12, 9, PF_RGBA, PF_UNSIGNED_BYTE, aux);
                pfPixelZoom(1.0f, 1.0f);
            }
            break;
        default: {
            const uint32_t sub = (r >> 19) % 8u;
            if (sub < 3) {                       /* vertex arrays: a random subset enabled, indexed or not */
                static float vp[48 * 3], vn[48 * 3], vt[48 * 2]; static uint8_t vc[48 * 4]; static uint16_t vi[60];
                if (persp && FUZZ_FAR() && !mv_far) { pfBindTexture(NULL); tex_bound = 0; }
                for (int v = 0; v < 48; v++) {
                    const float x = lcgf(), y = lcgf(), z = lcgf();
                    if (persp) { const float sc = FUZZ_FAR() ? 0.45f : 1.0f; vp[3 * v] = sc * (5.0f * x - 2.5f); vp[3 * v + 1] = sc * (4.0f * y - 2.0f); vp[3 * v + 2] = sc * (5.0f * z - 2.0f); }
                    else { vp[3 * v] = 20.0f + x * (float)(w - 40); vp[3 * v + 1] = 20.0f + y * (float)(h - 40); vp[3 * v + 2] = -0.9f * z; }
                    vn[3 * v] = lcgf() - 0.5f; vn[3 * v + 1] = lcgf() - 0.5f; vn[3 * v + 2] = lcgf() + 0.2f;
                    vt[2 * v] = 3.0f * lcgf() - 1.0f; vt[2 * v + 1] = 3.0f * lcgf() - 1.0f;
                    for (int k = 0; k < 4; k++) vc[4 * v + k] = (uint8_t)(lcg() >> 24);
                }
                for (int k = 0; k < 60; k++) vi[k] = (uint16_t)((lcg() >> 8) % 48u);
                pfEnable(PF_VERTEX_ARRAY); pfVertexPointer(3, PF_FLOAT, 0, vp);
                if ((r >> 5) & 1u) { pfEnable(PF_NORMAL_ARRAY); pfNormalPointer(PF_FLOAT, 0, vn); }
                if ((r >> 6) & 1u) { pfEnable(PF_TEXTURE_COORD_ARRAY); pfTexCoordPointer(PF_FLOAT, 0, vt); }
                if ((r >> 7) & 1u) { pfEnable(PF_COLOR_ARRAY); pfColorPointer(4, PF_UNSIGNED_BYTE, 0, vc); }
                static const PFdrawmode amodes[4] = { PF_TRIANGLES, PF_QUADS, PF_TRIANGLE_STRIP, PF_TRIANGLE_FAN };
                if ((r >> 8) & 1u) pfDrawElements(amodes[(r >> 9) % 4u], (PFsizei)(12 + (r >> 11) % 48u), PF_UNSIGNED_SHORT, vi);
                else pfDrawArrays(amodes[(r >> 9) % 4u], (PFint)((r >> 11) % 12u), (PFsizei)(12 + (r >> 15) % 24u));
                pfDisable(PF_VERTEX_ARRAY); pfDisable(PF_NORMAL_ARRAY); pfDisable(PF_TEXTURE_COORD_ARRAY); pfDisable(PF_COLOR_ARRAY);
            } else if (sub == 3 && fbo->texture) {     /* render to the 64x64 object, then sample it (2D state from scratch around it) */
                pfMatrixMode(PF_MODELVIEW); for (int k = 0; k < 40; k++) pfPopMatrix();
                (void)pfGetError();
                pfPolygonMode(PF_FRONT_AND_BACK, PF_FILL);
                pfBindFramebuffer(fbo); pfEnable(PF_FRAMEBUFFER);
                pfViewport(0, 0, 56, 56);
                pfMatrixMode(PF_PROJECTION); pfLoadIdentity(); pfOrtho(0.0f, 56.0f, 56.0f, 0.0f, 0.0f, 1.0f);
                pfMatrixMode(PF_MODELVIEW); pfLoadIdentity();
                if ((r >> 5) & 1u) pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
                pfBegin(PF_TRIANGLES);
                for (int v = 0; v < 9; v++) {
                    pfColor4ub((PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(128 + (lcg() >> 25)));
                    pfTexCoord2f(2.0f * lcgf(), 2.0f * lcgf());
                    pfVertex3f(2.0f + 50.0f * lcgf(), 2.0f + 50.0f * lcgf(), -0.9f * lcgf());
                }
                pfEnd();
                pfDisable(PF_FRAMEBUFFER);
                pfViewport(0, 0, (PFsizei)w, (PFsizei)h); ortho2d(w, h);
                persp = 0; clean = 1; mv_far = 1;
                {
                    const int was_tex = tex_enabled;
                    pfEnable(PF_TEXTURE_2D);
                    pfTextureParameter(fbo->texture, PF_REPEAT, PF_NEAREST);
                    draw_textured_quad(fbo->texture, 24.0f + lcgf() * (float)(w - 120), 24.0f + lcgf() * (float)(h - 110), 64.0f, 64.0f, 1.0f, 1.0f, (int)((r >> 6) & 1u));
                    if (!was_tex) pfDisable(PF_TEXTURE_2D);
                    tex_bound = 0;              /* draw_textured_quad leaves no texture bound */
                }
            } else if (sub == 4 && (r >> 5) % 4u == 0) {      /* fog over what has been drawn so far */
                const float fc[4] = { lcgf(), lcgf(), lcgf(), 0.3f + 0.7f * lcgf() };
                pfFogi(PF_FOG_MODE, (PFint)((r >> 8) % 3u)); pfFogf(PF_FOG_START, 0.2f * lcgf()); pfFogf(PF_FOG_END, 0.5f + 2.0f * lcgf());
                pfFogfv(PF_FOG_COLOR, (PFfloat *)fc);
                pfFogProcess();
            } else {
                PFcolor px[8 * 4];
                if (!persp) pfReadPixels(20 + (PFint)((r >> 5) % 64u), 20 + (PFint)((r >> 11) % 64u), 8, 4, PF_RGBA, PF_UNSIGNED_BYTE, px);
            }
            break; }
        }
    }
    for (int li = 0; li < 4; li++) if (lists[li]) pfDeleteList(&lists[li]);
    /* the second context's frame goes into a corner of the first one's, then the context goes away */
    cur = 0; pfMakeCurrent(ctxs[0]);
    for (int k = 0; k < 40; k++) { pfMatrixMode(PF_MODELVIEW); pfPopMatrix(); }
    (void)pfGetError();
    for (PFsizei li = 0; li < 4; li++) pfDisableLight(li);
    pfDisable(PF_BLEND); pfDisable(PF_DEPTH_TEST); pfDisable(PF_CULL_FACE); pfDisable(PF_TEXTURE_2D); pfDisable(PF_LIGHTING);
    pfDisable(PF_NORMALIZE); pfDisable(PF_COLOR_MATERIAL);
    pfPolygonMode(PF_FRONT_AND_BACK, PF_FILL); pfLightModel(PF_GOURAUD); pfShadeModel(PF_SMOOTH);
    pfMatrixMode(PF_TEXTURE); pfLoadIdentity(); pfMatrixMode(PF_MODELVIEW);
    pfBindTexture(NULL); pfViewport(0, 0, (PFsizei)w, (PFsizei)h);
    ortho2d(w, h);
    pfRasterPos2i(w - 136, 8); pfPixelZoom(1.0f, 1.0f);
    pfDrawPixels(128, 96, PF_RGBA, PF_UNSIGNED_BYTE, second);
    pfColor4ub(255, 255, 255, 255);
    pfDeleteContext(ctxs[1]);
    pfMakeCurrent(ctxs[0]);
    free(second);
#undef FUZZ_DIRTY
#undef FUZZ_FAR
#undef w
#undef h
#undef persp
#undef tex_enabled
#undef tex_bound
#undef mv_far
#undef clean
}

static const PFpixelformat target_formats_g[4] = { PF_RGBA, PF_BGRA, PF_RGB, PF_BGR };

/* ---- the runner ------------------------------------------------------------------------------------ */

#define MAX_BATCH_CTX 1024

typedef struct {
    char name[32];
    pfscene_cfg cfg;
    /* single-context scenes */
    PFcontext ctx; uint8_t *target; uint8_t *aux; uint8_t *texpx; PFtexture tex, tex2; mesh_t mesh; PFframebuffer fbo;
    /* "batch": n contexts */
    int n; PFcontext *ctxs; uint8_t **bufs; uint8_t **texpxs; PFtexture *texs; PFrenderlist (*lists)[3];
} scene_t;

static int cmp_double(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }

static void reset_counters(void)
{
#ifdef PFSCENE_HAVE_PFX
    pfxResetCounters();
#endif
}

SCN_API const char *pfscene_backend(void)
{
#ifdef PFSCENE_HAVE_PFX
    return pfxBackendName();
#else
    return "reference";
#endif
}

SCN_API void pfscene_close(void *handle);

/* "prims": points, lines and PF_POINT / PF_LINE polygon modes, interleaved with filled triangles so that the
 * submission order across primitive kinds matters.  variant: bit0 blend, bits1-3 blend mode, bit4 depth test,
 * bits5-7 depth function, bit8 perspective (3D lines are clipped against the frustum), bit9 thick lines and
 * large points.  Everything 2D stays 12 pixels inside the surface: the reference's line loop has no bounds
 * check and thick lines are shifted copies of the centre line. */
static void prims_scene(const pfscene_cfg *cfg)
{
    const int v = cfg->variant, w = cfg->width, h = cfg->height, n = cfg->size > 0 ? cfg->size : 40;
    lcg_state = (uint32_t)cfg->seed * 2654435761u + 12345u;
    pfClearColor(20, 10, 40, 255);
    pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
    const int persp = (v >> 8) & 1, thick = (v >> 9) & 1;
    if (persp) {
        pfViewport(0, 0, (PFsizei)w, (PFsizei)h);
        cam_perspective(60.0, (double)w / h, 0.1, 100.0);
        float eye[3] = { 0.2f, 0.3f, 3.0f }, at[3] = { 0, 0, 0 };
        cam_lookat(eye, at);
    } else ortho2d(w, h);
    if (v & 1) { pfEnable(PF_BLEND); pfBlendFunc((PFblendmode)((v >> 1) & 7)); }
    if (v & 16) { pfEnable(PF_DEPTH_TEST); pfDepthFunc((PFdepthmode)((v >> 5) & 7)); }
    pfDisable(PF_CULL_FACE);
    for (int i = 0; i < n; i++) {
        const int kind = i % 6;
        float x[4], y[4], z[4];
        for (int k = 0; k < 4; k++) {
            if (persp) { x[k] = lcgf() * 5.0f - 2.5f; y[k] = lcgf() * 4.0f - 2.0f; z[k] = lcgf() * 5.0f - 2.0f; }
            else { x[k] = 12.0f + lcgf() * (float)(w - 24); y[k] = 12.0f + lcgf() * (float)(h - 24); z[k] = 0.0f; }
        }
        /* thick 3D lines could be shifted outside the surface after clipping: thin ones only in perspective */
        pfLineWidth(thick && !persp ? 1.0f + (float)(i % 5) : 1.0f);
        pfPointSize(thick ? 1.0f + (float)(i % 7) : 1.0f);
        pfPolygonMode(PF_FRONT_AND_BACK, PF_FILL);
        switch (kind) {
        case 0: pfBegin(PF_POINTS); break;
        case 1: pfBegin(PF_LINES); break;
        case 2: pfBegin(PF_TRIANGLES); break;
        case 3: pfPolygonMode(PF_FRONT_AND_BACK, PF_LINE); pfBegin(PF_TRIANGLES); break;
        case 4: pfPolygonMode(PF_FRONT, PF_POINT); pfPolygonMode(PF_BACK, PF_LINE); pfBegin(PF_QUADS); break;
        default: pfPolygonMode(PF_FRONT, PF_LINE); pfBegin(PF_QUADS); break;        /* back faces stay filled */
        }
        const int nv = kind == 0 ? 3 : (kind == 1 ? 4 : (kind >= 4 ? 4 : 3));
        for (int k = 0; k < nv; k++) {
            pfColor4ub((PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(lcg() >> 24), (PFubyte)(96 + (lcg() >> 25)));
            if (kind >= 2 && !persp) {      /* keep polygons small so that many primitives overlap without filling the screen */
                const float cx = x[0], cy = y[0];
                pfVertex3f(cx + (x[k] - cx) * 0.25f, cy + (y[k] - cy) * 0.25f, z[k]);
            } else pfVertex3f(x[k], y[k], z[k]);
        }
        pfEnd();
    }
    pfPolygonMode(PF_FRONT_AND_BACK, PF_FILL); pfLineWidth(1.0f); pfPointSize(1.0f);
    pfDisable(PF_BLEND); pfDisable(PF_DEPTH_TEST); pfEnable(PF_CULL_FACE);
}

/* Creates the context(s), textures, meshes and fixed state of scene `name`.  Returns NULL on failure. */
SCN_API void *pfscene_open(const char *name, const pfscene_cfg *cfg)
{
    scene_t *s = (scene_t *)calloc(1, sizeof *s);
    if (!s) return NULL;
    strncpy(s->name, name, sizeof s->name - 1);
    s->cfg = *cfg;
    const int w = cfg->width, h = cfg->height;
#ifdef PFSCENE_HAVE_PFX
    pfxSetSyncMode(cfg->explicit_sync ? PF_TRUE : PF_FALSE);
#endif

    if (strcmp(name, "batch") == 0) {
        /* C5: `size` independent contexts, each with its own target, texture and the three gear render
           lists; replayed every frame with a per-context angle.  Textured + Gouraud-lit + depth. */
        const int n = cfg->size > 0 ? cfg->size : 4;
        s->n = n;
        s->ctxs = (PFcontext *)calloc((size_t)n, sizeof(PFcontext));
        s->bufs = (uint8_t **)calloc((size_t)n, sizeof(uint8_t *));
        s->texpxs = (uint8_t **)calloc((size_t)n, sizeof(uint8_t *));
        s->texs = (PFtexture *)calloc((size_t)n, sizeof(PFtexture));
        s->lists = (PFrenderlist (*)[3])calloc((size_t)n, sizeof(PFrenderlist[3]));
        for (int c = 0; c < n; c++) {
            s->bufs[c] = (uint8_t *)calloc((size_t)w * h * 4 + 64, 1);
            s->ctxs[c] = pfCreateContext(s->bufs[c], (PFsizei)w, (PFsizei)h, PF_RGBA, PF_UNSIGNED_BYTE);
            if (!s->ctxs[c]) { fprintf(stderr, "pfscene: pfCreateContext failed\n"); s->n = c; pfscene_close(s); return NULL; }
            pfMakeCurrent(s->ctxs[c]);
            if (cfg->variant & 128) {           /* lists that sample a 5-6-5 texture */
                int tb = 0;
                s->texpxs[c] = make_texture_pair(256, 256, PF_RGB, PF_UNSIGNED_SHORT_5_6_5, (uint32_t)(cfg->seed + c), &tb);
                s->texs[c] = pfGenTexture(s->texpxs[c], 256, 256, PF_RGB, PF_UNSIGNED_SHORT_5_6_5);
            } else {
                s->texpxs[c] = make_texture(256, 256, 4, (uint32_t)(cfg->seed + c), 96, 255, 255, 255);
                s->texs[c] = pfGenTexture(s->texpxs[c], 256, 256, (cfg->variant & 256) ? PF_BGRA : PF_RGBA, PF_UNSIGNED_BYTE);
            }
            gears_setup(w, h);
            pfEnable(PF_TEXTURE_2D);
            static const double gp[3][5] = { { 1.0, 4.0, 1.0, 20, 0.7 }, { 0.5, 2.0, 2.0, 10, 0.7 }, { 1.3, 2.0, 0.5, 10, 0.7 } };
            /* variant bits of "batch": 1 culling off (both face passes), 2 no PF_COLOR_MATERIAL at replay, 4 lists recorded
               with one colour per gear, 8 a texture matrix at replay, 16 front faces culled, 32 per-pixel Phong, 64 the
               first list is recorded again before every frame, 128 5-6-5 textures, 256 BGRA8 textures (both: lists replayed through
               the ordinary path) */
            for (int g = 0; g < 3; g++) {
                s->lists[c][g] = pfGenList();
                pfBindTexture(s->texs[c]);
                if (cfg->variant & 4) pfColor4ub((PFubyte)(90 + 60 * g), (PFubyte)(250 - 70 * g), (PFubyte)(120 + 40 * g), (PFubyte)(255 - 30 * g));
                pfNewList(s->lists[c][g]);
                pfTexCoord2f(0.25f * (float)g, 0.5f);
                gear(gp[g][0], gp[g][1], gp[g][2], (int)gp[g][3], gp[g][4]);
                pfEndList();
            }
            pfColor4ub(255, 255, 255, 255);
        }
        return s;
    }

    /* variant bits 24-25 (every single-context scene): layout of the target buffer - 0 RGBA8, 1 BGRA8, 2 RGB8, 3 BGR8.
       The buffer is always w*h*4 bytes (+ padding: the reference's RGB getter reads 4 bytes per pixel); 3-byte layouts
       use the first w*h*3 of them. */
    const PFpixelformat *target_formats = target_formats_g;
    const PFpixelformat tfmt = target_formats[(cfg->variant >> 24) & 3];
    s->target = (uint8_t *)calloc((size_t)w * h * 4 + 64, 1);
    s->ctx = pfCreateContext(s->target, (PFsizei)w, (PFsizei)h, tfmt, PF_UNSIGNED_BYTE);
    if (!s->ctx) { fprintf(stderr, "pfscene: pfCreateContext failed\n"); free(s->target); free(s); return NULL; }
    pfMakeCurrent(s->ctx);

    if (strcmp(name, "gears") == 0) {
        gears_setup(w, h);
    } else if (strcmp(name, "textured") == 0) {
        const int rgb = (cfg->variant >> 3) & 1;
        s->texpx = make_texture(1024, 1024, rgb ? 3 : 4, (uint32_t)cfg->seed, 0, 255, 128, 255);
        s->tex = pfGenTexture(s->texpx, 1024, 1024, rgb ? PF_RGB : PF_RGBA, PF_UNSIGNED_BYTE);
        pfTextureParameter(s->tex, (PFtexturewrap)(((cfg->variant >> 1) & 3) % 3), (cfg->variant & 1) ? PF_BILINEAR : PF_NEAREST);
        const int nu = cfg->size > 0 ? cfg->size : 256;
        s->mesh = make_torus(nu, nu / 2, 12.0f, 5.0f, ((cfg->variant >> 1) & 3) ? 1.6f : 1.0f);
        pfViewport(0, 0, (PFsizei)w, (PFsizei)h);
        cam_perspective(60.0, (double)w / h, 0.01, 1000.0);
        pfEnable(PF_TEXTURE_2D); pfEnable(PF_DEPTH_TEST);
        if (!(cfg->variant & 16)) { pfEnable(PF_BLEND); pfBlendFunc(PF_BLEND_ALPHA); }
        pfDisable(PF_CULL_FACE);
    } else if (strcmp(name, "phong") == 0) {
        const int n = cfg->size > 0 ? cfg->size : 708;
        s->mesh = make_heightfield(n);
        pfViewport(0, 0, (PFsizei)w, (PFsizei)h);
        cam_perspective(60.0, (double)w / h, 0.01, 1000.0);
        float eye[3] = { 0.0f, 0.0f, 2.6f }, at[3] = { 0, 0, 0 };
        cam_lookat(eye, at);
        float lp[3] = { 2.0f, 3.0f, 4.0f }, amb[3] = { 0.8f, 0.3f, 0.2f }, spec[3] = { 1.0f, 1.0f, 1.0f };
        pfEnable(PF_LIGHTING); pfLightModel(PF_PHONG);
        pfLightfv(PF_LIGHT0, PF_POSITION, lp); pfEnableLight(PF_LIGHT0);
        pfMaterialfv(PF_FRONT_AND_BACK, PF_AMBIENT_AND_DIFFUSE, amb); pfMaterialfv(PF_FRONT_AND_BACK, PF_SPECULAR, spec);
        pfMaterialf(PF_FRONT_AND_BACK, PF_SHININESS, 32.0f);
        pfEnable(PF_DEPTH_TEST); pfEnable(PF_CULL_FACE);
    } else if (strcmp(name, "overdraw") == 0) {
        /* variant: bit0 alpha-blend + depth test (the "4K textured+blended" target scene) instead of
           additive / no depth test; bit1 bilinear; the scene off its most specialised path: bit2 tinted vertex colours
           (a different colour per corner), bit3 CLAMP_TO_EDGE, bit4 RGB8 texture, bit5 two fragment states in one
           batch (odd layers blend additively) */
        const int comps = (cfg->variant & 16) ? 3 : 4;
        s->texpx = make_texture(512, 512, comps, (uint32_t)cfg->seed, 0, 3, (cfg->variant & 1) ? 96 : 0, (cfg->variant & 1) ? 200 : 3);
        if (cfg->variant & 1) { lcg_state = 7u; for (size_t i = 0; i < 512u * 512u; i++) for (int c = 0; c < 3; c++) s->texpx[i * comps + c] = (uint8_t)(lcg() >> 24); }
        s->tex = pfGenTexture(s->texpx, 512, 512, comps == 3 ? PF_RGB : PF_RGBA, PF_UNSIGNED_BYTE);
        pfTextureParameter(s->tex, (cfg->variant & 8) ? PF_CLAMP_TO_EDGE : PF_REPEAT, (cfg->variant & 2) ? PF_BILINEAR : PF_NEAREST);
        ortho2d(w, h);
        pfEnable(PF_TEXTURE_2D); pfEnable(PF_BLEND);
        if (cfg->variant & 1) { pfBlendFunc(PF_BLEND_ALPHA); pfEnable(PF_DEPTH_TEST); pfDepthFunc(PF_LEQUAL); }
        else pfBlendFunc(PF_BLEND_ADD);
    } else if (strcmp(name, "micro") == 0) {
        const int rgb = (cfg->variant >> 17) & 1, bgra = (cfg->variant >> 22) & 1, bgr = (cfg->variant >> 23) & 1;
        s->texpx = make_texture(61, 37, (rgb || bgr) ? 3 : 4, (uint32_t)cfg->seed ^ 0xabcdu, 0, 255, 0, 255);
        s->tex = pfGenTexture(s->texpx, 61, 37, bgr ? PF_BGR : (rgb ? PF_RGB : (bgra ? PF_BGRA : PF_RGBA)), PF_UNSIGNED_BYTE);
        /* the FBO is larger than the 96x80 viewport used to draw into it: for a viewport smaller than the
           MAIN buffer the reference computes vpMax = x+width (one past the last column/row,
           context.c:567-570), so "2D" triangles may touch column 96 / row 80 */
        if (cfg->variant & (1 << 20)) s->fbo = pfGenFramebuffer(104, 88, tfmt, PF_UNSIGNED_BYTE);       /* same layout as the target */
    } else if (strcmp(name, "texfmt") == 0) {
        /* "texfmt": the micro scene sampling a texture whose layout is the pair size = format * 16 + type */
        int tb = 0;
        const PFpixelformat f = (PFpixelformat)(cfg->size >> 4); const PFdatatype t = (PFdatatype)(cfg->size & 15);
        s->texpx = make_texture_pair(53, 29, f, t, (uint32_t)cfg->seed ^ 0x7e57u, &tb);
        s->tex = pfGenTexture(s->texpx, 53, 29, f, t);
        if (cfg->variant & (1 << 27)) {         /* bit 27: quads with a second, RGBA8 texture follow in the same batch */
            s->aux = make_texture(32, 32, 4, (uint32_t)cfg->seed ^ 0x2222u, 0, 255, 128, 255);
            s->tex2 = pfGenTexture(s->aux, 32, 32, PF_RGBA, PF_UNSIGNED_BYTE);
        }
    } else if (strcmp(name, "conform") == 0) {
        s->texpx = make_texture(32, 16, 4, (uint32_t)cfg->seed ^ 0xc0f0u, 0, 255, 0, 255);
        s->tex = pfGenTexture(s->texpx, 32, 16, PF_RGBA, PF_UNSIGNED_BYTE);
        s->fbo = pfGenFramebuffer(128, 64, PF_RGBA, PF_UNSIGNED_BYTE);
        s->aux = (uint8_t *)calloc((size_t)w * h * 4 + 4096 + 64, 1);
    } else if (strcmp(name, "fuzz") == 0) {
        s->texpx = make_texture(48, 40, 4, (uint32_t)cfg->seed ^ 0xf022u, 0, 255, 64, 255);
        s->tex = pfGenTexture(s->texpx, 48, 40, PF_RGBA, PF_UNSIGNED_BYTE);
        s->aux = (uint8_t *)calloc(4096, 1);
        s->fbo = pfGenFramebuffer(64, 64, PF_RGBA, PF_UNSIGNED_BYTE);
    } else if (strcmp(name, "examples") == 0) {
        s->texpx = make_texture(64, 64, 4, (uint32_t)cfg->seed ^ 0xe8a3u, 0, 255, 96, 255);
        s->tex = pfGenTexture(s->texpx, 64, 64, PF_RGBA, PF_UNSIGNED_BYTE);
        s->fbo = pfGenFramebuffer((PFsizei)w, (PFsizei)h, PF_RGBA, PF_UNSIGNED_BYTE);
        s->aux = (uint8_t *)calloc(4096, 1);
        s->mesh = make_torus(48, 24, 12.0f, 5.0f, 2.0f);
    } else if (strcmp(name, "prims") == 0) {
        /* no resources */
    } else if (strcmp(name, "api") == 0) {
        s->texpx = make_texture(64, 48, 4, (uint32_t)cfg->seed ^ 0x5151u, 0, 255, 64, 255);
        s->tex = pfGenTexture(s->texpx, 64, 48, PF_RGBA, PF_UNSIGNED_BYTE);
        pfTextureParameter(s->tex, PF_REPEAT, (cfg->variant & (1 << 18)) ? PF_BILINEAR : PF_NEAREST);
        s->aux = (uint8_t *)calloc((size_t)w * h * 4 + 64, 1);
    } else {
        fprintf(stderr, "pfscene: unknown scene '%s'\n", name);
        pfscene_close(s);
        return NULL;
    }
    return s;
}

/* Issues the API calls of one frame (clear + draw).  Does NOT wait for the GPU. */
SCN_API void pfscene_frame(void *handle, int frame)
{
    scene_t *s = (scene_t *)handle;
    const pfscene_cfg *cfg = &s->cfg;
    const int w = cfg->width, h = cfg->height;
    const char *name = s->name;
    if (strcmp(name, "batch") == 0) {
        for (int c = 0; c < s->n; c++) {
            pfMakeCurrent(s->ctxs[c]);
            float angle = 7.0f * (float)c + 1.44f * (float)frame;
            const int bv = cfg->variant;
            pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
            if (!(bv & 2)) { pfEnable(PF_COLOR_MATERIAL); pfColorMaterial(PF_FRONT_AND_BACK, PF_AMBIENT_AND_DIFFUSE); }
            if (bv & 1) pfDisable(PF_CULL_FACE); else { pfEnable(PF_CULL_FACE); pfCullFace((bv & 16) ? PF_FRONT : PF_BACK); }
            if (bv & 8) { pfMatrixMode(PF_TEXTURE); pfLoadIdentity(); pfScalef(2.0f, 0.5f, 1.0f); pfMatrixMode(PF_MODELVIEW); }
            pfLightModel((bv & 32) ? PF_PHONG : PF_GOURAUD);
            if ((bv & 64) && frame > 0) {
                pfBindTexture(s->texs[c]);
                pfNewList(s->lists[c][0]);
                pfTexCoord2f(0.125f * (float)(frame & 3), 0.25f);
                gear(1.0, 4.0, 1.0 + 0.25 * (frame & 1), 20, 0.7);
                pfEndList();
            }
            pfPushMatrix();
            pfRotatef(20.0f, 1.0f, 0.0f, 0.0f); pfRotatef(30.0f, 0.0f, 1.0f, 0.0f);
            static const float tr[3][2] = { { -3.0f, -2.0f }, { 3.1f, -2.0f }, { -3.1f, 4.2f } };
            static const PFubyte col[3][3] = { { 255, 64, 64 }, { 64, 255, 64 }, { 64, 64, 255 } };
            for (int g = 0; g < 3; g++) {
                pfPushMatrix();
                pfTranslatef(tr[g][0], tr[g][1], 0.0f);
                pfRotatef(g == 0 ? angle : (g == 1 ? -2.0f * angle - 9.0f : -2.0f * angle - 25.0f), 0.0f, 0.0f, 1.0f);
                pfColor3ub(col[g][0], col[g][1], col[g][2]);
                pfCallList(s->lists[c][g]);
                pfPopMatrix();
            }
            pfPopMatrix();
            pfDisable(PF_COLOR_MATERIAL);
            if (bv & 8) { pfMatrixMode(PF_TEXTURE); pfLoadIdentity(); pfMatrixMode(PF_MODELVIEW); }
        }
        return;
    }
    pfMakeCurrent(s->ctx);
    if (strcmp(name, "gears") == 0) {
        gears_frame(1.44f * (float)frame);
    } else if (strcmp(name, "textured") == 0) {
        double t = 0.35 + 0.05 * frame;
        float eye[3] = { (float)(35.0 * cos(t)), 30.0f, (float)(35.0 * sin(t)) }, at[3] = { 0.0f, 0.0f, 0.0f };
        if (cfg->variant & 64) { eye[0] = (float)(19.0 * cos(t)); eye[1] = 9.0f; eye[2] = (float)(19.0 * sin(t)); }   /* close-up: fills the screen */
        cam_lookat(eye, at);
        pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
        pfBindTexture(s->tex);
        pfColor4ub(255, 255, 255, 200);
        if ((cfg->variant & 128) && frame > 0) {
            /* bit 7: the application rewrites the vertex array in place between frames (and announces it when the
               arrays were declared static, pfx.h) */
            for (int k = 0; k < s->mesh.nverts; k++) s->mesh.pos[3 * k + 1] *= 1.0f + 0.02f * (float)((k + frame) & 3);
#ifdef PFSCENE_HAVE_PFX
            pfxHostModified(s->mesh.pos);
#endif
        }
        if (cfg->variant & 32) draw_mesh_arrays(&s->mesh); else draw_mesh_immediate(&s->mesh);
    } else if (strcmp(name, "phong") == 0) {
        pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
        pfColor4ub(255, 255, 255, 255);
        if (cfg->variant & 32) {
            pfEnable(PF_VERTEX_ARRAY); pfEnable(PF_NORMAL_ARRAY);
            pfVertexPointer(3, PF_FLOAT, 0, s->mesh.pos); pfNormalPointer(PF_FLOAT, 0, s->mesh.nrm);
            pfDrawElements(PF_TRIANGLES, (PFsizei)s->mesh.nidx, PF_UNSIGNED_INT, s->mesh.idx);
            pfDisable(PF_VERTEX_ARRAY); pfDisable(PF_NORMAL_ARRAY);
            g_api_tris += (unsigned long long)s->mesh.nidx / 3;
        } else {
            pfBegin(PF_TRIANGLES);
            for (int k = 0; k < s->mesh.nidx; k++) { uint32_t i = s->mesh.idx[k]; pfNormal3fv(s->mesh.nrm + 3 * i); pfVertex3fv(s->mesh.pos + 3 * i); }
            pfEnd();
            g_api_tris += (unsigned long long)s->mesh.nidx / 3;
        }
    } else if (strcmp(name, "overdraw") == 0) {
        const int layers = cfg->size > 0 ? cfg->size : 64;
        pfClearColor(0, 0, 0, 255);
        pfClear((PFclearflag)(PF_COLOR_BUFFER_BIT | PF_DEPTH_BUFFER_BIT));
        pfColor4ub(255, 255, 255, 255);
        /* CLAMP_TO_EDGE: texture coordinates 0..1 over the screen, so that every texel is still visited */
        const float ur = (cfg->variant & 8) ? 1.0f : (float)w / 512.0f, vr = (cfg->variant & 8) ? 1.0f : (float)h / 512.0f;
        for (int l = 0; l < layers; l++) {
            if (cfg->variant & 32) pfBlendFunc((l & 1) ? PF_BLEND_ADD : ((cfg->variant & 1) ? PF_BLEND_ALPHA : PF_BLEND_SCREEN));
            draw_textured_quad(s->tex, 0.0f, 0.0f, (float)w, (float)h, ur, vr, (cfg->variant & 4) != 0);
        }
    } else if (strcmp(name, "micro") == 0) {
        if (cfg->variant & (1 << 20)) {
            /* pass 1: render into the FBO; pass 2: draw it as a texture over the main buffer */
            pfscene_cfg sub = *cfg; sub.width = 96; sub.height = 80; sub.variant &= ~(1 << 20);
            pfBindFramebuffer(&s->fbo); pfEnable(PF_FRAMEBUFFER);
            micro_scene(&sub, s->tex);
            pfDisable(PF_FRAMEBUFFER);
            sub = *cfg; sub.variant &= ~(1 << 20); sub.seed += 17;
            micro_scene(&sub, s->tex);
            ortho2d(w, h);
            pfDisable(PF_LIGHTING); pfDisable(PF_DEPTH_TEST); pfEnable(PF_TEXTURE_2D); pfEnable(PF_BLEND); pfBlendFunc(PF_BLEND_ALPHA);
            pfTextureParameter(s->fbo.texture, PF_REPEAT, PF_NEAREST);   /* CLAMP would index row h of the FBO at v == 1 (reads past the reference's own allocation) */
            pfColor4ub(255, 255, 255, 255);
            draw_textured_quad(s->fbo.texture, 8.0f, 6.0f, 0.7f * w, 0.6f * h, 1.0f, 1.0f, 0);
        } else micro_scene(cfg, s->tex);
    } else if (strcmp(name, "texfmt") == 0) {
        pfscene_cfg sub = *cfg; sub.size = 0; sub.variant |= 512; sub.variant &= ~(1 << 27);
        micro_scene(&sub, s->tex);
        if (s->tex2) {                          /* the batch's states now differ in their texel layout */
            ortho2d(w, h); pfEnable(PF_TEXTURE_2D); pfDisable(PF_CULL_FACE);
            pfTextureParameter(s->tex2, PF_REPEAT, PF_NEAREST);
            draw_textured_quad(s->tex2, 6.0f, 8.0f, 60.0f, 40.0f, 2.0f, 2.0f, 1);
            pfTextureParameter(s->tex, PF_MIRRORED_REPEAT, PF_NEAREST);
            draw_textured_quad(s->tex, 80.0f, 70.0f, 70.0f, 40.0f, 3.0f, 2.0f, 0);
            draw_textured_quad(s->tex2, 100.0f, 20.0f, 50.0f, 30.0f, 1.0f, 1.0f, 0);
            pfDisable(PF_TEXTURE_2D);
        }
    } else if (strcmp(name, "fuzz") == 0) {
        fuzz_scene(cfg, s->tex, s->aux, &s->fbo);
    } else if (strcmp(name, "examples") == 0) {
        examples_scene(cfg, s->tex, &s->fbo, s->aux, &s->mesh, frame);
    } else if (strcmp(name, "conform") == 0) {
        conform_scene(cfg, s->tex, &s->fbo, s->aux, s->target, target_formats_g[(cfg->variant >> 24) & 3]);
    } else if (strcmp(name, "api") == 0) {
        api_scene(cfg, s->tex, s->aux);
    } else if (strcmp(name, "prims") == 0) {
        prims_scene(cfg);
    }
}

/* Waits for the frame and brings the caller-visible buffers up to date (no-op for the reference,
 * which is synchronous). */
SCN_API void pfscene_finish(void *handle)
{
    scene_t *s = (scene_t *)handle;
    if (s->n) { for (int c = 0; c < s->n; c++) { pfMakeCurrent(s->ctxs[c]); finish(); } }
    else { pfMakeCurrent(s->ctx); finish(); }
}

SCN_API void pfscene_make_current(void *handle, int index)
{
    scene_t *s = (scene_t *)handle;
    pfMakeCurrent(s->n ? s->ctxs[index % s->n] : s->ctx);
}

/* Copies out the colour (w*h*4 bytes) and optionally the depth buffer (context size-1 for "batch"). */
/* "batch" only: the caller-visible colour buffer of context `index`, as it is (no API call is made). */
SCN_API int pfscene_read_context(void *handle, int index, uint8_t *color_out)
{
    scene_t *s = (scene_t *)handle;
    if (!s->n || index < 0 || index >= s->n) return 0;
    memcpy(color_out, s->bufs[index], (size_t)s->cfg.width * s->cfg.height * 4);
    return 1;
}

/* Colour and (optionally) depth of context `index` ("batch": 0..n-1; single-context scenes: index 0), through the
 * public API only: the colour is the caller-owned target buffer after the frame was finished, the depth comes from
 * pfxReadDepth (product) / a pfPostProcess pass (reference). */
SCN_API int pfscene_read_index(void *handle, int index, uint8_t *color_out, float *depth_out)
{
    scene_t *s = (scene_t *)handle;
    const size_t bytes = (size_t)s->cfg.width * s->cfg.height * 4;
    if (s->n) {
        if (index < 0 || index >= s->n) return 0;
        pfMakeCurrent(s->ctxs[index]);
        if (color_out) memcpy(color_out, s->bufs[index], bytes);
    } else {
        if (index != 0) return 0;
        pfMakeCurrent(s->ctx);
        if (color_out) memcpy(color_out, s->target, bytes);
    }
    read_depth(depth_out, s->cfg.width);
    return 1;
}

SCN_API void pfscene_read(void *handle, uint8_t *color_out, float *depth_out)
{
    scene_t *s = (scene_t *)handle;
    const size_t bytes = (size_t)s->cfg.width * s->cfg.height * 4;
    if (s->n) { pfMakeCurrent(s->ctxs[s->n - 1]); if (color_out) memcpy(color_out, s->bufs[s->n - 1], bytes); }
    else { pfMakeCurrent(s->ctx); if (color_out) memcpy(color_out, s->target, bytes); }
    read_depth(depth_out, s->cfg.width);
}

SCN_API void pfscene_close(void *handle)
{
    scene_t *s = (scene_t *)handle;
    if (!s) return;
    if (s->ctxs) {
        for (int c = 0; c < s->n; c++) {
            pfMakeCurrent(s->ctxs[c]);
            for (int g = 0; g < 3; g++) if (s->lists[c][g]) pfDeleteList(&s->lists[c][g]);
            if (s->texs[c]) pfDeleteTexture(&s->texs[c], PF_FALSE);
            pfMakeCurrent(NULL);
            pfDeleteContext(s->ctxs[c]);
            free(s->bufs[c]); free(s->texpxs[c]);
        }
        free(s->ctxs); free(s->bufs); free(s->texpxs); free(s->texs); free(s->lists);
    } else if (s->ctx) {
        pfMakeCurrent(s->ctx);
        if (s->fbo.texture) pfDeleteFramebuffer(&s->fbo);
        if (s->tex) pfDeleteTexture(&s->tex, PF_FALSE);
        if (s->tex2) pfDeleteTexture(&s->tex2, PF_FALSE);
        free(s->texpx);
        if (s->mesh.pos) free_mesh(&s->mesh);
        pfMakeCurrent(NULL);
        pfDeleteContext(s->ctx);
        free(s->target); free(s->aux);
    }
    free(s);
}

/* Renders `cfg->warmup + cfg->frames` frames of scene `name` and times each frame with the wall clock
 * (API calls + wait + host buffers up to date).  The colour buffer and, when depth_out != NULL, the
 * depth buffer of the LAST frame are returned.  Counters cover the timed frames only. */
SCN_API int pfscene_render(const char *name, const pfscene_cfg *cfg, uint8_t *color_out, float *depth_out, pfscene_result *res)
{
    memset(res, 0, sizeof *res);
    void *s = pfscene_open(name, cfg);
    if (!s) return 2;
    const int total = cfg->warmup + cfg->frames;
    double *times = (double *)calloc((size_t)(cfg->frames > 0 ? cfg->frames : 1), sizeof(double));
    for (int f = 0; f < total; f++) {
        if (f == cfg->warmup) { reset_counters(); g_api_tris = 0; }
        double t0 = now_ms();
        pfscene_frame(s, cfg->first_frame + f);
        pfscene_finish(s);
        if (f >= cfg->warmup) times[f - cfg->warmup] = now_ms() - t0;
    }
    const int n = cfg->frames;
    res->api_triangles = g_api_tris;
    res->ms_total = 0; res->ms_min = n > 0 ? 1e300 : 0;
    for (int i = 0; i < n; i++) { res->ms_total += times[i]; if (times[i] < res->ms_min) res->ms_min = times[i]; }
    if (n > 0) { qsort(times, (size_t)n, sizeof(double), cmp_double); res->ms_median = times[n / 2]; }
#ifdef PFSCENE_HAVE_PFX
    { PFXcounters k; pfxGetCounters(&k);
      res->triangles_submitted = k.triangles_submitted; res->triangles_rasterised = k.triangles_rasterised;
      res->pixels_shaded = k.pixels_shaded; res->pixels_depth_failed = k.pixels_depth_failed; res->kernel_launches = k.kernel_launches; }
#endif
    pfscene_read(s, color_out, depth_out);
    pfscene_close(s);
    free(times);
    return 0;
}
