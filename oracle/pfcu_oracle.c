/*
 * pfcu_oracle.c - TEST INFRASTRUCTURE ONLY.  Scalar C restatement of the reference's per-fragment
 * triangle path, exposed through the same C-ABI as the product (include/pfcu.h) so that the host
 * state machine can be linked against it in tests.  Nothing in the product links or calls this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Parity status: PINNED.  The reference publishes no golden vectors; this
 * restatement is checked pixel-for-pixel (colour and depth) against the reference itself compiled
 * from /root/reference by oracle/build_ref.sh (tests/test_oracle_vs_ref.py) and against the
 * committed fixtures in tests/golden/ generated from that build.
 *
 * One pixel at a time, same operation order and roundings as the AVX2 lanes of
 *   src/internal/primitives/triangles.c:287-558  (Rasterize_Triangle, PF_TRIANGLE_TRAVEL_SIMD)
 *   src/internal/color.h:153-203                 (colour interpolation)
 *   src/internal/sampler.h:202-410               (wrap modes, nearest / bilinear)
 *   src/internal/pixel.h:2060-2078,2604-2632,2909-2920 (RGBA8/BGRA8/RGB8 get/set)
 *   src/internal/blend.h:137-274, depth.h:80-124, lighting/lighting.c:148-258
 *   src/internal/simd.h:183-304,1157-1245        (cephes log/exp, pow, rcp, rsqrt)
 * RCPPS/RSQRTPS are executed natively (_mm_rcp_ss / _mm_rsqrt_ss use the same hardware table).
 * Build with -ffp-contract=off and without -mfma / -ffast-math.
 */
#include "../include/pfcu.h"

#include <immintrin.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <limits.h>
#include "../pixelforge_b200/csrc/pf_pixfmt.h"

/* colour is held as canonical RGBA8 dwords whatever the caller's layout (fmt); upload / download convert */
struct pfcu_surface { uint32_t w, h; uint32_t *color; float *depth; int owned; uint32_t rank, world; int fmt; };
/* leader: texels come through the reference's BGRA8 getter, whose shuffle hands the first texel of every group of four
   pixels to all four (pixel.h:2915-2920, simd.h:563-583) */
struct pfcu_texture { uint32_t w, h; int fmt; uint8_t *pixels; int owned; pfcu_surface *alias; int leader; };
struct pfcu_batch   { pfcu_state *states; uint32_t n_states; pfcu_triangle *tris; uint32_t n_tris; };

static pfcu_counters g_cnt;
static char g_err[256];

/* ---- x86 SIMD lane semantics, scalar -------------------------------------------------------- */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float    u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static inline float rcp_x86(float x)   { return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x))); }    /* simd.h:1217-1225 */
static inline float rsqrt_x86(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }  /* simd.h:1237-1245 */
/* MINPS/MAXPS return the SECOND operand when either is NaN (and for equal zeros). */
static inline float min_x86(float a, float b) { return a < b ? a : b; }
static inline float max_x86(float a, float b) { return a > b ? a : b; }
static inline float clamp_x86(float x, float lo, float hi) { return min_x86(max_x86(x, lo), hi); } /* simd.h:1086-1093 */
/* CVTPS2DQ: round to nearest even, 0x80000000 when out of range or NaN (simd.h:921-929). */
static inline int32_t cvt_rne(float x) { return _mm_cvtss_si32(_mm_set_ss(x)); }
static inline int32_t cvt_trunc(float x) { return _mm_cvttss_si32(_mm_set_ss(x)); }

/* _mm256_log_ps, simd.h:183-252 */
static float log_cephes(float x)
{
    int invalid = (x <= 0.0f);
    x = max_x86(x, u2f(0x00800000u));
    int32_t imm0 = (int32_t)(f2u(x) >> 23);
    x = u2f((f2u(x) & ~0x7f800000u) | f2u(0.5f));
    imm0 -= 0x7f;
    float e = (float)imm0;
    e = e + 1.0f;
    int lt = (x < (float)0.707106781186547524);
    float tmp = lt ? x : 0.0f;
    x = x - 1.0f;
    e = e - (lt ? 1.0f : 0.0f);
    x = x + tmp;
    float z = x * x;
    float y = (float)7.0376836292E-2;
    y = y * x; y = y + (float)-1.1514610310E-1;
    y = y * x; y = y + (float)1.1676998740E-1;
    y = y * x; y = y + (float)-1.2420140846E-1;
    y = y * x; y = y + (float)+1.4249322787E-1;
    y = y * x; y = y + (float)-1.6668057665E-1;
    y = y * x; y = y + (float)+2.0000714765E-1;
    y = y * x; y = y + (float)-2.4999993993E-1;
    y = y * x; y = y + (float)+3.3333331174E-1;
    y = y * x;
    y = y * z;
    tmp = e * (float)-2.12194440e-4;
    y = y + tmp;
    tmp = z * 0.5f;
    y = y - tmp;
    tmp = e * (float)0.693359375;
    x = x + y;
    x = x + tmp;
    if (invalid) x = u2f(0xffffffffu);
    return x;
}

/* _mm256_exp_ps, simd.h:254-304 */
static float exp_cephes(float x)
{
    x = min_x86(x, 88.3762626647949f);
    x = max_x86(x, -88.3762626647949f);
    float fx = x * (float)1.44269504088896341;
    fx = fx + 0.5f;
    float tmp = floorf(fx);
    float mask = (tmp > fx) ? 1.0f : 0.0f;
    fx = tmp - mask;
    tmp = fx * (float)0.693359375;
    float z = fx * (float)-2.12194440e-4;
    x = x - tmp;
    x = x - z;
    z = x * x;
    float y = (float)1.9875691500E-4;
    y = y * x; y = y + (float)1.3981999507E-3;
    y = y * x; y = y + (float)8.3334519073E-3;
    y = y * x; y = y + (float)4.1665795894E-2;
    y = y * x; y = y + (float)1.6666665459E-1;
    y = y * x; y = y + (float)5.0000001201E-1;
    y = y * z;
    y = y + x;
    y = y + 1.0f;
    int32_t imm0 = cvt_trunc(fx);
    imm0 = (int32_t)((uint32_t)imm0 + 0x7fu);
    float pow2n = u2f((uint32_t)imm0 << 23);
    return y * pow2n;
}

static inline float pow_x86(float base, float e) { return exp_cephes(log_cephes(base) * e); }  /* simd.h:1157-1169 */

/* ---- colour helpers ------------------------------------------------------------------------- */

#define CH(c, i) ((int32_t)(((c) >> (8 * (i))) & 255u))     /* pfiColorSIMDToVecI_simd, color.h:77-83 */

static inline uint32_t pack_i(const int32_t *v, int n)     /* pfiColorSIMDFromVecI_simd, color.h:112-122 */
{
    uint32_t p = 0;
    for (int i = 0; i < n; i++) p |= (uint32_t)v[i] << (8 * i);
    return p;
}

static inline uint32_t pack_f(const float *v, int n)       /* pfiColorSIMDFromVecF_simd, color.h:124-135 */
{
    uint32_t p = 0;
    for (int i = 0; i < n; i++) {
        float c = clamp_x86(v[i], 0.0f, 1.0f) * 255.0f;
        p |= (uint32_t)cvt_rne(c) << (8 * i);
    }
    return p;
}

static const float INV255 = 1.0f / 255.0f;

/* pfiColorBarySmooth_simd, color.h:153-181 */
static uint32_t color_smooth(uint32_t c1, uint32_t c2, uint32_t c3, float w1, float w2, float w3)
{
    int32_t u1 = cvt_rne(w1 * 255.0f), u2 = cvt_rne(w2 * 255.0f), u3 = cvt_rne(w3 * 255.0f);
    int32_t r[4];
    for (int i = 0; i < 4; i++) {
        uint32_t s = (uint32_t)u1 * (uint32_t)CH(c1, i) + (uint32_t)u2 * (uint32_t)CH(c2, i);
        s += (uint32_t)u3 * (uint32_t)CH(c3, i);
        r[i] = (int32_t)((s * 257u) >> 16);     /* logical shift, simd.h:1326 */
    }
    return pack_i(r, 4);
}

/* pfiColorBaryFlat_simd, color.h:183-203 */
static uint32_t color_flat(uint32_t c1, uint32_t c2, uint32_t c3, float w1, float w2, float w3)
{
    float m = max_x86(w1, max_x86(w2, w3));
    return ((m == w1) ? c1 : 0u) | ((m == w2) ? c2 : 0u) | ((m == w3) ? c3 : 0u);
}

/* pfiColorLerpSmooth_simd with the Q7 one-token fix (color.h:137-144), pfiVec4LerpR_simd simd.h:3147 */
static uint32_t color_lerp(uint32_t a, uint32_t b, float t)
{
    float r[4];
    for (int i = 0; i < 4; i++) {
        float A = (float)CH(a, i) * INV255, B = (float)CH(b, i) * INV255;
        r[i] = A + t * (B - A);
    }
    return pack_f(r, 4);
}

/* ---- blending, blend.h:137-274 -------------------------------------------------------------- */

static uint32_t blend(int mode, uint32_t src, uint32_t dst)
{
    int32_t o[4];
    switch (mode) {
    case 0: for (int i = 0; i < 4; i++) o[i] = (int32_t)((uint32_t)(CH(src, i) + CH(dst, i)) >> 1); break;
    case 1: {
        int32_t alpha = CH(src, 3) + 1, inv = 256 - alpha;
        for (int i = 0; i < 3; i++) o[i] = (int32_t)((uint32_t)(CH(src, i) * alpha + CH(dst, i) * inv) >> 8);
        o[3] = (int32_t)((uint32_t)(255 * alpha + CH(dst, 3) * inv) >> 8);
    } break;
    case 2: for (int i = 0; i < 4; i++) { int32_t s = CH(src, i) + CH(dst, i); o[i] = s < 255 ? s : 255; } break;
    case 3: for (int i = 0; i < 4; i++) { int32_t s = CH(src, i) + CH(dst, i); o[i] = s > 0 ? s : 0; } break;   /* sic: adds (Q6) */
    case 4: for (int i = 0; i < 4; i++) o[i] = (int32_t)((uint32_t)(CH(src, i) * CH(dst, i)) >> 8); break;
    case 5: for (int i = 0; i < 4; i++) {
        int32_t s = (int32_t)((uint32_t)(CH(dst, i) * (255 - CH(src, i))) >> 8) + CH(src, i); o[i] = s < 255 ? s : 255; } break;
    case 6: for (int i = 0; i < 4; i++) o[i] = CH(src, i) > CH(dst, i) ? CH(src, i) : CH(dst, i); break;
    default: for (int i = 0; i < 4; i++) o[i] = CH(src, i) < CH(dst, i) ? CH(src, i) : CH(dst, i); break;
    }
    return pack_i(o, 4);
}

static inline uint32_t mul_color(uint32_t a, uint32_t b)   /* pfiBlendMultiplicative_simd, blend.h:199-212 */
{
    return blend(4, a, b);
}

/* ---- depth compare, depth.h:80-114 (ordered compares: false on NaN; NOTEQUAL == EQUAL, Q5) ---- */

static inline int depth_pass(int func, float z, float zb)
{
    switch (func) {
    case 0: return z == zb;
    case 1: return z == zb;
    case 2: return z <  zb;
    case 3: return z <= zb;
    case 4: return z >  zb;
    default: return z >= zb;
    }
}

/* ---- texturing, sampler.h:202-410 ------------------------------------------------------------ */

static inline int32_t abs_i32(int32_t v) { return v < 0 ? (int32_t)(0u - (uint32_t)v) : v; }   /* VPABSD */

static void tex_map(int wrap, uint32_t tw, uint32_t th, float u, float v, int32_t *x, int32_t *y)
{
    if (wrap == 0) {            /* REPEAT, sampler.h:202-218 */
        u = (u - truncf(u)) * (float)(tw - 1u);
        v = (v - truncf(v)) * (float)(th - 1u);
        *x = abs_i32(cvt_rne(u));
        *y = abs_i32(cvt_rne(v));
    } else if (wrap == 1) {     /* MIRRORED_REPEAT, sampler.h:220-240; mod = simd.h:1182-1195 */
        float au = u2f(f2u(u) & 0x7fffffffu), av = u2f(f2u(v) & 0x7fffffffu);
        float mu = au - floorf(au / 2.0f) * 2.0f;
        float mv = av - floorf(av / 2.0f) * 2.0f;
        float ru = 1.0f - (mu - 1.0f), rv = 1.0f - (mv - 1.0f);
        if (mu > 1.0f) mu = ru;
        if (mv > 1.0f) mv = rv;
        mu = mu * (float)(tw - 1u);
        mv = mv * (float)(th - 1u);
        *x = cvt_rne(mu + 0.5f);
        *y = cvt_rne(mv + 0.5f);
    } else {                    /* CLAMP_TO_EDGE, sampler.h:242-255 */
        float cu = clamp_x86(u, 0.0f, 1.0f) * (float)(tw - 1u);
        float cv = clamp_x86(v, 0.0f, 1.0f) * (float)(th - 1u);
        *x = cvt_rne(cu + 0.5f);
        *y = cvt_rne(cv + 0.5f);
    }
}

static uint32_t tex_fetch(const pfcu_texture *t, int32_t x, int32_t y)
{
    /* offsets = y*w + x in wrapping int32 (sampler.h:265-268), gather with a signed index */
    int32_t off = (int32_t)((uint32_t)y * t->w + (uint32_t)x);
    const uint8_t *base = t->alias ? (const uint8_t *)t->alias->color : t->pixels;
    size_t n = (size_t)t->w * t->h;
    /* the reference would read out of bounds here (e.g. CLAMP_TO_EDGE rounds v*(h-1)+0.5 up to row h);
       defined as "memory after the texture reads as zero": RGBA 0, and alpha 255 for 3-byte formats */
    if (t->fmt >= PFCU_TEX_PIX) {          /* the other texel layouts: the SIMD getters of pixel.h:2249-3040, restated in pf_pixfmt.h */
        static const uint8_t zeros[16];
        return (off < 0 || (size_t)off >= n) ? pfx_tex_get(zeros, 0u, t->fmt - PFCU_TEX_PIX) : pfx_tex_get(base, (uint32_t)off, t->fmt - PFCU_TEX_PIX);
    }
    if (off < 0 || (size_t)off >= n) return (t->fmt == PFCU_TEX_RGB8 || t->fmt == PFCU_TEX_BGR8) ? 0xff000000u : 0u;
    uint32_t raw;
    switch (t->fmt) {
    case PFCU_TEX_RGBA8: memcpy(&raw, base + 4 * (size_t)off, 4); return raw;                 /* pixel.h:2909-2913 */
    case PFCU_TEX_BGRA8: memcpy(&raw, base + 4 * (size_t)off, 4);                             /* pixel.h:2915-2920 */
        return (raw & 0xff00ff00u) | ((raw & 0xffu) << 16) | ((raw >> 16) & 0xffu);
    case PFCU_TEX_RGB8: {                                                                       /* pixel.h:2604-2632 */
        const uint8_t *p = base + 3 * (size_t)off;
        return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | 0xff000000u; }
    default: {                                                                                  /* BGR8 */
        const uint8_t *p = base + 3 * (size_t)off;
        return (uint32_t)p[2] | ((uint32_t)p[1] << 8) | ((uint32_t)p[0] << 16) | 0xff000000u; }
    }
}

/* The sampler in two steps, so that the caller can substitute another pixel's taps (the BGRA8 getter's behaviour):
 * taps[0] (nearest) or taps[0..3] = texels at (x0,y0) (x1,y0) (x0,y1) (x1,y1), plus the filter weights. */
static void tex_taps(const pfcu_state *st, float u, float v, uint32_t taps[4], float *fx, float *fy)
{
    const pfcu_texture *t = st->texture;
    uint32_t tw = t->w, th = t->h;
    int32_t x0, y0, x1, y1;
    tex_map(st->tex_wrap, tw, th, u, v, &x0, &y0);
    taps[0] = tex_fetch(t, x0, y0);
    *fx = *fy = 0.0f;
    if (st->tex_filter == 0) return;
    float tx = 1.0f / (float)tw, ty = 1.0f / (float)th;
    tex_map(st->tex_wrap, tw, th, u + tx, v + ty, &x1, &y1);
    *fx = clamp_x86(u * (float)tw - (float)x0, 0.0f, 1.0f);
    *fy = clamp_x86(v * (float)th - (float)y0, 0.0f, 1.0f);
    taps[1] = tex_fetch(t, x1, y0); taps[2] = tex_fetch(t, x0, y1); taps[3] = tex_fetch(t, x1, y1);
}

static uint32_t tex_combine(const pfcu_state *st, const uint32_t taps[4], float fx, float fy)
{
    if (st->tex_filter == 0) return taps[0];
    return color_lerp(color_lerp(taps[0], taps[1], fx), color_lerp(taps[2], taps[3], fx), fy);
}

static uint32_t tex_sample(const pfcu_state *st, float u, float v)
{
    const pfcu_texture *t = st->texture;
    uint32_t tw = t->w, th = t->h;
    int32_t x0, y0;
    tex_map(st->tex_wrap, tw, th, u, v, &x0, &y0);
    if (st->tex_filter == 0) return tex_fetch(t, x0, y0);
    /* bilinear, sampler.h:304-410 */
    int32_t x1, y1;
    float tx = 1.0f / (float)tw, ty = 1.0f / (float)th;      /* texture.c:55-56 */
    tex_map(st->tex_wrap, tw, th, u + tx, v + ty, &x1, &y1);
    float fx = u * (float)tw - (float)x0;
    float fy = v * (float)th - (float)y0;
    fx = clamp_x86(fx, 0.0f, 1.0f);
    fy = clamp_x86(fy, 0.0f, 1.0f);
    uint32_t c00 = tex_fetch(t, x0, y0), c10 = tex_fetch(t, x1, y0);
    uint32_t c01 = tex_fetch(t, x0, y1), c11 = tex_fetch(t, x1, y1);
    uint32_t c0 = color_lerp(c00, c10, fx);
    uint32_t c1 = color_lerp(c01, c11, fx);
    return color_lerp(c0, c1, fy);
}

/* ---- per-fragment Phong, lighting/lighting.c:148-258 ---------------------------------------- */

static inline float dot3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; } /* simd.h:2356 */

static void direction3(float d[3], const float a[3], const float b[3])   /* simd.h:2469-2491 */
{
    d[0] = a[0] - b[0]; d[1] = a[1] - b[1]; d[2] = a[2] - b[2];
    float l2 = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
    l2 = max_x86(l2, 1e-5f);
    float inv = rsqrt_x86(l2);
    d[0] = d[0] * inv; d[1] = d[1] * inv; d[2] = d[2] * inv;
}

static uint32_t phong(uint32_t frag, const pfcu_state *st, int face, const float P[3], const float N[3])
{
    const pfcu_material *m = &st->material[face];
    float D[3], A[3], S[3], V[3], acc[3] = { 0.0f, 0.0f, 0.0f };
    for (int i = 0; i < 3; i++) {
        D[i] = (float)CH(frag, i) * INV255;
        A[i] = ((float)CH(m->ambient, i) * INV255) * D[i];
        S[i] = (float)CH(m->specular, i) * INV255;
    }
    direction3(V, st->view_pos, P);
    for (uint32_t li = 0; li < st->n_lights; li++) {
        const pfcu_light *l = &st->lights[li];
        float La[3], Ld[3], Ls[3], L[3], amb[3], dif[3], spc[3];
        for (int i = 0; i < 3; i++) {
            La[i] = (float)CH(l->ambient, i) * INV255;
            Ld[i] = (float)CH(l->diffuse, i) * INV255;
            Ls[i] = (float)CH(l->specular, i) * INV255;
        }
        direction3(L, l->position, P);
        for (int i = 0; i < 3; i++) amb[i] = La[i] * A[i];
        float diff = max_x86(dot3(N, L), 0.0f);
        for (int i = 0; i < 3; i++) dif[i] = (Ld[i] * diff) * D[i];
        float H[3] = { L[0] + V[0], L[1] + V[1], L[2] + V[2] };
        float h2 = (H[0] * H[0] + H[1] * H[1]) + H[2] * H[2];        /* simd.h:2292-2310: no epsilon */
        float hinv = rsqrt_x86(h2);
        H[0] = H[0] * hinv; H[1] = H[1] * hinv; H[2] = H[2] * hinv;
        float spec = max_x86(dot3(N, H), 0.0f);
        spec = pow_x86(spec, m->shininess);
        for (int i = 0; i < 3; i++) spc[i] = (Ls[i] * spec) * S[i];
        if (l->inner_cutoff < (float)3.14159265358979323846) {
            float nd[3] = { 0.0f - l->direction[0], 0.0f - l->direction[1], 0.0f - l->direction[2] };
            float theta = dot3(L, nd);
            float eps = l->inner_cutoff - l->outer_cutoff;
            float in = (theta - l->outer_cutoff) / eps;
            in = clamp_x86(in, 0.0f, 1.0f);
            for (int i = 0; i < 3; i++) { dif[i] = dif[i] * in; spc[i] = spc[i] * in; }
        }
        if (l->att_linear != 0.0f || l->att_quadratic != 0.0f) {
            float d0 = l->position[0] - P[0], d1 = l->position[1] - P[1];
            float d2 = l->position[1] - P[1];                        /* sic: y twice (Q10, simd.h:2429) */
            float dsq = d0 * d0 + (d1 * d1 + d2 * d2);
            float dist = sqrtf(dsq);
            float att = rcp_x86(l->att_constant + (l->att_linear * dist + l->att_quadratic * dsq));
            for (int i = 0; i < 3; i++) { amb[i] = amb[i] * att; dif[i] = dif[i] * att; spc[i] = spc[i] * att; }
        }
        for (int i = 0; i < 3; i++) { acc[i] = acc[i] + amb[i]; acc[i] = acc[i] + dif[i]; acc[i] = acc[i] + spc[i]; }
    }
    return pack_f(acc, 3);      /* alpha byte = 0, emission unused (Q9) */
}

/* ---- Rasterize_Triangle, triangles.c:287-558 -------------------------------------------------- */

static inline int32_t to_int_x86(float f) { return cvt_trunc(f); }   /* (PFint)float == CVTTSS2SI */
static inline int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }
static inline int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
static inline int32_t iclamp(int32_t x, int32_t lo, int32_t hi) { return x < lo ? lo : (x > hi ? hi : x); }
#define WMUL(a, b) ((int32_t)((uint32_t)(a) * (uint32_t)(b)))
#define WADD(a, b) ((int32_t)((uint32_t)(a) + (uint32_t)(b)))
#define WSUB(a, b) ((int32_t)((uint32_t)(a) - (uint32_t)(b)))

static void raster_triangle(pfcu_surface *s, const pfcu_state *st, const pfcu_triangle *t)
{
    const pfcu_vertex *v1 = &t->v[0], *v2 = &t->v[1], *v3 = &t->v[2];
    int face = t->face;
    int32_t x1 = to_int_x86(v1->sx), y1 = to_int_x86(v1->sy);
    int32_t x2 = to_int_x86(v2->sx), y2 = to_int_x86(v2->sy);
    int32_t x3 = to_int_x86(v3->sx), y3 = to_int_x86(v3->sy);

    g_cnt.triangles_submitted++;
    float area = (float)WSUB(WMUL(WSUB(x2, x1), WSUB(y3, y1)), WMUL(WSUB(x3, x1), WSUB(y2, y1)));
    if ((face == 0 && area >= 0) || (face == 1 && area <= 0)) return;
    g_cnt.triangles_rasterised++;

    int32_t xMin = imin(x1, imin(x2, x3)), yMin = imin(y1, imin(y2, y3));
    int32_t xMax = imax(x1, imax(x2, x3)), yMax = imax(y1, imax(y2, y3));
    if (!t->is3d) {
        xMin = iclamp(xMin, st->vp_min[0], st->vp_max[0]);
        yMin = iclamp(yMin, st->vp_min[1], st->vp_max[1]);
        xMax = iclamp(xMax, st->vp_min[0], st->vp_max[0]);
        yMax = iclamp(yMax, st->vp_min[1], st->vp_max[1]);
    }
    int32_t w1X = WSUB(y3, y2), w1Y = WSUB(x2, x3);
    int32_t w2X = WSUB(y1, y3), w2Y = WSUB(x3, x1);
    int32_t w3X = WSUB(y2, y1), w3Y = WSUB(x1, x2);
    if (face == 1) {
        w1X = WSUB(0, w1X); w1Y = WSUB(0, w1Y); w2X = WSUB(0, w2X);
        w2Y = WSUB(0, w2Y); w3X = WSUB(0, w3X); w3Y = WSUB(0, w3Y);
    }
    int32_t w1R = WADD(WMUL(WSUB(xMin, x2), w1X), WMUL(w1Y, WSUB(yMin, y2)));
    int32_t w2R = WADD(WMUL(WSUB(xMin, x3), w2X), WMUL(w2Y, WSUB(yMin, y3)));
    int32_t w3R = WADD(WMUL(WSUB(xMin, x1), w3X), WMUL(w3Y, WSUB(yMin, y1)));
    float invSum = 1.0f / (float)WADD(WADD(w1R, w2R), w3R);

    int depth_on = (st->flags & PFCU_ST_DEPTH_TEST) != 0;
    int blend_on = (st->flags & PFCU_ST_BLEND) != 0;
    int tex_on = (st->flags & PFCU_ST_TEXTURE) && st->texture;
    int phong_on = (st->flags & PFCU_ST_PHONG) && st->n_lights > 0;
    int smooth = (st->flags & PFCU_ST_SMOOTH) != 0;

    /* the reference would index outside the surface here; the restatement (and the product) clip */
    int32_t ya = imax(yMin, 0), yb = imin(yMax, (int32_t)s->h - 1);
    int32_t xa = imax(xMin, 0), xb = imin(xMax - 1, (int32_t)s->w - 1);   /* x < xMax (Q4) */
    for (int32_t y = ya; y <= yb; y++) {
        if (s->world > 1) { /* tile-split emulation for CPU tests: 64x64 tiles */ }
        for (int32_t x = xa; x <= xb; x++) {
            if (s->world > 1) {
                uint32_t tiles_x = (s->w + 63u) / 64u;
                uint32_t tile = (uint32_t)x / 64u + ((uint32_t)y / 64u) * tiles_x;
                if (tile % s->world != s->rank) continue;
            }
            int32_t w1 = WADD(WADD(w1R, WMUL(WSUB(y, yMin), w1Y)), WMUL(WSUB(x, xMin), w1X));
            int32_t w2 = WADD(WADD(w2R, WMUL(WSUB(y, yMin), w2Y)), WMUL(WSUB(x, xMin), w2X));
            int32_t w3 = WADD(WADD(w3R, WMUL(WSUB(y, yMin), w3Y)), WMUL(WSUB(x, xMin), w3X));
            if (!((w1 | w2 | w3) > 0)) continue;
            float W1 = (float)w1 * invSum, W2 = (float)w2 * invSum, W3 = (float)w3 * invSum;
            float z = rcp_x86((v1->zinv * W1 + v2->zinv * W2) + v3->zinv * W3);
            size_t idx = (size_t)y * s->w + (size_t)x;
            if (depth_on && !depth_pass(st->depth_func, z, s->depth[idx])) { g_cnt.pixels_depth_failed++; continue; }

            uint32_t frag = smooth ? color_smooth(v1->rgba, v2->rgba, v3->rgba, W1, W2, W3)
                                   : color_flat(v1->rgba, v2->rgba, v3->rgba, W1, W2, W3);
            if (tex_on) {
                float u = (v1->u * W1 + v2->u * W2) + v3->u * W3;
                float v = (v1->v * W1 + v2->v * W2) + v3->v * W3;
                if (t->is3d) { u = u * z; v = v * z; }
                frag = mul_color(tex_sample(st, u, v), frag);
            }
            if (phong_on) {
                float N[3] = { (v1->nx * W1 + v2->nx * W2) + v3->nx * W3,
                               (v1->ny * W1 + v2->ny * W2) + v3->ny * W3,
                               (v1->nz * W1 + v2->nz * W2) + v3->nz * W3 };
                float P[3] = { (v1->px * W1 + v2->px * W2) + v3->px * W3,
                               (v1->py * W1 + v2->py * W2) + v3->py * W3,
                               (v1->pz * W1 + v2->pz * W2) + v3->pz * W3 };
                frag = phong(frag, st, face, P, N);
            }
            if (blend_on) frag = blend(st->blend_mode, frag, s->color[idx]);
            s->color[idx] = frag;
            s->depth[idx] = z;          /* written even when the depth test is off (Q11) */
            g_cnt.pixels_shaded++;
        }
    }
}

/* Rasterize_Triangle as the AVX2 lanes really behave for BGRA8, used for render targets other than RGBA8
 * and for textures read through the BGRA8 getter.  The row of the bounding box is walked in groups of FOUR pixels
 * from xMin (one 128-bit half of the reference's 8-pixel step): every lane runs the whole fragment program, covered
 * or not (triangles.c:404-445, 500-529); with a BGRA8 texture lanes 1..3 sample lane 0's texels (taps), with a BGRA8
 * target they blend against lane 0's pixel and store lane 0's final fragment (pixel.h:2069-2078, 2915-2920).  Reads
 * of a group precede its writes, as in the vector code.  Pixels outside the surface (the reference would run off its
 * buffer) read as 0 and are never written. */
static void raster_triangle_rows(pfcu_surface *s, const pfcu_state *st, const pfcu_triangle *t)
{
    const pfcu_vertex *v1 = &t->v[0], *v2 = &t->v[1], *v3 = &t->v[2];
    int face = t->face;
    int32_t x1 = to_int_x86(v1->sx), y1 = to_int_x86(v1->sy);
    int32_t x2 = to_int_x86(v2->sx), y2 = to_int_x86(v2->sy);
    int32_t x3 = to_int_x86(v3->sx), y3 = to_int_x86(v3->sy);

    g_cnt.triangles_submitted++;
    float area = (float)WSUB(WMUL(WSUB(x2, x1), WSUB(y3, y1)), WMUL(WSUB(x3, x1), WSUB(y2, y1)));
    if ((face == 0 && area >= 0) || (face == 1 && area <= 0)) return;
    g_cnt.triangles_rasterised++;

    int32_t xMin = imin(x1, imin(x2, x3)), yMin = imin(y1, imin(y2, y3));
    int32_t xMax = imax(x1, imax(x2, x3)), yMax = imax(y1, imax(y2, y3));
    if (!t->is3d) {
        xMin = iclamp(xMin, st->vp_min[0], st->vp_max[0]); yMin = iclamp(yMin, st->vp_min[1], st->vp_max[1]);
        xMax = iclamp(xMax, st->vp_min[0], st->vp_max[0]); yMax = iclamp(yMax, st->vp_min[1], st->vp_max[1]);
    }
    int32_t w1X = WSUB(y3, y2), w1Y = WSUB(x2, x3);
    int32_t w2X = WSUB(y1, y3), w2Y = WSUB(x3, x1);
    int32_t w3X = WSUB(y2, y1), w3Y = WSUB(x1, x2);
    if (face == 1) {
        w1X = WSUB(0, w1X); w1Y = WSUB(0, w1Y); w2X = WSUB(0, w2X);
        w2Y = WSUB(0, w2Y); w3X = WSUB(0, w3X); w3Y = WSUB(0, w3Y);
    }
    int32_t w1R = WADD(WMUL(WSUB(xMin, x2), w1X), WMUL(w1Y, WSUB(yMin, y2)));
    int32_t w2R = WADD(WMUL(WSUB(xMin, x3), w2X), WMUL(w2Y, WSUB(yMin, y3)));
    int32_t w3R = WADD(WMUL(WSUB(xMin, x1), w3X), WMUL(w3Y, WSUB(yMin, y1)));
    float invSum = 1.0f / (float)WADD(WADD(w1R, w2R), w3R);

    int depth_on = (st->flags & PFCU_ST_DEPTH_TEST) != 0;
    int blend_on = (st->flags & PFCU_ST_BLEND) != 0;
    int tex_on = (st->flags & PFCU_ST_TEXTURE) && st->texture;
    int phong_on = (st->flags & PFCU_ST_PHONG) && st->n_lights > 0;
    int smooth = (st->flags & PFCU_ST_SMOOTH) != 0;
    int tex_leader = tex_on && st->texture->leader, fb_leader = s->fmt == PFCU_TEX_BGRA8;
    uint32_t alpha_or = s->fmt >= PFCU_TEX_RGB8 ? 0xff000000u : 0u;

    if (!(xMin < xMax && yMin <= yMax && xMax > 0 && yMax >= 0 && xMin < (int32_t)s->w && yMin < (int32_t)s->h)) return;   /* nothing on the surface */
    int32_t ya = imax(yMin, 0), yb = imin(yMax, (int32_t)s->h - 1);
    long long gx0 = xMin < 0 ? (long long)xMin + ((-(long long)xMin) & ~3LL) : (long long)xMin;
    long long gxl = xMax < (int32_t)s->w - 1 ? xMax : (long long)s->w - 1;
    for (int32_t y = ya; y <= yb; y++) {
        for (long long gx = gx0; gx <= gxl; gx += 4) {
            int m[4], cov[4], ins[4]; float z[4]; uint32_t frag[4], dst[4], taps0[4] = { 0, 0, 0, 0 };
            size_t idx[4];
            for (int j = 0; j < 4; j++) {
                long long x = gx + j;
                ins[j] = x >= 0 && x < (long long)s->w;
                idx[j] = (size_t)y * s->w + (size_t)(ins[j] ? x : 0);
                int32_t xrel = (int32_t)(uint32_t)(x - (long long)xMin);
                int32_t w1 = WADD(WADD(w1R, WMUL(WSUB(y, yMin), w1Y)), WMUL(xrel, w1X));
                int32_t w2 = WADD(WADD(w2R, WMUL(WSUB(y, yMin), w2Y)), WMUL(xrel, w2X));
                int32_t w3 = WADD(WADD(w3R, WMUL(WSUB(y, yMin), w3Y)), WMUL(xrel, w3X));
                cov[j] = ((w1 | w2 | w3) > 0) && x < (long long)xMax && ins[j];
                float W1 = (float)w1 * invSum, W2 = (float)w2 * invSum, W3 = (float)w3 * invSum;
                z[j] = rcp_x86((v1->zinv * W1 + v2->zinv * W2) + v3->zinv * W3);
                float zb = ins[j] ? s->depth[idx[j]] : 0.0f;
                m[j] = cov[j] && (!depth_on || depth_pass(st->depth_func, z[j], zb));
                uint32_t f = smooth ? color_smooth(v1->rgba, v2->rgba, v3->rgba, W1, W2, W3)
                                    : color_flat(v1->rgba, v2->rgba, v3->rgba, W1, W2, W3);
                if (tex_on) {
                    float u = (v1->u * W1 + v2->u * W2) + v3->u * W3;
                    float v = (v1->v * W1 + v2->v * W2) + v3->v * W3;
                    if (t->is3d) { u = u * z[j]; v = v * z[j]; }
                    if (!m[j]) { u = 0.0f; v = 0.0f; }                      /* triangles.c:510 */
                    uint32_t taps[4] = { 0, 0, 0, 0 }; float fx, fy;
                    tex_taps(st, u, v, taps, &fx, &fy);
                    if (j == 0) memcpy(taps0, taps, sizeof taps0);
                    f = mul_color(tex_combine(st, tex_leader ? taps0 : taps, fx, fy), f);
                }
                if (phong_on) {
                    float N[3] = { (v1->nx * W1 + v2->nx * W2) + v3->nx * W3, (v1->ny * W1 + v2->ny * W2) + v3->ny * W3, (v1->nz * W1 + v2->nz * W2) + v3->nz * W3 };
                    float P[3] = { (v1->px * W1 + v2->px * W2) + v3->px * W3, (v1->py * W1 + v2->py * W2) + v3->py * W3, (v1->pz * W1 + v2->pz * W2) + v3->pz * W3 };
                    f = phong(f, st, face, P, N);
                }
                dst[j] = ins[j] ? s->color[idx[j]] : 0u;
                if (blend_on) f = blend(st->blend_mode, f, fb_leader ? dst[0] : dst[j]);
                frag[j] = f;
            }
            for (int j = 0; j < 4; j++) {
                if (m[j]) {
                    s->color[idx[j]] = (fb_leader ? frag[0] : frag[j]) | alpha_or;
                    s->depth[idx[j]] = z[j];
                    g_cnt.pixels_shaded++;
                } else if (cov[j]) g_cnt.pixels_depth_failed++;
            }
        }
    }
}

/* ---- C-ABI ----------------------------------------------------------------------------------- */

int  pfcu_init(int device) { (void)device; return PFCU_OK; }
void pfcu_shutdown(void) {}
const char *pfcu_last_error(void) { return g_err; }
const char *pfcu_backend_name(void) { return "oracle-c"; }
int  pfcu_set_stream(void *s) { (void)s; return PFCU_OK; }
void *pfcu_get_stream(void) { return NULL; }
void *pfcu_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void  pfcu_host_free(void *p) { free(p); }
int   pfcu_host_wait(const void *p) { (void)p; return PFCU_OK; }
int   pfcu_host_set_static(void *p, int on) { (void)p; (void)on; return PFCU_OK; }      /* host memory is read at every draw here */
int   pfcu_host_modified(void *p) { (void)p; return PFCU_OK; }
int   pfcu_host_is_static(const void *p) { (void)p; return 0; }
int   pfcu_host_register(void *p, size_t bytes) { (void)p; (void)bytes; return PFCU_OK; }
void  pfcu_host_unregister(void *p) { (void)p; }
int  pfcu_set_approx_tables(const uint32_t *rcp, int rb, const uint32_t *rs, int sb)
{ (void)rcp; (void)rb; (void)rs; (void)sb; return PFCU_OK; }   /* native RCPSS/RSQRTSS are used */

pfcu_surface *pfcu_surface_create(uint32_t w, uint32_t h) { return pfcu_surface_create_format(w, h, PFCU_TEX_RGBA8); }
int pfcu_surface_format(const pfcu_surface *s) { return s ? s->fmt : -1; }
pfcu_surface *pfcu_surface_create_format(uint32_t w, uint32_t h, int fmt)
{
    if (fmt < PFCU_TEX_RGBA8 || fmt > PFCU_TEX_BGR8) return NULL;
    pfcu_surface *s = (pfcu_surface *)calloc(1, sizeof *s);
    if (!s) return NULL;
    s->w = w; s->h = h; s->owned = 1; s->fmt = fmt;
    s->color = (uint32_t *)calloc((size_t)w * h + 16, 4);
    s->depth = (float *)calloc((size_t)w * h + 16, 4);
    if (!s->color || !s->depth) { free(s->color); free(s->depth); free(s); return NULL; }
    return s;
}
pfcu_surface *pfcu_surface_wrap(void *c, void *d, uint32_t w, uint32_t h)
{
    pfcu_surface *s = (pfcu_surface *)calloc(1, sizeof *s);
    if (!s) return NULL;
    s->w = w; s->h = h; s->color = (uint32_t *)c; s->depth = (float *)d;
    return s;
}
void pfcu_surface_destroy(pfcu_surface *s) { if (s) { if (s->owned) { free(s->color); free(s->depth); } free(s); } }
uint32_t pfcu_surface_width(const pfcu_surface *s) { return s->w; }
uint32_t pfcu_surface_height(const pfcu_surface *s) { return s->h; }
void *pfcu_surface_color_ptr(const pfcu_surface *s) { return s->color; }
void *pfcu_surface_depth_ptr(const pfcu_surface *s) { return s->depth; }

/* the scalar getters / setters of the reference for the four 8-bit layouts (pixel.h:233-360, 576-710), to / from
   canonical RGBA8 (alpha of a 3-byte pixel reads as 255) */
static uint32_t native_get(const void *px, size_t i, int fmt)
{
    const uint8_t *p = (const uint8_t *)px;
    switch (fmt) {
    case PFCU_TEX_BGRA8: return (uint32_t)p[4 * i + 2] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i] << 16) | ((uint32_t)p[4 * i + 3] << 24);
    case PFCU_TEX_RGB8:  return (uint32_t)p[3 * i] | ((uint32_t)p[3 * i + 1] << 8) | ((uint32_t)p[3 * i + 2] << 16) | 0xff000000u;
    case PFCU_TEX_BGR8:  return (uint32_t)p[3 * i + 2] | ((uint32_t)p[3 * i + 1] << 8) | ((uint32_t)p[3 * i] << 16) | 0xff000000u;
    default: { uint32_t v; memcpy(&v, p + 4 * i, 4); return v; }
    }
}
static void native_set(void *px, size_t i, int fmt, uint32_t c)
{
    uint8_t *p = (uint8_t *)px;
    switch (fmt) {
    case PFCU_TEX_BGRA8: p[4 * i] = (uint8_t)(c >> 16); p[4 * i + 1] = (uint8_t)(c >> 8); p[4 * i + 2] = (uint8_t)c; p[4 * i + 3] = (uint8_t)(c >> 24); break;
    case PFCU_TEX_RGB8:  p[3 * i] = (uint8_t)c; p[3 * i + 1] = (uint8_t)(c >> 8); p[3 * i + 2] = (uint8_t)(c >> 16); break;
    case PFCU_TEX_BGR8:  p[3 * i] = (uint8_t)(c >> 16); p[3 * i + 1] = (uint8_t)(c >> 8); p[3 * i + 2] = (uint8_t)c; break;
    default: memcpy(p + 4 * i, &c, 4); break;
    }
}

int pfcu_surface_upload(pfcu_surface *s, const void *c, const float *d, uint32_t y0, uint32_t rows)
{
    if (y0 > s->h || rows > s->h - y0) return PFCU_ERR_INVALID;
    size_t off = (size_t)y0 * s->w, n = (size_t)rows * s->w;
    if (c && s->fmt == PFCU_TEX_RGBA8) memcpy(s->color + off, (const uint32_t *)c + off, n * 4);
    else if (c) for (size_t i = off; i < off + n; i++) s->color[i] = native_get(c, i, s->fmt);
    if (d) memcpy(s->depth + off, d + off, n * 4);
    return PFCU_OK;
}
int pfcu_surface_download(pfcu_surface *s, void *c, float *d, uint32_t y0, uint32_t rows)
{
    if (y0 > s->h || rows > s->h - y0) return PFCU_ERR_INVALID;
    size_t off = (size_t)y0 * s->w, n = (size_t)rows * s->w;
    if (c && s->fmt == PFCU_TEX_RGBA8) memcpy((uint32_t *)c + off, s->color + off, n * 4);
    else if (c) for (size_t i = off; i < off + n; i++) native_set(c, i, s->fmt, s->color[i]);
    if (d) memcpy(d + off, s->depth + off, n * 4);
    return PFCU_OK;
}
int pfcu_surface_fill(pfcu_surface *s, int dc, uint32_t rgba, int dd, float depth)
{
    size_t n = (size_t)s->w * s->h;
    if (s->fmt >= PFCU_TEX_RGB8) rgba |= 0xff000000u;
    for (size_t i = 0; i < n; i++) { if (dc) s->color[i] = rgba; if (dd) s->depth[i] = depth; }
    return PFCU_OK;
}
/* pfClear SIMD path, context.c:696-747 */
int pfcu_surface_clear_ref(pfcu_surface *s, int dc, uint32_t rgba, int dd, float depth)
{
    uint32_t size = s->w * s->h, aligned = size - (size % 8u);
    if (s->fmt >= PFCU_TEX_RGB8) rgba |= 0xff000000u;
    for (uint32_t i = 8; i < aligned; i++) { if (dc) s->color[i] = rgba; if (dd) s->depth[i] = depth; }
    for (uint32_t i = aligned; i < size; i++) { if (dc) s->color[i] = s->color[0]; if (dd) s->depth[i] = s->depth[0]; }
    return PFCU_OK;
}
int pfcu_surface_set_tile_owner(pfcu_surface *s, uint32_t rank, uint32_t world) { s->rank = rank; s->world = world; return PFCU_OK; }

static uint32_t owned_tiles(const pfcu_surface *s, uint32_t rank, uint32_t world)
{
    uint32_t nt = ((s->w + 63u) / 64u) * ((s->h + 63u) / 64u);
    if (world <= 1) return nt;
    return nt / world + ((nt % world) > rank ? 1u : 0u);
}
size_t pfcu_surface_owned_bytes(const pfcu_surface *s, uint32_t rank, uint32_t world, int with_depth)
{ return (size_t)owned_tiles(s, rank, world) * 64u * 64u * 4u * (with_depth ? 2u : 1u); }

static int pack_unpack(pfcu_surface *s, uint32_t rank, uint32_t world, int with_depth, void *staging, int unpack)
{
    uint32_t tx = (s->w + 63u) / 64u, ty = (s->h + 63u) / 64u, k = 0;
    if (world == 0) world = 1;
    uint32_t *st = (uint32_t *)staging;
    for (uint32_t t = 0; t < tx * ty; t++) {
        if (t % world != rank) continue;
        uint32_t bx = (t % tx) * 64u, by = (t / tx) * 64u;
        uint32_t *dstc = st + (size_t)k * 4096u * (with_depth ? 2u : 1u);
        uint32_t *dstd = dstc + 4096u;
        for (uint32_t y = 0; y < 64; y++) for (uint32_t x = 0; x < 64; x++) {
            if (bx + x >= s->w || by + y >= s->h) continue;
            size_t gi = (size_t)(by + y) * s->w + bx + x;
            if (unpack) { s->color[gi] = dstc[y * 64 + x]; if (with_depth) memcpy(&s->depth[gi], &dstd[y * 64 + x], 4); }
            else { dstc[y * 64 + x] = s->color[gi]; if (with_depth) memcpy(&dstd[y * 64 + x], &s->depth[gi], 4); }
        }
        k++;
    }
    return PFCU_OK;
}
int pfcu_surface_pack_tiles(pfcu_surface *s, uint32_t r, uint32_t w, int wd, void *st) { return pack_unpack(s, r, w, wd, st, 0); }
int pfcu_surface_unpack_tiles(pfcu_surface *s, uint32_t r, uint32_t w, int wd, const void *st) { return pack_unpack(s, r, w, wd, (void *)st, 1); }

static size_t tex_bytes(uint32_t w, uint32_t h, int fmt)
{
    if (fmt >= PFCU_TEX_PIX) return (size_t)w * h * (size_t)pfx_bytes(fmt - PFCU_TEX_PIX);
    return (size_t)w * h * ((fmt == PFCU_TEX_RGBA8 || fmt == PFCU_TEX_BGRA8) ? 4u : 3u);
}

pfcu_texture *pfcu_texture_create(const void *px, uint32_t w, uint32_t h, int fmt)
{
    pfcu_texture *t = (pfcu_texture *)calloc(1, sizeof *t);
    if (!t) return NULL;
    t->w = w; t->h = h; t->fmt = fmt; t->owned = 1; t->leader = (fmt == PFCU_TEX_BGRA8);
    t->pixels = (uint8_t *)malloc(tex_bytes(w, h, fmt) + 16);
    if (!t->pixels) { free(t); return NULL; }
    memset(t->pixels, 0, tex_bytes(w, h, fmt) + 16);
    if (px) memcpy(t->pixels, px, tex_bytes(w, h, fmt));
    return t;
}
pfcu_texture *pfcu_texture_from_surface(pfcu_surface *s)
{
    pfcu_texture *t = (pfcu_texture *)calloc(1, sizeof *t);
    if (!t) return NULL;
    t->w = s->w; t->h = s->h; t->fmt = PFCU_TEX_RGBA8; t->alias = s; t->leader = (s->fmt == PFCU_TEX_BGRA8);
    return t;
}
int pfcu_texture_update(pfcu_texture *t, const void *px)
{
    if (!t || t->alias || !px) return PFCU_ERR_INVALID;
    memcpy(t->pixels, px, tex_bytes(t->w, t->h, t->fmt));
    return PFCU_OK;
}
void pfcu_texture_destroy(pfcu_texture *t) { if (t) { if (t->owned) free(t->pixels); free(t); } }

int pfcu_submit(pfcu_surface *s, const pfcu_state *states, uint32_t n_states, const pfcu_triangle *tris, uint32_t n_tris)
{
    for (uint32_t i = 0; i < n_tris; i++) {
        if (tris[i].state >= n_states) { snprintf(g_err, sizeof g_err, "state index out of range"); return PFCU_ERR_INVALID; }
        const pfcu_state *st = &states[tris[i].state];
        if (s->fmt != PFCU_TEX_RGBA8 || ((st->flags & PFCU_ST_TEXTURE) && st->texture && st->texture->leader)) raster_triangle_rows(s, st, &tris[i]);
        else raster_triangle(s, st, &tris[i]);
    }
    return PFCU_OK;
}
pfcu_batch *pfcu_batch_upload(const pfcu_state *states, uint32_t n_states, const pfcu_triangle *tris, uint32_t n_tris)
{
    pfcu_batch *b = (pfcu_batch *)calloc(1, sizeof *b);
    if (!b) return NULL;
    b->states = (pfcu_state *)malloc(sizeof(pfcu_state) * (n_states ? n_states : 1));
    b->tris = (pfcu_triangle *)malloc(sizeof(pfcu_triangle) * (n_tris ? n_tris : 1));
    memcpy(b->states, states, sizeof(pfcu_state) * n_states);
    memcpy(b->tris, tris, sizeof(pfcu_triangle) * n_tris);
    b->n_states = n_states; b->n_tris = n_tris;
    return b;
}
int  pfcu_batch_submit(pfcu_surface *s, pfcu_batch *b) { return pfcu_submit(s, b->states, b->n_states, b->tris, b->n_tris); }
void pfcu_batch_destroy(pfcu_batch *b) { if (b) { free(b->states); free(b->tris); free(b); } }
int pfcu_surface_download_async(pfcu_surface *s, void *c, float *d, uint32_t y0, uint32_t rows) { return pfcu_surface_download(s, c, d, y0, rows); }
int pfcu_surface_wait(pfcu_surface *s) { (void)s; return PFCU_OK; }
/* present over peer memory: CUDA only (the gloo tests use pack / gather / unpack) */
int pfcu_surface_ipc_handles(pfcu_surface *s, void *c, void *d) { (void)s; (void)c; (void)d; return PFCU_ERR_INVALID; }
int pfcu_surface_set_present_peer(pfcu_surface *s, const void *c, const void *d) { (void)s; (void)c; (void)d; return PFCU_ERR_INVALID; }
int pfcu_surface_set_present_surface(pfcu_surface *s, pfcu_surface *t) { (void)s; (void)t; return PFCU_ERR_INVALID; }
int pfcu_surface_clear_present(pfcu_surface *s) { (void)s; return PFCU_OK; }
int pfcu_surface_push_tiles(pfcu_surface *s, uint32_t r, uint32_t w, int d) { (void)s; (void)r; (void)w; (void)d; return PFCU_ERR_INVALID; }
int pfcu_submit_raw(pfcu_surface *s, const pfcu_state *states, uint32_t n_states, const pfcu_vparams_lit *vparams, uint32_t n_vparams,
                    const float *pow_tables, uint32_t n_pow_tables, const pfcu_rawtri *tris, uint32_t n_tris, uint32_t *n_out)
{   /* never advertised (pfcu_capabilities): with this library the front end runs the vertex stage itself */
    (void)s; (void)states; (void)n_states; (void)vparams; (void)n_vparams; (void)pow_tables; (void)n_pow_tables; (void)tris; (void)n_tris;
    if (n_out) *n_out = 0;
    return PFCU_ERR_INVALID;
}
/* device-resident render lists: never advertised (pfcu_capabilities), the front end replays lists on the host here */
pfcu_list *pfcu_list_create(const pfcu_rawtri *tris, uint32_t n) { (void)tris; (void)n; return NULL; }
void pfcu_list_destroy(pfcu_list *l) { (void)l; }
uint32_t pfcu_list_size(const pfcu_list *l) { (void)l; return 0; }
int pfcu_list_job_supported(const pfcu_surface *s, uint32_t n, uint32_t k) { (void)s; (void)n; (void)k; return 0; }
int pfcu_submit_list_jobs(const pfcu_list_job *jobs, uint32_t n) { (void)jobs; (void)n; return PFCU_ERR_INVALID; }
/* points and lines: the reference's scalar loops (lines.c:283-530, points.c:85-183), one primitive after the other */
#include "../pixelforge_b200/csrc/pf_prims.h"
static void prim_pixel(pfcu_surface *s, const pfcu_prim *p, uint32_t off, float z, uint32_t color, int test)
{
    if (off >= s->w * s->h) return;                     /* the reference would write outside its buffer */
    const uint32_t x = off % s->w, y = off / s->w;
    if (s->world > 1 && ((x / 64u) + (y / 64u) * ((s->w + 63u) / 64u)) % s->world != s->rank) return;
    if (test && !pfp_depth(p->depth_func, z, s->depth[off])) return;
    s->color[off] = ((p->flags & PFCU_ST_BLEND) ? pfp_blend(p->blend_mode, color, s->color[off]) : color) | (s->fmt >= PFCU_TEX_RGB8 ? 0xff000000u : 0u);
    s->depth[off] = z;
}
int pfcu_submit_prims(pfcu_surface *s, const pfcu_prim *prims, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) {
        const pfcu_prim *p = &prims[i];
        const int ztest = (p->flags & PFCU_ST_DEPTH_TEST) != 0;
        if (p->kind == PFP_KIND_POINT) {
            const int cx = (int)p->x1, cy = (int)p->y1;
            if (p->size <= 1.0f) { prim_pixel(s, p, (uint32_t)cy * s->w + (uint32_t)cx, p->z1, p->c1, ztest); continue; }
            const float r = p->size * 0.5f, r2 = r * r;
            const int R = (int)r;
            for (int y = -R; y <= R; y++)
                for (int x = -R; x <= R; x++)
                    if ((float)(y * y + x * x) <= r2) {
                        const uint32_t px = (uint32_t)(cx + x), py = (uint32_t)(cy + y);
                        if (px < s->w && py < s->h) prim_pixel(s, p, py * s->w + px, p->z1, p->c1, ztest);
                    }
            continue;
        }
        int axis;
        const unsigned nsub = pfp_thick_count(p->x1, p->y1, p->x2, p->y2, p->size, &axis);
        const int thick = p->size > 1.5f;
        for (unsigned sub = 0; sub < nsub; sub++) {
            const float sh = pfp_thick_shift(sub);
            pfp_line L;
            pfp_line_setup(&L, axis ? p->x1 : p->x1 + sh, axis ? p->y1 + sh : p->y1, axis ? p->x2 : p->x2 + sh, axis ? p->y2 + sh : p->y2);
            const int test = ztest || (thick && sub == 0);
            const unsigned steps = pfp_line_steps(&L);
            for (unsigned k = 0; k < steps; k++) {
                float t;
                const uint32_t off = pfp_line_step(&L, k, s->w, &t);
                prim_pixel(s, p, off, p->z1 + t * (p->z2 - p->z1), pfp_color_lerp(p->c1, p->c2, t), test);
            }
        }
    }
    return PFCU_OK;
}
/* ---- full-surface operations: the reference's own loops, one pixel after the other (context.c:1938-1977, 1988-2084,
 * 2275-2344, 2349-2395).  Scalar getters / setters = canonical RGBA8 with alpha 255 on 3-byte targets; scalar blend and
 * depth tables = pfp_blend / pfp_depth (pinned by the points / lines cases). */
int pfcu_surface_rect(pfcu_surface *s, int32_t x1, int32_t y1, int32_t x2, int32_t y2, uint32_t rgba)
{
    const uint32_t keep = s->fmt >= PFCU_TEX_RGB8 ? 0xff000000u : 0u;
    for (int32_t y = y1; y <= y2; y++)
        for (int32_t x = x1; x <= x2; x++) {
            const uint32_t o = (uint32_t)y * s->w + (uint32_t)x;           /* tex->setter(pixels, y*tex->w + x, color) */
            if (o < s->w * s->h) s->color[o] = rgba | keep;
        }
    return PFCU_OK;
}
int pfcu_surface_fog(pfcu_surface *s, const pfcu_fog *f)
{
    const uint32_t keep = s->fmt >= PFCU_TEX_RGB8 ? 0xff000000u : 0u, alpha = f->rgba >> 24, rgb = f->rgba & 0x00ffffffu;
    const size_t n = (size_t)s->w * s->h;
    for (size_t i = 0; i < n; i++) {
        const float depth = s->depth[i];
        if (depth >= f->end) {
            s->color[i] = (alpha == 255u ? f->rgba : pfp_blend(1, f->rgba, s->color[i])) | keep;
        } else if (depth > f->start) {
            float t = 0;
            switch (f->mode) {
            case 0: t = (depth - f->start) * f->inv_len; break;
            case 1: t = 1.0f - expf(-f->density * (depth - f->start)); break;      /* the host libm, as the reference */
            case 2: t = 1.0f - exp2f(-f->density * (depth - f->start)); break;
            }
            const uint32_t a = (uint8_t)(t * (float)alpha);
            s->color[i] = pfp_blend(1, rgb | (a << 24), s->color[i]) | keep;
        }
    }
    return PFCU_OK;
}
int pfcu_surface_draw_pixels(pfcu_surface *s, const pfcu_pixels *d)
{
    const uint32_t keep = s->fmt >= PFCU_TEX_RGB8 ? 0xff000000u : 0u;
    const uint32_t wm1 = d->width - 1u, hm1 = d->height - 1u;
    for (int32_t y = d->ymin; y <= d->ymax; y++) {
        const float v = (float)(y - d->ys) * d->inv_ylen;
        const uint32_t ysrc = (uint32_t)(v * hm1) * d->width;
        for (int32_t x = d->xmin; x <= d->xmax; x++) {
            const uint32_t o = (uint32_t)y * s->w + (uint32_t)x;
            if (o >= s->w * s->h) continue;
            if (!(d->flags & PFCU_ST_DEPTH_TEST) || pfp_depth(d->depth_func, d->z, s->depth[o])) {
                const float u = (float)(x - d->xs) * d->inv_xlen;
                const uint32_t so = ysrc + (uint32_t)(u * wm1);
                const uint32_t c = so < d->width * d->height ? pfx_get(d->pixels, so, d->format) : 0u;
                s->depth[o] = d->z;
                s->color[o] = ((d->flags & PFCU_ST_BLEND) ? pfp_blend(d->blend_mode, c, s->color[o]) : c) | keep;
            }
        }
    }
    return PFCU_OK;
}
int pfcu_surface_read_pixels(pfcu_surface *s, uint32_t x0, uint32_t y0, uint32_t cols, uint32_t rows, uint32_t dst_width, int format, void *host_pixels)
{
    if (cols == 0 || rows == 0) return PFCU_OK;
    if (x0 >= s->w || y0 >= s->h || cols > s->w - x0 || rows > s->h - y0 || cols > dst_width) return PFCU_ERR_INVALID;
    for (uint32_t y = 0; y < rows; y++)
        for (uint32_t x = 0; x < cols; x++)
            pfx_set(host_pixels, (size_t)y * dst_width + x, format, s->color[(size_t)(y0 + y) * s->w + x0 + x]);
    return PFCU_OK;
}
unsigned pfcu_capabilities(void) { return 0u; }     /* the oracle has no device vertex stage: the front end keeps it on the host */
int pfcu_draw_triangles(pfcu_surface *s, const pfcu_state *st, const pfcu_vparams *vp, const pfcu_draw *d, uint32_t *n)
{ (void)s; (void)st; (void)vp; (void)d; (void)n; return PFCU_ERR_INVALID; }
void pfcu_profile_enable(int on) { (void)on; }
void pfcu_set_raster_path(int path) { (void)path; }    /* one scalar path here */
int  pfcu_profile_read(pfcu_profile *out) { memset(out, 0, sizeof *out); return PFCU_OK; }
int  pfcu_fence(void) { return PFCU_OK; }
int  pfcu_finish(void) { return PFCU_OK; }
int  pfcu_get_counters(pfcu_counters *out) { *out = g_cnt; return PFCU_OK; }
void pfcu_reset_counters(void) { memset(&g_cnt, 0, sizeof g_cnt); }
