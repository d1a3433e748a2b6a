/*
 * pf_pixfmt.h - the reference's SCALAR pixel getters and setters for every supported (format, type) pair, written once
 * and compiled as CUDA device code (k_draw_pixels / k_read_pixels in pfcu_surface.cuh) and as C99 (the oracle, the front
 * end's validity check).  Restates, in behaviour, src/internal/pixel.h:76-406 (setters), :408-760 (getters), the tables
 * GC_pixelGetters / GC_pixelSetters (:764-860) and pfmFloatToHalf / pfmHalfToFloat (src/pfm.h:107-161).
 * pfReadPixels and pfDrawPixels take any of the 38 pairs (context.c:1998-2005, 2351-2362); framebuffers of the triangle
 * path exist in the four 8-bit layouts only.  Textures exist in every pair: pfx_tex_get at the end of this
 * file restates the SIMD texel getters (pixel.h:2249-3040), which are NOT the scalar ones (DESIGN Q22).
 * Arithmetic notes kept from upstream: single-channel HALF / FLOAT setters DIVIDE by 255.0f, the multi-channel ones
 * multiply by (float)(1.0/255); luminance = r*k*0.299f + g*k*0.587f + b*k*0.114f in float, left to right;
 * (PFubyte)(float) is CVTTSS2SI followed by a truncation to 8 bits; the 5-5-5-1 alpha threshold compares in double.
 */
#ifndef PF_PIXFMT_H
#define PF_PIXFMT_H

#include "pf_vstage.h"

/* code of a (PFpixelformat, PFdatatype) pair as the C-ABI carries it */
#ifndef PFCU_PIX
#define PFCU_PIX(format, type) ((int)(format) * 16 + (int)(type))
#endif
#ifndef PFCU_TEX_PIX
#define PFCU_TEX_PIX 256     /* include/pfcu.h */
#endif
enum { PFX_RED = 0, PFX_GREEN, PFX_BLUE, PFX_ALPHA, PFX_LUM, PFX_LUMA, PFX_RGB, PFX_RGBA, PFX_BGR, PFX_BGRA };       /* PFpixelformat */
enum { PFX_UBYTE = 0, PFX_565 = 2, PFX_5551 = 3, PFX_4444 = 4, PFX_HALF = 9, PFX_FLOAT = 10 };                      /* PFdatatype    */

#ifdef __CUDACC__
#  define PFX_HD __host__ __device__ __forceinline__
#else
#  define PFX_HD static inline
#endif
/* bytes per pixel of a pair, 0 when the reference has no getter / setter for it */
PFX_HD int pfx_bytes(int code)
{
    const int f = code >> 4, t = code & 15;
    if (f < 0 || f > PFX_BGRA) return 0;
    const int comps = f <= PFX_LUM ? 1 : (f == PFX_LUMA ? 2 : ((f == PFX_RGB || f == PFX_BGR) ? 3 : 4));
    if (t == PFX_UBYTE) return comps;
    if (t == PFX_HALF) return comps * 2;
    if (t == PFX_FLOAT) return comps * 4;
    if (t == PFX_565) return comps == 3 ? 2 : 0;
    if (t == PFX_5551 || t == PFX_4444) return comps == 4 ? 2 : 0;
    return 0;
}

PFV_FN uint32_t pfx_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
PFV_FN float pfx_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

PFV_FN uint16_t pfx_float_to_half(float x)         /* pfmFloatToHalfI */
{
    const uint32_t ui = pfx_f2u(x);
    const int s = (int)((ui >> 16) & 0x8000u);
    const int em = (int)(ui & 0x7fffffffu);
    int h = (em - (112 << 23) + (1 << 12)) >> 13;
    h = (em < (113 << 23)) ? 0 : h;
    h = (em >= (143 << 23)) ? 0x7c00 : h;
    h = (em > (255 << 23)) ? 0x7e00 : h;
    return (uint16_t)(s | h);
}

PFV_FN float pfx_half_to_float(uint16_t h)          /* pfmHalfToFloatI */
{
    const uint32_t s = (uint32_t)(h & 0x8000u) << 16;
    const int em = h & 0x7fff;
    int r = (int)((uint32_t)(em + (112 << 10)) << 13);
    r = (em < (1 << 10)) ? 0 : r;
    r = (int)((uint32_t)r + ((em >= (31 << 10)) ? (uint32_t)(112 << 23) : 0u));
    return pfx_u2f(s | (uint32_t)r);
}

PFV_FN uint32_t pfx_ub(float v) { return (uint32_t)PFV_F2I(v) & 255u; }         /* (PFubyte)(float) */
PFV_FN uint32_t pfx_pack(uint32_t r, uint32_t g, uint32_t b, uint32_t a) { return r | (g << 8) | (b << 16) | (a << 24); }

/* one component stored as UBYTE / HALF / FLOAT at element index e, as (PFubyte)(255 * x) - order of the product as upstream */
PFV_FN uint32_t pfx_comp_255x(const void *px, size_t e, int t)       /* (PFubyte)(255 * value) */
{
    if (t == PFX_UBYTE) return ((const uint8_t *)px)[e];
    if (t == PFX_HALF) return pfx_ub(PFV_MUL(255.0f, pfx_half_to_float(((const uint16_t *)px)[e])));
    return pfx_ub(PFV_MUL(255.0f, ((const float *)px)[e]));
}

PFV_FN uint32_t pfx_get(const void *px, size_t i, int code)
{
    const int f = code >> 4, t = code & 15;
    const uint8_t *p8 = (const uint8_t *)px; const uint16_t *p16 = (const uint16_t *)px;
    switch (f) {
    case PFX_RED:   return pfx_pack(pfx_comp_255x(px, i, t), 0, 0, 255);
    case PFX_GREEN: return pfx_pack(0, pfx_comp_255x(px, i, t), 0, 255);
    case PFX_BLUE:  return pfx_pack(0, 0, pfx_comp_255x(px, i, t), 255);
    case PFX_ALPHA: return pfx_pack(255, 255, 255, pfx_comp_255x(px, i, t));
    case PFX_LUM:   { const uint32_t g = pfx_comp_255x(px, i, t); return pfx_pack(g, g, g, 255); }
    case PFX_LUMA:  { const uint32_t g = pfx_comp_255x(px, 2 * i, t), a = pfx_comp_255x(px, 2 * i + 1, t); return pfx_pack(g, g, g, a); }
    case PFX_RGB: case PFX_BGR: {
        uint32_t c0, c1, c2;
        if (t == PFX_565) {
            const uint32_t v = p16[i];
            c0 = pfx_ub(PFV_MUL(PFV_U2F((v & 0xF800u) >> 11), 255.0f / 31)); c1 = pfx_ub(PFV_MUL(PFV_U2F((v & 0x7E0u) >> 5), 255.0f / 63));
            c2 = pfx_ub(PFV_MUL(PFV_U2F(v & 0x1Fu), 255.0f / 31));
        } else if (t == PFX_UBYTE) { c0 = p8[3 * i]; c1 = p8[3 * i + 1]; c2 = p8[3 * i + 2]; }
        else {
            const float a = t == PFX_HALF ? pfx_half_to_float(p16[3 * i]) : ((const float *)px)[3 * i];
            const float b = t == PFX_HALF ? pfx_half_to_float(p16[3 * i + 1]) : ((const float *)px)[3 * i + 1];
            const float c = t == PFX_HALF ? pfx_half_to_float(p16[3 * i + 2]) : ((const float *)px)[3 * i + 2];
            c0 = pfx_ub(PFV_MUL(a, 255.0f)); c1 = pfx_ub(PFV_MUL(b, 255.0f)); c2 = pfx_ub(PFV_MUL(c, 255.0f));
        }
        return f == PFX_RGB ? pfx_pack(c0, c1, c2, 255) : pfx_pack(c2, c1, c0, 255);
    }
    default: {      /* RGBA, BGRA */
        uint32_t c0, c1, c2, c3;
        if (t == PFX_5551) {
            const uint32_t v = p16[i];
            c0 = pfx_ub(PFV_MUL(PFV_U2F((v & 0xF800u) >> 11), 255.0f / 31)); c1 = pfx_ub(PFV_MUL(PFV_U2F((v & 0x7C0u) >> 6), 255.0f / 31));
            c2 = pfx_ub(PFV_MUL(PFV_U2F((v & 0x3Eu) >> 1), 255.0f / 31)); c3 = (v & 1u) * 255u;
        } else if (t == PFX_4444) {
            const uint32_t v = p16[i];
            c0 = pfx_ub(PFV_MUL(PFV_U2F((v & 0xF000u) >> 12), 255.0f / 15)); c1 = pfx_ub(PFV_MUL(PFV_U2F((v & 0xF00u) >> 8), 255.0f / 15));
            c2 = pfx_ub(PFV_MUL(PFV_U2F((v & 0xF0u) >> 4), 255.0f / 15)); c3 = pfx_ub(PFV_MUL(PFV_U2F(v & 0xFu), 255.0f / 15));
        } else if (t == PFX_UBYTE) { c0 = p8[4 * i]; c1 = p8[4 * i + 1]; c2 = p8[4 * i + 2]; c3 = p8[4 * i + 3]; }
        else {
            float v[4];
            for (int k = 0; k < 4; k++) v[k] = t == PFX_HALF ? pfx_half_to_float(p16[4 * i + k]) : ((const float *)px)[4 * i + k];
            c0 = pfx_ub(PFV_MUL(v[0], 255.0f)); c1 = pfx_ub(PFV_MUL(v[1], 255.0f)); c2 = pfx_ub(PFV_MUL(v[2], 255.0f)); c3 = pfx_ub(PFV_MUL(v[3], 255.0f));
        }
        /* the packed BGRA layouts name their fields in the same bit positions: first field = blue */
        return f == PFX_RGBA ? pfx_pack(c0, c1, c2, c3) : pfx_pack(c2, c1, c0, c3);
    }
    }
}

#ifdef __CUDACC__
#  define PFX_ROUND(x) roundf(x)
#else
#  define PFX_ROUND(x) roundf(x)
#endif
#define PFX_K255 ((float)(1.0 / 255))

PFV_FN float pfx_grey(uint32_t c)      /* PF_COLOR_GARYSCALE */
{
    const float r = PFV_MUL(PFV_MUL(PFV_U2F(c & 255u), PFX_K255), 0.299f), g = PFV_MUL(PFV_MUL(PFV_U2F((c >> 8) & 255u), PFX_K255), 0.587f);
    const float b = PFV_MUL(PFV_MUL(PFV_U2F((c >> 16) & 255u), PFX_K255), 0.114f);
    return PFV_ADD(PFV_ADD(r, g), b);
}

/* one component to element e: UBYTE stores the byte; HALF / FLOAT store `norm` (the caller's normalisation) */
PFV_FN void pfx_store_comp(void *px, size_t e, int t, uint32_t byte, float norm)
{
    if (t == PFX_UBYTE) ((uint8_t *)px)[e] = (uint8_t)byte;
    else if (t == PFX_HALF) ((uint16_t *)px)[e] = pfx_float_to_half(norm);
    else ((float *)px)[e] = norm;
}

PFV_FN void pfx_set(void *px, size_t i, int code, uint32_t c)
{
    const int f = code >> 4, t = code & 15;
    const uint32_t r = c & 255u, g = (c >> 8) & 255u, b = (c >> 16) & 255u, a = c >> 24;
    uint16_t *p16 = (uint16_t *)px;
    switch (f) {
    case PFX_RED:   pfx_store_comp(px, i, t, r, PFV_DIV(PFV_U2F(r), 255.0f)); return;
    case PFX_GREEN: pfx_store_comp(px, i, t, g, PFV_DIV(PFV_U2F(g), 255.0f)); return;
    case PFX_BLUE:  pfx_store_comp(px, i, t, b, PFV_DIV(PFV_U2F(b), 255.0f)); return;
    case PFX_ALPHA: pfx_store_comp(px, i, t, a, PFV_DIV(PFV_U2F(a), 255.0f)); return;
    case PFX_LUM:   { const float y = pfx_grey(c); pfx_store_comp(px, i, t, pfx_ub(PFV_MUL(255.0f, y)), y); return; }
    case PFX_LUMA:  { const float y = pfx_grey(c);
                      pfx_store_comp(px, 2 * i, t, pfx_ub(PFV_MUL(255.0f, y)), y);
                      pfx_store_comp(px, 2 * i + 1, t, a, PFV_MUL(PFV_U2F(a), PFX_K255)); return; }
    case PFX_RGB: case PFX_BGR: {
        const uint32_t c0 = f == PFX_RGB ? r : b, c2 = f == PFX_RGB ? b : r;
        const float n0 = PFV_MUL(PFV_U2F(c0), PFX_K255), n1 = PFV_MUL(PFV_U2F(g), PFX_K255), n2 = PFV_MUL(PFV_U2F(c2), PFX_K255);
        if (t == PFX_565) {
            const uint32_t q0 = pfx_ub(PFX_ROUND(PFV_MUL(n0, 31.0f))), q1 = pfx_ub(PFX_ROUND(PFV_MUL(n1, 63.0f))), q2 = pfx_ub(PFX_ROUND(PFV_MUL(n2, 31.0f)));
            p16[i] = (uint16_t)((q0 << 11) | (q1 << 5) | q2);
        } else { pfx_store_comp(px, 3 * i, t, c0, n0); pfx_store_comp(px, 3 * i + 1, t, g, n1); pfx_store_comp(px, 3 * i + 2, t, c2, n2); }
        return;
    }
    default: {
        /* Q21: upstream's PF_COLOR_BGRA_NORMALIZE lists r, g, b, a (pixel.h:60-66), so the BGRA setters that go through it
           (5-5-5-1, 4-4-4-4, half, float) store RED in the field the BGRA getters read as blue; only the 8-bit BGRA
           setter swaps.  A read-back into such a layout followed by a draw from it exchanges red and blue. */
        const int swap = f == PFX_BGRA && t == PFX_UBYTE;
        const uint32_t c0 = swap ? b : r, c2 = swap ? r : b;
        const float n0 = PFV_MUL(PFV_U2F(c0), PFX_K255), n1 = PFV_MUL(PFV_U2F(g), PFX_K255), n2 = PFV_MUL(PFV_U2F(c2), PFX_K255), n3 = PFV_MUL(PFV_U2F(a), PFX_K255);
        if (t == PFX_5551) {
            const uint32_t q0 = pfx_ub(PFX_ROUND(PFV_MUL(n0, 31.0f))), q1 = pfx_ub(PFX_ROUND(PFV_MUL(n1, 31.0f))), q2 = pfx_ub(PFX_ROUND(PFV_MUL(n2, 31.0f)));
            const uint32_t qa = ((double)n3 > (double)(float)50 * (1.0 / 255)) ? 1u : 0u;       /* PF_RGBA_5_5_5_1_ALPHA_THRESHOLD = 50, config.h:45 */
            p16[i] = (uint16_t)((q0 << 11) | (q1 << 6) | (q2 << 1) | qa);
        } else if (t == PFX_4444) {
            const uint32_t q0 = pfx_ub(PFX_ROUND(PFV_MUL(n0, 15.0f))), q1 = pfx_ub(PFX_ROUND(PFV_MUL(n1, 15.0f))), q2 = pfx_ub(PFX_ROUND(PFV_MUL(n2, 15.0f))), q3 = pfx_ub(PFX_ROUND(PFV_MUL(n3, 15.0f)));
            p16[i] = (uint16_t)((q0 << 12) | (q1 << 8) | (q2 << 4) | q3);
        } else { pfx_store_comp(px, 4 * i, t, c0, n0); pfx_store_comp(px, 4 * i + 1, t, g, n1); pfx_store_comp(px, 4 * i + 2, t, c2, n2); pfx_store_comp(px, 4 * i + 3, t, a, n3); }
        return;
    }
    }
}

/* ---- texel fetch of the TRIANGLE path: the reference's SIMD getters (pixel.h:2249-3040), lane by lane --------------------
 * Not the scalar getters above: every SIMD getter gathers a 32-bit word at the texel's byte offset and differs from its
 * scalar twin in ways that define the pixels a textured triangle gets -
 *   8-bit single channels and luminance take the gathered word >> 24, i.e. the byte THREE texels further on; ALPHA takes the
 *     right byte and leaves red / green / blue 0 (the scalar getter says 255); LUMINANCE_ALPHA shuffles bytes 2, 2, 2, 3 of the
 *     word, i.e. reads the NEXT texel;
 *   5-6-5 expands by shifts (r5 << 3), 5-5-5-1 and 4-4-4-4 multiply by the integers 255/31 = 8 and 255/15 = 17;
 *   half and float components go through CVTPS2DQ (round to nearest even, 0x80000000 out of range or NaN) after the
 *     multiplication by 255 and are OR-ed together unmasked: a component above 1.0 spills into its neighbours; BGR half / float
 *     come back with blue in the red channel (the getter ORs its first element into the low byte); half = the low 16 bits
 *     of the gathered word through pfmHalfToFloat (the F16C path is compiled out: FLT16_MAX is not defined, simd.h:973-990).
 * The word is assembled from bytes (texel offsets of the 2- and 3-byte layouts are not word aligned); bytes past the end of
 * the texture read as zero (DESIGN Q18: the texture copies carry a zeroed tail). */
#ifdef __CUDACC__
PFV_FN uint32_t pfx_rne(float x) { const int r = __float2int_rn(x); return (fabsf(x) < 2147483648.0f) ? (uint32_t)r : 0x80000000u; }
#else
#  include <xmmintrin.h>
PFV_FN uint32_t pfx_rne(float x) { return (uint32_t)_mm_cvtss_si32(_mm_set_ss(x)); }
#endif
PFV_FN uint32_t pfx_raw32(const uint8_t *p, size_t off)
{
    return (uint32_t)p[off] | ((uint32_t)p[off + 1] << 8) | ((uint32_t)p[off + 2] << 16) | ((uint32_t)p[off + 3] << 24);
}
PFV_FN uint32_t pfx_c255(float x) { return pfx_rne(PFV_MUL(x, 255.0f)); }

PFV_FN uint32_t pfx_tex_get(const uint8_t *px, uint32_t idx, int code)
{
    const int f = code >> 4, t = code & 15;
    const uint32_t A = 0xFF000000u;
    if (t == PFX_UBYTE) {
        const uint32_t w = pfx_raw32(px, f == PFX_LUMA ? 2u * (size_t)idx : (size_t)idx);
        switch (f) {
        case PFX_RED:   return (w >> 24) | A;
        case PFX_GREEN: return ((w >> 24) << 8) | A;
        case PFX_BLUE:  return ((w >> 24) << 16) | A;
        case PFX_ALPHA: return w << 24;
        case PFX_LUM:   { const uint32_t g = w >> 24; return A | g | (g << 8) | (g << 16); }
        default:        { const uint32_t g = (w >> 16) & 255u; return g | (g << 8) | (g << 16) | (w & A); }     /* LUMINANCE_ALPHA */
        }
    }
    if (t == PFX_565) {
        const uint32_t w = pfx_raw32(px, 2u * (size_t)idx);
        const uint32_t hi = ((w & 0xF800u) >> 11) << 3, mid = ((w & 0x07E0u) >> 5) << 2, lo = (w & 0x001Fu) << 3;
        return f == PFX_RGB ? (A | (lo << 16) | (mid << 8) | hi) : (A | (hi << 16) | (mid << 8) | lo);
    }
    if (t == PFX_5551 || t == PFX_4444) {
        const uint32_t w = pfx_raw32(px, 2u * (size_t)idx);
        uint32_t c0, c1, c2, a8;
        if (t == PFX_5551) { c0 = ((w >> 11) & 0x1Fu) * 8u; c1 = ((w >> 6) & 0x1Fu) * 8u; c2 = ((w >> 1) & 0x1Fu) * 8u; a8 = (w & 1u) * 255u; }
        else               { c0 = ((w >> 12) & 0xFu) * 17u; c1 = ((w >> 8) & 0xFu) * 17u; c2 = ((w >> 4) & 0xFu) * 17u; a8 = (w & 0xFu) * 17u; }
        /* first field = red (RGBA) or blue (BGRA) */
        return f == PFX_RGBA ? ((a8 << 24) | (c2 << 16) | (c1 << 8) | c0) : ((a8 << 24) | (c0 << 16) | (c1 << 8) | c2);
    }
    /* half / float components */
    const uint32_t n = f <= PFX_LUM ? 1u : (f == PFX_LUMA ? 2u : ((f == PFX_RGB || f == PFX_BGR) ? 3u : 4u));
    uint32_t e[4] = { 0, 0, 0, 0 };
    for (uint32_t k = 0; k < n; k++) {
        const size_t el = (size_t)idx * n + k;
        const float v = t == PFX_HALF ? pfx_half_to_float((uint16_t)(pfx_raw32(px, 2u * el) & 0xFFFFu)) : pfx_u2f(pfx_raw32(px, 4u * el));
        e[k] = pfx_c255(v);
    }
    switch (f) {
    case PFX_RED:   return e[0] | A;
    case PFX_GREEN: return (e[0] << 8) | A;
    case PFX_BLUE:  return (e[0] << 16) | A;
    case PFX_ALPHA: return e[0] << 24;
    case PFX_LUM:   return A | (e[0] << 16) | (e[0] << 8) | e[0];
    case PFX_LUMA:  return (e[1] << 24) | (e[0] << 16) | (e[0] << 8) | e[0];
    case PFX_RGB: case PFX_BGR: return A | (e[2] << 16) | (e[1] << 8) | e[0];          /* BGR: blue lands in the red channel, as upstream */
    case PFX_RGBA:  return (e[3] << 24) | (e[2] << 16) | (e[1] << 8) | e[0];
    default:        return (e[3] << 24) | (e[0] << 16) | (e[1] << 8) | e[2];            /* BGRA */
    }
}

#endif /* PF_PIXFMT_H */
