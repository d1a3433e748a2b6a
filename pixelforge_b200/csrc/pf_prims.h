/*
 * pf_prims.h - points and lines: the per-fragment arithmetic of the reference's scalar point / line
 * rasterisers, written once and compiled as CUDA device code (k_prims in pfcu.cu) and as C99 (the oracle).
 * Restates, in behaviour,
 *   Rasterize_Line_NODEPTH / _DEPTH / _THICK_*      src/internal/primitives/lines.c:283-530
 *   Rasterize_Point_NODEPTH / _DEPTH                src/internal/primitives/points.c:85-183
 *   pfiColorLerpSmooth (scalar)                     src/internal/color.h:28-37
 *   scalar blend table                              src/internal/blend.h:29-130
 *   scalar depth table                              src/internal/depth.h:28-78
 * These paths use the SCALAR tables, which differ from the SIMD ones of the triangle path: PF_NOTEQUAL really is
 * "not equal", PF_BLEND_SUB subtracts, multiply divides by 255.
 * Quirks kept: the last pixel of a line is not drawn (i != endVal); a pixel is addressed as y*W + x in
 * unsigned 32-bit arithmetic without a range check on x, so a column just outside the surface lands in the
 * neighbouring row (writes outside the buffer, which corrupt memory upstream, are dropped here); thick lines
 * draw their centre line with the depth-tested routine even when PF_DEPTH_TEST is off.
 */
#ifndef PF_PRIMS_H
#define PF_PRIMS_H

#include "pf_vstage.h"

#define PFP_KIND_POINT 0
#define PFP_KIND_LINE  1

typedef struct {            /* one plain line: what Rasterize_Line_* derives before its loop */
    int x1, y1, end_val, sgn_inc, dec_inc, y_longer;
    float inv_end_val;
} pfp_line;

PFV_FN int pfp_iabs(int v) { return v < 0 ? -v : v; }

PFV_FN void pfp_line_setup(pfp_line *L, float sx1, float sy1, float sx2, float sy2)
{
    const int x1 = PFV_F2I(sx1), y1 = PFV_F2I(sy1), x2 = PFV_F2I(sx2), y2 = PFV_F2I(sy2);
    int short_len = (int)((unsigned)y2 - (unsigned)y1), long_len = (int)((unsigned)x2 - (unsigned)x1);
    L->y_longer = 0;
    if (pfp_iabs(short_len) > pfp_iabs(long_len)) { const int t = short_len; short_len = long_len; long_len = t; L->y_longer = 1; }
    L->inv_end_val = PFV_DIV(1.0f, PFV_I2F(long_len));
    L->end_val = long_len;
    L->sgn_inc = 1;
    if (long_len < 0) { long_len = (int)(0u - (unsigned)long_len); L->sgn_inc = -1; }
    L->dec_inc = (long_len == 0) ? 0 : (int)((unsigned)short_len << 16) / long_len;
    L->x1 = x1; L->y1 = y1;
}

PFV_FN unsigned pfp_line_steps(const pfp_line *L) { return (unsigned)pfp_iabs(L->end_val); }

/* step k of the loop: linear pixel offset (unsigned wrap as upstream) and the interpolation parameter */
PFV_FN uint32_t pfp_line_step(const pfp_line *L, unsigned k, uint32_t W, float *t)
{
    const int i = (int)k * L->sgn_inc;
    const int j = (int)((unsigned)k * (unsigned)L->dec_inc);
    *t = PFV_MUL(PFV_I2F(i), L->inv_end_val);
    const int x = L->y_longer ? (int)((unsigned)L->x1 + (unsigned)(j >> 16)) : (int)((unsigned)L->x1 + (unsigned)i);
    const int y = L->y_longer ? (int)((unsigned)L->y1 + (unsigned)i) : (int)((unsigned)L->y1 + (unsigned)(j >> 16));
    return (uint32_t)y * W + (uint32_t)x;
}

PFV_FN uint32_t pfp_color_lerp(uint32_t a, uint32_t b, float t)      /* (PFubyte)(a + t*(b - a)) per channel */
{
    uint32_t o = 0;
    for (int i = 0; i < 4; i++) {
        const int ca = (int)((a >> (8 * i)) & 255u), cb = (int)((b >> (8 * i)) & 255u);
        const float v = PFV_ADD(PFV_I2F(ca), PFV_MUL(t, PFV_I2F(cb - ca)));
        o |= ((uint32_t)PFV_F2I(v) & 255u) << (8 * i);
    }
    return o;
}

PFV_FN uint32_t pfp_blend(int mode, uint32_t s, uint32_t d)
{
    const unsigned alpha = (s >> 24) + 1u, inv = 256u - alpha;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const int sv = (int)((s >> (8 * i)) & 255u), dv = (int)((d >> (8 * i)) & 255u);
        int o;
        switch (mode) {
        case 0:  o = (sv + dv) >> 1; break;
        case 1:  o = (int)(((i == 3 ? alpha * 255u : alpha * (unsigned)sv) + inv * (unsigned)dv) >> 8); break;
        case 2:  o = sv + dv > 255 ? 255 : sv + dv; break;
        case 3:  o = dv - sv < 0 ? 0 : dv - sv; break;
        case 4:  o = (sv * dv) / 255; break;
        case 5:  { const int v = ((dv * (255 - sv)) >> 8) + sv; o = v > 255 ? 255 : v; } break;
        case 6:  o = sv > dv ? sv : dv; break;
        default: o = sv < dv ? sv : dv; break;
        }
        r |= ((uint32_t)o & 255u) << (8 * i);
    }
    return r;
}

PFV_FN int pfp_depth(int func, float z, float zb)
{
    switch (func) {
    case 0: return z == zb;
    case 1: return z != zb;
    case 2: return z < zb;
    case 3: return z <= zb;
    case 4: return z > zb;
    default: return z >= zb;
    }
}

/* A thick line (lineWidth > 1.5) is its centre line followed by pairs of parallel lines shifted by -i, +i
 * along y (mostly horizontal) or x (otherwise).  Returns the number of plain lines; *axis = 1: shift y, 0: x. */
PFV_FN unsigned pfp_thick_count(float sx1, float sy1, float sx2, float sy2, float width, int *axis)
{
    const int x1 = PFV_F2I(sx1), y1 = PFV_F2I(sy1), x2 = PFV_F2I(sx2), y2 = PFV_F2I(sy2);
    const int dx = (int)((unsigned)x2 - (unsigned)x1), dy = (int)((unsigned)y2 - (unsigned)y1);
    *axis = 0;
    if (!(width > 1.5f)) return 1u;
    const float len2 = PFV_I2F((int)((unsigned)dx * (unsigned)dx + (unsigned)dy * (unsigned)dy));
    const float k = PFV_MUL(PFV_SUB(width, 1.0f), PFV_DIV(1.0f, PFV_SQRT(len2)));
    if (dx != 0 && pfp_iabs(dy / dx) < 1) {
        const int wy = PFV_F2I(PFV_MUL(k, PFV_I2F(pfp_iabs(dx)))) >> 1;
        *axis = 1;
        return wy > 0 ? 1u + 2u * (unsigned)wy : 1u;
    } else if (dy != 0) {
        const int wx = PFV_F2I(PFV_MUL(k, PFV_I2F(pfp_iabs(dy)))) >> 1;
        return wx > 0 ? 1u + 2u * (unsigned)wx : 1u;
    }
    return 1u;
}

/* the shift of plain line s (0 = centre, then -1, +1, -2, +2 ...) */
PFV_FN float pfp_thick_shift(unsigned s) { const int i = (int)((s + 1u) >> 1); return PFV_I2F((s & 1u) ? -i : i); }

#endif /* PF_PRIMS_H */
