/*
 * pf_vstage.h - the per-vertex stage (transform, Phong prologue, clipping, projection, triangle
 * emission), written once and compiled twice:
 *   - as C99 into the front end (pf_pipeline.c), for immediate mode, render lists and Gouraud lighting;
 *   - as CUDA device code into pfcu.cu (k_vertex_*), for large vertex-array draws, so that the host is
 *     not the triangle-rate limiter (SURVEY.md 8-f "next" row 1).
 * It restates, with the same order of float operations (no FMA on either side: -ffp-contract=off /
 * -fmad=false; IEEE division and square root on both), the reference's
 *   per-triangle prologue        src/internal/primitives/triangles.c:82-104  (minus Gouraud, host only)
 *   Process_ProjectAndClipTriangle, Process_ClipPolygonW/XYZ   triangles.c:157-280
 *   pfiLerpVertex, pfiHomogeneousToScreen                       src/internal/context/context.c:51-90
 */
#ifndef PF_VSTAGE_H
#define PF_VSTAGE_H

#include <stdint.h>
#include "pfcu.h"

#ifdef __CUDACC__
#  define PFV_FN __device__ __forceinline__
#  define PFV_SQRT(x) __fsqrt_rn(x)
#  define PFV_DIV(a, b) __fdiv_rn((a), (b))
#  define PFV_MUL(a, b) __fmul_rn((a), (b))
#  define PFV_ADD(a, b) __fadd_rn((a), (b))
#  define PFV_SUB(a, b) __fsub_rn((a), (b))
#  define PFV_U2F(u) __uint2float_rn(u)
#  define PFV_I2F(i) __int2float_rn(i)
#  define PFV_F2I(f) pfv_cvttss2si(f)
#else
#  include <math.h>
#  include <string.h>
#  define PFV_FN static inline
#  define PFV_SQRT(x) sqrtf(x)
#  define PFV_DIV(a, b) ((a) / (b))
#  define PFV_MUL(a, b) ((a) * (b))
#  define PFV_ADD(a, b) ((a) + (b))
#  define PFV_SUB(a, b) ((a) - (b))
#  define PFV_U2F(u) ((float)(u))
#  define PFV_I2F(i) ((float)(i))
#  define PFV_F2I(f) ((int)(f))
#endif

#define PFV_MAX_POLY 12
#define PFV_CLIP_EPSILON 1e-5f

typedef struct {                /* PFIvertex, src/internal/context/context.h:203-210 */
    float    homogeneous[4];
    float    screen[2];
    float    position[4];
    float    normal[3];
    float    texcoord[2];
    uint32_t color;             /* PFcolor dword */
} pfv_vertex;

typedef pfcu_vparams pfv_params;      /* declared in include/pfcu.h (it crosses the C-ABI) */

PFV_FN void pfv_copy(pfv_vertex *d, const pfv_vertex *s) { *d = *s; }

/* triangles.c:90-93: normal <- normalize(normal * matNormal); colour <- colour * diffuse / 255 */
PFV_FN void pfv_prologue(const pfv_params *p, int face, pfv_vertex *v)
{
    const float *m = p->normal_mat;
    const float nx = v->normal[0], ny = v->normal[1], nz = v->normal[2];
    float t[3];
    for (int j = 0; j < 3; j++)
        t[j] = PFV_ADD(PFV_ADD(PFV_ADD(PFV_MUL(m[j], nx), PFV_MUL(m[4 + j], ny)), PFV_MUL(m[8 + j], nz)), m[12 + j]);
    const float l2 = PFV_ADD(PFV_ADD(PFV_MUL(t[0], t[0]), PFV_MUL(t[1], t[1])), PFV_MUL(t[2], t[2]));
    if (l2 != 0.0f) {
        const float il = PFV_DIV(1.0f, PFV_SQRT(l2));
        t[0] = PFV_MUL(t[0], il); t[1] = PFV_MUL(t[1], il); t[2] = PFV_MUL(t[2], il);
    }
    /* pfmVec3Normalize leaves dst untouched for the zero vector: dst == the transformed vector here */
    v->normal[0] = t[0]; v->normal[1] = t[1]; v->normal[2] = t[2];
    const uint32_t c = v->color, d = p->diffuse[face];
    uint32_t o = 0;
    for (int i = 0; i < 4; i++) o |= ((((c >> (8 * i)) & 255u) * ((d >> (8 * i)) & 255u)) / 255u) << (8 * i);
    v->color = o;
}

/* ---- Gouraud vertex lighting: integer Blinn-Phong (lighting.c:23-144) --------------------------------------
 * One source for the host (real powf) and the device (the host-harvested specular table, see pfcu.h). */
#ifdef __CUDACC__
PFV_FN int pfv_cvttss2si(float f) { const int r = __float2int_rz(f); return (fabsf(f) < 2147483648.0f) ? r : (int)0x80000000; }
PFV_FN unsigned pfv_specular(const float *tab, float x)
{
    /* fmaxf(x, 0): NaN -> 0 */
    if (!(x > 0.0f)) x = 0.0f;
    int lo = 0, hi = PFCU_POW_TABLE_SIZE;                 /* number of thresholds <= x */
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (tab[mid] <= x) lo = mid + 1; else hi = mid; }
    return (unsigned)lo & 255u;                            /* (PFubyte) of 256 is 0 */
}
#  define PFV_SPECULAR(env, tabs, mface, x, shin) pfv_specular((tabs) + (size_t)(env)->pow_table[mface] * PFCU_POW_TABLE_SIZE, (x))
#else
#  define PFV_SPECULAR(env, tabs, mface, x, shin) ((unsigned)(uint8_t)(int)(255 * powf(fmaxf((x), 0.0f), (shin))))
#endif

PFV_FN unsigned pfv_min255(int n) { return (unsigned)(uint8_t)(n | ((255 - n) >> 31)); }

#define PFV_CH(c, i) ((int)(((c) >> (8 * (i))) & 255u))

PFV_FN uint32_t pfv_light_vertex(const pfcu_vparams_lit *e, const float *pow_tables, int mface, uint32_t diffuse,
                                 const float *P, const float *N)
{
    const pfcu_material *m = &e->material[mface];
    unsigned R = (unsigned)PFV_CH(m->emission, 0), G = (unsigned)PFV_CH(m->emission, 1), B = (unsigned)PFV_CH(m->emission, 2);
    const int dR = PFV_CH(diffuse, 0), dG = PFV_CH(diffuse, 1), dB = PFV_CH(diffuse, 2);
    const int aR = (int)(uint8_t)((PFV_CH(m->ambient, 0) * dR) / 255);
    const int aG = (int)(uint8_t)((PFV_CH(m->ambient, 1) * dG) / 255);
    const int aB = (int)(uint8_t)((PFV_CH(m->ambient, 2) * dB) / 255);

    float V[3], vl2 = 0.0f;
    for (int i = 0; i < 3; i++) { V[i] = PFV_SUB(e->view_pos[i], P[i]); vl2 = PFV_ADD(vl2, PFV_MUL(V[i], V[i])); }
    { const float il = PFV_DIV(1.0f, PFV_SQRT(vl2)); for (int i = 0; i < 3; i++) V[i] = PFV_MUL(V[i], il); }

    const float shininess = m->shininess;
    (void)shininess; (void)pow_tables;

    for (uint32_t li = 0; li < e->n_lights; li++) {
        const pfcu_light *l = &e->lights[li];
        unsigned lR = 0, lG = 0, lB = 0;
        float L[3] = { PFV_SUB(l->position[0], P[0]), PFV_SUB(l->position[1], P[1]), PFV_SUB(l->position[2], P[2]) };
        const float d2 = PFV_ADD(PFV_ADD(PFV_MUL(L[0], L[0]), PFV_MUL(L[1], L[1])), PFV_MUL(L[2], L[2]));
        float dist = 0.0f;
        if (d2 != 0.0f) {
            dist = PFV_SQRT(d2);
            const float il = PFV_DIV(1.0f, dist);
            L[0] = PFV_MUL(L[0], il); L[1] = PFV_MUL(L[1], il); L[2] = PFV_MUL(L[2], il);
        }
        unsigned intensity = 255;
        int skip = 0;
        if (l->inner_cutoff < (float)3.14159265358979323846) {
            const float theta = PFV_ADD(PFV_ADD(PFV_MUL(L[0], -l->direction[0]), PFV_MUL(L[1], -l->direction[1])), PFV_MUL(L[2], -l->direction[2]));
            const float eps = PFV_SUB(l->inner_cutoff, l->outer_cutoff);
            const int iv = PFV_F2I(PFV_DIV(PFV_MUL(255.0f, PFV_SUB(theta, l->outer_cutoff)), eps));
            intensity = (unsigned)(uint8_t)(iv < 0 ? 0 : (iv > 255 ? 255 : iv));
            if (intensity == 0) skip = 1;
        }
        unsigned attenuation = 255;
        if (!skip && (l->att_linear != 0.0f || l->att_quadratic != 0.0f)) {
            attenuation = (unsigned)(uint8_t)PFV_F2I(PFV_DIV(255.0f, PFV_ADD(PFV_ADD(l->att_constant, PFV_MUL(l->att_linear, dist)), PFV_MUL(l->att_quadratic, d2))));
            if (attenuation == 0) skip = 1;
        }
        if (!skip) {
            const unsigned factor = (unsigned)(uint8_t)((intensity * attenuation) / 255);
            const int di = PFV_F2I(PFV_MUL(255.0f, PFV_ADD(PFV_ADD(PFV_MUL(N[0], L[0]), PFV_MUL(N[1], L[1])), PFV_MUL(N[2], L[2]))));
            const int diff = (int)(uint8_t)(di > 0 ? di : 0);
            lR = pfv_min255((int)lR + (dR * PFV_CH(l->diffuse, 0) * diff) / (255 * 255));
            lG = pfv_min255((int)lG + (dG * PFV_CH(l->diffuse, 1) * diff) / (255 * 255));
            lB = pfv_min255((int)lB + (dB * PFV_CH(l->diffuse, 2) * diff) / (255 * 255));

            float H[3] = { PFV_ADD(L[0], V[0]), PFV_ADD(L[1], V[1]), PFV_ADD(L[2], V[2]) };
            const float hl2 = PFV_ADD(PFV_ADD(PFV_MUL(H[0], H[0]), PFV_MUL(H[1], H[1])), PFV_MUL(H[2], H[2]));
            if (hl2 != 0.0f) {                  /* pfmVec3Normalize leaves the zero vector alone */
                const float il = PFV_DIV(1.0f, PFV_SQRT(hl2));
                H[0] = PFV_MUL(H[0], il); H[1] = PFV_MUL(H[1], il); H[2] = PFV_MUL(H[2], il);
            }
            const float ndh = PFV_ADD(PFV_ADD(PFV_MUL(N[0], H[0]), PFV_MUL(N[1], H[1])), PFV_MUL(N[2], H[2]));
            const int spec = (int)PFV_SPECULAR(e, pow_tables, mface, ndh, shininess);
            lR = pfv_min255((int)lR + (PFV_CH(m->specular, 0) * PFV_CH(l->specular, 0) * spec) / (255 * 255));
            lG = pfv_min255((int)lG + (PFV_CH(m->specular, 1) * PFV_CH(l->specular, 1) * spec) / (255 * 255));
            lB = pfv_min255((int)lB + (PFV_CH(m->specular, 2) * PFV_CH(l->specular, 2) * spec) / (255 * 255));

            lR = (unsigned)(uint8_t)((lR * factor) / 255);
            lG = (unsigned)(uint8_t)((lG * factor) / 255);
            lB = (unsigned)(uint8_t)((lB * factor) / 255);
        }
        R = pfv_min255((int)(R + lR) + (aR * PFV_CH(l->ambient, 0)) / 255);
        G = pfv_min255((int)(G + lG) + (aG * PFV_CH(l->ambient, 1)) / 255);
        B = pfv_min255((int)(B + lB) + (aB * PFV_CH(l->ambient, 2)) / 255);
    }
    return R | (G << 8) | (B << 16) | (diffuse & 0xff000000u);
}

/* The per-triangle prologue of one vertex (triangles.c:88-103): normal transform, colour * diffuse, and
 * Gouraud lighting with the material picked by the side of the normal relative to the view axis. */
PFV_FN void pfv_prologue_lit(const pfcu_vparams_lit *e, const float *pow_tables, int face, pfv_vertex *v)
{
    pfv_prologue(&e->base, face, v);
    if (e->gouraud) {
        const float ndv = PFV_ADD(PFV_ADD(PFV_MUL(v->normal[0], e->view_z[0]), PFV_MUL(v->normal[1], e->view_z[1])), PFV_MUL(v->normal[2], e->view_z[2]));
        v->color = pfv_light_vertex(e, pow_tables, (ndv < 0) ? 0 : 1, v->color, v->position, v->normal);
    }
}

PFV_FN void pfv_to_screen(const pfv_params *p, pfv_vertex *v)
{
    v->screen[0] = PFV_ADD(PFV_ADD(PFV_I2F(p->vp_pos[0]), PFV_MUL(PFV_MUL(PFV_ADD(v->homogeneous[0], 1.0f), 0.5f), PFV_U2F(p->vp_dim[0]))), 0.5f);
    v->screen[1] = PFV_ADD(PFV_ADD(PFV_I2F(p->vp_pos[1]), PFV_MUL(PFV_MUL(PFV_SUB(1.0f, v->homogeneous[1]), 0.5f), PFV_U2F(p->vp_dim[1]))), 0.5f);
}

PFV_FN float pfv_mix(float a, float b, float t) { return PFV_ADD(a, PFV_MUL(t, PFV_SUB(b, a))); }

PFV_FN void pfv_lerp(pfv_vertex *r, const pfv_vertex *a, const pfv_vertex *b, float t)
{
    r->screen[0] = 0.0f; r->screen[1] = 0.0f;
    const int ut = (int)(uint8_t)(int)PFV_MUL(255.0f, t);            /* (PFubyte)(255*t) */
    uint32_t col = 0;
    for (int i = 0; i < 4; i++) {
        r->homogeneous[i] = pfv_mix(a->homogeneous[i], b->homogeneous[i], t);
        r->position[i] = pfv_mix(a->position[i], b->position[i], t);
        const int ca = (int)((a->color >> (8 * i)) & 255u), cb = (int)((b->color >> (8 * i)) & 255u);
        col |= (uint32_t)(uint8_t)(ca + (ut * (cb - ca)) / 255) << (8 * i);
        if (i < 2) r->texcoord[i] = pfv_mix(a->texcoord[i], b->texcoord[i], t);
        if (i < 3) r->normal[i] = pfv_mix(a->normal[i], b->normal[i], t);
    }
    r->color = col;
}

/* Sutherland-Hodgman against w >= eps (triangles.c:157-182) */
PFV_FN int pfv_clip_w(pfv_vertex *poly, int *n)
{
    pfv_vertex in[PFV_MAX_POLY];
    const int nin = *n;
    for (int i = 0; i < nin; i++) pfv_copy(&in[i], &poly[i]);
    *n = 0;
    const pfv_vertex *prev = &in[nin - 1];
    int pd = (prev->homogeneous[3] < PFV_CLIP_EPSILON) ? -1 : 1;
    for (int i = 0; i < nin; i++) {
        const int cd = (in[i].homogeneous[3] < PFV_CLIP_EPSILON) ? -1 : 1;
        if (pd * cd < 0 && *n < PFV_MAX_POLY) {
            const float t = PFV_DIV(PFV_SUB(PFV_CLIP_EPSILON, prev->homogeneous[3]), PFV_SUB(in[i].homogeneous[3], prev->homogeneous[3]));
            pfv_lerp(&poly[(*n)++], prev, &in[i], t);
        }
        if (cd > 0 && *n < PFV_MAX_POLY) pfv_copy(&poly[(*n)++], &in[i]);
        pd = cd; prev = &in[i];
    }
    return *n > 0;
}

/* ... then against +-x, +-y, +-z <= w (triangles.c:184-244) */
PFV_FN int pfv_clip_xyz(pfv_vertex *poly, int *n)
{
    for (int ax = 0; ax < 3; ax++) {
        if (*n == 0) return 0;
        for (int side = 0; side < 2; side++) {
            pfv_vertex in[PFV_MAX_POLY];
            const int nin = *n;
            for (int i = 0; i < nin; i++) pfv_copy(&in[i], &poly[i]);
            *n = 0;
            const pfv_vertex *prev = &in[nin - 1];
            int pd = ((side ? -prev->homogeneous[ax] : prev->homogeneous[ax]) <= prev->homogeneous[3]) ? 1 : -1;
            for (int i = 0; i < nin; i++) {
                const pfv_vertex *cur = &in[i];
                const int cd = ((side ? -cur->homogeneous[ax] : cur->homogeneous[ax]) <= cur->homogeneous[3]) ? 1 : -1;
                if (pd * cd <= 0 && *n < PFV_MAX_POLY) {
                    float t;
                    if (!side) {
                        const float pn = PFV_SUB(prev->homogeneous[3], prev->homogeneous[ax]);
                        t = PFV_DIV(pn, PFV_SUB(pn, PFV_SUB(cur->homogeneous[3], cur->homogeneous[ax])));
                    } else {
                        const float pn = PFV_ADD(prev->homogeneous[3], prev->homogeneous[ax]);
                        t = PFV_DIV(pn, PFV_SUB(pn, PFV_ADD(cur->homogeneous[3], cur->homogeneous[ax])));
                    }
                    pfv_lerp(&poly[(*n)++], prev, cur, t);
                }
                if (cd > 0 && *n < PFV_MAX_POLY) pfv_copy(&poly[(*n)++], cur);
                pd = cd; prev = cur;
            }
            if (*n == 0) return 0;
        }
    }
    return *n > 0;
}

PFV_FN void pfv_transform(const pfv_params *p, pfv_vertex *v)
{
    const float *m = p->mvp;
    const float x = v->position[0], y = v->position[1], z = v->position[2], w = v->position[3];
    for (int j = 0; j < 4; j++)
        v->homogeneous[j] = PFV_ADD(PFV_ADD(PFV_ADD(PFV_MUL(m[j], x), PFV_MUL(m[4 + j], y)), PFV_MUL(m[8 + j], z)), PFV_MUL(m[12 + j], w));
}

PFV_FN void pfv_perspective(const pfv_params *p, pfv_vertex *v)     /* triangles.c:265-276 */
{
    v->homogeneous[2] = PFV_DIV(1.0f, v->homogeneous[2]);
    v->texcoord[0] = PFV_MUL(v->texcoord[0], v->homogeneous[2]);
    v->texcoord[1] = PFV_MUL(v->texcoord[1], v->homogeneous[2]);
    const float iw = PFV_DIV(1.0f, v->homogeneous[3]);
    v->homogeneous[0] = PFV_MUL(v->homogeneous[0], iw);
    v->homogeneous[1] = PFV_MUL(v->homogeneous[1], iw);
    pfv_to_screen(p, v);
}

/* Process_ProjectAndClipTriangle (triangles.c:246-280).  poly[0..2] in, poly[0..*n-1] out; returns is3D.
 * Trivial accept: when all three vertices are inside every plane, Sutherland-Hodgman returns the polygon
 * unchanged, so the seven clipping passes are skipped (tests are written exactly as the clippers' tests;
 * NaNs fall through to the full path). */
PFV_FN int pfv_project_and_clip(const pfv_params *p, pfv_vertex *poly, int *n)
{
    float wsum = 0.0f;
    for (int i = 0; i < 3; i++) { pfv_transform(p, &poly[i]); wsum = PFV_ADD(wsum, poly[i].homogeneous[3]); }
    float dw = PFV_SUB(wsum, 3.0f); if (dw < 0.0f) dw = -dw;
    if (dw < PFV_CLIP_EPSILON) {
        for (int i = 0; i < 3; i++) pfv_to_screen(p, &poly[i]);
        return 0;
    }
    int inside = 1;
    for (int i = 0; inside && i < 3; i++) {
        const float *h = poly[i].homogeneous; const float w = h[3];
        inside = !(w < PFV_CLIP_EPSILON) && h[0] <= w && -h[0] <= w && h[1] <= w && -h[1] <= w && h[2] <= w && -h[2] <= w;
    }
    if (inside || (pfv_clip_w(poly, n) && pfv_clip_xyz(poly, n)))
        for (int i = 0; i < *n; i++) pfv_perspective(p, &poly[i]);
    return 1;
}

PFV_FN void pfv_emit(pfcu_triangle *t, const pfv_vertex *a, const pfv_vertex *b, const pfv_vertex *c, uint32_t state, int face, int is3d)
{
    const pfv_vertex *vs[3] = { a, b, c };
    for (int i = 0; i < 3; i++) {
        const pfv_vertex *v = vs[i];
        pfcu_vertex *o = &t->v[i];
        o->sx = v->screen[0]; o->sy = v->screen[1];
        o->zinv = v->homogeneous[2];
        o->u = v->texcoord[0]; o->v = v->texcoord[1];
        o->px = v->position[0]; o->py = v->position[1]; o->pz = v->position[2];
        o->nx = v->normal[0]; o->ny = v->normal[1]; o->nz = v->normal[2];
        o->rgba = v->color;
    }
    t->state = state; t->face = (uint8_t)face; t->is3d = (uint8_t)is3d; t->pad = 0;
}

#endif /* PF_VSTAGE_H */
