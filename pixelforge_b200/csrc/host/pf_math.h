/*
 * pf_math.h - the handful of matrix / vector routines the front end needs.
 *
 * Storage and conventions follow the reference's pfm.h (column-major storage with row-vector
 * products: v' = v * M, element [col*4 + row] in OpenGL terms -> out[j] = sum_k v[k] * M[k*4 + j]).
 * Operation ORDER matters: projected vertices are truncated to integer pixels by the rasteriser
 * (triangles.c:297-299), so every sum below is written in the reference's left-to-right order
 * (pfm.h:954-963,1507-1517,1996-2006,2082-2122,2176-2198,2218-2260,2368-2441) and the translation
 * unit is compiled with -ffp-contract=off.
 */
#ifndef PF_MATH_H
#define PF_MATH_H

#include <math.h>
#include <string.h>

static inline void m4_identity(float *m)
{
    memset(m, 0, 16 * sizeof(float));
    m[0] = m[5] = m[10] = m[15] = 1.0f;
}

static inline void m4_copy(float *d, const float *s) { memcpy(d, s, 16 * sizeof(float)); }

/* d = l * r ; d may alias l or r */
static inline void m4_mul(float *d, const float *l, const float *r)
{
    float t[16];
    for (int i = 0; i < 4; i++) {
        for (int j = 0; j < 4; j++) {
            float s = 0.0f;
            s += l[i * 4 + 0] * r[0 * 4 + j];
            s += l[i * 4 + 1] * r[1 * 4 + j];
            s += l[i * 4 + 2] * r[2 * 4 + j];
            s += l[i * 4 + 3] * r[3 * 4 + j];
            t[i * 4 + j] = s;
        }
    }
    memcpy(d, t, sizeof t);
}

static inline void m4_transpose(float *d, const float *s)
{
    float t[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) t[i * 4 + j] = s[j * 4 + i];
    memcpy(d, t, sizeof t);
}

/* cofactor inverse, same 2x2 sub-determinant grouping as pfm.h:2082-2122 */
static inline void m4_invert(float *d, const float *s)
{
    const float a00 = s[0], a01 = s[1], a02 = s[2], a03 = s[3];
    const float a10 = s[4], a11 = s[5], a12 = s[6], a13 = s[7];
    const float a20 = s[8], a21 = s[9], a22 = s[10], a23 = s[11];
    const float a30 = s[12], a31 = s[13], a32 = s[14], a33 = s[15];

    const float b00 = a00 * a11 - a01 * a10, b01 = a00 * a12 - a02 * a10;
    const float b02 = a00 * a13 - a03 * a10, b03 = a01 * a12 - a02 * a11;
    const float b04 = a01 * a13 - a03 * a11, b05 = a02 * a13 - a03 * a12;
    const float b06 = a20 * a31 - a21 * a30, b07 = a20 * a32 - a22 * a30;
    const float b08 = a20 * a33 - a23 * a30, b09 = a21 * a32 - a22 * a31;
    const float b10 = a21 * a33 - a23 * a31, b11 = a22 * a33 - a23 * a32;

    const float inv = 1.0f / (b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06);

    float t[16];
    t[0]  = (a11 * b11 - a12 * b10 + a13 * b09) * inv;
    t[1]  = (-a01 * b11 + a02 * b10 - a03 * b09) * inv;
    t[2]  = (a31 * b05 - a32 * b04 + a33 * b03) * inv;
    t[3]  = (-a21 * b05 + a22 * b04 - a23 * b03) * inv;
    t[4]  = (-a10 * b11 + a12 * b08 - a13 * b07) * inv;
    t[5]  = (a00 * b11 - a02 * b08 + a03 * b07) * inv;
    t[6]  = (-a30 * b05 + a32 * b02 - a33 * b01) * inv;
    t[7]  = (a20 * b05 - a22 * b02 + a23 * b01) * inv;
    t[8]  = (a10 * b10 - a11 * b08 + a13 * b06) * inv;
    t[9]  = (-a00 * b10 + a01 * b08 - a03 * b06) * inv;
    t[10] = (a30 * b04 - a31 * b02 + a33 * b00) * inv;
    t[11] = (-a20 * b04 + a21 * b02 - a23 * b00) * inv;
    t[12] = (-a10 * b09 + a11 * b07 - a12 * b06) * inv;
    t[13] = (a00 * b09 - a01 * b07 + a02 * b06) * inv;
    t[14] = (-a30 * b03 + a31 * b01 - a32 * b00) * inv;
    t[15] = (a20 * b03 - a21 * b01 + a22 * b00) * inv;
    memcpy(d, t, sizeof t);
}

static inline void m4_translate(float *m, float x, float y, float z)
{
    m4_identity(m);
    m[12] = x; m[13] = y; m[14] = z;
}

static inline void m4_scale(float *m, float x, float y, float z)
{
    memset(m, 0, 16 * sizeof(float));
    m[0] = x; m[5] = y; m[10] = z; m[15] = 1.0f;
}

/* axis-angle rotation (radians); the axis is normalised unless |axis|^2 is exactly 0 or 1 */
static inline void m4_rotate(float *m, float x, float y, float z, float angle)
{
    m4_identity(m);
    float l2 = x * x + y * y + z * z;
    if (l2 != 1.0f && l2 != 0.0f) {
        float il = 1.0f / sqrtf(l2);
        x *= il; y *= il; z *= il;
    }
    float s = sinf(angle), c = cosf(angle), t = 1.0f - c;
    m[0] = x * x * t + c;     m[1] = y * x * t + z * s; m[2]  = z * x * t - y * s;
    m[4] = x * y * t - z * s; m[5] = y * y * t + c;     m[6]  = z * y * t + x * s;
    m[8] = x * z * t + y * s; m[9] = y * z * t - x * s; m[10] = z * z * t + c;
}

static inline void m4_frustum(float *m, float l, float r, float b, float t, float n, float f)
{
    memset(m, 0, 16 * sizeof(float));
    float rl = r - l, tb = t - b, fn = f - n;
    m[0] = (n * 2.0f) / rl;
    m[5] = (n * 2.0f) / tb;
    m[8] = (r + l) / rl;
    m[9] = (t + b) / tb;
    m[10] = -(f + n) / fn;
    m[11] = -1.0f;
    m[14] = -(f * n * 2.0f) / fn;
}

static inline void m4_ortho(float *m, float l, float r, float b, float t, float n, float f)
{
    memset(m, 0, 16 * sizeof(float));
    float rl = r - l, tb = t - b, fn = f - n;
    m[0] = 2.0f / rl;
    m[5] = 2.0f / tb;
    m[10] = -2.0f / fn;
    m[12] = -(l + r) / rl;
    m[13] = -(t + b) / tb;
    m[14] = -(f + n) / fn;
    m[15] = 1.0f;
}

static inline void v4_transform(float *d, const float *v, const float *m)
{
    float t[4];
    for (int j = 0; j < 4; j++)
        t[j] = m[j] * v[0] + m[4 + j] * v[1] + m[8 + j] * v[2] + m[12 + j] * v[3];
    memcpy(d, t, sizeof t);
}

static inline void v3_transform(float *d, const float *v, const float *m)   /* implicit w = 1 */
{
    float t[3];
    for (int j = 0; j < 3; j++)
        t[j] = m[j] * v[0] + m[4 + j] * v[1] + m[8 + j] * v[2] + m[12 + j];
    memcpy(d, t, sizeof t);
}

static inline void v2_transform(float *d, const float *v, const float *m)
{
    float t0 = m[0] * v[0] + m[4] * v[1] + m[12];
    float t1 = m[1] * v[0] + m[5] * v[1] + m[13];
    d[0] = t0; d[1] = t1;
}

static inline float v3_dot(const float *a, const float *b)
{
    float s = 0.0f;
    s += a[0] * b[0]; s += a[1] * b[1]; s += a[2] * b[2];
    return s;
}

static inline void v3_normalize(float *d, const float *v)      /* leaves d untouched for the zero vector */
{
    float l2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    if (l2 == 0.0f) return;
    float il = 1.0f / sqrtf(l2);
    d[0] = v[0] * il; d[1] = v[1] * il; d[2] = v[2] * il;
}

#endif
