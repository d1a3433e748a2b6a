/*
 * pf_x86approx.c - capture the host CPU's RCPPS / RSQRTPS behaviour as lookup tables.
 *
 * The reference computes depth, perspective texcoords and every Phong normalisation with the x86
 * approximate-reciprocal instructions (src/internal/simd.h:1217-1245: _mm256_rcp_ps,
 * _mm256_rsqrt_ps).  Their results are implementation defined (they differ between CPU vendors),
 * so "identical depth-test masks as the reference on the same box" requires reproducing THIS
 * host's instruction on the GPU.  On every x86 implementation we know the result is a pure function
 * of the sign, the exponent (by plain scaling) and the top K mantissa bits; this file finds the
 * smallest such K by exhaustive verification over all 2^23 mantissas, verifies the exponent
 * scaling, and hands 2^K-entry tables to the CUDA layer (pfcu_set_approx_tables), where
 * rcp_x86()/rsqrt_x86() evaluate  table[mantissa >> (23-K)]  and re-bias the exponent.
 *
 * This is not a CPU fallback: nothing is rendered here.  It runs once per process (~20 ms).
 */
#include "pf_internal.h"

#include <immintrin.h>
#include <stdio.h>

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float    u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float hw_rcp(float x)   { return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x))); }
static inline float hw_rsqrt(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }

/* does f(1.m * 2^(ebias-127)) depend only on the top K mantissa bits?  exhaustive. */
static int depends_only_on_top_bits(float (*f)(float), uint32_t ebias, int K)
{
    const uint32_t shift = 23u - (uint32_t)K, e = ebias << 23;
    for (uint32_t top = 0; top < (1u << K); top++) {
        const uint32_t base = e | (top << shift);
        const uint32_t r0 = f2u(f(u2f(base)));
        for (uint32_t lo = 1; lo < (1u << shift); lo++)
            if (f2u(f(u2f(base | lo))) != r0) return 0;
    }
    return 1;
}

static int find_bits(float (*f)(float), uint32_t ebias)
{
    /* cheap screening with the extreme low-bit patterns, then the exhaustive proof */
    for (int K = 8; K <= 23; K++) {
        const uint32_t shift = 23u - (uint32_t)K, e = ebias << 23;
        int ok = 1;
        for (uint32_t top = 0; top < (1u << K) && ok; top++) {
            const uint32_t base = e | (top << shift);
            const uint32_t r0 = f2u(f(u2f(base)));
            if (shift && (f2u(f(u2f(base | ((1u << shift) - 1u)))) != r0 || f2u(f(u2f(base | (1u << (shift - 1))))) != r0)) ok = 0;
        }
        if (ok && depends_only_on_top_bits(f, ebias, K)) return K;
    }
    return 23;
}

int pfh_harvest_rcp(uint32_t **table, int *bits)
{
    int K = find_bits(hw_rcp, 127u);
    uint32_t *t = (uint32_t *)malloc(sizeof(uint32_t) << K);
    if (!t) return 0;
    for (uint32_t i = 0; i < (1u << K); i++) t[i] = f2u(hw_rcp(u2f((127u << 23) | (i << (23 - K)))));
    /* exponent scaling: rcp(1.m * 2^E) == rcp(1.m) * 2^-E (flushed to zero when subnormal) */
    for (uint32_t e = 1; e <= 254; e++) {
        for (uint32_t i = 0; i < (1u << K); i += 61) {
            uint32_t got = f2u(hw_rcp(u2f((e << 23) | (i << (23 - K)))));
            int32_t ex = (int32_t)(t[i] >> 23) + 127 - (int32_t)e;
            uint32_t want = ex <= 0 ? 0u : ((uint32_t)ex << 23) | (t[i] & 0x7fffffu);
            if (got != want) {
                fprintf(stderr, "pixelforge-b200: RCPPS on this CPU is not exponent-invariant (e=%u i=%u got %08x want %08x)\n", e, i, got, want);
                free(t); return 0;
            }
        }
    }
    *table = t; *bits = K;
    return 1;
}

int pfh_harvest_rsqrt(uint32_t **table, int *bits)
{
    int K0 = find_bits(hw_rsqrt, 127u), K1 = find_bits(hw_rsqrt, 128u);
    int K = K0 > K1 ? K0 : K1;
    uint32_t *t = (uint32_t *)malloc(sizeof(uint32_t) << (K + 1));
    if (!t) return 0;
    for (uint32_t odd = 0; odd < 2; odd++)
        for (uint32_t i = 0; i < (1u << K); i++)
            t[(odd << K) | i] = f2u(hw_rsqrt(u2f(((127u + odd) << 23) | (i << (23 - K)))));
    for (uint32_t e = 1; e <= 254; e++) {
        uint32_t odd = (e & 1u) ? 0u : 1u;               /* unbiased exponent parity */
        int32_t half = ((int32_t)e - 127 - (int32_t)odd) / 2;
        for (uint32_t i = 0; i < (1u << K); i += 61) {
            uint32_t got = f2u(hw_rsqrt(u2f((e << 23) | (i << (23 - K)))));
            uint32_t tv = t[(odd << K) | i];
            uint32_t want = (uint32_t)((int32_t)(tv >> 23) - half) << 23 | (tv & 0x7fffffu);
            if (got != want) {
                fprintf(stderr, "pixelforge-b200: RSQRTPS on this CPU is not exponent-invariant (e=%u i=%u got %08x want %08x)\n", e, i, got, want);
                free(t); return 0;
            }
        }
    }
    *table = t; *bits = K;
    return 1;
}

/* Host-side evaluation of the table formulas exactly as the CUDA device functions do it; exported
 * so that CPU-only tests can compare them with the hardware instructions. */
PF_API float pfh_rcp_from_table(const uint32_t *t, int K, float x)
{
    uint32_t u = f2u(x), s = u & 0x80000000u, e = (u >> 23) & 255u, m = u & 0x7fffffu;
    if (e == 255u) return m ? u2f(u | 0x00400000u) : u2f(s);           /* NaN -> quiet NaN, inf -> 0 */
    if (e == 0u) return u2f(s | 0x7f800000u);                           /* zero and denormals -> inf  */
    uint32_t tv = t[m >> (23 - K)];
    int32_t ex = (int32_t)(tv >> 23) + 127 - (int32_t)e;
    if (ex <= 0) return u2f(s);
    return u2f(s | ((uint32_t)ex << 23) | (tv & 0x7fffffu));
}

PF_API float pfh_rsqrt_from_table(const uint32_t *t, int K, float x)
{
    uint32_t u = f2u(x), s = u & 0x80000000u, e = (u >> 23) & 255u, m = u & 0x7fffffu;
    if (e == 255u && m) return u2f(u | 0x00400000u);
    if (e == 0u) return u2f(s | 0x7f800000u);                           /* +-0, denormals -> +-inf    */
    if (s) return u2f(0xffc00000u);                                      /* negative -> default NaN    */
    if (e == 255u) return 0.0f;
    uint32_t odd = (e & 1u) ? 0u : 1u;
    int32_t half = ((int32_t)e - 127 - (int32_t)odd) / 2;
    uint32_t tv = t[(odd << K) | (m >> (23 - K))];
    return u2f(((uint32_t)((int32_t)(tv >> 23) - half) << 23) | (tv & 0x7fffffu));
}

PF_API float pfh_hw_rcp(float x) { return hw_rcp(x); }
PF_API float pfh_hw_rsqrt(float x) { return hw_rsqrt(x); }
PF_API int pfh_harvest_tables(uint32_t **rcp, int *rcp_bits, uint32_t **rsqrt, int *rsqrt_bits)
{
    return pfh_harvest_rcp(rcp, rcp_bits) && pfh_harvest_rsqrt(rsqrt, rsqrt_bits);
}
