/* pfcu_device_math.cuh - device arithmetic shared by the kernels: x86 lane semantics, colour math, texturing, per-fragment Blinn-Phong.
 * Part of the single translation unit pfcu.cu (included there, in order; not a stand-alone header). */

/* ------------------------------------------------------------------------------------------------ */
/* device: x86 lane semantics                                                                       */
/* ------------------------------------------------------------------------------------------------ */

#define FM(a, b) __fmul_rn((a), (b))
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))
#define FD(a, b) __fdiv_rn((a), (b))

/* MINPS / MAXPS: second operand when either is NaN */
__device__ __forceinline__ float min_x86(float a, float b) { return (a < b) ? a : b; }
__device__ __forceinline__ float max_x86(float a, float b) { return (a > b) ? a : b; }
__device__ __forceinline__ float clamp_x86(float x, float lo, float hi) { return min_x86(max_x86(x, lo), hi); }

/* CVTPS2DQ (round to nearest even; 0x80000000 when out of range / NaN) */
__device__ __forceinline__ int cvt_rne_x86(float x)
{
    int r = __float2int_rn(x);
    return (fabsf(x) < 2147483648.0f) ? r : INT_MIN;
}
__device__ __forceinline__ int cvt_trunc_x86(float x)
{
    int r = __float2int_rz(x);
    return (fabsf(x) < 2147483648.0f) ? r : INT_MIN;
}

/* RCPPS via the host-harvested table (simd.h:1217-1225; see host/pf_x86approx.c) */
__device__ __forceinline__ float rcp_x86(float x)
{
    const unsigned u = __float_as_uint(x), s = u & 0x80000000u, e = (u >> 23) & 255u, m = u & 0x7fffffu;
    const unsigned tv = __ldg(c_rcp_tab + (m >> c_rcp_shift));
    const int ex = (int)(tv >> 23) + 127 - (int)e;
    unsigned r = s | ((unsigned)ex << 23) | (tv & 0x7fffffu);
    if (ex <= 0) r = s;
    if (e == 0u) r = s | 0x7f800000u;
    if (e == 255u) r = m ? (u | 0x00400000u) : s;
    return __uint_as_float(r);
}

/* RSQRTPS (simd.h:1237-1245).  tab_smem: shared-window byte address of a copy of the table (k_raster_frag<Phong> keeps one:
 * three look-ups per lit fragment were that kernel's main long-scoreboard stall), or 0 for the global one. */
__device__ __forceinline__ float rsqrt_x86(float x, unsigned tab_smem = 0u)
{
    const unsigned u = __float_as_uint(x), s = u & 0x80000000u, e = (u >> 23) & 255u, m = u & 0x7fffffu;
    const unsigned odd = (e & 1u) ^ 1u;
    const int half = ((int)e - 127 - (int)odd) >> 1;       /* even by construction */
    const unsigned ti = (odd << c_rsq_bits) | (m >> c_rsq_shift);
    unsigned tv;
    if (tab_smem) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tv) : "r"(tab_smem + (ti << 2)));
    else tv = __ldg(c_rsq_tab + ti);
    unsigned r = ((unsigned)((int)(tv >> 23) - half) << 23) | (tv & 0x7fffffu);
    if (e == 255u) r = 0u;
    if (s) r = 0xffc00000u;
    if (e == 0u) r = s | 0x7f800000u;
    if (e == 255u && m) r = u | 0x00400000u;
    return __uint_as_float(r);
}

/* _mm256_log_ps (simd.h:183-252) */
__device__ __forceinline__ float log_cephes(float x)
{
    const bool invalid = (x <= 0.0f);
    x = max_x86(x, __uint_as_float(0x00800000u));
    int imm0 = (int)(__float_as_uint(x) >> 23);
    x = __uint_as_float((__float_as_uint(x) & ~0x7f800000u) | 0x3f000000u);
    imm0 -= 0x7f;
    float e = __int2float_rn(imm0);
    e = FA(e, 1.0f);
    const bool lt = (x < 0.707106781186547524f);
    float tmp = lt ? x : 0.0f;
    x = FS(x, 1.0f);
    e = FS(e, lt ? 1.0f : 0.0f);
    x = FA(x, tmp);
    const float z = FM(x, x);
    float y = 7.0376836292E-2f;
    y = FM(y, x); y = FA(y, -1.1514610310E-1f);
    y = FM(y, x); y = FA(y, 1.1676998740E-1f);
    y = FM(y, x); y = FA(y, -1.2420140846E-1f);
    y = FM(y, x); y = FA(y, 1.4249322787E-1f);
    y = FM(y, x); y = FA(y, -1.6668057665E-1f);
    y = FM(y, x); y = FA(y, 2.0000714765E-1f);
    y = FM(y, x); y = FA(y, -2.4999993993E-1f);
    y = FM(y, x); y = FA(y, 3.3333331174E-1f);
    y = FM(y, x);
    y = FM(y, z);
    tmp = FM(e, -2.12194440e-4f);
    y = FA(y, tmp);
    tmp = FM(z, 0.5f);
    y = FS(y, tmp);
    tmp = FM(e, 0.693359375f);
    x = FA(x, y);
    x = FA(x, tmp);
    return invalid ? __uint_as_float(0xffffffffu) : x;
}

/* _mm256_exp_ps (simd.h:254-304) */
__device__ __forceinline__ float exp_cephes(float x)
{
    x = min_x86(x, 88.3762626647949f);
    x = max_x86(x, -88.3762626647949f);
    float fx = FM(x, 1.44269504088896341f);
    fx = FA(fx, 0.5f);
    float tmp = floorf(fx);
    const float mask = (tmp > fx) ? 1.0f : 0.0f;
    fx = FS(tmp, mask);
    tmp = FM(fx, 0.693359375f);
    float z = FM(fx, -2.12194440e-4f);
    x = FS(x, tmp);
    x = FS(x, z);
    z = FM(x, x);
    float y = 1.9875691500E-4f;
    y = FM(y, x); y = FA(y, 1.3981999507E-3f);
    y = FM(y, x); y = FA(y, 8.3334519073E-3f);
    y = FM(y, x); y = FA(y, 4.1665795894E-2f);
    y = FM(y, x); y = FA(y, 1.6666665459E-1f);
    y = FM(y, x); y = FA(y, 5.0000001201E-1f);
    y = FM(y, z);
    y = FA(y, x);
    y = FA(y, 1.0f);
    int imm0 = cvt_trunc_x86(fx);
    imm0 = (int)((unsigned)imm0 + 0x7fu);
    return FM(y, __uint_as_float((unsigned)imm0 << 23));
}

/* ------------------------------------------------------------------------------------------------ */
/* device: colour arithmetic                                                                        */
/* ------------------------------------------------------------------------------------------------ */

#define CHN(c, i) ((int)(((c) >> (8 * (i))) & 255u))
#define INV255 (1.0f / 255.0f)

__device__ __forceinline__ unsigned pack4(int r, int g, int b, int a)   /* OR of shifted lanes, no masking (color.h:112-122) */
{
    return (unsigned)r | ((unsigned)g << 8) | ((unsigned)b << 16) | ((unsigned)a << 24);
}

__device__ __forceinline__ unsigned quant(float v)                      /* color.h:124-135 */
{
    /* clamp_x86 maps NaN to 0 (MAXPS returns its second operand), so the product is always in [0, 255] and
       CVTPS2DQ's out-of-range result cannot occur */
    return (unsigned)__float2int_rn(FM(clamp_x86(v, 0.0f, 1.0f), 255.0f));
}

/* (texel * frag) >> 8 per channel (blend.h:199-212) */
__device__ __forceinline__ unsigned mul_color(unsigned a, unsigned b)
{
    return pack4((CHN(a, 0) * CHN(b, 0)) >> 8, (CHN(a, 1) * CHN(b, 1)) >> 8,
                 (CHN(a, 2) * CHN(b, 2)) >> 8, (CHN(a, 3) * CHN(b, 3)) >> 8);
}

/* color.h:137-144 (Q7 fixed), for t in [0, 1] (the samplers clamp fx, fy; NaN becomes 0).  A and B are in [0, 1] and
 * representable, and rounding is monotone, so A + t*(B - A) stays between A and B: the clamp of pfiColorPackFromF
 * (quant) is the identity here and is left out.
 * The int <-> float conversions are done with exact float tricks instead of I2F / F2I: those run on the XU pipe, which the
 * bilinear filter (36 conversions per fragment) saturated (ncu: pipe_xu 76 % on the 4K bilinear scene).
 *   float(b), b a byte:  bits 0x4b000000 | b are the float 2^23 + b; subtracting 2^23 is exact;
 *   RNE(x), 0 <= x <= 255:  x + 1.5 * 2^23 is rounded to an integer by the addition itself (ulp 1), to nearest-even
 *                           like CVTPS2DQ; the integer sits in the low bits of the sum. */
#define LERP_MAGIC 12582912.0f          /* 1.5 * 2^23 */
__device__ __forceinline__ float byte_unit(unsigned c, int i)       /* float(byte i of c) * (1/255) */
{
    return FM(FS(__uint_as_float(__byte_perm(c, 0x4b000000u, 0x7440u | (unsigned)i)), 8388608.0f), INV255);
}
/* one channel: the sum x*255 + LERP_MAGIC, whose low byte is the quantised result */
__device__ __forceinline__ float lerp_chan(float A, float B, float t) { return FA(FM(FA(A, FM(t, FS(B, A))), 255.0f), LERP_MAGIC); }
__device__ __forceinline__ float magic_unit(float y) { return FM(FS(y, LERP_MAGIC), INV255); }      /* float(low bits of y) * (1/255) */

__device__ __forceinline__ unsigned color_lerp(unsigned a, unsigned b, float t)
{
    float y[4];
#pragma unroll
    for (int i = 0; i < 4; i++) y[i] = lerp_chan(byte_unit(a, i), byte_unit(b, i), t);
    return __byte_perm(__byte_perm(__float_as_uint(y[0]), __float_as_uint(y[1]), 0x0040), __byte_perm(__float_as_uint(y[2]), __float_as_uint(y[3]), 0x0040), 0x5410);
}

/* the three lerps of a bilinear tap set; the two intermediate colours stay in their float form */
__device__ __forceinline__ unsigned color_bilerp(unsigned c00, unsigned c10, unsigned c01, unsigned c11, float fx, float fy)
{
    float y[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float top = lerp_chan(byte_unit(c00, i), byte_unit(c10, i), fx), bot = lerp_chan(byte_unit(c01, i), byte_unit(c11, i), fx);
        y[i] = lerp_chan(magic_unit(top), magic_unit(bot), fy);
    }
    return __byte_perm(__byte_perm(__float_as_uint(y[0]), __float_as_uint(y[1]), 0x0040), __byte_perm(__float_as_uint(y[2]), __float_as_uint(y[3]), 0x0040), 0x5410);
}

__device__ __forceinline__ unsigned blend_px(int mode, unsigned s, unsigned d)   /* blend.h:137-274 */
{
    int o[4];
    switch (mode) {
    case 0:
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = (CHN(s, i) + CHN(d, i)) >> 1;
        break;
    case 1: {
        const int alpha = CHN(s, 3) + 1, inv = 256 - alpha;
#pragma unroll
        for (int i = 0; i < 3; i++) o[i] = (CHN(s, i) * alpha + CHN(d, i) * inv) >> 8;
        o[3] = (255 * alpha + CHN(d, 3) * inv) >> 8;
    } break;
    case 2:
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = min(CHN(s, i) + CHN(d, i), 255);
        break;
    case 3:                                 /* "subtractive" adds (Q6) and may carry into the next channel */
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = max(CHN(s, i) + CHN(d, i), 0);
        break;
    case 4:
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = (CHN(s, i) * CHN(d, i)) >> 8;
        break;
    case 5:
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = min(((CHN(d, i) * (255 - CHN(s, i))) >> 8) + CHN(s, i), 255);
        break;
    case 6:
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = max(CHN(s, i), CHN(d, i));
        break;
    default:
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = min(CHN(s, i), CHN(d, i));
        break;
    }
    return pack4(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ bool depth_pass(int func, float z, float zb)   /* depth.h:80-114; NOTEQUAL == EQUAL (Q5) */
{
    switch (func) {
    case 0: case 1: return z == zb;
    case 2: return z < zb;
    case 3: return z <= zb;
    case 4: return z > zb;
    default: return z >= zb;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* device: texturing (sampler.h:202-410)                                                            */
/* ------------------------------------------------------------------------------------------------ */

struct TexRegs {            /* texture state kept in registers while a warp stays in one state */
    const unsigned char *base; unsigned tw, th, total; float wm1, hm1; int fmt, wrap, filter;
};

__device__ __forceinline__ int tex_coord(int wrap, float t, float sm1)
{
    if (wrap == 0) {                    /* REPEAT: |RNE((t - trunc t) * (size-1))| */
        const float f = FM(FS(t, truncf(t)), sm1);
        return cvt_rne_x86(fabsf(f));   /* == |RNE(f)|: RNE is symmetric, 0x80000000 stays 0x80000000 */
    } else if (wrap == 1) {             /* MIRRORED_REPEAT */
        const float a = fabsf(t);
        float m = FS(a, FM(floorf(FD(a, 2.0f)), 2.0f));
        const float r = FS(1.0f, FS(m, 1.0f));
        if (m > 1.0f) m = r;
        return cvt_rne_x86(FA(FM(m, sm1), 0.5f));
    } else {                            /* CLAMP_TO_EDGE */
        return cvt_rne_x86(FA(FM(clamp_x86(t, 0.0f, 1.0f), sm1), 0.5f));
    }
}

/* The texel layouts beyond the four 8-bit ones: the reference's SIMD getters lane by lane (pf_pixfmt.h); memory past
   the end of the texture reads as zero bytes.  Out of line: the 8-bit paths keep their registers. */
__device__ __noinline__ unsigned tex_fetch_pix(const unsigned char *base, unsigned total, int off, int code)
{
    const unsigned char zeros[16] = { 0 };
    return (unsigned)off >= total ? pfx_tex_get(zeros, 0u, code) : pfx_tex_get(base, (unsigned)off, code);
}

/* PIX: this kernel also samples the layouts beyond the four 8-bit ones (the generic tile rasteriser and the row-ordered
   one do; batches with such a texture are routed to them, launch_pipeline) */
template <bool PIX>
__device__ __forceinline__ unsigned tex_fetch(const TexRegs &t, int x, int y)
{
    const int off = (int)((unsigned)y * t.tw + (unsigned)x);
    /* the reference reads out of bounds here (CLAMP/MIRROR round v*(h-1)+0.5 up to row h); defined as
       "memory after the texture reads as zero": RGBA 0, and alpha 255 for the 3-byte formats */
    if (PIX && t.fmt >= PFCU_TEX_PIX) return tex_fetch_pix(t.base, t.total, off, t.fmt - PFCU_TEX_PIX);
    if ((unsigned)off >= t.total) return (t.fmt >= PFCU_TEX_RGB8) ? 0xff000000u : 0u;
    if (t.fmt == PFCU_TEX_RGBA8) return __ldg((const unsigned *)t.base + off);
    if (t.fmt == PFCU_TEX_BGRA8) { const unsigned r = __ldg((const unsigned *)t.base + off); return __byte_perm(r, 0, 0x3012); }
    const unsigned char *p = t.base + 3 * (size_t)off;
    const unsigned b0 = __ldg(p), b1 = __ldg(p + 1), b2 = __ldg(p + 2);
    return (t.fmt == PFCU_TEX_RGB8) ? (b0 | (b1 << 8) | (b2 << 16) | 0xff000000u) : (b2 | (b1 << 8) | (b0 << 16) | 0xff000000u);
}

template <bool PIX>
__device__ __forceinline__ unsigned tex_sample_bilinear(const TexRegs &t, const DevState *st, float u, float v)
{
    const int x0 = tex_coord(t.wrap, u, t.wm1), y0 = tex_coord(t.wrap, v, t.hm1);
    const float4 k = __ldg(reinterpret_cast<const float4 *>(&st->tex_fw));       /* fw, fh, 1/fw, 1/fh */
    const float fw = k.x, fh = k.y, tx = k.z, ty = k.w;
    const int x1 = tex_coord(t.wrap, FA(u, tx), t.wm1), y1 = tex_coord(t.wrap, FA(v, ty), t.hm1);
    const float fx = clamp_x86(FS(FM(u, fw), __int2float_rn(x0)), 0.0f, 1.0f);
    const float fy = clamp_x86(FS(FM(v, fh), __int2float_rn(y0)), 0.0f, 1.0f);
    const unsigned c00 = tex_fetch<PIX>(t, x0, y0), c10 = tex_fetch<PIX>(t, x1, y0);
    const unsigned c01 = tex_fetch<PIX>(t, x0, y1), c11 = tex_fetch<PIX>(t, x1, y1);
    return color_bilerp(c00, c10, c01, c11, fx, fy);
}

template <bool PIX>
__device__ __forceinline__ unsigned tex_sample(const TexRegs &t, const DevState *st, float u, float v)
{
    const int x0 = tex_coord(t.wrap, u, t.wm1), y0 = tex_coord(t.wrap, v, t.hm1);
    if (t.filter == 0) return tex_fetch<PIX>(t, x0, y0);
    const float4 k = __ldg(reinterpret_cast<const float4 *>(&st->tex_fw));       /* fw, fh, 1/fw, 1/fh */
    const float fw = k.x, fh = k.y, tx = k.z, ty = k.w;
    const int x1 = tex_coord(t.wrap, FA(u, tx), t.wm1), y1 = tex_coord(t.wrap, FA(v, ty), t.hm1);
    const float fx = clamp_x86(FS(FM(u, fw), __int2float_rn(x0)), 0.0f, 1.0f);
    const float fy = clamp_x86(FS(FM(v, fh), __int2float_rn(y0)), 0.0f, 1.0f);
    const unsigned c00 = tex_fetch<PIX>(t, x0, y0), c10 = tex_fetch<PIX>(t, x1, y0);
    const unsigned c01 = tex_fetch<PIX>(t, x0, y1), c11 = tex_fetch<PIX>(t, x1, y1);
    return color_bilerp(c00, c10, c01, c11, fx, fy);
}

/* ------------------------------------------------------------------------------------------------ */
/* device: per-fragment Blinn-Phong (lighting.c:148-258)                                            */
/* ------------------------------------------------------------------------------------------------ */

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    return FA(FA(FM(ax, bx), FM(ay, by)), FM(az, bz));
}

__device__ __noinline__ unsigned phong(unsigned frag, const DevState *st, int face,
                                       float Px, float Py, float Pz, float Nx, float Ny, float Nz, unsigned rsq_smem = 0u)
{
    const DevMaterial *m = &st->material[face];
    float D[3], A[3], S[3], acc[3] = { 0.0f, 0.0f, 0.0f };
#pragma unroll
    for (int i = 0; i < 3; i++) {
        D[i] = FM(__int2float_rn(CHN(frag, i)), INV255);
        A[i] = FM(m->amb_f[i], D[i]);
        S[i] = m->spc_f[i];
    }
    float Vx = FS(st->view_pos[0], Px), Vy = FS(st->view_pos[1], Py), Vz = FS(st->view_pos[2], Pz);
    {
        const float inv = rsqrt_x86(max_x86(dot3(Vx, Vy, Vz, Vx, Vy, Vz), 1e-5f), rsq_smem);
        Vx = FM(Vx, inv); Vy = FM(Vy, inv); Vz = FM(Vz, inv);
    }
    const float shininess = m->shininess;
    for (unsigned li = 0; li < st->n_lights; li++) {
        const DevLight *l = &st->lights[li];
        float Lx = FS(l->pos[0], Px), Ly = FS(l->pos[1], Py), Lz = FS(l->pos[2], Pz);
        {
            const float inv = rsqrt_x86(max_x86(dot3(Lx, Ly, Lz, Lx, Ly, Lz), 1e-5f), rsq_smem);
            Lx = FM(Lx, inv); Ly = FM(Ly, inv); Lz = FM(Lz, inv);
        }
        const float diff = max_x86(dot3(Nx, Ny, Nz, Lx, Ly, Lz), 0.0f);
        float Hx = FA(Lx, Vx), Hy = FA(Ly, Vy), Hz = FA(Lz, Vz);
        {
            const float inv = rsqrt_x86(dot3(Hx, Hy, Hz, Hx, Hy, Hz), rsq_smem);      /* no epsilon here */
            Hx = FM(Hx, inv); Hy = FM(Hy, inv); Hz = FM(Hz, inv);
        }
        float spec = max_x86(dot3(Nx, Ny, Nz, Hx, Hy, Hz), 0.0f);
        spec = exp_cephes(FM(log_cephes(spec), shininess));                 /* pow(0) -> e^88 (Q9) */
        float inten = 1.0f; bool spot = false;
        if (l->inner < 3.14159265358979323846f) {
            spot = true;
            const float theta = dot3(Lx, Ly, Lz, FS(0.0f, l->dir[0]), FS(0.0f, l->dir[1]), FS(0.0f, l->dir[2]));
            inten = clamp_x86(FD(FS(theta, l->outer), FS(l->inner, l->outer)), 0.0f, 1.0f);
        }
        float att = 1.0f; bool atten = false;
        if (l->attl != 0.0f || l->attq != 0.0f) {
            atten = true;
            const float d0 = FS(l->pos[0], Px), d1 = FS(l->pos[1], Py);     /* y twice, z dropped (Q10) */
            const float dsq = FA(FM(d0, d0), FA(FM(d1, d1), FM(d1, d1)));
            const float dist = __fsqrt_rn(dsq);
            att = rcp_x86(FA(l->attc, FA(FM(l->attl, dist), FM(l->attq, dsq))));
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
            float amb = FM(l->amb_f[i], A[i]);
            float dif = FM(FM(l->dif_f[i], diff), D[i]);
            float spc = FM(FM(l->spc_f[i], spec), S[i]);
            if (spot) { dif = FM(dif, inten); spc = FM(spc, inten); }
            if (atten) { amb = FM(amb, att); dif = FM(dif, att); spc = FM(spc, att); }
            acc[i] = FA(acc[i], amb); acc[i] = FA(acc[i], dif); acc[i] = FA(acc[i], spc);
        }
    }
    return quant(acc[0]) | (quant(acc[1]) << 8) | (quant(acc[2]) << 16);   /* alpha = 0 (Q9) */
}
