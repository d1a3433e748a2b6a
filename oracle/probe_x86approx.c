/* TEST INFRASTRUCTURE: characterises the host's RCPPS / RSQRTPS approximation
 * (reference uses them at src/internal/simd.h:1217-1245 via _mm256_rcp_ps/_mm256_rsqrt_ps).
 * Prints: smallest K such that the result mantissa/exponent depends only on the top K mantissa bits
 * (exhaustive over all 2^23 mantissas), exponent invariance, and an FNV hash of the table. */
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <immintrin.h>

static uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
static float rcp(float x){return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x)));}
static float rsq(float x){return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x)));}

int main(void){
    /* rcp: exponent 127 */
    for (int pass=0; pass<3; pass++){
        const char* name = pass==0?"rcp":(pass==1?"rsqrt_even(e=127)":"rsqrt_odd(e=128)");
        uint32_t ebits = (pass==2)?128u:127u;
        int K;
        for (K=8; K<=23; K++){
            int ok=1; uint32_t shift=23-K;
            for (uint32_t top=0; top<(1u<<K) && ok; top++){
                uint32_t base=(ebits<<23)|(top<<shift);
                float r0 = pass==0?rcp(u2f(base)):rsq(u2f(base));
                uint32_t step = shift>8 ? (1u<<(shift-8)) : 1; /* sample low bits, 256 samples + last */
                for (uint32_t lo=0; lo<(1u<<shift); lo+=step){
                    float r = pass==0?rcp(u2f(base|lo)):rsq(u2f(base|lo));
                    if (f2u(r)!=f2u(r0)){ok=0;break;}
                }
                if (ok && shift){ float r = pass==0?rcp(u2f(base|((1u<<shift)-1))):rsq(u2f(base|((1u<<shift)-1))); if (f2u(r)!=f2u(r0)) ok=0; }
            }
            if (ok) break;
        }
        /* exhaustive confirm for that K */
        int exhaustive=1; uint32_t shift=23-K;
        uint64_t h=1469598103934665603ull;
        for (uint32_t m=0; m<(1u<<23); m++){
            uint32_t base=(ebits<<23)|((m>>shift)<<shift);
            float r0 = pass==0?rcp(u2f(base)):rsq(u2f(base));
            float r = pass==0?rcp(u2f((ebits<<23)|m)):rsq(u2f((ebits<<23)|m));
            if (f2u(r)!=f2u(r0)) {exhaustive=0;}
            if ((m & ((1u<<shift)-1))==0){ h^=f2u(r0); h*=1099511628211ull; }
        }
        /* exponent invariance: compare mantissa of result across exponents */
        int expinv=1;
        for (int e=2; e<=252 && expinv; e+= (pass==0?1:2)){
            uint32_t ee = (pass==2)? (uint32_t)(e|1) : (pass==1? (uint32_t)((e&~1)|1) : (uint32_t)e);
            if (pass==1) ee = (uint32_t)(e|1);      /* odd biased exponent == even unbiased (127) */
            if (pass==2) ee = (uint32_t)(e&~1); if (pass==2 && ee<2) continue;
            for (uint32_t top=0; top<(1u<<K); top+=37){
                uint32_t m=top<<shift;
                float a = pass==0?rcp(u2f((ebits<<23)|m)):rsq(u2f((ebits<<23)|m));
                float b = pass==0?rcp(u2f((ee<<23)|m)):rsq(u2f((ee<<23)|m));
                if ((f2u(a)&0x7fffff)!=(f2u(b)&0x7fffff)) {expinv=0; printf("  expinv fail e=%u top=%u a=%08x b=%08x\n",ee,top,f2u(a),f2u(b)); break;}
            }
        }
        printf("%s: K=%d exhaustive=%d expinv=%d tablehash=%016llx  f(1.0)=%08x f(1.5)=%08x\n", name, K, exhaustive, expinv,
               (unsigned long long)h, f2u(pass==0?rcp(u2f(ebits<<23)):rsq(u2f(ebits<<23))), f2u(pass==0?rcp(u2f((ebits<<23)|0x400000)):rsq(u2f((ebits<<23)|0x400000))));
    }
    printf("rcp(0)=%08x rcp(-0)=%08x rcp(inf)=%08x rcp(denorm)=%08x rcp(2^126)=%08x rcp(1.5*2^126)=%08x rcp(2^127)=%08x rcp(nan)=%08x rcp(2^-126)=%08x\n",
        f2u(rcp(0.f)), f2u(rcp(-0.f)), f2u(rcp(u2f(0x7f800000))), f2u(rcp(u2f(0x00000100))), f2u(rcp(u2f(0x7e800000))), f2u(rcp(u2f(0x7ec00000))), f2u(rcp(u2f(0x7f000000))), f2u(rcp(u2f(0x7fc00000))), f2u(rcp(u2f(0x00800000))));
    printf("rsq(0)=%08x rsq(-1)=%08x rsq(inf)=%08x rsq(denorm)=%08x rsq(nan)=%08x rsq(2^-126)=%08x rsq(max)=%08x\n",
        f2u(rsq(0.f)), f2u(rsq(-1.f)), f2u(rsq(u2f(0x7f800000))), f2u(rsq(u2f(0x00000100))), f2u(rsq(u2f(0x7fc00000))), f2u(rsq(u2f(0x00800000))), f2u(rsq(u2f(0x7f7fffff))));
    return 0;
}
