/* tools/verify_acosf.c — brute-force check that the float arc cosine the device mesh builder uses (csrc/mesh_build.cuh: acosfExact, the
 * fdlibm float algorithm evaluated without FMA contraction) returns the bits of the host libm the reference calls in
 * PseudoNormalVertex (Source/Meshing/Mesh.cpp:216-242), for EVERY float in [-1, 1]:
 *     gcc -O2 -ffp-contract=off -fopenmp tools/verify_acosf.c -o /tmp/verify_acosf -lm && /tmp/verify_acosf
 * glibc 2.39 (this image): 0 mismatches of 2 130 706 434. */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>
/* candidate A: fdlibm e_acosf.c (original Sun / Cygnus float port) */
static float acosf_a(float x)
{
    static const float one=1.0f, pi=3.1415925026e+00f, pio2_hi=1.5707962513e+00f, pio2_lo=7.5497894159e-08f,
        pS0=1.6666667163e-01f, pS1=-3.2556581497e-01f, pS2=2.0121252537e-01f, pS3=-4.0055535734e-02f, pS4=7.9153501429e-04f, pS5=3.4793309169e-05f,
        qS1=-2.4033949375e+00f, qS2=2.0209457874e+00f, qS3=-6.8828397989e-01f, qS4=7.7038154006e-02f;
    float z,p,q,r,w,s,c,df; int32_t hx,ix; memcpy(&hx,&x,4); ix=hx&0x7fffffff;
    if(ix==0x3f800000){ if(hx>0) return 0.0f; else return pi+(float)2.0*pio2_lo; }
    else if(ix>0x3f800000) return (x-x)/(x-x);
    if(ix<0x3f000000){ if(ix<=0x23000000) return pio2_hi+pio2_lo;
        z=x*x; p=z*(pS0+z*(pS1+z*(pS2+z*(pS3+z*(pS4+z*pS5))))); q=one+z*(qS1+z*(qS2+z*(qS3+z*qS4))); r=p/q; return pio2_hi-(x-(pio2_lo-x*r)); }
    else if(hx<0){ z=(one+x)*(float)0.5; p=z*(pS0+z*(pS1+z*(pS2+z*(pS3+z*(pS4+z*pS5))))); q=one+z*(qS1+z*(qS2+z*(qS3+z*qS4))); s=sqrtf(z); r=p/q; w=r*s-pio2_lo; return pi-(float)2.0*(s+w); }
    else { int32_t idf; z=(one-x)*(float)0.5; s=sqrtf(z); df=s; memcpy(&idf,&df,4); idf&=0xfffff000; memcpy(&df,&idf,4); c=(z-df*df)/(s+df);
        p=z*(pS0+z*(pS1+z*(pS2+z*(pS3+z*(pS4+z*pS5))))); q=one+z*(qS1+z*(qS2+z*(qS3+z*qS4))); r=p/q; w=r*s+c; return (float)2.0*(df+w); }
}
int main(){
    long bad=0, total=0; uint32_t firstbad=0;
    #pragma omp parallel for reduction(+:bad,total) schedule(static)
    for(int64_t b=0;b<=0x3f800000;b+=1){
        for(int sgn=0;sgn<2;++sgn){ uint32_t u=(uint32_t)b|(sgn?0x80000000u:0u); float x; memcpy(&x,&u,4);
            float r=acosf(x), m=acosf_a(x); uint32_t ur,um; memcpy(&ur,&r,4); memcpy(&um,&m,4);
            if(ur!=um){ if(!bad) firstbad=u; bad++; } total++; }
    }
    printf("candidate A: %ld mismatches of %ld (first %08x)\n", bad,total,firstbad);
    return 0; }
