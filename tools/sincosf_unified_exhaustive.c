/* sincosf_unified_exhaustive.c -- the single-path formulation of the device sincosf (openkitchen_b200/csrc/ok_math.cuh:
 * one branch for every |y| < 120, signs applied to the binary32 results) restated in C and compared with the box's
 * libm sincosf on ALL 2^32 binary32 inputs (TEST INFRASTRUCTURE).
 *
 *   gcc -O2 -mfma -ffp-contract=off -fopenmp -o /tmp/scu tools/sincosf_unified_exhaustive.c -lm && /tmp/scu
 *
 * Recorded result (glibc 2.39, 8 threads, 13 s): bad=0. */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define C0 0x1p0
#define C1 -0x1.ffffffd0c621cp-2
#define C2 0x1.55553e1068f19p-5
#define C3 -0x1.6c087e89a359dp-10
#define C4 0x1.99343027bf8c3p-16
#define S1 -0x1.555545995a603p-3
#define S2 0x1.1107605230bc4p-7
#define S3 -0x1.994eb3774cf24p-13
static const uint32_t kInvPio4[24] = {
    0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
    0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
    0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
/* x: reduced argument WITHOUT the quadrant sign; q: quadrant used for sign/flip; n: quadrant used for the swap */
static inline void core(double x, int n, int q, float *sp, float *cp)
{
    const double x2 = x * x;
    const double x4 = x2 * x2;
    const double x3 = x2 * x;
    const double c2 = fma(x2, C4, C3);
    const double s1 = fma(x2, S3, S2);
    const double c1 = fma(x2, C1, C0);
    const double x5 = x3 * x2;
    const double x6 = x4 * x2;
    const double s  = fma(x3, S1, x);
    const double c  = fma(x4, C2, c1);
    uint32_t sv = f2u((float)fma(x5, s1, s));
    uint32_t cv = f2u((float)fma(x6, c2, c));
    sv ^= ((uint32_t)(q + 1) & 2u) << 30; /* sign table {1,-1,-1,1}[q & 3] */
    cv ^= ((uint32_t)q & 2u) << 30;       /* negated-cosine table for q & 2 */
    if (n & 1) { *sp = u2f(cv); *cp = u2f(sv); } else { *sp = u2f(sv); *cp = u2f(cv); }
}
static void new_sincosf(float y, float *sp, float *cp)
{
    const uint32_t top = (f2u(y) >> 20) & 0x7ffu;
    double x = (double)y;
    if (top < 0x42fu) {
        const double r = x * 0x1.45F306DC9C883p+23;
        const int    n = ((int32_t)r + 0x800000) >> 24;
        x = fma(-(double)n, 0x1.921FB54442D18p0, x);
        float s_, c_;
        core(x, n, n, &s_, &c_);
        if (top < 0x398u) { s_ = y; c_ = 1.0f; }
        *sp = s_; *cp = c_;
    } else if (top < 0x7f8u) {
        uint32_t xi = f2u(y);
        const int sign = (int)(xi >> 31);
        const uint32_t *arr = &kInvPio4[(xi >> 26) & 15];
        const int shift = (xi >> 23) & 7;
        xi = (xi & 0xffffffu) | 0x800000u;
        xi <<= shift;
        uint64_t res0 = (uint64_t)(uint32_t)(xi * arr[0]);
        const uint64_t res1 = (uint64_t)xi * arr[4];
        const uint64_t res2 = (uint64_t)xi * arr[8];
        res0 = (res2 >> 32) | (res0 << 32);
        res0 += res1;
        const uint64_t nn = (res0 + (1ULL << 61)) >> 62;
        res0 -= nn << 62;
        x = (double)(int64_t)res0 * 0x1.921FB54442D18p-62;
        const int n = (int)nn;
        core(x, n, n + sign, sp, cp);
    } else {
        *sp = *cp = y - y;
    }
}
int main(int argc, char **argv)
{
    uint64_t bad = 0;
    #pragma omp parallel for reduction(+:bad) schedule(static, 1<<20)
    for (uint64_t u = 0; u < (1ULL << 32); u++) {
        float y = u2f((uint32_t)u), s1, c1, s2, c2;
        new_sincosf(y, &s1, &c1);
        sincosf(y, &s2, &c2);
        if ((f2u(s1) != f2u(s2) || f2u(c1) != f2u(c2)) && !(s1 != s1 && s2 != s2)) {
            if (bad < 5) printf("mismatch y=%a new=(%a,%a) libm=(%a,%a)\n", y, s1, c1, s2, c2);
            bad++;
        }
    }
    printf("bad=%llu\n", (unsigned long long)bad);
    return bad != 0;
}
