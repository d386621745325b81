/* sincosf_exhaustive.c -- compares oko_sincosf (oracle/ok_oracle.c) with the box's libm sincosf on ALL
 * 2^32 binary32 inputs (TEST INFRASTRUCTURE).  ~2 minutes on one core.
 *
 *   gcc -O2 -mfma -ffp-contract=off -fopenmp -o /tmp/sincosf_exhaustive oracle/sincosf_exhaustive.c oracle/ok_oracle.c -lm
 *   /tmp/sincosf_exhaustive [stride]
 *
 * Recorded result (glibc 2.39-0ubuntu8.5, Xeon with FMA, 2026-10-18): total=4294967296 bad=0.
 * With -DOKO_SINCOS_NOFMA (glibc's SSE2 ifunc variant) the same run reports bad=34. */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ok_oracle.h"

int main(int argc, char **argv)
{
    uint64_t bad = 0, total = 0;
    uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;
    for (uint64_t u = 0; u < (1ULL << 32); u += stride) {
        uint32_t ui = (uint32_t)u, a, b, c, d;
        float    y, s1, c1, s2, c2;
        memcpy(&y, &ui, 4);
        oko_sincosf(y, &s1, &c1);
        sincosf(y, &s2, &c2);
        memcpy(&a, &s1, 4), memcpy(&b, &s2, 4), memcpy(&c, &c1, 4), memcpy(&d, &c2, 4);
        total++;
        if ((a != b || c != d) && !(s1 != s1 && s2 != s2)) {
            if (bad < 10)
                printf("mismatch y=%a oracle=(%a,%a) libm=(%a,%a)\n", y, s1, c1, s2, c2);
            bad++;
        }
    }
    printf("total=%llu bad=%llu\n", (unsigned long long)total, (unsigned long long)bad);
    return bad != 0;
}
