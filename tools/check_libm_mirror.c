// Exhaustive host check of csrc/libm_mirror.h against the host libm (glibc) log1pf / expm1f.
// usage: check_libm_mirror [stride]   (stride 1 = every float bit pattern; default 1)
// build: gcc -O2 -ffp-contract=off -fopenmp tools/check_libm_mirror.c -o /tmp/check_libm_mirror -lm
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "../image-preprocessing-pipeline_b200/csrc/libm_mirror.h"

int main(int argc, char **argv)
{
    unsigned long stride = argc > 1 ? strtoul(argv[1], 0, 10) : 1;
    unsigned long bad_l = 0, bad_e = 0, n = 0;
    unsigned first_l = 0, first_e = 0;
#pragma omp parallel for reduction(+ : bad_l, bad_e, n) schedule(static)
    for (unsigned long b = 0; b < 0x100000000ul; b += stride) {
        uint32_t u = (uint32_t)b;
        float x; memcpy(&x, &u, 4);
        if (isnan(x)) continue;
        float a = log1pf(x), m = b2s_log1pf(x);
        uint32_t ua, um; memcpy(&ua, &a, 4); memcpy(&um, &m, 4);
        if (ua != um && !(isnan(a) && isnan(m))) { bad_l++; if (!first_l) first_l = u; }
        a = expm1f(x); m = b2s_expm1f(x);
        memcpy(&ua, &a, 4); memcpy(&um, &m, 4);
        if (ua != um && !(isnan(a) && isnan(m))) { bad_e++; if (!first_e) first_e = u; }
        n++;
    }
    printf("{\"checked\": %lu, \"log1pf_mismatch\": %lu, \"expm1f_mismatch\": %lu, \"first_l\": \"0x%08x\", \"first_e\": \"0x%08x\"}\n",
           n, bad_l, bad_e, first_l, first_e);
    return (bad_l || bad_e) ? 1 : 0;
}
