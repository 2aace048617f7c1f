#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
static uint64_t s = 88172645463325252ULL;
static inline uint64_t rnd(void){ s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main(void){
    uint64_t bad = 0, bad1 = 0, n = 0;
    for (int k = 1; k <= 4096; k++) {
        const double b = (double)k, r = 1.0 / b;
        for (int t = 0; t < 150000; t++) {
            uint64_t u = rnd();
            double d;
            switch (t & 3) {
            case 0: d = (double)(int64_t)(u >> 11) * (1.0/512) - 4e12; break;                 /* typical deltas */
            case 1: { union { uint64_t i; double f; } v; v.i = (u & 0x800fffffffffffffULL) | ((uint64_t)(1023 - 60 + (u >> 52) % 160) << 52); d = v.f; } break; /* wide exponent range */
            case 2: d = (double)(u >> 1) - (double)(rnd() >> 1); break;                         /* u64 differences */
            default: d = (double)(int64_t)(u % 2000001) - 1000000.0; break;
            }
            const double want = d / b;
            double q0 = d * r;
            double r0 = fma(-b, q0, d);
            double q1 = fma(r0, r, q0);
            double r1 = fma(-b, q1, d);
            double q2 = fma(r1, r, q1);
            n++;
            if (q1 != want) bad1++;
            if (q2 != want) { if (bad < 5) printf("MISMATCH k=%d d=%a want=%a got=%a\n", k, d, want, q2); bad++; }
        }
    }
    /* quotients as close to a rounding boundary as a 53-bit dividend allows: d = RN(k * (2Q+1)) +- 0..2 ulp, scaled */
    uint64_t nm = 0, badm1 = 0, badm2 = 0;
    for (int k = 1; k <= 4096; k++) {
        const double b = (double)k, r = 1.0 / b;
        for (int t = 0; t < 60000; t++) {
            const uint64_t Q = (rnd() >> 11) | (1ULL << 52);
            unsigned __int128 tt = (unsigned __int128)k * (2 * (unsigned __int128)Q + 1);
            int sh = 0;
            while (tt >> 53) { tt >>= 1; sh++; }
            for (int dl = -2; dl <= 2; dl++) {
                double d = ldexp((double)(uint64_t)(tt + dl), sh - 1 - (int)(rnd() % 40));
                if (rnd() & 1) d = -d;
                const double want = d / b;
                const double q0 = d * r;
                const double q1 = fma(fma(-b, q0, d), r, q0);
                const double q2 = fma(fma(-b, q1, d), r, q1);
                nm++;
                if (q1 != want) badm1++;
                if (q2 != want) badm2++;
            }
        }
    }
    printf("near-midpoint trials %llu, one-correction mismatches %llu, two-correction mismatches %llu\n",
           (unsigned long long)nm, (unsigned long long)badm1, (unsigned long long)badm2);
    bad += badm1 + bad1;
    printf("trials %llu, one-correction mismatches %llu, two-correction mismatches %llu\n", (unsigned long long)n, (unsigned long long)bad1, (unsigned long long)bad);
    return bad != 0;
}
