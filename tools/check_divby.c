/* Brute-force check of div_by() (csrc/lockstep.cuh): a / d from the correctly rounded reciprocal y = 1/d and two
 * Newton-Markstein corrections must be BITWISE the IEEE quotient for operands in the safe exponent range.
 *
 *   q0 = a*y;  r0 = fma(-d, q0, a);  q1 = fma(r0, y, q0);  r1 = fma(-d, q1, a);  q2 = fma(r1, y, q1)
 *
 * q1 is a faithful quotient (relative error 2^-104 before its rounding), so r1 is exact and q2 = RN(a/d) by Markstein's theorem
 * (y is the correctly rounded reciprocal).  The test draws mantissas that stress the theorem's edge (all-ones divisors, powers of
 * two, quotients next to rounding boundaries) besides uniformly random ones.
 *
 *   gcc -O2 -mfma -ffp-contract=off tools/check_divby.c -o /tmp/check_divby -lm && /tmp/check_divby [millions]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double mk(uint64_t sign, int e, uint64_t mant) {
    uint64_t b = (sign << 63) | ((uint64_t)(e + 1023) << 52) | (mant & 0xFFFFFFFFFFFFFull);
    double v;
    memcpy(&v, &b, 8);
    return v;
}
static double div_by(double a, double d, double y) {
    const double q0 = a * y;
    const double r0 = fma(-d, q0, a);
    const double q1 = fma(r0, y, q0);
    const double r1 = fma(-d, q1, a);
    const double q2 = fma(r1, y, q1);
    return (a == 0.0) ? q0 : q2;
}
static uint64_t special_mant(void) {
    switch (rnd() % 8) {
        case 0: return 0xFFFFFFFFFFFFFull;                 /* all ones */
        case 1: return 0;                                  /* power of two */
        case 2: return 0xFFFFFFFFFFFFFull - (rnd() % 64);  /* just below 2 */
        case 3: return rnd() % 64;                         /* just above 1 */
        case 4: return (rnd() % 4096) << 40;               /* few significant bits */
        case 5: return 0x8000000000000ull + (rnd() % 16) - 8;
        default: return rnd();
    }
}

int main(int argc, char** argv) {
    const long long n = (argc > 1 ? atoll(argv[1]) : 200) * 1000000ll;
    long long bad = 0, zero = 0;
    for (long long i = 0; i < n; i++) {
        const int ea = (int)(rnd() % 601) - 300, ed = (int)(rnd() % 601) - 300;   /* the safe range of the device helper */
        double a = mk(rnd() & 1, ea, (i & 1) ? special_mant() : rnd());
        const double d = mk(rnd() & 1, ed, (i & 2) ? special_mant() : rnd());
        if (i % 1000 == 0) { a = (rnd() & 1) ? 0.0 : -0.0; zero++; }
        const double y = 1.0 / d;
        const double q = div_by(a, d, y), t = a / d;
        if (memcmp(&q, &t, 8) != 0) {
            if (bad < 10) printf("MISMATCH a=%a d=%a got %a want %a\n", a, d, q, t);
            bad++;
        }
    }
    printf("%lld cases (%lld signed zeros), %lld mismatches\n", n, zero, bad);
    return bad != 0;
}
