/* TEST ONLY: host restatement of xdiv_huge_den (solidboolean_b200/csrc/sb_raytri.cuh) -- the division the classifier uses for
 * the subnormal ray parameter -- checked bit for bit against the host's IEEE division: random operands in the ranges the
 * classifier produces (denominator ~2^1000..2^1023, numerator 2^-60..2^60) and CONSTRUCTED ties (true quotient exactly on /
 * one ulp beside a midpoint of the subnormal grid), where a scaled division rounds twice.  Run by tests/test_hostsim.py.
 *     gcc -O2 -ffp-contract=off -o div_huge_den div_huge_den.c -lm && ./div_huge_den 5000000 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
static inline uint64_t bits(double x){uint64_t b;memcpy(&b,&x,8);return b;}
static inline double frombits(uint64_t b){double x;memcpy(&x,&b,8);return x;}
static int fast_ok=0, generic=0, fixes=0;
static double div_huge_den(double n, double d)
{
    const int en = (int)((bits(n) >> 52) & 0x7ff), ed = (int)((bits(d) >> 52) & 0x7ff);
    if (ed >= 1023 + 1000 && ed < 0x7ff && en >= 1023 - 200 && en <= 1023 + 200) {
        const double dS = d * 0x1p-600;          /* exact */
        const double q = n / dS;                 /* normal range, correctly rounded */
        if (fabs(q) >= 0x1p-470) {
            double s = q * 0x1p-600;             /* RN-even, possibly into the subnormal range */
            if (fabs(s) < 0x1p-1022) {
                const double back = s * 0x1p600; /* exact */
                const double diff = q - back;    /* exact */
                if (fabs(diff) == 0x1p-475) {    /* q sat on a midpoint of the subnormal grid */
                    const double r = fma(-q, dS, n); /* exact sign of n - q dS */
                    if (r != 0.0) {
                        const int above = (r > 0.0) == (dS > 0.0); /* true quotient > q */
                        const double other = back + 2.0 * diff;
                        const double lo = back < other ? back : other, hi = back < other ? other : back;
                        s = (above ? hi : lo) * 0x1p-600;
                        ++fixes;
                    }
                }
            }
            ++fast_ok;
            return s;
        }
    }
    ++generic;
    return n / d;
}
static uint64_t rng=88172645463325252ull;
static uint64_t xr(){rng^=rng<<13;rng^=rng>>7;rng^=rng<<17;return rng;}
int main(int argc,char**argv){
    long N = argc>1?atol(argv[1]):100000000L; long bad=0;
    for(long i=0;i<N;++i){
        /* d: exponent 1000..1023, random mantissa, random sign; n: exponent -60..60 */
        uint64_t md = xr() & 0xfffffffffffffull, mn = xr() & 0xfffffffffffffull;
        int ed = 1023 + 1000 + (int)(xr()%24), en = 1023 - 60 + (int)(xr()%121);
        if ((i & 7) == 0) { md &= 0xfffff00000000ull; }            /* short mantissas: more exact / tie cases */
        if ((i & 15) == 0) { mn &= 0xff00000000000ull; }
        double d = frombits(((uint64_t)ed<<52)|md|((xr()&1)<<63));
        double n = frombits(((uint64_t)en<<52)|mn|((xr()&1)<<63));
        double a = div_huge_den(n,d), b = n/d;
        if (bits(a)!=bits(b)) { if (bad<10) printf("MISMATCH n=%a d=%a got=%a want=%a\n",n,d,a,b); ++bad; }
    }
    /* constructed ties: pick subnormal midpoint m = (k + 0.5) * 2^-1074, d power-of-two-ish with short mantissa, n = m*d (exact when it fits) and neighbours */
    long ties=0;
    for(long i=0;i<N/5;++i){
        uint64_t k = (xr() % ((1ull<<40))) + 1;
        double m = ldexp((double)k + 0.5, -1074);   /* not representable as subnormal: need exact arithmetic: use long double? build n = (2k+1) * dm * 2^-1075 */
        (void)m;
        uint64_t dm = (xr() & 0x3ff) | 0x400;        /* 11-bit integer mantissa */
        int sh = 1000 + (int)(xr()%13);
        double d = ldexp((double)dm, sh - 10);       /* ~2^sh */
        /* n = (2k+1) * dm * 2^(sh-10-1075): exact if (2k+1)*dm < 2^53 */
        unsigned __int128 prod = (unsigned __int128)(2*k+1) * dm;
        if (prod >> 53) continue;
        double n = ldexp((double)(uint64_t)prod, sh - 10 - 1075);
        for (int delta=-1; delta<=1; ++delta){
            double nn = delta==0 ? n : nextafter(n, delta>0? INFINITY : -INFINITY);
            if ((xr()&1)) { nn = -nn; }
            double a = div_huge_den(nn,d), b = nn/d;
            if (bits(a)!=bits(b)) { if (bad<20) printf("TIE MISMATCH n=%a d=%a got=%a want=%a\n",nn,d,a,b); ++bad; }
            ++ties;
        }
    }
    printf("random %ld, tie-probes %ld, mismatches %ld, fast %d generic %d fixes %d\n",N,ties,bad,fast_ok,generic,fixes);
    return bad!=0;
}
