// Checks dashing2_b200/csrc/xf80.h (software x87 arithmetic used by the compare kernel's finalisation)
// bit-for-bit against the CPU's native long double.  Built and run by tests/test_xf80.py.
#include "../../dashing2_b200/csrc/xf80.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

static bool same(long double a, long double b) {
    if (std::isnan(a) && std::isnan(b)) return true;
    return memcmp(&a, &b, 10) == 0;
}
int main(int argc, char **argv) {
    const long n = argc > 1 ? atol(argv[1]) : 2000000;
    std::mt19937_64 rng(12345);
    auto rnd_ld = [&](int kind) -> long double {
        uint64_t r = rng();
        switch (kind % 8) {
            case 0: return (long double)(r >> (rng() % 64));                       // integers
            case 1: return std::ldexp((long double)(r | (1ULL << 63)), int(rng() % 200) - 163); // wide range
            case 2: return (double)(r >> 11) * 0x1p-53;                             // doubles in [0,1)
            case 3: return (long double)(r % 8193) * (1.L / (long double)(1 + rng() % 8192)); // count * invdenom
            case 4: return 1.L + std::ldexp((long double)(r >> 1), -64 - int(rng() % 4)); // near 1
            case 5: return (rng() & 1) ? 2.L : 1.L;
            case 6: return -(long double)(double)(r >> 20);
            default: return std::ldexp((long double)r, -int(rng() % 130));
        }
    };
    long bad = 0;
    for (long i = 0; i < n && bad < 10; ++i) {
        long double a = rnd_ld(int(rng())), b = rnd_ld(int(rng()));
        if (rng() % 16 == 0) b = a + std::ldexp(a, -int(rng() % 70)); // near-cancellation
        if (rng() % 64 == 0) b = a;
        xf::f80 xa = xf::from_long_double(a), xb = xf::from_long_double(b);
        struct { const char *nm; long double want; xf::f80 got; } t[] = {
            {"add", a + b, xf::add(xa, xb)}, {"sub", a - b, xf::sub(xa, xb)},
            {"mul", a * b, xf::mul(xa, xb)}, {"div", a / b, xf::div(xa, xb)},
        };
        for (auto &c : t) {
            long double g = xf::to_long_double(c.got);
            if (!same(g, c.want)) { ++bad; printf("MISMATCH %s a=%La b=%La want=%La got=%La\n", c.nm, a, b, c.want, g); }
        }
        float wf = (float)a; float gf = xf::to_float(xa);
        if (memcmp(&wf, &gf, 4) && !(std::isnan(wf) && std::isnan(gf))) { ++bad; printf("MISMATCH to_float a=%La want=%a got=%a\n", a, wf, gf); }
        double wd = (double)a; double gd = xf::to_double(xa);
        if (memcmp(&wd, &gd, 8) && !(std::isnan(wd) && std::isnan(gd))) { ++bad; printf("MISMATCH to_double a=%La want=%a got=%a\n", a, wd, gd); }
        double d = (double)b; if (!same((long double)d, xf::to_long_double(xf::from_double(d)))) { ++bad; printf("MISMATCH from_double %a\n", d); }
        float f = (float)b; if (!same((long double)f, xf::to_long_double(xf::from_float(f)))) { ++bad; printf("MISMATCH from_float %a\n", f); }
        if ((a < b) != xf::lt(xa, xb) || (a <= b) != xf::le(xa, xb)) { ++bad; printf("MISMATCH cmp a=%La b=%La\n", a, b); }
        uint64_t u = rng() >> (rng() % 64);
        if (!same((long double)u, xf::to_long_double(xf::from_u64(u)))) { ++bad; printf("MISMATCH from_u64\n"); }
    }
    // small-float narrowing (subnormal floats) and overflow
    for (int e = -160; e <= 140 && bad < 10; ++e) for (int j = 0; j < 2000; ++j) {
        long double a = std::ldexp((long double)(rng() | (1ULL << 63)), e - 63);
        float wf = (float)a, gf = xf::to_float(xf::from_long_double(a));
        if (memcmp(&wf, &gf, 4)) { ++bad; printf("MISMATCH narrow e=%d a=%La want=%a got=%a\n", e, a, wf, gf); }
    }
    printf(bad ? "FAIL\n" : "OK\n");
    return bad != 0;
}
