// Checks the host instantiation of dashing2_b200/csrc/devlog.cuh::ref_log bit-for-bit against the
// host libm log() (the arithmetic the reference uses).  Built and run by tests/test_devlog.py.
#include "../../dashing2_b200/csrc/devlog.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
int main(int argc, char **argv) {
    const long n = argc > 1 ? atol(argv[1]) : 5000000;
    std::mt19937_64 rng(99);
    long bad = 0;
    auto check = [&](double x) {
        double a = std::log(x), b = d2g::ref_log(x);
        if (memcmp(&a, &b, 8) && !(std::isnan(a) && std::isnan(b))) { if (bad++ < 10) printf("MISMATCH x=%a libm=%a port=%a\n", x, a, b); }
    };
    for (long i = 0; i < n; ++i) {
        uint64_t r = rng();
        check((double)r * 0x1p-64);                       // the SetSketch argument: rv * 2^-64
        check((double)(r >> 12) * 0x1p-52);               // the BagMinHash argument: (wy >> 12) * 2^-52
        double u; uint64_t bits = r & 0x7fefffffffffffffULL; memcpy(&u, &bits, 8); check(u); // any positive finite
        check(1.0 + ((double)(int64_t)(r >> 11) - 0x1p52) * 0x1p-56);  // near 1 (both sides)
        float s = (float)((r >> 40) + 1) / 16777216.f; check(2. * s / (1. + s)); // Mash distance argument
    }
    double specials[] = {0., -0., 1., 0.5, 2., INFINITY, -1., NAN, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 0.9375, 1.0647};
    for (double x : specials) check(x);
    printf(bad ? "FAIL %ld\n" : "OK\n", bad);
    return bad != 0;
}
