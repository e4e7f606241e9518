// Host-side check of rings.cu's decision intervals: for every threshold / decision kind the float intervals found by
// bisection reproduce the reference's test on the double angle (acosf * 57.29578, folded) on 6e6 floats incl. the interval
// ends, their neighbours, NaN and |x| > 1.  Built and run by tests/test_host_logic_cpu.py (no GPU needed: host code only).
#include "rings.cu"
#include <cstdio>
#include <random>
int main() {
    std::mt19937 g(5);
    std::uniform_real_distribution<float> U(-1.2f, 1.2f);
    long bad = 0, n = 0;
    const float ths[] = {0.f, 5.f, 30.f, 45.3f, 60.f, 89.9999f, 90.f, 120.f, -3.f, NAN};
    for (float th : ths) {
        const double t = (double)th;
        for (int variant = 0; variant < 3; ++variant) {
            mkb::RingSet s;
            if (variant == 0) s = mkb::ring_decision_set([=](double a) { return a <= t; }, false);
            else if (variant == 1) s = mkb::ring_decision_set([=](double a) { return a >= t; }, false);
            else s = mkb::ring_decision_set([=](double a) { return a >= t; }, true);
            auto truth = [&](float x) {
                double a = (double)acosf(x) * 57.29578;
                if (a > 90.0) a = 180.0 - a;
                if (variant == 2) a = 90.0 - a;
                return variant == 0 ? a <= t : a >= t;
            };
            auto in = [&](float x) { return (x >= s.lo[0] && x <= s.hi[0]) || (x >= s.lo[1] && x <= s.hi[1]); };
            for (int i = 0; i < 200000; ++i) {
                float x = U(g);
                if (i % 7 == 0) x = std::nextafterf(s.hi[i % 2], (i % 3) ? 2.f : -2.f);
                if (i % 11 == 0) x = std::nextafterf(s.lo[i % 2], (i % 3) ? 2.f : -2.f);
                if (i % 13 == 0) x = s.lo[i % 2];
                if (i % 17 == 0) x = s.hi[i % 2];
                if (i % 1001 == 0) x = NAN;
                ++n;
                if (truth(x) != in(x)) { if (bad < 10) printf("th %g var %d x %.9g truth %d in %d  set [%g,%g] [%g,%g]\n", th, variant, x, truth(x), in(x), s.lo[0], s.hi[0], s.lo[1], s.hi[1]); ++bad; }
            }
        }
    }
    printf("checked %ld, mismatches %ld\n", n, bad);
    return bad != 0;
}
