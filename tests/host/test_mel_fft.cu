// Host-side check of the device FFT building blocks (flowtron_b200/csrc/mel_fft.cuh): runs the exact per-lane code of
// the fused mel kernel for all 32 "lanes" sequentially and compares with a direct O(N^2) double-precision DFT.
// Build + run: nvcc -std=c++17 -I flowtron_b200/csrc tests/host/test_mel_fft.cu -o /tmp/test_mel_fft && /tmp/test_mel_fft
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "mel_fft.cuh"

using namespace ft::melfft;

struct Src {
    const float* x;
    float2 operator()(int n) const { return make_float2(x[2 * n], x[2 * n + 1]); }
};

int main() {
    std::vector<float2> tw(1024);
    for (int j = 0; j < 1024; ++j) { const double a = -2.0 * M_PI * j / 1024.0; tw[j] = make_float2((float)cos(a), (float)sin(a)); }
    double worst = 0.0;
    for (int trial = 0; trial < 4; ++trial) {
        std::vector<float> x(NFFT);
        srand(17 + trial);
        for (auto& v : x) v = (float)rand() / RAND_MAX * 2.f - 1.f;
        std::vector<float> xr(XBUF), xi(XBUF);
        Src src{x.data()};
        for (int lane = 0; lane < 32; ++lane) for (int h = 0; h < 2; ++h) pass_a(lane + 32 * h, src, tw.data(), xr.data(), xi.data());
        cf regs[32][2][8];
        for (int lane = 0; lane < 32; ++lane) for (int h = 0; h < 2; ++h) pass_b_load(lane + 32 * h, xr.data(), xi.data(), regs[lane][h]);
        for (int lane = 0; lane < 32; ++lane) for (int h = 0; h < 2; ++h) pass_b_store(lane + 32 * h, regs[lane][h], tw.data(), xr.data(), xi.data());
        for (int lane = 0; lane < 32; ++lane) for (int h = 0; h < 2; ++h) pass_c_load(lane + 32 * h, xr.data(), xi.data(), regs[lane][h]);
        for (int lane = 0; lane < 32; ++lane) for (int h = 0; h < 2; ++h) pass_c_store(lane + 32 * h, regs[lane][h], xr.data(), xi.data());
        double scale = 0.0;
        std::vector<double> rr(NBINS), ri(NBINS);
        for (int k = 0; k < NBINS; ++k) {
            double sr = 0, si = 0;
            for (int n = 0; n < NFFT; ++n) { const double a = -2.0 * M_PI * n * k / NFFT; sr += x[n] * cos(a); si += x[n] * sin(a); }
            rr[k] = sr; ri[k] = si;
            scale = fmax(scale, hypot(sr, si));
        }
        for (int k = 0; k < NBINS; ++k) {
            cf X;
            if (k < NH) X = split_bin(k, xr.data(), xi.data(), tw.data());
            else X = {xr[0] - xi[0], 0.f};
            worst = fmax(worst, hypot(X.x - rr[k], X.y - ri[k]) / scale);
        }
    }
    printf("mel_fft host check: worst relative error %.3e\n", worst);
    return worst < 2e-6 ? 0 : 1;
}
