// development (host, no GPU): how many sweeps the one-sided Jacobi SVD of aep_math.cuh needs on the deformation gradients of a flowing
// state, and the per-sweep max cos^2 histogram (profiles/r2_svd_sweeps.txt).  Motivation of the SVD-free sand path.
#include <cmath>
#include <cstring>
#include <cstdio>
#include <random>
#define AEP_HOST_MATH_TEST
#define __device__
#define __host__
#define __forceinline__ inline
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline float __fdividef(float a, float b) { return a / b; }
using std::fmaf;
#include "../anisotropicelastoplasticity_b200/csrc/aep_math.cuh"   // build: g++ -O2 -o /tmp/svd_sweeps tools/svd_sweeps.cpp
using namespace aep;
// instrumented copy of svd3: records per-sweep max cos^2, and result quality for forced sweep counts
static void sweeps_trace(const float* F, float* mx_out, int nsw) {
    float A[3][3], W[3][3];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) { A[c][r] = F[3*r+c]; W[c][r] = r==c; }
    for (int s = 0; s < nsw; ++s) {
        float mx = jacobi_rot<0,1>(A, W); mx = fmaxf(mx, jacobi_rot<0,2>(A, W)); mx = fmaxf(mx, jacobi_rot<1,2>(A, W));
        mx_out[s] = mx;
    }
}
int main() {
    std::mt19937 rng(1); std::normal_distribution<float> N(0, 1);
    const char* names[] = {"I+5e-4 shear (flow, projected each step)", "I+2e-3 noise (perturb)", "I+1e-2 noise", "I+1e-1 noise", "snow-like 0.98 + 5e-3 noise"};
    for (int kind = 0; kind < 5; ++kind) {
        double hist[8][16] = {{0}}; int n = 200000; long cnt[9] = {0};
        for (int t = 0; t < n; ++t) {
            float F[9] = {1,0,0, 0,1,0, 0,0,1};
            if (kind == 0) { F[2] += 5e-4f * (1 + 0.1f*N(rng)); F[5] += 2e-4f * N(rng); for (int i = 0; i < 9; ++i) F[i] += 2e-5f * N(rng); }
            else { float e = kind == 1 ? 2e-3f : kind == 2 ? 1e-2f : kind == 3 ? 1e-1f : 5e-3f; for (int i = 0; i < 9; ++i) F[i] += e * N(rng); if (kind == 4) { F[0] -= 0.02f; F[4] -= 0.02f; F[8] -= 0.02f; } }
            float mx[8]; sweeps_trace(F, mx, 8);
            int need = 8; for (int s = 0; s < 8; ++s) if (mx[s] < 1e-12f) { need = s + 1; break; }
            cnt[need]++;
            for (int s = 0; s < 8; ++s) { int b = mx[s] <= 0 ? 15 : (int)fmin(15.0, fmax(0.0, -log10((double)mx[s]))); hist[s][b]++; }
        }
        printf("== %s\n sweeps executed by current rule:", names[kind]); for (int s = 1; s <= 8; ++s) printf(" %d:%.3f", s, cnt[s] / (double)n); printf("\n");
        for (int s = 0; s < 5; ++s) { printf("  sweep %d max cos^2 (before rot) -log10 hist:", s); for (int b = 0; b < 16; ++b) printf(" %.3f", hist[s][b] / n); printf("\n"); }
    }
}
