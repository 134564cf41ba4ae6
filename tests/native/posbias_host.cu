// Host-side harness (test infrastructure): runs the arithmetic of csrc/posbias.cuh -- the relative-position bias the opt-in
// attention kernels compute from positions -- on the CPU, so tests/test_posbias_host.py can pin it against the table formulation
// of the reference (aff.py:17-31,129-132,481-485) without a GPU.  Built by the test with `nvcc -shared` (host code only).
#include "../../autofocusformermod_b200/csrc/posbias.cuh"

using namespace clusten;

// q, k: [n][2] positions (x, y); w: [H][5], b: [H]; bias: [n][H]
extern "C" void posbias_host_bias(const float *q, const float *k, const float *w, const float *b, int n, int H, float *bias) {
    for (int h = 0; h < H; ++h) {
        const PosBiasW pw = {w[5 * h], w[5 * h + 1], w[5 * h + 2], w[5 * h + 3], w[5 * h + 4], b ? b[h] : 0.f};
        for (int i = 0; i < n; ++i)
            bias[(size_t)i * H + h] = pos_bias(pw, make_float2(q[2 * i], q[2 * i + 1]), make_float2(k[2 * i], k[2 * i + 1]));
    }
}

// grad[6] = sum_i ds[i] * [dx, dy, dist, dy/dist, dx/dist, 1]
extern "C" void posbias_host_grad(const float *q, const float *k, const float *ds, int n, float *grad) {
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < n; ++i) pos_bias_grad(acc, make_float2(q[2 * i], q[2 * i + 1]), make_float2(k[2 * i], k[2 * i + 1]), ds[i]);
    for (int c = 0; c < 6; ++c) grad[c] = acc[c];
}
