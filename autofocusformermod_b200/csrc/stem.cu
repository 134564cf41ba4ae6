// First half of the AFF stem in one pass (sm_100a, fp32 inference):
//     y = GELU( BatchNorm_eval( Conv2d(IC -> OC, 3x3, stride 2, padding 1)(x) ) )
// -- `self.act1(self.bn(self.proj1(x)))` of PatchEmbed.forward (mask2former/modeling/backbone/aff.py:527-529,549).
//
// With IC = 3 the convolution is 27 multiply-adds per output value: a memory-bound stencil, not a GEMM.  cuDNN answers it with an
// implicit-GEMM kernel, ATen adds the bias, cuDNN normalises, ATen applies GELU -- four passes over the [B, OC, H/2, W/2] map
// (0.24 ms of the 4.9 ms AFF-Mini forward at 512^2, batch 16; the map is 67 MB).  Here one thread owns one output pixel and all OC
// channels of it: 27 input loads (L1 serves the overlap between neighbouring pixels), the weights as [tap][OC] in shared memory
// read as broadcast LDS.128, bias / statistics / GELU in registers, one coalesced store per channel.  HBM traffic: x once, y once.
#include <algorithm>

#include "common.cuh"

namespace clusten {

constexpr int STEM_TX = 64, STEM_TY = 4;                 // output pixels of a CTA: 64 along x, 4 rows

template <int OC, int IC>
__global__ void __launch_bounds__(STEM_TX * STEM_TY)
stem_conv_bn_gelu_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                         const float *__restrict__ bn_mean, const float *__restrict__ bn_var, const float *__restrict__ bn_w,
                         const float *__restrict__ bn_b, float eps, float *__restrict__ y, int H, int W, int OH, int OW) {
    constexpr int TAPS = IC * 9;
    __shared__ __align__(16) float s_w[TAPS * OC];       // [tap][oc]
    __shared__ float s_bias[OC], s_mean[OC], s_inv[OC], s_g[OC], s_b[OC];
    const int tid = threadIdx.y * STEM_TX + threadIdx.x;
    for (int i = tid; i < TAPS * OC; i += STEM_TX * STEM_TY) {
        const int tap = i / OC, oc = i - tap * OC;
        s_w[i] = __ldg(w + oc * TAPS + tap);             // weight [OC][IC][3][3] -> [tap = (ic, ky, kx)][oc]
    }
    for (int c = tid; c < OC; c += STEM_TX * STEM_TY) {
        s_bias[c] = bias ? __ldg(bias + c) : 0.f;
        s_mean[c] = __ldg(bn_mean + c);
        s_inv[c] = 1.f / sqrtf(__ldg(bn_var + c) + eps);
        s_g[c] = bn_w ? __ldg(bn_w + c) : 1.f;
        s_b[c] = bn_b ? __ldg(bn_b + c) : 0.f;
    }
    __syncthreads();
    const int ox = blockIdx.x * STEM_TX + threadIdx.x, oy = blockIdx.y * STEM_TY + threadIdx.y, b = blockIdx.z;
    if (ox >= OW || oy >= OH) return;
    float in[TAPS];
    const float *xb = x + (int64_t)b * IC * H * W;
#pragma unroll
    for (int ic = 0; ic < IC; ++ic)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = 2 * oy - 1 + ky;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * ox - 1 + kx;
                in[(ic * 3 + ky) * 3 + kx] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(xb + ((int64_t)ic * H + iy) * W + ix) : 0.f;
            }
        }
    float acc[OC];
#pragma unroll
    for (int c = 0; c < OC; ++c) acc[c] = 0.f;
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
#pragma unroll
        for (int c4 = 0; c4 < OC / 4; ++c4) {
            const float4 wv = *reinterpret_cast<const float4 *>(s_w + t * OC + 4 * c4);
            acc[4 * c4] = fmaf(in[t], wv.x, acc[4 * c4]);
            acc[4 * c4 + 1] = fmaf(in[t], wv.y, acc[4 * c4 + 1]);
            acc[4 * c4 + 2] = fmaf(in[t], wv.z, acc[4 * c4 + 2]);
            acc[4 * c4 + 3] = fmaf(in[t], wv.w, acc[4 * c4 + 3]);
        }
    }
    float *yb = y + ((int64_t)b * OC * OH + oy) * OW + ox;
#pragma unroll
    for (int c = 0; c < OC; ++c) {
        float v = acc[c] + s_bias[c];
        v = fmaf((v - s_mean[c]) * s_inv[c], s_g[c], s_b[c]);
        v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        yb[(int64_t)c * OH * OW] = v;
    }
}

template <int OC>
static int stem_launch(const float *x, const float *w, const float *bias, const float *m, const float *v, const float *g, const float *b,
                       float eps, float *y, int B, int H, int W, cudaStream_t st) {
    const int OH = (H + 1) / 2, OW = (W + 1) / 2;
    const dim3 grid((OW + STEM_TX - 1) / STEM_TX, (OH + STEM_TY - 1) / STEM_TY, B), block(STEM_TX, STEM_TY);
    stem_conv_bn_gelu_kernel<OC, 3><<<grid, block, 0, st>>>(x, w, bias, m, v, g, b, eps, y, H, W, OH, OW);
    note_launches(1);
    return check_launch("stem_conv_bn_gelu");
}

}  // namespace clusten

using namespace clusten;

extern "C" int clusten_stem_conv_bn_gelu(const float *x, const float *weight, const float *bias, const float *bn_mean, const float *bn_var,
                                         const float *bn_weight, const float *bn_bias, float eps, float *y, int B, int IC, int H, int W,
                                         int OC, void *stream) {
    if (B < 0 || H <= 0 || W <= 0 || IC <= 0 || OC <= 0) return set_error(CLUSTEN_EINVAL, "stem: bad sizes B=%d IC=%d H=%d W=%d OC=%d", B, IC, H, W, OC);
    if (!x || !weight || !bn_mean || !bn_var || !y) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (IC != 3 || B > 65535) return set_error(CLUSTEN_EUNSUPPORTED, "stem: IC=%d (3 supported), B=%d (<= 65535)", IC, B);
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (OC) {
        case 16: return stem_launch<16>(x, weight, bias, bn_mean, bn_var, bn_weight, bn_bias, eps, y, B, H, W, st);
        case 24: return stem_launch<24>(x, weight, bias, bn_mean, bn_var, bn_weight, bn_bias, eps, y, B, H, W, st);
        case 32: return stem_launch<32>(x, weight, bias, bn_mean, bn_var, bn_weight, bn_bias, eps, y, B, H, W, st);
        case 48: return stem_launch<48>(x, weight, bias, bn_mean, bn_var, bn_weight, bn_bias, eps, y, B, H, W, st);
        case 64: return stem_launch<64>(x, weight, bias, bn_mean, bn_var, bn_weight, bn_bias, eps, y, B, H, W, st);
    }
    return set_error(CLUSTEN_EUNSUPPORTED, "stem: OC=%d (16, 24, 32, 48, 64 supported)", OC);
}
