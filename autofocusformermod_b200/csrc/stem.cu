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

// NHWC: y is written pixel-major [B, OH, OW, OC] (each thread stores its OC values as 16-byte pieces) -- the layout the im2col pass
// below gathers from; else NCHW, what cuDNN's second convolution takes.
template <int OC, int IC, bool NHWC>
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
#pragma unroll
    for (int c = 0; c < OC; ++c) {
        float v = acc[c] + s_bias[c];
        v = fmaf((v - s_mean[c]) * s_inv[c], s_g[c], s_b[c]);
        acc[c] = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    }
    if constexpr (NHWC) {
        float4 *yp = reinterpret_cast<float4 *>(y + (((int64_t)b * OH + oy) * OW + ox) * OC);
#pragma unroll
        for (int c4 = 0; c4 < OC / 4; ++c4) yp[c4] = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
    } else {
        float *yb = y + ((int64_t)b * OC * OH + oy) * OW + ox;
#pragma unroll
        for (int c = 0; c < OC; ++c) yb[(int64_t)c * OH * OW] = acc[c];
    }
}

// im2col of a 3x3 / stride 2 / padding 1 convolution over a pixel-major map: A[(b, py, px), (ky * 3 + kx) * C + c] =
// mid[b, 2 py - 1 + ky, 2 px - 1 + kx, c] (0 outside the map), columns 9 C .. Kp - 1 zero (K padded to the GEMM's chunk).  One thread per
// 16-byte piece of an A row: writes are contiguous, reads are C-float runs.  The second stem convolution (proj2, aff.py:530,549) then is
// ONE GEMM [B OH OW, Kp] x [E, Kp]^T on the tcgen05 Linear kernel, whose output rows are the tokens -- cuDNN answers that convolution
// with a SIMT sgemm (0.80 ms of the 13.6 ms AFF-Small forward) plus a bias pass plus the NCHW -> token-major copy.
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float4 *__restrict__ mid, float4 *__restrict__ A, int B, int H, int W, int C4, int OH, int OW, int Kp4) {
    const int64_t total = (int64_t)B * OH * OW * Kp4;
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = x / Kp4;
        const int p = (int)(x - row * Kp4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < 9 * C4) {
            const int tap = p / C4, c4 = p - tap * C4;
            const int ky = tap / 3, kx = tap - 3 * ky;
            const int px = (int)(row % OW);
            const int64_t t = row / OW;
            const int py = (int)(t % OH), b = (int)(t / OH);
            const int iy = 2 * py - 1 + ky, ix = 2 * px - 1 + kx;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(mid + (((int64_t)b * H + iy) * W + ix) * C4 + c4);
        }
        A[x] = v;
    }
}

template <int OC>
static int stem_launch(const float *x, const float *w, const float *bias, const float *m, const float *v, const float *g, const float *b,
                       float eps, float *y, int B, int H, int W, bool nhwc, cudaStream_t st) {
    const int OH = (H + 1) / 2, OW = (W + 1) / 2;
    const dim3 grid((OW + STEM_TX - 1) / STEM_TX, (OH + STEM_TY - 1) / STEM_TY, B), block(STEM_TX, STEM_TY);
    if (nhwc) stem_conv_bn_gelu_kernel<OC, 3, true><<<grid, block, 0, st>>>(x, w, bias, m, v, g, b, eps, y, H, W, OH, OW);
    else stem_conv_bn_gelu_kernel<OC, 3, false><<<grid, block, 0, st>>>(x, w, bias, m, v, g, b, eps, y, H, W, OH, OW);
    note_launches(1);
    return check_launch("stem_conv_bn_gelu");
}

}  // namespace clusten

using namespace clusten;

extern "C" int clusten_stem_conv_bn_gelu(const float *x, const float *weight, const float *bias, const float *bn_mean, const float *bn_var,
                                         const float *bn_weight, const float *bn_bias, float eps, float *y, int B, int IC, int H, int W,
                                         int OC, int channels_last, void *stream) {
    if (B < 0 || H <= 0 || W <= 0 || IC <= 0 || OC <= 0) return set_error(CLUSTEN_EINVAL, "stem: bad sizes B=%d IC=%d H=%d W=%d OC=%d", B, IC, H, W, OC);
    if (!x || !weight || !bn_mean || !bn_var || !y) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (IC != 3 || B > 65535) return set_error(CLUSTEN_EUNSUPPORTED, "stem: IC=%d (3 supported), B=%d (<= 65535)", IC, B);
    if (B == 0) return 0;
    if (channels_last && !aligned16(y)) return set_error(CLUSTEN_EUNSUPPORTED, "stem: the pixel-major output must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const bool nhwc = channels_last != 0;
    switch (OC) {
        case 16: return stem_launch<16>(x, weight, bias, bn_mean, bn_var, bn_weight, bn_bias, eps, y, B, H, W, nhwc, st);
        case 24: return stem_launch<24>(x, weight, bias, bn_mean, bn_var, bn_weight, bn_bias, eps, y, B, H, W, nhwc, st);
        case 32: return stem_launch<32>(x, weight, bias, bn_mean, bn_var, bn_weight, bn_bias, eps, y, B, H, W, nhwc, st);
        case 48: return stem_launch<48>(x, weight, bias, bn_mean, bn_var, bn_weight, bn_bias, eps, y, B, H, W, nhwc, st);
        case 64: return stem_launch<64>(x, weight, bias, bn_mean, bn_var, bn_weight, bn_bias, eps, y, B, H, W, nhwc, st);
    }
    return set_error(CLUSTEN_EUNSUPPORTED, "stem: OC=%d (16, 24, 32, 48, 64 supported)", OC);
}

extern "C" int clusten_stem_im2col(const float *mid, float *A, int B, int H, int W, int C, int Kp, void *stream) {
    if (B < 0 || H <= 0 || W <= 0 || C <= 0 || Kp < 9 * C) return set_error(CLUSTEN_EINVAL, "im2col: bad sizes B=%d H=%d W=%d C=%d Kp=%d", B, H, W, C, Kp);
    if (!mid || !A) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (C % 4 || Kp % 4 || !aligned16(mid) || !aligned16(A)) return set_error(CLUSTEN_EUNSUPPORTED, "im2col: C and Kp must be multiples of 4, 16-byte aligned buffers");
    if (B == 0) return 0;
    const int OH = (H + 1) / 2, OW = (W + 1) / 2;
    const int64_t total = (int64_t)B * OH * OW * (Kp / 4);
    const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 64);
    stem_im2col_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(mid), reinterpret_cast<float4 *>(A), B, H, W, C / 4, OH, OW, Kp / 4);
    note_launches(1);
    return check_launch("stem_im2col");
}
