// Gradients of the blank-token parameters of the fused ClusterAttention core (mask2former/modeling/backbone/aff.py:138-146
// run backwards):
//     d_blank_k[h,c] = sum_{b,i} dS_blank[b,h,i] * q[b,h,i,c]          d_blank_v[h,c] = sum_{b,i} P_blank[b,h,i] * dO[b,h,i,c]
// i.e. two column sums of row-scaled [B*N, H*C] matrices.  torch.einsum turns each into a skinny GEMM (nvjet 64x8 tiles,
// ~96 us per call at AFF-Tiny stage 0: 9 % of the training step in profiles/r1_launches_tiny_train_v6); here one pass reads q
// and dO once: a thread owns one 16-byte channel chunk and strides over the rows of its CTA's slab, partial sums meet in
// shared memory, one fp32 atomic per (CTA, channel).
#include "common.cuh"

namespace clusten {

template <typename T>
__global__ void __launch_bounds__(256)
blank_grad_kernel(const T *__restrict__ q, const T *__restrict__ dO, const float *__restrict__ dSb, const float *__restrict__ Pb,
                  float *__restrict__ d_bk, float *__restrict__ d_bv, int B, int H, int N, int C,
                  int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t g_sb, int64_t g_sh, int64_t g_sn, int rows_per_cta) {
    extern __shared__ float red[];                       // [row lanes][chunks][16]
    const int cpr = C >> 3, chunks = H * cpr;            // 16-byte chunks per head row / per token
    const int RL = blockDim.x / chunks;                  // row lanes of the CTA
    const int ck = threadIdx.x % chunks, rl = threadIdx.x / chunks;
    const int h = ck / cpr, c0 = 8 * (ck - h * cpr);
    float ak[8], av[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) ak[x] = av[x] = 0.f;
    const int R = B * N;                                 // (host-checked < 2^31)
    if (rl < RL) {
        const float *s1p = dSb + (int64_t)h * N, *s2p = Pb + (int64_t)h * N;
        const T *qp = q + h * q_sh + c0, *gp = dO + h * g_sh + c0;
        for (int slab = blockIdx.x * rows_per_cta; slab < R; slab += gridDim.x * rows_per_cta) {
            const int rend = min(slab + rows_per_cta, R);
            for (int r0 = slab + rl; r0 < rend; r0 += 4 * RL) {
                float fq[4][8], fg[4][8], s1[4], s2[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {                // four independent rows in flight
                    const int r = min(r0 + u * RL, rend - 1);
                    const int b = r / N, i = r - b * N;
                    const bool on = r0 + u * RL < rend;
                    s1[u] = on ? s1p[(int64_t)b * H * N + i] : 0.f;
                    s2[u] = on ? s2p[(int64_t)b * H * N + i] : 0.f;
                    load16(qp + b * q_sb + (int64_t)i * q_sn, fq[u]);
                    load16(gp + b * g_sb + (int64_t)i * g_sn, fg[u]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int x = 0; x < 8; ++x) { ak[x] = fmaf(s1[u], fq[u][x], ak[x]); av[x] = fmaf(s2[u], fg[u][x], av[x]); }
            }
        }
        float *dst = red + ((size_t)rl * chunks + ck) * 16;
#pragma unroll
        for (int x = 0; x < 8; ++x) { dst[x] = ak[x]; dst[8 + x] = av[x]; }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < chunks * 16; t += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < RL; ++l) s += red[(size_t)l * chunks * 16 + t];
        const int kc = t >> 4, x = t & 15;
        const int hh = kc / cpr, cc = 8 * (kc - hh * cpr) + (x & 7);
        if (s != 0.f) atomicAdd((x < 8 ? d_bk : d_bv) + hh * C + cc, s);
    }
}

}  // namespace clusten

using namespace clusten;

extern "C" int clusten_blank_grad(const void *q, const void *d_out, const float *dS_blank, const float *P_blank,
                                  float *d_blank_k, float *d_blank_v, int B, int H, int N, int C,
                                  int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t g_sb, int64_t g_sh, int64_t g_sn,
                                  int dtype, void *stream) {
    if (B < 0 || H <= 0 || N < 0 || C <= 0) return set_error(CLUSTEN_EINVAL, "bad sizes B=%d H=%d N=%d C=%d", B, H, N, C);
    if (!q || !d_out || !dS_blank || !P_blank || !d_blank_k || !d_blank_v) return set_error(CLUSTEN_EINVAL, "null pointer");
    if (dtype != CLUSTEN_F16 && dtype != CLUSTEN_BF16) return set_error(CLUSTEN_EUNSUPPORTED, "blank_grad: 16-bit types only");
    const int chunks = H * (C >> 3);
    if (C % 8 || chunks > 256 || !aligned16(q) || !aligned16(d_out) || q_sb % 8 || q_sh % 8 || q_sn % 8 || g_sb % 8 || g_sh % 8 || g_sn % 8)
        return set_error(CLUSTEN_EUNSUPPORTED, "blank_grad: needs C %% 8 == 0, H*C <= 2048 and 16-byte aligned rows");
    if ((int64_t)B * N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int RL = 256 / chunks;
    const int64_t R = (int64_t)B * N;
    if (R >= (1LL << 31)) return set_error(CLUSTEN_EUNSUPPORTED, "blank_grad: B*N too large");
    // ~4 CTAs per SM, each with a slab of at least 8 rows per row lane
    int rows_per_cta = (int)((R + 148 * 4 - 1) / (148 * 4));
    rows_per_cta = std::max(8 * RL, (rows_per_cta + RL - 1) / RL * RL);
    const int grid = (int)std::min<int64_t>((R + rows_per_cta - 1) / rows_per_cta, 148 * 4);
    const size_t smem = (size_t)RL * chunks * 16 * sizeof(float);
    if (dtype == CLUSTEN_BF16)
        blank_grad_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>((const __nv_bfloat16 *)q, (const __nv_bfloat16 *)d_out, dS_blank, P_blank,
                                                                 d_blank_k, d_blank_v, B, H, N, C, q_sb, q_sh, q_sn, g_sb, g_sh, g_sn, rows_per_cta);
    else
        blank_grad_kernel<__half><<<grid, 256, smem, st>>>((const __half *)q, (const __half *)d_out, dS_blank, P_blank, d_blank_k, d_blank_v,
                                                          B, H, N, C, q_sb, q_sh, q_sn, g_sb, g_sh, g_sn, rows_per_cta);
    note_launches(1);
    return check_launch("blank_grad");
}
