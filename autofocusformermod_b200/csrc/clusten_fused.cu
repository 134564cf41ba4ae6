// Fused ClusterAttention core, forward (sm_100a): QK logits + relative-position bias + cluster mask + blank token +
// softmax + attention-times-V in ONE kernel -- the glue the reference runs as >= 7 separate passes over the
// [B,H,N,M] tensor between its two CUDA ops (mask2former/modeling/backbone/aff.py:114-155):
//
//   logit[b,h,i,j] = q[b,h,i,:] . k[b,h,idx[b,i,j],:] + bias_tab[bias_idx[b,i,j], h] + (mask[b,i,j] ? 0 : -100)     :114-137
//   logit[b,h,i,M] = q[b,h,i,:] . blank_k[h,:]                                                                       :140
//   p = softmax over the M+1 logits                                                                                   :141-142
//   out[b,i,h,:]   = sum_j p[j] v[b,h,idx[b,i,j],:] + p[M] blank_v[h,:]                                               :145-155
//
// (bias_tab[r, h] = pos_embed(pre_table)[r, h] restricted to the rows a stage references, bias_idx = the matching
// inverse map; both are produced once per stage by the caller.)  One warp owns a 16-token tile and one head, exactly
// as in clusten_tile.cu: phase 1 runs the dot shape on the tensor cores and parks the selected logits in shared
// memory, phase 2 is a two-lanes-per-row softmax in shared memory, phase 3 runs the axpy shape with the
// probabilities as the A operand.  The [B,H,N,M] tensor never touches HBM (optionally the probabilities are written
// once, coalesced, for a backward pass).  Impure tokens and index tensors without octet structure take a
// one-warp-per-(token, head) generic kernel with the same arithmetic; the pack's device-side flag picks (no host sync).
#include "fused.cuh"

namespace clusten {

// ---- small pieces shared with clusten_tile.cu (kept local: every kernel file is self-contained) -----------------------
__device__ __forceinline__ void f_mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, __nv_bfloat16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void f_mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, __half) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void f_mma1688(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void f_split(float x, uint32_t &hi, uint32_t &lo) {      // truncating split (see clusten_tile.cu: tf32_split)
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void f_mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                       uint32_t b0h, uint32_t b1h, uint32_t b0l, uint32_t b1l) {
    f_mma1688(d, al, b0h, b1h);
    f_mma1688(d, ah, b0l, b1l);
    f_mma1688(d, ah, b0h, b1h);
}
__device__ __forceinline__ void f_cp16(void *smem, const void *gmem, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void f_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void f_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void f_ldsm4t(uint32_t (&r)[4], const void *p) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ uint32_t pack2(float a, float b, __nv_bfloat16) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&v);
}
__device__ __forceinline__ uint32_t pack2(float a, float b, __half) {
    const __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&v);
}
__device__ __forceinline__ uint32_t pack2(float, float, float) { return 0u; }
template <typename T> __device__ __forceinline__ void f_store2(T *p, float a, float b);
template <> __device__ __forceinline__ void f_store2<float>(float *p, float a, float b) { *reinterpret_cast<float2 *>(p) = make_float2(a, b); }
template <> __device__ __forceinline__ void f_store2<__half>(__half *p, float a, float b) { *reinterpret_cast<__half2 *>(p) = __floats2half2_rn(a, b); }
template <> __device__ __forceinline__ void f_store2<__nv_bfloat16>(__nv_bfloat16 *p, float a, float b) { *reinterpret_cast<__nv_bfloat162 *>(p) = __floats2bfloat162_rn(a, b); }

template <typename T, bool PB>
__global__ void __launch_bounds__(256)
attn_fused_generic_kernel(const FArgsOf<PB> a, const int *__restrict__ tile_flag) {
    extern __shared__ __align__(16) float dyn_f[];
    if (tile_flag && tile_flag[0] == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    float *sm = dyn_f + warp * (a.M + 2);
    const int64_t total = (int64_t)a.B * a.Nq * a.H;
    for (int64_t it = (int64_t)blockIdx.x * W + warp; it < total; it += (int64_t)gridDim.x * W) {
        const int h = (int)(it % a.H);
        const int64_t bi = it / a.H;
        fused_row_generic<T, PB>(a, (int)(bi / a.Nq), h, (int)(bi % a.Nq), sm, lane);
    }
}

// CH channels of one row as 32-bit registers
template <typename T, int CH> struct FFrag { uint32_t r[CH * sizeof(T) / 4]; };
template <typename T, int CH> __device__ __forceinline__ void f_load(FFrag<T, CH> &f, const T *p, bool pred) {
    constexpr int NB = CH * sizeof(T);
#pragma unroll
    for (int x = 0; x < NB / 4; ++x) f.r[x] = 0u;
    if (!pred) return;
    if constexpr (NB == 8) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        f.r[0] = v.x; f.r[1] = v.y;
    } else {
#pragma unroll
        for (int x = 0; x < NB / 16; ++x) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p) + x);
            f.r[4 * x] = v.x; f.r[4 * x + 1] = v.y; f.r[4 * x + 2] = v.z; f.r[4 * x + 3] = v.w;
        }
    }
}
template <typename T, int CH> __device__ __forceinline__ float f_elem(const FFrag<T, CH> &f, int x) {      // channel x of the chunk
    if constexpr (sizeof(T) == 4) return __uint_as_float(f.r[x]);
    else {
        const uint32_t w = f.r[x >> 1];
        const unsigned short hbits = (x & 1) ? (unsigned short)(w >> 16) : (unsigned short)(w & 0xffffu);
        if constexpr (std::is_same<T, __half>::value) return __half2float(__ushort_as_half(hbits));
        else return __bfloat162float(__ushort_as_bfloat16(hbits));
    }
}

// (3 CTAs of 8 warps per SM: without the bound the rare masked-neighbour path lifts the register count past 85 and costs the
// common path a third of its resident warps; with it the spills, if any, sit in that rare path)
template <typename T, int CH, int NT, bool PB>
__global__ void __launch_bounds__(256, 3)
attn_fused_tile_kernel(const FArgsOf<PB> a, const PackView pk, int smem_per_warp) {
    extern __shared__ __align__(16) unsigned char dyn[];
    if (pk.flags[0]) return;
    constexpr bool F32 = sizeof(T) == 4;
    constexpr int KS = F32 ? CH / 2 : CH / 4;                       // mma k-steps of the dot phase
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const int M = a.M, C = a.C, H = a.H, Nq = a.Nq, MP = M + 4;
    float *S = reinterpret_cast<float *>(dyn + (size_t)warp * smem_per_warp);        // [16][MP]: logits / e, [M] blank, [M+1] 1/sum
    unsigned char *stg = reinterpret_cast<unsigned char *>(S + 16 * MP);              // V staging, 2 buffers
    const int tile = blockIdx.x * W + warp;                                           // grid: (tiles / W, B * H)
    if (tile < pk.T) {
    const int b = blockIdx.y / H, h = blockIdx.y - b * H;
    const int bt = b * pk.T + tile, i0 = tile * TILE_TOK;
    const int g = lane >> 2, t = lane & 3;
    const int U = pk.tile_u[bt];
    const int *octp = pk.tile_oct + bt * U_MAX;
    const int o0 = octp[lane], o1 = octp[32 + (lane & 15)];
    auto octet = [&](int u) { u = min(u, U - 1); return __shfl_sync(FULL, u < 32 ? o0 : o1, u & 31); };
    const int8_t *sa = pk.slot_of + (bt * TILE_TOK + g) * U_MAX;
    const int8_t *sb = sa + 8 * U_MAX;
    const int ra = i0 + g, rb = ra + 8;
    uint32_t impm = 0;
    {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(pk.tok_imp + bt * TILE_TOK));
        if (v.x | v.y | v.z | v.w) {
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int qd = 0; qd < 4; ++qd)
#pragma unroll
                for (int kq = 0; kq < 4; ++kq) impm |= ((w4[qd] >> (8 * kq)) & 1u) << (4 * qd + kq);
        }
    }
    // per-warp 64-bit bases hidden from the optimiser + 32-bit element offsets (host-checked extents): one IMAD.WIDE per address
    const T *Q = t2::opaque(reinterpret_cast<const T *>(a.q) + b * a.q_sb + h * a.q_sh);
    const T *K = t2::opaque(reinterpret_cast<const T *>(a.k) + b * a.k_sb + h * a.k_sh);
    const T *V = t2::opaque(reinterpret_cast<const T *>(a.v) + b * a.v_sb + h * a.v_sh);
    const int q_sn = (int)a.q_sn, k_sn = (int)a.k_sn, v_sn = (int)a.v_sn;

    // ---- phase 1: logits of the 16 tokens against every union octet (tensor cores), selected blocks -> S ------------------
    {
        const bool cact = CH * t < C;
        const int cofs = cact ? CH * t : 0;
        FFrag<T, CH> xa, xb;
        // rows beyond Nq re-read the last row (their slots are all -1 and phase 2 zeroes them); inactive channel lanes zero q
        t2::ld_chunk<CH * (int)sizeof(T)>(xa.r, t2::at(Q, min(ra, Nq - 1) * q_sn + cofs));
        t2::ld_chunk<CH * (int)sizeof(T)>(xb.r, t2::at(Q, min(rb, Nq - 1) * q_sn + cofs));
        if (!cact) {
#pragma unroll
            for (int x = 0; x < CH * (int)sizeof(T) / 4; ++x) xa.r[x] = xb.r[x] = 0u;
        }
        {   // blank logit q . blank_k[h] (aff.py:140): partial over this lane's channels, reduced over the 4 lanes of the row
            const T *bk = reinterpret_cast<const T *>(a.blank_k) + h * C + CH * t;
            float pa = 0.f, pb = 0.f;
#pragma unroll
            for (int x = 0; x < CH; ++x) {
                const float w = (CH * t + x < C) ? to_f(bk[x]) : 0.f;
                pa = fmaf(f_elem<T, CH>(xa, x), w, pa);
                pb = fmaf(f_elem<T, CH>(xb, x), w, pb);
            }
            pa += __shfl_xor_sync(FULL, pa, 1); pa += __shfl_xor_sync(FULL, pa, 2);
            pb += __shfl_xor_sync(FULL, pb, 1); pb += __shfl_xor_sync(FULL, pb, 2);
            if (t == 0) { S[g * MP + M] = pa; S[(g + 8) * MP + M] = pb; }
        }
        uint32_t ah[KS][4], al[F32 ? KS : 1][4];
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            if constexpr (F32) {
                f_split(__uint_as_float(xa.r[2 * s]), ah[s][0], al[s][0]);
                f_split(__uint_as_float(xb.r[2 * s]), ah[s][1], al[s][1]);
                f_split(__uint_as_float(xa.r[2 * s + 1]), ah[s][2], al[s][2]);
                f_split(__uint_as_float(xb.r[2 * s + 1]), ah[s][3], al[s][3]);
            } else {
                ah[s][0] = xa.r[2 * s]; ah[s][1] = xb.r[2 * s]; ah[s][2] = xa.r[2 * s + 1]; ah[s][3] = xb.r[2 * s + 1];
            }
        }
        constexpr int UB = F32 ? 2 : 4;
        const int klast = a.Nk - 1;                      // (mask-aware packs may hold the partial last octet: rows are clamped)
        // PB: the lane owns the logits of rows g / g+8 against keys 2t, 2t+1 of each octet -> their bias from the four positions
        PosBiasW pw = {};
        float2 pqa = make_float2(0.f, 0.f), pqb = pqa;
        const float2 *PK = nullptr;
        if constexpr (PB) {
            pw = pos_bias_load(a.pe_w, a.pe_b, h);
            const float2 *PQ = reinterpret_cast<const float2 *>(a.pos_q) + (int64_t)b * Nq;
            pqa = __ldg(PQ + min(ra, Nq - 1));
            pqb = __ldg(PQ + min(rb, Nq - 1));
            PK = reinterpret_cast<const float2 *>(a.pos_k) + (int64_t)b * a.Nk;
        }
        for (int u0 = 0; u0 < U; u0 += UB) {
            FFrag<T, CH> y[UB];
            int sga[UB], sgb[UB], oct8[UB];
            uint32_t sa4, sb4;                           // slots of rows g / g+8 for the UB octets (positions >= U hold -1)
            if constexpr (UB == 4) {
                sa4 = __ldg(reinterpret_cast<const uint32_t *>(sa + u0));
                sb4 = __ldg(reinterpret_cast<const uint32_t *>(sb + u0));
            } else {
                sa4 = __ldg(reinterpret_cast<const unsigned short *>(sa + u0));
                sb4 = __ldg(reinterpret_cast<const unsigned short *>(sb + u0));
            }
#pragma unroll
            for (int j = 0; j < UB; ++j) {
                oct8[j] = octet(u0 + j) * 8;
                t2::ld_chunk<CH * (int)sizeof(T)>(y[j].r, t2::at(K, min(oct8[j] + g, klast) * k_sn + cofs));
                sga[j] = t2::sbyte(sa4, j);
                sgb[j] = t2::sbyte(sb4, j);
            }
#pragma unroll
            for (int j = 0; j < UB; ++j) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    if constexpr (F32) {
                        uint32_t b0h, b0l, b1h, b1l;
                        f_split(__uint_as_float(y[j].r[2 * s]), b0h, b0l);
                        f_split(__uint_as_float(y[j].r[2 * s + 1]), b1h, b1l);
                        f_mma3(acc, ah[s], al[s], b0h, b1h, b0l, b1l);
                    } else {
                        f_mma16816(acc, ah[s], y[j].r[2 * s], y[j].r[2 * s + 1], T());
                    }
                }
                if constexpr (PB) {
                    const float2 k0 = __ldg(PK + min(oct8[j] + 2 * t, klast)), k1 = __ldg(PK + min(oct8[j] + 2 * t + 1, klast));
                    acc[0] += pos_bias(pw, pqa, k0); acc[1] += pos_bias(pw, pqa, k1);
                    acc[2] += pos_bias(pw, pqb, k0); acc[3] += pos_bias(pw, pqb, k1);
                }
                if (sga[j] >= 0) *reinterpret_cast<float2 *>(S + g * MP + 8 * sga[j] + 2 * t) = make_float2(acc[0], acc[1]);
                if (sgb[j] >= 0) *reinterpret_cast<float2 *>(S + (g + 8) * MP + 8 * sgb[j] + 2 * t) = make_float2(acc[2], acc[3]);
            }
        }
    }
    __syncwarp();
    // ---- phase 2: + bias + mask, softmax over M + 1 logits; two lanes per token row, e_j stay unnormalised in S ---------------
    bool saw_mask = false;
    {
        const int row = lane >> 1, half = lane & 1;
        const int i = i0 + row;
        const bool rvalid = i < Nq && !((impm >> row) & 1u);
        float *Sr = S + row * MP;
        const int Mh = M >> 1, j0 = half * Mh, j1 = j0 + Mh;
        float mx = -INFINITY;
        if (rvalid) {
            const int32_t *bi = PB ? nullptr : a.bias_idx + ((int64_t)b * Nq + i) * M;
            const uint8_t *mk = a.mask ? a.mask + ((int64_t)b * Nq + i) * M : nullptr;
            for (int j = j0; j < j1; j += 4) {                       // M % 8 == 0 -> Mh % 4 == 0, 16-byte aligned rows
                float4 x = *reinterpret_cast<float4 *>(Sr + j);
                float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
                if constexpr (!PB) {                                 // (PB: the bias went in with the logits in phase 1)
                    const int4 bv = __ldg(reinterpret_cast<const int4 *>(bi + j));
                    g0 = __ldg(a.bias_tab + bv.x * H + h); g1 = __ldg(a.bias_tab + bv.y * H + h);
                    g2 = __ldg(a.bias_tab + bv.z * H + h); g3 = __ldg(a.bias_tab + bv.w * H + h);
                    x.x += g0; x.y += g1; x.z += g2; x.w += g3;
                }
                if (mk) {
                    const uchar4 m4 = *reinterpret_cast<const uchar4 *>(mk + j);
                    if (!(m4.x && m4.y && m4.z && m4.w)) {
                        // masked neighbours: the wildcard column's q.k term (and, PB, its bias) is dropped here; the exact logit
                        // is added by masked_logit_pass below (fused.cuh)
                        saw_mask = true;
                        if (!m4.x) x.x = g0;
                        if (!m4.y) x.y = g1;
                        if (!m4.z) x.z = g2;
                        if (!m4.w) x.w = g3;
                    }
                }
                *reinterpret_cast<float4 *>(Sr + j) = x;
                mx = fmaxf(fmaxf(mx, fmaxf(x.x, x.y)), fmaxf(x.z, x.w));
            }
            if (half == 0) mx = fmaxf(mx, Sr[M]);
        }
        if (__any_sync(FULL, saw_mask)) {                            // (rare: tiles that touch the padded cluster)
            __syncwarp();
            masked_logit_pass<T, PB>(a, b, h, i0, min(TILE_TOK, Nq - i0), impm, S, MP, lane);
            if (rvalid) {                                            // the row maxima again, now over the exact masked logits
                mx = -INFINITY;
                for (int j = j0; j < j1; j += 4) {
                    const float4 x = *reinterpret_cast<float4 *>(Sr + j);
                    mx = fmaxf(fmaxf(mx, fmaxf(x.x, x.y)), fmaxf(x.z, x.w));
                }
                if (half == 0) mx = fmaxf(mx, Sr[M]);
            }
        }
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 1));
        float sum = 0.f;
        if (rvalid) {
            for (int j = j0; j < j1; j += 4) {
                float4 x = *reinterpret_cast<float4 *>(Sr + j);
                x.x = f_exp<T>(x.x - mx); x.y = f_exp<T>(x.y - mx); x.z = f_exp<T>(x.z - mx); x.w = f_exp<T>(x.w - mx);
                sum += (x.x + x.y) + (x.z + x.w);
                *reinterpret_cast<float4 *>(Sr + j) = x;
            }
            if (half == 0) { const float e = f_exp<T>(Sr[M] - mx); Sr[M] = e; sum += e; }
        } else {
            for (int j = j0; j < j1; j += 4) *reinterpret_cast<float4 *>(Sr + j) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (half == 0) Sr[M] = 0.f;
        }
        sum += __shfl_xor_sync(FULL, sum, 1);
        if (half == 0) Sr[M + 1] = rvalid ? 1.f / sum : 0.f;
        if (a.lse && half == 0 && rvalid) a.lse[((int64_t)b * H + h) * Nq + i] = mx + logf(sum);
    }
    __syncwarp();
    // ---- phase 3: out = sum_j e_j v_j over the union octets (tensor cores, V staged key-major in shared memory) -------------
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    const T *vbase = V;
    const float *Sa = S + g * MP, *Sb = S + (g + 8) * MP;
    if (__any_sync(FULL, saw_mask)) {
        // masked entries: e_j leaves S and enters the accumulators with the row the reference reads, v[idx[b,i,j]] (fused.cuh); the
        // lanes that own the token's row in the mma layout take it
        masked_value_pass<T, PB>(a, b, i0, min(TILE_TOK, Nq - i0), impm, S, MP, lane, [&](int row, int64_t kidx, float e) {
            const T *vr = reinterpret_cast<const T *>(a.v) + b * a.v_sb + h * a.v_sh + kidx * a.v_sn + 2 * t;
            if (g == (row & 7)) {
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    if (8 * n + 2 * t < C) {
                        const float v0 = e * to_f(vr[8 * n]), v1 = e * to_f(vr[8 * n + 1]);
                        if (row < 8) { acc[n][0] += v0; acc[n][1] += v1; }
                        else { acc[n][2] += v0; acc[n][3] += v1; }
                    }
                }
            }
        });
    }
    if constexpr (!F32) {
        constexpr int ROWB = NT * 16 + 16, CPL = NT / 2;
        auto stage = [&](int p, int which) {
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const int ch = lane + 32 * j;
                const int row = ch / NT, blk = ch % NT;
                const int u = 2 * p + (row >> 3);
                const bool ok = u < U && 8 * blk < C;
                const T *src = t2::at(vbase, min(octet(u) * 8 + (row & 7), a.Nk - 1) * v_sn + (ok ? 8 * blk : 0));
                f_cp16(stg + which * 16 * ROWB + row * ROWB + blk * 16, src, ok);
            }
            f_commit();
        };
        const int P = (U + 1) >> 1;
        if (P > 0) stage(0, 0);
        for (int p = 0; p < P; ++p) {
            if (p + 1 < P) stage(p + 1, (p + 1) & 1);
            const int u0 = 2 * p, u1 = 2 * p + 1;
            const int s00 = sa[u0], s10 = sb[u0];
            const int s01 = u1 < U ? (int)sa[u1] : -1, s11 = u1 < U ? (int)sb[u1] : -1;
            uint32_t af[4];
            af[0] = s00 >= 0 ? pack2(Sa[8 * s00 + 2 * t], Sa[8 * s00 + 2 * t + 1], T()) : 0u;
            af[1] = s10 >= 0 ? pack2(Sb[8 * s10 + 2 * t], Sb[8 * s10 + 2 * t + 1], T()) : 0u;
            af[2] = s01 >= 0 ? pack2(Sa[8 * s01 + 2 * t], Sa[8 * s01 + 2 * t + 1], T()) : 0u;
            af[3] = s11 >= 0 ? pack2(Sb[8 * s11 + 2 * t], Sb[8 * s11 + 2 * t + 1], T()) : 0u;
            if (p + 1 < P) f_wait<1>(); else f_wait<0>();
            __syncwarp();
            const int mi = lane >> 3;
            const unsigned char *lrow = stg + (p & 1) * 16 * ROWB + (((mi & 1) << 3) + (lane & 7)) * ROWB + (mi >> 1) * 16;
#pragma unroll
            for (int n = 0; n < NT; n += 2) {
                uint32_t bfr[4];
                f_ldsm4t(bfr, lrow + n * 16);
                f_mma16816(acc[n], af, bfr[0], bfr[1], T());
                f_mma16816(acc[n + 1], af, bfr[2], bfr[3], T());
            }
            __syncwarp();
        }
    } else {
        constexpr int RS = NT * 8 + 8, CPL = NT / 2;
        float *stf = reinterpret_cast<float *>(stg);
        auto stage = [&](int u, int which) {
            const int o = octet(u);
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const int ch = lane + 32 * j;
                const int row = ch / (2 * NT), blk = ch % (2 * NT);
                const bool ok = 4 * blk < C;
                const T *src = t2::at(vbase, min(o * 8 + row, a.Nk - 1) * v_sn + (ok ? 4 * blk : 0));
                f_cp16(stf + which * 8 * RS + row * RS + blk * 4, src, ok);
            }
            f_commit();
        };
        if (U > 0) stage(0, 0);
        for (int u = 0; u < U; ++u) {
            if (u + 1 < U) stage(u + 1, (u + 1) & 1);
            const int s0 = sa[u], s1 = sb[u];
            uint32_t ah[4], al[4];
            f_split(s0 >= 0 ? Sa[8 * s0 + t] : 0.f, ah[0], al[0]);
            f_split(s1 >= 0 ? Sb[8 * s1 + t] : 0.f, ah[1], al[1]);
            f_split(s0 >= 0 ? Sa[8 * s0 + t + 4] : 0.f, ah[2], al[2]);
            f_split(s1 >= 0 ? Sb[8 * s1 + t + 4] : 0.f, ah[3], al[3]);
            if (u + 1 < U) f_wait<1>(); else f_wait<0>();
            __syncwarp();
            const float *bp = stf + (u & 1) * 8 * RS + t * RS + g;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                uint32_t b0h, b0l, b1h, b1l;
                f_split(bp[8 * n], b0h, b0l);
                f_split(bp[4 * RS + 8 * n], b1h, b1l);
                f_mma3(acc[n], ah, al, b0h, b1h, b0l, b1l);
            }
            __syncwarp();
        }
    }
    // ---- epilogue: + e_blank * blank_v, * 1/sum, token-major store ------------------------------------------------------------
    {
        const float inva = Sa[M + 1], invb = Sb[M + 1], eba = Sa[M], ebb = Sb[M];
        const T *bv = reinterpret_cast<const T *>(a.blank_v) + h * C;
        T *Ob = t2::opaque(reinterpret_cast<T *>(a.out) + b * a.o_sb + h * a.o_sh);
        T *oa = t2::at(Ob, ra * (int)a.o_sn + 2 * t);
        T *ob = t2::at(Ob, rb * (int)a.o_sn + 2 * t);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const int ch = 8 * n + 2 * t;
            if (ch < C) {
                const float v0 = to_f(bv[ch]), v1 = to_f(bv[ch + 1]);
                if (ra < Nq && !((impm >> g) & 1u)) f_store2<T>(oa + 8 * n, (acc[n][0] + eba * v0) * inva, (acc[n][1] + eba * v1) * inva);
                if (rb < Nq && !((impm >> (g + 8)) & 1u)) f_store2<T>(ob + 8 * n, (acc[n][2] + ebb * v0) * invb, (acc[n][3] + ebb * v1) * invb);
            }
        }
    }
    if (a.probs) {                                            // normalised probabilities, coalesced (16 rows x (M+1), rows adjacent)
        float *pr = a.probs + (((int64_t)b * H + h) * Nq + i0) * (M + 1);
        const int lim = min(TILE_TOK, Nq - i0) * (M + 1);
        for (int x = lane; x < lim; x += 32) {
            const int row = x / (M + 1), col = x - row * (M + 1);
            if (!((impm >> row) & 1u)) pr[x] = S[row * MP + col] * S[row * MP + M + 1];
        }
    }
    }
    __syncwarp();
    // impure tokens of the whole call, one (token, head) per warp, spread over the grid (t2::slow_items); S row 0 as scratch
    const t2::SlowIter si = t2::slow_items(pk, a.H);
    for (int it = si.first; it < si.n; it += si.stride) {
        const int gi = pk.imp_list[it / a.H], hh = it % a.H;
        const int bb = gi / a.Nq;
        fused_row_generic<T, PB>(a, bb, hh, gi - bb * a.Nq, S, lane);
    }
}

template <typename T> static bool f_rows_ok(const void *p, int64_t sb, int64_t sh, int64_t sn) {
    constexpr int VPT = 16 / sizeof(T);
    return aligned16(p) && sb % VPT == 0 && sh % VPT == 0 && sn % VPT == 0;
}

template <typename T, bool PB>
static int launch_fused(const FArgsOf<PB> &a, const void *pack, cudaStream_t st) {
    const int *flag = nullptr;
    const int M = a.M, C = a.C;
    const size_t per_warp = (size_t)16 * (M + 4) * 4 + 2560;
    const int W = per_warp * 8 <= 48 * 1024 ? 8 : 4;
    const bool shape_ok = C % 8 == 0 && C >= 8 && C <= 32 && M % 8 == 0 && M <= 256 && per_warp * W <= 48 * 1024;
    const bool align_ok = f_rows_ok<T>(a.q, a.q_sb, a.q_sh, a.q_sn) && f_rows_ok<T>(a.k, a.k_sb, a.k_sh, a.k_sn) &&
                          f_rows_ok<T>(a.v, a.v_sb, a.v_sh, a.v_sn) && f_rows_ok<T>(a.out, a.o_sb, a.o_sh, a.o_sn) &&
                          (PB || aligned16(a.bias_idx)) && (!a.mask || (reinterpret_cast<uintptr_t>(a.mask) & 3u) == 0);
    auto fits = [](int64_t v) { return v >= 0 && v < (1LL << 31); };
    const bool range_ok = fits(a.Nq * a.q_sn) && fits((int64_t)a.Nk * a.k_sn) && fits((int64_t)a.Nk * a.v_sn) && fits(a.Nq * a.o_sn) &&
                          (int64_t)a.B * a.H <= 65535 && fits((int64_t)a.B * ((a.Nq + 15) / 16) * 16 * U_MAX);
    bool taken = false;
    if (pack) {
        // the CTA-cooperative, TMA-staged kernel when the layout allows it (clusten_fused_tma.cu); else one warp per (tile, head)
        FusedArgsPB ap{};
        if constexpr (PB) ap = a; else static_cast<FusedArgs &>(ap) = a;
        constexpr int code = sizeof(T) == 4 ? CLUSTEN_F32 : std::is_same<T, __half>::value ? CLUSTEN_F16 : CLUSTEN_BF16;
        if (int e = fused_tma_launch(ap, PB, code, pack, st, &taken)) return e;
        if (taken) flag = reinterpret_cast<const int *>(pack);
    }
    if (!taken && pack && shape_ok && align_ok && range_ok) {
        const PackView pk = pack_view(const_cast<void *>(pack), a.B, a.Nq, a.Nk);
        const dim3 grid(ceil_div(pk.T, W), a.B * a.H);
        const size_t smem = per_warp * W;
        if (C <= 16) attn_fused_tile_kernel<T, 4, 2, PB><<<grid, W * 32, smem, st>>>(a, pk, (int)per_warp);
        else attn_fused_tile_kernel<T, 8, 4, PB><<<grid, W * 32, smem, st>>>(a, pk, (int)per_warp);
        note_launches(1);
        if (int e = check_launch("attn_fused_tile")) return e;
        flag = reinterpret_cast<const int *>(pack);
    }
    const int64_t total = (int64_t)a.B * a.Nq * a.H;
    int grid = ceil_div(total, 8);
    if (flag && grid > 148 * 8) grid = 148 * 8;
    const size_t smem = (size_t)8 * (M + 2) * sizeof(float);
    if (smem > 48 * 1024) return set_error(CLUSTEN_EUNSUPPORTED, "fused attention: M=%d too large", M);
    attn_fused_generic_kernel<T, PB><<<grid, 256, smem, st>>>(a, flag);
    note_launches(1);
    return check_launch("attn_fused_generic");
}

}  // namespace clusten

using namespace clusten;

extern "C" int clusten_attn_fwd(const void *q, const void *k, const void *v, const int64_t *nbhd_idx, const void *pack,
                                const float *bias_tab, const int32_t *bias_idx, const uint8_t *mask,
                                const void *blank_k, const void *blank_v, void *out, float *probs, float *lse,
                                int B, int H, int Nq, int Nk, int C, int M,
                                int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                                int64_t v_sb, int64_t v_sh, int64_t v_sn, int64_t o_sb, int64_t o_sh, int64_t o_sn,
                                int dtype, void *stream) {
    if (B < 0 || H <= 0 || Nq < 0 || Nk <= 0 || C <= 0 || M <= 0)
        return set_error(CLUSTEN_EINVAL, "bad sizes B=%d H=%d Nq=%d Nk=%d C=%d M=%d", B, H, Nq, Nk, C, M);
    if (!q || !k || !v || !nbhd_idx || !bias_tab || !bias_idx || !blank_k || !blank_v || !out)
        return set_error(CLUSTEN_EINVAL, "null pointer");
    if ((int64_t)B * Nq == 0) return 0;
    FusedArgs a{q, k, v, nbhd_idx, bias_tab, bias_idx, mask, blank_k, blank_v, out, probs, lse, B, H, Nq, Nk, C, M,
                q_sb, q_sh, q_sn, k_sb, k_sh, k_sn, v_sb, v_sh, v_sn, o_sb, o_sh, o_sn};
    cudaStream_t st = (cudaStream_t)stream;
    CLUSTEN_DISPATCH_DTYPE(dtype, return (launch_fused<T, false>(a, pack, st)));
    return 0;
}

// The same core with the relative-position bias computed from the token positions (posbias.cuh) instead of gathered from
// bias_tab[bias_idx]: pos_q [B,Nq,2] / pos_k [B,Nk,2] fp32 (x, y), pe_weight fp32 [H,5], pe_bias fp32 [H] or NULL.
extern "C" int clusten_attn_pos_fwd(const void *q, const void *k, const void *v, const int64_t *nbhd_idx, const void *pack,
                                    const float *pos_q, const float *pos_k, const float *pe_weight, const float *pe_bias,
                                    const uint8_t *mask, const void *blank_k, const void *blank_v, void *out, float *probs, float *lse,
                                    int B, int H, int Nq, int Nk, int C, int M,
                                    int64_t q_sb, int64_t q_sh, int64_t q_sn, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                                    int64_t v_sb, int64_t v_sh, int64_t v_sn, int64_t o_sb, int64_t o_sh, int64_t o_sn,
                                    int dtype, void *stream) {
    if (B < 0 || H <= 0 || Nq < 0 || Nk <= 0 || C <= 0 || M <= 0)
        return set_error(CLUSTEN_EINVAL, "bad sizes B=%d H=%d Nq=%d Nk=%d C=%d M=%d", B, H, Nq, Nk, C, M);
    if (!q || !k || !v || !nbhd_idx || !pos_q || !pos_k || !pe_weight || !blank_k || !blank_v || !out)
        return set_error(CLUSTEN_EINVAL, "null pointer");
    if ((reinterpret_cast<uintptr_t>(pos_q) | reinterpret_cast<uintptr_t>(pos_k)) & 7u)
        return set_error(CLUSTEN_EUNSUPPORTED, "positions must be 8-byte aligned");
    if ((int64_t)B * Nq == 0) return 0;
    FusedArgsPB a;
    static_cast<FusedArgs &>(a) = FusedArgs{q, k, v, nbhd_idx, nullptr, nullptr, mask, blank_k, blank_v, out, probs, lse, B, H, Nq, Nk, C, M,
                                            q_sb, q_sh, q_sn, k_sb, k_sh, k_sn, v_sb, v_sh, v_sn, o_sb, o_sh, o_sn};
    a.pos_q = pos_q; a.pos_k = pos_k; a.pe_w = pe_weight; a.pe_b = pe_bias;
    cudaStream_t st = (cudaStream_t)stream;
    CLUSTEN_DISPATCH_DTYPE(dtype, return (launch_fused<T, true>(a, pack, st)));
    return 0;
}
