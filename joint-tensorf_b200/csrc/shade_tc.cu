// K3 (tensor-core path) -- basis_mat projection + positional encoding + MLP_Fea
// shading head as ONE kernel on the 5th-generation tensor cores (tcgen05.mma,
// accumulators in TMEM), per tile of 128 appearance samples.
//
// Replaces reference basis_mat (tensoRF.py:156,270 / bateRF.py:130),
// positional_encoding (tensorBase.py:43-55) and MLPRender_Fea.forward
// (tensorBase.py:116-126) for app_dim 27 / hidden 64 / pe 2 (the Blender configs).
//
// Per tile: thread t owns sample row t (and TMEM lane t).
//   comps row (144 fp32, global) -> bf16 tile A0 (smem)        MMA1: feat  = A0 * Wb^T   [128x32]
//   feat (tcgen05.ld) -> PE -> bf16 tile A1 [128x160]           MMA2: h1    = A1 * W1^T   [128x64]
//   relu(h1) -> bf16 tile A2 [128x80]                           MMA3: h2    = A2 * W2^T   [128x64]
//   relu(h2) -> layer 3 (3x64 dot products in registers) -> sigmoid -> rgb
// Biases ride in the GEMMs: column 150 of A1 and column 64 of A2 are 1.0 and the
// staged weights carry b1 / b2 in those columns.
// SPLIT = 2 stores every operand as hi + lo bf16 terms and issues 3 MMAs
// (hi*hi + hi*lo + lo*hi): ~2^-16 relative product error, i.e. fp32-class results
// from bf16 tensor-core throughput. SPLIT = 1 is plain bf16.
#include <stdlib.h>
#include "head_tc.cuh"
#include "../../include/jt_vm.h"

namespace jt {
using namespace tc;

// ------------------------------------------------------------------ self test of the MMA plumbing
// mode 0: D[128][N] = A[128][K] * B[N][K]^T          (K-major operands)
// mode 1: D[m][n]   = sum_s X[s][m] * Y[s][n]        (MN-major operands; X [128][Ma], Y [128][N], K = 128 rows)
// mode 2: mode 0 with both operands stored as fp16
__global__ void __launch_bounds__(128) tc_selftest_kernel(int mode, const float* __restrict__ A, int lda,
                                                          const float* __restrict__ B, int ldb, float* __restrict__ D,
                                                          int K, int N, int Ma) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char* ta = smem;                       // A / X tile: 128 rows x 128 cols max
    unsigned char* tb = smem + tile_bytes(128, 256);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (mode == 0 || mode == 2) {
        for (int c = 0; c < K / 8; ++c) {           // thread = row of A
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = A[(size_t)tid * lda + c * 8 + i];
            if (mode == 2) store_chunk_f16(ta, 128, c, tid, v); else store_chunk(ta, nullptr, 128, c, tid, v);
        }
        if (mode == 2) {
            for (int idx = tid; idx < N * (K / 8); idx += 128) {
                const int chunk = idx / N, n = idx - chunk * N;
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = B[(size_t)n * ldb + chunk * 8 + i];
                store_chunk_f16(tb, N, chunk, n, v);
            }
        } else {
            stage_weight(tb, nullptr, B, ldb, N, K, N, K, nullptr, -1, 0);
        }
    } else {
        for (int c = 0; c < 128 / 8; ++c) {         // X tile padded to 128 columns with zeros
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (c * 8 + i < Ma) ? A[(size_t)tid * lda + c * 8 + i] : 0.f;
            store_chunk(ta, nullptr, 128, c, tid, v);
        }
        for (int c = 0; c < N / 8; ++c) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = B[(size_t)tid * ldb + c * 8 + i];
            store_chunk(tb, nullptr, 128, c, tid, v);
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        if (mode == 0) {
            issue_gemm_kmajor<1>(tmem, ta, nullptr, tb, nullptr, K, N, N);
        } else if (mode == 2) {
            const uint32_t idesc = idesc_f16(128, N, 0, 0, 0, 0);
            for (int ks = 0; ks < K / 16; ++ks)
                mma_bf16(tmem, smem_desc(smem_u32(ta) + ks * 2 * 128 * 16, 128 * 16, 128),
                         smem_desc(smem_u32(tb) + ks * 2 * N * 16, N * 16, 128), idesc, ks > 0);
        } else {
            const uint32_t idesc = idesc_bf16(128, N, 1, 1);
            for (int ks = 0; ks < 128 / 16; ++ks) {             // K = sample rows, 16 per step = 256 B
                const uint64_t ad = smem_desc(smem_u32(ta) + ks * 256, 128, 128 * 16);
                const uint64_t bd = smem_desc(smem_u32(tb) + ks * 256, 128, 128 * 16);
                mma_bf16(tmem, ad, bd, idesc, ks > 0);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(lane_addr + c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) if (c0 + i < N) D[(size_t)tid * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------ fused head forward
// `stage` != nullptr (training): the bf16 (hi) operand tiles A0 comps, A1 encoded input,
// A2 relu(h1), A3 relu(h2) of every tile are pushed to HBM with bulk async stores, in the
// UMMA canonical layout, for the backward kernels (exact relu masks + the B operands of
// the weight-gradient GEMMs). Two hi regions alternate so a store can drain while the
// next operand is being built.
template <int SPLIT, bool SAVE>
struct FwdSmem {
    static constexpr int WB = tile_bytes(NB, CT), W1 = tile_bytes(H_, K1), W2 = tile_bytes(H_, K2);
    static constexpr int A = tile_bytes(TM, K1);
    static constexpr int off_wb = 0, off_w1 = off_wb + SPLIT * WB, off_w2 = off_w1 + SPLIT * W1;
    static constexpr int off_hi0 = off_w2 + SPLIT * W2;
    static constexpr int off_hi1 = SAVE ? off_hi0 + A : off_hi0;
    static constexpr int off_lo = off_hi1 + A;
    static constexpr int off_w3 = off_lo + (SPLIT == 2 ? A : 0);      // fp32 [3][64] + b3[3]
    static constexpr int total = off_w3 + (3 * H_ + 4) * 4;
};

// Two threads share a sample row: thread (r, hh) with r = tid & 127 (= TMEM lane) and
// hh = tid >> 7 owns one half of the columns of every operand tile, so the per-row
// instruction chains are half as long and each SM sub-partition has two warps to switch
// between while one waits on memory / TMEM.
constexpr int NT = 2 * TM;

template <int SPLIT, bool SAVE>
__global__ void __launch_bounds__(NT) head_fwd_tc_kernel(const float* __restrict__ comps, const int* __restrict__ aidx,
                                                         const int* __restrict__ sidx, const float* __restrict__ rays_d,
                                                         int S, int normalize_dir, const float* __restrict__ Wb,
                                                         const float* __restrict__ W1, const float* __restrict__ b1,
                                                         const float* __restrict__ W2, const float* __restrict__ b2,
                                                         const float* __restrict__ W3, const float* __restrict__ b3,
                                                         const int* __restrict__ n_dev, int n_fixed, float fprog,
                                                         float vprog, float* __restrict__ rgb, float* __restrict__ feat_out,
                                                         unsigned char* __restrict__ stage, int ahead) {
    using L = FwdSmem<SPLIT, SAVE>;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float part[TM][3];                 // layer-3 partial sums of the hh = 1 half
    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & (TM - 1), hh = tid >> 7;
    const int n = n_dev ? *n_dev : n_fixed;

    unsigned char* wb_hi = smem + L::off_wb;  unsigned char* wb_lo = SPLIT == 2 ? wb_hi + L::WB : nullptr;
    unsigned char* w1_hi = smem + L::off_w1;  unsigned char* w1_lo = SPLIT == 2 ? w1_hi + L::W1 : nullptr;
    unsigned char* w2_hi = smem + L::off_w2;  unsigned char* w2_lo = SPLIT == 2 ? w2_hi + L::W2 : nullptr;
    unsigned char* hi0 = smem + L::off_hi0;   unsigned char* hi1 = smem + L::off_hi1;
    unsigned char* lo = SPLIT == 2 ? smem + L::off_lo : nullptr;
    float* w3s = reinterpret_cast<float*>(smem + L::off_w3);

    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    stage_weight(wb_hi, wb_lo, Wb, CT, F_, CT, NB, CT, nullptr, -1, 0);
    stage_weight(w1_hi, w1_lo, W1, IN_, H_, IN_, H_, K1, b1, BIAS1, 1);
    stage_weight(w2_hi, w2_lo, W2, H_, H_, H_, H_, K2, b2, H_, 0);
    for (int i = tid; i < 3 * H_ + 3; i += NT) w3s[i] = i < 3 * H_ ? W3[i] : b3[i - 3 * H_];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t T_FEAT = 0, T_H1 = 32, T_H2 = 96;
    uint32_t phase = 0;
    PEMask pm;
    pm.f0 = fminf(fmaxf(fprog * 2.f - 0.f, 0.f), 1.f); pm.f1 = fminf(fmaxf(fprog * 2.f - 1.f, 0.f), 1.f);
    pm.v0 = fminf(fmaxf(vprog * 2.f - 0.f, 0.f), 1.f); pm.v1 = fminf(fmaxf(vprog * 2.f - 1.f, 0.f), 1.f);

    // software prefetch: the 72 KB component tile (and the view direction) of the NEXT tile is
    // loaded into registers while the current tile runs its three GEMM phases, so the DRAM
    // latency of the only large global read of this kernel is off the critical path.
    float4 nx[CT / 8];                           // this thread's half row: 18 float4
    float pd[3];
    auto prefetch = [&](int t) {
        const long long rw = (long long)t * TM + r;
        const bool lv = rw < n;
        const float4* src = reinterpret_cast<const float4*>(comps + (size_t)(lv ? rw : 0) * CT) + 18 * hh;
#pragma unroll
        for (int c = 0; c < CT / 8; ++c) nx[c] = lv ? __ldcs(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        pd[0] = pd[1] = pd[2] = 0.f;
        if (lv) {
            const int ray = sidx[aidx[rw]] / S;
            pd[0] = rays_d[3 * ray]; pd[1] = rays_d[3 * ray + 1]; pd[2] = rays_d[3 * ray + 2];
        }
    };
    if (ahead) prefetch(blockIdx.x);

    for (int tile = blockIdx.x; (long long)tile * TM < n; tile += gridDim.x) {
        const int row = tile * TM + r;
        const bool live = row < n;
        unsigned char* st = SAVE ? stage + (size_t)tile * STAGE_TILE_BYTES : nullptr;
        // ---- A0: component row -> bf16 (hi0); this thread converts chunks [9 hh, 9 hh + 9)
        if (!ahead) prefetch(tile);
#pragma unroll
        for (int c = 0; c < CT / 16; ++c) {
            const float4 x = nx[2 * c], y = nx[2 * c + 1];
            const float v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
            store_chunk(hi0, lo, TM, 9 * hh + c, r, v);
        }
        float dir[3] = {pd[0], pd[1], pd[2]};
        if (normalize_dir && live) {
            const float nn = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
            dir[0] /= nn; dir[1] /= nn; dir[2] /= nn;
        }
        if (ahead) prefetch(tile + gridDim.x);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_gemm_kmajor<SPLIT>(tmem + T_FEAT, hi0, lo, wb_hi, wb_lo, CT, NB, NB);
            mma_commit(&bar);
            if (SAVE) {
                bulk_s2g(st + OFF_A0, hi0, SZ_A0);
                bulk_commit();
                bulk_wait_read1();             // previous tile's A3 store has left hi1
            }
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        if (SAVE) __syncthreads();
        // ---- feat -> encoded input A1 (hi1); chunks [10 hh, 10 hh + 10)
        float feat[32];
        tmem_ld32(lane_addr + T_FEAT, feat);
        if (feat_out && live && hh == 0) {
#pragma unroll
            for (int q = 0; q < 7; ++q)
                *reinterpret_cast<float4*>(feat_out + (size_t)row * 28 + 4 * q) =
                    make_float4(feat[4 * q], feat[4 * q + 1], feat[4 * q + 2], q == 6 ? 0.f : feat[4 * q + 3]);
        }
        if (hh == 0) {
#pragma unroll
            for (int c = 0; c < K1 / 16; ++c) {
                float v[8];
                encode_chunk(c, feat, dir, pm, v);
                store_chunk(hi1, lo, TM, c, r, v);
            }
        } else {
#pragma unroll
            for (int c = K1 / 16; c < K1 / 8; ++c) {
                float v[8];
                encode_chunk(c, feat, dir, pm, v);
                store_chunk(hi1, lo, TM, c, r, v);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_gemm_kmajor<SPLIT>(tmem + T_H1, hi1, lo, w1_hi, w1_lo, K1, H_, H_);
            mma_commit(&bar);
            if (SAVE) {
                bulk_s2g(st + OFF_A1, hi1, SZ_A1);
                bulk_commit();
                bulk_wait_read1();             // A0 store has left hi0
            }
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        if (SAVE) __syncthreads();
        // ---- relu(h1) -> A2 (hi0; col 64 = 1 carries b2); columns [32 hh, 32 hh + 32)
        {
            float h[32];
            tmem_ld32(lane_addr + T_H1 + 32 * hh, h);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaxf(h[c * 8 + i], 0.f);
                store_chunk(hi0, lo, TM, hh * 4 + c, r, v);
            }
            const float pad[8] = {hh == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            store_chunk(hi0, lo, TM, 8 + hh, r, pad);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_gemm_kmajor<SPLIT>(tmem + T_H2, hi0, lo, w2_hi, w2_lo, K2, H_, H_);
            mma_commit(&bar);
            if (SAVE) {
                bulk_s2g(st + OFF_A2, hi0, SZ_A2);
                bulk_commit();
                bulk_wait_read1();             // A1 store has left hi1
            }
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        if (SAVE) __syncthreads();
        // ---- relu(h2) -> layer 3 partial sums (and A3 tile into hi1 when saving)
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
        {
            float h[32];
            tmem_ld32(lane_addr + T_H2 + 32 * hh, h);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                h[i] = fmaxf(h[i], 0.f);
                o0 = fmaf(h[i], w3s[hh * 32 + i], o0);
                o1 = fmaf(h[i], w3s[H_ + hh * 32 + i], o1);
                o2 = fmaf(h[i], w3s[2 * H_ + hh * 32 + i], o2);
            }
            if (SAVE) {
#pragma unroll
                for (int c = 0; c < 4; ++c) store_chunk(hi1, nullptr, TM, hh * 4 + c, r, h + 8 * c);
                const float pad[8] = {hh == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                store_chunk(hi1, nullptr, TM, 8 + hh, r, pad);
                fence_async_smem();
            }
        }
        if (hh == 1) { part[r][0] = o0; part[r][1] = o1; part[r][2] = o2; }
        // all tcgen05.ld of this tile are complete (wait::ld) before the next tile's MMAs overwrite TMEM
        tc_fence_before();
        __syncthreads();
        if (hh == 0 && live) {
            o0 += part[r][0] + w3s[3 * H_]; o1 += part[r][1] + w3s[3 * H_ + 1]; o2 += part[r][2] + w3s[3 * H_ + 2];
            __stcs(reinterpret_cast<float4*>(rgb + 4 * (size_t)row),
                   make_float4(1.f / (1.f + expf(-o0)), 1.f / (1.f + expf(-o1)), 1.f / (1.f + expf(-o2)), 0.f));
        }
        if (SAVE && tid == 0) {
            bulk_s2g(st + OFF_A3, hi1, SZ_A3);
            bulk_commit();
            bulk_wait_read1();                 // A2 store has left hi0 (next tile's A0 goes there)
        }
        __syncthreads();                       // `part` and hi0 are free for the next tile
    }
    if (SAVE && tid == 0) bulk_wait0();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace jt

using namespace jt;

extern "C" int jt_tc_selftest(int mode, const float* A, int lda, const float* B, int ldb, float* D, int K, int N,
                              int Ma, cudaStream_t stream) {
    JT_CHECK_ARG(A && B && D && mode >= 0 && mode <= 2);
    JT_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256 && Ma <= 128);
    const int smem = tile_bytes(128, 256) * 2;
    if (int rc = set_smem(tc_selftest_kernel, smem)) return rc;
    g_launches += 1;
    tc_selftest_kernel<<<1, 128, smem, stream>>>(mode, A, lda, B, ldb, D, K, N, Ma);
    JT_RETURN_LAUNCH();
}

extern "C" int jt_head_fwd_tc(int split, const float* comps, const int* aidx, const int* sidx, const float* rays_d,
                              int n_samples, int normalize_dir, const float* Wb, const float* W1, const float* b1,
                              const float* W2, const float* b2, const float* W3, const float* b3, const int* n_dev,
                              int n_max, float fea_progress, float view_progress, float* rgb, float* feat_out,
                              void* stage, cudaStream_t stream) {
    JT_CHECK_ARG(comps && aidx && sidx && rays_d && Wb && W1 && b1 && W2 && b2 && W3 && b3 && rgb);
    JT_CHECK_ARG(split == 1 || split == 2);
    JT_CHECK_ARG((reinterpret_cast<uintptr_t>(stage) & 127) == 0);
    if (n_max <= 0) return JT_OK;
    long long tiles = ((long long)n_max + TM - 1) / TM;
    unsigned char* st = static_cast<unsigned char*>(stage);
    static const char* env_pf = getenv("JT_TC_PREFETCH");      // tuning: 0 loads the component tile in place
    const int ahead = env_pf ? atoi(env_pf) : 1;
    g_launches += 1;
#define JT_LAUNCH_FWD(SP, SV, PER_SM)                                                                                   \
    {                                                                                                                   \
        const int smem = FwdSmem<SP, SV>::total;                                                                        \
        if (int rc = set_smem(head_fwd_tc_kernel<SP, SV>, smem)) return rc;                                             \
        int grid = (int)(tiles < (PER_SM) * kNumSMs ? tiles : (PER_SM) * kNumSMs);                                      \
        head_fwd_tc_kernel<SP, SV><<<grid, NT, smem, stream>>>(comps, aidx, sidx, rays_d, n_samples, normalize_dir, Wb, \
                                                               W1, b1, W2, b2, W3, b3, n_dev, n_max, fea_progress,      \
                                                               view_progress, rgb, feat_out, st, ahead);                \
    }
    if (split == 1 && !st) JT_LAUNCH_FWD(1, false, 2)
    else if (split == 1) JT_LAUNCH_FWD(1, true, 1)
    else if (!st) JT_LAUNCH_FWD(2, false, 1)
    else JT_LAUNCH_FWD(2, true, 1)
#undef JT_LAUNCH_FWD
    JT_RETURN_LAUNCH();
}
