// K3 (tensor-core path), backward of basis_mat + positional encoding + MLP_Fea.
//
// Replaces the autograd of reference basis_mat (tensoRF.py:270), positional_encoding
// (tensorBase.py:43-55) and MLPRender_Fea.forward (tensorBase.py:116-126).
//
// Two kernels, both on tcgen05 tensor cores with TMEM accumulators:
//
//  head_bwd_data_kernel  (one CTA per SM, thread t = sample row t = TMEM lane t)
//     recomputes the forward activations of a 128-sample tile (nothing but the 144
//     component values per sample was kept from the forward pass), then runs the
//     input-gradient chain
//        dh2 = (dout W3) . [h2>0]         (registers)
//        dh1 = (dh2 W2) . [h1>0]          MMA4   [128x64]  = D2[128x64]  * W2
//        din = dh1 W1                     MMA5   [128x160] = D1[128x64]  * W1
//        dfeat = PE'(din)                 (registers)
//        dcomps = dfeat Wb                MMA6   [128x144] = DF[128x32]  * Wb
//     and writes dcomps (fp32) for the VM scatter kernel. Every bf16 operand tile it
//     builds on the way (A0 comps, A1 encoded input, A2 relu(h1), A3 relu(h2), D2, D1,
//     DF, DO = dout) is also pushed to a staging area in HBM with bulk async stores
//     (TMA engine), already in the UMMA canonical layout.
//
//  head_bwd_wgrad_kernel (persistent, one CTA per SM)
//     a TMA -> tcgen05 pipeline over those staged tiles: weight gradients are
//     K = samples GEMMs, dW[m][n] = sum_s A[s][m] B[s][n], i.e. both operands are read
//     MN-major from the same tiles (no transposes), accumulated in TMEM across all
//     tiles of the CTA (464 of the 512 columns), and reduced into global memory once.
//     Bias gradients fall out of the 1.0 columns the forward operands carry.
#include "head_tc.cuh"
#include "../../include/jt_vm.h"

namespace jt {
using namespace tc;

// staging block of one tile (bytes)
constexpr int SZ_A0 = tile_bytes(TM, CT), SZ_A1 = tile_bytes(TM, K1), SZ_A2 = tile_bytes(TM, K2), SZ_A3 = SZ_A2;
constexpr int SZ_D2 = tile_bytes(TM, H_), SZ_D1 = SZ_D2, SZ_DF = tile_bytes(TM, NB), SZ_DO = tile_bytes(TM, 8);
constexpr int OFF_A0 = 0, OFF_A1 = OFF_A0 + SZ_A0, OFF_A2 = OFF_A1 + SZ_A1, OFF_A3 = OFF_A2 + SZ_A2;
constexpr int OFF_D2 = OFF_A3 + SZ_A3, OFF_D1 = OFF_D2 + SZ_D2, OFF_DF = OFF_D1 + SZ_D1, OFF_DO = OFF_DF + SZ_DF;
constexpr int STAGE_TILE_BYTES = OFF_DO + SZ_DO;        // 161792

__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

struct BwdSmem {
    static constexpr int WB = tile_bytes(NB, CT), W1 = tile_bytes(H_, K1), W2 = tile_bytes(H_, K2);
    static constexpr int W2T = tile_bytes(H_, H_), W1T = tile_bytes(K1, H_), WBT = tile_bytes(CT, NB);
    static constexpr int off_wb = 0, off_w1 = off_wb + WB, off_w2 = off_w1 + W1, off_w2t = off_w2 + W2;
    static constexpr int off_w1t = off_w2t + W2T, off_wbt = off_w1t + W1T;
    static constexpr int off_x = off_wbt + WBT;                 // 40960: A0, then A2 | A3
    static constexpr int off_y = off_x + SZ_A1;                 // 40960: A1, then D2 | D1 | DF
    static constexpr int off_do = off_y + SZ_A1;                // 2048
    static constexpr int off_w3 = off_do + SZ_DO;               // fp32 [3][64]
    static constexpr int total = off_w3 + 3 * H_ * 4;
};

__global__ void __launch_bounds__(TM) head_bwd_data_kernel(
    const float* __restrict__ comps, const float* __restrict__ dout, const int* __restrict__ aidx,
    const int* __restrict__ sidx, const float* __restrict__ rays_d, int S, int normalize_dir,
    const float* __restrict__ Wb, const float* __restrict__ W1, const float* __restrict__ b1,
    const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ W3,
    const int* __restrict__ n_dev, int n_fixed, float fprog, float vprog, float* __restrict__ dcomps,
    unsigned char* __restrict__ stage) {
    using L = BwdSmem;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int n = n_dev ? *n_dev : n_fixed;

    unsigned char* wb = smem + L::off_wb;   unsigned char* w1 = smem + L::off_w1;   unsigned char* w2 = smem + L::off_w2;
    unsigned char* w2t = smem + L::off_w2t; unsigned char* w1t = smem + L::off_w1t; unsigned char* wbt = smem + L::off_wbt;
    unsigned char* X = smem + L::off_x;     unsigned char* Y = smem + L::off_y;     unsigned char* DO = smem + L::off_do;
    unsigned char* A2 = X;                  unsigned char* A3 = X + SZ_A2;
    unsigned char* D2 = Y;                  unsigned char* D1 = Y + SZ_D2;          unsigned char* DF = Y + SZ_D2 + SZ_D1;
    float* w3s = reinterpret_cast<float*>(smem + L::off_w3);

    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    stage_weight(wb, nullptr, Wb, CT, F_, CT, NB, CT, nullptr, -1, 0);
    stage_weight(w1, nullptr, W1, IN_, H_, IN_, H_, K1, b1, BIAS1, 1);
    stage_weight(w2, nullptr, W2, H_, H_, H_, H_, K2, b2, H_, 0);
    stage_tile(w2t, nullptr, H_, H_, [&](int i, int j) { return W2[(size_t)j * H_ + i]; });             // B(n=i,k=j)
    stage_tile(w1t, nullptr, K1, H_, [&](int c, int j) { const int r = ref_col_l1(c); return r >= 0 ? W1[(size_t)j * IN_ + r] : 0.f; });
    stage_tile(wbt, nullptr, CT, NB, [&](int ic, int m) { return m < F_ ? Wb[(size_t)m * CT + ic] : 0.f; });
    for (int i = tid; i < 3 * H_; i += TM) w3s[i] = W3[i];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t T_FEAT = 0, T_H1 = 32, T_H2 = 96, T_DH1 = 160, T_DIN = 96, T_DC = 0;
    uint32_t phase = 0;
    PEMask pm;
    pm.f0 = fminf(fmaxf(fprog * 2.f - 0.f, 0.f), 1.f); pm.f1 = fminf(fmaxf(fprog * 2.f - 1.f, 0.f), 1.f);
    pm.v0 = fminf(fmaxf(vprog * 2.f - 0.f, 0.f), 1.f); pm.v1 = fminf(fmaxf(vprog * 2.f - 1.f, 0.f), 1.f);

    for (int tile = blockIdx.x; (long long)tile * TM < n; tile += gridDim.x) {
        const int row = tile * TM + tid;
        const bool live = row < n;
        unsigned char* st = stage + (size_t)tile * STAGE_TILE_BYTES;
        // ---- S0: A0 (components) and DO (upstream gradient of the head's pre-activation)
        {
            const float4* src = reinterpret_cast<const float4*>(comps + (size_t)(live ? row : 0) * CT);
#pragma unroll 3
            for (int c = 0; c < CT / 8; ++c) {
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
                if (live) { x = __ldg(src + 2 * c); y = __ldg(src + 2 * c + 1); }
                const float v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
                store_chunk(X, nullptr, TM, c, tid, v);
            }
        }
        float go[3] = {0.f, 0.f, 0.f};
        float dir[3] = {0.f, 0.f, 0.f};
        if (live) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(dout) + row);
            go[0] = g4.x; go[1] = g4.y; go[2] = g4.z;
            const int ray = sidx[aidx[row]] / S;
            dir[0] = rays_d[3 * ray]; dir[1] = rays_d[3 * ray + 1]; dir[2] = rays_d[3 * ray + 2];
            if (normalize_dir) {
                const float nn = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
                dir[0] /= nn; dir[1] /= nn; dir[2] /= nn;
            }
        }
        {
            const float v[8] = {go[0], go[1], go[2], 0.f, 0.f, 0.f, 0.f, 0.f};
            store_chunk(DO, nullptr, TM, 0, tid, v);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_gemm_kmajor<1>(tmem + T_FEAT, X, nullptr, wb, nullptr, CT, NB, NB);
            mma_commit(&bar);
            bulk_s2g(st + OFF_A0, X, SZ_A0);
            bulk_s2g(st + OFF_DO, DO, SZ_DO);
            bulk_commit();
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        // ---- S1: feat -> A1
        float feat[32];
        tmem_ld32(lane_addr + T_FEAT, feat);
#pragma unroll
        for (int c = 0; c < K1 / 8; ++c) {
            float v[8];
            encode_chunk(c, feat, dir, pm, v);
            store_chunk(Y, nullptr, TM, c, tid, v);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_gemm_kmajor<1>(tmem + T_H1, Y, nullptr, w1, nullptr, K1, H_, H_);
            mma_commit(&bar);
            bulk_s2g(st + OFF_A1, Y, SZ_A1);
            bulk_commit();
            bulk_wait_read1();                 // A0 / DO stores have finished reading X
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        __syncthreads();                       // X is free for everyone
        // ---- S2: relu(h1) -> A2 (col 64 = 1)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float h[32];
            tmem_ld32(lane_addr + T_H1 + 32 * half, h);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaxf(h[c * 8 + i], 0.f);
                store_chunk(A2, nullptr, TM, half * 4 + c, tid, v);
            }
        }
        const float one[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store_chunk(A2, nullptr, TM, 8, tid, one);
        store_chunk(A2, nullptr, TM, 9, tid, zero);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_gemm_kmajor<1>(tmem + T_H2, A2, nullptr, w2, nullptr, K2, H_, H_);
            mma_commit(&bar);
            bulk_s2g(st + OFF_A2, A2, SZ_A2);
            bulk_commit();
            bulk_wait_read1();                 // A1 store has finished reading Y
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        __syncthreads();                       // Y is free for everyone
        // ---- S3: relu(h2) -> A3 ; dh2 = (dout W3) . [h2 > 0] -> D2
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float h[32];
            tmem_ld32(lane_addr + T_H2 + 32 * half, h);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v[8], g[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = half * 32 + c * 8 + i;
                    const float x = h[c * 8 + i];
                    v[i] = fmaxf(x, 0.f);
                    g[i] = x > 0.f ? go[0] * w3s[j] + go[1] * w3s[H_ + j] + go[2] * w3s[2 * H_ + j] : 0.f;
                }
                store_chunk(A3, nullptr, TM, half * 4 + c, tid, v);
                store_chunk(D2, nullptr, TM, half * 4 + c, tid, g);
            }
        }
        store_chunk(A3, nullptr, TM, 8, tid, one);
        store_chunk(A3, nullptr, TM, 9, tid, zero);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_gemm_kmajor<1>(tmem + T_DH1, D2, nullptr, w2t, nullptr, H_, H_, H_);
            mma_commit(&bar);
            bulk_s2g(st + OFF_A3, A3, SZ_A3);
            bulk_s2g(st + OFF_D2, D2, SZ_D2);
            bulk_commit();
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        // ---- S4: dh1 = dh1_pre . [h1 > 0] -> D1
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float h[32], g[32];
            tmem_ld32(lane_addr + T_H1 + 32 * half, h);
            tmem_ld32(lane_addr + T_DH1 + 32 * half, g);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = h[c * 8 + i] > 0.f ? g[c * 8 + i] : 0.f;
                store_chunk(D1, nullptr, TM, half * 4 + c, tid, v);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_gemm_kmajor<1>(tmem + T_DIN, D1, nullptr, w1t, nullptr, H_, K1, K1);
            mma_commit(&bar);
            bulk_s2g(st + OFF_D1, D1, SZ_D1);
            bulk_commit();
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        // ---- S5: din -> dfeat (chain rule through the encoding) -> DF
        float df[32];
        {
            float g[32];
            tmem_ld32(lane_addr + T_DIN, g);
#pragma unroll
            for (int e = 0; e < 32; ++e) df[e] = e < F_ ? g[e] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {                      // 8 source elements per 32 columns
            float g[32];
            tmem_ld32(lane_addr + T_DIN + 32 * (k + 1), g);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int e = 8 * k + q;
                if (e < F_) {
                    float s, co;
                    sincosf(feat[e], &s, &co);
                    const float s2 = 2.f * s * co, c2 = 1.f - 2.f * s * s;
                    df[e] += pm.f0 * (co * g[4 * q] - s * g[4 * q + 2]) + 2.f * pm.f1 * (c2 * g[4 * q + 1] - s2 * g[4 * q + 3]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) store_chunk(DF, nullptr, TM, c, tid, df + 8 * c);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_gemm_kmajor<1>(tmem + T_DC, DF, nullptr, wbt, nullptr, NB, CT, CT);
            mma_commit(&bar);
            bulk_s2g(st + OFF_DF, DF, SZ_DF);
            bulk_commit();
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        // ---- S6: dcomps row -> global (fp32)
        {
            float4* dst = reinterpret_cast<float4*>(dcomps + (size_t)(live ? row : 0) * CT);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float g[32];
                tmem_ld32(lane_addr + T_DC + 32 * k, g);
                if (live) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) dst[8 * k + q] = make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]);
                }
            }
            float g[16];
            tmem_ld16(lane_addr + T_DC + 128, g);
            if (live) {
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[32 + q] = make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]);
            }
        }
        if (tid == 0) bulk_wait_read0();       // every staged tile has left shared memory
        tc_fence_before();
        __syncthreads();
    }
    if (tid == 0) bulk_wait0();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------ weight gradients
constexpr int WG_STAGES = 3;
constexpr int WG_A_REGION = 32768;            // the M=128 A descriptor spans 16 groups x 2048 B
constexpr int WG_STAGE_BYTES = WG_A_REGION + SZ_A1;

struct WGroup { int a_off, a_bytes, b_off, b_bytes, n, col; };
__device__ __forceinline__ WGroup wgroup(int g) {
    switch (g) {
        case 0: return {OFF_D1, SZ_D1, OFF_A1, SZ_A1, K1, 0};        // dW1 (+db1 in the bias column)
        case 1: return {OFF_D2, SZ_D2, OFF_A2, SZ_A2, K2, 160};      // dW2 (+db2 in column 64)
        case 2: return {OFF_DO, SZ_DO, OFF_A3, SZ_A3, K2, 240};      // dW3 (+db3 in column 64)
        default: return {OFF_DF, SZ_DF, OFF_A0, SZ_A0, CT, 320};     // d basis_mat
    }
}

__global__ void __launch_bounds__(TM) head_bwd_wgrad_kernel(const unsigned char* __restrict__ stage,
                                                            const int* __restrict__ n_dev, int n_fixed,
                                                            float* __restrict__ gWb, float* __restrict__ gW1,
                                                            float* __restrict__ gb1, float* __restrict__ gW2,
                                                            float* __restrict__ gb2, float* __restrict__ gW3,
                                                            float* __restrict__ gb3) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full[WG_STAGES], empty[WG_STAGES], done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = n_dev ? *n_dev : n_fixed;
    const int ntiles = (int)(((long long)n + TM - 1) / TM);
    const int my_tiles = blockIdx.x < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&done, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int items = my_tiles * 4;

    if (warp == 0 && lane == 0) {                 // ---- TMA producer
        for (int i = 0; i < items; ++i) {
            const int s = i % WG_STAGES, round = i / WG_STAGES;
            mbar_wait(&empty[s], (round & 1) ^ 1);
            const int tile = blockIdx.x + (i >> 2) * gridDim.x;
            const WGroup g = wgroup(i & 3);
            const unsigned char* src = stage + (size_t)tile * STAGE_TILE_BYTES;
            unsigned char* dst = smem + s * WG_STAGE_BYTES;
            mbar_expect_tx(&full[s], (uint32_t)(g.a_bytes + g.b_bytes));
            bulk_g2s(dst, src + g.a_off, g.a_bytes, &full[s]);
            bulk_g2s(dst + WG_A_REGION, src + g.b_off, g.b_bytes, &full[s]);
        }
    } else if (warp == 1 && lane == 0) {          // ---- MMA issuer
        for (int i = 0; i < items; ++i) {
            const int s = i % WG_STAGES, round = i / WG_STAGES;
            mbar_wait(&full[s], round & 1);
            tc_fence_after();
            const WGroup g = wgroup(i & 3);
            const uint32_t a = smem_u32(smem + s * WG_STAGE_BYTES), b = a + WG_A_REGION;
            const uint32_t idesc = idesc_bf16(128, g.n, 1, 1);
#pragma unroll
            for (int ks = 0; ks < TM / 16; ++ks)       // K = 128 sample rows, 16 per MMA = 256 B
                mma_bf16(tmem + g.col, smem_desc(a + ks * 256, 128, TM * 16), smem_desc(b + ks * 256, 128, TM * 16), idesc,
                         (i >= 4 || ks > 0) ? 1u : 0u);
            mma_commit(&empty[s]);
        }
        mma_commit(&done);
    }
    __syncwarp();
    if (items > 0) {
        mbar_wait(&done, 0);
        tc_fence_after();
        if (warp < 2) {                               // real rows are all < 64
            const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
            const int m = tid;
            for (int c0 = 0; c0 < K1; c0 += 32) {     // dW1 / db1
                float v[32];
                tmem_ld32(lane_addr + 0 + c0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int r = ref_col_l1(c0 + i);
                    if (r >= 0) atomicAdd(gW1 + (size_t)m * IN_ + r, v[i]);
                    else if (r == -2) atomicAdd(gb1 + m, v[i]);
                }
            }
            for (int c0 = 0; c0 < K2; c0 += 16) {     // dW2 / db2, dW3 / db3
                float v[16], w[16];
                tmem_ld16(lane_addr + 160 + c0, v);
                tmem_ld16(lane_addr + 240 + c0, w);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = c0 + i;
                    if (c < H_) { atomicAdd(gW2 + (size_t)m * H_ + c, v[i]); if (m < 3) atomicAdd(gW3 + (size_t)m * H_ + c, w[i]); }
                    else if (c == H_) { atomicAdd(gb2 + m, v[i]); if (m < 3) atomicAdd(gb3 + m, w[i]); }
                }
            }
            for (int c0 = 0; c0 < CT; c0 += 16) {     // d basis_mat
                float v[16];
                tmem_ld16(lane_addr + 320 + c0, v);
                if (m < F_) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(gWb + (size_t)m * CT + c0 + i, v[i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace jt

using namespace jt;

extern "C" long long jt_head_bwd_tc_stage_bytes(int n_max) {
    return (((long long)n_max + TM - 1) / TM) * STAGE_TILE_BYTES;
}

extern "C" int jt_head_bwd_tc(const float* comps, const float* dout, const int* aidx, const int* sidx,
                              const float* rays_d, int n_samples, int normalize_dir, const float* Wb, const float* W1,
                              const float* b1, const float* W2, const float* b2, const float* W3, const int* n_dev,
                              int n_max, float fea_progress, float view_progress, float* dcomps, void* stage,
                              float* gWb, float* gW1, float* gb1, float* gW2, float* gb2, float* gW3, float* gb3,
                              cudaStream_t stream) {
    JT_CHECK_ARG(comps && dout && aidx && sidx && rays_d && Wb && W1 && b1 && W2 && b2 && W3 && dcomps && stage);
    JT_CHECK_ARG(gWb && gW1 && gb1 && gW2 && gb2 && gW3 && gb3);
    JT_CHECK_ARG((reinterpret_cast<uintptr_t>(stage) & 127) == 0);
    if (n_max <= 0) return JT_OK;
    long long tiles = ((long long)n_max + TM - 1) / TM;
    int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
    if (int rc = set_smem(head_bwd_data_kernel, BwdSmem::total)) return rc;
    if (int rc = set_smem(head_bwd_wgrad_kernel, WG_STAGES * WG_STAGE_BYTES)) return rc;
    g_launches += 2;
    head_bwd_data_kernel<<<grid, TM, BwdSmem::total, stream>>>(comps, dout, aidx, sidx, rays_d, n_samples, normalize_dir, Wb,
                                                               W1, b1, W2, b2, W3, n_dev, n_max, fea_progress,
                                                               view_progress, dcomps, static_cast<unsigned char*>(stage));
    head_bwd_wgrad_kernel<<<grid, TM, WG_STAGES * WG_STAGE_BYTES, stream>>>(static_cast<const unsigned char*>(stage), n_dev,
                                                                            n_max, gWb, gW1, gb1, gW2, gb2, gW3, gb3);
    JT_RETURN_LAUNCH();
}
