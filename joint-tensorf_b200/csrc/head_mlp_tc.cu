// K3 (tensor-core path) -- positional encoding + MLP_Fea shading head, per tile of 128
// appearance samples, fed by the feature/direction rows app_basis_fwd_kernel writes.
//
// Replaces reference positional_encoding (tensorBase.py:43-55) and MLPRender_Fea.forward
// (tensorBase.py:116-126) for app_dim 27 / hidden 64 / pe 2 (the Blender configs).
//
// Per tile (four threads per sample row, each owning a quarter of the columns):
//   feat/dir row (prefetched one tile ahead)        -> PE -> bf16 tile A1 [128x160]
//                                                    MMA: h1 = A1 * W1^T   [128x64] (TMEM)
//   relu(h1) -> bf16 tile A2 [128x80]                MMA: h2 = A2 * W2^T   [128x64] (TMEM)
//   relu(h2) -> layer 3 (3x64 dot products in registers) -> sigmoid -> rgb
// Biases ride in the GEMMs (a 1.0 column in A1 / A2, see head_tc.cuh). SPLIT = 2 stores
// every operand as hi + lo bf16 terms and issues 3 MMAs per product (fp32-class results);
// SPLIT = 1 is plain bf16. Training (SAVE) pushes the hi tiles A1, A2 and A3 = relu(h2) to
// the HBM staging area with bulk async stores for the backward kernels; each tile buffer
// is only re-written after the store issued three stores earlier has drained, so the
// stores overlap the next phases instead of stalling them.
#include "head_tc.cuh"
#include "../../include/jt_vm.h"

namespace jt {
using namespace tc;

template <int SPLIT, bool SAVE>
struct HSmem {
    static constexpr int W1 = tile_bytes(H_, K1), W2 = tile_bytes(H_, K2);
    static constexpr int A1 = tile_bytes(TM, K1), A2 = tile_bytes(TM, K2);
    static constexpr int NT = SPLIT == 2 ? 2 : 1;                      // operand terms kept in shared memory
    static constexpr int off_w1 = 0, off_w2 = off_w1 + NT * W1;
    static constexpr int off_a1_hi = off_w2 + NT * W2, off_a1_lo = off_a1_hi + A1;
    static constexpr int off_a2_hi = off_a1_lo + (SPLIT == 2 ? A1 : 0), off_a2_lo = off_a2_hi + A2;
    static constexpr int off_a3 = off_a2_lo + (SPLIT == 2 ? A2 : 0);
    static constexpr int off_w3 = off_a3 + (SAVE ? A2 : 0);            // fp32 [3][64] + b3[3]
    static constexpr int total = off_w3 + (3 * H_ + 4) * 4;
};

constexpr int HQ = 4;           // threads per sample row (column quarters): 16 warps keep the SIMT phases latency-tolerant
constexpr int HT = HQ * TM;     // thread (r, hq): r = tid & 127 = tile row = TMEM lane, hq = tid >> 7

// chunks of the encoded input owned by column-quarter hq: the 15 PE chunks (two sin/cos
// elements each) are spread 3/4/4/4, the 4 raw chunks go with the short PE share, the
// zero chunk 19 with the last
__device__ __forceinline__ void chunk_range(int hq, int& c0, int& c1) {
    c0 = hq == 0 ? 0 : 3 + 4 * hq;
    c1 = hq == 3 ? 20 : 7 + 4 * hq;
}

template <int SPLIT, bool SAVE>
__global__ void __launch_bounds__(HT) head_mlp_fwd_kernel(const float* __restrict__ featdir, const float* __restrict__ W1,
                                                          const float* __restrict__ b1, const float* __restrict__ W2,
                                                          const float* __restrict__ b2, const float* __restrict__ W3,
                                                          const float* __restrict__ b3, const int* __restrict__ n_dev,
                                                          int n_fixed, float fprog, float vprog, float* __restrict__ rgb,
                                                          unsigned char* __restrict__ stage) {
    using L = HSmem<SPLIT, SAVE>;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float part[2][HQ - 1][TM][3];      // layer-3 partial sums of column quarters 1..3, double-buffered
    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & (TM - 1), hq = tid >> 7;
    const int n = n_dev ? *n_dev : n_fixed;

    unsigned char* w1_hi = smem + L::off_w1;  unsigned char* w1_lo = SPLIT == 2 ? w1_hi + L::W1 : nullptr;
    unsigned char* w2_hi = smem + L::off_w2;  unsigned char* w2_lo = SPLIT == 2 ? w2_hi + L::W2 : nullptr;
    unsigned char* a1_hi = smem + L::off_a1_hi; unsigned char* a1_lo = SPLIT == 2 ? smem + L::off_a1_lo : nullptr;
    unsigned char* a2_hi = smem + L::off_a2_hi; unsigned char* a2_lo = SPLIT == 2 ? smem + L::off_a2_lo : nullptr;
    unsigned char* a3_hi = smem + L::off_a3;
    float* w3s = reinterpret_cast<float*>(smem + L::off_w3);

    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 128);
    constexpr bool F16 = SPLIT == 3;
    static_assert(!(F16 && SAVE), "the fp16 tiles are an inference-only format");
    stage_weight<F16>(w1_hi, w1_lo, W1, IN_, H_, IN_, H_, K1, b1, BIAS1, 1);
    stage_weight<F16>(w2_hi, w2_lo, W2, H_, H_, H_, H_, K2, b2, H_, 0);
    for (int i = tid; i < 3 * H_ + 3; i += HT) w3s[i] = i < 3 * H_ ? W3[i] : b3[i - 3 * H_];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t T_H1 = 0, T_H2 = 64;
    uint32_t phase = 0;
    PEMask pm;
    pm.f0 = fminf(fmaxf(fprog * 2.f - 0.f, 0.f), 1.f); pm.f1 = fminf(fmaxf(fprog * 2.f - 1.f, 0.f), 1.f);
    pm.v0 = fminf(fmaxf(vprog * 2.f - 0.f, 0.f), 1.f); pm.v1 = fminf(fmaxf(vprog * 2.f - 1.f, 0.f), 1.f);
    int ec0, ec1, pb = 0;
    chunk_range(hq, ec0, ec1);

    // the feat/dir row of the NEXT tile is fetched while the current tile computes. Only the 30
    // floats that are used are loaded: a dead destination register would be recycled by the
    // compiler while its load is still in flight and stall the consumer (seen in ncu).
    float4 nx[6];
    float2 nf2, nd2;
    float nf1, nd1;
    auto prefetch = [&](long long t) {
        const long long rw = t * TM + r;
        const bool lv = rw < n;
        const float* src = featdir + (size_t)(lv ? rw : 0) * FD;
#pragma unroll
        for (int c = 0; c < 6; ++c) nx[c] = lv ? __ldcs(reinterpret_cast<const float4*>(src) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        nf2 = lv ? __ldcs(reinterpret_cast<const float2*>(src + 24)) : make_float2(0.f, 0.f);
        nf1 = lv ? __ldcs(src + 26) : 0.f;
        nd2 = lv ? __ldcs(reinterpret_cast<const float2*>(src + 28)) : make_float2(0.f, 0.f);
        nd1 = lv ? __ldcs(src + 30) : 0.f;
    };
    prefetch(blockIdx.x);

    for (int tile = blockIdx.x; (long long)tile * TM < n; tile += gridDim.x) {
        const int row = tile * TM + r;
        const bool live = row < n;
        unsigned char* st = SAVE ? stage + (size_t)tile * STAGE_TILE_BYTES : nullptr;
        float feat[32];
#pragma unroll
        for (int c = 0; c < 6; ++c) { feat[4 * c] = nx[c].x; feat[4 * c + 1] = nx[c].y; feat[4 * c + 2] = nx[c].z; feat[4 * c + 3] = nx[c].w; }
        feat[24] = nf2.x; feat[25] = nf2.y; feat[26] = nf1;
        feat[27] = feat[28] = feat[29] = feat[30] = feat[31] = 0.f;
        const float dir[3] = {nd2.x, nd2.y, nd1};
        prefetch((long long)tile + gridDim.x);
        // ---- encoded input A1; this thread: chunks [ec0, ec1)
#pragma unroll
        for (int c = 0; c < K1 / 8; ++c) {
            if (c >= ec0 && c < ec1) {
                float v[8];
                encode_chunk<true>(c, feat, dir, pm, v);
                store_chunk_t<F16>(a1_hi, a1_lo, TM, c, r, v);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // Bulk-store bookkeeping (training): groups are committed in the order A1(t), A2(t), A3(t),
        // A1(t+1), ...; before a buffer is re-written the issuing thread waits until the store that
        // last read it has drained, and only then signals the MMA barrier every thread waits on.
        if (tid == 0) {
            tc_fence_after();
            issue_gemm_kmajor<SPLIT>(tmem + T_H1, a1_hi, a1_lo, w1_hi, w1_lo, K1, H_, H_);
            if (SAVE) {
                bulk_s2g(st + OFF_A1, a1_hi, SZ_A1);
                bulk_commit();
                bulk_wait_read2();             // A2(t-1) has left a2_hi
            }
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        // ---- relu(h1) -> A2 (col 64 = 1 carries b2); columns [16 hq, 16 hq + 16)
        {
            float h[16];
            tmem_ld16(lane_addr + T_H1 + 16 * hq, h);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaxf(h[c * 8 + i], 0.f);
                store_chunk_t<F16>(a2_hi, a2_lo, TM, hq * 2 + c, r, v);
            }
            if (hq < 2) {
                const float pad[8] = {hq == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                store_chunk_t<F16>(a2_hi, a2_lo, TM, 8 + hq, r, pad);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_gemm_kmajor<SPLIT>(tmem + T_H2, a2_hi, a2_lo, w2_hi, w2_lo, K2, H_, H_);
            if (SAVE) {
                bulk_s2g(st + OFF_A2, a2_hi, SZ_A2);
                bulk_commit();
                bulk_wait_read2();             // A3(t-1) has left a3_hi
            }
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        // ---- relu(h2) -> layer 3 partial sums (and the A3 tile when saving)
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
        {
            float h[16];
            tmem_ld16(lane_addr + T_H2 + 16 * hq, h);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                h[i] = fmaxf(h[i], 0.f);
                o0 = fmaf(h[i], w3s[hq * 16 + i], o0);
                o1 = fmaf(h[i], w3s[H_ + hq * 16 + i], o1);
                o2 = fmaf(h[i], w3s[2 * H_ + hq * 16 + i], o2);
            }
            if (SAVE) {
#pragma unroll
                for (int c = 0; c < 2; ++c) store_chunk(a3_hi, nullptr, TM, hq * 2 + c, r, h + 8 * c);
                if (hq < 2) {
                    const float pad[8] = {hq == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    store_chunk(a3_hi, nullptr, TM, 8 + hq, r, pad);
                }
                fence_async_smem();
            }
        }
        if (hq > 0) { part[pb][hq - 1][r][0] = o0; part[pb][hq - 1][r][1] = o1; part[pb][hq - 1][r][2] = o2; }
        if (SAVE && tid == 0) bulk_wait_read1();    // A1(t) has left a1_hi: the next tile's encode may overwrite it
        // all tcgen05.ld of this tile are complete (wait::ld) before the next tile's MMAs overwrite TMEM
        tc_fence_before();
        __syncthreads();
        if (SAVE && tid == 0) {
            bulk_s2g(st + OFF_A3, a3_hi, SZ_A3);
            bulk_commit();
        }
        if (hq == 0 && live) {
#pragma unroll
            for (int k = 0; k < HQ - 1; ++k) { o0 += part[pb][k][r][0]; o1 += part[pb][k][r][1]; o2 += part[pb][k][r][2]; }
            o0 += w3s[3 * H_]; o1 += w3s[3 * H_ + 1]; o2 += w3s[3 * H_ + 2];
            __stcs(reinterpret_cast<float4*>(rgb + 4 * (size_t)row),
                   make_float4(1.f / (1.f + expf(-o0)), 1.f / (1.f + expf(-o1)), 1.f / (1.f + expf(-o2)), 0.f));
        }
        pb ^= 1;                               // `part` is double-buffered: no barrier at the end of the tile
    }
    if (SAVE && tid == 0) bulk_wait0();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace jt

using namespace jt;

extern "C" int jt_head_mlp_fwd_tc(int split, const float* featdir, const float* W1, const float* b1, const float* W2,
                                  const float* b2, const float* W3, const float* b3, const int* n_dev, int n_max,
                                  float fea_progress, float view_progress, float* rgb, void* stage,
                                  cudaStream_t stream) {
    JT_CHECK_ARG(featdir && W1 && b1 && W2 && b2 && W3 && b3 && rgb);
    JT_CHECK_ARG(split >= 1 && split <= 3);
    JT_CHECK_ARG(split != 3 || !stage);                  // fp16 operands: inference only (no staged tiles)
    JT_CHECK_ARG((reinterpret_cast<uintptr_t>(stage) & 127) == 0);
    if (n_max <= 0) return JT_OK;
    long long tiles = ((long long)n_max + TM - 1) / TM;
    unsigned char* st = static_cast<unsigned char*>(stage);
    g_launches += 1;
#define JT_LAUNCH_H(SP, SV, PER_SM)                                                                                     \
    {                                                                                                                   \
        const int smem = HSmem<SP, SV>::total;                                                                          \
        if (int rc = set_smem(head_mlp_fwd_kernel<SP, SV>, smem)) return rc;                                            \
        int grid = (int)(tiles < (PER_SM) * kNumSMs ? tiles : (PER_SM) * kNumSMs);                                      \
        head_mlp_fwd_kernel<SP, SV><<<grid, HT, smem, stream>>>(featdir, W1, b1, W2, b2, W3, b3, n_dev, n_max,          \
                                                                fea_progress, view_progress, rgb, st);                  \
    }
    if (split == 3) JT_LAUNCH_H(3, false, 1)      // 87 registers: one CTA per SM (capping at 64 for two spills: 0.59 vs 0.50 ms)
    else if (split == 1 && !st) JT_LAUNCH_H(1, false, 2)
    else if (split == 1) JT_LAUNCH_H(1, true, 2)
    else if (!st) JT_LAUNCH_H(2, false, 1)
    else JT_LAUNCH_H(2, true, 1)
#undef JT_LAUNCH_H
    JT_RETURN_LAUNCH();
}
