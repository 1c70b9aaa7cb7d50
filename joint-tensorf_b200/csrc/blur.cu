// K5 -- separable component-wise 1-D blur of the VM planes and lines (forward and
// adjoint), staged into shared memory with the TMA engine's bulk copies.
//
// Replaces reference BAT_VMSplit.convolute_line (bateRF.py:8-19) and
// convolute_plane (bateRF.py:21-39): replicate-pad + conv1d (cross-correlation)
// along each plane axis / the line axis, per channel, 65 taps for
// c2f_kernel_size 64 (kernels.py:18), taps not normalised (kernels.py:20-21).
//
// Data is channel-last ([H][W][C]); one pass blurs along one spatial axis. A CTA
// owns a tile of TL consecutive positions of one "pencil" (a row for the W pass,
// a column for the H pass) with all C channels: it pulls the tile plus a halo of
// (ntaps-1)/2 on each side into shared memory with cp.async.bulk (one bulk copy
// for a row tile, which is contiguous; one per position for a column tile),
// waits on an mbarrier, then every thread produces 4 consecutive outputs of one
// channel from a sliding register window (2 LDS per 4 FMA).
//
// Forward:  y[i] = sum_t k[t] * x[clamp(i + t - h, 0, n-1)]            (h = ntaps/2)
// Adjoint:  dx[m] = sum_i sum_t k[t] dy[i] [clamp(i + t - h) == m]
//   interior m:  sum_t k[t] dy[m - t + h]   (zero outside [0,n))
//   m = 0     :  sum_{i<=h} dy[i] * sum_{t <= h-i} k[t]
//   m = n-1   :  sum_{i>=n-1-h} dy[i] * sum_{t >= n-1-i+h} k[t]
#include "jt_common.cuh"
#include "../../include/jt_vm.h"

namespace jt {

constexpr int TL = 128;          // outputs per CTA along the blurred axis
constexpr int MAX_TAPS = 257;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// global -> shared bulk copy executed by the TMA unit (SASS: UBLKCP); completion is
// signalled on the mbarrier as transaction bytes.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// pencils: n positions spaced `es` floats apart, each with C contiguous channels.
// pencil p starts at base + (p / inner) * outer_stride + (p % inner) * inner_stride.
struct BlurPass {
    int n;                 // length of the blurred axis
    long long es;          // element stride along the blurred axis (floats)
    int npencil, inner;
    long long outer_stride, inner_stride;
    int C;
    int ntaps;
    int adjoint;
};

__global__ void __launch_bounds__(256) blur_pass_kernel(BlurPass P, const float* __restrict__ in,
                                                        float* __restrict__ out, const float* __restrict__ taps) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ float ks[MAX_TAPS];       // taps (flipped for the adjoint)
    __shared__ float kcum[MAX_TAPS];     // adjoint only: prefix / suffix sums of the taps
    __shared__ float ksuf[MAX_TAPS];
    float* tile = reinterpret_cast<float*>(raw);

    const int h = P.ntaps >> 1;
    const int tiles_per = (P.n + TL - 1) / TL;
    const int pencil = blockIdx.x / tiles_per;
    const int t0 = (blockIdx.x - pencil * tiles_per) * TL;
    const long long base = (long long)(pencil / P.inner) * P.outer_stride + (long long)(pencil % P.inner) * P.inner_stride;
    const int lo = max(t0 - h, 0), hi = min(t0 + TL + h, P.n);       // staged range [lo, hi)
    const int span = hi - lo;
    const int C = P.C;
    const uint32_t row_bytes = (uint32_t)C * 4u;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = threadIdx.x; t < P.ntaps; t += blockDim.x) ks[t] = P.adjoint ? taps[P.ntaps - 1 - t] : taps[t];
    __syncthreads();
    if (threadIdx.x == 0) mbar_expect_tx(&bar, (uint32_t)span * row_bytes);
    __syncthreads();
    if (P.es == C) {                     // contiguous tile: one bulk copy
        if (threadIdx.x == 0) bulk_g2s(tile, in + base + (long long)lo * P.es, (uint32_t)span * row_bytes, &bar);
    } else {                             // strided tile: one bulk copy per position
        for (int i = threadIdx.x; i < span; i += blockDim.x)
            bulk_g2s(tile + (size_t)i * C, in + base + (long long)(lo + i) * P.es, row_bytes, &bar);
    }
    if (P.adjoint && threadIdx.x == 32) {     // tiny serial prefix sums, overlapped with the copy
        float s = 0.f;
        for (int t = 0; t < P.ntaps; ++t) { s += taps[t]; kcum[t] = s; }
        s = 0.f;
        for (int t = P.ntaps - 1; t >= 0; --t) { s += taps[t]; ksuf[t] = s; }
    }
    mbar_wait(&bar, 0);
    __syncthreads();

    const int ngroups = (min(TL, P.n - t0) + 3) >> 2;         // groups of 4 outputs
    for (int w = threadIdx.x; w < ngroups * C; w += blockDim.x) {
        const int c = w % C, gidx = w / C;
        const int p0 = t0 + gidx * 4;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        // value at absolute position q (forward: replicate; adjoint: zero outside)
        auto at = [&](int q) -> float {
            if (P.adjoint) return (q >= 0 && q < P.n) ? tile[(size_t)(q - lo) * C + c] : 0.f;
            q = min(max(q, 0), P.n - 1);
            return tile[(size_t)(q - lo) * C + c];
        };
        float v0 = at(p0 - h), v1 = at(p0 - h + 1), v2 = at(p0 - h + 2);
        for (int t = 0; t < P.ntaps; ++t) {
            const float v3 = at(p0 - h + t + 3);
            const float kv = ks[t];
            a0 = fmaf(kv, v0, a0); a1 = fmaf(kv, v1, a1); a2 = fmaf(kv, v2, a2); a3 = fmaf(kv, v3, a3);
            v0 = v1; v1 = v2; v2 = v3;
        }
        float r[4] = {a0, a1, a2, a3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int m = p0 + k;
            if (m >= P.n) break;
            float v = r[k];
            if (P.adjoint && (m == 0 || m == P.n - 1)) {      // fold the replicate padding back onto the edge
                v = 0.f;
                if (m == 0) {
                    for (int i = 0; i <= min(h, P.n - 1); ++i) v += tile[(size_t)(i - lo) * C + c] * kcum[h - i];
                } else {
                    for (int i = max(P.n - 1 - h, 0); i < P.n; ++i) v += tile[(size_t)(i - lo) * C + c] * ksuf[P.n - 1 - i + h];
                }
            }
            out[base + (long long)m * P.es + c] = v;
        }
    }
}

static int launch_pass(const BlurPass& P, const float* in, float* out, const float* taps, cudaStream_t stream) {
    const int h = P.ntaps >> 1;
    size_t smem = (size_t)(TL + 2 * h) * P.C * sizeof(float);
    if (smem > 200 * 1024) return JT_ERR_UNSUPPORTED;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        if (cudaFuncSetAttribute(blur_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)) != cudaSuccess)
            return JT_ERR_LAUNCH;
        configured = 200 * 1024;
    }
    const int tiles_per = (P.n + TL - 1) / TL;
    g_launches += 1;
    blur_pass_kernel<<<P.npencil * tiles_per, 256, smem, stream>>>(P, in, out, taps);
    return JT_OK;
}

}  // namespace jt

using namespace jt;

// Blur a channel-last [H][W][C] array. axes bit0: along W, bit1: along H (a line
// factor [L][C] is H=L, W=1, axes=2). tmp is required when both axes are set.
// Forward order is W then H (bateRF.py:29-36); the adjoint runs H then W.
extern "C" int jt_blur_cl(const float* in, float* out, float* tmp, int H, int W, int C, const float* taps,
                          int ntaps, int axes, int adjoint, cudaStream_t stream) {
    JT_CHECK_ARG(in && out && taps && H >= 1 && W >= 1 && C >= 4 && (C & 3) == 0);
    JT_CHECK_ARG(ntaps >= 1 && (ntaps & 1) == 1 && ntaps <= MAX_TAPS && axes >= 1 && axes <= 3);
    JT_CHECK_ARG(axes != 3 || tmp);
    JT_CHECK_ARG(in != out);
    JT_CHECK_ARG(!(axes & 1) || W >= 2);
    JT_CHECK_ARG(!(axes & 2) || H >= 2);
    BlurPass pw{W, (long long)C, H, 1, (long long)W * C, 0, C, ntaps, adjoint};           // rows
    BlurPass ph{H, (long long)W * C, W, W, 0, (long long)C, C, ntaps, adjoint};           // columns
    int rc = JT_OK;
    if (axes == 1) rc = launch_pass(pw, in, out, taps, stream);
    else if (axes == 2) rc = launch_pass(ph, in, out, taps, stream);
    else if (!adjoint) { rc = launch_pass(pw, in, tmp, taps, stream); if (!rc) rc = launch_pass(ph, tmp, out, taps, stream); }
    else { rc = launch_pass(ph, in, tmp, taps, stream); if (!rc) rc = launch_pass(pw, tmp, out, taps, stream); }
    if (rc) return rc;
    JT_RETURN_LAUNCH();
}
