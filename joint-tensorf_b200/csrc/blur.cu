// K5 -- separable component-wise 1-D blur of the VM planes and lines (forward and
// adjoint), staged into shared memory with the TMA engine's bulk copies.
//
// Replaces reference BAT_VMSplit.convolute_line (bateRF.py:8-19) and
// convolute_plane (bateRF.py:21-39): replicate-pad + conv1d (cross-correlation)
// along each plane axis / the line axis, per channel, 65 taps for
// c2f_kernel_size 64 (kernels.py:18), taps not normalised (kernels.py:20-21).
//
// Data is channel-last ([H][W][C]); one pass blurs along one spatial axis. The stencil is
// independent per channel, so for the pass along H the x position is folded into the channel
// axis: a CTA owns TL consecutive positions of one "pencil" = a row (W pass, C channels) or a
// strip of WX adjacent columns (H pass, WX*C contiguous floats per position).
//
// Per CTA:
//   1. the tile plus a halo of (ntaps-1)/2 positions on each side is pulled into shared memory by
//      cp.async.bulk (SASS UBLKCP; one copy per group of 8 contiguous positions, or one per
//      position when strided / at a clamped edge), completion on an mbarrier. Out-of-range
//      positions are materialised in the tile (forward: the replicated edge row, adjoint: zeros),
//      so the inner loop has no bounds logic;
//   2. a thread owns 8 consecutive outputs x 4 channels (32 fp32 accumulators). The input window
//      slides through two register halves of 8 float4 each: per block of 8 taps it issues 8
//      LDS.128 and 256 FMAs whose tap operand comes straight from the constant bank (the taps
//      travel in the kernel parameters). Position groups are padded in shared memory so that the
//      32 lanes of a warp always read 32 distinct 16-byte bank groups.
// Arithmetic per pass: 2*ntaps flop per element against 8 B of HBM traffic (65 taps: 16 flop/B),
// i.e. the pass is HBM-bound only if the FMA pipe is kept busy -- hence the register tiling.
//
// Forward:  y[i] = sum_t k[t] * x[clamp(i + t - h, 0, n-1)]            (h = ntaps/2)
// Adjoint:  dx[m] = sum_i sum_t k[t] dy[i] [clamp(i + t - h) == m]
//   interior m:  sum_t k[t] dy[m - t + h]   (zero outside [0,n))
//   m = 0     :  sum_{i<=h} dy[i] * sum_{t <= h-i} k[t]
//   m = n-1   :  sum_{i>=n-1-h} dy[i] * sum_{t >= n-1-i+h} k[t]
#include "jt_common.cuh"
#include "../../include/jt_vm.h"

namespace jt {

constexpr int MAX_TAPS = 257;
constexpr int KPAD = 272;        // taps padded with zeros to a multiple of 16
constexpr int BL_THREADS = 256;
constexpr int R_OUT = 8;         // outputs per thread along the blurred axis

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

// One pass over one array. Pencil p: positions i = 0..n-1 at  base(p) + i * es  (floats), each position holding
// cq(p) float4 of independent channels:
//   W pass:  p = row y,    base = y * W * C,        es = C,      cq = C/4
//   H pass:  p = strip s,  base = s * WX * C,       es = W * C,  cq = min(WX, W - s*WX) * C/4
struct BlurPass {
    const float* in;
    float* out;
    int n;                 // length of the blurred axis
    long long es;          // element stride along the blurred axis (floats)
    int npencil;
    long long pencil_stride;   // floats between pencil bases
    int cq_full;           // float4 per position of a full pencil
    int cq_last;           // float4 per position of the last pencil
    int tlg;               // output groups (of 8 positions) per tile
    int tiles_per;         // tiles per pencil
    int tpg;               // staged position groups per tile (outputs + halo + slack)
    int S;                 // float4 per staged position group in shared memory (8*cq + pad)
    int rp;                // pencils per CTA (narrow pencils are batched so a CTA has ~160 work items)
    int tapset;
    int cta_begin;         // first CTA of this pass in the launch
};
constexpr int MAX_ITEMS = 12;    // 6 density + 6 appearance factors in one launch
constexpr int MAX_SETS = 2;      // density taps, colour taps
// One launch = one pass over up to MAX_ITEMS arrays (large kernel parameters: ~5 KB)
struct BlurLaunch {
    int nitems, ntaps, adjoint;
    BlurPass it[MAX_ITEMS];
    float k[MAX_SETS][KPAD];         // taps (flipped for the adjoint), zero padded
    float kraw[MAX_SETS][MAX_TAPS];  // adjoint only: taps in the original order (edge formulas)
};

// acc += k * x on packed fp32 pairs (sm_100 FFMA2: two IEEE fmas per instruction and per fma-pipe slot). The
// 65-tap stencil is bound by fp32 FMA issue, not by HBM (2 x 65 FMAs per element against 8 bytes of traffic):
// scalar FFMA caps the two plane passes at ~0.12 ms for the 17.3 M factor elements of cfg2, FFMA2 at half that.
__device__ __forceinline__ void fma4(float4& a, const float k, const float4 x) {
    const float2 kk = make_float2(k, k);
    const float2 lo = __ffma2_rn(kk, make_float2(x.x, x.y), make_float2(a.x, a.y));
    const float2 hi = __ffma2_rn(kk, make_float2(x.z, x.w), make_float2(a.z, a.w));
    a = make_float4(lo.x, lo.y, hi.x, hi.y);
}

// 8 taps kk[0..7] applied to the window (lo = positions w..w+7, hi = w+8..w+15): acc[j] += kk[t] * x[w + j + t]
__device__ __forceinline__ void tap_block(float4 (&acc)[R_OUT], const float4 (&lo)[8], const float4 (&hi)[8], const float* kk) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const float kv = kk[t];
#pragma unroll
        for (int j = 0; j < R_OUT; ++j) {
            const int m = j + t;
            fma4(acc[j], kv, m < 8 ? lo[m] : hi[m - 8]);
        }
    }
}

__global__ void __launch_bounds__(BL_THREADS) blur_pass_kernel(const __grid_constant__ BlurLaunch M) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ __align__(8) uint64_t bar;
    float4* tile = reinterpret_cast<float4*>(raw);

    int item = 0;
    while (item + 1 < M.nitems && (int)blockIdx.x >= M.it[item + 1].cta_begin) ++item;
    const BlurPass& P = M.it[item];
    const float* __restrict__ in = P.in;
    float* __restrict__ out = P.out;
    const float* kf = M.k[P.tapset];
    const float* kraw = M.kraw[P.tapset];
    const int ntaps = M.ntaps, adjoint = M.adjoint;
    const int cta = blockIdx.x - P.cta_begin;
    const int h = ntaps >> 1;
    const int pgroup = cta / P.tiles_per;                             // group of rp pencils
    const int tidx = cta - pgroup * P.tiles_per;
    const int t0 = tidx * P.tlg * R_OUT;                              // first output position of the tile
    const int S = P.S;
    const int q0 = t0 - h;                                            // absolute position of staged position 0
    const int npos = P.tpg * 8;
    const int p_first = pgroup * P.rp;
    const int np = min(P.rp, P.npencil - p_first);                    // pencils of this CTA
    const size_t pen_f4 = (size_t)P.tpg * S;                          // float4 per staged pencil

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // bytes that will arrive: forward copies every staged position (clamped source); the adjoint
        // copies the in-range ones only
        int cnt = npos;
        if (adjoint) { const int lo = max(q0, 0), hi = min(q0 + npos, P.n); cnt = max(hi - lo, 0); }
        uint32_t bytes = 0;
        for (int r = 0; r < np; ++r) bytes += (uint32_t)cnt * (uint32_t)(p_first + r == P.npencil - 1 ? P.cq_last : P.cq_full) * 16u;
        mbar_expect_tx(&bar, bytes);
    }
    __syncthreads();
    // one work item per staged position, or per group of 8 positions where one copy can bring the whole group
    for (int itx = threadIdx.x; itx < np * npos; itx += blockDim.x) {
        const int r = itx / npos, it = itx - r * npos;
        const int pencil = p_first + r;
        const int cq = pencil == P.npencil - 1 ? P.cq_last : P.cq_full;
        const long long base = (long long)pencil * P.pencil_stride;
        const bool contiguous = P.es == (long long)cq * 4;
        const int g = it >> 3, i = it & 7;
        const int a0 = q0 + g * 8;
        float4* dst = tile + r * pen_f4 + (size_t)g * S;
        if (contiguous && a0 >= 0 && a0 + 8 <= P.n) {
            if (i == 0) bulk_g2s(dst, in + base + (long long)a0 * P.es, (uint32_t)cq * 128u, &bar);
            continue;
        }
        const int a = a0 + i;
        if ((a >= 0 && a < P.n) || !adjoint) {
            const int ac = min(max(a, 0), P.n - 1);
            bulk_g2s(dst + (size_t)i * cq, in + base + (long long)ac * P.es, (uint32_t)cq * 16u, &bar);
        } else {
            for (int c = 0; c < cq; ++c) dst[(size_t)i * cq + c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    mbar_wait(&bar, 0);
    __syncthreads();

    const int n_out = min(P.tlg * R_OUT, P.n - t0);                   // outputs of this tile (per pencil)
    const int ngroups = (n_out + R_OUT - 1) / R_OUT;
    const int nb2 = (ntaps / 16) * 2;                                 // tap blocks processed in pairs
    const int per_pencil = ngroups * P.cq_full;
    for (int wx = threadIdx.x; wx < np * per_pencil; wx += blockDim.x) {
        const int r = wx / per_pencil, w = wx - r * per_pencil;
        const int pencil = p_first + r;
        const int cq = pencil == P.npencil - 1 ? P.cq_last : P.cq_full;
        const int g = w / P.cq_full, c = w - g * P.cq_full;
        if (c >= cq) continue;                                        // narrower last strip
        const long long base = (long long)pencil * P.pencil_stride;
        const float4* ptile = tile + r * pen_f4;
        // staged position of (output j, tap t) = 8 g + j + t  ->  group g + (j+t)/8, slot (j+t)%8
        const float4* xp = ptile + (size_t)g * S + c;
        float4 acc[R_OUT];
#pragma unroll
        for (int j = 0; j < R_OUT; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 A[8], B[8];
        auto load8 = [&](float4 (&dst)[8], int grp) {
            const float4* s = xp + (size_t)grp * S;
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = s[i * cq];
        };
        load8(A, 0);
        load8(B, 1);
        for (int b = 0; b < nb2; b += 2) {
            tap_block(acc, A, B, kf + 8 * b);
            load8(A, b + 2);
            tap_block(acc, B, A, kf + 8 * b + 8);
            load8(B, b + 3);
        }
        for (int t = nb2 * 8; t < ntaps; ++t) {                       // leftover taps (65 taps: one)
            const float kv = kf[t];
#pragma unroll
            for (int j = 0; j < R_OUT; ++j) {
                const int m = j + t;
                fma4(acc[j], kv, xp[(size_t)(m >> 3) * S + (m & 7) * cq]);
            }
        }
#pragma unroll
        for (int j = 0; j < R_OUT; ++j) {
            const int m = t0 + g * R_OUT + j;
            if (m >= P.n || g * R_OUT + j >= n_out) break;
            float4 v = acc[j];
            if (adjoint && (m == 0 || m == P.n - 1)) {                // fold the replicate padding back onto the edge
                v = make_float4(0.f, 0.f, 0.f, 0.f);
                // dy[i] lives at staged position i - q0
                if (m == 0) {
                    float ks = 0.f;                                   // sum_{t <= h - i} k[t], built downwards from i = min(h, n-1)
                    const int i1 = min(h, P.n - 1);
                    for (int i = i1; i >= 0; --i) {
                        if (i == i1) { for (int t = 0; t <= h - i; ++t) ks += kraw[t]; }
                        else ks += kraw[h - i];
                        const int sp = i - q0;
                        fma4(v, ks, ptile[(size_t)(sp >> 3) * S + (sp & 7) * cq + c]);
                    }
                } else {
                    float ks = 0.f;                                   // sum_{t >= n-1-i+h} k[t], built upwards from i = n-1-h
                    const int i0 = max(P.n - 1 - h, 0);
                    for (int i = i0; i < P.n; ++i) {
                        if (i == i0) { for (int t = P.n - 1 - i + h; t < ntaps; ++t) ks += kraw[t]; }
                        else ks += kraw[P.n - 1 - i + h];
                        const int sp = i - q0;
                        fma4(v, ks, ptile[(size_t)(sp >> 3) * S + (sp & 7) * cq + c]);
                    }
                }
            }
            *reinterpret_cast<float4*>(out + base + (long long)m * P.es + 4 * c) = v;
        }
    }
}

static void plan_pass(BlurPass& P, int ntaps) {
    // tile length: balance the tiles of a pencil, at most 128 outputs (16 groups) each
    const int groups = (P.n + R_OUT - 1) / R_OUT;
    P.tiles_per = (groups + 15) / 16;
    P.tlg = (groups + P.tiles_per - 1) / P.tiles_per;
    // staged groups: outputs + halo (ntaps - 1 positions) + the look-ahead of the sliding window (2 groups)
    P.tpg = P.tlg + (ntaps - 1 + 7) / 8 + 2;
    // pad every position group so that consecutive work items (c fastest, then output group) of a warp
    // hit distinct 16-byte bank groups: S = 8 cq + (cq mod 8)  =>  g S + c == g cq + c (mod 8)
    P.S = 8 * P.cq_full + (P.cq_full & 7);
    int rp = 160 / (P.tlg * P.cq_full);
    rp = rp < 1 ? 1 : (rp > 8 ? 8 : rp);
    while (rp > 1 && (size_t)rp * P.tpg * P.S * 16 > 72 * 1024) --rp;
    if (rp > P.npencil) rp = P.npencil;
    P.rp = rp;
}
static void row_pass(BlurPass& P, const float* in, float* out, int H, int W, int C) {     // along W
    P.in = in; P.out = out;
    P.n = W; P.es = C; P.npencil = H; P.pencil_stride = (long long)W * C; P.cq_full = C / 4; P.cq_last = C / 4;
}
static void col_pass(BlurPass& P, const float* in, float* out, int H, int W, int C) {     // along H
    const int Q = C / 4;
    int WX = 24 / Q; if (WX < 1) WX = 1; if (WX > W) WX = W;       // strips of WX columns, <= 24 float4 per position
    P.in = in; P.out = out;
    P.n = H; P.es = (long long)W * C; P.npencil = (W + WX - 1) / WX; P.pencil_stride = (long long)WX * C;
    P.cq_full = WX * Q; P.cq_last = (W - (P.npencil - 1) * WX) * Q;
}

static int launch(BlurLaunch& L, cudaStream_t stream) {
    if (L.nitems == 0) return JT_OK;
    size_t smem = 0;
    int ctas = 0, items_max = 0;
    for (int i = 0; i < L.nitems; ++i) {
        BlurPass& P = L.it[i];
        plan_pass(P, L.ntaps);
        P.cta_begin = ctas;
        ctas += ((P.npencil + P.rp - 1) / P.rp) * P.tiles_per;
        const size_t b = (size_t)P.rp * P.tpg * P.S * 16;
        smem = b > smem ? b : smem;
        const int items = P.rp * P.tlg * P.cq_full;
        items_max = items > items_max ? items : items_max;
    }
    if (smem > 220 * 1024) return JT_ERR_UNSUPPORTED;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(blur_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess)
            return JT_ERR_LAUNCH;
        configured = true;
    }
    // block size: the work items (output groups x float4 columns) of the widest pass split evenly over the fewest rounds
    const int rounds = (items_max + BL_THREADS - 1) / BL_THREADS;
    int threads = (((items_max + rounds - 1) / rounds + 31) / 32) * 32;
    if (threads < 64) threads = 64;
    g_launches += 1;
    blur_pass_kernel<<<ctas, threads, smem, stream>>>(L);
    return cudaGetLastError() == cudaSuccess ? JT_OK : JT_ERR_LAUNCH;
}

}  // namespace jt

using namespace jt;

// Blur up to 12 channel-last arrays ([H][W][C] each) with two launches (one per pass): all W passes (and the
// single pass of arrays with one blurred axis) first, then the H passes of the two-axis arrays.
// h_dims = {H, W, C, axes} per array; axes bit0: along W, bit1: along H (a line factor [L][C] is H=L, W=1,
// axes=2). Forward order is W then H (bateRF.py:29-36); the adjoint runs H then W. h_tmp[i] is required when
// axes == 3. Array i uses tap set h_tapset[i] of h_taps [nsets][ntaps] (host).
extern "C" int jt_blur_multi(int n_arrays, const void* const* h_in, void* const* h_out, void* const* h_tmp,
                             const int* h_dims, const int* h_tapset, const float* h_taps, int nsets, int ntaps,
                             int adjoint, cudaStream_t stream) {
    JT_CHECK_ARG(n_arrays >= 0 && n_arrays <= MAX_ITEMS && nsets >= 1 && nsets <= MAX_SETS);
    if (n_arrays == 0) return JT_OK;
    JT_CHECK_ARG(h_in && h_out && h_dims && h_taps);
    JT_CHECK_ARG(ntaps >= 1 && (ntaps & 1) == 1 && ntaps <= MAX_TAPS);
    static thread_local BlurLaunch A, B;             // first / second pass
    for (BlurLaunch* L : {&A, &B}) {
        L->nitems = 0; L->ntaps = ntaps; L->adjoint = adjoint;
        for (int s = 0; s < MAX_SETS; ++s) {
            const float* t = h_taps + (size_t)(s < nsets ? s : 0) * ntaps;
            for (int i = 0; i < KPAD; ++i) L->k[s][i] = i < ntaps ? (adjoint ? t[ntaps - 1 - i] : t[i]) : 0.f;
            for (int i = 0; i < MAX_TAPS; ++i) L->kraw[s][i] = i < ntaps ? t[i] : 0.f;
        }
    }
    for (int i = 0; i < n_arrays; ++i) {
        const float* in = static_cast<const float*>(h_in[i]);
        float* out = static_cast<float*>(h_out[i]);
        float* tmp = h_tmp ? static_cast<float*>(h_tmp[i]) : nullptr;
        const int H = h_dims[4 * i], W = h_dims[4 * i + 1], C = h_dims[4 * i + 2], axes = h_dims[4 * i + 3];
        const int set = h_tapset ? h_tapset[i] : 0;
        JT_CHECK_ARG(in && out && H >= 1 && W >= 1 && C >= 4 && (C & 3) == 0 && axes >= 1 && axes <= 3);
        JT_CHECK_ARG(set >= 0 && set < nsets);
        JT_CHECK_ARG(axes != 3 || tmp);
        JT_CHECK_ARG(in != out);
        JT_CHECK_ARG(!(axes & 1) || W >= 2);
        JT_CHECK_ARG(!(axes & 2) || H >= 2);
        JT_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(tmp) & 15) == 0);
        BlurPass& a = A.it[A.nitems];
        a.tapset = set;
        if (axes == 1) { row_pass(a, in, out, H, W, C); A.nitems++; }
        else if (axes == 2) { col_pass(a, in, out, H, W, C); A.nitems++; }
        else {
            BlurPass& b = B.it[B.nitems];
            b.tapset = set;
            if (!adjoint) { row_pass(a, in, tmp, H, W, C); col_pass(b, tmp, out, H, W, C); }
            else { col_pass(a, in, tmp, H, W, C); row_pass(b, tmp, out, H, W, C); }
            A.nitems++; B.nitems++;
        }
    }
    if (int rc = launch(A, stream)) return rc;
    return launch(B, stream);
}

// One array (see jt_blur_multi).
extern "C" int jt_blur_cl(const float* in, float* out, float* tmp, int H, int W, int C, const float* h_taps,
                          int ntaps, int axes, int adjoint, cudaStream_t stream) {
    const void* ins[1] = {in};
    void* outs[1] = {out};
    void* tmps[1] = {tmp};
    const int dims[4] = {H, W, C, axes};
    const int set[1] = {0};
    JT_CHECK_ARG(h_taps);
    return jt_blur_multi(1, ins, outs, tmps, dims, set, h_taps, 1, ntaps, adjoint, stream);
}
