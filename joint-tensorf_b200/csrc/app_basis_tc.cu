// K2(appearance) + K3(basis) fused -- plane x line gather of the appearance components and the
// basis_mat projection on the tensor cores, per tile of 128 appearance samples.
//
// Replaces reference BAT_VMSplit.compute_appfeature (bateRF.py:97-130; twin
// tensoRF.py:254-270): 6 F.grid_sample calls, the concatenation of the three
// plane*line products [A,144] and `self.basis_mat(...)` (Linear 144 -> 27, no bias).
//
// The [A,144] component matrix never goes to HBM in fp32: the gather lanes convert their
// products to bf16 (hi + lo terms when SPLIT == 2) and store them straight into the UMMA
// canonical shared-memory tile; one tcgen05.mma chain (accumulator in TMEM) projects the
// tile onto basis_mat, and the epilogue writes the 27 features (+ the sample's view
// direction, so the shading-head kernel needs no index chain) as one 128-byte row.
// Training keeps the bf16 (hi) component tile for the basis_mat weight gradient: it is
// pushed to the HBM staging area with one bulk async store per tile.
//
// Thread mapping of the gather (16 warps): 4 lanes share a sample, lane `sub` owns channel
// quads sub, sub+4, sub+8 of each plane (64 contiguous bytes per tap across the 4 lanes),
// a warp covers 8 consecutive samples of (mostly) one ray -- same as vm_fwd_kernel.
#include "head_tc.cuh"
#include "../../include/jt_vm.h"

namespace jt {
using namespace tc;

constexpr int GT = 512;                       // threads per CTA: 16 warps x 8 samples = one 128-sample tile

template <int SPLIT>
struct GSmem {
    static constexpr int WB = tile_bytes(NB, CT), A0 = tile_bytes(TM, CT);
    static constexpr int off_wb = 0, off_a_hi = off_wb + SPLIT * WB, off_a_lo = off_a_hi + A0;
    static constexpr int total = off_a_lo + (SPLIT == 2 ? A0 : 0);
};

__device__ __forceinline__ uint2 pack4_bf16(float4 v) { return make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w)); }

// SH = true: the epilogue also shades the sample with SHRender (tensorBase.py:68-72,
// eval_sh_bases(2, .) sh.py:88-113): rgb = relu(sum_k Y_k(dir) feat[c*9+k] + 0.5); only the
// view direction (for the backward) and rgb leave the kernel, the 27 features never do.
template <int SPLIT, bool SAVE, bool SH, bool B16>
__global__ void __launch_bounds__(GT, 2) app_basis_fwd_kernel(Factors F, const float4* __restrict__ samp,
                                                              const int* __restrict__ aidx, const int* __restrict__ sidx,
                                                              const float* __restrict__ rays_d, int S, int normalize_dir,
                                                              const float* __restrict__ Wb, const int* __restrict__ n_dev,
                                                              int n_fixed, float* __restrict__ featdir,
                                                              float* __restrict__ rgb_sh,
                                                              unsigned char* __restrict__ stage) {
    using L = GSmem<SPLIT>;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float sdir[SH ? TM : 1][3];       // view directions of the tile's rows (SH epilogue)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sub = lane & 3, grp = lane >> 2;
    const int n = n_dev ? *n_dev : n_fixed;
    unsigned char* wb_hi = smem + L::off_wb;
    unsigned char* wb_lo = SPLIT == 2 ? wb_hi + L::WB : nullptr;
    unsigned char* a_hi = smem + L::off_a_hi;
    unsigned char* a_lo = SPLIT == 2 ? smem + L::off_a_lo : nullptr;

    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 32);
    stage_weight(wb_hi, wb_lo, Wb, CT, F_, CT, NB, CT, nullptr, -1, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    uint32_t phase = 0;
    const int r = warp * 8 + grp;                 // tile row gathered by this lane quad
    // Tuning notes (B200, cfg2: 0.62 ms): the kernel is latency-bound at 2 CTAs x 16 warps with exactly 64 registers --
    // predicating half of the tap loads off changes its time by 2 % (not an LSU/L2 throughput problem), one CTA per
    // SM with 128 registers is slower (0.82 ms), and fetching the next tile's slot / sample record / direction one
    // tile ahead costs registers the tap loads need (0.82 - 0.91 ms).

    for (int tile = blockIdx.x; (long long)tile * TM < n; tile += gridDim.x) {
        const int row = tile * TM + r;
        const bool live = row < n;
        // ---- gather: plane x line products of this sample -> bf16 tile (rows past n are zero)
        float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f);
        int j = 0;
        if (live) { j = aidx[row]; u4 = samp[j]; }
        const float u[3] = {u4.x, u4.y, u4.z};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const Tap tx = make_tap(u[mat0(i)], F.W[i]), ty = make_tap(u[mat1(i)], F.H[i]), tl = make_tap(u[vecm(i)], F.L[i]);
            const size_t C = CT / 3;
            const size_t r0 = (size_t)ty.i0 * F.W[i], r1 = (size_t)ty.i1 * F.W[i];
            // fp32 taps: six live pointers + immediate offsets (the form the 64-register budget of this kernel was
            // tuned for); bf16 taps: the same with 2-byte elements
            const float* p00 = F.plane[i] + (r0 + tx.i0) * C + sub * 4;
            const float* p10 = F.plane[i] + (r0 + tx.i1) * C + sub * 4;
            const float* p01 = F.plane[i] + (r1 + tx.i0) * C + sub * 4;
            const float* p11 = F.plane[i] + (r1 + tx.i1) * C + sub * 4;
            const float* l0 = F.line[i] + (size_t)tl.i0 * C + sub * 4;
            const float* l1 = F.line[i] + (size_t)tl.i1 * C + sub * 4;
            const unsigned short* P16 = reinterpret_cast<const unsigned short*>(F.plane[i]);
            const unsigned short* L16 = reinterpret_cast<const unsigned short*>(F.line[i]);
            const unsigned short* h00 = P16 + (r0 + tx.i0) * C + sub * 4;
            const unsigned short* h10 = P16 + (r0 + tx.i1) * C + sub * 4;
            const unsigned short* h01 = P16 + (r1 + tx.i0) * C + sub * 4;
            const unsigned short* h11 = P16 + (r1 + tx.i1) * C + sub * 4;
            const unsigned short* hl0 = L16 + (size_t)tl.i0 * C + sub * 4;
            const unsigned short* hl1 = L16 + (size_t)tl.i1 * C + sub * 4;
            const float w00 = tx.w0 * ty.w0, w10 = tx.w1 * ty.w0, w01 = tx.w0 * ty.w1, w11 = tx.w1 * ty.w1;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float4 a, b, c, d, la, lb;
                if (live) {
                    if (B16) {
                        auto ld = [](const unsigned short* p) { return bf16x4_to_f4(__ldg(reinterpret_cast<const uint2*>(p))); };
                        a = ld(h00 + 16 * k); b = ld(h10 + 16 * k); c = ld(h01 + 16 * k); d = ld(h11 + 16 * k);
                        la = ld(hl0 + 16 * k); lb = ld(hl1 + 16 * k);
                    } else {
                        a = ldg4(p00 + 16 * k); b = ldg4(p10 + 16 * k); c = ldg4(p01 + 16 * k); d = ldg4(p11 + 16 * k);
                        la = ldg4(l0 + 16 * k); lb = ldg4(l1 + 16 * k);
                    }
                } else {
                    a = b = c = d = la = lb = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                float4 v;
                v.x = (a.x * w00 + b.x * w10 + c.x * w01 + d.x * w11) * (la.x * tl.w0 + lb.x * tl.w1);
                v.y = (a.y * w00 + b.y * w10 + c.y * w01 + d.y * w11) * (la.y * tl.w0 + lb.y * tl.w1);
                v.z = (a.z * w00 + b.z * w10 + c.z * w01 + d.z * w11) * (la.z * tl.w0 + lb.z * tl.w1);
                v.w = (a.w * w00 + b.w * w10 + c.w * w01 + d.w * w11) * (la.w * tl.w0 + lb.w * tl.w1);
                const int c0 = i * (CT / 3) + 16 * k + 4 * sub;                 // first channel of this quad
                const int off = (c0 >> 3) * (TM * 16) + r * 16 + ((c0 >> 2) & 1) * 8;
                if (SPLIT == 2) {
                    uint2 h, l;
                    split_pair(v.x, v.y, h.x, l.x); split_pair(v.z, v.w, h.y, l.y);
                    *reinterpret_cast<uint2*>(a_hi + off) = h;
                    *reinterpret_cast<uint2*>(a_lo + off) = l;
                } else {
                    *reinterpret_cast<uint2*>(a_hi + off) = pack4_bf16(v);
                }
            }
        }
        // view direction of the sample (viewdirs = ray_dir, normalised for NDC rays: batBase.py:63-66)
        if (live && sub == 0) {
            const int ray = sidx[j] / S;
            float d0 = rays_d[3 * ray], d1 = rays_d[3 * ray + 1], d2 = rays_d[3 * ray + 2];
            if (normalize_dir) {
                const float nn = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
                d0 /= nn; d1 /= nn; d2 /= nn;
            }
            __stcs(reinterpret_cast<float4*>(featdir + (size_t)row * FD + 28), make_float4(d0, d1, d2, 0.f));
            if (SH) { sdir[r][0] = d0; sdir[r][1] = d1; sdir[r][2] = d2; }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- basis projection: feat[128 x 32] = A0[128 x 144] * Wb^T
        if (tid == 0) {
            tc_fence_after();
            issue_gemm_kmajor<SPLIT>(tmem, a_hi, a_lo, wb_hi, wb_lo, CT, NB, NB);
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        if (SAVE && tid == 0) {                      // the MMAs have consumed the tile; push its hi term to HBM
            bulk_s2g(stage + (size_t)tile * STAGE_TILE_BYTES + OFF_A0, a_hi, SZ_A0);
            bulk_commit();
        }
        // ---- epilogue: warps 0-3 own TMEM lanes 0..127 = tile rows; 7 x 16 B per row
        if (warp < 4) {
            float f[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), f);
            const int erow = tile * TM + tid;
            if (erow < n) {
                if (SH) {
                    const float d[3] = {sdir[tid][0], sdir[tid][1], sdir[tid][2]};
                    float y[9], c3[3];
                    sh9(d, y);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float s = 0.f;
#pragma unroll
                        for (int k = 0; k < 9; ++k) s += y[k] * f[c * 9 + k];
                        c3[c] = fmaxf(s + 0.5f, 0.f);
                    }
                    __stcs(reinterpret_cast<float4*>(rgb_sh) + erow, make_float4(c3[0], c3[1], c3[2], 0.f));
                } else {
                    float4* dst = reinterpret_cast<float4*>(featdir + (size_t)erow * FD);
#pragma unroll
                    for (int q = 0; q < 7; ++q)
                        __stcs(dst + q, make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], q == 6 ? 0.f : f[4 * q + 3]));
                }
            }
        }
        if (SAVE && tid == 0) bulk_wait_read0();     // the tile may be overwritten by the next gather
        tc_fence_before();
        __syncthreads();
    }
    if (SAVE && tid == 0) bulk_wait0();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 32);
}

}  // namespace jt

using namespace jt;

static int app_basis_launch(int split, int sh, const void* const* h_factors, const int* h_dims, const float* samp,
                            const int* aidx, const int* sidx, const float* rays_d, int n_samples, int normalize_dir,
                            const float* Wb, const int* n_dev, int n_max, float* featdir, float* rgb, void* stage,
                            cudaStream_t stream) {
    JT_CHECK_ARG(h_factors && h_dims && samp && aidx && sidx && rays_d && Wb && featdir && n_samples > 0);
    JT_CHECK_ARG(split == 1 || split == 2);
    JT_CHECK_ARG(!sh || rgb);
    JT_CHECK_ARG((reinterpret_cast<uintptr_t>(stage) & 127) == 0);
    if (n_max <= 0) return JT_OK;
    Factors F;
    if (int rc = fill_factors(F, h_factors, h_dims)) return rc;
    if (F.C[0] != CT / 3 || F.C[1] != CT / 3 || F.C[2] != CT / 3) return JT_ERR_UNSUPPORTED;
    long long tiles = ((long long)n_max + TM - 1) / TM;
    int grid = (int)(tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs);
    unsigned char* st = static_cast<unsigned char*>(stage);
    g_launches += 1;
#define JT_LAUNCH_GB(SP, SV, SHV, BV)                                                                                \
    {                                                                                                                \
        const int smem = GSmem<SP>::total;                                                                           \
        if (int rc = set_smem(app_basis_fwd_kernel<SP, SV, SHV, BV>, smem)) return rc;                               \
        app_basis_fwd_kernel<SP, SV, SHV, BV><<<grid, GT, smem, stream>>>(F, reinterpret_cast<const float4*>(samp),  \
                                                                          aidx, sidx, rays_d, n_samples,             \
                                                                          normalize_dir, Wb, n_dev, n_max, featdir,  \
                                                                          rgb, st);                                  \
    }
#define JT_LAUNCH_G(SP, SV, SHV) { if (F.bf16) JT_LAUNCH_GB(SP, SV, SHV, true) else JT_LAUNCH_GB(SP, SV, SHV, false) }
#define JT_LAUNCH_GS(SP, SV) { if (sh) JT_LAUNCH_G(SP, SV, true) else JT_LAUNCH_G(SP, SV, false) }
    if (split == 1 && !st) JT_LAUNCH_GS(1, false)
    else if (split == 1) JT_LAUNCH_GS(1, true)
    else if (!st) JT_LAUNCH_GS(2, false)
    else JT_LAUNCH_GS(2, true)
#undef JT_LAUNCH_GS
#undef JT_LAUNCH_G
#undef JT_LAUNCH_GB
    JT_RETURN_LAUNCH();
}

extern "C" int jt_app_basis_fwd_tc(int split, const void* const* h_factors, const int* h_dims, const float* samp,
                                   const int* aidx, const int* sidx, const float* rays_d, int n_samples,
                                   int normalize_dir, const float* Wb, const int* n_dev, int n_max, float* featdir,
                                   void* stage, cudaStream_t stream) {
    return app_basis_launch(split, 0, h_factors, h_dims, samp, aidx, sidx, rays_d, n_samples, normalize_dir, Wb, n_dev,
                            n_max, featdir, nullptr, stage, stream);
}

extern "C" int jt_app_basis_sh_fwd_tc(int split, const void* const* h_factors, const int* h_dims, const float* samp,
                                      const int* aidx, const int* sidx, const float* rays_d, int n_samples,
                                      int normalize_dir, const float* Wb, const int* n_dev, int n_max, float* featdir,
                                      float* rgb, void* stage, cudaStream_t stream) {
    JT_CHECK_ARG(rgb);
    return app_basis_launch(split, 1, h_factors, h_dims, samp, aidx, sidx, rays_d, n_samples, normalize_dir, Wb, n_dev,
                            n_max, featdir, rgb, stage, stream);
}
