// K2 backward, ray-walking form -- factor-gradient scatter with per-cell run merging.
//
// Replaces the 12 grid_sampler_2d_backward calls of one training step (autograd of
// bateRF.py:81-82,124-127 / tensoRF.py:245-248,263-266) plus the autograd of
// `rays_pts = o + d * t` / normalize_coord (tensorBase.py:502-503,597) that turns the
// coordinate gradients into d/d rays_o, d/d rays_d for the pose optimisation.
//
// Why a second scatter kernel: vm_bwd_kernel (vm_gather.cu) issues one 16-byte RED per
// (sample, tap, channel quad) and is bound by the LSU's RED rate (~1.3 cycles per lane-op
// per SM; 216 lane-ops per appearance sample). The marcher steps half a voxel
// (step_ratio 0.5, tensorBase.py:483-484), so consecutive samples of a ray stay in the
// same bilinear cell of a plane for ~2.1 samples and in the same line cell for ~3.9
// (measured on the cfg2 batch). Here a *walker* (C/4 lanes, one channel quad each) walks a
// segment of consecutive samples of the compacted, ray-major list, keeps the four corner
// gradients and the two line gradients of the current cell in registers, and flushes
// them with REDs only when the cell changes: 2.6 instead of 6 RED units per sample and
// plane, and the tap values are re-loaded only on a cell change as well.
// The coordinate gradient is reduced per ray in registers (sum_j du_j and sum_j du_j t_j
// are all the pose gradient needs) and flushed once per ray segment, so no per-sample
// dL/du array is written and no separate ray_bwd pass is needed.
#include <limits.h>
#include "jt_common.cuh"
#include "../../include/jt_vm.h"

namespace jt {

__device__ __forceinline__ float4 f4z() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_fma(float4& acc, float4 a, float s) {
    acc.x = fmaf(a.x, s, acc.x); acc.y = fmaf(a.y, s, acc.y); acc.z = fmaf(a.z, s, acc.z); acc.w = fmaf(a.w, s, acc.w);
}
__device__ __forceinline__ float4 f4_mul2(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float f4_dot2(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ bool f4_any(float4 a) { return a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f; }

// unclamped tap position on an axis of n texels (ATen grid_sampler, align_corners=True)
struct Pos { int i0; float f; float scale; };
__device__ __forceinline__ Pos axis_pos(float g, int n) {
    Pos p;
    p.scale = 0.5f * (float)(n - 1);
    const float x = (g + 1.0f) * p.scale;
    const float xf = floorf(x);
    p.f = x - xf;
    p.i0 = (int)fminf(fmaxf(xf, -2.0f), (float)n);       // NaN -> -2: both taps out of range
    return p;
}

struct ScatterArgs {
    Factors F;
    FactorGrads G;
    const float4* samp;     // [V] (u_x, u_y, u_z, t)
    const int* slot;        // element e -> sample slot (appearance list) or nullptr
    const int* sidx;        // [V] ray * S + k
    const int* n_dev;
    int n_fixed;
    const float* gin;       // APP: [n][ctot]; density: [n]
    float* d_o;             // [N][3], accumulated
    float* d_d;             // [N][3], accumulated
    float inv[3];           // 2 / aabbSize (normalize_coord)
    int S;                  // samples per ray (sidx decoding)
    int seg;                // samples per walker segment
};

// One walker = LW lanes = the LW channel quads of one plane (C = 4 LW). Work unit =
// (segment of `seg` consecutive elements, plane i). Walkers are laid out over the CTA's
// threads contiguously (they may straddle warps: no warp-level collectives are used).
template <bool APP>
__global__ void __launch_bounds__(256, 2) vm_scatter_walk_kernel(ScatterArgs A, int LW, int walkers_per_cta) {
    const Factors& F = A.F;
    const int n = A.n_dev ? *A.n_dev : A.n_fixed;
    const int wl = threadIdx.x / LW;                     // walker within the CTA
    const int q = (threadIdx.x - wl * LW) * 4;           // first channel of this lane's quad
    if (wl >= walkers_per_cta) return;
    const long long n_units = 3LL * ((n + A.seg - 1) / A.seg);
    const long long stride = (long long)gridDim.x * walkers_per_cta;
    for (long long unit = (long long)blockIdx.x * walkers_per_cta + wl; unit < n_units; unit += stride) {
        const int i = (int)(unit % 3);
        const int e0 = (int)(unit / 3) * A.seg;
        const int e1 = min(e0 + A.seg, n);
        const int W = F.W[i], H = F.H[i], L = F.L[i], C = F.C[i];
        if (q >= C) continue;
        const float* __restrict__ P = F.plane[i] + q;
        const float* __restrict__ Ln = F.line[i] + q;
        float* __restrict__ GP = A.G.plane[i] + q;
        float* __restrict__ GL = A.G.line[i] + q;
        const int ax = mat0(i), ay = mat1(i), al = vecm(i);

        // current plane cell / line cell state
        int cx = INT_MIN, cy = INT_MIN, cl = INT_MIN;
        unsigned o00 = 0, o10 = 0, o01 = 0, o11 = 0, ol0 = 0, ol1 = 0;   // element offsets (< 2^31, checked on the host)
        float mx0 = 0.f, mx1 = 0.f, my0 = 0.f, my1 = 0.f, ml0 = 0.f, ml1 = 0.f;     // in-range masks
        float4 a = f4z(), b = f4z(), c = f4z(), d = f4z(), la = f4z(), lb = f4z();
        float4 g00 = f4z(), g10 = f4z(), g01 = f4z(), g11 = f4z(), gl0 = f4z(), gl1 = f4z();
        // per-ray coordinate-gradient sums
        int ray = -1;
        int ray_end = -1;
        float so = 0.f, sdx = 0.f;   // axis ax:  sum du, sum du*t
        float sp = 0.f, sdy = 0.f;   // axis ay
        float sq = 0.f, sdl = 0.f;   // axis al

        auto flush_plane = [&]() {
            if (mx0 * my0 != 0.f && f4_any(g00)) red_add_v4(GP + o00, g00);
            if (mx1 * my0 != 0.f && f4_any(g10)) red_add_v4(GP + o10, g10);
            if (mx0 * my1 != 0.f && f4_any(g01)) red_add_v4(GP + o01, g01);
            if (mx1 * my1 != 0.f && f4_any(g11)) red_add_v4(GP + o11, g11);
            g00 = f4z(); g10 = f4z(); g01 = f4z(); g11 = f4z();
        };
        auto flush_line = [&]() {
            if (ml0 != 0.f && f4_any(gl0)) red_add_v4(GL + ol0, gl0);
            if (ml1 != 0.f && f4_any(gl1)) red_add_v4(GL + ol1, gl1);
            gl0 = f4z(); gl1 = f4z();
        };
        auto flush_ray = [&]() {
            if (ray >= 0) {
                const float ix = A.inv[ax], iy = A.inv[ay], il = A.inv[al];
                if (so != 0.f) atomicAdd(A.d_o + 3 * ray + ax, so * ix);
                if (sp != 0.f) atomicAdd(A.d_o + 3 * ray + ay, sp * iy);
                if (sq != 0.f) atomicAdd(A.d_o + 3 * ray + al, sq * il);
                if (sdx != 0.f) atomicAdd(A.d_d + 3 * ray + ax, sdx * ix);
                if (sdy != 0.f) atomicAdd(A.d_d + 3 * ray + ay, sdy * iy);
                if (sdl != 0.f) atomicAdd(A.d_d + 3 * ray + al, sdl * il);
            }
            so = sdx = sp = sdy = sq = sdl = 0.f;
        };

        // software pipeline: the sample record of element e+1 is fetched while e is processed
        int jn = 0, sn = 0;
        float4 un = f4z();
        if (e0 < e1) {
            jn = A.slot ? A.slot[e0] : e0;
            un = A.samp[jn];
            sn = A.sidx[jn];
        }
        for (int e = e0; e < e1; ++e) {
            const float4 u4 = un;
            const int sid = sn;
            if (e + 1 < e1) {
                jn = A.slot ? A.slot[e + 1] : e + 1;
                un = A.samp[jn];
                sn = A.sidx[jn];
            }
            const float u[3] = {u4.x, u4.y, u4.z};
            if (sid >= ray_end || ray < 0) {          // new ray (lists are ray-major)
                flush_ray();
                ray = sid / A.S;
                ray_end = (ray + 1) * A.S;
            }
            const Pos px = axis_pos(u[ax], W), py = axis_pos(u[ay], H), pl = axis_pos(u[al], L);
            if (px.i0 != cx || py.i0 != cy) {
                flush_plane();
                cx = px.i0; cy = py.i0;
                const int x0 = cx, x1 = cx + 1, y0 = cy, y1 = cy + 1;
                mx0 = (x0 >= 0 && x0 < W) ? 1.f : 0.f; mx1 = (x1 >= 0 && x1 < W) ? 1.f : 0.f;
                my0 = (y0 >= 0 && y0 < H) ? 1.f : 0.f; my1 = (y1 >= 0 && y1 < H) ? 1.f : 0.f;
                const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x1, 0), W - 1);
                const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y1, 0), H - 1);
                o00 = (unsigned)((yc0 * W + xc0) * C); o10 = (unsigned)((yc0 * W + xc1) * C);
                o01 = (unsigned)((yc1 * W + xc0) * C); o11 = (unsigned)((yc1 * W + xc1) * C);
                a = ldg4(P + o00); b = ldg4(P + o10); c = ldg4(P + o01); d = ldg4(P + o11);
            }
            if (pl.i0 != cl) {
                flush_line();
                cl = pl.i0;
                const int l0 = cl, l1 = cl + 1;
                ml0 = (l0 >= 0 && l0 < L) ? 1.f : 0.f; ml1 = (l1 >= 0 && l1 < L) ? 1.f : 0.f;
                ol0 = (unsigned)(min(max(l0, 0), L - 1) * C); ol1 = (unsigned)(min(max(l1, 0), L - 1) * C);
                la = ldg4(Ln + ol0); lb = ldg4(Ln + ol1);
            }
            float4 g4;
            if (APP) g4 = __ldcs(reinterpret_cast<const float4*>(A.gin + (size_t)e * F.ctot + F.off[i] + q));
            else { const float gs = A.gin[e]; g4 = make_float4(gs, gs, gs, gs); }

            const float wx0 = (1.f - px.f) * mx0, wx1 = px.f * mx1;
            const float wy0 = (1.f - py.f) * my0, wy1 = py.f * my1;
            const float wl0 = (1.f - pl.f) * ml0, wl1 = pl.f * ml1;
            const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
            float4 pv, lv;
            pv.x = a.x * w00 + b.x * w10 + c.x * w01 + d.x * w11; pv.y = a.y * w00 + b.y * w10 + c.y * w01 + d.y * w11;
            pv.z = a.z * w00 + b.z * w10 + c.z * w01 + d.z * w11; pv.w = a.w * w00 + b.w * w10 + c.w * w01 + d.w * w11;
            lv.x = la.x * wl0 + lb.x * wl1; lv.y = la.y * wl0 + lb.y * wl1;
            lv.z = la.z * wl0 + lb.z * wl1; lv.w = la.w * wl0 + lb.w * wl1;
            const float4 gl = f4_mul2(g4, lv);       // dL/dP (interpolated plane value)
            const float4 gp = f4_mul2(g4, pv);       // dL/dL (interpolated line value)
            f4_fma(g00, gl, w00); f4_fma(g10, gl, w10); f4_fma(g01, gl, w01); f4_fma(g11, gl, w11);
            f4_fma(gl0, gp, wl0); f4_fma(gl1, gp, wl1);
            // d/d index (ATen grid_sampler_2d_backward: out-of-range taps read as 0)
            float4 dpx, dpy, dl;
#define JT_DP(k)                                                                                         \
            dpx.k = (b.k * mx1 - a.k * mx0) * wy0 + (d.k * mx1 - c.k * mx0) * wy1;                       \
            dpy.k = (c.k * my1 - a.k * my0) * wx0 + (d.k * my1 - b.k * my0) * wx1;                       \
            dl.k = lb.k * ml1 - la.k * ml0;
            JT_DP(x) JT_DP(y) JT_DP(z) JT_DP(w)
#undef JT_DP
            const float t = u4.w;
            const float dux = f4_dot2(gl, dpx) * px.scale, duy = f4_dot2(gl, dpy) * py.scale,
                        dul = f4_dot2(gp, dl) * pl.scale;
            so += dux; sdx = fmaf(dux, t, sdx);
            sp += duy; sdy = fmaf(duy, t, sdy);
            sq += dul; sdl = fmaf(dul, t, sdl);
        }
        flush_plane();
        flush_line();
        flush_ray();
    }
}

}  // namespace jt

using namespace jt;

extern "C" int jt_vm_scatter_rays(int app, const void* const* h_factors, void* const* h_factor_grads,
                                  const int* h_dims, const float* samp, const int* slot, const int* sidx,
                                  const int* n_dev, int n_max, const float* gin, int n_samples, const float* h_inv,
                                  float* d_o, float* d_d, cudaStream_t stream) {
    JT_CHECK_ARG(h_factors && h_factor_grads && h_dims && samp && sidx && gin && h_inv && d_o && d_d && n_samples > 0);
    if (n_max <= 0) return JT_OK;
    ScatterArgs A;
    if (int rc = fill_factors(A.F, h_factors, h_dims)) return rc;
    int cmax = 0;
    for (int i = 0; i < 3; ++i) {
        A.G.plane[i] = static_cast<float*>(h_factor_grads[i]);
        A.G.line[i] = static_cast<float*>(h_factor_grads[3 + i]);
        JT_CHECK_ARG(A.G.plane[i] && A.G.line[i]);
        cmax = A.F.C[i] > cmax ? A.F.C[i] : cmax;
        JT_CHECK_ARG((long long)A.F.H[i] * A.F.W[i] * A.F.C[i] < 2147483647LL);
    }
    JT_CHECK_ARG((long long)n_max < 2147483647LL);
    A.samp = reinterpret_cast<const float4*>(samp);
    A.slot = slot; A.sidx = sidx; A.n_dev = n_dev; A.n_fixed = n_max; A.gin = gin;
    A.d_o = d_o; A.d_d = d_d;
    for (int a = 0; a < 3; ++a) A.inv[a] = h_inv[a];
    A.S = n_samples;
    A.seg = 32;
    const int LW = cmax / 4;                                  // lanes per walker
    JT_CHECK_ARG(LW >= 1 && LW <= 256);
    int wpc = 256 / LW;                                       // walkers per CTA (<= 256 threads, 2 CTAs per SM)
    int threads = ((wpc * LW + 31) / 32) * 32;
    long long units = 3LL * (((long long)n_max + A.seg - 1) / A.seg);
    long long want = (units + wpc - 1) / wpc;
    long long cap = (long long)kNumSMs * 16;
    int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    g_launches += 1;
    if (app) vm_scatter_walk_kernel<true><<<grid, threads, 0, stream>>>(A, LW, wpc);
    else vm_scatter_walk_kernel<false><<<grid, threads, 0, stream>>>(A, LW, wpc);
    JT_RETURN_LAUNCH();
}
