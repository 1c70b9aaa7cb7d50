// K3 (strict-fp32 path) -- basis_mat projection and the shading heads as plain
// SIMT fp32 GEMMs plus the positional-encoding and SH kernels around them.
//
// Replaces reference basis_mat (tensoRF.py:156,270 / bateRF.py:130),
// positional_encoding (tensorBase.py:43-55), MLPRender_Fea.forward
// (tensorBase.py:116-126), MLPRender_Fea_WeakView.forward (tensorBase.py:198-214),
// SHRender (tensorBase.py:68-72) + eval_sh_bases deg 2 (sh.py:88-113) and their
// autograd. This is the bit-faithful fp32 mode (<= 1e-4 parity class); the
// tensor-core (tcgen05) fused head lives in shade_tc.cu.
#include "jt_common.cuh"
#include "../../include/jt_vm.h"

namespace jt {

// ------------------------------------------------------------------ Y = act(X * W^T + b) [* (mask > 0)]
constexpr int BM = 128, BN = 64, KC = 32, XP = BM + 4;

__global__ void __launch_bounds__(256) gemm_nt_kernel(const float* __restrict__ X, int ldx,
                                                      const float* __restrict__ W, int ldw, int w_kn,
                                                      const float* __restrict__ bias, float* __restrict__ Y, int ldy,
                                                      const float* __restrict__ mask, int ldm,
                                                      const int* __restrict__ m_dev, int m_fixed, int N, int K,
                                                      int act) {
    __shared__ __align__(16) float Xs[KC][XP];
    __shared__ __align__(16) float Ws[KC][BN];
    const int M = m_dev ? *m_dev : m_fixed;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.y * BN;
    for (int tile = blockIdx.x; (long long)tile * BM < M; tile += gridDim.x) {
        const int m0 = tile * BM;
        float acc[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
        for (int k0 = 0; k0 < K; k0 += KC) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int idx = tid + it * 256;
                const int r = idx >> 3, kq = idx & 7;
                const int m = m0 + r, kc = k0 + kq * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < M && kc < K) {
                    const float* src = X + (size_t)m * ldx + kc;
                    if (kc + 3 < ldx) {
                        v = *reinterpret_cast<const float4*>(src);
                        if (kc + 1 >= K) v.y = 0.f;
                        if (kc + 2 >= K) v.z = 0.f;
                        if (kc + 3 >= K) v.w = 0.f;
                    } else {
                        v.x = src[0];
                        if (kc + 1 < K) v.y = src[1];
                        if (kc + 2 < K) v.z = src[2];
                    }
                }
                Xs[kq * 4 + 0][r] = v.x; Xs[kq * 4 + 1][r] = v.y;
                Xs[kq * 4 + 2][r] = v.z; Xs[kq * 4 + 3][r] = v.w;
            }
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int idx = tid + it * 256;
                const int kk = idx >> 6, nn = idx & 63;
                const int k = k0 + kk, n = n0 + nn;
                float w = 0.f;
                if (k < K && n < N) w = w_kn ? W[(size_t)k * ldw + n] : W[(size_t)n * ldw + k];
                Ws[kk][nn] = w;
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < KC; ++kk) {
                const float4 xa = *reinterpret_cast<const float4*>(&Xs[kk][ty * 8]);
                const float4 xb = *reinterpret_cast<const float4*>(&Xs[kk][ty * 8 + 4]);
                const float4 w = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
                const float xr[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
                const float wc[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(xr[r], wc[c], acc[r][c]);
            }
            __syncthreads();
        }
        const int n = n0 + tx * 4;
        const int wy = min(ldy, (N + 3) & ~3);      // writable width: N rounded up to a float4 (pad columns <- 0)
        if (n < wy) {
            float b[4] = {0.f, 0.f, 0.f, 0.f};
            if (bias) {
#pragma unroll
                for (int c = 0; c < 4; ++c) if (n + c < N) b[c] = bias[n + c];
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int m = m0 + ty * 8 + r;
                if (m >= M) continue;
                float v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float t = acc[r][c] + b[c];
                    if (act == 1) t = fmaxf(t, 0.f);
                    else if (act == 2) t = 1.0f / (1.0f + expf(-t));
                    if (mask && n + c < N) t = mask[(size_t)m * ldm + n + c] > 0.f ? t : 0.f;
                    v[c] = (n + c < N) ? t : 0.f;
                }
                float* dst = Y + (size_t)m * ldy + n;
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
}

// ------------------------------------------------------------------ dW[N][K] += dY^T X ; db[N] += sum dY
constexpr int TBM = 32;

__global__ void __launch_bounds__(256) gemm_tn_kernel(const float* __restrict__ dY, int ldy,
                                                      const float* __restrict__ X, int ldx,
                                                      const int* __restrict__ m_dev, int m_fixed, int N, int K,
                                                      float* __restrict__ dW, int ldw, float* __restrict__ db) {
    extern __shared__ __align__(16) float smem[];
    const int NP = (N + 3) & ~3;
    const int KP = (K + 1 + 7) & ~7;          // column K carries 1.0 -> bias gradient
    float* Ys = smem;                          // [TBM][NP]
    float* Xs = smem + TBM * NP;               // [TBM][KP]
    const int M = m_dev ? *m_dev : m_fixed;
    const int tid = threadIdx.x;
    const int ng = NP >> 2, kg = KP >> 3, ntile = ng * kg;
    float acc[2][4][8];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[s][a][b] = 0.f;
    for (int chunk = blockIdx.x; (long long)chunk * TBM < M; chunk += gridDim.x) {
        const int m0 = chunk * TBM;
        for (int idx = tid; idx < TBM * NP; idx += 256) {
            const int r = idx / NP, c = idx - r * NP;
            const int m = m0 + r;
            Ys[idx] = (m < M && c < N) ? dY[(size_t)m * ldy + c] : 0.f;
        }
        for (int idx = tid; idx < TBM * KP; idx += 256) {
            const int r = idx / KP, c = idx - r * KP;
            const int m = m0 + r;
            float v = 0.f;
            if (m < M) v = (c < K) ? X[(size_t)m * ldx + c] : (c == K ? 1.0f : 0.f);
            Xs[idx] = v;
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int t = tid + s * 256;
            if (t < ntile) {
                const int ni = t % ng, ki = t / ng;
#pragma unroll 4
                for (int m = 0; m < TBM; ++m) {
                    const float4 y = *reinterpret_cast<const float4*>(&Ys[m * NP + ni * 4]);
                    const float4 xa = *reinterpret_cast<const float4*>(&Xs[m * KP + ki * 8]);
                    const float4 xb = *reinterpret_cast<const float4*>(&Xs[m * KP + ki * 8 + 4]);
                    const float yv[4] = {y.x, y.y, y.z, y.w};
                    const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 8; ++b) acc[s][a][b] = fmaf(yv[a], xv[b], acc[s][a][b]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int t = tid + s * 256;
        if (t >= ntile) continue;
        const int ni = t % ng, ki = t / ng;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int n = ni * 4 + a;
            if (n >= N) continue;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int k = ki * 8 + b;
                const float v = acc[s][a][b];
                if (v == 0.f) continue;
                if (k < K) atomicAdd(dW + (size_t)n * ldw + k, v);
                else if (k == K && db) atomicAdd(db + n, v);
            }
        }
    }
}

// ------------------------------------------------------------------ positional encoding
// Column c of the encoded row (tensorBase.py:43-55): per source element e the
// block [sin(x*2^0..2^(F-1)), cos(x*2^0..2^(F-1))], each scaled by the
// annealing mask clamp(progress*F - l, 0, 1).
struct PEParams {
    int F;            // app_dim
    int fpe, vpe;     // frequency counts
    int mode;         // 0: MLP_Fea  [feat, dir, PE(feat), PE(dir)] -> out
                      // 1: WeakView [feat, PE(feat)] -> out ; [PE(dir)] -> out2 (cols 0..6*vpe)
    float fprog, vprog;
    int S;            // samples per ray (ray = sidx / S)
    int normalize_dir;
};

__device__ __forceinline__ float pe_value(float x, int r, int nf, float prog) {
    const int l = r % nf;
    const bool is_cos = r >= nf;
    const float m = fminf(fmaxf(prog * nf - (float)l, 0.f), 1.f);
    const float a = x * (float)(1 << l);
    return (is_cos ? cosf(a) : sinf(a)) * m;
}

__device__ __forceinline__ void load_dir(const float* __restrict__ rays_d, int ray, int normalize, float d[3]) {
    d[0] = rays_d[3 * ray]; d[1] = rays_d[3 * ray + 1]; d[2] = rays_d[3 * ray + 2];
    if (normalize) {
        const float n = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        d[0] /= n; d[1] /= n; d[2] /= n;
    }
}

__global__ void __launch_bounds__(256) pe_fwd_kernel(PEParams P, const float* __restrict__ feat, int ldf,
                                                     const int* __restrict__ aidx, const int* __restrict__ sidx,
                                                     const float* __restrict__ rays_d,
                                                     const int* __restrict__ n_dev, int n_fixed,
                                                     float* __restrict__ out, int ldo, float* __restrict__ out2,
                                                     int ldo2) {
    const int n = n_dev ? *n_dev : n_fixed;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int nfe = 2 * P.fpe * P.F, nve = 2 * P.vpe * 3;
    const int raw = P.F + (P.mode == 0 ? 3 : 0);
    const int main_cols = raw + nfe + (P.mode == 0 ? nve : 0);
    for (int a = warp; a < n; a += nwarps) {
        const int ray = sidx[aidx[a]] / P.S;
        float d[3];
        load_dir(rays_d, ray, P.normalize_dir, d);
        const float* f = feat + (size_t)a * ldf;
        for (int c = lane; c < ldo; c += 32) {
            float v = 0.f;
            if (c < P.F) v = f[c];
            else if (c < raw) v = d[c - P.F];
            else if (c < raw + nfe) { const int cc = c - raw; v = pe_value(f[cc / (2 * P.fpe)], cc % (2 * P.fpe), P.fpe, P.fprog); }
            else if (c < main_cols) { const int cc = c - raw - nfe; v = pe_value(d[cc / (2 * P.vpe)], cc % (2 * P.vpe), P.vpe, P.vprog); }
            out[(size_t)a * ldo + c] = v;
        }
        if (P.mode == 1 && lane < nve)
            out2[(size_t)a * ldo2 + lane] = pe_value(d[lane / (2 * P.vpe)], lane % (2 * P.vpe), P.vpe, P.vprog);
    }
}

// dfeat[a][e] = din[a][e] + sum_l 2^l m_l (cos(x 2^l) din_sin[l] - sin(x 2^l) din_cos[l])
__global__ void __launch_bounds__(256) pe_bwd_kernel(PEParams P, const float* __restrict__ feat, int ldf,
                                                     const float* __restrict__ din, int ldi,
                                                     const int* __restrict__ n_dev, int n_fixed,
                                                     float* __restrict__ dfeat, int ldd) {
    const int n = n_dev ? *n_dev : n_fixed;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int raw = P.F + (P.mode == 0 ? 3 : 0);
    for (int a = warp; a < n; a += nwarps) {
        const float* g = din + (size_t)a * ldi;
        for (int e = lane; e < ldd; e += 32) {
            float v = 0.f;
            if (e < P.F) {
                const float x = feat[(size_t)a * ldf + e];
                v = g[e];
                for (int l = 0; l < P.fpe; ++l) {
                    const float m = fminf(fmaxf(P.fprog * P.fpe - (float)l, 0.f), 1.f);
                    const float sc = (float)(1 << l);
                    float s, c;
                    sincosf(x * sc, &s, &c);
                    v += sc * m * (c * g[raw + e * 2 * P.fpe + l] - s * g[raw + e * 2 * P.fpe + P.fpe + l]);
                }
            }
            dfeat[(size_t)a * ldd + e] = v;
        }
    }
}

// ------------------------------------------------------------------ SH (deg 2) shading
// fwd: rgb[a][c] = relu(sum_k Y_k f[a][c*9+k] + 0.5); bwd: dfeat[a][c*9+k] = dout[a][c] * Y_k
__global__ void __launch_bounds__(256) sh_kernel(int bwd, const float* __restrict__ feat, int ldf,
                                                 const int* __restrict__ aidx, const int* __restrict__ sidx,
                                                 const float* __restrict__ rays_d, int S, int normalize_dir,
                                                 const int* __restrict__ n_dev, int n_fixed,
                                                 float* __restrict__ rgb, const float* __restrict__ dout,
                                                 float* __restrict__ dfeat, int ldd) {
    const int n = n_dev ? *n_dev : n_fixed;
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
        const int ray = sidx[aidx[a]] / S;
        float d[3], y[9];
        load_dir(rays_d, ray, normalize_dir, d);
        sh9(d, y);
        if (!bwd) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < 9; ++k) s += y[k] * feat[(size_t)a * ldf + c * 9 + k];
                rgb[4 * (size_t)a + c] = fmaxf(s + 0.5f, 0.f);
            }
            rgb[4 * (size_t)a + 3] = 0.f;
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int k = 0; k < 9; ++k) dfeat[(size_t)a * ldd + c * 9 + k] = dout[4 * (size_t)a + c] * y[k];
            for (int c = 27; c < ldd; ++c) dfeat[(size_t)a * ldd + c] = 0.f;
        }
    }
}

static int tiles_grid(int m_max, int per, int cap) {
    long long t = ((long long)m_max + per - 1) / per;
    if (t < 1) t = 1;
    return (int)(t < cap ? t : cap);
}

}  // namespace jt

using namespace jt;

extern "C" int jt_gemm_nt(const float* X, int ldx, const float* W, int ldw, int w_kn, const float* bias, float* Y,
                          int ldy, const float* mask, int ldm, const int* m_dev, int m_max, int N, int K, int act,
                          cudaStream_t stream) {
    JT_CHECK_ARG(X && W && Y && N > 0 && K > 0 && ldx >= K && ldy >= N);
    JT_CHECK_ARG((ldx & 3) == 0 && (ldy & 3) == 0 && act >= 0 && act <= 2);
    JT_CHECK_ARG((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0);
    if (m_max <= 0) return JT_OK;
    dim3 grid(tiles_grid(m_max, BM, kNumSMs * 4), (N + BN - 1) / BN);
    g_launches += 1;
    gemm_nt_kernel<<<grid, 256, 0, stream>>>(X, ldx, W, ldw, w_kn, bias, Y, ldy, mask, ldm, m_dev, m_max, N, K, act);
    JT_RETURN_LAUNCH();
}

extern "C" int jt_gemm_tn(const float* dY, int ldy, const float* X, int ldx, const int* m_dev, int m_max, int N,
                          int K, float* dW, int ldw, float* db, cudaStream_t stream) {
    JT_CHECK_ARG(dY && X && dW && N > 0 && K > 0 && ldy >= N && ldx >= K && ldw >= K);
    const int NP = (N + 3) & ~3, KP = (K + 1 + 7) & ~7;
    JT_CHECK_ARG((NP >> 2) * (KP >> 3) <= 512);
    if (m_max <= 0) return JT_OK;
    size_t smem = (size_t)TBM * (NP + KP) * sizeof(float);
    JT_CHECK_ARG(smem <= 48 * 1024);
    int grid = tiles_grid(m_max, TBM, kNumSMs * 2);
    g_launches += 1;
    gemm_tn_kernel<<<grid, 256, smem, stream>>>(dY, ldy, X, ldx, m_dev, m_max, N, K, dW, ldw, db);
    JT_RETURN_LAUNCH();
}

extern "C" int jt_pe_encode(int bwd, int app_dim, int fea_pe, int view_pe, int mode, float fea_progress,
                            float view_progress, int n_samples, int normalize_dir, const float* feat, int ldf,
                            const int* aidx, const int* sidx, const float* rays_d, const int* n_dev, int n_max,
                            float* out, int ldo, float* out2, int ldo2, const float* din, int ldi,
                            cudaStream_t stream) {
    JT_CHECK_ARG(feat && app_dim > 0 && fea_pe >= 0 && view_pe >= 0 && (mode == 0 || mode == 1));
    JT_CHECK_ARG(fea_pe <= 16 && view_pe <= 5);
    if (n_max <= 0) return JT_OK;
    PEParams P{app_dim, fea_pe, view_pe, mode, fea_progress, view_progress, n_samples, normalize_dir};
    int grid = tiles_grid(n_max, 8, kNumSMs * 8);
    g_launches += 1;
    if (!bwd) {
        JT_CHECK_ARG(aidx && sidx && rays_d && out && (mode == 0 || out2));
        const int cols = app_dim + (mode == 0 ? 3 : 0) + 2 * fea_pe * app_dim + (mode == 0 ? 6 * view_pe : 0);
        JT_CHECK_ARG(ldo >= cols && (mode == 0 || ldo2 >= 6 * view_pe));
        pe_fwd_kernel<<<grid, 256, 0, stream>>>(P, feat, ldf, aidx, sidx, rays_d, n_dev, n_max, out, ldo, out2, ldo2);
    } else {
        JT_CHECK_ARG(din && out && ldo >= app_dim);
        pe_bwd_kernel<<<grid, 256, 0, stream>>>(P, feat, ldf, din, ldi, n_dev, n_max, out, ldo);
    }
    JT_RETURN_LAUNCH();
}

extern "C" int jt_sh_shade(int bwd, const float* feat, int ldf, const int* aidx, const int* sidx,
                           const float* rays_d, int n_samples, int normalize_dir, const int* n_dev, int n_max,
                           float* rgb, const float* dout, float* dfeat, int ldd, cudaStream_t stream) {
    JT_CHECK_ARG(aidx && sidx && rays_d && ldf >= 27);
    JT_CHECK_ARG(bwd ? (dout && dfeat && ldd >= 27) : (feat && rgb));
    if (n_max <= 0) return JT_OK;
    int grid = tiles_grid(n_max, 256, kNumSMs * 8);
    g_launches += 1;
    sh_kernel<<<grid, 256, 0, stream>>>(bwd, feat, ldf, aidx, sidx, rays_d, n_samples, normalize_dir, n_dev, n_max,
                                        rgb, dout, dfeat, ldd);
    JT_RETURN_LAUNCH();
}
