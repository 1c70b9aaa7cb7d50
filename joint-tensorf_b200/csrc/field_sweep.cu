// Per-step full-factor sweeps: the regularisers the reference evaluates every
// training step on the VM factors and the optimiser update, as HBM-bound streaming kernels over the
// channel-last factor storage.
//   density_L1        tensoRF.py:212-216   sum_i mean|P_i| + mean|L_i|
//   TV_loss_density   tensoRF.py:218-222   sum_i 1e-2 * TVLoss(P_i)      (TVLoss: tensorBase.py:16-41)
//   TV_loss_app       tensoRF.py:224-228
//   torch.optim.Adam(betas=(0.9, 0.99))    tensorf.py:474-475, lr decay tensorf.py:431-436
// The reference runs ~10 ATen kernels per factor for the values, the same again in autograd, and the
// foreach Adam (6 passes over every tensor). Here: one launch for the values (x read once), one for
// the gradients (x read once, gradient read-modify-write), one for the update of every tensor.
#include "jt_common.cuh"
#include "../../include/jt_vm.h"

namespace jt {

constexpr int kMaxSweep = 12;     // the 12 VM factors of a field
constexpr int kMaxAdam = 32;      // tensors per multi-tensor Adam launch

struct SweepArr {
    const float* x;
    float* g;
    int H, W, Q;          // Q = C/4 float4 channel quads; memory is [H][W][Q] float4
    int term;             // 0: skip TV, 1: TV counted in the density slot, 2: in the appearance slot
    int l1;               // contributes mean|x| to L1
    float inv_n;          // 1 / numel
    float inv_ch, inv_cw; // 1 / count_h, 1 / count_w  (C*(H-1)*W, C*H*(W-1)); 0 when the count is 0
    int blk0;             // first block of this array in the launch
};
struct SweepArgs {
    SweepArr a[kMaxSweep];
    int n;
};

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
        t = warp_sum(t);
    }
    return t;     // valid in warp 0
}

__device__ __forceinline__ int find_arr(const SweepArgs& A) {
    int k = 0;
#pragma unroll 1
    for (int i = 1; i < A.n; ++i) if ((int)blockIdx.x >= A.a[i].blk0) k = i;
    return k;
}

constexpr int kSweepThreads = 256;
constexpr int kSweepPerThread = 8;   // float4 elements per thread
constexpr int kSweepPerBlock = kSweepThreads * kSweepPerThread;

// sums[i] = { sum|x|, sum_h (x[h+1]-x[h])^2, sum_w (x[w+1]-x[w])^2 } of array i, as doubles.
__global__ void __launch_bounds__(kSweepThreads) reg_sums_kernel(const __grid_constant__ SweepArgs A, double* sums) {
    __shared__ float red[kSweepThreads / 32];
    const int k = find_arr(A);
    const SweepArr& a = A.a[k];
    const int WQ = a.W * a.Q;
    const long long total = (long long)a.H * WQ;
    const long long base = (long long)(blockIdx.x - a.blk0) * kSweepPerBlock;
    const float4* x4 = reinterpret_cast<const float4*>(a.x);
    float s1 = 0.0f, sh = 0.0f, sw = 0.0f;
#pragma unroll
    for (int j = 0; j < kSweepPerThread; ++j) {
        long long idx = base + j * kSweepThreads + threadIdx.x;
        if (idx >= total) break;
        int h = (int)(idx / WQ);
        int r = (int)(idx - (long long)h * WQ);
        int w = r / a.Q;
        float4 v = __ldg(x4 + idx);
        s1 += (fabsf(v.x) + fabsf(v.y)) + (fabsf(v.z) + fabsf(v.w));
        if (a.term) {
            if (h + 1 < a.H) {
                float4 d = __ldg(x4 + idx + WQ);
                float dx = d.x - v.x, dy = d.y - v.y, dz = d.z - v.z, dw = d.w - v.w;
                sh += (dx * dx + dy * dy) + (dz * dz + dw * dw);
            }
            if (w + 1 < a.W) {
                float4 d = __ldg(x4 + idx + a.Q);
                float dx = d.x - v.x, dy = d.y - v.y, dz = d.z - v.z, dw = d.w - v.w;
                sw += (dx * dx + dy * dy) + (dz * dz + dw * dw);
            }
        }
    }
    float t1 = block_sum(s1, red);
    float th = block_sum(sh, red);
    float tw = block_sum(sw, red);
    if (threadIdx.x == 0) {
        if (a.l1) atomicAdd(sums + 3 * k + 0, (double)t1);
        if (a.term) {
            atomicAdd(sums + 3 * k + 1, (double)th);
            atomicAdd(sums + 3 * k + 2, (double)tw);
        }
    }
}

// out = { L1, TV_density, TV_app } as the reference's three scalar losses (fp32).
__global__ void reg_finalize_kernel(const __grid_constant__ SweepArgs A, const double* sums, float* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double l1 = 0.0, tv[2] = {0.0, 0.0};
    for (int i = 0; i < A.n; ++i) {
        const SweepArr& a = A.a[i];
        if (a.l1) l1 += sums[3 * i] * (double)a.inv_n;
        if (a.term) tv[a.term - 1] += 1e-2 * 2.0 * (sums[3 * i + 1] * (double)a.inv_ch + sums[3 * i + 2] * (double)a.inv_cw);
    }
    out[0] = (float)l1;
    out[1] = (float)tv[0];
    out[2] = (float)tv[1];
}

struct RegCoef {
    float h[3];           // host-side weights of {L1, TV_density, TV_app}
    const float* up;      // optional device upstream gradients [3] (autograd), multiplied in
};

// g += c_L1 * sign(x) / numel + c_TV * 1e-2 * 2 * d/dx [ sum_h (.)^2 / count_h + sum_w (.)^2 / count_w ]
__global__ void __launch_bounds__(kSweepThreads) reg_grad_kernel(const __grid_constant__ SweepArgs A, const RegCoef cf) {
    const int k = find_arr(A);
    const SweepArr& a = A.a[k];
    const int WQ = a.W * a.Q;
    const long long total = (long long)a.H * WQ;
    const long long base = (long long)(blockIdx.x - a.blk0) * kSweepPerBlock;
    const float4* x4 = reinterpret_cast<const float4*>(a.x);
    float4* g4 = reinterpret_cast<float4*>(a.g);
    float c1 = a.l1 ? cf.h[0] * (cf.up ? __ldg(cf.up) : 1.0f) * a.inv_n : 0.0f;
    float ct = a.term ? (a.term == 1 ? cf.h[1] : cf.h[2]) * (cf.up ? __ldg(cf.up + a.term) : 1.0f) * (1e-2f * 2.0f * 2.0f) : 0.0f;
    const float chh = ct * a.inv_ch, cww = ct * a.inv_cw;
    const bool tv = a.term != 0 && ct != 0.0f;
#pragma unroll
    for (int j = 0; j < kSweepPerThread; ++j) {
        long long idx = base + j * kSweepThreads + threadIdx.x;
        if (idx >= total) break;
        int h = (int)(idx / WQ);
        int r = (int)(idx - (long long)h * WQ);
        int w = r / a.Q;
        float4 v = __ldg(x4 + idx);
        float4 g = g4[idx];
        auto sgn = [](float t) { return t > 0.0f ? 1.0f : (t < 0.0f ? -1.0f : 0.0f); };
        g.x += c1 * sgn(v.x); g.y += c1 * sgn(v.y); g.z += c1 * sgn(v.z); g.w += c1 * sgn(v.w);
        if (tv) {
            // d/dx[h] of sum (x[h+1]-x[h])^2 = 2(x[h]-x[h-1]) - 2(x[h+1]-x[h])   (the 2 is inside ct)
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (h > 0) { float4 u = __ldg(x4 + idx - WQ);
                acc.x += chh * (v.x - u.x); acc.y += chh * (v.y - u.y); acc.z += chh * (v.z - u.z); acc.w += chh * (v.w - u.w); }
            if (h + 1 < a.H) { float4 d = __ldg(x4 + idx + WQ);
                acc.x -= chh * (d.x - v.x); acc.y -= chh * (d.y - v.y); acc.z -= chh * (d.z - v.z); acc.w -= chh * (d.w - v.w); }
            if (w > 0) { float4 u = __ldg(x4 + idx - a.Q);
                acc.x += cww * (v.x - u.x); acc.y += cww * (v.y - u.y); acc.z += cww * (v.z - u.z); acc.w += cww * (v.w - u.w); }
            if (w + 1 < a.W) { float4 d = __ldg(x4 + idx + a.Q);
                acc.x -= cww * (d.x - v.x); acc.y -= cww * (d.y - v.y); acc.z -= cww * (d.z - v.z); acc.w -= cww * (d.w - v.w); }
            g.x += acc.x; g.y += acc.y; g.z += acc.z; g.w += acc.w;
        }
        g4[idx] = g;
    }
}

static int fill_sweep(SweepArgs& A, int n, const void* const* h_x, void* const* h_g, const int* h_dims,
                      const int* h_term, const int* h_l1, bool need_g, int* n_blocks) {
    if (n < 1 || n > kMaxSweep || !h_x || !h_dims || !h_term || !h_l1) return JT_ERR_ARG;
    int blk = 0;
    A.n = 0;
    for (int i = 0; i < n; ++i) {
        SweepArr& a = A.a[A.n];
        int H = h_dims[3 * i], W = h_dims[3 * i + 1], C = h_dims[3 * i + 2];
        if (H < 1 || W < 1 || C < 4 || (C & 3) || !h_x[i]) return JT_ERR_ARG;
        if ((reinterpret_cast<uintptr_t>(h_x[i]) & 15) != 0) return JT_ERR_ARG;
        if (need_g && (!h_g || !h_g[i] || (reinterpret_cast<uintptr_t>(h_g[i]) & 15) != 0)) return JT_ERR_ARG;
        if (h_term[i] < 0 || h_term[i] > 2) return JT_ERR_ARG;
        if (!h_term[i] && !h_l1[i]) continue;
        a.x = static_cast<const float*>(h_x[i]);
        a.g = need_g ? static_cast<float*>(h_g[i]) : nullptr;
        a.H = H; a.W = W; a.Q = C / 4;
        a.term = h_term[i]; a.l1 = h_l1[i] ? 1 : 0;
        double numel = (double)H * W * C;
        a.inv_n = (float)(1.0 / numel);
        double ch = (double)C * (H - 1) * W, cw = (double)C * H * (W - 1);
        a.inv_ch = ch > 0 ? (float)(1.0 / ch) : 0.0f;
        a.inv_cw = cw > 0 ? (float)(1.0 / cw) : 0.0f;
        a.blk0 = blk;
        long long q = (long long)H * W * (C / 4);
        blk += (int)((q + kSweepPerBlock - 1) / kSweepPerBlock);
        ++A.n;     // sums_ws is indexed by the position in A (arrays with no term are dropped)
    }
    *n_blocks = blk;
    return JT_OK;
}

// ------------------------------------------------------------------------------------------- Adam
struct AdamTensor {
    float* p; float* g; float* m; float* v;
    long long n;
    float step_size;      // lr / (1 - beta1^t)
    float inv_bc2_sqrt;   // 1 / sqrt(1 - beta2^t)
    int blk0;
};
struct AdamArgs {
    AdamTensor t[kMaxAdam];
    int n;
    float omb1, beta2, omb2, eps, grad_scale;   // 1-beta1, beta2, 1-beta2 rounded from double as torch does
    int zero_grad;
};

constexpr int kAdamThreads = 256;
constexpr int kAdamPerThread = 4;    // float4 per thread
constexpr int kAdamPerBlock = kAdamThreads * kAdamPerThread * 4;   // floats per block

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float omb1, float b2, float omb2,
                                         float eps, float step_size, float inv_bc2_sqrt) {
    // torch/optim/adam.py _single_tensor_adam: exp_avg.lerp_(grad, 1-b1); exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2);
    // denom = sqrt(exp_avg_sq) / sqrt(bc2) + eps; param.addcdiv_(exp_avg, denom, value=-step_size)
    m = m + omb1 * (g - m);
    v = v * b2 + omb2 * g * g;
    float denom = sqrtf(v) * inv_bc2_sqrt + eps;
    p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(kAdamThreads) adam_multi_kernel(const __grid_constant__ AdamArgs A) {
    int k = 0;
#pragma unroll 1
    for (int i = 1; i < A.n; ++i) if ((int)blockIdx.x >= A.t[i].blk0) k = i;
    const AdamTensor& t = A.t[k];
    const long long base = (long long)(blockIdx.x - t.blk0) * kAdamPerBlock;
    const float o1 = A.omb1, b2 = A.beta2, o2 = A.omb2, eps = A.eps, gs = A.grad_scale;
    const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                       reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15) == 0;
#pragma unroll
    for (int j = 0; j < kAdamPerThread; ++j) {
        long long e = base + ((long long)j * kAdamThreads + threadIdx.x) * 4;
        if (e >= t.n) break;
        if (vec && e + 4 <= t.n) {
            float4 p = *reinterpret_cast<float4*>(t.p + e);
            float4 g = *reinterpret_cast<const float4*>(t.g + e);
            float4 m = *reinterpret_cast<float4*>(t.m + e);
            float4 v = *reinterpret_cast<float4*>(t.v + e);
            g.x *= gs; g.y *= gs; g.z *= gs; g.w *= gs;
            adam_one(p.x, g.x, m.x, v.x, o1, b2, o2, eps, t.step_size, t.inv_bc2_sqrt);
            adam_one(p.y, g.y, m.y, v.y, o1, b2, o2, eps, t.step_size, t.inv_bc2_sqrt);
            adam_one(p.z, g.z, m.z, v.z, o1, b2, o2, eps, t.step_size, t.inv_bc2_sqrt);
            adam_one(p.w, g.w, m.w, v.w, o1, b2, o2, eps, t.step_size, t.inv_bc2_sqrt);
            *reinterpret_cast<float4*>(t.p + e) = p;
            *reinterpret_cast<float4*>(t.m + e) = m;
            *reinterpret_cast<float4*>(t.v + e) = v;
            if (A.zero_grad) *reinterpret_cast<float4*>(t.g + e) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (long long q = e; q < e + 4 && q < t.n; ++q) {
                float p = t.p[q], g = t.g[q] * gs, m = t.m[q], v = t.v[q];
                adam_one(p, g, m, v, o1, b2, o2, eps, t.step_size, t.inv_bc2_sqrt);
                t.p[q] = p; t.m[q] = m; t.v[q] = v;
                if (A.zero_grad) t.g[q] = 0.0f;
            }
        }
    }
}

}  // namespace jt

using namespace jt;

extern "C" int jt_reg_values(int n_arrays, const void* const* h_x, const int* h_dims, const int* h_term,
                             const int* h_l1, double* sums_ws, float* out3, cudaStream_t stream) {
    JT_CHECK_ARG(sums_ws && out3);
    SweepArgs A;
    int nb = 0;
    int rc = fill_sweep(A, n_arrays, h_x, nullptr, h_dims, h_term, h_l1, false, &nb);
    if (rc != JT_OK) return rc;
    if (cudaMemsetAsync(sums_ws, 0, sizeof(double) * 3 * kMaxSweep, stream) != cudaSuccess) return JT_ERR_LAUNCH;
    if (nb > 0) { reg_sums_kernel<<<nb, kSweepThreads, 0, stream>>>(A, sums_ws); ++g_launches; }
    reg_finalize_kernel<<<1, 32, 0, stream>>>(A, sums_ws, out3);
    ++g_launches;
    JT_RETURN_LAUNCH();
}

extern "C" int jt_reg_grads(int n_arrays, const void* const* h_x, void* const* h_g, const int* h_dims,
                            const int* h_term, const int* h_l1, const float* h_coef3, const float* up3,
                            cudaStream_t stream) {
    JT_CHECK_ARG(h_coef3);
    SweepArgs A;
    int nb = 0;
    // arrays whose weight is zero on the host are not swept at all
    int term[kMaxSweep], l1[kMaxSweep];
    if (n_arrays < 1 || n_arrays > kMaxSweep || !h_term || !h_l1) return JT_ERR_ARG;
    for (int i = 0; i < n_arrays; ++i) {
        JT_CHECK_ARG(h_term[i] >= 0 && h_term[i] <= 2);
        term[i] = (h_term[i] && h_coef3[h_term[i]] != 0.0f) ? h_term[i] : 0;
        l1[i] = (h_l1[i] && h_coef3[0] != 0.0f) ? 1 : 0;
    }
    int rc = fill_sweep(A, n_arrays, h_x, h_g, h_dims, term, l1, true, &nb);
    if (rc != JT_OK) return rc;
    if (nb == 0) return JT_OK;
    RegCoef cf;
    cf.h[0] = h_coef3[0]; cf.h[1] = h_coef3[1]; cf.h[2] = h_coef3[2];
    cf.up = up3;
    reg_grad_kernel<<<nb, kSweepThreads, 0, stream>>>(A, cf);
    ++g_launches;
    JT_RETURN_LAUNCH();
}

extern "C" int jt_adam_multi(int n_tensors, void* const* h_p, void* const* h_g, void* const* h_m, void* const* h_v,
                             const long long* h_numel, const double* h_lr, const int* h_step, double beta1, double beta2,
                             double eps, double grad_scale, int zero_grad, cudaStream_t stream) {
    JT_CHECK_ARG(n_tensors >= 0 && h_p && h_g && h_m && h_v && h_numel && h_lr && h_step);
    int done = 0;
    while (done < n_tensors) {
        AdamArgs A;
        A.n = 0; A.omb1 = (float)(1.0 - beta1); A.beta2 = (float)beta2; A.omb2 = (float)(1.0 - beta2);
        A.eps = (float)eps; A.grad_scale = (float)grad_scale; A.zero_grad = zero_grad;
        long long blk = 0;
        while (done < n_tensors && A.n < kMaxAdam) {
            int i = done++;
            if (h_numel[i] == 0) continue;
            JT_CHECK_ARG(h_p[i] && h_g[i] && h_m[i] && h_v[i] && h_numel[i] > 0 && h_step[i] >= 1);
            AdamTensor& t = A.t[A.n++];
            t.p = static_cast<float*>(h_p[i]); t.g = static_cast<float*>(h_g[i]);
            t.m = static_cast<float*>(h_m[i]); t.v = static_cast<float*>(h_v[i]);
            t.n = h_numel[i];
            // bias corrections in double on the host, as python floats in torch/optim/adam.py
            double bc1 = 1.0 - pow(beta1, (double)h_step[i]);
            double bc2 = 1.0 - pow(beta2, (double)h_step[i]);
            t.step_size = (float)(h_lr[i] / bc1);
            t.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
            t.blk0 = (int)blk;
            blk += (t.n + kAdamPerBlock - 1) / kAdamPerBlock;
            JT_CHECK_ARG(blk < 0x7fffffffLL);
        }
        if (A.n == 0) break;
        adam_multi_kernel<<<(int)blk, kAdamThreads, 0, stream>>>(A);
        ++g_launches;
    }
    JT_RETURN_LAUNCH();
}
