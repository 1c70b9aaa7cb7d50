// Pose -> ray generation for the sampled pixels only, with its backward to the se(3) refinement.
//
// Replaces, for one training / test-time-optimisation step:
//   model/bat.py:350-353      pose = compose([lie.se3_to_SE3(se3_refine[idx]), pose])
//   camera.py:81-99           Lie.se3_to_SE3 (Taylor series of sin x / x, (1-cos x)/x^2, (x-sin x)/x^3, nth = 8)
//   camera.py:43-58           Pose.compose_pair
//   model/tensorf.py:144-166  camera.get_center_and_ray for ALL H*W pixels of ALL views, then [:, ray_idx]
//   camera.py:231-261         get_center_and_ray: grid_3D = [x+.5, y+.5, 1] K^-T; ray = grid_3D R; center = -(t^T R)
//   camera.py:303-340         convert_NDC
// and the autograd of all of it. The reference materialises B*H*W rays per step (and per render slice) and
// indexes them afterwards; here only the B*R requested rays are ever computed (24 B written per ray), and the
// backward reduces the 12 numbers per view the pose gradient needs before a per-view closed-form chain rule.
#include "jt_common.cuh"
#include "../../include/jt_vm.h"

namespace jt {

constexpr int PR_THREADS = 256;
constexpr int PR_NTH = 8;          // camera.py:92-94 (nth = 8)

struct PoseRaysArgs {
    const float* se3;        // [B][6] or nullptr (pose = base)
    const int* view_idx;     // [B] row of se3 used by view b (se3_refine.weight[var.idx]) or nullptr (row b)
    const float* base;       // [B][3][4] (or one pose when base_stride == 0)
    int base_stride;         // 12 or 0
    const float* intr_inv;   // [B][3][3] (or one matrix when kinv_stride == 0)
    int kinv_stride;
    const float* intr;       // NDC only: [B][3][3]
    int k_stride;
    const int* pix;          // [R] shared by the views, [B][R] when pix_per_view, or nullptr: pixel = pix_base + r
    int pix_per_view, pix_base;
    int B, R, W;
    int ndc, center_shift, detach_shift;
    float near_;
};

// Taylor coefficients of camera.py:121-145 in powers of theta^2; k = 0: A, 1: B, 2: C.
struct Series { float A, B, C, dA, dB, dC; };      // dX = X'(theta) / theta
__device__ __forceinline__ Series taylor_series(float th2) {
    Series s = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dA = 1.f, dB = 1.f, dC = 1.f, p = 1.f, pm = 0.f;     // p = theta^(2i), pm = theta^(2i-2)
    float sgn = 1.f;
#pragma unroll
    for (int i = 0; i <= PR_NTH; ++i) {
        if (i > 0) dA *= (float)((2 * i) * (2 * i + 1));
        dB *= (float)((2 * i + 1) * (2 * i + 2));
        dC *= (float)((2 * i + 2) * (2 * i + 3));
        s.A += sgn * (p / dA); s.B += sgn * (p / dB); s.C += sgn * (p / dC);
        // d/dtheta theta^(2i) = 2i theta^(2i-1); divided by theta -> 2i theta^(2i-2) (0 for i = 0)
        s.dA += sgn * ((float)(2 * i) * pm / dA); s.dB += sgn * ((float)(2 * i) * pm / dB); s.dC += sgn * ((float)(2 * i) * pm / dC);
        pm = p; p *= th2; sgn = -sgn;
    }
    return s;
}

__device__ __forceinline__ void mat3_mul(const float* a, const float* b, float* c) {          // c = a b
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
__device__ __forceinline__ void skew(const float* w, float* wx) {
    wx[0] = 0.f; wx[1] = -w[2]; wx[2] = w[1];
    wx[3] = w[2]; wx[4] = 0.f; wx[5] = -w[0];
    wx[6] = -w[1]; wx[7] = w[0]; wx[8] = 0.f;
}

// se3 (w, u) -> refine rotation Ra [9], V [9], translation ta [3] (camera.py:81-99)
__device__ __forceinline__ void se3_exp(const float* wu, float* Ra, float* V, float* ta, float* wx, float* wx2, Series& s) {
    skew(wu, wx);
    mat3_mul(wx, wx, wx2);
    s = taylor_series(wu[0] * wu[0] + wu[1] * wu[1] + wu[2] * wu[2]);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float I = (k == 0 || k == 4 || k == 8) ? 1.f : 0.f;
        Ra[k] = I + s.A * wx[k] + s.B * wx2[k];
        V[k] = I + s.B * wx[k] + s.C * wx2[k];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) ta[i] = V[3 * i] * wu[3] + V[3 * i + 1] * wu[4] + V[3 * i + 2] * wu[5];
}

// pose of view b: R [9], t [3] (world-to-camera), after the refinement (bat.py:350-353, camera.py:50-58)
__device__ __forceinline__ void view_pose(const PoseRaysArgs& A, int b, float* R, float* t) {
    const float* base = A.base + (size_t)b * A.base_stride;
    float Rb[9], tb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) Rb[3 * i + j] = base[4 * i + j];
        tb[i] = base[4 * i + 3];
    }
    if (A.se3 == nullptr) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = Rb[k];
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = tb[i];
        return;
    }
    const int row = A.view_idx ? A.view_idx[b] : b;
    float wu[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) wu[k] = A.se3[(size_t)row * 6 + k];
    float Ra[9], V[9], ta[3], wx[9], wx2[9];
    Series s;
    se3_exp(wu, Ra, V, ta, wx, wx2, s);
    mat3_mul(Rb, Ra, R);                                            // R_new = R_b R_a
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = Rb[3 * i] * ta[0] + Rb[3 * i + 1] * ta[1] + Rb[3 * i + 2] * ta[2] + tb[i];
}

__device__ __forceinline__ void pixel_cam(const PoseRaysArgs& A, int b, int r, float* g) {
    const int p = A.pix ? A.pix[A.pix_per_view ? (size_t)b * A.R + r : r] : A.pix_base + r;
    const float X = (float)(p % A.W) + 0.5f, Y = (float)(p / A.W) + 0.5f;
    const float* Ki = A.intr_inv + (size_t)b * A.kinv_stride;
#pragma unroll
    for (int i = 0; i < 3; ++i) g[i] = X * Ki[3 * i] + Y * Ki[3 * i + 1] + Ki[3 * i + 2];      // to_hom + img2cam
}

__global__ void __launch_bounds__(PR_THREADS) pose_rays_fwd_kernel(const PoseRaysArgs A, float* __restrict__ center,
                                                                   float* __restrict__ ray, float* __restrict__ pose_out) {
    __shared__ float sp[12];
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        float R[9], t[3];
        view_pose(A, b, R, t);
#pragma unroll
        for (int k = 0; k < 9; ++k) sp[k] = R[k];
#pragma unroll
        for (int i = 0; i < 3; ++i) sp[9 + i] = t[i];
        if (pose_out && blockIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) pose_out[b * 12 + 4 * i + j] = R[3 * i + j];
                pose_out[b * 12 + 4 * i + 3] = t[i];
            }
        }
    }
    __syncthreads();
    const int r = blockIdx.x * PR_THREADS + threadIdx.x;
    if (r >= A.R) return;
    float g[3];
    pixel_cam(A, b, r, g);
    float c[3], d[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        d[j] = g[0] * sp[j] + g[1] * sp[3 + j] + g[2] * sp[6 + j];                              // grid_3D @ R
        c[j] = -(sp[9] * sp[j] + sp[10] * sp[3 + j] + sp[11] * sp[6 + j]);                      // -(t^T R)
    }
    if (A.ndc) {                                                                                // camera.py:303-340
        const float* K = A.intr + (size_t)b * A.k_stride;
        const float sx = K[0] / K[2], sy = K[4] / K[5];
        if (A.center_shift) {
            const float s = (A.near_ - c[2]) / d[2];
#pragma unroll
            for (int j = 0; j < 3; ++j) c[j] = c[j] + s * d[j];
        }
        const float cxoz = c[0] / c[2], cyoz = c[1] / c[2], rxoz = d[0] / d[2], ryoz = d[1] / d[2];
        const float cn[3] = {sx * cxoz, sy * cyoz, 1.f - 2.f * A.near_ / c[2]};
        const float rn[3] = {sx * (rxoz - cxoz), sy * (ryoz - cyoz), 2.f * A.near_ / c[2]};
#pragma unroll
        for (int j = 0; j < 3; ++j) { c[j] = cn[j]; d[j] = rn[j]; }
    }
    const size_t o = ((size_t)b * A.R + r) * 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) { center[o + j] = c[j]; ray[o + j] = d[j]; }
}

// Per view: M[i][j] = sum_r g_i d_ray_j (9) and dcs[j] = sum_r d_center_j (3), in world (pre-NDC) terms.
__global__ void __launch_bounds__(PR_THREADS) pose_rays_bwd_reduce_kernel(const PoseRaysArgs A, const float* __restrict__ d_center,
                                                                          const float* __restrict__ d_ray, float* __restrict__ acc12) {
    __shared__ float sp[12];
    __shared__ float red[PR_THREADS / 32][12];
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        float R[9], t[3];
        view_pose(A, b, R, t);
#pragma unroll
        for (int k = 0; k < 9; ++k) sp[k] = R[k];
#pragma unroll
        for (int i = 0; i < 3; ++i) sp[9 + i] = t[i];
    }
    __syncthreads();
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = 0.f;
    for (int r = blockIdx.x * PR_THREADS + threadIdx.x; r < A.R; r += gridDim.x * PR_THREADS) {
        float g[3];
        pixel_cam(A, b, r, g);
        const size_t o = ((size_t)b * A.R + r) * 3;
        float dc[3] = {d_center[o], d_center[o + 1], d_center[o + 2]};
        float dd[3] = {d_ray[o], d_ray[o + 1], d_ray[o + 2]};
        if (A.ndc) {
            float c[3], d[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                d[j] = g[0] * sp[j] + g[1] * sp[3 + j] + g[2] * sp[6 + j];
                c[j] = -(sp[9] * sp[j] + sp[10] * sp[3 + j] + sp[11] * sp[6 + j]);
            }
            const float* K = A.intr + (size_t)b * A.k_stride;
            const float sx = K[0] / K[2], sy = K[4] / K[5];
            float s = 0.f;
            const float cz0 = c[2];
            if (A.center_shift) {
                s = (A.near_ - c[2]) / d[2];
#pragma unroll
                for (int j = 0; j < 3; ++j) c[j] = c[j] + s * d[j];
            }
            const float icz = 1.f / c[2], irz = 1.f / d[2];
            const float g_cxoz = sx * (dc[0] - dd[0]), g_cyoz = sy * (dc[1] - dd[1]);
            const float g_rxoz = sx * dd[0], g_ryoz = sy * dd[1];
            float gc[3], gr[3];
            gc[0] = g_cxoz * icz; gc[1] = g_cyoz * icz;
            gc[2] = (dc[2] - dd[2]) * 2.f * A.near_ * icz * icz - (g_cxoz * c[0] + g_cyoz * c[1]) * icz * icz;
            gr[0] = g_rxoz * irz; gr[1] = g_ryoz * irz;
            gr[2] = -(g_rxoz * d[0] + g_ryoz * d[1]) * irz * irz;
            if (A.center_shift && !A.detach_shift) {            // c' = c0 + s ray, s = (near - cz0) / rz
                const float ds = gc[0] * d[0] + gc[1] * d[1] + gc[2] * d[2];
#pragma unroll
                for (int j = 0; j < 3; ++j) gr[j] += s * gc[j];
                gr[2] += -ds * s * irz;
                gc[2] += -ds * irz;
                (void)cz0;
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) { dc[j] = gc[j]; dd[j] = gr[j]; }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) v[3 * i + j] = fmaf(g[i], dd[j], v[3 * i + j]);
#pragma unroll
        for (int j = 0; j < 3; ++j) v[9 + j] += dc[j];
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = warp_sum(v[k]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 12; ++k) red[warp][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < PR_THREADS / 32; ++w) s += red[w][threadIdx.x];
        if (gridDim.x == 1) acc12[b * 12 + threadIdx.x] = s;
        else atomicAdd(acc12 + b * 12 + threadIdx.x, s);
    }
}

// Per view: chain rule through get_center_and_ray, compose_pair and se3_to_SE3.
__global__ void pose_rays_bwd_pose_kernel(const PoseRaysArgs A, const float* __restrict__ acc12, float* __restrict__ d_se3,
                                          float* __restrict__ d_pose) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= A.B) return;
    const float* base = A.base + (size_t)b * A.base_stride;
    float Rb[9], tb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) Rb[3 * i + j] = base[4 * i + j];
        tb[i] = base[4 * i + 3];
    }
    float R[9], t[3], Ra[9], V[9], ta[3], wx[9], wx2[9], wu[6];
    Series s;
    const int row = A.view_idx ? A.view_idx[b] : b;
    if (A.se3) {
#pragma unroll
        for (int k = 0; k < 6; ++k) wu[k] = A.se3[(size_t)row * 6 + k];
        se3_exp(wu, Ra, V, ta, wx, wx2, s);
        mat3_mul(Rb, Ra, R);
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = Rb[3 * i] * ta[0] + Rb[3 * i + 1] * ta[1] + Rb[3 * i + 2] * ta[2] + tb[i];
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = Rb[k];
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = tb[i];
    }
    const float* M = acc12 + b * 12;
    const float* dcs = M + 9;
    // ray = g R, center = -(t^T R):  dR[i][j] = M[i][j] - t_i dcs_j,  dt_i = -sum_j R[i][j] dcs_j
    float dR[9], dt[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) dR[3 * i + j] = M[3 * i + j] - t[i] * dcs[j];
        dt[i] = -(R[3 * i] * dcs[0] + R[3 * i + 1] * dcs[1] + R[3 * i + 2] * dcs[2]);
    }
    if (d_pose) {                          // gradient w.r.t. the composed pose [B][3][4] (diagnostics / callers that own the pose)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) d_pose[b * 12 + 4 * i + j] = dR[3 * i + j];
            d_pose[b * 12 + 4 * i + 3] = dt[i];
        }
    }
    if (!A.se3 || !d_se3) return;
    // compose_pair(refine, base): R = Rb Ra, t = Rb ta + tb  ->  dRa = Rb^T dR, dta = Rb^T dt
    float dRa[9], dta[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) dRa[3 * i + j] = Rb[i] * dR[j] + Rb[3 + i] * dR[3 + j] + Rb[6 + i] * dR[6 + j];
        dta[i] = Rb[i] * dt[0] + Rb[3 + i] * dt[1] + Rb[6 + i] * dt[2];
    }
    // ta = V u
    float du[3], dV[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        du[i] = V[i] * dta[0] + V[3 + i] * dta[1] + V[6 + i] * dta[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) dV[3 * i + j] = dta[i] * wu[3 + j];
    }
    // Ra = I + A wx + B wx^2, V = I + B wx + C wx^2
    float gA = 0.f, gB = 0.f, gC = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) { gA += dRa[k] * wx[k]; gB += dRa[k] * wx2[k] + dV[k] * wx[k]; gC += dV[k] * wx2[k]; }
    // G2 = upstream of wx^2 = B dRa + C dV;  d(wx) = A dRa + B dV + G2 wx^T + wx^T G2
    float G2[9], dwx[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { G2[k] = s.B * dRa[k] + s.C * dV[k]; dwx[k] = s.A * dRa[k] + s.B * dV[k]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) a += G2[3 * i + k] * wx[3 * j + k] + wx[3 * k + i] * G2[3 * k + j];
            dwx[3 * i + j] += a;
        }
    float dw[3] = {dwx[7] - dwx[5], dwx[2] - dwx[6], dwx[3] - dwx[1]};
    // theta = |w|: d theta / d w = w / theta, folded into the series as X'(theta)/theta
    const float gth = gA * s.dA + gB * s.dB + gC * s.dC;
#pragma unroll
    for (int i = 0; i < 3; ++i) dw[i] += gth * wu[i];
    // several views may share one se3 row only if view_idx repeats; rows are distinct in the reference
#pragma unroll
    for (int i = 0; i < 3; ++i) { atomicAdd(d_se3 + (size_t)row * 6 + i, dw[i]); atomicAdd(d_se3 + (size_t)row * 6 + 3 + i, du[i]); }
}

static int fill_args(PoseRaysArgs& A, const float* se3, const int* view_idx, const float* base, int base_per_view,
                     const float* intr_inv, int kinv_per_view, const float* intr, int k_per_view, const int* pix,
                     int pix_per_view, int pix_base, int n_views, int n_rays_per_view, int width, int ndc,
                     int center_shift, int detach_shift, float near_) {
    if (!base || !intr_inv || n_views < 1 || n_rays_per_view < 1 || width < 1) return JT_ERR_ARG;
    if (ndc && !intr) return JT_ERR_ARG;
    if (n_views > 65535) return JT_ERR_ARG;
    A.se3 = se3; A.view_idx = view_idx; A.base = base; A.base_stride = base_per_view ? 12 : 0;
    A.intr_inv = intr_inv; A.kinv_stride = kinv_per_view ? 9 : 0; A.intr = intr; A.k_stride = k_per_view ? 9 : 0;
    A.pix = pix; A.pix_per_view = pix_per_view; A.pix_base = pix_base;
    A.B = n_views; A.R = n_rays_per_view; A.W = width;
    A.ndc = ndc; A.center_shift = center_shift; A.detach_shift = detach_shift; A.near_ = near_;
    return JT_OK;
}

}  // namespace jt

using namespace jt;

extern "C" int jt_pose_rays_fwd(const float* se3, const int* view_idx, const float* base, int base_per_view,
                                const float* intr_inv, int kinv_per_view, const float* intr, int k_per_view,
                                const int* pix, int pix_per_view, int pix_base, int n_views, int n_rays_per_view,
                                int width, int ndc, int center_shift, int detach_shift, float near_plane,
                                float* center, float* ray, float* pose_out, cudaStream_t stream) {
    PoseRaysArgs A;
    if (int rc = fill_args(A, se3, view_idx, base, base_per_view, intr_inv, kinv_per_view, intr, k_per_view, pix,
                           pix_per_view, pix_base, n_views, n_rays_per_view, width, ndc, center_shift, detach_shift,
                           near_plane)) return rc;
    JT_CHECK_ARG(center && ray);
    dim3 grid((n_rays_per_view + PR_THREADS - 1) / PR_THREADS, n_views);
    g_launches += 1;
    pose_rays_fwd_kernel<<<grid, PR_THREADS, 0, stream>>>(A, center, ray, pose_out);
    JT_RETURN_LAUNCH();
}

extern "C" int jt_pose_rays_bwd(const float* se3, const int* view_idx, const float* base, int base_per_view,
                                const float* intr_inv, int kinv_per_view, const float* intr, int k_per_view,
                                const int* pix, int pix_per_view, int pix_base, int n_views, int n_rays_per_view,
                                int width, int ndc, int center_shift, int detach_shift, float near_plane,
                                const float* d_center, const float* d_ray, float* scratch12, float* d_se3,
                                float* d_pose, cudaStream_t stream) {
    PoseRaysArgs A;
    if (int rc = fill_args(A, se3, view_idx, base, base_per_view, intr_inv, kinv_per_view, intr, k_per_view, pix,
                           pix_per_view, pix_base, n_views, n_rays_per_view, width, ndc, center_shift, detach_shift,
                           near_plane)) return rc;
    JT_CHECK_ARG(d_center && d_ray && scratch12 && (d_se3 || d_pose));
    // chunks of rays per view: enough CTAs to fill the machine when there are few views
    int chunks = (n_rays_per_view + PR_THREADS * 4 - 1) / (PR_THREADS * 4);
    const int want = (2 * kNumSMs + n_views - 1) / n_views;
    if (chunks > want) chunks = want;
    if (chunks < 1) chunks = 1;
    if (chunks > 1 && cudaMemsetAsync(scratch12, 0, sizeof(float) * 12 * n_views, stream) != cudaSuccess) return JT_ERR_LAUNCH;
    dim3 grid(chunks, n_views);
    g_launches += 2;
    pose_rays_bwd_reduce_kernel<<<grid, PR_THREADS, 0, stream>>>(A, d_center, d_ray, scratch12);
    pose_rays_bwd_pose_kernel<<<(n_views + 63) / 64, 64, 0, stream>>>(A, scratch12, d_se3, d_pose);
    JT_RETURN_LAUNCH();
}
