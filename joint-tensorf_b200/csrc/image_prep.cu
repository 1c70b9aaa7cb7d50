// 2-D supervision pre-processing and the render loss.
//
// Replaces, in the reference:
//   * Model.process_GT_images (model/nerf.py:57-113): separable 201-tap blur of every training image
//     (replicate pad + depth-wise conv1d along W, permute, the same along H), every 500 iterations and once
//     per blur scale of c2f_alternate_2D_scale_pool;
//   * Model.get_edge_mask (model/nerf.py:116-149): replicate pad + two 3x3 Sobel conv2d summed over the
//     3 colour channels, gradient magnitude, then soft (GG / max) or hard (GG > mean * thresh) masks;
//   * the render term of Graph.compute_loss (model/tensorf.py:99-124): gather of the supervising pixels
//     (image[:, ray_idx]) and edge masks, then plain / soft-edge / hard-edge MSE (nanmean, base.py:259-261).
//
// Image blur: a pass blurs along one axis. A CTA stages TL outputs + halo of 32 "lanes" (the other image axis)
// in shared memory as [position][33] (the W pass transposes while loading, so both passes run the same inner
// loop and every LDS is conflict-free), replicate padding is materialised by clamped loads. A thread owns
// 8 consecutive outputs of one lane: the input window slides through two register halves, per block of 8 taps
// 8 LDS and 64 FMAs with the tap operand read from the constant bank (taps travel in the kernel parameters).
// 201 taps = 402 flop per 8 B of HBM traffic: the pass is FMA-bound, hence the register tiling.
#include "jt_common.cuh"
#include "../../include/jt_vm.h"

namespace jt {

constexpr int IB_MAX_TAPS = 257;
constexpr int IB_KPAD = 264;       // taps zero-padded to a multiple of 8
constexpr int IB_THREADS = 256;
constexpr int IB_TL = 128;         // outputs per tile along the blurred axis (8 warps x 2 groups x 8)
constexpr int IB_LD = 33;          // padded lane stride in shared memory

struct ImgBlurArgs {
    const float* in;
    float* out;
    int n_img, H, W;
    int axis;          // 1: blur along x (lanes = rows), 0: blur along y (lanes = columns)
    int h;             // ntaps / 2
    int nblk;          // tap blocks of 8
    float k[IB_KPAD];
};

__global__ void __launch_bounds__(IB_THREADS) image_blur_pass_kernel(const __grid_constant__ ImgBlurArgs A) {
    extern __shared__ float tile[];                 // [(IB_TL + 8*nblk)][IB_LD]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = A.axis ? A.W : A.H;               // length of the blurred axis
    const int m = A.axis ? A.H : A.W;               // length of the lane axis
    const int i0 = blockIdx.x * IB_TL;              // first output of the tile
    const int c0 = blockIdx.y * 32;                 // first lane coordinate
    const float* img = A.in + (size_t)blockIdx.z * A.H * A.W;
    float* out = A.out + (size_t)blockIdx.z * A.H * A.W;
    const int npos = IB_TL + 8 * A.nblk;

    if (A.axis) {      // lanes = rows y, positions = x: coalesced along x, transposed into the tile
        for (int c = warp; c < 32; c += IB_THREADS / 32) {
            const int y = min(c0 + c, m - 1);
            const float* row = img + (size_t)y * A.W;
            for (int a = lane; a < npos; a += 32) {
                const int x = min(max(i0 - A.h + a, 0), n - 1);
                tile[a * IB_LD + c] = __ldg(row + x);
            }
        }
    } else {           // lanes = columns x, positions = y
        const int x = min(c0 + lane, m - 1);
        for (int a = warp; a < npos; a += IB_THREADS / 32) {
            const int y = min(max(i0 - A.h + a, 0), n - 1);
            tile[a * IB_LD + lane] = __ldg(img + (size_t)y * A.W + x);
        }
    }
    __syncthreads();

    const bool lane_ok = c0 + lane < m;
#pragma unroll 1
    for (int g = warp; g < IB_TL / 8; g += IB_THREADS / 32) {
        const int a0 = g * 8;
        if (i0 + a0 >= n) break;
        float acc[8], lo[8], hi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[j] = 0.f; lo[j] = tile[(a0 + j) * IB_LD + lane]; }
#pragma unroll 1
        for (int b = 0; b < A.nblk; ++b) {
#pragma unroll
            for (int j = 0; j < 8; ++j) hi[j] = tile[(a0 + 8 * (b + 1) + j) * IB_LD + lane];
            const float* kk = A.k + 8 * b;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float kv = kk[t];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int q = j + t;
                    acc[j] = fmaf(kv, q < 8 ? lo[q] : hi[q - 8], acc[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) lo[j] = hi[j];
        }
        if (!lane_ok) continue;
        if (A.axis) {
            float* dst = out + (size_t)(c0 + lane) * A.W + i0 + a0;
            if (i0 + a0 + 8 <= n && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                reinterpret_cast<float4*>(dst)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                reinterpret_cast<float4*>(dst)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (i0 + a0 + j < n) dst[j] = acc[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (i0 + a0 + j < n) out[(size_t)(i0 + a0 + j) * A.W + c0 + lane] = acc[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------- Sobel
constexpr int SB_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < blockDim.x / 32 ? sh[threadIdx.x] : 0.f;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < blockDim.x / 32 ? sh[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
    }
    __syncthreads();
    return r;
}

// gg[b][y][x] = sqrt(Gx^2 + Gy^2), Gx / Gy = 3x3 Sobel correlations of the replicate-padded image summed over the
// 3 channels (nerf.py:124-139). One CTA = 256 consecutive pixels of one image; partial sum / max per CTA.
__global__ void __launch_bounds__(SB_THREADS) sobel_mag_kernel(const float* __restrict__ img, int H, int W,
                                                               float* __restrict__ gg, float* __restrict__ part_sum,
                                                               float* __restrict__ part_max) {
    __shared__ float sh[8];
    const int b = blockIdx.y;
    const int hw = H * W;
    const int p = blockIdx.x * SB_THREADS + threadIdx.x;
    float v = 0.f;
    if (p < hw) {
        const int y = p / W, x = p - y * W;
        const int ym = max(y - 1, 0), yp = min(y + 1, H - 1), xm = max(x - 1, 0), xp = min(x + 1, W - 1);
        float gx = 0.f, gy = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* I = img + ((size_t)b * 3 + c) * hw;
            const float a00 = __ldg(I + ym * W + xm), a01 = __ldg(I + ym * W + x), a02 = __ldg(I + ym * W + xp);
            const float a10 = __ldg(I + y * W + xm), a12 = __ldg(I + y * W + xp);
            const float a20 = __ldg(I + yp * W + xm), a21 = __ldg(I + yp * W + x), a22 = __ldg(I + yp * W + xp);
            gx += (a00 - a02) + 2.f * (a10 - a12) + (a20 - a22);
            gy += (a00 + 2.f * a01 + a02) - (a20 + 2.f * a21 + a22);
        }
        v = sqrtf(gx * gx + gy * gy);
        gg[(size_t)b * hw + p] = v;
    }
    const float s = block_sum(v, sh);
    const float mx = block_max(v, sh);
    if (threadIdx.x == 0) {
        part_sum[(size_t)b * gridDim.x + blockIdx.x] = s;
        part_max[(size_t)b * gridDim.x + blockIdx.x] = mx;
    }
}

// stats[b] = {max, mean}: fixed-order reduction of the per-CTA partials (deterministic), sum in double.
__global__ void __launch_bounds__(SB_THREADS) sobel_stats_kernel(const float* __restrict__ part_sum,
                                                                 const float* __restrict__ part_max, int nparts,
                                                                 int hw, float* __restrict__ stats) {
    __shared__ double shs[SB_THREADS];
    __shared__ float shm[SB_THREADS];
    const int b = blockIdx.x;
    double s = 0.0;
    float mx = 0.f;
    for (int i = threadIdx.x; i < nparts; i += SB_THREADS) {
        s += (double)part_sum[(size_t)b * nparts + i];
        mx = fmaxf(mx, part_max[(size_t)b * nparts + i]);
    }
    shs[threadIdx.x] = s; shm[threadIdx.x] = mx;
    __syncthreads();
    for (int o = SB_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            shs[threadIdx.x] += shs[threadIdx.x + o];
            shm[threadIdx.x] = fmaxf(shm[threadIdx.x], shm[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        stats[2 * b] = shm[0];
        stats[2 * b + 1] = (float)(shs[0] / (double)hw);
    }
}

__global__ void __launch_bounds__(SB_THREADS) edge_mask_kernel(const float* __restrict__ gg, const float* __restrict__ stats,
                                                               int hw, int soft, float thresh, float* __restrict__ mask_f,
                                                               uint8_t* __restrict__ mask_u8) {
    const int b = blockIdx.y;
    const int p = blockIdx.x * SB_THREADS + threadIdx.x;
    if (p >= hw) return;
    const float v = gg[(size_t)b * hw + p];
    if (soft) mask_f[(size_t)b * hw + p] = v / stats[2 * b];                        // GG / GG_max (nerf.py:141-143)
    else mask_u8[(size_t)b * hw + p] = v > stats[2 * b + 1] * thresh ? 1 : 0;       // GG > GG_mean*thresh (145-148)
}

// ---------------------------------------------------------------------------------------------- render loss
constexpr int RL_THREADS = 1024;

struct LossArgs {
    const float* rgb;        // [B][n][3]
    const float* images;     // [n_cache][3][HW]
    const void* mask;        // [n_cache][HW] float or uint8, or NULL
    const int* ray_idx;      // [n] or NULL (identity: n == HW)
    const int* view_idx;     // [B] or NULL (identity)
    int B, n, hw, mask_kind; // 0 none, 1 float, 2 uint8
    int mode;                // 0 plain, 1 soft-edge (one MSE with m*fe + fn), 2 hard-edge (fe*MSE(m) + fn*MSE(1-m))
    float fe, fn;
};

__device__ __forceinline__ void loss_elem(const LossArgs& A, long long e, float& diff, float& m) {
    const int ch = (int)(e % 3);
    const long long r = e / 3;
    const int j = (int)(r % A.n), b = (int)(r / A.n);
    const int v = A.view_idx ? A.view_idx[b] : b;
    const int pix = A.ray_idx ? A.ray_idx[j] : j;
    const float t = __ldg(A.images + ((size_t)v * 3 + ch) * A.hw + pix);
    diff = A.rgb[e] - t;
    m = 1.f;
    if (A.mask_kind == 1) m = static_cast<const float*>(A.mask)[(size_t)v * A.hw + pix];
    else if (A.mask_kind == 2) m = (float)static_cast<const uint8_t*>(A.mask)[(size_t)v * A.hw + pix];
}

// Single CTA, fixed summation order (deterministic loss). ws = {S_a, S_b, cnt_a, cnt_b} doubles; loss[0] = value.
// pred*m - image*m is evaluated as the reference does (two products, then the difference).
__global__ void __launch_bounds__(RL_THREADS) render_loss_fwd_kernel(const LossArgs A, double* __restrict__ ws,
                                                                     float* __restrict__ loss) {
    __shared__ double sh[4][RL_THREADS / 32];
    const long long total = (long long)A.B * A.n * 3;
    double sa = 0.0, sb = 0.0, ca = 0.0, cb = 0.0;
    for (long long e = threadIdx.x; e < total; e += RL_THREADS) {
        const int ch = (int)(e % 3);
        const long long r = e / 3;
        const int j = (int)(r % A.n), b = (int)(r / A.n);
        const int v = A.view_idx ? A.view_idx[b] : b;
        const int pix = A.ray_idx ? A.ray_idx[j] : j;
        const float t = __ldg(A.images + ((size_t)v * 3 + ch) * A.hw + pix);
        const float p = A.rgb[e];
        float m = 1.f;
        if (A.mask_kind == 1) m = static_cast<const float*>(A.mask)[(size_t)v * A.hw + pix];
        else if (A.mask_kind == 2) m = (float)static_cast<const uint8_t*>(A.mask)[(size_t)v * A.hw + pix];
        if (A.mode == 0) {
            const float d = p - t;
            if (d == d) { sa += (double)(d * d); ca += 1.0; }
        } else if (A.mode == 1) {
            const float w = m * A.fe + A.fn;
            const float d = __fsub_rn(__fmul_rn(p, w), __fmul_rn(t, w));
            if (d == d) { sa += (double)(d * d); ca += 1.0; }
        } else {
            const float d1 = __fsub_rn(__fmul_rn(p, m), __fmul_rn(t, m));
            const float w = 1.f - m;
            const float d2 = __fsub_rn(__fmul_rn(p, w), __fmul_rn(t, w));
            if (d1 == d1) { sa += (double)(d1 * d1); ca += 1.0; }
            if (d2 == d2) { sb += (double)(d2 * d2); cb += 1.0; }
        }
    }
    double vals[4] = {sa, sb, ca, cb};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double v = vals[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double v = sh[k][threadIdx.x];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            vals[k] = v;
        }
        if (threadIdx.x == 0) {
            ws[0] = vals[0]; ws[1] = vals[1]; ws[2] = vals[2]; ws[3] = vals[3];
            float l;
            if (A.mode == 2) l = A.fe * (float)(vals[0] / vals[2]) + A.fn * (float)(vals[1] / vals[3]);
            else l = (float)(vals[0] / vals[2]);
            loss[0] = l;
        }
    }
}

// d loss / d rgb, times the upstream gradient g[0] (device scalar, NULL = 1).
__global__ void __launch_bounds__(256) render_loss_bwd_kernel(const LossArgs A, const double* __restrict__ ws,
                                                              const float* __restrict__ g, float* __restrict__ d_rgb) {
    const long long total = (long long)A.B * A.n * 3;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    float diff, m;
    loss_elem(A, e, diff, m);
    const float up = g ? g[0] : 1.f;
    float r;
    if (A.mode == 0) r = 2.f * diff / (float)ws[2];
    else if (A.mode == 1) { const float w = m * A.fe + A.fn; r = 2.f * diff * w * w / (float)ws[2]; }
    else { const float w = 1.f - m;
           r = A.fe * 2.f * diff * m * m / (float)ws[2] + A.fn * 2.f * diff * w * w / (float)ws[3]; }
    d_rgb[e] = up * r;
}

}  // namespace jt

using namespace jt;

extern "C" int jt_image_blur(const float* in, float* out, float* tmp, int n_img, int H, int W, const float* h_taps,
                             int ntaps, cudaStream_t stream) {
    JT_CHECK_ARG(in && out && tmp && h_taps && n_img > 0 && H > 0 && W > 0);
    JT_CHECK_ARG(ntaps >= 1 && ntaps <= IB_MAX_TAPS && (ntaps & 1));
    static bool attr_set = false;
    const int nblk = (ntaps + 7) / 8;
    const size_t smem = (size_t)(IB_TL + 8 * nblk) * IB_LD * sizeof(float);
    if (!attr_set) {
        if (cudaFuncSetAttribute(image_blur_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((IB_TL + IB_KPAD) * IB_LD * sizeof(float))) != cudaSuccess)
            return JT_ERR_LAUNCH;
        attr_set = true;
    }
    ImgBlurArgs A;
    A.n_img = n_img; A.H = H; A.W = W; A.h = ntaps / 2; A.nblk = nblk;
    for (int i = 0; i < IB_KPAD; ++i) A.k[i] = i < ntaps ? h_taps[i] : 0.f;
    // pass 1 along W (nerf.py:102-103), pass 2 along H (105-106)
    for (int pass = 0; pass < 2; ++pass) {
        A.axis = pass == 0 ? 1 : 0;
        A.in = pass == 0 ? in : tmp;
        A.out = pass == 0 ? tmp : out;
        const int n = A.axis ? W : H, m = A.axis ? H : W;
        for (int z0 = 0; z0 < n_img; z0 += 65535) {
            ImgBlurArgs B = A;
            B.in = A.in + (size_t)z0 * H * W;
            B.out = A.out + (size_t)z0 * H * W;
            dim3 grid((n + IB_TL - 1) / IB_TL, (m + 31) / 32, min(n_img - z0, 65535));
            image_blur_pass_kernel<<<grid, IB_THREADS, smem, stream>>>(B);
            ++g_launches;
        }
    }
    JT_RETURN_LAUNCH();
}

extern "C" long long jt_edge_mask_ws_floats(int n_img, int H, int W) {
    const long long nparts = ((long long)H * W + SB_THREADS - 1) / SB_THREADS;
    return 2 * (long long)n_img * nparts;
}

extern "C" int jt_edge_mask(const float* images, int n_img, int H, int W, int soft, float thresh, float* gg,
                            float* ws, float* stats, float* mask_f, uint8_t* mask_u8, cudaStream_t stream) {
    JT_CHECK_ARG(images && gg && ws && stats && n_img > 0 && n_img <= 65535 && H > 0 && W > 0);
    JT_CHECK_ARG(soft ? mask_f != nullptr : mask_u8 != nullptr);
    const int hw = H * W;
    const int nparts = (hw + SB_THREADS - 1) / SB_THREADS;
    float* part_sum = ws;
    float* part_max = ws + (size_t)n_img * nparts;
    sobel_mag_kernel<<<dim3(nparts, n_img), SB_THREADS, 0, stream>>>(images, H, W, gg, part_sum, part_max);
    sobel_stats_kernel<<<n_img, SB_THREADS, 0, stream>>>(part_sum, part_max, nparts, hw, stats);
    edge_mask_kernel<<<dim3(nparts, n_img), SB_THREADS, 0, stream>>>(gg, stats, hw, soft, thresh, mask_f, mask_u8);
    g_launches += 3;
    JT_RETURN_LAUNCH();
}

static int fill_loss(LossArgs& A, const float* rgb, const float* images, const void* mask, int mask_kind,
                     const int* ray_idx, const int* view_idx, int n_views, int n_rays, int hw, int mode, float fe,
                     float fn) {
    JT_CHECK_ARG(rgb && images && n_views > 0 && n_rays > 0 && hw > 0);
    JT_CHECK_ARG(mode >= 0 && mode <= 2 && mask_kind >= 0 && mask_kind <= 2);
    JT_CHECK_ARG(mode == 0 || (mask && mask_kind != 0));
    JT_CHECK_ARG(ray_idx || n_rays == hw);
    A.rgb = rgb; A.images = images; A.mask = mode == 0 ? nullptr : mask; A.ray_idx = ray_idx; A.view_idx = view_idx;
    A.B = n_views; A.n = n_rays; A.hw = hw; A.mask_kind = mode == 0 ? 0 : mask_kind; A.mode = mode; A.fe = fe; A.fn = fn;
    return JT_OK;
}

extern "C" int jt_render_loss_fwd(const float* rgb, const float* images, const void* mask, int mask_kind,
                                  const int* ray_idx, const int* view_idx, int n_views, int n_rays, int hw, int mode,
                                  float edge_factor, float non_edge_factor, double* ws4, float* loss,
                                  cudaStream_t stream) {
    LossArgs A;
    int rc = fill_loss(A, rgb, images, mask, mask_kind, ray_idx, view_idx, n_views, n_rays, hw, mode, edge_factor,
                       non_edge_factor);
    if (rc) return rc;
    JT_CHECK_ARG(ws4 && loss);
    render_loss_fwd_kernel<<<1, RL_THREADS, 0, stream>>>(A, ws4, loss);
    ++g_launches;
    JT_RETURN_LAUNCH();
}

extern "C" int jt_render_loss_bwd(const float* rgb, const float* images, const void* mask, int mask_kind,
                                  const int* ray_idx, const int* view_idx, int n_views, int n_rays, int hw, int mode,
                                  float edge_factor, float non_edge_factor, const double* ws4, const float* g_loss,
                                  float* d_rgb, cudaStream_t stream) {
    LossArgs A;
    int rc = fill_loss(A, rgb, images, mask, mask_kind, ray_idx, view_idx, n_views, n_rays, hw, mode, edge_factor,
                       non_edge_factor);
    if (rc) return rc;
    JT_CHECK_ARG(ws4 && d_rgb);
    const long long total = (long long)n_views * n_rays * 3;
    render_loss_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(A, ws4, g_loss, d_rgb);
    ++g_launches;
    JT_RETURN_LAUNCH();
}
