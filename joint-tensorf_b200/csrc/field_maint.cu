// Field maintenance between steps:
//   compute_alpha / getDenseAlpha   batBase.py:27-42, tensorBase.py:618-633
//   updateAlphaMask                 tensorBase.py:635-661 (clamp, max_pool3d(5), threshold, new aabb)
//   up_sampling_VM                  tensoRF.py:274-287   (F.interpolate bilinear, align_corners=True)
// The reference evaluates the dense alpha grid slice by slice (gx iterations x ~25 ATen kernels, a
// [gy*gz,3] boolean-mask gather/scatter each) and materialises the full xyz grid twice; here one kernel
// generates the grid points, tests the previous occupancy mask, gathers the density factors and writes
// alpha in the [gz][gy][gx] order the pooling needs; two more do the separable 5^3 max-pool with the
// threshold, the bit-packing of the new mask and the bounding-box reduction.
#include "jt_common.cuh"
#include "vm_taps.cuh"
#include "../../include/jt_vm.h"

namespace jt {

struct AlphaArgs {
    const float* xyz;            // explicit points [n][3], or NULL: dense grid
    const float* lin[3];         // torch.linspace(0, 1, g) tables (device) for the dense grid
    int g[3];                    // gx, gy, gz
    long long n;
    float shift, length;
    int act, use_mask;
};

// 4 lanes per point (each owns channel quads sub, sub+4, ..), 8 points per warp; point e of the dense
// grid is (ix, iy, iz) = (e % gx, (e / gx) % gy, e / (gx * gy)): output order [gz][gy][gx].
__global__ void __launch_bounds__(256) field_alpha_kernel(Factors F, Geom G, MaskGeom M, AlphaArgs A,
                                                          float* __restrict__ alpha) {
    const int lane = threadIdx.x & 31, sub = lane & 3, grp = lane >> 2;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = warp * 8; base < A.n; base += nwarps * 8) {
        const long long e = base + grp;
        const bool act = e < A.n;
        float acc = 0.f;
        bool keep = false;
        if (act) {
            float p[3];
            if (A.xyz) {
                p[0] = A.xyz[3 * e]; p[1] = A.xyz[3 * e + 1]; p[2] = A.xyz[3 * e + 2];
            } else {
                const int ix = (int)(e % A.g[0]);
                const long long r = e / A.g[0];
                const int idx[3] = {ix, (int)(r % A.g[1]), (int)(r / A.g[1])};
#pragma unroll
                for (int a = 0; a < 3; ++a) {       // aabb[0]*(1-s) + aabb[1]*s, separately rounded (tensorBase.py:626)
                    const float s = __ldg(A.lin[a] + idx[a]);
                    p[a] = __fadd_rn(__fmul_rn(G.a0[a], __fsub_rn(1.0f, s)), __fmul_rn(G.a1[a], s));
                }
            }
            keep = A.use_mask ? mask_keep(M, p) : true;
            if (keep) {
                float u[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) u[a] = __fsub_rn(__fmul_rn(__fsub_rn(p[a], G.a0[a]), G.inv[a]), 1.0f);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const PlaneTaps t = plane_taps(F, i, u);
                    const int C = F.C[i];
                    for (int q = sub * 4; q < C; q += 16) {
                        float4 a = ldg4(t.p00 + q), b = ldg4(t.p10 + q), c = ldg4(t.p01 + q), d = ldg4(t.p11 + q);
                        float4 la = ldg4(t.l0 + q), lb = ldg4(t.l1 + q);
                        acc += f4_dot(f4_bilin(a, t.w00, b, t.w10, c, t.w01, d, t.w11), f4_lerp2(la, t.tl.w0, lb, t.tl.w1));
                    }
                }
            }
        }
        acc = quad_sum(acc);
        if (act && sub == 0) {
            float sigma = keep ? density_act(acc + A.shift, A.act) : 0.0f;
            alpha[e] = 1.0f - exp_neg(-sigma * A.length);
        }
    }
}

// max over the 5x5 (y, x) window of clamp(alpha, 0, 1); out-of-range neighbours do not take part
// (max_pool3d pads with -inf).
__global__ void __launch_bounds__(256) pool_xy_kernel(const float* __restrict__ in, float* __restrict__ out, int W, int H,
                                                      long long total) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int x = (int)(e % W);
    const long long r = e / W;
    const int y = (int)(r % H);
    const float* plane = in + (r / H) * (long long)H * W;
    float m = 0.0f;                       // clamp(.,0,1) >= 0 and the centre voxel is always in range
    for (int dy = -2; dy <= 2; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            m = fmaxf(m, fminf(fmaxf(__ldg(plane + (long long)yy * W + xx), 0.0f), 1.0f));
        }
    }
    out[e] = m;
}

// max over z (5), threshold, float volume + bit-packed volume + bounding box of the kept voxels.
// stats = {min_x, max_x, min_y, max_y, min_z, max_z, count}; initialised by the host entry.
__global__ void __launch_bounds__(256) pool_z_pack_kernel(const float* __restrict__ in, int W, int H, int D,
                                                          long long total, float thres, float* __restrict__ vol,
                                                          uint32_t* __restrict__ bits, int* __restrict__ stats) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // grid covers whole 32-voxel words
    const long long HW = (long long)H * W;
    bool on = false;
    int x = 0, y = 0, z = 0;
    if (e < total) {
        x = (int)(e % W);
        const long long r = e / W;
        y = (int)(r % H);
        z = (int)(r / H);
        float m = 0.0f;
#pragma unroll
        for (int dz = -2; dz <= 2; ++dz) {
            const int zz = z + dz;
            if (zz < 0 || zz >= D) continue;
            m = fmaxf(m, __ldg(in + e + dz * HW));
        }
        on = m >= thres;
        vol[e] = on ? 1.0f : 0.0f;
    }
    const unsigned word = __ballot_sync(0xffffffffu, on);
    const int lane = threadIdx.x & 31;
    if (lane == 0 && (e >> 5) < ((total + 31) >> 5)) bits[e >> 5] = word;
    if (word == 0) return;
    int mn[3] = {on ? x : INT_MAX, on ? y : INT_MAX, on ? z : INT_MAX};
    int mx[3] = {on ? x : -1, on ? y : -1, on ? z : -1};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        mn[a] = __reduce_min_sync(0xffffffffu, mn[a]);
        mx[a] = __reduce_max_sync(0xffffffffu, mx[a]);
    }
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(stats + 2 * a, mn[a]); atomicMax(stats + 2 * a + 1, mx[a]); }
        atomicAdd(stats + 6, __popc(word));
    }
}

__global__ void stats_init_kernel(int* stats) {
    if (threadIdx.x < 7) stats[threadIdx.x] = (threadIdx.x == 6) ? 0 : ((threadIdx.x & 1) ? -1 : INT_MAX);
}

// F.interpolate(mode="bilinear", align_corners=True) on a channel-last [H][W][C] array (ATen
// UpSampleBilinear2d: scale = (in-1)/(out-1), src = scale*dst, i1 = i0 + (i0 < in-1), lambda = src - i0).
__global__ void __launch_bounds__(256) resize_bilinear_cl_kernel(const float4* __restrict__ in, int H, int W, int Q,
                                                                 float4* __restrict__ out, int H2, int W2, float sh,
                                                                 float sw, long long total) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int q = (int)(e % Q);
    const long long r = e / Q;
    const int x = (int)(r % W2), y = (int)(r / W2);
    const float fy = __fmul_rn(sh, (float)y), fx = __fmul_rn(sw, (float)x);
    const int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly1 = fminf(fmaxf(fy - (float)y0, 0.0f), 1.0f), lx1 = fminf(fmaxf(fx - (float)x0, 0.0f), 1.0f);
    const float ly0 = 1.0f - ly1, lx0 = 1.0f - lx1;
    const float4 a = __ldg(in + ((long long)y0 * W + x0) * Q + q), b = __ldg(in + ((long long)y0 * W + x1) * Q + q);
    const float4 c = __ldg(in + ((long long)y1 * W + x0) * Q + q), d = __ldg(in + ((long long)y1 * W + x1) * Q + q);
    auto mix = [&](float va, float vb, float vc, float vd) {
        // h0lambda * (w0lambda * x00 + w1lambda * x01) + h1lambda * (w0lambda * x10 + w1lambda * x11), no contraction
        float top = __fadd_rn(__fmul_rn(lx0, va), __fmul_rn(lx1, vb));
        float bot = __fadd_rn(__fmul_rn(lx0, vc), __fmul_rn(lx1, vd));
        return __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
    };
    out[e] = make_float4(mix(a.x, b.x, c.x, d.x), mix(a.y, b.y, c.y, d.y), mix(a.z, b.z, c.z, d.z), mix(a.w, b.w, c.w, d.w));
}

}  // namespace jt

using namespace jt;

extern "C" int jt_field_alpha(const void* const* h_factors, const int* h_dims, const float* h_geom, const float* xyz,
                              long long n_points, const float* lin_x, const float* lin_y, const float* lin_z,
                              const int* h_grid3, const uint32_t* mask_bits, const int* h_mask_dims,
                              const float* h_mask_geom, float density_shift, int act, float length, float* alpha,
                              cudaStream_t stream) {
    JT_CHECK_ARG(h_factors && h_dims && h_geom && alpha);
    JT_CHECK_ARG(xyz || (lin_x && lin_y && lin_z && h_grid3));
    JT_CHECK_ARG(!mask_bits || (h_mask_dims && h_mask_geom));
    JT_CHECK_ARG(act == 0 || act == 1);
    Factors F;
    if (int rc = fill_factors(F, h_factors, h_dims)) return rc;
    if (F.bf16) return JT_ERR_UNSUPPORTED;          // maintenance reads the fp32 master factors
    AlphaArgs A{};
    A.xyz = xyz;
    if (xyz) {
        A.n = n_points;
    } else {
        JT_CHECK_ARG(h_grid3[0] > 0 && h_grid3[1] > 0 && h_grid3[2] > 0);
        A.lin[0] = lin_x; A.lin[1] = lin_y; A.lin[2] = lin_z;
        A.g[0] = h_grid3[0]; A.g[1] = h_grid3[1]; A.g[2] = h_grid3[2];
        A.n = (long long)h_grid3[0] * h_grid3[1] * h_grid3[2];
    }
    if (A.n <= 0) return JT_OK;
    A.shift = density_shift; A.length = length; A.act = act; A.use_mask = mask_bits != nullptr;
    Geom G = make_geom(h_geom);
    MaskGeom M = make_mask(mask_bits, h_mask_dims, h_mask_geom);
    long long blocks = (A.n + 63) / 64;                       // 64 points per 256-thread block
    int grid = (int)(blocks < (long long)kNumSMs * 16 ? blocks : (long long)kNumSMs * 16);
    field_alpha_kernel<<<grid, 256, 0, stream>>>(F, G, M, A, alpha);
    ++g_launches;
    JT_RETURN_LAUNCH();
}

extern "C" int jt_alpha_mask_build(const float* alpha, int W, int H, int D, float thres, float* tmp, float* vol,
                                   uint32_t* bits, int* stats7, cudaStream_t stream) {
    JT_CHECK_ARG(alpha && tmp && vol && bits && stats7 && W > 0 && H > 0 && D > 0);
    const long long total = (long long)W * H * D;
    const long long padded = (total + 31) / 32 * 32;
    stats_init_kernel<<<1, 32, 0, stream>>>(stats7);
    pool_xy_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(alpha, tmp, W, H, total);
    pool_z_pack_kernel<<<(unsigned)((padded + 255) / 256), 256, 0, stream>>>(tmp, W, H, D, total, thres, vol, bits, stats7);
    g_launches += 3;
    JT_RETURN_LAUNCH();
}

extern "C" int jt_resize_bilinear_cl(const float* in, int H, int W, int C, float* out, int H2, int W2,
                                     cudaStream_t stream) {
    JT_CHECK_ARG(in && out && H > 0 && W > 0 && H2 > 0 && W2 > 0 && C >= 4 && (C & 3) == 0);
    JT_CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
    const float sh = H2 > 1 ? (float)(H - 1) / (float)(H2 - 1) : 0.0f;
    const float sw = W2 > 1 ? (float)(W - 1) / (float)(W2 - 1) : 0.0f;
    const long long total = (long long)H2 * W2 * (C / 4);
    resize_bilinear_cl_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(in), H, W, C / 4, reinterpret_cast<float4*>(out), H2, W2, sh, sw, total);
    ++g_launches;
    JT_RETURN_LAUNCH();
}
