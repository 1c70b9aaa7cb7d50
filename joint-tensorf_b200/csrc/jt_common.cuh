// Shared device/host helpers for the joint-tensorf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define JT_OK 0
#define JT_ERR_ARG -1        // bad argument (null pointer, size constraint violated)
#define JT_ERR_LAUNCH -2     // cudaGetLastError() after a launch
#define JT_ERR_UNSUPPORTED -3

#define JT_CHECK_ARG(cond) do { if (!(cond)) return JT_ERR_ARG; } while (0)
#define JT_RETURN_LAUNCH() do { return cudaGetLastError() == cudaSuccess ? JT_OK : JT_ERR_LAUNCH; } while (0)

namespace jt {

extern long long g_launches;   // kernels launched through the C ABI (jt_launch_count)

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; persistent grids are multiples of this

// Field geometry, computed on the host with torch exactly as the reference does
// (tensorBase.py:477-488) and handed over as floats.
struct Geom {
    float a0[3], a1[3];   // aabb
    float inv[3];         // 2 / aabbSize
    float step;           // stepSize
    float near_, far_;
};

// Optional occupancy volume (AlphaGridMask, tensorBase.py:80-98), bit-packed:
// voxel n = (z*H + y)*W + x -> word n>>5, bit n&31.
struct MaskGeom {
    const uint32_t* bits;
    float a0[3], inv[3];  // mask aabb[0], (1/size)*2
    int W, H, D;
};

// 3 plane + 3 line factors in channel-last layout: plane i is [H_i][W_i][C_i],
// line i is [L_i][C_i]. x (fastest spatial axis, size W) is sampled with
// u[mat0(i)], y with u[mat1(i)], the line with u[vec(i)].
struct Factors {
    const float* plane[3];
    const float* line[3];
    int H[3], W[3], L[3], C[3];
    int off[3];           // channel offset of plane i in the concatenated feature vector
    int ctot;
    int bf16;             // element type the plane / line pointers address: 0 fp32, 1 bf16 (a gather-side copy of
                          // the fp32 master factors; offsets are in elements either way)
};
struct FactorGrads {
    float* plane[3];
    float* line[3];
};

__device__ __forceinline__ int mat0(int i) { return i == 2 ? 1 : 0; }   // matMode = [[0,1],[0,2],[1,2]]
__device__ __forceinline__ int mat1(int i) { return i == 0 ? 1 : 2; }
__device__ __forceinline__ int vecm(int i) { return 2 - i; }            // vecMode = [2,1,0]

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float quad_sum(float v) {       // over the 4 lanes sharing a sample
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// Vector reduction into global memory: one 16-byte RED per call (sm_90+).
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    // no "memory" clobber on purpose: the gradient buffers are never read by the issuing kernel, and a
    // clobber would pin every later tap load behind the reduction (it serialised the scatter kernels)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// four bf16 (8 bytes) -> float4: a bf16 is the upper half of the fp32 with the same value
__device__ __forceinline__ float4 bf16x4_to_f4(uint2 h) {
    return make_float4(__uint_as_float(h.x << 16), __uint_as_float(h.x & 0xFFFF0000u),
                       __uint_as_float(h.y << 16), __uint_as_float(h.y & 0xFFFF0000u));
}
// one channel quad of a factor tap: `base` addresses fp32 or (B16) bf16 elements, `off` is an element offset
template <bool B16>
__device__ __forceinline__ float4 ld_tap4(const float* base, size_t off) {
    if (B16) return bf16x4_to_f4(__ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(base) + off)));
    return __ldg(reinterpret_cast<const float4*>(base + off));
}

// 1-D linear interpolation setup on an axis of n texels, align_corners=True,
// zero padding (ATen grid_sampler semantics): index = (g+1)/2*(n-1).
struct Tap {
    int i0, i1;      // clamped indices (always safe to load)
    float w0, w1;    // weights, zeroed for out-of-range taps
    float m0, m1;    // 1 if the tap is in range else 0 (needed for d/dcoord)
    float scale;     // d index / d g = (n-1)/2
};
__device__ __forceinline__ Tap make_tap(float g, int n) {
    Tap t;
    t.scale = 0.5f * (float)(n - 1);
    float x = (g + 1.0f) * t.scale;
    float xf = floorf(x);
    float f = x - xf;
    float xc = fminf(fmaxf(xf, -2.0f), (float)n);    // NaN -> -2 (both taps out of range)
    int i0 = (int)xc;
    int i1 = i0 + 1;
    t.m0 = (i0 >= 0 && i0 < n) ? 1.0f : 0.0f;
    t.m1 = (i1 >= 0 && i1 < n) ? 1.0f : 0.0f;
    t.w0 = (1.0f - f) * t.m0;
    t.w1 = f * t.m1;
    t.i0 = min(max(i0, 0), n - 1);
    t.i1 = min(max(i1, 0), n - 1);
    return t;
}

// host: unpack the C-ABI factor description (include/jt_vm.h "Layouts"); dims[12] = element type (0 fp32, 1 bf16)
inline int fill_factors(Factors& F, const void* const* ptrs, const int* dims) {
    int off = 0;
    for (int i = 0; i < 3; ++i) {
        F.plane[i] = static_cast<const float*>(ptrs[i]);
        F.line[i] = static_cast<const float*>(ptrs[3 + i]);
        F.H[i] = dims[i]; F.W[i] = dims[3 + i]; F.L[i] = dims[6 + i]; F.C[i] = dims[9 + i];
        if (!F.plane[i] || !F.line[i]) return JT_ERR_ARG;
        if (F.C[i] <= 0 || (F.C[i] & 3) || F.H[i] < 1 || F.W[i] < 1 || F.L[i] < 1) return JT_ERR_ARG;
        F.off[i] = off;
        off += F.C[i];
    }
    F.ctot = off;
    F.bf16 = dims[12] ? 1 : 0;
    return JT_OK;
}

// degree-2 real spherical-harmonics basis of a direction (reference sh.py:88-113, eval_sh_bases(2, .))
__device__ __forceinline__ void sh9(const float d[3], float y[9]) {
    const float x = d[0], yy_ = d[1], z = d[2];
    y[0] = 0.28209479177387814f;
    y[1] = -0.4886025119029199f * yy_;
    y[2] = 0.4886025119029199f * z;
    y[3] = -0.4886025119029199f * x;
    const float xx = x * x, yy = yy_ * yy_, zz = z * z;
    y[4] = 1.0925484305920792f * (x * yy_);
    y[5] = -1.0925484305920792f * (yy_ * z);
    y[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    y[7] = -1.0925484305920792f * (x * z);
    y[8] = 0.5462742152960396f * (xx - yy);
}

}  // namespace jt
