// bf16 gather-side copy of the VM factors ("bf16 factor storage", north star item 2).
//
// The reference keeps its factors (tensoRF.py:159-169) in fp32 and samples them with F.grid_sample. Here the fp32
// tensors stay the master parameters (optimizer state, checkpoints, gradients are untouched); when
// `B200_VMSplit.factor_storage == "bf16"` the gather / scatter kernels read their taps from a bf16 copy in the same
// channel-last layout, refreshed by this kernel whenever a factor changed (one launch for all 12 arrays: 69 MB read
// + 35 MB written at cfg2). Half the bytes per tap through L2 and the L1 data pipe, which is what binds those kernels.
#include <cuda_bf16.h>
#include "jt_common.cuh"
#include "../../include/jt_vm.h"

namespace jt {

constexpr int kMaxCast = 12;
struct CastArgs {
    const float4* src[kMaxCast];
    uint2* dst[kMaxCast];
    long long quads[kMaxCast];      // elements / 4
    int n;
};

__global__ void __launch_bounds__(256) cast_bf16_kernel(const CastArgs A) {
    const int a = blockIdx.y;
    const float4* __restrict__ s = A.src[a];
    uint2* __restrict__ d = A.dst[a];
    const long long n = A.quads[a];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldcs(s + i);
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        d[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
}

}  // namespace jt

using namespace jt;

extern "C" int jt_cast_bf16_multi(int n_arrays, const void* const* h_src, void* const* h_dst, const long long* h_count,
                                  cudaStream_t stream) {
    JT_CHECK_ARG(n_arrays >= 1 && n_arrays <= kMaxCast && h_src && h_dst && h_count);
    CastArgs A;
    A.n = n_arrays;
    long long most = 0;
    for (int i = 0; i < n_arrays; ++i) {
        JT_CHECK_ARG(h_src[i] && h_dst[i] && h_count[i] >= 0 && h_count[i] % 4 == 0);
        JT_CHECK_ARG((reinterpret_cast<uintptr_t>(h_src[i]) & 15) == 0 && (reinterpret_cast<uintptr_t>(h_dst[i]) & 7) == 0);
        A.src[i] = static_cast<const float4*>(h_src[i]);
        A.dst[i] = static_cast<uint2*>(h_dst[i]);
        A.quads[i] = h_count[i] / 4;
        most = A.quads[i] > most ? A.quads[i] : most;
    }
    if (most == 0) return JT_OK;
    long long want = (most + 255) / 256;
    const int gx = (int)(want < 4 * kNumSMs ? want : 4 * kNumSMs);
    g_launches += 1;
    cast_bf16_kernel<<<dim3(gx, n_arrays), 256, 0, stream>>>(A);
    JT_RETURN_LAUNCH();
}
