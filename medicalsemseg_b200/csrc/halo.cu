// Halo reduction for z-slab partitioning across GPUs: the receiving rank adds its neighbour's partial
// weighted sums for the planes both ranks' windows cover (no counterpart in the single-GPU reference;
// it is the multi-GPU form of the `+=` at engine/utils.py:147).
#include "common.cuh"

namespace mss {

template <bool VEC>
__global__ void __launch_bounds__(256) halo_add_kernel(float* __restrict__ dst, long long dst_pitch,
                                                       const float* __restrict__ src, long long src_pitch, long long n_rows,
                                                       long long row_len) {
    constexpr int E = VEC ? 4 : 1;
    const long long per_row = row_len / E;
    const long long total = per_row * n_rows;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / per_row, c = (i - r * per_row) * E;
        if (VEC) {
            float4 a = *reinterpret_cast<const float4*>(dst + r * dst_pitch + c);
            const float4 b = ld_stream_f4(src + r * src_pitch + c);
            a.x = __fadd_rn(a.x, b.x);
            a.y = __fadd_rn(a.y, b.y);
            a.z = __fadd_rn(a.z, b.z);
            a.w = __fadd_rn(a.w, b.w);
            *reinterpret_cast<float4*>(dst + r * dst_pitch + c) = a;
        } else {
            dst[r * dst_pitch + c] = __fadd_rn(dst[r * dst_pitch + c], src[r * src_pitch + c]);
        }
    }
}

}  // namespace mss

using namespace mss;

extern "C" int mss_halo_add(float* dst, int64_t dst_pitch, const float* src, int64_t src_pitch, int64_t n_rows,
                            int64_t row_len, void* stream) {
    MSS_REQUIRE(dst != nullptr && src != nullptr, MSS_E_ARG, "halo_add: null argument");
    MSS_REQUIRE(n_rows > 0 && row_len > 0 && dst_pitch >= row_len && src_pitch >= row_len, MSS_E_ARG,
                "halo_add: need positive sizes and pitches >= row_len");
    const bool vec = row_len % 4 == 0 && dst_pitch % 4 == 0 && src_pitch % 4 == 0 &&
                     reinterpret_cast<uintptr_t>(dst) % 16 == 0 && reinterpret_cast<uintptr_t>(src) % 16 == 0;
    const long long work = (vec ? row_len / 4 : row_len) * n_rows;
    long long blocks = (work + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    if (vec)
        halo_add_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(dst, dst_pitch, src, src_pitch,
                                                                                            n_rows, row_len);
    else
        halo_add_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(dst, dst_pitch, src, src_pitch,
                                                                                             n_rows, row_len);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}
