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

// dst += src over a 5-D box whose innermost dimension is contiguous in both operands and whose 4 outer dimensions
// have arbitrary element strides: one launch for any face of a block partition, and `src` may be PEER memory
// (a neighbour's accumulator mapped over NVLink) - then the add is the transfer.
struct HaloNdParams {
    float* dst;
    const float* src;
    long long ds[4], ss[4];
    long long n[4];
    long long row_len;
};

template <bool VEC>
__global__ void __launch_bounds__(256) halo_add_nd_kernel(const __grid_constant__ HaloNdParams p) {
    constexpr int E = VEC ? 4 : 1;
    const long long per_row = p.row_len / E;
    const long long total = per_row * p.n[0] * p.n[1] * p.n[2] * p.n[3];
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long r = i / per_row;
        const long long c = (i - r * per_row) * E;
        const long long i3 = r % p.n[3];
        r /= p.n[3];
        const long long i2 = r % p.n[2];
        r /= p.n[2];
        const long long i1 = r % p.n[1];
        const long long i0 = r / p.n[1];
        float* d = p.dst + i0 * p.ds[0] + i1 * p.ds[1] + i2 * p.ds[2] + i3 * p.ds[3] + c;
        const float* s = p.src + i0 * p.ss[0] + i1 * p.ss[1] + i2 * p.ss[2] + i3 * p.ss[3] + c;
        if (VEC) {
            float4 a = *reinterpret_cast<const float4*>(d);
            const float4 b = ld_stream_f4(s);
            a.x = __fadd_rn(a.x, b.x);
            a.y = __fadd_rn(a.y, b.y);
            a.z = __fadd_rn(a.z, b.z);
            a.w = __fadd_rn(a.w, b.w);
            *reinterpret_cast<float4*>(d) = a;
        } else {
            *d = __fadd_rn(*d, *s);
        }
    }
}

}  // namespace mss

using namespace mss;

extern "C" int mss_halo_add_nd(float* dst, const int64_t dst_strides[4], const float* src, const int64_t src_strides[4],
                               const int64_t dims[4], int64_t row_len, void* stream) {
    MSS_REQUIRE(dst && src && dst_strides && src_strides && dims, MSS_E_ARG, "halo_add_nd: null argument");
    MSS_REQUIRE(row_len > 0, MSS_E_ARG, "halo_add_nd: row_len must be positive");
    HaloNdParams p;
    p.dst = dst;
    p.src = src;
    p.row_len = row_len;
    bool vec = row_len % 4 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0 && reinterpret_cast<uintptr_t>(src) % 16 == 0;
    long long rows = 1;
    for (int a = 0; a < 4; ++a) {
        MSS_REQUIRE(dims[a] > 0 && dst_strides[a] >= 0 && src_strides[a] >= 0, MSS_E_ARG,
                    "halo_add_nd: dims must be positive, strides non-negative");
        p.n[a] = dims[a];
        p.ds[a] = dst_strides[a];
        p.ss[a] = src_strides[a];
        if (dims[a] > 1 && (dst_strides[a] % 4 != 0 || src_strides[a] % 4 != 0)) vec = false;
        rows *= dims[a];
    }
    const long long work = rows * (vec ? row_len / 4 : row_len);
    long long blocks = (work + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    if (vec)
        halo_add_nd_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(p);
    else
        halo_add_nd_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(p);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}

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
