// Per-class Dice confusion counts (replaces the one-hot + fp32 reductions of MONAI DiceMetric /
// AsDiscrete as wired at engine/test.py:28-31,50-56).
//
// TP[c] = #(pred==c & label==c), P[c] = #(pred==c), Y[c] = #(label==c) as exact int64 - the reference
// sums fp32 one-hots, which stops being exact above 2^24 voxels per class; Dice is derived from the
// integers on the host in float64.  Each thread streams 16 voxels per step and counts into packed
// 8-bit fields held in 64-bit registers (K <= 16: two registers per quantity), spilling to a
// shared-memory histogram every 240 voxels and to global int64 atomics once per block.
#include "common.cuh"

namespace mss {

constexpr int kDiceThreads = 256;

// add one to byte field (c & 7) of lo (c < 8) or hi (c >= 8); c >= 16 adds nothing
__device__ __forceinline__ void bump(unsigned long long& lo, unsigned long long& hi, unsigned c) {
    const unsigned long long one = 1ull << (8 * (c & 7u));
    lo += (c < 8u) ? one : 0ull;
    hi += (c >= 8u && c < 16u) ? one : 0ull;
}

__device__ __forceinline__ void spill(unsigned long long& lo, unsigned long long& hi, unsigned* sh) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const unsigned a = static_cast<unsigned>(lo >> (8 * c)) & 0xffu;
        const unsigned b = static_cast<unsigned>(hi >> (8 * c)) & 0xffu;
        if (a) atomicAdd(&sh[c], a);
        if (b) atomicAdd(&sh[8 + c], b);
    }
    lo = 0ull;
    hi = 0ull;
}

template <typename LabelT>
__device__ __forceinline__ unsigned label_class(LabelT v);
template <>
__device__ __forceinline__ unsigned label_class<uint8_t>(uint8_t v) { return v; }
template <>
__device__ __forceinline__ unsigned label_class<float>(float v) {
    // integer-valued float labels (engine/test.py:40); anything else belongs to no class
    const int i = __float2int_rz(v);
    return (static_cast<float>(i) == v && i >= 0) ? static_cast<unsigned>(i) : 0xffu;
}

template <typename LabelT>
__global__ void __launch_bounds__(kDiceThreads) dice_kernel(const uint8_t* __restrict__ pred, const LabelT* __restrict__ label,
                                                            long long n, int K, long long* __restrict__ counts, int vec_ok) {
    __shared__ unsigned sh[3][16];
    if (threadIdx.x < 48) (&sh[0][0])[threadIdx.x] = 0u;
    __syncthreads();
    unsigned long long tp_lo = 0, tp_hi = 0, p_lo = 0, p_hi = 0, y_lo = 0, y_hi = 0;
    int pending = 0;
    const unsigned Ku = static_cast<unsigned>(K);
    auto one = [&](unsigned pc, unsigned yc) {
        pc = pc < Ku ? pc : 0xffu;
        yc = yc < Ku ? yc : 0xffu;
        bump(p_lo, p_hi, pc);
        bump(y_lo, y_hi, yc);
        bump(tp_lo, tp_hi, pc == yc ? pc : 0xffu);
    };
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long n16 = vec_ok ? n / 16 : 0;
    for (long long i = tid; i < n16; i += stride) {
        const uint4 pv = ld_stream_u4(pred + i * 16);
        const unsigned pw[4] = {pv.x, pv.y, pv.z, pv.w};
        if (sizeof(LabelT) == 1) {
            const uint4 yv = ld_stream_u4(reinterpret_cast<const uint8_t*>(label) + i * 16);
            const unsigned yw[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int w = 0; w < 4; ++w)
#pragma unroll
                for (int j = 0; j < 4; ++j) one((pw[w] >> (8 * j)) & 0xffu, (yw[w] >> (8 * j)) & 0xffu);
        } else {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const float4 yf = ld_stream_f4(reinterpret_cast<const float*>(label) + i * 16 + w * 4);
                one(pw[w] & 0xffu, label_class<float>(yf.x));
                one((pw[w] >> 8) & 0xffu, label_class<float>(yf.y));
                one((pw[w] >> 16) & 0xffu, label_class<float>(yf.z));
                one((pw[w] >> 24) & 0xffu, label_class<float>(yf.w));
            }
        }
        pending += 16;
        if (pending >= 240) {  // 8-bit fields hold 255
            spill(tp_lo, tp_hi, sh[0]);
            spill(p_lo, p_hi, sh[1]);
            spill(y_lo, y_hi, sh[2]);
            pending = 0;
        }
    }
    for (long long v = n16 * 16 + tid; v < n; v += stride) {
        one(pred[v], label_class<LabelT>(label[v]));
        if (++pending >= 240) {
            spill(tp_lo, tp_hi, sh[0]);
            spill(p_lo, p_hi, sh[1]);
            spill(y_lo, y_hi, sh[2]);
            pending = 0;
        }
    }
    spill(tp_lo, tp_hi, sh[0]);
    spill(p_lo, p_hi, sh[1]);
    spill(y_lo, y_hi, sh[2]);
    __syncthreads();
    if (threadIdx.x < 48) {
        const int q = threadIdx.x / 16, c = threadIdx.x % 16;
        const unsigned v = sh[q][c];
        if (c < K && v) atomicAdd(reinterpret_cast<unsigned long long*>(counts) + q * K + c, static_cast<unsigned long long>(v));
    }
}

}  // namespace mss

using namespace mss;

extern "C" int mss_dice_counts(const uint8_t* pred, const void* label, int32_t label_dtype, int64_t n_voxels,
                               int32_t n_classes, long long* counts, void* stream) {
    MSS_REQUIRE(pred != nullptr && label != nullptr && counts != nullptr, MSS_E_ARG, "dice_counts: null argument");
    MSS_REQUIRE(n_voxels > 0, MSS_E_ARG, "dice_counts: n_voxels must be positive");
    MSS_REQUIRE(n_classes >= 1 && n_classes <= 16, MSS_E_UNSUPPORTED, "dice_counts: n_classes %d outside [1, 16]", n_classes);
    MSS_REQUIRE(label_dtype == 0 || label_dtype == 1, MSS_E_ARG, "dice_counts: label_dtype must be 0 (uint8) or 1 (float32)");
    // a block's shared histogram is 32-bit: bound the voxels one block can see below 2^32
    long long blocks = (n_voxels / 16 + kDiceThreads - 1) / kDiceThreads + 1;
    if (blocks > 148LL * 8) blocks = 148LL * 8;
    MSS_REQUIRE(n_voxels / blocks < (1LL << 31), MSS_E_UNSUPPORTED, "dice_counts: volume too large for one call");
    const int vec_ok = reinterpret_cast<uintptr_t>(pred) % 16 == 0 && reinterpret_cast<uintptr_t>(label) % 16 == 0;
    cudaStream_t s = as_stream(stream);
    if (label_dtype == 0)
        dice_kernel<uint8_t><<<static_cast<unsigned>(blocks), kDiceThreads, 0, s>>>(
            pred, static_cast<const uint8_t*>(label), n_voxels, n_classes, counts, vec_ok);
    else
        dice_kernel<float><<<static_cast<unsigned>(blocks), kDiceThreads, 0, s>>>(pred, static_cast<const float*>(label),
                                                                                  n_voxels, n_classes, counts, vec_ok);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}
