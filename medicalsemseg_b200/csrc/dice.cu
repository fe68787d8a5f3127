// Per-class Dice confusion counts (replaces the one-hot + fp32 reductions of MONAI DiceMetric /
// AsDiscrete as wired at engine/test.py:28-31,50-56).
//
// TP[c] = #(pred==c & label==c), P[c] = #(pred==c), Y[c] = #(label==c) as exact int64 - the reference
// sums fp32 one-hots, which stops being exact above 2^24 voxels per class; Dice is derived from the
// integers on the host in float64.
//
// The kernel has 2 bytes of traffic per voxel, i.e. ~11 instructions per voxel at the HBM roofline, so
// the labels are counted bit-sliced: a thread turns 32 voxels (2 x 16-byte loads per map) into four
// 32-bit bit planes (nibble pack + a 4x4 bit-matrix transpose: ~24 instructions), after which the
// indicator of class c over all 32 voxels is one LOP3 and its count one POPC.  Per-thread counts live in
// registers as 16-bit pairs; a warp reduces them with REDUX and adds them to a shared histogram, one
// int64 atomic per class and block reaches global memory.  Labels >= 16 (legal: they belong to no class)
// send their 32-voxel chunk down a scalar path.
#include "common.cuh"
#include "bitslice.cuh"

namespace mss {

constexpr int kDiceThreads = 256;
constexpr int kDiceChunk = 32;           // voxels per thread and iteration
constexpr int kDiceFlushIters = 63;      // 16-bit halves survive the packed warp sum: 63 * 32 voxels * 32 lanes < 65536

template <typename LabelT>
__device__ __forceinline__ unsigned label_class(LabelT v);
template <>
__device__ __forceinline__ unsigned label_class<uint8_t>(uint8_t v) { return v; }
template <>
__device__ __forceinline__ unsigned label_class<float>(float v) {
    // integer-valued float labels (engine/test.py:40); anything else belongs to no class
    const int i = __float2int_rz(v);
    return (static_cast<float>(i) == v && i >= 0 && i < 255) ? static_cast<unsigned>(i) : 0xffu;
}

__device__ __forceinline__ unsigned pack4(float4 f) {
    return label_class<float>(f.x) | (label_class<float>(f.y) << 8) | (label_class<float>(f.z) << 16) |
           (label_class<float>(f.w) << 24);
}

// raw 16-byte loads of one 32-voxel chunk (kept as loaded so the next chunk can be in flight during the math)
template <typename LabelT>
struct ChunkRegs;
template <>
struct ChunkRegs<uint8_t> {
    uint4 p0, p1, y0, y1;
    __device__ __forceinline__ void load(const uint8_t* pred, const uint8_t* label, long long i) {
        p0 = ld_stream_u4(pred + i * kDiceChunk), p1 = ld_stream_u4(pred + i * kDiceChunk + 16);
        y0 = ld_stream_u4(label + i * kDiceChunk), y1 = ld_stream_u4(label + i * kDiceChunk + 16);
    }
    __device__ __forceinline__ void label_words(unsigned (&yw)[8]) const {
        yw[0] = y0.x, yw[1] = y0.y, yw[2] = y0.z, yw[3] = y0.w, yw[4] = y1.x, yw[5] = y1.y, yw[6] = y1.z, yw[7] = y1.w;
    }
};
template <>
struct ChunkRegs<float> {
    uint4 p0, p1;
    float4 f[8];
    __device__ __forceinline__ void load(const uint8_t* pred, const float* label, long long i) {
        p0 = ld_stream_u4(pred + i * kDiceChunk), p1 = ld_stream_u4(pred + i * kDiceChunk + 16);
#pragma unroll
        for (int w = 0; w < 8; ++w) f[w] = ld_stream_f4(label + i * kDiceChunk + 4 * w);
    }
    __device__ __forceinline__ void label_words(unsigned (&yw)[8]) const {
#pragma unroll
        for (int w = 0; w < 8; ++w) yw[w] = pack4(f[w]);
    }
};

// KP = ceil(K / 2) class pairs
template <typename LabelT, int KP>
__global__ void __launch_bounds__(kDiceThreads) dice_kernel(const uint8_t* __restrict__ pred_all,
                                                            const LabelT* __restrict__ label_all, long long n, int K,
                                                            long long* __restrict__ counts_all, int vec_ok) {
    // blockIdx.y = volume: a batch of label maps is counted by one launch, each into its own [3][K] slot
    const uint8_t* __restrict__ pred = pred_all + static_cast<long long>(blockIdx.y) * n;
    const LabelT* __restrict__ label = label_all + static_cast<long long>(blockIdx.y) * n;
    long long* __restrict__ counts = counts_all + static_cast<long long>(blockIdx.y) * 3 * K;
    __shared__ unsigned sh[3][16];  // TP, P, Y
    if (threadIdx.x < 48) (&sh[0][0])[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned Ku = static_cast<unsigned>(K);
    auto slow_one = [&](unsigned pc, unsigned yc) {
        if (pc < Ku) atomicAdd(&sh[1][pc], 1u);
        if (yc < Ku) atomicAdd(&sh[2][yc], 1u);
        if (pc == yc && pc < Ku) atomicAdd(&sh[0][pc], 1u);
    };

    unsigned tp[8], pp[8], yy[8];  // classes (2i, 2i+1) as 16-bit halves
#pragma unroll
    for (int i = 0; i < 8; ++i) tp[i] = pp[i] = yy[i] = 0u;
    auto flush = [&]() {  // warp sums of the packed counters (REDUX), one shared atomic per class and warp
#pragma unroll
        for (int i = 0; i < KP; ++i) {
            const unsigned a = __reduce_add_sync(0xffffffffu, tp[i]);
            const unsigned b = __reduce_add_sync(0xffffffffu, pp[i]);
            const unsigned c = __reduce_add_sync(0xffffffffu, yy[i]);
            if ((threadIdx.x & 31) == 0) {
                if (a & 0xffffu) atomicAdd(&sh[0][2 * i], a & 0xffffu);
                if (a >> 16) atomicAdd(&sh[0][2 * i + 1], a >> 16);
                if (b & 0xffffu) atomicAdd(&sh[1][2 * i], b & 0xffffu);
                if (b >> 16) atomicAdd(&sh[1][2 * i + 1], b >> 16);
                if (c & 0xffffu) atomicAdd(&sh[2][2 * i], c & 0xffffu);
                if (c >> 16) atomicAdd(&sh[2][2 * i + 1], c >> 16);
            }
            tp[i] = pp[i] = yy[i] = 0u;
        }
    };

    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long nchunks = vec_ok ? n / kDiceChunk : 0;
    // every lane of a warp runs the same number of iterations (REDUX needs the full warp): the loop bound is
    // per warp, lanes past the end work on an empty chunk
    const long long warp_first = tid - (threadIdx.x & 31);
    int iters = 0;
    ChunkRegs<LabelT> cur, nxt;
    if (tid < nchunks) cur.load(pred, label, tid);
    for (long long i = tid; warp_first + (i - tid) < nchunks; i += stride) {
        const bool have = i < nchunks;
        const long long inext = i + stride;
        if (inext < nchunks) nxt.load(pred, label, inext);  // in flight while this chunk is counted
        if (have) {
            unsigned pw[8], yw[8];
            pw[0] = cur.p0.x, pw[1] = cur.p0.y, pw[2] = cur.p0.z, pw[3] = cur.p0.w;
            pw[4] = cur.p1.x, pw[5] = cur.p1.y, pw[6] = cur.p1.z, pw[7] = cur.p1.w;
            cur.label_words(yw);
            if (has_wide_label(pw) || has_wide_label(yw)) {  // some label >= 16: voxel by voxel, re-read as bytes
#pragma unroll 1
                for (int v = 0; v < kDiceChunk; ++v)
                    slow_one(pred[i * kDiceChunk + v], label_class<LabelT>(label[i * kDiceChunk + v]));
            } else {
                dice_chunk<KP>(pw, yw, tp, pp, yy);
            }
        }
        cur = nxt;
        if (++iters == kDiceFlushIters) {
            flush();
            iters = 0;
        }
    }
    flush();
    // scalar tail (and the whole array when a pointer is not 16-byte aligned)
    for (long long v = nchunks * kDiceChunk + tid; v < n; v += stride) slow_one(pred[v], label_class<LabelT>(label[v]));
    __syncthreads();
    if (threadIdx.x < 48) {
        const int q = threadIdx.x / 16, c = threadIdx.x % 16;
        const unsigned v = sh[q][c];
        if (c < K && v) atomicAdd(reinterpret_cast<unsigned long long*>(counts) + q * K + c, static_cast<unsigned long long>(v));
    }
}

template <typename LabelT, int KP>
static void launch_dice_kp(dim3 blocks, cudaStream_t s, const uint8_t* pred, const LabelT* label, long long n, int K,
                           long long* counts, int vec_ok) {
    if (blocks.y == 1) {  // one volume: ONE wave of CTAs (a second, partial wave of a 30 us launch is mostly ramp and tail)
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dice_kernel<LabelT, KP>, kDiceThreads, 0) == cudaSuccess && occ > 0 &&
            blocks.x > 148u * occ)
            blocks.x = 148u * occ;
    }
    dice_kernel<LabelT, KP><<<blocks, kDiceThreads, 0, s>>>(pred, label, n, K, counts, vec_ok);
}

template <typename LabelT>
static void launch_dice(dim3 blocks, cudaStream_t s, const uint8_t* pred, const LabelT* label, long long n, int K,
                        long long* counts, int vec_ok) {
    switch ((K + 1) / 2) {
        case 1: launch_dice_kp<LabelT, 1>(blocks, s, pred, label, n, K, counts, vec_ok); break;
        case 2: launch_dice_kp<LabelT, 2>(blocks, s, pred, label, n, K, counts, vec_ok); break;
        case 3: launch_dice_kp<LabelT, 3>(blocks, s, pred, label, n, K, counts, vec_ok); break;
        case 4: launch_dice_kp<LabelT, 4>(blocks, s, pred, label, n, K, counts, vec_ok); break;
        case 5: launch_dice_kp<LabelT, 5>(blocks, s, pred, label, n, K, counts, vec_ok); break;
        case 6: launch_dice_kp<LabelT, 6>(blocks, s, pred, label, n, K, counts, vec_ok); break;
        case 7: launch_dice_kp<LabelT, 7>(blocks, s, pred, label, n, K, counts, vec_ok); break;
        default: launch_dice_kp<LabelT, 8>(blocks, s, pred, label, n, K, counts, vec_ok); break;
    }
}

}  // namespace mss

using namespace mss;

static int dice_counts_impl(const uint8_t* pred, const void* label, int32_t label_dtype, int64_t n_voxels, int64_t n_volumes,
                            int32_t n_classes, long long* counts, void* stream) {
    MSS_REQUIRE(pred != nullptr && label != nullptr && counts != nullptr, MSS_E_ARG, "dice_counts: null argument");
    MSS_REQUIRE(n_voxels > 0 && n_volumes > 0 && n_volumes <= 65535, MSS_E_ARG,
                "dice_counts: n_voxels must be positive, n_volumes in [1, 65535]");
    MSS_REQUIRE(n_classes >= 1 && n_classes <= 16, MSS_E_UNSUPPORTED, "dice_counts: n_classes %d outside [1, 16]", n_classes);
    MSS_REQUIRE(label_dtype == 0 || label_dtype == 1, MSS_E_ARG, "dice_counts: label_dtype must be 0 (uint8) or 1 (float32)");
    // a block's shared histogram is 32-bit: bound the voxels one block can see below 2^32
    long long blocks = (n_voxels / kDiceChunk + kDiceThreads - 1) / kDiceThreads + 1;
    if (blocks > 148LL * 4) blocks = 148LL * 4;
    MSS_REQUIRE(n_voxels / blocks < (1LL << 31), MSS_E_UNSUPPORTED, "dice_counts: volume too large for one call");
    // every volume of a batch must keep the 16-byte alignment of the first one for the vector path
    const int esz = label_dtype == 0 ? 1 : 4;
    const int vec_ok = reinterpret_cast<uintptr_t>(pred) % 16 == 0 && reinterpret_cast<uintptr_t>(label) % 16 == 0 &&
                       (n_volumes == 1 || (n_voxels % 16 == 0 && (n_voxels * esz) % 16 == 0));
    cudaStream_t s = as_stream(stream);
    const dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>(n_volumes));
    if (label_dtype == 0)
        launch_dice<uint8_t>(grid, s, pred, static_cast<const uint8_t*>(label), n_voxels, n_classes, counts, vec_ok);
    else
        launch_dice<float>(grid, s, pred, static_cast<const float*>(label), n_voxels, n_classes, counts, vec_ok);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}

extern "C" int mss_dice_counts(const uint8_t* pred, const void* label, int32_t label_dtype, int64_t n_voxels,
                               int32_t n_classes, long long* counts, void* stream) {
    return dice_counts_impl(pred, label, label_dtype, n_voxels, 1, n_classes, counts, stream);
}

extern "C" int mss_dice_counts_batched(const uint8_t* pred, const void* label, int32_t label_dtype, int64_t n_voxels,
                                       int64_t n_volumes, int32_t n_classes, long long* counts, void* stream) {
    return dice_counts_impl(pred, label, label_dtype, n_voxels, n_volumes, n_classes, counts, stream);
}
