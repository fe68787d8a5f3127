// Per-voxel ensemble majority vote (replaces get_class_votes / get_new_label, majority_vote.py:23-37).
//
// votes[0] is the constant 1 (the reference never counts background, majority_vote.py:28,32),
// votes[c] = #{m : map_m == c} for 1 <= c < K, labels >= K match nothing, first maximum wins.
//
// (M+1) bytes of traffic per voxel leave ~30 instructions per voxel at the HBM roofline, an order of
// magnitude less than a per-voxel histogram costs, so the vote is bit-sliced (M <= 8): a thread turns
// 32 voxels of every map into four bit planes (bitslice.cuh); per class the M indicator planes are
// summed by a carry-save adder tree into a 1..4-bit sliced count, compared against the sliced running
// maximum (strictly greater, classes ascending = first maximum wins; the maximum starts at the
// background's single vote) and the winner's label planes are updated - every step a LOP3 over 32
// voxels.  The M x K one-hot temporaries of the NumPy version (M*K*V bytes + K*V*8 bytes) never exist.
// Chunks holding a label >= 16, ensembles above 8 maps and unaligned buffers take the scalar kernel.
#include "common.cuh"
#include "bitslice.cuh"

namespace mss {

struct VoteParams {
    const uint8_t* maps[MSS_MAX_VOTE_MAPS];
    int n_maps;
    int n_classes;
    long long n_voxels;
    uint8_t* out;
};

constexpr int kVoteThreads = 256;
constexpr int kVoteChunk = 32;  // voxels per thread and iteration of the sliced kernel

// scalar vote of one voxel: labels of the maps are the bytes at `shift` of lab[m]
__device__ __forceinline__ unsigned vote_one(const unsigned* lab, int n_maps, int n_classes, int shift) {
    unsigned long long votes = 1ull;  // background: exactly one vote; sixteen 4-bit counters
    for (int m = 0; m < n_maps; ++m) {
        const unsigned l = (lab[m] >> shift) & 0xffu;
        if (l >= 1u && l < static_cast<unsigned>(n_classes)) votes += 1ull << (4 * l);
    }
    unsigned best = 0, best_n = 0;
    for (int c = 0; c < n_classes; ++c) {
        const unsigned n = static_cast<unsigned>(votes >> (4 * c)) & 0xfu;
        if (n > best_n) {
            best_n = n;
            best = c;
        }
    }
    return best;
}

template <int M>
__global__ void __launch_bounds__(kVoteThreads) vote_sliced_kernel(const __grid_constant__ VoteParams p) {
    const int K = p.n_classes;
    const long long nchunks = p.n_voxels / kVoteChunk;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (long long i = tid; i < nchunks; i += stride) {
        unsigned w[M][8];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const uint4 a = ld_stream_u4(p.maps[m] + i * kVoteChunk), b = ld_stream_u4(p.maps[m] + i * kVoteChunk + 16);
            w[m][0] = a.x, w[m][1] = a.y, w[m][2] = a.z, w[m][3] = a.w;
            w[m][4] = b.x, w[m][5] = b.y, w[m][6] = b.z, w[m][7] = b.w;
        }
        unsigned wide = 0u;
#pragma unroll
        for (int m = 0; m < M; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) wide |= w[m][j];
        if ((wide & 0xF0F0F0F0u) != 0u) {  // a label >= 16 in the chunk: voxel by voxel, re-read as bytes
#pragma unroll 1
            for (int v = 0; v < kVoteChunk; ++v) {
                unsigned lab[M];
#pragma unroll
                for (int m = 0; m < M; ++m) lab[m] = p.maps[m][i * kVoteChunk + v];
                p.out[i * kVoteChunk + v] = static_cast<uint8_t>(vote_one(lab, M, K, 0));
            }
            continue;
        }
        unsigned res[8];
        vote_chunk<M>(w, K, res);
        uint4* dst = reinterpret_cast<uint4*>(p.out + i * kVoteChunk);
        dst[0] = make_uint4(res[0], res[1], res[2], res[3]);
        dst[1] = make_uint4(res[4], res[5], res[6], res[7]);
    }
    // scalar tail
    for (long long v = nchunks * kVoteChunk + tid; v < p.n_voxels; v += stride) {
        unsigned lab[M];
#pragma unroll
        for (int m = 0; m < M; ++m) lab[m] = p.maps[m][v];
        p.out[v] = static_cast<uint8_t>(vote_one(lab, M, K, 0));
    }
}

// any M <= MSS_MAX_VOTE_MAPS, any alignment: one voxel per thread and step
__global__ void __launch_bounds__(kVoteThreads) vote_scalar_kernel(const __grid_constant__ VoteParams p) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < p.n_voxels; v += stride) {
        unsigned lab[MSS_MAX_VOTE_MAPS];
#pragma unroll
        for (int m = 0; m < MSS_MAX_VOTE_MAPS; ++m) lab[m] = m < p.n_maps ? p.maps[m][v] : 0u;
        p.out[v] = static_cast<uint8_t>(vote_one(lab, p.n_maps, p.n_classes, 0));
    }
}

template <int M>
static void launch_sliced(unsigned blocks, cudaStream_t s, const VoteParams& p) {
    vote_sliced_kernel<M><<<blocks, kVoteThreads, 0, s>>>(p);
}

}  // namespace mss

using namespace mss;

extern "C" int mss_majority_vote(const uint8_t* const* maps, int32_t n_maps, int32_t n_classes, int64_t n_voxels,
                                 uint8_t* voted_out, void* stream) {
    MSS_REQUIRE(maps != nullptr && voted_out != nullptr, MSS_E_ARG, "majority_vote: null argument");
    MSS_REQUIRE(n_maps >= 1 && n_maps <= MSS_MAX_VOTE_MAPS, MSS_E_UNSUPPORTED, "majority_vote: n_maps %d outside [1, %d]",
                n_maps, MSS_MAX_VOTE_MAPS);
    MSS_REQUIRE(n_classes >= 1 && n_classes <= MSS_MAX_VOTE_CLASSES, MSS_E_UNSUPPORTED,
                "majority_vote: n_classes %d outside [1, %d]", n_classes, MSS_MAX_VOTE_CLASSES);
    MSS_REQUIRE(n_voxels > 0, MSS_E_ARG, "majority_vote: n_voxels must be positive");
    VoteParams p;
    bool aligned = reinterpret_cast<uintptr_t>(voted_out) % 16 == 0;
    for (int m = 0; m < MSS_MAX_VOTE_MAPS; ++m) {
        p.maps[m] = m < n_maps ? maps[m] : nullptr;
        if (m < n_maps) {
            MSS_REQUIRE(maps[m] != nullptr, MSS_E_ARG, "majority_vote: map %d is null", m);
            if (reinterpret_cast<uintptr_t>(maps[m]) % 16 != 0) aligned = false;
        }
    }
    p.n_maps = n_maps;
    p.n_classes = n_classes;
    p.n_voxels = n_voxels;
    p.out = voted_out;
    cudaStream_t s = as_stream(stream);
    if (aligned && n_maps <= 8) {
        long long blocks = (n_voxels / kVoteChunk + kVoteThreads - 1) / kVoteThreads + 1;
        if (blocks > 148LL * 8) blocks = 148LL * 8;
        const unsigned b = static_cast<unsigned>(blocks);
        switch (n_maps) {
            case 1: launch_sliced<1>(b, s, p); break;
            case 2: launch_sliced<2>(b, s, p); break;
            case 3: launch_sliced<3>(b, s, p); break;
            case 4: launch_sliced<4>(b, s, p); break;
            case 5: launch_sliced<5>(b, s, p); break;
            case 6: launch_sliced<6>(b, s, p); break;
            case 7: launch_sliced<7>(b, s, p); break;
            default: launch_sliced<8>(b, s, p); break;
        }
    } else {
        long long blocks = (n_voxels + kVoteThreads - 1) / kVoteThreads;
        if (blocks > 148LL * 16) blocks = 148LL * 16;
        vote_scalar_kernel<<<static_cast<unsigned>(blocks), kVoteThreads, 0, s>>>(p);
    }
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}
