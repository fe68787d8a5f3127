// Per-voxel ensemble majority vote (replaces get_class_votes / get_new_label, majority_vote.py:23-37).
//
// votes[0] is the constant 1 (the reference never counts background, majority_vote.py:28,32),
// votes[c] = #{m : map_m == c} for 1 <= c < K, labels >= K match nothing, first maximum wins.
// Each thread owns 16 consecutive voxels (one 16-byte load per map); the per-voxel votes live in one
// 64-bit register as sixteen 4-bit counters, so the M x K one-hot temporaries of the NumPy version
// (M*K*V bytes + K*V*8 bytes) never exist.
#include "common.cuh"

namespace mss {

struct VoteParams {
    const uint8_t* maps[MSS_MAX_VOTE_MAPS];
    int n_maps;
    int n_classes;
    long long n_voxels;
    uint8_t* out;
};

__device__ __forceinline__ unsigned vote_one(const unsigned (&lab)[MSS_MAX_VOTE_MAPS], int n_maps, int n_classes, int shift) {
    unsigned long long votes = 1ull;  // background: exactly one vote
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

template <bool VEC>
__global__ void __launch_bounds__(256) vote_kernel(const __grid_constant__ VoteParams p) {
    const long long n16 = VEC ? p.n_voxels / 16 : 0;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (long long i = tid; i < n16; i += stride) {
        uint4 in[MSS_MAX_VOTE_MAPS];
#pragma unroll
        for (int m = 0; m < MSS_MAX_VOTE_MAPS; ++m)
            if (m < p.n_maps) in[m] = ld_stream_u4(p.maps[m] + i * 16);
        uint4 res;
        unsigned* rw = reinterpret_cast<unsigned*>(&res);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            unsigned lab[MSS_MAX_VOTE_MAPS];
#pragma unroll
            for (int m = 0; m < MSS_MAX_VOTE_MAPS; ++m)
                lab[m] = m < p.n_maps ? reinterpret_cast<const unsigned*>(&in[m])[w] : 0u;
            unsigned r = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) r |= vote_one(lab, p.n_maps, p.n_classes, 8 * j) << (8 * j);
            rw[w] = r;
        }
        *reinterpret_cast<uint4*>(p.out + i * 16) = res;
    }
    // scalar tail (and the whole array when a pointer is not 16-byte aligned)
    for (long long v = n16 * 16 + tid; v < p.n_voxels; v += stride) {
        unsigned lab[MSS_MAX_VOTE_MAPS];
#pragma unroll
        for (int m = 0; m < MSS_MAX_VOTE_MAPS; ++m) lab[m] = m < p.n_maps ? p.maps[m][v] : 0u;
        p.out[v] = static_cast<uint8_t>(vote_one(lab, p.n_maps, p.n_classes, 0));
    }
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
    long long blocks = (n_voxels / 16 + 255) / 256 + 1;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    if (aligned)
        vote_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(p);
    else
        vote_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(p);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}
