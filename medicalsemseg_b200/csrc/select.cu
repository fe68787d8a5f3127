// Order statistics on the device without a sort: the k-th smallest of (a masked subset of) a 32-bit array by a three-level
// radix histogram (11 + 11 + 10 bits), for up to two ranks at once - the two neighbours numpy.percentile interpolates.
// Used by
//   * the 95th-percentile Hausdorff distance (engine/test.py:31,55-57 -> MONAI compute_percent_hausdorff_distance ->
//     np.percentile over the squared distances at the surface voxels): keys = exact squared distances (int32), mask = the
//     other surface, ranks derived ON THE DEVICE from the number of surface voxels - no compaction, no sort, no host sync;
//   * monai.transforms.ScaleIntensityRangePercentiles as configured at data/dataset_builder.py:343-353
//     (np.percentile(img, 5 / 95)): keys = float32 voxels mapped to order-preserving unsigned integers.
// Every launch is stream-ordered; the result lands in device memory (n, key_lo, key_hi, key_max).
#include "common.cuh"

namespace mss {

constexpr int kSelBins = 2048;
constexpr int kSelThreads = 256;

struct SelState {            // device scratch (zeroed by the first kernel)
    unsigned long long n;    // valid elements
    unsigned long long rank[2];
    unsigned prefix[2];
    unsigned key_max;
    unsigned pad;
    unsigned hist[2][kSelBins];
};

struct SelParams {
    const unsigned* keys;
    const uint8_t* mask;  // nullable
    long long n_elems;
    int key_kind;         // 0: unsigned as stored, 1: float32 -> order-preserving unsigned
    int rank_mode;        // 0: absolute ranks, 1: numpy percentile ranks from the valid count
    double quant;
    long long k[2];
    SelState* st;
    unsigned long long* out;  // [4]: n, key_lo, key_hi, key_max (keys in the mapped unsigned domain)
};

__device__ __forceinline__ unsigned map_key(unsigned raw, int kind) {
    if (kind == 0) return raw;
    return (raw & 0x80000000u) ? ~raw : (raw | 0x80000000u);  // float32 bits -> unsigned with the same order
}

__global__ void sel_zero_kernel(SelState* st) {
    unsigned* w = reinterpret_cast<unsigned*>(st);
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(SelState) / 4); i += blockDim.x) w[i] = 0u;
}

// level 0: bits 31..21, level 1: bits 20..10, level 2: bits 9..0
__device__ __forceinline__ int sel_shift(int level) { return level == 0 ? 21 : (level == 1 ? 10 : 0); }
__device__ __forceinline__ int sel_bits(int level) { return level == 2 ? 10 : 11; }

__global__ void __launch_bounds__(kSelThreads) sel_hist_kernel(const __grid_constant__ SelParams p, int level) {
    __shared__ unsigned sh[2][kSelBins];
    for (int i = threadIdx.x; i < 2 * kSelBins; i += kSelThreads) (&sh[0][0])[i] = 0u;
    __syncthreads();
    const int shift = sel_shift(level), bits = sel_bits(level);
    const unsigned bmask = (1u << bits) - 1u;
    const unsigned pre0 = p.st->prefix[0], pre1 = p.st->prefix[1];
    const bool same = level == 0 || pre0 == pre1;  // both ranks still in the same bucket chain: one histogram serves both
    unsigned kmax = 0u;
    auto handle = [&](unsigned raw) {
        const unsigned key = map_key(raw, p.key_kind);
        const unsigned hi = level == 0 ? 0u : key >> (shift + bits);
        const unsigned b = (key >> shift) & bmask;
        if (level == 0) {
            atomicAdd(&sh[0][b], 1u);
            kmax = max(kmax, key);
        } else {
            if (hi == pre0) atomicAdd(&sh[0][b], 1u);
            if (!same && hi == pre1) atomicAdd(&sh[1][b], 1u);
        }
    };
    const long long tid = static_cast<long long>(blockIdx.x) * kSelThreads + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * kSelThreads;
    // 16 elements per step: one 16-byte load of the mask (most of a surface mask is zero: skipped without touching the keys)
    // or, unmasked, four 16-byte loads of keys
    const bool vec = reinterpret_cast<uintptr_t>(p.keys) % 16 == 0 && (p.mask == nullptr || reinterpret_cast<uintptr_t>(p.mask) % 16 == 0);
    const long long n16 = vec ? p.n_elems / 16 : 0;
    for (long long i = tid; i < n16; i += stride) {
        if (p.mask != nullptr) {
            const uint4 m = ld_stream_u4(p.mask + i * 16);
            const unsigned mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                if (mw[w] == 0u) continue;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if ((mw[w] >> (8 * e)) & 0xffu) handle(p.keys[i * 16 + 4 * w + e]);
            }
        } else {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const uint4 k4 = ld_stream_u4(p.keys + i * 16 + 4 * w);
                handle(k4.x), handle(k4.y), handle(k4.z), handle(k4.w);
            }
        }
    }
    for (long long i = n16 * 16 + tid; i < p.n_elems; i += stride) {
        if (p.mask != nullptr && p.mask[i] == 0) continue;
        handle(p.keys[i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSelBins; i += kSelThreads) {
        if (sh[0][i]) atomicAdd(&p.st->hist[0][i], sh[0][i]);
        if (sh[1][i]) atomicAdd(&p.st->hist[1][i], sh[1][i]);
    }
    if (level == 0) {
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        if ((threadIdx.x & 31) == 0 && kmax) atomicMax(&p.st->key_max, kmax);
    }
}

// one CTA: turn the histograms of this level into the next prefix / remaining rank of both targets
__global__ void __launch_bounds__(kSelThreads) sel_pick_kernel(const __grid_constant__ SelParams p, int level) {
    __shared__ unsigned long long s_cum[kSelThreads];
    __shared__ int s_bucket[2];
    __shared__ unsigned long long s_before[2];
    SelState* st = p.st;
    const int bits = sel_bits(level);
    const int nb = 1 << bits;
    const bool same = level == 0 || st->prefix[0] == st->prefix[1];
    if (level == 0) {
        // the valid count is the sum of the level-0 histogram
        unsigned long long part = 0;
        for (int i = threadIdx.x; i < nb; i += kSelThreads) part += st->hist[0][i];
        s_cum[threadIdx.x] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long n = 0;
            for (int i = 0; i < kSelThreads; ++i) n += s_cum[i];
            st->n = n;
            unsigned long long lo = 0, hi = 0;
            if (n > 0) {
                if (p.rank_mode == 1) {  // numpy.percentile, method='linear': virtual index (n - 1) * q / 100
                    const double pos = static_cast<double>(n - 1) * p.quant;
                    lo = pos <= 0.0 ? 0ull : static_cast<unsigned long long>(floor(pos));
                    if (lo > n - 1) lo = n - 1;
                    hi = lo + 1 > n - 1 ? n - 1 : lo + 1;
                } else {
                    lo = static_cast<unsigned long long>(p.k[0] < 0 ? 0 : p.k[0]);
                    hi = static_cast<unsigned long long>(p.k[1] < 0 ? 0 : p.k[1]);
                    if (lo > n - 1) lo = n - 1;
                    if (hi > n - 1) hi = n - 1;
                }
            }
            st->rank[0] = lo;
            st->rank[1] = hi;
        }
        __syncthreads();
    }
    if (st->n == 0) {
        if (threadIdx.x == 0 && level == 2) {
            p.out[0] = 0;
            p.out[1] = p.out[2] = p.out[3] = 0;
        }
        return;
    }
    // sequential scan by one thread per target (2048 bins: negligible next to the histogram pass)
    if (threadIdx.x < 2) {
        const int j = threadIdx.x;
        const unsigned* h = st->hist[(same || j == 0) ? 0 : 1];
        const unsigned long long r = st->rank[j];
        unsigned long long cum = 0;
        int b = 0;
        for (; b < nb; ++b) {
            const unsigned long long c = h[b];
            if (cum + c > r) break;
            cum += c;
        }
        if (b >= nb) b = nb - 1;  // (cannot happen: the ranks are < the population of the chain)
        s_bucket[j] = b;
        s_before[j] = cum;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int j = threadIdx.x;
        st->rank[j] -= s_before[j];
        st->prefix[j] = (level == 0 ? 0u : (st->prefix[j] << bits)) | static_cast<unsigned>(s_bucket[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * kSelBins; i += kSelThreads) (&st->hist[0][0])[i] = 0u;
    if (level == 2 && threadIdx.x == 0) {
        p.out[0] = st->n;
        p.out[1] = st->prefix[0];
        p.out[2] = st->prefix[1];
        p.out[3] = st->key_max;
    }
}

}  // namespace mss

using namespace mss;

extern "C" int64_t mss_select_scratch_bytes(void) { return static_cast<int64_t>(sizeof(SelState)); }

extern "C" int mss_select2(const void* keys, int32_t key_kind, const uint8_t* mask, int64_t n_elems, int32_t rank_mode,
                           double quant, int64_t k_lo, int64_t k_hi, void* scratch, unsigned long long* out, void* stream) {
    MSS_REQUIRE(keys != nullptr && scratch != nullptr && out != nullptr, MSS_E_ARG, "select2: null argument");
    MSS_REQUIRE(n_elems > 0, MSS_E_ARG, "select2: n_elems must be positive");
    MSS_REQUIRE(key_kind == 0 || key_kind == 1, MSS_E_ARG, "select2: key_kind %d (0 = uint32, 1 = float32)", key_kind);
    MSS_REQUIRE(rank_mode == 0 || rank_mode == 1, MSS_E_ARG, "select2: rank_mode %d (0 = absolute, 1 = percentile)", rank_mode);
    MSS_REQUIRE(reinterpret_cast<uintptr_t>(scratch) % 8 == 0 && reinterpret_cast<uintptr_t>(keys) % 4 == 0, MSS_E_ALIGN,
                "select2: scratch must be 8-byte aligned");
    SelParams p;
    p.keys = static_cast<const unsigned*>(keys);
    p.mask = mask;
    p.n_elems = n_elems;
    p.key_kind = key_kind;
    p.rank_mode = rank_mode;
    p.quant = quant;
    p.k[0] = k_lo;
    p.k[1] = k_hi;
    p.st = static_cast<SelState*>(scratch);
    p.out = out;
    cudaStream_t s = as_stream(stream);
    long long blocks = (n_elems + kSelThreads * 8 - 1) / (kSelThreads * 8);
    if (blocks > 148LL * 8) blocks = 148LL * 8;
    if (blocks < 1) blocks = 1;
    sel_zero_kernel<<<1, 256, 0, s>>>(p.st);
    for (int level = 0; level < 3; ++level) {
        sel_hist_kernel<<<static_cast<unsigned>(blocks), kSelThreads, 0, s>>>(p, level);
        sel_pick_kernel<<<1, kSelThreads, 0, s>>>(p, level);
    }
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}
