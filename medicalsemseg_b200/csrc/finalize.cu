// Normalise + softmax + argmax -> uint8 label map
// (replaces `output_image / count_map` of engine/utils.py:151 and the softmax -> D2H -> np.argmax -> uint8
// of engine/test.py:140-141 / :81-82; the eval variant AsDiscrete(argmax=True) of engine/test.py:29).
//
// One thread owns 4 consecutive W voxels and streams the K class planes once (16-byte loads, KU planes
// in flight per thread, three 256-thread CTAs per SM).  Three specialisations share the body:
//   labels only      - the common case.  The first-max argmax is taken on the sums as stored: dividing
//                      every class of a voxel by its (positive) weight count cannot reorder them, so the
//                      count is not even computed; two classes can only collapse into a tie through the
//                      rounding of the division, which is far inside the near-tie tolerance (counted).
//   normalised logits - divides by the window-weight count (the K-channel-replicated count_map of the
//                      reference, engine/utils.py:142,148, is never materialised: the count of a voxel is
//                      re-derived from the geometry table and the L2-resident importance map as the
//                      ascending fp32 sum over its covering windows - same additions, same order).
//   probabilities    - adds the two softmax sweeps (engine/test.py:140); kept out of the hot variants.
#include "common.cuh"
#include "labels.cuh"

namespace mss {

constexpr int kFinThreads = 256;

struct FinParams {
    Geo g;
    const float* logits;
    const float* imp;
    int box_lo[3];
    int box_n[3];
    int nq;
    uint8_t* labels;
    int label_pitch;
    float* logits_out;
    float* probs_out;
    float tie_tol;
    unsigned long long* near_ties;
};

// ascending fp32 sum of the importance weights of all windows covering the 4 voxels at global (gd, gh, gw..gw+3)
__device__ __forceinline__ void weight_count(const Geo& g, const float* __restrict__ imp, int gd, int gh, int gw,
                                             const bool (&valid)[4], float (&cnt)[4]) {
    const int rh = g.roi[1], rw = g.roi[2];
    const int cvd = g.cover[0][gd], cvh = g.cover[1][gh];
    int wlo = 0x7fffffff, whi = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (valid[e]) {
            const int c = g.cover[2][gw + e];
            wlo = min(wlo, c & 0xffff);
            whi = max(whi, c >> 16);
        }
#pragma unroll
    for (int e = 0; e < 4; ++e) cnt[e] = 0.f;
    for (int id = cvd & 0xffff; id < (cvd >> 16); ++id)
        for (int ih = cvh & 0xffff; ih < (cvh >> 16); ++ih) {
            const float* row = imp + (static_cast<long long>(gd - g.starts[0][id]) * rh + (gh - g.starts[1][ih])) * rw;
            for (int iw = wlo; iw < whi; ++iw) {
                const int ww = gw - g.starts[2][iw];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (valid[e] && ww + e >= 0 && ww + e < rw) cnt[e] = __fadd_rn(cnt[e], __ldg(row + ww + e));
            }
        }
}

enum : int { kFinLabels = 0, kFinNormalise = 1, kFinProbs = 2 };

template <int MODE, int KU>
__global__ void __launch_bounds__(kFinThreads, MODE == kFinLabels ? 3 : 2) finalize_kernel(const __grid_constant__ FinParams p) {
    const Geo& g = p.g;
    const int t = blockIdx.x * kFinThreads + threadIdx.x;
    const bool in_box = t < p.nq * p.box_n[1];
    const int row = in_box ? t / p.nq : 0;
    const int ld = p.box_lo[0] + blockIdx.y;
    const int lh = p.box_lo[1] + row;
    const int lw = p.box_lo[2] + (in_box ? t - row * p.nq : 0) * 4;
    const int b = blockIdx.z;
    const int hi_w = min(p.box_lo[2] + p.box_n[2], g.ext[2]);
    bool valid[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) valid[e] = in_box && lw + e < hi_w;
    unsigned ties = 0;
    if (valid[0]) {
        const int K = g.K;
        const long long plane = static_cast<long long>(g.ext[1]) * g.pitch;
        const long long cstride = static_cast<long long>(g.ext[0]) * plane;
        const long long vox = static_cast<long long>(b) * K * cstride + static_cast<long long>(ld) * plane +
                              static_cast<long long>(lh) * g.pitch + lw;
        const float* src = p.logits + vox;
        const bool full = valid[3];

        float cnt[4] = {1.f, 1.f, 1.f, 1.f};
        const bool divide = MODE != kFinLabels && p.imp != nullptr;
        if (divide) weight_count(g, p.imp, ld + g.org[0], lh + g.org[1], lw + g.org[2], valid, cnt);

        ArgmaxState am[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) am[e].reset();

        auto consume = [&](float4 v, int k) {
            if (MODE != kFinLabels) {
                if (divide) {
                    v.x = __fdiv_rn(v.x, cnt[0]);  // engine/utils.py:151
                    v.y = __fdiv_rn(v.y, cnt[1]);
                    v.z = __fdiv_rn(v.z, cnt[2]);
                    v.w = __fdiv_rn(v.w, cnt[3]);
                }
                if (p.logits_out != nullptr) {
                    float* dst = p.logits_out + vox + k * cstride;
                    if (full) {
                        *reinterpret_cast<float4*>(dst) = v;
                    } else {
                        dst[0] = v.x;
                        if (valid[1]) dst[1] = v.y;
                        if (valid[2]) dst[2] = v.z;
                    }
                }
            }
            am[0].push(v.x, k);
            am[1].push(v.y, k);
            am[2].push(v.z, k);
            am[3].push(v.w, k);
        };

        int k0 = 0;
        for (; k0 + KU <= K; k0 += KU) {  // full groups: KU independent 16-byte loads in flight
            float4 v[KU];
#pragma unroll
            for (int k = 0; k < KU; ++k) v[k] = ld_stream_f4(src + (k0 + k) * cstride);
#pragma unroll
            for (int k = 0; k < KU; ++k) consume(v[k], k0 + k);
        }
        if (k0 < K) {
            float4 v[KU];
#pragma unroll
            for (int k = 0; k < KU; ++k)
                if (k0 + k < K) v[k] = ld_stream_f4(src + (k0 + k) * cstride);
#pragma unroll
            for (int k = 0; k < KU; ++k)
                if (k0 + k < K) consume(v[k], k0 + k);
        }

        if (MODE == kFinProbs && p.probs_out != nullptr) {
            // softmax(dim=classes) = exp(x - max) / sum exp(x - max): second and third sweep over the planes
            const float* nsrc = (p.logits_out != nullptr) ? p.logits_out + vox : src;
            const bool renorm = divide && p.logits_out == nullptr;
            const float m[4] = {am[0].best, am[1].best, am[2].best, am[3].best};
            float s[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k = 0; k < K; ++k) {
                const float4 x = *reinterpret_cast<const float4*>(nsrc + k * cstride);
                float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (renorm) xv[e] = __fdiv_rn(xv[e], cnt[e]);
                    s[e] += expf(xv[e] - m[e]);
                }
            }
            for (int k = 0; k < K; ++k) {
                const float4 x = *reinterpret_cast<const float4*>(nsrc + k * cstride);
                float xv[4] = {x.x, x.y, x.z, x.w};
                float* dst = p.probs_out + vox + k * cstride;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (renorm) xv[e] = __fdiv_rn(xv[e], cnt[e]);
                    if (valid[e]) dst[e] = expf(xv[e] - m[e]) / s[e];
                }
            }
        }

        if (p.labels != nullptr) {
            uint8_t* lab = p.labels + (static_cast<long long>(b) * g.ext[0] + ld) * g.ext[1] * p.label_pitch +
                           static_cast<long long>(lh) * p.label_pitch + lw;
            unsigned packed = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (valid[e]) {
                    packed |= static_cast<unsigned>(am[e].label()) << (8 * e);
                    ties += am[e].near_tie(p.tie_tol) ? 1u : 0u;
                }
            if (full && ((reinterpret_cast<uintptr_t>(lab) & 3u) == 0)) {
                *reinterpret_cast<unsigned*>(lab) = packed;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (valid[e]) lab[e] = static_cast<uint8_t>(packed >> (8 * e));
            }
        }
    }
    // near-tie census: one atomic per warp that saw any
    if (p.near_ties != nullptr) {
        const unsigned warp_ties = __reduce_add_sync(0xffffffffu, ties);
        if (warp_ties != 0u && (threadIdx.x & 31) == 0) atomicAdd(p.near_ties, static_cast<unsigned long long>(warp_ties));
    }
}

// Labels only, few classes (K <= 4, e.g. BraTS' 3): a voxel carries so few bytes that a thread takes QPT quads of
// the plane (a CTA-stride apart, so every load stays coalesced) and issues all QPT * K loads before the first compare.
template <int K, int QPT>
__global__ void __launch_bounds__(kFinThreads) finalize_labels_smallk_kernel(const __grid_constant__ FinParams p) {
    const Geo& g = p.g;
    const int ld = p.box_lo[0] + blockIdx.y;
    const int b = blockIdx.z;
    const int hi_w = min(p.box_lo[2] + p.box_n[2], g.ext[2]);
    const long long plane = static_cast<long long>(g.ext[1]) * g.pitch;
    const long long cstride = static_cast<long long>(g.ext[0]) * plane;
    const long long base = static_cast<long long>(b) * K * cstride + static_cast<long long>(ld) * plane;
    const int per_plane = p.nq * p.box_n[1];
    float4 v[QPT][K];
    int lh[QPT], lw[QPT];
    bool in_box[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        const int t = (blockIdx.x * QPT + q) * kFinThreads + threadIdx.x;
        in_box[q] = t < per_plane;
        const int row = in_box[q] ? t / p.nq : 0;
        lh[q] = p.box_lo[1] + row;
        lw[q] = p.box_lo[2] + (in_box[q] ? t - row * p.nq : 0) * 4;
        in_box[q] = in_box[q] && lw[q] < hi_w;
        const float* src = p.logits + base + static_cast<long long>(lh[q]) * g.pitch + lw[q];
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (in_box[q]) v[q][k] = ld_stream_f4(src + k * cstride);
    }
    unsigned ties = 0;
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
        if (!in_box[q]) continue;
        ArgmaxState am[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) am[e].reset();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            am[0].push(v[q][k].x, k);
            am[1].push(v[q][k].y, k);
            am[2].push(v[q][k].z, k);
            am[3].push(v[q][k].w, k);
        }
        uint8_t* lab = p.labels + (static_cast<long long>(b) * g.ext[0] + ld) * g.ext[1] * p.label_pitch +
                       static_cast<long long>(lh[q]) * p.label_pitch + lw[q];
        unsigned packed = 0;
        const int nv = min(4, hi_w - lw[q]);
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (e < nv) {
                packed |= static_cast<unsigned>(am[e].label()) << (8 * e);
                ties += am[e].near_tie(p.tie_tol) ? 1u : 0u;
            }
        if (nv == 4 && ((reinterpret_cast<uintptr_t>(lab) & 3u) == 0)) {
            *reinterpret_cast<unsigned*>(lab) = packed;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (e < nv) lab[e] = static_cast<uint8_t>(packed >> (8 * e));
        }
    }
    if (p.near_ties != nullptr) {
        const unsigned warp_ties = __reduce_add_sync(0xffffffffu, ties);
        if (warp_ties != 0u && (threadIdx.x & 31) == 0) atomicAdd(p.near_ties, static_cast<unsigned long long>(warp_ties));
    }
}

template <int MODE>
static void launch_fin(dim3 grid, cudaStream_t s, const FinParams& p) {
    if (p.g.K % 7 == 0)
        finalize_kernel<MODE, 7><<<grid, kFinThreads, 0, s>>>(p);
    else
        finalize_kernel<MODE, 8><<<grid, kFinThreads, 0, s>>>(p);
}

}  // namespace mss

using namespace mss;

extern "C" int mss_finalize_labels(const mss_layout_t* lay, const float* logits, const float* importance_map,
                                   int32_t normalise, const int32_t box_lo[3], const int32_t box_hi[3], uint8_t* labels,
                                   int32_t label_pitch_w, float* logits_out, float* probs_out, float tie_tol,
                                   unsigned long long* near_ties, void* stream) {
    FinParams p;
    int rc = make_geo(lay, &p.g);
    if (rc != MSS_OK) return rc;
    const Geo& g = p.g;
    MSS_REQUIRE(logits != nullptr && box_lo != nullptr && box_hi != nullptr, MSS_E_ARG, "finalize_labels: null argument");
    MSS_REQUIRE(labels != nullptr || logits_out != nullptr || probs_out != nullptr, MSS_E_ARG,
                "finalize_labels: nothing to write");
    MSS_REQUIRE(!normalise || importance_map != nullptr, MSS_E_ARG, "finalize_labels: normalise needs the importance map");
    MSS_REQUIRE(g.pitch % 4 == 0 && reinterpret_cast<uintptr_t>(logits) % 16 == 0, MSS_E_ALIGN,
                "finalize_labels: logits need a pitch multiple of 4 and a 16-byte aligned base");
    MSS_REQUIRE((logits_out == nullptr || reinterpret_cast<uintptr_t>(logits_out) % 16 == 0) &&
                    (probs_out == nullptr || reinterpret_cast<uintptr_t>(probs_out) % 16 == 0),
                MSS_E_ALIGN, "finalize_labels: outputs must be 16-byte aligned");
    MSS_REQUIRE(labels == nullptr || (g.K <= 255 && label_pitch_w >= g.ext[2]), MSS_E_ARG,
                "finalize_labels: uint8 labels need K <= 255 and label_pitch_w >= extent W");
    for (int a = 0; a < 3; ++a) {
        MSS_REQUIRE(0 <= box_lo[a] && box_lo[a] < box_hi[a] && box_hi[a] <= g.ext[a], MSS_E_ARG,
                    "finalize_labels: axis %d box [%d,%d) outside the buffer (%d)", a, box_lo[a], box_hi[a], g.ext[a]);
        p.box_lo[a] = box_lo[a];
        p.box_n[a] = box_hi[a] - box_lo[a];
    }
    MSS_REQUIRE(box_lo[2] % 4 == 0, MSS_E_ALIGN, "finalize_labels: box_lo W (%d) must be a multiple of 4", box_lo[2]);
    p.nq = (p.box_n[2] + 3) / 4;
    p.logits = logits;
    p.imp = normalise ? importance_map : nullptr;
    p.labels = labels;
    p.label_pitch = label_pitch_w;
    p.logits_out = logits_out;
    p.probs_out = probs_out;
    p.tie_tol = tie_tol;
    p.near_ties = near_ties;
    const long long per_plane = static_cast<long long>(p.nq) * p.box_n[1];
    dim3 grid(static_cast<unsigned>((per_plane + kFinThreads - 1) / kFinThreads), static_cast<unsigned>(p.box_n[0]),
              static_cast<unsigned>(g.nb));
    MSS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, MSS_E_UNSUPPORTED, "finalize_labels: box too large for one launch");
    cudaStream_t s = as_stream(stream);
    if (probs_out != nullptr)
        launch_fin<kFinProbs>(grid, s, p);
    else if (logits_out != nullptr)
        launch_fin<kFinNormalise>(grid, s, p);
    else if (g.K <= 4) {
        constexpr int kQ = 4;
        grid.x = static_cast<unsigned>((per_plane + kFinThreads * kQ - 1) / (kFinThreads * kQ));
        switch (g.K) {
            case 1: finalize_labels_smallk_kernel<1, kQ><<<grid, kFinThreads, 0, s>>>(p); break;
            case 2: finalize_labels_smallk_kernel<2, kQ><<<grid, kFinThreads, 0, s>>>(p); break;
            case 3: finalize_labels_smallk_kernel<3, kQ><<<grid, kFinThreads, 0, s>>>(p); break;
            default: finalize_labels_smallk_kernel<4, kQ><<<grid, kFinThreads, 0, s>>>(p); break;
        }
    } else
        launch_fin<kFinLabels>(grid, s, p);  // labels only: the weight count cannot change the argmax
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}

// ---- multi-GPU: the finalise step that IS the exchange ---------------------------------------------------------------
// One volume, its window list cut into one contiguous range per rank (flat partition): every rank holds raw weighted sums
// of ITS windows over the bounding box of their footprints (zero where none of them reaches).  The rank that owns a box of
// the volume finishes it by reading every contributing accumulator - its own and its peers', mapped over NVLink - adding
// them in ascending rank order (= ascending window order between ranks), and taking the argmax: no halo buffers, no
// send / recv, no separate add pass.
namespace mss {

constexpr int kMaxGatherSrc = 16;

struct GatherParams {
    Geo g;  // org / ext = the owned box (global frame), pitch = pitch of logits_out
    const float* src[kMaxGatherSrc];
    int so[kMaxGatherSrc][3];  // source box origin (global), extent, W pitch
    int se[kMaxGatherSrc][3];
    int sp[kMaxGatherSrc];
    int n_src;
    const float* imp;
    uint8_t* labels;
    int label_pitch;
    float* logits_out;
    float tie_tol;
    unsigned long long* near_ties;
    int nq;
};

__global__ void __launch_bounds__(kFinThreads) finalize_gather_kernel(const __grid_constant__ GatherParams p) {
    const Geo& g = p.g;
    const int t = blockIdx.x * kFinThreads + threadIdx.x;
    const bool in_box = t < p.nq * g.ext[1];
    const int row = in_box ? t / p.nq : 0;
    const int ld = blockIdx.y, lh = row, lw = (in_box ? t - row * p.nq : 0) * 4;
    const int b = blockIdx.z;
    const int gd = ld + g.org[0], gh = lh + g.org[1], gw = lw + g.org[2];
    bool valid[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) valid[e] = in_box && lw + e < g.ext[2];
    unsigned ties = 0;
    if (valid[0]) {
        const int K = g.K;
        // the accumulators that hold this quad (a box test; what a rank did not touch inside its box is zero)
        const float* ptr[kMaxGatherSrc];
        long long cs[kMaxGatherSrc];
        int n_hit = 0;
#pragma unroll
        for (int q = 0; q < kMaxGatherSrc; ++q) {
            if (q < p.n_src) {
                const int d = gd - p.so[q][0], h = gh - p.so[q][1], w = gw - p.so[q][2];
                if (static_cast<unsigned>(d) < static_cast<unsigned>(p.se[q][0]) &&
                    static_cast<unsigned>(h) < static_cast<unsigned>(p.se[q][1]) && w >= 0 && w < p.sp[q]) {
                    const long long plane = static_cast<long long>(p.se[q][1]) * p.sp[q];
                    cs[n_hit] = static_cast<long long>(p.se[q][0]) * plane;
                    ptr[n_hit] = p.src[q] + static_cast<long long>(b) * K * cs[n_hit] + static_cast<long long>(d) * plane +
                                 static_cast<long long>(h) * p.sp[q] + w;
                    ++n_hit;
                }
            }
        }
        float cnt[4] = {1.f, 1.f, 1.f, 1.f};
        if (p.logits_out != nullptr) weight_count(g, p.imp, gd, gh, gw, valid, cnt);
        ArgmaxState am[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) am[e].reset();
        const long long oplane = static_cast<long long>(g.ext[1]) * g.pitch;
        const long long ocs = static_cast<long long>(g.ext[0]) * oplane;
        for (int k = 0; k < K; ++k) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int q = 0; q < n_hit; ++q) {
                const float4 s = ld_stream_f4(ptr[q] + k * cs[q]);
                v.x = __fadd_rn(v.x, s.x);
                v.y = __fadd_rn(v.y, s.y);
                v.z = __fadd_rn(v.z, s.z);
                v.w = __fadd_rn(v.w, s.w);
            }
            if (p.logits_out != nullptr) {
                v.x = __fdiv_rn(v.x, cnt[0]);  // engine/utils.py:151
                v.y = __fdiv_rn(v.y, cnt[1]);
                v.z = __fdiv_rn(v.z, cnt[2]);
                v.w = __fdiv_rn(v.w, cnt[3]);
                float* dst = p.logits_out + (static_cast<long long>(b) * K + k) * ocs + static_cast<long long>(ld) * oplane +
                             static_cast<long long>(lh) * g.pitch + lw;
                *reinterpret_cast<float4*>(dst) = v;
            }
            am[0].push(v.x, k);
            am[1].push(v.y, k);
            am[2].push(v.z, k);
            am[3].push(v.w, k);
        }
        if (p.labels != nullptr) {
            uint8_t* lab = p.labels + (static_cast<long long>(b) * g.ext[0] + ld) * g.ext[1] * p.label_pitch +
                           static_cast<long long>(lh) * p.label_pitch + lw;
            unsigned packed = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (valid[e]) {
                    packed |= static_cast<unsigned>(am[e].label()) << (8 * e);
                    ties += am[e].near_tie(p.tie_tol) ? 1u : 0u;
                }
            if (valid[3] && ((reinterpret_cast<uintptr_t>(lab) & 3u) == 0)) {
                *reinterpret_cast<unsigned*>(lab) = packed;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (valid[e]) lab[e] = static_cast<uint8_t>(packed >> (8 * e));
            }
        }
    }
    if (p.near_ties != nullptr) {
        const unsigned warp_ties = __reduce_add_sync(0xffffffffu, ties);
        if (warp_ties != 0u && (threadIdx.x & 31) == 0) atomicAdd(p.near_ties, static_cast<unsigned long long>(warp_ties));
    }
}

}  // namespace mss

extern "C" int mss_finalize_gather(const mss_layout_t* lay, int32_t n_src, const float* const* src_acc,
                                   const int32_t* src_origin, const int32_t* src_extent, const int32_t* src_pitch_w,
                                   const float* importance_map, uint8_t* labels, int32_t label_pitch_w, float* logits_out,
                                   float tie_tol, unsigned long long* near_ties, void* stream) {
    GatherParams p;
    int rc = make_geo(lay, &p.g, false);
    if (rc != MSS_OK) return rc;
    const Geo& g = p.g;
    MSS_REQUIRE(src_acc && src_origin && src_extent && src_pitch_w, MSS_E_ARG, "finalize_gather: null argument");
    MSS_REQUIRE(n_src >= 1 && n_src <= kMaxGatherSrc, MSS_E_ARG, "finalize_gather: n_src %d outside [1, %d]", n_src,
                kMaxGatherSrc);
    MSS_REQUIRE(labels != nullptr || logits_out != nullptr, MSS_E_ARG, "finalize_gather: nothing to write");
    MSS_REQUIRE(logits_out == nullptr || importance_map != nullptr, MSS_E_ARG,
                "finalize_gather: normalised logits need the importance map");
    MSS_REQUIRE(labels == nullptr || (g.K <= 255 && label_pitch_w >= g.ext[2]), MSS_E_ARG,
                "finalize_gather: uint8 labels need K <= 255 and label_pitch_w >= extent W");
    MSS_REQUIRE(g.org[2] % 4 == 0, MSS_E_ALIGN, "finalize_gather: the owned box must start on a multiple of 4 along W");
    MSS_REQUIRE(logits_out == nullptr || (g.pitch % 4 == 0 && reinterpret_cast<uintptr_t>(logits_out) % 16 == 0), MSS_E_ALIGN,
                "finalize_gather: logits_out needs a pitch multiple of 4 and a 16-byte aligned base");
    for (int q = 0; q < n_src; ++q) {
        MSS_REQUIRE(src_acc[q] != nullptr && reinterpret_cast<uintptr_t>(src_acc[q]) % 16 == 0, MSS_E_ALIGN,
                    "finalize_gather: source %d is null or not 16-byte aligned", q);
        MSS_REQUIRE(src_pitch_w[q] % 4 == 0 && src_origin[3 * q + 2] % 4 == 0 && src_pitch_w[q] >= src_extent[3 * q + 2],
                    MSS_E_ALIGN, "finalize_gather: source %d needs W origin and pitch multiples of 4", q);
        p.src[q] = src_acc[q];
        p.sp[q] = src_pitch_w[q];
        for (int a = 0; a < 3; ++a) {
            MSS_REQUIRE(src_extent[3 * q + a] > 0 && src_origin[3 * q + a] >= 0, MSS_E_ARG,
                        "finalize_gather: source %d has an empty box", q);
            p.so[q][a] = src_origin[3 * q + a];
            p.se[q][a] = src_extent[3 * q + a];
        }
    }
    p.n_src = n_src;
    p.imp = importance_map;
    p.labels = labels;
    p.label_pitch = label_pitch_w;
    p.logits_out = logits_out;
    p.tie_tol = tie_tol;
    p.near_ties = near_ties;
    p.nq = (g.ext[2] + 3) / 4;
    const long long per_plane = static_cast<long long>(p.nq) * g.ext[1];
    dim3 grid(static_cast<unsigned>((per_plane + kFinThreads - 1) / kFinThreads), static_cast<unsigned>(g.ext[0]),
              static_cast<unsigned>(g.nb));
    MSS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, MSS_E_UNSUPPORTED, "finalize_gather: box too large for one launch");
    finalize_gather_kernel<<<grid, kFinThreads, 0, as_stream(stream)>>>(p);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}
