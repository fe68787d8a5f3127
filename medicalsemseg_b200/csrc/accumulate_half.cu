// Half-precision aggregation of the nnU-Net-style tiler's `all_in_gpu` branch
// (models/segmentors/nnformer_official/neural_network.py:346-372, :399-406, :420-423): the importance map, the aggregated
// results and the aggregated counts are HALF tensors there, so every step rounds to binary16:
//     patch  = half(pred_fp32 * float(w_half))                 (:566 `result_torch[:, :] *= mult`, then :400 `.half()`)
//     agg    = half(float(agg) + float(patch))                 (:405, one tile after the other, x-y-z order)
//     nb     = half(float(nb)  + float(w_half))                (:406)
//     prob   = half(float(agg) / float(nb))                    (:420)
//     seg    = first-max argmax over classes of prob           (:423)
// Output-stationary like accumulate.cu: one thread owns one voxel, walks its covering windows in ascending window index
// (the tile loop order of :385-406) once per class, and keeps the running half in a register - no accumulator traffic.
// A second stitching policy behind the same window tables; not a roofline kernel (the 8 mirrored backbone passes dominate).
#include <cuda_fp16.h>

#include "acc_common.cuh"

namespace mss {

__global__ void __launch_bounds__(256) accumulate_half_kernel(const __grid_constant__ AccParams p, float* __restrict__ probs) {
    const Geo& g = p.g;
    const int lw = blockIdx.x * 256 + threadIdx.x;
    if (lw >= g.ext[2]) return;
    const int lh = blockIdx.y % g.ext[1], ld = blockIdx.y / g.ext[1];
    const int b = blockIdx.z;
    const int gd = ld + g.org[0], gh = lh + g.org[1], gw = lw + g.org[2];
    const int rh = g.roi[1], rw = g.roi[2];
    const long long R = static_cast<long long>(g.roi[0]) * rh * rw;
    const int cvd = g.cover[0][gd], cvh = g.cover[1][gh], cvw = g.cover[2][gw];
    const int dlo = cvd & 0xffff, dhi = cvd >> 16, hlo = cvh & 0xffff, hhi = cvh >> 16, wlo = cvw & 0xffff, whi = cvw >> 16;
    const long long vol0 = static_cast<long long>(b) * g.n_local;

    // count of weights: the same for every class (the reference replicates it K times)
    __half nb = __float2half_rn(0.f);
    for (int id = dlo; id < dhi; ++id)
        for (int ih = hlo; ih < hhi; ++ih)
            for (int iw = wlo; iw < whi; ++iw) {
                const int o = ((gd - g.starts[0][id]) * rh + (gh - g.starts[1][ih])) * rw + (gw - g.starts[2][iw]);
                nb = __float2half_rn(__fadd_rn(__half2float(nb), __ldg(p.imp + o)));
            }
    const float nbf = __half2float(nb);
    const long long plane = static_cast<long long>(g.ext[1]) * g.pitch;
    const long long cstride = static_cast<long long>(g.ext[0]) * plane;
    float* out = probs + static_cast<long long>(b) * g.K * cstride + static_cast<long long>(ld) * plane +
                 static_cast<long long>(lh) * g.pitch + lw;
    float best = __int_as_float(0xff800000);
    int best_k = 0;
    bool poisoned = false;
    for (int k = 0; k < g.K; ++k) {
        __half agg = __float2half_rn(0.f);
        for (int id = dlo; id < dhi; ++id)
            for (int ih = hlo; ih < hhi; ++ih)
                for (int iw = wlo; iw < whi; ++iw) {
                    const long long n = (static_cast<long long>(id) * g.ns[1] + ih) * g.ns[2] + iw;
                    const long long gi = vol0 + n;
                    const long long bi = gi / p.sw_batch;
                    const float* base = static_cast<const float*>(p.batch[bi]) + ((gi - bi * p.sw_batch) * g.K + k) * R;
                    const int o = ((gd - g.starts[0][id]) * rh + (gh - g.starts[1][ih])) * rw + (gw - g.starts[2][iw]);
                    const __half patch = __float2half_rn(__fmul_rn(__ldg(base + o), __ldg(p.imp + o)));
                    agg = __float2half_rn(__fadd_rn(__half2float(agg), __half2float(patch)));
                }
        const float prob = __half2float(__float2half_rn(__fdiv_rn(__half2float(agg), nbf)));
        out[k * cstride] = prob;
        poisoned |= prob != prob;  // torch.argmax treats NaN as the maximum; report the first NaN like it does
        if (prob > best) {
            best = prob;
            best_k = k;
        }
    }
    if (p.labels != nullptr) {
        int label = best_k;
        if (poisoned)
            for (int k = 0; k < g.K; ++k) {
                const float v = out[k * cstride];
                if (v != v) {
                    label = k;
                    break;
                }
            }
        p.labels[(static_cast<long long>(b) * g.ext[0] + ld) * g.ext[1] * p.label_pitch + static_cast<long long>(lh) * p.label_pitch +
                 lw] = static_cast<uint8_t>(label);
    }
}

}  // namespace mss

using namespace mss;

extern "C" int mss_accumulate_half(const mss_layout_t* lay, const void* const* batch_ptrs, int32_t n_batches, int32_t sw_batch,
                                   const float* importance_map_half_valued, float* probs_out, uint8_t* labels,
                                   int32_t label_pitch_w, void* stream) {
    AccParams q;
    int rc = make_geo(lay, &q.g);
    if (rc != MSS_OK) return rc;
    const Geo& g = q.g;
    MSS_REQUIRE(batch_ptrs != nullptr && importance_map_half_valued != nullptr && probs_out != nullptr, MSS_E_ARG,
                "accumulate_half: null argument");
    MSS_REQUIRE(n_batches > 0 && n_batches <= MSS_MAX_BATCH_PTRS && sw_batch > 0, MSS_E_ARG,
                "accumulate_half: n_batches %d outside [1, %d]", n_batches, MSS_MAX_BATCH_PTRS);
    const long long total = g.n_local * g.nb;
    MSS_REQUIRE(static_cast<long long>(n_batches) * sw_batch >= total && static_cast<long long>(n_batches - 1) * sw_batch < total,
                MSS_E_ARG, "accumulate_half: %d batches of %d do not hold the %lld windows (one call takes them all)", n_batches,
                sw_batch, total);
    for (int a = 0; a < 3; ++a)
        MSS_REQUIRE(g.wlo[a] == 0 && g.whi[a] == g.ns[a] && g.org[a] == 0 && g.ext[a] == g.img[a], MSS_E_ARG,
                    "accumulate_half: the buffer must be the whole stitched volume (axis %d)", a);
    MSS_REQUIRE(labels == nullptr || (g.K <= 255 && label_pitch_w >= g.ext[2]), MSS_E_ARG,
                "accumulate_half: uint8 labels need K <= 255 and label_pitch_w >= extent W");
    MSS_REQUIRE(static_cast<long long>(g.ext[0]) * g.ext[1] <= 0x7fffffffLL && g.nb <= 65535, MSS_E_UNSUPPORTED,
                "accumulate_half: volume too large for one launch");
    for (int i = 0; i < n_batches; ++i) {
        MSS_REQUIRE(batch_ptrs[i] != nullptr, MSS_E_ARG, "accumulate_half: batch pointer %d is null", i);
        q.batch[i] = batch_ptrs[i];
    }
    q.sw_batch = sw_batch;
    q.imp = importance_map_half_valued;
    q.labels = labels;
    q.label_pitch = label_pitch_w;
    dim3 grid(static_cast<unsigned>((g.ext[2] + 255) / 256), static_cast<unsigned>(g.ext[0] * g.ext[1]),
              static_cast<unsigned>(g.nb));
    accumulate_half_kernel<<<grid, 256, 0, as_stream(stream)>>>(q, probs_out);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}
