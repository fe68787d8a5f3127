// Dice + cross-entropy of the stitched logits against the label map, the evaluation loss of the reference
// (run_evaluation.py:53 `DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True)`, applied at engine/test.py:48 on
// `outputs.cpu(), labels.cpu()`: a 2.9 GB device-to-host copy and a CPU softmax for a cfg2 volume).
//
// One pass over the K logit planes (the same streaming pattern as finalize): per voxel softmax in registers, then
//   I[c] += p_c [y == c],   P[c] += p_c^2 (or p_c),   G[c] += [y == c],   CE += -log p_y
// kept per thread over its quads, then reduced per warp with shuffles, per CTA in shared memory and across CTAs with
// float64 atomics.  The host turns the
// 3K + 1 sums into the loss (MONAI DiceLoss: 1 - (2 I + 1e-5) / (G + P + 1e-5), mean over classes; CrossEntropyLoss:
// mean over voxels).  SURVEY.md section 8f rank 4.
#include "common.cuh"

namespace mss {

constexpr int kLossMaxK = 16;
constexpr int kLossThreads = 128;

struct LossParams {
    const float* logits;   // class c, row r, column x at logits[c * class_stride + r * row_pitch + x]
    const void* labels;    // [rows, row_len] contiguous, uint8 or float32
    int label_f32;
    long long class_stride, row_pitch, n_rows;
    int row_len, K, squared;
    double* sums;          // [3K + 1]: I[K], P[K], G[K], CE
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// E voxels per thread (2: 8-byte loads, ~80 registers, 6 CTAs/SM; the first version took 4 with 150 registers and three
// CTAs/SM and ran at 0.29 of the HBM roofline).  The sums that only the labelled class receives (I, G) live in a private
// shared-memory column per thread ([class][thread]: conflict-free, no atomics); P stays in registers.
constexpr int kLossE = 2;

// K is a template parameter: with a run-time class count the 16-wide unrolled loops carried a compare and a branch per
// class and stage (970 warp instructions per 2 voxels, 64 % issue-active: instruction-bound at 0.41 of the roofline).
template <int K>
__global__ void __launch_bounds__(kLossThreads, 6) dice_ce_kernel(const __grid_constant__ LossParams p) {
    __shared__ double sh[3 * kLossMaxK + 1];
    __shared__ float sI[K][kLossThreads], sG[K][kLossThreads];
    for (int i = threadIdx.x; i < 3 * kLossMaxK + 1; i += kLossThreads) sh[i] = 0.0;
#pragma unroll
    for (int c = 0; c < K; ++c) sI[c][threadIdx.x] = 0.f, sG[c][threadIdx.x] = 0.f;
    __syncthreads();
    const int nq = (p.row_len + kLossE - 1) / kLossE;
    const long long total = p.n_rows * nq;
    const bool vec = (p.row_len % kLossE == 0) && (p.row_pitch % kLossE == 0) && (p.class_stride % kLossE == 0) &&
                     (reinterpret_cast<uintptr_t>(p.logits) % (4 * kLossE) == 0);
    const int lane = threadIdx.x & 31;
    float aP[K], ce = 0.f;
#pragma unroll
    for (int c = 0; c < K; ++c) aP[c] = 0.f;
    auto flush = [&]() {  // warp sums -> float64 in shared memory; every lane of the warp takes part
        ce = warp_sum(ce);
        if (lane == 0 && ce != 0.f) atomicAdd(&sh[3 * kLossMaxK], static_cast<double>(ce));
        ce = 0.f;
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const float inter = warp_sum(sI[c][threadIdx.x]), psq = warp_sum(aP[c]), g = warp_sum(sG[c][threadIdx.x]);
            if (lane == 0) {
                if (inter != 0.f) atomicAdd(&sh[c], static_cast<double>(inter));
                if (psq != 0.f) atomicAdd(&sh[kLossMaxK + c], static_cast<double>(psq));
                if (g != 0.f) atomicAdd(&sh[2 * kLossMaxK + c], static_cast<double>(g));
            }
            aP[c] = 0.f;
            sI[c][threadIdx.x] = 0.f;
            sG[c][threadIdx.x] = 0.f;
        }
    };
    int iters = 0;
    // the trip count is per warp (lanes past the end idle), so the periodic flush can use full-warp shuffles
    for (long long i0 = static_cast<long long>(blockIdx.x) * kLossThreads + (threadIdx.x & ~31); i0 < total;
         i0 += static_cast<long long>(gridDim.x) * kLossThreads) {
        if ((++iters & 127) == 0) flush();  // at most 256 float additions per accumulator between float64 hand-overs
        const long long i = i0 + lane;
        if (i >= total) continue;
        const long long row = i / nq;
        const int x0 = static_cast<int>(i - row * nq) * kLossE;
        const int nv = min(kLossE, p.row_len - x0);
        float v[K][kLossE];
        const float* src = p.logits + row * p.row_pitch + x0;
#pragma unroll
        for (int c = 0; c < K; ++c) {
            if (vec) {
                float2 f;
                asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(f.x), "=f"(f.y) : "l"(src + c * p.class_stride));
                v[c][0] = f.x, v[c][1] = f.y;
            } else {
#pragma unroll
                for (int e = 0; e < kLossE; ++e) v[c][e] = e < nv ? __ldg(src + c * p.class_stride + e) : 0.f;
            }
        }
        int y[kLossE];
#pragma unroll
        for (int e = 0; e < kLossE; ++e) {
            y[e] = -1;
            if (e < nv) {
                const long long li = row * p.row_len + x0 + e;
                y[e] = p.label_f32 ? __float2int_rn(static_cast<const float*>(p.labels)[li])
                                   : static_cast<int>(static_cast<const uint8_t*>(p.labels)[li]);
            }
        }
        // softmax per voxel (max-subtracted, like torch), log-softmax of the labelled class for the cross-entropy
#pragma unroll
        for (int e = 0; e < kLossE; ++e) {
            if (e >= nv) continue;
            float m = -INFINITY;
#pragma unroll
            for (int c = 0; c < K; ++c) m = fmaxf(m, v[c][e]);
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < K; ++c) {
                v[c][e] = __expf(v[c][e] - m);  // ex2.approx path: ~2 ulp, far inside the 1e-5 bar of a 50 M-term mean
                s += v[c][e];
            }
            const float inv = __fdividef(1.f, s);
            float py = 1.f;
#pragma unroll
            for (int c = 0; c < K; ++c) {
                const float pc = v[c][e] * inv;
                aP[c] += p.squared ? pc * pc : pc;
                py = c == y[e] ? pc : py;
            }
            if (y[e] >= 0 && y[e] < K) {
                ce -= __logf(py);
                sI[y[e]][threadIdx.x] += py;
                sG[y[e]][threadIdx.x] += 1.f;
            }
        }
    }
    flush();
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * K + 1; i += kLossThreads) {
        const int q = i / K, c = i - q * K;
        const double val = i == 3 * K ? sh[3 * kLossMaxK] : sh[q * kLossMaxK + c];
        if (val != 0.0) atomicAdd(p.sums + i, val);
    }
}

}  // namespace mss

using namespace mss;

extern "C" int mss_dice_ce_sums(const float* logits, int64_t class_stride, int64_t row_pitch, int64_t n_rows, int32_t row_len,
                                int32_t n_classes, const void* labels, int32_t label_dtype, int32_t squared_pred, double* sums,
                                void* stream) {
    MSS_REQUIRE(logits && labels && sums, MSS_E_ARG, "dice_ce_sums: null argument");
    MSS_REQUIRE(n_classes >= 1 && n_classes <= kLossMaxK, MSS_E_UNSUPPORTED, "dice_ce_sums: n_classes %d outside [1, %d]",
                n_classes, kLossMaxK);
    MSS_REQUIRE(n_rows > 0 && row_len > 0 && row_pitch >= row_len && class_stride > 0, MSS_E_ARG,
                "dice_ce_sums: need positive sizes and row_pitch >= row_len");
    MSS_REQUIRE(label_dtype == 0 || label_dtype == 1, MSS_E_ARG, "dice_ce_sums: label_dtype must be 0 (uint8) or 1 (float32)");
    LossParams p;
    p.logits = logits;
    p.labels = labels;
    p.label_f32 = label_dtype;
    p.class_stride = class_stride;
    p.row_pitch = row_pitch;
    p.n_rows = n_rows;
    p.row_len = row_len;
    p.K = n_classes;
    p.squared = squared_pred != 0;
    p.sums = sums;
    const long long work = n_rows * ((row_len + kLossE - 1) / kLossE);
    long long blocks = (work + kLossThreads - 1) / kLossThreads;
    if (blocks > 148LL * 6) blocks = 148LL * 6;  // one resident wave: the per-thread partial sums are reduced once per thread
    const unsigned nb = static_cast<unsigned>(blocks);
    cudaStream_t st = as_stream(stream);
    switch (n_classes) {
#define MSS_LOSS_CASE(KK) case KK: dice_ce_kernel<KK><<<nb, kLossThreads, 0, st>>>(p); break;
        MSS_LOSS_CASE(1) MSS_LOSS_CASE(2) MSS_LOSS_CASE(3) MSS_LOSS_CASE(4) MSS_LOSS_CASE(5) MSS_LOSS_CASE(6) MSS_LOSS_CASE(7)
        MSS_LOSS_CASE(8) MSS_LOSS_CASE(9) MSS_LOSS_CASE(10) MSS_LOSS_CASE(11) MSS_LOSS_CASE(12) MSS_LOSS_CASE(13)
        MSS_LOSS_CASE(14) MSS_LOSS_CASE(15) MSS_LOSS_CASE(16)
#undef MSS_LOSS_CASE
    }
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}
