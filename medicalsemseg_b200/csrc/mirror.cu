// Mirror test-time augmentation around the backbone, the nnU-Net-style predictor of the reference
// (models/segmentors/nnformer_official/neural_network.py:511-568 `_internal_maybe_mirror_and_pred_3D`):
//   result = 0;  for every enabled mirror m:  result += (1 / n) * flip_m(softmax(net(flip_m(x))))
// flip_m mirrors the spatial axes named by the bits of m (bit 0 = W, bit 1 = H, bit 2 = D - the reference's
// torch.flip dims 4, 3, 2).  Two kernels replace the 2 x 8 torch.flip copies and the 8 scaled adds:
//   mss_flip_copy    - one mirrored copy of the patch batch for the backbone's input;
//   mss_mirror_merge - reads the n predictions ONCE, un-mirrors on the fly and accumulates them in the reference's
//                      order with the reference's roundings: fadd_rn(result, fmul_rn(1/n, pred_m)), 1/n a power of two.
// One thread owns 4 consecutive W voxels of the result; a W mirror reads the mirrored aligned quad and reverses it
// in registers, so every access stays a 16-byte vector (W % 4 == 0; scalar kernel otherwise).
#include "common.cuh"

namespace mss {

constexpr int kMaxMirrors = 8;

struct MirrorParams {
    const float* src[kMaxMirrors];
    int mask[kMaxMirrors];
    int n_terms;
    float scale;
    float* out;
    long long n_outer;  // batch * channels
    int d, h, w;
};

__device__ __forceinline__ float4 reverse4(float4 v) { return make_float4(v.w, v.z, v.y, v.x); }

template <bool VEC, bool MERGE>
__global__ void __launch_bounds__(256) mirror_kernel(const __grid_constant__ MirrorParams p) {
    constexpr int E = VEC ? 4 : 1;
    const int wq = p.w / E;
    const int per = p.d * p.h * wq;  // < 2^31 (checked by the host); 32-bit index math, volumes on grid.y
    // volumes per thread and step: the plain flip copy wants 2 loads in flight per thread (98 % of peak), the merge
    // already has n_terms and wants resident warps instead (U = 2 cost it a third of its bandwidth on B200)
    constexpr int U = MERGE ? 1 : 2;
    const long long vol = static_cast<long long>(p.d) * p.h * p.w;
    for (long long o0 = static_cast<long long>(blockIdx.y) * U; o0 < p.n_outer; o0 += static_cast<long long>(gridDim.y) * U)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
            int r = i;
            const int q = r % wq;
            r /= wq;
            const int y = r % p.h, z = r / p.h;
            float4 v[U][kMaxMirrors];
#pragma unroll
            for (int m = 0; m < kMaxMirrors; ++m) {
                if (m < p.n_terms) {
                    const int mk = p.mask[m];
                    const int zz = (mk & 4) ? p.d - 1 - z : z;
                    const int yy = (mk & 2) ? p.h - 1 - y : y;
                    const int qq = (mk & 1) ? wq - 1 - q : q;
                    const long long off = (static_cast<long long>(zz) * p.h + yy) * p.w + qq * E;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (o0 + u < p.n_outer) {
                            const float* s = p.src[m] + (o0 + u) * vol + off;
                            if (VEC) {
                                v[u][m] = ld_stream_f4(s);
                                if (mk & 1) v[u][m] = reverse4(v[u][m]);
                            } else {
                                v[u][m].x = __ldg(s);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (o0 + u >= p.n_outer) break;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (MERGE) {
#pragma unroll
                    for (int m = 0; m < kMaxMirrors; ++m) {
                        if (m < p.n_terms) {  // result_torch += 1 / num_results * pred   (neural_network.py:537-565)
                            acc.x = __fadd_rn(acc.x, __fmul_rn(p.scale, v[u][m].x));
                            if (VEC) {
                                acc.y = __fadd_rn(acc.y, __fmul_rn(p.scale, v[u][m].y));
                                acc.z = __fadd_rn(acc.z, __fmul_rn(p.scale, v[u][m].z));
                                acc.w = __fadd_rn(acc.w, __fmul_rn(p.scale, v[u][m].w));
                            }
                        }
                    }
                } else {
                    acc = v[u][0];
                }
                float* dst = p.out + (o0 + u) * vol + (static_cast<long long>(z) * p.h + y) * p.w + q * E;
                if (VEC) *reinterpret_cast<float4*>(dst) = acc;
                else *dst = acc.x;
            }
        }
}

static int launch_mirror(const MirrorParams& p, bool merge, cudaStream_t s) {
    bool vec = p.w % 4 == 0 && reinterpret_cast<uintptr_t>(p.out) % 16 == 0;
    for (int m = 0; m < p.n_terms; ++m)
        if (reinterpret_cast<uintptr_t>(p.src[m]) % 16 != 0) vec = false;
    const long long per = static_cast<long long>(p.d) * p.h * (vec ? p.w / 4 : p.w);
    MSS_REQUIRE(per < (1LL << 31), MSS_E_UNSUPPORTED, "mirror: one volume exceeds 2^31 elements");
    long long bx = (per + 255) / 256;
    if (bx > 148LL * 8) bx = 148LL * 8;
    const long long by = merge ? p.n_outer : (p.n_outer + 1) / 2;
    const dim3 nb(static_cast<unsigned>(bx), static_cast<unsigned>(by < 65535 ? by : 65535));
    if (vec && merge) mirror_kernel<true, true><<<nb, 256, 0, s>>>(p);
    else if (vec) mirror_kernel<true, false><<<nb, 256, 0, s>>>(p);
    else if (merge) mirror_kernel<false, true><<<nb, 256, 0, s>>>(p);
    else mirror_kernel<false, false><<<nb, 256, 0, s>>>(p);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}

}  // namespace mss

using namespace mss;

extern "C" int mss_flip_copy(const float* in, float* out, int64_t n_outer, const int32_t dims[3], int32_t mirror_mask,
                             void* stream) {
    MSS_REQUIRE(in != nullptr && out != nullptr && dims != nullptr && in != out, MSS_E_ARG,
                "flip_copy: null argument (or in == out: the copy is not in place)");
    MSS_REQUIRE(n_outer > 0 && dims[0] > 0 && dims[1] > 0 && dims[2] > 0, MSS_E_ARG, "flip_copy: sizes must be positive");
    MSS_REQUIRE(mirror_mask >= 0 && mirror_mask < 8, MSS_E_ARG, "flip_copy: mirror_mask %d outside [0, 8)", mirror_mask);
    MirrorParams p;
    for (int m = 0; m < kMaxMirrors; ++m) p.src[m] = nullptr, p.mask[m] = 0;
    p.src[0] = in;
    p.mask[0] = mirror_mask;
    p.n_terms = 1;
    p.scale = 1.f;
    p.out = out;
    p.n_outer = n_outer;
    p.d = dims[0], p.h = dims[1], p.w = dims[2];
    return launch_mirror(p, false, as_stream(stream));
}

extern "C" int mss_mirror_merge(const float* const* preds, const int32_t* mirror_masks, int32_t n_terms, float scale,
                                float* out, int64_t n_outer, const int32_t dims[3], void* stream) {
    MSS_REQUIRE(preds != nullptr && mirror_masks != nullptr && out != nullptr && dims != nullptr, MSS_E_ARG,
                "mirror_merge: null argument");
    MSS_REQUIRE(n_terms >= 1 && n_terms <= kMaxMirrors, MSS_E_ARG, "mirror_merge: n_terms %d outside [1, %d]", n_terms,
                kMaxMirrors);
    MSS_REQUIRE(n_outer > 0 && dims[0] > 0 && dims[1] > 0 && dims[2] > 0, MSS_E_ARG, "mirror_merge: sizes must be positive");
    MirrorParams p;
    for (int m = 0; m < kMaxMirrors; ++m) p.src[m] = nullptr, p.mask[m] = 0;
    for (int m = 0; m < n_terms; ++m) {
        MSS_REQUIRE(preds[m] != nullptr && preds[m] != out, MSS_E_ARG, "mirror_merge: prediction %d is null or aliases out", m);
        MSS_REQUIRE(mirror_masks[m] >= 0 && mirror_masks[m] < 8, MSS_E_ARG, "mirror_merge: mask %d outside [0, 8)", mirror_masks[m]);
        p.src[m] = preds[m];
        p.mask[m] = mirror_masks[m];
    }
    p.n_terms = n_terms;
    p.scale = scale;
    p.out = out;
    p.n_outer = n_outer;
    p.d = dims[0], p.h = dims[1], p.w = dims[2];
    return launch_mirror(p, true, as_stream(stream));
}
