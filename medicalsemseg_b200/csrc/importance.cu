// Importance-map generation (replaces MONAI compute_importance_map called at engine/utils.py:113-115).
//
// MONAI 0.8 filters a unit impulse with three zero-padded 1-D gaussian passes (axis 0 first) in
// float32, which is exactly the rounded outer product ((p_d[i] * p_h[j]) * p_w[k]) of the three
// per-axis profiles; it then divides by the maximum and clamps to the smallest non-zero entry.
// The kernels below reproduce that rounding order: __fmul_rn twice, __fdiv_rn once.
#include "common.cuh"

namespace mss {

// MONAI 0.8 gaussian_1d(approx="erf") taps laid on the patch axis with the impulse at n//2:
//   profile[i] = 0.5 * (erf(t (x + .5)) - erf(t (x - .5))),  x = (n//2) - i,  t = 0.70710678 / sigma,
// zero beyond the truncation tail int(max(4 sigma, .5) + .5).
__global__ void gaussian_profile_kernel(float* __restrict__ out, int n, float sigma, int variant) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (variant == MSS_GAUSS_MONAI08_ERF) {
        const int tail = static_cast<int>(fmaxf(sigma * 4.0f, 0.5f) + 0.5f);
        const int xi = n / 2 - i;
        float v = 0.0f;
        if (xi >= -tail && xi <= tail) {
            const float x = static_cast<float>(xi);
            const float t = __fdiv_rn(0.70710678f, fabsf(sigma));
            const float hi = erff(__fmul_rn(t, __fadd_rn(x, 0.5f)));
            const float lo = erff(__fmul_rn(t, __fsub_rn(x, 0.5f)));
            v = fmaxf(__fmul_rn(0.5f, __fsub_rn(hi, lo)), 0.0f);
        }
        out[i] = v;
    } else {
        // MONAI >= 1.2: x on the half-integer grid -(n-1)/2 .. (n-1)/2, exp(x^2 / (-2 sigma^2))
        const float x = __fadd_rn(-(static_cast<float>(n) - 1.0f) * 0.5f, static_cast<float>(i));
        const float den = __fmul_rn(-2.0f, __fmul_rn(sigma, sigma));
        out[i] = expf(__fdiv_rn(__fmul_rn(x, x), den));
    }
}

__device__ __forceinline__ float block_max_of(const float* __restrict__ p, int n, float* sh) {
    float m = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, p[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    float r = 0.0f;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) r = fmaxf(r, sh[w]);
    return r;
}

// pass 1: map = ((pd*ph)*pw)/max, track min over non-zero entries (scratch[0], float bits) and over all (scratch[1])
__global__ void importance_outer_kernel(float* __restrict__ map, int rd, int rh, int rw, const float* __restrict__ pd,
                                        const float* __restrict__ ph, const float* __restrict__ pw,
                                        unsigned int* __restrict__ scratch, int divide_by_max) {
    __shared__ float sh[32];
    const float md = block_max_of(pd, rd, sh);
    const float mh = block_max_of(ph, rh, sh);
    const float mw = block_max_of(pw, rw, sh);
    const float vmax = __fmul_rn(__fmul_rn(md, mh), mw);  // rounding is monotone: max of the rounded products
    const long long total = static_cast<long long>(rd) * rh * rw;
    unsigned int lmin_nz = 0x7f800000u, lmin = 0x7f800000u;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(e % rw);
        const long long r = e / rw;
        const int j = static_cast<int>(r % rh);
        const int i = static_cast<int>(r / rh);
        float v = __fmul_rn(__fmul_rn(pd[i], ph[j]), pw[k]);
        if (divide_by_max) v = __fdiv_rn(v, vmax);  // MONAI 0.8 normalises by the maximum, >= 1.2 does not
        map[e] = v;
        const unsigned int bits = __float_as_uint(v);  // v >= 0: float order == unsigned order
        lmin = min(lmin, bits);
        if (v != 0.0f) lmin_nz = min(lmin_nz, bits);
    }
    lmin_nz = __reduce_min_sync(0xffffffffu, lmin_nz);
    lmin = __reduce_min_sync(0xffffffffu, lmin);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&scratch[0], lmin_nz);
        atomicMin(&scratch[1], lmin);
    }
}

// pass 2: clamp to the floor: min non-zero (MONAI 0.8) or max(min, floor_abs) (MONAI >= 1.2)
__global__ void importance_clamp_kernel(float* __restrict__ map, long long total, const unsigned int* __restrict__ scratch,
                                        float floor_abs) {
    const float floor_v = floor_abs > 0.0f ? fmaxf(__uint_as_float(scratch[1]), floor_abs) : __uint_as_float(scratch[0]);
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x)
        map[e] = fmaxf(map[e], floor_v);
}

__global__ void fill_kernel(float* __restrict__ p, long long total, float v) {
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x)
        p[e] = v;
}

}  // namespace mss

using namespace mss;

extern "C" {

int mss_gaussian_profile(float* profile_out, int32_t n, float sigma, int32_t variant, void* stream) {
    MSS_REQUIRE(profile_out != nullptr && n > 0, MSS_E_ARG, "gaussian_profile: null output or n <= 0");
    MSS_REQUIRE(sigma > 0.0f, MSS_E_ARG, "gaussian_profile: sigma must be positive");
    MSS_REQUIRE(variant == MSS_GAUSS_MONAI08_ERF || variant == MSS_GAUSS_MONAI12_EXP, MSS_E_ARG,
                "gaussian_profile: unknown variant %d", variant);
    gaussian_profile_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(profile_out, n, sigma, variant);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}

int mss_importance_map(float* map_out, const int32_t roi[3], int32_t mode, const float* prof_d, const float* prof_h,
                       const float* prof_w, float floor_abs, void* scratch, void* stream) {
    MSS_REQUIRE(map_out != nullptr && roi != nullptr, MSS_E_ARG, "importance_map: null argument");
    MSS_REQUIRE(roi[0] > 0 && roi[1] > 0 && roi[2] > 0, MSS_E_ARG, "importance_map: roi must be positive");
    const long long total = static_cast<long long>(roi[0]) * roi[1] * roi[2];
    const int blocks = static_cast<int>(total / 256 + 1 < 148 * 8 ? total / 256 + 1 : 148 * 8);
    cudaStream_t s = as_stream(stream);
    if (mode == MSS_BLEND_CONSTANT) {
        fill_kernel<<<blocks, 256, 0, s>>>(map_out, total, 1.0f);
        MSS_CUDA(cudaGetLastError());
        return MSS_OK;
    }
    MSS_REQUIRE(mode == MSS_BLEND_PROFILES, MSS_E_ARG, "importance_map: unknown mode %d", mode);
    MSS_REQUIRE(prof_d && prof_h && prof_w && scratch, MSS_E_ARG, "importance_map: profiles / scratch are null");
    MSS_CUDA(cudaMemsetAsync(scratch, 0x7f, 8, s));  // 0x7f7f7f7f = 3.39e38: above every weight
    importance_outer_kernel<<<blocks, 256, 0, s>>>(map_out, roi[0], roi[1], roi[2], prof_d, prof_h, prof_w,
                                                   static_cast<unsigned int*>(scratch), floor_abs > 0.0f ? 0 : 1);
    MSS_CUDA(cudaGetLastError());
    importance_clamp_kernel<<<blocks, 256, 0, s>>>(map_out, total, static_cast<const unsigned int*>(scratch),
                                                   floor_abs);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}

}  // extern "C"
