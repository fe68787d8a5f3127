// Test-time intensity transforms of the volume, one fused elementwise pass
// (replaces the chain built by data/dataset_builder.py:322-370: ScaleIntensityRange / ScaleCubedIntensityRange
// (data/transforms.py:17-71) followed by NormalizeIntensity - the step right before the sliding window,
// SURVEY.md section 8f rank 2).
//
//   [x = cbrt(x)]                                     MSS_INT_CBRT      data/transforms.py:54
//   [x = (x - a_min) / (a_max - a_min)]               MSS_INT_SCALE     data/transforms.py:62
//   [x = x * (b_max - b_min) + b_min]                 MSS_INT_RESCALE   data/transforms.py:63-64
//   [x = clip(x, b_min, b_max)]                       MSS_INT_CLIP_LO / _HI   data/transforms.py:65-66
//   [x = (x - sub) / div  (only where x != 0)]        MSS_INT_NORM (+ MSS_INT_NONZERO)   MONAI NormalizeIntensity
// Every step is a separately rounded operation (sub, div, mul, add; no FMA contraction) in the order the NumPy /
// torch expression evaluates it - in float32, or with MSS_INT_F64 in float64 rounded to float32 once at the end,
// which is what NumPy >= 2 does to the reference's cubed scaler (it subtracts the float64 scalar np.cbrt(a_min) from
// a float32 array).  The fixed-range and normalise paths are bit-identical to NumPy; cbrtf differs from glibc's by
// at most 1 ulp.  8 bytes of traffic per voxel, 16-byte accesses.
#include "common.cuh"

namespace mss {

struct IntensityParams {
    int flags;
    double a_min, denom, b_scale, b_min, b_max, sub, div;
};

template <typename T>
struct Arith;
template <>
struct Arith<float> {
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
};
template <>
struct Arith<double> {
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
};

// The expensive steps (cube root, the two IEEE divisions) are compile-time switches; rescale, clip and the non-zero
// mask are cheap and stay runtime flags / neutral bounds.
template <typename T, bool CBRT, bool SCALE, bool NORM>
__device__ __forceinline__ float intensity_one(float x0, const IntensityParams& p, T lo, T hi) {
    using A = Arith<T>;
    if (CBRT) x0 = cbrtf(x0);  // np.cbrt of a float32 array stays float32
    T x = static_cast<T>(x0);
    if (SCALE) x = A::div(A::sub(x, static_cast<T>(p.a_min)), static_cast<T>(p.denom));
    if (p.flags & MSS_INT_RESCALE) x = A::add(A::mul(x, static_cast<T>(p.b_scale)), static_cast<T>(p.b_min));
    x = (x < lo) ? lo : x;  // np.clip / torch.clamp: NaN stays NaN; bounds are -inf / +inf when a side is off
    x = (x > hi) ? hi : x;
    if (NORM) {
        if (!(p.flags & MSS_INT_NONZERO) || x != static_cast<T>(0))
            x = A::div(A::sub(x, static_cast<T>(p.sub)), static_cast<T>(p.div));
    }
    return static_cast<float>(x);
}

template <bool VEC, typename T, bool CBRT, bool SCALE, bool NORM>
__global__ void __launch_bounds__(256) intensity_kernel(const float* __restrict__ in, float* __restrict__ out, long long n,
                                                        const IntensityParams p) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long n4 = VEC ? n / 4 : 0;
    const T inf = static_cast<T>(__int_as_float(0x7f800000));
    const T lo = (p.flags & MSS_INT_CLIP_LO) ? static_cast<T>(p.b_min) : -inf;
    const T hi = (p.flags & MSS_INT_CLIP_HI) ? static_cast<T>(p.b_max) : inf;
    constexpr int U = 4;  // independent 16-byte loads in flight per thread
    for (long long i = tid; i < n4; i += stride * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < n4) v[u] = ld_stream_f4(in + (i + u * stride) * 4);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < n4) {
                float4 r;
                r.x = intensity_one<T, CBRT, SCALE, NORM>(v[u].x, p, lo, hi);
                r.y = intensity_one<T, CBRT, SCALE, NORM>(v[u].y, p, lo, hi);
                r.z = intensity_one<T, CBRT, SCALE, NORM>(v[u].z, p, lo, hi);
                r.w = intensity_one<T, CBRT, SCALE, NORM>(v[u].w, p, lo, hi);
                *reinterpret_cast<float4*>(out + (i + u * stride) * 4) = r;
            }
    }
    for (long long i = n4 * 4 + tid; i < n; i += stride) out[i] = intensity_one<T, CBRT, SCALE, NORM>(in[i], p, lo, hi);
}

template <bool VEC, typename T>
static void launch_intensity(unsigned nb, cudaStream_t s, const float* in, float* out, long long n, const IntensityParams& p) {
    const int key = ((p.flags & MSS_INT_CBRT) ? 4 : 0) | ((p.flags & MSS_INT_SCALE) ? 2 : 0) | ((p.flags & MSS_INT_NORM) ? 1 : 0);
    switch (key) {
        case 0: intensity_kernel<VEC, T, false, false, false><<<nb, 256, 0, s>>>(in, out, n, p); break;
        case 1: intensity_kernel<VEC, T, false, false, true><<<nb, 256, 0, s>>>(in, out, n, p); break;
        case 2: intensity_kernel<VEC, T, false, true, false><<<nb, 256, 0, s>>>(in, out, n, p); break;
        case 3: intensity_kernel<VEC, T, false, true, true><<<nb, 256, 0, s>>>(in, out, n, p); break;
        case 4: intensity_kernel<VEC, T, true, false, false><<<nb, 256, 0, s>>>(in, out, n, p); break;
        case 5: intensity_kernel<VEC, T, true, false, true><<<nb, 256, 0, s>>>(in, out, n, p); break;
        case 6: intensity_kernel<VEC, T, true, true, false><<<nb, 256, 0, s>>>(in, out, n, p); break;
        default: intensity_kernel<VEC, T, true, true, true><<<nb, 256, 0, s>>>(in, out, n, p); break;
    }
}

}  // namespace mss

using namespace mss;

extern "C" int mss_intensity_transform(const float* in, float* out, int64_t n, int32_t flags, double a_min,
                                       double a_max_minus_a_min, double b_min, double b_max, double subtrahend, double divisor,
                                       void* stream) {
    MSS_REQUIRE(in != nullptr && out != nullptr && n > 0, MSS_E_ARG, "intensity_transform: null argument or empty volume");
    MSS_REQUIRE((flags & ~0xff) == 0, MSS_E_ARG, "intensity_transform: unknown flag bits 0x%x", flags);
    IntensityParams p;
    p.flags = flags;
    const bool f64 = (flags & MSS_INT_F64) != 0;
    p.a_min = a_min;
    p.denom = a_max_minus_a_min;
    p.b_min = b_min;
    p.b_max = b_max;
    // (b_max - b_min) as the host expression computes it: float64 difference (Python floats), rounded to the working type
    p.b_scale = b_max - b_min;
    p.sub = subtrahend;
    p.div = divisor;
    if (!f64) {  // float32 arithmetic: constants are rounded to float32 once, like NumPy does with Python scalars
        p.a_min = static_cast<float>(p.a_min), p.denom = static_cast<float>(p.denom), p.b_min = static_cast<float>(p.b_min);
        p.b_max = static_cast<float>(p.b_max), p.b_scale = static_cast<float>(p.b_scale);
        p.sub = static_cast<float>(p.sub), p.div = static_cast<float>(p.div);
    }
    const bool vec = reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
    long long blocks = (n / 16 + 255) / 256 + 1;
    if (blocks > 148LL * 8) blocks = 148LL * 8;
    const unsigned nb = static_cast<unsigned>(blocks);
    cudaStream_t s = as_stream(stream);
    if (vec && f64) launch_intensity<true, double>(nb, s, in, out, n, p);
    else if (vec) launch_intensity<true, float>(nb, s, in, out, n, p);
    else if (f64) launch_intensity<false, double>(nb, s, in, out, n, p);
    else launch_intensity<false, float>(nb, s, in, out, n, p);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}
