// Shared host/device plumbing for libmss_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/mss_b200.h"

namespace mss {

// ---- error text (per calling thread) -------------------------------------------------------
char* last_error_buf();
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define MSS_REQUIRE(cond, code, ...)                   \
    do {                                               \
        if (!(cond)) return ::mss::fail((code), __VA_ARGS__); \
    } while (0)

#define MSS_CUDA(expr)                                           \
    do {                                                         \
        cudaError_t _e = (expr);                                 \
        if (_e != cudaSuccess) return ::mss::cuda_fail(_e, #expr); \
    } while (0)

// ---- geometry table layout (int32 words) ---------------------------------------------------
constexpr int32_t kTableMagic = 0x4D535331;  // "MSS1"
constexpr int kHdrMagic = 0, kHdrImage = 1, kHdrRoi = 4, kHdrN = 7, kHdrOffStarts = 10, kHdrOffCover = 13,
              kHdrWords = 16;

// Geometry handed to kernels by value.  Pointers are device pointers into the geometry table.
struct Geo {
    int img[3];
    int roi[3];
    int ns[3];    // starts per axis (whole grid)
    int wlo[3];   // owned window index box
    int whi[3];
    int nwl[3];   // whi - wlo
    int org[3];   // global coordinate of buffer voxel (0,0,0)
    int ext[3];   // buffer dims
    int pitch;    // accumulator row pitch (elements)
    int nb;       // volumes
    int K;        // classes
    long long n_local;  // owned windows per volume
    const int* starts[3];
    const int* cover[3];  // per global coordinate: lo | (hi << 16), hi exclusive
};

// Validates `lay` and fills `g`; host side only.  windows_inside: the owned windows must lie inside the buffer box (not
// the case for the box a rank merely OWNS in mss_finalize_gather).
int make_geo(const mss_layout_t* lay, Geo* g, bool windows_inside = true);

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#ifdef __CUDACC__
// owned-window enumeration index -> (volume, window index per axis)
__device__ __forceinline__ void decode_window(const Geo& g, long long idx, int& b, int& id, int& ih, int& iw) {
    b = static_cast<int>(idx / g.n_local);
    int n = static_cast<int>(idx - static_cast<long long>(b) * g.n_local);
    iw = n % g.nwl[2] + g.wlo[2];
    n /= g.nwl[2];
    ih = n % g.nwl[1] + g.wlo[1];
    id = n / g.nwl[1] + g.wlo[0];
}

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming 16-byte load that does not allocate in L1 (data touched once)
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
#endif

}  // namespace mss
