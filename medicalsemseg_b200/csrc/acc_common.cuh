// Pieces shared by the two accumulation kernels (accumulate.cu: the general per-quad kernel; accumulate_cells.cu: the
// cell-uniform kernel): launch parameters, logit loads per dtype, the per-thread cp.async ring.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "labels.cuh"

namespace mss {

constexpr int kAccThreads = 128;
constexpr int kMaxCand = 64;  // windows listed per shared-memory chunk (general kernel)
constexpr int kCellMaxSeg = 64;  // segments per axis / W tiles the cell kernel's tables hold

struct AccParams {
    Geo g;
    const void* batch[MSS_MAX_BATCH_PTRS];
    int sw_batch;
    long long g0, g1;  // owned-window range of this call, over n_volumes * n_local
    long long own0, own1;  // windows of the box outside [own0, own1) belong to another rank: they do not exist here
    const float* imp;
    float* acc;
    uint8_t* labels;
    int label_pitch;
    int fuse;
    float tie_tol;
    unsigned long long* near_ties;
    int box_lo[3];  // local box this launch covers; box_lo[2] is a multiple of 4
    int box_n[3];
    int nq;        // quads per row of the box
    int tq, th;    // tile: quads per row, rows
    int n_wtiles;  // tiles along W
    int b_lo;      // first volume touched
    int vec_ok;    // logits pointers and roi allow 16-byte (8-byte for 16-bit logits) vector loads
};

template <typename LT>
struct LogitLoad;
template <>
struct LogitLoad<float> {
    static __device__ __forceinline__ float4 quad(const float* p) { return ld_stream_f4(p); }
    static __device__ __forceinline__ float one(const float* p) { return __ldg(p); }
};
template <>
struct LogitLoad<__half> {
    static __device__ __forceinline__ float4 quad(const __half* p) {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ float one(const __half* p) { return __half2float(__ldg(p)); }
};
template <>
struct LogitLoad<__nv_bfloat16> {
    static __device__ __forceinline__ float4 quad(const __nv_bfloat16* p) {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                           __uint_as_float(r.y & 0xffff0000u));
    }
    static __device__ __forceinline__ float one(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }
};

__device__ __forceinline__ float& comp(float4& v, int e) { return e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w)); }

enum : int { kBefore = 0, kNow = 1, kAfter = 2, kForeign = 3 };

// ---- per-thread asynchronous copies into the thread's own shared-memory ring slots -----------------------
// (shared-memory operands are 32-bit shared-window addresses computed once per thread)
__device__ __forceinline__ void cp_async16(unsigned smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ float4 lds_f4(unsigned smem) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(smem));
    return r;
}
__device__ __forceinline__ uint2 lds_u2(unsigned smem) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(smem));
    return r;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename LT>
struct Slot;  // one thread's 4 logits of one class in the ring
template <>
struct Slot<float> {
    static constexpr int kBytes = 16;
    static constexpr bool kUnaligned = true;  // a quad at any 4-byte offset can be fetched as four 4-byte copies
    static __device__ __forceinline__ void fetch(unsigned s, const void* g) { cp_async16(s, g); }
    static __device__ __forceinline__ void fetch_unaligned(unsigned s, const void* g) {
        const char* b = static_cast<const char*>(g);
        cp_async4(s, b), cp_async4(s + 4, b + 4), cp_async4(s + 8, b + 8), cp_async4(s + 12, b + 12);
    }
    static __device__ __forceinline__ float4 read(unsigned s) { return lds_f4(s); }
};
template <>
struct Slot<__half> {
    static constexpr int kBytes = 8;
    static constexpr bool kUnaligned = false;
    static __device__ __forceinline__ void fetch(unsigned s, const void* g) { cp_async8(s, g); }
    static __device__ __forceinline__ void fetch_unaligned(unsigned, const void*) {}
    static __device__ __forceinline__ float4 read(unsigned s) {
        const uint2 r = lds_u2(s);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
};
template <>
struct Slot<__nv_bfloat16> {
    static constexpr int kBytes = 8;
    static constexpr bool kUnaligned = false;
    static __device__ __forceinline__ void fetch(unsigned s, const void* g) { cp_async8(s, g); }
    static __device__ __forceinline__ void fetch_unaligned(unsigned, const void*) {}
    static __device__ __forceinline__ float4 read(unsigned s) {
        const uint2 r = lds_u2(s);
        return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                           __uint_as_float(r.y & 0xffff0000u));
    }
};

template <typename LT, int KC, int S>
constexpr size_t acc_smem_bytes() {
    return static_cast<size_t>(S) * kAccThreads * (16 + KC * Slot<LT>::kBytes);
}


// accumulate_rows.cu (few classes, every window in one launch, labels out)
int launch_rows(const mss_layout_t* lay, const AccParams& p, int logits_dtype, cudaStream_t s, cudaError_t* err);

// accumulate_cells.cu
int launch_cells(const mss_layout_t* lay, const AccParams& p, int logits_dtype, cudaStream_t s, cudaError_t* err);

}  // namespace mss
