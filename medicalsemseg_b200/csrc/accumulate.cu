// Weighted overlap accumulation with fused normalise / argmax
// (replaces the Python scatter loop of engine/utils.py:137-151 and, when fused, engine/test.py:140-141).
//
// Output-stationary: one thread owns 4 consecutive W voxels of the stitched volume for all classes.
// It walks the windows of THIS call that cover its voxels in ascending window order (the order of
// engine/utils.py:146-148) and performs acc = fadd_rn(acc, fmul_rn(w, logit)) per window - the same
// two roundings as `output_image[idx] += importance_map * seg_prob[i]`, so sums are bit-identical to
// the reference's.  No atomics: every accumulator element has exactly one writer per launch.  The
// accumulator is read only if a window of an earlier call touched the voxel (no memset needed) and
// written once; a voxel whose last covering window is in this call is finished on the spot
// (divide by the weight count, optionally argmax to uint8) and never travels through HBM again.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "labels.cuh"

namespace mss {

constexpr int kAccThreads = 128;
constexpr int kAccTK = 8;  // classes held in registers at a time

struct AccParams {
    Geo g;
    const void* batch[MSS_MAX_BATCH_PTRS];
    int sw_batch;
    long long g0, g1;  // owned-window range of this call, over n_volumes * n_local
    const float* imp;
    float* acc;
    uint8_t* labels;
    int label_pitch;
    int fuse;
    float tie_tol;
    unsigned long long* near_ties;
    int box_lo[3];  // local box this launch covers; box_lo[2] is a multiple of 4
    int box_n[3];
    int nq;    // quads per row of the box
    int b_lo;  // first volume touched
    int vec_ok;  // logits pointers and roi allow 16-byte (8-byte for 16-bit logits) vector loads
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

template <typename LT>
__global__ void __launch_bounds__(kAccThreads) accumulate_kernel(const __grid_constant__ AccParams p) {
    const Geo& g = p.g;
    const int t = blockIdx.x * kAccThreads + threadIdx.x;
    if (t >= p.nq * p.box_n[1]) return;
    const int q = t % p.nq;
    const int ld = p.box_lo[0] + blockIdx.y;
    const int lh = p.box_lo[1] + t / p.nq;
    const int lw = p.box_lo[2] + q * 4;
    const int b = p.b_lo + blockIdx.z;
    const int gd = ld + g.org[0], gh = lh + g.org[1], gw = lw + g.org[2];
    const int rd = g.roi[0], rh = g.roi[1], rw = g.roi[2];
    const long long R = static_cast<long long>(rd) * rh * rw;

    // this volume's slice of the call's window range
    const long long vol0 = static_cast<long long>(b) * g.n_local;
    const long long n0 = p.g0 > vol0 ? p.g0 - vol0 : 0;
    const long long n1 = (p.g1 - vol0) < g.n_local ? (p.g1 - vol0) : g.n_local;

    // cover ranges (global window indices per axis), W per element
    const int cvd = g.cover[0][gd], cvh = g.cover[1][gh];
    const int dlo = cvd & 0xffff, dhi = cvd >> 16, hlo = cvh & 0xffff, hhi = cvh >> 16;
    bool valid[4];
    int wlo = 0x7fffffff, whi = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        valid[e] = lw + e < g.ext[2];
        if (valid[e]) {
            const int c = g.cover[2][gw + e];
            wlo = min(wlo, c & 0xffff);
            whi = max(whi, c >> 16);
        }
    }
    if (!valid[0]) return;
    // clipped to the windows this buffer owns
    const int odlo = max(dlo, g.wlo[0]), odhi = min(dhi, g.whi[0]);
    const int ohlo = max(hlo, g.wlo[1]), ohhi = min(hhi, g.whi[1]);
    const int owlo = max(wlo, g.wlo[2]), owhi = min(whi, g.whi[2]);

    // pass 1 (integers only): which elements were touched before / are touched now / will be touched later
    unsigned before = 0, now = 0, after = 0;
    for (int id = odlo; id < odhi; ++id)
        for (int ih = ohlo; ih < ohhi; ++ih)
            for (int iw = owlo; iw < owhi; ++iw) {
                const long long n =
                    (static_cast<long long>(id - g.wlo[0]) * g.nwl[1] + (ih - g.wlo[1])) * g.nwl[2] + (iw - g.wlo[2]);
                const int ww = gw - g.starts[2][iw];
                unsigned m = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (valid[e] && ww + e >= 0 && ww + e < rw) m |= 1u << e;
                if (n < n0) before |= m;
                else if (n >= n1) after |= m;
                else now |= m;
            }
    if (now == 0) return;
    const unsigned complete = p.fuse != MSS_FUSE_NONE ? (now & ~after) : 0u;

    // window-weight count of finished voxels: ascending fp32 sum over ALL covering windows of the grid
    // (engine/utils.py:148 accumulates the same weights in the same order into count_map)
    float cnt[4] = {0.f, 0.f, 0.f, 0.f};
    if (complete) {
        for (int id = dlo; id < dhi; ++id)
            for (int ih = hlo; ih < hhi; ++ih) {
                const float* row = p.imp + (static_cast<long long>(gd - g.starts[0][id]) * rh + (gh - g.starts[1][ih])) * rw;
                for (int iw = wlo; iw < whi; ++iw) {
                    const int ww = gw - g.starts[2][iw];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (valid[e] && ww + e >= 0 && ww + e < rw) cnt[e] = __fadd_rn(cnt[e], __ldg(row + ww + e));
                }
            }
    }

    const long long plane = static_cast<long long>(g.ext[1]) * g.pitch;
    const long long vox = static_cast<long long>(ld) * plane + static_cast<long long>(lh) * g.pitch + lw;
    float* accb = p.acc != nullptr ? p.acc + static_cast<long long>(b) * g.K * g.ext[0] * plane + vox : nullptr;
    const long long cstride = static_cast<long long>(g.ext[0]) * plane;

    ArgmaxState am[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) am[e].reset();

    for (int k0 = 0; k0 < g.K; k0 += kAccTK) {
        float4 a[kAccTK];
#pragma unroll
        for (int k = 0; k < kAccTK; ++k) {
            a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (before && k0 + k < g.K) {
                const float4 v = *reinterpret_cast<const float4*>(accb + (k0 + k) * cstride);
                a[k].x = (before & 1u) ? v.x : 0.f;
                a[k].y = (before & 2u) ? v.y : 0.f;
                a[k].z = (before & 4u) ? v.z : 0.f;
                a[k].w = (before & 8u) ? v.w : 0.f;
            }
        }
        for (int id = odlo; id < odhi; ++id)
            for (int ih = ohlo; ih < ohhi; ++ih)
                for (int iw = owlo; iw < owhi; ++iw) {
                    const long long n =
                        (static_cast<long long>(id - g.wlo[0]) * g.nwl[1] + (ih - g.wlo[1])) * g.nwl[2] + (iw - g.wlo[2]);
                    if (n < n0 || n >= n1) continue;
                    const int ww = gw - g.starts[2][iw];
                    if (ww <= -4 || ww >= rw) continue;
                    const long long gi = vol0 + n - p.g0;  // position inside this call's window range
                    const int bi = static_cast<int>(gi / p.sw_batch);
                    const int bj = static_cast<int>(gi - static_cast<long long>(bi) * p.sw_batch);
                    const long long off = (static_cast<long long>(gd - g.starts[0][id]) * rh + (gh - g.starts[1][ih])) * rw + ww;
                    const LT* lg = static_cast<const LT*>(p.batch[bi]) + static_cast<long long>(bj) * g.K * R + off;
                    const float* wp = p.imp + off;
                    if (p.vec_ok && ww >= 0 && ww + 3 < rw && (ww & 3) == 0 && valid[3]) {
                        const float4 w4 = ldg_f4(wp);
                        float4 l[kAccTK];
#pragma unroll
                        for (int k = 0; k < kAccTK; ++k)
                            if (k0 + k < g.K) l[k] = LogitLoad<LT>::quad(lg + (k0 + k) * R);
#pragma unroll
                        for (int k = 0; k < kAccTK; ++k)
                            if (k0 + k < g.K) {
                                a[k].x = __fadd_rn(a[k].x, __fmul_rn(w4.x, l[k].x));
                                a[k].y = __fadd_rn(a[k].y, __fmul_rn(w4.y, l[k].y));
                                a[k].z = __fadd_rn(a[k].z, __fmul_rn(w4.z, l[k].z));
                                a[k].w = __fadd_rn(a[k].w, __fmul_rn(w4.w, l[k].w));
                            }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (!(valid[e] && ww + e >= 0 && ww + e < rw)) continue;
                            const float w1 = __ldg(wp + e);
#pragma unroll
                            for (int k = 0; k < kAccTK; ++k)
                                if (k0 + k < g.K) {
                                    float& dst = comp(a[k], e);
                                    dst = __fadd_rn(dst, __fmul_rn(w1, LogitLoad<LT>::one(lg + (k0 + k) * R + e)));
                                }
                        }
                    }
                }
        if (complete) {
#pragma unroll
            for (int k = 0; k < kAccTK; ++k)
                if (k0 + k < g.K) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (complete & (1u << e)) {
                            float& v = comp(a[k], e);
                            v = __fdiv_rn(v, cnt[e]);  // engine/utils.py:151
                            if (p.fuse == MSS_FUSE_LABELS) am[e].push(v, k0 + k);
                        }
                }
        }
        // store: skipped only when the whole quad was finished into labels
        const unsigned live = (valid[0] ? 1u : 0u) | (valid[1] ? 2u : 0u) | (valid[2] ? 4u : 0u) | (valid[3] ? 8u : 0u);
        const bool all_to_labels = p.fuse == MSS_FUSE_LABELS && (complete & live) == live;
        if (accb != nullptr && !all_to_labels) {
#pragma unroll
            for (int k = 0; k < kAccTK; ++k)
                if (k0 + k < g.K) *reinterpret_cast<float4*>(accb + (k0 + k) * cstride) = a[k];
        }
    }

    if (p.fuse == MSS_FUSE_LABELS && complete) {
        uint8_t* lab = p.labels + (static_cast<long long>(b) * g.ext[0] + ld) * g.ext[1] * p.label_pitch +
                       static_cast<long long>(lh) * p.label_pitch + lw;
        unsigned ties = 0;
        unsigned packed = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (complete & (1u << e)) {
                const int lbl = am[e].label();
                packed |= static_cast<unsigned>(lbl) << (8 * e);
                ties += am[e].near_tie(p.tie_tol) ? 1u : 0u;
            }
        if (complete == 0xFu && ((reinterpret_cast<uintptr_t>(lab) & 3u) == 0)) {
            *reinterpret_cast<unsigned*>(lab) = packed;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (complete & (1u << e)) lab[e] = static_cast<uint8_t>(packed >> (8 * e));
        }
        if (ties && p.near_ties != nullptr) atomicAdd(p.near_ties, static_cast<unsigned long long>(ties));
    }
}

// bounding box (local buffer coordinates) of the owned windows [n0, n1) of one volume
static void window_range_box(const mss_layout_t* lay, long long n0, long long n1, int lo[3], int hi[3]) {
    const int32_t* t = lay->table_host;
    const int32_t* st[3] = {t + t[kHdrOffStarts], t + t[kHdrOffStarts + 1], t + t[kHdrOffStarts + 2]};
    const int nwl[3] = {lay->win_hi[0] - lay->win_lo[0], lay->win_hi[1] - lay->win_lo[1], lay->win_hi[2] - lay->win_lo[2]};
    int i0[3], i1[3];  // inclusive owned-window index ranges per axis
    const long long hw = static_cast<long long>(nwl[1]) * nwl[2];
    i0[0] = static_cast<int>(n0 / hw);
    i1[0] = static_cast<int>((n1 - 1) / hw);
    if (i0[0] == i1[0]) {
        i0[1] = static_cast<int>((n0 / nwl[2]) % nwl[1]);
        i1[1] = static_cast<int>(((n1 - 1) / nwl[2]) % nwl[1]);
        if (i0[1] == i1[1]) {
            i0[2] = static_cast<int>(n0 % nwl[2]);
            i1[2] = static_cast<int>((n1 - 1) % nwl[2]);
        } else {
            i0[2] = 0;
            i1[2] = nwl[2] - 1;
        }
    } else {
        i0[1] = 0;
        i1[1] = nwl[1] - 1;
        i0[2] = 0;
        i1[2] = nwl[2] - 1;
    }
    for (int a = 0; a < 3; ++a) {
        lo[a] = st[a][lay->win_lo[a] + i0[a]] - lay->origin[a];
        hi[a] = st[a][lay->win_lo[a] + i1[a]] + lay->roi[a] - lay->origin[a];
    }
}

}  // namespace mss

using namespace mss;

extern "C" int mss_accumulate(const mss_layout_t* lay, const void* const* batch_ptrs, int32_t n_batches, int32_t sw_batch,
                              int32_t logits_dtype, int64_t first_window, int64_t n_windows, const float* importance_map,
                              float* acc, int32_t fuse, uint8_t* labels, int32_t label_pitch_w, float tie_tol,
                              unsigned long long* near_ties, void* stream) {
    AccParams p;
    int rc = make_geo(lay, &p.g);
    if (rc != MSS_OK) return rc;
    const Geo& g = p.g;
    MSS_REQUIRE(batch_ptrs != nullptr && importance_map != nullptr, MSS_E_ARG, "accumulate: null argument");
    MSS_REQUIRE(n_batches > 0 && n_batches <= MSS_MAX_BATCH_PTRS, MSS_E_ARG, "accumulate: n_batches %d outside [1, %d]",
                n_batches, MSS_MAX_BATCH_PTRS);
    MSS_REQUIRE(sw_batch > 0 && n_windows > 0, MSS_E_ARG, "accumulate: need sw_batch > 0 and n_windows > 0");
    MSS_REQUIRE(n_windows <= static_cast<int64_t>(n_batches) * sw_batch &&
                    n_windows > static_cast<int64_t>(n_batches - 1) * sw_batch,
                MSS_E_ARG, "accumulate: %lld windows do not fill %d batches of %d", static_cast<long long>(n_windows),
                n_batches, sw_batch);
    const long long total = g.n_local * g.nb;
    MSS_REQUIRE(first_window >= 0 && first_window + n_windows <= total, MSS_E_ARG,
                "accumulate: windows [%lld, +%lld) outside [0, %lld)", static_cast<long long>(first_window),
                static_cast<long long>(n_windows), total);
    MSS_REQUIRE(fuse == MSS_FUSE_NONE || fuse == MSS_FUSE_LOGITS || fuse == MSS_FUSE_LABELS, MSS_E_ARG,
                "accumulate: unknown fuse mode %d", fuse);
    MSS_REQUIRE(g.pitch % 4 == 0, MSS_E_ALIGN, "accumulate: pitch_w (%d) must be a multiple of 4", g.pitch);
    const bool covers_all = first_window == 0 && n_windows == total;
    if (fuse == MSS_FUSE_LABELS) {
        MSS_REQUIRE(labels != nullptr && label_pitch_w >= g.ext[2], MSS_E_ARG, "accumulate: labels buffer / pitch invalid");
        MSS_REQUIRE(g.K <= 255, MSS_E_UNSUPPORTED, "accumulate: uint8 labels need K <= 255");
        MSS_REQUIRE(acc != nullptr || covers_all, MSS_E_ARG,
                    "accumulate: acc may be NULL only when one call covers every owned window");
        for (int a = 0; a < 3; ++a)
            MSS_REQUIRE(g.wlo[a] == 0 && g.whi[a] == g.ns[a], MSS_E_ARG,
                        "accumulate: fused finishing needs a buffer that owns every window (axis %d)", a);
    } else {
        MSS_REQUIRE(acc != nullptr, MSS_E_ARG, "accumulate: acc is null");
        if (fuse == MSS_FUSE_LOGITS)
            for (int a = 0; a < 3; ++a)
                MSS_REQUIRE(g.wlo[a] == 0 && g.whi[a] == g.ns[a], MSS_E_ARG,
                            "accumulate: fused finishing needs a buffer that owns every window (axis %d)", a);
    }
    MSS_REQUIRE(acc == nullptr || reinterpret_cast<uintptr_t>(acc) % 16 == 0, MSS_E_ALIGN,
                "accumulate: acc must be 16-byte aligned");
    MSS_REQUIRE(logits_dtype == MSS_F32 || logits_dtype == MSS_F16 || logits_dtype == MSS_BF16, MSS_E_ARG,
                "accumulate: unknown logits dtype %d", logits_dtype);
    const int esz = logits_dtype == MSS_F32 ? 4 : 2;
    const long long R = static_cast<long long>(g.roi[0]) * g.roi[1] * g.roi[2];
    int vec_ok = (g.roi[2] % 4 == 0) && (reinterpret_cast<uintptr_t>(importance_map) % 16 == 0);
    for (int i = 0; i < n_batches; ++i) {
        MSS_REQUIRE(batch_ptrs[i] != nullptr, MSS_E_ARG, "accumulate: batch pointer %d is null", i);
        p.batch[i] = batch_ptrs[i];
        if (reinterpret_cast<uintptr_t>(batch_ptrs[i]) % (4 * esz) != 0) vec_ok = 0;
    }
    (void)R;
    p.sw_batch = sw_batch;
    p.g0 = first_window;
    p.g1 = first_window + n_windows;
    p.imp = importance_map;
    p.acc = acc;
    p.labels = labels;
    p.label_pitch = label_pitch_w;
    p.fuse = fuse;
    p.tie_tol = tie_tol;
    p.near_ties = near_ties;
    p.vec_ok = vec_ok;

    // union of the per-volume bounding boxes of the windows in [g0, g1)
    const int b_lo = static_cast<int>(p.g0 / g.n_local);
    const int b_hi = static_cast<int>((p.g1 - 1) / g.n_local);
    int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {0, 0, 0};
    for (int b = b_lo; b <= b_hi; ++b) {
        const long long vol0 = static_cast<long long>(b) * g.n_local;
        const long long n0 = p.g0 > vol0 ? p.g0 - vol0 : 0;
        const long long n1 = (p.g1 - vol0) < g.n_local ? (p.g1 - vol0) : g.n_local;
        int l[3], h[3];
        window_range_box(lay, n0, n1, l, h);
        for (int a = 0; a < 3; ++a) {
            lo[a] = l[a] < lo[a] ? l[a] : lo[a];
            hi[a] = h[a] > hi[a] ? h[a] : hi[a];
        }
    }
    lo[2] &= ~3;
    for (int a = 0; a < 3; ++a) {
        p.box_lo[a] = lo[a];
        p.box_n[a] = hi[a] - lo[a];
    }
    p.nq = (p.box_n[2] + 3) / 4;
    p.b_lo = b_lo;
    const long long per_plane = static_cast<long long>(p.nq) * p.box_n[1];
    dim3 grid(static_cast<unsigned>((per_plane + kAccThreads - 1) / kAccThreads), static_cast<unsigned>(p.box_n[0]),
              static_cast<unsigned>(b_hi - b_lo + 1));
    MSS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, MSS_E_UNSUPPORTED, "accumulate: box too large for one launch");
    cudaStream_t s = as_stream(stream);
    if (logits_dtype == MSS_F32)
        accumulate_kernel<float><<<grid, kAccThreads, 0, s>>>(p);
    else if (logits_dtype == MSS_F16)
        accumulate_kernel<__half><<<grid, kAccThreads, 0, s>>>(p);
    else
        accumulate_kernel<__nv_bfloat16><<<grid, kAccThreads, 0, s>>>(p);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}
