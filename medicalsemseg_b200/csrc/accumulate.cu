// Weighted overlap accumulation with fused normalise / argmax
// (replaces the Python scatter loop of engine/utils.py:137-151 and, when fused, engine/test.py:140-141).
//
// Output-stationary: one thread owns 4 consecutive W voxels of the stitched volume for all classes; a
// CTA owns a tile of one D-plane (a few H rows x up to 128 W voxels).  The CTA first lists, in shared
// memory, the windows of the grid that intersect its tile (ascending window index, i.e. the order of
// engine/utils.py:146-148) with a ready-made base pointer per window, so the per-voxel inner loop is
// nothing but {coverage test, one 16-byte weight load, K 16-byte logit loads, fmul_rn + fadd_rn}.
// acc = fadd_rn(acc, fmul_rn(w, logit)) are the same two roundings as
// `output_image[idx] += importance_map * seg_prob[i]`, so sums are bit-identical to the reference's.
// No atomics: every accumulator element has exactly one writer per launch.  The accumulator is read
// only if a window of an earlier launch touched the voxel (no memset needed) and written once; a voxel
// whose last covering window is in this launch is finished on the spot (divide by the weight count, or
// argmax straight to uint8) and never travels through HBM again.
#include "acc_common.cuh"

namespace mss {

// KC classes are held in registers per pass; S ring stages (windows) are in flight per thread.
template <typename LT, int KC, int S, int MINB>
__global__ void __launch_bounds__(kAccThreads, MINB) accumulate_kernel(const __grid_constant__ AccParams p) {
    extern __shared__ __align__(16) unsigned char ring[];  // [S][128] weights (float4) then [S][KC][128] logit slots
    __shared__ const LT* s_ptr[kMaxCand];  // logits of class 0 such that ptr[gh * rw + gw] is voxel (gd, gh, gw)
    __shared__ int s_wofs[kMaxCand];       // same re-basing for the importance map
    __shared__ int s_sh[kMaxCand], s_sw[kMaxCand], s_cls[kMaxCand];
    __shared__ int s_any_now;
    constexpr int IB = Slot<LT>::kBytes;

    const Geo& g = p.g;
    const int tid = threadIdx.x;
    const int rd = g.roi[0], rh = g.roi[1], rw = g.roi[2];
    const long long R = static_cast<long long>(rd) * rh * rw;
    // 32-bit shared-window addresses of this thread's ring slots
    const unsigned ring_s = static_cast<unsigned>(__cvta_generic_to_shared(ring));
    const unsigned ring_w = ring_s + tid * 16;                                 // + stage * kWStage
    const unsigned ring_l = ring_s + S * kAccThreads * 16 + tid * IB;          // + stage * kLStage + k * kLSlot
    constexpr unsigned kWStage = kAccThreads * 16, kLSlot = kAccThreads * IB, kLStage = KC * kLSlot;
    const unsigned R_bytes = static_cast<unsigned>(R) * static_cast<unsigned>(sizeof(LT));

    // ---- tile -----------------------------------------------------------------------------------
    const int wt = blockIdx.x % p.n_wtiles, ht = blockIdx.x / p.n_wtiles;
    const int ld = p.box_lo[0] + blockIdx.y;
    const int b = p.b_lo + blockIdx.z;
    const int lh0 = p.box_lo[1] + ht * p.th;
    const int rows = min(p.th, p.box_lo[1] + p.box_n[1] - lh0);
    const int q0 = wt * p.tq;
    const int nqt = min(p.tq, p.nq - q0);
    const int lw0 = p.box_lo[2] + q0 * 4;
    const int gd = ld + g.org[0], gh0 = lh0 + g.org[1], gw0 = lw0 + g.org[2];
    const int gw_last = min(gw0 + nqt * 4, g.org[2] + g.ext[2]) - 1;  // last real voxel of the tile row

    // this volume's slice of the call's window range
    const long long vol0 = static_cast<long long>(b) * g.n_local;
    const long long n0 = p.g0 > vol0 ? p.g0 - vol0 : 0;
    const long long n1 = (p.g1 - vol0) < g.n_local ? (p.g1 - vol0) : g.n_local;

    // windows (global index ranges per axis) that can touch the tile, clipped to the ones this buffer owns
    const int cvd = g.cover[0][gd];
    const int dlo = max(cvd & 0xffff, g.wlo[0]), dhi = min(cvd >> 16, g.whi[0]);
    const int hlo = max(g.cover[1][gh0] & 0xffff, g.wlo[1]), hhi = min(g.cover[1][gh0 + rows - 1] >> 16, g.whi[1]);
    const int wlo = max(g.cover[2][gw0] & 0xffff, g.wlo[2]), whi = min(g.cover[2][gw_last] >> 16, g.whi[2]);
    const int nh = hhi - hlo, nw = whi - wlo;
    const int ncand = (dhi - dlo) * nh * nw;
    if (ncand <= 0) return;
    const int nchunks = (ncand + kMaxCand - 1) / kMaxCand;

    auto build_chunk = [&](int c0) {
        const int c = c0 + tid;
        if (tid < kMaxCand && c < ncand) {
            const int iw = wlo + c % nw;
            const int ih = hlo + (c / nw) % nh;
            const int id = dlo + c / (nw * nh);
            const long long n =
                (static_cast<long long>(id - g.wlo[0]) * g.nwl[1] + (ih - g.wlo[1])) * g.nwl[2] + (iw - g.wlo[2]);
            const long long ng = vol0 + n;
            const int cls = (ng < p.own0 || ng >= p.own1) ? kForeign : (n < n0 ? kBefore : (n >= n1 ? kAfter : kNow));
            const int sd = g.starts[0][id], sh = g.starts[1][ih], sw = g.starts[2][iw];
            const int wofs = ((gd - sd) * rh - sh) * rw - sw;
            s_sh[tid] = sh;
            s_sw[tid] = sw;
            s_cls[tid] = cls;
            s_wofs[tid] = wofs;
            const LT* ptr = nullptr;
            if (cls == kNow) {
                const long long gi = vol0 + n - p.g0;  // position inside this call's window range
                const long long bi = gi / p.sw_batch;
                const long long bj = gi - bi * p.sw_batch;
                ptr = static_cast<const LT*>(p.batch[bi]) + bj * g.K * R + wofs;
                s_any_now = 1;
            }
            s_ptr[tid] = ptr;
        }
    };

    // ---- this thread's quad ------------------------------------------------------------------------
    const int r = tid / p.tq, qi = tid - r * p.tq;
    const bool active = r < rows && qi < nqt;
    const int lh = lh0 + r, lw = lw0 + qi * 4;
    const int gh = gh0 + r, gw = gw0 + qi * 4;
    const int hw = gh * rw + gw;
    unsigned vmask = 0;  // elements of the quad that are real voxels
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (active && lw + e < g.ext[2]) vmask |= 1u << e;

    // element mask of window (sh, sw) over this quad
    auto cover_mask = [&](int sh, int sw) -> unsigned {
        if (static_cast<unsigned>(gh - sh) >= static_cast<unsigned>(rh)) return 0u;
        const int ww = gw - sw;
        unsigned m = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (static_cast<unsigned>(ww + e) < static_cast<unsigned>(rw)) m |= 1u << e;
        return m & vmask;
    };

    // ---- phase 1: who touched / touches / will touch each element; which listed windows feed this quad ------
    // fast: the window of this launch covers the whole quad -> ring pipeline (16-byte copies; four 4-byte copies when
    // the window sits at an odd W offset such as BraTS' clamped start 59: `unal`); slow: partial cover
    if (tid == 0) s_any_now = 0;
    unsigned before = 0, now = 0, after = 0;
    unsigned long long fast = 0ull, slow = 0ull, unal = 0ull;
    auto scan_chunk = [&](int cn, bool flags) {
        fast = 0ull;
        slow = 0ull;
        unal = 0ull;
        for (int c = 0; c < cn; ++c) {
            const int sw = s_sw[c];
            const unsigned m = cover_mask(s_sh[c], sw);
            const int cls = s_cls[c];
            if (flags) {
                before |= cls == kBefore ? m : 0u;
                after |= cls == kAfter ? m : 0u;
                now |= cls == kNow ? m : 0u;
            }
            if (cls == kNow && m != 0u) {
                const bool aligned = ((gw - sw) & 3) == 0;
                if (m == 0xFu && p.vec_ok && (aligned || Slot<LT>::kUnaligned)) {
                    fast |= 1ull << c;
                    if (!aligned) unal |= 1ull << c;
                } else {
                    slow |= 1ull << c;
                }
            }
        }
    };
    for (int ch = 0; ch < nchunks; ++ch) {
        __syncthreads();
        build_chunk(ch * kMaxCand);
        __syncthreads();
        scan_chunk(min(kMaxCand, ncand - ch * kMaxCand), true);
    }
    if (s_any_now == 0) return;  // uniform: nothing of this launch lands in the tile
    const unsigned complete = p.fuse != MSS_FUSE_NONE ? (now & ~after) : 0u;

    const long long plane = static_cast<long long>(g.ext[1]) * g.pitch;
    const long long cstride = static_cast<long long>(g.ext[0]) * plane;
    float* accb = p.acc != nullptr ? p.acc + static_cast<long long>(b) * g.K * cstride + static_cast<long long>(ld) * plane +
                                         static_cast<long long>(lh) * g.pitch + lw
                                   : nullptr;

    // window-weight count of the voxels finished here: ascending fp32 sum over ALL covering windows, i.e. what
    // engine/utils.py:148 accumulates into count_map.  Weights only (L1/L2-resident map), logits mode only.
    float cnt[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.fuse == MSS_FUSE_LOGITS) {
        for (int ch = 0; ch < nchunks; ++ch) {
            if (nchunks > 1) {
                __syncthreads();
                build_chunk(ch * kMaxCand);
                __syncthreads();
            }
            const int cn = min(kMaxCand, ncand - ch * kMaxCand);
            if (complete == 0u) continue;
            for (int c = 0; c < cn; ++c) {
                const unsigned m = cover_mask(s_sh[c], s_sw[c]);
                if (m == 0u) continue;
                const float* wp = p.imp + (s_wofs[c] + hw);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (m & (1u << e)) cnt[e] = __fadd_rn(cnt[e], __ldg(wp + e));
            }
        }
    }

    ArgmaxState am[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) am[e].reset();

    // ---- phase 2: KC classes at a time, S windows in flight per thread -------------------------------------------
    for (int k0 = 0; k0 < g.K; k0 += KC) {
        const int kc = min(KC, g.K - k0);
        float4 a[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (now != 0u && before != 0u && k < kc) {
                const float4 v = *reinterpret_cast<const float4*>(accb + (k0 + k) * cstride);
                a[k].x = (before & 1u) ? v.x : 0.f;
                a[k].y = (before & 2u) ? v.y : 0.f;
                a[k].z = (before & 4u) ? v.z : 0.f;
                a[k].w = (before & 8u) ? v.w : 0.f;
            }
        }
        for (int ch = 0; ch < nchunks; ++ch) {
            const int cn = min(kMaxCand, ncand - ch * kMaxCand);
            if (nchunks > 1) {  // the single-chunk list and masks of phase 1 are still in place otherwise
                __syncthreads();
                build_chunk(ch * kMaxCand);
                __syncthreads();
                scan_chunk(cn, false);
            }
            if ((fast | slow) == 0ull) continue;

            // full: all KC classes of this pass exist (no predicates in the unrolled loops)
            auto run = [&](auto full_tag) {
                constexpr bool kFull = decltype(full_tag)::value;
                auto issue = [&](int c, int st) {  // start fetching window c of the list into ring stage st
                    if (c < cn && ((fast >> c) & 1ull)) {
                        const float* wg = p.imp + (s_wofs[c] + hw);
                        const char* lg = reinterpret_cast<const char*>(s_ptr[c] + hw) + static_cast<size_t>(k0) * R_bytes;
                        const unsigned dst = ring_l + st * kLStage;
                        if (!((unal >> c) & 1ull)) {
                            cp_async16(ring_w + st * kWStage, wg);
#pragma unroll
                            for (int k = 0; k < KC; ++k)
                                if (kFull || k < kc) Slot<LT>::fetch(dst + k * kLSlot, lg + static_cast<unsigned>(k) * R_bytes);
                        } else {
                            Slot<float>::fetch_unaligned(ring_w + st * kWStage, wg);
#pragma unroll
                            for (int k = 0; k < KC; ++k)
                                if (kFull || k < kc)
                                    Slot<LT>::fetch_unaligned(dst + k * kLSlot, lg + static_cast<unsigned>(k) * R_bytes);
                        }
                    }
                    cp_async_commit();
                };
                int st_issue = 0;
#pragma unroll
                for (int c = 0; c < S - 1; ++c) {
                    issue(c, st_issue);
                    st_issue = st_issue + 1 == S ? 0 : st_issue + 1;
                }
                int st_cons = 0;
                for (int c = 0; c < cn; ++c) {
                    issue(c + S - 1, st_issue);
                    st_issue = st_issue + 1 == S ? 0 : st_issue + 1;
                    cp_async_wait<S - 1>();  // everything up to window c has landed
                    if ((fast >> c) & 1ull) {
                        const float4 w4 = lds_f4(ring_w + st_cons * kWStage);
                        const unsigned src = ring_l + st_cons * kLStage;
#pragma unroll
                        for (int k = 0; k < KC; ++k)
                            if (kFull || k < kc) {
                                const float4 l = Slot<LT>::read(src + k * kLSlot);
                                a[k].x = __fadd_rn(a[k].x, __fmul_rn(w4.x, l.x));
                                a[k].y = __fadd_rn(a[k].y, __fmul_rn(w4.y, l.y));
                                a[k].z = __fadd_rn(a[k].z, __fmul_rn(w4.z, l.z));
                                a[k].w = __fadd_rn(a[k].w, __fmul_rn(w4.w, l.w));
                            }
                    } else if ((slow >> c) & 1ull) {
                        const unsigned m = cover_mask(s_sh[c], s_sw[c]);
                        const float* wp = p.imp + (s_wofs[c] + hw);
                        const LT* lg = s_ptr[c] + hw + static_cast<long long>(k0) * R;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (!(m & (1u << e))) continue;
                            const float w1 = __ldg(wp + e);
#pragma unroll
                            for (int k = 0; k < KC; ++k)
                                if (kFull || k < kc) {
                                    float& dst = comp(a[k], e);
                                    dst = __fadd_rn(dst, __fmul_rn(w1, LogitLoad<LT>::one(lg + k * R + e)));
                                }
                        }
                    }
                    st_cons = st_cons + 1 == S ? 0 : st_cons + 1;
                }
            };
            if (kc == KC) run(std::true_type{});
            else run(std::false_type{});
            cp_async_wait<0>();
        }
        if (now == 0u) continue;
        if (p.fuse == MSS_FUSE_LOGITS && complete) {
#pragma unroll
            for (int k = 0; k < KC; ++k)
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (complete & (1u << e)) {
                        float& v = comp(a[k], e);
                        v = __fdiv_rn(v, cnt[e]);  // engine/utils.py:151
                    }
        }
        if (p.fuse == MSS_FUSE_LABELS && complete) {
            // argmax of the raw weighted sums: dividing every class by the same positive count cannot reorder them
#pragma unroll
            for (int k = 0; k < KC; ++k)
                if (k < kc) {
                    am[0].push(a[k].x, k0 + k);
                    am[1].push(a[k].y, k0 + k);
                    am[2].push(a[k].z, k0 + k);
                    am[3].push(a[k].w, k0 + k);
                }
        }
        // store: skipped only when the whole quad was finished into labels
        const bool all_to_labels = p.fuse == MSS_FUSE_LABELS && (complete & vmask) == vmask;
        if (accb != nullptr && !all_to_labels) {
#pragma unroll
            for (int k = 0; k < KC; ++k)
                if (k < kc) *reinterpret_cast<float4*>(accb + (k0 + k) * cstride) = a[k];
        }
    }

    if (p.fuse == MSS_FUSE_LABELS && complete) {
        uint8_t* lab = p.labels + (static_cast<long long>(b) * g.ext[0] + ld) * g.ext[1] * p.label_pitch +
                       static_cast<long long>(lh) * p.label_pitch + lw;
        unsigned ties = 0, packed = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (complete & (1u << e)) {
                packed |= static_cast<unsigned>(am[e].label()) << (8 * e);
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

template <typename LT, int KC, int S, int MINB>
static cudaError_t launch_one(dim3 grid, cudaStream_t s, const AccParams& p) {
    constexpr size_t smem = acc_smem_bytes<LT, KC, S>();
    // the opt-in is per device (context): set it on every launch - a process may stitch on several GPUs
    cudaError_t e = cudaFuncSetAttribute(accumulate_kernel<LT, KC, S, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    accumulate_kernel<LT, KC, S, MINB><<<grid, kAccThreads, smem, s>>>(p);
    return cudaGetLastError();
}

template <typename LT>
static cudaError_t launch_for_k(int K, dim3 grid, cudaStream_t s, const AccParams& p) {
    // classes per register pass (KC): exact divisors avoid predication, 8 with a predicated tail otherwise.
    // Ring depth S and the 4-CTA/SM register cap were picked on B200 with benchmarks/kernel_bench.py: 16 resident
    // warps/SM matter more than a deeper ring (K=14: KC=7,S=3 -> 5.6 TB/s; KC=14,S=3 -> 4.4 TB/s; KC=7,S=6 -> 2.8 TB/s).
    static const int variant = getenv("MSS_ACC_VARIANT") ? atoi(getenv("MSS_ACC_VARIANT")) : 0;  // tuning knob
    if (K % 14 == 0 && variant == 2) return launch_one<LT, 14, 3, 2>(grid, s, p);
    if (K % 7 == 0 && variant == 1) return launch_one<LT, 7, 4, 4>(grid, s, p);
    if (K % 7 == 0) return launch_one<LT, 7, 3, 4>(grid, s, p);
    if (K % 8 == 0) return launch_one<LT, 8, 3, 4>(grid, s, p);
    if (K % 5 == 0) return launch_one<LT, 5, 4, 4>(grid, s, p);
    if (K % 4 == 0) return launch_one<LT, 4, 4, 4>(grid, s, p);
    if (K % 3 == 0) return launch_one<LT, 3, 6, 4>(grid, s, p);
    if (K == 2) return launch_one<LT, 2, 8, 4>(grid, s, p);
    return launch_one<LT, 8, 3, 4>(grid, s, p);
}

}  // namespace mss

using namespace mss;

static thread_local int t_last_path = -1;  // which kernel the calling thread's last accumulate call launched

static int accumulate_impl(const mss_layout_t* lay, const void* const* batch_ptrs, int32_t n_batches, int32_t sw_batch,
                           int32_t logits_dtype, int64_t first_window, int64_t n_windows, int64_t own_first,
                           int64_t own_count, const float* importance_map, float* acc, int32_t fuse, uint8_t* labels,
                           int32_t label_pitch_w, float tie_tol, unsigned long long* near_ties, void* stream) {
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
    MSS_REQUIRE(static_cast<long long>(g.img[1]) * g.roi[2] + g.img[2] < (1LL << 31) &&
                    static_cast<long long>(g.roi[0]) * g.roi[1] * g.roi[2] < (1LL << 30),
                MSS_E_UNSUPPORTED, "accumulate: roi / image too large for 32-bit window offsets");
    if (own_count < 0) own_first = 0, own_count = total;  // the whole box
    MSS_REQUIRE(own_first >= 0 && own_count > 0 && own_first + own_count <= total, MSS_E_ARG,
                "accumulate: owned range [%lld, +%lld) outside [0, %lld)", static_cast<long long>(own_first),
                static_cast<long long>(own_count), total);
    MSS_REQUIRE(first_window >= own_first && first_window + n_windows <= own_first + own_count, MSS_E_ARG,
                "accumulate: windows [%lld, +%lld) outside the owned range [%lld, +%lld)", static_cast<long long>(first_window),
                static_cast<long long>(n_windows), static_cast<long long>(own_first), static_cast<long long>(own_count));
    p.own0 = own_first;
    p.own1 = own_first + own_count;
    const bool covers_all = first_window == 0 && n_windows == total;
    if (fuse != MSS_FUSE_NONE) {
        for (int a = 0; a < 3; ++a)
            MSS_REQUIRE(g.wlo[a] == 0 && g.whi[a] == g.ns[a], MSS_E_ARG,
                        "accumulate: fused finishing needs a buffer that owns every window (axis %d)", a);
        MSS_REQUIRE(own_first == 0 && own_count == total, MSS_E_ARG,
                    "accumulate: fused finishing needs a buffer that owns every window (owned range is partial)");
    }
    if (fuse == MSS_FUSE_LABELS) {
        MSS_REQUIRE(labels != nullptr && label_pitch_w >= g.ext[2], MSS_E_ARG, "accumulate: labels buffer / pitch invalid");
        MSS_REQUIRE(g.K <= 255, MSS_E_UNSUPPORTED, "accumulate: uint8 labels need K <= 255");
        MSS_REQUIRE(acc != nullptr || covers_all, MSS_E_ARG,
                    "accumulate: acc may be NULL only when one call covers every owned window");
    } else {
        MSS_REQUIRE(acc != nullptr, MSS_E_ARG, "accumulate: acc is null");
    }
    MSS_REQUIRE(acc == nullptr || reinterpret_cast<uintptr_t>(acc) % 16 == 0, MSS_E_ALIGN,
                "accumulate: acc must be 16-byte aligned");
    MSS_REQUIRE(logits_dtype == MSS_F32 || logits_dtype == MSS_F16 || logits_dtype == MSS_BF16, MSS_E_ARG,
                "accumulate: unknown logits dtype %d", logits_dtype);
    const int esz = logits_dtype == MSS_F32 ? 4 : 2;
    int vec_ok = (g.roi[2] % 4 == 0) && (reinterpret_cast<uintptr_t>(importance_map) % 16 == 0);
    for (int i = 0; i < n_batches; ++i) {
        MSS_REQUIRE(batch_ptrs[i] != nullptr, MSS_E_ARG, "accumulate: batch pointer %d is null", i);
        p.batch[i] = batch_ptrs[i];
        if (reinterpret_cast<uintptr_t>(batch_ptrs[i]) % (4 * esz) != 0) vec_ok = 0;
    }
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
    p.b_lo = b_lo;
    p.nq = (p.box_n[2] + 3) / 4;
    cudaStream_t s = as_stream(stream);
    // few classes, every window in this launch, labels out: the row-staged kernel (accumulate_rows.cu); otherwise the
    // cell-uniform kernel (accumulate_cells.cu) serves every geometry that fits its tables, and the general kernel the rest
    {
        cudaError_t cerr = cudaSuccess;
        if (launch_rows(lay, p, logits_dtype, s, &cerr) == 0) {
            MSS_CUDA(cerr);
            t_last_path = MSS_ACC_PATH_ROWS;
            return MSS_OK;
        }
        if (launch_cells(lay, p, logits_dtype, s, &cerr) == 0) {
            MSS_CUDA(cerr);
            t_last_path = MSS_ACC_PATH_CELLS;
            return MSS_OK;
        }
    }
    // tile: up to 32 quads (128 voxels) along W, split evenly; as many rows as fit 128 threads
    p.n_wtiles = (p.nq + 31) / 32;
    p.tq = (p.nq + p.n_wtiles - 1) / p.n_wtiles;
    p.th = kAccThreads / p.tq;
    const int n_htiles = (p.box_n[1] + p.th - 1) / p.th;
    dim3 grid(static_cast<unsigned>(p.n_wtiles * n_htiles), static_cast<unsigned>(p.box_n[0]),
              static_cast<unsigned>(b_hi - b_lo + 1));
    MSS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, MSS_E_UNSUPPORTED, "accumulate: box too large for one launch");
    cudaError_t e;
    if (logits_dtype == MSS_F32)
        e = launch_for_k<float>(g.K, grid, s, p);
    else if (logits_dtype == MSS_F16)
        e = launch_for_k<__half>(g.K, grid, s, p);
    else
        e = launch_for_k<__nv_bfloat16>(g.K, grid, s, p);
    MSS_CUDA(e);
    t_last_path = MSS_ACC_PATH_GENERAL;
    return MSS_OK;
}

extern "C" int mss_accumulate_last_path(void) { return t_last_path; }

extern "C" int mss_accumulate(const mss_layout_t* lay, const void* const* batch_ptrs, int32_t n_batches, int32_t sw_batch,
                              int32_t logits_dtype, int64_t first_window, int64_t n_windows, const float* importance_map,
                              float* acc, int32_t fuse, uint8_t* labels, int32_t label_pitch_w, float tie_tol,
                              unsigned long long* near_ties, void* stream) {
    return accumulate_impl(lay, batch_ptrs, n_batches, sw_batch, logits_dtype, first_window, n_windows, 0, -1, importance_map,
                           acc, fuse, labels, label_pitch_w, tie_tol, near_ties, stream);
}

extern "C" int mss_accumulate_range(const mss_layout_t* lay, const void* const* batch_ptrs, int32_t n_batches,
                                    int32_t sw_batch, int32_t logits_dtype, int64_t first_window, int64_t n_windows,
                                    int64_t own_first, int64_t own_count, const float* importance_map, float* acc,
                                    void* stream) {
    return accumulate_impl(lay, batch_ptrs, n_batches, sw_batch, logits_dtype, first_window, n_windows, own_first, own_count,
                           importance_map, acc, MSS_FUSE_NONE, nullptr, 0, 0.f, nullptr, stream);
}
