// Cell-uniform weighted overlap accumulation (engine/utils.py:137-151 [+ engine/test.py:140-141 when fused]).
//
// The window starts and ends of an axis cut it into SEGMENTS inside which the set of covering windows is constant;
// the product of one segment per axis is a CELL, and every voxel of a cell is covered by exactly the same windows,
// each of them completely.  A CTA works inside one cell, so everything the general kernel (accumulate.cu) decides per
// quad and per window - coverage masks, before / now / after, which windows to fetch - is decided ONCE per CTA and W
// segment (one warp each, from the parameter block alone) and is the same for every quad of the segment: the
// per-voxel loop is {K + 1 asynchronous 16-byte copies into the thread's ring, K + 1 shared-memory reads,
// 4 K fmul_rn + 4 K fadd_rn} per window and nothing else.  A CTA takes whole rows of a column of W segments (so the
// pieces of a logits row are fetched together and meet in L2) for several rows and planes, which amortises the
// window listing; a thread's copies form ONE stream across its quads, S - 1 windows in flight.
// Arithmetic, window order and outputs are those of the general kernel (bit-identical sums).
//
// A quad column that straddles a W segment boundary (a clamped window start such as BraTS' 59 that is not a multiple of 4)
// or holds the ragged end of a row is a tile of its own: still uniform, but its windows carry an element mask and are
// fetched element-wise.  Geometries that exceed the tables (more than 32 windows over one voxel, more than 64 segments
// per axis) and 16-bit logits with unaligned windows run on the general kernel instead.
#include "acc_common.cuh"

namespace mss {

constexpr int kCellMaxCand = 32;   // windows covering a cell (3 x 3 x 3 = 27 at overlap <= 0.5 with clamped last windows)
constexpr int kCellMaxTiles = 8;   // W tiles (segments / straddling quad columns) one CTA column spans
constexpr int kCellMaxColQuads = 128;

struct CellParams {
    AccParams a;
    int starts[3][kCellMaxSeg];       // window starts per axis (global coordinates), indexed by window
    int seg_lo[2][kCellMaxSeg];       // axis 0 (D), 1 (H): first local coordinate of every segment,
    int seg_n[2][kCellMaxSeg];        //                    planes / rows in it,
    int seg_w0[2][kCellMaxSeg];       //                    owned windows [w0, w1) covering it
    int seg_w1[2][kCellMaxSeg];
    int n_seg[2];
    int wt_q0[kCellMaxSeg];           // W tiles: first local quad; a tile lies inside one W segment, or is ONE quad column
    int wt_nq[kCellMaxSeg];           //          quads                      that straddles segments / the ragged row end,
    int wt_w0[kCellMaxSeg];           //          owned windows [w0, w1) covering it
    int wt_w1[kCellMaxSeg];
    int col_t0[kCellMaxSeg];          // CTA columns: first W tile, tiles, rows per item
    int col_nt[kCellMaxSeg];
    int col_hr[kCellMaxSeg];
    int x_prefix[kCellMaxSeg + 1];    // blockIdx.x -> column: items (row chunks) before column i
    int y_prefix[kCellMaxSeg + 1];    // blockIdx.y -> D segment: plane groups before segment i
    int n_col;
    int pd;                           // planes of an item
};

// flags of a listed window: bits 0-3 = elements of the tile's quads it covers, bit 4 = fetch element-wise (4-byte copies:
// the window covers the quads partly, or starts off the 16-byte lattice such as BraTS' clamped 59)
constexpr int kCellScalar = 16;

template <typename LT, int KC, int S, int MINB, bool FULL, int QD>
__global__ void __launch_bounds__(kAccThreads, MINB) accumulate_cells_kernel(const __grid_constant__ CellParams cp) {
    extern __shared__ __align__(16) unsigned char ring[];  // [S * QD][128] weights (float4) then [S * QD][KC][128] logit slots
    // per W tile of this CTA's column: the windows of THIS launch covering its cells, ascending window index ...
    __shared__ const LT* s_base[kCellMaxTiles][kCellMaxCand];  // class-0 logits of the window
    __shared__ int s_wc[kCellMaxTiles][kCellMaxCand];          // (sd * rh + sh) * rw + sw
    __shared__ unsigned char s_flag[kCellMaxTiles][kCellMaxCand];  // element mask | kCellScalar
    __shared__ int s_allwc[kCellMaxTiles][kCellMaxCand];       // ... and ALL covering windows (weight count)
    __shared__ unsigned char s_allm[kCellMaxTiles][kCellMaxCand];
    __shared__ int s_info[kCellMaxTiles][5];                   // n_now, n_all, element masks of before / after / now
    __shared__ unsigned char s_qtile[kCellMaxColQuads];        // quad of the column -> tile
    constexpr int IB = Slot<LT>::kBytes;
    const AccParams& p = cp.a;
    const Geo& g = p.g;
    const int tid = threadIdx.x;
    const int K = g.K;
    const int rh = g.roi[1], rw = g.roi[2];
    const long long R = static_cast<long long>(g.roi[0]) * rh * rw;
    const unsigned ring_s = static_cast<unsigned>(__cvta_generic_to_shared(ring));
    const unsigned ring_w = ring_s + tid * 16;
    const unsigned ring_l = ring_s + S * QD * kAccThreads * 16 + tid * IB;
    constexpr unsigned kWStage = kAccThreads * 16, kLSlot = kAccThreads * IB, kLStage = KC * kLSlot;
    const unsigned R_bytes = static_cast<unsigned>(R) * static_cast<unsigned>(sizeof(LT));

    // ---- which item: (column of W tiles, chunk of rows of one H segment) from blockIdx.x, planes of one D segment from
    // blockIdx.y.  A CTA takes whole rows of its column, so the pieces of a logits row are fetched together
    int x = blockIdx.x, col = 0;
    while (x >= cp.x_prefix[col + 1]) ++col;
    x -= cp.x_prefix[col];
    const int hr = cp.col_hr[col];
    int hs = 0;
    for (;; ++hs) {
        const int nch = (cp.seg_n[1][hs] + hr - 1) / hr;
        if (x < nch) break;
        x -= nch;
    }
    const int row0 = cp.seg_lo[1][hs] + x * hr;
    const int nrows = min(hr, cp.seg_lo[1][hs] + cp.seg_n[1][hs] - row0);
    int y = blockIdx.y, ds = 0;
    while (y >= cp.y_prefix[ds + 1]) ++ds;
    const int pl0 = cp.seg_lo[0][ds] + (y - cp.y_prefix[ds]) * cp.pd;
    const int npl = min(cp.pd, cp.seg_lo[0][ds] + cp.seg_n[0][ds] - pl0);
    const int b = p.b_lo + blockIdx.z;
    const int t0 = cp.col_t0[col], nt = cp.col_nt[col];
    const int cq0 = cp.wt_q0[t0];                                        // first quad of the column
    const int nqc = cp.wt_q0[t0 + nt - 1] + cp.wt_nq[t0 + nt - 1] - cq0;  // quads per row of the column
    const int gw_end = g.org[2] + g.ext[2] - 1;                          // last real voxel of a buffer row

    // ---- window lists, one warp per tile (everything comes from the parameter block: no global loads) ---------------
    const long long vol0 = static_cast<long long>(b) * g.n_local;
    const long long n0 = p.g0 > vol0 ? p.g0 - vol0 : 0;
    const long long n1 = (p.g1 - vol0) < g.n_local ? (p.g1 - vol0) : g.n_local;
    for (int ti = tid >> 5; ti < nt; ti += kAccThreads / 32) {
        const int lane = tid & 31;
        const int wt = t0 + ti;
        const int dlo = cp.seg_w0[0][ds], dhi = cp.seg_w1[0][ds];
        const int hlo = cp.seg_w0[1][hs], hhi = cp.seg_w1[1][hs];
        const int wlo = cp.wt_w0[wt], whi = cp.wt_w1[wt];
        const int nh = hhi - hlo, nw = whi - wlo;
        const int ncand = (dhi - dlo) * nh * nw;  // <= kCellMaxCand (checked by the host)
        const int gw_a = cp.wt_q0[wt] * 4 + g.org[2];
        for (int qq = lane; qq < cp.wt_nq[wt]; qq += 32) s_qtile[cp.wt_q0[wt] - cq0 + qq] = static_cast<unsigned char>(ti);
        int cls = kForeign, wc = 0, flag = 0;
        const LT* base = nullptr;
        if (lane < ncand) {
            const int iw = wlo + lane % nw, ih = hlo + (lane / nw) % nh, id = dlo + lane / (nw * nh);
            const long long n =
                (static_cast<long long>(id - g.wlo[0]) * g.nwl[1] + (ih - g.wlo[1])) * g.nwl[2] + (iw - g.wlo[2]);
            const long long ng = vol0 + n;
            cls = (ng < p.own0 || ng >= p.own1) ? kForeign : (n < n0 ? kBefore : (n >= n1 ? kAfter : kNow));
            const int sd = cp.starts[0][id], sh = cp.starts[1][ih], sw = cp.starts[2][iw];
            wc = (sd * rh + sh) * rw + sw;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (static_cast<unsigned>(gw_a + e - sw) < static_cast<unsigned>(rw) && gw_a + e <= gw_end) flag |= 1 << e;
            if (flag != 0xF || ((gw_a - sw) & 3) != 0) flag |= kCellScalar;
            if (cls == kNow) {
                const long long gi = ng - p.g0;  // position inside this call's window range
                const long long bi = gi / p.sw_batch;
                base = static_cast<const LT*>(p.batch[bi]) + (gi - bi * p.sw_batch) * K * R;
            }
        }
        const int m = flag & 0xF;
        const unsigned now_b = __ballot_sync(0xffffffffu, cls == kNow);
        const unsigned all_b = __ballot_sync(0xffffffffu, cls != kForeign);
        const int bef_m = __reduce_or_sync(0xffffffffu, cls == kBefore ? m : 0);
        const int aft_m = __reduce_or_sync(0xffffffffu, cls == kAfter ? m : 0);
        const int now_m = __reduce_or_sync(0xffffffffu, cls == kNow ? m : 0);
        const unsigned lt = (1u << lane) - 1u;
        if (cls == kNow) {
            const int j = __popc(now_b & lt);
            s_base[ti][j] = base;
            s_wc[ti][j] = wc;
            s_flag[ti][j] = static_cast<unsigned char>(flag);
        }
        if (cls != kForeign) {
            const int j = __popc(all_b & lt);
            s_allwc[ti][j] = wc;
            s_allm[ti][j] = static_cast<unsigned char>(m);
        }
        if (lane == 0) {
            s_info[ti][0] = __popc(now_b);
            s_info[ti][1] = __popc(all_b);
            s_info[ti][2] = bef_m;
            s_info[ti][3] = aft_m;
            s_info[ti][4] = now_m;
        }
    }
    __syncthreads();

    // ---- this thread's quads: positions tid, tid + 128, ... of the item's rows x column quads, QD planes at a time ---
    const int npos = nrows * nqc;
    if (tid >= npos) return;
    const int step_r = kAccThreads / nqc, step_q = kAccThreads - step_r * nqc;
    const long long plane = static_cast<long long>(g.ext[1]) * g.pitch;
    const long long cstride = static_cast<long long>(g.ext[0]) * plane;
    const int gd0 = pl0 + g.org[0], gh0 = row0 + g.org[1], gwc = cq0 * 4 + g.org[2];
    const int r_first = tid / nqc, q_first = tid - r_first * nqc;
    const int dplane = rh * rw;  // window-relative offset of the next plane
    unsigned ties = 0;

    // ONE stream of (plane group, quad, class pass, window) fetches per thread, S - 1 of them in flight: the copies of
    // the next quad are already under way while this one is finished.  With QD = 2 (few classes) a fetch brings the
    // window's data of the SAME quad in two consecutive planes - same tile, same windows, same masks - so the
    // bookkeeping of an iteration is shared by twice the bytes.  A quad whose tile has no window of this launch is
    // skipped by both cursors alike.
    int i_pl = 0, i_r = r_first, i_q = q_first, i_pos = tid, i_k0 = 0, i_j = 0, st_issue = 0;
    int i_t = s_qtile[i_q];
    bool i_done = false;
    auto i_advance = [&]() {
        i_pos += kAccThreads;
        i_r += step_r;
        i_q += step_q;
        if (i_q >= nqc) i_q -= nqc, ++i_r;
        if (i_pos >= npos) {
            i_pos = tid, i_r = r_first, i_q = q_first;
            i_pl += QD;
            if (i_pl >= npl) i_done = true;
        }
        i_t = s_qtile[i_q];
    };
    while (!i_done && s_info[i_t][0] == 0) i_advance();
    auto issue_next = [&]() {
        if (!i_done) {
            const int tofs = (((gd0 + i_pl) * rh) + (gh0 + i_r)) * rw + gwc + i_q * 4;
            const int o = tofs - s_wc[i_t][i_j];
            const int flag = s_flag[i_t][i_j];
            const int kc = FULL ? KC : min(KC, K - i_k0);
            const LT* lbase = s_base[i_t][i_j];
#pragma unroll
            for (int d = 0; d < QD; ++d) {
                if (d > 0 && i_pl + d >= npl) break;
                const float* wg = p.imp + (o + d * dplane);
                const char* lg = reinterpret_cast<const char*>(lbase + (o + d * dplane)) + static_cast<size_t>(i_k0) * R_bytes;
                const unsigned dw = ring_w + (st_issue * QD + d) * kWStage, dl = ring_l + (st_issue * QD + d) * kLStage;
                if (!(flag & kCellScalar)) {
                    cp_async16(dw, wg);
#pragma unroll
                    for (int k = 0; k < KC; ++k)
                        if (FULL || k < kc) Slot<LT>::fetch(dl + k * kLSlot, lg + static_cast<unsigned>(k) * R_bytes);
                } else if (Slot<LT>::kUnaligned) {  // element-wise (16-bit logits: the host never lists such windows)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (flag & (1 << e)) {
                            cp_async4(dw + 4 * e, wg + e);
#pragma unroll
                            for (int k = 0; k < KC; ++k)
                                if (FULL || k < kc)
                                    cp_async4(dl + k * kLSlot + 4 * e, lg + static_cast<unsigned>(k) * R_bytes + 4 * e);
                        }
                }
            }
            if (++i_j == s_info[i_t][0]) {
                i_j = 0;
                i_k0 += KC;
                if (i_k0 >= K) {
                    i_k0 = 0;
                    i_advance();
                    while (!i_done && s_info[i_t][0] == 0) i_advance();
                }
            }
        }
        cp_async_commit();  // one group per call, possibly empty
        st_issue = st_issue + 1 == S ? 0 : st_issue + 1;
    };
#pragma unroll
    for (int i = 0; i < S - 1; ++i) issue_next();
    int st_cons = 0;

    for (int pl = 0; pl < npl; pl += QD) {
        const int nd = min(QD, npl - pl);  // planes of this group that exist
        int rr = r_first, qq = q_first;
        for (int pos = tid; pos < npos; pos += kAccThreads, rr += step_r, qq += step_q) {
            if (qq >= nqc) qq -= nqc, ++rr;
            const int ti = s_qtile[qq];
            const int n_now = s_info[ti][0];
            if (n_now == 0) continue;  // nothing of this launch lands in this cell
            const int n_all = s_info[ti][1];
            const unsigned before = s_info[ti][2], now = s_info[ti][4];
            const unsigned complete = p.fuse != MSS_FUSE_NONE ? (now & ~static_cast<unsigned>(s_info[ti][3])) : 0u;
            const int lh = row0 + rr, lw = (cq0 + qq) * 4;
            const int gw = lw + g.org[2];
            unsigned vmask = 0;  // elements of the quad that are real voxels
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (gw + e <= gw_end) vmask |= 1u << e;
            const bool all_to_labels = p.fuse == MSS_FUSE_LABELS && (complete & vmask) == vmask;
            const int tofs = (((gd0 + pl) * rh) + (gh0 + rr)) * rw + gw;  // this quad (first plane) in window-relative units
            float* accb = p.acc != nullptr ? p.acc + static_cast<long long>(b) * K * cstride +
                                                 static_cast<long long>(pl0 + pl) * plane + static_cast<long long>(lh) * g.pitch + lw
                                           : nullptr;
            // weight count of the voxels finished here (logits mode): ascending fp32 sum over ALL covering windows,
            // i.e. what engine/utils.py:148 accumulates into count_map
            float cnt[QD][4];
#pragma unroll
            for (int d = 0; d < QD; ++d)
#pragma unroll
                for (int e = 0; e < 4; ++e) cnt[d][e] = 0.f;
            if (p.fuse == MSS_FUSE_LOGITS && complete) {
                for (int j = 0; j < n_all; ++j) {
                    const int m = s_allm[ti][j];
#pragma unroll
                    for (int d = 0; d < QD; ++d) {
                        if (d >= nd) break;
                        const int o = tofs + d * dplane - s_allwc[ti][j];
                        const float* wp = p.imp + o;
                        if (m == 0xF && (o & 3) == 0) {
                            const float4 w4 = ldg_f4(wp);
                            cnt[d][0] = __fadd_rn(cnt[d][0], w4.x);
                            cnt[d][1] = __fadd_rn(cnt[d][1], w4.y);
                            cnt[d][2] = __fadd_rn(cnt[d][2], w4.z);
                            cnt[d][3] = __fadd_rn(cnt[d][3], w4.w);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (m & (1 << e)) cnt[d][e] = __fadd_rn(cnt[d][e], __ldg(wp + e));
                        }
                    }
                }
            }
            ArgmaxState am[QD][4];
#pragma unroll
            for (int d = 0; d < QD; ++d)
#pragma unroll
                for (int e = 0; e < 4; ++e) am[d][e].reset();

            for (int k0 = 0; k0 < K; k0 += KC) {
                const int kc = FULL ? KC : min(KC, K - k0);
                float4 a[QD][KC];
#pragma unroll
                for (int d = 0; d < QD; ++d)
#pragma unroll
                    for (int k = 0; k < KC; ++k) {
                        a[d][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (before != 0u && d < nd && (FULL || k < kc)) {
                            const float4 v = *reinterpret_cast<const float4*>(accb + d * plane + (k0 + k) * cstride);
                            a[d][k].x = (before & 1u) ? v.x : 0.f;
                            a[d][k].y = (before & 2u) ? v.y : 0.f;
                            a[d][k].z = (before & 4u) ? v.z : 0.f;
                            a[d][k].w = (before & 8u) ? v.w : 0.f;
                        }
                    }
                for (int j = 0; j < n_now; ++j) {
                    issue_next();
                    cp_async_wait<S - 1>();  // the group of this window (and every earlier one) has landed
                    const int m = s_flag[ti][j] & 0xF;
#pragma unroll
                    for (int d = 0; d < QD; ++d) {
                        if (d >= nd) break;
                        const float4 w4 = lds_f4(ring_w + (st_cons * QD + d) * kWStage);
                        const unsigned src = ring_l + (st_cons * QD + d) * kLStage;
                        if (m == 0xF) {
#pragma unroll
                            for (int k = 0; k < KC; ++k)
                                if (FULL || k < kc) {
                                    const float4 l = Slot<LT>::read(src + k * kLSlot);
                                    a[d][k].x = __fadd_rn(a[d][k].x, __fmul_rn(w4.x, l.x));
                                    a[d][k].y = __fadd_rn(a[d][k].y, __fmul_rn(w4.y, l.y));
                                    a[d][k].z = __fadd_rn(a[d][k].z, __fmul_rn(w4.z, l.z));
                                    a[d][k].w = __fadd_rn(a[d][k].w, __fmul_rn(w4.w, l.w));
                                }
                        } else {  // a window that covers part of the quad: the other ring lanes hold stale bytes
#pragma unroll
                            for (int k = 0; k < KC; ++k)
                                if (FULL || k < kc) {
                                    const float4 l = Slot<LT>::read(src + k * kLSlot);
                                    if (m & 1) a[d][k].x = __fadd_rn(a[d][k].x, __fmul_rn(w4.x, l.x));
                                    if (m & 2) a[d][k].y = __fadd_rn(a[d][k].y, __fmul_rn(w4.y, l.y));
                                    if (m & 4) a[d][k].z = __fadd_rn(a[d][k].z, __fmul_rn(w4.z, l.z));
                                    if (m & 8) a[d][k].w = __fadd_rn(a[d][k].w, __fmul_rn(w4.w, l.w));
                                }
                        }
                    }
                    st_cons = st_cons + 1 == S ? 0 : st_cons + 1;
                }
#pragma unroll
                for (int d = 0; d < QD; ++d) {
                    if (d >= nd) break;
                    if (p.fuse == MSS_FUSE_LOGITS && complete) {
#pragma unroll
                        for (int k = 0; k < KC; ++k) {  // engine/utils.py:151
                            if (complete & 1u) a[d][k].x = __fdiv_rn(a[d][k].x, cnt[d][0]);
                            if (complete & 2u) a[d][k].y = __fdiv_rn(a[d][k].y, cnt[d][1]);
                            if (complete & 4u) a[d][k].z = __fdiv_rn(a[d][k].z, cnt[d][2]);
                            if (complete & 8u) a[d][k].w = __fdiv_rn(a[d][k].w, cnt[d][3]);
                        }
                    }
                    if (p.fuse == MSS_FUSE_LABELS && complete) {
                        // argmax of the raw weighted sums: dividing every class by the same positive count cannot reorder them
#pragma unroll
                        for (int k = 0; k < KC; ++k)
                            if (FULL || k < kc) {
                                am[d][0].push(a[d][k].x, k0 + k);
                                am[d][1].push(a[d][k].y, k0 + k);
                                am[d][2].push(a[d][k].z, k0 + k);
                                am[d][3].push(a[d][k].w, k0 + k);
                            }
                    }
                    if (accb != nullptr && !all_to_labels) {  // skipped only when the whole quad was finished into labels
#pragma unroll
                        for (int k = 0; k < KC; ++k)
                            if (FULL || k < kc) *reinterpret_cast<float4*>(accb + d * plane + (k0 + k) * cstride) = a[d][k];
                    }
                }
            }
            if (p.fuse == MSS_FUSE_LABELS && complete) {
#pragma unroll
                for (int d = 0; d < QD; ++d) {
                    if (d >= nd) break;
                    uint8_t* lab = p.labels + (static_cast<long long>(b) * g.ext[0] + pl0 + pl + d) * g.ext[1] * p.label_pitch +
                                   static_cast<long long>(lh) * p.label_pitch + lw;
                    unsigned packed = 0;
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (complete & (1u << e)) {
                            packed |= static_cast<unsigned>(am[d][e].label()) << (8 * e);
                            ties += am[d][e].near_tie(p.tie_tol) ? 1u : 0u;
                        }
                    if (complete == 0xFu && (reinterpret_cast<uintptr_t>(lab) & 3u) == 0) {
                        *reinterpret_cast<unsigned*>(lab) = packed;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (complete & (1u << e)) lab[e] = static_cast<uint8_t>(packed >> (8 * e));
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    if (ties && p.near_ties != nullptr) atomicAdd(p.near_ties, static_cast<unsigned long long>(ties));
}

template <typename LT, int KC, int S, int MINB, int QD>
static cudaError_t launch_cells_one(dim3 grid, cudaStream_t s, const CellParams& cp) {
    constexpr size_t smem = acc_smem_bytes<LT, KC, S * QD>();
    const bool full = cp.a.g.K % KC == 0;
    // the opt-in is per device (context): set it on every launch - a process may stitch on several GPUs
    cudaError_t e = full ? cudaFuncSetAttribute(accumulate_cells_kernel<LT, KC, S, MINB, true, QD>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))
                         : cudaFuncSetAttribute(accumulate_cells_kernel<LT, KC, S, MINB, false, QD>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    if (full)
        accumulate_cells_kernel<LT, KC, S, MINB, true, QD><<<grid, kAccThreads, smem, s>>>(cp);
    else
        accumulate_cells_kernel<LT, KC, S, MINB, false, QD><<<grid, kAccThreads, smem, s>>>(cp);
    return cudaGetLastError();
}

template <typename LT>
static cudaError_t launch_cells_for_k(int K, dim3 grid, cudaStream_t s, const CellParams& cp) {
    // classes per register pass (KC), ring depth (S), CTAs per SM, planes per fetch (QD): picked on B200 with
    // benchmarks/kernel_bench.py
    static const int variant = getenv("MSS_ACC_VARIANT") ? atoi(getenv("MSS_ACC_VARIANT")) : 0;  // tuning knob
    if (K % 7 == 0 && variant == 1) return launch_cells_one<LT, 7, 4, 4, 1>(grid, s, cp);
    if (K % 7 == 0) return launch_cells_one<LT, 7, 3, 4, 1>(grid, s, cp);
    if (K % 8 == 0) return launch_cells_one<LT, 8, 3, 4, 1>(grid, s, cp);
    if (K % 5 == 0) return launch_cells_one<LT, 5, 4, 4, 1>(grid, s, cp);
    if (K % 4 == 0) return launch_cells_one<LT, 4, 4, 4, 1>(grid, s, cp);
    if (K % 3 == 0 && variant == 1) return launch_cells_one<LT, 3, 6, 4, 1>(grid, s, cp);
    if (K % 3 == 0) return launch_cells_one<LT, 3, 3, 4, 2>(grid, s, cp);
    if (K == 2) return launch_cells_one<LT, 2, 4, 4, 2>(grid, s, cp);
    return launch_cells_one<LT, 8, 3, 4, 1>(grid, s, cp);
}

// Cuts the launch box into cells and launches the cell kernel.  Returns 0 when it ran, < 0 when the geometry does not fit
// its tables (the caller then runs the general kernel).
int launch_cells(const mss_layout_t* lay, const AccParams& p, int logits_dtype, cudaStream_t s, cudaError_t* err) {
    *err = cudaSuccess;
    const Geo& g = p.g;
    if (!p.vec_ok) return -1;
    // window-relative offsets of the cell kernel are 32-bit: (plane * roi_h + row) * roi_w + column over the whole image
    if ((static_cast<long long>(g.img[0]) * g.roi[1] + g.img[1]) * g.roi[2] + g.img[2] >= (1LL << 31)) return -1;
    static const int disabled = getenv("MSS_ACC_NO_CELLS") ? atoi(getenv("MSS_ACC_NO_CELLS")) : 0;
    if (disabled) return -1;
    const int32_t* t = lay->table_host;
    CellParams cp;
    cp.a = p;
    // per axis: window starts, breakpoints (local coordinates) of the owned windows inside the box
    int bp[3][2 * kCellMaxSeg + 4];
    int nbp[3];
    const int32_t* starts[3];
    for (int a = 0; a < 3; ++a) {
        const int32_t* st = t + t[kHdrOffStarts + a];
        starts[a] = st;
        if (lay->n_starts[a] > kCellMaxSeg) return -1;
        for (int i = 0; i < lay->n_starts[a]; ++i) cp.starts[a][i] = st[i];
        const int lo = p.box_lo[a], hi = p.box_lo[a] + p.box_n[a];
        int n = 0;
        bp[a][n++] = lo;
        bp[a][n++] = hi;
        for (int i = lay->win_lo[a]; i < lay->win_hi[a]; ++i) {
            const int v[2] = {st[i] - lay->origin[a], st[i] + lay->roi[a] - lay->origin[a]};
            for (int e = 0; e < 2; ++e)
                if (v[e] > lo && v[e] < hi) {
                    if (n >= 2 * kCellMaxSeg) return -1;
                    bp[a][n++] = v[e];
                }
        }
        for (int i = 1; i < n; ++i) {  // insertion sort, then unique
            const int v = bp[a][i];
            int j = i - 1;
            for (; j >= 0 && bp[a][j] > v; --j) bp[a][j + 1] = bp[a][j];
            bp[a][j + 1] = v;
        }
        int m = 0;
        for (int i = 0; i < n; ++i)
            if (m == 0 || bp[a][i] != bp[a][m - 1]) bp[a][m++] = bp[a][i];
        nbp[a] = m;
        if (m - 1 > kCellMaxSeg) return -1;
    }
    // owned windows [w0, w1) covering any of the local coordinates [x0, x1]
    auto cover = [&](int a, int x0, int x1, int* w0, int* w1) {
        int lo = lay->win_hi[a], hi = lay->win_lo[a];
        for (int i = lay->win_lo[a]; i < lay->win_hi[a]; ++i) {
            const int s0 = starts[a][i] - lay->origin[a];
            if (s0 <= x1 && x0 < s0 + lay->roi[a]) {
                lo = i < lo ? i : lo;
                hi = i + 1 > hi ? i + 1 : hi;
            }
        }
        if (hi < lo) hi = lo;
        *w0 = lo;
        *w1 = hi;
    };
    int cover_max[3] = {0, 0, 0};
    for (int a = 0; a < 2; ++a) {
        cp.n_seg[a] = nbp[a] - 1;
        for (int i = 0; i + 1 < nbp[a]; ++i) {
            cp.seg_lo[a][i] = bp[a][i];
            cp.seg_n[a][i] = bp[a][i + 1] - bp[a][i];
            cover(a, bp[a][i], bp[a][i], &cp.seg_w0[a][i], &cp.seg_w1[a][i]);
            const int c = cp.seg_w1[a][i] - cp.seg_w0[a][i];
            cover_max[a] = c > cover_max[a] ? c : cover_max[a];
        }
    }
    // W: quads completely inside one segment and completely real -> one tile per segment; every other quad column (it
    // straddles a segment boundary or holds the ragged end of the row) is a tile of its own
    const int q_lo = p.box_lo[2] / 4, q_hi = (p.box_lo[2] + p.box_n[2] + 3) / 4;
    int n_wt = 0;
    bool any_scalar = false;
    const int last_real = g.ext[2] - 1;
    auto add_tile = [&](int q0, int nq, bool single) {
        if (n_wt >= kCellMaxSeg) return false;
        cp.wt_q0[n_wt] = q0;
        cp.wt_nq[n_wt] = nq;
        const int x1 = single ? (q0 * 4 + 3 < last_real ? q0 * 4 + 3 : last_real) : q0 * 4;
        cover(2, q0 * 4, x1, &cp.wt_w0[n_wt], &cp.wt_w1[n_wt]);
        const int c = cp.wt_w1[n_wt] - cp.wt_w0[n_wt];
        cover_max[2] = c > cover_max[2] ? c : cover_max[2];
        ++n_wt;
        if (single) any_scalar = true;
        return true;
    };
    int q_done = q_lo;  // quads below this are assigned
    for (int i = 0; i + 1 < nbp[2]; ++i) {
        const int a0 = bp[2][i], b0 = bp[2][i + 1] < g.ext[2] ? bp[2][i + 1] : g.ext[2];
        int qa = (a0 + 3) / 4;
        const int qb = b0 / 4;  // interior quads [qa, qb)
        qa = qa < q_hi ? qa : q_hi;
        for (int q = q_done; q < qa; ++q)
            if (!add_tile(q, 1, true)) return -1;
        q_done = qa > q_done ? qa : q_done;
        while (qb > q_done) {  // (a segment longer than a CTA column is cut)
            const int nq = qb - q_done < kCellMaxColQuads ? qb - q_done : kCellMaxColQuads;
            if (!add_tile(q_done, nq, false)) return -1;
            q_done += nq;
        }
    }
    for (int q = q_done; q < q_hi; ++q)
        if (!add_tile(q, 1, true)) return -1;
    if (n_wt == 0) return -1;
    if (cover_max[0] * cover_max[1] * cover_max[2] > kCellMaxCand || cover_max[0] * cover_max[1] * cover_max[2] == 0) return -1;
    for (int i = lay->win_lo[2]; i < lay->win_hi[2]; ++i)  // windows at an odd W offset are fetched element-wise too
        if (((lay->origin[2] - starts[2][i]) & 3) != 0) any_scalar = true;
    if (any_scalar && logits_dtype != MSS_F32) return -1;  // 16-bit logits have no 4-byte-granular copy path
    // CTA columns: consecutive tiles, as even as possible, each <= 64 quads and <= kCellMaxTiles tiles
    const int total_q = q_hi - q_lo;
    const int want_cols = (total_q + 63) / 64;
    static const int cols_mode = getenv("MSS_ACC_COLS") ? atoi(getenv("MSS_ACC_COLS")) : 0;  // tuning knob: 1 = a column per tile
    int n_col = 0;
    for (int w = 0; w < n_wt;) {
        if (n_col >= kCellMaxSeg) return -1;
        // few classes: a column per W tile (the launch is latency-bound and every lane of a warp then runs the same window
        // list and copy path); many classes: whole rows per CTA, so the pieces of a logits row are fetched together
        const bool per_tile = cols_mode == 1 || (cols_mode == 0 && g.K <= 4);
        const int target = per_tile ? 1 : (total_q + want_cols - 1) / want_cols;
        int nq = 0, ntile = 0;
        while (w + ntile < n_wt && ntile < kCellMaxTiles && (ntile == 0 || nq + cp.wt_nq[w + ntile] <= (target > 64 ? 64 : target)) &&
               nq + cp.wt_nq[w + ntile] <= kCellMaxColQuads) {
            nq += cp.wt_nq[w + ntile];
            ++ntile;
        }
        cp.col_t0[n_col] = w;
        cp.col_nt[n_col] = ntile;
        // rows per item: `ri` passes of 128 positions
        cp.col_hr[n_col] = nq;  // (quads for now; turned into rows below)
        ++n_col;
        w += ntile;
    }
    cp.n_col = n_col;
    // how much a thread takes per item: as much as leaves >= ~6 items per resident CTA slot (148 SMs x 4)
    const long long nz = static_cast<long long>((p.g1 - 1) / g.n_local) - p.b_lo + 1;
    const long long quads = static_cast<long long>(p.box_n[0]) * p.box_n[1] * total_q * nz;
    const long long batches = (quads + kAccThreads - 1) / kAccThreads;
    int ri = 1, pd = 1;
    const int choices[4][2] = {{2, 2}, {1, 2}, {1, 2}, {1, 2}};  // (pd stays even: two planes share a fetch when K is small)
    for (int c = 0; c < 4; ++c) {
        ri = choices[c][0];
        pd = choices[c][1];
        if (batches / (ri * pd) >= 148LL * 4 * 6) break;
    }
    static const int force_ri = getenv("MSS_ACC_RI") ? atoi(getenv("MSS_ACC_RI")) : 0;
    static const int force_pd = getenv("MSS_ACC_PD") ? atoi(getenv("MSS_ACC_PD")) : 0;
    if (force_ri > 0) ri = force_ri;
    if (force_pd > 0) pd = force_pd;
    cp.pd = pd;
    long long nx = 0;
    for (int c = 0; c < n_col; ++c) {
        const int nq = cp.col_hr[c];
        int hr = (ri * kAccThreads + nq / 2) / nq;  // rows whose positions fill `ri` passes best
        hr = hr < 1 ? 1 : hr;
        cp.col_hr[c] = hr;
        cp.x_prefix[c] = static_cast<int>(nx);
        for (int h = 0; h < cp.n_seg[1]; ++h) nx += (cp.seg_n[1][h] + hr - 1) / hr;
    }
    cp.x_prefix[n_col] = static_cast<int>(nx);
    long long ny = 0;
    for (int d = 0; d < cp.n_seg[0]; ++d) {
        cp.y_prefix[d] = static_cast<int>(ny);
        ny += (cp.seg_n[0][d] + pd - 1) / pd;
    }
    cp.y_prefix[cp.n_seg[0]] = static_cast<int>(ny);
    if (nx <= 0 || nx > 0x7fffffffLL || ny <= 0 || ny > 65535 || nz <= 0 || nz > 65535) return -1;
    dim3 grid(static_cast<unsigned>(nx), static_cast<unsigned>(ny), static_cast<unsigned>(nz));
    if (logits_dtype == MSS_F32)
        *err = launch_cells_for_k<float>(g.K, grid, s, cp);
    else if (logits_dtype == MSS_F16)
        *err = launch_cells_for_k<__half>(g.K, grid, s, cp);
    else
        *err = launch_cells_for_k<__nv_bfloat16>(g.K, grid, s, cp);
    return 0;
}

}  // namespace mss
