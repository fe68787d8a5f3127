// Row-staged weighted overlap accumulation + normalise / argmax (engine/utils.py:137-151 + engine/test.py:140-141): the kernel
// of a volume whose windows all sit in ONE launch (cfg2, cfg4, cfg5): labels (or the normalised logits) straight out, no
// accumulator round trip.  fp32 logits, K <= 16, grids with <= 4 window positions along W.
//
// Why a third accumulation kernel: in the cell kernel (accumulate_cells.cu) every thread issues its own 16-byte cp.async per
// class, plane and window (four 4-byte copies for a window that starts off the 16-byte lattice, such as BraTS' clamped start
// 59).  With few classes that is instruction-bound (~280 instructions per voxel on 57 bytes, ncu:
// profiles/r2_ncu_acc_cells_k3.md), and at K = 14 it stops at 0.9 of the copy bandwidth.  Here the data movement costs the
// threads next to nothing: a CTA works on whole rows of the volume inside one (D segment, H segment) CELL - all its rows
// are covered by the same windows along D and H - and for every window (ascending index, all W positions) it stages the
// K + 1 row sets (importance map + K logit planes) of a tile of rows with cp.async.bulk (global -> shared, completing on an
// mbarrier), one copy per run of rows that share a plane: a window row is contiguous and 16-byte aligned whatever the
// window's start in the volume, so an off-lattice window costs nothing extra on the way in.  The threads then read their
// 4 voxels from the shared-memory ring (one 16-byte read per plane; scalar reads where the window is off the lattice or
// covers the quad partly) and do 8 K flops per covering window.
// Arithmetic and window order are those of the other kernels: acc = fadd_rn(acc, fmul_rn(w, logit)) in ascending window
// index; labels: first-max argmax of the raw sums (a common positive divisor cannot reorder them), near-ties counted;
// logits: fdiv_rn(sum, ascending fp32 sum of the covering windows' weights).
// Measured on B200: cfg2 (K = 14) 2.9 ms = 1.02-1.04 of the measured copy bandwidth (cell kernel 3.4 ms, 0.89); real BraTS
// geometry (K = 3) 0.10 ms = 0.75-0.77 (cell kernel 0.19 ms, 0.41).
#include "acc_common.cuh"

namespace mss {

constexpr int kRowsMaxSeg = 64;    // segments / window starts per axis
constexpr int kRowsMaxWin = 64;    // windows over one (D, H) cell, all W positions
constexpr int kRowsMaxK = 16;      // classes (one float4 accumulator per class, row and thread)
constexpr int kRowsMaxWinW = 4;    // W positions of the grid (every window of a cell is staged for whole rows, so a thread's quad
                                   // should be covered by most of them: 2 of 4 at cfg2, 1.9 of 3 at BraTS size)
constexpr int kRowsThreads = 256;
constexpr int kRowsMaxTr = 64;     // rows of a tile
constexpr int kRowsMaxRuns = 4;    // planes a tile's rows may touch (one bulk copy per plane-run)

struct RowsParams {
    const void* batch[MSS_MAX_BATCH_PTRS];
    int sw_batch;
    const float* imp;
    uint8_t* labels;
    int label_pitch;
    float* out;   // LOGITS mode: normalised logits [Nb, K, D, H, pitch]
    int pitch;
    float tie_tol;
    unsigned long long* near_ties;
    int roi[3], img[3], ns[3];
    long long n_local;
    int starts[3][kRowsMaxSeg];
    int seg_lo[2][kRowsMaxSeg];  // axes 0 (D), 1 (H): first coordinate of a segment,
    int seg_n[2][kRowsMaxSeg];   // its length,
    int seg_w0[2][kRowsMaxSeg];  // the windows [w0, w1) that cover it
    int seg_w1[2][kRowsMaxSeg];
    int n_seg[2];
    int tr;             // rows per tile
    int tiles_per_cta;
};

__device__ __forceinline__ unsigned rows_smem(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void rows_bar_init(uint64_t* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rows_smem(bar)));
}
__device__ __forceinline__ void rows_bar_expect(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rows_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rows_bar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(rows_smem(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void rows_bulk_copy(float* dst, const float* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rows_smem(dst)),
                 "l"(src), "r"(bytes), "r"(rows_smem(bar))
                 : "memory");
}

template <int K, int S, int RPT, bool LOGITS>  // classes, ring depth, rows per thread, sum / count out instead of labels
__global__ void __launch_bounds__(kRowsThreads, K <= 4 ? 4 : (K <= 8 ? 3 : 2)) accumulate_rows_kernel(const __grid_constant__ RowsParams p) {
    extern __shared__ __align__(128) float ring[];  // [S] stages of [K + 1][TR][roi_w]
    __shared__ __align__(8) uint64_t full[S];
    __shared__ const float* s_base[kRowsMaxWin];  // class-0 logits of the cell's windows, ascending window index
    __shared__ int s_term[kRowsMaxWin];           // (start_d * roi_h + start_h) * roi_w  [floats]
    __shared__ int s_sw[kRowsMaxWin];             // start_w
    const int tid = threadIdx.x;

    // ---- the cell (one D segment x one H segment, whole rows) and this CTA's run of tiles inside it --------------------
    const int sh = blockIdx.y;
    const int b = blockIdx.z / p.n_seg[0], sd = blockIdx.z - b * p.n_seg[0];
    const int n1 = p.seg_n[1][sh];
    const int rows_cell = p.seg_n[0][sd] * n1;  // rows (plane, row) of the cell, plane-major
    const int TR = p.tr;
    const int tiles = (rows_cell + TR - 1) / TR;
    const int tile0 = blockIdx.x * p.tiles_per_cta;
    if (tile0 >= tiles) return;
    const int ntile = min(p.tiles_per_cta, tiles - tile0);
    const int W = p.img[2], nq = (W + 3) >> 2;  // a thread owns 4 consecutive voxels of a row
    const int rh = p.roi[1], rw = p.roi[2];
    const long long R = static_cast<long long>(p.roi[0]) * rh * rw;
    const int d_lo = p.seg_lo[0][sd], h_lo = p.seg_lo[1][sh];
    const float inv_n1 = 1.f / static_cast<float>(n1);
    const int stage_floats = (K + 1) * TR * rw;
    const unsigned row_bytes = static_cast<unsigned>(rw) * 4u;

    const int dw0 = p.seg_w0[0][sd], hw0 = p.seg_w0[1][sh];
    const int nh = p.seg_w1[1][sh] - hw0, nw = p.ns[2];
    const int nwin = (p.seg_w1[0][sd] - dw0) * nh * nw;  // <= kRowsMaxWin (host)
    if (tid < nwin) {
        const int iw = tid % nw, ih = hw0 + (tid / nw) % nh, id = dw0 + tid / (nw * nh);
        const long long gi = static_cast<long long>(b) * p.n_local + (static_cast<long long>(id) * p.ns[1] + ih) * p.ns[2] + iw;
        const long long bi = gi / p.sw_batch;
        s_base[tid] = static_cast<const float*>(p.batch[bi]) + (gi - bi * p.sw_batch) * K * R;
        s_sw[tid] = p.starts[2][iw];
        s_term[tid] = (p.starts[0][id] * rh + p.starts[1][ih]) * rw;
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) rows_bar_init(&full[s]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // (plane, row) of row r of tile T (global coordinates); rows_cell < 2^22, so the float quotient is exact after the fix-up
    auto row_dh = [&](int T, int r, int* d, int* h) {
        const int rho = T * TR + r;
        int q = __float2int_rz((static_cast<float>(rho) + 0.5f) * inv_n1);
        q -= q * n1 > rho ? 1 : 0;
        q += (q + 1) * n1 <= rho ? 1 : 0;
        *d = d_lo + q;
        *h = h_lo + rho - q * n1;
    };
    // ---- producer side: the rows of a tile that lie in one plane are contiguous in the window (and in the stage), so a
    // step is (K + 1) x <= kRowsMaxRuns bulk copies of up to TR rows each; the copy slots are dealt out one per warp first
    // (a warp issues its bulk copies one lane at a time) ---------------------------------------------------------------
    const int slot = (tid & 31) * (kRowsThreads / 32) + (tid >> 5);  // warp w, lane l -> slot l * 8 + w
    const int cp_c = slot / kRowsMaxRuns, cp_u = slot - cp_c * kRowsMaxRuns;
    auto issue = [&](int t, int j, int s) {  // stage tile t's rows of window j
        if (cp_c > K) return;
        const int rho0 = (tile0 + t) * TR;
        const int nrows = min(TR, rows_cell - rho0);
        int q = __float2int_rz((static_cast<float>(rho0) + 0.5f) * inv_n1);  // plane of the tile's first row (see row_dh)
        q -= q * n1 > rho0 ? 1 : 0;
        q += (q + 1) * n1 <= rho0 ? 1 : 0;
        const int pl = q + cp_u;                                 // this slot's plane (cell-relative)
        const int ra = max(rho0, pl * n1), rb = min(rho0 + nrows, (pl + 1) * n1);  // its rows [ra, rb) of the cell
        if (rb <= ra) return;
        const long long off = static_cast<long long>(((d_lo + pl) * rh + h_lo + (ra - pl * n1)) * rw - s_term[j]);
        const float* src = cp_c == 0 ? p.imp + off : s_base[j] + (cp_c - 1) * R + off;
        rows_bulk_copy(ring + static_cast<size_t>(s) * stage_floats + (cp_c * TR + (ra - rho0)) * rw, src,
                       static_cast<unsigned>(rb - ra) * row_bytes, &full[s]);
    };
    auto expect = [&](int t, int s) {  // thread 0, before the copies of that step are issued
        const int nrows = min(TR, rows_cell - (tile0 + t) * TR);
        rows_bar_expect(&full[s], static_cast<unsigned>((K + 1) * nrows) * row_bytes);
    };

    const int total = ntile * nwin;
    int it = 0, ij = 0, iq = 0;  // producer cursor: step iq = (tile it, window ij)
    {
        const int pro = min(S, total);
        if (tid == 0) {
            int t = 0, j = 0;
            for (int q = 0; q < pro; ++q) {
                expect(t, q);
                if (++j == nwin) j = 0, ++t;
            }
        }
        __syncthreads();
        for (; iq < pro; ++iq) {
            issue(it, ij, iq);
            if (++ij == nwin) ij = 0, ++it;
        }
    }

    // ---- consumer side: 4 voxels of each of RPT rows (rows r_c, r_c + TRP, ...: more rows per step = fewer steps, and a step
    // costs a wait, an arrival, a block barrier and the copy issue whatever it carries) -----------------------------------
    const int TRP = kRowsThreads / nq;  // rows one pass of the CTA's threads covers
    const int r_c = tid / nq, g = tid - r_c * nq;
    float4 a[RPT][K];
    float4 cnt[LOGITS ? RPT : 1];  // LOGITS: the weight count (engine/utils.py:148: ascending fp32 sum of the covering windows' weights)
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr) {
        if (LOGITS) cnt[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < K; ++k) a[rr][k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    unsigned ties = 0;
    int t = 0, j = 0;
    int s = 0;
    unsigned phase = 0;  // ring stage of step q and the parity of its barrier phase
    int roff[RPT];       // this thread's rows inside a stage plane [floats]; -1: no such row
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr) {
        const int r = r_c + rr * TRP;
        roff[rr] = r_c < TRP && r < TR ? r * rw : -1;
    }
    const int plane_floats = TR * rw;
    const float* stg = ring;  // stage of step q
    auto fma4 = [&](int rr, int k, const float4& w4, const float4& v) {  // engine/utils.py:147: product, then sum
        a[rr][k].x = __fadd_rn(a[rr][k].x, __fmul_rn(w4.x, v.x));
        a[rr][k].y = __fadd_rn(a[rr][k].y, __fmul_rn(w4.y, v.y));
        a[rr][k].z = __fadd_rn(a[rr][k].z, __fmul_rn(w4.z, v.z));
        a[rr][k].w = __fadd_rn(a[rr][k].w, __fmul_rn(w4.w, v.w));
    };
    auto add_cnt = [&](int rr, const float4& w4) {
        if (LOGITS) {
            cnt[rr].x = __fadd_rn(cnt[rr].x, w4.x), cnt[rr].y = __fadd_rn(cnt[rr].y, w4.y);
            cnt[rr].z = __fadd_rn(cnt[rr].z, w4.z), cnt[rr].w = __fadd_rn(cnt[rr].w, w4.w);
        }
    };
    int nrows_t = min(TR, rows_cell - tile0 * TR);  // rows of the current tile
    for (int q = 0; q < total; ++q) {
        const int T = tile0 + t;
        const int l = g * 4 - s_sw[j];  // window-local column of the quad's first voxel
        const bool last_win = j + 1 == nwin;
        const bool covered = l > -4 && l < rw, inside = l >= 0 && l + 4 <= rw;
        rows_bar_wait(&full[s], phase);
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) {
            const int r = r_c + rr * TRP;
            const bool active = roff[rr] >= 0 && r < nrows_t;
            if (active && covered) {
                const float* st = stg + roff[rr] + l;
                if (inside && (l & 3) == 0) {
                    const float4 w4 = *reinterpret_cast<const float4*>(st);
                    add_cnt(rr, w4);
#pragma unroll
                    for (int k = 0; k < K; ++k) fma4(rr, k, w4, *reinterpret_cast<const float4*>(st + (k + 1) * plane_floats));
                } else {  // off the 16-byte lattice (BraTS' clamped start 59), or the window covers the quad partly: scalar reads
                          // (two aligned 16-byte reads per plane + a register shuffle were measured slower: twice the shared-memory
                          // wavefronts)
                    const bool c0 = l >= 0, c1 = l + 1 >= 0 && l + 1 < rw, c2 = l + 2 >= 0 && l + 2 < rw, c3 = l + 3 < rw;
                    const float w0 = c0 ? st[0] : 0.f, w1 = c1 ? st[1] : 0.f, w2 = c2 ? st[2] : 0.f, w3 = c3 ? st[3] : 0.f;
                    if (LOGITS) {
                        if (c0) cnt[rr].x = __fadd_rn(cnt[rr].x, w0);
                        if (c1) cnt[rr].y = __fadd_rn(cnt[rr].y, w1);
                        if (c2) cnt[rr].z = __fadd_rn(cnt[rr].z, w2);
                        if (c3) cnt[rr].w = __fadd_rn(cnt[rr].w, w3);
                    }
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const float* lp = st + (k + 1) * plane_floats;
                        if (c0) a[rr][k].x = __fadd_rn(a[rr][k].x, __fmul_rn(w0, lp[0]));
                        if (c1) a[rr][k].y = __fadd_rn(a[rr][k].y, __fmul_rn(w1, lp[1]));
                        if (c2) a[rr][k].z = __fadd_rn(a[rr][k].z, __fmul_rn(w2, lp[2]));
                        if (c3) a[rr][k].w = __fadd_rn(a[rr][k].w, __fmul_rn(w3, lp[3]));
                    }
                }
            }
            if (active && last_win && LOGITS) {  // the tile's voxels are complete: sum / count (engine/utils.py:151)
                int d, h;
                row_dh(T, r, &d, &h);
                const int nv = min(4, W - g * 4);
                const long long cs = static_cast<long long>(p.img[0]) * p.img[1] * p.pitch;
                float* o = p.out + static_cast<long long>(b) * K * cs + (static_cast<long long>(d) * p.img[1] + h) * p.pitch + g * 4;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float4 v = a[rr][k];
                    v.x = __fdiv_rn(v.x, cnt[rr].x);
                    v.y = nv > 1 ? __fdiv_rn(v.y, cnt[rr].y) : 0.f;  // (the pad voxels of a ragged row stay 0)
                    v.z = nv > 2 ? __fdiv_rn(v.z, cnt[rr].z) : 0.f;
                    v.w = nv > 3 ? __fdiv_rn(v.w, cnt[rr].w) : 0.f;
                    *reinterpret_cast<float4*>(o + k * cs) = v;
                    a[rr][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                cnt[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (active && last_win && !LOGITS) {  // the tile's voxels are complete: first-max argmax of the raw sums
                int d, h;
                row_dh(T, r, &d, &h);
                ArgmaxState am[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) am[e].reset();
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    am[0].push(a[rr][k].x, k);
                    am[1].push(a[rr][k].y, k);
                    am[2].push(a[rr][k].z, k);
                    am[3].push(a[rr][k].w, k);
                    a[rr][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                uint8_t* lab = p.labels + ((static_cast<long long>(b) * p.img[0] + d) * p.img[1] + h) * p.label_pitch + g * 4;
                const int nv = min(4, W - g * 4);
                unsigned packed = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    packed |= static_cast<unsigned>(am[e].label()) << (8 * e);
                    if (e < nv) ties += am[e].near_tie(p.tie_tol) ? 1u : 0u;
                }
                if (nv == 4 && (reinterpret_cast<uintptr_t>(lab) & 3u) == 0) {
                    *reinterpret_cast<unsigned*>(lab) = packed;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (e < nv) lab[e] = static_cast<uint8_t>(packed >> (8 * e));
                }
            }
        }
        if (last_win) {
            j = 0, ++t;
            nrows_t = min(TR, rows_cell - (tile0 + t) * TR);
        } else {
            ++j;
        }
        if (iq < total) {  // refill the stage just read with step q + S (uniform over the CTA)
            if (tid == 0) expect(it, s);
            __syncthreads();  // everybody has read the stage; the arrival above precedes every copy's complete_tx
            issue(it, ij, s);
            ++iq;
            if (++ij == nwin) ij = 0, ++it;
        }
        stg += stage_floats;
        if (++s == S) s = 0, phase ^= 1u, stg = ring;
    }
    if (ties && p.near_ties != nullptr) atomicAdd(p.near_ties, static_cast<unsigned long long>(ties));
}

// Returns 0 when the launch was made, < 0 when this launch is not the kernel's case (the caller goes on to the cell kernel).
int launch_rows(const mss_layout_t* lay, const AccParams& ap, int logits_dtype, cudaStream_t s, cudaError_t* err) {
    *err = cudaSuccess;
    const Geo& g = ap.g;
    static const int disabled = getenv("MSS_ACC_NO_ROWS") ? atoi(getenv("MSS_ACC_NO_ROWS")) : 0;
    static const int max_k = getenv("MSS_ROWS_MAXK") ? atoi(getenv("MSS_ROWS_MAXK")) : kRowsMaxK;  // tuning knob
    if (disabled || logits_dtype != MSS_F32 || ap.fuse == MSS_FUSE_NONE || !ap.vec_ok || g.K < 1 || g.K > max_k || g.K > 16) return -1;
    const bool logits_out = ap.fuse == MSS_FUSE_LOGITS;
    if (logits_out && (ap.acc == nullptr || g.pitch % 4 != 0 || g.pitch < g.img[2])) return -1;
    const long long total = g.n_local * g.nb;
    if (ap.g0 != 0 || ap.g1 != total || ap.own0 != 0 || ap.own1 != total) return -1;  // every window, one launch
    for (int a = 0; a < 3; ++a)
        if (g.org[a] != 0 || g.ext[a] != g.img[a] || g.wlo[a] != 0 || g.whi[a] != g.ns[a] || g.ns[a] > kRowsMaxSeg) return -1;
    // whole rows per CTA: small volumes only (every W position of the grid is staged for every row)
    const int nq = (g.img[2] + 3) / 4;
    if (g.ns[2] > kRowsMaxWinW || nq > kRowsThreads || g.nb > 1024) return -1;
    RowsParams p;
    for (int i = 0; i < MSS_MAX_BATCH_PTRS; ++i) p.batch[i] = ap.batch[i];
    p.sw_batch = ap.sw_batch;
    p.imp = ap.imp;
    p.labels = ap.labels;
    p.out = ap.acc;
    p.pitch = g.pitch;
    p.label_pitch = ap.label_pitch;
    p.tie_tol = ap.tie_tol;
    p.near_ties = ap.near_ties;
    p.n_local = g.n_local;
    const int32_t* t = lay->table_host;
    int cover_max[2] = {0, 0}, rows_max[2] = {0, 0};
    for (int a = 0; a < 3; ++a) {
        p.roi[a] = g.roi[a];
        p.img[a] = g.img[a];
        p.ns[a] = g.ns[a];
        const int32_t* st = t + t[kHdrOffStarts + a];
        for (int i = 0; i < g.ns[a]; ++i) p.starts[a][i] = st[i];
        if (a == 2) {  // along W every voxel must be covered (it is, by a sliding-window grid)
            for (int i = 0; i + 1 < g.ns[a]; ++i)
                if (st[i + 1] > st[i] + g.roi[a]) return -1;
            if (st[0] != 0 || st[g.ns[a] - 1] + g.roi[a] < g.img[a]) return -1;
            continue;
        }
        int bp[2 * kRowsMaxSeg + 2], n = 0;
        bp[n++] = 0;
        bp[n++] = g.img[a];
        for (int i = 0; i < g.ns[a]; ++i) {
            const int v[2] = {st[i], st[i] + g.roi[a]};
            for (int e = 0; e < 2; ++e)
                if (v[e] > 0 && v[e] < g.img[a]) bp[n++] = v[e];
        }
        for (int i = 1; i < n; ++i) {  // insertion sort, then unique
            const int v = bp[i];
            int j = i - 1;
            for (; j >= 0 && bp[j] > v; --j) bp[j + 1] = bp[j];
            bp[j + 1] = v;
        }
        int m = 0;
        for (int i = 0; i < n; ++i)
            if (m == 0 || bp[i] != bp[m - 1]) bp[m++] = bp[i];
        if (m - 1 > kRowsMaxSeg) return -1;
        p.n_seg[a] = m - 1;
        for (int i = 0; i + 1 < m; ++i) {
            p.seg_lo[a][i] = bp[i];
            p.seg_n[a][i] = bp[i + 1] - bp[i];
            int lo = g.ns[a], hi = 0;  // windows covering the segment: a contiguous range (ascending starts, one roi)
            for (int w = 0; w < g.ns[a]; ++w)
                if (st[w] <= bp[i] && bp[i] < st[w] + g.roi[a]) {
                    lo = w < lo ? w : lo;
                    hi = w + 1 > hi ? w + 1 : hi;
                }
            if (hi <= lo) return -1;  // a voxel no window covers: not a sliding-window grid
            for (int w = lo; w < hi; ++w)
                if (!(st[w] <= bp[i] && bp[i + 1] <= st[w] + g.roi[a])) return -1;  // not contiguous / not complete
            p.seg_w0[a][i] = lo;
            p.seg_w1[a][i] = hi;
            cover_max[a] = hi - lo > cover_max[a] ? hi - lo : cover_max[a];
            rows_max[a] = p.seg_n[a][i] > rows_max[a] ? p.seg_n[a][i] : rows_max[a];
        }
    }
    if (cover_max[0] * cover_max[1] * g.ns[2] > kRowsMaxWin) return -1;
    if (static_cast<long long>(rows_max[0]) * rows_max[1] >= (1 << 22)) return -1;
    if ((static_cast<long long>(g.img[0]) * g.roi[1] + g.img[1]) * g.roi[2] + g.img[2] >= (1LL << 31)) return -1;  // 32-bit row terms
    // tiles: as many rows as rows x quads fill the CTA; a CTA takes `tiles_per_cta` consecutive tiles of a cell
    static const int force_tr = getenv("MSS_ROWS_TR") ? atoi(getenv("MSS_ROWS_TR")) : 0;  // tuning knob
    static const int force_rpt = getenv("MSS_ROWS_RPT") ? atoi(getenv("MSS_ROWS_RPT")) : 0;     // tuning knob
    int trp0 = kRowsThreads / nq;
    trp0 = trp0 > 16 ? 16 : trp0;
    // two rows per thread when three stages of them leave >= 3 CTAs per SM (<= 74 KB), else one
    int rpt = static_cast<size_t>(3) * (g.K + 1) * 2 * trp0 * g.roi[2] * sizeof(float) <= 74 * 1024 ? 2 : 1;
    if (force_rpt == 1 || force_rpt == 2) rpt = force_rpt;
    int trp = kRowsThreads / nq;
    trp = trp > 16 ? 16 : trp;
    int tr = rpt * trp;
    if (force_tr > 0 && force_tr < tr) tr = force_tr;
    tr = tr > kRowsMaxTr ? kRowsMaxTr : tr;
    int n1_min = 1 << 30;
    for (int i = 0; i < p.n_seg[1]; ++i) n1_min = p.seg_n[1][i] < n1_min ? p.seg_n[1][i] : n1_min;
    while (tr > 1 && (tr - 1) / n1_min + 2 > kRowsMaxRuns) --tr;  // a tile's rows touch at most kRowsMaxRuns planes
    p.tr = tr;
    const int rows_cell_max = rows_max[0] * rows_max[1];
    const int tiles_cell = (rows_cell_max + tr - 1) / tr;
    const long long tiles_total = static_cast<long long>(tiles_cell) * p.n_seg[0] * p.n_seg[1] * g.nb;
    static const int force_tpc = getenv("MSS_ROWS_TPC") ? atoi(getenv("MSS_ROWS_TPC")) : 0;  // tuning knob
    long long tpc = tiles_total / (148LL * 32);  // ~8 waves of CTAs (measured: BraTS 2 tiles per CTA best, 1 / 4 / 8 within 10 %)
    tpc = tpc < 2 ? 2 : (tpc > 8 ? 8 : tpc);  // (cfg2, K = 14: 1 / 2 / 8 / 16 / 42 tiles per CTA = 3.10 / 2.94 / 2.95 / 3.00 / 2.98 ms)
    if (force_tpc > 0) tpc = force_tpc;
    p.tiles_per_cta = static_cast<int>(tpc);
    const long long nx = (tiles_cell + p.tiles_per_cta - 1) / p.tiles_per_cta;
    static const int force_st = getenv("MSS_ROWS_STAGES") ? atoi(getenv("MSS_ROWS_STAGES")) : 0;  // tuning knob
    const int stages = force_st == 2 || force_st == 4 ? force_st : 3;  // 3: 55 KB per CTA at BraTS size, 4 CTAs per SM
    if (g.K > 4 && stages != 3) return -1;  // (the other depths are instantiated for few classes only)
    const size_t smem = static_cast<size_t>(stages) * (g.K + 1) * tr * g.roi[2] * sizeof(float);
    const long long nz = static_cast<long long>(g.nb) * p.n_seg[0];
    if (nx <= 0 || nx > 0x7fffffffLL || nz > 65535 || smem > 200 * 1024) return -1;
    const dim3 grid(static_cast<unsigned>(nx), static_cast<unsigned>(p.n_seg[1]), static_cast<unsigned>(nz));
    auto launch = [&](auto kernel) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));  // per device
        if (e != cudaSuccess) return e;
        kernel<<<grid, kRowsThreads, smem, s>>>(p);
        return cudaGetLastError();
    };
#define MSS_ROWS_PICK(KK, SS, LL) \
    (rpt == 1 ? launch(accumulate_rows_kernel<KK, SS, 1, LL>) : launch(accumulate_rows_kernel<KK, SS, 2, LL>))
#define MSS_ROWS_STAGED(KK, LL) \
    (stages == 2 ? MSS_ROWS_PICK(KK, 2, LL) : stages == 3 ? MSS_ROWS_PICK(KK, 3, LL) : MSS_ROWS_PICK(KK, 4, LL))
#define MSS_ROWS_CASE(KK)                                                           \
    case KK:                                                                        \
        *err = logits_out ? MSS_ROWS_STAGED(KK, true) : MSS_ROWS_STAGED(KK, false); \
        break;
#define MSS_ROWS_CASE3(KK)                                                            \
    case KK:                                                                          \
        *err = logits_out ? MSS_ROWS_PICK(KK, 3, true) : MSS_ROWS_PICK(KK, 3, false); \
        break;
    switch (g.K) {
        MSS_ROWS_CASE(1)
        MSS_ROWS_CASE(2)
        MSS_ROWS_CASE(3)
        MSS_ROWS_CASE(4)
        MSS_ROWS_CASE3(5)
        MSS_ROWS_CASE3(6)
        MSS_ROWS_CASE3(7)
        MSS_ROWS_CASE3(8)
        MSS_ROWS_CASE3(9)
        MSS_ROWS_CASE3(10)
        MSS_ROWS_CASE3(11)
        MSS_ROWS_CASE3(12)
        MSS_ROWS_CASE3(13)
        MSS_ROWS_CASE3(14)
        MSS_ROWS_CASE3(15)
        MSS_ROWS_CASE3(16)
    }
#undef MSS_ROWS_CASE3
#undef MSS_ROWS_STAGED
#undef MSS_ROWS_PICK
#undef MSS_ROWS_CASE
    return 0;
}

}  // namespace mss
