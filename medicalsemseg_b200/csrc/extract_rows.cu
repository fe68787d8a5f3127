// Volume-stationary ROI patch extraction (engine/utils.py:122-125,133): every row of the volume is read from HBM ONCE and
// written to all the windows that contain it.
//
// The window-stationary kernels of extract.cu copy window by window, so a voxel covered by 8 windows (overlap 0.5) is read 8
// times; the re-reads along W and H hit L2, the ones along D do not (a layer of 100 windows writes 350 MB between them):
// ncu counts 0.46 GB of DRAM reads per 300 windows where 0.21 GB are compulsory, and the TMA kernel stops at 0.72 of the
// roofline on compulsory bytes (4 V Cin + 4 N Cin R).  Here a CTA (one warp; no thread ever touches the data) works on whole
// rows inside one (D segment, H segment) cell - all its rows are covered by the same windows along D and H:
//   * per tile of rows and per window position along W, one cp.async.bulk (global -> shared, mbarrier complete_tx) per row
//     brings the row's segment [start_w, start_w + roi_w) into the stage plane of that W position - the segments of the
//     W positions overlap, but they are requested together and meet in L2;
//   * per window of the cell inside this call's range, ONE cp.async.bulk (shared -> global, bulk_group) per run of rows that
//     share a plane writes TR x roi_w floats: consecutive rows of a window are contiguous in the patch.
// Three stages: the loads of tile t + 1 go out while the stores of tile t drain.
// Needs what the TMA kernel needs: every window inside the volume, rows and window starts 16-byte tiled along W.
#include <cstdlib>

#include "common.cuh"

namespace mss {

constexpr int kXrMaxSeg = 64;
constexpr int kXrMaxWin = 64;   // windows of a cell (all W positions) inside the call's range
constexpr int kXrMaxWinW = 8;   // window positions along W
constexpr int kXrStages = 3;
constexpr int kXrMaxRuns = 4;

struct XrParams {
    const float* vol;
    float* out;
    float* centers;
    long long first_window;
    int n_windows;
    int n_channels;
    int nb;
    int vorg[3], vext[3];
    int roi[3], img[3], ns[3];
    long long n_local;
    int starts[3][kXrMaxSeg];
    int seg_lo[2][kXrMaxSeg], seg_n[2][kXrMaxSeg], seg_w0[2][kXrMaxSeg], seg_w1[2][kXrMaxSeg];
    int n_seg[2];
    int tr;
    int tiles_per_cta;
};

__device__ __forceinline__ unsigned xr_smem(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void xr_bar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(xr_smem(bar)),
        "r"(parity)
        : "memory");
}

// relative window centre, engine/utils.py:126-128 (the same arithmetic as extract.cu::write_centers)
__device__ __forceinline__ void xr_write_centers(const XrParams& p, int w) {
    const long long gi = p.first_window + w;
    int n = static_cast<int>(gi % p.n_local);
    const int iw = n % p.ns[2];
    n /= p.ns[2];
    const int idx[3] = {n / p.ns[1], n % p.ns[1], iw};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int stop = p.starts[a][idx[a]] + p.roi[a];
        p.centers[w * 3 + a] = static_cast<float>(static_cast<double>(stop - p.roi[a] / 2) / static_cast<double>(p.img[a]));
    }
}

__global__ void __launch_bounds__(32) extract_rows_kernel(const __grid_constant__ XrParams p) {
    extern __shared__ __align__(128) float ring[];  // [kXrStages][ns_w][TR][roi_w]
    __shared__ __align__(8) uint64_t full[kXrStages];
    __shared__ float* s_out[kXrMaxWin];   // the window's patch of this channel
    __shared__ int s_sd[kXrMaxWin], s_sh[kXrMaxWin], s_iw[kXrMaxWin];
    const int lane = threadIdx.x;

    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && p.centers != nullptr)
        for (int w = lane; w < p.n_windows; w += 32) xr_write_centers(p, w);

    const int sh = blockIdx.y;
    const int sd = blockIdx.z % p.n_seg[0];
    const int bc = blockIdx.z / p.n_seg[0];  // volume * channels + channel
    const int b = bc / p.n_channels, c = bc - b * p.n_channels;
    const int n1 = p.seg_n[1][sh];
    const int rows_cell = p.seg_n[0][sd] * n1;
    const int TR = p.tr;
    const int tiles = (rows_cell + TR - 1) / TR;
    const int tile0 = blockIdx.x * p.tiles_per_cta;
    if (tile0 >= tiles) return;
    const int ntile = min(p.tiles_per_cta, tiles - tile0);
    const int rd = p.roi[0], rh = p.roi[1], rw = p.roi[2];
    const long long R = static_cast<long long>(rd) * rh * rw;
    const int d_lo = p.seg_lo[0][sd], h_lo = p.seg_lo[1][sh];
    const int nw = p.ns[2];
    const unsigned row_bytes = static_cast<unsigned>(rw) * 4u;
    const int plane_floats = TR * rw, stage_floats = nw * plane_floats;

    // ---- the cell's windows inside [first_window, first_window + n_windows), ascending ------------------------------
    const int dw0 = p.seg_w0[0][sd], hw0 = p.seg_w0[1][sh];
    const int nh = p.seg_w1[1][sh] - hw0;
    const int ncand = (p.seg_w1[0][sd] - dw0) * nh * nw;  // <= kXrMaxWin (host)
    int nwin = 0;
    unsigned w_used = 0;  // W positions some window of the range uses: only those are loaded
    for (int c0 = 0; c0 < ncand; c0 += 32) {  // (<= 64 candidates: two passes of the warp)
        const int cand = c0 + lane;
        bool mine = false;
        long long gi = 0;
        int id = 0, ih = 0, iw = 0;
        if (cand < ncand) {
            iw = cand % nw, ih = hw0 + (cand / nw) % nh, id = dw0 + cand / (nw * nh);
            gi = static_cast<long long>(b) * p.n_local + (static_cast<long long>(id) * p.ns[1] + ih) * p.ns[2] + iw;
            mine = gi >= p.first_window && gi < p.first_window + p.n_windows;
        }
        const unsigned in_range = __ballot_sync(0xffffffffu, mine);
        if (mine) {
            const int j = nwin + __popc(in_range & ((1u << lane) - 1u));
            s_out[j] = p.out + ((gi - p.first_window) * p.n_channels + c) * R;
            s_sd[j] = p.starts[0][id];
            s_sh[j] = p.starts[1][ih];
            s_iw[j] = iw;
        }
        for (int l = 0; l < 32; ++l)
            if (in_range >> l & 1u) w_used |= 1u << ((c0 + l) % nw);
        nwin += __popc(in_range);
    }
    if (nwin == 0) return;  // nothing of this call lands in this cell
    const int n_wused = __popc(w_used);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kXrStages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xr_smem(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    const float inv_n1 = 1.f / static_cast<float>(n1);
    auto plane_of = [&](int rho) {  // rho / n1 (rows_cell < 2^22: the float quotient is exact after the fix-up)
        int q = __float2int_rz((static_cast<float>(rho) + 0.5f) * inv_n1);
        q -= q * n1 > rho ? 1 : 0;
        q += (q + 1) * n1 <= rho ? 1 : 0;
        return q;
    };
    const float* vol_c = p.vol + static_cast<long long>(bc) * p.vext[0] * p.vext[1] * p.vext[2];
    const long long vplane = static_cast<long long>(p.vext[1]) * p.vext[2];

    auto load_tile = [&](int t, int s) {  // every lane: rows x used W positions, one bulk copy each
        const int rho0 = (tile0 + t) * TR;
        const int nrows = min(TR, rows_cell - rho0);
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xr_smem(&full[s])),
                         "r"(static_cast<unsigned>(nrows * n_wused) * row_bytes)
                         : "memory");
        __syncwarp();
        float* st = ring + static_cast<size_t>(s) * stage_floats;
        for (int i = lane; i < nrows * nw; i += 32) {
            const int wi = i / nrows, r = i - wi * nrows;
            if (!(w_used >> wi & 1u)) continue;
            const int rho = rho0 + r, pl = plane_of(rho);
            const int d = d_lo + pl - p.vorg[0], h = h_lo + (rho - pl * n1) - p.vorg[1], w = p.starts[2][wi] - p.vorg[2];
            const float* src = vol_c + d * vplane + static_cast<long long>(h) * p.vext[2] + w;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             xr_smem(st + wi * plane_floats + r * rw)),
                         "l"(src), "r"(row_bytes), "r"(xr_smem(&full[s]))
                         : "memory");
        }
    };
    auto store_tile = [&](int t, int s) {  // every lane: windows x plane-runs, one bulk store each
        const int rho0 = (tile0 + t) * TR;
        const int nrows = min(TR, rows_cell - rho0);
        const int pl0 = plane_of(rho0);
        const float* st = ring + static_cast<size_t>(s) * stage_floats;
        for (int i = lane; i < nwin * kXrMaxRuns; i += 32) {
            const int j = i / kXrMaxRuns, u = i - j * kXrMaxRuns;
            const int pl = pl0 + u;
            const int ra = max(rho0, pl * n1), rb = min(rho0 + nrows, (pl + 1) * n1);
            if (rb <= ra) continue;
            const int d = d_lo + pl, h = h_lo + (ra - pl * n1);
            float* dst = s_out[j] + (static_cast<long long>(d - s_sd[j]) * rh + (h - s_sh[j])) * rw;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                         "r"(xr_smem(st + s_iw[j] * plane_floats + (ra - rho0) * rw)), "r"(static_cast<unsigned>(rb - ra) * row_bytes)
                         : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");  // (one group per lane and tile, possibly empty)
    };

    load_tile(0, 0);
    int s = 0;
    unsigned phase = 0;
    for (int t = 0; t < ntile; ++t) {
        if (t + 1 < ntile) {
            // the stage of tile t + 1 was last read by the stores of tile t - 2: at most the stores of tile t - 1 may be pending
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            load_tile(t + 1, s + 1 == kXrStages ? 0 : s + 1);
        }
        xr_bar_wait(&full[s], phase);
        store_tile(t, s);
        if (++s == kXrStages) s = 0, phase ^= 1u;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the stores have left shared memory and reached global memory
}

// Returns 0 when the launch was made, < 0 when the geometry is not this kernel's (the caller goes on to the others).
int launch_extract_rows(const mss_layout_t* lay, const Geo& g, const float* volume, const int32_t vol_origin[3],
                        const int32_t vol_extent[3], int n_channels, long long first_window, int n_windows, float* patches_out,
                        float* centers_out, cudaStream_t s, cudaError_t* err) {
    *err = cudaSuccess;
    for (int a = 0; a < 3; ++a)
        if (g.wlo[a] != 0 || g.whi[a] != g.ns[a] || g.ns[a] > kXrMaxSeg) return -1;  // the whole window grid
    if (g.ns[2] > kXrMaxWinW) return -1;
    XrParams p;
    p.vol = volume;
    p.out = patches_out;
    p.centers = centers_out;
    p.first_window = first_window;
    p.n_windows = n_windows;
    p.n_channels = n_channels;
    p.nb = g.nb;
    p.n_local = g.n_local;
    const int32_t* t = lay->table_host;
    int cover_max[2] = {0, 0}, rows_max[2] = {0, 0};
    for (int a = 0; a < 3; ++a) {
        p.vorg[a] = vol_origin[a];
        p.vext[a] = vol_extent[a];
        p.roi[a] = g.roi[a];
        p.img[a] = g.img[a];
        p.ns[a] = g.ns[a];
        const int32_t* st = t + t[kHdrOffStarts + a];
        for (int i = 0; i < g.ns[a]; ++i) p.starts[a][i] = st[i];
        if (a == 2) continue;
        // segments between window starts / ends over the span the windows cover
        const int lo0 = st[0], hi0 = st[g.ns[a] - 1] + g.roi[a];
        int bp[2 * kXrMaxSeg + 2], n = 0;
        bp[n++] = lo0;
        bp[n++] = hi0;
        for (int i = 0; i < g.ns[a]; ++i) {
            const int v[2] = {st[i], st[i] + g.roi[a]};
            for (int e = 0; e < 2; ++e)
                if (v[e] > lo0 && v[e] < hi0) bp[n++] = v[e];
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
        if (m - 1 > kXrMaxSeg) return -1;
        p.n_seg[a] = m - 1;
        for (int i = 0; i + 1 < m; ++i) {
            p.seg_lo[a][i] = bp[i];
            p.seg_n[a][i] = bp[i + 1] - bp[i];
            int lo = g.ns[a], hi = 0;
            for (int w = 0; w < g.ns[a]; ++w)
                if (st[w] <= bp[i] && bp[i] < st[w] + g.roi[a]) {
                    lo = w < lo ? w : lo;
                    hi = w + 1 > hi ? w + 1 : hi;
                }
            if (hi <= lo) return -1;  // a gap between windows (overlap < 0 never happens; be safe)
            for (int w = lo; w < hi; ++w)
                if (!(st[w] <= bp[i] && bp[i + 1] <= st[w] + g.roi[a])) return -1;
            p.seg_w0[a][i] = lo;
            p.seg_w1[a][i] = hi;
            cover_max[a] = hi - lo > cover_max[a] ? hi - lo : cover_max[a];
            rows_max[a] = p.seg_n[a][i] > rows_max[a] ? p.seg_n[a][i] : rows_max[a];
        }
    }
    if (cover_max[0] * cover_max[1] * g.ns[2] > kXrMaxWin) return -1;
    if (static_cast<long long>(rows_max[0]) * rows_max[1] >= (1 << 22)) return -1;
    if (static_cast<long long>(vol_extent[0]) * vol_extent[1] * vol_extent[2] >= (1LL << 31)) return -1;  // 32-bit row offsets
    static const int force_tr = getenv("MSS_XROWS_TR") ? atoi(getenv("MSS_XROWS_TR")) : 0;  // tuning knobs
    static const int force_tpc = getenv("MSS_XROWS_TPC") ? atoi(getenv("MSS_XROWS_TPC")) : 0;
    int tr = force_tr > 0 ? force_tr : 4;  // (cfg2, 296 windows: 2 / 4 / 8 / 16 rows per tile = 0.213 / 0.207 / 0.222 / 0.235 ms)
    int n1_min = 1 << 30;
    for (int i = 0; i < p.n_seg[1]; ++i) n1_min = p.seg_n[1][i] < n1_min ? p.seg_n[1][i] : n1_min;
    while (tr > 1 && (tr - 1) / n1_min + 2 > kXrMaxRuns) --tr;
    p.tr = tr;
    const size_t smem = static_cast<size_t>(kXrStages) * g.ns[2] * tr * g.roi[2] * sizeof(float);
    if (smem > 200 * 1024) return -1;
    const int rows_cell_max = rows_max[0] * rows_max[1];
    const int tiles_cell = (rows_cell_max + tr - 1) / tr;
    p.tiles_per_cta = force_tpc > 0 ? force_tpc : 4;  // (1 / 2 / 4 / 8 / 16 tiles per CTA: 0.219 / 0.207 / 0.208 / 0.214 / 0.222 ms)
    const long long nx = (tiles_cell + p.tiles_per_cta - 1) / p.tiles_per_cta;
    const long long nz = static_cast<long long>(g.nb) * n_channels * p.n_seg[0];
    if (nx <= 0 || nx > 0x7fffffffLL || p.n_seg[1] > 65535 || nz > 65535) return -1;
    cudaError_t e = cudaFuncSetAttribute(extract_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));  // per device
    if (e != cudaSuccess) {
        *err = e;
        return 0;
    }
    const dim3 grid(static_cast<unsigned>(nx), static_cast<unsigned>(p.n_seg[1]), static_cast<unsigned>(nz));
    extract_rows_kernel<<<grid, 32, smem, s>>>(p);
    *err = cudaGetLastError();
    return 0;
}

}  // namespace mss
