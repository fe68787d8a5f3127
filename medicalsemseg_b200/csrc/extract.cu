// ROI patch extraction (replaces the per-window slice + torch.cat of engine/utils.py:122-133 and the
// per-window centre tensors of :126-130).
//
// TMA path: the volume [Nb*Cin, D, H, W] and the patch batch [B*Cin, rd, rh, rw] are described by two
// rank-4 tensor maps.  Each CTA is one warp whose elected lane drives a ring of shared-memory stages:
// cp.async.bulk.tensor (global -> smem, mbarrier complete_tx) followed by cp.async.bulk.tensor
// (smem -> global, bulk_group).  No thread ever touches the data; the SMs only issue descriptors.
// TMA wants every box to start on a 16-byte boundary of the innermost dimension, so it serves the groups whose
// W-starts are all multiples of 4 (every regular start of a 96^3 / overlap-.5 grid; a clamped last start such as
// BraTS' 155 - 96 = 59 is not).  Those groups take the shifted-vector kernel: two aligned 16-byte loads per
// 16-byte store, the second one an L1 hit on the neighbour's first, re-assembled in registers.
// Scalar fallback (row pitch or roi_w not a multiple of 4, unaligned base, or a window that reaches into the
// reference's constant pad): plain predicated loads, 16-byte stores where the layout allows.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace mss {

constexpr int kExtractRowsAuto = 1;  // use_tma = 1 (auto) prefers the volume-stationary kernel where it applies (0.207 vs 0.253 ms)
constexpr int kTmaStages = 4;
constexpr int kTmaBoxH = 16;
constexpr int kTmaBoxD = 2;

struct ExtractParams {
    Geo g;
    long long first_window;
    int n_windows;
    int n_channels;
    int vorg[3];  // stitched-frame coordinate of volume voxel (0,0,0)
    int vext[3];  // volume dims
    float cval;
    float* centers;
};

// relative window centre, engine/utils.py:126-128: (slice.stop - roi//2) / image_size evaluated in
// double (Python floats) and rounded once to float32 by torch.tensor(...)
__device__ __forceinline__ void write_centers(const ExtractParams& p, int w) {
    int b, id, ih, iw;
    decode_window(p.g, p.first_window + w, b, id, ih, iw);
    const int idx[3] = {id, ih, iw};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int stop = p.g.starts[a][idx[a]] + p.g.roi[a];
        p.centers[w * 3 + a] =
            static_cast<float>(static_cast<double>(stop - p.g.roi[a] / 2) / static_cast<double>(p.g.img[a]));
    }
}

// ---- TMA path ---------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// L2 policies: the volume is read by up to 8 overlapping windows and should stay in the 126 MB L2 (evict_last); the patches
// are written once and read much later by the backbone (evict_first), so 1 GB of stores does not flush the volume out
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3,
                                             uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;" ::"l"(map),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// tile t of this launch -> (window in batch, channel, d-tile, h-tile); coordinates for both maps
struct TileCoord {
    int in_c[4];
    int out_c[4];
};

__device__ __forceinline__ TileCoord tile_coord(const ExtractParams& p, long long t, int tiles_h, int tiles_d) {
    const int th = static_cast<int>(t % tiles_h);
    t /= tiles_h;
    const int td = static_cast<int>(t % tiles_d);
    t /= tiles_d;
    const int c = static_cast<int>(t % p.n_channels);
    const int w = static_cast<int>(t / p.n_channels);
    int b, id, ih, iw;
    decode_window(p.g, p.first_window + w, b, id, ih, iw);
    TileCoord tc;
    tc.in_c[0] = p.g.starts[2][iw] - p.vorg[2];
    tc.in_c[1] = p.g.starts[1][ih] - p.vorg[1] + th * kTmaBoxH;
    tc.in_c[2] = p.g.starts[0][id] - p.vorg[0] + td * kTmaBoxD;
    tc.in_c[3] = b * p.n_channels + c;
    tc.out_c[0] = 0;
    tc.out_c[1] = th * kTmaBoxH;
    tc.out_c[2] = td * kTmaBoxD;
    tc.out_c[3] = w * p.n_channels + c;
    return tc;
}

constexpr int kCoordRing = 64;  // tile coordinates computed ahead by the whole warp

__global__ void __launch_bounds__(32) extract_tma_kernel(const __grid_constant__ CUtensorMap in_map,
                                                         const __grid_constant__ CUtensorMap out_map,
                                                         const ExtractParams p, int stage_bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full[kTmaStages];
    __shared__ TileCoord s_tc[kCoordRing];
    const int lane = threadIdx.x;

    if (blockIdx.x == 0 && p.centers != nullptr)
        for (int w = lane; w < p.n_windows; w += 32) write_centers(p, w);

    const int tiles_h = (p.g.roi[1] + kTmaBoxH - 1) / kTmaBoxH;
    const int tiles_d = (p.g.roi[0] + kTmaBoxD - 1) / kTmaBoxD;
    const long long n_tiles = static_cast<long long>(p.n_windows) * p.n_channels * tiles_d * tiles_h;
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const long long mine = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (mine == 0) return;

    if (lane == 0) {
        for (int s = 0; s < kTmaStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // Where a tile comes from and goes to (window decode: divisions and three dependent loads of window starts) is worked
    // out by ALL 32 lanes, 32 tiles at a time, one batch ahead of the elected lane that drives the copies - the first
    // version did it in the elected lane, twice per tile, and spent ~1.8 us per 12 KB tile there (0.70 of the roofline).
    auto fill = [&](long long first) {  // coordinates of tiles [first, first + 32) of this CTA into the ring
        const long long i = first + lane;
        if (i < mine) s_tc[i % kCoordRing] = tile_coord(p, blockIdx.x + i * gridDim.x, tiles_h, tiles_d);
    };
    fill(0);
    fill(32);
    __syncwarp();

    // the rows a box pulls that lie beyond roi_h / roi_d are never stored (the store clips at the patch
    // tensor's extent), so the byte count of every load is the full box
    const uint32_t tx = static_cast<uint32_t>(stage_bytes);
    const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
    auto issue_load = [&](long long i) {
        const int s = static_cast<int>(i % kTmaStages);
        const TileCoord tc = s_tc[i % kCoordRing];
        mbar_expect_tx(&full[s], tx);
        tma_load_4d(smem + static_cast<size_t>(s) * stage_bytes, &in_map, &full[s], tc.in_c[0], tc.in_c[1], tc.in_c[2],
                    tc.in_c[3], pol_keep);
    };

    if (lane == 0) {
        const long long prologue = mine < kTmaStages ? mine : kTmaStages;
        for (long long i = 0; i < prologue; ++i) issue_load(i);
    }
    for (long long base = 0; base < mine; base += 32) {
        if (lane == 0) {
            const long long end = base + 32 < mine ? base + 32 : mine;
            for (long long i = base; i < end; ++i) {
                const int s = static_cast<int>(i % kTmaStages);
                mbar_wait(&full[s], static_cast<uint32_t>((i / kTmaStages) & 1));
                const TileCoord tc = s_tc[i % kCoordRing];
                tma_store_4d(&out_map, smem + static_cast<size_t>(s) * stage_bytes, tc.out_c[0], tc.out_c[1], tc.out_c[2],
                             tc.out_c[3], pol_stream);
                tma_store_commit();
                // refill the stage whose store was issued one iteration ago: allow only the newest store to be pending
                if (i >= 1 && i - 1 + kTmaStages < mine) {
                    tma_store_wait_read<1>();
                    issue_load(i - 1 + kTmaStages);  // <= base + 30 + kTmaStages < base + 64: inside the ring
                }
            }
        }
        __syncwarp();
        fill(base + kCoordRing);  // overwrites the batch just finished
        __syncwarp();
    }
    if (lane == 0) tma_store_wait_read<0>();  // smem must outlive the last bulk reads
}

// ---- fallback path ------------------------------------------------------------------------------

template <bool VEC4>
__global__ void __launch_bounds__(256) extract_plain_kernel(const float* __restrict__ vol, float* __restrict__ out,
                                                            const ExtractParams p) {
    if (blockIdx.x == 0 && p.centers != nullptr)
        for (int w = threadIdx.x; w < p.n_windows; w += blockDim.x) write_centers(p, w);

    constexpr int E = VEC4 ? 4 : 1;
    const int rd = p.g.roi[0], rh = p.g.roi[1], rw = p.g.roi[2];
    const int rwq = rw / E;
    const long long per_win = static_cast<long long>(p.n_channels) * rd * rh * rwq;
    const long long total = per_win * p.n_windows;
    const long long plane = static_cast<long long>(p.vext[1]) * p.vext[2];
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long r = e;
        const int lq = static_cast<int>(r % rwq);
        r /= rwq;
        const int lh = static_cast<int>(r % rh);
        r /= rh;
        const int ld = static_cast<int>(r % rd);
        r /= rd;
        const int c = static_cast<int>(r % p.n_channels);
        const int w = static_cast<int>(r / p.n_channels);
        int b, id, ih, iw;
        decode_window(p.g, p.first_window + w, b, id, ih, iw);
        const int vd = p.g.starts[0][id] + ld - p.vorg[0];
        const int vh = p.g.starts[1][ih] + lh - p.vorg[1];
        const int vw = p.g.starts[2][iw] + lq * E - p.vorg[2];
        const bool row_in = vd >= 0 && vd < p.vext[0] && vh >= 0 && vh < p.vext[1];
        const float* src = vol + (static_cast<long long>(b) * p.n_channels + c) * p.vext[0] * plane +
                           static_cast<long long>(vd) * plane + static_cast<long long>(vh) * p.vext[2];
        float v[E];
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const int x = vw + k;
            v[k] = (row_in && x >= 0 && x < p.vext[2]) ? __ldg(src + x) : p.cval;
        }
        if (VEC4)
            *reinterpret_cast<float4*>(out + e * 4) = make_float4(v[0], v[E > 1 ? 1 : 0], v[E > 2 ? 2 : 0], v[E > 3 ? 3 : 0]);
        else
            out[e] = v[0];
    }
}

// ---- shifted-vector path: aligned rows, window starts at any W offset ------------------------------
// requires: all windows inside the volume, vext[2] % 4 == 0, roi_w % 4 == 0, 16-byte aligned bases
// grid: x = 256-thread tiles over one (rh x rw/4) plane, y = groups of kShiftPlanes patch planes, z = (window, channel);
// a thread copies the same quad of kShiftPlanes consecutive planes (independent loads, all in flight together)
constexpr int kShiftPlanes = 8;

__global__ void __launch_bounds__(256) extract_shifted_kernel(const float* __restrict__ vol, float* __restrict__ out,
                                                              const ExtractParams p) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && p.centers != nullptr)
        for (int w = threadIdx.x; w < p.n_windows; w += blockDim.x) write_centers(p, w);
    const int rd = p.g.roi[0], rh = p.g.roi[1], rwq = p.g.roi[2] / 4;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= rh * rwq) return;
    const int lh = idx / rwq, lq = idx - lh * rwq;
    const int ld0 = blockIdx.y * kShiftPlanes;
    const int w = blockIdx.z / p.n_channels, c = blockIdx.z - w * p.n_channels;
    int b, id, ih, iw;
    decode_window(p.g, p.first_window + w, b, id, ih, iw);  // uniform per block
    const int vd = p.g.starts[0][id] + ld0 - p.vorg[0];
    const int vh = p.g.starts[1][ih] + lh - p.vorg[1];
    const int vw = p.g.starts[2][iw] + lq * 4 - p.vorg[2];
    const int sh = vw & 3;  // uniform per window
    const long long plane = static_cast<long long>(p.vext[1]) * p.vext[2];
    const float* src = vol + (static_cast<long long>(b) * p.n_channels + c) * p.vext[0] * plane +
                       static_cast<long long>(vd) * plane + static_cast<long long>(vh) * p.vext[2] + (vw - sh);
    float* dst = out + ((static_cast<long long>(blockIdx.z) * rd + ld0) * rh + lh) * (rwq * 4) + lq * 4;
    const long long dst_plane = static_cast<long long>(rh) * rwq * 4;
    float4 a[kShiftPlanes], n[kShiftPlanes];
#pragma unroll
    for (int d = 0; d < kShiftPlanes; ++d)
        if (ld0 + d < rd) {
            a[d] = ld_stream_f4(src + d * plane);
            if (sh != 0) n[d] = __ldg(reinterpret_cast<const float4*>(src + d * plane + 4));  // inside the padded row
        }
#pragma unroll
    for (int d = 0; d < kShiftPlanes; ++d)
        if (ld0 + d < rd) {
            float4 v = a[d];
            if (sh == 1) v = make_float4(a[d].y, a[d].z, a[d].w, n[d].x);
            else if (sh == 2) v = make_float4(a[d].z, a[d].w, n[d].x, n[d].y);
            else if (sh == 3) v = make_float4(a[d].w, n[d].x, n[d].y, n[d].z);
            *reinterpret_cast<float4*>(dst + d * dst_plane) = v;
        }
}

// ---- host side ----------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

static int encode_map(CUtensorMap* map, const void* base, const long long dims[4], const int box[4]) {
    EncodeTiledFn fn = encode_tiled_fn();
    MSS_REQUIRE(fn != nullptr, MSS_E_DRIVER, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(dims[0]), static_cast<cuuint64_t>(dims[1]),
                          static_cast<cuuint64_t>(dims[2]), static_cast<cuuint64_t>(dims[3])};
    cuuint64_t gstr[3] = {gdim[0] * 4, gdim[0] * gdim[1] * 4, gdim[0] * gdim[1] * gdim[2] * 4};
    cuuint32_t bdim[4] = {static_cast<cuuint32_t>(box[0]), static_cast<cuuint32_t>(box[1]), static_cast<cuuint32_t>(box[2]),
                          static_cast<cuuint32_t>(box[3])};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gdim, gstr, bdim, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSS_REQUIRE(r == CUDA_SUCCESS, MSS_E_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return MSS_OK;
}

// extract_rows.cu: the volume-stationary kernel
int launch_extract_rows(const mss_layout_t* lay, const Geo& g, const float* volume, const int32_t vol_origin[3],
                        const int32_t vol_extent[3], int n_channels, long long first_window, int n_windows, float* patches_out,
                        float* centers_out, cudaStream_t s, cudaError_t* err);

}  // namespace mss

using namespace mss;

extern "C" int mss_extract_patches(const float* volume, const int32_t vol_origin[3], const int32_t vol_extent[3],
                                   int32_t n_channels, float cval, const mss_layout_t* lay, int64_t first_window,
                                   int32_t n_windows, float* patches_out, float* centers_out, int32_t use_tma,
                                   void* stream) {
    MSS_REQUIRE(volume && vol_origin && vol_extent && patches_out, MSS_E_ARG, "extract_patches: null argument");
    ExtractParams p;
    int rc = make_geo(lay, &p.g);
    if (rc != MSS_OK) return rc;
    MSS_REQUIRE(n_channels > 0 && n_windows > 0, MSS_E_ARG, "extract_patches: need n_channels > 0 and n_windows > 0");
    const long long total_windows = p.g.n_local * p.g.nb;
    MSS_REQUIRE(first_window >= 0 && first_window + n_windows <= total_windows, MSS_E_ARG,
                "extract_patches: windows [%lld, +%d) outside [0, %lld)", static_cast<long long>(first_window), n_windows,
                total_windows);
    p.first_window = first_window;
    p.n_windows = n_windows;
    p.n_channels = n_channels;
    p.cval = cval;
    p.centers = centers_out;
    bool inside = true;  // do all owned windows lie inside the real volume (no constant pad involved)?
    const int32_t* t = lay->table_host;
    for (int a = 0; a < 3; ++a) {
        MSS_REQUIRE(vol_extent[a] > 0, MSS_E_ARG, "extract_patches: volume extent must be positive");
        p.vorg[a] = vol_origin[a];
        p.vext[a] = vol_extent[a];
        const int32_t* st = t + t[kHdrOffStarts + a];
        if (st[lay->win_lo[a]] < vol_origin[a] || st[lay->win_hi[a] - 1] + lay->roi[a] > vol_origin[a] + vol_extent[a])
            inside = false;
    }
    cudaStream_t s = as_stream(stream);
    const int rw = p.g.roi[2];
    // vector paths: all windows inside the volume, rows and patches 16-byte tiled
    const bool vec_layout = inside && (vol_extent[2] % 4 == 0) && (rw % 4 == 0) &&
                            (reinterpret_cast<uintptr_t>(volume) % 16 == 0) &&
                            (reinterpret_cast<uintptr_t>(patches_out) % 16 == 0);
    bool starts_aligned = true;  // TMA boxes must start on a 16-byte boundary of the innermost dimension
    {
        const int32_t* st = t + t[kHdrOffStarts + 2];
        for (int i = lay->win_lo[2]; i < lay->win_hi[2]; ++i)
            if ((st[i] - vol_origin[2]) % 4 != 0) starts_aligned = false;
    }
    // volume-stationary kernel (extract_rows.cu): every volume row read once, written to all its windows; use_tma 3 asks for
    // it, 1 (auto) takes it when MSS_EXTRACT_ROWS is not 0
    {
        static const int rows_auto = getenv("MSS_EXTRACT_ROWS") ? atoi(getenv("MSS_EXTRACT_ROWS")) : kExtractRowsAuto;
        if ((use_tma == 3 || (use_tma == 1 && rows_auto)) && vec_layout && starts_aligned) {
            cudaError_t cerr = cudaSuccess;
            if (launch_extract_rows(lay, p.g, volume, vol_origin, vol_extent, n_channels, first_window, n_windows, patches_out,
                                    centers_out, s, &cerr) == 0) {
                MSS_CUDA(cerr);
                return MSS_OK;
            }
        }
        if (use_tma == 3) use_tma = 1;  // not this kernel's geometry: auto
    }
    // the box {rw, kTmaBoxH, kTmaBoxD, 1} must fit the patch tensor; a failed descriptor encode (driver / shape the
    // driver rejects) falls through to the shifted-vector kernel, which serves the same layout
    bool tma_ok = use_tma == 1 && vec_layout && starts_aligned && rw <= 256 && p.g.roi[1] >= kTmaBoxH &&
                  p.g.roi[0] >= kTmaBoxD && encode_tiled_fn() != nullptr;
    CUtensorMap in_map, out_map;
    if (tma_ok) {
        const long long in_dims[4] = {vol_extent[2], vol_extent[1], vol_extent[0],
                                      static_cast<long long>(p.g.nb) * n_channels};
        const long long out_dims[4] = {rw, p.g.roi[1], p.g.roi[0], static_cast<long long>(n_windows) * n_channels};
        const int box[4] = {rw, kTmaBoxH, kTmaBoxD, 1};
        if (encode_map(&in_map, volume, in_dims, box) != MSS_OK || encode_map(&out_map, patches_out, out_dims, box) != MSS_OK)
            tma_ok = false;
    }
    if (tma_ok) {
        const int stage_bytes = rw * kTmaBoxH * kTmaBoxD * 4;
        const size_t smem = static_cast<size_t>(stage_bytes) * kTmaStages;
        MSS_CUDA(cudaFuncSetAttribute(extract_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));  // per device
        const long long n_tiles = static_cast<long long>(n_windows) * n_channels * ((p.g.roi[0] + kTmaBoxD - 1) / kTmaBoxD) *
                                  ((p.g.roi[1] + kTmaBoxH - 1) / kTmaBoxH);
        int ctas_per_sm = static_cast<int>((200 * 1024) / (smem + 1024));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        if (ctas_per_sm > 8) ctas_per_sm = 8;
        long long grid = 148LL * ctas_per_sm;
        if (grid > n_tiles) grid = n_tiles;
        extract_tma_kernel<<<static_cast<unsigned>(grid), 32, smem, s>>>(in_map, out_map, p, stage_bytes);
        MSS_CUDA(cudaGetLastError());
        return MSS_OK;
    }
    const long long elems = static_cast<long long>(n_windows) * n_channels * p.g.roi[0] * p.g.roi[1] * rw;
    if (vec_layout && use_tma != 2 && p.g.roi[0] <= 65535 && static_cast<long long>(n_windows) * n_channels <= 65535) {
        dim3 grid(static_cast<unsigned>((p.g.roi[1] * (rw / 4) + 255) / 256),
                  static_cast<unsigned>((p.g.roi[0] + kShiftPlanes - 1) / kShiftPlanes),
                  static_cast<unsigned>(n_windows * n_channels));
        extract_shifted_kernel<<<grid, 256, 0, s>>>(volume, patches_out, p);
        MSS_CUDA(cudaGetLastError());
        return MSS_OK;
    }
    const bool vec = (rw % 4 == 0) && (reinterpret_cast<uintptr_t>(patches_out) % 16 == 0);
    const long long work = vec ? elems / 4 : elems;
    long long blocks = (work + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    if (vec)
        extract_plain_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, s>>>(volume, patches_out, p);
    else
        extract_plain_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, s>>>(volume, patches_out, p);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}
