// Surface extraction and exact Euclidean distance transform for the 95th-percentile Hausdorff distance
// (MONAI HausdorffDistanceMetric(percentile=95) as wired at engine/test.py:31,55-57: get_mask_edges = binary_erosion
// XOR mask on the bounding box of pred | gt, get_surface_distance = scipy distance_transform_edt of the other surface).
// SURVEY.md section 8f rank 4.
//
//   mss_mask_edges : edges[x] = (label[x] == c) and some 6-neighbour (along axes of box extent > 1; outside the box
//                    counts as background - scipy's border_value = 0 on the cropped array) is not c.
//                    Also writes the EDT input h = 0 on edge voxels, kEdtInf elsewhere.
//   mss_edt_pass   : one axis of the exact squared EDT (csrc/edt.cuh), one thread per line; consecutive threads own
//                    lines that are adjacent in memory, so every access of the line loop is coalesced for the two
//                    strided axes (the host transposes before the pass along the contiguous axis).
#include "common.cuh"
#include "edt.cuh"

namespace mss {

struct EdgeParams {
    const uint8_t* labels;
    int dims[3];
    int lo[3], n[3];
    int cls;
    uint8_t* edges;
    int* h;
};

// grid: x over the (y, x-chunk) pairs of one plane of the box, y over its planes.  A thread decides kEdgeRun = 8 voxels
// of a row at once on packed bytes: the five rows it needs (centre, +-1 row, +-1 plane) come in as aligned 32-bit words
// funnel-shifted to the run's first byte, `byte == class` is a zero-byte test of `word ^ class` (the six neighbours are OR-ed
// first, so a word costs one test), the W neighbours are the centre
// row's words shifted by one byte, and a run without a single voxel of the class (most runs of an organ's box) stops
// after the centre row.  ~60 instructions per 8 voxels where the voxel-by-voxel form (mask_edge_at, kept for the first
// and last bytes of the volume and as the host-checked definition) spent ~25 per voxel.
constexpr int kEdgeRun = 8;

// 0x80 in every byte of x that is zero (exact: no carries cross byte lanes)
__device__ __forceinline__ unsigned zero_bytes(unsigned x) {
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}

// bytes [0, 8) at `ptr` (any alignment) as two little-endian words, from aligned word loads
__device__ __forceinline__ uint2 load_run8(const uint8_t* ptr) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(ptr);
    const unsigned* w = reinterpret_cast<const unsigned*>(a & ~static_cast<uintptr_t>(3));
    const unsigned sh = static_cast<unsigned>(a & 3u) * 8u;
    const unsigned w0 = __ldg(w), w1 = __ldg(w + 1), w2 = sh ? __ldg(w + 2) : 0u;
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}

__global__ void __launch_bounds__(256) mask_edges_kernel(const __grid_constant__ EdgeParams p) {
    const long long sy = p.dims[2], sz = static_cast<long long>(p.dims[1]) * p.dims[2];
    const int chunks = (p.n[2] + kEdgeRun - 1) / kEdgeRun;
    const int work = p.n[1] * chunks;
    const unsigned cls = static_cast<unsigned>(p.cls);
    const unsigned cls4 = cls * 0x01010101u;
    const uint8_t* vol_end = p.labels + static_cast<long long>(p.dims[0]) * sz;
    for (int z = blockIdx.y; z < p.n[0]; z += gridDim.y)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < work; i += gridDim.x * blockDim.x) {
            const int y = i / chunks, x0 = (i - y * chunks) * kEdgeRun;
            const uint8_t* row = p.labels + (p.lo[0] + z) * sz + (p.lo[1] + y) * sy + p.lo[2];
            const long long o = (static_cast<long long>(z) * p.n[1] + y) * p.n[2];
            const int x1 = min(x0 + kEdgeRun, p.n[2]);
            const uint8_t* c = row + x0;
            unsigned long long packed = 0ull;  // the run's 8 edge bytes, stored with one 8-byte store where aligned
            // word loads reach up to 4 bytes before / 16 bytes after the run's first byte (of any of the five rows): only
            // inside the volume
            if (p.h == nullptr && c - sz - 4 >= p.labels && c + sz + 16 <= vol_end) {
                // centre row, bytes [-1, 9): left neighbours, the run, right neighbours
                const uintptr_t a = reinterpret_cast<uintptr_t>(c - 1);
                const unsigned* w = reinterpret_cast<const unsigned*>(a & ~static_cast<uintptr_t>(3));
                const unsigned sh = static_cast<unsigned>(a & 3u) * 8u;
                const unsigned w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = sh > 8u ? __ldg(w + 3) : 0u;
                const unsigned lo = __funnelshift_r(w0, w1, sh), mid = __funnelshift_r(w1, w2, sh), hi = __funnelshift_r(w2, w3, sh);
                // validity of the run's bytes (the last run of a row may be short)
                const int nv = x1 - x0;
                const unsigned v0 = nv >= 4 ? 0xFFFFFFFFu : (1u << (8 * nv)) - 1u;
                const unsigned v1 = nv >= 8 ? 0xFFFFFFFFu : (nv > 4 ? (1u << (8 * (nv - 4))) - 1u : 0u);
                // byte == class  <=>  (byte ^ class) == 0; zero_bytes() marks those bytes with 0x80
                const unsigned c0 = zero_bytes(__funnelshift_r(lo, mid, 8) ^ cls4) & v0;
                const unsigned c1 = zero_bytes(__funnelshift_r(mid, hi, 8) ^ cls4) & v1;
                if (c0 | c1) {
                    // t: OR of (neighbour ^ class) over the neighbours that exist - a zero byte = every neighbour is of the
                    // class; a neighbour outside the box is background: its byte is forced non-zero
                    unsigned t0 = 0u, t1 = 0u;
                    if (p.n[2] > 1) {
                        t0 = (lo ^ cls4) | (__funnelshift_r(lo, mid, 16) ^ cls4);
                        t1 = (mid ^ cls4) | (__funnelshift_r(mid, hi, 16) ^ cls4);
                        if (x0 == 0) t0 |= 0xFFu;
                        const int xl = p.n[2] - 1 - x0;  // the voxel at x = n - 1 has no right neighbour inside the box
                        if (xl < 4) t0 |= 0xFFu << (8 * xl);
                        else if (xl < 8) t1 |= 0xFFu << (8 * (xl - 4));
                    }
                    if (p.n[1] > 1) {
                        if (y == 0 || y == p.n[1] - 1) {
                            t0 = t1 = 0xFFFFFFFFu;
                        } else {
                            const uint2 u = load_run8(c - sy), d = load_run8(c + sy);
                            t0 |= (u.x ^ cls4) | (d.x ^ cls4);
                            t1 |= (u.y ^ cls4) | (d.y ^ cls4);
                        }
                    }
                    if (p.n[0] > 1) {
                        if (z == 0 || z == p.n[0] - 1) {
                            t0 = t1 = 0xFFFFFFFFu;
                        } else {
                            const uint2 u = load_run8(c - sz), d = load_run8(c + sz);
                            t0 |= (u.x ^ cls4) | (d.x ^ cls4);
                            t1 |= (u.y ^ cls4) | (d.y ^ cls4);
                        }
                    }
                    const unsigned e0 = (c0 & ~zero_bytes(t0)) >> 7, e1 = (c1 & ~zero_bytes(t1)) >> 7;
                    packed = static_cast<unsigned long long>(e0) | (static_cast<unsigned long long>(e1) << 32);
                }
            } else {
                for (int x = x0; x < x1; ++x) {
                    const bool edge = mask_edge_at(row + x, sz, sy, z, y, x, p.n, cls);
                    packed |= static_cast<unsigned long long>(edge ? 1u : 0u) << (8 * (x - x0));
                    if (p.h != nullptr) p.h[o + x] = edge ? 0 : kEdtInf;
                }
            }
            uint8_t* dst = p.edges + o + x0;
            const unsigned al = static_cast<unsigned>(reinterpret_cast<uintptr_t>(dst)) & 7u;
            if (x1 - x0 == kEdgeRun && al == 0) {
                *reinterpret_cast<unsigned long long*>(dst) = packed;
            } else if (x1 - x0 == kEdgeRun && (al & 3u) == 0) {
                *reinterpret_cast<unsigned*>(dst) = static_cast<unsigned>(packed);
                *reinterpret_cast<unsigned*>(dst + 4) = static_cast<unsigned>(packed >> 32);
            } else {
                for (int x = x0; x < x1; ++x) dst[x - x0] = static_cast<uint8_t>(packed >> (8 * (x - x0)));
            }
        }
}

struct EdtParams {
    const int* in;
    const uint8_t* mask;  // first pass straight from a uint8 feature mask (in == nullptr)
    int* out;
    int* s;
    int* t;
    int n[3];
    int axis;
};

constexpr int kEdtRecipMax = 2048;  // line lengths up to this divide by table (8 KB of shared memory)

// MASK: the line values come from a uint8 feature mask; RECIP: line length <= kEdtRecipMax, the envelope's floor divisions
// run on a shared-memory table of reciprocals (csrc/edt.cuh::edt_div2k) built once per CTA
template <bool MASK, bool RECIP>
__global__ void __launch_bounds__(128) edt_pass_kernel(const __grid_constant__ EdtParams p) {
    __shared__ unsigned s_recip[RECIP ? kEdtRecipMax : 1];
    const long long sy = p.n[2], sz = static_cast<long long>(p.n[1]) * p.n[2];
    long long lines, stride;
    int len;
    if (p.axis == 0) lines = sz, stride = sz, len = p.n[0];
    else if (p.axis == 1) lines = static_cast<long long>(p.n[0]) * p.n[2], stride = sy, len = p.n[1];
    else lines = static_cast<long long>(p.n[0]) * p.n[1], stride = 1, len = p.n[2];
    if (RECIP) {
        for (int k = threadIdx.x; k < len; k += blockDim.x) s_recip[k] = k ? edt_recip(static_cast<unsigned>(k)) : 0u;
        __syncthreads();
    }
    for (long long l = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; l < lines;
         l += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long o;
        if (p.axis == 0) o = l;
        else if (p.axis == 1) o = (l / p.n[2]) * sz + (l % p.n[2]);
        else o = l * p.n[2];
        if (MASK) edt_line_mask_cached<long long>(p.mask + o, p.out + o, p.s + o, p.t + o, len, stride);
        else edt_line_cached<long long>(p.in + o, p.out + o, p.s + o, p.t + o, len, stride, RECIP ? s_recip : nullptr);
    }
}

// First EDT pass along the CONTIGUOUS axis straight from a uint8 feature mask: no envelope needed - the squared distance
// to the nearest feature of the same row.  One warp per row, 32 voxels per step: a ballot of the feature bits gives every
// lane its nearest set position inside the chunk (clz / ffs on the masked ballot), a carried position covers the chunks
// before / after.  Forward sweep stores the left distance, backward sweep combines it with the right one and squares.
__global__ void __launch_bounds__(256) edt_row_mask_kernel(const uint8_t* __restrict__ mask, int* __restrict__ out,
                                                           long long n_rows, int w) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int chunks = (w + 31) >> 5;
    constexpr int kFar = 1 << 20;
    for (long long r = warp; r < n_rows; r += n_warps) {
        const uint8_t* m = mask + r * w;
        int* o = out + r * w;
        int last = -kFar;
        for (int c = 0; c < chunks; ++c) {
            const int x = c * 32 + lane;
            const unsigned bal = __ballot_sync(0xffffffffu, x < w && m[x] != 0);
            const unsigned le = bal & (0xffffffffu >> (31 - lane));  // features at or left of this lane
            const int pos = le ? c * 32 + 31 - __clz(le) : last;
            if (x < w) o[x] = x - pos;
            if (bal) last = c * 32 + 31 - __clz(bal);
        }
        int next = kFar;
        for (int c = chunks - 1; c >= 0; --c) {
            const int x = c * 32 + lane;
            const unsigned bal = __ballot_sync(0xffffffffu, x < w && m[x] != 0);
            const unsigned ge = bal & (0xffffffffu << lane);  // features at or right of this lane
            const int pos = ge ? c * 32 + __ffs(ge) - 1 : next;
            if (x < w) {
                const int dl = o[x], dr = pos - x;
                const long long d = dl < dr ? dl : dr;
                o[x] = d >= (1 << 14) ? kEdtInf : static_cast<int>(d * d);
            }
            if (bal) next = c * 32 + __ffs(bal) - 1;
        }
    }
}

// Bounding boxes of (pred == c) | (label == c) for every class c < K in one pass over both label maps
// (MONAI generate_spatial_bounding_box per class, get_mask_edges): boxes[c] = {lo_d, lo_h, lo_w, hi_d, hi_h, hi_w}
// (hi exclusive; lo > hi for a class that occurs in neither map).  Shared-memory min / max per CTA, then global atomics.
constexpr int kBoxMaxClasses = 32;

__global__ void __launch_bounds__(256) class_boxes_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int d, int h,
                                                          int w, int k, int* __restrict__ boxes) {
    __shared__ int s_lo[kBoxMaxClasses][3], s_hi[kBoxMaxClasses][3];
    for (int i = threadIdx.x; i < kBoxMaxClasses * 3; i += blockDim.x) {
        (&s_lo[0][0])[i] = 1 << 30;
        (&s_hi[0][0])[i] = -1;
    }
    __syncthreads();
    const long long rows = static_cast<long long>(d) * h;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const int z = static_cast<int>(r / h), y = static_cast<int>(r - static_cast<long long>(z) * h);
        const uint8_t* pa = a + r * w;
        const uint8_t* pb = b + r * w;
        unsigned seen = 0;  // classes this thread met in this row, with their W range
        for (int x = threadIdx.x; x < w; x += blockDim.x) {
            const unsigned ca = pa[x], cb = pb[x];
            if (ca < static_cast<unsigned>(k)) {
                seen |= 1u << ca;
                atomicMin(&s_lo[ca][2], x);
                atomicMax(&s_hi[ca][2], x);
            }
            if (cb < static_cast<unsigned>(k) && cb != ca) {
                seen |= 1u << cb;
                atomicMin(&s_lo[cb][2], x);
                atomicMax(&s_hi[cb][2], x);
            }
        }
        while (seen) {
            const int c = __ffs(seen) - 1;
            seen &= seen - 1;
            atomicMin(&s_lo[c][0], z);
            atomicMax(&s_hi[c][0], z);
            atomicMin(&s_lo[c][1], y);
            atomicMax(&s_hi[c][1], y);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k * 3; i += blockDim.x) {
        const int c = i / 3, ax = i - c * 3;
        if (s_hi[c][ax] >= 0) {
            atomicMin(&boxes[c * 6 + ax], s_lo[c][ax]);
            atomicMax(&boxes[c * 6 + 3 + ax], s_hi[c][ax] + 1);
        }
    }
}

}  // namespace mss

using namespace mss;

extern "C" int mss_class_boxes(const uint8_t* pred, const uint8_t* label, const int32_t dims[3], int32_t n_classes,
                               int32_t* boxes_out, void* stream) {
    MSS_REQUIRE(pred && label && dims && boxes_out, MSS_E_ARG, "class_boxes: null argument");
    MSS_REQUIRE(n_classes > 0 && n_classes <= kBoxMaxClasses, MSS_E_UNSUPPORTED, "class_boxes: n_classes %d outside [1, %d]",
                n_classes, kBoxMaxClasses);
    for (int a = 0; a < 3; ++a) MSS_REQUIRE(dims[a] > 0, MSS_E_ARG, "class_boxes: dims must be positive");
    // boxes_out must hold {2^30, 2^30, 2^30, 0, 0, 0} per class on entry (the caller fills it; min / max are accumulated)
    long long rows = static_cast<long long>(dims[0]) * dims[1];
    long long blocks = rows < 148LL * 8 ? rows : 148LL * 8;
    class_boxes_kernel<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(pred, label, dims[0], dims[1], dims[2],
                                                                                  n_classes, boxes_out);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}

extern "C" int mss_edt_row_mask(const uint8_t* mask, int32_t* out, const int32_t dims[3], void* stream) {
    MSS_REQUIRE(mask && out && dims, MSS_E_ARG, "edt_row_mask: null argument");
    for (int a = 0; a < 3; ++a)
        MSS_REQUIRE(dims[a] > 0 && dims[a] < (1 << 14), MSS_E_UNSUPPORTED, "edt_row_mask: dims must be in [1, 16384)");
    const long long rows = static_cast<long long>(dims[0]) * dims[1];
    long long blocks = (rows + 7) / 8;  // 8 warps (rows) per CTA
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    edt_row_mask_kernel<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(mask, out, rows, dims[2]);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}

extern "C" int mss_mask_edges(const uint8_t* labels, const int32_t dims[3], int32_t cls, const int32_t box_lo[3],
                              const int32_t box_hi[3], uint8_t* edges_out, int32_t* edt_input_out, void* stream) {
    MSS_REQUIRE(labels && dims && box_lo && box_hi && edges_out, MSS_E_ARG, "mask_edges: null argument");
    MSS_REQUIRE(cls >= 0 && cls <= 255, MSS_E_ARG, "mask_edges: class %d outside uint8", cls);
    EdgeParams p;
    for (int a = 0; a < 3; ++a) {
        MSS_REQUIRE(dims[a] > 0 && 0 <= box_lo[a] && box_lo[a] < box_hi[a] && box_hi[a] <= dims[a], MSS_E_ARG,
                    "mask_edges: axis %d box [%d,%d) outside the volume (%d)", a, box_lo[a], box_hi[a], dims[a]);
        p.dims[a] = dims[a];
        p.lo[a] = box_lo[a];
        p.n[a] = box_hi[a] - box_lo[a];
    }
    p.labels = labels;
    p.cls = cls;
    p.edges = edges_out;
    p.h = edt_input_out;
    MSS_REQUIRE(static_cast<long long>(p.n[1]) * p.n[2] < (1LL << 31), MSS_E_UNSUPPORTED, "mask_edges: plane too large");
    long long bx = (static_cast<long long>(p.n[1]) * ((p.n[2] + kEdgeRun - 1) / kEdgeRun) + 255) / 256;
    if (bx > 148LL * 4) bx = 148LL * 4;
    const dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(p.n[0] < 65535 ? p.n[0] : 65535));
    mask_edges_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}

static int edt_pass_impl(const int32_t* in, const uint8_t* mask, int32_t* out, int32_t* scratch_s, int32_t* scratch_t,
                         const int32_t dims[3], int32_t axis, void* stream) {
    MSS_REQUIRE((in != nullptr) != (mask != nullptr), MSS_E_ARG, "edt_pass: need exactly one input");
    MSS_REQUIRE(out && scratch_s && scratch_t && dims && in != out, MSS_E_ARG,
                "edt_pass: null argument (or in == out: the pass is not in place)");
    MSS_REQUIRE(axis >= 0 && axis < 3, MSS_E_ARG, "edt_pass: axis %d outside [0, 3)", axis);
    EdtParams p;
    for (int a = 0; a < 3; ++a) {
        MSS_REQUIRE(dims[a] > 0 && dims[a] < (1 << 14), MSS_E_UNSUPPORTED, "edt_pass: dims must be in [1, 16384)");
        p.n[a] = dims[a];
    }
    p.in = in;
    p.mask = mask;
    p.out = out;
    p.s = scratch_s;
    p.t = scratch_t;
    p.axis = axis;
    const long long lines = static_cast<long long>(dims[0]) * dims[1] * dims[2] / dims[axis];
    long long blocks = (lines + 127) / 128;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    const unsigned nb = static_cast<unsigned>(blocks);
    if (mask != nullptr) edt_pass_kernel<true, false><<<nb, 128, 0, as_stream(stream)>>>(p);
    else if (dims[axis] <= kEdtRecipMax) edt_pass_kernel<false, true><<<nb, 128, 0, as_stream(stream)>>>(p);
    else edt_pass_kernel<false, false><<<nb, 128, 0, as_stream(stream)>>>(p);
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}

extern "C" int mss_edt_pass(const int32_t* in, int32_t* out, int32_t* scratch_s, int32_t* scratch_t, const int32_t dims[3],
                            int32_t axis, void* stream) {
    MSS_REQUIRE(in != nullptr, MSS_E_ARG, "edt_pass: null argument");
    return edt_pass_impl(in, nullptr, out, scratch_s, scratch_t, dims, axis, stream);
}

extern "C" int mss_edt_pass_mask(const uint8_t* mask, int32_t* out, int32_t* scratch_s, int32_t* scratch_t,
                                 const int32_t dims[3], int32_t axis, void* stream) {
    MSS_REQUIRE(mask != nullptr, MSS_E_ARG, "edt_pass_mask: null argument");
    return edt_pass_impl(nullptr, mask, out, scratch_s, scratch_t, dims, axis, stream);
}
