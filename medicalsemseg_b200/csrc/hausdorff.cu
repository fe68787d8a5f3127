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

// grid: x over the (y, x-chunk) pairs of one plane of the box, y over its planes; a thread walks kEdgeRun voxels of a
// row with one set of row pointers
constexpr int kEdgeRun = 8;

__global__ void __launch_bounds__(256) mask_edges_kernel(const __grid_constant__ EdgeParams p) {
    const long long sy = p.dims[2], sz = static_cast<long long>(p.dims[1]) * p.dims[2];
    const int chunks = (p.n[2] + kEdgeRun - 1) / kEdgeRun;
    const int work = p.n[1] * chunks;
    const unsigned cls = static_cast<unsigned>(p.cls);
    for (int z = blockIdx.y; z < p.n[0]; z += gridDim.y)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < work; i += gridDim.x * blockDim.x) {
            const int y = i / chunks, x0 = (i - y * chunks) * kEdgeRun;
            const uint8_t* row = p.labels + (p.lo[0] + z) * sz + (p.lo[1] + y) * sy + p.lo[2];
            const long long o = (static_cast<long long>(z) * p.n[1] + y) * p.n[2];
            const int x1 = min(x0 + kEdgeRun, p.n[2]);
            for (int x = x0; x < x1; ++x) {
                const bool edge = mask_edge_at(row + x, sz, sy, z, y, x, p.n, cls);
                p.edges[o + x] = edge ? 1 : 0;
                if (p.h != nullptr) p.h[o + x] = edge ? 0 : kEdtInf;
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

__global__ void __launch_bounds__(128) edt_pass_kernel(const __grid_constant__ EdtParams p) {
    const long long sy = p.n[2], sz = static_cast<long long>(p.n[1]) * p.n[2];
    long long lines, stride;
    int len;
    if (p.axis == 0) lines = sz, stride = sz, len = p.n[0];
    else if (p.axis == 1) lines = static_cast<long long>(p.n[0]) * p.n[2], stride = sy, len = p.n[1];
    else lines = static_cast<long long>(p.n[0]) * p.n[1], stride = 1, len = p.n[2];
    for (long long l = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; l < lines;
         l += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long o;
        if (p.axis == 0) o = l;
        else if (p.axis == 1) o = (l / p.n[2]) * sz + (l % p.n[2]);
        else o = l * p.n[2];
        if (p.mask != nullptr) edt_line_mask<long long>(p.mask + o, p.out + o, p.s + o, p.t + o, len, stride);
        else edt_line<long long>(p.in + o, p.out + o, p.s + o, p.t + o, len, stride);
    }
}

}  // namespace mss

using namespace mss;

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
    edt_pass_kernel<<<static_cast<unsigned>(blocks), 128, 0, as_stream(stream)>>>(p);
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
