// Exact squared Euclidean distance transform, one line at a time (Meijster / Felzenszwalb lower envelope of parabolas),
// shared by the GPU pass kernel (csrc/hausdorff.cu) and the host test harness (tests/csrc/host_chunks.cpp).
//
//   out[x] = min over i of (x - i)^2 + h[i]          (h[i] >= kEdtInf: no parabola at i)
//
// Applied along the three axes in turn, starting from h = 0 on feature voxels and kEdtInf elsewhere, it yields the
// exact integer squared distance to the nearest feature voxel - what scipy.ndimage.distance_transform_edt returns
// after a float64 sqrt (MONAI get_surface_distance, used by HausdorffDistanceMetric at engine/test.py:31,55).
// All arithmetic is integer, so host and device agree bit for bit.
#pragma once

#ifdef __CUDACC__
#define MSS_EDT_HD __host__ __device__ __forceinline__
#else
#define MSS_EDT_HD inline
#endif

namespace mss {

constexpr int kEdtInf = 1 << 29;  // "no feature": larger than any squared distance in a volume of side < 2^14

// s / t: the parabola stack of the line (positions and take-over points), n ints each, strided like the line itself.
// h(u) returns the line's input at position u (a functor, so the first pass can read the uint8 surface mask directly).
template <typename Idx, typename HFn>
MSS_EDT_HD void edt_line_fn(HFn h, int* out, int* s, int* t, int n, Idx stride) {
    int q = -1;  // top of the stack
    for (int u = 0; u < n; ++u) {
        const long long hu = h(u);
        if (hu >= kEdtInf) continue;
        while (q >= 0) {
            const long long i = s[q * stride], hi = h(static_cast<int>(i)), x = t[q * stride];
            // parabola u is at or below parabola i at the point where i took over: i never wins
            if ((x - i) * (x - i) + hi > (x - u) * (x - u) + hu) --q;
            else break;
        }
        if (q < 0) {
            q = 0;
            s[0] = u;
            t[0] = 0;
        } else {
            const long long i = s[q * stride], hi = h(static_cast<int>(i));
            // first integer x where parabola u is strictly below parabola i (u > i):  x > (u^2 - i^2 + hu - hi) / (2 (u - i))
            const long long num = static_cast<long long>(u) * u - i * i + hu - hi, den = 2 * (u - i);
            long long w = num >= 0 ? num / den : -((-num + den - 1) / den);  // floor division
            w += 1;
            if (w < n) {
                ++q;
                s[q * stride] = u;
                t[q * stride] = static_cast<int>(w < 0 ? 0 : w);
            }
        }
    }
    if (q < 0) {
        for (int x = 0; x < n; ++x) out[x * stride] = kEdtInf;
        return;
    }
    for (int x = n - 1; x >= 0; --x) {
        while (q > 0 && t[q * stride] > x) --q;
        const long long i = s[q * stride];
        const long long v = (x - i) * (x - i) + h(static_cast<int>(i));
        out[x * stride] = v >= kEdtInf ? kEdtInf : static_cast<int>(v);
    }
}

template <typename Idx>
MSS_EDT_HD void edt_line(const int* h, int* out, int* s, int* t, int n, Idx stride) {
    edt_line_fn<Idx>([h, stride](int u) -> long long { return h[u * stride]; }, out, s, t, n, stride);
}

// first pass straight from a uint8 feature mask: h = 0 on features, "none" elsewhere
template <typename Idx>
MSS_EDT_HD void edt_line_mask(const unsigned char* m, int* out, int* s, int* t, int n, Idx stride) {
    edt_line_fn<Idx>([m, stride](int u) -> long long { return m[u * stride] ? 0 : kEdtInf; }, out, s, t, n, stride);
}

// Is the voxel at `c` (class `cls`, position (z, y, x) inside a box of n[0] x n[1] x n[2] voxels, strides sz / sy / 1)
// on the surface of its class?  binary_erosion(mask) XOR mask with scipy's 6-connected structure and border_value 0 on
// the CROPPED box: a neighbour outside the box is background, and an axis along which the box is one voxel thick does
// not exist (MONAI squeezes it away before the erosion).
MSS_EDT_HD bool mask_edge_at(const unsigned char* c, long long sz, long long sy, int z, int y, int x, const int (&n)[3],
                             unsigned cls) {
    if (*c != cls) return false;
    bool edge = false;
    if (n[0] > 1) edge |= (z == 0 || c[-sz] != cls) || (z == n[0] - 1 || c[sz] != cls);
    if (n[1] > 1) edge |= (y == 0 || c[-sy] != cls) || (y == n[1] - 1 || c[sy] != cls);
    if (n[2] > 1) edge |= (x == 0 || c[-1] != cls) || (x == n[2] - 1 || c[1] != cls);
    return edge;
}

}  // namespace mss
