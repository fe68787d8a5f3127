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

// floor(a / (2 k)) for 0 <= a < 2^31, 1 <= k < 2^15, from the reciprocal r = ceil(2^31 / k) = ceil(2^32 / (2 k)): the high word
// of a * r is the quotient or one above it (a * r / 2^32 = a / (2 k) + a * e / (2 k * 2^32) with e < 2 k, an excess below 1 / 2),
// one multiply-back decides.  Replaces the ~30-instruction emulated integer division (I2F, MUFU.RCP, F2I, two multiply-highs
// and corrections) the envelope paid per pushed parabola.
MSS_EDT_HD unsigned edt_recip(unsigned k) { return (0x80000000u + k - 1u) / k; }
MSS_EDT_HD unsigned edt_div2k(unsigned a, unsigned k, unsigned r) {
#ifdef __CUDA_ARCH__
    unsigned q = __umulhi(a, r);
#else
    unsigned q = static_cast<unsigned>((static_cast<unsigned long long>(a) * r) >> 32);
#endif
    return q * (2u * k) > a ? q - 1u : q;
}

// The same lower envelope with the TOP of the parabola stack kept in registers: a push spills the old top (one packed
// position / take-over word into st[], its value into hv[] - two independent stores nobody waits for), only a pop loads
// (the entry below).  In the common case - the new parabola just joins the envelope - an iteration touches no scratch
// memory at all, where edt_line_fn pays two dependent round trips (s[q], then h[s[q]]).  Positions and take-over points
// are < 2^14 (volume sides < 16384), so both fit one 32-bit word.  Same comparisons, same integer arithmetic, same result.
// `recip` (optional): edt_recip(k) for 0 < k < n, e.g. a shared-memory table built once per CTA - then no division is executed.
template <typename Idx, typename HFn>
MSS_EDT_HD void edt_line_cached_fn(HFn h, int* out, int* st, int* hv, int n, Idx stride, const unsigned* recip = nullptr) {
    int q = -1;                       // index of the top entry (entries 0 .. q-1 live in st / hv)
    // 32-bit arithmetic throughout: positions < 2^14, finite values < 2^29, so (x - i)^2 + h < 2^28 + 2^29 < 2^31
    int ts = 0, tt = 0, th = 0;  // the top: position, take-over point, value
    constexpr int kAhead = 8;          // line values fetched together: one memory latency per 8 steps, not per step
    for (int u0 = 0; u0 < n; u0 += kAhead) {
        int hb[kAhead];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < kAhead; ++k) hb[k] = u0 + k < n ? static_cast<int>(h(u0 + k)) : kEdtInf;  // (values > kEdtInf: "none" too)
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < kAhead; ++k) {
            const int u = u0 + k;
            const int hu = hb[k];
            if (hu >= kEdtInf) continue;
            while (q >= 0) {
                // parabola u is at or below the top parabola at the point where the top took over: the top never wins
                if ((tt - ts) * (tt - ts) + th > (tt - u) * (tt - u) + hu) {
                    --q;
                    if (q >= 0) {
                        const int e = st[q * stride];
                        ts = e & 0xffff;
                        tt = e >> 16;
                        th = hv[q * stride];
                    }
                } else {
                    break;
                }
            }
            if (q < 0) {
                q = 0;
                ts = u;
                tt = 0;
                th = hu;
            } else {
                const int num = u * u - ts * ts + hu - th, den = 2 * (u - ts);
                int w;  // floor(num / den)
                if (recip != nullptr) {
                    const unsigned k = static_cast<unsigned>(u - ts), r = recip[k];
                    w = num >= 0 ? static_cast<int>(edt_div2k(static_cast<unsigned>(num), k, r))
                                 : -static_cast<int>(edt_div2k(static_cast<unsigned>(-num + den - 1), k, r));
                } else {
                    w = num >= 0 ? num / den : -((-num + den - 1) / den);
                }
                w += 1;
                if (w < n) {
                    st[q * stride] = ts | (tt << 16);
                    hv[q * stride] = th;
                    ++q;
                    ts = u;
                    tt = w < 0 ? 0 : w;
                    th = hu;
                }
            }
        }
    }
    if (q < 0) {
        for (int x = 0; x < n; ++x) out[x * stride] = kEdtInf;
        return;
    }
    for (int x = n - 1; x >= 0; --x) {
        while (q > 0 && tt > x) {
            --q;
            const int e = st[q * stride];
            ts = e & 0xffff;
            tt = e >> 16;
            th = hv[q * stride];
        }
        const int v = (x - ts) * (x - ts) + th;
        out[x * stride] = v >= kEdtInf ? kEdtInf : v;
    }
}

template <typename Idx>
MSS_EDT_HD void edt_line_cached(const int* h, int* out, int* st, int* hv, int n, Idx stride, const unsigned* recip = nullptr) {
    edt_line_cached_fn<Idx>([h, stride](int u) -> int { return h[u * stride]; }, out, st, hv, n, stride, recip);
}

template <typename Idx>
MSS_EDT_HD void edt_line_mask_cached(const unsigned char* m, int* out, int* st, int* hv, int n, Idx stride) {
    edt_line_cached_fn<Idx>([m, stride](int u) -> int { return m[u * stride] ? 0 : kEdtInf; }, out, st, hv, n, stride);
}

// First pass along a line straight from a feature mask without any envelope: the squared distance to the nearest
// feature of the SAME line is (distance to the nearest set position)^2.  Host form of the row-scan kernel.
MSS_EDT_HD void edt_row_from_mask(const unsigned char* m, int* out, int n) {
    int last = -(1 << 20);
    for (int x = 0; x < n; ++x) {
        if (m[x]) last = x;
        out[x] = x - last;  // distance to the nearest feature on the left (huge when none)
    }
    int next = 1 << 20;
    for (int x = n - 1; x >= 0; --x) {
        if (m[x]) next = x;
        const long long d = out[x] < next - x ? out[x] : next - x;
        out[x] = d >= (1 << 14) ? kEdtInf : static_cast<int>(d * d);
    }
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
