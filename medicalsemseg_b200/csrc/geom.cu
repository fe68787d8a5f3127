// Host-side geometry: window starts, cover tables, layout validation, error text.
// Replaces the window-grid part of engine/utils.py:95-110 (MONAI fall_back/scan-interval/dense_patch_slices).
#include <cstring>

#include "common.cuh"

namespace mss {

char* last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(last_error_buf(), 512, "CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
    return static_cast<int>(e);
}

int make_geo(const mss_layout_t* lay, Geo* g, bool windows_inside) {
    MSS_REQUIRE(lay != nullptr, MSS_E_ARG, "layout is null");
    MSS_REQUIRE(lay->table_host != nullptr && lay->table_dev != nullptr, MSS_E_ARG, "layout tables are null");
    const int32_t* t = lay->table_host;
    MSS_REQUIRE(t[kHdrMagic] == kTableMagic, MSS_E_ARG, "geometry table has a bad magic word");
    for (int a = 0; a < 3; ++a) {
        MSS_REQUIRE(lay->image[a] > 0 && lay->roi[a] > 0 && lay->roi[a] <= lay->image[a], MSS_E_ARG,
                    "axis %d: need 0 < roi (%d) <= image (%d)", a, lay->roi[a], lay->image[a]);
        MSS_REQUIRE(t[kHdrImage + a] == lay->image[a] && t[kHdrRoi + a] == lay->roi[a] &&
                        t[kHdrN + a] == lay->n_starts[a],
                    MSS_E_ARG, "axis %d: layout disagrees with its geometry table", a);
        MSS_REQUIRE(lay->n_starts[a] > 0 && lay->n_starts[a] < 65536, MSS_E_ARG, "axis %d: bad window count", a);
        MSS_REQUIRE(0 <= lay->win_lo[a] && lay->win_lo[a] < lay->win_hi[a] && lay->win_hi[a] <= lay->n_starts[a],
                    MSS_E_ARG, "axis %d: owned window box [%d,%d) outside [0,%d)", a, lay->win_lo[a], lay->win_hi[a],
                    lay->n_starts[a]);
        MSS_REQUIRE(lay->extent[a] > 0 && lay->origin[a] >= 0 && lay->origin[a] + lay->extent[a] <= lay->image[a],
                    MSS_E_ARG, "axis %d: buffer box [%d,+%d) outside the image (%d)", a, lay->origin[a],
                    lay->extent[a], lay->image[a]);
        // every owned window must lie inside the buffer
        const int32_t* st = t + t[kHdrOffStarts + a];
        MSS_REQUIRE(!windows_inside || (st[lay->win_lo[a]] >= lay->origin[a] &&
                                        st[lay->win_hi[a] - 1] + lay->roi[a] <= lay->origin[a] + lay->extent[a]),
                    MSS_E_ARG, "axis %d: owned windows reach outside the buffer box", a);
        g->img[a] = lay->image[a];
        g->roi[a] = lay->roi[a];
        g->ns[a] = lay->n_starts[a];
        g->wlo[a] = lay->win_lo[a];
        g->whi[a] = lay->win_hi[a];
        g->nwl[a] = lay->win_hi[a] - lay->win_lo[a];
        g->org[a] = lay->origin[a];
        g->ext[a] = lay->extent[a];
        g->starts[a] = lay->table_dev + t[kHdrOffStarts + a];
        g->cover[a] = lay->table_dev + t[kHdrOffCover + a];
    }
    MSS_REQUIRE(lay->pitch_w >= lay->extent[2], MSS_E_ARG, "pitch_w (%d) < extent W (%d)", lay->pitch_w,
                lay->extent[2]);
    MSS_REQUIRE(lay->n_volumes > 0 && lay->n_classes > 0, MSS_E_ARG, "need n_volumes > 0 and n_classes > 0");
    g->pitch = lay->pitch_w;
    g->nb = lay->n_volumes;
    g->K = lay->n_classes;
    g->n_local = static_cast<long long>(g->nwl[0]) * g->nwl[1] * g->nwl[2];
    return MSS_OK;
}

}  // namespace mss

using namespace mss;

extern "C" {

int mss_abi_version(void) { return MSS_ABI_VERSION; }

const char* mss_last_error(void) { return last_error_buf(); }

int mss_axis_starts(int32_t image, int32_t roi, int32_t interval, int32_t* starts_out, int32_t cap) {
    MSS_REQUIRE(image > 0 && roi > 0 && roi <= image && interval >= 0, MSS_E_ARG,
                "axis_starts: need 0 < roi <= image and interval >= 0 (image=%d roi=%d interval=%d)", image, roi,
                interval);
    int n = 1;
    if (interval > 0) {
        const int upper = (image + interval - 1) / interval;  // ceil(image / interval)
        n = 1;                                                // "1 if none" branch of dense_patch_slices
        for (int d = 0; d < upper; ++d) {
            if (static_cast<long long>(d) * interval + roi >= image) {
                n = d + 1;
                break;
            }
        }
    }
    if (starts_out != nullptr) {
        MSS_REQUIRE(cap >= n, MSS_E_ARG, "axis_starts: capacity %d < %d windows", cap, n);
        for (int k = 0; k < n; ++k) {
            int s = k * interval;
            const int over = s + roi - image;
            if (over > 0) s -= over;  // last window is pulled back inside the image (clamped, not padded)
            starts_out[k] = s;
        }
    }
    return n;
}

int64_t mss_geom_table_len(const int32_t image[3], const int32_t n_starts[3]) {
    if (image == nullptr || n_starts == nullptr) return MSS_E_ARG;
    int64_t len = kHdrWords;
    for (int a = 0; a < 3; ++a) {
        if (image[a] <= 0 || n_starts[a] <= 0) return MSS_E_ARG;
        len += n_starts[a] + image[a];
    }
    return len;
}

int mss_geom_table_build(const int32_t image[3], const int32_t roi[3], const int32_t n_starts[3],
                         const int32_t* starts_d, const int32_t* starts_h, const int32_t* starts_w,
                         int32_t* table_out, int64_t table_len) {
    MSS_REQUIRE(image && roi && n_starts && starts_d && starts_h && starts_w && table_out, MSS_E_ARG,
                "geom_table_build: null argument");
    const int64_t need = mss_geom_table_len(image, n_starts);
    MSS_REQUIRE(need > 0 && table_len >= need, MSS_E_ARG, "geom_table_build: table too small (%lld < %lld)",
                static_cast<long long>(table_len), static_cast<long long>(need));
    const int32_t* starts[3] = {starts_d, starts_h, starts_w};
    int32_t* t = table_out;
    memset(t, 0, sizeof(int32_t) * kHdrWords);
    t[kHdrMagic] = kTableMagic;
    int32_t off = kHdrWords;
    for (int a = 0; a < 3; ++a) {
        MSS_REQUIRE(roi[a] > 0 && roi[a] <= image[a] && n_starts[a] < 65536, MSS_E_ARG,
                    "geom_table_build: axis %d: need 0 < roi <= image", a);
        t[kHdrImage + a] = image[a];
        t[kHdrRoi + a] = roi[a];
        t[kHdrN + a] = n_starts[a];
        t[kHdrOffStarts + a] = off;
        for (int i = 0; i < n_starts[a]; ++i) {
            const int s = starts[a][i];
            MSS_REQUIRE(s >= 0 && s + roi[a] <= image[a], MSS_E_ARG, "axis %d: window %d starts at %d, outside", a, i,
                        s);
            MSS_REQUIRE(i == 0 || s > starts[a][i - 1], MSS_E_ARG, "axis %d: starts must increase strictly", a);
            t[off + i] = s;
        }
        off += n_starts[a];
    }
    for (int a = 0; a < 3; ++a) {
        t[kHdrOffCover + a] = off;
        int lo = 0, hi = 0;  // windows [lo, hi) cover coordinate x
        for (int x = 0; x < image[a]; ++x) {
            while (hi < n_starts[a] && starts[a][hi] <= x) ++hi;
            while (lo < hi && starts[a][lo] + roi[a] <= x) ++lo;
            MSS_REQUIRE(lo < hi, MSS_E_ARG, "axis %d: coordinate %d is covered by no window", a, x);
            t[off + x] = lo | (hi << 16);
        }
        off += image[a];
    }
    return MSS_OK;
}

}  // extern "C"
