// Test-only: runs the bit-sliced per-chunk logic of the vote and Dice kernels (medicalsemseg_b200/csrc/bitslice.cuh,
// the very code the kernels inline) on the CPU, so the CPU suite can check it against the oracle without a GPU.
// Built by tests/test_bitslice_host.py with g++ into tests/_build/; never loaded by the product.
#include <cstdint>
#include <cstring>

#include "../../medicalsemseg_b200/csrc/bitslice.cuh"

using namespace mss;

template <int M>
static void vote_all(const uint8_t* const* maps, int K, long long nchunks, uint8_t* out) {
    for (long long i = 0; i < nchunks; ++i) {
        unsigned w[M][8], res[8];
        for (int m = 0; m < M; ++m) std::memcpy(w[m], maps[m] + i * 32, 32);
        vote_chunk<M>(w, K, res);
        std::memcpy(out + i * 32, res, 32);
    }
}

extern "C" int host_vote_chunks(const uint8_t* const* maps, int M, int K, long long n_voxels, uint8_t* out) {
    const long long nchunks = n_voxels / 32;
    switch (M) {
        case 1: vote_all<1>(maps, K, nchunks, out); break;
        case 2: vote_all<2>(maps, K, nchunks, out); break;
        case 3: vote_all<3>(maps, K, nchunks, out); break;
        case 4: vote_all<4>(maps, K, nchunks, out); break;
        case 5: vote_all<5>(maps, K, nchunks, out); break;
        case 6: vote_all<6>(maps, K, nchunks, out); break;
        case 7: vote_all<7>(maps, K, nchunks, out); break;
        case 8: vote_all<8>(maps, K, nchunks, out); break;
        default: return -1;
    }
    return 0;
}

// counts[3][16] (TP, P, Y), all labels must be < 16
extern "C" int host_dice_chunks(const uint8_t* pred, const uint8_t* label, int K, long long n_voxels, long long* counts) {
    const long long nchunks = n_voxels / 32;
    unsigned tp[8] = {0}, pp[8] = {0}, yy[8] = {0};
    auto flush = [&]() {
        for (int i = 0; i < 8; ++i) {
            counts[0 * 16 + 2 * i] += tp[i] & 0xffffu, counts[0 * 16 + 2 * i + 1] += tp[i] >> 16;
            counts[1 * 16 + 2 * i] += pp[i] & 0xffffu, counts[1 * 16 + 2 * i + 1] += pp[i] >> 16;
            counts[2 * 16 + 2 * i] += yy[i] & 0xffffu, counts[2 * 16 + 2 * i + 1] += yy[i] >> 16;
            tp[i] = pp[i] = yy[i] = 0;
        }
    };
    int iters = 0;
    for (long long i = 0; i < nchunks; ++i) {
        unsigned pw[8], yw[8];
        std::memcpy(pw, pred + i * 32, 32);
        std::memcpy(yw, label + i * 32, 32);
        if (has_wide_label(pw) || has_wide_label(yw)) return -2;
        switch ((K + 1) / 2) {
            case 1: dice_chunk<1>(pw, yw, tp, pp, yy); break;
            case 2: dice_chunk<2>(pw, yw, tp, pp, yy); break;
            case 3: dice_chunk<3>(pw, yw, tp, pp, yy); break;
            case 4: dice_chunk<4>(pw, yw, tp, pp, yy); break;
            case 5: dice_chunk<5>(pw, yw, tp, pp, yy); break;
            case 6: dice_chunk<6>(pw, yw, tp, pp, yy); break;
            case 7: dice_chunk<7>(pw, yw, tp, pp, yy); break;
            default: dice_chunk<8>(pw, yw, tp, pp, yy); break;
        }
        if (++iters == 2040) flush(), iters = 0;  // per-thread 16-bit halves: 2040 * 32 < 65536
    }
    flush();
    return 0;
}

extern "C" void host_bitplanes_roundtrip(const uint8_t* in32, uint8_t* out32, unsigned* planes4) {
    unsigned w[8], q[4], r[8];
    std::memcpy(w, in32, 32);
    bitplanes32(w, q);
    std::memcpy(planes4, q, 16);
    labels_from_bitplanes32(q, r);
    std::memcpy(out32, r, 32);
}

// ---- exact EDT line routine (csrc/edt.cuh) run over a whole volume on the host: three passes like the GPU driver ----
#include "../../medicalsemseg_b200/csrc/edt.cuh"
#include <vector>

// feature: uint8 mask [d, h, w] (non-zero = feature voxel); out: int32 squared distances (kEdtInf when there is none)
extern "C" void host_edt_squared(const uint8_t* feature, int d, int h, int w, int* out) {
    const long long n = static_cast<long long>(d) * h * w;
    std::vector<int> a(n), b(n), s(n), t(n);
    // axis 0 (stride h*w) straight from the mask, axis 1 (stride w), axis 2 (stride 1); never in place, so ping-pong
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const long long o = static_cast<long long>(y) * w + x;
            mss::edt_line_mask<long long>(feature + o, b.data() + o, s.data() + o, t.data() + o, d, static_cast<long long>(h) * w);
        }
    for (int z = 0; z < d; ++z)
        for (int x = 0; x < w; ++x) {
            const long long o = static_cast<long long>(z) * h * w + x;
            mss::edt_line<long long>(b.data() + o, a.data() + o, s.data() + o, t.data() + o, h, w);
        }
    for (int z = 0; z < d; ++z)
        for (int y = 0; y < h; ++y) {
            const long long o = (static_cast<long long>(z) * h + y) * w;
            mss::edt_line<long long>(a.data() + o, b.data() + o, s.data() + o, t.data() + o, w, 1);
        }
    for (long long i = 0; i < n; ++i) out[i] = b[i];
}

// the order and the routines the GPU driver uses since round 2: row scan along W from the mask, then the envelope with
// the register-cached stack top along H and D
// use_recip: divide by the reciprocal table (what the pass kernel does for lines up to 2048 voxels)
extern "C" void host_edt_squared_v2(const uint8_t* feature, int d, int h, int w, int use_recip, int* out) {
    const long long n = static_cast<long long>(d) * h * w;
    std::vector<int> a(n), b(n), s(n), t(n);
    std::vector<unsigned> recip(static_cast<size_t>(d > h ? d : h) + 1, 0u);
    for (size_t k = 1; k < recip.size(); ++k) recip[k] = mss::edt_recip(static_cast<unsigned>(k));
    const unsigned* rp = use_recip ? recip.data() : nullptr;
    for (long long r = 0; r < static_cast<long long>(d) * h; ++r) mss::edt_row_from_mask(feature + r * w, a.data() + r * w, w);
    for (int z = 0; z < d; ++z)
        for (int x = 0; x < w; ++x) {
            const long long o = static_cast<long long>(z) * h * w + x;
            mss::edt_line_cached<long long>(a.data() + o, b.data() + o, s.data() + o, t.data() + o, h, w, rp);
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const long long o = static_cast<long long>(y) * w + x;
            mss::edt_line_cached<long long>(b.data() + o, a.data() + o, s.data() + o, t.data() + o, d, static_cast<long long>(h) * w, rp);
        }
    for (long long i = 0; i < n; ++i) out[i] = a[i];
}

// edt_div2k against the plain floor division over a strided sweep of numerators for every k < kmax: returns the mismatches
extern "C" long long host_edt_div_check(int kmax, unsigned step) {
    long long bad = 0;
    for (unsigned k = 1; k < static_cast<unsigned>(kmax); ++k) {
        const unsigned r = mss::edt_recip(k);
        for (unsigned long long a = 0; a < (1ull << 31); a += step + k) bad += mss::edt_div2k(static_cast<unsigned>(a), k, r) != a / (2 * k);
        for (unsigned a = (1u << 31) - 4096; a < (1u << 31); ++a) bad += mss::edt_div2k(a, k, r) != a / (2 * k);
        for (unsigned a = 0; a < 8 * k + 8; ++a) bad += mss::edt_div2k(a, k, r) != a / (2 * k);
    }
    return bad;
}

// surface voxels of class cls inside the box [lo, hi) of a label map [dims], as csrc/hausdorff.cu computes them
extern "C" void host_mask_edges(const uint8_t* labels, const int* dims, int cls, const int* lo, const int* hi, uint8_t* edges) {
    const int n[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    const long long sy = dims[2], sz = static_cast<long long>(dims[1]) * dims[2];
    for (int z = 0; z < n[0]; ++z)
        for (int y = 0; y < n[1]; ++y)
            for (int x = 0; x < n[2]; ++x) {
                const uint8_t* c = labels + (lo[0] + z) * sz + (lo[1] + y) * sy + (lo[2] + x);
                edges[(static_cast<long long>(z) * n[1] + y) * n[2] + x] = mss::mask_edge_at(c, sz, sy, z, y, x, n, cls) ? 1 : 0;
            }
}
