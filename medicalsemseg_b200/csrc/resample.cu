// Nearest-neighbour resampling of a uint8 label map back to the original grid
// (replaces utils/misc.py:420-425 `resample_3d` = scipy.ndimage.zoom(order=0, prefilter=False), called at
// engine/test.py:143-147 between the argmax and the NIfTI write that majority_vote.py consumes).
//
// scipy's zoom is separable: output index k of an axis reads input index floor(k * zoom + 0.5) with
// zoom = (n_in - 1) / (n_out - 1) in float64 (1.0 when n_out == 1), and - mode 'constant' - answers cval = 0
// when k * zoom falls outside [0, n_in - 1], which float rounding makes happen for the LAST index of some
// (n_in, n_out) pairs.  mss_zoom_index_table reproduces that rule per axis on the host (-1 = constant);
// the kernels are then pure byte gathers: out[x, y, z] = in[ix[x], iy[y], iz[z]].
// Rows kernel: a CTA walks a range of output rows (b, x, y) in passes of `rp` rows.  A pass stages its source rows
// in shared memory with aligned 16-byte copies, then every thread gathers the 16 voxels of its z-chunk from shared
// memory - per output word three aligned 4-byte loads, two funnel shifts and one PRMT whose selector was computed
// once (the chunk of a thread never changes, so its gather plan lives in registers for the whole kernel) - and
// stores them with one 16-byte store (funnel-shifted 4-byte stores where the output row is not 16-byte aligned).
// Gather kernel: the general fallback (rows wider than 4096 voxels or too long for shared memory), byte loads via L1.
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace mss {

struct ResampleParams {
    const uint8_t* in;
    uint8_t* out;
    const int* ix;
    const int* iy;
    const int* iz;
    int in_dims[3];
    int out_dims[3];
    long long n_volumes;
    // rows kernel
    int n_chunks;      // 16-voxel chunks per output row
    int rp;            // rows per pass
    int slot_pitch;    // bytes of one staged source row (multiple of 16)
    int rows_per_block;
};

// 16 gathered voxels (little-endian in w) -> dst, with the widest stores the address allows
__device__ __forceinline__ void store_chunk(uint8_t* dst, const unsigned (&w)[4], int nz) {
    const unsigned a = static_cast<unsigned>(reinterpret_cast<uintptr_t>(dst));
    if (nz == 16) {
        if ((a & 15u) == 0) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
            return;
        }
        const unsigned m = a & 3u;
        if (m == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<unsigned*>(dst + 4 * q) = w[q];
            return;
        }
        // head bytes up to the next 4-byte boundary, three aligned words built by funnel shifts, tail bytes
        const unsigned head = 4u - m;  // 1..3
        dst[0] = static_cast<uint8_t>(w[0]);
        if (head > 1) dst[1] = static_cast<uint8_t>(w[0] >> 8);
        if (head > 2) dst[2] = static_cast<uint8_t>(w[0] >> 16);
        const unsigned sh = 8u * head;
        *reinterpret_cast<unsigned*>(dst + head) = __funnelshift_r(w[0], w[1], sh);
        *reinterpret_cast<unsigned*>(dst + head + 4) = __funnelshift_r(w[1], w[2], sh);
        *reinterpret_cast<unsigned*>(dst + head + 8) = __funnelshift_r(w[2], w[3], sh);
        const unsigned tail = w[3] >> sh;  // the m = 1..3 last bytes
        dst[head + 12] = static_cast<uint8_t>(tail);
        if (m > 1) dst[head + 13] = static_cast<uint8_t>(tail >> 8);
        if (m > 2) dst[head + 14] = static_cast<uint8_t>(tail >> 16);
        return;
    }
    for (int e = 0; e < nz; ++e) {
        const unsigned word = e < 4 ? w[0] : (e < 8 ? w[1] : (e < 12 ? w[2] : w[3]));
        dst[e] = static_cast<uint8_t>(word >> (8 * (e & 3)));
    }
}

constexpr int kRowsPerThread = 4;  // rows a thread gathers per pass (same z-chunk, rows `rp` apart)

__global__ void __launch_bounds__(256) resample_rows_kernel(const __grid_constant__ ResampleParams p) {
    extern __shared__ __align__(16) uint8_t stage[];  // [rp * kRowsPerThread][slot_pitch]
    __shared__ int s_off[256 * kRowsPerThread];       // per slot: byte offset of the row inside its slot, -1 = constant row
    __shared__ const uint8_t* s_src[256 * kRowsPerThread];
    const int tid = threadIdx.x;
    const int oz = p.out_dims[2], izd = p.in_dims[2], oyd = p.out_dims[1], oxd = p.out_dims[0];
    const int chunk = tid % p.n_chunks, slot = tid / p.n_chunks;
    const bool active = slot < p.rp;
    const int n_rows = static_cast<int>(p.n_volumes) * oxd * oyd;  // < 2^31 (checked by the host)
    const int row_begin = blockIdx.x * p.rows_per_block;
    const int row_end = min(row_begin + p.rows_per_block, n_rows);
    const int nz = min(16, oz - chunk * 16);
    // Gather plan of this thread's 16 outputs, fixed for the whole kernel: per output word q the first source index,
    // a PRMT selector of the 4 sources relative to it (valid when they span <= 8 bytes, i.e. zoom factors up to ~2.3)
    // and a byte mask of the outputs that are real (not past the row, not scipy's constant).
    int first[4];
    unsigned sel[4], keep[4];
    bool narrow = true;
    {
        int zi[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) zi[e] = (active && e < nz) ? __ldg(p.iz + chunk * 16 + e) : -1;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int f = -1;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (f < 0 && zi[4 * q + e] >= 0) f = zi[4 * q + e];
            first[q] = f < 0 ? 0 : f;
            sel[q] = 0u;
            keep[q] = 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = zi[4 * q + e];
                if (i >= 0) {
                    const int rel = i - first[q];
                    if (rel < 0 || rel > 7) narrow = false;
                    sel[q] |= static_cast<unsigned>(rel & 7) << (4 * e);
                    keep[q] |= 0xFFu << (8 * e);
                }
            }
        }
    }
    const uint8_t* in_end = p.in + p.n_volumes * p.in_dims[0] * p.in_dims[1] * static_cast<long long>(izd);
    const int pass_rows = p.rp * kRowsPerThread;

    for (int row0 = row_begin; row0 < row_end; row0 += pass_rows) {
        // ---- where do the source rows of this pass start? (one thread per row) ---------------------------------
        for (int s2 = tid; s2 < pass_rows; s2 += 256) {
            const int row = row0 + s2;
            const uint8_t* src = nullptr;
            if (row < row_end) {
                const int r2 = row / oyd, oy = row - r2 * oyd;
                const int b = r2 / oxd, ox = r2 - b * oxd;
                const int sx = __ldg(p.ix + ox), sy = __ldg(p.iy + oy);
                if (sx >= 0 && sy >= 0)
                    src = p.in + ((static_cast<long long>(b) * p.in_dims[0] + sx) * p.in_dims[1] + sy) * izd;
            }
            s_src[s2] = src;
            s_off[s2] = src != nullptr ? static_cast<int>(reinterpret_cast<uintptr_t>(src) & 15u) : -1;
        }
        __syncthreads();
        // ---- stage them: one warp per row, lane j copies the j-th aligned 16-byte vector (asynchronous, all in
        // flight together); only the first / last rows of the tensor can touch bytes outside it -------------------
        for (int s2 = tid >> 5; s2 < pass_rows; s2 += 8) {
            const uint8_t* src = s_src[s2];
            if (src == nullptr) continue;
            const int m = s_off[s2];
            const uint8_t* g0 = src - m;
            uint8_t* d0 = stage + static_cast<size_t>(s2) * p.slot_pitch;
            const int nv = (m + izd + 15) >> 4;
            const bool interior = g0 >= p.in && g0 + 16 * nv <= in_end;  // uniform per warp
            for (int j = tid & 31; j < nv; j += 32) {
                if (interior) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                                     static_cast<unsigned>(__cvta_generic_to_shared(d0 + 16 * j))),
                                 "l"(g0 + 16 * j)
                                 : "memory");
                } else {
                    for (int e = 0; e < 16; ++e) {
                        const uint8_t* g = g0 + 16 * j + e;
                        d0[16 * j + e] = (g >= p.in && g < in_end) ? *g : 0;
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // ---- gather -----------------------------------------------------------------------------------------
        if (active) {
#pragma unroll
            for (int rr = 0; rr < kRowsPerThread; ++rr) {
                const int s = slot + rr * p.rp;
                const int row = row0 + s;
                if (row >= row_end) break;
                const int off = s_off[s];
                const uint8_t* base = stage + static_cast<size_t>(s) * p.slot_pitch + (off < 0 ? 0 : off);
                unsigned w[4] = {0u, 0u, 0u, 0u};
                if (off >= 0) {
                    if (narrow) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            // 8 source bytes starting at first[q]: three aligned words, funnel-shifted to the byte
                            const unsigned pos = static_cast<unsigned>(first[q]) + (base - stage);
                            const unsigned* wp = reinterpret_cast<const unsigned*>(stage + (pos & ~3u));
                            const unsigned sh = (pos & 3u) * 8u;
                            const unsigned a = wp[0], b = wp[1], c = wp[2];
                            const unsigned lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);
                            w[q] = __byte_perm(lo, hi, sel[q]) & keep[q];
                        }
                    } else {  // extreme zoom factors: byte by byte
#pragma unroll 1
                        for (int e = 0; e < nz; ++e) {
                            const int i = __ldg(p.iz + chunk * 16 + e);
                            const unsigned v = i >= 0 ? static_cast<unsigned>(base[i]) : 0u;
                            if (e < 4) w[0] |= v << (8 * e);
                            else if (e < 8) w[1] |= v << (8 * (e - 4));
                            else if (e < 12) w[2] |= v << (8 * (e - 8));
                            else w[3] |= v << (8 * (e - 12));
                        }
                    }
                }
                store_chunk(p.out + static_cast<long long>(row) * oz + chunk * 16, w, nz);
            }
        }
        __syncthreads();
    }
}

// Stream kernel (the default).  The output is walked as ONE flat byte stream in chunks of CB bytes, so every store is an
// aligned CB-byte vector whatever the row length (rows of 147 voxels start at every alignment).  A chunk's position in
// its row repeats with period P = oz / gcd(CB, oz) chunks = RS = CB / gcd(CB, oz) rows: thread c of a period slot always
// gets the chunk that starts at byte CB * c of a period, so its gather plan - the z source offsets of its CB bytes and
// how many of them still belong to the chunk's first row - is fixed and lives in registers; per period it only needs
// the source rows of (at most) two output rows, which it tracks incrementally.  No staging, no barriers: the byte loads
// go through L1, where the source rows of neighbouring threads meet.
struct StreamParams {
    const uint8_t* in;
    uint8_t* out;
    const int* ix;
    const int* iy;
    const int* iz;
    int in_dims[3];
    int out_dims[3];
    int n_rows;         // n_volumes * ox * oy
    long long total;    // output bytes
    int P, RS, NS;      // chunks and rows per period, period slots per CTA
    int periods;        // ceil(n_rows / RS)
    int per_block;      // periods per CTA (a multiple of NS)
    int pf;             // periods the L2 prefetch runs ahead (0 = none)
};

// one output row's source: advanced by a fixed number of rows per step
struct RowCursor {
    int oy, ox, b;
    int xrow;     // (b * in_x + sx) * in_y, with sx = 0 when the x index is scipy's constant
    bool xvalid;
    __device__ __forceinline__ void load_x(const StreamParams& p) {
        const int sx = __ldg(p.ix + ox);
        xvalid = sx >= 0;
        xrow = (b * p.in_dims[0] + (sx < 0 ? 0 : sx)) * p.in_dims[1];
    }
    __device__ __forceinline__ void init(const StreamParams& p, int row) {
        const int r2 = row / p.out_dims[1];
        oy = row - r2 * p.out_dims[1];
        b = r2 / p.out_dims[0];
        ox = r2 - b * p.out_dims[0];
        load_x(p);
    }
    __device__ __forceinline__ void advance(const StreamParams& p, int step) {
        oy += step;
        if (oy >= p.out_dims[1]) {
            const int q = oy / p.out_dims[1];
            oy -= q * p.out_dims[1];
            ox += q;
            if (ox >= p.out_dims[0]) {
                const int q2 = ox / p.out_dims[0];
                ox -= q2 * p.out_dims[0];
                b += q2;
            }
            load_x(p);
        }
    }
    // first voxel of the source row (the tensor's first row when the output row is scipy's constant), and whether it is real
    __device__ __forceinline__ unsigned src(const StreamParams& p, bool in_range, bool* valid) const {
        const int sy = __ldg(p.iy + oy);
        *valid = in_range && xvalid && sy >= 0;
        return *valid ? static_cast<unsigned>(xrow + sy) * static_cast<unsigned>(p.in_dims[2]) : 0u;  // input < 4 GB (host)
    }
};

template <int CB>
struct ChunkStore;
template <>
struct ChunkStore<16> {
    static __device__ __forceinline__ void st(uint8_t* d, const unsigned* w) {
        *reinterpret_cast<uint4*>(d) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};
template <>
struct ChunkStore<8> {
    static __device__ __forceinline__ void st(uint8_t* d, const unsigned* w) { *reinterpret_cast<uint2*>(d) = make_uint2(w[0], w[1]); }
};
template <>
struct ChunkStore<4> {
    static __device__ __forceinline__ void st(uint8_t* d, const unsigned* w) { *reinterpret_cast<unsigned*>(d) = w[0]; }
};

// The loop of one thread.  CROSS = the thread's chunk straddles a row boundary (two source rows, a select per byte);
// the threads of a slot are ordered so that the few chunks that do (RS - 1 of P) sit in their own warps.
template <int CB, bool CROSS>
__device__ __forceinline__ void stream_loop(const StreamParams& p, int c, int sub) {
    constexpr int NW = CB / 4;
    const int oz = p.out_dims[2];
    const int ro = (c * CB) / oz, z0 = c * CB - ro * oz;
    const int n_first = CROSS ? oz - z0 : CB;  // bytes of the chunk that lie in its first row (oz >= CB: at most two rows)
    // the fixed gather plan
    unsigned off[CB], second[CROSS ? CB : 1];  // z source offset; all ones when the byte belongs to the chunk's second row
    unsigned keep0[NW], keep1[NW];             // byte masks of the real voxels of the first / second row
    bool clean = true;
#pragma unroll
    for (int q = 0; q < NW; ++q) keep0[q] = keep1[q] = 0u;
#pragma unroll
    for (int e = 0; e < CB; ++e) {
        int z = z0 + e;
        z = z >= oz ? z - oz : z;
        const int i = __ldg(p.iz + z);
        off[e] = i < 0 ? 0u : static_cast<unsigned>(i);
        if (CROSS) second[e] = e < n_first ? 0u : ~0u;
        clean = clean && i >= 0;
        const unsigned m = i >= 0 ? 0xFFu << (8 * (e & 3)) : 0u;
        keep0[e >> 2] |= e < n_first ? m : 0u;
        keep1[e >> 2] |= e < n_first ? 0u : m;
    }
    const int first = blockIdx.x * p.per_block;
    const int end = min(p.periods, first + p.per_block);
    int per = first + sub;
    if (per >= end) return;
    const int step = p.NS * p.RS;
    int row = per * p.RS + ro;  // (n_rows + RS < 2^31: checked by the host)
    RowCursor r0, r1, rp;
    r0.init(p, row);
    if (CROSS) r1.init(p, row + 1);
    // a third cursor runs `pf` periods ahead and asks L2 for the sectors this thread will gather then: the byte loads of a
    // period are one DRAM round trip otherwise (ncu: half of all stall samples on the first use of a loaded byte)
    const int pf_rows = p.pf * step;
    if (p.pf > 0) rp.init(p, row + pf_rows);
    long long o = (static_cast<long long>(per) * p.P + c) * CB;
    const long long o_step = static_cast<long long>(p.NS) * p.P * CB;
    // software pipeline: the byte loads of the NEXT period are issued before this period's bytes are packed and stored
    unsigned b[CB], bn[CB];
    bool v0, v1 = true, vn0 = false, vn1 = true;
    auto fetch = [&](unsigned (&dst)[CB], bool* w0, bool* w1) {
        const unsigned s0 = r0.src(p, row < p.n_rows, w0);
        unsigned d1 = 0u;
        if (CROSS) d1 = r1.src(p, row + 1 < p.n_rows, w1) - s0;
#pragma unroll
        for (int e = 0; e < CB; ++e) dst[e] = __ldg(p.in + (s0 + off[e] + (CROSS ? (d1 & second[e]) : 0u)));
    };
    auto prefetch = [&]() {
        if (p.pf > 0) {
            bool ok;
            const unsigned s = rp.src(p, row + pf_rows < p.n_rows, &ok);
            if (ok) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.in + (s + off[0])));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.in + (s + off[CROSS ? 0 : CB - 1])));
            }
            rp.advance(p, step);
        }
    };
    fetch(b, &v0, &v1);
    for (;;) {
        const bool more = per + p.NS < end;
        if (more) {
            prefetch();
            row += step;
            r0.advance(p, step);
            if (CROSS) r1.advance(p, step);
            fetch(bn, &vn0, &vn1);
        }
        unsigned w[NW];
#pragma unroll
        for (int q = 0; q < NW; ++q)
            w[q] = __byte_perm(__byte_perm(b[4 * q], b[4 * q + 1], 0x0040), __byte_perm(b[4 * q + 2], b[4 * q + 3], 0x0040), 0x5410);
        if (!(clean && v0 && v1)) {
#pragma unroll
            for (int q = 0; q < NW; ++q) w[q] &= (v0 ? keep0[q] : 0u) | (v1 ? keep1[q] : 0u);
        }
        if (o + CB <= p.total) {
            ChunkStore<CB>::st(p.out + o, w);
        } else {
#pragma unroll
            for (int e = 0; e < CB; ++e)
                if (o + e < p.total) p.out[o + e] = static_cast<uint8_t>(w[e >> 2] >> (8 * (e & 3)));
        }
        if (!more) break;
        per += p.NS;
        o += o_step;
        v0 = vn0, v1 = vn1;
#pragma unroll
        for (int e = 0; e < CB; ++e) b[e] = bn[e];
    }
}

template <int CB>
__global__ void __launch_bounds__(512) resample_stream_kernel(const __grid_constant__ StreamParams p) {
    const int tid = threadIdx.x;
    // RS - 1 chunks of a period straddle a row boundary: chunk floor(k * oz / CB) for k = 1 .. RS - 1.  The slots' other
    // chunks take the first NS * (P - RS + 1) threads, the straddling ones the threads after them.
    const int nc = p.RS - 1, plain = p.P - nc;
    int sub, c;
    bool cross;
    if (tid < p.NS * plain) {
        sub = tid / plain;
        c = tid - sub * plain;
        for (int k = 1; k <= nc; ++k)  // the c-th chunk that does not straddle
            if ((k * p.out_dims[2]) / CB <= c) ++c;
        cross = false;
    } else {
        const int i = tid - p.NS * plain;
        if (i >= p.NS * nc) return;
        sub = i / nc;
        c = ((i - sub * nc + 1) * p.out_dims[2]) / CB;
        cross = true;
    }
    if (__any_sync(__activemask(), cross)) stream_loop<CB, true>(p, c, sub);
    else stream_loop<CB, false>(p, c, sub);
}

__global__ void __launch_bounds__(256) resample_gather_kernel(const __grid_constant__ ResampleParams p) {
    const int oz_chunks = (p.out_dims[2] + 15) / 16;
    const long long rows = p.n_volumes * p.out_dims[0] * p.out_dims[1];
    const long long total = rows * oz_chunks;
    const long long in_plane = static_cast<long long>(p.in_dims[1]) * p.in_dims[2];
    const long long in_vol = in_plane * p.in_dims[0];
    for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = t / oz_chunks;
        const int z0 = static_cast<int>(t - row * oz_chunks) * 16;
        const int oy = static_cast<int>(row % p.out_dims[1]);
        const long long r2 = row / p.out_dims[1];
        const int ox = static_cast<int>(r2 % p.out_dims[0]);
        const long long b = r2 / p.out_dims[0];
        const int sx = __ldg(p.ix + ox), sy = __ldg(p.iy + oy);
        const bool row_in = sx >= 0 && sy >= 0;
        const uint8_t* src = p.in + b * in_vol + static_cast<long long>(row_in ? sx : 0) * in_plane +
                             static_cast<long long>(row_in ? sy : 0) * p.in_dims[2];
        unsigned w[4] = {0u, 0u, 0u, 0u};
        const int nz = min(16, p.out_dims[2] - z0);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            if (e < nz) {
                const int sz = __ldg(p.iz + z0 + e);
                const unsigned v = (row_in && sz >= 0) ? static_cast<unsigned>(__ldg(src + sz)) : 0u;
                w[e >> 2] |= v << (8 * (e & 3));
            }
        }
        store_chunk(p.out + row * p.out_dims[2] + z0, w, nz);
    }
}

}  // namespace mss

using namespace mss;

extern "C" int mss_zoom_index_table(int32_t n_in, int32_t n_out, int32_t* table_out) {
    MSS_REQUIRE(n_in > 0 && n_out > 0 && table_out != nullptr, MSS_E_ARG, "zoom_index_table: need positive sizes and a table");
    // scipy.ndimage.zoom, grid_mode=False: zoom = (n_in - 1) / (n_out - 1), 1.0 where the divisor is 0
    const double zoom = n_out > 1 ? static_cast<double>(n_in - 1) / static_cast<double>(n_out - 1) : 1.0;
    for (int32_t k = 0; k < n_out; ++k) {
        const double cc = static_cast<double>(k) * zoom;
        if (cc < 0.0 || cc > static_cast<double>(n_in - 1)) {
            table_out[k] = -1;  // mode='constant': outside the input -> cval (0)
        } else {
            int32_t i = static_cast<int32_t>(std::floor(cc + 0.5));  // order 0: nearest
            table_out[k] = i > n_in - 1 ? n_in - 1 : i;
        }
    }
    return MSS_OK;
}

extern "C" int mss_resample_nearest(const uint8_t* labels_in, const int32_t in_dims[3], uint8_t* labels_out,
                                    const int32_t out_dims[3], int64_t n_volumes, const int32_t* index_x,
                                    const int32_t* index_y, const int32_t* index_z, void* stream) {
    MSS_REQUIRE(labels_in && in_dims && labels_out && out_dims && index_x && index_y && index_z, MSS_E_ARG,
                "resample_nearest: null argument");
    ResampleParams p;
    for (int a = 0; a < 3; ++a) {
        MSS_REQUIRE(in_dims[a] > 0 && out_dims[a] > 0, MSS_E_ARG, "resample_nearest: dims must be positive");
        p.in_dims[a] = in_dims[a];
        p.out_dims[a] = out_dims[a];
    }
    MSS_REQUIRE(n_volumes > 0, MSS_E_ARG, "resample_nearest: n_volumes must be positive");
    p.in = labels_in;
    p.out = labels_out;
    p.ix = index_x;
    p.iy = index_y;
    p.iz = index_z;
    p.n_volumes = n_volumes;
    cudaStream_t s = as_stream(stream);
    const long long n_rows = n_volumes * out_dims[0] * out_dims[1];
    MSS_REQUIRE(n_rows < (1LL << 31), MSS_E_UNSUPPORTED, "resample_nearest: too many output rows for one call");
    // ---- stream kernel: aligned CB-byte chunks of the flat output, fixed gather plan per thread --------------------
    static const int force_cb = getenv("MSS_RESAMPLE_CB") ? atoi(getenv("MSS_RESAMPLE_CB")) : -1;  // tuning knob; 0 = rows kernel
    if (force_cb != 0 && (reinterpret_cast<uintptr_t>(labels_out) & 15u) == 0 &&
        static_cast<long long>(n_volumes) * in_dims[0] * in_dims[1] * in_dims[2] < (1LL << 32)) {
        const int cands[3] = {force_cb > 0 ? force_cb : 8, 16, 4};  // 8 measured best on B200 (kernel_bench)
        for (int ci = 0; ci < 3; ++ci) {
            const int cb = cands[ci];
            if ((cb != 4 && cb != 8 && cb != 16) || out_dims[2] < cb) continue;
            int g = cb, r = out_dims[2] % cb;
            while (r) {
                const int t = g % r;
                g = r;
                r = t;
            }
            StreamParams sp;
            sp.P = out_dims[2] / g;
            sp.RS = cb / g;
            if (sp.P > 512) continue;
            sp.NS = sp.P <= 256 ? 256 / sp.P : 1;
            const int threads = (sp.NS * sp.P + 31) / 32 * 32;
            sp.in = labels_in;
            sp.out = labels_out;
            sp.ix = index_x;
            sp.iy = index_y;
            sp.iz = index_z;
            for (int a = 0; a < 3; ++a) sp.in_dims[a] = in_dims[a], sp.out_dims[a] = out_dims[a];
            if (n_rows >= (1LL << 30)) continue;
            sp.n_rows = static_cast<int>(n_rows);
            static const int pf = getenv("MSS_RESAMPLE_PF") ? atoi(getenv("MSS_RESAMPLE_PF")) : 4;  // tuning knob
            sp.pf = pf;
            sp.total = n_rows * out_dims[2];
            sp.periods = static_cast<int>((n_rows + sp.RS - 1) / sp.RS);
            // one wave: as many CTAs as are resident together (a second, partial wave would run at a fraction of the occupancy)
            int resident = 0;
            MSS_CUDA(cb == 16  ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, resample_stream_kernel<16>, threads, 0)
                     : cb == 8 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, resample_stream_kernel<8>, threads, 0)
                               : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, resample_stream_kernel<4>, threads, 0));
            int dev = 0, sms = 148;
            if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const long long want = static_cast<long long>(sms) * (resident < 1 ? 1 : resident);
            const long long units = (sp.periods + sp.NS - 1) / sp.NS;  // loop iterations of all CTAs together
            const long long per_cta = (units + want - 1) / want;
            sp.per_block = static_cast<int>(per_cta) * sp.NS;
            const unsigned blocks = static_cast<unsigned>((sp.periods + sp.per_block - 1) / sp.per_block);
            if (cb == 16)
                resample_stream_kernel<16><<<blocks, threads, 0, s>>>(sp);
            else if (cb == 8)
                resample_stream_kernel<8><<<blocks, threads, 0, s>>>(sp);
            else
                resample_stream_kernel<4><<<blocks, threads, 0, s>>>(sp);
            MSS_CUDA(cudaGetLastError());
            return MSS_OK;
        }
    }
    p.n_chunks = (out_dims[2] + 15) / 16;
    p.slot_pitch = (in_dims[2] + 15 + 12 + 15) / 16 * 16;  // row + alignment offset + the 3-word gather window
    constexpr int kMaxStage = 96 * 1024;
    p.rp = p.n_chunks <= 256 ? 256 / p.n_chunks : 0;
    if (p.rp > kMaxStage / (p.slot_pitch * kRowsPerThread)) p.rp = kMaxStage / (p.slot_pitch * kRowsPerThread);
    if (p.rp >= 1) {
        // about 4 CTAs per SM's worth of row ranges, each a multiple of the pass size
        long long per_block = (n_rows + 148LL * 4 - 1) / (148LL * 4);
        const int pass_rows = p.rp * kRowsPerThread;
        per_block = (per_block + pass_rows - 1) / pass_rows * pass_rows;
        p.rows_per_block = static_cast<int>(per_block);
        const long long blocks = (n_rows + per_block - 1) / per_block;
        const size_t smem = static_cast<size_t>(pass_rows) * p.slot_pitch;
        MSS_CUDA(cudaFuncSetAttribute(resample_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxStage));  // per device
        resample_rows_kernel<<<static_cast<unsigned>(blocks), 256, smem, s>>>(p);
    } else {
        p.rows_per_block = 0;
        const long long total = n_rows * p.n_chunks;
        long long blocks = (total + 255) / 256;
        if (blocks > 148LL * 16) blocks = 148LL * 16;
        resample_gather_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(p);
    }
    MSS_CUDA(cudaGetLastError());
    return MSS_OK;
}
