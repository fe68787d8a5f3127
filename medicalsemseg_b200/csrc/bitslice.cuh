// Bit-sliced views of small-integer label maps, shared by the vote and Dice kernels.
//
// 32 uint8 labels, all < 16, arrive as eight 32-bit words (two 16-byte loads).  Packing word i with
// word i+4 as low/high nibbles gives four words of eight 4-bit labels; a 4x4 bit-matrix transpose over
// those four words (two rounds of masked swaps) leaves plane j holding bit j of all 32 labels.  The
// voxel -> bit position mapping is a fixed permutation (voxel 4i+b of the chunk sits at bit
// 8b + 4(i>>2) + (i&3)), which is irrelevant for counting and is undone by `labels_from_bitplanes32`.
//
// Everything here is plain integer code marked __host__ __device__, so tests/csrc/host_chunks.cu can run
// the very same per-chunk logic on the CPU against the oracle (no GPU in the build container).
#pragma once

#ifdef __CUDACC__
#define MSS_HD __host__ __device__ __forceinline__
#else
#define MSS_HD inline
#endif

namespace mss {

MSS_HD int popc32(unsigned x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

MSS_HD void transpose4x4(unsigned (&n)[4]) {
    unsigned t;
    t = ((n[0] >> 2) ^ n[2]) & 0x33333333u, n[2] ^= t, n[0] ^= t << 2;
    t = ((n[1] >> 2) ^ n[3]) & 0x33333333u, n[3] ^= t, n[1] ^= t << 2;
    t = ((n[0] >> 1) ^ n[1]) & 0x55555555u, n[1] ^= t, n[0] ^= t << 1;
    t = ((n[2] >> 1) ^ n[3]) & 0x55555555u, n[3] ^= t, n[2] ^= t << 1;
}

// w: 32 labels (< 16) as loaded; q[j]: bit j of every label
MSS_HD void bitplanes32(const unsigned (&w)[8], unsigned (&q)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = w[i + 4] * 16u + w[i];  // nibble pack (IMAD: the FMA pipe is idle otherwise)
    transpose4x4(q);
}

// inverse of bitplanes32: planes of 4-bit labels -> the eight words of 32 uint8 labels
MSS_HD void labels_from_bitplanes32(const unsigned (&q)[4], unsigned (&w)[8]) {
    unsigned n[4] = {q[0], q[1], q[2], q[3]};
    transpose4x4(n);  // the transpose is an involution
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        w[i] = n[i] & 0x0F0F0F0Fu;
        w[i + 4] = (n[i] >> 4) & 0x0F0F0F0Fu;
    }
}

// the four minterms of two planes: m[v] has a bit set where (bit_lo, bit_hi) spell the 2-bit value v
MSS_HD void pair_minterms(unsigned lo, unsigned hi, unsigned (&m)[4]) {
    m[0] = ~(lo | hi);
    m[1] = lo & ~hi;
    m[2] = ~lo & hi;
    m[3] = lo & hi;
}

// true when one of the 8 words holds a byte >= 16 (the chunk cannot be nibble-packed)
MSS_HD bool has_wide_label(const unsigned (&w)[8]) {
    return ((w[0] | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7]) & 0xF0F0F0F0u) != 0u;
}

// ---- carry-save population count of N one-bit planes into sliced sum bits s[BIT], s[BIT+1], ... ----
MSS_HD unsigned maj3(unsigned a, unsigned b, unsigned c) { return (a & b) | (a & c) | (b & c); }

template <int N, int BIT, int NB>
MSS_HD void sliced_count(const unsigned (&in)[N], unsigned (&s)[NB]) {
    constexpr int NC = N / 2;
    unsigned carry[NC > 0 ? NC : 1];
    unsigned acc = in[0];
#pragma unroll
    for (int i = 1; i + 1 < N; i += 2) {  // full adders
        carry[(i - 1) / 2] = maj3(acc, in[i], in[i + 1]);
        acc = acc ^ in[i] ^ in[i + 1];
    }
    if (N % 2 == 0) {  // one half adder left
        carry[NC - 1] = acc & in[N - 1];
        acc ^= in[N - 1];
    }
    s[BIT] = acc;
    if constexpr (NC > 0) sliced_count<NC, BIT + 1, NB>(carry, s);
}

constexpr int vote_count_bits(int m) { return m >= 8 ? 4 : (m >= 4 ? 3 : (m >= 2 ? 2 : 1)); }

// Majority vote (majority_vote.py:23-37) over one chunk of 32 voxels of M maps, all labels < 16:
// background holds one vote, class c >= 1 one vote per map equal to c, first maximum wins.
template <int M>
MSS_HD void vote_chunk(const unsigned (&w)[M][8], int K, unsigned (&res)[8]) {
    constexpr int NB = vote_count_bits(M);
    unsigned lo[M][4], hi[M][4];  // minterms of label bits (0,1) and (2,3) per map
#pragma unroll
    for (int m = 0; m < M; ++m) {
        unsigned q[4];
        bitplanes32(w[m], q);
        pair_minterms(q[0], q[1], lo[m]);
        pair_minterms(q[2], q[3], hi[m]);
    }
    unsigned best[NB], lab[4] = {0u, 0u, 0u, 0u};
    best[0] = 0xffffffffu;  // background holds one vote everywhere
#pragma unroll
    for (int j = 1; j < NB; ++j) best[j] = 0u;
#pragma unroll
    for (int c = 1; c < 16; ++c) {
        if (c >= K) break;  // uniform
        unsigned ind[M], s[NB];
#pragma unroll
        for (int m = 0; m < M; ++m) ind[m] = lo[m][c & 3] & hi[m][c >> 2];
        sliced_count<M, 0, NB>(ind, s);
        unsigned gt = 0u;  // votes[c] > running maximum, least significant bit first
#pragma unroll
        for (int j = 0; j < NB; ++j) gt = (s[j] & ~best[j]) | (~(s[j] ^ best[j]) & gt);
#pragma unroll
        for (int j = 0; j < NB; ++j) best[j] = (gt & s[j]) | (~gt & best[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) lab[j] = ((c >> j) & 1) ? (lab[j] | gt) : (lab[j] & ~gt);
    }
    labels_from_bitplanes32(lab, res);
}

// Dice counts of one chunk of 32 voxels (pred words pw, label words yw, all < 16) added into per-thread counters:
// tp/pp/yy[i] hold classes 2i (low 16 bits) and 2i+1 (high 16 bits), KP = ceil(K / 2) pairs (compile time, so
// the POPCs of all classes interleave).  The half of class K (odd K) collects counts that are never reported.
template <int KP>
MSS_HD void dice_chunk(const unsigned (&pw)[8], const unsigned (&yw)[8], unsigned (&tp)[8], unsigned (&pp)[8],
                       unsigned (&yy)[8]) {
    unsigned qp[4], qy[4], ap[4], bp[4], ay[4], by[4];
    bitplanes32(pw, qp);
    bitplanes32(yw, qy);
    pair_minterms(qp[0], qp[1], ap);
    pair_minterms(qp[2], qp[3], bp);
    pair_minterms(qy[0], qy[1], ay);
    pair_minterms(qy[2], qy[3], by);
#pragma unroll
    for (int c2 = 0; c2 < KP; ++c2) {
        const int c = 2 * c2;
        const unsigned mp0 = ap[c & 3] & bp[c >> 2], my0 = ay[c & 3] & by[c >> 2];
        const unsigned mp1 = ap[(c + 1) & 3] & bp[(c + 1) >> 2], my1 = ay[(c + 1) & 3] & by[(c + 1) >> 2];
        pp[c2] = popc32(mp1) * 65536u + (pp[c2] + popc32(mp0));  // IADD + IMAD
        yy[c2] = popc32(my1) * 65536u + (yy[c2] + popc32(my0));
        tp[c2] = popc32(mp1 & my1) * 65536u + (tp[c2] + popc32(mp0 & my0));
    }
}

}  // namespace mss
