// Streaming first-max argmax with top-2 tracking, shared by the fused accumulate and the finalize kernel.
//
// engine/test.py:140-141 takes np.argmax (first maximum) of softmax(logits).  softmax is monotone, so
// the label is the first-max argmax of the normalised logits themselves except where float32 rounding
// inside softmax merges two classes whose logits differ by < ~1.2e-7; such voxels are far inside the
// near-tie tolerance of the parity criterion and are what `near_tie` counts.  A NaN or +inf logit turns
// the whole softmax row into NaN, for which np.argmax answers 0 - reproduced via `poisoned`.
#pragma once

namespace mss {

struct ArgmaxState {
    float best, second;
    int idx;
    bool poisoned;
    __device__ __forceinline__ void reset() {
        best = __int_as_float(0xff800000);  // -inf
        second = best;
        idx = 0;
        poisoned = false;
    }
    __device__ __forceinline__ void push(float v, int k) {
        poisoned |= !(v < __int_as_float(0x7f800000));  // NaN or +inf
        idx = v > best ? k : idx;                       // strictly greater: the first maximum keeps its index
        second = fmaxf(second, fminf(best, v));         // branch-free top-2 (a NaN leaves both untouched)
        best = fmaxf(best, v);
    }
    __device__ __forceinline__ int label() const { return poisoned ? 0 : idx; }
    // top-2 gap relative to the larger magnitude strictly below tol
    __device__ __forceinline__ bool near_tie(float tol) const {
        const float gap = best - second;
        const float scale = fmaxf(fabsf(best), fabsf(second));
        return !poisoned && gap < tol * scale;
    }
};

}  // namespace mss
