/*
 * mss_b200.h - C ABI of libmss_b200.so: the B200 (sm_100a) kernels behind the sliding-window
 * volumetric inference path of zouyunkai/MedicalSemSeg.
 *
 * The reference has no native/FFI boundary for this path: it is Python calling Python
 * (engine/utils.py:19-159 -> MONAI helpers + ATen ops).  This header is therefore the boundary the
 * new host code binds (medicalsemseg_b200/_lib.py, ctypes); every entry point names the reference
 * lines whose work it takes over.  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void*.
 *   - every device buffer is owned by the caller (PyTorch's caching allocator in practice); the
 *     library allocates nothing on the device and keeps no global mutable device state.
 *   - all launches are stream-ordered on `stream`; no entry point synchronises the host.
 *   - return value: 0 = ok, < 0 = argument error (MSS_E_*), > 0 = a cudaError_t.  A text for the
 *     last failing call of the calling thread is available from mss_last_error().
 *   - spatial dims are always ordered (D, H, W) with W the fastest-varying one (NCDHW).
 */
#ifndef MSS_B200_H_
#define MSS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSS_ABI_VERSION 1

#define MSS_OK 0
#define MSS_E_ARG (-1)         /* null pointer / non-positive size / inconsistent layout            */
#define MSS_E_UNSUPPORTED (-2) /* valid request the kernels do not implement (e.g. K > MSS_MAX_CLASSES) */
#define MSS_E_ALIGN (-3)       /* a pointer or pitch violates the documented alignment              */
#define MSS_E_DRIVER (-4)      /* CUDA driver entry point (tensor-map encode) unavailable            */

#define MSS_MAX_CLASSES 32       /* K for the fused label path; accumulate itself takes any K      */
#define MSS_MAX_BATCH_PTRS 640   /* predictor batches per mss_accumulate call                      */
#define MSS_MAX_VOTE_MAPS 15     /* ensemble size M (4-bit vote counters)                          */
#define MSS_MAX_VOTE_CLASSES 16

/* dtype of the predictor's logits (engine/utils.py:135; fp16/bf16 under autocast, engine/test.py:127) */
#define MSS_F32 0
#define MSS_F16 1
#define MSS_BF16 2

/* importance-map modes (MONAI BlendMode, engine/utils.py:26,113-115) */
#define MSS_BLEND_CONSTANT 0
#define MSS_BLEND_PROFILES 1 /* outer product of three caller-supplied 1-D profiles              */

/* 1-D profile variants generated on the device by mss_gaussian_profile */
#define MSS_GAUSS_MONAI08_ERF 0 /* MONAI 0.8: erf-integrated taps centred at i//2 (oracle/monai08.py) */
#define MSS_GAUSS_MONAI12_EXP 1 /* MONAI >= 1.2: exp(-x^2/2s^2) on the half-integer grid            */

/* what mss_accumulate does with a voxel once its last covering window has been applied */
#define MSS_FUSE_NONE 0   /* keep raw weighted sums (multi-GPU slabs: halo exchange comes first)  */
#define MSS_FUSE_LOGITS 1 /* store sum / count  (engine/utils.py:151)                             */
#define MSS_ACC_PATH_GENERAL 0 /* mss_accumulate_last_path(): per-quad general kernel               */
#define MSS_ACC_PATH_CELLS 1   /* cell-uniform kernel (per-thread cp.async rings)                   */
#define MSS_ACC_PATH_ROWS 2    /* row-staged kernel (bulk copies; K <= 4, one launch, labels out)   */
#define MSS_FUSE_LABELS 2 /* store only the uint8 argmax label (engine/test.py:140-141)          */

/*
 * Window layout of one stitched volume (or of one rank's slab of it).
 *
 * `image` is the stitched image size after the reference's pad-to-roi (engine/utils.py:97); window
 * starts per axis come from MONAI dense_patch_slices semantics (engine/utils.py:105-110) and live,
 * together with per-coordinate cover ranges, in a table built by mss_geom_table_build().  A buffer
 * handed to the kernels covers the global box [origin, origin + extent) and owns the windows of the
 * index box [win_lo, win_hi) (single GPU: everything).  Owned windows are enumerated in C order
 * (D slowest, W fastest) per volume, volumes outermost - the order of engine/utils.py:120-125.
 */
typedef struct mss_layout {
    int32_t image[3];
    int32_t roi[3];
    int32_t n_starts[3];
    int32_t win_lo[3];
    int32_t win_hi[3];
    int32_t origin[3];
    int32_t extent[3];
    int32_t pitch_w;   /* elements per W-row of accumulator / label / count buffers (>= extent[2])   */
    int32_t n_volumes; /* Nb                                                                         */
    int32_t n_classes; /* K                                                                          */
    const int32_t* table_host; /* table from mss_geom_table_build (host copy, read by the wrappers)  */
    const int32_t* table_dev;  /* the same table in device memory (read by the kernels)              */
} mss_layout_t;

/* ---- host-only helpers (no GPU needed) ------------------------------------------------------ */

int mss_abi_version(void);
const char* mss_last_error(void);

/* Window starts along one axis: MONAI dense_patch_slices per-axis rule (engine/utils.py:108):
 * n = 1 + min{d >= 0 : d*interval + roi >= image}, start_k = k*interval - max(k*interval+roi-image, 0).
 * Writes up to `cap` starts, returns the count (or MSS_E_ARG). */
int mss_axis_starts(int32_t image, int32_t roi, int32_t interval, int32_t* starts_out, int32_t cap);

/* Number of int32 entries of the geometry table for this image / window grid. */
int64_t mss_geom_table_len(const int32_t image[3], const int32_t n_starts[3]);

/* Build the geometry table (header, per-axis starts, per-coordinate cover ranges) into host memory. */
int mss_geom_table_build(const int32_t image[3], const int32_t roi[3], const int32_t n_starts[3],
                         const int32_t* starts_d, const int32_t* starts_h, const int32_t* starts_w,
                         int32_t* table_out, int64_t table_len);

/* ---- kernels ------------------------------------------------------------------------------- */

/* 1-D gaussian profile of length n on the device (MONAI compute_importance_map's per-axis factor,
 * engine/utils.py:113-115).  variant: MSS_GAUSS_*.  `profile_out` is device memory, n floats. */
int mss_gaussian_profile(float* profile_out, int32_t n, float sigma, int32_t variant, void* stream);

/* Importance map [roi_d, roi_h, roi_w] fp32 (engine/utils.py:113-115): ones, or
 * floor_abs <= 0 (MONAI 0.8): clamp_min(((p_d[i]*p_h[j])*p_w[k]) / max, smallest non-zero entry);
 * floor_abs > 0 (MONAI >= 1.2): clamp_min((p_d[i]*p_h[j])*p_w[k], max(min entry, floor_abs)).  `profiles` are device
 * pointers; `scratch` is 8 bytes of device memory. */
int mss_importance_map(float* map_out, const int32_t roi[3], int32_t mode, const float* prof_d,
                       const float* prof_h, const float* prof_w, float floor_abs, void* scratch,
                       void* stream);

/* Patch extraction (engine/utils.py:122-133): gathers owned windows [first_window, first_window +
 * n_windows) of `lay` into patches_out[n_windows, Cin, roi_d, roi_h, roi_w] fp32 and writes their
 * relative centres (engine/utils.py:126-130) into centers_out[n_windows, 3].
 * `volume` is [Nb, Cin, vol_extent] fp32, contiguous; its voxel (0,0,0) sits at stitched-frame
 * coordinate vol_origin (= the reference's pad offsets, engine/utils.py:98-103); anything outside
 * reads as `cval`.  Three kernels: TMA 3-D tiled copies (W, roi_w multiples of 4, 16-byte aligned bases, every
 * window inside the volume and starting on a multiple of 4 along W); a shifted-vector copy for the same layout
 * with arbitrary W starts; scalar loads for everything else (constant pad, odd row pitch).
 * use_tma: 1 = auto, 0 = never TMA (shifted-vector or scalar), 2 = scalar kernel only, 3 = the volume-stationary kernel
 * (cp.async.bulk: every volume row read once and written to all its windows; same layout conditions as TMA, the whole
 * window grid owned, <= 8 window positions along W) where it applies, else auto. */
int mss_extract_patches(const float* volume, const int32_t vol_origin[3], const int32_t vol_extent[3],
                        int32_t n_channels, float cval, const mss_layout_t* lay, int64_t first_window,
                        int32_t n_windows, float* patches_out, float* centers_out, int32_t use_tma,
                        void* stream);

/* Weighted overlap accumulation (engine/utils.py:137-151), output-stationary and atomics-free.
 * Applies owned windows [first_window, first_window + n_windows) whose logits live in `n_batches`
 * predictor outputs (batch_ptrs[i] -> [sw_batch, K, roi] of logits_dtype; the last batch may be
 * ragged) to the fp32 accumulator acc[Nb, K, extent_d, extent_h, pitch_w]: every covered voxel adds
 * w*logit per window in ascending window order with separately rounded multiply and add, reads the
 * accumulator only if an earlier window touched it and writes it once.  `fuse`: MSS_FUSE_*; for
 * voxels completed by this call MSS_FUSE_LOGITS stores sum/count (count = ascending fp32 sum of the
 * weights of ALL covering windows of the grid), MSS_FUSE_LABELS stores the first-max argmax into
 * labels[Nb, extent_d, extent_h, label_pitch_w] and bumps near_ties (uint64, device) for voxels
 * whose top-2 relative gap is < tie_tol.  acc may be NULL only with MSS_FUSE_LABELS when the call
 * covers every owned window.
 * Three kernels sit behind this entry point, all with the same operation order (bit-identical sums): the row-staged
 * kernel (fp32 logits, K <= 16, one call covering every window with a fused mode, <= 4 window positions along W: whole
 * window rows staged by cp.async.bulk into a shared-memory ring), the cell-uniform kernel (per-thread cp.async rings;
 * several calls per volume, raw sums, 16-bit logits, wide grids) and the general per-quad kernel (geometries beyond the
 * cell tables). */
int mss_accumulate(const mss_layout_t* lay, const void* const* batch_ptrs, int32_t n_batches,
                   int32_t sw_batch, int32_t logits_dtype, int64_t first_window, int64_t n_windows,
                   const float* importance_map, float* acc, int32_t fuse, uint8_t* labels,
                   int32_t label_pitch_w, float tie_tol, unsigned long long* near_ties, void* stream);

/* The raw-sum form of mss_accumulate for a buffer whose window box is SHARED with other ranks (multi-GPU: the window
 * list of engine/utils.py:120-125 is cut into one contiguous range per rank): only windows [own_first, own_first +
 * own_count) of the box exist for this buffer, the others are ignored entirely.  Always MSS_FUSE_NONE: acc receives
 * raw weighted sums of this rank's windows (voxels none of them covers are not written - clear the buffer first).
 * [first_window, +n_windows) must lie inside the owned range. */
int mss_accumulate_range(const mss_layout_t* lay, const void* const* batch_ptrs, int32_t n_batches,
                         int32_t sw_batch, int32_t logits_dtype, int64_t first_window, int64_t n_windows,
                         int64_t own_first, int64_t own_count, const float* importance_map, float* acc,
                         void* stream);

/* Which kernel the calling thread's last successful mss_accumulate / mss_accumulate_range call launched (MSS_ACC_PATH_*;
 * -1 before the first call): lets tests and benches state which path they measured.  No reference counterpart. */
int mss_accumulate_last_path(void);

/* Normalise + softmax-argmax -> uint8 (engine/utils.py:151 + engine/test.py:140-141) over the local
 * box [box_lo, box_hi) of `logits[Nb, K, extent_d, extent_h, pitch_w]`.  normalise != 0 divides by
 * the window-weight count first (computed on the fly from the table and the importance map);
 * logits_out (optional, may alias `logits`) receives the normalised values, probs_out (optional)
 * the softmax probabilities. */
int mss_finalize_labels(const mss_layout_t* lay, const float* logits, const float* importance_map,
                        int32_t normalise, const int32_t box_lo[3], const int32_t box_hi[3],
                        uint8_t* labels, int32_t label_pitch_w, float* logits_out, float* probs_out,
                        float tie_tol, unsigned long long* near_ties, void* stream);

/* Multi-GPU finalise that IS the exchange (no counterpart in the single-process reference; it is the `+=` of
 * engine/utils.py:147 between ranks followed by :151 and engine/test.py:140-141).  `lay` describes the box this rank OWNS
 * (origin / extent in the global frame, origin W a multiple of 4) over the GLOBAL window grid.  src_acc[i] is the raw-sum
 * accumulator [Nb, K, extent_i, pitch_i] of rank i over its box src_origin[3i..] / src_extent[3i..] (W origin and pitch
 * multiples of 4, zero where the rank accumulated nothing) - its own memory or a peer's mapped over NVLink (torch
 * symmetric memory).  Every owned voxel adds the sources that hold it in ascending order, then either stores the first-max
 * argmax into labels[Nb, extent, label_pitch_w] or (logits_out != NULL, [Nb, K, extent_d, extent_h, lay->pitch_w]) also
 * the sums divided by the window-weight count of the global grid. */
int mss_finalize_gather(const mss_layout_t* lay, int32_t n_src, const float* const* src_acc, const int32_t* src_origin,
                        const int32_t* src_extent, const int32_t* src_pitch_w, const float* importance_map,
                        uint8_t* labels, int32_t label_pitch_w, float* logits_out, float tie_tol,
                        unsigned long long* near_ties, void* stream);

/* Ensemble majority vote (majority_vote.py:23-37): votes[0] = 1, votes[c>=1] = #{m : map_m == c},
 * labels >= n_classes ignored, first-max argmax.  maps[i] are device pointers to n_voxels uint8. */
int mss_majority_vote(const uint8_t* const* maps, int32_t n_maps, int32_t n_classes, int64_t n_voxels,
                      uint8_t* voted_out, void* stream);

/* Per-class Dice counts (MONAI DiceMetric inputs, engine/test.py:50-56): counts[0][c] = #(pred==c &
 * label==c), counts[1][c] = #(pred==c), counts[2][c] = #(label==c) as int64[3][n_classes], ADDED to
 * the existing contents of `counts` (device).  label_dtype: 0 = uint8, 1 = float32 (integer valued,
 * as the reference's loaders provide).  Values outside [0, n_classes) are counted nowhere. */
int mss_dice_counts(const uint8_t* pred, const void* label, int32_t label_dtype, int64_t n_voxels,
                    int32_t n_classes, long long* counts, void* stream);

/* The same for a batch: pred / label hold n_volumes maps of n_voxels each, back to back; counts is
 * int64[n_volumes][3][n_classes] (ADDED to).  One launch for all volumes a rank evaluates (BASELINE.json
 * configs[4]: per-volume Dice, engine/test.py:59-69), so the GPU sees enough work to run at HBM speed. */
int mss_dice_counts_batched(const uint8_t* pred, const void* label, int32_t label_dtype, int64_t n_voxels,
                            int64_t n_volumes, int32_t n_classes, long long* counts, void* stream);

/* Halo reduction for z-slab partitioning: dst[i] += src[i] over a [n_rows, row_len] fp32 block with
 * independent row pitches (the receiving rank adds its neighbour's partial sums). */
int mss_halo_add(float* dst, int64_t dst_pitch, const float* src, int64_t src_pitch, int64_t n_rows,
                 int64_t row_len, void* stream);

/* The same reduction over a 5-D box [dims[0..3], row_len] whose innermost dimension is contiguous in both
 * operands and whose outer dimensions have arbitrary element strides: any face of a block partition in ONE
 * launch.  `src` may be a peer GPU's accumulator mapped into this process (NVLink peer memory, e.g. a
 * torch symmetric-memory buffer): the add then IS the transfer - no staging copy, no send/recv. */
int mss_halo_add_nd(float* dst, const int64_t dst_strides[4], const float* src, const int64_t src_strides[4],
                    const int64_t dims[4], int64_t row_len, void* stream);

/* ---- after the argmax: back to the original voxel grid ---------------- */

/* Per-axis source-index table of scipy.ndimage.zoom(order=0, prefilter=False, mode='constant') as the
 * reference calls it (utils/misc.py:420-425): table_out[k] = floor(k*zoom + 0.5) with
 * zoom = (n_in-1)/(n_out-1) in float64 (1.0 when n_out == 1), or -1 where k*zoom leaves [0, n_in-1]
 * (scipy then writes cval = 0; rounding makes this hit the last index of some size pairs).  Host only. */
int mss_zoom_index_table(int32_t n_in, int32_t n_out, int32_t* table_out);

/* Nearest-neighbour resampling of uint8 label maps [n_volumes, in_dims] -> [n_volumes, out_dims]
 * (utils/misc.py:420-425 resample_3d, called at engine/test.py:143-147): out[x,y,z] =
 * in[index_x[x], index_y[y], index_z[z]], 0 where any index is -1.  index_* are DEVICE int32 tables of
 * out_dims[a] entries, as produced by mss_zoom_index_table. */
int mss_resample_nearest(const uint8_t* labels_in, const int32_t in_dims[3], uint8_t* labels_out,
                         const int32_t out_dims[3], int64_t n_volumes, const int32_t* index_x,
                         const int32_t* index_y, const int32_t* index_z, void* stream);

/* ---- before the sliding window: test-time intensity transforms -------- */

#define MSS_INT_CBRT 1     /* x = cbrt(x)                          data/transforms.py:54 (ScaleCubedIntensityRange) */
#define MSS_INT_SCALE 2    /* x = (x - a_min) / (a_max - a_min)    data/transforms.py:62 / MONAI ScaleIntensityRange */
#define MSS_INT_RESCALE 4  /* x = x * (b_max - b_min) + b_min      data/transforms.py:63-64                          */
#define MSS_INT_CLIP_LO 8  /* x = max(x, b_min)                    data/transforms.py:65-66                          */
#define MSS_INT_CLIP_HI 16 /* x = min(x, b_max)                                                                      */
#define MSS_INT_NORM 32    /* x = (x - subtrahend) / divisor       MONAI NormalizeIntensity, data/dataset_builder.py:352-368 */
#define MSS_INT_NONZERO 64 /* ... only where x != 0 (nonzero=True)                                                   */
#define MSS_INT_F64 128    /* evaluate in float64, round to float32 once (NumPy >= 2 with float64 scalar bounds)     */

/* The chain data/dataset_builder.py:322-370 applies to the test volume, fused into one elementwise pass over
 * n float32 voxels (in may equal out).  Steps are applied in the order of the flag values above, each as a
 * separately rounded operation of the working type (float32; float64 with MSS_INT_F64).  a_max_minus_a_min is
 * (a_max - a_min) as the host language computed it (Python: a float64 difference); constants are rounded to the
 * working type once. */
int mss_intensity_transform(const float* in, float* out, int64_t n, int32_t flags, double a_min,
                            double a_max_minus_a_min, double b_min, double b_max, double subtrahend, double divisor,
                            void* stream);

/* ---- mirror test-time augmentation of the nnU-Net-style predictor ------ */

/* out = flip(in) over the spatial axes named by mirror_mask (bit 0 = W, bit 1 = H, bit 2 = D: torch.flip dims
 * 4, 3, 2 at models/segmentors/nnformer_official/neural_network.py:541-565) for n_outer = batch*channels
 * volumes of dims (D, H, W).  Not in place. */
int mss_flip_copy(const float* in, float* out, int64_t n_outer, const int32_t dims[3], int32_t mirror_mask,
                  void* stream);

/* out = sum over m, in order, of fadd_rn(out, fmul_rn(scale, flip_{mask[m]}(preds[m]))) starting from 0:
 * `result_torch += 1 / num_results * torch.flip(pred, ...)` of neural_network.py:537-565 for all mirrors in
 * one pass.  preds[m] are device pointers to [n_outer, D, H, W] fp32; n_terms <= 8; out must not alias. */
int mss_mirror_merge(const float* const* preds, const int32_t* mirror_masks, int32_t n_terms, float scale,
                     float* out, int64_t n_outer, const int32_t dims[3], void* stream);

/* Half-precision aggregation of the tiler's `all_in_gpu` branch (neural_network.py:346-372, :399-406, :420-423): importance
 * map, aggregated results and counts are binary16 there.  batch_ptrs[i] -> fp32 [sw_batch, K, roi] tile predictions (the
 * output of mss_mirror_merge, NOT yet weighted), ALL windows of `lay` in one call; importance_map_half_valued is the fp32 copy
 * of the map after `.half()` (every value exactly representable in binary16).  Per voxel and class, in tile order:
 * agg = half(float(agg) + float(half(pred * w))), nb = half(float(nb) + w); probs_out[Nb, K, extent, pitch_w] receives
 * float(half(agg / nb)) and labels (optional) the first-max argmax. */
int mss_accumulate_half(const mss_layout_t* lay, const void* const* batch_ptrs, int32_t n_batches, int32_t sw_batch,
                        const float* importance_map_half_valued, float* probs_out, uint8_t* labels, int32_t label_pitch_w,
                        void* stream);

/* ---- 95th-percentile Hausdorff distance (engine/test.py:31,55-57) -------- */

/* Bounding boxes of (pred == c) | (label == c) for every class c < n_classes (<= 32) of two uint8 label maps [dims] in one
 * pass (MONAI generate_spatial_bounding_box per class inside get_mask_edges).  boxes_out: device int32 [n_classes][6] =
 * {lo_d, lo_h, lo_w, hi_d, hi_h, hi_w} (hi exclusive), which the caller initialises to {2^30, 2^30, 2^30, 0, 0, 0}: a
 * class in neither map keeps lo > hi. */
int mss_class_boxes(const uint8_t* pred, const uint8_t* label, const int32_t dims[3], int32_t n_classes,
                    int32_t* boxes_out, void* stream);

/* Two order statistics of (a masked subset of) a 32-bit array on the device, without a sort and without a host sync:
 * three radix-histogram levels (11 + 11 + 10 bits).  keys: n_elems values, key_kind 0 = unsigned / non-negative int32 as
 * stored, 1 = float32 (mapped to order-preserving unsigned); mask (optional uint8, non-zero = take part).  rank_mode 0:
 * the k_lo-th and k_hi-th smallest (0-based) of the valid elements; rank_mode 1: the two neighbours numpy.percentile
 * (linear) interpolates for quantile `quant` in [0, 1]: floor((n - 1) * quant) and the next one, n = valid count.
 * scratch: mss_select_scratch_bytes() of device memory (8-byte aligned); out (device, uint64[4]) = {n, key_lo, key_hi,
 * key_max} in the mapped unsigned domain.  Behind np.percentile at engine/test.py:55-57 (HD95) and
 * data/dataset_builder.py:343-353 (ScaleIntensityRangePercentiles). */
int64_t mss_select_scratch_bytes(void);
int mss_select2(const void* keys, int32_t key_kind, const uint8_t* mask, int64_t n_elems, int32_t rank_mode, double quant,
                int64_t k_lo, int64_t k_hi, void* scratch, unsigned long long* out, void* stream);

/* Surface voxels of class `cls` inside the box [box_lo, box_hi) of a uint8 label map [dims] (MONAI
 * get_mask_edges: binary_erosion XOR mask on the bounding box of pred | gt; outside the box = background; axes
 * along which the box is one voxel thick are ignored, as the reference squeezes them away).  Writes
 * edges_out[box] (0/1) and, when edt_input_out is not NULL, the int32 distance-transform input (0 on edges, 2^29
 * elsewhere; not needed when the first pass is mss_edt_pass_mask). */
int mss_mask_edges(const uint8_t* labels, const int32_t dims[3], int32_t cls, const int32_t box_lo[3],
                   const int32_t box_hi[3], uint8_t* edges_out, int32_t* edt_input_out, void* stream);

/* One axis of the exact squared Euclidean distance transform (scipy distance_transform_edt, squared, as
 * MONAI get_surface_distance uses it): out[x] = min_i (x - i)^2 + in[i] along `axis` of an int32 volume [dims];
 * entries >= 2^29 mean "no feature".  scratch_s / scratch_t: int32 buffers of the volume's size.  Apply to axes
 * 0, 1, 2 in turn (any order), ping-ponging in/out.  Not in place. */
int mss_edt_pass(const int32_t* in, int32_t* out, int32_t* scratch_s, int32_t* scratch_t, const int32_t dims[3],
                 int32_t axis, void* stream);

/* The first pass along the CONTIGUOUS axis (2) straight from a uint8 feature mask [dims] without an envelope scan: out[x] =
 * (distance to the nearest non-zero voxel of the same row)^2, 2^29 when the row has none.  One warp per row (ballot scan).
 * Follow with mss_edt_pass along axes 1 and 0: the result is laid out like the input, no transposes. */
int mss_edt_row_mask(const uint8_t* mask, int32_t* out, const int32_t dims[3], void* stream);

/* The first pass straight from a uint8 feature mask [dims] (non-zero = feature voxel) along any axis: saves writing and
 * reading the int32 input volume. */
int mss_edt_pass_mask(const uint8_t* mask, int32_t* out, int32_t* scratch_s, int32_t* scratch_t,
                      const int32_t dims[3], int32_t axis, void* stream);

/* Sums behind the evaluation loss DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True) of
 * run_evaluation.py:53 / engine/test.py:48, in one pass over the stitched logits (class c, row r, column x at
 * logits[c*class_stride + r*row_pitch + x]) and the label map [n_rows, row_len] (label_dtype 0 = uint8,
 * 1 = float32): ADDS into sums[3K+1] (float64, device) I[c] = sum softmax_c [y==c], P[c] = sum softmax_c^2
 * (softmax_c when squared_pred == 0), G[c] = #(y==c), and CE = sum -log softmax_y.  K <= 16. */
int mss_dice_ce_sums(const float* logits, int64_t class_stride, int64_t row_pitch, int64_t n_rows, int32_t row_len,
                     int32_t n_classes, const void* labels, int32_t label_dtype, int32_t squared_pred, double* sums,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSS_B200_H_ */
