/* values_b200.h -- C-ABI of the B200-native ValUES C2+C3 uncertainty hot path.
 *
 * The reference (IML-DKFZ/values) is pure Python and has no FFI layer; its boundary for
 * this path is a set of Python callables.  Every entry point below
 * names the reference callable it replaces.  Conventions:
 *   - all data pointers are DEVICE pointers unless the name ends in `_host`;
 *   - sizes / strides are in ELEMENTS, int64_t; the voxel axis is contiguous (stride 1);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - nothing is allocated or owned by the library: outputs and workspaces belong to the
 *     caller; workspace sizes come from the matching *_workspace_bytes function;
 *   - return value: VALUES_OK or a negative VALUES_ERR_* code; values_last_error() gives
 *     a thread-local message.  No entry point synchronises the device.
 * There is deliberately NO CPU fallback: without a CUDA device every compute entry point
 * returns VALUES_ERR_CUDA.
 */
#ifndef VALUES_B200_H
#define VALUES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VALUES_ABI_VERSION 8

typedef enum {
    VALUES_F32 = 0, VALUES_F64 = 1, VALUES_BF16 = 2,
    /* label / segmentation maps (K4 statistics only) */
    VALUES_U8 = 3, VALUES_I32 = 4, VALUES_I64 = 5
} values_dtype_t;

#define VALUES_OK 0
#define VALUES_ERR_INVALID_ARG (-1) /* -> ValueError on the Python side            */
#define VALUES_ERR_UNSUPPORTED (-2) /* shape outside what the kernels tile for      */
#define VALUES_ERR_CUDA (-3)        /* a CUDA runtime call / launch failed          */
#define VALUES_ERR_WORKSPACE (-4)   /* workspace pointer NULL or too small          */

int values_abi_version(void);
const char* values_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py reports it). */
int64_t values_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * K1: fused N x C reduction.
 * Replaces calculate_uncertainty(softmax_preds, ssn)  uncertainty_modeling/test_3D.py:486-518
 * (PE / EE / MI with the NaN-skip rule, fp32 accumulators in class order) and the arg-max
 * of the mean / of each sample taken at save time (data_carrier_3D.py:253-259, 281-285;
 * test_2D.py:119-127).  One sweep over HBM.
 *
 *   probs   [B, N, C, V] with element strides (stride_b, stride_n, stride_c, 1)
 *   pe/ee/mi      float [B, V]   predictive entropy / expected entropy / mutual information
 *                               (the `ssn` key swap is a host-side relabelling); volume b of each
 *                               map starts map_stride_b elements after volume b-1 (0 = V), so
 *                               pe, ee = pe + V, mi = pe + 2V with map_stride_b = 3V gives the
 *                               volume-major layout [B, 3, V] that K2b consumes
 *   mean_argmax   uint8 [B, V]      or NULL
 *   sample_argmax uint8 [B, N, V]   or NULL
 *   scores  double rows or NULL: per map (pe, ee, mi): {sum, sum over v>=thr, count v>=thr}
 *           = image_level_aggregation / threshold_aggregation numerators fused into the sweep
 *           (evaluation/uncertainty_aggregation/aggregate_uncertainties.py:34-37, 61-62); the row of
 *           (volume b, map k) starts at scores[(3 b + k) * score_stride]: score_stride = 3 is a packed
 *           [B, 3, 3] array, 7 writes columns 0..2 of a [B, 3, 7] score table in place.
 *   thresholds_host  3 doubles (pe, ee, mi) or NULL (then thr columns are 0).
 *   workspace: values_uncertainty_workspace_bytes(B, V, dtype) bytes when scores != NULL.
 */
size_t values_uncertainty_workspace_bytes(int64_t B, int64_t V, int dtype);
int values_uncertainty_fused(const void* probs, int dtype, int64_t B, int64_t N, int64_t C,
                             int64_t V, int64_t stride_b, int64_t stride_n, int64_t stride_c,
                             float* pe, float* ee, float* mi, int64_t map_stride_b,
                             uint8_t* mean_argmax, uint8_t* sample_argmax, double* scores,
                             int64_t score_stride, const double* thresholds_host, void* workspace,
                             size_t workspace_bytes, int variant, int tiles_per_cta, void* stream);
/* `variant` and `tiles_per_cta` choose the kernel for THIS call (no process-wide state); 0 / 0 is the
 * automatic choice.  The alternatives exist so that tests and benchmarks can run two implementations
 * of the same arithmetic against each other -- every variant gives bit-identical outputs:
 * 1 = register-stream kernel instead of the bulk-copy (TMA) ring, 2 = ring with 4-row stages where
 * 8-row stages are the default, 3 = 8-row stages x 3 at two CTAs per SM, 4 = (fp64 stacks) sample-outer
 * kernel instead of the class-outer ring; tiles_per_cta 1..64 fixes the voxel tiles a CTA walks. */

/* Replaces calculate_one_minus_msr(softmax_pred)  test_3D.py:521-525 and
 * ExperimentDataloader.get_max_softmax_pred  evaluation/experiment_dataloader.py:38-49.
 *   probs [B, C, V] (strides stride_b, stride_c, 1) -> out [B, V] of the same dtype = 1 - max_c p. */
int values_one_minus_msr(const void* probs, int dtype, int64_t B, int64_t C, int64_t V,
                         int64_t stride_b, int64_t stride_c, void* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * K2a: image-level and threshold aggregation of M stored maps (f32 or f64).
 * Replaces image_level_aggregation (aggregate_uncertainties.py:34-37) and
 * threshold_aggregation (:40-67; mask is `>=`).  Deterministic block-then-grid reduction.
 *   maps [M, V] (stride_m, 1); thresholds_host[n_thresholds] (<= 16), map m uses entry
 *   m % n_thresholds; NULL / 0 -> no threshold.
 *   out double [M, 3] = {sum, sum over v>=thr, count v>=thr}
 */
size_t values_map_reduce_workspace_bytes(int64_t M, int64_t V);
int values_map_reduce(const void* maps, int dtype, int64_t M, int64_t V, int64_t stride_m,
                      const double* thresholds_host, int n_thresholds, double* out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* K2b: patch-level aggregation.
 * Replaces patch_level_aggregation(image, patch_size, mean)  aggregate_uncertainties.py:13-31:
 * fp64 box-sum over every fully-inside window (scipy convolve mode="valid" with a ones
 * kernel), max_score = max, bbox_lo = FIRST C-order window index with
 * |sum - max| <= atol + rtol*|max| (np.isclose defaults rtol 1e-5, atol 1e-8).
 *   maps [M, shape[0], shape[1], shape[2]] (stride_m between maps; last axis contiguous);
 *   ndim in {1,2,3}: unused leading axes must have shape 1 and patch 1.
 *   max_score double, map m at max_score[m * score_stride]; bbox_lo: the corner of map m at
 *   bbox_lo[m * bbox_stride + 0..2], as int64 (bbox_dtype VALUES_I64) or as double (VALUES_F64, for
 *   an fp64 score table written in place); leading unused axes report 0, (-1, -1, -1) on a NaN map.
 * Returns VALUES_ERR_INVALID_ARG if any patch > shape (the reference raises ValueError).
 */
size_t values_patch_max_workspace_bytes(int64_t M, const int64_t* shape3_host,
                                        const int64_t* patch3_host, int path);
int values_patch_max(const void* maps, int dtype, int64_t M, int64_t stride_m,
                     const int64_t* shape3_host, const int64_t* patch3_host, int mean_flag,
                     double rtol, double atol, double* max_score, int64_t score_stride,
                     void* bbox_lo, int bbox_dtype, int64_t bbox_stride,
                     void* workspace, size_t workspace_bytes, int path, void* stream);
/* `path` selects the implementation for THIS call (no process-wide state): 0 automatic -- for
 * 10x10 in-plane patches (every reference config) an fp32 filter pass streams the maps by TMA
 * (cp.async.bulk.tensor) and lists the few (tile, z sub-chunk) entries that can hold the answer,
 * the exact fp64 march kernel then walks only those; other patches use the fused tile kernel when
 * they fit its shared memory, else the generic tiled path.  5 = march kernel over every tile (no
 * filter), 4 = fused tile kernel, 2 = generic tiled path: all give the same scores and boxes
 * (tests/test_gpu_parity.py runs them against each other).  Other values: VALUES_ERR_INVALID_ARG. */

/* ---------------------------------------------------------------------------------------
 * K3: sliding-window stitch accumulator (atomic-free, output-stationary).
 * Replaces the `+=` slab updates of DataCarrier3D.concat_data
 * (uncertainty_modeling/data_carrier_3D.py:154-179): uniform weights, patches summed in
 * list order, count incremented once per covering patch.
 *   patches  [N, n_patches_total, C, p0, p1, p2]; element strides patch_stride_n (between
 *            samples) and patch_stride_p (between patches); [C, p0, p1, p2] contiguous.
 *   patch_index int32 [n_sel] (device) or NULL (= 0..n_sel-1): which patches to use
 *   crop_lo  int32 [n_sel, 3] (device): lower corner (x0, y0, z0) of each selected patch
 *   out_sum  [N, C, vol0, vol1, vol2] contiguous, dtype out_dtype (F64 parity / F32)
 *   out_count double [vol0, vol1, vol2] or NULL (written when non-NULL; `accumulate` applies)
 *   accumulate != 0: out += (read-modify-write of every voxel);  == 0: out = (uncovered -> 0)
 */
int values_stitch_accumulate(const void* patches, int patch_dtype, int64_t patch_stride_n,
                             int64_t patch_stride_p, const int32_t* patch_index,
                             const int32_t* crop_lo, int64_t n_sel, int64_t N, int64_t C,
                             const int64_t* patch3_host, const int64_t* vol3_host,
                             void* out_sum, int out_dtype, double* out_count, int accumulate,
                             int path, void* stream);
/* `path`: 0 automatic (output boxes fed by tensor-map copies when the rows are 16-byte aligned, else the
 * register-staged vector kernel, else the scalar kernel), 1 scalar kernel, 2 vector kernel; same results. */

/* The same accumulator with a per-patch importance map (BASELINE.json north_star: "Gaussian-
 * weighted sliding-window patch stitching"; the reference itself accumulates with uniform weights,
 * SURVEY.md D1, so this is an opt-in extension and weight == NULL is exactly
 * values_stitch_accumulate):  out_sum += weight * patch,  out_count += weight.
 *   weight double [p0, p1, p2] (device) or NULL. */
int values_stitch_accumulate_weighted(const void* patches, int patch_dtype, int64_t patch_stride_n,
                                      int64_t patch_stride_p, const int32_t* patch_index,
                                      const int32_t* crop_lo, const double* weight, int64_t n_sel,
                                      int64_t N, int64_t C, const int64_t* patch3_host,
                                      const int64_t* vol3_host, void* out_sum, int out_dtype,
                                      double* out_count, int accumulate, int path, void* stream);

/* The weighted accumulator for a SEPARABLE importance map (the Gaussian of sliding-window inference is one):
 *   weight[x][y][z] = fl(fl(wx[x] * wy[y]) * wz[z])   (two fp64 roundings, in this order)
 * with wx / wy / wz double [p0] / [p1] / [p2] on the device.  Results are bit-identical to
 * values_stitch_accumulate_weighted on the materialised map; the factors live in shared memory, so the
 * weights cost no memory traffic.  Only the box kernel takes them: returns VALUES_ERR_UNSUPPORTED (nothing
 * launched) when path != 0, a patch edge exceeds 128 or the rows are not 16-byte aligned -- the caller then
 * materialises the map and calls values_stitch_accumulate_weighted. */
int values_stitch_accumulate_separable(const void* patches, int patch_dtype, int64_t patch_stride_n,
                                       int64_t patch_stride_p, const int32_t* patch_index,
                                       const int32_t* crop_lo, const double* wx, const double* wy,
                                       const double* wz, int64_t n_sel, int64_t N, int64_t C,
                                       const int64_t* patch3_host, const int64_t* vol3_host, void* out_sum,
                                       int out_dtype, double* out_count, int accumulate, int path,
                                       void* stream);

/* Save-time normalisation (data_carrier_3D.py:215-217, 326-329):
 *   out[m, v] = (double) maps[m, v] / max(count[v], clip_min)   -> fp64 as written to NIfTI;
 *   clip_min = 1 is the reference's np.clip(count, 1, None); clip_min = 0 divides by the weight
 *   sum of a weighted stitch and leaves uncovered voxels (count == 0) unscaled. */
int values_normalize_maps(const void* maps, int dtype, int64_t M, int64_t V, int64_t stride_m,
                          const double* count, double clip_min, double* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * K4: whole-map statistics either side of the C2 -> C3 path.
 *
 * Replaces calculate_foreground_quantile_image(image)
 * evaluation/uncertainty_aggregation/find_threshold.py:11-13 (np.count_nonzero):
 *   *count += #{ i : data[i] != 0 }   (device counter, caller zeroes it; NaN counts, -0.0 does not)
 *   dtype: U8 / I32 / I64 / F32 / F64. */
int values_count_nonzero(const void* data, int dtype, int64_t n, unsigned long long* count,
                         void* stream);

/* Building blocks of the exact np.quantile(images, q) in calculate_threshold_image /
 * find_threshold (find_threshold.py:63-68, 90-96) as a most-significant-digit radix select.
 * Values map to order-preserving unsigned keys (32 bit for F32, 64 bit for F64; NaN -> all
 * ones = last, as np.sort; -0.0 == +0.0).
 *   hist[d] += #{ i : key_i >> (bits - prefix_bits) == prefix and
 *                     (key_i >> (bits - prefix_bits - digit_bits)) & (2^digit_bits - 1) == d }
 * hist is a device array of 2^digit_bits counters (digit_bits <= 11) that ACCUMULATES, so the
 * maps of a validation set (and, after an all-reduce, of every rank) add into one histogram;
 * the caller walks the digits from the top, keeping the prefix of the bucket that holds the
 * wanted rank.  values_min_key_above: *out = min(*out, min key_i > key) -- the next order
 * statistic when the interpolation partner is not a duplicate of the selected value. */
int values_radix_histogram(const void* data, int dtype, int64_t n, uint64_t prefix,
                           int prefix_bits, int digit_bits, unsigned long long* hist,
                           void* stream);
int values_min_key_above(const void* data, int dtype, int64_t n, uint64_t key,
                         unsigned long long* out, void* stream);
/* The same walk without a host round trip per digit: `state` is a device array of 5 counters
 * {prefix, prefix_bits, rank (0-based, among the elements that match the prefix), count in the chosen
 * bucket, count in the last bucket of the FIRST digit (fp32: the NaN count)}.
 * values_radix_histogram_dev = values_radix_histogram with prefix / prefix_bits read from state;
 * values_radix_select picks the bucket of `hist` that holds the rank, updates state (prefix_bits becomes
 * all ones if the rank is out of range) and zeroes hist for the next digit.  After the last digit
 * state[0] is the key of the wanted order statistic, rank_initial - state[2] the number of smaller
 * elements and state[3] the number of elements equal to it. */
int values_radix_histogram_dev(const void* data, int dtype, int64_t n, const unsigned long long* state,
                               int digit_bits, unsigned long long* hist, void* stream);
int values_radix_select(unsigned long long* hist, int digit_bits, unsigned long long* state, void* stream);
/* The same two sweeps over a SET of maps in one launch per 96 maps (find_threshold.py:31-40, 90-96 walk a
 * validation set of separate images; a launch per image and digit made the quantile launch-bound):
 * maps_host / counts_host are HOST arrays of n_maps device pointers / element counts (empty maps allowed).
 * values_radix_histogram_set: state != NULL reads the prefix from the device state (as _dev), else prefix /
 * prefix_bits are the arguments (as values_radix_histogram). */
int values_radix_histogram_set(const void* const* maps_host, const int64_t* counts_host, int64_t n_maps,
                               int dtype, uint64_t prefix, int prefix_bits,
                               const unsigned long long* state, int digit_bits,
                               unsigned long long* hist, void* stream);
int values_min_key_above_set(const void* const* maps_host, const int64_t* counts_host, int64_t n_maps,
                             int dtype, uint64_t key, unsigned long long* out, void* stream);

/* Replaces the reductions of compute_ncc(gt_unc_map, pred_unc_map) evaluation/metrics/ncc.py:9-25.
 *   a [M, V] (stride_a, 1), b [M, V] (stride_b, 1), F32 or F64 each; shift: device double [M, 2]
 *   (sa, sb per pair) or NULL (= 0).
 *   out double [M, 5] = { sum(a-sa), sum(b-sb), sum (a-sa)^2, sum (b-sb)^2, sum (a-sa)(b-sb) }
 * Two calls give the reference's two-pass result: shift = NULL -> means, shift = means ->
 * centred moments.  Deterministic block-then-grid reduction in fp64. */
size_t values_pair_moments_workspace_bytes(int64_t M, int64_t V);
int values_pair_moments(const void* a, int dtype_a, int64_t stride_a, const void* b, int dtype_b,
                        int64_t stride_b, int64_t M, int64_t V, const double* shift, double* out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Replaces the binning of calib_stats(correct, calib_confids) evaluation/metrics/ace.py:49-81:
 * binids = np.digitize(prob, edges) - 1 with the caller's n_bins + 1 increasing edges
 * (np.linspace(0, 1 + 1e-8, 21); n_bins must be 20 as the reference hard-codes).
 *   prob F32/F64 [n]; correct U8/I32/I64 [n] (non-zero = true)
 *   out double [3, n_bins + 1] = { bin_total, bin_sums (sum prob), bin_true } per slot;
 *   counts are exact.  Elements below edges[0] are dropped (the reference raises ValueError
 *   before binning; the Python mirror does too). */
size_t values_calib_bins_workspace_bytes(int64_t n);
int values_calib_bins(const void* prob, int dtype, const void* correct, int label_dtype, int64_t n,
                      const double* edges_host, int n_bins, double* out, void* workspace,
                      size_t workspace_bytes, void* stream);

/* The per-image body of calibration_error (ace.py:96-127) in one sweep: for every rater r and
 * voxel v,  correct = (ref_segs[r, v] == pred_seg[v]),  conf = 1 / (1 + exp(-unc[v] * a + b))
 * (platt_scale_confid, ace.py:42-46, evaluated in the map's dtype as numpy does), skipping
 * ref_segs[r, v] == ignore_value when has_ignore; binned as values_calib_bins.
 *   unc F32/F64 [V]; pred_seg [V], ref_segs [R, V] of label_dtype (U8/I32/I64)
 *   workspace: values_calib_bins_workspace_bytes(V). */
int values_calib_bins_fused(const void* unc, int dtype, const void* pred_seg, const void* ref_segs,
                            int label_dtype, int64_t V, int64_t R, double a, double b,
                            int has_ignore, int64_t ignore_value, const double* edges_host,
                            int n_bins, double* out, void* workspace, size_t workspace_bytes,
                            void* stream);

/* Pairwise confusion matrices of label maps: the integer statistics behind calculate_ged and the
 * torchmetrics `dice` calls around it (uncertainty_modeling/test_3D.py:284-358; per-sample arg-max
 * maps come from values_uncertainty_fused's sample_argmax output).
 *   labels_a [Na, V] (stride_a, 1), labels_b [Nb, V] (stride_b, 1), one label dtype (U8/I32/I64)
 *   out uint64 [Na, Nb, C, C] (device, ACCUMULATES, caller zeroes):
 *       out[ia, ib, a, b] += #{ v : A[ia, v] == a and B[ib, v] == b };  labels outside [0, C) are dropped.
 *   n_classes <= 32.  Exact (integer atomics, order-free). */
int values_confusion_counts(const void* labels_a, int64_t Na, int64_t stride_a, const void* labels_b,
                            int64_t Nb, int64_t stride_b, int label_dtype, int64_t V, int n_classes,
                            unsigned long long* out, void* stream);

/* The sums behind calculate_test_metrics (uncertainty_modeling/test_3D.py:250-281): SoftDiceLoss
 * (uncertainty_modeling/loss_modules.py:7-90, smooth 1e-5, background included) + torch.nn.NLLLoss of
 * log(mean softmax), per rater.  probs [C, V] (F32 / F64, class stride stride_c), labels [R, V] (U8 / I32 /
 * I64, rater stride stride_r), C <= 8.  With CT = C rounded up to 2, 4 or 8:
 *   out double [R, 3 CT + 1] (device, overwritten):
 *     out[r][c]          = sum_v probs[c][v] [labels[r][v] == c]     (intersect)
 *     out[r][CT + c]     = #{ v : labels[r][v] == c }
 *     out[r][2 CT + c]   = sum_v probs[c][v]
 *     out[r][3 CT]       = sum_v log(probs[labels[r][v]][v])          (NaN if a label is outside [0, C))
 * fp64, fixed summation order (deterministic). */
size_t values_seg_loss_workspace_bytes(int64_t R, int C, int64_t V);
int values_seg_loss_terms(const void* probs, int dtype, int64_t stride_c, const void* labels,
                          int label_dtype, int64_t stride_r, int64_t R, int C, int64_t V, double* out,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Axis-order conversion for the hand-off files: medpy.io.load / save
 * (data_carrier_3D.py:233-371, experiment_dataloader.py:38-49, aggregate_uncertainties.py:77-79)
 * present a NIfTI payload -- stored x-fastest, a C-order [Z][Y][X] array -- to Python as an array
 * indexed [x][y][z]; cv2's [H][W] image likewise becomes [W][H].
 *   in  [n0, n1, n2] C-order, out [n2, n1, n0] C-order: out[c][b][a] = in[a][b][c]
 *   elem_bytes 1, 2, 4 or 8 (pure byte movement, any dtype of that size); in != out. */
int values_reverse_axes(const void* in, void* out, int elem_bytes, int64_t n0, int64_t n1,
                        int64_t n2, void* stream);

/* Pure host function: the error bound of K2b's fp32 filter pass, |fp32 box sum - exact box sum| <=
 * coef * max |input| for 10x10 in-plane patches, z-chunks of zc output planes and p0 planes per window.
 * tests/test_filter_bound.py checks it against an emulation of the kernel's operation order. */
double values_patch_filter_err_coef(int zc, int p0);

#ifdef __cplusplus
}
#endif
#endif /* VALUES_B200_H */
