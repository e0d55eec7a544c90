/*
 * debvader_b200 — C-ABI of the B200-native debvader hot path.
 *
 * The reference (astrodeepnet/debvader) has no FFI: the path sits behind plain
 * Python functions that call TensorFlow/Keras/TFP/scipy.  This header is the
 * boundary a maintainer would bind (ctypes/cffi) directly
 * beneath those functions.  Each entry point cites the reference code it
 * replaces (paths relative to the reference's src/debvader/).
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types; every function returns
 *     0 on success or a negative dbv_status, never throws; the message of the
 *     last failure on the calling thread is dbv_last_error().
 *   - "dev" pointers are device memory on the ctx's device, "host" pointers are
 *     host memory (pinned memory makes the copies asynchronous and fast).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default
 *     stream).  All device-pointer calls only enqueue work on it; no hidden
 *     synchronisation.  The *_host calls are synchronous.
 *   - the caller owns every buffer; the ctx owns repacked weights + workspace.
 *   - one ctx per device, not re-entrant.
 *   - tensors are dense, row-major, NHWC: stamps (B,59,59,6), fields (1,F,F,C).
 */
#ifndef DEBVADER_B200_H
#define DEBVADER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBV_ABI_VERSION 6  /* 2: dbv_deblend_host gained mean_dev / stddev_dev; precisions FP16X3, MIXED
                              3: dbv_window_axpy_ex, dbv_spline_* (sub-pixel placement), dbv_shift_objective
                              4: dbv_window_axpy_rect (rectangular local regions, caller scratch, in-place),
                                 dbv_sqdiff_sum_rect, dbv_shift_objective_batch, dbv_fp16_overflow
                              5: DBV_PREC_FP32TC; the binning scratch of dbv_window_axpy_rect holds 16-byte records and must
                                 be 16-byte aligned (its size still comes from dbv_window_axpy_scratch_bytes)
                              6: dbv_detect* (device detector), dbv_layer_kernel */

typedef enum {
  DBV_OK = 0,
  DBV_ERR_INVALID = -1,     /* bad argument / shape / key */
  DBV_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed */
  DBV_ERR_STATE = -3,       /* e.g. compute before dbv_finalize_weights */
  DBV_ERR_UNSUPPORTED = -4  /* not the DC2 architecture, not an sm_100 device, ... */
} dbv_status;

/* arithmetic of the network path */
typedef enum {
  DBV_PREC_FP32 = 0,   /* fp32 SIMT kernels: <=1e-5 of peak flux vs the fp64 oracle        */
  DBV_PREC_BF16 = 1,   /* tcgen05 bf16 x bf16 -> fp32, activations stored once in bf16      */
  DBV_PREC_BF16X3 = 2, /* tcgen05, hi/lo bf16 split of activations and weights (3 MMAs per
                          product, ~16-bit mantissa): <=1e-3 of peak flux                   */
  DBV_PREC_FP16X3 = 3, /* tcgen05, hi/lo fp16 split, same cost as BF16X3; activations saturate
                          at +-65504 (fp16 range)                                           */
  DBV_PREC_MIXED = 4,  /* BF16X3, except that the activations entering the four large-image
                          decoder layers (convT6, convT7, convT8, head: 55 % of the step) are
                          stored ONCE in fp16 (11-bit mantissa, saturating at +-65504) and
                          multiplied by fp16 hi/lo weights: 2 MMAs per product instead of 3
                          and half the activation traffic there; ~5e-4 of peak flux         */
  DBV_PREC_FP32TC = 5  /* the <=1e-5 tier on the tensor cores: FP16X3 operands (~22-bit mantissa)
                          with the tcgen05 accumulation chain cut every <=128 values of K and the
                          partial sums promoted into fp32 registers with round-to-nearest adds
                          (tcgen05 accumulates without rounding to nearest: ~1 ulp lost per MMA) */
} dbv_precision;

typedef enum { DBV_F32 = 0, DBV_F64 = 1 } dbv_dtype;

typedef struct dbv_ctx dbv_ctx;

/* ---- housekeeping ------------------------------------------------------------------- */
int dbv_abi_version(void);
const char* dbv_last_error(void);

/* Replaces create_model_vae / load_deblender's graph construction (model/model.py:164-218,
 * 221-259): allocates the workspace for `chunk` stamps per pass (0 = default). */
int dbv_create(dbv_ctx** out, int device, int precision, int64_t chunk);
int dbv_destroy(dbv_ctx* ctx);

/* Replaces net.load_weights (model/model.py:262-266).  `key` is the TF2 checkpoint key
 * without the "/.ATTRIBUTES/VARIABLE_VALUE" suffix, e.g.
 * "layer_with_weights-0/layer_with_weights-1/kernel"; data is host fp32 in the
 * checkpoint's own layout (Conv2D HWIO, Conv2DTranspose HWOI, Dense IO, PReLU HWC). */
int dbv_set_weights(dbv_ctx* ctx, const char* key, const float* host, const int64_t* shape, int ndim);
/* Validates the 64 tensors against the DC2 architecture, repacks and uploads them. */
int dbv_finalize_weights(dbv_ctx* ctx);

/* ---- network: device buffers ------------------------------------------------------------ */
/* encoder(x): model/model.py:61-100.  x (B,59,59,6) f32 -> params (B,560) f32. */
int dbv_encode(dbv_ctx* ctx, const float* x_dev, int64_t B, float* params_dev, void* stream);

/* tfp.layers.MultivariateNormalTriL(32) / MvNormal.__call__: model/model.py:43-58, 211-214.
 * z = loc + L eps.  eps_dev (B,32) or NULL; with NULL, `sample`!=0 draws eps ~ N(0,1) from a
 * Philox stream keyed by (seed, first_stamp + row) and `sample`==0 gives z = loc.
 * Optional outputs: loc (B,32), stddev (B,32) = sqrt(sum_j L_ij^2). */
int dbv_latent(dbv_ctx* ctx, const float* params_dev, const float* eps_dev, uint64_t seed, int sample,
               int64_t first_stamp, int64_t B, float* z_dev, float* loc_dev, float* stddev_dev, void* stream);

/* decoder(z): model/model.py:103-161.  z (B,32) -> mean, stddev (B,59,59,6) f32
 * (Normal(loc=t[...,:6], scale=1e-4+t[...,6:]) of model/model.py:154-159). */
int dbv_decode(dbv_ctx* ctx, const float* z_dev, int64_t B, float* mean_dev, float* stddev_dev, void* stream);

/* net(x) as deblend() uses it: deblend_cutout/deblender.py:18,24.  The benchmarked unit.
 * stddev_dev / z_dev may be NULL. */
int dbv_deblend(dbv_ctx* ctx, const float* x_dev, int64_t B, const float* eps_dev, uint64_t seed, int sample,
                float* mean_dev, float* stddev_dev, float* z_dev, void* stream);

/* ---- network: host buffers (the end-to-end call of deblend()) ----------------------------- */
/* deblend_cutout/deblender.py:6-24 including tf.cast(images, tf.float32) (x_dtype = DBV_F32 or
 * DBV_F64; the cast is done on the device) and outimg.mean().numpy().  Chunks are pipelined:
 * H2D of chunk k+1, compute of chunk k and D2H of chunk k-1 overlap on three streams.
 * stddev_host / z_host / eps_host may be NULL.  mean_dev / stddev_dev (device, (B,59,59,6) fp32, may be
 * NULL) keep the outputs resident: the decoder then writes them directly, and a NULL stddev_host
 * skips that device-to-host copy — the distribution object deblend() returns fetches the stddev
 * only when a caller asks for it (the reference's callers mostly use the mean).  Synchronous. */
int dbv_deblend_host(dbv_ctx* ctx, const void* x_host, int x_dtype, int64_t B, const float* eps_host,
                     uint64_t seed, int sample, float* mean_host, float* stddev_host, float* z_host,
                     float* mean_dev, float* stddev_dev);

/* The piece sizes dbv_deblend_host uses for a batch of B stamps on a context of `chunk` stamps (pure host function, no
 * device needed): fills counts[0..min(n, max_pieces)) and returns n.  Small first piece, pieces growing as fast as the
 * copies keep up with the compute, one short last piece; DBV_HOST_PIECE=n selects fixed-size pieces instead. */
int64_t dbv_host_schedule(int64_t B, int64_t chunk, int64_t* counts_host, int64_t max_pieces);

/* ---- field operators: device buffers -------------------------------------------------------- */
/* extract_cutouts: extract/extraction.py:21-36.  For k in [0,N): copies the window of `field`
 * (1,F,F,C) whose first row/col is (sx[k], sy[k]) into out[slot[k]] (S,S,C).  flags[k] bit0 / bit1
 * = the source has length 1 along rows / cols and is broadcast (numpy assignment semantics).
 * The host planner (which applies int() truncation and slice.indices) owns acceptance.
 * sx, sy (int32), flags (uint8), slot (int64) are device arrays.  out is f64 (reference-faithful)
 * or f32 (fused tf.cast of deblender.py:18). */
int dbv_extract(const void* field_dev, int field_dtype, int64_t F, int C, const int32_t* sx_dev,
                const int32_t* sy_dev, const uint8_t* flags_dev, const int64_t* slot_dev, int64_t N, int S,
                void* out_dev, int out_dtype, void* stream);

/* get_residual_field / get_predicted_field with integer positions: deblend/field_deblender.py:46-97,
 * 99-189.  out = in + alpha * sum_k paste(stamps[k] at rows x0[k].., cols y0[k]..), clipped to the
 * field; in_dev == NULL means zeros; in_dev may equal out_dev.  Contributions are applied to each
 * pixel in ascending k with one rounding per addition (no atomics, no FMA contraction), so the
 * result is bit-identical to the sequential numpy loop.  stamps (N,S,S,C) f32. */
int dbv_window_axpy(const void* in_dev, void* out_dev, int field_dtype, int64_t F, int C, const float* stamps_dev,
                    const int32_t* x0_dev, const int32_t* y0_dev, int64_t N, int S, double alpha, void* stream);

/* The same with f32 or f64 stamps (stamp_dtype), interleaved (N,S,S,C) or planar (N,C,S,S; stamp_planar != 0): the
 * sub-pixel path pastes planar f64 windows. */
int dbv_window_axpy_ex(const void* in_dev, void* out_dev, int field_dtype, int64_t F, int C, const void* stamps_dev,
                       int stamp_dtype, int stamp_planar, const int32_t* x0_dev, const int32_t* y0_dev, int64_t N, int S,
                       double alpha, void* stream);

/* The same operator on a RECTANGULAR field (1,FH,FW,C) — a rank's local region (owner tile + 30-px halo) of a field
 * tiled across GPUs, positions relative to the region's origin (windows are clipped to the region) — with the binning
 * scratch supplied by the caller (dbv_window_axpy_scratch_bytes(FH, FW) bytes of device memory, used only on `stream`:
 * calls on different streams never share state; NULL = no binning, every tile scans all stamps).
 * in_dev == out_dev is the IN-PLACE form for a device-resident iterative loop: only the elements a stamp covers are
 * read and written back (the algorithmic traffic of SURVEY section 8d: window read-modify-write + stamp), tiles no stamp
 * touches return at once.  Same ordering guarantee: bit-identical to the sequential loop restricted to the region. */
int64_t dbv_window_axpy_scratch_bytes(int64_t FH, int64_t FW);
int dbv_window_axpy_rect(const void* in_dev, void* out_dev, int field_dtype, int64_t FH, int64_t FW, int C,
                         const void* stamps_dev, int stamp_dtype, int stamp_planar, const int32_t* x0_dev,
                         const int32_t* y0_dev, int64_t N, int S, double alpha, void* scratch_dev, int64_t scratch_bytes,
                         void* stream);

/* Sub-pixel placement: scipy.ndimage.shift(padded_canvas, shift=(x_pos, y_pos)) of
 * deblend/field_deblender.py:66-95, 121-182 and deblend_cutout/optimization.py:27-29, 41-44 (order 3,
 * mode 'constant', prefilter=True; scipy==1.11.2, requirements.txt:7), evaluated on the window of the
 * canvas where the result is not negligible: per axis the segment [origin-P, origin+S+P) clipped to
 * the canvas is prefiltered (the prefilter's response decays as 0.268^d, P = 28 truncates at 1e-16 of
 * the data's peak) with scipy's mirror initialisation wherever the segment ends on the canvas edge, so
 * a canvas smaller than the segment is filtered whole, exactly as scipy does.
 *   data (N,S,S,C) f32/f64: the block of non-zero canvas samples of each item; its sample 0 sits at
 *   canvas row origin_x[k] / col origin_y[k] (both NULL: `origin` for all, i.e. int((F-S)/2) of
 *   field_deblender.py:72).  pos = the shift.  Output window k has side E = dbv_spline_extent(S,P)
 *   = S+2P+2 and starts at canvas row ax[k], col ay[k] (the caller passes origin - P - 1 + floor(pos));
 *   placed (N,C,E,E) f64 — PLANAR — is then pasted with dbv_window_axpy_ex(..., DBV_F64, 1, ax, ay, N, E, ...).
 *   scratch: dbv_spline_scratch_doubles(N,S,C,P) doubles (pass-X output + interpolation weights).  S + 2P <= 192. */
int dbv_spline_extent(int S, int P);
int64_t dbv_spline_scratch_doubles(int64_t N, int S, int C, int P);
int dbv_spline_place(const void* data_dev, int data_dtype, int64_t N, int S, int C, int64_t F, int origin,
                     const int32_t* origin_x_dev, const int32_t* origin_y_dev, const double* pos_x_dev,
                     const double* pos_y_dev, const int32_t* ax_dev, const int32_t* ay_dev, int P, double* scratch_dev,
                     double* placed_dev, void* stream);

/* Objective of position_optimization (deblend_cutout/optimization.py:21-33):
 * out[0] = mean over the F x F canvas of (field[..., band] - shifted)^2 where `shifted` is zero but for
 * the placed window (E,E) f64 at (ax, ay): (sumsq_field + sum_window(T^2 - 2 field T)) / F^2, with
 * sumsq_field = sum field[..., band]^2 from dbv_band_sumsq (scratch as for dbv_mse). */
int dbv_band_sumsq(const double* field_dev, int64_t F, int C, int band, double* out_dev, void* scratch_dev,
                   int64_t scratch_bytes, void* stream);
int dbv_shift_objective(const double* field_dev, int64_t F, int C, int band, const double* placed_dev, int E, int ax,
                        int ay, double sumsq_field, double* out_dev, void* stream);
/* The objective for a BATCH of placed windows (the batched position fit evaluates every galaxy of a field, at all the trial
 * shifts of one optimiser iteration, with one dbv_spline_place + one call of this): placed (M,E,E) f64, window i at
 * (ax[i], ay[i]) (device int32 arrays); out[i] = (sumsq_field + sum_window_i(T^2 - 2 field T)) / F^2. */
int dbv_shift_objective_batch(const double* field_dev, int64_t F, int C, int band, const double* placed_dev, int E,
                              const int32_t* ax_dev, const int32_t* ay_dev, int64_t M, double sumsq_field, double* out_dev,
                              void* stream);
/* One evaluation of `fun(x)` of optimization.py:21-33 for the already placed prediction placed1 (E1,E1) f64
 * at (a1x, a1y) (= shift(r_band_prediction, galaxy_distance_to_center), optimization.py:41-44): a second
 * placement by x = (x0, x1) into placed2 (E2 = dbv_spline_extent(E1,P) squared; scratch dbv_spline_scratch_doubles(1,E1,1,P) doubles), then
 * dbv_shift_objective.  Synchronous: the value is written to out_host (and out_dev). */
int dbv_position_objective(const double* field_dev, int64_t F, int C, int band, const double* placed1_dev, int E1, int a1x,
                           int a1y, double x0, double x1, int P, double sumsq_field, double* scratch_dev,
                           double* placed2_dev, double* out_dev, double* out_host, void* stream);

/* mse of the centre window: deblend/field_deblender.py:323-332.  out[k] = mean over
 * [lo,hi)x[lo,hi)xC of (cutouts[k] - mean[k])^2 in fp64. */
int dbv_center_mse(const void* cutouts_dev, int cutouts_dtype, const float* mean_dev, int64_t N, int S, int C,
                   int lo, int hi, double* out_dev, void* stream);

/* training/metrics.py:4-12 on two fields (deblend_iterative/iterative_deblender.py:52,75):
 * out[0] = mean((a-b)^2) over n elements, fp64, fixed reduction order. */
int dbv_mse(const void* a_dev, const void* b_dev, int dtype, int64_t n, double* out_dev, void* scratch_dev,
            int64_t scratch_bytes, void* stream);
int64_t dbv_mse_scratch_bytes(void);
/* out[0] = sum((a-b)^2) over a rows x cols sub-rectangle of two row-pitched arrays (pitches in elements): the partial
 * sum of a rank's OWNER TILE inside its local region; a tiled field's mse (iterative_deblender.py:52,75) is the
 * all-reduced sum of these divided by the element count.  scratch as for dbv_mse. */
int dbv_sqdiff_sum_rect(const void* a_dev, const void* b_dev, int dtype, int64_t rows, int64_t cols, int64_t pitch_a,
                        int64_t pitch_b, double* out_dev, void* scratch_dev, int64_t scratch_bytes, void* stream);

/* DBV_PREC_MIXED stores the activations entering convT6, convT7, convT8 and the head in fp16 (range +-65504).  The
 * epilogues that write them watch for saturation and set a sticky host-mapped flag: returns 1 if any activation of a
 * call completed so far left the fp16 range (its result is then NOT within the 1e-3 tolerance: use DBV_PREC_BF16X3,
 * which has the fp32 range), 0 otherwise; reset != 0 clears the flag.  Valid after the caller has synchronised the
 * stream the call was enqueued on.  Always 0 for the other precisions. */
int dbv_fp16_overflow(dbv_ctx* ctx, int reset);

/* ---- detection ---------------------------------------------------------------------
 * Replaces the two calls reference detect/detection.py:5-56 makes into the CPU library `sep`
 * (sep.Background(r_band) :15 and sep.extract(r_band - bkg, thresh=1.5, err=bkg.globalrms, minarea=4,
 * filter_kernel=<7x7>, filter_type="conv") :37-46) and the centre arithmetic of :48-54, on a field that stays on the
 * device.  Restated from the published SExtractor algorithm (oracle/detect_numpy.py, bit-exact with it): mesh
 * background, matched filter, threshold, 8-connected components of >= minarea pixels in the order Lutz's scan
 * completes them, barycentres.  NOT restated: multi-threshold deblending and the `clean` pass.  Parity with sep
 * itself is unpinned (sep is not installable where this was built).
 *   field   (H, pitch, C) f64 or f32 on the device; band = the plane detection runs on (2 = r)
 *   taps    HOST pointer, kh*kw float32, already divided by the sum of their absolute values
 *   cy, cx  subtracted from the barycentres before rounding (the reference: int(F/2))
 *   outputs (device): n_found[0] = objects found (may exceed max_objects: only the first max_objects are written),
 *           xy (max_objects,2) f64 = (x, y) barycentres in pixels, centres (max_objects,2) f64 = (row, col) offsets
 *           rounded half-to-even, npix (max_objects) i32, stats[0..2] = global background, global rms, threshold
 *   scratch 256-byte aligned device memory of dbv_detect_scratch_bytes(H, W, max_objects) bytes, private to the call
 * Everything is enqueued on `stream`; no host synchronisation. */
int64_t dbv_detect_scratch_bytes(int64_t H, int64_t W, int64_t max_objects);
int dbv_detect(const void* field, int dtype, int64_t H, int64_t W, int64_t pitch, int C, int band, const float* taps, int kh,
               int kw, double thresh_sigma, int minarea, int cy, int cx, int64_t max_objects, void* scratch,
               int64_t scratch_bytes, int32_t* n_found, double* xy, double* centres, int32_t* npix, float* stats, void* stream);
/* The same detection on a field TILED over GPUs (one rank = owner tile + halo = a (RH, RW) region at (gy0, gx0) of the (H, W)
 * field; reference: none — the reference is single-process).  Two phases around ONE exchange of the tiny mesh maps:
 *   dbv_detect_meshes   band plane of the region into `scratch`, statistics of every 64x64 mesh lying wholly inside the region
 *                       written at the mesh's place in the caller's field-sized (ny, nx) maps back0 / sig0 (other entries are
 *                       left alone: preset -inf, then max-reduce the maps over the ranks);
 *   dbv_detect_objects  everything else on the region, with the COMPLETE maps: the mesh post-processing (replicated), background,
 *                       filter (exact kh/2, kw/2 pixels inside the region's inner edges), components, and the objects whose LAST
 *                       pixel lies in the owner tile [oy0, oy1) x [ox0, ox1) (field coordinates), in ascending order of
 *                       last_out = raster index of that pixel in the whole field.  Merging the ranks' lists by last_out gives
 *                       the single-GPU list bit for bit.  flags[0] != 0: an owned object reaches the rim of what the region
 *                       sees — detect on the assembled field instead.
 * dbv_detect == both phases with the field as its own region and owner tile. */
int64_t dbv_detect_scratch_bytes_region(int64_t H, int64_t W, int64_t RH, int64_t RW, int64_t max_objects);
int dbv_detect_meshes(const void* region, int dtype, int64_t RH, int64_t RW, int64_t pitch, int C, int band, int64_t gy0, int64_t gx0,
                      int64_t H, int64_t W, int64_t max_objects, void* scratch, int64_t scratch_bytes, float* back0, float* sig0,
                      void* stream);
int dbv_detect_objects(int64_t RH, int64_t RW, int64_t gy0, int64_t gx0, int64_t H, int64_t W, const float* back0, const float* sig0,
                       const float* taps, int kh, int kw, double thresh_sigma, int minarea, int cy, int cx, int64_t oy0, int64_t oy1,
                       int64_t ox0, int64_t ox1, int64_t max_objects, void* scratch, int64_t scratch_bytes, int32_t* n_found,
                       double* xy, double* centres, int32_t* npix, int64_t* last_out, int32_t* flags, float* stats, void* stream);
/* device pointer to an intermediate plane of the last call on `scratch` (per-stage parity tests): 0 foreground f32, 1 filtered
 * f32, 2 labels i32 (region-sized), 3 / 4 filtered mesh background / sigma f32 (ny,nx), 5 / 6 raw ones (dbv_detect only) */
const void* dbv_detect_plane(void* scratch, int64_t H, int64_t W, int64_t max_objects, int what);
const void* dbv_detect_plane_region(void* scratch, int64_t H, int64_t W, int64_t RH, int64_t RW, int64_t max_objects, int what);

/* ---- introspection -------------------------------------------------------------------------- */
/* number of kernels this library has launched on behalf of ctx (bench.py's gpu_launches) */
int64_t dbv_launch_count(const dbv_ctx* ctx);
int64_t dbv_global_launch_count(void);
/* name / elapsed ms per call of the per-layer CUDA-event timers, averaged over the dbv_deblend calls made since
 * profiling was enabled (the caller synchronises first); returns the number of layers. */
int dbv_set_profiling(dbv_ctx* ctx, int enabled);
int dbv_layer_times(dbv_ctx* ctx, int max_layers, float* ms_out, char* names_out /* max_layers*32 bytes */);
/* the __global__ function (named as the ncu launch list names it, leading template arguments only) that runs `layer`
 * under the ctx's precision and tuned plan; out_bytes >= 48.  bench.py groups the per-layer times by it. */
int dbv_layer_kernel(dbv_ctx* ctx, const char* layer, char* out, int out_bytes);
/* copy an internal activation buffer (debug / per-layer parity): writes fp32 NHWC */
int dbv_debug_activation(dbv_ctx* ctx, const char* name, int64_t B, float* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEBVADER_B200_H */
