/*
 * dstr_b200.h - C-ABI of the B200-native destripe engine (libdstr_b200.so).
 *
 * Drop-in boundary for the per-plane streak-removal hot path of
 * AllenNeuralDynamics/aind-smartspim-destripe.  The reference is pure Python, so the
 * binding a maintainer adds is a ctypes stub; every entry point
 * below names the reference interface it replaces.  Plain pointers and sizes only; no
 * torch / numpy types.  All functions return 0 on success, a negative DSTR_E_* code for
 * an invalid argument, or a positive cudaError_t value; dstr_last_error() gives text.
 *
 * Threading: one dstr_ctx per host thread per GPU.  A context owns its device workspace
 * and CUDA streams; calls on one context are serialised by the caller, calls on
 * different contexts are independent (reference: one OS process per plane worker,
 * zarr_destriper.py:1151-1165).
 */
#ifndef DSTR_B200_H
#define DSTR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dstr_ctx dstr_ctx;

/* element types of caller buffers */
#define DSTR_U16 0
#define DSTR_F32 1

/* error codes (negative) */
#define DSTR_E_ARG -1       /* bad argument (NULL, non-positive size, bad dtype ...)      */
#define DSTR_E_SHAPE -2     /* plane shape unsupported by this context                    */
#define DSTR_E_STATE -3     /* call not valid in the current state (e.g. no flat/dark)    */
#define DSTR_E_UNSUPPORTED -4

/* mode of dstr_filter_chunk */
#define DSTR_MODE_LOGSPACE 0 /* every plane filtered with `no_cells` (filtering.py:139-224) */
#define DSTR_MODE_DISPATCH 1 /* per-plane cells / no_cells choice (filtering.py:459-467)    */

/* flags of dstr_filter_chunk */
#define DSTR_FLAG_SHADOW 1      /* dark/flat epilogue + clip + truncate (filtering.py:338-414) */
#define DSTR_FLAG_EXPM1 2       /* corrected inverse exp(y)-1 instead of the reference's exp(y)+1 */
#define DSTR_FLAG_STACK_OTSU 4  /* 3-D input semantics: one Otsu threshold per level for the whole
                                   chunk (filtering.py:182-183,210-213)                          */
#define DSTR_FLAG_NO_SYNC 8     /* enqueue and return without synchronising (device buffers, or pinned
                                   host buffers: both stay owned by the call until dstr_synchronize; the
                                   next chunk's H2D then overlaps this chunk's D2H) */

#define DSTR_FLAG_NOTCH_ONLY 16 /* no Otsu mask and no median in-painting: every cH coefficient goes through the
                                   notch (the sub-band filter of the dual-band mode) */

/* Filter parameters = the reference's config dict {wavelet:"db3", level, sigma, max_threshold}
 * (run_capsule.py:377-388).  level < 0 means "None" (maximum level). */
typedef struct dstr_params {
    float sigma;
    float max_threshold;
    int level;
} dstr_params;

/* debug stages for dstr_set_debug_stop / dstr_debug_fetch */
#define DSTR_STAGE_NONE 0
#define DSTR_STAGE_ANALYSIS 1 /* stop after all DWT analysis levels (cA_l, cH_l valid)        */
#define DSTR_STAGE_OTSU 2     /* + min/max, histograms, thresholds                           */
#define DSTR_STAGE_FILTER 3   /* + cH_l replaced by dH_l = cH'_l - cH_l                      */
#define DSTR_STAGE_SYNTH 4    /* + cA_l replaced by dA_l for l < L (all but the final level) */

/* what to fetch */
#define DSTR_FETCH_CA 0    /* float32 [Z][H_l][W_l] approximation (or dA_l after SYNTH)       */
#define DSTR_FETCH_CH 1    /* float32 [Z][H_l][W_l] horizontal detail (or dH_l after FILTER)  */
#define DSTR_FETCH_STATS 2 /* float32 [Z][8]: qmin, qmax, otsu_raw, threshold, use_cells, fg_mean,
                              bg_mean, otsu_bin (level-independent entries repeated)           */
#define DSTR_FETCH_HIST 3  /* uint32 [Z][256] histogram of cH_l^2                             */

/* ---- lifetime -------------------------------------------------------------------------
 * Replaces: worker-process start-up (zarr_destriper.py:1151-1165).  A context is sized for
 * chunks of at most `max_planes` planes of exactly H x W pixels. */
int dstr_create(int device, int max_planes, int H, int W, dstr_ctx** out);
int dstr_destroy(dstr_ctx* ctx);
const char* dstr_last_error(const dstr_ctx* ctx); /* ctx may be NULL: last global error */

/* ---- shadow-correction fields ----------------------------------------------------------
 * Replaces: shadow_correction["flatfield"/"darkfield"] handed to flatfield_correction
 * (filtering.py:470-489, :338-414).  Host pointers, H*W elements each, already cropped to
 * the plane shape (filtering.py:377); NULL clears.  Copied to the device. */
int dstr_set_flat_dark(dstr_ctx* ctx, const float* flat, const float* dark);

/* ---- the hot path -----------------------------------------------------------------------
 * Replaces: the plane loop of execute_worker (zarr_destriper.py:319-327) calling
 * filter_stripes (filtering.py:417-491) / log_space_fft_filtering (filtering.py:139-224).
 *   in      : Z planes of H*W elements, contiguous, dtype in_dtype; host or device pointer
 *   out     : Z planes, dtype out_dtype (DSTR_U16: clip [0,65535] + truncate; DSTR_F32: raw)
 *   cells / no_cells : parameter sets; `cells` may be NULL in DSTR_MODE_LOGSPACE
 *   high_int: microscope_high_int of filter_stripes (2500 in the Zarr path)
 * Host buffers are streamed through the device in sub-chunks with H2D / compute / D2H
 * overlapped on three CUDA streams (replaces producer/consumer, zarr_destriper.py:797-906);
 * pinned host memory (dstr_host_alloc / dstr_host_register) makes the copies asynchronous. */
int dstr_filter_chunk(dstr_ctx* ctx, const void* in, int in_dtype, void* out, int out_dtype, int Z,
                      const dstr_params* cells, const dstr_params* no_cells, float high_int,
                      int mode, int flags);

/* Replaces: get_foreground_background_mean (filtering.py:54-88) for Z planes at once.
 * fg_mean / bg_mean: host arrays of Z doubles (0.0 when the class is empty); use_cells
 * (nullable): the filter_stripes decision fg > bg && fg > high_int (filtering.py:462);
 * threshold_mask: the reference's keyword (0.3 in the hot path). */
int dstr_plane_stats(dstr_ctx* ctx, const void* in, int in_dtype, int Z, double* fg_mean,
                     double* bg_mean, int* use_cells, float high_int, float threshold_mask);

/* Replaces: flatfield_correction (filtering.py:338-414) as a standalone call.  Host pointers;
 * img / flat / dark hold n_outer * H * W float32 elements (identical shapes, as the reference
 * requires, filtering.py:379-391); baseline is n_outer floats or NULL (zeros);
 * out = uint16( trunc( clip( (img <= dark ? 0 : img - dark) / flat - baseline, 0, 65535 ) ) ). */
int dstr_flatfield_correction(int device, const float* img, const float* flat, const float* dark,
                              const float* baseline, uint16_t* out, int n_outer, int H, int W);

/* ---- next row f3: multiscale pyramid ---------------------------------------------------------
 * Replaces: compute_pyramid / xarray_multiscale.windowed_mean with preserve_dtype, scale (2,2,2)
 * (zarr_destriper.py:365-407, :741-749): out = uint16(floor(mean of each 2x2x2 block)); odd
 * trailing planes / rows / columns are cropped.  Host or device pointers. */
int dstr_downscale2x(dstr_ctx* ctx, const uint16_t* in, int Z, int H, int W, uint16_t* out);
/* Fused variant: while a destriped uint16 batch is still resident, the following
 * dstr_filter_chunk calls also emit pyramid level 1 ((Z/2, H/2, W/2)) and, if non-NULL, level 2
 * ((Z/4, H/4, W/4), the windowed mean of level 1, like the reference which re-reads the written
 * level).  Z of the chunk should be a multiple of 4.  NULL level1 disables. */
int dstr_set_pyramid_outputs(dstr_ctx* ctx, void* level1, void* level2);

/* ---- geometry / tables (host only, no GPU work) ---------------------------------------------
 * Replaces: pywt.dwtn_max_level / dwt_coeff_len bookkeeping behind pywt.wavedec2
 * (filtering.py:176). */
int dstr_max_level(int H, int W);
int dstr_level_shape(int H, int W, int level, int* H_l, int* W_l);
/* Smallest float16 value v with sigmoid((v - 400) / 20) > threshold_mask under float16
 * arithmetic, i.e. the foreground rule of get_foreground_background_mean (filtering.py:78-81)
 * as a threshold on float16(pixel); -inf / +inf when the rule holds everywhere / nowhere. */
float dstr_foreground_threshold(float threshold_mask);
/* The same rule as a threshold on the float32 pixel value (what the kernels compare against):
 * float16 rounding is monotone, so float16(v) >= h  <=>  v >= t for the smallest float32 t whose
 * float16 rounding is >= h.  -inf when the rule holds everywhere, NaN when it holds nowhere. */
float dstr_foreground_threshold_f32(float threshold_mask);
/* Time-domain form of the reference's packed-rfft notch (filtering.py:206-215):
 * irfft(rfft(x) * g) = x - B x with B[t][v] = hp[(t - v) mod n] + hq[(t + v) mod n].
 * Writes n doubles to each of hp, hq. */
int dstr_notch_kernels(int n, double s, double* hp, double* hq);

/* The device evaluates B x as A x_e + Bo x_o (even / odd parts of the row): a compact FIR for
 * the smooth part of each plus a rank-J cosine correction for the kink of the packed layout;
 * eps is the truncation tolerance relative to the operator gain (0 = dense, exact kernels).
 * dstr_notch_design reports {ntap_e, ue_lo, ntap_o, uo_lo, J, Jpad}; dstr_notch_apply_host
 * evaluates y = B x on the host from exactly the float32 tables the device uses. */
int dstr_notch_design(int n, double s, double eps, int* info /*[6]*/);
int dstr_notch_apply_host(int n, double s, double eps, const double* x, double* y);
int dstr_set_notch_tolerance(dstr_ctx* ctx, double eps);

/* ---- pinned host memory ---------------------------------------------------------------------- */
int dstr_host_alloc(void** ptr, uint64_t bytes);
int dstr_host_free(void* ptr);
int dstr_host_register(void* ptr, uint64_t bytes);
int dstr_host_unregister(void* ptr);

/* ---- device memory (for callers that keep chunks resident, e.g. the benchmark) ------------- */
int dstr_device_alloc(dstr_ctx* ctx, void** ptr, uint64_t bytes);
int dstr_device_free(dstr_ctx* ctx, void* ptr);
int dstr_memcpy_h2d(dstr_ctx* ctx, void* dst, const void* src, uint64_t bytes);
int dstr_memcpy_d2h(dstr_ctx* ctx, void* dst, const void* src, uint64_t bytes);
int dstr_synchronize(dstr_ctx* ctx);
/* cudaStream_t of the compute stream (so a caller can record its own events on it) */
void* dstr_compute_stream(dstr_ctx* ctx);

/* ---- instrumentation ------------------------------------------------------------------------ */
#define DSTR_NUM_TIMERS 9
/* timers: 0 analysis L1, 1 analysis L2+, 2 histogram, 3 otsu, 4 row filter (all levels), 5 synthesis
 *         L2+, 6 final synthesis+epilogue, 7 whole chunk, 8 row filter level 1 alone (the single
 *         dominant kernel launch).  Device time in ms, accumulated; profiling mode issues the pass
 *         stage by stage on one stream. */
int dstr_set_profiling(dstr_ctx* ctx, int enabled);
int dstr_get_timers(dstr_ctx* ctx, double* ms_out /*[DSTR_NUM_TIMERS]*/, uint64_t* launches_out);
int dstr_reset_timers(dstr_ctx* ctx);
int dstr_set_debug_stop(dstr_ctx* ctx, int stage);
/* Copies an intermediate of the LAST sub-chunk processed (host chunks are cut into dstr_set_subchunk planes, default
 * 4; device chunks into max_planes).  Z of the layouts above is that sub-chunk's plane count and host_bytes of a
 * coefficient fetch must match it exactly (DSTR_E_ARG otherwise). */
int dstr_debug_fetch(dstr_ctx* ctx, int what, int level, void* host_buf, uint64_t host_bytes);
/* 1 (default): the per-level histogram / Otsu / row-filter branches run on side streams next to
 * the analysis and synthesis chains; 0: every kernel on the compute stream in stage order */
int dstr_set_overlap(dstr_ctx* ctx, int enabled);
/* 1 (default): level-1 analysis through the TMA-staged kernel (cp.async.bulk ring + mbarrier) when
 * W >= 256 and rows are 16-byte multiples; 0: always the register-streaming kernel */
int dstr_set_tma(dstr_ctx* ctx, int enabled);

/* Classic dual-band mode (pystripe `filter_streaks`; the picture under "Dual-band" in the reference README; not in
 * the reference sources, see SURVEY.md Appendix B — parity is against oracle/dual_band.py only):
 *   bg = min(img, T), fg = max(img, T); each band through log1p -> wavedec2(db3, level) -> packed-rfft notch on
 *   every cH_l with s_l = H_l sigma / H (no mask, no median) -> waverec2 -> exp(y) - 1;
 *   f = sigmoid((img - T) / crossover) (filtering.py:13-51);  out = fg_f f + bg_f (1 - f);  then `- dark` when
 *   dark > 0, `/ flat` when flat != NULL, clip to [0, 65535], truncate to uint16.
 * sigma_fg == sigma_bg: one band, no blend; a sigma of 0 leaves that band unfiltered (the image itself); both 0:
 * only dark / flat / clip.  in / out: host or device pointers ([Z][H][W]); thresholds: host, one per plane;
 * flat: host [H][W] or NULL; level < 0 = maximum level.  Needs 3 float32 planes of workspace per max_planes. */
int dstr_dual_band_chunk(dstr_ctx* ctx, const void* in, int in_dtype, uint16_t* out, int Z, float sigma_fg,
                         float sigma_bg, int level, const float* thresholds, float crossover, float dark,
                         const float* flat);
/* Exact histogram of every uint16 plane, hist[Z][65536] (host).  The host derives the dual-band threshold from it
 * (skimage threshold_otsu on an integer image counts every value). */
int dstr_histogram_u16(dstr_ctx* ctx, const uint16_t* in, int Z, uint32_t* hist);
/* 1: the row filter (filtering.py:195-217) of every band 96 <= W_l <= 1056 runs on the 5th-generation
 * tensor cores (tcgen05.mma kind::f16 on fp16 hi/lo operand pairs, accumulators in TMEM, operands
 * staged by TMA bulk copies; csrc/dstr_notch_umma.cuh); 0 (default, or environment DSTR_UMMA=1 to flip
 * it): the mma.sync kernel below, which measures 40 % faster on B200 (1.14 vs 1.93 ms for the level-1 launch of a
 * 128 x 2048 x 2048 chunk; DESIGN.md section 5b explains why).  Both paths pass the same parity tests. */
int dstr_set_umma(dstr_ctx* ctx, int enabled);
/* Row filter variant when dstr_set_umma is off: 1 (default; environment DSTR_ROW_FILTER) = the even / odd FIRs and the
 * rank-J correction as mma.sync m16n8k16 products on fp16 hi/lo operand pairs, 8 rows per block
 * (csrc/dstr_rows_mma.cuh), 2 = the same kernel with 4 rows per block (what bands too long for the 8-row form fall back
 * to), 0 = the register-tiled FMA kernel of round 1.  Same operator design, same parity tests. */
int dstr_set_row_filter(dstr_ctx* ctx, int kind);
/* geometry of the tensor-core row filter for a band of width n:
 * info = {eligible, passes, outputs per pass, k chunks, table bytes, shared memory bytes, outputs, padded K}
 * (tables are sized for notch width s; s <= 0 reports the geometry only) */
int dstr_notch_umma_info(int n, double s, int* info /*[8]*/);
/* y = x - irfft(rfft(x) * notch(n, s)) (filtering.py:206-215) evaluated on the HOST through the data
 * path of the tensor-core kernel (fp16 hi/lo Hankel tables addressed like the UMMA descriptors, banded /
 * remainder product lists, power-of-two pre-scaling from `thr`), double accumulation; x has n entries
 * bounded by thr.  info = {band radius, three-product remainder flag, MMAs per 128-row item, compact band
 * tables}.  Test support: no GPU needed. */
int dstr_notch_umma_apply_host(int n, double s, double thr, const double* x, double* y, int* info /*[4]*/);
/* ---- Zarr chunk codec (next-row f1) ------------------------------------------------------------
 * Blosc1 frames as the reference writes them (numcodecs Blosc(cname="zstd", clevel=3, shuffle=SHUFFLE),
 * zarr_destriper.py:1066-1074): zstd (compressor 4) or lz4 (1) streams with optional byte shuffle, host
 * code over the system libzstd / liblz4 (dlopen).  compress: dst_capacity >= nbytes + 16; returns the frame
 * size.  decompress: returns the number of bytes written (dst == NULL: the size the frame expands to).
 * Negative return = DSTR_E_*.  Thread safe; no GPU involved. */
int dstr_blosc_available(int compressor);
int64_t dstr_blosc_compress(const void* src, uint64_t nbytes, int typesize, int clevel, int shuffle, int compressor,
                            uint64_t blocksize, void* dst, uint64_t dst_capacity);
int64_t dstr_blosc_decompress(const void* frame, uint64_t frame_bytes, void* dst, uint64_t dst_capacity);

/* PNG row reconstruction (filter types 0-4) for the TIFF / PNG front-end (reference readers.py:64-89 reads PNG
 * through imageio): scan = h rows of (1 filter byte + stride bytes) as inflated, out = h * stride bytes. */
int dstr_png_unfilter(const uint8_t* scan, int h, int stride, int bpp, uint8_t* out);
/* sub-chunk size (planes) used when streaming host buffers; 0 restores the default */
int dstr_set_subchunk(dstr_ctx* ctx, int planes);

#ifdef __cplusplus
}
#endif
#endif /* DSTR_B200_H */
