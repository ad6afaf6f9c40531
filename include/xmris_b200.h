/* xmris_b200 -- C ABI of the B200 FID->spectrum hot path (libxmris_b200.so).
 *
 * The reference (andrewendlinger/xmris v0.6.1) is pure Python and has no FFI of its own: its plugin boundary is
 * the xarray accessor `.xmr` (src/xmris/core/accessor.py:707-710) whose methods forward to the functions in
 * src/xmris/processing/{fid,fourier,phasing}.py.  Each entry point below names the reference function it
 * replaces; `xmris_b200/processing.py` is the Python side that binds them with ctypes.
 *
 * Conventions
 *   - complex64 data, interleaved (re, im) float pairs, batch-major: element (b, k) at base[(b*n + k)*2].
 *     The transform axis is last and contiguous.  Rows must be 8-byte aligned; the TMA (bulk-copy) fast path is
 *     taken when the base pointer is 16-byte aligned and n_in is even, otherwise a plain-load path is used.
 *   - `*_dev` pointers are caller-owned DEVICE memory on the current CUDA device; `*_host` are host memory.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream) unless it
 *     says otherwise.  No hidden allocation except a per-(device, N) twiddle-table cache guarded by a mutex.
 *   - return value: 0 on success, else one of XMR_ERR_*; xmr_last_error() returns the thread-local message.
 *   - supported transform lengths n_out: powers of two in [16, 8192] (one spectrum per CTA-resident slot); the Python
 *     layer composes other lengths (<= 4096) from these entry points as a chirp-z transform.
 */
#ifndef XMRIS_B200_H
#define XMRIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XMR_OK 0
#define XMR_ERR_BAD_ARG 1        /* NULL pointer, negative size, n_in > n_out, bad pad_left ...            */
#define XMR_ERR_UNSUPPORTED_N 2  /* n_out is not a power of two in [16, 8192]                              */
#define XMR_ERR_CUDA 3           /* a CUDA runtime call failed; message holds cudaGetErrorString           */
#define XMR_ERR_NO_CONVERGE 4    /* reserved                                                               */

/* window_mode */
#define XMR_WIN_NONE 0      /* multiply by `scale` only                                                    */
#define XMR_WIN_TABLE 1     /* window_dev[k], k < n_out  (already includes any 1/sqrt(N))                  */
#define XMR_WIN_SEPARABLE 2 /* w[M*n1 + n2] = window_dev[n2] * win_rows_host[n1], M = min(n_out, 256)      */

/* phase_mode of the fused epilogue */
#define XMR_PHASE_NONE 0
#define XMR_PHASE_UNIFORM 1 /* out[b,m] *= exp(2*pi*i*(ph_a_turns + ph_b_turns*m)), same for every spectrum */

/* autophase objective (src/xmris/processing/phasing.py:100-157) */
#define XMR_METHOD_ACME 0
#define XMR_METHOD_PEAK_MINIMA 1
#define XMR_METHOD_POSITIVITY 2

int xmr_version(void);
const char* xmr_last_error(void);

/* Fused zero_fill -> apodize -> FFT(ortho via the window scale) -> fftshift [-> uniform phase].
 * Replaces: zero_fill   src/xmris/processing/fid.py:201-285  (implicit: padded points never touch HBM)
 *           apodize_exp src/xmris/processing/fid.py:105-144  (window table / separable window)
 *           to_spectrum src/xmris/processing/fid.py:9-42 = fft (fourier.py:117-173) + fftshift (fourier.py:10-32)
 *           phase       src/xmris/processing/phasing.py:56-73 (optional epilogue, uniform coordinates)
 * Also serves to_fid (fid.py:45-102) with inverse=1, in_shift=(n+1)/2, out_shift=0.
 *
 *   fid_dev   [batch, n_in] complex64      spec_dev [batch, n_out] complex64 (may be NULL: statistics only)
 *   pad_left  zeros before the data ("symmetric" position: (n_out-n_in)/2; "end": 0)
 *   in_shift  input index rotation: x[k] is read from fid[(k + in_shift) mod n_in]; requires n_in == n_out
 *   out_shift output index rotation: bin j is stored at (j + out_shift) mod n_out  (n_out/2 = fftshift)
 *   absmax_dev/argmax_dev  optional [batch]: max |S| (float) and its first index (in stored order) per spectrum;
 *                          argmax_dev may be NULL (maxima only: the cheapest statistics pass)
 */
int xmr_fid_to_spectrum_c64(const void* fid_dev, void* spec_dev, int64_t batch, int n_in, int n_out, int pad_left,
                            int window_mode, const float* window_dev, const float* win_rows_host, float scale,
                            int inverse, int in_shift, int out_shift, float* absmax_dev, int* argmax_dev,
                            int phase_mode, double ph_a_turns, double ph_b_turns, void* stream);

/* Pass 1 of autophase(mode="single"): per-spectrum max |S| for the GLOBAL argmax of phasing.py:229-231, with branch and
 * bound: after the second FFT stage every output obeys |X|^2 <= 16 * sum_b max_c |Z[k1][c][b]|^2; spectra whose bound is
 * below the running global maximum cannot hold the argmax and skip their last stage (absmax entry 0).  The global
 * maximum and its first row are exact and run-independent.  running_max2_dev: one float of caller-owned scratch, shared
 * by the chunks of one data set (reset_running_max = 1 on the first chunk).  Geometries outside the specialised path
 * (zero filling, table windows, n_out > 4096) run the plain statistics pass.
 */
int xmr_fid_absmax_pruned_c64(const void* fid_dev, int64_t batch, int n_in, int n_out, int pad_left, int window_mode,
                              const float* window_dev, const float* win_rows_host, float scale, float* absmax_dev,
                              float* running_max2_dev, int reset_running_max, void* stream);

/* Standalone elementwise ops for the un-fused accessor calls.
 * xmr_zero_fill_c64  replaces zero_fill   (fid.py:251)  out[b, pad_left + k] = in[b, k], zeros elsewhere
 * xmr_scale_rows_c64 replaces apodize_exp (fid.py:139)  out[b, k] = in[b, k] * w_dev[k]          (real weights)
 * xmr_rotate_rows_c64 replaces phase      (phasing.py:73) out[b, k] = in[b, k] * rot_dev[k]      (complex weights)
 */
int xmr_zero_fill_c64(const void* in_dev, void* out_dev, int64_t batch, int n_in, int n_out, int pad_left,
                      void* stream);
 /* xmr_roll_rows_c64  replaces fftshift / ifftshift (fourier.py:31-32, 57-58) out[b, (k+shift) mod n] = in[b, k]        */
int xmr_roll_rows_c64(const void* in_dev, void* out_dev, int64_t batch, int n, int shift, void* stream);
int xmr_scale_rows_c64(const void* in_dev, void* out_dev, int64_t batch, int n, const float* w_dev, void* stream);
int xmr_rotate_rows_c64(const void* in_dev, void* out_dev, int64_t batch, int n, const void* rot_dev, void* stream);
/* The same with both rotations of the index folded in and strided input rows (the pre- / post-chirp steps of the chirp-z
 * transform that serves lengths that are not powers of two, e.g. the 1972-point FIDs of bruker.py:84 after
 * remove_digital_filter; to_spectrum's ifftshift / fftshift, fourier.py:31-32):
 *   out[b, (j + out_shift) mod n] = in[b * in_stride + (j + in_shift) mod n] * rot_dev[j],   j < n <= in_stride        */
int xmr_rotate_rows_shift_c64(const void* in_dev, int64_t in_stride, void* out_dev, int64_t batch, int n, const void* rot_dev,
                              int in_shift, int out_shift, void* stream);

/* Per-spectrum phase: out[b,m] = in[b,m] * exp(2*pi*i*(a_turns[b] + b_turns[b]*m)).  (phasing.py:56-73 per voxel) */
int xmr_phase_each_c64(const void* in_dev, void* out_dev, int64_t batch, int n, const double* a_turns_dev,
                       const double* b_turns_dev, void* stream);

/* Global first-occurrence argmax over a [batch] array of per-spectrum maxima (phasing.py:229-231).
 * Writes {max value, flat index = spectrum*n + argmax[spectrum]} to out_dev (float, then int64 at byte offset 8);
 * argmax_dev may be NULL (flat index = spectrum*n: the row only). */
int xmr_global_argmax(const float* absmax_dev, const int* argmax_dev, int64_t batch, int n, void* out_dev,
                      void* stream);

/* Per-row statistics of an existing spectrum array: max |S| and its first index (numpy argmax order),
 * the first step of autophase / phase(pivot=None)  (phasing.py:49-53, 229-231). */
int xmr_row_absmax_c64(const void* spec_dev, int64_t batch, int n, float* absmax_dev, int* argmax_dev, void* stream);

/* autophase(mode="single") optimiser on ONE spectrum.  Replaces the scipy.optimize.differential_evolution call of
 * src/xmris/processing/phasing.py:270-287 (bounds p0 in [-180,180], p1 in [-4000,4000]; p0_only -> p1 = 0) by a
 * deterministic dense grid (float32) + nested float64 zoom refinement of the reference's objective
 * (method: ACME phasing.py:100-122, peak_minima :125-139, positivity :142-157).
 *   u0, du        u_m = u0 + du*m = (x_m - pivot)/(x_max - x_min)  -- the reference's phase ramp (phasing.py:56-69)
 *   target_idx, index_width   ROI of the local methods (phasing.py:233-247)
 *   result_dev    double[4] on the device: {p0 deg, p1 deg, objective value, 0}
 *   workspace_dev at least xmr_autophase_workspace_bytes() bytes of device memory, caller-owned
 */
int64_t xmr_autophase_workspace_bytes(void);
/* Geometry of the search (process-wide; not part of the reference's interface -- the reference exposes the optimiser's
 * knobs through **kwargs of differential_evolution, phasing.py:276-284): coarse grid steps in degrees, number of
 * mutually distinct coarse cells refined side by side (<= 8), nested 21x21 zoom levels, how many leading
 * levels run in float32, how many basins the remaining float64 levels keep, and the window shrink factor between zoom
 * levels 0 and 1 (5 between all later ones; each level's window = +-`half-width`, spacing half-width/10).
 * Defaults: 6, 15, 4, 4, 2, 2, 2.5 (final spacing 0.01 x 0.024 deg). */
int xmr_autophase_search_tuning(double p0_step_deg, double p1_step_deg, int starts, int levels, int f32_levels,
                                int late_starts, double first_ratio);
/* ACME only: how many mutually distinct basins are finished by the bounded Newton polish on the analytic gradient of the
 * objective (default 3; 0 = the float64 zoom levels of `levels` instead), how many float32 zoom levels localise them first
 * (default 1) and whether the finest direct-search level runs for wall-type optima (default 1).  The polish converges to
 * the local minimum to ~1e-4 deg -- where the reference's own optimiser lands when run to convergence. */
int xmr_autophase_search_polish(int starts, int f32_levels, int fine_b);
int xmr_autophase_search_c64(const void* spec_dev, int n, double u0, double du, int method, int target_idx,
                             int index_width, int p0_only, double* result_dev, void* workspace_dev, void* stream);

/* The autophase objective itself at k given candidates -- what the searches minimise, exposed so that it can be checked
 * against the reference's own score functions: _acme_score (src/xmris/processing/phasing.py:100-122),
 * _peak_minima_score (:125-139), _roi_positivity_score (:142-157).  Arguments as xmr_autophase_search_c64;
 *   p0_dev, p1_dev  double[k] degrees (device); out_dev double[k] objective values (device)
 *   use_f64         1: float64 accumulation (the searches' final levels); 0: float32 (their coarse levels)
 * Returns the reference's formula as written, including ACME's negative branch max(Re) <= 0 that the searches reject. */
int xmr_autophase_score_c64(const void* spec_dev, int n, double u0, double du, int method, int target_idx,
                            int index_width, const double* p0_dev, const double* p1_dev, int k, int use_f64,
                            double* out_dev, void* stream);

/* The whole chain per voxel in ONE pass (autophase mode="all"; the reference's NotImplementedError branch,
 * phasing.py:219-222, defined as "the reference's autophase applied to every 1-D spectrum on its own"):
 *   zero_fill -> apodize -> FFT -> fftshift -> per-spectrum (p0, p1) search on the shared-memory resident
 *   spectrum (pivot = that spectrum's own |S| maximum, phasing.py:229-238) -> phase -> store.
 *   in_dev            [batch, n_in] FIDs, or (input_is_spectrum=1, n_in == n_out) spectra in stored order
 *   du                u_m = u0 + du*m; u0 = -du*argmax per voxel, or u0_fixed when fixed_pivot=1 (target_coord)
 *   p0_dev, p1_dev    double[batch] degrees; pivot_dev int[batch] (index of the maximum); fun_dev float[batch]
 * n_out must be a power of two in [512, 8192].
 */
int xmr_chain_each_c64(const void* in_dev, void* out_dev, int64_t batch, int n_in, int n_out, int pad_left,
                       int input_is_spectrum, int window_mode, const float* window_dev, const float* win_rows_host,
                       float scale, int method, double du, int fixed_pivot, double u0_fixed, int fixed_target,
                       int index_width, int p0_only, double* p0_dev, double* p1_dev, int* pivot_dev, float* fun_dev,
                       void* stream);

/* The whole chain on HOST buffers in one synchronous call (numpy-side bindings need nothing but ctypes):
 *   fid_host [batch, n_in] complex64  ->  out_host [batch, n_out] complex64
 *   zero_fill -> apodize (window_host) -> to_spectrum [-> autophase]           (README.md:66-73 of the reference)
 * Uploads, kernels and downloads are pipelined over voxel chunks on internal streams; pageable buffers are page-locked
 * for the duration of the call.  Device scratch is a per-thread workspace (xmr_host_workspace_release() frees it).
 *   autophase_mode 0: none; 1: the reference's mode="single" (two passes over the device-resident FIDs);
 *                  2: per spectrum (mode="all", one pass).
 *   result_host    mode 1: double[4] = {p0 deg, p1 deg, index of the pivot (global |S| maximum) on the output axis, objective}
 *   p0_host .. fun_host  mode 2: per-voxel results [batch] (fun_host may be NULL)
 */
typedef struct xmr_host_chain_desc {
    int n_in, n_out, pad_left;
    const double* window_host; /* n_out weights incl. 1/sqrt(n_out) (fid.py:136 with the ortho norm), or NULL          */
    float scale;               /* used when window_host is NULL; 0 -> 1/sqrt(n_out)                                     */
    int autophase_mode;
    int method;                /* XMR_METHOD_*                                                                          */
    int index_width;           /* ROI half width in points (phasing.py:245-247)                                         */
    int p0_only;
    int fixed_pivot;           /* target_coord given: u0_fixed / fixed_target apply to every spectrum                   */
    double u0_fixed;
    int fixed_target;
    double du;                 /* (x[1]-x[0]) / (x_max - x_min) of the output axis (phasing.py:56-69)                   */
    int chunk;                 /* spectra per pipeline chunk; 0 -> 8192                                                 */
} xmr_host_chain_desc;

int xmr_chain_host_c64(const xmr_host_chain_desc* desc, const void* fid_host, void* out_host, int64_t batch,
                       double* result_host, double* p0_host, double* p1_host, int* pivot_host, float* fun_host);
int xmr_host_workspace_release(void);
/* mode 1 keeps the whole FID batch in device memory between its two passes when it fits; otherwise -- or when the batch
 * exceeds `bytes` set here (0, the default: whatever is free) -- the FID chunks are uploaded twice (pass 1 over
 * double-buffered chunks, the winning row fetched again, pass 2 over the re-uploaded chunks): data sets larger than HBM,
 * and long-lived processes that must not hold gigabytes per thread.  Same results either way.                           */
int xmr_host_chain_resident_limit(int64_t bytes);

/* The mode="single" chain on DEVICE-resident data in one call (the reference's `.xmr.zero_fill().xmr.apodize_exp()
 * .xmr.to_spectrum().xmr.autophase()` on an array that already lives in HBM): branch-and-bound pass 1 -> global argmax
 * (phasing.py:229-231) -> the winning spectrum -> (p0, p1) search (phasing.py:270-287) -> pass 2 with the fused phase
 * (phasing.py:289-290).  Only `n_in, n_out, pad_left, scale, method, index_width, p0_only, fixed_pivot, u0_fixed,
 * fixed_target, du` of the descriptor are read; the window is passed as prepared for xmr_fid_to_spectrum_c64
 * (window_mode / window_dev / win_rows_host).
 *   workspace_dev  at least xmr_chain_single_workspace_bytes(batch, n_out) bytes of device memory, caller-owned
 *   result_host    double[6] = {p0 deg, p1 deg, pivot index on the output axis, objective, global max |S|, winning row}
 * No host read-back between the passes: the winning row, its pivot and the phase parameters of pass 2 stay in device memory
 * (workspace control block); the call returns -- with pass 2 enqueued on `stream` -- once the 256-byte control block has
 * arrived on a side stream. */
int64_t xmr_chain_single_workspace_bytes(int64_t batch, int n_out);
/* diagnostic: how many calls of xmr_chain_single_dev_c64 replayed their captured CUDA graph (process-wide) */
int64_t xmr_chain_single_graph_launches(void);
/* {front part ms (pass 1 ... search's final kernel), pass 2 ms} of this thread's last device chain, from CUDA events recorded
 * on the caller's stream (waits for that chain's pass 2). */
int xmr_chain_single_last_timing(double* ms_out);

/* The same chain split at its one exchange point, for voxels sharded over several GPUs (one process per GPU).  The reference's
 * mode="single" needs the GLOBAL |S| argmax (phasing.py:229-231):
 *   front  pass 1 on this rank's shard, then its candidate slot {best FID row, max |S|, global row = row + row_offset}
 *          (xmr_chain_single_slot_bytes(n_in) bytes at slot_dev)
 *   ---    ONE all-gather of the slots (torch.distributed / NCCL), nothing else crosses the GPUs
 *   back   every rank picks the winner among the `world` gathered slots ON THE DEVICE (ties: lowest global row), transforms and
 *          searches it redundantly (deterministic: identical angles on every rank) and runs pass 2 on its shard.
 * No host read-back between the passes on any rank; result_host as xmr_chain_single_dev_c64 (winning row = global row). */
int64_t xmr_chain_single_slot_bytes(int n_in);
int xmr_chain_single_front_c64(const xmr_host_chain_desc* desc, const void* fid_dev, int64_t batch, int window_mode,
                               const float* window_dev, const float* win_rows_host, void* workspace_dev, int64_t row_offset,
                               void* slot_dev, void* stream);
int xmr_chain_single_back_c64(const xmr_host_chain_desc* desc, const void* fid_dev, void* spec_dev, int64_t batch,
                              int window_mode, const float* window_dev, const float* win_rows_host, void* workspace_dev,
                              const void* gathered_dev, int world, double* result_host, void* stream);
int xmr_chain_single_dev_c64(const xmr_host_chain_desc* desc, const void* fid_dev, void* spec_dev, int64_t batch,
                             int window_mode, const float* window_dev, const float* win_rows_host, void* workspace_dev,
                             double* result_host, void* stream);

/* baseline_als(da, dim, lam, p, n_iter) -- asymmetric least squares baseline correction of the REAL part, the step after
 * autophase in the reference's pipeline (src/xmris/processing/baseline.py:10-119; accessor core/accessor.py:552-597).
 * Replaces the per-spectrum scipy.sparse spsolve loop (baseline.py:27-37) by a batched banded LDL^T in float64, one thread
 * per spectrum, with the reference's re-weighting w = p*(y > z) + (1-p)*(y < z) between the n_iter solves.
 *   in_dev          [batch, n] complex64 (in_is_complex=1: the real part is used, baseline.py:84-85) or float32
 *   out_dev         [batch, n] float32: real part minus its baseline (baseline.py:99)
 *   workspace_dev   factor scratch, xmr_baseline_als_workspace_bytes(batch, n) bytes for full occupancy (any multiple of
 *                   one CTA's share works: fewer CTAs run at a time) */
int64_t xmr_baseline_als_workspace_bytes(int64_t batch, int n);
int xmr_baseline_als(const void* in_dev, int in_is_complex, float* out_dev, int64_t batch, int n, double lam, double p,
                     int n_iter, void* workspace_dev, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XMRIS_B200_H */
