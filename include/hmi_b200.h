/*
 * hmi_b200.h — C ABI of the B200-native FD-PDE inpainting solver.
 *
 * The reference (coronis-computing/heightmap_interpolation v1.0.7) is pure Python and has no
 * native boundary for this path; the boundary it does have is
 *
 *     Inpainter(**options).inpaint(image, mask)
 *         heightmap_interpolation/inpainting/fd_pde_inpainter.py:52-122 (options), :160-185 (inpaint)
 *     created by  create_inpainter_from_params(params)
 *         heightmap_interpolation/apps/apps_common.py:172-215
 *     and called from  interpolate()   heightmap_interpolation/apps/interpolate_netcdf4.py:200-204
 *
 * Every entry point below names the reference code it replaces.  All functions return 0 on success
 * and a negative hmi_error on failure; hmi_last_error() gives the message.  No C++ exceptions and
 * no torch types cross this boundary: plain pointers, sizes and PODs only.  A solver handle is not
 * thread-safe.  There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef HMI_B200_H
#define HMI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMI_ABI_VERSION 1

typedef struct hmi_solver hmi_solver;

enum hmi_error {
    HMI_OK = 0,
    HMI_ERR_BAD_ARG = -1,     /* maps to Python ValueError */
    HMI_ERR_CUDA = -2,        /* CUDA runtime/driver failure (RuntimeError) */
    HMI_ERR_NO_DEVICE = -3,   /* no CUDA device / wrong architecture (RuntimeError) */
    HMI_ERR_STATE = -4,       /* call out of order, e.g. run before set_problem */
    HMI_ERR_UNSUPPORTED = -5  /* valid in the reference, not built here yet */
};

enum hmi_dtype { HMI_F32 = 0, HMI_F64 = 1 };

/* Which subclass step function to run:
 *   HMI_HARMONIC  SobolevInpainter.step_fun   inpainting/sobolev_inpainter.py:15-17
 *   HMI_CCST      CCSTInpainter.step_fun      inpainting/ccst_inpainter.py:23-34
 *   HMI_TV        TVInpainter.step_fun        inpainting/tv_inpainter.py:20-33
 *   HMI_AMLE      AMLEInpainter.step_fun      inpainting/amle_inpainter.py:41-59 (2-D branch) */
enum hmi_method { HMI_HARMONIC = 0, HMI_CCST = 1, HMI_TV = 2, HMI_AMLE = 3 };

/* The `convolver` option decomposes into an orientation and a border rule
 * (inpainting/convolver.py:11-69, inpainting/convolve_at_mask.py:39-63):
 *   "opencv"                    -> +1 (correlation), HMI_BORDER_REFLECT101
 *   "masked", "masked-parallel" -> -1 (convolution), HMI_BORDER_REFLECT101
 *   "scipy-signal", "scipy-ndimage" -> -1, HMI_BORDER_ZERO  (convolver.py:17-18 maps both names
 *                                      to scipy.signal.convolve)
 *   HMI_BORDER_SYMMETRIC is scipy.ndimage's 'reflect' (only reachable through AMLE convolve_in_1d). */
enum hmi_border { HMI_BORDER_REFLECT101 = 0, HMI_BORDER_SYMMETRIC = 1, HMI_BORDER_ZERO = 2 };

/* term_check()  inpainting/fd_pde_inpainter.py:368-379.  "absolute_percent" is HMI_ABSOLUTE with
 * term_thres pre-multiplied by the z-range by the caller (set_term_criteria, :149-155). */
enum hmi_criteria { HMI_RELATIVE = 0, HMI_ABSOLUTE = 1 };

/* Kernel selection; AUTO is the product default, the others exist for tests and benchmarks. */
enum hmi_kernel { HMI_KERNEL_AUTO = 0, HMI_KERNEL_GENERIC = 1, HMI_KERNEL_STREAM = 2 };

/* Constructor options (FDPDEInpainter.__init__, inpainting/fd_pde_inpainter.py:71-100, plus the
 * subclass options tension / epsilon) after the Python layer has validated them. */
typedef struct hmi_params {
    int32_t struct_size;      /* sizeof(hmi_params), for ABI checking */
    int32_t device;           /* CUDA device ordinal */
    int32_t rows, cols;       /* interior rows of this slab (whole grid on one GPU), columns */
    int32_t dtype;            /* hmi_dtype: arithmetic follows the input grid's dtype */
    int32_t method;           /* hmi_method */
    int32_t orientation;      /* +1 correlation, -1 convolution */
    int32_t border;           /* hmi_border */
    int32_t criteria;         /* hmi_criteria */
    int32_t kernel;           /* hmi_kernel */
    int32_t temporal_block;   /* sweeps fused per launch by the streaming kernel; 0 = auto */
    int32_t ghost_top;        /* ghost rows above the slab: 0 = global top edge (border rule applies) */
    int32_t ghost_bottom;     /* ghost rows below the slab: 0 = global bottom edge */
    int32_t reserved;
    double dt;                /* update_step_size */
    double tension;           /* CCST */
    double epsilon;           /* TV */
    double relaxation;        /* 0 or (1,2]: over-relaxation (fd_pde_inpainter.py:317-318) */
    double term_thres;
    int64_t term_check_iters;
    int64_t max_iters;
} hmi_params;

typedef struct hmi_status {
    int64_t updates_done;     /* sweeps executed; the reference only prints this */
    double last_diff;         /* value of the last convergence measure (100000 before any check) */
    int32_t converged;        /* 1 if a check passed (the reference returns), 0 = max_iters reached */
    int32_t checks;           /* number of convergence checks evaluated */
} hmi_status;

/* Partial convergence sums of one slab, for the multi-GPU all-reduce:
 * sum (fnew-f)^2, sum fnew^2 (ncclSum) and max|fnew-f|, NaN flag (ncclMax). */
typedef struct hmi_norm {
    double sum_d2, sum_f2, max_abs, has_nan;
} hmi_norm;

int hmi_abi_version(void);
const char *hmi_last_error(void);

/* Number of CUDA devices visible, or a negative hmi_error. */
int hmi_device_count(void);

/* Create / destroy a solver for one grid (or one row slab of it).  Allocates the two ping-pong
 * grids, the mask and the reduction scratch on `device`.
 * Replaces: FDPDEInpainter.__init__ state + the arrays inpaint_grid() allocates per call
 * (fd_pde_inpainter.py:281-309). */
int hmi_create(const hmi_params *params, hmi_solver **out);
void hmi_destroy(hmi_solver *s);

/* Load the *initialised* grid (unknown cells already filled by the initialiser; known cells are
 * what update_at_mask re-imposes every sweep, update_at_mask.py:6-14) and the mask
 * (1 byte per cell, non-zero = known).  Arrays cover ghost_top + rows + ghost_bottom rows, dense
 * row-major with the given row pitches in BYTES.  _host copies from host memory (pageable or
 * pinned) inside the call; _device reads device pointers on `stream` (a cudaStream_t, NULL = default).
 * Replaces: `f = image`, `mask_inv = 1 - mask` (fd_pde_inpainter.py:291-303). */
int hmi_set_problem_host(hmi_solver *s, const void *image, size_t image_pitch_bytes,
                         const uint8_t *mask, size_t mask_pitch_bytes);
int hmi_set_problem_device(hmi_solver *s, const void *image_dev, size_t image_pitch_bytes,
                           const uint8_t *mask_dev, size_t mask_pitch_bytes, void *stream);

/* n raw relaxation sweeps f <- Pi(f + dt*step(f)) without convergence checks, asynchronous on
 * `stream`.  Replaces n iterations of the loop body fd_pde_inpainter.py:311-326 between checks. */
int hmi_sweeps(hmi_solver *s, int64_t n, void *stream);

/* Row slabs: n sweeps like hmi_sweeps, but the last temporal block is split so that the rows the
 * neighbouring slabs need (the first ghost_top and last ghost_bottom interior rows) are produced
 * first, on `side_stream`, and the interior rows on `main_stream`.  The caller starts the halo
 * exchange on `side_stream` right after this call and makes `main_stream` wait for it before the
 * next call, so the exchange overlaps the interior update.  Halos must be fresh on entry. */
int hmi_sweeps_overlap(hmi_solver *s, int64_t n, void *main_stream, void *side_stream);

/* Peer-memory halo links between row slabs — the multi-GPU data path (csrc/hmi_halo.cu).  The
 * reference has nothing comparable (one process, whole grid in RAM: fd_pde_inpainter.py:281-366);
 * these calls replace what a grouped ncclSend/ncclRecv per neighbour would do, without occupying SMs:
 * the producer's copy engine pushes its rows next to a seam into a staging area in the neighbour's HBM
 * (double buffered by exchange parity) and publishes the exchange number in a flag there with a stream
 * memory operation; the consumer's next temporal block waits for the flag and moves the rows into its
 * ghost rows.  Works between devices of one process (hmi_halo_connect) and between processes
 * (hmi_halo_connect_ipc with the 64-byte CUDA IPC handle hmi_halo_setup returns; exchange the handles
 * with any transport, e.g. torch.distributed.all_gather).  All slabs of one solve must use the same
 * ghost depth and call the same sequence of hmi_sweeps* functions.  Nobody may destroy a slab while a
 * neighbour can still push into it (synchronise / barrier first). */
typedef struct hmi_halo_info {
    void *block;                  /* this slab's halo block (device pointer, valid in this process) */
    uint64_t block_bytes;
    int32_t device;
    int32_t reserved;
    unsigned char ipc_handle[64]; /* cudaIpcMemHandle_t of `block` */
} hmi_halo_info;
int hmi_halo_setup(hmi_solver *s, hmi_halo_info *out);
/* side 0 = the slab above (it feeds my top ghost rows), side 1 = the slab below */
int hmi_halo_connect(hmi_solver *s, int32_t side, void *peer_block, int32_t peer_device);
int hmi_halo_connect_ipc(hmi_solver *s, int32_t side, const unsigned char *ipc_handle);

/* n <= exchange depth sweeps on a linked slab, then the halo exchange of the result, overlapped with
 * the interior update of the last temporal block: stale or in-flight ghost rows are brought up to date
 * first; the rows next to the seams are produced first on `side_stream` (NULL = the solver's own
 * high-priority stream) and pushed to the neighbours from there while `main_stream` still updates the
 * interior.  hmi_sweeps / hmi_sweeps_norm on a linked slab refresh the ghost rows themselves too, but
 * leave them stale (no push). */
int hmi_sweeps_exchange(hmi_solver *s, int64_t n, void *main_stream, void *side_stream);

/* n sweeps (n >= 1); the convergence sums between the last two iterates are reduced on the device,
 * fused into the last sweep, and returned (synchronises `stream`).
 * Replaces: the loop body + term_check() on a check iteration (fd_pde_inpainter.py:322-323, 368-379). */
int hmi_sweeps_norm(hmi_solver *s, int64_t n, void *stream, hmi_norm *out);

/* The whole reference loop on one GPU: checks at i % term_check_iters == 0, returns after i+1
 * updates when a check passes, else after max_iters (status->converged = 0; the reference prints a
 * warning).  Replaces: FDPDEInpainter.inpaint_grid  fd_pde_inpainter.py:281-366. */
int hmi_run(hmi_solver *s, void *stream, hmi_status *status);

/* Copy the current iterate (interior rows only) out.  _host synchronises. */
int hmi_get_result_host(hmi_solver *s, void *out, size_t out_pitch_bytes);
int hmi_get_result_device(hmi_solver *s, void *out_dev, size_t out_pitch_bytes, void *stream);

/* One-shot convenience over host buffers: create + set_problem_host + run + get_result_host +
 * destroy; host<->device copies happen inside.  This is the end-to-end call bench.py times.
 * Replaces: FDPDEInpainter.inpaint_grid(image, mask) called with NumPy arrays. */
int hmi_inpaint_host(const hmi_params *params, const void *image, const uint8_t *mask, void *out,
                     hmi_status *status);

/* On large grids (>= 2^26 cells, >= 4096 rows, streaming kernels) the one-shot calls upload the grid in row
 * bands and overlap the upload with the first temporal blocks of the bands that have already landed
 * (csrc/hmi_capi.cu: upload_pipelined); the result is bit-identical to upload-then-solve.  Environment
 * variable HMI_UPLOAD_PIPELINE=0 turns it off (=2 forces it on small grids, for tests).  The counter says
 * how many one-shot solves of this process took the pipelined route (tests, bench). */
int64_t hmi_upload_pipeline_count(void);

/* Introspection for the CPU tests (pure host functions, no device needed).
 * hmi_debug_band_steps: the launch order of the overlapped copies for nb row bands.  phase 0 = upload head start
 * (every band runs nb launches), phase 1 = download drain of m launches per band.  Writes 4 ints per step —
 * q, j_lo, j_hi, final: "launch q on the bands j_lo ... j_hi as one kernel launch; afterwards band `final` (or -1)
 * is final" — and, in phase 0, a step with q = -1 when band j_lo has landed in HBM.  Returns the number of steps.
 * hmi_debug_pick_chunks: the chunk heights (interior strips, column-edge strips) the wave-aware picker of the
 * streaming kernels chooses for a launch of nrows x nwin strips. */
int hmi_debug_band_steps(int32_t nb, int32_t m, int32_t phase, int32_t *out, int32_t cap);
int hmi_debug_pick_chunks(int32_t nrows, int32_t nwin, int32_t n_edge, int32_t halo, int32_t wpb, int32_t cta_slots,
                          double edge_cost, int32_t min_chunk, int32_t *chunk_rows, int32_t *chunk_rows_edge);

/* hmi_inpaint_host with the 'zeros' / 'mean' initialiser fused in: after the upload the unknown cells
 * (mask == 0) of the device copy are set to `fill_value`, so the caller can initialise its own host
 * array (the reference does that in place, inpainting/initializer.py:15-18) concurrently with the
 * upload instead of before it.  `image` itself is only read. */
int hmi_inpaint_host_fill(const hmi_params *params, const void *image, const uint8_t *mask, double fill_value,
                          void *out, hmi_status *status);

/* hmi_inpaint_host_fill that also initialises the caller's array in place (unknown cells of `image` <-
 * fill_value, what Initializer.initialize does first: inpainting/initializer.py:15-18) — on a host
 * thread, after the upload has read the array and while the device solves. */
int hmi_inpaint_host_init(const hmi_params *params, void *image, const uint8_t *mask, double fill_value, void *out,
                          hmi_status *status);

/* Multigrid (FDPDEInpainter.inpaint_multigrid, inpainting/fd_pde_inpainter.py:255-264): load the next
 * finer level into `fine` — upload its image and mask like hmi_set_problem_host, then replace the
 * device copy by
 *     image = cv2.resize(coarse solution, (cols, rows)) * (~mask) + image * mask
 * where the coarse solution is the current iterate of `coarse` (it never leaves the device).  The
 * bilinear taps (sx, sx1 / sy, sy1: source column / row pairs) and weights (ax0, ax1 / by0, by1, in
 * the grid's dtype) are 1-D tables built by the host exactly as OpenCV's INTER_LINEAR does (half-pixel
 * centres); products and sums are individually rounded in the order of the NumPy restatement.  With
 * scrub_nan the NaNs of the uploaded image are zeroed first (level 0, :259-260) and
 * info->image_had_nan reports whether there were any, so that the caller can do the same to its own
 * array.  info->min/max are the extrema of the blended grid (NaN if it holds one): the z-range
 * set_term_criteria needs for 'absolute_percent'. */
typedef struct hmi_blend_info {
    double min, max;
    int32_t has_nan;          /* the blended grid contains a NaN */
    int32_t image_had_nan;    /* the uploaded image contained a NaN (only checked with scrub_nan) */
} hmi_blend_info;
int hmi_set_problem_host_blend(hmi_solver *fine, const void *image, size_t image_pitch_bytes, const uint8_t *mask,
                               size_t mask_pitch_bytes, hmi_solver *coarse, const int32_t *sx, const int32_t *sx1,
                               const void *ax0, const void *ax1, const int32_t *sy, const int32_t *sy1,
                               const void *by0, const void *by1, int32_t scrub_nan, hmi_blend_info *info);

/* Several GPUs driven by one process: the grid is cut into contiguous row slabs, one per device,
 * linked through peer memory (hmi_halo_*), and all of them are driven by the calling thread.  This is
 * what `Inpainter(..., devices=[...]).inpaint(image, mask)` runs on, so that the reference's single call
 * site (apps/interpolate_netcdf4.py:200-204) can use the whole box.  `params` describes the whole grid
 * (ghost_* = 0; `device` is ignored); `exchange_every` = sweeps between halo exchanges (0 = default: 18
 * for the radius-1 methods — 16 for fp32 harmonic —, 12 for CCST), each slab carries radius * exchange_every ghost rows.  Small
 * grids use fewer devices than listed.  The functions mirror their single-solver namesakes. */
typedef struct hmi_group hmi_group;
int hmi_group_create(const hmi_params *params, const int32_t *devices, int32_t n_devices, int32_t exchange_every,
                     hmi_group **out);
void hmi_group_destroy(hmi_group *g);
int hmi_group_set_problem_host(hmi_group *g, const void *image, size_t image_pitch_bytes, const uint8_t *mask,
                               size_t mask_pitch_bytes);
int hmi_group_fill_unknown(hmi_group *g, double value);      /* 'zeros' / 'mean' initialiser on the device copies */
int hmi_group_sweeps(hmi_group *g, int64_t n);
int hmi_group_sweeps_norm(hmi_group *g, int64_t n, hmi_norm *out);
int hmi_group_run(hmi_group *g, hmi_status *status);         /* FDPDEInpainter.inpaint_grid, fd_pde_inpainter.py:281-366 */
int hmi_group_set_term_thres(hmi_group *g, double term_thres);
int hmi_group_get_result_host(hmi_group *g, void *out, size_t out_pitch_bytes);
int hmi_group_sync(hmi_group *g);
/* Multigrid with the pyramid on the device(s) (FDPDEInpainter.inpaint_multigrid, fd_pde_inpainter.py:187-279):
 * load pyramid level l = log2(stride) into `level` without touching the host again.  `finest` holds the
 * full-resolution image and mask (hmi_group_set_problem_host); level l is every stride-th cell of it
 * (halve_image applied l times, misc/image_proc.py:4-12) with the mask ANDed with "image is not NaN"
 * (:223-224); its unknown cells take cv2.resize(current iterate of `coarse`) and its known cells the image,
 * by the reference's arithmetic blend (:255-263; taps and weights as for hmi_set_problem_host_blend, sy /
 * by* indexed by the level's global row).  stride 1 blends `finest` itself in place (level == finest), with
 * scrub_nan as at level 0 (:259-260).  The three groups may use different device lists; rows are read
 * across GPUs through peer memory. */
int hmi_group_set_problem_pyramid(hmi_group *level, hmi_group *finest, int32_t stride, hmi_group *coarse,
                                  const int32_t *sx, const int32_t *sx1, const void *ax0, const void *ax1,
                                  const int32_t *sy, const int32_t *sy1, const void *by0, const void *by1,
                                  int32_t scrub_nan, hmi_blend_info *info);
int hmi_group_info(const hmi_group *g, int32_t *n_slabs, int32_t *exchange_every, int32_t *ghost_rows);
int hmi_group_slab(hmi_group *g, int32_t i, hmi_solver **solver, int32_t *row0, int32_t *row1);
int64_t hmi_group_launch_count(const hmi_group *g);
/* One-shot over host buffers on several devices.  fill: 0 = hmi_inpaint_host, 1 = hmi_inpaint_host_fill,
 * 2 = hmi_inpaint_host_init (`image` is then written: its unknown cells take fill_value). */
int hmi_inpaint_host_devices(const hmi_params *params, const int32_t *devices, int32_t n_devices, const void *image,
                             const uint8_t *mask, int32_t fill, double fill_value, void *out, hmi_status *status);

/* One evaluation of the step function on a host array: out = step(f) where `work_mask` is non-zero
 * (everywhere if it is NULL), the input value elsewhere.  Uses params->rows, cols, dtype, method,
 * orientation, border, tension, epsilon, device.  With a work mask the result equals the reference's at
 * every cell whose stencil support lies inside the mask (all unknown cells, for the dilated mask the
 * reference passes, fd_pde_inpainter.py:293-296); outside it the reference's masked convolvers return
 * by-products of copying, which only harmonic reproduces (the input value).
 * Replaces: step_fun(f, mask)  sobolev_inpainter.py:15-17, ccst_inpainter.py:23-34, tv_inpainter.py:20-33,
 * amle_inpainter.py:22-59 (the abstract hook fd_pde_inpainter.py:406-408). */
int hmi_step_fun_host(const hmi_params *params, const void *f, size_t f_pitch_bytes, const uint8_t *work_mask,
                      size_t mask_pitch_bytes, void *out, size_t out_pitch_bytes);

/* One 3x3 filter on a host array with a convolver's orientation (+1 correlation = "opencv", -1
 * convolution = the masked and SciPy convolvers) and border rule; evaluated where `mask` is non-zero
 * (everywhere if NULL), copying the input elsewhere.  kernel9 = the 3x3 kernel in row-major order.
 * Replaces: Convolver.__call__(image, kernel, mask)  inpainting/convolver.py:28-29 and the back ends
 * :47-69 (cv2.filter2D, scipy.signal.convolve, nb_convolve_at_mask convolve_at_mask.py:39-63). */
int hmi_convolve3x3_host(int32_t rows, int32_t cols, int32_t dtype, int32_t device, const void *image,
                         size_t image_pitch_bytes, const double *kernel9, int32_t orientation, int32_t border,
                         const uint8_t *mask, size_t mask_pitch_bytes, void *out, size_t out_pitch_bytes);

/* The gridded branch of the reference CLI with the grid resident on the device
 * (interpolate(), apps/interpolate_netcdf4.py:176-212; same code in apps/interpolate_xyz.py:255-279).
 * hmi_areas_create uploads `elevation` and `mask_int` (non-zero = cell to interpolate) once and keeps
 * elevation_int = copy(elevation) on the device (:56).  Per work area: hmi_areas_set_work_area uploads the
 * area's layer (NULL = one area covering the grid, netcdf_data_io.py:262); hmi_areas_load_window builds the
 * window problem [r0, r0 + rows) x [c0, c0 + cols) (the sizes of `g`) straight into the group's slabs —
 * crop (:187-188), known = ~mask_int | ~area (:187-190), NaN -> 0 (:192) — and returns the count, sum, minimum
 * and maximum of the known cells (what the 'mean' initialiser and the 'absolute_percent' z-range need);
 * after the solve hmi_areas_paste writes the inpainted cells, and only those, into elevation_int
 * (:210-212); hmi_areas_get_result_host downloads elevation_int. */
typedef struct hmi_areas hmi_areas;
typedef struct hmi_window_info {
    int64_t known, unknown;
    double known_sum, known_min, known_max;
} hmi_window_info;
int hmi_areas_create(int32_t rows, int32_t cols, int32_t dtype, int32_t device, const void *elevation,
                     size_t elevation_pitch_bytes, const uint8_t *mask_int, size_t mask_pitch_bytes, hmi_areas **out);
void hmi_areas_destroy(hmi_areas *a);
int hmi_areas_set_work_area(hmi_areas *a, const uint8_t *work_area, size_t pitch_bytes);
int hmi_areas_load_window(hmi_areas *a, hmi_group *g, int32_t r0, int32_t c0, hmi_window_info *info);
int hmi_areas_paste(hmi_areas *a, hmi_group *g, int32_t r0, int32_t c0);
int hmi_areas_get_result_host(hmi_areas *a, void *elevation_int, size_t pitch_bytes);

/* Multi-GPU plumbing: device address and row pitch (bytes) of the current iterate and of the mask,
 * so the host layer can exchange ghost rows with NCCL (row 0 = first ghost row).  The addresses
 * change after every hmi_sweeps*() call (ping-pong). */
int hmi_current_grid(hmi_solver *s, void **grid_dev, size_t *pitch_bytes);
int hmi_mask_grid(hmi_solver *s, uint8_t **mask_dev, size_t *pitch_bytes);

/* Host-side initialiser helpers (multi-threaded): image[i] = value where mask[i] == 0, and the mean
 * of the known cells.  Replaces: Initializer.initialize 'zeros' / 'mean', inpainting/initializer.py:15-18
 * (in place on the caller's host array, as the reference does). */
int hmi_host_fill_unknown(void *image, const uint8_t *mask, int64_t n, int32_t dtype, double value);
int hmi_host_known_mean(const void *image, const uint8_t *mask, int64_t n, int32_t dtype, double *mean);

/* The 'nearest' initialiser on the device: every unknown cell (mask == 0) of the host array `image`
 * takes the value of its Euclidean-nearest known cell, in place; known cells are untouched.  Host to
 * device, an exact two-pass nearest-cell search and device to host all happen inside the call.
 * Fails with HMI_ERR_BAD_ARG if there is no known cell.  Ties between equidistant known cells are
 * broken towards the smaller column offset, then left, then up (SciPy's choice depends on its
 * KD-tree's leaf order; the distance is always the same).
 * Replaces: Initializer.initialize with init_with='nearest' (the CLI default, apps/apps_common.py:52),
 * i.e. scipy.interpolate.griddata(method='nearest') over all known cells, inpainting/initializer.py:19-34. */
int hmi_init_nearest_host(void *image, size_t image_pitch_bytes, const uint8_t *mask, size_t mask_pitch_bytes,
                          int32_t rows, int32_t cols, int32_t dtype, int32_t device);

/* Minimum and maximum over the known cells (NaN-propagating like np.max / np.min) and their count,
 * multi-threaded: with the fill value this gives the z-range of the initialised grid that
 * set_term_criteria needs for 'absolute_percent' (inpainting/fd_pde_inpainter.py:149-155) without two
 * more passes over the array. */
int hmi_host_known_minmax(const void *image, const uint8_t *mask, int64_t n, int32_t dtype, double *min_out,
                          double *max_out, int64_t *known_out);

/* Pre-fault a freshly allocated host buffer with several threads (the result array of inpaint()). */
int hmi_host_prefault(void *buf, size_t bytes);

/* Device blocks of destroyed solvers are kept in a per-process cache (the reference builds a fresh
 * inpainter per work area, apps/interpolate_netcdf4.py:200); this returns them to the driver. */
int hmi_release_cache(void);

/* Tuning knobs for tests and benchmarks: "chunk_rows" (rows per warp chunk of the streaming kernel,
 * 0 = wave-aware automatic choice, -1 = the older fixed rule), "temporal_block" (sweeps fused per launch),
 * "vec" (fp64 cells per lane: 2 or 4) and "tma" (row staging of the harmonic / CCST kernel: 0 = cp.async
 * ring, 1 = one 1-D bulk copy per row, 2 = 2-D tensor-map boxes on interior strips). */
int hmi_set_option(hmi_solver *s, const char *key, int64_t value);

/* Replace the stop threshold of hmi_run.  'absolute_percent' scales term_thres by the z-range of the
 * initialised grid (set_term_criteria, inpainting/fd_pde_inpainter.py:149-155), which the multigrid
 * driver only knows after the device-side blend (info->min/max of hmi_set_problem_host_blend). */
int hmi_set_term_thres(hmi_solver *s, double term_thres);

/* Introspection for tests and bench: sweeps fused per launch, kernels launched so far, and the
 * name of the kernel family the solver selected ("generic", "stream"). */
int hmi_temporal_block(const hmi_solver *s);
int64_t hmi_launch_count(const hmi_solver *s);
const char *hmi_kernel_name(const hmi_solver *s);

#ifdef __cplusplus
}
#endif
#endif /* HMI_B200_H */
