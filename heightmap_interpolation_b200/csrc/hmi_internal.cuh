// Internals shared by the translation units that implement the C ABI (hmi_capi.cu: single solver,
// hmi_halo.cu: peer-memory halo links, hmi_group.cu: several GPUs driven by one process).
// Nothing here is part of the public contract (include/hmi_b200.h).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <vector>

#include "hmi_common.cuh"

namespace hmi {

cudaError_t device_count_patiently(int *ndev);   // cudaGetDeviceCount that rides out a driver still starting (hmi_capi.cu)


int fail(int code, const char *fmt, ...);           // records the message hmi_last_error() returns

#define CUDA_TRY(expr)                                                                               \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return hmi::fail(HMI_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),   \
                             __FILE__, __LINE__);                                                    \
    } while (0)

inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

// per-process cache of freed device blocks (exact size match, per device)
cudaError_t cached_malloc(int device, void **p, size_t bytes);
void cached_free(int device, void *p, size_t bytes);

// Driver entry points resolved at run time (the library is not linked against libcuda, so that it
// still loads — and fails loudly at the first compute call — on a machine without a driver).
struct DriverApi {
    CUresult (*streamWriteValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    CUresult (*streamWaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    CUresult (*tensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
};
const DriverApi *driver_api();                      // nullptr (with the error recorded) if unavailable

// One side of a slab's halo link.  Side 0 = the neighbour above (feeds my top ghost rows), side 1 = below.
struct HaloSide {
    char *peer_block = nullptr;     // the neighbour's halo block mapped into this process / device
    bool ipc = false;               // peer_block came from cudaIpcOpenMemHandle
};

}  // namespace hmi

struct hmi_solver {
    hmi_params p;
    int elem = 0;            // bytes per cell
    int radius = 1;          // stencil radius of one sweep
    long long pitch = 0;     // elements per row
    long long mpitch = 0;    // bytes per mask row
    int phys_rows = 0;       // ghost_top + rows + ghost_bottom
    char *buf[2] = {nullptr, nullptr};
    uint8_t *mask = nullptr;
    int cur = 0;
    void *d_scratch = nullptr;   // one block: partials | result | counter | error word
    size_t scratch_bytes = 0;
    double *d_partials = nullptr;
    unsigned int *d_counter = nullptr;
    unsigned int *d_error = nullptr;   // set by device code that gave up (halo wait timed out)
    double *d_result = nullptr;
    double *h_result = nullptr;  // pinned, recycled: [0..3] norm sums, [4] error word
    int max_blocks = 0;
    bool use_stream = false;
    bool use_nl = false;     // streaming kernel for TV / AMLE
    int tb = 1;              // sweeps fused per launch
    int chunk_rows = 0;      // 0 = auto
    int vec = 0;             // streaming kernel: cells per lane (0 = default)
    int tma = 0;             // streaming kernel: 1 = 1-D bulk copies, 2 = 2-D tensor-map boxes (interior strips)
    long long launches = 0;
    bool has_problem = false;
    size_t grid_bytes = 0, mask_bytes = 0;
    cudaEvent_t ev = nullptr;       // main -> side stream dependency of the split last block
    cudaEvent_t ev_side = nullptr;  // side -> main: edge rows produced and pushed
    cudaStream_t own = nullptr;     // this solver's stream for uploads, downloads and one-shot solves
    cudaStream_t own_side = nullptr;  // high-priority side stream (edge rows + halo push) when the caller passes none
    cudaStream_t last_main = nullptr, last_side = nullptr;   // streams the latest sweeps were queued on
    bool side_pending = false;      // work on last_side that last_main has not waited for yet

    // peer-memory halo link (hmi_halo.cu)
    char *halo_block = nullptr;     // [stage top p0 | top p1 | bottom p0 | bottom p1 | flags]
    size_t halo_bytes = 0, halo_stage = 0;
    hmi::HaloSide peer[2];
    unsigned halo_gen = 0;          // exchanges issued so far (both neighbours count in lockstep)
    int halo_state = 0;             // 0 = ghost rows fresh, 1 = pushes in flight (wait + copy in), 2 = stale
    bool linked = false;
    CUtensorMap *tmaps = nullptr;   // host copies of the tensor maps of buf[0], buf[1], mask (tma == 2)
    int tmaps_v = 0;                // cells per lane they were encoded for
};

// One grid cut into row slabs over several devices of this process (hmi_group.cu)
struct hmi_group {
    hmi_params p;                       // the whole grid (rows = global rows, ghost_* = 0)
    int n = 0;
    int radius = 1;
    int exchange_every = 1;             // sweeps between halo exchanges
    int ghost = 0;                      // radius * exchange_every
    std::vector<int> devices;
    std::vector<hmi_solver *> slab;
    std::vector<int> r0, r1;            // global interior rows [r0, r1) of each slab
    bool has_problem = false;
};

namespace hmi {

// hmi_capi.cu
int run_block(hmi_solver *s, int k, int lo, int hi, bool do_norm, cudaStream_t st, bool flip = true);
int do_sweeps(hmi_solver *s, long long n, bool norm_on_last, cudaStream_t st);
int do_sweeps_split(hmi_solver *s, long long n, cudaStream_t ms, cudaStream_t ss);   // last block: edge rows first, on ss
int fetch_norm(hmi_solver *s, cudaStream_t st, hmi_norm *out);
double norm_to_diff(const hmi_params &p, const hmi_norm &n);
int sync_solver(hmi_solver *s);                       // wait for everything queued for this solver
cudaStream_t pick_stream(hmi_solver *s, void *stream);   // NULL -> the legacy default stream, as documented

int fill_unknown(hmi_solver *s, double value);         // unknown cells of the current grid <- value (own stream)

// The reference loop (fd_pde_inpainter.py:311-351), event driven: sweeps are issued in batches up to
// the next iteration index i with i % T == 0, where the convergence measure of the sweep i -> i+1 is
// taken; returns after i+1 updates when it passes, else after max_iters.  `sweep(n, norm, &nm)` runs n
// sweeps and, if `norm`, reduces the measure of the last one.
template <class Sweep>
int reference_loop(const hmi_params &p, Sweep &&sweep, hmi_status *status)
{
    const long long T = p.term_check_iters, maxit = p.max_iters;
    long long i = 0;
    status->updates_done = 0;
    status->last_diff = 100000.0;  // fd_pde_inpainter.py:309
    status->converged = 0;
    status->checks = 0;
    while (i < maxit) {
        const long long c = (i + T - 1) / T * T;  // next iteration index with c % T == 0
        if (c >= maxit) {                         // no check left before max_iters
            const int rc = sweep(maxit - i, false, nullptr);
            if (rc != HMI_OK) return rc;
            i = maxit;
            break;
        }
        hmi_norm nm;
        const int rc = sweep(c - i + 1, true, &nm);
        if (rc != HMI_OK) return rc;
        i = c + 1;
        status->checks += 1;
        status->last_diff = norm_to_diff(p, nm);
        if (status->last_diff < p.term_thres) {   // NaN compares false: never terminates on NaN
            status->converged = 1;
            break;
        }
    }
    status->updates_done = i;
    return HMI_OK;
}

// hmi_halo.cu
int halo_ensure_fresh(hmi_solver *s, cudaStream_t st);   // ghost rows valid for the current iterate
int halo_push(hmi_solver *s, cudaStream_t st);           // my edge rows -> the neighbours' staging + flag
void halo_release(hmi_solver *s);

}  // namespace hmi
