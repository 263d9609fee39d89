// C ABI of the B200 FD-PDE inpainting solver (see include/hmi_b200.h for the contract and the
// reference code each entry point replaces).  Host-side orchestration only: buffer ownership,
// kernel selection, the temporal-block schedule and the reference's check cadence
// (heightmap_interpolation/inpainting/fd_pde_inpainter.py:311-351).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <atomic>
#include <chrono>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "hmi_internal.cuh"
#include "hmi_stream.cuh"
#include "hmi_stream_nl.cuh"
#include "hmi_init.cuh"
#include "hmi_xfer.cuh"

namespace hmi {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// A small per-process cache of device allocations: the reference builds a fresh inpainter (and
// fresh arrays) per work area (apps/interpolate_netcdf4.py:200), which on the device would mean a
// cudaMalloc/cudaFree pair per 100+ MiB grid per call.  Freed blocks are kept (per device, exact
// size match) up to a byte budget and handed back to the next solver.  The budget is the smaller of
// 24 GiB and a quarter of the device's memory, so that other allocators in the process (torch, NCCL)
// are not starved; hmi_release_cache() returns everything to the driver.
struct CachedBlock { int device; size_t bytes; void *ptr; };
std::mutex g_cache_mu;
std::vector<CachedBlock> g_cache;
size_t g_cache_bytes = 0;
size_t g_cache_budget = 0;                      // 0 = not sized yet

static void drop_blocks(std::vector<CachedBlock> &drop)
{
    int prev = 0;
    cudaGetDevice(&prev);
    for (auto &b : drop) {
        cudaSetDevice(b.device);
        cudaFree(b.ptr);
    }
    cudaSetDevice(prev);
}

cudaError_t cached_malloc(int device, void **p, size_t bytes)
{
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (size_t i = 0; i < g_cache.size(); ++i)
            if (g_cache[i].device == device && g_cache[i].bytes == bytes) {
                *p = g_cache[i].ptr;
                g_cache_bytes -= bytes;
                g_cache.erase(g_cache.begin() + i);
                return cudaSuccess;
            }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {       // drop the cache and retry once
        cudaGetLastError();
        std::vector<CachedBlock> drop;
        {
            std::lock_guard<std::mutex> lk(g_cache_mu);
            drop.swap(g_cache);
            g_cache_bytes = 0;
        }
        drop_blocks(drop);
        e = cudaMalloc(p, bytes);
    }
    return e;
}

void cached_free(int device, void *p, size_t bytes)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        if (g_cache_budget == 0) {
            size_t free_b = 0, total_b = 0;
            g_cache_budget = (size_t)24 << 30;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b / 4 < g_cache_budget) g_cache_budget = total_b / 4;
            cudaGetLastError();
        }
        if (bytes >= ((size_t)1 << 20) && g_cache_bytes + bytes <= g_cache_budget && g_cache.size() < 64) {
            g_cache.push_back({device, bytes, p});
            g_cache_bytes += bytes;
            return;
        }
    }
    cudaFree(p);
}

const DriverApi *driver_api()
{
    static DriverApi api;
    static int state = 0;                       // 0 = not tried, 1 = ok, -1 = failed
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (state == 0) {
        struct { const char *name; void **fn; } want[] = {
            {"cuStreamWriteValue32", (void **)&api.streamWriteValue32},
            {"cuStreamWaitValue32", (void **)&api.streamWaitValue32},
            {"cuTensorMapEncodeTiled", (void **)&api.tensorMapEncodeTiled},
        };
        state = 1;
        for (auto &w : want) {
            cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
            const cudaError_t e = cudaGetDriverEntryPoint(w.name, w.fn, cudaEnableDefault, &q);
            if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*w.fn) {
                cudaGetLastError();
                fail(HMI_ERR_CUDA, "driver entry point %s is not available", w.name);
                state = -1;
                break;
            }
        }
    }
    return state == 1 ? &api : nullptr;
}

}  // namespace hmi

using hmi::fail;
using hmi::round_up;
namespace hmi {
// The first CUDA call of a process on a box that has just come up can fail with an initialisation
// error while the driver is still starting (seen once on a fresh B200 box, profiles/README.md); that
// is not "no device", so ask again a few times before giving up.
cudaError_t device_count_patiently(int *ndev)
{
    cudaError_t e = cudaSuccess;
    for (int attempt = 0; attempt < 5; ++attempt) {
        e = cudaGetDeviceCount(ndev);
        if (e != cudaErrorInitializationError && e != cudaErrorSystemNotReady && e != cudaErrorDevicesUnavailable) break;
        cudaGetLastError();
        std::this_thread::sleep_for(std::chrono::milliseconds(1000));
    }
    return e;
}
}  // namespace hmi

using hmi::cached_malloc;
using hmi::cached_free;

namespace {

using hmi::g_cache_mu;

std::vector<double *> g_pinned;                 // recycled 4-double pinned result buffers

double *pinned_get()
{
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        if (!g_pinned.empty()) {
            double *p = g_pinned.back();
            g_pinned.pop_back();
            return p;
        }
    }
    double *p = nullptr;
    if (cudaMallocHost((void **)&p, 8 * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void pinned_put(double *p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_pinned.push_back(p);
}

int g_cc_major[64] = {0};                       // compute capability major per device (0 = unknown)

// The streaming kernel packs four mask bytes with one multiply, which needs them to be exactly 0/1.
__global__ void normalize_mask_kernel(uint8_t *m, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint32_t *w = reinterpret_cast<uint32_t *>(m);
    for (size_t q = i; q < n / 4; q += stride) {
        const uint32_t v = w[q];
        w[q] = ((v & 0xffu) ? 1u : 0u) | ((v & 0xff00u) ? 0x100u : 0u) |
               ((v & 0xff0000u) ? 0x10000u : 0u) | ((v & 0xff000000u) ? 0x1000000u : 0u);
    }
}

// Fused 'zeros' / 'mean' initialiser: unknown cells of the uploaded grid take `value`.
template <typename T>
__global__ void fill_unknown_kernel(T *f, long long pitch, const uint8_t *m, long long mpitch, int rows, int cols, T value)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    for (int i = blockIdx.y; i < rows; i += gridDim.y)
        if (!m[(long long)i * mpitch + j]) f[(long long)i * pitch + j] = value;
}

// Pipelined upload (inpaint_host_impl): fold the convergence sums of one row band into the accumulator.
__global__ void accumulate_norm_kernel(double *acc, const double *r, int first)
{
    if (first) { acc[0] = r[0]; acc[1] = r[1]; acc[2] = r[2]; acc[3] = r[3]; return; }
    acc[0] += r[0];
    acc[1] += r[1];
    acc[2] = fmax(acc[2], r[2]);
    acc[3] = fmax(acc[3], r[3]);
}

}  // namespace

namespace hmi {

// Tensor maps for the 2-D TMA staging of the streaming kernel (tma == 2): one per ping-pong grid and
// one for the mask, boxes of 32*V columns x STREAM_BOX_ROWS rows, encoded once per solver and V.
static int ensure_tmaps(hmi_solver *s, int V)
{
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
    if (s->tmaps && s->tmaps_v == V) return HMI_OK;
    const DriverApi *drv = driver_api();
    if (!drv) return HMI_ERR_CUDA;
    if (!s->tmaps) s->tmaps = new (std::nothrow) CUtensorMap[3];
    if (!s->tmaps) return fail(HMI_ERR_CUDA, "out of host memory");
    const cuuint32_t estr[2] = {1, 1};
    for (int i = 0; i < 3; ++i) {
        const bool is_mask = (i == 2);
        // the mask box starts at the 16-byte aligned column below the strip and is 16 bytes wider (hmi_stream.cu)
        const cuuint32_t box[2] = {(cuuint32_t)(32 * V + (is_mask ? 16 : 0)), (cuuint32_t)STREAM_BOX_ROWS};
        const cuuint64_t dims[2] = {(cuuint64_t)(is_mask ? s->mpitch : s->pitch), (cuuint64_t)s->phys_rows};
        const cuuint64_t strides[1] = {(cuuint64_t)(is_mask ? s->mpitch : s->pitch * s->elem)};
        const CUtensorMapDataType dt = is_mask ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                               : (s->elem == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        void *base = is_mask ? (void *)s->mask : (void *)s->buf[i];
        const CUresult r = drv->tensorMapEncodeTiled(&s->tmaps[i], dt, 2, base, dims, strides, box, estr,
                                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            delete[] s->tmaps;
            s->tmaps = nullptr;
            return fail(HMI_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        }
    }
    s->tmaps_v = V;
    return HMI_OK;
}

template <typename T>
static int run_block_t(hmi_solver *s, int k, int lo, int hi, bool do_norm, cudaStream_t st, bool flip)
{
    const hmi_params &p = s->p;
    // Over-relaxation (fd_pde_inpainter.py:317-318): fnew = f(1-w) + (f + dt step)w = f + (w dt) step for
    // every method, i.e. a larger step; the streaming kernels take it as such.  (The generic kernel
    // keeps the reference's own expression.)
    const double dt_eff = p.relaxation > 1.0 ? p.dt * p.relaxation : p.dt;
    const T *src = reinterpret_cast<const T *>(s->buf[s->cur]) + (long long)p.ghost_top * s->pitch;
    T *dst = reinterpret_cast<T *>(s->buf[s->cur ^ 1]) + (long long)p.ghost_top * s->pitch;
    const uint8_t *mask = s->mask + (long long)p.ghost_top * s->mpitch;
    int nblocks = 0;
    cudaError_t e;
    if (s->use_nl) {
        if (k < 1 || k > hmi::stream_nl_max_k(p.method, p.dtype, p.border))
            return fail(HMI_ERR_STATE, "the TV/AMLE kernel fuses at most two (fp32: three) sweeps");
        hmi::NLLaunch<T> l;
        memset(&l, 0, sizeof(l));
        l.src = src; l.dst = dst; l.mask = mask;
        l.pitch = s->pitch; l.mpitch = s->mpitch;
        l.rows = p.rows; l.cols = p.cols;
        l.row_lo = lo; l.row_hi = hi;
        l.read_lo = -p.ghost_top; l.read_hi = p.rows + p.ghost_bottom;
        l.top_border = p.ghost_top == 0; l.bottom_border = p.ghost_bottom == 0;
        l.method = p.method; l.sgn = p.orientation; l.border = p.border;
        l.chunk_rows = s->chunk_rows;
        l.k = k;
        l.do_norm = do_norm ? 1 : 0;
        l.dt = (T)dt_eff; l.eps2 = (T)(p.epsilon * p.epsilon);
        l.partials = s->d_partials; l.counter = s->d_counter; l.result = s->d_result;
        e = hmi::launch_stream_nl<T>(l, st, &nblocks);
    } else if (s->use_stream) {
        hmi::StreamArgs<T> a;
        memset(&a, 0, sizeof(a));
        a.src = src; a.dst = dst; a.mask = mask;
        a.pitch = s->pitch; a.mpitch = s->mpitch;
        a.rows = p.rows; a.cols = p.cols;
        a.row_lo = lo; a.row_hi = hi;
        a.read_lo = -p.ghost_top; a.read_hi = p.rows + p.ghost_bottom;
        a.top_border = p.ghost_top == 0; a.bottom_border = p.ghost_bottom == 0;
        a.chunk_rows = s->chunk_rows;
        a.vec = s->vec;
        a.border = p.border;
        // fp64 harmonic on grids of many waves: 4 cells per lane (half the shuffles per cell, 112 of 128 columns
        // of a strip useful instead of 52 of 64) beat 2 by 8-13 % at 16384^2 and 32768^2 and lose 6 % at 4096^2,
        // where the launch is a single wave (profiles/r04_tune_peeled_steady_loop.log); CCST needs 234 registers
        // with 4 cells and gains nothing.  Results are bit-identical either way.
        if (s->vec == 0 && s->tma == 0 && p.dtype == HMI_F64 && p.method == HMI_HARMONIC && !do_norm && k >= 4 &&
            (long long)(hi - lo) * p.cols >= (1LL << 26))
            a.vec = 4;
        a.tma = s->tma;
        if (s->tma == 2) {
            const int rc = ensure_tmaps(s, hmi::stream_cells_per_lane(p.dtype, s->vec));
            if (rc != HMI_OK) return rc;
            a.tmap_f = s->tmaps[s->cur];
            a.tmap_m = s->tmaps[2];
            a.phys_row0 = p.ghost_top;
        }
        a.dt = (T)dt_eff; a.tension = (T)p.tension; a.one_minus_tension = (T)(1.0 - p.tension);
        a.do_norm = do_norm ? 1 : 0;
        a.partials = s->d_partials; a.counter = s->d_counter; a.result = s->d_result;
        e = hmi::launch_stream<T>(p.method, k, a, st, &nblocks);
    } else {
        if (k != 1) return fail(HMI_ERR_STATE, "generic kernel runs one sweep per launch");
        hmi::SweepArgs<T> a;
        memset(&a, 0, sizeof(a));
        a.src = src; a.dst = dst; a.mask = mask;
        a.pitch = s->pitch; a.mpitch = s->mpitch;
        a.rows = p.rows; a.cols = p.cols;
        a.row_lo = lo; a.row_hi = hi;
        a.top_border = p.ghost_top == 0; a.bottom_border = p.ghost_bottom == 0;
        a.border = p.border; a.s = p.orientation;
        a.dt = (T)p.dt; a.tension = (T)p.tension; a.one_minus_tension = (T)(1.0 - p.tension);
        a.eps2 = (T)(p.epsilon * p.epsilon);
        a.relax = p.relaxation > 1.0 ? 1 : 0;
        a.relax_w = (T)p.relaxation; a.one_minus_relax_w = (T)(1.0 - p.relaxation);
        a.do_norm = do_norm ? 1 : 0;
        a.partials = s->d_partials; a.counter = s->d_counter; a.result = s->d_result;
        e = hmi::launch_generic<T>(p.method, a, st, &nblocks);
    }
    if (e != cudaSuccess)
        return fail(HMI_ERR_CUDA, "sweep kernel launch failed: %s", cudaGetErrorString(e));
    if (nblocks > s->max_blocks)
        return fail(HMI_ERR_STATE, "internal: %d blocks exceed the reduction scratch (%d)", nblocks,
                    s->max_blocks);
    s->launches += 1;
    if (flip) s->cur ^= 1;
    return HMI_OK;
}

int run_block(hmi_solver *s, int k, int lo, int hi, bool do_norm, cudaStream_t st, bool flip)
{
    return s->p.dtype == HMI_F64 ? run_block_t<double>(s, k, lo, hi, do_norm, st, flip)
                                 : run_block_t<float>(s, k, lo, hi, do_norm, st, flip);
}

static int check_ghost_depth(const hmi_solver *s, long long n)
{
    const hmi_params &p = s->p;
    const long long need = n * s->radius;
    if ((p.ghost_top > 0 && need > p.ghost_top) || (p.ghost_bottom > 0 && need > p.ghost_bottom))
        return fail(HMI_ERR_BAD_ARG, "%lld sweeps of radius %d need %lld ghost rows, the slab has %d/%d", n,
                    s->radius, need, p.ghost_top, p.ghost_bottom);
    return HMI_OK;
}

// Blocks [done, upto) of an n-sweep call: block j also advances the ghost rows that later blocks of
// the call still need (a shrinking trapezoid), so no exchange is needed inside the call.
static int trapezoid_blocks(hmi_solver *s, long long n, long long upto, bool norm_on_last, cudaStream_t st)
{
    const hmi_params &p = s->p;
    long long done = 0;
    while (done < upto) {
        long long k = upto - done;
        if (k > s->tb) k = s->tb;
        // the streaming kernel fuses the convergence sums into a single-sweep launch: keep the
        // last sweep of a check window on its own
        if (norm_on_last && (s->use_stream || s->use_nl) && done + k == n && k > 1) k -= 1;
        const long long after = n - (done + k);  // sweeps still to run after this block
        const int lo = p.ghost_top > 0 ? (int)(-after * s->radius) : 0;
        const int hi = p.rows + (p.ghost_bottom > 0 ? (int)(after * s->radius) : 0);
        const bool last = (done + k == n);
        const int rc = run_block(s, (int)k, lo, hi, norm_on_last && last, st, true);
        if (rc != HMI_OK) return rc;
        done += k;
    }
    return HMI_OK;
}

// Work queued on a side stream by an earlier call (edge rows, halo pushes) must be finished before
// `st` touches the grids again.
static int join_side(hmi_solver *s, cudaStream_t st)
{
    if (s->side_pending && s->ev_side) CUDA_TRY(cudaStreamWaitEvent(st, s->ev_side, 0));
    s->side_pending = false;
    return HMI_OK;
}

// n sweeps as a sequence of temporal blocks.  On a slab with ghost rows the halos must be fresh
// on entry and n*radius must not exceed the ghost depth.
int do_sweeps(hmi_solver *s, long long n, bool norm_on_last, cudaStream_t st)
{
    if (!s->has_problem) return fail(HMI_ERR_STATE, "hmi_set_problem_* has not been called");
    if (n < 0) return fail(HMI_ERR_BAD_ARG, "negative sweep count");
    int rc = check_ghost_depth(s, n);
    if (rc != HMI_OK) return rc;
    if ((rc = join_side(s, st)) != HMI_OK) return rc;
    s->last_main = st;
    if (n == 0) return HMI_OK;
    // linked slabs (hmi_halo_*): bring the ghost rows up to date first; afterwards they are stale
    if (s->linked && (rc = halo_ensure_fresh(s, st)) != HMI_OK) return rc;
    rc = trapezoid_blocks(s, n, n, norm_on_last, st);
    if (s->linked) s->halo_state = 2;
    return rc;
}

// n sweeps with the last temporal block split: the rows the neighbouring slabs need (the first
// ghost_top and last ghost_bottom interior rows) are produced first, on the side stream, and the
// interior on the main stream.  On return ev_side has been recorded on `ss` behind the edge launches;
// callers queue the halo transfer on `ss` and record ev_side again.
int do_sweeps_split(hmi_solver *s, long long n, cudaStream_t ms, cudaStream_t ss)
{
    const hmi_params &p = s->p;
    if (!s->has_problem) return fail(HMI_ERR_STATE, "hmi_set_problem_* has not been called");
    if (n < 1) return fail(HMI_ERR_BAD_ARG, "the split sweep call needs n >= 1");
    int rc = check_ghost_depth(s, n);
    if (rc != HMI_OK) return rc;
    if ((rc = join_side(s, ms)) != HMI_OK) return rc;
    const int bt = p.ghost_top, bb = p.ghost_bottom;     // rows each neighbour needs from this slab
    if (!s->ev) CUDA_TRY(cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming));
    if (!s->ev_side) CUDA_TRY(cudaEventCreateWithFlags(&s->ev_side, cudaEventDisableTiming));
    s->last_main = ms;
    s->last_side = ss;
    if ((bt == 0 && bb == 0) || p.rows - bt - bb < 8 || ss == ms) {
        // nothing to overlap with: everything on the main stream; the side stream still has to see it
        if ((rc = trapezoid_blocks(s, n, n, false, ms)) != HMI_OK) return rc;
        CUDA_TRY(cudaEventRecord(s->ev, ms));
        CUDA_TRY(cudaStreamWaitEvent(ss, s->ev, 0));
        CUDA_TRY(cudaEventRecord(s->ev_side, ss));
        s->side_pending = (ss != ms);
        return HMI_OK;
    }
    long long k_last = n % s->tb == 0 ? s->tb : n % s->tb;
    if (k_last > n) k_last = n;
    // all blocks but the last: the shrinking trapezoid, on the main stream
    if ((rc = trapezoid_blocks(s, n, n - k_last, false, ms)) != HMI_OK) return rc;
    // last block: the rows the neighbours are waiting for go first, on the side stream; the
    // interior follows on the main stream
    CUDA_TRY(cudaEventRecord(s->ev, ms));
    CUDA_TRY(cudaStreamWaitEvent(ss, s->ev, 0));
    if (bt > 0) rc = run_block(s, (int)k_last, 0, bt, false, ss, false);
    if (rc == HMI_OK && bb > 0) rc = run_block(s, (int)k_last, p.rows - bb, p.rows, false, ss, false);
    if (rc == HMI_OK) rc = run_block(s, (int)k_last, bt, p.rows - bb, false, ms, true);
    if (rc != HMI_OK) return rc;
    CUDA_TRY(cudaEventRecord(s->ev_side, ss));
    s->side_pending = true;
    return HMI_OK;
}

int fetch_norm(hmi_solver *s, cudaStream_t st, hmi_norm *out)
{
    // result[0..3] and the device error word behind the counter travel in one copy
    CUDA_TRY(cudaMemcpyAsync(s->h_result, s->d_result, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    out->sum_d2 = s->h_result[0];
    out->sum_f2 = s->h_result[1];
    out->max_abs = s->h_result[2];
    out->has_nan = s->h_result[3];
    const unsigned int *w = reinterpret_cast<const unsigned int *>(s->h_result + 4);
    if (w[2] != 0u)
        return fail(HMI_ERR_CUDA, "halo exchange timed out: a neighbouring slab did not deliver its rows "
                                  "(device error word %u)", w[2]);
    return HMI_OK;
}

double norm_to_diff(const hmi_params &p, const hmi_norm &n)
{
    if (p.criteria == HMI_RELATIVE) return sqrt(n.sum_d2) / sqrt(n.sum_f2);  // 0/0 -> NaN like NumPy
    return n.has_nan != 0.0 ? nan("") : n.max_abs;
}

cudaStream_t pick_stream(hmi_solver *, void *stream) { return (cudaStream_t)stream; }

int sync_solver(hmi_solver *s)
{
    CUDA_TRY(cudaSetDevice(s->p.device));
    CUDA_TRY(cudaStreamSynchronize(s->last_main));        // NULL = the legacy default stream
    if (s->last_side) CUDA_TRY(cudaStreamSynchronize(s->last_side));
    if (s->own) CUDA_TRY(cudaStreamSynchronize(s->own));
    if (s->own_side) CUDA_TRY(cudaStreamSynchronize(s->own_side));
    s->side_pending = false;
    return HMI_OK;
}

}  // namespace hmi

namespace {

void destroy(hmi_solver *s)
{
    if (!s) return;
    cudaSetDevice(s->p.device);
    hmi::sync_solver(s);                         // nothing may still be using the blocks we recycle
    cudaGetLastError();
    hmi::halo_release(s);
    cached_free(s->p.device, s->buf[0], s->grid_bytes);
    cached_free(s->p.device, s->buf[1], s->grid_bytes);
    cached_free(s->p.device, s->mask, s->mask_bytes);
    cached_free(s->p.device, s->d_scratch, s->scratch_bytes);
    pinned_put(s->h_result);
    if (s->ev) cudaEventDestroy(s->ev);
    if (s->ev_side) cudaEventDestroy(s->ev_side);
    if (s->own) cudaStreamDestroy(s->own);
    if (s->own_side) cudaStreamDestroy(s->own_side);
    delete[] s->tmaps;
    delete s;
}

}  // namespace

namespace hmi {

// 'zeros' / 'mean' initialiser on the device copy (both ghost and interior rows), on the solver's stream
int fill_unknown(hmi_solver *s, double value)
{
    if (!s->has_problem) return fail(HMI_ERR_STATE, "no problem loaded");
    CUDA_TRY(cudaSetDevice(s->p.device));
    const dim3 grid((s->p.cols + 255) / 256, s->phys_rows < 1024 ? s->phys_rows : 1024);
    char *f = s->buf[s->cur];
    if (s->p.dtype == HMI_F64)
        fill_unknown_kernel<double><<<grid, 256, 0, s->own>>>((double *)f, s->pitch, s->mask, s->mpitch, s->phys_rows,
                                                              s->p.cols, value);
    else
        fill_unknown_kernel<float><<<grid, 256, 0, s->own>>>((float *)f, s->pitch, s->mask, s->mpitch, s->phys_rows,
                                                             s->p.cols, (float)value);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(HMI_ERR_CUDA, "fill kernel launch failed: %s", cudaGetErrorString(e));
    s->last_main = s->own;
    return HMI_OK;
}

}  // namespace hmi

using hmi::do_sweeps;
using hmi::fetch_norm;
using hmi::norm_to_diff;
using hmi::run_block;

extern "C" {

int hmi_abi_version(void) { return HMI_ABI_VERSION; }

const char *hmi_last_error(void) { return hmi::g_err; }

int hmi_device_count(void)
{
    int n = 0;
    const cudaError_t e = hmi::device_count_patiently(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(HMI_ERR_NO_DEVICE, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    return n;
}

int hmi_create(const hmi_params *params, hmi_solver **out)
{
    if (!params || !out) return fail(HMI_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    if (params->struct_size != (int32_t)sizeof(hmi_params))
        return fail(HMI_ERR_BAD_ARG, "hmi_params size mismatch: caller %d, library %d",
                    params->struct_size, (int)sizeof(hmi_params));
    const hmi_params &p = *params;
    if (p.rows < 2 || p.cols < 2) return fail(HMI_ERR_BAD_ARG, "grid must be at least 2x2");
    if (p.dtype != HMI_F32 && p.dtype != HMI_F64) return fail(HMI_ERR_BAD_ARG, "bad dtype");
    if (p.method < HMI_HARMONIC || p.method > HMI_AMLE) return fail(HMI_ERR_BAD_ARG, "bad method");
    if (p.orientation != 1 && p.orientation != -1) return fail(HMI_ERR_BAD_ARG, "orientation must be +1/-1");
    if (p.border < HMI_BORDER_REFLECT101 || p.border > HMI_BORDER_ZERO) return fail(HMI_ERR_BAD_ARG, "bad border");
    if (p.criteria != HMI_RELATIVE && p.criteria != HMI_ABSOLUTE) return fail(HMI_ERR_BAD_ARG, "bad criteria");
    if (!(p.dt > 0.0)) return fail(HMI_ERR_BAD_ARG, "update_step_size must be larger than zero");
    if (!(p.term_thres > 0.0)) return fail(HMI_ERR_BAD_ARG, "term_thres must be larger than zero");
    if (p.max_iters <= 0) return fail(HMI_ERR_BAD_ARG, "max_iters must be larger than zero");
    if (p.term_check_iters <= 0) return fail(HMI_ERR_BAD_ARG, "term_check_iters must be larger than zero");
    if (p.relaxation != 0.0 && (p.relaxation < 1.0 || p.relaxation > 2.0))
        return fail(HMI_ERR_BAD_ARG, "relaxation must be a number between 1 and 2 (0 to deactivate)");
    if (p.method == HMI_CCST && (p.tension < 0.0 || p.tension > 1.0))
        return fail(HMI_ERR_BAD_ARG, "tension parameter must be a number between 0 and 1 (included)");
    if (p.method == HMI_TV && !(p.epsilon > 0.0)) return fail(HMI_ERR_BAD_ARG, "epsilon must be larger than zero");
    if (p.ghost_top < 0 || p.ghost_bottom < 0) return fail(HMI_ERR_BAD_ARG, "negative ghost rows");
    if (p.temporal_block < 0) return fail(HMI_ERR_BAD_ARG, "negative temporal_block");

    int ndev = 0;
    cudaError_t e = hmi::device_count_patiently(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(HMI_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (p.device < 0 || p.device >= ndev) return fail(HMI_ERR_BAD_ARG, "device %d out of range (%d devices)", p.device, ndev);
    if (p.device < 64 && g_cc_major[p.device] == 0) {
        int major = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, p.device));
        g_cc_major[p.device] = major;
    }
    if (p.device < 64 && g_cc_major[p.device] != 10)
        return fail(HMI_ERR_NO_DEVICE, "device %d has compute capability %d.x; this library is built for sm_100a (B200) only",
                    p.device, g_cc_major[p.device]);
    CUDA_TRY(cudaSetDevice(p.device));

    hmi_solver *s = new (std::nothrow) hmi_solver();
    if (!s) return fail(HMI_ERR_CUDA, "out of host memory");
    s->p = p;
    s->elem = p.dtype == HMI_F64 ? 8 : 4;
    s->radius = hmi::method_radius(p.method);
    s->pitch = round_up(p.cols, 32);
    s->mpitch = round_up(p.cols, 128);
    s->phys_rows = p.ghost_top + p.rows + p.ghost_bottom;

    // kernel selection
    // (zero padding = the SciPy convolvers; the symmetric rule is only reachable through AMLE convolve_in_1d)
    const bool stream_ok = (p.method == HMI_HARMONIC || p.method == HMI_CCST) &&
                           (p.border == HMI_BORDER_REFLECT101 || p.border == HMI_BORDER_ZERO);
    int tb = 1;
    bool use_stream = false;
    if (p.kernel == HMI_KERNEL_STREAM && !stream_ok && (p.method == HMI_HARMONIC || p.method == HMI_CCST)) {
        delete s;
        return fail(HMI_ERR_UNSUPPORTED, "streaming kernel needs reflect-101 or zero-padded borders");
    }
    if (p.kernel != HMI_KERNEL_GENERIC && stream_ok) {
        const int maxk = hmi::stream_max_k(p.method, p.dtype);
        // measured: fp64 CCST 3, harmonic 6 (profiles/r02_tune_*); fp32 CCST 4, harmonic 8 since the
        // steady-state row loop (profiles/r05_tune_f32_temporal_block.log: +4..16 % / +2..12 % over 3 / 6
        // from 4096^2 to 32768^2)
        const bool f32 = p.dtype == HMI_F32;
        tb = p.temporal_block > 0 ? p.temporal_block : (p.method == HMI_CCST ? (f32 ? 4 : 3) : (f32 ? 8 : 6));
        if (tb > maxk) tb = maxk;
        if (p.border == HMI_BORDER_ZERO)
            while (tb > 1 && !hmi::stream_zero_supports(p.method, tb)) --tb;
        // the mirrored ghost loads need the mirror cell to exist: shrink the block on small grids
        const int min_dim = p.cols < p.rows ? p.cols : p.rows;
        while (tb > 1 && 2 * hmi::stream_halo(p.method, tb) + 2 > min_dim) --tb;
        use_stream = 2 * hmi::stream_halo(p.method, tb) + 2 <= min_dim;
        if (!use_stream && p.kernel == HMI_KERNEL_STREAM) {
            delete s;
            return fail(HMI_ERR_UNSUPPORTED, "grid too small for the streaming kernel");
        }
        if (!use_stream) tb = 1;
    }
    s->use_stream = use_stream;
    s->tb = tb;
    // TV / AMLE: streaming kernel (reflect-101, grid not tiny)
    const bool nl_ok = hmi::stream_nl_supports(p.method, p.orientation, p.border) && p.rows >= 8 && p.cols >= 8;
    if (p.kernel == HMI_KERNEL_STREAM && (p.method == HMI_TV || p.method == HMI_AMLE) && !nl_ok) {
        delete s;
        return fail(HMI_ERR_UNSUPPORTED, "streaming TV/AMLE kernel: this orientation / border pair is served by the generic kernel (or the grid is under 8x8)");
    }
    s->use_nl = nl_ok && p.kernel != HMI_KERNEL_GENERIC;
    if (s->use_nl) {
        // two fused sweeps by default (profiles/r02_tune_tv_amle_k2.log); one on grids too small for the halo
        const int maxk = hmi::stream_nl_max_k(p.method, p.dtype, p.border);
        int k = p.temporal_block > 0 ? p.temporal_block : 2;
        if (k > maxk) k = maxk;
        if (p.rows < 16 || p.cols < 16) k = 1;
        s->tb = k;
    }

    const int mb_gen = hmi::generic_max_blocks(s->phys_rows, p.cols);
    const int mb_str = hmi::stream_max_blocks(p.method, s->phys_rows, p.cols);
    s->max_blocks = mb_gen > mb_str ? mb_gen : mb_str;

    const size_t grid_bytes = (size_t)s->phys_rows * s->pitch * s->elem;
    const size_t mask_bytes = (size_t)s->phys_rows * s->mpitch;
    s->grid_bytes = grid_bytes;
    s->mask_bytes = mask_bytes;
    int rc = HMI_OK;
    do {
        if ((e = cached_malloc(p.device, (void **)&s->buf[0], grid_bytes)) != cudaSuccess) break;
        if ((e = cached_malloc(p.device, (void **)&s->buf[1], grid_bytes)) != cudaSuccess) break;
        if ((e = cached_malloc(p.device, (void **)&s->mask, mask_bytes)) != cudaSuccess) break;
        {
            const size_t part = (size_t)s->max_blocks * 4 * sizeof(double);
            s->scratch_bytes = (size_t)round_up((long long)(part + 128), 1 << 20);   // result, counter, error word, band accumulator
            if ((e = cached_malloc(p.device, &s->d_scratch, s->scratch_bytes)) != cudaSuccess) break;
            s->d_partials = (double *)s->d_scratch;
            s->d_result = (double *)((char *)s->d_scratch + part);
            s->d_counter = (unsigned int *)((char *)s->d_scratch + part + 4 * sizeof(double));
            s->d_error = s->d_counter + 2;
        }
        s->h_result = pinned_get();
        if (!s->h_result) { e = cudaErrorMemoryAllocation; break; }
        // the solver's own streams: uploads / downloads / one-shot solves never touch the legacy
        // default stream, so several solvers (threads, GPUs) do not serialise on it
        if ((e = cudaStreamCreateWithFlags(&s->own, cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((e = cudaMemsetAsync(s->buf[0], 0, grid_bytes, s->own)) != cudaSuccess) break;
        if ((e = cudaMemsetAsync(s->buf[1], 0, grid_bytes, s->own)) != cudaSuccess) break;
        if ((e = cudaMemsetAsync(s->mask, 1, mask_bytes, s->own)) != cudaSuccess) break;
        if ((e = cudaMemsetAsync(s->d_result, 0, 6 * sizeof(double), s->own)) != cudaSuccess) break;
        // callers may continue on any (non-blocking) stream: the zero fill must not land later
        if ((e = cudaStreamSynchronize(s->own)) != cudaSuccess) break;
    } while (0);
    if (e != cudaSuccess) {
        rc = fail(HMI_ERR_CUDA, "device allocation failed (%zu bytes per grid): %s", grid_bytes, cudaGetErrorString(e));
        cudaGetLastError();
        destroy(s);
        return rc;
    }
    *out = s;
    return HMI_OK;
}

void hmi_destroy(hmi_solver *s) { destroy(s); }

int hmi_set_option(hmi_solver *s, const char *key, int64_t value)
{
    if (!s || !key) return fail(HMI_ERR_BAD_ARG, "null argument");
    if (!strcmp(key, "chunk_rows")) {
        // 0 = wave-aware automatic choice, -1 = the older fixed-wave rule (A/B measurements)
        if (value < -1) return fail(HMI_ERR_BAD_ARG, "chunk_rows must be >= -1");
        s->chunk_rows = value <= 0 ? (int)value : (value < 8 ? 8 : (int)value);
        return HMI_OK;
    }
    if (!strcmp(key, "tma")) {
        // 0 = cp.async ring (default), 1 = 1-D bulk copies per row, 2 = 2-D tensor-map boxes (interior strips)
        // 3 = the warp-specialised pipeline kernel (hmi_pipe.cu) for the launches that fuse >= 2 sweeps
        if (value < 0 || value > 3) return fail(HMI_ERR_BAD_ARG, "tma must be 0, 1, 2 or 3");
        s->tma = (int)value;
        return HMI_OK;
    }
    if (!strcmp(key, "vec")) {
        if (value != 0 && value != 2 && value != 4 && value != 8) return fail(HMI_ERR_BAD_ARG, "vec must be 0, 2, 4 or 8");
        s->vec = (int)value;
        return HMI_OK;
    }
    if (!strcmp(key, "temporal_block")) {
        if (s->use_nl) {
            const int maxk = hmi::stream_nl_max_k(s->p.method, s->p.dtype, s->p.border);
            if (value < 1 || value > maxk || (value > 1 && (s->p.rows < 16 || s->p.cols < 16)))
                return fail(HMI_ERR_BAD_ARG, "temporal_block out of range");
            s->tb = (int)value;
            return HMI_OK;
        }
        if (!s->use_stream) return fail(HMI_ERR_UNSUPPORTED, "temporal_block applies to the streaming kernel");
        const int maxk = hmi::stream_max_k(s->p.method, s->p.dtype);
        const int min_dim = s->p.cols < s->p.rows ? s->p.cols : s->p.rows;
        if (value < 1 || value > maxk || 2 * hmi::stream_halo(s->p.method, (int)value) + 2 > min_dim ||
            (s->p.border == HMI_BORDER_ZERO && !hmi::stream_zero_supports(s->p.method, (int)value)))
            return fail(HMI_ERR_BAD_ARG, "temporal_block out of range");
        s->tb = (int)value;
        return HMI_OK;
    }
    return fail(HMI_ERR_BAD_ARG, "unknown option '%s'", key);
}

int hmi_set_problem_host(hmi_solver *s, const void *image, size_t image_pitch_bytes,
                         const uint8_t *mask, size_t mask_pitch_bytes)
{
    if (!s || !image || !mask) return fail(HMI_ERR_BAD_ARG, "null argument");
    const size_t wbytes = (size_t)s->p.cols * s->elem;
    if (image_pitch_bytes < wbytes || mask_pitch_bytes < (size_t)s->p.cols)
        return fail(HMI_ERR_BAD_ARG, "pitch smaller than a row");
    CUDA_TRY(cudaSetDevice(s->p.device));
    s->cur = 0;
    // pinned memory: straight DMA (one linear copy when dense); pageable memory (NumPy arrays): staged
    // through pinned ring buffers by several host threads (hmi_xfer.cu)
    int rc = hmi::sync_solver(s);                // earlier sweeps may still be reading the grids
    if (rc != HMI_OK) return rc;
    CUDA_TRY(hmi::copy_h2d_2d(s->p.device, s->buf[0], (size_t)s->pitch * s->elem, image, image_pitch_bytes, wbytes,
                              s->phys_rows, s->own));
    CUDA_TRY(hmi::copy_h2d_2d(s->p.device, s->mask, (size_t)s->mpitch, mask, mask_pitch_bytes, (size_t)s->p.cols,
                              s->phys_rows, s->own));
    normalize_mask_kernel<<<592, 256, 0, s->own>>>(s->mask, (size_t)s->phys_rows * s->mpitch);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(s->own));     // the caller may reuse its host buffers on return
    s->has_problem = true;
    s->halo_state = 0;                           // slabs are loaded with their neighbours' rows
    return HMI_OK;
}

int hmi_set_problem_device(hmi_solver *s, const void *image_dev, size_t image_pitch_bytes,
                           const uint8_t *mask_dev, size_t mask_pitch_bytes, void *stream)
{
    if (!s || !image_dev || !mask_dev) return fail(HMI_ERR_BAD_ARG, "null argument");
    const size_t wbytes = (size_t)s->p.cols * s->elem;
    if (image_pitch_bytes < wbytes || mask_pitch_bytes < (size_t)s->p.cols)
        return fail(HMI_ERR_BAD_ARG, "pitch smaller than a row");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(s->p.device));
    s->cur = 0;
    s->last_main = st;
    s->halo_state = 0;
    CUDA_TRY(cudaMemcpy2DAsync(s->buf[0], (size_t)s->pitch * s->elem, image_dev, image_pitch_bytes,
                               wbytes, s->phys_rows, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpy2DAsync(s->mask, (size_t)s->mpitch, mask_dev, mask_pitch_bytes,
                               (size_t)s->p.cols, s->phys_rows, cudaMemcpyDeviceToDevice, st));
    normalize_mask_kernel<<<592, 256, 0, st>>>(s->mask, (size_t)s->phys_rows * s->mpitch);
    CUDA_TRY(cudaGetLastError());
    s->has_problem = true;
    return HMI_OK;
}

int hmi_sweeps(hmi_solver *s, int64_t n, void *stream)
{
    if (!s) return fail(HMI_ERR_BAD_ARG, "null solver");
    CUDA_TRY(cudaSetDevice(s->p.device));
    return do_sweeps(s, n, false, (cudaStream_t)stream);
}

int hmi_sweeps_overlap(hmi_solver *s, int64_t n, void *main_stream, void *side_stream)
{
    if (!s) return fail(HMI_ERR_BAD_ARG, "null solver");
    if (n < 1) return fail(HMI_ERR_BAD_ARG, "hmi_sweeps_overlap needs n >= 1");
    CUDA_TRY(cudaSetDevice(s->p.device));
    // on every path the side stream ends up ordered behind the rows the neighbours need, so the
    // caller's exchange on it never sends rows that are still being computed
    const int rc = hmi::do_sweeps_split(s, n, (cudaStream_t)main_stream, (cudaStream_t)side_stream);
    s->side_pending = false;     // this entry point leaves joining the two streams to the caller
    return rc;
}

int hmi_sweeps_norm(hmi_solver *s, int64_t n, void *stream, hmi_norm *out)
{
    if (!s || !out) return fail(HMI_ERR_BAD_ARG, "null argument");
    if (n < 1) return fail(HMI_ERR_BAD_ARG, "hmi_sweeps_norm needs n >= 1");
    CUDA_TRY(cudaSetDevice(s->p.device));
    const int rc = do_sweeps(s, n, true, (cudaStream_t)stream);
    if (rc != HMI_OK) return rc;
    return fetch_norm(s, (cudaStream_t)stream, out);
}

int hmi_run(hmi_solver *s, void *stream, hmi_status *status)
{
    if (!s || !status) return fail(HMI_ERR_BAD_ARG, "null argument");
    const hmi_params &p = s->p;
    if (p.ghost_top || p.ghost_bottom)
        return fail(HMI_ERR_STATE, "hmi_run drives a whole grid; slabs are driven by hmi_group_run or the host layer");
    if (!s->has_problem) return fail(HMI_ERR_STATE, "hmi_set_problem_* has not been called");
    CUDA_TRY(cudaSetDevice(p.device));
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = hmi::reference_loop(p, [&](long long n, bool norm, hmi_norm *nm) {
        int r = do_sweeps(s, n, norm, st);
        if (r == HMI_OK && norm) r = fetch_norm(s, st, nm);
        return r;
    }, status);
    if (rc != HMI_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(st));
    return HMI_OK;
}

int hmi_get_result_host(hmi_solver *s, void *out, size_t out_pitch_bytes)
{
    if (!s || !out) return fail(HMI_ERR_BAD_ARG, "null argument");
    if (!s->has_problem) return fail(HMI_ERR_STATE, "no problem loaded");
    const size_t wbytes = (size_t)s->p.cols * s->elem;
    if (out_pitch_bytes < wbytes) return fail(HMI_ERR_BAD_ARG, "pitch smaller than a row");
    const int rc = hmi::sync_solver(s);
    if (rc != HMI_OK) return rc;
    const char *src = s->buf[s->cur] + (size_t)s->p.ghost_top * s->pitch * s->elem;
    CUDA_TRY(hmi::copy_d2h_2d(s->p.device, out, out_pitch_bytes, src, (size_t)s->pitch * s->elem, wbytes, s->p.rows,
                              s->own));
    return HMI_OK;
}

int hmi_get_result_device(hmi_solver *s, void *out_dev, size_t out_pitch_bytes, void *stream)
{
    if (!s || !out_dev) return fail(HMI_ERR_BAD_ARG, "null argument");
    if (!s->has_problem) return fail(HMI_ERR_STATE, "no problem loaded");
    const size_t wbytes = (size_t)s->p.cols * s->elem;
    if (out_pitch_bytes < wbytes) return fail(HMI_ERR_BAD_ARG, "pitch smaller than a row");
    CUDA_TRY(cudaSetDevice(s->p.device));
    if (s->side_pending && s->ev_side) CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, s->ev_side, 0));
    const char *src = s->buf[s->cur] + (size_t)s->p.ghost_top * s->pitch * s->elem;
    CUDA_TRY(cudaMemcpy2DAsync(out_dev, out_pitch_bytes, src, (size_t)s->pitch * s->elem, wbytes,
                               s->p.rows, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return HMI_OK;
}

// fill: 0 = the grid is already initialised; 1 = unknown cells of the device copy <- fill_value;
// 2 = as 1 and the caller's host array is initialised in place too (what the reference's Initializer
// does first, initializer.py:15-18) — after the upload has read it, while the device solves.
// ---- one-shot solve with the host <-> device copies overlapped with the solve -------------------
//
// hmi_inpaint_host* on a large grid is upload -> loop -> download, and the 18 GB a 32768^2 fp64 problem
// moves over PCIe are 0.34 s of a 1.84 s call in which the SMs have nothing to do.  Both halves can be
// hidden behind temporal blocks that run on row BANDS in a skewed order:
//   * head start: the grid goes up in row bands, and as soon as band b has landed the bands above it
//     advance — band b-1 does its first launch (iteration 0: one sweep with the convergence sums), band
//     b-2 its second (a block of `tb` fused sweeps), band b-3 its third ...  A launch of a band reads
//     radius*k halo rows of both neighbours at the previous time level, so a band may be one launch ahead
//     of the band below it and no more; issuing each anti-diagonal (band + launch index = const) from
//     its bottom band upwards on one stream satisfies every read-after-write and write-after-read
//     dependence of the two ping-pong grids.  When the last band has landed the triangle is completed
//     the other way round — the bands nb-1-q ... nb-1 are then all at time level q, so launch q runs on
//     them as ONE whole-wave launch — until every band has done the same nb launches = 1 + tb*(nb-1)
//     sweeps, and the loop carries on with whole-grid launches;
//   * drain: when a request of the loop is known to be the last one (it reaches max_iters — whatever its
//     check says, the grid it leaves is the result), its last nb launches run skewed the same way:
//     first the ramp that lets band 0 run ahead (launch q on the bands 0 ... nb-1-q, one whole-wave
//     launch each), then band by band along the anti-diagonals, so band 0 is final nb-1 diagonals before
//     the last band, and every band is copied to the host as soon as it is final, while the bands below
//     it are still being updated.  (A solve that stops on a
//     passed check cannot do this: the result is only known once the check has been read.)
// The kernels are indifferent to how sweeps are cut into launches, chunks or row ranges, so the result
// is bit-identical to the plain path; convergence sums taken band-wise are the band sums folded in band
// order (the plain path folds CTA partials in CTA order: the last bits of last_diff can differ).
// If iteration 0 already meets the threshold (the reference returns after one update) the caller redoes
// the solve on the plain path.
static std::atomic<long long> g_pipelined_uploads{0};

struct BandPlan {
    int nb = 0;                 // row bands
    std::vector<int> r0;        // band b = rows [r0[b], r0[b + 1])
    long long extra = 0;        // head start: sweeps every band is ahead of iteration 0 when it is over
};

struct BandLaunch { int k; bool norm; };   // sweeps fused, with the convergence sums

static bool plan_bands(const hmi_solver *s, BandPlan *pl)
{
    const hmi_params &p = s->p;
    const char *env = getenv("HMI_UPLOAD_PIPELINE");      // 0 = off, 2 = also on small grids (tests)
    const int mode = env ? atoi(env) : 1;
    if (getenv("HMI_DEBUG"))
        fprintf(stderr, "[hmi] plan_bands: mode %d stream %d nl %d tb %d radius %d rows %d T %lld maxit %lld\n", mode,
                (int)s->use_stream, (int)s->use_nl, s->tb, s->radius, p.rows, (long long)p.term_check_iters, (long long)p.max_iters);
    if (mode == 0 || p.ghost_top || p.ghost_bottom || !(s->use_stream || s->use_nl)) return false;
    if (mode != 2 && ((long long)p.rows * p.cols < (1LL << 26) || p.rows < 4096)) return false;
    const long long first_window = p.term_check_iters < p.max_iters - 1 ? p.term_check_iters : p.max_iters - 1;
    // bands of 512 rows, at most 64.  The band-by-band launches are partial waves (~1.35x their share of a whole-grid
    // launch), but they run under the copies; what they hide is the triangle nb/2 launches deep, which should
    // match the copy time (32768^2 fp64: 175 ms up / 156 ms down against 4.5 ms per launch).  Measured on the bench
    // workload, end to end as a share of the resident rate: no overlap 82.0 %, 32 bands 88.4 %, 64 bands 90.3 %
    // (profiles/r05_tune_pipeline_bands_merged_ramps.log; before the ramps ran as whole-wave launches 87.6 / 87.3 %)
    int nb = mode == 2 ? 8 : p.rows / 512;
    if (nb > 64) nb = 64;
    if (const char *eb = getenv("HMI_PIPELINE_BANDS")) nb = atoi(eb) < nb ? atoi(eb) : nb;    // tuning
    // the second request of the loop (T sweeps up to the next check, or max_iters - 1) must cover the head start,
    // and a band must be taller than the halo a launch reads from it
    while (nb >= 4 && ((long long)s->tb * (nb - 1) + 1 > first_window || p.rows / nb < 2 * s->radius * s->tb + 8)) nb /= 2;
    if (nb < 4) return false;
    if (getenv("HMI_DEBUG"))
        fprintf(stderr, "[hmi] one-shot copies pipelined: %d bands of ~%d rows, head start %lld sweeps\n", nb, p.rows / nb,
                (long long)s->tb * (nb - 1) + 1);
    pl->nb = nb;
    pl->extra = (long long)s->tb * (nb - 1);
    pl->r0.resize(nb + 1);
    for (int b = 0; b <= nb; ++b) pl->r0[b] = (int)((long long)p.rows * b / nb) & ~7;
    pl->r0[nb] = p.rows;
    return true;
}

// ---- the order of the launches, as pure functions (tests/test_host_api_cpu.py checks every read-after-write and
// write-after-read dependence of these orders through hmi_debug_band_steps) ----------------------------------
// A step = launch q on the contiguous bands [j_lo, j_hi], all of which are at time level q: ONE kernel launch.
struct BandStep { int q, j_lo, j_hi, final_band; };     // final_band: the band that is final after this step, or -1

// Head start.  landed = 1 ... nb-1: band `landed` is in HBM now — anti-diagonal landed-1, bottom band first, band by
// band.  landed = nb: the catch-up after the last band — the bands nb-1-q ... nb-1 are all at time level q, so launch q
// covers them in one whole-wave launch.  Every band ends up having run launches 0 ... nb-1.
static void band_head_start_steps(int nb, int landed, std::vector<BandStep> &out)
{
    if (landed < nb) {
        const int d = landed - 1;
        for (int j = d; j >= 0; --j) out.push_back({d - j, j, j, -1});
    } else {
        for (int q = 0; q < nb; ++q) out.push_back({q, nb - 1 - q, nb - 1, -1});
    }
}

// Drain of M launches per band.  Ramp: launch q covers the bands 0 ... M-1-q (all at time level q, one whole-wave
// launch), after which band 0 is final; then anti-diagonal after anti-diagonal, bottom band first, band by band: band j
// is final after diagonal j + M - 1.
static void band_drain_steps(int nb, int M, std::vector<BandStep> &out)
{
    for (int q = 0; q < M; ++q) {
        const int j_hi = M - 1 - q < nb - 1 ? M - 1 - q : nb - 1;
        out.push_back({q, 0, j_hi, q == M - 1 ? 0 : -1});
    }
    for (int d = M; d <= nb - 1 + M - 1; ++d) {
        const int j_hi = d < nb - 1 ? d : nb - 1, j_lo = d - (M - 1) > 0 ? d - (M - 1) : 0;
        for (int j = j_hi; j >= j_lo; --j) out.push_back({d - j, j, j, d - j == M - 1 ? j : -1});
    }
}

// Queues the steps: every band runs seq[0], seq[1], ... seq[M-1]; launch q reads grid (base + q) & 1 and writes the
// other one.  Band sums of launches with `norm` are folded on the device in issue order.
struct BandSchedule {
    hmi_solver *s;
    const BandPlan &pl;
    std::vector<BandLaunch> seq;
    int base;
    cudaStream_t st;
    double *acc;
    bool acc_started = false;

    int issue(const BandStep &b)
    {
        s->cur = (base + b.q) & 1;
        const int rc = run_block(s, seq[b.q].k, pl.r0[b.j_lo], pl.r0[b.j_hi + 1], seq[b.q].norm, st, false);
        if (rc != HMI_OK) return rc;
        if (seq[b.q].norm) {
            accumulate_norm_kernel<<<1, 1, 0, st>>>(acc, s->d_result, acc_started ? 0 : 1);
            acc_started = true;
            if (cudaGetLastError() != cudaSuccess) return fail(HMI_ERR_CUDA, "norm accumulation launch failed");
        }
        return HMI_OK;
    }
    // after the last step: the current grid, and the folded sums where the result of a check is read
    int finish(bool with_norm)
    {
        s->cur = (base + (int)seq.size()) & 1;
        s->last_main = st;
        if (with_norm) CUDA_TRY(cudaMemcpyAsync(s->d_result, acc, 4 * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return HMI_OK;
    }
};

// Upload in bands + head start.  On return every band has run iteration 0 and `extra` more sweeps, *nm holds
// the convergence sums of iteration 0, and the caller's arrays have been read completely.
static int upload_pipelined(hmi_solver *s, const BandPlan &pl, const void *image, const uint8_t *mask, int fill,
                            double fill_value, hmi_norm *nm)
{
    const hmi_params &p = s->p;
    const int nb = pl.nb;
    CUDA_TRY(cudaSetDevice(p.device));
    int rc = hmi::sync_solver(s);
    if (rc != HMI_OK) return rc;
    cudaStream_t st = s->own, cs = nullptr;
    CUDA_TRY(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    const size_t wbytes = (size_t)p.cols * s->elem;
    s->cur = 0;
    s->has_problem = true;
    s->halo_state = 0;
    BandSchedule sch{s, pl, {}, 0, st, s->d_result + 8};
    sch.seq.push_back({1, true});                                  // iteration 0
    for (int q = 1; q < nb; ++q) sch.seq.push_back({s->tb, false});
    std::vector<BandStep> steps;
    cudaError_t e = cudaSuccess;
    for (int b = 0; b < nb && rc == HMI_OK; ++b) {
        const int rows_b = pl.r0[b + 1] - pl.r0[b];
        char *gdst = s->buf[0] + (size_t)pl.r0[b] * s->pitch * s->elem;
        uint8_t *mdst = s->mask + (size_t)pl.r0[b] * s->mpitch;
        // (synchronous on return: the band is in HBM before anything that reads it is queued)
        if ((e = hmi::copy_h2d_2d(p.device, gdst, (size_t)s->pitch * s->elem, (const char *)image + (size_t)pl.r0[b] * wbytes,
                                  wbytes, wbytes, rows_b, cs)) != cudaSuccess) break;
        if ((e = hmi::copy_h2d_2d(p.device, mdst, (size_t)s->mpitch, mask + (size_t)pl.r0[b] * p.cols, (size_t)p.cols,
                                  (size_t)p.cols, rows_b, cs)) != cudaSuccess) break;
        normalize_mask_kernel<<<148, 256, 0, st>>>(mdst, (size_t)rows_b * s->mpitch);
        if (fill) {
            const dim3 grid((p.cols + 255) / 256, rows_b < 256 ? rows_b : 256);
            if (p.dtype == HMI_F64)
                fill_unknown_kernel<double><<<grid, 256, 0, st>>>((double *)gdst, s->pitch, mdst, s->mpitch, rows_b, p.cols, fill_value);
            else
                fill_unknown_kernel<float><<<grid, 256, 0, st>>>((float *)gdst, s->pitch, mdst, s->mpitch, rows_b, p.cols, (float)fill_value);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if (b >= 1) {                              // band b has landed: bands b-1 ... 0 advance by one launch
            steps.clear();
            band_head_start_steps(nb, b, steps);
            for (size_t i = 0; i < steps.size() && rc == HMI_OK; ++i) rc = sch.issue(steps[i]);
        }
    }
    // everything is in HBM: the bands catch up with band 0 in whole-wave launches (band launches are partial waves
    // and cost ~1.35x their share)
    steps.clear();
    band_head_start_steps(nb, nb, steps);
    for (size_t i = 0; i < steps.size() && rc == HMI_OK && e == cudaSuccess; ++i) rc = sch.issue(steps[i]);
    if (rc == HMI_OK && e == cudaSuccess) rc = sch.finish(true);
    if (cs) { cudaStreamSynchronize(cs); cudaStreamDestroy(cs); }
    if (rc != HMI_OK) return rc;
    if (e != cudaSuccess) return fail(HMI_ERR_CUDA, "pipelined upload failed: %s", cudaGetErrorString(e));
    return fetch_norm(s, st, nm);
}

// The last request of the loop (n sweeps, the last one with the convergence sums if `norm`) with the download
// overlapped: whole-grid launches first, then the last launches band-skewed, every band copied out once final.
static int drain_pipelined(hmi_solver *s, const BandPlan &pl, long long n, bool norm, hmi_norm *nm, void *out)
{
    const hmi_params &p = s->p;
    const int nb = pl.nb;
    cudaStream_t st = s->own;
    // the skewed tail: nb launches per band, the last of them the single sweep that carries the sums
    std::vector<BandLaunch> seq;
    long long tail = 0;
    for (int q = 0; q < nb; ++q) {
        const bool last = q == nb - 1;
        seq.push_back({norm && last ? 1 : s->tb, norm && last});
        tail += seq.back().k;
    }
    if (n < tail) {                                // short request: a shorter tail (at least the band of the sums)
        seq.clear();
        tail = 0;
        long long left = norm ? n - 1 : n;         // (< tb * (nb - 1) with the sums, < tb * nb without)
        while (left > 0) {
            const int k = left < s->tb ? (int)left : s->tb;
            seq.push_back({k, false});
            left -= k;
        }
        if (norm) seq.push_back({1, true});
        for (auto &l : seq) tail += l.k;
        if (seq.empty()) return fail(HMI_ERR_STATE, "internal: empty drain");
    }
    int rc = do_sweeps(s, n - tail, false, st);
    if (rc != HMI_OK) return rc;
    BandSchedule sch{s, pl, seq, s->cur, st, s->d_result + 8};
    const int final_cur = (s->cur + (int)seq.size()) & 1;
    std::vector<cudaEvent_t> ev(nb, nullptr);
    cudaError_t e = cudaSuccess;
    for (int b = 0; b < nb && e == cudaSuccess; ++b) e = cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming);
    // The copies run on their own host thread: the launches below fill the driver's queue, so the thread that
    // issues them only returns when most of them have already executed.
    std::atomic<int> recorded{0};
    std::atomic<bool> give_up{false};
    cudaError_t e_dl = cudaSuccess;
    const size_t wbytes = (size_t)p.cols * s->elem;
    std::thread downloader;
    if (e == cudaSuccess)
        downloader = std::thread([&]() {
            cudaStream_t cs = nullptr;
            if ((e_dl = cudaSetDevice(p.device)) != cudaSuccess) return;
            if ((e_dl = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking)) != cudaSuccess) return;
            for (int b = 0; b < nb && e_dl == cudaSuccess; ++b) {
                while (recorded.load(std::memory_order_acquire) <= b && !give_up.load()) std::this_thread::sleep_for(std::chrono::microseconds(20));
                if (give_up.load()) break;
                if ((e_dl = cudaEventSynchronize(ev[b])) != cudaSuccess) break;     // band b is final
                const int rows_b = pl.r0[b + 1] - pl.r0[b];
                const char *src = s->buf[final_cur] + (size_t)pl.r0[b] * s->pitch * s->elem;
                e_dl = hmi::copy_d2h_2d(p.device, (char *)out + (size_t)pl.r0[b] * wbytes, wbytes, src, (size_t)s->pitch * s->elem,
                                        wbytes, rows_b, cs);
            }
            cudaStreamSynchronize(cs);
            cudaStreamDestroy(cs);
        });
    // ramp (band 0 runs ahead, whole-wave launches), then the bands below finish one anti-diagonal after the other
    // while the copies run; a band that is final is handed to the downloader at once
    std::vector<BandStep> steps;
    band_drain_steps(nb, (int)seq.size(), steps);
    for (size_t i = 0; i < steps.size() && rc == HMI_OK && e == cudaSuccess; ++i) {
        rc = sch.issue(steps[i]);
        const int fb = steps[i].final_band;
        if (rc == HMI_OK && fb >= 0) {             // (bands become final in order 0, 1, 2 ...)
            e = cudaEventRecord(ev[fb], st);
            if (e == cudaSuccess) recorded.store(fb + 1, std::memory_order_release);
        }
    }
    if (rc == HMI_OK && e == cudaSuccess) rc = sch.finish(norm);
    if (rc != HMI_OK || e != cudaSuccess) give_up.store(true);
    if (downloader.joinable()) downloader.join();
    for (auto &x : ev) if (x) cudaEventDestroy(x);
    if (rc != HMI_OK) return rc;
    if (e != cudaSuccess || e_dl != cudaSuccess)
        return fail(HMI_ERR_CUDA, "pipelined download failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e_dl));
    if (norm) return fetch_norm(s, st, nm);
    CUDA_TRY(cudaStreamSynchronize(st));
    return HMI_OK;
}

static int inpaint_host_impl(const hmi_params *params, const void *image, const uint8_t *mask, int fill,
                             double fill_value, void *out, hmi_status *status)
{
    if (!params || !image || !mask || !out || !status) return fail(HMI_ERR_BAD_ARG, "null argument");
    hmi_solver *s = nullptr;
    int rc = hmi_create(params, &s);
    if (rc != HMI_OK) return rc;
    const size_t elem = params->dtype == HMI_F64 ? 8 : 4;
    std::thread host_init;
    auto start_host_init = [&]() {                 // the caller's array has been read: initialise it in place
        if (fill == 2 && !host_init.joinable())
            host_init = std::thread(hmi_host_fill_unknown, const_cast<void *>(image), mask,
                                    (int64_t)params->rows * params->cols, params->dtype, fill_value);
    };
    BandPlan pl;
    bool plain = !plan_bands(s, &pl);
    bool downloaded = false;
    if (!plain) {
        bool first = true;
        long long ahead = 0, done = 0;
        rc = hmi::reference_loop(s->p, [&](long long n, bool norm, hmi_norm *nm) {
            int r;
            if (first) {                           // iteration 0 (n == 1, with the sums) + the head start
                first = false;
                r = upload_pipelined(s, pl, image, mask, fill, fill_value, nm);
                if (r == HMI_OK) g_pipelined_uploads.fetch_add(1);
                start_host_init();
                ahead = pl.extra;
                done = 1;
                return r;
            }
            const long long m = n - ahead;         // >= 1 by plan_bands
            ahead = 0;
            done += n;
            if (done == s->p.max_iters) {          // the last request, whatever its check says: overlap the download
                r = drain_pipelined(s, pl, m, norm, nm, out);
                downloaded = r == HMI_OK;
                return r;
            }
            r = do_sweeps(s, m, norm, s->own);
            if (r == HMI_OK && norm) r = fetch_norm(s, s->own, nm);
            return r;
        }, status);
        if (rc == HMI_OK) rc = cudaStreamSynchronize(s->own) == cudaSuccess ? HMI_OK : fail(HMI_ERR_CUDA, "solve failed");
        // converged at iteration 0: the reference returns the grid after ONE update, the bands are further on
        if (rc == HMI_OK && status->converged && status->updates_done == 1) plain = true;
        if (plain && host_init.joinable()) host_init.join();
    }
    if (plain && rc == HMI_OK) {
        rc = hmi_set_problem_host(s, image, (size_t)params->cols * elem, mask, (size_t)params->cols);
        if (rc == HMI_OK) start_host_init();
        if (rc == HMI_OK && fill) rc = hmi::fill_unknown(s, fill_value);
        if (rc == HMI_OK) rc = hmi_run(s, s->own, status);
    }
    if (rc == HMI_OK && !downloaded) rc = hmi_get_result_host(s, out, (size_t)params->cols * elem);
    if (host_init.joinable()) host_init.join();
    hmi_destroy(s);
    return rc;
}

int hmi_inpaint_host(const hmi_params *params, const void *image, const uint8_t *mask, void *out,
                     hmi_status *status)
{
    return inpaint_host_impl(params, image, mask, 0, 0.0, out, status);
}

int hmi_inpaint_host_fill(const hmi_params *params, const void *image, const uint8_t *mask, double fill_value,
                          void *out, hmi_status *status)
{
    return inpaint_host_impl(params, image, mask, 1, fill_value, out, status);
}

int hmi_inpaint_host_init(const hmi_params *params, void *image, const uint8_t *mask, double fill_value, void *out,
                          hmi_status *status)
{
    return inpaint_host_impl(params, image, mask, 2, fill_value, out, status);
}

int hmi_init_nearest_host(void *image, size_t image_pitch_bytes, const uint8_t *mask, size_t mask_pitch_bytes,
                          int32_t rows, int32_t cols, int32_t dtype, int32_t device)
{
    if (!image || !mask) return fail(HMI_ERR_BAD_ARG, "null argument");
    if (rows < 1 || cols < 1) return fail(HMI_ERR_BAD_ARG, "empty grid");
    if (dtype != HMI_F32 && dtype != HMI_F64) return fail(HMI_ERR_BAD_ARG, "bad dtype");
    const size_t elem = dtype == HMI_F64 ? 8 : 4;
    const size_t wbytes = (size_t)cols * elem;
    if (image_pitch_bytes < wbytes || mask_pitch_bytes < (size_t)cols)
        return fail(HMI_ERR_BAD_ARG, "pitch smaller than a row");
    int ndev = 0;
    cudaError_t e = hmi::device_count_patiently(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(HMI_ERR_NO_DEVICE, "no CUDA device available; this library has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(HMI_ERR_BAD_ARG, "device %d out of range (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    const long long pitch = round_up(cols, 32), mpitch = round_up(cols, 128);
    const size_t gbytes = (size_t)rows * pitch * elem, mbytes = (size_t)rows * mpitch;
    const size_t obytes = (size_t)rows * pitch * sizeof(int);
    char *d_img = nullptr, *d_out = nullptr;
    uint8_t *d_mask = nullptr;
    int *d_off = nullptr, *d_status = nullptr;
    int rc = HMI_OK, status = 0;
    do {
        if ((e = cached_malloc(device, (void **)&d_img, gbytes)) != cudaSuccess) break;
        if ((e = cached_malloc(device, (void **)&d_out, gbytes)) != cudaSuccess) break;
        if ((e = cached_malloc(device, (void **)&d_mask, mbytes)) != cudaSuccess) break;
        if ((e = cached_malloc(device, (void **)&d_off, obytes)) != cudaSuccess) break;
        if ((e = cudaMalloc((void **)&d_status, sizeof(int))) != cudaSuccess) break;
        if ((e = hmi::copy_h2d_2d(device, d_img, (size_t)pitch * elem, image, image_pitch_bytes, wbytes, rows)) !=
            cudaSuccess) break;
        if ((e = hmi::copy_h2d_2d(device, d_mask, (size_t)mpitch, mask, mask_pitch_bytes, (size_t)cols, rows)) !=
            cudaSuccess) break;
        if (dtype == HMI_F64)
            e = hmi::launch_init_nearest<double>((const double *)d_img, pitch, d_mask, mpitch, rows, cols, d_off,
                                                 pitch, (double *)d_out, pitch, d_status, nullptr);
        else
            e = hmi::launch_init_nearest<float>((const float *)d_img, pitch, d_mask, mpitch, rows, cols, d_off,
                                                pitch, (float *)d_out, pitch, d_status, nullptr);
        if (e != cudaSuccess) break;
        if ((e = cudaMemcpy(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if (status != 0) break;
        e = hmi::copy_d2h_2d(device, image, image_pitch_bytes, d_out, (size_t)pitch * elem, wbytes, rows);
    } while (0);
    if (e != cudaSuccess) {
        rc = fail(HMI_ERR_CUDA, "nearest initialiser failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
    } else if (status != 0) {
        rc = fail(HMI_ERR_BAD_ARG, "the nearest initialiser needs at least one known cell");
    }
    cudaDeviceSynchronize();
    cached_free(device, d_img, gbytes);
    cached_free(device, d_out, gbytes);
    cached_free(device, d_mask, mbytes);
    cached_free(device, d_off, obytes);
    if (d_status) cudaFree(d_status);
    return rc;
}

int hmi_set_problem_host_blend(hmi_solver *fine, const void *image, size_t image_pitch_bytes, const uint8_t *mask,
                               size_t mask_pitch_bytes, hmi_solver *coarse, const int32_t *sx, const int32_t *sx1,
                               const void *ax0, const void *ax1, const int32_t *sy, const int32_t *sy1,
                               const void *by0, const void *by1, int32_t scrub_nan, hmi_blend_info *info)
{
    if (!fine || !coarse || !sx || !sx1 || !ax0 || !ax1 || !sy || !sy1 || !by0 || !by1 || !info)
        return fail(HMI_ERR_BAD_ARG, "null argument");
    if (fine == coarse) return fail(HMI_ERR_BAD_ARG, "fine and coarse solver must differ");
    if (fine->p.dtype != coarse->p.dtype || fine->p.device != coarse->p.device)
        return fail(HMI_ERR_BAD_ARG, "fine and coarse level must share dtype and device");
    if (fine->p.ghost_top || fine->p.ghost_bottom || coarse->p.ghost_top || coarse->p.ghost_bottom)
        return fail(HMI_ERR_STATE, "the multigrid blend works on whole grids, not slabs");
    if (!coarse->has_problem) return fail(HMI_ERR_STATE, "the coarse level has no solution yet");
    int rc = hmi_set_problem_host(fine, image, image_pitch_bytes, mask, mask_pitch_bytes);
    if (rc != HMI_OK) return rc;
    const int rows = fine->p.rows, cols = fine->p.cols;
    const size_t elem = fine->elem;
    // tables: [sx | sx1 | sy | sy1] ints, then [ax0 | ax1 | by0 | by1] values, then partials and the NaN flag
    const size_t n_int = 2 * (size_t)cols + 2 * (size_t)rows;
    const size_t off_val = (n_int * sizeof(int) + 15) / 16 * 16;
    const size_t n_val = n_int;
    const size_t off_part = (off_val + n_val * elem + 15) / 16 * 16;
    const int max_blocks = hmi::upsample_max_blocks(cols);
    const size_t off_flag = off_part + (size_t)max_blocks * 3 * sizeof(double);
    const size_t total = off_flag + 16;
    char *d = nullptr;
    CUDA_TRY(cudaMalloc((void **)&d, total));
    std::vector<char> h(off_part);
    int *hi_ = reinterpret_cast<int *>(h.data());
    memcpy(hi_, sx, cols * sizeof(int));
    memcpy(hi_ + cols, sx1, cols * sizeof(int));
    memcpy(hi_ + 2 * cols, sy, rows * sizeof(int));
    memcpy(hi_ + 2 * cols + rows, sy1, rows * sizeof(int));
    char *hv = h.data() + off_val;
    memcpy(hv, ax0, cols * elem);
    memcpy(hv + cols * elem, ax1, cols * elem);
    memcpy(hv + 2 * cols * elem, by0, rows * elem);
    memcpy(hv + (2 * cols + rows) * elem, by1, rows * elem);
    cudaError_t e = cudaMemcpy(d, h.data(), off_part, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d + off_flag, 0, 16);
    int nblocks = 0;
    if (e == cudaSuccess) {
        const int *di = reinterpret_cast<const int *>(d);
        const char *dv = d + off_val;
        double *dpart = reinterpret_cast<double *>(d + off_part);
        int *dflag = reinterpret_cast<int *>(d + off_flag);
        if (fine->p.dtype == HMI_F64)
            e = hmi::launch_upsample_blend<double>(
                (double *)fine->buf[0], fine->pitch, fine->mask, fine->mpitch, rows, cols,
                (const double *)coarse->buf[coarse->cur], coarse->pitch, di, di + cols, (const double *)dv,
                (const double *)dv + cols, di + 2 * cols, di + 2 * cols + rows, (const double *)dv + 2 * cols,
                (const double *)dv + 2 * cols + rows, scrub_nan, dpart, dflag, &nblocks, nullptr);
        else
            e = hmi::launch_upsample_blend<float>(
                (float *)fine->buf[0], fine->pitch, fine->mask, fine->mpitch, rows, cols,
                (const float *)coarse->buf[coarse->cur], coarse->pitch, di, di + cols, (const float *)dv,
                (const float *)dv + cols, di + 2 * cols, di + 2 * cols + rows, (const float *)dv + 2 * cols,
                (const float *)dv + 2 * cols + rows, scrub_nan, dpart, dflag, &nblocks, nullptr);
    }
    std::vector<double> part;
    int flag = 0;
    if (e == cudaSuccess) {
        part.resize((size_t)nblocks * 3);
        e = cudaMemcpy(part.data(), d + off_part, part.size() * sizeof(double), cudaMemcpyDeviceToHost);
    }
    if (e == cudaSuccess) e = cudaMemcpy(&flag, d + off_flag, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(HMI_ERR_CUDA, "multigrid blend failed: %s", cudaGetErrorString(e));
    }
    double lo = INFINITY, hi = -INFINITY, nanv = 0.0;
    for (int b = 0; b < nblocks; ++b) {
        lo = part[3 * b] < lo ? part[3 * b] : lo;
        hi = part[3 * b + 1] > hi ? part[3 * b + 1] : hi;
        nanv = part[3 * b + 2] > nanv ? part[3 * b + 2] : nanv;
    }
    info->min = nanv != 0.0 ? nan("") : lo;      // np.min / np.max propagate NaN
    info->max = nanv != 0.0 ? nan("") : hi;
    info->has_nan = nanv != 0.0 ? 1 : 0;
    info->image_had_nan = flag;
    return HMI_OK;
}

int hmi_current_grid(hmi_solver *s, void **grid_dev, size_t *pitch_bytes)
{
    if (!s || !grid_dev || !pitch_bytes) return fail(HMI_ERR_BAD_ARG, "null argument");
    *grid_dev = s->buf[s->cur];
    *pitch_bytes = (size_t)s->pitch * s->elem;
    return HMI_OK;
}

int hmi_mask_grid(hmi_solver *s, uint8_t **mask_dev, size_t *pitch_bytes)
{
    if (!s || !mask_dev || !pitch_bytes) return fail(HMI_ERR_BAD_ARG, "null argument");
    *mask_dev = s->mask;
    *pitch_bytes = (size_t)s->mpitch;
    return HMI_OK;
}

int64_t hmi_upload_pipeline_count(void) { return (int64_t)g_pipelined_uploads.load(); }

int hmi_debug_band_steps(int32_t nb, int32_t m, int32_t phase, int32_t *out, int32_t cap)
{
    if (nb < 1 || m < 1 || !out || cap < 0) return fail(HMI_ERR_BAD_ARG, "bad argument");
    std::vector<BandStep> steps;
    if (phase == 0) {
        steps.push_back({-1, 0, 0, -1});                       // band 0 lands
        for (int b = 1; b < nb; ++b) {
            steps.push_back({-1, b, b, -1});                   // band b lands ...
            band_head_start_steps(nb, b, steps);               // ... and the bands above it advance
        }
        band_head_start_steps(nb, nb, steps);
    } else {
        band_drain_steps(nb, m, steps);
    }
    if ((long long)steps.size() > cap) return fail(HMI_ERR_BAD_ARG, "%zu steps do not fit %d", steps.size(), cap);
    for (size_t i = 0; i < steps.size(); ++i) {
        out[4 * i + 0] = steps[i].q; out[4 * i + 1] = steps[i].j_lo;
        out[4 * i + 2] = steps[i].j_hi; out[4 * i + 3] = steps[i].final_band;
    }
    return (int)steps.size();
}

int hmi_debug_pick_chunks(int32_t nrows, int32_t nwin, int32_t n_edge, int32_t halo, int32_t wpb, int32_t cta_slots,
                          double edge_cost, int32_t min_chunk, int32_t *chunk_rows, int32_t *chunk_rows_edge)
{
    if (nrows < 1 || nwin < 1 || wpb < 1 || cta_slots < 1 || !chunk_rows || !chunk_rows_edge)
        return fail(HMI_ERR_BAD_ARG, "bad argument");
    int ch = 0, che = 0;
    hmi::stream_pick_chunks(nrows, nwin, n_edge, halo, wpb, cta_slots, edge_cost, &ch, &che, min_chunk);
    *chunk_rows = ch;
    *chunk_rows_edge = che;
    return HMI_OK;
}

int hmi_release_cache(void)
{
    std::vector<hmi::CachedBlock> drop;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        drop.swap(hmi::g_cache);
        hmi::g_cache_bytes = 0;
    }
    hmi::drop_blocks(drop);
    return HMI_OK;
}

int hmi_set_term_thres(hmi_solver *s, double term_thres)
{
    if (!s) return fail(HMI_ERR_BAD_ARG, "null solver");
    if (!(term_thres > 0.0) && term_thres == term_thres)     // NaN is allowed: it never terminates, as in the reference
        return fail(HMI_ERR_BAD_ARG, "term_thres must be larger than zero");
    s->p.term_thres = term_thres;
    return HMI_OK;
}

int hmi_temporal_block(const hmi_solver *s) { return s ? s->tb : 0; }
int64_t hmi_launch_count(const hmi_solver *s) { return s ? s->launches : 0; }
const char *hmi_kernel_name(const hmi_solver *s)
{
    return !s ? "" : (s->use_nl ? "stream_nl" : (s->use_stream ? "stream" : "generic"));
}

}  // extern "C"
