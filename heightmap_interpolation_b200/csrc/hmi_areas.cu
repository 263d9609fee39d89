// The gridded branch of the reference CLI's interpolate() with the grid resident on the device
// (include/hmi_b200.h: hmi_areas_*).
//
// Reference (apps/interpolate_netcdf4.py:176-212, same code in apps/interpolate_xyz.py:255-279), per
// work area, on the host: bounding-box crop of elevation and mask (two copies), `~mask | ~area`,
// isnan -> 0, the inpainter's initialiser, the solve, and a boolean-indexed paste of the inpainted cells
// into elevation_int — seven full passes over the window in NumPy around the solve.  Here the elevation
// grid and mask_int are uploaded once, elevation_int is a device copy, and per area one kernel builds
// the window problem straight into the solver's slabs (crop + NaN -> 0 + known mask, with the sums the
// 'mean' initialiser and the z-range of 'absolute_percent' need), the solve runs on the device group,
// and one kernel pastes the inpainted cells into elevation_int, which is downloaded once at the end.
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "hmi_internal.cuh"
#include "hmi_xfer.cuh"

struct hmi_areas {
    int device = 0, rows = 0, cols = 0, dtype = HMI_F64;
    size_t elem = 8;
    long long pitch = 0, mpitch = 0;     // elements / bytes per row
    char *elev = nullptr, *elev_int = nullptr;
    uint8_t *mask_int = nullptr, *area = nullptr;
    size_t gbytes = 0, mbytes = 0;
    cudaStream_t st = nullptr;
    bool area_loaded = false;
};

using hmi::fail;

namespace {

constexpr int kPrepBlocksY = 64;

// One slab of the window problem.  Row p of the slab (ghost rows included) is window row
// first_row + p, i.e. grid row r0 + first_row + p.  known = not flagged for interpolation, or outside
// the work area (interpolate_netcdf4.py:187-190); NaN -> 0 everywhere in the window (:192).
template <typename T>
__global__ void window_prepare_kernel(T *__restrict__ f, long long fpitch, uint8_t *__restrict__ m, long long fmpitch,
                                      int phys_rows, int wcols, int first_row, int interior_lo, int interior_hi,
                                      const T *__restrict__ elev, long long pitch, const uint8_t *__restrict__ mask_int,
                                      const uint8_t *__restrict__ area, long long mpitch, int r0, int c0,
                                      double *__restrict__ partials)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double sum = 0.0, cnt = 0.0, lo = INFINITY, hi = -INFINITY, nanv = 0.0;
    if (j < wcols) {
        for (int p = blockIdx.y; p < phys_rows; p += gridDim.y) {
            const long long gr = (long long)r0 + first_row + p, gc = (long long)c0 + j;
            T v = elev[gr * pitch + gc];
            if (v != v) v = T(0);
            const bool known = !mask_int[gr * mpitch + gc] || (area && !area[gr * mpitch + gc]);
            f[(long long)p * fpitch + j] = v;
            m[(long long)p * fmpitch + j] = known ? 1 : 0;
            if (known && p >= interior_lo && p < interior_hi) {      // statistics over the slab's own rows only
                const double d = (double)v;
                sum += d; cnt += 1.0;
                lo = d < lo ? d : lo;
                hi = d > hi ? d : hi;
            }
        }
    }
    __shared__ double sh[5][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { sh[0][warp] = sum; sh[1][warp] = cnt; sh[2][warp] = lo; sh[3][warp] = hi; sh[4][warp] = nanv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 1; w < nw; ++w) {
            sum += sh[0][w]; cnt += sh[1][w]; lo = fmin(lo, sh[2][w]); hi = fmax(hi, sh[3][w]);
        }
        const long long b = (long long)blockIdx.y * gridDim.x + blockIdx.x;
        partials[4 * b + 0] = sum; partials[4 * b + 1] = cnt; partials[4 * b + 2] = lo; partials[4 * b + 3] = hi;
    }
}

// elevation_int[window][unknown] = result[unknown]   (interpolate_netcdf4.py:210-212)
template <typename T>
__global__ void window_paste_kernel(const T *__restrict__ res, long long rpitch, const uint8_t *__restrict__ m, long long rmpitch,
                                    int nrows, int wcols, int first_row, T *__restrict__ elev_int, long long pitch, int r0, int c0)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= wcols) return;
    for (int i = blockIdx.y; i < nrows; i += gridDim.y)
        if (!m[(long long)i * rmpitch + j])
            elev_int[((long long)r0 + first_row + i) * pitch + c0 + j] = res[(long long)i * rpitch + j];
}

template <typename T>
int prepare_slab(hmi_areas *a, hmi_solver *s, int first_row, int r0, int c0, std::vector<double> &part)
{
    const int wcols = s->p.cols;
    const dim3 grid((wcols + 127) / 128, s->phys_rows < kPrepBlocksY ? s->phys_rows : kPrepBlocksY);
    const size_t nblk = (size_t)grid.x * grid.y;
    double *d_part = nullptr;
    CUDA_TRY(cudaMalloc((void **)&d_part, nblk * 4 * sizeof(double)));
    s->cur = 0;
    window_prepare_kernel<T><<<grid, 128, 0, s->own>>>(
        (T *)s->buf[0], s->pitch, s->mask, s->mpitch, s->phys_rows, wcols, first_row, s->p.ghost_top,
        s->p.ghost_top + s->p.rows, (const T *)a->elev, a->pitch, a->mask_int, a->area_loaded ? a->area : nullptr, a->mpitch,
        r0, c0, d_part);
    cudaError_t e = cudaGetLastError();
    part.resize(nblk * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(part.data(), d_part, part.size() * sizeof(double), cudaMemcpyDeviceToHost, s->own);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->own);
    cudaFree(d_part);
    if (e != cudaSuccess) return fail(HMI_ERR_CUDA, "window kernel failed: %s", cudaGetErrorString(e));
    s->has_problem = true;
    s->halo_state = 0;
    s->last_main = s->own;
    return HMI_OK;
}

}  // namespace

extern "C" {

void hmi_areas_destroy(hmi_areas *a)
{
    if (!a) return;
    cudaSetDevice(a->device);
    if (a->st) { cudaStreamSynchronize(a->st); cudaStreamDestroy(a->st); }
    hmi::cached_free(a->device, a->elev, a->gbytes);
    hmi::cached_free(a->device, a->elev_int, a->gbytes);
    hmi::cached_free(a->device, a->mask_int, a->mbytes);
    hmi::cached_free(a->device, a->area, a->mbytes);
    cudaGetLastError();
    delete a;
}

int hmi_areas_create(int32_t rows, int32_t cols, int32_t dtype, int32_t device, const void *elevation,
                     size_t elevation_pitch_bytes, const uint8_t *mask_int, size_t mask_pitch_bytes, hmi_areas **out)
{
    if (!elevation || !mask_int || !out) return fail(HMI_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    if (rows < 1 || cols < 1) return fail(HMI_ERR_BAD_ARG, "empty grid");
    if (dtype != HMI_F32 && dtype != HMI_F64) return fail(HMI_ERR_BAD_ARG, "bad dtype");
    const size_t elem = dtype == HMI_F64 ? 8 : 4;
    if (elevation_pitch_bytes < (size_t)cols * elem || mask_pitch_bytes < (size_t)cols)
        return fail(HMI_ERR_BAD_ARG, "pitch smaller than a row");
    int ndev = 0;
    const cudaError_t e0 = hmi::device_count_patiently(&ndev);
    if (e0 != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(HMI_ERR_NO_DEVICE, "no CUDA device available; this library has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(HMI_ERR_BAD_ARG, "device %d out of range (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    hmi_areas *a = new (std::nothrow) hmi_areas();
    if (!a) return fail(HMI_ERR_CUDA, "out of host memory");
    a->device = device; a->rows = rows; a->cols = cols; a->dtype = dtype; a->elem = elem;
    a->pitch = hmi::round_up(cols, 32);
    a->mpitch = hmi::round_up(cols, 128);
    a->gbytes = (size_t)rows * a->pitch * elem;
    a->mbytes = (size_t)rows * a->mpitch;
    cudaError_t e = cudaStreamCreateWithFlags(&a->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = hmi::cached_malloc(device, (void **)&a->elev, a->gbytes);
    if (e == cudaSuccess) e = hmi::cached_malloc(device, (void **)&a->elev_int, a->gbytes);
    if (e == cudaSuccess) e = hmi::cached_malloc(device, (void **)&a->mask_int, a->mbytes);
    if (e == cudaSuccess) e = hmi::cached_malloc(device, (void **)&a->area, a->mbytes);
    if (e == cudaSuccess)
        e = hmi::copy_h2d_2d(device, a->elev, (size_t)a->pitch * elem, elevation, elevation_pitch_bytes, (size_t)cols * elem, rows, a->st);
    if (e == cudaSuccess) e = hmi::copy_h2d_2d(device, a->mask_int, (size_t)a->mpitch, mask_int, mask_pitch_bytes, (size_t)cols, rows, a->st);
    // elevation_int starts as a copy of elevation (interpolate_netcdf4.py:56)
    if (e == cudaSuccess) e = cudaMemcpyAsync(a->elev_int, a->elev, a->gbytes, cudaMemcpyDeviceToDevice, a->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(a->st);
    if (e != cudaSuccess) {
        const int rc = fail(HMI_ERR_CUDA, "loading the grid failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        hmi_areas_destroy(a);
        return rc;
    }
    *out = a;
    return HMI_OK;
}

int hmi_areas_set_work_area(hmi_areas *a, const uint8_t *work_area, size_t pitch_bytes)
{
    if (!a) return fail(HMI_ERR_BAD_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(a->device));
    if (!work_area) { a->area_loaded = false; return HMI_OK; }      // one area covering the grid
    if (pitch_bytes < (size_t)a->cols) return fail(HMI_ERR_BAD_ARG, "pitch smaller than a row");
    CUDA_TRY(hmi::copy_h2d_2d(a->device, a->area, (size_t)a->mpitch, work_area, pitch_bytes, (size_t)a->cols, a->rows, a->st));
    a->area_loaded = true;
    return HMI_OK;
}

int hmi_areas_load_window(hmi_areas *a, hmi_group *g, int32_t r0, int32_t c0, hmi_window_info *info)
{
    if (!a || !g || !info) return fail(HMI_ERR_BAD_ARG, "null argument");
    if (g->p.dtype != a->dtype) return fail(HMI_ERR_BAD_ARG, "the solver group and the grid differ in dtype");
    if (r0 < 0 || c0 < 0 || r0 + g->p.rows > a->rows || c0 + g->p.cols > a->cols)
        return fail(HMI_ERR_BAD_ARG, "window [%d, %d) x [%d, %d) exceeds the %d x %d grid", r0, r0 + g->p.rows, c0,
                    c0 + g->p.cols, a->rows, a->cols);
    CUDA_TRY(cudaSetDevice(a->device));
    CUDA_TRY(cudaStreamSynchronize(a->st));
    double sum = 0.0, cnt = 0.0, lo = INFINITY, hi = -INFINITY;
    for (int i = 0; i < g->n; ++i) {
        hmi_solver *s = g->slab[i];
        if (s->p.device != a->device) {
            // slabs on other GPUs read the grid through peer memory
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, s->p.device, a->device));
            if (!can) return fail(HMI_ERR_UNSUPPORTED, "device %d cannot access device %d's memory", s->p.device, a->device);
            CUDA_TRY(cudaSetDevice(s->p.device));
            const cudaError_t e = cudaDeviceEnablePeerAccess(a->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(HMI_ERR_CUDA, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
        CUDA_TRY(cudaSetDevice(s->p.device));
        int rc = hmi::sync_solver(s);
        if (rc != HMI_OK) return rc;
        std::vector<double> part;
        const int first_row = g->r0[i] - s->p.ghost_top;
        rc = a->dtype == HMI_F64 ? prepare_slab<double>(a, s, first_row, r0, c0, part) : prepare_slab<float>(a, s, first_row, r0, c0, part);
        if (rc != HMI_OK) return rc;
        for (size_t b = 0; b + 3 < part.size(); b += 4) {
            sum += part[b]; cnt += part[b + 1];
            lo = part[b + 2] < lo ? part[b + 2] : lo;
            hi = part[b + 3] > hi ? part[b + 3] : hi;
        }
    }
    g->has_problem = true;
    info->known = (int64_t)cnt;
    info->unknown = (int64_t)g->p.rows * g->p.cols - (int64_t)cnt;
    info->known_sum = sum;
    info->known_min = lo;
    info->known_max = hi;
    return HMI_OK;
}

int hmi_areas_paste(hmi_areas *a, hmi_group *g, int32_t r0, int32_t c0)
{
    if (!a || !g) return fail(HMI_ERR_BAD_ARG, "null argument");
    if (!g->has_problem) return fail(HMI_ERR_STATE, "no window loaded");
    if (r0 < 0 || c0 < 0 || r0 + g->p.rows > a->rows || c0 + g->p.cols > a->cols) return fail(HMI_ERR_BAD_ARG, "window exceeds the grid");
    for (int i = 0; i < g->n; ++i) {
        hmi_solver *s = g->slab[i];
        const int rc = hmi::sync_solver(s);
        if (rc != HMI_OK) return rc;
        CUDA_TRY(cudaSetDevice(s->p.device));
        const dim3 grid((s->p.cols + 127) / 128, s->p.rows < kPrepBlocksY ? s->p.rows : kPrepBlocksY);
        const char *res = s->buf[s->cur] + (size_t)s->p.ghost_top * s->pitch * s->elem;
        const uint8_t *m = s->mask + (size_t)s->p.ghost_top * s->mpitch;
        if (a->dtype == HMI_F64)
            window_paste_kernel<double><<<grid, 128, 0, s->own>>>((const double *)res, s->pitch, m, s->mpitch, s->p.rows, s->p.cols,
                                                                  g->r0[i], (double *)a->elev_int, a->pitch, r0, c0);
        else
            window_paste_kernel<float><<<grid, 128, 0, s->own>>>((const float *)res, s->pitch, m, s->mpitch, s->p.rows, s->p.cols,
                                                                 g->r0[i], (float *)a->elev_int, a->pitch, r0, c0);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(s->own));
    }
    return HMI_OK;
}

int hmi_areas_get_result_host(hmi_areas *a, void *elevation_int, size_t pitch_bytes)
{
    if (!a || !elevation_int) return fail(HMI_ERR_BAD_ARG, "null argument");
    if (pitch_bytes < (size_t)a->cols * a->elem) return fail(HMI_ERR_BAD_ARG, "pitch smaller than a row");
    CUDA_TRY(cudaSetDevice(a->device));
    CUDA_TRY(hmi::copy_d2h_2d(a->device, elevation_int, pitch_bytes, a->elev_int, (size_t)a->pitch * a->elem, (size_t)a->cols * a->elem,
                              a->rows, a->st));
    return HMI_OK;
}

}  // extern "C"
