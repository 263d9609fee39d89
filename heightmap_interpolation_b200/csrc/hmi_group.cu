// Several GPUs of one box driven by ONE process (include/hmi_b200.h: hmi_group_*).
//
// The reference's one call site is a single process, inpainter.inpaint(image, mask)
// (apps/interpolate_netcdf4.py:200-204), so the drop-in has to be able to use the whole box without
// the caller starting one process per GPU.  A group cuts the grid into contiguous row slabs, one
// solver per device, links neighbouring slabs through peer memory (hmi_halo.cu: copy-engine pushes
// + flags, no SM and no collective library involved) and drives all of them from the calling thread:
// every launch, copy and flag operation is asynchronous, so one host thread keeps eight GPUs busy.
// Per convergence check each slab's four partial sums come back to the host and are folded in slab
// order.  Uploads and downloads run on one host thread per device.
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "hmi_init.cuh"
#include "hmi_internal.cuh"
#include "hmi_xfer.cuh"

using hmi::fail;

namespace {

void slab_bounds(int rows, int n, std::vector<int> &r0, std::vector<int> &r1)
{
    const int base = rows / n, rem = rows % n;
    int a = 0;
    for (int k = 0; k < n; ++k) {
        const int b = a + base + (k < rem ? 1 : 0);
        r0.push_back(a);
        r1.push_back(b);
        a = b;
    }
}

// run fn(i) for every slab, one host thread per device; the first failure wins
template <class F>
int for_each_slab_parallel(hmi_group *g, F &&fn)
{
    std::vector<int> rc(g->n, HMI_OK);
    std::vector<std::string> msg(g->n);
    std::vector<std::thread> th;
    auto body = [&](int i) {
        rc[i] = fn(i);
        if (rc[i] != HMI_OK) msg[i] = hmi_last_error();     // the error text is thread local
    };
    for (int i = 1; i < g->n; ++i) th.emplace_back(body, i);
    body(0);
    for (auto &t : th) t.join();
    for (int i = 0; i < g->n; ++i)
        if (rc[i] != HMI_OK) return fail(rc[i], "slab %d (device %d): %s", i, g->devices[i], msg[i].c_str());
    return HMI_OK;
}

// n sweeps on every slab, in blocks of at most exchange_every with a halo exchange behind each
// block; with `norm` the convergence sums of the last sweep are reduced and folded in slab order.
int group_sweeps(hmi_group *g, long long n, bool norm, hmi_norm *out)
{
    if (!g->has_problem) return fail(HMI_ERR_STATE, "hmi_group_set_problem_host has not been called");
    if (n < 0) return fail(HMI_ERR_BAD_ARG, "negative sweep count");
    if (norm && n < 1) return fail(HMI_ERR_BAD_ARG, "a convergence check needs at least one sweep");
    long long left = n;
    while (left > 0) {
        const long long k = g->n == 1 ? left : (left < g->exchange_every ? left : g->exchange_every);
        const bool last = (k == left);
        for (int i = 0; i < g->n; ++i) {
            hmi_solver *s = g->slab[i];
            CUDA_TRY(cudaSetDevice(s->p.device));
            int rc;
            if (g->n == 1 || (last && norm)) rc = hmi::do_sweeps(s, k, last && norm, s->own);
            else rc = hmi_sweeps_exchange(s, k, s->own, s->own_side);
            if (rc != HMI_OK) return rc;
        }
        left -= k;
    }
    if (!norm) return HMI_OK;
    hmi_norm tot = {0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < g->n; ++i) {
        hmi_solver *s = g->slab[i];
        CUDA_TRY(cudaSetDevice(s->p.device));
        hmi_norm nm;
        const int rc = hmi::fetch_norm(s, s->own, &nm);
        if (rc != HMI_OK) return rc;
        tot.sum_d2 += nm.sum_d2;
        tot.sum_f2 += nm.sum_f2;
        tot.max_abs = nm.max_abs > tot.max_abs ? nm.max_abs : tot.max_abs;
        tot.has_nan = nm.has_nan > tot.has_nan ? nm.has_nan : tot.has_nan;
    }
    *out = tot;
    return HMI_OK;
}

int group_sync(hmi_group *g)
{
    for (hmi_solver *s : g->slab) {         // (a group whose creation failed half-way holds fewer than n slabs)
        const int rc = hmi::sync_solver(s);
        if (rc != HMI_OK) return rc;
    }
    return HMI_OK;
}

}  // namespace

extern "C" {

void hmi_group_destroy(hmi_group *g)
{
    if (!g) return;
    group_sync(g);                        // nobody may still be pushing into a block about to be freed
    for (hmi_solver *s : g->slab) hmi_destroy(s);
    delete g;
}

int hmi_group_create(const hmi_params *params, const int32_t *devices, int32_t n_devices, int32_t exchange_every,
                     hmi_group **out)
{
    if (!params || !devices || !out) return fail(HMI_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    if (params->struct_size != (int32_t)sizeof(hmi_params)) return fail(HMI_ERR_BAD_ARG, "hmi_params size mismatch");
    if (n_devices < 1 || n_devices > 64) return fail(HMI_ERR_BAD_ARG, "n_devices must be between 1 and 64");
    if (params->ghost_top || params->ghost_bottom) return fail(HMI_ERR_BAD_ARG, "a group takes the whole grid (no ghost rows)");
    if (exchange_every < 0) return fail(HMI_ERR_BAD_ARG, "negative exchange_every");
    // (a device may be listed more than once: its slabs then share that GPU — how a one-GPU box
    // exercises the whole halo protocol in the tests)
    const int radius = hmi::method_radius(params->method);
    // slabs need a few rows each: fewer devices for a small grid
    int n = n_devices;
    while (n > 1 && params->rows / n < 8 * radius + 8) --n;

    hmi_group *g = new (std::nothrow) hmi_group();
    if (!g) return fail(HMI_ERR_CUDA, "out of host memory");
    g->p = *params;
    g->n = n;
    g->radius = radius;
    g->devices.assign(devices, devices + n);
    slab_bounds(params->rows, n, g->r0, g->r1);
    int min_rows = params->rows;
    for (int i = 0; i < n; ++i) min_rows = (g->r1[i] - g->r0[i]) < min_rows ? (g->r1[i] - g->r0[i]) : min_rows;
    // a whole number of temporal blocks by default (harmonic 6, TV/AMLE 2, CCST 3 sweeps per launch)
    // a whole number of temporal blocks per exchange: CCST 3 / 4 sweeps per launch, harmonic 6 (fp32: 8), TV / AMLE 2
    int E = exchange_every > 0 ? exchange_every
                               : (radius == 1 ? (params->method == HMI_HARMONIC && params->dtype == HMI_F32 ? 16 : 18) : 12);
    if (n > 1 && E > (min_rows - 1) / radius) E = (min_rows - 1) / radius;
    if (E < 1) E = 1;
    g->exchange_every = E;
    g->ghost = n > 1 ? radius * E : 0;

    int rc = HMI_OK;
    for (int i = 0; i < n && rc == HMI_OK; ++i) {
        hmi_params sp = *params;
        sp.device = g->devices[i];
        sp.rows = g->r1[i] - g->r0[i];
        sp.ghost_top = i > 0 ? g->ghost : 0;
        sp.ghost_bottom = i < n - 1 ? g->ghost : 0;
        hmi_solver *s = nullptr;
        rc = hmi_create(&sp, &s);
        if (rc == HMI_OK) g->slab.push_back(s);
    }
    std::vector<hmi_halo_info> info(n);
    for (int i = 0; i < n && rc == HMI_OK && n > 1; ++i) rc = hmi_halo_setup(g->slab[i], &info[i]);
    for (int i = 0; i < n && rc == HMI_OK && n > 1; ++i) {
        if (i > 0) rc = hmi_halo_connect(g->slab[i], 0, info[i - 1].block, info[i - 1].device);
        if (rc == HMI_OK && i < n - 1) rc = hmi_halo_connect(g->slab[i], 1, info[i + 1].block, info[i + 1].device);
    }
    if (rc != HMI_OK) {
        const std::string keep = hmi_last_error();
        hmi_group_destroy(g);
        return fail(rc, "%s", keep.c_str());
    }
    *out = g;
    return HMI_OK;
}

int hmi_group_set_problem_host(hmi_group *g, const void *image, size_t image_pitch_bytes, const uint8_t *mask,
                               size_t mask_pitch_bytes)
{
    if (!g || !image || !mask) return fail(HMI_ERR_BAD_ARG, "null argument");
    const unsigned hw = std::thread::hardware_concurrency();
    int per = hw ? (int)(hw / (unsigned)g->n) : 2;
    hmi::set_xfer_threads(per < 2 ? 2 : per);
    const int rc = for_each_slab_parallel(g, [&](int i) {
        hmi_solver *s = g->slab[i];
        const long long first = (long long)g->r0[i] - s->p.ghost_top;      // global row of the slab's first physical row
        return hmi_set_problem_host(s, (const char *)image + first * (long long)image_pitch_bytes, image_pitch_bytes,
                                    mask + first * (long long)mask_pitch_bytes, mask_pitch_bytes);
    });
    hmi::set_xfer_threads(6);
    g->has_problem = rc == HMI_OK;
    return rc;
}

int hmi_group_fill_unknown(hmi_group *g, double value)
{
    if (!g) return fail(HMI_ERR_BAD_ARG, "null group");
    for (hmi_solver *s : g->slab) {
        const int rc = hmi::fill_unknown(s, value);
        if (rc != HMI_OK) return rc;
    }
    return HMI_OK;
}

int hmi_group_sweeps(hmi_group *g, int64_t n)
{
    if (!g) return fail(HMI_ERR_BAD_ARG, "null group");
    return group_sweeps(g, n, false, nullptr);
}

int hmi_group_sweeps_norm(hmi_group *g, int64_t n, hmi_norm *out)
{
    if (!g || !out) return fail(HMI_ERR_BAD_ARG, "null argument");
    return group_sweeps(g, n, true, out);
}

int hmi_group_sync(hmi_group *g)
{
    if (!g) return fail(HMI_ERR_BAD_ARG, "null group");
    return group_sync(g);
}

int hmi_group_run(hmi_group *g, hmi_status *status)
{
    if (!g || !status) return fail(HMI_ERR_BAD_ARG, "null argument");
    if (!g->has_problem) return fail(HMI_ERR_STATE, "hmi_group_set_problem_host has not been called");
    const int rc = hmi::reference_loop(g->p, [&](long long n, bool norm, hmi_norm *nm) { return group_sweeps(g, n, norm, nm); },
                                       status);
    if (rc != HMI_OK) return rc;
    return group_sync(g);
}

int hmi_group_set_term_thres(hmi_group *g, double term_thres)
{
    if (!g) return fail(HMI_ERR_BAD_ARG, "null group");
    if (!(term_thres > 0.0) && term_thres == term_thres) return fail(HMI_ERR_BAD_ARG, "term_thres must be larger than zero");
    g->p.term_thres = term_thres;
    for (hmi_solver *s : g->slab) s->p.term_thres = term_thres;
    return HMI_OK;
}

int hmi_group_get_result_host(hmi_group *g, void *out, size_t out_pitch_bytes)
{
    if (!g || !out) return fail(HMI_ERR_BAD_ARG, "null argument");
    if (!g->has_problem) return fail(HMI_ERR_STATE, "no problem loaded");
    const unsigned hw = std::thread::hardware_concurrency();
    int per = hw ? (int)(hw / (unsigned)g->n) : 2;
    hmi::set_xfer_threads(per < 2 ? 2 : per);
    const int rc = for_each_slab_parallel(g, [&](int i) {
        return hmi_get_result_host(g->slab[i], (char *)out + (long long)g->r0[i] * (long long)out_pitch_bytes, out_pitch_bytes);
    });
    hmi::set_xfer_threads(6);
    return rc;
}

}  // extern "C"

// ---- multigrid: the pyramid lives on the device(s) ------------------------------------------------

static int enable_peer(int from, int to)
{
    if (from == to) return HMI_OK;
    int can = 0;
    CUDA_TRY(cudaDeviceCanAccessPeer(&can, from, to));
    if (!can) return fail(HMI_ERR_UNSUPPORTED, "device %d cannot access device %d's memory", from, to);
    CUDA_TRY(cudaSetDevice(from));
    const cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(HMI_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", from, to, cudaGetErrorString(e));
    cudaGetLastError();
    return HMI_OK;
}

static int make_view(const hmi_group *g, bool mask, hmi::SlabView *v)
{
    if (g->n > hmi::MAX_VIEW_SLABS) return fail(HMI_ERR_UNSUPPORTED, "more than %d slabs", hmi::MAX_VIEW_SLABS);
    v->n = g->n;
    for (int k = 0; k < g->n; ++k) {
        const hmi_solver *s = g->slab[k];
        v->row0[k] = g->r0[k];
        v->base[k] = mask ? (const char *)s->mask + (size_t)s->p.ghost_top * s->mpitch
                          : (const char *)s->buf[s->cur] + (size_t)s->p.ghost_top * s->pitch * s->elem;
    }
    v->row0[g->n] = g->p.rows;
    v->pitch = mask ? g->slab[0]->mpitch : g->slab[0]->pitch * g->slab[0]->elem;
    return HMI_OK;
}

template <typename T>
static int pyramid_slab(hmi_solver *s, int row_first, int level_rows, int stride, int scrub_nan, const hmi::SlabView &fimg,
                        const hmi::SlabView &fmask, const hmi::SlabView &coarse, const int32_t *sx, const int32_t *sx1,
                        const void *ax0, const void *ax1, const int32_t *sy, const int32_t *sy1, const void *by0,
                        const void *by1, double *lo, double *hi, double *nanv, int *had_nan)
{
    const int cols = s->p.cols;
    const size_t n_int = 2 * (size_t)cols + 2 * (size_t)level_rows;
    const size_t off_val = (n_int * sizeof(int) + 15) / 16 * 16;
    const size_t off_part = (off_val + n_int * sizeof(T) + 15) / 16 * 16;
    const int max_blocks = hmi::upsample_max_blocks(cols);
    const size_t off_flag = off_part + (size_t)max_blocks * 3 * sizeof(double);
    const size_t total = off_flag + 16;
    CUDA_TRY(cudaSetDevice(s->p.device));
    char *d = nullptr;
    CUDA_TRY(cudaMalloc((void **)&d, total));
    std::vector<char> h(off_part);
    int *hi_ = reinterpret_cast<int *>(h.data());
    memcpy(hi_, sx, cols * sizeof(int));
    memcpy(hi_ + cols, sx1, cols * sizeof(int));
    memcpy(hi_ + 2 * cols, sy, level_rows * sizeof(int));
    memcpy(hi_ + 2 * cols + level_rows, sy1, level_rows * sizeof(int));
    T *hv = reinterpret_cast<T *>(h.data() + off_val);
    memcpy(hv, ax0, cols * sizeof(T));
    memcpy(hv + cols, ax1, cols * sizeof(T));
    memcpy(hv + 2 * cols, by0, level_rows * sizeof(T));
    memcpy(hv + 2 * cols + level_rows, by1, level_rows * sizeof(T));
    cudaError_t e = cudaMemcpyAsync(d, h.data(), off_part, cudaMemcpyHostToDevice, s->own);
    if (e == cudaSuccess) e = cudaMemsetAsync(d + off_flag, 0, 16, s->own);
    int nblocks = 0;
    if (e == cudaSuccess) {
        hmi::PyramidBlend<T> a;
        s->cur = 0;
        a.f = (T *)s->buf[0];
        a.mask = s->mask;
        a.pitch = s->pitch; a.mpitch = s->mpitch;
        a.phys_rows = s->phys_rows; a.cols = cols;
        a.row_first = row_first;
        a.stride = stride; a.scrub_nan = scrub_nan;
        a.fine_img = fimg; a.fine_mask = fmask; a.coarse = coarse;
        const int *di = reinterpret_cast<const int *>(d);
        const T *dv = reinterpret_cast<const T *>(d + off_val);
        a.sx = di; a.sx1 = di + cols; a.sy = di + 2 * cols; a.sy1 = di + 2 * cols + level_rows;
        a.ax0 = dv; a.ax1 = dv + cols; a.by0 = dv + 2 * cols; a.by1 = dv + 2 * cols + level_rows;
        a.partials = reinterpret_cast<double *>(d + off_part);
        a.nan_flag = reinterpret_cast<int *>(d + off_flag);
        e = hmi::launch_pyramid_blend<T>(a, &nblocks, s->own);
    }
    std::vector<double> part;
    int flag = 0;
    if (e == cudaSuccess) {
        part.resize((size_t)nblocks * 3);
        e = cudaMemcpyAsync(part.data(), d + off_part, part.size() * sizeof(double), cudaMemcpyDeviceToHost, s->own);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, d + off_flag, sizeof(int), cudaMemcpyDeviceToHost, s->own);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->own);
    cudaFree(d);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(HMI_ERR_CUDA, "multigrid level transition failed: %s", cudaGetErrorString(e));
    }
    for (int b = 0; b < nblocks; ++b) {
        *lo = part[3 * b] < *lo ? part[3 * b] : *lo;
        *hi = part[3 * b + 1] > *hi ? part[3 * b + 1] : *hi;
        *nanv = part[3 * b + 2] > *nanv ? part[3 * b + 2] : *nanv;
    }
    *had_nan |= flag;
    s->has_problem = true;
    s->halo_state = 0;              // ghost rows were computed like every other row of the level
    s->last_main = s->own;
    return HMI_OK;
}

extern "C" {

int hmi_group_set_problem_pyramid(hmi_group *level, hmi_group *finest, int32_t stride, hmi_group *coarse,
                                  const int32_t *sx, const int32_t *sx1, const void *ax0, const void *ax1,
                                  const int32_t *sy, const int32_t *sy1, const void *by0, const void *by1,
                                  int32_t scrub_nan, hmi_blend_info *info)
{
    if (!level || !finest || !coarse || !sx || !sx1 || !ax0 || !ax1 || !sy || !sy1 || !by0 || !by1 || !info)
        return fail(HMI_ERR_BAD_ARG, "null argument");
    if (level == coarse) return fail(HMI_ERR_BAD_ARG, "a level cannot be its own coarser level");
    if (stride < 1) return fail(HMI_ERR_BAD_ARG, "stride must be >= 1");
    if ((stride == 1) != (level == finest)) return fail(HMI_ERR_BAD_ARG, "stride 1 means the finest level itself, in place");
    if (level->p.dtype != finest->p.dtype || level->p.dtype != coarse->p.dtype)
        return fail(HMI_ERR_BAD_ARG, "all levels must share the dtype");
    if (!finest->has_problem) return fail(HMI_ERR_STATE, "the finest level has not been loaded");
    if (!coarse->has_problem) return fail(HMI_ERR_STATE, "the coarser level has no solution yet");
    const int want_rows = (finest->p.rows + stride - 1) / stride, want_cols = (finest->p.cols + stride - 1) / stride;
    if (level->p.rows != want_rows || level->p.cols != want_cols)
        return fail(HMI_ERR_BAD_ARG, "a stride-%d level of a %dx%d grid is %dx%d, not %dx%d", stride, finest->p.rows,
                    finest->p.cols, want_rows, want_cols, level->p.rows, level->p.cols);
    int rc = group_sync(finest);
    if (rc == HMI_OK) rc = group_sync(coarse);
    if (rc == HMI_OK) rc = group_sync(level);
    if (rc != HMI_OK) return rc;
    for (int i = 0; i < level->n && rc == HMI_OK; ++i) {
        for (int k = 0; k < finest->n && rc == HMI_OK; ++k) rc = enable_peer(level->devices[i], finest->devices[k]);
        for (int k = 0; k < coarse->n && rc == HMI_OK; ++k) rc = enable_peer(level->devices[i], coarse->devices[k]);
    }
    hmi::SlabView vimg, vmask, vcoarse;
    if (rc == HMI_OK) rc = make_view(finest, false, &vimg);
    if (rc == HMI_OK) rc = make_view(finest, true, &vmask);
    if (rc == HMI_OK) rc = make_view(coarse, false, &vcoarse);
    if (rc != HMI_OK) return rc;
    double lo = INFINITY, hi = -INFINITY, nanv = 0.0;
    int had_nan = 0;
    for (int i = 0; i < level->n; ++i) {
        hmi_solver *s = level->slab[i];
        const int row_first = level->r0[i] - s->p.ghost_top;
        rc = level->p.dtype == HMI_F64
                 ? pyramid_slab<double>(s, row_first, level->p.rows, stride, scrub_nan, vimg, vmask, vcoarse, sx, sx1, ax0, ax1,
                                        sy, sy1, by0, by1, &lo, &hi, &nanv, &had_nan)
                 : pyramid_slab<float>(s, row_first, level->p.rows, stride, scrub_nan, vimg, vmask, vcoarse, sx, sx1, ax0, ax1,
                                       sy, sy1, by0, by1, &lo, &hi, &nanv, &had_nan);
        if (rc != HMI_OK) return rc;
    }
    level->has_problem = true;
    info->min = nanv != 0.0 ? nan("") : lo;      // np.min / np.max propagate NaN
    info->max = nanv != 0.0 ? nan("") : hi;
    info->has_nan = nanv != 0.0 ? 1 : 0;
    info->image_had_nan = had_nan;
    return HMI_OK;
}

int hmi_group_info(const hmi_group *g, int32_t *n_slabs, int32_t *exchange_every, int32_t *ghost_rows)
{
    if (!g) return fail(HMI_ERR_BAD_ARG, "null group");
    if (n_slabs) *n_slabs = g->n;
    if (exchange_every) *exchange_every = g->exchange_every;
    if (ghost_rows) *ghost_rows = g->ghost;
    return HMI_OK;
}

int hmi_group_slab(hmi_group *g, int32_t i, hmi_solver **solver, int32_t *row0, int32_t *row1)
{
    if (!g) return fail(HMI_ERR_BAD_ARG, "null group");
    if (i < 0 || i >= g->n) return fail(HMI_ERR_BAD_ARG, "slab index %d out of range (%d slabs)", i, g->n);
    if (solver) *solver = g->slab[i];
    if (row0) *row0 = g->r0[i];
    if (row1) *row1 = g->r1[i];
    return HMI_OK;
}

int64_t hmi_group_launch_count(const hmi_group *g)
{
    long long n = 0;
    if (g) for (const hmi_solver *s : g->slab) n += s->launches;
    return n;
}

int hmi_inpaint_host_devices(const hmi_params *params, const int32_t *devices, int32_t n_devices, const void *image,
                             const uint8_t *mask, int32_t fill, double fill_value, void *out, hmi_status *status)
{
    if (!params || !devices || !image || !mask || !out || !status) return fail(HMI_ERR_BAD_ARG, "null argument");
    hmi_group *g = nullptr;
    int rc = hmi_group_create(params, devices, n_devices, 0, &g);
    if (rc != HMI_OK) return rc;
    const size_t elem = params->dtype == HMI_F64 ? 8 : 4;
    rc = hmi_group_set_problem_host(g, image, (size_t)params->cols * elem, mask, (size_t)params->cols);
    std::thread host_init;           // fill == 2: the caller's array is initialised in place once it has been read
    if (rc == HMI_OK && fill == 2)
        host_init = std::thread(hmi_host_fill_unknown, const_cast<void *>(image), mask,
                                (int64_t)params->rows * params->cols, params->dtype, fill_value);
    if (rc == HMI_OK && fill) rc = hmi_group_fill_unknown(g, fill_value);
    if (rc == HMI_OK) rc = hmi_group_run(g, status);
    if (rc == HMI_OK) rc = hmi_group_get_result_host(g, out, (size_t)params->cols * elem);
    if (host_init.joinable()) host_init.join();
    std::string keep = rc == HMI_OK ? "" : hmi_last_error();
    hmi_group_destroy(g);
    if (rc != HMI_OK) return fail(rc, "%s", keep.c_str());
    return rc;
}

}  // extern "C"
