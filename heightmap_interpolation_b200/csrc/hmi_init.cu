// 'nearest' initialiser on the device.
//
// Reference: Initializer.initialize with init_with='nearest'
// (heightmap_interpolation/inpainting/initializer.py:19-34): scipy.interpolate.griddata(method=
// 'nearest') over ALL known cells, i.e. every unknown cell takes the value of its Euclidean-nearest
// known cell (a cKDTree query per cell; 3 s of host time on a 2048^2 grid, and the CLI default,
// apps/apps_common.py:52).
//
// Here: an exact two-pass Euclidean nearest-known-cell search.
//   1. columns: for every cell the signed row offset to the nearest known cell of ITS OWN column
//      (one thread per column, a downward and an upward scan; coalesced across columns);
//   2. rows: the nearest known cell of (i, j) is, for some column j', the column-nearest cell of
//      (i, j'), so  d^2(i, j) = min_j' (j - j')^2 + off(i, j')^2.  Each unknown cell walks outwards
//      k = 1, 2, ... over the columns j -+ k and stops as soon as k^2 >= best: exact, and as cheap as
//      the hole is small.
// Ties (several known cells at the same distance) are resolved towards the smaller |column offset|,
// then the left column, then the upper row.  SciPy's tie-break is whatever order its KD-tree visits
// leaves in, so parity with the reference is "same distance, value of a cell at that distance"; where
// the nearest cell is unique the values are identical (tests/test_gpu_parity.py).
#include "hmi_common.cuh"
#include "hmi_init.cuh"

namespace hmi {

namespace {

constexpr int NONE = 0x3fffffff;     // no known cell in this column

__global__ void nearest_cols_kernel(const uint8_t *__restrict__ mask, long long mpitch, int rows, int cols,
                                    int *__restrict__ off, long long opitch)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    int last = -NONE;
#pragma unroll 8
    for (int i = 0; i < rows; ++i) {                       // nearest known cell at or above
        if (mask[(long long)i * mpitch + j]) last = i;
        off[(long long)i * opitch + j] = last == -NONE ? NONE : last - i;
    }
    int nxt = NONE;
#pragma unroll 8
    for (int i = rows - 1; i >= 0; --i) {                  // ... at or below; keep the closer (upper on ties)
        if (mask[(long long)i * mpitch + j]) nxt = i;
        const int up = off[(long long)i * opitch + j];
        const int dn = nxt == NONE ? NONE : nxt - i;
        const int aup = up == NONE ? NONE : -up;
        off[(long long)i * opitch + j] = (aup <= dn) ? up : dn;
    }
}

template <typename T>
__device__ __forceinline__ void nearest_cell(const T *__restrict__ img, long long pitch,
                                             const uint8_t *__restrict__ mask, long long mpitch,
                                             const int *__restrict__ off, long long opitch, int cols,
                                             T *__restrict__ out, long long out_pitch, int *__restrict__ status,
                                             const int i, const int j)
{
    const long long r = (long long)i * pitch;
    if (mask[(long long)i * mpitch + j]) {
        out[(long long)i * out_pitch + j] = img[r + j];
        return;
    }
    const int *orow = off + (long long)i * opitch;
    long long best = 0x7fffffffffffffffLL;
    int bj = -1, bo = 0;
    {
        const int o = orow[j];
        if (o != NONE) { best = (long long)o * o; bj = j; bo = o; }
    }
    const int kmax = max(j, cols - 1 - j);
    for (int k = 1; k <= kmax; ++k) {
        const long long k2 = (long long)k * k;
        if (k2 >= best) break;
        const int jl = j - k, jr = j + k;
        if (jl >= 0) {
            const int o = orow[jl];
            if (o != NONE) {
                const long long d = k2 + (long long)o * o;
                if (d < best) { best = d; bj = jl; bo = o; }
            }
        }
        if (jr < cols) {
            const int o = orow[jr];
            if (o != NONE) {
                const long long d = k2 + (long long)o * o;
                if (d < best) { best = d; bj = jr; bo = o; }
            }
        }
    }
    if (bj < 0) {                                          // no known cell anywhere
        atomicExch(status, 1);
        out[(long long)i * out_pitch + j] = img[r + j];
        return;
    }
    out[(long long)i * out_pitch + j] = img[(long long)(i + bo) * pitch + bj];
}

template <typename T>
__global__ void nearest_rows_kernel(const T *__restrict__ img, long long pitch, const uint8_t *__restrict__ mask,
                                    long long mpitch, const int *__restrict__ off, long long opitch, int rows,
                                    int cols, T *__restrict__ out, long long out_pitch, int *__restrict__ status)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    for (int i = blockIdx.y; i < rows; i += gridDim.y)
        nearest_cell<T>(img, pitch, mask, mpitch, off, opitch, cols, out, out_pitch, status, i, j);
}
}  // namespace

template <typename T>
cudaError_t launch_init_nearest(const T *img, long long pitch, const uint8_t *mask, long long mpitch, int rows,
                                int cols, int *off, long long opitch, T *out, long long out_pitch, int *status,
                                cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    nearest_cols_kernel<<<(cols + 63) / 64, 64, 0, st>>>(mask, mpitch, rows, cols, off, opitch);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const dim3 grid((cols + 255) / 256, rows < 65535 ? rows : 65535);
    nearest_rows_kernel<T><<<grid, 256, 0, st>>>(img, pitch, mask, mpitch, off, opitch, rows, cols, out, out_pitch,
                                                 status);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Multigrid: coarse solution -> next finer level (fd_pde_inpainter.py:255-263)
//   upscaled = cv2.resize(coarse, (cols, rows))            default INTER_LINEAR, half-pixel centres
//   image    = upscaled * (~mask) + image * mask
// The interpolation taps and weights of both axes come from the host (they are 1-D tables, computed
// exactly as the host restatement does); every product and sum here is a separate IEEE rounding in
// the same order as that restatement, and the blend is the reference's arithmetic one, so a NaN left
// in an unknown cell of a decimated image poisons the level exactly as it does in the reference.
namespace {

template <typename T>
__global__ void upsample_blend_kernel(T *__restrict__ f, long long pitch, const uint8_t *__restrict__ mask,
                                      long long mpitch, int rows, int cols, const T *__restrict__ coarse,
                                      long long cpitch, const int *__restrict__ sx, const int *__restrict__ sx1,
                                      const T *__restrict__ ax0, const T *__restrict__ ax1,
                                      const int *__restrict__ sy, const int *__restrict__ sy1,
                                      const T *__restrict__ by0, const T *__restrict__ by1, int scrub_nan,
                                      double *__restrict__ partials, int *__restrict__ nan_flag)
{
    typedef Op<T> O;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double lo = INFINITY, hi = -INFINITY, nanv = 0.0;
    if (j < cols) {
        const int x0 = sx[j], x1 = sx1[j];
        const T a0 = ax0[j], a1 = ax1[j];
        for (int i = blockIdx.y; i < rows; i += gridDim.y) {
            const T *c0 = coarse + (long long)sy[i] * cpitch, *c1 = coarse + (long long)sy1[i] * cpitch;
            const T t0 = O::add(O::mul(c0[x0], a0), O::mul(c0[x1], a1));
            const T t1 = O::add(O::mul(c1[x0], a0), O::mul(c1[x1], a1));
            const T up = O::add(O::mul(t0, by0[i]), O::mul(t1, by1[i]));
            T img = f[(long long)i * pitch + j];
            if (scrub_nan && img != img) { img = T(0); *nan_flag = 1; }
            const T m = mask[(long long)i * mpitch + j] ? T(1) : T(0);
            const T v = O::add(O::mul(up, T(1) - m), O::mul(img, m));
            f[(long long)i * pitch + j] = v;
            const double d = (double)v;
            if (d != d) nanv = 1.0;
            lo = d < lo ? d : lo;
            hi = d > hi ? d : hi;
        }
    }
    // block min / max / NaN flag -> one partial per block (folded on the host)
    __shared__ double sh[3][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        nanv = fmax(nanv, __shfl_xor_sync(0xffffffffu, nanv, o));
    }
    if (lane == 0) { sh[0][warp] = lo; sh[1][warp] = hi; sh[2][warp] = nanv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 1; w < nw; ++w) {
            lo = fmin(lo, sh[0][w]); hi = fmax(hi, sh[1][w]); nanv = fmax(nanv, sh[2][w]);
        }
        const long long b = (long long)blockIdx.y * gridDim.x + blockIdx.x;
        partials[3 * b + 0] = lo; partials[3 * b + 1] = hi; partials[3 * b + 2] = nanv;
    }
}

// The same level transition for a pyramid that lives on the device(s): the level's image and mask are
// gathered with stride 2^l from the finest level (halve_image applied l times picks every 2^l-th cell,
// misc/image_proc.py:4-12; the mask is ANDed with "not NaN", fd_pde_inpainter.py:223-224), the coarser
// solution is up-sampled from wherever its row slabs live (peer memory included) and blended in.
template <typename T>
__device__ __forceinline__ const T *view_row(const SlabView &v, int y)
{
    int k = 0;
    while (k + 1 < v.n && y >= v.row0[k + 1]) ++k;
    return reinterpret_cast<const T *>(v.base[k] + (long long)(y - v.row0[k]) * v.pitch);
}

template <typename T>
__global__ void pyramid_blend_kernel(const PyramidBlend<T> a)
{
    typedef Op<T> O;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double lo = INFINITY, hi = -INFINITY, nanv = 0.0;
    if (j < a.cols) {
        const int x0 = a.sx[j], x1 = a.sx1[j];
        const T a0 = a.ax0[j], a1 = a.ax1[j];
        for (int p = blockIdx.y; p < a.phys_rows; p += gridDim.y) {
            const int i = a.row_first + p;                       // row of this level's (global) grid
            const T *c0 = view_row<T>(a.coarse, a.sy[i]), *c1 = view_row<T>(a.coarse, a.sy1[i]);
            const T t0 = O::add(O::mul(c0[x0], a0), O::mul(c0[x1], a1));
            const T t1 = O::add(O::mul(c1[x0], a0), O::mul(c1[x1], a1));
            const T up = O::add(O::mul(t0, a.by0[i]), O::mul(t1, a.by1[i]));
            T img;
            bool known;
            if (a.stride == 1) {                                 // the finest level itself, in place
                img = a.f[(long long)p * a.pitch + j];
                known = a.mask[(long long)p * a.mpitch + j] != 0;
                if (a.scrub_nan && img != img) { img = T(0); *a.nan_flag = 1; }
            } else {
                img = view_row<T>(a.fine_img, i * a.stride)[(long long)j * a.stride];
                known = view_row<uint8_t>(a.fine_mask, i * a.stride)[(long long)j * a.stride] != 0 && img == img;
                a.mask[(long long)p * a.mpitch + j] = known ? 1 : 0;
            }
            const T m = known ? T(1) : T(0);
            const T v = O::add(O::mul(up, T(1) - m), O::mul(img, m));
            a.f[(long long)p * a.pitch + j] = v;
            const double d = (double)v;
            if (d != d) nanv = 1.0;
            lo = d < lo ? d : lo;
            hi = d > hi ? d : hi;
        }
    }
    __shared__ double sh[3][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        nanv = fmax(nanv, __shfl_xor_sync(0xffffffffu, nanv, o));
    }
    if (lane == 0) { sh[0][warp] = lo; sh[1][warp] = hi; sh[2][warp] = nanv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 1; w < nw; ++w) {
            lo = fmin(lo, sh[0][w]); hi = fmax(hi, sh[1][w]); nanv = fmax(nanv, sh[2][w]);
        }
        const long long b = (long long)blockIdx.y * gridDim.x + blockIdx.x;
        a.partials[3 * b + 0] = lo; a.partials[3 * b + 1] = hi; a.partials[3 * b + 2] = nanv;
    }
}

}  // namespace

template <typename T>
cudaError_t launch_pyramid_blend(const PyramidBlend<T> &a, int *nblocks, cudaStream_t st)
{
    const dim3 grid((a.cols + 127) / 128, a.phys_rows < UPSAMPLE_MAX_GRID_Y ? a.phys_rows : UPSAMPLE_MAX_GRID_Y);
    *nblocks = (int)(grid.x * grid.y);
    pyramid_blend_kernel<T><<<grid, 128, 0, st>>>(a);
    return cudaGetLastError();
}

template cudaError_t launch_pyramid_blend<float>(const PyramidBlend<float> &, int *, cudaStream_t);
template cudaError_t launch_pyramid_blend<double>(const PyramidBlend<double> &, int *, cudaStream_t);

template <typename T>
cudaError_t launch_upsample_blend(T *f, long long pitch, const uint8_t *mask, long long mpitch, int rows, int cols,
                                  const T *coarse, long long cpitch, const int *sx, const int *sx1, const T *ax0,
                                  const T *ax1, const int *sy, const int *sy1, const T *by0, const T *by1,
                                  int scrub_nan, double *partials, int *nan_flag, int *nblocks, cudaStream_t st)
{
    const dim3 grid((cols + 127) / 128, rows < UPSAMPLE_MAX_GRID_Y ? rows : UPSAMPLE_MAX_GRID_Y);
    *nblocks = (int)(grid.x * grid.y);
    upsample_blend_kernel<T><<<grid, 128, 0, st>>>(f, pitch, mask, mpitch, rows, cols, coarse, cpitch, sx, sx1, ax0,
                                                   ax1, sy, sy1, by0, by1, scrub_nan, partials, nan_flag);
    return cudaGetLastError();
}

template cudaError_t launch_upsample_blend<float>(float *, long long, const uint8_t *, long long, int, int,
                                                  const float *, long long, const int *, const int *, const float *,
                                                  const float *, const int *, const int *, const float *,
                                                  const float *, int, double *, int *, int *, cudaStream_t);
template cudaError_t launch_upsample_blend<double>(double *, long long, const uint8_t *, long long, int, int,
                                                   const double *, long long, const int *, const int *,
                                                   const double *, const double *, const int *, const int *,
                                                   const double *, const double *, int, double *, int *, int *,
                                                   cudaStream_t);

template cudaError_t launch_init_nearest<float>(const float *, long long, const uint8_t *, long long, int, int, int *,
                                                long long, float *, long long, int *, cudaStream_t);
template cudaError_t launch_init_nearest<double>(const double *, long long, const uint8_t *, long long, int, int,
                                                 int *, long long, double *, long long, int *, cudaStream_t);

}  // namespace hmi
