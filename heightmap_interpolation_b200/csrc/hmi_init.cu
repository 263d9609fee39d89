// 'nearest' initialiser on the device.
//
// Reference: Initializer.initialize with init_with='nearest'
// (heightmap_interpolation/inpainting/initializer.py:19-34): scipy.interpolate.griddata(method=
// 'nearest') over ALL known cells, i.e. every unknown cell takes the value of its Euclidean-nearest
// known cell (a cKDTree query per cell; minutes of host time on a 4096^2 grid, and the CLI default,
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

template cudaError_t launch_init_nearest<float>(const float *, long long, const uint8_t *, long long, int, int, int *,
                                                long long, float *, long long, int *, cudaStream_t);
template cudaError_t launch_init_nearest<double>(const double *, long long, const uint8_t *, long long, int, int,
                                                 int *, long long, double *, long long, int *, cudaStream_t);

}  // namespace hmi
