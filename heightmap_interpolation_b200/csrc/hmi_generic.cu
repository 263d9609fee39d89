// Generic single-sweep kernel: one thread per grid cell, every method x border rule x orientation.
//
// This kernel is the exact restatement of the reference semantics on the device: it evaluates the
// closed-form per-cell step functions of SURVEY.md A.3 with the border rule applied to every
// field a 3x3 filter is applied to (intermediates included), keeps multiplies and adds as
// separate IEEE roundings in the reference's expression order, and accumulates each 3x3 sum in
// double before rounding to the grid dtype (what convolve_at_mask.py:51-62 does).  It serves
// the methods/borders the streaming kernel does not cover, tiny grids, and it is the on-device
// cross-check for the fast kernels.
//
// Reference semantics restated (paths under heightmap_interpolation/inpainting/):
//   harmonic  sobolev_inpainter.py:15-17        ccst  ccst_inpainter.py:31-34
//   tv        tv_inpainter.py:24-33 + differential.py:146-187
//   amle      amle_inpainter.py:41-59           update  update_at_mask.py:6-14
//   relaxation / term_check  fd_pde_inpainter.py:317-318, 368-379
#include "hmi_common.cuh"

namespace hmi {

template <typename T>
struct Field {
    const T *__restrict__ f;
    long long pitch;
    int rows, cols, border, tb, bb;

    // Is (i, j) a cell of the global grid (ghost rows are)?
    __device__ __forceinline__ bool inside(int i, int j) const
    {
        if (j < 0 || j >= cols) return false;
        if (i < 0 && tb) return false;
        if (i >= rows && bb) return false;
        return true;
    }
    // Fold an outside position by the border rule; false = tap dropped (zero padding).
    __device__ __forceinline__ bool fold(int &i, int &j) const
    {
        if (j < 0) {
            if (border == HMI_BORDER_ZERO) return false;
            j = (border == HMI_BORDER_REFLECT101) ? -j : -j - 1;
        } else if (j >= cols) {
            if (border == HMI_BORDER_ZERO) return false;
            j = (border == HMI_BORDER_REFLECT101) ? 2 * (cols - 1) - j : 2 * cols - 1 - j;
        }
        if (i < 0 && tb) {
            if (border == HMI_BORDER_ZERO) return false;
            i = (border == HMI_BORDER_REFLECT101) ? -i : -i - 1;
        } else if (i >= rows && bb) {
            if (border == HMI_BORDER_ZERO) return false;
            i = (border == HMI_BORDER_REFLECT101) ? 2 * (rows - 1) - i : 2 * rows - 1 - i;
        }
        return true;
    }
    // f with the border rule
    __device__ __forceinline__ T F(int i, int j) const
    {
        if (!fold(i, j)) return T(0);
        return f[(long long)i * pitch + j];
    }
    // 5-point Laplacian at an inside position, taps in row-major order, double accumulator
    __device__ __forceinline__ T lap(int i, int j) const
    {
        double num = 0.0;
        num += (double)F(i - 1, j);
        num += (double)F(i, j - 1);
        num += -4.0 * (double)F(i, j);
        num += (double)F(i, j + 1);
        num += (double)F(i + 1, j);
        return (T)num;
    }
    // Laplacian field with the border rule applied to it (an intermediate array in the reference)
    __device__ __forceinline__ T lapB(int i, int j) const
    {
        if (!fold(i, j)) return T(0);
        return lap(i, j);
    }
    // two-tap differences: (T)((double)a - (double)b) == a - b in T, so plain subtraction is exact
    __device__ __forceinline__ T diffx(int i, int j, int s) const { return Op<T>::sub(F(i, j + s), F(i, j)); }
    __device__ __forceinline__ T diffy(int i, int j, int s) const { return Op<T>::sub(F(i + s, j), F(i, j)); }
};

template <typename T, int METHOD>
__device__ __forceinline__ T step_at(const Field<T> &g, const SweepArgs<T> &a, int i, int j)
{
    typedef Op<T> O;
    const int s = a.s;
    if (METHOD == HMI_HARMONIC) {
        return g.lap(i, j);
    } else if (METHOD == HMI_CCST) {
        // h = L f ; b = L h ; step = -1*((1-t)*b - t*h)
        const T h = g.lap(i, j);
        double num = 0.0;
        num += (double)g.lapB(i - 1, j);
        num += (double)g.lapB(i, j - 1);
        num += -4.0 * (double)h;
        num += (double)g.lapB(i, j + 1);
        num += (double)g.lapB(i + 1, j);
        const T b = (T)num;
        const T x = O::mul(a.one_minus_tension, b);
        const T y = O::mul(a.tension, h);
        return O::sub(y, x);  // -(x - y) exactly
    } else if (METHOD == HMI_TV) {
        // n = g / sqrt(gx^2 + gy^2 + eps^2) at (i,j), (i,j-s), (i-s,j); step = d/dx n0 + d/dy n1
        T n0[3], n1[3];
        const int pi[3] = {i, i, i - s};
        const int pj[3] = {j, j - s, j};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            int ii = pi[q], jj = pj[q];
            if (!g.fold(ii, jj)) { n0[q] = T(0); n1[q] = T(0); continue; }
            const T g0 = g.diffx(ii, jj, s);
            const T g1 = g.diffy(ii, jj, s);
            const T ampl = O::sqrt(O::add(O::add(O::mul(g0, g0), O::mul(g1, g1)), a.eps2));
            n0[q] = O::div(g0, ampl);
            n1[q] = O::div(g1, ampl);
        }
        const T d0 = O::sub(n0[0], n0[1]);
        const T d1 = O::sub(n1[0], n1[2]);
        return O::add(d0, d1);
    } else {
        // AMLE
        const T ux = g.diffy(i, j, s), uy = g.diffx(i, j, s);
        T ux_up = T(0), ux_left = T(0), uy_up = T(0), uy_left = T(0);
        {
            int ii = i - s, jj = j;
            if (g.fold(ii, jj)) { ux_up = g.diffy(ii, jj, s); uy_up = g.diffx(ii, jj, s); }
            ii = i; jj = j - s;
            if (g.fold(ii, jj)) { ux_left = g.diffy(ii, jj, s); uy_left = g.diffx(ii, jj, s); }
        }
        const T uxx = O::sub(ux, ux_up), uxy = O::sub(ux, ux_left);
        const T uyx = O::sub(uy, uy_up), uyy = O::sub(uy, uy_left);
        T v0 = O::sub(g.F(i + s, j), g.F(i - s, j));
        T v1 = O::sub(g.F(i, j + s), g.F(i, j - s));
        const T den = O::sqrt(O::add(O::add(O::mul(v0, v0), O::mul(v1, v1)), T(1e-15)));
        v0 = O::div(v0, den);
        v1 = O::div(v1, den);
        const T t0 = O::mul(O::mul(uxx, v0), v0);
        const T t1 = O::mul(O::mul(uyy, v1), v1);
        const T t2 = O::mul(O::mul(O::add(uxy, uyx), v0), v1);
        return O::add(O::add(t0, t1), t2);
    }
}

constexpr int GEN_BX = 64, GEN_BY = 4;

template <typename T, int METHOD>
__global__ void __launch_bounds__(GEN_BX *GEN_BY) generic_sweep_kernel(const SweepArgs<T> a)
{
    typedef Op<T> O;
    __shared__ double sh[128];
    const int j = blockIdx.x * GEN_BX + threadIdx.x;
    const int i = a.row_lo + blockIdx.y * GEN_BY + threadIdx.y;
    double d2 = 0.0, f2 = 0.0, mx = 0.0, nanf = 0.0;
    if (j < a.cols && i < a.row_hi) {
        Field<T> g{a.src, a.pitch, a.rows, a.cols, a.border, a.top_border, a.bottom_border};
        const long long o = (long long)i * a.pitch + j;
        const T fc = a.src[o];
        T fn = fc;  // known cells keep the image value they hold (update_at_mask)
        if (!a.mask[(long long)i * a.mpitch + j]) {
            const T step = step_at<T, METHOD>(g, a, i, j);
            fn = O::add(fc, O::mul(a.dt, step));
            if (a.relax) fn = O::add(O::mul(fc, a.one_minus_relax_w), O::mul(fn, a.relax_w));
        }
        a.dst[o] = fn;
        if (a.do_norm && i >= 0 && i < a.rows) {  // interior rows only: ghosts belong to a neighbour
            const T d = O::sub(fn, fc);
            d2 = (double)O::mul(d, d);
            f2 = (double)O::mul(fn, fn);
            const double ad = fabs((double)d);
            mx = (ad == ad) ? ad : 0.0;
            nanf = (ad == ad) ? 0.0 : 1.0;
        }
    }
    if (a.do_norm) block_norm_reduce(d2, f2, mx, nanf, sh, a.partials, a.counter, a.result);
}

// step_fun(f, mask) of the reference (abstract hook fd_pde_inpainter.py:406-408, the four subclasses):
// out = step where the work mask is set (or everywhere without one), the input value elsewhere —
// what a masked convolver leaves in cells it does not evaluate (convolve_at_mask.py:46).
template <typename T, int METHOD>
__global__ void __launch_bounds__(GEN_BX *GEN_BY) generic_step_kernel(const SweepArgs<T> a)
{
    const int j = blockIdx.x * GEN_BX + threadIdx.x;
    const int i = blockIdx.y * GEN_BY + threadIdx.y;
    if (j >= a.cols || i >= a.rows) return;
    Field<T> g{a.src, a.pitch, a.rows, a.cols, a.border, 1, 1};
    const long long o = (long long)i * a.pitch + j;
    const bool work = !a.mask || a.mask[(long long)i * a.mpitch + j];
    a.dst[o] = work ? step_at<T, METHOD>(g, a, i, j) : a.src[o];
}

template <typename T>
cudaError_t launch_generic_step(int method, const SweepArgs<T> &a, cudaStream_t st)
{
    dim3 grid((a.cols + GEN_BX - 1) / GEN_BX, (a.rows + GEN_BY - 1) / GEN_BY), block(GEN_BX, GEN_BY);
    switch (method) {
    case HMI_HARMONIC: generic_step_kernel<T, HMI_HARMONIC><<<grid, block, 0, st>>>(a); break;
    case HMI_CCST:     generic_step_kernel<T, HMI_CCST><<<grid, block, 0, st>>>(a); break;
    case HMI_TV:       generic_step_kernel<T, HMI_TV><<<grid, block, 0, st>>>(a); break;
    default:           generic_step_kernel<T, HMI_AMLE><<<grid, block, 0, st>>>(a); break;
    }
    return cudaGetLastError();
}

template cudaError_t launch_generic_step<float>(int, const SweepArgs<float> &, cudaStream_t);
template cudaError_t launch_generic_step<double>(int, const SweepArgs<double> &, cudaStream_t);

// Convolver.__call__(image, kernel, mask) of the reference (convolver.py:28-29): one 3x3 filter with
// the convolver's orientation (s = +1 correlation, -1 true convolution: the kernel flipped in both
// axes) and border rule, evaluated where `mask` is set (everywhere without one) and copying the input
// elsewhere (convolve_at_mask.py:46).  Taps in row-major order of the (flipped) kernel, every tap
// multiplied and added — zero weights included, like the numba loops (convolve_at_mask.py:51-62) —
// double accumulator, one rounding to the grid dtype at the end.
template <typename T>
__global__ void __launch_bounds__(GEN_BX *GEN_BY) conv3x3_kernel(const T *__restrict__ src, T *__restrict__ dst,
                                                                 long long pitch, const uint8_t *__restrict__ mask,
                                                                 long long mpitch, int rows, int cols, int border, int s,
                                                                 Kernel9 k)
{
    const int j = blockIdx.x * GEN_BX + threadIdx.x;
    const int i = blockIdx.y * GEN_BY + threadIdx.y;
    if (j >= cols || i >= rows) return;
    const long long o = (long long)i * pitch + j;
    if (mask && !mask[(long long)i * mpitch + j]) { dst[o] = src[o]; return; }
    Field<T> g{src, pitch, rows, cols, border, 1, 1};
    double num = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            // correlation: weight k[a][b] on f[i-1+a][j-1+b]; convolution: weight k[2-a][2-b] on the same cell
            const double w = s > 0 ? k.w[3 * a + b] : k.w[3 * (2 - a) + (2 - b)];
            num += w * (double)g.F(i - 1 + a, j - 1 + b);
        }
    dst[o] = (T)num;
}

template <typename T>
cudaError_t launch_conv3x3(const T *src, T *dst, long long pitch, const uint8_t *mask, long long mpitch, int rows, int cols,
                           int border, int s, const Kernel9 &k, cudaStream_t st)
{
    dim3 grid((cols + GEN_BX - 1) / GEN_BX, (rows + GEN_BY - 1) / GEN_BY), block(GEN_BX, GEN_BY);
    conv3x3_kernel<T><<<grid, block, 0, st>>>(src, dst, pitch, mask, mpitch, rows, cols, border, s, k);
    return cudaGetLastError();
}

template cudaError_t launch_conv3x3<float>(const float *, float *, long long, const uint8_t *, long long, int, int, int, int,
                                           const Kernel9 &, cudaStream_t);
template cudaError_t launch_conv3x3<double>(const double *, double *, long long, const uint8_t *, long long, int, int, int,
                                            int, const Kernel9 &, cudaStream_t);

int generic_max_blocks(int rows_total, int cols)
{
    return ((cols + GEN_BX - 1) / GEN_BX) * ((rows_total + GEN_BY - 1) / GEN_BY);
}

template <typename T>
cudaError_t launch_generic(int method, const SweepArgs<T> &a, cudaStream_t st, int *nblocks_out)
{
    const int nrows = a.row_hi - a.row_lo;
    dim3 grid((a.cols + GEN_BX - 1) / GEN_BX, (nrows + GEN_BY - 1) / GEN_BY), block(GEN_BX, GEN_BY);
    if (nblocks_out) *nblocks_out = grid.x * grid.y;
    switch (method) {
    case HMI_HARMONIC: generic_sweep_kernel<T, HMI_HARMONIC><<<grid, block, 0, st>>>(a); break;
    case HMI_CCST:     generic_sweep_kernel<T, HMI_CCST><<<grid, block, 0, st>>>(a); break;
    case HMI_TV:       generic_sweep_kernel<T, HMI_TV><<<grid, block, 0, st>>>(a); break;
    default:           generic_sweep_kernel<T, HMI_AMLE><<<grid, block, 0, st>>>(a); break;
    }
    return cudaGetLastError();
}

template cudaError_t launch_generic<float>(int, const SweepArgs<float> &, cudaStream_t, int *);
template cudaError_t launch_generic<double>(int, const SweepArgs<double> &, cudaStream_t, int *);

}  // namespace hmi
