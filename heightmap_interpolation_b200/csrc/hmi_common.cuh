// Shared device/host definitions for the B200 FD-PDE inpainting kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hmi_b200.h"

namespace hmi {

// Arguments common to every sweep kernel.  All row indices are slab-local: row 0 is the first
// interior row, negative rows are ghost rows above, rows >= `rows` are ghost rows below.
// `src`/`dst`/`mask` point at interior row 0.
template <typename T>
struct SweepArgs {
    const T *src;
    T *dst;
    const uint8_t *mask;       // non-zero = known
    long long pitch;           // elements per row of src/dst
    long long mpitch;          // bytes per row of mask
    int rows, cols;            // interior rows of the slab, columns
    int row_lo, row_hi;        // rows updated by this launch [row_lo, row_hi)
    int top_border, bottom_border;  // 1 = the slab edge is the global grid edge (border rule applies)
    int border;                // hmi_border
    int s;                     // orientation: +1 correlation, -1 convolution
    T dt, tension, one_minus_tension, eps2, relax_w, one_minus_relax_w;
    int relax;                 // 1 = over-relaxation active (relaxation > 1)
    int do_norm;               // 1 = reduce the convergence sums of this sweep
    double *partials;          // [gridDim][4] per-block partial sums
    unsigned int *counter;     // blocks-done counter (self-resetting)
    double *result;            // [4] sum_d2, sum_f2, max_abs, has_nan
};

// Non-contracting arithmetic: the generic kernel keeps every multiply and add a separate IEEE
// rounding so it reproduces the NumPy expression order of the reference bit for bit.
template <typename T> struct Op;
template <> struct Op<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
};
template <> struct Op<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Fused convergence reduction: per-thread values -> warp shuffle -> block (shared memory) ->
// one partial per block -> the last block to finish folds all partials in index order, so the
// result does not depend on block scheduling.  Must be called by every thread of the block.
// `sh` is block-shared scratch of at least 4*32 doubles.
__device__ __forceinline__ void block_norm_reduce(double d2, double f2, double mx, double nan_flag,
                                                  double *sh, double *partials,
                                                  unsigned int *counter, double *result)
{
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    const int lane = tid & 31, warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
    d2 = warp_sum(d2);
    f2 = warp_sum(f2);
    mx = warp_max(mx);
    nan_flag = warp_max(nan_flag);
    if (lane == 0) {
        sh[warp] = d2; sh[32 + warp] = f2; sh[64 + warp] = mx; sh[96 + warp] = nan_flag;
    }
    __syncthreads();
    __shared__ bool is_last;
    const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (warp == 0) {
        double a = lane < nwarps ? sh[lane] : 0.0;
        double b = lane < nwarps ? sh[32 + lane] : 0.0;
        double c = lane < nwarps ? sh[64 + lane] : 0.0;
        double d = lane < nwarps ? sh[96 + lane] : 0.0;
        a = warp_sum(a); b = warp_sum(b); c = warp_max(c); d = warp_max(d);
        if (lane == 0) {
            partials[4 * bid + 0] = a; partials[4 * bid + 1] = b;
            partials[4 * bid + 2] = c; partials[4 * bid + 3] = d;
            __threadfence();
            const unsigned int prev = atomicAdd(counter, 1u);
            is_last = (prev == nblocks - 1);
        }
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // fixed-order fold: thread t sums partials t, t+n, t+2n ... then a fixed tree
        double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
        for (unsigned int q = tid; q < nblocks; q += nthreads) {
            a += __ldcg(&partials[4 * q + 0]); b += __ldcg(&partials[4 * q + 1]);
            c = fmax(c, __ldcg(&partials[4 * q + 2])); d = fmax(d, __ldcg(&partials[4 * q + 3]));
        }
        a = warp_sum(a); b = warp_sum(b); c = warp_max(c); d = warp_max(d);
        __syncthreads();
        if (lane == 0) { sh[warp] = a; sh[32 + warp] = b; sh[64 + warp] = c; sh[96 + warp] = d; }
        __syncthreads();
        if (warp == 0) {
            a = lane < nwarps ? sh[lane] : 0.0;
            b = lane < nwarps ? sh[32 + lane] : 0.0;
            c = lane < nwarps ? sh[64 + lane] : 0.0;
            d = lane < nwarps ? sh[96 + lane] : 0.0;
            a = warp_sum(a); b = warp_sum(b); c = warp_max(c); d = warp_max(d);
            if (lane == 0) {
                result[0] = a; result[1] = b; result[2] = c; result[3] = d;
                *counter = 0u;  // ready for the next check
            }
        }
    }
}

// Stencil radius of one sweep (rows of ghost validity a sweep consumes).
__host__ __device__ inline int method_radius(int method) { return method == HMI_CCST ? 2 : 1; }

// Launchers implemented in the .cu files.  Each returns the cudaError_t of the launch.
template <typename T>
cudaError_t launch_generic(int method, const SweepArgs<T> &a, cudaStream_t st, int *nblocks_out);
int generic_max_blocks(int rows_total, int cols);

// one evaluation of the step function on a whole grid (a.mask = work mask or null; a.dst = step)
template <typename T>
cudaError_t launch_generic_step(int method, const SweepArgs<T> &a, cudaStream_t st);

// one 3x3 filter (Convolver.__call__)
struct Kernel9 { double w[9]; };
template <typename T>
cudaError_t launch_conv3x3(const T *src, T *dst, long long pitch, const uint8_t *mask, long long mpitch, int rows, int cols,
                           int border, int s, const Kernel9 &k, cudaStream_t st);

}  // namespace hmi
