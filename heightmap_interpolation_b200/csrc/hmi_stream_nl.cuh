// Interface of the streaming kernel for TV and AMLE (hmi_stream_nl.cu).
#pragma once

#include "hmi_common.cuh"

namespace hmi {

// What the host passes (actual coordinates)
template <typename T>
struct NLLaunch {
    const T *src;
    T *dst;
    const uint8_t *mask;
    long long pitch, mpitch;
    int rows, cols;
    int row_lo, row_hi;          // rows written [row_lo, row_hi)
    int read_lo, read_hi;        // rows that exist in memory
    int top_border, bottom_border;
    int method;                  // HMI_TV or HMI_AMLE
    int sgn;                     // +1 correlation ("opencv"), -1 convolution ("masked*")
    int border;                  // hmi_border (zero: sgn -1 only; symmetric: AMLE with sgn -1 only)
    int chunk_rows;              // 0 = auto (wave-aware), < 0 = legacy auto rule
    int k;                       // sweeps fused per launch (1..2, fp32 1..3; the norm needs 1)
    int do_norm;
    T dt, eps2;
    double *partials;
    unsigned int *counter;
    double *result;
};

// What the kernel sees (canonical coordinates: identity for sgn = +1, point mirror for sgn = -1)
template <typename T>
struct StreamNLArgs {
    const T *src;
    T *dst;
    const uint8_t *mask;
    long long pitch, mpitch;
    int rows;                    // interior rows of the slab (for the row mirror)
    int clo, chi, cbase;         // canonical column range of the grid, first strip's origin
    int row_lo, row_hi, read_lo, read_hi, top_border, bottom_border;   // canonical
    int chunk_rows, nwin, nchunks;
    int n_edge, nchunks_e, chunk_rows_e;   // column-edge strips: how many (listed first), their chunking
    T dt, eps2;
    double *partials;
    unsigned int *counter;
    double *result;
};

template <typename T>
cudaError_t launch_stream_nl(const NLLaunch<T> &l, cudaStream_t st, int *nblocks_out);

int stream_nl_auto_chunk_rows(int nrows, int nwin);
int stream_nl_max_k(int method, int dtype, int border = HMI_BORDER_REFLECT101);   // 2; TV fp32 reflect-101 (four cells of lane halo): 3
bool stream_nl_supports(int method, int sgn, int border);

}  // namespace hmi
