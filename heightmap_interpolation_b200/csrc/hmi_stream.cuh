// Interface of the streaming temporal-blocked kernel (hmi_stream.cu).
#pragma once

#include <cuda.h>

#include "hmi_common.cuh"

namespace hmi {

constexpr int STREAM_MAX_K_HARMONIC = 8;
constexpr int STREAM_MAX_K_CCST = 4;

template <typename T>
struct alignas(64) StreamArgs {
    CUtensorMap tmap_f, tmap_m;  // tma == 2: 2-D tensor maps of the whole src buffer and of the mask (boxes of 32*V x 3)
    const T *src;
    T *dst;
    const uint8_t *mask;
    long long pitch, mpitch;
    int rows, cols;             // interior rows of the slab, columns
    int row_lo, row_hi;         // rows written by this launch [row_lo, row_hi)
    int read_lo, read_hi;       // rows that exist in memory [read_lo, read_hi) (ghosts included)
    int top_border, bottom_border;
    int chunk_rows;             // rows per warp chunk; 0 = auto (wave-aware), < 0 = legacy auto rule
    int vec;                    // cells per lane (fp64: 2 or 4, pipeline kernel 4 or 8; 0 = default)
    int tma;                    // row staging: 0 = cp.async ring, 1 = 1-D bulk copies + mbarriers, 2 = 2-D tensor-map boxes, 3 = pipeline kernel
    int phys_row0;              // tma == 2: tensor-map row of slab row 0 (= ghost_top)
    int border;                 // HMI_BORDER_REFLECT101 or HMI_BORDER_ZERO (zero: default vec, cp.async staging)
    int nwin, nchunks;          // filled by the launcher
    int n_edge, nchunks_e, chunk_rows_e;   // column-edge strips: how many (listed first), their chunking
    T dt, tension, one_minus_tension;
    int do_norm;
    double *partials;
    unsigned int *counter;
    double *result;
};

// k sweeps fused in one launch.  method in {HMI_HARMONIC, HMI_CCST}, 1 <= k <= stream_max_k().
template <typename T>
cudaError_t launch_stream(int method, int k, const StreamArgs<T> &a, cudaStream_t st, int *nblocks_out);

// The same k fused sweeps on the warp-specialised pipeline kernel (hmi_pipe.cu; tma == 3): one loader warp and one
// warp per sweep, rows handed on through shared-memory rings.  No convergence sums; k >= 2.
template <typename T>
cudaError_t launch_pipe(int method, int k, const StreamArgs<T> &a, cudaStream_t st, int *nblocks_out);
bool pipe_supports(int method, int k);
bool stream_zero_supports(int method, int k);       // fused-sweep depths instantiated for zero padding

int stream_max_k(int method, int dtype);
int stream_halo(int method, int k);                 // ghost rows/columns k fused sweeps consume
int stream_auto_chunk_rows(int nrows, int nwin, int halo);      // legacy rule (chunk_rows < 0)
int stream_cta_slots(const void *kernel, int threads, size_t smem);   // CTAs resident on the GPU at once
void stream_pick_chunks(int nrows, int nwin, int n_edge, int halo, int wpb, int cta_slots, double edge_cost,
                        int *chunk_rows, int *chunk_rows_edge, int min_chunk = 48);                  // wave-aware (chunk_rows == 0)
int stream_max_blocks(int method, int rows_total, int cols);
int stream_cells_per_lane(int dtype, int vec);       // V of the instantiation launch_stream picks
constexpr int STREAM_BOX_ROWS = 3;                   // rows per tensor-map box (tma == 2)

}  // namespace hmi
