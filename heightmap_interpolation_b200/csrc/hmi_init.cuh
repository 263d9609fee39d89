// Interface of the device-side 'nearest' initialiser (hmi_init.cu).
#pragma once

#include "hmi_common.cuh"

namespace hmi {

// out[i][j] = img[i][j] where mask != 0, else the value of the Euclidean-nearest known cell.
// `off` is scratch of rows x opitch ints; *status becomes 1 if the grid has no known cell at all.
// img and out must not alias.
template <typename T>
cudaError_t launch_init_nearest(const T *img, long long pitch, const uint8_t *mask, long long mpitch, int rows,
                                int cols, int *off, long long opitch, T *out, long long out_pitch, int *status,
                                cudaStream_t st);

// Multigrid: f[i][j] <- up(i, j) * (1 - m) + f[i][j] * m with up = bilinear interpolation of `coarse`
// from host-built tap/weight tables (cv2.resize INTER_LINEAR); optionally NaNs of f are zeroed first
// (level 0, fd_pde_inpainter.py:259-260; *nan_flag is set if there were any).  One (min, max, nan)
// triple per block is written to `partials` (at most UPSAMPLE_MAX_BLOCKS(cols) blocks).
constexpr int UPSAMPLE_MAX_GRID_Y = 64;
inline int upsample_max_blocks(int cols) { return ((cols + 127) / 128) * UPSAMPLE_MAX_GRID_Y; }

template <typename T>
cudaError_t launch_upsample_blend(T *f, long long pitch, const uint8_t *mask, long long mpitch, int rows, int cols,
                                  const T *coarse, long long cpitch, const int *sx, const int *sx1, const T *ax0,
                                  const T *ax1, const int *sy, const int *sy1, const T *by0, const T *by1,
                                  int scrub_nan, double *partials, int *nan_flag, int *nblocks, cudaStream_t st);

// A grid cut into row slabs that may live on several devices (all readable from the launching one):
// slab k holds global rows [row0[k], row0[k+1]); base[k] points at its row row0[k].
constexpr int MAX_VIEW_SLABS = 16;
struct SlabView {
    int n;
    int row0[MAX_VIEW_SLABS + 1];
    const char *base[MAX_VIEW_SLABS];
    long long pitch;                 // bytes per row
};

// One slab of one pyramid level: image and mask gathered with `stride` from the finest level (stride 1 =
// the finest level itself, blended in place), unknown cells <- bilinear up-sampling of the coarser solution.
template <typename T>
struct PyramidBlend {
    T *f;                            // this slab, physical row 0 (ghost rows included)
    uint8_t *mask;
    long long pitch, mpitch;         // elements / bytes per row
    int phys_rows, cols;
    int row_first;                   // row of the level's grid that physical row 0 is
    int stride, scrub_nan;
    SlabView fine_img, fine_mask, coarse;
    const int *sx, *sx1, *sy, *sy1;  // taps: per column, per row of the level's grid
    const T *ax0, *ax1, *by0, *by1;
    double *partials;                // (min, max, nan) per block
    int *nan_flag;
};
template <typename T>
cudaError_t launch_pyramid_blend(const PyramidBlend<T> &a, int *nblocks, cudaStream_t st);

}  // namespace hmi
