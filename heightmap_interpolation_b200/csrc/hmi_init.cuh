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

}  // namespace hmi
