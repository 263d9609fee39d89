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

}  // namespace hmi
