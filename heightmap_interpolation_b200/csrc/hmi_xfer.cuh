// Host <-> device copies of pitched 2-D arrays that are fast for *pageable* host memory too.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

namespace hmi {

// cudaMemcpy2D semantics (synchronous on return), ordered after the work already queued on `stream`.  Pinned / registered host memory goes to the DMA
// engine directly (one linear copy when both sides are dense).  Pageable memory — what a NumPy array
// is — is staged by a few host threads through pinned ring buffers, so the page-by-page copy the
// driver would do on one thread (~10 GB/s) runs on several and overlaps the DMA.
cudaError_t copy_h2d_2d(int device, void *dst_dev, size_t dpitch, const void *src_host, size_t spitch, size_t width,
                        size_t rows, cudaStream_t stream = nullptr);
cudaError_t copy_d2h_2d(int device, void *dst_host, size_t dpitch, const void *src_dev, size_t spitch, size_t width,
                        size_t rows, cudaStream_t stream = nullptr);

// Host threads per staged copy (1..6, default 6).  A multi-GPU solve that moves its slabs
// concurrently lowers it so that the copies of all devices together do not oversubscribe the host.
void set_xfer_threads(int n);

}  // namespace hmi
