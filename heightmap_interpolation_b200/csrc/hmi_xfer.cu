// Staged host <-> device copies (see hmi_xfer.cuh).  The reference hands the solver NumPy arrays
// (heightmap_interpolation/apps/interpolate_netcdf4.py:188-204), i.e. pageable memory: the stock
// cudaMemcpy path moved a 4096^2 fp64 grid at ~10 GB/s (13.5 ms up, 8-14 ms down,
// profiles/r02_e2e_breakdown.log) against 2.9 / 2.4 ms from pinned buffers.
#include "hmi_xfer.cuh"

#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace hmi {

namespace {

constexpr int kThreads = 6;                       // host threads copying through one device's staging ring (at most)
constexpr size_t kChunk = (size_t)4 << 20;        // bytes per staging buffer (two per thread)
constexpr size_t kMinStaged = (size_t)8 << 20;    // smaller copies are left to the driver
constexpr int kMaxDevices = 64;

// one pinned ring and one "busy" lock per device, so that the slabs of a multi-GPU solve are
// uploaded / downloaded concurrently (one staged copy at a time per device)
std::mutex g_mu;
char *g_stage[kMaxDevices] = {nullptr};           // kThreads * 2 * kChunk pinned bytes each, allocated on first use
bool g_stage_failed = false;
std::mutex g_busy[kMaxDevices];
int g_threads = kThreads;                         // threads per staged copy (set_xfer_threads)

char *staging(int device)
{
    if (device < 0 || device >= kMaxDevices) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_stage[device] && !g_stage_failed) {
        void *p = nullptr;
        if (cudaHostAlloc(&p, (size_t)kThreads * 2 * kChunk, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            g_stage_failed = true;
        } else {
            g_stage[device] = (char *)p;
        }
    }
    return g_stage[device];
}

bool is_pageable(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// one worker: chunks t, t + kThreads, ... of `nchunks`, each `rpc` rows (the last one shorter)
void worker(int t, int nthreads, int device, bool to_device, char *dev, size_t dev_pitch, char *host, size_t host_pitch,
            size_t width, size_t rows, size_t rpc, char *stage, cudaError_t *err)
{
    *err = cudaSetDevice(device);
    if (*err != cudaSuccess) return;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    if ((*err = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess) return;
    for (int b = 0; b < 2 && *err == cudaSuccess; ++b) *err = cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming);
    const size_t nchunks = (rows + rpc - 1) / rpc;
    auto stage_buf = [&](int b) { return stage + ((size_t)t * 2 + b) * kChunk; };
    auto host_copy = [&](char *dst, size_t dp, const char *src, size_t sp, size_t n) {
        if (dp == width && sp == width) { memcpy(dst, src, n * width); return; }
        for (size_t r = 0; r < n; ++r) memcpy(dst + r * dp, src + r * sp, width);
    };
    if (*err == cudaSuccess && to_device) {
        int b = 0;
        for (size_t c = t; c < nchunks && *err == cudaSuccess; c += nthreads, b ^= 1) {
            const size_t r0 = c * rpc, n = (r0 + rpc <= rows) ? rpc : rows - r0;
            if ((*err = cudaEventSynchronize(ev[b])) != cudaSuccess) break;       // buffer free again?
            host_copy(stage_buf(b), width, host + r0 * host_pitch, host_pitch, n);
            *err = cudaMemcpy2DAsync(dev + r0 * dev_pitch, dev_pitch, stage_buf(b), width, width, n,
                                     cudaMemcpyHostToDevice, st);
            if (*err == cudaSuccess) *err = cudaEventRecord(ev[b], st);
        }
    } else if (*err == cudaSuccess) {
        // device -> staging of chunk i+1 is in flight while chunk i is copied out to the caller's array
        auto issue = [&](size_t c, int b) {
            const size_t r0 = c * rpc, n = (r0 + rpc <= rows) ? rpc : rows - r0;
            cudaError_t e = cudaMemcpy2DAsync(stage_buf(b), width, dev + r0 * dev_pitch, dev_pitch, width, n,
                                              cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaEventRecord(ev[b], st);
            return e;
        };
        int b = 0;
        if ((size_t)t < nchunks) *err = issue(t, 0);
        for (size_t c = t; c < nchunks && *err == cudaSuccess; c += nthreads, b ^= 1) {
            if (c + nthreads < nchunks) *err = issue(c + nthreads, b ^ 1);
            if (*err != cudaSuccess) break;
            if ((*err = cudaEventSynchronize(ev[b])) != cudaSuccess) break;
            const size_t r0 = c * rpc, n = (r0 + rpc <= rows) ? rpc : rows - r0;
            host_copy(host + r0 * host_pitch, host_pitch, stage_buf(b), width, n);
        }
    }
    if (st) {
        const cudaError_t e = cudaStreamSynchronize(st);
        if (*err == cudaSuccess) *err = e;
    }
    for (int b = 0; b < 2; ++b) if (ev[b]) cudaEventDestroy(ev[b]);
    if (st) cudaStreamDestroy(st);
}

cudaError_t staged(int device, bool to_device, void *dev, size_t dev_pitch, void *host, size_t host_pitch,
                   size_t width, size_t rows, cudaStream_t stream)
{
    char *stage = staging(device);
    if (!stage || width > kChunk) return cudaErrorNotSupported;
    std::lock_guard<std::mutex> lk(g_busy[device]);
    cudaError_t e = cudaStreamSynchronize(stream);  // the copy runs on its own streams, after the caller's work
    if (e != cudaSuccess) return e;
    size_t rpc = kChunk / width;
    if (rpc < 1) rpc = 1;
    const int nt = g_threads < 1 ? 1 : (g_threads > kThreads ? kThreads : g_threads);
    cudaError_t errs[kThreads];
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t)
        th.emplace_back(worker, t, nt, device, to_device, (char *)dev, dev_pitch, (char *)host, host_pitch, width, rows,
                        rpc, stage, &errs[t]);
    worker(0, nt, device, to_device, (char *)dev, dev_pitch, (char *)host, host_pitch, width, rows, rpc, stage, &errs[0]);
    for (auto &x : th) x.join();
    for (int t = 0; t < nt; ++t)
        if (errs[t] != cudaSuccess) return errs[t];
    return cudaSuccess;
}

}  // namespace

void set_xfer_threads(int n) { g_threads = n < 1 ? 1 : (n > kThreads ? kThreads : n); }

cudaError_t copy_h2d_2d(int device, void *dst_dev, size_t dpitch, const void *src_host, size_t spitch, size_t width,
                        size_t rows, cudaStream_t stream)
{
    if (rows == 0 || width == 0) return cudaSuccess;
    if (width * rows >= kMinStaged && is_pageable(src_host)) {
        const cudaError_t e = staged(device, true, dst_dev, dpitch, const_cast<void *>(src_host), spitch, width, rows,
                                     stream);
        if (e != cudaErrorNotSupported) return e;
    }
    cudaError_t e;
    if (dpitch == width && spitch == width)
        e = cudaMemcpyAsync(dst_dev, src_host, width * rows, cudaMemcpyHostToDevice, stream);
    else
        e = cudaMemcpy2DAsync(dst_dev, dpitch, src_host, spitch, width, rows, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);
}

cudaError_t copy_d2h_2d(int device, void *dst_host, size_t dpitch, const void *src_dev, size_t spitch, size_t width,
                        size_t rows, cudaStream_t stream)
{
    if (rows == 0 || width == 0) return cudaSuccess;
    if (width * rows >= kMinStaged && is_pageable(dst_host)) {
        const cudaError_t e = staged(device, false, const_cast<void *>(src_dev), spitch, dst_host, dpitch, width, rows,
                                     stream);
        if (e != cudaErrorNotSupported) return e;
    }
    cudaError_t e;
    if (dpitch == width && spitch == width)
        e = cudaMemcpyAsync(dst_host, src_dev, width * rows, cudaMemcpyDeviceToHost, stream);
    else
        e = cudaMemcpy2DAsync(dst_host, dpitch, src_dev, spitch, width, rows, cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);
}

}  // namespace hmi
