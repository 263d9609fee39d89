// Single-operator entry points of the C ABI (include/hmi_b200.h): the reference's host-side hooks
// step_fun(f, mask) (fd_pde_inpainter.py:406-408 and the four subclasses) and Convolver.__call__
// (convolver.py:28-29), evaluated on the device by the generic kernels of hmi_generic.cu.  They exist
// so that code written against the reference's hooks keeps working; the solver itself never calls
// them (its step functions are fused into the sweep kernels).
#include "hmi_internal.cuh"
#include "hmi_xfer.cuh"

using hmi::fail;

namespace {

struct DevGrid {
    int device = 0;
    char *in = nullptr, *out = nullptr;
    uint8_t *mask = nullptr;
    size_t gbytes = 0, mbytes = 0;
    long long pitch = 0, mpitch = 0;
    cudaStream_t st = nullptr;
    ~DevGrid()
    {
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
        hmi::cached_free(device, in, gbytes);
        hmi::cached_free(device, out, gbytes);
        hmi::cached_free(device, mask, mbytes);
    }
};

int check_device(int device)
{
    int ndev = 0;
    const cudaError_t e = hmi::device_count_patiently(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(HMI_ERR_NO_DEVICE, "no CUDA device available; this library has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(HMI_ERR_BAD_ARG, "device %d out of range (%d devices)", device, ndev);
    return HMI_OK;
}

// upload `image` (and `mask`, if any) into freshly sized device grids
int stage_in(DevGrid &d, int device, int rows, int cols, size_t elem, const void *image, size_t pitch_b, const uint8_t *mask,
             size_t mpitch_b)
{
    int rc = check_device(device);
    if (rc != HMI_OK) return rc;
    if (rows < 1 || cols < 1) return fail(HMI_ERR_BAD_ARG, "empty grid");
    if (pitch_b < (size_t)cols * elem || (mask && mpitch_b < (size_t)cols)) return fail(HMI_ERR_BAD_ARG, "pitch smaller than a row");
    CUDA_TRY(cudaSetDevice(device));
    d.device = device;
    d.pitch = hmi::round_up(cols, 32);
    d.mpitch = hmi::round_up(cols, 128);
    d.gbytes = (size_t)rows * d.pitch * elem;
    d.mbytes = (size_t)rows * d.mpitch;
    CUDA_TRY(cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking));
    CUDA_TRY(hmi::cached_malloc(device, (void **)&d.in, d.gbytes));
    CUDA_TRY(hmi::cached_malloc(device, (void **)&d.out, d.gbytes));
    CUDA_TRY(hmi::copy_h2d_2d(device, d.in, (size_t)d.pitch * elem, image, pitch_b, (size_t)cols * elem, rows, d.st));
    if (mask) {
        CUDA_TRY(hmi::cached_malloc(device, (void **)&d.mask, d.mbytes));
        CUDA_TRY(hmi::copy_h2d_2d(device, d.mask, (size_t)d.mpitch, mask, mpitch_b, (size_t)cols, rows, d.st));
    }
    return HMI_OK;
}

}  // namespace

extern "C" {

int hmi_step_fun_host(const hmi_params *params, const void *f, size_t f_pitch_bytes, const uint8_t *work_mask,
                      size_t mask_pitch_bytes, void *out, size_t out_pitch_bytes)
{
    if (!params || !f || !out) return fail(HMI_ERR_BAD_ARG, "null argument");
    if (params->struct_size != (int32_t)sizeof(hmi_params)) return fail(HMI_ERR_BAD_ARG, "hmi_params size mismatch");
    const hmi_params &p = *params;
    if (p.dtype != HMI_F32 && p.dtype != HMI_F64) return fail(HMI_ERR_BAD_ARG, "bad dtype");
    if (p.method < HMI_HARMONIC || p.method > HMI_AMLE) return fail(HMI_ERR_BAD_ARG, "bad method");
    if (p.orientation != 1 && p.orientation != -1) return fail(HMI_ERR_BAD_ARG, "orientation must be +1/-1");
    if (p.border < HMI_BORDER_REFLECT101 || p.border > HMI_BORDER_ZERO) return fail(HMI_ERR_BAD_ARG, "bad border");
    const size_t elem = p.dtype == HMI_F64 ? 8 : 4;
    if (out_pitch_bytes < (size_t)p.cols * elem) return fail(HMI_ERR_BAD_ARG, "pitch smaller than a row");
    DevGrid d;
    int rc = stage_in(d, p.device, p.rows, p.cols, elem, f, f_pitch_bytes, work_mask, mask_pitch_bytes);
    if (rc != HMI_OK) return rc;
    cudaError_t e;
    if (p.dtype == HMI_F64) {
        hmi::SweepArgs<double> a = {};
        a.src = (const double *)d.in; a.dst = (double *)d.out; a.mask = d.mask;
        a.pitch = d.pitch; a.mpitch = d.mpitch; a.rows = p.rows; a.cols = p.cols;
        a.top_border = a.bottom_border = 1; a.border = p.border; a.s = p.orientation;
        a.tension = p.tension; a.one_minus_tension = 1.0 - p.tension; a.eps2 = p.epsilon * p.epsilon;
        e = hmi::launch_generic_step<double>(p.method, a, d.st);
    } else {
        hmi::SweepArgs<float> a = {};
        a.src = (const float *)d.in; a.dst = (float *)d.out; a.mask = d.mask;
        a.pitch = d.pitch; a.mpitch = d.mpitch; a.rows = p.rows; a.cols = p.cols;
        a.top_border = a.bottom_border = 1; a.border = p.border; a.s = p.orientation;
        a.tension = (float)p.tension; a.one_minus_tension = (float)(1.0 - p.tension);
        a.eps2 = (float)(p.epsilon * p.epsilon);
        e = hmi::launch_generic_step<float>(p.method, a, d.st);
    }
    if (e != cudaSuccess) return fail(HMI_ERR_CUDA, "step kernel launch failed: %s", cudaGetErrorString(e));
    CUDA_TRY(hmi::copy_d2h_2d(p.device, out, out_pitch_bytes, d.out, (size_t)d.pitch * elem, (size_t)p.cols * elem, p.rows, d.st));
    return HMI_OK;
}

int hmi_convolve3x3_host(int32_t rows, int32_t cols, int32_t dtype, int32_t device, const void *image,
                         size_t image_pitch_bytes, const double *kernel9, int32_t orientation, int32_t border,
                         const uint8_t *mask, size_t mask_pitch_bytes, void *out, size_t out_pitch_bytes)
{
    if (!image || !kernel9 || !out) return fail(HMI_ERR_BAD_ARG, "null argument");
    if (dtype != HMI_F32 && dtype != HMI_F64) return fail(HMI_ERR_BAD_ARG, "bad dtype");
    if (orientation != 1 && orientation != -1) return fail(HMI_ERR_BAD_ARG, "orientation must be +1/-1");
    if (border < HMI_BORDER_REFLECT101 || border > HMI_BORDER_ZERO) return fail(HMI_ERR_BAD_ARG, "bad border");
    const size_t elem = dtype == HMI_F64 ? 8 : 4;
    if (out_pitch_bytes < (size_t)cols * elem) return fail(HMI_ERR_BAD_ARG, "pitch smaller than a row");
    DevGrid d;
    int rc = stage_in(d, device, rows, cols, elem, image, image_pitch_bytes, mask, mask_pitch_bytes);
    if (rc != HMI_OK) return rc;
    hmi::Kernel9 k;
    for (int i = 0; i < 9; ++i) k.w[i] = kernel9[i];
    cudaError_t e;
    if (dtype == HMI_F64)
        e = hmi::launch_conv3x3<double>((const double *)d.in, (double *)d.out, d.pitch, d.mask, d.mpitch, rows, cols, border,
                                        orientation, k, d.st);
    else
        e = hmi::launch_conv3x3<float>((const float *)d.in, (float *)d.out, d.pitch, d.mask, d.mpitch, rows, cols, border,
                                       orientation, k, d.st);
    if (e != cudaSuccess) return fail(HMI_ERR_CUDA, "filter kernel launch failed: %s", cudaGetErrorString(e));
    CUDA_TRY(hmi::copy_d2h_2d(device, out, out_pitch_bytes, d.out, (size_t)d.pitch * elem, (size_t)cols * elem, rows, d.st));
    return HMI_OK;
}

}  // extern "C"
