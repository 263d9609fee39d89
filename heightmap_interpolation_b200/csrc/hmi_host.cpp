// Host-side initialiser helpers of the C ABI (include/hmi_b200.h): the 'zeros' / 'mean' fillers of
// the reference's Initializer (heightmap_interpolation/inpainting/initializer.py:15-18), done in
// place on the caller's host array with a few threads and branch-free inner loops.
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/hmi_b200.h"

namespace {

int num_threads(int64_t n)
{
    // one thread per million cells (starting a thread costs about as much as filling 200k cells)
    const unsigned hw = std::thread::hardware_concurrency();
    const int64_t cap = hw == 0 ? 4 : (hw > 16 ? 16 : (int64_t)hw);
    const int64_t want = n >> 20;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

template <typename T>
void fill_range(T *__restrict__ img, const uint8_t *__restrict__ mask, int64_t a, int64_t b, T v)
{
    for (int64_t i = a; i < b; ++i) img[i] = mask[i] ? img[i] : v;
}

template <typename T>
void fill_unknown(T *img, const uint8_t *mask, int64_t n, T v)
{
    const int nt = num_threads(n);
    const int64_t per = (n + nt - 1) / nt;
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) {
        const int64_t a = t * per, b = (a + per < n) ? a + per : n;
        if (a < b) th.emplace_back(fill_range<T>, img, mask, a, b, v);
    }
    fill_range<T>(img, mask, 0, per < n ? per : n, v);
    for (auto &x : th) x.join();
}

template <typename T>
void sum_range(const T *__restrict__ img, const uint8_t *__restrict__ mask, int64_t a, int64_t b,
               double *sum, int64_t *cnt)
{
    double s = 0.0;
    int64_t c = 0;
    for (int64_t i = a; i < b; ++i) {
        s += mask[i] ? (double)img[i] : 0.0;
        c += mask[i] ? 1 : 0;
    }
    *sum = s;
    *cnt = c;
}

template <typename T>
double known_mean(const T *img, const uint8_t *mask, int64_t n)
{
    const int nt = num_threads(n);
    const int64_t per = (n + nt - 1) / nt;
    std::vector<double> sums(nt, 0.0);
    std::vector<int64_t> cnts(nt, 0);
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) {
        const int64_t a = t * per, b = (a + per < n) ? a + per : n;
        if (a < b) th.emplace_back(sum_range<T>, img, mask, a, b, &sums[t], &cnts[t]);
    }
    sum_range<T>(img, mask, 0, per < n ? per : n, &sums[0], &cnts[0]);
    for (auto &x : th) x.join();
    double s = 0.0;
    int64_t c = 0;
    for (int t = 0; t < nt; ++t) { s += sums[t]; c += cnts[t]; }
    return c ? s / (double)c : std::nan("");
}

// min / max over the known cells, NaN-propagating like np.max / np.min
template <typename T>
void minmax_range(const T *__restrict__ img, const uint8_t *__restrict__ mask, int64_t a, int64_t b,
                  double *mn, double *mx, int64_t *cnt, int *has_nan)
{
    double lo = INFINITY, hi = -INFINITY;
    int64_t c = 0;
    int nanf = 0;
    for (int64_t i = a; i < b; ++i) {
        if (!mask[i]) continue;
        const double v = (double)img[i];
        if (v != v) nanf = 1;
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
        ++c;
    }
    *mn = lo; *mx = hi; *cnt = c; *has_nan = nanf;
}

template <typename T>
int64_t known_minmax(const T *img, const uint8_t *mask, int64_t n, double *mn, double *mx)
{
    const int nt = num_threads(n);
    const int64_t per = (n + nt - 1) / nt;
    std::vector<double> los(nt, INFINITY), his(nt, -INFINITY);
    std::vector<int64_t> cnts(nt, 0);
    std::vector<int> nans(nt, 0);
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) {
        const int64_t a = t * per, b = (a + per < n) ? a + per : n;
        if (a < b) th.emplace_back(minmax_range<T>, img, mask, a, b, &los[t], &his[t], &cnts[t], &nans[t]);
    }
    minmax_range<T>(img, mask, 0, per < n ? per : n, &los[0], &his[0], &cnts[0], &nans[0]);
    for (auto &x : th) x.join();
    double lo = INFINITY, hi = -INFINITY;
    int64_t c = 0;
    int nanf = 0;
    for (int t = 0; t < nt; ++t) {
        lo = los[t] < lo ? los[t] : lo;
        hi = his[t] > hi ? his[t] : hi;
        c += cnts[t];
        nanf |= nans[t];
    }
    if (nanf) lo = hi = std::nan("");
    *mn = lo;
    *mx = hi;
    return c;
}

}  // namespace

extern "C" {

int hmi_host_known_minmax(const void *image, const uint8_t *mask, int64_t n, int32_t dtype, double *min_out,
                          double *max_out, int64_t *known_out)
{
    if (!image || !mask || !min_out || !max_out || !known_out || n < 0) return HMI_ERR_BAD_ARG;
    if (dtype == HMI_F64) *known_out = known_minmax<double>((const double *)image, mask, n, min_out, max_out);
    else if (dtype == HMI_F32) *known_out = known_minmax<float>((const float *)image, mask, n, min_out, max_out);
    else return HMI_ERR_BAD_ARG;
    return HMI_OK;
}

int hmi_host_fill_unknown(void *image, const uint8_t *mask, int64_t n, int32_t dtype, double value)
{
    if (!image || !mask || n < 0) return HMI_ERR_BAD_ARG;
    if (dtype == HMI_F64) fill_unknown<double>((double *)image, mask, n, value);
    else if (dtype == HMI_F32) fill_unknown<float>((float *)image, mask, n, (float)value);
    else return HMI_ERR_BAD_ARG;
    return HMI_OK;
}

int hmi_host_known_mean(const void *image, const uint8_t *mask, int64_t n, int32_t dtype, double *mean)
{
    if (!image || !mask || !mean || n < 0) return HMI_ERR_BAD_ARG;
    if (dtype == HMI_F64) *mean = known_mean<double>((const double *)image, mask, n);
    else if (dtype == HMI_F32) *mean = known_mean<float>((const float *)image, mask, n);
    else return HMI_ERR_BAD_ARG;
    return HMI_OK;
}

int hmi_host_prefault(void *buf, size_t bytes)
{
    // touch one byte per page from several threads, so a freshly allocated result array does not
    // take its page faults one by one inside the device-to-host copy
    if (!buf) return HMI_ERR_BAD_ARG;
    const int64_t n = (int64_t)bytes;
    const int nt = num_threads(n);
    const int64_t per = ((n + nt - 1) / nt + 4095) / 4096 * 4096;
    auto touch = [](volatile unsigned char *p, int64_t a, int64_t b) {
        for (int64_t i = a; i < b; i += 4096) p[i] = 0;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) {
        const int64_t a = t * per, b = (a + per < n) ? a + per : n;
        if (a < b) th.emplace_back(touch, (volatile unsigned char *)buf, a, b);
    }
    touch((volatile unsigned char *)buf, 0, per < n ? per : n);
    for (auto &x : th) x.join();
    return HMI_OK;
}

}  // extern "C"
