/*
 * TEST INFRASTRUCTURE ONLY — CPU oracle for the FD-PDE inpainting hot path.
 *
 * This header is included twice by fdpde_oracle.c, once with T=double (SUF=f64)
 * and once with T=float (SUF=f32).  It restates, in plain C, the algorithm of
 * the reference's explicit finite-difference PDE inpainters.  Citations are
 * relative to /root/reference/heightmap_interpolation/inpainting/ :
 *
 *   conv3x3           <- convolver.py:47-69 (which library call each name maps to),
 *                        convolve_at_mask.py:39-63 (masked true convolution, reflect-101
 *                        index folding, identity outside the work mask),
 *                        cv2.filter2D semantics = correlation + BORDER_REFLECT_101
 *   stencil constants <- differential.py:7-36
 *   step_harmonic     <- sobolev_inpainter.py:15-17
 *   step_ccst         <- ccst_inpainter.py:23-34
 *   step_tv           <- tv_inpainter.py:20-33 + differential.py:146-187
 *   step_amle         <- amle_inpainter.py:22-59 (2-D convolver branch)
 *   update_at_mask    <- update_at_mask.py:6-14
 *   term_check        <- fd_pde_inpainter.py:368-383
 *   inpaint_grid      <- fd_pde_inpainter.py:281-366
 *   dilate3x3         <- fd_pde_inpainter.py:293-296 (cv2.dilate with a 3x3 ones kernel)
 *
 * Like the reference, every operator produces a full intermediate array, and
 * the border rule is applied to every array a 3x3 filter is applied to
 * (including intermediates).  Element-wise arithmetic is done in the array
 * dtype T with Python-float scalars narrowed to T (NumPy weak-scalar rule);
 * each 3x3 sum is accumulated in double and rounded to T on store (what the
 * numba kernel does, convolve_at_mask.py:51-62).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * call this code.  It is never on the product path.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* 3x3 filter of `img` into `out`.
 * k[9]      : row-major 3x3 kernel as written in differential.py
 * orient    : +1 correlation (cv2.filter2D), -1 true convolution (numba / SciPy)
 * border    : BORDER_* rule for out-of-range taps
 * work      : NULL = whole grid; else evaluate only where work[i,j] != 0 and
 *             copy img elsewhere (convolve_at_mask.py:46,49)                     */
static void FN(conv3x3)(const T *img, T *out, int rows, int cols, const double k[9],
                        int orient, int border, const uint8_t *work)
{
    /* non-zero taps in the row-major order the numba kernel visits them */
    int nt = 0, ta[9], tb[9];
    double tw[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            const double w = (orient > 0) ? k[a * 3 + b] : k[(2 - a) * 3 + (2 - b)];
            if (w == 0.0) continue;
            ta[nt] = a - 1; tb[nt] = b - 1; tw[nt] = w; ++nt;
        }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < rows; ++i) {
        const int row_inner = (i > 0 && i < rows - 1);
        for (int j = 0; j < cols; ++j) {
            const size_t o = (size_t)i * cols + j;
            if (work && !work[o]) { out[o] = img[o]; continue; }
            double num = 0.0;
            if (row_inner && j > 0 && j < cols - 1) {
                for (int t = 0; t < nt; ++t)
                    num += tw[t] * (double)img[o + (ptrdiff_t)ta[t] * cols + tb[t]];
            } else {
                for (int t = 0; t < nt; ++t) {
                    int inside = 1;
                    const int ii = fold_index(i + ta[t], rows, border, &inside);
                    const int jj = fold_index(j + tb[t], cols, border, &inside);
                    if (!inside) continue; /* zero padding */
                    num += tw[t] * (double)img[(size_t)ii * cols + jj];
                }
            }
            out[o] = (T)num;
        }
    }
}

typedef struct {
    int rows, cols;
    size_t n;
    T *h, *b;                 /* ccst */
    T *g0, *g1, *d0, *d1;     /* tv: gradient pair, divergence terms */
    T *ux, *uy, *uxx, *uxy, *uyx, *uyy, *v0, *v1; /* amle */
} FN(scratch);

static T *FN(alloc)(size_t n) { return (T *)malloc(n * sizeof(T)); }

/* step = L * f                                   sobolev_inpainter.py:17 */
static void FN(step_harmonic)(const T *f, T *step, int rows, int cols, const oracle_params *p,
                              const uint8_t *work)
{
    FN(conv3x3)(f, step, rows, cols, K_LAPLACIAN, p->orientation, p->border, work);
}

/* step = -1*((1-t)*L(L f) - t*L f)               ccst_inpainter.py:31-34 */
static void FN(step_ccst)(const T *f, T *step, FN(scratch) *s, const oracle_params *p,
                          const uint8_t *work)
{
    FN(conv3x3)(f, s->h, s->rows, s->cols, K_LAPLACIAN, p->orientation, p->border, work);
    FN(conv3x3)(s->h, s->b, s->rows, s->cols, K_LAPLACIAN, p->orientation, p->border, work);
    const T one_minus_t = (T)(1.0 - p->tension), t = (T)p->tension;
#pragma omp parallel for schedule(static)
    for (size_t q = 0; q < s->n; ++q) {
        const T x = one_minus_t * s->b[q];
        const T y = t * s->h[q];
        step[q] = (T)(-1) * (x - y);
    }
}

/* step = div( grad f / sqrt(|grad f|^2 + eps^2) )  tv_inpainter.py:24-33
 * gradient(): [horizontal fwd, vertical fwd]       differential.py:163-168
 * divergence(): bwd horizontal of [0] + bwd vertical of [1]   differential.py:187 */
static void FN(step_tv)(const T *f, T *step, FN(scratch) *s, const oracle_params *p,
                        const uint8_t *work)
{
    FN(conv3x3)(f, s->g0, s->rows, s->cols, K_FWD_HORZ, p->orientation, p->border, work);
    FN(conv3x3)(f, s->g1, s->rows, s->cols, K_FWD_VERT, p->orientation, p->border, work);
    const T eps2 = (T)(p->epsilon * p->epsilon); /* Python-float product, then narrowed */
#pragma omp parallel for schedule(static)
    for (size_t q = 0; q < s->n; ++q) {
        const T a0 = s->g0[q] * s->g0[q];
        const T a1 = s->g1[q] * s->g1[q];
        const T ampl = (T)sqrt((double)((a0 + a1) + eps2)); /* sqrt is correctly rounded in T */
        s->g0[q] = s->g0[q] / ampl;
        s->g1[q] = s->g1[q] / ampl;
    }
    FN(conv3x3)(s->g0, s->d0, s->rows, s->cols, K_BWD_HORZ, p->orientation, p->border, work);
    FN(conv3x3)(s->g1, s->d1, s->rows, s->cols, K_BWD_VERT, p->orientation, p->border, work);
#pragma omp parallel for schedule(static)
    for (size_t q = 0; q < s->n; ++q) step[q] = s->d0[q] + s->d1[q];
}

/* step = uxx*v0*v0 + uyy*v1*v1 + (uxy+uyx)*v0*v1   amle_inpainter.py:41-59 */
static void FN(step_amle)(const T *f, T *step, FN(scratch) *s, const oracle_params *p,
                          const uint8_t *work)
{
    const int R = s->rows, C = s->cols, o = p->orientation, bd = p->border;
    FN(conv3x3)(f, s->ux, R, C, K_FWD_VERT, o, bd, work);
    FN(conv3x3)(f, s->uy, R, C, K_FWD_HORZ, o, bd, work);
    FN(conv3x3)(s->ux, s->uxx, R, C, K_BWD_VERT, o, bd, work);
    FN(conv3x3)(s->ux, s->uxy, R, C, K_BWD_HORZ, o, bd, work);
    FN(conv3x3)(s->uy, s->uyx, R, C, K_BWD_VERT, o, bd, work);
    FN(conv3x3)(s->uy, s->uyy, R, C, K_BWD_HORZ, o, bd, work);
    FN(conv3x3)(f, s->v0, R, C, K_CEN_VERT, o, bd, work);
    FN(conv3x3)(f, s->v1, R, C, K_CEN_HORZ, o, bd, work);
    const T tiny = (T)1e-15;
#pragma omp parallel for schedule(static)
    for (size_t q = 0; q < s->n; ++q) {
        T v0 = s->v0[q], v1 = s->v1[q];
        const T den = (T)sqrt((double)((v0 * v0 + v1 * v1) + tiny));
        v0 = v0 / den;
        v1 = v1 / den;
        const T t0 = (s->uxx[q] * v0) * v0;
        const T t1 = (s->uyy[q] * v1) * v1;
        const T t2 = ((s->uxy[q] + s->uyx[q]) * v0) * v1;
        step[q] = (t0 + t1) + t2;
    }
}

static void FN(step_fun)(const T *f, T *step, FN(scratch) *s, const oracle_params *p,
                         const uint8_t *work)
{
    switch (p->method) {
    case ORACLE_HARMONIC: FN(step_harmonic)(f, step, s->rows, s->cols, p, work); break;
    case ORACLE_CCST:     FN(step_ccst)(f, step, s, p, work); break;
    case ORACLE_TV:       FN(step_tv)(f, step, s, p, work); break;
    default:              FN(step_amle)(f, step, s, p, work); break;
    }
}

static int FN(scratch_init)(FN(scratch) *s, int rows, int cols, int method)
{
    memset(s, 0, sizeof(*s));
    s->rows = rows; s->cols = cols; s->n = (size_t)rows * cols;
    int ok = 1;
#define GET(field) do { s->field = FN(alloc)(s->n); ok = ok && s->field; } while (0)
    if (method == ORACLE_CCST) { GET(h); GET(b); }
    if (method == ORACLE_TV) { GET(g0); GET(g1); GET(d0); GET(d1); }
    if (method == ORACLE_AMLE) { GET(ux); GET(uy); GET(uxx); GET(uxy); GET(uyx); GET(uyy); GET(v0); GET(v1); }
#undef GET
    return ok;
}

static void FN(scratch_free)(FN(scratch) *s)
{
    free(s->h); free(s->b); free(s->g0); free(s->g1); free(s->d0); free(s->d1);
    free(s->ux); free(s->uy); free(s->uxx); free(s->uxy); free(s->uyx); free(s->uyy);
    free(s->v0); free(s->v1);
}

/* One evaluation of the subclass step function on the whole grid (for the
 * single-step golden vectors).  work_mask may be NULL.                       */
int FN(oracle_step_fun)(const T *f, const uint8_t *work_mask, int rows, int cols,
                        const oracle_params *p, T *step_out)
{
    FN(scratch) s;
    if (!FN(scratch_init)(&s, rows, cols, p->method)) { FN(scratch_free)(&s); return -2; }
    FN(step_fun)(f, step_out, &s, p, work_mask);
    FN(scratch_free)(&s);
    return 0;
}

/* term_check                                    fd_pde_inpainter.py:368-379 */
static double FN(term_check)(const T *f, const T *fnew, size_t n, int criteria)
{
    if (criteria == ORACLE_RELATIVE) {
        double num = 0.0, den = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : num, den)
        for (size_t q = 0; q < n; ++q) {
            const T d = fnew[q] - f[q];
            num += (double)(T)(d * d);
            den += (double)(T)(fnew[q] * fnew[q]);
        }
        return sqrt(num) / sqrt(den); /* 0/0 -> NaN -> never terminates, like the reference */
    }
    double mx = 0.0;
    int has_nan = 0;
#pragma omp parallel for schedule(static) reduction(max : mx) reduction(| : has_nan)
    for (size_t q = 0; q < n; ++q) {
        const T d = fnew[q] - f[q];
        const double a = fabs((double)d);
        if (a != a) has_nan = 1;
        if (a > mx) mx = a;
    }
    return has_nan ? NAN : mx; /* np.max propagates NaN */
}

/* inpaint_grid                                  fd_pde_inpainter.py:281-366
 * image : initialised grid (unknown cells already filled by the initialiser)
 * mask  : 1 = known, 0 = unknown
 * out   : rows*cols result.  status: updates_done, last_diff, converged.
 * snapshot_every > 0 is not supported; callers wanting K-sweep states pass
 * max_iters = K with an unreachable threshold.                               */
int FN(oracle_inpaint_grid)(const T *image, const uint8_t *mask, int rows, int cols,
                            const oracle_params *p, T *out, long long *updates_done,
                            double *last_diff, int *converged)
{
    const size_t n = (size_t)rows * cols;
    if (rows < 2 || cols < 2) return -1;
    if (p->term_check_iters <= 0) return -1; /* reference: ZeroDivisionError */
    T *f = FN(alloc)(n), *fnew = FN(alloc)(n), *step = FN(alloc)(n);
    uint8_t *unknown = (uint8_t *)malloc(n), *work = NULL;
    FN(scratch) s;
    int ok = FN(scratch_init)(&s, rows, cols, p->method);
    if (!f || !fnew || !step || !unknown || !ok) {
        free(f); free(fnew); free(step); free(unknown); FN(scratch_free)(&s);
        return -2;
    }
    for (size_t q = 0; q < n; ++q) unknown[q] = mask[q] ? 0 : 1; /* mask_inv = 1 - mask (:291) */
    if (p->masked) {                                             /* :293-296 */
        work = (uint8_t *)malloc(n);
        dilate3x3(unknown, work, rows, cols);
    }
    memcpy(f, image, n * sizeof(T));                             /* f = image (:303) */

    const T dt = (T)p->dt;
    const T w = (T)p->relaxation, one_minus_w = (T)(1.0 - p->relaxation);
    double diff = 100000.0;                                      /* :309 */
    long long done = 0;
    int term = 0, stopped = 0;
    for (long long i = 0; i < p->max_iters; ++i) {               /* :311 */
        FN(step_fun)(f, step, &s, p, work);
        /* fnew = update_at_mask(image, f + dt*step, mask_inv)      :314 */
#pragma omp parallel for schedule(static)
        for (size_t q = 0; q < n; ++q) {
            const T cand = f[q] + dt * step[q];
            fnew[q] = unknown[q] ? cand : image[q];
        }
        if (p->relaxation > 1.0) {                               /* :317-318 */
#pragma omp parallel for schedule(static)
            for (size_t q = 0; q < n; ++q) {
                const T g = f[q] * one_minus_w + fnew[q] * w;
                fnew[q] = unknown[q] ? g : image[q];
            }
        }
        const int check = (i % p->term_check_iters) == 0;        /* :322 */
        if (check) {
            diff = FN(term_check)(f, fnew, n, p->criteria);
            term = diff < p->term_thres;
        }
        T *tmp = f; f = fnew; fnew = tmp;                        /* f = fnew (:326) */
        done = i + 1;
        if (check && term) { stopped = 1; break; }               /* :342-351 */
    }
    memcpy(out, f, n * sizeof(T));
    if (updates_done) *updates_done = done;
    if (last_diff) *last_diff = diff;
    if (converged) *converged = stopped;
    free(f); free(fnew); free(step); free(unknown); free(work);
    FN(scratch_free)(&s);
    return 0;
}

#undef FN
#undef CAT
#undef CAT_
