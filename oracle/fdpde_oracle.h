/* TEST INFRASTRUCTURE ONLY — C interface of the CPU oracle (see fdpde_oracle.c). */
#ifndef FDPDE_ORACLE_H
#define FDPDE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_HARMONIC = 0, ORACLE_CCST = 1, ORACLE_TV = 2, ORACLE_AMLE = 3 };
enum { ORACLE_BORDER_REFLECT101 = 0, ORACLE_BORDER_SYMMETRIC = 1, ORACLE_BORDER_ZERO = 2 };
enum { ORACLE_RELATIVE = 0, ORACLE_ABSOLUTE = 1 };

typedef struct {
    int method;      /* ORACLE_HARMONIC ...                                                  */
    int orientation; /* +1 correlation (convolver "opencv"), -1 convolution (masked*, scipy-*) */
    int border;      /* ORACLE_BORDER_*                                                      */
    int masked;      /* 1: convolver evaluates only on the dilated unknown set (masked*)      */
    int criteria;    /* ORACLE_RELATIVE / ORACLE_ABSOLUTE (absolute_percent pre-scaled)       */
    int reserved;
    double dt, tension, epsilon, relaxation, term_thres;
    long long term_check_iters, max_iters;
} oracle_params;

int oracle_inpaint_grid_f64(const double *image, const uint8_t *mask, int rows, int cols,
                            const oracle_params *p, double *out, long long *updates_done,
                            double *last_diff, int *converged);
int oracle_inpaint_grid_f32(const float *image, const uint8_t *mask, int rows, int cols,
                            const oracle_params *p, float *out, long long *updates_done,
                            double *last_diff, int *converged);
int oracle_step_fun_f64(const double *f, const uint8_t *work_mask, int rows, int cols,
                        const oracle_params *p, double *step_out);
int oracle_step_fun_f32(const float *f, const uint8_t *work_mask, int rows, int cols,
                        const oracle_params *p, float *step_out);
void oracle_dilate3x3(const uint8_t *src, uint8_t *dst, int rows, int cols);

#ifdef __cplusplus
}
#endif
#endif
