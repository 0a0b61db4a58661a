/* Minimal prototypes for the four OpenBLAS symbols the reference's BLAS path uses
 * (Network.cpp:239-240, 375-380, 404-409, 1543). The image has an OpenBLAS shared
 * object (inside the opencv / scipy wheels) but no cblas.h. Test infrastructure only. */
#ifndef LB2_SHIM_CBLAS_H
#define LB2_SHIM_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
#ifdef LB2_SCIPY_OPENBLAS
#define cblas_sgemm scipy_cblas_sgemm
#define cblas_sgemv scipy_cblas_sgemv
#define openblas_set_num_threads scipy_openblas_set_num_threads
#define openblas_get_corename scipy_openblas_get_corename
#endif
void cblas_sgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb,
                 int m, int n, int k, float alpha, const float* a, int lda,
                 const float* b, int ldb, float beta, float* c, int ldc);
void cblas_sgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, int m, int n, float alpha,
                 const float* a, int lda, const float* x, int incx, float beta,
                 float* y, int incy);
void openblas_set_num_threads(int n);
char* openblas_get_corename(void);
#ifdef __cplusplus
}
#endif
#endif
