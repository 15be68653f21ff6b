/* Minimal cblas.h shim (test infrastructure, written for this repo).
 * The reference includes <cblas.h> (include/utils/math_functions.hh:9) but this image ships no
 * OpenBLAS headers; the OpenBLAS 0.3.15 shared object bundled with opencv_python_headless exports the
 * standard LP64 CBLAS symbols, which is all the GNN path calls (math_functions.cpp:148,226,313,359,365). */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
void cblas_sgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int M, int N, int K, float alpha,
                 const float* A, int lda, const float* B, int ldb, float beta, float* C, int ldc);
void cblas_sgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, int M, int N, float alpha, const float* A, int lda,
                 const float* x, int incx, float beta, float* y, int incy);
void cblas_sscal(int n, float alpha, float* x, int incx);
void cblas_saxpy(int n, float alpha, const float* x, int incx, float* y, int incy);
void cblas_scopy(int n, const float* x, int incx, float* y, int incy);
float cblas_sdot(int n, const float* x, int incx, const float* y, int incy);
void openblas_set_num_threads(int n);
#ifdef __cplusplus
}
#endif
