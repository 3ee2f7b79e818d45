// Microbenchmark (host CPU): the stand-in loops under the reference build (oracle/standin/RcppArmadillo.h: gemv_n, gemv_t with
// `omp simd` reductions, -O3 -march=x86-64-v3) against OpenBLAS dgemv (the library bundled in the scipy wheel, one thread) on the
// shapes of a fit: how much a tuned BLAS under the reference would change the CPU baseline.
// g++ -std=c++17 -O3 -march=x86-64-v3 -fopenmp-simd -fno-math-errno gemv_probe.cpp -o gemv_probe -ldl; ./gemv_probe <libscipy_openblas*.so>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <vector>
typedef void (*dgemv_t)(int order, int trans, int m, int n, double alpha, const double* a, int lda, const double* x, int incx, double beta, double* y, int incy);
typedef void (*setthr_t)(int);
static void gemv_n(const double* __restrict A, const double* __restrict x, double* __restrict y, int r, int c) {
	for (int i = 0; i < r; ++i) y[i] = 0.0;
	for (int j = 0; j < c; ++j) { const double* __restrict a = A + (size_t) j * r; const double xj = x[j]; for (int i = 0; i < r; ++i) y[i] += a[i] * xj; }
}
static void gemv_t(const double* __restrict A, const double* __restrict x, double* __restrict y, int r, int c) {
	for (int j = 0; j < c; ++j) { const double* __restrict a = A + (size_t) j * r; double s = 0.0;
#pragma omp simd reduction(+ : s)
 for (int i = 0; i < r; ++i) s += a[i] * x[i]; y[j] = s; }
}
int main(int argc, char** argv) {
	void* h = dlopen(argv[1], RTLD_NOW);
	if (!h) { printf("dlopen failed: %s\n", dlerror()); return 1; }
	dgemv_t dgemv = (dgemv_t) dlsym(h, "scipy_cblas_dgemv");
	setthr_t setthr = (setthr_t) dlsym(h, "scipy_openblas_set_num_threads");
	if (setthr) setthr(1);
	for (int cfg = 0; cfg < 3; ++cfg) {
		const int r = cfg == 0 ? 120 : (cfg == 1 ? 90 : 400), c = cfg == 0 ? 100 : (cfg == 1 ? 50 : 200);
		std::vector<double> A((size_t) r * c), x(c > r ? c : r, 0.5), y(c > r ? c : r);
		for (size_t i = 0; i < A.size(); ++i) A[i] = 1e-3 * (i % 977);
		const int it = 200000 / (cfg == 2 ? 8 : 1);
		double sink = 0;
		auto t0 = std::chrono::steady_clock::now();
		for (int k = 0; k < it; ++k) { gemv_n(A.data(), x.data(), y.data(), r, c); sink += y[k % r]; gemv_t(A.data(), y.data(), x.data(), r, c); sink += x[k % c]; for (int q = 0; q < c; ++q) x[q] = 0.5; }
		double a = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		t0 = std::chrono::steady_clock::now();
		for (int k = 0; k < it; ++k) { dgemv(102, 111, r, c, 1.0, A.data(), r, x.data(), 1, 0.0, y.data(), 1); sink += y[k % r]; dgemv(102, 112, r, c, 1.0, A.data(), r, y.data(), 1, 0.0, x.data(), 1); sink += x[k % c]; for (int q = 0; q < c; ++q) x[q] = 0.5; }
		double b = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		printf("%3d x %3d: stand-in loops %.2f us per (A x, A'y) pair, OpenBLAS dgemv %.2f us: ratio %.2f  (%g)\n", r, c, a / it * 1e6, b / it * 1e6, a / b, sink);
	}
	return 0;
}
