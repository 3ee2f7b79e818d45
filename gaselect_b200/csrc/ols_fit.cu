// ols_fit.cu -- K3: ordinary least squares with intercept + BIC/AIC/R2 epilogue, one CTA per
// chromosome.  Replaces LMEvaluator::evaluate (src/LMEvaluator.cpp:27-67):
//   Xsub = [1 | X[:, sel]]; beta = arma::solve(Xsub, y); RSS = sum (y - Xsub beta)^2; statistic.
// The reference solves by LAPACK least squares (QR); here the slope coefficients come from an
// LDL' factorisation (Cholesky's pivots, no square roots) of the centred Gram sub-matrix (gathered from
// the all-rows Gram table into shared memory, packed lower triangle), and -- as SURVEY.md Appendix B.1 shows is
// necessary for 1e-8 parity when R^2 -> 1 -- the RSS is accumulated from EXPLICIT residuals
// over coalesced column segments of X[:, sel], not from the quadratic form.
// A pivot that is not safely positive marks the subset GSB_STATUS_SINGULAR (the reference
// reports "system could not be solved", src/LMEvaluator.cpp:60-61).
#include "kernels.cuh"

#include <algorithm>
#include <cmath>

namespace gsb {

namespace {

constexpr int kOlsThreads = 128;
constexpr double kPivotTol = 1e-12; // relative to the original diagonal entry

__device__ __forceinline__ double blockSum1(double v, double* red) {
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
	const int warp = threadIdx.x >> 5;
	__syncthreads();
	if ((threadIdx.x & 31) == 0) red[warp] = v;
	__syncthreads();
	double s = red[0];
#pragma unroll
	for (int w = 1; w < kOlsThreads / 32; ++w) s += red[w];
	return s;
}

// GLOB: the packed triangle of a subset too wide for shared memory (k > ~230) is factorised in a global scratch, one region
// per CTA of the launch (L2-resident while it runs); everything else stays in shared memory
template <bool GLOB>
__global__ void __launch_bounds__(kOlsThreads) olsFitKernel(const OlsParams prm) {
	extern __shared__ __align__(16) double smem[];
	const int chrom = prm.bucket[blockIdx.x];
	const uint32_t selBegin = prm.offsets[chrom];
	const int k = static_cast<int>(prm.offsets[chrom + 1] - selBegin);
	const int tid = threadIdx.x;
	const int ntri = k * (k + 1) / 2;
	const int ntriS = GLOB ? 0 : ntri;
	double* __restrict__ L = GLOB ? prm.scratch + static_cast<size_t>(blockIdx.x) * prm.scratchStride : smem; // packed lower triangle, row-major
	double* __restrict__ rhs = smem + ntriS + (ntriS & 1); // s0 -> z -> beta
	double* __restrict__ diag0 = rhs + k + (k & 1);
	double* __restrict__ xms = diag0 + k + (k & 1);
	double* red = xms + k + (k & 1);
	uint32_t* __restrict__ sel = reinterpret_cast<uint32_t*>(red + 8);
	__shared__ int singular;

	if (tid == 0) singular = 0;
	for (int i = tid; i < k; i += kOlsThreads) sel[i] = prm.idx[selBegin + i];
	__syncthreads();
	for (int t = tid; t < ntri; t += kOlsThreads) {
		int a = static_cast<int>((sqrtf(8.0f * static_cast<float>(t) + 1.0f) - 1.0f) * 0.5f);
		while ((a + 1) * (a + 2) / 2 <= t) ++a;
		while (a * (a + 1) / 2 > t) --a;
		const int b = t - a * (a + 1) / 2;
		const double g = __ldg(prm.gram + tri(sel[a]) + sel[b]);
		L[t] = g;
		if (a == b) diag0[a] = g;
	}
	for (int i = tid; i < k; i += kOlsThreads) {
		rhs[i] = __ldg(prm.s0 + sel[i]);
		xms[i] = __ldg(prm.xm + sel[i]);
	}
	__syncthreads();

	// Right-looking LDL' (the Cholesky pivots without the square roots) of the packed triangle, the right-hand side
	// carried as one more row (row k): after column j every entry (r, c > j) has lost A[r][j] A[c][j] / d_j, so the
	// last row ends as M^-1 s0 scaled by D -- the forward substitution is done -- and column j is never written again
	// after step j - 1: ONE barrier per column (the scaled-column form needed three, and two more per column for the
	// substitutions: a 50-column subset paid 350 barriers).
	for (int j = 0; j < k; ++j) {
		const int tj = j * (j + 1) / 2;
		const double piv = L[tj + j];
		if (!(piv > kPivotTol * diag0[j])) { // uniform: every thread reads the same value
			if (tid == 0) singular = 1;
			break;
		}
		const double rp = 1.0 / piv;
		for (int r = j + 1 + tid; r <= k; r += kOlsThreads) {
			double* __restrict__ row = (r < k) ? L + r * (r + 1) / 2 : rhs;
			const double f = row[j] * rp;
			const int cend = (r < k) ? r : k - 1;
			for (int c = j + 1; c <= cend; ++c) row[c] = fma(-f, L[c * (c + 1) / 2 + j], row[c]);
		}
		__syncthreads();
	}
	__syncthreads();
	if (singular) {
		if (tid == 0) { prm.fitness[chrom] = 0.0; prm.status[chrom] = 3; }
		return;
	}
	// backward substitution M' beta = D^-1 (last row), column oriented, by ONE warp: lane c mod 32 owns entry c, beta_j
	// travels by shuffle, no CTA barrier: beta_j = acc[j] / d_j; acc[c] -= A[j][c] beta_j for c < j (row j is contiguous)
	if (tid < 32) {
		for (int j = k - 1; j >= 0; --j) {
			const int tj = j * (j + 1) / 2;
			double bj = 0.0;
			if ((j & 31) == tid) { bj = rhs[j] / L[tj + j]; rhs[j] = bj; }
			bj = __shfl_sync(0xffffffffu, bj, j & 31);
			for (int c = tid; c < j; c += 32) rhs[c] = fma(-L[tj + c], bj, rhs[c]);
		}
	}
	__syncthreads();
	// intercept on the (globally centred) data: b0 = ym - xm'beta
	double part = 0.0;
	for (int i = tid; i < k; i += kOlsThreads) part = fma(xms[i], rhs[i], part);
	const double b0 = prm.ym - blockSum1(part, red);
	// explicit residuals over all rows (src/LMEvaluator.cpp:35-37)
	const double* __restrict__ ycol = prm.z + static_cast<size_t>(prm.p) * prm.ldz;
	double rss = 0.0;
	for (int r = tid; r < prm.n; r += kOlsThreads) {
		double pred = b0;
#pragma unroll 4
		for (int c = 0; c < k; ++c) pred = fma(__ldg(prm.z + static_cast<size_t>(sel[c]) * prm.ldz + r), rhs[c], pred);
		const double e = __ldg(ycol + r) - pred;
		rss = fma(e, e, rss);
	}
	const double RSS = blockSum1(rss, red);
	if (tid == 0) {
		const double dn = static_cast<double>(prm.n), dm = static_cast<double>(k + 1);
		double fit;
		if (prm.statistic == 0) fit = -(dn * log(RSS / dn) + dm * log(dn));          // :41
		else if (prm.statistic == 1) fit = -(dn * log(RSS / dn) + dm * 2.0);          // :45
		else if (prm.statistic == 2) {                                                // :49-50
			const double r2 = 1.0 - (RSS / prm.r2denom);
			fit = 1.0 - (((dn - 1.0) * (1.0 - r2)) / (dn - dm - 1.0));
		} else fit = 1.0 - (RSS / prm.r2denom);                                       // :54
		prm.fitness[chrom] = fit;
		prm.status[chrom] = 0;
	}
}

int olsSharedDoubles(int k, bool glob = false) {
	const int ntri = glob ? 0 : k * (k + 1) / 2;
	return ntri + (ntri & 1) + 3 * (k + (k & 1)) + 8 + (k + 1) / 2 + 1;
}

} // namespace

int olsKernelMaxK(int smemLimit) {
	int k = 1;
	while (k < 65535 && olsSharedDoubles(k + 1, true) * 8 <= smemLimit) ++k; // with the triangle in the global scratch
	return k;
}

// doubles of global scratch per chromosome for subsets up to kMax columns (0: the triangle fits in shared memory)
unsigned long long olsKernelScratchPerCta(int kMax, int smemLimit) {
	if (olsSharedDoubles(kMax) * 8 <= smemLimit) return 0ull;
	return (static_cast<unsigned long long>(kMax) * (kMax + 1) / 2 + 1ull) & ~1ull;
}

cudaError_t launchOlsKernel(const OlsParams& prm, int kMax, int smemLimit, cudaStream_t stream) {
	if (prm.bucketCount <= 0) return cudaSuccess;
	const unsigned long long per = olsKernelScratchPerCta(kMax, smemLimit);
	if (per == 0) {
		const size_t bytes = static_cast<size_t>(olsSharedDoubles(kMax)) * sizeof(double);
		cudaError_t err = cudaFuncSetAttribute(olsFitKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
		if (err != cudaSuccess) return err;
		olsFitKernel<false><<<static_cast<unsigned>(prm.bucketCount), kOlsThreads, bytes, stream>>>(prm);
		return cudaGetLastError();
	}
	// the chromosomes of the class in chunks that fit the scratch the engine provides (same stream: reused in order)
	const size_t bytes = static_cast<size_t>(olsSharedDoubles(kMax, true)) * sizeof(double);
	if (static_cast<int>(bytes) > smemLimit || prm.scratch == nullptr || prm.scratchCapacity < per) return cudaErrorInvalidValue;
	cudaError_t err = cudaFuncSetAttribute(olsFitKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
	if (err != cudaSuccess) return err;
	const int chunk = static_cast<int>(std::min<unsigned long long>(static_cast<unsigned long long>(prm.bucketCount), prm.scratchCapacity / per));
	for (int c0 = 0; c0 < prm.bucketCount; c0 += chunk) {
		OlsParams sub = prm;
		sub.bucket = prm.bucket + c0;
		sub.bucketCount = std::min(chunk, prm.bucketCount - c0);
		sub.scratchStride = per;
		olsFitKernel<true><<<static_cast<unsigned>(sub.bucketCount), kOlsThreads, bytes, stream>>>(sub);
		err = cudaGetLastError();
		if (err != cudaSuccess) return err;
	}
	return cudaSuccess;
}

} // namespace gsb
