// cv_reduce.cu -- K2, the per-chromosome epilogue.
//
// Replaces the data-dependent tail of PLSEvaluator::estSEP (src/PLSEvaluator.cpp:284-335):
//   * per (replication, outer fold): mean and sample SD over the I inner folds of the MSEP of
//     every component count (OnlineStddev, src/OnlineStddev.h:36-44), the "one-SE rule"
//     (src/PLSEvaluator.cpp:284-308) -> optNComp;
//   * residual statistics of the outer fit at optNComp, pooled over the outer folds of a
//     replication; SEP = sample SD of the pooled residuals (:323-335);
//   * fitness = -(1 + gamma k) * mean over replications (src/PLSEvaluator.cpp:87);
//   * status 4 (GSB_STATUS_FLAGGED) when a comparison of the rule was a near-tie (relative band gsb_config.tie_band):
//     the fitness is valid for this engine, but a flipped optNComp moves it by O(1e-2), so the host may re-score.
// and of BICEvaluator::getRSS / evaluate (src/BICEvaluator.cpp:191-226, :64-88): the same rule
// over the K folds, RSS of the all-rows fit at optNComp, then BIC / AIC / adjusted R2 / R2.
#include "kernels.cuh"

#include <cmath>

namespace gsb {

namespace {

constexpr int kReduceThreads = 64;

__global__ void __launch_bounds__(kReduceThreads) cvReduceKernel(const ReduceParams prm) {
	const int chrom = blockIdx.x;
	const int k = static_cast<int>(prm.offsets[chrom + 1] - prm.offsets[chrom]);
	if (k == 0) { // empty subset: src/PLSEvaluator.cpp:72-75
		if (threadIdx.x == 0) { prm.fitness[chrom] = 0.0; prm.status[chrom] = 1; }
		return;
	}
	if (prm.refit != nullptr && prm.refit[chrom] != 0) { // the kernel-space kernel declined it: to be scored again (engine.cpp, resolveRetries)
		if (threadIdx.x == 0) { prm.fitness[chrom] = 0.0; prm.status[chrom] = kStatusRetry; }
		return;
	}
	extern __shared__ double sh[]; // per group: sum e, sum e^2, flag
	double* gE = sh;
	double* gEE = sh + prm.groups;
	int* gBad = reinterpret_cast<int*>(sh + 2 * prm.groups);

	const int A = min(prm.maxNComp, k);
	const int I = prm.inner;
	const double* mseC = prm.mse + static_cast<size_t>(chrom) * prm.innerDests * prm.maxNComp;
	const double* ostC = prm.ostat + static_cast<size_t>(chrom) * prm.outerDests * prm.maxNComp * 2;

	for (int g = threadIdx.x; g < prm.groups; g += kReduceThreads) {
		const double* m = mseC + static_cast<size_t>(g) * I * prm.maxNComp;
		const OneSe rule = oneSeRule(m, I, A, prm.maxNComp, prm.sdfact, prm.tieBand);
		const bool bad = rule.bad;
		const int opt = rule.opt;
		const double* o = ostC + (static_cast<size_t>(g) * prm.maxNComp + (opt - 1)) * 2;
		gE[g] = o[0];
		gEE[g] = o[1];
		gBad[g] = ((bad || isnan(o[1])) ? 1 : 0) | (rule.tie ? 2 : 0);
	}
	__syncthreads();
	if (threadIdx.x != 0) return;

	bool bad = false, tie = false;
	for (int g = 0; g < prm.groups; ++g) {
		bad = bad || (gBad[g] & 1) != 0;
		tie = tie || (gBad[g] & 2) != 0;
	}
	if (bad) { // the reference throws: src/PLSEvaluator.cpp:337-340, src/BICEvaluator.cpp:227-230
		prm.fitness[chrom] = 0.0;
		prm.status[chrom] = 2;
		return;
	}
	double fit;
	if (prm.kind == 1) { // PLS_CV
		double sumSEP = 0.0;
		for (int rep = 0; rep < prm.replications; ++rep) {
			double se = 0.0, see = 0.0;
			int cnt = 0;
			for (int o = 0; o < prm.outerPerRep; ++o) {
				const int g = rep * prm.outerPerRep + o;
				se += gE[g];
				see += gEE[g];
				cnt += prm.outerRows[g];
			}
			const double m2 = see - se * se / static_cast<double>(cnt);
			sumSEP += sqrt(m2 / static_cast<double>(cnt - 1)); // :335
		}
		fit = sumSEP * (-(1.0 + prm.gamma * static_cast<double>(k)) / static_cast<double>(prm.replications)); // :87
	} else { // PLS_FIT
		const double RSS = gEE[0];
		const double dn = static_cast<double>(prm.n);
		const double df = static_cast<double>(k) + 2.0; // src/BICEvaluator.cpp:64
		if (prm.statistic == 0) fit = -(dn * log(RSS / dn) + df * log(dn));           // :70
		else if (prm.statistic == 1) fit = -(dn * log(RSS / dn) + df * 2.0);           // :74
		else if (prm.statistic == 2) {                                                 // :78-79
			const double r2 = 1.0 - (RSS / prm.r2denom);
			fit = 1.0 - (((dn - 1.0) * (1.0 - r2)) / (dn - df - 1.0));
		} else fit = 1.0 - (RSS / prm.r2denom);                                        // :83
	}
	prm.fitness[chrom] = fit;
	prm.status[chrom] = tie ? 4 : 0; // GSB_STATUS_FLAGGED: the fitness is this engine's, a near-tie decided optNComp
}

__global__ void markEmptyKernel(const uint32_t* offsets, int npop, double* fitness, int* status) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c < npop && offsets[c + 1] == offsets[c]) { fitness[c] = 0.0; status[c] = 1; }
}

// LMEvaluator does not refuse an empty subset (src/LMEvaluator.cpp:27-67): Xsub is the intercept column alone, the
// solve returns the mean of y, RSS = sum (y - mean)^2 = r2denom, Xsub.n_cols = 1.
__global__ void lmEmptyKernel(const uint32_t* offsets, int npop, int statistic, int n, double r2denom, double* fitness, int* status) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= npop || offsets[c + 1] != offsets[c]) return;
	const double dn = static_cast<double>(n), dm = 1.0, RSS = r2denom;
	double fit;
	if (statistic == 0) fit = -(dn * log(RSS / dn) + dm * log(dn));   // :41
	else if (statistic == 1) fit = -(dn * log(RSS / dn) + dm * 2.0);   // :45
	else if (statistic == 2) {                                         // :49-50
		const double r2 = 1.0 - (RSS / r2denom);
		fit = 1.0 - (((dn - 1.0) * (1.0 - r2)) / (dn - dm - 1.0));
	} else fit = 1.0 - (RSS / r2denom);                                // :54
	fitness[c] = fit;
	status[c] = 0;
}

} // namespace

void launchReduceKernel(const ReduceParams& prm, cudaStream_t stream) {
	if (prm.npop <= 0) return;
	const size_t bytes = static_cast<size_t>(prm.groups) * (2 * sizeof(double) + sizeof(int)) + 16;
	cvReduceKernel<<<static_cast<unsigned>(prm.npop), kReduceThreads, bytes, stream>>>(prm);
}

void launchMarkEmpty(const uint32_t* offsets, int npop, double* fitness, int* status, cudaStream_t stream) {
	if (npop <= 0) return;
	markEmptyKernel<<<(npop + 255) / 256, 256, 0, stream>>>(offsets, npop, fitness, status);
}

void launchLmEmpty(const uint32_t* offsets, int npop, int statistic, int n, double r2denom, double* fitness, int* status, cudaStream_t stream) {
	if (npop <= 0) return;
	lmEmptyKernel<<<(npop + 255) / 256, 256, 0, stream>>>(offsets, npop, statistic, n, r2denom, fitness, status);
}

} // namespace gsb
