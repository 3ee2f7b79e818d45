// dropin_shim.cpp -- TEST INFRASTRUCTURE.  C entry points that construct the product's
// reference-side binding (gaselect_b200/host/GpuEvaluator) as a plain `Evaluator*`, so that the
// reference's OWN population drivers (compiled from /root/reference/src into libgaselect_ref.so,
// entry ref_ga_run of ref_shim.cpp) can be run on top of the CUDA engine: the drop-in test.
#include "GpuEvaluator.h"

namespace {
arma::mat wrapMatrix(const double* colmajor, int n, int p) {
	arma::mat X(static_cast<arma::uword>(n), static_cast<arma::uword>(p));
	for (size_t i = 0; i < static_cast<size_t>(n) * p; ++i) X.memptr()[i] = colmajor[i];
	return X;
}
arma::vec wrapVector(const double* v, int n) {
	arma::vec y(static_cast<arma::uword>(n));
	for (int i = 0; i < n; ++i) y[i] = v[i];
	return y;
}
}

extern "C" {

void* dropin_pls_create(const double* X, const double* y, int n, int p, int numReplications, int maxNComp,
                        const uint32_t* seedvec, int innerSegments, int outerSegments, double testSetSize, double sdfact,
                        double complexityPenalty) {
	try {
		const std::vector<uint32_t> seed(seedvec, seedvec + 624);
		Evaluator* ev = GpuEvaluator::makePLS(wrapMatrix(X, n, p), wrapVector(y, n), static_cast<uint16_t>(numReplications),
		                                      static_cast<uint16_t>(maxNComp), seed, OFF, static_cast<uint16_t>(innerSegments),
		                                      static_cast<uint16_t>(outerSegments), testSetSize, sdfact, complexityPenalty);
		return ev;
	} catch (...) {
		return nullptr;
	}
}

void* dropin_lm_create(const double* X, const double* y, int n, int p, int statistic) {
	try {
		Evaluator* ev = GpuEvaluator::makeLM(wrapMatrix(X, n, p), wrapVector(y, n), statistic, OFF);
		return ev;
	} catch (...) {
		return nullptr;
	}
}

// evaluate one subset through the virtual interface; returns 0 ok, 1 EvaluatorException, 2 other
int dropin_evaluate_one(void* evaluator, const uint32_t* cols, int k, double* fitness) {
	try {
		arma::uvec subset(static_cast<arma::uword>(k));
		for (int i = 0; i < k; ++i) subset[i] = cols[i];
		*fitness = static_cast<Evaluator*>(evaluator)->evaluate(subset);
		return 0;
	} catch (const Evaluator::EvaluatorException&) {
		return 1;
	} catch (...) {
		return 2;
	}
}

void dropin_destroy(void* evaluator) { delete static_cast<Evaluator*>(evaluator); }

}
