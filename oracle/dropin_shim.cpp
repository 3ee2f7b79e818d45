// dropin_shim.cpp -- TEST INFRASTRUCTURE.  C entry points that construct the product's
// reference-side binding (gaselect_b200/host/GpuEvaluator) as a plain `Evaluator*`, so that the
// reference's OWN population drivers (compiled from /root/reference/src into libgaselect_ref.so,
// entry ref_ga_run of ref_shim.cpp) can be run on top of the CUDA engine: the drop-in test.
#include "BatchedPopulation.h"
#include "GpuEvaluator.h"

#include <chrono>

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

// genAlg() on the product's BatchedPopulation (same Control as ref_ga_run of ref_shim.cpp builds for the
// reference's drivers).  stats: engine evaluations, engine calls, passes, cache hits.
int dropin_ga_run_batched(void* evaluator, int chromosomeSize, int popSize, int numGenerations, int elitism, int minVariables,
                          int maxVariables, double mutationProb, int numThreads, int maxDuplicateEliminationTries, double badSolutionThreshold,
                          double convergenceThreshold, int convergenceGenerations, int crossover, int fitnessScaling,
                          const uint32_t* seedvec, int resultCapacity, long long idxCapacity, int* nResult, uint32_t* resIdx,
                          uint32_t* resOffsets, double* resFitness, int evolutionCapacity, int* nEvolution,
                          double* fitnessEvolution, unsigned long long* stats, long long* elapsed_ns) {
	try {
		GpuEvaluator* gpu = dynamic_cast<GpuEvaluator*>(static_cast<Evaluator*>(evaluator));
		if (!gpu) return 3;
		Control ctrl(static_cast<uint16_t>(chromosomeSize), static_cast<uint16_t>(popSize), static_cast<uint16_t>(numGenerations),
		             static_cast<uint16_t>(elitism), static_cast<uint16_t>(minVariables), static_cast<uint16_t>(maxVariables),
		             mutationProb, static_cast<uint16_t>(numThreads), static_cast<uint16_t>(maxDuplicateEliminationTries), badSolutionThreshold,
		             convergenceThreshold, static_cast<uint16_t>(convergenceGenerations), static_cast<CrossoverType>(crossover),
		             static_cast<FitnessScaling>(fitnessScaling), OFF);
		const std::vector<uint32_t> seed(seedvec, seedvec + 624);
		BatchedPopulation pop(ctrl, *gpu, seed);
		const auto t0 = std::chrono::steady_clock::now();
		pop.run();
		const auto t1 = std::chrono::steady_clock::now();
		if (elapsed_ns) *elapsed_ns = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
		if (stats) {
			stats[0] = pop.stats().engineEvaluations;
			stats[1] = pop.stats().engineCalls;
			stats[2] = pop.stats().passes;
			stats[3] = pop.stats().cacheHits;
		}
		int rc = 0, count = 0;
		long long pos = 0;
		const Population::SortedChromosomes result = pop.getResult();
		for (Population::SortedChromosomes::const_iterator it = result.begin(); it != result.end(); ++it) {
			const arma::uvec cols = it->toColumnSubset();
			if (count >= resultCapacity || pos + static_cast<long long>(cols.n_elem) > idxCapacity) { rc = 1; break; }
			resOffsets[count] = static_cast<uint32_t>(pos);
			for (arma::uword j = 0; j < cols.n_elem; ++j) resIdx[pos++] = static_cast<uint32_t>(cols[j]);
			resFitness[count++] = it->getFitness();
		}
		resOffsets[count] = static_cast<uint32_t>(pos);
		*nResult = count;
		const std::vector<double>& evo = pop.getFitnessEvolution();
		*nEvolution = static_cast<int>(std::min<size_t>(evo.size(), static_cast<size_t>(evolutionCapacity)));
		for (int i = 0; i < *nEvolution; ++i) fitnessEvolution[i] = evo[static_cast<size_t>(i)];
		return rc;
	} catch (...) {
		return 2;
	}
}

}
