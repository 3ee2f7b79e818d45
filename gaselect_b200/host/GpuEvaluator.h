// GpuEvaluator.h -- the reference-side binding: gaselect's own Evaluator plugin interface
// (src/Evaluator.h:19-40 of dakep/gaselect) implemented on top of the C ABI of the B200 engine
// (include/gaselect_b200.h).  It is compiled INSIDE the gaselect package, next to the reference's
// sources (it includes the reference's own "Evaluator.h"; nothing of the reference is copied here),
// and drops into SingleThreadPopulation / MultiThreadedPopulation / the C_evaluate batch loop
// unchanged: construction sites src/GenAlg.cpp:119-127 (PLS), :131-141 (LM), :142-156 (FIT),
// :262-303 (evaluate).  See INTEGRATION.md.
//
// Error convention: per-chromosome status 1..3 -> Evaluator::EvaluatorException; status 4
// (GSB_STATUS_FLAGGED: valid fitness, near-tie in the one-SE rule) is counted (flaggedCount()), not thrown
// (callers re-draw the slot, src/MultiThreadedPopulation.cpp:202-206); invalid constructor arguments
// -> std::invalid_argument (src/PLSEvaluator.cpp:43,51); any other engine failure -> std::runtime_error.
#ifndef GASELECT_B200_GPU_EVALUATOR_H
#define GASELECT_B200_GPU_EVALUATOR_H

#include <memory>
#include <mutex>
#include <stdexcept>
#include <vector>

#include "Evaluator.h" // the reference's interface, found on the package's include path

#include "gaselect_b200.h"

class GpuEvaluator : public Evaluator {
public:
	// PLSEvaluator(pls, numReplications, maxNComp, seed, verbosity, innerSegments, outerSegments,
	//              testSetSize, sdfact, complexityPenalty)                 src/PLSEvaluator.cpp:27-57
	static GpuEvaluator* makePLS(const arma::mat& X, const arma::vec& y, uint16_t numReplications, uint16_t maxNComp,
	                             const std::vector<uint32_t>& seed, VerbosityLevel verbosity, uint16_t innerSegments,
	                             uint16_t outerSegments = 1, double testSetSize = 0.0, double sdfact = 1.0,
	                             double complexityPenalty = 0.0, int device = 0);
	// BICEvaluator(pls, maxNComp, seed, verbosity, numSegments, stat, sdfact)   src/BICEvaluator.cpp:22-44
	static GpuEvaluator* makeFit(const arma::mat& X, const arma::vec& y, uint16_t maxNComp, const std::vector<uint32_t>& seed,
	                             VerbosityLevel verbosity, uint16_t numSegments, int statistic, double sdfact = 1.0,
	                             int device = 0);
	// LMEvaluator(X, y, statistic, verbosity)                                   src/LMEvaluator.cpp:17-25
	static GpuEvaluator* makeLM(const arma::mat& X, const arma::vec& y, int statistic, VerbosityLevel verbosity, int device = 0);

	// Devices of every evaluator made afterwards: more than one ordinal makes the engine deal each batch to all of them
	// (one process, one CUDA context per GPU; include/gaselect_b200.h, gsb_config.devices) -- the role numThreads plays in
	// the reference (src/GenAlg.cpp:173-181).  Empty list: the `device` argument of make*.  Without a call, the
	// environment variable GASELECT_B200_DEVICES ("all" or a comma-separated list of ordinals) is consulted.
	static void setDefaultDevices(const std::vector<int>& devices);

	~GpuEvaluator() {}

	double evaluate(arma::uvec& columnSubset);   // src/Evaluator.h:30
	double evaluate(Chromosome& ch) {            // src/Evaluator.h:31, as src/PLSEvaluator.h:33-38
		arma::uvec columnSubset = ch.toColumnSubset();
		const double fitness = this->evaluate(columnSubset);
		ch.setFitness(fitness);
		return fitness;
	}
	Evaluator* clone() const;                    // src/Evaluator.h:33
	std::vector<arma::uvec> getSegmentation() const; // src/Evaluator.h:35-37

	// The whole point of the engine: score many subsets in ONE call (what a batching population
	// driver and the loop of src/GenAlg.cpp:312-325 use).  status[i] != 0 marks the subsets for
	// which the reference would have thrown EvaluatorException.
	void evaluateBatch(const std::vector<arma::uvec>& subsets, std::vector<double>& fitness, std::vector<int>& status);
	// The same for chromosomes as they are stored: `count` chromosomes of `words` 64-bit parts each, in the layout of
	// src/Chromosome.cpp:55-77,419-438 (gsb_evaluate_bitsets decodes it on the engine's side).
	void evaluateBitsets(const uint64_t* parts, int words, size_t count, std::vector<double>& fitness, std::vector<int>& status);

	// Evaluations (over all clones) whose one-SE rule met a near-tie (GSB_STATUS_FLAGGED, relative band 1e-9): their
	// fitness was returned as usual; the reference, summing in another order, may have fitted another optNComp there.
	size_t flaggedCount() const;

private:
	struct Engine { // shared by clones: an engine is not re-entrant, calls are serialised
		gsb_engine* handle = nullptr;
		mutable std::mutex lock;
		size_t flagged = 0;
		~Engine() { gsb_destroy(handle); }
	};
	GpuEvaluator(std::shared_ptr<Engine> engine, VerbosityLevel verbosity) : Evaluator(verbosity), engine(engine) {}
	static GpuEvaluator* make(const gsb_config& cfg, const arma::mat& X, const arma::vec& y, const std::vector<uint32_t>* seed,
	                          VerbosityLevel verbosity);
	static void check(gsb_engine* e, int rc);
	std::shared_ptr<Engine> engine;
};

#endif
