// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A plain-C entry layer around the reference's OWN classes, compiled out of tree
// from /root/reference/src (never copied into this repo) against the stand-in
// oracle/standin/RcppArmadillo.h.  The resulting oracle/_ref/libgaselect_ref.so is
//   * the authority the CPU restatement (oracle/gaselect_oracle.c) is pinned against,
//   * the generator of the golden vectors under tests/golden/, and
//   * the "reference" arm / cpu_baseline of bench.py (kind = "reference").
// It bypasses only src/GenAlg.cpp (Rcpp::List glue, needs a real R): objects are
// constructed exactly as src/GenAlg.cpp:80-94,99-103,110-181,251-325 constructs them.
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.

#include <RcppArmadillo.h>

#include <atomic>
#include <chrono>
#include <memory>
#include <thread>

#include "config.h"
#include "RNG.h"
#include "ShuffledSet.h"
#include "Control.h"
#include "Chromosome.h"
#include "Evaluator.h"
#include "PLS.h"
#include "PLSEvaluator.h"
#include "BICEvaluator.h"
#include "LMEvaluator.h"
#include "Population.h"
#include "SingleThreadPopulation.h"
#include "MultiThreadedPopulation.h"

// ---------------------------------------------------------------------------
// R runtime symbols the reference links against
// ---------------------------------------------------------------------------
static bool shim_verbose() {
	static const bool v = (std::getenv("GASELECT_REF_VERBOSE") != NULL);
	return v;
}

extern "C" {

void Rprintf(const char* fmt, ...) {
	if (!shim_verbose()) return;
	va_list ap; va_start(ap, fmt); std::vfprintf(stdout, fmt, ap); va_end(ap);
}
void REprintf(const char* fmt, ...) {
	if (!shim_verbose()) return;
	va_list ap; va_start(ap, fmt); std::vfprintf(stderr, fmt, ap); va_end(ap);
}
void R_FlushConsole(void) {}
void R_CheckUserInterrupt(void) {}
Rboolean R_ToplevelExec(void (*fun)(void*), void* data) { fun(data); return TRUE; }

// x^n for integer n by binary exponentiation, as R's nmath does (NOT std::pow):
// multiply the result in for every set bit of n while squaring the base.
double R_pow_di(double x, int n) {
	double result = 1.0;
	if (std::isnan(x)) return x;
	if (n == 0) return result;
	if (!std::isfinite(x)) return std::pow(x, static_cast<double>(n));
	const bool invert = (n < 0);
	unsigned int e = invert ? static_cast<unsigned int>(-static_cast<long long>(n)) : static_cast<unsigned int>(n);
	while (true) {
		if (e & 1u) result *= x;
		e >>= 1;
		if (e == 0) break;
		x *= x;
	}
	return invert ? 1.0 / result : result;
}

} // extern "C"

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
namespace {

std::vector<uint32_t> seed_vector(const uint32_t* seedvec) {
	return std::vector<uint32_t>(seedvec, seedvec + RNG::SEED_SIZE);
}

arma::mat wrap_matrix(const double* X, int n, int p) {
	arma::mat m(static_cast<arma::uword>(n), static_cast<arma::uword>(p));
	std::memcpy(m.memptr(), X, sizeof(double) * static_cast<size_t>(n) * static_cast<size_t>(p));
	return m;
}

arma::vec wrap_vector(const double* y, int n) {
	arma::vec v(static_cast<arma::uword>(n));
	std::memcpy(v.memptr(), y, sizeof(double) * static_cast<size_t>(n));
	return v;
}

int status_of(const char* what) {
	const std::string w(what);
	if (w.find("empty") != std::string::npos) return 1;
	if (w.find("underflow") != std::string::npos) return 2;
	return 3;
}

// Counts evaluate() calls; clones share the counter (the reference's own
// PLSEvaluator::counter is debug-only and racy, src/PLSEvaluator.cpp:24,77).
class CountingEvaluator : public Evaluator {
public:
	CountingEvaluator(Evaluator* inner, bool ownsInner, std::shared_ptr<std::atomic<uint64_t> > counter)
		: Evaluator(OFF), inner(inner), ownsInner(ownsInner), counter(counter) {}
	~CountingEvaluator() { if (ownsInner) delete inner; }
	double evaluate(arma::uvec& columnSubset) { ++(*counter); return inner->evaluate(columnSubset); }
	double evaluate(Chromosome& ch) { ++(*counter); return inner->evaluate(ch); }
	Evaluator* clone() const { return new CountingEvaluator(inner->clone(), true, counter); }
	std::vector<arma::uvec> getSegmentation() const { return inner->getSegmentation(); }
private:
	Evaluator* inner;
	bool ownsInner;
	std::shared_ptr<std::atomic<uint64_t> > counter;
};

} // namespace

extern "C" {

// ---------------------------------------------------------------------------
// RNG / ShuffledSet known-answer entry points
// ---------------------------------------------------------------------------
void ref_rng_u32_stream(uint32_t seed, int n, uint32_t* out) {
	RNG rng(seed);
	for (int i = 0; i < n; ++i) out[i] = rng();
}

// the two-stage seeding of src/GenAlg.cpp:99-103
void ref_seedvec(uint32_t seed, uint32_t* out624) {
	RNG rng(seed);
	for (uint32_t i = 0; i < RNG::SEED_SIZE; ++i) out624[i] = rng();
}

void ref_rng_vec_stream(const uint32_t* seedvec, int n, uint32_t* out) {
	RNG rng(seed_vector(seedvec));
	for (int i = 0; i < n; ++i) out[i] = rng();
}

void ref_rng_vec_uniform(const uint32_t* seedvec, double lo, double hi, int n, double* out) {
	RNG rng(seed_vector(seedvec));
	for (int i = 0; i < n; ++i) out[i] = rng(lo, hi);
}

// `ncalls` consecutive shuffleAll() calls on ONE ShuffledSet (state persists, as in
// src/PLSEvaluator.cpp:98,145); out holds ncalls*size entries
void ref_shuffle_all(const uint32_t* seedvec, int size, int ncalls, uint32_t* out) {
	RNG rng(seed_vector(seedvec));
	ShuffledSet set(static_cast<arma::uword>(size));
	for (int c = 0; c < ncalls; ++c) {
		const arma::uvec& s = set.shuffleAll(rng);
		for (int i = 0; i < size; ++i) out[c * size + i] = static_cast<uint32_t>(s[i]);
	}
}

// ---------------------------------------------------------------------------
// Evaluators (constructed as src/GenAlg.cpp:110-163 / :251-303 does)
// ---------------------------------------------------------------------------
void* ref_pls_evaluator_create(const double* X, const double* y, int n, int p, int numReplications, int maxNComp,
                               const uint32_t* seedvec, int innerSegments, int outerSegments, double testSetSize,
                               double sdfact, double complexityPenalty) {
	try {
		arma::mat Xm = wrap_matrix(X, n, p);
		arma::vec yv = wrap_vector(y, n);
		std::unique_ptr<PLS> pls = PLS::getInstance(SIMPLS, Xm, yv);
		return static_cast<Evaluator*>(new PLSEvaluator(std::move(pls), static_cast<uint16_t>(numReplications),
			static_cast<uint16_t>(maxNComp), seed_vector(seedvec), OFF, static_cast<uint16_t>(innerSegments),
			static_cast<uint16_t>(outerSegments), testSetSize, sdfact, complexityPenalty));
	} catch (...) {
		return NULL;
	}
}

void* ref_fit_evaluator_create(const double* X, const double* y, int n, int p, int maxNComp, const uint32_t* seedvec,
                               int numSegments, int statistic, double sdfact) {
	try {
		arma::mat Xm = wrap_matrix(X, n, p);
		arma::vec yv = wrap_vector(y, n);
		std::unique_ptr<PLS> pls = PLS::getInstance(SIMPLS, Xm, yv);
		return static_cast<Evaluator*>(new BICEvaluator(std::move(pls), static_cast<uint16_t>(maxNComp), seed_vector(seedvec), OFF,
			static_cast<uint16_t>(numSegments), static_cast<BICEvaluator::Statistic>(statistic), sdfact));
	} catch (...) {
		return NULL;
	}
}

void* ref_lm_evaluator_create(const double* X, const double* y, int n, int p, int statistic) {
	try {
		arma::mat Xm = wrap_matrix(X, n, p);
		arma::vec yv = wrap_vector(y, n);
		return static_cast<Evaluator*>(new LMEvaluator(Xm, yv, static_cast<LMEvaluator::Statistic>(statistic), OFF));
	} catch (...) {
		return NULL;
	}
}

void ref_evaluator_destroy(void* h) { delete static_cast<Evaluator*>(h); }

int ref_segmentation_count(void* h) {
	return static_cast<int>(static_cast<Evaluator*>(h)->getSegmentation().size());
}

long long ref_segmentation_total(void* h) {
	const std::vector<arma::uvec> seg = static_cast<Evaluator*>(h)->getSegmentation();
	long long total = 0;
	for (size_t i = 0; i < seg.size(); ++i) total += static_cast<long long>(seg[i].n_elem);
	return total;
}

// rows: concatenated index vectors; offsets: count+1 entries
void ref_segmentation_get(void* h, uint32_t* rows, uint32_t* offsets) {
	const std::vector<arma::uvec> seg = static_cast<Evaluator*>(h)->getSegmentation();
	uint32_t pos = 0;
	for (size_t i = 0; i < seg.size(); ++i) {
		offsets[i] = pos;
		for (arma::uword j = 0; j < seg[i].n_elem; ++j) rows[pos++] = static_cast<uint32_t>(seg[i][j]);
	}
	offsets[seg.size()] = pos;
}

// The batch loop of src/GenAlg.cpp:312-325 over `nsub` column subsets (CSR: idx/offsets,
// ascending 0-based columns), split into contiguous slices over `nthreads` host threads,
// each with its own clone() -- the same sharing rule as src/MultiThreadedPopulation.cpp:311.
// status: 0 ok, 1 empty subset, 2 underflow, 3 other EvaluatorException.
// Returns the wall time of the loop in nanoseconds through *elapsed_ns (may be NULL).
int ref_evaluate(void* h, const uint32_t* idx, const uint32_t* offsets, int nsub, int nthreads, double* fitness,
                 int32_t* status, long long* elapsed_ns) {
	Evaluator* base = static_cast<Evaluator*>(h);
	if (nthreads < 1) nthreads = 1;
	if (nthreads > nsub) nthreads = (nsub > 0) ? nsub : 1;

	std::vector<Evaluator*> evals(static_cast<size_t>(nthreads), NULL);
	evals[0] = base;
	for (int t = 1; t < nthreads; ++t) evals[t] = base->clone();

	auto work = [&](int t) {
		const int per = nsub / nthreads, rem = nsub % nthreads;
		const int begin = t * per + (t < rem ? t : rem);
		const int end = begin + per + (t < rem ? 1 : 0);
		for (int s = begin; s < end; ++s) {
			const uint32_t k = offsets[s + 1] - offsets[s];
			arma::uvec cols(static_cast<arma::uword>(k));
			for (uint32_t j = 0; j < k; ++j) cols[j] = idx[offsets[s] + j];
			try {
				fitness[s] = evals[t]->evaluate(cols);
				status[s] = 0;
			} catch (const Evaluator::EvaluatorException& ee) {
				fitness[s] = 0.0;
				status[s] = status_of(ee.what());
			}
		}
	};

	const auto t0 = std::chrono::steady_clock::now();
	if (nthreads == 1) {
		work(0);
	} else {
		std::vector<std::thread> threads;
		for (int t = 1; t < nthreads; ++t) threads.emplace_back(work, t);
		work(0);
		for (size_t t = 0; t < threads.size(); ++t) threads[t].join();
	}
	const auto t1 = std::chrono::steady_clock::now();
	if (elapsed_ns) *elapsed_ns = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();

	for (int t = 1; t < nthreads; ++t) delete evals[t];
	return 0;
}

// Raw SIMPLS on a (rows, cols) view: the hidden C_simpls entry (src/GenAlg.cpp:351-371),
// extended with a row view.  coef: k x ncomp column-major, b0: ncomp.  nrows == 0 -> all rows.
int ref_simpls(const double* X, const double* y, int n, int p, const uint32_t* cols, int k, const uint32_t* rows,
               int nrows, int ncomp, double* coef, double* b0) {
	try {
		std::unique_ptr<PLS> pls = PLS::getInstance(SIMPLS, wrap_matrix(X, n, p), wrap_vector(y, n));
		arma::uvec c(static_cast<arma::uword>(k));
		for (int j = 0; j < k; ++j) c[j] = cols[j];
		pls->viewSelectColumns(c);
		if (nrows > 0) {
			arma::uvec r(static_cast<arma::uword>(nrows));
			for (int j = 0; j < nrows; ++j) r[j] = rows[j];
			pls->viewSelectRows(r);
		} else {
			pls->viewSelectAllRows();
		}
		pls->fit(static_cast<uint16_t>(ncomp));
		const arma::mat& B = pls->getCoefficients();
		const arma::vec& I = pls->getIntercepts();
		if (static_cast<int>(B.n_cols) != ncomp) return 2;
		std::memcpy(coef, B.memptr(), sizeof(double) * B.n_elem);
		std::memcpy(b0, I.memptr(), sizeof(double) * I.n_elem);
		return 0;
	} catch (...) {
		return 1;
	}
}

// ---------------------------------------------------------------------------
// GA driver: Control + Population exactly as src/GenAlg.cpp:80-94,173-181,205-222
// ---------------------------------------------------------------------------
// Outputs: result chromosomes (elite U last generation, std::set order = ascending fitness)
// as CSR column lists + fitness; fitnessEvolution (3 doubles per generation: best, mean, sd).
// Capacities are given by the caller; returns 0 on success, 1 if a capacity was too small,
// 2 on an exception.
int ref_ga_run(void* evaluator, int chromosomeSize, int popSize, int numGenerations, int elitism, int minVariables,
               int maxVariables, double mutationProb, int numThreads, int maxDuplicateEliminationTries,
               double badSolutionThreshold, double convergenceThreshold, int convergenceGenerations, int crossover,
               int fitnessScaling, const uint32_t* seedvec,
               int resultCapacity, long long idxCapacity, int* nResult, uint32_t* resIdx, uint32_t* resOffsets,
               double* resFitness, int evolutionCapacity, int* nEvolution, double* fitnessEvolution,
               unsigned long long* evalCount, long long* elapsed_ns) {
	try {
		Control ctrl(static_cast<uint16_t>(chromosomeSize), static_cast<uint16_t>(popSize), static_cast<uint16_t>(numGenerations),
			static_cast<uint16_t>(elitism), static_cast<uint16_t>(minVariables), static_cast<uint16_t>(maxVariables),
			mutationProb, static_cast<uint16_t>(numThreads), static_cast<uint16_t>(maxDuplicateEliminationTries),
			badSolutionThreshold, convergenceThreshold, static_cast<uint16_t>(convergenceGenerations),
			static_cast<CrossoverType>(crossover), static_cast<FitnessScaling>(fitnessScaling), OFF);
		const std::vector<uint32_t> seed = seed_vector(seedvec);

		std::shared_ptr<std::atomic<uint64_t> > counter(new std::atomic<uint64_t>(0));
		CountingEvaluator counting(static_cast<Evaluator*>(evaluator), false, counter);

		std::unique_ptr<Population> pop;
		if (numThreads > 1) {
			pop.reset(new MultiThreadedPopulation(ctrl, counting, seed));
		} else {
			pop.reset(new SingleThreadPopulation(ctrl, counting, seed));
		}

		const auto t0 = std::chrono::steady_clock::now();
		pop->run();
		const auto t1 = std::chrono::steady_clock::now();
		if (elapsed_ns) *elapsed_ns = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
		if (evalCount) *evalCount = counter->load();

		Population::SortedChromosomes result = pop->getResult();
		int r = 0;
		long long pos = 0;
		int rc = 0;
		for (Population::SortedChromosomes::iterator it = result.begin(); it != result.end(); ++it) {
			const arma::uvec cols = it->toColumnSubset();
			if (r >= resultCapacity || pos + static_cast<long long>(cols.n_elem) > idxCapacity) { rc = 1; break; }
			resOffsets[r] = static_cast<uint32_t>(pos);
			for (arma::uword j = 0; j < cols.n_elem; ++j) resIdx[pos++] = static_cast<uint32_t>(cols[j]);
			resFitness[r] = it->getFitness();
			++r;
		}
		resOffsets[r] = static_cast<uint32_t>(pos);
		*nResult = r;

		const std::vector<double>& evo = pop->getFitnessEvolution();
		const int ne = static_cast<int>(evo.size());
		*nEvolution = (ne <= evolutionCapacity) ? ne : evolutionCapacity;
		for (int i = 0; i < *nEvolution; ++i) fitnessEvolution[i] = evo[i];
		if (ne > evolutionCapacity) rc = 1;
		return rc;
	} catch (...) {
		return 2;
	}
}

// ---------------------------------------------------------------------------
// Chromosome known-answer entry points (host GA pieces)
// ---------------------------------------------------------------------------
// Creates two random chromosomes A, B with RNG(seedvec) and a fresh ShuffledSet(p), mates
// them, mutates both children, and returns the four column lists + the next RNG draw.
// out lists are CSR: offsets has 7 entries for [A, B, child1, child2, child1mut, child2mut].
int ref_chromosome_kat(int p, int popSize, int minVariables, int maxVariables, double mutationProb, int crossover,
                       const uint32_t* seedvec, uint32_t* idx, uint32_t* offsets, int* mutated, uint32_t* nextDraw) {
	try {
		Control ctrl(static_cast<uint16_t>(p), static_cast<uint16_t>(popSize), 8, 5, static_cast<uint16_t>(minVariables),
			static_cast<uint16_t>(maxVariables), mutationProb, 1, 0, 2.0, 1e-6, 0,
			static_cast<CrossoverType>(crossover), NONE, OFF);
		RNG rng(seed_vector(seedvec));
		ShuffledSet set(static_cast<arma::uword>(p));
		Chromosome a(ctrl, set, rng), b(ctrl, set, rng);
		Chromosome c1(ctrl, set, rng, false), c2(ctrl, set, rng, false);
		uint32_t pos = 0;
		int slot = 0;
		auto emit = [&](const Chromosome& ch) {
			const arma::uvec cols = ch.toColumnSubset();
			offsets[slot++] = pos;
			for (arma::uword j = 0; j < cols.n_elem; ++j) idx[pos++] = static_cast<uint32_t>(cols[j]);
		};
		emit(a); emit(b);
		a.mateWith(b, rng, c1, c2);
		emit(c1); emit(c2);
		mutated[0] = c1.mutate(rng) ? 1 : 0;
		mutated[1] = c2.mutate(rng) ? 1 : 0;
		emit(c1); emit(c2);
		offsets[slot] = pos;
		*nextDraw = rng();
		return 0;
	} catch (...) {
		return 2;
	}
}

int ref_hardware_threads(void) { return static_cast<int>(std::thread::hardware_concurrency()); }

} // extern "C"
