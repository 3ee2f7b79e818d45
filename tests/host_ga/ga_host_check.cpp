// tests/host_ga/ga_host_check.cpp -- TEST INFRASTRUCTURE (CPU only, no GPU, no CUDA).
//
// The host side of the drop-in -- gaselect_b200/host/BatchedPopulation (speculate, confirm or replay; chromosome bit parts
// as cache keys and engine input) and GpuEvaluator -- against the reference's OWN drivers, SingleThreadPopulation and
// MultiThreadedPopulation, compiled from /root/reference/src where they lie.  The CUDA engine behind the C ABI is replaced
// by the stub below (a deterministic fake fitness of the selected columns, some subsets "throwing"); the reference's
// drivers run on an Evaluator subclass that computes the same fake fitness.  Same seed => the same accept / reject
// decisions, elite, final population and fitness evolution, bit for bit -- with gaselect's default badSolutionThreshold
// (no rejections: every speculative pass is confirmed) and with a tight one (real rejections: passes are replayed),
// with and without duplicate elimination, for one and for several (virtual) threads.
#include <RcppArmadillo.h>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "config.h"
#include "RNG.h"
#include "ShuffledSet.h"
#include "Control.h"
#include "Chromosome.h"
#include "Evaluator.h"
#include "Population.h"
#include "SingleThreadPopulation.h"
#include "MultiThreadedPopulation.h"

#include "BatchedPopulation.h"
#include "GpuEvaluator.h"

// ---- R runtime symbols the reference links against ------------------------------------------------------------------
extern "C" {
void Rprintf(const char*, ...) {}
void REprintf(const char*, ...) {}
void R_FlushConsole(void) {}
void R_CheckUserInterrupt(void) {}
Rboolean R_ToplevelExec(void (*fun)(void*), void* data) { fun(data); return TRUE; }
double R_pow_di(double x, int n) { // binary exponentiation, as R's nmath
	double result = 1.0;
	if (n == 0) return 1.0;
	const bool neg = n < 0;
	if (neg) n = -n;
	for (;;) {
		if (n & 1) result *= x;
		if (n >>= 1) x *= x; else break;
	}
	return neg ? 1.0 / result : result;
}
}

// ---- the fake fitness: a pure function of the ascending column list -------------------------------------------------
static double fakeFitness(const uint32_t* idx, int k, int* status) {
	if (k == 0) { *status = GSB_STATUS_EMPTY; return 0.0; }
	uint64_t h = 1469598103934665603ull;
	for (int i = 0; i < k; ++i) { h ^= idx[i]; h *= 1099511628211ull; }
	if (h % 41 == 0) { *status = GSB_STATUS_UNDERFLOW; return 0.0; } // the evaluator "throws" on 2-3 % of the subsets
	*status = (h % 37 == 0) ? GSB_STATUS_FLAGGED : GSB_STATUS_OK;     // near-tie flag: valid fitness, must not disturb anything
	return -(1.0 + static_cast<double>(h >> 11) * (1.0 / 9007199254740992.0)) * (1.0 + 0.01 * k);
}

// ---- stub of the C ABI (the entry points GpuEvaluator uses) ----------------------------------------------------------
struct gsb_engine { int n, p; };
static unsigned long long g_calls = 0, g_evals = 0;
extern "C" {
int gsb_create(gsb_engine** out, const gsb_config*) { *out = new gsb_engine(); return 0; }
void gsb_destroy(gsb_engine* e) { delete e; }
int gsb_last_error(const gsb_engine*, char* buf, size_t n) { if (n) buf[0] = 0; return 0; }
int gsb_device_count(int32_t* c) { *c = 1; return 0; }
int gsb_set_data(gsb_engine* e, const double*, const double*, int32_t n, int32_t p) { e->n = n; e->p = p; return 0; }
int gsb_init_segmentation(gsb_engine*, const uint32_t*) { return 0; }
int gsb_get_segmentation(const gsb_engine*, uint32_t*, uint32_t*, int32_t* nvec, int64_t* total) { *nvec = 0; *total = 0; return 0; }
int gsb_evaluate_indices(gsb_engine*, const uint32_t* idx, const uint32_t* off, int32_t npop, double* fit, int32_t* st) {
	for (int c = 0; c < npop; ++c) { int s = 0; fit[c] = fakeFitness(idx + off[c], static_cast<int>(off[c + 1] - off[c]), &s); st[c] = s; }
	++g_calls; g_evals += static_cast<unsigned long long>(npop);
	return 0;
}
int gsb_evaluate_bitsets(gsb_engine* e, const uint64_t* bits, int32_t words, int32_t npop, double* fit, int32_t* st) {
	// the layout of src/Chromosome.cpp:419-438: variable v <-> bit (unusedBits + v), LSB first from part 0
	const int unused = words * 64 - e->p;
	std::vector<uint32_t> idx;
	for (int c = 0; c < npop; ++c) {
		idx.clear();
		for (int w = 0; w < words; ++w) {
			uint64_t b = bits[static_cast<size_t>(c) * words + w];
			if (w == 0 && unused > 0) b &= ~((uint64_t(1) << unused) - 1);
			while (b) { idx.push_back(static_cast<uint32_t>(w * 64 + __builtin_ctzll(b) - unused)); b &= b - 1; }
		}
		int s = 0;
		fit[c] = fakeFitness(idx.data(), static_cast<int>(idx.size()), &s);
		st[c] = s;
	}
	++g_calls; g_evals += static_cast<unsigned long long>(npop);
	return 0;
}
}

// ---- the same fitness behind the reference's plugin interface --------------------------------------------------------
class HashEvaluator : public Evaluator {
public:
	HashEvaluator() : Evaluator(OFF) {}
	double evaluate(arma::uvec& cols) {
		std::vector<uint32_t> idx(cols.n_elem);
		for (arma::uword i = 0; i < cols.n_elem; ++i) idx[i] = static_cast<uint32_t>(cols[i]);
		int s = 0;
		const double f = fakeFitness(idx.data(), static_cast<int>(idx.size()), &s);
		if (s != GSB_STATUS_OK && s != GSB_STATUS_FLAGGED) throw Evaluator::EvaluatorException("fake failure");
		return f;
	}
	double evaluate(Chromosome& ch) {
		arma::uvec cols = ch.toColumnSubset();
		const double f = this->evaluate(cols);
		ch.setFitness(f);
		return f;
	}
	Evaluator* clone() const { return new HashEvaluator(); }
};

struct Outcome {
	std::vector<std::vector<unsigned long long> > subsets;
	std::vector<double> fitness;
	std::vector<double> evolution;
};

static Outcome collect(Population& pop) {
	Outcome o;
	const Population::SortedChromosomes result = pop.getResult();
	for (Population::SortedChromosomes::const_iterator it = result.begin(); it != result.end(); ++it) {
		const arma::uvec cols = it->toColumnSubset();
		std::vector<unsigned long long> s(cols.n_elem);
		for (arma::uword j = 0; j < cols.n_elem; ++j) s[j] = cols[j];
		o.subsets.push_back(s);
		o.fitness.push_back(it->getFitness());
	}
	o.evolution = pop.getFitnessEvolution();
	return o;
}

static bool same(const Outcome& a, const Outcome& b) {
	return a.subsets == b.subsets && a.fitness == b.fitness && a.evolution == b.evolution;
}

// usage: ga_host_check ref|batched THRESHOLD DUP THREADS
// prints one line: a digest of the outcome (every selected subset, every fitness and the fitness evolution, bit for bit) and
// the driver's statistics.  `ref` runs the reference's own driver (SingleThreadPopulation, or MultiThreadedPopulation with real
// threads -- whose hand-rolled condition-variable hand-over can miss a wake-up when an evaluation takes no time at all, so
// tests/test_host_ga.py runs the multi-threaded reference only through tests/host_ga/make_digests.py, with a watchdog, and
// compares against the committed digests), `batched` runs BatchedPopulation on GpuEvaluator over the stub engine.
int main(int argc, char** argv) {
	if (argc < 5) { std::fprintf(stderr, "usage: %s ref|batched THRESHOLD DUP THREADS\n", argv[0]); return 2; }
	const bool batched = std::string(argv[1]) == "batched";
	const double threshold = std::atof(argv[2]);
	const int dup = std::atoi(argv[3]), nthreads = std::atoi(argv[4]);
	// GA_HOST_SCALE="P POP GENS MINV MAXV ELITE" (profiling aid: the host side alone at another scale, e.g. config 3's
	// "1000 500 50 10 200 10"); prints the wall time of run() on stderr
	int p = 200, popSize = 60, gens = 12, elite = 6, minV = 3, maxV = 25;
	if (const char* sc = std::getenv("GA_HOST_SCALE")) std::sscanf(sc, "%d %d %d %d %d %d", &p, &popSize, &gens, &minV, &maxV, &elite);
	const int n = 40;
	arma::mat X(n, p);
	arma::vec y(n);
	for (int i = 0; i < n; ++i) { y[i] = i; for (int j = 0; j < p; ++j) X(i, j) = (i * 31 + j * 17) % 101; }
	std::vector<uint32_t> seed(624);
	for (int i = 0; i < 624; ++i) seed[i] = 2654435761u * static_cast<uint32_t>(i + 1) + 12345u;
	Control ctrl(p, popSize, gens, elite, minV, maxV, 0.05, nthreads, dup, threshold, 1e-6, 0, static_cast<CrossoverType>(0), static_cast<FitnessScaling>(0), OFF);
	Outcome o;
	unsigned long long calls = 0, evals = 0, passes = 0, confirmed = 0, flagged = 0;
	const auto t0 = std::chrono::steady_clock::now();
	if (batched) {
		std::unique_ptr<GpuEvaluator> gpu(GpuEvaluator::makePLS(X, y, 2, 5, seed, OFF, 3));
		BatchedPopulation pop(ctrl, *gpu, seed);
		pop.run();
		o = collect(pop);
		calls = pop.stats().engineCalls; evals = pop.stats().engineEvaluations; passes = pop.stats().passes; confirmed = pop.stats().confirmed;
		flagged = gpu->flaggedCount();
	} else {
		HashEvaluator hash;
		if (nthreads > 1) { MultiThreadedPopulation pop(ctrl, hash, seed); pop.run(); o = collect(pop); }
		else { SingleThreadPopulation pop(ctrl, hash, seed); pop.run(); o = collect(pop); }
	}
	if (std::getenv("GA_HOST_SCALE")) {
		std::fprintf(stderr, "run + collect: %.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
	}
	// FNV-1a over the exact bits of the outcome
	uint64_t h = 1469598103934665603ull;
	auto mix = [&](const void* ptr, size_t bytes) { const unsigned char* b = static_cast<const unsigned char*>(ptr); for (size_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; } };
	for (size_t i = 0; i < o.subsets.size(); ++i) {
		const unsigned long long len = o.subsets[i].size();
		mix(&len, sizeof(len));
		mix(o.subsets[i].data(), len * sizeof(unsigned long long));
	}
	mix(o.fitness.data(), o.fitness.size() * sizeof(double));
	mix(o.evolution.data(), o.evolution.size() * sizeof(double));
	std::printf("digest %016llx chromosomes %zu best %.17g calls %llu evaluations %llu passes %llu confirmed %llu flagged %llu\n",
	            static_cast<unsigned long long>(h), o.subsets.size(), o.fitness.empty() ? 0.0 : o.fitness.back(), calls, evals, passes, confirmed, flagged);
	return o.subsets.empty() ? 1 : 0;
}
