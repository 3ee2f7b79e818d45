// BatchedPopulation.h -- a Population (src/Population.h:33-420 of dakep/gaselect) that hands a whole
// generation of candidate chromosomes to the B200 engine in ONE evaluateBatch call and still
// produces exactly the run of the reference's drivers: SingleThreadPopulation
// (src/SingleThreadPopulation.cpp:43-251) when ctrl.numThreads <= 1, MultiThreadedPopulation
// (src/MultiThreadedPopulation.cpp:255-465) otherwise -- the latter's result depends on numThreads (every
// worker owns a random stream and a slice of the population, :301-323,:471-482), so the workers are
// reproduced as "virtual threads": same seeds, same slices, same per-slice loops, run one after the other
// on the host (they never interact inside a generation).  Same accept / reject decisions, same elite, same result.
//
// Why it is not simply "collect, evaluate, continue": inside a generation the reference interleaves
// RNG draws with decisions that depend on each child's fitness (accept iff fitness > cutoff,
// src/SingleThreadPopulation.cpp:171-184; EvaluatorException -> the slot is drawn again, :185-189), so
// the set of candidates is only known after their fitness is.  SURVEY.md Appendix F protocol:
//   1. snapshot the generation's state (RNG, ShuffledSet, child slots, elite, counters);
//   2. run the generation against the fitness cache; a chromosome that is not cached yet is recorded
//      and provisionally treated as accepted;
//   3. if anything was recorded: evaluate all of it in one engine call.  Every provisional decision is then checked
//      against the real fitness (accepted iff fitness > cutoff, no EvaluatorException): when all of them stand, the
//      speculative pass WAS the reference's generation up to the fitness-dependent bookkeeping (child fitness, the
//      generation minimum, the elite insertions), which is redone in the recorded order -- no replay.  Otherwise the
//      snapshot is restored and the pass runs again against the now larger cache (go to 2);
//   4. a pass without a miss IS the reference's generation (fitness is a pure function of the subset).
// At gaselect's defaults no child is rejected, so every generation costs one engine call and one host pass.
//
// Compiled inside the gaselect package next to the reference sources (it includes the package's own
// Population.h); selected at src/GenAlg.cpp:173-181 in place of SingleThreadPopulation.  See INTEGRATION.md.
#ifndef GASELECT_B200_BATCHED_POPULATION_H
#define GASELECT_B200_BATCHED_POPULATION_H

#include <string>
#include <unordered_map>
#include <vector>

#include "Population.h" // the reference's base class, on the package's include path

#include "ChromosomeBits.h"
#include "GpuEvaluator.h"

class BatchedPopulation : public Population {
public:
	BatchedPopulation(const Control& ctrl, GpuEvaluator& evaluator, const std::vector<uint32_t>& seed);
	~BatchedPopulation();

	void run(); // src/Population.h:96

	// bookkeeping for benchmarks: chromosomes scored by the engine, engine calls, replayed passes
	struct Stats {
		unsigned long long engineEvaluations = 0, engineCalls = 0, passes = 0, cacheHits = 0;
		unsigned long long confirmed = 0; // speculative passes whose every provisional decision stood (no replay)
	};
	const Stats& stats() const { return this->counters; }

	// optional cooperative interrupt (the reference polls R_CheckUserInterrupt, src/SingleThreadPopulation.cpp:30-38)
	void setInterruptCheck(bool (*check)()) { this->interruptCheck = check; }

private:
	enum Lookup { HIT_OK, HIT_FAILED, MISS };
	struct Entry { double fitness; int status; };
	struct LoopState; // everything a generation mutates outside the base class
	struct Stream;    // one (virtual) worker of the multi-threaded driver

	// a decision taken on a chromosome whose fitness was not known yet
	struct Provisional {
		Chromosome* ch;
		size_t pendingIndex; // into pending / lastFitness / lastStatus
		double cutoff;       // the decision stands iff fitness > cutoff ...
		bool forced;         // ... or the child was kept regardless (too many discarded)
	};

	Lookup lookup(Chromosome& ch, size_t* pendingIndex = nullptr);
	void flush();
	bool confirmProvisional();
	void snapshotSlots();
	void restoreSlots();
	void generateInitial(LoopState& st);
	void mateGeneration(LoopState& st);
	void runSingle();
	void runVirtualThreads();
	void generateInitialSlice(Stream& w);
	void mateSlice(Stream& w, double sumFitness);
	bool pollInterrupt() { return this->interruptCheck != nullptr && this->interruptCheck(); }

	GpuEvaluator& gpu;
	// exact keys: the bytes of a chromosome's bit parts (ChromosomeBits.h); the parts also go to the engine as they are
	std::unordered_map<std::string, Entry> cache;
	std::vector<uint64_t> pendingParts;                 // `words` parts per pending chromosome
	std::vector<std::string> pendingKeys;
	std::unordered_map<std::string, size_t> pendingSet; // key -> index into pendingKeys / pendingParts
	size_t words = 0;                                   // parts per chromosome: ceil(chromosomeSize / 64)
	std::vector<double> lastFitness;                    // results of the most recent flush, by pending index
	std::vector<int> lastStatus;
	std::vector<Provisional> provisional;               // this pass's decisions on unknown fitness
	std::vector<Chromosome*> acceptedOrder;             // single-thread driver: children in the order they were accepted
	std::vector<Chromosome*> slots; // the generation being built (owned)
	std::vector<Chromosome> slotsBefore; // snapshot of the slots (storage kept across generations)
	bool (*interruptCheck)() = nullptr;
	Stats counters;
};

#endif
