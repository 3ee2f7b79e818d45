// BatchedPopulation.cpp -- see BatchedPopulation.h.
#include "BatchedPopulation.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <utility>

namespace {

// exact key: the bytes of the chromosome's bit parts (equal parts <=> equal chromosomes, src/Chromosome.cpp:440-442)
std::string keyOf(const std::vector<IntChromosome>& parts) {
	return std::string(reinterpret_cast<const char*>(parts.data()), parts.size() * sizeof(IntChromosome));
}

} // namespace

// Everything one generation reads and writes besides the (protected) elite of the base class.
struct BatchedPopulation::LoopState {
	RNG rng;
	ShuffledSet shuffledSet;
	double sumFitness = 0.0;
	double minFitness = 0.0;
	uint8_t tries1 = 0, tries2 = 0;           // persist across generations like the reference's locals
	std::pair<bool, bool> duplicated{false, false};
	LoopState(const std::vector<uint32_t>& seed, arma::uword chromosomeSize) : rng(seed), shuffledSet(chromosomeSize) {}
};

BatchedPopulation::BatchedPopulation(const Control& ctrl, GpuEvaluator& evaluator, const std::vector<uint32_t>& seed)
	: Population(ctrl, evaluator, seed), gpu(evaluator) {}

BatchedPopulation::~BatchedPopulation() {
	for (size_t i = 0; i < this->slots.size(); ++i) delete this->slots[i];
}

BatchedPopulation::Lookup BatchedPopulation::lookup(Chromosome& ch, size_t* pendingIndex) {
	const std::vector<IntChromosome>& parts = gsb_host::partsOf(ch);
	std::string key = keyOf(parts);
	std::unordered_map<std::string, Entry>::const_iterator hit = this->cache.find(key);
	if (hit != this->cache.end()) {
		++this->counters.cacheHits;
		if (hit->second.status != 0) return HIT_FAILED; // the reference's evaluator would have thrown
		ch.setFitness(hit->second.fitness);
		return HIT_OK;
	}
	std::unordered_map<std::string, size_t>::const_iterator known = this->pendingSet.find(key);
	size_t index;
	if (known == this->pendingSet.end()) {
		index = this->pendingKeys.size();
		this->words = parts.size();
		this->pendingSet.emplace(key, index);
		this->pendingKeys.push_back(std::move(key));
		this->pendingParts.insert(this->pendingParts.end(), parts.begin(), parts.end());
	} else {
		index = known->second;
	}
	if (pendingIndex) *pendingIndex = index;
	ch.setFitness(0.0); // provisional: confirmed or replayed after the engine call
	return MISS;
}

void BatchedPopulation::flush() {
	const size_t count = this->pendingKeys.size();
	this->gpu.evaluateBitsets(this->pendingParts.data(), static_cast<int>(this->words), count, this->lastFitness, this->lastStatus);
	for (size_t i = 0; i < count; ++i) {
		Entry e = {this->lastFitness[i], this->lastStatus[i]};
		this->cache.emplace(std::move(this->pendingKeys[i]), e);
	}
	this->counters.engineEvaluations += count;
	++this->counters.engineCalls;
	this->pendingParts.clear();
	this->pendingKeys.clear();
	this->pendingSet.clear();
}

// After a flush: did every decision this pass took on a yet unknown fitness turn out right?  If so the chromosomes get
// their fitness and the pass stands (the caller redoes the fitness-dependent bookkeeping); if not the caller replays.
bool BatchedPopulation::confirmProvisional() {
	for (size_t i = 0; i < this->provisional.size(); ++i) {
		const Provisional& d = this->provisional[i];
		if (this->lastStatus[d.pendingIndex] != 0) return false; // EvaluatorException: the slot would have been drawn again
		if (!d.forced && !(this->lastFitness[d.pendingIndex] > d.cutoff)) return false; // the child would have been rejected
	}
	for (size_t i = 0; i < this->provisional.size(); ++i) {
		this->provisional[i].ch->setFitness(this->lastFitness[this->provisional[i].pendingIndex]);
	}
	++this->counters.confirmed;
	return true;
}

// snapshot of the generation being built, reusing the storage of the previous one (no allocation per chromosome)
void BatchedPopulation::snapshotSlots() {
	if (this->slotsBefore.size() != this->slots.size()) {
		this->slotsBefore.clear();
		this->slotsBefore.reserve(this->slots.size());
		for (size_t i = 0; i < this->slots.size(); ++i) this->slotsBefore.push_back(*this->slots[i]);
	} else {
		for (size_t i = 0; i < this->slots.size(); ++i) this->slotsBefore[i] = *this->slots[i];
	}
}

void BatchedPopulation::restoreSlots() {
	for (size_t i = 0; i < this->slots.size(); ++i) *this->slots[i] = this->slotsBefore[i];
}

// One pass over the initial population (src/SingleThreadPopulation.cpp:73-104).
void BatchedPopulation::generateInitial(LoopState& st) {
	while (this->slots.size() < this->ctrl.populationSize && !this->interrupted) {
		Chromosome* candidate = new Chromosome(this->ctrl, st.shuffledSet, st.rng);
		bool known = false;
		for (size_t i = 0; i < this->slots.size() && !known; ++i) known = (*this->slots[i] == *candidate);
		bool keep = false;
		if (!known) {
			size_t pendingIndex = 0;
			const Lookup res = this->lookup(*candidate, &pendingIndex);
			keep = res != HIT_FAILED;
			if (keep) {
				if (res == MISS) {
					const Provisional d = {candidate, pendingIndex, -std::numeric_limits<double>::infinity(), false};
					this->provisional.push_back(d);
				}
				if (candidate->getFitness() < st.minFitness) st.minFitness = candidate->getFitness();
				this->addChromosomeToElite(*candidate);
				this->acceptedOrder.push_back(candidate);
			}
		}
		if (keep) this->slots.push_back(candidate); else delete candidate;
		if (this->pollInterrupt()) this->interrupted = true;
	}
}

// One pass over one generation (the mating loop of src/SingleThreadPopulation.cpp:129-233).
void BatchedPopulation::mateGeneration(LoopState& st) {
	const uint32_t maxDiscarded = Population::MAX_DISCARDED_SOLUTIONS_RATIO * this->ctrl.populationSize;
	uint32_t discarded1 = 0, discarded2 = 0;
	st.minFitness = 0.0;
	ChVecIt front = this->slots.begin();   // child 1 fills from the front,
	ChVecRIt back = this->slots.rbegin();  // child 2 from the back
	while (front < back.base() && !this->interrupted) {
		const bool twoChildren = (front + 1 != back.base());
		Chromosome* parentA = this->drawChromosomeFromCurrentGeneration(st.rng(0.0, st.sumFitness));
		Chromosome* parentB = this->drawChromosomeFromCurrentGeneration(st.rng(0.0, st.sumFitness), parentA);
		parentA->mateWith(*parentB, st.rng, **front, **back);
		const double betterParent = std::max(parentA->getFitness(), parentB->getFitness());
		(*front)->mutate(st.rng);
		if (twoChildren) (*back)->mutate(st.rng);
		// reject children clearly worse than the better parent (the reference calls it minParentFitness)
		const double cutoff = betterParent - this->ctrl.badSolutionThreshold * std::fabs(betterParent);
		if (this->ctrl.maxDuplicateEliminationTries > 0) {
			st.duplicated = Population::checkDuplicated(this->slots.begin(), this->slots.rbegin(), front, back);
		}
		for (int which = 0; which < 2; ++which) {
			if (which == 1 && !twoChildren) break;
			const bool dup = (which == 0) ? st.duplicated.first : st.duplicated.second;
			uint8_t& tries = (which == 0) ? st.tries1 : st.tries2;
			uint32_t& discarded = (which == 0) ? discarded1 : discarded2;
			Chromosome& child = (which == 0) ? **front : **back;
			if (dup && !(++tries > this->ctrl.maxDuplicateEliminationTries)) continue; // try this slot again
			if (dup) child.randomlyReset(st.rng, st.shuffledSet); // too many duplicate tries: random restart
			tries = 0;
			size_t pendingIndex = 0;
			const Lookup res = this->lookup(child, &pendingIndex);
			if (res == HIT_FAILED) continue; // EvaluatorException: the slot is drawn again
			const bool accepted = (res == MISS) || (child.getFitness() > cutoff) || (discarded > maxDiscarded);
			if (accepted) {
				if (res == MISS) {
					const Provisional d = {&child, pendingIndex, cutoff, discarded > maxDiscarded};
					this->provisional.push_back(d);
				}
				if (child.getFitness() < st.minFitness) st.minFitness = child.getFitness();
				this->addChromosomeToElite(child);
				this->acceptedOrder.push_back(&child);
				if (which == 0) ++front; else ++back;
				discarded = 0;
			} else {
				++discarded;
			}
		}
		if (this->pollInterrupt()) this->interrupted = true;
	}
}

void BatchedPopulation::run() {
	if (this->ctrl.numThreads > 1) this->runVirtualThreads(); else this->runSingle();
}

void BatchedPopulation::runSingle() {
	LoopState st(this->seed, this->ctrl.chromosomeSize);

	// ---- initial population: speculate, evaluate in one call, replay until a pass has no miss --------
	{
		const LoopState before = st;
		const SortedChromosomes eliteBefore = this->elite;
		const double minEliteBefore = this->minEliteFitness;
		for (;;) {
			++this->counters.passes;
			this->provisional.clear();
			this->acceptedOrder.clear();
			this->generateInitial(st);
			if (this->pendingKeys.empty()) break;
			this->flush();
			this->elite = eliteBefore;
			this->minEliteFitness = minEliteBefore;
			if (this->confirmProvisional()) { // the pass stands: redo what depended on the fitness, in the same order
				st.minFitness = before.minFitness;
				for (size_t i = 0; i < this->acceptedOrder.size(); ++i) {
					if (this->acceptedOrder[i]->getFitness() < st.minFitness) st.minFitness = this->acceptedOrder[i]->getFitness();
					this->addChromosomeToElite(*this->acceptedOrder[i]);
				}
				break;
			}
			for (size_t i = 0; i < this->slots.size(); ++i) delete this->slots[i];
			this->slots.clear();
			st = before;
		}
	}
	this->initCurrentGeneration(st.shuffledSet, st.rng);
	st.sumFitness = this->updateCurrentGeneration(this->slots, st.minFitness, true);

	// ---- generations ------------------------------------------------------------------------------------
	for (int g = this->ctrl.numGenerations; g > 0 && !this->interrupted; --g) {
		const LoopState before = st;
		const SortedChromosomes eliteBefore = this->elite;
		const double minEliteBefore = this->minEliteFitness;
		this->snapshotSlots();
		for (;;) {
			++this->counters.passes;
			this->provisional.clear();
			this->acceptedOrder.clear();
			this->mateGeneration(st);
			if (this->pendingKeys.empty()) break;
			this->flush();
			this->elite = eliteBefore;
			this->minEliteFitness = minEliteBefore;
			if (this->confirmProvisional()) { // the pass stands: redo what depended on the fitness, in the same order
				st.minFitness = 0.0;
				for (size_t i = 0; i < this->acceptedOrder.size(); ++i) {
					if (this->acceptedOrder[i]->getFitness() < st.minFitness) st.minFitness = this->acceptedOrder[i]->getFitness();
					this->addChromosomeToElite(*this->acceptedOrder[i]);
				}
				break;
			}
			st = before;
			this->restoreSlots();
		}
		st.sumFitness = this->updateCurrentGeneration(this->slots, st.minFitness, false);
	}
}

// ---------------------------------------------------------------------------------------------------------
// numThreads > 1: the workers of MultiThreadedPopulation as virtual threads
// ---------------------------------------------------------------------------------------------------------
struct BatchedPopulation::Stream {
	RNG rng;
	ShuffledSet shuffledSet;
	uint16_t offset = 0, count = 0; // the worker's slice of the population
	Stream(const RNG& rng, arma::uword chromosomeSize) : rng(rng), shuffledSet(chromosomeSize) {}
};

// A worker's share of the initial population (src/MultiThreadedPopulation.cpp:88-123): duplicates are only
// looked for inside the worker's own slice.
void BatchedPopulation::generateInitialSlice(Stream& w) {
	uint16_t filled = 0;
	while (filled < w.count && !this->interrupted) {
		Chromosome* candidate = new Chromosome(this->ctrl, w.shuffledSet, w.rng);
		bool known = false;
		for (uint16_t i = 0; i < filled && !known; ++i) known = (*this->slots[w.offset + i] == *candidate);
		size_t pendingIndex = 0;
		const Lookup res = known ? HIT_FAILED : this->lookup(*candidate, &pendingIndex);
		if (!known && res != HIT_FAILED) {
			if (res == MISS) {
				const Provisional d = {candidate, pendingIndex, -std::numeric_limits<double>::infinity(), false};
				this->provisional.push_back(d);
			}
			this->slots[w.offset + filled] = candidate;
			++filled;
		} else {
			delete candidate;
		}
		if (this->pollInterrupt()) this->interrupted = true;
	}
}

// A worker's share of one generation (src/MultiThreadedPopulation.cpp:129-250).  Differences to the
// single-thread loop: counters are per call, the elite is not touched here, and a slot whose children keep
// being rejected is eventually accepted as it is (:199-203).
void BatchedPopulation::mateSlice(Stream& w, double sumFitness) {
	const ChVecIt sliceBegin = this->slots.begin() + w.offset;
	const ChVecRIt sliceEnd(sliceBegin + w.count);
	const uint32_t maxDiscarded = static_cast<uint32_t>(Population::MAX_DISCARDED_SOLUTIONS_RATIO) * w.count;
	ChVecIt front = sliceBegin;
	ChVecRIt back = sliceEnd;
	uint8_t tries[2] = {0, 0};
	uint32_t discarded[2] = {0, 0};
	std::pair<bool, bool> duplicated(false, false);
	while (front < back.base() && !this->interrupted) {
		const bool twoChildren = (front + 1 != back.base());
		Chromosome* parentA = this->drawChromosomeFromCurrentGeneration(w.rng(0.0, sumFitness));
		Chromosome* parentB = this->drawChromosomeFromCurrentGeneration(w.rng(0.0, sumFitness), parentA);
		parentA->mateWith(*parentB, w.rng, **front, **back);
		const double betterParent = std::max(parentA->getFitness(), parentB->getFitness());
		(*front)->mutate(w.rng);
		if (twoChildren) (*back)->mutate(w.rng);
		const double cutoff = betterParent - this->ctrl.badSolutionThreshold * std::fabs(betterParent);
		if (this->ctrl.maxDuplicateEliminationTries > 0) {
			duplicated = Population::checkDuplicated(sliceBegin, sliceEnd, front, back);
		}
		for (int which = 0; which < 2; ++which) {
			if (which == 1 && !twoChildren) break;
			const bool dup = (which == 0) ? duplicated.first : duplicated.second;
			Chromosome& child = (which == 0) ? **front : **back;
			if (dup && !(++tries[which] > this->ctrl.maxDuplicateEliminationTries)) continue;
			if (dup) child.randomlyReset(w.rng, w.shuffledSet);
			tries[which] = 0;
			size_t pendingIndex = 0;
			const Lookup res = this->lookup(child, &pendingIndex);
			if (res == HIT_FAILED) continue; // EvaluatorException: the slot is drawn again
			bool accepted = (res == MISS) || (child.getFitness() > cutoff);
			if (res == MISS) { // stands iff fitness > cutoff (a rejection would also have moved the discarded counter)
				const Provisional d = {&child, pendingIndex, cutoff, false};
				this->provisional.push_back(d);
			}
			if (!accepted && ++discarded[which] > maxDiscarded) { // "the algorithm may be stuck": keep the child
				discarded[which] = 0;
				accepted = true;
			}
			if (accepted) { if (which == 0) ++front; else ++back; }
		}
		if (this->pollInterrupt()) this->interrupted = true;
	}
}

void BatchedPopulation::runVirtualThreads() {
	const uint16_t T = this->ctrl.numThreads;
	RNG mainRng(this->seed);
	ShuffledSet mainSet(this->ctrl.chromosomeSize);
	this->slots.assign(this->ctrl.populationSize, static_cast<Chromosome*>(nullptr));
	this->initCurrentGeneration(mainSet, mainRng);

	// slices and seeds in the order the reference hands them out (src/MultiThreadedPopulation.cpp:260-262,301-323)
	const uint16_t perThread = this->ctrl.populationSize / T;
	int remaining = this->ctrl.populationSize % T;
	std::vector<Stream> workers;
	workers.reserve(T);
	uint16_t offset = 0;
	for (int i = T - 2; i >= 0; --i) {
		const uint32_t workerSeed = mainRng();
		Stream w(RNG(workerSeed), this->ctrl.chromosomeSize);
		w.count = perThread;
		if (remaining > 0) { --remaining; ++w.count; }
		w.offset = offset;
		offset += w.count;
		workers.push_back(w);
	}
	{ // the main thread mates the last slice with the stream the seeds were drawn from
		Stream w(mainRng, this->ctrl.chromosomeSize);
		w.shuffledSet = mainSet;
		w.count = perThread;
		w.offset = offset;
		workers.push_back(w);
	}
	auto minFitnessOfSlots = [&]() {
		double m = this->slots[0]->getFitness();
		for (size_t i = 1; i < this->slots.size(); ++i) m = std::min(m, this->slots[i]->getFitness());
		return m;
	};

	// ---- initial population ---------------------------------------------------------------------------------
	{
		const std::vector<Stream> before = workers;
		for (;;) {
			++this->counters.passes;
			this->provisional.clear();
			for (size_t t = 0; t < workers.size(); ++t) this->generateInitialSlice(workers[t]);
			if (this->pendingKeys.empty()) break;
			this->flush();
			if (this->confirmProvisional()) break; // nothing in a worker's loop depends on the fitness but these decisions
			for (size_t i = 0; i < this->slots.size(); ++i) { delete this->slots[i]; this->slots[i] = nullptr; }
			workers = before;
		}
	}
	if (this->interrupted) return;
	double sumFitness = this->updateCurrentGeneration(this->slots, minFitnessOfSlots(), true, true);

	// ---- generations ------------------------------------------------------------------------------------------
	for (int g = this->ctrl.numGenerations; g > 0 && !this->interrupted; --g) {
		const std::vector<Stream> before = workers;
		this->snapshotSlots();
		for (;;) {
			++this->counters.passes;
			this->provisional.clear();
			for (size_t t = 0; t < workers.size(); ++t) this->mateSlice(workers[t], sumFitness);
			if (this->pendingKeys.empty()) break;
			this->flush();
			if (this->confirmProvisional()) break;
			workers = before;
			this->restoreSlots();
		}
		sumFitness = this->updateCurrentGeneration(this->slots, minFitnessOfSlots(), false, true);
	}
}
