// plan.cpp -- see plan.hpp.
#include "plan.hpp"

#include <algorithm>
#include <cmath>
#include <map>

#include "host_rng.hpp"

namespace gsb {

namespace {

RowSet sortedSlice(const std::vector<uint32_t>& v, size_t from, size_t to) {
	RowSet out(v.begin() + static_cast<std::ptrdiff_t>(from), v.begin() + static_cast<std::ptrdiff_t>(to));
	std::sort(out.begin(), out.end());
	return out;
}

} // namespace

std::string buildPlsSegmentation(int n, int numReplications, int maxNComp, int innerSegments, int outerSegments,
                                 double testSetSize, double sdfact, const uint32_t* seed624,
                                 std::vector<RowSet>& segmentation, CvShape& shape) {
	segmentation.clear();
	if (n < 4 || numReplications < 1) return "PLS evaluator needs at least 4 observations and 1 replication";
	const int outer = std::max(1, outerSegments);                                  // src/PLSEvaluator.cpp:33
	const int inner = (outer <= 1 && testSetSize == 0.0) ? innerSegments - 1 : innerSegments; // :34
	if (inner < 1) return "The number of inner segments is too small";
	if (outer > 1) testSetSize = 1.0 / static_cast<double>(outer);                 // :46-48
	if (testSetSize < 0.0 || testSetSize >= 1.0) {                                 // :50-52
		return "The test set size must be within the interval (0, 1)";
	}
	if (testSetSize == 0.0 && outer == 1) testSetSize = 1.0 / (inner + 1.0);       // :101-103

	const uint64_t N = static_cast<uint64_t>(n);
	const uint64_t outerLen = static_cast<uint64_t>(static_cast<double>(N) * testSetSize); // :109 (double product, truncated)
	if (outerLen == 0 || outerLen >= N) return "The test set would be empty or the whole data set";
	const uint64_t outerRem0 = N % outerLen;                                       // :110
	const uint64_t innerLen = (N - outerLen) / static_cast<uint64_t>(inner);       // :121
	const uint64_t innerRem0 = (N - outerLen) % static_cast<uint64_t>(inner);      // :122
	if (innerLen == 0) return "Too many inner segments for the number of observations";
	const uint64_t innerLenBigOuter = innerLen - ((outerRem0 > 0 && innerRem0 == 0) ? 1 : 0); // :123-127
	if (N < outerLen + innerLen + 3) return "Too few observations for this segmentation";

	shape.replications = numReplications;
	shape.outer = outer;
	shape.inner = inner;
	shape.sdfact = sdfact / std::sqrt(static_cast<double>(inner));                 // :35
	const uint64_t minFit = N - outerLen - innerLen - 2;                           // :135
	shape.maxNComp = (maxNComp == 0 || minFit <= static_cast<uint64_t>(maxNComp)) ? static_cast<int>(minFit) : maxNComp; // :136-138

	Well19937a rng(seed624);                                                       // :97
	RowShuffler rows(static_cast<uint32_t>(n));                                    // :98
	segmentation.reserve(static_cast<size_t>(2) * numReplications * (inner + 1) * outer); // :112

	for (int rep = 0; rep < numReplications; ++rep) {
		std::vector<uint32_t> sh = rows.shuffleAll(rng);                            // :145 (copy of the persistent state)
		uint64_t outerRem = outerRem0;
		uint64_t innerRem = (outerRem > 0) ? ((innerRem0 == 0) ? static_cast<uint64_t>(inner) - 1 : innerRem0 - 1) : innerRem0; // :147-156
		for (int o = 0; o < outer; ++o) {
			uint64_t olen, ilenFixed = innerLen;
			if (outerRem > 0) {                                                     // :161-164
				olen = outerLen + 1;
				ilenFixed = innerLenBigOuter;
				--outerRem;
			} else {                                                                // :165-168
				olen = outerLen;
				innerRem = innerRem0;
			}
			uint64_t pos = 0;
			for (int j = 0; j < inner; ++j) {
				const uint64_t ilen = ilenFixed + ((static_cast<uint64_t>(j) < innerRem) ? 1 : 0); // :174-178
				RowSet train;                                                        // sh[0,pos) + sh[pos+ilen, n-olen)  :183-193
				train.reserve(static_cast<size_t>(N - olen - ilen));
				train.insert(train.end(), sh.begin(), sh.begin() + static_cast<std::ptrdiff_t>(pos));
				train.insert(train.end(), sh.begin() + static_cast<std::ptrdiff_t>(pos + ilen), sh.begin() + static_cast<std::ptrdiff_t>(N - olen));
				std::sort(train.begin(), train.end());
				segmentation.push_back(std::move(train));                            // :196
				segmentation.push_back(sortedSlice(sh, pos, pos + ilen));            // :199
				pos += ilen;
			}
			segmentation.push_back(sortedSlice(sh, 0, pos));                        // :210 outer training set
			segmentation.push_back(sortedSlice(sh, pos, N));                        // :213 outer test set
			std::rotate(sh.begin(), sh.begin() + static_cast<std::ptrdiff_t>(pos), sh.end()); // :220 last block to the front
		}
	}
	return "";
}

std::string buildFitSegmentation(int n, int maxNComp, int numSegments, double sdfact, const uint32_t* seed624,
                                 std::vector<RowSet>& segmentation, CvShape& shape) {
	segmentation.clear();
	if (numSegments < 2) return "For CV at least 2 segments are needed";            // src/BICEvaluator.cpp:31-33
	if (n < numSegments || n < 4) return "Too few observations for this segmentation";
	shape.replications = 1;
	shape.outer = 1;
	shape.inner = numSegments;
	shape.sdfact = sdfact / std::sqrt(static_cast<double>(numSegments));           // :26
	if (maxNComp <= 1) maxNComp = n - 1;                                           // :39-41
	const int len = n / numSegments, rem = n % numSegments;                        // :99-100
	if (n - len - 2 < maxNComp) maxNComp = n - len - 2;                            // :106-108
	if (maxNComp < 1) return "Too few observations for this segmentation";
	shape.maxNComp = maxNComp;

	Well19937a rng(seed624);
	RowShuffler rows(static_cast<uint32_t>(n));
	const std::vector<uint32_t> sh = rows.shuffleAll(rng);                          // :98
	size_t pos = 0;
	for (int i = 0; i < numSegments; ++i) {
		const size_t segLen = static_cast<size_t>(len) + ((i < rem) ? 1 : 0);      // :113-116
		RowSet train;                                                              // :118-128
		train.reserve(static_cast<size_t>(n) - segLen);
		train.insert(train.end(), sh.begin(), sh.begin() + static_cast<std::ptrdiff_t>(pos));
		train.insert(train.end(), sh.begin() + static_cast<std::ptrdiff_t>(pos + segLen), sh.end());
		std::sort(train.begin(), train.end());
		segmentation.push_back(std::move(train));
		segmentation.push_back(sortedSlice(sh, pos, pos + segLen));                 // :134
		pos += segLen;
	}
	return "";
}

namespace {

struct Interner {
	std::map<RowSet, int> ids;
	template <class Vec, class Make>
	int intern(const RowSet& rows, Vec& store, Make make) {
		std::map<RowSet, int>::iterator it = ids.find(rows);
		if (it != ids.end()) return it->second;
		const int id = static_cast<int>(store.size());
		store.push_back(make(rows));
		ids.insert(std::make_pair(rows, id));
		return id;
	}
};

bool isAllRows(const RowSet& rows, int n) {
	if (static_cast<int>(rows.size()) != n) return false;
	for (int i = 0; i < n; ++i) if (rows[static_cast<size_t>(i)] != static_cast<uint32_t>(i)) return false;
	return true;
}

PlanBlock makeBlock(const RowSet& rows, int n) {
	PlanBlock b;
	b.rows = rows;
	b.allRows = isAllRows(rows, n);
	return b;
}

} // namespace

std::string buildCvPlan(int n, const std::vector<RowSet>& segmentation, const CvShape& shape, bool finalAllRows,
                        FitPlan& plan) {
	plan = FitPlan();
	plan.n = n;
	const int groups = shape.replications * shape.outer;
	const int I = shape.inner;
	const size_t perGroup = finalAllRows ? static_cast<size_t>(2 * I) : static_cast<size_t>(2 * (I + 1));
	if (segmentation.size() != perGroup * static_cast<size_t>(groups)) {
		return "segmentation has the wrong number of index vectors for this evaluator";
	}
	for (size_t v = 0; v < segmentation.size(); ++v) {
		const RowSet& r = segmentation[v];
		if (r.empty()) return "segmentation contains an empty index vector";
		for (size_t i = 0; i < r.size(); ++i) {
			if (r[i] >= static_cast<uint32_t>(n)) return "segmentation row index out of range";
			if (i > 0 && r[i] <= r[i - 1]) return "segmentation index vectors must be strictly ascending";
		}
	}
	plan.groups = groups;
	plan.innerPerGroup = I;
	plan.groupsPerReplication = shape.outer;
	plan.fitsPerEvaluation = groups * (I + 1);
	plan.outerRows.assign(static_cast<size_t>(groups), 0);

	Interner fitIds, blockIds;
	std::vector<PlanTask> tasks;
	std::vector<char> mark(static_cast<size_t>(n));

	auto internBlock = [&](const RowSet& rows) {
		return blockIds.intern(rows, plan.blocks, [&](const RowSet& r) { return makeBlock(r, n); });
	};
	// candidates: blocks that are expected to tile the complement of `train`; verified, with the
	// complement itself as the fallback block
	auto internFit = [&](const RowSet& train, const std::vector<int>& candidates) {
		std::map<RowSet, int>::iterator it = fitIds.ids.find(train);
		if (it != fitIds.ids.end()) return it->second;
		std::fill(mark.begin(), mark.end(), 0);
		for (size_t i = 0; i < train.size(); ++i) mark[train[i]] = 1;
		size_t covered = train.size();
		bool ok = true;
		for (size_t c = 0; c < candidates.size() && ok; ++c) {
			const RowSet& br = plan.blocks[static_cast<size_t>(candidates[c])].rows;
			for (size_t i = 0; i < br.size(); ++i) {
				if (mark[br[i]]) { ok = false; break; }
				mark[br[i]] = 1;
			}
			covered += br.size();
		}
		PlanFit f;
		f.rows = train;
		if (ok && covered == static_cast<size_t>(n)) {
			f.exclBlocks = candidates;
		} else if (train.size() < static_cast<size_t>(n)) {
			RowSet comp;
			std::fill(mark.begin(), mark.end(), 0);
			for (size_t i = 0; i < train.size(); ++i) mark[train[i]] = 1;
			for (int i = 0; i < n; ++i) if (!mark[static_cast<size_t>(i)]) comp.push_back(static_cast<uint32_t>(i));
			f.exclBlocks.push_back(internBlock(comp));
		}
		const int id = static_cast<int>(plan.fits.size());
		plan.fits.push_back(f);
		fitIds.ids.insert(std::make_pair(train, id));
		return id;
	};

	size_t pos = 0;
	for (int g = 0; g < groups; ++g) {
		int outerBlock = -1;
		if (!finalAllRows) {
			outerBlock = internBlock(segmentation[pos + static_cast<size_t>(2 * I) + 1]);
		}
		for (int j = 0; j < I; ++j) {
			const RowSet& train = segmentation[pos + static_cast<size_t>(2 * j)];
			const RowSet& test = segmentation[pos + static_cast<size_t>(2 * j) + 1];
			const int b = internBlock(test);
			std::vector<int> cand;
			cand.push_back(b);
			if (outerBlock >= 0) cand.push_back(outerBlock);
			const int f = internFit(train, cand);
			PlanTask t = {f, b, TASK_INNER, g * I + j};
			tasks.push_back(t);
		}
		if (!finalAllRows) {
			const RowSet& train = segmentation[pos + static_cast<size_t>(2 * I)];
			std::vector<int> cand(1, outerBlock);
			const int f = internFit(train, cand);
			PlanTask t = {f, outerBlock, TASK_OUTER, g};
			tasks.push_back(t);
			plan.outerRows[static_cast<size_t>(g)] = static_cast<int>(plan.blocks[static_cast<size_t>(outerBlock)].rows.size());
		}
		pos += perGroup;
	}
	if (finalAllRows) {
		RowSet all(static_cast<size_t>(n));
		for (int i = 0; i < n; ++i) all[static_cast<size_t>(i)] = static_cast<uint32_t>(i);
		const int b = internBlock(all);
		const int f = internFit(all, std::vector<int>());
		PlanTask t = {f, b, TASK_OUTER, 0};
		tasks.push_back(t);
		plan.outerRows[0] = n;
	}

	std::stable_sort(tasks.begin(), tasks.end(), [](const PlanTask& a, const PlanTask& b) { return a.fit < b.fit; });
	plan.tasks = tasks;
	size_t t = 0;
	for (size_t f = 0; f < plan.fits.size(); ++f) {
		plan.fits[f].taskBegin = static_cast<int>(t);
		while (t < tasks.size() && tasks[t].fit == static_cast<int>(f)) ++t;
		plan.fits[f].taskCount = static_cast<int>(t) - plan.fits[f].taskBegin;
	}
	return "";
}

void buildAllRowsPlan(int n, FitPlan& plan) {
	plan = FitPlan();
	plan.n = n;
	PlanFit f;
	f.rows.resize(static_cast<size_t>(n));
	for (int i = 0; i < n; ++i) f.rows[static_cast<size_t>(i)] = static_cast<uint32_t>(i);
	plan.fits.push_back(f);
	plan.fitsPerEvaluation = 1;
}

} // namespace gsb
