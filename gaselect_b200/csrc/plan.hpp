// plan.hpp -- host-side CV segmentation (bit-exact with the reference) and the fit plan.
//
// The reference builds its row partitions once per run in the evaluator constructor and then
// walks them with an iterator on every evaluate() (src/PLSEvaluator.cpp:56,228-343).  Here the
// same partitions are built on the host (buildPlsSegmentation / buildFitSegmentation) and
// compiled into a FitPlan: the distinct training sets ("fits"), the distinct held-out row
// blocks, and the (fit, block) prediction tasks with the slot of the MSEP table or pooled
// residual statistic each one feeds.
#ifndef GASELECT_B200_PLAN_HPP
#define GASELECT_B200_PLAN_HPP

#include <cstdint>
#include <string>
#include <vector>

namespace gsb {

typedef std::vector<uint32_t> RowSet; // ascending 0-based row indices

struct CvShape {
	int replications = 1; // R
	int outer = 1;        // O (>= 1)
	int inner = 1;        // I, effective (src/PLSEvaluator.cpp:34) or numSegments (BIC)
	int maxNComp = 0;     // after the reference's clamps
	double sdfact = 1.0;  // after division by sqrt(I) (src/PLSEvaluator.cpp:35, src/BICEvaluator.cpp:26)
};

// PLSEvaluator ctor + initSegmentation, src/PLSEvaluator.cpp:27-57,96-223.
// Returns "" or the std::invalid_argument message of the reference.
std::string buildPlsSegmentation(int n, int numReplications, int maxNComp, int innerSegments, int outerSegments,
                                 double testSetSize, double sdfact, const uint32_t* seed624,
                                 std::vector<RowSet>& segmentation, CvShape& shape);

// BICEvaluator ctor + initSegmentation, src/BICEvaluator.cpp:22-44,96-139.
std::string buildFitSegmentation(int n, int maxNComp, int numSegments, double sdfact, const uint32_t* seed624,
                                 std::vector<RowSet>& segmentation, CvShape& shape);

enum TaskKind { TASK_INNER = 0, TASK_OUTER = 1 };

struct PlanTask {
	int fit;   // index into FitPlan::fits
	int block; // index into FitPlan::blocks
	int kind;  // TaskKind
	int dest;  // INNER: group * inner + fold ; OUTER: group
};

struct PlanFit {
	RowSet rows;                 // training rows
	std::vector<int> exclBlocks; // blocks whose disjoint union is the complement of `rows`
	int taskBegin = 0, taskCount = 0;
};

struct PlanBlock {
	RowSet rows;
	bool allRows = false; // the block is 0..n-1 (BIC final fit): aliases the full matrix
};

struct FitPlan {
	int n = 0;
	int groups = 0;        // R*O (PLS_CV), 1 (PLS_FIT), 0 (LM)
	int innerPerGroup = 0; // I
	int groupsPerReplication = 0; // O
	std::vector<PlanFit> fits;
	std::vector<PlanBlock> blocks;
	std::vector<PlanTask> tasks;    // sorted by fit
	std::vector<int> outerRows;     // per group: number of pooled held-out rows
	int fitsPerEvaluation = 0;      // what the reference runs: groups * (I + 1)
};

// PLS_CV / PLS_FIT: segmentation in the reference's order; finalAllRows adds the all-rows refit
// of src/BICEvaluator.cpp:223-226.  Returns "" or an error message.
std::string buildCvPlan(int n, const std::vector<RowSet>& segmentation, const CvShape& shape, bool finalAllRows,
                        FitPlan& plan);

// LM: one all-rows "fit" (its Gram is what the OLS solve gathers from), no tasks.
void buildAllRowsPlan(int n, FitPlan& plan);

} // namespace gsb

#endif
