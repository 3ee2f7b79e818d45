// kernels.cuh -- device-side tables shared by the CUDA kernels and the engine.
//
// HBM layout (all FP64 unless noted; "P2" = p + 2 augmented columns [Xc | yc | 1]):
//   Z        ldz x P2   column-major, globally centred X and y, ones column; rows padded to a
//                       multiple of 4 with zeros (they do not change any Gram)
//   blocks   per distinct held-out block b: ld_b x P2 column-major copy of Z[rows_b, :] so that a
//                       selected column of a block is one contiguous, coalesced segment
//   unitTab  per packed entry (i, j <= i) of the p x p lower triangle: the moment sum_rows z_i z_j of every
//                       "unit" (unit 0 = all rows, unit u >= 1 = a held-out block some training set excludes), the
//                       units of one entry contiguous: [tri(i) + j][upad].  One run of upad doubles is everything
//                       any training set's Gram needs of that entry: no 8-byte gathers from per-fit tables
//   unitSum, unitXy     per column i: sum_rows z_i and sum_rows z_i y of every unit, [i][upad]
//   gpack    per generation, per (chromosome, distinct training set T): packed lower triangle (tri(k) doubles,
//                       padded to an even count) of the centred Gram  Xc_T' Xc_T - n_T xm xm'  of the SELECTED
//                       columns, written by gramPackKernel, read back contiguously (one TMA bulk copy) by K1
//   pvec     per (chromosome, T): s0 = Xc_T' yc_T and xm = column means over T of the selected columns
#ifndef GASELECT_B200_KERNELS_CUH
#define GASELECT_B200_KERNELS_CUH

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace gsb {

struct FitDesc {
	double ym;                  // mean of yc over the training rows
	int ntr;
	int taskBegin;
	int taskCount;
	int pad;
};

// a training set as full data minus excluded units (gramPackKernel)
struct PackFit {
	int u1, u2;   // excluded units (indices into a unitTab run), -1: none
	double nT;    // number of training rows
	double sy;    // sum of yc over the training rows (unit sums combined in the same order as every other moment)
	double rnT;   // 1 / nT: the column means are sums * rnT everywhere (vectors and Gram centring agree to the bit)
};

struct TaskDesc {
	int block;
	int kind; // 0 inner (MSEP table), 1 outer (pooled residual statistics)
	int dest;
	int pad;
};

struct BlockDesc {
	unsigned long long zOff; // offset (doubles) into the block pool (or into Z when allRows)
	int ld;
	int nrows;
};

// everything K1 needs, passed by value
struct FitParams {
	const double* gpack;               // per chromosome: nfits slabs of packedDoubles(k) (the fits' centred packed Grams)
	const double* pvec;                // per chromosome: nfits slabs of 2 x roundUp(k, 2): s0, then xm
	const unsigned long long* packOff; // per chromosome: offset (doubles) of its first slab in gpack
	const unsigned long long* vecOff;  // per chromosome: offset (doubles) of its first slab in pvec
	const double* zpool;
	const FitDesc* fits;
	const TaskDesc* tasks;
	const BlockDesc* blocks;
	const uint32_t* idx;     // CSR column lists of the population
	const uint32_t* offsets;
	const int* bucket;       // chromosome ids handled by this launch
	int bucketCount;
	int nfits;
	int p;
	int maxNComp;            // table stride (components)
	int innerDests;          // groups * I
	int outerDests;          // groups
	double* mse;             // [chrom][innerDest][maxNComp]
	double* ostat;           // [chrom][outerDest][maxNComp][2] = (sum e, sum e^2)
	double* coefOut;         // optional (gsb_simpls): [k x A] + A intercepts for fit 0 / bucket[0]
	long long* trace;        // optional (profiling): clock64() stamps of one CTA per launch
	int maxTasks;            // most prediction tasks any fit has
	int maxBlockLd;          // largest padded row count of a held-out block
	// fit-only launches (launchFitOnlyKernel): no prediction in the kernel; per (chromosome, fit) the rows
	// [coef (k x A); -1 (A); intercepts (A)] go to coefScratch for predictBlocksKernel (predict_blocks.cu)
	double* coefScratch;
	const unsigned long long* coefOff; // per chromosome: offset (doubles) of its first fit's rows
	// generic kernel, many components on wide subsets: stored loadings and coefficients in a global scratch, one region of
	// genStride doubles per CTA of a launch (launchFitKernel walks the class in chunks of chromosomes that fit genCapacity)
	double* genScratch;
	unsigned long long genCapacity, genStride;
	int* guardFail;                    // GSB_GUARD (fast-path instances): shared-memory canary, as in KParams
};

// ---- K1p (predict_blocks.cu): held-out prediction of the per-fit path, one CTA per (held-out block, chromosome) ----
constexpr int kPredMaxK = 1024;   // largest subset (column list staged in shared memory)
constexpr int kPredMaxTasks = 6;  // (fit, block) tasks per pass: 6 x 10 component counts <= 64 columns

struct PredTask {
	int fit;
	int kind; // 0 inner (MSEP table), 1 outer (pooled residual statistics)
	int dest;
	int pad;
};

struct PredictParams {
	const double* zpool;
	const BlockDesc* blocks;
	const int* activeBlocks;  // blocks that some task predicts
	int nActive;
	const int* blockTaskOff;  // per active block: its tasks in ptasks (CSR)
	const PredTask* ptasks;
	const double* coef;       // coefScratch of the fit-only kernel
	const unsigned long long* coefOff;
	const uint32_t* idx;
	const uint32_t* offsets;
	const int* bucket;
	int bucketCount;
	int p;
	int maxNComp;
	int innerDests;
	int outerDests;
	int lda;                  // leading dimension of the staged block slice: >= min(128, largest block ld), == 4 mod 16
	double* mse;
	double* ostat;
};

// The "one-SE rule" of PLSEvaluator::estSEP (src/PLSEvaluator.cpp:284-308; BICEvaluator::getRSS, src/BICEvaluator.cpp:191-221)
// for one (replication, outer fold): m[j * stride + a] = MSEP of inner fold j with a + 1 components.  ONE definition for
// K2 (cv_reduce.cu) and for the kernel-space kernel, which stops its outer fits at optNComp: both take the same decisions
// to the bit.  `tie`: a comparison the rule took was closer than tieBand (relative) -- another summation order, e.g. the
// reference's own, may take it the other way (GSB_STATUS_FLAGGED).
struct OneSe {
	int opt;   // optNComp, 1-based
	bool bad;  // a fold mean is NaN (the reference throws, :337-340)
	bool tie;
};

#ifdef __CUDACC__
__device__ inline OneSe oneSeRule(const double* m, int I, int A, int stride, double sdfact, double tieBand) {
	OneSe r;
	r.bad = false;
	r.tie = false;
	// first minimum of the fold-mean MSEP (strict '<', :284-292)
	int minIdx = 0;
	double best = 0.0;
	for (int a = 0; a < A; ++a) {
		double s = 0.0;
		for (int j = 0; j < I; ++j) s += m[j * stride + a];
		s /= static_cast<double>(I);
		if (isnan(s)) r.bad = true;
		if (a > 0 && fabs(s - best) <= tieBand * fabs(best)) r.tie = true;
		if (a == 0 || s < best) { best = s; minIdx = a; }
	}
	// sample SD over the folds at the minimum (I - 1 denominator; NaN when I == 1)
	double ss = 0.0;
	for (int j = 0; j < I; ++j) { const double d = m[j * stride + minIdx] - best; ss = fma(d, d, ss); }
	const double cutoff = best + sqrt(ss / static_cast<double>(I - 1)) * sdfact; // :296
	if (minIdx == 0) {
		r.opt = 1; // :298-299
	} else {
		int opt = 0; // :301-307
		while (opt < minIdx) {
			double s = 0.0;
			for (int j = 0; j < I; ++j) s += m[j * stride + opt];
			s /= static_cast<double>(I);
			if (fabs(s - cutoff) <= tieBand * fabs(cutoff)) r.tie = true;
			if (!(s > cutoff)) break;
			++opt;
		}
		r.opt = opt + 1;
	}
	return r;
}
#endif

struct ReduceParams {
	const double* mse;
	const double* ostat;
	const uint32_t* offsets;
	const int* outerRows; // per group
	int npop;
	int groups;
	int inner;
	int outerPerRep;
	int replications;
	int maxNComp;
	int innerDests;
	int outerDests;
	int kind;      // gsb_evaluator_kind
	int statistic;
	int n;
	double sdfact;
	double gamma;
	double r2denom;
	double tieBand; // relative band of the near-tie flag (GSB_STATUS_FLAGGED); negative: off
	const int* refit; // optional: chromosomes the kernel-space kernel wants re-scored (status kStatusRetry)
	double* fitness;
	int* status;
};

struct OlsParams {
	const double* gram;   // packed centred Gram of all rows
	const double* s0;
	const double* xm;
	double ym;
	const double* z;      // Z (ldz x P2)
	int ldz;
	int n;
	int p;
	const uint32_t* idx;
	const uint32_t* offsets;
	const int* bucket;
	int bucketCount;
	int statistic;
	double r2denom;
	double* fitness;
	int* status;
	double* scratch;                 // subsets too wide for shared memory: the packed triangle is factorised here
	unsigned long long scratchCapacity, scratchStride;
};


// ---- K1x (simpls_scores.cu): score-space SIMPLS on the FP64 tensor cores, 8 fits of one replication per CTA --
// Applies when the training sets of a replication are unions of a few disjoint row blocks ("atoms": the outer
// segments of a nested CV whose inner segments coincide with them, plus the never-held-out rest).  Rows are kept in
// atom order (one row-permuted copy of Z per replication); a fit is a bit mask of excluded atoms.
struct XRepDesc {
	unsigned long long zOff; // row-permuted copy of Z for this replication: nr x (p + 2), column-major
	int nslots;              // 32-row slots; every atom starts a new slot (padding rows are zero rows)
	int natoms;
	int slotAtom[6];         // atom of each slot
	int slotRows[6];         // live rows of each slot (<= 32)
};

struct XFitDesc {
	uint32_t exclMask; // atoms that are NOT training rows of this fit
	int taskBegin;     // into xtasks
	int ntasks;
	int pad;
};

struct XTaskDesc {
	int atom;
	int kind; // 0 inner (MSEP table), 1 outer (pooled residual statistics)
	int dest;
	int rows; // rows of the atom
};

struct XGroupDesc {
	int rep;
	int nfits;    // <= 16
	int fitBegin; // into xfits
	int pad;
};

struct XParams {
	const double* zpool;
	const XRepDesc* reps;
	const XGroupDesc* groups;
	const XFitDesc* xfits;
	const XTaskDesc* xtasks;
	const uint32_t* idx;
	const uint32_t* offsets;
	const int* bucket;
	int bucketCount;
	int ngroups;
	int nr;       // rows of the permuted copies: 32 x slots (the same for every replication)
	int nf;       // fits per group: 8 or 16 (= warps per CTA)
	long long* trace; // optional (profiling): clock64() stamps of one CTA per launch
	double* vscratch;    // stored loadings of the vGlobal configurations: one region per CTA
	long long vstride;   // doubles per region
	int vslots;          // > 0: regions are indexed by %smid (at most one CTA of the launch per SM), else by CTA
	int desync;          // 16-fit single-CTA launches: 0 CTA-wide barriers, 1 the two 8-fit tiles run on named barriers, 2/3 and start out of phase
	int p;
	int maxNComp;
	int innerDests;
	int outerDests;
	double* mse;
	double* ostat;
};

// ---- K1k (simpls_kspace.cu): SIMPLS in the space of the rows; one K = Z Z' per chromosome, any training sets --------
constexpr int kKspaceComponents = 10;
constexpr int kKspaceSlots = 5; // 32-row slots per lane: n <= 160

struct KTaskDesc {
	uint32_t rows[kKspaceSlots]; // the held-out rows this task predicts
	int kind;                    // 0 inner (MSEP table), 1 outer (pooled residual statistics); -1: none
	int dest;
	int nrows;
	double rr;                   // 1 / nrows
};

// everything a warp needs to start a fit, in one 128-byte record: no dependent loads, no divisions, no reductions
struct KFitDesc {
	uint32_t train[kKspaceSlots]; // bit r % 32 of word r / 32: row r is a training row
	int taskBegin;                // into ktasks
	int ntasks;
	int ntr;                      // number of training rows
	double ym;                    // mean of the (globally centred) response over the training rows (src/PLSSimpls.cpp:28-49)
	double syc;                   // sum over the training rows of (y - ym): the rounding residue of that centring
	double rntr;                  // 1 / ntr
	KTaskDesc first[2];           // copies of the fit's first two tasks (kind -1 when absent)
};

// the fits one CTA of a chromosome runs: fitOrder[begin .. begin + n0 + n1); the first n0 need every component (they
// feed the MSEP table), the last n1 only predict outer folds: once the first are done the CTA runs the one-SE rule
// itself and stops each of those at its optNComp (src/PLSEvaluator.cpp:316-317 fits exactly optNComp components)
struct KPartDesc {
	int begin, n0, n1, pad;
};

struct KParams {
	const double* zpool; // globally centred Z = [X y 1], column-major, leading dimension ldz, rows n..ldz-1 zero
	const KFitDesc* kfits;
	const KTaskDesc* ktasks;
	const uint32_t* idx;
	const uint32_t* offsets;
	const int* bucket;
	int bucketCount;
	int nfits;
	int fullCount; // the first fullCount chromosomes (largest first) run on one CTA each ...
	int tailSplit; // ... the others on tailSplit CTAs each (every CTA builds K and takes a share of the fits)
	int nf;     // fits per pass = warps per CTA: 8, or 16 (stored vectors in tensor memory)
	int n, n8, ldz, p;
	int maxNComp;
	int innerDests;
	int outerDests;
	double* mse;
	double* ostat;
	long long* trace; // optional (profiling): clock64() stamps of one CTA
	const int* fitOrder;      // the order the fits are run in (indices into kfits), per part, for tailSplit CTAs per chromosome
	const KPartDesc* parts;   // tailSplit entries
	const int* fitOrderFull;  // the same for one CTA per chromosome
	const KPartDesc* partsFull;
	int inner;                // I: inner folds per (replication, outer fold) of the MSEP table
	double sdfact;            // of the one-SE rule (after division by sqrt(I))
	uint32_t* sched;          // per CTA 2 x schedStride words: (components << 24 | fit) per slot of its passes, and the sort's staging
	int schedStride;
	unsigned long long* products; // += column tiles x products this launch issued (FP64 tensor work actually done)
	int* refit;                   // per chromosome: set when a fit lost orthogonality (v'v by subtraction) or seemed to underflow
	int* guardFail;               // GSB_GUARD: != nullptr: 64 bytes of canary behind the CTA's shared memory, counted here when overwritten
};

// status K2 gives a chromosome the kernel-space kernel marked in KParams::refit (GSB_STATUS_RETRY of the C ABI): the
// engine scores it again on the Gram-space kernel when the results are downloaded
constexpr int kStatusRetry = 5;


__host__ __device__ inline unsigned long long tri(unsigned long long i) { return i * (i + 1ull) / 2ull; }
// doubles of one packed k x k lower triangle in gpack (even: 16-byte aligned slabs for the bulk copies)
__host__ __device__ inline unsigned long long packedDoubles(unsigned long long k) { return (tri(k) + 1ull) & ~1ull; }

// gramPackKernel: the chromosomes of one launch
struct PackParams {
	const double* unitTab;  // [tri(p)][upad]
	const double* unitSum;  // [p][upad]
	const double* unitXy;   // [p][upad]
	const PackFit* pfits;
	const uint32_t* idx;
	const uint32_t* offsets;
	const int* bucket;
	int bucketCount;
	int nfits;
	int upad;
	double* gpack;
	double* pvec;
	const unsigned long long* packOff;
	const unsigned long long* vecOff;
};

// launchers (definitions next to the kernels)
void launchGramSyrk(const double* const* unitPtrs, const int* unitLd, const int* unitRows, double* const* outPtrs,
                    int units, int P2, int ldm, cudaStream_t stream);
void launchGramCombine(const double* mFull, const double* const* mBlocks, const int* exclIndex, const int* exclCount,
                       const int* fitIds, const int* fitNtr, int nfitsChunk, int p, int ldm, double* gramPool,
                       const unsigned long long* gramOffs, double* s0, double* xm, double* ym, cudaStream_t stream);
void launchUnitInterleave(const double* const* mUnits, int unitBegin, int unitCount, int p, int ldm, int upad, double* unitTab,
                          double* unitSum, double* unitXy, cudaStream_t stream);
cudaError_t launchGramPack(const PackParams& prm, int kMax, cudaStream_t stream);
long long gramPackSharedBytes(int upad, int nfits);
void launchBlockGather(const double* z, int ldz, int ncols, const uint32_t* cols, const uint32_t* rows, int nrows,
                       double* dst, int ld, cudaStream_t stream);
unsigned long long fitKernelScratchPerCta(int kMax, int A, int maxTasks, int maxBlockLd, int smemLimit);
int fitKernelSharedBytes(int k, int A);
int fitKernelMaxK(int A, int smemLimit);
cudaError_t launchFitKernel(const FitParams& prm, int kMax, int A, int smemLimit, cudaStream_t stream, long long* ctas);
bool fitOnlyKernelServes(int kMax, int A, int smemLimit);
cudaError_t launchFitOnlyKernel(const FitParams& prm, int kMax, int A, int smemLimit, cudaStream_t stream, long long* ctas);
int predictKernelLda(int maxBlockLd);
cudaError_t launchPredictKernel(const PredictParams& prm, cudaStream_t stream, long long* ctas);
int scoreKernelMaxK(int nr, int nf, int clusterSize, int vGlobal, int smemLimit);
long long scoreKernelScratchDoubles(int kMax, int nf, int clusterSize);
int scoreKernelScratchSlots(int kMax, int nf, int clusterSize, int nr, long long ctas);
constexpr int kScoreSmSlots = 256; // >= the largest %smid + 1
cudaError_t launchScoreKernel(const XParams& prm, int kMax, int A, int clusterSize, int vGlobal, int smemLimit, cudaStream_t stream,
                              long long* ctas);
constexpr int kScoreComponents = 10;
constexpr int kScoreMaxSlots = 6; // 32-row slots: at most 192 rows including the padding of every atom to 32
int kspaceKernelSharedBytes(int n, int nf);
int kspaceKernelFits(int n, int A, int smemLimit, int nf);
cudaError_t launchKspaceKernel(const KParams& prm, cudaStream_t stream, long long* ctas);
void launchReduceKernel(const ReduceParams& prm, cudaStream_t stream);
cudaError_t launchOlsKernel(const OlsParams& prm, int kMax, int smemLimit, cudaStream_t stream);
int olsKernelMaxK(int smemLimit);
unsigned long long olsKernelScratchPerCta(int kMax, int smemLimit);
void launchMarkEmpty(const uint32_t* offsets, int npop, double* fitness, int* status, cudaStream_t stream);
void launchLmEmpty(const uint32_t* offsets, int npop, int statistic, int n, double r2denom, double* fitness, int* status, cudaStream_t stream);

} // namespace gsb

#endif
