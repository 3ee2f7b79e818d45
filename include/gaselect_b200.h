/* gaselect_b200.h -- C ABI of the B200-native fitness engine for gaselect.
 *
 * This is the drop-in boundary for ONE path of dakep/gaselect: scoring chromosomes (variable
 * subsets) by repeated cross-validated PLS (SIMPLS) or OLS/BIC fits.  Every entry point
 * replaces a piece of the reference's C++ evaluator layer; the file:line of the reference
 * interface it stands in for is cited on each declaration (paths relative to the reference
 * repository root).  Plain pointers and sizes only; no C++ or torch types; no exception crosses
 * this boundary.  All functions return GSB_OK (0) or a nonzero gsb_error; gsb_last_error()
 * gives the message.  An engine is NOT re-entrant (like a reference Evaluator instance,
 * src/PLS.h:100-103): one engine per host thread.  An engine drives one GPU, or -- with
 * gsb_config.num_devices > 1 -- several GPUs of the node from ONE process: every batch is dealt to
 * per-device replicas (longest-processing-time-first on the kernels' cost model, equal counts), the
 * replicas run concurrently on their own CUDA contexts (one persistent host thread each), and the
 * fitness / status vectors are gathered (on the host by gsb_evaluate_*, with peer copies by the
 * resident-population calls).  That is the role the numThreads slice deal plays in the reference
 * (src/MultiThreadedPopulation.cpp:260-262,301-323, selected at src/GenAlg.cpp:173-181).
 *
 * There is no CPU fallback: every evaluate call runs hand-written sm_100a CUDA kernels and
 * fails with GSB_ERR_CUDA when no usable device is present.
 */
#ifndef GASELECT_B200_H
#define GASELECT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define GSB_ABI_VERSION 4
#define GSB_SEED_WORDS 624 /* RNG::SEED_SIZE, src/RNG.h:64 */
#define GSB_MAX_DEVICES 16

typedef struct gsb_engine gsb_engine;

/* values follow enum EvaluatorClass, src/GenAlg.h:23-28 (USER = 0 is out of scope) */
typedef enum gsb_evaluator_kind {
	GSB_EVAL_PLS_CV = 1,  /* PLSEvaluator  (R: evaluatorPLS),  src/PLSEvaluator.h:25 */
	GSB_EVAL_LM = 2,      /* LMEvaluator   (R: evaluatorLM),   src/LMEvaluator.h:18 */
	GSB_EVAL_PLS_FIT = 3  /* BICEvaluator  (R: evaluatorFit),  src/BICEvaluator.h:24 */
} gsb_evaluator_kind;

/* enum Statistic of src/BICEvaluator.h:26-31 and src/LMEvaluator.h:20-25 */
typedef enum gsb_statistic {
	GSB_STAT_BIC = 0,
	GSB_STAT_AIC = 1,
	GSB_STAT_ADJ_R2 = 2,
	GSB_STAT_R2 = 3
} gsb_statistic;

/* per-chromosome outcome; nonzero values are what the reference reports by throwing
 * Evaluator::EvaluatorException (src/Evaluator.h:21-25) */
typedef enum gsb_status {
	GSB_STATUS_OK = 0,
	GSB_STATUS_EMPTY = 1,     /* empty subset: src/PLSEvaluator.cpp:72-75, src/BICEvaluator.cpp:55-58 */
	GSB_STATUS_UNDERFLOW = 2, /* ||t|| < 1e-25: src/PLSSimpls.cpp:141-143 -> src/PLSEvaluator.cpp:337-340 */
	GSB_STATUS_SINGULAR = 3,  /* system could not be solved: src/LMEvaluator.cpp:60-61 */
	/* NOT an error: fitness is valid, but a comparison of the one-SE rule (src/PLSEvaluator.cpp:284-308: the first
	 * minimum of the fold-mean MSEP, or a fold mean against the cutoff) was a near-tie within the relative band
	 * gsb_config.tie_band.  The reference, summing in another order, may take that comparison the other way and fit a
	 * different optNComp (moves the fitness by O(1e-2)); a host that needs the reference's exact choice re-scores the
	 * flagged chromosomes there.  The wrappers return the fitness and count the flag; they do not throw. */
	GSB_STATUS_FLAGGED = 4,
	/* Transient, never returned by gsb_evaluate_* / gsb_population_download: the kernel-space kernel found the chromosome
	 * ill-conditioned (it takes |v|^2 by subtraction on K = Z Z') and left it to the Gram-space kernel; the download
	 * re-scores such chromosomes before it returns (gsb_timing.retries counts them).  Only a caller that reads the
	 * device buffers of gsb_population_device_results itself can see it. */
	GSB_STATUS_RETRY = 5
} gsb_status;

typedef enum gsb_error {
	GSB_OK = 0,
	GSB_ERR_INVALID_ARGUMENT = 1, /* std::invalid_argument of the reference constructors */
	GSB_ERR_STATE = 2,            /* call order: data / segmentation not set */
	GSB_ERR_CUDA = 3,             /* CUDA runtime failure or no device */
	GSB_ERR_MEMORY = 4            /* Gram tables do not fit in device memory */
} gsb_error;

/* Constructor arguments exactly as the reference constructors take them (before any of the
 * adjustments the constructors make themselves):
 *   PLS_CV : PLSEvaluator(pls, numReplications, maxNComp, seed, verbosity, innerSegments,
 *            outerSegments, testSetSize, sdfact, complexityPenalty)   src/PLSEvaluator.cpp:27-57
 *   PLS_FIT: BICEvaluator(pls, maxNComp, seed, verbosity, numSegments, stat, sdfact)
 *            (numSegments -> inner_segments)                          src/BICEvaluator.cpp:22-44
 *   LM     : LMEvaluator(X, y, statistic, verbosity)                  src/LMEvaluator.cpp:17-25 */
typedef struct gsb_config {
	int32_t struct_size;       /* sizeof(gsb_config), for ABI evolution */
	int32_t device;            /* CUDA device ordinal */
	int32_t evaluator;         /* gsb_evaluator_kind */
	int32_t statistic;         /* gsb_statistic (LM, PLS_FIT) */
	int32_t num_replications;  /* PLS_CV */
	int32_t inner_segments;    /* PLS_CV innerSegments / PLS_FIT numSegments */
	int32_t outer_segments;    /* PLS_CV; <= 1 selects srCV */
	int32_t max_ncomp;         /* 0 = as many as possible */
	double test_set_size;      /* PLS_CV srCV only */
	double sdfact;
	double complexity_penalty; /* PLS_CV */
	/* devices of a multi-device engine (ABI 3): num_devices <= 1 uses `device` alone; otherwise the batch of every
	 * evaluate call is dealt to devices[0 .. num_devices) (distinct CUDA ordinals) and `device` is ignored */
	int32_t num_devices;
	int32_t devices[GSB_MAX_DEVICES];
	int32_t reserved;
	/* near-tie band of GSB_STATUS_FLAGGED, relative (ABI 4): 0 selects the default 1e-9, negative switches the flag off */
	double tie_band;
} gsb_config;

/* read-only facts about a prepared engine (sizes of the fit plan and the device tables) */
typedef struct gsb_info {
	int32_t n, p;
	int32_t max_ncomp;         /* after the reference's clamps, src/PLSEvaluator.cpp:135-138 */
	int32_t inner_segments;    /* effective I, src/PLSEvaluator.cpp:34 */
	int32_t outer_segments;
	int32_t num_replications;
	int32_t segmentation_vectors; /* 2*R*O*(I+1) for PLS_CV, src/PLSEvaluator.cpp:112 */
	int32_t fits_per_evaluation;  /* F = R*O*(I+1): what the reference runs per evaluate() */
	int32_t distinct_fits;        /* after de-duplicating identical training sets */
	int32_t distinct_blocks;      /* distinct held-out row blocks */
	int32_t prediction_tasks;     /* (fit, block) pairs */
	int32_t sm_count;
	int64_t gram_bytes;           /* per-split centred Gram tables resident in HBM */
	int64_t block_bytes;          /* held-out row blocks resident in HBM */
	double sdfact_effective;      /* sdfact / sqrt(I), src/PLSEvaluator.cpp:35 */
	double setup_ms;              /* one-off Gram build (K0), device time */
	double gram_dmma_ms;          /* DMMA contraction part of setup_ms */
	double gram_flops;            /* algorithmic FLOPs of the DMMA contraction */
} gsb_info;

/* device-time breakdown of the most recent evaluate call (CUDA events on the engine's stream) */
typedef struct gsb_timing {
	double h2d_ms;
	double fit_kernel_ms;     /* K1: gather + SIMPLS + held-out prediction */
	double reduce_kernel_ms;  /* K2: one-SE rule + pooled SD / BIC epilogue */
	double d2h_ms;
	double total_ms;
	int64_t fit_ctas;         /* CTAs launched by K1 */
	int32_t kernel_launches;  /* kernels launched by the call */
	int32_t retries;          /* chromosomes the download re-scored on the Gram-space kernel (GSB_STATUS_RETRY) */
	double fit_tensor_gflop;  /* FP64 tensor-core work of K1 (kernel-space kernel): DMMA.8x8x4 issued x 512 FLOP, from
	                             the launch geometry (passes x products x tiles, plus the K = Z Z' builds) */
} gsb_timing;

/* ---- lifetime ----------------------------------------------------------------------------- */

/* Validates the configuration as the reference constructors do (std::invalid_argument ->
 * GSB_ERR_INVALID_ARGUMENT).  Touches no CUDA state.  src/PLSEvaluator.cpp:27-57,
 * src/BICEvaluator.cpp:22-44, src/LMEvaluator.cpp:17-25. */
int gsb_create(gsb_engine** out, const gsb_config* cfg);

/* ~Evaluator(): releases host and device memory. src/Evaluator.h:28 */
void gsb_destroy(gsb_engine* e);

/* Message of the last failing call on this engine (or of gsb_create when e == NULL). */
int gsb_last_error(const gsb_engine* e, char* buf, size_t buflen);

int gsb_abi_version(void);

/* Number of usable CUDA devices (0 and GSB_ERR_CUDA when there is none): what a host needs to build the device list
 * of a multi-device engine without linking the CUDA runtime itself. */
int gsb_device_count(int32_t* count);

/* ---- data and CV segmentation (host side, bit-exact) -------------------------------------- */

/* X: n x p column-major (R / Armadillo layout), y: n.  The engine keeps its own copy, like
 * PLS::PLS copies X and Y (src/PLS.h:25,95-96; src/LMEvaluator.cpp:17). */
int gsb_set_data(gsb_engine* e, const double* X_colmajor, const double* y, int32_t n, int32_t p);

/* The two-stage seeding of src/GenAlg.cpp:99-103: 624 draws of RNG(uint32 seed). */
int gsb_seed_vector(uint32_t seed, uint32_t* out624);

/* Builds the CV row partitions on the host exactly as PLSEvaluator::initSegmentation
 * (src/PLSEvaluator.cpp:96-223) / BICEvaluator::initSegmentation (src/BICEvaluator.cpp:96-139)
 * do, from the 624-word seed (RNG(seed vector), src/RNG.cpp:61-73; ShuffledSet::shuffleAll,
 * src/ShuffledSet.cpp:38-44).  Also applies the maxNComp clamps.  No-op for GSB_EVAL_LM. */
int gsb_init_segmentation(gsb_engine* e, const uint32_t* seed624);

/* Alternative: adopt an externally built segmentation in the reference's order
 * [train, test, train, test, ...] (what Evaluator::getSegmentation returns,
 * src/PLSEvaluator.h:42-44).  rows: concatenated ascending 0-based row indices;
 * offsets: nvec + 1 entries. */
int gsb_set_segmentation(gsb_engine* e, const uint32_t* rows, const uint32_t* offsets, int32_t nvec);

/* Evaluator::getSegmentation(), src/Evaluator.h:35-37.  Call with rows == NULL to size:
 * *nvec and *total are always written. */
int gsb_get_segmentation(const gsb_engine* e, uint32_t* rows, uint32_t* offsets, int32_t* nvec, int64_t* total);

/* Uploads X, y, builds the fit plan and runs the one-off Gram build (K0) on the device.
 * Called implicitly by the first evaluate call. */
int gsb_prepare(gsb_engine* e);

int gsb_get_info(const gsb_engine* e, gsb_info* info);

/* Run all of this engine's device work on the caller's stream (a cudaStream_t, e.g. the host
 * framework's current stream) instead of the engine's own; NULL restores the engine's stream. */
int gsb_set_stream(gsb_engine* e, void* cuda_stream);

/* ---- the hot path -------------------------------------------------------------------------- */

/* Evaluator::evaluate(arma::uvec&) over a batch (the loop of src/GenAlg.cpp:312-325):
 * CSR lists of ascending 0-based column indices, HOST buffers.  fitness[npop], status[npop]
 * (gsb_status) are caller-allocated.  fitness is meaningful where status == 0 or
 * status == GSB_STATUS_FLAGGED. */
int gsb_evaluate_indices(gsb_engine* e, const uint32_t* idx, const uint32_t* offsets, int32_t npop,
                         double* fitness, int32_t* status);

/* Evaluator::evaluate(Chromosome&) over a batch: bitsets in the reference layout
 * (src/Chromosome.cpp:55-77,419-438): words_per_chrom = ceil(p/64) 64-bit parts per chromosome,
 * variable v <-> bit (unusedBits + v) counted LSB-first from part 0,
 * unusedBits = 64*words_per_chrom - p. */
int gsb_evaluate_bitsets(gsb_engine* e, const uint64_t* bitsets, int32_t words_per_chrom, int32_t npop,
                         double* fitness, int32_t* status);

/* The same evaluation split into its three stages, for callers that keep a population resident
 * in device memory (bench "value" leg, the multi-GPU host): upload copies the CSR lists from
 * HOST buffers to the device and sorts the chromosomes into subset-size buckets; evaluate runs
 * the kernels on the resident population (synchronous; device time in gsb_last_timing);
 * download copies fitness[npop] / status[npop] back.  gsb_evaluate_indices == the three in a row. */
int gsb_population_upload(gsb_engine* e, const uint32_t* idx, const uint32_t* offsets, int32_t npop);
int gsb_population_evaluate(gsb_engine* e);
/* gsb_population_evaluate without the final synchronisation and without the timing record: the
 * kernels are only enqueued on the engine's stream (see gsb_set_stream). */
int gsb_population_enqueue(gsb_engine* e);
int gsb_population_download(gsb_engine* e, double* fitness, int32_t* status);

/* Device addresses of the resident results (fitness: npop doubles, status: npop int32) so that a
 * multi-GPU host can hand them to NCCL without a host round trip.  Valid until the next upload.
 * Single-device engines only (a multi-device engine gathers to the host itself). */
int gsb_population_device_results(gsb_engine* e, void** d_fitness, void** d_status, int32_t* npop);

int gsb_last_timing(const gsb_engine* e, gsb_timing* t);

/* The deal of a multi-device engine, as a pure host function (testable without a GPU): chromosome c of a batch with
 * the given CSR offsets goes to device_of[c] in [0, ndev).  Longest-processing-time-first on the cost model of the
 * kernels that will run (kspace != 0: the kernel-space kernel takes subsets of more than 40 columns, its cost flat in
 * the subset size; else k^2), at most ceil(npop / ndev) chromosomes per device, ties to the lower device:
 * deterministic.  The role of the contiguous slices of src/MultiThreadedPopulation.cpp:260-262,302-308. */
int gsb_deal(const uint32_t* offsets, int32_t npop, int32_t ndev, int32_t kspace, int32_t* device_of);

/* Profiling aid: with GSB_TRACE set in the environment when the engine first touches the device, one
 * CTA of every K1 launch stamps clock64() at its phase boundaries; this returns the (up to 64)
 * stamps of the most recent launch.  All zeros when tracing is off. */
int gsb_debug_trace(gsb_engine* e, int64_t* out, int32_t n);

/* The hidden C_simpls entry (src/GenAlg.cpp:351-371): SIMPLS coefficients of ONE fit on the
 * device.  cols: k ascending column indices; rows: training rows (NULL/0 = all rows).
 * coef: k x ncomp column-major (cumulative columns), intercepts: ncomp.  Needs gsb_set_data only. */
int gsb_simpls(gsb_engine* e, const uint32_t* cols, int32_t k, const uint32_t* rows, int32_t nrows, int32_t ncomp,
               double* coef, double* intercepts);

/* ---- host RNG known-answer hooks (bit-exactness of the host side is testable without a GPU) */
/* n draws of RNG(seed624) (src/RNG.cpp:61-73,92-167) */
int gsb_rng_stream(const uint32_t* seed624, int32_t n, uint32_t* out);
/* ncalls consecutive ShuffledSet(size).shuffleAll(rng) permutations on one persistent set */
int gsb_shuffle_all(const uint32_t* seed624, int32_t size, int32_t ncalls, uint32_t* out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif

#endif /* GASELECT_B200_H */
