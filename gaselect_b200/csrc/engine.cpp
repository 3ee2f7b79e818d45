// engine.cpp -- the extern "C" layer of include/gaselect_b200.h: host-side state, the one-off
// device setup (K0) and the per-generation launch sequence (K1 -> K2, or K3 for the LM evaluator).
//
// What the reference does per chromosome inside Evaluator::evaluate (src/PLSEvaluator.cpp:71-91,
// src/BICEvaluator.cpp:54-91, src/LMEvaluator.cpp:27-67) is done here for a whole population per
// call.  There is no CPU fallback: without a CUDA device every evaluate call fails (GSB_ERR_CUDA).
#include "../../include/gaselect_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "host_rng.hpp"
#include "kernels.cuh"
#include "plan.hpp"

namespace {

std::string g_createError;

int roundUp(int v, int m) { return (v + m - 1) / m * m; }

struct CudaFailure {
	cudaError_t code;
	const char* what;
	const char* file;
	int line;
};

#define GSB_CUDA(expr)                                                     \
	do {                                                                   \
		cudaError_t err__ = (expr);                                        \
		if (err__ != cudaSuccess) throw CudaFailure{err__, #expr, __FILE__, __LINE__}; \
	} while (0)

// GSB_GUARD=1 (read when the first engine is created): every device array gets a 256-byte guard zone behind its last
// element, filled with a pattern; gsb_population_evaluate / download check all of them (and the shared-memory canaries of
// the K1 kernels) and fail with GSB_ERR_CUDA when one was overwritten.  compute-sanitizer is not available on the pool
// this was developed on: this is the engine's own evidence against out-of-bounds WRITES (tests/test_gpu_guard.py).
bool g_guard = false;
constexpr size_t kGuardBytes = 256;
constexpr unsigned char kGuardByte = 0xA5;

template <class T>
struct DeviceArray {
	T* ptr = nullptr;
	size_t capacity = 0; // elements
	bool guarded = false;
	void reserve(size_t count) {
		if (count <= capacity) return;
		release();
		const size_t want = std::max<size_t>(count, 16);
		GSB_CUDA(cudaMalloc(reinterpret_cast<void**>(&ptr), want * sizeof(T) + (g_guard ? kGuardBytes : 0)));
		capacity = want;
		guarded = g_guard;
		if (guarded) GSB_CUDA(cudaMemset(reinterpret_cast<unsigned char*>(ptr) + want * sizeof(T), kGuardByte, kGuardBytes));
	}
	// true when the guard zone still holds its pattern (or there is none)
	bool guardIntact() const {
		if (!ptr || !guarded) return true;
		unsigned char host[kGuardBytes];
		if (cudaMemcpy(host, reinterpret_cast<const unsigned char*>(ptr) + capacity * sizeof(T), kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
		for (size_t i = 0; i < kGuardBytes; ++i) if (host[i] != kGuardByte) return false;
		return true;
	}
	void release() {
		if (ptr) cudaFree(ptr);
		ptr = nullptr;
		capacity = 0;
	}
	DeviceArray() {}
	~DeviceArray() { release(); }
	DeviceArray(const DeviceArray&) = delete;
	DeviceArray& operator=(const DeviceArray&) = delete;
	void upload(const T* host, size_t count, cudaStream_t s) {
		reserve(count);
		if (count) GSB_CUDA(cudaMemcpyAsync(ptr, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
	}
	void upload(const std::vector<T>& host, cudaStream_t s) { upload(host.data(), host.size(), s); }
};

template <class T>
struct PinnedArray {
	T* ptr = nullptr;
	size_t capacity = 0;
	void reserve(size_t count) {
		if (count <= capacity) return;
		release();
		const size_t want = std::max<size_t>(count + count / 2, 64);
		GSB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ptr), want * sizeof(T), cudaHostAllocDefault));
		capacity = want;
	}
	void release() {
		if (ptr) cudaFreeHost(ptr);
		ptr = nullptr;
		capacity = 0;
	}
	PinnedArray() {}
	~PinnedArray() { release(); }
	PinnedArray(const PinnedArray&) = delete;
	PinnedArray& operator=(const PinnedArray&) = delete;
};

enum BucketKernel { BK_FIT = 0, BK_SCORES = 1, BK_KSPACE = 2 };

struct Bucket {
	int begin;
	int count;
	int kMax;
	int kernel;        // BucketKernel, decided when the population is uploaded
	bool split;        // BK_FIT: fit-only kernel + predictBlocksKernel (else the per-fit kernel predicts in place)
	unsigned long long genPerCta; // BK_FIT: doubles of global scratch per CTA (many components on wide subsets), 0: none
	int nf, cs, vg;    // K1x configuration (fits per CTA, cluster size, loadings in the global scratch)
	long long vPerCta, vOffset;
	int vSlots;
};

// subset-size classes: one K1 launch per class, shared memory sized for the class maximum
int bucketCeil(int k) {
	if (k <= 64) return roundUp(k, 8);
	if (k <= 256) return roundUp(k, 16);
	return roundUp(k, 32);
}

// profiling / experiment switches, read from the environment ONCE per engine (ensureDevice)
struct EnvFlags {
	bool serial = false, noScores = false, noKspace = false, trace = false, noPredict = false, forcePredict = false, noEarlyStop = false;
	int kspaceNf = 8, kspaceMinK = 40, kspaceSplit = 0, desync = 0;
	int forcedNf = 0, forcedCs = 0, forcedVg = 0;
};

// One persistent host thread per device of a multi-device engine: the per-device engines are driven concurrently
// (what MultiThreadedPopulation's worker threads are to the reference's cloned evaluators,
// src/MultiThreadedPopulation.cpp:255-323), without creating threads per call.
class DevicePool {
public:
	explicit DevicePool(int n) : jobs_(static_cast<size_t>(n)), rc_(static_cast<size_t>(n), 0), pending_(static_cast<size_t>(n), 0) {
		for (int i = 0; i < n; ++i) threads_.emplace_back([this, i] { loop(i); });
	}
	~DevicePool() {
		{
			std::lock_guard<std::mutex> lk(m_);
			stop_ = true;
		}
		cv_.notify_all();
		for (size_t i = 0; i < threads_.size(); ++i) threads_[i].join();
	}
	// runs job(i) on thread i for every i, waits for all; returns the first nonzero result (and its index)
	int run(const std::function<int(int)>& job, int* which) {
		{
			std::lock_guard<std::mutex> lk(m_);
			for (size_t i = 0; i < jobs_.size(); ++i) { jobs_[i] = job; pending_[i] = 1; }
			left_ = static_cast<int>(jobs_.size());
		}
		cv_.notify_all();
		std::unique_lock<std::mutex> lk(m_);
		done_.wait(lk, [this] { return left_ == 0; });
		for (size_t i = 0; i < rc_.size(); ++i) {
			if (rc_[i] != 0) { if (which) *which = static_cast<int>(i); return rc_[i]; }
		}
		return 0;
	}
private:
	void loop(int i) {
		for (;;) {
			std::function<int(int)> job;
			{
				std::unique_lock<std::mutex> lk(m_);
				cv_.wait(lk, [this, i] { return stop_ || pending_[static_cast<size_t>(i)]; });
				if (stop_) return;
				job = jobs_[static_cast<size_t>(i)];
				pending_[static_cast<size_t>(i)] = 0;
			}
			const int rc = job(i);
			{
				std::lock_guard<std::mutex> lk(m_);
				rc_[static_cast<size_t>(i)] = rc;
				--left_;
			}
			done_.notify_all();
		}
	}
	std::mutex m_;
	std::condition_variable cv_, done_;
	std::vector<std::function<int(int)>> jobs_;
	std::vector<int> rc_;
	std::vector<char> pending_;
	std::vector<std::thread> threads_;
	int left_ = 0;
	bool stop_ = false;
};

// What a multi-device engine keeps besides its per-device engines: the deal of the resident population and the
// gather buffers (device 0 of the list, then pinned host memory)
struct MultiState {
	DevicePool* pool = nullptr;
	std::vector<std::vector<int>> shares;        // per device: caller indices of its chromosomes, ascending
	std::vector<std::vector<uint32_t>> idx, off; // per device: its CSR lists
	std::vector<int> first;                      // per device: offset of its results in the gathered vectors
	int npop = 0;
	bool populationReady = false, resultsReady = false;
	double* dFitness = nullptr;                  // on devices[0]: the shares' results back to back
	int* dStatus = nullptr;
	double* hFitness = nullptr;                  // pinned
	int* hStatus = nullptr;
	size_t capacity = 0;
	cudaStream_t stream = nullptr;               // on devices[0]
	std::vector<double> gatherF;                 // gsb_evaluate_indices: the replicas' results back to back (host)
	std::vector<int32_t> gatherS;
	~MultiState() { delete pool; }
};

} // namespace

struct gsb_engine {
	gsb_config cfg;
	std::string error;
	// multi-device engine (cfg.num_devices > 1): one single-device engine per GPU; this object only deals and gathers
	std::vector<gsb_engine*> replicas;
	MultiState* multi = nullptr;

	// host copies (PLS::PLS copies X and Y, src/PLS.h:25,95-96)
	int n = 0, p = 0;
	std::vector<double> X, y;
	std::vector<double> xMean; // global column means (the device works on globally centred data)
	double yMean = 0.0;
	double r2denom = 0.0;
	bool haveData = false, haveSegmentation = false, prepared = false, deviceReady = false, zReady = false;

	std::vector<gsb::RowSet> segmentation;
	gsb::CvShape shape;
	gsb::FitPlan plan;
	EnvFlags env;

	// device
	cudaStream_t stream = nullptr;    // the stream all work is enqueued on
	cudaStream_t ownStream = nullptr; // created by the engine
	bool callerStream = false;
	// K1 launches of different subset-size classes run concurrently: CTAs of small subsets fill the
	// shared memory / issue slots that the one or two resident CTAs of a large subset leave idle
	static constexpr int kSideStreams = 8;
	cudaStream_t side[kSideStreams] = {};
	cudaEvent_t evFork = nullptr, evJoin[kSideStreams] = {};
	cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	int smCount = 0, smemLimit = 0;
	int ldz = 0;
	DeviceArray<double> zpool;
	DeviceArray<double> gram, s0, xm; // LMEvaluator only: the all-rows centred Gram (packed) and vectors
	DeviceArray<gsb::FitDesc> fits;
	DeviceArray<gsb::TaskDesc> tasks;
	DeviceArray<gsb::BlockDesc> blocks;
	std::vector<gsb::BlockDesc> blockDesc;
	DeviceArray<int> outerRows;
	double ymAllRows = 0.0;
	int maxTasks = 0, maxBlockLd = 0;

	// per-fit Gram-space kernel: the unit moment tables (built the first time a size class needs them)
	bool unitsReady = false;
	std::vector<int> unitBlocks;      // block id of unit u >= 1
	std::vector<int> unitOfBlock;     // unit of a block, -1: no training set excludes it
	int upad = 0;
	DeviceArray<double> unitTab, unitSum, unitXy;
	DeviceArray<gsb::PackFit> packFits;
	// ... and, per resident population, the packed Grams / vectors of every (chromosome, training set)
	DeviceArray<double> gpack, pvec;
	DeviceArray<unsigned long long> dPackOff, dVecOff, dCoefOff;
	PinnedArray<unsigned long long> hPackOff, hVecOff, hCoefOff;
	// ... and the fits' coefficient rows for the prediction kernel, whose task lists are per held-out block
	DeviceArray<double> coefScratch;
	DeviceArray<double> genScratch; // stored loadings + coefficients of the generic kernel when they outgrow shared memory
	DeviceArray<int> activeBlocks, blockTaskOff;
	DeviceArray<gsb::PredTask> predTasks;
	int nActiveBlocks = 0;

	// resident population
	int npop = 0, kMaxPop = 0, tableA = 1;
	bool populationReady = false, resultsReady = false;
	std::vector<Bucket> buckets;
	DeviceArray<uint32_t> dIdx, dOffsets;
	DeviceArray<int> dBucket, dStatus;
	DeviceArray<double> dMse, dOstat, dFitness;
	PinnedArray<uint32_t> hIdx, hOffsets;
	PinnedArray<int> hBucket, hStatus;
	PinnedArray<double> hFitness;
	// K1x (score-space tensor-core kernel): row-permuted copies of Z per replication and the atom tables
	// (built the first time a size class needs them)
	bool scoresTried = false;
	DeviceArray<double> xz, scoreScratch;
	DeviceArray<gsb::XRepDesc> xReps;
	DeviceArray<gsb::XGroupDesc> xGroups8, xGroups16;
	DeviceArray<gsb::XFitDesc> xFits;
	DeviceArray<gsb::XTaskDesc> xTasks;
	int scoreGroups8 = 0, scoreGroups16 = 0, scoreNr = 0;
	// K1k (kernel-space tensor-core kernel): training-row masks of the fits, held-out-row masks of the tasks
	DeviceArray<gsb::KFitDesc> kFits;
	DeviceArray<gsb::KTaskDesc> kTasks;
	int kspaceFits = 0; // 0: not available for this plan
	// ... and the order its CTAs run the fits in (buildKspaceSchedule): fits that feed the MSEP table first, outer-only fits last
	std::vector<int> kFitPhase, kFitComp; // per fit: 1 = outer-only; dependency component (fits and CV groups that share MSEP tables)
	int kComponents = 0;
	struct KSchedule { // for one number of CTAs per chromosome (index) and kSchedNf fits per pass
		bool ready = false, early = false;
		int stride = 2; // schedule words a CTA needs
		DeviceArray<int> fitOrder;
		DeviceArray<gsb::KPartDesc> parts;
	};
	static constexpr int kMaxSplit = 8;
	KSchedule kSchedules[kMaxSplit + 1];
	int kSchedNf = -1;
	DeviceArray<uint32_t> kSched;
	DeviceArray<unsigned long long> kProducts;
	bool kProductsPending = false;
	// chromosomes K1k declines (ill-conditioned: it takes |v|^2 by subtraction) are scored again on the Gram-space kernel by
	// a child engine without K1k, created the first time it is needed (resolveRetries)
	DeviceArray<int> dRefit;
	DeviceArray<int> dGuardFail;   // GSB_GUARD: shared-memory canaries the K1 kernels found overwritten
	bool refitArmed = false;       // the most recent enqueue ran K1k: its results may hold kStatusRetry
	gsb_engine* retryEngine = nullptr;
	bool forceNoKspace = false;    // this IS such a child engine
	bool h2dPending = false;       // gsb_evaluate_indices: the upload's copies and the kernels are in flight, timed at the download
	DeviceArray<long long> traceBuf; // allocated only when GSB_TRACE is set in the environment
	double lastH2dMs = 0.0;

	gsb_info info;
	gsb_timing timing;

	gsb_engine() {
		std::memset(&cfg, 0, sizeof(cfg));
		std::memset(&info, 0, sizeof(info));
		std::memset(&timing, 0, sizeof(timing));
	}
};

namespace {

int fail(gsb_engine* e, int code, const std::string& msg) {
	if (e) e->error = msg; else g_createError = msg;
	return code;
}

int failCuda(gsb_engine* e, const CudaFailure& f) {
	char buf[512];
	std::snprintf(buf, sizeof(buf), "CUDA error %d (%s) in %s at %s:%d", static_cast<int>(f.code),
	              cudaGetErrorString(f.code), f.what, f.file, f.line);
	cudaGetLastError(); // clear the sticky-less error state
	return fail(e, GSB_ERR_CUDA, buf);
}

void releaseDevice(gsb_engine* e) {
	e->zpool.release(); e->gram.release(); e->s0.release(); e->xm.release();
	e->fits.release(); e->tasks.release(); e->blocks.release(); e->outerRows.release();
	e->traceBuf.release();
	e->unitTab.release(); e->unitSum.release(); e->unitXy.release(); e->packFits.release();
	e->gpack.release(); e->pvec.release(); e->dPackOff.release(); e->dVecOff.release(); e->dCoefOff.release();
	e->hPackOff.release(); e->hVecOff.release(); e->hCoefOff.release();
	e->coefScratch.release(); e->genScratch.release(); e->activeBlocks.release(); e->blockTaskOff.release(); e->predTasks.release();
	e->xz.release(); e->scoreScratch.release(); e->xReps.release(); e->xGroups8.release(); e->xGroups16.release(); e->xFits.release(); e->xTasks.release();
	e->kFits.release(); e->kTasks.release();
	for (int i = 0; i <= gsb_engine::kMaxSplit; ++i) { e->kSchedules[i].fitOrder.release(); e->kSchedules[i].parts.release(); e->kSchedules[i].ready = false; }
	e->kSched.release(); e->kProducts.release(); e->dRefit.release(); e->dGuardFail.release();
	e->kSchedNf = -1;
	e->dIdx.release(); e->dOffsets.release(); e->dBucket.release(); e->dStatus.release();
	e->dMse.release(); e->dOstat.release(); e->dFitness.release();
	e->hIdx.release(); e->hOffsets.release(); e->hBucket.release(); e->hStatus.release(); e->hFitness.release();
	for (int i = 0; i < 6; ++i) { if (e->ev[i]) cudaEventDestroy(e->ev[i]); e->ev[i] = nullptr; }
	for (int i = 0; i < gsb_engine::kSideStreams; ++i) {
		if (e->side[i]) cudaStreamDestroy(e->side[i]);
		if (e->evJoin[i]) cudaEventDestroy(e->evJoin[i]);
		e->side[i] = nullptr;
		e->evJoin[i] = nullptr;
	}
	if (e->evFork) cudaEventDestroy(e->evFork);
	e->evFork = nullptr;
	if (e->ownStream) cudaStreamDestroy(e->ownStream);
	e->ownStream = nullptr;
	e->stream = nullptr;
	e->callerStream = false;
	e->deviceReady = e->zReady = e->prepared = e->populationReady = e->resultsReady = false;
	e->unitsReady = e->scoresTried = false;
}

// Selects the device and creates the stream; fails loudly when there is no usable GPU.
void ensureDevice(gsb_engine* e) {
	if (e->deviceReady) {
		GSB_CUDA(cudaSetDevice(e->cfg.device));
		return;
	}
	int count = 0;
	GSB_CUDA(cudaGetDeviceCount(&count));
	if (count <= 0 || e->cfg.device >= count) throw CudaFailure{cudaErrorNoDevice, "device ordinal not present", __FILE__, __LINE__};
	GSB_CUDA(cudaSetDevice(e->cfg.device));
	cudaDeviceProp prop;
	GSB_CUDA(cudaGetDeviceProperties(&prop, e->cfg.device));
	e->smCount = prop.multiProcessorCount;
	e->smemLimit = static_cast<int>(prop.sharedMemPerBlockOptin) - (g_guard ? 64 : 0); // GSB_GUARD: room for the kernels' canaries
	GSB_CUDA(cudaStreamCreateWithFlags(&e->ownStream, cudaStreamNonBlocking));
	if (!e->callerStream) e->stream = e->ownStream;
	for (int i = 0; i < 6; ++i) GSB_CUDA(cudaEventCreate(&e->ev[i]));
	for (int i = 0; i < gsb_engine::kSideStreams; ++i) {
		GSB_CUDA(cudaStreamCreateWithFlags(&e->side[i], cudaStreamNonBlocking));
		GSB_CUDA(cudaEventCreateWithFlags(&e->evJoin[i], cudaEventDisableTiming));
	}
	GSB_CUDA(cudaEventCreateWithFlags(&e->evFork, cudaEventDisableTiming));
	// profiling / experiment switches: read once, here
	EnvFlags& env = e->env;
	env = EnvFlags();
	env.serial = std::getenv("GSB_SERIAL") != nullptr;
	env.noScores = std::getenv("GSB_NO_SCORES") != nullptr;
	env.noKspace = std::getenv("GSB_NO_KSPACE") != nullptr;
	env.trace = std::getenv("GSB_TRACE") != nullptr;
	env.noEarlyStop = std::getenv("GSB_NO_EARLYSTOP") != nullptr; // outer-only fits run every component (the arrangement before ABI 4)
	env.noPredict = std::getenv("GSB_NO_PREDICT") != nullptr; // per-fit kernel predicts in place whatever the block size
	env.forcePredict = std::getenv("GSB_PREDICT") != nullptr; // fit-only kernel + predictBlocksKernel whatever the block size
	env.noKspace = env.noKspace || e->forceNoKspace;
	if (const char* v = std::getenv("GSB_KSPACE_NF")) env.kspaceNf = std::atoi(v) == 16 ? 16 : 8;
	if (const char* v = std::getenv("GSB_KSPACE_MINK")) env.kspaceMinK = std::atoi(v);
	if (const char* v = std::getenv("GSB_KSPACE_SPLIT")) env.kspaceSplit = std::atoi(v);
	if (const char* v = std::getenv("GSB_DESYNC")) env.desync = std::atoi(v);
	if (const char* v = std::getenv("GSB_SCORE_CFG")) std::sscanf(v, "%d,%d,%d", &env.forcedNf, &env.forcedCs, &env.forcedVg);
	if (env.trace) { // profiling aid: K1 stamps clock64() at its phase boundaries
		e->traceBuf.reserve(64);
		GSB_CUDA(cudaMemset(e->traceBuf.ptr, 0, 64 * sizeof(long long)));
	}
	e->deviceReady = true;
}

// Z = [X - colmeans | y - mean | 1], rows padded with zeros to a multiple of 4, column-major.
void uploadZ(gsb_engine* e, size_t poolDoubles) {
	const int n = e->n, p = e->p, P2 = p + 2;
	e->ldz = roundUp(n, 4);
	const size_t zDoubles = static_cast<size_t>(e->ldz) * P2;
	std::vector<double> Z(zDoubles, 0.0);
	e->xMean.assign(static_cast<size_t>(p), 0.0);
	for (int c = 0; c < p; ++c) {
		const double* col = e->X.data() + static_cast<size_t>(c) * n;
		long double s = 0.0L;
		for (int r = 0; r < n; ++r) s += col[r];
		const double m = static_cast<double>(s / n);
		e->xMean[static_cast<size_t>(c)] = m;
		double* dst = Z.data() + static_cast<size_t>(c) * e->ldz;
		for (int r = 0; r < n; ++r) dst[r] = col[r] - m;
	}
	long double sy = 0.0L;
	for (int r = 0; r < n; ++r) sy += e->y[static_cast<size_t>(r)];
	e->yMean = static_cast<double>(sy / n);
	double* yc = Z.data() + static_cast<size_t>(p) * e->ldz;
	double* one = Z.data() + static_cast<size_t>(p + 1) * e->ldz;
	double ss = 0.0;
	for (int r = 0; r < n; ++r) {
		yc[r] = e->y[static_cast<size_t>(r)] - e->yMean;
		one[r] = 1.0;
		ss += yc[r] * yc[r];
	}
	e->r2denom = ss; // n * var_pop(y) (src/BICEvaluator.cpp:37) == sum (y - mean)^2 (src/LMEvaluator.cpp:24)
	e->zpool.reserve(std::max(poolDoubles, zDoubles));
	GSB_CUDA(cudaMemcpyAsync(e->zpool.ptr, Z.data(), zDoubles * sizeof(double), cudaMemcpyHostToDevice, e->stream));
	GSB_CUDA(cudaStreamSynchronize(e->stream));
	e->zReady = true;
}

// One SYRK pass (K0's contraction) over the units [u0, u0 + cnt): unit 0 = all rows of Z, unit u >= 1 = the block copy of
// unitBlocks[u - 1].  Results in mPool (cnt matrices of P2 x ldm); device pointer array in dOutPtrs.
void syrkUnits(gsb_engine* e, int u0, int cnt, DeviceArray<double>& mPool, DeviceArray<double*>& dOutPtrs, std::vector<double*>& outPtrs,
               double& flops) {
	const int n = e->n, P2 = e->p + 2, ldm = roundUp(P2, 4);
	const size_t mDoubles = static_cast<size_t>(P2) * ldm;
	mPool.reserve(mDoubles * static_cast<size_t>(cnt));
	GSB_CUDA(cudaMemsetAsync(mPool.ptr, 0, mDoubles * static_cast<size_t>(cnt) * sizeof(double), e->stream));
	std::vector<const double*> unitPtrs(static_cast<size_t>(cnt));
	std::vector<int> unitLd(static_cast<size_t>(cnt)), unitRows(static_cast<size_t>(cnt));
	outPtrs.assign(static_cast<size_t>(cnt), nullptr);
	for (int i = 0; i < cnt; ++i) {
		const int u = u0 + i;
		outPtrs[static_cast<size_t>(i)] = mPool.ptr + mDoubles * static_cast<size_t>(i);
		if (u == 0) {
			unitPtrs[static_cast<size_t>(i)] = e->zpool.ptr;
			unitLd[static_cast<size_t>(i)] = e->ldz;
			unitRows[static_cast<size_t>(i)] = e->ldz;
			flops += static_cast<double>(n) * P2 * (P2 + 1.0);
		} else {
			const gsb::BlockDesc& d = e->blockDesc[static_cast<size_t>(e->unitBlocks[static_cast<size_t>(u - 1)])];
			unitPtrs[static_cast<size_t>(i)] = e->zpool.ptr + d.zOff;
			unitLd[static_cast<size_t>(i)] = d.ld;
			unitRows[static_cast<size_t>(i)] = d.ld;
			flops += static_cast<double>(d.nrows) * P2 * (P2 + 1.0);
		}
	}
	DeviceArray<const double*> dUnitPtrs;
	DeviceArray<int> dUnitLd, dUnitRows;
	dUnitPtrs.upload(unitPtrs, e->stream);
	dOutPtrs.upload(outPtrs, e->stream);
	dUnitLd.upload(unitLd, e->stream);
	dUnitRows.upload(unitRows, e->stream);
	// grid.y is limited to 65535 units per launch
	for (int c0 = 0; c0 < cnt; c0 += 32768) {
		const int c = std::min(32768, cnt - c0);
		gsb::launchGramSyrk(dUnitPtrs.ptr + c0, dUnitLd.ptr + c0, dUnitRows.ptr + c0, dOutPtrs.ptr + c0, c, P2, ldm, e->stream);
	}
	GSB_CUDA(cudaGetLastError());
	GSB_CUDA(cudaStreamSynchronize(e->stream)); // the host arrays above go out of scope
}

// LMEvaluator: the centred Gram of ALL rows over all columns (packed), s0 and the column means: what K3 gathers from
void buildAllRowsGram(gsb_engine* e) {
	const int n = e->n, p = e->p, P2 = p + 2, ldm = roundUp(P2, 4);
	const size_t triP = static_cast<size_t>(gsb::tri(static_cast<unsigned long long>(p)));
	DeviceArray<double> mPool, dYm;
	DeviceArray<double*> dOutPtrs;
	std::vector<double*> outPtrs;
	double flops = 0.0;
	GSB_CUDA(cudaEventRecord(e->ev[0], e->stream));
	syrkUnits(e, 0, 1, mPool, dOutPtrs, outPtrs, flops);
	GSB_CUDA(cudaEventRecord(e->ev[1], e->stream));
	DeviceArray<const double*> dMOfBlock;
	DeviceArray<int> dZero2, dZero1, dN1;
	DeviceArray<unsigned long long> dOff0;
	const double* noBlock = nullptr;
	const int zeros2[2] = {0, 0};
	const int zero1 = 0;
	const unsigned long long off0 = 0;
	dMOfBlock.upload(&noBlock, 1, e->stream);
	dZero2.upload(zeros2, 2, e->stream);
	dZero1.upload(&zero1, 1, e->stream);
	dN1.upload(&n, 1, e->stream);
	dOff0.upload(&off0, 1, e->stream);
	dYm.reserve(1);
	e->gram.reserve(triP);
	e->s0.reserve(static_cast<size_t>(p));
	e->xm.reserve(static_cast<size_t>(p));
	gsb::launchGramCombine(mPool.ptr, dMOfBlock.ptr, dZero2.ptr, dZero1.ptr, dZero1.ptr, dN1.ptr, 1, p, ldm, e->gram.ptr, dOff0.ptr,
	                       e->s0.ptr, e->xm.ptr, dYm.ptr, e->stream);
	GSB_CUDA(cudaGetLastError());
	GSB_CUDA(cudaEventRecord(e->ev[2], e->stream));
	GSB_CUDA(cudaMemcpyAsync(&e->ymAllRows, dYm.ptr, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
	GSB_CUDA(cudaStreamSynchronize(e->stream));
	float msSyrk = 0.f, msAll = 0.f;
	GSB_CUDA(cudaEventElapsedTime(&msSyrk, e->ev[0], e->ev[1]));
	GSB_CUDA(cudaEventElapsedTime(&msAll, e->ev[0], e->ev[2]));
	e->info.setup_ms = msAll;
	e->info.gram_dmma_ms = msSyrk;
	e->info.gram_flops = flops;
	e->info.gram_bytes = static_cast<int64_t>((triP + 2 * static_cast<size_t>(p)) * sizeof(double));
}

// K0 for the per-fit Gram-space kernel, built the first time a subset-size class needs it: the FP64 tensor-core SYRK
// M_u = Z_u'Z_u of every unit (the full data and each block some training set excludes), the units interleaved per packed
// entry (gram_build.cu).  The SYRK scratch is walked in chunks of units that fit beside the tables.
void buildUnitTables(gsb_engine* e) {
	if (e->unitsReady) return;
	const int p = e->p, P2 = p + 2, ldm = roundUp(P2, 4);
	const gsb::FitPlan& plan = e->plan;
	const int nfits = static_cast<int>(plan.fits.size());
	const int units = 1 + static_cast<int>(e->unitBlocks.size());
	const int upad = roundUp(units, 4);
	if (gsb::gramPackSharedBytes(upad, nfits) > e->smemLimit || nfits > 65535) {
		char buf[256];
		std::snprintf(buf, sizeof(buf), "the per-fit kernel's pack stage holds the %d unit blocks and %d training sets of this segmentation in shared memory "
		              "(%lld bytes, %d available): use fewer replications or segments", units - 1, nfits, gsb::gramPackSharedBytes(upad, nfits), e->smemLimit);
		throw std::string(buf);
	}
	const size_t mDoubles = static_cast<size_t>(P2) * ldm;
	const size_t triP = static_cast<size_t>(gsb::tri(static_cast<unsigned long long>(p)));
	const double tableBytes = 8.0 * (static_cast<double>(triP) + 2.0 * p) * upad;
	size_t freeB = 0, totalB = 0;
	GSB_CUDA(cudaMemGetInfo(&freeB, &totalB));
	const double room = 0.90 * static_cast<double>(freeB) - tableBytes;
	if (room < 8.0 * static_cast<double>(mDoubles)) {
		char buf[256];
		std::snprintf(buf, sizeof(buf), "the Gram tables of the per-fit kernel need %.1f GB of device memory, %.1f GB are free",
		              (tableBytes + 8.0 * static_cast<double>(mDoubles)) / 1e9, static_cast<double>(freeB) / 1e9);
		throw std::string(buf);
	}
	int chunk = static_cast<int>(std::min<double>(units, std::floor(room / (8.0 * static_cast<double>(mDoubles)))));
	chunk = std::max(1, std::min(chunk, 32768));
	e->unitTab.reserve(triP * static_cast<size_t>(upad));
	e->unitSum.reserve(static_cast<size_t>(p) * upad);
	e->unitXy.reserve(static_cast<size_t>(p) * upad);
	std::vector<double> unitSy(static_cast<size_t>(units), 0.0);
	DeviceArray<double> mPool;
	DeviceArray<double*> dOutPtrs;
	std::vector<double*> outPtrs;
	double flops = 0.0;
	float msSyrkAll = 0.f, msAll = 0.f;
	for (int u0 = 0; u0 < units; u0 += chunk) {
		const int cnt = std::min(chunk, units - u0);
		GSB_CUDA(cudaEventRecord(e->ev[0], e->stream));
		syrkUnits(e, u0, cnt, mPool, dOutPtrs, outPtrs, flops);
		GSB_CUDA(cudaEventRecord(e->ev[1], e->stream));
		gsb::launchUnitInterleave(reinterpret_cast<const double* const*>(dOutPtrs.ptr), u0, cnt, p, ldm, upad, e->unitTab.ptr,
		                          e->unitSum.ptr, e->unitXy.ptr, e->stream);
		GSB_CUDA(cudaGetLastError());
		GSB_CUDA(cudaEventRecord(e->ev[2], e->stream));
		// sum of yc over the rows of each unit: row "ones", column "y" of its moment matrix
		for (int i = 0; i < cnt; ++i) {
			GSB_CUDA(cudaMemcpyAsync(&unitSy[static_cast<size_t>(u0 + i)], outPtrs[static_cast<size_t>(i)] + static_cast<size_t>(p + 1) * ldm + p,
			                         sizeof(double), cudaMemcpyDeviceToHost, e->stream));
		}
		GSB_CUDA(cudaStreamSynchronize(e->stream));
		float a = 0.f, b = 0.f;
		GSB_CUDA(cudaEventElapsedTime(&a, e->ev[0], e->ev[1]));
		GSB_CUDA(cudaEventElapsedTime(&b, e->ev[0], e->ev[2]));
		msSyrkAll += a;
		msAll += b;
	}
	mPool.release();
	// training sets as full data minus excluded units; every moment is combined in this order (full, - first, - second)
	std::vector<gsb::PackFit> pf(static_cast<size_t>(nfits));
	for (int f = 0; f < nfits; ++f) {
		const gsb::PlanFit& fit = plan.fits[static_cast<size_t>(f)];
		gsb::PackFit& d = pf[static_cast<size_t>(f)];
		d.u1 = d.u2 = -1;
		double sy = unitSy[0];
		for (size_t x = 0; x < fit.exclBlocks.size() && x < 2; ++x) {
			const int u = e->unitOfBlock[static_cast<size_t>(fit.exclBlocks[x])];
			(x == 0 ? d.u1 : d.u2) = u;
			sy -= unitSy[static_cast<size_t>(u)];
		}
		d.nT = static_cast<double>(fit.rows.size());
		d.sy = sy;
		d.rnT = 1.0 / d.nT;
	}
	e->packFits.upload(pf, e->stream);
	GSB_CUDA(cudaStreamSynchronize(e->stream));
	e->upad = upad;
	e->unitsReady = true;
	e->info.setup_ms += msAll;
	e->info.gram_dmma_ms += msSyrkAll;
	e->info.gram_flops += flops;
	e->info.gram_bytes = static_cast<int64_t>(tableBytes);
}

// K1x tables, built the first time a subset-size class is given to the score-space kernel (since K1k: plans with
// 160 < n <= 192 rows, or GSB_NO_KSPACE=1)
void buildScoreTables(gsb_engine* e) {
	if (e->scoresTried) return;
	e->scoresTried = true;
	const int n = e->n, P2 = e->p + 2, ldz = e->ldz;
	const gsb::FitPlan& plan = e->plan;
	const int nfits = static_cast<int>(plan.fits.size());
	// ---- K1x: atoms (disjoint row blocks) per replication, fits as masks of excluded atoms
	e->scoreGroups8 = e->scoreGroups16 = 0;
	if (e->cfg.evaluator == GSB_EVAL_PLS_CV && nfits > 0 && !e->env.noScores) {
		const int I = std::max(1, plan.innerPerGroup), O = std::max(1, plan.groupsPerReplication);
		const int R = std::max(1, plan.groups / O);
		std::vector<int> fitRep(static_cast<size_t>(nfits), 1 << 30);
		for (size_t t = 0; t < plan.tasks.size(); ++t) {
			const gsb::PlanTask& pt = plan.tasks[t];
			const int g = pt.kind == gsb::TASK_OUTER ? pt.dest : pt.dest / I;
			fitRep[static_cast<size_t>(pt.fit)] = std::min(fitRep[static_cast<size_t>(pt.fit)], g / O);
		}
		std::vector<gsb::XRepDesc> reps(static_cast<size_t>(R));
		std::vector<gsb::XGroupDesc> groups8, groups16;
		std::vector<gsb::XFitDesc> xfits;
		std::vector<gsb::XTaskDesc> xtasks;
		std::vector<std::vector<uint32_t> > permRows(static_cast<size_t>(R)); // 32 x slots each, 0xffffffff = zero row
		int maxSlots = 0;
		bool ok = true;
		for (int r = 0; r < R && ok; ++r) {
			std::vector<int> fitsHere;
			for (int f = 0; f < nfits; ++f) if (fitRep[static_cast<size_t>(f)] == r) fitsHere.push_back(f);
			if (fitsHere.empty()) { ok = false; break; }
			// atoms: every block some fit of the replication excludes or predicts; they must be pairwise disjoint
			std::vector<int> atomBlocks;
			for (size_t x = 0; x < fitsHere.size(); ++x) {
				const gsb::PlanFit& pf = plan.fits[static_cast<size_t>(fitsHere[x])];
				for (size_t q = 0; q < pf.exclBlocks.size(); ++q) atomBlocks.push_back(pf.exclBlocks[q]);
				for (int t = pf.taskBegin; t < pf.taskBegin + pf.taskCount; ++t) atomBlocks.push_back(plan.tasks[static_cast<size_t>(t)].block);
			}
			std::sort(atomBlocks.begin(), atomBlocks.end());
			atomBlocks.erase(std::unique(atomBlocks.begin(), atomBlocks.end()), atomBlocks.end());
			std::vector<int> owner(static_cast<size_t>(n), -1);
			for (size_t a = 0; a < atomBlocks.size() && ok; ++a) {
				const gsb::PlanBlock& pb = plan.blocks[static_cast<size_t>(atomBlocks[a])];
				for (size_t i = 0; i < pb.rows.size(); ++i) {
					if (owner[pb.rows[i]] >= 0) { ok = false; break; }
					owner[pb.rows[i]] = static_cast<int>(a);
				}
			}
			if (!ok) break;
			int natoms = static_cast<int>(atomBlocks.size());
			bool rest = false;
			for (int i = 0; i < n; ++i) if (owner[static_cast<size_t>(i)] < 0) { owner[static_cast<size_t>(i)] = natoms; rest = true; }
			if (rest) ++natoms;
			if (natoms > 32) { ok = false; break; }
			// every atom starts a new 32-row slot
			gsb::XRepDesc& rd = reps[static_cast<size_t>(r)];
			rd.natoms = natoms;
			rd.nslots = 0;
			std::vector<int> atomRowCount(static_cast<size_t>(natoms), 0);
			std::vector<uint32_t>& perm = permRows[static_cast<size_t>(r)];
			for (int a = 0; a < natoms && ok; ++a) {
				std::vector<uint32_t> rowsA;
				for (int i = 0; i < n; ++i) if (owner[static_cast<size_t>(i)] == a) rowsA.push_back(static_cast<uint32_t>(i));
				atomRowCount[static_cast<size_t>(a)] = static_cast<int>(rowsA.size());
				for (size_t at = 0; at < rowsA.size(); at += 32) {
					if (rd.nslots >= gsb::kScoreMaxSlots) { ok = false; break; }
					const int live = static_cast<int>(std::min<size_t>(32, rowsA.size() - at));
					rd.slotAtom[rd.nslots] = a;
					rd.slotRows[rd.nslots] = live;
					++rd.nslots;
					for (int j = 0; j < 32; ++j) perm.push_back(j < live ? rowsA[at + static_cast<size_t>(j)] : 0xffffffffu);
				}
			}
			if (!ok) break;
			for (int u = rd.nslots; u < gsb::kScoreMaxSlots; ++u) { rd.slotAtom[u] = -1; rd.slotRows[u] = 0; }
			maxSlots = std::max(maxSlots, rd.nslots);
			// fits of the replication: one group of <= 16, or groups of <= 8 (balanced chunks)
			const int fitBase = static_cast<int>(xfits.size());
			for (size_t x = 0; x < fitsHere.size(); ++x) {
				const gsb::PlanFit& pf = plan.fits[static_cast<size_t>(fitsHere[x])];
				gsb::XFitDesc xf;
				xf.exclMask = 0u;
				xf.taskBegin = static_cast<int>(xtasks.size());
				xf.ntasks = pf.taskCount;
				xf.pad = 0;
				size_t exclRows = 0;
				for (size_t q = 0; q < pf.exclBlocks.size(); ++q) {
					const int a = static_cast<int>(std::lower_bound(atomBlocks.begin(), atomBlocks.end(), pf.exclBlocks[q]) - atomBlocks.begin());
					xf.exclMask |= 1u << a;
					exclRows += plan.blocks[static_cast<size_t>(pf.exclBlocks[q])].rows.size();
				}
				if (exclRows + pf.rows.size() != static_cast<size_t>(n)) ok = false;
				for (int t = pf.taskBegin; t < pf.taskBegin + pf.taskCount; ++t) {
					const gsb::PlanTask& pt = plan.tasks[static_cast<size_t>(t)];
					gsb::XTaskDesc td;
					td.atom = static_cast<int>(std::lower_bound(atomBlocks.begin(), atomBlocks.end(), pt.block) - atomBlocks.begin());
					td.kind = pt.kind;
					td.dest = pt.dest;
					td.rows = atomRowCount[static_cast<size_t>(td.atom)];
					if (!((xf.exclMask >> td.atom) & 1u)) ok = false; // a predicted block must be held out of the fit
					xtasks.push_back(td);
				}
				xfits.push_back(xf);
			}
			for (int width = 8; width <= 16; width += 8) {
				std::vector<gsb::XGroupDesc>& groups = width == 8 ? groups8 : groups16;
				const size_t count = fitsHere.size(), chunks = (count + width - 1) / width, per = (count + chunks - 1) / chunks;
				for (size_t c0 = 0; c0 < count; c0 += per) {
					gsb::XGroupDesc g;
					g.rep = r;
					g.nfits = static_cast<int>(std::min(count, c0 + per) - c0);
					g.fitBegin = fitBase + static_cast<int>(c0);
					g.pad = 0;
					groups.push_back(g);
				}
			}
		}
		if (ok) {
			const int nr = 32 * maxSlots;
			for (int r = 0; r < R; ++r) {
				permRows[static_cast<size_t>(r)].resize(static_cast<size_t>(nr), 0xffffffffu);
				reps[static_cast<size_t>(r)].zOff = static_cast<unsigned long long>(r) * nr * P2;
			}
			std::vector<uint32_t> permAll;
			for (int r = 0; r < R; ++r) permAll.insert(permAll.end(), permRows[static_cast<size_t>(r)].begin(), permRows[static_cast<size_t>(r)].end());
			e->xz.reserve(static_cast<size_t>(R) * nr * P2);
			DeviceArray<uint32_t> dPerm;
			dPerm.upload(permAll, e->stream);
			for (int r = 0; r < R; ++r) {
				gsb::launchBlockGather(e->zpool.ptr, ldz, P2, nullptr, dPerm.ptr + static_cast<size_t>(r) * nr, nr,
				                       e->xz.ptr + static_cast<size_t>(r) * nr * P2, nr, e->stream);
			}
			GSB_CUDA(cudaGetLastError());
			e->xReps.upload(reps, e->stream);
			e->xGroups8.upload(groups8, e->stream);
			e->xGroups16.upload(groups16, e->stream);
			e->xFits.upload(xfits, e->stream);
			e->xTasks.upload(xtasks, e->stream);
			GSB_CUDA(cudaStreamSynchronize(e->stream));
			e->scoreGroups8 = static_cast<int>(groups8.size());
			e->scoreGroups16 = static_cast<int>(groups16.size());
			e->scoreNr = nr;
		}
	}

}

// K0 at prepare time: the globally centred data, contiguous copies of the held-out blocks, the descriptors of the fit plan
// and the K1k row masks.  The moment tables of the per-fit kernel (buildUnitTables) and the K1x copies (buildScoreTables)
// are built when a subset-size class first needs them; the LM evaluator's all-rows Gram is built here.
void prepareDevice(gsb_engine* e) {
	ensureDevice(e);
	const int n = e->n, p = e->p, P2 = p + 2;
	const gsb::FitPlan& plan = e->plan;
	const int nfits = static_cast<int>(plan.fits.size());
	const int nblocks = static_cast<int>(plan.blocks.size());
	const int ldz = roundUp(n, 4);
	e->unitsReady = e->scoresTried = false;
	e->scoreGroups8 = e->scoreGroups16 = 0;
	std::memset(&e->info, 0, sizeof(e->info));

	// ---- held-out row blocks, each a contiguous column-major copy (coalesced prediction reads)
	std::vector<gsb::BlockDesc>& bdesc = e->blockDesc;
	bdesc.assign(static_cast<size_t>(nblocks), gsb::BlockDesc());
	std::vector<uint32_t> blockRows;
	std::vector<size_t> blockRowOff(static_cast<size_t>(nblocks), 0);
	size_t pool = static_cast<size_t>(ldz) * P2;
	for (int b = 0; b < nblocks; ++b) {
		const gsb::PlanBlock& pb = plan.blocks[static_cast<size_t>(b)];
		gsb::BlockDesc& d = bdesc[static_cast<size_t>(b)];
		d.nrows = static_cast<int>(pb.rows.size());
		if (pb.allRows) {
			d.zOff = 0;
			d.ld = ldz;
		} else {
			d.ld = roundUp(d.nrows, 4);
			d.zOff = pool;
			pool += static_cast<size_t>(d.ld) * P2;
			blockRowOff[static_cast<size_t>(b)] = blockRows.size();
			blockRows.insert(blockRows.end(), pb.rows.begin(), pb.rows.end());
		}
	}
	e->zReady = false;
	uploadZ(e, pool);
	DeviceArray<uint32_t> dRows;
	dRows.upload(blockRows, e->stream);
	for (int b = 0; b < nblocks; ++b) {
		if (plan.blocks[static_cast<size_t>(b)].allRows) continue;
		const gsb::BlockDesc& d = bdesc[static_cast<size_t>(b)];
		gsb::launchBlockGather(e->zpool.ptr, ldz, P2, nullptr, dRows.ptr + blockRowOff[static_cast<size_t>(b)], d.nrows,
		                       e->zpool.ptr + d.zOff, d.ld, e->stream);
	}
	GSB_CUDA(cudaGetLastError());
	GSB_CUDA(cudaStreamSynchronize(e->stream));

	// ---- units of the moment tables: the full data, then every block some fit excludes
	e->unitOfBlock.assign(static_cast<size_t>(nblocks), -1);
	e->unitBlocks.clear();
	for (int f = 0; f < nfits; ++f) {
		for (size_t x = 0; x < plan.fits[static_cast<size_t>(f)].exclBlocks.size(); ++x) {
			const int b = plan.fits[static_cast<size_t>(f)].exclBlocks[x];
			if (e->unitOfBlock[static_cast<size_t>(b)] < 0) {
				e->unitBlocks.push_back(b);
				e->unitOfBlock[static_cast<size_t>(b)] = static_cast<int>(e->unitBlocks.size());
			}
		}
	}

	// ---- descriptors
	std::vector<gsb::FitDesc> fdesc(static_cast<size_t>(nfits));
	for (int f = 0; f < nfits; ++f) {
		const gsb::PlanFit& pf = plan.fits[static_cast<size_t>(f)];
		gsb::FitDesc& d = fdesc[static_cast<size_t>(f)];
		// mean of the (globally centred) response over the training rows (src/PLSSimpls.cpp:28-49)
		double sy = 0.0;
		for (size_t i = 0; i < pf.rows.size(); ++i) sy += e->y[pf.rows[i]] - e->yMean;
		d.ym = sy / static_cast<double>(std::max<size_t>(1, pf.rows.size()));
		d.ntr = static_cast<int>(pf.rows.size());
		d.taskBegin = pf.taskBegin;
		d.taskCount = pf.taskCount;
		d.pad = 0;
	}
	std::vector<gsb::TaskDesc> tdesc(plan.tasks.size());
	for (size_t t = 0; t < plan.tasks.size(); ++t) {
		tdesc[t].block = plan.tasks[t].block;
		tdesc[t].kind = plan.tasks[t].kind;
		tdesc[t].dest = plan.tasks[t].dest;
		tdesc[t].pad = 0;
	}
	e->maxTasks = 0;
	e->maxBlockLd = 0;
	for (int f = 0; f < nfits; ++f) e->maxTasks = std::max(e->maxTasks, plan.fits[static_cast<size_t>(f)].taskCount);
	for (size_t t = 0; t < plan.tasks.size(); ++t) e->maxBlockLd = std::max(e->maxBlockLd, bdesc[static_cast<size_t>(plan.tasks[t].block)].ld);
	{ // prediction tasks grouped by held-out block (predict_blocks.cu)
		std::vector<std::vector<gsb::PredTask>> perBlock(static_cast<size_t>(nblocks));
		for (size_t t = 0; t < plan.tasks.size(); ++t) {
			gsb::PredTask pt;
			pt.fit = plan.tasks[t].fit;
			pt.kind = plan.tasks[t].kind;
			pt.dest = plan.tasks[t].dest;
			pt.pad = 0;
			perBlock[static_cast<size_t>(plan.tasks[t].block)].push_back(pt);
		}
		std::vector<int> active, off(1, 0);
		std::vector<gsb::PredTask> flat;
		for (int b = 0; b < nblocks; ++b) {
			if (perBlock[static_cast<size_t>(b)].empty()) continue;
			active.push_back(b);
			flat.insert(flat.end(), perBlock[static_cast<size_t>(b)].begin(), perBlock[static_cast<size_t>(b)].end());
			off.push_back(static_cast<int>(flat.size()));
		}
		e->nActiveBlocks = static_cast<int>(active.size());
		e->activeBlocks.upload(active, e->stream);
		e->blockTaskOff.upload(off, e->stream);
		e->predTasks.upload(flat, e->stream);
	}
	e->fits.upload(fdesc, e->stream);
	e->tasks.upload(tdesc, e->stream);
	e->blocks.upload(bdesc, e->stream);
	e->outerRows.upload(plan.outerRows, e->stream);
	GSB_CUDA(cudaStreamSynchronize(e->stream));

	if (e->cfg.evaluator == GSB_EVAL_LM) buildAllRowsGram(e);

	// ---- K1k: per-lane row masks (lane = row mod 32) of every distinct fit and prediction task
	e->kspaceFits = 0;
	// (PLS_CV and PLS_FIT: the all-rows refit of BICEvaluator is a fit whose "held-out" block is every row)
	if ((e->cfg.evaluator == GSB_EVAL_PLS_CV || e->cfg.evaluator == GSB_EVAL_PLS_FIT) && nfits > 0 &&
	    gsb::kspaceKernelFits(n, 1, e->smemLimit, 8) && !e->env.noKspace) {
		std::vector<gsb::KFitDesc> kf(static_cast<size_t>(nfits));
		std::vector<gsb::KTaskDesc> kt(plan.tasks.size());
		for (size_t t = 0; t < plan.tasks.size(); ++t) {
			const gsb::PlanTask& pt = plan.tasks[t];
			const gsb::PlanBlock& pb = plan.blocks[static_cast<size_t>(pt.block)];
			gsb::KTaskDesc& d = kt[t];
			for (int u = 0; u < gsb::kKspaceSlots; ++u) d.rows[u] = 0u;
			for (size_t i = 0; i < pb.rows.size(); ++i) d.rows[pb.rows[i] >> 5] |= 1u << (pb.rows[i] & 31u);
			d.kind = pt.kind;
			d.dest = pt.dest;
			d.nrows = static_cast<int>(pb.rows.size());
			d.rr = 1.0 / static_cast<double>(std::max<size_t>(1, pb.rows.size()));
		}
		for (int f = 0; f < nfits; ++f) {
			const gsb::PlanFit& pf = plan.fits[static_cast<size_t>(f)];
			gsb::KFitDesc& d = kf[static_cast<size_t>(f)];
			for (int u = 0; u < gsb::kKspaceSlots; ++u) d.train[u] = 0u;
			for (size_t i = 0; i < pf.rows.size(); ++i) d.train[pf.rows[i] >> 5] |= 1u << (pf.rows[i] & 31u);
			d.taskBegin = pf.taskBegin;
			d.ntasks = pf.taskCount;
			d.ntr = static_cast<int>(pf.rows.size());
			// centring of the response on the training rows, in the arithmetic of the device data (y - global mean)
			double sy = 0.0;
			for (size_t i = 0; i < pf.rows.size(); ++i) sy += e->y[pf.rows[i]] - e->yMean;
			d.rntr = 1.0 / static_cast<double>(std::max<size_t>(1, pf.rows.size()));
			d.ym = sy * d.rntr;
			double res = 0.0;
			for (size_t i = 0; i < pf.rows.size(); ++i) res += (e->y[pf.rows[i]] - e->yMean) - d.ym;
			d.syc = res;
			for (int j = 0; j < 2; ++j) {
				if (j < pf.taskCount) {
					d.first[j] = kt[static_cast<size_t>(pf.taskBegin + j)];
				} else {
					for (int u = 0; u < gsb::kKspaceSlots; ++u) d.first[j].rows[u] = 0u;
					d.first[j].kind = -1;
					d.first[j].dest = 0;
					d.first[j].nrows = 1;
					d.first[j].rr = 1.0;
				}
			}
		}
		e->kFits.upload(kf, e->stream);
		e->kTasks.upload(kt, e->stream);
		GSB_CUDA(cudaStreamSynchronize(e->stream));
		e->kspaceFits = nfits;
		// outer-only fits, and the dependency components of (fits, CV groups): a fit belongs with every group whose MSEP
		// table it feeds or whose optNComp it is fitted with (in a nested CV: one component per replication)
		const int I = std::max(1, plan.innerPerGroup);
		std::vector<int> parent(static_cast<size_t>(std::max(1, plan.groups)));
		std::iota(parent.begin(), parent.end(), 0);
		std::function<int(int)> find = [&](int x) { while (parent[static_cast<size_t>(x)] != x) x = parent[static_cast<size_t>(x)] = parent[static_cast<size_t>(parent[static_cast<size_t>(x)])]; return x; };
		auto groupOf = [&](const gsb::PlanTask& pt) { return pt.kind == gsb::TASK_OUTER ? pt.dest : pt.dest / I; };
		e->kFitPhase.assign(static_cast<size_t>(nfits), 1);
		for (int f = 0; f < nfits; ++f) {
			const gsb::PlanFit& pf = plan.fits[static_cast<size_t>(f)];
			if (pf.taskCount == 0) e->kFitPhase[static_cast<size_t>(f)] = 0;
			for (int t = pf.taskBegin; t < pf.taskBegin + pf.taskCount; ++t) {
				if (plan.tasks[static_cast<size_t>(t)].kind != gsb::TASK_OUTER) e->kFitPhase[static_cast<size_t>(f)] = 0;
				const int a = find(groupOf(plan.tasks[static_cast<size_t>(pf.taskBegin)])), b = find(groupOf(plan.tasks[static_cast<size_t>(t)]));
				if (a != b) parent[static_cast<size_t>(std::max(a, b))] = std::min(a, b);
			}
		}
		std::vector<int> compId(parent.size(), -1);
		e->kComponents = 0;
		for (size_t g = 0; g < parent.size(); ++g) {
			const int r = find(static_cast<int>(g));
			if (compId[static_cast<size_t>(r)] < 0) compId[static_cast<size_t>(r)] = e->kComponents++;
			compId[g] = compId[static_cast<size_t>(r)];
		}
		e->kFitComp.assign(static_cast<size_t>(nfits), 0);
		for (int f = 0; f < nfits; ++f) {
			const gsb::PlanFit& pf = plan.fits[static_cast<size_t>(f)];
			if (pf.taskCount > 0) e->kFitComp[static_cast<size_t>(f)] = compId[static_cast<size_t>(groupOf(plan.tasks[static_cast<size_t>(pf.taskBegin)]))];
		}
		for (int i = 0; i <= gsb_engine::kMaxSplit; ++i) e->kSchedules[i].ready = false;
		e->kSchedNf = -1;
		e->kProducts.reserve(1);
	}


	gsb_info& inf = e->info;
	inf.n = n; inf.p = p;
	inf.max_ncomp = e->shape.maxNComp;
	inf.inner_segments = e->shape.inner;
	inf.outer_segments = e->shape.outer;
	inf.num_replications = e->shape.replications;
	inf.segmentation_vectors = static_cast<int32_t>(e->segmentation.size());
	inf.fits_per_evaluation = plan.fitsPerEvaluation;
	inf.distinct_fits = nfits;
	inf.distinct_blocks = nblocks;
	inf.prediction_tasks = static_cast<int32_t>(plan.tasks.size());
	inf.sm_count = e->smCount;
	inf.block_bytes = static_cast<int64_t>((pool - static_cast<size_t>(ldz) * P2) * sizeof(double));
	inf.sdfact_effective = e->shape.sdfact;
	e->prepared = true;
	e->populationReady = e->resultsReady = false;
}

// Which K1 kernel serves each subset-size class of the resident population, decided once per upload; builds the tables
// a kernel needs the first time a class is given to it, and lays out the per-generation packed Grams of the per-fit kernel.
void planBuckets(gsb_engine* e) {
	const EnvFlags& env = e->env;
	const int nfits = static_cast<int>(e->plan.fits.size());
	const bool kspaceOn = e->kspaceFits > 0 && gsb::kspaceKernelFits(e->n, e->tableA, e->smemLimit, 8) && !env.noKspace;
	// K1x configurations in order of preference (measured, profiles/kb_cfg.sh): the cost is dominated by a per-CTA,
	// per-component overhead, so the fewest CTAs win: 16 fits per CTA, the stored loadings in the global scratch
	// (vGlobal) as soon as that keeps the subset in one CTA, then clusters of 2 and 3 CTAs splitting the columns
	static const int kScoreMinK = 40;
	static const int kScoreCfg[8][3] = {{16, 1, 0}, {8, 1, 0}, {8, 1, 1}, {8, 2, 0}, {16, 2, 1}, {8, 2, 1}, {16, 3, 1}, {8, 3, 1}};
	bool wantScores = false, wantFit = false;
	// K1k takes every class above kspaceMinK (its cost per fit does not depend on the subset size)
	for (size_t b = 0; b < e->buckets.size(); ++b) {
		Bucket& bk = e->buckets[b];
		bk.kernel = (kspaceOn && bk.kMax > env.kspaceMinK) ? BK_KSPACE : BK_FIT;
		if (bk.kernel == BK_FIT && (bk.kMax > kScoreMinK || env.forcedNf > 0)) wantScores = true;
	}
	// the K1k classes must be contiguous at the top (one launch): a class below a per-fit class goes back to the per-fit kernel
	bool top = true;
	for (size_t b = e->buckets.size(); b-- > 0;) {
		if (e->buckets[b].kernel != BK_KSPACE) top = false;
		else if (!top) e->buckets[b].kernel = BK_FIT;
	}
	wantScores = wantScores && e->cfg.evaluator == GSB_EVAL_PLS_CV && e->tableA <= gsb::kScoreComponents && !env.noScores &&
	             e->n <= 32 * gsb::kScoreMaxSlots;
	if (wantScores) buildScoreTables(e);
	const bool scoresOn = wantScores && e->scoreGroups8 > 0;
	long long vTotal = 0;
	for (size_t b = 0; b < e->buckets.size(); ++b) {
		Bucket& bk = e->buckets[b];
		if (bk.kernel != BK_FIT) continue;
		if (scoresOn && env.forcedNf > 0 && bk.kMax <= gsb::scoreKernelMaxK(e->scoreNr, env.forcedNf, env.forcedCs, env.forcedVg, e->smemLimit)) {
			bk.kernel = BK_SCORES;
			bk.nf = env.forcedNf; bk.cs = env.forcedCs; bk.vg = env.forcedVg;
		}
		// small subsets: the per-fit kernel (one CTA per fit, many CTAs per SM) is faster below ~40 columns
		for (int c = 0; scoresOn && env.forcedNf == 0 && bk.kMax > kScoreMinK && c < 8 && bk.kernel == BK_FIT; ++c) {
			if (bk.kMax <= gsb::scoreKernelMaxK(e->scoreNr, kScoreCfg[c][0], kScoreCfg[c][1], kScoreCfg[c][2], e->smemLimit)) {
				bk.kernel = BK_SCORES;
				bk.nf = kScoreCfg[c][0]; bk.cs = kScoreCfg[c][1]; bk.vg = kScoreCfg[c][2];
			}
		}
		if (bk.kernel == BK_SCORES && bk.vg) {
			const long long groupsHere = bk.nf == 16 ? e->scoreGroups16 : e->scoreGroups8;
			bk.vPerCta = gsb::scoreKernelScratchDoubles(bk.kMax, bk.nf, bk.cs);
			bk.vOffset = vTotal;
			bk.vSlots = gsb::scoreKernelScratchSlots(bk.kMax, bk.nf, bk.cs, e->scoreNr, groupsHere * bk.count * bk.cs);
			vTotal += bk.vPerCta * (bk.vSlots < 0 ? -bk.vSlots : bk.vSlots);
		}
		if (bk.kernel == BK_FIT) wantFit = true;
	}
	if (vTotal > 0) e->scoreScratch.reserve(static_cast<size_t>(vTotal));

	// per-fit kernel: one slab of packed Gram and one of vectors per (chromosome, training set)
	e->hPackOff.reserve(static_cast<size_t>(e->npop) + 1);
	e->hVecOff.reserve(static_cast<size_t>(e->npop) + 1);
	e->hCoefOff.reserve(static_cast<size_t>(e->npop) + 1);
	for (int c = 0; c <= e->npop; ++c) e->hPackOff.ptr[c] = e->hVecOff.ptr[c] = e->hCoefOff.ptr[c] = 0ull;
	if (!wantFit) return;
	try {
		buildUnitTables(e);
	} catch (const std::string&) {
		// the moment tables of the Gram-space path do not fit in device memory (p in the tens of thousands): when the
		// kernel-space kernel can serve the plan it takes the small subsets too (it needs no tables at all)
		bool allKspace = kspaceOn && !e->unitsReady;
		for (size_t b = 0; b < e->buckets.size() && allKspace; ++b) allKspace = e->buckets[b].kernel == BK_KSPACE || e->buckets[b].kernel == BK_FIT;
		if (!allKspace) throw;
		for (size_t b = 0; b < e->buckets.size(); ++b) e->buckets[b].kernel = BK_KSPACE;
		return;
	}
	unsigned long long packTotal = 0, vecTotal = 0, coefTotal = 0;
	for (size_t b = 0; b < e->buckets.size(); ++b) {
		Bucket& bk = e->buckets[b];
		if (bk.kernel != BK_FIT) continue;
		// fit-only kernel + one tensor-core prediction per (chromosome, held-out block) where the fast path serves the class
		// (measured, profiles/pred_probe.py: the separate prediction wins by 20-30 % with the 100-row blocks of config 4 and
		// loses 10-20 % with the 30-row blocks of config 3, whose tiles leave the tensor-core product mostly empty)
		bk.split = !env.noPredict && e->nActiveBlocks > 0 && (e->maxBlockLd >= 64 || env.forcePredict) && bk.kMax <= gsb::kPredMaxK &&
		           gsb::fitOnlyKernelServes(bk.kMax, e->tableA, e->smemLimit);
		for (int i = 0; i < bk.count; ++i) {
			const int c = e->hBucket.ptr[bk.begin + i];
			const unsigned long long k = e->hOffsets.ptr[c + 1] - e->hOffsets.ptr[c];
			e->hPackOff.ptr[c] = packTotal;
			e->hVecOff.ptr[c] = vecTotal;
			packTotal += gsb::packedDoubles(k) * static_cast<unsigned long long>(nfits);
			vecTotal += 2ull * ((k + 1ull) & ~1ull) * static_cast<unsigned long long>(nfits);
			if (bk.split) {
				e->hCoefOff.ptr[c] = coefTotal;
				const unsigned long long A = std::min<unsigned long long>(static_cast<unsigned long long>(e->tableA), k);
				coefTotal += ((k + 2ull) * A * static_cast<unsigned long long>(nfits) + 1ull) & ~1ull;
			}
		}
	}
	// generic kernel (more than ten components): when the stored loadings and coefficients of a class outgrow shared
	// memory they live in a global scratch, one region per CTA of a launch; the class is then walked in chunks of
	// chromosomes (launchFitKernel).  One region for all such classes: they are launched on one stream.
	unsigned long long genWant = 0, genMin = 0;
	for (size_t b = 0; b < e->buckets.size(); ++b) {
		Bucket& bk = e->buckets[b];
		bk.genPerCta = 0;
		if (bk.kernel != BK_FIT || bk.split) continue;
		bk.genPerCta = gsb::fitKernelScratchPerCta(bk.kMax, e->tableA, e->maxTasks, e->maxBlockLd, e->smemLimit);
		const unsigned long long perChrom = bk.genPerCta * static_cast<unsigned long long>(nfits);
		genMin = std::max(genMin, perChrom);
		genWant = std::max(genWant, perChrom * static_cast<unsigned long long>(std::min(bk.count, 2 * std::max(1, e->smCount))));
	}
	if (genWant > e->genScratch.capacity) {
		size_t freeB = 0, totalB = 0;
		GSB_CUDA(cudaMemGetInfo(&freeB, &totalB));
		const unsigned long long room = static_cast<unsigned long long>(0.25 * static_cast<double>(freeB)) / 8ull;
		const unsigned long long take = std::max(genMin, std::min(genWant, room));
		if (8.0 * static_cast<double>(take) > 0.9 * static_cast<double>(freeB)) {
			throw std::string("the stored loadings of this many components on subsets this wide do not fit in device memory");
		}
		e->genScratch.reserve(static_cast<size_t>(take));
	}
	if (packTotal + vecTotal + coefTotal > e->gpack.capacity + e->pvec.capacity + e->coefScratch.capacity) {
		size_t freeB = 0, totalB = 0;
		GSB_CUDA(cudaMemGetInfo(&freeB, &totalB));
		const double need = 8.0 * static_cast<double>(packTotal + vecTotal + coefTotal);
		const double have = static_cast<double>(freeB) + 8.0 * static_cast<double>(e->gpack.capacity + e->pvec.capacity + e->coefScratch.capacity);
		if (need > 0.95 * have) {
			char buf[256];
			std::snprintf(buf, sizeof(buf), "the packed Grams of this population need %.1f GB of device memory, %.1f GB are free: evaluate it in smaller batches",
			              need / 1e9, have / 1e9);
			throw std::string(buf);
		}
	}
	e->gpack.reserve(static_cast<size_t>(packTotal) + 2);
	e->pvec.reserve(static_cast<size_t>(vecTotal) + 2);
	if (coefTotal > 0) e->coefScratch.reserve(static_cast<size_t>(coefTotal) + 2);
}

int prepareChecked(gsb_engine* e) {
	if (!e->haveData) return fail(e, GSB_ERR_STATE, "gsb_set_data has not been called");
	if (e->cfg.evaluator != GSB_EVAL_LM && !e->haveSegmentation) {
		return fail(e, GSB_ERR_STATE, "no CV segmentation: call gsb_init_segmentation or gsb_set_segmentation");
	}
	try {
		if (e->cfg.evaluator == GSB_EVAL_LM) {
			gsb::buildAllRowsPlan(e->n, e->plan);
			e->shape = gsb::CvShape();
		} else {
			const std::string msg = gsb::buildCvPlan(e->n, e->segmentation, e->shape, e->cfg.evaluator == GSB_EVAL_PLS_FIT, e->plan);
			if (!msg.empty()) return fail(e, GSB_ERR_INVALID_ARGUMENT, msg);
		}
		prepareDevice(e);
	} catch (const CudaFailure& f) {
		return failCuda(e, f);
	} catch (const std::string& msg) {
		return fail(e, GSB_ERR_MEMORY, msg);
	} catch (const std::bad_alloc&) {
		return fail(e, GSB_ERR_MEMORY, "host allocation failed");
	}
	return GSB_OK;
}

std::string deriveShape(gsb_engine* e, std::vector<gsb::RowSet>& seg, const uint32_t* seed624) {
	const gsb_config& c = e->cfg;
	if (c.evaluator == GSB_EVAL_PLS_CV) {
		return gsb::buildPlsSegmentation(e->n, c.num_replications, c.max_ncomp, c.inner_segments, c.outer_segments,
		                                 c.test_set_size, c.sdfact, seed624, seg, e->shape);
	}
	return gsb::buildFitSegmentation(e->n, c.max_ncomp, c.inner_segments, c.sdfact, seed624, seg, e->shape);
}


// ---- multi-device engine -----------------------------------------------------------------------------------
// The per-device engines ("replicas") are complete single-device engines: each holds its own copy of the data, builds
// the same segmentation from the same seed and its own device tables (nothing is exchanged between GPUs at setup).
// Per batch: deal -> concurrent upload / kernels on every replica -> peer copies of the result vectors to the first
// device -> one copy to pinned host memory -> un-permute into the caller's order.

int resolveRetries(gsb_engine* e); // below, next to the download
std::string checkGuards(gsb_engine* e);

int failFrom(gsb_engine* e, int rc, const gsb_engine* child, int dev) {
	char buf[768];
	std::snprintf(buf, sizeof(buf), "device %d: %s", dev, child ? child->error.c_str() : "");
	return fail(e, rc, buf);
}

// runs fn(replica) on every replica concurrently (the pool's threads), propagates the first failure
int onReplicas(gsb_engine* e, const std::function<int(gsb_engine*)>& fn) {
	int which = 0;
	const int rc = e->multi->pool->run([&](int i) { return fn(e->replicas[static_cast<size_t>(i)]); }, &which);
	if (rc != GSB_OK) return failFrom(e, rc, e->replicas[static_cast<size_t>(which)], e->cfg.devices[which]);
	return GSB_OK;
}

// cost model of the deal (multigpu.py holds the same constants for the one-process-per-GPU host): milliseconds per 296
// chromosomes measured at BASELINE config 3 for the kernel-space kernel, k^2 for the per-fit Gram-space kernel
double dealCost(double k, bool kspace) {
	if (kspace) return k > 40.0 ? 1.24 + 0.00023 * k : 0.38 + 0.0131 * k;
	return k * k + 1.0;
}

void dealPopulation(const uint32_t* offsets, int npop, int ndev, bool kspace, std::vector<int>& deviceOf) {
	deviceOf.assign(static_cast<size_t>(npop), 0);
	if (ndev <= 1) return;
	std::vector<double> cost(static_cast<size_t>(npop));
	std::vector<int> order(static_cast<size_t>(npop));
	for (int c = 0; c < npop; ++c) {
		cost[static_cast<size_t>(c)] = dealCost(static_cast<double>(offsets[c + 1] - offsets[c]), kspace);
		order[static_cast<size_t>(c)] = c;
	}
	std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[static_cast<size_t>(a)] > cost[static_cast<size_t>(b)]; });
	const int cap = (npop + ndev - 1) / ndev;
	std::vector<double> load(static_cast<size_t>(ndev), 0.0);
	std::vector<int> count(static_cast<size_t>(ndev), 0);
	for (int x = 0; x < npop; ++x) {
		const int c = order[static_cast<size_t>(x)];
		int best = -1;
		for (int d = 0; d < ndev; ++d) {
			if (count[static_cast<size_t>(d)] >= cap) continue;
			if (best < 0 || load[static_cast<size_t>(d)] < load[static_cast<size_t>(best)]) best = d;
		}
		deviceOf[static_cast<size_t>(c)] = best;
		load[static_cast<size_t>(best)] += cost[static_cast<size_t>(c)];
		count[static_cast<size_t>(best)] += 1;
	}
}

bool multiUsesKspace(const gsb_engine* e) {
	const gsb_engine* r0 = e->replicas[0];
	return r0->kspaceFits > 0 && !r0->env.noKspace;
}

void multiReleaseGather(gsb_engine* e) {
	MultiState* ms = e->multi;
	if (!ms) return;
	if (ms->dFitness || ms->stream) cudaSetDevice(e->cfg.devices[0]);
	if (ms->dFitness) cudaFree(ms->dFitness);
	if (ms->dStatus) cudaFree(ms->dStatus);
	if (ms->hFitness) cudaFreeHost(ms->hFitness);
	if (ms->hStatus) cudaFreeHost(ms->hStatus);
	if (ms->stream) cudaStreamDestroy(ms->stream);
	ms->dFitness = nullptr; ms->dStatus = nullptr; ms->hFitness = nullptr; ms->hStatus = nullptr; ms->stream = nullptr;
	ms->capacity = 0;
}

// deal + the per-device CSR lists of one batch (ms->shares / idx / off / first)
int multiDeal(gsb_engine* e, const uint32_t* idx, const uint32_t* offsets, int32_t npop) {
	MultiState* ms = e->multi;
	const int ndev = static_cast<int>(e->replicas.size());
	ms->populationReady = ms->resultsReady = false;
	for (int c = 0; c < npop; ++c) {
		if (offsets[c + 1] < offsets[c]) return fail(e, GSB_ERR_INVALID_ARGUMENT, "population offsets must be non-decreasing");
	}
	// the replicas must be prepared before the deal: the cost model depends on which kernels the plan admits
	bool allPrepared = true;
	for (int d = 0; d < ndev; ++d) allPrepared = allPrepared && e->replicas[static_cast<size_t>(d)]->prepared;
	if (!allPrepared) {
		const int rc = onReplicas(e, [](gsb_engine* r) { return gsb_prepare(r); });
		if (rc != GSB_OK) return rc;
	}
	std::vector<int> deviceOf;
	dealPopulation(offsets, npop, ndev, multiUsesKspace(e), deviceOf);
	ms->shares.resize(static_cast<size_t>(ndev));
	ms->idx.resize(static_cast<size_t>(ndev));
	ms->off.resize(static_cast<size_t>(ndev));
	for (int d = 0; d < ndev; ++d) ms->shares[static_cast<size_t>(d)].clear(); // clear, not assign: the capacities of the last batch are kept
	for (int c = 0; c < npop; ++c) ms->shares[static_cast<size_t>(deviceOf[static_cast<size_t>(c)])].push_back(c);
	ms->first.assign(static_cast<size_t>(ndev) + 1, 0);
	for (int d = 0; d < ndev; ++d) {
		ms->first[static_cast<size_t>(d) + 1] = ms->first[static_cast<size_t>(d)] + static_cast<int>(ms->shares[static_cast<size_t>(d)].size());
	}
	ms->npop = npop;
	return GSB_OK;
}

// the CSR lists of device d's share, copied out of the caller's arrays (on that device's thread: at 20 000 chromosomes
// of 200 columns the lists are 16 MB)
void multiBuildShare(MultiState* ms, size_t d, const uint32_t* idx, const uint32_t* offsets) {
	std::vector<uint32_t>& di = ms->idx[d];
	std::vector<uint32_t>& dof = ms->off[d];
	const std::vector<int>& sh = ms->shares[d];
	size_t total = 0;
	for (size_t i = 0; i < sh.size(); ++i) total += offsets[sh[i] + 1] - offsets[sh[i]];
	di.resize(std::max<size_t>(total, 1));
	dof.resize(sh.size() + 1);
	dof[0] = 0u;
	size_t at = 0;
	for (size_t i = 0; i < sh.size(); ++i) {
		const size_t len = offsets[sh[i] + 1] - offsets[sh[i]];
		if (len) std::memcpy(di.data() + at, idx + offsets[sh[i]], len * sizeof(uint32_t));
		at += len;
		dof[i + 1] = static_cast<uint32_t>(at);
	}
	if (total == 0) di[0] = 0u;
}

int multiFailed(gsb_engine* e, int rc, const char* what) {
	for (size_t d = 0; d < e->replicas.size(); ++d) if (!e->replicas[d]->error.empty()) return failFrom(e, rc, e->replicas[d], e->cfg.devices[d]);
	return fail(e, rc, what);
}

void multiCollectTiming(gsb_engine* e) {
	gsb_timing& t = e->timing;
	std::memset(&t, 0, sizeof(t));
	for (size_t d = 0; d < e->replicas.size(); ++d) {
		const gsb_timing& c = e->replicas[d]->timing;
		if (e->replicas[d]->npop == 0) continue;
		t.h2d_ms = std::max(t.h2d_ms, c.h2d_ms);
		t.fit_kernel_ms = std::max(t.fit_kernel_ms, c.fit_kernel_ms);
		t.reduce_kernel_ms = std::max(t.reduce_kernel_ms, c.reduce_kernel_ms);
		t.d2h_ms = std::max(t.d2h_ms, c.d2h_ms);
		t.total_ms = std::max(t.total_ms, c.total_ms);
		t.fit_ctas += c.fit_ctas;
		t.kernel_launches += c.kernel_launches;
		t.fit_tensor_gflop += c.fit_tensor_gflop;
		t.retries += c.retries;
	}
}

int multiUpload(gsb_engine* e, const uint32_t* idx, const uint32_t* offsets, int32_t npop) {
	MultiState* ms = e->multi;
	int rc = multiDeal(e, idx, offsets, npop);
	if (rc != GSB_OK) return rc;
	rc = e->multi->pool->run([&](int i) {
		const size_t d = static_cast<size_t>(i);
		try {
			multiBuildShare(ms, d, idx, offsets);
		} catch (const std::bad_alloc&) {
			return fail(e->replicas[d], GSB_ERR_MEMORY, "host allocation failed");
		}
		return gsb_population_upload(e->replicas[d], ms->idx[d].data(), ms->off[d].data(), static_cast<int32_t>(ms->shares[d].size()));
	}, nullptr);
	if (rc != GSB_OK) return multiFailed(e, rc, "population upload failed");
	ms->populationReady = true;
	return GSB_OK;
}

// gsb_evaluate_indices on a multi-device engine: deal, then ONE hand-over to the device threads -- every replica runs
// its whole single-device call (its lists copied out of the caller's arrays, upload, kernels, re-scoring of declined
// chromosomes, download into its slice of the gathered host vectors) -- and the un-permutation into the caller's order.
// No copy between GPUs and a third of the hand-overs of upload + evaluate + download (which gather through device 0, so
// that the results are also resident on one device): 0.96 -> 0.49 ms per 500 chromosomes on 8 GPUs, kernels 0.36 ms.
int multiEvaluateIndices(gsb_engine* e, const uint32_t* idx, const uint32_t* offsets, int32_t npop, double* fitness, int32_t* status) {
	MultiState* ms = e->multi;
	int rc = multiDeal(e, idx, offsets, npop);
	if (rc != GSB_OK) return rc;
	if (npop == 0) return GSB_OK;
	try {
		ms->gatherF.resize(static_cast<size_t>(npop));
		ms->gatherS.resize(static_cast<size_t>(npop));
	} catch (const std::bad_alloc&) {
		return fail(e, GSB_ERR_MEMORY, "host allocation failed");
	}
	rc = ms->pool->run([&](int i) {
		const size_t d = static_cast<size_t>(i);
		gsb_engine* r = e->replicas[d];
		const int32_t cnt = static_cast<int32_t>(ms->shares[d].size());
		if (cnt == 0) { r->npop = 0; return static_cast<int>(GSB_OK); }
		try {
			multiBuildShare(ms, d, idx, offsets);
		} catch (const std::bad_alloc&) {
			return fail(r, GSB_ERR_MEMORY, "host allocation failed");
		}
		const size_t at = static_cast<size_t>(ms->first[d]);
		return gsb_evaluate_indices(r, ms->idx[d].data(), ms->off[d].data(), cnt, ms->gatherF.data() + at, ms->gatherS.data() + at);
	}, nullptr);
	if (rc != GSB_OK) return multiFailed(e, rc, "evaluation failed");
	for (size_t d = 0; d < e->replicas.size(); ++d) {
		const std::vector<int>& sh = ms->shares[d];
		const size_t at = static_cast<size_t>(ms->first[d]);
		for (size_t i = 0; i < sh.size(); ++i) {
			fitness[sh[i]] = ms->gatherF[at + i];
			status[sh[i]] = ms->gatherS[at + i];
		}
	}
	multiCollectTiming(e);
	ms->populationReady = true; // every replica holds its share and its results: gsb_population_download works afterwards
	ms->resultsReady = true;
	return GSB_OK;
}

int multiEvaluate(gsb_engine* e, bool synchronous) {
	MultiState* ms = e->multi;
	if (!ms->populationReady) return fail(e, GSB_ERR_STATE, "no resident population: call gsb_population_upload first");
	int rc;
	if (synchronous) {
		rc = onReplicas(e, [](gsb_engine* r) { return r->npop > 0 ? gsb_population_evaluate(r) : GSB_OK; });
	} else {
		// enqueue only: the launches are asynchronous, one device after the other from this thread
		rc = GSB_OK;
		for (size_t d = 0; d < e->replicas.size() && rc == GSB_OK; ++d) {
			gsb_engine* r = e->replicas[d];
			if (r->npop > 0) rc = gsb_population_enqueue(r);
			if (rc != GSB_OK) return failFrom(e, rc, r, e->cfg.devices[d]);
		}
	}
	if (rc != GSB_OK) return rc;
	multiCollectTiming(e);
	ms->resultsReady = true;
	return GSB_OK;
}

int multiDownload(gsb_engine* e, double* fitness, int32_t* status) {
	MultiState* ms = e->multi;
	if (!ms->resultsReady) return fail(e, GSB_ERR_STATE, "no results: call gsb_population_evaluate first");
	const int ndev = static_cast<int>(e->replicas.size());
	const size_t np = static_cast<size_t>(ms->npop);
	if (np == 0) return GSB_OK;
	try {
		const int dev0 = e->cfg.devices[0];
		// every replica's kernels must have finished before its result vectors are read from another device
		for (int d = 0; d < ndev; ++d) {
			gsb_engine* r = e->replicas[static_cast<size_t>(d)];
			if (r->npop == 0) continue;
			GSB_CUDA(cudaSetDevice(e->cfg.devices[d]));
			GSB_CUDA(cudaStreamSynchronize(r->stream));
			if (r->refitArmed) { // chromosomes its kernel-space kernel declined: re-scored on that device before the gather
				r->hStatus.reserve(static_cast<size_t>(r->npop) + 1);
				r->hFitness.reserve(static_cast<size_t>(r->npop) + 1);
				GSB_CUDA(cudaMemcpyAsync(r->hStatus.ptr, r->dStatus.ptr, static_cast<size_t>(r->npop) * sizeof(int), cudaMemcpyDeviceToHost, r->stream));
				GSB_CUDA(cudaStreamSynchronize(r->stream));
				const int rc = resolveRetries(r);
				if (rc != GSB_OK) return failFrom(e, rc, r, e->cfg.devices[d]);
				e->timing.retries += r->timing.retries;
			}
		}
		GSB_CUDA(cudaSetDevice(dev0));
		if (!ms->stream) GSB_CUDA(cudaStreamCreateWithFlags(&ms->stream, cudaStreamNonBlocking));
		if (np > ms->capacity) {
			const size_t want = np + np / 2 + 64;
			if (ms->dFitness) cudaFree(ms->dFitness);
			if (ms->dStatus) cudaFree(ms->dStatus);
			if (ms->hFitness) cudaFreeHost(ms->hFitness);
			if (ms->hStatus) cudaFreeHost(ms->hStatus);
			ms->dFitness = nullptr; ms->dStatus = nullptr; ms->hFitness = nullptr; ms->hStatus = nullptr; ms->capacity = 0;
			GSB_CUDA(cudaMalloc(reinterpret_cast<void**>(&ms->dFitness), want * sizeof(double)));
			GSB_CUDA(cudaMalloc(reinterpret_cast<void**>(&ms->dStatus), want * sizeof(int)));
			GSB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ms->hFitness), want * sizeof(double), cudaHostAllocDefault));
			GSB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ms->hStatus), want * sizeof(int), cudaHostAllocDefault));
			ms->capacity = want;
		}
		// gather: peer copies (NVLink when peer access is enabled, else staged by the driver) into the first device
		for (int d = 0; d < ndev; ++d) {
			gsb_engine* r = e->replicas[static_cast<size_t>(d)];
			const size_t cnt = ms->shares[static_cast<size_t>(d)].size();
			if (cnt == 0) continue;
			const size_t at = static_cast<size_t>(ms->first[static_cast<size_t>(d)]);
			GSB_CUDA(cudaMemcpyPeerAsync(ms->dFitness + at, dev0, r->dFitness.ptr, e->cfg.devices[d], cnt * sizeof(double), ms->stream));
			GSB_CUDA(cudaMemcpyPeerAsync(ms->dStatus + at, dev0, r->dStatus.ptr, e->cfg.devices[d], cnt * sizeof(int), ms->stream));
		}
		GSB_CUDA(cudaMemcpyAsync(ms->hFitness, ms->dFitness, np * sizeof(double), cudaMemcpyDeviceToHost, ms->stream));
		GSB_CUDA(cudaMemcpyAsync(ms->hStatus, ms->dStatus, np * sizeof(int), cudaMemcpyDeviceToHost, ms->stream));
		GSB_CUDA(cudaStreamSynchronize(ms->stream));
		for (int d = 0; d < ndev; ++d) {
			const std::vector<int>& sh = ms->shares[static_cast<size_t>(d)];
			const size_t at = static_cast<size_t>(ms->first[static_cast<size_t>(d)]);
			for (size_t i = 0; i < sh.size(); ++i) {
				fitness[sh[i]] = ms->hFitness[at + i];
				status[sh[i]] = ms->hStatus[at + i];
			}
		}
	} catch (const CudaFailure& f) {
		return failCuda(e, f);
	}
	return GSB_OK;
}

} // namespace

extern "C" {

int gsb_abi_version(void) { return GSB_ABI_VERSION; }

int gsb_device_count(int32_t* count) {
	if (!count) return GSB_ERR_INVALID_ARGUMENT;
	int n = 0;
	const cudaError_t err = cudaGetDeviceCount(&n);
	if (err != cudaSuccess) cudaGetLastError();
	*count = (err == cudaSuccess) ? n : 0;
	return (err == cudaSuccess && n > 0) ? GSB_OK : GSB_ERR_CUDA;
}

int gsb_create(gsb_engine** out, const gsb_config* cfg) {
	if (!out || !cfg) return fail(nullptr, GSB_ERR_INVALID_ARGUMENT, "null argument");
	*out = nullptr;
	if (cfg->struct_size != static_cast<int32_t>(sizeof(gsb_config))) {
		return fail(nullptr, GSB_ERR_INVALID_ARGUMENT, "gsb_config.struct_size does not match this library");
	}
	if (cfg->device < 0) return fail(nullptr, GSB_ERR_INVALID_ARGUMENT, "negative device ordinal");
	if (cfg->num_devices < 0 || cfg->num_devices > GSB_MAX_DEVICES) return fail(nullptr, GSB_ERR_INVALID_ARGUMENT, "num_devices out of range");
	for (int i = 0; i < cfg->num_devices; ++i) {
		if (cfg->devices[i] < 0) return fail(nullptr, GSB_ERR_INVALID_ARGUMENT, "negative device ordinal in the device list");
		for (int j = 0; j < i; ++j) if (cfg->devices[j] == cfg->devices[i]) return fail(nullptr, GSB_ERR_INVALID_ARGUMENT, "the device list must hold distinct ordinals");
	}
	switch (cfg->evaluator) {
	case GSB_EVAL_PLS_CV:
		if (cfg->num_replications < 1) return fail(nullptr, GSB_ERR_INVALID_ARGUMENT, "numReplications must be at least 1");
		// src/PLSEvaluator.cpp:50-52 (rdCV overrides the test set size, :46-48)
		if (cfg->outer_segments <= 1 && (cfg->test_set_size < 0.0 || cfg->test_set_size >= 1.0)) {
			return fail(nullptr, GSB_ERR_INVALID_ARGUMENT, "The test set size must be within the interval (0, 1)");
		}
		if (cfg->inner_segments < 1) return fail(nullptr, GSB_ERR_INVALID_ARGUMENT, "innerSegments must be positive");
		if (cfg->max_ncomp < 0) return fail(nullptr, GSB_ERR_INVALID_ARGUMENT, "maxNComp must not be negative");
		break;
	case GSB_EVAL_PLS_FIT:
		if (cfg->inner_segments < 2) return fail(nullptr, GSB_ERR_INVALID_ARGUMENT, "For CV at least 2 segments are needed");
		if (cfg->statistic < GSB_STAT_BIC || cfg->statistic > GSB_STAT_R2) return fail(nullptr, GSB_ERR_INVALID_ARGUMENT, "unknown statistic");
		break;
	case GSB_EVAL_LM:
		if (cfg->statistic < GSB_STAT_BIC || cfg->statistic > GSB_STAT_R2) return fail(nullptr, GSB_ERR_INVALID_ARGUMENT, "unknown statistic");
		break;
	default:
		return fail(nullptr, GSB_ERR_INVALID_ARGUMENT, "unknown evaluator kind (the user-function evaluator is out of scope)");
	}
	gsb_engine* e = new (std::nothrow) gsb_engine();
	if (!e) return fail(nullptr, GSB_ERR_MEMORY, "host allocation failed");
	g_guard = std::getenv("GSB_GUARD") != nullptr; // (arrays remember whether they were allocated with a guard zone)
	e->cfg = *cfg;
	if (cfg->num_devices == 1) e->cfg.device = cfg->devices[0];
	if (cfg->num_devices > 1) {
		// one complete single-device engine per GPU; this object deals the batches and gathers the results
		try {
			for (int i = 0; i < cfg->num_devices; ++i) {
				gsb_config one = *cfg;
				one.num_devices = 0;
				one.device = cfg->devices[i];
				gsb_engine* r = nullptr;
				const int rc = gsb_create(&r, &one);
				if (rc != GSB_OK) { gsb_destroy(e); return rc; }
				e->replicas.push_back(r);
			}
			e->multi = new MultiState();
			e->multi->pool = new DevicePool(cfg->num_devices);
		} catch (const std::exception&) {
			gsb_destroy(e);
			return fail(nullptr, GSB_ERR_MEMORY, "host allocation failed");
		}
	}
	*out = e;
	return GSB_OK;
}

void gsb_destroy(gsb_engine* e) {
	if (!e) return;
	if (e->multi) {
		multiReleaseGather(e);
		delete e->multi; // joins the device threads
		e->multi = nullptr;
	}
	for (size_t i = 0; i < e->replicas.size(); ++i) gsb_destroy(e->replicas[i]);
	e->replicas.clear();
	if (e->retryEngine) gsb_destroy(e->retryEngine);
	e->retryEngine = nullptr;
	if (e->deviceReady) cudaSetDevice(e->cfg.device);
	releaseDevice(e);
	delete e;
}

int gsb_last_error(const gsb_engine* e, char* buf, size_t buflen) {
	if (!buf || buflen == 0) return GSB_ERR_INVALID_ARGUMENT;
	const std::string& msg = e ? e->error : g_createError;
	std::snprintf(buf, buflen, "%s", msg.c_str());
	return GSB_OK;
}

int gsb_set_data(gsb_engine* e, const double* X_colmajor, const double* y, int32_t n, int32_t p) {
	if (!e) return GSB_ERR_INVALID_ARGUMENT;
	if (!X_colmajor || !y || n < 4 || p < 1) return fail(e, GSB_ERR_INVALID_ARGUMENT, "X must be n x p with n >= 4, p >= 1");
	if (n > 65535 || p > 65535) return fail(e, GSB_ERR_INVALID_ARGUMENT, "n and p are limited to 65535 (src/Control.h:67-68, src/OnlineStddev.h:62)");
	if (!e->replicas.empty()) { // every replica keeps its own copy (it uploads from it); this object only records the shape
		for (size_t i = 0; i < e->replicas.size(); ++i) {
			const int rc = gsb_set_data(e->replicas[i], X_colmajor, y, n, p);
			if (rc != GSB_OK) return failFrom(e, rc, e->replicas[i], e->cfg.devices[i]);
		}
		e->n = n;
		e->p = p;
		e->haveData = true;
		e->multi->populationReady = e->multi->resultsReady = false;
		return GSB_OK;
	}
	try {
		e->X.assign(X_colmajor, X_colmajor + static_cast<size_t>(n) * p);
		e->y.assign(y, y + n);
	} catch (const std::bad_alloc&) {
		return fail(e, GSB_ERR_MEMORY, "host allocation failed");
	}
	e->n = n;
	e->p = p;
	e->haveData = true;
	e->haveSegmentation = false;
	e->prepared = e->zReady = e->populationReady = e->resultsReady = false;
	return GSB_OK;
}

int gsb_seed_vector(uint32_t seed, uint32_t* out624) {
	if (!out624) return GSB_ERR_INVALID_ARGUMENT;
	gsb::Well19937a rng(seed);
	for (int i = 0; i < GSB_SEED_WORDS; ++i) out624[i] = rng.next();
	return GSB_OK;
}

int gsb_rng_stream(const uint32_t* seed624, int32_t n, uint32_t* out) {
	if (!seed624 || !out || n < 0) return GSB_ERR_INVALID_ARGUMENT;
	gsb::Well19937a rng(seed624);
	for (int i = 0; i < n; ++i) out[i] = rng.next();
	return GSB_OK;
}

int gsb_shuffle_all(const uint32_t* seed624, int32_t size, int32_t ncalls, uint32_t* out) {
	if (!seed624 || !out || size < 1 || ncalls < 0) return GSB_ERR_INVALID_ARGUMENT;
	gsb::Well19937a rng(seed624);
	gsb::RowShuffler rows(static_cast<uint32_t>(size));
	for (int c = 0; c < ncalls; ++c) {
		const std::vector<uint32_t>& s = rows.shuffleAll(rng);
		std::copy(s.begin(), s.end(), out + static_cast<size_t>(c) * size);
	}
	return GSB_OK;
}

int gsb_init_segmentation(gsb_engine* e, const uint32_t* seed624) {
	if (!e) return GSB_ERR_INVALID_ARGUMENT;
	if (e->cfg.evaluator == GSB_EVAL_LM) return GSB_OK;
	if (!seed624) return fail(e, GSB_ERR_INVALID_ARGUMENT, "null seed");
	if (!e->haveData) return fail(e, GSB_ERR_STATE, "gsb_set_data has not been called");
	for (size_t i = 0; i < e->replicas.size(); ++i) { // the same host-built segmentation on every replica
		const int rc = gsb_init_segmentation(e->replicas[i], seed624);
		if (rc != GSB_OK) return failFrom(e, rc, e->replicas[i], e->cfg.devices[i]);
	}
	if (!e->replicas.empty()) { e->multi->populationReady = e->multi->resultsReady = false; return GSB_OK; }
	const std::string msg = deriveShape(e, e->segmentation, seed624);
	if (!msg.empty()) return fail(e, GSB_ERR_INVALID_ARGUMENT, msg);
	e->haveSegmentation = true;
	e->prepared = false;
	return GSB_OK;
}

int gsb_set_segmentation(gsb_engine* e, const uint32_t* rows, const uint32_t* offsets, int32_t nvec) {
	if (!e) return GSB_ERR_INVALID_ARGUMENT;
	if (e->cfg.evaluator == GSB_EVAL_LM) return GSB_OK;
	if (!rows || !offsets || nvec < 2) return fail(e, GSB_ERR_INVALID_ARGUMENT, "null or empty segmentation");
	if (!e->haveData) return fail(e, GSB_ERR_STATE, "gsb_set_data has not been called");
	for (size_t i = 0; i < e->replicas.size(); ++i) {
		const int rc = gsb_set_segmentation(e->replicas[i], rows, offsets, nvec);
		if (rc != GSB_OK) return failFrom(e, rc, e->replicas[i], e->cfg.devices[i]);
	}
	if (!e->replicas.empty()) { e->multi->populationReady = e->multi->resultsReady = false; return GSB_OK; }
	// segment lengths and the maxNComp clamp depend only on n and the configuration
	std::vector<uint32_t> zeros(GSB_SEED_WORDS, 0u);
	zeros[0] = 1u;
	std::vector<gsb::RowSet> own;
	const std::string msg = deriveShape(e, own, zeros.data());
	if (!msg.empty()) return fail(e, GSB_ERR_INVALID_ARGUMENT, msg);
	if (static_cast<size_t>(nvec) != own.size()) return fail(e, GSB_ERR_INVALID_ARGUMENT, "segmentation has the wrong number of index vectors for this evaluator");
	std::vector<gsb::RowSet> seg(static_cast<size_t>(nvec));
	for (int v = 0; v < nvec; ++v) {
		if (offsets[v + 1] < offsets[v]) return fail(e, GSB_ERR_INVALID_ARGUMENT, "segmentation offsets must be non-decreasing");
		seg[static_cast<size_t>(v)].assign(rows + offsets[v], rows + offsets[v + 1]);
	}
	e->segmentation.swap(seg);
	e->haveSegmentation = true;
	e->prepared = false;
	return GSB_OK;
}

int gsb_get_segmentation(const gsb_engine* e, uint32_t* rows, uint32_t* offsets, int32_t* nvec, int64_t* total) {
	if (!e || !nvec || !total) return GSB_ERR_INVALID_ARGUMENT;
	if (!e->replicas.empty()) return gsb_get_segmentation(e->replicas[0], rows, offsets, nvec, total);
	int64_t tot = 0;
	for (size_t v = 0; v < e->segmentation.size(); ++v) tot += static_cast<int64_t>(e->segmentation[v].size());
	*nvec = static_cast<int32_t>(e->segmentation.size());
	*total = tot;
	if (!rows || !offsets) return GSB_OK;
	uint32_t pos = 0;
	for (size_t v = 0; v < e->segmentation.size(); ++v) {
		offsets[v] = pos;
		std::copy(e->segmentation[v].begin(), e->segmentation[v].end(), rows + pos);
		pos += static_cast<uint32_t>(e->segmentation[v].size());
	}
	offsets[e->segmentation.size()] = pos;
	return GSB_OK;
}

int gsb_prepare(gsb_engine* e) {
	if (!e) return GSB_ERR_INVALID_ARGUMENT;
	if (!e->replicas.empty()) return onReplicas(e, [](gsb_engine* r) { return gsb_prepare(r); }); // K0 on every GPU at once
	if (e->prepared) return GSB_OK;
	return prepareChecked(e);
}

int gsb_set_stream(gsb_engine* e, void* cuda_stream) {
	if (!e) return GSB_ERR_INVALID_ARGUMENT;
	if (!e->replicas.empty()) return fail(e, GSB_ERR_STATE, "a multi-device engine runs on its replicas' own streams");
	try {
		ensureDevice(e);
		GSB_CUDA(cudaStreamSynchronize(e->stream));
		e->callerStream = cuda_stream != nullptr;
		e->stream = e->callerStream ? static_cast<cudaStream_t>(cuda_stream) : e->ownStream;
	} catch (const CudaFailure& f) {
		return failCuda(e, f);
	}
	return GSB_OK;
}

int gsb_get_info(const gsb_engine* e, gsb_info* info) {
	if (!e || !info) return GSB_ERR_INVALID_ARGUMENT;
	if (!e->replicas.empty()) return gsb_get_info(e->replicas[0], info);
	*info = e->info;
	if (!e->prepared) { // host-side facts are available before the device setup
		info->n = e->n;
		info->p = e->p;
		info->max_ncomp = e->shape.maxNComp;
		info->inner_segments = e->shape.inner;
		info->outer_segments = e->shape.outer;
		info->num_replications = e->shape.replications;
		info->segmentation_vectors = static_cast<int32_t>(e->segmentation.size());
		info->sdfact_effective = e->shape.sdfact;
	}
	return GSB_OK;
}

static int uploadPopulation(gsb_engine* e, const uint32_t* idx, const uint32_t* offsets, int32_t npop, bool sync);

int gsb_population_upload(gsb_engine* e, const uint32_t* idx, const uint32_t* offsets, int32_t npop) {
	if (!e) return GSB_ERR_INVALID_ARGUMENT;
	if (!offsets || npop < 0 || (!idx && npop > 0 && offsets[npop] > 0)) return fail(e, GSB_ERR_INVALID_ARGUMENT, "null population");
	if (!e->replicas.empty()) return multiUpload(e, idx, offsets, npop);
	return uploadPopulation(e, idx, offsets, npop, true);
}

// sync = false (gsb_evaluate_indices): the copies are only enqueued; the call's one synchronisation is the download's
static int uploadPopulation(gsb_engine* e, const uint32_t* idx, const uint32_t* offsets, int32_t npop, bool sync) {
	if (!e->prepared) {
		const int rc = prepareChecked(e);
		if (rc != GSB_OK) return rc;
	}
	// validate the whole CSR before any resident state is touched
	const int lmMaxK = std::min(gsb::olsKernelMaxK(e->smemLimit), e->n - 2);
	std::vector<unsigned char> unsortedFlag(static_cast<size_t>(npop), 0);
	for (int c = 0; c < npop; ++c) {
		if (offsets[c + 1] < offsets[c]) return fail(e, GSB_ERR_INVALID_ARGUMENT, "population offsets must be non-decreasing");
		const int k = static_cast<int>(offsets[c + 1] - offsets[c]);
		// largest index and sortedness of the list in one branch-free pass (this loop runs over every index of every call)
		const uint32_t* list = idx + offsets[c];
		uint32_t largest = 0, descents = 0;
		for (int i = 0; i < k; ++i) largest = std::max(largest, list[i]);
		for (int i = 1; i < k; ++i) descents |= static_cast<uint32_t>(list[i] < list[i - 1]);
		if (k > 0 && largest >= static_cast<uint32_t>(e->p)) return fail(e, GSB_ERR_INVALID_ARGUMENT, "column index out of range");
		unsortedFlag[static_cast<size_t>(c)] = static_cast<unsigned char>(descents);
		if (e->cfg.evaluator == GSB_EVAL_LM && k > lmMaxK) return fail(e, GSB_ERR_INVALID_ARGUMENT, "subset too large for the OLS kernel (k <= n - 2)");
		if (e->cfg.evaluator != GSB_EVAL_LM && k > 1024) return fail(e, GSB_ERR_INVALID_ARGUMENT, "subset too large for the SIMPLS kernel (k <= 1024)");
	}
	e->populationReady = e->resultsReady = false;
	try {
		ensureDevice(e);
		if (e->h2dPending) { // a previous un-synchronised upload never reached its download: its copies may still read the pinned buffers
			GSB_CUDA(cudaStreamSynchronize(e->stream));
			e->h2dPending = false;
		}
		const size_t total = npop > 0 ? offsets[npop] : 0;
		e->hIdx.reserve(total + 1);
		e->hOffsets.reserve(static_cast<size_t>(npop) + 1);
		e->hBucket.reserve(static_cast<size_t>(npop) + 1);
		int kMaxPop = 0;
		if (total > 0) std::memcpy(e->hIdx.ptr, idx, total * sizeof(uint32_t));
		for (int c = 0; c < npop; ++c) {
			const uint32_t b = offsets[c], en = offsets[c + 1];
			if (unsortedFlag[static_cast<size_t>(c)]) std::sort(e->hIdx.ptr + b, e->hIdx.ptr + en); // the fit does not depend on the column order
			e->hOffsets.ptr[c] = b;
			kMaxPop = std::max(kMaxPop, static_cast<int>(en - b));
		}
		e->hOffsets.ptr[npop] = static_cast<uint32_t>(total);

		// counting sort of the non-empty chromosomes into subset-size classes (stable: caller order inside a class)
		std::vector<int> order;
		if (kMaxPop > 1024) { // (LM subsets may be wider than the SIMPLS kernels take: no table of classes for those)
			order.reserve(static_cast<size_t>(npop));
			for (int c = 0; c < npop; ++c) if (offsets[c + 1] > offsets[c]) order.push_back(c);
			std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
				return bucketCeil(static_cast<int>(offsets[a + 1] - offsets[a])) < bucketCeil(static_cast<int>(offsets[b + 1] - offsets[b]));
			});
		} else {
			int classCount[80] = {0}; // classes: multiples of 8 up to 64, of 16 up to 256, of 32 up to 1024
			auto classOf = [](int k) { return k <= 64 ? (k + 7) / 8 : (k <= 256 ? 8 + (k - 64 + 15) / 16 : 20 + (k - 256 + 31) / 32); };
			int live = 0;
			for (int c = 0; c < npop; ++c) {
				const int k = static_cast<int>(offsets[c + 1] - offsets[c]);
				if (k > 0) { ++classCount[classOf(k)]; ++live; }
			}
			int start[80];
			for (int i = 0, at = 0; i < 80; ++i) { start[i] = at; at += classCount[i]; }
			order.assign(static_cast<size_t>(live), 0);
			for (int c = 0; c < npop; ++c) {
				const int k = static_cast<int>(offsets[c + 1] - offsets[c]);
				if (k > 0) order[static_cast<size_t>(start[classOf(k)]++)] = c;
			}
		}
		e->buckets.clear();
		for (size_t i = 0; i < order.size(); ++i) {
			const int k = static_cast<int>(offsets[order[i] + 1] - offsets[order[i]]);
			const int cls = bucketCeil(k);
			if (e->buckets.empty() || bucketCeil(e->buckets.back().kMax) != cls) {
				Bucket b = Bucket();
				b.begin = static_cast<int>(i);
				e->buckets.push_back(b);
			}
			e->buckets.back().count += 1;
			e->buckets.back().kMax = std::max(e->buckets.back().kMax, k);
			e->hBucket.ptr[i] = order[i];
		}

		e->npop = npop;
		e->kMaxPop = kMaxPop;
		e->tableA = std::max(1, std::min(e->shape.maxNComp, kMaxPop));
		if (e->cfg.evaluator != GSB_EVAL_LM) planBuckets(e);
		GSB_CUDA(cudaEventRecord(e->ev[0], e->stream));
		e->dIdx.upload(e->hIdx.ptr, total, e->stream);
		e->dOffsets.upload(e->hOffsets.ptr, static_cast<size_t>(npop) + 1, e->stream);
		e->dBucket.upload(e->hBucket.ptr, order.size(), e->stream);
		if (e->cfg.evaluator != GSB_EVAL_LM) {
			e->dPackOff.upload(e->hPackOff.ptr, static_cast<size_t>(npop) + 1, e->stream);
			e->dVecOff.upload(e->hVecOff.ptr, static_cast<size_t>(npop) + 1, e->stream);
			e->dCoefOff.upload(e->hCoefOff.ptr, static_cast<size_t>(npop) + 1, e->stream);
		}
		GSB_CUDA(cudaEventRecord(e->ev[1], e->stream));
		e->dFitness.reserve(static_cast<size_t>(npop));
		e->dStatus.reserve(static_cast<size_t>(npop));
		if (e->cfg.evaluator != GSB_EVAL_LM) {
			const size_t innerDests = static_cast<size_t>(e->plan.groups) * e->plan.innerPerGroup;
			e->dMse.reserve(static_cast<size_t>(npop) * innerDests * e->tableA);
			e->dOstat.reserve(static_cast<size_t>(npop) * e->plan.groups * e->tableA * 2);
		}
		if (sync) {
			GSB_CUDA(cudaStreamSynchronize(e->stream));
			float ms = 0.f;
			GSB_CUDA(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]));
			e->lastH2dMs = ms;
		} else {
			e->h2dPending = true;
		}
		e->populationReady = true;
		e->resultsReady = false;
	} catch (const CudaFailure& f) {
		return failCuda(e, f);
	} catch (const std::string& msg) {
		return fail(e, GSB_ERR_MEMORY, msg);
	} catch (const std::bad_alloc&) {
		return fail(e, GSB_ERR_MEMORY, "host allocation failed");
	}
	return GSB_OK;
}

} // extern "C"

namespace {

// The order the CTAs of the kernel-space kernel run a chromosome's fits in, for `split` CTAs per chromosome and nf fits
// per pass.  With early stopping every CTA takes whole dependency components (replications): their MSEP fits first,
// then their outer-only fits, which the kernel stops at optNComp.  Without it (fewer components than CTAs, no
// outer-only fits, GSB_NO_EARLYSTOP): plan order, equal numbers of passes.  Built once per (split, nf).
gsb_engine::KSchedule& kspaceSchedule(gsb_engine* e, int split, int nf) {
	const int nfits = e->kspaceFits;
	if (e->kSchedNf != nf) {
		for (int i = 0; i <= gsb_engine::kMaxSplit; ++i) e->kSchedules[i].ready = false;
		e->kSchedNf = nf;
	}
	gsb_engine::KSchedule& sc = e->kSchedules[split];
	if (sc.ready) return sc;
	bool anyOuterOnly = false;
	for (int f = 0; f < nfits; ++f) anyOuterOnly = anyOuterOnly || e->kFitPhase[static_cast<size_t>(f)] == 1;
	// whole components per CTA only when they deal out evenly: an uneven deal costs more passes on the longest CTA than
	// stopping the outer fits early saves (measured, profiles/split_probe.py: 10 replications over 4 CTAs)
	sc.early = !e->env.noEarlyStop && anyOuterOnly && e->kComponents >= split && e->kComponents % split == 0 && nfits < (1 << 24);
	std::vector<int> order;
	std::vector<gsb::KPartDesc> parts(static_cast<size_t>(split), gsb::KPartDesc());
	if (sc.early) {
		for (int part = 0; part < split; ++part) {
			gsb::KPartDesc& pd = parts[static_cast<size_t>(part)];
			pd.begin = static_cast<int>(order.size());
			pd.n0 = pd.n1 = pd.pad = 0;
			for (int phase = 0; phase < 2; ++phase) {
				for (int f = 0; f < nfits; ++f) {
					const int comp = e->kFitComp[static_cast<size_t>(f)];
					if (static_cast<long long>(comp) * split / e->kComponents != part || e->kFitPhase[static_cast<size_t>(f)] != phase) continue;
					order.push_back(f);
					(phase == 0 ? pd.n0 : pd.n1) += 1;
				}
			}
		}
	} else {
		const int passes = (nfits + nf - 1) / nf, perPart = (passes + split - 1) / split;
		for (int f = 0; f < nfits; ++f) order.push_back(f);
		for (int part = 0; part < split; ++part) {
			gsb::KPartDesc& pd = parts[static_cast<size_t>(part)];
			pd.begin = std::min(nfits, part * perPart * nf);
			pd.n0 = std::min(nfits, (part + 1) * perPart * nf) - pd.begin;
			pd.n1 = pd.pad = 0;
		}
	}
	// per CTA: one schedule word per fit of its part, and as many again to stage the sort
	sc.stride = 2;
	for (int part = 0; part < split; ++part) sc.stride = std::max(sc.stride, roundUp(parts[static_cast<size_t>(part)].n0 + parts[static_cast<size_t>(part)].n1, 2));
	GSB_CUDA(cudaStreamSynchronize(e->stream)); // no launch in flight reads a previous schedule
	sc.fitOrder.upload(order, e->stream);
	sc.parts.upload(parts, e->stream);
	GSB_CUDA(cudaStreamSynchronize(e->stream)); // the host vectors go out of scope
	sc.ready = true;
	return sc;
}

// K1 per subset-size class, then K2 (or K3 for the LM evaluator); events ev[2..4] bracket them
void enqueuePopulation(gsb_engine* e, long long& ctas, int& launches) {
		e->timing.fit_tensor_gflop = 0.0;
		e->timing.retries = 0;
		e->refitArmed = false;
		GSB_CUDA(cudaEventRecord(e->ev[2], e->stream));
		if (e->cfg.evaluator == GSB_EVAL_LM) {
			// the reference's LMEvaluator scores an empty subset as the intercept-only model (src/LMEvaluator.cpp:27-67)
			gsb::launchLmEmpty(e->dOffsets.ptr, e->npop, e->cfg.statistic, e->n, e->r2denom, e->dFitness.ptr, e->dStatus.ptr, e->stream);
			++launches;
			// size classes whose shared memory stays small go out as ONE launch (a launch of a few dozen single-CTA
			// Choleskys is latency-bound: six of them in a row were 0.4 of the 0.5 ms of a 200-chromosome generation)
			for (size_t b = 0; b < e->buckets.size();) {
				size_t last = b;
				int kMax = e->buckets[b].kMax, count = e->buckets[b].count;
				while (last + 1 < e->buckets.size() && e->buckets[last + 1].kMax <= 96) {
					++last;
					kMax = std::max(kMax, e->buckets[last].kMax);
					count += e->buckets[last].count;
				}
				gsb::OlsParams prm;
				prm.gram = e->gram.ptr;
				prm.s0 = e->s0.ptr;
				prm.xm = e->xm.ptr;
				prm.ym = e->ymAllRows;
				prm.z = e->zpool.ptr;
				prm.ldz = e->ldz;
				prm.n = e->n;
				prm.p = e->p;
				prm.idx = e->dIdx.ptr;
				prm.offsets = e->dOffsets.ptr;
				prm.bucket = e->dBucket.ptr + e->buckets[b].begin;
				prm.bucketCount = count;
				prm.statistic = e->cfg.statistic;
				prm.r2denom = e->r2denom;
				prm.fitness = e->dFitness.ptr;
				prm.status = e->dStatus.ptr;
				// subsets too wide for shared memory (k > ~230): the packed triangle is factorised in a global scratch
				const unsigned long long per = gsb::olsKernelScratchPerCta(kMax, e->smemLimit);
				if (per > 0) e->genScratch.reserve(static_cast<size_t>(per * static_cast<unsigned long long>(std::min(count, 2 * std::max(1, e->smCount)))));
				prm.scratch = e->genScratch.ptr;
				prm.scratchCapacity = e->genScratch.capacity;
				prm.scratchStride = 0;
				GSB_CUDA(gsb::launchOlsKernel(prm, kMax, e->smemLimit, e->stream));
				ctas += prm.bucketCount;
				++launches;
				b = last + 1;
			}
			GSB_CUDA(cudaEventRecord(e->ev[3], e->stream));
			GSB_CUDA(cudaEventRecord(e->ev[4], e->stream));
		} else {
			const int innerDests = e->plan.groups * e->plan.innerPerGroup;
			const EnvFlags& env = e->env;
			// largest subsets first, classes dealt round-robin to the side streams (fork / join on events)
			const int nside = (e->buckets.size() > 1 && !env.serial) ? gsb_engine::kSideStreams : 0;
			if (nside) {
				GSB_CUDA(cudaEventRecord(e->evFork, e->stream));
				for (int i = 0; i < nside; ++i) GSB_CUDA(cudaStreamWaitEvent(e->side[i], e->evFork, 0));
			}
			// K1k takes every size class above kspaceMinK in ONE launch (its cost per fit does not depend on the subset
			// size; below, the per-fit kernel is cheaper): the classes are contiguous in dBucket
			size_t kspaceFirst = e->buckets.size();
			for (size_t b = e->buckets.size(); b-- > 0;) {
				if (e->buckets[b].kernel != BK_KSPACE) break;
				kspaceFirst = b;
			}
			int slot = 0;
			if (kspaceFirst < e->buckets.size()) {
				cudaStream_t launchStream = nside ? e->side[slot % nside] : e->stream;
				++slot;
				const int kspaceNf = (env.kspaceNf == 16 && gsb::kspaceKernelFits(e->n, e->tableA, e->smemLimit, 16)) ? 16 : 8;
				gsb::KParams kp;
				kp.zpool = e->zpool.ptr;
				kp.kfits = e->kFits.ptr;
				kp.ktasks = e->kTasks.ptr;
				kp.idx = e->dIdx.ptr;
				kp.offsets = e->dOffsets.ptr;
				kp.bucket = e->dBucket.ptr + e->buckets[kspaceFirst].begin;
				kp.bucketCount = 0;
				for (size_t b = kspaceFirst; b < e->buckets.size(); ++b) kp.bucketCount += e->buckets[b].count;
				kp.nfits = e->kspaceFits;
				// CTAs per chromosome: every CTA builds its own K, and CTAs run in waves of one per SM.  Whole waves of
				// chromosomes get one CTA each; the chromosomes of the last, partial wave (all of a small population) are
				// split over the number of CTAs that minimises waves x (passes per CTA x pass time + K build time) and
				// come last in the grid.  Measured (k-cycles, B200): a component of a pass ~6, its first scores ~6,
				// K = Z Z' ~0.37 per column.  GSB_KSPACE_SPLIT forces one split for every chromosome.
				const int passes = (e->kspaceFits + kspaceNf - 1) / kspaceNf;
				kp.nf = kspaceNf;
				const int maxSplit = std::min(gsb_engine::kMaxSplit, passes);
				if (env.kspaceSplit > 0) {
					kp.fullCount = 0;
					kp.tailSplit = std::max(1, std::min(env.kspaceSplit, maxSplit));
				} else {
					const int S = std::max(1, e->smCount);
					kp.fullCount = kp.bucketCount / S * S;
					const int tail = kp.bucketCount - kp.fullCount;
					double kbar = 0.0; // mean subset size of the tail (the smallest subsets of the launch)
					for (int i = 0; i < tail; ++i) {
						const int c = e->hBucket.ptr[e->buckets[kspaceFirst].begin + i];
						kbar += static_cast<double>(e->hOffsets.ptr[c + 1] - e->hOffsets.ptr[c]);
					}
					kbar /= std::max(1, tail);
					const double passTime = (6.0 * e->tableA + 6.0) * (kspaceNf / 8), buildTime = 0.37 * kbar;
					double best = 0.0;
					kp.tailSplit = 1;
					for (int sp = 1; sp <= maxSplit && tail > 0; ++sp) {
						const long long waves = (static_cast<long long>(tail) * sp + S - 1) / S;
						// passes of the longest part: whole dependency components (replications) per CTA when there are enough
						int partPasses = (passes + sp - 1) / sp;
						if (e->kComponents >= sp && e->kComponents % sp == 0 && !env.noEarlyStop) {
							const int comps = e->kComponents / sp;
							partPasses = (comps * ((e->kspaceFits + e->kComponents - 1) / e->kComponents) + kspaceNf - 1) / kspaceNf;
						}
						const double t = static_cast<double>(waves) * (partPasses * passTime + buildTime);
						if (sp == 1 || t < best * 0.98) { best = t; kp.tailSplit = sp; }
					}
				}
				kp.n = e->n;
				kp.n8 = roundUp(e->n, 8);
				kp.ldz = e->ldz;
				kp.p = e->p;
				kp.maxNComp = e->tableA;
				kp.innerDests = innerDests;
				kp.outerDests = e->plan.groups;
				kp.mse = e->dMse.ptr;
				kp.ostat = e->dOstat.ptr;
				kp.trace = e->traceBuf.ptr;
				gsb_engine::KSchedule& scTail = kspaceSchedule(e, kp.tailSplit, kp.nf);
				gsb_engine::KSchedule& scFull = kspaceSchedule(e, 1, kp.nf);
				kp.fitOrder = scTail.fitOrder.ptr;
				kp.parts = scTail.parts.ptr;
				kp.fitOrderFull = scFull.fitOrder.ptr;
				kp.partsFull = scFull.parts.ptr;
				kp.inner = std::max(1, e->plan.innerPerGroup);
				kp.sdfact = e->shape.sdfact;
				kp.schedStride = std::max(scTail.stride, scFull.stride);
				e->kSched.reserve((static_cast<size_t>(kp.fullCount) + static_cast<size_t>(kp.bucketCount - kp.fullCount) * kp.tailSplit) * 2 * static_cast<size_t>(kp.schedStride));
				kp.sched = e->kSched.ptr;
				kp.products = e->kProducts.ptr;
				e->dRefit.reserve(static_cast<size_t>(e->npop));
				kp.refit = e->dRefit.ptr;
				if (g_guard && !e->dGuardFail.ptr) { e->dGuardFail.reserve(1); GSB_CUDA(cudaMemset(e->dGuardFail.ptr, 0, sizeof(int))); }
				kp.guardFail = e->dGuardFail.ptr;
				e->refitArmed = true;
				GSB_CUDA(cudaMemsetAsync(e->dRefit.ptr, 0, static_cast<size_t>(e->npop) * sizeof(int), launchStream));
				GSB_CUDA(cudaMemsetAsync(e->kProducts.ptr, 0, sizeof(unsigned long long), launchStream));
				GSB_CUDA(gsb::launchKspaceKernel(kp, launchStream, &ctas));
				++launches;
				e->kProductsPending = true;
				{ // tensor-core work of the K = Z Z' builds (DMMA.8x8x4 = 512 FLOP); the products are counted by the kernel
					const double nb = (kp.n8 / 8 + 1) / 2;
					double dmma = 0.0;
					for (int i = 0; i < kp.bucketCount; ++i) {
						const int c = e->hBucket.ptr[e->buckets[kspaceFirst].begin + i];
						const int k = static_cast<int>(e->hOffsets.ptr[c + 1] - e->hOffsets.ptr[c]);
						dmma += static_cast<double>(i < kp.bucketCount - kp.fullCount ? kp.tailSplit : 1) * (nb * (nb + 1) / 2) * 4.0 * ((k + 3) / 4);
					}
					e->timing.fit_tensor_gflop = dmma * 512.0 * 1e-9;
				}
			}
			for (size_t bi = kspaceFirst; bi-- > 0; ++slot) {
				const Bucket& bk = e->buckets[bi];
				cudaStream_t launchStream = nside ? e->side[slot % nside] : e->stream;
				if (nside && bk.kernel == BK_FIT && bk.genPerCta > 0) launchStream = e->side[0]; // the classes that share the global scratch run in order
				if (bk.kernel == BK_SCORES) {
					gsb::XParams xp;
					xp.zpool = e->xz.ptr;
					xp.reps = e->xReps.ptr;
					xp.groups = bk.nf == 16 ? e->xGroups16.ptr : e->xGroups8.ptr;
					xp.xfits = e->xFits.ptr;
					xp.xtasks = e->xTasks.ptr;
					xp.idx = e->dIdx.ptr;
					xp.offsets = e->dOffsets.ptr;
					xp.bucket = e->dBucket.ptr + bk.begin;
					xp.bucketCount = bk.count;
					xp.ngroups = bk.nf == 16 ? e->scoreGroups16 : e->scoreGroups8;
					xp.nr = e->scoreNr;
					xp.nf = bk.nf;
					xp.trace = e->traceBuf.ptr;
					xp.vscratch = bk.vg ? e->scoreScratch.ptr + bk.vOffset : nullptr;
					xp.vstride = bk.vPerCta;
					xp.vslots = (bk.vg && bk.vSlots < 0) ? -bk.vSlots : 0;
					xp.desync = env.desync;
					xp.p = e->p;
					xp.maxNComp = e->tableA;
					xp.innerDests = innerDests;
					xp.outerDests = e->plan.groups;
					xp.mse = e->dMse.ptr;
					xp.ostat = e->dOstat.ptr;
					GSB_CUDA(gsb::launchScoreKernel(xp, bk.kMax, e->tableA, bk.cs, bk.vg, e->smemLimit, launchStream, &ctas));
					++launches;
					continue;
				}
				// per-fit Gram-space kernel: first the packed Grams and vectors of every (chromosome, training set) of the
				// class, then one CTA per (chromosome, training set), on the same stream
				gsb::PackParams pp;
				pp.unitTab = e->unitTab.ptr;
				pp.unitSum = e->unitSum.ptr;
				pp.unitXy = e->unitXy.ptr;
				pp.pfits = e->packFits.ptr;
				pp.idx = e->dIdx.ptr;
				pp.offsets = e->dOffsets.ptr;
				pp.bucket = e->dBucket.ptr + bk.begin;
				pp.bucketCount = bk.count;
				pp.nfits = static_cast<int>(e->plan.fits.size());
				pp.upad = e->upad;
				pp.gpack = e->gpack.ptr;
				pp.pvec = e->pvec.ptr;
				pp.packOff = e->dPackOff.ptr;
				pp.vecOff = e->dVecOff.ptr;
				GSB_CUDA(gsb::launchGramPack(pp, bk.kMax, launchStream));
				launches += 2;
				gsb::FitParams prm;
				prm.gpack = e->gpack.ptr;
				prm.pvec = e->pvec.ptr;
				prm.packOff = e->dPackOff.ptr;
				prm.vecOff = e->dVecOff.ptr;
				prm.zpool = e->zpool.ptr;
				prm.fits = e->fits.ptr;
				prm.tasks = e->tasks.ptr;
				prm.blocks = e->blocks.ptr;
				prm.idx = e->dIdx.ptr;
				prm.offsets = e->dOffsets.ptr;
				prm.bucket = e->dBucket.ptr + bk.begin;
				prm.bucketCount = bk.count;
				prm.nfits = static_cast<int>(e->plan.fits.size());
				prm.p = e->p;
				prm.maxNComp = e->tableA;
				prm.innerDests = innerDests;
				prm.outerDests = e->plan.groups;
				prm.mse = e->dMse.ptr;
				prm.ostat = e->dOstat.ptr;
				prm.coefOut = nullptr;
				prm.trace = e->traceBuf.ptr;
				prm.maxTasks = e->maxTasks;
				prm.maxBlockLd = e->maxBlockLd;
				prm.coefScratch = e->coefScratch.ptr;
				prm.coefOff = e->dCoefOff.ptr;
				prm.genScratch = e->genScratch.ptr;
				prm.genCapacity = e->genScratch.capacity;
				prm.genStride = 0;
				if (g_guard && !e->dGuardFail.ptr) { e->dGuardFail.reserve(1); GSB_CUDA(cudaMemset(e->dGuardFail.ptr, 0, sizeof(int))); }
				prm.guardFail = e->dGuardFail.ptr;
				if (bk.split) {
					GSB_CUDA(gsb::launchFitOnlyKernel(prm, bk.kMax, e->tableA, e->smemLimit, launchStream, &ctas));
					gsb::PredictParams pr;
					pr.zpool = e->zpool.ptr;
					pr.blocks = e->blocks.ptr;
					pr.activeBlocks = e->activeBlocks.ptr;
					pr.nActive = e->nActiveBlocks;
					pr.blockTaskOff = e->blockTaskOff.ptr;
					pr.ptasks = e->predTasks.ptr;
					pr.coef = e->coefScratch.ptr;
					pr.coefOff = e->dCoefOff.ptr;
					pr.idx = e->dIdx.ptr;
					pr.offsets = e->dOffsets.ptr;
					pr.bucket = e->dBucket.ptr + bk.begin;
					pr.bucketCount = bk.count;
					pr.p = e->p;
					pr.maxNComp = e->tableA;
					pr.innerDests = innerDests;
					pr.outerDests = e->plan.groups;
					pr.lda = gsb::predictKernelLda(e->maxBlockLd);
					pr.mse = e->dMse.ptr;
					pr.ostat = e->dOstat.ptr;
					GSB_CUDA(gsb::launchPredictKernel(pr, launchStream, nullptr)); // fit_ctas counts the fits' CTAs
					launches += 2;
				} else {
					GSB_CUDA(gsb::launchFitKernel(prm, bk.kMax, e->tableA, e->smemLimit, launchStream, &ctas));
					++launches;
				}
			}
			for (int i = 0; i < nside; ++i) {
				GSB_CUDA(cudaEventRecord(e->evJoin[i], e->side[i]));
				GSB_CUDA(cudaStreamWaitEvent(e->stream, e->evJoin[i], 0));
			}
			GSB_CUDA(cudaEventRecord(e->ev[3], e->stream));
			gsb::ReduceParams rp;
			rp.mse = e->dMse.ptr;
			rp.ostat = e->dOstat.ptr;
			rp.offsets = e->dOffsets.ptr;
			rp.outerRows = e->outerRows.ptr;
			rp.npop = e->npop;
			rp.groups = e->plan.groups;
			rp.inner = e->plan.innerPerGroup;
			rp.outerPerRep = e->plan.groupsPerReplication;
			rp.replications = e->shape.replications;
			rp.maxNComp = e->tableA;
			rp.innerDests = innerDests;
			rp.outerDests = e->plan.groups;
			rp.kind = e->cfg.evaluator;
			rp.statistic = e->cfg.statistic;
			rp.n = e->n;
			rp.sdfact = e->shape.sdfact;
			rp.gamma = e->cfg.evaluator == GSB_EVAL_PLS_CV ? e->cfg.complexity_penalty : 0.0;
			rp.r2denom = e->r2denom;
			rp.tieBand = e->cfg.tie_band == 0.0 ? 1e-9 : e->cfg.tie_band;
			rp.refit = e->refitArmed ? e->dRefit.ptr : nullptr;
			rp.fitness = e->dFitness.ptr;
			rp.status = e->dStatus.ptr;
			gsb::launchReduceKernel(rp, e->stream);
			GSB_CUDA(cudaGetLastError());
			++launches;
			GSB_CUDA(cudaEventRecord(e->ev[4], e->stream));
		}
}

// GSB_GUARD: every device array's guard zone and the kernels' shared-memory canaries; "" or what was overwritten
std::string checkGuards(gsb_engine* e) {
	if (!g_guard) return "";
	std::string bad;
#define GSB_CHECK_GUARD(member) do { if (!e->member.guardIntact()) bad += std::string(bad.empty() ? "" : ", ") + #member; } while (0)
	GSB_CHECK_GUARD(zpool); GSB_CHECK_GUARD(gram); GSB_CHECK_GUARD(s0); GSB_CHECK_GUARD(xm); GSB_CHECK_GUARD(fits); GSB_CHECK_GUARD(tasks);
	GSB_CHECK_GUARD(blocks); GSB_CHECK_GUARD(outerRows); GSB_CHECK_GUARD(unitTab); GSB_CHECK_GUARD(unitSum); GSB_CHECK_GUARD(unitXy);
	GSB_CHECK_GUARD(packFits); GSB_CHECK_GUARD(gpack); GSB_CHECK_GUARD(pvec); GSB_CHECK_GUARD(dPackOff); GSB_CHECK_GUARD(dVecOff);
	GSB_CHECK_GUARD(dCoefOff); GSB_CHECK_GUARD(coefScratch); GSB_CHECK_GUARD(genScratch); GSB_CHECK_GUARD(activeBlocks);
	GSB_CHECK_GUARD(blockTaskOff); GSB_CHECK_GUARD(predTasks); GSB_CHECK_GUARD(dIdx); GSB_CHECK_GUARD(dOffsets); GSB_CHECK_GUARD(dBucket);
	GSB_CHECK_GUARD(dStatus); GSB_CHECK_GUARD(dMse); GSB_CHECK_GUARD(dOstat); GSB_CHECK_GUARD(dFitness); GSB_CHECK_GUARD(xz);
	GSB_CHECK_GUARD(scoreScratch); GSB_CHECK_GUARD(xReps); GSB_CHECK_GUARD(xGroups8); GSB_CHECK_GUARD(xGroups16); GSB_CHECK_GUARD(xFits);
	GSB_CHECK_GUARD(xTasks); GSB_CHECK_GUARD(kFits); GSB_CHECK_GUARD(kTasks); GSB_CHECK_GUARD(kSched); GSB_CHECK_GUARD(kProducts);
	GSB_CHECK_GUARD(dRefit); GSB_CHECK_GUARD(traceBuf);
	for (int i = 0; i <= gsb_engine::kMaxSplit; ++i) { GSB_CHECK_GUARD(kSchedules[i].fitOrder); GSB_CHECK_GUARD(kSchedules[i].parts); }
#undef GSB_CHECK_GUARD
	if (e->dGuardFail.ptr) {
		int fails = 0;
		if (cudaMemcpy(&fails, e->dGuardFail.ptr, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess && fails != 0) {
			bad += std::string(bad.empty() ? "" : ", ") + "shared-memory canary of a K1 kernel (" + std::to_string(fails) + " CTAs)";
		}
	}
	return bad;
}

// Chromosomes the kernel-space kernel declined (kStatusRetry in hStatus[0 .. npop)): scored again by a child engine that
// has no K1k (the Gram-space kernel takes its norms from explicit vectors, like the reference); host AND device copies
// of the results are patched.  Returns GSB_OK or the child's error.
int resolveRetries(gsb_engine* e) {
	std::vector<int> ids;
	for (int c = 0; c < e->npop; ++c) if (e->hStatus.ptr[c] == gsb::kStatusRetry) ids.push_back(c);
	if (ids.empty()) return GSB_OK;
	if (!e->retryEngine) {
		gsb_config one = e->cfg;
		one.num_devices = 0;
		gsb_engine* r = nullptr;
		int rc = gsb_create(&r, &one);
		if (rc != GSB_OK) return fail(e, rc, "retry engine: " + g_createError);
		r->forceNoKspace = true;
		rc = gsb_set_data(r, e->X.data(), e->y.data(), e->n, e->p);
		if (rc == GSB_OK) { // adopt the parent's segmentation as it is
			std::vector<uint32_t> rows, offs(1, 0u);
			for (size_t v = 0; v < e->segmentation.size(); ++v) {
				rows.insert(rows.end(), e->segmentation[v].begin(), e->segmentation[v].end());
				offs.push_back(static_cast<uint32_t>(rows.size()));
			}
			rc = gsb_set_segmentation(r, rows.data(), offs.data(), static_cast<int32_t>(e->segmentation.size()));
		}
		if (rc != GSB_OK) {
			const std::string msg = "retry engine: " + r->error;
			gsb_destroy(r);
			return fail(e, rc, msg);
		}
		e->retryEngine = r;
	}
	std::vector<uint32_t> idx, off(1, 0u);
	for (size_t i = 0; i < ids.size(); ++i) {
		const int c = ids[i];
		idx.insert(idx.end(), e->hIdx.ptr + e->hOffsets.ptr[c], e->hIdx.ptr + e->hOffsets.ptr[c + 1]);
		off.push_back(static_cast<uint32_t>(idx.size()));
	}
	if (idx.empty()) idx.push_back(0u);
	std::vector<double> fit(ids.size(), 0.0);
	std::vector<int32_t> st(ids.size(), 0);
	const int rc = gsb_evaluate_indices(e->retryEngine, idx.data(), off.data(), static_cast<int32_t>(ids.size()), fit.data(), st.data());
	if (rc != GSB_OK) return fail(e, rc, "retry engine: " + e->retryEngine->error);
	GSB_CUDA(cudaSetDevice(e->cfg.device));
	for (size_t i = 0; i < ids.size(); ++i) {
		const int c = ids[i];
		e->hFitness.ptr[c] = fit[i];
		e->hStatus.ptr[c] = st[i];
		GSB_CUDA(cudaMemcpyAsync(e->dFitness.ptr + c, e->hFitness.ptr + c, sizeof(double), cudaMemcpyHostToDevice, e->stream));
		GSB_CUDA(cudaMemcpyAsync(e->dStatus.ptr + c, e->hStatus.ptr + c, sizeof(int), cudaMemcpyHostToDevice, e->stream));
	}
	GSB_CUDA(cudaStreamSynchronize(e->stream));
	e->timing.retries = static_cast<int32_t>(ids.size());
	return GSB_OK;
}

} // namespace

extern "C" {

int gsb_population_enqueue(gsb_engine* e) {
	if (!e) return GSB_ERR_INVALID_ARGUMENT;
	if (!e->replicas.empty()) return multiEvaluate(e, false);
	if (!e->populationReady) return fail(e, GSB_ERR_STATE, "no resident population: call gsb_population_upload first");
	try {
		ensureDevice(e);
		long long ctas = 0;
		int launches = 0;
		enqueuePopulation(e, ctas, launches);
		e->timing.fit_ctas = ctas;
		e->timing.kernel_launches = launches;
		e->resultsReady = true;
	} catch (const CudaFailure& f) {
		return failCuda(e, f);
	}
	return GSB_OK;
}

int gsb_population_evaluate(gsb_engine* e) {
	if (!e) return GSB_ERR_INVALID_ARGUMENT;
	if (!e->replicas.empty()) return multiEvaluate(e, true);
	if (!e->populationReady) return fail(e, GSB_ERR_STATE, "no resident population: call gsb_population_upload first");
	try {
		ensureDevice(e);
		long long ctas = 0;
		int launches = 0;
		e->kProductsPending = false;
		enqueuePopulation(e, ctas, launches);
		unsigned long long products = 0ull;
		if (e->kProductsPending) GSB_CUDA(cudaMemcpyAsync(&products, e->kProducts.ptr, sizeof(products), cudaMemcpyDeviceToHost, e->stream));
		GSB_CUDA(cudaStreamSynchronize(e->stream));
		if (e->kProductsPending) { // products the kernel-space kernel issued (column tiles of 8 fits): n8/8 row tiles x n8/4 steps of DMMA.8x8x4
			const double n8 = roundUp(e->n, 8);
			e->timing.fit_tensor_gflop += static_cast<double>(products) * (n8 / 8.0) * (n8 / 4.0) * 512.0 * 1e-9;
		}
		float msFit = 0.f, msRed = 0.f;
		GSB_CUDA(cudaEventElapsedTime(&msFit, e->ev[2], e->ev[3]));
		GSB_CUDA(cudaEventElapsedTime(&msRed, e->ev[3], e->ev[4]));
		e->timing.h2d_ms = e->lastH2dMs;
		e->timing.fit_kernel_ms = msFit;
		e->timing.reduce_kernel_ms = msRed;
		e->timing.d2h_ms = 0.0;
		e->timing.total_ms = e->lastH2dMs + msFit + msRed;
		e->timing.fit_ctas = ctas;
		e->timing.kernel_launches = launches;
		e->resultsReady = true;
		const std::string overwritten = checkGuards(e);
		if (!overwritten.empty()) return fail(e, GSB_ERR_CUDA, "GSB_GUARD: out-of-bounds write behind " + overwritten);
	} catch (const CudaFailure& f) {
		return failCuda(e, f);
	}
	return GSB_OK;
}

static int downloadResults(gsb_engine* e, double* fitness, int32_t* status, bool fused);

int gsb_population_download(gsb_engine* e, double* fitness, int32_t* status) {
	if (!e) return GSB_ERR_INVALID_ARGUMENT;
	if (!fitness || !status) return fail(e, GSB_ERR_INVALID_ARGUMENT, "null output buffer");
	if (!e->replicas.empty()) return multiDownload(e, fitness, status);
	if (!e->resultsReady) return fail(e, GSB_ERR_STATE, "no results: call gsb_population_evaluate first");
	return downloadResults(e, fitness, status, false);
}

// fused = true (gsb_evaluate_indices): upload, kernels and these copies were enqueued back to back; this is the call's ONE
// synchronisation and every stage is timed from its events afterwards
static int downloadResults(gsb_engine* e, double* fitness, int32_t* status, bool fused) {
	try {
		ensureDevice(e);
		const size_t np = static_cast<size_t>(e->npop);
		e->hFitness.reserve(np + 1);
		e->hStatus.reserve(np + 1);
		if (!fused) GSB_CUDA(cudaEventRecord(e->ev[4], e->stream)); // fused: ev[4] is the end of K2, where these copies start
		GSB_CUDA(cudaMemcpyAsync(e->hFitness.ptr, e->dFitness.ptr, np * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
		GSB_CUDA(cudaMemcpyAsync(e->hStatus.ptr, e->dStatus.ptr, np * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
		GSB_CUDA(cudaEventRecord(e->ev[5], e->stream));
		unsigned long long products = 0ull;
		if (fused && e->kProductsPending) GSB_CUDA(cudaMemcpyAsync(&products, e->kProducts.ptr, sizeof(products), cudaMemcpyDeviceToHost, e->stream));
		GSB_CUDA(cudaStreamSynchronize(e->stream));
		e->h2dPending = false;
		float ms = 0.f;
		if (fused) {
			float msH2d = 0.f, msFit = 0.f, msRed = 0.f;
			GSB_CUDA(cudaEventElapsedTime(&msH2d, e->ev[0], e->ev[1]));
			GSB_CUDA(cudaEventElapsedTime(&msFit, e->ev[2], e->ev[3]));
			GSB_CUDA(cudaEventElapsedTime(&msRed, e->ev[3], e->ev[4]));
			e->lastH2dMs = msH2d;
			e->timing.h2d_ms = msH2d;
			e->timing.fit_kernel_ms = msFit;
			e->timing.reduce_kernel_ms = msRed;
			if (e->kProductsPending) {
				const double n8 = roundUp(e->n, 8);
				e->timing.fit_tensor_gflop += static_cast<double>(products) * (n8 / 8.0) * (n8 / 4.0) * 512.0 * 1e-9;
			}
		}
		GSB_CUDA(cudaEventElapsedTime(&ms, e->ev[4], e->ev[5]));
		e->timing.d2h_ms = ms;
		e->timing.total_ms = e->timing.h2d_ms + e->timing.fit_kernel_ms + e->timing.reduce_kernel_ms + ms;
		if (e->refitArmed) {
			const int rc = resolveRetries(e);
			if (rc != GSB_OK) return rc;
		}
		const std::string overwritten = checkGuards(e);
		if (!overwritten.empty()) return fail(e, GSB_ERR_CUDA, "GSB_GUARD: out-of-bounds write behind " + overwritten);
		std::memcpy(fitness, e->hFitness.ptr, np * sizeof(double));
		for (size_t i = 0; i < np; ++i) status[i] = e->hStatus.ptr[i];
	} catch (const CudaFailure& f) {
		return failCuda(e, f);
	}
	return GSB_OK;
}

int gsb_population_device_results(gsb_engine* e, void** d_fitness, void** d_status, int32_t* npop) {
	if (!e || !d_fitness || !d_status || !npop) return GSB_ERR_INVALID_ARGUMENT;
	if (!e->replicas.empty()) return fail(e, GSB_ERR_STATE, "a multi-device engine gathers its results to the host (gsb_population_download)");
	if (!e->populationReady) return fail(e, GSB_ERR_STATE, "no resident population");
	*d_fitness = e->dFitness.ptr;
	*d_status = e->dStatus.ptr;
	*npop = e->npop;
	return GSB_OK;
}

int gsb_evaluate_indices(gsb_engine* e, const uint32_t* idx, const uint32_t* offsets, int32_t npop, double* fitness,
                         int32_t* status) {
	if (e && e->replicas.empty() && offsets && npop > 0 && idx && fitness && status) {
		// one device: upload, kernels and download enqueued back to back, ONE synchronisation (three, and their event
		// queries, cost 0.05 ms of a 2 ms call)
		int rc = uploadPopulation(e, idx, offsets, npop, false);
		if (rc != GSB_OK) return rc;
		try {
			long long ctas = 0;
			int launches = 0;
			e->kProductsPending = false;
			enqueuePopulation(e, ctas, launches);
			e->timing.fit_ctas = ctas;
			e->timing.kernel_launches = launches;
			e->resultsReady = true;
		} catch (const CudaFailure& f) {
			return failCuda(e, f);
		}
		return downloadResults(e, fitness, status, true);
	}
	if (e && !e->replicas.empty() && offsets && npop > 0 && idx && fitness && status) return multiEvaluateIndices(e, idx, offsets, npop, fitness, status);
	int rc = gsb_population_upload(e, idx, offsets, npop);
	if (rc != GSB_OK) return rc;
	if (npop == 0) return GSB_OK;
	rc = gsb_population_evaluate(e);
	if (rc != GSB_OK) return rc;
	return gsb_population_download(e, fitness, status);
}

int gsb_evaluate_bitsets(gsb_engine* e, const uint64_t* bitsets, int32_t words_per_chrom, int32_t npop, double* fitness,
                         int32_t* status) {
	if (!e) return GSB_ERR_INVALID_ARGUMENT;
	if (!e->haveData) return fail(e, GSB_ERR_STATE, "gsb_set_data has not been called");
	if (!bitsets || npop < 0) return fail(e, GSB_ERR_INVALID_ARGUMENT, "null population");
	const int words = (e->p + 63) / 64;
	if (words_per_chrom != words) return fail(e, GSB_ERR_INVALID_ARGUMENT, "words_per_chrom must be ceil(p / 64)");
	// Chromosome::toColumnSubset (src/Chromosome.cpp:419-438): variable v is bit unusedBits + v
	const int unused = words * 64 - e->p;
	std::vector<uint32_t> idx, offsets(static_cast<size_t>(npop) + 1, 0u);
	try {
		size_t total = 0;
		for (size_t w = 0; w < static_cast<size_t>(npop) * words; ++w) total += static_cast<size_t>(__builtin_popcountll(bitsets[w]));
		idx.reserve(total + 1);
		for (int c = 0; c < npop; ++c) {
			const uint64_t* w = bitsets + static_cast<size_t>(c) * words;
			for (int part = 0; part < words; ++part) {
				uint64_t bits = w[part];
				if (part == 0 && unused > 0) bits &= ~((uint64_t(1) << unused) - 1);
				while (bits) {
					const int bit = __builtin_ctzll(bits);
					idx.push_back(static_cast<uint32_t>(part * 64 + bit - unused));
					bits &= bits - 1;
				}
			}
			offsets[static_cast<size_t>(c) + 1] = static_cast<uint32_t>(idx.size());
		}
	} catch (const std::bad_alloc&) {
		return fail(e, GSB_ERR_MEMORY, "host allocation failed");
	}
	if (idx.empty()) idx.push_back(0u);
	return gsb_evaluate_indices(e, idx.data(), offsets.data(), npop, fitness, status);
}

int gsb_debug_trace(gsb_engine* e, int64_t* out, int32_t n) {
	if (!e || !out || n < 0) return GSB_ERR_INVALID_ARGUMENT;
	if (!e->replicas.empty()) return gsb_debug_trace(e->replicas[0], out, n);
	for (int i = 0; i < n; ++i) out[i] = 0;
	if (!e->traceBuf.ptr) return GSB_OK;
	try {
		ensureDevice(e);
		GSB_CUDA(cudaStreamSynchronize(e->stream));
		GSB_CUDA(cudaMemcpy(out, e->traceBuf.ptr, sizeof(long long) * static_cast<size_t>(n < 64 ? n : 64), cudaMemcpyDeviceToHost));
	} catch (const CudaFailure& f) {
		return failCuda(e, f);
	}
	return GSB_OK;
}

int gsb_deal(const uint32_t* offsets, int32_t npop, int32_t ndev, int32_t kspace, int32_t* device_of) {
	if (!offsets || !device_of || npop < 0 || ndev < 1 || ndev > GSB_MAX_DEVICES) return GSB_ERR_INVALID_ARGUMENT;
	for (int c = 0; c < npop; ++c) if (offsets[c + 1] < offsets[c]) return GSB_ERR_INVALID_ARGUMENT;
	std::vector<int> d;
	dealPopulation(offsets, npop, ndev, kspace != 0, d);
	for (int c = 0; c < npop; ++c) device_of[c] = d[static_cast<size_t>(c)];
	return GSB_OK;
}

int gsb_last_timing(const gsb_engine* e, gsb_timing* t) {
	if (!e || !t) return GSB_ERR_INVALID_ARGUMENT;
	*t = e->timing;
	return GSB_OK;
}

int gsb_simpls(gsb_engine* e, const uint32_t* cols, int32_t k, const uint32_t* rows, int32_t nrows, int32_t ncomp,
               double* coef, double* intercepts) {
	if (!e) return GSB_ERR_INVALID_ARGUMENT;
	if (!e->haveData) return fail(e, GSB_ERR_STATE, "gsb_set_data has not been called");
	if (!e->replicas.empty()) {
		const int rc = gsb_simpls(e->replicas[0], cols, k, rows, nrows, ncomp, coef, intercepts);
		return rc == GSB_OK ? rc : failFrom(e, rc, e->replicas[0], e->cfg.devices[0]);
	}
	if (!cols || k < 1 || k > 1024 || ncomp < 1 || !coef || !intercepts) return fail(e, GSB_ERR_INVALID_ARGUMENT, "invalid SIMPLS request");
	const int n = e->n, p = e->p;
	std::vector<uint32_t> trainRows;
	if (rows && nrows > 0) trainRows.assign(rows, rows + nrows);
	else { trainRows.resize(static_cast<size_t>(n)); std::iota(trainRows.begin(), trainRows.end(), 0u); }
	const int ntr = static_cast<int>(trainRows.size());
	for (int i = 0; i < ntr; ++i) if (trainRows[static_cast<size_t>(i)] >= static_cast<uint32_t>(n)) return fail(e, GSB_ERR_INVALID_ARGUMENT, "row index out of range");
	for (int i = 0; i < k; ++i) {
		if (cols[i] >= static_cast<uint32_t>(p)) return fail(e, GSB_ERR_INVALID_ARGUMENT, "column index out of range");
		if (i > 0 && cols[i] <= cols[i - 1]) return fail(e, GSB_ERR_INVALID_ARGUMENT, "columns must be strictly ascending");
	}
	// PLSSimpls::fit clamps the number of components (src/PLSSimpls.cpp:107-110)
	const int A = std::min(ncomp, std::min(k, ntr - 1));
	if (A < 1) return fail(e, GSB_ERR_INVALID_ARGUMENT, "too few training rows");
	try {
		ensureDevice(e);
		if (!e->zReady) uploadZ(e, 0);
		const int K2 = k + 2, ld = roundUp(ntr, 4), ldm = roundUp(K2, 4);
		std::vector<uint32_t> colSel(cols, cols + k);
		colSel.push_back(static_cast<uint32_t>(p));
		colSel.push_back(static_cast<uint32_t>(p + 1));
		DeviceArray<uint32_t> dCols, dRows, dIdx, dOff;
		DeviceArray<double> dSub, dM, dGram, dS0, dXm, dYm, dOut, dDummy;
		DeviceArray<const double*> dUnit, dMBlocks;
		DeviceArray<double*> dOutPtr;
		DeviceArray<int> dLd, dNr, dExclIdx, dExclCnt, dFitId, dFitNtr, dBucket;
		DeviceArray<unsigned long long> dGOff;
		DeviceArray<gsb::FitDesc> dFit;
		dCols.upload(colSel, e->stream);
		dRows.upload(trainRows, e->stream);
		dSub.reserve(static_cast<size_t>(ld) * K2);
		gsb::launchBlockGather(e->zpool.ptr, e->ldz, K2, dCols.ptr, dRows.ptr, ntr, dSub.ptr, ld, e->stream);
		dM.reserve(static_cast<size_t>(K2) * ldm);
		GSB_CUDA(cudaMemsetAsync(dM.ptr, 0, static_cast<size_t>(K2) * ldm * sizeof(double), e->stream));
		const double* unitPtr = dSub.ptr;
		double* outPtr = dM.ptr;
		const double* noBlock = nullptr;
		const int zero = 0, two[2] = {0, 0};
		const unsigned long long goff = 0;
		dUnit.upload(&unitPtr, 1, e->stream);
		dOutPtr.upload(&outPtr, 1, e->stream);
		dLd.upload(&ld, 1, e->stream);
		dNr.upload(&ld, 1, e->stream);
		dMBlocks.upload(&noBlock, 1, e->stream);
		dExclIdx.upload(two, 2, e->stream);
		dExclCnt.upload(&zero, 1, e->stream);
		dFitId.upload(&zero, 1, e->stream);
		dFitNtr.upload(&ntr, 1, e->stream);
		dGOff.upload(&goff, 1, e->stream);
		gsb::launchGramSyrk(dUnit.ptr, dLd.ptr, dNr.ptr, dOutPtr.ptr, 1, K2, ldm, e->stream);
		dGram.reserve(static_cast<size_t>(gsb::packedDoubles(static_cast<unsigned long long>(k))));
		GSB_CUDA(cudaMemsetAsync(dGram.ptr, 0, static_cast<size_t>(gsb::packedDoubles(static_cast<unsigned long long>(k))) * sizeof(double), e->stream));
		dS0.reserve(static_cast<size_t>(k));
		dXm.reserve(static_cast<size_t>(k));
		dYm.reserve(1);
		gsb::launchGramCombine(dM.ptr, dMBlocks.ptr, dExclIdx.ptr, dExclCnt.ptr, dFitId.ptr, dFitNtr.ptr, 1, k, ldm, dGram.ptr,
		                       dGOff.ptr, dS0.ptr, dXm.ptr, dYm.ptr, e->stream);
		GSB_CUDA(cudaGetLastError());
		double ymc = 0.0;
		GSB_CUDA(cudaMemcpyAsync(&ymc, dYm.ptr, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
		GSB_CUDA(cudaStreamSynchronize(e->stream));
		gsb::FitDesc fd;
		fd.ym = ymc; fd.ntr = ntr; fd.taskBegin = 0; fd.taskCount = 0; fd.pad = 0;
		std::vector<uint32_t> ident(static_cast<size_t>(k));
		std::iota(ident.begin(), ident.end(), 0u);
		const uint32_t off[2] = {0u, static_cast<uint32_t>(k)};
		dFit.upload(&fd, 1, e->stream);
		dIdx.upload(ident, e->stream);
		dOff.upload(off, 2, e->stream);
		dBucket.upload(&zero, 1, e->stream);
		dOut.reserve(static_cast<size_t>(k) * A + A);
		dDummy.reserve(16);
		// the kernel's slab layout: s0 then xm, each padded to an even count (the packed Gram is padded in place)
		const int kp = roundUp(k, 2);
		DeviceArray<double> dVec;
		DeviceArray<unsigned long long> dZeroOff;
		dVec.reserve(static_cast<size_t>(2 * kp));
		GSB_CUDA(cudaMemsetAsync(dVec.ptr, 0, static_cast<size_t>(2 * kp) * sizeof(double), e->stream));
		GSB_CUDA(cudaMemcpyAsync(dVec.ptr, dS0.ptr, static_cast<size_t>(k) * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
		GSB_CUDA(cudaMemcpyAsync(dVec.ptr + kp, dXm.ptr, static_cast<size_t>(k) * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
		dZeroOff.upload(&goff, 1, e->stream);
		gsb::FitParams prm;
		prm.gpack = dGram.ptr; prm.pvec = dVec.ptr; prm.packOff = dZeroOff.ptr; prm.vecOff = dZeroOff.ptr; prm.zpool = dSub.ptr;
		prm.fits = dFit.ptr; prm.tasks = nullptr; prm.blocks = nullptr;
		prm.idx = dIdx.ptr; prm.offsets = dOff.ptr; prm.bucket = dBucket.ptr; prm.bucketCount = 1; prm.nfits = 1;
		prm.p = k; prm.maxNComp = A; prm.innerDests = 0; prm.outerDests = 0;
		prm.mse = dDummy.ptr; prm.ostat = dDummy.ptr; prm.coefOut = dOut.ptr; prm.trace = nullptr; prm.maxTasks = 0; prm.maxBlockLd = 0;
		prm.coefScratch = nullptr; prm.coefOff = nullptr;
		DeviceArray<double> dGen; // many components on a wide subset: stored loadings and coefficients outside shared memory
		const unsigned long long genPer = gsb::fitKernelScratchPerCta(k, A, 0, 0, e->smemLimit);
		if (genPer > 0) dGen.reserve(static_cast<size_t>(genPer));
		prm.genScratch = dGen.ptr; prm.genCapacity = dGen.capacity; prm.genStride = 0;
		prm.guardFail = nullptr;
		GSB_CUDA(gsb::launchFitKernel(prm, k, A, e->smemLimit, e->stream, nullptr));
		std::vector<double> out(static_cast<size_t>(k) * A + A);
		GSB_CUDA(cudaMemcpyAsync(out.data(), dOut.ptr, out.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
		GSB_CUDA(cudaStreamSynchronize(e->stream));
		// back to the caller's (uncentred) coordinates: slopes are unchanged, b0 = ybar + b0c - xbar'B
		for (int a = 0; a < A; ++a) {
			double shift = 0.0;
			for (int i = 0; i < k; ++i) {
				const double b = out[static_cast<size_t>(a) * k + i];
				coef[static_cast<size_t>(a) * k + i] = b;
				shift += e->xMean[cols[i]] * b;
			}
			intercepts[a] = e->yMean + out[static_cast<size_t>(k) * A + a] - shift;
		}
		for (int a = A; a < ncomp; ++a) { // components beyond the clamp do not exist in the reference either
			intercepts[a] = std::nan("");
			for (int i = 0; i < k; ++i) coef[static_cast<size_t>(a) * k + i] = std::nan("");
		}
	} catch (const CudaFailure& f) {
		return failCuda(e, f);
	} catch (const std::bad_alloc&) {
		return fail(e, GSB_ERR_MEMORY, "host allocation failed");
	}
	return GSB_OK;
}

} // extern "C"
