// gram_build.cu -- K0, the one-off setup contraction (runs once per engine, not per generation).
//
// The reference recomputes X_tr'y and (implicitly, through t = X S and v = X't) X_tr'X_tr inside
// every fit (src/PLSSimpls.cpp:126,135,147).  Because the CV segmentation is fixed for a whole
// run (src/PLSEvaluator.cpp:56), those moments are hoisted here:
//   1. gramSyrkKernel  -- FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) SYRK  M_u = Z_u' Z_u  of the
//      augmented matrix Z = [Xc | yc | 1] for the full data and for every distinct held-out block;
//      lower-triangular 64x64 output tiles, operands staged through shared memory.
//   2. combine kernels -- each training set's centred moments from full-data sums minus the sums
//      of the blocks it excludes:  G_T = (M_full - sum_b M_b)[0:p,0:p] - n_T xm xm'  etc.
//      (SURVEY.md Appendix B), written as packed lower triangles that K1 gathers from.
#include "kernels.cuh"

namespace gsb {

namespace {

constexpr int kTile = 64;   // output tile edge
constexpr int kRows = 32;   // rows (samples) staged per pipeline stage
constexpr int kLds = kRows + 4; // padded row stride: (4 * col) mod 16 distinct for 4 consecutive cols
constexpr int kStages = 3;  // TMA pipeline depth
constexpr int kSyrkThreads = 128;
constexpr int kStageDoubles = 2 * kTile * kLds; // A panel + B panel

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
	             : "+d"(c0), "+d"(c1)
	             : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbarInit(uint64_t* bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_LOOP:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra WAIT_DONE;\n"
		"bra WAIT_LOOP;\n"
		"WAIT_DONE:\n"
		"}\n" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared (SASS UBLKCP), completion signalled on the mbarrier
__device__ __forceinline__ void tmaLoad1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smemAddr(dst)),
	             "l"(src), "r"(bytes), "r"(smemAddr(bar))
	             : "memory");
}

// grid.x: lower-triangular tile pair (ti >= tj); grid.y: unit (full data or one block).
// Operand panels (32 rows x 64 columns of Z_u, column-major: one contiguous 256-byte run per column) are
// staged by TMA bulk copies into a 3-stage shared-memory ring, each stage guarded by an mbarrier; the four
// warps (2 x 2, 32 x 32 outputs each) issue FP64 tensor-core MMAs (DMMA.8x8x4) from the stage that has landed
// while the next two stages are in flight.
__global__ void __launch_bounds__(kSyrkThreads) gramSyrkKernel(const double* const* __restrict__ unitPtrs,
                                                               const int* __restrict__ unitLd,
                                                               const int* __restrict__ unitRows,
                                                               double* const* __restrict__ outPtrs, int P2, int ldm) {
	extern __shared__ __align__(128) double ring[]; // kStages x [A panel | B panel]
	__shared__ __align__(8) uint64_t full[kStages];

	// decode blockIdx.x -> (ti, tj), tj <= ti
	int ti = static_cast<int>((sqrtf(8.0f * static_cast<float>(blockIdx.x) + 1.0f) - 1.0f) * 0.5f);
	while ((ti + 1) * (ti + 2) / 2 <= static_cast<int>(blockIdx.x)) ++ti;
	while (ti * (ti + 1) / 2 > static_cast<int>(blockIdx.x)) --ti;
	const int tj = static_cast<int>(blockIdx.x) - ti * (ti + 1) / 2;
	const int i0 = ti * kTile, j0 = tj * kTile;

	const int unit = blockIdx.y;
	const double* __restrict__ Zu = unitPtrs[unit];
	const int ld = unitLd[unit];
	const int nrows = unitRows[unit]; // padded to a multiple of 4 with zero rows
	double* __restrict__ M = outPtrs[unit];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int wi = warp >> 1, wj = warp & 1; // 2 x 2 warps, 32 x 32 outputs each
	const int group = lane >> 2, tig = lane & 3;

	// thread t < 64 feeds column t of the A panel, thread t >= 64 column t - 64 of the B panel
	const bool isA = tid < kTile;
	const int myCol = isA ? i0 + tid : j0 + tid - kTile;
	const bool myColValid = myCol < P2; // columns past the matrix are never stored: their stage rows may hold anything
	const int colsA = min(kTile, P2 - i0), colsB = min(kTile, P2 - j0);
	const double* __restrict__ mySrc = Zu + static_cast<size_t>(myColValid ? myCol : 0) * ld;
	const int myDst = (isA ? tid : kTile + (tid - kTile)) * kLds;

	if (tid == 0) {
		for (int s = 0; s < kStages; ++s) mbarInit(&full[s], 1);
		asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
	}
	__syncthreads();

	const int steps = (nrows + kRows - 1) / kRows;
	auto issue = [&](int step) {
		const int stage = step % kStages;
		const int r0 = step * kRows;
		const int rowsThis = min(kRows, nrows - r0); // multiple of 4
		if (tid == 0) mbarExpectTx(&full[stage], static_cast<uint32_t>((colsA + colsB) * rowsThis * sizeof(double)));
		if (myColValid) tmaLoad1d(ring + stage * kStageDoubles + myDst, mySrc + r0, static_cast<uint32_t>(rowsThis * sizeof(double)), &full[stage]);
	};
	for (int s = 0; s < kStages - 1 && s < steps; ++s) issue(s);

	double acc[4][4][2];
#pragma unroll
	for (int a = 0; a < 4; ++a)
#pragma unroll
		for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

	for (int step = 0; step < steps; ++step) {
		const int stage = step % kStages;
		// the stage consumed at step - 1 is free (barrier at the end of that step): refill it two steps ahead
		if (step + kStages - 1 < steps) issue(step + kStages - 1);
		mbarWait(&full[stage], static_cast<uint32_t>((step / kStages) & 1));
		const double* __restrict__ As = ring + stage * kStageDoubles;
		const double* __restrict__ Bs = As + kTile * kLds;
		const int ksteps = min(kRows, nrows - step * kRows) >> 2;
		for (int kk = 0; kk < ksteps; ++kk) {
			double a[4], b[4];
#pragma unroll
			for (int m = 0; m < 4; ++m) a[m] = As[(wi * 32 + m * 8 + group) * kLds + kk * 4 + tig];
#pragma unroll
			for (int n = 0; n < 4; ++n) b[n] = Bs[(wj * 32 + n * 8 + group) * kLds + kk * 4 + tig];
#pragma unroll
			for (int m = 0; m < 4; ++m)
#pragma unroll
				for (int n = 0; n < 4; ++n) dmma884(acc[m][n][0], acc[m][n][1], a[m], b[n]);
		}
		__syncthreads(); // everyone is done with this stage
	}

	// C fragment: row = group, cols = 2*tig, 2*tig+1
#pragma unroll
	for (int m = 0; m < 4; ++m) {
		const int i = i0 + wi * 32 + m * 8 + group;
		if (i >= P2) continue;
#pragma unroll
		for (int n = 0; n < 4; ++n) {
			const int j = j0 + wj * 32 + n * 8 + tig * 2;
			if (j < P2) M[static_cast<size_t>(i) * ldm + j] = acc[m][n][0];
			if (j + 1 < P2) M[static_cast<size_t>(i) * ldm + j + 1] = acc[m][n][1];
		}
	}
}

// per training set: column sums -> xm, s0 = X'y - sx*sy/n, ym
__global__ void combineVectorsKernel(const double* __restrict__ mFull, const double* const* __restrict__ mBlocks,
                                     const int* __restrict__ exclIndex, const int* __restrict__ exclCount,
                                     const int* __restrict__ fitIds, const int* __restrict__ fitNtr, int p, int ldm,
                                     double* __restrict__ s0, double* __restrict__ xm, double* __restrict__ ym) {
	const int f = blockIdx.y;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int fit = fitIds[f];
	const double nT = static_cast<double>(fitNtr[f]);
	const int ne = exclCount[f];
	const size_t rowY = static_cast<size_t>(p) * ldm, row1 = static_cast<size_t>(p + 1) * ldm;
	double sy = mFull[row1 + p];
	for (int e = 0; e < ne; ++e) sy -= mBlocks[exclIndex[f * 2 + e]][row1 + p];
	if (i == 0) ym[fit] = sy / nT;
	if (i >= p) return;
	double sx = mFull[row1 + i], xy = mFull[rowY + i];
	for (int e = 0; e < ne; ++e) {
		const double* mb = mBlocks[exclIndex[f * 2 + e]];
		sx -= mb[row1 + i];
		xy -= mb[rowY + i];
	}
	xm[static_cast<size_t>(fit) * p + i] = sx / nT;
	s0[static_cast<size_t>(fit) * p + i] = xy - sx * sy / nT;
}

// per training set: packed lower triangle of the centred Gram
__global__ void combineGramKernel(const double* __restrict__ mFull, const double* const* __restrict__ mBlocks,
                                  const int* __restrict__ exclIndex, const int* __restrict__ exclCount,
                                  const int* __restrict__ fitIds, const int* __restrict__ fitNtr, int p, int ldm,
                                  double* __restrict__ gramPool, const unsigned long long* __restrict__ gramOffs,
                                  const double* __restrict__ xm) {
	const int f = blockIdx.z;
	const int i = blockIdx.y;
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j > i) return;
	const int fit = fitIds[f];
	const double nT = static_cast<double>(fitNtr[f]);
	const int ne = exclCount[f];
	const size_t at = static_cast<size_t>(i) * ldm + j;
	double g = mFull[at];
	for (int e = 0; e < ne; ++e) g -= mBlocks[exclIndex[f * 2 + e]][at];
	const double* xmf = xm + static_cast<size_t>(fit) * p;
	g -= nT * xmf[i] * xmf[j];
	gramPool[gramOffs[fit] + tri(static_cast<unsigned long long>(i)) + j] = g;
}

// dst[:, c] = z[rows, cols ? cols[c] : c], zero-padded to ld rows
__global__ void blockGatherKernel(const double* __restrict__ z, int ldz, const uint32_t* __restrict__ cols,
                                  const uint32_t* __restrict__ rows, int nrows, double* __restrict__ dst, int ld) {
	const int c = blockIdx.y;
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= ld) return;
	const size_t src = cols ? cols[c] : static_cast<uint32_t>(c);
	// a row index of 0xffffffff asks for a zero row (padding inside a row-permuted copy)
	dst[static_cast<size_t>(c) * ld + r] = (r < nrows && rows[r] != 0xffffffffu) ? z[src * ldz + rows[r]] : 0.0;
}

} // namespace

void launchGramSyrk(const double* const* unitPtrs, const int* unitLd, const int* unitRows, double* const* outPtrs,
                    int units, int P2, int ldm, cudaStream_t stream) {
	if (units <= 0) return;
	const int tiles = (P2 + kTile - 1) / kTile;
	dim3 grid(static_cast<unsigned>(tiles * (tiles + 1) / 2), static_cast<unsigned>(units));
	const int bytes = kStages * kStageDoubles * static_cast<int>(sizeof(double));
	cudaFuncSetAttribute(gramSyrkKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
	gramSyrkKernel<<<grid, kSyrkThreads, bytes, stream>>>(unitPtrs, unitLd, unitRows, outPtrs, P2, ldm);
}

void launchGramCombine(const double* mFull, const double* const* mBlocks, const int* exclIndex, const int* exclCount,
                       const int* fitIds, const int* fitNtr, int nfitsChunk, int p, int ldm, double* gramPool,
                       const unsigned long long* gramOffs, double* s0, double* xm, double* ym, cudaStream_t stream) {
	if (nfitsChunk <= 0) return;
	dim3 gv(static_cast<unsigned>((p + 255) / 256), static_cast<unsigned>(nfitsChunk));
	combineVectorsKernel<<<gv, 256, 0, stream>>>(mFull, mBlocks, exclIndex, exclCount, fitIds, fitNtr, p, ldm, s0, xm, ym);
	dim3 gg(static_cast<unsigned>((p + 255) / 256), static_cast<unsigned>(p), static_cast<unsigned>(nfitsChunk));
	combineGramKernel<<<gg, 256, 0, stream>>>(mFull, mBlocks, exclIndex, exclCount, fitIds, fitNtr, p, ldm, gramPool,
	                                          gramOffs, xm);
}

void launchBlockGather(const double* z, int ldz, int ncols, const uint32_t* cols, const uint32_t* rows, int nrows,
                       double* dst, int ld, cudaStream_t stream) {
	dim3 grid(static_cast<unsigned>((ld + 127) / 128), static_cast<unsigned>(ncols));
	blockGatherKernel<<<grid, 128, 0, stream>>>(z, ldz, cols, rows, nrows, dst, ld);
}

} // namespace gsb
