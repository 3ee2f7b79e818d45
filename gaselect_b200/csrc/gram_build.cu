// gram_build.cu -- K0, the one-off setup contraction (runs once per engine, not per generation).
//
// The reference recomputes X_tr'y and (implicitly, through t = X S and v = X't) X_tr'X_tr inside
// every fit (src/PLSSimpls.cpp:126,135,147).  Because the CV segmentation is fixed for a whole
// run (src/PLSEvaluator.cpp:56), those moments are hoisted here:
//   1. gramSyrkKernel  -- FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) SYRK  M_u = Z_u' Z_u  of the
//      augmented matrix Z = [Xc | yc | 1] for the full data and for every distinct held-out block;
//      lower-triangular 64x64 output tiles, operands staged through shared memory.
//   2. unitInterleaveKernel -- the moments of all units side by side per packed entry (i, j):
//      unitTab[tri(i) + j][u] = M_u[i][j]; likewise the column sums and x'y per column.  Once per engine.
//   3. gramPackVectorsKernel / gramPackKernel (PER GENERATION, ahead of K1) -- each training set's centred
//      moments of a chromosome's SELECTED columns from full-data sums minus the sums of the blocks it
//      excludes:  G_T = (M_full - sum_b M_b)[sel, sel] - n_T xm xm'  etc..  One
//      contiguous run of unitTab holds the entry of every unit, so the DRAM sectors it reads are fully
//      used by the chromosome's training sets, and K1 reads its packed Gram back as one contiguous slab.
//   4. combine kernels -- the same combination over ALL columns for one training set (LMEvaluator's
//      all-rows Gram, gsb_simpls).
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

// unitTab[(tri(i) + j) * upad + u] = M_u[i][j] for the units u in [unitBegin, unitBegin + unitCount) (<= kInterleaveUnits per
// launch), whose moment matrices are mUnits[srcBegin ...]; rows i = p and p + 1 of M_u are x'y and the column sums:
// they go to unitXy / unitSum [j][upad].
constexpr int kInterleaveUnits = 64;
__global__ void __launch_bounds__(256) unitInterleaveKernel(const double* const* __restrict__ mUnits, int srcBegin, int unitBegin, int unitCount,
                                                           int p, int ldm, int upad, double* __restrict__ unitTab,
                                                           double* __restrict__ unitSum, double* __restrict__ unitXy) {
	__shared__ double tile[32][kInterleaveUnits + 1];
	const int i = blockIdx.x;
	const int j0 = blockIdx.y * 32;
	const int jEnd = min(i < p ? i + 1 : p, j0 + 32);
	if (j0 >= jEnd) return;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int ul = warp; ul < unitCount; ul += 8) {
		const double* __restrict__ m = mUnits[srcBegin + ul];
		if (j0 + lane < jEnd) tile[lane][ul] = m[static_cast<size_t>(i) * ldm + j0 + lane];
	}
	__syncthreads();
	const int nj = jEnd - j0;
	for (int t = threadIdx.x; t < nj * unitCount; t += 256) {
		const int jl = t / unitCount, ul = t - jl * unitCount;
		const size_t j = static_cast<size_t>(j0 + jl);
		double* dst = i < p ? unitTab + (tri(static_cast<unsigned long long>(i)) + j) * upad : (i == p ? unitXy : unitSum) + j * upad;
		dst[unitBegin + ul] = tile[jl][ul];
	}
}

constexpr int kPackThreads = 256;

// s0 = X_T'y_T - sx sy / n_T and xm = sx / n_T of the selected columns for every training set of a chromosome
// (the arithmetic of combineVectorsKernel); grid: (32-column tiles, chromosomes of the launch)
__global__ void __launch_bounds__(kPackThreads) gramPackVectorsKernel(const PackParams prm) {
	extern __shared__ __align__(16) double psm[];
	const int chrom = prm.bucket[blockIdx.y];
	const uint32_t selBegin = prm.offsets[chrom];
	const int k = static_cast<int>(prm.offsets[chrom + 1] - selBegin);
	const int a0 = blockIdx.x * 32;
	if (a0 >= k) return;
	const int na = min(32, k - a0);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int ldt = prm.upad + 1;
	double* __restrict__ tSum = psm;
	double* __restrict__ tXy = psm + 32 * ldt;
	for (int a = warp; a < na; a += kPackThreads / 32) {
		const size_t col = prm.idx[selBegin + a0 + a];
		for (int u = lane; u < prm.upad; u += 32) {
			tSum[a * ldt + u] = prm.unitSum[col * prm.upad + u];
			tXy[a * ldt + u] = prm.unitXy[col * prm.upad + u];
		}
	}
	__syncthreads();
	const int kp = (k + 1) & ~1;
	double* __restrict__ out = prm.pvec + prm.vecOff[chrom];
	if (lane < na) {
		for (int f = warp; f < prm.nfits; f += kPackThreads / 32) {
			const PackFit pf = prm.pfits[f];
			double sx = tSum[lane * ldt], xy = tXy[lane * ldt];
			if (pf.u1 >= 0) { sx -= tSum[lane * ldt + pf.u1]; xy -= tXy[lane * ldt + pf.u1]; }
			if (pf.u2 >= 0) { sx -= tSum[lane * ldt + pf.u2]; xy -= tXy[lane * ldt + pf.u2]; }
			out[static_cast<size_t>(f) * 2 * kp + a0 + lane] = xy - sx * pf.sy * pf.rnT;
			out[static_cast<size_t>(f) * 2 * kp + kp + a0 + lane] = sx * pf.rnT;
		}
	}
}

// The packed centred Gram of the selected columns for every training set of a chromosome (the arithmetic of
// combineGramKernel).  One CTA per (tile of 2 rows x 32 columns of the lower triangle, chromosome):
//   1  the unit runs of its <= 64 entries are fetched (upad contiguous doubles each: every sector touched is used; four
//      runs per warp in flight: the column ids come from shared memory),
//   2  a warp per training set: the set's column means of the tile's 32 columns (one coalesced read of the vectors
//      gramPackVectorsKernel has just written) and of its 2 rows, g = t_full - t_u1 - t_u2 - n_T xm_i xm_j, written as
//      runs of 32 consecutive doubles of the set's packed slab.
// The kernel is bound by what it writes (150 slabs per chromosome at config 3/4); shared-memory traffic is 7 wavefronts
// per 32 entries written.
constexpr int kPackRows = 2;
constexpr int kPackCols = 32;

// tiles of a k-column subset: row pairs p = 0 .. ceil(k / 2) - 1 (rows 2p, 2p + 1), column blocks 0 .. (2p + 1) / 32
__host__ __device__ inline int packTiles(int k) {
	const int pairs = (k + 1) >> 1;
	const int m = pairs >> 4, r = pairs & 15;
	return pairs + 8 * m * (m - 1) + r * m;
}

__global__ void __launch_bounds__(kPackThreads, 4) gramPackKernel(const PackParams prm) {
	extern __shared__ __align__(16) double psm[];
	const int chrom = prm.bucket[blockIdx.y];
	const uint32_t selBegin = prm.offsets[chrom];
	const int k = static_cast<int>(prm.offsets[chrom + 1] - selBegin);
	if (static_cast<int>(blockIdx.x) >= packTiles(k)) return;
	// tile -> (row pair, column block): pairs [16 m, 16 m + 16) have m + 1 column blocks each
	int t = static_cast<int>(blockIdx.x), m = 0;
	while (t >= 16 * (m + 1)) { t -= 16 * (m + 1); ++m; }
	const int pair = 16 * m + t / (m + 1), jb = t % (m + 1);
	const int i0 = 2 * pair, j0 = 32 * jb;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int upad = prm.upad, ldt = upad | 1; // odd leading dimension: conflict-free across entries
	constexpr int NC = kPackRows + kPackCols;  // its rows, then its columns
	constexpr int NE = kPackRows * kPackCols;
	double* __restrict__ tile = psm;                                          // [NE][ldt] unit runs of the entries (row-major over the tile)
	double* __restrict__ xr = tile + NE * ldt;                                // [2][nfits] nT x (mean of the tile's row 0 / row 1) per training set
	PackFit* __restrict__ pfs = reinterpret_cast<PackFit*>(xr + 2 * prm.nfits); // [nfits]
	uint32_t* __restrict__ cols = reinterpret_cast<uint32_t*>(pfs + prm.nfits); // [NC] global column ids

	if (tid < NC) {
		const int local = tid < kPackRows ? i0 + tid : j0 + (tid - kPackRows);
		cols[tid] = prm.idx[selBegin + min(local, k - 1)];
	}
	{
		const int kp = (k + 1) & ~1;
		const double* __restrict__ vec = prm.pvec + prm.vecOff[chrom];
		const int i1c = min(i0 + 1, k - 1);
		for (int f = tid; f < prm.nfits; f += kPackThreads) {
			const PackFit pf = prm.pfits[f];
			const double* __restrict__ xmf = vec + static_cast<size_t>(f) * 2 * kp + kp;
			pfs[f] = pf;
			xr[f] = pf.nT * __ldg(xmf + i0);
			xr[prm.nfits + f] = pf.nT * __ldg(xmf + i1c);
		}
	}
	__syncthreads();

	// ---- 1: the runs.  Entry e = row * 32 + column; valid when column <= row < k (lower triangle).  8-byte cp.async
	// straight into the tile (its leading dimension is odd): every piece of the CTA is in flight at once -- through
	// registers, four runs per warp at a time, the kernel was latency-bound at 46 % of the HBM rate
	for (int e = warp; e < NE; e += kPackThreads / 32) {
		const int i = i0 + e / kPackCols, j = j0 + (e % kPackCols);
		double* __restrict__ dst = tile + e * ldt;
		if (i < k && j <= i) {
			// ascending column lists: sel[i] >= sel[j]
			const unsigned long long gi = cols[e / kPackCols], gj = cols[kPackRows + (e % kPackCols)];
			const double* __restrict__ src = prm.unitTab + (tri(gi) + gj) * upad;
			for (int u = lane; u < upad; u += 32) {
				asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst + u))), "l"(src + u));
			}
		} else {
			for (int u = lane; u < upad; u += 32) dst[u] = 0.0;
		}
	}
	asm volatile("cp.async.commit_group;\n" ::: "memory");
	// ---- 2: a warp per training set; lane = column of the tile
	const int kp = (k + 1) & ~1;
	const size_t slab = static_cast<size_t>(packedDoubles(static_cast<unsigned long long>(k)));
	double* __restrict__ out = prm.gpack + prm.packOff[chrom];
	const double* __restrict__ vec = prm.pvec + prm.vecOff[chrom];
	const int j = j0 + lane;
	const int jc = min(j, k - 1);
	const bool on0 = i0 < k && j <= i0, on1 = i0 + 1 < k && j <= i0 + 1;
	const double* __restrict__ tr0 = tile + lane * ldt;
	const double* __restrict__ tr1 = tile + (kPackCols + lane) * ldt;
	const size_t at0 = static_cast<size_t>(tri(static_cast<unsigned long long>(i0))) + j;
	const size_t at1 = static_cast<size_t>(tri(static_cast<unsigned long long>(i0 + 1))) + j;
	constexpr int FU = 8; // training sets per iteration: their vector reads are in flight together
	double xj[FU];
	auto loadVectors = [&](int f0) {
#pragma unroll
		for (int q = 0; q < FU; ++q) {
			const int f = min(f0 + q * (kPackThreads / 32), prm.nfits - 1);
			xj[q] = __ldg(vec + static_cast<size_t>(f) * 2 * kp + kp + jc);
		}
	};
	loadVectors(warp); // in flight with the runs
	asm volatile("cp.async.wait_group 0;\n" ::: "memory");
	__syncthreads();
	const double full0 = tr0[0], full1 = tr1[0];
	for (int f0 = warp; f0 < prm.nfits; f0 += FU * (kPackThreads / 32)) {
		if (f0 != warp) loadVectors(f0);
#pragma unroll
		for (int q = 0; q < FU; ++q) {
			const int f = f0 + q * (kPackThreads / 32);
			if (f < prm.nfits) {
				const int u1 = pfs[f].u1, u2 = pfs[f].u2;
				double* __restrict__ of = out + static_cast<size_t>(f) * slab;
				double g0 = full0, g1 = full1;
				if (u1 >= 0) { g0 -= tr0[u1]; g1 -= tr1[u1]; }
				if (u2 >= 0) { g0 -= tr0[u2]; g1 -= tr1[u2]; }
				g0 -= xr[f] * xj[q];
				g1 -= xr[prm.nfits + f] * xj[q];
				if (on0) of[at0] = g0;
				if (on1) of[at1] = g1;
			}
		}
	}
	// the padding double of an odd triangle
	if (blockIdx.x == 0 && ((k * (k + 1) / 2) & 1)) {
		for (int f = tid; f < prm.nfits; f += kPackThreads) out[static_cast<size_t>(f) * slab + static_cast<size_t>(k) * (k + 1) / 2] = 0.0;
	}
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

// mUnits[0 .. unitCount) are the moment matrices of the units unitBegin .. unitBegin + unitCount - 1
void launchUnitInterleave(const double* const* mUnits, int unitBegin, int unitCount, int p, int ldm, int upad, double* unitTab,
                          double* unitSum, double* unitXy, cudaStream_t stream) {
	for (int u0 = 0; u0 < unitCount; u0 += kInterleaveUnits) {
		const int cnt = min(kInterleaveUnits, unitCount - u0);
		dim3 grid(static_cast<unsigned>(p + 2), static_cast<unsigned>((p + 31) / 32));
		unitInterleaveKernel<<<grid, 256, 0, stream>>>(mUnits, u0, unitBegin + u0, cnt, p, ldm, upad, unitTab, unitSum, unitXy);
	}
}

// dynamic shared memory of the two pack kernels (the larger one): the engine refuses a plan whose unit runs and training
// sets outgrow it before any table is built
long long gramPackSharedBytes(int upad, int nfits) {
	const long long NC = kPackRows + kPackCols, NE = kPackRows * kPackCols;
	const long long vbytes = 2LL * 32 * (upad + 1) * static_cast<long long>(sizeof(double));
	const long long gbytes = (NE * (upad | 1) + 2LL * nfits) * static_cast<long long>(sizeof(double)) + static_cast<long long>(nfits) * static_cast<long long>(sizeof(PackFit)) +
	                         NC * static_cast<long long>(sizeof(uint32_t)) + 16;
	return vbytes > gbytes ? vbytes : gbytes;
}

// the vectors and then the packed Grams of every training set for the chromosomes of one launch (subsets <= kMax)
cudaError_t launchGramPack(const PackParams& prm, int kMax, cudaStream_t stream) {
	if (prm.bucketCount <= 0 || prm.nfits <= 0 || kMax <= 0) return cudaSuccess;
	if (prm.bucketCount > 65535 || prm.nfits > 65535) return cudaErrorInvalidConfiguration;
	const int ldt = prm.upad + 1;
	const int vbytes = 2 * 32 * ldt * static_cast<int>(sizeof(double));
	cudaError_t err = cudaFuncSetAttribute(gramPackVectorsKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, vbytes);
	if (err != cudaSuccess) return err;
	dim3 gv(static_cast<unsigned>((kMax + 31) / 32), static_cast<unsigned>(prm.bucketCount));
	gramPackVectorsKernel<<<gv, kPackThreads, static_cast<size_t>(vbytes), stream>>>(prm);
	const int ldg = prm.upad | 1;
	const int NC = kPackRows + kPackCols, NE = kPackRows * kPackCols;
	const int gbytes = (NE * ldg + 2 * prm.nfits) * static_cast<int>(sizeof(double)) + prm.nfits * static_cast<int>(sizeof(PackFit)) +
	                   NC * static_cast<int>(sizeof(uint32_t)) + 16;
	err = cudaFuncSetAttribute(gramPackKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gbytes);
	if (err != cudaSuccess) return err;
	dim3 gg(static_cast<unsigned>(packTiles(kMax)), static_cast<unsigned>(prm.bucketCount));
	gramPackKernel<<<gg, kPackThreads, static_cast<size_t>(gbytes), stream>>>(prm);
	return cudaGetLastError();
}

void launchBlockGather(const double* z, int ldz, int ncols, const uint32_t* cols, const uint32_t* rows, int nrows,
                       double* dst, int ld, cudaStream_t stream) {
	dim3 grid(static_cast<unsigned>((ld + 127) / 128), static_cast<unsigned>(ncols));
	blockGatherKernel<<<grid, 128, 0, stream>>>(z, ldz, cols, rows, nrows, dst, ld);
}

} // namespace gsb
