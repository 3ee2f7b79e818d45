// predict_blocks.cu -- K1p: the held-out predictions of the per-fit Gram-space path as ONE FP64 tensor-core product
// per (chromosome, held-out block).
//
// Replaces, for every fit that predicts the block,
//   PLS::predict                       src/PLS.cpp:34-47            newX * coef[:, a] + intercept[a]
//   the MSEP / residual accumulation   src/PLSEvaluator.cpp:263-268 (inner folds), :323-325 (outer folds)
//
// A held-out block is predicted by several fits (config 3/4: the outer fit of its fold and the four inner fits that
// hold it out), each with A component counts (coefficient columns are cumulative, src/PLSSimpls.cpp:152).  The fit
// kernel (simpls_fit.cu, fit-only instance) leaves, per (chromosome, fit), the rows [coef(k x A); -1; intercept(A)];
// with the matching columns [Z_b[:, sel] | y_b | 1] of the block this kernel forms
//
//   -E = [Z_b[:, sel] | y_b | 1] . [coef; -1; b0]          (rows of the block) x (tasks x A)
//
// i.e. the negated residuals of every (task, component count) in one product, on mma.sync.m8n8k4.f64 (SASS DMMA;
// tcgen05 has no FP64 kind).  The selected columns of the block are read ONCE for all its tasks (the fit kernel used
// to stage them once per task: 5x the gather traffic at config 4) and every operand fragment is shared by up to 14
// DMMAs.  Sum e and sum e^2 per column leave the kernel: MSEP for inner folds, pooled residual sums for outer folds.
//
// Shared memory, double-buffered over chunks of 32 reduction columns (16-byte / 8-byte cp.async):
//   As[32][lda]  column-major slice of the block: up to 128 rows of 32 selected columns (leading dimension == 4
//                mod 16: the 8x4 A fragments are bank-conflict free; 100 at config 4, 132 for 128-row chunks)
//   Bs[32][68]   the same 32 coefficient rows of up to 6 tasks x A <= 10 columns (68 == 4 mod 16)
// Warp w owns the 8-row tiles w, w + 8 of the 128-row chunk and every 8-column tile (blocks of fewer rows: the warps
// split the column tiles instead); a block of more than 128 rows is walked in chunks, more than 6 tasks in groups.
#include "kernels.cuh"

namespace gsb {

namespace {

constexpr int kPredThreads = 256;
constexpr int kPredKC = 32;      // reduction columns per stage
constexpr int kPredRows = 128;   // block rows per chunk
constexpr int kPredLdb = 68;
constexpr int kPredCols = 64;    // (task, component) columns per pass

__device__ __forceinline__ void dmmaP(double& d0, double& d1, double a, double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cpAsync16P(double* dst, const double* src) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src));
}
__device__ __forceinline__ void cpAsync8P(double* dst, const double* src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src));
}
__device__ __forceinline__ void cpAsyncCommitP() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpAsyncWaitP() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// The products of one staged chunk for a warp that owns MI row tiles and NI column tiles (both known when the chunk loop
// starts: as template arguments the loop body is loads and DMMAs only -- with the tile tests inside it every DMMA sat
// behind its own predicate, branch and WARPSYNC, 19 other instructions per DMMA in the ncu source view)
template <int MI, int NI>
__device__ __forceinline__ void predSteps(double (&acc)[2][8][2], const double* __restrict__ ap, const double* __restrict__ bp,
                                          int steps, int lda4, int aoff, int boff) {
	for (int s = 0; s < steps; ++s) {
		double a[MI], b[NI];
#pragma unroll
		for (int mi = 0; mi < MI; ++mi) a[mi] = ap[mi * aoff];
#pragma unroll
		for (int ni = 0; ni < NI; ++ni) b[ni] = bp[ni * boff];
#pragma unroll
		for (int mi = 0; mi < MI; ++mi)
#pragma unroll
			for (int ni = 0; ni < NI; ++ni) dmmaP(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
		ap += lda4;
		bp += 4 * kPredLdb;
	}
}

__global__ void __launch_bounds__(kPredThreads, 2) predictBlocksKernel(const PredictParams prm) {
	extern __shared__ __align__(16) double psm[];
	__shared__ uint32_t sel[kPredMaxK + 2];
	__shared__ double red[8][kPredCols][2]; // per warp slot (wrot: results do not depend on the rotation): sum e, sum e^2 of its row tiles, per column

	// block-major: the CTAs in flight share one or two blocks (their columns stay in L2)
	const int ab = blockIdx.x / prm.bucketCount;
	const int ci = blockIdx.x - ab * prm.bucketCount;
	const int chrom = prm.bucket[ci];
	const uint32_t selBegin = prm.offsets[chrom];
	const int k = static_cast<int>(prm.offsets[chrom + 1] - selBegin);
	const int A = min(prm.maxNComp, k);
	const int K2 = k + 2; // reduction length: the selected columns, y, the ones column
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int lrow = lane >> 2, lcol = lane & 3;
	const int kPredLda = prm.lda;
	const int kPredStageDoubles = kPredKC * (kPredLda + kPredLdb);
	const bool evenA = (A & 1) == 0;

	const BlockDesc bd = prm.blocks[prm.activeBlocks[ab]];
	const int taskBegin = prm.blockTaskOff[ab], taskEnd = prm.blockTaskOff[ab + 1];
	const double* __restrict__ zb = prm.zpool + bd.zOff;
	const double* __restrict__ coefC = prm.coef + prm.coefOff[chrom];
	const size_t fitStride = static_cast<size_t>(K2) * A;

	for (int i = tid; i < K2; i += kPredThreads) sel[i] = (i < k) ? prm.idx[selBegin + i] : static_cast<uint32_t>(prm.p + (i - k));

	// warp grid over the output tiles: WM warps along the rows (tiles wm, wm + WM), WN = 8 / WM along the columns
	const int mt0 = (min(kPredRows, bd.nrows) + 7) >> 3;
	const int WM = mt0 > 4 ? 8 : (mt0 > 2 ? 4 : (mt0 > 1 ? 2 : 1));
	const int WN = 8 / WM;
	// heavy and light warps alternate between the CTAs of an SM (13 row tiles over 8 warps at 100-row blocks)
	const int wrot = (warp + ((blockIdx.x & 1) ? 3 : 0)) & 7;
	const int wm = wrot % WM, wn = wrot / WM;
	const int tasksPerPass = max(1, min(kPredMaxTasks, kPredCols / A));

	for (int tb = taskBegin; tb < taskEnd; tb += tasksPerPass) {
		const int nt = min(tasksPerPass, taskEnd - tb);
		const int ncol = nt * A;
		const int ntiles = (ncol + 7) >> 3;
		for (int i = tid; i < 8 * kPredCols * 2; i += kPredThreads) (&red[0][0][0])[i] = 0.0;
		// the coefficient columns this lane copies in every staged row: 16-byte pieces (columns 2 lane, 2 lane + 1: a row of
		// a fit starts on an even offset when A is even) or two 8-byte pieces (columns lane, lane + 32)
		const double* bsrc[2];
		bool bval[2];
#pragma unroll
		for (int q = 0; q < 2; ++q) {
			const int j = evenA ? 2 * lane : lane + 32 * q;
			bval[q] = j < ncol && (evenA ? q == 0 : true);
			const int t = bval[q] ? j / A : 0, a = bval[q] ? j - t * A : 0;
			bsrc[q] = coefC + static_cast<size_t>(prm.ptasks[tb + t].fit) * fitStride + a;
		}
		const int nni = max(0, min(WM, (ntiles - wn + WN - 1) / WN)); // column tiles wn + ni WN < ntiles of this warp

		for (int m0 = 0; m0 < bd.nrows; m0 += kPredRows) {
			const int rows = min(kPredRows, bd.nrows - m0);
			const int rowsLd = min(kPredRows, bd.ld - m0); // multiple of 4; the copy is zero padded up to ld
			const int mtiles = (rows + 7) >> 3;
			double acc[2][8][2];
#pragma unroll
			for (int mi = 0; mi < 2; ++mi)
#pragma unroll
				for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

			const int nchunks = (K2 + kPredKC - 1) / kPredKC;
			auto stage = [&](int chunk) {
				double* __restrict__ As = psm + (chunk & 1) * kPredStageDoubles;
				double* __restrict__ Bs = As + kPredKC * kPredLda;
				const int c0 = chunk * kPredKC;
				const int nc = min(kPredKC, K2 - c0);
				const int ncPad = (nc + 3) & ~3;
				const int pieces = rowsLd >> 1;
				// a warp per staged column (row of Bs), the lanes along it: no index arithmetic per piece
				for (int cc = warp; cc < ncPad; cc += 8) {
					// padding columns of the last chunk: the ones column (finite) against zero coefficient rows
					const size_t col = sel[min(c0 + cc, K2 - 1)];
					const double* __restrict__ src = zb + col * bd.ld + m0;
					double* __restrict__ dst = As + cc * kPredLda;
					for (int piece = lane; piece < pieces; piece += 32) cpAsync16P(dst + 2 * piece, src + 2 * piece);
					double* __restrict__ brow = Bs + cc * kPredLdb;
					const size_t roff = static_cast<size_t>(c0 + cc) * A;
					if (evenA) {
						if (lane < ntiles * 4) {
							if (cc < nc && bval[0]) cpAsync16P(brow + 2 * lane, bsrc[0] + roff);
							else { brow[2 * lane] = 0.0; brow[2 * lane + 1] = 0.0; }
						}
					} else {
#pragma unroll
						for (int q = 0; q < 2; ++q) {
							const int j = lane + 32 * q;
							if (j < ntiles * 8) {
								if (cc < nc && bval[q]) cpAsync8P(brow + j, bsrc[q] + roff);
								else brow[j] = 0.0;
							}
						}
					}
				}
				cpAsyncCommitP();
			};

			__syncthreads(); // sel / red visible; the previous pass's buffers are free
			stage(0);
			for (int chunk = 0; chunk < nchunks; ++chunk) {
				if (chunk + 1 < nchunks) {
					stage(chunk + 1);
					cpAsyncWaitP<1>();
				} else {
					cpAsyncWaitP<0>();
				}
				__syncthreads();
				const double* __restrict__ As = psm + (chunk & 1) * kPredStageDoubles;
				const double* __restrict__ Bs = As + kPredKC * kPredLda;
				const int nc = min(kPredKC, K2 - chunk * kPredKC);
				const int steps = (nc + 3) >> 2;
				const double* __restrict__ ap = As + lcol * kPredLda + lrow + wm * 8;
				const double* __restrict__ bp = Bs + lcol * kPredLdb + lrow + wn * 8;
				const int nmi = (wm < mtiles ? 1 : 0) + (wm + WM < mtiles ? 1 : 0);
#define GSB_PRED_CASE(MI, NI) case (MI) * 16 + (NI): predSteps<MI, NI>(acc, ap, bp, steps, 4 * kPredLda, WM * 8, WN * 8); break;
				switch (nmi * 16 + nni) {
					GSB_PRED_CASE(1, 1) GSB_PRED_CASE(1, 2) GSB_PRED_CASE(1, 3) GSB_PRED_CASE(1, 4)
					GSB_PRED_CASE(1, 5) GSB_PRED_CASE(1, 6) GSB_PRED_CASE(1, 7) GSB_PRED_CASE(1, 8)
					GSB_PRED_CASE(2, 1) GSB_PRED_CASE(2, 2) GSB_PRED_CASE(2, 3) GSB_PRED_CASE(2, 4)
					GSB_PRED_CASE(2, 5) GSB_PRED_CASE(2, 6) GSB_PRED_CASE(2, 7) GSB_PRED_CASE(2, 8)
					default: break; // no tile of this chunk belongs to the warp
				}
#undef GSB_PRED_CASE
				__syncthreads(); // this buffer is refilled two chunks on
			}

			// acc = yhat + b0 - y = -e for row m0 + 8 mt + lrow, columns 8 nt8 + 2 lcol + {0, 1}
#pragma unroll
			for (int ni = 0; ni < 8; ++ni) {
				const int nt8 = wn + ni * WN;
				if (ni < WM && nt8 < ntiles) { // warp-uniform
					double se[2] = {0.0, 0.0}, sq[2] = {0.0, 0.0};
#pragma unroll
					for (int mi = 0; mi < 2; ++mi) {
						const int mt = wm + mi * WM;
						if (mt < mtiles && mt * 8 + lrow < rows) {
#pragma unroll
							for (int j = 0; j < 2; ++j) {
								se[j] -= acc[mi][ni][j];
								sq[j] = fma(acc[mi][ni][j], acc[mi][ni][j], sq[j]);
							}
						}
					}
#pragma unroll
					for (int m = 4; m <= 16; m <<= 1) {
#pragma unroll
						for (int j = 0; j < 2; ++j) {
							se[j] += __shfl_xor_sync(0xffffffffu, se[j], m);
							sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], m);
						}
					}
					if (lrow == 0) {
#pragma unroll
						for (int j = 0; j < 2; ++j) {
							red[wrot][nt8 * 8 + 2 * lcol + j][0] += se[j];
							red[wrot][nt8 * 8 + 2 * lcol + j][1] += sq[j];
						}
					}
				}
			}
		}
		__syncthreads();
		// totals over the warps in a fixed order; one thread per (task, component count)
		if (tid < ncol) {
			double se = 0.0, see = 0.0;
#pragma unroll
			for (int w = 0; w < 8; ++w) { se += red[w][tid][0]; see += red[w][tid][1]; }
			const int t = tid / A, a = tid - t * A;
			const PredTask pt = prm.ptasks[tb + t];
			if (pt.kind == 0) {
				// mean squared error of prediction of this fold (src/PLSEvaluator.cpp:266-268)
				prm.mse[(static_cast<size_t>(chrom) * prm.innerDests + pt.dest) * prm.maxNComp + a] = see / static_cast<double>(bd.nrows);
			} else {
				double* o = prm.ostat + ((static_cast<size_t>(chrom) * prm.outerDests + pt.dest) * prm.maxNComp + a) * 2;
				o[0] = se;
				o[1] = see;
			}
		}
		__syncthreads();
	}
}

} // namespace

// leading dimension of the staged block slice: at least the rows of a chunk, == 4 (mod 16)
int predictKernelLda(int maxBlockLd) {
	const int r = maxBlockLd < kPredRows ? maxBlockLd : kPredRows;
	return r + (((4 - r) % 16) + 16) % 16;
}

// One launch for the chromosomes of one subset-size class: CTA = (active held-out block, chromosome)
cudaError_t launchPredictKernel(const PredictParams& prm, cudaStream_t stream, long long* ctas) {
	const long long grid = static_cast<long long>(prm.nActive) * prm.bucketCount;
	if (grid <= 0) return cudaSuccess;
	if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
	// the 8-row tiles of the last chunk column may read up to 7 rows past the slice: into Bs, or the slack below
	const int bytes = (2 * kPredKC * (prm.lda + kPredLdb) + 8) * static_cast<int>(sizeof(double));
	cudaError_t err = cudaFuncSetAttribute(predictBlocksKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
	if (err != cudaSuccess) return err;
	predictBlocksKernel<<<static_cast<unsigned>(grid), kPredThreads, static_cast<size_t>(bytes), stream>>>(prm);
	if (ctas) *ctas += grid;
	return cudaGetLastError();
}

} // namespace gsb
