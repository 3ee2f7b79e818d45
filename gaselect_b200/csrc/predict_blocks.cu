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
				for (int e = tid; e < ncPad * pieces; e += kPredThreads) {
					const int cc = e / pieces, piece = e - cc * pieces;
					// padding columns of the last chunk: the ones column (finite) against zero coefficient rows
					const size_t col = sel[min(c0 + cc, K2 - 1)];
					cpAsync16P(As + cc * kPredLda + 2 * piece, zb + col * bd.ld + m0 + 2 * piece);
				}
				if (evenA) { // 16-byte pieces: coefficient rows start on even offsets
					const int half = ntiles * 4;
					for (int e = tid; e < ncPad * half; e += kPredThreads) {
						const int cc = e / half, j = 2 * (e - cc * half);
						double* dst = Bs + cc * kPredLdb + j;
						if (cc < nc && j < ncol) {
							const int t = j / A, a = j - t * A;
							cpAsync16P(dst, coefC + static_cast<size_t>(prm.ptasks[tb + t].fit) * fitStride + static_cast<size_t>(c0 + cc) * A + a);
						} else {
							dst[0] = 0.0;
							dst[1] = 0.0;
						}
					}
				} else {
					for (int e = tid; e < ncPad * (ntiles * 8); e += kPredThreads) {
						const int cc = e / (ntiles * 8), j = e - cc * (ntiles * 8);
						double* dst = Bs + cc * kPredLdb + j;
						if (cc < nc && j < ncol) {
							const int t = j / A, a = j - t * A;
							cpAsync8P(dst, coefC + static_cast<size_t>(prm.ptasks[tb + t].fit) * fitStride + static_cast<size_t>(c0 + cc) * A + a);
						} else {
							*dst = 0.0;
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
				const double* __restrict__ ap = As + lcol * kPredLda + lrow;
				const double* __restrict__ bp = Bs + lcol * kPredLdb + lrow;
				for (int s = 0; s < steps; ++s) {
					double a[2], b[8];
#pragma unroll
					for (int mi = 0; mi < 2; ++mi) {
						const int mt = wm + mi * WM;
						a[mi] = (mt < mtiles) ? ap[mt * 8] : 0.0;
					}
#pragma unroll
					for (int ni = 0; ni < 8; ++ni) {
						const int nt8 = wn + ni * WN;
						if (ni < WM && nt8 < ntiles) b[ni] = bp[nt8 * 8];
					}
#pragma unroll
					for (int mi = 0; mi < 2; ++mi) {
						if (wm + mi * WM < mtiles) {
#pragma unroll
							for (int ni = 0; ni < 8; ++ni) {
								if (ni < WM && wn + ni * WN < ntiles) dmmaP(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
							}
						}
					}
					ap += 4 * kPredLda;
					bp += 4 * kPredLdb;
				}
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
