// simpls_fit.cu -- K1, the steady-state hot kernel: one CTA per (chromosome, training set).
//
// Replaces, per CTA, what the reference does with
//   PLS::viewSelectColumns / viewSelectRows   src/PLS.cpp:13-23      (gather)
//   PLSSimpls::fit                             src/PLSSimpls.cpp:106-166
//   PLS::predict + MSEP / residual updates     src/PLS.cpp:34-47, src/PLSEvaluator.cpp:263-268,323-325
// in Gram space: with G = Xc'Xc and s0 = Xc'yc of the training rows,
//   w = G S, tn = sqrt(S'w), v = w/tn, q = s0'S/tn, then the reference's orthogonalisation against
//   the stored loadings, cumulative coefficients and deflation (src/PLSSimpls.cpp:151-162).
// All component counts 1..A come out of one fit (coefficient columns are cumulative,
// src/PLSSimpls.cpp:152), so the outer refit with optNComp components (src/PLSEvaluator.cpp:316-317)
// is column optNComp-1 of a fit with A components and every fit of a chromosome is independent.
//
// Data movement: the k(k+1)/2 selected entries of the training set's packed Gram are gathered
// from HBM/L2 into shared memory once (row-wise, no index decoding); the symmetric matrix-vector
// product reads the packed triangle conflict-free (consecutive rows start at consecutive
// triangular numbers, which are distinct modulo 16); held-out rows are read as contiguous column
// segments of the block copy with lanes = rows and the columns split over the warps.
//
// Reductions per component: ONE round for {S'w, s0'S, V_j'w (j < comp)} and ONE for {v'v, v'S}.
// The projections on the stored (orthonormal) loadings are therefore taken from w before the
// subtraction (all at once) rather than one after the other as src/PLSSimpls.cpp:57-63 does; the
// two orders differ by the loss of orthogonality of V, i.e. at rounding level.
#include "kernels.cuh"

namespace gsb {

namespace {

constexpr int kChunk = 8;   // components per prediction pass (accumulators held in registers)
constexpr int kDots = 6;    // loadings projected in the same reduction round as S'w and s0'S
constexpr int kRedMax = 16; // widest reduction (values per thread)

__device__ __forceinline__ double shflXor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

template <int NWARPS>
__device__ __forceinline__ void ctaSync() {
	if (NWARPS > 1) __syncthreads(); else __syncwarp();
}

// Sum NV values over the CTA; every thread receives bitwise-identical totals.
template <int NWARPS, int NV>
__device__ __forceinline__ void blockSum(double (&v)[NV], double* red, int& phase) {
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
		for (int i = 0; i < NV; ++i) v[i] += shflXor(v[i], m);
	}
	if (NWARPS > 1) {
		double* buf = red + phase * (NWARPS * kRedMax);
		const int warp = threadIdx.x >> 5;
		if ((threadIdx.x & 31) == 0) {
#pragma unroll
			for (int i = 0; i < NV; ++i) buf[warp * NV + i] = v[i];
		}
		__syncthreads();
#pragma unroll
		for (int i = 0; i < NV; ++i) {
			double s = buf[i];
#pragma unroll
			for (int w = 1; w < NWARPS; ++w) s += buf[w * NV + i];
			v[i] = s;
		}
		phase ^= 1;
	}
}

__host__ __device__ inline int roundUp(int v, int m) { return (v + m - 1) / m * m; }

// shared-memory carve-up (in doubles)
struct Carve {
	int P, S, s0, xm, V, coef, b0, red, pred, sel, total;
};

__host__ __device__ inline Carve carve(int k, int A, int nwarps, bool gramInSmem) {
	Carve c;
	const int Apad = roundUp(A, kChunk);
	int o = 0;
	c.P = o; o += gramInSmem ? roundUp(k * (k + 1) / 2, 2) : 0;
	c.S = o; o += roundUp(k, 2);
	c.s0 = o; o += roundUp(k, 2);
	c.xm = o; o += roundUp(k, 2);
	c.V = o; o += roundUp(A * k, 2);
	c.coef = o; o += k * Apad;
	c.b0 = o; o += Apad;
	c.red = o; o += 2 * nwarps * kRedMax;                // double-buffered
	c.pred = o; o += (nwarps > 1) ? nwarps * 32 * kChunk : 0; // cross-warp prediction partials
	c.sel = o; o += roundUp(k, 2) / 2 + 1;               // uint32 column ids
	c.total = o;
	return c;
}

template <int NWARPS, int RPT, bool GSMEM>
__global__ void __launch_bounds__(NWARPS * 32) simplsFitKernel(const FitParams prm) {
	constexpr int NT = NWARPS * 32;
	extern __shared__ __align__(16) double smem[];

	const int ci = blockIdx.x % prm.bucketCount;
	const int fi = blockIdx.x / prm.bucketCount;
	const int chrom = prm.bucket[ci];
	const uint32_t selBegin = prm.offsets[chrom];
	const int k = static_cast<int>(prm.offsets[chrom + 1] - selBegin);
	const int A = min(prm.maxNComp, k);
	const int Apad = roundUp(A, kChunk);
	const FitDesc fd = prm.fits[fi];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	const Carve cv = carve(k, A, NWARPS, GSMEM);
	double* __restrict__ P = smem + cv.P;
	double* __restrict__ S = smem + cv.S;
	double* __restrict__ s0s = smem + cv.s0;
	double* __restrict__ xms = smem + cv.xm;
	double* __restrict__ V = smem + cv.V;
	double* __restrict__ coef = smem + cv.coef;
	double* __restrict__ b0 = smem + cv.b0;
	double* red = smem + cv.red;
	double* predBuf = smem + cv.pred;
	uint32_t* __restrict__ sel = reinterpret_cast<uint32_t*>(smem + cv.sel);
	int phase = 0;

	// ---- gather (PLS::viewSelectColumns + viewSelectRows, as Gram entries) -----------------
	for (int i = tid; i < k; i += NT) sel[i] = prm.idx[selBegin + i];
	ctaSync<NWARPS>();
	const double* __restrict__ gsrc = prm.gram + fd.gramOff;
	if (GSMEM) {
		// packed row a = entries (a, 0..a); rows are dealt to the warps, four rows in flight per pass
		for (int a0 = warp * 4; a0 < k; a0 += NWARPS * 4) {
#pragma unroll
			for (int u = 0; u < 4; ++u) {
				const int a = a0 + u;
				if (a < k) {
					const double* __restrict__ grow = gsrc + tri(sel[a]);
					double* __restrict__ prow = P + a * (a + 1) / 2;
					for (int b = lane; b <= a; b += 32) prow[b] = __ldg(grow + sel[b]);
				}
			}
		}
	}
	{
		const double* __restrict__ s0g = prm.s0 + static_cast<size_t>(fi) * prm.p;
		const double* __restrict__ xmg = prm.xm + static_cast<size_t>(fi) * prm.p;
		for (int i = tid; i < k; i += NT) {
			const double s = __ldg(s0g + sel[i]);
			S[i] = s;
			s0s[i] = s;
			xms[i] = __ldg(xmg + sel[i]);
		}
		for (int i = tid; i < k * Apad; i += NT) coef[i] = 0.0;
	}
	ctaSync<NWARPS>();

	// rows of the k x k system owned by this thread
	int row[RPT];
	bool valid[RPT];
	int rowTri[RPT];
	unsigned long long rowTriG[RPT];
	uint32_t rowSel[RPT];
	double sOwn[RPT], s0Own[RPT];
#pragma unroll
	for (int u = 0; u < RPT; ++u) {
		const int r = tid + u * NT;
		valid[u] = r < k;
		row[u] = valid[u] ? r : k - 1;
		rowTri[u] = row[u] * (row[u] + 1) / 2;
		rowSel[u] = sel[row[u]];
		rowTriG[u] = tri(rowSel[u]);
		sOwn[u] = valid[u] ? S[row[u]] : 0.0;
		s0Own[u] = valid[u] ? s0s[row[u]] : 0.0;
	}

	// ---- SIMPLS components (src/PLSSimpls.cpp:133-163 in Gram space) -----------------------
	int done = A;
	for (int comp = 0; comp < A; ++comp) {
		double w[RPT];
#pragma unroll
		for (int u = 0; u < RPT; ++u) w[u] = 0.0;
		if (GSMEM) {
			int tc = 0;
#pragma unroll 4
			for (int c = 0; c < k; ++c) {
				const double sc = S[c];
#pragma unroll
				for (int u = 0; u < RPT; ++u) {
					const int addr = (c <= row[u]) ? rowTri[u] + c : tc + row[u];
					w[u] = fma(P[addr], sc, w[u]);
				}
				tc += c + 1;
			}
		} else {
			for (int c = 0; c < k; ++c) {
				const double sc = S[c];
				const uint32_t selc = sel[c];
				const unsigned long long tcg = tri(selc);
#pragma unroll
				for (int u = 0; u < RPT; ++u) {
					const unsigned long long addr = (c <= row[u]) ? rowTriG[u] + selc : tcg + rowSel[u];
					w[u] = fma(__ldg(gsrc + addr), sc, w[u]);
				}
			}
		}
		// round 1: tn^2 = S'GS = ||Xc S||^2 (:138), q*tn = s0'S (:148), projections V_j'w
		double r1[2 + kDots];
#pragma unroll
		for (int i = 0; i < 2 + kDots; ++i) r1[i] = 0.0;
#pragma unroll
		for (int u = 0; u < RPT; ++u) {
			if (valid[u]) {
				r1[0] = fma(sOwn[u], w[u], r1[0]);
				r1[1] = fma(s0Own[u], sOwn[u], r1[1]);
#pragma unroll
				for (int j = 0; j < kDots; ++j) {
					if (j < comp) r1[2 + j] = fma(V[j * k + row[u]], w[u], r1[2 + j]);
				}
			}
		}
		blockSum<NWARPS, 2 + kDots>(r1, red, phase);
		if (!(r1[0] >= 1e-50)) { // ||t|| < 1e-25 (NORM_TOL, :16,:141-143); also catches NaN / negative rounding
			done = comp;
			break;
		}
		const double tn = sqrt(r1[0]);
		const double q = r1[1] / tn;
		double v[RPT];
#pragma unroll
		for (int u = 0; u < RPT; ++u) v[u] = w[u];
#pragma unroll
		for (int j = 0; j < kDots; ++j) {
			if (j < comp) {
#pragma unroll
				for (int u = 0; u < RPT; ++u) v[u] = fma(-V[j * k + row[u]], r1[2 + j], v[u]);
			}
		}
		// more than kDots stored loadings (unconstrained component counts): further rounds of 8
		for (int j0 = kDots; j0 < comp; j0 += 8) {
			double d[8];
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				d[j] = 0.0;
				if (j0 + j < comp) {
#pragma unroll
					for (int u = 0; u < RPT; ++u) if (valid[u]) d[j] = fma(V[(j0 + j) * k + row[u]], w[u], d[j]);
				}
			}
			blockSum<NWARPS, 8>(d, red, phase);
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				if (j0 + j < comp) {
#pragma unroll
					for (int u = 0; u < RPT; ++u) v[u] = fma(-V[(j0 + j) * k + row[u]], d[j], v[u]);
				}
			}
		}
		// cumulative coefficients from S BEFORE deflation (:152-156); v = X't has the factor 1/tn (:145-147)
		const double f = q / tn;
		double r2[2] = {0.0, 0.0};
#pragma unroll
		for (int u = 0; u < RPT; ++u) {
			v[u] = v[u] / tn;
			if (valid[u]) {
				double* cr = coef + row[u] * Apad;
				cr[comp] = ((comp > 0) ? cr[comp - 1] : 0.0) + sOwn[u] * f;
				r2[0] = fma(v[u], v[u], r2[0]);
				r2[1] = fma(v[u], sOwn[u], r2[1]);
			}
		}
		// round 2: deflate S, normalise v (:160, :70-88)
		blockSum<NWARPS, 2>(r2, red, phase);
		const double dd = r2[1] / r2[0];
		const double vn = sqrt(r2[0]);
#pragma unroll
		for (int u = 0; u < RPT; ++u) {
			if (valid[u]) {
				sOwn[u] = fma(-v[u], dd, sOwn[u]);
				S[row[u]] = sOwn[u];
				V[comp * k + row[u]] = v[u] / vn;
			}
		}
		ctaSync<NWARPS>();
	}
	if (done < A) {
		// the reference throws std::underflow_error here; poison the remaining components so
		// that the epilogue reports GSB_STATUS_UNDERFLOW where the reference would have thrown
		const double nan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
		for (int u = 0; u < RPT; ++u) {
			if (valid[u]) for (int a = done; a < A; ++a) coef[row[u] * Apad + a] = nan;
		}
	}
	ctaSync<NWARPS>();

	// ---- intercepts b0[a] = ym - xm'coef[:,a] (:162) --------------------------------------
	for (int a0 = 0; a0 < A; a0 += kChunk) {
		double acc[kChunk];
#pragma unroll
		for (int u = 0; u < kChunk; ++u) acc[u] = 0.0;
		for (int r = tid; r < k; r += NT) {
			const double x = xms[r];
			const double* cr = coef + r * Apad + a0;
#pragma unroll
			for (int u = 0; u < kChunk; ++u) acc[u] = fma(x, cr[u], acc[u]);
		}
		blockSum<NWARPS, kChunk>(acc, red, phase);
		if (tid < kChunk) {
			double mine = 0.0;
#pragma unroll
			for (int u = 0; u < kChunk; ++u) if (u == tid) mine = acc[u];
			b0[a0 + tid] = fd.ym - mine;
		}
	}
	ctaSync<NWARPS>();

	if (prm.coefOut != nullptr) { // gsb_simpls: raw coefficients of this fit
		for (int i = tid; i < k * A; i += NT) prm.coefOut[i] = coef[(i % k) * Apad + (i / k)];
		for (int a = tid; a < A; a += NT) prm.coefOut[k * A + a] = b0[a];
	}

	// ---- held-out prediction and statistics -------------------------------------------------
	// lanes = rows of the block (coalesced column segments), warps split the selected columns;
	// the warps' partial predictions meet in shared memory and warp 0 finishes the statistics.
	for (int t = 0; t < fd.taskCount; ++t) {
		const TaskDesc td = prm.tasks[fd.taskBegin + t];
		const BlockDesc bd = prm.blocks[td.block];
		const double* __restrict__ zb = prm.zpool + bd.zOff;
		const double* __restrict__ ycol = zb + static_cast<size_t>(prm.p) * bd.ld;
		for (int a0 = 0; a0 < A; a0 += kChunk) {
			double st[2 * kChunk];
#pragma unroll
			for (int u = 0; u < 2 * kChunk; ++u) st[u] = 0.0;
			for (int r0 = 0; r0 < bd.nrows; r0 += 32) {
				const int r = r0 + lane;
				const bool live = r < bd.nrows;
				const double* __restrict__ zr = zb + (live ? r : 0);
				double acc[kChunk];
#pragma unroll
				for (int u = 0; u < kChunk; ++u) acc[u] = 0.0;
				int c = warp;
				for (; c + 3 * NWARPS < k; c += 4 * NWARPS) {
					double x[4];
#pragma unroll
					for (int j = 0; j < 4; ++j) x[j] = __ldg(zr + static_cast<size_t>(sel[c + j * NWARPS]) * bd.ld);
#pragma unroll
					for (int j = 0; j < 4; ++j) {
						const double* cr = coef + (c + j * NWARPS) * Apad + a0;
#pragma unroll
						for (int u = 0; u < kChunk; ++u) acc[u] = fma(x[j], cr[u], acc[u]);
					}
				}
				for (; c < k; c += NWARPS) {
					const double x = __ldg(zr + static_cast<size_t>(sel[c]) * bd.ld);
					const double* cr = coef + c * Apad + a0;
#pragma unroll
					for (int u = 0; u < kChunk; ++u) acc[u] = fma(x, cr[u], acc[u]);
				}
				if (NWARPS > 1) {
					__syncthreads(); // previous pass has been consumed
#pragma unroll
					for (int u = 0; u < kChunk; ++u) predBuf[(warp * kChunk + u) * 32 + lane] = acc[u];
					__syncthreads();
					if (warp == 0) {
#pragma unroll
						for (int u = 0; u < kChunk; ++u) {
							double s = predBuf[u * 32 + lane];
#pragma unroll
							for (int w2 = 1; w2 < NWARPS; ++w2) s += predBuf[(w2 * kChunk + u) * 32 + lane];
							acc[u] = s;
						}
					}
				}
				if (warp == 0 && live) {
					const double y = __ldg(ycol + r);
#pragma unroll
					for (int u = 0; u < kChunk; ++u) {
						const double e = y - (acc[u] + b0[a0 + u]);
						st[u] += e;
						st[kChunk + u] = fma(e, e, st[kChunk + u]);
					}
				}
			}
			if (warp == 0) {
#pragma unroll
				for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
					for (int u = 0; u < 2 * kChunk; ++u) st[u] += shflXor(st[u], m);
				}
				if (lane < kChunk && a0 + lane < A) {
					const int a = a0 + lane;
					double se = 0.0, see = 0.0;
#pragma unroll
					for (int u = 0; u < kChunk; ++u) {
						if (u == lane) { se = st[u]; see = st[kChunk + u]; }
					}
					if (td.kind == 0) {
						// mean squared error of prediction for this fold (src/PLSEvaluator.cpp:266-268)
						prm.mse[(static_cast<size_t>(chrom) * prm.innerDests + td.dest) * prm.maxNComp + a] =
							see / static_cast<double>(bd.nrows);
					} else {
						double* o = prm.ostat + ((static_cast<size_t>(chrom) * prm.outerDests + td.dest) * prm.maxNComp + a) * 2;
						o[0] = se;
						o[1] = see;
					}
				}
			}
		}
	}
}

template <int NWARPS, int RPT, bool GSMEM>
cudaError_t launchOne(const FitParams& prm, int kMax, int A, cudaStream_t stream, long long* ctas) {
	const Carve cv = carve(kMax, A, NWARPS, GSMEM);
	const size_t bytes = static_cast<size_t>(cv.total) * sizeof(double);
	cudaError_t err = cudaFuncSetAttribute(simplsFitKernel<NWARPS, RPT, GSMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                                       static_cast<int>(bytes));
	if (err != cudaSuccess) return err;
	const long long grid = static_cast<long long>(prm.nfits) * prm.bucketCount;
	if (grid <= 0) return cudaSuccess;
	if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
	simplsFitKernel<NWARPS, RPT, GSMEM><<<static_cast<unsigned>(grid), NWARPS * 32, bytes, stream>>>(prm);
	if (ctas) *ctas += grid;
	return cudaGetLastError();
}

int warpsFor(int kMax) {
	if (kMax <= 64) return 1;
	if (kMax <= 128) return 2;
	if (kMax <= 256) return 4;
	return 8;
}

} // namespace

int fitKernelSharedBytes(int k, int A) { return carve(k, A, warpsFor(k), true).total * static_cast<int>(sizeof(double)); }

// largest k whose packed Gram still fits in shared memory next to the loadings
int fitKernelMaxK(int A, int smemLimit) {
	int k = 1;
	while (k < 1024 && carve(k + 1, (A < k + 1) ? A : k + 1, warpsFor(k + 1), true).total * 8 <= smemLimit) ++k;
	return k;
}

// One launch for a bucket of chromosomes whose subset sizes are all <= kMax.
cudaError_t launchFitKernel(const FitParams& prm, int kMax, int A, int smemLimit, cudaStream_t stream, long long* ctas) {
	const int Ak = (A < kMax) ? A : kMax;
	const bool fits = carve(kMax, Ak, warpsFor(kMax), true).total * 8 <= smemLimit;
	if (fits) {
		if (kMax <= 32) return launchOne<1, 1, true>(prm, kMax, Ak, stream, ctas);
		if (kMax <= 64) return launchOne<1, 2, true>(prm, kMax, Ak, stream, ctas);
		if (kMax <= 128) return launchOne<2, 2, true>(prm, kMax, Ak, stream, ctas);
		if (kMax <= 256) return launchOne<4, 2, true>(prm, kMax, Ak, stream, ctas);
		return launchOne<8, 4, true>(prm, kMax, Ak, stream, ctas);
	}
	if (kMax <= 256) {
		if (carve(kMax, Ak, 4, false).total * 8 > smemLimit) return cudaErrorInvalidValue;
		return launchOne<4, 2, false>(prm, kMax, Ak, stream, ctas);
	}
	if (carve(kMax, Ak, 8, false).total * 8 > smemLimit) return cudaErrorInvalidValue;
	if (kMax <= 512) return launchOne<8, 2, false>(prm, kMax, Ak, stream, ctas);
	if (kMax <= 1024) return launchOne<8, 4, false>(prm, kMax, Ak, stream, ctas);
	return cudaErrorInvalidValue;
}

} // namespace gsb
