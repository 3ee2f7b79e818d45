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
// Two kernels, one launch per subset-size class of the population (engine.cpp, concurrent streams):
//   simplsFitFast<10, SPILL>   A <= 10 components, k <= 384 -- everything at BASELINE configs 3-5.  One thread
//       per Gram row; the per-row SIMPLS state (S, s0, x-bar, the stored loadings, the cumulative
//       coefficients) lives in registers; shared memory holds the packed Gram, S and reduction scratch.
//       SPILL: the packed rows that do not fit in shared memory (k > 212) are read in place from the fit's slab (L2).
//   simplsFitKernel<RPT, GSMEM> the generic path: any number of components (BASELINE config 1 runs up to 30
//       per fit), loadings and coefficients in shared memory; GSMEM = false reads the Gram in place.
//
// Data movement
//   * the training set's packed centred Gram of the selected columns (k(k+1)/2 doubles) has been written as one
//     contiguous slab by gramPackKernel (gram_build.cu) just before this launch; the fast path fetches it with ONE
//     TMA bulk copy (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier) issued by one thread;
//   * the symmetric matrix-vector product reads the packed triangle conflict-free (consecutive
//     rows start at consecutive triangular numbers, distinct modulo 16); one thread per row,
//     four independent accumulators, warp-uniform split into "left of the diagonal band" (own
//     packed row), the 32-wide band, and "below the band" (column access, lane = row);
//   * held-out rows: after the last component the Gram area is dead; the selected column segments (+ y) of
//     the fit's 1-2 held-out blocks are staged there with 16-byte cp.async, one or two blocks per pass
//     (whatever costs no occupancy); work units (block x half of the component counts) run on separate
//     warps with lanes = rows, so no partial sums cross warps.
//
// Reductions per component: ONE round for {S'w, s0'S, V_j'w (j < comp)} and ONE for {v'v, v'S}; a round over
// 16 values costs 16 shuffle-adds (the lanes halve the value set at every butterfly step); 1/tn and 1/|v|
// come from rsqrt.  The projections on the stored (orthonormal) loadings are therefore taken from w before
// the subtraction (all at once) rather than one after the other as src/PLSSimpls.cpp:57-63 does; the two
// orders differ by the loss of orthogonality of V, i.e. at rounding level.
#include "kernels.cuh"
#include "fit_common.cuh"

namespace gsb {

namespace {

constexpr int kChunk = 10;  // components per prediction pass (accumulators held in registers)
constexpr int kDots = 14;   // loadings projected in the same reduction round as S'w and s0'S (16 values: one blockSum16)
constexpr int kRedMax = 20; // widest reduction (values per thread)
constexpr int kMaxWarps = 13; // 416 threads: subsets up to k = 416 on the spill path

// Sum NV values over the CTA (nw warps); every thread receives bitwise-identical totals.
template <int NV>
__device__ __forceinline__ void blockSum(double (&v)[NV], double* red, int& phase, int nw) {
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
		for (int i = 0; i < NV; ++i) v[i] += shflXor(v[i], m);
	}
	if (nw > 1) {
		double* buf = red + phase * (nw * kRedMax);
		const int warp = threadIdx.x >> 5;
		if ((threadIdx.x & 31) == 0) {
#pragma unroll
			for (int i = 0; i < NV; ++i) buf[warp * NV + i] = v[i];
		}
		__syncthreads();
#pragma unroll
		for (int i = 0; i < NV; ++i) v[i] = buf[i];
		for (int w = 1; w < nw; ++w) {
#pragma unroll
			for (int i = 0; i < NV; ++i) v[i] += buf[w * NV + i];
		}
		phase ^= 1;
	}
}



// shared-memory carve-up (in doubles)
struct Carve {
	int P, S, s0, xm, V, coef, b0, red, sel, total;
};

// staging area of the generic path after the fit: the warps' partial sums + at least min(k, 32) column tiles
__host__ __device__ inline int stageMin(int k, int nw) { return nw * 32 * kChunk + 32 * min(k, 32); }

// vGlobal: the stored loadings V (A x k) and the cumulative coefficients (k x Apad) live in a global scratch (one region
// per CTA of the launch, L2-resident while it runs) instead of shared memory: unconstrained component counts on wide
// subsets (evaluatorPLS defaults on spectra: A = min(k, n_tr - 1) ~ 100 components of 200 columns = 320 KB)
// The staging area of the held-out rows (after the fit) is the packed Gram AND the stored loadings behind it: both are
// dead by then (with only the Gram area, a 30-column subset with 30 components took 26 KB per fit for 10 KB of staging
// tiles: 8 single-warp CTAs per SM; now 11).
__host__ __device__ inline int stageCapOf(int k, int A, bool gramInSmem, int nw, bool vGlobal) {
	const int sizeP = gramInSmem ? roundUp(k * (k + 1) / 2, 2) : 0;
	const int sizeV = vGlobal ? 0 : roundUp(A * k, 2);
	return roundUp(max(sizeP + sizeV, stageMin(k, nw)), 2);
}
__host__ __device__ inline Carve carve(int k, int A, bool gramInSmem, int nw, bool vGlobal = false) {
	Carve c;
	const int Apad = roundUp(A, kChunk);
	int o = 0;
	// packed Gram, then the stored loadings; afterwards the staging area of the held-out rows (>= one 32-row column per warp)
	c.P = o;
	c.V = o + (gramInSmem ? roundUp(k * (k + 1) / 2, 2) : 0);
	o += stageCapOf(k, A, gramInSmem, nw, vGlobal);
	c.S = o; o += roundUp(k, 2);
	c.s0 = o; o += roundUp(k, 2);
	c.xm = o; o += roundUp(k, 2);
	c.coef = o; o += vGlobal ? 0 : k * Apad;
	c.b0 = o; o += Apad;
	c.red = o; o += 2 * nw * kRedMax;  // double-buffered
	c.sel = o; o += roundUp(k, 2) / 2 + 1;    // uint32 column ids
	c.total = o;
	return c;
}

// RPT rows of the k x k system per thread (1 on the shared-memory path: blockDim.x >= k)
template <int RPT, bool GSMEM, bool VGLOB = false>
__global__ void __launch_bounds__(kMaxWarps * 32) simplsFitKernel(const FitParams prm) {
	extern __shared__ __align__(16) double smem[];
	const int NT = blockDim.x, NW = NT >> 5;

	const int ci = blockIdx.x % prm.bucketCount;
	const int fi = blockIdx.x / prm.bucketCount;
	const int chrom = prm.bucket[ci];
	const uint32_t selBegin = prm.offsets[chrom];
	const int k = static_cast<int>(prm.offsets[chrom + 1] - selBegin);
	const int A = min(prm.maxNComp, k);
	const int Apad = roundUp(A, kChunk);
	const FitDesc fd = prm.fits[fi];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	const Carve cv = carve(k, A, GSMEM, NW, VGLOB);
	double* __restrict__ P = smem + cv.P;
	double* __restrict__ S = smem + cv.S;
	double* __restrict__ s0s = smem + cv.s0;
	double* __restrict__ xms = smem + cv.xm;
	double* __restrict__ V = VGLOB ? prm.genScratch + static_cast<size_t>(blockIdx.x) * prm.genStride : smem + cv.V;
	double* __restrict__ coef = VGLOB ? V + roundUp(A * k, 2) : smem + cv.coef;
	double* __restrict__ b0 = smem + cv.b0;
	double* red = smem + cv.red;
	uint32_t* __restrict__ sel = reinterpret_cast<uint32_t*>(smem + cv.sel);
	int phase = 0;

	// ---- the fit's packed Gram and vectors (PLS::viewSelectColumns + viewSelectRows, as moments; gramPackKernel) ----
	for (int i = tid; i < k; i += NT) sel[i] = prm.idx[selBegin + i];
	const int kp = roundUp(k, 2);
	const size_t slab = static_cast<size_t>(packedDoubles(static_cast<unsigned long long>(k)));
	const double* __restrict__ gsrc = prm.gpack + prm.packOff[chrom] + static_cast<size_t>(fi) * slab;
	const double* __restrict__ vsrc = prm.pvec + prm.vecOff[chrom] + static_cast<size_t>(fi) * 2 * kp;
	if (GSMEM) {
		for (int i = tid; i < static_cast<int>(slab >> 1); i += NT) cpAsync16(P + 2 * i, gsrc + 2 * i);
	}
	{
		for (int i = tid; i < k; i += NT) {
			const double sv = __ldg(vsrc + i);
			S[i] = sv;
			s0s[i] = sv;
			xms[i] = __ldg(vsrc + kp + i);
		}
		for (int i = tid; i < k * Apad; i += NT) coef[i] = 0.0;
	}
	cpAsyncWaitAll();
	__syncthreads();

	// rows of the k x k system owned by this thread
	int row[RPT];
	bool valid[RPT];
	double sOwn[RPT], s0Own[RPT];
#pragma unroll
	for (int u = 0; u < RPT; ++u) {
		const int r = tid + u * NT;
		valid[u] = r < k;
		row[u] = valid[u] ? r : k - 1;
		sOwn[u] = valid[u] ? S[row[u]] : 0.0;
		s0Own[u] = valid[u] ? s0s[row[u]] : 0.0;
	}

	// ---- SIMPLS components (src/PLSSimpls.cpp:133-163 in Gram space) -----------------------
	int done = A;
	for (int comp = 0; comp < A; ++comp) {
		double w[RPT];
		if (GSMEM) {
#pragma unroll
			for (int u = 0; u < RPT; ++u) {
				const int r = row[u];
				// rows of a warp are consecutive: [rlo, rhi] is warp-uniform
				const int rlo = min(warp * 32 + u * NT, k - 1);
				const int rhi = min(rlo + 31, k - 1);
				const double* __restrict__ pr = P + r * (r + 1) / 2;
				double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
				int c = 0;
				for (; c + 3 < rlo; c += 4) { // left of the band: own packed row
					a0 = fma(pr[c], S[c], a0);
					a1 = fma(pr[c + 1], S[c + 1], a1);
					a2 = fma(pr[c + 2], S[c + 2], a2);
					a3 = fma(pr[c + 3], S[c + 3], a3);
				}
				for (; c < rlo; ++c) a0 = fma(pr[c], S[c], a0);
				int tc = rlo * (rlo + 1) / 2;
				for (; c <= rhi; ++c) {       // the band: either side of the diagonal
					const double g = (c <= r) ? pr[c] : P[tc + r];
					a1 = fma(g, S[c], a1);
					tc += c + 1;
				}
				const double* __restrict__ pc = P + r; // below the band: column r of rows c
				for (; c + 3 < k; c += 4) {
					const int t1 = tc + c + 1, t2 = t1 + c + 2, t3 = t2 + c + 3;
					a0 = fma(pc[tc], S[c], a0);
					a1 = fma(pc[t1], S[c + 1], a1);
					a2 = fma(pc[t2], S[c + 2], a2);
					a3 = fma(pc[t3], S[c + 3], a3);
					tc = t3 + c + 4;
				}
				for (; c < k; ++c) {
					a2 = fma(pc[tc], S[c], a2);
					tc += c + 1;
				}
				w[u] = (a0 + a1) + (a2 + a3);
			}
		} else {
#pragma unroll
			for (int u = 0; u < RPT; ++u) {
				const int rt = row[u] * (row[u] + 1) / 2;
				double a0 = 0.0;
				for (int c = 0; c < k; ++c) {
					const int addr = (c <= row[u]) ? rt + c : c * (c + 1) / 2 + row[u];
					a0 = fma(__ldg(gsrc + addr), S[c], a0);
				}
				w[u] = a0;
			}
		}
		// round 1: tn^2 = S'GS = ||Xc S||^2 (:138), q*tn = s0'S (:148), projections V_j'w on the first kDots stored
		// loadings -- 16 values, one butterfly in which the lanes halve the value set at every step
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
		blockSum16<2 + kDots, kRedMax>(r1, red, phase, NW, lane, warp);
		if (!(r1[0] >= 1e-50)) { // ||t|| < 1e-25 (NORM_TOL, :16,:141-143); also catches NaN / negative rounding
			done = comp;
			break;
		}
		const double rtn = rsqrt(r1[0]); // 1/tn: no division, no square root (as in the fast kernel)
		const double q = r1[1] * rtn;
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
		// more than kDots stored loadings (unconstrained component counts): further rounds of 16
		for (int j0 = kDots; j0 < comp; j0 += 16) {
			double d[16];
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				d[j] = 0.0;
				if (j0 + j < comp) {
#pragma unroll
					for (int u = 0; u < RPT; ++u) if (valid[u]) d[j] = fma(V[(j0 + j) * k + row[u]], w[u], d[j]);
				}
			}
			blockSum16<16, kRedMax>(d, red, phase, NW, lane, warp);
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				if (j0 + j < comp) {
#pragma unroll
					for (int u = 0; u < RPT; ++u) v[u] = fma(-V[(j0 + j) * k + row[u]], d[j], v[u]);
				}
			}
		}
		// cumulative coefficients from S BEFORE deflation (:152-156); v = X't has the factor 1/tn (:145-147)
		const double f = q * rtn;
		double r2[2] = {0.0, 0.0};
#pragma unroll
		for (int u = 0; u < RPT; ++u) {
			v[u] = v[u] * rtn;
			if (valid[u]) {
				double* cr = coef + row[u] * Apad;
				cr[comp] = ((comp > 0) ? cr[comp - 1] : 0.0) + sOwn[u] * f;
				r2[0] = fma(v[u], v[u], r2[0]);
				r2[1] = fma(v[u], sOwn[u], r2[1]);
			}
		}
		// round 2: deflate S, normalise v (:160, :70-88)
		blockSum2<kRedMax>(r2, red, phase, NW, lane, warp);
		const double rvn = rsqrt(r2[0]);
		const double dd = (r2[1] * rvn) * rvn;
#pragma unroll
		for (int u = 0; u < RPT; ++u) {
			if (valid[u]) {
				sOwn[u] = fma(-v[u], dd, sOwn[u]);
				S[row[u]] = sOwn[u];
				V[comp * k + row[u]] = v[u] * rvn;
			}
		}
		__syncthreads();
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
	__syncthreads();

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
		blockSum<kChunk>(acc, red, phase, NW);
		if (tid < kChunk) {
			double mine = 0.0;
#pragma unroll
			for (int u = 0; u < kChunk; ++u) if (u == tid) mine = acc[u];
			b0[a0 + tid] = fd.ym - mine;
		}
	}
	__syncthreads();

	if (prm.coefOut != nullptr) { // gsb_simpls: raw coefficients of this fit
		for (int i = tid; i < k * A; i += NT) prm.coefOut[i] = coef[(i % k) * Apad + (i / k)];
		for (int a = tid; a < A; a += NT) prm.coefOut[k * A + a] = b0[a];
	}

	// ---- held-out prediction and statistics -------------------------------------------------
	// The Gram area is dead now: it stages [column pass][32-row slab] tiles of the held-out block
	// (16-byte cp.async, everything in flight at once) and, at its end, the warps' partial sums.
	const int stageCap = stageCapOf(k, A, GSMEM, NW, VGLOB);
	double* __restrict__ part = P + stageCap - NW * 32 * kChunk; // [warp][u][lane]
	const int colsPerPass = max(1, min(k, (stageCap - NW * 32 * kChunk) / 32));
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
				const int slab = min(32, bd.ld - r0); // multiple of 4 rows, zero padded in the block copy
				double acc[kChunk];
#pragma unroll
				for (int u = 0; u < kChunk; ++u) acc[u] = 0.0;
				for (int c0 = 0; c0 < k; c0 += colsPerPass) {
					const int nc = min(colsPerPass, k - c0);
					const int quads = slab >> 1; // 16-byte pieces per column
					for (int e = tid; e < nc * quads; e += NT) {
						const int c = e / quads, piece = e - c * quads;
						cpAsync16(P + c * 32 + piece * 2, zb + static_cast<size_t>(sel[c0 + c]) * bd.ld + r0 + piece * 2);
					}
					cpAsyncWaitAll();
					__syncthreads();
					if (lane < slab) {
						for (int c = warp; c < nc; c += NW) {
							const double x = P[c * 32 + lane];
							const double* cr = coef + (c0 + c) * Apad + a0;
#pragma unroll
							for (int u = 0; u < kChunk; ++u) acc[u] = fma(x, cr[u], acc[u]);
						}
					}
					__syncthreads(); // staging area free again
				}
				if (NW > 1) {
#pragma unroll
					for (int u = 0; u < kChunk; ++u) part[(warp * kChunk + u) * 32 + lane] = acc[u];
					__syncthreads();
					if (warp == 0) {
#pragma unroll
						for (int u = 0; u < kChunk; ++u) {
							double s = part[u * 32 + lane];
							for (int w2 = 1; w2 < NW; ++w2) s += part[(w2 * kChunk + u) * 32 + lane];
							acc[u] = s;
						}
					}
					__syncthreads();
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


// ---------------------------------------------------------------------------------------------
// Fast path: A <= AREG components, k <= 256.  The per-row state -- S, s0, xm, the stored loadings
// V[0..A) and the cumulative coefficients -- is only ever touched by the row's own thread
// (src/PLSSimpls.cpp:150-162 are all row-wise), so it lives in registers; shared memory holds the
// packed Gram, S (broadcast operand of the matrix-vector product) and the reduction scratch.
// That halves the footprint of a fit and doubles the CTAs resident per SM.  After the last
// component the Gram area is dead and takes the coefficients and the held-out rows.
//
// Reductions: ONE round per component for {S'w, s0'S, V_j'w for every stored loading} and one for
// {v'v, v'S}.  A round over 16 values costs 16 shuffle-adds (the lanes halve the value set at every
// butterfly step) instead of 80, and 1/tn, 1/|v| come from rsqrt (no division, no sqrt).
// shared memory of the spill path beside the packed rows: the spilled rows' own products (k - rs) and, per warp, its share
// of the column products of the spilled entries (k each)
__host__ __device__ inline int spillScratchDoubles(int k, int rs) {
	if (rs >= k) return 0;
	return roundUp(k - rs, 2) + (roundUp(k, 32) / 32) * roundUp(k, 2);
}

struct FastCarve {
	int P, S, red, sel, total;
	int spill;                        // spill path: wrow[k - rs], then part[warp][k]
	int coef, part, stage, stageCols; // inside the P area, after the fit
	int xt;                           // > 0: whole held-out blocks are staged at P + xt ...
	int xtTasks;                      // ... this many per pass (1 or 2)
};

// wantXt < 0: host decision; else the number of held-out blocks staged per pass (0, 1 or 2), as told
// rowsInSmem: packed Gram rows kept in shared memory (k on the plain path; the rest is read in place from the fit's slab in L2)
__host__ __device__ inline FastCarve fastCarve(int k, int A, int nw, int maxTasks, int maxBlockLd, int wantXt, int rowsInSmem) {
	FastCarve c;
	const int Apad = roundUp(A, 2);
	const int rs = min(k, rowsInSmem);
	const int ntri = rs * (rs + 1) / 2;
	const int coefD = k * Apad, partD = (nw > 1) ? nw * 32 * kChunk : 0;
	const int minStage = 32 * min(k, 8);
	const int plain = roundUp(max(ntri, coefD + partD + minStage), 2);
	const int xt1 = (k + 1) * 32;
	int stagedTasks = wantXt;
	if (wantXt < 0) {
		// stage both blocks at once when that keeps the CTAs per SM, else one block per pass when the
		// footprint grows by at most a third; large blocks (more than 32 rows) take the slab path
		stagedTasks = 0;
		if (maxTasks >= 1 && maxTasks <= 2 && maxBlockLd >= 1) {
			const int fixed = roundUp(k, 2) + 2 * nw * 16 + roundUp(k, 2) / 2 + 1;
			const int perPlain = (plain + fixed) * 8 + 1024;
			const int residentPlain = min(233472 / perPlain, 2048 / (nw * 32));
			for (int tasks = maxTasks; tasks >= 1 && stagedTasks == 0; --tasks) {
				const int area = roundUp(max(ntri, coefD + tasks * xt1), 2);
				const int per = (area + fixed) * 8 + 1024;
				if (per > 232448) continue;
				if (233472 / per >= residentPlain || (tasks == 1 && 3 * area <= 4 * plain)) stagedTasks = tasks;
			}
		}
	}
	const int area = stagedTasks > 0 ? roundUp(max(ntri, coefD + stagedTasks * xt1), 2) : plain;
	int o = 0;
	c.P = o; o += area;
	c.S = o; o += roundUp(k, 2);
	c.red = o; o += 2 * nw * 16; // double-buffered partial sums, 16 values per warp
	c.sel = o; o += roundUp(k, 2) / 2 + 1;
	c.spill = o; o += spillScratchDoubles(k, rs);
	c.total = o;
	c.coef = 0;
	c.part = coefD;
	c.stage = coefD + partD;
	c.stageCols = (area - c.stage) / 32;
	c.xt = stagedTasks > 0 ? coefD : 0;
	c.xtTasks = stagedTasks;
	return c;
}

// fit-only instances (no prediction in the kernel: predict_blocks.cu): the packed Gram rows [0, rowsInSmem), S, reduction scratch
__host__ __device__ inline FastCarve fitOnlyCarve(int k, int nw, int rowsInSmem) {
	FastCarve c;
	const int rs = min(k, rowsInSmem);
	int o = 0;
	c.P = o; o += roundUp(rs * (rs + 1) / 2, 2);
	c.S = o; o += roundUp(k, 2);
	c.red = o; o += 2 * nw * 16;
	c.sel = o;
	c.spill = o; o += spillScratchDoubles(k, rs);
	c.total = o;
	c.coef = c.part = c.stage = c.stageCols = c.xt = c.xtTasks = 0;
	return c;
}

// One row of w = G S from the packed triangle in shared memory.  Thread r owns row r; the rows of a warp
// are consecutive, [rlo, rhi] is warp-uniform.  Three warp-uniform phases: left of the 32-wide diagonal band
// (own packed row), the band (either side of the diagonal), below the band (column r of the rows c).
__device__ __forceinline__ double symvRow(const double* __restrict__ P, const double* __restrict__ S, int k, int r, int rlo, int rhi) {
	const double* __restrict__ pr = P + r * (r + 1) / 2;
	const double* __restrict__ pc = P + r;
	double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
	int c = 0;
	for (; c + 7 < rlo; c += 8) {
		const double2 s01 = *reinterpret_cast<const double2*>(S + c);
		const double2 s23 = *reinterpret_cast<const double2*>(S + c + 2);
		const double2 s45 = *reinterpret_cast<const double2*>(S + c + 4);
		const double2 s67 = *reinterpret_cast<const double2*>(S + c + 6);
		a0 = fma(pr[c], s01.x, a0);
		a1 = fma(pr[c + 1], s01.y, a1);
		a2 = fma(pr[c + 2], s23.x, a2);
		a3 = fma(pr[c + 3], s23.y, a3);
		a0 = fma(pr[c + 4], s45.x, a0);
		a1 = fma(pr[c + 5], s45.y, a1);
		a2 = fma(pr[c + 6], s67.x, a2);
		a3 = fma(pr[c + 7], s67.y, a3);
	}
	for (; c < rlo; ++c) a0 = fma(pr[c], S[c], a0);
	int tc = rlo * (rlo + 1) / 2;
	for (; c + 1 <= rhi; c += 2) {
		const double2 s01 = *reinterpret_cast<const double2*>(S + c);
		const int t1 = tc + c + 1;
		const double g0 = (c <= r) ? pr[c] : pc[tc];
		const double g1 = (c + 1 <= r) ? pr[c + 1] : pc[t1];
		a0 = fma(g0, s01.x, a0);
		a1 = fma(g1, s01.y, a1);
		tc = t1 + c + 2;
	}
	for (; c <= rhi; ++c) {
		const double g = (c <= r) ? pr[c] : pc[tc];
		a2 = fma(g, S[c], a2);
		tc += c + 1;
	}
	for (; c + 7 < k; c += 8) {
		const double2 s01 = *reinterpret_cast<const double2*>(S + c);
		const double2 s23 = *reinterpret_cast<const double2*>(S + c + 2);
		const double2 s45 = *reinterpret_cast<const double2*>(S + c + 4);
		const double2 s67 = *reinterpret_cast<const double2*>(S + c + 6);
		const int t1 = tc + c + 1, t2 = t1 + c + 2, t3 = t2 + c + 3, t4 = t3 + c + 4;
		const int t5 = t4 + c + 5, t6 = t5 + c + 6, t7 = t6 + c + 7;
		a0 = fma(pc[tc], s01.x, a0);
		a1 = fma(pc[t1], s01.y, a1);
		a2 = fma(pc[t2], s23.x, a2);
		a3 = fma(pc[t3], s23.y, a3);
		a0 = fma(pc[t4], s45.x, a0);
		a1 = fma(pc[t5], s45.y, a1);
		a2 = fma(pc[t6], s67.x, a2);
		a3 = fma(pc[t7], s67.y, a3);
		tc = t7 + c + 8;
	}
	for (; c < k; ++c) {
		a3 = fma(pc[tc], S[c], a3);
		tc += c + 1;
	}
	return (a0 + a1) + (a2 + a3);
}

// w = G S when only the packed rows [0, rs) are in shared memory (rs is a multiple of 32 below k, so a warp's rows are
// all on one side); the rows [rs, k) stay in the fit's packed slab gp in global memory (L2-resident while the CTA runs).
//   * warps whose rows are in shared memory: symvRow on the rs x rs triangle;
//   * then EVERY warp takes its share of the spilled rows (row rs + warp, rs + warp + warps, ...), one row at a time with
//     the lanes across its columns: coalesced 256-byte reads, every spilled entry fetched ONCE and used for both of its
//     products (row c: summed over the lanes; column j: a per-lane accumulator, colacc[m] for j = lane + 32 m).  All
//     warps, because a row costs an L2 round trip: with the spilled rows left to the few warps that own them a
//     256-column subset (one such warp) took twice as long as with per-thread runs;
//   * one barrier; every row then adds the warps' column shares (fixed order: bitwise reproducible) and, if it is a
//     spilled row, its own product.
// (The first form read the spilled rows as per-thread runs -- a 32-byte sector per 8 useful bytes -- and every spilled
// entry twice; at k = 384 the step took 3.9 x what the k^2 trend of the shared-memory path predicts.)
constexpr int kSpillCols = kMaxWarps; // column slots of 32 per lane
__device__ __forceinline__ double symvSpill(const double* __restrict__ P, const double* __restrict__ S, const double* __restrict__ gp,
                                            int k, int rs, int r, int rlo, int rhi, int warp, int lane, double* __restrict__ wrow,
                                            double* __restrict__ part) {
	const int kp = roundUp(k, 2);
	const int nsw = roundUp(k, 32) / 32; // warps of this subset (a class of subsets shares one multiple of 32)
	double wa = 0.0;
	if (rlo < rs) wa = symvRow(P, S, rs, r, rlo, rhi);
	{
		const int sw = warp;
		const int cFirst = sw < nsw ? rs + sw : k;
		double colacc[kSpillCols];
#pragma unroll
		for (int m = 0; m < kSpillCols; ++m) colacc[m] = 0.0;
		// two rows per iteration: their loads are in flight together (a row is one L2 round trip)
		for (int c = cFirst; c < k; c += 2 * nsw) {
			const int c2 = c + nsw;
			const bool two = c2 < k;
			const double* __restrict__ row = gp + static_cast<size_t>(c) * (c + 1) / 2;
			const double* __restrict__ row2 = gp + static_cast<size_t>(two ? c2 : c) * ((two ? c2 : c) + 1) / 2;
			double g[kSpillCols], h[kSpillCols];
#pragma unroll
			for (int m = 0; m < kSpillCols; ++m) {
				const int j = lane + 32 * m;
				g[m] = (j <= c) ? __ldcg(row + j) : 0.0;
				h[m] = (two && j <= c2) ? __ldcg(row2 + j) : 0.0;
			}
			const double sc = S[c], sc2 = S[two ? c2 : c];
			double acc = 0.0, acc2 = 0.0;
#pragma unroll
			for (int m = 0; m < kSpillCols; ++m) {
				const int j = lane + 32 * m;
				if (32 * m <= (two ? c2 : c)) { // warp-uniform
					const double sj = S[min(j, k - 1)];
					acc = fma(g[m], sj, acc);
					acc2 = fma(h[m], sj, acc2);
					if (j < c) colacc[m] = fma(g[m], sc, colacc[m]);
					if (two && j < c2) colacc[m] = fma(h[m], sc2, colacc[m]);
				}
			}
#pragma unroll
			for (int m = 16; m >= 1; m >>= 1) { acc += shflXor(acc, m); acc2 += shflXor(acc2, m); }
			if (lane == 0) {
				wrow[c - rs] = acc;
				if (two) wrow[c2 - rs] = acc2;
			}
		}
#pragma unroll
		for (int m = 0; m < kSpillCols; ++m) {
			const int j = lane + 32 * m;
			if (j < k && sw < nsw) part[sw * kp + j] = colacc[m];
		}
	}
	__syncthreads();
	double w = wa;
	for (int s2 = 0; s2 < nsw; ++s2) w += part[s2 * kp + r];
	if (r >= rs) w += wrow[r - rs];
	return w;
}

// SPILL: packed Gram rows >= rSplit (a multiple of 32) do not fit in shared memory; every matrix-vector product reads
// them in place from the fit's slab (L2-resident while the CTA runs): own rows as per-thread runs, column accesses
// coalesced.
// TRACE: the instance with the clock64() stamps, launched only when prm.trace is set (even predicated off they cost
// a few per cent in the hot loops, measured on the kernel-space kernel)
// PRED: the held-out rows are predicted in the kernel; else the coefficients leave it for predictBlocksKernel
template <int AREG, bool SPILL, bool TRACE, bool PRED>
__global__ void __launch_bounds__(SPILL ? kMaxWarps * 32 : 256, SPILL ? 1 : 2) simplsFitFast(const FitParams prm, const int stageAll, const int rSplit) {
	extern __shared__ __align__(16) double smem[];
	static_assert(AREG % 2 == 0 && AREG + 2 <= 16, "one reduction round holds tn^2, q*tn and every projection");
	const int NT = blockDim.x, NW = NT >> 5;

	const int ci = blockIdx.x % prm.bucketCount;
	const int fi = blockIdx.x / prm.bucketCount;
	const int chrom = prm.bucket[ci];
	const uint32_t selBegin = prm.offsets[chrom];
	const int k = static_cast<int>(prm.offsets[chrom + 1] - selBegin);
	const int A = min(prm.maxNComp, k);
	const int Apad = roundUp(A, 2);
	const FitDesc fd = prm.fits[fi];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	const int rs = SPILL ? min(k, rSplit) : k; // Gram rows [0, rs) live in shared memory
	const FastCarve cv = PRED ? fastCarve(k, A, NW, prm.maxTasks, prm.maxBlockLd, stageAll, rs) : fitOnlyCarve(k, NW, rs);
	double* __restrict__ P = smem + cv.P;
	double* __restrict__ S = smem + cv.S;
	double* red = smem + cv.red;
	uint32_t* __restrict__ sel = reinterpret_cast<uint32_t*>(smem + cv.sel);
	int phase = 0;
	__shared__ __align__(8) uint64_t gramBar;

	// GSB_GUARD: a canary behind the carve (the launch asks for 64 bytes more), checked where the CTA leaves
	constexpr unsigned long long kCanary = 0xA5C3A5C3A5C3A5C3ull;
	unsigned long long* canary = reinterpret_cast<unsigned long long*>(smem + cv.total);
	if (prm.guardFail != nullptr && tid < 8) canary[tid] = kCanary;
	auto checkCanary = [&]() {
		if (prm.guardFail != nullptr) {
			__syncthreads();
			if (tid < 8 && canary[tid] != kCanary) atomicAdd(prm.guardFail, 1);
		}
	};

	const bool tracing = TRACE && prm.trace != nullptr && blockIdx.x == gridDim.x / 2 && tid == 0;
	int stamp = 0;
#define GSB_STAMP() do { if (TRACE && tracing) prm.trace[stamp++] = clock64(); } while (0)
	GSB_STAMP();

	// ---- the fit's packed Gram: one contiguous slab (gramPackKernel), fetched by TMA bulk copies ------------------
	const int kp = roundUp(k, 2);
	const size_t slab = static_cast<size_t>(packedDoubles(static_cast<unsigned long long>(k)));
	const double* __restrict__ gsrc = prm.gpack + prm.packOff[chrom] + static_cast<size_t>(fi) * slab;
	const double* __restrict__ vsrc = prm.pvec + prm.vecOff[chrom] + static_cast<size_t>(fi) * 2 * kp;
	if (tid == 0) {
		mbarInitF(&gramBar, 1);
		// rows [0, rs) (all of them unless SPILL); the slab is padded to an even number of doubles
		const uint32_t bytes = static_cast<uint32_t>(packedDoubles(static_cast<unsigned long long>(rs)) * sizeof(double));
		mbarExpectTxF(&gramBar, bytes);
		for (uint32_t o = 0; o < bytes; o += 32768u) {
			bulkLoadF(reinterpret_cast<char*>(P) + o, reinterpret_cast<const char*>(gsrc) + o, min(32768u, bytes - o), &gramBar);
		}
	}
	if (PRED) for (int i = tid; i < k; i += NT) sel[i] = prm.idx[selBegin + i];
	// descriptors of the held-out blocks this fit predicts (needed much later: load them now)
	TaskDesc td[2];
	BlockDesc bd[2];
#pragma unroll
	for (int t = 0; t < 2; ++t) {
		if (PRED && t < fd.taskCount) {
			td[t] = prm.tasks[fd.taskBegin + t];
			bd[t] = prm.blocks[td[t].block];
		} else {
			td[t].block = 0; td[t].kind = 0; td[t].dest = 0; td[t].pad = 0;
			bd[t].zOff = 0; bd[t].ld = 0; bd[t].nrows = 0;
		}
	}
	const bool valid = tid < k;
	const int r = valid ? tid : k - 1;
	double sOwn, s0Own, xmOwn;
	{
		s0Own = __ldg(vsrc + r);
		xmOwn = __ldg(vsrc + kp + r);
		sOwn = s0Own;
		if (valid) S[r] = sOwn;
	}
	double vOwn[AREG], cOwn[AREG];
#pragma unroll
	for (int j = 0; j < AREG; ++j) { vOwn[j] = 0.0; cOwn[j] = 0.0; }
	__syncthreads();          // sel, S; the barrier's initialisation
	mbarWaitF(&gramBar, 0u);  // the packed Gram has landed
	GSB_STAMP(); // 1: gather done

	// rows of a warp are consecutive: [rlo, rhi] is warp-uniform
	const int rlo = min(warp * 32, k - 1);
	const int rhi = min(rlo + 31, k - 1);
	const double nanv = __longlong_as_double(0x7ff8000000000000LL);

	// ---- SIMPLS components -------------------------------------------------------------------
	double cum = 0.0;
#pragma unroll 1
	for (int comp = 0; comp < A; ++comp) {
		const double w = SPILL ? symvSpill(P, S, gsrc, k, rs, r, rlo, rhi, warp, lane, smem + cv.spill, smem + cv.spill + roundUp(k - rs, 2))
		                       : symvRow(P, S, k, r, rlo, rhi);

		// ONE round: tn^2 = S'GS (:138), q*tn = s0'S (:148), projections V_j'w on every stored loading
		double r1[2 + AREG];
		r1[0] = valid ? sOwn * w : 0.0;
		r1[1] = valid ? s0Own * sOwn : 0.0;
#pragma unroll
		for (int j = 0; j < AREG; ++j) r1[2 + j] = (valid && j < comp) ? vOwn[j] * w : 0.0;
		blockSum16<2 + AREG>(r1, red, phase, NW, lane, warp);
		const bool under = !(r1[0] >= 1e-50); // ||t|| < 1e-25 (NORM_TOL, :16,:141-143); also NaN / negative rounding
		const double rtn = rsqrt(r1[0]);      // 1/tn; tn itself is never needed
		const double q = r1[1] * rtn;
		double v = w;
#pragma unroll
		for (int j = 0; j < AREG; ++j) v = fma(-vOwn[j], r1[2 + j], v);
		// cumulative coefficients from S BEFORE deflation (:152-156); v = X't carries 1/tn (:145-147)
		v *= rtn;
		cum += sOwn * (q * rtn);
		double r2[2];
		r2[0] = valid ? v * v : 0.0;
		r2[1] = valid ? v * sOwn : 0.0;
		blockSum2<16>(r2, red, phase, NW, lane, warp); // second round: deflate S, normalise v (:160, :70-88)
		const double rvn = rsqrt(r2[0]);
		const double dd = (r2[1] * rvn) * rvn;
		sOwn = fma(-v, dd, sOwn);
		if (under) { sOwn = nanv; cum = nanv; } // the reference throws: poison this and every later component
		if (valid) S[r] = sOwn;
#pragma unroll
		for (int j = 0; j < AREG; ++j) {
			if (j == comp) { vOwn[j] = v * rvn; cOwn[j] = cum; }
		}
		__syncthreads();
		GSB_STAMP(); // 2 + comp
	}

	if (!PRED) {
		// ---- the rows [coef; -1; intercepts] of this fit leave the kernel (predict_blocks.cu) ---------------
		double* __restrict__ cdst = prm.coefScratch + prm.coefOff[chrom] + static_cast<size_t>(fi) * (static_cast<size_t>(k) + 2) * A;
		if (valid) {
#pragma unroll
			for (int j = 0; j < AREG; ++j) if (j < A) cdst[static_cast<size_t>(r) * A + j] = cOwn[j];
		}
		// intercepts b0[a] = ym - xm'coef[:,a] (:162)
		double b0s[AREG];
#pragma unroll
		for (int j = 0; j < AREG; ++j) b0s[j] = (valid && j < A) ? xmOwn * cOwn[j] : 0.0;
		blockSum16<AREG>(b0s, red, phase, NW, lane, warp);
		if (tid == 0) {
#pragma unroll
			for (int j = 0; j < AREG; ++j) {
				if (j < A) {
					cdst[static_cast<size_t>(k) * A + j] = -1.0;
					cdst[static_cast<size_t>(k + 1) * A + j] = fd.ym - b0s[j];
				}
			}
		}
		GSB_STAMP();
		checkCanary();
		return;
	}

	// ---- the Gram area is dead: coefficients go there, the held-out rows are fetched ---------------
	double* __restrict__ coef = P + cv.coef;
	const bool staged = cv.xt > 0;
	const int tasksPerPass = cv.xtTasks;
	double* __restrict__ XT = P + cv.xt; // [slot][column c <= k][32 rows]; column k is the block's y
	// slot t - firstTask <- rows [32 * slab, 32 * slab + 32) of the selected columns (+ y) of held-out block t
	auto stageBlocks = [&](int firstTask, int slab) {
#pragma unroll
		for (int t = 0; t < 2; ++t) {
			if (t >= firstTask && t < firstTask + tasksPerPass && t < fd.taskCount && slab * 32 < bd[t].ld) {
				const double* __restrict__ zb = prm.zpool + bd[t].zOff + slab * 32;
				const int pieces = min(32, bd[t].ld - slab * 32) >> 1; // 16-byte pieces per column
				double* __restrict__ dst = XT + (t - firstTask) * (k + 1) * 32;
				for (int e = tid; e < (k + 1) * 16; e += NT) {
					const int c = e >> 4, piece = e & 15;
					if (piece < pieces) {
						const size_t col = (c < k) ? sel[c] : static_cast<size_t>(prm.p);
						cpAsync16(dst + c * 32 + piece * 2, zb + col * bd[t].ld + piece * 2);
					}
				}
			}
		}
		asm volatile("cp.async.commit_group;\n" ::: "memory");
	};
	if (staged) stageBlocks(0, 0);
	if (valid) {
#pragma unroll
		for (int j = 0; j < AREG; ++j) if (j < Apad) coef[r * Apad + j] = (j < A) ? cOwn[j] : 0.0;
	}
	// intercepts b0[a] = ym - xm'coef[:,a] (:162)
	double b0v[AREG];
#pragma unroll
	for (int j = 0; j < AREG; ++j) b0v[j] = (valid && j < A) ? xmOwn * cOwn[j] : 0.0;
	blockSum16<AREG>(b0v, red, phase, NW, lane, warp);
#pragma unroll
	for (int j = 0; j < AREG; ++j) b0v[j] = fd.ym - b0v[j];
	__syncthreads();

	if (prm.coefOut != nullptr) { // gsb_simpls: raw coefficients of this fit
		for (int i = tid; i < k * A; i += NT) prm.coefOut[i] = coef[(i % k) * Apad + (i / k)];
		if (tid == 0) {
#pragma unroll
			for (int j = 0; j < AREG; ++j) if (j < A) prm.coefOut[k * A + j] = b0v[j];
		}
	}
	GSB_STAMP(); // intercepts done

	// ---- held-out prediction and statistics -------------------------------------------------
	if (staged) {
		// work unit = (task, half of the component counts): no partial sums cross warps.  Blocks of more than 32
		// rows go through the staging tiles slab by slab; a unit's sums over the slabs stay in its warp's registers.
		constexpr int H = (AREG / 2 + 1) & ~1; // components of the first half, even (6 of 10); second half AREG - H
		for (int firstTask = 0; firstTask < fd.taskCount; firstTask += tasksPerPass) {
			const int tasksHere = min(tasksPerPass, fd.taskCount - firstTask);
			const int units = tasksHere * 2;
			int maxLd = 0;
#pragma unroll
			for (int t = 0; t < 2; ++t) if (t >= firstTask && t < firstTask + tasksHere) maxLd = max(maxLd, bd[t].ld);
			const int slabs = (maxLd + 31) >> 5;
			double tot[4] = {0.0, 0.0, 0.0, 0.0}; // this lane's value (lane >> 1) of up to 4 units handled by this warp
			for (int slab = 0; slab < slabs; ++slab) {
				if (firstTask > 0 || slab > 0) {
					__syncthreads(); // previous tiles consumed
					stageBlocks(firstTask, slab);
				}
				cpAsyncWaitAll();
				__syncthreads();
#pragma unroll
				for (int ui = 0; ui < 4; ++ui) {
					const int u = warp + ui * NW;
					if (u >= units) break;
					const int t = firstTask + (u >> 1);
					const bool second = (u & 1) != 0;
					const int base = second ? H : 0;
					const int nrows = (t == 0) ? bd[0].nrows : bd[1].nrows;
					const int rowsHere = nrows - slab * 32; // rows of this block in this slab (may be <= 0)
					if (base >= A || rowsHere <= 0) continue;
					const double* __restrict__ xt = XT + (u >> 1) * (k + 1) * 32 + lane;
					const double* __restrict__ cr = coef + base;
					double acc[H];
#pragma unroll
					for (int i = 0; i < H; ++i) acc[i] = 0.0;
					// Apad is even and >= A; the pairs of this half that lie inside Apad (warp-uniform count)
					const int pairs = min((second ? AREG - H : H), Apad - base) >> 1;
					if (pairs == H / 2) {
#pragma unroll 4
						for (int c = 0; c < k; ++c) {
							const double x = xt[c * 32];
#pragma unroll
							for (int i = 0; i < H; i += 2) {
								const double2 cc = *reinterpret_cast<const double2*>(cr + c * Apad + i);
								acc[i] = fma(x, cc.x, acc[i]);
								acc[i + 1] = fma(x, cc.y, acc[i + 1]);
							}
						}
					} else if (pairs == (AREG - H) / 2) {
#pragma unroll 4
						for (int c = 0; c < k; ++c) {
							const double x = xt[c * 32];
#pragma unroll
							for (int i = 0; i < AREG - H; i += 2) {
								const double2 cc = *reinterpret_cast<const double2*>(cr + c * Apad + i);
								acc[i] = fma(x, cc.x, acc[i]);
								acc[i + 1] = fma(x, cc.y, acc[i + 1]);
							}
						}
					} else {
						for (int c = 0; c < k; ++c) {
							const double x = xt[c * 32];
#pragma unroll
							for (int i = 0; i < H; i += 2) {
								if ((i >> 1) < pairs) {
									const double2 cc = *reinterpret_cast<const double2*>(cr + c * Apad + i);
									acc[i] = fma(x, cc.x, acc[i]);
									acc[i + 1] = fma(x, cc.y, acc[i + 1]);
								}
							}
						}
					}
					const double y = xt[k * 32];
					double st[16];
#pragma unroll
					for (int i = 0; i < 16; ++i) st[i] = 0.0;
#pragma unroll
					for (int i = 0; i < H; ++i) {
						double b0i = 0.0;
#pragma unroll
						for (int j = 0; j < AREG; ++j) if (j == (second ? H : 0) + i) b0i = b0v[j];
						const double e = (lane < rowsHere) ? y - (acc[i] + b0i) : 0.0;
						st[i] = e;
						st[H + i] = e * e;
					}
					// lane L: value (L >> 1): sum e (< H), sum e^2 (H ..); accumulated over the slabs
					const double slabTot = warpSum16(st, lane);
#pragma unroll
					for (int q = 0; q < 4; ++q) if (q == ui) tot[q] += slabTot;
				}
			}
#pragma unroll
			for (int ui = 0; ui < 4; ++ui) {
				const int u = warp + ui * NW;
				if (u >= units) break;
				const int t = firstTask + (u >> 1);
				const int base = (u & 1) ? H : 0;
				const int nrows = (t == 0) ? bd[0].nrows : bd[1].nrows;
				const int kind = (t == 0) ? td[0].kind : td[1].kind;
				const int dest = (t == 0) ? td[0].dest : td[1].dest;
				const int j = lane >> 1;
				double mine = 0.0;
#pragma unroll
				for (int q = 0; q < 4; ++q) if (q == ui) mine = tot[q];
				if ((lane & 1) == 0 && j < 2 * H) {
					const int a = base + (j < H ? j : j - H);
					if (a < A) {
						if (kind == 0) {
							// mean squared error of prediction for this fold (src/PLSEvaluator.cpp:266-268)
							if (j >= H) prm.mse[(static_cast<size_t>(chrom) * prm.innerDests + dest) * prm.maxNComp + a] = mine / static_cast<double>(nrows);
						} else {
							prm.ostat[((static_cast<size_t>(chrom) * prm.outerDests + dest) * prm.maxNComp + a) * 2 + (j >= H ? 1 : 0)] = mine;
						}
					}
				}
			}
		}
		GSB_STAMP();
	} else {
		double* __restrict__ part = P + cv.part;   // [warp][u][lane]
		double* __restrict__ stage = P + cv.stage; // [column][32 rows]
		const int colsPerPass = max(1, min(k, cv.stageCols));
		for (int t = 0; t < fd.taskCount; ++t) {
			const TaskDesc tdt = prm.tasks[fd.taskBegin + t];
			const BlockDesc bdt = prm.blocks[tdt.block];
			const double* __restrict__ zb = prm.zpool + bdt.zOff;
			const double* __restrict__ ycol = zb + static_cast<size_t>(prm.p) * bdt.ld;
			double st[2 * AREG];
#pragma unroll
			for (int u = 0; u < 2 * AREG; ++u) st[u] = 0.0;
			for (int r0 = 0; r0 < bdt.nrows; r0 += 32) {
				const int tr = r0 + lane;
				const bool live = tr < bdt.nrows;
				const int slab = min(32, bdt.ld - r0); // multiple of 4 rows, zero padded in the block copy
				double acc[AREG];
#pragma unroll
				for (int u = 0; u < AREG; ++u) acc[u] = 0.0;
				for (int c0 = 0; c0 < k; c0 += colsPerPass) {
					const int nc = min(colsPerPass, k - c0);
					const int pieces = slab >> 1; // 16-byte pieces per column
					for (int e = tid; e < nc * 16; e += NT) {
						const int c = e >> 4, piece = e & 15;
						if (piece < pieces) cpAsync16(stage + c * 32 + piece * 2, zb + static_cast<size_t>(sel[c0 + c]) * bdt.ld + r0 + piece * 2);
					}
					cpAsyncWaitAll();
					__syncthreads();
					if (lane < slab) {
						for (int c = warp; c < nc; c += NW) {
							const double x = stage[c * 32 + lane];
							const double* cr = coef + (c0 + c) * Apad;
#pragma unroll
							for (int u = 0; u < AREG; u += 2) {
								if (u < Apad) {
									const double2 cc = *reinterpret_cast<const double2*>(cr + u);
									acc[u] = fma(x, cc.x, acc[u]);
									acc[u + 1] = fma(x, cc.y, acc[u + 1]);
								}
							}
						}
					}
					__syncthreads(); // staging area free again
				}
				if (NW > 1) {
#pragma unroll
					for (int u = 0; u < AREG; ++u) part[(warp * kChunk + u) * 32 + lane] = acc[u];
					__syncthreads();
					if (warp == 0) {
#pragma unroll
						for (int u = 0; u < AREG; ++u) {
							double sum = part[u * 32 + lane];
							for (int w2 = 1; w2 < NW; ++w2) sum += part[(w2 * kChunk + u) * 32 + lane];
							acc[u] = sum;
						}
					}
					__syncthreads();
				}
				if (warp == 0 && live) {
					const double y = __ldg(ycol + tr);
#pragma unroll
					for (int u = 0; u < AREG; ++u) {
						const double e = y - (acc[u] + b0v[u]);
						st[u] += e;
						st[AREG + u] = fma(e, e, st[AREG + u]);
					}
				}
			}
			if (warp == 0) {
#pragma unroll
				for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
					for (int u = 0; u < 2 * AREG; ++u) st[u] += shflXor(st[u], m);
				}
				if (lane < A) {
					double se = 0.0, see = 0.0;
#pragma unroll
					for (int u = 0; u < AREG; ++u) {
						if (u == lane) { se = st[u]; see = st[AREG + u]; }
					}
					if (tdt.kind == 0) {
						prm.mse[(static_cast<size_t>(chrom) * prm.innerDests + tdt.dest) * prm.maxNComp + lane] =
							see / static_cast<double>(bdt.nrows);
					} else {
						double* o = prm.ostat + ((static_cast<size_t>(chrom) * prm.outerDests + tdt.dest) * prm.maxNComp + lane) * 2;
						o[0] = se;
						o[1] = see;
					}
				}
			}
			GSB_STAMP(); // one per prediction task
		}
	}
	checkCanary();
#undef GSB_STAMP
}

// rows of the packed Gram that fit in shared memory next to everything else (a multiple of 32, or k)
inline int fastRowsInSmem(int k, int A, int maxTasks, int maxBlockLd, int smemLimit) {
	const int nw = roundUp(k, 32) / 32;
	if (fastCarve(k, A, nw, maxTasks, maxBlockLd, -1, k).total * 8 <= smemLimit) return k;
	int rs = (k / 32) * 32;
	while (rs >= 32 && fastCarve(k, A, nw, maxTasks, maxBlockLd, -1, rs).total * 8 > smemLimit) rs -= 32;
	return rs >= 32 ? rs : 0;
}

// fit-only instances: rows of the packed Gram that fit in shared memory (a multiple of 32, or k)
inline int fitOnlyRowsInSmem(int k, int smemLimit) {
	const int nw = roundUp(k, 32) / 32;
	if (fitOnlyCarve(k, nw, k).total * 8 <= smemLimit) return k;
	int rs = (k / 32) * 32;
	while (rs >= 32 && fitOnlyCarve(k, nw, rs).total * 8 > smemLimit) rs -= 32;
	return rs >= 32 ? rs : 0;
}

template <int AREG, bool PRED>
cudaError_t launchFast(const FitParams& prm, int kMax, int A, int smemLimit, cudaStream_t stream, long long* ctas) {
	const int threads = roundUp(kMax, 32);
	const int rs = PRED ? fastRowsInSmem(kMax, A, prm.maxTasks, prm.maxBlockLd, smemLimit) : fitOnlyRowsInSmem(kMax, smemLimit);
	if (rs <= 0) return cudaErrorInvalidValue;
	const FastCarve cv = PRED ? fastCarve(kMax, A, threads / 32, prm.maxTasks, prm.maxBlockLd, -1, rs) : fitOnlyCarve(kMax, threads / 32, rs);
	const size_t bytes = static_cast<size_t>(cv.total) * sizeof(double) + (prm.guardFail != nullptr ? 64 : 0);
	const long long grid = static_cast<long long>(prm.nfits) * prm.bucketCount;
	if (grid <= 0) return cudaSuccess;
	if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
	cudaError_t err;
	if (rs >= kMax) {
		if (prm.trace != nullptr) {
			err = cudaFuncSetAttribute(simplsFitFast<AREG, false, true, PRED>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
			if (err != cudaSuccess) return err;
			simplsFitFast<AREG, false, true, PRED><<<static_cast<unsigned>(grid), threads, bytes, stream>>>(prm, cv.xtTasks, kMax);
		} else {
			err = cudaFuncSetAttribute(simplsFitFast<AREG, false, false, PRED>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
			if (err != cudaSuccess) return err;
			simplsFitFast<AREG, false, false, PRED><<<static_cast<unsigned>(grid), threads, bytes, stream>>>(prm, cv.xtTasks, kMax);
		}
	} else {
		err = cudaFuncSetAttribute(simplsFitFast<AREG, true, false, PRED>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
		if (err != cudaSuccess) return err;
		simplsFitFast<AREG, true, false, PRED><<<static_cast<unsigned>(grid), threads, bytes, stream>>>(prm, cv.xtTasks, rs);
	}
	if (ctas) *ctas += grid;
	return cudaGetLastError();
}

constexpr int kFastComponents = 10;

template <int RPT, bool GSMEM, bool VGLOB = false>
cudaError_t launchOne(const FitParams& prm, int kMax, int A, int threads, cudaStream_t stream, long long* ctas) {
	const Carve cv = carve(kMax, A, GSMEM, threads / 32, VGLOB);
	const size_t bytes = static_cast<size_t>(cv.total) * sizeof(double);
	cudaError_t err = cudaFuncSetAttribute(simplsFitKernel<RPT, GSMEM, VGLOB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                                       static_cast<int>(bytes));
	if (err != cudaSuccess) return err;
	const long long grid = static_cast<long long>(prm.nfits) * prm.bucketCount;
	if (grid <= 0) return cudaSuccess;
	if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
	simplsFitKernel<RPT, GSMEM, VGLOB><<<static_cast<unsigned>(grid), threads, bytes, stream>>>(prm);
	if (ctas) *ctas += grid;
	return cudaGetLastError();
}

// doubles of global scratch one CTA of the vGlobal instances needs for a subset of k columns and A components
inline unsigned long long genScratchPerCta(int k, int A) {
	return static_cast<unsigned long long>(roundUp(A * k, 2)) + static_cast<unsigned long long>(k) * roundUp(A, kChunk) + 2ull;
}

// the vGlobal instances, the chromosomes of the class walked in chunks that fit the scratch the engine provides
template <int RPT, bool GSMEM>
cudaError_t launchGlobalV(const FitParams& prm, int kMax, int A, int threads, cudaStream_t stream, long long* ctas) {
	const unsigned long long per = genScratchPerCta(kMax, A);
	if (prm.genScratch == nullptr || prm.genCapacity < per * static_cast<unsigned long long>(prm.nfits)) return cudaErrorMemoryAllocation;
	const int chunk = static_cast<int>(std::min<unsigned long long>(static_cast<unsigned long long>(prm.bucketCount),
	                                                                   prm.genCapacity / (per * static_cast<unsigned long long>(prm.nfits))));
	for (int c0 = 0; c0 < prm.bucketCount; c0 += chunk) {
		FitParams sub = prm;
		sub.bucket = prm.bucket + c0;
		sub.bucketCount = std::min(chunk, prm.bucketCount - c0);
		sub.genStride = per;
		const cudaError_t err = launchOne<RPT, GSMEM, true>(sub, kMax, A, threads, stream, ctas); // same stream: the chunks reuse the scratch in order
		if (err != cudaSuccess) return err;
	}
	return cudaSuccess;
}

} // namespace

// doubles of global scratch per CTA the per-fit kernel needs for a class of subsets up to kMax columns with A components
// (0: everything fits in shared memory)
unsigned long long fitKernelScratchPerCta(int kMax, int A, int maxTasks, int maxBlockLd, int smemLimit) {
	const int Ak = (A < kMax) ? A : kMax;
	if (Ak <= kFastComponents && kMax <= kMaxWarps * 32 && fastRowsInSmem(kMax, Ak, maxTasks, maxBlockLd, smemLimit) > 0) return 0ull;
	if (kMax <= kMaxWarps * 32 && carve(kMax, Ak, true, roundUp(kMax, 32) / 32).total * 8 <= smemLimit) return 0ull;
	return genScratchPerCta(kMax, Ak);
}

int fitKernelSharedBytes(int k, int A) { return carve(k, A, true, roundUp(k, 32) / 32).total * static_cast<int>(sizeof(double)); }

// largest k whose packed Gram still fits in shared memory next to the loadings
int fitKernelMaxK(int A, int smemLimit) {
	int k = 1;
	while (k < kMaxWarps * 32 && carve(k + 1, (A < k + 1) ? A : k + 1, true, roundUp(k + 1, 32) / 32).total * 8 <= smemLimit) ++k;
	return k;
}

// One launch for a bucket of chromosomes whose subset sizes are all <= kMax.
cudaError_t launchFitKernel(const FitParams& prm, int kMax, int A, int smemLimit, cudaStream_t stream, long long* ctas) {
	const int Ak = (A < kMax) ? A : kMax;
	if (Ak <= kFastComponents && kMax <= kMaxWarps * 32 && fastRowsInSmem(kMax, Ak, prm.maxTasks, prm.maxBlockLd, smemLimit) > 0) {
		return launchFast<kFastComponents, true>(prm, kMax, Ak, smemLimit, stream, ctas); // per-row state in registers
	}
	if (kMax <= kMaxWarps * 32 && carve(kMax, Ak, true, roundUp(kMax, 32) / 32).total * 8 <= smemLimit) {
		return launchOne<1, true>(prm, kMax, Ak, roundUp(kMax, 32), stream, ctas); // one thread per row
	}
	// the stored loadings and coefficients no longer fit beside the Gram: they go to the global scratch before the Gram
	// does (measured at k = 200, 20 components: Gram in place 93 ms per 74 chromosomes, loadings in the scratch far less)
	if (kMax <= kMaxWarps * 32 && prm.genScratch != nullptr && carve(kMax, Ak, true, roundUp(kMax, 32) / 32, true).total * 8 <= smemLimit) {
		return launchGlobalV<1, true>(prm, kMax, Ak, roundUp(kMax, 32), stream, ctas);
	}
	if (kMax <= 4 * kMaxWarps * 32 && carve(kMax, Ak, false, kMaxWarps).total * 8 <= smemLimit) {
		return launchOne<4, false>(prm, kMax, Ak, kMaxWarps * 32, stream, ctas);     // Gram read in place (L2)
	}
	if (kMax > 4 * kMaxWarps * 32 || carve(kMax, Ak, false, kMaxWarps, true).total * 8 > smemLimit) return cudaErrorInvalidValue;
	return launchGlobalV<4, false>(prm, kMax, Ak, kMaxWarps * 32, stream, ctas);
}

// The fit-only instance of the fast path (A <= 10 components, k <= 384) serves this class: its coefficients go to
// predictBlocksKernel
bool fitOnlyKernelServes(int kMax, int A, int smemLimit) {
	const int Ak = (A < kMax) ? A : kMax;
	return Ak <= kFastComponents && kMax <= kMaxWarps * 32 && fitOnlyRowsInSmem(kMax, smemLimit) > 0;
}

cudaError_t launchFitOnlyKernel(const FitParams& prm, int kMax, int A, int smemLimit, cudaStream_t stream, long long* ctas) {
	const int Ak = (A < kMax) ? A : kMax;
	if (!fitOnlyKernelServes(kMax, A, smemLimit) || prm.coefScratch == nullptr) return cudaErrorInvalidValue;
	return launchFast<kFastComponents, false>(prm, kMax, Ak, smemLimit, stream, ctas);
}

} // namespace gsb
