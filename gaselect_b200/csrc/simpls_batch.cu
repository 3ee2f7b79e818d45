// simpls_batch.cu -- K1b, the replication-batched SIMPLS kernel: one CTA per (chromosome, group of fits).
//
// Same arithmetic as simpls_fit.cu (PLSSimpls::fit in Gram space, src/PLSSimpls.cpp:106-166; PLS::predict,
// src/PLS.cpp:34-47; MSEP / residual sums, src/PLSEvaluator.cpp:263-268,323-325), organised around what the
// <= 16 training sets of one CV replication have in common: every training Gram is the FULL-data Gram minus
// the Gram of the 1-2 excluded row blocks (this is how K0 derives them, too).  So instead of gathering 15
// different k x k Grams and running 15 matrix-vector products per component, the CTA gathers the selected
// entries of the full-data Gram ONCE and computes, for all fits f of the group at the same time,
//
//     W[:, f] = M S[:, f]  -  sum_{b excluded from f} Z_b' (Z_b S[:, f])  -  n_f xm_f (xm_f' S[:, f])
//
// M = Z'Z over all rows (Z globally centred), Z_b = rows of block b, xm_f = training mean relative to the
// global centre.  The first term is a k x k by k x NF product: each Gram entry read from shared memory feeds NF
// FMAs (the per-fit kernel reads one entry per FMA and is bound by shared-memory bandwidth).  The second term
// costs 2 |b| k per (block, fit) pair, and its inner factor U = Z_b S[:, f] is exactly the un-normalised score of
// the held-out rows: the predictions of the block by fit f come for free,
//
//     yhat_a = yhat_{a-1} + (U - xm_f'S) q / tn          (coefficients are cumulative, src/PLSSimpls.cpp:152-156)
//
// so no coefficient matrix is kept at all.  Per component: 6 CTA barriers, phases
//   A  U = Z_b S for every (block, fit) pair (a warp per 32-row slab, lanes = rows);  xm'S, s0'S partial sums
//   B  W = M S (one thread per Gram row and FPT fits, packed triangle, conflict-free)
//   C  W -= Z_b' U (one thread per column, walking its column of the block);  W -= n xm (xm'S)
//   D  S'W and V_j'W for every stored loading (thread = (fit, row stripe), warp-level partial sums)
//   E  v = (W - V c) / tn;  v'v, v'S partial sums
//   F  deflate S, store the new loading;  update the held-out predictions and write sum e, sum e^2
// The stored loadings and the per-fit s0 / xm columns live in a per-CTA global scratch (L2-resident while the
// CTA runs); shared memory holds the packed Gram, S, W, U and the partial sums.
#include "kernels.cuh"

namespace gsb {

namespace {

constexpr int kBT = 256;             // threads per CTA
constexpr int kBW = kBT / 32;        // warps
constexpr int kBA = kBatchComponents;
constexpr int kPG = 8;               // most pairs per pair-group

__device__ __forceinline__ double shflXorD(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

__device__ __forceinline__ void cpAsync8b(double* dst, const double* src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src));
}
__device__ __forceinline__ void cpAsyncWaitAllB() {
	asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

__host__ __device__ inline int roundUpB(int v, int m) { return (v + m - 1) / m * m; }

// shared-memory carve-up (in doubles)
struct BatchCarve {
	int P, S, W, U, partA, partD, partE, scal, sel, total;
};

constexpr int kScalRows = 8; // ntr, ym, m, rtn, fac, under, (spare)

__host__ __device__ inline BatchCarve batchCarve(int k, int NF, int uDoubles) {
	BatchCarve c;
	int o = 0;
	c.P = o; o += roundUpB(k * (k + 1) / 2, 2);
	c.S = o; o += k * NF;
	c.W = o; o += k * (NF + 2);
	c.U = o; o += roundUpB(uDoubles, 2);
	c.partA = o; o += kBW * 2 * NF;
	c.partD = o; o += kBW * (kBA + 1) * NF;
	c.partE = o; o += kBW * 2 * NF;
	c.scal = o; o += kScalRows * NF;
	c.sel = o; o += roundUpB(k, 2) / 2 + 1;
	c.total = o;
	return c;
}

// Sum 16 values per lane over the warp: afterwards every lane L holds the total of value (L >> 1).
__device__ __forceinline__ double warpSum16b(double (&v)[16], int lane) {
#pragma unroll
	for (int h = 8; h >= 1; h >>= 1) {
		const bool up = (lane & (2 * h)) != 0;
#pragma unroll
		for (int i = 0; i < h; ++i) {
			const double send = up ? v[i] : v[i + h];
			const double keep = up ? v[i + h] : v[i];
			v[i] = keep + shflXorD(send, 2 * h);
		}
	}
	return v[0] + shflXorD(v[0], 1);
}

// acc[0..FPT) += g * Srow[0..FPT)
template <int FPT>
__device__ __forceinline__ void fmaRow(double (&acc)[FPT], double g, const double* __restrict__ srow) {
#pragma unroll
	for (int f = 0; f < FPT; f += 2) {
		const double2 s = *reinterpret_cast<const double2*>(srow + f);
		acc[f] = fma(g, s.x, acc[f]);
		acc[f + 1] = fma(g, s.y, acc[f + 1]);
	}
}

// One row of W = M S for FPT fits from the packed triangle in shared memory.  Thread r owns row r; the rows of a
// warp are consecutive, [rlo, rhi] is warp-uniform: left of the 32-wide diagonal band (own packed row), the band
// (either side of the diagonal), below the band (column r of the rows c; lane = row, conflict-free).
template <int NF, int FPT>
__device__ __forceinline__ void symvMulti(const double* __restrict__ P, const double* __restrict__ Sg, int k, int r, int rlo,
                                          int rhi, double (&acc)[FPT]) {
	const double* __restrict__ pr = P + r * (r + 1) / 2;
	const double* __restrict__ pc = P + r;
	int c = 0;
#pragma unroll 2
	for (; c < rlo; ++c) fmaRow<FPT>(acc, pr[c], Sg + c * NF);
	int tc = rlo * (rlo + 1) / 2;
	for (; c <= rhi; ++c) {
		const double g = (c <= r) ? pr[c] : pc[tc];
		fmaRow<FPT>(acc, g, Sg + c * NF);
		tc += c + 1;
	}
#pragma unroll 2
	for (; c < k; ++c) {
		fmaRow<FPT>(acc, pc[tc], Sg + c * NF);
		tc += c + 1;
	}
}

// phase A for one (pair-group, slab) unit with NP pairs: U[row][q] = sum_c Z_b[row, sel_c] S[c, fit_q]
template <int NF, int NP>
__device__ __forceinline__ void scoreUnit(const double* __restrict__ zrow, bool live, int ld, const uint32_t* __restrict__ sel,
                                          const double* __restrict__ S, int k, const BatchPair* __restrict__ pairs,
                                          double* __restrict__ urow) {
	int fo[NP];
#pragma unroll
	for (int q = 0; q < NP; ++q) fo[q] = pairs[q].fitLocal;
	double acc[NP];
#pragma unroll
	for (int q = 0; q < NP; ++q) acc[q] = 0.0;
#pragma unroll 8
	for (int c = 0; c < k; ++c) {
		const double z = live ? __ldg(zrow + static_cast<size_t>(sel[c]) * ld) : 0.0;
		const double* __restrict__ sc = S + c * NF;
#pragma unroll
		for (int q = 0; q < NP; ++q) acc[q] = fma(z, sc[fo[q]], acc[q]);
	}
#pragma unroll
	for (int q = 0; q < NP; ++q) urow[q] = acc[q];
}

// phase C for one pair-group with NP pairs: acc[q] = sum_r Z_b[r, sel_i] U[r][q]
template <int NP>
__device__ __forceinline__ void corrGroup(const double* __restrict__ zcol, int ld, const double* __restrict__ Ub, int pgp,
                                          double (&acc)[kPG]) {
#pragma unroll 2
	for (int r = 0; r < ld; r += 2) {
		const double2 z = __ldg(reinterpret_cast<const double2*>(zcol + r));
		const double* __restrict__ u0 = Ub + r * pgp;
		const double* __restrict__ u1 = u0 + pgp;
#pragma unroll
		for (int q = 0; q < NP; ++q) {
			acc[q] = fma(z.x, u0[q], acc[q]);
			acc[q] = fma(z.y, u1[q], acc[q]);
		}
	}
}

template <int NF, int FPT>
__global__ void __launch_bounds__(kBT, 2) simplsBatchKernel(const BatchParams prm) {
	extern __shared__ __align__(16) double smem[];
	static_assert(NF == 8 || NF == 16, "fits per group (padded)");
	static_assert(NF % FPT == 0 && FPT % 2 == 0, "fits per thread");
	constexpr int LDW = NF + 2;
	constexpr int NS = kBT / NF; // row stripes of the (fit, stripe) thread mapping
	constexpr int NV = kBA + 1;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int ci = blockIdx.x % prm.bucketCount;
	const int gi = blockIdx.x / prm.bucketCount;
	const int chrom = prm.bucket[ci];
	const uint32_t selBegin = prm.offsets[chrom];
	const int k = static_cast<int>(prm.offsets[chrom + 1] - selBegin);
	const int A = min(prm.maxNComp, k);
	const BatchGroup grp = prm.groups[gi];
	const int PGP = prm.pgp;

	const BatchCarve cv = batchCarve(k, NF, prm.uDoubles);
	double* __restrict__ P = smem + cv.P;
	double* __restrict__ S = smem + cv.S;
	double* __restrict__ W = smem + cv.W;
	double* __restrict__ U = smem + cv.U;
	double* __restrict__ partA = smem + cv.partA;
	double* __restrict__ partD = smem + cv.partD;
	double* __restrict__ partE = smem + cv.partE;
	double* __restrict__ scal = smem + cv.scal;
	uint32_t* __restrict__ sel = reinterpret_cast<uint32_t*>(smem + cv.sel);
	double* __restrict__ scNtr = scal, * __restrict__ scYm = scal + NF, * __restrict__ scM = scal + 2 * NF;
	double* __restrict__ scFac = scal + 4 * NF;

	double* scr = prm.scratch + static_cast<size_t>(blockIdx.x) * prm.scratchStride;
	double* Vg = scr;                                  // [comp][row][fit]
	double* s0g = scr + static_cast<size_t>(kBA) * k * NF; // [row][fit]
	double* xmg = s0g + k * NF;
	double* predg = xmg + k * NF;                      // [slab][pair][lane]

	// ---- gather -------------------------------------------------------------------------------------------
	for (int i = tid; i < k; i += kBT) sel[i] = prm.idx[selBegin + i];
	__syncthreads();
	{
		const double* __restrict__ gsrc = prm.gfull;
		uint32_t selB[kBW];
#pragma unroll
		for (int j = 0; j < kBW; ++j) selB[j] = (lane + 32 * j < k) ? sel[lane + 32 * j] : 0u;
		for (int a = warp; a < k; a += kBW) {
			const double* __restrict__ grow = gsrc + tri(sel[a]);
			double* __restrict__ prow = P + a * (a + 1) / 2;
#pragma unroll
			for (int j = 0; j < kBW; ++j) {
				if (32 * j <= a) {
					const int b = lane + 32 * j;
					if (b <= a) cpAsync8b(prow + b, grow + selB[j]);
				}
			}
		}
	}
	for (int e = tid; e < k * NF; e += kBT) {
		const int i = e / NF, fl = e % NF;
		double s = 0.0, x = 0.0;
		if (fl < grp.nfits) {
			const size_t at = static_cast<size_t>(prm.groupFits[grp.fitBegin + fl]) * prm.p + sel[i];
			s = __ldg(prm.s0 + at);
			x = __ldg(prm.xm + at);
		}
		S[e] = s;
		s0g[e] = s;
		xmg[e] = x;
	}
	for (int e = tid; e < prm.uDoubles; e += kBT) predg[e] = 0.0;
	if (tid < NF) {
		double ntr = 0.0, ym = 0.0;
		if (tid < grp.nfits) {
			const FitDesc fd = prm.fits[prm.groupFits[grp.fitBegin + tid]];
			ntr = static_cast<double>(fd.ntr);
			ym = fd.ym;
		}
		scNtr[tid] = ntr;
		scYm[tid] = ym;
	}
	cpAsyncWaitAllB();
	__syncthreads();

	const int f = tid % NF, stripe = tid / NF;
	const bool active = f < grp.nfits;
	const int KP = roundUpB(k, 32);
	const double nanv = __longlong_as_double(0x7ff8000000000000LL);

#pragma unroll 1
	for (int comp = 0; comp < A; ++comp) {
		// ---- A: scores of the block rows; xm'S and s0'S ---------------------------------------------------
		for (int u = warp; u < grp.nunits; u += kBW) {
			const int2 ud = prm.units[grp.unitBegin + u];
			const BatchPairGroup pg = prm.pgs[ud.x];
			const BlockDesc bd = prm.blocks[pg.block];
			const int row = ud.y * 32 + lane;
			const bool live = row < bd.ld;
			const double* __restrict__ zrow = prm.zpool + bd.zOff + (live ? row : 0);
			const BatchPair* __restrict__ pr = prm.pairs + pg.pairBegin;
			double* __restrict__ urow = U + (static_cast<size_t>(pg.slabBegin + ud.y) * 32 + lane) * PGP;
			switch (pg.npairs) {
			case 1: scoreUnit<NF, 1>(zrow, live, bd.ld, sel, S, k, pr, urow); break;
			case 2: scoreUnit<NF, 2>(zrow, live, bd.ld, sel, S, k, pr, urow); break;
			case 3: scoreUnit<NF, 3>(zrow, live, bd.ld, sel, S, k, pr, urow); break;
			case 4: scoreUnit<NF, 4>(zrow, live, bd.ld, sel, S, k, pr, urow); break;
			case 5: scoreUnit<NF, 5>(zrow, live, bd.ld, sel, S, k, pr, urow); break;
			case 6: scoreUnit<NF, 6>(zrow, live, bd.ld, sel, S, k, pr, urow); break;
			case 7: scoreUnit<NF, 7>(zrow, live, bd.ld, sel, S, k, pr, urow); break;
			default: scoreUnit<NF, 8>(zrow, live, bd.ld, sel, S, k, pr, urow); break;
			}
		}
		{
			double am = 0.0, aq = 0.0;
			for (int i = stripe; i < k; i += NS) {
				const double s = S[i * NF + f];
				am = fma(xmg[i * NF + f], s, am);
				aq = fma(s0g[i * NF + f], s, aq);
			}
#pragma unroll
			for (int m = 16; m >= NF; m >>= 1) {
				am += shflXorD(am, m);
				aq += shflXorD(aq, m);
			}
			if (lane < NF) {
				partA[(warp * 2 + 0) * NF + f] = am;
				partA[(warp * 2 + 1) * NF + f] = aq;
			}
		}
		__syncthreads();

		// ---- B: W = M S -------------------------------------------------------------------------------------
		if (tid < NF) {
			double m = partA[tid];
			for (int w = 1; w < kBW; ++w) m += partA[(w * 2) * NF + tid];
			scM[tid] = m;
		}
		if (tid < KP * (NF / FPT)) {
			const int i = tid % KP, g = tid / KP;
			const int r = min(i, k - 1);
			const int rlo = min((i >> 5) * 32, k - 1);
			const int rhi = min(rlo + 31, k - 1);
			double acc[FPT];
#pragma unroll
			for (int q = 0; q < FPT; ++q) acc[q] = 0.0;
			symvMulti<NF, FPT>(P, S + g * FPT, k, r, rlo, rhi, acc);
			if (i < k) {
#pragma unroll
				for (int q = 0; q < FPT; q += 2) *reinterpret_cast<double2*>(W + i * LDW + g * FPT + q) = make_double2(acc[q], acc[q + 1]);
			}
		}
		__syncthreads();

		// ---- C: W -= Z_b' U for the excluded blocks; centring term -----------------------------------------
		if (tid < k) {
			const int i = tid;
			const size_t col = sel[i];
			double* __restrict__ wr = W + i * LDW;
			for (int pgi = 0; pgi < grp.npg; ++pgi) {
				const BatchPairGroup pg = prm.pgs[grp.pgBegin + pgi];
				if (pg.ncorr == 0) continue;
				const BlockDesc bd = prm.blocks[pg.block];
				const double* __restrict__ zcol = prm.zpool + bd.zOff + col * bd.ld;
				const double* __restrict__ Ub = U + static_cast<size_t>(pg.slabBegin) * 32 * PGP;
				double acc[kPG];
#pragma unroll
				for (int q = 0; q < kPG; ++q) acc[q] = 0.0;
				switch (pg.npairs) {
				case 1: corrGroup<1>(zcol, bd.ld, Ub, PGP, acc); break;
				case 2: corrGroup<2>(zcol, bd.ld, Ub, PGP, acc); break;
				case 3: corrGroup<3>(zcol, bd.ld, Ub, PGP, acc); break;
				case 4: corrGroup<4>(zcol, bd.ld, Ub, PGP, acc); break;
				case 5: corrGroup<5>(zcol, bd.ld, Ub, PGP, acc); break;
				case 6: corrGroup<6>(zcol, bd.ld, Ub, PGP, acc); break;
				case 7: corrGroup<7>(zcol, bd.ld, Ub, PGP, acc); break;
				default: corrGroup<8>(zcol, bd.ld, Ub, PGP, acc); break;
				}
				const BatchPair* __restrict__ pr = prm.pairs + pg.pairBegin;
#pragma unroll
				for (int q = 0; q < kPG; ++q) {
					if (q < pg.npairs) {
						const BatchPair bp = pr[q];
						if (bp.corr) wr[bp.fitLocal] -= acc[q];
					}
				}
			}
			const double* xr = xmg + i * NF;
#pragma unroll
			for (int q = 0; q < NF; q += 2) {
				const double2 x = *reinterpret_cast<const double2*>(xr + q);
				double2 w = *reinterpret_cast<double2*>(wr + q);
				w.x = fma(-scNtr[q] * x.x, scM[q], w.x);
				w.y = fma(-scNtr[q + 1] * x.y, scM[q + 1], w.y);
				*reinterpret_cast<double2*>(wr + q) = w;
			}
		}
		__syncthreads();

		// ---- D: tn^2 = S'W (:138) and the projections V_j'W on every stored loading ---------------------------
		{
			double d[NV];
#pragma unroll
			for (int v = 0; v < NV; ++v) d[v] = 0.0;
			for (int i = stripe; i < k; i += NS) {
				const double w = W[i * LDW + f];
				d[0] = fma(S[i * NF + f], w, d[0]);
#pragma unroll
				for (int j = 0; j < kBA; ++j) {
					if (j < comp) d[1 + j] = fma(Vg[(static_cast<size_t>(j) * k + i) * NF + f], w, d[1 + j]);
				}
			}
#pragma unroll
			for (int m = 16; m >= NF; m >>= 1) {
#pragma unroll
				for (int v = 0; v < NV; ++v) d[v] += shflXorD(d[v], m);
			}
			if (lane < NF) {
#pragma unroll
				for (int v = 0; v < NV; ++v) if (v <= comp) partD[(warp * NV + v) * NF + f] = d[v];
			}
		}
		__syncthreads();

		// ---- E: v = (W - V c) / tn; v'v and v'S ------------------------------------------------------------
		bool under;
		{
			double tn2 = partD[f], qtn = partA[NF + f];
			for (int w = 1; w < kBW; ++w) {
				tn2 += partD[(w * NV) * NF + f];
				qtn += partA[(w * 2 + 1) * NF + f];
			}
			double c[kBA];
#pragma unroll
			for (int j = 0; j < kBA; ++j) {
				c[j] = 0.0;
				if (j < comp) {
					double t = partD[(1 + j) * NF + f];
					for (int w = 1; w < kBW; ++w) t += partD[(w * NV + 1 + j) * NF + f];
					c[j] = t;
				}
			}
			under = active && !(tn2 >= 1e-50); // ||t|| < 1e-25 (NORM_TOL, :16,:141-143); also NaN / negative rounding
			const double rtn = active ? rsqrt(tn2) : 0.0;
			const double q = qtn * rtn;
			if (stripe == 0) scFac[f] = under ? nanv : q * rtn; // coefficient increment is S q / tn (:152-156)
			double vv = 0.0, vs = 0.0;
			for (int i = stripe; i < k; i += NS) {
				double v = W[i * LDW + f];
#pragma unroll
				for (int j = 0; j < kBA; ++j) {
					if (j < comp) v = fma(-Vg[(static_cast<size_t>(j) * k + i) * NF + f], c[j], v);
				}
				v *= rtn;
				W[i * LDW + f] = v;
				vv = fma(v, v, vv);
				vs = fma(v, S[i * NF + f], vs);
			}
#pragma unroll
			for (int m = 16; m >= NF; m >>= 1) {
				vv += shflXorD(vv, m);
				vs += shflXorD(vs, m);
			}
			if (lane < NF) {
				partE[(warp * 2 + 0) * NF + f] = vv;
				partE[(warp * 2 + 1) * NF + f] = vs;
			}
		}
		__syncthreads();

		// ---- F: deflate S, store the loading (:160, :70-88); predictions of the held-out rows ---------------
		{
			double vv = partE[f], vs = partE[NF + f];
			for (int w = 1; w < kBW; ++w) {
				vv += partE[(w * 2) * NF + f];
				vs += partE[(w * 2 + 1) * NF + f];
			}
			const double rvn = active ? rsqrt(vv) : 0.0;
			const double dd = active ? (vs * rvn) * rvn : 0.0;
			double* Vn = Vg + static_cast<size_t>(comp) * k * NF;
			const bool keep = comp + 1 < A;
			for (int i = stripe; i < k; i += NS) {
				const double v = W[i * LDW + f];
				double s = fma(-v, dd, S[i * NF + f]);
				if (under) s = nanv; // the reference throws: poison this and every later component
				S[i * NF + f] = s;
				if (keep) Vn[i * NF + f] = v * rvn;
			}
		}
		for (int pgi = warp; pgi < grp.npg; pgi += kBW) {
			const BatchPairGroup pg = prm.pgs[grp.pgBegin + pgi];
			const BlockDesc bd = prm.blocks[pg.block];
			const BatchPair* __restrict__ pr = prm.pairs + pg.pairBegin;
			const double* __restrict__ ycol = prm.zpool + bd.zOff + static_cast<size_t>(prm.p) * bd.ld;
			int fo[kPG], kind[kPG];
#pragma unroll
			for (int q = 0; q < kPG; ++q) {
				fo[q] = (q < pg.npairs) ? pr[q].fitLocal : 0;
				kind[q] = (q < pg.npairs) ? pr[q].kind : -1;
			}
			double st[16];
#pragma unroll
			for (int q = 0; q < 16; ++q) st[q] = 0.0;
			for (int slab = 0; slab < pg.nslabs; ++slab) {
				const int row = slab * 32 + lane;
				const bool live = row < bd.nrows;
				const double y = live ? __ldg(ycol + row) : 0.0;
				const double* __restrict__ urow = U + (static_cast<size_t>(pg.slabBegin + slab) * 32 + lane) * PGP;
				double* pslab = predg + static_cast<size_t>(pg.slabBegin + slab) * PGP * 32 + lane;
#pragma unroll
				for (int q = 0; q < kPG; ++q) {
					if (kind[q] >= 0) {
						const double pe = fma(urow[q] - scM[fo[q]], scFac[fo[q]], pslab[q * 32]);
						pslab[q * 32] = pe;
						const double e = live ? y - (scYm[fo[q]] + pe) : 0.0;
						st[q] += e;
						st[8 + q] = fma(e, e, st[8 + q]);
					}
				}
			}
			const double tot = warpSum16b(st, lane);
			const int v = lane >> 1, q = v & 7;
			if ((lane & 1) == 0 && q < pg.npairs) {
				const BatchPair bp = pr[q];
				if (bp.kind == 0) {
					// mean squared error of prediction for this fold (src/PLSEvaluator.cpp:266-268)
					if (v >= 8) prm.mse[(static_cast<size_t>(chrom) * prm.innerDests + bp.dest) * prm.maxNComp + comp] = tot / static_cast<double>(bd.nrows);
				} else if (bp.kind == 1) {
					prm.ostat[((static_cast<size_t>(chrom) * prm.outerDests + bp.dest) * prm.maxNComp + comp) * 2 + (v >= 8 ? 1 : 0)] = tot;
				}
			}
		}
		__syncthreads();
	}
}

template <int NF, int FPT>
cudaError_t launchBatchOne(const BatchParams& prm, int kMax, cudaStream_t stream, long long* ctas) {
	const BatchCarve cv = batchCarve(kMax, NF, prm.uDoubles);
	const size_t bytes = static_cast<size_t>(cv.total) * sizeof(double);
	const long long grid = static_cast<long long>(prm.ngroups) * prm.bucketCount;
	if (grid <= 0) return cudaSuccess;
	if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
	cudaError_t err = cudaFuncSetAttribute(simplsBatchKernel<NF, FPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
	if (err != cudaSuccess) return err;
	simplsBatchKernel<NF, FPT><<<static_cast<unsigned>(grid), kBT, bytes, stream>>>(prm);
	if (ctas) *ctas += grid;
	return cudaGetLastError();
}

inline int batchNF(int nfMax) { return nfMax <= 8 ? 8 : 16; }

} // namespace

// largest subset size the batched kernel takes (packed Gram + S + W + U in shared memory)
int batchKernelMaxK(int nfMax, int uDoubles, int smemLimit) {
	if (nfMax < 1 || nfMax > 16) return 0;
	const int NF = batchNF(nfMax);
	int k = 0;
	while (k < kBT && batchCarve(k + 1, NF, uDoubles).total * 8 <= smemLimit) ++k;
	return k;
}

long long batchKernelScratchDoubles(int kMax, int nfMax, int uDoubles) {
	const int NF = batchNF(nfMax);
	return roundUpB((kBA + 2) * kMax * NF + uDoubles, 2);
}

cudaError_t launchBatchKernel(const BatchParams& prm, int kMax, int A, int smemLimit, cudaStream_t stream, long long* ctas) {
	if (A > kBA || kMax > batchKernelMaxK(prm.nfMax, prm.uDoubles, smemLimit) || prm.pgp > kPG || (prm.pgp & 1)) return cudaErrorInvalidValue;
	if (batchNF(prm.nfMax) == 16) {
		if (kMax > 128) return launchBatchOne<16, 16>(prm, kMax, stream, ctas);
		if (kMax > 64) return launchBatchOne<16, 8>(prm, kMax, stream, ctas);
		return launchBatchOne<16, 4>(prm, kMax, stream, ctas);
	}
	if (kMax > 128) return launchBatchOne<8, 8>(prm, kMax, stream, ctas);
	if (kMax > 64) return launchBatchOne<8, 4>(prm, kMax, stream, ctas);
	return launchBatchOne<8, 2>(prm, kMax, stream, ctas);
}

} // namespace gsb
