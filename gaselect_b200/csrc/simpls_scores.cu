// simpls_scores.cu -- K1x: score-space SIMPLS on the FP64 tensor cores, 8 fits of one CV replication per CTA.
//
// What it replaces per CTA (same as simpls_fit.cu): PLS::viewSelectColumns / viewSelectRows (src/PLS.cpp:13-23),
// PLSSimpls::fit (src/PLSSimpls.cpp:106-166), PLS::predict (src/PLS.cpp:34-47) and the MSEP / residual sums of
// PLSEvaluator::estSEP (src/PLSEvaluator.cpp:263-268, 323-325) -- for EIGHT training sets at a time.
//
// In a nested CV whose inner segments coincide with the outer ones (BASELINE configs 3 and 5: I = O - 1, n divisible
// by O) the 15 distinct training sets of a replication are unions of the same 5 row blocks ("atoms").  The CTA keeps
// the selected columns of ALL rows in shared memory once, rows in atom order (Zs, n x k, gathered with 16-byte
// cp.async from a row-permuted copy of Z), and runs the reference's own X-space recurrences with the 8 fits as the
// N dimension of two FP64 tensor-core products per component (mma.sync m8n8k4 -> DMMA; tcgen05 has no FP64 kind):
//
//   P1  T  = Zs S              scores of every row for every fit              (n x k) . (k x 8)      DMMA
//   P2  one warp per fit: mask to the fit's training rows, centre (mean over the training rows: this IS the
//       column-centring of src/PLSSimpls.cpp:28-49 applied to t = X S), tn = |t| (:138), q = y't / tn (:148);
//       the masked-out rows are the held-out rows: yhat += (t - mean) q / tn (cumulative coefficients, :152-156),
//       residual sums per prediction task leave the kernel here
//   P3  Wv = Zs' (t / tn)      loadings v = X't (:145-147)                    (k x n) . (n x 8)      DMMA
//   P4  one warp per fit: project v on the stored orthonormal loadings, deflate S (:57-63, :70-88, :160)
//
// 4 CTA barriers per component; everything between them is either a tensor-core product out of shared memory or
// warp-local.  No Gram, no per-fit tables: the only inputs are Z (with y as column p) and the atom structure.
// Subsets whose columns do not fit one CTA's shared memory next to the 9 stored loadings are split over a cluster
// of two CTAs by columns: P1 partial products are exchanged through distributed shared memory, P3 is local, the
// P4 dot products are exchanged as 11 partial sums per fit (3 cluster barriers per component).
//
// Numerics: as in simpls_fit.cu the projections on the stored loadings are taken all at once; sums run in a fixed
// order (bitwise reproducible, independent of the batch composition).
#include "kernels.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace gsb {

namespace {

constexpr int kXT = 256;                  // threads per CTA: one warp per fit
constexpr int kXF = 8;                    // fits per CTA
constexpr int kXA = kScoreComponents;
constexpr int kXRU = kScoreMaxRows / 32;  // rows per lane
constexpr int kXCU = 4;                   // local columns per lane (<= 128 local columns)

__host__ __device__ inline int roundUpX(int v, int m) { return (v + m - 1) / m * m; }

__device__ __forceinline__ void dmma884x(double& d0, double& d1, double a, double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cpAsync16x(double* dst, const double* src) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src));
}
__device__ __forceinline__ void cpAsyncWaitAllX() {
	asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ double shflXorX(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double warpSumX(double v) {
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1) v += shflXorX(v, m);
	return v;
}
// Sum 16 values per lane over the warp: afterwards every lane L holds the total of value (L >> 1).
__device__ __forceinline__ double warpSum16x(double (&v)[16], int lane) {
#pragma unroll
	for (int h = 8; h >= 1; h >>= 1) {
		const bool up = (lane & (2 * h)) != 0;
#pragma unroll
		for (int i = 0; i < h; ++i) {
			const double send = up ? v[i] : v[i + h];
			const double keep = up ? v[i + h] : v[i];
			v[i] = keep + shflXorX(send, 2 * h);
		}
	}
	return v[0] + shflXorX(v[0], 1);
}

// shared-memory carve-up (in doubles).  Leading dimensions are = 4 or 12 (mod 16): the 8x4 / 4x8 tensor-core
// fragments (lane -> 4 * (lane / 4) + lane % 4 or 4 * (lane % 4) + lane / 4 rows/columns) are then bank-conflict free.
struct XCarve {
	int klp, ldz, lds, ldt;
	int Z, S, Wv, V, TA, TB, xch, sel, total;
};

__host__ __device__ inline XCarve xCarve(int kLocal, int nr) {
	XCarve c;
	c.klp = roundUpX(kLocal > 0 ? kLocal : 1, 8);
	c.ldz = nr + 4;
	c.ldt = nr + 4;
	c.lds = c.klp + 4;
	int o = 0;
	c.Z = o; o += c.klp * c.ldz;              // [column][row]
	c.S = o; o += kXF * c.lds;                // [fit][column]
	c.Wv = o; o += kXF * c.lds;
	c.V = o; o += (kXA - 1) * kXF * c.klp;    // [loading][fit][column]
	c.TA = o; o += kXF * c.ldt;               // [fit][row]
	c.TB = o; o += kXF * (c.ldt > c.lds ? c.ldt : c.lds);
	c.xch = o; o += kXF * 16 + kXF * 2;       // partial sums exchanged inside a cluster
	c.sel = o; o += c.klp / 2 + 1;            // uint32 column ids
	c.total = o;
	return c;
}

// columns of the first CTA of a cluster of CS
__host__ __device__ inline int xLocalColumns(int k, int CS) { return CS == 1 ? k : roundUpX((k + 1) / 2, 4); }

// Both products run on mma.sync.m8n8k4 (DMMA): A = 8 x 4 slice of Zs, B = 4 x 8 slice of the fit-major operand,
// C = 8 rows x 8 fits.  NWS warps share one range of the reduction dimension; warp slot `ws` owns the output tiles
// ws, ws + NWS, ...  (NT of them, compile time: no predicates in the loop) and loads each B fragment once for all of
// its tiles.  Per reduction step: 1 + NT shared-memory loads, NT DMMAs, two pointer increments.
template <int NT>
__device__ __forceinline__ void scoreTiles(const double* __restrict__ zp, const double* __restrict__ sp, int nsteps, int ldz4,
                                           int tileStride, double* __restrict__ dst, int ldt) {
	double c0[NT], c1[NT];
#pragma unroll
	for (int j = 0; j < NT; ++j) { c0[j] = 0.0; c1[j] = 0.0; }
#pragma unroll 2
	for (int s = 0; s < nsteps; ++s) {
		const double b = *sp;
#pragma unroll
		for (int j = 0; j < NT; ++j) dmma884x(c0[j], c1[j], zp[j * tileStride], b);
		zp += ldz4;
		sp += 4;
	}
#pragma unroll
	for (int j = 0; j < NT; ++j) {
		dst[j * tileStride] = c0[j];
		dst[j * tileStride + ldt] = c1[j];
	}
}

// dst_h[fit][row] = sum over the columns of half h of Zs[row][column] S[fit][column]; KS halves, outputs dst0 / dst1
template <int KS>
__device__ __forceinline__ void gemmScores(const double* __restrict__ Zs, const double* __restrict__ Sx, double* __restrict__ dst0,
                                           double* __restrict__ dst1, int klp, int nr, int ldz, int lds, int ldt, int warp, int lane) {
	constexpr int NWS = 8 / KS;
	const int g = lane >> 2, t = lane & 3;
	const int half = warp / NWS, ws = warp % NWS;
	const int mtiles = nr >> 3, ksteps = klp >> 2, kper = (ksteps + KS - 1) / KS;
	const int kb = half * kper, nsteps = max(0, min(kper, ksteps - kb));
	const int ntiles = (mtiles - ws + NWS - 1) / NWS; // tiles ws, ws + NWS, ...
	const double* __restrict__ zp = Zs + (4 * kb + t) * ldz + 8 * ws + g;
	const double* __restrict__ sp = Sx + g * lds + 4 * kb + t;
	double* __restrict__ dst = (half ? dst1 : dst0) + (2 * t) * ldt + 8 * ws + g;
	switch (ntiles) {
	case 1: scoreTiles<1>(zp, sp, nsteps, 4 * ldz, 8 * NWS, dst, ldt); break;
	case 2: scoreTiles<2>(zp, sp, nsteps, 4 * ldz, 8 * NWS, dst, ldt); break;
	case 3: scoreTiles<3>(zp, sp, nsteps, 4 * ldz, 8 * NWS, dst, ldt); break;
	case 4: scoreTiles<4>(zp, sp, nsteps, 4 * ldz, 8 * NWS, dst, ldt); break;
	case 5: scoreTiles<5>(zp, sp, nsteps, 4 * ldz, 8 * NWS, dst, ldt); break;
	case 6: scoreTiles<6>(zp, sp, nsteps, 4 * ldz, 8 * NWS, dst, ldt); break;
	default: break;
	}
}

template <int NT>
__device__ __forceinline__ void loadingTiles(const double* __restrict__ zp, const double* __restrict__ tp, int nsteps, int tileStrideZ,
                                             int tileStrideD, double* __restrict__ dst, int lds) {
	double c0[NT], c1[NT];
#pragma unroll
	for (int j = 0; j < NT; ++j) { c0[j] = 0.0; c1[j] = 0.0; }
#pragma unroll 2
	for (int s = 0; s < nsteps; ++s) {
		const double b = *tp;
#pragma unroll
		for (int j = 0; j < NT; ++j) dmma884x(c0[j], c1[j], zp[j * tileStrideZ], b);
		zp += 4;
		tp += 4;
	}
#pragma unroll
	for (int j = 0; j < NT; ++j) {
		dst[j * tileStrideD] = c0[j];
		dst[j * tileStrideD + lds] = c1[j];
	}
}

// dst_h[fit][column] = sum over the rows of half h of Zs[row][column] Tc[fit][row]
template <int KS>
__device__ __forceinline__ void gemmLoadings(const double* __restrict__ Zs, const double* __restrict__ Tc, double* __restrict__ dst0,
                                             double* __restrict__ dst1, int klp, int nr, int ldz, int lds, int ldt, int warp, int lane) {
	constexpr int NWS = 8 / KS;
	const int g = lane >> 2, t = lane & 3;
	const int half = warp / NWS, ws = warp % NWS;
	const int mtiles = klp >> 3, ksteps = nr >> 2, kper = (ksteps + KS - 1) / KS;
	const int kb = half * kper, nsteps = max(0, min(kper, ksteps - kb));
	const int ntiles = (mtiles - ws + NWS - 1) / NWS;
	const double* __restrict__ zp = Zs + (8 * ws + g) * ldz + 4 * kb + t;
	const double* __restrict__ tp = Tc + g * ldt + 4 * kb + t;
	double* __restrict__ dst = (half ? dst1 : dst0) + (2 * t) * lds + 8 * ws + g;
	switch (ntiles) {
	case 1: loadingTiles<1>(zp, tp, nsteps, 8 * NWS * ldz, 8 * NWS, dst, lds); break;
	case 2: loadingTiles<2>(zp, tp, nsteps, 8 * NWS * ldz, 8 * NWS, dst, lds); break;
	case 3: loadingTiles<3>(zp, tp, nsteps, 8 * NWS * ldz, 8 * NWS, dst, lds); break;
	case 4: loadingTiles<4>(zp, tp, nsteps, 8 * NWS * ldz, 8 * NWS, dst, lds); break;
	default: break;
	}
}

template <int CS>
__device__ __forceinline__ void ctaOrClusterSync() {
	if (CS == 1) __syncthreads();
	else cg::this_cluster().sync();
}

template <int CS>
__global__ void __launch_bounds__(kXT, 1) simplsScoreKernel(const XParams prm) {
	extern __shared__ __align__(16) double smem[];
	constexpr int KS = (CS == 1) ? 2 : 1; // a single CTA splits the reduction dimension of both products in two

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int work = blockIdx.x / CS;
	const int rank = (CS == 1) ? 0 : static_cast<int>(cg::this_cluster().block_rank());
	const int ci = work % prm.bucketCount;
	const int gi = work / prm.bucketCount;
	const int chrom = prm.bucket[ci];
	const uint32_t selBegin = prm.offsets[chrom];
	const int k = static_cast<int>(prm.offsets[chrom + 1] - selBegin);
	const int A = min(prm.maxNComp, k);
	const XGroupDesc grp = prm.groups[gi];
	const XRepDesc rep = prm.reps[grp.rep];
	const int n = prm.n, nr = prm.nr;

	const int kh = xLocalColumns(k, CS);
	const int cbeg = rank * kh;
	const int kL = max(0, min(k - cbeg, kh)); // this CTA's columns
	const XCarve cv = xCarve(kh, nr);         // same layout in both CTAs of a cluster
	const int klp = cv.klp, ldz = cv.ldz, lds = cv.lds, ldt = cv.ldt;
	double* __restrict__ Zs = smem + cv.Z;
	double* __restrict__ Sx = smem + cv.S;
	double* __restrict__ Wv = smem + cv.Wv;
	double* __restrict__ V = smem + cv.V;
	double* __restrict__ TA = smem + cv.TA;
	double* __restrict__ TB = smem + cv.TB;
	double* __restrict__ xch = smem + cv.xch;
	double* __restrict__ xch2 = xch + kXF * 16;
	uint32_t* __restrict__ sel = reinterpret_cast<uint32_t*>(smem + cv.sel);
	double* __restrict__ tcBuf = (CS == 1) ? TA : TB; // masked, centred, normalised scores (B operand of P3)
	// partner's buffers (distributed shared memory)
	const double* TAr = TA;
	const double* xchr = xch;
	const double* xch2r = xch2;
	if (CS > 1) {
		cg::cluster_group cluster = cg::this_cluster();
		TAr = cluster.map_shared_rank(TA, rank ^ 1);
		xchr = cluster.map_shared_rank(xch, rank ^ 1);
		xch2r = cluster.map_shared_rank(xch2, rank ^ 1);
	}
	const double* TA0 = (rank == 0) ? TA : TAr;   // rank 0's partial first: both CTAs add in the same order
	const double* TA1 = (rank == 0) ? TAr : TA;
	const double* xc0 = (rank == 0) ? xch : xchr;
	const double* xc1 = (rank == 0) ? xchr : xch;
	const double* xd0 = (rank == 0) ? xch2 : xch2r;
	const double* xd1 = (rank == 0) ? xch2r : xch2;

	const double* __restrict__ zsrc = prm.zpool + rep.zOff;
	const double nanv = __longlong_as_double(0x7ff8000000000000LL);

	// ---- gather: selected columns of every row, atom order (PLS::viewSelectColumns) ---------------------------
	for (int j = tid; j < kL; j += kXT) sel[j] = prm.idx[selBegin + cbeg + j];
	__syncthreads();
	{
		const int pieces = nr >> 1; // 16-byte pieces per column
		for (int e = tid; e < kL * pieces; e += kXT) {
			const int j = e / pieces, piece = e - j * pieces;
			cpAsync16x(Zs + j * ldz + 2 * piece, zsrc + static_cast<size_t>(sel[j]) * nr + 2 * piece);
		}
		for (int e = tid; e < (klp - kL) * nr; e += kXT) {
			const int j = kL + e / nr, r = e % nr;
			Zs[j * ldz + r] = 0.0;
		}
	}

	// ---- this warp's fit: rows of the lane, training mask, tasks ----------------------------------------------
	const int f = warp;
	const bool fitActive = f < grp.nfits;
	XFitDesc xf;
	xf.exclMask = 0u; xf.taskBegin = 0; xf.ntasks = 0; xf.pad = 0;
	if (fitActive) xf = prm.xfits[grp.fitBegin + f];
	double yv[kXRU], pred[kXRU];
	int atomOf[kXRU];
	uint32_t trainBits = 0u, liveBits = 0u;
	{
		const int* __restrict__ ar = prm.atomRows + rep.atomBegin;
		const double* __restrict__ ycol = zsrc + static_cast<size_t>(prm.p) * nr;
#pragma unroll
		for (int u = 0; u < kXRU; ++u) {
			const int r = lane + 32 * u;
			yv[u] = 0.0;
			pred[u] = 0.0;
			atomOf[u] = -1;
			if (r < n) {
				yv[u] = __ldg(ycol + r);
				int a = 0;
				while (a + 1 < rep.natoms && r >= ar[a + 1]) ++a;
				atomOf[u] = a;
				liveBits |= 1u << u;
				if (!((xf.exclMask >> a) & 1u)) trainBits |= 1u << u;
			}
		}
	}
	double ntr = 0.0, ym = 0.0;
	{
		double cnt = 0.0, sy = 0.0;
#pragma unroll
		for (int u = 0; u < kXRU; ++u) {
			if ((trainBits >> u) & 1u) { cnt += 1.0; sy += yv[u]; }
		}
		ntr = warpSumX(cnt);
		ym = warpSumX(sy) / ntr;
	}
	const double rntr = 1.0 / ntr;
	// S_0 = X_c' y_c (:126) as a P3 product: masked, centred response in place of the scores
#pragma unroll
	for (int u = 0; u < kXRU; ++u) {
		const int r = lane + 32 * u;
		if (r < nr) tcBuf[f * ldt + r] = (fitActive && ((trainBits >> u) & 1u)) ? yv[u] - ym : 0.0;
	}
	cpAsyncWaitAllX();
	ctaOrClusterSync<CS>(); // also: the partner CTA runs before any distributed-shared-memory access
	gemmLoadings<KS>(Zs, tcBuf, Wv, TB, klp, nr, ldz, lds, ldt, warp, lane);
	__syncthreads();
#pragma unroll
	for (int u = 0; u < kXCU; ++u) {
		const int i = lane + 32 * u;
		if (i < klp) Sx[f * lds + i] = Wv[f * lds + i] + (KS == 2 ? TB[f * lds + i] : 0.0);
	}
	__syncthreads();

#pragma unroll 1
	for (int comp = 0; comp < A; ++comp) {
		// ---- P1: scores of every row --------------------------------------------------------------------------
		gemmScores<KS>(Zs, Sx, TA, TB, klp, nr, ldz, lds, ldt, warp, lane);
		ctaOrClusterSync<CS>();

		// ---- P2: centre, normalise, predict the held-out rows -------------------------------------------------
		bool under = false;
		if (fitActive) {
			double tc[kXRU];
			double sum = 0.0;
#pragma unroll
			for (int u = 0; u < kXRU; ++u) {
				const int r = lane + 32 * u;
				tc[u] = 0.0;
				if (r < nr) tc[u] = TA0[f * ldt + r] + ((KS == 2) ? TB[f * ldt + r] : TA1[f * ldt + r]);
				if ((trainBits >> u) & 1u) sum += tc[u];
			}
			const double mu = warpSumX(sum) * rntr;
			double a2 = 0.0, b2 = 0.0;
#pragma unroll
			for (int u = 0; u < kXRU; ++u) {
				tc[u] -= mu;
				if ((trainBits >> u) & 1u) {
					a2 = fma(tc[u], tc[u], a2);
					b2 = fma(yv[u], tc[u], b2);
				}
			}
			a2 = warpSumX(a2);
			b2 = warpSumX(b2);
			under = !(a2 >= 1e-50); // ||t|| < 1e-25 (NORM_TOL, :16,:141-143); also NaN
			const double rtn = rsqrt(a2);
			const double q = b2 * rtn;
			const double fac = under ? nanv : q * rtn; // coefficient increment is S q / tn (:152-156)
#pragma unroll
			for (int u = 0; u < kXRU; ++u) {
				if (((liveBits & ~trainBits) >> u) & 1u) pred[u] = fma(tc[u], fac, pred[u]);
			}
			// residual sums of the prediction tasks, two tasks per reduction
			for (int tk = 0; tk < xf.ntasks; tk += 2) {
				const XTaskDesc td0 = prm.xtasks[xf.taskBegin + tk];
				const bool two = tk + 1 < xf.ntasks;
				XTaskDesc td1 = td0;
				if (two) td1 = prm.xtasks[xf.taskBegin + tk + 1];
				const int atom1 = two ? td1.atom : -2;
				double se0 = 0.0, see0 = 0.0, se1 = 0.0, see1 = 0.0;
#pragma unroll
				for (int u = 0; u < kXRU; ++u) {
					const double e = yv[u] - (ym + pred[u]);
					if (atomOf[u] == td0.atom) { se0 += e; see0 = fma(e, e, see0); }
					if (atomOf[u] == atom1) { se1 += e; see1 = fma(e, e, see1); }
				}
#pragma unroll
				for (int m = 16; m >= 1; m >>= 1) {
					se0 += shflXorX(se0, m);
					see0 += shflXorX(see0, m);
					se1 += shflXorX(se1, m);
					see1 += shflXorX(see1, m);
				}
				if (lane < 2 && rank == 0 && (lane == 0 || two)) {
					const XTaskDesc td = lane ? td1 : td0;
					const double se = lane ? se1 : se0, see = lane ? see1 : see0;
					if (td.kind == 0) {
						// mean squared error of prediction for this fold (src/PLSEvaluator.cpp:266-268)
						const int rows = prm.atomRows[rep.atomBegin + td.atom + 1] - prm.atomRows[rep.atomBegin + td.atom];
						prm.mse[(static_cast<size_t>(chrom) * prm.innerDests + td.dest) * prm.maxNComp + comp] = see / static_cast<double>(rows);
					} else {
						double* o = prm.ostat + ((static_cast<size_t>(chrom) * prm.outerDests + td.dest) * prm.maxNComp + comp) * 2;
						o[0] = se;
						o[1] = see;
					}
				}
			}
#pragma unroll
			for (int u = 0; u < kXRU; ++u) {
				const int r = lane + 32 * u;
				if (r < nr) tcBuf[f * ldt + r] = ((trainBits >> u) & 1u) ? (under ? nanv : tc[u] * rtn) : 0.0;
			}
		}
		__syncthreads();

		// ---- P3: loadings v = X't --------------------------------------------------------------------------------
		gemmLoadings<KS>(Zs, tcBuf, Wv, TB, klp, nr, ldz, lds, ldt, warp, lane);
		__syncthreads();

		// ---- P4: orthogonalise against the stored loadings, deflate S ------------------------------------------
		double v[kXCU], s[kXCU];
#pragma unroll
		for (int u = 0; u < kXCU; ++u) {
			const int i = lane + 32 * u;
			v[u] = 0.0;
			s[u] = 0.0;
			if (fitActive && i < klp) {
				v[u] = Wv[f * lds + i] + (KS == 2 ? TB[f * lds + i] : 0.0);
				s[u] = Sx[f * lds + i];
			}
		}
		double cj[kXA - 1];
		{
			double c[16];
#pragma unroll
			for (int j = 0; j < 16; ++j) c[j] = 0.0;
#pragma unroll
			for (int j = 0; j < kXA - 1; ++j) {
				if (j < comp) {
#pragma unroll
					for (int u = 0; u < kXCU; ++u) {
						const int i = lane + 32 * u;
						if (i < klp) c[j] = fma(V[(j * kXF + f) * klp + i], v[u], c[j]);
					}
				}
			}
			double tot = warpSum16x(c, lane);
			if (CS > 1) {
				if ((lane & 1) == 0) xch[f * 16 + (lane >> 1)] = tot;
				cg::this_cluster().sync();
				tot = xc0[f * 16 + (lane >> 1)] + xc1[f * 16 + (lane >> 1)];
			}
#pragma unroll
			for (int j = 0; j < kXA - 1; ++j) cj[j] = __shfl_sync(0xffffffffu, tot, 2 * j);
		}
		double vv = 0.0, vs = 0.0;
#pragma unroll
		for (int u = 0; u < kXCU; ++u) {
			const int i = lane + 32 * u;
			if (i < klp) {
#pragma unroll
				for (int j = 0; j < kXA - 1; ++j) {
					if (j < comp) v[u] = fma(-V[(j * kXF + f) * klp + i], cj[j], v[u]);
				}
			}
			vv = fma(v[u], v[u], vv);
			vs = fma(v[u], s[u], vs);
		}
		vv = warpSumX(vv);
		vs = warpSumX(vs);
		if (CS > 1) {
			if (lane == 0) {
				xch2[f * 2] = vv;
				xch2[f * 2 + 1] = vs;
			}
			cg::this_cluster().sync();
			vv = xd0[f * 2] + xd1[f * 2];
			vs = xd0[f * 2 + 1] + xd1[f * 2 + 1];
		}
		if (fitActive) {
			const double rvn = rsqrt(vv);
			const double dd = (vs * rvn) * rvn;
#pragma unroll
			for (int u = 0; u < kXCU; ++u) {
				const int i = lane + 32 * u;
				if (i < klp) {
					double sn = fma(-v[u], dd, s[u]);
					if (under) sn = nanv; // the reference throws: poison this and every later component
					Sx[f * lds + i] = sn;
					if (comp < kXA - 1) V[(comp * kXF + f) * klp + i] = v[u] * rvn;
				}
			}
		}
		__syncthreads();
	}
	if (CS > 1) cg::this_cluster().sync(); // no CTA exits while its partner may still read its shared memory
}

template <int CS>
cudaError_t launchScoreOne(const XParams& prm, int kMax, cudaStream_t stream, long long* ctas) {
	const XCarve cv = xCarve(xLocalColumns(kMax, CS), prm.nr);
	const size_t bytes = static_cast<size_t>(cv.total) * sizeof(double);
	const long long grid = static_cast<long long>(prm.ngroups) * prm.bucketCount * CS;
	if (grid <= 0) return cudaSuccess;
	if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
	cudaError_t err = cudaFuncSetAttribute(simplsScoreKernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
	if (err != cudaSuccess) return err;
	cudaLaunchConfig_t cfg;
	cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
	cfg.blockDim = dim3(kXT, 1, 1);
	cfg.dynamicSmemBytes = bytes;
	cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = CS;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	err = cudaLaunchKernelEx(&cfg, simplsScoreKernel<CS>, prm);
	if (err != cudaSuccess) return err;
	if (ctas) *ctas += grid;
	return cudaGetLastError();
}

} // namespace

// largest subset size a cluster of `clusterSize` CTAs takes (selected columns + 9 loadings in shared memory)
int scoreKernelMaxK(int n, int clusterSize, int smemLimit) {
	if (n > kScoreMaxRows || clusterSize < 1 || clusterSize > 2) return 0;
	const int nr = roundUpX(n, 8);
	int k = 0;
	while (k < 32 * kXCU * clusterSize) {
		const XCarve cv = xCarve(xLocalColumns(k + 1, clusterSize), nr);
		if (cv.klp > 32 * kXCU || cv.total * 8 > smemLimit) break;
		++k;
	}
	return k;
}

cudaError_t launchScoreKernel(const XParams& prm, int kMax, int A, int smemLimit, cudaStream_t stream, long long* ctas) {
	if (A > kXA || prm.n > kScoreMaxRows) return cudaErrorInvalidValue;
	if (kMax <= scoreKernelMaxK(prm.n, 1, smemLimit)) return launchScoreOne<1>(prm, kMax, stream, ctas);
	if (kMax <= scoreKernelMaxK(prm.n, 2, smemLimit)) return launchScoreOne<2>(prm, kMax, stream, ctas);
	return cudaErrorInvalidValue;
}

} // namespace gsb
