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

constexpr int kXA = kScoreComponents;
constexpr int kXRU = kScoreMaxSlots;      // 32-row slots: lane = row within the slot

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

// Sum NV (a power of two <= 16) values per lane over the warp with NV - 1 + log2(32 / NV) shuffles: the lanes halve
// the value set at every butterfly step.  Afterwards lane L holds the total of value L / (32 / NV).
template <int NV>
__device__ __forceinline__ double warpSumN(double (&v)[NV], int lane) {
	int m = 16;
#pragma unroll
	for (int h = NV / 2; h >= 1; h >>= 1, m >>= 1) {
		const bool up = (lane & m) != 0;
#pragma unroll
		for (int i = 0; i < h; ++i) {
			const double send = up ? v[i] : v[i + h];
			const double keep = up ? v[i + h] : v[i];
			v[i] = keep + shflXorX(send, m);
		}
	}
	double r = v[0];
#pragma unroll
	for (; m >= 1; m >>= 1) r += shflXorX(r, m);
	return r;
}

// shared-memory carve-up (in doubles).  Leading dimensions are = 4 or 12 (mod 16): the 8x4 / 4x8 tensor-core
// fragments (lane -> 4 * (lane / 4) + lane % 4 or 4 * (lane % 4) + lane / 4 rows/columns) are then bank-conflict free.
struct XCarve {
	int klp, ldz, lds, ldt;
	int Z, S, Wv, V, TA, TB, xch, sel, total;
};

__host__ __device__ inline XCarve xCarve(int kLocal, int nr, int NF, bool vInSmem) {
	XCarve c;
	c.klp = roundUpX(kLocal > 0 ? kLocal : 1, 8);
	c.ldz = nr + 4;
	c.ldt = nr + 4;
	c.lds = c.klp + 4;
	int o = 0;
	c.Z = o; o += c.klp * c.ldz;              // [column][row]
	c.S = o; o += NF * c.lds;                 // [fit][column]
	c.Wv = o; o += NF * c.lds;
	c.V = o; o += vInSmem ? (kXA - 1) * NF * c.klp : 0; // [loading][fit][column]; else in a global scratch (L2)
	c.TA = o; o += NF * c.ldt;                // [fit][row]
	c.TB = o; o += NF * (c.ldt > c.lds ? c.ldt : c.lds);
	c.xch = o; o += NF * 16 + NF * 2;         // partial sums exchanged inside a cluster
	c.sel = o; o += c.klp / 2 + 1;            // uint32 column ids
	c.total = o;
	return c;
}

// columns of every CTA of a cluster of CS but the last (which takes the rest)
__host__ __device__ inline int xLocalColumns(int k, int CS) { return CS == 1 ? k : roundUpX((k + CS - 1) / CS, 4); }

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
	case 5: loadingTiles<5>(zp, tp, nsteps, 8 * NWS * ldz, 8 * NWS, dst, lds); break;
	case 6: loadingTiles<6>(zp, tp, nsteps, 8 * NWS * ldz, 8 * NWS, dst, lds); break;
	default: break;
	}
}

// barrier of one 8-fit tile (256 threads) of a decoupled 16-fit CTA; id 1 or 2
__device__ __forceinline__ void tileSync(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }

template <int CS>
__device__ __forceinline__ void ctaOrClusterSync() {
	if (CS == 1) __syncthreads();
	else cg::this_cluster().sync();
}

// the same buffer in the shared memory of CTA `r` of the cluster
template <int CS>
__device__ __forceinline__ const double* rankView(const double* mine, int r, int rank) {
	if (CS == 1 || r == rank) return mine;
	return cg::this_cluster().map_shared_rank(const_cast<double*>(mine), r);
}

// projections of v on the first `comp` stored loadings (NV >= comp values per lane), summed over the warp and, in a
// cluster, over the CTAs; cj[j] = V_j'v on every lane
template <int NF, int CS, int CU, int NV>
__device__ __forceinline__ void projectOnLoadings(const double* __restrict__ V, const double (&v)[CU], int f, int klp, int lane, int comp,
                                                  double* __restrict__ xch, int rank,
                                                  double (&cj)[kXA - 1]) {
	double c[NV];
#pragma unroll
	for (int j = 0; j < NV; ++j) {
		c[j] = 0.0;
		if (j < comp) {
#pragma unroll
			for (int u = 0; u < CU; ++u) {
				const int i = lane + 32 * u;
				if (i < klp) c[j] = fma(V[(j * NF + f) * klp + i], v[u], c[j]);
			}
		}
	}
	double tot = warpSumN<NV>(c, lane);
	constexpr int LP = 32 / NV; // lanes per value
	if (CS > 1) {
		if ((lane & (LP - 1)) == 0) xch[f * 16 + lane / LP] = tot;
		cg::this_cluster().sync();
		tot = 0.0;
#pragma unroll
		for (int rk = 0; rk < CS; ++rk) tot += rankView<CS>(xch, rk, rank)[f * 16 + lane / LP];
	}
#pragma unroll
	for (int j = 0; j < kXA - 1; ++j) cj[j] = (j < NV) ? __shfl_sync(0xffffffffu, tot, (j < NV ? j : 0) * LP) : 0.0;
}

// NF fits (= warps) per CTA, CS CTAs per cluster (columns split over the CTAs); VG: the stored loadings live in a
// per-CTA global scratch (L2-resident while the CTA runs) instead of shared memory, which almost doubles the columns
// one CTA can take
template <int NF, int CS, bool VG>
__global__ void __launch_bounds__(32 * NF, 1) simplsScoreKernel(const XParams prm) {
	extern __shared__ __align__(16) double smem[];
	constexpr int KS = (CS == 1) ? 2 : 1; // a single CTA splits the reduction dimension of both products in two
	constexpr int NTHR = 32 * NF;
	constexpr int CU = VG ? ((NF == 16) ? 4 : 5) : ((NF == 16) ? 2 : 3); // local columns per lane

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int w8 = warp & 7, half16 = warp >> 3; // fits come in tensor-core column tiles of 8; 8 warps work on each
	const int work = blockIdx.x / CS;
	const int rank = (CS == 1) ? 0 : static_cast<int>(cg::this_cluster().block_rank());
	const int ci = work % prm.bucketCount;
	const int gi = work / prm.bucketCount;
	const int chrom = prm.bucket[ci];
	const uint32_t selBegin = prm.offsets[chrom];
	const int k = static_cast<int>(prm.offsets[chrom + 1] - selBegin);
	const int A = min(prm.maxNComp, k);
	const XGroupDesc grp = prm.groups[gi];
	const XRepDesc* __restrict__ rep = prm.reps + grp.rep;
	const int nr = prm.nr;
	const int nslots = rep->nslots;

	const int kh = xLocalColumns(k, CS);
	const int cbeg = rank * kh;
	const int kL = max(0, min(k - cbeg, kh)); // this CTA's columns
	const XCarve cv = xCarve(kh, nr, NF, !VG); // same layout in every CTA of a cluster
	const int klp = cv.klp, ldz = cv.ldz, lds = cv.lds, ldt = cv.ldt;
	double* __restrict__ Zs = smem + cv.Z;
	double* __restrict__ Sx = smem + cv.S;
	double* __restrict__ Wv = smem + cv.Wv;
	// vGlobal: one scratch region per CTA, or -- when at most one CTA of this launch fits an SM (vslots > 0) -- per SM
	unsigned vslot = blockIdx.x;
	if (VG && prm.vslots > 0) asm volatile("mov.u32 %0, %%smid;" : "=r"(vslot));
	double* __restrict__ V = VG ? prm.vscratch + static_cast<size_t>(vslot) * prm.vstride : smem + cv.V;
	double* __restrict__ TA = smem + cv.TA;
	double* __restrict__ TB = smem + cv.TB;
	double* __restrict__ xch = smem + cv.xch;
	double* __restrict__ xch2 = xch + NF * 16;
	uint32_t* __restrict__ sel = reinterpret_cast<uint32_t*>(smem + cv.sel);
	// operands of this warp's column tile of 8 fits
	const double* __restrict__ SxH = Sx + half16 * 8 * lds;
	double* __restrict__ WvH = Wv + half16 * 8 * lds;
	double* __restrict__ TAH = TA + half16 * 8 * ldt;
	double* __restrict__ TBsH = TB + half16 * 8 * (ldt > lds ? ldt : lds); // TB as the second partial of P1 (stride ldt) ...
	double* __restrict__ TBwH = TBsH;                                      // ... and of P3 (stride lds); one region per tile
	const double* __restrict__ tcH = (CS == 1) ? TAH : TBsH;
	const int fl = warp & 7; // this warp's fit inside its tile
	double* __restrict__ tcW = (CS == 1) ? TAH : TBsH; // masked, centred, normalised scores of the tile (B operand of P3)

	const bool tracing = prm.trace != nullptr && blockIdx.x == (gridDim.x / (2 * CS)) * CS && (tid & 255) == 0;
	int stamp = 0;
	const int stampBase = 32 * (tid >> 8); // the two column tiles of 8 fits stamp separately
#define GSB_XSTAMP() do { if (tracing && stamp < 32) prm.trace[stampBase + stamp++] = clock64(); } while (0)
	GSB_XSTAMP();

	const double* __restrict__ zsrc = prm.zpool + rep->zOff;
	const double nanv = __longlong_as_double(0x7ff8000000000000LL);

	// ---- gather: selected columns of every row, atom order (PLS::viewSelectColumns) ---------------------------
	for (int j = tid; j < kL; j += NTHR) sel[j] = prm.idx[selBegin + cbeg + j];
	__syncthreads();
	{
		const int pieces = nr >> 1; // 16-byte pieces per column
		for (int j = warp; j < klp; j += NF) {
			double* __restrict__ dcol = Zs + j * ldz;
			if (j < kL) {
				const double* __restrict__ scol = zsrc + static_cast<size_t>(sel[j]) * nr;
				for (int piece = lane; piece < pieces; piece += 32) cpAsync16x(dcol + 2 * piece, scol + 2 * piece);
			} else {
				for (int r = lane; r < nr; r += 32) dcol[r] = 0.0;
			}
		}
	}

	// ---- this warp's fit: slot masks (warp-uniform), centred response of the lane's rows -----------------------
	const int f = warp;
	const bool fitActive = f < grp.nfits;
	XFitDesc xf;
	xf.exclMask = 0u; xf.taskBegin = 0; xf.ntasks = 0; xf.pad = 0;
	if (fitActive) xf = prm.xfits[grp.fitBegin + f];
	// the fit's first two prediction tasks and the atom of every slot live in registers (no global loads in the loop)
	XTaskDesc tdA, tdB;
	tdA.atom = -2; tdA.kind = -1; tdA.dest = 0; tdA.rows = 1;
	tdB = tdA;
	if (xf.ntasks > 0) tdA = prm.xtasks[xf.taskBegin];
	if (xf.ntasks > 1) tdB = prm.xtasks[xf.taskBegin + 1];
	const double rrowsA = 1.0 / static_cast<double>(tdA.rows), rrowsB = 1.0 / static_cast<double>(tdB.rows);
	uint32_t slotAtoms = 0u; // 5 bits per slot
#pragma unroll
	for (int u = 0; u < kXRU; ++u) if (u < nslots) slotAtoms |= static_cast<uint32_t>(rep->slotAtom[u] & 31) << (5 * u);
	uint32_t trainSlots = 0u, liveBits = 0u; // bit u: slot u is training data of this fit / the lane's row of slot u exists
	double yc[kXRU], pred[kXRU];
	double rntr, ym, syc;
	{
		const double* __restrict__ ycol = zsrc + static_cast<size_t>(prm.p) * nr;
		double cnt = 0.0, sy = 0.0;
#pragma unroll
		for (int u = 0; u < kXRU; ++u) {
			yc[u] = 0.0;
			pred[u] = 0.0;
			if (u < nslots) {
				const bool live = lane < rep->slotRows[u];
				if (live) liveBits |= 1u << u;
				yc[u] = live ? __ldg(ycol + 32 * u + lane) : 0.0;
				if (!((xf.exclMask >> ((slotAtoms >> (5 * u)) & 31u)) & 1u)) {
					trainSlots |= 1u << u;
					cnt += live ? 1.0 : 0.0;
					sy += yc[u];
				}
			}
		}
#pragma unroll
		for (int m = 16; m >= 1; m >>= 1) {
			cnt += shflXorX(cnt, m);
			sy += shflXorX(sy, m);
		}
		ym = sy / cnt;
		rntr = 1.0 / cnt;
		double s2 = 0.0;
#pragma unroll
		for (int u = 0; u < kXRU; ++u) {
			yc[u] = ((liveBits >> u) & 1u) ? yc[u] - ym : 0.0;
			if ((trainSlots >> u) & 1u) s2 += yc[u];
		}
#pragma unroll
		for (int m = 16; m >= 1; m >>= 1) s2 += shflXorX(s2, m);
		syc = s2; // rounding residue of the centring, ~ 1e-16
	}
	// S_0 = X_c' y_c (:126) as a P3 product: masked, centred response in place of the scores
#pragma unroll
	for (int u = 0; u < kXRU; ++u) {
		if (u < nslots) tcW[fl * ldt + 32 * u + lane] = (fitActive && ((trainSlots >> u) & 1u)) ? yc[u] : 0.0;
	}
	cpAsyncWaitAllX();
	ctaOrClusterSync<CS>(); // also: the partner CTAs run before any distributed-shared-memory access
	GSB_XSTAMP(); // 1: gather done
	// a 16-fit single CTA may run its two 8-fit tiles independently from here on (they share only the read-only Zs):
	// one tile's warp-local FP64 phases then overlap the other's tensor-core products
	const bool decoupled = (NF == 16 && CS == 1) && prm.desync > 0;
#define GSB_PHASE_SYNC() do { if (decoupled) tileSync(1 + half16); else __syncthreads(); } while (0)
	if (decoupled && prm.desync >= 2 && half16 == 1) asm volatile("bar.sync 3, 512;" ::: "memory"); // start half a component late
	gemmLoadings<KS>(Zs, tcH, WvH, TBwH, klp, nr, ldz, lds, ldt, w8, lane);
	GSB_PHASE_SYNC();
#pragma unroll
	for (int u = 0; u < CU; ++u) {
		const int i = lane + 32 * u;
		if (i < klp) Sx[f * lds + i] = Wv[f * lds + i] + (KS == 2 ? TBwH[fl * lds + i] : 0.0);
	}
	GSB_PHASE_SYNC();
	GSB_XSTAMP(); // 2: S_0 done

#pragma unroll 1
	for (int comp = 0; comp < A; ++comp) {
		// ---- P1: scores of every row --------------------------------------------------------------------------
		gemmScores<KS>(Zs, SxH, TAH, TBsH, klp, nr, ldz, lds, ldt, w8, lane);
		if (CS == 1) { GSB_PHASE_SYNC(); } else ctaOrClusterSync<CS>();
		GSB_XSTAMP();
		if (decoupled && prm.desync == 2 && comp == 0 && half16 == 0) asm volatile("bar.arrive 3, 512;" ::: "memory");

		// ---- P2: centre, normalise, predict the held-out rows -------------------------------------------------
		bool under = false;
		if (fitActive) {
			double t[kXRU];
			double s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
			for (int u = 0; u < kXRU; ++u) {
				t[u] = 0.0;
				if (u < nslots) {
					const int at = f * ldt + 32 * u + lane;
					if (CS == 1) {
						t[u] = TA[at] + TBsH[fl * ldt + 32 * u + lane];
					} else {
						// rank 0's partial first: every CTA of the cluster adds in the same order
						double part[CS];
#pragma unroll
						for (int rk = 0; rk < CS; ++rk) part[rk] = rankView<CS>(TA, rk, rank)[at];
#pragma unroll
						for (int rk = 0; rk < CS; ++rk) t[u] += part[rk];
					}
					if ((trainSlots >> u) & 1u) { // padding rows are zero rows of Z: t = 0, yc = 0
						s1 += t[u];
						s2 = fma(t[u], t[u], s2);
						s3 = fma(yc[u], t[u], s3);
					}
				}
			}
#pragma unroll
			for (int m = 16; m >= 1; m >>= 1) {
				s1 += shflXorX(s1, m);
				s2 += shflXorX(s2, m);
				s3 += shflXorX(s3, m);
			}
			// centring of the scores = column-centring of X (:28-49): |t - mu|^2 = sum t^2 - n mu^2, y_c'(t - mu)
			const double mu = s1 * rntr;
			const double a2 = fma(-s1, mu, s2);
			under = !(a2 >= 1e-50); // ||t|| < 1e-25 (NORM_TOL, :16,:141-143); also NaN
			const double b2 = fma(-mu, syc, s3);
			const double rtn = rsqrt(a2);
			const double q = b2 * rtn;
			const double fac = under ? nanv : q * rtn; // coefficient increment is S q / tn (:152-156)
			const double scale = under ? nanv : rtn;
#pragma unroll
			for (int u = 0; u < kXRU; ++u) {
				if (u < nslots) {
					const double tc = t[u] - mu;
					if ((trainSlots >> u) & 1u) {
						tcW[fl * ldt + 32 * u + lane] = tc * scale;
					} else {
						tcW[fl * ldt + 32 * u + lane] = 0.0;
						pred[u] = fma(tc, fac, pred[u]);
					}
				}
			}
			// residual sums of the prediction tasks, two tasks per reduction
			for (int tk = 0; tk < xf.ntasks; tk += 2) {
				XTaskDesc td0 = tdA, td1 = tdB;
				if (tk > 0) { // more than two tasks on one fit: not in a nested CV, but legal
					td0 = prm.xtasks[xf.taskBegin + tk];
					td1.atom = -2;
					if (tk + 1 < xf.ntasks) td1 = prm.xtasks[xf.taskBegin + tk + 1];
				}
				const bool two = tk + 1 < xf.ntasks;
				const int atom1 = two ? td1.atom : -2;
				double st[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
				for (int u = 0; u < kXRU; ++u) {
					if (u < nslots) {
						const int au = static_cast<int>((slotAtoms >> (5 * u)) & 31u);
						if (au == td0.atom || au == atom1) {
							const double e = ((liveBits >> u) & 1u) ? yc[u] - pred[u] : 0.0;
							if (au == td0.atom) { st[0] += e; st[1] = fma(e, e, st[1]); }
							if (au == atom1) { st[2] += e; st[3] = fma(e, e, st[3]); }
						}
					}
				}
				const double tot = warpSumN<4>(st, lane); // lane L: value L / 8
				if ((lane & 7) == 0 && rank == 0) {
					const int vi = lane >> 3;
					const XTaskDesc td = (vi >= 2) ? td1 : td0;
					if (vi < 2 || two) {
						if (td.kind == 0) {
							// mean squared error of prediction for this fold (src/PLSEvaluator.cpp:266-268)
							if (vi & 1) prm.mse[(static_cast<size_t>(chrom) * prm.innerDests + td.dest) * prm.maxNComp + comp] = tot * ((tk > 0) ? 1.0 / static_cast<double>(td.rows) : ((vi >= 2) ? rrowsB : rrowsA));
						} else {
							prm.ostat[((static_cast<size_t>(chrom) * prm.outerDests + td.dest) * prm.maxNComp + comp) * 2 + (vi & 1)] = tot;
						}
					}
				}
			}
		}

		GSB_PHASE_SYNC();
		GSB_XSTAMP();
		if (decoupled && prm.desync == 3 && comp == 0 && half16 == 0) asm volatile("bar.arrive 3, 512;" ::: "memory");

		// ---- P3: loadings v = X't --------------------------------------------------------------------------------
		gemmLoadings<KS>(Zs, tcH, WvH, TBwH, klp, nr, ldz, lds, ldt, w8, lane);
		GSB_PHASE_SYNC();
		GSB_XSTAMP();

		// ---- P4: orthogonalise against the stored loadings, deflate S ------------------------------------------
		double v[CU], s[CU];
#pragma unroll
		for (int u = 0; u < CU; ++u) {
			const int i = lane + 32 * u;
			v[u] = 0.0;
			s[u] = 0.0;
			if (fitActive && i < klp) {
				v[u] = Wv[f * lds + i] + (KS == 2 ? TBwH[fl * lds + i] : 0.0);
				s[u] = Sx[f * lds + i];
			}
		}
		if (comp > 0) {
			double cj[kXA - 1];
			if (comp <= 2) projectOnLoadings<NF, CS, CU, 2>(V, v, f, klp, lane, comp, xch, rank, cj);
			else if (comp <= 4) projectOnLoadings<NF, CS, CU, 4>(V, v, f, klp, lane, comp, xch, rank, cj);
			else if (comp <= 8) projectOnLoadings<NF, CS, CU, 8>(V, v, f, klp, lane, comp, xch, rank, cj);
			else projectOnLoadings<NF, CS, CU, 16>(V, v, f, klp, lane, comp, xch, rank, cj);
#pragma unroll
			for (int u = 0; u < CU; ++u) {
				const int i = lane + 32 * u;
				if (i < klp) {
					// two partial sums: half the dependent chain
					double d0 = 0.0, d1 = 0.0;
#pragma unroll
					for (int j = 0; j < kXA - 1; j += 2) {
						if (j < comp) d0 = fma(V[(j * NF + f) * klp + i], cj[j], d0);
						if (j + 1 < kXA - 1 && j + 1 < comp) d1 = fma(V[((j + 1) * NF + f) * klp + i], cj[j + 1], d1);
					}
					v[u] -= d0 + d1;
				}
			}
		}
		double vv = 0.0, vs = 0.0;
#pragma unroll
		for (int u = 0; u < CU; ++u) {
			vv = fma(v[u], v[u], vv);
			vs = fma(v[u], s[u], vs);
		}
#pragma unroll
		for (int m = 16; m >= 1; m >>= 1) {
			vv += shflXorX(vv, m);
			vs += shflXorX(vs, m);
		}
		if (CS > 1) {
			if (lane == 0) {
				xch2[f * 2] = vv;
				xch2[f * 2 + 1] = vs;
			}
			cg::this_cluster().sync();
			vv = 0.0;
			vs = 0.0;
#pragma unroll
			for (int rk = 0; rk < CS; ++rk) {
				const double* x2 = rankView<CS>(xch2, rk, rank);
				vv += x2[f * 2];
				vs += x2[f * 2 + 1];
			}
		}
		if (fitActive) {
			const double dd = vs / vv;
			const double rvn = rsqrt(vv);
#pragma unroll
			for (int u = 0; u < CU; ++u) {
				const int i = lane + 32 * u;
				if (i < klp) {
					double sn = fma(-v[u], dd, s[u]);
					if (under) sn = nanv; // the reference throws: poison this and every later component
					Sx[f * lds + i] = sn;
					if (comp < kXA - 1) V[(comp * NF + f) * klp + i] = v[u] * rvn;
				}
			}
		}
		GSB_PHASE_SYNC();
		GSB_XSTAMP();
	}
	if (decoupled && prm.desync >= 2 && half16 == 0 && A == 0) asm volatile("bar.arrive 3, 512;" ::: "memory");
#undef GSB_XSTAMP
#undef GSB_PHASE_SYNC
	if (CS > 1) cg::this_cluster().sync(); // no CTA exits while a partner may still read its shared memory
}

template <int NF, int CS, bool VG>
cudaError_t launchScoreOne(const XParams& prm, int kMax, cudaStream_t stream, long long* ctas) {
	const XCarve cv = xCarve(xLocalColumns(kMax, CS), prm.nr, NF, !VG);
	if (VG && (prm.vscratch == nullptr || prm.vstride < static_cast<long long>(kXA - 1) * NF * cv.klp)) return cudaErrorInvalidValue;
	const size_t bytes = static_cast<size_t>(cv.total) * sizeof(double);
	const long long grid = static_cast<long long>(prm.ngroups) * prm.bucketCount * CS;
	if (grid <= 0) return cudaSuccess;
	if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
	cudaError_t err = cudaFuncSetAttribute(simplsScoreKernel<NF, CS, VG>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
	if (err != cudaSuccess) return err;
	cudaLaunchConfig_t cfg;
	cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
	cfg.blockDim = dim3(32 * NF, 1, 1);
	cfg.dynamicSmemBytes = bytes;
	cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = CS;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	err = cudaLaunchKernelEx(&cfg, simplsScoreKernel<NF, CS, VG>, prm);
	if (err != cudaSuccess) return err;
	if (ctas) *ctas += grid;
	return cudaGetLastError();
}

} // namespace

// largest subset size a cluster of `clusterSize` CTAs with nf fits each takes (0: configuration not available);
// vGlobal: stored loadings in the global scratch
int scoreKernelMaxK(int nr, int nf, int clusterSize, int vGlobal, int smemLimit) {
	if (nr > 32 * kScoreMaxSlots || clusterSize < 1 || clusterSize > 3 || (nf != 8 && nf != 16)) return 0;
	const int cu = vGlobal ? ((nf == 16) ? 4 : 5) : ((nf == 16) ? 2 : 3);
	int k = 0;
	while (k < 32 * cu * clusterSize) {
		const XCarve cv = xCarve(xLocalColumns(k + 1, clusterSize), nr, nf, !vGlobal);
		if (cv.klp > 32 * cu || cv.total * 8 > smemLimit) break;
		++k;
	}
	return k;
}

template <int NF, int CS>
int scoreOccupancy(int kMax, int nr) {
	const XCarve cv = xCarve(xLocalColumns(kMax, CS), nr, NF, false);
	int nb = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, simplsScoreKernel<NF, CS, true>, 32 * NF, static_cast<size_t>(cv.total) * sizeof(double)) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return nb;
}

// Scratch regions a vGlobal launch of `ctas` CTAs needs: one per CTA, or kScoreSmSlots (indexed by %smid) when the
// launch is large and at most one of its CTAs fits an SM at a time (returned negative: index by %smid).
int scoreKernelScratchSlots(int kMax, int nf, int clusterSize, int nr, long long ctas) {
	if (ctas <= kScoreSmSlots) return static_cast<int>(ctas);
	int nb = 0;
	if (nf == 16) nb = clusterSize == 1 ? scoreOccupancy<16, 1>(kMax, nr) : clusterSize == 2 ? scoreOccupancy<16, 2>(kMax, nr) : scoreOccupancy<16, 3>(kMax, nr);
	else nb = clusterSize == 1 ? scoreOccupancy<8, 1>(kMax, nr) : clusterSize == 2 ? scoreOccupancy<8, 2>(kMax, nr) : scoreOccupancy<8, 3>(kMax, nr);
	if (nb == 1) return -kScoreSmSlots; // negative: indexed by %smid
	return ctas > 2147483647LL ? 0 : static_cast<int>(ctas);
}

// doubles of global scratch one CTA needs for its stored loadings (vGlobal configurations)
long long scoreKernelScratchDoubles(int kMax, int nf, int clusterSize) {
	return static_cast<long long>(kXA - 1) * nf * roundUpX(xLocalColumns(kMax, clusterSize), 8);
}

cudaError_t launchScoreKernel(const XParams& prm, int kMax, int A, int clusterSize, int vGlobal, int smemLimit, cudaStream_t stream,
                              long long* ctas) {
	if (A > kXA || kMax > scoreKernelMaxK(prm.nr, prm.nf, clusterSize, vGlobal, smemLimit)) return cudaErrorInvalidValue;
	if (prm.nf == 16) {
		if (vGlobal) {
			if (clusterSize == 1) return launchScoreOne<16, 1, true>(prm, kMax, stream, ctas);
			if (clusterSize == 2) return launchScoreOne<16, 2, true>(prm, kMax, stream, ctas);
			return launchScoreOne<16, 3, true>(prm, kMax, stream, ctas);
		}
		if (clusterSize == 1) return launchScoreOne<16, 1, false>(prm, kMax, stream, ctas);
		if (clusterSize == 2) return launchScoreOne<16, 2, false>(prm, kMax, stream, ctas);
		return launchScoreOne<16, 3, false>(prm, kMax, stream, ctas);
	}
	if (vGlobal) {
		if (clusterSize == 1) return launchScoreOne<8, 1, true>(prm, kMax, stream, ctas);
		if (clusterSize == 2) return launchScoreOne<8, 2, true>(prm, kMax, stream, ctas);
		return launchScoreOne<8, 3, true>(prm, kMax, stream, ctas);
	}
	if (clusterSize == 1) return launchScoreOne<8, 1, false>(prm, kMax, stream, ctas);
	if (clusterSize == 2) return launchScoreOne<8, 2, false>(prm, kMax, stream, ctas);
	return launchScoreOne<8, 3, false>(prm, kMax, stream, ctas);
}

} // namespace gsb
