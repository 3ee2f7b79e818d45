// simpls_kspace.cu -- K1k: SIMPLS in the space of the rows ("kernel space") on the FP64 tensor cores.
//
// What it replaces per CTA (same as simpls_fit.cu / simpls_scores.cu): PLS::viewSelectColumns / viewSelectRows
// (src/PLS.cpp:13-23), PLSSimpls::fit (src/PLSSimpls.cpp:106-166), PLS::predict (src/PLS.cpp:34-47) and the MSEP /
// residual sums of PLSEvaluator::estSEP (src/PLSEvaluator.cpp:263-268, 323-325).
//
// Every vector the reference's recurrences touch lies in the range of X' (S_0 = X_c'y_c, the loadings v = X_c't), so
// the whole fit can be carried by vectors over the ROWS.  With Z = the selected columns of the globally centred data
// (all rows), K = Z Z' (n x n, ONE matrix per chromosome, shared by every fit of every replication), a fit given by
// its training-row mask m, and n-vectors g, h with S = Z'g, v = Z'h:
//
//   scores of every row      t = Z S            = K g        kept by recurrence: t <- t - (K h) (v'S)/(v'v)   (:135,:160)
//   centring                 mean over the training rows of t  (= column-centring of X applied to t, :28-49)
//   tc = m.(t - mu)/|.|      normalised training scores (:138-145);   q = y_c'tc (:148)
//   predictions              yhat += (t - mu) q / |t|   on the held-out rows (cumulative coefficients, :152-156)
//   loading                  v = Z'tc,  K tc = w  is THE tensor-core product of the component  (152 x 152) . (152 x 8)
//   projections              V_j'v = (K h_j)'tc = U_j'tc   with U_j = K h_j / |v_j| stored per fit (:57-63)
//   orthogonalised           K h = w - sum_j (U_j'tc) U_j = u,  v'v = tc'w - sum_j (U_j'tc)^2,  v'S = tc't   (:70-88)
//
// One product per component instead of the two of score space (Z S and Z't), its size independent of the subset size,
// no stored loadings, no size classes, no clusters: the subset size only enters the one-off K = Z Z' (a DMMA SYRK over
// column chunks streamed through shared memory with cp.async).  One warp per fit: lane = row (mod 32), the fit's
// vectors (t, y_c, yhat, tc, the nine U_j) live in registers; masks are per-lane bits, so ANY training sets and
// held-out blocks work (no atom structure needed).  Two CTA barriers per component (around the product).
// A fit starts from one 128-byte record (KFitDesc: masks, the host-computed centring constants, its first two tasks).
// NF = 8 fits per pass is the default; NF = 16 (opt-in) keeps the U_j in tensor memory (tcgen05.st / tcgen05.ld as
// per-thread scratch) because 512 threads leave only 128 registers each -- measured no faster, see DESIGN.md section 4.
//
// Numerics: mathematically the reference's SIMPLS; K squares the conditioning like the Gram-space kernel does.
// tests/test_kspace_model.py holds the numpy model of exactly these recurrences (<= 6e-12 against the reference on
// the parity data set, the same band as two CPU builds of the reference's algorithm that differ in summation order).
#include "kernels.cuh"

namespace gsb {

namespace {

constexpr int kKA = kKspaceComponents;
constexpr int kKNS = kKspaceSlots;   // 32-row slots per lane
// NF = fits per pass = warps per CTA: 8 (one tensor-core column tile; the stored vectors of a fit in registers) or 16
// (two column tiles sharing every fragment of K; 128 registers per thread, the stored vectors in tensor memory)
constexpr int kKProductWarps = 8;    // warps that issue the products (enough to saturate the FP64 tensor pipe)
constexpr int kKUStride = 12;        // TMEM columns per stored vector: 5 doubles as .x8 + .x4

__host__ __device__ inline int roundUpK(int v, int m) { return (v + m - 1) / m * m; }

__device__ __forceinline__ void dmma884k(double& d0, double& d1, double a, double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cpAsync16k(double* dst, const double* src) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src));
}
__device__ __forceinline__ void cpAsyncWaitAllK() {
	asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ double shflXorK(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// Tensor memory as per-thread scratch (32x32b: thread i of the warp <-> lane 32 (warp % 4) + i, consecutive columns).
// Measured (profiles/micro/tmem_probe.cu): 415-640 B/clk/SM for loads, 3-5 x shared memory; ~65 cycles ld + wait.
__device__ __forceinline__ void tmemLd8(uint32_t (&r)[12], uint32_t addr) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
	             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
	             : "r"(addr));
}
__device__ __forceinline__ void tmemLd4(uint32_t (&r)[12], uint32_t addr) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]) : "r"(addr));
}
__device__ __forceinline__ void tmemSt12(uint32_t addr, const uint32_t (&r)[12]) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
	             "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
	asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(addr + 8), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]));
}
__device__ __forceinline__ void tmemWaitLd() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmemWaitSt() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// Sum NV (a power of two <= 16) values per lane over the warp with NV - 1 + log2(32 / NV) shuffles: the lanes halve
// the value set at every butterfly step.  Afterwards lane L holds the total of value L / (32 / NV).
template <int NV>
__device__ __forceinline__ double warpSumNK(double (&v)[NV], int lane) {
	int m = 16;
#pragma unroll
	for (int h = NV / 2; h >= 1; h >>= 1, m >>= 1) {
		const bool up = (lane & m) != 0;
#pragma unroll
		for (int i = 0; i < h; ++i) {
			const double send = up ? v[i] : v[i + h];
			const double keep = up ? v[i + h] : v[i];
			v[i] = keep + shflXorK(send, m);
		}
	}
	double r = v[0];
#pragma unroll
	for (; m >= 1; m >>= 1) r += shflXorK(r, m);
	return r;
}

struct KCarve {
	int ldk;
	int K, TA, TB, ys, total; // in doubles; the column ids of a chunk alias TA, the staged chunk aliases K
};

__host__ __device__ inline KCarve kCarve(int n8, int nf) {
	KCarve c;
	c.ldk = n8 + 4; // = 4 or 12 (mod 16): the 8x4 / 4x8 fragment loads are bank-conflict free
	int o = 0;
	c.K = o; o += n8 * c.ldk;
	c.TA = o; o += nf * c.ldk; // [fit][row]: B operand of the product
	c.TB = o; o += nf * c.ldk; // [fit][row]: its result
	c.ys = o; o += n8;
	c.total = o;
	return c;
}

// dst[fit][row] = sum_r K[row][r] src[fit][r]: 8 warps share the row tiles (tiles ws, ws + 8, ...), NT per warp; CT
// column tiles of 8 fits each reuse every fragment of K
template <int NT, int CT>
__device__ __forceinline__ void productTiles(const double* __restrict__ kp, const double* __restrict__ bp, int nsteps, int tileStride,
                                             double* __restrict__ dst, int ldk) {
	double c0[CT][NT], c1[CT][NT];
#pragma unroll
	for (int c = 0; c < CT; ++c) {
#pragma unroll
		for (int j = 0; j < NT; ++j) { c0[c][j] = 0.0; c1[c][j] = 0.0; }
	}
#pragma unroll 2
	for (int s = 0; s < nsteps; ++s) {
		double b[CT];
#pragma unroll
		for (int c = 0; c < CT; ++c) b[c] = bp[c * 8 * ldk];
#pragma unroll
		for (int j = 0; j < NT; ++j) {
			const double a = kp[j * tileStride];
#pragma unroll
			for (int c = 0; c < CT; ++c) dmma884k(c0[c][j], c1[c][j], a, b[c]);
		}
		kp += 4;
		bp += 4;
	}
#pragma unroll
	for (int c = 0; c < CT; ++c) {
#pragma unroll
		for (int j = 0; j < NT; ++j) {
			dst[c * 8 * ldk + j * 64] = c0[c][j];
			dst[c * 8 * ldk + j * 64 + ldk] = c1[c][j];
		}
	}
}

// executed by the first 8 warps of the CTA; ct = column tiles in use (1 or 2)
__device__ __forceinline__ void kernelProduct(const double* __restrict__ Km, const double* __restrict__ src, double* __restrict__ dst,
                                              int n8, int ldk, int warp, int lane, int ct) {
	if (warp >= kKProductWarps) return;
	const int g = lane >> 2, t = lane & 3;
	const int nt = n8 >> 3, nsteps = n8 >> 2;
	const int ntiles = (nt - warp + kKProductWarps - 1) / kKProductWarps;
	const double* __restrict__ kp = Km + (8 * warp + g) * ldk + t;
	const double* __restrict__ bp = src + g * ldk + t;
	double* __restrict__ dp = dst + (2 * t) * ldk + 8 * warp + g;
	if (ct == 1) {
		switch (ntiles) {
		case 1: productTiles<1, 1>(kp, bp, nsteps, 64 * ldk, dp, ldk); break;
		case 2: productTiles<2, 1>(kp, bp, nsteps, 64 * ldk, dp, ldk); break;
		case 3: productTiles<3, 1>(kp, bp, nsteps, 64 * ldk, dp, ldk); break;
		default: break;
		}
	} else {
		switch (ntiles) {
		case 1: productTiles<1, 2>(kp, bp, nsteps, 64 * ldk, dp, ldk); break;
		case 2: productTiles<2, 2>(kp, bp, nsteps, 64 * ldk, dp, ldk); break;
		case 3: productTiles<3, 2>(kp, bp, nsteps, 64 * ldk, dp, ldk); break;
		default: break;
		}
	}
}

// what a component step needs besides the fit's own vectors
struct KCtx {
	const KParams* prm;
	const double* Km;
	double* TA;
	double* TB;
	double* taW;       // this warp's row of TA
	const double* tbW; // this warp's row of TB
	int n8, ldk, warp, lane, A, chrom;
	const KFitDesc* fd;
	bool fitActive;
	bool tracing;
};

// the vectors of one fit, one warp per fit, lane = row (mod 32): all in registers -- but for the nine stored
// vectors U of the 16-fit configuration, which live in tensor memory (12 columns per vector at tmemU)
template <int NF>
struct KFitState {
	double yc[kKNS], tt[kKNS], pred[kKNS];
	double U[NF == 8 ? kKA - 1 : 1][kKNS];
	uint32_t tmemU;
	uint32_t trainBits, rowBits;  // bit u: the lane's row of slot u is a training row / exists in shared memory
	uint32_t taskBitsA, taskBitsB; // bit u: the lane's row of slot u belongs to the fit's first / second prediction task
	double* outA;                 // where the task's statistic of component 0 goes (kind 0: mse, kind 1: sum e, sum e^2)
	double* outB;
	int kindA, kindB;             // -1: no such task
	double rrA, rrB;              // 1 / rows of the task
	double rntr, syc;
	bool under;
	bool ill; // the orthogonalised loading lost (nearly) all of its norm: v'v by subtraction is not trustworthy
};

// statistic of one prediction task for component COMP from its residual sums (src/PLSEvaluator.cpp:266-268, 323-325)
__device__ __forceinline__ void storeTaskSum(int kind, double* out, double rr, int which, double tot, int comp) {
	if (kind == 0) {
		if (which) out[comp] = tot * rr; // mean squared error of prediction for this fold
	} else if (kind == 1) {
		out[2 * comp + which] = tot;     // sum e, sum e^2 of the pooled outer residuals
	}
}

// First half of component `comp` (0-based), the same code for every component: centre and normalise the scores,
// predict the held-out rows, hand tc to the product, residual sums of the fit's first two tasks (left in st[] for the
// second half's reduction when one follows, else reduced and stored here).
template <int NF>
__device__ __forceinline__ void kspaceScores(KFitState<NF>& fs, const KCtx& cx, int comp, bool more, double (&tc)[kKNS], double (&st)[4]) {
	const KParams& prm = *cx.prm;
	const int lane = cx.lane;
	const double nanv = __longlong_as_double(0x7ff8000000000000LL);
	// ---- centre, normalise, predict the held-out rows ----------------------------------------------------------
	{
		double sv[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
		for (int u = 0; u < kKNS; ++u) {
			if ((fs.trainBits >> u) & 1u) {
				sv[0] += fs.tt[u];
				sv[1] = fma(fs.tt[u], fs.tt[u], sv[1]);
				sv[2] = fma(fs.yc[u], fs.tt[u], sv[2]);
			}
		}
		const double tot = warpSumNK<4>(sv, lane); // lane L: value L / 8
		const double s1 = __shfl_sync(0xffffffffu, tot, 0), s2 = __shfl_sync(0xffffffffu, tot, 8), s3 = __shfl_sync(0xffffffffu, tot, 16);
		// centring of the scores = column-centring of X (:28-49): |t - mu|^2 = sum t^2 - n mu^2, y_c'(t - mu)
		const double mu = s1 * fs.rntr;
		const double a2 = fma(-s1, mu, s2);
		fs.under = fs.under || !(a2 >= 1e-50); // ||t|| < 1e-25 (NORM_TOL, :16,:141-143); also NaN
		const double b2 = fma(-mu, fs.syc, s3);
		const double rtn = rsqrt(a2);
		const double q = b2 * rtn;
		const double fac = fs.under ? nanv : q * rtn; // coefficient increment is S q / tn (:152-156)
		const double scale = fs.under ? nanv : rtn;
#pragma unroll
		for (int u = 0; u < kKNS; ++u) {
			const double d = fs.tt[u] - mu;
			tc[u] = ((fs.trainBits >> u) & 1u) ? d * scale : 0.0;
			fs.pred[u] = fma(d, fac, fs.pred[u]); // only the held-out rows are read
		}
	}
	if (more) {
#pragma unroll
		for (int u = 0; u < kKNS; ++u) {
			if ((fs.rowBits >> u) & 1u) cx.taW[32 * u + lane] = tc[u];
		}
	}
	// residual sums of the fit's (first two) prediction tasks: reduced together with the dot products below
#pragma unroll
	for (int i = 0; i < 4; ++i) st[i] = 0.0;
#pragma unroll
	for (int u = 0; u < kKNS; ++u) {
		const double e = fs.yc[u] - fs.pred[u];
		const double eA = ((fs.taskBitsA >> u) & 1u) ? e : 0.0;
		const double eB = ((fs.taskBitsB >> u) & 1u) ? e : 0.0;
		st[0] += eA;
		st[1] = fma(eA, eA, st[1]);
		st[2] += eB;
		st[3] = fma(eB, eB, st[3]);
	}
	// more than two tasks on one fit: not in a nested CV, but legal
	const KFitDesc& fd = *cx.fd;
	for (int tk = 2; cx.fitActive && tk < fd.ntasks; ++tk) {
		const KTaskDesc td = prm.ktasks[fd.taskBegin + tk];
		double sx[2] = {0.0, 0.0};
#pragma unroll
		for (int u = 0; u < kKNS; ++u) {
			const double e = ((td.rows[u] >> lane) & 1u) ? fs.yc[u] - fs.pred[u] : 0.0;
			sx[0] += e;
			sx[1] = fma(e, e, sx[1]);
		}
		const double tot = warpSumNK<2>(sx, lane); // lane L: value L / 16
		if ((lane & 15) == 0) {
			double* out = td.kind == 0 ? prm.mse + (static_cast<size_t>(cx.chrom) * prm.innerDests + td.dest) * prm.maxNComp
			                           : prm.ostat + (static_cast<size_t>(cx.chrom) * prm.outerDests + td.dest) * prm.maxNComp * 2;
			storeTaskSum(td.kind, out, 1.0 / static_cast<double>(td.nrows), lane >> 4, tot, comp);
		}
	}
	if (!more) {
		const double tot = warpSumNK<4>(st, lane); // lane L: value L / 8
		if ((lane & 7) == 0) {
			const int vi = lane >> 3;
			storeTaskSum(vi < 2 ? fs.kindA : fs.kindB, vi < 2 ? fs.outA : fs.outB, vi < 2 ? fs.rrA : fs.rrB, vi & 1, tot, comp);
		}
	}
}

// Second half of component `comp` in [JLO, JHI], after w = K tc: orthogonalise against the `comp` stored vectors,
// deflate the scores, store the new vector (:57-63, :70-88, :160).  The stored vectors live in registers, so their
// index must be a compile-time constant: an instance covers the components JLO..JHI with the terms j >= comp
// predicated off.  The first half of a component is deliberately NOT part of these instances: unrolling whole
// components made 230 KB of code, an instruction-cache hit rate of 63 % and "no instruction" the second largest stall.
template <int NF, int JLO, int JHI>
__device__ __forceinline__ void kspaceDeflate(KFitState<NF>& fs, const KCtx& cx, int comp, const double (&tc)[kKNS], const double (&st)[4]) {
	const int lane = cx.lane;
	const double nanv = __longlong_as_double(0x7ff8000000000000LL);
	double w[kKNS];
#pragma unroll
	for (int u = 0; u < kKNS; ++u) w[u] = ((fs.rowBits >> u) & 1u) ? cx.tbW[32 * u + lane] : 0.0;
	// one reduction for: V_j'v (JHI slots), v'S, |X'tc|^2, and the four residual sums
	constexpr int NV = (JHI + 6 <= 8) ? 8 : 16;
	double sv[NV];
#pragma unroll
	for (int j = 0; j < NV; ++j) sv[j] = 0.0;
	if (NF == 8) {
#pragma unroll
		for (int u = 0; u < kKNS; ++u) {
#pragma unroll
			for (int j = 0; j < JHI; ++j) {
				if (j < JLO || j < comp) sv[j] = fma(fs.U[j][u], tc[u], sv[j]); // V_j'v = U_j'tc
			}
		}
	} else {
#pragma unroll
		for (int j0 = 0; j0 < JHI; j0 += 2) { // two stored vectors per round trip to tensor memory
			uint32_t r[2][12];
#pragma unroll
			for (int jj = 0; jj < 2; ++jj) {
				if (j0 + jj < JHI) {
					tmemLd8(r[jj], fs.tmemU + kKUStride * (j0 + jj));
					tmemLd4(r[jj], fs.tmemU + kKUStride * (j0 + jj) + 8);
				}
			}
			tmemWaitLd();
#pragma unroll
			for (int jj = 0; jj < 2; ++jj) {
				if (j0 + jj < JHI && (j0 + jj < JLO || j0 + jj < comp)) {
#pragma unroll
					for (int u = 0; u < kKNS; ++u) sv[j0 + jj] = fma(__hiloint2double(r[jj][2 * u + 1], r[jj][2 * u]), tc[u], sv[j0 + jj]);
				}
			}
		}
	}
#pragma unroll
	for (int u = 0; u < kKNS; ++u) {
		sv[JHI] = fma(tc[u], fs.tt[u], sv[JHI]);     // v'S = tc't
		sv[JHI + 1] = fma(tc[u], w[u], sv[JHI + 1]); // |X'tc|^2 = tc'K tc
	}
#pragma unroll
	for (int i = 0; i < 4; ++i) sv[JHI + 2 + i] = st[i];
	const double tot = warpSumNK<NV>(sv, lane);
	constexpr int LP = 32 / NV; // lanes per value
	{
		const int vi = lane / LP - (JHI + 2);
		if ((lane % LP) == 0 && vi >= 0 && vi < 4) {
			storeTaskSum(vi < 2 ? fs.kindA : fs.kindB, vi < 2 ? fs.outA : fs.outB, vi < 2 ? fs.rrA : fs.rrB, vi & 1, tot, comp);
		}
	}
	double cj[kKA];
	const double vvRaw = __shfl_sync(0xffffffffu, tot, (JHI + 1) * LP); // |X'tc|^2 before the projections are taken off
	double vv = vvRaw, vv2 = 0.0;
	const double vs = __shfl_sync(0xffffffffu, tot, JHI * LP);
#pragma unroll
	for (int j = 0; j < JHI; ++j) {
		cj[j] = __shfl_sync(0xffffffffu, tot, j * LP); // exactly 0 for j >= comp
		// |v - sum_j (V_j'v) V_j|^2 for orthonormal V_j (two chains)
		if (j & 1) vv2 = fma(cj[j], cj[j], vv2);
		else vv = fma(-cj[j], cj[j], vv);
	}
	vv -= vv2;
	// v'v comes from a subtraction on K = Z Z' (squared conditioning); the reference takes the norm of the explicit
	// vector (src/PLSSimpls.cpp:57-63).  When next to nothing is left the chromosome goes to the Gram-space kernel.
	fs.ill = fs.ill || !(vv > 1e-8 * vvRaw);
	const double rvn = rsqrt(vv);
	const double dd = vs * rvn * rvn; // v'S / v'v
	double dsum[kKNS]; // sum_j (V_j'v) U_j
#pragma unroll
	for (int u = 0; u < kKNS; ++u) dsum[u] = 0.0;
	if (NF == 8) {
#pragma unroll
		for (int u = 0; u < kKNS; ++u) {
			// two partial sums: half the dependent chain
			double d0 = 0.0, d1 = 0.0;
#pragma unroll
			for (int j = 0; j < JHI; j += 2) {
				if (j < JLO || j < comp) d0 = fma(fs.U[j][u], cj[j], d0);
				if (j + 1 < JHI && (j + 1 < JLO || j + 1 < comp)) d1 = fma(fs.U[j + 1][u], cj[j + 1], d1);
			}
			dsum[u] = d0 + d1;
		}
	} else {
#pragma unroll
		for (int j0 = 0; j0 < JHI; j0 += 2) {
			uint32_t r[2][12];
#pragma unroll
			for (int jj = 0; jj < 2; ++jj) {
				if (j0 + jj < JHI) {
					tmemLd8(r[jj], fs.tmemU + kKUStride * (j0 + jj));
					tmemLd4(r[jj], fs.tmemU + kKUStride * (j0 + jj) + 8);
				}
			}
			tmemWaitLd();
#pragma unroll
			for (int jj = 0; jj < 2; ++jj) {
				if (j0 + jj < JHI && (j0 + jj < JLO || j0 + jj < comp)) {
#pragma unroll
					for (int u = 0; u < kKNS; ++u) dsum[u] = fma(__hiloint2double(r[jj][2 * u + 1], r[jj][2 * u]), cj[j0 + jj], dsum[u]);
				}
			}
		}
	}
	uint32_t packed[12];
#pragma unroll
	for (int u = 0; u < kKNS; ++u) {
		const double uu = w[u] - dsum[u];   // K h for the orthogonalised loading v = Z'h
		double tn = fma(-uu, dd, fs.tt[u]); // Z (S - v v'S / v'v)
		if (fs.under) tn = nanv;            // the reference throws: poison this and every later component
		fs.tt[u] = tn;
		const double un = uu * rvn;
		if (NF == 8) {
#pragma unroll
			for (int j = JLO; j <= JHI; ++j) {
				if (j == comp) fs.U[NF == 8 ? j : 0][u] = un;
			}
		} else {
			packed[2 * u] = static_cast<uint32_t>(__double2loint(un));
			packed[2 * u + 1] = static_cast<uint32_t>(__double2hiint(un));
		}
	}
	if (NF != 8) {
		packed[10] = 0u;
		packed[11] = 0u;
		tmemSt12(fs.tmemU + kKUStride * comp, packed);
		tmemWaitSt();
	}
}

// The one-SE rule inside the kernel: called by every thread of the CTA once its fits that feed the MSEP table are done
// (their global writes are visible CTA-wide after the barrier).  Slots [first, total) of the part's fit order are
// outer-only fits: each gets the optNComp of its outer fold(s) (oneSeRule, the function K2 uses: the same decision to
// the bit) and they are sorted, longest first, into sched[first ...) -- staged behind the schedule at sched + stride.
__device__ __forceinline__ void kspaceRule(const KParams& prm, const int* fitOrder, const KPartDesc& pd, uint32_t* sched, int first, int total, int chrom, int A,
                                        int tid, int nthreads) {
	__syncthreads();
	uint32_t* raw = sched + prm.schedStride;
	const int rem = total - first;
	for (int j = tid; j < rem; j += nthreads) {
		const int fj = fitOrder[pd.begin + first + j];
		const int tb = prm.kfits[fj].taskBegin, ntk = prm.kfits[fj].ntasks;
		int nc = 1;
		for (int tk = 0; tk < ntk; ++tk) {
			const int grp = prm.ktasks[tb + tk].dest; // outer task: dest = (replication, outer fold)
			const double* m = prm.mse + (static_cast<size_t>(chrom) * prm.innerDests + static_cast<size_t>(grp) * prm.inner) * prm.maxNComp;
			const OneSe rule = oneSeRule(m, prm.inner, A, prm.maxNComp, prm.sdfact, -1.0);
			nc = max(nc, rule.bad ? A : rule.opt);
		}
		raw[j] = (static_cast<uint32_t>(nc) << 24) | static_cast<uint32_t>(fj);
	}
	__syncthreads();
	for (int j = tid; j < rem; j += nthreads) { // rank sort, longest first; ties keep their order
		const uint32_t key = raw[j];
		int rank = 0;
		for (int i = 0; i < rem; ++i) {
			const uint32_t ki = raw[i] >> 24;
			rank += (ki > (key >> 24) || (ki == (key >> 24) && i < j)) ? 1 : 0;
		}
		sched[first + rank] = key;
	}
	__syncthreads();
}

// TRACE: the instance that stamps clock64() at its phase boundaries (launched only when prm.trace is set: the stamps
// and the loads of their destination cost ~3 % in the component loop even when predicated off)
template <int NF, bool TRACE>
__global__ void __launch_bounds__(32 * NF, 1) simplsKspaceKernel(const KParams prm) {
	constexpr int kKNF = NF, kKThreads = 32 * NF;
	constexpr int kKBlocks = (55 + NF - 1) / NF; // 2x2-tile blocks of the lower triangle of K per warp (20 tiles: 55 blocks)
	extern __shared__ __align__(16) double smem[];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int g = lane >> 2, t4 = lane & 3;
	const int n = prm.n, n8 = prm.n8, ldz = prm.ldz;
	const KCarve cv = kCarve(n8, NF);
	const int ldk = cv.ldk;
	double* __restrict__ Km = smem + cv.K;
	double* __restrict__ TA = smem + cv.TA;
	double* __restrict__ TB = smem + cv.TB;
	double* __restrict__ ys = smem + cv.ys;
	uint32_t* __restrict__ sel = reinterpret_cast<uint32_t*>(TA);

	// Largest subsets first (the bucket list is ascending): their K costs the most.  The first fullCount chromosomes get
	// one CTA each (whole waves of one CTA per SM); the rest -- the last, partial wave, or all of a small population --
	// are split over tailSplit CTAs each, which come last in the grid so that they fill the SMs as those run dry.
	int cj, part;
	const KPartDesc* partTab;
	const int* fitOrder;
	if (static_cast<int>(blockIdx.x) < prm.fullCount) {
		cj = static_cast<int>(blockIdx.x);
		part = 0;
		partTab = prm.partsFull;
		fitOrder = prm.fitOrderFull;
	} else {
		const int t = static_cast<int>(blockIdx.x) - prm.fullCount;
		cj = prm.fullCount + t / prm.tailSplit;
		part = t % prm.tailSplit;
		partTab = prm.parts;
		fitOrder = prm.fitOrder;
	}
	const int chrom = prm.bucket[prm.bucketCount - 1 - cj];
	const uint32_t selBegin = prm.offsets[chrom];
	const int k = static_cast<int>(prm.offsets[chrom + 1] - selBegin);
	const int A = min(prm.maxNComp, k);
	const double nanv = __longlong_as_double(0x7ff8000000000000LL);
	// this CTA's fits: fitOrder[pd.begin ...), the first n0 with every component, then n1 that only predict outer folds
	const KPartDesc pd = partTab[part];
	const int total = pd.n0 + pd.n1;
	if (total == 0) return;

	// GSB_GUARD: a canary behind the carve (the launch asks for 64 bytes more); checked at the single exit below
	constexpr unsigned long long kCanary = 0xA5C3A5C3A5C3A5C3ull;
	unsigned long long* canary = reinterpret_cast<unsigned long long*>(smem + cv.total);
	if (prm.guardFail != nullptr && tid < 8) canary[tid] = kCanary;
	const bool tracing = TRACE && prm.trace != nullptr && blockIdx.x == gridDim.x / 2 && tid == 0;
	int stamp = 0;
#define GSB_KSTAMP() do { if (tracing && stamp < 64) prm.trace[stamp++] = clock64(); } while (0)
	GSB_KSTAMP();

	// globally centred response of every row (column p of Z)
	for (int r = tid; r < n8; r += kKThreads) ys[r] = (r < n) ? __ldg(prm.zpool + static_cast<size_t>(prm.p) * ldz + r) : 0.0;

	// the 16-fit configuration keeps the stored vectors of its fits in tensor memory: 4 warps per 32-lane quadrant,
	// 128 columns each (9 vectors x 12 columns used).  Allocated here, released at the single exit below.
	uint32_t tmemBase = 0u;
	if (NF != 8) {
		uint32_t* slot = reinterpret_cast<uint32_t*>(TB); // TB is idle until the first product
		if (warp == 0) {
			asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(slot))));
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
		}
		asm volatile("tcgen05.fence::before_thread_sync;\n");
		__syncthreads();
		asm volatile("tcgen05.fence::after_thread_sync;\n");
		tmemBase = *slot;
	}

	// ---- K = Z Z' over column chunks (PLS::viewSelectColumns + the implicit X X' of the recurrences) -------------
	{
		const int nt = n8 >> 3, nb = (nt + 1) >> 1, nblocks = nb * (nb + 1) / 2;
		double acc[kKBlocks][8];
		int rowA[kKBlocks], rowB[kKBlocks]; // first row of the block's tile pairs; -1: no block
#pragma unroll
		for (int b = 0; b < kKBlocks; ++b) {
#pragma unroll
			for (int i = 0; i < 8; ++i) acc[b][i] = 0.0;
			const int id = warp + kKNF * b;
			rowA[b] = -1;
			rowB[b] = 0;
			if (id < nblocks) {
				int bi = 0;
				while ((bi + 1) * (bi + 2) / 2 <= id) ++bi;
				rowA[b] = 16 * bi;
				rowB[b] = 16 * (id - bi * (bi + 1) / 2);
			}
		}
		const int chunk = n8; // columns per chunk: the staged chunk [column][ldk] fills exactly the area of K
		for (int c0 = 0; c0 < k; c0 += chunk) {
			const int kc = min(chunk, k - c0), kc4 = roundUpK(kc, 4);
			__syncthreads(); // the previous chunk has been consumed
			for (int j = tid; j < kc; j += kKThreads) sel[j] = prm.idx[selBegin + c0 + j];
			__syncthreads();
			const int pieces = ldz >> 1; // 16-byte pieces per column (ldz is a multiple of 4)
			for (int j = warp; j < kc4; j += kKNF) {
				double* __restrict__ dcol = Km + j * ldk;
				if (j < kc) {
					const double* __restrict__ scol = prm.zpool + static_cast<size_t>(sel[j]) * ldz;
					for (int piece = lane; piece < pieces; piece += 32) cpAsync16k(dcol + 2 * piece, scol + 2 * piece);
					for (int r = ldz + lane; r < ldk; r += 32) dcol[r] = 0.0;
				} else {
					for (int r = lane; r < ldk; r += 32) dcol[r] = 0.0;
				}
			}
			cpAsyncWaitAllK();
			__syncthreads();
			const int nsteps = kc4 >> 2;
#pragma unroll
			for (int b = 0; b < kKBlocks; ++b) {
				if (rowA[b] >= 0) {
					// tile rows beyond n8 (odd tile count) read the padding / the next column: results discarded
					const double* __restrict__ pa = Km + t4 * ldk + rowA[b] + g;
					const double* __restrict__ pb = Km + t4 * ldk + rowB[b] + g;
#pragma unroll 2
					for (int s = 0; s < nsteps; ++s) {
						const double a0 = pa[0], a1 = pa[8], b0 = pb[0], b1 = pb[8];
						dmma884k(acc[b][0], acc[b][1], a0, b0);
						dmma884k(acc[b][2], acc[b][3], a0, b1);
						dmma884k(acc[b][4], acc[b][5], a1, b0);
						dmma884k(acc[b][6], acc[b][7], a1, b1);
						pa += 4 * ldk;
						pb += 4 * ldk;
					}
				}
			}
		}
		__syncthreads(); // every warp is done with the staged columns: K takes their place
#pragma unroll
		for (int b = 0; b < kKBlocks; ++b) {
			if (rowA[b] >= 0) {
#pragma unroll
				for (int i = 0; i < 2; ++i) {
#pragma unroll
					for (int j = 0; j < 2; ++j) {
						const int r0 = rowA[b] + 8 * i, c0 = rowB[b] + 8 * j;
						if (r0 < n8 && c0 < n8 && r0 >= c0) {
							const double v0 = acc[b][(i * 2 + j) * 2], v1 = acc[b][(i * 2 + j) * 2 + 1];
							Km[(r0 + g) * ldk + c0 + 2 * t4] = v0;
							Km[(r0 + g) * ldk + c0 + 2 * t4 + 1] = v1;
							if (r0 != c0) {
								Km[(c0 + 2 * t4) * ldk + r0 + g] = v0;
								Km[(c0 + 2 * t4 + 1) * ldk + r0 + g] = v1;
							}
						}
					}
				}
			}
		}
	}
	__syncthreads();
	GSB_KSTAMP(); // 1: K done

	// ---- passes of NF fits: one warp per fit ---------------------------------------------------------------------
	// The reference fits its outer models with exactly optNComp components (src/PLSEvaluator.cpp:316-317).  Once the
	// passes holding the fits that feed the MSEP table are done, the CTA runs the one-SE rule itself (oneSeRule, the
	// function K2 uses: same decision to the bit), sorts the remaining outer-only fits by their optNComp and runs each
	// pass only as far as its longest fit needs.  (An outer-only fit sharing a pass with MSEP fits runs every component.)
	const int passes = (total + kKNF - 1) / kKNF;
	const int ruleAt = (pd.n1 > 0 && pd.n0 > 0) ? (pd.n0 + kKNF - 1) / kKNF : passes;
	// the CTA's schedule: one word per slot, (components to run) << 24 | fit; the slots from ruleAt * NF on are written
	// by kspaceRule below
	uint32_t* sched = prm.sched + static_cast<size_t>(blockIdx.x) * 2 * prm.schedStride;
	for (int j = tid; j < min(total, ruleAt * kKNF); j += kKThreads) {
		sched[j] = (static_cast<uint32_t>(A) << 24) | static_cast<uint32_t>(fitOrder[pd.begin + j]);
	}
	unsigned long long nprod = 0ull;
	double* __restrict__ taW = TA + warp * ldk;
	const double* __restrict__ tbW = TB + warp * ldk;
	__syncthreads();
#pragma unroll 1
	for (int pass = 0; pass < passes; ++pass) {
		if (pass == ruleAt) kspaceRule(prm, fitOrder, pd, sched, ruleAt * kKNF, total, chrom, A, tid, kKThreads);
		const int slot = min(pass * kKNF + warp, total - 1);
		const bool fitActive = pass * kKNF + warp < total;
		const int fi = static_cast<int>(sched[slot] & 0xffffffu);
		const int Apass = min(A, static_cast<int>(sched[pass * kKNF] >> 24));
		// one 128-byte record per fit: masks, centring constants and its first two tasks (no dependent loads)
		const KFitDesc fd = prm.kfits[fi];
		KFitState<NF> fs;
		fs.tmemU = tmemBase + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + static_cast<uint32_t>(128 * (warp >> 2));
		fs.trainBits = 0u; // bit u: the lane's row of slot u is a training row of the fit
		fs.rowBits = 0u;   // bit u: the lane's row of slot u exists in shared memory
		fs.under = false;
		fs.ill = false;
		fs.taskBitsA = 0u; fs.taskBitsB = 0u;
		fs.kindA = fitActive ? fd.first[0].kind : -1;
		fs.kindB = fitActive ? fd.first[1].kind : -1;
		fs.rrA = fd.first[0].rr;
		fs.rrB = fd.first[1].rr;
		fs.outA = fd.first[0].kind == 0 ? prm.mse + (static_cast<size_t>(chrom) * prm.innerDests + fd.first[0].dest) * prm.maxNComp
		                                : prm.ostat + (static_cast<size_t>(chrom) * prm.outerDests + fd.first[0].dest) * prm.maxNComp * 2;
		fs.outB = fd.first[1].kind == 0 ? prm.mse + (static_cast<size_t>(chrom) * prm.innerDests + fd.first[1].dest) * prm.maxNComp
		                                : prm.ostat + (static_cast<size_t>(chrom) * prm.outerDests + fd.first[1].dest) * prm.maxNComp * 2;
		fs.rntr = fd.rntr;
		fs.syc = fd.syc;
#pragma unroll
		for (int u = 0; u < kKNS; ++u) {
			const int r = 32 * u + lane;
			if (r < n8) fs.rowBits |= 1u << u;
			if (fitActive) {
				if ((fd.train[u] >> lane) & 1u) fs.trainBits |= 1u << u;
				if ((fd.first[0].rows[u] >> lane) & 1u) fs.taskBitsA |= 1u << u;
				if ((fd.first[1].rows[u] >> lane) & 1u) fs.taskBitsB |= 1u << u;
			}
			fs.yc[u] = (r < n) ? ys[r] - fd.ym : 0.0; // response centred on the training rows (src/PLSSimpls.cpp:28-49)
			fs.pred[u] = 0.0;
		}
		// t_0 = Z S_0 = K (m . y_c)   (S_0 = X_c'y_c, :126)
#pragma unroll
		for (int u = 0; u < kKNS; ++u) {
			if ((fs.rowBits >> u) & 1u) taW[32 * u + lane] = ((fs.trainBits >> u) & 1u) ? fs.yc[u] : 0.0;
		}
		const int ct = (min(total - pass * kKNF, kKNF) + 7) >> 3; // column tiles of 8 fits in use
		nprod += static_cast<unsigned long long>(ct * Apass);
		__syncthreads();
		kernelProduct(Km, TA, TB, n8, ldk, warp, lane, ct);
		__syncthreads();
#pragma unroll
		for (int u = 0; u < kKNS; ++u) fs.tt[u] = ((fs.rowBits >> u) & 1u) ? tbW[32 * u + lane] : 0.0;
		KCtx cx;
		cx.prm = &prm; cx.Km = Km; cx.TA = TA; cx.TB = TB; cx.taW = taW; cx.tbW = tbW;
		cx.n8 = n8; cx.ldk = ldk; cx.warp = warp; cx.lane = lane; cx.A = A; cx.chrom = chrom;
		cx.fd = &fd;
		cx.fitActive = fitActive;
		cx.tracing = tracing && pass == 0;
		if (cx.tracing) GSB_KSTAMP(); // 2: first scores
		// the barriers below are uniform: Apass is a property of the pass.  (The loop bound is deliberately A, not Apass,
		// and kspaceRule is force-inlined: with either of the other forms nvcc 12.9 emits the scores and deflation code
		// twice -- 10.9 k instead of 7.7 k instructions -- and the instruction cache costs 4 % of the launch.)
#pragma unroll 1
		for (int comp = 0; comp < A; ++comp) {
			const bool more = comp + 1 < Apass;
			double tc[kKNS], st[4];
			kspaceScores<NF>(fs, cx, comp, more, tc, st);
			if (!more) break;
			// ---- w = K tc: the loading v = X't seen from the rows (:145-147) ---------------------------------------
			__syncthreads();
			if (cx.tracing) GSB_KSTAMP();
			kernelProduct(Km, TA, TB, n8, ldk, warp, lane, ct);
			__syncthreads();
			if (cx.tracing) GSB_KSTAMP();
			// one instance per component (measured: faster than three instances covering 0-2, 3-5, 6-8 with
			// predicated terms, although each then runs once per pass out of a cold instruction cache)
			switch (comp) {
			case 0: kspaceDeflate<NF, 0, 0>(fs, cx, comp, tc, st); break;
			case 1: kspaceDeflate<NF, 1, 1>(fs, cx, comp, tc, st); break;
			case 2: kspaceDeflate<NF, 2, 2>(fs, cx, comp, tc, st); break;
			case 3: kspaceDeflate<NF, 3, 3>(fs, cx, comp, tc, st); break;
			case 4: kspaceDeflate<NF, 4, 4>(fs, cx, comp, tc, st); break;
			case 5: kspaceDeflate<NF, 5, 5>(fs, cx, comp, tc, st); break;
			case 6: kspaceDeflate<NF, 6, 6>(fs, cx, comp, tc, st); break;
			case 7: kspaceDeflate<NF, 7, 7>(fs, cx, comp, tc, st); break;
			default: kspaceDeflate<NF, 8, 8>(fs, cx, comp, tc, st); break;
			}
			if (cx.tracing) GSB_KSTAMP();
		}
		// ill-conditioned (or apparently underflowing) here: the engine scores the chromosome again on the Gram-space
		// kernel, which takes its norms from explicit vectors, before anyone sees the result
		if ((fs.ill || fs.under) && fitActive && lane == 0 && prm.refit != nullptr) prm.refit[chrom] = 1;
	}
#undef GSB_KSTAMP
	if (tid == 0 && prm.products != nullptr) atomicAdd(prm.products, nprod);
	if (prm.guardFail != nullptr) {
		__syncthreads();
		if (tid < 8 && canary[tid] != kCanary) atomicAdd(prm.guardFail, 1);
	}
	if (NF != 8) {
		__syncthreads();
		if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmemBase));
	}
}

} // namespace

int kspaceKernelSharedBytes(int n, int nf) {
	return kCarve(roundUpK(n, 8), nf).total * static_cast<int>(sizeof(double));
}

// 1 when the kernel-space kernel with nf (8 or 16) fits per pass can take a plan with n rows and A components under
// the shared-memory limit
int kspaceKernelFits(int n, int A, int smemLimit, int nf) {
	const int n8 = roundUpK(n, 8);
	if (nf != 8 && nf != 16) return 0;
	if (n < 8 || n8 > 32 * kKNS || A > kKA || A < 1) return 0;
	if ((n8 >> 3) > 20) return 0; // 2x2 blocks of the lower triangle of K: at most 55
	return kspaceKernelSharedBytes(n, nf) <= smemLimit ? 1 : 0;
}

template <int NF, bool TRACE>
static cudaError_t launchKspaceOne(const KParams& prm, int bytes, long long grid, cudaStream_t stream) {
	cudaError_t err = cudaFuncSetAttribute(simplsKspaceKernel<NF, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
	if (err != cudaSuccess) return err;
	simplsKspaceKernel<NF, TRACE><<<static_cast<unsigned>(grid), 32 * NF, static_cast<size_t>(bytes), stream>>>(prm);
	return cudaGetLastError();
}

cudaError_t launchKspaceKernel(const KParams& prm, cudaStream_t stream, long long* ctas) {
	const int bytes = kspaceKernelSharedBytes(prm.n, prm.nf) + (prm.guardFail != nullptr ? 64 : 0);
	const long long grid = static_cast<long long>(prm.fullCount) + static_cast<long long>(prm.bucketCount - prm.fullCount) * prm.tailSplit;
	if (prm.bucketCount <= 0) return cudaSuccess;
	if (grid > 2147483647LL || prm.tailSplit < 1 || prm.fullCount < 0 || prm.fullCount > prm.bucketCount || (prm.nf != 8 && prm.nf != 16)) return cudaErrorInvalidConfiguration;
	const bool tr = prm.trace != nullptr;
	const cudaError_t err = prm.nf == 8 ? (tr ? launchKspaceOne<8, true>(prm, bytes, grid, stream) : launchKspaceOne<8, false>(prm, bytes, grid, stream))
	                                    : (tr ? launchKspaceOne<16, true>(prm, bytes, grid, stream) : launchKspaceOne<16, false>(prm, bytes, grid, stream));
	if (err != cudaSuccess) return err;
	if (ctas) *ctas += grid;
	return cudaSuccess;
}

} // namespace gsb
