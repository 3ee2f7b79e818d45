// simpls_tmem.cu -- K1t: the per-fit Gram-space SIMPLS kernel with the fit's Gram held in TENSOR MEMORY.
//
// Same arithmetic and the same outputs as the fit-only instance of simplsFitFast (simpls_fit.cu): one CTA per
// (chromosome, training set), A <= 10 components of src/PLSSimpls.cpp:106-166 in Gram space, the rows
// [coef; -1; intercepts] left for predictBlocksKernel.  What changes is where the k x k Gram lives while the ten
// matrix-vector products w = G S run.
//
// The product is one multiply-add per Gram entry with nothing to reuse (a fit has ONE vector S per component and no two
// fits share a Gram), so with the Gram in shared memory it is bound by the 128 B/clk of the shared-memory crossbar:
// 16 FMA/clk/SM against the 58 the FP64 pipe issues.  Tensor memory is the one other on-chip store large enough for a
// Gram -- 128 lanes x 512 32-bit columns = 32 K doubles per SM, the full square of k = 181 -- and tcgen05.ld streams it at
// 415-640 B/clk/SM (measured, profiles/micro/tmem_probe_b200.txt).  No tcgen05.mma is involved (there is no FP64 kind):
// TMEM serves as a thread-private scratch, 32x32b shape, thread i of warp w <-> lane 32 (w % 4) + i.
//
//   row r < 128       thread r, lane r:       the whole row, k doubles at column 0
//   row r >= 128      thread r, lane r - 128: its first cB doubles behind the first group's (cB = what is left of the
//                     256 doubles of the lane), the columns [cB, k) from the packed slab in shared memory as before
//
// The slab (one contiguous packed lower triangle written by gramPackKernel) arrives by TMA bulk copies; every thread then
// expands its row into its lane (tcgen05.st), once per fit: k^2 shared-memory reads, the cost of ONE of the old products.
// Per component: tcgen05.ld x32 pairs (32 doubles per wait), S as 16-byte broadcast loads, four accumulators.
//
// Instances (engine: one launch per subset-size class):
//   128 columns, k <= 64:   up to 4 CTAs per SM      256 columns, k <= 128: 2 CTAs per SM
//   512 columns, 128 < k <= ~232 (whole slab in shared memory): 1 CTA per SM
// The dynamic shared-memory request is padded so that no more CTAs become resident than tensor memory can serve.
#include "kernels.cuh"
#include "fit_common.cuh"

namespace gsb {

namespace {

constexpr int kTA = 10; // components held in registers (the fast path's bound)

#define GSB_R8(r, o) "=r"(r[o]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
#define GSB_W8(r, o) "r"(r[o]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7])

// 16 doubles (32 columns) of this thread's lane
__device__ __forceinline__ void tmemLd32(uint32_t (&r)[32], uint32_t addr) {
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		"{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
		: GSB_R8(r, 0), GSB_R8(r, 8), GSB_R8(r, 16), GSB_R8(r, 24)
		: "r"(addr));
}
// 8 doubles (16 columns)
__device__ __forceinline__ void tmemSt16(uint32_t addr, const uint32_t (&r)[16]) {
	asm volatile(
		"tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(addr),
		GSB_W8(r, 0), GSB_W8(r, 8));
}
// The loaded registers are written asynchronously until the wait: tie them to it, so that no use is scheduled before it.
__device__ __forceinline__ void tmemWaitLd32(uint32_t (&a)[32]) {
	asm volatile("tcgen05.wait::ld.sync.aligned;\n"
	             : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]),
	               "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]), "+r"(a[16]), "+r"(a[17]), "+r"(a[18]),
	               "+r"(a[19]), "+r"(a[20]), "+r"(a[21]), "+r"(a[22]), "+r"(a[23]), "+r"(a[24]), "+r"(a[25]), "+r"(a[26]), "+r"(a[27]),
	               "+r"(a[28]), "+r"(a[29]), "+r"(a[30]), "+r"(a[31])
	             :
	             : "memory");
}
__device__ __forceinline__ void tmemWaitStT() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ double pairToDouble(uint32_t lo, uint32_t hi) { return __hiloint2double(static_cast<int>(hi), static_cast<int>(lo)); }

struct TmemCarve {
	int P, S, red, total;
};
// packed slab, S (zero padded to the next multiple of 32 columns past k), reduction scratch
__host__ __device__ inline TmemCarve tmemCarve(int k, int nw) {
	TmemCarve c;
	int o = 0;
	c.P = o; o += roundUp(k * (k + 1) / 2, 2);
	c.S = o; o += roundUp(k, 32) + 32;
	c.red = o; o += 2 * nw * 16;
	c.total = o;
	return c;
}

// columns [cstart, k) of row r of w = G S from the packed triangle in shared memory (cstart a multiple of 16, below the
// warp's first row): the three warp-uniform phases of symvRow (simpls_fit.cu)
__device__ __forceinline__ double symvRowFrom(const double* __restrict__ P, const double* __restrict__ S, int k, int r, int rlo, int rhi, int cstart) {
	const double* __restrict__ pr = P + r * (r + 1) / 2;
	const double* __restrict__ pc = P + r;
	double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
	int c = cstart;
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
	int tc = c * (c + 1) / 2;
	for (; c <= rhi; ++c) {
		const double g = (c <= r) ? pr[c] : pc[tc];
		a1 = fma(g, S[c], a1);
		tc += c + 1;
	}
	for (; c + 3 < k; c += 4) {
		const int t1 = tc + c + 1, t2 = t1 + c + 2, t3 = t2 + c + 3;
		a0 = fma(pc[tc], S[c], a0);
		a1 = fma(pc[t1], S[c + 1], a1);
		a2 = fma(pc[t2], S[c + 2], a2);
		a3 = fma(pc[t3], S[c + 3], a3);
		tc = t3 + c + 4;
	}
	for (; c < k; ++c) {
		a3 = fma(pc[tc], S[c], a3);
		tc += c + 1;
	}
	return (a0 + a1) + (a2 + a3);
}

// NCOL tensor-memory columns per CTA (128 / 256 / 512): 512 / NCOL CTAs per SM
template <int NCOL, bool TRACE>
__global__ void __launch_bounds__(NCOL / 2, 512 / NCOL) simplsFitTmem(const FitParams prm) {
	extern __shared__ __align__(16) double smem[];
	constexpr int LANE_DOUBLES = NCOL / 2;
	const int NT = blockDim.x, NW = NT >> 5;

	const int ci = blockIdx.x % prm.bucketCount;
	const int fi = blockIdx.x / prm.bucketCount;
	const int chrom = prm.bucket[ci];
	const uint32_t selBegin = prm.offsets[chrom];
	const int k = static_cast<int>(prm.offsets[chrom + 1] - selBegin);
	const int A = min(prm.maxNComp, k);
	const FitDesc fd = prm.fits[fi];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	const TmemCarve cv = tmemCarve(k, NW);
	double* __restrict__ P = smem + cv.P;
	double* __restrict__ S = smem + cv.S;
	double* red = smem + cv.red;
	int phase = 0;
	__shared__ __align__(8) uint64_t gramBar;
	__shared__ uint32_t tmemSlot;

	const bool tracing = TRACE && prm.trace != nullptr && blockIdx.x == gridDim.x / 2 && tid == 0;
	int stamp = 0;
#define GSB_TSTAMP() do { if (TRACE && tracing) prm.trace[stamp++] = clock64(); } while (0)
	GSB_TSTAMP();

	// ---- the fit's packed Gram: one contiguous slab (gramPackKernel), fetched by TMA bulk copies ------------------
	const int kp = roundUp(k, 2);
	const size_t slab = static_cast<size_t>(packedDoubles(static_cast<unsigned long long>(k)));
	const double* __restrict__ gsrc = prm.gpack + prm.packOff[chrom] + static_cast<size_t>(fi) * slab;
	const double* __restrict__ vsrc = prm.pvec + prm.vecOff[chrom] + static_cast<size_t>(fi) * 2 * kp;
	if (tid == 0) {
		mbarInitF(&gramBar, 1);
		const uint32_t bytes = static_cast<uint32_t>(slab * sizeof(double));
		mbarExpectTxF(&gramBar, bytes);
		for (uint32_t o = 0; o < bytes; o += 32768u) {
			bulkLoadF(reinterpret_cast<char*>(P) + o, reinterpret_cast<const char*>(gsrc) + o, min(32768u, bytes - o), &gramBar);
		}
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smemAddrF(&tmemSlot)), "n"(NCOL));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
	}
	const bool valid = tid < k;
	const int r = valid ? tid : k - 1;
	double sOwn, s0Own, xmOwn;
	{
		s0Own = __ldg(vsrc + r);
		xmOwn = __ldg(vsrc + kp + r);
		sOwn = s0Own;
		if (valid) S[r] = sOwn;
		for (int i = k + tid; i < roundUp(k, 32) + 32; i += NT) S[i] = 0.0; // the products run over whole 16-column pieces
	}
	double vOwn[kTA], cOwn[kTA];
#pragma unroll
	for (int j = 0; j < kTA; ++j) { vOwn[j] = 0.0; cOwn[j] = 0.0; }
	asm volatile("tcgen05.fence::before_thread_sync;\n");
	__syncthreads(); // S; the barrier's initialisation; the tensor-memory address
	asm volatile("tcgen05.fence::after_thread_sync;\n");
	const uint32_t tmemBase = tmemSlot;
	mbarWaitF(&gramBar, 0u); // the packed Gram has landed
	GSB_TSTAMP(); // 1: slab in shared memory

	// rows of a warp are consecutive: [rlo, rhi] is warp-uniform
	const int rlo = min(warp * 32, k - 1);
	const int rhi = min(rlo + 31, k - 1);
	const bool live = warp * 32 < k; // warps without rows skip the products
	// this thread's piece of its lane: group 0 (rows < 128) the whole row, group 1 what is left of the lane
	const int grp = warp >> 2;
	const int k16 = roundUp(k, 16);
	const int cT = grp == 0 ? k : max(0, min(k, (LANE_DOUBLES - k16) & ~15)); // columns [0, cT) of the row in tensor memory
	const int cTpad = roundUp(cT, 16);
	const uint32_t tbase = tmemBase + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(grp == 0 ? 0 : 2 * k16);

	// ---- expand the row into tensor memory (once per fit) --------------------------------------------------------
	if (live) {
		const double* __restrict__ pr = P + r * (r + 1) / 2;
		const double* __restrict__ pc = P + r;
		for (int c0 = 0; c0 < cTpad; c0 += 8) {
			uint32_t w8[16];
			int tc = c0 * (c0 + 1) / 2;
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				const int c = c0 + j;
				double g = 0.0;
				if (c < cT) g = (c <= r) ? pr[c] : pc[tc];
				tc += c + 1;
				w8[2 * j] = static_cast<uint32_t>(__double2loint(g));
				w8[2 * j + 1] = static_cast<uint32_t>(__double2hiint(g));
			}
			tmemSt16(tbase + static_cast<uint32_t>(2 * c0), w8);
		}
		tmemWaitStT();
	}
	GSB_TSTAMP(); // 2: Gram in tensor memory

	const double nanv = __longlong_as_double(0x7ff8000000000000LL);

	// ---- SIMPLS components -------------------------------------------------------------------
	double cum = 0.0;
#pragma unroll 1
	for (int comp = 0; comp < A; ++comp) {
		double w = 0.0;
		if (live) {
			double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
			for (int c0 = 0; c0 < cTpad; c0 += 32) {
				uint32_t ga[32], gb[32];
				const bool two = c0 + 16 < cTpad; // warp-uniform
				tmemLd32(ga, tbase + static_cast<uint32_t>(2 * c0));
				if (two) tmemLd32(gb, tbase + static_cast<uint32_t>(2 * c0 + 32));
				else {
#pragma unroll
					for (int j = 0; j < 32; ++j) gb[j] = 0u;
				}
				tmemWaitLd32(ga);
				tmemWaitLd32(gb);
				const double* __restrict__ sp = S + c0;
#pragma unroll
				for (int j = 0; j < 16; j += 4) {
					const double2 s01 = *reinterpret_cast<const double2*>(sp + j);
					const double2 s23 = *reinterpret_cast<const double2*>(sp + j + 2);
					a0 = fma(pairToDouble(ga[2 * j], ga[2 * j + 1]), s01.x, a0);
					a1 = fma(pairToDouble(ga[2 * j + 2], ga[2 * j + 3]), s01.y, a1);
					a2 = fma(pairToDouble(ga[2 * j + 4], ga[2 * j + 5]), s23.x, a2);
					a3 = fma(pairToDouble(ga[2 * j + 6], ga[2 * j + 7]), s23.y, a3);
				}
				if (two) {
#pragma unroll
					for (int j = 0; j < 16; j += 4) {
						const double2 s01 = *reinterpret_cast<const double2*>(sp + 16 + j);
						const double2 s23 = *reinterpret_cast<const double2*>(sp + 16 + j + 2);
						a0 = fma(pairToDouble(gb[2 * j], gb[2 * j + 1]), s01.x, a0);
						a1 = fma(pairToDouble(gb[2 * j + 2], gb[2 * j + 3]), s01.y, a1);
						a2 = fma(pairToDouble(gb[2 * j + 4], gb[2 * j + 5]), s23.x, a2);
						a3 = fma(pairToDouble(gb[2 * j + 6], gb[2 * j + 7]), s23.y, a3);
					}
				}
			}
			w = (a0 + a1) + (a2 + a3);
			if (cT < k) w += symvRowFrom(P, S, k, r, rlo, rhi, cT); // rows >= 128: the rest of the row from the slab
		}

		// ONE round: tn^2 = S'GS (:138), q*tn = s0'S (:148), projections V_j'w on every stored loading
		double r1[2 + kTA];
		r1[0] = valid ? sOwn * w : 0.0;
		r1[1] = valid ? s0Own * sOwn : 0.0;
#pragma unroll
		for (int j = 0; j < kTA; ++j) r1[2 + j] = (valid && j < comp) ? vOwn[j] * w : 0.0;
		blockSum16<2 + kTA>(r1, red, phase, NW, lane, warp);
		const bool under = !(r1[0] >= 1e-50); // ||t|| < 1e-25 (NORM_TOL, :16,:141-143); also NaN / negative rounding
		const double rtn = rsqrt(r1[0]);      // 1/tn; tn itself is never needed
		const double q = r1[1] * rtn;
		double v = w;
#pragma unroll
		for (int j = 0; j < kTA; ++j) v = fma(-vOwn[j], r1[2 + j], v);
		// cumulative coefficients from S BEFORE deflation (:152-156); v = X't carries 1/tn (:145-147)
		v *= rtn;
		cum += sOwn * (q * rtn);
		double r2[2];
		r2[0] = valid ? v * v : 0.0;
		r2[1] = valid ? v * sOwn : 0.0;
		blockSum16<2>(r2, red, phase, NW, lane, warp); // second round: deflate S, normalise v (:160, :70-88)
		const double rvn = rsqrt(r2[0]);
		const double dd = (r2[1] * rvn) * rvn;
		sOwn = fma(-v, dd, sOwn);
		if (under) { sOwn = nanv; cum = nanv; } // the reference throws: poison this and every later component
		if (valid) S[r] = sOwn;
#pragma unroll
		for (int j = 0; j < kTA; ++j) {
			if (j == comp) { vOwn[j] = v * rvn; cOwn[j] = cum; }
		}
		__syncthreads();
		GSB_TSTAMP(); // 3 + comp
	}

	// ---- the rows [coef; -1; intercepts] of this fit leave the kernel (predict_blocks.cu) ---------------
	double* __restrict__ cdst = prm.coefScratch + prm.coefOff[chrom] + static_cast<size_t>(fi) * (static_cast<size_t>(k) + 2) * A;
	if (valid) {
#pragma unroll
		for (int j = 0; j < kTA; ++j) if (j < A) cdst[static_cast<size_t>(r) * A + j] = cOwn[j];
	}
	// intercepts b0[a] = ym - xm'coef[:,a] (:162)
	double b0s[kTA];
#pragma unroll
	for (int j = 0; j < kTA; ++j) b0s[j] = (valid && j < A) ? xmOwn * cOwn[j] : 0.0;
	blockSum16<kTA>(b0s, red, phase, NW, lane, warp);
	if (tid == 0) {
#pragma unroll
		for (int j = 0; j < kTA; ++j) {
			if (j < A) {
				cdst[static_cast<size_t>(k) * A + j] = -1.0;
				cdst[static_cast<size_t>(k + 1) * A + j] = fd.ym - b0s[j];
			}
		}
	}
	GSB_TSTAMP();
#undef GSB_TSTAMP
	asm volatile("tcgen05.fence::before_thread_sync;\n");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmemBase), "n"(NCOL));
}

// dynamic shared memory: what the fit needs, padded so that at most 512 / ncol CTAs fit an SM
int tmemSmemBytes(int kMax, int threads, int ncol) {
	const int need = tmemCarve(kMax, threads / 32).total * static_cast<int>(sizeof(double));
	const int ctas = 512 / ncol;
	const int pad = 233472 / (ctas + 1) + 16;
	return need > pad ? need : pad;
}

template <int NCOL>
cudaError_t launchTmemOne(const FitParams& prm, int kMax, int smemLimit, cudaStream_t stream, long long* ctas) {
	const int threads = roundUp(kMax, 32);
	const int bytes = tmemSmemBytes(kMax, threads, NCOL);
	if (bytes > smemLimit) return cudaErrorInvalidValue;
	const long long grid = static_cast<long long>(prm.nfits) * prm.bucketCount;
	if (grid <= 0) return cudaSuccess;
	if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
	cudaError_t err;
	if (prm.trace != nullptr) {
		err = cudaFuncSetAttribute(simplsFitTmem<NCOL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
		if (err != cudaSuccess) return err;
		simplsFitTmem<NCOL, true><<<static_cast<unsigned>(grid), threads, static_cast<size_t>(bytes), stream>>>(prm);
	} else {
		err = cudaFuncSetAttribute(simplsFitTmem<NCOL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
		if (err != cudaSuccess) return err;
		simplsFitTmem<NCOL, false><<<static_cast<unsigned>(grid), threads, static_cast<size_t>(bytes), stream>>>(prm);
	}
	if (ctas) *ctas += grid;
	return cudaGetLastError();
}

} // namespace

// A <= 10 components and a subset whose whole packed Gram fits in shared memory next to S
bool tmemKernelServes(int kMax, int A, int smemLimit) {
	const int Ak = (A < kMax) ? A : kMax;
	if (Ak > kTA || kMax < 1 || kMax > 256) return false;
	const int ncol = kMax <= 64 ? 128 : (kMax <= 128 ? 256 : 512);
	return tmemSmemBytes(kMax, roundUp(kMax, 32), ncol) <= smemLimit;
}

cudaError_t launchTmemKernel(const FitParams& prm, int kMax, int A, int smemLimit, cudaStream_t stream, long long* ctas) {
	if (!tmemKernelServes(kMax, A, smemLimit) || prm.coefScratch == nullptr) return cudaErrorInvalidValue;
	if (kMax <= 64) return launchTmemOne<128>(prm, kMax, smemLimit, stream, ctas);
	if (kMax <= 128) return launchTmemOne<256>(prm, kMax, smemLimit, stream, ctas);
	return launchTmemOne<512>(prm, kMax, smemLimit, stream, ctas);
}

} // namespace gsb
