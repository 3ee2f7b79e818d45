// fit_common.cuh -- device helpers shared by the Gram-space fit kernels (simpls_fit.cu; profiles/experiments/simpls_tmem.cu): warp / CTA
// reductions with bitwise-identical totals in every thread, cp.async, TMA bulk copies on an mbarrier.
#ifndef GASELECT_B200_FIT_COMMON_CUH
#define GASELECT_B200_FIT_COMMON_CUH

#include "kernels.cuh"

namespace gsb {

namespace {

__host__ __device__ inline int roundUp(int v, int m) { return (v + m - 1) / m * m; }

__device__ __forceinline__ double shflXor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

__device__ __forceinline__ void cpAsync16(double* dst, const double* src) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src));
}
__device__ __forceinline__ void cpAsyncWaitAll() {
	asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// TMA bulk copy (global -> shared, SASS UBLKCP) completing on an mbarrier
__device__ __forceinline__ uint32_t smemAddrF(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarInitF(uint64_t* bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smemAddrF(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbarExpectTxF(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smemAddrF(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWaitF(uint64_t* bar, uint32_t parity) {
	uint32_t done;
	do {
		asm volatile(
			"{\n"
			".reg .pred p;\n"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
			"selp.u32 %0, 1, 0, p;\n"
			"}\n" : "=r"(done) : "r"(smemAddrF(bar)), "r"(parity) : "memory");
	} while (!done);
}
__device__ __forceinline__ void bulkLoadF(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smemAddrF(dst)),
	             "l"(src), "r"(bytes), "r"(smemAddrF(bar))
	             : "memory");
}

// Sum 16 values per lane over the warp: afterwards every lane L holds the total of value (L >> 1).
__device__ __forceinline__ double warpSum16(double (&v)[16], int lane) {
#pragma unroll
	for (int h = 8; h >= 1; h >>= 1) {
		const bool up = (lane & (2 * h)) != 0;
#pragma unroll
		for (int i = 0; i < h; ++i) {
			const double send = up ? v[i] : v[i + h];
			const double keep = up ? v[i + h] : v[i];
			v[i] = keep + shflXor(send, 2 * h);
		}
	}
	return v[0] + shflXor(v[0], 1);
}

// CTA-wide totals of NV <= 16 values; every thread receives all of them, bitwise identical.
// STRIDE: doubles per warp of one of the two phases of `red` (a kernel that also uses wider reductions passes theirs)
template <int NV, int STRIDE = 16>
__device__ __forceinline__ void blockSum16(double (&x)[NV], double* red, int& phase, int nw, int lane, int warp) {
	double v[16];
#pragma unroll
	for (int i = 0; i < 16; ++i) v[i] = (i < NV) ? x[i] : 0.0;
	double tot = warpSum16(v, lane);
	if (nw > 1) {
		double* buf = red + phase * (nw * STRIDE);
		if ((lane & 1) == 0) buf[warp * STRIDE + (lane >> 1)] = tot;
		__syncthreads();
		tot = buf[lane & 15];
		for (int w = 1; w < nw; ++w) tot += buf[w * STRIDE + (lane & 15)];
		phase ^= 1;
#pragma unroll
		for (int i = 0; i < NV; ++i) x[i] = __shfl_sync(0xffffffffu, tot, i);
	} else {
#pragma unroll
		for (int i = 0; i < NV; ++i) x[i] = __shfl_sync(0xffffffffu, tot, 2 * i);
	}
}

// CTA-wide totals of TWO values (the second reduction round of a component: v'v and v'S): a plain butterfly -- 30
// instructions per warp where the 16-value form spends 108 on its selects, shuffles and broadcasts
template <int STRIDE = 16>
__device__ __forceinline__ void blockSum2(double (&x)[2], double* red, int& phase, int nw, int lane, int warp) {
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1) {
		x[0] += shflXor(x[0], m);
		x[1] += shflXor(x[1], m);
	}
	if (nw > 1) {
		double* buf = red + phase * (nw * STRIDE);
		if (lane == 0) *reinterpret_cast<double2*>(buf + warp * STRIDE) = make_double2(x[0], x[1]);
		__syncthreads();
		double2 t = *reinterpret_cast<const double2*>(buf);
		for (int w = 1; w < nw; ++w) {
			const double2 o = *reinterpret_cast<const double2*>(buf + w * STRIDE);
			t.x += o.x;
			t.y += o.y;
		}
		x[0] = t.x;
		x[1] = t.y;
		phase ^= 1;
	}
}

} // namespace

} // namespace gsb

#endif
