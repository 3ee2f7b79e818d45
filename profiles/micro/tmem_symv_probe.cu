// Microbenchmark: tensor memory as the store of a Gram matrix for a matrix-vector product on sm_100a.
// Each thread streams ROWLEN doubles of its own TMEM lane (tcgen05.ld.32x32b) and multiplies them with a vector read from
// shared memory as 16-byte broadcast loads (the K1t inner loop), in several shapes:
//   mode 0: ld.x16 x4 + wait, no arithmetic        mode 1: ld.x32 x2 + wait, no arithmetic
//   mode 2: ld.x64 x1 + wait, no arithmetic        mode 3: ld.x32 x2 + wait + 32 DFMA (S from shared memory)
//   mode 4: ld.x64 + wait + 32 DFMA                mode 5: shared-memory baseline: 32 LDS.64 (conflict-free) + 32 DFMA
//   mode 6: ld.x32 double-buffered (next pair in flight during the DFMAs)
// Reported: bytes of matrix per clock per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_symv_probe tmem_symv_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define R8(r, o) "=r"(r[o]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
#define P8(r, o) "+r"(r[o]), "+r"(r[o + 1]), "+r"(r[o + 2]), "+r"(r[o + 3]), "+r"(r[o + 4]), "+r"(r[o + 5]), "+r"(r[o + 6]), "+r"(r[o + 7])

__device__ __forceinline__ void ld16(uint32_t* r, uint32_t addr) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
	             : R8(r, 0), R8(r, 8) : "r"(addr));
}
__device__ __forceinline__ void ld32(uint32_t* r, uint32_t addr) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
	             : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24) : "r"(addr));
}
__device__ __forceinline__ void ld64(uint32_t* r, uint32_t addr) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
	             "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n"
	             : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24), R8(r, 32), R8(r, 40), R8(r, 48), R8(r, 56) : "r"(addr));
}
__device__ __forceinline__ void waitLd() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tie32(uint32_t* r) { asm volatile("" : P8(r, 0), P8(r, 8), P8(r, 16), P8(r, 24)); }
__device__ __forceinline__ void st16(uint32_t addr, const uint32_t* r) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(addr),
	             "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
	             "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}
__device__ __forceinline__ double d2(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }

template <int NCOL>
__global__ void kProbe(double* out, long long* clk, int iters, int mode, int rowlen) {
	extern __shared__ __align__(16) double sm[];
	__shared__ uint32_t slot;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "n"(NCOL));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
	}
	double* S = sm;             // 512 doubles
	double* G = sm + 512;       // mode 5: 32 rows x rowlen per warp, row stride rowlen + 1 ... kept simple: [warp][c][32]
	for (int i = threadIdx.x; i < 512; i += blockDim.x) S[i] = 1.0 / (1 + i);
	for (int i = threadIdx.x; i < (int)(blockDim.x / 32) * 1024; i += blockDim.x) G[i] = 1e-3 * i;
	asm volatile("tcgen05.fence::before_thread_sync;\n");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;\n");
	const uint32_t base = slot;
	const int grp = warp >> 2;
	const int ngrp = (blockDim.x + 127) / 128;
	const int cols = NCOL / ngrp; // 32-bit columns of this thread's share of its lane
	const uint32_t mine = base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(grp * cols);
	{
		uint32_t r[16];
		for (int c = 0; c + 16 <= cols; c += 16) {
			for (int j = 0; j < 16; j += 2) { const double v = 1.0 + 1e-3 * (c + j + threadIdx.x); r[j] = (uint32_t)__double2loint(v); r[j + 1] = (uint32_t)__double2hiint(v); }
			st16(mine + c, r);
		}
		asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
	}
	__syncthreads();
	const int chunks = (2 * rowlen) / 64; // 64 columns = 32 doubles per step
	double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
	long long t0 = clock64();
	for (int it = 0; it < iters; ++it) {
		if (mode == 0) {
			for (int ch = 0; ch < chunks; ++ch) {
				uint32_t r[64];
				const uint32_t a = mine + (uint32_t)((ch * 64) % cols);
				ld16(r, a); ld16(r + 16, a + 16); ld16(r + 32, a + 32); ld16(r + 48, a + 48);
				waitLd(); tie32(r); tie32(r + 32);
				a0 += d2(r[0], r[1]); a1 += d2(r[34], r[35]);
			}
		} else if (mode == 1) {
			for (int ch = 0; ch < chunks; ++ch) {
				uint32_t r[64];
				const uint32_t a = mine + (uint32_t)((ch * 64) % cols);
				ld32(r, a); ld32(r + 32, a + 32);
				waitLd(); tie32(r); tie32(r + 32);
				a0 += d2(r[0], r[1]); a1 += d2(r[34], r[35]);
			}
		} else if (mode == 2) {
			for (int ch = 0; ch < chunks; ++ch) {
				uint32_t r[64];
				const uint32_t a = mine + (uint32_t)((ch * 64) % cols);
				ld64(r, a);
				waitLd(); tie32(r); tie32(r + 32);
				a0 += d2(r[0], r[1]); a1 += d2(r[34], r[35]);
			}
		} else if (mode == 3 || mode == 4) {
			for (int ch = 0; ch < chunks; ++ch) {
				uint32_t r[64];
				const uint32_t a = mine + (uint32_t)((ch * 64) % cols);
				if (mode == 3) { ld32(r, a); ld32(r + 32, a + 32); } else ld64(r, a);
				waitLd(); tie32(r); tie32(r + 32);
				const double* sp = S + ((ch * 32) & 255);
#pragma unroll
				for (int j = 0; j < 32; j += 4) {
					const double2 s01 = *reinterpret_cast<const double2*>(sp + j);
					const double2 s23 = *reinterpret_cast<const double2*>(sp + j + 2);
					a0 = fma(d2(r[2 * j], r[2 * j + 1]), s01.x, a0);
					a1 = fma(d2(r[2 * j + 2], r[2 * j + 3]), s01.y, a1);
					a2 = fma(d2(r[2 * j + 4], r[2 * j + 5]), s23.x, a2);
					a3 = fma(d2(r[2 * j + 6], r[2 * j + 7]), s23.y, a3);
				}
			}
		} else if (mode == 5) {
			const double* g = G + (size_t)warp * 1024 + lane; // the same 32 x 32 tile every step: bandwidth only
			for (int ch = 0; ch < chunks; ++ch) {
				const double* sp = S + ((ch * 32) & 255);
				const double* gp = g;
#pragma unroll
				for (int j = 0; j < 32; j += 4) {
					const double2 s01 = *reinterpret_cast<const double2*>(sp + j);
					const double2 s23 = *reinterpret_cast<const double2*>(sp + j + 2);
					a0 = fma(gp[j * 32], s01.x, a0);
					a1 = fma(gp[(j + 1) * 32], s01.y, a1);
					a2 = fma(gp[(j + 2) * 32], s23.x, a2);
					a3 = fma(gp[(j + 3) * 32], s23.y, a3);
				}
			}
		} else if (mode == 7) {
			// symmetric 32 x 32 tile per step: lane l holds G[row l ^ j][column l] (j = 0..31) in its TMEM lane; direct product into
			// acc[j], transposed product (S of the row block permuted in registers) into accT: 64 DFMA per 32 doubles, no LSU traffic
			double acc[32], sI[32];
#pragma unroll
			for (int j = 0; j < 32; ++j) { acc[j] = 0.0; sI[j] = S[(lane ^ j) + 32 * (it & 7)]; }
			double accT = 0.0, accU = 0.0, accV = 0.0, accW = 0.0;
			for (int ch = 0; ch < chunks; ++ch) {
				uint32_t r[64];
				const uint32_t a = mine + (uint32_t)((ch * 64) % cols);
				ld64(r, a);
				const double sJ = S[((ch * 32) & 255) + lane];
				waitLd(); tie32(r); tie32(r + 32);
#pragma unroll
				for (int j = 0; j < 32; j += 4) {
					const double g0 = d2(r[2 * j], r[2 * j + 1]), g1 = d2(r[2 * j + 2], r[2 * j + 3]);
					const double g2 = d2(r[2 * j + 4], r[2 * j + 5]), g3 = d2(r[2 * j + 6], r[2 * j + 7]);
					acc[j] = fma(g0, sJ, acc[j]);
					accT = fma(g0, sI[j], accT);
					acc[j + 1] = fma(g1, sJ, acc[j + 1]);
					accU = fma(g1, sI[j + 1], accU);
					acc[j + 2] = fma(g2, sJ, acc[j + 2]);
					accV = fma(g2, sI[j + 2], accV);
					acc[j + 3] = fma(g3, sJ, acc[j + 3]);
					accW = fma(g3, sI[j + 3], accW);
				}
			}
			accT = (accT + accU) + (accV + accW);
#pragma unroll
			for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
				for (int j = 0; j < m; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j + m], m);
			}
			a0 += acc[0] + accT;
		} else {
			uint32_t ra[32], rb[32];
			ld32(ra, mine);
			for (int ch = 0; ch < 2 * chunks; ch += 2) {
				waitLd(); tie32(ra);
				ld32(rb, mine + (uint32_t)(((ch + 1) * 32) % cols));
				const double* sp = S + ((ch * 16) & 255);
#pragma unroll
				for (int j = 0; j < 16; j += 4) {
					const double2 s01 = *reinterpret_cast<const double2*>(sp + j);
					const double2 s23 = *reinterpret_cast<const double2*>(sp + j + 2);
					a0 = fma(d2(ra[2 * j], ra[2 * j + 1]), s01.x, a0);
					a1 = fma(d2(ra[2 * j + 2], ra[2 * j + 3]), s01.y, a1);
					a2 = fma(d2(ra[2 * j + 4], ra[2 * j + 5]), s23.x, a2);
					a3 = fma(d2(ra[2 * j + 6], ra[2 * j + 7]), s23.y, a3);
				}
				waitLd(); tie32(rb);
				ld32(ra, mine + (uint32_t)(((ch + 2) * 32) % cols));
#pragma unroll
				for (int j = 0; j < 16; j += 4) {
					const double2 s01 = *reinterpret_cast<const double2*>(sp + 16 + j);
					const double2 s23 = *reinterpret_cast<const double2*>(sp + 16 + j + 2);
					a0 = fma(d2(rb[2 * j], rb[2 * j + 1]), s01.x, a0);
					a1 = fma(d2(rb[2 * j + 2], rb[2 * j + 3]), s01.y, a1);
					a2 = fma(d2(rb[2 * j + 4], rb[2 * j + 5]), s23.x, a2);
					a3 = fma(d2(rb[2 * j + 6], rb[2 * j + 7]), s23.y, a3);
				}
			}
			waitLd(); tie32(ra);
			a0 += d2(ra[0], ra[1]);
		}
	}
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = (a0 + a1) + (a2 + a3);
	if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
	asm volatile("tcgen05.fence::before_thread_sync;\n");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base), "n"(NCOL));
}

int main() {
	double* out;
	long long* clk;
	cudaMalloc(&out, 148 * 4 * 1024 * sizeof(double));
	cudaMalloc(&clk, sizeof(long long));
	const int iters = 200;
	const char* names[] = {"ld.x16 x4 + wait", "ld.x32 x2 + wait", "ld.x64 + wait", "ld.x32 x2 + wait + 32 DFMA", "ld.x64 + wait + 32 DFMA", "LDS.64 x32 + 32 DFMA (smem)", "ld.x32 double-buffered + DFMA", "sym. tile: ld.x64 + 64 DFMA, S in regs"};
	// (CTAs per SM, threads per CTA): TMEM columns per CTA = 512 / CTAs
	const int cfgs[][2] = {{1, 128}, {1, 256}, {2, 128}, {2, 256}, {4, 64}, {4, 128}};
	for (auto& cfg : cfgs) {
		const int ctas = cfg[0], threads = cfg[1];
		const int ncol = 512 / ctas;
		const int ngrp = (threads + 127) / 128;
		const int rowlen = (ncol / ngrp) / 2; // doubles per thread
		if (rowlen < 32) continue;
		const int smem = ctas == 1 ? 200 * 1024 : (ctas == 2 ? 100 * 1024 : 50 * 1024);
		for (int mode = 0; mode < 8; ++mode) {
			if (ncol == 512) { cudaFuncSetAttribute(kProbe<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); kProbe<512><<<148 * ctas, threads, smem>>>(out, clk, iters, mode, rowlen); }
			else if (ncol == 256) { cudaFuncSetAttribute(kProbe<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); kProbe<256><<<148 * ctas, threads, smem>>>(out, clk, iters, mode, rowlen); }
			else { cudaFuncSetAttribute(kProbe<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); kProbe<128><<<148 * ctas, threads, smem>>>(out, clk, iters, mode, rowlen); }
			cudaError_t err = cudaDeviceSynchronize();
			if (err != cudaSuccess) { printf("ctas=%d threads=%d mode=%d: %s\n", ctas, threads, mode, cudaGetErrorString(err)); return 1; }
			long long c = 0;
			cudaMemcpy(&c, clk, sizeof(c), cudaMemcpyDeviceToHost);
			const int chunks = (2 * rowlen) / 64;
			const double bytes = double(iters) * chunks * 32 * 8 * threads * ctas; // per SM
			printf("CTAs/SM=%d threads=%3d rowlen=%3d  %-40s %7.1f B/clk/SM  %5.1f FMA/clk/SM (%6.1f clk per 32-double step of a warp)\n", ctas, threads, rowlen, names[mode],
			       bytes / double(c), (mode == 7 ? 2.0 : (mode >= 3 ? 1.0 : 0.0)) * bytes / 8.0 / double(c), double(c) / (double(iters) * chunks));
		}
	}
	return 0;
}
