// Microbenchmark: tensor memory (TMEM) as per-thread scratch on sm_100a -- tcgen05.st / tcgen05.ld (32x32b) round
// trip, load latency and load / store throughput per SM.  Question behind it: could the nine stored vectors of a
// fit (K1k, 45 doubles per lane) live in TMEM instead of registers?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmemLd16(uint32_t (&r)[16], uint32_t addr) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
	             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
	               "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
	             : "r"(addr));
}
__device__ __forceinline__ void tmemSt16(uint32_t addr, const uint32_t (&r)[16]) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(addr),
	             "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
	             "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}
__device__ __forceinline__ void tmemWaitLd() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmemWaitSt() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// mode 0: round trip check; 1: dependent load latency; 2: load throughput (4 loads in flight); 3: store throughput
__global__ void kTmem(uint32_t* out, long long* clk, int iters, int mode) {
	__shared__ uint32_t baseSlot;
	const int warp = threadIdx.x >> 5;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&baseSlot))));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
	}
	asm volatile("tcgen05.fence::before_thread_sync;\n");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;\n");
	const uint32_t base = baseSlot;
	// a warp reaches the 32 lanes of its quadrant (warp % 4); the warps of a quadrant split the 512 columns
	const int wq = warp >> 2, nq = blockDim.x >> 7;   // warps per quadrant
	const int cols = 512 / (nq > 0 ? nq : 1);
	const uint32_t mine = base + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + static_cast<uint32_t>(wq * cols);
	uint32_t r[16], acc = 0;
	for (int j = 0; j < 16; ++j) r[j] = threadIdx.x * 131u + j;
	for (int c = 0; c + 16 <= cols; c += 16) {
		for (int j = 0; j < 16; ++j) r[j] = (threadIdx.x * 131u + j) ^ (c * 2654435761u);
		tmemSt16(mine + c, r);
	}
	tmemWaitSt();
	__syncthreads();
	long long t0 = clock64();
	if (mode == 0) {
		uint32_t bad = 0;
		for (int c = 0; c + 16 <= cols; c += 16) {
			tmemLd16(r, mine + c);
			tmemWaitLd();
			for (int j = 0; j < 16; ++j) bad += r[j] != ((threadIdx.x * 131u + j) ^ (c * 2654435761u));
		}
		acc = bad;
	} else if (mode == 1) {
		uint32_t c = 0;
		for (int i = 0; i < iters; ++i) {
			tmemLd16(r, mine + c);
			tmemWaitLd();
			c = (r[0] & 0u) + ((c + 16) % cols); // dependent address
			acc += r[5];
		}
	} else if (mode == 2) {
		uint32_t r1[16], r2[16], r3[16];
		for (int i = 0; i < iters; ++i) {
			const uint32_t c = (i * 64) % cols;
			tmemLd16(r, mine + c);
			tmemLd16(r1, mine + (c + 16) % cols);
			tmemLd16(r2, mine + (c + 32) % cols);
			tmemLd16(r3, mine + (c + 48) % cols);
			tmemWaitLd();
			acc += r[1] + r1[2] + r2[3] + r3[4];
		}
	} else {
		for (int i = 0; i < iters; ++i) {
			const uint32_t c = (i * 64) % cols;
			r[0] = i;
			tmemSt16(mine + c, r);
			tmemSt16(mine + (c + 16) % cols, r);
			tmemSt16(mine + (c + 32) % cols, r);
			tmemSt16(mine + (c + 48) % cols, r);
		}
		tmemWaitSt();
	}
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(base));
}

int main() {
	uint32_t* out;
	long long* clk;
	cudaMalloc(&out, 148 * 1024 * sizeof(uint32_t));
	cudaMalloc(&clk, sizeof(long long));
	const int iters = 2000;
	cudaFuncSetAttribute(kTmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	for (int threads : {128, 256, 512}) {
		for (int mode = 0; mode < 4; ++mode) {
			kTmem<<<148, threads, 200 * 1024>>>(out, clk, iters, mode); // 200 KB of shared memory: one CTA per SM
			cudaError_t err = cudaDeviceSynchronize();
			if (err != cudaSuccess) { printf("threads=%d mode=%d: %s\n", threads, mode, cudaGetErrorString(err)); return 1; }
			long long c = 0;
			cudaMemcpy(&c, clk, sizeof(c), cudaMemcpyDeviceToHost);
			if (mode == 0) {
				uint32_t h[1024];
				cudaMemcpy(h, out, threads * sizeof(uint32_t), cudaMemcpyDeviceToHost);
				unsigned bad = 0;
				for (int i = 0; i < threads; ++i) bad += h[i];
				printf("threads=%d round trip: %u mismatches\n", threads, bad);
			} else if (mode == 1) {
				printf("threads=%d dependent ld.x16 + wait: %.1f cycles\n", threads, double(c) / iters);
			} else {
				const double bytes = double(iters) * 4 * 16 * 4 * threads; // per CTA = per SM
				printf("threads=%d %s: %.1f B/clk/SM\n", threads, mode == 2 ? "ld.x16 x4 in flight" : "st.x16", bytes / double(c));
			}
		}
	}
	return 0;
}
