// Microbenchmark: inner loops for the FP64 symmetric matrix-vector product w = G S with G in shared memory (sm_100a).
// A warp owns a 32-row block and walks NJ column blocks of 32 (one 32 x 32 tile each).
//   mode 0  row per lane (the r01/r02 kernel): lane = row, per column one LDS.64 of G (conflict-free) and half a broadcast
//           LDS.128 of S: 16 bytes delivered per FMA and lane
//   mode 1  XOR diagonal: lane l holds S[32 J + l] in a register and reads G[row l ^ j][column l] at step j, accumulating into
//           acc[j] (the row l ^ j): 8 bytes per FMA; the 32 accumulators are folded at the end of the row block by a butterfly
//           without selects (at mask m a lane keeps acc[j], j & m == 0, and receives the partner's acc[j | m]: the same row)
//   mode 2  XOR diagonal on a symmetric tile: every loaded entry also feeds the transposed product (S of the row block by
//           SHFL.BFLY with the static mask j): 2 FMAs per 8 bytes
//   mode 3  as mode 2, the transposed product's S element by a second (lane-distinct) LDS.64 instead of the shuffle
// Reported: FMAs per clock per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o symv_probe symv_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int LD = 32; // tile row stride in doubles (a multiple of 16: the bank depends on the lane only)

__global__ void kSymv(double* out, long long* clk, int iters, int mode, int nj) {
	extern __shared__ __align__(16) double sm[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int nw = blockDim.x >> 5;
	double* S = sm;                      // 1024 doubles
	double* G = sm + 1024;               // per warp: one 32 x 32 tile, re-used (bandwidth only)
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) S[i] = 1.0 / (1 + i);
	for (int i = threadIdx.x; i < nw * 1024; i += blockDim.x) G[i] = 1e-3 * (i % 977);
	__syncthreads();
	const double* gw = G + warp * 1024;
	double res = 0.0;
	long long t0 = clock64();
	for (int it = 0; it < iters; ++it) {
		if (mode == 0) {
			double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
			for (int J = 0; J < nj; ++J) {
				const double* gp = gw + lane; // [column][32 rows]: lane = row
				const double* sp = S + ((J * 32) & 1023);
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
			res += (a0 + a1) + (a2 + a3);
		} else {
			double acc[32];
#pragma unroll
			for (int j = 0; j < 32; ++j) acc[j] = 0.0;
			double accT = 0.0;
			const double sI = S[lane + 32 * (it & 7)];
			for (int J = 0; J < nj; ++J) {
				const double* gp = gw + lane; // [row][LD]: lane = column
				const double sJ = S[((J * 32) & 1023) + lane];
				if (mode == 1) {
#pragma unroll
					for (int j = 0; j < 32; ++j) acc[j] = fma(gp[(lane ^ j) * LD], sJ, acc[j]);
				} else if (mode == 2) {
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						const double g = gp[(lane ^ j) * LD];
						acc[j] = fma(g, sJ, acc[j]);
						accT = fma(g, __shfl_xor_sync(0xffffffffu, sI, j), accT);
					}
				} else {
					const double* sp = S + 32 * (it & 7);
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						const double g = gp[(lane ^ j) * LD];
						acc[j] = fma(g, sJ, acc[j]);
						accT = fma(g, sp[lane ^ j], accT);
					}
				}
			}
			// fold: at mask m a lane keeps acc[j], j < m, and receives the partner's acc[j + m] (the same row l ^ j)
#pragma unroll
			for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
				for (int j = 0; j < m; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j + m], m);
			}
			res += acc[0] + accT;
		}
	}
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = res;
	if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
	double* out;
	long long* clk;
	cudaMalloc(&out, 148 * 4 * 1024 * sizeof(double));
	cudaMalloc(&clk, sizeof(long long));
	const int iters = 200;
	const char* names[] = {"row per lane, S broadcast (16 B/FMA)", "XOR diagonal, S in a register (8 B/FMA)", "XOR diagonal symmetric, S by SHFL (4 B/FMA + SHFL)",
	                       "XOR diagonal symmetric, S by LDS.64 (8 B/FMA)"};
	const int cfgs[][2] = {{1, 128}, {1, 256}, {2, 128}, {2, 256}, {4, 128}, {3, 256}, {4, 256}};
	for (int nj : {4, 7}) {
		for (auto& cfg : cfgs) {
			const int ctas = cfg[0], threads = cfg[1];
			const int smem = (1024 + (threads / 32) * 1024) * 8;
			const int pad = 233472 / (ctas + 1) + 16; // exactly `ctas` CTAs per SM
			const int bytes = smem > pad ? smem : pad;
			cudaFuncSetAttribute(kSymv, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
			for (int mode = 0; mode < 4; ++mode) {
				kSymv<<<148 * ctas, threads, bytes>>>(out, clk, iters, mode, nj);
				cudaError_t err = cudaGetLastError();
				if (err == cudaSuccess) err = cudaDeviceSynchronize();
				if (err != cudaSuccess) { printf("ctas=%d threads=%d mode=%d: %s\n", ctas, threads, mode, cudaGetErrorString(err)); return 1; }
				long long c = 0;
				cudaMemcpy(&c, clk, sizeof(c), cudaMemcpyDeviceToHost);
				const double fmas = double(iters) * nj * 1024.0 * (threads / 32) * ctas * (mode >= 2 ? 2.0 : 1.0);
				printf("k=%3d CTAs/SM=%d threads=%3d  %-52s %6.1f FMA/clk/SM  (%6.0f clk per row block of a warp)\n", nj * 32, ctas, threads, names[mode],
				       fmas / double(c), double(c) / iters);
			}
		}
	}
	return 0;
}
