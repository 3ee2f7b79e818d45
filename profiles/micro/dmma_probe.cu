// Microbenchmark: FP64 mma.sync m8n8k4 (DMMA) and DFMA issue rate / latency on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void kDmma(double* out, long long* clk, int iters) {
	double c0[CH], c1[CH];
	for (int j = 0; j < CH; ++j) { c0[j] = threadIdx.x; c1[j] = j; }
	double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
	__syncthreads();
	long long t0 = clock64();
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int j = 0; j < CH; ++j) dmma(c0[j], c1[j], a, b);
	}
	long long t1 = clock64();
	double s = 0;
	for (int j = 0; j < CH; ++j) s += c0[j] + c1[j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int CH>
__global__ void kDfma(double* out, long long* clk, int iters) {
	double c[CH];
	for (int j = 0; j < CH; ++j) c[j] = threadIdx.x + j;
	double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
	__syncthreads();
	long long t0 = clock64();
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int j = 0; j < CH; ++j) c[j] = fma(c[j], a, b);
	}
	long long t1 = clock64();
	double s = 0;
	for (int j = 0; j < CH; ++j) s += c[j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <class F>
void run(const char* name, F kern, int threads, int chains, int iters) {
	double* out; long long* clk; long long h;
	cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&clk, 8);
	kern<<<148, threads>>>(out, clk, iters);
	kern<<<148, threads>>>(out, clk, iters);
	cudaDeviceSynchronize();
	cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
	double per = double(h) / (double(iters) * chains);
	printf("%-6s warps/SM=%2d chains=%2d: %8.1f clk per iteration-of-all-chains, %6.2f clk per instruction per warp, %6.2f clk per instruction per SMSP\n",
	       name, threads / 32, chains, double(h) / iters, per, per / ((threads / 32 + 3) / 4));
	cudaFree(out); cudaFree(clk);
}
int main() {
	const int it = 2000;
	for (int w : {1, 4, 8, 16, 32}) {
		run("DMMA", kDmma<1>, 32 * w, 1, it);
		run("DMMA", kDmma<2>, 32 * w, 2, it);
		run("DMMA", kDmma<4>, 32 * w, 4, it);
		run("DMMA", kDmma<8>, 32 * w, 8, it);
	}
	for (int w : {1, 4, 8, 16, 32}) {
		run("DFMA", kDfma<1>, 32 * w, 1, it);
		run("DFMA", kDfma<4>, 32 * w, 4, it);
		run("DFMA", kDfma<8>, 32 * w, 8, it);
	}
	return 0;
}
