// fp64_peak.cu — measures the FP64 pipe's DFMA throughput of the GPU (SURVEY.md 8d asks for this denominator: MEASURED_PEAKS.json
// has HBM and bf16 figures only).  The gather kernels of the step (density, ECS, explicit forces, PPE assembly, gradient) are
// bound by FP64 issue (true IEEE sqrt and division per particle pair), so their roofline is this number, not the HBM one.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu && tools/fp64_peak
//
// Every thread runs 8 independent DFMA chains (enough ILP to cover the pipe latency at full occupancy); the result is written
// so that nothing is optimised away.  Prints DFMA/s, TFLOP/s (2 flops per DFMA) and, for orientation, DSQRT+DDIV pairs per second
// (the per-pair cost the gather kernels actually pay).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, const double a, const double b, const int iters)
{
	double x[8];
	for (int k = 0; k < 8; k++) x[k] = threadIdx.x * 1e-3 + k;
	for (int i = 0; i < iters; i++)
	{
#pragma unroll
		for (int k = 0; k < 8; k++) x[k] = fma(x[k], a, b);
	}
	double s = 0;
	for (int k = 0; k < 8; k++) s += x[k];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_sqrt_div(double* out, const double a, const int iters)
{
	double x[4];
	for (int k = 0; k < 4; k++) x[k] = 1.0 + threadIdx.x * 1e-3 + k;
	for (int i = 0; i < iters; i++)
	{
#pragma unroll
		for (int k = 0; k < 4; k++) x[k] = a / sqrt(x[k] + 1.0); // one r = sqrt(r2) and one r_e / r per pair, like weight()
	}
	double s = 0;
	for (int k = 0; k < 4; k++) s += x[k];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	const int blocks = p.multiProcessorCount * 8, threads = 256;
	double* out = nullptr;
	cudaMalloc(&out, sizeof(double) * blocks * threads);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	float ms = 0;

	const int iters = 1 << 16;
	k_dfma<<<blocks, threads>>>(out, 0.999999, 1e-6, 1024); // warm-up
	double best = 0;
	for (int rep = 0; rep < 5; rep++)
	{
		cudaEventRecord(e0);
		k_dfma<<<blocks, threads>>>(out, 0.999999, 1e-6, iters);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		cudaEventElapsedTime(&ms, e0, e1);
		const double rate = 8.0 * iters * blocks * threads / (ms * 1e-3);
		if (rate > best) best = rate;
	}
	std::printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_per_s\": %.4e, \"fp64_tflops\": %.2f", p.name, p.multiProcessorCount, best, 2.0 * best / 1e12);

	const int iters2 = 1 << 13;
	k_sqrt_div<<<blocks, threads>>>(out, 2.4, 256);
	best = 0;
	for (int rep = 0; rep < 5; rep++)
	{
		cudaEventRecord(e0);
		k_sqrt_div<<<blocks, threads>>>(out, 2.4, iters2);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		cudaEventElapsedTime(&ms, e0, e1);
		const double rate = 4.0 * iters2 * blocks * threads / (ms * 1e-3);
		if (rate > best) best = rate;
	}
	std::printf(", \"sqrt_div_pairs_per_s\": %.4e}\n", best);
	cudaFree(out);
	return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
