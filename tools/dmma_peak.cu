// dmma_peak.cu -- does the FP64 tensor path (mma.sync m8n8k4 f64, SASS DMMA) of the B200 add
// throughput over the FP64 FMA pipe?  Decides whether the expanded-form cross term
// Sym = Y * M^T should be issued as DMMA tiles.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu && ./dmma_peak
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
	             : "+d"(c0), "+d"(c1)
	             : "d"(a), "d"(b));
}

// MODE 0: DMMA only; 1: DMMA + DFMA interleaved 1:NF; 2: DFMA only
template <int MODE, int CH, int NF>
__global__ void __launch_bounds__(256) k(double *out, int iters)
{
	double c[CH][2], f[NF > 0 ? NF : 1];
	const double a = threadIdx.x * 1e-3 + 1.0, b = threadIdx.x * 1e-4 + 0.5;
#pragma unroll
	for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = i;
#pragma unroll
	for (int i = 0; i < (NF > 0 ? NF : 1); ++i) f[i] = i;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < CH; ++i) {
			if (MODE != 2) dmma(c[i][0], c[i][1], a, b);
			if (MODE != 0) {
#pragma unroll
				for (int j = 0; j < NF; ++j) f[j] = fma(a, b, f[j]);
			}
		}
	}
	double s = 0;
#pragma unroll
	for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
#pragma unroll
	for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += f[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int CH, int NF>
void run(const char *name, int ctas_per_sm)
{
	int sms;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	double *out;
	cudaMalloc(&out, sizeof(double) * sms * ctas_per_sm * 256);
	const int iters = 4000;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	k<MODE, CH, NF><<<sms * ctas_per_sm, 256>>>(out, iters);
	cudaEventRecord(e0);
	k<MODE, CH, NF><<<sms * ctas_per_sm, 256>>>(out, iters);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double warps = (double)sms * ctas_per_sm * 8;
	const double fma_dmma = MODE == 2 ? 0 : warps * iters * CH * 256.0;              // 8x8x4 per DMMA
	const double fma_dfma = MODE == 0 ? 0 : warps * iters * CH * NF * 32.0;
	printf("%-26s acc=%d nf=%d ctas/sm=%d  %.3f ms  DMMA %.2f TFMA/s + DFMA %.2f TFMA/s = %.2f TFLOP/s\n",
	       name, CH, NF, ctas_per_sm, ms, fma_dmma / ms * 1e-9, fma_dfma / ms * 1e-9,
	       2 * (fma_dmma + fma_dfma) / ms * 1e-9);
	cudaFree(out);
}

int main()
{
	run<0, 8, 0>("DMMA only", 2);
	run<0, 8, 0>("DMMA only", 4);
	run<0, 16, 0>("DMMA only", 2);
	run<2, 8, 8>("DFMA only", 2);
	run<1, 8, 2>("DMMA + 2 DFMA", 2);
	run<1, 8, 4>("DMMA + 4 DFMA", 2);
	run<1, 8, 8>("DMMA + 8 DFMA", 2);
	run<1, 8, 8>("DMMA + 8 DFMA", 4);
	return 0;
}
