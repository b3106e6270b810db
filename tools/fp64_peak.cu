// fp64_peak.cu -- calibrates the FP64 pipe of the B200 for the likelihood kernels' op mix.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int CH>
__global__ void __launch_bounds__(256) k(double *out, double m, int iters)
{
	double acc[CH], y[CH];
#pragma unroll
	for (int c = 0; c < CH; ++c) {
		acc[c] = threadIdx.x * 1e-9 + c;
		y[c] = threadIdx.x * 1e-3 + c;
	}
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int c = 0; c < CH; ++c) {
			if (MODE == 0) {            // DFMA only
				acc[c] = fma(y[c], y[c], acc[c]);
			} else if (MODE == 1) {     // DADD + DFMA (the chi-square inner op)
				const double d = m - y[c];
				acc[c] = fma(d, d, acc[c]);
				y[c] = d;           // keep the DADD live and dependent per chain
			} else {                    // DADD feeding DFMA, DADD independent of the chain
				const double d = m - y[c];
				acc[c] = fma(d, d, acc[c]);
			}
		}
		if (MODE == 2) m += 1e-9;
	}
	double s = 0;
#pragma unroll
	for (int c = 0; c < CH; ++c) s += acc[c] + y[c];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int CH>
void run(const char *name, int ctas_per_sm)
{
	int sms;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	double *out;
	cudaMalloc(&out, sizeof(double) * sms * ctas_per_sm * 256);
	const int iters = 20000;
	cudaEvent_t a, b;
	cudaEventCreate(&a);
	cudaEventCreate(&b);
	k<MODE, CH><<<sms * ctas_per_sm, 256>>>(out, 0.5, iters);
	cudaEventRecord(a);
	k<MODE, CH><<<sms * ctas_per_sm, 256>>>(out, 0.5, iters);
	cudaEventRecord(b);
	cudaEventSynchronize(b);
	float ms;
	cudaEventElapsedTime(&ms, a, b);
	const double ops = (double)sms * ctas_per_sm * 256 * iters * CH * (MODE == 0 ? 1 : 2);
	printf("%-28s chains=%d ctas/sm=%d  %.3f ms  %.2f Tinstr/s  (%.1f per clk per SM at 1.9 GHz)\n", name,
	       CH, ctas_per_sm, ms, ops / ms * 1e-9, ops / (ms * 1e-3) / sms / 1.9e9);
	cudaFree(out);
}

int main()
{
	run<0, 8>("DFMA only", 2);
	run<0, 8>("DFMA only", 4);
	run<0, 16>("DFMA only", 2);
	run<1, 8>("DADD->DFMA dependent", 2);
	run<1, 8>("DADD->DFMA dependent", 4);
	run<2, 8>("DADD+DFMA", 2);
	run<2, 8>("DADD+DFMA", 4);
	run<2, 16>("DADD+DFMA", 2);
	run<2, 16>("DADD+DFMA", 1);
	return 0;
}
