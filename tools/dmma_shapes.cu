// dmma_shapes.cu -- FP64 tensor-path throughput of the B200 by mma.sync shape: m8n8k4 (what
// round 1 used) against the sm_90+ shapes m16n8k4 / m16n8k8 / m16n8k16.  Same arithmetic per
// FMA; what differs is how many FMAs one issued instruction (and one set of operand registers)
// carries: 256 / 512 / 1024 / 2048.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_shapes dmma_shapes.cu && ./dmma_shapes
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE>
__device__ __forceinline__ void mma(double (&c)[4], const double (&a)[8], const double (&b)[4])
{
	if (SHAPE == 0) {          // 2 x m8n8k4 (rows g and g+8 as two instructions), k = 4
		asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
		             : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
		asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
		             : "+d"(c[2]), "+d"(c[3]) : "d"(a[1]), "d"(b[0]));
	} else if (SHAPE == 1) {   // m16n8k4
		asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
		             : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
	} else if (SHAPE == 2) {   // m16n8k8
		asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
		             : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
		             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
	} else {                   // m16n8k16
		asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
		             "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
		             : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
		             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
		               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
	}
}

template <int SHAPE, int CH>
__global__ void __launch_bounds__(256) k(double *out, int iters)
{
	double c[CH][4], a[8], b[4];
#pragma unroll
	for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + 1.0 + i;
#pragma unroll
	for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 1e-4 + 0.5 + i;
#pragma unroll
	for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = i;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < CH; ++i) mma<SHAPE>(c[i], a, b);
	}
	double s = 0;
#pragma unroll
	for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int CH>
void run(const char *name, int ctas_per_sm, int kdepth)
{
	int sms;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	double *out;
	cudaMalloc(&out, sizeof(double) * sms * ctas_per_sm * 256);
	const int iters = 4000 * 4 / kdepth;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	k<SHAPE, CH><<<sms * ctas_per_sm, 256>>>(out, iters);
	cudaEventRecord(e0);
	k<SHAPE, CH><<<sms * ctas_per_sm, 256>>>(out, iters);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double warps = (double)sms * ctas_per_sm * 8;
	const double fmas = warps * iters * CH * 16.0 * 8.0 * kdepth;
	printf("%-12s acc tiles=%d ctas/sm=%d  %.3f ms  %.2f TFMA/s = %.2f TFLOP/s\n", name, CH, ctas_per_sm, ms,
	       fmas / ms * 1e-9, 2 * fmas / ms * 1e-9);
	cudaFree(out);
}

int main()
{
	run<0, 4>("2 x m8n8k4", 2, 4);
	run<0, 8>("2 x m8n8k4", 2, 4);
	run<1, 4>("m16n8k4", 2, 4);
	run<1, 8>("m16n8k4", 2, 4);
	run<2, 4>("m16n8k8", 2, 8);
	run<2, 8>("m16n8k8", 2, 8);
	run<3, 2>("m16n8k16", 2, 16);
	run<3, 4>("m16n8k16", 2, 16);
	run<3, 8>("m16n8k16", 2, 16);
	run<3, 4>("m16n8k16", 1, 16);
	return 0;
}
