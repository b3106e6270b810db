// common.cuh -- shared helpers of libmdns_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <utility>

#include "../../include/mdns_b200.h"

namespace mdns {

// ---- error plumbing ------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;
extern std::atomic<const char *> g_last_kernel;

#define MDNS_CUDA(call)                                                              \
	do {                                                                         \
		cudaError_t e__ = (call);                                            \
		if (e__ != cudaSuccess) {                                            \
			mdns::set_error("%s failed at %s:%d: %s", #call, __FILE__,   \
			                __LINE__, cudaGetErrorString(e__));          \
			return MDNS_ECUDA;                                           \
		}                                                                    \
	} while (0)

// Count a kernel launch and surface launch-configuration errors.
#define MDNS_LAUNCHED(name)                                                          \
	do {                                                                         \
		mdns::g_launches.fetch_add(1, std::memory_order_relaxed);            \
		mdns::g_last_kernel.store(name, std::memory_order_relaxed);          \
		cudaError_t e__ = cudaGetLastError();                                \
		if (e__ != cudaSuccess) {                                            \
			mdns::set_error("launch of %s failed at %s:%d: %s", name,    \
			                __FILE__, __LINE__, cudaGetErrorString(e__)); \
			return MDNS_ECUDA;                                           \
		}                                                                    \
	} while (0)

// The same for helper launches that should not show up as "the kernel of the last launch".
#define MDNS_LAUNCHED_HELPER(name)                                                   \
	do {                                                                         \
		mdns::g_launches.fetch_add(1, std::memory_order_relaxed);            \
		cudaError_t e__ = cudaGetLastError();                                \
		if (e__ != cudaSuccess) {                                            \
			mdns::set_error("launch of %s failed at %s:%d: %s", name,    \
			                __FILE__, __LINE__, cudaGetErrorString(e__)); \
			return MDNS_ECUDA;                                           \
		}                                                                    \
	} while (0)

// Programmatic dependent launch (round 2): a kernel launched with `pdl` may start while the
// kernel before it on the stream is still running, once every CTA of that kernel has executed
// pdl_trigger() (or exited); whatever it reads of the earlier kernel's output comes after its
// pdl_wait(), which returns when the earlier grid has completed and its writes are visible.
// Only kernels that call pdl_wait() unconditionally on every thread are ever launched this way.
// Inside a stream capture the dependency becomes a programmatic edge of the graph.
// MDNS_NO_PDL=1 launches everything fully serialised.
bool pdl_enabled();
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t st, bool pdl, Args &&...args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid;
	cfg.blockDim = block;
	cfg.dynamicSmemBytes = smem;
	cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = pdl && pdl_enabled() ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
#endif

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

#ifdef __CUDACC__
// ---- device helpers --------------------------------------------------------

// 128-bit streaming load of two doubles: read-only path, no L1 allocation
// (each data-set row is read exactly once per pass).
__device__ __forceinline__ double2 ldg_stream(const double2 *p)
{
	double2 v;
	asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
	             : "=d"(v.x), "=d"(v.y)
	             : "l"(p));
	return v;
}

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}

// mbarrier + 1-D bulk TMA copy (cp.async.bulk -> SASS UBLKCP): global -> shared.
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
	// make the initialised barrier visible to the async (TMA) proxy
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
	             "r"(bytes)
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra DONE_%=;\n"
	    "bra WAIT_%=;\n"
	    "DONE_%=:\n"
	    "}\n" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}
// bytes must be a multiple of 16; src and dst 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                            uint64_t *bar)
{
	asm volatile(
	    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
	        "r"(smem_u32(smem_dst)),
	    "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
	    : "memory");
}

__device__ __forceinline__ double shfl_xor_f64(double v, int lane_mask)
{
	return __shfl_xor_sync(0xffffffffu, v, lane_mask);
}
#endif  // __CUDACC__

}  // namespace mdns
