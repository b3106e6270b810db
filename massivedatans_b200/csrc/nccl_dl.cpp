// nccl_dl.cpp -- see nccl_dl.h
#include "nccl_dl.h"

#include <dlfcn.h>

#include <mutex>

namespace mdns {
void set_error(const char *fmt, ...);

const NcclApi *nccl_api()
{
	static NcclApi api;
	static bool tried = false, ok = false;
	static std::mutex mu;
	std::lock_guard<std::mutex> lock(mu);
	if (tried) {
		if (!ok) set_error("NCCL is not available (libnccl.so.2 could not be loaded)");
		return ok ? &api : nullptr;
	}
	tried = true;
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!h) {
		set_error("NCCL is not available: %s", dlerror());
		return nullptr;
	}
#define MDNS_SYM(field, name)                                                   \
	api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));      \
	if (!api.field) {                                                       \
		set_error("NCCL symbol %s is missing", name);                   \
		return nullptr;                                                 \
	}
	MDNS_SYM(GetVersion, "ncclGetVersion")
	MDNS_SYM(GetUniqueId, "ncclGetUniqueId")
	MDNS_SYM(CommInitRank, "ncclCommInitRank")
	MDNS_SYM(CommDestroy, "ncclCommDestroy")
	MDNS_SYM(AllReduce, "ncclAllReduce")
	MDNS_SYM(AllGather, "ncclAllGather")
	MDNS_SYM(GetErrorString, "ncclGetErrorString")
#undef MDNS_SYM
	ok = true;
	return &api;
}

}  // namespace mdns
