// nccl_dyn.cpp — see nccl_dyn.h
#include <dlfcn.h>
#include <stdio.h>
#include <mutex>
#include "nccl_dyn.h"

namespace rxb {
void set_error(const char *fmt, ...);

static NcclApi g_api;
static bool g_ok = false;
static char g_why[256] = "";
static std::once_flag g_once;

static void load_once()
{
	// an NCCL the process already mapped (e.g. torch's) is reused; otherwise the system library
	void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);
	if (!lib) { lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL); }
	if (!lib) { lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL); }
	if (!lib) { snprintf(g_why, sizeof g_why, "libnccl.so.2 not loadable: %s", dlerror()); return; }
	struct { const char *name; void **slot; } syms[] = {
		{"ncclGetUniqueId", (void **)&g_api.GetUniqueId}, {"ncclCommInitRank", (void **)&g_api.CommInitRank},
		{"ncclCommInitAll", (void **)&g_api.CommInitAll}, {"ncclCommDestroy", (void **)&g_api.CommDestroy},
		{"ncclAllGather", (void **)&g_api.AllGather}, {"ncclGroupStart", (void **)&g_api.GroupStart},
		{"ncclGroupEnd", (void **)&g_api.GroupEnd}, {"ncclGetErrorString", (void **)&g_api.GetErrorString},
		{"ncclGetVersion", (void **)&g_api.GetVersion},
	};
	for (auto &s : syms) {
		*s.slot = dlsym(lib, s.name);
		if (!*s.slot) { snprintf(g_why, sizeof g_why, "libnccl.so.2 lacks %s", s.name); return; }
	}
	g_ok = true;
}

const NcclApi *nccl_api()
{
	std::call_once(g_once, load_once);
	if (!g_ok) { set_error("%s", g_why); return nullptr; }
	return &g_api;
}

}  // namespace rxb
