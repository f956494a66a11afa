// nccl_dl.cpp — see nccl_dl.h.
#include "nccl_dl.h"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace {
struct UniqueId { char internal[128]; };  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
using comm_t = void*;
constexpr int kNcclFloat = 7;  // ncclFloat32

struct Api {
    void* handle = nullptr;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(comm_t*, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string load_error;
};
Api g_api;
std::once_flag g_once;

void load() {
    // BPT_NCCL_LIB names the library to load instead of the default sonames (a site-specific build; the tests use it
    // to exercise the NCCL-missing path)
    const char* override_name = getenv("BPT_NCCL_LIB");
    const char* names[] = {override_name && *override_name ? override_name : "libnccl.so.2",
                           override_name && *override_name ? nullptr : "libnccl.so"};
    std::string why;
    for (const char* n : names) {
        if (!n) continue;
        g_api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_api.handle) break;
        const char* e = dlerror();  // one call: dlerror() clears the message it returns
        if (why.empty()) why = e ? e : "unknown";
    }
    if (!g_api.handle) {
        g_api.load_error = std::string("cannot dlopen ") + names[0] + ": " + why;
        return;
    }
#define SYM(field, name)                                                              \
    g_api.field = reinterpret_cast<decltype(g_api.field)>(dlsym(g_api.handle, name)); \
    if (!g_api.field) { g_api.load_error = std::string("NCCL symbol missing: ") + name; return; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllGather, "ncclAllGather")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
}

bool ready(std::string* err) {
    std::call_once(g_once, load);
    if (!g_api.load_error.empty() || !g_api.AllGather) {
        if (err) *err = g_api.load_error.empty() ? "NCCL not loaded" : g_api.load_error;
        return false;
    }
    return true;
}
int check(int rc, const char* what, std::string* err) {
    if (rc == 0) return 0;
    if (err) *err = std::string(what) + ": " + g_api.GetErrorString(rc);
    return -1;
}
}  // namespace

int bpt_nccl_get_unique_id(uint8_t id[128], std::string* err) {
    if (!ready(err)) return -1;
    UniqueId u;
    if (check(g_api.GetUniqueId(&u), "ncclGetUniqueId", err)) return -1;
    std::memcpy(id, u.internal, 128);
    return 0;
}
int bpt_nccl_comm_init(void** comm, const uint8_t id[128], int rank, int nranks, std::string* err) {
    if (!ready(err)) return -1;
    UniqueId u;
    std::memcpy(u.internal, id, 128);
    comm_t c = nullptr;
    if (check(g_api.CommInitRank(&c, nranks, u, rank), "ncclCommInitRank", err)) return -1;
    *comm = c;
    return 0;
}
void bpt_nccl_comm_destroy(void* comm) {
    if (comm && g_api.CommDestroy) g_api.CommDestroy(comm);
}
int bpt_nccl_allgather_f32(void* comm, const void* send, void* recv, size_t count, cudaStream_t st, std::string* err) {
    if (!ready(err)) return -1;
    return check(g_api.AllGather(send, recv, count, kNcclFloat, comm, st), "ncclAllGather", err);
}
