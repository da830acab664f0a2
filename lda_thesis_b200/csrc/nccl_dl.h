// NCCL through dlopen: the library has no link-time dependency on libnccl, so it loads (and the single-GPU
// path runs) on a box without NCCL, and inside a process that already loaded torch's bundled libnccl.so.2 the
// same copy is reused (dlopen resolves by SONAME).  Only what the sweep needs: unique id, comm init/destroy,
// int32 sum all-reduce in place on a stream.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string>

namespace nccl_dl {
struct UniqueId { char internal[128]; };
typedef int (*fn_get_unique_id)(UniqueId *);
typedef int (*fn_comm_init_rank)(void **, int, UniqueId, int);
typedef int (*fn_comm_destroy)(void *);
typedef int (*fn_all_reduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*fn_get_error_string)(int);

static void *g_lib = nullptr;
static fn_get_unique_id p_get_unique_id = nullptr;
static fn_comm_init_rank p_comm_init_rank = nullptr;
static fn_comm_destroy p_comm_destroy = nullptr;
static fn_all_reduce p_all_reduce = nullptr;
static fn_get_error_string p_get_error_string = nullptr;

static bool load(std::string *why) {
    if (g_lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD);          // already in the process (torch)?
        if (!lib) lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) { if (why) *why = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
    p_get_unique_id = (fn_get_unique_id)dlsym(lib, "ncclGetUniqueId");
    p_comm_init_rank = (fn_comm_init_rank)dlsym(lib, "ncclCommInitRank");
    p_comm_destroy = (fn_comm_destroy)dlsym(lib, "ncclCommDestroy");
    p_all_reduce = (fn_all_reduce)dlsym(lib, "ncclAllReduce");
    p_get_error_string = (fn_get_error_string)dlsym(lib, "ncclGetErrorString");
    if (!p_get_unique_id || !p_comm_init_rank || !p_comm_destroy || !p_all_reduce || !p_get_error_string) {
        if (why) *why = "libnccl.so.2 lacks a required symbol";
        return false;
    }
    g_lib = lib;
    return true;
}

static const char *error_string(int r) { return p_get_error_string ? p_get_error_string(r) : "NCCL not loaded"; }
static int get_unique_id(char id[128]) { return p_get_unique_id(reinterpret_cast<UniqueId *>(id)); }
static int comm_init_rank(void **comm, int nranks, const char id[128], int rank) {
    UniqueId u;
    for (int i = 0; i < 128; ++i) u.internal[i] = id[i];
    return p_comm_init_rank(comm, nranks, u, rank);
}
static int comm_destroy(void *comm) { return p_comm_destroy ? p_comm_destroy(comm) : 0; }
// ncclInt32 == 2, ncclSum == 0 (nccl.h, stable across NCCL 2.x)
static int all_reduce_i32(int *buf, size_t n, void *comm, cudaStream_t s) { return p_all_reduce(buf, buf, n, 2, 0, comm, s); }
}  // namespace nccl_dl
