// Multi-GPU plumbing of the engine: one context per GPU, NCCL over NVLink / NVSwitch INSIDE the library, so that a host
// that can only make C calls (the reference is one JVM, SpeciesIdentifier.java:668-681) reaches the sharded path.
// NCCL is bound at run time (dlopen): the library has no link-time dependency on it and a single-GPU host never loads it;
// when the process already holds an NCCL (torch's bundled one in bench.py) that very copy is used.
#pragma once
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only

#include <cstdlib>
#include <string>

namespace tdna {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};

// nullptr + *why on failure.  The handle is process-wide and never closed.
inline const NcclApi *nccl_api(std::string *why) {
    static NcclApi api;
    static bool tried = false, ok = false;
    static std::string err;
    if (!tried) {
        tried = true;
        // The copy the process already uses, if any: two NCCLs of one SONAME cannot live in one process, and a host that loads
        // another one LATER (torch imports its bundled libnccl.so.2) would be bound to whatever was loaded first.  TDNA_NCCL_LIB
        // names a file to load instead of the system's (taxondna_b200/_lib.py preloads the pip-installed copy the same way).
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) {
            const char *path = getenv("TDNA_NCCL_LIB");
            if (path && *path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
        }
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            const char *e = dlerror();
            err = std::string("libnccl.so.2 not found: ") + (e ? e : "?");
        } else {
            api.handle = h;
            bool all = true;
#define TDNA_NCCL_SYM(field, name)                                        \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name)); \
    if (!api.field) { all = false; err = std::string("NCCL symbol missing: ") + name; }
            TDNA_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
            TDNA_NCCL_SYM(CommInitRank, "ncclCommInitRank")
            TDNA_NCCL_SYM(CommDestroy, "ncclCommDestroy")
            TDNA_NCCL_SYM(GetErrorString, "ncclGetErrorString")
            TDNA_NCCL_SYM(AllReduce, "ncclAllReduce")
            TDNA_NCCL_SYM(AllGather, "ncclAllGather")
            TDNA_NCCL_SYM(Send, "ncclSend")
            TDNA_NCCL_SYM(Recv, "ncclRecv")
            TDNA_NCCL_SYM(GroupStart, "ncclGroupStart")
            TDNA_NCCL_SYM(GroupEnd, "ncclGroupEnd")
            TDNA_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef TDNA_NCCL_SYM
            ok = all;
        }
    }
    if (!ok && why) *why = err;
    return ok ? &api : nullptr;
}

struct Comm {
    const NcclApi *api = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    bool enabled = true;  // TDNA_OPT_MULTI: 0 = whole-range calls run on this GPU alone although a communicator exists
    bool force = false;   // TDNA_OPT_MULTI = 2: the collective path even with ONE rank (how a one-GPU box tests the choreography)
    bool active() const { return comm != nullptr && enabled && (nranks > 1 || force); }
};

}  // namespace tdna
