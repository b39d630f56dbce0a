// Exchange step of the row-sharded sparse path: all-reduce(sum) of float64 buffers over
// NCCL, callable from a host that is NOT torch (the jax.ffi binding, a C++ trainer).  Reference: in GPJax the sum over data
// rows is the `jnp.sum` / matmul contraction inside collapsed_elbo (gpjax/objectives.py:380-398); sharding it over devices
// turns that contraction into exactly one all-reduce of the (M+2)^2 statistics and one of the gradient vector.
//
// NCCL is resolved at run time from the copy ALREADY loaded in the process (dlopen RTLD_NOLOAD on its SONAME, so a
// communicator created by the host framework's NCCL is used with that same NCCL), else loaded by name.  Nothing here links
// against libnccl: the library builds and loads on a box without it, and then every entry point returns GPB_ERR_UNSUPPORTED.
#include <dlfcn.h>

#include <atomic>
#include <cstddef>
#include <cstdint>
#include <mutex>

#include "../../include/gpjax_b200.h"

namespace {
// nccl.h (2.x), restated: ncclResult_t ncclSuccess = 0; ncclDataType_t ncclFloat64 = 8; ncclRedOp_t ncclSum = 0;
// ncclUniqueId = struct { char internal[128]; } passed BY VALUE to ncclCommInitRank.
struct UniqueId { char internal[128]; };
typedef int (*GetUniqueIdFn)(UniqueId*);
typedef int (*CommInitRankFn)(void**, int, UniqueId, int);
typedef int (*CommDestroyFn)(void*);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, void*);
typedef int (*GetVersionFn)(int*);
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

struct Nccl {
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    AllReduceFn all_reduce = nullptr;
    GetVersionFn get_version = nullptr;
    bool ok = false;
};

const Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOLOAD | RTLD_LAZY);  // the host framework's copy, if one is mapped
        if (!h) h = dlopen("libnccl.so.2", RTLD_LAZY | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_LAZY | RTLD_LOCAL);
        if (!h) return;
        n.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(h, "ncclGetUniqueId"));
        n.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(h, "ncclCommInitRank"));
        n.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(h, "ncclCommDestroy"));
        n.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(h, "ncclAllReduce"));
        n.get_version = reinterpret_cast<GetVersionFn>(dlsym(h, "ncclGetVersion"));
        n.ok = n.get_unique_id && n.comm_init_rank && n.comm_destroy && n.all_reduce;
    });
    return n;
}
}  // namespace

extern "C" {

int gpb_nccl_version(void) {
    const Nccl& n = nccl();
    int v = 0;
    if (!n.ok || !n.get_version || n.get_version(&v) != 0) return 0;
    return v;
}

int gpb_nccl_unique_id(void* id_out_128_bytes) {
    const Nccl& n = nccl();
    if (!n.ok) return GPB_ERR_UNSUPPORTED;
    if (!id_out_128_bytes) return GPB_ERR_INVALID;
    return n.get_unique_id(static_cast<UniqueId*>(id_out_128_bytes)) == 0 ? GPB_OK : GPB_ERR_LAUNCH;
}

int gpb_nccl_comm_init_rank(void** comm_out, int nranks, const void* id_128_bytes, int rank) {
    const Nccl& n = nccl();
    if (!n.ok) return GPB_ERR_UNSUPPORTED;
    if (!comm_out || !id_128_bytes || nranks < 1 || rank < 0 || rank >= nranks) return GPB_ERR_INVALID;
    UniqueId id = *static_cast<const UniqueId*>(id_128_bytes);
    return n.comm_init_rank(comm_out, nranks, id, rank) == 0 ? GPB_OK : GPB_ERR_LAUNCH;
}

int gpb_nccl_comm_destroy(void* comm) {
    const Nccl& n = nccl();
    if (!n.ok) return GPB_ERR_UNSUPPORTED;
    if (!comm) return GPB_ERR_INVALID;
    return n.comm_destroy(comm) == 0 ? GPB_OK : GPB_ERR_LAUNCH;
}

int gpb_allreduce_f64(void* comm, void* stream, double* buf, int64_t count) {
    const Nccl& n = nccl();
    if (!n.ok) return GPB_ERR_UNSUPPORTED;
    if (!comm || count < 0 || (count > 0 && !buf)) return GPB_ERR_INVALID;
    if (count == 0) return GPB_OK;
    return n.all_reduce(buf, buf, (size_t)count, kNcclFloat64, kNcclSum, comm, stream) == 0 ? GPB_OK : GPB_ERR_LAUNCH;
}

}  // extern "C"
