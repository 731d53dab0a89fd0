// zb_nccl.cpp -- NCCL through dlopen.  The library carries no link-time NCCL dependency: a process that
// already loaded NCCL (e.g. through torch) shares that copy, otherwise libnccl.so.2 is loaded on first use.
#include <dlfcn.h>

#include "zb_host.h"

namespace zb {

// ---- pooled device memory (see zb_host.h) ----
namespace {
struct PoolState {
    cudaStream_t stream = nullptr;
    bool ready = false;
};
PoolState g_pool[64];
PoolState* pool_for_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    PoolState& ps = g_pool[dev];
    if (!ps.ready) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) return nullptr;
        unsigned long long keep = ~0ull;  // never trim: freed buffers stay in the pool for the next allocation
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        if (cudaStreamCreateWithFlags(&ps.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        ps.ready = true;
    }
    return &ps;
}
}  // namespace

cudaError_t dev_alloc(void** p, size_t bytes) {
    PoolState* ps = pool_for_current_device();
    if (!ps) return cudaMalloc(p, bytes);
    cudaError_t e = cudaMallocAsync(p, bytes, ps->stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(ps->stream);  // the block is now usable on every stream
}
void dev_free(void* p) {
    if (!p) return;
    PoolState* ps = pool_for_current_device();
    cudaDeviceSynchronize();  // nothing in flight may still touch the block
    if (!ps || cudaFreeAsync(p, ps->stream) != cudaSuccess) cudaFree(p);
}

namespace {
struct ncclUniqueId_ {
    char internal[128];
};
typedef int (*fn_get_unique_id)(ncclUniqueId_*);
typedef int (*fn_comm_init_rank)(void**, int, ncclUniqueId_, int);
typedef int (*fn_comm_destroy)(void*);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_send)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_recv)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_group)(void);
typedef const char* (*fn_get_error_string)(int);

struct Api {
    void* handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_send send = nullptr;
    fn_recv recv = nullptr;
    fn_group group_start = nullptr, group_end = nullptr;
    fn_get_error_string get_error_string = nullptr;
};

Api& api() {
    static Api a;
    if (a.handle) return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        a.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) {
        // symbols may already be global in the process even if the soname differs
        if (dlsym(RTLD_DEFAULT, "ncclCommInitRank")) a.handle = dlopen(nullptr, RTLD_NOW);
    }
    for (const char* n : names) {
        if (a.handle) break;
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    ZB_REQUIRE(a.handle, ZB_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
    a.get_unique_id = (fn_get_unique_id)dlsym(a.handle, "ncclGetUniqueId");
    a.comm_init_rank = (fn_comm_init_rank)dlsym(a.handle, "ncclCommInitRank");
    a.comm_destroy = (fn_comm_destroy)dlsym(a.handle, "ncclCommDestroy");
    a.all_reduce = (fn_all_reduce)dlsym(a.handle, "ncclAllReduce");
    a.all_gather = (fn_all_gather)dlsym(a.handle, "ncclAllGather");
    a.send = (fn_send)dlsym(a.handle, "ncclSend");
    a.recv = (fn_recv)dlsym(a.handle, "ncclRecv");
    a.group_start = (fn_group)dlsym(a.handle, "ncclGroupStart");
    a.group_end = (fn_group)dlsym(a.handle, "ncclGroupEnd");
    a.get_error_string = (fn_get_error_string)dlsym(a.handle, "ncclGetErrorString");
    ZB_REQUIRE(a.get_unique_id && a.comm_init_rank && a.comm_destroy && a.all_reduce && a.all_gather && a.send && a.recv &&
                   a.group_start && a.group_end, ZB_ERR_COMM,
               "libnccl is missing required symbols");
    return a;
}

void check(int rc, const char* what) {
    if (rc != 0) {
        const char* msg = api().get_error_string ? api().get_error_string(rc) : "?";
        throw Error(ZB_ERR_COMM, fmt("%s failed: NCCL error %d (%s)", what, rc, msg));
    }
}
}  // namespace

void Nccl::unique_id(uint8_t* out128) {
    ncclUniqueId_ id;
    check(api().get_unique_id(&id), "ncclGetUniqueId");
    memcpy(out128, id.internal, 128);
}

void Nccl::init(const uint8_t* id128, int rank_, int world_, int device) {
    ZB_CUDA(cudaSetDevice(device));
    ncclUniqueId_ id;
    memcpy(id.internal, id128, 128);
    check(api().comm_init_rank(&comm, world_, id, rank_), "ncclCommInitRank");
    rank = rank_;
    world = world_;
}

void Nccl::destroy() {
    if (comm) api().comm_destroy(comm);
    comm = nullptr;
}

void Nccl::allreduce(void* d_buf, size_t count, Type t, Op op, cudaStream_t s) {
    if (!count) return;
    check(api().all_reduce(d_buf, d_buf, count, (int)t, (int)op, comm, s), "ncclAllReduce");
}

void Nccl::allgather(const void* d_send, void* d_recv, size_t bytes_per_rank, cudaStream_t s) {
    if (!bytes_per_rank) return;
    check(api().all_gather(d_send, d_recv, bytes_per_rank, (int)U8, comm, s), "ncclAllGather");
}

void Nccl::group_start() { check(api().group_start(), "ncclGroupStart"); }
void Nccl::group_end() { check(api().group_end(), "ncclGroupEnd"); }
void Nccl::send(const void* d_buf, size_t bytes, int peer, cudaStream_t s) {
    if (!bytes) return;
    check(api().send(d_buf, bytes, (int)U8, peer, comm, s), "ncclSend");
}
void Nccl::recv(void* d_buf, size_t bytes, int peer, cudaStream_t s) {
    if (!bytes) return;
    check(api().recv(d_buf, bytes, (int)U8, peer, comm, s), "ncclRecv");
}

}  // namespace zb
