// zb_host.h -- host-side utilities of libzebra_b200: errors, device buffers, NCCL shim.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/zebra_b200.h"

namespace zb {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline std::string fmt(const char* f, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return std::string(buf);
}

#define ZB_CUDA(expr)                                                                                              \
    do {                                                                                                           \
        cudaError_t e__ = (expr);                                                                                  \
        if (e__ != cudaSuccess)                                                                                    \
            throw zb::Error(e__ == cudaErrorMemoryAllocation ? ZB_ERR_OOM : ZB_ERR_CUDA,                            \
                            zb::fmt("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__)); \
    } while (0)
#define ZB_REQUIRE(cond, code, ...) \
    do {                            \
        if (!(cond)) throw zb::Error((code), zb::fmt(__VA_ARGS__)); \
    } while (0)

extern thread_local std::string t_last_error;  // message of the calling thread's last failure (zb_last_error)

// every extern "C" entry point: exceptions become a zb_status + message
#define ZB_API_BEGIN try {
#define ZB_API_END                                   \
    return ZB_OK;                                    \
    }                                                \
    catch (const zb::Error& e) {                     \
        zb::t_last_error = e.what();                 \
        return e.code;                               \
    }                                                \
    catch (const std::bad_alloc&) {                  \
        zb::t_last_error = "host allocation failed"; \
        return ZB_ERR_OOM;                           \
    }                                                \
    catch (const std::exception& e) {                \
        zb::t_last_error = e.what();                 \
        return ZB_ERR_INVALID;                       \
    }

extern size_t g_device_bytes;  // not thread-exact; per-process accounting for zb_stats

// Device memory comes from the device's stream-ordered pool with the release threshold lifted, so that the buffers of
// a destroyed index (or of a grown workspace) are recycled instead of going back to the driver: cudaMalloc / cudaFree
// of multi-GB buffers cost milliseconds each and dominated the bulk build.  dev_alloc returns memory that is ready for
// use on any stream; dev_free waits for the device first (what cudaFree did implicitly).
cudaError_t dev_alloc(void** p, size_t bytes);
void dev_free(void* p);

// Grow-only device buffer.
template <class T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0;
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    ~DBuf() { release(); }
    void release() {
        if (p) {
            dev_free(p);
            g_device_bytes -= cap * sizeof(T);
        }
        p = nullptr;
        cap = 0;
    }
    size_t bytes() const { return cap * sizeof(T); }
    // Ensure capacity for n elements.  keep > 0: preserve the first `keep` elements (device copy on `s`).
    void ensure(size_t n, size_t keep = 0, cudaStream_t s = 0, bool exact = false) {
        if (n <= cap) return;
        size_t ncap = exact ? n : (cap ? cap + cap / 2 : n);
        if (ncap < n) ncap = n;
        T* np = nullptr;
        cudaError_t e = dev_alloc((void**)&np, ncap * sizeof(T));
        if (e != cudaSuccess && ncap > n) {  // retry with the exact size before giving up
            cudaGetLastError();
            ncap = n;
            e = dev_alloc((void**)&np, ncap * sizeof(T));
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            throw Error(ZB_ERR_OOM, fmt("cudaMalloc of %zu bytes failed: %s", ncap * sizeof(T), cudaGetErrorString(e)));
        }
        if (p && keep) {
            ZB_CUDA(cudaMemcpyAsync(np, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s));
            ZB_CUDA(cudaStreamSynchronize(s));
        }
        if (p) {
            dev_free(p);
            g_device_bytes -= cap * sizeof(T);
        }
        p = np;
        cap = ncap;
        g_device_bytes += cap * sizeof(T);
    }
};

// Pinned host staging buffer (grow-only).
template <class T>
struct HBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~HBuf() {
        if (p) cudaFreeHost(p);
    }
    void ensure(size_t n) {
        if (n <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        size_t ncap = n + n / 4;
        ZB_CUDA(cudaMallocHost(&p, ncap * sizeof(T)));
        cap = ncap;
    }
};

// ---- NCCL, loaded at run time (no link-time dependency; reuses the copy a host framework already loaded) ----
struct Nccl {
    void* comm = nullptr;
    int rank = 0, world = 1;
    static void unique_id(uint8_t* out128);
    void init(const uint8_t* id128, int rank, int world, int device);
    void destroy();
    enum Type { I32 = 2, U32 = 3, I64 = 4, U64 = 5, U8 = 1 };
    enum Op { SUM = 0, MAX = 2, MIN = 3 };
    void allreduce(void* d_buf, size_t count, Type t, Op op, cudaStream_t s);
    void allgather(const void* d_send, void* d_recv, size_t bytes_per_rank, cudaStream_t s);
    // point-to-point exchange (bucket-major store build): calls between group_start / group_end form one fused operation
    void group_start();
    void group_end();
    void send(const void* d_buf, size_t bytes, int peer, cudaStream_t s);
    void recv(void* d_buf, size_t bytes, int peer, cudaStream_t s);
};

}  // namespace zb
