// common.hpp -- error handling, RAII device buffers, timers shared by the host side of libgeneob200.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace geneo {

struct Error : public std::runtime_error {
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};

#define GENEO_CHECK(cond, msg)                                                                   \
  do {                                                                                           \
    if (!(cond)) throw ::geneo::Error(std::string("geneo_b200: ") + (msg) + " [" #cond "] at " + \
                                      __FILE__ + ":" + std::to_string(__LINE__));                \
  } while (0)

#define CUDA_CHECK(call)                                                                              \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess)                                                                           \
      throw ::geneo::Error(std::string("geneo_b200: CUDA error ") + cudaGetErrorString(e__) + " in " + \
                           #call + " at " + __FILE__ + ":" + std::to_string(__LINE__));               \
  } while (0)

inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Every kernel launch of the library goes through GENEO_TICK (wrapped around the grid argument of <<< >>>), so the
// number of launches inside a timed region is a counted fact (bench.py "gpu_launches"), not an estimate.
extern std::atomic<unsigned long long> g_kernel_launches;  // (launches are issued by more than one host thread)
extern unsigned long long g_h2d_bytes, g_d2h_bytes;
template <class G>
inline G launch_tick(G g) { g_kernel_launches.fetch_add(1, std::memory_order_relaxed); return g; }
// GENEO_PROFILE=1: a CUDA event is recorded on the (single, in-order) stream right before every launch, so consecutive
// events bracket each kernel (plus whatever memset / idle gap follows it); geneo_profile_dump() aggregates by launch site.
extern bool g_profile;
void profile_tick(const char* file, int line);
// every host wait of the library goes through here: in profile mode a marker event closes the interval of the last
// kernel, so that the idle gap while the host works is accounted to "host" and not to that kernel
extern thread_local double g_sync_wait_s;  // host time this THREAD spent waiting for the device (what is left of a phase is host-side work)
inline cudaError_t sync_stream(cudaStream_t s) {
  if (g_profile) profile_tick("<host wait / idle>", 0);
  const double t0 = now_s();
  const cudaError_t e = cudaStreamSynchronize(s);
  g_sync_wait_s += now_s() - t0;
  return e;
}
// GENEO_HOSTPROF=1: named host-side stopwatches (seconds, calls), printed by host_prof_report()
void host_prof_add(const char* name, double seconds);
void host_prof_report(const char* title);
struct HostProfScope {
  const char* name; double t0;
  explicit HostProfScope(const char* n) : name(n), t0(now_s()) {}
  ~HostProfScope() { host_prof_add(name, now_s() - t0); }
};
#define GENEO_TICK(g) ((::geneo::g_profile ? ::geneo::profile_tick(__FILE__, __LINE__) : (void)0), ::geneo::launch_tick(g))

// The product has NO CPU fallback: every numeric entry point calls this first.
inline void require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    throw Error("geneo_b200: no CUDA device available -- this library has no CPU fallback (" +
                std::string(cudaGetErrorString(e)) + ")");
}

// Device memory goes through a small stream-ordered block cache: every kernel and copy of the library runs on ONE in-order
// stream, so a released block can be handed to the next request without the implicit device synchronisation (and the
// page-table work) of cudaFree / cudaMalloc -- which otherwise sit between the factorizations of a re-setup.
void* dev_alloc(size_t bytes, size_t* cap);
void dev_free(void* p, size_t cap);
void dev_cache_flush();  // give every cached block back to the driver
// host -> device copy of a PAGEABLE buffer through two pinned bounce buffers (the driver's own staging of pageable memory
// moves ~2 GB/s on these boxes; memcpy into pinned memory overlapped with the DMA of the previous chunk moves 3-4x that).
// Like cudaMemcpyAsync from pageable memory, the source may be reused as soon as the call returns.
void h2d_staged(void* dst, const void* src, size_t bytes, cudaStream_t st);

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  size_t cap = 0;  // bytes of the underlying block
  DevBuf() {}
  explicit DevBuf(size_t n_) { alloc(n_); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = 0; o.cap = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; cap = o.cap; o.p = nullptr; o.n = 0; o.cap = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) dev_free(p, cap);
    p = nullptr; n = 0; cap = 0;
  }
  void alloc(size_t n_) {
    release();
    n = n_;
    if (n) p = static_cast<T*>(dev_alloc(n * sizeof(T), &cap));
  }
  void zero(cudaStream_t s = 0) { if (n) CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const T* h, size_t cnt, cudaStream_t s = 0) {
    if (cnt > n) alloc(cnt);
    if (cnt) h2d_staged(p, h, cnt * sizeof(T), s);
    g_h2d_bytes += cnt * sizeof(T);
  }
  void upload(const std::vector<T>& h, cudaStream_t s = 0) {
    if (h.size() != n) alloc(h.size());
    upload(h.data(), h.size(), s);
  }
  void download(T* h, size_t cnt, cudaStream_t s = 0) const {
    if (cnt) CUDA_CHECK(cudaMemcpyAsync(h, p, cnt * sizeof(T), cudaMemcpyDeviceToHost, s));
    g_d2h_bytes += cnt * sizeof(T);
    CUDA_CHECK(::geneo::sync_stream(s));
  }
  std::vector<T> to_host(cudaStream_t s = 0) const {
    std::vector<T> h(n);
    download(h.data(), n, s);
    return h;
  }
  size_t bytes() const { return n * sizeof(T); }
};

// Host CSR (square unless stated), 32-bit column indices, 64-bit row pointers where nnz may exceed 2^31.
struct CsrHost {
  int n = 0, ncols = 0;
  std::vector<int64_t> ptr;
  std::vector<int> idx;
  std::vector<double> val;
  int64_t nnz() const { return ptr.empty() ? 0 : ptr.back(); }
};

}  // namespace geneo
