// common.cuh — context, error handling, device buffers and launch accounting.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "core_math.cuh"

namespace flgp {

// ---- errors: C++ exceptions inside, int status + flgp_last_error() at the C boundary -----------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] inline void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}
#define FLGP_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      ::flgp::fail(3, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__,    \
                   cudaGetErrorString(e_));                                                      \
  } while (0)

// ---- NCCL, bound at run time (dlopen) so that single-GPU users need no NCCL at all -------------
struct Nccl;  // comm.cu

struct StageRec {
  std::string name;
  cudaEvent_t beg, end;
  uint64_t launches;
  double flops, bytes;  // algorithmic work declared by the stage
};

struct Ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // multi-GPU (one process per GPU)
  int rank = 0, nranks = 1;
  Nccl* nccl = nullptr;
  // accounting
  uint64_t launches = 0;
  bool timing = false;
  std::vector<StageRec> stages;
  // pinned scratch for small device->host reads
  int64_t* pinned = nullptr;  // 64 words
};

// every kernel launch goes through this so that `gpu_launches` is counted, not guessed
#define FLGP_LAUNCH(ctx, kernel, grid, block, smem, ...)                                         \
  do {                                                                                           \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                             \
    (ctx)->launches++;                                                                           \
    FLGP_CUDA(cudaGetLastError());                                                               \
  } while (0)

// ---- device memory pool ----------------------------------------------------------------------------
// cudaMalloc/cudaFree synchronise the device and cost milliseconds for the 30-250 MB work buffers of a
// fit; a repeated fit asks for exactly the same sizes, so freed blocks are kept and handed back by size.
// All work of a process runs on one stream, which orders reuse; with more than one live context the pool
// is bypassed.  (capi.cu owns the singleton.)
// hostcopy.cu: host <-> device copies; large copies from / to pageable memory are staged by several host threads
void h2d_copy(void* d, const void* h, size_t bytes, cudaStream_t st);
void d2h_copy(void* h, const void* d, size_t bytes, cudaStream_t st);
void* pool_alloc(size_t bytes);
void pool_free(void* p, size_t bytes);
void pool_trim();
void pool_ctx_count(int delta);

// ---- RAII device buffer ---------------------------------------------------------------------------
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p;
      n = o.n;
      o.p = nullptr;
      o.n = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) p = static_cast<T*>(pool_alloc(count * sizeof(T)));
  }
  void release() {
    if (p) pool_free(p, n * sizeof(T));
    p = nullptr;
    n = 0;
  }
  void zero(cudaStream_t st) {
    if (n) FLGP_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), st));
  }
  void upload(const T* h, size_t count, cudaStream_t st) { h2d_copy(p, h, count * sizeof(T), st); }
  void download(T* h, size_t count, cudaStream_t st) const { d2h_copy(h, p, count * sizeof(T), st); }
};

inline void sync(Ctx* c) { FLGP_CUDA(cudaStreamSynchronize(c->stream)); }

// stage timing (CUDA events on the library's stream)
struct StageScope {
  Ctx* c;
  int idx = -1;
  StageScope(Ctx* ctx, const char* name, double flops = 0, double bytes = 0) : c(ctx) {
    if (!c->timing) return;
    StageRec r;
    r.name = name;
    r.flops = flops;
    r.bytes = bytes;
    FLGP_CUDA(cudaEventCreate(&r.beg));
    FLGP_CUDA(cudaEventCreate(&r.end));
    r.launches = c->launches;
    FLGP_CUDA(cudaEventRecord(r.beg, c->stream));
    c->stages.push_back(r);
    idx = (int)c->stages.size() - 1;
  }
  void stop() {
    if (idx < 0) return;
    StageRec& r = c->stages[idx];
    cudaEventRecord(r.end, c->stream);
    r.launches = c->launches - r.launches;
    idx = -1;
  }
  ~StageScope() { stop(); }
};

// ---- collectives (comm.cu): no-ops when nranks == 1 ---------------------------------------------
void comm_allreduce_i64(Ctx* c, int64_t* dbuf, size_t count);
void comm_allreduce_i64_to(Ctx* c, const int64_t* src, int64_t* dst, size_t count);  // out of place
void comm_allreduce_f64(Ctx* c, double* dbuf, size_t count);
void comm_allreduce_max_f64(Ctx* c, double* dbuf, size_t count);
void comm_allgather_f64(Ctx* c, const double* send, double* recv, size_t count_per_rank);
void comm_destroy(Ctx* c);

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace flgp
