// pool.cu — cache of device blocks keyed by (device, size) (declared in common.cuh): a block is only ever handed
// back on the device it was allocated on.
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace flgp {
namespace {
struct Pool {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void*> free_blocks;  // (device, bytes)
  size_t cached = 0;
  int live_ctx = 0;
  static constexpr size_t kLimit = (size_t)48 << 30;  // bytes kept at most
} g_pool;
size_t round_up(size_t b) { return (b + 511) & ~(size_t)511; }
int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}
}  // namespace

void* pool_alloc(size_t bytes) {
  bytes = round_up(bytes);
  const int dev = current_device();
  {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    if (g_pool.live_ctx <= 1) {
      auto it = g_pool.free_blocks.find(std::make_pair(dev, bytes));
      if (it != g_pool.free_blocks.end()) {
        void* p = it->second;
        g_pool.free_blocks.erase(it);
        g_pool.cached -= bytes;
        return p;
      }
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    pool_trim();  // give cached blocks back and retry once
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) fail(3, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
  return p;
}

void pool_free(void* p, size_t bytes) {
  bytes = round_up(bytes);
  cudaPointerAttributes at;
  int dev = current_device();
  if (cudaPointerGetAttributes(&at, p) == cudaSuccess) dev = at.device;  // the device that owns the block
  {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    if (g_pool.live_ctx <= 1 && g_pool.cached + bytes <= Pool::kLimit) {
      g_pool.free_blocks.emplace(std::make_pair(dev, bytes), p);
      g_pool.cached += bytes;
      return;
    }
  }
  cudaFree(p);
}

void pool_ctx_count(int delta) {
  std::lock_guard<std::mutex> lk(g_pool.mu);
  g_pool.live_ctx += delta;
}

void pool_trim() {
  std::multimap<std::pair<int, size_t>, void*> blocks;
  {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    blocks.swap(g_pool.free_blocks);
    g_pool.cached = 0;
  }
  for (auto& kv : blocks) cudaFree(kv.second);
}
}  // namespace flgp

