// hostcopy.cu — large copies between PAGEABLE host memory and the device.
//
// The reference-side caller (the R shim, numpy arrays) hands the C ABI ordinary pageable buffers.  The driver stages
// such a copy through its own pinned buffer on one thread: ~13 GB/s on this host against ~55 GB/s for pinned memory,
// i.e. 23 of the 65 ms of a C4 fit+predict from pageable memory.  Here the staging is done by several host threads,
// each with two pinned bounce buffers and its own stream: thread t takes chunks t, t+T, ...; while one chunk is on the
// bus the thread fills (or drains) its other buffer.  Pinned or registered host memory and small copies take the plain
// cudaMemcpyAsync.  Ordering: the copy streams wait for everything already queued on the caller's stream, and the
// caller's stream waits for the copies; a download returns when the data is in the destination.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace flgp {

namespace {

constexpr size_t HC_MIN = (size_t)8 << 20;  // smaller copies are left to the driver
constexpr int HC_MAXT = 16;
size_t hc_chunk() {  // bytes per bounce buffer
  static const size_t v = (size_t)(std::getenv("FLGP_HC_CHUNK_MB") ? std::atoi(std::getenv("FLGP_HC_CHUNK_MB")) : 4) << 20;
  return v;
}

struct Lane {
  char* bounce[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool used[2] = {false, false};
  cudaStream_t st = nullptr;
};
struct Stager {
  std::mutex mu;
  bool ready = false;
  int nl = 0;
  Lane lane[HC_MAXT];
  cudaEvent_t gate = nullptr;
};
Stager g_stager[64];  // one per device ordinal

bool staged_enabled() {
  static const bool on = std::getenv("FLGP_NO_STAGED_COPY") == nullptr;
  return on;
}

// true for memory the CUDA driver knows nothing about (malloc, R vectors, numpy arrays)
bool is_pageable(const void* h) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, h) != cudaSuccess) {
    cudaGetLastError();  // older drivers report unregistered memory as an error
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

void stager_init(Stager& s) {  // called with s.mu held
  if (!s.ready) {
    const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
    s.nl = (int)std::min<unsigned>(8, std::max(2u, hw / 2));
    if (std::getenv("FLGP_HC_LANES")) s.nl = std::max(1, std::min(HC_MAXT, std::atoi(std::getenv("FLGP_HC_LANES"))));
    FLGP_CUDA(cudaEventCreateWithFlags(&s.gate, cudaEventDisableTiming));
    for (int t = 0; t < s.nl; ++t) {
      FLGP_CUDA(cudaStreamCreateWithFlags(&s.lane[t].st, cudaStreamNonBlocking));
      for (int b = 0; b < 2; ++b) {
        FLGP_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&s.lane[t].bounce[b]), hc_chunk(), cudaHostAllocDefault));
        FLGP_CUDA(cudaEventCreateWithFlags(&s.lane[t].ev[b], cudaEventDisableTiming));
      }
    }
    s.ready = true;
  }
}

// runs body(t) on nl host threads bound to `dev`; the first error is re-thrown in the caller
template <class F>
void run_lanes(int dev, int nl, F body) {
  std::vector<std::string> err(nl);
  std::vector<std::thread> th;
  for (int t = 0; t < nl; ++t)
    th.emplace_back([&, t] {
      try {
        FLGP_CUDA(cudaSetDevice(dev));
        body(t);
      } catch (const std::exception& e) {
        err[t] = e.what();
      }
    });
  for (auto& x : th) x.join();
  for (int t = 0; t < nl; ++t)
    if (!err[t].empty()) fail(3, "%s", err[t].c_str());
}

}  // namespace

void h2d_copy(void* d, const void* h, size_t bytes, cudaStream_t st) {
  if (bytes < HC_MIN || !staged_enabled() || !is_pageable(h)) {
    FLGP_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st));
    return;
  }
  int dev = 0;
  FLGP_CUDA(cudaGetDevice(&dev));
  Stager& s = g_stager[dev & 63];
  std::lock_guard<std::mutex> lk(s.mu);
  stager_init(s);
  FLGP_CUDA(cudaEventRecord(s.gate, st));  // the destination may still be in use by work queued on st
  const size_t HC_CHUNK = hc_chunk();
  const size_t nchunks = (bytes + HC_CHUNK - 1) / HC_CHUNK;
  const int nl = (int)std::min<size_t>(s.nl, nchunks);
  run_lanes(dev, nl, [&](int t) {
    Lane& L = s.lane[t];
    FLGP_CUDA(cudaStreamWaitEvent(L.st, s.gate, 0));
    int b = 0;
    for (size_t k = t; k < nchunks; k += nl, b ^= 1) {
      const size_t off = k * HC_CHUNK, len = std::min(HC_CHUNK, bytes - off);
      if (L.used[b]) FLGP_CUDA(cudaEventSynchronize(L.ev[b]));  // the buffer's previous transfer has left it
      std::memcpy(L.bounce[b], static_cast<const char*>(h) + off, len);
      FLGP_CUDA(cudaMemcpyAsync(static_cast<char*>(d) + off, L.bounce[b], len, cudaMemcpyHostToDevice, L.st));
      FLGP_CUDA(cudaEventRecord(L.ev[b], L.st));
      L.used[b] = true;
    }
  });
  for (int t = 0; t < nl; ++t)
    for (int b = 0; b < 2; ++b)
      if (s.lane[t].used[b]) FLGP_CUDA(cudaStreamWaitEvent(st, s.lane[t].ev[b], 0));
}

void d2h_copy(void* h, const void* d, size_t bytes, cudaStream_t st) {
  if (bytes < HC_MIN || !staged_enabled() || !is_pageable(h)) {
    FLGP_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st));
    return;
  }
  int dev = 0;
  FLGP_CUDA(cudaGetDevice(&dev));
  Stager& s = g_stager[dev & 63];
  std::lock_guard<std::mutex> lk(s.mu);
  stager_init(s);
  FLGP_CUDA(cudaEventRecord(s.gate, st));  // the source is produced by work queued on st
  const size_t HC_CHUNK = hc_chunk();
  const size_t nchunks = (bytes + HC_CHUNK - 1) / HC_CHUNK;
  const int nl = (int)std::min<size_t>(s.nl, nchunks);
  run_lanes(dev, nl, [&](int t) {
    Lane& L = s.lane[t];
    FLGP_CUDA(cudaStreamWaitEvent(L.st, s.gate, 0));
    for (int b = 0; b < 2; ++b)
      if (L.used[b]) FLGP_CUDA(cudaEventSynchronize(L.ev[b]));  // an earlier upload may still read the buffer
    int b = 0;
    size_t prev_off = 0, prev_len = 0;
    bool have_prev = false;
    for (size_t k = t; k < nchunks; k += nl, b ^= 1) {
      const size_t off = k * HC_CHUNK, len = std::min(HC_CHUNK, bytes - off);
      FLGP_CUDA(cudaMemcpyAsync(L.bounce[b], static_cast<const char*>(d) + off, len, cudaMemcpyDeviceToHost, L.st));
      FLGP_CUDA(cudaEventRecord(L.ev[b], L.st));
      L.used[b] = true;
      if (have_prev) {  // drain the other buffer while this chunk is on the bus
        FLGP_CUDA(cudaEventSynchronize(L.ev[b ^ 1]));
        std::memcpy(static_cast<char*>(h) + prev_off, L.bounce[b ^ 1], prev_len);
      }
      prev_off = off;
      prev_len = len;
      have_prev = true;
    }
    if (have_prev) {
      FLGP_CUDA(cudaEventSynchronize(L.ev[b ^ 1]));
      std::memcpy(static_cast<char*>(h) + prev_off, L.bounce[b ^ 1], prev_len);
    }
  });
}

}  // namespace flgp
