// comm.cu — the only collectives the path needs, over NCCL/NVLink.
//
// One process per GPU.  NCCL is bound with dlopen so that the single-GPU drop-in has no NCCL
// dependency and so that, under torchrun, the copy torch already loaded is the one used.
// Everything reduced on the data path is either int64 fixed-point limbs (associative => the
// result is bit-identical for every rank count) or small fp64 K x K blocks.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "common.cuh"

namespace flgp {

struct Nccl {
  void* handle = nullptr;
  ncclComm_t comm = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static Nccl* nccl_open() {
  Nccl* n = new Nccl;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    n->handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n->handle) break;
  }
  if (!n->handle) {
    delete n;
    fail(4, "multi-GPU requested but libnccl.so.2 cannot be loaded: %s", dlerror());
  }
#define BIND(field, sym)                                             \
  *(void**)(&n->field) = dlsym(n->handle, sym);                      \
  if (!n->field) {                                                   \
    delete n;                                                        \
    fail(4, "NCCL symbol %s missing", sym);                          \
  }
  BIND(GetUniqueId, "ncclGetUniqueId");
  BIND(CommInitRank, "ncclCommInitRank");
  BIND(AllReduce, "ncclAllReduce");
  BIND(AllGather, "ncclAllGather");
  BIND(CommDestroy, "ncclCommDestroy");
  BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
  return n;
}

#define FLGP_NCCL(n, call)                                                             \
  do {                                                                                 \
    ncclResult_t r_ = (call);                                                          \
    if (r_ != ncclSuccess) fail(4, "NCCL error at %s:%d: %s", __FILE__, __LINE__,      \
                                (n)->GetErrorString(r_));                              \
  } while (0)

void comm_unique_id(void* out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  Nccl* n = nccl_open();
  ncclUniqueId id;
  ncclResult_t r = n->GetUniqueId(&id);
  if (r != ncclSuccess) {
    std::string m = n->GetErrorString(r);
    delete n;
    fail(4, "ncclGetUniqueId: %s", m.c_str());
  }
  std::memcpy(out128, &id, 128);
  delete n;  // the library handle stays loaded (refcounted by dlopen)
}

void comm_init(Ctx* c, const void* id128, int rank, int nranks) {
  if (nranks < 1 || rank < 0 || rank >= nranks) fail(2, "bad rank %d / nranks %d", rank, nranks);
  comm_destroy(c);
  c->rank = rank;
  c->nranks = nranks;
  if (nranks == 1) return;
  Nccl* n = nccl_open();
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  FLGP_CUDA(cudaSetDevice(c->device));
  FLGP_NCCL(n, n->CommInitRank(&n->comm, nranks, id, rank));
  c->nccl = n;
}

void comm_destroy(Ctx* c) {
  if (c->nccl) {
    if (c->nccl->comm) c->nccl->CommDestroy(c->nccl->comm);
    delete c->nccl;
    c->nccl = nullptr;
  }
  c->rank = 0;
  c->nranks = 1;
}

void comm_allreduce_i64(Ctx* c, int64_t* d, size_t count) {
  if (c->nranks == 1 || count == 0) return;
  FLGP_NCCL(c->nccl, c->nccl->AllReduce(d, d, count, ncclInt64, ncclSum, c->nccl->comm, c->stream));
}
// out of place: dst = sum over ranks of src (src untouched)
void comm_allreduce_i64_to(Ctx* c, const int64_t* src, int64_t* dst, size_t count) {
  if (count == 0) return;
  if (c->nranks == 1) {
    if (dst != src) FLGP_CUDA(cudaMemcpyAsync(dst, src, sizeof(int64_t) * count, cudaMemcpyDeviceToDevice, c->stream));
    return;
  }
  FLGP_NCCL(c->nccl, c->nccl->AllReduce(src, dst, count, ncclInt64, ncclSum, c->nccl->comm, c->stream));
}
void comm_allreduce_f64(Ctx* c, double* d, size_t count) {
  if (c->nranks == 1 || count == 0) return;
  FLGP_NCCL(c->nccl, c->nccl->AllReduce(d, d, count, ncclFloat64, ncclSum, c->nccl->comm, c->stream));
}
// recv[q * count .. (q+1) * count) = send of rank q (bit copies: the only fp64 data-path exchange besides the K x K tail)
void comm_allgather_f64(Ctx* c, const double* send, double* recv, size_t count) {
  if (count == 0) return;
  if (c->nranks == 1) {
    if (recv != send) FLGP_CUDA(cudaMemcpyAsync(recv, send, sizeof(double) * count, cudaMemcpyDeviceToDevice, c->stream));
    return;
  }
  FLGP_NCCL(c->nccl, c->nccl->AllGather(send, recv, count, ncclFloat64, c->nccl->comm, c->stream));
}
void comm_allreduce_max_f64(Ctx* c, double* d, size_t count) {
  if (c->nranks == 1 || count == 0) return;
  FLGP_NCCL(c->nccl, c->nccl->AllReduce(d, d, count, ncclFloat64, ncclMax, c->nccl->comm, c->stream));
}

}  // namespace flgp
