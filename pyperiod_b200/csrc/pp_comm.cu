// pyperiod_b200 -- the one collective of the path: gather of the compact per-window results to one rank
// (SURVEY.md 8e), issued through NCCL on the compute stream right behind the last kernel.
//
// Windows are independent, so the data path has no exchange step; what crosses NVLink is ~124 B per window of
// periods / powers / status.  The library has no link-time dependency on NCCL: the symbols are resolved at run time
// from the libnccl.so.2 the process already has (the one torch.distributed loaded), so a build without NCCL still
// loads and every other entry point works.  ncclGather exists from NCCL 2.28; older libraries get the equivalent
// grouped ncclSend / ncclRecv.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>

#include "../../include/pyperiod_b200.h"
#include "pp_host.cuh"

namespace pp {

// the handful of NCCL declarations used here (stable since NCCL 2.0; nccl.h is not on every include path)
typedef struct ncclComm* nccl_comm_t;
typedef struct { char internal[128]; } nccl_unique_id;
constexpr int kNcclSuccess = 0;
constexpr int kNcclUint8 = 1;

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(nccl_unique_id*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*Gather)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

static NcclApi g_nccl;   // resolved once; read-only afterwards

static int nccl_load(const char* path) {
  if (g_nccl.handle) return 0;
  void* h = nullptr;
  if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // already in the process (torch)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(-4, "NCCL library not found (pass its path to pp_comm_load)%s");
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  a.Gather = reinterpret_cast<decltype(a.Gather)>(dlsym(h, "ncclGather"));   // NCCL >= 2.28
  a.Send = reinterpret_cast<decltype(a.Send)>(dlsym(h, "ncclSend"));
  a.Recv = reinterpret_cast<decltype(a.Recv)>(dlsym(h, "ncclRecv"));
  a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(h, "ncclGroupStart"));
  a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(h, "ncclGetVersion"));
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Send || !a.Recv || !a.GroupStart || !a.GroupEnd)
    return fail(-4, "NCCL library lacks a required symbol%s");
  g_nccl = a;
  return 0;
}

static int nccl_check(int rc, const char* what) {
  if (rc == kNcclSuccess) return 0;
  return fail(-5, "NCCL error: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : what);
}

struct Comm {
  nccl_comm_t comm;
  int nranks, rank;
};

}  // namespace pp

using namespace pp;

extern "C" {

int pp_comm_load(const char* library_path) { return nccl_load(library_path); }

int pp_comm_version(void) {
  if (nccl_load(nullptr)) return -1;
  int v = 0;
  if (!g_nccl.GetVersion || g_nccl.GetVersion(&v) != kNcclSuccess) return 0;
  return v;
}

int pp_comm_unique_id(void* id_out_128_bytes) {
  if (!id_out_128_bytes) return fail(-1, "null id buffer%s");
  if (int rc = nccl_load(nullptr)) return rc;
  nccl_unique_id id;
  if (int rc = nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId")) return rc;
  memcpy(id_out_128_bytes, &id, sizeof(id));
  return 0;
}

int pp_comm_init(void** comm_out, int32_t nranks, int32_t rank, const void* id_128_bytes) {
  if (!comm_out || !id_128_bytes || nranks < 1 || rank < 0 || rank >= nranks) return fail(-1, "bad communicator arguments%s");
  if (int rc = nccl_load(nullptr)) return rc;
  nccl_unique_id id;
  memcpy(&id, id_128_bytes, sizeof(id));
  Comm* c = new Comm{nullptr, nranks, rank};
  if (int rc = nccl_check(g_nccl.CommInitRank(&c->comm, nranks, id, rank), "ncclCommInitRank")) {
    delete c;
    return rc;
  }
  *comm_out = c;
  return 0;
}

int pp_comm_destroy(void* comm) {
  if (!comm) return 0;
  Comm* c = reinterpret_cast<Comm*>(comm);
  const int rc = g_nccl.CommDestroy ? nccl_check(g_nccl.CommDestroy(c->comm), "ncclCommDestroy") : 0;
  delete c;
  return rc;
}

int pp_gather(void* comm, const void* send, void* recv, size_t bytes, int32_t root, void* stream) {
  if (!comm || !send) return fail(-1, "null communicator or send buffer%s");
  Comm* c = reinterpret_cast<Comm*>(comm);
  if (root < 0 || root >= c->nranks) return fail(-1, "bad root%s");
  if (c->rank == root && !recv) return fail(-1, "root needs a receive buffer of nranks * bytes%s");
  if (bytes == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (g_nccl.Gather) return nccl_check(g_nccl.Gather(send, recv, bytes, kNcclUint8, root, c->comm, s), "ncclGather");
  // NCCL < 2.28: the same exchange as one group of point-to-point operations
  if (int rc = nccl_check(g_nccl.GroupStart(), "ncclGroupStart")) return rc;
  int rc = nccl_check(g_nccl.Send(send, bytes, kNcclUint8, root, c->comm, s), "ncclSend");
  if (rc == 0 && c->rank == root)
    for (int r = 0; r < c->nranks && rc == 0; ++r)
      rc = nccl_check(g_nccl.Recv(reinterpret_cast<char*>(recv) + (size_t)r * bytes, bytes, kNcclUint8, r, c->comm, s),
                      "ncclRecv");
  const int rc2 = nccl_check(g_nccl.GroupEnd(), "ncclGroupEnd");
  return rc ? rc : rc2;
}

}  // extern "C"
