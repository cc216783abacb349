// Data-parallel communicator: one flat-buffer all-reduce of the gradients per step over
// NVLink 5 / NVSwitch.  The reference has no collective at all (its multi-worker mode is
// Hogwild over shared CPU memory: main.py:394-405, ddpg.py:104-108, shared_adam.py:16-17);
// this is the synchronous-DP equivalent described in SURVEY.md section 8e.
// NCCL is bound at run time (dlopen) so the .so has no link-time dependency on it.
#include "internal.cuh"
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <new>

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclSum = 0 };

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.ok) return D4PG_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (lib) break; }   // torch's copy, if loaded
  const char* env = getenv("D4PG_NCCL_LIB");
  if (!lib && env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  for (const char* n : names) { if (lib) break; lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); }
  if (!lib) { d4pg::set_error("NCCL not loadable: %s (set D4PG_NCCL_LIB)", dlerror()); return D4PG_ENCCL; }
  g_nccl.lib = lib;
  *(void**)(&g_nccl.GetUniqueId) = dlsym(lib, "ncclGetUniqueId");
  *(void**)(&g_nccl.CommInitRank) = dlsym(lib, "ncclCommInitRank");
  *(void**)(&g_nccl.CommDestroy) = dlsym(lib, "ncclCommDestroy");
  *(void**)(&g_nccl.AllReduce) = dlsym(lib, "ncclAllReduce");
  *(void**)(&g_nccl.GetErrorString) = dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
    d4pg::set_error("NCCL symbols missing in loaded library");
    return D4PG_ENCCL;
  }
  g_nccl.ok = true;
  return D4PG_OK;
}
#define NCCL_OK(expr)                                                                          \
  do {                                                                                         \
    ncclResult_t _r = (expr);                                                                  \
    if (_r != 0) {                                                                             \
      d4pg::set_error("%s -> NCCL error %d (%s)", #expr, _r,                                   \
                      g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?");                \
      return D4PG_ENCCL;                                                                       \
    }                                                                                          \
  } while (0)
}  // namespace

struct d4pg_comm { ncclComm_t comm; int rank, world; };

namespace d4pg {
int comm_allreduce(d4pg_comm* c, float* buf, int64_t n, cudaStream_t st) {
  D4PG_REQUIRE(c && buf && n > 0, D4PG_EINVAL, "comm_allreduce: bad arguments");
  NCCL_OK(g_nccl.AllReduce(buf, buf, size_t(n), ncclFloat32, ncclSum, c->comm, st));
  return D4PG_OK;
}
}  // namespace d4pg

extern "C" int32_t d4pg_comm_unique_id(uint8_t* id128) {
  D4PG_REQUIRE(id128, D4PG_EINVAL, "d4pg_comm_unique_id: null argument");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  NCCL_OK(g_nccl.GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return D4PG_OK;
}

extern "C" int32_t d4pg_comm_create(const uint8_t* id128, int32_t rank, int32_t world, d4pg_comm_t** out) {
  D4PG_REQUIRE(id128 && out && world >= 1 && rank >= 0 && rank < world, D4PG_EINVAL, "d4pg_comm_create: bad arguments");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  d4pg_comm* c = new (std::nothrow) d4pg_comm();
  D4PG_REQUIRE(c, D4PG_EINVAL, "d4pg_comm_create: out of host memory");
  c->rank = rank; c->world = world;
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (r != 0) {
    d4pg::set_error("ncclCommInitRank failed: %d (%s)", r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    delete c; return D4PG_ENCCL;
  }
  *out = c;
  return D4PG_OK;
}

extern "C" int32_t d4pg_comm_destroy(d4pg_comm_t* c) {
  if (!c) return D4PG_OK;
  if (g_nccl.ok && c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
  return D4PG_OK;
}

extern "C" int32_t d4pg_comm_allreduce_sum(d4pg_comm_t* c, float* buf, int64_t n, d4pg_stream_t stream) {
  return d4pg::comm_allreduce(c, buf, n, d4pg::as_stream(stream));
}
