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
#include <algorithm>

using d4pg::D4PG_MAX_PEERS;
using d4pg::PeerInfo;

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

struct d4pg_comm {
  ncclComm_t comm; int rank, world;
  // ---- fused all-reduce over peer memory (d4pg_comm_peer_*) ----------------------------------------------------
  // One cudaMalloc block per rank, exported with CUDA IPC: [2][n] gradient halves (double buffer), [n] reduced
  // gradient, two flag lines (256 B apart).
  float* xbuf; int64_t xn; unsigned long long* flags;       // local block
  void* peer_base[D4PG_MAX_PEERS];                          // opened IPC mappings (nullptr for self)
  float* peer_x[D4PG_MAX_PEERS]; unsigned long long* peer_flag[D4PG_MAX_PEERS];
  bool peer_ready;
};

namespace d4pg {
// "my gradient half of this step is complete" + "wait until every peer's is" as a launch of its own (the level plan has
// several dW launches): one CTA publishes, then polls its local inbox.
__global__ void peer_barrier_kernel(PeerSignal sig) {
  if (threadIdx.x == 0) peer_signal_last_cta(sig, 1u);
  __syncthreads();
  peer_wait_all(sig.local, sig.world);
  __threadfence_system();
}

PeerSignal comm_peer_signal(const PeerInfo& info, int kind) {
  PeerSignal s{};
  s.world = info.world;
  s.local = info.flag[info.rank] + kind * 64;
  for (int p = 0; p < info.world; ++p) s.inbox[p] = info.flag[p] + kind * 64 + 8 + info.rank;
  return s;
}

bool comm_peer_info(d4pg_comm* c, PeerInfo* out) {
  if (!c || !c->peer_ready) return false;
  if (!out) return true;
  out->world = c->world; out->rank = c->rank; out->n = c->xn;
  for (int r = 0; r < c->world; ++r) {
    out->x[r] = c->peer_x[r]; out->red[r] = c->peer_x[r] + 2 * c->xn;
    out->flag[r] = c->peer_flag[r]; out->flag2[r] = c->peer_flag[r] + 64;
  }
  return true;
}
int comm_peer_barrier(d4pg_comm* c, cudaStream_t st) {
  PeerInfo info{};
  D4PG_REQUIRE(comm_peer_info(c, &info), D4PG_ESTATE, "comm_peer_barrier: peers are not open");
  peer_barrier_kernel<<<1, 32, 0, st>>>(comm_peer_signal(info, 0));
  D4PG_LAUNCH_OK();
  return D4PG_OK;
}

// ---- reduce-scatter + all-gather over peer memory ---------------------------------------------------------------
// Rank r owns slice r of the flat gradient.  After every rank's dW published flag1 for this step, rank r reads slice r of
// all N halves (N-1 of them over NVLink), sums them IN RANK ORDER and stores the result into slice r of every rank's
// reduced buffer (N-1 remote stores).  Every element is reduced by exactly one rank, so all replicas consume the same
// bits.  Per rank: (N-1)/N x 1.15 MB in and out, instead of pulling all N halves (N x 1.15 MB in).  The last CTA
// publishes flag2; the fused Adam kernel waits for all ranks' flag2 and then streams its LOCAL reduced buffer.
struct PeerRSArgs {
  int world, rank; int64_t n4, lo4, hi4, half_off;
  const float* g[D4PG_MAX_PEERS]; float* red[D4PG_MAX_PEERS];
  const unsigned long long* my_f1;              // this rank's flag block of signal 0 (local inbox: every rank's dW is done)
  d4pg::PeerSignal sig2;                        // signal 1: this rank's slice is reduced and pushed
};
__global__ void __launch_bounds__(256) peer_reduce_scatter_kernel(const PeerRSArgs a) {
  d4pg::peer_wait_all(a.my_f1, a.world);
  for (int64_t i = a.lo4 + int64_t(blockIdx.x) * 256 + threadIdx.x; i < a.hi4; i += int64_t(gridDim.x) * 256) {
    float4 t[D4PG_MAX_PEERS];
#pragma unroll
    for (int r = 0; r < D4PG_MAX_PEERS; ++r)
      if (r < a.world) t[r] = __ldcg(reinterpret_cast<const float4*>(a.g[r] + a.half_off) + i);     // all loads in flight
    float4 s = t[0];
#pragma unroll
    for (int r = 1; r < D4PG_MAX_PEERS; ++r)
      if (r < a.world) { s.x += t[r].x; s.y += t[r].y; s.z += t[r].z; s.w += t[r].w; }
#pragma unroll
    for (int r = 0; r < D4PG_MAX_PEERS; ++r)
      if (r < a.world) reinterpret_cast<float4*>(a.red[r])[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence_system(); d4pg::peer_signal_last_cta(a.sig2, gridDim.x); }
}
int comm_peer_reduce_scatter(d4pg_comm* c, int parity, cudaStream_t st) {
  PeerInfo info{};
  D4PG_REQUIRE(comm_peer_info(c, &info), D4PG_ESTATE, "comm_peer_reduce_scatter: peers are not open");
  PeerRSArgs a{};
  a.world = info.world; a.rank = info.rank; a.n4 = info.n >> 2;
  const int64_t per = (a.n4 + info.world - 1) / info.world;
  a.lo4 = std::min<int64_t>(a.n4, per * info.rank); a.hi4 = std::min<int64_t>(a.n4, a.lo4 + per);
  a.half_off = int64_t(parity) * info.n;
  for (int r = 0; r < info.world; ++r) { a.g[r] = info.x[r]; a.red[r] = info.red[r]; }
  a.my_f1 = info.flag[info.rank]; a.sig2 = comm_peer_signal(info, 1);
  const int blocks = int(std::max<int64_t>(1, std::min<int64_t>(296, (a.hi4 - a.lo4 + 255) / 256)));
  peer_reduce_scatter_kernel<<<blocks, 256, 0, st>>>(a);
  D4PG_LAUNCH_OK();
  return D4PG_OK;
}

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
  c->xbuf = nullptr; c->xn = 0; c->flags = nullptr; c->peer_ready = false;
  for (int i = 0; i < D4PG_MAX_PEERS; ++i) { c->peer_base[i] = nullptr; c->peer_x[i] = nullptr; c->peer_flag[i] = nullptr; }
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
  for (int i = 0; i < D4PG_MAX_PEERS; ++i) if (c->peer_base[i]) cudaIpcCloseMemHandle(c->peer_base[i]);
  if (c->xbuf) cudaFree(c->xbuf);
  delete c;
  return D4PG_OK;
}

extern "C" int32_t d4pg_comm_allreduce_sum(d4pg_comm_t* c, float* buf, int64_t n, d4pg_stream_t stream) {
  return d4pg::comm_allreduce(c, buf, n, d4pg::as_stream(stream));
}

// ---- fused all-reduce over peer memory -------------------------------------------------------------------------
extern "C" int32_t d4pg_comm_peer_alloc(d4pg_comm_t* c, int64_t n_floats, uint8_t* handle64) {
  D4PG_REQUIRE(c && handle64 && n_floats > 0, D4PG_EINVAL, "d4pg_comm_peer_alloc: bad arguments");
  D4PG_REQUIRE(c->world <= D4PG_MAX_PEERS, D4PG_ENOTSUP, "d4pg_comm_peer_alloc: at most %d ranks (one node)", D4PG_MAX_PEERS);
  D4PG_REQUIRE(!c->xbuf, D4PG_ESTATE, "d4pg_comm_peer_alloc: already allocated");
  const int64_t n = (n_floats + 31) & ~int64_t(31);
  const size_t bytes = size_t(3 * n) * sizeof(float) + 1024;  // [2][n] halves + [n] reduced + two flag blocks of 512 B
  D4PG_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&c->xbuf), bytes));
  D4PG_CUDA_OK(cudaMemset(c->xbuf, 0, bytes));
  D4PG_CUDA_OK(cudaDeviceSynchronize());
  c->xn = n; c->flags = reinterpret_cast<unsigned long long*>(c->xbuf + 3 * n);
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
  D4PG_CUDA_OK(cudaIpcGetMemHandle(&h, c->xbuf));
  memcpy(handle64, &h, 64);
  return D4PG_OK;
}

extern "C" int32_t d4pg_comm_peer_open(d4pg_comm_t* c, const uint8_t* all_handles) {
  D4PG_REQUIRE(c && all_handles && c->xbuf, D4PG_ESTATE, "d4pg_comm_peer_open: call d4pg_comm_peer_alloc first");
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) { c->peer_x[r] = c->xbuf; c->peer_flag[r] = c->flags; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, all_handles + size_t(r) * 64, 64);
    void* base = nullptr;
    D4PG_CUDA_OK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_base[r] = base;
    c->peer_x[r] = static_cast<float*>(base);
    c->peer_flag[r] = reinterpret_cast<unsigned long long*>(static_cast<float*>(base) + 3 * c->xn);
  }
  c->peer_ready = true;
  return D4PG_OK;
}
extern "C" int32_t d4pg_comm_peer_ready(const d4pg_comm_t* c) { return (c && c->peer_ready) ? 1 : 0; }
/* some rank could not map its peers: every rank drops back to the NCCL all-reduce */
extern "C" int32_t d4pg_comm_peer_disable(d4pg_comm_t* c) {
  D4PG_REQUIRE(c, D4PG_EINVAL, "d4pg_comm_peer_disable: null handle");
  c->peer_ready = false;
  for (int i = 0; i < D4PG_MAX_PEERS; ++i) {
    if (c->peer_base[i]) { cudaIpcCloseMemHandle(c->peer_base[i]); c->peer_base[i] = nullptr; }
    c->peer_x[i] = nullptr; c->peer_flag[i] = nullptr;
  }
  return D4PG_OK;
}
