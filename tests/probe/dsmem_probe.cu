// Standalone hardware probe: cost of exchanging a 32x32 fp32 tile (4 KB) from every CTA of an
// 8-CTA cluster to all 8 CTAs (each receives a 32 KB plane), as the chain kernels do between layers.
//   mode 0: st.async.v4 (register source) with mbarrier complete_tx at the destination
//   mode 1: one cp.async.bulk shared::cta -> shared::cluster per peer (4 KB each), mbarrier complete_tx
//   mode 2: global-memory exchange: st.global + barrier.cluster release/acquire + cp.async (what v2 does)
//   mode 3: plain st.shared::cluster.v4 + barrier.cluster release/acquire
// Also times one mma.sync.m16n8k8 tf32 stream per warp (mode 4) to learn the legacy tensor rate on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_probe dsmem_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int CL = 8, ROUNDS = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
          smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(256, 1)
probe(int mode, float* gplanes, long long* cycles_out, float* check_out) {
  extern __shared__ __align__(128) float smem[];
  float* plane = smem;                 // [2][256][32] receive planes (ping-pong)
  float* stage = smem + 2 * 8192;      // [32][32]
  __shared__ __align__(8) uint64_t bars[ROUNDS];
  const int tid = threadIdx.x;
  const uint32_t rank = cta_rank();
  const int cluster_id = blockIdx.x / CL;
  if (tid == 0) {
    for (int i = 0; i < ROUNDS; ++i) { mbar_init(&bars[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < ROUNDS; ++i) mbar_expect_tx(&bars[i], 32768);
  }
  __syncthreads();
  cluster_sync_all();
  float4 v = make_float4(float(rank), float(tid), 1.f, 2.f);
  const int col = tid & 31, row4 = (tid >> 5) * 4;        // thread's 4 rows of column `col` of the tile
  long long t0 = clock64();
  for (int r = 0; r < ROUNDS; ++r) {
    float* dst_plane = plane + (r & 1) * 8192;
    const int off = (rank * 32 + col) * 32 + row4;        // plane[k = rank*32+col][row4..row4+3]
    if (mode == 0) {
      const uint32_t la = smem_u32(dst_plane + off), lb = smem_u32(&bars[r]);
#pragma unroll
      for (uint32_t p = 0; p < CL; ++p) {
        const uint32_t ra = mapa(la, p), rb = mapa(lb, p);
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(ra),
                     "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(rb)
                     : "memory");
      }
      mbar_wait(&bars[r], 0);
    } else if (mode == 1) {
      *reinterpret_cast<float4*>(stage + col * 32 + row4) = v;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid < CL) {
        const uint32_t ra = mapa(smem_u32(dst_plane + rank * 1024), tid), rb = mapa(smem_u32(&bars[r]), tid);
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ra),
                     "r"(smem_u32(stage)), "r"(4096), "r"(rb)
                     : "memory");
      }
      mbar_wait(&bars[r], 0);
      __syncthreads();            // stage may be rewritten next round (source reads are done once every peer got its copy... approximately)
    } else if (mode == 2) {
      float* gp = gplanes + size_t(cluster_id) * 2 * 8192 + (r & 1) * 8192;
      *reinterpret_cast<float4*>(gp + off) = v;
      cluster_sync_all();
      for (int e = tid; e < 2048; e += 256) {
        const uint32_t d = smem_u32(dst_plane + e * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gp + e * 4) : "memory");
      }
      asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    } else if (mode == 3) {
      const uint32_t la = smem_u32(dst_plane + off);
#pragma unroll
      for (uint32_t p = 0; p < CL; ++p) {
        const uint32_t ra = mapa(la, p);
        asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      }
      cluster_sync_all();
    }
    v.x += dst_plane[(tid * 37) & 8191];                  // consume something so rounds depend on each other
  }
  long long t1 = clock64();
  if (tid == 0) cycles_out[blockIdx.x] = (t1 - t0) / ROUNDS;
  if (tid == 0) check_out[blockIdx.x] = plane[(ROUNDS - 1 & 1) * 8192 + (7 * 32 + 5) * 32 + 9] + v.x * 0.f;
  cluster_sync_all();
}

// legacy tensor-core rate: each warp issues `n` dependent-free mma.sync m16n8k8 tf32 (8 accumulator tiles)
__global__ void __launch_bounds__(256, 1) mma_rate(int n, long long* cycles_out, float* sink) {
  const int tid = threadIdx.x;
  float acc[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  uint32_t a[4] = {0x3f800000u + tid, 0x3f900000u, 0x3fa00000u, 0x3fb00000u}, b[2] = {0x3f800000u, 0x3fc00000u + tid};
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
  sink[blockIdx.x * 256 + tid] = s;
  if (tid == 0) cycles_out[blockIdx.x] = t1 - t0;
}

// the same stream with the operand traffic of a 3xTF32 tile step: per 8-deep k-step a warp loads 8 A + 8 B
// fragment registers from shared memory, splits them into hi/lo (cvt.rna.tf32 + sub) and issues 24 mma
__global__ void __launch_bounds__(256, 1) mma_tile_rate(int n, long long* cycles_out, float* sink) {
  __shared__ float As[64 * 40], Bs[64 * 40];
  const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  for (int i = tid; i < 64 * 40; i += 256) { As[i] = 1.f + i * 1e-3f; Bs[i] = 0.5f + i * 1e-3f; }
  __syncthreads();
  float acc[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
    const int k0 = (it & 7) * 8;
    uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float x = As[(k0 + t4 + (r >> 1) * 4) * 40 + m * 16 + g + (r & 1) * 8];
        uint32_t h; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
        const float lo = x - __uint_as_float(h);
        uint32_t l; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(lo));
        ah[m][r] = h; al[m][r] = l;
      }
#pragma unroll
    for (int nn = 0; nn < 4; ++nn)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float x = Bs[(k0 + t4 + r * 4) * 40 + nn * 8 + g];
        uint32_t h; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
        const float lo = x - __uint_as_float(h);
        uint32_t l; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(lo));
        bh[nn][r] = h; bl[nn][r] = l;
      }
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) {
        float* c = acc[m * 4 + nn];
#define MMA(A, B) asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(B[0]), "r"(B[1]))
        MMA(al[m], bh[nn]); MMA(ah[m], bl[nn]); MMA(ah[m], bh[nn]);
      }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
  sink[blockIdx.x * 256 + tid] = s;
  if (tid == 0) cycles_out[blockIdx.x] = t1 - t0;
}

int main() {
  const int nclusters = 16, grid = nclusters * CL;
  long long* cyc; float* chk; float* gplanes; float* sink;
  CK(cudaMalloc(&cyc, 256 * sizeof(long long))); CK(cudaMalloc(&chk, 256 * sizeof(float)));
  CK(cudaMalloc(&gplanes, size_t(nclusters) * 2 * 8192 * sizeof(float)));
  CK(cudaMalloc(&sink, 148 * 256 * sizeof(float)));
  const size_t smem = (2 * 8192 + 1024) * sizeof(float);
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const char* names[] = {"st.async.v4 + mbarrier", "cp.async.bulk smem->dsmem + mbarrier", "global + cluster barrier + cp.async",
                         "st.shared::cluster.v4 + cluster barrier"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      probe<<<grid, 256, smem>>>(mode, gplanes, cyc, chk);
      CK(cudaDeviceSynchronize());
    }
    long long h[grid]; float c[grid];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(c, chk, sizeof(c), cudaMemcpyDeviceToHost));
    long long mx = 0, mn = 1ll << 60;
    for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
    printf("mode %d  %-42s  cycles/round min %lld max %lld   check %.1f\n", mode, names[mode], mn, mx, c[0]);
  }
  for (int n : {64, 256}) {
    mma_rate<<<148, 256>>>(n, cyc, sink);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long h[148];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    // 8 warps x n x 8 mma per CTA on one SM
    printf("mma.sync m16n8k8 tf32: n=%d  %lld cycles for %d mma/SM -> %.2f cycles per mma per SM sub-partition, %.0f MAC/clk/SM\n", n,
           h[0], 8 * n * 8, double(h[0]) / (2.0 * n * 8), double(8 * n * 8) * 1024.0 / double(h[0]));
  }
  for (int n : {32, 128}) {
    mma_tile_rate<<<148, 256>>>(n, cyc, sink);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long h[148];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("3xTF32 tile step (8 warps x n=%d k-steps: 16 LDS + 48 cvt/sub + 24 mma each): %lld cycles -> %.0f cycles per k-step per CTA; a 32x32x256 tile = 32 k-steps over 8 warps = 4 per warp -> %.0f cycles\n",
           n, h[0], double(h[0]) / n, double(h[0]) / n * 4);
  }
  return 0;
}
