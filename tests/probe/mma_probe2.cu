// tcgen05.mma issue-rate probe, second pass: is the ~150 cycles per MMA of mma_probe.cu the instruction itself or the
// code around it?  Variants: V0 = one thread in a divergent branch, per-iteration descriptor arithmetic (as mma_probe.cu);
// V1 = whole warp runs the loop (warp-uniform values -> uniform registers), instruction predicated by elect.sync,
// fully unrolled body of NACC MMAs with compile-time accumulator / descriptor offsets.
#include "tc_common.cuh"
#include <stdio.h>
using namespace d4pg::tc;

__device__ __forceinline__ void mma_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}

template <int M, int N, int NACC, int V>
__global__ void probe(int nmma, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < (M + N) * 128 * 4 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 63);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tslot, 512);
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tslot;
  const uint32_t idesc = make_idesc(FMT_TF32, false, false, M, N);
  const uint64_t tmpl = make_smem_desc(0, 16, 1024, 2);
  const uint32_t a0 = smem_u32(smem) >> 4, b0 = smem_u32(smem + M * 128 * 4) >> 4;
  if (threadIdx.x < 32) {
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      if (V == 0) {
        if (threadIdx.x == 0)
          for (int i = 0; i < nmma; i += NACC) {
#pragma unroll
            for (int a = 0; a < NACC; ++a)
              mma_tf32(tmem + uint32_t(a * N), tmpl + (a0 + 2 * (a & 3)), tmpl + (b0 + 2 * (a & 3)), idesc, i > 0);
          }
      } else {
        for (int i = 0; i < nmma; i += NACC) {
#pragma unroll
          for (int a = 0; a < NACC; ++a)
            mma_elect(tmem + uint32_t(a * N), tmpl + (a0 + 2 * (a & 3)), tmpl + (b0 + 2 * (a & 3)), idesc, uint32_t(i > 0));
        }
      }
      const long long t1 = clock64();
      if (V == 0) { if (threadIdx.x == 0) mma_commit(&bar); } else commit_elect(&bar);
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      if (rep == 2 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
      __syncwarp();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int M, int N, int NACC, int V>
void run(int nmma) {
  long long* d; long long h[2];
  cudaMalloc(&d, 16);
  const int smem = (M + N) * 128 * 4 + 1024;
  cudaFuncSetAttribute(probe<M, N, NACC, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<M, N, NACC, V><<<1, 128, smem>>>(nmma, d);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("V%d M=%3d N=%3d nmma=%3d nacc=%2d : issue %6lld cyc (%5.1f/mma)  complete %6lld cyc (%5.1f/mma)  floor %d/mma  %s\n", V, M, N, nmma, NACC,
         h[0], double(h[0]) / nmma, h[1], double(h[1]) / nmma, (M > 128 ? M : 128) * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, 32, 1, 0>(96); run<64, 32, 4, 0>(96); run<64, 32, 8, 0>(96);
  run<64, 32, 1, 1>(96); run<64, 32, 4, 1>(96); run<64, 32, 8, 1>(96);
  run<128, 32, 8, 1>(96); run<128, 64, 8, 1>(96); run<128, 128, 4, 1>(96); run<128, 256, 2, 1>(96);
  run<64, 64, 8, 1>(96); run<64, 128, 4, 1>(96); run<64, 256, 2, 1>(96);
  return 0;
}
