// tcgen05.mma issue / dependency-latency probe (sm_100a).  One CTA, one issuing thread: NMMA kind::tf32 MMAs of shape
// M x N x 8 (SS mode, K-major SWIZZLE_128B operands in shared memory) accumulating round-robin into NACC independent
// TMEM accumulators; clock64 from first issue to tcgen05.commit completion.  Answers: is a K-loop of tiny MMAs bound by
// the accumulate dependency (time ~ 1/NACC) or by instruction issue (time independent of NACC)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I d4pg-pytorch_b200/csrc tests/probe/mma_probe.cu -o /tmp/mma_probe
#include "tc_common.cuh"
#include <stdio.h>
using namespace d4pg::tc;

template <int M, int N>
__global__ void probe(int nmma, int nacc, long long* out) {
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
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(FMT_TF32, false, false, M, N);
    const uint64_t tmpl = make_smem_desc(0, 16, 1024, 2);
    const uint32_t a0 = smem_u32(smem) >> 4, b0 = smem_u32(smem + M * 128 * 4) >> 4;     // 4 chunks of A, then 4 of B
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < nmma; ++i) {
        const int ks = i & 3, ch = (i >> 2) & 3;
        mma_tf32(tmem + uint32_t((i % nacc) * N), tmpl + (a0 + ch * (M * 128 >> 4) + 2 * ks), tmpl + (b0 + ch * (N * 128 >> 4) + 2 * ks), idesc, i >= nacc);
      }
      const long long t1 = clock64();
      mma_commit(&bar);
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      if (rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int M, int N>
void run(int nmma, int nacc) {
  long long* d; long long h[2];
  cudaMalloc(&d, 16);
  const int smem = (M + N) * 128 * 4 + 1024;
  cudaFuncSetAttribute(probe<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<M, N><<<1, 128, smem>>>(nmma, nacc, d);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("M=%3d N=%3d nmma=%3d nacc=%2d : issue %6lld cyc (%5.1f/mma)  complete %6lld cyc (%5.1f/mma)  floor %d/mma  %s\n", M, N, nmma, nacc,
         h[0], double(h[0]) / nmma, h[1], double(h[1]) / nmma, (M > 128 ? M : 128) * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int nacc : {1, 2, 4, 8, 12}) run<64, 32>(96, nacc);
  for (int nacc : {1, 2, 4, 8}) run<128, 32>(96, nacc);
  for (int nacc : {1, 2, 4, 8}) run<64, 64>(96, nacc);
  for (int nacc : {1, 2, 4}) run<128, 64>(96, nacc);
  for (int nacc : {1, 2}) run<128, 128>(96, nacc);
  for (int nacc : {1, 2}) run<128, 256>(48, nacc);
  return 0;
}
