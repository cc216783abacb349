// Standalone hardware probe for the tcgen05 building blocks (UMMA descriptors, SWIZZLE_128B K-major
// and MN-major operand layouts, TMEM accumulate + tcgen05.ld, 3xTF32 split).  Not part of the
// library: it exists so descriptor encodings are validated on a B200 against a CPU reference before
// the production kernel depends on them.   build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../d4pg-pytorch_b200/csrc/tc_common.cuh"

using namespace d4pg::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// D[128 x N] = A * B^T-ish.  A_MN=false: A[m*lda+k]; true: A[k*lda+m].  B_MN=false: B[n*ldb+k]; true: B[k*ldb+n].
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(128) probe_gemm(const float* A, const float* B, float* D, int N, int K,
                                                  int lda, int ldb, int passes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nchunk = K / 32;
  const uint32_t a_chunk_bytes = 128 * 128, b_chunk_bytes = uint32_t(N) * 128;
  uint8_t* Ahi = smem;
  uint8_t* Alo = Ahi + nchunk * a_chunk_bytes;
  uint8_t* Bhi = Alo + nchunk * a_chunk_bytes;
  uint8_t* Blo = Bhi + nchunk * b_chunk_bytes;

  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }

  // stage operands (generic proxy) into the canonical SWIZZLE_128B layouts, hi/lo split
  for (int e = tid; e < 128 * K; e += 128) {
    int m, k;
    if (A_MN) { k = e / 128; m = e % 128; } else { m = e / K; k = e % K; }
    float x = A_MN ? A[size_t(k) * lda + m] : A[size_t(m) * lda + k];
    int c = k / 32, kk = k % 32;
    uint32_t off = c * a_chunk_bytes + (A_MN ? uint32_t((m / 32) * 4096) + sw128b32_mnmajor_off(kk, m % 32) : sw128_kmajor_off(m, kk));
    float hi = tf32_hi(x);
    *reinterpret_cast<float*>(Ahi + off) = hi;
    *reinterpret_cast<float*>(Alo + off) = tf32_lo(x, hi);
  }
  for (int e = tid; e < N * K; e += 128) {
    int n, k;
    if (B_MN) { k = e / N; n = e % N; } else { n = e / K; k = e % K; }
    float x = B_MN ? B[size_t(k) * ldb + n] : B[size_t(n) * ldb + k];
    int c = k / 32, kk = k % 32;
    uint32_t off = c * b_chunk_bytes + (B_MN ? uint32_t((n / 32) * 4096) + sw128b32_mnmajor_off(kk, n % 32) : sw128_kmajor_off(n, kk));
    float hi = tf32_hi(x);
    *reinterpret_cast<float*>(Bhi + off) = hi;
    *reinterpret_cast<float*>(Blo + off) = tf32_lo(x, hi);
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = make_idesc(FMT_TF32, A_MN, B_MN, 128, uint32_t(N));
    bool acc = false;
    for (int p = 0; p < passes; ++p) {
      const uint8_t* Ap = (p == 2) ? Alo : Ahi;
      const uint8_t* Bp = (p == 1) ? Blo : Bhi;
      for (int c = 0; c < nchunk; ++c) {
        for (int ks = 0; ks < 4; ++ks) {
          uint64_t ad = A_MN ? make_smem_desc(smem_u32(Ap + c * a_chunk_bytes + ks * 1024), 4096, 512, 1)
                             : make_smem_desc(smem_u32(Ap + c * a_chunk_bytes + ks * 32), 16, 1024);
          uint64_t bd = B_MN ? make_smem_desc(smem_u32(Bp + c * b_chunk_bytes + ks * 1024), 4096, 512, 1)
                             : make_smem_desc(smem_u32(Bp + c * b_chunk_bytes + ks * 32), 16, 1024);
          mma_tf32(tmem_d, ad, bd, idesc, acc);
          acc = true;
        }
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float r[32];
    tmem_ld_32x32(tmem_d + (uint32_t(warp * 32) << 16) + uint32_t(c0), r);
    const int row = warp * 32 + (tid & 31);
    for (int j = 0; j < 32 && c0 + j < N; ++j) D[size_t(row) * N + c0 + j] = r[j];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 64);
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <bool A_MN, bool B_MN>
static int run_case(int N, int K, int passes) {
  const int M = 128;
  std::vector<float> A(size_t(M) * K), B(size_t(N) * K), D(size_t(M) * N, -777.f);
  srand(1234 + N * 7 + K);
  for (auto& x : A) x = (rand() / float(RAND_MAX) - 0.5f) * 2.f;
  for (auto& x : B) x = (rand() / float(RAND_MAX) - 0.5f) * 2.f;
  // logical A(m,k), B(n,k) stored according to majorness
  auto a_at = [&](int m, int k) { return A_MN ? A[size_t(k) * M + m] : A[size_t(m) * K + k]; };
  auto b_at = [&](int n, int k) { return B_MN ? B[size_t(k) * N + n] : B[size_t(n) * K + k]; };
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice));
  size_t smem = size_t(K / 32) * (128 * 128 + N * 128) * 2 + 2048;
  CK(cudaFuncSetAttribute(probe_gemm<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  probe_gemm<A_MN, B_MN><<<1, 128, smem>>>(dA, dB, dD, N, K, A_MN ? M : K, B_MN ? N : K, passes);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) {
        float a = a_at(m, k), b = b_at(n, k);
        if (passes == 1) { a = trunc_tf32(a); b = trunc_tf32(b); }
        ref += double(a) * double(b);
      }
      maxerr = fmax(maxerr, fabs(ref - double(D[size_t(m) * N + n])));
      maxref = fmax(maxref, fabs(ref));
    }
  const double tol = passes == 1 ? 2e-5 : 2e-5;
  const bool ok = maxerr <= tol * fmax(1.0, maxref);
  printf("probe A_%s B_%s N=%3d K=%3d passes=%d  max_abs_err=%.3e  max_ref=%.3f  %s\n", A_MN ? "MN" : "K ", B_MN ? "MN" : "K ",
         N, K, passes, maxerr, maxref, ok ? "OK" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return ok ? 0 : 1;
}

int main() {
  int fails = 0;
  fails += run_case<false, false>(16, 32, 1);
  fails += run_case<false, false>(64, 128, 1);
  fails += run_case<false, false>(64, 128, 3);
  fails += run_case<false, true>(32, 32, 1);
  fails += run_case<false, true>(64, 128, 3);
  fails += run_case<true, true>(32, 32, 1);
  fails += run_case<true, true>(64, 128, 3);
  fails += run_case<true, false>(64, 64, 1);
  printf("tc_probe: %d failing case(s)\n", fails);
  return fails ? 1 : 0;
}
