// Exact-fp32 grouped GEMM for the actor/critic MLP layers (precision mode 0).
//
// Replaces the ATen/MKL nn.Linear forward calls of models.py:33-40,77-83 and their autograd
// backward (ddpg.py:230,242).  One launch runs up to 8 independent layer problems (the actor,
// critic and target networks advance in lock-step through the step's dependency levels), each
// tiled 32x32 so a 256x256x256 layer spreads over 64 CTAs: at batch 256 the whole step is
// latency-bound, so small tiles on many SMs beat big tiles on few.  Accumulation is plain FFMA
// in k order, i.e. a true fp32 dot product (needed for the 1e-5 parity of config 2).
#include "gemm_ffma.cuh"

namespace d4pg {

constexpr int BM = 32, BN = 32, KC = 64;
constexpr int LDS_A = BM + 4;   // multiple of 4 (float4 reads)
constexpr int LDS_B = BN + 4;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_WARPS = GEMM_THREADS / 32;
constexpr int KW = KC / GEMM_WARPS;                  // k values per warp per chunk (intra-CTA split-K)
constexpr int PER_THREAD = BM * KC / GEMM_THREADS;   // 8 staged elements per thread per operand per chunk

// Stage one K-chunk of both operands global -> registers.  All 16 loads of a thread are issued
// back to back (fully unrolled, predicated) so one chunk costs ONE L2 round trip, and the next
// chunk's loads are in flight while the current one is being multiplied.
__device__ __forceinline__ void load_chunk(const GemmProblem& P, int m0, int n0, int k0, int tid,
                                           float (&ra)[PER_THREAD], float (&rb)[PER_THREAD]) {
#pragma unroll
  for (int r = 0; r < PER_THREAD; ++r) {
    const int e = tid + r * GEMM_THREADS;
    float va = 0.f, vb = 0.f;
    if (P.mode == GEMM_DW) {                       // A(i,k) = dZ[k*lda + i]: i fastest
      const int kk = e / BM, i = e % BM, gi = m0 + i, gk = k0 + kk;
      if (gi < P.M && gk < P.K) va = __ldg(P.A + size_t(gk) * P.lda + gi);
    } else {                                       // A(i,k) = A[i*lda + k]: k fastest (+ concat)
      const int i = e / KC, kk = e % KC, gi = m0 + i, gk = k0 + kk;
      if (gi < P.M && gk < P.K)
        va = (gk < P.K1) ? __ldg(P.A + size_t(gi) * P.lda + gk) : __ldg(P.A2 + size_t(gi) * P.lda2 + (gk - P.K1));
    }
    if (P.mode == GEMM_FWD) {                      // B(k,j) = W[j*ldb + k]: k fastest
      const int j = e / KC, kk = e % KC, gj = n0 + j, gk = k0 + kk;
      if (gj < P.N && gk < P.K) vb = __ldg(P.Bm + size_t(gj) * P.ldb + gk);
    } else {                                       // B(k,j) = B[k*ldb + j]: j fastest
      const int kk = e / BN, j = e % BN, gj = n0 + j, gk = k0 + kk;
      if (gj < P.N && gk < P.K) vb = __ldg(P.Bm + size_t(gk) * P.ldb + gj);
    }
    ra[r] = va; rb[r] = vb;
  }
}

__device__ __forceinline__ void store_chunk(const GemmProblem& P, int tid, float* __restrict__ As, float* __restrict__ Bs,
                                            const float (&ra)[PER_THREAD], const float (&rb)[PER_THREAD]) {
#pragma unroll
  for (int r = 0; r < PER_THREAD; ++r) {
    const int e = tid + r * GEMM_THREADS;
    if (P.mode == GEMM_DW) As[(e / BM) * LDS_A + (e % BM)] = ra[r];
    else As[(e % KC) * LDS_A + (e / KC)] = ra[r];
    if (P.mode == GEMM_FWD) Bs[(e % KC) * LDS_B + (e / KC)] = rb[r];
    else Bs[(e / BN) * LDS_B + (e % BN)] = rb[r];
  }
}

// Per-CTA structure (ncu of the first version showed 2 warps/scheduler stalled on LDS->FFMA
// dependencies, IPC 0.38): the 8 warps now split K inside the CTA.  Each warp owns the whole
// 32x32 tile for 1/8 of every K-chunk with an 8x4 register tile per lane (32 independent FFMAs
// per 3 LDS.128), and the 8 partial tiles are summed through shared memory in fixed warp order
// (deterministic).  Per-warp serial instruction count drops ~3x, which is what bounds a
// latency-bound 256^3 layer at batch 256.
__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_ffma_kernel(const __grid_constant__ GemmBatch batch) {
  __shared__ __align__(16) float smem[2 * KC * LDS_A + 2 * KC * LDS_B];
  float* As0 = smem;                      // [2][KC*LDS_A]
  float* Bs0 = smem + 2 * KC * LDS_A;     // [2][KC*LDS_B]
  static_assert(2 * KC * LDS_A + 2 * KC * LDS_B >= GEMM_WARPS * BM * BN, "partial-tile buffer must fit");

  int pi = 0;
#pragma unroll
  for (int i = 1; i < GEMM_MAX_PROBLEMS; ++i)
    if (i < batch.n && int(blockIdx.x) >= batch.p[i].tile_begin) pi = i;
  const GemmProblem& P = batch.p[pi];
  const int tile = blockIdx.x - P.tile_begin;
  const int tm = tile / P.tiles_n, tn = tile - tm * P.tiles_n;
  const int m0 = tm * BM, n0 = tn * BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = (lane >> 3) * 8, c0 = (lane & 7) * 4;      // lane's 8x4 sub-tile

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float colsum = 0.f;                       // DW: bias gradient, threads < BM of the tn==0 tiles
  const bool want_bias_grad = (P.mode == GEMM_DW) && (P.bias_grad != nullptr) && (tn == 0);

  float ra[PER_THREAD], rb[PER_THREAD];
  const int nchunks = (P.K + KC - 1) / KC;
  load_chunk(P, m0, n0, 0, tid, ra, rb);
  store_chunk(P, tid, As0, Bs0, ra, rb);
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    const int cur = c & 1;
    if (c + 1 < nchunks) load_chunk(P, m0, n0, (c + 1) * KC, tid, ra, rb);     // in flight during the FMAs
    const float* __restrict__ as = As0 + cur * KC * LDS_A;
    const float* __restrict__ bs = Bs0 + cur * KC * LDS_B;
#pragma unroll
    for (int k = 0; k < KW; ++k) {                                              // zero-padded past K
      const int kk = warp * KW + k;
      const float4 a0 = *reinterpret_cast<const float4*>(&as[kk * LDS_A + r0]);
      const float4 a1 = *reinterpret_cast<const float4*>(&as[kk * LDS_A + r0 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&bs[kk * LDS_B + c0]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (want_bias_grad && tid < BM) {
#pragma unroll 8
      for (int kk = 0; kk < KC; ++kk) colsum += as[kk * LDS_A + tid];
    }
    if (c + 1 < nchunks) store_chunk(P, tid, As0 + (cur ^ 1) * KC * LDS_A, Bs0 + (cur ^ 1) * KC * LDS_B, ra, rb);
    __syncthreads();
  }

  // ---- cross-warp reduction of the 8 partial tiles (fixed order w = 0..7) ---------------------
  float* red = smem;                        // [GEMM_WARPS][BM][BN]
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(&red[(warp * BM + r0 + i) * BN + c0]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  __syncthreads();
  const int orow = tid >> 3, ocol = (tid & 7) * 4;          // thread's 4 outputs
  float4 sum = *reinterpret_cast<const float4*>(&red[orow * BN + ocol]);
#pragma unroll
  for (int w = 1; w < GEMM_WARPS; ++w) {
    const float4 t = *reinterpret_cast<const float4*>(&red[(w * BM + orow) * BN + ocol]);
    sum.x += t.x; sum.y += t.y; sum.z += t.z; sum.w += t.w;
  }

  // ---- epilogue ------------------------------------------------------------------------------
  const int gi = m0 + orow;
  if (gi < P.M) {
    const float v[4] = {sum.x, sum.y, sum.z, sum.w};
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int gj = n0 + ocol + cc;
      if (gj >= P.N) continue;
      float x = v[cc];
      switch (P.epi) {
        case EPI_BIAS: x += __ldg(P.bias + gj); break;
        case EPI_BIAS_RELU: x = fmaxf(x + __ldg(P.bias + gj), 0.f); break;
        case EPI_BIAS_TANH: x = tanhf(x + __ldg(P.bias + gj)); break;
        case EPI_RELU_MASK: x = (__ldg(P.aux + size_t(gi) * P.ldaux + gj) > 0.f) ? x : 0.f; break;
        case EPI_TANH_MASK: { const float t = __ldg(P.aux + size_t(gi) * P.ldaux + gj); x *= (1.f - t * t); } break;
        default: break;
      }
      P.C[size_t(gi) * P.ldc + gj] = x;
    }
  }
  if (want_bias_grad && tid < BM && m0 + tid < P.M) P.bias_grad[m0 + tid] = colsum;
}

// ---- host side -------------------------------------------------------------------------------
GemmProblem gemm_fwd(const float* X, int ldx, const float* X2, int ldx2, int K1, const float* W, int ldw,
                     const float* bias, float* Y, int ldy, int M, int N, int K, int epi) {
  GemmProblem p{};
  p.A = X; p.lda = ldx; p.A2 = X2 ? X2 : X; p.lda2 = X2 ? ldx2 : ldx; p.K1 = X2 ? K1 : K;
  p.Bm = W; p.ldb = ldw; p.bias = bias; p.C = Y; p.ldc = ldy; p.M = M; p.N = N; p.K = K;
  p.mode = GEMM_FWD; p.epi = epi;
  return p;
}
GemmProblem gemm_dx(const float* dZ, int lddz, const float* W, int ldw, float* dX, int lddx,
                    int M, int N_in, int K_out, int epi, const float* aux, int ldaux) {
  GemmProblem p{};
  p.A = dZ; p.lda = lddz; p.A2 = dZ; p.lda2 = lddz; p.K1 = K_out;
  p.Bm = W; p.ldb = ldw; p.C = dX; p.ldc = lddx; p.M = M; p.N = N_in; p.K = K_out;
  p.mode = GEMM_DX; p.epi = epi; p.aux = aux; p.ldaux = ldaux;
  return p;
}
GemmProblem gemm_dw(const float* dZ, int lddz, const float* X, int ldx, float* dW, int lddw,
                    float* db, int N_out, int K_in, int M_batch) {
  GemmProblem p{};
  p.A = dZ; p.lda = lddz; p.A2 = dZ; p.lda2 = lddz; p.K1 = M_batch;
  p.Bm = X; p.ldb = ldx; p.C = dW; p.ldc = lddw; p.bias_grad = db;
  p.M = N_out; p.N = K_in; p.K = M_batch;
  p.mode = GEMM_DW; p.epi = EPI_NONE;
  return p;
}
void gemm_batch_begin(GemmBatch& b) { b.n = 0; b.total_tiles = 0; }
void gemm_batch_add(GemmBatch& b, const GemmProblem& pin) {
  GemmProblem p = pin;
  p.tiles_m = cdiv(p.M, BM); p.tiles_n = cdiv(p.N, BN); p.tile_begin = b.total_tiles;
  b.total_tiles += p.tiles_m * p.tiles_n;
  b.p[b.n++] = p;
}
int gemm_batch_launch(const GemmBatch& b, cudaStream_t st) {
  D4PG_REQUIRE(b.n > 0 && b.n <= GEMM_MAX_PROBLEMS, D4PG_EINVAL, "gemm_batch_launch: %d problems", b.n);
  gemm_ffma_kernel<<<b.total_tiles, GEMM_THREADS, 0, st>>>(b);
  D4PG_LAUNCH_OK();
  return D4PG_OK;
}

}  // namespace d4pg
