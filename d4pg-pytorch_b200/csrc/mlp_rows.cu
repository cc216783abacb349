// Row-owner MLP chains: a CTA carries R batch rows through EVERY layer of a network chain.
//
// Why: at batch 256 the actor/critic passes of DDPG.train (models.py:32-41,76-88; backward ddpg.py:230,242)
// are 14 dependent 256-wide layers.  Tiling each layer over many CTAs (gemm_ffma.cu: one launch per
// level; mlp_chain.cu: one cluster per 32 rows) needs an all-gather of the activations between layers,
// and on B200 that exchange costs ~1.4 us per layer however it is done (tests/probe/dsmem_probe.cu:
// global + cluster barrier 2.6-3.1k cycles, cp.async.bulk to DSMEM 2.6-3.4k, st.async 4.1-6.8k).
// Rows of the batch never mix in an MLP, so here nothing is exchanged at all: a CTA owns R rows
// (5 at batch 256 -> 52 CTAs per chain), keeps their activations in shared memory from the first layer
// to the last, and streams every weight matrix of the chain through a cp.async ring (32 contraction
// rows per stage, 5 stages in flight) -- weights do not depend on activations, so the prefetch runs
// ahead across layer boundaries and the only synchronisation is one __syncthreads per stage.
// Thread t owns output column t of the current layer: an fp32 dot product accumulated in k order
// (exact FFMA, no split, no cross-thread reduction); the R activations of a k are warp-broadcast reads.
//   per stage and thread: 8 x (1 LDS.128 weights + R LDS.128 broadcast) + 32 R FFMA
// Every CTA re-reading the same weights from L2 saturates it (136 CTAs x 70 KB per stage: 1.9 us/stage measured),
// so CTAs run in clusters of 8 that share each fetch: a producer warp per CTA issues 1/8 of every stage as
// cp.async.bulk ... .multicast::cluster copies that land in all 8 CTAs, completion on per-slot mbarriers
// (full: transaction bytes; empty: one arrival per consumer warp of every CTA of the cluster).
// Shared-memory traffic is 4+R wavefronts per 4R warp-FFMAs (weights are read once per CTA), the L2->SM
// stream is one pass over the chain's parameters per CTA (~1.1 MB for actor+critic).
#include "gemm_ffma_dev.cuh"
#include "mlp_rows.cuh"
#include <algorithm>

namespace d4pg {

constexpr int ROWS_WP = ROWS_KC + 4;                         // FWD stage pitch: W[j][k0..k0+31]; 36 = 4 x odd: conflict-free LDS.128
constexpr int ROWS_STAGE_FLOATS = ROWS_THREADS * ROWS_WP;    // 36.9 KB (DX stages need <= 32 x 256 floats)
constexpr int ROWS_CS = 8;                                   // CTAs per cluster: one L2 read of the weights feeds 8 SMs
constexpr int ROWS_BLOCK = ROWS_THREADS + 32;                // 8 consumer warps + 1 producer warp
constexpr int ROWS_MAX_STAGES = 6;

__device__ __forceinline__ unsigned long long rows_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- mbarrier / cluster primitives -----------------------------------------------------------------
__device__ __forceinline__ uint32_t rows_su32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rows_su32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rows_su32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n RW_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra RD_%=;\n bra RW_%=;\n RD_%=:\n}\n" ::"r"(
          rows_su32(b)),
      "r"(parity)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* b, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(rows_su32(b)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// global -> shared memory of EVERY CTA of the cluster (same offset), completing `bytes` on each CTA's barrier
__device__ __forceinline__ void bulk_multicast(float* dst, const float* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                   rows_su32(dst)),
               "l"(src), "r"(bytes), "r"(rows_su32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void rows_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t rows_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(ROWS_THREADS) : "memory"); }

// bytes of the weight stage (layer L, contraction rows [k0, k0+32))
__device__ __forceinline__ uint32_t rows_stage_bytes(const RowLayer& L, int k0) {
  if (L.mode == GEMM_FWD) return uint32_t(L.N) * uint32_t(min(ROWS_KC, ((L.K + 3) & ~3) - k0)) * 4u;
  return uint32_t(min(ROWS_KC, L.K - k0)) * uint32_t((L.N + 3) & ~3) * 4u;
}
// this CTA's share of one weight stage, multicast to the whole cluster (one warp)
__device__ __forceinline__ void rows_issue_stage(const RowLayer& L, int k0, float* Ws, uint64_t* bar, uint32_t rank, int lane) {
  const float* __restrict__ W = L.W;
  const int ldw = L.ldw;
  const uint16_t mask = uint16_t((1u << ROWS_CS) - 1u);
  if (L.mode == GEMM_FWD) {                                  // Ws[j][k - k0] = W[j][k]: one 128-B row per copy
    const uint32_t bytes = uint32_t(min(ROWS_KC, ((L.K + 3) & ~3) - k0)) * 4u;
    for (int j = int(rank) * 32 + lane; j < L.N; j += ROWS_CS * 32)
      bulk_multicast(Ws + j * ROWS_WP, W + size_t(j) * ldw + k0, bytes, bar, mask);
  } else {                                                   // Ws[k - k0][n] = W[k][n]: one N-float row per copy
    const int kn = min(ROWS_KC, L.K - k0), N4 = (L.N + 3) & ~3;
    for (int kk = lane * ROWS_CS + int(rank); kk < kn; kk += 32 * ROWS_CS)
      bulk_multicast(Ws + kk * N4, W + size_t(k0 + kk) * ldw, uint32_t(N4) * 4u, bar, mask);
  }
}

// One stage of thread `tid`'s output column.  The additions are blocked like a vectorised CPU dot
// product: four interleaved partial sums per stage (16 terms each), folded pairwise into the layer total.
template <int R>
__device__ __forceinline__ void rows_compute_stage(const RowLayer& L, int k0, const float* __restrict__ Ws,
                                                   const float* __restrict__ xin, int pitch, float (&acc)[R], int tid) {
  if (tid >= L.N) return;
  float p[R][4];
#pragma unroll
  for (int r = 0; r < R; ++r) { p[r][0] = 0.f; p[r][1] = 0.f; p[r][2] = 0.f; p[r][3] = 0.f; }
  if (L.mode == GEMM_FWD) {
    const int K4 = (L.K + 3) & ~3;
    const int kn4 = min(ROWS_KC, K4 - k0);
    const float* __restrict__ wrow = Ws + tid * ROWS_WP;
    const float* __restrict__ xk = xin + k0;
#pragma unroll 4
    for (int k4 = 0; k4 < kn4; k4 += 4) {
      const float4 w = *reinterpret_cast<const float4*>(wrow + k4);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(xk + r * pitch + k4);
        p[r][0] = fmaf(w.x, x.x, p[r][0]); p[r][1] = fmaf(w.y, x.y, p[r][1]);
        p[r][2] = fmaf(w.z, x.z, p[r][2]); p[r][3] = fmaf(w.w, x.w, p[r][3]);
      }
    }
  } else {
    const int kn = min(ROWS_KC, L.K - k0), N4 = (L.N + 3) & ~3;
    const int kfull = kn & ~3;
    const float* __restrict__ wcol = Ws + tid;
    const float* __restrict__ xk = xin + k0;
#pragma unroll 4
    for (int kk = 0; kk < kfull; kk += 4) {
      const float w0 = wcol[kk * N4], w1 = wcol[(kk + 1) * N4], w2 = wcol[(kk + 2) * N4], w3 = wcol[(kk + 3) * N4];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(xk + r * pitch + kk);
        p[r][0] = fmaf(x.x, w0, p[r][0]); p[r][1] = fmaf(x.y, w1, p[r][1]);
        p[r][2] = fmaf(x.z, w2, p[r][2]); p[r][3] = fmaf(x.w, w3, p[r][3]);
      }
    }
    for (int kk = kfull; kk < kn; ++kk) {
      const float w = wcol[kk * N4];
#pragma unroll
      for (int r = 0; r < R; ++r) p[r][0] = fmaf(xk[r * pitch + kk], w, p[r][0]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] += (p[r][0] + p[r][1]) + (p[r][2] + p[r][3]);
}

// the 8 consumer warps: R rows through every layer of the chain
template <int R>
__device__ __forceinline__ void rows_consume(const RowsArgs& a, const RowChain& ch, int m0, float* smem, uint64_t* full,
                                             uint64_t* empty, unsigned long long* tr) {
  const int tid = threadIdx.x, lane = tid & 31, B = a.B;
  const int nst = a.nstages;
  float* ring = smem + a.ring_off;

  // ---- the chain's inputs: R rows of the sampled batch / of the logit gradients ---------------------
  {
    float* xg = smem + a.xoff[XB_IN];
    const int pg = a.pitch[XB_IN], k4 = (ch.k_in0 + 3) & ~3;
    for (int r = 0; r < R; ++r)
      for (int c = tid * 4; c < k4; c += ROWS_THREADS * 4) {
        if (m0 + r < B) cp_async16(xg + r * pg + c, ch.in0 + size_t(m0 + r) * ch.ld_in0 + c);
        else *reinterpret_cast<float4*>(xg + r * pg + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    if (ch.in1) {                                            // replay action -> tail of the concatenated input
      float* xc = smem + a.xoff[XB_CAT] + ch.in1_off;
      const int pc = a.pitch[XB_CAT], kpad = (ch.k_in1 + 3) & ~3;
      for (int e = tid; e < R * kpad; e += ROWS_THREADS) {
        const int r = e / kpad, c = e - r * kpad;
        xc[r * pc + c] = (c < ch.k_in1 && m0 + r < B) ? __ldg(ch.in1 + size_t(m0 + r) * ch.ld_in1 + c) : 0.f;
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    consumer_sync();
  }

  float acc[R], em[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { acc[r] = 0.f; em[r] = 0.f; }
  int slot = 0; uint32_t phase = 0;
  long long t_wait = 0, t_comp = 0, t_epi = 0;               // SM cycles of thread 0 (trace only)
  for (int l = 0; l < ch.nlayers; ++l) {
    const RowLayer& L = ch.layer[l];
    const float* xin = smem + a.xoff[L.in_buf];
    const int pin = a.pitch[L.in_buf];
    // epilogue operands (bias or forward activations): independent of the chain, fetched early
    float eb = 0.f;
    if (tid < L.N) {
      if (L.mode == GEMM_FWD) eb = __ldg(L.bias + tid);
      else if (L.epi != EPI_NONE) {
#pragma unroll
        for (int r = 0; r < R; ++r) em[r] = (m0 + r < B) ? __ldg(L.aux + size_t(m0 + r) * L.ldaux + tid) : 0.f;
      }
    }
    if (tr) tr[2 * l] = rows_gtime();
    for (int k0 = 0; k0 < L.K; k0 += ROWS_KC) {
      const long long c0 = tr ? clock64() : 0;
      mbar_wait(&full[slot], phase);                         // the multicast stage has landed in this CTA
      const long long c1 = tr ? clock64() : 0;
      rows_compute_stage<R>(L, k0, ring + slot * a.stage_floats, xin, pin, acc, tid);
      __syncwarp();
      if (tr) { t_wait += c1 - c0; t_comp += clock64() - c1; }
      if (lane < ROWS_CS) mbar_arrive_remote(&empty[slot], uint32_t(lane));   // this warp is done with the slot, cluster-wide
      if (++slot == nst) { slot = 0; phase ^= 1u; }
    }
    if (tr) tr[2 * l + 1] = rows_gtime();
    const long long c2 = tr ? clock64() : 0;
    // ---- layer epilogue: this thread's column for all R rows --------------------------------------
    const int N = L.N, N4 = (N + 3) & ~3;
    if (tid < N4) {
      float* xout = L.out_buf >= 0 ? smem + a.xoff[L.out_buf] + L.out_off + tid : nullptr;
      const int pout = L.out_buf >= 0 ? a.pitch[L.out_buf] : 0;
      float* __restrict__ C = L.C;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float v = acc[r];
        acc[r] = 0.f;
        if (tid < N) {
          switch (L.epi) {
            case EPI_BIAS: v += eb; break;
            case EPI_BIAS_RELU: v = fmaxf(v + eb, 0.f); break;
            case EPI_BIAS_TANH: v = tanhf(v + eb); break;
            case EPI_RELU_MASK: v = (em[r] > 0.f) ? v : 0.f; break;
            case EPI_TANH_MASK: v *= (1.f - em[r] * em[r]); break;
            default: break;
          }
          if (C && m0 + r < B) C[size_t(m0 + r) * L.ldc + tid] = v;
        } else v = 0.f;                                      // pad columns up to a multiple of 4 read as zero
        if (xout) xout[r * pout] = v;
      }
    }
    consumer_sync();                                         // the next layer reads these activations
    if (tr) t_epi += clock64() - c2;
  }
  if (tr) { tr[2 * ROWS_MAX_LAYERS + 0] = t_wait; tr[2 * ROWS_MAX_LAYERS + 1] = t_comp; tr[2 * ROWS_MAX_LAYERS + 2] = t_epi; }
}

// CTA = 8 consumer warps + 1 producer warp; cluster = 8 CTAs of the same chain sharing every weight fetch
__global__ void __launch_bounds__(ROWS_BLOCK, 1) mlp_rows_kernel(const __grid_constant__ RowsArgs a) {
  extern __shared__ __align__(128) float rows_smem[];
  __shared__ __align__(8) uint64_t full_bar[ROWS_MAX_STAGES], empty_bar[ROWS_MAX_STAGES];
  int c = 0;
  if (a.nchains > 1 && int(blockIdx.x) >= a.chain[1].cta_begin) c = 1;
  if (a.nchains > 2 && int(blockIdx.x) >= a.chain[2].cta_begin) c = 2;
  const RowChain& ch = a.chain[c];
  const int m0 = (int(blockIdx.x) - ch.cta_begin) * ch.R;    // CTAs past the batch (cluster padding) own no rows
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nst = a.nstages;
  if (tid == 0) {
    for (int i = 0; i < nst; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], (ROWS_THREADS / 32) * ROWS_CS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  rows_cluster_sync();                                       // every CTA's barriers exist before any remote traffic

  if (warp == ROWS_THREADS / 32) {
    // ---- producer warp: stream the chain's weights, stage by stage, 1/8 of each stage per CTA --------
    const uint32_t rank = rows_cta_rank();
    float* ring = rows_smem + a.ring_off;
    int slot = 0; uint32_t phase = 0; bool wrapped = false;
    const bool ptrace = a.trace && blockIdx.x == 0 && lane == 0;
    long long t_empty = 0, t_issue = 0;
    for (int l = 0; l < ch.nlayers; ++l) {
      const RowLayer& L = ch.layer[l];
      for (int k0 = 0; k0 < L.K; k0 += ROWS_KC) {
        const long long c0 = ptrace ? clock64() : 0;
        if (wrapped) { if (lane == 0) mbar_wait(&empty_bar[slot], phase ^ 1u); __syncwarp(); }   // all 8 CTAs left the slot
        const long long c1 = ptrace ? clock64() : 0;
        if (lane == 0) mbar_expect_tx(&full_bar[slot], rows_stage_bytes(L, k0));
        __syncwarp();
        rows_issue_stage(L, k0, ring + slot * a.stage_floats, &full_bar[slot], rank, lane);
        if (ptrace) { t_empty += c1 - c0; t_issue += clock64() - c1; }
        if (++slot == nst) { slot = 0; phase ^= 1u; wrapped = true; }
      }
    }
    if (ptrace) { a.trace[2 * ROWS_MAX_LAYERS + 3] = t_empty; a.trace[2 * ROWS_MAX_LAYERS + 4] = t_issue; }
  } else {
    unsigned long long* tr = (a.trace && blockIdx.x == 0 && tid == 0) ? a.trace : nullptr;
    switch (ch.R) {
      case 5: rows_consume<5>(a, ch, m0, rows_smem, full_bar, empty_bar, tr); break;
      case 6: rows_consume<6>(a, ch, m0, rows_smem, full_bar, empty_bar, tr); break;
      case 8: rows_consume<8>(a, ch, m0, rows_smem, full_bar, empty_bar, tr); break;
      default: rows_consume<12>(a, ch, m0, rows_smem, full_bar, empty_bar, tr); break;
    }
  }
  rows_cluster_sync();                                       // no CTA leaves while peers may still signal its barriers
}

// ---- host side -------------------------------------------------------------------------------------

void rows_args_begin(RowsArgs& a, int B, int obs_dim, int act_dim) {
  a = RowsArgs{};
  a.B = B;
  (void)obs_dim; (void)act_dim;
}
RowLayer rows_fwd(const float* W, int ldw, const float* bias, int N, int K, int epi, float* C, int ldc,
                  int in_buf, int out_buf, int out_off) {
  RowLayer l{};
  l.W = W; l.ldw = ldw; l.bias = bias; l.N = N; l.K = K; l.epi = epi; l.C = C; l.ldc = ldc; l.mode = GEMM_FWD;
  l.in_buf = in_buf; l.out_buf = out_buf; l.out_off = out_off;
  return l;
}
RowLayer rows_dx(const float* W, int ldw, int N_in, int K_out, int epi, const float* aux, int ldaux, float* C, int ldc,
                 int in_buf, int out_buf, int out_off) {
  RowLayer l{};
  l.W = W; l.ldw = ldw; l.N = N_in; l.K = K_out; l.epi = epi; l.aux = aux; l.ldaux = ldaux; l.C = C; l.ldc = ldc;
  l.mode = GEMM_DX; l.in_buf = in_buf; l.out_buf = out_buf; l.out_off = out_off;
  return l;
}
void rows_chain_input(RowsArgs& a, int c, const float* in0, int ld, int k) {
  if (c >= a.nchains) a.nchains = c + 1;
  a.chain[c].in0 = in0; a.chain[c].ld_in0 = ld; a.chain[c].k_in0 = k;
}
void rows_chain_input2(RowsArgs& a, int c, const float* in1, int ld, int k, int col_off) {
  a.chain[c].in1 = in1; a.chain[c].ld_in1 = ld; a.chain[c].k_in1 = k; a.chain[c].in1_off = col_off;
}
void rows_add(RowsArgs& a, int c, const RowLayer& l) {
  if (c >= a.nchains) a.nchains = c + 1;
  RowChain& ch = a.chain[c];
  if (ch.nlayers < ROWS_MAX_LAYERS) ch.layer[ch.nlayers] = l;
  ++ch.nlayers;                                              // overflow is reported by rows_finalize
}

static int rows_stage_count(const RowChain& ch) {
  int s = 0;
  for (int l = 0; l < ch.nlayers; ++l) s += cdiv(ch.layer[l].K, ROWS_KC);
  return s;
}

int rows_finalize(RowsArgs& a) {
  D4PG_REQUIRE(a.nchains > 0 && a.nchains <= ROWS_MAX_CHAINS, D4PG_EINVAL, "rows_finalize: %d chains", a.nchains);
  int pc = D4PG_HIDDEN, pg = 4;
  for (int c = 0; c < a.nchains; ++c) {
    const RowChain& ch = a.chain[c];
    D4PG_REQUIRE(ch.nlayers > 0 && ch.nlayers <= ROWS_MAX_LAYERS, D4PG_ENOTSUP, "rows_finalize: chain %d has %d layers", c, ch.nlayers);
    D4PG_REQUIRE(ch.in0 && ch.ld_in0 % 4 == 0 && (reinterpret_cast<uintptr_t>(ch.in0) & 15) == 0, D4PG_EINVAL,
                 "rows_finalize: chain %d input must be 16-B aligned with a 16-B row pitch", c);
    pg = std::max(pg, (ch.k_in0 + 3) & ~3);
    if (ch.in1) pc = std::max(pc, ch.in1_off + ((ch.k_in1 + 3) & ~3));
    for (int l = 0; l < ch.nlayers; ++l) {
      const RowLayer& L = ch.layer[l];
      D4PG_REQUIRE(L.N > 0 && L.N <= ROWS_THREADS, D4PG_ENOTSUP, "rows_finalize: layer width %d > %d", L.N, ROWS_THREADS);
      D4PG_REQUIRE(L.ldw % 4 == 0 && (reinterpret_cast<uintptr_t>(L.W) & 15) == 0, D4PG_EINVAL, "rows_finalize: weights must be 16-B pitched");
      D4PG_REQUIRE(L.in_buf >= 0 && L.in_buf < XB_COUNT && L.out_buf < XB_COUNT && L.in_buf != L.out_buf, D4PG_EINVAL,
                   "rows_finalize: bad activation buffers in chain %d layer %d", c, l);
      D4PG_REQUIRE(L.mode == GEMM_FWD ? L.ldw >= ((L.K + 3) & ~3) : L.ldw >= ((L.N + 3) & ~3), D4PG_EINVAL,
                   "rows_finalize: weight pitch too small in chain %d layer %d", c, l);
      const int need_in = (L.K + 3) & ~3, need_out = L.out_off + ((L.N + 3) & ~3);
      if (L.in_buf == XB_CAT) pc = std::max(pc, need_in);
      if (L.in_buf == XB_IN) pg = std::max(pg, need_in);
      if (L.out_buf == XB_CAT) pc = std::max(pc, need_out);
      D4PG_REQUIRE((L.in_buf > XB_PONG || need_in <= D4PG_HIDDEN) && (L.out_buf > XB_PONG || L.out_buf < 0 || need_out <= D4PG_HIDDEN),
                   D4PG_ENOTSUP, "rows_finalize: chain %d layer %d does not fit the ping-pong buffers", c, l);
    }
  }
  // rows per CTA: minimise waves x (stages x per-stage cost) over R in {5, 6, 8, 12} per chain.  CTAs come in
  // clusters of 8 (one per chain slice); a B200 holds 16 such clusters at once (2 per GPC).
  const int cand[4] = {5, 6, 8, 12};
  const int cluster_capacity = 16;
  double best = 1e300; int bestR[ROWS_MAX_CHAINS] = {12, 12, 12};
  int idx[ROWS_MAX_CHAINS] = {0, 0, 0};
  const int combos = a.nchains == 1 ? 4 : a.nchains == 2 ? 16 : 64;
  for (int m = 0; m < combos; ++m) {
    idx[0] = m % 4; idx[1] = (m / 4) % 4; idx[2] = m / 16;
    int clusters = 0; double worst = 0;
    for (int c = 0; c < a.nchains; ++c) {
      const int R = cand[idx[c]];
      clusters += cdiv(cdiv(a.B, R), ROWS_CS);
      worst = std::max(worst, double(rows_stage_count(a.chain[c])) * (64.0 * (4 + R) + 100.0));
    }
    const double cost = worst * cdiv(clusters, cluster_capacity) + 1e-3 * clusters;
    if (cost < best) { best = cost; for (int c = 0; c < a.nchains; ++c) bestR[c] = cand[idx[c]]; }
  }
  int Rmax = 0, begin = 0;
  for (int c = 0; c < a.nchains; ++c) {
    a.chain[c].R = bestR[c]; a.chain[c].cta_begin = begin; a.chain[c].ctas = cdiv(cdiv(a.B, bestR[c]), ROWS_CS) * ROWS_CS;
    begin += a.chain[c].ctas; Rmax = std::max(Rmax, bestR[c]);
  }
  a.total_ctas = begin;
  a.pitch[XB_PING] = a.pitch[XB_PONG] = D4PG_HIDDEN; a.pitch[XB_CAT] = pc; a.pitch[XB_IN] = pg;
  int off = 0;
  for (int b = 0; b < XB_COUNT; ++b) { a.xoff[b] = off; off += Rmax * a.pitch[b]; }
  a.ring_off = (off + 31) & ~31; a.stage_floats = ROWS_STAGE_FLOATS;       // 128-B aligned bulk-copy destinations
  const int64_t budget = 220 * 1024 / 4 - a.ring_off;
  a.nstages = int(std::min<int64_t>(ROWS_MAX_STAGES, budget / ROWS_STAGE_FLOATS));
  D4PG_REQUIRE(a.nstages >= 2, D4PG_ENOTSUP, "rows_finalize: activations of %d rows leave no room for the weight ring", Rmax);
  return D4PG_OK;
}

int launch_mlp_rows(RowsArgs& a, cudaStream_t st) {
  D4PG_REQUIRE(a.total_ctas > 0 && a.nstages >= 2, D4PG_ESTATE, "launch_mlp_rows: call rows_finalize first");
  const size_t smem = size_t(a.ring_off + a.nstages * a.stage_floats) * sizeof(float);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    D4PG_CUDA_OK(cudaFuncSetAttribute(mlp_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    smem_set = smem;
  }
  D4PG_MAX_CARVEOUT(mlp_rows_kernel);
  a.trace = debug_trace_buffer() ? debug_trace_buffer() + a.trace_base : nullptr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a.total_ctas); cfg.blockDim = dim3(ROWS_BLOCK); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ROWS_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  D4PG_CUDA_OK(cudaLaunchKernelEx(&cfg, mlp_rows_kernel, a));
  return D4PG_OK;
}

}  // namespace d4pg
