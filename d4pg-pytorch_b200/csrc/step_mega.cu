// Persistent "one kernel per gradient step" variant of the learner (precision fp32).
//
// At batch 256 a step is ~18 small dependent kernels and the kernel boundaries (grid drain + next
// launch, ~3-5 us each on B200) cost about as much as the kernels (profiles/README.md).  This kernel
// runs the SAME device code (sample_body, gemm_tile, heads_row, tree_write_body, adam_segment) as
// phases of one cooperative launch with one CTA pair per SM resident for the whole step, separated
// by a grid-wide barrier (one atomic + spin on an L2 flag, ~1 us) instead of a kernel boundary:
//
//   phase 0        sample + gather (ddpg.py:202)                          CTAs [0, B/32)
//   phases 1..7    forward levels  (ddpg.py:205-208,236)                  tiles strided over all CTAs
//   phase 8        projection / losses / priorities / logit gradients     one warp per row
//   phases 9..15   backward levels (ddpg.py:229-243)                      tiles strided over CTAs [0, G-1)
//                  ... while the LAST CTA writes the priorities into the trees (ddpg.py:252-255)
//                      and leaves (it takes part in no later barrier)
//   phase 16       Adam + Polyak + reported losses + clock advance (ddpg.py:232,244,247,250)
//
// The barrier counter only ever increases: step k uses the targets k*T + (cumulative arrivals),
// with k read from the device clock (advanced by the last phase), so a graph replay needs no reset.
#include "gemm_ffma_dev.cuh"
#include "heads_dev.cuh"
#include "replay_dev.cuh"
#include "adam_dev.cuh"
#include "step_mega.cuh"

namespace d4pg {

// Arrivals are counted with one atomic each on `bar[0]`; the LAST arriver publishes the reached target
// in `bar[16]` (its own 128-B line) and everybody else polls that read-mostly line with back-off, so
// the atomics and the polls do not fight over one L2 line.
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long old = atomicAdd(bar, 1ull);
    unsigned long long* flag = bar + 16;
    if (old + 1 == target) {
      asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flag), "l"(target) : "memory");
    } else {
      unsigned long long v;
      while (true) {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= target) break;
        __nanosleep(40);
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ unsigned long long mega_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define MTRACE(slot) do { if (p.trace && bid == 0 && tid == 0) p.trace[slot] = mega_gtime(); } while (0)

union MegaSmem {
  float gemm[GEMM_SMEM_FLOATS];
  HeadsWarpSmem heads[GEMM_WARPS];
  SampleSmem sample;
  float red[2][8];
  float tree_red[32];
};

__global__ void __launch_bounds__(GEMM_THREADS, 2) step_mega_kernel(const __grid_constant__ MegaParams p) {
  __shared__ __align__(16) MegaSmem sm;
  const int G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x;
  const int n_levels = p.n_fwd + p.n_bwd;
  // arrivals per step: (1 + n_fwd + 1) barriers with G CTAs, then n_bwd barriers with G-1 (tree CTA gone)
  const unsigned long long per_step = (unsigned long long)(2 + p.n_fwd) * G + (unsigned long long)p.n_bwd * (G - 1);
  unsigned long long target = (unsigned long long)(p.clock->mega_epoch) * per_step;

  int slot = 0;
  MTRACE(slot); ++slot;
  // ---- phase 0: sample + gather ----------------------------------------------------------------------
  if (bid < cdiv(p.sample.B, SAMPLE_ROWS)) sample_body(p.sample, bid, sm.sample);
  MTRACE(slot); ++slot;
  target += G; grid_barrier(p.barrier, target);
  MTRACE(slot); ++slot;

  // ---- forward levels ---------------------------------------------------------------------------------
  for (int lv = 0; lv < p.n_fwd; ++lv) {
    const GemmBatchLite& b = p.level[lv];
    for (int tile = bid; tile < b.total_tiles; tile += G) {
      int pi = 0;
      for (int i = 1; i < b.n; ++i) if (tile >= b.p[i].tile_begin) pi = i;
      const GemmProblem P = b.p[pi];
      gemm_tile_dispatch<false>(P, sm.gemm, tile - P.tile_begin);
      __syncthreads();                                     // smem reuse by the next tile
    }
    target += G; grid_barrier(p.barrier, target);
    MTRACE(slot); ++slot;
  }

  // ---- heads: one warp per batch row -------------------------------------------------------------------
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int row = bid * GEMM_WARPS + warp; row < p.heads.B; row += G * GEMM_WARPS) {
      if (p.heads.N <= 64) {
        if (p.heads_mode == 0) heads_row<0, 2>(p.heads, row, lane, sm.heads[warp]);
        else heads_row<1, 2>(p.heads, row, lane, sm.heads[warp]);
      } else {
        if (p.heads_mode == 0) heads_row<0, 4>(p.heads, row, lane, sm.heads[warp]);
        else heads_row<1, 4>(p.heads, row, lane, sm.heads[warp]);
      }
      __syncwarp();
    }
  }
  MTRACE(slot); ++slot;
  target += G; grid_barrier(p.barrier, target);
  MTRACE(slot); ++slot;

  // ---- priorities -> trees on the last CTA, which then retires -------------------------------------------
  const int Gw = G - 1;                                        // CTAs that keep working
  if (bid == G - 1) {
    if (p.do_tree) tree_write_body<TREE_UPDATE, GEMM_THREADS>(p.tree, sm.tree_red);
    return;
  }

  // ---- backward levels ---------------------------------------------------------------------------------
  for (int lv = p.n_fwd; lv < n_levels; ++lv) {
    const GemmBatchLite& b = p.level[lv];
    for (int tile = bid; tile < b.total_tiles; tile += Gw) {
      int pi = 0;
      for (int i = 1; i < b.n; ++i) if (tile >= b.p[i].tile_begin) pi = i;
      const GemmProblem P = b.p[pi];
      gemm_tile_dispatch<false>(P, sm.gemm, tile - P.tile_begin);
      __syncthreads();
    }
    target += Gw; grid_barrier(p.barrier, target);
    MTRACE(slot); ++slot;
  }

  // ---- Adam + Polyak (+ losses, clock) -------------------------------------------------------------------
  for (int seg = 0; seg < p.adam.nseg; ++seg) adam_segment(p.adam, seg, bid, Gw);
  if (bid == 0) {
    adam_tail(p.adam, sm.red);
    if (tid == 0) p.clock->mega_epoch += 1;     // every CTA read it before the first barrier of this step
  }
  MTRACE(slot);
}

int launch_step_mega(const MegaParams& p, cudaStream_t st) {
  static int grid = 0;
  if (grid == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    D4PG_CUDA_OK(cudaGetDevice(&dev));
    D4PG_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    D4PG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_mega_kernel, GEMM_THREADS, 0));
    D4PG_REQUIRE(per_sm >= 1, D4PG_ECUDA, "step_mega_kernel does not fit on an SM");
    grid = sms * (per_sm > 2 ? 2 : per_sm);
  }
  D4PG_REQUIRE(cdiv(p.sample.B, SAMPLE_ROWS) <= grid, D4PG_ENOTSUP, "batch too large for the persistent step kernel");
  void* args[] = {const_cast<MegaParams*>(&p)};
  D4PG_CUDA_OK(cudaLaunchCooperativeKernel((const void*)step_mega_kernel, dim3(grid), dim3(GEMM_THREADS), args, 0, st));
  return D4PG_OK;
}

}  // namespace d4pg
