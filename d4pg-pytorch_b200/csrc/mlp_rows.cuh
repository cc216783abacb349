// Row-owner MLP chains (exact fp32): a CTA carries R batch rows through every layer of a chain.
#pragma once
#include "gemm_ffma.cuh"

namespace d4pg {

constexpr int ROWS_MAX_LAYERS = 8;
constexpr int ROWS_MAX_CHAINS = 3;
constexpr int ROWS_THREADS = 256;        // one thread per output column of a 256-wide layer
constexpr int ROWS_KC = 32;              // contraction rows per streamed weight stage

// activation buffers in shared memory
enum { XB_PING = 0, XB_PONG = 1, XB_CAT = 2, XB_IN = 3, XB_COUNT = 4 };

struct RowLayer {
  const float* W; const float* bias; const float* aux;
  float* C;                         // row-major global copy of the output [B][ldc] (nullptr: none)
  int ldw, ldaux, ldc;
  int N, K;
  int mode, epi;                    // GEMM_FWD: y = act(W x + b), W[N][ldw];  GEMM_DX: y = (x W) * act', W[K][ldw]
  int in_buf, out_buf, out_off;     // out_buf < 0: the output only goes to global memory
};

struct RowChain {
  RowLayer layer[ROWS_MAX_LAYERS];
  int nlayers;
  const float* in0; int ld_in0, k_in0;              // global rows -> XB_IN columns [0, k_in0)
  const float* in1; int ld_in1, k_in1, in1_off;     // optional second input -> XB_CAT columns [in1_off, in1_off + k_in1)
  int R;                                            // batch rows per CTA (5, 6, 8 or 12)
  int cta_begin, ctas;
};

struct RowsArgs {
  RowChain chain[ROWS_MAX_CHAINS];
  int nchains, B;
  int pitch[XB_COUNT];              // floats per row of each activation buffer (multiples of 4)
  int xoff[XB_COUNT];               // float offset of each buffer in dynamic shared memory (sized for the largest R)
  int ring_off, stage_floats, nstages;
  int total_ctas;
  unsigned long long* trace;        // optional phase stamps of CTA 0 (D4PG_TC_TRACE)
  int trace_base;
};

void rows_args_begin(RowsArgs& a, int B, int obs_dim, int act_dim);
RowLayer rows_fwd(const float* W, int ldw, const float* bias, int N, int K, int epi, float* C, int ldc,
                  int in_buf, int out_buf, int out_off);
RowLayer rows_dx(const float* W, int ldw, int N_in, int K_out, int epi, const float* aux, int ldaux, float* C, int ldc,
                 int in_buf, int out_buf, int out_off);
void rows_chain_input(RowsArgs& a, int c, const float* in0, int ld, int k);
void rows_chain_input2(RowsArgs& a, int c, const float* in1, int ld, int k, int col_off);
void rows_add(RowsArgs& a, int c, const RowLayer& l);
// picks the rows per CTA of every chain (so that the launch is one wave when possible) and the smem plan
int rows_finalize(RowsArgs& a);
int launch_mlp_rows(RowsArgs& a, cudaStream_t st);

}  // namespace d4pg
