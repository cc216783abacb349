// Fused Adam (beta=(0.9,0.9) by default) + Polyak soft target update over flat fp32 buffers.
//
// Replaces (reference, relative to /root/reference):
//   shared_adam.py:3-17 + torch.optim.Adam.step as called at ddpg.py:232,244.  Arithmetic follows
//     torch 2.11 `_single_tensor_adam` (the form the reference executes in this image):
//       m <- lerp(m, g, 1-b1); v <- v*b2 + (1-b2)*g*g;
//       p <- p - (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps),  bc = 1 - b^step
//   ddpg.py:118-120 sync_local_global: identity (local and global share storage)
//   ddpg.py:110-116 update_target_parameters: t <- (1-tau)*t + tau*p   (p = post-step value)
// Pure streaming kernel: 4 reads + 4 writes of P floats -> HBM/L2 bound, float4 vectorised.
#include "adam.cuh"
#include <math.h>
#include <algorithm>

namespace d4pg {

// mean over the batch of the per-row loss terms (ddpg.py:217 `.mean()`, ddpg.py:238 `.mean()`),
// fixed-order single-block reduction so the reported scalars are run-to-run deterministic.
__device__ void loss_reduce_block(const float* loss_rows, const float* pi_rows, int B, float inv_count, float* out) {
  __shared__ float red[2][8];
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < B; i += 256) { a += loss_rows[i]; if (pi_rows) b += pi_rows[i]; }
  a = warp_sum(a); b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sa = 0.f, sb = 0.f;
    for (int w = 0; w < 8; ++w) { sa += red[0][w]; sb += red[1][w]; }
    out[0] = sa * inv_count; out[1] = sb * inv_count;
  }
}

__global__ void __launch_bounds__(256) adam_polyak_kernel(const AdamArgs a) {
  pdl_trigger();
  pdl_wait();
  if (int(blockIdx.y) == a.nseg) {
    // tail slice: reported losses + advance the base counters for the NEXT step (nobody in this
    // kernel reads them: the derived scalars were written by the step's first kernel)
    if (blockIdx.x != 0) return;
    if (a.loss_out) loss_reduce_block(a.loss_rows, a.pi_rows, a.B, a.inv_count, a.loss_out);
    if (threadIdx.x == 0 && a.clock) { a.clock->adam_step += 1; a.clock->beta_t += 1; a.clock->steps_done += 1; }
    return;
  }
  const AdamSeg& s = a.seg[blockIdx.y];
  const float nss = (a.clock && s.clock_slot >= 0) ? a.clock->neg_step_size[s.clock_slot] : s.neg_step_size;
  const float bc2s = a.clock ? a.clock->bc2_sqrt : a.bc2_sqrt;
  const int64_t n4 = s.n >> 2;
  float4* p4 = reinterpret_cast<float4*>(s.p);
  const float4* g4 = reinterpret_cast<const float4*>(s.g);
  float4* m4 = reinterpret_cast<float4*>(s.m);
  float4* v4 = reinterpret_cast<float4*>(s.v);
  float4* t4 = reinterpret_cast<float4*>(s.target);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
    float4 t = s.target ? t4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float* pp = &p.x; float* gg = &g.x; float* mm = &m.x; float* vv = &v.x; float* tt = &t.x;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float gr = gg[c] * a.grad_scale;
      mm[c] = fmaf(a.w1, gr - mm[c], mm[c]);                                   // lerp, weight < 0.5
      vv[c] = __fadd_rn(__fmul_rn(vv[c], a.beta2), __fmul_rn(__fmul_rn(a.w2, gr), gr));
      const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv[c]), bc2s), a.eps);
      pp[c] = __fadd_rn(pp[c], __fmul_rn(nss, __fdiv_rn(mm[c], denom)));
      tt[c] = __fadd_rn(__fmul_rn(a.one_minus_tau, tt[c]), __fmul_rn(a.tau, pp[c]));
    }
    p4[i] = p; m4[i] = m; v4[i] = v;
    if (s.target) t4[i] = t;
  }
}

int launch_adam(const AdamArgs& a, cudaStream_t st) {
  int64_t nmax = 0;
  for (int i = 0; i < a.nseg; ++i) nmax = a.seg[i].n > nmax ? a.seg[i].n : nmax;
  int blocks = int((nmax / 4 + 255) / 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  D4PG_MAX_CARVEOUT(adam_polyak_kernel);
  const int tail = (a.clock || a.loss_out) ? 1 : 0;
  D4PG_CUDA_OK(launch_pdl(adam_polyak_kernel, dim3(blocks, a.nseg + tail), dim3(256), 0, st, a));
  return D4PG_OK;
}

}  // namespace d4pg

using namespace d4pg;

extern "C" int32_t d4pg_adam_polyak(float* p, const float* g, float* m, float* v, float* target, int64_t n,
                                    double lr, double beta1, double beta2, double eps, int64_t step,
                                    double tau, float grad_scale, d4pg_stream_t stream) {
  D4PG_REQUIRE(p && g && m && v, D4PG_EINVAL, "d4pg_adam_polyak: null buffer");
  D4PG_REQUIRE(n > 0 && n % 4 == 0, D4PG_EINVAL, "d4pg_adam_polyak: n must be a positive multiple of 4 (flat layout is 4-aligned)");
  D4PG_REQUIRE(step >= 1, D4PG_EINVAL, "d4pg_adam_polyak: step is the post-increment count (>= 1)");
  AdamArgs a{};
  const double bc1 = 1.0 - pow(beta1, double(step));
  const double bc2 = 1.0 - pow(beta2, double(step));
  a.seg[0] = AdamSeg{p, g, m, v, target, n, float(-(lr / bc1)), -1};
  a.nseg = 1;
  a.w1 = float(1.0 - beta1); a.w2 = float(1.0 - beta2); a.beta2 = float(beta2); a.eps = float(eps);
  a.bc2_sqrt = float(sqrt(bc2)); a.tau = float(tau); a.one_minus_tau = float(1.0 - tau);
  a.grad_scale = grad_scale; a.clock = nullptr; a.loss_out = nullptr;
  return launch_adam(a, as_stream(stream));
}

namespace d4pg {
__global__ void polyak_kernel(float* t, const float* s, int64_t n, float tau, float omt) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    t[i] = __fadd_rn(__fmul_rn(omt, t[i]), __fmul_rn(tau, s[i]));
}
}  // namespace d4pg

extern "C" int32_t d4pg_polyak(float* target, const float* src, int64_t n, double tau, d4pg_stream_t stream) {
  D4PG_REQUIRE(target && src && n > 0, D4PG_EINVAL, "d4pg_polyak: bad arguments");
  int blocks = int(std::min<int64_t>(592, (n + 255) / 256));
  polyak_kernel<<<blocks, 256, 0, as_stream(stream)>>>(target, src, n, float(tau), float(1.0 - tau));
  D4PG_LAUNCH_OK();
  return D4PG_OK;
}

extern "C" int32_t d4pg_copy_f32(float* dst, const float* src, int64_t n, d4pg_stream_t stream) {
  D4PG_REQUIRE(dst && src && n >= 0, D4PG_EINVAL, "d4pg_copy_f32: bad arguments");
  D4PG_CUDA_OK(cudaMemcpyAsync(dst, src, size_t(n) * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(stream)));
  return D4PG_OK;
}
