// Fused categorical projection + critic CE loss + TD proxy / priority + logit gradients
// (+ the policy-loss head).  One warp per batch row; the atom support is staged in shared
// memory and every row reduction is a warp shuffle.
//
// Replaces (reference, relative to /root/reference):
//   ddpg.py:142-185  DDPG.reproject2              (proj_mode 0, the live projection)
//   ddpg.py:122-140  DDPG.reproj_categorical_dist (proj_mode 1, gamma**n, config 5)
//   ddpg.py:217      qdist_loss = -(m*log(q+1e-10)).sum(1).mean()
//   ddpg.py:220-222  td_errors  = -(m*q).sum(1);  ddpg.py:253 priorities = |td| + eps
//   ddpg.py:236-238  policy_loss = -critic(s,actor(s)).matmul(bin_centers).mean()
//   models.py:83     F.softmax(fc3(out), dim=1) for the three critic heads
//
// Bit-exactness plan (SURVEY.md section 7 "hard parts"): the bin index path is fp64 with explicit
// round-to-nearest intrinsics (no FMA contraction of r + c_j); the projected mass of mode 0 is
// accumulated per bin in atom order j=0..N-1, each add done in fp64 and rounded to fp32, which
// is exactly what NumPy does for `proj_distr[rows, l] += p * (u - b)` on an fp32 array.
#include "internal.cuh"

namespace d4pg {

constexpr int HEAD_WARPS = 4;

// softmax of one row held as 4 values per lane (atom k = lane + 32*t); fp32, max-subtracted,
// exp then divide (models.py:83 -> torch softmax).
template <int NT>
__device__ __forceinline__ void row_softmax(const float* __restrict__ logits, int N, int lane, float (&p)[NT],
                                            bool already_probs = false) {
  if (already_probs) {
#pragma unroll
    for (int t = 0; t < NT; ++t) { int k = lane + 32 * t; p[t] = (k < N) ? __ldg(logits + k) : 0.f; }
    return;
  }
  float x[NT];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    int k = lane + 32 * t;
    x[t] = (k < N) ? __ldg(logits + k) : -INFINITY;
    mx = fmaxf(mx, x[t]);
  }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    int k = lane + 32 * t;
    p[t] = (k < N) ? expf(x[t] - mx) : 0.f;
    s += p[t];
  }
  s = warp_sum(s);
#pragma unroll
  for (int t = 0; t < NT; ++t) p[t] = p[t] / s;
}

template <int MODE, int NT>
__global__ void __launch_bounds__(HEAD_WARPS * 32) heads_kernel(const HeadsArgs a) {
  __shared__ float  p_s[HEAD_WARPS][D4PG_MAX_ATOMS];
  __shared__ int    l_s[HEAD_WARPS][D4PG_MAX_ATOMS];
  __shared__ int    u_s[HEAD_WARPS][D4PG_MAX_ATOMS];
  __shared__ double wl_s[HEAD_WARPS][D4PG_MAX_ATOMS];
  __shared__ double wu_s[HEAD_WARPS][D4PG_MAX_ATOMS];

  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * HEAD_WARPS + warp;
  if (row >= a.B) return;
  const int N = a.N;
  const size_t ro = size_t(row) * a.ld;

  // ---- target distribution ------------------------------------------------------------
  float p[NT];
  row_softmax(a.target_logits + ro, N, lane, p, (a.flags & D4PG_PROJ_TARGET_IS_PROBS) != 0);
  const double r = a.rewards[row];
  const bool done = a.dones[row] != 0;

  float mk[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) mk[t] = 0.f;

  if (MODE == 0 && done) {
    // ddpg.py:165-181: zero the row, Dirac at clip(r); weights cast f64 -> f32
    double tz = fmin(a.v_max, fmax(a.v_min, r));
    double b = __ddiv_rn(__dsub_rn(tz, a.v_min), a.delta);
    double lf = floor(b), uf = ceil(b);
    int l = int(lf), u = int(uf);
    float wl = (l == u) ? 1.0f : __double2float_rn(__dsub_rn(uf, b));
    float wu = __double2float_rn(__dsub_rn(b, lf));
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      int k = lane + 32 * t;
      if (k < N) {
        if (k == l) mk[t] = wl;
        else if (k == u) mk[t] = wu;
        if (a.bins_l) { a.bins_l[ro + k] = l; a.bins_u[ro + k] = u; }
      }
    }
  } else {
    // per-atom bins and weights in fp64 (ddpg.py:155-158 / ddpg.py:129-134)
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      int j = lane + 32 * t;
      if (j < N) {
        double zj = __dadd_rn(a.v_min, __dmul_rn(double(j), a.delta));
        double c;
        if (MODE == 0) c = __dmul_rn(zj, a.discount);                        // (v_min+j*delta)*gamma
        else c = __dmul_rn(__dmul_rn(a.discount, done ? 0.0 : 1.0), zj);      // gamma^n*(1-d)*z_j
        double tz = fmin(a.v_max, fmax(a.v_min, __dadd_rn(r, c)));
        double b = __ddiv_rn(__dsub_rn(tz, a.v_min), a.delta);
        double lf = floor(b), uf = ceil(b);
        int l = int(lf), u = int(uf);
        double wl, wu;
        if (MODE == 0) {
          if (l == u) { wl = 1.0; wu = 0.0; }
          else { wl = __dsub_rn(uf, b); wu = __dsub_rn(b, lf); }
        } else {
          if (l == u && u > 0) l -= 1;                                        // ddpg.py:133
          if (l == u && l < N - 1) u += 1;                                    // ddpg.py:134
          wl = __dsub_rn(double(u), b);
          wu = __dsub_rn(b, double(l));
        }
        p_s[warp][j] = p[t];
        l_s[warp][j] = l; u_s[warp][j] = u;
        wl_s[warp][j] = wl; wu_s[warp][j] = wu;
        if (a.bins_l) { a.bins_l[ro + j] = l; a.bins_u[ro + j] = u; }
      }
    }
    __syncwarp();
    // ordered per-bin accumulation (gather form: lane owns output bins, visits atoms in order).
    // b_j is non-decreasing in j, so the atoms that touch bin k (l_j == k or u_j == k) form one
    // contiguous run [j0, j1): two binary searches over the shared tables bound the loop to the few
    // atoms that matter (gamma < 1 => ~2-3 per bin; clamped atoms pile up only on the edge bins).
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int k = lane + 32 * t;
      if (k >= N) continue;
      int lo = 0, hi = N;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (u_s[warp][mid] >= k) hi = mid; else lo = mid + 1; }
      const int j0 = lo;
      hi = N;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (l_s[warp][mid] > k) hi = mid; else lo = mid + 1; }
      const int j1 = lo;
      if (MODE == 0) {
        float acc = 0.f;
        for (int j = j0; j < j1; ++j) {
          const int l = l_s[warp][j], u = u_s[warp][j];
          const double pj = double(p_s[warp][j]);
          if (k == l) {
            // eq: f32+f32 add; ne: f32 + (f64 product) in f64, rounded to f32
            const double term = (l == u) ? pj : __dmul_rn(pj, wl_s[warp][j]);
            acc = __double2float_rn(__dadd_rn(double(acc), term));
          } else if (k == u) {
            acc = __double2float_rn(__dadd_rn(double(acc), __dmul_rn(pj, wu_s[warp][j])));
          }
        }
        mk[t] = acc;
      } else {
        double acc = 0.;
        for (int j = j0; j < j1; ++j) {
          const int l = l_s[warp][j], u = u_s[warp][j];
          const double pj = double(p_s[warp][j]);
          if (k == l) acc = __dadd_rn(acc, __dmul_rn(pj, wl_s[warp][j]));
          if (k == u) acc = __dadd_rn(acc, __dmul_rn(pj, wu_s[warp][j]));
        }
        mk[t] = __double2float_rn(acc);
      }
    }
  }

  // ---- online critic: CE loss, TD proxy, priority, d loss / d logits -------------------
  float q[NT];
  row_softmax(a.q_logits + ro, N, lane, q, (a.flags & D4PG_PROJ_Q_IS_PROBS) != 0);
  float ce = 0.f, mq = 0.f, sq = 0.f;
  float gq[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    int k = lane + 32 * t;
    gq[t] = 0.f;
    if (k < N) {
      float qe = q[t] + 1e-10f;
      ce += mk[t] * logf(qe);
      mq += mk[t] * q[t];
      gq[t] = -(mk[t] / qe) * a.grad_scale;          // d mean-loss / d q_k
      sq += q[t] * gq[t];
    }
  }
  ce = warp_sum(ce); mq = warp_sum(mq); sq = warp_sum(sq);
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    int k = lane + 32 * t;
    if (k < N) {
      if (a.m) a.m[ro + k] = mk[t];
      if (a.target_probs) a.target_probs[ro + k] = p[t];
      if (a.q_probs) a.q_probs[ro + k] = q[t];
      if (a.dlogits_q) a.dlogits_q[ro + k] = q[t] * (gq[t] - sq);   // softmax backward
    }
  }
  if (lane == 0) {
    float tdv = -mq;
    if (a.loss_rows) a.loss_rows[row] = -ce;
    if (a.td) a.td[row] = tdv;
    if (a.prio) a.prio[row] = fabsf(tdv) + float(a.prio_eps);       // np.abs(f32) + 1e-6 (f32)
  }

  // ---- policy head: -E_q[z] and its logit gradient --------------------------------------
  if (a.pi_logits) {
    float qp[NT];
    row_softmax(a.pi_logits + ro, N, lane, qp);
    float ez = 0.f;
    float z[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      int k = lane + 32 * t;
      z[t] = (k < N) ? float(__dadd_rn(a.v_min, __dmul_rn(double(k), a.delta))) : 0.f;  // ddpg.py:47,238
      ez += qp[t] * z[t];
    }
    ez = warp_sum(ez);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      int k = lane + 32 * t;
      if (k < N && a.dlogits_pi) a.dlogits_pi[ro + k] = -a.grad_scale * qp[t] * (z[t] - ez);
    }
    if (lane == 0 && a.pi_rows) a.pi_rows[row] = -ez;
  }
}

int launch_heads(const HeadsArgs& a, int mode, cudaStream_t st) {
  dim3 grid(cdiv(a.B, HEAD_WARPS)), block(HEAD_WARPS * 32);
  D4PG_MAX_CARVEOUT((heads_kernel<0, 2>)); D4PG_MAX_CARVEOUT((heads_kernel<1, 2>));
  D4PG_MAX_CARVEOUT((heads_kernel<0, 4>)); D4PG_MAX_CARVEOUT((heads_kernel<1, 4>));
  // NT = atom slots per lane: 2 covers N<=64 (51 atoms), 4 covers N<=128 (101 atoms)
  if (a.N <= 64) {
    if (mode == 0) D4PG_CUDA_OK(launch_pdl(heads_kernel<0, 2>, grid, block, 0, st, a));
    else D4PG_CUDA_OK(launch_pdl(heads_kernel<1, 2>, grid, block, 0, st, a));
  } else {
    if (mode == 0) D4PG_CUDA_OK(launch_pdl(heads_kernel<0, 4>, grid, block, 0, st, a));
    else D4PG_CUDA_OK(launch_pdl(heads_kernel<1, 4>, grid, block, 0, st, a));
  }
  return D4PG_OK;
}

}  // namespace d4pg

extern "C" int32_t d4pg_proj_loss(const float* target_logits, const float* q_logits, const float* pi_logits,
                                  const double* rewards, const uint8_t* dones,
                                  int32_t B, int32_t N, double v_min, double v_max, double discount,
                                  int32_t proj_mode, int32_t flags, double prio_eps, float grad_scale,
                                  float* m, int32_t* bins_l, int32_t* bins_u,
                                  float* target_probs, float* q_probs,
                                  float* loss_rows, float* td, float* prio, float* dlogits_q,
                                  float* pi_rows, float* dlogits_pi, d4pg_stream_t stream) {
  using namespace d4pg;
  D4PG_REQUIRE(target_logits && q_logits && rewards && dones, D4PG_EINVAL, "d4pg_proj_loss: null input");
  D4PG_REQUIRE(B > 0 && N >= 2 && N <= D4PG_MAX_ATOMS, D4PG_EINVAL, "d4pg_proj_loss: need B>0, 2<=N<=%d (got B=%d N=%d)", D4PG_MAX_ATOMS, B, N);
  D4PG_REQUIRE(proj_mode == 0 || proj_mode == 1, D4PG_EINVAL, "d4pg_proj_loss: proj_mode must be 0 or 1");
  D4PG_REQUIRE((bins_l == nullptr) == (bins_u == nullptr), D4PG_EINVAL, "d4pg_proj_loss: bins_l/bins_u must both be set or both NULL");
  D4PG_REQUIRE(v_max > v_min, D4PG_EINVAL, "d4pg_proj_loss: v_max <= v_min");
  HeadsArgs a;
  a.target_logits = target_logits; a.q_logits = q_logits; a.pi_logits = pi_logits;
  a.rewards = rewards; a.dones = dones; a.B = B; a.N = N; a.flags = flags; a.ld = N;
  a.v_min = v_min; a.v_max = v_max;
  a.delta = (v_max - v_min) / double(N - 1);        // ddpg.py:46
  a.discount = discount; a.prio_eps = prio_eps; a.grad_scale = grad_scale;
  a.m = m; a.bins_l = bins_l; a.bins_u = bins_u; a.target_probs = target_probs; a.q_probs = q_probs;
  a.loss_rows = loss_rows; a.td = td; a.prio = prio; a.dlogits_q = dlogits_q;
  a.pi_rows = pi_rows; a.dlogits_pi = dlogits_pi;
  return launch_heads(a, proj_mode, as_stream(stream));
}
