// Per-family log_prob over the particle axis — warp-coalesced, 128-bit vectorised, HBM-bound.
// Replaces pyprob/distributions/distribution.py:38-43 as driven per particle by pyprob/state.py.
// Algorithmic bytes per element: Normal/Uniform 16 B, Poisson 12 B, Categorical 4C+12 B,
// Mixture-Normal (3K+2)*4 B, Mixture-TruncatedNormal (3K+4)*4 B.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

// Streaming 128-bit load that bypasses L1 (each element is touched once).
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float ldg_stream(const float* p) {
  float r;
  asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

struct Param {
  const float* p;
  int stride;  // 0 = scalar broadcast, 1 = per particle
  __device__ __forceinline__ float at(int64_t i) const { return stride ? __ldg(p + i) : __ldg(p); }
  __device__ __forceinline__ void load4(int64_t i, float (&o)[4]) const {
    if (stride) {
      float4 v = ldg_stream4(p + i);
      o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    } else {
      float s = __ldg(p);
      o[0] = o[1] = o[2] = o[3] = s;
    }
  }
  bool vec_ok() const { return stride == 0 || ((((uintptr_t)p) & 15u) == 0); }
};

struct Sink {
  float* lp;
  double* acc;
  double scale;
  __device__ __forceinline__ void put(int64_t i, float v) const {
    if (lp) lp[i] = v;
    if (acc) acc[i] += scale * (double)v;
  }
  __device__ __forceinline__ void put4(int64_t i, const float (&v)[4]) const {
    if (lp) *reinterpret_cast<float4*>(lp + i) = make_float4(v[0], v[1], v[2], v[3]);
    if (acc) {
      double2 a0 = *reinterpret_cast<double2*>(acc + i);
      double2 a1 = *reinterpret_cast<double2*>(acc + i + 2);
      a0.x += scale * (double)v[0]; a0.y += scale * (double)v[1];
      a1.x += scale * (double)v[2]; a1.y += scale * (double)v[3];
      *reinterpret_cast<double2*>(acc + i) = a0;
      *reinterpret_cast<double2*>(acc + i + 2) = a1;
    }
  }
  bool vec_ok() const { return (!lp || ((((uintptr_t)lp) & 15u) == 0)) && (!acc || ((((uintptr_t)acc) & 15u) == 0)); }
};

// ---- two-parameter families ---------------------------------------------------------------------
// MUFU approximations (<= 2 ulp): the scoring kernels are bound by instruction issue, not HBM, as soon as they carry an IEEE
// division or a libm logf/expf (ncu: profiles/r02c_ncu_scoring.md); results stay within 1e-6 relative of the libm forms.
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

struct NormalOp {
  static constexpr bool kTable = false;
  // torch/distributions/normal.py log_prob: -((v - mu)^2) / (2 var) - log(sigma) - log(sqrt(2 pi))
  __device__ __forceinline__ float operator()(float v, float mu, float sigma, const float*) const {
    const float z = (v - mu) * fast_rcp(sigma);
    return fmaf(-0.5f * z, z, -fast_lg2(sigma) * PPB_LN2) - PPB_LOG_SQRT_2PI;
  }
};
struct UniformOp {
  // torch/distributions/uniform.py log_prob: log(lb*ub) - log(high-low), lb = low<=v, ub = high>v
  static constexpr bool kTable = false;
  __device__ __forceinline__ float operator()(float v, float lo, float hi, const float*) const {
    float inside = (lo <= v && hi > v) ? 0.0f : -INFINITY;
    return inside - logf(hi - lo);
  }
};
// log(k!) for the counts that actually occur (k < 64), correctly rounded from the double-precision lgamma: lgammaf costs
// ~40 instructions per particle and made the Poisson kernel ALU-bound at 39 % of HBM; other values take lgammaf.
// The table is copied to shared memory by every CTA: the lanes of a warp hold different counts, and a constant-bank read
// with divergent indices is replayed once per distinct address (63 % of HBM), shared memory serves them in one pass.
__constant__ float c_log_factorial[64];
struct PoissonOp {
  static constexpr bool kTable = true;
  // torch/distributions/poisson.py log_prob: xlogy(v, rate) - rate - lgamma(v+1)
  __device__ __forceinline__ float operator()(float v, float rate, float, const float* tab) const {
    float xl = (v == 0.0f) ? 0.0f : v * (fast_lg2(rate) * PPB_LN2);
    const int k = (int)v;
    const float lg = (v >= 0.0f && v < 64.0f && (float)k == v) ? tab[k] : lgammaf(v + 1.0f);
    return xl - rate - lg;
  }
};

template <class Op, bool VEC>
__global__ void __launch_bounds__(kThreads) k_score2(const float* __restrict__ value, Param a, Param b, Sink out,
                                                      int64_t n, Op op) {
  __shared__ float tab[Op::kTable ? 64 : 1];
  if (Op::kTable) {
    if (threadIdx.x < 64) tab[threadIdx.x] = c_log_factorial[threadIdx.x];
    __syncthreads();
  }
  int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t nth = (int64_t)gridDim.x * blockDim.x;
  if (VEC) {
    int64_t n4 = n >> 2;
    for (int64_t q = tid; q < n4; q += nth) {
      int64_t i = q << 2;
      float4 vv = ldg_stream4(value + i);
      float v[4] = {vv.x, vv.y, vv.z, vv.w};
      float pa[4], pb[4], r[4];
      a.load4(i, pa);
      b.load4(i, pb);
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = op(v[j], pa[j], pb[j], tab);
      out.put4(i, r);
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += nth) out.put(i, op(__ldg(value + i), a.at(i), b.at(i), tab));
  } else {
    for (int64_t i = tid; i < n; i += nth) out.put(i, op(__ldg(value + i), a.at(i), b.at(i), tab));
  }
}

template <class Op>
int launch_score2(const float* value, Param a, Param b, Sink out, int64_t n, void* stream, Op op) {
  if (n == 0) return PPB_OK;
  bool vec = aligned16(value) && a.vec_ok() && b.vec_ok() && out.vec_ok();
  cudaStream_t st = (cudaStream_t)stream;
  int grid = ppb_grid_for(n, kThreads, 4);
  if (vec)
    k_score2<Op, true><<<grid, kThreads, 0, st>>>(value, a, b, out, n, op);
  else
    k_score2<Op, false><<<grid, kThreads, 0, st>>>(value, a, b, out, n, op);
  PPB_LAUNCH_CHECK();
  return PPB_OK;
}

// ---- categorical --------------------------------------------------------------------------------
// log_prob = log(clamp(p[v] / sum(p))) (torch Categorical(probs=...): normalise, probs_to_logits clamps)
__global__ void __launch_bounds__(kThreads) k_categorical(const float* __restrict__ value,
                                                           const float* __restrict__ probs, int64_t row_stride,
                                                           int C, Sink out, int64_t n) {
  int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n; i += nth) {
    const float* p = probs + i * row_stride;
    int v = (int)__ldg(value + i);
    float s = 0.0f, pv = 0.0f;
    bool vec = (C % 4 == 0) && aligned16(p);
    if (vec) {
      for (int c = 0; c < C; c += 4) {
        float4 q = __ldg(reinterpret_cast<const float4*>(p + c));
        s += q.x; s += q.y; s += q.z; s += q.w;
        if (v >= c && v < c + 4) pv = (v == c) ? q.x : (v == c + 1) ? q.y : (v == c + 2) ? q.z : q.w;
      }
    } else {
      for (int c = 0; c < C; ++c) {
        float q = __ldg(p + c);
        s += q;
        if (c == v) pv = q;
      }
    }
    float lp = (v >= 0 && v < C) ? logf(ppb_clamp_prob(pv / s)) : NAN;
    out.put(i, lp);
  }
}

// ---- mixtures ------------------------------------------------------------------------------------
// Mixture.log_prob (pyprob/distributions/mixture.py:15-16, :38-45):
//   w = probs / sum(probs); lw = log(clamp(w)); lp = logsumexp_k(lw_k + lp_k(v))
// Arithmetic: with w_k = clamp(p_k / sum p) and z_k = (v - mu_k) / sigma_k,
//   logsumexp_k(log w_k + log N(v; mu_k, sigma_k)) = max_k(-z_k^2 / 2) + log sum_k (w_k / sigma_k) exp(-z_k^2 / 2 - max) - log sqrt(2 pi)
// (truncated components: sigma_k -> sigma_k Z_k, and -inf outside [low, high]): one exp and one reciprocal per component and
// ONE log per particle instead of two logs, an exp and two divisions per component — these kernels are bound by the
// transcendental/ALU rate, not by HBM (ncu: profiles/).  Same value as the reference's formula up to fp32 rounding.
// EXACT: K == KMAX is known at compile time (the component loops carry no k < K predicates)
template <int KMAX, bool TRUNC, bool EXACT = false>
__device__ __forceinline__ float mixture_row(float v, const float* __restrict__ m, const float* __restrict__ s,
                                             const float* __restrict__ p, int K_rt, float lo, float hi) {
  const int K = EXACT ? KMAX : K_rt;
  // Reciprocals, the exponentials and the final log use the hardware approximations (MUFU.RCP / EX2 / LG2: <= 2 ulp on the
  // terms that matter — exp arguments are <= 0 and the largest term is exp(0) = 1 exactly): with IEEE divisions and expf the
  // kernel issued 660 instructions per particle at K = 10 and was issue-bound at 0.72 of HBM (profiles/r02c_ncu_scoring.md).
  float pk[KMAX], a[KMAX], scale[KMAX];
  float psum = 0.0f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    pk[k] = (k < K) ? p[k] : 0.0f;
    psum += pk[k];
  }
  const float inv_psum = fast_rcp(psum);
  float mx = -INFINITY;
  bool bad = false;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K) {
      const float mu = m[k], sg = s[k];
      const float inv_sg = fast_rcp(sg);
      const float z = (v - mu) * inv_sg;
      float inv_norm = inv_sg;               // 1 / (sigma * truncated mass)
      bool neg = !(sg >= 0.0f);
      if (TRUNC) {
        const float alpha = (lo - mu) * inv_sg, beta = (hi - mu) * inv_sg;
        const float mass = ppb_std_normal_cdf(beta) - ppb_std_normal_cdf(alpha);
        inv_norm = inv_sg * fast_rcp(mass);
        neg = neg || !(mass >= 0.0f);
      }
      a[k] = (-0.5f * PPB_LOG2E) * z * z;   // exponent in base 2
      scale[k] = ppb_clamp_prob(pk[k] * inv_psum) * inv_norm;
      bad = bad || neg || (a[k] != a[k]) || (scale[k] != scale[k]);
      mx = fmaxf(mx, a[k]);
    }
  }
  if (bad) return NAN;                       // log of a negative / NaN parameter poisons the row like the reference
  if (TRUNC && !(v >= lo && v <= hi)) return -INFINITY;   // log(lb * ub) = -inf in every component
  if (mx == -INFINITY) return -INFINITY;     // torch.logsumexp of all -inf
  float acc = 0.0f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (k < K) acc = fmaf(scale[k], fast_ex2(a[k] - mx), acc);
  return (mx + fast_lg2(acc)) * PPB_LN2 - PPB_LOG_SQRT_2PI;
}

template <int KMAX, bool TRUNC, bool EXACT = false>
__global__ void __launch_bounds__(kThreads) k_mixture(const float* __restrict__ value,
                                                       const float* __restrict__ means,
                                                       const float* __restrict__ stddevs,
                                                       const float* __restrict__ probs, int64_t row_stride, int K,
                                                       Param low, Param high, Sink out, int64_t n) {
  int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n; i += nth) {
    float lo = 0.f, hi = 0.f;
    if (TRUNC) { lo = low.at(i); hi = high.at(i); }
    out.put(i, mixture_row<KMAX, TRUNC, EXACT>(__ldg(value + i), means + i * row_stride, stddevs + i * row_stride,
                                               probs + i * row_stride, K, lo, hi));
  }
}

// Same arithmetic, parameters staged through shared memory.  A thread of k_mixture reads its own [K] rows with 3K strided
// 4-byte loads: one warp-level load touches K different 128-byte lines, i.e. 3K * K load wavefronts per 32 particles (300 at
// K = 10) — the load/store unit, not HBM, bounds that kernel.  Here a WARP owns 32 consecutive particles: it copies their
// contiguous 32 x K parameter block with fully coalesced 4-byte loads (one wavefront each) into shared memory at an odd row
// pitch, then every lane reads its own row conflict-free.  Only __syncwarp, no CTA barrier.  Requires densely packed rows
// (row_stride == K).  Opt-in (PPB_MIXTURE_STAGED=1): results must equal k_mixture's to rounding
// (tests/test_scoring_staged_gpu.py).
template <int KMAX, bool TRUNC>
__global__ void __launch_bounds__(kThreads) k_mixture_staged(const float* __restrict__ value,
                                                              const float* __restrict__ means,
                                                              const float* __restrict__ stddevs,
                                                              const float* __restrict__ probs, int K, Param low,
                                                              Param high, Sink out, int64_t n) {
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pitch = K | 1;
  float* wm = sm + (size_t)warp * 3 * 32 * pitch;
  float* ws = wm + 32 * pitch;
  float* wp = ws + 32 * pitch;
  const int dr = 32 / K, dk = 32 % K;       // (row, component) of flat element idx + 32
  const int64_t stride = (int64_t)gridDim.x * (kThreads / 32) * 32;
  for (int64_t w0 = ((int64_t)blockIdx.x * (kThreads / 32) + warp) * 32; w0 < n; w0 += stride) {
    const int rows = (int)((n - w0 < 32) ? (n - w0) : 32);
    const int cnt = rows * K;
    const float* gm = means + w0 * K;
    const float* gs = stddevs + w0 * K;
    const float* gp = probs + w0 * K;
    // one array at a time: its K loads are issued together (independent requests in flight), then scattered into shared memory
    const int r0 = lane / K, k0 = lane % K;
    auto stage = [&](const float* __restrict__ g, float* __restrict__ w) {
      float v[KMAX];
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        const int idx = lane + 32 * j;
        v[j] = (j < K && idx < cnt) ? ldg_stream(g + idx) : 0.0f;
      }
      int r = r0, k = k0;
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        if (j < K && lane + 32 * j < cnt) w[r * pitch + k] = v[j];
        r += dr; k += dk;
        if (k >= K) { k -= K; ++r; }
      }
    };
    stage(gm, wm);
    stage(gs, ws);
    stage(gp, wp);
    __syncwarp();
    if (lane < rows) {
      const int64_t i = w0 + lane;
      float lo = 0.f, hi = 0.f;
      if (TRUNC) { lo = low.at(i); hi = high.at(i); }
      out.put(i, mixture_row<KMAX, TRUNC>(__ldg(value + i), wm + lane * pitch, ws + lane * pitch, wp + lane * pitch, K, lo, hi));
    }
    __syncwarp();
  }
}

inline bool mixture_staged_enabled() {
  static const bool on = [] {
    const char* e = getenv("PPB_MIXTURE_STAGED");
    return e && e[0] == '1';
  }();
  return on;
}

template <bool TRUNC>
int launch_mixture(const float* value, const float* means, const float* stddevs, const float* probs,
                   int64_t row_stride, int K, Param low, Param high, Sink out, int64_t n, void* stream) {
  if (n == 0) return PPB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = ppb_grid_for(n, kThreads, 1);
  if (mixture_staged_enabled() && row_stride == K && K <= 10) {
    size_t smem = (size_t)3 * kThreads * (K | 1) * sizeof(float);
    if (K <= 4)
      k_mixture_staged<4, TRUNC><<<grid, kThreads, smem, st>>>(value, means, stddevs, probs, K, low, high, out, n);
    else
      k_mixture_staged<10, TRUNC><<<grid, kThreads, smem, st>>>(value, means, stddevs, probs, K, low, high, out, n);
    PPB_LAUNCH_CHECK();
    return PPB_OK;
  }
  if (K <= 4)
    k_mixture<4, TRUNC><<<grid, kThreads, 0, st>>>(value, means, stddevs, probs, row_stride, K, low, high, out, n);
  else if (K == 10)   // the proposal heads' component count (pyprob/nn/proposal_normal_mixture.py: mixture_components = 10)
    k_mixture<10, TRUNC, true><<<grid, kThreads, 0, st>>>(value, means, stddevs, probs, row_stride, K, low, high, out, n);
  else if (K <= 10)
    k_mixture<10, TRUNC><<<grid, kThreads, 0, st>>>(value, means, stddevs, probs, row_stride, K, low, high, out, n);
  else if (K <= 32)
    k_mixture<32, TRUNC><<<grid, kThreads, 0, st>>>(value, means, stddevs, probs, row_stride, K, low, high, out, n);
  else {
    ppb_set_error("mixture log_prob: K=%d > 32 not supported", K);
    return PPB_ENOTSUP;
  }
  PPB_LAUNCH_CHECK();
  return PPB_OK;
}

}  // namespace

extern "C" {

int ppb_normal_log_prob(const float* value, const float* mean, int mean_stride, const float* stddev,
                        int stddev_stride, float* lp_out, double* acc, double acc_scale, int64_t n, void* stream) {
  if (n == 0) return PPB_OK;
  PPB_CHECK_ARG(n >= 0 && value && mean && stddev, "null pointer or negative n");
  PPB_CHECK_ARG((mean_stride | 1) == 1 && (stddev_stride | 1) == 1, "strides must be 0 or 1");
  return launch_score2(value, Param{mean, mean_stride}, Param{stddev, stddev_stride}, Sink{lp_out, acc, acc_scale}, n,
                       stream, NormalOp{});
}

int ppb_uniform_log_prob(const float* value, const float* low, int low_stride, const float* high, int high_stride,
                         float* lp_out, double* acc, double acc_scale, int64_t n, void* stream) {
  if (n == 0) return PPB_OK;
  PPB_CHECK_ARG(n >= 0 && value && low && high, "null pointer or negative n");
  PPB_CHECK_ARG((low_stride | 1) == 1 && (high_stride | 1) == 1, "strides must be 0 or 1");
  return launch_score2(value, Param{low, low_stride}, Param{high, high_stride}, Sink{lp_out, acc, acc_scale}, n, stream,
                       UniformOp{});
}

int ppb_poisson_log_prob(const float* value, const float* rate, int rate_stride, float* lp_out, double* acc,
                         double acc_scale, int64_t n, void* stream) {
  if (n == 0) return PPB_OK;
  PPB_CHECK_ARG(n >= 0 && value && rate, "null pointer or negative n");
  PPB_CHECK_ARG((rate_stride | 1) == 1, "strides must be 0 or 1");
  static bool table_ready = false;
  if (!table_ready) {
    float t[64];
    for (int k = 0; k < 64; ++k) t[k] = (float)lgamma((double)k + 1.0);
    PPB_CUDA(cudaMemcpyToSymbol(c_log_factorial, t, sizeof(t)));
    table_ready = true;
  }
  return launch_score2(value, Param{rate, rate_stride}, Param{rate, 0}, Sink{lp_out, acc, acc_scale}, n, stream,
                       PoissonOp{});
}

int ppb_categorical_log_prob(const float* value, const float* probs, int64_t probs_row_stride, int num_categories,
                             float* lp_out, double* acc, double acc_scale, int64_t n, void* stream) {
  if (n == 0) return PPB_OK;
  PPB_CHECK_ARG(n >= 0 && value && probs && num_categories > 0, "bad arguments");
  PPB_CHECK_ARG(probs_row_stride == 0 || probs_row_stride >= num_categories, "row stride < num_categories");
  if (n == 0) return PPB_OK;
  int grid = ppb_grid_for(n, kThreads, 1);
  k_categorical<<<grid, kThreads, 0, (cudaStream_t)stream>>>(value, probs, probs_row_stride, num_categories,
                                                             Sink{lp_out, acc, acc_scale}, n);
  PPB_LAUNCH_CHECK();
  return PPB_OK;
}

int ppb_mixture_normal_log_prob(const float* value, const float* means, const float* stddevs, const float* probs,
                                int64_t row_stride, int K, float* lp_out, double* acc, double acc_scale, int64_t n,
                                void* stream) {
  if (n == 0) return PPB_OK;
  PPB_CHECK_ARG(n >= 0 && value && means && stddevs && probs && K > 0, "bad arguments");
  return launch_mixture<false>(value, means, stddevs, probs, row_stride, K, Param{nullptr, 0}, Param{nullptr, 0},
                               Sink{lp_out, acc, acc_scale}, n, stream);
}

int ppb_mixture_truncated_normal_log_prob(const float* value, const float* means, const float* stddevs,
                                          const float* probs, int64_t row_stride, int K, const float* low,
                                          int low_stride, const float* high, int high_stride, float* lp_out,
                                          double* acc, double acc_scale, int64_t n, void* stream) {
  if (n == 0) return PPB_OK;
  PPB_CHECK_ARG(n >= 0 && value && means && stddevs && probs && low && high && K > 0, "bad arguments");
  return launch_mixture<true>(value, means, stddevs, probs, row_stride, K, Param{low, low_stride},
                              Param{high, high_stride}, Sink{lp_out, acc, acc_scale}, n, stream);
}

}  // extern "C"
