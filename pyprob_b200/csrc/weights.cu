// Importance-weight normalisation: fp64 log-sum-exp, ESS and normalised logits.
// Replaces pyprob/distributions/empirical.py:298-302 (Categorical(logits=log_weights.double())),
// :759-766 (ESS = 1/sum p^2) and pyprob/util.py:398-399.  Algorithmic bytes: 4 B read per particle per
// pass (two passes) + 8 B fp64 logits written.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kItems = 8;  // elements per thread per block pass

struct Tri {
  double m, s, s2;  // max, sum exp(w-m), sum exp(2(w-m))
};

__device__ __forceinline__ Tri tri_merge(Tri a, Tri b) {
  if (b.m == -INFINITY) return a;
  if (a.m == -INFINITY) return b;
  Tri r;
  r.m = fmax(a.m, b.m);
  double ea = exp(a.m - r.m), eb = exp(b.m - r.m);
  r.s = a.s * ea + b.s * eb;
  r.s2 = a.s2 * ea * ea + b.s2 * eb * eb;
  return r;
}

__device__ __forceinline__ Tri tri_block_reduce(Tri t) {
  __shared__ double sm[3][kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Tri u;
    u.m = __shfl_xor_sync(0xffffffffu, t.m, o);
    u.s = __shfl_xor_sync(0xffffffffu, t.s, o);
    u.s2 = __shfl_xor_sync(0xffffffffu, t.s2, o);
    t = tri_merge(t, u);
  }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sm[0][w] = t.m; sm[1][w] = t.s; sm[2][w] = t.s2; }
  __syncthreads();
  if (w == 0) {
    Tri u;
    if (l < kThreads / 32) { u.m = sm[0][l]; u.s = sm[1][l]; u.s2 = sm[2][l]; }
    else { u.m = -INFINITY; u.s = 0; u.s2 = 0; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Tri v;
      v.m = __shfl_xor_sync(0xffffffffu, u.m, o);
      v.s = __shfl_xor_sync(0xffffffffu, u.s, o);
      v.s2 = __shfl_xor_sync(0xffffffffu, u.s2, o);
      u = tri_merge(u, v);
    }
    t = u;
  }
  __syncthreads();
  return t;  // valid in warp 0
}

__global__ void __launch_bounds__(kThreads) k_partials(const float* __restrict__ w, int64_t n,
                                                        double* __restrict__ partials) {
  // grid-stride over tiles of kThreads*kItems elements; per tile: max in fp32 (exact), sums in fp64
  Tri t; t.m = -INFINITY; t.s = 0; t.s2 = 0;
  const int64_t tile = (int64_t)kThreads * kItems;
  for (int64_t base = (int64_t)blockIdx.x * tile; base < n; base += (int64_t)gridDim.x * tile) {
    float v[kItems];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      int64_t i = base + (int64_t)j * kThreads + threadIdx.x;
      v[j] = (i < n) ? __ldg(w + i) : -INFINITY;
      mx = fmaxf(mx, v[j]);
    }
    if (mx != -INFINITY) {
      Tri u; u.m = (double)mx; u.s = 0; u.s2 = 0;
#pragma unroll
      for (int j = 0; j < kItems; ++j) {
        double e = exp((double)v[j] - u.m);
        u.s += e;
        u.s2 += e * e;
      }
      t = tri_merge(t, u);
    }
  }
  t = tri_block_reduce(t);
  if (threadIdx.x == 0) {
    partials[3 * blockIdx.x + 0] = t.m;
    partials[3 * blockIdx.x + 1] = t.s;
    partials[3 * blockIdx.x + 2] = t.s2;
  }
}

__global__ void __launch_bounds__(kThreads) k_finalize(const float* __restrict__ w, int64_t n,
                                                        const double* __restrict__ partials, int npart,
                                                        double* __restrict__ stats, double* __restrict__ logits) {
  // every block re-combines the (small) partial list; block 0 publishes the statistics
  Tri t; t.m = -INFINITY; t.s = 0; t.s2 = 0;
  for (int p = threadIdx.x; p < npart; p += kThreads) {
    Tri u; u.m = partials[3 * p]; u.s = partials[3 * p + 1]; u.s2 = partials[3 * p + 2];
    t = tri_merge(t, u);
  }
  t = tri_block_reduce(t);
  __shared__ double s_lse;
  if (threadIdx.x == 0) {
    double lse = t.m + log(t.s);
    s_lse = lse;
    if (blockIdx.x == 0) {
      stats[0] = lse;
      stats[1] = (t.s * t.s) / t.s2;  // 1 / sum p^2, p = e^{w-m}/s
      stats[2] = t.m;
      stats[3] = t.s;
    }
  }
  __syncthreads();
  if (!logits) return;
  double lse = s_lse;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    logits[i] = (double)__ldg(w + i) - lse;
}

__global__ void __launch_bounds__(kThreads) k_cast(const double* __restrict__ acc, float* __restrict__ out,
                                                    uint8_t* __restrict__ invalid, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    double a = acc[i];
    out[i] = (float)a;
    if (invalid) invalid[i] = (isnan(a) || isinf(a)) ? 1 : 0;
  }
}

}  // namespace

extern "C" {

int ppb_weights_num_partials(int64_t n) {
  int64_t b = (n + (int64_t)kThreads * kItems - 1) / ((int64_t)kThreads * kItems);
  if (b < 1) b = 1;
  if (b > 4 * PPB_NUM_SMS) b = 4 * PPB_NUM_SMS;  // 4 resident CTAs per SM, grid-stride beyond that
  return (int)b;
}

int ppb_weights_cast(const double* acc, float* log_w_out, uint8_t* invalid_out, int64_t n, void* stream) {
  PPB_CHECK_ARG(n >= 0 && acc && log_w_out, "bad arguments");
  if (n == 0) return PPB_OK;
  k_cast<<<ppb_grid_for(n, kThreads, 1), kThreads, 0, (cudaStream_t)stream>>>(acc, log_w_out, invalid_out, n);
  PPB_LAUNCH_CHECK();
  return PPB_OK;
}

int ppb_weights_partials(const float* log_w, int64_t n, double* partials, void* stream) {
  PPB_CHECK_ARG(n > 0 && log_w && partials, "bad arguments");
  k_partials<<<ppb_weights_num_partials(n), kThreads, 0, (cudaStream_t)stream>>>(log_w, n, partials);
  PPB_LAUNCH_CHECK();
  return PPB_OK;
}

int ppb_weights_finalize(const float* log_w, int64_t n, const double* partials, int npartials, double* stats4,
                         double* logits_out, void* stream) {
  PPB_CHECK_ARG(n >= 0 && partials && npartials > 0 && stats4, "bad arguments");
  PPB_CHECK_ARG(!logits_out || log_w, "logits requested without weights");
  int grid = logits_out ? ppb_grid_for(n, kThreads, 1) : 1;
  k_finalize<<<grid, kThreads, 0, (cudaStream_t)stream>>>(log_w, n, partials, npartials, stats4, logits_out);
  PPB_LAUNCH_CHECK();
  return PPB_OK;
}

}  // extern "C"
