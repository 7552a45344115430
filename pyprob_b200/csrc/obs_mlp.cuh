// Fused observation-embedding MLP (pyprob/nn/inference_network.py:132-139 + embedding_feedforward.py:35-48):
// per-observable Linear+ReLU chains -> concat -> final Linear+ReLU chain, all layers in ONE kernel per direction.
// Used in the small-batch regime where a GEMM launch per layer is latency-bound (each layer is a few MFLOP);
// wide / large-batch cases fall back to the grouped GEMM path.
#pragma once
#include "common.cuh"

namespace obsmlp {

constexpr int TB = 4;        // traces per CTA: the chain is latency-bound, so favour CTAs in flight (64 at B=256)
constexpr int WMAX = 256;    // widest activation handled on chip
constexpr int LD = WMAX + 1;
constexpr int NC = 64;       // output columns per staged weight chunk
constexpr int kThreads = 256;

struct Bufs {  // global activation buffers (forward writes, backward reads) — same layout as the GEMM path
  float* obs_act[PPB_MAX_OBS][PPB_MAX_FF_LAYERS];
  float* obs_cat;
  float* fin_act[PPB_MAX_FF_LAYERS];
  float* obs_emb;
};

struct Net {
  int num_obs, obs_in_total, E;
  ppb_ff_desc obs_ff[PPB_MAX_OBS];
  ppb_ff_desc obs_final;
};

inline size_t smem_bytes() { return (size_t)(3 * TB * LD + WMAX * (NC + 1)) * sizeof(float); }

// out_s[r][n] = relu(b[n] + sum_k in_s[r][k] W[n][k]) for the CTA's TB rows; also streamed to global (ldg)
__device__ __forceinline__ void dense_fwd(const float (*in_s)[LD], int in_dim, const float* __restrict__ W,
                                          const float* __restrict__ b, int out_dim, float (*out_s)[LD], int out_col0,
                                          float* ws, float* gout, int64_t ldg, int gcol0, int r0, int B) {
  const int tid = threadIdx.x;
  for (int n0 = 0; n0 < out_dim; n0 += NC) {
    const int nc = out_dim - n0 < NC ? out_dim - n0 : NC;
    for (int idx = tid; idx < nc * in_dim; idx += kThreads) {  // k fastest: coalesced read of W rows
      int nn = idx / in_dim, k = idx - nn * in_dim;
      ws[k * (NC + 1) + nn] = __ldg(W + (int64_t)(n0 + nn) * in_dim + k);
    }
    __syncthreads();
    for (int o = tid; o < TB * NC; o += kThreads) {
      int r = o / NC, nn = o % NC;
      if (nn < nc) {
        float acc = __ldg(b + n0 + nn);
#pragma unroll 4
        for (int k = 0; k < in_dim; ++k) acc = fmaf(in_s[r][k], ws[k * (NC + 1) + nn], acc);
        acc = fmaxf(acc, 0.0f);
        out_s[r][out_col0 + n0 + nn] = acc;
        if (r0 + r < B) gout[(int64_t)(r0 + r) * ldg + gcol0 + n0 + nn] = acc;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads) k_fwd(Net net, const float* __restrict__ arena, const float* __restrict__ obs,
                                                   int B, Bufs bufs) {
  extern __shared__ float smem[];
  float (*cat)[LD] = reinterpret_cast<float (*)[LD]>(smem);
  float (*pa)[LD] = reinterpret_cast<float (*)[LD]>(smem + TB * LD);
  float (*pb)[LD] = reinterpret_cast<float (*)[LD]>(smem + 2 * TB * LD);
  float* ws = smem + 3 * TB * LD;
  const int r0 = blockIdx.x * TB, tid = threadIdx.x;
  int in_off = 0, out_off = 0;
  for (int j = 0; j < net.num_obs; ++j) {
    const ppb_ff_desc& ff = net.obs_ff[j];
    for (int idx = tid; idx < TB * ff.in_dim; idx += kThreads) {
      int r = idx / ff.in_dim, k = idx % ff.in_dim;
      pa[r][k] = (r0 + r < B) ? __ldg(obs + (int64_t)(r0 + r) * net.obs_in_total + in_off + k) : 0.0f;
    }
    __syncthreads();
    float (*cur)[LD] = pa;
    float (*nxt)[LD] = pb;
    for (int l = 0; l < ff.num_layers; ++l) {
      const ppb_linear_desc& L = ff.layers[l];
      const bool last = l == ff.num_layers - 1;
      if (last) dense_fwd(cur, L.in_dim, arena + L.w_off, arena + L.b_off, L.out_dim, cat, out_off, ws, bufs.obs_cat, net.E, out_off, r0, B);
      else dense_fwd(cur, L.in_dim, arena + L.w_off, arena + L.b_off, L.out_dim, nxt, 0, ws, bufs.obs_act[j][l], L.out_dim, 0, r0, B);
      float (*t)[LD] = cur; cur = nxt; nxt = t;
    }
    in_off += ff.in_dim;
    out_off += ff.out_dim;
  }
  float (*cur)[LD] = cat;
  float (*nxt)[LD] = pa;
  for (int l = 0; l < net.obs_final.num_layers; ++l) {
    const ppb_linear_desc& L = net.obs_final.layers[l];
    const bool last = l == net.obs_final.num_layers - 1;
    dense_fwd(cur, L.in_dim, arena + L.w_off, arena + L.b_off, L.out_dim, nxt, 0, ws, last ? bufs.obs_emb : bufs.fin_act[l],
              L.out_dim, 0, r0, B);
    cur = nxt;
    nxt = (nxt == pa) ? pb : pa;
  }
}

// One layer of the backward chain for the CTA's TB rows.
//   dy_s : d(loss)/d(pre-activation) of this layer  [TB][out]      (already ReLU-masked)
//   x    : the layer's input rows (global, ld)       -> staged in x_s [TB][in]
//   dW += dy^T x, db += colsum(dy) (atomics), dx_s[r][k] = (x[r][k] > 0) * sum_n dy[r][n] W[n][k]  (if want_dx)
__device__ __forceinline__ void dense_bwd(const float (*dy_s)[LD], int dy_col0, int out_dim, const float* __restrict__ x,
                                          int64_t ldx, int xcol0, int in_dim, const float* __restrict__ W,
                                          float* __restrict__ dW, float* __restrict__ db, float (*x_s)[LD],
                                          float (*dx_s)[LD], bool want_dx, int r0, int B) {
  const int tid = threadIdx.x;
  for (int idx = tid; idx < TB * in_dim; idx += kThreads) {
    int r = idx / in_dim, k = idx % in_dim;
    x_s[r][k] = (r0 + r < B) ? __ldg(x + (int64_t)(r0 + r) * ldx + xcol0 + k) : 0.0f;
  }
  __syncthreads();
  for (int idx = tid; idx < out_dim * in_dim; idx += kThreads) {  // weight gradient: one atomic per (n, k) per CTA
    int n = idx / in_dim, k = idx - n * in_dim;
    float s = 0.0f;
#pragma unroll
    for (int r = 0; r < TB; ++r) s = fmaf(dy_s[r][dy_col0 + n], x_s[r][k], s);
    if (s != 0.0f) atomicAdd(dW + idx, s);
  }
  for (int n = tid; n < out_dim; n += kThreads) {
    float s = 0.0f;
#pragma unroll
    for (int r = 0; r < TB; ++r) s += dy_s[r][dy_col0 + n];
    if (s != 0.0f) atomicAdd(db + n, s);
  }
  if (want_dx) {
    for (int idx = tid; idx < TB * in_dim; idx += kThreads) {  // lanes along k: coalesced reads of W rows
      int r = idx / in_dim, k = idx - r * in_dim;
      float s = 0.0f;
      for (int n = 0; n < out_dim; ++n) s = fmaf(dy_s[r][dy_col0 + n], __ldg(W + (int64_t)n * in_dim + k), s);
      dx_s[r][k] = x_s[r][k] > 0.0f ? s : 0.0f;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads) k_bwd(Net net, const float* __restrict__ arena, float* __restrict__ grad,
                                                   const float* __restrict__ obs, const float* __restrict__ d_obs_emb,
                                                   int B, Bufs bufs) {
  extern __shared__ float smem[];
  float (*da)[LD] = reinterpret_cast<float (*)[LD]>(smem);
  float (*dbuf)[LD] = reinterpret_cast<float (*)[LD]>(smem + TB * LD);
  float (*xs)[LD] = reinterpret_cast<float (*)[LD]>(smem + 2 * TB * LD);
  const int r0 = blockIdx.x * TB, tid = threadIdx.x, E = net.E;
  // gradient w.r.t. the pre-activation of the last final layer: mask by its (post-ReLU) output
  for (int idx = tid; idx < TB * E; idx += kThreads) {
    int r = idx / E, k = idx % E;
    float g = 0.0f;
    if (r0 + r < B) {
      int64_t o = (int64_t)(r0 + r) * E + k;
      g = bufs.obs_emb[o] > 0.0f ? d_obs_emb[o] : 0.0f;
    }
    da[r][k] = g;
  }
  __syncthreads();
  float (*cur)[LD] = da;
  float (*nxt)[LD] = dbuf;
  for (int l = net.obs_final.num_layers - 1; l >= 0; --l) {
    const ppb_linear_desc& L = net.obs_final.layers[l];
    const float* x = l == 0 ? bufs.obs_cat : bufs.fin_act[l - 1];
    dense_bwd(cur, 0, L.out_dim, x, L.in_dim, 0, L.in_dim, arena + L.w_off, grad + L.w_off, grad + L.b_off, xs, nxt, true, r0, B);
    float (*t)[LD] = cur; cur = nxt; nxt = t;
  }
  // cur now holds d(obs_cat pre-activations) [TB][E]; keep it and walk every observable's chain
  float (*dcat)[LD] = cur;
  float (*tmp)[LD] = nxt;
  // a third scratch is needed for ping-pong inside a chain: reuse xs' neighbour region is not available, so chains
  // alternate between `tmp` and the tail of the shared allocation
  float (*tmp2)[LD] = reinterpret_cast<float (*)[LD]>(smem + 3 * TB * LD);
  int in_off = 0, out_off = 0;
  for (int j = 0; j < net.num_obs; ++j) {
    const ppb_ff_desc& ff = net.obs_ff[j];
    const float (*dy)[LD] = dcat;
    int dy_col0 = out_off;
    float (*o1)[LD] = tmp;
    float (*o2)[LD] = tmp2;
    for (int l = ff.num_layers - 1; l >= 0; --l) {
      const ppb_linear_desc& L = ff.layers[l];
      const float* x = l == 0 ? obs : bufs.obs_act[j][l - 1];
      int64_t ldx = l == 0 ? net.obs_in_total : ff.layers[l - 1].out_dim;
      int xcol0 = l == 0 ? in_off : 0;
      dense_bwd(dy, dy_col0, L.out_dim, x, ldx, xcol0, L.in_dim, arena + L.w_off, grad + L.w_off, grad + L.b_off, xs, o1,
                l > 0, r0, B);
      dy = o1; dy_col0 = 0;
      float (*t)[LD] = o1; o1 = o2; o2 = t;
    }
    in_off += ff.in_dim;
    out_off += ff.out_dim;
  }
}

inline size_t smem_bytes_bwd() { return (size_t)(4 * TB * LD) * sizeof(float); }

}  // namespace obsmlp
