// LSTM time step with the cell update fused into the recurrent GEMM's epilogue.
//
//   pre[:, g*H + u] = h_{t-1} W_hh^T  (tcgen05, 3xTF32 or TF32)  + P_obs[trace] + P_step[segment] + smp_emb W_smp
//   i,f,o = sigmoid, g = tanh;  c_t = f c_{t-1} + i g;  h_t = o tanh(c_t)          (torch.nn.LSTM, gate order i,f,g,o)
//
// replaces, for t >= 1, the pair (tcg::k_grouped<X3,0> writing fp32 gate pre-activations, k_cell_fwd re-reading
// them): pyprob/nn/inference_network_lstm.py:186-188.  The trick is the weight layout: W_hh is packed with
// GATE-INTERLEAVED rows — packed row ub*128 + g*32 + j  <-  original row g*H + ub*32 + j — so that the 128 output
// columns of one tile are the four gates of the SAME 32 hidden units, and epilogue warp (q, cb) holds gate cb of rows
// 32q..32q+31.  Each warp finishes its gate (adds the projections, applies the activation, stores it for the
// backward pass, keeps it in its staging block); after a 128-thread named barrier per row quadrant the four warps
// each complete 8 rows of the cell: c, h, and the K-/MN-format tile images of h that the next step, the heads and
// the weight-gradient GEMMs read.
//
// Mainloop, pipeline, descriptors and the three-accumulator 3xTF32 scheme are those of tc_grouped.cuh.
//
// STATUS: written after the round-1 GPU budget was spent — compiles, never run.  Opt-in: PPB_FUSED_CELL=1 launches
// k_lstm_step once per time step, PPB_FUSED_CELL=2 runs all steps in one persistent launch (k_lstm_seq).
#pragma once
#include "tc_grouped.cuh"

namespace tcl {

using namespace tc;

struct CellIO {               // element-wise operands of the cell update, shared by all steps of a launch
  const float* p_obs;         // [traces, 4H]  observation projection, original gate-major columns
  const float* p_step;        // [steps, 4H]   step-embedding projection + both biases
  const float* smp_emb;       // [rows, S]     previous-sample embedding of every row
  const float* w_smp_t;       // [S, 4H]       W_ih columns of the sample embedding, transposed
  const int* row_trace;       // [rows] trace of a row, -1 for padding rows
  const int* row_step;        // [rows] step (segment) id of a row
  const int* row_prev;        // [rows] row of the same trace at step t-1
  float* gates;               // [rows, 4H] activated gates (i,f,g,o), gate-major columns — read by the backward pass
  float* c;                   // [rows, H]
  float* h;                   // [rows, H]
  float* hk_hi; float* hk_lo; float* hmn_hi; float* hmn_lo;   // tile images of h (both formats)
  int hkb;                    // column blocks of the h images (H / 32)
  int H, S;                   // hidden size (reduction length; N = 4H), sample-embedding width (<= 8)
  int no_b_prefetch;          // 1: the cluster kernel does not request W_hh tiles ahead of its PDL wait (A/B switch)
};

struct Step {                 // one (time step t >= 1, sub-batch) segment: one launch per time step
  tcg::Operand a;             // h image, rows of the segment at step t-1 (K-major, row0 = previous row origin)
  tcg::Operand b;             // gate-interleaved W_hh image (K-major, 4H rows, H columns)
  int M;                      // segment rows, padded to 128
  int row0;                   // first global row of the segment at step t (multiple of 128)
  int tile_start, tiles_m, tiles_n;
  CellIO io;
};

struct Seq {                  // one sub-batch, ALL its steps t = 1 .. T-1: persistent launch (k_lstm_seq)
  tcg::Operand a;             // h image (K-major); the row origin is set per step
  tcg::Operand b;             // gate-interleaved W_hh image
  int M;                      // rows of the sub-batch, padded to 128 (constant over its steps)
  int T;                      // trace length of the sub-batch
  int seg_off;                // offset of the sub-batch inside every step's row block
  int prog0;                  // index of its first progress counter (one counter per 128-row tile)
  int tile_start, tiles_m, tiles_n;
  const int* row_off;         // [T_max + 1] device: first row of step t
  int* progress;              // device counters, zero before the launch; [n_counters] = error flag
  int n_counters;
  CellIO io;
};

struct __align__(1024) Smem {
  float a_hi[tcg::kStages][kTileFloats];
  float a_lo[tcg::kStages][kTileFloats];
  float b_hi[tcg::kStages][kTileFloats];
  float b_lo[tcg::kStages][kTileFloats];
  uint64_t full[tcg::kStages];
  uint64_t empty[tcg::kStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
  union { Step step; Seq seq; };
};
static_assert(tcg::kEpiWarps * 32 * 33 * 4 <= 2 * tcg::kStages * kTileBytes, "staging blocks must fit in the A stages");

__device__ __forceinline__ void quad_barrier(int q) {  // the four epilogue warps that share TMEM lane quadrant q
  asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(128) : "memory");
}

// Epilogue of one 128-row x 32-unit tile: gate activations (phase 1), then the cell update (phase 2).
// `staging` = base of the idle operand stages; `tile_row0` = global row of the tile's first row.  Called by the 16
// epilogue warps only, after the accumulators of the step are complete.
template <bool X3>
__device__ __forceinline__ void cell_epilogue(float* staging, const CellIO& io, uint32_t tmem, int warp, int lane,
                                              int KC, int nt, int64_t tile_row0) {
  const int q = warp & 3;              // TMEM lane quadrant = 32-row block of the tile
  const int g = (warp - 2) >> 2;       // 32-column chunk of the tile = gate (i, f, g, o)
  const int qi = (warp - 2) & 3;       // position of this quadrant's warp inside every group of four staging blocks
  float (*stg)[33] = reinterpret_cast<float (*)[33]>(staging + (warp - 2) * 32 * 33);
  const int H = io.H, H4 = 4 * io.H, S = io.S;
  const int u = nt * 32 + lane;        // hidden unit owned by this lane
  const int col = g * H + u;           // its column in the gate-major [.., 4H] arrays
  float v[32];
  tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + g * 32, v);
  if (X3) {
    float w[32];
    if (KC > 1) {
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + tcg::kBN + g * 32, w);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += w[j];
    }
    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 2 * tcg::kBN + g * 32, w);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += w[j];
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; ++j) stg[lane][j] = v[j];   // thread = row  ->  lane = column
  __syncwarp();
  float wsmp[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) wsmp[s] = (s < S) ? __ldg(io.w_smp_t + (int64_t)s * H4 + col) : 0.0f;
  const int64_t row_base = tile_row0 + q * 32;
  // ---- phase 1: this warp's gate for its 32 rows ------------------------------------------------------------------
  for (int r = 0; r < 32; ++r) {
    const int64_t row = row_base + r;
    const int tr = __ldg(io.row_trace + row);
    float act = 0.0f;
    if (tr >= 0) {   // warp-uniform
      const int st = __ldg(io.row_step + row);
      // same order of additions as k_cell_fwd: (P_obs + P_step) + recurrent, then the sample-embedding FMAs
      float x = __ldg(io.p_obs + (int64_t)tr * H4 + col) + __ldg(io.p_step + (int64_t)st * H4 + col);
      x += stg[r][lane];
#pragma unroll
      for (int s = 0; s < 8; ++s)
        if (s < S) x = fmaf(__ldg(io.smp_emb + row * S + s), wsmp[s], x);
      act = (g == 2) ? ppb_cell_tanh(x) : ppb_cell_sigmoid(x);
    }
    tcg::st_global(io.gates + row * H4 + col, act);
    stg[r][lane] = act;
  }
  quad_barrier(q);
  // ---- phase 2: cell update, 8 rows per warp; the four gates come from the four staging blocks of the quadrant --------
  float (*sg_i)[33] = reinterpret_cast<float (*)[33]>(staging + (0 * 4 + qi) * 32 * 33);
  float (*sg_f)[33] = reinterpret_cast<float (*)[33]>(staging + (1 * 4 + qi) * 32 * 33);
  float (*sg_g)[33] = reinterpret_cast<float (*)[33]>(staging + (2 * 4 + qi) * 32 * 33);
  float (*sg_o)[33] = reinterpret_cast<float (*)[33]>(staging + (3 * 4 + qi) * 32 * 33);
  const int64_t hkb = io.hkb;
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    const int r = g * 8 + rr;
    const int64_t row = row_base + r;
    const int tr = __ldg(io.row_trace + row);
    float cn = 0.0f, hn = 0.0f;
    if (tr >= 0) {
      const int64_t rp = __ldg(io.row_prev + row);
      const float cp = tcg::ld_global(io.c + rp * H + u);
      cn = sg_f[r][lane] * cp + sg_i[r][lane] * sg_g[r][lane];
      hn = sg_o[r][lane] * ppb_cell_tanh(cn);
    }
    tcg::st_global(io.c + row * H + u, cn);
    tcg::st_global(io.h + row * H + u, hn);
    // image position of (row, u): tile (row / 128, nt), row span of 32 floats, swizzled chunk
    const int64_t span = ((row >> 7) * hkb + nt) * kTileFloats + (row & 127) * 32;
    float hh, hl;
    split_tf32(hn, hh, hl);
    const int64_t pos_k = span + ((((lane >> 2) ^ (int)(row & 7))) << 2) + (lane & 3);
    tcg::st_global(io.hk_hi + pos_k, hh);
    tcg::st_global(io.hk_lo + pos_k, hl);
    const int64_t pos_mn = span + ((((lane >> 3) ^ (int)(row & 3))) << 3) + (lane & 7);
    tcg::st_global(io.hmn_hi + pos_mn, hh);
    tcg::st_global(io.hmn_lo + pos_mn, hl);
  }
}

template <bool X3>
__global__ void __launch_bounds__(tcg::kThreads, 1) k_lstm_step(const Step* __restrict__ steps, int n_steps) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  int lo_i = 0, hi_i = n_steps - 1;
  while (lo_i < hi_i) {
    int mid = (lo_i + hi_i + 1) >> 1;
    if (steps[mid].tile_start <= tile) lo_i = mid; else hi_i = mid - 1;
  }
  for (int i = threadIdx.x; i < (int)(sizeof(Step) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&sm.step)[i] = reinterpret_cast<const uint32_t*>(steps + lo_i)[i];
  __syncthreads();
  const Step& P = sm.step;
  const int local = tile - P.tile_start;
  const int mt = local / P.tiles_n, nt = local % P.tiles_n;   // nt = block of 32 hidden units
  const int KC = (P.io.H + 31) / 32;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < tcg::kStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<tcg::kTmemCols>(&sm.tmem_base);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t bytes = (tcg::stage_bytes(P.a, mt) + tcg::stage_bytes(P.b, nt)) * (X3 ? 2u : 1u);
      for (int c = 0; c < KC; ++c) {
        int s = c % tcg::kStages;
        uint32_t ph = (c / tcg::kStages) & 1;
        mbar_wait(&sm.empty[s], ph ^ 1);
        mbar_expect_tx(&sm.full[s], bytes);
        tcg::load_operand(P.a, mt, c, sm.a_hi[s], sm.a_lo[s], X3, &sm.full[s]);
        tcg::load_operand(P.b, nt, c, sm.b_hi[s], sm.b_lo[s], X3, &sm.full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(128, tcg::kBN, 0, 0);
      for (int c = 0; c < KC; ++c) {
        int s = c % tcg::kStages;
        uint32_t ph = (c / tcg::kStages) & 1;
        mbar_wait(&sm.full[s], ph);
        fence_after_sync();
        uint32_t sa_hi = smem_u32(sm.a_hi[s]), sa_lo = smem_u32(sm.a_lo[s]);
        uint32_t sb_hi = smem_u32(sm.b_hi[s]), sb_lo = smem_u32(sm.b_lo[s]);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint64_t ah = tcg::operand_desc(false, sa_hi, ks), bh = tcg::operand_desc(false, sb_hi, ks);
          if (X3) {
            uint64_t al = tcg::operand_desc(false, sa_lo, ks), bl = tcg::operand_desc(false, sb_lo, ks);
            mma_tf32(tmem + 2 * tcg::kBN, al, bh, idesc, (c == 0 && ks == 0) ? 0u : 1u);
            mma_tf32(tmem + 2 * tcg::kBN, ah, bl, idesc, 1u);
            mma_tf32(tmem + (c & 1) * tcg::kBN, ah, bh, idesc, (c < 2 && ks == 0) ? 0u : 1u);
          } else {
            mma_tf32(tmem, ah, bh, idesc, (c == 0 && ks == 0) ? 0u : 1u);
          }
        }
        mma_commit(&sm.empty[s]);
      }
      mma_commit(&sm.tmem_full);
    }
  } else {
    mbar_wait(&sm.tmem_full, 0);
    fence_after_sync();
    cell_epilogue<X3>(reinterpret_cast<float*>(sm.a_hi), P.io, tmem, warp, lane, KC, nt, (int64_t)P.row0 + mt * 128);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc<tcg::kTmemCols>(tmem);
  }
}

// ---- persistent variant: ALL time steps t >= 1 of every sub-batch in one launch ---------------------------------------------
// Grid = (128-row tiles) x (32-unit blocks) of the sub-batches, at most one CTA per SM so that every CTA is resident.
// CTA (sub-batch, mt, nt) loops over t; step t needs h_{t-1} of ALL unit blocks of its row tile, so a launch
// boundary is replaced by one arrival counter per row tile: after its h-image stores of step t a CTA does
// __threadfence + atomicAdd; before the first bulk load of step t the producer thread spins (ld.acquire.gpu) until the
// counter shows (t-1) * tiles_n arrivals, then orders the async proxy after the acquire (fence.proxy.async) because
// the bulk copies read what other CTAs wrote with ordinary stores.  Waits give up after 2 s and raise the error flag.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <bool X3>
__global__ void __launch_bounds__(tcg::kThreads, 1) k_lstm_seq(const Seq* __restrict__ seqs, int n_seqs) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  int lo_i = 0, hi_i = n_seqs - 1;
  while (lo_i < hi_i) {
    int mid = (lo_i + hi_i + 1) >> 1;
    if (seqs[mid].tile_start <= tile) lo_i = mid; else hi_i = mid - 1;
  }
  for (int i = threadIdx.x; i < (int)(sizeof(Seq) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&sm.seq)[i] = reinterpret_cast<const uint32_t*>(seqs + lo_i)[i];
  __syncthreads();
  const Seq& P = sm.seq;
  const int local = tile - P.tile_start;
  const int mt = local / P.tiles_n, nt = local % P.tiles_n;
  const int KC = (P.io.H + 31) / 32;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < tcg::kStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<tcg::kTmemCols>(&sm.tmem_base);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  int* counter = P.progress + P.prog0 + mt;

  for (int t = 1; t < P.T; ++t) {
    const int row0 = __ldg(P.row_off + t) + P.seg_off, prev0 = __ldg(P.row_off + t - 1) + P.seg_off;
    const int cbase = (t - 1) * KC;     // the stage ring and its phases keep running across the steps
    if (warp == 0) {
      if (lane == 0) {
        if (t > 1) {
          const int target = (t - 1) * P.tiles_n;
          const unsigned long long t0 = timer_ns();
          while (ld_acquire_gpu(counter) < target) {
            if (timer_ns() - t0 > 2000000000ull) { P.progress[P.n_counters] = 1; break; }
          }
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        tcg::Operand a = P.a;
        a.row0 = prev0;
        const uint32_t bytes = (tcg::stage_bytes(a, mt) + tcg::stage_bytes(P.b, nt)) * (X3 ? 2u : 1u);
        for (int c = 0; c < KC; ++c) {
          const int cg = cbase + c, s = cg % tcg::kStages;
          const uint32_t ph = (cg / tcg::kStages) & 1;
          mbar_wait(&sm.empty[s], ph ^ 1);
          mbar_expect_tx(&sm.full[s], bytes);
          tcg::load_operand(a, mt, c, sm.a_hi[s], sm.a_lo[s], X3, &sm.full[s]);
          tcg::load_operand(P.b, nt, c, sm.b_hi[s], sm.b_lo[s], X3, &sm.full[s]);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = idesc_tf32(128, tcg::kBN, 0, 0);
        for (int c = 0; c < KC; ++c) {
          const int cg = cbase + c, s = cg % tcg::kStages;
          const uint32_t ph = (cg / tcg::kStages) & 1;
          mbar_wait(&sm.full[s], ph);
          fence_after_sync();
          uint32_t sa_hi = smem_u32(sm.a_hi[s]), sa_lo = smem_u32(sm.a_lo[s]);
          uint32_t sb_hi = smem_u32(sm.b_hi[s]), sb_lo = smem_u32(sm.b_lo[s]);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            uint64_t ah = tcg::operand_desc(false, sa_hi, ks), bh = tcg::operand_desc(false, sb_hi, ks);
            if (X3) {
              uint64_t al = tcg::operand_desc(false, sa_lo, ks), bl = tcg::operand_desc(false, sb_lo, ks);
              mma_tf32(tmem + 2 * tcg::kBN, al, bh, idesc, (c == 0 && ks == 0) ? 0u : 1u);
              mma_tf32(tmem + 2 * tcg::kBN, ah, bl, idesc, 1u);
              mma_tf32(tmem + (c & 1) * tcg::kBN, ah, bh, idesc, (c < 2 && ks == 0) ? 0u : 1u);
            } else {
              mma_tf32(tmem, ah, bh, idesc, (c == 0 && ks == 0) ? 0u : 1u);
            }
          }
          mma_commit(&sm.empty[s]);
        }
        mma_commit(&sm.tmem_full);
      }
    } else {
      mbar_wait(&sm.tmem_full, (uint32_t)((t - 1) & 1));
      fence_after_sync();
      cell_epilogue<X3>(reinterpret_cast<float*>(sm.a_hi), P.io, tmem, warp, lane, KC, nt, (int64_t)row0 + mt * 128);
      __threadfence();   // this thread's h / c / image stores are visible device-wide before the arrival below
    }
    // the accumulators are drained and the staging blocks (aliasing the operand stages) are free again; the next
    // step's bulk copies (async proxy) overwrite shared memory this step's epilogue wrote through the generic proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(counter, 1);
    }
  }
  if (warp == 1) tmem_dealloc<tcg::kTmemCols>(tmem);
}

inline size_t smem_bytes() { return sizeof(Smem) + 1024; }

// W_hh [4H, H] (row-major, gate-major rows) -> K-format hi / lo tile image with gate-interleaved rows
static __global__ void __launch_bounds__(256) k_pack_whh_interleaved(const float* __restrict__ w_hh, int H,
                                                               float* __restrict__ img_hi, float* __restrict__ img_lo) {
  const int kb = H / 32;
  const int rt = blockIdx.x;             // packed row tile = block of 32 hidden units
  for (int qi = threadIdx.x + blockIdx.y * blockDim.x; qi < 128 * kb * 8; qi += blockDim.x * gridDim.y) {
    const int c16 = qi & 7, t = qi >> 3;
    const int cb = t % kb, r = t / kb;   // r = packed row inside the tile = gate * 32 + j
    const int src_row = (r >> 5) * H + rt * 32 + (r & 31);
    const int k0 = cb * 32 + c16 * 4;
    float h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_tf32(__ldg(w_hh + (int64_t)src_row * H + k0 + j), h[j], l[j]);
    const int64_t o = ((int64_t)rt * kb + cb) * kTileFloats + r * 32 + ((c16 ^ (r & 7)) << 2);
    *reinterpret_cast<float4*>(img_hi + o) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(img_lo + o) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

}  // namespace tcl
