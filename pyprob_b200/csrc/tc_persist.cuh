// Persistent form of the grouped tcgen05 GEMM (tc_grouped.cuh): one CTA per SM walks output tiles blockIdx.x, + gridDim.x, ...
//
// Why: with one tile per CTA every tile pays the prologue (barrier init, TMEM allocation, descriptor fetch, first operand
// latency: ~2.5 us) and an epilogue during which the tensor pipe idles — 1-2.5 us for a plain fp32 store, 7-10 us for the
// image epilogue (five words per element at ~32 B/clk of SM store bandwidth).  For a 16-chunk tile (K = 512, 10 us of
// mainloop) that is half of the time.  Here the roles run free of each other across tiles:
//   producer   : streams operand chunks of tile after tile into a 2-stage ring (full/empty mbarriers, one running chunk counter)
//   MMA thread : waits until the accumulators are drained (tmem_empty), issues the tile's MMAs, commits tmem_full
//   epilogue   : drains TMEM into a DEDICATED staging buffer (16 warps x 32 x 33 floats), releases the accumulators
//                (tmem_empty) and only then runs the store loop — while the MMAs of the next tile are already running.
// 3xTF32 needs three accumulators (384 of the 512 TMEM columns), so there is no second accumulator set to ping-pong with;
// the early release after the drain is what overlaps epilogue and mainloop.  128 KB of operand stages (2 x 64 KB in 3xTF32,
// 4 x 32 KB single-pass) + 67.6 KB of staging.
// Same Problem descriptors, same arithmetic and the same epilogue flavours as tcg::k_grouped; results are identical.
#pragma once
#include "tc_grouped.cuh"

namespace tcp {

using namespace tc;

// 128 KB of operand stages either way: 3xTF32 moves 64 KB per 32-element chunk (A/B x hi/lo) -> 2 stages; single-pass TF32
// moves 32 KB -> 4 stages (with two, a CTA kept ~58 GB/s of loads in flight and the mainloop waited on latency)
template <bool X3> struct StageCfg { static constexpr int kStages = X3 ? 2 : 4; static constexpr int kFloats = (X3 ? 4 : 2) * kTileFloats; };

constexpr int kMaxStages = 4;
struct __align__(1024) Smem {
  float ring[8 * kTileFloats];            // 128 KB: stage s at s * kFloats: [a_hi | b_hi | a_lo | b_lo] (lo parts: 3xTF32 only)
  float stg[tcg::kEpiWarps][32][33];
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint32_t tmem_base;
};
inline size_t smem_bytes() { return sizeof(Smem) + 1024; }

// index of the problem that owns `tile` (tile_start is ascending)
__device__ __forceinline__ int find_problem(const tcg::Problem* __restrict__ probs, int n_probs, int tile) {
  int lo = 0, hi = n_probs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (probs[mid].tile_start <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}

struct TileGeom { int mt, nt, c0, c1; };
__device__ __forceinline__ TileGeom tile_geom(const tcg::Problem& P, int tile) {
  const int local = tile - P.tile_start;
  const int tiles_mn = P.tiles_m * P.tiles_n;
  const int split = local / tiles_mn, rem = local % tiles_mn;
  const int KC = (P.K + 31) / 32;
  const int nsplit = P.k_splits > 1 ? P.k_splits : 1;
  TileGeom g;
  g.mt = rem / P.tiles_n; g.nt = rem % P.tiles_n;
  g.c0 = (int)((int64_t)KC * split / nsplit); g.c1 = (int)((int64_t)KC * (split + 1) / nsplit);
  return g;
}

template <bool X3, int EPI>
__global__ void __launch_bounds__(tcg::kThreads, 1) k_grouped_persistent(const tcg::Problem* __restrict__ probs, int n_probs,
                                                                          int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kStages = StageCfg<X3>::kStages, kStageFloats = StageCfg<X3>::kFloats;
  auto st_a_hi = [&](int s) { return sm.ring + s * kStageFloats; };
  auto st_b_hi = [&](int s) { return sm.ring + s * kStageFloats + kTileFloats; };
  auto st_a_lo = [&](int s) { return sm.ring + s * kStageFloats + 2 * kTileFloats; };   // 3xTF32 only
  auto st_b_lo = [&](int s) { return sm.ring + s * kStageFloats + 3 * kTileFloats; };
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.tmem_full, 1);
    mbar_init(&sm.tmem_empty, tcg::kEpiWarps);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<tcg::kTmemCols>(&sm.tmem_base);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  ppb_pdl_trigger();
  ppb_pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t kc = 0;   // chunks issued so far (all tiles)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const tcg::Problem& P = probs[find_problem(probs, n_probs, tile)];
        const TileGeom g = tile_geom(P, tile);
        const tcg::Operand A = P.a, B = P.b;
        const uint32_t bytes = (tcg::stage_bytes(A, g.mt) + tcg::stage_bytes(B, g.nt)) * (X3 ? 2u : 1u);
        for (int c = g.c0; c < g.c1; ++c, ++kc) {
          const int s = kc % kStages;
          const uint32_t ph = (kc / kStages) & 1;
          mbar_wait(&sm.empty[s], ph ^ 1);
          mbar_expect_tx(&sm.full[s], bytes);
          tcg::load_operand(A, g.mt, c, st_a_hi(s), st_a_lo(s), X3, &sm.full[s]);
          tcg::load_operand(B, g.nt, c, st_b_hi(s), st_b_lo(s), X3, &sm.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t kc = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const tcg::Problem& P = probs[find_problem(probs, n_probs, tile)];
        const TileGeom g = tile_geom(P, tile);
        const bool amn = P.a.mn != 0, bmn = P.b.mn != 0;
        const uint32_t idesc = idesc_tf32(128, tcg::kBN, amn, bmn);
        // the accumulators are free once the epilogue warps have drained the previous tile (a fresh barrier passes parity 1)
        mbar_wait(&sm.tmem_empty, (uint32_t)(it & 1) ^ 1u);
        fence_after_sync();
        for (int c = g.c0; c < g.c1; ++c, ++kc) {
          const int s = kc % kStages;
          const uint32_t ph = (kc / kStages) & 1;
          mbar_wait(&sm.full[s], ph);
          fence_after_sync();
          const uint32_t sa_hi = smem_u32(st_a_hi(s)), sa_lo = smem_u32(st_a_lo(s));
          const uint32_t sb_hi = smem_u32(st_b_hi(s)), sb_lo = smem_u32(st_b_lo(s));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ah = tcg::operand_desc(amn, sa_hi, ks), bh = tcg::operand_desc(bmn, sb_hi, ks);
            if (X3) {
              const uint64_t al = tcg::operand_desc(amn, sa_lo, ks), bl = tcg::operand_desc(bmn, sb_lo, ks);
              mma_tf32(tmem + 2 * tcg::kBN, al, bh, idesc, (c == g.c0 && ks == 0) ? 0u : 1u);
              mma_tf32(tmem + 2 * tcg::kBN, ah, bl, idesc, 1u);
              mma_tf32(tmem + (c & 1) * tcg::kBN, ah, bh, idesc, (c - g.c0 < 2 && ks == 0) ? 0u : 1u);
            } else {
              mma_tf32(tmem, ah, bh, idesc, (c == g.c0 && ks == 0) ? 0u : 1u);
            }
          }
          mma_commit(&sm.empty[s]);
        }
        mma_commit(&sm.tmem_full);
      }
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int cb = (warp - 2) >> 2;         // its 32-column chunk of the 128-column tile
    float (*stg)[33] = sm.stg[warp - 2];
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int pi = 0;
      if (lane == 0) pi = find_problem(probs, n_probs, tile);
      pi = __shfl_sync(0xffffffffu, pi, 0);
      const tcg::Problem& P = probs[pi];
      const TileGeom g = tile_geom(P, tile);
      const int pM = P.M, pN = P.N, flags = P.flags;
      const int m_base = g.mt * 128 + q * 32;
      const bool do_relu = (flags & tcg::kRelu) != 0, do_mask = (flags & tcg::kMaskImg) != 0;
      const int rows = (pM - m_base) < 32 ? (pM - m_base) : 32;
      const int valid_rows = (flags & tcg::kZeroInvalid) ? (P.m_valid - m_base) : 32;
      const float* bias_p = P.bias;
      const float* mask_p = P.mask_hi;
      float* c_p = P.c;
      float* ok_hi = P.o_k_hi; float* ok_lo = P.o_k_lo; float* omn_hi = P.o_mn_hi; float* omn_lo = P.o_mn_lo;
      const int64_t ldc = P.ldc, o_kb = P.o_kb, orow0 = (int64_t)P.o_row0 + m_base, ocb0 = P.o_col0 / 32 + g.nt * 4;
      const int n0 = g.nt * tcg::kBN + cb * 32;
      const bool work = n0 < ((pN + 31) & ~31) && m_base < pM;   // warp-uniform
      mbar_wait(&sm.tmem_full, (uint32_t)(it & 1));
      fence_after_sync();
      float v[32];
      if (work) {
        if (g.c1 > g.c0) {
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (X3 ? (g.c0 & 1) * tcg::kBN : 0) + cb * 32, v);
          if (X3) {
            float u[32];
            if (g.c1 - g.c0 > 1) {
              tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ((g.c0 & 1) ^ 1) * tcg::kBN + cb * 32, u);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += u[j];
            }
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 2 * tcg::kBN + cb * 32, u);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += u[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.0f;
        }
      }
      // the accumulators are in registers: hand TMEM back to the MMA thread before the (long) store loop
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tmem_empty);
      if (!work) continue;
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane][j] = v[j];
      __syncwarp();
      const int n = n0 + lane;
      const bool col_ok = n < pN;
      const float bias = (bias_p && col_ok) ? __ldg(bias_p + n) : 0.0f;
      const int64_t span0 = ((orow0 >> 7) * o_kb + (ocb0 + cb)) * kTileFloats + (orow0 & 127) * 32;
      float* cp = c_p ? c_p + (int64_t)m_base * ldc + n : nullptr;
      const bool c_ok = cp != nullptr && col_ok;
#pragma unroll
      for (int rb = 0; rb < 32; rb += 8) {
        float mk[8];
        if (EPI == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = rb + j;
            const int64_t pos_k = span0 + r * 32 + ((((lane >> 2) ^ (r & 7))) << 2) + (lane & 3);
            mk[j] = 1.0f;
            if (do_mask && r < rows) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(mk[j]) : "l"(mask_p + pos_k));
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = rb + j;
          float x = stg[r][lane] + bias;
          x = do_relu ? fmaxf(x, 0.0f) : x;
          const bool live = r < rows;
          x = (col_ok && r < valid_rows) ? x : 0.0f;
          if (EPI == 0) {
            if (live && c_ok) tcg::st_global(cp + (int64_t)r * ldc, x);
          } else if (EPI == 1) {
            if (live && c_ok && x != 0.0f) tcg::red_add_global(cp + (int64_t)r * ldc, x);
          } else {
            const int64_t pos_k = span0 + r * 32 + ((((lane >> 2) ^ (r & 7))) << 2) + (lane & 3);
            if (do_mask && live) x = (mk[j] > 0.0f) ? x : 0.0f;
            if (live && c_ok) tcg::st_global(cp + (int64_t)r * ldc, x);
            float h, l;
            split_tf32(x, h, l);
            if (live && ok_hi) { tcg::st_global(ok_hi + pos_k, h); tcg::st_global(ok_lo + pos_k, l); }
            if (live && omn_hi) {
              const int64_t pos_mn = span0 + r * 32 + ((((lane >> 3) ^ (r & 3))) << 3) + (lane & 7);
              tcg::st_global(omn_hi + pos_mn, h);
              tcg::st_global(omn_lo + pos_mn, l);
            }
          }
        }
      }
      __syncwarp();   // the staging block is overwritten by this warp's next tile
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc<tcg::kTmemCols>(tmem);
  }
}

}  // namespace tcp
