// Grouped tcgen05 GEMM over packed tile images (the tensor-core workhorse of the proposal network).
//
//   C[m, n] (op)= epi( sum_k A(m, k) * B(n, k) )      fp32-faithful 3xTF32 or single-pass TF32
//
// Every operand is a tile image (tc.cuh): 128-row x 32-column tiles, tf32 hi / lo parts.  An operand is read
//   * K-major  when the reduction runs along the image COLUMNS (image rows = M or N index)        [fmt K]
//   * MN-major when the reduction runs along the image ROWS    (image columns = M or N index)      [fmt MN]
// so forward GEMMs, input-gradient GEMMs and weight-gradient GEMMs all read the same row-major tensors
// without transposed copies (a tensor that is read both ways keeps one image per format).
// One pipeline stage = 32 reduction elements: a 16 KB tile (K-major) or 4 x 4 KB row pieces (MN-major).
// Warp roles (576 threads): warp 0 bulk-TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-17 epilogue
// (a lone warp per scheduler issues ~0.2 instr/clk on dependent code; 16 warps hide that latency).
// One 128 x 128 output tile per CTA.
#pragma once
#include "common.cuh"
#include "tc.cuh"

namespace tcg {

using namespace tc;

// explicit global-space accesses: pointers fetched from the on-chip descriptor are generic to the compiler
__device__ __forceinline__ void st_global(float* p, float v) { asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add_global(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ float ld_global(const float* p) { float v; asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }

enum : int {
  kRelu = 1,        // max(x, 0)
  kAccumulate = 2,  // fp32 output: C += result (exclusive owner)
  kMaskImg = 4,     // result = mask_img(m, n) > 0 ? result : 0   (mask image: K-format hi part, output geometry)
  kZeroInvalid = 8, // rows m >= m_valid produce zeros (padding rows of a segment)
};
// Split-K: when k_splits > 1 the reduction is divided over k_splits CTAs per output tile and the fp32 result
// is combined with atomicAdd (the caller zero-initialises / accumulates into C; no images, bias or mask).

struct Operand {
  const float* hi;
  const float* lo;
  const int* k_rows;  // MN-major only: image row origin of each 32-row reduction chunk (null: row0 + 32 c)
  int kb;             // column blocks of the image
  int mn;             // 0 = K-major, 1 = MN-major
  int row0, col0;     // K-major: (first M/N row [mult of 128], first reduction col [mult of 32])
                      // MN-major: (first reduction row [mult of 32], first M/N col [mult of 32])
};

struct Problem {
  Operand a, b;
  int M, N, K;        // logical dims (K = reduction length)
  int m_valid;        // rows >= m_valid are padding (kZeroInvalid)
  int flags;
  float* c;           // fp32 output or null
  int64_t ldc;
  const float* bias;  // [N] or null
  float* o_k_hi; float* o_k_lo;    // K-format output image (or null)
  float* o_mn_hi; float* o_mn_lo;  // MN-format output image (or null)
  const float* mask_hi;            // K-format image with the geometry of the output
  int o_kb;           // column blocks of the output / mask images
  int o_row0, o_col0; // origin of C(0,0) inside the output images (mult of 128 / 32)
  int tile_start, tiles_m, tiles_n;
  int k_splits;
};

constexpr int kStages = 3;
constexpr int kEpiWarps = 16;                     // 4 TMEM lane quadrants x 4 column chunks: one 32 x 32 block per warp
constexpr int kThreads = 64 + 32 * kEpiWarps;     // + bulk-TMA producer warp + MMA/TMEM warp
constexpr int kBN = 128;
constexpr int kTmemCols = 512;
constexpr int kPiece = 32 * 128;  // bytes: 32 rows x 128 B

struct __align__(1024) Smem {
  float a_hi[kStages][kTileFloats];
  float a_lo[kStages][kTileFloats];
  float b_hi[kStages][kTileFloats];
  float b_lo[kStages][kTileFloats];
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
  Problem prob;  // on-chip copy of the descriptor
};
// The epilogue warps transpose their blocks through the operand stages, which are idle once the last MMA has
// completed (16 warps x 32 x 33 floats = 66 KB <= the 96 KB of a_hi + a_lo).
static_assert(kEpiWarps * 32 * 33 * 4 <= 2 * kStages * kTileBytes, "transpose buffers must fit in the A stages");

__device__ __forceinline__ uint32_t stage_bytes(const Operand& o, int tile_idx) {
  if (!o.mn) return kTileBytes;
  int n = 0;
  for (int j = 0; j < 4; ++j) n += (o.col0 / 32 + tile_idx * 4 + j < o.kb);
  return (uint32_t)n * kPiece;
}

__device__ __forceinline__ void load_operand(const Operand& o, int tile_idx, int chunk, float* s_hi, float* s_lo,
                                             bool x3, uint64_t* bar) {
  if (!o.mn) {
    int64_t rt = o.row0 / 128 + tile_idx, cb = o.col0 / 32 + chunk;
    int64_t off = (rt * o.kb + cb) * kTileFloats;
    bulk_g2s(s_hi, o.hi + off, kTileBytes, bar);
    if (x3) bulk_g2s(s_lo, o.lo + off, kTileBytes, bar);
  } else {
    int r0 = o.k_rows ? o.k_rows[chunk] : o.row0 + 32 * chunk;
    int64_t rt = r0 >> 7, sub = (r0 & 127) >> 3;
    for (int j = 0; j < 4; ++j) {
      int cb = o.col0 / 32 + tile_idx * 4 + j;
      if (cb < o.kb) {
        int64_t off = (rt * o.kb + cb) * kTileFloats + sub * 256;
        bulk_g2s(s_hi + j * (kPiece / 4), o.hi + off, kPiece, bar);
        if (x3) bulk_g2s(s_lo + j * (kPiece / 4), o.lo + off, kPiece, bar);
      }
    }
  }
}

__device__ __forceinline__ uint64_t operand_desc(bool mn, uint32_t smem_addr, int ks) {
  return mn ? smem_desc_sw128_mn(smem_addr + ks * 1024, kPiece, 512) : smem_desc_sw128(smem_addr) + (uint64_t)(ks * 2);
}

// EPI selects the (compile-time) epilogue flavour so that the row loop is straight-line code:
//   0 = fp32 store (+bias, relu)   1 = fp32 reduction (red.add: weight gradients, split-K)   2 = tile images (+fp32)
template <bool X3, int EPI>
__global__ void __launch_bounds__(kThreads, 1) k_grouped(const Problem* __restrict__ probs, int n_probs,
                                                          unsigned long long* __restrict__ trace) {
#define TCG_TRACE(slot)                                                                                   \
  do {                                                                                                    \
    if (trace && blockIdx.x == 0) {                                                                       \
      unsigned long long _t;                                                                              \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                                              \
      trace[slot] = _t;                                                                                   \
    }                                                                                                     \
  } while (0)
  if (threadIdx.x == 0) TCG_TRACE(0);
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  int lo_i = 0, hi_i = n_probs - 1;
  while (lo_i < hi_i) {
    int mid = (lo_i + hi_i + 1) >> 1;
    if (probs[mid].tile_start <= tile) lo_i = mid; else hi_i = mid - 1;
  }
  for (int i = threadIdx.x; i < (int)(sizeof(Problem) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&sm.prob)[i] = reinterpret_cast<const uint32_t*>(probs + lo_i)[i];
  __syncthreads();
  const Problem& P = sm.prob;
  const int local = tile - P.tile_start;
  const int tiles_mn = P.tiles_m * P.tiles_n;
  const int split = local / tiles_mn, rem = local % tiles_mn;
  const int mt = rem / P.tiles_n, nt = rem % P.tiles_n;
  const int KC = (P.K + 31) / 32;
  const int nsplit = P.k_splits > 1 ? P.k_splits : 1;
  const int c0 = (int)((int64_t)KC * split / nsplit), c1 = (int)((int64_t)KC * (split + 1) / nsplit);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(&sm.tmem_base);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  if (threadIdx.x == 0) TCG_TRACE(1);
  // PDL (common.cuh): everything above touched only the host-uploaded descriptor table, shared memory and TMEM
  ppb_pdl_trigger();
  ppb_pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t bytes = (stage_bytes(P.a, mt) + stage_bytes(P.b, nt)) * (X3 ? 2u : 1u);
      for (int c = c0; c < c1; ++c) {
        int s = (c - c0) % kStages;
        uint32_t ph = ((c - c0) / kStages) & 1;
        mbar_wait(&sm.empty[s], ph ^ 1);
        mbar_expect_tx(&sm.full[s], bytes);
        load_operand(P.a, mt, c, sm.a_hi[s], sm.a_lo[s], X3, &sm.full[s]);
        load_operand(P.b, nt, c, sm.b_hi[s], sm.b_lo[s], X3, &sm.full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(128, kBN, P.a.mn, P.b.mn);
      const bool amn = P.a.mn != 0, bmn = P.b.mn != 0;
      for (int c = c0; c < c1; ++c) {
        int s = (c - c0) % kStages;
        uint32_t ph = ((c - c0) / kStages) & 1;
        mbar_wait(&sm.full[s], ph);
        if (c == c0) TCG_TRACE(2);
        fence_after_sync();
        uint32_t sa_hi = smem_u32(sm.a_hi[s]), sa_lo = smem_u32(sm.a_lo[s]);
        uint32_t sb_hi = smem_u32(sm.b_hi[s]), sb_lo = smem_u32(sm.b_lo[s]);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint64_t ah = operand_desc(amn, sa_hi, ks), bh = operand_desc(bmn, sb_hi, ks);
          if (X3) {
            uint64_t al = operand_desc(amn, sa_lo, ks), bl = operand_desc(bmn, sb_lo, ks);
            // three accumulators (see tc_gemm.cu): cross terms, hi*hi of even chunks, hi*hi of odd chunks
            mma_tf32(tmem + 2 * kBN, al, bh, idesc, (c == c0 && ks == 0) ? 0u : 1u);
            mma_tf32(tmem + 2 * kBN, ah, bl, idesc, 1u);
            mma_tf32(tmem + (c & 1) * kBN, ah, bh, idesc, (c - c0 < 2 && ks == 0) ? 0u : 1u);
          } else {
            mma_tf32(tmem, ah, bh, idesc, (c == c0 && ks == 0) ? 0u : 1u);
          }
        }
        mma_commit(&sm.empty[s]);
      }
      mma_commit(&sm.tmem_full);
      TCG_TRACE(3);
    }
  } else {
    // Epilogue.  TMEM hands each thread one ROW (32 consecutive columns per load); writing rows straight out would
    // scatter every store instruction over 32 cache lines (measured: 8.7 us per tile).  Each warp therefore
    // transposes its 32 x 32 block through shared memory and emits whole 128-byte row spans: lane = column.
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int cb = (warp - 2) >> 2;         // its 32-column chunk of the 128-column tile
    float (*stg)[33] = reinterpret_cast<float (*)[33]>(reinterpret_cast<float*>(sm.a_hi) + (warp - 2) * 32 * 33);
    mbar_wait(&sm.tmem_full, 0);
    if (threadIdx.x == 64) TCG_TRACE(4);
    fence_after_sync();
    const int m_base = mt * 128 + q * 32;  // first row of C handled by this warp
    // hoist everything that does not depend on the row out of the (fully unrolled) row loop
    const bool do_relu = (P.flags & kRelu) != 0, do_mask = (P.flags & kMaskImg) != 0;
    const int rows = (P.M - m_base) < 32 ? (P.M - m_base) : 32;
    const int valid_rows = (P.flags & kZeroInvalid) ? (P.m_valid - m_base) : 32;
    const float* bias_p = P.bias;
    const float* mask_p = P.mask_hi;
    float* c_p = P.c;
    float* ok_hi = P.o_k_hi; float* ok_lo = P.o_k_lo; float* omn_hi = P.o_mn_hi; float* omn_lo = P.o_mn_lo;
    const int64_t ldc = P.ldc, o_kb = P.o_kb, orow0 = (int64_t)P.o_row0 + m_base, ocb0 = P.o_col0 / 32 + nt * 4;
    {
      const int n0 = nt * kBN + cb * 32;
      if (n0 < ((P.N + 31) & ~31) && m_base < P.M) {  // warp-uniform
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (X3 ? (c0 & 1) * kBN : 0) + cb * 32, v);
      if (X3) {
        float u[32];
        if (c1 - c0 > 1) {
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ((c0 & 1) ^ 1) * kBN + cb * 32, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += u[j];
        }
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 2 * kBN + cb * 32, u);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += u[j];
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane][j] = v[j];
      __syncwarp();
      const int n = n0 + lane;
      const bool col_ok = n < P.N;
      const float bias = (bias_p && col_ok) ? __ldg(bias_p + n) : 0.0f;
      // the warp's 32 rows sit inside one 128-row image tile (m_base % 32 == 0, o_row0 % 128 == 0): row r of the block is
      // image row (orow0 + r), so (orow & 7) == (r & 7) and the span offset advances by 32 floats per row
      const int64_t span0 = ((orow0 >> 7) * o_kb + (ocb0 + cb)) * kTileFloats + (orow0 & 127) * 32;
      float* cp = c_p ? c_p + (int64_t)m_base * ldc + n : nullptr;
      const bool c_ok = cp != nullptr && col_ok;
      // rows in batches of eight: the ReLU-mask loads of a batch are issued together (one L2 round trip per batch instead of
      // one per row — the mask read used to serialise the whole epilogue: 10 us for a one-chunk GEMM)
#pragma unroll
      for (int rb = 0; rb < 32; rb += 8) {
        float mk[8];
        if (EPI == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = rb + j;
            const int64_t pos_k = span0 + r * 32 + ((((lane >> 2) ^ (r & 7))) << 2) + (lane & 3);
            // explicit loads into distinct registers, issued back to back (the compiler folded the predicated __ldg's into
            // one register and serialised them)
            mk[j] = 1.0f;
            if (do_mask && r < rows) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(mk[j]) : "l"(mask_p + pos_k));
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = rb + j;
          float x = stg[r][lane] + bias;
          x = do_relu ? fmaxf(x, 0.0f) : x;
          const bool live = r < rows;                        // row exists in C
          x = (col_ok && r < valid_rows) ? x : 0.0f;
          if (EPI == 0) {
            if (live && c_ok) st_global(cp + (int64_t)r * ldc, x);
          } else if (EPI == 1) {
            if (live && c_ok && x != 0.0f) red_add_global(cp + (int64_t)r * ldc, x);
          } else {
            const int64_t pos_k = span0 + r * 32 + ((((lane >> 2) ^ (r & 7))) << 2) + (lane & 3);
            if (do_mask && live) x = (mk[j] > 0.0f) ? x : 0.0f;
            if (live && c_ok) st_global(cp + (int64_t)r * ldc, x);
            float h, l;
            split_tf32(x, h, l);
            if (live && ok_hi) { st_global(ok_hi + pos_k, h); st_global(ok_lo + pos_k, l); }
            if (live && omn_hi) {
              const int64_t pos_mn = span0 + r * 32 + ((((lane >> 3) ^ (r & 3))) << 3) + (lane & 7);
              st_global(omn_hi + pos_mn, h);
              st_global(omn_lo + pos_mn, l);
            }
          }
        }
      }
      }
    }
  }
  if (threadIdx.x == 64) TCG_TRACE(5);
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) TCG_TRACE(6);
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem);
  }
  if (threadIdx.x == 0 && trace && blockIdx.x == 0) { trace[8] = (unsigned long long)P.M; trace[9] = (unsigned long long)P.N; trace[10] = (unsigned long long)P.K; trace[11] = (unsigned long long)gridDim.x; trace[12] = (unsigned long long)(c1 - c0); }
}

inline size_t smem_bytes() { return sizeof(Smem) + 1024; }

// image element writers for element-wise kernels (one value at logical (row, col) of an image with KB blocks)
__device__ __forceinline__ void img_store(float* k_hi, float* k_lo, float* mn_hi, float* mn_lo, int64_t row,
                                          int64_t col, int64_t KB, float x) {
  float h, l;
  split_tf32(x, h, l);
  if (k_hi) {
    int64_t o = packed_offset(row, col, KB);
    k_hi[o] = h;
    if (k_lo) k_lo[o] = l;
  }
  if (mn_hi) {
    int64_t o = packed_offset_mn(row, col, KB);
    mn_hi[o] = h;
    if (mn_lo) mn_lo[o] = l;
  }
}

}  // namespace tcg
