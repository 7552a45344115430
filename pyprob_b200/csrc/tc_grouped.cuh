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
// Warp roles (192 threads): warp 0 bulk-TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-5 epilogue.
// One 128 x 128 output tile per CTA.
#pragma once
#include "common.cuh"
#include "tc.cuh"

namespace tcg {

using namespace tc;

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
constexpr int kThreads = 192;
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

template <bool X3>
__global__ void __launch_bounds__(kThreads, 1) k_grouped(const Problem* __restrict__ probs, int n_probs) {
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
    }
  } else {
    const int q = warp & 3;
    mbar_wait(&sm.tmem_full, 0);
    fence_after_sync();
    const int m = mt * 128 + q * 32 + lane;  // row of C handled by this thread
    const bool row_ok = m < P.M;
    const bool row_valid = row_ok && (!(P.flags & kZeroInvalid) || m < P.m_valid);
#pragma unroll 1
    for (int cb = 0; cb < kBN / 32; ++cb) {
      const int n0 = nt * kBN + cb * 32;
      if (n0 >= ((P.N + 31) & ~31)) break;  // warp-uniform
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
      // image coordinates of this thread's 32-column span
      const int64_t orow = (int64_t)P.o_row0 + m, ocb = P.o_col0 / 32 + nt * 4 + cb;
      const int64_t tile_off = ((orow >> 7) * P.o_kb + ocb) * kTileFloats + (orow & 127) * 32;
      const int rr = (int)(orow & 7), r3 = (int)(orow & 3);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        int n = n0 + j;
        float x = v[j];
        if (P.bias && n < P.N) x += __ldg(P.bias + n);
        if (P.flags & kRelu) x = fmaxf(x, 0.0f);
        if (!row_valid || n >= P.N) x = 0.0f;
        v[j] = x;
      }
      if ((P.flags & kMaskImg) && row_ok) {
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16) {
          float4 mk = *reinterpret_cast<const float4*>(P.mask_hi + tile_off + ((c16 ^ rr) << 2));
          if (!(mk.x > 0.0f)) v[c16 * 4 + 0] = 0.0f;
          if (!(mk.y > 0.0f)) v[c16 * 4 + 1] = 0.0f;
          if (!(mk.z > 0.0f)) v[c16 * 4 + 2] = 0.0f;
          if (!(mk.w > 0.0f)) v[c16 * 4 + 3] = 0.0f;
        }
      }
      if (P.c && row_ok) {
        float* crow = P.c + (int64_t)m * P.ldc;
        if (nsplit > 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (n0 + j < P.N && v[j] != 0.0f) atomicAdd(crow + n0 + j, v[j]);
        } else if (P.flags & kAccumulate) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (n0 + j < P.N) crow[n0 + j] += v[j];
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (n0 + j < P.N) crow[n0 + j] = v[j];
        }
      }
      if ((P.o_k_hi || P.o_mn_hi) && row_ok) {
        float h[32], l[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) split_tf32(v[j], h[j], l[j]);
        if (P.o_k_hi) {
#pragma unroll
          for (int c16 = 0; c16 < 8; ++c16) {
            int64_t o = tile_off + ((c16 ^ rr) << 2);
            *reinterpret_cast<float4*>(P.o_k_hi + o) = make_float4(h[c16 * 4], h[c16 * 4 + 1], h[c16 * 4 + 2], h[c16 * 4 + 3]);
            if (P.o_k_lo) *reinterpret_cast<float4*>(P.o_k_lo + o) = make_float4(l[c16 * 4], l[c16 * 4 + 1], l[c16 * 4 + 2], l[c16 * 4 + 3]);
          }
        }
        if (P.o_mn_hi) {
#pragma unroll
          for (int c32 = 0; c32 < 4; ++c32) {
            int64_t o = tile_off + ((c32 ^ r3) << 3);
            *reinterpret_cast<float4*>(P.o_mn_hi + o) = make_float4(h[c32 * 8], h[c32 * 8 + 1], h[c32 * 8 + 2], h[c32 * 8 + 3]);
            *reinterpret_cast<float4*>(P.o_mn_hi + o + 4) = make_float4(h[c32 * 8 + 4], h[c32 * 8 + 5], h[c32 * 8 + 6], h[c32 * 8 + 7]);
            if (P.o_mn_lo) {
              *reinterpret_cast<float4*>(P.o_mn_lo + o) = make_float4(l[c32 * 8], l[c32 * 8 + 1], l[c32 * 8 + 2], l[c32 * 8 + 3]);
              *reinterpret_cast<float4*>(P.o_mn_lo + o + 4) = make_float4(l[c32 * 8 + 4], l[c32 * 8 + 5], l[c32 * 8 + 6], l[c32 * 8 + 7]);
            }
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem);
  }
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
