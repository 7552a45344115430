// Tensor-core building blocks: operand packing + a tcgen05 GEMM over packed images.
//   C[M,N] = A[M,K] * B[N,K]^T (+bias) (relu), 3xTF32 (fp32-faithful) or single-pass TF32.
// Warp roles (192 threads): warp 0 = bulk-TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> global).  One 128x128 output tile per CTA.
#include "common.cuh"
#include "tc.cuh"

namespace {

using namespace tc;

// ---- pack: row-major fp32 -> swizzled tile image(s) ------------------------------------------------
template <bool MN>
__global__ void __launch_bounds__(256) k_pack(const float* __restrict__ X, int64_t rows, int64_t K, int64_t ldx,
                                               float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t KB = (K + kTileK - 1) / kTileK;
  const int64_t RT = (rows + kTileRows - 1) / kTileRows;
  const int64_t chunks = RT * kTileRows * KB * 8;  // 16-byte chunks in the padded image
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < chunks; q += (int64_t)gridDim.x * blockDim.x) {
    int64_t c = q & 7, t = q >> 3;
    int64_t row = t / KB, kb = t % KB;  // consecutive threads walk along K of one row: coalesced reads
    int64_t k0 = kb * kTileK + c * 4;
    float x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = (row < rows && k0 + j < K) ? __ldg(X + row * ldx + k0 + j) : 0.0f;
    float h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_tf32(x[j], h[j], l[j]);
    int64_t off = MN ? packed_offset_mn(row, k0, KB) : packed_offset(row, k0, KB);
    *reinterpret_cast<float4*>(hi + off) = make_float4(h[0], h[1], h[2], h[3]);
    if (lo) *reinterpret_cast<float4*>(lo + off) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// ---- GEMM --------------------------------------------------------------------------------------------
constexpr int kStages = 3;
constexpr int kGemmThreads = 192;
constexpr int kBN = 128;
// The tensor core adds each MMA result into the fp32 TMEM accumulator with truncation (measured: error
// grows linearly with the number of accumulating instructions, ~0.4 ulp each, biased toward zero).  To stay
// fp32-faithful the 3xTF32 path therefore spreads the work over three accumulators that the epilogue
// sums with round-to-nearest adds: hi*hi of even k-blocks, hi*hi of odd k-blocks, and the two small
// cross terms (lo*hi + hi*lo, 2^-11 of the main magnitude).
constexpr int kTmemCols = 512;
struct __align__(1024) GemmSmem {
  float a_hi[kStages][kTileFloats];
  float a_lo[kStages][kTileFloats];
  float b_hi[kStages][kTileFloats];
  float b_lo[kStages][kTileFloats];
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

template <bool X3>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm_packed(const float* __restrict__ A_hi, const float* __restrict__ A_lo, const float* __restrict__ B_hi,
              const float* __restrict__ B_lo, float* __restrict__ C, int M, int N, int K, int64_t ldc,
              const float* __restrict__ bias, int relu, unsigned long long* __restrict__ trace) {
#define PPB_TRACE(slot)                                                                                   \
  do {                                                                                                    \
    if (trace && blockIdx.x == 0 && blockIdx.y == 0) {                                                    \
      unsigned long long _t;                                                                              \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                                              \
      trace[slot] = _t;                                                                                   \
    }                                                                                                     \
  } while (0)
  extern __shared__ uint8_t smem_raw[];
  GemmSmem& sm = *reinterpret_cast<GemmSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = (K + kTileK - 1) / kTileK;
  const int mt = blockIdx.y, nt = blockIdx.x;
  if (threadIdx.x == 0) PPB_TRACE(0);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(&sm.tmem_base);
  if (threadIdx.x == 32) PPB_TRACE(1);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  if (threadIdx.x == 0) PPB_TRACE(2);

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t bytes = (X3 ? 4u : 2u) * kTileBytes;
      for (int kb = 0; kb < KB; ++kb) {
        int s = kb % kStages;
        uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&sm.empty[s], ph ^ 1);
        mbar_expect_tx(&sm.full[s], bytes);
        int64_t ao = ((int64_t)mt * KB + kb) * kTileFloats, bo = ((int64_t)nt * KB + kb) * kTileFloats;
        bulk_g2s(sm.a_hi[s], A_hi + ao, kTileBytes, &sm.full[s]);
        bulk_g2s(sm.b_hi[s], B_hi + bo, kTileBytes, &sm.full[s]);
        if (X3) {
          bulk_g2s(sm.a_lo[s], A_lo + ao, kTileBytes, &sm.full[s]);
          bulk_g2s(sm.b_lo[s], B_lo + bo, kTileBytes, &sm.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_tf32(128, kBN);
      for (int kb = 0; kb < KB; ++kb) {
        int s = kb % kStages;
        uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&sm.full[s], ph);
        if (kb == 0) PPB_TRACE(3);
        fence_after_sync();
        uint64_t ah = smem_desc_sw128(smem_u32(sm.a_hi[s])), bh = smem_desc_sw128(smem_u32(sm.b_hi[s]));
        uint64_t al = smem_desc_sw128(smem_u32(sm.a_lo[s])), bl = smem_desc_sw128(smem_u32(sm.b_lo[s]));
#pragma unroll
        for (int k = 0; k < kTileK / 8; ++k) {
          uint64_t adv = (uint64_t)(k * 8 * 4 >> 4);  // 32 bytes per UMMA_K step, encoded >> 4
          if (X3) {
            uint32_t first_lo = (kb == 0 && k == 0) ? 0u : 1u;
            uint32_t first_hi = (kb < 2 && k == 0) ? 0u : 1u;  // even / odd k-blocks own separate accumulators
            mma_tf32(tmem + 2 * kBN, al + adv, bh + adv, idesc, first_lo);
            mma_tf32(tmem + 2 * kBN, ah + adv, bl + adv, idesc, 1u);
            mma_tf32(tmem + (kb & 1) * kBN, ah + adv, bh + adv, idesc, first_hi);
          } else {
            mma_tf32(tmem, ah + adv, bh + adv, idesc, (kb == 0 && k == 0) ? 0u : 1u);
          }
        }
        mma_commit(&sm.empty[s]);  // frees the stage once these MMAs have read it
      }
      mma_commit(&sm.tmem_full);
      PPB_TRACE(4);
    }
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), +32)
    const int q = warp & 3;
    mbar_wait(&sm.tmem_full, 0);
    if (threadIdx.x == 64) PPB_TRACE(5);
    fence_after_sync();
    const int row = mt * 128 + q * 32 + lane;
#pragma unroll 1
    for (int cb = 0; cb < kBN / 32; ++cb) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + cb * 32, v);
      if (X3) {
        float u[32];
        if (KB > 1) {
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + kBN + cb * 32, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += u[j];
        }
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 2 * kBN + cb * 32, u);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += u[j];
      }
      int col0 = nt * kBN + cb * 32;
      if (relu & 8) {  // bring-up experiment: skip the stores except one value (keeps the loads alive)
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += v[j];
        if (row < M && cb == 0) C[(int64_t)row * ldc] = acc;
      } else if (row < M) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          int col = col0 + j;
          if (col < N) {
            float x = v[j] + (bias ? __ldg(bias + col) : 0.0f);
            if (relu & 1) x = fmaxf(x, 0.0f);
            C[(int64_t)row * ldc + col] = x;
          }
        }
      }
    }
  }
  if (threadIdx.x == 64) PPB_TRACE(6);
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) PPB_TRACE(7);
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem);
    if (lane == 0) PPB_TRACE(8);
  }
}

// ---- TN bring-up: C[M,N] = sum_r X[r, m] * Y[r, n] with X, Y given as row-major-packed images -------------------
// (both operands MN-major over the SAME tile format: no transposed copies).  One 128x128 tile per CTA;
// a stage holds 32 reduction rows: 4 column-blocks x 4 KB per operand image.
constexpr int kTnRows = 32;                     // reduction rows per stage
constexpr int kTnPiece = kTnRows * 128;         // bytes of one column-block piece (32 rows x 128 B)
struct __align__(1024) TnSmem {
  float a_hi[kStages][4 * kTnPiece / 4];
  float a_lo[kStages][4 * kTnPiece / 4];
  float b_hi[kStages][4 * kTnPiece / 4];
  float b_lo[kStages][4 * kTnPiece / 4];
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

template <bool X3>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm_packed_tn(const float* __restrict__ X_hi, const float* __restrict__ X_lo, const float* __restrict__ Y_hi,
                 const float* __restrict__ Y_lo, float* __restrict__ C, int M, int N, int R, int64_t ldc) {
  extern __shared__ uint8_t smem_raw[];
  TnSmem& sm = *reinterpret_cast<TnSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KBx = (M + kTileK - 1) / kTileK, KBy = (N + kTileK - 1) / kTileK;  // column blocks of the images
  const int RC = (R + kTnRows - 1) / kTnRows;                                  // reduction chunks
  const int mt = blockIdx.y, nt = blockIdx.x;

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
      for (int rc = 0; rc < RC; ++rc) {
        int s = rc % kStages;
        uint32_t ph = (rc / kStages) & 1;
        mbar_wait(&sm.empty[s], ph ^ 1);
        // zero-fill column blocks that do not exist in the image (ragged M / N tile edges)
        int na = 0, nb = 0;
        for (int j = 0; j < 4; ++j) { na += (mt * 4 + j < KBx); nb += (nt * 4 + j < KBy); }
        mbar_expect_tx(&sm.full[s], (uint32_t)((na + nb) * (X3 ? 2 : 1) * kTnPiece));
        int r0 = rc * kTnRows;
        int64_t rt = r0 >> 7, sub = (r0 & 127) >> 3;  // row tile, first atom inside it
        for (int j = 0; j < 4; ++j) {
          int cb = mt * 4 + j;
          if (cb < KBx) {
            int64_t off = (rt * KBx + cb) * kTileFloats + sub * 256;
            bulk_g2s(&sm.a_hi[s][j * kTnPiece / 4], X_hi + off, kTnPiece, &sm.full[s]);
            if (X3) bulk_g2s(&sm.a_lo[s][j * kTnPiece / 4], X_lo + off, kTnPiece, &sm.full[s]);
          }
          cb = nt * 4 + j;
          if (cb < KBy) {
            int64_t off = (rt * KBy + cb) * kTileFloats + sub * 256;
            bulk_g2s(&sm.b_hi[s][j * kTnPiece / 4], Y_hi + off, kTnPiece, &sm.full[s]);
            if (X3) bulk_g2s(&sm.b_lo[s][j * kTnPiece / 4], Y_lo + off, kTnPiece, &sm.full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_tf32(128, kBN, 1, 1);
      for (int rc = 0; rc < RC; ++rc) {
        int s = rc % kStages;
        uint32_t ph = (rc / kStages) & 1;
        mbar_wait(&sm.full[s], ph);
        fence_after_sync();
#pragma unroll
        for (int k = 0; k < kTnRows / 8; ++k) {
          uint32_t adv = k * 1024;  // next 8-row atom
          uint64_t ah = smem_desc_sw128_mn(smem_u32(sm.a_hi[s]) + adv, kTnPiece, 512);
          uint64_t al = smem_desc_sw128_mn(smem_u32(sm.a_lo[s]) + adv, kTnPiece, 512);
          uint64_t bh = smem_desc_sw128_mn(smem_u32(sm.b_hi[s]) + adv, kTnPiece, 512);
          uint64_t bl = smem_desc_sw128_mn(smem_u32(sm.b_lo[s]) + adv, kTnPiece, 512);
          if (X3) {
            uint32_t first_lo = (rc == 0 && k == 0) ? 0u : 1u;
            uint32_t first_hi = (rc < 2 && k == 0) ? 0u : 1u;
            mma_tf32(tmem + 2 * kBN, al, bh, idesc, first_lo);
            mma_tf32(tmem + 2 * kBN, ah, bl, idesc, 1u);
            mma_tf32(tmem + (rc & 1) * kBN, ah, bh, idesc, first_hi);
          } else {
            mma_tf32(tmem, ah, bh, idesc, (rc == 0 && k == 0) ? 0u : 1u);
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
    const int row = mt * 128 + q * 32 + lane;
#pragma unroll 1
    for (int cb = 0; cb < kBN / 32; ++cb) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + cb * 32, v);
      if (X3) {
        float u[32];
        if (RC > 1) {
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + kBN + cb * 32, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += u[j];
        }
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 2 * kBN + cb * 32, u);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += u[j];
      }
      int col0 = nt * kBN + cb * 32;
      if (row < M) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < N) C[(int64_t)row * ldc + col0 + j] = v[j];
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

}  // namespace

unsigned long long* g_trace = nullptr;  // bring-up only: device buffer of timestamps (ppb_debug_trace)
int g_trace_launch = 0;

extern "C" {

// bring-up instrumentation: when a 16-slot device buffer is set, CTA (0,0) of ppb_gemm_packed records globaltimer
// stamps at its phase boundaries (start, alloc, sync, first data, last commit, accumulators ready, stores done,
// final sync, dealloc)
int ppb_debug_trace(void* buf16_dev) { g_trace = (unsigned long long*)buf16_dev; g_trace_launch = 0; return PPB_OK; }

int64_t ppb_packed_floats(int64_t rows, int64_t K) {
  int64_t RT = (rows + tc::kTileRows - 1) / tc::kTileRows, KB = (K + tc::kTileK - 1) / tc::kTileK;
  return RT * KB * tc::kTileFloats;
}

int ppb_pack_tf32(const float* X, int64_t rows, int64_t K, int64_t ldx, float* hi_out, float* lo_out, void* stream) {
  PPB_CHECK_ARG(X && hi_out && rows > 0 && K > 0 && ldx >= K, "bad arguments");
  int64_t chunks = ppb_packed_floats(rows, K) / 4;
  k_pack<false><<<ppb_grid_for(chunks, 256, 1), 256, 0, (cudaStream_t)stream>>>(X, rows, K, ldx, hi_out, lo_out);
  PPB_LAUNCH_CHECK();
  return PPB_OK;
}

int ppb_pack_tf32_mn(const float* X, int64_t rows, int64_t K, int64_t ldx, float* hi_out, float* lo_out, void* stream) {
  PPB_CHECK_ARG(X && hi_out && rows > 0 && K > 0 && ldx >= K, "bad arguments");
  int64_t chunks = ppb_packed_floats(rows, K) / 4;
  k_pack<true><<<ppb_grid_for(chunks, 256, 1), 256, 0, (cudaStream_t)stream>>>(X, rows, K, ldx, hi_out, lo_out);
  PPB_LAUNCH_CHECK();
  return PPB_OK;
}

int ppb_gemm_packed(const float* A_hi, const float* A_lo, const float* B_hi, const float* B_lo, float* C, int64_t M,
                    int64_t N, int64_t K, int64_t ldc, const float* bias, int relu, int precision, void* stream) {
  PPB_CHECK_ARG(A_hi && B_hi && C && M > 0 && N > 0 && K > 0 && ldc >= N, "bad arguments");
  PPB_CHECK_ARG(precision == PPB_PREC_TF32X3 || precision == PPB_PREC_TF32, "precision must be TF32X3 or TF32");
  PPB_CHECK_ARG(precision == PPB_PREC_TF32 || (A_lo && B_lo), "3xTF32 needs the lo images");
  dim3 grid((unsigned)((N + kBN - 1) / kBN), (unsigned)((M + 127) / 128));
  size_t smem = sizeof(GemmSmem) + 1024;
  if (precision == PPB_PREC_TF32X3) {
    PPB_CUDA(cudaFuncSetAttribute(k_gemm_packed<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gemm_packed<true><<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(A_hi, A_lo, B_hi, B_lo, C, (int)M, (int)N,
                                                                            (int)K, ldc, bias, relu, g_trace);
  } else {
    PPB_CUDA(cudaFuncSetAttribute(k_gemm_packed<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gemm_packed<false><<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(A_hi, A_lo, B_hi, B_lo, C, (int)M, (int)N,
                                                                             (int)K, ldc, bias, relu, g_trace);
  }
  PPB_LAUNCH_CHECK();
  return PPB_OK;
}

int ppb_gemm_packed_tn(const float* X_hi, const float* X_lo, const float* Y_hi, const float* Y_lo, float* C, int64_t M,
                       int64_t N, int64_t R, int64_t ldc, int precision, void* stream) {
  PPB_CHECK_ARG(X_hi && Y_hi && C && M > 0 && N > 0 && R > 0 && ldc >= N, "bad arguments");
  PPB_CHECK_ARG(precision == PPB_PREC_TF32X3 || precision == PPB_PREC_TF32, "precision must be TF32X3 or TF32");
  PPB_CHECK_ARG(precision == PPB_PREC_TF32 || (X_lo && Y_lo), "3xTF32 needs the lo images");
  dim3 grid((unsigned)((N + kBN - 1) / kBN), (unsigned)((M + 127) / 128));
  size_t smem = sizeof(TnSmem) + 1024;
  if (precision == PPB_PREC_TF32X3) {
    PPB_CUDA(cudaFuncSetAttribute(k_gemm_packed_tn<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gemm_packed_tn<true><<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(X_hi, X_lo, Y_hi, Y_lo, C, (int)M, (int)N, (int)R, ldc);
  } else {
    PPB_CUDA(cudaFuncSetAttribute(k_gemm_packed_tn<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gemm_packed_tn<false><<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(X_hi, X_lo, Y_hi, Y_lo, C, (int)M, (int)N, (int)R, ldc);
  }
  PPB_LAUNCH_CHECK();
  return PPB_OK;
}

}  // extern "C"
