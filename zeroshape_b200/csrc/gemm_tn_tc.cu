// tcgen05 "TN" GEMM for weight gradients:   C[N, Kc] (+)= A[M, N]^T * Bm[M, Kc]
//
//   A  = dY, fp32 row-major [M, N]                      (gradient of a layer's output, one row per sample / token / pixel)
//   Bm = the layer's input, fp32: a dense row-major matrix [M, Kc] (nn.Linear, 1x1 convolutions) or the im2col view of an
//        NHWC image (Kc = KH*KW*Cin, gathered on the fly -- the matrix is never materialised)
//   C  = dW in the parameter's own [N, Kc] layout (OHWI filters for convolutions)
// Replaces torch autograd's weight gradients of every nn.Linear / nn.Conv2d of the training step (reference
// model/shape_engine.py:248-277 `loss.backward()`), formerly zs_gemm_tn_f32 / zs_conv2d_nhwc_wgrad_f32 (FFMA).
//
// The reduction runs over the ROWS of both operands, i.e. over the slow index of both fp32 matrices, so both UMMA operands
// are "MN-major": element (n, m) of the A operand sits next to (n+1, m).  The shared-memory tiles use the canonical
// MN-major SWIZZLE_128B layout (64 MN-elements = 128 B per K-row, 8 K-rows per 1024-byte atom, atoms tiled MN-first) so a
// producer thread turns 8 consecutive fp32 of one row (two float4 loads, fully coalesced) into ONE 16-byte bf16 chunk per
// precision pass -- no transposition in registers or shared memory.  (Round-1 bring-up cross-checked these descriptors on B200
// against a K-major variant with transposing producers: both exact on integer operands; the transposing variant was dropped.)
//
// CTA = 672 threads: warps 0-15 producers (fp32 -> (hi, lo) bf16 tiles), warps 16-19 epilogue (TMEM -> red.global.add into
// C), warp 20 MMA issuer.  Work item = (128 x 256 output tile, split of the row range); split-K partial sums are combined with
// fp32 reductions in L2 (C is zeroed by the wrapper unless `accumulate`).  2 smem stages x 96 KB (64 reduced rows each), the
// producers fetch one stage ahead of their conversion (4 x 48 KB stages with two stages of look-ahead measured 20 % slower);
// 2 x 256 TMEM columns.
#include "common.cuh"
#include "tc_common.cuh"

namespace zs {
using namespace tc;

constexpr int TN_BM = 128, TN_BN = 256, TN_BK = 64;     // output tile 128 (n) x 256 (k-columns); 64 rows reduced per chunk
constexpr int TN_STAGES = 2;
constexpr int TN_A_TILE = TN_BM * TN_BK * 2;            // 16 KB (one of hi / lo)
constexpr int TN_B_TILE = TN_BN * TN_BK * 2;            // 32 KB
constexpr int TN_STAGE_BYTES = 2 * TN_A_TILE + 2 * TN_B_TILE;
constexpr int TN_PROD_WARPS = 16;                       // latency-bound fp32 loads: 4 producer warps per scheduler
constexpr int TN_PRODUCERS = TN_PROD_WARPS * 32;
constexpr int TN_MMA_WARP = TN_PROD_WARPS + 4;         // epilogue warps = TN_PROD_WARPS .. +3 (TMEM lane quarter = warp & 3)
constexpr int TN_THREADS = (TN_PROD_WARPS + 5) * 32;
constexpr int TN_SMEM = TN_STAGES * TN_STAGE_BYTES + 1024 + 256 + 2 * TN_BK * 16;     // + align slack, barriers, 2 row tables

struct TnParams {
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;
  int M, N, Kc;
  int n_tiles, k_tiles, splits, chunks_per_split, total_chunks;
  int precision;     // 0 = bf16x3, 1 = bf16
  int layout;        // 0 = MN-major operand tiles (the only layout; validated on B200 against K-major transposing producers)
  // BMODE 1: Bm = im2col of an NHWC image; row m = output pixel (b, oh, ow), column = (kh, kw, ci)
  int cB, cH, cW, cCin, cKH, cKW, cStride, cPadT, cPadL, cOH, cOW;
};

// MN-major SWIZZLE_128B operand descriptor: `lbo` = byte distance between 64-element atoms along M/N, `sbo` = byte distance
// between groups of 8 K-rows (cute/atom/mma_traits_sm100.hpp: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

// SPLIT: compile-time copy of `precision == 0` (single-pass bf16 skips the hi/lo residual arithmetic in the producers, which are
// bound by instruction issue: profiles/r1_ncu_train_gemm.md)
template <int BMODE, bool SPLIT>
__global__ void __launch_bounds__(TN_THREADS, 1) gemm_tn_tc_kernel(TnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + TN_STAGES * TN_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * TN_STAGES + 8u * s; };
  auto tfull_bar = [&](int a) { return bar_base + 16u * TN_STAGES + 8u * a; };
  auto tempty_bar = [&](int a) { return bar_base + 16u * TN_STAGES + 16u + 8u * a; };
  const uint32_t tmem_slot = bar_base + 16u * TN_STAGES + 32u;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + TN_STAGES * TN_STAGE_BYTES + 16 * TN_STAGES + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool split = SPLIT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TN_STAGES; ++s) {
      mbar_init(full_bar(s), TN_PROD_WARPS);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    fence_mbar_init();
  }
  if (warp == TN_MMA_WARP) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int total_work = p.n_tiles * p.k_tiles * p.splits;
  auto chunk_range = [&](int w, int& c0, int& c1) {
    const int sp = w % p.splits;
    c0 = sp * p.chunks_per_split;
    c1 = min(p.total_chunks, c0 + p.chunks_per_split);
  };

  if (warp < TN_PROD_WARPS) {
    // ================= producers =================
    // thread = (row r of the chunk, 16-byte bf16 chunk c of that row) of the MN-major tiles:
    //   A: 64 rows x 16 chunks (128 n-columns): c = t & 15, rows (t >> 4) + 32 i, i < 2
    //   B: 64 rows x 32 chunks (256 k-columns): c = t & 31, rows (t >> 5) + 16 i, i < 4
    // A warp reads whole 512-byte / 1-KB row segments and writes whole 128-byte swizzled smem rows (conflict-free).
    const int t = threadIdx.x;
    int stage = 0; uint32_t phase = 0;
    const bool vecA = ((p.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    const bool vecB = BMODE == 1 ? true : (((p.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0));
    const int ca = t & 15, ra = t >> 4;
    const int cb = t & 31, rb = t >> 5;
    constexpr int NA = TN_BK * 16 / TN_PRODUCERS, NB = TN_BK * 32 / TN_PRODUCERS;      // units per thread: 2, 4
    constexpr int SA = TN_PRODUCERS / 16, SB = TN_PRODUCERS / 32;                      // row strides: 32, 16
    struct Buf { float4 a[NA][2], b[NB][2]; };

    // 8 consecutive columns of row m of A / Bm (zero outside the matrix)
    auto load8A = [&](int m, int col, float4& x0, float4& x1) {
      x0 = x1 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m >= p.M || col >= p.N) return;
      const float* src = p.A + (int64_t)m * p.lda + col;
      if (vecA && col + 8 <= p.N) {
        x0 = __ldg(reinterpret_cast<const float4*>(src));
        x1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
      } else {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (col + e < p.N) ? __ldg(src + e) : 0.f;
        x0 = make_float4(v[0], v[1], v[2], v[3]); x1 = make_float4(v[4], v[5], v[6], v[7]);
      }
    };
    // BMODE 1: the thread's 8 columns lie inside one filter tap (Cin % 8 == 0) = 32 contiguous bytes of one input pixel.  The
    // output pixel of each of the chunk's 64 rows is decoded ONCE per chunk into a shared table {first pixel of the image, ih0, iw0}
    // (the per-unit div / mod chains were 1/3 of the producers' instructions and kept the XU pipe 22 % busy).
    int4* row_tab = reinterpret_cast<int4*>(smem_gen + TN_STAGES * TN_STAGE_BYTES + 256);
    int fetches = 0;
    int tap_kh = 0, tap_kw = 0, tap_ci = 0;
    auto load8B = [&](int r, int m, int col, float4& x0, float4& x1) {
      x0 = x1 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (BMODE == 1) {
        const int4 e = row_tab[(fetches & 1) * TN_BK + r];
        if (e.x < 0 || col >= p.Kc) return;
        const int ih = e.y + tap_kh, iw = e.z + tap_kw;
        if ((unsigned)ih >= (unsigned)p.cH || (unsigned)iw >= (unsigned)p.cW) return;
        const float4* src = reinterpret_cast<const float4*>(p.B + ((int64_t)e.x + ih * p.cW + iw) * p.cCin + tap_ci);
        x0 = __ldg(src); x1 = __ldg(src + 1);
        return;
      }
      if (m >= p.M || col >= p.Kc) return;
      const float* src = p.B + (int64_t)m * p.ldb + col;
      if (vecB && col + 8 <= p.Kc) {
        x0 = __ldg(reinterpret_cast<const float4*>(src));
        x1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
      } else {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (col + e < p.Kc) ? __ldg(src + e) : 0.f;
        x0 = make_float4(v[0], v[1], v[2], v[3]); x1 = make_float4(v[4], v[5], v[6], v[7]);
      }
    };
    auto put = [&](uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, const float4& x0, const float4& x1) {
      if (SPLIT) {
        uint4 hi, lo;
        split_bf16x2(x0.x, x0.y, hi.x, lo.x);
        split_bf16x2(x0.z, x0.w, hi.y, lo.y);
        split_bf16x2(x1.x, x1.y, hi.z, lo.z);
        split_bf16x2(x1.z, x1.w, hi.w, lo.w);
        *reinterpret_cast<uint4*>(hi_tile + off) = hi;
        *reinterpret_cast<uint4*>(lo_tile + off) = lo;
      } else {
        *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(pack_bf16x2(x0.x, x0.y), pack_bf16x2(x0.z, x0.w), pack_bf16x2(x1.x, x1.y),
                                                              pack_bf16x2(x1.z, x1.w));
      }
    };
    // fetch iterator over (work item, chunk): runs one chunk ahead of the conversion, so a chunk's L2 / DRAM round trip
    // overlaps the MMAs that still own the smem slot
    int fw = blockIdx.x, fc = 0, fc1 = 0;
    if (fw < total_work) chunk_range(fw, fc, fc1);
    auto fetch_next = [&](Buf& buf) {
      if (fw >= total_work) return;
      const int tile = fw / p.splits;
      const int nt = tile / p.k_tiles, kt = tile - nt * p.k_tiles;
      const int m0 = fc * TN_BK;
      const int colA = nt * TN_BM + ca * 8, colB = kt * TN_BN + cb * 8;
      if (BMODE == 1) {
        const int tap = colB / p.cCin;
        tap_ci = colB - tap * p.cCin;
        tap_kh = tap / p.cKW;
        tap_kw = tap - tap_kh * p.cKW;
        if (t < TN_BK) {
          const int m = m0 + t;
          int4 e = make_int4(-1, 0, 0, 0);
          if (m < p.M) {
            const int ow = m % p.cOW, tt = m / p.cOW;
            const int oh = tt % p.cOH, b = tt / p.cOH;
            e = make_int4(b * p.cH * p.cW, oh * p.cStride - p.cPadT, ow * p.cStride - p.cPadL, 0);
          }
          row_tab[(fetches & 1) * TN_BK + t] = e;
        }
        // producers only (named barrier 1); the table of the fetch before last is free again: every producer passed the
        // barrier of the last fetch after its final read of it
        asm volatile("bar.sync 1, %0;" ::"n"(TN_PRODUCERS) : "memory");
      }
#pragma unroll
      for (int i = 0; i < NA; ++i) load8A(m0 + ra + SA * i, colA, buf.a[i][0], buf.a[i][1]);
#pragma unroll
      for (int i = 0; i < NB; ++i) load8B(rb + SB * i, m0 + rb + SB * i, colB, buf.b[i][0], buf.b[i][1]);
      ++fetches;
      if (++fc == fc1) {
        fw += gridDim.x;
        if (fw < total_work) chunk_range(fw, fc, fc1);
      }
    };
    auto consume = [&](const Buf& buf) {
      mbar_wait(empty_bar(stage), phase ^ 1);
      uint8_t* a_hi = smem_gen + stage * TN_STAGE_BYTES;
      uint8_t* a_lo = a_hi + TN_A_TILE;
      uint8_t* b_hi = a_hi + 2 * TN_A_TILE;
      uint8_t* b_lo = b_hi + TN_B_TILE;
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        const int r = ra + SA * i;                      // K-row of the chunk
        const uint32_t off = (uint32_t)(r >> 3) * (2u * 1024u) + (uint32_t)(ca >> 3) * 1024u + (uint32_t)(r & 7) * 128u +
                             (uint32_t)(((ca & 7) ^ (r & 7)) << 4);
        put(a_hi, a_lo, off, buf.a[i][0], buf.a[i][1]);
      }
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int r = rb + SB * i;
        const uint32_t off = (uint32_t)(r >> 3) * (4u * 1024u) + (uint32_t)(cb >> 3) * 1024u + (uint32_t)(r & 7) * 128u +
                             (uint32_t)(((cb & 7) ^ (r & 7)) << 4);
        put(b_hi, b_lo, off, buf.b[i][0], buf.b[i][1]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(stage));      // one arrival per producer warp (its 32 stores are fenced and ordered)
      if (++stage == TN_STAGES) { stage = 0; phase ^= 1; }
    };
    Buf buf0;
    fetch_next(buf0);
    int64_t items = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) { int c0, c1; chunk_range(w, c0, c1); items += c1 - c0; }
    for (int64_t it = 0; it < items; ++it) {
      consume(buf0);
      fetch_next(buf0);
    }
  } else if (warp == TN_MMA_WARP) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint32_t idesc = umma_idesc_bf16(TN_BM, TN_BN) | (1u << 15) | (1u << 16);     // A and B MN-major
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int c0, c1;
        chunk_range(w, c0, c1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TN_BN;
        for (int c = c0; c < c1; ++c) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * TN_STAGE_BYTES;
          const uint32_t sb = sa + 2 * TN_A_TILE;
#pragma unroll
          for (int k = 0; k < TN_BK / 16; ++k) {
            // one MMA consumes 16 K-rows = two 8-row groups: A groups are 2 KB apart (2 atoms of 64 n), B groups 4 KB
            const uint32_t aoff = (uint32_t)k * 2u * 2048u, boff = (uint32_t)k * 2u * 4096u;
            const uint64_t a_hi = umma_desc_mn_sw128(sa + aoff, 1024u, 2048u), a_lo = umma_desc_mn_sw128(sa + TN_A_TILE + aoff, 1024u, 2048u);
            const uint64_t b_hi = umma_desc_mn_sw128(sb + boff, 1024u, 4096u), b_lo = umma_desc_mn_sw128(sb + TN_B_TILE + boff, 1024u, 4096u);
            umma_bf16(d_tmem, a_hi, b_hi, idesc, (c != c0 || k != 0) ? 1u : 0u);
            if (split) {
              umma_bf16(d_tmem, a_hi, b_lo, idesc, 1);
              umma_bf16(d_tmem, a_lo, b_hi, idesc, 1);
            }
          }
          umma_commit(empty_bar(stage));
          if (c == c1 - 1) umma_commit(tfull_bar(acc));
          if (++stage == TN_STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue: warps 16-19, warp q owns TMEM lanes 32q .. 32q+31 (rows n of the tile) =================
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int tile = w / p.splits;
      const int nt = tile / p.k_tiles, kt = tile - nt * p.k_tiles;
      const int n = nt * TN_BM + q * 32 + lane;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * TN_BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < TN_BN; cc += 32) {
        uint32_t rr[32];
        tmem_ld_32x32(taddr + cc, rr);
        tmem_ld_wait();
        const int col0 = kt * TN_BN + cc;
        if (n < p.N && col0 < p.Kc) {
          float* crow = p.C + (int64_t)n * p.ldc + col0;
          if (col0 + 32 <= p.Kc && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4(crow + j, __uint_as_float(rr[j]), __uint_as_float(rr[j + 1]), __uint_as_float(rr[j + 2]), __uint_as_float(rr[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.Kc) red_add(crow + j, __uint_as_float(rr[j]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TN_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int BMODE>
static int launch_tn(TnParams& p, int accumulate, cudaStream_t st, const char* what) {
  static thread_local bool configured = false;
  if (!configured) {
    ZS_CUDA_CALL(cudaFuncSetAttribute(gemm_tn_tc_kernel<BMODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM));
    ZS_CUDA_CALL(cudaFuncSetAttribute(gemm_tn_tc_kernel<BMODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM));
    configured = true;
  }
  p.n_tiles = (p.N + TN_BM - 1) / TN_BM;
  p.k_tiles = (p.Kc + TN_BN - 1) / TN_BN;
  p.total_chunks = (p.M + TN_BK - 1) / TN_BK;
  const int tiles = p.n_tiles * p.k_tiles;
  int splits = sm_count() / tiles;                         // at most one work item per SM: a second item on a few CTAs
  if (splits > p.total_chunks) splits = p.total_chunks;    // would double the kernel's duration (one wave, no tail)
  if (splits < 1) splits = 1;
  p.chunks_per_split = (p.total_chunks + splits - 1) / splits;
  p.splits = (p.total_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
  if (!accumulate) {
    if (p.ldc == p.Kc) ZS_CUDA_CALL(cudaMemsetAsync(p.C, 0, sizeof(float) * (size_t)p.N * p.Kc, st));
    else ZS_CUDA_CALL(cudaMemset2DAsync(p.C, sizeof(float) * (size_t)p.ldc, 0, sizeof(float) * (size_t)p.Kc, (size_t)p.N, st));
  }
  const int work = tiles * p.splits;
  const int grid = work < sm_count() ? work : sm_count();
  if (p.precision == 0) gemm_tn_tc_kernel<BMODE, true><<<grid, TN_THREADS, TN_SMEM, st>>>(p);
  else gemm_tn_tc_kernel<BMODE, false><<<grid, TN_THREADS, TN_SMEM, st>>>(p);
  ZS_CUDA_CHECK_LAUNCH(what);
  return ZS_OK;
}

}  // namespace zs

using namespace zs;

extern "C" int zs_gemm_tn_tc(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int64_t M, int N, int K,
                             int accumulate, int precision, int layout, void* stream) {
  ZS_REQUIRE(A && B && C, "zs_gemm_tn_tc: null pointer");
  ZS_REQUIRE(M > 0 && M < (1LL << 31) && N > 0 && K > 0 && lda >= N && ldb >= K && ldc >= K, "zs_gemm_tn_tc: bad shape");
  ZS_REQUIRE(precision == 0 || precision == 1, "zs_gemm_tn_tc: precision must be 0 (bf16x3) or 1 (bf16)");
  ZS_REQUIRE(layout == 0, "zs_gemm_tn_tc: layout must be 0 (MN-major operand tiles)");
  TnParams p{};
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
  p.M = (int)M; p.N = N; p.Kc = K; p.precision = precision; p.layout = layout;
  return launch_tn<0>(p, accumulate, as_stream(stream), "zs_gemm_tn_tc");
}

extern "C" int zs_conv2d_nhwc_wgrad_tc(const float* x, int B, int H, int W, int Cin, const float* dy, float* dw, int Cout, int KH,
                                       int KW, int stride, int pad_top, int pad_left, int OH, int OW, int accumulate,
                                       int precision, int layout, void* stream) {
  ZS_REQUIRE(x && dy && dw, "zs_conv2d_nhwc_wgrad_tc: null pointer");
  ZS_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && OH > 0 && OW > 0,
             "zs_conv2d_nhwc_wgrad_tc: bad shape");
  ZS_REQUIRE((Cin & 7) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
             "zs_conv2d_nhwc_wgrad_tc: needs Cin %% 8 == 0 and a 16-byte aligned image");
  ZS_REQUIRE(precision == 0 || precision == 1, "zs_conv2d_nhwc_wgrad_tc: precision must be 0 (bf16x3) or 1 (bf16)");
  ZS_REQUIRE(layout == 0, "zs_conv2d_nhwc_wgrad_tc: layout must be 0 (MN-major operand tiles)");
  const int64_t M64 = (int64_t)B * OH * OW;
  ZS_REQUIRE(M64 < (1LL << 31), "zs_conv2d_nhwc_wgrad_tc: too many output pixels");
  TnParams p{};
  p.A = dy; p.lda = Cout; p.B = x; p.ldb = 0; p.C = dw; p.ldc = KH * KW * Cin;
  p.M = (int)M64; p.N = Cout; p.Kc = KH * KW * Cin; p.precision = precision; p.layout = layout;
  p.cB = B; p.cH = H; p.cW = W; p.cCin = Cin; p.cKH = KH; p.cKW = KW; p.cStride = stride; p.cPadT = pad_top; p.cPadL = pad_left;
  p.cOH = OH; p.cOW = OW;
  return launch_tn<1>(p, accumulate, as_stream(stream), "zs_conv2d_nhwc_wgrad_tc");
}
