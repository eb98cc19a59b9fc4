// tcgen05 GEMM with split-bf16 operands (fp32-grade accuracy on the bf16 tensor pipe) and fused epilogue.
//
//   C[M,N] = epi( A[M,K] * W[N,K]^T ),  A fp32 row-major in global memory, W pre-packed (zs_gemm_tc_pack).
//
// Precision modes
//   0 "bf16x3": A = Ah + Al, W = Wh + Wl (bf16 each); D += Ah*Wh + Ah*Wl + Al*Wh in fp32 TMEM accumulators
//               -> ~2^-16 relative error per product (parity mode, meets the 1e-3 bar with margin)
//   1 "bf16"  : D += Ah*Wh only (fast mode, ~1e-2 relative on the decoder logits -- not parity)
//
// CTA = 576 threads, persistent over 128x256 output tiles, warp-specialised:
//   warps 0-7   A producers : fp32 rows -> (hi,lo) bf16 -> 128B-swizzled K-major smem tiles (generic proxy
//                              stores + fence.proxy.async), coalesced: thread = (row, 16-byte chunk)
//   warps 8-15  epilogue    : tcgen05.ld accumulator rows -> bias / activation / residual -> fp32 global
//   warp  16    MMA issuer  : one elected lane issues tcgen05.mma (M=128,N=256,K=16), commits to mbarriers
//   warp  17    W loader    : cp.async.bulk (TMA 1-D) of pre-swizzled weight tiles, complete_tx on mbarriers
// smem: 2 stages x (A_hi 16K + A_lo 16K + W_hi 32K + W_lo 32K) = 192 KB; TMEM: 2 x 256 fp32 columns
// (accumulator double buffer: the epilogue of tile i overlaps the MMAs of tile i+1).
#include "common.cuh"
#include "tc_common.cuh"

namespace zs {
using namespace tc;

constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 64;
constexpr int TC_STAGES = 2;
constexpr int TC_A_TILE = TC_BM * TC_BK * 2;   // 16 KB (one of hi/lo)
constexpr int TC_B_TILE = TC_BN * TC_BK * 2;   // 32 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_TILE + 2 * TC_B_TILE;  // 96 KB
constexpr int TC_PROD_WARPS = 8;                                     // warps 0-7 A producers, 8-15 epilogue, 16 MMA issuer, 17 W loader
constexpr int TC_MMA_WARP = TC_PROD_WARPS + 8, TC_W_WARP = TC_PROD_WARPS + 9;
constexpr int TC_THREADS = (TC_PROD_WARPS + 10) * 32;
constexpr int TC_SMEM = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 1024 /*softmax exchange*/;

struct TcParams {
  const float* A; int lda;
  const uint8_t* Wp;          // packed: [n_tiles][k_chunks][hi 32K | lo 32K]
  const float* bias;
  const float* res; int ldres; int res_mode;
  float* C; int ldc;
  int M, N, K;
  int act;
  int precision;
  int m_tiles, n_tiles, k_chunks;
  // grouped mode (attention): n-tile `nt` reads its A columns from nt*a_nt_off and writes its C columns at
  // nt*c_nt_off; only n_tile_valid (<= 256) columns of a tile are real and the MMA runs with N = mma_n.
  int a_nt_off, c_nt_off, n_tile_valid, mma_n;
  uint32_t w_tile_bytes;       // bytes of one (hi or lo) weight tile actually needed: mma_n rows x 128 B
  int epi_mode;                // 0: bias / activation / residual;  1: attention softmax (tc_epilogue_softmax)
  const float* qkv; int ld_qkv; float* R; float* R_inv; float scale; int n_keys;
  const float* rowscale;       // optional [M, n_tiles]: accumulator row scale applied before bias/residual (P.V normalisation)
  // implicit-GEMM convolution (AMODE 1): A row m = output pixel (b, oh, ow) of an NHWC image, K index = (kh, kw, ci)
  int cB, cH, cW, cCin, cKH, cKW, cStride, cPadT, cPadL, cOH, cOW, cPreRelu;
  // split-K (few-tile layers, see tc_launch): work item = (tile, split s), split s covers K-chunks [s k_chunks / k_splits, (s+1) ...)
  // and writes its raw fp32 partial tile to ws[s][M][N]; splitk_finalize_kernel sums them in a fixed order (deterministic)
  int k_splits; float* ws;
};

template <int ACT>
__device__ __forceinline__ float act_ct(float x) {
  if (ACT == ZS_ACT_RELU) return fmaxf(x, 0.0f);
  if (ACT == ZS_ACT_GELU) return act_gelu_erf(x);
  if (ACT == ZS_ACT_SOFTPLUS100) return act_softplus100(x);
  if (ACT == ZS_ACT_SIGMOID) return act_sigmoid(x);
  if (ACT == ZS_ACT_CLAMP01) return fminf(fmaxf(x, 0.0f), 1.0f);
  return x;
}

// one epilogue thread: its row, the 128 accumulator columns [hsel*128, hsel*128+128) of n-tile `nt`, in 4 pieces of 32
template <int ACT>
__device__ __forceinline__ void tc_epilogue_half(const TcParams& p, uint32_t taddr, int m, int nt, int hsel) {
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    if (hsel * 128 + c0 >= p.mma_n) break;                // warp-uniform: nothing was accumulated beyond the MMA width
    uint32_t rr[32];
    tmem_ld_32x32(taddr + c0, rr);
    tmem_ld_wait();
    const int lc = hsel * 128 + c0;                       // column inside the tile
    const int n0 = nt * p.c_nt_off + lc;                  // column in C / bias / residual
    int nvalid = p.n_tile_valid - lc;
    if (p.N - n0 < nvalid) nvalid = p.N - n0;
    if (m < p.M && nvalid > 0) {
      float* crow = p.C + (int64_t)m * p.ldc + n0;
      const float* rrow = p.res_mode != ZS_RES_NONE ? p.res + (int64_t)m * p.ldres + n0 : nullptr;
      const bool full = nvalid >= 32;
      const bool vec_c = full && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0);
      const bool vec_r = rrow && full && ((reinterpret_cast<uintptr_t>(rrow) & 15) == 0);
      float v[32];
      if (p.rowscale) {
        const float rs = __ldg(p.rowscale + (int64_t)m * p.n_tiles + nt);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]) * rs;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
      }
      if (p.bias) {
        if (full && ((reinterpret_cast<uintptr_t>(p.bias + n0) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j);
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += (j < nvalid) ? __ldg(p.bias + n0 + j) : 0.f;
        }
      }
      float rv[32];
      if (rrow) {
        if (vec_r) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 x = __ldg(reinterpret_cast<const float4*>(rrow) + j);
            rv[4 * j] = x.x; rv[4 * j + 1] = x.y; rv[4 * j + 2] = x.z; rv[4 * j + 3] = x.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) rv[j] = (j < nvalid) ? __ldg(rrow + j) : 0.f;
        }
        if (p.res_mode == ZS_RES_BEFORE_ACT) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = act_ct<ACT>(v[j] + rv[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = act_ct<ACT>(v[j]) + rv[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = act_ct<ACT>(v[j]);
      }
      if (vec_c) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          reinterpret_cast<float4*>(crow)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) crow[j] = v[j];
      }
    }
  }
}

// Attention scores epilogue (model/shape/implicit.py:38-57).  The accumulator row holds the raw scores q_h . k_j of
// one query point against the n_keys latent keys of head `nt`; the softmax also runs over the point's own key
// (s_self = q_h . k_p,h).  The two threads that share a row (column halves) exchange max / sum through `xch`.
// Writes normalised probabilities P[m, nt*c_nt_off + j] (0 for the padding columns) and the self-value term
// R[m, 32*nt + d] = p_self * v_p,h[d], which the P.V GEMM adds as its residual.
__device__ __forceinline__ void tc_epilogue_softmax(const TcParams& p, uint32_t taddr, int m, int nt, int hsel, int row,
                                                    float* xch) {
  const float sl2 = p.scale * 1.4426950408889634f;
  const bool ok = m < p.M;
  const float* qrow = p.qkv + (int64_t)(ok ? m : 0) * p.ld_qkv + nt * 32;      // q | k (+256) | v (+512)
  float s_self = 0.f;
#pragma unroll
  for (int d = 0; d < 32; d += 4) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(qrow + d));
    const float4 k = __ldg(reinterpret_cast<const float4*>(qrow + 256 + d));
    s_self = fmaf(q.x, k.x, s_self); s_self = fmaf(q.y, k.y, s_self); s_self = fmaf(q.z, k.z, s_self); s_self = fmaf(q.w, k.w, s_self);
  }
  const int lc0 = hsel * 128;
  float mx = -INFINITY;
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t rr[32];
    tmem_ld_32x32(taddr + c0, rr);
    tmem_ld_wait();
    if (lc0 + c0 + 32 <= p.n_keys) {
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(rr[j]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (lc0 + c0 + j < p.n_keys) mx = fmaxf(mx, __uint_as_float(rr[j]));
    }
  }
  xch[hsel * 128 + row] = mx;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  mx = fmaxf(fmaxf(xch[row], xch[128 + row]), s_self);
  asm volatile("bar.sync 1, 256;" ::: "memory");
  // single sweep: unnormalised e_j = exp(scale*(s_j - max)) goes to P, the row sum is applied later as a per-(row, head)
  // scale in the P.V epilogue (p.rowscale there) -- one exp per score, no third TMEM pass.
  float sum = 0.f;
  const float mxs = mx * sl2;
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t rr[32];
    tmem_ld_32x32(taddr + c0, rr);
    tmem_ld_wait();
    const int lc = lc0 + c0;
    const int nvalid = p.n_tile_valid - lc;
    if (nvalid > 0) {
      float v[32];
      if (lc + 32 <= p.n_keys) {
#pragma unroll
        for (int j = 0; j < 32; ++j) { v[j] = fast_ex2(fmaf(__uint_as_float(rr[j]), sl2, -mxs)); sum += v[j]; }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = (lc + j < p.n_keys) ? fast_ex2(fmaf(__uint_as_float(rr[j]), sl2, -mxs)) : 0.f;
          sum += v[j];
        }
      }
      if (ok) {
        float* prow = p.C + (int64_t)m * p.ldc + nt * p.c_nt_off + lc;
        if (nvalid >= 32 && ((reinterpret_cast<uintptr_t>(prow) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            reinterpret_cast<float4*>(prow)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < nvalid) prow[j] = v[j];
        }
      }
    }
  }
  xch[hsel * 128 + row] = sum;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const float e_self = fast_ex2(fmaf(s_self, sl2, -mxs));
  const float inv = 1.0f / (xch[row] + xch[128 + row] + e_self);
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (hsel == 0 && ok) {
    p.R_inv[(int64_t)m * 8 + nt] = inv;
    const float ps = e_self * inv;
    float* rrow = p.R + (int64_t)m * 256 + nt * 32;
#pragma unroll
    for (int d = 0; d < 32; d += 4) {
      const float4 vv = __ldg(reinterpret_cast<const float4*>(qrow + 512 + d));
      *reinterpret_cast<float4*>(rrow + d) = make_float4(ps * vv.x, ps * vv.y, ps * vv.z, ps * vv.w);
    }
  }
}

// AMODE 0: A is a dense row-major fp32 matrix;  AMODE 1: A rows are gathered from an NHWC image (im2col on the fly);
// AMODE 2: A rows are gathered from the output gradient of a convolution (data gradient as an implicit GEMM)
// SPLIT: compile-time copy of `precision == 0` for the A producers (single-pass bf16 skips the hi/lo residual arithmetic: the
// producers are bound by instruction issue, profiles/r1_ncu_train_gemm.md)
// split-K: the raw partial accumulator of split s goes to ws[s][m][n] (n in [0, N)), 128 bytes per thread and 32-column piece
__device__ __forceinline__ void tc_epilogue_split(const TcParams& p, uint32_t taddr, int m, int nt, int hsel, int s) {
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    if (hsel * 128 + c0 >= p.mma_n) break;
    uint32_t rr[32];
    tmem_ld_32x32(taddr + c0, rr);
    tmem_ld_wait();
    const int n0 = nt * TC_BN + hsel * 128 + c0;
    const int nvalid = p.N - n0;
    if (m < p.M && nvalid > 0) {
      float* dst = p.ws + ((int64_t)s * p.M + m) * p.N + n0;
      if (nvalid >= 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          reinterpret_cast<float4*>(dst)[j] = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]),
                                                          __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) dst[j] = __uint_as_float(rr[j]);
      }
    }
  }
}

// C = act(sum_s ws[s] + bias) combined with the residual as tc_epilogue_half does; splits summed in index order
__global__ void splitk_finalize_kernel(const float* __restrict__ ws, int ks, int M, int N, const float* __restrict__ bias,
                                       const float* __restrict__ res, int ldres, int res_mode, int act, float* __restrict__ C, int ldc) {
  const int64_t total = (int64_t)M * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const int64_t m = i / N;
    float v = ws[i];
    for (int s = 1; s < ks; ++s) v += ws[(int64_t)s * total + i];
    if (bias) v += __ldg(bias + n);
    if (res_mode == ZS_RES_NONE) {
      v = apply_act(v, act);
    } else {
      const float r = __ldg(res + m * ldres + n);
      v = res_mode == ZS_RES_BEFORE_ACT ? apply_act(v + r, act) : apply_act(v, act) + r;
    }
    C[m * ldc + n] = v;
  }
}

template <int AMODE, bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // 1024-aligned operand tiles
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + TC_STAGES * TC_STAGE_BYTES;
  // barriers: full[2], empty[2], tmem_full[2], tmem_empty[2], then tmem base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 16u + 8u * s; };
  auto tfull_bar = [&](int a) { return bar_base + 32u + 8u * a; };
  auto tempty_bar = [&](int a) { return bar_base + 48u + 8u * a; };
  const uint32_t tmem_slot = bar_base + 64u;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + TC_STAGES * TC_STAGE_BYTES + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool split = SPLIT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar(s), TC_PROD_WARPS + 1);        // one arrival per A-producer warp + 1 expect_tx arrival of the W loader
      mbar_init(empty_bar(s), 1);        // one tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);        // one tcgen05.commit
      mbar_init(tempty_bar(a), 256);     // 256 epilogue threads
    }
    fence_mbar_init();
  }
  if (warp == TC_MMA_WARP) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int ksp = p.k_splits > 1 ? p.k_splits : 1;           // entry points that never split leave the field zero
  const int total_items = total_tiles * ksp;                 // work item w = (tile w % total_tiles, K split w / total_tiles)
  auto kbeg = [&](int w) { return (int)((int64_t)(w / total_tiles) * p.k_chunks / ksp); };
  auto kend = [&](int w) { return (int)((int64_t)(w / total_tiles + 1) * p.k_chunks / ksp); };

  if (warp < TC_PROD_WARPS) {
    // ================= A producers =================
    // thread = (16-byte bf16 chunk c of the 64-wide K-chunk, rows rb + 32 i): a warp reads four 256-byte row segments per
    // instruction pair (coalesced) and writes whole 128-byte swizzled smem rows (conflict-free).  Eight warps, because one
    // warp per scheduler converting a chunk is latency-bound (profiles/r1_gemm_tc_v0.md, the chained kernels' lesson).
    const int c = threadIdx.x & 7, rb = threadIdx.x >> 3;
    int stage = 0; uint32_t phase = 0;
    const bool vec_ok = ((p.lda & 3) == 0) && ((p.a_nt_off & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    // The 8 fp32 of each unit are fetched into registers one K-chunk ahead of their conversion (also across tile boundaries),
    // so the global / L2 latency of the next chunk overlaps the MMAs that still own the smem slot.  (A second register buffer
    // -- two chunks of look-ahead -- was measured on B200: the conv modes spill at the 96-register cap of an 18-warp CTA and
    // every GEMM class of the training step got 5-25 % slower; profiles/r1_train_launches.md.)
    float4 buf0[4][2];
    int cb[4] = {0, 0, 0, 0}, cih0[4] = {0, 0, 0, 0}, ciw0[4] = {0, 0, 0, 0};   // conv modes: pixel of each of the thread's rows
    bool crow_ok[4] = {false, false, false, false};
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch = [&](float4 (&buf)[4][2], int t, int kc, bool first) {
      const int m_base = (t / p.n_tiles) * TC_BM + rb;
      const int k = kc * TC_BK + c * 8;
      if (AMODE != 0 && first) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int m = m_base + 32 * i;
          crow_ok[i] = m < p.M;
          if (AMODE == 1) {
            const int ow = m % p.cOW, tt = m / p.cOW;
            cb[i] = tt / p.cOH;
            cih0[i] = (tt % p.cOH) * p.cStride - p.cPadT;
            ciw0[i] = ow * p.cStride - p.cPadL;
          } else {
            const int iw = m % p.cW, tt = m / p.cW;
            cb[i] = tt / p.cH;
            cih0[i] = (tt % p.cH) + p.cPadT;
            ciw0[i] = iw + p.cPadL;
          }
        }
      }
      if (AMODE == 1 || AMODE == 2) {
        // K index = (kh, kw, channel); the two float4 of a unit share a filter tap when the channel count is a multiple of 8
        int kh[2], kw[2], ch[2];
        bool kok[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int kk = k + 4 * h;
          if (h == 1 && (p.cCin & 7) == 0) { kh[1] = kh[0]; kw[1] = kw[0]; ch[1] = ch[0] + 4; kok[1] = kok[0]; continue; }
          const int tap = kk / p.cCin;
          ch[h] = kk - tap * p.cCin;
          kh[h] = tap / p.cKW;
          kw[h] = tap - kh[h] * p.cKW;
          kok[h] = kk < p.K;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float4 v = zero4;
            if (AMODE == 1) {
              const int ih = cih0[i] + kh[h], iw = ciw0[i] + kw[h];
              if (crow_ok[i] && kok[h] && (unsigned)ih < (unsigned)p.cH && (unsigned)iw < (unsigned)p.cW)
                v = __ldg(reinterpret_cast<const float4*>(p.A + (((int64_t)cb[i] * p.cH + ih) * p.cW + iw) * p.cCin + ch[h]));
              if (p.cPreRelu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
            } else {
              // data gradient: rows are INPUT pixels, the gather runs over dY [B, OH, OW, Cout] (p.cCin holds Cout):
              // oh = (ih + pad_top - kh) / stride when exact and in range, else zero
              const int a = cih0[i] - kh[h], bb = ciw0[i] - kw[h];
              int oh = a, ow = bb;
              bool ok = crow_ok[i] && kok[h] && a >= 0 && bb >= 0;
              if (p.cStride != 1) {
                oh = a / p.cStride; ow = bb / p.cStride;
                ok = ok && oh * p.cStride == a && ow * p.cStride == bb;
              }
              ok = ok && oh < p.cOH && ow < p.cOW;
              if (ok) v = __ldg(reinterpret_cast<const float4*>(p.A + (((int64_t)cb[i] * p.cOH + oh) * p.cOW + ow) * p.cCin + ch[h]));
            }
            buf[i][h] = v;
          }
        }
        return;
      }
      const int a_off = (t % p.n_tiles) * p.a_nt_off;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m_base + 32 * i;
        const bool row_ok = m < p.M;
        const float* arow = p.A + (int64_t)(row_ok ? m : 0) * p.lda + a_off;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int kk = k + 4 * h;
          if (row_ok && vec_ok && kk + 4 <= p.K) {
            buf[i][h] = __ldg(reinterpret_cast<const float4*>(arow + kk));
          } else {
            float tt[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) tt[e] = (row_ok && kk + e < p.K) ? __ldg(arow + kk + e) : 0.f;
            buf[i][h] = make_float4(tt[0], tt[1], tt[2], tt[3]);
          }
        }
      }
    };
    auto consume = [&](float4 (&buf)[4][2]) {
      mbar_wait(empty_bar(stage), phase ^ 1);
      uint8_t* a_hi = smem_gen + stage * TC_STAGE_BYTES;
      uint8_t* a_lo = a_hi + TC_A_TILE;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 x0 = buf[i][0], x1 = buf[i][1];
        const uint32_t off = swizzle128_offset(rb + 32 * i, c);
        if (SPLIT) {
          uint4 hi, lo;
          split_bf16x2(x0.x, x0.y, hi.x, lo.x);
          split_bf16x2(x0.z, x0.w, hi.y, lo.y);
          split_bf16x2(x1.x, x1.y, hi.z, lo.z);
          split_bf16x2(x1.z, x1.w, hi.w, lo.w);
          *reinterpret_cast<uint4*>(a_hi + off) = hi;
          *reinterpret_cast<uint4*>(a_lo + off) = lo;
        } else {
          *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(pack_bf16x2(x0.x, x0.y), pack_bf16x2(x0.z, x0.w), pack_bf16x2(x1.x, x1.y),
                                                             pack_bf16x2(x1.z, x1.w));
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(stage));      // one arrival per producer warp (its 32 stores are fenced and ordered)
      if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
    };
    // fetch iterator (runs one (tile, chunk) item ahead of the consume iterator, in the same order)
    int fw = blockIdx.x, fkc = fw < total_items ? kbeg(fw) : 0;
    auto fetch_next = [&](float4 (&buf)[4][2]) {
      if (fw >= total_items) return;
      fetch(buf, fw % total_tiles, fkc, fkc == kbeg(fw));
      if (++fkc == kend(fw)) { fw += gridDim.x; fkc = fw < total_items ? kbeg(fw) : 0; }
    };
    fetch_next(buf0);
    int64_t items = 0;
    for (int w = blockIdx.x; w < total_items; w += gridDim.x) items += kend(w) - kbeg(w);
    for (int64_t it = 0; it < items; ++it) {
      consume(buf0);
      fetch_next(buf0);
    }
  } else if (warp == TC_W_WARP) {
    // ================= W loader =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t bytes = split ? 2u * p.w_tile_bytes : p.w_tile_bytes;
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const int nt = (w % total_tiles) % p.n_tiles;
        for (int kc = kbeg(w); kc < kend(w); ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint8_t* src = p.Wp + ((size_t)nt * p.k_chunks + kc) * (2u * TC_B_TILE);
          const uint32_t dst = smem_base + stage * TC_STAGE_BYTES + 2 * TC_A_TILE;
          mbar_arrive_expect_tx(full_bar(stage), bytes);
          bulk_g2s(dst, src, p.w_tile_bytes, full_bar(stage));
          if (split) bulk_g2s(dst + TC_B_TILE, src + TC_B_TILE, p.w_tile_bytes, full_bar(stage));
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == TC_MMA_WARP) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint32_t idesc = umma_idesc_bf16(TC_BM, (uint32_t)p.mma_n);
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TC_BN;
        const int kb = kbeg(w), ke = kend(w);
        for (int kc = kb; kc < ke; ++kc) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * TC_STAGE_BYTES;
          const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + TC_A_TILE);
          const uint64_t b_hi = umma_desc_sw128(sa + 2 * TC_A_TILE), b_lo = umma_desc_sw128(sa + 2 * TC_A_TILE + TC_B_TILE);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t koff = (uint64_t)((k * 32) >> 4);   // 16 bf16 = 32 B along K inside the swizzle atom
            umma_bf16(d_tmem, a_hi + koff, b_hi + koff, idesc, (kc != kb || k != 0) ? 1u : 0u);
            if (split) {
              umma_bf16(d_tmem, a_hi + koff, b_lo + koff, idesc, 1);
              umma_bf16(d_tmem, a_lo + koff, b_hi + koff, idesc, 1);
            }
          }
          umma_commit(empty_bar(stage));                 // smem stage reusable when these MMAs retire
          if (kc == ke - 1) umma_commit(tfull_bar(acc));
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue: 8 warps; warp e -> TMEM lane quarter e&3, column half e>>2 =================
    // (two warps per SM sub-partition and a compile-time activation: a single warp per scheduler running a
    //  jump-table per element was latency-bound at ~0.2 IPC -- profiles/r1_gemm_tc_v0.md)
    const int e = warp - TC_PROD_WARPS, q = e & 3, hsel = e >> 2;
    float* xch = reinterpret_cast<float*>(smem_gen + TC_STAGES * TC_STAGE_BYTES + 256);   // [2][128] softmax exchange
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
      const int t = w % total_tiles;
      const int mt = t / p.n_tiles, nt = t % p.n_tiles;
      const int m = mt * TC_BM + q * 32 + lane;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * TC_BN + ((uint32_t)(q * 32) << 16) + hsel * 128;
      if (p.k_splits > 1) {
        tc_epilogue_split(p, taddr, m, nt, hsel, w / total_tiles);
      } else if (p.epi_mode == 1) {
        tc_epilogue_softmax(p, taddr, m, nt, hsel, q * 32 + lane, xch);
      } else {
        switch (p.act) {
          case ZS_ACT_RELU: tc_epilogue_half<ZS_ACT_RELU>(p, taddr, m, nt, hsel); break;
          case ZS_ACT_GELU: tc_epilogue_half<ZS_ACT_GELU>(p, taddr, m, nt, hsel); break;
          case ZS_ACT_SOFTPLUS100: tc_epilogue_half<ZS_ACT_SOFTPLUS100>(p, taddr, m, nt, hsel); break;
          case ZS_ACT_SIGMOID: tc_epilogue_half<ZS_ACT_SIGMOID>(p, taddr, m, nt, hsel); break;
          case ZS_ACT_CLAMP01: tc_epilogue_half<ZS_ACT_CLAMP01>(p, taddr, m, nt, hsel); break;
          default: tc_epilogue_half<ZS_ACT_NONE>(p, taddr, m, nt, hsel); break;
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- weight packing: fp32 W[N,K] -> per (n_tile, k_chunk) [hi | lo] swizzled bf16 tiles --------------
// fmt 0: bf16 (gemm_tc / encoder / training kernels), 1: fp16 (decoder chain kernels, csrc/chain_tc.cu)
__global__ void gemm_tc_pack_kernel(const float* __restrict__ W, int ldw, int N, int K, uint8_t* __restrict__ out,
                                    int n_tiles, int k_chunks, int fmt) {
  // one thread per (tile row, 16-byte chunk)
  int64_t total = (int64_t)n_tiles * k_chunks * TC_BN * 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i & 7);
    int64_t t = i >> 3;
    int r = (int)(t % TC_BN); t /= TC_BN;
    int kc = (int)(t % k_chunks);
    int nt = (int)(t / k_chunks);
    int n = nt * TC_BN + r, k = kc * TC_BK + c * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (n < N && k + e < K) ? W[(int64_t)n * ldw + k + e] : 0.f;
    uint4 hi, lo;
    if (fmt == 1) {
      tc::split_f16x2(v[0], v[1], hi.x, lo.x);
      tc::split_f16x2(v[2], v[3], hi.y, lo.y);
      tc::split_f16x2(v[4], v[5], hi.z, lo.z);
      tc::split_f16x2(v[6], v[7], hi.w, lo.w);
    } else {
      tc::split_bf16x2(v[0], v[1], hi.x, lo.x);
      tc::split_bf16x2(v[2], v[3], hi.y, lo.y);
      tc::split_bf16x2(v[4], v[5], hi.z, lo.z);
      tc::split_bf16x2(v[6], v[7], hi.w, lo.w);
    }
    uint8_t* tile = out + ((size_t)nt * k_chunks + kc) * (2u * TC_B_TILE);
    uint32_t off = tc::swizzle128_offset(r, c);
    *reinterpret_cast<uint4*>(tile + off) = hi;
    *reinterpret_cast<uint4*>(tile + TC_B_TILE + off) = lo;
  }
}

}  // namespace zs

using namespace zs;

// A layer narrower than one 256-column tile runs its MMAs at N = round_up(N, 16) and copies only those weight rows (the
// first mma_n rows of a packed K-major tile are contiguous): less tensor work, 1/8 .. 1/2 of the weight traffic from L2.
static void tc_narrow(zs::TcParams& p) {
  if (p.n_tiles != 1 || p.N >= zs::TC_BN) return;
  const int n16 = (p.N + 15) / 16 * 16;
  p.mma_n = n16 < 16 ? 16 : n16;
  p.n_tile_valid = p.N;
  p.w_tile_bytes = (uint32_t)p.mma_n * 128u;
}

// Launch of a plain forward GEMM / convolution (epi_mode 0).  Few-tile layers -- the deep, small-image layers of the encoder: one
// 128 x 256 tile per SM is busy for K / 64 chunk steps of ~0.8 us while most SMs idle -- are split along K: ks work items per tile,
// raw partial tiles to a stream-ordered workspace, one deterministic finalize pass (bias / activation / residual).
static int g_splitk_enable = 1, g_splitk_min_chunks = 2;     // at least this many 64-wide K-chunks per split (a chunk step is ~2 us
                                                              // of latency for a lone CTA: 5.46 -> 5.14 ms for the batch-1 encoder against 4)
template <int AMODE>
static int tc_launch(zs::TcParams& p, cudaStream_t st, const char* name) {
  using namespace zs;
  const int tiles = p.m_tiles * p.n_tiles, sms = sm_count();
  int ks = 1;
  if (g_splitk_enable && p.epi_mode == 0 && p.rowscale == nullptr && tiles * 2 <= sms && p.k_chunks >= g_splitk_min_chunks * 2) {
    ks = sms / tiles;
    if (ks > p.k_chunks / g_splitk_min_chunks) ks = p.k_chunks / g_splitk_min_chunks;
    if (ks > 16) ks = 16;
    if (ks < 2) ks = 1;
  }
  p.k_splits = ks; p.ws = nullptr;
  if (ks > 1) {
    // stream-ordered workspace from the device's default memory pool; the pool must KEEP its memory across synchronisation points
    // (release threshold 0, the default, hands it back to the driver at every sync: measured 10-100 ms per re-allocation)
    static thread_local int pool_dev = -1;
    int dev = 0;
    ZS_CUDA_CALL(cudaGetDevice(&dev));
    if (pool_dev != dev) {
      cudaMemPool_t pool;
      ZS_CUDA_CALL(cudaDeviceGetDefaultMemPool(&pool, dev));
      unsigned long long keep = ~0ull;
      ZS_CUDA_CALL(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
      pool_dev = dev;
    }
    ZS_CUDA_CALL(cudaMallocAsync(reinterpret_cast<void**>(&p.ws), (size_t)ks * p.M * p.N * sizeof(float), st));
  }
  const int items = tiles * ks;
  const int grid = items < sms ? items : sms;
  if (p.precision == 0) gemm_tc_kernel<AMODE, true><<<grid, TC_THREADS, TC_SMEM, st>>>(p);
  else gemm_tc_kernel<AMODE, false><<<grid, TC_THREADS, TC_SMEM, st>>>(p);
  ZS_CUDA_CHECK_LAUNCH(name);
  if (ks > 1) {
    const int64_t total = (int64_t)p.M * p.N;
    int fgrid = (int)((total + 255) / 256);
    if (fgrid > sms * 8) fgrid = sms * 8;
    splitk_finalize_kernel<<<fgrid, 256, 0, st>>>(p.ws, ks, p.M, p.N, p.bias, p.res, p.ldres, p.res_mode, p.act, p.C, p.ldc);
    ZS_CUDA_CHECK_LAUNCH(name);
    ZS_CUDA_CALL(cudaFreeAsync(p.ws, st));
  }
  return ZS_OK;
}

/* debug / A-B: 0 disables the split-K path of zs_gemm_tc_f32 / zs_conv2d_nhwc_tc (process-wide) */
extern "C" int zs_debug_gemm_splitk(int enable) {      // 0 = off, 1 = on (default granularity), n >= 2 = on with n chunks per split
  g_splitk_enable = enable != 0;
  g_splitk_min_chunks = enable >= 2 ? enable : 2;
  return ZS_OK;
}

extern "C" size_t zs_gemm_tc_packed_bytes(int N, int K) {
  if (N <= 0 || K <= 0) return 0;
  size_t nt = (N + TC_BN - 1) / TC_BN, kc = (K + TC_BK - 1) / TC_BK;
  return nt * kc * 2u * TC_B_TILE;
}

extern "C" int zs_gemm_tc_pack(const float* W, int ldw, int N, int K, void* packed, void* stream) {
  return zs_gemm_tc_pack_fmt(W, ldw, N, K, packed, 0, stream);
}

extern "C" int zs_gemm_tc_pack_fmt(const float* W, int ldw, int N, int K, void* packed, int fmt, void* stream) {
  ZS_REQUIRE(W && packed && N > 0 && K > 0 && ldw >= K && (fmt == 0 || fmt == 1), "zs_gemm_tc_pack: bad args");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "zs_gemm_tc_pack: packed buffer must be 16-byte aligned");
  int nt = (N + TC_BN - 1) / TC_BN, kc = (K + TC_BK - 1) / TC_BK;
  int64_t total = (int64_t)nt * kc * TC_BN * 8;
  int grid = (int)((total + 255) / 256);
  if (grid > 4096) grid = 4096;
  gemm_tc_pack_kernel<<<grid, 256, 0, as_stream(stream)>>>(W, ldw, N, K, reinterpret_cast<uint8_t*>(packed), nt, kc, fmt);
  ZS_CUDA_CHECK_LAUNCH("zs_gemm_tc_pack");
  return ZS_OK;
}

extern "C" int zs_gemm_tc_f32(const float* A, int lda, const void* Wpacked, const float* bias,
                              const float* res, int ldres, int res_mode, float* C, int ldc,
                              int M, int N, int K, int act, int precision, void* stream) {
  ZS_REQUIRE(A && Wpacked && C, "zs_gemm_tc_f32: null pointer");
  ZS_REQUIRE(M >= 0 && N > 0 && K > 0 && lda >= K && ldc >= N, "zs_gemm_tc_f32: bad shape");
  ZS_REQUIRE(res_mode == ZS_RES_NONE || res != nullptr, "zs_gemm_tc_f32: residual requested but res==NULL");
  ZS_REQUIRE(precision == 0 || precision == 1, "zs_gemm_tc_f32: precision must be 0 (bf16x3) or 1 (bf16)");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(Wpacked) & 15) == 0, "zs_gemm_tc_f32: packed weights must be 16-byte aligned");
  if (M == 0) return ZS_OK;
  static thread_local bool configured = false;
  if (!configured) {
    ZS_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    ZS_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    configured = true;
  }
  TcParams p{};
  p.A = A; p.lda = lda; p.Wp = reinterpret_cast<const uint8_t*>(Wpacked); p.bias = bias;
  p.res = res; p.ldres = ldres; p.res_mode = res_mode; p.C = C; p.ldc = ldc;
  p.M = M; p.N = N; p.K = K; p.act = act; p.precision = precision;
  p.m_tiles = (M + TC_BM - 1) / TC_BM; p.n_tiles = (N + TC_BN - 1) / TC_BN; p.k_chunks = (K + TC_BK - 1) / TC_BK;
  p.a_nt_off = 0; p.c_nt_off = TC_BN; p.n_tile_valid = TC_BN; p.mma_n = TC_BN; p.w_tile_bytes = TC_B_TILE; p.epi_mode = 0;
  tc_narrow(p);
  p.qkv = nullptr; p.ld_qkv = 0; p.R = nullptr; p.R_inv = nullptr; p.scale = 0.f; p.n_keys = 0; p.rowscale = nullptr;
  return tc_launch<0>(p, as_stream(stream), "zs_gemm_tc_f32");
}

// ---- implicit-GEMM convolution on the tensor cores (NHWC fp32 activations, OHWI filters packed as W[Cout, KH*KW*Cin]) ----
extern "C" int zs_conv2d_nhwc_tc(const float* x, int B, int H, int W, int Cin, const void* Wpacked, const float* bias,
                                 const float* res, int res_mode, float* y, int Cout, int KH, int KW, int stride,
                                 int pad_top, int pad_left, int OH, int OW, int act, int pre_relu, int precision, void* stream) {
  ZS_REQUIRE(x && Wpacked && y, "zs_conv2d_nhwc_tc: null pointer");
  ZS_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && OH > 0 && OW > 0,
             "zs_conv2d_nhwc_tc: bad shape");
  ZS_REQUIRE((Cin & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "zs_conv2d_nhwc_tc: needs Cin %% 4 == 0 and a 16-byte aligned image");
  ZS_REQUIRE(res_mode == ZS_RES_NONE || res != nullptr, "zs_conv2d_nhwc_tc: residual requested but res==NULL");
  ZS_REQUIRE(precision == 0 || precision == 1, "zs_conv2d_nhwc_tc: precision must be 0 (bf16x3) or 1 (bf16)");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(Wpacked) & 15) == 0, "zs_conv2d_nhwc_tc: packed weights must be 16-byte aligned");
  const int64_t M64 = (int64_t)B * OH * OW;
  ZS_REQUIRE(M64 < (1LL << 31), "zs_conv2d_nhwc_tc: too many output pixels");
  static thread_local bool configured = false;
  if (!configured) {
    ZS_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    ZS_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    configured = true;
  }
  TcParams p{};
  p.A = x; p.lda = 0; p.Wp = reinterpret_cast<const uint8_t*>(Wpacked); p.bias = bias;
  p.res = res; p.ldres = Cout; p.res_mode = res_mode; p.C = y; p.ldc = Cout;
  p.M = (int)M64; p.N = Cout; p.K = KH * KW * Cin; p.act = act; p.precision = precision;
  p.m_tiles = (p.M + TC_BM - 1) / TC_BM; p.n_tiles = (p.N + TC_BN - 1) / TC_BN; p.k_chunks = (p.K + TC_BK - 1) / TC_BK;
  p.a_nt_off = 0; p.c_nt_off = TC_BN; p.n_tile_valid = TC_BN; p.mma_n = TC_BN; p.w_tile_bytes = TC_B_TILE; p.epi_mode = 0;
  tc_narrow(p);
  p.cB = B; p.cH = H; p.cW = W; p.cCin = Cin; p.cKH = KH; p.cKW = KW; p.cStride = stride; p.cPadT = pad_top; p.cPadL = pad_left;
  p.cOH = OH; p.cOW = OW; p.cPreRelu = pre_relu;
  return tc_launch<1>(p, as_stream(stream), "zs_conv2d_nhwc_tc");
}

// ---- attention on the tensor cores (two grouped launches of gemm_tc_kernel) ------------------------------------
// scores + softmax numerators:  P[m, h*208 + j] = exp(scale*(q_h(m).k_lat_h(j) - max)),  Rinv[m,h] = 1/sum (incl. the point's own
// key), R[m, 32h+d] = p_self * v_p,h[d]
// `Kpacked`: 8 tiles (one per head) packed with zs_gemm_tc_pack from K_lat_h [208 (197 + zero rows), 32].
extern "C" int zs_attn_scores_tc(const float* qkv, int ld_qkv, const void* Kpacked, int M, int n_keys, float scale,
                                 float* P, float* R, float* Rinv, int precision, void* stream) {
  ZS_REQUIRE(qkv && Kpacked && P && R && Rinv && M >= 0, "zs_attn_scores_tc: null pointer");
  ZS_REQUIRE(n_keys > 0 && n_keys <= 208, "zs_attn_scores_tc: n_keys must be in [1, 208]");
  ZS_REQUIRE((ld_qkv & 3) == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(R) & 15) == 0,
             "zs_attn_scores_tc: qkv / R must be 16-byte aligned with ld % 4 == 0");
  ZS_REQUIRE(precision == 0 || precision == 1, "zs_attn_scores_tc: bad precision");
  if (M == 0) return ZS_OK;
  static thread_local bool configured = false;
  if (!configured) {
    ZS_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    ZS_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    configured = true;
  }
  TcParams p{};
  p.A = qkv; p.lda = ld_qkv; p.Wp = reinterpret_cast<const uint8_t*>(Kpacked); p.bias = nullptr;
  p.res = nullptr; p.ldres = 0; p.res_mode = ZS_RES_NONE; p.C = P; p.ldc = 8 * 208;
  p.M = M; p.N = 8 * 208; p.K = 32; p.act = ZS_ACT_NONE; p.precision = precision;
  p.m_tiles = (M + TC_BM - 1) / TC_BM; p.n_tiles = 8; p.k_chunks = 1;
  p.a_nt_off = 32; p.c_nt_off = 208; p.n_tile_valid = 208; p.mma_n = 208; p.w_tile_bytes = 208 * 128; p.epi_mode = 1;
  p.qkv = qkv; p.ld_qkv = ld_qkv; p.R = R; p.R_inv = Rinv; p.scale = scale; p.n_keys = n_keys;
  int tiles = p.m_tiles * p.n_tiles;
  int grid = tiles < sm_count() ? tiles : sm_count();
  if (p.precision == 0) gemm_tc_kernel<0, true><<<grid, TC_THREADS, TC_SMEM, as_stream(stream)>>>(p);
  else gemm_tc_kernel<0, false><<<grid, TC_THREADS, TC_SMEM, as_stream(stream)>>>(p);
  ZS_CUDA_CHECK_LAUNCH("zs_attn_scores_tc");
  return ZS_OK;
}

// O[m, 32h + d] = Rinv[m,h] * sum_j P[m, h*208 + j] * v_lat_h[j, d] + R[m, 32h + d]
// `Vpacked`: 8 heads x 4 K-chunks, packed with zs_gemm_tc_pack from V_lat_h^T [32, 208].
extern "C" int zs_attn_pv_tc(const float* P, const void* Vpacked, const float* R, const float* Rinv, float* O, int M, int precision,
                             void* stream) {
  ZS_REQUIRE(P && Vpacked && R && Rinv && O && M >= 0, "zs_attn_pv_tc: null pointer");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(P) & 15) == 0, "zs_attn_pv_tc: P must be 16-byte aligned");
  ZS_REQUIRE(precision == 0 || precision == 1, "zs_attn_pv_tc: bad precision");
  if (M == 0) return ZS_OK;
  static thread_local bool configured = false;
  if (!configured) {
    ZS_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    ZS_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    configured = true;
  }
  TcParams p{};
  p.A = P; p.lda = 8 * 208; p.Wp = reinterpret_cast<const uint8_t*>(Vpacked); p.bias = nullptr;
  p.res = R; p.ldres = 256; p.res_mode = ZS_RES_AFTER_ACT; p.C = O; p.ldc = 256;
  p.M = M; p.N = 256; p.K = 208; p.act = ZS_ACT_NONE; p.precision = precision;
  p.m_tiles = (M + TC_BM - 1) / TC_BM; p.n_tiles = 8; p.k_chunks = 4;
  p.a_nt_off = 208; p.c_nt_off = 32; p.n_tile_valid = 32; p.mma_n = 32; p.w_tile_bytes = 32 * 128; p.epi_mode = 0;
  p.rowscale = Rinv;
  int tiles = p.m_tiles * p.n_tiles;
  int grid = tiles < sm_count() ? tiles : sm_count();
  if (p.precision == 0) gemm_tc_kernel<0, true><<<grid, TC_THREADS, TC_SMEM, as_stream(stream)>>>(p);
  else gemm_tc_kernel<0, false><<<grid, TC_THREADS, TC_SMEM, as_stream(stream)>>>(p);
  ZS_CUDA_CHECK_LAUNCH("zs_attn_pv_tc");
  return ZS_OK;
}

// ---- data gradient of an NHWC convolution on the tensor cores ------------------------------------------------------
// dx[b,ih,iw,ci] = sum_{kh,kw,co} dy[b,oh,ow,co] w[co,kh,kw,ci]: gemm_tc_kernel<2> with rows = input pixels, the K index
// (kh, kw, co) gathered from dy, and `Wpacked` = zs_gemm_tc_pack of the filter re-laid as Wd[Cin, KH*KW*Cout].
// Replaces torch autograd's conv2d input gradient in the training step (formerly zs_conv2d_nhwc_dgrad_f32, FFMA).
extern "C" int zs_conv2d_nhwc_dgrad_tc(const float* dy, int B, int H, int W, int Cin, const void* Wpacked, float* dx, int Cout,
                                       int KH, int KW, int stride, int pad_top, int pad_left, int OH, int OW, int precision,
                                       void* stream) {
  ZS_REQUIRE(dy && Wpacked && dx, "zs_conv2d_nhwc_dgrad_tc: null pointer");
  ZS_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (Cout & 3) == 0 && KH > 0 && KW > 0 && stride > 0 && OH > 0 && OW > 0,
             "zs_conv2d_nhwc_dgrad_tc: bad shape (Cout %% 4 == 0 required)");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wpacked) & 15) == 0,
             "zs_conv2d_nhwc_dgrad_tc: dy / packed weights must be 16-byte aligned");
  ZS_REQUIRE(precision == 0 || precision == 1, "zs_conv2d_nhwc_dgrad_tc: precision must be 0 (bf16x3) or 1 (bf16)");
  const int64_t M64 = (int64_t)B * H * W;
  ZS_REQUIRE(M64 < (1LL << 31), "zs_conv2d_nhwc_dgrad_tc: too many pixels");
  static thread_local bool configured = false;
  if (!configured) {
    ZS_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    ZS_CUDA_CALL(cudaFuncSetAttribute(gemm_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    configured = true;
  }
  TcParams p{};
  p.A = dy; p.lda = 0; p.Wp = reinterpret_cast<const uint8_t*>(Wpacked); p.bias = nullptr;
  p.res = nullptr; p.ldres = Cin; p.res_mode = ZS_RES_NONE; p.C = dx; p.ldc = Cin;
  p.M = (int)M64; p.N = Cin; p.K = KH * KW * Cout; p.act = ZS_ACT_NONE; p.precision = precision;
  p.m_tiles = (p.M + TC_BM - 1) / TC_BM; p.n_tiles = (p.N + TC_BN - 1) / TC_BN; p.k_chunks = (p.K + TC_BK - 1) / TC_BK;
  p.a_nt_off = 0; p.c_nt_off = TC_BN; p.n_tile_valid = TC_BN; p.mma_n = TC_BN; p.w_tile_bytes = TC_B_TILE; p.epi_mode = 0;
  tc_narrow(p);
  p.cB = B; p.cH = H; p.cW = W; p.cCin = Cout; p.cKH = KH; p.cKW = KW; p.cStride = stride; p.cPadT = pad_top; p.cPadL = pad_left;
  p.cOH = OH; p.cOW = OW; p.cPreRelu = 0;
  int tiles = p.m_tiles * p.n_tiles;
  int grid = tiles < sm_count() ? tiles : sm_count();
  if (p.precision == 0) gemm_tc_kernel<2, true><<<grid, TC_THREADS, TC_SMEM, as_stream(stream)>>>(p);
  else gemm_tc_kernel<2, false><<<grid, TC_THREADS, TC_SMEM, as_stream(stream)>>>(p);
  ZS_CUDA_CHECK_LAUNCH("zs_conv2d_nhwc_dgrad_tc");
  return ZS_OK;
}
