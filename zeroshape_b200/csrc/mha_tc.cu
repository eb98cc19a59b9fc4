// Token self-attention on the tensor cores (tcgen05 + tensor memory): softmax(q k^T * scale) v for the 12 ViT blocks of the
// DPT-hybrid backbone (timm Attention.forward inside Block, model/depth/vit.py:149-150) and the latent branch of the implicit
// decoder (model/shape/implicit.py:65-71).  T <= 208 tokens (197 = 14 x 14 + cls), head dim 32 or 64.
//
// One CTA per (image, head, 128-query tile).  All 256 threads convert the head's q tile, k and v from the packed fp32 qkv buffer
// into split-fp16 UMMA operand tiles in shared memory (v transposed: it is the B operand of P V, K-major over the keys, stored as
// [Vh rows | Vl rows] so that Ph Vh and Ph Vl are one instruction).  Then, the construction of chain_qkvattn2_kernel:
//   S = Q K^T              (M = 128, N = 208, K = head dim; Qh Kh + Ql Kh + Qh Kl, fp32 accumulation in tensor memory)
//   softmax                (4 warps, thread = query row: max sweep, then exp2 sweep that overwrites the scores IN PLACE with the
//                           split-fp16 probabilities, 32 score columns -> 16 packed hi + 16 packed lo columns)
//   O = P V                (A operand read from tensor memory, tcgen05.mma [d], [a], b: 13 K-steps of 16 keys)
//   out = O / rowsum       (thread = row, 16-byte stores)
// No pipelining: a CTA lives for one tile (~6 MFLOP); the launch fills the SMs with B x heads x 2 CTAs.
#include "common.cuh"
#include "tc_common.cuh"

namespace zs {
using namespace tc;

constexpr int MT_THREADS = 256;
constexpr int MT_OFF_Q = 0;                       // q tile  hi | lo : 2 x 128 rows x 128 B
constexpr int MT_OFF_K = 32 * 1024;               // k       hi | lo : 2 x 208 (256) rows x 128 B
constexpr int MT_OFF_V = MT_OFF_K + 64 * 1024;    // v^T: 4 key chunks x [Vh rows | Vl rows] x 128 B  (<= 4 x 16 KB)
constexpr int MT_OFF_BAR = MT_OFF_V + 64 * 1024;
constexpr int MT_SMEM = MT_OFF_BAR + 64 + 1024;   // + slack for the 1024-byte alignment of the base

struct MhaTcParams {
  const float* qkv; float* out; int B, T, heads, hd; float scale; int precision;
};

template <int NK, bool MASKED>
__device__ __forceinline__ void mt_exp_tmem(const uint32_t* rr, int k0, int n_keys, float sl2, float mxs, float2& sum2,
                                            uint32_t taddr, bool split) {
  uint32_t hi[NK / 2], lo[NK / 2];
#pragma unroll
  for (int j = 0; j < NK / 2; ++j) {
    const int i = 2 * j;
    const float2 a = fma2(make_float2(__uint_as_float(rr[i]), __uint_as_float(rr[i + 1])), bc2(sl2), bc2(-mxs));
    float2 v = make_float2(fast_ex2(a.x), fast_ex2(a.y));
    if (MASKED) { if (k0 + i >= n_keys) v.x = 0.f; if (k0 + i + 1 >= n_keys) v.y = 0.f; }
    sum2 = add2(sum2, v);
    split_f16x2(v.x, v.y, hi[j], lo[j]);
  }
  if constexpr (NK == 32) {
    tmem_st_32x16(taddr, hi);
    if (split) tmem_st_32x16(taddr + 16, lo);
  } else {
    tmem_st_32x8(taddr, hi);
    if (split) tmem_st_32x8(taddr + 8, lo);
  }
}

template <int HD>
__global__ void __launch_bounds__(MT_THREADS, 1) mha_tc_kernel(MhaTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_s = smem_base + MT_OFF_BAR, bar_o = bar_s + 8, tmem_slot = bar_s + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = p.precision == 0;
  const int T = p.T, C = p.heads * HD;
  const int bh = blockIdx.x, b = bh / p.heads, h = bh % p.heads;
  const int q0 = blockIdx.y * 128;
  constexpr int VCH = 2 * HD * 128;                 // bytes of one 64-key chunk of v^T: (Vh | Vl) rows x 128 B
  const float* base = p.qkv + (int64_t)b * T * 3 * C + h * HD;

  if (threadIdx.x == 0) { mbar_init(bar_s, 1); mbar_init(bar_o, 1); fence_mbar_init(); }
  if (warp == 4) tmem_alloc(tmem_slot, 512);

  // ---- operand tiles: thread = (token, 8 consecutive dims) ----
  constexpr int CPR = HD / 8;                        // 16-byte fp16 chunks per row
  for (int i = threadIdx.x; i < 128 * CPR; i += MT_THREADS) {
    const int r = i / CPR, c = i % CPR, t = q0 + r;
    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
    if (t < T) {
      const float4* src = reinterpret_cast<const float4*>(base + (int64_t)t * 3 * C + 8 * c);
      x0 = __ldg(src); x1 = __ldg(src + 1);
    }
    uint4 hi, lo;
    split_f16x2(x0.x, x0.y, hi.x, lo.x); split_f16x2(x0.z, x0.w, hi.y, lo.y);
    split_f16x2(x1.x, x1.y, hi.z, lo.z); split_f16x2(x1.z, x1.w, hi.w, lo.w);
    const uint32_t off = swizzle128_offset(r, c);
    *reinterpret_cast<uint4*>(smem + MT_OFF_Q + off) = hi;
    *reinterpret_cast<uint4*>(smem + MT_OFF_Q + 16384 + off) = lo;
  }
  for (int i = threadIdx.x; i < 208 * CPR; i += MT_THREADS) {
    const int j = i / CPR, c = i % CPR;
    float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f), k1 = k0, v0 = k0, v1 = k0;
    if (j < T) {
      const float4* ks = reinterpret_cast<const float4*>(base + (int64_t)j * 3 * C + C + 8 * c);
      const float4* vs = reinterpret_cast<const float4*>(base + (int64_t)j * 3 * C + 2 * C + 8 * c);
      k0 = __ldg(ks); k1 = __ldg(ks + 1); v0 = __ldg(vs); v1 = __ldg(vs + 1);
    }
    uint4 hi, lo;
    split_f16x2(k0.x, k0.y, hi.x, lo.x); split_f16x2(k0.z, k0.w, hi.y, lo.y);
    split_f16x2(k1.x, k1.y, hi.z, lo.z); split_f16x2(k1.z, k1.w, hi.w, lo.w);
    const uint32_t off = swizzle128_offset(j, c);
    *reinterpret_cast<uint4*>(smem + MT_OFF_K + off) = hi;
    *reinterpret_cast<uint4*>(smem + MT_OFF_K + 32768 + off) = lo;
    // v^T: element (dim d, key j) -> row d (hi) / HD + d (lo) of key chunk j >> 6, fp16 column j & 63
    const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint8_t* vch = smem + MT_OFF_V + (j >> 6) * VCH;
    const int kc = j & 63;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int d = 8 * c + u;
      const __half hh = __float2half_rn(vv[u]);
      const __half ll = __float2half_rn(vv[u] - __half2float(hh));
      *reinterpret_cast<__half*>(vch + swizzle128_offset(d, kc >> 3) + (kc & 7) * 2) = hh;
      *reinterpret_cast<__half*>(vch + swizzle128_offset(HD + d, kc >> 3) + (kc & 7) * 2) = ll;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + MT_OFF_BAR + 16);
  const uint32_t d_s = tmem_base, d_o = tmem_base + 256;

  if (warp == 4 && lane == 0) {
    // ---- S = Q K^T ----
    const uint32_t idesc_s = umma_idesc_f16(128, 208);
    const uint64_t qh = umma_desc_sw128(smem_base + MT_OFF_Q), ql = umma_desc_sw128(smem_base + MT_OFF_Q + 16384);
    const uint64_t kh = umma_desc_sw128(smem_base + MT_OFF_K), kl = umma_desc_sw128(smem_base + MT_OFF_K + 32768);
#pragma unroll
    for (int k = 0; k < HD / 16; ++k) {
      umma_bf16(d_s, qh + 2 * k, kh + 2 * k, idesc_s, k > 0 ? 1u : 0u);
      if (split) {
        umma_bf16(d_s, ql + 2 * k, kh + 2 * k, idesc_s, 1u);
        umma_bf16(d_s, qh + 2 * k, kl + 2 * k, idesc_s, 1u);
      }
    }
    umma_commit(bar_s);
  }

  float inv = 0.f;
  if (warp < 4) {
    // ---- softmax: thread = query row ----
    const uint32_t s_tm = d_s + ((uint32_t)(warp * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = -3.0e38f;
#pragma unroll 1
    for (int c = 0; c < 6; ++c) {
      uint32_t rr[32];
      tmem_ld_32x32(s_tm + 32 * c, rr);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) if (32 * c + j < T) mx = fmaxf(mx, __uint_as_float(rr[j]));
    }
    {
      uint32_t rr[16];
      tmem_ld_32x16(s_tm + 192, rr);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) if (192 + j < T) mx = fmaxf(mx, __uint_as_float(rr[j]));
    }
    const float mxs = mx * sl2;
    float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int c = 0; c < 6; ++c) {
      uint32_t rr[32];
      tmem_ld_32x32(s_tm + 32 * c, rr);
      tmem_ld_wait();
      if (32 * c + 32 <= T) mt_exp_tmem<32, false>(rr, 32 * c, T, sl2, mxs, sum2, s_tm + 32 * c, split);
      else mt_exp_tmem<32, true>(rr, 32 * c, T, sl2, mxs, sum2, s_tm + 32 * c, split);
    }
    {
      uint32_t rr[16];
      tmem_ld_32x16(s_tm + 192, rr);
      tmem_ld_wait();
      mt_exp_tmem<16, true>(rr, 192, T, sl2, mxs, sum2, s_tm + 192, split);
    }
    tmem_st_wait();
    inv = 1.0f / (sum2.x + sum2.y);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 4 && lane == 0) {
    // ---- O = P V: A = the probabilities in tensor memory, B = v^T chunks; Ph [Vh | Vl] is one N = 2 HD instruction ----
    const uint32_t idesc_hi = split ? umma_idesc_f16(128, 2 * HD) : umma_idesc_f16(128, HD), idesc_lo = umma_idesc_f16(128, HD);
#pragma unroll 1
    for (int j = 0; j < 13; ++j) {                  // K-step j = keys 16 j .. 16 j + 15
      const uint32_t a_hi = d_s + 32u * (j >> 1) + 8u * (j & 1);
      const uint32_t a_lo = a_hi + (j == 12 ? 8u : 16u);
      const uint64_t vv = umma_desc_sw128(smem_base + MT_OFF_V + (j >> 2) * VCH) + 2 * (j & 3);
      umma_ts(d_o, a_hi, vv, idesc_hi, j > 0 ? 1u : 0u);
      if (split) umma_ts(d_o, a_lo, vv, idesc_lo, 1u);
    }
    umma_commit(bar_o);
  }

  if (warp < 4) {
    const uint32_t o_tm = d_o + ((uint32_t)(warp * 32) << 16);
    const int t = q0 + warp * 32 + lane;
    mbar_wait(bar_o, 0);
    tc_fence_after();
    float* dst = p.out + ((int64_t)b * T + (t < T ? t : 0)) * C + h * HD;
#pragma unroll 1
    for (int c = 0; c < HD / 16; ++c) {
      uint32_t rr[16], r2[16];
      tmem_ld_32x16(o_tm + 16 * c, rr);
      if (split) tmem_ld_32x16(o_tm + HD + 16 * c, r2);
      tmem_ld_wait();
      if (t < T) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a = __uint_as_float(rr[4 * j + i]);
            if (split) a += __uint_as_float(r2[4 * j + i]);
            o[i] = a * inv;
          }
          *reinterpret_cast<float4*>(dst + 16 * c + 4 * j) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}


// ===============================================================================================================
// Forward of the decoder's point attention inside the training tape (ImplFuncAttention, model/shape/implicit.py:25-79: every query
// point attends to the image's L <= 208 latent tokens plus its OWN key / value): the construction of mha_tc_kernel with the keys and
// values read from the latent-side buffers and one extra softmax column per row handled in registers --
//   s_self = q . k_self (fp32), e_self = exp(scale (s_self - max)), out = (P V_lat + e_self v_self) / (sum_lat + e_self).
// One CTA per (image, head, 128-point tile), head dim 32.  Replaces the one-thread-per-point FFMA kernel (zs_point_attention_f32)
// when the training engine runs on the tensor cores in its single-pass mode.
struct PaFwdParams {
  const float* qkv_p; const float* k_lat; const float* v_lat; int ld_lat; float* out; int B, P, L, heads; float scale; int precision;
};

__global__ void __launch_bounds__(MT_THREADS, 1) pa_fwd_tc_kernel(PaFwdParams p) {
  constexpr int HD = 32;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_s = smem_base + MT_OFF_BAR, bar_o = bar_s + 8, tmem_slot = bar_s + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = p.precision == 0;
  const int L = p.L, C = p.heads * HD;
  const int bh = blockIdx.x, b = bh / p.heads, h = bh % p.heads;
  const int q0 = blockIdx.y * 128;
  constexpr int VCH = 2 * HD * 128;
  const float* qbase = p.qkv_p + (int64_t)b * p.P * 3 * C + h * HD;           // row t: q | + C: k_self | + 2 C: v_self
  const float* kbase = p.k_lat + (int64_t)b * L * p.ld_lat + h * HD;
  const float* vbase = p.v_lat + (int64_t)b * L * p.ld_lat + h * HD;

  if (threadIdx.x == 0) { mbar_init(bar_s, 1); mbar_init(bar_o, 1); fence_mbar_init(); }
  if (warp == 4) tmem_alloc(tmem_slot, 512);

  constexpr int CPR = HD / 8;
  for (int i = threadIdx.x; i < 128 * CPR; i += MT_THREADS) {
    const int r = i / CPR, c = i % CPR, t = q0 + r;
    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
    if (t < p.P) {
      const float4* src = reinterpret_cast<const float4*>(qbase + (int64_t)t * 3 * C + 8 * c);
      x0 = __ldg(src); x1 = __ldg(src + 1);
    }
    uint4 hi, lo;
    split_f16x2(x0.x, x0.y, hi.x, lo.x); split_f16x2(x0.z, x0.w, hi.y, lo.y);
    split_f16x2(x1.x, x1.y, hi.z, lo.z); split_f16x2(x1.z, x1.w, hi.w, lo.w);
    const uint32_t off = swizzle128_offset(r, c);
    *reinterpret_cast<uint4*>(smem + MT_OFF_Q + off) = hi;
    *reinterpret_cast<uint4*>(smem + MT_OFF_Q + 16384 + off) = lo;
  }
  for (int i = threadIdx.x; i < 208 * CPR; i += MT_THREADS) {
    const int j = i / CPR, c = i % CPR;
    float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f), k1 = k0, v0 = k0, v1 = k0;
    if (j < L) {
      const float4* ks = reinterpret_cast<const float4*>(kbase + (int64_t)j * p.ld_lat + 8 * c);
      const float4* vs = reinterpret_cast<const float4*>(vbase + (int64_t)j * p.ld_lat + 8 * c);
      k0 = __ldg(ks); k1 = __ldg(ks + 1); v0 = __ldg(vs); v1 = __ldg(vs + 1);
    }
    uint4 hi, lo;
    split_f16x2(k0.x, k0.y, hi.x, lo.x); split_f16x2(k0.z, k0.w, hi.y, lo.y);
    split_f16x2(k1.x, k1.y, hi.z, lo.z); split_f16x2(k1.z, k1.w, hi.w, lo.w);
    const uint32_t off = swizzle128_offset(j, c);
    *reinterpret_cast<uint4*>(smem + MT_OFF_K + off) = hi;
    *reinterpret_cast<uint4*>(smem + MT_OFF_K + 32768 + off) = lo;
    const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint8_t* vch = smem + MT_OFF_V + (j >> 6) * VCH;
    const int kc = j & 63;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int d = 8 * c + u;
      const __half hh = __float2half_rn(vv[u]);
      const __half ll = __float2half_rn(vv[u] - __half2float(hh));
      *reinterpret_cast<__half*>(vch + swizzle128_offset(d, kc >> 3) + (kc & 7) * 2) = hh;
      *reinterpret_cast<__half*>(vch + swizzle128_offset(HD + d, kc >> 3) + (kc & 7) * 2) = ll;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + MT_OFF_BAR + 16);
  const uint32_t d_s = tmem_base, d_o = tmem_base + 256;

  if (warp == 4 && lane == 0) {
    const uint32_t idesc_s = umma_idesc_f16(128, 208);
    const uint64_t qh = umma_desc_sw128(smem_base + MT_OFF_Q), ql = umma_desc_sw128(smem_base + MT_OFF_Q + 16384);
    const uint64_t kh = umma_desc_sw128(smem_base + MT_OFF_K), kl = umma_desc_sw128(smem_base + MT_OFF_K + 32768);
#pragma unroll
    for (int k = 0; k < HD / 16; ++k) {
      umma_bf16(d_s, qh + 2 * k, kh + 2 * k, idesc_s, k > 0 ? 1u : 0u);
      if (split) {
        umma_bf16(d_s, ql + 2 * k, kh + 2 * k, idesc_s, 1u);
        umma_bf16(d_s, qh + 2 * k, kl + 2 * k, idesc_s, 1u);
      }
    }
    umma_commit(bar_s);
  }

  float inv = 0.f, e_self = 0.f;
  const int t = q0 + (warp & 3) * 32 + lane;
  if (warp < 4) {
    // the point's own key: s_self in fp32 while the S product runs
    float s_self = -3.0e38f;
    if (t < p.P) {
      const float4* q4 = reinterpret_cast<const float4*>(qbase + (int64_t)t * 3 * C);
      const float4* k4 = reinterpret_cast<const float4*>(qbase + (int64_t)t * 3 * C + C);
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < HD / 4; ++i) {
        const float4 a = __ldg(q4 + i), kk = __ldg(k4 + i);
        acc = fmaf(a.x, kk.x, acc); acc = fmaf(a.y, kk.y, acc); acc = fmaf(a.z, kk.z, acc); acc = fmaf(a.w, kk.w, acc);
      }
      s_self = acc;
    }
    const uint32_t s_tm = d_s + ((uint32_t)(warp * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = s_self;
#pragma unroll 1
    for (int c = 0; c < 6; ++c) {
      uint32_t rr[32];
      tmem_ld_32x32(s_tm + 32 * c, rr);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) if (32 * c + j < L) mx = fmaxf(mx, __uint_as_float(rr[j]));
    }
    {
      uint32_t rr[16];
      tmem_ld_32x16(s_tm + 192, rr);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) if (192 + j < L) mx = fmaxf(mx, __uint_as_float(rr[j]));
    }
    const float mxs = mx * sl2;
    float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int c = 0; c < 6; ++c) {
      uint32_t rr[32];
      tmem_ld_32x32(s_tm + 32 * c, rr);
      tmem_ld_wait();
      if (32 * c + 32 <= L) mt_exp_tmem<32, false>(rr, 32 * c, L, sl2, mxs, sum2, s_tm + 32 * c, split);
      else mt_exp_tmem<32, true>(rr, 32 * c, L, sl2, mxs, sum2, s_tm + 32 * c, split);
    }
    {
      uint32_t rr[16];
      tmem_ld_32x16(s_tm + 192, rr);
      tmem_ld_wait();
      mt_exp_tmem<16, true>(rr, 192, L, sl2, mxs, sum2, s_tm + 192, split);
    }
    tmem_st_wait();
    e_self = t < p.P ? fast_ex2(fmaf(s_self, sl2, -mxs)) : 0.f;
    inv = 1.0f / (sum2.x + sum2.y + e_self);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 4 && lane == 0) {
    const uint32_t idesc_hi = split ? umma_idesc_f16(128, 2 * HD) : umma_idesc_f16(128, HD), idesc_lo = umma_idesc_f16(128, HD);
#pragma unroll 1
    for (int j = 0; j < 13; ++j) {
      const uint32_t a_hi = d_s + 32u * (j >> 1) + 8u * (j & 1);
      const uint32_t a_lo = a_hi + (j == 12 ? 8u : 16u);
      const uint64_t vv = umma_desc_sw128(smem_base + MT_OFF_V + (j >> 2) * VCH) + 2 * (j & 3);
      umma_ts(d_o, a_hi, vv, idesc_hi, j > 0 ? 1u : 0u);
      if (split) umma_ts(d_o, a_lo, vv, idesc_lo, 1u);
    }
    umma_commit(bar_o);
  }

  if (warp < 4) {
    const uint32_t o_tm = d_o + ((uint32_t)(warp * 32) << 16);
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const int tt = t < p.P ? t : 0;
    float* dst = p.out + ((int64_t)b * p.P + tt) * C + h * HD;
    const float4* v4 = reinterpret_cast<const float4*>(qbase + (int64_t)tt * 3 * C + 2 * C);
#pragma unroll 1
    for (int c = 0; c < HD / 16; ++c) {
      uint32_t rr[16], r2[16];
      tmem_ld_32x16(o_tm + 16 * c, rr);
      if (split) tmem_ld_32x16(o_tm + HD + 16 * c, r2);
      tmem_ld_wait();
      if (t < p.P) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 vs = __ldg(v4 + 4 * c + j);
          const float vself[4] = {vs.x, vs.y, vs.z, vs.w};
          float o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a = __uint_as_float(rr[4 * j + i]);
            if (split) a += __uint_as_float(r2[4 * j + i]);
            o[i] = fmaf(e_self, vself[i], a) * inv;
          }
          *reinterpret_cast<float4*>(dst + 16 * c + 4 * j) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ===============================================================================================================
// Backward of the token self-attention on the tensor cores (single fp16 pass, fp32 accumulation: the precision class of the
// bf16 training mode; the FFMA kernels zs_mha_bwd_f32 stay the fp32-grade path).  With P = softmax(S), S = scale Q K^T:
//     dV = P^T dO,   dP = dO V^T,   D = rowsum(P o dP),   dS = scale P o (dP - D),   dQ = dS K,   dK = dS^T Q.
// Two launches, each the one-shot construction of mha_tc_kernel (operands converted by all threads, MMAs by one thread, the
// A operands P / dS re-written IN PLACE in tensor memory as packed fp16 over the fp32 scores):
//   mha_bwd_q_kernel  (image, head, 128-query tile): S and dP side by side in tensor memory; a thread per query row makes P, D
//                     and dS; dQ = dS K (B operand = K^T, scatter-transposed in shared memory); row statistics (max * scale * log2 e,
//                     1 / sum, D) to `stats` for the second launch.
//   mha_bwd_kv_kernel (image, head, 128-key tile): the TRANSPOSED products S^T = K Q^T, dP^T = V dO^T (operands in their natural
//                     row-major form), a thread per key row rebuilds P^T and dS^T from the per-query statistics, then dV = P^T dO and
//                     dK = dS^T Q with dO^T / Q^T as scatter-transposed B operands; one accumulator, read out twice.
constexpr int MB_OFF_A0 = 0;                    // q kernel: Q tile            kv kernel: K tile           (16 KB)
constexpr int MB_OFF_A1 = 16 * 1024;            //           dO tile                      V tile           (16 KB)
constexpr int MB_OFF_B0 = 32 * 1024;            //           K  [208 x hd]                Q  [208 x hd]    (32 KB)
constexpr int MB_OFF_B1 = 64 * 1024;            //           V  [208 x hd]                dO [208 x hd]    (32 KB)
constexpr int MB_OFF_T0 = 96 * 1024;            //           K^T chunks                   dO^T chunks      (32 KB)
constexpr int MB_OFF_T1 = 128 * 1024;           //           --                           Q^T chunks       (32 KB)
constexpr int MB_OFF_ST = 160 * 1024;           // kv kernel: [208][3] per-query statistics (2.5 KB)
constexpr int MB_OFF_BAR = 164 * 1024;
constexpr int MB_SMEM = MB_OFF_BAR + 64 + 1024;

struct MhaBwdParams {
  const float* qkv; const float* dO; float* dqkv; float* stats; int B, T, heads, hd; float scale;
};

// rows [r0, r0 + nrows) of a [T, hd] matrix (row stride ld floats) -> single-fp16 K-major SW128 tile; optionally also its transpose
// as 64-column chunks (rows = dims, columns = tokens) for use as the B operand of a product that contracts over the tokens
template <int HD, bool TRANSPOSE>
__device__ __forceinline__ void mb_stage(const float* src, int64_t ld, int r0, int nrows, int T, uint8_t* tile, uint8_t* ttile) {
  constexpr int CPR = HD / 8;
  for (int i = threadIdx.x; i < nrows * CPR; i += MT_THREADS) {
    const int r = i / CPR, c = i % CPR, t = r0 + r;
    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
    if (t < T) {
      const float4* s4 = reinterpret_cast<const float4*>(src + (int64_t)t * ld + 8 * c);
      x0 = __ldg(s4); x1 = __ldg(s4 + 1);
    }
    const uint4 h = make_uint4(cvt_f16x2_sat(x0.x, x0.y), cvt_f16x2_sat(x0.z, x0.w), cvt_f16x2_sat(x1.x, x1.y), cvt_f16x2_sat(x1.z, x1.w));
    *reinterpret_cast<uint4*>(tile + swizzle128_offset(r, c)) = h;
    if (TRANSPOSE) {
      const float vv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      uint8_t* ch = ttile + (r >> 6) * (HD * 128);
      const int kc = r & 63;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        *reinterpret_cast<__half*>(ch + swizzle128_offset(8 * c + u, kc >> 3) + (kc & 7) * 2) = __float2half_rn(vv[u]);
    }
  }
}

// 32 (16) fp32 values -> packed fp16 pairs in the first 16 (8) columns of their own block (A operand of a TS MMA, single pass)
template <int NK>
__device__ __forceinline__ void mb_store_packed(uint32_t taddr, const float* v) {
  uint32_t h[NK / 2];
#pragma unroll
  for (int j = 0; j < NK / 2; ++j) h[j] = cvt_f16x2_sat(v[2 * j], v[2 * j + 1]);
  if constexpr (NK == 32) tmem_st_32x16(taddr, *reinterpret_cast<uint32_t(*)[16]>(h));
  else tmem_st_32x8(taddr, *reinterpret_cast<uint32_t(*)[8]>(h));
}

// one thread issues D[128 x N] = A[tmem, packed fp16 over `keys` tokens] . B^T with B = the transposed chunks at tchunks
template <int HD>
__device__ __forceinline__ void mb_ts_product(uint32_t d_acc, uint32_t a_tm, uint32_t tchunks_addr, bool accumulate = false) {
  const uint32_t idesc = umma_idesc_f16(128, HD);
#pragma unroll 1
  for (int j = 0; j < 13; ++j) {
    const uint32_t a = a_tm + 32u * (j >> 1) + 8u * (j & 1);
    const uint64_t b = umma_desc_sw128(tchunks_addr + (j >> 2) * (HD * 128)) + 2 * (j & 3);
    umma_ts(d_acc, a, b, idesc, (accumulate || j > 0) ? 1u : 0u);
  }
}

template <int HD>
__global__ void __launch_bounds__(MT_THREADS, 1) mha_bwd_q_kernel(MhaBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + MB_OFF_BAR, bar1 = bar0 + 8, tmem_slot = bar0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, C = p.heads * HD;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads, q0 = blockIdx.y * 128;
  const float* base = p.qkv + (int64_t)b * T * 3 * C + h * HD;
  const float* dob = p.dO + (int64_t)b * T * C + h * HD;
  if (threadIdx.x == 0) { mbar_init(bar0, 1); mbar_init(bar1, 1); fence_mbar_init(); }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  mb_stage<HD, false>(base, 3 * C, q0, 128, T, smem + MB_OFF_A0, nullptr);                // Q tile
  mb_stage<HD, false>(dob, C, q0, 128, T, smem + MB_OFF_A1, nullptr);                     // dO tile
  mb_stage<HD, true>(base + C, 3 * C, 0, 208, T, smem + MB_OFF_B0, smem + MB_OFF_T0);     // K and K^T
  mb_stage<HD, false>(base + 2 * C, 3 * C, 0, 208, T, smem + MB_OFF_B1, nullptr);         // V
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + MB_OFF_BAR + 16);
  const uint32_t d_s = tmem_base, d_p = tmem_base + 208, d_q = tmem_base + 416;
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = umma_idesc_f16(128, 208);
    const uint64_t qa = umma_desc_sw128(smem_base + MB_OFF_A0), da = umma_desc_sw128(smem_base + MB_OFF_A1);
    const uint64_t kb = umma_desc_sw128(smem_base + MB_OFF_B0), vb = umma_desc_sw128(smem_base + MB_OFF_B1);
#pragma unroll
    for (int k = 0; k < HD / 16; ++k) umma_bf16(d_s, qa + 2 * k, kb + 2 * k, idesc, k > 0 ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < HD / 16; ++k) umma_bf16(d_p, da + 2 * k, vb + 2 * k, idesc, k > 0 ? 1u : 0u);
    umma_commit(bar0);
  }
  if (warp < 4) {
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const uint32_t s_tm = d_s + lane_off, p_tm = d_p + lane_off;
    const float sl2 = p.scale * 1.4426950408889634f;
    mbar_wait(bar0, 0);
    tc_fence_after();
    float mx = -3.0e38f;
#pragma unroll 1
    for (int c = 0; c < 7; ++c) {
      uint32_t rr[32];
      if (c < 6) tmem_ld_32x32(s_tm + 32 * c, rr); else tmem_ld_32x16(s_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rr));
      tmem_ld_wait();
      const int n = c < 6 ? 32 : 16;
      for (int j = 0; j < n; ++j) if (32 * c + j < T) mx = fmaxf(mx, __uint_as_float(rr[j]));
    }
    const float mxs = mx * sl2;
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < 7; ++c) {
      uint32_t rr[32];
      if (c < 6) tmem_ld_32x32(s_tm + 32 * c, rr); else tmem_ld_32x16(s_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rr));
      tmem_ld_wait();
      const int n = c < 6 ? 32 : 16;
      for (int j = 0; j < n; ++j) if (32 * c + j < T) sum += fast_ex2(fmaf(__uint_as_float(rr[j]), sl2, -mxs));
    }
    const float inv = 1.0f / sum;
    float D = 0.f;
#pragma unroll 1
    for (int c = 0; c < 7; ++c) {
      uint32_t rs[32], rp[32];
      if (c < 6) { tmem_ld_32x32(s_tm + 32 * c, rs); tmem_ld_32x32(p_tm + 32 * c, rp); }
      else { tmem_ld_32x16(s_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rs)); tmem_ld_32x16(p_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rp)); }
      tmem_ld_wait();
      const int n = c < 6 ? 32 : 16;
      for (int j = 0; j < n; ++j)
        if (32 * c + j < T) D = fmaf(fast_ex2(fmaf(__uint_as_float(rs[j]), sl2, -mxs)) * inv, __uint_as_float(rp[j]), D);
    }
#pragma unroll 1
    for (int c = 0; c < 7; ++c) {
      uint32_t rs[32], rp[32];
      float ds[32];
      if (c < 6) { tmem_ld_32x32(s_tm + 32 * c, rs); tmem_ld_32x32(p_tm + 32 * c, rp); }
      else { tmem_ld_32x16(s_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rs)); tmem_ld_32x16(p_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rp)); }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float pr = fast_ex2(fmaf(__uint_as_float(rs[j]), sl2, -mxs)) * inv;
        ds[j] = (32 * c + j < T && (c < 6 || j < 16)) ? p.scale * pr * (__uint_as_float(rp[j]) - D) : 0.f;
      }
      if (c < 6) mb_store_packed<32>(p_tm + 32 * c, ds); else mb_store_packed<16>(p_tm + 192, ds);
    }
    tmem_st_wait();
    const int t = q0 + warp * 32 + lane;
    if (t < T) {
      float* st = p.stats + (((int64_t)b * p.heads + h) * T + t) * 3;
      st[0] = mxs; st[1] = inv; st[2] = D;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && lane == 0) {
    mb_ts_product<HD>(d_q, d_p, smem_base + MB_OFF_T0);        // dQ = dS K
    umma_commit(bar1);
  }
  if (warp < 4) {
    const uint32_t o_tm = d_q + ((uint32_t)(warp * 32) << 16);
    const int t = q0 + warp * 32 + lane;
    mbar_wait(bar1, 0);
    tc_fence_after();
    float* dst = p.dqkv + ((int64_t)b * T + (t < T ? t : 0)) * 3 * C + h * HD;
#pragma unroll 1
    for (int c = 0; c < HD / 16; ++c) {
      uint32_t rr[16];
      tmem_ld_32x16(o_tm + 16 * c, rr);
      tmem_ld_wait();
      if (t < T) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(dst + 16 * c + 4 * j) = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]),
                                                                        __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int HD>
__global__ void __launch_bounds__(MT_THREADS, 1) mha_bwd_kv_kernel(MhaBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + MB_OFF_BAR, bar1 = bar0 + 8, bar2 = bar0 + 24, tmem_slot = bar0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T, C = p.heads * HD;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads, k0 = blockIdx.y * 128;
  const float* base = p.qkv + (int64_t)b * T * 3 * C + h * HD;
  const float* dob = p.dO + (int64_t)b * T * C + h * HD;
  float* stq = reinterpret_cast<float*>(smem + MB_OFF_ST);
  if (threadIdx.x == 0) { mbar_init(bar0, 1); mbar_init(bar1, 1); mbar_init(bar2, 1); fence_mbar_init(); }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  mb_stage<HD, false>(base + C, 3 * C, k0, 128, T, smem + MB_OFF_A0, nullptr);            // K tile
  mb_stage<HD, false>(base + 2 * C, 3 * C, k0, 128, T, smem + MB_OFF_A1, nullptr);        // V tile
  mb_stage<HD, true>(base, 3 * C, 0, 208, T, smem + MB_OFF_B0, smem + MB_OFF_T1);         // Q and Q^T
  mb_stage<HD, true>(dob, C, 0, 208, T, smem + MB_OFF_B1, smem + MB_OFF_T0);              // dO and dO^T
  for (int i = threadIdx.x; i < 208 * 3; i += MT_THREADS)
    stq[i] = i < T * 3 ? __ldg(p.stats + ((int64_t)b * p.heads + h) * T * 3 + i) : 0.f;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + MB_OFF_BAR + 16);
  const uint32_t d_s = tmem_base, d_p = tmem_base + 208, d_o = tmem_base + 416;
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = umma_idesc_f16(128, 208);
    const uint64_t ka = umma_desc_sw128(smem_base + MB_OFF_A0), va = umma_desc_sw128(smem_base + MB_OFF_A1);
    const uint64_t qb = umma_desc_sw128(smem_base + MB_OFF_B0), db = umma_desc_sw128(smem_base + MB_OFF_B1);
#pragma unroll
    for (int k = 0; k < HD / 16; ++k) umma_bf16(d_s, ka + 2 * k, qb + 2 * k, idesc, k > 0 ? 1u : 0u);    // S^T  = K Q^T
#pragma unroll
    for (int k = 0; k < HD / 16; ++k) umma_bf16(d_p, va + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);    // dP^T = V dO^T
    umma_commit(bar0);
  }
  if (warp < 4) {
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const uint32_t s_tm = d_s + lane_off, p_tm = d_p + lane_off;
    const float sl2 = p.scale * 1.4426950408889634f;
    mbar_wait(bar0, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < 7; ++c) {          // columns = queries 32 c ...: P^T and dS^T from the per-query statistics
      uint32_t rs[32], rp[32];
      float pt[32], ds[32];
      if (c < 6) { tmem_ld_32x32(s_tm + 32 * c, rs); tmem_ld_32x32(p_tm + 32 * c, rp); }
      else { tmem_ld_32x16(s_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rs)); tmem_ld_32x16(p_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rp)); }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int qi = 32 * c + j;
        const bool ok = qi < T && (c < 6 || j < 16);
        const float* sq = stq + (ok ? qi : 0) * 3;
        const float pr = ok ? fast_ex2(fmaf(__uint_as_float(rs[j]), sl2, -sq[0])) * sq[1] : 0.f;
        pt[j] = pr;
        ds[j] = ok ? p.scale * pr * (__uint_as_float(rp[j]) - sq[2]) : 0.f;
      }
      if (c < 6) { mb_store_packed<32>(s_tm + 32 * c, pt); mb_store_packed<32>(p_tm + 32 * c, ds); }
      else { mb_store_packed<16>(s_tm + 192, pt); mb_store_packed<16>(p_tm + 192, ds); }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && lane == 0) {
    mb_ts_product<HD>(d_o, d_s, smem_base + MB_OFF_T0);        // dV = P^T dO
    umma_commit(bar1);
  }
  const int t = k0 + (warp & 3) * 32 + lane;
  auto read_out = [&](int which) {                             // accumulator -> dqkv[b, t, which, h, :]
    const uint32_t o_tm = d_o + ((uint32_t)(warp * 32) << 16);
    float* dst = p.dqkv + ((int64_t)b * T + (t < T ? t : 0)) * 3 * C + which * C + h * HD;
#pragma unroll 1
    for (int c = 0; c < HD / 16; ++c) {
      uint32_t rr[16];
      tmem_ld_32x16(o_tm + 16 * c, rr);
      tmem_ld_wait();
      if (t < T) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(dst + 16 * c + 4 * j) = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]),
                                                                        __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
      }
    }
  };
  if (warp < 4) {
    mbar_wait(bar1, 0);
    tc_fence_after();
    read_out(2);
  }
  tc_fence_before();
  __syncthreads();                                             // the accumulator has been read: dK may overwrite it
  tc_fence_after();
  if (warp == 4 && lane == 0) {
    mb_ts_product<HD>(d_o, d_p, smem_base + MB_OFF_T1);        // dK = dS^T Q
    umma_commit(bar2);
  }
  if (warp < 4) {
    mbar_wait(bar2, 0);
    tc_fence_after();
    read_out(1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}


// ===============================================================================================================
// Backward of the decoder's point -> (latents + the point's own key) attention (ImplFuncAttention, model/shape/implicit.py:38-57;
// forward: zs_point_attention_f32) on the tensor cores: the two-launch scheme of the token self-attention backward above with
//   * queries = the P query points of an image (128 per CTA), keys = its L <= 208 latent tokens, head dim 32;
//   * the point's OWN key / value as an extra softmax column handled per row in registers (s_self = q . k_self, its probability,
//     dv_self = p_self dO, dk_self = ds_self q, dq += ds_self k_self);
//   * the key-side launch looping over the points in chunks of 208 with dV and dK accumulating in tensor memory.
// Single fp16 pass, fp32 accumulation (precision class of the bf16 training mode).
struct PaBwdParams {
  const float* qkv_p; const float* k_lat; const float* v_lat; int ld_lat; const float* dO;
  float* dqkv_p; float* dk_lat; float* dv_lat; int ld_dlat; float* stats; int B, P, L, heads; float scale;
};

__global__ void __launch_bounds__(MT_THREADS, 1) pa_bwd_q_kernel(PaBwdParams p) {
  constexpr int HD = 32;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + MB_OFF_BAR, bar1 = bar0 + 8, tmem_slot = bar0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = p.P, L = p.L, C = p.heads * HD;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads, p0 = blockIdx.y * 128;
  const float* qb = p.qkv_p + (int64_t)b * P * 3 * C + h * HD;
  const float* dob = p.dO + (int64_t)b * P * C + h * HD;
  const float* kl = p.k_lat + (int64_t)b * L * p.ld_lat + h * HD;
  const float* vl = p.v_lat + (int64_t)b * L * p.ld_lat + h * HD;
  if (threadIdx.x == 0) { mbar_init(bar0, 1); mbar_init(bar1, 1); fence_mbar_init(); }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  mb_stage<HD, false>(qb, 3 * C, p0, 128, P, smem + MB_OFF_A0, nullptr);
  mb_stage<HD, false>(dob, C, p0, 128, P, smem + MB_OFF_A1, nullptr);
  mb_stage<HD, true>(kl, p.ld_lat, 0, 208, L, smem + MB_OFF_B0, smem + MB_OFF_T0);
  mb_stage<HD, false>(vl, p.ld_lat, 0, 208, L, smem + MB_OFF_B1, nullptr);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + MB_OFF_BAR + 16);
  const uint32_t d_s = tmem_base, d_p = tmem_base + 208, d_q = tmem_base + 416;
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = umma_idesc_f16(128, 208);
    const uint64_t qa = umma_desc_sw128(smem_base + MB_OFF_A0), da = umma_desc_sw128(smem_base + MB_OFF_A1);
    const uint64_t kb = umma_desc_sw128(smem_base + MB_OFF_B0), vb = umma_desc_sw128(smem_base + MB_OFF_B1);
#pragma unroll
    for (int k = 0; k < HD / 16; ++k) umma_bf16(d_s, qa + 2 * k, kb + 2 * k, idesc, k > 0 ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < HD / 16; ++k) umma_bf16(d_p, da + 2 * k, vb + 2 * k, idesc, k > 0 ? 1u : 0u);
    umma_commit(bar0);
  }
  float ps = 0.f, ds_self = 0.f;
  const int pt = p0 + (warp & 3) * 32 + lane;
  const bool live = pt < P;
  const float4* qrow = reinterpret_cast<const float4*>(qb + (int64_t)(live ? pt : 0) * 3 * C);
  const float4* grow = reinterpret_cast<const float4*>(dob + (int64_t)(live ? pt : 0) * C);
  if (warp < 4) {
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const uint32_t s_tm = d_s + lane_off, p_tm = d_p + lane_off;
    const float sl2 = p.scale * 1.4426950408889634f;
    // the point's own key / value: raw self score and dO . v_self
    float s_self = 0.f, dp_self = 0.f;
#pragma unroll
    for (int j = 0; j < HD / 4; ++j) {
      const float4 q4 = __ldg(qrow + j), k4 = __ldg(qrow + C / 4 + j), v4 = __ldg(qrow + 2 * C / 4 + j), g4 = __ldg(grow + j);
      s_self = fmaf(q4.x, k4.x, fmaf(q4.y, k4.y, fmaf(q4.z, k4.z, fmaf(q4.w, k4.w, s_self))));
      dp_self = fmaf(g4.x, v4.x, fmaf(g4.y, v4.y, fmaf(g4.z, v4.z, fmaf(g4.w, v4.w, dp_self))));
    }
    mbar_wait(bar0, 0);
    tc_fence_after();
    float mx = s_self;
#pragma unroll 1
    for (int c = 0; c < 7; ++c) {
      uint32_t rr[32];
      if (c < 6) tmem_ld_32x32(s_tm + 32 * c, rr); else tmem_ld_32x16(s_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rr));
      tmem_ld_wait();
      const int n = c < 6 ? 32 : 16;
      for (int j = 0; j < n; ++j) if (32 * c + j < L) mx = fmaxf(mx, __uint_as_float(rr[j]));
    }
    const float mxs = mx * sl2;
    const float e_self = fast_ex2(fmaf(s_self, sl2, -mxs));
    float sum = e_self;
#pragma unroll 1
    for (int c = 0; c < 7; ++c) {
      uint32_t rr[32];
      if (c < 6) tmem_ld_32x32(s_tm + 32 * c, rr); else tmem_ld_32x16(s_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rr));
      tmem_ld_wait();
      const int n = c < 6 ? 32 : 16;
      for (int j = 0; j < n; ++j) if (32 * c + j < L) sum += fast_ex2(fmaf(__uint_as_float(rr[j]), sl2, -mxs));
    }
    const float inv = 1.0f / sum;
    ps = e_self * inv;
    float D = ps * dp_self;
#pragma unroll 1
    for (int c = 0; c < 7; ++c) {
      uint32_t rs[32], rp[32];
      if (c < 6) { tmem_ld_32x32(s_tm + 32 * c, rs); tmem_ld_32x32(p_tm + 32 * c, rp); }
      else { tmem_ld_32x16(s_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rs)); tmem_ld_32x16(p_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rp)); }
      tmem_ld_wait();
      const int n = c < 6 ? 32 : 16;
      for (int j = 0; j < n; ++j)
        if (32 * c + j < L) D = fmaf(fast_ex2(fmaf(__uint_as_float(rs[j]), sl2, -mxs)) * inv, __uint_as_float(rp[j]), D);
    }
    ds_self = p.scale * ps * (dp_self - D);
#pragma unroll 1
    for (int c = 0; c < 7; ++c) {
      uint32_t rs[32], rp[32];
      float ds[32];
      if (c < 6) { tmem_ld_32x32(s_tm + 32 * c, rs); tmem_ld_32x32(p_tm + 32 * c, rp); }
      else { tmem_ld_32x16(s_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rs)); tmem_ld_32x16(p_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rp)); }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float pr = fast_ex2(fmaf(__uint_as_float(rs[j]), sl2, -mxs)) * inv;
        ds[j] = (32 * c + j < L && (c < 6 || j < 16)) ? p.scale * pr * (__uint_as_float(rp[j]) - D) : 0.f;
      }
      if (c < 6) mb_store_packed<32>(p_tm + 32 * c, ds); else mb_store_packed<16>(p_tm + 192, ds);
    }
    tmem_st_wait();
    if (live) {
      float* st = p.stats + (((int64_t)b * p.heads + h) * P + pt) * 3;
      st[0] = mxs; st[1] = inv; st[2] = D;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && lane == 0) {
    mb_ts_product<HD>(d_q, d_p, smem_base + MB_OFF_T0);        // dq (latent part) = dS K_lat
    umma_commit(bar1);
  }
  if (warp < 4) {
    const uint32_t o_tm = d_q + ((uint32_t)(warp * 32) << 16);
    mbar_wait(bar1, 0);
    tc_fence_after();
    float4* dst = reinterpret_cast<float4*>(p.dqkv_p + ((int64_t)b * P + (live ? pt : 0)) * 3 * C + h * HD);
#pragma unroll 1
    for (int c = 0; c < HD / 16; ++c) {
      uint32_t rr[16];
      tmem_ld_32x16(o_tm + 16 * c, rr);
      tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int jj = 4 * c + j;
          const float4 q4 = __ldg(qrow + jj), k4 = __ldg(qrow + C / 4 + jj), g4 = __ldg(grow + jj);
          dst[jj] = make_float4(fmaf(ds_self, k4.x, __uint_as_float(rr[4 * j])), fmaf(ds_self, k4.y, __uint_as_float(rr[4 * j + 1])),
                                fmaf(ds_self, k4.z, __uint_as_float(rr[4 * j + 2])), fmaf(ds_self, k4.w, __uint_as_float(rr[4 * j + 3])));
          dst[C / 4 + jj] = make_float4(ds_self * q4.x, ds_self * q4.y, ds_self * q4.z, ds_self * q4.w);          // dk_self
          dst[2 * C / 4 + jj] = make_float4(ps * g4.x, ps * g4.y, ps * g4.z, ps * g4.w);                          // dv_self
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

__global__ void __launch_bounds__(MT_THREADS, 1) pa_bwd_kv_kernel(PaBwdParams p) {
  constexpr int HD = 32;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + MB_OFF_BAR, bar1 = bar0 + 8, tmem_slot = bar0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = p.P, L = p.L, C = p.heads * HD;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads, k0 = blockIdx.y * 128;
  const float* qb = p.qkv_p + (int64_t)b * P * 3 * C + h * HD;
  const float* dob = p.dO + (int64_t)b * P * C + h * HD;
  const float* kl = p.k_lat + (int64_t)b * L * p.ld_lat + h * HD;
  const float* vl = p.v_lat + (int64_t)b * L * p.ld_lat + h * HD;
  const float* stg = p.stats + ((int64_t)b * p.heads + h) * P * 3;
  float* stq = reinterpret_cast<float*>(smem + MB_OFF_ST);
  if (threadIdx.x == 0) { mbar_init(bar0, 1); mbar_init(bar1, 1); fence_mbar_init(); }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  mb_stage<HD, false>(kl, p.ld_lat, k0, 128, L, smem + MB_OFF_A0, nullptr);        // this CTA's 128 latent keys / values
  mb_stage<HD, false>(vl, p.ld_lat, k0, 128, L, smem + MB_OFF_A1, nullptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + MB_OFF_BAR + 16);
  const uint32_t d_s = tmem_base, d_p = tmem_base + 208, d_v = tmem_base + 416, d_k = tmem_base + 448;
  const float sl2 = p.scale * 1.4426950408889634f;
  const int n_chunks = (P + 207) / 208;
  uint32_t ph = 0;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int c0 = ch * 208;
    mb_stage<HD, true>(qb, 3 * C, c0, 208, P, smem + MB_OFF_B0, smem + MB_OFF_T1);       // Q chunk and its transpose
    mb_stage<HD, true>(dob, C, c0, 208, P, smem + MB_OFF_B1, smem + MB_OFF_T0);          // dO chunk and its transpose
    for (int i = threadIdx.x; i < 208 * 3; i += MT_THREADS) stq[i] = c0 * 3 + i < P * 3 ? __ldg(stg + c0 * 3 + i) : 0.f;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4 && lane == 0) {
      const uint32_t idesc = umma_idesc_f16(128, 208);
      const uint64_t ka = umma_desc_sw128(smem_base + MB_OFF_A0), va = umma_desc_sw128(smem_base + MB_OFF_A1);
      const uint64_t qd = umma_desc_sw128(smem_base + MB_OFF_B0), dd = umma_desc_sw128(smem_base + MB_OFF_B1);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) umma_bf16(d_s, ka + 2 * k, qd + 2 * k, idesc, k > 0 ? 1u : 0u);    // S^T  = K_lat Q^T
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) umma_bf16(d_p, va + 2 * k, dd + 2 * k, idesc, k > 0 ? 1u : 0u);    // dP^T = V_lat dO^T
      umma_commit(bar0);
    }
    if (warp < 4) {
      const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
      const uint32_t s_tm = d_s + lane_off, p_tm = d_p + lane_off;
      mbar_wait(bar0, ph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 7; ++c) {
        uint32_t rs[32], rp[32];
        float ptv[32], ds[32];
        if (c < 6) { tmem_ld_32x32(s_tm + 32 * c, rs); tmem_ld_32x32(p_tm + 32 * c, rp); }
        else { tmem_ld_32x16(s_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rs)); tmem_ld_32x16(p_tm + 192, *reinterpret_cast<uint32_t(*)[16]>(rp)); }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int qi = 32 * c + j;
          const bool ok = c0 + qi < P && (c < 6 || j < 16);
          const float* sq = stq + (ok ? qi : 0) * 3;
          const float pr = ok ? fast_ex2(fmaf(__uint_as_float(rs[j]), sl2, -sq[0])) * sq[1] : 0.f;
          ptv[j] = pr;
          ds[j] = ok ? p.scale * pr * (__uint_as_float(rp[j]) - sq[2]) : 0.f;
        }
        if (c < 6) { mb_store_packed<32>(s_tm + 32 * c, ptv); mb_store_packed<32>(p_tm + 32 * c, ds); }
        else { mb_store_packed<16>(s_tm + 192, ptv); mb_store_packed<16>(p_tm + 192, ds); }
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4 && lane == 0) {
      mb_ts_product<HD>(d_v, d_s, smem_base + MB_OFF_T0, ch > 0);        // dV += P^T dO
      mb_ts_product<HD>(d_k, d_p, smem_base + MB_OFF_T1, ch > 0);        // dK += dS^T Q
      umma_commit(bar1);
    }
    // the next chunk overwrites the operand tiles and the score columns: every thread waits for these MMAs
    mbar_wait(bar1, ph);
    tc_fence_after();
    ph ^= 1;
  }
  if (warp < 4) {
    const int key = k0 + warp * 32 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float* dvd = p.dv_lat + ((int64_t)b * L + (key < L ? key : 0)) * p.ld_dlat + h * HD;
    float* dkd = p.dk_lat + ((int64_t)b * L + (key < L ? key : 0)) * p.ld_dlat + h * HD;
#pragma unroll 1
    for (int c = 0; c < HD / 16; ++c) {
      uint32_t rv[16], rk[16];
      tmem_ld_32x16(d_v + lane_off + 16 * c, rv);
      tmem_ld_32x16(d_k + lane_off + 16 * c, rk);
      tmem_ld_wait();
      if (key < L) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<float4*>(dvd + 16 * c + 4 * j) = make_float4(__uint_as_float(rv[4 * j]), __uint_as_float(rv[4 * j + 1]),
                                                                        __uint_as_float(rv[4 * j + 2]), __uint_as_float(rv[4 * j + 3]));
          *reinterpret_cast<float4*>(dkd + 16 * c + 4 * j) = make_float4(__uint_as_float(rk[4 * j]), __uint_as_float(rk[4 * j + 1]),
                                                                        __uint_as_float(rk[4 * j + 2]), __uint_as_float(rk[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace zs

using namespace zs;

extern "C" int zs_mha_tc_f32(const float* qkv, float* out, int B, int T, int heads, int hd, float scale, int precision,
                             void* stream) {
  ZS_REQUIRE(qkv && out && B >= 0 && heads > 0, "zs_mha_tc_f32: null pointer / bad sizes");
  ZS_REQUIRE(T >= 1 && T <= 208 && (hd == 32 || hd == 64), "zs_mha_tc_f32: T must be in [1, 208] and the head dim 32 or 64");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "zs_mha_tc_f32: qkv / out must be 16-byte aligned");
  ZS_REQUIRE(precision == 0 || precision == 1, "zs_mha_tc_f32: bad precision");
  if (B == 0) return ZS_OK;
  MhaTcParams p{qkv, out, B, T, heads, hd, scale, precision};
  dim3 grid(B * heads, (T + 127) / 128);
  if (hd == 64) {
    ZS_CUDA_CALL(cudaFuncSetAttribute(mha_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM));
    mha_tc_kernel<64><<<grid, MT_THREADS, MT_SMEM, as_stream(stream)>>>(p);
  } else {
    ZS_CUDA_CALL(cudaFuncSetAttribute(mha_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM));
    mha_tc_kernel<32><<<grid, MT_THREADS, MT_SMEM, as_stream(stream)>>>(p);
  }
  ZS_CUDA_CHECK_LAUNCH("zs_mha_tc_f32");
  return ZS_OK;
}

extern "C" int zs_point_attention_tc_f32(const float* qkv_p, const float* k_lat, const float* v_lat, int ld_lat, float* out, int B,
                                        int P, int L, int heads, int hd, float scale, int precision, void* stream) {
  ZS_REQUIRE(qkv_p && k_lat && v_lat && out && B >= 0 && P >= 0 && heads > 0, "zs_point_attention_tc_f32: null pointer / bad sizes");
  ZS_REQUIRE(L >= 1 && L <= 208 && hd == 32, "zs_point_attention_tc_f32: L must be in [1, 208] and the head dim 32");
  ZS_REQUIRE(((reinterpret_cast<uintptr_t>(qkv_p) | reinterpret_cast<uintptr_t>(k_lat) | reinterpret_cast<uintptr_t>(v_lat) |
               reinterpret_cast<uintptr_t>(out)) & 15) == 0 && (ld_lat & 3) == 0 && ld_lat >= heads * hd,
             "zs_point_attention_tc_f32: buffers must be 16-byte aligned, ld_lat a multiple of 4");
  ZS_REQUIRE(precision == 0 || precision == 1, "zs_point_attention_tc_f32: bad precision");
  if (B == 0 || P == 0) return ZS_OK;
  PaFwdParams p{qkv_p, k_lat, v_lat, ld_lat, out, B, P, L, heads, scale, precision};
  ZS_CUDA_CALL(cudaFuncSetAttribute(pa_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM));
  pa_fwd_tc_kernel<<<dim3(B * heads, (P + 127) / 128), MT_THREADS, MT_SMEM, as_stream(stream)>>>(p);
  ZS_CUDA_CHECK_LAUNCH("zs_point_attention_tc_f32");
  return ZS_OK;
}

extern "C" size_t zs_mha_bwd_tc_ws_bytes(int B, int T, int heads) {
  return B > 0 && T > 0 && heads > 0 ? (size_t)B * heads * T * 3 * sizeof(float) : 0;
}

extern "C" int zs_mha_bwd_tc_f32(const float* qkv, const float* dO, float* dqkv, int B, int T, int heads, int hd, float scale,
                                 void* ws, void* stream) {
  ZS_REQUIRE(qkv && dO && dqkv && ws && B >= 0 && heads > 0, "zs_mha_bwd_tc_f32: null pointer / bad sizes");
  ZS_REQUIRE(T >= 1 && T <= 208 && (hd == 32 || hd == 64), "zs_mha_bwd_tc_f32: T must be in [1, 208] and the head dim 32 or 64");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dO) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0,
             "zs_mha_bwd_tc_f32: qkv / dO / dqkv / ws must be 16-byte aligned");
  if (B == 0) return ZS_OK;
  MhaBwdParams p{qkv, dO, dqkv, reinterpret_cast<float*>(ws), B, T, heads, hd, scale};
  dim3 grid(B * heads, (T + 127) / 128);
  cudaStream_t st = as_stream(stream);
  if (hd == 64) {
    ZS_CUDA_CALL(cudaFuncSetAttribute(mha_bwd_q_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
    ZS_CUDA_CALL(cudaFuncSetAttribute(mha_bwd_kv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
    mha_bwd_q_kernel<64><<<grid, MT_THREADS, MB_SMEM, st>>>(p);
    ZS_CUDA_CHECK_LAUNCH("zs_mha_bwd_tc_f32");
    mha_bwd_kv_kernel<64><<<grid, MT_THREADS, MB_SMEM, st>>>(p);
  } else {
    ZS_CUDA_CALL(cudaFuncSetAttribute(mha_bwd_q_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
    ZS_CUDA_CALL(cudaFuncSetAttribute(mha_bwd_kv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
    mha_bwd_q_kernel<32><<<grid, MT_THREADS, MB_SMEM, st>>>(p);
    ZS_CUDA_CHECK_LAUNCH("zs_mha_bwd_tc_f32");
    mha_bwd_kv_kernel<32><<<grid, MT_THREADS, MB_SMEM, st>>>(p);
  }
  ZS_CUDA_CHECK_LAUNCH("zs_mha_bwd_tc_f32");
  return ZS_OK;
}

extern "C" size_t zs_point_attention_bwd_tc_ws_bytes(int B, int P, int heads) {
  return B > 0 && P > 0 && heads > 0 ? (size_t)B * heads * P * 3 * sizeof(float) : 0;
}

extern "C" int zs_point_attention_bwd_tc_f32(const float* qkv_p, const float* k_lat, const float* v_lat, int ld_lat, const float* dO,
                                             float* dqkv_p, float* dk_lat, float* dv_lat, int ld_dlat, int B, int P, int L, int heads,
                                             int hd, float scale, void* ws, void* stream) {
  ZS_REQUIRE(qkv_p && k_lat && v_lat && dO && dqkv_p && dk_lat && dv_lat && ws, "zs_point_attention_bwd_tc_f32: null pointer");
  ZS_REQUIRE(hd == 32 && B > 0 && P > 0 && L > 0 && L <= 208 && heads > 0, "zs_point_attention_bwd_tc_f32: head dim 32, 1 <= L <= 208");
  ZS_REQUIRE((ld_lat & 3) == 0 && (ld_dlat & 3) == 0, "zs_point_attention_bwd_tc_f32: latent row strides must be multiples of 4");
  const void* ptrs[] = {qkv_p, k_lat, v_lat, dO, dqkv_p, dk_lat, dv_lat, ws};
  for (const void* q : ptrs) ZS_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, "zs_point_attention_bwd_tc_f32: 16-byte alignment");
  PaBwdParams p{qkv_p, k_lat, v_lat, ld_lat, dO, dqkv_p, dk_lat, dv_lat, ld_dlat, reinterpret_cast<float*>(ws), B, P, L, heads, scale};
  cudaStream_t st = as_stream(stream);
  ZS_CUDA_CALL(cudaFuncSetAttribute(pa_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
  ZS_CUDA_CALL(cudaFuncSetAttribute(pa_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
  pa_bwd_q_kernel<<<dim3(B * heads, (P + 127) / 128), MT_THREADS, MB_SMEM, st>>>(p);
  ZS_CUDA_CHECK_LAUNCH("zs_point_attention_bwd_tc_f32");
  pa_bwd_kv_kernel<<<dim3(B * heads, (L + 127) / 128), MT_THREADS, MB_SMEM, st>>>(p);
  ZS_CUDA_CHECK_LAUNCH("zs_point_attention_bwd_tc_f32");
  return ZS_OK;
}

