// Chained tcgen05 kernels of the implicit decoder: consecutive layers of a 128-point tile stay on chip.
//
//   zs_chain_mlp_fwd : x <- x + fc2( GELU( fc1( LayerNorm(x) ) ) )           (timm Mlp inside ImplFuncBlock,
//                                                                              model/shape/implicit.py:94-108)
//   zs_chain_occ_fwd : logit = MLPBlocks([xyz, LayerNorm(x)])                 (final norm + impl_mlp,
//                                                                              implicit.py:275,168-184)
// Both are instances of one dataflow.  Per 128-row tile a sequence of GEMM steps D[128 x 256] (+)= A * W^T
// runs on the tensor pipe; every A operand is a stream of 64-wide K-chunks in one of two smem rings:
//   ring L : chunks converted from global fp32 (LayerNorm applied on the fly) by 4 loader warps,
//   ring E : chunks written by the 8 epilogue warps straight from the previous step's TMEM accumulator
//            (bias + GELU / Softplus + bf16 hi/lo split), i.e. layer l+1 consumes layer l without touching HBM/L2.
// Accumulators ping-pong between the two 256-column halves of TMEM, so the epilogue of step s overlaps the MMAs
// of step s+1 chunk by chunk.  Weights stream through a 3-slot ring of 32 KB tiles (cp.async.bulk / TMA 1-D) from
// a blob that lists the (hi, lo) tiles in exactly the order the MMA warp consumes them.
// Split-bf16 arithmetic as in gemm_tc.cu (precision 0: Ah*Wh + Al*Wh + Ah*Wl; 1: Ah*Wh).
#include "common.cuh"
#include "tc_common.cuh"

namespace zs {
using namespace tc;

constexpr int CT_THREADS = 448;                 // warps: 0-3 loader, 4-11 epilogue, 12 MMA, 13 W loader
constexpr int CT_TILE_BYTES = 256 * 64 * 2;     // one W tile (hi or lo): 256 rows x 64 bf16 = 32 KB
constexpr int CT_A_HALF = 128 * 64 * 2;         // A chunk hi (or lo): 16 KB
constexpr int CT_WSLOTS = 3, CT_LSLOTS = 2, CT_ESLOTS = 2;
constexpr int CT_OFF_W = 0;
constexpr int CT_OFF_L = CT_OFF_W + CT_WSLOTS * CT_TILE_BYTES;          // 96 KB
constexpr int CT_OFF_E = CT_OFF_L + CT_LSLOTS * 2 * CT_A_HALF;          // +64 KB
constexpr int CT_OFF_BAR = CT_OFF_E + CT_ESLOTS * 2 * CT_A_HALF;        // +64 KB = 224 KB
constexpr int CT_SMEM_USED = CT_OFF_BAR + 256 + 2048;   // barriers + 2 KB of per-kernel scratch (row partials / self scores)
constexpr int CT_SMEM = 227 * 1024;                     // the whole opt-in window: 768 B of slack for a base that is not 1024-aligned

struct ChainParams {
  float* x; int ldx;             // [M, 256] residual stream (MLP: in/out; OCC: in)
  const float* points;           // OCC: [M,3]
  int M;
  const float* ln_w; const float* ln_b; float ln_eps;
  const uint8_t* blob;           // weight tiles in consumption order: [pair][hi 32K | lo 32K]
  const float* bias;             // MLP: b1[1024] ; OCC: [8][256]
  const float* bias2;            // MLP: b2[256]  ; OCC: w8[256]
  float b8;                      // OCC: last-layer bias
  float* out;                    // OCC: [M]
  int apply_sigmoid;
  int precision;
  unsigned long long* trace;     // debug: per-role (tag, clock64) events of CTA 0 (nullptr in production)
  // chain_lin_kernel: out[M, 256*n_tiles] = LN?(x) W^T + bias (+ res)
  const float* res; int ldres; int ldo; int n_tiles; int do_ln;
  // chain_qkvattn_kernel: latent key / value operand images of the image, softmax scale, per-GEMM pass policy
  const uint8_t* kblob; const uint8_t* vblob; int n_keys; float scale; int flags;
  // chain_pmlp_kernel: tile-blocked attention output, packed proj weight (4 K-chunk pairs), proj bias
  const float* a_blk; const uint8_t* pblob; const float* bias_p;
  // "points mode" of the first block (chain_qkvattn2_kernel<true> with flags & 64, chain_pmlp_kernel with pp != nullptr): the
  // residual stream entering the block is x = LinearProj3D(points) (model/shape/implicit.py:128-131) and is RECOMPUTED from the
  // 12-byte points instead of being read (and first written by a separate launch): pp = [4][256] = w[:,0] | w[:,1] | w[:,2] | bias,
  // pp_stat = the 14 moments of those four rows that give the LayerNorm mean / variance of a row in closed form
  const float* pp; const float* pp_stat;
};

struct Bars {
  uint32_t base;
  __device__ uint32_t wfull(int i) const { return base + 8u * i; }
  __device__ uint32_t wempty(int i) const { return base + 24u + 8u * i; }
  __device__ uint32_t lfull(int i) const { return base + 48u + 8u * i; }
  __device__ uint32_t lempty(int i) const { return base + 64u + 8u * i; }
  __device__ uint32_t efull(int i) const { return base + 80u + 8u * i; }
  __device__ uint32_t eempty(int i) const { return base + 96u + 8u * i; }
  __device__ uint32_t tfull(int i) const { return base + 112u + 8u * i; }
  __device__ uint32_t tempty(int i) const { return base + 128u + 8u * i; }
  __device__ uint32_t tmem_slot() const { return base + 144u; }
};

struct Ring {   // index/phase walker of an n-slot mbarrier ring
  int idx = 0; uint32_t phase = 0; int n;
  __device__ explicit Ring(int n_) : n(n_) {}
  __device__ void advance() { if (++idx == n) { idx = 0; phase ^= 1; } }
};

constexpr int CT_TRACE_EVENTS = 512;   // per role
__device__ __forceinline__ void trace_ev(unsigned long long* tr, int role, int& n, int tag) {
  if (tr != nullptr && blockIdx.x == 0 && n < CT_TRACE_EVENTS) tr[role * CT_TRACE_EVENTS + n++] = ((unsigned long long)clock64() << 8) | (unsigned)tag;
}

__device__ __forceinline__ float fast_softplus100(float x) {
  // torch Softplus(beta=100, threshold=20) = x if 100x > 20 else log1p(exp(100x))/100
  // Accurate to ~1e-7 relative with ~20 instructions: log1p(e^t) = max(t,0) + log1p(e^-|t|); log1p(u), u in (0,1], via
  // 2*atanh(u/(2+u)) (odd series, z <= 1/3) -- avoids the cancellation of log(1+u) for small u.
  const float bx = 100.0f * x;
  if (bx > 20.0f) return x;
  const float u = fast_ex2(-fabsf(bx) * 1.4426950408889634f);
  const float z = u * fast_rcp(2.0f + u);
  const float z2 = z * z;
  const float l = 2.0f * z * (1.0f + z2 * (0.33333334f + z2 * (0.2f + z2 * (0.14285715f + z2 * (0.11111111f + z2 * (0.09090909f + z2 * 0.07692308f))))));
  return (fmaxf(bx, 0.0f) + l) * 0.01f;
}

// exact-erf GELU (torch.nn.GELU default) with erf from Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7, i.e. the
// same order as the rounding of torch's own fp32 erff path: measured 4.7e-7 max abs on GELU vs fp64, torch fp32 1.2e-6):
// ~17 instructions (1 MUFU.RCP + 1 MUFU.EX2) instead of ~35 for erff.
__device__ __forceinline__ float fast_gelu_erf(float x) {
  const float z = x * 0.70710678118654752440f;
  const float az = fabsf(z);
  const float t = fast_rcp(fmaf(0.3275911f, az, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = fast_ex2(-az * az * 1.4426950408889634f);
  const float erf_abs = fmaf(-poly, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, z));
}

// The same two activations on PAIRS (FFMA2 / FMUL2 / FADD2): the polynomial parts cost half the issue slots; only
// |.|, MUFU, copysign / max / select stay per element.  ~10 (GELU) and ~14 (Softplus) instructions per element.
__device__ __forceinline__ float2 fast_gelu_erf2(float2 x) {
  const float2 z = mul2(x, bc2(0.70710678118654752440f));
  const float2 az = make_float2(fabsf(z.x), fabsf(z.y));
  const float2 den = fma2(bc2(0.3275911f), az, bc2(1.0f));
  const float2 t = make_float2(fast_rcp(den.x), fast_rcp(den.y));
  float2 np = fma2(bc2(-1.061405429f), t, bc2(1.453152027f));   // the NEGATED polynomial (exact sign symmetry of fma)
  np = fma2(np, t, bc2(-1.421413741f));
  np = fma2(np, t, bc2(0.284496736f));
  np = fma2(np, t, bc2(-0.254829592f));
  np = mul2(np, t);
  const float2 ea = mul2(mul2(az, az), bc2(-1.4426950408889634f));
  const float2 e = make_float2(fast_ex2(ea.x), fast_ex2(ea.y));
  const float2 erf_abs = fma2(np, e, bc2(1.0f));
  const float2 er = make_float2(copysignf(erf_abs.x, z.x), copysignf(erf_abs.y, z.y));
  const float2 hx = mul2(x, bc2(0.5f));
  return fma2(hx, er, hx);
}

// Softplus(beta = 100, threshold = 20) on pairs: max(x, 0) + ln2 / 100 * lg2(1 + 2^(-|100 x| log2 e)).  The activation is ADDED to
// other O(0.1) terms downstream, so what matters is its ABSOLUTE error: 1 + u rounds u to 6e-8 and lg2.approx adds ~2e-7, i.e. < 3e-9
// after the 1 / 100 -- three orders below the fp16x3 operand split.  (fast_softplus100 above keeps the relative-accuracy series; this
// form is 6 instructions + 2 MUFU per element instead of 14 + 2.)
__device__ __forceinline__ float2 fast_softplus100_2(float2 x) {
  const float2 bx = mul2(x, bc2(100.0f));
  const float2 ua = mul2(make_float2(-fabsf(bx.x), -fabsf(bx.y)), bc2(1.4426950408889634f));
  const float2 u1 = add2(make_float2(fast_ex2(ua.x), fast_ex2(ua.y)), bc2(1.0f));
  const float2 l = make_float2(fast_lg2(u1.x), fast_lg2(u1.y));
  const float2 o = fma2(l, bc2(0.0069314718055994531f), make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
  return make_float2(bx.x > 20.0f ? x.x : o.x, bx.y > 20.0f ? x.y : o.y);
}

// ---------------------------------------------------------------------------------------------------------------
// common prologue: barriers + TMEM
__device__ __forceinline__ uint32_t chain_setup(const Bars& B, uint8_t* smem_gen, uint32_t smem_base, int warp, int mma_warp = 12,
                                                int loader_threads = 128) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < CT_WSLOTS; ++i) { mbar_init(B.wfull(i), 1); mbar_init(B.wempty(i), 1); }
    for (int i = 0; i < CT_LSLOTS; ++i) { mbar_init(B.lfull(i), loader_threads); mbar_init(B.lempty(i), 1); }
    for (int i = 0; i < CT_ESLOTS; ++i) { mbar_init(B.efull(i), 256); mbar_init(B.eempty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(B.tfull(i), 1); mbar_init(B.tempty(i), 256); }
    fence_mbar_init();
  }
  if (warp == mma_warp) tmem_alloc(B.tmem_slot(), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(smem_gen + (B.tmem_slot() - smem_base));
}

// MMA warp: consume one A chunk (hi at a_addr, lo at a_addr + 16K) against the next weight tile pair
__device__ __forceinline__ void mma_chunk(const Bars& B, uint32_t smem_base, Ring& wr, uint32_t a_addr, uint32_t d_tmem,
                                          bool first, bool split, uint32_t a_empty_bar, unsigned long long* tr = nullptr,
                                          int* tn = nullptr) {
  const uint32_t idesc = umma_idesc_f16(128, 256);
  const uint64_t a_hi = umma_desc_sw128(a_addr), a_lo = umma_desc_sw128(a_addr + CT_A_HALF);
  mbar_wait(B.wfull(wr.idx), wr.phase);
  if (tr) trace_ev(tr, 0, *tn, 3);
  tc_fence_after();
  {
    const uint64_t w = umma_desc_sw128(smem_base + CT_OFF_W + wr.idx * CT_TILE_BYTES);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      umma_bf16(d_tmem, a_hi + 2 * k, w + 2 * k, idesc, (!first || k > 0) ? 1u : 0u);
      if (split) umma_bf16(d_tmem, a_lo + 2 * k, w + 2 * k, idesc, 1u);
    }
    umma_commit(B.wempty(wr.idx));
    wr.advance();
  }
  if (tr) trace_ev(tr, 0, *tn, 4);
  if (split) {
    mbar_wait(B.wfull(wr.idx), wr.phase);
    if (tr) trace_ev(tr, 0, *tn, 5);
    tc_fence_after();
    const uint64_t w = umma_desc_sw128(smem_base + CT_OFF_W + wr.idx * CT_TILE_BYTES);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_hi + 2 * k, w + 2 * k, idesc, 1u);
    umma_commit(B.wempty(wr.idx));
    wr.advance();
  }
  if (a_empty_bar != 0u) umma_commit(a_empty_bar);
  if (tr) trace_ev(tr, 0, *tn, 6);
}

// W loader warp (one lane): stream `pairs` (hi,lo) tile pairs of the blob through the W ring
__device__ __forceinline__ void w_stream(const Bars& B, uint32_t smem_base, Ring& wr, const uint8_t* blob, int pairs, bool split) {
  for (int i = 0; i < pairs; ++i) {
    const uint8_t* src = blob + (size_t)i * 2 * CT_TILE_BYTES;
    for (int h = 0; h < (split ? 2 : 1); ++h) {
      mbar_wait(B.wempty(wr.idx), wr.phase ^ 1);
      mbar_arrive_expect_tx(B.wfull(wr.idx), CT_TILE_BYTES);
      bulk_g2s(smem_base + CT_OFF_W + wr.idx * CT_TILE_BYTES, src + (size_t)h * CT_TILE_BYTES, CT_TILE_BYTES, B.wfull(wr.idx));
      wr.advance();
    }
  }
}

// loader thread: write 64 values (already in registers as 16 float4) as a swizzled (hi,lo) chunk row
__device__ __forceinline__ void store_chunk_row(uint8_t* slot, int r, const float4* buf, bool split) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 x0 = buf[2 * c], x1 = buf[2 * c + 1];
    uint4 hi, lo;
    split_f16x2(x0.x, x0.y, hi.x, lo.x);
    split_f16x2(x0.z, x0.w, hi.y, lo.y);
    split_f16x2(x1.x, x1.y, hi.z, lo.z);
    split_f16x2(x1.z, x1.w, hi.w, lo.w);
    const uint32_t off = swizzle128_offset(r, c);
    *reinterpret_cast<uint4*>(slot + off) = hi;
    if (split) *reinterpret_cast<uint4*>(slot + CT_A_HALF + off) = lo;
  }
}

// ---- coalesced loader (4 warps; warp w owns tile rows 32w .. 32w+31) ------------------------------------------
// A thread-per-row loader reads 16 B out of every 1 KB row per instruction (32 sectors per request, L1 thrashing):
// measured 44k cycles per tile for the LayerNorm statistics alone.  Here a warp walks its rows with the lanes along
// the columns, so every request is one or two fully used 256..1024-byte segments.
// LayerNorm statistics of the warp's 32 rows (two-pass variance on register-held data); lane L receives
// (sc, sh) = (rstd, -mean * rstd) of row m0 + L, or (0, 0) past M so that padded rows normalise to zero.
__device__ __forceinline__ void ln_load4(const float* __restrict__ x, int ldx, int m0, int M, int lane, float4* aa, float4* bb) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int m = m0 + u;
    const float4* p4 = reinterpret_cast<const float4*>(x + (int64_t)(m < M ? m : 0) * ldx);
    aa[u] = __ldg(p4 + lane);
    bb[u] = __ldg(p4 + 32 + lane);
  }
}
__device__ __forceinline__ void ln_reduce4(const float4* a, const float4* b, int m0, int i0, int M, float eps, int lane, float& sc, float& sh) {
  float mean[4], q[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) mean[u] = ((a[u].x + a[u].y) + (a[u].z + a[u].w)) + ((b[u].x + b[u].y) + (b[u].z + b[u].w));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int u = 0; u < 4; ++u) mean[u] += __shfl_xor_sync(0xffffffffu, mean[u], o);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    mean[u] *= (1.0f / 256.0f);
    const float d0 = a[u].x - mean[u], d1 = a[u].y - mean[u], d2 = a[u].z - mean[u], d3 = a[u].w - mean[u];
    const float d4 = b[u].x - mean[u], d5 = b[u].y - mean[u], d6 = b[u].z - mean[u], d7 = b[u].w - mean[u];
    q[u] = ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3)) + ((d4 * d4 + d5 * d5) + (d6 * d6 + d7 * d7));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] += __shfl_xor_sync(0xffffffffu, q[u], o);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float rstd = rsqrtf(q[u] * (1.0f / 256.0f) + eps);
    if (lane == i0 + u && m0 + i0 + u < M) { sc = rstd; sh = -mean[u] * rstd; }
  }
}
__device__ __forceinline__ void warp_ln_stats(const float* __restrict__ x, int ldx, int m0, int M, float eps, int lane,
                                              float& sc, float& sh) {
  sc = 0.f; sh = 0.f;
  // 4 rows per step in two register sets: the next step's 8 loads are in flight while the current rows are reduced
  // (rolled loop: a fully unrolled version measured slower, the role code must stay small for the instruction cache)
  float4 a0[4], b0[4], a1[4], b1[4];
  ln_load4(x, ldx, m0, M, lane, a0, b0);
#pragma unroll 1
  for (int i0 = 0; i0 < 32; i0 += 8) {
    ln_load4(x, ldx, m0 + i0 + 4, M, lane, a1, b1);
    ln_reduce4(a0, b0, m0, i0, M, eps, lane, sc, sh);
    if (i0 + 8 < 32) ln_load4(x, ldx, m0 + i0 + 8, M, lane, a0, b0);
    ln_reduce4(a1, b1, m0, i0 + 4, M, eps, lane, sc, sh);
  }
}

// L2 prefetch of 256 fp32 columns [col0, col0+256) of the warp's 32 rows m0..m0+31 (8 x 128-byte lines per row, one line per
// lane per instruction).  The loaders issue it one tile ahead: with a single 32 KB chunk in flight per SM the HBM stream
// is latency-bound (Little: 6.4 TB/s x ~1.5 us needs ~65 KB in flight per SM); afterwards every demand load is an L2 hit.
__device__ __forceinline__ void prefetch_rows_l2(const float* __restrict__ x, int ldx, int m0, int M, int col0, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + 4 * i + (lane >> 3);
    if (m < M) asm volatile("prefetch.global.L2 [%0];" ::"l"(x + (int64_t)m * ldx + col0 + (lane & 7) * 32));
  }
}

// 64 columns [col0, col0+64) of the warp's 32 rows: instruction j covers rows 2j, 2j+1 (16 lanes x float4 each)
__device__ __forceinline__ void fetch_chunk_co(const float* __restrict__ x, int ldx, int m0, int M, int col0, int lane, float4* buf) {
  const int sub = lane >> 4, q = lane & 15;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int m = m0 + 2 * j + sub;
    buf[j] = m < M ? __ldg(reinterpret_cast<const float4*>(x + (int64_t)m * ldx + col0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// normalise (v * sc[row] + sh[row], optional affine g/b of this lane's 4 columns), split, store into the swizzled chunk
__device__ __forceinline__ void store_chunk_co(uint8_t* slot, int w, int lane, const float4* buf, bool normalise, float sc, float sh,
                                               const float* __restrict__ g, const float* __restrict__ b, bool split) {
  const int sub = lane >> 4, q = lane & 15;
  float4 gg = make_float4(1.f, 1.f, 1.f, 1.f), bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g != nullptr) { gg = __ldg(reinterpret_cast<const float4*>(g) + q); bb = __ldg(reinterpret_cast<const float4*>(b) + q); }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int rl = 2 * j + sub;
    float2 v0 = make_float2(buf[j].x, buf[j].y), v1 = make_float2(buf[j].z, buf[j].w);
    if (normalise) {
      const float s = __shfl_sync(0xffffffffu, sc, rl), h = __shfl_sync(0xffffffffu, sh, rl);
      v0 = fma2(v0, bc2(s), bc2(h)); v1 = fma2(v1, bc2(s), bc2(h));
      if (g != nullptr) {
        v0 = fma2(v0, make_float2(gg.x, gg.y), make_float2(bb.x, bb.y));
        v1 = fma2(v1, make_float2(gg.z, gg.w), make_float2(bb.z, bb.w));
      }
    }
    uint2 hi, lo;
    split_f16x2(v0.x, v0.y, hi.x, lo.x);
    split_f16x2(v1.x, v1.y, hi.y, lo.y);
    const uint32_t off = swizzle128_offset(32 * w + rl, q >> 1) + ((q & 1) << 3);
    *reinterpret_cast<uint2*>(slot + off) = hi;
    if (split) *reinterpret_cast<uint2*>(slot + CT_A_HALF + off) = lo;
  }
}

// epilogue thread: 32 accumulator columns -> act(d + bias) -> (hi,lo) bf16 -> its 4 16-byte chunks of a ring-E row
template <int ACT>
__device__ __forceinline__ void epi_to_ring(uint8_t* slot, int row, int hsel, const uint32_t (&rr)[32], const float* __restrict__ bias,
                                            bool split) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c * 8 + 4));
    const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
    float2 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = add2(make_float2(__uint_as_float(rr[c * 8 + 2 * j]), __uint_as_float(rr[c * 8 + 2 * j + 1])), bb[j]);
      v[j] = ACT == ZS_ACT_GELU ? fast_gelu_erf2(t) : fast_softplus100_2(t);
    }
    uint4 hi, lo;
    split_f16x2(v[0].x, v[0].y, hi.x, lo.x);
    split_f16x2(v[1].x, v[1].y, hi.y, lo.y);
    split_f16x2(v[2].x, v[2].y, hi.z, lo.z);
    split_f16x2(v[3].x, v[3].y, hi.w, lo.w);
    const uint32_t off = swizzle128_offset(row, hsel * 4 + c);
    *reinterpret_cast<uint4*>(slot + off) = hi;
    if (split) *reinterpret_cast<uint4*>(slot + CT_A_HALF + off) = lo;
  }
}

// ===============================================================================================================
// x <- x + fc2(GELU(fc1(LN(x))))      blob order: for g in 0..3: fc1[g] (4 pairs), fc2[:, g] (4 pairs)
__global__ void __launch_bounds__(CT_THREADS, 1) chain_mlp_kernel(ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  Bars B{smem_base + CT_OFF_BAR};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = p.precision == 0;
  const uint32_t tmem_base = chain_setup(B, smem_gen, smem_base, warp);
  const int n_tiles = (p.M + 127) / 128;

  if (warp < 4) {
    // ---------------- loader: LN(x) chunks, 4 groups x 4 chunks per tile ----------------
    const int r = threadIdx.x;
    Ring lr(CT_LSLOTS);
    float4 buf[16];
    int tn = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = t * 128 + warp * 32;
      float sc, sh;
      if (r == 0) trace_ev(p.trace, 1, tn, 14);
      if (t + (int)gridDim.x < n_tiles) prefetch_rows_l2(p.x, p.ldx, m0 + (int)gridDim.x * 128, p.M, 0, lane);
      warp_ln_stats(p.x, p.ldx, m0, p.M, p.ln_eps, lane, sc, sh);
      fetch_chunk_co(p.x, p.ldx, m0, p.M, 0, lane, buf);
      if (r == 0) trace_ev(p.trace, 1, tn, 15);
      for (int i = 0; i < 16; ++i) {
        const int kc = i & 3;
        if (r == 0) trace_ev(p.trace, 1, tn, 10);
        mbar_wait(B.lempty(lr.idx), lr.phase ^ 1);
        if (r == 0) trace_ev(p.trace, 1, tn, 11);
        store_chunk_co(smem_gen + CT_OFF_L + lr.idx * 2 * CT_A_HALF, warp, lane, buf, true, sc, sh,
                       p.ln_w ? p.ln_w + kc * 64 : nullptr, p.ln_b ? p.ln_b + kc * 64 : nullptr, split);
        fence_proxy_async_smem();
        mbar_arrive(B.lfull(lr.idx));
        if (r == 0) trace_ev(p.trace, 1, tn, 12);
        lr.advance();
        if (i + 1 < 16) fetch_chunk_co(p.x, p.ldx, m0, p.M, ((i + 1) & 3) * 64, lane, buf);
        if (r == 0) trace_ev(p.trace, 1, tn, 13);
      }
    }
  } else if (warp == 13) {
    if (lane == 0) {
      Ring wr(CT_WSLOTS);
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) w_stream(B, smem_base, wr, p.blob, 32, split);
    }
  } else if (warp == 12) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      Ring wr(CT_WSLOTS), lr(CT_LSLOTS), er(CT_ESLOTS);
      uint32_t te_phase[2] = {0, 0};
      int tn = 0;
      const uint32_t d0 = tmem_base, d1 = tmem_base + 256;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int g = 0; g < 4; ++g) {
          trace_ev(p.trace, 0, tn, 7);
          mbar_wait(B.tempty(0), te_phase[0] ^ 1); te_phase[0] ^= 1;
          tc_fence_after();
          for (int kc = 0; kc < 4; ++kc) {
            trace_ev(p.trace, 0, tn, 1);
            mbar_wait(B.lfull(lr.idx), lr.phase);
            trace_ev(p.trace, 0, tn, 2);
            tc_fence_after();
            mma_chunk(B, smem_base, wr, smem_base + CT_OFF_L + lr.idx * 2 * CT_A_HALF, d0, kc == 0, split, B.lempty(lr.idx),
                      p.trace, &tn);
            lr.advance();
          }
          umma_commit(B.tfull(0));
          if (g == 0) { mbar_wait(B.tempty(1), te_phase[1] ^ 1); te_phase[1] ^= 1; tc_fence_after(); }
          for (int kc = 0; kc < 4; ++kc) {
            trace_ev(p.trace, 0, tn, 8);
            mbar_wait(B.efull(er.idx), er.phase);
            trace_ev(p.trace, 0, tn, 9);
            tc_fence_after();
            mma_chunk(B, smem_base, wr, smem_base + CT_OFF_E + er.idx * 2 * CT_A_HALF, d1, g == 0 && kc == 0, split, B.eempty(er.idx),
                      p.trace, &tn);
            er.advance();
          }
        }
        umma_commit(B.tfull(1));
      }
    }
  } else {
    // ---------------- epilogue: 8 warps; lane quarter q, column half hsel ----------------
    const int e = warp - 4, q = e & 3, hsel = e >> 2;
    const int row = q * 32 + lane;
    Ring er(CT_ESLOTS);
    uint32_t tf_phase[2] = {0, 0};
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int tn = 0;
    const bool tr0 = (warp == 4 && lane == 0);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m = t * 128 + row;
      for (int g = 0; g < 4; ++g) {
        if (tr0) trace_ev(p.trace, 2, tn, 20);
        mbar_wait(B.tfull(0), tf_phase[0]); tf_phase[0] ^= 1;
        if (tr0) trace_ev(p.trace, 2, tn, 21);
        tc_fence_after();
        for (int c = 0; c < 4; ++c) {
          uint32_t rr[32];
          tmem_ld_32x32(tmem_base + lane_off + c * 64 + hsel * 32, rr);
          tmem_ld_wait();
          if (tr0) trace_ev(p.trace, 2, tn, 22);
          mbar_wait(B.eempty(er.idx), er.phase ^ 1);
          if (tr0) trace_ev(p.trace, 2, tn, 23);
          epi_to_ring<ZS_ACT_GELU>(smem_gen + CT_OFF_E + er.idx * 2 * CT_A_HALF, row, hsel, rr,
                                   p.bias + g * 256 + c * 64 + hsel * 32, split);
          fence_proxy_async_smem();
          mbar_arrive(B.efull(er.idx));
          if (tr0) trace_ev(p.trace, 2, tn, 24);
          er.advance();
        }
        tc_fence_before();
        mbar_arrive(B.tempty(0));
      }
      mbar_wait(B.tfull(1), tf_phase[1]); tf_phase[1] ^= 1;
      tc_fence_after();
      // x += fc2 + b2, transposed through smem (ring E is idle here: every fc2 MMA of the tile has completed) so that
      // global memory sees whole 256-byte row segments instead of 16 bytes per row per request.
      float* tr_buf = reinterpret_cast<float*>(smem_gen + CT_OFF_E);   // [128][68] fp32 (padded rows: conflict-free both ways)
      const int sub = lane >> 4, q4 = lane & 15;
      for (int c = 0; c < 4; ++c) {
        uint32_t rr[32];
        tmem_ld_32x32(tmem_base + 256 + lane_off + c * 64 + hsel * 32, rr);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(tr_buf + row * 68 + hsel * 32 + 4 * j) =
              make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias2 + c * 64) + q4);
        // all 8 residual loads first (they alias the stores below, so the compiler would otherwise serialise
        // load -> store -> load: 8 dependent L2 round trips per column group)
        float4 xin[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int mm = t * 128 + i * 16 + e * 2 + sub;
          xin[i] = *(reinterpret_cast<const float4*>(p.x + (int64_t)(mm < p.M ? mm : 0) * p.ldx + c * 64) + q4);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 16 + e * 2 + sub;
          const int mm = t * 128 + rl;
          const float4 a = *reinterpret_cast<const float4*>(tr_buf + rl * 68 + 4 * q4);
          float4 xv = xin[i];
          xv.x += a.x + bv.x; xv.y += a.y + bv.y; xv.z += a.z + bv.z; xv.w += a.w + bv.w;
          if (mm < p.M) *(reinterpret_cast<float4*>(p.x + (int64_t)mm * p.ldx + c * 64) + q4) = xv;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      tc_fence_before();
      mbar_arrive(B.tempty(1));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// 16-row variants of the coalesced loader helpers (warp w owns tile rows 16w .. 16w+15): with 8 loader warps a chunk
// conversion is split over two warps per scheduler (one warp converting a chunk alone is latency-bound at ~0.3 IPC)
__device__ __forceinline__ void prefetch_rows16_l2(const float* __restrict__ x, int ldx, int m0, int M, int col0, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + 4 * i + (lane >> 3);
    if (m < M) asm volatile("prefetch.global.L2 [%0];" ::"l"(x + (int64_t)m * ldx + col0 + (lane & 7) * 32));
  }
}
__device__ __forceinline__ void warp_ln_stats16(const float* __restrict__ x, int ldx, int m0, int M, float eps, int lane,
                                                float& sc, float& sh) {
  sc = 0.f; sh = 0.f;
  float4 a0[4], b0[4];          // one register set: the 96-register budget of a 576-thread CTA has no room for two
#pragma unroll 1
  for (int i0 = 0; i0 < 16; i0 += 4) {
    ln_load4(x, ldx, m0 + i0, M, lane, a0, b0);
    ln_reduce4(a0, b0, m0, i0, M, eps, lane, sc, sh);
  }
}
__device__ __forceinline__ void fetch_chunk16(const float* __restrict__ x, int ldx, int m0, int M, int col0, int lane, float4* buf) {
  const int sub = lane >> 4, q = lane & 15;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int m = m0 + 2 * j + sub;
    buf[j] = m < M ? __ldg(reinterpret_cast<const float4*>(x + (int64_t)m * ldx + col0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void store_chunk16(uint8_t* slot, int w, int lane, const float4* buf, bool normalise, float sc, float sh, bool split) {
  const int sub = lane >> 4, q = lane & 15;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int rl = 2 * j + sub;
    float2 v0 = make_float2(buf[j].x, buf[j].y), v1 = make_float2(buf[j].z, buf[j].w);
    if (normalise) {
      const float s = __shfl_sync(0xffffffffu, sc, rl), h = __shfl_sync(0xffffffffu, sh, rl);
      v0 = fma2(v0, bc2(s), bc2(h)); v1 = fma2(v1, bc2(s), bc2(h));
    }
    uint2 hi, lo;
    split_f16x2(v0.x, v0.y, hi.x, lo.x);
    split_f16x2(v1.x, v1.y, hi.y, lo.y);
    const uint32_t off = swizzle128_offset(16 * w + rl, q >> 1) + ((q & 1) << 3);
    *reinterpret_cast<uint2*>(slot + off) = hi;
    if (split) *reinterpret_cast<uint2*>(slot + CT_A_HALF + off) = lo;
  }
}

// points mode: the chunk's values are LayerNorm(LinearProj3D(point)) of the rows' points, formed as four packed FMAs per pair of
// columns with the row's LayerNorm scale folded into the point (w0 (x s) + w1 (y s) + w2 (z s) + (b s + h)); lane L < 16 holds the
// point / LayerNorm factors of row m0 + L
__device__ __forceinline__ void store_chunk16_pts(uint8_t* slot, int w, int lane, int kc, float px, float py, float pz, float sc, float sh,
                                                  const float* __restrict__ pp, bool split) {
  const int sub = lane >> 4, q = lane & 15;
  const float4 w0 = __ldg(reinterpret_cast<const float4*>(pp + kc * 64) + q), w1 = __ldg(reinterpret_cast<const float4*>(pp + 256 + kc * 64) + q);
  const float4 w2 = __ldg(reinterpret_cast<const float4*>(pp + 512 + kc * 64) + q), bb = __ldg(reinterpret_cast<const float4*>(pp + 768 + kc * 64) + q);
  const float xs_l = px * sc, ys_l = py * sc, zs_l = pz * sc;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int rl = 2 * j + sub;
    const float2 x = bc2(__shfl_sync(0xffffffffu, xs_l, rl)), y = bc2(__shfl_sync(0xffffffffu, ys_l, rl)), z = bc2(__shfl_sync(0xffffffffu, zs_l, rl));
    const float2 s = bc2(__shfl_sync(0xffffffffu, sc, rl)), h = bc2(__shfl_sync(0xffffffffu, sh, rl));
    const float2 v0 = fma2(make_float2(w2.x, w2.y), z, fma2(make_float2(w1.x, w1.y), y, fma2(make_float2(w0.x, w0.y), x, fma2(make_float2(bb.x, bb.y), s, h))));
    const float2 v1 = fma2(make_float2(w2.z, w2.w), z, fma2(make_float2(w1.z, w1.w), y, fma2(make_float2(w0.z, w0.w), x, fma2(make_float2(bb.z, bb.w), s, h))));
    uint2 hi, lo;
    split_f16x2(v0.x, v0.y, hi.x, lo.x);
    split_f16x2(v1.x, v1.y, hi.y, lo.y);
    const uint32_t off = swizzle128_offset(16 * w + rl, q >> 1) + ((q & 1) << 3);
    *reinterpret_cast<uint2*>(slot + off) = hi;
    if (split) *reinterpret_cast<uint2*>(slot + CT_A_HALF + off) = lo;
  }
}
// LayerNorm factors (rstd, -mean * rstd) of LinearProj3D(point) in closed form from the moments of the projection's rows
__device__ __forceinline__ void pts_ln_factors(const float* __restrict__ st, float x, float y, float z, float eps, float& sc, float& sh) {
  const float mean = fmaf(x, __ldg(st), fmaf(y, __ldg(st + 1), fmaf(z, __ldg(st + 2), __ldg(st + 3))));
  float var = fmaf(x * x, __ldg(st + 4), fmaf(y * y, __ldg(st + 5), fmaf(z * z, __ldg(st + 6), __ldg(st + 7))));
  var += 2.0f * fmaf(x * y, __ldg(st + 8), fmaf(x * z, __ldg(st + 9), fmaf(y * z, __ldg(st + 10),
                fmaf(x, __ldg(st + 11), fmaf(y, __ldg(st + 12), z * __ldg(st + 13))))));
  sc = rsqrtf(fmaxf(var, 0.f) + eps);
  sh = -mean * sc;
}

// The same 64-column chunk of 16 rows from a TILE-BLOCKED matrix (the attention output of chain_qkvattn2_kernel<true>: the
// 16-byte chunk j of row r of tile t is float4 number (t * 64 + j) * 128 + r).  Lane = (row r = lane & 15, parity qq = lane >> 4),
// instruction j covers chunks 2j + qq: two 256-byte runs per request; buf[j] = chunk 2j + qq of row 16 w + r.
__device__ __forceinline__ void fetch_chunk16_blk(const float* __restrict__ x, int t, int w, int kc, int lane, float4* buf) {
  const float4* base = reinterpret_cast<const float4*>(x) + ((size_t)t * 64 + kc * 16 + (lane >> 4)) * 128 + 16 * w + (lane & 15);
#pragma unroll
  for (int j = 0; j < 8; ++j) buf[j] = __ldg(base + (size_t)(2 * j) * 128);
}
__device__ __forceinline__ void store_chunk16_blk(uint8_t* slot, int w, int lane, const float4* buf, bool split) {
  const int r = 16 * w + (lane & 15), qq = lane >> 4;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint2 hi, lo;
    split_f16x2(buf[j].x, buf[j].y, hi.x, lo.x);
    split_f16x2(buf[j].z, buf[j].w, hi.y, lo.y);
    const uint32_t off = swizzle128_offset(r, j) + (qq << 3);     // float4 chunk 2j + qq = 8-byte half qq of 16-byte fp16 chunk j
    *reinterpret_cast<uint2*>(slot + off) = hi;
    if (split) *reinterpret_cast<uint2*>(slot + CT_A_HALF + off) = lo;
  }
}
__device__ __forceinline__ void prefetch_tile_blk_l2(const float* __restrict__ x, int t, int w, int lane) {
  const char* base = reinterpret_cast<const char*>(x) + (size_t)t * 131072 + w * 16384 + lane * 128;     // 128 KB per tile
#pragma unroll
  for (int i = 0; i < 4; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + i * 4096));
}

// ===============================================================================================================
// out[M, 256*NT] = LN?(x)[M,256] . W[256*NT, 256]^T + bias (+ res)      (qkv and proj of ImplFuncAttention,
// model/shape/implicit.py:30,74 with norm1 of ImplFuncBlock :105 folded in: the LayerNorm statistics are computed by
// the loader warps, its affine is folded into W / bias at pack time).
// Same dataflow as chain_mlp_kernel without the second GEMM: coalesced LN loader -> ring L, weights through the TMA
// ring (blob = zs_gemm_tc_pack image of W: [n_tile][k_chunk][hi | lo]), accumulators ping-pong between the TMEM halves,
// and the epilogue transposes each 32x32 accumulator block inside its warp (4 KB of the idle ring E per warp) so that
// bias / residual / store run on whole 128-byte row segments.
constexpr int LN_THREADS = 576;   // chain_lin_kernel warps: 0-7 loaders (16 rows each), 8-15 epilogue, 16 MMA, 17 W loader
__global__ void __launch_bounds__(LN_THREADS, 1) chain_lin_kernel(ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  Bars B{smem_base + CT_OFF_BAR};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = p.precision == 0;
  const uint32_t tmem_base = chain_setup(B, smem_gen, smem_base, warp, 16, 256);
  const int n_tiles = (p.M + 127) / 128;
  const int NT = p.n_tiles;

  if (warp < 8) {
    Ring lr(CT_LSLOTS);
    float4 buf[8];
    int tn = 0;
    const bool tr0 = threadIdx.x == 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = t * 128 + warp * 16;
      float sc = 1.f, sh = 0.f;
      if (tr0) trace_ev(p.trace, 1, tn, 14);
      const bool blk = p.do_ln == 2;          // A is tile-blocked (and padded to whole tiles); no LayerNorm in this mode
      if (t + (int)gridDim.x < n_tiles) {
        if (blk) prefetch_tile_blk_l2(p.x, t + (int)gridDim.x, warp, lane);
        else prefetch_rows16_l2(p.x, p.ldx, m0 + (int)gridDim.x * 128, p.M, 0, lane);
        if (p.res)
          for (int nt = 0; nt < NT; ++nt) prefetch_rows16_l2(p.res, p.ldres, m0 + (int)gridDim.x * 128, p.M, nt * 256, lane);
      }
      if (p.do_ln == 1) warp_ln_stats16(p.x, p.ldx, m0, p.M, p.ln_eps, lane, sc, sh);
      if (blk) fetch_chunk16_blk(p.x, t, warp, 0, lane, buf);
      else fetch_chunk16(p.x, p.ldx, m0, p.M, 0, lane, buf);
      if (tr0) trace_ev(p.trace, 1, tn, 15);
      const int n_chunks = 4 * NT;
      for (int i = 0; i < n_chunks; ++i) {
        if (tr0) trace_ev(p.trace, 1, tn, 10);
        mbar_wait(B.lempty(lr.idx), lr.phase ^ 1);
        if (tr0) trace_ev(p.trace, 1, tn, 11);
        if (blk) store_chunk16_blk(smem_gen + CT_OFF_L + lr.idx * 2 * CT_A_HALF, warp, lane, buf, split);
        else store_chunk16(smem_gen + CT_OFF_L + lr.idx * 2 * CT_A_HALF, warp, lane, buf, p.do_ln == 1, sc, sh, split);
        fence_proxy_async_smem();
        mbar_arrive(B.lfull(lr.idx));
        if (tr0) trace_ev(p.trace, 1, tn, 12);
        lr.advance();
        if (i + 1 < n_chunks) {
          if (blk) fetch_chunk16_blk(p.x, t, warp, (i + 1) & 3, lane, buf);
          else fetch_chunk16(p.x, p.ldx, m0, p.M, ((i + 1) & 3) * 64, lane, buf);
        }
        if (tr0) trace_ev(p.trace, 1, tn, 13);
      }
    }
  } else if (warp == 17) {
    if (lane == 0) {
      Ring wr(CT_WSLOTS);
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) w_stream(B, smem_base, wr, p.blob, 4 * NT, split);
    }
  } else if (warp == 16) {
    if (lane == 0) {
      Ring wr(CT_WSLOTS), lr(CT_LSLOTS);
      uint32_t te_phase[2] = {0, 0};
      int acc = 0, tn = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int nt = 0; nt < NT; ++nt) {
          trace_ev(p.trace, 0, tn, 7);
          mbar_wait(B.tempty(acc), te_phase[acc] ^ 1); te_phase[acc] ^= 1;
          tc_fence_after();
          for (int kc = 0; kc < 4; ++kc) {
            trace_ev(p.trace, 0, tn, 1);
            mbar_wait(B.lfull(lr.idx), lr.phase);
            trace_ev(p.trace, 0, tn, 2);
            tc_fence_after();
            mma_chunk(B, smem_base, wr, smem_base + CT_OFF_L + lr.idx * 2 * CT_A_HALF, tmem_base + acc * 256, kc == 0, split,
                      B.lempty(lr.idx), p.trace, &tn);
            lr.advance();
          }
          umma_commit(B.tfull(acc));
          acc ^= 1;
        }
      }
    }
  } else {
    const int e = warp - 8, q = e & 3, hsel = e >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint8_t* wscr = smem_gen + CT_OFF_E + e * 4096;       // this warp's [32 rows][32 cols] fp32 transpose scratch
    const int sub = lane >> 3, q8 = lane & 7;
    uint32_t tf_phase[2] = {0, 0};
    int acc = 0, tn = 0;
    const bool tr0 = (warp == 8 && lane == 0);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int nt = 0; nt < NT; ++nt) {
        if (tr0) trace_ev(p.trace, 2, tn, 20);
        mbar_wait(B.tfull(acc), tf_phase[acc]); tf_phase[acc] ^= 1;
        if (tr0) trace_ev(p.trace, 2, tn, 21);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int col0 = c * 64 + hsel * 32;
          uint32_t rr[32];
          tmem_ld_32x32(tmem_base + acc * 256 + lane_off + col0, rr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(wscr + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
          __syncwarp();
          const int n0 = nt * 256 + col0 + 4 * q8;
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0));
          // residual loads batched 4 at a time: `res` may alias `out` (in-place update), which would otherwise serialise
          // load -> store -> load; 8 at a time spills at the 96-register budget of a 576-thread CTA
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            float4 rin[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int mm = t * 128 + q * 32 + 4 * (4 * hb + i) + sub;
              rin[i] = p.res ? *reinterpret_cast<const float4*>(p.res + (int64_t)(mm < p.M ? mm : 0) * p.ldres + n0)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rl = 4 * (4 * hb + i) + sub;
              const int mm = t * 128 + q * 32 + rl;
              const float4 a = *reinterpret_cast<const float4*>(wscr + rl * 128 + ((q8 ^ (rl & 7)) << 4));
              if (mm < p.M)
                *reinterpret_cast<float4*>(p.out + (int64_t)mm * p.ldo + n0) =
                    make_float4(a.x + bv.x + rin[i].x, a.y + bv.y + rin[i].y, a.z + bv.z + rin[i].z, a.w + bv.w + rin[i].w);
            }
          }
          __syncwarp();
          if (tr0) trace_ev(p.trace, 2, tn, 24);
        }
        tc_fence_before();
        mbar_arrive(B.tempty(acc));
        acc ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ===============================================================================================================
// logit = MLPBlocks([xyz, LN(x)])   (8 hidden layers, skips at 2,4,6, Softplus(100), last layer 256 -> 1)
// A chunks of ring L: 4 x LN(x) then 1 x [xyz, 0...]  (K order = [feat | xyz]; weights permuted to match at pack time)
// blob order: l0: 5 pairs (feat 4, xyz 1); l1: 4; l2: 5 (inputs) + 4 (h); l3: 4; l4: 5+4; l5: 4; l6: 5+4; l7: 4   = 48 pairs
__global__ void __launch_bounds__(CT_THREADS, 1) chain_occ_kernel(ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  Bars B{smem_base + CT_OFF_BAR};
  float* row_scratch = reinterpret_cast<float*>(smem_gen + CT_OFF_BAR + 256);   // [128] partial dots
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = p.precision == 0;
  const uint32_t tmem_base = chain_setup(B, smem_gen, smem_base, warp);
  const int n_tiles = (p.M + 127) / 128;

  if (warp < 4) {
    const int r = threadIdx.x;
    Ring lr(CT_LSLOTS);
    float4 buf[16];
    (void)r;
    const int sub = lane >> 4, q = lane & 15;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = t * 128 + warp * 32;
      float sc, sh;
      if (t + (int)gridDim.x < n_tiles) prefetch_rows_l2(p.x, p.ldx, m0 + (int)gridDim.x * 128, p.M, 0, lane);
      warp_ln_stats(p.x, p.ldx, m0, p.M, p.ln_eps, lane, sc, sh);
      for (int rep = 0; rep < 4; ++rep) {
        for (int kc = 0; kc < 5; ++kc) {
          if (kc < 4) {
            fetch_chunk_co(p.x, p.ldx, m0, p.M, kc * 64, lane, buf);
          } else {
            // the xyz chunk: columns 0..2 = the query point, the rest zero (lanes q == 0 own columns 0..3 of their rows)
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int m = m0 + 2 * j + sub;
              buf[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (q == 0 && m < p.M) {
                const float* pp = p.points + (int64_t)m * 3;
                buf[j] = make_float4(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2), 0.f);
              }
            }
          }
          mbar_wait(B.lempty(lr.idx), lr.phase ^ 1);
          store_chunk_co(smem_gen + CT_OFF_L + lr.idx * 2 * CT_A_HALF, warp, lane, buf, kc < 4, sc, sh,
                         (kc < 4 && p.ln_w) ? p.ln_w + kc * 64 : nullptr, (kc < 4 && p.ln_b) ? p.ln_b + kc * 64 : nullptr, split);
          fence_proxy_async_smem();
          mbar_arrive(B.lfull(lr.idx));
          lr.advance();
        }
      }
    }
  } else if (warp == 13) {
    if (lane == 0) {
      Ring wr(CT_WSLOTS);
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) w_stream(B, smem_base, wr, p.blob, 48, split);
    }
  } else if (warp == 12) {
    if (lane == 0) {
      Ring wr(CT_WSLOTS), lr(CT_LSLOTS), er(CT_ESLOTS);
      uint32_t te_phase[2] = {0, 0};
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int l = 0; l < 8; ++l) {
          const int half = l & 1;
          const uint32_t d = tmem_base + half * 256;
          mbar_wait(B.tempty(half), te_phase[half] ^ 1); te_phase[half] ^= 1;
          tc_fence_after();
          const int nL = (l & 1) ? 0 : 5, nE = l == 0 ? 0 : 4;
          for (int i = 0; i < nL; ++i) {
            mbar_wait(B.lfull(lr.idx), lr.phase);
            tc_fence_after();
            mma_chunk(B, smem_base, wr, smem_base + CT_OFF_L + lr.idx * 2 * CT_A_HALF, d, i == 0, split, B.lempty(lr.idx));
            lr.advance();
          }
          for (int i = 0; i < nE; ++i) {
            mbar_wait(B.efull(er.idx), er.phase);
            tc_fence_after();
            mma_chunk(B, smem_base, wr, smem_base + CT_OFF_E + er.idx * 2 * CT_A_HALF, d, nL == 0 && i == 0, split, B.eempty(er.idx));
            er.advance();
          }
          umma_commit(B.tfull(half));
        }
      }
    }
  } else {
    const int e = warp - 4, q = e & 3, hsel = e >> 2;
    const int row = q * 32 + lane;
    Ring er(CT_ESLOTS);
    uint32_t tf_phase[2] = {0, 0};
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m = t * 128 + row;
      for (int l = 0; l < 8; ++l) {
        const int half = l & 1;
        mbar_wait(B.tfull(half), tf_phase[half]); tf_phase[half] ^= 1;
        tc_fence_after();
        float dot = 0.f;
        for (int c = 0; c < 4; ++c) {
          uint32_t rr[32];
          tmem_ld_32x32(tmem_base + half * 256 + lane_off + c * 64 + hsel * 32, rr);
          tmem_ld_wait();
          const float* bl = p.bias + l * 256 + c * 64 + hsel * 32;
          if (l < 7) {
            mbar_wait(B.eempty(er.idx), er.phase ^ 1);
            epi_to_ring<ZS_ACT_SOFTPLUS100>(smem_gen + CT_OFF_E + er.idx * 2 * CT_A_HALF, row, hsel, rr, bl, split);
            fence_proxy_async_smem();
            mbar_arrive(B.efull(er.idx));
            er.advance();
          } else {
            const float* w8 = p.bias2 + c * 64 + hsel * 32;
            float2 dot2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bl) + j), w4 = __ldg(reinterpret_cast<const float4*>(w8) + j);
              const float2 s0 = fast_softplus100_2(add2(make_float2(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1])), make_float2(b4.x, b4.y)));
              const float2 s1 = fast_softplus100_2(add2(make_float2(__uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3])), make_float2(b4.z, b4.w)));
              dot2 = fma2(s0, make_float2(w4.x, w4.y), dot2);
              dot2 = fma2(s1, make_float2(w4.z, w4.w), dot2);
            }
            dot += dot2.x + dot2.y;
          }
        }
        tc_fence_before();
        mbar_arrive(B.tempty(half));
        if (l == 7) {
          // combine the two column halves of each row: hsel 1 publishes, hsel 0 finishes (named barrier 1, 256 threads)
          if (hsel == 1) row_scratch[row] = dot;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (hsel == 0 && m < p.M) {
            float v = dot + row_scratch[row] + p.b8;
            p.out[m] = p.apply_sigmoid ? 1.0f / (1.0f + __expf(-v)) : v;
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ===============================================================================================================
// Point-to-latent attention of one image (model/shape/implicit.py:38-57), flash-style.  Per 128-point tile the 8 heads
// are processed as 4 PAIRS (2p, 2p+1); the two heads of a pair are owned by two independent groups of 4 epilogue
// warps (one thread = one query row of one head, so the softmax needs no cross-thread exchange):
//   loader   : q[:, 64p:64p+64] -> ring L as one split-bf16 A chunk (head 2p = K-steps 0,1; head 2p+1 = K-steps 2,3),
//              the points' own scores q_h.k_h -> smem;
//   MMA      : S_g = Q_h K_lat_h^T (N=208, own TMEM buffer per group) ; O_g = P_h V_lat_h (N=32) chunk by chunk;
//   group g  : row max (pass 1 over TMEM) -> e_j = exp(scale (s_j - max)) -> its ring-E slot (split bf16) -> row sum;
//              (O_g + e_self v_p,h) / sum -> transposed in smem -> coalesced store.
// The probabilities never leave the SM.  Weight ring (4 x 32 KB) per pair: K pair tile hi, lo ([208 keys x 64] : the
// two heads' 32 dims side by side), V_2p, V_2p+1 (each 4 key-chunks x [hi 4K | lo 4K], rows = head dims).
// ChainParams reuse: x = qkv [M,768] (q|k|v, ldx), blob = K pair tiles (4 x [hi 32K | lo 32K]), points -> V blob
// (8 heads x 32 KB), out = O [M,256], b8 = softmax scale, apply_sigmoid = number of latent keys (<= 208).
constexpr int AT_WSLOTS = 4;
constexpr int AT_OFF_L = AT_WSLOTS * CT_TILE_BYTES;          // 128 KB: the pair's q chunk (hi 16K | lo 16K)
constexpr int AT_OFF_E = AT_OFF_L + 2 * CT_A_HALF;           // 160 KB: one 32 KB slot per group
constexpr int AT_OFF_BAR = AT_OFF_E + 2 * 2 * CT_A_HALF;     // 224 KB
constexpr int AT_OFF_SELF = AT_OFF_BAR + 256;                // [2 pair parities][2 heads][128 rows] fp32 = 2 KB
static_assert(AT_OFF_BAR == CT_OFF_BAR && AT_OFF_SELF + 2048 <= CT_SMEM_USED, "attention smem layout");

struct ABars {
  uint32_t base;
  __device__ uint32_t wfull(int i) const { return base + 8u * i; }
  __device__ uint32_t wempty(int i) const { return base + 32u + 8u * i; }
  __device__ uint32_t lfull() const { return base + 64u; }
  __device__ uint32_t lempty() const { return base + 72u; }
  // ring-E slot of group g is handed over in two 32-key halves (K-steps {0,1} and {2,3}): h = 0, 1
  __device__ uint32_t efull(int g, int h) const { return base + (h ? 184u : 80u) + 8u * g; }
  __device__ uint32_t eempty(int g, int h) const { return base + (h ? 200u : 96u) + 8u * g; }
  __device__ uint32_t sfull(int g) const { return base + 112u + 8u * g; }
  __device__ uint32_t sempty(int g) const { return base + 128u + 8u * g; }
  __device__ uint32_t ofull(int g) const { return base + 144u + 8u * g; }
  __device__ uint32_t oempty(int g) const { return base + 160u + 8u * g; }
  __device__ uint32_t tmem_slot() const { return base + 176u; }
};

// e = exp2(s * sl2 - mxs) of 2*NP scores (keys k0 ...), masked past n_keys; accumulates the row sum; writes the split-bf16
// values as NP/4 16-byte chunks of the row (16-byte chunk index c16_0 ...) of a swizzled A chunk
template <int NP>
__device__ __forceinline__ void attn_exp_block(const uint32_t* rr, int k0, int n_keys, float sl2, float mxs, float2& sum2,
                                               uint8_t* slot, int row, int c16_0, bool split) {
  const bool full = k0 + 2 * NP <= n_keys;
#pragma unroll
  for (int cc = 0; cc < NP / 4; ++cc) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = cc * 8 + 2 * j;
      const float2 a = fma2(make_float2(__uint_as_float(rr[i]), __uint_as_float(rr[i + 1])), bc2(sl2), bc2(-mxs));
      float2 v = make_float2(fast_ex2(a.x), fast_ex2(a.y));
      if (!full) { if (k0 + i >= n_keys) v.x = 0.f; if (k0 + i + 1 >= n_keys) v.y = 0.f; }
      sum2 = add2(sum2, v);
      split_f16x2(v.x, v.y, hi[j], lo[j]);
    }
    const uint32_t off = swizzle128_offset(row, c16_0 + cc);
    *reinterpret_cast<uint4*>(slot + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (split) *reinterpret_cast<uint4*>(slot + CT_A_HALF + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

template <int N>
__device__ __forceinline__ float attn_max_block(const uint32_t* rr, int k0, int n_keys, float mx) {
  if (k0 + N <= n_keys) {
#pragma unroll
    for (int j = 0; j < N; ++j) mx = fmaxf(mx, __uint_as_float(rr[j]));
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) if (k0 + j < n_keys) mx = fmaxf(mx, __uint_as_float(rr[j]));
  }
  return mx;
}

__global__ void __launch_bounds__(CT_THREADS, 1) chain_attn_kernel(ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  if (smem_base - smem_u32(smem_raw) > CT_SMEM - CT_SMEM_USED) __trap();   // dynamic smem base less aligned than budgeted
  ABars B{smem_base + AT_OFF_BAR};
  float* sself = reinterpret_cast<float*>(smem_gen + AT_OFF_SELF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = p.precision == 0;
  const int n_tiles = (p.M + 127) / 128;
  const int n_keys = p.apply_sigmoid;
  const uint8_t* vblob = reinterpret_cast<const uint8_t*>(p.points);
  constexpr uint32_t K_TILE_BYTES = 208 * 128;     // rows actually read by the N=208 MMA

  if (threadIdx.x == 0) {
    for (int i = 0; i < AT_WSLOTS; ++i) { mbar_init(B.wfull(i), 1); mbar_init(B.wempty(i), 1); }
    mbar_init(B.lfull(), 128); mbar_init(B.lempty(), 1);
    for (int g = 0; g < 2; ++g) {
      for (int h = 0; h < 2; ++h) { mbar_init(B.efull(g, h), 128); mbar_init(B.eempty(g, h), 1); }
      mbar_init(B.sfull(g), 1); mbar_init(B.sempty(g), 128);
      mbar_init(B.ofull(g), 1); mbar_init(B.oempty(g), 128);
    }
    fence_mbar_init();
  }
  if (warp == 12) tmem_alloc(B.tmem_slot(), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (B.tmem_slot() - smem_base));
  // TMEM columns: S of group 0 [0,208), O of group 0 [208,240), S of group 1 [256,464), O of group 1 [464,496)

  if (warp < 4) {
    // ---------------- loader: per pair the q chunk (coalesced) and the points' own scores ----------------
    const int sub = lane >> 4, q = lane & 15;
    uint32_t lph = 0;
    float4 buf[16];
    int tn = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = t * 128 + warp * 32;
      for (int pr = 0; pr < 4; ++pr) {
        {
          // pull the NEXT pair's q / k / v column blocks (2 x 128-byte lines each) of this warp's 32 rows into L2, one
          // pair (~10k cycles) ahead of their use by this loader (q, k) and by the epilogue (the points' own values v)
          const int tn2 = pr == 3 ? t + (int)gridDim.x : t, pn = (pr + 1) & 3;
          const int mrow = tn2 * 128 + warp * 32 + lane;
          if (tn2 < n_tiles && mrow < p.M) {
            const float* rowp = p.x + (int64_t)mrow * p.ldx + pn * 64;
#pragma unroll
            for (int l = 0; l < 6; ++l)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + (l >> 1) * 256 + (l & 1) * 32));
          }
        }
        fetch_chunk_co(p.x, p.ldx, m0, p.M, pr * 64, lane, buf);
        float dd[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int m = m0 + 2 * j + sub;
          float4 k4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m < p.M) k4 = __ldg(reinterpret_cast<const float4*>(p.x + (int64_t)m * p.ldx + 256 + pr * 64) + q);
          float d = fmaf(buf[j].x, k4.x, fmaf(buf[j].y, k4.y, fmaf(buf[j].z, k4.z, buf[j].w * k4.w)));
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          dd[j] = d;
        }
        if (threadIdx.x == 0) trace_ev(p.trace, 1, tn, 10);
        mbar_wait(B.lempty(), lph ^ 1);
        if (threadIdx.x == 0) trace_ev(p.trace, 1, tn, 11);
        // (the previous user of this sself parity, pair pr-2, is finished: S(pr-1) was issued after both groups released it)
        if ((q & 7) == 0) {
          float* ss = sself + (pr & 1) * 256 + (q >> 3) * 128 + warp * 32 + sub;
#pragma unroll
          for (int j = 0; j < 16; ++j) ss[2 * j] = dd[j];
        }
        store_chunk_co(smem_gen + AT_OFF_L, warp, lane, buf, false, 0.f, 0.f, nullptr, nullptr, split);
        fence_proxy_async_smem();
        mbar_arrive(B.lfull());
        if (threadIdx.x == 0) trace_ev(p.trace, 1, tn, 12);
        lph ^= 1;
      }
    }
  } else if (warp == 13) {
    // ---------------- W loader: per pair K hi, K lo, V_2p, V_2p+1 ----------------
    if (lane == 0) {
      Ring wr(AT_WSLOTS);
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int pr = 0; pr < 4; ++pr) {
          for (int j = 0; j < 4; ++j) {
            if (j == 1 && !split) continue;
            const uint8_t* src = j < 2 ? p.blob + (size_t)pr * 2 * CT_TILE_BYTES + (size_t)j * CT_TILE_BYTES
                                       : vblob + (size_t)(2 * pr + (j - 2)) * CT_TILE_BYTES;
            const uint32_t bytes = j < 2 ? K_TILE_BYTES : (uint32_t)CT_TILE_BYTES;
            mbar_wait(B.wempty(wr.idx), wr.phase ^ 1);
            mbar_arrive_expect_tx(B.wfull(wr.idx), bytes);
            bulk_g2s(smem_base + CT_OFF_W + wr.idx * CT_TILE_BYTES, src, bytes, B.wfull(wr.idx));
            wr.advance();
          }
        }
      }
    }
  } else if (warp == 12) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      Ring wr(AT_WSLOTS);
      uint32_t lph = 0, se_ph[2] = {0, 0}, oe_ph[2] = {0, 0}, ef_ph[2][2] = {{0, 0}, {0, 0}};
      const uint32_t idesc_s = umma_idesc_f16(128, 208), idesc_o = umma_idesc_f16(128, 32);
      const uint32_t d_s[2] = {tmem_base, tmem_base + 256}, d_o[2] = {tmem_base + 208, tmem_base + 464};
      const uint32_t a_addr = smem_base + AT_OFF_L;
      const uint64_t a_hi = umma_desc_sw128(a_addr), a_lo = umma_desc_sw128(a_addr + CT_A_HALF);
      int tn = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int pr = 0; pr < 4; ++pr) {
          // ---- S_g = Q_h K_h^T : head g of the pair = K-steps 2g, 2g+1 of the q chunk and of the K pair tile ----
          trace_ev(p.trace, 0, tn, 1);
          mbar_wait(B.lfull(), lph); lph ^= 1;
          tc_fence_after();
          mbar_wait(B.wfull(wr.idx), wr.phase);
          tc_fence_after();
          {
            const uint64_t w = umma_desc_sw128(smem_base + CT_OFF_W + wr.idx * CT_TILE_BYTES);
            for (int g = 0; g < 2; ++g) {
              mbar_wait(B.sempty(g), se_ph[g] ^ 1); se_ph[g] ^= 1;
              tc_fence_after();
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const int kk = 2 * g + k;
                umma_bf16(d_s[g], a_hi + 2 * kk, w + 2 * kk, idesc_s, k > 0 ? 1u : 0u);
                if (split) umma_bf16(d_s[g], a_lo + 2 * kk, w + 2 * kk, idesc_s, 1u);
              }
              if (!split) umma_commit(B.sfull(g));
            }
            umma_commit(B.wempty(wr.idx));
            wr.advance();
          }
          if (split) {
            mbar_wait(B.wfull(wr.idx), wr.phase);
            tc_fence_after();
            const uint64_t w = umma_desc_sw128(smem_base + CT_OFF_W + wr.idx * CT_TILE_BYTES);
            for (int g = 0; g < 2; ++g) {
#pragma unroll
              for (int k = 0; k < 2; ++k) umma_bf16(d_s[g], a_hi + 2 * (2 * g + k), w + 2 * (2 * g + k), idesc_s, 1u);
              umma_commit(B.sfull(g));
            }
            umma_commit(B.wempty(wr.idx));
            wr.advance();
          }
          umma_commit(B.lempty());
          trace_ev(p.trace, 0, tn, 2);
          // ---- O_g = P_h V_h : key chunks from the groups' ring-E slots, the two heads interleaved ----
          uint32_t v_addr[2]; int v_slot[2];
          for (int g = 0; g < 2; ++g) {
            mbar_wait(B.wfull(wr.idx), wr.phase);
            v_slot[g] = wr.idx; v_addr[g] = smem_base + CT_OFF_W + wr.idx * CT_TILE_BYTES;
            wr.advance();
            mbar_wait(B.oempty(g), oe_ph[g] ^ 1); oe_ph[g] ^= 1;
          }
          tc_fence_after();
          for (int c = 0; c < 4; ++c) {
            const int nh = c == 3 ? 1 : 2;        // keys 192..207 are one K-step (half 0 only); 208.. do not exist
            for (int hh = 0; hh < nh; ++hh) {
              for (int g = 0; g < 2; ++g) {
                mbar_wait(B.efull(g, hh), ef_ph[g][hh]); ef_ph[g][hh] ^= 1;
                tc_fence_after();
                const uint32_t e_addr = smem_base + AT_OFF_E + g * 2 * CT_A_HALF;
                const uint64_t e_hi = umma_desc_sw128(e_addr), e_lo = umma_desc_sw128(e_addr + CT_A_HALF);
                const uint64_t v_hi = umma_desc_sw128(v_addr[g] + c * 8192), v_lo = umma_desc_sw128(v_addr[g] + c * 8192 + 4096);
                const int k0 = 2 * hh, k1 = c == 3 ? 1 : 2 * hh + 2;
                for (int k = k0; k < k1; ++k) {
                  umma_bf16(d_o[g], e_hi + 2 * k, v_hi + 2 * k, idesc_o, (c > 0 || k > 0) ? 1u : 0u);
                  if (split) {
                    umma_bf16(d_o[g], e_lo + 2 * k, v_hi + 2 * k, idesc_o, 1u);
                    umma_bf16(d_o[g], e_hi + 2 * k, v_lo + 2 * k, idesc_o, 1u);
                  }
                }
                umma_commit(B.eempty(g, hh));
                if (c == 3) umma_commit(B.ofull(g));
                trace_ev(p.trace, 0, tn, 3);
              }
            }
          }
          umma_commit(B.wempty(v_slot[0]));
          umma_commit(B.wempty(v_slot[1]));
        }
      }
    }
  } else {
    // ---------------- epilogue: group g = one head of the pair; one thread = one query row ----------------
    const int e = warp - 4, wq = e & 3, g = e >> 2;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t s_tm = tmem_base + lane_off + (g ? 256u : 0u), o_tm = tmem_base + lane_off + (g ? 464u : 208u);
    uint8_t* eslot = smem_gen + AT_OFF_E + g * 2 * CT_A_HALF;
    uint8_t* wscr = eslot + wq * 4096;            // this warp's own 32 rows of the slot's hi half: [32][32] fp32 transpose scratch
    uint32_t sf_ph = 0, of_ph = 0, ee_ph[2] = {0, 0};
    const float sl2 = p.b8 * 1.4426950408889634f;
    const bool tr0 = (warp == 4 && lane == 0);
    int tn = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int pr = 0; pr < 4; ++pr) {
        const int h = 2 * pr + g;
        if (tr0) trace_ev(p.trace, 2, tn, 20);
        mbar_wait(B.sfull(g), sf_ph); sf_ph ^= 1;
        if (tr0) trace_ev(p.trace, 2, tn, 21);
        tc_fence_after();
        const float s_self = sself[(pr & 1) * 256 + g * 128 + row];
        // pass 1: row max over the latent keys and the point's own key
        float mx = s_self;
#pragma unroll 1
        for (int b = 0; b < 6; ++b) {
          uint32_t rr[32];
          tmem_ld_32x32(s_tm + b * 32, rr);
          tmem_ld_wait();
          mx = attn_max_block<32>(rr, b * 32, n_keys, mx);
        }
        {
          uint32_t rr[16];
          tmem_ld_32x16(s_tm + 192, rr);
          tmem_ld_wait();
          mx = attn_max_block<16>(rr, 192, n_keys, mx);
        }
        const float mxs = mx * sl2;
        if (tr0) trace_ev(p.trace, 2, tn, 22);
        // pass 2: e_j -> this group's ring-E slot, chunk by chunk (64 keys; the last chunk has 16)
        float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          if (c < 3) {
            uint32_t rr[32];
            tmem_ld_32x32(s_tm + c * 64, rr);
            tmem_ld_wait();
            mbar_wait(B.eempty(g, 0), ee_ph[0] ^ 1); ee_ph[0] ^= 1;
            attn_exp_block<16>(rr, c * 64, n_keys, sl2, mxs, sum2, eslot, row, 0, split);
            fence_proxy_async_smem();
            mbar_arrive(B.efull(g, 0));
            tmem_ld_32x32(s_tm + c * 64 + 32, rr);
            tmem_ld_wait();
            mbar_wait(B.eempty(g, 1), ee_ph[1] ^ 1); ee_ph[1] ^= 1;
            attn_exp_block<16>(rr, c * 64 + 32, n_keys, sl2, mxs, sum2, eslot, row, 4, split);
            fence_proxy_async_smem();
            mbar_arrive(B.efull(g, 1));
          } else {
            uint32_t rr[16];
            tmem_ld_32x16(s_tm + 192, rr);
            tmem_ld_wait();
            mbar_wait(B.eempty(g, 0), ee_ph[0] ^ 1); ee_ph[0] ^= 1;
            attn_exp_block<8>(rr, 192, n_keys, sl2, mxs, sum2, eslot, row, 0, split);
            fence_proxy_async_smem();
            mbar_arrive(B.efull(g, 0));
          }
          if (tr0) trace_ev(p.trace, 2, tn, 23);
        }
        tc_fence_before();
        mbar_arrive(B.sempty(g));
        const float e_self = fast_ex2(fmaf(s_self, sl2, -mxs));
        const float inv = 1.0f / ((sum2.x + sum2.y) + e_self);
        const float ps = e_self * inv;
        // the points' own values v_p,h of the 8 (row, 4-column) pieces this lane stores below: issued together (L2 hits, the
        // loader prefetched them a pair ago) so their latency overlaps the wait for O instead of serialising 8 round trips
        const int q8 = lane & 7;
        float4 vself[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int mm = t * 128 + wq * 32 + 4 * i + (lane >> 3);
          vself[i] = __ldg(reinterpret_cast<const float4*>(p.x + (int64_t)(mm < p.M ? mm : 0) * p.ldx + 512 + h * 32) + q8);
        }
        // O_h: 32 accumulator columns of this row -> normalise -> transpose inside the warp -> coalesced store
        if (tr0) trace_ev(p.trace, 2, tn, 24);
        mbar_wait(B.ofull(g), of_ph); of_ph ^= 1;
        if (tr0) trace_ev(p.trace, 2, tn, 25);
        tc_fence_after();
        {
          uint32_t rr[32];
          tmem_ld_32x32(o_tm, rr);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(B.oempty(g));
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(wscr + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_float4(__uint_as_float(rr[4 * j]) * inv, __uint_as_float(rr[4 * j + 1]) * inv,
                            __uint_as_float(rr[4 * j + 2]) * inv, __uint_as_float(rr[4 * j + 3]) * inv);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = 4 * i + (lane >> 3);
          const float psr = __shfl_sync(0xffffffffu, ps, rl);
          const int mm = t * 128 + wq * 32 + rl;
          const float4 a = *reinterpret_cast<const float4*>(wscr + rl * 128 + ((q8 ^ (rl & 7)) << 4));
          const float4 vv = vself[i];
          if (mm < p.M)
            *(reinterpret_cast<float4*>(p.out + (int64_t)mm * 256 + h * 32) + q8) =
                make_float4(fmaf(psr, vv.x, a.x), fmaf(psr, vv.y, a.y), fmaf(psr, vv.z, a.z), fmaf(psr, vv.w, a.w));
        }
        __syncwarp();
        if (tr0) trace_ev(p.trace, 2, tn, 26);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}


// ===============================================================================================================
// LayerNorm + qkv + point->latent attention of one image in ONE kernel: norm1 of ImplFuncBlock and ImplFuncAttention
// up to (not including) the output projection (model/shape/implicit.py:105, 30-57).  q, k, v of the query points, the
// scores and the probabilities never leave the SM; HBM sees x (read) and the attention output O (written).
//
// Per 128-point tile the 8 heads run as 4 pairs.  Work item of the pipeline = one pair ("unit"):
//   loaders (8 warps) : LN(x) -> split fp16 -> ring L, 4 K-chunks per unit (the qkv GEMM of a pair has N = 192:
//                       [q of heads 2p, 2p+1 | k | v], weight tile rows in that order, norm1 affine folded at pack time);
//   MMA               : acc[128 x 192] = LN(x) Wp^T for the NEXT unit while the current unit's attention runs;
//                       per head S = Q_h K_h^T (N = 208) -> one S buffer, O_h = P_h V_h (N = 32) -> two O buffers;
//   softmax warps (8) : thread = (row, half).  epi-1: half g drains q_g | k_g | v_g of head 2p+g from the accumulator:
//                       q_g + b -> fp16 split -> its K-steps of the Q chunk, the point's own score q.k and value v stay
//                       in registers.  Per head both halves share the keys (half 0: K-steps {0,1} of every 64-key
//                       chunk, half 1: K-steps {2,3}): row max and row sum are exchanged through smem (named barriers),
//                       probabilities go to the P slot in 32-key halves (pfull/pempty per half), and the half that owns
//                       the head finishes (O + e_self v_self) / sum -> transpose through the idle P slot -> global.
// TMEM (this kernel): qkv accumulator [0,192), S [192,400), O of even heads [400,432), of odd heads [432,464).
// flags: 1 = k, v columns single-pass (Ah Wh only), 2 = scores without Qh Kl, 4 = P V without Ph Vl
//        (profiles/r2_precision_study.md; precision 1 = everything single-pass).
constexpr int QA_THREADS = 576;                                 // warps 0-7 loaders, 8-15 softmax, 16 MMA, 17 W loader
constexpr int QA_WSLOTS = 3;
constexpr int QA_OFF_W = 0;                                     // 3 x 32 KB weight / key / value tiles
constexpr int QA_OFF_L = QA_OFF_W + QA_WSLOTS * CT_TILE_BYTES;  //  96 KB: 2 x (hi 16K | lo 16K) LN(x) chunks
constexpr int QA_OFF_Q = QA_OFF_L + 2 * 2 * CT_A_HALF;          // 160 KB: the unit's q chunk (hi | lo)
constexpr int QA_OFF_P = QA_OFF_Q + 2 * CT_A_HALF;              // 192 KB: probability slot (hi | lo)
constexpr int QA_OFF_BAR = QA_OFF_P + 2 * CT_A_HALF;            // 224 KB
constexpr int QA_OFF_X = QA_OFF_BAR + 256;                      // [2 (max, sum)][2 halves][128 rows] fp32 = 2 KB
static_assert(QA_OFF_BAR == CT_OFF_BAR && QA_OFF_X + 2048 <= CT_SMEM_USED, "qkv-attention smem layout");
constexpr uint32_t QA_QKV_TILE_BYTES = 192 * 128;               // rows of a packed 256-row tile actually read (N = 192)
constexpr uint32_t QA_Q_TILE_BYTES = 64 * 128;                  // ... when only the q rows take the pass (N = 64)
constexpr uint32_t QA_K_TILE_BYTES = 208 * 128;

struct QBars {
  uint32_t base;
  __device__ uint32_t wfull(int i) const { return base + 8u * i; }
  __device__ uint32_t wempty(int i) const { return base + 24u + 8u * i; }
  __device__ uint32_t lfull(int i) const { return base + 48u + 8u * i; }
  __device__ uint32_t lempty(int i) const { return base + 64u + 8u * i; }
  __device__ uint32_t accfull() const { return base + 80u; }
  __device__ uint32_t accempty() const { return base + 88u; }
  __device__ uint32_t qfull() const { return base + 96u; }
  __device__ uint32_t qempty() const { return base + 104u; }
  __device__ uint32_t sfull() const { return base + 112u; }
  __device__ uint32_t sempty() const { return base + 120u; }
  __device__ uint32_t pfull(int h) const { return base + 128u + 8u * h; }
  __device__ uint32_t pempty(int h) const { return base + 144u + 8u * h; }
  __device__ uint32_t ofull(int i) const { return base + 160u + 8u * i; }
  __device__ uint32_t oempty(int i) const { return base + 176u + 8u * i; }
  __device__ uint32_t tmem_slot() const { return base + 192u; }
};

__global__ void __launch_bounds__(QA_THREADS, 1) chain_qkvattn_kernel(ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  if (smem_base - smem_u32(smem_raw) > CT_SMEM - CT_SMEM_USED) __trap();   // dynamic smem base less aligned than budgeted
  QBars B{smem_base + QA_OFF_BAR};
  float* xbuf = reinterpret_cast<float*>(smem_gen + QA_OFF_X);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = p.precision == 0;
  const bool kv1 = split && (p.flags & 1), s2 = split && (p.flags & 2), pv2 = split && (p.flags & 4);
  const int n_tiles = (p.M + 127) / 128;
  const int n_keys = p.n_keys;

  if (threadIdx.x == 0) {
    for (int i = 0; i < QA_WSLOTS; ++i) { mbar_init(B.wfull(i), 1); mbar_init(B.wempty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(B.lfull(i), 256); mbar_init(B.lempty(i), 1); }
    mbar_init(B.accfull(), 1); mbar_init(B.accempty(), 256);
    mbar_init(B.qfull(), 256); mbar_init(B.qempty(), 1);
    mbar_init(B.sfull(), 1); mbar_init(B.sempty(), 256);
    for (int i = 0; i < 2; ++i) {
      mbar_init(B.pfull(i), 128); mbar_init(B.pempty(i), 1);
      mbar_init(B.ofull(i), 1); mbar_init(B.oempty(i), 128);
    }
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(B.tmem_slot(), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (B.tmem_slot() - smem_base));

  if (warp < 8) {
    // ---------------- loaders: LN(x) chunks, 4 units x 4 K-chunks per tile (as chain_lin_kernel) ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");      // the 8 softmax warps take the registers (v_self stays live)
    Ring lr(2);
    float4 buf[8];
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = t * 128 + warp * 16;
      float sc = 1.f, sh = 0.f;
      if (t + (int)gridDim.x < n_tiles) prefetch_rows16_l2(p.x, p.ldx, m0 + (int)gridDim.x * 128, p.M, 0, lane);
      warp_ln_stats16(p.x, p.ldx, m0, p.M, p.ln_eps, lane, sc, sh);
      fetch_chunk16(p.x, p.ldx, m0, p.M, 0, lane, buf);
      for (int i = 0; i < 16; ++i) {
        mbar_wait(B.lempty(lr.idx), lr.phase ^ 1);
        store_chunk16(smem_gen + QA_OFF_L + lr.idx * 2 * CT_A_HALF, warp, lane, buf, true, sc, sh, split);
        fence_proxy_async_smem();
        mbar_arrive(B.lfull(lr.idx));
        lr.advance();
        if (i + 1 < 16) fetch_chunk16(p.x, p.ldx, m0, p.M, ((i + 1) & 3) * 64, lane, buf);
      }
    }
  } else if (warp == 17) {
    // ---------------- W loader: tiles in exactly the order the MMA thread consumes them ----------------
    if (lane == 0) {
      Ring wr(QA_WSLOTS);
      auto put = [&](const uint8_t* src, uint32_t bytes) {
        mbar_wait(B.wempty(wr.idx), wr.phase ^ 1);
        mbar_arrive_expect_tx(B.wfull(wr.idx), bytes);
        bulk_g2s(smem_base + QA_OFF_W + wr.idx * CT_TILE_BYTES, src, bytes, B.wfull(wr.idx));
        wr.advance();
      };
      auto qkv_chunk = [&](int pr, int kc) {
        const uint8_t* src = p.blob + ((size_t)pr * 4 + kc) * 2 * CT_TILE_BYTES;
        put(src, QA_QKV_TILE_BYTES);
        if (split) put(src + CT_TILE_BYTES, kv1 ? QA_Q_TILE_BYTES : QA_QKV_TILE_BYTES);
      };
      bool first = true;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int pr = 0; pr < 4; ++pr) {
          if (first) { for (int kc = 0; kc < 4; ++kc) qkv_chunk(0, kc); first = false; }
          const bool has_next = pr < 3 || t + (int)gridDim.x < n_tiles;
          const int pn = (pr + 1) & 3;
          for (int g = 0; g < 2; ++g) {
            put(p.kblob + (size_t)pr * 2 * CT_TILE_BYTES, QA_K_TILE_BYTES);
            if (split && !s2) put(p.kblob + (size_t)pr * 2 * CT_TILE_BYTES + CT_TILE_BYTES, QA_K_TILE_BYTES);
            if (has_next) { qkv_chunk(pn, 2 * g); qkv_chunk(pn, 2 * g + 1); }
            put(p.vblob + (size_t)(2 * pr + g) * CT_TILE_BYTES, CT_TILE_BYTES);
          }
        }
      }
    }
  } else if (warp == 16) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      Ring wr(QA_WSLOTS), lr(2);
      uint32_t ph_accempty = 0, ph_qfull = 0, ph_sempty = 0, ph_pfull = 0, ph_oempty = 0;   // the last two: bit i = phase of barrier i
      const uint32_t idesc_qkv = umma_idesc_f16(128, 192), idesc_q2 = kv1 ? umma_idesc_f16(128, 64) : idesc_qkv;
      const uint32_t idesc_s = umma_idesc_f16(128, 208), idesc_o = umma_idesc_f16(128, 32);
      const uint32_t d_acc = tmem_base, d_s = tmem_base + 192;
      const uint64_t q_hi = umma_desc_sw128(smem_base + QA_OFF_Q), q_lo = umma_desc_sw128(smem_base + QA_OFF_Q + CT_A_HALF);
      const uint64_t e_hi = umma_desc_sw128(smem_base + QA_OFF_P), e_lo = umma_desc_sw128(smem_base + QA_OFF_P + CT_A_HALF);
      auto qkv_chunk = [&](int kc) {
        mbar_wait(B.lfull(lr.idx), lr.phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + QA_OFF_L + lr.idx * 2 * CT_A_HALF;
        const uint64_t a_hi = umma_desc_sw128(a_addr), a_lo = umma_desc_sw128(a_addr + CT_A_HALF);
        mbar_wait(B.wfull(wr.idx), wr.phase);
        tc_fence_after();
        {
          const uint64_t w = umma_desc_sw128(smem_base + QA_OFF_W + wr.idx * CT_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(d_acc, a_hi + 2 * k, w + 2 * k, idesc_qkv, (kc > 0 || k > 0) ? 1u : 0u);
            if (split) umma_bf16(d_acc, a_lo + 2 * k, w + 2 * k, idesc_q2, 1u);
          }
          umma_commit(B.wempty(wr.idx));
          wr.advance();
        }
        if (split) {
          mbar_wait(B.wfull(wr.idx), wr.phase);
          tc_fence_after();
          const uint64_t w = umma_desc_sw128(smem_base + QA_OFF_W + wr.idx * CT_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_acc, a_hi + 2 * k, w + 2 * k, idesc_q2, 1u);
          umma_commit(B.wempty(wr.idx));
          wr.advance();
        }
        umma_commit(B.lempty(lr.idx));
        lr.advance();
      };
      // the first unit's qkv GEMM; afterwards unit u+1's runs inside unit u
      mbar_wait(B.accempty(), ph_accempty ^ 1); ph_accempty ^= 1;
      tc_fence_after();
      for (int kc = 0; kc < 4; ++kc) qkv_chunk(kc);
      umma_commit(B.accfull());
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int pr = 0; pr < 4; ++pr) {
          const bool has_next = pr < 3 || t + (int)gridDim.x < n_tiles;
          mbar_wait(B.qfull(), ph_qfull); ph_qfull ^= 1;
          tc_fence_after();
          for (int g = 0; g < 2; ++g) {
            // ---- S = Q_h K_h^T: head g of the pair = K-steps 2g, 2g+1 of the q chunk and of the K pair tile ----
            mbar_wait(B.sempty(), ph_sempty ^ 1); ph_sempty ^= 1;
            tc_fence_after();
            mbar_wait(B.wfull(wr.idx), wr.phase);
            tc_fence_after();
            {
              const uint64_t w = umma_desc_sw128(smem_base + QA_OFF_W + wr.idx * CT_TILE_BYTES);
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const int kk = 2 * g + k;
                umma_bf16(d_s, q_hi + 2 * kk, w + 2 * kk, idesc_s, k > 0 ? 1u : 0u);
                if (split) umma_bf16(d_s, q_lo + 2 * kk, w + 2 * kk, idesc_s, 1u);
              }
              umma_commit(B.wempty(wr.idx));
              wr.advance();
            }
            if (split && !s2) {
              mbar_wait(B.wfull(wr.idx), wr.phase);
              tc_fence_after();
              const uint64_t w = umma_desc_sw128(smem_base + QA_OFF_W + wr.idx * CT_TILE_BYTES);
#pragma unroll
              for (int k = 0; k < 2; ++k) umma_bf16(d_s, q_hi + 2 * (2 * g + k), w + 2 * (2 * g + k), idesc_s, 1u);
              umma_commit(B.wempty(wr.idx));
              wr.advance();
            }
            umma_commit(B.sfull());
            if (g == 1) umma_commit(B.qempty());
            // ---- the next unit's qkv GEMM, two K-chunks per head: fills the tensor pipe while the softmax runs ----
            if (has_next) {
              if (g == 0) { mbar_wait(B.accempty(), ph_accempty ^ 1); ph_accempty ^= 1; tc_fence_after(); }
              qkv_chunk(2 * g);
              qkv_chunk(2 * g + 1);
              if (g == 1) umma_commit(B.accfull());
            }
            // ---- O_h = P_h V_h: 32-key halves from the P slot ----
            mbar_wait(B.wfull(wr.idx), wr.phase);
            const int v_slot = wr.idx;
            const uint32_t v_addr = smem_base + QA_OFF_W + wr.idx * CT_TILE_BYTES;
            wr.advance();
            mbar_wait(B.oempty(g), ((ph_oempty >> g) & 1u) ^ 1u); ph_oempty ^= 1u << g;
            tc_fence_after();
            const uint32_t d_o = tmem_base + 400u + 32u * g;
            for (int c = 0; c < 4; ++c) {
              const int nh = c == 3 ? 1 : 2;        // keys 192..207 are one K-step (half 0 only); 208.. do not exist
              for (int hh = 0; hh < nh; ++hh) {
                mbar_wait(B.pfull(hh), (ph_pfull >> hh) & 1u); ph_pfull ^= 1u << hh;
                tc_fence_after();
                const uint64_t v_hi = umma_desc_sw128(v_addr + c * 8192), v_lo = umma_desc_sw128(v_addr + c * 8192 + 4096);
                const int k0 = 2 * hh, k1 = c == 3 ? 1 : 2 * hh + 2;
                for (int k = k0; k < k1; ++k) {
                  umma_bf16(d_o, e_hi + 2 * k, v_hi + 2 * k, idesc_o, (c > 0 || k > 0) ? 1u : 0u);
                  if (split) {
                    umma_bf16(d_o, e_lo + 2 * k, v_hi + 2 * k, idesc_o, 1u);
                    if (!pv2) umma_bf16(d_o, e_hi + 2 * k, v_lo + 2 * k, idesc_o, 1u);
                  }
                }
                umma_commit(B.pempty(hh));
              }
            }
            umma_commit(B.ofull(g));
            umma_commit(B.wempty(v_slot));
          }
        }
      }
    }
  } else {
    // ---------------- softmax / epilogue warps: thread = (query row, half) ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
    const int e = warp - 8, wq = e & 3, half = e >> 2;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t acc_tm = tmem_base + lane_off, s_tm = acc_tm + 192u, o_tm = acc_tm + 400u + 32u * half;
    uint8_t* qslot = smem_gen + QA_OFF_Q;
    uint8_t* pslot = smem_gen + QA_OFF_P;
    uint8_t* wscr = pslot + wq * 4096;            // transpose scratch of the finishing half: the P slot is idle then (see below)
    uint32_t ph_accfull = 0, ph_qempty = 0, ph_sfull = 0, ph_pempty = 0, ph_ofull = 0;
    const float sl2 = p.scale * 1.4426950408889634f;
    const int q8 = lane & 7;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int pr = 0; pr < 4; ++pr) {
        // ---- epi-1: this half's head of the pair out of the qkv accumulator ----
        const int hm = 2 * pr + half;
        float v[32];
        float s_self;
        mbar_wait(B.accfull(), ph_accfull); ph_accfull ^= 1;
        tc_fence_after();
        {
          uint32_t rq[32], rk[32];
          tmem_ld_32x32(acc_tm + 32u * half, rq);
          tmem_ld_32x32(acc_tm + 64u + 32u * half, rk);
          tmem_ld_wait();
          const float4* bq = reinterpret_cast<const float4*>(p.bias + hm * 32);
          const float4* bk = reinterpret_cast<const float4*>(p.bias + 256 + hm * 32);
          float2 d2 = make_float2(0.f, 0.f);
          mbar_wait(B.qempty(), ph_qempty ^ 1); ph_qempty ^= 1;     // the previous unit's score MMAs have read the q chunk
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float4 b4 = __ldg(bq + 2 * c + j), k4 = __ldg(bk + 2 * c + j);
              const int i = 8 * c + 4 * j;
              const float2 qa = add2(make_float2(__uint_as_float(rq[i]), __uint_as_float(rq[i + 1])), make_float2(b4.x, b4.y));
              const float2 qb = add2(make_float2(__uint_as_float(rq[i + 2]), __uint_as_float(rq[i + 3])), make_float2(b4.z, b4.w));
              const float2 ka = add2(make_float2(__uint_as_float(rk[i]), __uint_as_float(rk[i + 1])), make_float2(k4.x, k4.y));
              const float2 kb = add2(make_float2(__uint_as_float(rk[i + 2]), __uint_as_float(rk[i + 3])), make_float2(k4.z, k4.w));
              d2 = fma2(qa, ka, d2);
              d2 = fma2(qb, kb, d2);
              split_f16x2(qa.x, qa.y, hi[2 * j], lo[2 * j]);
              split_f16x2(qb.x, qb.y, hi[2 * j + 1], lo[2 * j + 1]);
            }
            const uint32_t off = swizzle128_offset(row, 4 * half + c);
            *reinterpret_cast<uint4*>(qslot + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (split) *reinterpret_cast<uint4*>(qslot + CT_A_HALF + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          s_self = d2.x + d2.y;
          fence_proxy_async_smem();
          mbar_arrive(B.qfull());
        }
        {
          uint32_t rv[32];
          tmem_ld_32x32(acc_tm + 128u + 32u * half, rv);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(B.accempty());
          const float4* bv = reinterpret_cast<const float4*>(p.bias + 512 + hm * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(bv + j);
            v[4 * j] = __uint_as_float(rv[4 * j]) + b4.x; v[4 * j + 1] = __uint_as_float(rv[4 * j + 1]) + b4.y;
            v[4 * j + 2] = __uint_as_float(rv[4 * j + 2]) + b4.z; v[4 * j + 3] = __uint_as_float(rv[4 * j + 3]) + b4.w;
          }
        }
        // ---- the two heads of the pair, one after the other (one S buffer); both halves share each head's keys ----
#pragma unroll 1
        for (int hs = 0; hs < 2; ++hs) {
          const int h = 2 * pr + hs;
          const bool mine = hs == half;
          mbar_wait(B.sfull(), ph_sfull); ph_sfull ^= 1;
          tc_fence_after();
          // pass 1: max over this half's keys (and the point's own key for the owner)
          float mx = mine ? s_self : -3.0e38f;
#pragma unroll 1
          for (int c = 0; c < 3; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(s_tm + c * 64 + 32 * half, rr);
            tmem_ld_wait();
            mx = attn_max_block<32>(rr, c * 64 + 32 * half, n_keys, mx);
          }
          if (half == 0) {
            uint32_t rr[16];
            tmem_ld_32x16(s_tm + 192, rr);
            tmem_ld_wait();
            mx = attn_max_block<16>(rr, 192, n_keys, mx);
          }
          xbuf[half * 128 + row] = mx;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          mx = fmaxf(mx, xbuf[(half ^ 1) * 128 + row]);
          const float mxs = mx * sl2;
          // pass 2: e_j -> the P slot, 32 keys per piece (half 0: 4 pieces, the last one 16 keys; half 1: 3 pieces)
          float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll 1
          for (int c = 0; c < 3; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(s_tm + c * 64 + 32 * half, rr);
            tmem_ld_wait();
            mbar_wait(B.pempty(half), ph_pempty ^ 1); ph_pempty ^= 1;
            attn_exp_block<16>(rr, c * 64 + 32 * half, n_keys, sl2, mxs, sum2, pslot, row, 4 * half, split);
            fence_proxy_async_smem();
            mbar_arrive(B.pfull(half));
          }
          if (half == 0) {
            uint32_t rr[16];
            tmem_ld_32x16(s_tm + 192, rr);
            tmem_ld_wait();
            mbar_wait(B.pempty(0), ph_pempty ^ 1); ph_pempty ^= 1;
            attn_exp_block<8>(rr, 192, n_keys, sl2, mxs, sum2, pslot, row, 0, split);
            fence_proxy_async_smem();
            mbar_arrive(B.pfull(0));
          }
          tc_fence_before();
          mbar_arrive(B.sempty());
          float sum = sum2.x + sum2.y, e_self = 0.f;
          if (mine) { e_self = fast_ex2(fmaf(s_self, sl2, -mxs)); sum += e_self; }
          xbuf[256 + half * 128 + row] = sum;
          asm volatile("bar.sync 2, 256;" ::: "memory");
          if (mine) {
            // this half owns head h: (O_h + e_self v_self) / sum.  The other half moves on to the next head, but writes the
            // P slot only after named barrier 1 of that head, which this half joins after the stores below; every P V MMA of
            // head h has completed (ofull): the slot is idle and serves as the transpose scratch.
            const float inv = 1.0f / (sum + xbuf[256 + (half ^ 1) * 128 + row]);
            mbar_wait(B.ofull(half), ph_ofull); ph_ofull ^= 1;
            tc_fence_after();
            {
              uint32_t rr[32];
              tmem_ld_32x32(o_tm, rr);
              tmem_ld_wait();
              tc_fence_before();
              mbar_arrive(B.oempty(half));
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(wscr + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                    make_float4(fmaf(e_self, v[4 * j], __uint_as_float(rr[4 * j])) * inv,
                                fmaf(e_self, v[4 * j + 1], __uint_as_float(rr[4 * j + 1])) * inv,
                                fmaf(e_self, v[4 * j + 2], __uint_as_float(rr[4 * j + 2])) * inv,
                                fmaf(e_self, v[4 * j + 3], __uint_as_float(rr[4 * j + 3])) * inv);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rl = 4 * i + (lane >> 3);
              const int mm = t * 128 + wq * 32 + rl;
              const float4 a = *reinterpret_cast<const float4*>(wscr + rl * 128 + ((q8 ^ (rl & 7)) << 4));
              if (mm < p.M) *(reinterpret_cast<float4*>(p.out + (int64_t)mm * p.ldo + h * 32) + q8) = a;
            }
            __syncwarp();
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}


// ===============================================================================================================
// Same operation as chain_qkvattn_kernel with the probabilities kept in TENSOR MEMORY: the softmax threads overwrite
// the fp32 scores of their 32-key blocks in place with the split-fp16 probabilities (32 score columns -> 16 packed hi
// columns + 16 packed lo columns, tcgen05.st) and O_h = P_h V_h takes its A operand from TMEM (tcgen05.mma [d], [a], b).
// One hand-over per head (pfull) replaces the seven P-slot round trips per head of the smem variant, P never crosses the
// shared-memory port, and the freed 32 KB are a 4th weight-ring slot (the MMA thread was exposed to every tile-load
// latency with 3).  S(h+1) overwrites the columns P(h) was read from: both are tcgen05.mma of the same issuing thread and
// execute in order, and every softmax thread has finished its loads / stores of the buffer before pfull(h).
// The attention output goes straight from registers to global memory (16 bytes per row and store; the rows' 128-byte
// segments merge in L2).
// Per head the W loader streams: K hi, K lo, [next unit's qkv weights: 2 K-chunks x (hi, lo)], V -- the order of use.
constexpr int QB_WSLOTS = 4;
constexpr int QB_OFF_W = 0;                                     // 4 x 32 KB
constexpr int QB_OFF_L = QB_OFF_W + QB_WSLOTS * CT_TILE_BYTES;  // 128 KB: 2 x (hi 16K | lo 16K) LN(x) chunks
constexpr int QB_OFF_Q = QB_OFF_L + 2 * 2 * CT_A_HALF;          // 192 KB: the unit's q chunk (hi | lo)
static_assert(QB_OFF_Q + 2 * CT_A_HALF == QA_OFF_BAR, "qkv-attention (TMEM P) smem layout");

struct QBars2 {
  uint32_t base;
  __device__ uint32_t wfull(int i) const { return base + 8u * i; }
  __device__ uint32_t wempty(int i) const { return base + 32u + 8u * i; }
  __device__ uint32_t lfull(int i) const { return base + 64u + 8u * i; }
  __device__ uint32_t lempty(int i) const { return base + 80u + 8u * i; }
  __device__ uint32_t accfull() const { return base + 96u; }
  __device__ uint32_t accempty() const { return base + 104u; }
  __device__ uint32_t qfull(int g) const { return base + 112u + 8u * g; }    // head g of the unit: written by half g
  __device__ uint32_t qempty(int g) const { return base + 128u + 8u * g; }
  __device__ uint32_t sfull() const { return base + 144u; }
  __device__ uint32_t pfull() const { return base + 152u; }
  __device__ uint32_t ofull() const { return base + 160u; }
  __device__ uint32_t oempty() const { return base + 168u; }
  __device__ uint32_t tmem_slot() const { return base + 176u; }
};

// e = exp2(s * sl2 - mxs) of NK scores (keys k0 ...), masked past n_keys; accumulates the row sum; the split-fp16 values go
// back into the scores' own TMEM columns: NK/2 packed hi columns at taddr, NK/2 packed lo columns behind them
template <int NK, bool MASKED>
__device__ __forceinline__ void attn_exp_tmem(const uint32_t* rr, int k0, int n_keys, float sl2, float mxs, float2& sum2,
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

// REGS: third generation of the softmax role (see the comment in its branch): every score is read from tensor memory once.
// Register pool of the REGS variant: setmaxnreg moves registers only inside the CTA's own launch allocation (USETMAXREG ...
// CTAPOOL) and warps are allocated in groups of four, so a 576-thread CTA gets 20 x 32 x 96 registers but can use only 18 warps'
// worth.  The variant is launched with 640 threads: warps 18-19 exist only to complete the fifth warpgroup, which hands back
// its registers as a whole; 8 loader warps x 56 + 8 softmax warps x 152 + 4 x 56 = 1888 <= 20 x 96 = 1920.
constexpr int QA3_THREADS = 640;
template <bool REGS>
__global__ void __maxnreg__(96) chain_qkvattn2_kernel(ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  if (smem_base - smem_u32(smem_raw) > CT_SMEM - CT_SMEM_USED) __trap();   // dynamic smem base less aligned than budgeted
  QBars2 B{smem_base + QA_OFF_BAR};
  float* xbuf = reinterpret_cast<float*>(smem_gen + QA_OFF_X);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = p.precision == 0;
  const bool kv1 = split && (p.flags & 1), s2 = split && (p.flags & 2), pv2 = split && (p.flags & 4);
  const int n_tiles = (p.M + 127) / 128;
  const int n_keys = p.n_keys;

  if (threadIdx.x == 0) {
    for (int i = 0; i < QB_WSLOTS; ++i) { mbar_init(B.wfull(i), 1); mbar_init(B.wempty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(B.lfull(i), 256); mbar_init(B.lempty(i), 1); }
    mbar_init(B.accfull(), 1); mbar_init(B.accempty(), 256);
    for (int i = 0; i < 2; ++i) { mbar_init(B.qfull(i), 128); mbar_init(B.qempty(i), 1); }
    mbar_init(B.sfull(), 1); mbar_init(B.pfull(), 256);
    mbar_init(B.ofull(), 1); mbar_init(B.oempty(), REGS ? 256 : 128);
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(B.tmem_slot(), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (B.tmem_slot() - smem_base));

  if constexpr (REGS) {
    if (warp >= 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");     // MMA, W loader and the two filler warps
  }
  if (warp < 8) {
    // ---------------- loaders: LN(x) chunks, 4 units x 4 K-chunks per tile ----------------
    if constexpr (REGS) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    Ring lr(2);
    float4 buf[8];
    int tn = 0;
    const bool tr0 = threadIdx.x == 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = t * 128 + warp * 16;
      float sc = 1.f, sh = 0.f;
      if (tr0) trace_ev(p.trace, 1, tn, 14);
      // the REGS variant does NOT prefetch the next tile into L2: measured (ncu, 129^3 pass) 2.69 GB of DRAM reads with the prefetch
      // against the algorithmic 2.20 GB without, at the same run time (the prefetched lines do not survive a whole tile time)
      if (!REGS && !(p.flags & 32) && t + (int)gridDim.x < n_tiles) prefetch_rows16_l2(p.x, p.ldx, m0 + (int)gridDim.x * 128, p.M, 0, lane);
      const bool pts = REGS && (p.flags & 64);       // x = LinearProj3D(points), recomputed (first block)
      float px = 0.f, py = 0.f, pz = 0.f;
      if (pts) {
        sc = 0.f; sh = 0.f;
        if (lane < 16 && m0 + lane < p.M) {
          const float* pt = p.points + (int64_t)(m0 + lane) * 3;
          px = __ldg(pt); py = __ldg(pt + 1); pz = __ldg(pt + 2);
          pts_ln_factors(p.pp_stat, px, py, pz, p.ln_eps, sc, sh);
        }
      } else {
        warp_ln_stats16(p.x, p.ldx, m0, p.M, p.ln_eps, lane, sc, sh);
        fetch_chunk16(p.x, p.ldx, m0, p.M, 0, lane, buf);
      }
      if (tr0) trace_ev(p.trace, 1, tn, 15);
      for (int i = 0; i < 16; ++i) {
        if (tr0) trace_ev(p.trace, 1, tn, 10);
        mbar_wait(B.lempty(lr.idx), lr.phase ^ 1);
        if (tr0) trace_ev(p.trace, 1, tn, 11);
        if (pts) store_chunk16_pts(smem_gen + QB_OFF_L + lr.idx * 2 * CT_A_HALF, warp, lane, i & 3, px, py, pz, sc, sh, p.pp, split);
        else store_chunk16(smem_gen + QB_OFF_L + lr.idx * 2 * CT_A_HALF, warp, lane, buf, true, sc, sh, split);
        fence_proxy_async_smem();
        mbar_arrive(B.lfull(lr.idx));
        if (tr0) trace_ev(p.trace, 1, tn, 12);
        lr.advance();
        if (!pts && i + 1 < 16) fetch_chunk16(p.x, p.ldx, m0, p.M, ((i + 1) & 3) * 64, lane, buf);
      }
    }
  } else if (warp == 17) {
    // ---------------- W loader ----------------
    if (lane == 0) {
      Ring wr(QB_WSLOTS);
      auto put = [&](const uint8_t* src, uint32_t bytes) {
        mbar_wait(B.wempty(wr.idx), wr.phase ^ 1);
        mbar_arrive_expect_tx(B.wfull(wr.idx), bytes);
        bulk_g2s(smem_base + QB_OFF_W + wr.idx * CT_TILE_BYTES, src, bytes, B.wfull(wr.idx));
        wr.advance();
      };
      auto qkv_chunk = [&](int pr, int kc) {
        const uint8_t* src = p.blob + ((size_t)pr * 4 + kc) * 2 * CT_TILE_BYTES;
        put(src, QA_QKV_TILE_BYTES);
        if (split) put(src + CT_TILE_BYTES, kv1 ? QA_Q_TILE_BYTES : QA_QKV_TILE_BYTES);
      };
      bool first = true;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int pr = 0; pr < 4; ++pr) {
          if (first) { for (int kc = 0; kc < 4; ++kc) qkv_chunk(0, kc); first = false; }
          const bool has_next = pr < 3 || t + (int)gridDim.x < n_tiles;
          const int pn = (pr + 1) & 3;
          for (int g = 0; g < 2; ++g) {
            put(p.kblob + (size_t)pr * 2 * CT_TILE_BYTES, QA_K_TILE_BYTES);
            if (split && !s2) put(p.kblob + (size_t)pr * 2 * CT_TILE_BYTES + CT_TILE_BYTES, QA_K_TILE_BYTES);
            if (has_next) { qkv_chunk(pn, 2 * g); qkv_chunk(pn, 2 * g + 1); }
            put(p.vblob + (size_t)(2 * pr + g) * CT_TILE_BYTES, CT_TILE_BYTES);
          }
        }
      }
    }
  } else if (warp == 16) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      Ring wr(QB_WSLOTS), lr(2);
      uint32_t ph_accempty = 0, ph_qfull = 0, ph_pfull = 0, ph_oempty = 0;   // ph_qfull: bit g = phase of qfull(g)
      const uint32_t idesc_qkv = umma_idesc_f16(128, 192), idesc_q2 = kv1 ? umma_idesc_f16(128, 64) : idesc_qkv;
      const uint32_t idesc_s = umma_idesc_f16(128, 208), idesc_o = umma_idesc_f16(128, 32), idesc_o2 = umma_idesc_f16(128, 64);
      const uint32_t d_acc = tmem_base, d_s = tmem_base + 192, d_o = tmem_base + 400;
      const uint64_t q_hi = umma_desc_sw128(smem_base + QB_OFF_Q), q_lo = umma_desc_sw128(smem_base + QB_OFF_Q + CT_A_HALF);
      int tn = 0;
      auto qkv_chunk = [&](int kc) {
        trace_ev(p.trace, 0, tn, 1);
        mbar_wait(B.lfull(lr.idx), lr.phase);
        trace_ev(p.trace, 0, tn, 2);
        tc_fence_after();
        const uint32_t a_addr = smem_base + QB_OFF_L + lr.idx * 2 * CT_A_HALF;
        const uint64_t a_hi = umma_desc_sw128(a_addr), a_lo = umma_desc_sw128(a_addr + CT_A_HALF);
        mbar_wait(B.wfull(wr.idx), wr.phase);
        tc_fence_after();
        {
          const uint64_t w = umma_desc_sw128(smem_base + QB_OFF_W + wr.idx * CT_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(d_acc, a_hi + 2 * k, w + 2 * k, idesc_qkv, (kc > 0 || k > 0) ? 1u : 0u);
            if (split) umma_bf16(d_acc, a_lo + 2 * k, w + 2 * k, idesc_q2, 1u);
          }
          umma_commit(B.wempty(wr.idx));
          wr.advance();
        }
        if (split) {
          mbar_wait(B.wfull(wr.idx), wr.phase);
          tc_fence_after();
          const uint64_t w = umma_desc_sw128(smem_base + QB_OFF_W + wr.idx * CT_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_acc, a_hi + 2 * k, w + 2 * k, idesc_q2, 1u);
          umma_commit(B.wempty(wr.idx));
          wr.advance();
        }
        umma_commit(B.lempty(lr.idx));
        lr.advance();
        trace_ev(p.trace, 0, tn, 3);
      };
      mbar_wait(B.accempty(), ph_accempty ^ 1); ph_accempty ^= 1;
      tc_fence_after();
      for (int kc = 0; kc < 4; ++kc) qkv_chunk(kc);
      umma_commit(B.accfull());
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int pr = 0; pr < 4; ++pr) {
          const bool has_next = pr < 3 || t + (int)gridDim.x < n_tiles;
          for (int g = 0; g < 2; ++g) {
            // ---- S = Q_h K_h^T (overwrites the probabilities of the previous head: in-order tensor pipe) ----
            trace_ev(p.trace, 0, tn, 4);
            mbar_wait(B.qfull(g), (ph_qfull >> g) & 1u); ph_qfull ^= 1u << g;     // half g has written this head's q (epi-1)
            mbar_wait(B.wfull(wr.idx), wr.phase);
            tc_fence_after();
            {
              const uint64_t w = umma_desc_sw128(smem_base + QB_OFF_W + wr.idx * CT_TILE_BYTES);
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const int kk = 2 * g + k;
                umma_bf16(d_s, q_hi + 2 * kk, w + 2 * kk, idesc_s, k > 0 ? 1u : 0u);
                if (split) umma_bf16(d_s, q_lo + 2 * kk, w + 2 * kk, idesc_s, 1u);
              }
              umma_commit(B.wempty(wr.idx));
              wr.advance();
            }
            if (split && !s2) {
              mbar_wait(B.wfull(wr.idx), wr.phase);
              tc_fence_after();
              const uint64_t w = umma_desc_sw128(smem_base + QB_OFF_W + wr.idx * CT_TILE_BYTES);
#pragma unroll
              for (int k = 0; k < 2; ++k) umma_bf16(d_s, q_hi + 2 * (2 * g + k), w + 2 * (2 * g + k), idesc_s, 1u);
              umma_commit(B.wempty(wr.idx));
              wr.advance();
            }
            umma_commit(B.sfull());
            umma_commit(B.qempty(g));
            trace_ev(p.trace, 0, tn, 5);
            // ---- the next unit's qkv GEMM, two K-chunks per head, while the softmax of this head runs ----
            if (has_next) {
              if (g == 0) { mbar_wait(B.accempty(), ph_accempty ^ 1); ph_accempty ^= 1; tc_fence_after(); }
              qkv_chunk(2 * g);
              qkv_chunk(2 * g + 1);
              if (g == 1) umma_commit(B.accfull());
            }
            // ---- O_h = P_h V_h: A = the probabilities in TMEM (in the score columns), B = the head's value tile ----
            mbar_wait(B.wfull(wr.idx), wr.phase);
            const int v_slot = wr.idx;
            const uint32_t v_addr = smem_base + QB_OFF_W + wr.idx * CT_TILE_BYTES;
            wr.advance();
            trace_ev(p.trace, 0, tn, 6);
            mbar_wait(B.oempty(), ph_oempty ^ 1); ph_oempty ^= 1;
            mbar_wait(B.pfull(), ph_pfull); ph_pfull ^= 1;
            trace_ev(p.trace, 0, tn, 7);
            tc_fence_after();
            // an MMA whose A operand comes from tensor memory costs ~64 cycles whatever its N (the 128 x 16 A block is
            // fetched per instruction), so Ph Vh and Ph Vl are ONE instruction with N = 64: the value tile chunk is stored
            // as [hi rows 0..31 | lo rows 32..63] = one 64-row operand tile, and the accumulator holds Ph Vh (+ Pl Vh) in
            // columns [0,32) and Ph Vl in [32,64); the finishing threads add the two halves
            const uint32_t idesc_hi = (split && !pv2) ? idesc_o2 : idesc_o;
#pragma unroll 1
            for (int j = 0; j < 13; ++j) {                 // K-step j = keys 16j .. 16j+15
              const uint32_t a_hi = d_s + 32u * (j >> 1) + 8u * (j & 1);
              const uint32_t a_lo = a_hi + (j == 12 ? 8u : 16u);
              const uint64_t vv = umma_desc_sw128(v_addr + (j >> 2) * 8192) + 2 * (j & 3);
              umma_ts(d_o, a_hi, vv, idesc_hi, j > 0 ? 1u : 0u);
              if (split) umma_ts(d_o, a_lo, vv, idesc_o, 1u);
            }
            umma_commit(B.ofull());
            umma_commit(B.wempty(v_slot));
            trace_ev(p.trace, 0, tn, 8);
          }
        }
      }
    }
  } else if (warp >= 18) {
    // filler warps of the REGS variant: nothing to do
  } else if constexpr (REGS) {
    // ---------------- softmax / epilogue warps, scores in REGISTERS: thread = (query row, half) ----------------
    // Tensor memory is read at 64 B/clk per SM: the 128 x 208 fp32 scores of a head take ~1.7k cycles per sweep, and the
    // two sweeps (max, then exp) of the branch below were ~3.3k of the ~11k-cycle critical path of a head.  Here a thread
    // loads its 104 (112) scores once, keeps them in registers across the max exchange and converts them in place.  The
    // two threads of a row meet at a 64-thread named barrier (their two warps) instead of a 256-thread one, and BOTH
    // finish the head (16 of its 32 output columns each), which halves the serial tail O read -> normalise -> store.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    int tn = 0;
    const bool tr0 = (warp == 8 && lane == 0);
    const int e = warp - 8, wq = e & 3, half = e >> 2;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t acc_tm = tmem_base + lane_off, s_tm = acc_tm + 192u, o_tm = acc_tm + 400u;
    const bool o_two = split && !pv2;
    uint8_t* qslot = smem_gen + QB_OFF_Q;
    uint32_t ph_accfull = 0, ph_qempty = 0, ph_sfull = 0, ph_ofull = 0;
    const float sl2 = p.scale * 1.4426950408889634f;
    const int bar_max = 1 + wq, bar_sum = 5 + wq;
    float s_self = 0.f;
    float* oblk_next = nullptr;
    // epi-1 of unit (tile t, head pair pr): runs one unit AHEAD of the heads loop -- for the first unit of the CTA in the
    // prologue, afterwards in the shadow of the P V MMAs of the previous unit's second head (its accumulator has been
    // ready since the softmax of that head began), so that the S MMA of the next head never waits for q.
    auto epi1 = [&](int t, int pr) {
        // ---- epi-1: q and the self score of this half's head; 16 value columns of BOTH heads ----
        const int hm = 2 * pr + half;
        // this thread's four 16-byte chunks of head 2 pr in the tile-blocked output (see zs_chain_qkvattn_fwd): chunk j of
        // row r lives at float4 index (tile * 64 + j) * 128 + r, so a warp's 32 rows are 512 contiguous bytes per chunk
        float* const oblk = p.out + ((size_t)t * 8192 + row) * 4 + (size_t)((2 * pr) * 8 + 4 * half) * 512;
        oblk_next = oblk;
        if (tr0) trace_ev(p.trace, 2, tn, 20);
        mbar_wait(B.accfull(), ph_accfull); ph_accfull ^= 1;
        if (tr0) trace_ev(p.trace, 2, tn, 21);
        tc_fence_after();
        {
          uint32_t rq[32], rk[32];
          tmem_ld_32x32(acc_tm + 32u * half, rq);
          tmem_ld_32x32(acc_tm + 64u + 32u * half, rk);
          tmem_ld_wait();
          const float4* bq = reinterpret_cast<const float4*>(p.bias + hm * 32);
          const float4* bk = reinterpret_cast<const float4*>(p.bias + 256 + hm * 32);
          float2 d2 = make_float2(0.f, 0.f);
          mbar_wait(B.qempty(half), ph_qempty ^ 1); ph_qempty ^= 1;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float4 b4 = __ldg(bq + 2 * c + j), k4 = __ldg(bk + 2 * c + j);
              const int i = 8 * c + 4 * j;
              const float2 qa = add2(make_float2(__uint_as_float(rq[i]), __uint_as_float(rq[i + 1])), make_float2(b4.x, b4.y));
              const float2 qb = add2(make_float2(__uint_as_float(rq[i + 2]), __uint_as_float(rq[i + 3])), make_float2(b4.z, b4.w));
              const float2 ka = add2(make_float2(__uint_as_float(rk[i]), __uint_as_float(rk[i + 1])), make_float2(k4.x, k4.y));
              const float2 kb = add2(make_float2(__uint_as_float(rk[i + 2]), __uint_as_float(rk[i + 3])), make_float2(k4.z, k4.w));
              d2 = fma2(qa, ka, d2);
              d2 = fma2(qb, kb, d2);
              split_f16x2(qa.x, qa.y, hi[2 * j], lo[2 * j]);
              split_f16x2(qb.x, qb.y, hi[2 * j + 1], lo[2 * j + 1]);
            }
            const uint32_t off = swizzle128_offset(row, 4 * half + c);
            *reinterpret_cast<uint4*>(qslot + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (split) *reinterpret_cast<uint4*>(qslot + CT_A_HALF + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          s_self = d2.x + d2.y;
          fence_proxy_async_smem();
          mbar_arrive(B.qfull(half));
        }
        {
          uint32_t ra[16], rb[16];
          tmem_ld_32x16(acc_tm + 128u + 16u * half, ra);
          tmem_ld_32x16(acc_tm + 160u + 16u * half, rb);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(B.accempty());
          // the points' own value rows (v + bias) wait in the OUTPUT buffer, at the very addresses the finished head will
          // overwrite: 32 registers less across the two softmax passes (which need them for instruction-level parallelism)
          const float4* ba = reinterpret_cast<const float4*>(p.bias + 512 + (2 * pr) * 32 + 16 * half);
          const float4* bb = reinterpret_cast<const float4*>(p.bias + 512 + (2 * pr + 1) * 32 + 16 * half);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 a4 = __ldg(ba + j), b4 = __ldg(bb + j);
            *reinterpret_cast<float4*>(oblk + j * 512) =
                make_float4(__uint_as_float(ra[4 * j]) + a4.x, __uint_as_float(ra[4 * j + 1]) + a4.y,
                            __uint_as_float(ra[4 * j + 2]) + a4.z, __uint_as_float(ra[4 * j + 3]) + a4.w);
            *reinterpret_cast<float4*>(oblk + (8 + j) * 512) =
                make_float4(__uint_as_float(rb[4 * j]) + b4.x, __uint_as_float(rb[4 * j + 1]) + b4.y,
                            __uint_as_float(rb[4 * j + 2]) + b4.z, __uint_as_float(rb[4 * j + 3]) + b4.w);
          }
        }
    };
    if ((int)blockIdx.x < n_tiles) epi1(blockIdx.x, 0);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int pr = 0; pr < 4; ++pr) {
        float* const oblk = oblk_next;
        // ---- the two heads of the pair; the halves split each head's keys (half 0: 32-key blocks 0,2,4,6; half 1: 1,3,5) ----
#pragma unroll 1
        for (int hs = 0; hs < 2; ++hs) {
          const int h = 2 * pr + hs;
          const bool mine = hs == half;
          if (tr0) trace_ev(p.trace, 2, tn, 22);
          mbar_wait(B.sfull(), ph_sfull); ph_sfull ^= 1;
          if (tr0) trace_ev(p.trace, 2, tn, 23);
          tc_fence_after();
          uint32_t sa[32], sb[32], sc3[32], sd[16];
          const int ka = 32 * half, kb = 64 + 32 * half, kc3 = 128 + 32 * half;   // first keys of this thread's blocks
          tmem_ld_32x32(s_tm + ka, sa);
          tmem_ld_32x32(s_tm + kb, sb);
          tmem_ld_32x32(s_tm + kc3, sc3);
          if (half == 0) tmem_ld_32x16(s_tm + 192, sd);
          tmem_ld_wait();
          float mx = mine ? s_self : -3.0e38f;
          mx = attn_max_block<32>(sa, ka, n_keys, mx);
          mx = attn_max_block<32>(sb, kb, n_keys, mx);
          mx = attn_max_block<32>(sc3, kc3, n_keys, mx);
          if (half == 0) mx = attn_max_block<16>(sd, 192, n_keys, mx);
          xbuf[half * 128 + row] = mx;
          if (mine) xbuf[256 + (half ^ 1) * 128 + row] = s_self;      // the other thread of the row reads it from ITS sum slot
          if (tr0) trace_ev(p.trace, 2, tn, 24);
          asm volatile("bar.sync %0, 64;" ::"r"(bar_max) : "memory");
          if (tr0) trace_ev(p.trace, 2, tn, 25);
          mx = fmaxf(mx, xbuf[(half ^ 1) * 128 + row]);
          const float ss = mine ? s_self : xbuf[256 + half * 128 + row];
          const float mxs = mx * sl2;
          float2 sum2 = make_float2(0.f, 0.f);
          // a block is masked only when the key count ends inside it (CTA-uniform)
          if (ka + 32 <= n_keys) attn_exp_tmem<32, false>(sa, ka, n_keys, sl2, mxs, sum2, s_tm + ka, split);
          else attn_exp_tmem<32, true>(sa, ka, n_keys, sl2, mxs, sum2, s_tm + ka, split);
          if (kb + 32 <= n_keys) attn_exp_tmem<32, false>(sb, kb, n_keys, sl2, mxs, sum2, s_tm + kb, split);
          else attn_exp_tmem<32, true>(sb, kb, n_keys, sl2, mxs, sum2, s_tm + kb, split);
          if (kc3 + 32 <= n_keys) attn_exp_tmem<32, false>(sc3, kc3, n_keys, sl2, mxs, sum2, s_tm + kc3, split);
          else attn_exp_tmem<32, true>(sc3, kc3, n_keys, sl2, mxs, sum2, s_tm + kc3, split);
          if (half == 0) attn_exp_tmem<16, true>(sd, 192, n_keys, sl2, mxs, sum2, s_tm + 192, split);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(B.pfull());
          if (tr0) trace_ev(p.trace, 2, tn, 26);
          const float e_self = fast_ex2(fmaf(ss, sl2, -mxs));
          const float sum = sum2.x + sum2.y + (mine ? e_self : 0.f);
          xbuf[256 + half * 128 + row] = sum;
          asm volatile("bar.sync %0, 64;" ::"r"(bar_sum) : "memory");
          if (tr0) trace_ev(p.trace, 2, tn, 27);
          const float inv = 1.0f / (sum + xbuf[256 + (half ^ 1) * 128 + row]);
          if (hs == 1) {          // s_self of this unit is dead (e_self is computed): the next unit's epi-1 may overwrite it
            const int t2 = pr < 3 ? t : t + (int)gridDim.x;     // one call site: the role's code must stay small (instruction cache)
            if (t2 < n_tiles) epi1(t2, (pr + 1) & 3);
          }
          float4 vs[4];                                        // the parked value chunks of this head (L2 hits, issued early)
          float* const oh = oblk + hs * (8 * 512);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(vs[j].x), "=f"(vs[j].y), "=f"(vs[j].z), "=f"(vs[j].w) : "l"(oh + j * 512) : "memory");
          mbar_wait(B.ofull(), ph_ofull); ph_ofull ^= 1;
          if (tr0) trace_ev(p.trace, 2, tn, 28);
          tc_fence_after();
          {
            uint32_t rr[16], r2[16];
            tmem_ld_32x16(o_tm + 16u * half, rr);
            if (o_two) tmem_ld_32x16(o_tm + 32u + 16u * half, r2);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(B.oempty());
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float a0 = __uint_as_float(rr[4 * j]), a1 = __uint_as_float(rr[4 * j + 1]);
              float a2 = __uint_as_float(rr[4 * j + 2]), a3 = __uint_as_float(rr[4 * j + 3]);
              if (o_two) {
                a0 += __uint_as_float(r2[4 * j]); a1 += __uint_as_float(r2[4 * j + 1]);
                a2 += __uint_as_float(r2[4 * j + 2]); a3 += __uint_as_float(r2[4 * j + 3]);
              }
              *reinterpret_cast<float4*>(oh + j * 512) =
                  make_float4(fmaf(e_self, vs[j].x, a0) * inv, fmaf(e_self, vs[j].y, a1) * inv,
                              fmaf(e_self, vs[j].z, a2) * inv, fmaf(e_self, vs[j].w, a3) * inv);
            }
          }
          if (tr0) trace_ev(p.trace, 2, tn, 29);
        }
      }
    }
  } else {
    // ---------------- softmax / epilogue warps: thread = (query row, half) ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
    int tn = 0;
    const bool tr0 = (warp == 8 && lane == 0);
    const int e = warp - 8, wq = e & 3, half = e >> 2;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t acc_tm = tmem_base + lane_off, s_tm = acc_tm + 192u, o_tm = acc_tm + 400u;
    const bool o_two = split && !pv2;             // the accumulator carries Ph Vl in a second 32-column half
    uint8_t* qslot = smem_gen + QB_OFF_Q;
    uint32_t ph_accfull = 0, ph_qempty = 0, ph_sfull = 0, ph_ofull = 0;
    const float sl2 = p.scale * 1.4426950408889634f;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int mrow = t * 128 + row;
      for (int pr = 0; pr < 4; ++pr) {
        // ---- epi-1: this half's head of the pair out of the qkv accumulator ----
        const int hm = 2 * pr + half;
        float v[32];
        float s_self;
        if (tr0) trace_ev(p.trace, 2, tn, 20);
        mbar_wait(B.accfull(), ph_accfull); ph_accfull ^= 1;
        if (tr0) trace_ev(p.trace, 2, tn, 21);
        tc_fence_after();
        {
          uint32_t rq[32], rk[32];
          tmem_ld_32x32(acc_tm + 32u * half, rq);
          tmem_ld_32x32(acc_tm + 64u + 32u * half, rk);
          tmem_ld_wait();
          const float4* bq = reinterpret_cast<const float4*>(p.bias + hm * 32);
          const float4* bk = reinterpret_cast<const float4*>(p.bias + 256 + hm * 32);
          float2 d2 = make_float2(0.f, 0.f);
          mbar_wait(B.qempty(half), ph_qempty ^ 1); ph_qempty ^= 1;     // the previous unit's S MMA of this head slot has read its q
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float4 b4 = __ldg(bq + 2 * c + j), k4 = __ldg(bk + 2 * c + j);
              const int i = 8 * c + 4 * j;
              const float2 qa = add2(make_float2(__uint_as_float(rq[i]), __uint_as_float(rq[i + 1])), make_float2(b4.x, b4.y));
              const float2 qb = add2(make_float2(__uint_as_float(rq[i + 2]), __uint_as_float(rq[i + 3])), make_float2(b4.z, b4.w));
              const float2 ka = add2(make_float2(__uint_as_float(rk[i]), __uint_as_float(rk[i + 1])), make_float2(k4.x, k4.y));
              const float2 kb = add2(make_float2(__uint_as_float(rk[i + 2]), __uint_as_float(rk[i + 3])), make_float2(k4.z, k4.w));
              d2 = fma2(qa, ka, d2);
              d2 = fma2(qb, kb, d2);
              split_f16x2(qa.x, qa.y, hi[2 * j], lo[2 * j]);
              split_f16x2(qb.x, qb.y, hi[2 * j + 1], lo[2 * j + 1]);
            }
            const uint32_t off = swizzle128_offset(row, 4 * half + c);
            *reinterpret_cast<uint4*>(qslot + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (split) *reinterpret_cast<uint4*>(qslot + CT_A_HALF + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          s_self = d2.x + d2.y;
          fence_proxy_async_smem();
          mbar_arrive(B.qfull(half));
        }
        {
          uint32_t rv[32];
          tmem_ld_32x32(acc_tm + 128u + 32u * half, rv);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(B.accempty());
          const float4* bv = reinterpret_cast<const float4*>(p.bias + 512 + hm * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(bv + j);
            v[4 * j] = __uint_as_float(rv[4 * j]) + b4.x; v[4 * j + 1] = __uint_as_float(rv[4 * j + 1]) + b4.y;
            v[4 * j + 2] = __uint_as_float(rv[4 * j + 2]) + b4.z; v[4 * j + 3] = __uint_as_float(rv[4 * j + 3]) + b4.w;
          }
        }
        // ---- the two heads of the pair; both halves share each head's keys (half 0: 32-key blocks 0,2,4,6; half 1: 1,3,5) ----
#pragma unroll 1
        for (int hs = 0; hs < 2; ++hs) {
          const int h = 2 * pr + hs;
          const bool mine = hs == half;
          if (tr0) trace_ev(p.trace, 2, tn, 22);
          mbar_wait(B.sfull(), ph_sfull); ph_sfull ^= 1;
          if (tr0) trace_ev(p.trace, 2, tn, 23);
          tc_fence_after();
          float mx = mine ? s_self : -3.0e38f;
#pragma unroll 1
          for (int c = 0; c < 3; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(s_tm + c * 64 + 32 * half, rr);
            tmem_ld_wait();
            mx = attn_max_block<32>(rr, c * 64 + 32 * half, n_keys, mx);
          }
          if (half == 0) {
            uint32_t rr[16];
            tmem_ld_32x16(s_tm + 192, rr);
            tmem_ld_wait();
            mx = attn_max_block<16>(rr, 192, n_keys, mx);
          }
          xbuf[half * 128 + row] = mx;
          if (tr0) trace_ev(p.trace, 2, tn, 24);
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (tr0) trace_ev(p.trace, 2, tn, 25);
          mx = fmaxf(mx, xbuf[(half ^ 1) * 128 + row]);
          const float mxs = mx * sl2;
          float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll 1
          for (int c = 0; c < 3; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(s_tm + c * 64 + 32 * half, rr);
            tmem_ld_wait();
            const int k0 = c * 64 + 32 * half;     // a block is masked only when the key count ends inside it (CTA-uniform)
            if (k0 + 32 <= n_keys) attn_exp_tmem<32, false>(rr, k0, n_keys, sl2, mxs, sum2, s_tm + k0, split);
            else attn_exp_tmem<32, true>(rr, k0, n_keys, sl2, mxs, sum2, s_tm + k0, split);
          }
          if (half == 0) {
            uint32_t rr[16];
            tmem_ld_32x16(s_tm + 192, rr);
            tmem_ld_wait();
            attn_exp_tmem<16, true>(rr, 192, n_keys, sl2, mxs, sum2, s_tm + 192, split);
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(B.pfull());
          if (tr0) trace_ev(p.trace, 2, tn, 26);
          float sum = sum2.x + sum2.y, e_self = 0.f;
          if (mine) { e_self = fast_ex2(fmaf(s_self, sl2, -mxs)); sum += e_self; }
          xbuf[256 + half * 128 + row] = sum;
          asm volatile("bar.sync 2, 256;" ::: "memory");
          if (tr0) trace_ev(p.trace, 2, tn, 27);
          if (mine) {
            const float inv = 1.0f / (sum + xbuf[256 + (half ^ 1) * 128 + row]);
            mbar_wait(B.ofull(), ph_ofull);
            if (tr0) trace_ev(p.trace, 2, tn, 28);
            tc_fence_after();
            float4* dst = reinterpret_cast<float4*>(p.out + (int64_t)mrow * p.ldo + h * 32);
#pragma unroll
            for (int q = 0; q < 2; ++q) {            // 16 output columns at a time (register budget)
              uint32_t rr[16], r2[16];
              tmem_ld_32x16(o_tm + 16u * q, rr);
              if (o_two) tmem_ld_32x16(o_tm + 32u + 16u * q, r2);
              tmem_ld_wait();
              if (q == 1) { tc_fence_before(); mbar_arrive(B.oempty()); }
              if (mrow < p.M) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float o[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    float a = __uint_as_float(rr[4 * j + i]);
                    if (o_two) a += __uint_as_float(r2[4 * j + i]);
                    o[i] = fmaf(e_self, v[16 * q + 4 * j + i], a) * inv;
                  }
                  dst[4 * q + j] = make_float4(o[0], o[1], o[2], o[3]);
                }
              }
            }
            if (tr0) trace_ev(p.trace, 2, tn, 29);
          }
          ph_ofull ^= 1;      // one accumulator, one barrier: a phase per head, whichever half owns it
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}


// ===============================================================================================================
// Second generation of the two chained MLP kernels: the activations that feed the NEXT GEMM stay in TENSOR MEMORY.
// The epilogue threads overwrite each 32-column block of the fp32 accumulator in place with the split-fp16 activation
// (16 packed hi columns | 16 packed lo columns, tcgen05.st), and the next layer's MMAs take that block as their A operand
// from TMEM (tcgen05.mma [d], [a], b) -- the construction validated in chain_qkvattn2_kernel.  Versus ring E:
//   * the activations never cross the shared-memory port (the N = 256 MMAs read 12 KB of smem operands per 128 cycles
//     = 96 of the port's 128 B/clk; ring-E chunks added 32 KB written + 48 KB read per 64-wide K-chunk);
//   * the 64 KB of ring E become weight-ring slots (3 -> 4 for the MLP, 5 for the occupancy MLP): the weight stream is
//     latency-bound (a slot is released only when the MMAs that read it have completed), so depth is throughput;
//   * no eempty hand-back: the accumulator half is protected by the in-order tensor pipe and tempty.
// A TS MMA costs >= 64 cycles whatever its N (the 128 x 16 A block is fetched per instruction); at N = 256 that is free.
constexpr int M2_WSLOTS = 4, O2_WSLOTS = 5;
constexpr int M2_OFF_L = M2_WSLOTS * CT_TILE_BYTES;            // 128 KB
constexpr int M2_OFF_T = M2_OFF_L + CT_LSLOTS * 2 * CT_A_HALF; // 192 KB: 8 x 4 KB transpose scratch of the final epilogue
constexpr int O2_OFF_L = O2_WSLOTS * CT_TILE_BYTES;            // 160 KB
static_assert(M2_OFF_T + 8 * 4096 == CT_OFF_BAR && O2_OFF_L + CT_LSLOTS * 2 * CT_A_HALF == CT_OFF_BAR, "chain2 smem layout");

struct Bars5 {   // up to 5 weight slots
  uint32_t base;
  __device__ uint32_t wfull(int i) const { return base + 8u * i; }
  __device__ uint32_t wempty(int i) const { return base + 40u + 8u * i; }
  __device__ uint32_t lfull(int i) const { return base + 80u + 8u * i; }
  __device__ uint32_t lempty(int i) const { return base + 96u + 8u * i; }
  // one barrier per 64-column activation chunk: a single barrier would let the epilogue run two phases ahead of the MMA
  // thread (parity waits cannot tell phase k from k + 2); per chunk it is at most one, because the next layer's tfull needs
  // the MMA thread to have consumed all four
  __device__ uint32_t efull(int c) const { return base + 112u + 8u * c; }
  __device__ uint32_t tfull(int i) const { return base + 144u + 8u * i; }
  __device__ uint32_t tempty(int i) const { return base + 160u + 8u * i; }
  __device__ uint32_t tmem_slot() const { return base + 176u; }
};

__device__ __forceinline__ uint32_t chain2_setup(const Bars5& B, uint8_t* smem_gen, uint32_t smem_base, int warp, int wslots) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < wslots; ++i) { mbar_init(B.wfull(i), 1); mbar_init(B.wempty(i), 1); }
    for (int i = 0; i < CT_LSLOTS; ++i) { mbar_init(B.lfull(i), 128); mbar_init(B.lempty(i), 1); }
    for (int i = 0; i < 4; ++i) mbar_init(B.efull(i), 256);
    for (int i = 0; i < 2; ++i) { mbar_init(B.tfull(i), 1); mbar_init(B.tempty(i), 256); }
    fence_mbar_init();
  }
  if (warp == 12) tmem_alloc(B.tmem_slot(), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(smem_gen + (B.tmem_slot() - smem_base));
}

// MMA thread: one 64-wide K-chunk against the next weight tile pair; A from shared memory (a_addr) or, when `ts`,
// from tensor memory (two 32-value blocks of [16 hi | 16 lo] columns: K-step kk at a_tmem + 32 (kk >> 1) + 8 (kk & 1))
__device__ __forceinline__ void mma_chunk2(const Bars5& B, uint32_t smem_base, Ring& wr, bool ts, uint32_t a_addr, uint32_t a_tmem,
                                           uint32_t d_tmem, bool first, bool split, uint32_t a_empty_bar, int ksteps = 4) {
  // ksteps < 4: only the first 16 * ksteps of the chunk's 64 K-columns are non-zero (the xyz chunk of the occupancy MLP: 3 columns)
  const uint32_t idesc = umma_idesc_f16(128, 256);
  const uint64_t a_hi = umma_desc_sw128(a_addr), a_lo = umma_desc_sw128(a_addr + CT_A_HALF);
  mbar_wait(B.wfull(wr.idx), wr.phase);
  tc_fence_after();
  {
    const uint64_t w = umma_desc_sw128(smem_base + wr.idx * CT_TILE_BYTES);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k >= ksteps) break;
      const uint32_t acc = (!first || k > 0) ? 1u : 0u;
      if (ts) {
        const uint32_t at = a_tmem + 32u * (k >> 1) + 8u * (k & 1);
        umma_ts(d_tmem, at, w + 2 * k, idesc, acc);
        if (split) umma_ts(d_tmem, at + 16u, w + 2 * k, idesc, 1u);
      } else {
        umma_bf16(d_tmem, a_hi + 2 * k, w + 2 * k, idesc, acc);
        if (split) umma_bf16(d_tmem, a_lo + 2 * k, w + 2 * k, idesc, 1u);
      }
    }
    umma_commit(B.wempty(wr.idx));
    wr.advance();
  }
  if (split) {
    mbar_wait(B.wfull(wr.idx), wr.phase);
    tc_fence_after();
    const uint64_t w = umma_desc_sw128(smem_base + wr.idx * CT_TILE_BYTES);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k >= ksteps) break;
      if (ts) umma_ts(d_tmem, a_tmem + 32u * (k >> 1) + 8u * (k & 1), w + 2 * k, idesc, 1u);
      else umma_bf16(d_tmem, a_hi + 2 * k, w + 2 * k, idesc, 1u);
    }
    umma_commit(B.wempty(wr.idx));
    wr.advance();
  }
  if (a_empty_bar != 0u) umma_commit(a_empty_bar);
}

__device__ __forceinline__ void w_stream2(const Bars5& B, uint32_t smem_base, Ring& wr, const uint8_t* blob, int pairs, bool split) {
  for (int i = 0; i < pairs; ++i) {
    const uint8_t* src = blob + (size_t)i * 2 * CT_TILE_BYTES;
    for (int h = 0; h < (split ? 2 : 1); ++h) {
      mbar_wait(B.wempty(wr.idx), wr.phase ^ 1);
      mbar_arrive_expect_tx(B.wfull(wr.idx), CT_TILE_BYTES);
      bulk_g2s(smem_base + wr.idx * CT_TILE_BYTES, src + (size_t)h * CT_TILE_BYTES, CT_TILE_BYTES, B.wfull(wr.idx));
      wr.advance();
    }
  }
}

// epilogue thread: 32 accumulator columns -> act(d + bias) -> split fp16 -> back into the same 32 TMEM columns [16 hi | 16 lo]
template <int ACT>
__device__ __forceinline__ void epi_to_tmem(uint32_t taddr, const uint32_t (&rr)[32], const float* __restrict__ bias, bool split) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c * 8 + 4));
    const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = add2(make_float2(__uint_as_float(rr[c * 8 + 2 * j]), __uint_as_float(rr[c * 8 + 2 * j + 1])), bb[j]);
      const float2 v = ACT == ZS_ACT_GELU ? fast_gelu_erf2(t) : fast_softplus100_2(t);
      split_f16x2(v.x, v.y, hi[4 * c + j], lo[4 * c + j]);
    }
  }
  tmem_st_32x16(taddr, hi);
  if (split) tmem_st_32x16(taddr + 16u, lo);
}

// x <- x + fc2(GELU(fc1(LN(x))))      blob order as chain_mlp_kernel
__global__ void __launch_bounds__(CT_THREADS, 1) chain_mlp2_kernel(ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  if (smem_base - smem_u32(smem_raw) > CT_SMEM - CT_SMEM_USED) __trap();
  Bars5 B{smem_base + CT_OFF_BAR};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = p.precision == 0;
  const uint32_t tmem_base = chain2_setup(B, smem_gen, smem_base, warp, M2_WSLOTS);
  const int n_tiles = (p.M + 127) / 128;

  if (warp < 4) {
    // ---------------- loader: LN(x) chunks, 4 groups x 4 chunks per tile ----------------
    Ring lr(CT_LSLOTS);
    float4 buf[16];
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = t * 128 + warp * 32;
      float sc, sh;
      if (t + (int)gridDim.x < n_tiles) prefetch_rows_l2(p.x, p.ldx, m0 + (int)gridDim.x * 128, p.M, 0, lane);
      warp_ln_stats(p.x, p.ldx, m0, p.M, p.ln_eps, lane, sc, sh);
      fetch_chunk_co(p.x, p.ldx, m0, p.M, 0, lane, buf);
      for (int i = 0; i < 16; ++i) {
        const int kc = i & 3;
        mbar_wait(B.lempty(lr.idx), lr.phase ^ 1);
        store_chunk_co(smem_gen + M2_OFF_L + lr.idx * 2 * CT_A_HALF, warp, lane, buf, true, sc, sh,
                       p.ln_w ? p.ln_w + kc * 64 : nullptr, p.ln_b ? p.ln_b + kc * 64 : nullptr, split);
        fence_proxy_async_smem();
        mbar_arrive(B.lfull(lr.idx));
        lr.advance();
        if (i + 1 < 16) fetch_chunk_co(p.x, p.ldx, m0, p.M, ((i + 1) & 3) * 64, lane, buf);
      }
    }
  } else if (warp == 13) {
    if (lane == 0) {
      Ring wr(M2_WSLOTS);
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) w_stream2(B, smem_base, wr, p.blob, 32, split);
    }
  } else if (warp == 12) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      Ring wr(M2_WSLOTS), lr(CT_LSLOTS);
      uint32_t te_phase[2] = {0, 0}, ef_phase = 0;
      const uint32_t d0 = tmem_base, d1 = tmem_base + 256;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int g = 0; g < 4; ++g) {
          mbar_wait(B.tempty(0), te_phase[0] ^ 1); te_phase[0] ^= 1;
          tc_fence_after();
          for (int kc = 0; kc < 4; ++kc) {
            mbar_wait(B.lfull(lr.idx), lr.phase);
            tc_fence_after();
            mma_chunk2(B, smem_base, wr, false, smem_base + M2_OFF_L + lr.idx * 2 * CT_A_HALF, 0u, d0, kc == 0, split, B.lempty(lr.idx));
            lr.advance();
          }
          umma_commit(B.tfull(0));
          if (g == 0) { mbar_wait(B.tempty(1), te_phase[1] ^ 1); te_phase[1] ^= 1; tc_fence_after(); }
          for (int kc = 0; kc < 4; ++kc) {      // fc2: A = GELU(fc1) chunk kc, in place in d0's columns [64 kc, 64 kc + 64)
            mbar_wait(B.efull(kc), ef_phase);
            tc_fence_after();
            mma_chunk2(B, smem_base, wr, true, 0u, d0 + 64u * kc, d1, g == 0 && kc == 0, split, 0u);
          }
          ef_phase ^= 1;
        }
        umma_commit(B.tfull(1));
      }
    }
  } else {
    // ---------------- epilogue: 8 warps; lane quarter q, column half hsel ----------------
    const int e = warp - 4, q = e & 3, hsel = e >> 2;
    uint32_t tf_phase[2] = {0, 0};
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint8_t* wscr = smem_gen + M2_OFF_T + e * 4096;       // this warp's [32 rows][32 cols] fp32 transpose scratch
    const int sub = lane >> 3, q8 = lane & 7;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int g = 0; g < 4; ++g) {
        mbar_wait(B.tfull(0), tf_phase[0]); tf_phase[0] ^= 1;
        tc_fence_after();
        for (int c = 0; c < 4; ++c) {
          uint32_t rr[32];
          const uint32_t ta = tmem_base + lane_off + c * 64 + hsel * 32;
          tmem_ld_32x32(ta, rr);
          tmem_ld_wait();
          epi_to_tmem<ZS_ACT_GELU>(ta, rr, p.bias + g * 256 + c * 64 + hsel * 32, split);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(B.efull(c));
        }
        tc_fence_before();
        mbar_arrive(B.tempty(0));
      }
      mbar_wait(B.tfull(1), tf_phase[1]); tf_phase[1] ^= 1;
      tc_fence_after();
      // x += fc2 + b2: every 32 x 32 accumulator block is transposed inside its warp so that global memory sees whole
      // 128-byte row segments (as chain_lin_kernel)
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col0 = c * 64 + hsel * 32;
        uint32_t rr[32];
        tmem_ld_32x32(tmem_base + 256 + lane_off + col0, rr);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(wscr + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
        __syncwarp();
        const int n0 = col0 + 4 * q8;
        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias2 + n0));
        float4 xin[8];                       // all residual loads first: they alias the stores below
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int mm = t * 128 + q * 32 + 4 * i + sub;
          xin[i] = *reinterpret_cast<const float4*>(p.x + (int64_t)(mm < p.M ? mm : 0) * p.ldx + n0);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = 4 * i + sub;
          const int mm = t * 128 + q * 32 + rl;
          const float4 a = *reinterpret_cast<const float4*>(wscr + rl * 128 + ((q8 ^ (rl & 7)) << 4));
          if (mm < p.M)
            *reinterpret_cast<float4*>(p.x + (int64_t)mm * p.ldx + n0) =
                make_float4(xin[i].x + a.x + bv.x, xin[i].y + a.y + bv.y, xin[i].z + a.z + bv.z, xin[i].w + a.w + bv.w);
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(B.tempty(1));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ===============================================================================================================
// Attention output projection + residual AND the MLP of an ImplFuncBlock in one kernel (model/shape/implicit.py:74,105-108):
//     x' = x + A Wp^T + bp          (A = the tile-blocked attention output of chain_qkvattn2_kernel<true>)
//     x  = x' + fc2(GELU(fc1(LayerNorm(x'))))
// chain_mlp2_kernel with one more GEMM in front.  Per tile: the loaders stream the four 64-column chunks of A (512 contiguous
// bytes per warp request in the blocked layout) through ring L, the proj MMAs accumulate into the FIRST accumulator half
// (free since fc2 of the previous tile's last group has read it -- so they overlap that tile's final epilogue, which reads the
// second half), and the epilogue warps finish x' = acc + bp + x in the transposed (row-segment) layout, write it back in place
// and reduce the LayerNorm sums of every row on the way (16 values per thread, 8 lanes by shuffles, the two column halves
// through shared memory).  The loaders then read x' back (L2 hits, ld.global.cg: it was written by other warps of this CTA)
// exactly as chain_mlp2_kernel reads x.  HBM sees A and x once and x once more for the result: the separate proj launch
// (4.3 KB per point, the one HBM-bound launch of the decoder) is gone.
struct BarsP : Bars5 {
  __device__ uint32_t xready() const { return base + 184u; }
};

// 64 columns of a TILE-BLOCKED matrix for a 32-row loader warp: instruction j = 16-byte chunk kc * 16 + j of rows 32 w + lane
__device__ __forceinline__ void fetch_chunk32_blk(const float* __restrict__ a, int t, int w, int kc, int lane, float4* buf) {
  const float4* base = reinterpret_cast<const float4*>(a) + ((size_t)t * 64 + kc * 16) * 128 + 32 * w + lane;
#pragma unroll
  for (int j = 0; j < 16; ++j) buf[j] = __ldg(base + (size_t)j * 128);
}
__device__ __forceinline__ void store_chunk32_blk(uint8_t* slot, int w, int lane, const float4* buf, bool split) {
  const int r = 32 * w + lane;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    uint2 hi, lo;
    split_f16x2(buf[j].x, buf[j].y, hi.x, lo.x);
    split_f16x2(buf[j].z, buf[j].w, hi.y, lo.y);
    const uint32_t off = swizzle128_offset(r, j >> 1) + ((j & 1) << 3);
    *reinterpret_cast<uint2*>(slot + off) = hi;
    if (split) *reinterpret_cast<uint2*>(slot + CT_A_HALF + off) = lo;
  }
}
// fetch_chunk_co with L2-coherent loads (the rows were written by other warps of this CTA a moment ago)
__device__ __forceinline__ void fetch_chunk_co_cg(const float* x, int ldx, int m0, int M, int col0, int lane, float4* buf) {
  const int sub = lane >> 4, q = lane & 15;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int m = m0 + 2 * j + sub;
    buf[j] = m < M ? __ldcg(reinterpret_cast<const float4*>(x + (int64_t)m * ldx + col0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(CT_THREADS, 1) chain_pmlp_kernel(ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  if (smem_base - smem_u32(smem_raw) > CT_SMEM - CT_SMEM_USED) __trap();
  BarsP B{{smem_base + CT_OFF_BAR}};
  float2* stats = reinterpret_cast<float2*>(smem_gen + CT_OFF_BAR + 256);     // [2 column halves][128 rows] (sum, sum of squares)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = p.precision == 0;
  if (threadIdx.x == 0) mbar_init(B.xready(), 256);
  const uint32_t tmem_base = chain2_setup(B, smem_gen, smem_base, warp, M2_WSLOTS);
  const int n_tiles = (p.M + 127) / 128;

  if (warp < 4) {
    // ---------------- loader: 4 chunks of A, then (after x' is written) 4 groups x 4 chunks of LayerNorm(x') ----------------
    Ring lr(CT_LSLOTS);
    float4 buf[16];
    uint32_t ph_x = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = t * 128 + warp * 32;
      if (t + (int)gridDim.x < n_tiles) {
        const char* nb = reinterpret_cast<const char*>(p.a_blk) + (size_t)(t + gridDim.x) * 131072 + warp * 32768 + lane * 128;
        // OFF by default: measured with ncu on the 129^3 pass, a whole-tile-ahead L2 prefetch of A and x nearly DOUBLES the DRAM
        // reads of this kernel (8.19 GB vs the algorithmic 4.40 GB: the lines are evicted again before their demand loads, the
        // working set of 148 CTAs x [A, x, x' re-read four times, next A, next x] exceeds what L2 keeps) and is 2 % slower
        if (p.flags & 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + i * 4096));
        }
        if (p.flags & 4) prefetch_rows_l2(p.x, p.ldx, m0 + (int)gridDim.x * 128, p.M, 0, lane);
      }
      fetch_chunk32_blk(p.a_blk, t, warp, 0, lane, buf);
      for (int kc = 0; kc < 4; ++kc) {
        mbar_wait(B.lempty(lr.idx), lr.phase ^ 1);
        store_chunk32_blk(smem_gen + M2_OFF_L + lr.idx * 2 * CT_A_HALF, warp, lane, buf, split);
        fence_proxy_async_smem();
        mbar_arrive(B.lfull(lr.idx));
        lr.advance();
        if (kc + 1 < 4) fetch_chunk32_blk(p.a_blk, t, warp, kc + 1, lane, buf);
      }
      mbar_wait(B.xready(), ph_x); ph_x ^= 1;
      fetch_chunk_co_cg(p.x, p.ldx, m0, p.M, 0, lane, buf);
      float sc = 0.f, sh = 0.f;
      if (m0 + lane < p.M) {
        const float2 s0 = stats[32 * warp + lane], s1 = stats[128 + 32 * warp + lane];
        const float mean = (s0.x + s1.x) * (1.0f / 256.0f);
        const float var = fmaxf((s0.y + s1.y) * (1.0f / 256.0f) - mean * mean, 0.f);
        sc = rsqrtf(var + p.ln_eps);
        sh = -mean * sc;
      }
      for (int i = 0; i < 16; ++i) {
        const int kc = i & 3;
        mbar_wait(B.lempty(lr.idx), lr.phase ^ 1);
        store_chunk_co(smem_gen + M2_OFF_L + lr.idx * 2 * CT_A_HALF, warp, lane, buf, true, sc, sh,
                       p.ln_w ? p.ln_w + kc * 64 : nullptr, p.ln_b ? p.ln_b + kc * 64 : nullptr, split);
        fence_proxy_async_smem();
        mbar_arrive(B.lfull(lr.idx));
        lr.advance();
        if (i + 1 < 16) fetch_chunk_co_cg(p.x, p.ldx, m0, p.M, ((i + 1) & 3) * 64, lane, buf);
      }
    }
  } else if (warp == 13) {
    if (lane == 0) {
      Ring wr(M2_WSLOTS);
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        w_stream2(B, smem_base, wr, p.pblob, 4, split);
        w_stream2(B, smem_base, wr, p.blob, 32, split);
      }
    }
  } else if (warp == 12) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      Ring wr(M2_WSLOTS), lr(CT_LSLOTS);
      uint32_t te_phase[2] = {0, 0}, ef_phase = 0;
      const uint32_t d0 = tmem_base, d1 = tmem_base + 256;
      int tn = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int g = -1; g < 4; ++g) {            // g = -1: the projection, into the first half like an fc1 group
          trace_ev(p.trace, 0, tn, 7);
          mbar_wait(B.tempty(0), te_phase[0] ^ 1); te_phase[0] ^= 1;
          tc_fence_after();
          for (int kc = 0; kc < 4; ++kc) {
            trace_ev(p.trace, 0, tn, 1);
            mbar_wait(B.lfull(lr.idx), lr.phase);
            trace_ev(p.trace, 0, tn, 2);
            tc_fence_after();
            mma_chunk2(B, smem_base, wr, false, smem_base + M2_OFF_L + lr.idx * 2 * CT_A_HALF, 0u, d0, kc == 0, split, B.lempty(lr.idx));
            trace_ev(p.trace, 0, tn, 6);
            lr.advance();
          }
          umma_commit(B.tfull(0));
          if (g < 0) continue;
          if (g == 0) { mbar_wait(B.tempty(1), te_phase[1] ^ 1); te_phase[1] ^= 1; tc_fence_after(); }
          for (int kc = 0; kc < 4; ++kc) {
            trace_ev(p.trace, 0, tn, 8);
            mbar_wait(B.efull(kc), ef_phase);
            trace_ev(p.trace, 0, tn, 9);
            tc_fence_after();
            mma_chunk2(B, smem_base, wr, true, 0u, d0 + 64u * kc, d1, g == 0 && kc == 0, split, 0u);
            trace_ev(p.trace, 0, tn, 4);
          }
          ef_phase ^= 1;
        }
        umma_commit(B.tfull(1));
      }
    }
  } else {
    // ---------------- epilogue: 8 warps; lane quarter q, column half hsel ----------------
    const int e = warp - 4, q = e & 3, hsel = e >> 2;
    uint32_t tf_phase[2] = {0, 0};
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint8_t* wscr = smem_gen + M2_OFF_T + e * 4096;
    const int sub = lane >> 3, q8 = lane & 7;
    // P phase of tile t: x' = x + proj + bp written in place, LayerNorm sums published.  It runs one tile AHEAD of the final epilogue
    // (x'' of the previous tile): fc1 of tile t+1 waits for it, nothing waits for the final epilogue but fc2 of tile t+1.
    auto p_phase = [&](int t) {
      // ---- x' = x + proj + bp, written in place; LayerNorm sums of x' ----
      mbar_wait(B.tfull(0), tf_phase[0]); tf_phase[0] ^= 1;
      tc_fence_after();
      float s1[8], s2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col0 = c * 64 + hsel * 32;
        uint32_t rr[32];
        tmem_ld_32x32(tmem_base + lane_off + col0, rr);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(wscr + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
        __syncwarp();
        const int n0 = col0 + 4 * q8;
        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias_p + n0));
        float4 xin[8];
        if (p.pp != nullptr) {              // points mode: the residual is LinearProj3D(points), recomputed (see ChainParams)
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.pp + n0)), w1 = __ldg(reinterpret_cast<const float4*>(p.pp + 256 + n0));
          const float4 w2 = __ldg(reinterpret_cast<const float4*>(p.pp + 512 + n0)), bb = __ldg(reinterpret_cast<const float4*>(p.pp + 768 + n0));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int mm = t * 128 + q * 32 + 4 * i + sub;
            const float* pt = p.points + (int64_t)(mm < p.M ? mm : 0) * 3;
            const float x = __ldg(pt), y = __ldg(pt + 1), z = __ldg(pt + 2);
            xin[i] = make_float4(fmaf(w2.x, z, fmaf(w1.x, y, w0.x * x)) + bb.x, fmaf(w2.y, z, fmaf(w1.y, y, w0.y * x)) + bb.y,
                                 fmaf(w2.z, z, fmaf(w1.z, y, w0.z * x)) + bb.z, fmaf(w2.w, z, fmaf(w1.w, y, w0.w * x)) + bb.w);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int mm = t * 128 + q * 32 + 4 * i + sub;
            xin[i] = *reinterpret_cast<const float4*>(p.x + (int64_t)(mm < p.M ? mm : 0) * p.ldx + n0);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = 4 * i + sub;
          const int mm = t * 128 + q * 32 + rl;
          const float4 a = *reinterpret_cast<const float4*>(wscr + rl * 128 + ((q8 ^ (rl & 7)) << 4));
          const float4 v = make_float4(xin[i].x + a.x + bv.x, xin[i].y + a.y + bv.y, xin[i].z + a.z + bv.z, xin[i].w + a.w + bv.w);
          if (mm < p.M) *reinterpret_cast<float4*>(p.x + (int64_t)mm * p.ldx + n0) = v;
          s1[i] += (v.x + v.y) + (v.z + v.w);
          s2[i] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(B.tempty(0));
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], o);
          s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], o);
        }
      }
      if (q8 == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) stats[hsel * 128 + q * 32 + 4 * i + sub] = make_float2(s1[i], s2[i]);
      }
      __threadfence_block();
      mbar_arrive(B.xready());
    };
    if ((int)blockIdx.x < n_tiles) p_phase(blockIdx.x);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      // ---- the MLP, as chain_mlp2_kernel ----
      for (int g = 0; g < 4; ++g) {
        mbar_wait(B.tfull(0), tf_phase[0]); tf_phase[0] ^= 1;
        tc_fence_after();
        for (int c = 0; c < 4; ++c) {
          uint32_t rr[32];
          const uint32_t ta = tmem_base + lane_off + c * 64 + hsel * 32;
          tmem_ld_32x32(ta, rr);
          tmem_ld_wait();
          epi_to_tmem<ZS_ACT_GELU>(ta, rr, p.bias + g * 256 + c * 64 + hsel * 32, split);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(B.efull(c));
        }
        tc_fence_before();
        mbar_arrive(B.tempty(0));
      }
      if (t + (int)gridDim.x < n_tiles) p_phase(t + (int)gridDim.x);
      mbar_wait(B.tfull(1), tf_phase[1]); tf_phase[1] ^= 1;
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col0 = c * 64 + hsel * 32;
        uint32_t rr[32];
        tmem_ld_32x32(tmem_base + 256 + lane_off + col0, rr);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(wscr + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
        __syncwarp();
        const int n0 = col0 + 4 * q8;
        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias2 + n0));
        float4 xin[8];                       // x' of this thread's own earlier stores (same row / column mapping as above)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int mm = t * 128 + q * 32 + 4 * i + sub;
          xin[i] = *reinterpret_cast<const float4*>(p.x + (int64_t)(mm < p.M ? mm : 0) * p.ldx + n0);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = 4 * i + sub;
          const int mm = t * 128 + q * 32 + rl;
          const float4 a = *reinterpret_cast<const float4*>(wscr + rl * 128 + ((q8 ^ (rl & 7)) << 4));
          if (mm < p.M)
            *reinterpret_cast<float4*>(p.x + (int64_t)mm * p.ldx + n0) =
                make_float4(xin[i].x + a.x + bv.x, xin[i].y + a.y + bv.y, xin[i].z + a.z + bv.z, xin[i].w + a.w + bv.w);
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(B.tempty(1));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// logit = MLPBlocks([xyz, LN(x)]); chunk / blob order as chain_occ_kernel
__global__ void __launch_bounds__(CT_THREADS, 1) chain_occ2_kernel(ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  if (smem_base - smem_u32(smem_raw) > CT_SMEM - CT_SMEM_USED) __trap();
  Bars5 B{smem_base + CT_OFF_BAR};
  float* row_scratch = reinterpret_cast<float*>(smem_gen + CT_OFF_BAR + 256);   // [128] partial dots
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = p.precision == 0;
  const uint32_t tmem_base = chain2_setup(B, smem_gen, smem_base, warp, O2_WSLOTS);
  const int n_tiles = (p.M + 127) / 128;

  if (warp < 4) {
    Ring lr(CT_LSLOTS);
    float4 buf[16];
    const int sub = lane >> 4, q = lane & 15;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = t * 128 + warp * 32;
      float sc, sh;
      if (t + (int)gridDim.x < n_tiles) prefetch_rows_l2(p.x, p.ldx, m0 + (int)gridDim.x * 128, p.M, 0, lane);
      warp_ln_stats(p.x, p.ldx, m0, p.M, p.ln_eps, lane, sc, sh);
      for (int rep = 0; rep < 4; ++rep) {
        for (int kc = 0; kc < 5; ++kc) {
          if (kc < 4) {
            fetch_chunk_co(p.x, p.ldx, m0, p.M, kc * 64, lane, buf);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int m = m0 + 2 * j + sub;
              buf[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (q == 0 && m < p.M) {
                const float* pp = p.points + (int64_t)m * 3;
                buf[j] = make_float4(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2), 0.f);
              }
            }
          }
          mbar_wait(B.lempty(lr.idx), lr.phase ^ 1);
          store_chunk_co(smem_gen + O2_OFF_L + lr.idx * 2 * CT_A_HALF, warp, lane, buf, kc < 4, sc, sh,
                         (kc < 4 && p.ln_w) ? p.ln_w + kc * 64 : nullptr, (kc < 4 && p.ln_b) ? p.ln_b + kc * 64 : nullptr, split);
          fence_proxy_async_smem();
          mbar_arrive(B.lfull(lr.idx));
          lr.advance();
        }
      }
    }
  } else if (warp == 13) {
    if (lane == 0) {
      Ring wr(O2_WSLOTS);
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) w_stream2(B, smem_base, wr, p.blob, 48, split);
    }
  } else if (warp == 12) {
    if (lane == 0) {
      Ring wr(O2_WSLOTS), lr(CT_LSLOTS);
      uint32_t te_phase[2] = {0, 0}, ef_phase = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int l = 0; l < 8; ++l) {
          const int half = l & 1;
          const uint32_t d = tmem_base + half * 256, dprev = tmem_base + (half ^ 1) * 256;
          mbar_wait(B.tempty(half), te_phase[half] ^ 1); te_phase[half] ^= 1;
          tc_fence_after();
          const int nL = (l & 1) ? 0 : 5, nE = l == 0 ? 0 : 4;
          for (int i = 0; i < nL; ++i) {
            mbar_wait(B.lfull(lr.idx), lr.phase);
            tc_fence_after();
            mma_chunk2(B, smem_base, wr, false, smem_base + O2_OFF_L + lr.idx * 2 * CT_A_HALF, 0u, d, i == 0, split, B.lempty(lr.idx),
                       i == 4 ? 1 : 4);      // chunk 4 = [xyz | 61 zero columns]: one K-step
            lr.advance();
          }
          for (int i = 0; i < nE; ++i) {          // A = the previous layer's activations, in place in the other half
            mbar_wait(B.efull(i), ef_phase);
            tc_fence_after();
            mma_chunk2(B, smem_base, wr, true, 0u, dprev + 64u * i, d, nL == 0 && i == 0, split, 0u);
          }
          if (nE) ef_phase ^= 1;
          umma_commit(B.tfull(half));
        }
      }
    }
  } else {
    const int e = warp - 4, q = e & 3, hsel = e >> 2;
    const int row = q * 32 + lane;
    uint32_t tf_phase[2] = {0, 0};
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m = t * 128 + row;
      for (int l = 0; l < 8; ++l) {
        const int half = l & 1;
        mbar_wait(B.tfull(half), tf_phase[half]); tf_phase[half] ^= 1;
        tc_fence_after();
        float dot = 0.f;
        for (int c = 0; c < 4; ++c) {
          uint32_t rr[32];
          const uint32_t ta = tmem_base + half * 256 + lane_off + c * 64 + hsel * 32;
          tmem_ld_32x32(ta, rr);
          tmem_ld_wait();
          const float* bl = p.bias + l * 256 + c * 64 + hsel * 32;
          if (l < 7) {
            epi_to_tmem<ZS_ACT_SOFTPLUS100>(ta, rr, bl, split);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(B.efull(c));
          } else {
            const float* w8 = p.bias2 + c * 64 + hsel * 32;
            float2 dot2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bl) + j), w4 = __ldg(reinterpret_cast<const float4*>(w8) + j);
              const float2 s0 = fast_softplus100_2(add2(make_float2(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1])), make_float2(b4.x, b4.y)));
              const float2 s1 = fast_softplus100_2(add2(make_float2(__uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3])), make_float2(b4.z, b4.w)));
              dot2 = fma2(s0, make_float2(w4.x, w4.y), dot2);
              dot2 = fma2(s1, make_float2(w4.z, w4.w), dot2);
            }
            dot += dot2.x + dot2.y;
          }
        }
        if (l == 7) {
          // nothing reads this half any more: release it; the activations of layers < 7 stay until the next layer's MMAs
          // have read them, which the in-order tensor pipe and the next tempty wait of that half guarantee
          tc_fence_before();
          mbar_arrive(B.tempty(half));
          if (hsel == 1) row_scratch[row] = dot;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (hsel == 0 && m < p.M) {
            float v = dot + row_scratch[row] + p.b8;
            p.out[m] = p.apply_sigmoid ? 1.0f / (1.0f + __expf(-v)) : v;
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        } else {
          tc_fence_before();
          mbar_arrive(B.tempty(half));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static unsigned long long* g_chain_trace = nullptr;   // debug only (zs_debug_chain_trace)
static int g_chain_dbg = 0;       // zs_debug_chain_variant bits 1.. : experiment switches (chain_pmlp_kernel: 2 = L2-prefetch the next A tile, 4 = the next x tile; 8 = chain_qkvattn2_kernel without its x prefetch)
static int g_chain_variant = 1;   // zs_chain_mlp_fwd / zs_chain_occ_fwd: 1 = activations in tensor memory (chain_*2_kernel), 0 = ring E

static int chain_launch(void (*kern)(ChainParams), ChainParams p, cudaStream_t st, const char* name, int threads = CT_THREADS) {
  p.trace = g_chain_trace;
  ZS_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
  int tiles = (p.M + 127) / 128;
  int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, threads, CT_SMEM, st>>>(p);
  ZS_CUDA_CHECK_LAUNCH(name);
  return ZS_OK;
}

}  // namespace zs

using namespace zs;

extern "C" size_t zs_chain_mlp_blob_bytes(void) { return (size_t)32 * 2 * CT_TILE_BYTES; }
extern "C" size_t zs_chain_occ_blob_bytes(void) { return (size_t)48 * 2 * CT_TILE_BYTES; }

extern "C" int zs_chain_mlp_fwd(float* x, int ldx, int M, const float* ln_w, const float* ln_b, float ln_eps,
                                const void* blob, const float* b1, const float* b2, int precision, void* stream) {
  ZS_REQUIRE(x && blob && b1 && b2 && M >= 0 && ((ln_w == nullptr) == (ln_b == nullptr)), "zs_chain_mlp_fwd: null pointer");
  ZS_REQUIRE(ldx >= 256 && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "zs_chain_mlp_fwd: x must be 16B aligned, ldx%4==0");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(blob) & 15) == 0 && (precision == 0 || precision == 1), "zs_chain_mlp_fwd: bad blob/precision");
  if (M == 0) return ZS_OK;
  ChainParams p{};
  p.x = x; p.ldx = ldx; p.M = M; p.ln_w = ln_w; p.ln_b = ln_b; p.ln_eps = ln_eps;
  p.blob = reinterpret_cast<const uint8_t*>(blob); p.bias = b1; p.bias2 = b2; p.precision = precision;
  return chain_launch(g_chain_variant ? chain_mlp2_kernel : chain_mlp_kernel, p, as_stream(stream), "zs_chain_mlp_fwd");
}

extern "C" int zs_chain_pmlp_fwd(float* x, int ldx, int M, const float* a_blk, const void* proj_blob, const float* proj_bias,
                                 float ln_eps, const void* mlp_blob, const float* b1, const float* b2, const float* points,
                                 const float* pp, int precision, void* stream) {
  ZS_REQUIRE((points == nullptr) == (pp == nullptr) && (reinterpret_cast<uintptr_t>(pp) & 15) == 0,
             "zs_chain_pmlp_fwd: points and pp go together; pp must be 16-byte aligned");
  ZS_REQUIRE(x && a_blk && proj_blob && proj_bias && mlp_blob && b1 && b2 && M >= 0, "zs_chain_pmlp_fwd: null pointer");
  ZS_REQUIRE(ldx >= 256 && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "zs_chain_pmlp_fwd: x must be 16B aligned, ldx%%4==0");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(a_blk) & 15) == 0 && (reinterpret_cast<uintptr_t>(proj_blob) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(mlp_blob) & 15) == 0 && (reinterpret_cast<uintptr_t>(proj_bias) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(b1) & 15) == 0 && (reinterpret_cast<uintptr_t>(b2) & 15) == 0,
             "zs_chain_pmlp_fwd: blobs / biases / a_blk must be 16-byte aligned");
  ZS_REQUIRE(precision == 0 || precision == 1, "zs_chain_pmlp_fwd: bad precision");
  if (M == 0) return ZS_OK;
  ChainParams p{};
  p.x = x; p.ldx = ldx; p.M = M; p.ln_eps = ln_eps; p.a_blk = a_blk; p.pblob = reinterpret_cast<const uint8_t*>(proj_blob);
  p.bias_p = proj_bias; p.blob = reinterpret_cast<const uint8_t*>(mlp_blob); p.bias = b1; p.bias2 = b2; p.precision = precision;
  p.flags = g_chain_dbg; p.points = points; p.pp = pp;
  return chain_launch(chain_pmlp_kernel, p, as_stream(stream), "zs_chain_pmlp_fwd");
}

extern "C" int zs_chain_lin_fwd(const float* x, int ldx, int M, int do_ln, float ln_eps, const void* blob, int n_tiles,
                                const float* bias, const float* res, int ldres, float* out, int ldo, int precision, void* stream) {
  ZS_REQUIRE(x && blob && out && M >= 0, "zs_chain_lin_fwd: null pointer");
  ZS_REQUIRE(n_tiles >= 1 && n_tiles <= 16, "zs_chain_lin_fwd: n_tiles must be in [1, 16]");
  ZS_REQUIRE(do_ln >= 0 && do_ln <= 2 && (do_ln != 2 || n_tiles == 1), "zs_chain_lin_fwd: do_ln must be 0, 1 or 2 (tile-blocked x, one n-tile)");
  ZS_REQUIRE(ldx >= 256 && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "zs_chain_lin_fwd: x must be 16B aligned, ldx%%4==0");
  ZS_REQUIRE(ldo >= 256 * n_tiles && (ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "zs_chain_lin_fwd: out must be 16B aligned, ldo%%4==0");
  ZS_REQUIRE(res == nullptr || ((ldres & 3) == 0 && ldres >= 256 * n_tiles && (reinterpret_cast<uintptr_t>(res) & 15) == 0),
             "zs_chain_lin_fwd: res must be 16B aligned, ldres%%4==0");
  ZS_REQUIRE(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "zs_chain_lin_fwd: bias must be 16B aligned");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(blob) & 15) == 0 && (precision == 0 || precision == 1), "zs_chain_lin_fwd: bad blob/precision");
  if (M == 0) return ZS_OK;
  ChainParams p{};
  p.x = const_cast<float*>(x); p.ldx = ldx; p.M = M; p.ln_eps = ln_eps; p.do_ln = do_ln;
  p.blob = reinterpret_cast<const uint8_t*>(blob); p.bias = bias; p.res = res; p.ldres = ldres; p.out = out; p.ldo = ldo;
  p.n_tiles = n_tiles; p.precision = precision;
  return chain_launch(chain_lin_kernel, p, as_stream(stream), "zs_chain_lin_fwd", LN_THREADS);
}

// debug: while `buf` (device, [3][512] uint64) is non-null every chained kernel records (clock64 << 8 | tag) events of the
// MMA thread, loader thread 0 and epilogue warp 4 lane 0 of CTA 0 into it (tools/trace_chain.py).  Not thread-safe.
extern "C" int zs_debug_chain_trace(unsigned long long* buf) { g_chain_trace = buf; return ZS_OK; }
// A/B switch of the two MLP kernels (tools/diag_decoder.py, tests): 1 (default) = TMEM-resident activations, 0 = smem ring E
extern "C" int zs_debug_chain_variant(int v) { g_chain_variant = (v & 1) != 0; g_chain_dbg = v & ~1; return ZS_OK; }

extern "C" int zs_chain_attn_fwd(const float* qkv, int ld_qkv, int M, const void* Kblob, const void* Vblob, int n_keys,
                                 float scale, float* O, int precision, void* stream) {
  ZS_REQUIRE(qkv && Kblob && Vblob && O && M >= 0, "zs_chain_attn_fwd: null pointer");
  ZS_REQUIRE(n_keys > 0 && n_keys <= 208, "zs_chain_attn_fwd: n_keys must be in [1, 208]");
  ZS_REQUIRE(ld_qkv >= 768 && (ld_qkv & 3) == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(O) & 15) == 0, "zs_chain_attn_fwd: qkv/O must be 16B aligned, ld % 4 == 0");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(Kblob) & 15) == 0 && (reinterpret_cast<uintptr_t>(Vblob) & 15) == 0 &&
             (precision == 0 || precision == 1), "zs_chain_attn_fwd: bad blob / precision");
  if (M == 0) return ZS_OK;
  ChainParams p{};
  p.x = const_cast<float*>(qkv); p.ldx = ld_qkv; p.M = M; p.blob = reinterpret_cast<const uint8_t*>(Kblob);
  p.points = reinterpret_cast<const float*>(Vblob); p.out = O; p.b8 = scale; p.apply_sigmoid = n_keys; p.precision = precision;
  return chain_launch(chain_attn_kernel, p, as_stream(stream), "zs_chain_attn_fwd");
}

extern "C" int zs_chain_occ_fwd(const float* x, int ldx, const float* points, int M, const float* ln_w, const float* ln_b,
                                float ln_eps, const void* blob, const float* biases, const float* w8, float b8,
                                float* out, int apply_sigmoid, int precision, void* stream) {
  ZS_REQUIRE(x && points && blob && biases && w8 && out && M >= 0 && ((ln_w == nullptr) == (ln_b == nullptr)),
             "zs_chain_occ_fwd: null pointer");
  ZS_REQUIRE(ldx >= 256 && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "zs_chain_occ_fwd: x must be 16B aligned, ldx%4==0");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(blob) & 15) == 0 && (precision == 0 || precision == 1), "zs_chain_occ_fwd: bad blob/precision");
  if (M == 0) return ZS_OK;
  ChainParams p{};
  p.x = const_cast<float*>(x); p.ldx = ldx; p.points = points; p.M = M; p.ln_w = ln_w; p.ln_b = ln_b; p.ln_eps = ln_eps;
  p.blob = reinterpret_cast<const uint8_t*>(blob); p.bias = biases; p.bias2 = w8; p.b8 = b8; p.out = out;
  p.apply_sigmoid = apply_sigmoid; p.precision = precision; p.flags = g_chain_dbg;
  return chain_launch(g_chain_variant ? chain_occ2_kernel : chain_occ_kernel, p, as_stream(stream), "zs_chain_occ_fwd");
}

extern "C" size_t zs_chain_qkvattn_blob_bytes(void) { return (size_t)4 * 4 * 2 * CT_TILE_BYTES; }

// points mode of zs_chain_qkvattn_fwd (flags must contain 8 | 16): x = LinearProj3D(points) is recomputed inside the kernel
extern "C" int zs_chain_qkvattn_pts_fwd(const float* points, int M, const float* pp, const float* pp_stat, float ln_eps,
                                        const void* Wblob, const float* bias_qkv, const void* Kblob, const void* Vblob, int n_keys,
                                        float scale, float* O, int precision, int flags, void* stream) {
  ZS_REQUIRE(points && pp && pp_stat && Wblob && bias_qkv && Kblob && Vblob && O && M >= 0, "zs_chain_qkvattn_pts_fwd: null pointer");
  ZS_REQUIRE(n_keys > 0 && n_keys <= 208, "zs_chain_qkvattn_pts_fwd: n_keys must be in [1, 208]");
  ZS_REQUIRE((flags & 24) == 24 && flags >= 0 && flags < 32, "zs_chain_qkvattn_pts_fwd: flags must contain 8 | 16");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(Wblob) & 15) == 0 && (reinterpret_cast<uintptr_t>(Kblob) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(Vblob) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias_qkv) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(pp) & 15) == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0,
             "zs_chain_qkvattn_pts_fwd: blobs / bias / pp / O must be 16-byte aligned");
  ZS_REQUIRE(precision == 0 || precision == 1, "zs_chain_qkvattn_pts_fwd: bad precision");
  if (M == 0) return ZS_OK;
  ChainParams p{};
  p.points = points; p.pp = pp; p.pp_stat = pp_stat; p.ldx = 256; p.M = M; p.ln_eps = ln_eps;
  p.blob = reinterpret_cast<const uint8_t*>(Wblob); p.bias = bias_qkv; p.kblob = reinterpret_cast<const uint8_t*>(Kblob);
  p.vblob = reinterpret_cast<const uint8_t*>(Vblob); p.n_keys = n_keys; p.scale = scale; p.out = O; p.ldo = 256;
  p.precision = precision; p.flags = flags | 64;
  return chain_launch(chain_qkvattn2_kernel<true>, p, as_stream(stream), "zs_chain_qkvattn_pts_fwd", QA3_THREADS);
}

extern "C" int zs_chain_qkvattn_fwd(const float* x, int ldx, int M, float ln_eps, const void* Wblob, const float* bias_qkv,
                                    const void* Kblob, const void* Vblob, int n_keys, float scale, float* O, int ldo,
                                    int precision, int flags, void* stream) {
  ZS_REQUIRE(x && Wblob && bias_qkv && Kblob && Vblob && O && M >= 0, "zs_chain_qkvattn_fwd: null pointer");
  ZS_REQUIRE(n_keys > 0 && n_keys <= 208, "zs_chain_qkvattn_fwd: n_keys must be in [1, 208]");
  ZS_REQUIRE(ldx >= 256 && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "zs_chain_qkvattn_fwd: x must be 16B aligned, ldx%%4==0");
  ZS_REQUIRE(ldo >= 256 && (ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0, "zs_chain_qkvattn_fwd: O must be 16B aligned, ldo%%4==0");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(Wblob) & 15) == 0 && (reinterpret_cast<uintptr_t>(Kblob) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(Vblob) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias_qkv) & 15) == 0,
             "zs_chain_qkvattn_fwd: blobs / bias must be 16-byte aligned");
  ZS_REQUIRE((precision == 0 || precision == 1) && flags >= 0 && flags < 32, "zs_chain_qkvattn_fwd: bad precision / flags");
  if (M == 0) return ZS_OK;
  ChainParams p{};
  p.x = const_cast<float*>(x); p.ldx = ldx; p.M = M; p.ln_eps = ln_eps; p.blob = reinterpret_cast<const uint8_t*>(Wblob);
  p.bias = bias_qkv; p.kblob = reinterpret_cast<const uint8_t*>(Kblob); p.vblob = reinterpret_cast<const uint8_t*>(Vblob);
  p.n_keys = n_keys; p.scale = scale; p.out = O; p.ldo = ldo; p.precision = precision; p.flags = flags | ((g_chain_dbg & 8) ? 32 : 0);
  // flags & 8: the variant that keeps the probabilities in tensor memory (chain_qkvattn2_kernel); & 16: its softmax role with
  // the scores held in registers (one tensor-memory sweep per head)
  if (flags & 16) return chain_launch(chain_qkvattn2_kernel<true>, p, as_stream(stream), "zs_chain_qkvattn_fwd", QA3_THREADS);
  return chain_launch((flags & 8) ? chain_qkvattn2_kernel<false> : chain_qkvattn_kernel, p, as_stream(stream), "zs_chain_qkvattn_fwd", QA_THREADS);
}
