// fp32 FFMA GEMM / implicit-GEMM convolution with fused epilogue.
//
// This is the bit-faithful (plain fp32) path of the library: every nn.Linear / nn.Conv2d of the
// hot path can run through it with results that differ from the PyTorch fp32 reference only by
// summation order.  It is the first correct CUDA path of every stage and the in-library ground
// truth the tcgen05 kernels are compared against on the GPU.
//
//   C[m,n] = epi( sum_k A(m,k) * W[n,k] )         A(m,k) either a dense row-major matrix or an
//                                                  on-the-fly NHWC im2col gather.
// Tiling: 128x128x16 CTA tile, 256 threads, 8x8 register tile per thread, double-buffered smem.
#include "common.cuh"

namespace zs {

struct DenseA {
  const float* A;
  int lda;
  int M, K;
  bool vec;  // lda%4==0 && K%4==0 && 16B aligned
  struct Row { const float* p; bool ok; };
  __device__ __forceinline__ Row row(int m) const { return {A + (int64_t)m * lda, m < M}; }
  __device__ __forceinline__ float4 load4(const Row& r, int k) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!r.ok) return v;
    if (vec) {
      if (k < K) v = __ldg(reinterpret_cast<const float4*>(r.p + k));
    } else {
      if (k + 0 < K) v.x = __ldg(r.p + k + 0);
      if (k + 1 < K) v.y = __ldg(r.p + k + 1);
      if (k + 2 < K) v.z = __ldg(r.p + k + 2);
      if (k + 3 < K) v.w = __ldg(r.p + k + 3);
    }
    return v;
  }
};

struct ConvA {
  const float* x;
  int B, H, W, Cin, KH, KW, stride, pad_top, pad_left, OH, OW;
  int M, K;
  bool vec;  // Cin%4==0
  bool pre_relu;
  struct Row { int b, ih0, iw0; bool ok; };
  __device__ __forceinline__ Row row(int m) const {
    Row r;
    r.ok = m < M;
    int ow = m % OW;
    int t = m / OW;
    int oh = t % OH;
    r.b = t / OH;
    r.ih0 = oh * stride - pad_top;
    r.iw0 = ow * stride - pad_left;
    return r;
  }
  __device__ __forceinline__ float elem(const Row& r, int k) const {
    if (k >= K) return 0.f;
    int ci = k % Cin;
    int t = k / Cin;
    int kw = t % KW, kh = t / KW;
    int ih = r.ih0 + kh, iw = r.iw0 + kw;
    if ((unsigned)ih >= (unsigned)H || (unsigned)iw >= (unsigned)W) return 0.f;
    float v = __ldg(x + (((int64_t)r.b * H + ih) * W + iw) * Cin + ci);
    return pre_relu ? fmaxf(v, 0.f) : v;
  }
  __device__ __forceinline__ float4 load4(const Row& r, int k) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!r.ok) return v;
    if (vec) {
      if (k >= K) return v;
      int ci = k % Cin;
      int t = k / Cin;
      int kw = t % KW, kh = t / KW;
      int ih = r.ih0 + kh, iw = r.iw0 + kw;
      if ((unsigned)ih >= (unsigned)H || (unsigned)iw >= (unsigned)W) return v;
      v = __ldg(reinterpret_cast<const float4*>(x + (((int64_t)r.b * H + ih) * W + iw) * Cin + ci));
      if (pre_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    } else {
      v.x = elem(r, k); v.y = elem(r, k + 1); v.z = elem(r, k + 2); v.w = elem(r, k + 3);
    }
    return v;
  }
};

struct Epilogue {
  const float* bias;
  const float* res;
  int ldres;
  int res_mode;
  float* C;
  int ldc;
  int act;
};

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, NTHREADS = 256;
constexpr int PADM = 4;  // smem row padding (floats)

template <class ALoader>
__global__ void __launch_bounds__(NTHREADS, 2)
gemm_f32_kernel(ALoader a, const float* __restrict__ Wt, int ldw, bool wvec, int M, int N, int K, Epilogue ep) {
  __shared__ __align__(16) float As[2][BK][BM + PADM];
  __shared__ __align__(16) float Bs[2][BK][BN + PADM];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // loader mapping: each thread loads 2 float4 of A and 2 of W per k-tile
  const int lrow = tid >> 2;         // 0..63
  const int lk = (tid & 3) * 4;      // 0,4,8,12
  typename ALoader::Row ar0 = a.row(m0 + lrow), ar1 = a.row(m0 + lrow + 64);
  DenseA wl{Wt, ldw, N, K, wvec};
  DenseA::Row wr0 = wl.row(n0 + lrow), wr1 = wl.row(n0 + lrow + 64);

  const int tx = tid & 15, ty = tid >> 4;  // 16x16 thread grid, each 8x8 outputs (split 4+4 for conflict-free LDS.128)
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra0, ra1, rb0, rb1;
  auto gload = [&](int k0) {
    ra0 = a.load4(ar0, k0 + lk);
    ra1 = a.load4(ar1, k0 + lk);
    rb0 = wl.load4(wr0, k0 + lk);
    rb1 = wl.load4(wr1, k0 + lk);
  };
  auto sstore = [&](int buf) {
    As[buf][lk + 0][lrow] = ra0.x; As[buf][lk + 1][lrow] = ra0.y; As[buf][lk + 2][lrow] = ra0.z; As[buf][lk + 3][lrow] = ra0.w;
    As[buf][lk + 0][lrow + 64] = ra1.x; As[buf][lk + 1][lrow + 64] = ra1.y; As[buf][lk + 2][lrow + 64] = ra1.z; As[buf][lk + 3][lrow + 64] = ra1.w;
    Bs[buf][lk + 0][lrow] = rb0.x; Bs[buf][lk + 1][lrow] = rb0.y; Bs[buf][lk + 2][lrow] = rb0.z; Bs[buf][lk + 3][lrow] = rb0.w;
    Bs[buf][lk + 0][lrow + 64] = rb1.x; Bs[buf][lk + 1][lrow + 64] = rb1.y; Bs[buf][lk + 2][lrow + 64] = rb1.z; Bs[buf][lk + 3][lrow + 64] = rb1.w;
  };

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4 + 64]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4 + 64]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: rows m0 + ty*4 + {0..3} and +64; cols n0 + tx*4 + {0..3} and +64
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * 4 + (i & 3) + (i >> 2) * 64;
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      int nb = n0 + tx * 4 + jh * 64;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int n = nb + j;
        if (n >= N) continue;
        float v = acc[i][jh * 4 + j];
        if (ep.bias) v += __ldg(ep.bias + n);
        if (ep.res_mode == ZS_RES_BEFORE_ACT) v += __ldg(ep.res + (int64_t)m * ep.ldres + n);
        v = apply_act(v, ep.act);
        if (ep.res_mode == ZS_RES_AFTER_ACT) v += __ldg(ep.res + (int64_t)m * ep.ldres + n);
        ep.C[(int64_t)m * ep.ldc + n] = v;
      }
    }
  }
}

// small-N / small-M friendly variant: 32x32 tile, 4x... used when the 128x128 grid would leave
// most SMs idle (e.g. M=197 tokens) -- 64x64x16 tile, 256 threads, 4x4 per thread.
constexpr int SBM = 64, SBN = 64, SBK = 16;
template <class ALoader>
__global__ void __launch_bounds__(256, 3)
gemm_f32_small_kernel(ALoader a, const float* __restrict__ Wt, int ldw, bool wvec, int M, int N, int K, Epilogue ep) {
  __shared__ __align__(16) float As[2][SBK][SBM + PADM];
  __shared__ __align__(16) float Bs[2][SBK][SBN + PADM];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  const int lrow = tid >> 2;     // 0..63
  const int lk = (tid & 3) * 4;  // 0..12
  typename ALoader::Row ar0 = a.row(m0 + lrow);
  DenseA wl{Wt, ldw, N, K, wvec};
  DenseA::Row wr0 = wl.row(n0 + lrow);
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float4 ra0, rb0;
  auto gload = [&](int k0) { ra0 = a.load4(ar0, k0 + lk); rb0 = wl.load4(wr0, k0 + lk); };
  auto sstore = [&](int buf) {
    As[buf][lk + 0][lrow] = ra0.x; As[buf][lk + 1][lrow] = ra0.y; As[buf][lk + 2][lrow] = ra0.z; As[buf][lk + 3][lrow] = ra0.w;
    Bs[buf][lk + 0][lrow] = rb0.x; Bs[buf][lk + 1][lrow] = rb0.y; Bs[buf][lk + 2][lrow] = rb0.z; Bs[buf][lk + 3][lrow] = rb0.w;
  };
  const int nk = (K + SBK - 1) / SBK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * SBK);
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float av[4] = {a0.x, a0.y, a0.z, a0.w};
      float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (ep.bias) v += __ldg(ep.bias + n);
      if (ep.res_mode == ZS_RES_BEFORE_ACT) v += __ldg(ep.res + (int64_t)m * ep.ldres + n);
      v = apply_act(v, ep.act);
      if (ep.res_mode == ZS_RES_AFTER_ACT) v += __ldg(ep.res + (int64_t)m * ep.ldres + n);
      ep.C[(int64_t)m * ep.ldc + n] = v;
    }
  }
}

template <class ALoader>
static int launch_gemm(const ALoader& a, const float* W, int ldw, int M, int N, int K, const Epilogue& ep,
                       cudaStream_t st, const char* what) {
  bool wvec = (ldw % 4 == 0) && (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  long big_ctas = (long)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  if (big_ctas >= 2L * sm_count()) {
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    gemm_f32_kernel<ALoader><<<grid, NTHREADS, 0, st>>>(a, W, ldw, wvec, M, N, K, ep);
  } else {
    dim3 grid((N + SBN - 1) / SBN, (M + SBM - 1) / SBM);
    gemm_f32_small_kernel<ALoader><<<grid, 256, 0, st>>>(a, W, ldw, wvec, M, N, K, ep);
  }
  ZS_CUDA_CHECK_LAUNCH(what);
  return ZS_OK;
}

}  // namespace zs

extern "C" int zs_gemm_f32(const float* A, int lda, const float* W, int ldw, const float* bias,
                           const float* res, int ldres, int res_mode, float* C, int ldc,
                           int M, int N, int K, int act, void* stream) {
  using namespace zs;
  ZS_REQUIRE(A && W && C, "zs_gemm_f32: null pointer");
  ZS_REQUIRE(M >= 0 && N > 0 && K > 0, "zs_gemm_f32: bad shape M=%d N=%d K=%d", M, N, K);
  ZS_REQUIRE(lda >= K && ldw >= K && ldc >= N, "zs_gemm_f32: bad leading dims");
  ZS_REQUIRE(res_mode == ZS_RES_NONE || res != nullptr, "zs_gemm_f32: residual requested but res==NULL");
  if (M == 0) return ZS_OK;
  DenseA a{A, lda, M, K, (lda % 4 == 0) && (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0)};
  Epilogue ep{bias, res, ldres, res_mode, C, ldc, act};
  return launch_gemm(a, W, ldw, M, N, K, ep, as_stream(stream), "zs_gemm_f32");
}

extern "C" int zs_conv2d_nhwc_f32(const float* x, int B, int H, int W, int Cin,
                                  const float* w, const float* bias, const float* res, int res_mode,
                                  float* y, int Cout, int KH, int KW, int stride, int pad_top, int pad_left,
                                  int OH, int OW, int act, int pre_relu, void* stream) {
  using namespace zs;
  ZS_REQUIRE(x && w && y, "zs_conv2d_nhwc_f32: null pointer");
  ZS_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && OH > 0 && OW > 0,
             "zs_conv2d_nhwc_f32: bad shape");
  ZS_REQUIRE(res_mode == ZS_RES_NONE || res != nullptr, "zs_conv2d_nhwc_f32: residual requested but res==NULL");
  int64_t M64 = (int64_t)B * OH * OW;
  ZS_REQUIRE(M64 < (1LL << 31), "zs_conv2d_nhwc_f32: too many output pixels");
  int M = (int)M64, K = KH * KW * Cin;
  Epilogue ep{bias, res, Cout, res_mode, y, Cout, act};
  if (KH == 1 && KW == 1 && stride == 1 && pad_top == 0 && pad_left == 0 && !pre_relu) {
    DenseA a{x, Cin, M, K, (Cin % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)};
    return launch_gemm(a, w, K, M, Cout, K, ep, as_stream(stream), "zs_conv2d_nhwc_f32(1x1)");
  }
  ConvA a{x, B, H, W, Cin, KH, KW, stride, pad_top, pad_left, OH, OW, M, K,
          (Cin % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0), pre_relu != 0};
  return launch_gemm(a, w, K, M, Cout, K, ep, as_stream(stream), "zs_conv2d_nhwc_f32");
}

// =====================================================================================================================
// Convolution backward (training of the seen-surface encoder): both gradients are
// implicit GEMMs over the same NHWC tensors.
//   dgrad: dX[b,ih,iw,ci] = sum_{kh,kw,co} dY[b,oh,ow,co] w[co,kh,kw,ci],  oh = (ih + pad_top - kh) / stride (exact, in range)
//          = gemm_f32_kernel with rows = INPUT pixels, K = (kh,kw,co) gathered from dY, weights re-laid as Wd[ci][(kh,kw,co)]
//   wgrad: dW[co,(kh,kw,ci)] = sum_m dY[m,co] * im2col(x)[m,(kh,kw,ci)]     (A^T B with B gathered by ConvA)
namespace zs {

struct DgradA {
  const float* dy;
  int B, H, W, Cout, KH, KW, stride, pad_top, pad_left, OH, OW;
  int M, K;
  struct Row { int b, ih, iw; bool ok; };
  __device__ __forceinline__ Row row(int m) const {
    Row r;
    r.ok = m < M;
    r.iw = m % W;
    int t = m / W;
    r.ih = t % H;
    r.b = t / H;
    return r;
  }
  __device__ __forceinline__ float4 load4(const Row& r, int k) const {   // Cout % 4 == 0: a float4 never straddles a tap
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!r.ok || k >= K) return v;
    const int co = k % Cout;
    const int t = k / Cout;
    const int kw = t % KW, kh = t / KW;
    const int a = r.ih + pad_top - kh, bb = r.iw + pad_left - kw;
    if (a < 0 || bb < 0) return v;
    const int oh = a / stride, ow = bb / stride;
    if (oh * stride != a || ow * stride != bb || oh >= OH || ow >= OW) return v;
    return __ldg(reinterpret_cast<const float4*>(dy + (((int64_t)r.b * OH + oh) * OW + ow) * Cout + co));
  }
};

__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ dY, ConvA xa, float* __restrict__ dW, int M, int N,
                                                         int K, int m_per_split) {
  __shared__ float As[16][65], Bs[16][65];
  const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const int mb = blockIdx.z * m_per_split, me = min(M, mb + m_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int m0 = mb; m0 < me; m0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int r = i >> 6, c = i & 63;
      const int m = m0 + r;
      As[r][c] = (m < me && n0 + c < N) ? dY[(int64_t)m * N + n0 + c] : 0.f;
      float bv = 0.f;
      if (m < me && k0 + c < K) { ConvA::Row rw = xa.row(m); bv = xa.elem(rw, k0 + c); }
      Bs[r][c] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[r][ty * 4 + i]; b[i] = Bs[r][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + ty * 4 + i, k = k0 + tx * 4 + j;
      if (n < N && k < K) atomicAdd(dW + (int64_t)n * K + k, acc[i][j]);
    }
}

// Data gradient for a stem convolution (Cin <= 4, e.g. the 7x7 / stride-2 XYZ stem of the seen-surface encoder): as a GEMM it
// has N = Cin = 3 useful columns of a 128-wide tile.  Here one warp owns one input pixel: lanes run over the output channels of
// every filter tap that reaches the pixel (coalesced dy rows, L1-resident weights), a shuffle reduction finishes the Cin sums.
__global__ void __launch_bounds__(256) conv_dgrad_smallcin_kernel(const float* __restrict__ dy, const float* __restrict__ wd,
                                                                  float* __restrict__ dx, int B, int H, int W, int Cin, int Cout, int KH,
                                                                  int KW, int stride, int pad_top, int pad_left, int OH, int OW) {
  const int64_t pix = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pix >= (int64_t)B * H * W) return;
  const int iw = (int)(pix % W);
  const int64_t t = pix / W;
  const int ih = (int)(t % H), b = (int)(t / H);
  const int K = KH * KW * Cout;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int kh = 0; kh < KH; ++kh) {
    const int a = ih + pad_top - kh;
    if (a < 0) break;                                   // a only decreases with kh
    const int oh = a / stride;
    if (oh * stride != a || oh >= OH) continue;
    for (int kw = 0; kw < KW; ++kw) {
      const int c = iw + pad_left - kw;
      if (c < 0) break;
      const int ow = c / stride;
      if (ow * stride != c || ow >= OW) continue;
      const float* g = dy + (((int64_t)b * OH + oh) * OW + ow) * Cout;
      const float* wk = wd + (kh * KW + kw) * Cout;
      for (int co = lane; co < Cout; co += 32) {
        const float gv = __ldg(g + co);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
          if (ci < Cin) acc[ci] = fmaf(gv, __ldg(wk + (int64_t)ci * K + co), acc[ci]);
      }
    }
  }
#pragma unroll
  for (int ci = 0; ci < 4; ++ci) acc[ci] = warp_sum(acc[ci]);
  if (lane < Cin) dx[pix * Cin + lane] = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
}

// The stride-2 stem (7x7, Cin = 3, Cout = 64 at 224 x 224: 1.6 M input pixels per batch of 32) one 16 x 16 tile of input pixels per
// CTA.  The dy patch the tile touches ((16 + KH) / 2 + 1 rows and columns, zero outside the image, rows padded to Cout + 4 floats
// against bank conflicts) and the whole filter sit in shared memory.  The pixels of a warp share their parity (ih & 1, iw & 1),
// so the taps that reach them -- kh = (ih + pad) & 1, +2, ... -- are the same for the whole warp: the filter reads are
// broadcasts, the dy reads hit 32 different rows, and the inner loop has no bounds test and no reduction across lanes.
constexpr int DG2_TILE = 16;
constexpr int DG2_MAXK = 7;
constexpr int DG2_OD = DG2_TILE / 2 + (DG2_MAXK + 1) / 2 + 1;      // 13

__device__ __forceinline__ int floor_div2(int a) { return a >= 0 ? a / 2 : -((1 - a) / 2); }

__global__ void __launch_bounds__(256) conv_dgrad_stem_s2_kernel(const float* __restrict__ dy, const float* __restrict__ wd,
                                                                 float* __restrict__ dx, int B, int H, int W, int Cin, int Cout, int KH,
                                                                 int KW, int pad_top, int pad_left, int OH, int OW) {
  extern __shared__ __align__(16) float dg_smem[];
  const int ldy = Cout + 4;                                        // floats per dy pixel in shared memory
  float* sdy = dg_smem;                                            // [DG2_OD][DG2_OD][ldy]
  float* sw = dg_smem + DG2_OD * DG2_OD * ldy;                     // [Cin][KH * KW][Cout]
  const int tiles_w = (W + DG2_TILE - 1) / DG2_TILE, tiles_h = (H + DG2_TILE - 1) / DG2_TILE;
  const int tw = blockIdx.x % tiles_w, th = (blockIdx.x / tiles_w) % tiles_h, b = blockIdx.x / (tiles_w * tiles_h);
  const int ih0 = th * DG2_TILE, iw0 = tw * DG2_TILE;
  const int oh0 = floor_div2(ih0 + pad_top - (KH - 1)), ow0 = floor_div2(iw0 + pad_left - (KW - 1));
  const int c4 = Cout / 4, taps = KH * KW;
  for (int i = threadIdx.x; i < DG2_OD * DG2_OD * c4; i += 256) {
    const int q = i % c4, px = i / c4;
    const int oh = oh0 + px / DG2_OD, ow = ow0 + px % DG2_OD;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (oh >= 0 && oh < OH && ow >= 0 && ow < OW) v = __ldg(reinterpret_cast<const float4*>(dy + (((int64_t)b * OH + oh) * OW + ow) * Cout) + q);
    *reinterpret_cast<float4*>(sdy + px * ldy + q * 4) = v;
  }
  for (int i = threadIdx.x; i < Cin * taps * c4; i += 256)
    *reinterpret_cast<float4*>(sw + i * 4) = __ldg(reinterpret_cast<const float4*>(wd) + i);
  __syncthreads();
  // 4 parity classes x 64 pixels: warps 2c, 2c + 1 own class c = (py, px)
  const int cls = threadIdx.x >> 6, j = threadIdx.x & 63;
  const int py = cls >> 1, px_ = cls & 1;
  const int ih = ih0 + 2 * (j >> 3) + py, iw = iw0 + 2 * (j & 7) + px_;
  float acc[4][4];
#pragma unroll
  for (int ci = 0; ci < 4; ++ci)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[ci][e] = 0.f;
  for (int kh = (ih + pad_top) & 1; kh < KH; kh += 2) {
    const int r = (ih + pad_top - kh) / 2 - oh0;                   // exact: the numerator is even
    for (int kw = (iw + pad_left) & 1; kw < KW; kw += 2) {
      const int c = (iw + pad_left - kw) / 2 - ow0;
      const float4* g = reinterpret_cast<const float4*>(sdy + (r * DG2_OD + c) * ldy);
      const float4* wk = reinterpret_cast<const float4*>(sw + (kh * KW + kw) * Cout);
      for (int q = 0; q < c4; ++q) {
        const float4 gv = g[q];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
          if (ci < Cin) {
            const float4 wv = wk[ci * taps * c4 + q];
            acc[ci][0] = fmaf(gv.x, wv.x, acc[ci][0]);
            acc[ci][1] = fmaf(gv.y, wv.y, acc[ci][1]);
            acc[ci][2] = fmaf(gv.z, wv.z, acc[ci][2]);
            acc[ci][3] = fmaf(gv.w, wv.w, acc[ci][3]);
          }
      }
    }
  }
  if (ih < H && iw < W) {
    float* o = dx + (((int64_t)b * H + ih) * W + iw) * Cin;
#pragma unroll
    for (int ci = 0; ci < 4; ++ci)
      if (ci < Cin) o[ci] = (acc[ci][0] + acc[ci][1]) + (acc[ci][2] + acc[ci][3]);
  }
}

}  // namespace zs

extern "C" int zs_conv2d_nhwc_dgrad_f32(const float* dy, int B, int H, int W, int Cin, const float* w_dgrad, float* dx, int Cout,
                                        int KH, int KW, int stride, int pad_top, int pad_left, int OH, int OW, void* stream) {
  using namespace zs;
  ZS_REQUIRE(dy && w_dgrad && dx, "zs_conv2d_nhwc_dgrad_f32: null pointer");
  ZS_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (Cout & 3) == 0 && KH > 0 && KW > 0 && stride > 0 && OH > 0 && OW > 0,
             "zs_conv2d_nhwc_dgrad_f32: bad shape (Cout %% 4 == 0 required)");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(dy) & 15) == 0, "zs_conv2d_nhwc_dgrad_f32: dy must be 16-byte aligned");
  const int64_t M64 = (int64_t)B * H * W;
  ZS_REQUIRE(M64 < (1LL << 31), "zs_conv2d_nhwc_dgrad_f32: too many pixels");
  const int M = (int)M64, K = KH * KW * Cout;
  if (Cin <= 4 && stride == 2 && KH <= DG2_MAXK && KW <= DG2_MAXK && Cout <= 64 && (reinterpret_cast<uintptr_t>(w_dgrad) & 15) == 0) {
    const size_t smem = sizeof(float) * ((size_t)DG2_OD * DG2_OD * (Cout + 4) + (size_t)Cin * KH * KW * Cout);
    static bool attr_set = false;
    if (!attr_set) {
      ZS_CUDA_CALL(cudaFuncSetAttribute(conv_dgrad_stem_s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr_set = true;
    }
    const int64_t blocks = (int64_t)B * ((H + DG2_TILE - 1) / DG2_TILE) * ((W + DG2_TILE - 1) / DG2_TILE);
    ZS_REQUIRE(smem <= 100 * 1024 && blocks < (1LL << 31), "zs_conv2d_nhwc_dgrad_f32: stem tile does not fit");
    conv_dgrad_stem_s2_kernel<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(dy, w_dgrad, dx, B, H, W, Cin, Cout, KH, KW, pad_top,
                                                                                 pad_left, OH, OW);
    ZS_CUDA_CHECK_LAUNCH("zs_conv2d_nhwc_dgrad_f32(stride-2 stem)");
    return ZS_OK;
  }
  if (Cin <= 4) {
    conv_dgrad_smallcin_kernel<<<(unsigned)((M64 + 7) / 8), 256, 0, as_stream(stream)>>>(dy, w_dgrad, dx, B, H, W, Cin, Cout, KH, KW, stride,
                                                                                       pad_top, pad_left, OH, OW);
    ZS_CUDA_CHECK_LAUNCH("zs_conv2d_nhwc_dgrad_f32(small Cin)");
    return ZS_OK;
  }
  DgradA a{dy, B, H, W, Cout, KH, KW, stride, pad_top, pad_left, OH, OW, M, K};
  Epilogue ep{nullptr, nullptr, Cin, ZS_RES_NONE, dx, Cin, ZS_ACT_NONE};
  return launch_gemm(a, w_dgrad, K, M, Cin, K, ep, as_stream(stream), "zs_conv2d_nhwc_dgrad_f32");
}

extern "C" int zs_conv2d_nhwc_wgrad_f32(const float* x, int B, int H, int W, int Cin, const float* dy, float* dw, int Cout, int KH,
                                        int KW, int stride, int pad_top, int pad_left, int OH, int OW, int accumulate, void* stream) {
  using namespace zs;
  ZS_REQUIRE(x && dy && dw, "zs_conv2d_nhwc_wgrad_f32: null pointer");
  ZS_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && OH > 0 && OW > 0,
             "zs_conv2d_nhwc_wgrad_f32: bad shape");
  const int64_t M64 = (int64_t)B * OH * OW;
  ZS_REQUIRE(M64 < (1LL << 31), "zs_conv2d_nhwc_wgrad_f32: too many pixels");
  const int M = (int)M64, K = KH * KW * Cin;
  cudaStream_t st = as_stream(stream);
  if (!accumulate) ZS_CUDA_CALL(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * K, st));
  ConvA a{x, B, H, W, Cin, KH, KW, stride, pad_top, pad_left, OH, OW, M, K, false, false};
  const int tiles = ((Cout + 63) / 64) * ((K + 63) / 64);
  int splits = (2 * sm_count() + tiles - 1) / tiles;
  const int max_splits = (M + 127) / 128;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int mps = ((M + splits - 1) / splits + 15) / 16 * 16;
  splits = (M + mps - 1) / mps;
  dim3 grid((Cout + 63) / 64, (K + 63) / 64, splits);
  conv_wgrad_kernel<<<grid, 256, 0, st>>>(dy, a, dw, M, Cout, K, mps);
  ZS_CUDA_CHECK_LAUNCH("zs_conv2d_nhwc_wgrad_f32");
  return ZS_OK;
}
