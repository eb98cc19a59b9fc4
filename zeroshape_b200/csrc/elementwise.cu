// HBM-bound helpers: normalisations, pooling, resampling, layout changes, error plumbing.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "common.cuh"

namespace zs {

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int sm_count() {
  static thread_local int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

// ---- LayerNorm: one warp per row ------------------------------------------------------------
__global__ void layernorm_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ g,
                                 const float* __restrict__ b, float* __restrict__ y, int ldy, int rows, int cols,
                                 float eps) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (int64_t)row * ldx;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += xr[c];
  float mean = warp_sum(s) / cols;
  float v = 0.f;
  for (int c = lane; c < cols; c += 32) { float d = xr[c] - mean; v += d * d; }
  float rstd = rsqrtf(warp_sum(v) / cols + eps);
  float* yr = y + (int64_t)row * ldy;
  for (int c = lane; c < cols; c += 32) yr[c] = (xr[c] - mean) * rstd * g[c] + b[c];
}

// ---- GroupNorm NHWC: one CTA per (b, group): two-pass over HW x (C/groups) -------------------
__global__ void groupnorm_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ res,
                                      float* __restrict__ y, int HW, int C, int groups, float eps, int relu) {
  int b = blockIdx.x / groups, gi = blockIdx.x % groups;
  int cg = C / groups;
  const float* xb = x + (int64_t)b * HW * C + gi * cg;
  float* yb = y + (int64_t)b * HW * C + gi * cg;
  const float* rb = res ? res + (int64_t)b * HW * C + gi * cg : nullptr;
  int64_t n = (int64_t)HW * cg;
  __shared__ float sh[34];
  // pass 1: mean
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += xb[(i / cg) * C + (i % cg)];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) sh[32] = t / (float)n;
  }
  __syncthreads();
  float mean = sh[32];
  float v = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { float d = xb[(i / cg) * C + (i % cg)] - mean; v += d * d; }
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) sh[33] = rsqrtf(t / (float)n + eps);
  }
  __syncthreads();
  float rstd = sh[33];
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    int c = (int)(i % cg);
    int64_t off = (i / cg) * C + c;
    float o = (xb[off] - mean) * rstd * gamma[gi * cg + c] + beta[gi * cg + c];
    if (rb) o += rb[off];
    yb[off] = relu ? fmaxf(o, 0.f) : o;
  }
}

// ---- GroupNorm NHWC, tiled (C / 4 a power of two <= 256, the ResNetV2 / DPT shapes): two launches, both fully coalesced -------------
// The one-CTA-per-(image, group) kernel above walks 32-byte group slices with a 4 * C byte stride three times on B * 32 CTAs
// (47 us for [1,56,56,256], a quarter of the encoder's device time).  Here every CTA streams whole pixels (all channels, 128-bit
// loads); with 256 threads and C / 4 | 256 a thread always sees the same four channels, so its group(s), gamma and beta are fixed.
//   stats: grid (chunks, B); per lane fp32 sums of (x - c) and (x - c)^2 with c = the group's first element of the image (a shift
//          near the mean: no cancellation in E[d^2] - E[d]^2), combined in DOUBLE per group in shared memory, one partial per
//          (image, chunk, group) to the workspace -- no atomics on global memory, deterministic.
//   apply: grid (chunks', B); every CTA sums the <= GN_MAX_CHUNKS partials of its image in double -> mean, rstd per group, then
//          y = (x - mean) * rstd * gamma + beta (+ res) (ReLU) in the original operation order.
constexpr int GN_MAX_CHUNKS = 32;

__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const float* __restrict__ x, double* __restrict__ ws, int HW, int C,
                                                              int groups, int pix_per_chunk) {
  __shared__ double acc[2 * 256];                       // [group][sum, sumsq], groups <= 256
  const int b = blockIdx.y, chunk = blockIdx.x, c4 = C >> 2, cg = C / groups;
  for (int i = threadIdx.x; i < 2 * groups; i += 256) acc[i] = 0.0;
  __syncthreads();
  const int q = threadIdx.x % c4;                       // this thread's channel quad
  const float* xb = x + (int64_t)b * HW * C;
  float shift[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) shift[e] = __ldg(xb + ((q * 4 + e) / cg) * cg);
  const int p0 = chunk * pix_per_chunk, p1 = min(HW, p0 + pix_per_chunk);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
  const float4* x4 = reinterpret_cast<const float4*>(xb);
  for (int64_t i = (int64_t)p0 * c4 + threadIdx.x; i < (int64_t)p1 * c4; i += 256) {
    const float4 v = __ldg(x4 + i);
    const float d0 = v.x - shift[0], d1 = v.y - shift[1], d2 = v.z - shift[2], d3 = v.w - shift[3];
    s[0] += d0; s[1] += d1; s[2] += d2; s[3] += d3;
    ss[0] = fmaf(d0, d0, ss[0]); ss[1] = fmaf(d1, d1, ss[1]); ss[2] = fmaf(d2, d2, ss[2]); ss[3] = fmaf(d3, d3, ss[3]);
  }
  if (cg >= 4) {                                         // the quad lies inside one group
    const int g = (q * 4) / cg;
    atomicAdd(&acc[2 * g], (double)s[0] + (double)s[1] + (double)s[2] + (double)s[3]);
    atomicAdd(&acc[2 * g + 1], (double)ss[0] + (double)ss[1] + (double)ss[2] + (double)ss[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int g = (q * 4 + e) / cg;
      atomicAdd(&acc[2 * g], (double)s[e]);
      atomicAdd(&acc[2 * g + 1], (double)ss[e]);
    }
  }
  __syncthreads();
  double* o = ws + ((int64_t)b * gridDim.x + chunk) * 2 * groups;
  for (int i = threadIdx.x; i < 2 * groups; i += 256) o[i] = acc[i];
}

__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ res,
                                                              float* __restrict__ y, const double* __restrict__ ws, int HW, int C,
                                                              int groups, int stat_chunks, int pix_per_chunk, float eps, int relu) {
  __shared__ float smean[256], srstd[256];
  const int b = blockIdx.y, c4 = C >> 2, cg = C / groups;
  const float* xb = x + (int64_t)b * HW * C;
  for (int g = threadIdx.x; g < groups; g += 256) {
    double su = 0.0, sq = 0.0;
    for (int k = 0; k < stat_chunks; ++k) {
      const double* w = ws + ((int64_t)b * stat_chunks + k) * 2 * groups + 2 * g;
      su += w[0]; sq += w[1];
    }
    const double n = (double)HW * cg, md = su / n;
    double var = sq / n - md * md;
    if (var < 0.0) var = 0.0;
    smean[g] = (float)((double)__ldg(xb + g * cg) + md);
    srstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int q = threadIdx.x % c4;
  float mean[4], rstd[4], ga[4], be[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = q * 4 + e, g = c / cg;
    mean[e] = smean[g]; rstd[e] = srstd[g]; ga[e] = __ldg(gamma + c); be[e] = __ldg(beta + c);
  }
  const int p0 = blockIdx.x * pix_per_chunk, p1 = min(HW, p0 + pix_per_chunk);
  const float4* x4 = reinterpret_cast<const float4*>(xb);
  const float4* r4 = res ? reinterpret_cast<const float4*>(res + (int64_t)b * HW * C) : nullptr;
  float4* y4 = reinterpret_cast<float4*>(y + (int64_t)b * HW * C);
  for (int64_t i = (int64_t)p0 * c4 + threadIdx.x; i < (int64_t)p1 * c4; i += 256) {
    const float4 v = __ldg(x4 + i);
    float4 o;
    o.x = (v.x - mean[0]) * rstd[0] * ga[0] + be[0];
    o.y = (v.y - mean[1]) * rstd[1] * ga[1] + be[1];
    o.z = (v.z - mean[2]) * rstd[2] * ga[2] + be[2];
    o.w = (v.w - mean[3]) * rstd[3] * ga[3] + be[3];
    if (r4) { const float4 r = __ldg(r4 + i); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    y4[i] = o;
  }
}

__global__ void channel_affine_kernel(const float* __restrict__ x, const float* __restrict__ sc,
                                      const float* __restrict__ sh, const float* __restrict__ res,
                                      float* __restrict__ y, int64_t n, int C, int act) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float v = x[i] * sc[c] + sh[c];
    if (res) v += res[i];
    y[i] = apply_act(v, act);
  }
}

__global__ void axpby_kernel(const float* __restrict__ a, float alpha, const float* __restrict__ b, float beta,
                             float* __restrict__ y, int64_t n, int act) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = a[i] * alpha;
    if (b) v += b[i] * beta;
    y[i] = apply_act(v, act);
  }
}

__global__ void maxpool3x3s2_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C,
                                    int pt, int pl, int OH, int OW) {
  int64_t total = (int64_t)B * OH * OW * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t = i / C;
    int ow = (int)(t % OW); t /= OW;
    int oh = (int)(t % OH);
    int b = (int)(t / OH);
    float m = -INFINITY;
    for (int kh = 0; kh < 3; ++kh) {
      int ih = oh * 2 - pt + kh;
      if ((unsigned)ih >= (unsigned)H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        int iw = ow * 2 - pl + kw;
        if ((unsigned)iw >= (unsigned)W) continue;
        m = fmaxf(m, x[(((int64_t)b * H + ih) * W + iw) * C + c]);
      }
    }
    y[i] = m;
  }
}

__global__ void avgpool_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int HW, int C) {
  int b = blockIdx.y;
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* xb = x + (int64_t)b * HW * C + c;
  float s = 0.f;
  for (int i = 0; i < HW; ++i) s += xb[(int64_t)i * C];
  y[(int64_t)b * C + c] = s / (float)HW;
}

// PyTorch upsample_bilinear2d source-index rules (aten/native/UpSample.h):
//   align_corners: src = dst * (in-1)/(out-1)      (0 if out==1)
//   else         : src = max((dst+0.5)*in/out - 0.5, 0)
__device__ __forceinline__ void bilinear_src(int dst, int in, int out, int align, int& i0, int& i1, float& l1) {
  float src;
  if (align) {
    float sc = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    src = sc * dst;
  } else {
    float sc = (float)in / (float)out;
    src = sc * (dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
  }
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 < in - 1 ? i0 + 1 : i0;
  l1 = src - (float)i0;
}

__global__ void bilinear_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C,
                                     int OH, int OW, int align) {
  int64_t total = (int64_t)B * OH * OW * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t = i / C;
    int ow = (int)(t % OW); t /= OW;
    int oh = (int)(t % OH);
    int b = (int)(t / OH);
    int h0, h1, w0, w1;
    float lh, lw;
    bilinear_src(oh, H, OH, align, h0, h1, lh);
    bilinear_src(ow, W, OW, align, w0, w1, lw);
    const float* xb = x + (int64_t)b * H * W * C + c;
    float v00 = xb[((int64_t)h0 * W + w0) * C], v01 = xb[((int64_t)h0 * W + w1) * C];
    float v10 = xb[((int64_t)h1 * W + w0) * C], v11 = xb[((int64_t)h1 * W + w1) * C];
    float hh0 = 1.f - lh, ww0 = 1.f - lw;
    y[i] = hh0 * (ww0 * v00 + lw * v01) + lh * (ww0 * v10 + lw * v11);
  }
}

// four channels per thread (C % 4 == 0, 16-byte aligned rows): 128-bit loads / stores, 32-bit index arithmetic (the 64-bit
// divisions of the scalar kernel are software routines: it was instruction-bound at 30 % of the DRAM rate)
__global__ void bilinear_nhwc_v4_kernel(const float4* __restrict__ x, float4* __restrict__ y, int B, int H, int W, int C4,
                                        int OH, int OW, int align) {
  const unsigned total = (unsigned)B * OH * OW * C4;          // < 2^31 (checked by the caller)
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned pix = i / (unsigned)C4, c = i - pix * C4;
    const unsigned t = pix / (unsigned)OW, ow = pix - t * OW;
    const unsigned b = t / (unsigned)OH, oh = t - b * OH;
    int h0, h1, w0, w1;
    float lh, lw;
    bilinear_src((int)oh, H, OH, align, h0, h1, lh);
    bilinear_src((int)ow, W, OW, align, w0, w1, lw);
    const float4* xb = x + (size_t)b * H * W * C4 + c;
    const float4 v00 = __ldg(xb + (h0 * W + w0) * C4), v01 = __ldg(xb + (h0 * W + w1) * C4);
    const float4 v10 = __ldg(xb + (h1 * W + w0) * C4), v11 = __ldg(xb + (h1 * W + w1) * C4);
    const float hh0 = 1.f - lh, ww0 = 1.f - lw;
    float4 o;                                             // the same expression as the scalar kernel, per component
    o.x = hh0 * (ww0 * v00.x + lw * v01.x) + lh * (ww0 * v10.x + lw * v11.x);
    o.y = hh0 * (ww0 * v00.y + lw * v01.y) + lh * (ww0 * v10.y + lw * v11.y);
    o.z = hh0 * (ww0 * v00.z + lw * v01.z) + lh * (ww0 * v10.z + lw * v11.z);
    o.w = hh0 * (ww0 * v00.w + lw * v01.w) + lh * (ww0 * v10.w + lw * v11.w);
    __stcs(y + i, o);                                     // streamed: read once by the next layer, keep x in L2 instead
  }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C, int H, int W,
                                    float scale, float shift) {
  int64_t total = (int64_t)B * C * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t = i / C;
    int w = (int)(t % W); t /= W;
    int h = (int)(t % H);
    int b = (int)(t / H);
    y[i] = x[(((int64_t)b * C + c) * H + h) * W + w] * scale + shift;
  }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C, int H, int W) {
  int64_t total = (int64_t)B * C * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int w = (int)(i % W);
    int64_t t = i / W;
    int h = (int)(t % H); t /= H;
    int c = (int)(t % C);
    int b = (int)(t / C);
    y[i] = x[(((int64_t)b * H + h) * W + w) * C + c];
  }
}

__global__ void concat2_kernel(const float* __restrict__ a, int lda, int Ca, const float* __restrict__ b, int ldb, int Cb,
                               float s, float* __restrict__ y, int ldy, int64_t rows) {
  int Ct = Ca + Cb;
  int64_t total = rows * Ct;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % Ct);
    int64_t r = i / Ct;
    float v = c < Ca ? a[r * lda + c] : b[r * ldb + (c - Ca)];
    y[r * ldy + c] = v / s;
  }
}

// torch.linspace(start,end,steps) fp32 CPU/CUDA kernel semantics: step=(end-start)/(steps-1);
// i < steps/2 ? start + step*i : end - step*(steps-1-i)
__global__ void dense_grid_kernel(float* __restrict__ out, int n, float rmin, float rmax, int x0, int x1) {
  int64_t total = (int64_t)(x1 - x0) * n * n;
  float step = (rmax - rmin) / (float)(n - 1);
  int half = n / 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int k = (int)(i % n);
    int64_t t = i / n;
    int j = (int)(t % n);
    int ii = (int)(t / n) + x0;
    auto lin = [&](int q) { return q < half ? rmin + step * (float)q : rmax - step * (float)(n - 1 - q); };
    out[i * 3 + 0] = lin(ii);
    out[i * 3 + 1] = lin(j);
    out[i * 3 + 2] = lin(k);
  }
}

static inline int grid_for(int64_t n, int threads = 256) {
  int64_t g = (n + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace zs

using namespace zs;

extern "C" const char* zs_last_error(void) { return zs::g_err; }
extern "C" int zs_abi_version(void) { return 1; }
extern "C" long long zs_launch_count(void) { return zs::g_launches.load(std::memory_order_relaxed); }
// launches replayed from a CUDA graph never pass through the entry points: the graph owners (zeroshape_b200/graphed.py, the
// graphed inference encoder) add the number of launches recorded at capture time once per replay
extern "C" void zs_launch_count_add(long long n) { zs::g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int zs_device_cc(void) {
  int dev = 0, maj = 0, min = 0;
  ZS_CUDA_CALL(cudaGetDevice(&dev));
  ZS_CUDA_CALL(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  ZS_CUDA_CALL(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  return maj * 10 + min;
}

extern "C" int zs_layernorm_f32(const float* x, int ldx, const float* gamma, const float* beta, float* y, int ldy,
                                int rows, int cols, float eps, void* stream) {
  ZS_REQUIRE(x && gamma && beta && y && rows >= 0 && cols > 0, "zs_layernorm_f32: bad args");
  if (rows == 0) return ZS_OK;
  int wpb = 8;
  layernorm_kernel<<<(rows + wpb - 1) / wpb, wpb * 32, 0, as_stream(stream)>>>(x, ldx, gamma, beta, y, ldy, rows, cols, eps);
  ZS_CUDA_CHECK_LAUNCH("zs_layernorm_f32");
  return ZS_OK;
}

extern "C" int zs_groupnorm_nhwc_f32(const float* x, const float* gamma, const float* beta, const float* res, float* y,
                                     int B, int HW, int C, int groups, float eps, int relu, void* stream) {
  ZS_REQUIRE(x && gamma && beta && y && B > 0 && HW > 0 && C > 0 && groups > 0 && C % groups == 0,
             "zs_groupnorm_nhwc_f32: bad args");
  groupnorm_nhwc_kernel<<<B * groups, 512, 0, as_stream(stream)>>>(x, gamma, beta, res, y, HW, C, groups, eps, relu);
  ZS_CUDA_CHECK_LAUNCH("zs_groupnorm_nhwc_f32");
  return ZS_OK;
}

static bool gn_tiled_ok(const float* x, const float* res, const float* y, int HW, int C, int groups) {
  const int c4 = C >> 2;
  return (C & 3) == 0 && c4 >= 1 && c4 <= 256 && (c4 & (c4 - 1)) == 0 && groups <= 256 && HW >= 16 &&
         ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(res)) & 15) == 0;
}
static void gn_chunks(int HW, int& chunks, int& ppc) {
  chunks = (HW + 63) / 64;                                 // >= 64 pixels per stats CTA, at most GN_MAX_CHUNKS partials per image
  if (chunks > zs::GN_MAX_CHUNKS) chunks = zs::GN_MAX_CHUNKS;
  ppc = (HW + chunks - 1) / chunks;
  chunks = (HW + ppc - 1) / ppc;
}

extern "C" size_t zs_groupnorm_ws_bytes(int B, int HW, int C, int groups) {
  if (B <= 0 || HW <= 0 || C <= 0 || groups <= 0) return 0;
  return (size_t)B * zs::GN_MAX_CHUNKS * 2 * groups * sizeof(double);
}

extern "C" int zs_groupnorm_nhwc_ws_f32(const float* x, const float* gamma, const float* beta, const float* res, float* y,
                                        int B, int HW, int C, int groups, float eps, int relu, void* ws, void* stream) {
  ZS_REQUIRE(x && gamma && beta && y && B > 0 && HW > 0 && C > 0 && groups > 0 && C % groups == 0,
             "zs_groupnorm_nhwc_ws_f32: bad args");
  if (!ws || !gn_tiled_ok(x, res, y, HW, C, groups) || (reinterpret_cast<uintptr_t>(ws) & 7) != 0)
    return zs_groupnorm_nhwc_f32(x, gamma, beta, res, y, B, HW, C, groups, eps, relu, stream);
  int chunks, ppc;
  gn_chunks(HW, chunks, ppc);
  cudaStream_t st = as_stream(stream);
  groupnorm_stats_kernel<<<dim3(chunks, B), 256, 0, st>>>(x, reinterpret_cast<double*>(ws), HW, C, groups, ppc);
  ZS_CUDA_CHECK_LAUNCH("zs_groupnorm_nhwc_ws_f32(stats)");
  int achunks = (HW + 31) / 32;                            // apply: ~32 pixels per CTA, at most ~8 CTAs per SM in flight
  const int cap = (8 * sm_count() + B - 1) / B;
  if (achunks > cap) achunks = cap;
  const int appc = (HW + achunks - 1) / achunks;
  achunks = (HW + appc - 1) / appc;
  groupnorm_apply_kernel<<<dim3(achunks, B), 256, 0, st>>>(x, gamma, beta, res, y, reinterpret_cast<const double*>(ws), HW, C, groups,
                                                          chunks, appc, eps, relu);
  ZS_CUDA_CHECK_LAUNCH("zs_groupnorm_nhwc_ws_f32(apply)");
  return ZS_OK;
}

extern "C" int zs_channel_affine_f32(const float* x, const float* scale, const float* shift, const float* res,
                                     float* y, int64_t rows, int C, int act, void* stream) {
  ZS_REQUIRE(x && scale && shift && y && rows >= 0 && C > 0, "zs_channel_affine_f32: bad args");
  if (rows == 0) return ZS_OK;
  channel_affine_kernel<<<grid_for(rows * C), 256, 0, as_stream(stream)>>>(x, scale, shift, res, y, rows * C, C, act);
  ZS_CUDA_CHECK_LAUNCH("zs_channel_affine_f32");
  return ZS_OK;
}

extern "C" int zs_axpby_f32(const float* a, float alpha, const float* b, float beta, float* y, int64_t n, int act, void* stream) {
  ZS_REQUIRE(a && y && n >= 0, "zs_axpby_f32: bad args");
  if (n == 0) return ZS_OK;
  axpby_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(a, alpha, b, beta, y, n, act);
  ZS_CUDA_CHECK_LAUNCH("zs_axpby_f32");
  return ZS_OK;
}

extern "C" int zs_maxpool3x3s2_nhwc_f32(const float* x, float* y, int B, int H, int W, int C, int pad_top, int pad_left,
                                        int OH, int OW, void* stream) {
  ZS_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0 && OH > 0 && OW > 0, "zs_maxpool3x3s2_nhwc_f32: bad args");
  maxpool3x3s2_kernel<<<grid_for((int64_t)B * OH * OW * C), 256, 0, as_stream(stream)>>>(x, y, B, H, W, C, pad_top, pad_left, OH, OW);
  ZS_CUDA_CHECK_LAUNCH("zs_maxpool3x3s2_nhwc_f32");
  return ZS_OK;
}

extern "C" int zs_avgpool_nhwc_f32(const float* x, float* y, int B, int HW, int C, void* stream) {
  ZS_REQUIRE(x && y && B > 0 && HW > 0 && C > 0, "zs_avgpool_nhwc_f32: bad args");
  dim3 grid((C + 127) / 128, B);
  avgpool_nhwc_kernel<<<grid, 128, 0, as_stream(stream)>>>(x, y, HW, C);
  ZS_CUDA_CHECK_LAUNCH("zs_avgpool_nhwc_f32");
  return ZS_OK;
}

extern "C" int zs_bilinear_nhwc_f32(const float* x, float* y, int B, int H, int W, int C, int OH, int OW,
                                    int align_corners, void* stream) {
  ZS_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0 && OH > 0 && OW > 0, "zs_bilinear_nhwc_f32: bad args");
  if ((C & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 && (int64_t)B * OH * OW * (C / 4) < (1LL << 31) && (int64_t)H * W * (C / 4) < (1LL << 31)) {
    bilinear_nhwc_v4_kernel<<<grid_for((int64_t)B * OH * OW * (C / 4)), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), B, H, W, C / 4, OH, OW, align_corners);
    ZS_CUDA_CHECK_LAUNCH("zs_bilinear_nhwc_f32");
    return ZS_OK;
  }
  bilinear_nhwc_kernel<<<grid_for((int64_t)B * OH * OW * C), 256, 0, as_stream(stream)>>>(x, y, B, H, W, C, OH, OW, align_corners);
  ZS_CUDA_CHECK_LAUNCH("zs_bilinear_nhwc_f32");
  return ZS_OK;
}

extern "C" int zs_nchw_to_nhwc_f32(const float* x, float* y, int B, int C, int H, int W, float scale, float shift, void* stream) {
  ZS_REQUIRE(x && y && B > 0 && C > 0 && H > 0 && W > 0, "zs_nchw_to_nhwc_f32: bad args");
  nchw_to_nhwc_kernel<<<grid_for((int64_t)B * C * H * W), 256, 0, as_stream(stream)>>>(x, y, B, C, H, W, scale, shift);
  ZS_CUDA_CHECK_LAUNCH("zs_nchw_to_nhwc_f32");
  return ZS_OK;
}
extern "C" int zs_nhwc_to_nchw_f32(const float* x, float* y, int B, int C, int H, int W, void* stream) {
  ZS_REQUIRE(x && y && B > 0 && C > 0 && H > 0 && W > 0, "zs_nhwc_to_nchw_f32: bad args");
  nhwc_to_nchw_kernel<<<grid_for((int64_t)B * C * H * W), 256, 0, as_stream(stream)>>>(x, y, B, C, H, W);
  ZS_CUDA_CHECK_LAUNCH("zs_nhwc_to_nchw_f32");
  return ZS_OK;
}

extern "C" int zs_concat2_f32(const float* a, int lda, int Ca, const float* b, int ldb, int Cb, float s, float* y, int ldy,
                              int64_t rows, void* stream) {
  ZS_REQUIRE(a && b && y && Ca > 0 && Cb > 0 && rows >= 0 && ldy >= Ca + Cb, "zs_concat2_f32: bad args");
  if (rows == 0) return ZS_OK;
  concat2_kernel<<<grid_for(rows * (Ca + Cb)), 256, 0, as_stream(stream)>>>(a, lda, Ca, b, ldb, Cb, s, y, ldy, rows);
  ZS_CUDA_CHECK_LAUNCH("zs_concat2_f32");
  return ZS_OK;
}

extern "C" int zs_dense_grid_f32(float* out, int n, float rmin, float rmax, int x0, int x1, void* stream) {
  ZS_REQUIRE(out && n >= 2 && x0 >= 0 && x1 <= n && x0 <= x1, "zs_dense_grid_f32: bad args");
  if (x0 == x1) return ZS_OK;
  dense_grid_kernel<<<grid_for((int64_t)(x1 - x0) * n * n), 256, 0, as_stream(stream)>>>(out, n, rmin, rmax, x0, x1);
  ZS_CUDA_CHECK_LAUNCH("zs_dense_grid_f32");
  return ZS_OK;
}

// out[m, c] = W[c,0]*p[m,0] + W[c,1]*p[m,1] + W[c,2]*p[m,2] + b[c]    (LinearProj3D, model/shape/implicit.py:128-131)
// A K=3 "GEMM" is pure output bandwidth: one float4 of one row per thread, the thread's 4 weight rows in registers.
namespace zs {
__global__ void point_proj_kernel(const float* __restrict__ pts, int64_t M, const float* __restrict__ W, const float* __restrict__ b,
                                  float* __restrict__ out, int C) {
  const int c4 = C >> 2;                                   // float4 groups per row
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;  // a multiple of c4 (host guarantees) -> fixed column group per thread
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int cg = (int)(i0 % c4);
  float w[4][3], bb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cg * 4 + j;
    w[j][0] = __ldg(W + c * 3); w[j][1] = __ldg(W + c * 3 + 1); w[j][2] = __ldg(W + c * 3 + 2);
    bb[j] = b ? __ldg(b + c) : 0.f;
  }
  const int64_t total = M * c4;
  for (int64_t i = i0; i < total; i += stride) {
    const int64_t m = i / c4;
    const float x = __ldg(pts + m * 3), y = __ldg(pts + m * 3 + 1), z = __ldg(pts + m * 3 + 2);
    float4 v;
    // same association as the FFMA GEMM it replaces: ((b? no) k = 0,1,2 accumulated in order, bias added last
    v.x = fmaf(w[0][2], z, fmaf(w[0][1], y, w[0][0] * x)) + bb[0];
    v.y = fmaf(w[1][2], z, fmaf(w[1][1], y, w[1][0] * x)) + bb[1];
    v.z = fmaf(w[2][2], z, fmaf(w[2][1], y, w[2][0] * x)) + bb[2];
    v.w = fmaf(w[3][2], z, fmaf(w[3][1], y, w[3][0] * x)) + bb[3];
    reinterpret_cast<float4*>(out)[i] = v;
  }
}
}  // namespace zs

extern "C" int zs_point_proj_f32(const float* points, int64_t M, const float* W, const float* bias, float* out, int C, void* stream) {
  ZS_REQUIRE(points && W && out && M >= 0, "zs_point_proj_f32: null pointer");
  ZS_REQUIRE(C > 0 && (C & 3) == 0 && 256 % (C >> 2) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "zs_point_proj_f32: C/4 must divide 256 and out must be 16-byte aligned");
  if (M == 0) return ZS_OK;
  int64_t total = M * (C >> 2);
  int64_t blocks = (total + 255) / 256;
  int maxb = zs::sm_count() * 16;
  int grid = (int)(blocks < maxb ? blocks : maxb);
  zs::point_proj_kernel<<<grid, 256, 0, zs::as_stream(stream)>>>(points, M, W, bias, out, C);
  ZS_CUDA_CHECK_LAUNCH("zs_point_proj_f32");
  return ZS_OK;
}

// ---- CoordEmb front end (model/shape/seen_coord_enc.py:49-72): Linear(3 -> C) of every pixel of the XYZ map, the learned token
// for invalid pixels, the window partition, the fixed 2-D sin-cos embedding local to each window and the per-window cls row --
// seven tensor-sized torch ops (incl. two boolean scatters and a 6-D permute) as one output-bandwidth-bound launch.
//   out[(b, wy, wx), 0, c]        = cls[c] + pos[0, c]
//   out[(b, wy, wx), 1 + j, c]    = (mask ? w[c,:] . xyz + bias[c] : invalid[c]) + pos[1 + j, c],   j = dy * ws + dx
namespace zs {
__global__ void coord_embed_windows_kernel(const float* __restrict__ coord, const float* __restrict__ mask, const float* __restrict__ w,
                                           const float* __restrict__ bias, const float* __restrict__ invalid,
                                           const float* __restrict__ pos, const float* __restrict__ cls, float* __restrict__ out, int B,
                                           int H, int W, int C, int ws) {
  const int nwy = H / ws, nwx = W / ws, T = ws * ws + 1;
  const int64_t total = (int64_t)B * nwy * nwx * T * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t r = i / C;
    const int tok = (int)(r % T);
    const int64_t win = r / T;
    float v;
    if (tok == 0) {
      v = cls[c];
    } else {
      const int j = tok - 1, dy = j / ws, dx = j - dy * ws;
      const int wx = (int)(win % nwx), wy = (int)((win / nwx) % nwy), b = (int)(win / ((int64_t)nwx * nwy));
      const int64_t pix = ((int64_t)b * H + wy * ws + dy) * W + wx * ws + dx;
      if (mask[pix] > 0.5f) {
        const float x = coord[pix * 3], y = coord[pix * 3 + 1], z = coord[pix * 3 + 2];
        v = fmaf(z, w[c * 3 + 2], fmaf(y, w[c * 3 + 1], x * w[c * 3])) + bias[c];
      } else {
        v = invalid[c];
      }
    }
    out[i] = v + pos[(int64_t)tok * C + c];
  }
}
}  // namespace zs

extern "C" int zs_coord_embed_windows_f32(const float* coord, const float* mask, const float* w, const float* bias, const float* invalid,
                                          const float* pos, const float* cls, float* out, int B, int H, int W, int C, int ws,
                                          void* stream) {
  ZS_REQUIRE(coord && mask && w && bias && invalid && pos && cls && out, "zs_coord_embed_windows_f32: null pointer");
  ZS_REQUIRE(B > 0 && C > 0 && ws > 0 && H >= ws && W >= ws && H % ws == 0 && W % ws == 0,
             "zs_coord_embed_windows_f32: the map must tile into ws x ws windows");
  const int64_t total = (int64_t)B * (H / ws) * (W / ws) * (ws * ws + 1) * C;
  zs::coord_embed_windows_kernel<<<zs::grid_for(total), 256, 0, zs::as_stream(stream)>>>(coord, mask, w, bias, invalid, pos, cls, out, B, H, W,
                                                                                        C, ws);
  ZS_CUDA_CHECK_LAUNCH("zs_coord_embed_windows_f32");
  return ZS_OK;
}

// ---- debug: effective SM clock right now (cycles of clock64 per globaltimer microsecond over a ~20k-cycle spin) ----
// Used by tools/diag_decoder.py and bench.py to see power-cap clock droop between the tensor-heavy kernels.
namespace zs {
__global__ void clock_probe_kernel(float* out) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  const long long c0 = clock64();
  long long c1 = c0;
  while (c1 - c0 < 20000) c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  *out = t1 > t0 ? (float)((double)(c1 - c0) * 1e3 / (double)(t1 - t0)) : 0.f;   // MHz
}
}  // namespace zs

extern "C" int zs_debug_clock_mhz(float* out, void* stream) {
  ZS_REQUIRE(out, "zs_debug_clock_mhz: null pointer");
  zs::clock_probe_kernel<<<1, 1, 0, zs::as_stream(stream)>>>(out);
  ZS_CUDA_CHECK_LAUNCH("zs_debug_clock_mhz");
  return ZS_OK;
}

// ---- out[a, n] = mean_m x[a, m, n]  (Z-mean of the attention maps for the attention movie, utils/eval_3D.py:47-52) ----
namespace zs {
__global__ void mean_axis1_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t A, int M, int N) {
  const int64_t total = A * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = i / N;
    const int n = (int)(i % N);
    const float* p = x + a * (int64_t)M * N + n;
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += p[(int64_t)m * N];
    out[i] = s / (float)M;
  }
}
}  // namespace zs

extern "C" int zs_mean_axis1_f32(const float* x, float* out, int64_t A, int M, int N, void* stream) {
  ZS_REQUIRE(x && out && A >= 0 && M > 0 && N > 0, "zs_mean_axis1_f32: bad args");
  if (A == 0) return ZS_OK;
  int64_t blocks = (A * N + 255) / 256;
  int maxb = zs::sm_count() * 16;
  zs::mean_axis1_kernel<<<(int)(blocks < maxb ? blocks : maxb), 256, 0, zs::as_stream(stream)>>>(x, out, A, M, N);
  ZS_CUDA_CHECK_LAUNCH("zs_mean_axis1_f32");
  return ZS_OK;
}
