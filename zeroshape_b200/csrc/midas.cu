// MiDaS scale-and-shift-invariant depth loss, forward + gradient w.r.t. the prediction, on the device.
//
// Replaces `Loss.depth_loss` = `MidasLoss.forward` of the reference (utils/loss.py:30-34 -> model/depth/midas_loss.py:145-185
// with masked_shift_and_scale :33-62, masked_l1_loss :6-9, compute_scale_and_shift :11-30, gradient_loss :88-107 over 4 scales,
// image-based reduction :75-84; mask_shrink = False as in options/depth.yaml) and what torch autograd derives from it -- ~120
// small launches (boolean gathers, nanmedian sorts, strided slices) per call in the reference.
//
// Three launches, one CTA per image for the first two (a 224 x 224 map is 50k pixels = 49 per thread, everything stays in L2):
//   midas_stats_kernel : valid count, the two masked lower medians (4-pass radix select on order-preserving keys), mean absolute
//                        deviations, the SSI-MAE numerator and the three sums its gradient needs, the least-squares scale / shift of
//                        the (inverse) depth, r = m (q - t) for the gradient-matching term
//   midas_grad_kernel  : gradient-matching loss of 4 scales by GATHER (each pixel reads its +-1, 2, 4, 8 neighbours: no atomics),
//                        the chain through scale / shift (five sums), the inverse-depth map and the SSI alignment -> d loss / d pred
//   midas_finish_kernel: loss = sum_b ssi_b / (N + 1e-6) + alpha * sum_b reg_b
// The formulas are those of oracle/midas.py::midas_loss_grad, which is pinned to the reference module's autograd.
// Ties: when several valid pixels share the median value (e.g. predictions clamped to 0 or 1), the median's gradient goes to the LOWEST
// pixel index among them; torch.nanmedian's choice among equal values is implementation-defined (CPU and CUDA differ), the loss value and
// every other pixel's gradient are unaffected.
#include "common.cuh"

namespace zs {

constexpr int MD_THREADS = 1024;
constexpr int MD_STATS = 32;       // doubles per image in the workspace
// stats slots
enum { S_N = 0, S_TG, S_SG, S_TP, S_SP, S_MIDX, S_SSI, S_SE, S_SEP, S_SSG, S_X0, S_X1, S_A00, S_A01, S_A11, S_B0, S_B1, S_DET,
       S_M0, S_M1, S_M2, S_M3, S_REG };

__device__ __forceinline__ double block_sum(double v, double* red) {      // all threads get the sum
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < MD_THREADS / 32; ++w) s += red[w];
  return s;
}

__device__ __forceinline__ unsigned order_key(float x) {
  const unsigned u = __float_as_uint(x);
  return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key_value(unsigned k) {
  return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}

// k-th smallest (0-based) of x over the valid pixels: radix select, 8 bits per pass
__device__ unsigned radix_select(const float* __restrict__ x, const float* __restrict__ mask, int HW, int k, unsigned* hist,
                                 unsigned* bcast) {
  unsigned prefix = 0, known = 0;
  for (int pass = 3; pass >= 0; --pass) {
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += MD_THREADS) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < HW; i += MD_THREADS) {
      if (mask[i] > 0.5f) {
        const unsigned key = order_key(x[i]);
        if ((key & known) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned cum = 0;
      int b = 0;
      for (; b < 255; ++b) {
        if (cum + hist[b] > (unsigned)k) break;
        cum += hist[b];
      }
      bcast[0] = (unsigned)b;
      bcast[1] = cum;
    }
    __syncthreads();
    prefix |= bcast[0] << (8 * pass);
    known |= 0xffu << (8 * pass);
    k -= (int)bcast[1];
  }
  return prefix;
}

__global__ void __launch_bounds__(MD_THREADS) midas_stats_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                                 const float* __restrict__ mask, int H, int W, int inverse,
                                                                 double* __restrict__ stats, float* __restrict__ rbuf) {
  __shared__ double red[MD_THREADS / 32];
  __shared__ unsigned hist[256];
  __shared__ unsigned bcast[2];
  __shared__ int s_midx;
  const int b = blockIdx.x, HW = H * W;
  const float* P = pred + (size_t)b * HW;
  const float* T = gt + (size_t)b * HW;
  const float* Mk = mask + (size_t)b * HW;
  double* st = stats + (size_t)b * MD_STATS;
  // valid counts: all pixels and the 4 subsampled grids
  double cnt[5] = {0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < HW; i += MD_THREADS) {
    if (Mk[i] > 0.5f) {
      const int y = i / W, x = i - y * W;
      cnt[0] += 1.0;
#pragma unroll
      for (int s = 0; s < 4; ++s)
        if (((y | x) & ((1 << s) - 1)) == 0) cnt[1 + s] += 1.0;
    }
  }
  double tot[5];
  for (int j = 0; j < 5; ++j) tot[j] = block_sum(cnt[j], red);
  const int n = (int)tot[0];
  float tg = 0.f, tp = 0.f;
  int midx = -1;
  if (n > 0) {
    const int k = (n - 1) / 2;                                   // torch.nanmedian: the lower median
    tg = key_value(radix_select(T, Mk, HW, k, hist, bcast));
    const unsigned kp = radix_select(P, Mk, HW, k, hist, bcast);
    tp = key_value(kp);
    if (threadIdx.x == 0) s_midx = 0x7fffffff;
    __syncthreads();
    for (int i = threadIdx.x; i < HW; i += MD_THREADS)
      if (Mk[i] > 0.5f && order_key(P[i]) == kp) atomicMin(&s_midx, i);
    __syncthreads();
    midx = s_midx;
  }
  // mean absolute deviations
  double sg = 0.0, sp = 0.0;
  for (int i = threadIdx.x; i < HW; i += MD_THREADS) {
    if (Mk[i] > 0.5f) { sg += (double)fabsf(T[i] - tg); sp += (double)fabsf(P[i] - tp); }
  }
  const float s_gt = (float)(block_sum(sg, red) / (double)(n + 1));
  const float s_p = (float)(block_sum(sp, red) / (double)(n + 1));
  // SSI-MAE numerator and the sums of its gradient; least-squares sums of the (inverse) depth
  double ssi = 0.0, se = 0.0, sep = 0.0, ssg = 0.0, a00 = 0.0, a01 = 0.0, b0 = 0.0, b1 = 0.0;
  for (int i = threadIdx.x; i < HW; i += MD_THREADS) {
    if (Mk[i] > 0.5f) {
      const float dp = P[i] - tp;
      const float ap = dp / (s_p + 1e-6f), ag = (T[i] - tg) / (s_gt + 1e-6f);
      const float d = ap - ag;
      const float e = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      ssi += (double)fabsf(d);
      se += e; sep += (double)(e * dp);
      ssg += dp > 0.f ? 1.0 : (dp < 0.f ? -1.0 : 0.0);
      const float p = inverse ? 1.0f / (P[i] + 1e-6f) : P[i];
      const float t = inverse ? 1.0f / (T[i] + 1e-6f) : T[i];
      a00 += (double)p * p; a01 += p; b0 += (double)p * t; b1 += t;
    }
  }
  ssi = block_sum(ssi, red); se = block_sum(se, red); sep = block_sum(sep, red); ssg = block_sum(ssg, red);
  a00 = block_sum(a00, red); a01 = block_sum(a01, red); b0 = block_sum(b0, red); b1 = block_sum(b1, red);
  const double a11 = tot[0];
  const double det = a00 * a11 - a01 * a01;
  double x0 = 0.0, x1 = 0.0;
  if (det != 0.0) { x0 = (a11 * b0 - a01 * b1) / (det + 1e-6); x1 = (-a01 * b0 + a00 * b1) / (det + 1e-6); }
  if (threadIdx.x == 0) {
    st[S_N] = n; st[S_TG] = tg; st[S_SG] = s_gt; st[S_TP] = tp; st[S_SP] = s_p; st[S_MIDX] = midx;
    st[S_SSI] = ssi; st[S_SE] = se; st[S_SEP] = sep; st[S_SSG] = ssg; st[S_X0] = x0; st[S_X1] = x1;
    st[S_A00] = a00; st[S_A01] = a01; st[S_A11] = a11; st[S_B0] = b0; st[S_B1] = b1; st[S_DET] = det;
    st[S_M0] = tot[1]; st[S_M1] = tot[2]; st[S_M2] = tot[3]; st[S_M3] = tot[4];
  }
  // r = m (q - t), q = x0 p + x1: the quantity the gradient-matching term differences
  const float fx0 = (float)x0, fx1 = (float)x1;
  float* r = rbuf + (size_t)b * HW;
  for (int i = threadIdx.x; i < HW; i += MD_THREADS) {
    float v = 0.f;
    if (Mk[i] > 0.5f) {
      const float p = inverse ? 1.0f / (P[i] + 1e-6f) : P[i];
      const float t = inverse ? 1.0f / (T[i] + 1e-6f) : T[i];
      v = (fx0 * p + fx1) - t;
    }
    r[i] = v;
  }
}

__global__ void __launch_bounds__(MD_THREADS) midas_grad_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                                const float* __restrict__ mask, int B, int H, int W, int inverse,
                                                                float alpha, float grad_scale, double* __restrict__ stats,
                                                                const float* __restrict__ rbuf, float* __restrict__ gqbuf,
                                                                float* __restrict__ dpred) {
  __shared__ double red[MD_THREADS / 32];
  const int b = blockIdx.x, HW = H * W;
  const float* P = pred + (size_t)b * HW;
  const float* T = gt + (size_t)b * HW;
  const float* Mk = mask + (size_t)b * HW;
  const float* r = rbuf + (size_t)b * HW;
  float* gq = gqbuf + (size_t)b * HW;
  double* st = stats + (size_t)b * MD_STATS;
  double N = 1e-6;
  for (int j = 0; j < B; ++j) N += stats[(size_t)j * MD_STATS + S_N];
  float w[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const double M = st[S_M0 + s];
    w[s] = (float)((M != 0.0 ? 1.0 / M : 1.0) / (double)B);
  }
  const float x0 = (float)st[S_X0];
  double reg = 0.0, gx0 = 0.0, gx1 = 0.0;
  if (alpha > 0.f) {
    for (int i = threadIdx.x; i < HW; i += MD_THREADS) {
      const int y = i / W, x = i - y * W;
      float g = 0.f;
      if (Mk[i] > 0.5f) {
        const float ri = r[i];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int stp = 1 << s;
          if (((y | x) & (stp - 1)) != 0) break;                 // not on this (or any coarser) grid
          // left / right neighbours on the scale's grid
          if (x - stp >= 0 && Mk[i - stp] > 0.5f) { const float D = ri - r[i - stp]; g += w[s] * (D > 0.f ? 1.f : (D < 0.f ? -1.f : 0.f)); }
          if (x + stp < W && Mk[i + stp] > 0.5f) {
            const float D = r[i + stp] - ri;
            g -= w[s] * (D > 0.f ? 1.f : (D < 0.f ? -1.f : 0.f));
            reg += (double)(w[s] * fabsf(D));
          }
          if (y - stp >= 0 && Mk[i - stp * W] > 0.5f) { const float D = ri - r[i - stp * W]; g += w[s] * (D > 0.f ? 1.f : (D < 0.f ? -1.f : 0.f)); }
          if (y + stp < H && Mk[i + stp * W] > 0.5f) {
            const float D = r[i + stp * W] - ri;
            g -= w[s] * (D > 0.f ? 1.f : (D < 0.f ? -1.f : 0.f));
            reg += (double)(w[s] * fabsf(D));
          }
        }
        const float p = inverse ? 1.0f / (P[i] + 1e-6f) : P[i];
        gx0 += (double)g * p;
        gx1 += g;
      }
      gq[i] = g;
    }
    reg = block_sum(reg, red); gx0 = block_sum(gx0, red); gx1 = block_sum(gx1, red);
  }
  if (threadIdx.x == 0) st[S_REG] = reg;
  if (dpred == nullptr) return;
  // chain through the least-squares scale / shift (oracle/midas.py::midas_loss_grad)
  double Ga00 = 0.0, Ga01 = 0.0, Gb0 = 0.0;
  const double det = st[S_DET];
  if (alpha > 0.f && det != 0.0) {
    const double D = det + 1e-6, a01 = st[S_A01], a11 = st[S_A11], b0 = st[S_B0], b1 = st[S_B1], X0 = st[S_X0], X1 = st[S_X1];
    const double dx0_a00 = -X0 * a11 / D, dx0_a01 = (-b1 + 2 * a01 * X0) / D, dx0_b0 = a11 / D;
    const double dx1_a00 = (b1 - X1 * a11) / D, dx1_a01 = (-b0 + 2 * a01 * X1) / D, dx1_b0 = -a01 / D;
    Ga00 = gx0 * dx0_a00 + gx1 * dx1_a00;
    Ga01 = gx0 * dx0_a01 + gx1 * dx1_a01;
    Gb0 = gx0 * dx0_b0 + gx1 * dx1_b0;
  }
  // SSI alignment
  const double n = st[S_N];
  const float tp = (float)st[S_TP], tg = (float)st[S_TG], s_p = (float)st[S_SP], s_gt = (float)st[S_SG];
  const double c = 1.0 / ((double)s_p + 1e-6);
  const double Gs = -(c * c) / N * st[S_SEP];
  const double Gt = -c / N * st[S_SE] - Gs * st[S_SSG] / (n + 1.0);
  const int midx = (int)st[S_MIDX];
  float* out = dpred + (size_t)b * HW;
  for (int i = threadIdx.x; i < HW; i += MD_THREADS) {
    double g = 0.0;
    if (Mk[i] > 0.5f) {
      const float dp = P[i] - tp;
      const float d = dp / (s_p + 1e-6f) - (T[i] - tg) / (s_gt + 1e-6f);
      const double e = d > 0.f ? 1.0 : (d < 0.f ? -1.0 : 0.0);
      const double sgn = dp > 0.f ? 1.0 : (dp < 0.f ? -1.0 : 0.0);
      g = e * c / N + Gs * sgn / (n + 1.0);
      if (i == midx) g += Gt;
      if (alpha > 0.f && det != 0.0) {
        const float p = inverse ? 1.0f / (P[i] + 1e-6f) : P[i];
        const float t = inverse ? 1.0f / (T[i] + 1e-6f) : T[i];
        double gp = (double)gq[i] * x0 + (2.0 * p * Ga00 + Ga01 + (double)t * Gb0);
        if (inverse) gp *= -((double)p * p);
        g += (double)alpha * gp;
      }
    }
    out[i] = (float)(g * (double)grad_scale);
  }
}

__global__ void midas_finish_kernel(const double* __restrict__ stats, int B, float alpha, float* __restrict__ loss) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double N = 1e-6, ssi = 0.0, reg = 0.0;
  for (int b = 0; b < B; ++b) {
    N += stats[(size_t)b * MD_STATS + S_N];
    ssi += stats[(size_t)b * MD_STATS + S_SSI];
    reg += stats[(size_t)b * MD_STATS + S_REG];
  }
  *loss = (float)(ssi / N + (alpha > 0.f ? (double)alpha * reg : 0.0));
}

// MidasLoss.erode_mask (midas_loss.py:158-167): valid where a whole pool x pool block of the raw mask is 1
// (1 - mask -> max_pool2d(pool) -> nearest upsampling back to H x W -> == 0)
__global__ void mask_erode_kernel(const float* __restrict__ mask, float* __restrict__ out, int B, int H, int W, int pool) {
  const int Hp = H / pool, Wp = W / pool;
  const float sh = (float)Hp / (float)H, sw = (float)Wp / (float)W;
  const int64_t total = (int64_t)B * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / ((int64_t)W * H));
    const int py = min((int)floorf(y * sh), Hp - 1), px = min((int)floorf(x * sw), Wp - 1);
    float mx = -INFINITY;
    for (int dy = 0; dy < pool; ++dy)
      for (int dx = 0; dx < pool; ++dx) mx = fmaxf(mx, 1.0f - mask[((int64_t)b * H + py * pool + dy) * W + px * pool + dx]);
    out[i] = mx == 0.0f ? 1.0f : 0.0f;
  }
}

// ---- DepthMetric.compute_metrics (utils/eval_depth.py:41-110): least-squares alignment of the predicted disparity to the ground
// truth disparity per image, aligned depth map, then d > threshold fractions, RMSE, L1 and absolute relative error over the mask.
// One CTA per image: five sums -> scale / shift -> one sweep that writes the aligned depth and accumulates the metrics.
constexpr int DM_MAX_T = 8;
__global__ void __launch_bounds__(MD_THREADS) depth_metric_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                                  const float* __restrict__ mask, int H, int W, const float* __restrict__ thr,
                                                                  int T, float depth_cap, int disparity_input,
                                                                  float* __restrict__ metrics, float* __restrict__ depth_out) {
  __shared__ double red[MD_THREADS / 32];
  const int b = blockIdx.x, HW = H * W;
  const float* P = pred + (size_t)b * HW;
  const float* G = gt + (size_t)b * HW;
  const float* Mk = mask + (size_t)b * HW;
  double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0;
  for (int i = threadIdx.x; i < HW; i += MD_THREADS) {
    if (Mk[i] > 0.5f) {
      const float p = disparity_input ? P[i] : 1.0f / (P[i] + 1.e-6f);
      const float t = 1.0f / G[i];
      a00 += (double)p * p; a01 += p; a11 += 1.0; b0 += (double)p * t; b1 += t;
    }
  }
  a00 = block_sum(a00, red); a01 = block_sum(a01, red); a11 = block_sum(a11, red); b0 = block_sum(b0, red); b1 = block_sum(b1, red);
  const double det = a00 * a11 - a01 * a01;
  float x0 = 0.f, x1 = 0.f;
  if (det > 0.0) { x0 = (float)((a11 * b0 - a01 * b1) / det); x1 = (float)((-a01 * b0 + a00 * b1) / det); }
  const float dcap = depth_cap > 0.f ? 1.0f / depth_cap : 0.f;
  double cnt[DM_MAX_T] = {}, se = 0, l1 = 0, rel = 0;
  for (int i = threadIdx.x; i < HW; i += MD_THREADS) {
    const bool v = Mk[i] > 0.5f;
    const float p = v ? (disparity_input ? P[i] : 1.0f / (P[i] + 1.e-6f)) : 0.f;
    float al = x0 * p + x1;
    if (depth_cap > 0.f && al < dcap) al = dcap;
    const float d = 1.0f / al;
    depth_out[(size_t)b * HW + i] = d;
    if (v) {
      const float g = G[i];
      const float ratio = fmaxf(d / g, g / d);
      for (int k = 0; k < T; ++k) cnt[k] += ratio > thr[k] ? 1.0 : 0.0;
      const float df = d - g;
      se += (double)(df * df); l1 += (double)fabsf(df); rel += (double)(fabsf(df) / g);
    }
  }
  for (int k = 0; k < T; ++k) cnt[k] = block_sum(cnt[k], red);
  se = block_sum(se, red); l1 = block_sum(l1, red); rel = block_sum(rel, red);
  if (threadIdx.x == 0) {
    float* m = metrics + (size_t)b * (T + 3);
    const float n = (float)a11;
    for (int k = 0; k < T; ++k) m[k] = (float)cnt[k] / n;
    m[T] = sqrtf((float)se / n);
    m[T + 1] = (float)l1 / n;
    m[T + 2] = (float)rel / n;
  }
}

}  // namespace zs

using namespace zs;

extern "C" int zs_depth_metrics_f32(const float* pred, const float* gt, const float* mask, int B, int H, int W, const float* thresholds,
                                    int T, float depth_cap, int disparity_input, float* metrics, float* depth_out, void* stream) {
  ZS_REQUIRE(pred && gt && mask && metrics && depth_out && B > 0 && H > 0 && W > 0, "zs_depth_metrics_f32: bad args");
  ZS_REQUIRE(T >= 0 && T <= DM_MAX_T && (T == 0 || thresholds), "zs_depth_metrics_f32: 0 <= T <= 8 thresholds (device array)");
  depth_metric_kernel<<<B, MD_THREADS, 0, as_stream(stream)>>>(pred, gt, mask, H, W, thresholds, T, depth_cap, disparity_input, metrics,
                                                              depth_out);
  ZS_CUDA_CHECK_LAUNCH("zs_depth_metrics_f32");
  return ZS_OK;
}

extern "C" int zs_mask_erode_f32(const float* mask, float* out, int B, int H, int W, int pool, void* stream) {
  ZS_REQUIRE(mask && out && B > 0 && pool > 0 && H >= pool && W >= pool, "zs_mask_erode_f32: bad args");
  const int64_t total = (int64_t)B * H * W;
  int grid = (int)((total + 255) / 256);
  if (grid > 4096) grid = 4096;
  mask_erode_kernel<<<grid, 256, 0, as_stream(stream)>>>(mask, out, B, H, W, pool);
  ZS_CUDA_CHECK_LAUNCH("zs_mask_erode_f32");
  return ZS_OK;
}

extern "C" size_t zs_midas_ws_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return sizeof(double) * (size_t)B * MD_STATS + 2 * sizeof(float) * (size_t)B * H * W;
}

extern "C" int zs_midas_loss_f32(const float* pred, const float* gt, const float* mask, int B, int H, int W, float alpha,
                                 int inverse_depth, float grad_scale, void* ws, float* loss, float* dpred, void* stream) {
  ZS_REQUIRE(pred && gt && mask && ws && loss && B > 0 && H > 0 && W > 0, "zs_midas_loss_f32: bad args");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 7) == 0, "zs_midas_loss_f32: workspace must be 8-byte aligned");
  ZS_REQUIRE((int64_t)H * W < (1LL << 30), "zs_midas_loss_f32: map too large");
  double* stats = reinterpret_cast<double*>(ws);
  float* rbuf = reinterpret_cast<float*>(stats + (size_t)B * MD_STATS);
  float* gqbuf = rbuf + (size_t)B * H * W;
  cudaStream_t st = as_stream(stream);
  midas_stats_kernel<<<B, MD_THREADS, 0, st>>>(pred, gt, mask, H, W, inverse_depth, stats, rbuf);
  ZS_CUDA_CHECK_LAUNCH("zs_midas_loss_f32(stats)");
  midas_grad_kernel<<<B, MD_THREADS, 0, st>>>(pred, gt, mask, B, H, W, inverse_depth, alpha, grad_scale, stats, rbuf, gqbuf, dpred);
  ZS_CUDA_CHECK_LAUNCH("zs_midas_loss_f32(grad)");
  midas_finish_kernel<<<1, 32, 0, st>>>(stats, B, alpha, loss);
  ZS_CUDA_CHECK_LAUNCH("zs_midas_loss_f32(finish)");
  return ZS_OK;
}
