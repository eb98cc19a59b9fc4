// Backward / training kernels of the implicit decoder (first slice: the decoder's
// training step -- Implicit.forward on the GT sample points, BCE shape loss, backward, AdamW).
//   reference: model/compute_graph/graph_shape.py:185 (pred_sample_occ), utils/loss.py:18-28 (shape_loss),
//              model/shape_engine.py:248-277 (loss.backward(); optim.step()), torch autograd for every layer of
//              model/shape/implicit.py.
// All kernels are fp32 (gradient parity against torch autograd on the oracle); no tensor-core path yet.
#include <cstring>
#include "common.cuh"

namespace zs {

static inline int grid_for_n(int64_t n, int block = 256) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * 32;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// ---- BCE-with-logits shape loss (utils/loss.py:18-28) ---------------------------------------------------------
// loss = mean_i w_i * bce(x_i, y_i), y = (sdf < 0), w = impt_weight where |sdf| < impt_thres else 1
// bce(x, y) = max(x, 0) - x*y + log1p(exp(-|x|))  (torch's stable form);  dloss/dx_i = w_i (sigmoid(x_i) - y_i) / n
__global__ void bce_fwd_kernel(const float* __restrict__ x, const float* __restrict__ sdf, int64_t n, float thres, float weight,
                               double* __restrict__ acc) {
  double local = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xi = x[i], s = sdf[i];
    const float y = s < 0.f ? 1.f : 0.f;
    const float w = fabsf(s) < thres ? weight : 1.f;
    local += (double)(w * (fmaxf(xi, 0.f) - xi * y + log1pf(expf(-fabsf(xi)))));
  }
  __shared__ double red[256];
  red[threadIdx.x] = local;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(acc, red[0]);
}
__global__ void bce_finish_kernel(const double* acc, int64_t n, float* loss) { *loss = (float)(*acc / (double)n); }
__global__ void bce_bwd_kernel(const float* __restrict__ x, const float* __restrict__ sdf, int64_t n, float thres, float weight,
                               float gscale, float* __restrict__ dx) {
  const float inv_n = gscale / (float)n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xi = x[i], s = sdf[i];
    const float y = s < 0.f ? 1.f : 0.f;
    const float w = fabsf(s) < thres ? weight : 1.f;
    dx[i] = w * (1.0f / (1.0f + expf(-xi)) - y) * inv_n;
  }
}

// ---- activation backward: dx = dy * act'(z), z = pre-activation -------------------------------------------------
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, float* __restrict__ dx, int64_t n, int act) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float zi = z[i];
    float d;
    if (act == ZS_ACT_GELU) {              // d/dz [0.5 z (1 + erf(z/sqrt2))] = 0.5 (1 + erf(z/sqrt2)) + z exp(-z^2/2) / sqrt(2 pi)
      d = 0.5f * (1.0f + erff(zi * 0.70710678118654752440f)) + zi * expf(-0.5f * zi * zi) * 0.39894228040143267794f;
    } else if (act == ZS_ACT_SOFTPLUS100) {  // softplus(beta=100, threshold=20): sigmoid(100 z), 1 past the threshold
      const float bz = 100.0f * zi;
      d = bz > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-bz));
    } else if (act == ZS_ACT_RELU) {
      d = zi > 0.f ? 1.f : 0.f;
    } else if (act == ZS_ACT_CLAMP01) {      // clamp(relu(z), 0, 1): passes gradient strictly inside (0, 1)
      d = (zi > 0.f && zi < 1.f) ? 1.f : 0.f;
    } else {
      d = 1.f;
    }
    dx[i] = dy[i] * d;
  }
}

// ---- column sums (bias gradients): out[n] (+)= sum_m A[m, n] ----------------------------------------------------
__global__ void colsum_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B2, int ldb, int64_t M, int N,
                              float* __restrict__ out) {
  // block = 32 x 8: 32 consecutive columns, 8 row lanes; grid.x over column groups, grid.y over row slabs (atomics combine slabs)
  // B2 != nullptr: column-wise dot products sum_m A[m,n] * B2[m,n]
  const int n = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (n < N)
    for (int64_t m = blockIdx.y * 8 + threadIdx.y; m < M; m += (int64_t)gridDim.y * 8)
      acc += B2 ? A[m * lda + n] * B2[m * ldb + n] : A[m * lda + n];
  __shared__ float red[8][33];
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(out + n, s);
  }
}

// ---- C[N,K] += A[M,N]^T B[M,K]  (weight gradients: dW = dY^T X) --------------------------------------------------
// 64x64 output tile per CTA, 256 threads x (4x4) outputs, M split over grid.z and combined with atomics.
__global__ void __launch_bounds__(256) gemm_tn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                      float* __restrict__ C, int ldc, int64_t M, int N, int K, int64_t m_per_split) {
  __shared__ float As[16][65], Bs[16][65];
  const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const int64_t mb = blockIdx.z * m_per_split, me = (mb + m_per_split < M) ? mb + m_per_split : M;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // outputs n = n0 + ty*4 + i, k = k0 + tx*4 + j
  float acc[4][4] = {};
  for (int64_t m0 = mb; m0 < me; m0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int r = i >> 6, c = i & 63;
      const int64_t m = m0 + r;
      As[r][c] = (m < me && n0 + c < N) ? A[m * lda + n0 + c] : 0.f;
      Bs[r][c] = (m < me && k0 + c < K) ? B[m * ldb + k0 + c] : 0.f;
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
      if (n < N && k < K) atomicAdd(C + (int64_t)n * ldc + k, acc[i][j]);
    }
}

// ---- LayerNorm backward (C == 256): warp per row --------------------------------------------------------------
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma;  dgamma += dy * xhat, dbeta += dy
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ gamma, float eps, float* __restrict__ dx,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float gam[8], dg[8], db[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { gam[i] = gamma ? gamma[lane + 32 * i] : 1.f; dg[i] = 0.f; db[i] = 0.f; }
  for (int64_t r = blockIdx.x * 8 + warp; r < rows; r += (int64_t)gridDim.x * 8) {
    float xv[8], dv[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { xv[i] = x[r * 256 + lane + 32 * i]; dv[i] = dy[r * 256 + lane + 32 * i]; s += xv[i]; }
    const float mean = warp_sum(s) * (1.0f / 256.0f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { xv[i] -= mean; q += xv[i] * xv[i]; }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / 256.0f) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      xv[i] *= rstd;                      // xhat
      const float g = dv[i] * gam[i];
      s1 += g; s2 += g * xv[i];
      dg[i] += dv[i] * xv[i]; db[i] += dv[i];
    }
    s1 = warp_sum(s1) * (1.0f / 256.0f);
    s2 = warp_sum(s2) * (1.0f / 256.0f);
#pragma unroll
    for (int i = 0; i < 8; ++i) dx[r * 256 + lane + 32 * i] = rstd * (dv[i] * gam[i] - s1 - xv[i] * s2);
  }
  if (dgamma == nullptr) return;
  __shared__ float rg[8][256], rb[8][256];
#pragma unroll
  for (int i = 0; i < 8; ++i) { rg[warp][lane + 32 * i] = dg[i]; rb[warp][lane + 32 * i] = db[i]; }
  __syncthreads();
  const int c = threadIdx.x;
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { a += rg[w][c]; b += rb[w][c]; }
  atomicAdd(dgamma + c, a);
  atomicAdd(dbeta + c, b);
}

// ---- point -> (latents + self) attention backward (head dim 32) --------------------------------------------------
// One CTA per (head, image): walks the image's P query points in tiles of 256 (thread = point) and keeps the latent-side
// gradients dK_lat_h, dV_lat_h [L,32] of its (image, head) in shared memory for the whole walk -> no atomics, deterministic.
// Per tile and block of 16 keys the threads publish ds[p][j] and p[p][j]; the CTA then reduces dK += ds^T Q, dV += p^T dO.
constexpr int PAB_NT = 256, PAB_KB = 16;     // threads (= points per tile) and latent keys per reduction block
template <int HD>
__global__ void __launch_bounds__(PAB_NT) point_attention_bwd_kernel(
    const float* __restrict__ qkv, const float* __restrict__ k_lat, const float* __restrict__ v_lat, int ld_lat,
    const float* __restrict__ O, const float* __restrict__ dO, float* __restrict__ dqkv, float* __restrict__ dk_lat,
    float* __restrict__ dv_lat, int ld_dlat, int P, int L, int heads, float scale) {
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;                   // [L][HD]
  float* Vs = Ks + L * HD;          // [L][HD]
  float* dKs = Vs + L * HD;         // [L][HD]
  float* dVs = dKs + L * HD;        // [L][HD]
  float* Qs = dVs + L * HD;         // [NT][HD+1]
  float* Gs = Qs + PAB_NT * (HD + 1);  // [NT][HD+1]  (dO)
  float* dsT = Gs + PAB_NT * (HD + 1); // [KB][NT+1]
  float* pT = dsT + PAB_KB * (PAB_NT + 1);       // [KB][NT+1]
  const int h = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const int C = heads * HD;
  for (int i = t; i < L * HD; i += PAB_NT) {
    const int j = i / HD, d = i % HD;
    Ks[i] = k_lat[((int64_t)b * L + j) * ld_lat + h * HD + d];
    Vs[i] = v_lat[((int64_t)b * L + j) * ld_lat + h * HD + d];
    dKs[i] = 0.f; dVs[i] = 0.f;
  }
  __syncthreads();
  for (int p0 = 0; p0 < P; p0 += PAB_NT) {
    const int p = p0 + t;
    const bool live = p < P;
    const int64_t row = (int64_t)b * P + (live ? p : 0);
    float q[HD], g[HD], dq[HD];
    float D = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      q[d] = live ? qkv[row * 3 * C + h * HD + d] : 0.f;
      g[d] = live ? dO[row * C + h * HD + d] : 0.f;
      D = fmaf(g[d], live ? O[row * C + h * HD + d] : 0.f, D);
      dq[d] = 0.f;
      Qs[t * (HD + 1) + d] = q[d];
      Gs[t * (HD + 1) + d] = g[d];
    }
    // pass 1: softmax statistics over the latent keys and the point's own key
    float s_self = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) s_self = fmaf(q[d], live ? qkv[row * 3 * C + C + h * HD + d] : 0.f, s_self);
    s_self *= scale;
    // one online-softmax sweep over the latent keys (running max + rescaled sum); K rows are read as broadcast float4
    float mx = s_self, sum = 1.0f;
    for (int j = 0; j < L; ++j) {
      const float4* kr = reinterpret_cast<const float4*>(Ks + j * HD);
      float s = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < HD / 4; ++d4) {
        const float4 k = kr[d4];
        s = fmaf(q[4 * d4], k.x, s); s = fmaf(q[4 * d4 + 1], k.y, s); s = fmaf(q[4 * d4 + 2], k.z, s); s = fmaf(q[4 * d4 + 3], k.w, s);
      }
      s *= scale;
      if (s > mx) { sum *= expf(mx - s); mx = s; }
      sum += expf(s - mx);
    }
    const float inv = 1.0f / sum;
    // the point's own key / value
    if (live) {
      const float ps = expf(s_self - mx) * inv;
      float dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) dp = fmaf(g[d], qkv[row * 3 * C + 2 * C + h * HD + d], dp);
      const float ds = ps * (dp - D) * scale;
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        const float ks = qkv[row * 3 * C + C + h * HD + d];
        dq[d] = fmaf(ds, ks, dq[d]);
        dqkv[row * 3 * C + C + h * HD + d] = ds * q[d];
        dqkv[row * 3 * C + 2 * C + h * HD + d] = ps * g[d];
      }
    }
    // pass 2: latent keys in blocks of 32
    for (int jb = 0; jb < L; jb += PAB_KB) {
      for (int jj = 0; jj < PAB_KB; ++jj) {
        const int j = jb + jj;
        float dsv = 0.f, pv = 0.f;
        if (j < L && live) {
          float s = 0.f, dp = 0.f;
          const float4* kr = reinterpret_cast<const float4*>(Ks + j * HD);
          const float4* vr = reinterpret_cast<const float4*>(Vs + j * HD);
          float4 kk[HD / 4];
#pragma unroll
          for (int d4 = 0; d4 < HD / 4; ++d4) {
            const float4 k = kr[d4], v = vr[d4];
            kk[d4] = k;
            s = fmaf(q[4 * d4], k.x, s); s = fmaf(q[4 * d4 + 1], k.y, s); s = fmaf(q[4 * d4 + 2], k.z, s); s = fmaf(q[4 * d4 + 3], k.w, s);
            dp = fmaf(g[4 * d4], v.x, dp); dp = fmaf(g[4 * d4 + 1], v.y, dp); dp = fmaf(g[4 * d4 + 2], v.z, dp); dp = fmaf(g[4 * d4 + 3], v.w, dp);
          }
          pv = expf(s * scale - mx) * inv;
          dsv = pv * (dp - D) * scale;
#pragma unroll
          for (int d4 = 0; d4 < HD / 4; ++d4) {
            dq[4 * d4] = fmaf(dsv, kk[d4].x, dq[4 * d4]); dq[4 * d4 + 1] = fmaf(dsv, kk[d4].y, dq[4 * d4 + 1]);
            dq[4 * d4 + 2] = fmaf(dsv, kk[d4].z, dq[4 * d4 + 2]); dq[4 * d4 + 3] = fmaf(dsv, kk[d4].w, dq[4 * d4 + 3]);
          }
        }
        dsT[jj * (PAB_NT + 1) + t] = dsv;
        pT[jj * (PAB_NT + 1) + t] = pv;
      }
      __syncthreads();
      {
        // thread -> key jj = t / TPK, dims d0 = (t % TPK) * DPT .. + DPT
        constexpr int TPK = PAB_NT / PAB_KB, DPT = HD / TPK;
        const int jj = t / TPK, d0 = (t % TPK) * DPT;
        const int j = jb + jj;
        if (j < L) {
          float ak[DPT] = {}, av[DPT] = {};
          for (int pp = 0; pp < PAB_NT; ++pp) {
            const float dsv = dsT[jj * (PAB_NT + 1) + pp], pv = pT[jj * (PAB_NT + 1) + pp];
#pragma unroll
            for (int i = 0; i < DPT; ++i) {
              ak[i] = fmaf(dsv, Qs[pp * (HD + 1) + d0 + i], ak[i]);
              av[i] = fmaf(pv, Gs[pp * (HD + 1) + d0 + i], av[i]);
            }
          }
#pragma unroll
          for (int i = 0; i < DPT; ++i) { dKs[j * HD + d0 + i] += ak[i]; dVs[j * HD + d0 + i] += av[i]; }
        }
      }
      __syncthreads();
    }
    if (live) {
#pragma unroll
      for (int d = 0; d < HD; ++d) dqkv[row * 3 * C + h * HD + d] = dq[d];
    }
    __syncthreads();
  }
  for (int i = t; i < L * HD; i += PAB_NT) {
    const int j = i / HD, d = i % HD;
    dk_lat[((int64_t)b * L + j) * ld_dlat + h * HD + d] = dKs[i];
    dv_lat[((int64_t)b * L + j) * ld_dlat + h * HD + d] = dVs[i];
  }
}

// ---- token self-attention backward (latent branch of the decoder, ViT blocks of the depth encoder; timm Attention) ----
// Two passes, no atomics.  Pass 1 (rows): one CTA per (image, head, row slab), K and V in shared memory, one warp per query
// row i: p_i = softmax(scale q_i K^T), dp_ij = dO_i . v_j, ds_ij = p_ij (dp_ij - sum_j p_ij dp_ij) scale, dq_i = ds_i K;
// the probability and ds rows go to a [B*heads, T, T] workspace.  Pass 2 (columns): dV = P^T dO, dK = dS^T Q, one CTA per
// (image, head, 32 key columns) with Q / dO of the head and the column block of P / dS in shared memory.
constexpr int MHA_BWD_WARPS = 16;
__global__ void __launch_bounds__(MHA_BWD_WARPS * 32) mha_bwd_rows_kernel(const float* __restrict__ qkv, const float* __restrict__ dO,
                                                                          float* __restrict__ dqkv, float* __restrict__ Pg,
                                                                          float* __restrict__ Sg, int T, int heads, int hd, float scale,
                                                                          int rows_per_cta) {
  extern __shared__ __align__(16) float sm[];
  const int ld = hd + 4;                // float4 rows; lanes 68 floats apart hit distinct bank quads
  float* Ks = sm;                       // [T][ld]
  float* Vs = Ks + (size_t)T * ld;
  float* Ps = Vs + (size_t)T * ld;      // [warps][T]   p_j
  float* Ds = Ps + (size_t)MHA_BWD_WARPS * T;   // [warps][T]   dp_j, then ds_j
  float* Qw = Ds + (size_t)MHA_BWD_WARPS * T;   // [warps][hd]
  float* Gw = Qw + (size_t)MHA_BWD_WARPS * hd;  // [warps][hd]
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int C = heads * hd;
  const float* base = qkv + (int64_t)b * T * 3 * C;
  for (int i = threadIdx.x; i < T * (hd >> 2); i += blockDim.x) {
    const int t = i / (hd >> 2), d = (i % (hd >> 2)) * 4;
    *reinterpret_cast<float4*>(Ks + t * ld + d) = __ldg(reinterpret_cast<const float4*>(base + (int64_t)t * 3 * C + C + h * hd + d));
    *reinterpret_cast<float4*>(Vs + t * ld + d) = __ldg(reinterpret_cast<const float4*>(base + (int64_t)t * 3 * C + 2 * C + h * hd + d));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* P = Ps + warp * T;
  float* Dv = Ds + warp * T;
  float* Q = Qw + warp * hd;
  float* G = Gw + warp * hd;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(T, r0 + rows_per_cta);
  float* Pout = Pg + (int64_t)bh * T * T;
  float* Sout = Sg + (int64_t)bh * T * T;
  for (int i = r0 + warp; i < r1; i += MHA_BWD_WARPS) {
    for (int d = lane; d < hd; d += 32) {
      Q[d] = base[(int64_t)i * 3 * C + h * hd + d];
      G[d] = dO[((int64_t)b * T + i) * C + h * hd + d];
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) {
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < hd; d += 4) {
        const float4 q = *reinterpret_cast<const float4*>(Q + d), g = *reinterpret_cast<const float4*>(G + d);
        const float4 k = *reinterpret_cast<const float4*>(Ks + j * ld + d), v = *reinterpret_cast<const float4*>(Vs + j * ld + d);
        s = fmaf(q.x, k.x, s); s = fmaf(q.y, k.y, s); s = fmaf(q.z, k.z, s); s = fmaf(q.w, k.w, s);
        dp = fmaf(g.x, v.x, dp); dp = fmaf(g.y, v.y, dp); dp = fmaf(g.z, v.z, dp); dp = fmaf(g.w, v.w, dp);
      }
      s *= scale;
      P[j] = s;
      Dv[j] = dp;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) { const float e = expf(P[j] - mx); P[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float Dp = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float p = P[j] * inv;
      P[j] = p;
      Dp = fmaf(p, Dv[j], Dp);
    }
    Dp = warp_sum(Dp);
    for (int j = lane; j < T; j += 32) {
      const float ds = P[j] * (Dv[j] - Dp) * scale;
      Dv[j] = ds;
      Pout[(int64_t)i * T + j] = P[j];
      Sout[(int64_t)i * T + j] = ds;
    }
    __syncwarp();
    for (int d = lane; d < hd; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < T; ++j) acc = fmaf(Dv[j], Ks[j * ld + d], acc);
      dqkv[((int64_t)b * T + i) * 3 * C + h * hd + d] = acc;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) mha_bwd_cols_kernel(const float* __restrict__ qkv, const float* __restrict__ dO,
                                                           float* __restrict__ dqkv, const float* __restrict__ Pg,
                                                           const float* __restrict__ Sg, int T, int heads, int hd) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                        // [T][hd]
  float* Gs = Qs + (size_t)T * hd;       // [T][hd]
  float* Pc = Gs + (size_t)T * hd;       // [T][32]  column block of P
  float* Sc = Pc + (size_t)T * 32;       // [T][32]  column block of dS
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int C = heads * hd;
  const int j0 = blockIdx.y * 32;
  const float* base = qkv + (int64_t)b * T * 3 * C;
  for (int i = threadIdx.x; i < T * (hd >> 2); i += blockDim.x) {
    const int t = i / (hd >> 2), d = (i % (hd >> 2)) * 4;
    *reinterpret_cast<float4*>(Qs + t * hd + d) = __ldg(reinterpret_cast<const float4*>(base + (int64_t)t * 3 * C + h * hd + d));
    *reinterpret_cast<float4*>(Gs + t * hd + d) = __ldg(reinterpret_cast<const float4*>(dO + ((int64_t)b * T + t) * C + h * hd + d));
  }
  const float* Pin = Pg + (int64_t)bh * T * T;
  const float* Sin = Sg + (int64_t)bh * T * T;
  for (int i = threadIdx.x; i < T * 32; i += blockDim.x) {
    const int t = i >> 5, c = i & 31;
    const bool ok = j0 + c < T;
    Pc[i] = ok ? Pin[(int64_t)t * T + j0 + c] : 0.f;
    Sc[i] = ok ? Sin[(int64_t)t * T + j0 + c] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool two = hd > 32;              // hd <= 64: lane owns d = lane and lane + 32
  float av[4][2] = {}, ak[4][2] = {};
  for (int i = 0; i < T; ++i) {
    const float4 pv = *reinterpret_cast<const float4*>(Pc + i * 32 + warp * 4);
    const float4 sv = *reinterpret_cast<const float4*>(Sc + i * 32 + warp * 4);
    const float g0 = lane < hd ? Gs[i * hd + lane] : 0.f, q0 = lane < hd ? Qs[i * hd + lane] : 0.f;
    const float g1 = two ? Gs[i * hd + lane + 32] : 0.f, q1 = two ? Qs[i * hd + lane + 32] : 0.f;
    const float pp[4] = {pv.x, pv.y, pv.z, pv.w}, ss[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      av[c][0] = fmaf(pp[c], g0, av[c][0]); av[c][1] = fmaf(pp[c], g1, av[c][1]);
      ak[c][0] = fmaf(ss[c], q0, ak[c][0]); ak[c][1] = fmaf(ss[c], q1, ak[c][1]);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int j = j0 + warp * 4 + c;
    if (j >= T) continue;
    float* row = dqkv + ((int64_t)b * T + j) * 3 * C + h * hd;
    if (lane < hd) { row[C + lane] = ak[c][0]; row[2 * C + lane] = av[c][0]; }
    if (two) { row[C + lane + 32] = ak[c][1]; row[2 * C + lane + 32] = av[c][1]; }
  }
}

// ---- BatchNorm in training mode (torchvision ResNet-50 / Bottleneck_Conv of CoordEncRes, seen_coord_enc.py:145-178) ----
// statistics over the rows of x [M, C] (NHWC: M = B*H*W): double accumulators, atomics across row slabs
__global__ void bn_stats_kernel(const float* __restrict__ x, int64_t M, int C, double* __restrict__ acc) {   // acc[0:C] sum, [C:2C] sumsq
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0, q = 0.0;
  if (c < C)
    for (int64_t m = blockIdx.y * 8 + threadIdx.y; m < M; m += (int64_t)gridDim.y * 8) { const double v = x[m * C + c]; s += v; q += v * v; }
  __shared__ double rs[8][33], rq[8][33];
  rs[threadIdx.y][threadIdx.x] = s; rq[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a += rs[i][threadIdx.x]; b += rq[i][threadIdx.x]; }
    atomicAdd(acc + c, a);
    atomicAdd(acc + C + c, b);
  }
}
__global__ void bn_stats_finish_kernel(const double* __restrict__ acc, int64_t M, int C, float eps, float* __restrict__ mean,
                                       float* __restrict__ var, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = acc[c] / (double)M;
  double v = acc[C + c] / (double)M - mu * mu;
  if (v < 0.0) v = 0.0;
  mean[c] = (float)mu; var[c] = (float)v; rstd[c] = (float)(1.0 / sqrt(v + (double)eps));
}
// dgamma[c] = sum_m dy * xhat, dbeta[c] = sum_m dy
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, int64_t M, int C, double* __restrict__ acc) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0, q = 0.0;
  if (c < C) {
    const float mu = mean[c], rs = rstd[c];
    for (int64_t m = blockIdx.y * 8 + threadIdx.y; m < M; m += (int64_t)gridDim.y * 8) {
      const float d = dy[m * C + c];
      s += (double)d; q += (double)(d * ((x[m * C + c] - mu) * rs));
    }
  }
  __shared__ double rs_[8][33], rq_[8][33];
  rs_[threadIdx.y][threadIdx.x] = s; rq_[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a += rs_[i][threadIdx.x]; b += rq_[i][threadIdx.x]; }
    atomicAdd(acc + c, a);          // dbeta
    atomicAdd(acc + C + c, b);      // dgamma
  }
}
// dx = gamma * rstd * (dy - dbeta / M - xhat * dgamma / M);  also converts the double accumulators to the fp32 gradients
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ gamma, const double* __restrict__ acc,
                                    int64_t M, int C, float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int64_t n = M * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float db = (float)(acc[c] / (double)M), dg = (float)(acc[C + c] / (double)M);
    const float xh = (x[i] - mean[c]) * rstd[c];
    dx[i] = gamma[c] * rstd[c] * (dy[i] - db - xh * dg);
    if (i < C) { dbeta[i] += (float)acc[i]; dgamma[i] += (float)acc[C + i]; }
  }
}

// ---- 3x3 stride-2 max-pool backward (NHWC): the gradient goes to the first maximum in scan order, like PyTorch ----
__global__ void maxpool3x3s2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int B, int H,
                                        int W, int C, int pad_top, int pad_left, int OH, int OW) {
  const int64_t n = (int64_t)B * OH * OW * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t t = i / C;
    const int ow = (int)(t % OW); t /= OW;
    const int oh = (int)(t % OH);
    const int b = (int)(t / OH);
    float best = -INFINITY;
    int64_t arg = -1;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const int ih = oh * 2 - pad_top + kh, iw = ow * 2 - pad_left + kw;
        if ((unsigned)ih >= (unsigned)H || (unsigned)iw >= (unsigned)W) continue;
        const int64_t j = (((int64_t)b * H + ih) * W + iw) * C + c;
        const float v = x[j];
        if (v > best || arg < 0) { if (v > best || arg < 0) { best = v; arg = j; } }
      }
    if (arg >= 0) atomicAdd(dx + arg, dy[i]);
  }
}

// dx[b, hw, c] = dy[b, c] / HW   (global average pool backward)
__global__ void avgpool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int HW, int C) {
  const int64_t n = (int64_t)B * HW * C;
  const float inv = 1.0f / (float)HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int b = (int)(i / ((int64_t)HW * C));
    dx[i] = dy[(int64_t)b * C + c] * inv;
  }
}

// ---- LayerNorm backward, any C (multiple of 4): dx only (dgamma = coldot(dy, xhat), dbeta = colsum(dy)) -----------
__global__ void __launch_bounds__(256) layernorm_bwd_generic_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                    const float* __restrict__ gamma, float eps, float* __restrict__ dx,
                                                                    float* __restrict__ xhat, int64_t rows, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = blockIdx.x * 8 + warp; r < rows; r += (int64_t)gridDim.x * 8) {
    const float* xr = x + r * C;
    const float* dr = dy + r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float g = dr[c] * (gamma ? gamma[c] : 1.f);
      s1 += g; s2 += g * (xr[c] - mean) * rstd;
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
    for (int c = lane; c < C; c += 32) {
      const float xh = (xr[c] - mean) * rstd;
      dx[r * C + c] = rstd * (dr[c] * (gamma ? gamma[c] : 1.f) - s1 - xh * s2);
      if (xhat) xhat[r * C + c] = xh;
    }
  }
}

// ---- GroupNorm backward (NHWC; timm GroupNormAct of the ResNetV2 backbone): one CTA per (image, group) -----------
__global__ void __launch_bounds__(512) groupnorm_bwd_nhwc_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                 const float* __restrict__ gamma, float* __restrict__ dx,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta, int HW, int C,
                                                                 int groups, float eps) {
  const int b = blockIdx.x / groups, gi = blockIdx.x % groups;
  const int cg = C / groups;
  const int64_t base = (int64_t)b * HW * C + gi * cg;
  const int64_t n = (int64_t)HW * cg;
  __shared__ float sh[34];
  __shared__ float accg[64], accb[64];      // cg <= 64
  auto bsum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
      t = warp_sum(t);
      if (threadIdx.x == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
  };
  if (threadIdx.x < 64) { accg[threadIdx.x] = 0.f; accb[threadIdx.x] = 0.f; }
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[base + (i / cg) * C + (i % cg)];
  const float mean = bsum(s) / (float)n;
  float v = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const float d = x[base + (i / cg) * C + (i % cg)] - mean; v += d * d; }
  const float rstd = rsqrtf(bsum(v) / (float)n + eps);
  float s1 = 0.f, s2 = 0.f, lg = 0.f, lb = 0.f;      // blockDim % cg == 0: a thread always sees the same channel
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = (int)(i % cg);
    const int64_t off = base + (i / cg) * C + c;
    const float xh = (x[off] - mean) * rstd, d = dy[off];
    const float g = d * gamma[gi * cg + c];
    s1 += g; s2 += g * xh;
    lg += d * xh; lb += d;
  }
  atomicAdd(&accg[threadIdx.x % cg], lg);
  atomicAdd(&accb[threadIdx.x % cg], lb);
  s1 = bsum(s1) / (float)n;
  s2 = bsum(s2) / (float)n;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = (int)(i % cg);
    const int64_t off = base + (i / cg) * C + c;
    const float xh = (x[off] - mean) * rstd;
    dx[off] = rstd * (dy[off] * gamma[gi * cg + c] - s1 - xh * s2);
  }
  __syncthreads();
  if (threadIdx.x < cg) { atomicAdd(dgamma + gi * cg + threadIdx.x, accg[threadIdx.x]); atomicAdd(dbeta + gi * cg + threadIdx.x, accb[threadIdx.x]); }
}

// ---- bilinear resize backward (NHWC): scatter with atomics ------------------------------------------------------
__device__ __forceinline__ void bilinear_src_t(int dst, int in, int out, int align, int& i0, int& i1, float& l1) {
  float src;
  if (align) {
    const float sc = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    src = sc * dst;
  } else {
    const float sc = (float)in / (float)out;
    src = sc * (dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
  }
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 < in - 1 ? i0 + 1 : i0;
  l1 = src - (float)i0;
}
__global__ void bilinear_bwd_nhwc_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C, int OH, int OW,
                                         int align) {
  const int64_t total = (int64_t)B * OH * OW * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t t = i / C;
    const int ow = (int)(t % OW); t /= OW;
    const int oh = (int)(t % OH);
    const int b = (int)(t / OH);
    int h0, h1, w0, w1;
    float lh, lw;
    bilinear_src_t(oh, H, OH, align, h0, h1, lh);
    bilinear_src_t(ow, W, OW, align, w0, w1, lw);
    float* xb = dx + (int64_t)b * H * W * C + c;
    const float g = dy[i], hh0 = 1.f - lh, ww0 = 1.f - lw;
    atomicAdd(xb + ((int64_t)h0 * W + w0) * C, g * hh0 * ww0);
    atomicAdd(xb + ((int64_t)h0 * W + w1) * C, g * hh0 * lw);
    atomicAdd(xb + ((int64_t)h1 * W + w0) * C, g * lh * ww0);
    atomicAdd(xb + ((int64_t)h1 * W + w1) * C, g * lh * lw);
  }
}

// The same gradient as a GATHER, four channels per thread: every input pixel (h, w) sums the output pixels whose two taps include
// it, found by running the forward's own source-index function over the few candidate rows / columns (an input row h is touched
// by the outputs whose source coordinate lies in (h - 1, h + 1): ~2 x the scale factor of them).  No zero fill, no atomics,
// deterministic; 128-bit loads of dy that stay in L1 / L2 across neighbouring pixels.
__device__ __forceinline__ void bilinear_dst_range(int i, int in, int out, int align, int& lo, int& hi) {
  if (in == 1 || out == 1) { lo = 0; hi = out - 1; return; }
  float a, b;                                             // destination coordinates of the sources i - 1 and i + 1
  if (align) {
    const float inv = (float)(out - 1) / (float)(in - 1);
    a = (i - 1) * inv; b = (i + 1) * inv;
  } else {
    const float inv = (float)out / (float)in;
    a = (i - 0.5f) * inv - 0.5f; b = (i + 1.5f) * inv - 0.5f;
  }
  lo = (int)floorf(a) - 1; hi = (int)ceilf(b) + 1;        // one spare candidate on either side: its weight comes out as 0
  if (i == 0 || lo < 0) lo = 0;                           // (sources clamped to 0 / in - 1 land on the border rows)
  if (i == in - 1 || hi > out - 1) hi = out - 1;
}

constexpr int BG_MAXT = 10;                                 // candidate rows / columns per input pixel kept in registers
__global__ void bilinear_bwd_gather_v4_kernel(const float4* __restrict__ dy, float4* __restrict__ dx, int B, int H, int W, int C4,
                                              int OH, int OW, int align) {
  const unsigned total = (unsigned)B * H * W * C4;          // < 2^31 (checked by the caller)
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned pix = i / (unsigned)C4, c = i - pix * C4;
    const unsigned t = pix / (unsigned)W;
    const int w = (int)(pix - t * W);
    const unsigned b = t / (unsigned)H;
    const int h = (int)(t - b * H);
    int oh_lo, oh_hi, ow_lo, ow_hi;
    bilinear_dst_range(h, H, OH, align, oh_lo, oh_hi);
    bilinear_dst_range(w, W, OW, align, ow_lo, ow_hi);
    // column weights once per pixel (not once per candidate row)
    float ww[BG_MAXT];
    const int ncol = ow_hi - ow_lo + 1;
    const bool cached = ncol <= BG_MAXT;
    if (cached) {
#pragma unroll
      for (int k = 0; k < BG_MAXT; ++k) {
        ww[k] = 0.f;
        if (k < ncol) {
          int w0, w1;
          float lw;
          bilinear_src_t(ow_lo + k, W, OW, align, w0, w1, lw);
          ww[k] = (w0 == w ? 1.f - lw : 0.f) + (w1 == w ? lw : 0.f);
        }
      }
    }
    const float4* g = dy + (size_t)b * OH * OW * C4 + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      int h0, h1;
      float lh;
      bilinear_src_t(oh, H, OH, align, h0, h1, lh);
      const float wh = (h0 == h ? 1.f - lh : 0.f) + (h1 == h ? lh : 0.f);
      if (wh == 0.f) continue;
      const float4* grow = g + (size_t)oh * OW * C4;
      if (cached) {
#pragma unroll
        for (int k = 0; k < BG_MAXT; ++k) {
          if (k < ncol && ww[k] != 0.f) {
            const float4 v = __ldg(grow + (ow_lo + k) * C4);
            const float kk = wh * ww[k];
            acc.x = fmaf(v.x, kk, acc.x); acc.y = fmaf(v.y, kk, acc.y); acc.z = fmaf(v.z, kk, acc.z); acc.w = fmaf(v.w, kk, acc.w);
          }
        }
      } else {
        for (int ow = ow_lo; ow <= ow_hi; ++ow) {
          int w0, w1;
          float lw;
          bilinear_src_t(ow, W, OW, align, w0, w1, lw);
          const float wc = (w0 == w ? 1.f - lw : 0.f) + (w1 == w ? lw : 0.f);
          if (wc == 0.f) continue;
          const float4 v = __ldg(grow + ow * C4);
          const float kk = wh * wc;
          acc.x = fmaf(v.x, kk, acc.x); acc.y = fmaf(v.y, kk, acc.y); acc.z = fmaf(v.z, kk, acc.z); acc.w = fmaf(v.w, kk, acc.w);
        }
      }
    }
    dx[i] = acc;
  }
}

// ---- unproject + normalise backward (utils/camera.py:52-108, graph_shape.py:132-141) ------------------------------
// out_i = (X_i - m) / s for valid pixels, X_i = d_i * r_i, r_i = Kinv (u, v, 1)^T, m = mean_valid X, s = max_valid |X - m|.
// Given g_i = dL/dout_i:  dL/dX_i = g_i / s + [i == j] * ds * n_j + dm / N,   ds = -sum_i g_i . out_i / s,  n_j = out_j,
// dm = -sum_i g_i / s - ds * n_j (j = the pixel that attains the maximum);  dL/dd_i = dL/dX_i . r_i,
// dL/dKinv = sum_i d_i * (dL/dX_i) (u, v, 1).   One CTA (1024 threads) per image.
__global__ void __launch_bounds__(1024) unproject_normalize_bwd_kernel(const float* __restrict__ depth, const float* __restrict__ mask,
                                                                       const float* __restrict__ K, const float* __restrict__ out,
                                                                       const float* __restrict__ scale, const float* __restrict__ g,
                                                                       float* __restrict__ ddepth, float* __restrict__ dKinv, int H, int W) {
  __shared__ float sh[33];
  __shared__ float kinv[9];
  __shared__ int sj;
  const int b = blockIdx.x, HW = H * W;
  if (threadIdx.x == 0) {
    const float* k = K + b * 9;
    const float a = k[0], bb = k[1], c = k[2], d = k[3], e = k[4], f = k[5], gg = k[6], h = k[7], i = k[8];
    const float A = e * i - f * h, Bc = -(d * i - f * gg), Cc = d * h - e * gg;
    const float r = 1.0f / (a * A + bb * Bc + c * Cc);
    kinv[0] = A * r;  kinv[1] = -(bb * i - c * h) * r; kinv[2] = (bb * f - c * e) * r;
    kinv[3] = Bc * r; kinv[4] = (a * i - c * gg) * r;  kinv[5] = -(a * f - c * d) * r;
    kinv[6] = Cc * r; kinv[7] = -(a * h - bb * gg) * r; kinv[8] = (a * e - bb * d) * r;
    sj = 0x7fffffff;
  }
  __syncthreads();
  auto bred = [&](float v, bool is_max) {
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : (is_max ? -INFINITY : 0.f);
      t = is_max ? warp_max(t) : warp_sum(t);
      if (threadIdx.x == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
  };
  const float* dp = depth + (int64_t)b * HW;
  const float* mp = mask + (int64_t)b * HW;
  const float* op = out + (int64_t)b * HW * 3;
  const float* gp = g + (int64_t)b * HW * 3;
  const float s = scale[b];
  float gx = 0.f, gy = 0.f, gz = 0.f, go = 0.f, cnt = 0.f, nmax = -INFINITY;
  for (int idx = threadIdx.x; idx < HW; idx += blockDim.x) {
    if (mp[idx] > 0.5f) {
      const float ox = op[idx * 3], oy = op[idx * 3 + 1], oz = op[idx * 3 + 2];
      const float a0 = gp[idx * 3], a1 = gp[idx * 3 + 1], a2 = gp[idx * 3 + 2];
      gx += a0; gy += a1; gz += a2; go += a0 * ox + a1 * oy + a2 * oz; cnt += 1.f;
      nmax = fmaxf(nmax, ox * ox + oy * oy + oz * oz);
    }
  }
  gx = bred(gx, false); gy = bred(gy, false); gz = bred(gz, false); go = bred(go, false); cnt = bred(cnt, false);
  nmax = bred(nmax, true);
  for (int idx = threadIdx.x; idx < HW; idx += blockDim.x) {     // first pixel attaining the maximum norm
    if (mp[idx] > 0.5f) {
      const float ox = op[idx * 3], oy = op[idx * 3 + 1], oz = op[idx * 3 + 2];
      if (ox * ox + oy * oy + oz * oz == nmax) atomicMin(&sj, idx);
    }
  }
  __syncthreads();
  const int j = sj;
  const float ds = -go / s;
  float njx = 0.f, njy = 0.f, njz = 0.f;
  if (j < HW) {
    njx = op[j * 3]; njy = op[j * 3 + 1]; njz = op[j * 3 + 2];
    const float nn = sqrtf(njx * njx + njy * njy + njz * njz);
    if (nn > 0.f) { njx /= nn; njy /= nn; njz /= nn; }
  }
  const float dmx = (-gx / s - ds * njx) / cnt, dmy = (-gy / s - ds * njy) / cnt, dmz = (-gz / s - ds * njz) / cnt;
  float m[9] = {};
  for (int idx = threadIdx.x; idx < HW; idx += blockDim.x) {
    float dd = 0.f;
    if (mp[idx] > 0.5f) {
      float dxv = gp[idx * 3] / s + dmx, dyv = gp[idx * 3 + 1] / s + dmy, dzv = gp[idx * 3 + 2] / s + dmz;
      if (idx == j) { dxv += ds * njx; dyv += ds * njy; dzv += ds * njz; }
      const float px = (float)(idx % W), py = (float)(idx / W);
      const float rx = kinv[0] * px + kinv[1] * py + kinv[2], ry = kinv[3] * px + kinv[4] * py + kinv[5],
                  rz = kinv[6] * px + kinv[7] * py + kinv[8];
      dd = dxv * rx + dyv * ry + dzv * rz;
      const float dz = dp[idx];
      m[0] += dz * dxv * px; m[1] += dz * dxv * py; m[2] += dz * dxv;
      m[3] += dz * dyv * px; m[4] += dz * dyv * py; m[5] += dz * dyv;
      m[6] += dz * dzv * px; m[7] += dz * dzv * py; m[8] += dz * dzv;
    }
    ddepth[(int64_t)b * HW + idx] = dd;
  }
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    const float t = bred(m[q], false);
    if (threadIdx.x == 0) dKinv[b * 9 + q] = t;
  }
}

// ---- AdamW (torch.optim.AdamW semantics, model/shape_engine.py:132) ---------------------------------------------
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             int64_t n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i] * (1.0f - lr * wd);
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// all parameter tensors of the optimizer in ONE launch: `table` holds, per tensor, {param, grad, exp_avg, exp_avg_sq, numel}
// as five 64-bit words; blockIdx.y = tensor, the CTAs of a row stride over its elements
__global__ void adamw_multi_kernel(const unsigned long long* __restrict__ table, float lr, float b1, float b2, float eps, float wd,
                                   float bc1, float bc2_sqrt) {
  const unsigned long long* e = table + (size_t)blockIdx.y * 5;
  float* __restrict__ p = reinterpret_cast<float*>(e[0]);
  const float* __restrict__ g = reinterpret_cast<const float*>(e[1]);
  float* __restrict__ m = reinterpret_cast<float*>(e[2]);
  float* __restrict__ v = reinterpret_cast<float*>(e[3]);
  const int64_t n = (int64_t)e[4];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i] * (1.0f - lr * wd);
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// the same update with the step-dependent scalars read from device memory: hyper = {lr, beta1, beta2, eps, weight_decay,
// 1 - beta1^step, sqrt(1 - beta2^step)}, and the tensor table passed BY VALUE in the kernel parameters.  Nothing that changes
// from step to step is a kernel argument and no host buffer is read at run time, so the launches can sit in a CUDA graph that
// is replayed every step while the host refreshes `hyper` with a stream-ordered copy in front of each replay.
constexpr int ADAMW_ROWS = 96;                       // 96 x 40 B = 3840 B of kernel parameters
struct AdamwRows { unsigned long long w[ADAMW_ROWS * 5]; };

__global__ void adamw_multi_dev_kernel(const __grid_constant__ AdamwRows rows, const float* __restrict__ hyper) {
  const unsigned long long* e = rows.w + (size_t)blockIdx.y * 5;
  float* __restrict__ p = reinterpret_cast<float*>(e[0]);
  const float* __restrict__ g = reinterpret_cast<const float*>(e[1]);
  float* __restrict__ m = reinterpret_cast<float*>(e[2]);
  float* __restrict__ v = reinterpret_cast<float*>(e[3]);
  const int64_t n = (int64_t)e[4];
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], bc1 = hyper[5], bc2_sqrt = hyper[6];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i] * (1.0f - lr * wd);
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

}  // namespace zs

using namespace zs;

extern "C" int zs_adamw_multi_dev_f32(const void* host_table, int n_tensors, const float* hyper, void* stream) {
  ZS_REQUIRE(host_table && hyper && n_tensors >= 0, "zs_adamw_multi_dev_f32: bad args");
  const unsigned long long* t = reinterpret_cast<const unsigned long long*>(host_table);
  for (int first = 0; first < n_tensors; first += ADAMW_ROWS) {
    const int n = n_tensors - first < ADAMW_ROWS ? n_tensors - first : ADAMW_ROWS;
    AdamwRows rows;
    memcpy(rows.w, t + (size_t)first * 5, sizeof(unsigned long long) * 5 * n);
    adamw_multi_dev_kernel<<<dim3(48, n), 256, 0, as_stream(stream)>>>(rows, hyper);
    ZS_CUDA_CHECK_LAUNCH("zs_adamw_multi_dev_f32");
  }
  return ZS_OK;
}

extern "C" int zs_adamw_multi_f32(const void* table, int n_tensors, float lr, float beta1, float beta2, float eps, float weight_decay,
                                  int step, void* stream) {
  ZS_REQUIRE(table && n_tensors >= 0 && n_tensors <= 65535 && step >= 1, "zs_adamw_multi_f32: bad args");
  if (n_tensors == 0) return ZS_OK;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2 = sqrtf(1.0f - powf(beta2, (float)step));
  adamw_multi_kernel<<<dim3(48, n_tensors), 256, 0, as_stream(stream)>>>(reinterpret_cast<const unsigned long long*>(table), lr, beta1,
                                                                         beta2, eps, weight_decay, bc1, bc2);
  ZS_CUDA_CHECK_LAUNCH("zs_adamw_multi_f32");
  return ZS_OK;
}

extern "C" int zs_bce_logits_fwd(const float* logits, const float* sdf, int64_t n, float impt_thres, float impt_weight,
                                 double* ws, float* loss, void* stream) {
  ZS_REQUIRE(logits && sdf && ws && loss && n > 0, "zs_bce_logits_fwd: bad args");
  cudaStream_t st = as_stream(stream);
  ZS_CUDA_CALL(cudaMemsetAsync(ws, 0, sizeof(double), st));
  bce_fwd_kernel<<<grid_for_n(n), 256, 0, st>>>(logits, sdf, n, impt_thres, impt_weight, ws);
  ZS_CUDA_CHECK_LAUNCH("zs_bce_logits_fwd");
  bce_finish_kernel<<<1, 1, 0, st>>>(ws, n, loss);
  ZS_CUDA_CHECK_LAUNCH("zs_bce_logits_fwd(finish)");
  return ZS_OK;
}

extern "C" int zs_bce_logits_bwd(const float* logits, const float* sdf, int64_t n, float impt_thres, float impt_weight,
                                 float grad_scale, float* dlogits, void* stream) {
  ZS_REQUIRE(logits && sdf && dlogits && n > 0, "zs_bce_logits_bwd: bad args");
  bce_bwd_kernel<<<grid_for_n(n), 256, 0, as_stream(stream)>>>(logits, sdf, n, impt_thres, impt_weight, grad_scale, dlogits);
  ZS_CUDA_CHECK_LAUNCH("zs_bce_logits_bwd");
  return ZS_OK;
}

extern "C" int zs_act_bwd_f32(const float* dy, const float* z, float* dx, int64_t n, int act, void* stream) {
  ZS_REQUIRE(dy && z && dx && n >= 0, "zs_act_bwd_f32: bad args");
  if (n == 0) return ZS_OK;
  act_bwd_kernel<<<grid_for_n(n), 256, 0, as_stream(stream)>>>(dy, z, dx, n, act);
  ZS_CUDA_CHECK_LAUNCH("zs_act_bwd_f32");
  return ZS_OK;
}

extern "C" int zs_colsum_f32(const float* A, int lda, int64_t M, int N, float* out, int accumulate, void* stream) {
  ZS_REQUIRE(A && out && M >= 0 && N > 0 && lda >= N, "zs_colsum_f32: bad args");
  cudaStream_t st = as_stream(stream);
  if (!accumulate) ZS_CUDA_CALL(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
  if (M == 0) return ZS_OK;
  int64_t slabs = (M + 8 * 64 - 1) / (8 * 64);
  if (slabs > 256) slabs = 256;
  dim3 grid((N + 31) / 32, (unsigned)(slabs > 0 ? slabs : 1));
  colsum_kernel<<<grid, dim3(32, 8), 0, st>>>(A, lda, nullptr, 0, M, N, out);
  ZS_CUDA_CHECK_LAUNCH("zs_colsum_f32");
  return ZS_OK;
}

extern "C" int zs_coldot_f32(const float* A, int lda, const float* B2, int ldb, int64_t M, int N, float* out, int accumulate,
                             void* stream) {
  ZS_REQUIRE(A && B2 && out && M >= 0 && N > 0 && lda >= N && ldb >= N, "zs_coldot_f32: bad args");
  cudaStream_t st = as_stream(stream);
  if (!accumulate) ZS_CUDA_CALL(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
  if (M == 0) return ZS_OK;
  int64_t slabs = (M + 8 * 64 - 1) / (8 * 64);
  if (slabs > 256) slabs = 256;
  dim3 grid((N + 31) / 32, (unsigned)(slabs > 0 ? slabs : 1));
  colsum_kernel<<<grid, dim3(32, 8), 0, st>>>(A, lda, B2, ldb, M, N, out);
  ZS_CUDA_CHECK_LAUNCH("zs_coldot_f32");
  return ZS_OK;
}

extern "C" int zs_gemm_tn_f32(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int64_t M, int N, int K,
                              int accumulate, void* stream) {
  ZS_REQUIRE(A && B && C && M >= 0 && N > 0 && K > 0 && lda >= N && ldb >= K && ldc >= K, "zs_gemm_tn_f32: bad args");
  cudaStream_t st = as_stream(stream);
  if (!accumulate) ZS_CUDA_CALL(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * K, N, st));
  if (M == 0) return ZS_OK;
  const int tiles = ((N + 63) / 64) * ((K + 63) / 64);
  int64_t splits = (2 * (int64_t)sm_count() + tiles - 1) / tiles;
  const int64_t max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t mps = ((M + splits - 1) / splits + 15) / 16 * 16;
  splits = (M + mps - 1) / mps;
  dim3 grid((N + 63) / 64, (K + 63) / 64, (unsigned)splits);
  gemm_tn_kernel<<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, mps);
  ZS_CUDA_CHECK_LAUNCH("zs_gemm_tn_f32");
  return ZS_OK;
}

extern "C" int zs_layernorm_bwd_f32(const float* dy, const float* x, const float* gamma, float eps, float* dx, float* dgamma,
                                    float* dbeta, int64_t rows, int cols, void* stream) {
  ZS_REQUIRE(dy && x && dx && rows >= 0, "zs_layernorm_bwd_f32: null pointer");
  ZS_REQUIRE(cols == 256, "zs_layernorm_bwd_f32: only 256 columns (the decoder width) are supported");
  ZS_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "zs_layernorm_bwd_f32: dgamma and dbeta go together");
  if (rows == 0) return ZS_OK;
  int64_t g = (rows + 7) / 8;
  const int cap = sm_count() * 8;
  layernorm_bwd_kernel<<<(int)(g < cap ? g : cap), 256, 0, as_stream(stream)>>>(dy, x, gamma, eps, dx, dgamma, dbeta, rows);
  ZS_CUDA_CHECK_LAUNCH("zs_layernorm_bwd_f32");
  return ZS_OK;
}

extern "C" int zs_point_attention_bwd_f32(const float* qkv_p, const float* k_lat, const float* v_lat, int ld_lat, const float* O,
                                          const float* dO, float* dqkv_p, float* dk_lat, float* dv_lat, int ld_dlat, int B, int P,
                                          int L, int heads, int hd, float scale, void* stream) {
  ZS_REQUIRE(qkv_p && k_lat && v_lat && O && dO && dqkv_p && dk_lat && dv_lat, "zs_point_attention_bwd_f32: null pointer");
  ZS_REQUIRE(hd == 32 && B > 0 && P > 0 && L > 0 && heads > 0, "zs_point_attention_bwd_f32: head dim must be 32");
  const size_t smem = sizeof(float) * ((size_t)4 * L * 32 + 2 * PAB_NT * 33 + 2 * PAB_KB * (PAB_NT + 1));
  ZS_REQUIRE(smem <= 227 * 1024, "zs_point_attention_bwd_f32: too many latent tokens for shared memory");
  ZS_CUDA_CALL(cudaFuncSetAttribute(point_attention_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  point_attention_bwd_kernel<32><<<dim3(heads, B), PAB_NT, smem, as_stream(stream)>>>(qkv_p, k_lat, v_lat, ld_lat, O, dO, dqkv_p, dk_lat,
                                                                                  dv_lat, ld_dlat, P, L, heads, scale);
  ZS_CUDA_CHECK_LAUNCH("zs_point_attention_bwd_f32");
  return ZS_OK;
}

extern "C" size_t zs_mha_bwd_ws_bytes(int B, int T, int heads) {
  if (B <= 0 || T <= 0 || heads <= 0) return 0;
  return sizeof(float) * 2 * (size_t)B * heads * T * T;
}

extern "C" int zs_mha_bwd_f32(const float* qkv, const float* dO, float* dqkv, int B, int T, int heads, int hd, float scale, void* ws,
                              void* stream) {
  ZS_REQUIRE(qkv && dO && dqkv && ws && B > 0 && T > 0 && heads > 0 && hd > 0, "zs_mha_bwd_f32: bad args");
  ZS_REQUIRE((hd & 3) == 0 && hd <= 64, "zs_mha_bwd_f32: head dim must be a multiple of 4, at most 64");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dO) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(ws) & 15) == 0, "zs_mha_bwd_f32: qkv / dO / ws must be 16-byte aligned");
  float* Pg = reinterpret_cast<float*>(ws);
  float* Sg = Pg + (size_t)B * heads * T * T;
  const size_t smem1 = sizeof(float) * ((size_t)2 * T * (hd + 4) + (size_t)2 * MHA_BWD_WARPS * T + (size_t)2 * MHA_BWD_WARPS * hd);
  const size_t smem2 = sizeof(float) * ((size_t)2 * T * hd + (size_t)2 * T * 32);
  ZS_REQUIRE(smem1 <= 227 * 1024 && smem2 <= 227 * 1024, "zs_mha_bwd_f32: sequence too long for shared memory");
  ZS_CUDA_CALL(cudaFuncSetAttribute(mha_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  ZS_CUDA_CALL(cudaFuncSetAttribute(mha_bwd_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  int slabs = (2 * sm_count() + B * heads - 1) / (B * heads);        // enough CTAs for two waves at small batch
  if (slabs < 2) slabs = 2;
  int rows_per_cta = (T + slabs - 1) / slabs;
  if (rows_per_cta < MHA_BWD_WARPS) rows_per_cta = MHA_BWD_WARPS;
  slabs = (T + rows_per_cta - 1) / rows_per_cta;
  cudaStream_t st = as_stream(stream);
  mha_bwd_rows_kernel<<<dim3(B * heads, slabs), MHA_BWD_WARPS * 32, smem1, st>>>(qkv, dO, dqkv, Pg, Sg, T, heads, hd, scale, rows_per_cta);
  ZS_CUDA_CHECK_LAUNCH("zs_mha_bwd_f32(rows)");
  mha_bwd_cols_kernel<<<dim3(B * heads, (T + 31) / 32), 256, smem2, st>>>(qkv, dO, dqkv, Pg, Sg, T, heads, hd);
  ZS_CUDA_CHECK_LAUNCH("zs_mha_bwd_f32(cols)");
  return ZS_OK;
}

extern "C" int zs_adamw_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                            float beta2, float eps, float weight_decay, int step, void* stream) {
  ZS_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "zs_adamw_f32: bad args");
  if (n == 0) return ZS_OK;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2 = sqrtf(1.0f - powf(beta2, (float)step));
  adamw_kernel<<<grid_for_n(n), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                           bc1, bc2);
  ZS_CUDA_CHECK_LAUNCH("zs_adamw_f32");
  return ZS_OK;
}

extern "C" int zs_bn_stats_f32(const float* x, int64_t M, int C, float eps, double* ws, float* mean, float* var, float* rstd,
                               void* stream) {
  ZS_REQUIRE(x && ws && mean && var && rstd && M > 0 && C > 0, "zs_bn_stats_f32: bad args");
  cudaStream_t st = as_stream(stream);
  ZS_CUDA_CALL(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  int64_t slabs = (M + 8 * 32 - 1) / (8 * 32);
  if (slabs > 512) slabs = 512;
  bn_stats_kernel<<<dim3((C + 31) / 32, (unsigned)slabs), dim3(32, 8), 0, st>>>(x, M, C, ws);
  ZS_CUDA_CHECK_LAUNCH("zs_bn_stats_f32");
  bn_stats_finish_kernel<<<(C + 255) / 256, 256, 0, st>>>(ws, M, C, eps, mean, var, rstd);
  ZS_CUDA_CHECK_LAUNCH("zs_bn_stats_f32(finish)");
  return ZS_OK;
}

extern "C" int zs_bn_bwd_f32(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int64_t M,
                             int C, double* ws, float* dx, float* dgamma, float* dbeta, void* stream) {
  ZS_REQUIRE(dy && x && mean && rstd && gamma && ws && dx && dgamma && dbeta && M > 0 && C > 0, "zs_bn_bwd_f32: bad args");
  cudaStream_t st = as_stream(stream);
  ZS_CUDA_CALL(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  int64_t slabs = (M + 8 * 32 - 1) / (8 * 32);
  if (slabs > 512) slabs = 512;
  bn_bwd_reduce_kernel<<<dim3((C + 31) / 32, (unsigned)slabs), dim3(32, 8), 0, st>>>(dy, x, mean, rstd, M, C, ws);
  ZS_CUDA_CHECK_LAUNCH("zs_bn_bwd_f32(reduce)");
  bn_bwd_apply_kernel<<<grid_for_n(M * C), 256, 0, st>>>(dy, x, mean, rstd, gamma, ws, M, C, dx, dgamma, dbeta);
  ZS_CUDA_CHECK_LAUNCH("zs_bn_bwd_f32(apply)");
  return ZS_OK;
}

extern "C" int zs_maxpool3x3s2_bwd_nhwc_f32(const float* x, const float* dy, float* dx, int B, int H, int W, int C, int pad_top,
                                            int pad_left, int OH, int OW, void* stream) {
  ZS_REQUIRE(x && dy && dx && B > 0 && H > 0 && W > 0 && C > 0 && OH > 0 && OW > 0, "zs_maxpool3x3s2_bwd_nhwc_f32: bad args");
  cudaStream_t st = as_stream(stream);
  ZS_CUDA_CALL(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * H * W * C, st));
  maxpool3x3s2_bwd_kernel<<<grid_for_n((int64_t)B * OH * OW * C), 256, 0, st>>>(x, dy, dx, B, H, W, C, pad_top, pad_left, OH, OW);
  ZS_CUDA_CHECK_LAUNCH("zs_maxpool3x3s2_bwd_nhwc_f32");
  return ZS_OK;
}

extern "C" int zs_avgpool_bwd_nhwc_f32(const float* dy, float* dx, int B, int HW, int C, void* stream) {
  ZS_REQUIRE(dy && dx && B > 0 && HW > 0 && C > 0, "zs_avgpool_bwd_nhwc_f32: bad args");
  avgpool_bwd_kernel<<<grid_for_n((int64_t)B * HW * C), 256, 0, as_stream(stream)>>>(dy, dx, B, HW, C);
  ZS_CUDA_CHECK_LAUNCH("zs_avgpool_bwd_nhwc_f32");
  return ZS_OK;
}

extern "C" int zs_layernorm_bwd_generic_f32(const float* dy, const float* x, const float* gamma, float eps, float* dx, float* xhat,
                                            int64_t rows, int cols, void* stream) {
  ZS_REQUIRE(dy && x && dx && rows >= 0 && cols > 0, "zs_layernorm_bwd_generic_f32: bad args");
  if (rows == 0) return ZS_OK;
  int64_t g = (rows + 7) / 8;
  const int cap = sm_count() * 8;
  layernorm_bwd_generic_kernel<<<(int)(g < cap ? g : cap), 256, 0, as_stream(stream)>>>(dy, x, gamma, eps, dx, xhat, rows, cols);
  ZS_CUDA_CHECK_LAUNCH("zs_layernorm_bwd_generic_f32");
  return ZS_OK;
}

extern "C" int zs_groupnorm_bwd_nhwc_f32(const float* dy, const float* x, const float* gamma, float* dx, float* dgamma, float* dbeta,
                                         int B, int HW, int C, int groups, float eps, void* stream) {
  ZS_REQUIRE(dy && x && gamma && dx && dgamma && dbeta && B > 0 && HW > 0 && C > 0 && groups > 0 && C % groups == 0 && C / groups <= 64 &&
                 512 % (C / groups) == 0,
             "zs_groupnorm_bwd_nhwc_f32: bad args (channels per group must divide 512 and be at most 64)");
  groupnorm_bwd_nhwc_kernel<<<B * groups, 512, 0, as_stream(stream)>>>(dy, x, gamma, dx, dgamma, dbeta, HW, C, groups, eps);
  ZS_CUDA_CHECK_LAUNCH("zs_groupnorm_bwd_nhwc_f32");
  return ZS_OK;
}

extern "C" int zs_bilinear_bwd_nhwc_f32(const float* dy, float* dx, int B, int H, int W, int C, int OH, int OW, int align_corners,
                                        void* stream) {
  ZS_REQUIRE(dy && dx && B > 0 && H > 0 && W > 0 && C > 0 && OH > 0 && OW > 0, "zs_bilinear_bwd_nhwc_f32: bad args");
  cudaStream_t st = as_stream(stream);
  if ((C & 3) == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0 && (int64_t)B * H * W * (C / 4) < (1LL << 31) &&
      (int64_t)OH * OW * (C / 4) < (1LL << 31)) {
    bilinear_bwd_gather_v4_kernel<<<grid_for_n((int64_t)B * H * W * (C / 4)), 256, 0, st>>>(
        reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(dx), B, H, W, C / 4, OH, OW, align_corners);
    ZS_CUDA_CHECK_LAUNCH("zs_bilinear_bwd_nhwc_f32");
    return ZS_OK;
  }
  ZS_CUDA_CALL(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * H * W * C, st));
  bilinear_bwd_nhwc_kernel<<<grid_for_n((int64_t)B * OH * OW * C), 256, 0, st>>>(dy, dx, B, H, W, C, OH, OW, align_corners);
  ZS_CUDA_CHECK_LAUNCH("zs_bilinear_bwd_nhwc_f32");
  return ZS_OK;
}

extern "C" int zs_unproject_normalize_bwd_f32(const float* depth, const float* mask, const float* K, const float* seen_points,
                                              const float* scale, const float* dseen, float* ddepth, float* dKinv, int B, int H, int W,
                                              void* stream) {
  ZS_REQUIRE(depth && mask && K && seen_points && scale && dseen && ddepth && dKinv && B > 0 && H > 0 && W > 0,
             "zs_unproject_normalize_bwd_f32: bad args");
  unproject_normalize_bwd_kernel<<<B, 1024, 0, as_stream(stream)>>>(depth, mask, K, seen_points, scale, dseen, ddepth, dKinv, H, W);
  ZS_CUDA_CHECK_LAUNCH("zs_unproject_normalize_bwd_f32");
  return ZS_OK;
}
