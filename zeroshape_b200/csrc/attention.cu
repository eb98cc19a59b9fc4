// fp32 attention kernels (bit-faithful path).
//   zs_mha_f32             : token self-attention (ViT blocks, latent branch of the implicit decoder)
//   zs_point_attention_f32 : query-point -> (latents + self) attention of ImplFuncAttention
#include "common.cuh"

namespace zs {

// One CTA per (b*heads, query chunk).  K,V of that head staged in smem (row stride hd+1 floats so that
// lane-varying key index is bank-conflict free); one warp per query.
__global__ void mha_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T, int heads, int hd,
                           float scale, int q_per_cta) {
  extern __shared__ float smem[];
  const int ldk = hd + 1;
  float* Ks = smem;                    // [T][hd+1]
  float* Vs = Ks + (size_t)T * ldk;    // [T][hd+1]
  const int nwarps = blockDim.x >> 5;
  float* Ps = Vs + (size_t)T * ldk;    // [nwarps][T]
  float* Qs = Ps + (size_t)nwarps * T; // [nwarps][hd]
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int C = heads * hd;
  const float* base = qkv + (int64_t)b * T * 3 * C;
  for (int i = threadIdx.x; i < T * hd; i += blockDim.x) {
    int t = i / hd, d = i % hd;
    Ks[t * ldk + d] = base[(int64_t)t * 3 * C + C + h * hd + d];
    Vs[t * ldk + d] = base[(int64_t)t * 3 * C + 2 * C + h * hd + d];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.y * q_per_cta;
  const int q1 = min(T, q0 + q_per_cta);
  float* P = Ps + warp * T;
  float* Q = Qs + warp * hd;
  for (int q = q0 + warp; q < q1; q += nwarps) {
    for (int d = lane; d < hd; d += 32) Q[d] = base[(int64_t)q * 3 * C + h * hd + d];
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) {
      float s = 0.f;
      for (int d = 0; d < hd; ++d) s = fmaf(Q[d], Ks[j * ldk + d], s);
      s *= scale;
      P[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
      float e = expf(P[j] - mx);
      P[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    float inv = 1.0f / sum;
    __syncwarp();
    for (int d = lane; d < hd; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < T; ++j) acc = fmaf(P[j] * inv, Vs[j * ldk + d], acc);
      out[((int64_t)b * T + q) * C + h * hd + d] = acc;
    }
    __syncwarp();
  }
}

// grid (ceil(P/128), B); 128 threads; each thread owns one query point and loops over heads.
template <int HD>
__global__ void __launch_bounds__(128)
point_attention_kernel(const float* __restrict__ qkv_p, const float* __restrict__ k_lat,
                       const float* __restrict__ v_lat, int ld_lat, float* __restrict__ out,
                       float* __restrict__ attn, float attn_scale, int attn_accumulate,
                       int P, int L, int heads, float scale) {
  extern __shared__ float smem[];
  float* Ks = smem;            // [L][HD]
  float* Vs = Ks + L * HD;     // [L][HD]
  const int b = blockIdx.y;
  const int p = blockIdx.x * 128 + threadIdx.x;
  const bool live = p < P;
  const int C = heads * HD;
  const float* qrow = qkv_p + ((int64_t)b * P + (live ? p : 0)) * 3 * C;
  float* arow = attn ? attn + ((int64_t)b * P + (live ? p : 0)) * L : nullptr;
  for (int h = 0; h < heads; ++h) {
    __syncthreads();
    for (int i = threadIdx.x; i < L * HD; i += 128) {
      int t = i / HD, d = i % HD;
      Ks[i] = k_lat[((int64_t)b * L + t) * ld_lat + h * HD + d];
      Vs[i] = v_lat[((int64_t)b * L + t) * ld_lat + h * HD + d];
    }
    __syncthreads();
    if (!live) continue;
    float q[HD], acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) { q[d] = qrow[h * HD + d]; acc[d] = 0.f; }
    // self score
    float s_self = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) s_self = fmaf(q[d], qrow[C + h * HD + d], s_self);
    s_self *= scale;
    float mx = s_self, sum = 1.0f;  // running max / sum, self term first
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = qrow[2 * C + h * HD + d];
    for (int j = 0; j < L; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) s = fmaf(q[d], Ks[j * HD + d], s);
      s *= scale;
      if (s > mx) {
        float c = expf(mx - s);
        sum *= c;
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] *= c;
        mx = s;
      }
      float e = expf(s - mx);
      sum += e;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] = fmaf(e, Vs[j * HD + d], acc[d]);
    }
    float inv = 1.0f / sum;
#pragma unroll
    for (int d = 0; d < HD; ++d) out[((int64_t)b * P + p) * C + h * HD + d] = acc[d] * inv;
    if (arow) {
      float w = attn_scale / (float)heads;
      for (int j = 0; j < L; ++j) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) s = fmaf(q[d], Ks[j * HD + d], s);
        float pj = expf(s * scale - mx) * inv * w;
        if (h == 0 && !attn_accumulate) arow[j] = pj; else arow[j] += pj;
      }
    }
  }
}

}  // namespace zs

using namespace zs;

extern "C" int zs_mha_f32(const float* qkv, float* out, int B, int T, int heads, int hd, float scale, void* stream) {
  ZS_REQUIRE(qkv && out && B > 0 && T > 0 && heads > 0 && hd > 0, "zs_mha_f32: bad args");
  const int threads = 256, nwarps = threads / 32;
  size_t smem = ((size_t)2 * T * (hd + 1) + (size_t)nwarps * T + (size_t)nwarps * hd) * sizeof(float);
  ZS_REQUIRE(smem <= 227 * 1024, "zs_mha_f32: T=%d hd=%d needs %zu B smem", T, hd, smem);
  static thread_local size_t configured = 0;
  if (smem > configured) {
    ZS_CUDA_CALL(cudaFuncSetAttribute(mha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int ctas_bh = B * heads;
  int chunks = (2 * sm_count() + ctas_bh - 1) / ctas_bh;
  if (chunks < 1) chunks = 1;
  int q_per = (T + chunks - 1) / chunks;
  if (q_per < nwarps) q_per = nwarps;
  chunks = (T + q_per - 1) / q_per;
  dim3 grid(ctas_bh, chunks);
  mha_kernel<<<grid, threads, smem, as_stream(stream)>>>(qkv, out, T, heads, hd, scale, q_per);
  ZS_CUDA_CHECK_LAUNCH("zs_mha_f32");
  return ZS_OK;
}

extern "C" int zs_point_attention_f32(const float* qkv_p, const float* k_lat, const float* v_lat, int ld_lat,
                                      float* out, float* attn, float attn_scale, int attn_accumulate,
                                      int B, int P, int L, int heads, int hd, float scale, void* stream) {
  ZS_REQUIRE(qkv_p && k_lat && v_lat && out && B > 0 && P >= 0 && L > 0 && heads > 0, "zs_point_attention_f32: bad args");
  ZS_REQUIRE(hd == 32, "zs_point_attention_f32: head dim %d unsupported (32 only)", hd);
  if (P == 0) return ZS_OK;
  size_t smem = (size_t)2 * L * hd * sizeof(float);
  ZS_REQUIRE(smem <= 200 * 1024, "zs_point_attention_f32: L too large");
  static thread_local size_t configured = 0;
  if (smem > configured) {
    ZS_CUDA_CALL(cudaFuncSetAttribute(point_attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((P + 127) / 128, B);
  point_attention_kernel<32><<<grid, 128, smem, as_stream(stream)>>>(qkv_p, k_lat, v_lat, ld_lat, out, attn, attn_scale,
                                                                    attn_accumulate, P, L, heads, scale);
  ZS_CUDA_CHECK_LAUNCH("zs_point_attention_f32");
  return ZS_OK;
}
