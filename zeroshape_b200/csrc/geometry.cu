// Geometry glue between the depth/intrinsics encoder and the seen-surface encoder, as ONE launch
// per op and no host synchronisation.
//   zs_intr_param2mtx_f32       <- Graph.intr_param2mtx        (model/compute_graph/graph_shape.py:89-113)
//   zs_unproject_normalize_f32  <- unproj_depth                (utils/camera.py:88-108)
//                                  valid_norm_fac              (utils/camera.py:52-78; a Python loop over
//                                                               the batch with boolean indexing = B host syncs)
//                                  normalise + zero background (graph_shape.py:140-141)
#include "common.cuh"

namespace zs {

__global__ void intr_param2mtx_kernel(const float* __restrict__ p, float* __restrict__ K, int B, float H, float W) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float f = 1.3875f;
  float sf = powf(4.0f, tanhf(p[b * 3 + 0]));
  float* k = K + b * 9;
  k[0] = (float)((double)1.3875 * (double)W) * sf;  // python: f * opt.W is a double product, then * fp32 tensor
  k[1] = 0.f;
  k[2] = W * 0.5f + tanhf(p[b * 3 + 1]) * W / 2.0f;
  k[3] = 0.f;
  k[4] = (float)((double)1.3875 * (double)H) * sf;
  k[5] = H * 0.5f + tanhf(p[b * 3 + 2]) * H / 2.0f;
  k[6] = 0.f; k[7] = 0.f; k[8] = 1.f;
  (void)f;
}

__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
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
}

// one CTA (1024 threads) per sample; three passes over H*W pixels (L2 resident)
__global__ void __launch_bounds__(1024)
unproject_normalize_kernel(const float* __restrict__ depth, const float* __restrict__ mask, const float* __restrict__ K,
                           float* __restrict__ pts, float* __restrict__ mean_out, float* __restrict__ scale_out,
                           int H, int W) {
  __shared__ float sh[33];
  __shared__ float kinv[9];
  const int b = blockIdx.x;
  const int HW = H * W;
  if (threadIdx.x == 0) {
    const float* k = K + b * 9;
    // general 3x3 inverse by adjugate (torch.linalg.inv in the reference)
    float a = k[0], bb = k[1], c = k[2], d = k[3], e = k[4], f = k[5], g = k[6], h = k[7], i = k[8];
    float A = e * i - f * h, Bc = -(d * i - f * g), C = d * h - e * g;
    float det = a * A + bb * Bc + c * C;
    float r = 1.0f / det;
    kinv[0] = A * r;  kinv[1] = -(bb * i - c * h) * r; kinv[2] = (bb * f - c * e) * r;
    kinv[3] = Bc * r; kinv[4] = (a * i - c * g) * r;   kinv[5] = -(a * f - c * d) * r;
    kinv[6] = C * r;  kinv[7] = -(a * h - bb * g) * r; kinv[8] = (a * e - bb * d) * r;
  }
  __syncthreads();
  const float* dp = depth + (int64_t)b * HW;
  const float* mp = mask ? mask + (int64_t)b * HW : nullptr;
  float* op = pts + (int64_t)b * HW * 3;
  auto point = [&](int idx, float& x, float& y, float& z) {
    float px = (float)(idx % W), py = (float)(idx / W);
    float dz = dp[idx];
    x = (kinv[0] * px + kinv[1] * py + kinv[2]) * dz;
    y = (kinv[3] * px + kinv[4] * py + kinv[5]) * dz;
    z = (kinv[6] * px + kinv[7] * py + kinv[8]) * dz;
  };
  if (mask == nullptr) {   // raw unprojection only (utils/camera.py:88-108)
    for (int idx = threadIdx.x; idx < HW; idx += blockDim.x) {
      float x, y, z;
      point(idx, x, y, z);
      op[idx * 3 + 0] = x; op[idx * 3 + 1] = y; op[idx * 3 + 2] = z;
    }
    return;
  }
  float sx = 0.f, sy = 0.f, sz = 0.f, cnt = 0.f;
  for (int idx = threadIdx.x; idx < HW; idx += blockDim.x) {
    if (mp[idx] > 0.5f) {
      float x, y, z;
      point(idx, x, y, z);
      sx += x; sy += y; sz += z; cnt += 1.f;
    }
  }
  sx = block_reduce(sx, sh, false);
  sy = block_reduce(sy, sh, false);
  sz = block_reduce(sz, sh, false);
  cnt = block_reduce(cnt, sh, false);
  float mx = sx / cnt, my = sy / cnt, mz = sz / cnt;   // cnt==0 -> NaN, like torch.mean of empty
  float md = -INFINITY;
  for (int idx = threadIdx.x; idx < HW; idx += blockDim.x) {
    if (mp[idx] > 0.5f) {
      float x, y, z;
      point(idx, x, y, z);
      x -= mx; y -= my; z -= mz;
      md = fmaxf(md, sqrtf(x * x + y * y + z * z));
    }
  }
  md = block_reduce(md, sh, true);
  if (threadIdx.x == 0) {
    mean_out[b * 3 + 0] = mx; mean_out[b * 3 + 1] = my; mean_out[b * 3 + 2] = mz;
    scale_out[b] = md;
  }
  for (int idx = threadIdx.x; idx < HW; idx += blockDim.x) {
    float x = 0.f, y = 0.f, z = 0.f;
    if (mp[idx] > 0.5f) {
      point(idx, x, y, z);
      x = (x - mx) / md; y = (y - my) / md; z = (z - mz) / md;
    }
    op[idx * 3 + 0] = x; op[idx * 3 + 1] = y; op[idx * 3 + 2] = z;
  }
}

}  // namespace zs

using namespace zs;

extern "C" int zs_intr_param2mtx_f32(const float* params, float* K, int B, int H, int W, void* stream) {
  ZS_REQUIRE(params && K && B > 0 && H > 0 && W > 0, "zs_intr_param2mtx_f32: bad args");
  intr_param2mtx_kernel<<<(B + 63) / 64, 64, 0, as_stream(stream)>>>(params, K, B, (float)H, (float)W);
  ZS_CUDA_CHECK_LAUNCH("zs_intr_param2mtx_f32");
  return ZS_OK;
}

extern "C" size_t zs_unproject_ws_bytes(int B) { (void)B; return 0; }

extern "C" int zs_unproject_normalize_f32(const float* depth, const float* mask, const float* K,
                                          float* seen_points, float* mean, float* scale,
                                          int B, int H, int W, void* ws, void* stream) {
  (void)ws;
  ZS_REQUIRE(depth && K && seen_points && B > 0 && H > 0 && W > 0, "zs_unproject_normalize_f32: bad args");
  ZS_REQUIRE(mask == nullptr || (mean && scale), "zs_unproject_normalize_f32: mean/scale required with a mask");
  unproject_normalize_kernel<<<B, 1024, 0, as_stream(stream)>>>(depth, mask, K, seen_points, mean, scale, H, W);
  ZS_CUDA_CHECK_LAUNCH("zs_unproject_normalize_f32");
  return ZS_OK;
}
