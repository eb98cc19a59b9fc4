// Input pipeline of demo.py / data/*.py on the device: the byte-exact crop + PIL-bicubic resize of
// an RGBA image, the mask compositing of preprocess_image, and demo.py's erode_mask.
//
//   zs_rgba_crop_resize_u8 : torchvision_F.crop(image, top, left, h, w)  (PIL crop: zero = transparent black outside the image,
//                            demo.py:33-41, data/synthetic.py:201-210)  followed by  image.resize((W, H))  (demo.py:45,
//                            data/synthetic.py:195): PIL's default BICUBIC on an RGBA image, i.e. RGBA -> premultiplied RGBa,
//                            two separable 8-bit passes with 22-bit fixed-point coefficients (horizontal first, clipped to a
//                            byte between the passes), RGBa -> RGBA.  Integer arithmetic throughout: bit-exact with Pillow.
//   zs_rgba_composite_f32  : torchvision to_tensor (byte / 255) + `rgb * mask + bgcolor * (1 - mask)`, `mask > 0.5`
//                            (demo.py:46-52) -> rgb [3,H,W], mask [1,H,W] fp32, the same fp32 operations in the same order.
//   zs_erode_square_f32    : cv2.erode(mask, ones(3,3), iterations=r) of demo.py:70-75 = minimum over the (2r+1)^2 window
//                            clipped to the image (OpenCV's default erosion border never erodes).
// The coefficient tables are built on the host exactly as Pillow's precompute_coeffs / normalize_coeffs_8bpc do (double
// arithmetic) and are tiny: bounds [out][2] (first source index, tap count) and kk [out][ksize] int32.
#include "common.cuh"

namespace zs {

__device__ __forceinline__ int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
// Pillow MULDIV255: round(a * b / 255) without a division
__device__ __forceinline__ int muldiv255(int a, int b) {
  const int t = a * b + 128;
  return ((t >> 8) + t) >> 8;
}

struct ResizeParams {
  const uint8_t* src; int H0, W0;           // source RGBA image [H0, W0, 4]
  int left, top, cw, ch;                    // crop window (may leave the image)
  uint8_t* out; int OH, OW;
  const int* xb; const int* xk; int xks;    // horizontal: bounds [OW][2], coefficients [OW][xks]
  const int* yb; const int* yk; int yks;    // vertical:   bounds [OH][2], coefficients [OH][yks]
};

// premultiplied pixel of the CROPPED image at (cy, cx); zero outside the source image
__device__ __forceinline__ void load_premul(const ResizeParams& p, int cy, int cx, int (&c)[4]) {
  const int sy = cy + p.top, sx = cx + p.left;
  if (sy < 0 || sy >= p.H0 || sx < 0 || sx >= p.W0) { c[0] = c[1] = c[2] = c[3] = 0; return; }
  const uchar4 v = *reinterpret_cast<const uchar4*>(p.src + ((size_t)sy * p.W0 + sx) * 4);
  c[0] = muldiv255(v.x, v.w); c[1] = muldiv255(v.y, v.w); c[2] = muldiv255(v.z, v.w); c[3] = v.w;
}

// one thread per output pixel: for every source row of its vertical footprint the horizontally resampled byte is formed on
// the fly (exactly Pillow's intermediate image), then the vertical pass; at 224 x 224 outputs that is ~10^2 taps^2 per pixel
__global__ void rgba_crop_resize_kernel(ResizeParams p) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y * blockDim.y + threadIdx.y;
  if (ox >= p.OW || oy >= p.OH) return;
  constexpr int PB = 22;
  const int xmin = p.xb[2 * ox], xcnt = p.xb[2 * ox + 1];
  const int ymin = p.yb[2 * oy], ycnt = p.yb[2 * oy + 1];
  const int* kx = p.xk + (size_t)ox * p.xks;
  const int* ky = p.yk + (size_t)oy * p.yks;
  const bool horiz = p.cw != p.OW, vert = p.ch != p.OH;       // Pillow skips a pass whose size does not change
  int acc[4] = {1 << (PB - 1), 1 << (PB - 1), 1 << (PB - 1), 1 << (PB - 1)};
  int res[4];
  for (int r = 0; r < (vert ? ycnt : 1); ++r) {
    const int cy = vert ? ymin + r : oy;
    int h[4];
    if (horiz) {
      int ss[4] = {1 << (PB - 1), 1 << (PB - 1), 1 << (PB - 1), 1 << (PB - 1)};
      for (int x = 0; x < xcnt; ++x) {
        int c[4];
        load_premul(p, cy, xmin + x, c);
        const int k = kx[x];
#pragma unroll
        for (int b = 0; b < 4; ++b) ss[b] += c[b] * k;
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) h[b] = clip8(ss[b] >> PB);
    } else {
      load_premul(p, cy, ox, h);
    }
    if (vert) {
      const int k = ky[r];
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[b] += h[b] * k;
    } else {
#pragma unroll
      for (int b = 0; b < 4; ++b) res[b] = h[b];
    }
  }
  if (vert) {
#pragma unroll
    for (int b = 0; b < 4; ++b) res[b] = clip8(acc[b] >> PB);
  }
  // RGBa -> RGBA (Pillow rgba2rgbA)
  const int a = res[3];
  if (a != 255 && a != 0) {
#pragma unroll
    for (int b = 0; b < 3; ++b) res[b] = clip8((255 * res[b]) / a);
  }
  *reinterpret_cast<uchar4*>(p.out + ((size_t)oy * p.OW + ox) * 4) = make_uchar4(res[0], res[1], res[2], res[3]);
}

__global__ void rgba_composite_kernel(const uint8_t* __restrict__ img, int H, int W, int use_bg, float bg, float* __restrict__ rgb,
                                      float* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const uchar4 v = *reinterpret_cast<const uchar4*>(img + (size_t)i * 4);
  float c[3] = {__fdiv_rn((float)v.x, 255.0f), __fdiv_rn((float)v.y, 255.0f), __fdiv_rn((float)v.z, 255.0f)};
  float m = __fdiv_rn((float)v.w, 255.0f);
  if (use_bg) {
    const float bgm = __fmul_rn(bg, __fsub_rn(1.0f, m));        // opt.data.bgcolor * (1 - mask)
#pragma unroll
    for (int b = 0; b < 3; ++b) c[b] = __fadd_rn(__fmul_rn(c[b], m), bgm);
    m = m > 0.5f ? 1.0f : 0.0f;
  }
#pragma unroll
  for (int b = 0; b < 3; ++b) rgb[(size_t)b * H * W + i] = c[b];
  mask[i] = m;
}

__global__ void erode_square_kernel(const float* __restrict__ mask, float* __restrict__ out, int B, int H, int W, int r) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H * W) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const float* img = mask + (i / ((int64_t)H * W)) * H * W;
  float m = 3.0e38f;
  for (int dy = -r; dy <= r; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    for (int dx = -r; dx <= r; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= W) continue;
      m = fminf(m, img[(size_t)yy * W + xx]);
    }
  }
  out[i] = m;
}

}  // namespace zs

using namespace zs;

extern "C" int zs_rgba_crop_resize_u8(const uint8_t* src, int H0, int W0, int left, int top, int cw, int ch, uint8_t* out, int OH,
                                      int OW, const int* xbounds, const int* xcoef, int xksize, const int* ybounds,
                                      const int* ycoef, int yksize, void* stream) {
  ZS_REQUIRE(src && out && xbounds && xcoef && ybounds && ycoef, "zs_rgba_crop_resize_u8: null pointer");
  ZS_REQUIRE(H0 > 0 && W0 > 0 && cw > 0 && ch > 0 && OH > 0 && OW > 0 && xksize > 0 && yksize > 0, "zs_rgba_crop_resize_u8: bad sizes");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(src) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
             "zs_rgba_crop_resize_u8: images must be 4-byte aligned");
  ResizeParams p{src, H0, W0, left, top, cw, ch, out, OH, OW, xbounds, xcoef, xksize, ybounds, ycoef, yksize};
  dim3 block(16, 16), grid((OW + 15) / 16, (OH + 15) / 16);
  rgba_crop_resize_kernel<<<grid, block, 0, as_stream(stream)>>>(p);
  ZS_CUDA_CHECK_LAUNCH("zs_rgba_crop_resize_u8");
  return ZS_OK;
}

extern "C" int zs_rgba_composite_f32(const uint8_t* img, int H, int W, int use_bgcolor, float bgcolor, float* rgb, float* mask,
                                     void* stream) {
  ZS_REQUIRE(img && rgb && mask && H > 0 && W > 0, "zs_rgba_composite_f32: null pointer / bad sizes");
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(img) & 3) == 0, "zs_rgba_composite_f32: image must be 4-byte aligned");
  rgba_composite_kernel<<<(H * W + 255) / 256, 256, 0, as_stream(stream)>>>(img, H, W, use_bgcolor, bgcolor, rgb, mask);
  ZS_CUDA_CHECK_LAUNCH("zs_rgba_composite_f32");
  return ZS_OK;
}

extern "C" int zs_erode_square_f32(const float* mask, float* out, int B, int H, int W, int radius, void* stream) {
  ZS_REQUIRE(mask && out && B > 0 && H > 0 && W > 0 && radius >= 0, "zs_erode_square_f32: null pointer / bad sizes");
  const int64_t n = (int64_t)B * H * W;
  erode_square_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(mask, out, B, H, W, radius);
  ZS_CUDA_CHECK_LAUNCH("zs_erode_square_f32");
  return ZS_OK;
}
