// GPU marching cubes (classify -> scan -> emit, welded vertices) and area-weighted surface sampling.
//
// Replaces the GPU->CPU->GPU hop of utils/eval_3D.py:233-263 (PyMCubes 0.1.4 `marching_cubes` on a
// numpy volume + trimesh 4.0.8 `Trimesh.sample`).  Semantics restated from those libraries:
//   * corner "inside" iff value <= iso; one vertex per crossed grid edge (welded), placed by linear
//     interpolation evaluated in double: x1 + (x2-x1)*(iso-f1)/(f2-f1) (midpoint if f1==f2);
//   * vertices in array-index units, axis order i,j,k = x,y,z; classic 256-case triangle table;
//   * sampling: face ~ area, uniform barycentric with the u+v>1 reflection.
// HBM-bound: volume read twice (classify, emit) + one flag/scan workspace pass.
#include "common.cuh"
#include "mc_tables.h"

namespace zs {

// ---- generic 3-phase exclusive scan (1024 elements per CTA) -----------------------------------
template <typename T>
__device__ __forceinline__ T block_exclusive_scan_256(T v, T* sh, T& total) {
  // 256 threads; returns exclusive prefix of v across the block, total = block sum
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) sh[w] = inc;
  __syncthreads();
  if (w == 0) {
    T s = lane < 8 ? sh[lane] : T(0);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      T n = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += n;
    }
    if (lane < 8) sh[lane] = s;
  }
  __syncthreads();
  T base = w > 0 ? sh[w - 1] : T(0);
  total = sh[7];
  __syncthreads();
  return base + inc - v;
}

template <typename T>
__global__ void scan_phase1(const T* __restrict__ in, T* __restrict__ out, T* __restrict__ bsum, int64_t n) {
  __shared__ T sh[8];
  int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
  T v[4];
  T s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[i] = base + i < n ? in[base + i] : T(0); s += v[i]; }
  T total;
  T ex = block_exclusive_scan_256<T>(s, sh, total);
#pragma unroll
  for (int i = 0; i < 4; ++i) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}
template <typename T>
__global__ void scan_phase2(T* __restrict__ bsum, int nb, T* __restrict__ total_out) {
  // single CTA, sequential chunks of 256
  __shared__ T sh[8];
  __shared__ T carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int c = 0; c < nb; c += 256) {
    int i = c + threadIdx.x;
    T v = i < nb ? bsum[i] : T(0);
    T total;
    T ex = block_exclusive_scan_256<T>(v, sh, total);
    T carry = carry_s;
    if (i < nb) bsum[i] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}
template <typename T>
__global__ void scan_phase3(T* __restrict__ out, const T* __restrict__ bsum, int64_t n) {
  int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
  T add = bsum[blockIdx.x];
#pragma unroll
  for (int i = 0; i < 4; ++i) if (base + i < n) out[base + i] += add;
}
template <typename T>
static void exclusive_scan(const T* in, T* out, T* bsum, int64_t n, T* total_out, cudaStream_t st) {
  int nb = (int)((n + 1023) / 1024);
  scan_phase1<T><<<nb, 256, 0, st>>>(in, out, bsum, n);
  scan_phase2<T><<<1, 256, 0, st>>>(bsum, nb, total_out);
  scan_phase3<T><<<nb, 256, 0, st>>>(out, bsum, n);
}

// ---- marching cubes ---------------------------------------------------------------------------
struct McWs {
  int32_t* vcnt;    // [n^3] per grid point: number of crossed +x/+y/+z edges, then exclusive scan
  int32_t* tcnt;    // [n^3] per cell (stored at its low corner): triangle count, then exclusive scan
  int32_t* vbs;     // block sums
  int32_t* tbs;
  uint8_t* vflags;  // [n^3] bit0 +x crossed, bit1 +y, bit2 +z
  uint8_t* cases;   // [n^3] cube case index of the cell at this low corner (0 if none)
};
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static McWs carve(void* ws, int nx, int n) {
  size_t n3 = (size_t)nx * n * n, nb = (n3 + 1023) / 1024;
  uint8_t* p = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(ws), 256));
  McWs w;
  w.vcnt = reinterpret_cast<int32_t*>(p); p += align_up(n3 * 4, 256);
  w.tcnt = reinterpret_cast<int32_t*>(p); p += align_up(n3 * 4, 256);
  w.vbs = reinterpret_cast<int32_t*>(p); p += align_up(nb * 4, 256);
  w.tbs = reinterpret_cast<int32_t*>(p); p += align_up(nb * 4, 256);
  w.vflags = p; p += align_up(n3, 256);
  w.cases = p;
  return w;
}

// vol is [nx, n, n] (an x-slab of the (n)^3 grid, or the whole grid when nx == n): cells whose low corner lies in slices
// 0 .. nx-2, one vertex per crossed edge owned by its low corner
__global__ void mc_classify_kernel(const float* __restrict__ vol, int nx, int n, float iso, McWs w) {
  int64_t n3 = (int64_t)nx * n * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n3; idx += (int64_t)gridDim.x * blockDim.x) {
    int k = (int)(idx % n);
    int64_t t = idx / n;
    int j = (int)(t % n), i = (int)(t / n);
    auto in = [&](int di, int dj, int dk) { return vol[idx + ((int64_t)di * n + dj) * n + dk] <= iso; };
    bool c0 = in(0, 0, 0);
    bool hx = i + 1 < nx, hy = j + 1 < n, hz = k + 1 < n;
    bool cx = hx ? in(1, 0, 0) : c0, cy = hy ? in(0, 1, 0) : c0, cz = hz ? in(0, 0, 1) : c0;
    int flags = (hx && cx != c0 ? 1 : 0) | (hy && cy != c0 ? 2 : 0) | (hz && cz != c0 ? 4 : 0);
    int cs = 0;
    if (hx && hy && hz) {
      // corner numbering: 0:(0,0,0) 1:(1,0,0) 2:(1,1,0) 3:(0,1,0) 4:(0,0,1) 5:(1,0,1) 6:(1,1,1) 7:(0,1,1)
      cs = (c0 ? 1 : 0) | (cx ? 2 : 0) | (in(1, 1, 0) ? 4 : 0) | (cy ? 8 : 0) | (cz ? 16 : 0) |
           (in(1, 0, 1) ? 32 : 0) | (in(1, 1, 1) ? 64 : 0) | (in(0, 1, 1) ? 128 : 0);
    }
    w.vflags[idx] = (uint8_t)flags;
    w.cases[idx] = (uint8_t)cs;
    w.vcnt[idx] = __popc(flags);
    w.tcnt[idx] = kMcNumTris[cs];
  }
}

__global__ void mc_emit_kernel(const float* __restrict__ vol, int nx, int n, float iso, McWs w, float* __restrict__ verts,
                               int32_t* __restrict__ faces, int x_offset) {
  int64_t n3 = (int64_t)nx * n * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n3; idx += (int64_t)gridDim.x * blockDim.x) {
    int k = (int)(idx % n);
    int64_t t = idx / n;
    int j = (int)(t % n), i = (int)(t / n);
    int flags = w.vflags[idx];
    if (flags) {
      int vid = w.vcnt[idx];
      double f1 = (double)vol[idx];
      const int64_t strides[3] = {(int64_t)n * n, n, 1};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (!(flags & (1 << a))) continue;
        double f2 = (double)vol[idx + strides[a]];
        double tpar = (f2 == f1) ? 0.5 : ((double)iso - f1) / (f2 - f1);
        float px = (float)(i + x_offset), py = (float)j, pz = (float)k;
        if (a == 0) px = (float)((double)(i + x_offset) + tpar);
        if (a == 1) py = (float)((double)j + tpar);
        if (a == 2) pz = (float)((double)k + tpar);
        verts[(int64_t)vid * 3 + 0] = px;
        verts[(int64_t)vid * 3 + 1] = py;
        verts[(int64_t)vid * 3 + 2] = pz;
        ++vid;
      }
    }
    int cs = w.cases[idx];
    int nt = kMcNumTris[cs];
    if (nt) {
      int64_t f0 = w.tcnt[idx];
      for (int tt = 0; tt < nt; ++tt) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          int e = kMcTris[cs][tt * 3 + c];
          int64_t oidx = idx + ((int64_t)kMcEdgeOwner[e][0] * n + kMcEdgeOwner[e][1]) * n + kMcEdgeOwner[e][2];
          int axis = kMcEdgeOwner[e][3];
          int of = w.vflags[oidx];
          int rank = __popc(of & ((1 << axis) - 1));
          faces[(f0 + tt) * 3 + c] = w.vcnt[oidx] + rank;
        }
      }
    }
  }
}

// ---- surface sampling -------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
__device__ __forceinline__ double u01(uint64_t seed, uint64_t ctr, uint64_t stream) {
  uint64_t r = splitmix64(splitmix64(seed ^ (stream * 0xD1B54A32D192ED03ULL)) + ctr);
  return (double)(r >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void face_area_kernel(const float* __restrict__ verts, const int32_t* __restrict__ faces, int F,
                                 float vscale, float voffset, double* __restrict__ area) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
    double p[3][3];
    for (int c = 0; c < 3; ++c) {
      const float* v = verts + (int64_t)faces[f * 3 + c] * 3;
      for (int d = 0; d < 3; ++d) p[c][d] = (double)(v[d] * vscale + voffset);
    }
    double ax = p[1][0] - p[0][0], ay = p[1][1] - p[0][1], az = p[1][2] - p[0][2];
    double bx = p[2][0] - p[0][0], by = p[2][1] - p[0][1], bz = p[2][2] - p[0][2];
    double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
    area[f] = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
  }
}

__global__ void mesh_sample_kernel(const float* __restrict__ verts, const int32_t* __restrict__ faces, int F,
                                   float vscale, float voffset, const double* __restrict__ cdf_excl,
                                   const double* __restrict__ total, int S, uint64_t seed, float* __restrict__ out) {
  double tot = *total;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
    double pick = u01(seed, s, 0) * tot;
    // largest f with cdf_excl[f] <= pick  (== searchsorted on the inclusive cumsum)
    int lo = 0, hi = F - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (cdf_excl[mid] <= pick) lo = mid; else hi = mid - 1;
    }
    double r1 = u01(seed, s, 1), r2 = u01(seed, s, 2);
    if (r1 + r2 > 1.0) { r1 = 1.0 - r1; r2 = 1.0 - r2; }
    float p[3][3];
    for (int c = 0; c < 3; ++c) {
      const float* v = verts + (int64_t)faces[lo * 3 + c] * 3;
      for (int d = 0; d < 3; ++d) p[c][d] = v[d] * vscale + voffset;
    }
    for (int d = 0; d < 3; ++d)
      out[(int64_t)s * 3 + d] = (float)((double)p[0][d] + r1 * ((double)p[1][d] - p[0][d]) + r2 * ((double)p[2][d] - p[0][d]));
  }
}

static inline int grid_for(int64_t n) {
  int64_t g = (n + 255) / 256, cap = (int64_t)sm_count() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace zs

using namespace zs;

extern "C" size_t zs_mc_ws_bytes(int n) { return zs_mc_slab_ws_bytes(n, n); }

extern "C" size_t zs_mc_slab_ws_bytes(int nx, int n) {
  if (n < 2 || nx < 1) return 0;
  size_t n3 = (size_t)nx * n * n, nb = (n3 + 1023) / 1024;
  return 256 + 2 * align_up(n3 * 4, 256) + 2 * align_up(nb * 4, 256) + 2 * align_up(n3, 256);
}

extern "C" int zs_mc_count(const float* vol, int n, float iso, void* ws, int32_t* counts, void* stream) {
  return zs_mc_slab_count(vol, n, n, iso, ws, counts, stream);
}

extern "C" int zs_mc_emit(const float* vol, int n, float iso, void* ws, float* verts, int32_t* faces, void* stream) {
  return zs_mc_slab_emit(vol, n, n, iso, ws, verts, faces, 0, stream);
}

extern "C" int zs_mc_slab_count(const float* vol, int nx, int n, float iso, void* ws, int32_t* counts, void* stream) {
  ZS_REQUIRE(vol && ws && counts && n >= 2 && n <= 1024 && nx >= 1 && nx <= n, "zs_mc_slab_count: bad args (nx=%d n=%d)", nx, n);
  cudaStream_t st = as_stream(stream);
  McWs w = carve(ws, nx, n);
  int64_t n3 = (int64_t)nx * n * n;
  mc_classify_kernel<<<grid_for(n3), 256, 0, st>>>(vol, nx, n, iso, w);
  exclusive_scan<int32_t>(w.vcnt, w.vcnt, w.vbs, n3, counts + 0, st);
  exclusive_scan<int32_t>(w.tcnt, w.tcnt, w.tbs, n3, counts + 1, st);
  count_launches(6);
  ZS_CUDA_CHECK_LAUNCH("zs_mc_slab_count");
  return ZS_OK;
}

extern "C" int zs_mc_slab_emit(const float* vol, int nx, int n, float iso, void* ws, float* verts, int32_t* faces, int x_offset,
                               void* stream) {
  ZS_REQUIRE(vol && ws && n >= 2 && n <= 1024 && nx >= 1 && nx <= n, "zs_mc_slab_emit: bad args");
  McWs w = carve(ws, nx, n);
  int64_t n3 = (int64_t)nx * n * n;
  mc_emit_kernel<<<grid_for(n3), 256, 0, as_stream(stream)>>>(vol, nx, n, iso, w, verts, faces, x_offset);
  ZS_CUDA_CHECK_LAUNCH("zs_mc_slab_emit");
  return ZS_OK;
}

extern "C" size_t zs_mesh_sample_ws_bytes(int F) {
  if (F <= 0) return 256;
  size_t nb = ((size_t)F + 1023) / 1024;
  return 512 + align_up((size_t)F * 8, 256) + align_up(nb * 8, 256) + 256;
}

extern "C" int zs_mesh_sample(const float* verts, const int32_t* faces, int V, int F, float vscale, float voffset,
                              int S, uint64_t seed, void* ws, float* points, void* stream) {
  ZS_REQUIRE(points && S >= 0 && F >= 0, "zs_mesh_sample: bad args");
  cudaStream_t st = as_stream(stream);
  if (S == 0) return ZS_OK;
  if (F == 0) {  // utils/eval_3D.py:262: empty mesh -> zeros
    ZS_CUDA_CALL(cudaMemsetAsync(points, 0, (size_t)S * 3 * sizeof(float), st));
    return ZS_OK;
  }
  ZS_REQUIRE(verts && faces && ws && V > 0, "zs_mesh_sample: null pointer");
  uint8_t* p = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(ws), 256));
  double* area = reinterpret_cast<double*>(p); p += align_up((size_t)F * 8, 256);
  double* bsum = reinterpret_cast<double*>(p); p += align_up((((size_t)F + 1023) / 1024) * 8, 256);
  double* total = reinterpret_cast<double*>(p);
  face_area_kernel<<<grid_for(F), 256, 0, st>>>(verts, faces, F, vscale, voffset, area);
  exclusive_scan<double>(area, area, bsum, F, total, st);
  count_launches(4);
  mesh_sample_kernel<<<grid_for(S), 256, 0, st>>>(verts, faces, F, vscale, voffset, area, total, S, seed, points);
  ZS_CUDA_CHECK_LAUNCH("zs_mesh_sample");
  return ZS_OK;
}
