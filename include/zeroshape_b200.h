/*
 * zeroshape_b200 -- C ABI of the B200 (sm_100a) implementation of ZeroShape's per-image hot path.
 *
 * This header is the drop-in boundary.  The reference (zxhuang1698/ZeroShape) is pure Python over
 * PyTorch plus one pybind11 CUDA extension; it has no FFI registry, so each entry point below cites
 * the reference Python/CUDA function (file:line under the reference tree) whose arithmetic it
 * replaces.  The Python host package `zeroshape_b200` binds these symbols with ctypes and re-exposes
 * the reference's own module surface (model.compute_graph.graph_shape.Graph, utils.eval_3D.*,
 * external.chamfer3D.dist_chamfer_3D.chamfer_3DDist) -- see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - tensors are dense, row-major, fp32 unless stated; image tensors are NHWC inside the library;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); nothing synchronises;
 *   - nothing allocates: callers pass outputs and (where stated) a workspace;
 *   - return value: 0 = ok, negative = error (message via zs_last_error(), thread-local);
 *   - all functions are re-entrant and may be called from several host threads.
 */
#ifndef ZEROSHAPE_B200_H
#define ZEROSHAPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZS_OK 0
#define ZS_ERR_ARG (-1)
#define ZS_ERR_CUDA (-2)
#define ZS_ERR_UNSUPPORTED (-3)

/* activation codes for fused epilogues */
#define ZS_ACT_NONE 0
#define ZS_ACT_RELU 1
#define ZS_ACT_GELU 2      /* exact erf GELU (torch.nn.GELU default) */
#define ZS_ACT_SOFTPLUS100 3 /* torch Softplus(beta=100, threshold=20): model/shape/implicit.py:166 */
#define ZS_ACT_SIGMOID 4
#define ZS_ACT_CLAMP01 5    /* relu then clamp(max=1): DPT depth head, model/depth/dpt_depth.py:106,119 */

/* residual placement in the fused epilogue: y = act(acc + bias [+ res]) [+ res] */
#define ZS_RES_NONE 0
#define ZS_RES_BEFORE_ACT 1
#define ZS_RES_AFTER_ACT 2

const char* zs_last_error(void);
int zs_abi_version(void);
/* compute capability (major*10+minor) of the current device, or negative error (library plumbing: no reference counterpart;
 * the reference asserts a CUDA device at utils/options.py:101) */
int zs_device_cc(void);
/* number of CUDA kernels this library has launched since it was loaded (all threads) */
long long zs_launch_count(void);
/* Adds n to that counter: a CUDA graph replays its launches without passing through these entry points; the owner of the graph
 * adds the count recorded at capture once per replay. */
void zs_launch_count_add(long long n);

/* ------------------------------------------------------------------------------------------------
 * Dense building blocks (replace the cuBLAS/cuDNN calls PyTorch issues for nn.Linear / nn.Conv2d /
 * LayerNorm / GroupNorm / F.interpolate / MaxPool on the hot path; SURVEY.md section 2.3 last row).
 * ---------------------------------------------------------------------------------------------- */

/* C[M,N] = epilogue( A[M,K] * W[N,K]^T ), fp32 accumulate in fp32 FFMA (bit-faithful path).
 * Replaces F.linear (e.g. model/shape/implicit.py:30,74,178-181; timm Block qkv/proj/fc1/fc2). */
int zs_gemm_f32(const float* A, int lda, const float* W, int ldw, const float* bias,
                const float* res, int ldres, int res_mode, float* C, int ldc,
                int M, int N, int K, int act, void* stream);

/* Tensor-core GEMM (tcgen05.mma, TMEM accumulators, TMA bulk-copied weights) with the same epilogue: the F.linear calls above
 * (model/shape/implicit.py:30,74,178-181; timm Block qkv / proj / fc1 / fc2 under model/depth/vit.py:149-150).
 * Weights are packed once (fp32 W[N,K] -> (hi,lo) bf16 tiles in the UMMA 128B-swizzled K-major smem image).
 * precision 0 = "bf16x3" (Ah*Wh + Ah*Wl + Al*Wh, ~2^-16 relative: parity mode), 1 = "bf16" (single pass). */
size_t zs_gemm_tc_packed_bytes(int N, int K);
int zs_gemm_tc_pack(const float* W, int ldw, int N, int K, void* packed, void* stream);
/* same image with the 16-bit operand format chosen: fmt 0 = bf16 (zs_gemm_tc_f32, convolutions, training GEMMs),
 * fmt 1 = fp16 (every zs_chain_* kernel of the decoder: hi/lo fp16 carries 22 significand bits, 8x tighter than the
 * bf16 split at the same MMA count -- profiles/r2_precision_study.md) */
int zs_gemm_tc_pack_fmt(const float* W, int ldw, int N, int K, void* packed, int fmt, void* stream);
int zs_gemm_tc_f32(const float* A, int lda, const void* Wpacked, const float* bias,
                   const float* res, int ldres, int res_mode, float* C, int ldc,
                   int M, int N, int K, int act, int precision, void* stream);

/* Point-to-latent attention on the tensor cores (model/shape/implicit.py:38-57), two grouped tcgen05 launches:
 *   zs_attn_scores_tc: P[m, 208h + j] = exp(scale*(q_h(m).k_lat_h(j) - rowmax)) (softmax numerators over the latent keys
 *                      j < n_keys; the row max / sum also cover the point's own key), Rinv[m, h] = 1/sum,
 *                      R[m, 32h + d] = p_self * v_p,h[d]                               (qkv [M,768] fp32 = q|k|v)
 *   zs_attn_pv_tc    : O[m, 32h + d] = Rinv[m,h] * sum_j P[m, 208h + j] * v_lat_h[j, d] + R[m, 32h + d]
 * Kpacked: 8 head tiles from zs_gemm_tc_pack(K_lat_h [208 rows (n_keys real, rest 0), 32]);
 * Vpacked: 8 heads x 4 K-chunks from zs_gemm_tc_pack(V_lat_h^T [32, 208]).  P is [M, 1664], R and O [M, 256], Rinv [M, 8]. */
int zs_attn_scores_tc(const float* qkv, int ld_qkv, const void* Kpacked, int M, int n_keys, float scale,
                      float* P, float* R, float* Rinv, int precision, void* stream);
int zs_attn_pv_tc(const float* P, const void* Vpacked, const float* R, const float* Rinv, float* O, int M, int precision,
                  void* stream);

/* Same attention, flash-style in ONE kernel (scores, softmax numerators and P.V of a 128-point tile stay on the SM;
 * csrc/chain_tc.cu chain_attn_kernel).  Operand images of every zs_chain_* kernel are fp16 (zs_gemm_tc_pack_fmt, fmt 1);
 * their `precision` 0 = "fp16x3" (Ah*Wh + Al*Wh + Ah*Wl, ~2^-22 relative: parity mode), 1 = "fp16" (single pass).
 * The 8 heads are processed as 4 pairs (2p, 2p+1):
 *   Kblob: 4 pair tiles from zs_gemm_tc_pack_fmt(K_lat[:, 64p:64p+64] zero-padded to 256 rows) = 4 x [hi 32 KB | lo 32 KB]
 *          (the two heads' 32 dims side by side, keys along the rows);
 *   Vblob: 8 heads x 32 KB = 4 key-chunks x [hi 4 KB | lo 4 KB] (the first 4 KB of each hi / lo tile of Vpacked above).
 * n_keys <= 208.  O [M,256]. */
int zs_chain_attn_fwd(const float* qkv, int ld_qkv, int M, const void* Kblob, const void* Vblob, int n_keys,
                      float scale, float* O, int precision, void* stream);
/* LayerNorm + qkv + the same attention in ONE kernel: norm1 of ImplFuncBlock and ImplFuncAttention up to the output
 * projection (model/shape/implicit.py:105, 30-57; csrc/chain_tc.cu chain_qkvattn_kernel).  q, k, v of the query points,
 * scores and probabilities stay on the SM: HBM sees x [M,256] (read) and O [M,256] (written).
 *   Wblob   : zs_chain_qkvattn_blob_bytes() = 4 pairs x zs_gemm_tc_pack_fmt(Wp [256 rows, 256], fmt 1) where the rows of Wp
 *             are [q rows of heads 2p, 2p+1 (64) | their k rows (64) | their v rows (64) | 64 zero rows] of
 *             qkv.weight * norm1.weight (the LayerNorm affine folded in);
 *   bias_qkv: [768] = qkv.bias + qkv.weight @ norm1.bias, original q | k | v order;
 *   Kblob / Vblob / n_keys / scale: as zs_chain_attn_fwd.
 *   flags   : per-GEMM pass policy on top of precision 0: 1 = k, v columns single-pass, 2 = scores without Qh*Kl,
 *             4 = P*V without Ph*Vl (0 = every contraction three passes); 8 = keep the probabilities in tensor memory
 *             (chain_qkvattn2_kernel: P overwrites the scores in place, P*V reads its A operand from TMEM);
 *             16 (with 8) = the softmax role reads every score from tensor memory ONCE and keeps it in registers, and O is
 *             written TILE-BLOCKED: the 16-byte chunk j (columns 4j..4j+3) of row r of 128-row tile t is float4 number
 *             (t * 64 + j) * 128 + r, ldo is ignored and O must hold ceil(M / 128) * 128 * 256 floats (padded rows are
 *             written too).  zs_chain_lin_fwd consumes that layout with do_ln = 2. */
size_t zs_chain_qkvattn_blob_bytes(void);
int zs_chain_qkvattn_fwd(const float* x, int ldx, int M, float ln_eps, const void* Wblob, const float* bias_qkv,
                         const void* Kblob, const void* Vblob, int n_keys, float scale, float* O, int ldo,
                         int precision, int flags, void* stream);
/* debug: while buf (device, [3][512] uint64) is non-NULL the chained kernels record role-level clock64 events of
 * CTA 0 (MMA thread, loader thread 0, epilogue warp 4 lane 0) into it; see tools/trace_chain.py.  Not thread-safe. */
int zs_debug_chain_trace(unsigned long long* buf);
/* debug / A-B: zs_chain_mlp_fwd and zs_chain_occ_fwd run chain_mlp2 / chain_occ2 (activations of the next GEMM kept in tensor
 * memory, 4 / 5 weight-ring slots) when v != 0 (default) and the round-1 kernels (shared-memory ring E, 3 slots) when v == 0.
 * Identical arithmetic.  Process-wide, not thread-safe. */
int zs_debug_chain_variant(int v);
/* debug / A-B: 0 disables the split-K path that zs_gemm_tc_f32 / zs_conv2d_nhwc_tc take for few-tile layers (tiles * 2 <= SMs and
 * K >= 512: up to 16 K-splits per tile, partial tiles in a stream-ordered workspace, deterministic finalize pass).  Process-wide. */
int zs_debug_gemm_splitk(int enable);

/* Chained tcgen05 kernels of the implicit decoder (consecutive layers of a 128-point tile stay on chip; see
 * csrc/chain_tc.cu).  `blob` = weight tiles in consumption order, each sub-matrix packed with zs_gemm_tc_pack
 * (N=256 rows, K padded to 64) and concatenated:
 *   mlp: for g in 0..3: fc1.weight[256g:256g+256, :] (K=256), fc2.weight[:, 256g:256g+256] (K=256)
 *   occ: layer l = 0..7 of impl_mlp with K reordered to [feat(256) | xyz(3)] for the `inputs` part and the
 *        skip layers' 1/sqrt(2) folded in:  l0: [W0[:,3:259] | W0[:,0:3]] ; odd l: W_l ;
 *        l in {2,4,6}: [W_l[:,259:515] | W_l[:,256:259]]/sqrt2 then W_l[:,0:256]/sqrt2.
 * zs_chain_mlp_fwd:  x <- x + fc2(GELU(fc1(LayerNorm(x))))           (model/shape/implicit.py:94-108, timm Mlp)
 * zs_chain_occ_fwd:  out = MLPBlocks([xyz, LayerNorm(x)]) (+sigmoid)  (model/shape/implicit.py:275,168-184) */
/* zs_chain_lin_fwd:  out[M, 256*n_tiles] = LN?(x)[M,256] W^T + bias (+ res)   (qkv / proj of ImplFuncAttention,
 * model/shape/implicit.py:30,74; do_ln = the block's norm1 statistics computed in-kernel, affine folded into W / bias).
 * `blob` = zs_gemm_tc_pack image of W[256*n_tiles, 256].  `res`/`out` may alias (in-place residual update).
 * do_ln: 0 = x as is, 1 = LayerNorm(x), 2 = x is the tile-blocked attention output of zs_chain_qkvattn_fwd (flags & 16;
 * ldx ignored, n_tiles must be 1, no LayerNorm). */
int zs_chain_lin_fwd(const float* x, int ldx, int M, int do_ln, float ln_eps, const void* blob, int n_tiles,
                     const float* bias, const float* res, int ldres, float* out, int ldo, int precision, void* stream);
size_t zs_chain_mlp_blob_bytes(void);
size_t zs_chain_occ_blob_bytes(void);
int zs_chain_mlp_fwd(float* x, int ldx, int M, const float* ln_w, const float* ln_b, float ln_eps,
                     const void* blob, const float* b1, const float* b2, int precision, void* stream);
/* zs_chain_pmlp_fwd: x <- x' + fc2(GELU(fc1(LayerNorm(x')))) with x' = x + A proj.weight^T + proj.bias: the attention
 * output projection + residual (model/shape/implicit.py:74, ImplFuncBlock :105) fused in front of zs_chain_mlp_fwd.
 * a_blk = the tile-blocked attention output of zs_chain_qkvattn_fwd (flags & 16), ceil(M/128)*128*256 floats;
 * proj_blob = zs_gemm_tc_pack_fmt(proj.weight [256,256], fmt 1); mlp_blob / b1 / b2 as zs_chain_mlp_fwd (norm2's affine
 * folded into fc1).  The LayerNorm variance is the one-pass E[x^2] - mean^2 of pairwise-reduced fp32 sums.
 * Points mode (points, pp both non-null; the FIRST decoder block): the incoming residual stream is x = LinearProj3D(points)
 * (model/shape/implicit.py:128-131) and is recomputed from the [M,3] points instead of being read: pp = [4][256] floats =
 * point_proj.weight[:,0] | [:,1] | [:,2] | point_proj.bias; x is then written only (x' and the result). */
int zs_chain_pmlp_fwd(float* x, int ldx, int M, const float* a_blk, const void* proj_blob, const float* proj_bias,
                      float ln_eps, const void* mlp_blob, const float* b1, const float* b2, const float* points,
                      const float* pp, int precision, void* stream);
/* zs_chain_qkvattn_fwd (flags 8 | 16) for the first block in points mode: LayerNorm(LinearProj3D(points)) is formed in the
 * loader warps from the points, pp (as above) and pp_stat = 14 floats: the means m0 m1 m2 mb of the four rows of pp over the
 * 256 channels, then their (biased) covariances C00 C11 C22 Cbb, C01 C02 C12, C0b C1b C2b -- the LayerNorm mean and variance
 * of a row are a linear / quadratic form of its point.  No [M,256] input exists at all. */
int zs_chain_qkvattn_pts_fwd(const float* points, int M, const float* pp, const float* pp_stat, float ln_eps, const void* Wblob,
                             const float* bias_qkv, const void* Kblob, const void* Vblob, int n_keys, float scale, float* O,
                             int precision, int flags, void* stream);
int zs_chain_occ_fwd(const float* x, int ldx, const float* points, int M, const float* ln_w, const float* ln_b,
                     float ln_eps, const void* blob, const float* biases, const float* w8, float b8,
                     float* out, int apply_sigmoid, int precision, void* stream);

/* NHWC convolution as implicit GEMM: y[b,oh,ow,co] = epi( sum x[b,oh*s+kh-pt,ow*s+kw-pl,ci] w[co,kh,kw,ci] ).
 * `pre_relu` applies ReLU to x on load (ResidualConvUnit_custom: model/depth/blocks.py:274-281).
 * Replaces nn.Conv2d in model/depth/blocks.py:58-70,247-253,305, dpt_depth.py:100-108,
 * utils/layers.py:79-82, torchvision resnet50 and the timm ResNetV2 stem/stages. */
int zs_conv2d_nhwc_f32(const float* x, int B, int H, int W, int Cin,
                       const float* w, const float* bias, const float* res, int res_mode,
                       float* y, int Cout, int KH, int KW, int stride, int pad_top, int pad_left,
                       int OH, int OW, int act, int pre_relu, void* stream);

/* The same convolution on the tensor cores: gemm_tc_kernel with an im2col-on-the-fly A producer (the K-chunks of an
 * output pixel are 256-byte runs of one input pixel when Cin % 64 == 0).  `Wpacked` = zs_gemm_tc_pack of the OHWI filter
 * viewed as W[Cout, KH*KW*Cin].  Needs Cin % 4 == 0 (every conv of the path except the two RGB/XYZ stems). */
int zs_conv2d_nhwc_tc(const float* x, int B, int H, int W, int Cin, const void* Wpacked, const float* bias,
                      const float* res, int res_mode, float* y, int Cout, int KH, int KW, int stride,
                      int pad_top, int pad_left, int OH, int OW, int act, int pre_relu, int precision, void* stream);

/* LayerNorm over the last dim (eps given).  Replaces nn.LayerNorm in timm Block / implicit.py:89,95,225 */
int zs_layernorm_f32(const float* x, int ldx, const float* gamma, const float* beta, float* y, int ldy,
                     int rows, int cols, float eps, void* stream);

/* GroupNorm (+optional ReLU) on NHWC.  Replaces timm GroupNormAct in the ResNetV2 stem/stages of vit_base_resnet50_384
 * (created at model/depth/vit.py:482; timm itself is third-party). */
int zs_groupnorm_nhwc_f32(const float* x, const float* gamma, const float* beta, const float* res, float* y,
                          int B, int HW, int C, int groups, float eps, int relu, void* stream);
/* The same GroupNorm on the tiled kernels (two fully coalesced launches: per-chunk group statistics in double to `ws`, then the
 * normalisation) when C / 4 is a power of two <= 256 and the tensors are 16-byte aligned; any other shape (or ws == NULL) takes
 * zs_groupnorm_nhwc_f32.  ws: zs_groupnorm_ws_bytes(B, HW, C, groups) bytes, 8-byte aligned, need not be initialised. */
size_t zs_groupnorm_ws_bytes(int B, int HW, int C, int groups);
int zs_groupnorm_nhwc_ws_f32(const float* x, const float* gamma, const float* beta, const float* res, float* y,
                             int B, int HW, int C, int groups, float eps, int relu, void* ws, void* stream);

/* Per-channel affine y = act(x*scale[c] + shift[c] (+res)) on [rows, C]: eval-mode BatchNorm of Bottleneck_Conv
 * (utils/layers.py:76-100) and of the torchvision ResNet-50 in CoordEncRes (model/shape/seen_coord_enc.py:141-194). */
int zs_channel_affine_f32(const float* x, const float* scale, const float* shift, const float* res,
                          float* y, int64_t rows, int C, int act, void* stream);

/* Elementwise: y = act(a*alpha + b*beta) (b may be NULL): residual adds / scalings such as model/depth/blocks.py:285,331,
 * utils/util.py:345 (coordmap / (1 + 1e-6)), the final sigmoid of utils/eval_3D.py:43. */
int zs_axpby_f32(const float* a, float alpha, const float* b, float beta, float* y, int64_t n, int act, void* stream);

/* Max-pool 3x3 stride 2 on NHWC with explicit top/left padding (pad value -inf): timm MaxPool2dSame of the ResNetV2 stem and
 * torchvision resnet50.maxpool (model/shape/seen_coord_enc.py:148-160). */
int zs_maxpool3x3s2_nhwc_f32(const float* x, float* y, int B, int H, int W, int C, int pad_top, int pad_left,
                             int OH, int OW, void* stream);

/* Global average pool NHWC [B,HW,C] -> [B,C].  (nn.AdaptiveAvgPool2d((1,1)), graph_shape.py:25) */
int zs_avgpool_nhwc_f32(const float* x, float* y, int B, int HW, int C, void* stream);

/* Bilinear resize NHWC (align_corners 0/1).  model/depth/blocks.py:336-338, dpt_depth.py:102,
 * model/depth/vit.py:110 (pos-embed), utils/util.py:341-342. */
int zs_bilinear_nhwc_f32(const float* x, float* y, int B, int H, int W, int C, int OH, int OW,
                         int align_corners, void* stream);

/* Input pipeline (csrc/preprocess.cu; reference demo.py:33-75, data/synthetic.py:178-210).
 * zs_rgba_crop_resize_u8: PIL crop of the RGBA image [H0,W0,4] to the window (left, top, cw, ch) -- zeros outside the image, as
 *   torchvision_F.crop on a PIL image -- then PIL `Image.resize((OW, OH))` with its default BICUBIC filter on an RGBA image
 *   (premultiplied alpha, two 8-bit passes with 22-bit fixed-point coefficients).  bounds [out][2] = (first source index, taps),
 *   coef [out][ksize] int32 = Pillow's normalized coefficients (host: zeroshape_b200/data/preprocess.py:resize_coeffs).  Byte-exact.
 * zs_rgba_composite_f32: to_tensor + `rgb * mask + bgcolor * (1 - mask)`, `mask > 0.5` (demo.py:46-52) -> rgb [3,H,W], mask [1,H,W].
 * zs_erode_square_f32: cv2.erode(mask, ones(3,3), iterations=radius) of demo.py:70-75 (minimum over the clipped square window). */
int zs_rgba_crop_resize_u8(const uint8_t* src, int H0, int W0, int left, int top, int cw, int ch, uint8_t* out, int OH, int OW,
                           const int* xbounds, const int* xcoef, int xksize, const int* ybounds, const int* ycoef, int yksize,
                           void* stream);
int zs_rgba_composite_f32(const uint8_t* img, int H, int W, int use_bgcolor, float bgcolor, float* rgb, float* mask, void* stream);
int zs_erode_square_f32(const float* mask, float* out, int B, int H, int W, int radius, void* stream);

/* NCHW <-> NHWC (the reference keeps NCHW throughout; scale / shift fold `image * 2 - 1` of model/depth/dpt_depth.py:116) */
int zs_nchw_to_nhwc_f32(const float* x, float* y, int B, int C, int H, int W, float scale, float shift, void* stream);
int zs_nhwc_to_nchw_f32(const float* x, float* y, int B, int C, int H, int W, void* stream);

/* Multi-head self-attention over T tokens from a packed qkv buffer [B,T,3,heads,hd] -> out [B,T,heads*hd].
 * softmax(q k^T * scale) v.  Replaces timm Attention.forward and the latent branch implicit.py:65-71. */
int zs_mha_f32(const float* qkv, float* out, int B, int T, int heads, int hd, float scale, void* stream);
/* The same operation on the tensor cores (csrc/mha_tc.cu): S = Q K^T and O = P V as tcgen05 MMAs with split-fp16 operands
 * (precision 0: three passes, fp32-grade; 1: one fp16 pass), fp32 accumulation and the probabilities kept in tensor memory.
 * T <= 208 tokens, hd 32 or 64.  The 12 ViT blocks of the DPT-hybrid backbone (timm Block, model/depth/vit.py:149-150). */
int zs_mha_tc_f32(const float* qkv, float* out, int B, int T, int heads, int hd, float scale, int precision, void* stream);
/* Backward of the same operation on the tensor cores (csrc/mha_tc.cu: mha_bwd_q_kernel, mha_bwd_kv_kernel): S / dP and their
 * transposes as tcgen05 MMAs with single-pass fp16 operands and fp32 accumulation, P and dS re-written in place in tensor memory
 * as the A operands of dQ = dS K, dV = P^T dO, dK = dS^T Q.  Precision class of the bf16 training mode (zs_mha_bwd_f32 is the
 * fp32-grade path).  dqkv [B,T,3C] is written completely; ws: zs_mha_bwd_tc_ws_bytes(B, T, heads), 16-byte aligned. */
size_t zs_mha_bwd_tc_ws_bytes(int B, int T, int heads);
int zs_mha_bwd_tc_f32(const float* qkv, const float* dO, float* dqkv, int B, int T, int heads, int hd, float scale, void* ws,
                      void* stream);
/* zs_point_attention_f32 on the tensor cores for the training tape (csrc/mha_tc.cu: pa_fwd_tc_kernel; head dim 32, L <= 208, no
 * attention-map output): per 128-point tile S = Q K_lat^T and O = P V_lat as tcgen05 MMAs with the probabilities in tensor memory,
 * the point's own key / value as one extra softmax column in registers.  precision 0 = split fp16 (three passes), 1 = one fp16 pass. */
int zs_point_attention_tc_f32(const float* qkv_p, const float* k_lat, const float* v_lat, int ld_lat, float* out, int B, int P, int L,
                              int heads, int hd, float scale, int precision, void* stream);
/* Backward of zs_point_attention_f32 on the tensor cores (csrc/mha_tc.cu: pa_bwd_q_kernel per 128-point tile, pa_bwd_kv_kernel
 * per 128 latent keys looping over the points in chunks of 208; the point's own key / value handled per row): same outputs as
 * zs_point_attention_bwd_f32 (dqkv_p [B,P,3C] complete, dk_lat / dv_lat [B,L,C] with row stride ld_dlat), single fp16 pass.
 * head dim 32, L <= 208; ws: zs_point_attention_bwd_tc_ws_bytes(B, P, heads), 16-byte aligned. */
size_t zs_point_attention_bwd_tc_ws_bytes(int B, int P, int heads);
int zs_point_attention_bwd_tc_f32(const float* qkv_p, const float* k_lat, const float* v_lat, int ld_lat, const float* dO,
                                  float* dqkv_p, float* dk_lat, float* dv_lat, int ld_dlat, int B, int P, int L, int heads, int hd,
                                  float scale, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Implicit decoder (model/shape/implicit.py:251-288), query-point side.
 * ---------------------------------------------------------------------------------------------- */

/* Point-to-latent attention of ImplFuncAttention (implicit.py:38-57): each query point attends to the
 * L latent tokens and to itself.  qkv_p [B,P,3,heads,hd]; k_lat,v_lat [B,L,heads*hd] (row stride ld_lat).
 * out [B,P,heads*hd].  attn (optional) [B,P,L] accumulates `attn_scale` * mean-over-heads probability. */
int zs_point_attention_f32(const float* qkv_p, const float* k_lat, const float* v_lat, int ld_lat,
                           float* out, float* attn, float attn_scale, int attn_accumulate,
                           int B, int P, int L, int heads, int hd, float scale, void* stream);

/* Dense query grid of utils/eval_3D.py:10-20 for x-slices [x0,x1): out [x1-x0, n, n, 3], n = N+1,
 * coordinates = linspace(rmin,rmax,n) exactly as torch.linspace computes them. */
int zs_dense_grid_f32(float* out, int n, float rmin, float rmax, int x0, int x1, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training step of the implicit decoder (decoder slice): fp32 backward kernels.
 * Replace torch autograd over model/shape/implicit.py for the call at model/compute_graph/graph_shape.py:185,
 * utils/loss.py:18-28 (shape_loss) and torch.optim.AdamW (model/shape_engine.py:132, 276).
 * ---------------------------------------------------------------------------------------------- */
/* loss = mean_i w_i * BCEWithLogits(logits_i, sdf_i < 0), w_i = impt_weight where |sdf_i| < impt_thres else 1.
 * `ws` = one double of scratch; `loss` = one float on the device. */
int zs_bce_logits_fwd(const float* logits, const float* sdf, int64_t n, float impt_thres, float impt_weight,
                      double* ws, float* loss, void* stream);
/* dlogits_i = grad_scale * w_i * (sigmoid(logits_i) - y_i) / n */
int zs_bce_logits_bwd(const float* logits, const float* sdf, int64_t n, float impt_thres, float impt_weight,
                      float grad_scale, float* dlogits, void* stream);
/* dx = dy * act'(z) for the pre-activation z (ZS_ACT_GELU exact-erf, ZS_ACT_SOFTPLUS100, ZS_ACT_RELU, ZS_ACT_NONE) */
int zs_act_bwd_f32(const float* dy, const float* z, float* dx, int64_t n, int act, void* stream);
/* out[n] (+)= sum_m A[m,n]  (bias gradients) */
int zs_colsum_f32(const float* A, int lda, int64_t M, int N, float* out, int accumulate, void* stream);
/* C[N,K] (+)= A[M,N]^T B[M,K]  (weight gradients dW = dY^T X) */
int zs_gemm_tn_f32(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int64_t M, int N, int K,
                   int accumulate, void* stream);
/* LayerNorm backward over 256 columns; dgamma/dbeta (optional, both or neither) are ACCUMULATED into. */
int zs_layernorm_bwd_f32(const float* dy, const float* x, const float* gamma, float eps, float* dx, float* dgamma,
                         float* dbeta, int64_t rows, int cols, void* stream);
/* Backward of zs_point_attention_f32 (head dim 32): O = forward output, dO its gradient; writes dqkv_p [B,P,3C] (all of it)
 * and the latent-side gradients dk_lat / dv_lat [B,L,C] (row stride ld_dlat). */
int zs_point_attention_bwd_f32(const float* qkv_p, const float* k_lat, const float* v_lat, int ld_lat, const float* O,
                               const float* dO, float* dqkv_p, float* dk_lat, float* dv_lat, int ld_dlat, int B, int P,
                               int L, int heads, int hd, float scale, void* stream);
/* Backward of zs_mha_f32: dqkv [B,T,3C] from qkv and dO [B,T,C] (head dim <= 64).  Two passes without atomics: the
 * probability and score-gradient rows go through `ws` (zs_mha_bwd_ws_bytes = 2 * B * heads * T * T floats, 16-byte aligned). */
size_t zs_mha_bwd_ws_bytes(int B, int T, int heads);
int zs_mha_bwd_f32(const float* qkv, const float* dO, float* dqkv, int B, int T, int heads, int hd, float scale, void* ws,
                   void* stream);
/* Front end of the transformer seen-surface encoder (CoordEmb.forward, model/shape/seen_coord_enc.py:49-72): per-pixel
 * Linear(3 -> C) of the XYZ map coord [B,H,W,3], `invalid` token where mask [B,H,W] <= 0.5, ws x ws window partition, the fixed
 * 2-D sin-cos embedding pos [ws*ws+1, C] local to each window and the cls row -> out [B*(H/ws)*(W/ws), ws*ws+1, C]. */
int zs_coord_embed_windows_f32(const float* coord, const float* mask, const float* w, const float* bias, const float* invalid,
                               const float* pos, const float* cls, float* out, int B, int H, int W, int C, int ws, void* stream);
/* DepthMetric.compute_metrics (utils/eval_depth.py:41-110): per image, least-squares scale / shift of the predicted disparity
 * (1 / (pred + 1e-6), or pred itself when disparity_input) to 1 / gt over mask > 0.5, aligned depth = 1 / max(aligned disparity,
 * 1 / depth_cap) (depth_cap <= 0: no cap), then metrics [B, T + 3] = {fraction with max(d/g, g/d) > thresholds[k]} (thresholds: T <= 8
 * floats in device memory), RMSE, L1, absolute relative error; depth_out [B,1,H,W] = the aligned depth map (also outside the mask). */
int zs_depth_metrics_f32(const float* pred, const float* gt, const float* mask, int B, int H, int W, const float* thresholds,
                         int T, float depth_cap, int disparity_input, float* metrics, float* depth_out, void* stream);
/* MidasLoss.erode_mask (midas_loss.py:158-167, `training.depth_loss.mask_shrink`): out = 1 where a whole pool x pool block of the raw
 * mask [B,1,H,W] equals 1 (1 - mask -> max_pool2d -> nearest upsampling -> == 0), else 0. */
int zs_mask_erode_f32(const float* mask, float* out, int B, int H, int W, int pool, void* stream);
/* MiDaS scale-and-shift-invariant depth loss (model/depth/midas_loss.py:145-185 as configured by utils/loss.py:14-16,30-34:
 * SSI-MAE on median / mean-absolute-deviation aligned maps + alpha x gradient matching over 4 scales of the least-squares aligned
 * (inverse) depth, image-based reduction, valid = mask > 0.5) and its gradient w.r.t. the prediction.
 * pred, gt, mask: [B, 1, H, W] fp32 contiguous; `loss`: one float on the device; `dpred` (optional): grad_scale * d loss / d pred;
 * `ws`: zs_midas_ws_bytes(B, H, W), 8-byte aligned.  Three launches, no host sync. */
size_t zs_midas_ws_bytes(int B, int H, int W);
int zs_midas_loss_f32(const float* pred, const float* gt, const float* mask, int B, int H, int W, float alpha,
                      int inverse_depth, float grad_scale, void* ws, float* loss, float* dpred, void* stream);
/* torch.optim.AdamW step (decoupled weight decay, bias correction) on one flat tensor; `step` counts from 1. */
int zs_adamw_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                 float beta2, float eps, float weight_decay, int step, void* stream);
/* The same step for every parameter tensor of the optimizer in one launch.  `table` (device memory) = n_tensors rows of five
 * 64-bit words {param, grad, exp_avg, exp_avg_sq (device addresses), numel}; identical arithmetic to zs_adamw_f32. */
int zs_adamw_multi_f32(const void* table, int n_tensors, float lr, float beta1, float beta2, float eps, float weight_decay,
                       int step, void* stream);
/* The same update with every step-dependent scalar in device memory: hyper[7] = {lr, beta1, beta2, eps, weight_decay,
 * 1 - beta1^step, sqrt(1 - beta2^step)}.  `host_table` is the same table in HOST memory: its rows travel by value in the kernel
 * parameters (96 tensors per launch), so the call reads no host buffer at run time and is capturable in a CUDA graph; the host
 * refreshes `hyper` with a stream-ordered copy before each replay (zeroshape_b200/graphed.py). */
int zs_adamw_multi_dev_f32(const void* host_table, int n_tensors, const float* hyper, void* stream);

/* Training of the seen-surface encoder (CoordEncRes: torchvision ResNet-50 + Bottleneck_Conv heads with batch-statistics
 * BatchNorm, model/shape/seen_coord_enc.py:141-194; the `optim.fix_dpt` configuration of options/shape.yaml).
 * zs_conv2d_nhwc_dgrad_f32: dx [B,H,W,Cin] from dy [B,OH,OW,Cout]; `w_dgrad` = the OHWI filter re-laid as [Cin][KH][KW][Cout].
 * zs_conv2d_nhwc_wgrad_f32: dw [Cout,KH,KW,Cin] (+)= sum over output pixels of dy x im2col(x). */
int zs_conv2d_nhwc_dgrad_f32(const float* dy, int B, int H, int W, int Cin, const float* w_dgrad, float* dx, int Cout,
                             int KH, int KW, int stride, int pad_top, int pad_left, int OH, int OW, void* stream);
int zs_conv2d_nhwc_wgrad_f32(const float* x, int B, int H, int W, int Cin, const float* dy, float* dw, int Cout, int KH,
                             int KW, int stride, int pad_top, int pad_left, int OH, int OW, int accumulate, void* stream);

/* The three gradient GEMMs of the training step on the tensor cores (tcgen05, split-bf16 operands, fp32 TMEM accumulators;
 * `precision` 0 = bf16x3, 1 = bf16 as in zs_gemm_tc_f32).  They replace what torch autograd runs for every nn.Linear /
 * nn.Conv2d in `loss.backward()` of the reference (model/shape_engine.py:268-272):
 *  - dX = dY W        : zs_gemm_tc_f32 with zs_gemm_tc_pack(W^T)  (no new entry point)
 *  - zs_conv2d_nhwc_dgrad_tc: dx [B,H,W,Cin] from dy [B,OH,OW,Cout] (Cout % 4 == 0); `Wpacked` = zs_gemm_tc_pack of the
 *    filter re-laid as Wd[Cin, KH*KW*Cout] (the operand of zs_conv2d_nhwc_dgrad_f32)
 *  - zs_gemm_tn_tc          : C[N,K] (+)= A[M,N]^T B[M,K]  (dW = dY^T X), split over the rows, fp32 reductions into C
 *  - zs_conv2d_nhwc_wgrad_tc: dw [Cout,KH,KW,Cin] (+)= dY^T im2col(x), the im2col gathered on the fly (Cin % 8 == 0)
 * `layout` names the shared-memory operand layout of the TN kernels and must be 0: MN-major SWIZZLE_128B tiles (the reduction
 * runs over the rows of both fp32 matrices, so no transposition is needed anywhere). */
int zs_conv2d_nhwc_dgrad_tc(const float* dy, int B, int H, int W, int Cin, const void* Wpacked, float* dx, int Cout,
                            int KH, int KW, int stride, int pad_top, int pad_left, int OH, int OW, int precision, void* stream);
int zs_gemm_tn_tc(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int64_t M, int N, int K,
                  int accumulate, int precision, int layout, void* stream);
int zs_conv2d_nhwc_wgrad_tc(const float* x, int B, int H, int W, int Cin, const float* dy, float* dw, int Cout, int KH,
                            int KW, int stride, int pad_top, int pad_left, int OH, int OW, int accumulate, int precision,
                            int layout, void* stream);
/* per-channel batch statistics of x [M,C]: mean, biased variance, rstd = 1/sqrt(var + eps); `ws` = 2*C doubles */
int zs_bn_stats_f32(const float* x, int64_t M, int C, float eps, double* ws, float* mean, float* var, float* rstd, void* stream);
/* BatchNorm (batch statistics) backward: dx, and dgamma / dbeta ACCUMULATED into; `ws` = 2*C doubles */
int zs_bn_bwd_f32(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int64_t M,
                  int C, double* ws, float* dx, float* dgamma, float* dbeta, void* stream);
int zs_maxpool3x3s2_bwd_nhwc_f32(const float* x, const float* dy, float* dx, int B, int H, int W, int C, int pad_top,
                                 int pad_left, int OH, int OW, void* stream);
int zs_avgpool_bwd_nhwc_f32(const float* dy, float* dx, int B, int HW, int C, void* stream);

/* Backward kernels of the depth estimator's layers (DPT-hybrid: timm ResNetV2 GroupNorm, ViT LayerNorm over 768 columns,
 * align_corners bilinear upsampling of model/depth/blocks.py:321-342) and of the geometry glue (utils/camera.py:52-108). */
int zs_coldot_f32(const float* A, int lda, const float* B2, int ldb, int64_t M, int N, float* out, int accumulate, void* stream);
int zs_layernorm_bwd_generic_f32(const float* dy, const float* x, const float* gamma, float eps, float* dx, float* xhat,
                                 int64_t rows, int cols, void* stream);
int zs_groupnorm_bwd_nhwc_f32(const float* dy, const float* x, const float* gamma, float* dx, float* dgamma, float* dbeta,
                              int B, int HW, int C, int groups, float eps, void* stream);
int zs_bilinear_bwd_nhwc_f32(const float* dy, float* dx, int B, int H, int W, int C, int OH, int OW, int align_corners, void* stream);
/* ddepth [B,H,W] and dKinv [B,3,3] (gradient w.r.t. the INVERSE intrinsics) from dseen [B,HW,3]; seen_points / scale are the
 * forward outputs of zs_unproject_normalize_f32. */
int zs_unproject_normalize_bwd_f32(const float* depth, const float* mask, const float* K, const float* seen_points,
                                   const float* scale, const float* dseen, float* ddepth, float* dKinv, int B, int H, int W,
                                   void* stream);

/* debug: effective SM clock in MHz at this point of the stream (one-thread spin kernel, ~10 us). */
int zs_debug_clock_mhz(float* out, void* stream);

/* out[A,N] = mean over the middle axis of x[A,M,N] (Z-mean of the attention maps, utils/eval_3D.py:47-52). */
int zs_mean_axis1_f32(const float* x, float* out, int64_t A, int M, int N, void* stream);

/* LinearProj3D (model/shape/implicit.py:128-131): out[M,C] = points[M,3] W[C,3]^T + bias.  Output-bandwidth bound. */
int zs_point_proj_f32(const float* points, int64_t M, const float* W, const float* bias, float* out, int C, void* stream);

/* concat-and-divide used by MLPBlocks skip layers (implicit.py:179-180): y[r,:] = [a[r,:Ca], b[r,:Cb]] / s */
int zs_concat2_f32(const float* a, int lda, int Ca, const float* b, int ldb, int Cb, float s, float* y, int ldy,
                   int64_t rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Geometry glue (utils/camera.py:52-108, model/compute_graph/graph_shape.py:89-144, utils/util.py:336-345)
 * ---------------------------------------------------------------------------------------------- */

/* intr_param2mtx (graph_shape.py:89-113): params [B,3] -> K [B,3,3]. */
int zs_intr_param2mtx_f32(const float* params, float* K, int B, int H, int W, void* stream);

/* unproj_depth + valid_norm_fac + normalise + zero background (graph_shape.py:136-141), no host sync.
 * depth [B,H*W], mask [B,H*W] (valid iff > 0.5), K [B,3,3] -> seen_points [B,H*W,3], mean [B,3], scale [B].
 * mask == NULL: raw unprojection only (utils/camera.py:88-108), mean/scale untouched.
 * workspace: zs_unproject_ws_bytes(B) bytes (currently 0). */
size_t zs_unproject_ws_bytes(int B);
int zs_unproject_normalize_f32(const float* depth, const float* mask, const float* K,
                               float* seen_points, float* mean, float* scale,
                               int B, int H, int W, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Chamfer (external/chamfer3D/chamfer3D.cu:12-196, external/chamfer3D/chamfer_cuda.cpp:17-32)
 * ---------------------------------------------------------------------------------------------- */

/* Bidirectional nearest neighbour: squared distances + int32 argmin (lowest index on ties), exactly the
 * contract of chamfer_3D.forward(xyz1,xyz2,dist1,dist2,idx1,idx2) but on the caller's stream and with all
 * SMs busy at b=1.  ws: zs_chamfer_ws_bytes(b,n,m) bytes of scratch (packed 64-bit min keys). */
size_t zs_chamfer_ws_bytes(int b, int n, int m);
int zs_chamfer_nn_fwd(const float* xyz1, const float* xyz2, int b, int n, int m,
                      float* dist1, float* dist2, int32_t* idx1, int32_t* idx2, void* ws, void* stream);
/* chamfer_3D.backward: gradients are ACCUMULATED into gradxyz1/gradxyz2 (caller zero-fills), like the reference. */
int zs_chamfer_nn_bwd(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                      const int32_t* idx1, const int32_t* idx2, int b, int n, int m,
                      float* gradxyz1, float* gradxyz2, void* stream);
/* Exact nearest neighbours through a flat bounding-box hierarchy: the same (squared distance, lowest-index-on-ties) results
 * as zs_chamfer_nn_fwd -- distances bit-identical, evaluated on the caller's coordinates -- at ~1/20 of the pair evaluations.
 * Built for the brute-force pose-search evaluator (utils/eval_3D.py:140-207: 6912 rotations x a 10k x 10k Chamfer per shape).
 *   zs_nn_bvh_build: `sets` point sets [sets, n, 3] (n <= 16384) -> `bvh` (zs_nn_bvh_bytes, 16-byte aligned): per set the points
 *                    sorted along a Morton curve (x, y, z, original index) + one tight box per cluster of 32.
 *   zs_nn_bvh_query: for b < batch: queries q[(sets_q == 1 ? 0 : b)] [nq, 3] against target set (sets_t == 1 ? 0 : b);
 *                    writes dist [batch, nq] (squared) and idx [batch, nq] (index into the ORIGINAL target order).
 *                    `q_order` (optional, [nq] int32): processing order of the queries (e.g. a Morton order, so that
 *                    neighbouring threads walk the same boxes); results are stored at the query's own index.
 *                    `variant` 0 = one thread per query, 1 = warp-cooperative (a warp answers its 32 queries one at a time with all
 *                    lanes: no divergence); identical results. */
size_t zs_nn_bvh_bytes(int sets, int n);
int zs_nn_bvh_build(const float* pts, int sets, int n, void* bvh, void* stream);
int zs_nn_bvh_query(const void* bvh, int sets_t, int n, const float* q, int sets_q, int nq, int batch,
                    const int32_t* q_order, float* dist, int32_t* idx, int variant, void* stream);

/* compute_fscore + means (utils/eval_3D.py:131-137,215-231) from squared distances: sqrt, mean, and
 * strict-< threshold fractions.  `squared`=1: inputs are squared distances (sqrt taken first, like
 * eval_3D.py:268-269), 0: already sqrt-ed.  mean1/2 [b], frac1/2 [b,T]. */
int zs_chamfer_stats(const float* sqdist1, const float* sqdist2, int b, int n, int m,
                     const float* thresholds, int T, int squared, float* mean1, float* mean2, float* frac1, float* frac2,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * Marching cubes + surface sampling (utils/eval_3D.py:233-263; PyMCubes 0.1.4 / trimesh 4.0.8 semantics)
 * ---------------------------------------------------------------------------------------------- */

/* Pass 1: count.  vol [n,n,n] fp32 (index order i,j,k = x,y,z); a corner is "inside" iff vol <= iso
 * (PyMCubes).  Writes counts[0] = #vertices (one per crossed grid edge, welded), counts[1] = #triangles.
 * ws: zs_mc_ws_bytes(n). */
size_t zs_mc_ws_bytes(int n);
int zs_mc_count(const float* vol, int n, float iso, void* ws, int32_t* counts, void* stream);
/* Pass 2: emit (after the caller read counts and allocated).  verts [V,3] fp32 in array-index units,
 * faces [F,3] int32.  Deterministic ordering (by edge id / cell id). */
int zs_mc_emit(const float* vol, int n, float iso, void* ws, float* verts, int32_t* faces, void* stream);
/* The same two passes on an x-SLAB vol [nx, n, n] of the (n)^3 grid (multi-GPU partitioning of SURVEY.md 8e A: the slab of
 * slices [x0, x1) plus the one-slice halo x1): cells whose low corner lies in local slices 0 .. nx-2, so every cell of the grid
 * is owned by exactly one slab; vertex x coordinates are emitted in GLOBAL index units (local slice + x_offset).  The y/z-edge
 * vertices of the halo slice are emitted by both neighbours (bit-identical coordinates): seam vertices are duplicated, faces
 * are not.  nx == n, x_offset == 0 reproduces zs_mc_count / zs_mc_emit. */
size_t zs_mc_slab_ws_bytes(int nx, int n);
int zs_mc_slab_count(const float* vol, int nx, int n, float iso, void* ws, int32_t* counts, void* stream);
int zs_mc_slab_emit(const float* vol, int nx, int n, float iso, void* ws, float* verts, int32_t* faces, int x_offset,
                    void* stream);
/* Area-weighted surface sampling with an explicit counter-based RNG seed (replaces trimesh.sample).
 * ws: zs_mesh_sample_ws_bytes(F). points [S,3]; vertices are scaled v*vscale+voffset first
 * (eval_3D.py:252-255). */
size_t zs_mesh_sample_ws_bytes(int F);
int zs_mesh_sample(const float* verts, const int32_t* faces, int V, int F, float vscale, float voffset,
                   int S, uint64_t seed, void* ws, float* points, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ZEROSHAPE_B200_H */
