"""ctypes binding of libzeroshape_b200.so (the C ABI declared in include/zeroshape_b200.h).

The product has NO fallback: if the library is missing or a symbol is absent, importing this
module raises.  Nothing here imports `oracle/`.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZEROSHAPE_B200_LIB", os.path.join(_HERE, "libzeroshape_b200.so"))

P = c_void_p  # device pointer


# name -> (restype, argtypes); must list every symbol of include/zeroshape_b200.h
SIGNATURES = {
    "zs_last_error": (c_char_p, []),
    "zs_abi_version": (c_int, []),
    "zs_device_cc": (c_int, []),
    "zs_launch_count": (ctypes.c_longlong, []),
    "zs_launch_count_add": (None, [ctypes.c_longlong]),
    "zs_gemm_f32": (c_int, [P, c_int, P, c_int, P, P, c_int, c_int, P, c_int, c_int, c_int, c_int, c_int, P]),
    "zs_gemm_tc_packed_bytes": (c_size_t, [c_int, c_int]),
    "zs_gemm_tc_pack": (c_int, [P, c_int, c_int, c_int, P, P]),
    "zs_gemm_tc_pack_fmt": (c_int, [P, c_int, c_int, c_int, P, c_int, P]),
    "zs_chain_qkvattn_blob_bytes": (c_size_t, []),
    "zs_chain_qkvattn_fwd": (c_int, [P, c_int, c_int, c_float, P, P, P, P, c_int, c_float, P, c_int, c_int, c_int, P]),
    "zs_gemm_tc_f32": (c_int, [P, c_int, P, P, P, c_int, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "zs_attn_scores_tc": (c_int, [P, c_int, P, c_int, c_int, c_float, P, P, P, c_int, P]),
    "zs_attn_pv_tc": (c_int, [P, P, P, P, P, c_int, c_int, P]),
    "zs_debug_chain_trace": (c_int, [P]),
    "zs_debug_chain_variant": (c_int, [c_int]),
    "zs_debug_gemm_splitk": (c_int, [c_int]),
    "zs_chain_attn_fwd": (c_int, [P, c_int, c_int, P, P, c_int, c_float, P, c_int, P]),
    "zs_debug_clock_mhz": (c_int, [P, P]),
    "zs_bce_logits_fwd": (c_int, [P, P, c_int64, c_float, c_float, P, P, P]),
    "zs_bce_logits_bwd": (c_int, [P, P, c_int64, c_float, c_float, c_float, P, P]),
    "zs_act_bwd_f32": (c_int, [P, P, P, c_int64, c_int, P]),
    "zs_colsum_f32": (c_int, [P, c_int, c_int64, c_int, P, c_int, P]),
    "zs_gemm_tn_f32": (c_int, [P, c_int, P, c_int, P, c_int, c_int64, c_int, c_int, c_int, P]),
    "zs_layernorm_bwd_f32": (c_int, [P, P, P, c_float, P, P, P, c_int64, c_int, P]),
    "zs_point_attention_bwd_f32": (c_int, [P, P, P, c_int, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P]),
    "zs_mha_bwd_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "zs_mha_bwd_f32": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_float, P, P]),
    "zs_conv2d_nhwc_dgrad_tc": (c_int, [P, c_int, c_int, c_int, c_int, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_int, P]),
    "zs_gemm_tn_tc": (c_int, [P, c_int, P, c_int, P, c_int, c_int64, c_int, c_int, c_int, c_int, c_int, P]),
    "zs_conv2d_nhwc_wgrad_tc": (c_int, [P, c_int, c_int, c_int, c_int, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_int, c_int, P]),
    "zs_conv2d_nhwc_dgrad_f32": (c_int, [P, c_int, c_int, c_int, c_int, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "zs_conv2d_nhwc_wgrad_f32": (c_int, [P, c_int, c_int, c_int, c_int, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_int, P]),
    "zs_bn_stats_f32": (c_int, [P, c_int64, c_int, c_float, P, P, P, P, P]),
    "zs_bn_bwd_f32": (c_int, [P, P, P, P, P, c_int64, c_int, P, P, P, P, P]),
    "zs_maxpool3x3s2_bwd_nhwc_f32": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "zs_avgpool_bwd_nhwc_f32": (c_int, [P, P, c_int, c_int, c_int, P]),
    "zs_coldot_f32": (c_int, [P, c_int, P, c_int, c_int64, c_int, P, c_int, P]),
    "zs_layernorm_bwd_generic_f32": (c_int, [P, P, P, c_float, P, P, c_int64, c_int, P]),
    "zs_groupnorm_bwd_nhwc_f32": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_float, P]),
    "zs_bilinear_bwd_nhwc_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "zs_unproject_normalize_bwd_f32": (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, P]),
    "zs_adamw_f32": (c_int, [P, P, P, P, c_int64, c_float, c_float, c_float, c_float, c_float, c_int, P]),
    "zs_coord_embed_windows_f32": (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "zs_depth_metrics_f32": (c_int, [P, P, P, c_int, c_int, c_int, P, c_int, c_float, c_int, P, P, P]),
    "zs_mask_erode_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "zs_midas_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "zs_midas_loss_f32": (c_int, [P, P, P, c_int, c_int, c_int, c_float, c_int, c_float, P, P, P, P]),
    "zs_adamw_multi_f32": (c_int, [P, c_int, c_float, c_float, c_float, c_float, c_float, c_int, P]),
    "zs_adamw_multi_dev_f32": (c_int, [P, c_int, P, P]),
    "zs_mean_axis1_f32": (c_int, [P, P, c_int64, c_int, c_int, P]),
    "zs_point_proj_f32": (c_int, [P, c_int64, P, P, P, c_int, P]),
    "zs_chain_lin_fwd": (c_int, [P, c_int, c_int, c_int, c_float, P, c_int, P, P, c_int, P, c_int, c_int, P]),
    "zs_chain_mlp_blob_bytes": (c_size_t, []),
    "zs_chain_occ_blob_bytes": (c_size_t, []),
    "zs_chain_mlp_fwd": (c_int, [P, c_int, c_int, P, P, c_float, P, P, P, c_int, P]),
    "zs_chain_occ_fwd": (c_int, [P, c_int, P, c_int, P, P, c_float, P, P, P, c_float, P, c_int, c_int, P]),
    "zs_chain_pmlp_fwd": (c_int, [P, c_int, c_int, P, P, P, c_float, P, P, P, P, P, c_int, P]),
    "zs_chain_qkvattn_pts_fwd": (c_int, [P, c_int, P, P, c_float, P, P, P, P, c_int, c_float, P, c_int, c_int, P]),
    "zs_conv2d_nhwc_f32": (c_int, [P, c_int, c_int, c_int, c_int, P, P, P, c_int, P, c_int, c_int, c_int, c_int,
                                   c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "zs_conv2d_nhwc_tc": (c_int, [P, c_int, c_int, c_int, c_int, P, P, P, c_int, P, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "zs_layernorm_f32": (c_int, [P, c_int, P, P, P, c_int, c_int, c_int, c_float, P]),
    "zs_groupnorm_nhwc_f32": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_float, c_int, P]),
    "zs_groupnorm_ws_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "zs_groupnorm_nhwc_ws_f32": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_float, c_int, P, P]),
    "zs_channel_affine_f32": (c_int, [P, P, P, P, P, c_int64, c_int, c_int, P]),
    "zs_axpby_f32": (c_int, [P, c_float, P, c_float, P, c_int64, c_int, P]),
    "zs_maxpool3x3s2_nhwc_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "zs_avgpool_nhwc_f32": (c_int, [P, P, c_int, c_int, c_int, P]),
    "zs_bilinear_nhwc_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "zs_nchw_to_nhwc_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, c_float, c_float, P]),
    "zs_nhwc_to_nchw_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "zs_mha_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, c_float, P]),
    "zs_mha_tc_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, c_float, c_int, P]),
    "zs_point_attention_tc_f32": (c_int, [P, P, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_float, c_int, P]),
    "zs_mha_bwd_tc_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "zs_mha_bwd_tc_f32": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_float, P, P]),
    "zs_point_attention_bwd_tc_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "zs_point_attention_bwd_tc_f32": (c_int, [P, P, P, c_int, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, P]),
    "zs_rgba_crop_resize_u8": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, c_int, P, P, c_int, P, P, c_int, P]),
    "zs_rgba_composite_f32": (c_int, [P, c_int, c_int, c_int, c_float, P, P, P]),
    "zs_erode_square_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "zs_point_attention_f32": (c_int, [P, P, P, c_int, P, P, c_float, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_float, P]),
    "zs_dense_grid_f32": (c_int, [P, c_int, c_float, c_float, c_int, c_int, P]),
    "zs_concat2_f32": (c_int, [P, c_int, c_int, P, c_int, c_int, c_float, P, c_int, c_int64, P]),
    "zs_intr_param2mtx_f32": (c_int, [P, P, c_int, c_int, c_int, P]),
    "zs_unproject_ws_bytes": (c_size_t, [c_int]),
    "zs_unproject_normalize_f32": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, P, P]),
    "zs_chamfer_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "zs_chamfer_nn_fwd": (c_int, [P, P, c_int, c_int, c_int, P, P, P, P, P, P]),
    "zs_chamfer_nn_bwd": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, P, P, P]),
    "zs_nn_bvh_bytes": (c_size_t, [c_int, c_int]),
    "zs_nn_bvh_build": (c_int, [P, c_int, c_int, P, P]),
    "zs_nn_bvh_query": (c_int, [P, c_int, c_int, P, c_int, c_int, c_int, P, P, P, c_int, P]),
    "zs_chamfer_stats": (c_int, [P, P, c_int, c_int, c_int, P, c_int, c_int, P, P, P, P, P]),
    "zs_mc_ws_bytes": (c_size_t, [c_int]),
    "zs_mc_count": (c_int, [P, c_int, c_float, P, P, P]),
    "zs_mc_emit": (c_int, [P, c_int, c_float, P, P, P, P]),
    "zs_mc_slab_ws_bytes": (c_size_t, [c_int, c_int]),
    "zs_mc_slab_count": (c_int, [P, c_int, c_int, c_float, P, P, P]),
    "zs_mc_slab_emit": (c_int, [P, c_int, c_int, c_float, P, P, P, c_int, P]),
    "zs_mesh_sample_ws_bytes": (c_size_t, [c_int]),
    "zs_mesh_sample": (c_int, [P, P, c_int, c_int, c_float, c_float, c_int, c_uint64, P, P, P]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"zeroshape_b200: CUDA library not found at {LIB_PATH}. Build it with "
            f"`python -m zeroshape_b200.build` (nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError(f"zeroshape_b200: {LIB_PATH} does not export `{name}` (stale build?)") from e
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class NativeError(RuntimeError):
    pass


def check(status, what=""):
    if status != 0:
        msg = lib.zs_last_error()
        raise NativeError(f"{what or 'zeroshape_b200'} failed ({status}): {msg.decode() if msg else ''}")
