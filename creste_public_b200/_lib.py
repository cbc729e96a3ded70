"""ctypes binding of libcreste_b200.so (the C ABI declared in include/creste_b200.h).

The library is the product: there is no PyTorch / CPU fallback.  Importing this module never
needs a GPU (the driver tier imports the package on a CPU box), but every op raises loudly if
the shared library is missing or its tensors are not CUDA tensors.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcreste_b200.so")

EXPORTS = [
    "creste_version", "creste_last_error", "creste_num_sms", "creste_launch_count",
    "creste_vi_workspace_bytes", "creste_vi_solve",
    "creste_svf_workspace_bytes", "creste_svf",
    "creste_frustum_to_bev", "creste_camera_to_world", "creste_points_to_voxels", "creste_zmlp_concat", "creste_zmlp_concat_ex",
    "creste_splat_workspace_bytes", "creste_splat_soft", "creste_splat_bwd_workspace_bytes", "creste_splat_soft_bwd",
    "creste_frustum_bwd", "creste_depth_expectation_bwd", "creste_dilate", "creste_phase_slice",
    "creste_lidar_raster", "creste_depth_expectation", "creste_bin_depths",
    "creste_conv2d", "creste_conv2d_ex", "creste_conv2d_presplit", "creste_conv2d_split_out", "creste_conv2d_presplit_split_out", "creste_upsample_concat_split", "creste_conv2d_workspace_bytes", "creste_conv2d_tc_supported",
    "creste_conv2d_tc_layout", "creste_conv2d_tc_debug",
    "creste_f16_split", "creste_f16_split_amax", "creste_chan_axpby_act", "creste_chan_affine_amax", "creste_relu_bwd_amax", "creste_chan_affine_act_amax", "creste_chan_axpby_amax", "creste_conv2d_wgrad_tc_presplit", "creste_affine_warp", "creste_depth_augment", "creste_traverse_to_bev", "creste_dwconv_num_parts", "creste_dwconv_tile_parts", "creste_dwconv_parts", "creste_dwconv_bn_swish", "creste_dwconv_bn_swish_ex", "creste_se_gate",
    "creste_upsample_concat", "creste_maxpool2_concat",
    "creste_nchw_to_nhwc", "creste_nhwc_to_nchw", "creste_proj_head",
    "creste_expert_visitation",
    "creste_chan_affine", "creste_relu_bwd", "creste_chan_dot_workspace_bytes", "creste_chan_dot", "creste_chan_stats",
    "creste_maxpool2_bwd", "creste_maxpool2_gather", "creste_upsample_adjoint", "creste_upsample_adjoint_slice",
    "creste_conv2d_wgrad_workspace_bytes", "creste_conv2d_wgrad",
    "creste_row_dot", "creste_row_scale", "creste_row_normalize",
    "creste_grad_penalty_workspace_bytes", "creste_grad_penalty", "creste_grad_penalty_bwd",
    "creste_adam_step", "creste_stage1_depth_losses", "creste_masked_mse",
    "creste_chan_reduce_workspace_bytes", "creste_chan_moments", "creste_chan_affine_act",
    "creste_bn_act_bwd", "creste_chan_axpby", "creste_dwconv_fwd", "creste_dwconv_dgrad",
    "creste_dwconv_wgrad_workspace_bytes", "creste_dwconv_wgrad", "creste_sample_dot",
    "creste_sample_affine", "creste_act", "creste_act_bwd", "creste_add_scaled", "creste_chan_slice",
    "creste_wgrad_strided_workspace_bytes", "creste_wgrad_strided", "creste_ce_depth_bwd",
    "creste_masked_mse_bwd",
    "creste_conv2d_wgrad_tc_supported", "creste_conv2d_wgrad_tc_workspace_bytes", "creste_conv2d_wgrad_tc",
    "creste_wgrad_rows", "creste_bn_fwd_finalize", "creste_bn_bwd_finalize",
    "creste_pack_weight_f16",
    "creste_smooth_l1", "creste_smooth_l1_bwd", "creste_ce_weighted", "creste_ce_weighted_bwd",
    "creste_l2norm_rows", "creste_l2norm_rows_bwd", "creste_supcon_fwd", "creste_supcon_bwd",
]


class ConvDesc(C.Structure):
    """creste_conv_desc (include/creste_b200.h)."""
    _fields_ = [(n, C.c_int) for n in (
        "N", "H", "W", "C", "K", "R", "S", "stride", "pad_t", "pad_l", "P", "Q", "act",
        "out_nchw", "precision")]


_lib = None


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into csrc/libcreste_b200.so (nvcc cross-compiles on CPU)."""
    import subprocess
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    res = subprocess.run(cmd, capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libcreste_b200.so failed:\n" + (res.stdout or "") + (res.stderr or ""))
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(creste_public_b200 has no CPU or PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        L.creste_last_error.restype = C.c_char_p
        L.creste_launch_count.restype = C.c_ulonglong
        for name in ("creste_vi_workspace_bytes", "creste_svf_workspace_bytes",
                     "creste_splat_workspace_bytes", "creste_splat_bwd_workspace_bytes", "creste_conv2d_workspace_bytes",
                     "creste_chan_dot_workspace_bytes", "creste_conv2d_wgrad_workspace_bytes",
                     "creste_grad_penalty_workspace_bytes", "creste_chan_reduce_workspace_bytes",
                     "creste_dwconv_wgrad_workspace_bytes", "creste_wgrad_strided_workspace_bytes",
                     "creste_conv2d_wgrad_tc_workspace_bytes"):
            getattr(L, name).restype = C.c_size_t
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().creste_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def ptr(t, dtype=None):
    """Raw device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("creste_public_b200 ops need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("creste_public_b200 ops need contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"expected dtype {dtype}, got {t.dtype}")
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
