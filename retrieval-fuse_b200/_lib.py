"""ctypes binding of librf_b200.so (the C ABI declared in include/rf_b200.h).

There is no CPU fallback: if the shared library is missing, importing any op
raises.  The library is built in-tree by `__graft_entry__.build()` (nvcc,
-gencode arch=compute_100a,code=sm_100a) and travels to the GPU box.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_long, c_size_t, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librf_b200.so")

_c_float_p = c_void_p  # device pointers travel as integers
_int3 = c_int * 3
_ptr4 = c_void_p * 8

# name -> (restype, argtypes); must list every symbol of include/rf_b200.h
PROTOTYPES = {
    "rf_last_error": (c_char_p, []),
    "rf_version": (c_int, []),
    "rf_device_info": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "rf_create": (c_int, [c_int, POINTER(c_void_p)]),
    "rf_destroy": (c_int, [c_void_p]),
    "rf_handle_device": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int)]),
    "rf_unfold3d": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "rf_fold3d": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "rf_unfold3d_pad_stride": (c_int, [c_void_p, c_void_p, c_int, c_int, _int3, _int3, _int3, _int3, c_float, c_float,
                                       c_float, c_void_p]),
    "rf_recompose_patches": (c_int, [c_void_p, c_void_p, c_int, c_int, _int3, _int3, _int3, _int3, _int3, c_float,
                                     c_void_p]),
    "rf_conv3d_fwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_float, c_void_p]),
    "rf_linear_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "rf_groupnorm_stats": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                   c_int, c_int, c_float, c_void_p]),
    "rf_maxpool3d_2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rf_upsample_nearest_2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rf_l2_normalize_rows": (c_int, [c_void_p, c_void_p, c_long, c_int, c_float, c_void_p]),
    "rf_tc_weight_image_bytes": (c_size_t, [c_int, c_int]),
    "rf_tc_weight_image": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "rf_tc_linear_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_long, c_int, c_int, c_int, c_float,
                                 c_void_p]),
    "rf_tc_mlp_supported": (c_int, [c_int * 9, c_int]),
    "rf_tc_mlp_weight_image_bytes": (c_size_t, [c_int, c_int]),
    "rf_tc_mlp_weight_image": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "rf_tc_mlp_debug_read": (c_int, [c_void_p]),
    "rf_tc_mlp_fwd": (c_int, [c_void_p, c_int, _ptr4, _ptr4, c_int * 9, c_int, c_int, c_float, c_int, c_float, c_void_p, c_int,
                              c_long, c_void_p]),
    "rf_cl_gn_stats_workspace_bytes": (c_size_t, [c_int, c_int]),
    "rf_cl_gn_stats": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_float, c_void_p, c_void_p]),
    "rf_cl_norm_split": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_long, c_long,
                                 c_int, c_int, c_float, c_void_p]),
    "rf_cl_maxpool3d_2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rf_cl_transpose": (c_int, [c_void_p, c_void_p, c_long, c_long, c_int, c_int, c_void_p]),
    "rf_tc_conv_weight_image_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "rf_tc_conv_weight_image": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "rf_tc_conv3d_fwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                 c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p]),
    "rf_halo_act_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "rf_cl_norm_split_halo": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                      c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "rf_tc_conv_halo_weight_image_bytes": (c_size_t, [c_int, c_int, c_int]),
    "rf_tc_conv_halo_weight_image": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "rf_tc_conv3d_halo_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "rf_tc_conv3d_halo_geometry": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int)]),
    "rf_tc_conv3d_halo_debug_read": (c_int, [c_void_p]),
    "rf_tc_conv3d_halo_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_int, c_float, c_float, c_int, c_void_p]),
    "rf_halo_s2_act_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "rf_cl_split_parity_planes": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "rf_tc_conv3d_halo_s2_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "rf_tc_conv3d_halo_s2_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_int, c_float, c_float, c_int, c_void_p]),
    "rf_cl_norm_split_halo_wp": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                         c_int, c_int, c_int, c_float, c_void_p]),
    "rf_tc_conv_halo_wp_weight_image_bytes": (c_size_t, [c_int, c_int, c_int]),
    "rf_tc_conv_halo_wp_weight_image": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "rf_tc_conv3d_halo_wp_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "rf_tc_conv3d_halo_wp_pool_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "rf_tc_conv3d_halo_wp_pool_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                              c_int, c_int, c_float, c_float, c_void_p]),
    "rf_tc_conv3d_halo_gn_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "rf_tc_conv3d_halo_gn_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_int, c_float, c_float, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p, c_void_p, c_int,
                                         c_void_p]),
    "rf_tc_conv3d_halo_wp_geometry": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "rf_tc_conv3d_halo_wp_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "rf_unet_front16_fwd": (c_int, [c_void_p, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float,
                                    c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "rf_wrun_act_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "rf_cl_norm_split_wrun": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_float, c_void_p]),
    "rf_tc_conv_wrun_weight_image_bytes": (c_size_t, [c_int, c_int]),
    "rf_tc_conv_wrun_weight_image": (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "rf_tc_conv3d_wrun_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "rf_tc_conv3d_wrun_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_float, c_float, c_int, c_void_p]),
    "rf_cl_pointwise_head": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_int, c_float, c_void_p]),
    "rf_conv3d_cin1_cl_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                      c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "rf_mlp_encode_workspace_bytes": (c_size_t, [c_long, c_int * 9, c_int]),
    "rf_mlp_encode_fwd": (c_int, [c_void_p, _ptr4, _ptr4, c_int * 9, c_int, c_int, c_void_p, c_long, c_void_p, c_size_t,
                                  c_void_p]),
    "rf_knn_workspace_bytes": (c_size_t, [c_long, c_long, c_int, c_int]),
    "rf_knn_l2_topk": (c_int, [c_void_p, c_long, c_long, c_void_p, c_long, c_int, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_size_t, c_void_p]),
    "rf_knn_tc_stats": (c_int, [c_void_p, POINTER(c_int), POINTER(c_float), c_void_p]),
    "rf_knn_bank_method": (c_int, [c_long, c_int]),
    "rf_knn_bank_image_bytes": (c_size_t, [c_long, c_int]),
    "rf_knn_bank_scratch_bytes": (c_size_t, [c_long, c_int]),
    "rf_knn_bank_prepare": (c_int, [c_void_p, c_long, c_int, c_void_p, c_long, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "rf_knn_prepared_workspace_bytes": (c_size_t, [c_long, c_long, c_int, c_int]),
    "rf_knn_l2_topk_prepared": (c_int, [c_void_p, c_long, c_long, c_void_p, c_int, c_void_p, c_long, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p, c_size_t, c_void_p]),
    "rf_knn_merge": (c_int, [c_void_p, c_void_p, c_int, c_long, c_int, c_void_p, c_void_p, c_void_p]),
    "rf_knn_demote_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p]),
    "rf_compose_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, _int3, _int3,
                                  c_float, c_float, c_float, c_float, c_void_p]),
    "rf_compose_gather_patches": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, _int3, _int3,
                                          c_float, c_float, c_float, c_float, c_void_p]),
    "rf_attention_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "rf_attention_fuse_fwd": (c_int, [c_void_p, c_void_p, _ptr4, _ptr4, _ptr4, _ptr4, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "rf_attention_fuse_patched_fwd": (c_int, [c_void_p, c_void_p, _ptr4, _ptr4, _ptr4, _ptr4, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                              c_size_t, c_void_p]),
    "rf_attention_features": (c_int, [c_void_p, c_void_p, c_void_p, _ptr4, _ptr4, _ptr4, _ptr4, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "rf_sobel_normals": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "rf_occupancy_counts": (c_int, [c_void_p, c_void_p, c_int, c_long, c_void_p, c_void_p]),
    "rf_chamfer_nn": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "rf_conv3d_wgrad": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rf_act_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_long, c_int, c_float, c_void_p]),
    "rf_channel_sum": (c_int, [c_void_p, c_int, c_int, c_long, c_void_p, c_void_p]),
    "rf_gn_bwd_workspace_bytes": (c_size_t, [c_int, c_int]),
    "rf_gn_bwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                          c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rf_upsample2_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "rf_maxpool3d_2_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rf_attention_epilogue_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_int, c_int,
                                          c_int, c_int, c_int, c_float, c_void_p]),
    "rf_attention_epilogue_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_long, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "rf_ntxent_workspace_bytes": (c_size_t, [c_int]),
    "rf_ntxent_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_float, c_float, c_float, c_int, c_void_p, c_void_p,
                              c_size_t, c_void_p]),
}

_lib = None


class RfError(RuntimeError):
    pass


def lib():
    """Loads the shared library once; raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RfError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU or PyTorch fallback for the rf_b200 kernels.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


_handles = {}


def handle(device_index):
    """rf_create(device) once per device of this process (eager per-device kernel setup); kept until exit."""
    h = _handles.get(device_index)
    if h is None:
        out = c_void_p()
        check(lib().rf_create(int(device_index), ctypes.byref(out)), "rf_create")
        _handles[device_index] = h = out
    return h


def check(rc, what=""):
    if rc != 0:
        msg = lib().rf_last_error().decode(errors="replace")
        raise RfError(f"{what or 'rf_b200 call'} failed (rc={rc}): {msg}")


def int3(v):
    if isinstance(v, int):
        v = (v, v, v)
    return _int3(int(v[0]), int(v[1]), int(v[2]))


def ptr_array(ptrs):
    arr = _ptr4()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
