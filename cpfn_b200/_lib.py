"""ctypes binding of libcpfn_b200.so (the C ABI declared in include/cpfn_b200.h).

There is no fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  Build it with ``python -m cpfn_b200.build`` (or
``__graft_entry__.build()``).
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcpfn_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "cpfn_b200.h")

_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol of include/cpfn_b200.h
# (tests/test_abi.py checks this table against the header).
SIGNATURES = {
    "cpfn_version": (c_int, []),
    "cpfn_error_string": (ctypes.c_char_p, [c_int]),
    "cpfn_last_cuda_error": (ctypes.c_char_p, []),
    "cpfn_sm_count": (c_int, []),
    "cpfn_debug_fps_profile": (c_int, [c_void_p]),
    "cpfn_debug_chain_profile": (c_int, [c_void_p, c_int, c_int]),
    "cpfn_sym_eigh_small": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_void_p]),
    "cpfn_small_solve": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p, c_void_p]),
    "cpfn_fps_rounds_supported": (c_int, [c_int, c_int]),
    "cpfn_furthest_point_sampling_rounds": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                                    c_size_t, c_size_t, c_void_p]),
    "cpfn_fps_workspace_bytes": (c_size_t, [c_int, c_int]),
    "cpfn_furthest_point_sampling": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                             c_size_t, c_void_p]),
    "cpfn_ball_query": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int,
                                c_void_p, c_void_p]),
    "cpfn_ball_query_grid_workspace_bytes": (c_size_t, [c_int, c_int]),
    "cpfn_ball_query_grid": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    "cpfn_gather_points": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                   c_void_p]),
    "cpfn_gather_points_grad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                        c_void_p]),
    "cpfn_group_points": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p]),
    "cpfn_group_points_grad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                       c_void_p, c_void_p]),
    "cpfn_three_nn": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                              c_void_p]),
    "cpfn_three_weighted_sum": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p]),
    "cpfn_three_weighted_sum_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                             c_int, c_void_p, c_void_p]),
    "cpfn_moments_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "cpfn_weighted_moments": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                      c_void_p]),
    "cpfn_weighted_moments_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                           c_void_p, c_void_p]),
    "cpfn_three_nn_weights": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cpfn_three_nn_weights_sorted": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cpfn_linear_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cpfn_gather_xyz": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cpfn_normalise_patches": (c_int, [c_void_p, ctypes.c_longlong, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cpfn_seg_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "cpfn_label_membership_sums": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cpfn_hungarian_matching": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                        c_void_p]),
    "cpfn_zero_fill": (c_int, [c_void_p, c_size_t, c_void_p]),
    "cpfn_rng_set": (c_int, [c_void_p, ctypes.c_ulonglong, ctypes.c_ulonglong, c_void_p]),
    "cpfn_dropout_mask_bits": (c_int, [c_void_p, c_int, c_int, c_int, c_float, ctypes.c_longlong, c_void_p, c_void_p]),
    "cpfn_spfn_post": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p]),
    "cpfn_spfn_post_scatter": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "cpfn_extract_patches_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "cpfn_extract_patches": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_size_t, c_void_p]),
    "cpfn_merge_inverse_bytes": (c_size_t, [c_int, c_int]),
    "cpfn_merge_inverse_index": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cpfn_merge_similarity_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "cpfn_merge_similarity": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p, c_void_p, c_size_t, c_void_p]),
    "cpfn_heuristic_merging_host": (c_int, [c_void_p, c_void_p, ctypes.c_int64, c_void_p, ctypes.c_int64, c_void_p]),
    "cpfn_merge_solve_host": (c_int, [c_void_p, ctypes.c_int64, ctypes.c_double, c_void_p, c_void_p]),
    "cpfn_merge_solve_host_f32": (c_int, [c_void_p, ctypes.c_int64, c_float, c_void_p, c_void_p]),
    "cpfn_merge_solve_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "cpfn_merge_solve": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    "cpfn_merge_point_labels": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                        c_int, c_int, c_void_p, c_void_p]),
    "cpfn_merge_dense_labels": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cpfn_merge_normals_types": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_int, c_void_p, c_void_p, c_void_p]),
    "cpfn_primitive_residues": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int,
                                        c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "cpfn_p_coverage": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                c_void_p, c_void_p]),
    "cpfn_fps_dense_workspace_bytes": (c_size_t, []),
    "cpfn_fps_dense": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                               c_void_p]),
    "cpfn_ball_query_grid_build": (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_size_t, c_void_p]),
    "cpfn_ball_query_grid_query": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                           c_size_t, c_void_p]),
    "cpfn_ball_query_grid_query_range": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int,
                                                 c_void_p, c_void_p, c_size_t, c_void_p]),
    "cpfn_furthest_point_sampling_xyz": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                                 c_void_p]),
    "cpfn_mlp_packed_bytes": (c_size_t, [c_int, c_int]),
    "cpfn_mlp_pack_weights_host": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "cpfn_mlp_chain": (c_int, [c_void_p, c_void_p]),
    "cpfn_fit_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "cpfn_fit_primitives": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_size_t, c_void_p]),
}


def header_symbols():
    """Every function name declared in include/cpfn_b200.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return sorted(set(re.findall(r"CPFN_API[^;(]*?\b(cpfn_\w+)\s*\(", text)))


def lib():
    """Load the shared library once; fail loudly when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "cpfn_b200: %s is missing -- build it with `python -m cpfn_b200.build`. "
                "There is no CPU or PyTorch fallback for this path." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code, what):
    if code != 0:
        L = lib()
        msg = L.cpfn_error_string(code).decode()
        if code == -2:
            msg += ": " + L.cpfn_last_cuda_error().decode()
        raise RuntimeError("cpfn_b200.%s failed: %s" % (what, msg))
