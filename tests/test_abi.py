"""The C-ABI library loads and exports every symbol include/cpfn_b200.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import subprocess

import pytest
import torch

from cpfn_b200 import _lib, cuda_ops


def test_header_and_signature_table_agree(built_lib):
    assert set(_lib.header_symbols()) == set(_lib.SIGNATURES)
    assert len(_lib.header_symbols()) >= 14


def test_library_exports_every_header_symbol(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True,
                         check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = set(_lib.header_symbols()) - exported
    assert not missing, missing


def test_library_loads_and_reports_version(built_lib):
    L = _lib.lib()
    assert L.cpfn_version() >= 100
    assert L.cpfn_error_string(0) == b"ok"
    assert b"workspace" in L.cpfn_error_string(-3)
    assert L.cpfn_fps_workspace_bytes(4, 8192) == 0
    assert L.cpfn_fps_workspace_bytes(2, 100000) == 2 * 100000 * 4


def test_no_torch_types_in_abi(built_lib):
    out = subprocess.run(["nm", "-D", "-C", built_lib], capture_output=True, text=True,
                         check=True).stdout
    assert "at::" not in out and "c10::" not in out and "torch::" not in out


def test_library_is_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    archs = {ln.rsplit(".", 2)[-2] for ln in out.splitlines() if ln.strip().endswith(".cubin")}
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("call", [
    lambda: cuda_ops.farthest_point_sampling(torch.zeros(1, 8, 3), 4),
    lambda: cuda_ops.ball_query(torch.zeros(1, 2, 3), torch.zeros(1, 8, 3), 0.2, 4),
    lambda: cuda_ops.three_nn(torch.zeros(1, 8, 3), torch.zeros(1, 4, 3)),
    lambda: cuda_ops.gather_points(torch.zeros(1, 2, 8), torch.zeros(1, 4, dtype=torch.int32)),
    lambda: cuda_ops.group_points(torch.zeros(1, 2, 8), torch.zeros(1, 4, 2, dtype=torch.int32)),
    lambda: cuda_ops.three_weighted_sum(torch.zeros(1, 2, 8), torch.zeros(1, 4, 3, dtype=torch.int32),
                                        torch.zeros(1, 4, 3)),
])
def test_cpu_tensors_are_rejected_like_the_reference(call):
    # src/*.cpp: TORCH_CHECK(false, "CPU not supported")
    with pytest.raises(RuntimeError, match="CPU not supported"):
        call()


def test_dtype_and_layout_checks_match_reference_messages():
    # include/utils.h:5-25
    with pytest.raises(RuntimeError, match="must be a float tensor"):
        cuda_ops.farthest_point_sampling(torch.zeros(1, 8, 3, dtype=torch.float64), 4)
    with pytest.raises(RuntimeError, match="must be an int tensor"):
        cuda_ops.gather_points(torch.zeros(1, 2, 8), torch.zeros(1, 4, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="must be a contiguous tensor"):
        cuda_ops.farthest_point_sampling(torch.zeros(1, 3, 8).permute(0, 2, 1), 4)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_argument_validation_needs_no_gpu(built_lib):
    """Every entry point checks its arguments before it touches CUDA: invalid calls return CPFN_EINVAL (-1) /
    CPFN_EWORKSPACE (-3) on a machine without a GPU, and the workspace-size helpers are pure host arithmetic."""
    L = _lib.lib()
    null = None
    one = ctypes.c_void_p(16)                       # a non-null, never dereferenced pointer
    assert L.cpfn_furthest_point_sampling(null, 2, 100, 10, one, null, 0, null) == -1
    assert L.cpfn_ball_query(null, null, 2, 100, 10, 0.2, 4, null, null) == -1
    assert L.cpfn_extract_patches(one, 100, one, 1, 200, one, null, null, one, 1 << 30, null) == -1       # k > N
    assert L.cpfn_extract_patches(one, 100000, one, 1, 20000, one, null, null, one, 1 << 30, null) == -1  # k > 16384
    assert L.cpfn_extract_patches(one, 100, one, 1, 10, one, null, null, null, 0, null) == -3             # no workspace
    assert L.cpfn_extract_patches_workspace_bytes(131072, 32, 8192) >= 32 * (131072 * 4 + 8192 * 8)
    assert L.cpfn_merge_inverse_bytes(32, 131072) == 32 * 131072 * 4
    assert L.cpfn_merge_similarity_workspace_bytes(32, 21, 28) == 700 * 700 * 8
    assert L.cpfn_merge_similarity(one, one, one, one, 2, 64, 40, 1000, 28, one, one, 1 << 30, null) == -1  # Kl > 32
    assert L.cpfn_merge_point_labels(one, one, one, one, one, 2, 64, 21, 1000, 28, 5000, one, null) == -1   # L too large
    assert L.cpfn_merge_normals_types(one, one, one, one, one, 2, 64, 1000, 9, one, one, null) == -1        # n_types > 8
    assert L.cpfn_primitive_residues(one, one, one, 0, 0, 1, 4, 4, 16, null, 4, one, one, null) == -1       # no class list
    assert L.cpfn_p_coverage(one, one, one, one, 1, 4, 4, 100, one, 5, one, null) == -1                      # > 4 thresholds
    assert L.cpfn_fps_dense(one, 100, null, null, 0, 100, 10, one, one, 1 << 20, null) == -1                 # start outside the cloud
    assert L.cpfn_fps_dense(one, 100, null, null, 0, 0, 10, one, null, 0, null) == -3
    assert L.cpfn_fps_dense_workspace_bytes() > 0
    assert L.cpfn_heuristic_merging_host(null, null, 0, null, 4, null) == -1
    assert L.cpfn_merge_solve_host_f32(null, 4, 0.0, null, null) == -1
    # round-2 entry points
    assert L.cpfn_three_nn_weights_sorted(one, one, 2, 100, 128, one, one, null) == -1         # m below the grid range
    assert L.cpfn_three_nn_weights_sorted(null, one, 2, 100, 512, one, one, null) == -1
    assert L.cpfn_sym_eigh_small(one, 4, 4, one, one, null) == -1                               # D = 2 or 3 only
    assert L.cpfn_sym_eigh_small(null, 4, 3, one, one, null) == -1
    assert L.cpfn_small_solve(one, one, 4, 4, 0, one, null) == -1                               # D <= 3
    assert L.cpfn_small_solve(one, null, 4, 3, 0, one, null) == -1
    assert L.cpfn_debug_chain_profile(null, 4, 0) == -1                                         # profiling not enabled
    assert L.cpfn_seg_workspace_bytes(2, 1000, 28, 28) >= 2 * 29 * 29 * 8
    assert L.cpfn_seg_workspace_bytes(2, 1000, 80, 28) == 0                                      # > 64 slots
