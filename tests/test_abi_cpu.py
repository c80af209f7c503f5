"""CPU: the C-ABI library loads and exports every symbol include/tricolo_b200.h declares;
argument validation that needs no GPU; host-side logic of the drop-in modules."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tricolo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tcl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from tricolo_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 14
    raw = ctypes.CDLL(str(_lib.lib_path()))
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding table and header disagree"
    assert _lib.LIB.tcl_version() == 1
    assert _lib.LIB.tcl_last_error_string() is not None


def test_bwd_struct_layout_matches_header():
    from tricolo_b200 import _lib

    # tcl_bwd_segment: 5 pointers + 2 floats = 48 bytes; tcl_bwd_job: 4 pointers + 2 int32 + 2 segments
    assert ctypes.sizeof(_lib.BwdSegment) == 48
    assert ctypes.sizeof(_lib.BwdJob) == 4 * 8 + 8 + 2 * 48


def test_argument_validation_without_gpu():
    """Shape/alignment checks run before any CUDA call, so they are testable on the CPU box."""
    from tricolo_b200 import _lib

    L = _lib.LIB
    buf = ctypes.create_string_buffer(4096)
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = L.tcl_topk_rank(p, 8, 4, 8, 17, p, 0, None, p, p, p, p, None)  # k > 16
    assert rc == 4 and b"k must be" in L.tcl_last_error_string()
    rc = L.tcl_sim_gemm(p, p, 4, 4, 12, 1, p, 4, None)  # dim not a multiple of 8
    assert rc == 1
    arr = (ctypes.c_void_p * 1)(p.value)
    rc = L.tcl_ntxent_fwd(1, arr, arr, 128, 128, 100, 0, 0, 0, 10.0, p, p, p, p, 1 << 20, None)  # dim % 64
    assert rc == 1
    rc = L.tcl_ntxent_fwd(1, arr, arr, 128, 128, 128, 0, 0, 0, 1000.0, p, p, p, p, 1 << 20, None)  # tau too small
    assert rc == 4 and b"temperature" in L.tcl_last_error_string()
    # norm=False entry points: dim % 16, dtype, pair indices
    i3 = (ctypes.c_int32 * 1)(0)
    i4 = (ctypes.c_int32 * 1)(1)
    arr2 = (ctypes.c_void_p * 2)(p.value, p.value)
    rc = L.tcl_ntxent_raw_fwd(2, arr2, 0, 64, 100, 100, 1, i3, i4, 10.0, 0.25, p, 1 << 20, p, 1 << 20, p, None)
    assert rc == 1 and b"multiple of 16" in L.tcl_last_error_string()
    rc = L.tcl_ntxent_raw_fwd(2, arr2, 3, 64, 128, 128, 1, i3, i4, 10.0, 0.25, p, 1 << 20, p, 1 << 20, p, None)  # f64
    assert rc == 4
    rc = L.tcl_ntxent_raw_fwd(2, arr2, 0, 64, 128, 128, 1, i3, i3, 10.0, 0.25, p, 1 << 20, p, 1 << 20, p, None)  # pair (0, 0)
    assert rc == 4
    assert L.tcl_ntxent_raw_state_bytes(3, 8192) == 3 * 8192 * 2 * 4
    assert L.tcl_ntxent_raw_workspace_bytes(3, 8192) > 3 * 128 * 8192 * 8
    # gather with per-(destination, tensor) addresses: destination 0 is mandatory, NULL further destinations are skipped
    inv = (ctypes.c_void_p * 1)(p.value)
    dst = (ctypes.c_void_p * 2)(None, p.value)
    rc = L.tcl_l2norm_fwd_bcast(1, arr, 0, 8, 64, 64, 2, dst, 64, 0, inv, 1e-12, None)
    assert rc == 2 and b"destination 0" in L.tcl_last_error_string()
    rc = L.tcl_copy_rows(p, 64, p, 32, 64, 4, None)  # source pitch below the row width
    assert rc == 1
    assert L.tcl_ntxent_fwd_workspace_bytes(3, 8192, 8192) > 0
    assert L.tcl_ntxent_bwd_workspace_bytes(3, 8192, 512) >= 3 * 8192 * 512 * 4


def test_sharded_backward_plan_and_validation_without_gpu():
    """The plan of the sharded shared-G backward (buffer sizes, slot counts) is a pure host function of the sizes, and
    the sharded entry points validate their arguments before any CUDA call."""
    from tricolo_b200 import _lib, ops

    L = _lib.LIB
    pairs = [(0, 1), (0, 2), (1, 2)]
    p8 = ops.ShardedBwdPlan(3, pairs, (1, 1, 1), 1024, 8, 512)
    p2 = ops.ShardedBwdPlan(3, pairs, (1, 1, 1), 4096, 2, 512)
    # receive buffer: 1 KB header + 2 column-side tensors x world x slots x b_loc x dim x 2 bytes (fp16 partials)
    assert (p8.recv_bytes - 1024) % (2 * 8 * 1024 * 512 * 2) == 0 and p8.recv_bytes > 1024
    assert p2.workspace_bytes > 3 * 4096 * 8192 * 2  # holds the rank's row block of G for every pair
    # text never needs a gradient: only its column-side partner tensors get jobs; still a valid plan
    p = ops.ShardedBwdPlan(3, pairs, (0, 1, 1), 1024, 8, 512)
    assert p.recv_bytes == p8.recv_bytes
    assert not ops.ShardedBwdPlan.supported(1000, 512, 8) and not ops.ShardedBwdPlan.supported(1024, 256, 8)
    with pytest.raises(ValueError):
        ops.ShardedBwdPlan(3, pairs, (1, 1, 1), 1000, 8, 512)  # rows per rank not a multiple of 128
    assert L.tcl_shard_sync_bytes() == 4096
    assert L.tcl_shard_stats_bytes(3, 1024, 8) == 8 * 3 * (8192 + 2048) * 4
    assert L.tcl_shard_stats_bytes(4, 1024, 8) == 0 and L.tcl_rank_metrics_workspace_bytes() > 0
    buf = ctypes.create_string_buffer(8192)
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    arr = (ctypes.c_void_p * 8)(*([ptr.value] * 8))
    rc = L.tcl_l2norm_fwd_push(3, arr, 0, 1024, 512, 512, 9, 8, arr, 1536, 0, arr, 1e-12, arr, 1, None)  # rank 9 of 8
    assert rc == 4 and b"rank" in L.tcl_last_error_string()
    rc = L.tcl_ntxent_finalize_sharded(3, 1024, 8000, 0, 8, 10.0, 0.25, arr, None, ptr, ptr, ptr, None)  # b_glob != W * b_loc
    assert rc == 4
    rc = L.tcl_rank_metrics(ptr, 100, 17, ptr, ptr, 1 << 20, None)  # k > 16
    assert rc == 4


def test_no_cpu_fallback():
    import torch

    from tricolo_b200.loss import NTXentLoss

    fn = NTXentLoss(0.1, 0.25)
    with pytest.raises(RuntimeError, match="no CPU path"):
        fn(torch.randn(8, 64), torch.randn(8, 64))


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tricolo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_host_metrics_finalise_matches_oracle():
    from oracle import retrieval_oracle as RO
    from tricolo_b200.evaluation import construct_embeddings_matrix, metrics_from_ranks

    tuples = RO.make_val_shaped(seed=3, n_shapes=300, n_queries=1000, dim=128, round_bf16=True)
    ref = RO.compute_metrics(tuples)
    text, gal, labels, fit_labels, m2l, n, l2m = construct_embeddings_matrix("x", {"caption_embedding_tuples": tuples})
    r_text, r_gal, r_labels, _, _, _ = RO.build_matrices(tuples)
    assert np.array_equal(text, r_text) and np.array_equal(gal, r_gal) and np.array_equal(labels, r_labels)
    assert n == 1000 and l2m[m2l["m5"]] == "m5"
    got = metrics_from_ranks(ref["_indices"], ref["_rank"], labels, 5, fit_labels)
    for k in ("precision", "recall", "recall_rate", "ndcg"):
        assert np.array_equal(got[k], ref[k])
    assert got["mrr"] == ref["mrr"]


def test_rank_count_closed_form_matches_reference_order_finalise():
    """metrics_from_rank_counts (what the K5 device reduction feeds) against metrics_from_ranks on the same ranks."""
    from oracle import retrieval_oracle as RO
    from tricolo_b200.evaluation import metrics_from_rank_counts, metrics_from_ranks

    tuples = RO.make_val_shaped(seed=3, n_shapes=300, n_queries=1000, dim=128, round_bf16=True)
    ref = RO.compute_metrics(tuples)
    rank, labels = ref["_rank"], RO.build_matrices(tuples)[2]
    counts = [(rank == j + 1).sum() for j in range(5)]
    got = metrics_from_rank_counts(counts, float((1.0 / rank).sum()), len(rank))
    want = metrics_from_ranks(ref["_indices"], rank, labels, 5, np.arange(300))
    for k in ("precision", "recall", "recall_rate", "ndcg"):
        assert np.abs(got[k] - want[k]).max() <= 1e-14
    assert abs(got["mrr"] - want["mrr"]) <= 1e-14


def test_distance_flip_quirk():
    from tricolo_b200.evaluation.eval_retrieval import _flip_distances_like_reference

    v = np.arange(20.0).reshape(10, 2)
    assert np.array_equal(_flip_distances_like_reference(v, None), v[::-1])
    out = _flip_distances_like_reference(v, 4)
    assert np.array_equal(out[:4], v[:4][::-1]) and np.array_equal(out[8:], v[8:][::-1])
