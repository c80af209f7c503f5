"""HBM fraction of the standalone top-k / rank kernel (K4) at the shard sizes of BASELINE configs[4]:
   python profiles/topk_time.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tricolo_b200 import _lib, ops
from tricolo_b200.evaluation.eval_retrieval import two_kernel_block_queries
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6550.4
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for g in (25000, 50000, 100000, 200000):
    rows = two_kernel_block_queries(g)
    ld = (g + 31) // 32 * 32
    s = torch.randn(rows, ld, device="cuda")
    lab = torch.randint(0, g, (rows,), device="cuda")
    for _ in range(2):
        ops.topk_rank(s, g, 5, lab)
    _lib.profile_enable(True)
    for _ in range(5):
        flush.zero_()
        ops.topk_rank(s, g, 5, lab)
    torch.cuda.synchronize()
    ms, n = _lib.profile_read()["topk_rank"]
    _lib.profile_enable(False)
    gbs = (rows * g * 4 + rows * 48) / (ms / n * 1e-3) / 1e9
    print(f"gallery {g:7d} rows/launch {rows:6d}: {ms / n:.3f} ms  {gbs:7.0f} GB/s  {gbs / peak:.3f} of the measured HBM peak")
    del s
