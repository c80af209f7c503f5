import os, sys, torch, time
sys.path.insert(0, os.getcwd())
from tricolo_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
gal = torch.randn(200000, 512, generator=g, device="cuda").bfloat16()
q = torch.randn(8192, 512, generator=g, device="cuda").bfloat16()
out = torch.empty(8192, 200000, device="cuda", dtype=torch.float32)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n
print("TCL_DBG", os.environ.get("TCL_DBG"), "gemm ms", t(lambda: ops.sim_gemm(q, gal, out=out)))
if not os.environ.get("TCL_DBG"):
    big = torch.empty(1<<30, device="cuda", dtype=torch.float32)
    ms = t(lambda: big.zero_()); print("zero_ 4GiB write GB/s", 4.295/ms*1e3)
    ms = t(lambda: big[:1<<29].copy_(big[1<<29:])); print("copy 2GiB->2GiB total GB/s", 4.295/ms*1e3)
    ms = t(lambda: torch.matmul(q, gal.t())); print("cublas bf16 gemm (bf16 out) ms", ms)
