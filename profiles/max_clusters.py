"""Active clusters the B200 can hold for the producer/consumer backward kernel's footprint, per cluster size."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tricolo_b200 import _lib
torch.cuda.init()
for cs in (1, 2, 4, 8, 16):
    n = C.c_int(0)
    try:
        _lib.check(_lib.LIB.tcl_debug_max_clusters(cs, C.byref(n)))
        print(f"cluster size {cs}: {n.value} clusters = {n.value * cs} SMs")
    except Exception as e:
        print(f"cluster size {cs}: {e}")
