"""Epilogue probes: the epilogue-bound conv shapes with parts of the kernel switched off (TFMQ_IGEMM_DBG: 64 = prologue and
teardown only, 16 = no TMA store, 32 = no fold / STS, 1 = no TMA loads, 8 = no unpack) next to a plain device memset / copy
of the output's size.  Graph-replayed GPU times."""
import os, sys
sys.path[:0] = [os.path.dirname(os.path.abspath(__file__))]
import torch
import microbench_conv as mb
print("TFMQ_IGEMM_DBG =", os.environ.get("TFMQ_IGEMM_DBG", "0"))
if os.environ.get("TFMQ_IGEMM_DBG", "0") == "0":
    for mbytes in (58.7, 235):
        x = torch.empty(int(mbytes * 1e6 / 4), device="cuda")
        y = torch.empty_like(x)
        print(f"memset {mbytes} MB: {mb.timeit(lambda: x.zero_()):7.1f} us   copy: {mb.timeit(lambda: y.copy_(x)):7.1f} us")
mb.w4a8(16, 64, 64, 224, 224, 1, False, False, False)
mb.w4a8(16, 64, 64, 224, 224, 3, False, False, False)
mb.w4a8(16, 64, 64, 224, 224, 3, True, True, True)
mb.w4a8(16, 32, 32, 448, 448, 3, True, True, True)
mb.w4a8(16, 8, 8, 896, 896, 3, True, True, True)
mb.fp(16, 32, 32, 448, 448, 1, True, 3, True)
mb.fp(16, 64, 64, 448, 224, 1, False, 3, True)
