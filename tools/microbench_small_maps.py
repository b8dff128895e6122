"""The small-map w4a8 layers of one LDM-4 step (8x8 and 16x16, batch 16), per launch:
  TFMQ_IGEMM_KSPLIT=0|2|3|4 python tools/microbench_small_maps.py      (0: split-K off, default: the cost model decides)"""
import os, sys
sys.path[:0] = [os.path.dirname(os.path.abspath(__file__))]
import microbench_conv as mb
print("TFMQ_IGEMM_KSPLIT =", os.environ.get("TFMQ_IGEMM_KSPLIT", "(default)"), " TFMQ_IGEMM_XFGW =", os.environ.get("TFMQ_IGEMM_XFGW", "(default)"))
mb.w4a8(16, 8, 8, 896, 896, 3, True, False, True)
mb.w4a8(16, 8, 8, 896, 896, 3, False, True, True)
mb.w4a8(16, 8, 8, 1792, 896, 3, False, True, True)
mb.w4a8(16, 8, 8, 672, 896, 3, False, True, True)
mb.w4a8(16, 16, 16, 672, 672, 3, True, False, True)
mb.w4a8(16, 16, 16, 1344, 672, 3, False, True, True)
mb.w4a8(16, 16, 16, 448, 672, 3, False, True, True)
