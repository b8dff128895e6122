"""w4a8 conv layers of LDM-4 that fold a residual in their epilogue (conv2 of every ResBlock), per launch:
  TFMQ_IGEMM_EPI_BUFS=2 python tools/microbench_epi.py     (two epilogue chunk buffers: residual load one chunk ahead)
  TFMQ_IGEMM_EPI_BUFS=3 python tools/microbench_epi.py     (three: two chunks ahead)"""
import os
import sys

sys.path[:0] = [os.path.dirname(os.path.abspath(__file__))]
import microbench_conv as mb  # noqa: E402

print("TFMQ_IGEMM_EPI_BUFS =", os.environ.get("TFMQ_IGEMM_EPI_BUFS", "(default)"))
for res in (True, False):
    mb.w4a8(16, 64, 64, 224, 224, 3, res, not res, True)
    mb.w4a8(16, 32, 32, 448, 448, 3, res, not res, True)
    mb.w4a8(16, 16, 16, 672, 672, 3, res, not res, True)
    mb.w4a8(16, 8, 8, 896, 896, 3, res, not res, True)
mb.w4a8(16, 64, 64, 672, 224, 3, True, False, True)
mb.w4a8(16, 32, 32, 1120, 448, 3, True, False, True)
