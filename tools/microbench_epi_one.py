"""One epilogue-bound w4a8 layer (224->224 3x3 @ 64x64, batch 16, residual + statistics) for an ncu source-level capture:
  ncu --set full --import-source on --clock-control none -k regex:igemm -c 1 -s 3 -o gpurun_out/epi python tools/microbench_epi_one.py"""
import os, sys
sys.path[:0] = [os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."), os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tfmq-dm_b200"), os.path.dirname(os.path.abspath(__file__))]
import microbench_conv as mb
mb.w4a8(16, 64, 64, 224, 224, 3, True, False, True)
