import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/tfmq-dm_b200", "/root/repo/tools"]
import microbench_conv as mb
mb.w4a8(16, 32, 32, 896, 448, 3, False, False, False)
if not os.environ.get("ONE"):
    mb.w4a8(16, 64, 64, 448, 448, 3, False, False, False)
