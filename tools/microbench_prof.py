import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/tfmq-dm_b200", "/root/repo/tools"]
import microbench_conv as mb
mb.w4a8(16, 64, 64, 224, 224, 1, False, False, False)
mb.w4a8(16, 64, 64, 224, 224, 3, False, False, False)
mb.w4a8(16, 64, 64, 224, 224, 3, True, False, True)
mb.w4a8(16, 32, 32, 448, 448, 3, True, False, True)
