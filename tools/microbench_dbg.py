"""Timing experiments on the w4a8 conv main loop (TFMQ_IGEMM_DBG bits: 1 no TMA, 2 B only, 4 A only, 8 no unpack).
Results are WRONG by construction for dbg != 0; only the times mean anything."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1 or sys.argv[1] != "child":
    for dbg in [int(a) for a in sys.argv[1:]] or (0, 1, 8, 9):
        env = dict(os.environ, TFMQ_IGEMM_DBG=str(dbg))
        print(f"== dbg {dbg}", flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=env)
    sys.exit(0)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tfmq-dm_b200"), os.path.join(ROOT, "tools")]
import microbench_conv as mb  # noqa: E402

mb.w4a8(16, 64, 64, 224, 224, 3, False, False, False)
mb.w4a8(16, 64, 64, 224, 224, 3, True, False, True)
mb.w4a8(16, 32, 32, 448, 448, 3, False, False, False)
mb.w4a8(16, 32, 32, 896, 448, 3, False, False, False)
mb.w4a8(16, 16, 16, 896, 896, 3, False, False, False)
mb.w4a8(16, 8, 8, 896, 896, 3, False, False, False)
if os.environ.get("TFMQ_IGEMM_DBG") == "0":
    mb.fp(16, 32, 32, 448, 1344, 1, False, 3, True)
    mb.fp(16, 64, 64, 448, 224, 1, False, 3, True)
