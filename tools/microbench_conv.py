"""Micro-benchmark of the w4a8 / tf32 conv kernels on the LDM-4 top-level layer shape, with parts of the
epilogue switched off, to separate main-loop time from epilogue time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tfmq-dm_b200")]
import torch  # noqa: E402

from tfmq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


PROF = os.environ.get("TFMQ_IGEMM_PROF") is not None   # kernel prints its phase cycle counters: one timed call


def timeit(fn, iters=20):
    """GPU time per call: `iters` calls captured in a CUDA graph and replayed (a Python-side launch of these kernels costs
    more CPU time than the kernel runs, so back-to-back eager launches measure the host, not the device)."""
    if PROF:
        fn()
        torch.cuda.synchronize()
        fn()
        torch.cuda.synchronize()
        return 0.0
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for _ in range(iters):
                fn()
    graph.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        graph.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (3 * iters) * 1e3


def w4a8(n, h, w, cin, cout, ks, res, emb, stats):
    halo = ks // 2
    act = torch.randint(0, 255, (n, h + 2 * halo, w + 2 * halo, cin), dtype=torch.uint8, device=dev)
    wt = torch.randn(cout, ks * ks * cin, device=dev) * 0.05
    delta = wt.abs().amax(1) / 7.5
    zp = torch.full((cout,), 8.0, device=dev)
    _, packed, wsum = ops.pack_w4(wt, delta, zp)
    out = torch.zeros((n, h, w, cout), device=dev)
    aq = torch.tensor([0.02, 128.0], device=dev)
    bias = torch.zeros(cout, device=dev)
    e = torch.zeros((n, cout), device=dev) if emb else None
    st = [(torch.zeros((n, 32, 2), dtype=torch.float64, device=dev), cout // 32, 0)] if stats else None
    fn = lambda: ops.conv_w4a8(act, ks, packed, zp.to(torch.int32), delta, wsum, bias, aq, out, emb=e,  # noqa: E731
                               res=out if res else None, stats=st)
    us = timeit(fn)
    gop = 2 * n * h * w * cout * ks * ks * cin / 1e9
    print(f"w4a8 n={n} {h}x{w} {cin}->{cout} k{ks} res={int(res)} emb={int(emb)} stats={int(stats)}: {us:7.1f} us  "
          f"{gop / us * 1e-3:7.3f} POP/s", flush=True)


def fp(n, h, w, cin, cout, ks, res, passes=3, wlo=True):
    x = torch.randn((n, h, w, cin), device=dev)
    wt = torch.randn(cout, ks * ks * cin, device=dev) * 0.05
    hi, lo = ops.split_tf32(wt)
    out = torch.zeros((n, h, w, cout), device=dev)
    fn = lambda: ops.conv_fp(x, ks, 1, ks // 2, hi, lo if wlo else None, out, res=out if res else None,  # noqa: E731
                             passes=passes)
    us = timeit(fn)
    gf = 2 * n * h * w * cout * ks * ks * cin / 1e9
    print(f"tf32 n={n} {h}x{w} {cin}->{cout} k{ks} res={int(res)} passes={passes} wlo={int(wlo)}: {us:7.1f} us  "
          f"{gf / us * 1e-3:7.3f} PFLOP/s (algorithmic)", flush=True)


def i8gemm(mm, nn, kk):
    a = torch.randint(0, 255, (mm, kk), dtype=torch.uint8, device=dev)
    b = torch.randint(-8, 8, (nn, kk), dtype=torch.int8, device=dev)
    out = torch.empty((mm, nn), dtype=torch.int32, device=dev)
    us = timeit(lambda: ops.gemm_i8_peak(a, b, out))
    print(f"i8 gemm {mm}x{nn}x{kk}: {us:7.1f} us  {2 * mm * nn * kk / us * 1e-9:7.3f} POP/s", flush=True)


if __name__ == "__main__":
  i8gemm(8192, 8192, 8192)
  i8gemm(148 * 128, 256, 4096)
  i8gemm(8192, 8576, 8192)
  if os.environ.get('TFMQ_ONLY_GEMM'):
      sys.exit(0)
  w4a8(16, 64, 64, 224, 224, 3, True, True, True)
  w4a8(16, 64, 64, 224, 224, 3, True, False, False)
  w4a8(16, 64, 64, 224, 224, 3, False, False, False)
  w4a8(16, 64, 64, 224, 224, 1, False, False, False)
  w4a8(16, 64, 64, 224, 224, 1, True, False, False)
  w4a8(16, 32, 32, 448, 448, 3, True, True, True)
  w4a8(16, 16, 16, 672, 672, 3, True, True, True)
  w4a8(16, 8, 8, 896, 896, 3, True, True, True)
  w4a8(64, 64, 64, 256, 256, 3, False, False, False)
  fp(16, 64, 64, 224, 224, 3, False, 3, False)
  fp(16, 32, 32, 448, 1344, 1, False, 3, True)
  fp(16, 32, 32, 448, 1344, 1, False, 1, True)
  fp(16, 32, 32, 448, 448, 1, True, 3, True)
  fp(16, 64, 64, 448, 224, 1, False, 3, True)
