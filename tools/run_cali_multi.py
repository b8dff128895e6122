"""Data-parallel calibration (`cali_model_multi`, reference quant/calibration.py:228-389) on N GPUs of one node over NCCL:
one process per GPU (mp.spawn, as the reference's scripts do), every rank calibrates on its 1/N slice of each timestep
interval, alpha gradients are SUM-all-reduced per reconstruction unit, activation deltas averaged, rank 0 saves.
Checks on the hardware run: every rank ends with bit-identical alphas and activation tables (the all-reduced gradients
drive identical Adam steps), and the checkpoint has the single-GPU schema.   python tools/run_cali_multi.py [world=2]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfmq-dm_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def data():
    from helpers import synth
    w = (synth.latents((32, 3, 32, 32), 31), torch.randint(0, 1000, (32,), generator=torch.Generator().manual_seed(1)).float())
    a = (synth.latents((64, 3, 32, 32), 32), torch.cat([torch.full((32,), 980.0), torch.full((32,), 960.0)]))
    return w, a


def worker(gpu, world, url, path, ret):
    from helpers import fp_model
    from tfmq_b200.quant.calibration import cali_model_multi
    from tfmq_b200.quant.quant_layer import QMODE, Scaler
    from tfmq_b200.quant.reconstruction_util import RLOSS
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    torch.manual_seed(0)                     # same randperm sequence on every rank, different data shards
    w, a = data()
    kw = dict(wq_params=dict(bits=4, channel_wise=True, scaler=Scaler.MINMAX),
              aq_params=dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True), softmax_a_bit=8,
              aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value], iters=20, batch_size=8, w=0.01, asym=True, warmup=0.2,
              opt_mode=RLOSS.MSE)
    t0 = time.time()
    ckpt = cali_model_multi(gpu, "nccl", world, url, 0, world, fp_model("cifar"), True, path, w, a, 32, True, kw)
    torch.cuda.synchronize()
    dt = time.time() - t0
    # every rank holds the same calibrated state: the all-reduced gradients drive identical Adam steps, the deltas are averaged
    ok = dist.is_initialized() and dist.get_world_size() == world and dist.get_backend() == "nccl"
    # (the reference averages the activation DELTAS only, quant/quant_model.py:127-132: zero points stay per rank)
    sig = [float(sum(v.double().sum() for k, v in ckpt["weight"].items() if k.endswith("alpha"))),
           float(sum(v.double().sum() for k, v in ckpt["act_0"].items() if k.endswith("delta")))]
    sigs = [None] * world
    dist.all_gather_object(sigs, sig)
    same = [all(s_[i] == sigs[0][i] for s_ in sigs) for i in range(2)]
    if gpu == 0:
        print("rank consistency: alphas identical", same[0], "| activation deltas identical", same[1], "| signatures", sigs, flush=True)
    ok = ok and all(same)
    ret[gpu] = (ok, dt)
    dist.barrier()
    dist.destroy_process_group()


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    assert torch.cuda.device_count() >= world, f"needs {world} GPUs, found {torch.cuda.device_count()}"
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    path = os.path.join(tempfile.mkdtemp(), "cali_multi.pth")
    ret = mp.Manager().dict()
    mp.spawn(worker, args=(world, f"tcp://127.0.0.1:{port}", path, ret), nprocs=world, join=True)
    ckpt = torch.load(path, map_location="cpu")
    alphas = [k for k in ckpt["weight"] if k.endswith("wqtizer.alpha")]
    acts = [k for k in ckpt if k.startswith("act_")]
    print(f"cali_model_multi over NCCL on {world} GPUs: {[(r, ok, round(dt, 1)) for r, (ok, dt) in sorted(ret.items())]} "
          f"(rank, nccl group ok and alphas / activation deltas bit-identical across ranks, seconds); checkpoint: {len(ckpt['weight'])} weight keys, {len(alphas)} alphas, tables {acts}, "
          f"{len(ckpt[acts[0]])} activation parameters per table")
    assert all(ok for ok, _ in ret.values()) and len(alphas) > 50 and len(acts) == 2
    # the same data on ONE GPU (world 1 through the same entry point): same schema
    ret1 = mp.Manager().dict()
    path1 = os.path.join(tempfile.mkdtemp(), "cali_single.pth")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(worker, args=(1, f"tcp://127.0.0.1:{port}", path1, ret1), nprocs=1, join=True)
    c1 = torch.load(path1, map_location="cpu")
    assert sorted(c1["weight"]) == sorted(ckpt["weight"]) and sorted(c1) == sorted(ckpt)
    d = max((c1["act_0"][k].float() - ckpt["act_0"][k].float()).abs().max().item() / max(1e-6, c1["act_0"][k].float().abs().max().item())
            for k in c1["act_0"] if k.endswith("delta"))
    print(f"single-GPU run of the same entry point: {round(ret1[0][1], 1)} s; identical key sets; activation deltas of table 0 differ by "
          f"at most {d:.3e} relative (each rank calibrates on its own slice; deltas are averaged)")


if __name__ == "__main__":
    main()
