"""Multi-GPU parity check (run under torchrun, one process per GPU; not collected by pytest directly):
the sharded, collectively-controlled train step must reproduce the single-device step on the
concatenated batch - identical iteration counts / NFE, values within 1e-5.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import easy_neural_ode_b200 as eno  # noqa: E402
from easy_neural_ode_b200 import mnist_step  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    Bg = int(os.environ.get("ENO_DIST_B", "96"))
    tol = float(os.environ.get("ENO_DIST_TOL", "1.4e-8"))
    # ENO_DIST_REGTYPE=fin runs the odeint_sepaux / 'fin' arm (eps_bar block, two scalar states per sample)
    kw = (dict(reg="none", reg_type="fin", lam_fro=1e-2, lam_kin=1e-2) if os.environ.get("ENO_DIST_REGTYPE") == "fin"
          else dict(reg="r3", lam=6e-5))
    gen = torch.Generator().manual_seed(0)
    images = torch.randint(0, 256, (Bg, 28, 28, 1), generator=gen, dtype=torch.uint8).to(dev)
    labels = torch.randint(0, 10, (Bg,), generator=gen).to(dev)
    eps = (torch.randint(0, 2, (Bg, 784), generator=gen).float() * 2 - 1).to(dev)

    # single-device reference on the full batch (every rank computes it redundantly, before going collective)
    ref_model = mnist_step.MnistNode(rtol=tol, atol=tol, device=dev, **kw)
    want = ref_model.loss_and_grads(images, labels, eps)
    want_f, want_b = dict(want["fwd_stats"]), dict(want["bwd_stats"][0])

    eno.init_distributed()
    sl = eno.shard_rows(Bg, world, rank)
    model = mnist_step.MnistNode(rtol=tol, atol=tol, device=dev, world=world, **kw)
    if Bg % world:   # uneven shards (e.g. ENO_DIST_B=100 on 8 GPUs: 13,13,13,13,12,12,12,12): declare the global batch
        eno.set_shards(Bg)
        model.rows_global = Bg
    got = model.loss_and_grads(images[sl], labels[sl], eps[sl])
    f, b = got["fwd_stats"], got["bwd_stats"][0]
    errs = dict(y1=rel(got["y1"], want["y1"][sl]), g_ode=rel(got["grads"]["ode"], want["grads"]["ode"]),
                g_w=rel(got["grads"]["post_w"], want["grads"]["post_w"]))
    ok = (f["nfe"] == want_f["nfe"] and f["n_iters"] == want_f["n_iters"] and f["n_accepted"] == want_f["n_accepted"]
          and b["n_iters"] == want_b["n_iters"] and b["n_accepted"] == want_b["n_accepted"]
          and all(v < 1e-5 for v in errs.values()))
    # all ranks must report the same controller trajectory
    t = torch.tensor([f["last_dt"], f["last_t"], b["last_dt"], b["last_t"]], device=dev, dtype=torch.float64)
    ts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(ts, t)
    same = all(torch.equal(ts[0], x) for x in ts)
    print(f"[rank {rank}] fwd {f['n_iters']}/{want_f['n_iters']} iters nfe {f['nfe']}/{want_f['nfe']}  bwd {b['n_iters']}/"
          f"{want_b['n_iters']} iters  errs {errs}  ranks_identical {same}  launches fwd {f['n_launches']} bwd {b['n_launches']}",
          flush=True)
    # timing of the collective step
    for _ in range(3):
        model.loss_and_grads(images[sl], labels[sl], eps[sl])
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        model.loss_and_grads(images[sl], labels[sl], eps[sl])
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"collective train step, B_global {Bg} on {world} GPUs: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
    eno.shutdown_distributed()
    dist.destroy_process_group()
    if not (ok and same):
        sys.exit(1)
    if rank == 0:
        print("DIST PARITY OK", flush=True)


if __name__ == "__main__":
    main()
