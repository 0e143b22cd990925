"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: row sharding, the rendezvous of the
communicator id through torch.distributed, and the loud failure of the collective set-up when there is
no CUDA device.  The device-side protocol itself is checked on GPUs by tests/dist_check.py."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

import easy_neural_ode_b200 as eno

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_partitions_the_batch():
    got = [eno.shard_rows(100, 4, r) for r in range(4)]
    idx = torch.arange(100)
    assert torch.equal(torch.cat([idx[s] for s in got]), idx)
    assert all((s.stop - s.start) == 25 for s in got)
    # uneven shards: 100 rows on 8 ranks = 13, 13, 13, 13, 12, 12, 12, 12, contiguous and covering
    got = [eno.shard_rows(100, 8, r) for r in range(8)]
    assert [s.stop - s.start for s in got] == [13, 13, 13, 13, 12, 12, 12, 12]
    assert torch.equal(torch.cat([idx[s] for s in got]), idx)
    with pytest.raises(ValueError):
        eno.shard_rows(3, 8, 0)


def test_init_distributed_requires_a_process_group():
    with pytest.raises(RuntimeError):
        eno.init_distributed()


WORKER = textwrap.dedent('''
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    import easy_neural_ode_b200 as eno
    from easy_neural_ode_b200 import _cabi
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # 1. the id travels from rank 0 to everybody exactly as init_distributed ships it
    box = [bytes(range(128)) if rank == 0 else bytes(128)]
    dist.broadcast_object_list(box, src=0)
    assert box[0] == bytes(range(128))
    # 2. shards are disjoint and cover the batch; a sharded sum equals the global sum in fp64
    x = torch.arange(96, dtype=torch.float64) ** 2
    part = x[eno.shard_rows(96, world, rank)].sum().reshape(1)
    parts = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, part)
    assert float(sum(parts)) == float(x.sum())
    # 2b. the loss terms of the sharded latent step (latent_step.LatentGen._global_sum): the VALUE is the sum over ranks,
    #     the derivative wrt this rank's term is 1 - so every rank differentiates the same global loss wrt its own rows and
    #     the all-reduced parameter gradient is the gradient of the concatenated batch
    import types
    from easy_neural_ode_b200 import latent_step
    w = torch.tensor([1.0 + rank, 2.0], dtype=torch.float64, requires_grad=True)
    local = (w * w).sum()
    tot = latent_step.LatentGen._global_sum(types.SimpleNamespace(world=world), local)
    assert float(tot) == sum((1.0 + r) ** 2 + 4.0 for r in range(world))
    loss = torch.log(tot)
    g, = torch.autograd.grad(loss, w)
    assert torch.allclose(g, 2 * w.detach() / float(tot))
    # 3. without a GPU the collective set-up fails loudly (no CPU fallback)
    try:
        eno.init_distributed()
        raise SystemExit("init_distributed succeeded without a CUDA device")
    except (_cabi.EnoLibraryError, RuntimeError, AssertionError):
        pass
    dist.destroy_process_group()
    print("worker", rank, "ok")
''') % ROOT


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU/gloo variant; GPUs run tests/dist_check.py")
def test_gloo_world2_host_logic(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


@pytest.mark.gpu
@pytest.mark.parametrize("reg_type", ["our", "fin"])
def test_two_gpu_collective_step_matches_single_device(reg_type):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    if reg_type == "fin" and os.environ.get("ENO_PATH") == "stream":
        pytest.skip("the 'fin' adjoint exists on the cluster-resident path only")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", {"our": "29534", "fin": "29536"}[reg_type],   # no port reuse
           os.path.join(ROOT, "tests", "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, ENO_DIST_REGTYPE=reg_type))
    assert res.returncode == 0 and "DIST PARITY OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("rows", [None, "7"])
def test_two_gpu_collective_flat_engine_matches_single_device(rows):
    """ConcatSquash / ConcatConv2D split solves sharded over 2 GPUs (even shards, then 7 rows = 4 + 3) against the
    single-device solve of the concatenated batch (tests/dist_check_flat.py)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29538" if rows is None else "29540",
           os.path.join(ROOT, "tests", "dist_check_flat.py")]
    env = dict(os.environ) if rows is None else dict(os.environ, ENO_DIST_FLAT_B=rows)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0 and "FLAT DIST PARITY OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
