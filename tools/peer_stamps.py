"""Development (2+ GPUs, libeno_phases.so): where one peer-memory exchange of the adjoint spends its cycles.
  ENO_LIB=.../libeno_phases.so python -m torch.distributed.run --nproc-per-node 2 ... tools/peer_stamps.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ENO_LIB"] = os.path.join(ROOT, "easy-neural-ode_b200", "libeno_phases.so")
import ctypes as C, torch, torch.distributed as dist
import easy_neural_ode_b200 as eno
from easy_neural_ode_b200 import mnist_step, _cabi
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eno.init_distributed()
B = 100
gen = torch.Generator().manual_seed(rank)
images = torch.randint(0, 256, (B, 28, 28, 1), generator=gen, dtype=torch.uint8).to(dev)
labels = torch.randint(0, 10, (B,), generator=gen).to(dev)
eps = (torch.randint(0, 2, (B, 784), generator=gen).float() * 2 - 1).to(dev)
m = mnist_step.MnistNode(device=dev, world=world)
L = _cabi.lib()
L.eno_debug_peer_stamps.argtypes = [C.POINTER(C.c_longlong)]
for it in range(4):
    m.loss_and_grads(images, labels, eps)
    st = (C.c_longlong * 8)()
    L.eno_debug_peer_stamps(st)
    d = [st[i + 1] - st[i] for i in range(6)]
    print(f"[rank {rank}] CTA0: push {d[0]} wait-all {d[1]} slice {d[2]} | (ticket gap {d[3]}) round2 {d[4]} finalise {d[5]}  total {st[6] - st[0]} cycles", flush=True)
dist.barrier()
eno.shutdown_distributed()
dist.destroy_process_group()
