"""Where does the multi-rank step spend its time?  torchrun --nproc-per-node 2 tools/dbg_ddp.py [batch]
Prints, per mode, the host enqueue time and the device time of one step (development aid)."""
import os, sys, time, torch, torch.distributed as dist
from types import SimpleNamespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from equiadapt_b200.images.canonicalization.discrete_group import GroupEquivariantImageCanonicalization
from equiadapt_b200.images.canonicalization_networks.custom_equivariant_networks import CustomEquivariantNetwork

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
net = CustomEquivariantNetwork((3, 96, 96), 32, 5, "rotation", 8, 3, device="cuda")
can = GroupEquivariantImageCanonicalization(net, SimpleNamespace(beta=1.0, input_crop_ratio=0.8, resize_shape=96), (3, 224, 224)).eval()
x = torch.rand(B, 3, 224, 224, device="cuda")


def step():
    y = can(x)
    z = can.invert_canonicalization(y, induced_rep_type="scalar")
    return z, can.get_prior_regularization_loss(), can.get_identity_metric()


for sync, prefetch in ((False, False), (True, False), (True, True)):
    can.sync_prior_across_ranks, can.prefetch_prior_allreduce = sync, prefetch
    with torch.no_grad():
        for _ in range(5): step()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        t0 = time.perf_counter(); e0.record()
        for _ in range(n): step()
        e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
    if rank == 0:
        print(f"world {world} batch {B} sync {sync} prefetch {prefetch}: host enqueue {1e3 * (t1 - t0) / n:.3f} ms/step, device {e0.elapsed_time(e1) / n:.3f} ms/step", flush=True)
if world > 1:
    dist.destroy_process_group()
