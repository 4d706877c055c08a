"""end-to-end env step rate (pinned host actions in, reward / info rows and next actions out) against the number of env groups
(handles) kept in flight by HostEnvGroups: python profiles/e2e_groups.py [envs] [ticks] [G ...]   (run under torchrun for N > 1)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(dev)
barrier = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device(dev))
    barrier = dist.barrier
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 300
for G in [int(x) for x in sys.argv[3:]] or [1, 2, 3, 4, 6, 8]:
    for packed in (False, True):
        rate, errs = bench.e2e_double_buffered(n, ticks, 1500, dev, 0, rank, world, G, barrier, packed=packed)
        if rank == 0:
            print("world %d  groups %d  %s action rows: %.1f M env steps/s end to end (errs %d)" % (world, G, "uint8" if packed else "int32", rate / 1e6, errs), flush=True)
if world > 1:
    dist.destroy_process_group()
