import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device="cuda")
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast_ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else 0,
          "signal_pad", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
    t.fill_(rank + 1.0); hdl.barrier()
    if hdl.multicast_ptr:
        out = torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize(); print(rank, "multimem all_reduce ->", t[:2].tolist(), flush=True)
except Exception as e:
    print(rank, "symm_mem failed:", type(e).__name__, e, flush=True)
dist.barrier(); dist.destroy_process_group()
