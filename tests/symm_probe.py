"""Probe (not a test; run under torchrun on 2+ GPUs): does torch's symmetric memory give peer and MULTICAST pointers here?"""
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    names = [n for n in dir(hdl) if not n.startswith("_")]
    if rank == 0:
        print("handle attrs:", names)
    info = {"rank": rank, "world": hdl.world_size, "buffer_ptrs": [hex(p) for p in hdl.buffer_ptrs],
            "multicast_ptr": hex(getattr(hdl, "multicast_ptr", 0) or 0),
            "signal_pad_ptrs": [hex(p) for p in getattr(hdl, "signal_pad_ptrs", [])]}
    print(info, flush=True)
    # sanity: multimem all-reduce op shipped with torch, if present
    t.fill_(rank + 1.0)
    hdl.barrier()
    try:
        out = torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        print(rank, "multimem_all_reduce_ ->", float(out[0]), flush=True)
    except Exception as e:   # noqa: BLE001
        print(rank, "multimem_all_reduce_ failed:", repr(e)[:300], flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
