"""Probe: can this box do peer writes through torch symmetric memory / CUDA IPC? (torchrun, 2 ranks)"""
import os, sys, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n = 64 * 1024 * 1024
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty((world, n), dtype=torch.bfloat16, device=dev)
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
    src = torch.full((n,), float(rank + 1), dtype=torch.bfloat16, device=dev)
    peers = [hdl.get_buffer(r, (world, n), torch.bfloat16) for r in range(world)]
    side = torch.cuda.Stream()
    for it in range(3):
        hdl.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(side):
            for r in range(world):
                peers[r][rank].copy_(src, non_blocking=True)
        side.synchronize()
        hdl.barrier()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            print(f"symm peer copies: {dt*1e3:.3f} ms for {n*2/1e6:.0f} MB per peer -> {n*2/dt/1e9:.1f} GB/s per direction")
    ok = all(float(t[r][0]) == r + 1 and float(t[r][-1]) == r + 1 for r in range(world))
    print(f"rank {rank}: symmetric memory ok={ok}")
except Exception as e:
    print(f"rank {rank}: symmetric memory failed: {type(e).__name__}: {e}")
dist.destroy_process_group()
