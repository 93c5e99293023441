"""2-rank probe: is torch symmetric memory (peer-mapped buffers over NVLink) usable on this box?"""
import os, sys, time
import torch
import torch.distributed as dist

def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    print(rank, "can_access_peer", [torch.cuda.can_device_access_peer(dev.index, j) for j in range(world) if j != dev.index], flush=True)
    try:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(1 << 20, dtype=torch.bfloat16, device=dev)
        t.fill_(rank + 1)
        hdl = symm.rendezvous(t, group=dist.group.WORLD)
        print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad", len(hdl.signal_pad_ptrs), flush=True)
        hdl.barrier()
        peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.bfloat16)
        print(rank, "peer value", float(peer[0]), flush=True)
        # bandwidth of a peer copy (local -> remote)
        src = torch.ones(64 << 20, dtype=torch.bfloat16, device=dev)
        big = symm.empty(64 << 20, dtype=torch.bfloat16, device=dev)
        h2 = symm.rendezvous(big, group=dist.group.WORLD)
        dst = h2.get_buffer((rank + 1) % world, (64 << 20,), torch.bfloat16)
        h2.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            dst.copy_(src)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        print(rank, f"peer write {128 / dt / 1e3:.1f} GB/s", flush=True)
        h2.barrier()
    except Exception as e:
        print(rank, "symmetric memory FAILED:", type(e).__name__, str(e)[:300], flush=True)
    dist.barrier()
    dist.destroy_process_group()

main()
