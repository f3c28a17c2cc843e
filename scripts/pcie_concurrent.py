"""Host <-> device copy rates with 1 .. N ranks copying at the same time (torchrun): does the end-to-end leg of
bench.py stop scaling because of the GPUs or because the ranks share the host side (PCIe root / IOMMU / DRAM)?

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 scripts/pcie_concurrent.py
"""
import json, os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)

def rate(active, reps=6):
    """GB/s each way of this rank while the ranks in `active` copy concurrently (others idle)."""
    dist.barrier(); torch.cuda.synchronize()
    r = 0.0
    if rank in active:
        both(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): both()
        torch.cuda.synchronize()
        r = n * reps / (time.perf_counter() - t0) / 1e9
    t = torch.tensor([r], dtype=torch.float64, device="cuda")
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    return [round(float(x), 1) for x in g]

out = {}
k = 1
while k <= world:
    res = rate(set(range(k)))
    out[f"{k}_ranks_concurrent"] = {"per_rank_GBps_each_way": res[:k], "aggregate_each_way": round(sum(res), 1)}
    k *= 2
out["each_rank_alone"] = [rate({r})[r] for r in range(world)]
if rank == 0:
    try:
        import subprocess
        out["nvidia_smi_topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout.splitlines()[:12]
        out["lscpu"] = [l for l in subprocess.run(["lscpu"], capture_output=True, text=True).stdout.splitlines()
                        if any(k in l for k in ("Model name", "Socket", "NUMA node", "CPU(s):"))][:8]
    except Exception as e:
        out["topo_error"] = repr(e)
    print(json.dumps(out))
dist.destroy_process_group()
