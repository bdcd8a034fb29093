"""Host<->device ceiling of the box with N GPUs busy at once: raw duplex cudaMemcpyAsync between pinned host buffers and HBM.
  one process per GPU :  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/exp_pcie_multi.py
  one process, N GPUs :  python profiles/exp_pcie_multi.py --single N
Prints JSON lines: per-direction and duplex GB/s per GPU and the aggregate -- what the end-to-end path (1 GB in + 1 GB out per
GPU and step) can reach at best."""
import json, os, sys, time
import torch

MB = 1 << 20
n = 512 * MB


def bufs(dev):
    torch.cuda.set_device(dev)
    return dict(dev=dev, hi=torch.empty(n, dtype=torch.uint8).pin_memory(), ho=torch.empty(n, dtype=torch.uint8).pin_memory(),
                di=torch.empty(n, dtype=torch.uint8, device=f"cuda:{dev}"), do=torch.empty(n, dtype=torch.uint8, device=f"cuda:{dev}"),
                s1=torch.cuda.Stream(device=dev), s2=torch.cuda.Stream(device=dev))


def run(bs, h2d, d2h, reps, barrier=None):
    for b in bs:
        torch.cuda.synchronize(b["dev"])
    if barrier:
        barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        for b in bs:
            if h2d:
                with torch.cuda.stream(b["s1"]):
                    b["di"].copy_(b["hi"], non_blocking=True)
            if d2h:
                with torch.cuda.stream(b["s2"]):
                    b["ho"].copy_(b["do"], non_blocking=True)
    for b in bs:
        torch.cuda.synchronize(b["dev"])
    if barrier:
        barrier()
    return n * reps / (time.perf_counter() - t0) / 1e9  # GB/s per GPU and direction


if "--single" in sys.argv:
    k = int(sys.argv[sys.argv.index("--single") + 1])
    bs = [bufs(d) for d in range(k)]
    run(bs, True, True, 2)
    r = {"mode": "one process", "gpus": k, "h2d_alone": run(bs, True, False, 6), "d2h_alone": run(bs, False, True, 6), "duplex_each_way": run(bs, True, True, 6)}
    r["aggregate_duplex_both_ways"] = r["duplex_each_way"] * 2 * k
    print(json.dumps({k_: (round(v, 1) if isinstance(v, float) else v) for k_, v in r.items()}))
else:
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if "--split-affinity" in sys.argv and world > 1:
        cores = sorted(os.sched_getaffinity(0)); per = max(1, len(cores) // world)
        os.sched_setaffinity(0, cores[local * per:(local + 1) * per])
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    barrier = (lambda: dist.barrier()) if world > 1 else None
    bs = [bufs(local)]
    run(bs, True, True, 2, barrier)
    r = {"mode": "one process per GPU" + (" (split affinity)" if "--split-affinity" in sys.argv else ""), "gpus": world,
         "h2d_alone": run(bs, True, False, 6, barrier), "d2h_alone": run(bs, False, True, 6, barrier), "duplex_each_way": run(bs, True, True, 6, barrier)}
    r["aggregate_duplex_both_ways"] = r["duplex_each_way"] * 2 * world
    if rank == 0:
        print(json.dumps({k_: (round(v, 1) if isinstance(v, float) else v) for k_, v in r.items()}))
    if world > 1:
        dist.destroy_process_group()
