"""Host<->device copy bandwidth with N ranks copying AT THE SAME TIME (one rank per GPU, torchrun): what limits the
end-to-end (host-buffer) figure when the GPUs of one box are used together.  Pinned 64 MiB buffers, CUDA events, every
rank starts after a barrier; prints one JSON object on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 \
        tools/experiments/pcie_bandwidth_ranks.py [--bind]
--bind: set the thread's CPU affinity to the GPU's NVML-reported ideal CPUs before allocating the pinned buffers."""
import json, os, sys, time
import torch, torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
info = {"rank": rank}
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(lr)
    n_cpu = os.cpu_count() or 1
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
    cpus = [i for i in range(n_cpu) if (mask[i // 64] >> (i % 64)) & 1]
    info["gpu_ideal_cpus"] = "%d-%d (%d)" % (cpus[0], cpus[-1], len(cpus)) if cpus else "none"
    try:
        info["pcie_gen_width"] = "gen%d x%d" % (pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h), pynvml.nvmlDeviceGetCurrPcieLinkWidth(h))
    except Exception:
        pass
    if "--bind" in sys.argv and cpus:
        os.sched_setaffinity(0, cpus)
        info["bound"] = True
except Exception as e:
    info["nvml"] = str(e)
info["sched_affinity"] = len(os.sched_getaffinity(0))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 64 << 20
h1, h2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
h1.fill_(1); h2.fill_(2)   # first touch on this thread's NUMA node
d1, d2 = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    return reps * n * (h2d + d2h) / dt / 1e9
for name, a in (("h2d", (1, 0)), ("d2h", (0, 1)), ("both", (1, 1))):
    run(*a, reps=2)
    info[name + "_gbs"] = round(run(*a), 1)
if world > 1:
    allinfo = [None] * world
    dist.all_gather_object(allinfo, info)
    dist.destroy_process_group()
else:
    allinfo = [info]
if rank == 0:
    tot = {k: round(sum(i[k] for i in allinfo), 1) for k in ("h2d_gbs", "d2h_gbs", "both_gbs")}
    print(json.dumps({"ranks": world, "bind": "--bind" in sys.argv, "host_cpus": os.cpu_count(), "sum_over_ranks": tot, "per_rank": allinfo}))
