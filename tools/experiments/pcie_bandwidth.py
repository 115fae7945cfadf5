import torch, time
n = 64 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10, chunk=None):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if chunk is None:
            if h2d:
                with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
        else:
            for o in range(0, n, chunk):
                if h2d:
                    with torch.cuda.stream(s1): d1[o:o+chunk].copy_(h1[o:o+chunk], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2): h2[o:o+chunk].copy_(d2[o:o+chunk], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    return reps * n * (h2d + d2h) / dt / 1e9
for args in ((1,0),(0,1),(1,1)):
    run(*args, reps=2)
    print("h2d,d2h", args, "GB/s total: %.1f" % run(*args), " 1.2MB chunks: %.1f" % run(*args, chunk=1228800))
