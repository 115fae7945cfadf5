"""Bus floor of the bench's end-to-end step on this box: 39.3 MB up and 39.3 MB down per step (pinned buffers), as whole
copies and cut into pieces, one direction and both.  python tools/experiments/pcie_floor.py"""
import time, json, torch
torch.cuda.set_device(0)
PLANE = 8 * 480 * 640
h_in = [torch.empty(PLANE).pin_memory() for _ in range(4)]
h_out = [torch.empty(PLANE).pin_memory() for _ in range(4)]
d_in = [torch.empty(PLANE, device="cuda") for _ in range(4)]
d_out = [torch.empty(PLANE, device="cuda") for _ in range(4)]
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, dn, pieces, n=30):
    def once():
        step = PLANE // pieces
        for p in range(pieces):
            sl = slice(p * step, (p + 1) * step)
            if up:
                with torch.cuda.stream(s_up):
                    for a, b in zip(d_in, h_in):
                        a[sl].copy_(b[sl], non_blocking=True)
            if dn:
                with torch.cuda.stream(s_dn):
                    for a, b in zip(h_out, d_out):
                        a[sl].copy_(b[sl], non_blocking=True)
        s_up.synchronize(); s_dn.synchronize()
    for _ in range(3):
        once()
    t0 = time.perf_counter()
    for _ in range(n):
        once()
    return round((time.perf_counter() - t0) / n * 1e3, 4)
res = {}
for pieces in (1, 2, 4, 8):
    res["up_only_pieces%d" % pieces] = run(True, False, pieces)
    res["down_only_pieces%d" % pieces] = run(False, True, pieces)
    res["both_pieces%d" % pieces] = run(True, True, pieces)
print(json.dumps(res, indent=1))
