"""ncu target: one launch (after one warm-up) of every kernel family at the BASELINE sizes -- batch 8 x 480x640, 12 frame
pairs for ProjNN / CrossCheck, 16384^2 for NN.  python tools/experiments/all_ops_profile_target.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import connecting_the_dots_b200 as ctd
from connecting_the_dots_b200 import _lib, synth
tx = ctd.torchext
B, H, W = 8, 480, 640
dev = torch.device("cuda", 0)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
base = synth.make_batch(B, H, W)
d = {k: cu(base[k]) for k in ("im", "es", "ta", "go", "std", "pat_lcn", "disp")}
sums = torch.zeros(2, device=dev)
o1, o2, o3 = torch.empty(B, 1, H, W, device=dev), torch.empty(B, 1, H, W, device=dev), torch.empty(B, 1, H, W, device=dev)
st = torch.cuda.current_stream().cuda_stream
P = lambda t: t.data_ptr()
xyz, K, poses = synth.make_clouds(4, H, W)
pairs = [(i, j) for i in range(4) for j in range(4) if i != j]
x0 = cu(np.stack([synth.transform(xyz[i], poses[j]) for i, j in pairs])); x1 = cu(np.stack([xyz[j] for i, j in pairs])); Kd = cu(K)
idx = torch.empty(12, H, W, dtype=torch.int64, device=dev)
m = torch.empty(12 * H * W, dtype=torch.uint8, device=dev)
rng = np.random.RandomState(0)
p0, p1 = cu(rng.randn(16384, 3).astype(np.float32)), cu(rng.randn(16384, 3).astype(np.float32)); nno = torch.empty(16384, dtype=torch.int64, device=dev)
gd = synth.make_depth_pairs(B, H, W, seed=0); g = {k: cu(v) for k, v in gd.items()}; ray = tx.projection_rays(gd["Ki"], H, W).to(dev)
ga, gb, gs = torch.empty(B, 1, H, W, device=dev), torch.zeros(B, 1, H, W, device=dev), torch.zeros(2, 2, device=dev)
edge = torch.sigmoid(torch.randn(B, 1, H, W, device=dev))
pat = d["pat_lcn"][:1].contiguous()
vol = torch.empty(B, 128, H, W, device=dev) if "--xcorr" in sys.argv else None
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)   # 256 MB: evict L2 between the profiled launches
def run():
    c = lambda *a: _lib.call(*a)
    c("ctd_lcn_f32", P(d["im"]), P(o1), P(o2), B, H, W, 5, 0.05, st); flush.zero_()
    c("ctd_lcn_bwd_f32", P(d["im"]), P(o1), P(o2), P(d["go"]), P(d["go"]), P(o3), B, H, W, 5, 0.05, st); flush.zero_()
    c("ctd_photometric_fwd_bwd_masked_f32", P(d["es"]), P(d["ta"]), P(d["go"]), P(d["std"]), P(o1), P(o2), P(sums), B, 1, H, W, 9, 1, 0.5, st); flush.zero_()
    c("ctd_photometric_fwd_bwd_masked_f32", P(d["es"]), P(d["ta"]), P(d["go"]), P(d["std"]), P(o1), P(o2), P(sums), B, 1, H, W, 9, 3, 0.5, st); flush.zero_()
    c("ctd_photometric_fwd_f32", P(d["es"]), P(d["ta"]), P(o1), B, 1, H, W, 9, 3, 0.5, st); flush.zero_()
    c("ctd_photometric_bwd_f32", P(d["es"]), P(d["ta"]), P(d["go"]), P(o2), B, 1, H, W, 9, 3, 0.5, st); flush.zero_()
    c("ctd_warp_pattern_fwd_f32", P(pat), P(d["disp"]), P(o1), B, 1, H, W, H, W, st); flush.zero_()
    c("ctd_warp_pattern_bwd_f32", P(pat), P(d["disp"]), P(d["go"]), P(o2), B, 1, H, W, H, W, st); flush.zero_()
    c("ctd_pattern_similarity_f32", P(pat), P(d["disp"]), P(d["ta"]), P(d["std"]), P(d["std"]), P(o1), P(o2), P(o3), P(sums), B, 1, H, W, H, W, 3, 0.5, st); flush.zero_()
    c("ctd_proj_nn_f32", P(x0), P(x1), P(Kd), P(idx), 12, H, W, 3, st); flush.zero_()
    c("ctd_proj_nn_f32", P(x0), P(x1), P(Kd), P(idx), 12, H, W, 5, st); flush.zero_()
    c("ctd_crosscheck", P(idx), P(idx), P(m), 12 * H * W, 12 * H * W, st); flush.zero_()
    c("ctd_nn_f32", P(p0), P(p1), P(nno), 16384, 16384, st); flush.zero_()
    c("ctd_depth_similarity_f32", P(g["depth0"]), P(g["depth1"]), P(ray), P(g["K"]), P(g["R0"]), P(g["t0"]), P(g["R1"]), P(g["t1"]), P(ga), P(gb), P(gs[0]), B, H, W, 0.1, 1.0 / (B * H * W), 0, st); flush.zero_()
    c("ctd_disparity_loss_f32", P(d["disp"]), P(edge), P(o1), P(o2), P(sums), B, H, W, 1.0 / (B * H * W), st); flush.zero_()
    c("ctd_masked_sums_f32", P(d["es"]), P(d["std"]), B * H * W, P(sums), P(ws), st); flush.zero_()
    if "--xcorr" in sys.argv:
        c("ctd_xcorrvol_f32", P(d["ta"]), P(d["pat_lcn"]), P(vol), B, 1, H, W, 128, 9, st)
ws = torch.zeros(int(_lib.lib().ctd_masked_sums_workspace_bytes()), dtype=torch.uint8, device=dev)
_lib.set_option("xcorr_serial", 1)
run(); torch.cuda.synchronize()   # the profiled pass: run under `ncu -c N`, the kernels appear in the order of run()
