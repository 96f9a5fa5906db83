"""Per-layer timeline of the tensor-core kernel (CTA 0, first tiles).  Run under gpurun."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nerf_dbr_b200.host import ops, lib as L
from nerf_dbr_b200.host.synthetic import orbit_pose

dev = torch.device("cuda", 0)
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ckpt_lego_stuffed_fp16.npz"))
net = ops.pack_weights({k: torch.from_numpy(z[k].astype(np.float32)).to(dev) for k in z.files}, dev)
pose = orbit_pose(1, 40)
TRAIN = len(sys.argv) > 1 and sys.argv[1] == "train"      # timeline of the TRAIN forward variant instead
FP8 = len(sys.argv) > 1 and sys.argv[1] == "fp8"          # timeline of the FP8-mode kernel
if FP8:
    net8 = ops.pack_weights_fp8({k: torch.from_numpy(z[k].astype(np.float32)).to(dev) for k in z.files}, dev, packed=net)
if TRAIN:
    import nerf_dbr_b200 as nb
    model = nb.NeRFModel().to(dev)
    g = torch.Generator().manual_seed(0)
    ro = torch.zeros(4096, 3) + torch.tensor([0.0, 0.0, 4.0])
    rd = torch.nn.functional.normalize(torch.randn(4096, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, -1.0]), dim=-1)
    tgt = torch.rand(4096, 3, generator=g)

    def run():
        ops.train_fwd_bwd(model, ro.to(dev), rd.to(dev), tgt.to(dev), 128, mode=L.BF16)
elif FP8:
    def run():
        ops.render_image(net8, pose, 800, 600, 128, mode=L.FP8)
else:
    def run():
        ops.render_image(net, pose, 800, 600, 128, mode=1)
run()
torch.cuda.synchronize()
buf = torch.zeros(6 * 9 * 8, dtype=torch.int64, device=dev)
L.load_library().nerf_b200_set_trace_buffer(ctypes.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
L.load_library().nerf_b200_set_trace_buffer(None)
t = buf.cpu().numpy().reshape(6, 9, 8).astype(np.int64)
t0 = t[t > 0].min()
print("slots: mma_first h0_commit last_commit | epi_h0_acc epi_h0_arrive epi_done epi_h1_acc || layer_period  E_h0(acc->arrive)  E_h1(acc->done)")
for ti in range(2, 6):
    for l in range(9):
        r = t[ti, l]
        rel = np.where(r > 0, r - t0, -1)
        nxt = t[ti, l + 1, 0] if l < 8 else (t[ti + 1, 0, 0] if ti + 1 < 6 else 0)
        print(f"{ti:2d} {l:2d} " + " ".join(f"{int(v):9d}" for v in rel[:7]) + f" || {int(nxt - r[0]) if nxt else -1:8d} "
              f"{int(r[4]-r[3]) if r[4] > 0 else -1:8d} {int(r[5]-r[6]) if r[6] > 0 else -1:8d}")
print("tile period (cycles):", (t[5, 0, 0] - t[2, 0, 0]) / 3)
