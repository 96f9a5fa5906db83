"""First-contact GPU diagnostic: correctness numbers and rough timings for both precision modes,
printed even when something is off (tests only say pass/fail).  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from nerf_dbr_b200.host import ops, lib as L
from oracle import nerf_oracle as O
import ctypes

dev = torch.device("cuda", 0)
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ckpt_lego_stuffed_fp16.npz"))
weights = {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files}
net = ops.pack_weights({k: v.to(dev) for k, v in weights.items()}, dev)
torch.cuda.synchronize()
pose = O.generic_pose()
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def stats(tag, a, b):
    a, b = a.detach().cpu().double().numpy(), b.double().numpy()
    d = np.abs(a - b)
    mse = np.mean((a - b) ** 2)
    print(f"  {tag}: max {d.max():.3e} mean {d.mean():.3e} psnr {(-10*np.log10(mse) if mse > 0 else 99):.1f} "
          f"nan {int(np.isnan(a).sum())} ref[min {b.min():.3f} max {b.max():.3f}] got[min {np.nanmin(a):.3f} max {np.nanmax(a):.3f}]")


def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if which in ("all", "fp32"):
    for (W, H, S) in [(64, 48, 16), (64, 48, 32)]:
        ref = O.render_image(weights, pose, W, H, S)
        rgb, dep = ops.render_image(net, pose, W, H, S, mode=0)
        torch.cuda.synchronize()
        print(f"fp32 {W}x{H}x{S}")
        stats("rgb", rgb, ref[0]); stats("depth", dep, ref[1])
    ms = timed(lambda: ops.render_image(net, pose, 400, 300, 64, mode=0), 2)
    print(f"fp32 400x300x64: {ms:.2f} ms  {400*300*64*1.055744e6/ms/1e9:.2f} TFLOP/s")

if which in ("all", "bf16"):
    word = torch.zeros(1, dtype=torch.int32, device=dev)
    L.load_library().nerf_b200_set_watchdog_word(ctypes.c_void_p(word.data_ptr()))
    for (W, H, S) in [(64, 48, 128), (64, 48, 32), (64, 48, 16), (40, 30, 256)]:
        ref = O.render_image(weights, pose, W, H, S)
        try:
            rgb, dep = ops.render_image(net, pose, W, H, S, mode=1)
            torch.cuda.synchronize()
        except Exception as e:
            print("bf16 launch failed:", e, "watchdog word", hex(int(word.cpu().item()) & 0xffffffff) if False else "")
            raise
        print(f"bf16 {W}x{H}x{S}  watchdog={hex(int(word.item()) & 0xffffffff)}")
        stats("rgb", rgb, ref[0]); stats("depth", dep, ref[1])
        if S == 128:
            r = rgb.cpu().reshape(-1, 3); rr = ref[0].reshape(-1, 3)
            print("   first rays got", r[:3].tolist(), "ref", rr[:3].tolist())
    for (W, H, S, it) in [(400, 300, 64, 5), (800, 600, 128, 5)]:
        ms3 = timed(lambda: ops.render_image(net, pose, W, H, S, mode=2), 2)
        print(f"bf16x3 {W}x{H}x{S}: {ms3:.3f} ms  {W*H/ms3/1e3:.3f} Mrays/s")
        ms = timed(lambda: ops.render_image(net, pose, W, H, S, mode=1), it)
        print(f"bf16 {W}x{H}x{S}: {ms:.3f} ms  {W*H/ms/1e3:.3f} Mrays/s  {W*H*S*1.055744e6/ms/1e9:.1f} TFLOP/s")
