"""The reference suite's benchmark loop (UnifiedBenchmarkSuite.run_benchmark, src/benchmark/benchmark_suite.py:151-240)
over the B200 renderers, written as the CSV the suite's generate_report produces (columns of :244-255) plus Mrays/s,
Msamples/s, TFLOP/s and the fraction of the measured bf16 peak -- so the rows read side by side with the reference's
six methods.  Same defaults as the suite: resolutions 400x300 and 800x600, 64 and 128 samples per ray, 2 views;
timing through the renderer's own performance_monitor() (wall clock, device-synchronised), one untimed warm-up
view per configuration.  Run under gpurun:  python tools/suite_report.py > profiles/r1_suite_report.csv"""
import csv
import json
import math
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerf_dbr_b200 as nb  # noqa: E402

FLOP_PER_SAMPLE = 1_055_744


def test_poses(n):
    """generate_test_poses (benchmark_suite.py:132-149): rotation about y by 2 pi i / n, translation (0, 0, 4)."""
    out = []
    for i in range(n):
        th = 2.0 * math.pi * i / n
        out.append(torch.tensor([[math.cos(th), 0.0, math.sin(th), 0.0], [0.0, 1.0, 0.0, 0.0],
                                 [-math.sin(th), 0.0, math.cos(th), 4.0], [0.0, 0.0, 0.0, 1.0]], dtype=torch.float32))
    return out


def main():
    peak = 1671.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["bf16_tflops"]
    z = np.load(os.path.join(ROOT, "tests", "golden", "ckpt_lego_stuffed_fp16.npz"))
    sd = {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files}
    with tempfile.TemporaryDirectory() as tmp:
        ck = os.path.join(tmp, "ck.pth")
        torch.save({"coarse_model": sd, "fine_model": sd}, ck)
        w = csv.writer(sys.stdout)
        w.writerow(["Method", "Device", "Resolution", "Samples/Ray", "Render Time (s)", "Memory (MB)", "Rays/Second", "Device Info",
                    "Mrays/s", "Msamples/s", "TFLOP/s", "Fraction of bf16 peak"])
        for precision in ("bf16", "bf16x3", "fp32"):
            r = nb.B200Renderer(precision=precision)
            r.setup(ck)
            for res in ((400, 300), (800, 600)):
                for s in (64, 128):
                    poses = test_poses(2)
                    r.render_image(poses[0], res, s)                      # warm-up
                    times, mem = [], []
                    for pose in poses:
                        with r.performance_monitor():
                            rgb, depth = r.render_image(pose, res, s)
                        assert tuple(rgb.shape) == (res[1], res[0], 3) and torch.isfinite(rgb).all()
                        times.append(r.last_render_time)
                        mem.append(r.peak_memory_mb)
                    t, rays = float(np.mean(times)), res[0] * res[1]
                    tfl = FLOP_PER_SAMPLE * rays * s / t / 1e12
                    w.writerow([r.name, r.device, f"{res[0]}x{res[1]}", s, f"{t:.6f}", f"{np.mean(mem):.1f}", f"{rays / t:.0f}",
                                r.get_device_info(), f"{rays / t / 1e6:.3f}", f"{rays * s / t / 1e6:.1f}", f"{tfl:.1f}", f"{tfl / peak:.3f}"])


if __name__ == "__main__":
    main()
