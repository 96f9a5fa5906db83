"""GPU: the B200 renderer inside the reference's OWN benchmark suite.  The unmodified
``UnifiedBenchmarkSuite`` (src/benchmark/benchmark_suite.py:45-94, 151-302, imported from the vendored sources,
tools/vendor_reference.sh) gets ``B200Renderer`` appended to its renderer list exactly as INTEGRATION.md section 2
tells a maintainer to, then runs ``run_benchmark`` on a tiny grid next to "PyTorch CPU" and writes its report.
Checked: the CSV the suite itself writes has a row per method with the suite's columns, and the images the suite
itself saved for the renderers agree pixel for pixel (8-bit) with the reference renderer's."""
import os

import numpy as np
import pytest
import torch

from conftest import load_npz

pytestmark = pytest.mark.gpu


def test_b200_renderer_in_the_reference_benchmark_suite(tmp_path, monkeypatch):
    from oracle import refload
    if refload.reference_root() is None:
        pytest.skip("reference sources not vendored on this box (tools/vendor_reference.sh)")
    refload.import_reference()
    import pandas as pd
    from PIL import Image
    from src.benchmark.benchmark_suite import UnifiedBenchmarkSuite
    from src.benchmark.pytorch_renderers import PyTorchCPURenderer
    import nerf_dbr_b200 as nb

    monkeypatch.chdir(tmp_path)                           # the suite writes outputs/ relative to the working directory
    z = load_npz("ckpt_lego_stuffed_fp16.npz")
    lego = {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files}
    ck = str(tmp_path / "final_model.pth")
    torch.save({"coarse_model": lego, "fine_model": lego}, ck)

    suite = UnifiedBenchmarkSuite()
    suite.renderers.append(PyTorchCPURenderer())          # what add_available_renderers() always adds (:50)
    # ---- INTEGRATION.md section 2, verbatim -------------------------------------------------
    try:
        from nerf_dbr_b200 import B200Renderer
        suite.renderers.append(B200Renderer("bf16"))      # tcgen05 tensor cores
        suite.renderers.append(B200Renderer("fp32"))      # CUDA-core parity mode (max-abs <= 1e-4)
    except (ImportError, RuntimeError) as e:
        pytest.fail(f"B200 renderer not available: {e}")
    # ------------------------------------------------------------------------------------------
    suite.run_benchmark(ck, resolutions=[(64, 48)], samples_per_ray_options=[16], n_views=2)
    # matplotlib is not installed in this image: the plotting half of generate_report is skipped, the CSV half runs as is
    monkeypatch.setattr(suite, "_create_performance_plots", lambda df: None)
    df = suite.generate_report()

    assert list(df.columns) == ["Method", "Device", "Resolution", "Samples/Ray", "Render Time (s)", "Memory (MB)",
                                "Rays/Second", "Device Info"]
    csv = pd.read_csv(tmp_path / "outputs" / "benchmark_results.csv")
    assert list(csv["Method"]) == ["PyTorch CPU", "B200 BF16", "B200 FP32"]
    assert list(csv["Device"]) == ["cpu", "cuda", "cuda"] and set(csv["Resolution"]) == {"64x48"}
    assert (csv["Rays/Second"] > 0).all() and (csv["Render Time (s)"] > 0).all() and (csv["Memory (MB)"] > 0).all()
    assert all("CUDA" in s for s in csv["Device Info"][1:])
    # the sample renders the suite saved: view 0 and 1 of every method (benchmark_suite.py:96-125)
    ref = {}
    for method, tol in (("PyTorch_CPU", 0), ("B200_FP32", 1), ("B200_BF16", 6)):
        for view in (0, 1):
            path = tmp_path / "outputs" / "sample_renders" / method / f"view_{view}_rgb.png"
            assert path.exists(), path
            img = np.asarray(Image.open(path)).astype(np.int32)
            assert img.shape == (48, 64, 3)
            if method == "PyTorch_CPU":
                ref[view] = img
                assert img.std() > 5                       # not a blank frame
            else:
                d = np.abs(img - ref[view])
                print(f"{method} view {view}: max 8-bit difference to PyTorch CPU {d.max()}, differing pixels {(d > 0).mean():.4f}")
                assert d.max() <= tol, (method, view, d.max())
    assert isinstance(suite.renderers[1], nb.B200Renderer) and suite.renderers[1].last_render_time > 0
