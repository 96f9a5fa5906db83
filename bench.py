#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric: Mrays/s at 800x600, 128 samples/ray.

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --steps K --warmup W    # the UNMODIFIED reference (PyTorchCPURenderer) on the host cores
    python bench.py --workload train|hierarchical ...        # configs[3] / configs[4] with their own metric names

A step = one view of the 40-view synthetic orbit (reference benchmark_suite.py:132-149) rendered
through the fused kernel: 480,000 rays x 128 samples = 61.44 M network queries = 64.865 TFLOP
(1,055,744 FLOP per sample, unpadded).  At N > 1 every view is cut into N
row bands, one per GPU (strong scaling, no data-path collective); value = all rays of all ranks /
max-over-ranks device time.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE line (the JSON): everything else any library prints to fd 1 (NCCL's version
# banner, torchrun notices) is sent to stderr; emit() writes the result to the real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

W, H, S = 800, 600, 128
N_VIEWS = 40
FLOP_PER_SAMPLE = 1_055_744
METRIC = "Mrays/s at 800x600, 128 samples/ray"
L2_FLUSH_BYTES = 256 << 20          # > 126 MB L2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


def lego_weights():
    import numpy as np
    import torch
    z = np.load(os.path.join(ROOT, "tests", "golden", "ckpt_lego_stuffed_fp16.npz"))
    return {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (NVML, 100 ms period)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                     nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:  # noqa: BLE001 -- clocks are evidence, not a dependency
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arms are meant to use every host core, so the
    thread count is set explicitly (the round-1 scaling record's reference arm ran single-threaded at N > 1)."""
    import torch
    n = host_cores()
    torch.set_num_threads(n)
    return torch.get_num_threads()


class ReferenceCPU:
    """The reference's own CPU implementation of the path, unmodified: ``PyTorchCPURenderer`` imported from the
    vendored copy (oracle/_ref, tools/vendor_reference.sh) -- ``kind`` "reference".  When that copy is absent (a box
    that never received it) the oracle port of the same algorithm stands in -- ``kind`` "port"."""

    def __init__(self):
        import torch
        from oracle import refload
        self.cores = use_all_host_threads()
        self.weights = lego_weights()
        self.kind, self.renderer = "port", None
        if refload.reference_root() is not None:
            try:
                import tempfile
                refload.import_reference()
                from src.benchmark.pytorch_renderers import PyTorchCPURenderer
                self._tmp = tempfile.TemporaryDirectory()
                path = os.path.join(self._tmp.name, "lego_stuffed.pth")
                torch.save({"coarse_model": self.weights, "fine_model": self.weights}, path)
                r = PyTorchCPURenderer()
                r.setup(path)
                self.renderer, self.kind = r, "reference"
            except Exception as e:  # noqa: BLE001 -- fall back to the port, say why
                print(f"reference import failed ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)

    def rays(self, pose, width=W, height=H):
        import torch
        with torch.no_grad():
            if self.renderer is not None:
                ro, rd = self.renderer.generate_rays(pose, width, height)
            else:
                from oracle import nerf_oracle as O
                ro, rd = O.camera_rays(pose, width, height)
        return ro.reshape(-1, 3), rd.reshape(-1, 3)

    def render_first_rays(self, pose, n, samples=S):
        """The first ``n`` rays of render_image(pose, (W, H), samples): the 512-ray chunks render_image itself forms
        (pytorch_renderers.py:137-149), each through the reference's _render_ray_chunk."""
        import torch
        ro, rd = self.rays(pose)
        with torch.no_grad():
            for s0 in range(0, n, 512):
                if self.renderer is not None:
                    self.renderer._render_ray_chunk(ro[s0:min(s0 + 512, n)], rd[s0:min(s0 + 512, n)], samples)
                else:
                    from oracle import nerf_oracle as O
                    O.render_rays(self.weights, ro[s0:min(s0 + 512, n)], rd[s0:min(s0 + 512, n)], samples)

    def chunk_seconds(self, pose):
        self.render_first_rays(pose, 512)                                 # warm the thread pool / allocator
        t0 = time.perf_counter()
        self.render_first_rays(pose, 512)
        return max(time.perf_counter() - t0, 1e-3)


def bench_pose(i):
    """View i of the 40-view orbit as a fp32 4x4 (the product helper; equal to the reference's generate_test_poses,
    tests/test_orbit_pose.py)."""
    from nerf_dbr_b200.host.synthetic import orbit_pose
    return orbit_pose(i % N_VIEWS, N_VIEWS)


def cpu_baseline_sample(budget_s: float):
    """cpu_baseline of the headline line: a bounded band of view 0 of the same workload, all host threads."""
    ref = ReferenceCPU()
    pose = bench_pose(0)
    chunks = max(1, min(int(budget_s / ref.chunk_seconds(pose)), W * H // 512))
    n = chunks * 512
    t0 = time.perf_counter()
    ref.render_first_rays(pose, n)
    dt = time.perf_counter() - t0
    return {"value": n / dt / 1e6, "unit": "Mrays/s", "cores": ref.cores, "kind": ref.kind,
            "sample": f"first {n} rays of view 0 at 800x600x128 through PyTorchCPURenderer._render_ray_chunk, 512-ray chunks ({dt:.1f} s)"}


def cpu_optimized_baseline():
    """north_star: "the reference's PyTorch CPU and CPU_Optimized paths are timed on the same box's host cores".
    CPUOptimizedRenderer.render_image has no chunking (cpu_optimized_renderer.py:187-223): every [R*S, 256] activation
    of both networks is alive at once, ~9 GB at 200x150x32 and ~600 GB at 800x600x128 -- so it is timed at configs[0]
    (200x150, 32 samples per ray) only, next to PyTorchCPURenderer at the same size for a like-for-like ratio."""
    import tempfile
    import torch
    from oracle import refload
    if refload.reference_root() is None:
        return {"unavailable": "oracle/_ref not present on this box (tools/vendor_reference.sh)"}
    cores = use_all_host_threads()
    try:
        refload.import_reference()
        from src.benchmark.cpu_optimized_renderer import CPUOptimizedRenderer
        from src.benchmark.pytorch_renderers import PyTorchCPURenderer
        w = lego_weights()
        out = {"unit": "Mrays/s", "cores": cores, "kind": "reference", "workload": "200x150, 32 samples/ray, view 0 (configs[0])"}
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "lego_stuffed.pth")
            torch.save({"coarse_model": w, "fine_model": w}, path)
            for key, cls in (("value", CPUOptimizedRenderer), ("pytorch_cpu_same_size", PyTorchCPURenderer)):
                r = cls()
                r.setup(path)
                with torch.no_grad():
                    t0 = time.perf_counter()
                    r.render_image(bench_pose(0), (200, 150), 32)
                    out[key] = 200 * 150 / (time.perf_counter() - t0) / 1e6
        out["sample"] = "one full 200x150x32 frame each (CPUOptimizedRenderer.render_image, PyTorchCPURenderer.render_image), no warm-up"
        return out
    except Exception as e:  # noqa: BLE001 -- a baseline must not take the headline down
        return {"unavailable": f"{type(e).__name__}: {e}"}


def gpu_eager_baseline(dev):
    """The reference's only GPU path on the same B200: PyTorchCUDARenderer (pytorch_renderers.py:173-246) -- eager torch
    ops, cuBLAS SGEMM, 4096-ray chunks with a .cpu() per chunk -- at the headline size, TF32 off (torch's default, what
    the reference runs) and on.  Library kernels: timed as a baseline only, never on this repo's path."""
    import tempfile
    import torch
    from oracle import refload
    if refload.reference_root() is None:
        return {"unavailable": "oracle/_ref not present on this box (tools/vendor_reference.sh)"}
    try:
        refload.import_reference()
        from src.benchmark.pytorch_renderers import PyTorchCUDARenderer
        w = lego_weights()
        out = {"unit": "Mrays/s", "kind": "reference", "renderer": "PyTorchCUDARenderer (eager torch, cuBLAS), chunk 4096",
               "workload": "800x600x128, orbit views 0-2"}
        keep = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        with tempfile.TemporaryDirectory() as tmp, torch.cuda.device(dev):
            path = os.path.join(tmp, "lego_stuffed.pth")
            torch.save({"coarse_model": w, "fine_model": w}, path)
            r = PyTorchCUDARenderer()
            r.setup(path)
            for key, tf32 in (("value", False), ("value_tf32", True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                with torch.no_grad():
                    r.render_image(bench_pose(0), (W, H), S)             # warm-up (cuBLAS handles, allocator)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for i in (1, 2):
                        r.render_image(bench_pose(i), (W, H), S)
                    torch.cuda.synchronize()
                    out[key] = 2 * W * H / (time.perf_counter() - t0) / 1e6
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = keep
        torch.cuda.empty_cache()
        out["sample"] = "2 full frames after 1 warm-up per setting, wall clock around render_image (host image out, as the reference returns it)"
        return out
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (unmodified PyTorchCPURenderer from the
    vendored sources; the oracle port only if they are missing), all host threads whatever torchrun exported, a
    bounded band of every view per step.  Rank 0 alone runs; other ranks exit 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = ReferenceCPU()
    steps, warm = args.steps, args.warmup
    per_step_budget = min(3.0, 150.0 / max(1, steps + warm))
    chunks = max(1, int(per_step_budget / ref.chunk_seconds(bench_pose(0))))
    n = min(chunks * 512, W * H)
    times = []
    for i in range(steps + warm):
        t0 = time.perf_counter()
        ref.render_first_rays(bench_pose(i), n)                          # includes generate_rays of the whole view, as render_image does
        if i >= warm:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = n * steps / total / 1e6
    sample = (f"first {n} rays ({n / W:.1f} rows) of each 800x600x128 view: render_image's own 512-ray chunks through "
              f"PyTorchCPURenderer._render_ray_chunk, fine network" if ref.kind == "reference" else
              f"first {n} rays of each 800x600x128 view, oracle port of PyTorchCPURenderer (vendored reference sources absent)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * total / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "800x600x128 render, 40-view synthetic orbit (reference benchmark_suite.py:132-149)",
                       "weights": "lego_stuffed_fp16 fixture", "sample_rays_per_step": n,
                       "host_threads": ref.cores, "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": ref.cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def train_roofline(precision, n_samples, flop, ms):
    """Step-level roofline of the training workload, per GPU (n_samples, flop: this rank's share).  BF16 mode is HBM-bound by construction: the forward writes the
    bf16 operand blocks (4544 B/sample + 88 B of masks and head rows), the dgrad chain writes dpre (4352 B, reads
    88 B), the eleven wgrad launches read every operand once (9600 B), the ray kernel 36 B -- 18.7 KB per sample
    (DESIGN.md 4.3).  That traffic is what the step is measured against; the part streams writes at 6.2-6.3 TB/s and reads at
    6.9 TB/s (tools/probe/hbm_store_probe.cu), and what bounds the forward and dgrad kernels below that is the serial chain of
    their epilogue warps (accumulator -> pack -> stage -> bulk store -> mask per layer half), not HBM."""
    tfl = flop / (ms * 1e-3) / 1e12
    if precision != "bf16":
        return {"bound": "tensor", "achieved": tfl, "peak": peaks()["bf16_tflops"], "unit": "TFLOP/s",
                "frac": tfl / peaks()["bf16_tflops"], "traffic": None, "note": "FP32 CUDA-core kernels (gradient parity mode)"}
    bytes_per_sample = 4544 + 88 + 4352 + 88 + 9600 + 36
    gbs = bytes_per_sample * n_samples / (ms * 1e-3) / 1e9
    traffic = None                                # measured DRAM bytes per 4096-ray step (ncu), when the batch is configs[3]
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and n_samples == 4096 * 192:
        with open(tpath) as fh:
            traffic = json.load(fh).get("train_bf16_step_dram_bytes")
    return {"bound": "hbm", "achieved": gbs, "peak": peaks()["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks()["hbm_gbs"],
            "traffic": traffic, "algorithmic_bytes_per_sample": bytes_per_sample, "tflops": tfl,
            "tflops_frac_of_bf16_peak": tfl / peaks()["bf16_tflops"],
            "note": ("whole step: forward, dgrad chain and wgrad on tcgen05 (bf16 operands, fp32 accumulate); activations and "
                     "pre-activation gradients visit HBM once as bf16 (TMA bulk stores); the forward and dgrad kernels are bound by "
                     "their epilogue warps' per-half chain, wgrad by HBM reads (5.9 TB/s)")}


def measure_train(dev, rank, world, weak, steps, warmup, precision, fused=True, transport="auto"):
    """BASELINE.json configs[3]: 4096-ray batch (strong: in total; weak: per GPU), coarse 64 jittered + fine 128 uniform
    samples, forward + backward of mse(coarse)+mse(fine), gradient exchange, clip, Adam, re-pack -- one iteration of
    NeRFTrainer.train_step (trainer.py:83-138) per step.  ``fused``: TrainEngine (one CUDA graph of this library's
    kernels, gradients summed over NVLink peer memory); else the unfused sequence (B200TrainStep + NCCL all-reduce +
    torch's fused Adam).  A fresh batch is copied into the static input buffers inside the timed region every step.
    Requires torch.distributed to be initialised when world > 1.  Returns the result dict (meaningful on rank 0)."""
    import torch
    import torch.distributed as dist
    from nerf_dbr_b200.host import lib as L
    from nerf_dbr_b200.host import ops
    from nerf_dbr_b200.host.engine import TrainEngine
    from nerf_dbr_b200.host.parallel import ray_shard
    from nerf_dbr_b200.host.synthetic import seeded_models
    from nerf_dbr_b200.host.trainer import B200TrainStep

    n_c, n_f = 64, 128
    n_rays = 4096 * (world if weak else 1)          # global batch: configs[3] (strong), or configs[3] per GPU (weak)
    coarse, fine = seeded_models(5, 30.0, dev)      # seeded default init, density heads x30 (semi-opaque volume)
    mode = L.BF16 if precision == "bf16" else L.FP32
    pose = torch.eye(4); pose[2, 3] = 4.0
    ro, rd = (t.cpu() for t in ops.generate_rays(pose, 200, 150, device=dev))
    g = torch.Generator().manual_seed(0)
    image = torch.rand(150, 200, 3, generator=g)
    first, count = ray_shard(rank, world, n_rays)
    batches = []
    for i in range(min(steps + warmup, 24)):        # a ring of distinct batches, resident in HBM
        sel = (torch.randperm(200 * 150, generator=g)[:n_rays] if n_rays <= 200 * 150 else
               torch.randint(0, 200 * 150, (n_rays,), generator=g))[first:first + count]
        batches.append((ro.reshape(-1, 3)[sel].to(dev), rd.reshape(-1, 3)[sel].to(dev), image.reshape(-1, 3)[sel].to(dev),
                        torch.rand(count, n_c, generator=g).to(dev)))
    if fused:
        eng = TrainEngine(coarse, fine, count, n_c, n_f, mode=mode, lr=5e-4, gamma=0.1 ** (1 / 250000), max_norm=1.0,
                          n_rays_global=n_rays, transport=transport)

        def one(i):
            eng.step(*batches[i % len(batches)])
        path = f"TrainEngine: CUDA graph, gradient exchange '{eng.transport}', fused clip+Adam+decay"
    else:
        step = B200TrainStep(coarse, fine, n_c, n_f, mode=mode)
        opt = torch.optim.Adam(step.parameters(), lr=5e-4, fused=True)

        def one(i):
            b = batches[i % len(batches)]
            step(b[0], b[1], b[2], t_rand=b[3], n_rays_global=n_rays)
            torch.nn.utils.clip_grad_norm_(step.parameters(), 1.0)
            opt.step()
        path = "B200TrainStep + NCCL all-reduce + torch clip_grad_norm_ + torch fused Adam (eager launches)"

    for i in range(warmup):
        one(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        one(warmup + k)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    launches = int(ops.launch_count() - n0)
    loss = eng.loss() if fused else None
    graph = bool(eng.graph) if fused else False
    if graph:                                           # replays do not pass through the library's host entry points:
        launches = eng.launches_per_step * steps        # kernel nodes per replay (counted on the eager first step) x steps
    flop = 3_095_808 * n_rays * (n_c + n_f)
    out = {"metric": "training rays/s, 4096-ray batch, 64 coarse + 128 fine samples, fwd+bwd+exchange+clip+Adam",
           "value": n_rays / (ms * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
           "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
           "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
           "config": {"workload": "BASELINE.json configs[3]: 4096-ray batch fwd+bwd MSE, data-parallel gradient exchange, Adam",
                      "global_rays": n_rays, "rays_per_gpu": count, "path": path, "cuda_graph": graph, "loss_last": loss,
                      "inputs": "fresh batch copied device-to-device into the static buffers inside the timed region, every step"},
           "gpu_launches": launches,
           "kernels_per_step": (eng.launches_per_step if fused else launches // max(1, steps)),
           "roofline": train_roofline(precision, n_rays * (n_c + n_f) // world, flop / world, ms)}      # per GPU
    del batches
    torch.cuda.empty_cache()
    return out


def run_train(args):
    """--workload train: configs[3] with its own metric name (extra evidence; the default run carries the same numbers
    under extras.train)."""
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = measure_train(dev, rank, world, args.weak, args.steps, args.warmup, args.precision, fused=not args.unfused,
                        transport=args.transport)
    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def run_hierarchical(args):
    """--workload hierarchical: BASELINE.json configs[4] -- 1600x1200, 128 coarse + 128 importance samples (inverse-CDF
    fine pass on the 256-sample sorted union), a 64-view orbit, every view cut into row bands over the GPUs.  Three
    launches per view and rank: fused coarse render (compositing weights out), fused sampling (coarse depths +
    inverse CDF + sorted union; uniforms drawn in the kernel), fused fine render at explicit depths.  A step = one view.
    The reference's importance_sample is dead code (src/utils/rendering.py:54-100 raises at :89), so there is no
    reference arm for this workload; FLOPs: 1,055,744 per network query, 128 coarse + 256 fine queries per ray."""
    import torch
    import torch.distributed as dist
    from nerf_dbr_b200.host import lib as L
    from nerf_dbr_b200.host import ops
    from nerf_dbr_b200.host.parallel import row_band
    from nerf_dbr_b200.host.synthetic import orbit_pose
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w, h, n_c, n_i, views = 1600, 1200, 128, 128, 64
    mode = L.BF16 if args.precision == "bf16" else L.FP32
    weights = lego_weights()
    net = ops.pack_weights({k: v.to(dev) for k, v in weights.items()}, dev)
    row0, n_rows = row_band(rank, world, h)
    poses = [orbit_pose(i % views, views) for i in range(args.steps + args.warmup)]
    marks = []

    def view(i, timed):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if timed else None
        if timed:
            ev[0].record()
        ro, rd = ops.generate_rays(poses[i], w, h, row0=row0, n_rows=n_rows, device=dev)
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        if timed:
            ev[1].record()
        rgb_c, _, wts = ops.render_rays(net, ro, rd, n_c, mode, want_weights=True)
        if timed:
            ev[2].record()
        z_all = ops.hierarchical_samples(wts, n_i, seed=i)
        if timed:
            ev[3].record()
        rgb_f, dep_f = ops.render_rays(net, ro, rd, n_c + n_i, mode, z_vals=z_all)
        if timed:
            ev[4].record()
            marks.append(ev)
        return rgb_f

    for i in range(args.warmup):
        view(i, False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n0 = ops.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    for k in range(args.steps):
        out = view(args.warmup + k, True)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = ops.launch_count() - n0
    stage = [sum(ev[j].elapsed_time(ev[j + 1]) for ev in marks) / args.steps for j in range(4)]
    total = sum(ev[0].elapsed_time(ev[4]) for ev in marks)
    t = torch.tensor([total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    if rank == 0:
        pk = peaks()
        flop_rank = FLOP_PER_SAMPLE * n_rows * w * (n_c + n_c + n_i)
        mlp_ms = stage[1] + stage[3]
        tfl = flop_rank / (mlp_ms * 1e-3) / 1e12
        samp_bytes = n_rows * w * 4 * (n_c + n_c + n_i)            # weights in, union out (uniforms drawn in the kernel)
        emit({"metric": "Mrays/s at 1600x1200, 128 coarse + 128 importance samples/ray (hierarchical)", "value": w * h / (ms * 1e-3) / 1e6,
              "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if mode == L.BF16 else "f32",
              "data": "synthetic",
              "config": {"workload": "BASELINE.json configs[4]: 1600x1200, 128 coarse + 128 inverse-CDF samples, fine pass on the sorted "
                                     "union of 256, 64-view orbit, row bands over the GPUs", "rows_per_gpu": n_rows,
                         "ms_rank0": {"generate_rays": stage[0], "coarse_render": stage[1], "sampling": stage[2], "fine_render": stage[3]},
                         "launches_per_view": launches / args.steps, "finite": bool(torch.isfinite(out).all()),
                         "timing": "CUDA events per view on the launching stream, summed; max over ranks",
                         "l2": "per-view working set (weights + union: 3 GB at N = 1) exceeds the 126 MB L2"},
              "gpu_launches": int(launches), "clocks": clocks,
              "roofline": {"bound": "tensor", "achieved": tfl, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": tfl / pk["bf16_tflops"],
                           "traffic": None, "kernel": "fused_render_kernel<SRC_RAYS> (coarse + fine passes)",
                           "algorithmic_flop_per_view_rank0": flop_rank,
                           "sampling_kernel": {"bound": "hbm", "limited_by": "instruction issue + shared-memory probes of the sort / rank searches (ncu: issue slots 64 %, "
                                                             "L1 81 %, DRAM 12 % busy; profiles/r2_hbm_kernels_full_ncu.txt)",
                                               "algorithmic_bytes": samp_bytes, "ms": stage[2],
                                               "achieved_gbs": samp_bytes / (stage[2] * 1e-3) / 1e9, "peak_gbs": pk["hbm_gbs"],
                                               "frac": samp_bytes / (stage[2] * 1e-3) / 1e9 / pk["hbm_gbs"]}}})
    if world > 1:
        dist.destroy_process_group()


def measure_hbm_kernels(dev):
    """"standalone raygen/sampling and standalone compositing -> HBM bandwidth".  The staged API's kernels at the
    headline shape (800x600x128), inputs resident, CUDA events around 10 launches after 3 warm-ups; outputs (0.7-2 GB each)
    are larger than the 126 MB L2.  Algorithmic bytes per launch as in SURVEY 8d; peak = the measured copy bandwidth."""
    import torch
    from nerf_dbr_b200.host import ops
    pk = peaks()
    Wd, Hd, Sd = 800, 600, 128
    R = Wd * Hd
    pose = torch.eye(4)
    pose[2, 3] = 4.0

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    out = {}

    def add(name, ms, nbytes, kernel):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = {"kernel": kernel, "ms": ms, "algorithmic_bytes": nbytes, "achieved_gbs": gbs, "frac": gbs / pk["hbm_gbs"]}

    with torch.cuda.device(dev):
        ro, rd = ops.generate_rays(pose, Wd, Hd, device=dev)
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        add("sample_points_on_rays", timed(lambda: ops.sample_points(ro, rd, Sd)), 16 * R * Sd + 24 * R, "sample_points_rays_kernel")
        _, z = ops.sample_points(ro, rd, Sd)
        sigma = torch.rand(R, Sd, device=dev)
        col = torch.rand(R, Sd, 3, device=dev)
        add("execute_volume_rendering", timed(lambda: ops.composite(sigma, col, z, rd)), 20 * R * Sd + 12 * R + 16 * R, "composite4_kernel")
        x = torch.rand(1 << 23, 3, device=dev)
        add("positional_encoding_L10", timed(lambda: ops.positional_encoding(x, 10)), (12 + 4 * 63) * (1 << 23), "encode_rows_kernel<10>")
        wts = torch.rand(R, Sd, device=dev)
        u = torch.rand(R, Sd, device=dev)
        add("importance_sample", timed(lambda: ops.importance_sample(ro, rd, z, wts, u)), (12 + 24) * R * Sd + 24 * R, "importance_warp_kernel")
    return {"peak_gbs": pk["hbm_gbs"], "peak_source": pk["source"], "shape": "800x600x128 (encoding: 2^23 points)",
            "timing": "CUDA events around 10 launches through the tensor-level wrapper (output allocation included), after 3 warm-ups",
            "kernels": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--weak", action="store_true", help="train workload: 4096 rays per GPU instead of 4096 in total")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "fp8"],
                    help="render workload; bf16 = the headline (north_star's tensor-core mode); fp8 = the lossy e4m3 mode, extra evidence")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="render", choices=["render", "train", "hierarchical"],
                    help="render = the headline metric (default); train = BASELINE.json configs[3]; hierarchical = configs[4]")
    ap.add_argument("--unfused", action="store_true", help="train workload: B200TrainStep + NCCL + torch Adam instead of TrainEngine")
    ap.add_argument("--transport", default="auto", choices=["auto", "p2p", "multimem", "nccl"],
                    help="train workload: gradient exchange of TrainEngine")
    ap.add_argument("--no-extras", action="store_true", help="render workload: skip extras.train and the extra baselines")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    if args.workload == "train":
        return run_train(args)
    if args.workload == "hierarchical":
        return run_hierarchical(args)

    import torch
    import torch.distributed as dist
    from nerf_dbr_b200.host import ops, lib as L
    from nerf_dbr_b200.host.synthetic import orbit_pose

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mode = {"bf16": L.BF16, "fp32": L.FP32, "fp8": L.FP8}[args.precision]

    weights = lego_weights()
    net = ops.pack_weights({k: v.to(dev) for k, v in weights.items()}, dev)
    if mode == L.FP8:                                   # quantised buffer, scales calibrated on points of the orbit
        net = ops.pack_weights_fp8({k: v.to(dev) for k, v in weights.items()}, dev, packed=net)
    # row band of this rank
    from nerf_dbr_b200.host.parallel import row_band
    row0, n_rows = row_band(rank, world, H)
    rgb = torch.empty(n_rows, W, 3, device=dev)
    depth = torch.empty(n_rows, W, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    poses = [orbit_pose(i % N_VIEWS, N_VIEWS) for i in range(args.steps + args.warmup)]

    def step(i):
        ops.render_image(net, poses[i], W, H, S, mode, row0=row0, n_rows=n_rows, out_rgb=rgb, out_depth=depth)

    # ---- device-timed: K steps, CUDA events on the launching stream, L2 flushed between steps ----
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = ops.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()                                   # evicts L2 between timed steps (not timed)
        ev[k][0].record()
        step(args.warmup + k)
        ev[k][1].record()
    torch.cuda.synchronize()
    launches = ops.launch_count() - n0
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = W * H * args.steps / (total_ms * 1e-3) / 1e6

    # ---- end to end through the public renderer API: checkpoint file -> setup(); host poses in, host images out ----
    import tempfile
    import nerf_dbr_b200 as nb
    r = nb.B200Renderer(args.precision, device_index=local)
    with tempfile.TemporaryDirectory() as tmp:
        ck = os.path.join(tmp, f"lego_stuffed_rank{rank}.pth")
        torch.save({"coarse_model": weights, "fine_model": weights}, ck)
        r.setup(ck)                                      # the reference's own entry: SharedNeRFModel + weight packing
    host_poses = [p.pin_memory() for p in poses]
    checksum = 0.0
    for rgb_h, _ in r.render_views(host_poses[:min(3, args.warmup)], (W, H), S, row0, n_rows):
        checksum += float(rgb_h[0, 0, 0])
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for rgb_h, depth_h in r.render_views(host_poses[args.warmup:args.warmup + args.steps], (W, H), S, row0, n_rows):
        checksum += float(rgb_h[0, 0, 0]) + float(depth_h[-1, -1])      # the caller touches every image it receives
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = W * H * args.steps / float(t.item()) / 1e6

    # ---- thermally settled figure: 120 views back to back, no flush (1 s bursts flatter a power-capped part) ----
    n_settle = 120
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(n_settle):
        ops.render_image(net, poses[i % len(poses)], W, H, S, mode, row0=row0, n_rows=n_rows, out_rgb=rgb, out_depth=depth)
    s1.record()
    torch.cuda.synchronize()
    t = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    settled_value = W * H * n_settle / (float(t.item()) * 1e-3) / 1e6

    # ---- configs[3] in the same record: training step, strong and weak data-parallel scaling (all ranks take part) ----
    extras = {}
    if not args.no_extras:
        del flush
        torch.cuda.empty_cache()
        try:
            tr = {}
            for tag, weak in (("strong", False), ("weak", True)):
                if world == 1 and weak:
                    continue                             # identical to strong at one GPU
                m = measure_train(dev, rank, world, weak, 40, 8, "bf16")
                tr[tag] = {k: m[k] for k in ("value", "unit", "ms_per_step", "gpu_launches", "kernels_per_step", "roofline", "config")}
            extras["train"] = tr
        except Exception as e:  # noqa: BLE001 -- extras must not take the headline down
            extras["train"] = {"unavailable": f"{type(e).__name__}: {e}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    if mode == L.FP8:                                    # no measured fp8 peak on file: twice the measured bf16 figure (the nominal ratio)
        pk = dict(pk, bf16_tflops=2 * pk["bf16_tflops"], source=pk["source"] + " x 2 for e4m3 (nominal fp8 : bf16 ratio)",
                  bf16_tflops_sustained=2 * pk["bf16_tflops_sustained"] if pk["bf16_tflops_sustained"] else None)
    ms_per_launch = dev_ms / max(1, launches)            # rank 0's kernel; one launch per step
    flops_per_launch = FLOP_PER_SAMPLE * n_rows * W * S
    achieved = flops_per_launch / (ms_per_launch * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get("fused_render_kernel_dram_bytes_per_launch") if mode != L.FP8 else None
    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": {L.BF16: "bf16", L.FP32: "f32", L.FP8: "fp8 e4m3 (bf16 encoded-position inputs, fp32 accumulate)"}[mode], "data": "synthetic",
        "config": {"workload": "800x600x128 render, 40-view synthetic orbit (BASELINE.json configs[2]), "
                               "fine network, uniform samples, image row bands sharded across GPUs",
                   "rays_per_step": W * H, "samples_per_ray": S, "msamples_per_s": value * S,
                   "weights": "lego_stuffed_fp16 fixture (tests/golden)", "l2": "flushed between timed steps (256 MiB memset)",
                   "timing": "CUDA events per step on the launching stream, summed; max over ranks",
                   "settled_mrays_per_s_120_views": settled_value,
                   "settled_note": "120 views back to back, one pair of CUDA events, no L2 flush: the power-capped steady state"},
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 64,
                "d2h_bytes_per_step": n_rows * W * 16,
                "path": "B200Renderer.setup(checkpoint.pth); B200Renderer.render_views(host poses) -> pinned host rgb+depth per "
                        "view (copy of view k overlaps the render of view k+1), wall clock over all views incl. the last copy"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": achieved / pk["bf16_tflops"], "traffic": traffic,
                     "kernel": "fused_render_fp8_kernel<SRC_POSE>" if mode == L.FP8 else "fused_render_kernel<SRC_POSE>",
                     "algorithmic_flop_per_launch": flops_per_launch,
                     "peak_source": pk["source"] + ", burst cuBLAS bf16",
                     "frac_of_sustained": (achieved / pk["bf16_tflops_sustained"]) if pk["bf16_tflops_sustained"] else None},
    }
    if not args.no_extras:
        try:
            extras["hbm_kernels"] = measure_hbm_kernels(dev)
        except Exception as e:  # noqa: BLE001 -- extras must not take the headline down
            extras["hbm_kernels"] = {"unavailable": f"{type(e).__name__}: {e}"}
    if extras:
        line["extras"] = extras
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(12.0)
        if not args.no_extras:
            line["gpu_eager_baseline"] = gpu_eager_baseline(dev)
            line["cpu_baseline_optimized"] = cpu_optimized_baseline()
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
