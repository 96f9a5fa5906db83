"""CPU: the product's synthetic camera orbit (``nerf_dbr_b200.host.synthetic.orbit_pose``, what bench.py and the
full-size tests render) equals the oracle's ``benchmark_pose``, which tests/golden/make_golden*.py assert equal to the
reference's ``UnifiedBenchmarkSuite.generate_test_poses`` (src/benchmark/benchmark_suite.py:132-149) bit for bit."""
import numpy as np
import torch

from conftest import load_npz
from oracle import nerf_oracle as O


def test_orbit_pose_equals_oracle_benchmark_pose():
    from nerf_dbr_b200.host.synthetic import orbit_pose
    for n in (1, 2, 3, 7, 40, 64):
        for i in range(n):
            a, b = orbit_pose(i, n), O.benchmark_pose(i, n)
            assert a.dtype == torch.float32 and torch.equal(a, b), (i, n)


def test_orbit_pose_equals_reference_poses_in_the_config_goldens():
    from nerf_dbr_b200.host.synthetic import orbit_pose
    g = load_npz("golden_configs.npz")          # poses written by the reference's generate_test_poses(40)
    assert np.array_equal(orbit_pose(5, 40).numpy(), g["pose_c2"])
    assert np.array_equal(orbit_pose(7, 40).numpy(), g["pose_c3"])
