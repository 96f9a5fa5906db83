"""CPU: the three numerical arguments the rewritten standalone kernels (csrc/geometry.cu, csrc/data.cu) rest on, checked
without a GPU -- each is the host-side statement of what the kernel computes, so a change to the kernel's recipe has a
test to break here before it breaks parity on the device.

  * composite_white4_kernel: c / 255.0 from a product and two fused corrections is the correctly rounded quotient.
  * warp_build_cdf: the double-precision running sum of the quotients is exact (so any order gives the serial result)
    whenever every quotient is >= 2^-28, which compositing weights always satisfy.
  * encode_rows_kernel: sin / cos from the 64-bit fixed-point phase of fl(pi_f x) and two short polynomials stay
    within the 3e-7 gate of the reference's torch.sin / torch.cos of fl(fl(2^k pi) x)."""
from fractions import Fraction as F

import numpy as np
import torch


def test_div255_is_correctly_rounded_for_every_byte():
    inv = 1.0 / 255.0
    for c in range(256):
        q0 = float(F(c) * F(inv))                       # __dmul_rn(c, 1/255)
        r = float(F(c) - F(q0) * F(255))                # __fma_rn(-q0, 255, c)
        q = float(F(q0) + F(r) * F(inv))                # __fma_rn(r, 1/255, q0)
        assert q == float(F(c, 255)) == c / 255.0, c


def test_cdf_running_sum_is_exact_when_quotients_are_at_least_2_pow_minus_28():
    g = torch.Generator().manual_seed(0)
    for S in (32, 64, 128, 256):
        w = torch.rand(4000, S, generator=g) ** 8                 # compositing-weight-like rows in [0, 1]
        w[::5] = 0.0
        w[1::9, 3] = 1.0
        v = w + 1e-5
        q = (v / v.sum(-1, keepdim=True)).numpy()                  # fp32 quotients, as importance_kernel forms them
        assert q.min() >= 2.0 ** -28
        serial = np.cumsum(q.astype(np.float64), axis=-1)          # the reference's order (double accumulator)
        exact = np.array([[float(x) for x in np.cumsum([F(float(t)) for t in row])] for row in q[:50]])
        assert np.array_equal(serial[:50], exact)                  # no addition rounded
        # any other association gives the same doubles: blocked lane sums + scan over lanes, as the warp kernel does
        spl = S // 32
        blocks = q.astype(np.float64).reshape(-1, 32, spl)
        local = np.cumsum(blocks, axis=-1)
        lane_tot = local[..., -1]
        offs = np.concatenate([np.zeros_like(lane_tot[:, :1]), np.cumsum(lane_tot, axis=-1)[:, :-1]], axis=-1)
        # a Hillis-Steele scan adds in yet another order; pairwise tree here
        tree = lane_tot.copy()
        o = 1
        while o < 32:
            shifted = np.concatenate([np.zeros_like(tree[:, :o]), tree[:, :-o]], axis=-1)
            tree = tree + shifted
            o <<= 1
        assert np.array_equal(tree - lane_tot, offs)
        par = (offs[..., None] + local).reshape(-1, S)
        assert np.array_equal(par, serial)
        assert np.array_equal(par.astype(np.float32), torch.cumsum(torch.from_numpy(q), -1).numpy())   # torch's CPU cumsum


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def test_phase_shift_sincos_stays_inside_the_encoding_gate():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-4, 4, 200000), rng.normal(0, 300, 20000), rng.uniform(-16384, 16384, 20000),
                        [0.0, 1e-8, -1e-8, 0.5, 1.0, 2.0, 4.0, 16384.0]]).astype(np.float32)
    f32 = np.float32
    pi_f = f32(3.14159274101257324)
    a = (pi_f * x).astype(np.float32)
    turns = a.astype(np.float64) * 0.15915494309189535
    ph = np.rint((turns - np.rint(turns)) * 2.0 ** 63).astype(np.int64).astype(np.uint64)
    one = np.ones_like(x)
    for k in range(10):
        frac = ((ph << np.uint64(k + 1)) >> np.uint64(32)).astype(np.uint32)
        t = frac + np.uint32(0x20000000)
        q = t >> np.uint32(30)
        r = ((t & np.uint32(0x3FFFFFFF)).astype(np.int64) - 0x20000000).astype(np.float32) * f32(1.4629180792671596e-9)
        z = (r * r).astype(np.float32)
        ps = _fma(_fma(f32(-1.9515295891e-4) * one, z, f32(8.3321608736e-3) * one), z, f32(-1.6666654611e-1) * one)
        sr = _fma((ps * z).astype(np.float32), r, r)
        pc = _fma(_fma(f32(2.443315711809948e-5) * one, z, f32(-1.388731625493765e-3) * one), z, f32(4.166664568298827e-2) * one)
        cr = _fma((pc * z).astype(np.float32), z, _fma(z, f32(-0.5) * one, one))
        s = np.where(q == 0, sr, np.where(q == 1, cr, np.where(q == 2, -sr, -cr)))
        c = np.where(q == 0, cr, np.where(q == 1, -sr, np.where(q == 2, -cr, sr)))
        arg = torch.tensor(2.0 ** k) * torch.pi * torch.from_numpy(x)              # nerf.py:42-43: fp32 product
        assert np.array_equal(arg.numpy(), (f32(2.0 ** k) * a).astype(np.float32))  # = 2^k fl(pi_f x): the shift is exact
        assert np.abs(s - torch.sin(arg).numpy()).max() <= 2.5e-7, k
        assert np.abs(c - torch.cos(arg).numpy()).max() <= 2.5e-7, k
        assert np.abs(s - np.sin(arg.double().numpy())).max() <= 1.3e-7 and np.abs(c - np.cos(arg.double().numpy())).max() <= 1.3e-7, k


def _count_below(arr, v, log, strict):
    """count_below<LOG, STRICT> of csrc/geometry.cu, step for step."""
    pos = 0
    step = 1 << (log - 1)
    while step > 0:
        x = arr[pos + step - 1]
        if (x < v) if strict else (x <= v):
            pos += step
        step >>= 1
    x = arr[pos]
    if (x < v) if strict else (x <= v):
        pos += 1
    return pos


def test_fixed_probe_descent_equals_searchsorted():
    """The rank searches of the warp-per-ray sampling kernels: log2(n) + 1 probes without bounds over a sorted row of
    2^LOG entries (padded with +inf) return numpy's searchsorted for every query, ties and out-of-range values included."""
    rng = np.random.default_rng(5)
    for log in (5, 6, 7, 8):
        n = 1 << log
        for n_real in (n, n - 1, n // 2 + 3, 1):
            row = np.sort(rng.integers(0, 40, n_real).astype(np.float32))          # many ties
            padded = np.concatenate([row, np.full(n - n_real, np.inf, np.float32)])
            for v in np.concatenate([np.unique(row), np.unique(row) + 0.5, [-1.0, 1e9]]).astype(np.float32):
                assert _count_below(padded, v, log, False) == np.searchsorted(row, v, side="right")
                assert _count_below(padded, v, log, True) == np.searchsorted(row, v, side="left")


def test_register_bitonic_network_sorts():
    """warp_bitonic_sort<NPL>: element e = lane * NPL + i, partner e ^ j, ascending blocks where (e & k) == 0 -- the same
    (k, j) schedule and min/max rule, emulated on whole arrays, sorts every input incl. ties and +inf padding."""
    rng = np.random.default_rng(6)
    for npl in (1, 2, 4, 8):
        n = 32 * npl
        for trial in range(20):
            v = rng.integers(0, 50, n).astype(np.float32)
            v[rng.integers(0, n, 5)] = np.inf
            ref = np.sort(v)
            e = np.arange(n)
            k = 2
            while k <= n:
                j = k >> 1
                while j > 0:
                    partner = v[e ^ j]
                    take_min = ((e & j) == 0) == ((e & k) == 0)
                    v = np.where(take_min, np.minimum(v, partner), np.maximum(v, partner))
                    j >>= 1
                k <<= 1
            assert np.array_equal(v, ref)


def test_rank_merge_positions_are_a_permutation():
    """The rank merge of the sampling kernels: a_i -> i + #{b < a_i}, b_j -> j + #{a <= b_j} fills every slot of the union
    exactly once and yields sort(cat(a, b)), ties between and inside the lists included."""
    rng = np.random.default_rng(7)
    for na, nb in ((128, 128), (64, 100), (32, 7), (96, 1)):
        a = np.sort(rng.integers(0, 60, na).astype(np.float32))
        b = np.sort(rng.integers(0, 60, nb).astype(np.float32))
        out = np.full(na + nb, np.nan, np.float32)
        pa = np.arange(na) + np.searchsorted(b, a, side="left")
        pb = np.arange(nb) + np.searchsorted(a, b, side="right")
        assert len(set(pa) | set(pb)) == na + nb
        out[pa] = a
        out[pb] = b
        assert np.array_equal(out, np.sort(np.concatenate([a, b])))


def test_compositing_transmittance_association_is_invisible_in_fp32():
    """composite4_kernel multiplies a lane's four keep factors locally and scans the lane products across the warp; the
    reference's cumprod is a serial double product.  The two doubles differ in their last bits only, so the fp32
    transmittance differs at most by one rounding tie: <= 1 ulp, and in well under 1e-4 of the samples."""
    rng = np.random.default_rng(8)
    alpha = (rng.random((4000, 128)) ** 3).astype(np.float32)
    keep = ((np.float32(1.0) - alpha) + np.float32(1e-10)).astype(np.float32).astype(np.float64)
    serial = np.cumprod(keep, axis=-1)
    excl_serial = np.concatenate([np.ones_like(serial[:, :1]), serial[:, :-1]], axis=-1).astype(np.float32)
    blocks = keep.reshape(-1, 32, 4)
    local_excl = np.concatenate([np.ones_like(blocks[..., :1]), np.cumprod(blocks, axis=-1)[..., :-1]], axis=-1)
    lane_tot = np.prod(blocks, axis=-1)
    incl = lane_tot.copy()                                      # Hillis-Steele inclusive scan, as warp_incl_prod
    o = 1
    while o < 32:
        shifted = np.concatenate([np.ones_like(incl[:, :o]), incl[:, :-o]], axis=-1)
        incl = incl * shifted
        o <<= 1
    lead = np.concatenate([np.ones_like(incl[:, :1]), incl[:, :-1]], axis=-1)
    par = (lead[..., None] * local_excl).reshape(-1, 128).astype(np.float32)
    diff = par != excl_serial
    assert diff.mean() < 1e-4
    ulp = np.abs(par.view(np.int32).astype(np.int64) - excl_serial.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1
