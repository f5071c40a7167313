"""Two-dimensional phase unwrapping (the skimage.restoration.unwrap_phase step of the reference's energy,
tensor_tools.py:531): oracle (oracle/unwrap_herraez.c) against known answers, the kernel sources in emulation and
the CUDA library against the oracle.

scikit-image is not available, so there is no golden vector of the reference for this step (PARITY UNPINNED, see the
oracle's header); what is checked instead:
  * exact recovery (up to one global multiple of 2 pi) of smooth fields wrapped into (-pi, pi], numpy.unwrap on
    fields varying along one axis, results that differ from the input by exact integer multiples of 2 pi;
  * the product (offset-carrying union-find, device-side keys and sort) and the oracle (linked pixel lists, qsort)
    produce the SAME integer field — bit for bit when both start from the same wrapped angles, vortices and noise
    included;
  * the energy with unwrap_mode 2 equals the oracle's eng_expect restatement fed with the oracle unwrapper, on the
    golden runs where that number is insensitive to rounding noise.
"""
import os

import numpy as np
import pytest

from oracle import spinor_oracle as orc
from oracle.unwrap_oracle import unwrap_phase as oracle_unwrap

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
TWO_PI = 6.283185307179586
# golden runs whose Herraez energy moves by < 1e-13 under 3e-15 relative perturbations of psi_k (the other runs have
# (nearly) real negative psi somewhere: the global 2 pi offset of the unwrapped field, and with it the phase gradient
# across the mask edge, depends on rounding noise — in the reference as well)
ENERGY_ROBUST = [('cgrad_64', 0), ('cgrad_64', 1), ('ground_64', 0), ('nocoupl_64', 0)]


def wrap(a):
    return np.angle(np.exp(1j * a))


def fields(ny, nx, seed):
    """name -> complex field; smooth ramps, vortices and noise."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:ny, 0:nx].astype(np.float64)
    smooth = 0.31 * x + 0.17 * y + 3.0 * np.sin(x / 7.0) * np.cos(y / 5.0)
    amp = 1.0 + 0.3 * np.cos(x / 3.0)
    vort = (x - 0.31 * nx + 1j * (y - 0.37 * ny)) * (x - 0.62 * nx - 1j * (y - 0.71 * ny)) * np.exp(0.2j * x)
    noise = rng.normal(size=(ny, nx)) + 1j * rng.normal(size=(ny, nx))
    return {'smooth': amp * np.exp(1j * smooth), 'vortices': vort + 0.5 * noise, 'noise': noise,
            'flat': np.ones((ny, nx), dtype=complex) * np.exp(0.3j)}


def increments(unwrapped, wrapped):
    k = (unwrapped - wrapped) / TWO_PI
    assert np.abs(k - np.round(k)).max() < 1e-9
    return np.round(k).astype(np.int64)


# ----------------------------------------------------------------------------- oracle: known answers
def test_oracle_recovers_smooth_field():
    y, x = np.mgrid[0:48, 0:80].astype(np.float64)
    true = 0.21 * x + 0.13 * y + 3 * np.sin(x / 17.0) * np.cos(y / 9.0)
    out, inc = oracle_unwrap(wrap(true), return_increments=True)
    d = (out - true) / TWO_PI
    assert np.abs(d - np.round(d[0, 0])).max() < 1e-12          # one global multiple of 2 pi
    assert inc.max() - inc.min() >= 3                           # the ramp really wraps several times


def test_oracle_matches_numpy_unwrap_on_one_dimensional_variation():
    t = np.linspace(0, 40, 128)
    for axis, w in ((1, wrap(t[None, :] * np.ones((32, 1)))), (0, wrap(t[:, None] * np.ones((1, 32))))):
        out = oracle_unwrap(w)
        ref = np.unwrap(w, axis=axis)
        np.testing.assert_allclose(out - out[0, 0], ref - ref[0, 0], atol=1e-12)


def test_oracle_changes_by_integer_multiples_only():
    for name, f in fields(40, 56, 5).items():
        w = np.angle(f)
        out, inc = oracle_unwrap(w, return_increments=True)
        np.testing.assert_array_equal(out, w + TWO_PI * inc)
        if name == 'flat':
            assert not inc.any()


def test_oracle_unwrapped_neighbours_are_continuous_away_from_residues():
    f = fields(64, 64, 1)['smooth']
    out = oracle_unwrap(np.angle(f))
    assert np.abs(np.diff(out, axis=0)).max() < np.pi and np.abs(np.diff(out, axis=1)).max() < np.pi


# ----------------------------------------------------------------------------- kernel sources in emulation (CPU)
@pytest.mark.parametrize('shape', [(32, 32), (32, 64), (64, 32), (128, 64)])
def test_emulated_unwrap_equals_oracle_on_same_angles(shape):
    from tests.emu_harness import EmuPlan
    ny, nx = shape
    pl = EmuPlan(nx, ny)
    fs = fields(ny, nx, ny + nx)
    ang = np.stack([np.angle(f) for f in fs.values()])
    got = pl.unwrap_phase(ang)                                   # kind 1: float64 wrapped angles, four planes
    for k, name in enumerate(fs):
        want, inc = oracle_unwrap(ang[k], return_increments=True)
        np.testing.assert_array_equal(got[k], want, err_msg=name)
    pl.close()


@pytest.mark.parametrize('shape', [(32, 32), (48, 80)])
def test_emulated_device_merging_equals_host_merging(shape):
    """Spanning tree by Boruvka rounds on the device + anchor pass over the tree edges (the default) against the
    offset-carrying union-find over ALL edges on the host (option unwrap_merge = 1): the same integer field, global
    offset included — also when every reliability ties (flat / exactly linear phase: rank = edge id)."""
    from tests.emu_harness import EmuPlan
    ny, nx = shape
    pl = EmuPlan(nx, ny)
    fs = fields(ny, nx, 5 * ny + nx)
    y, x = np.mgrid[0:ny, 0:nx].astype(np.float64)
    planes = [np.angle(f) for f in fs.values()] + [np.zeros((ny, nx)), wrap(0.25 * x), wrap(0.5 * x - 0.25 * y)]
    ang = np.stack(planes)
    got = pl.unwrap_phase(ang)
    pl.set_option('unwrap_anchor', 1)         # the surviving group by bisection on the device, down to single pixels
    pl.set_option('unwrap_tail', 0)
    np.testing.assert_array_equal(pl.unwrap_phase(ang), got)
    pl.set_option('unwrap_tail', 100)         # ... and with the last levels on the host
    np.testing.assert_array_equal(pl.unwrap_phase(ang), got)
    pl.set_option('unwrap_merge', 1)
    host = pl.unwrap_phase(ang)
    np.testing.assert_array_equal(got, host)
    for k in range(len(planes)):
        np.testing.assert_array_equal(got[k], oracle_unwrap(ang[k]), err_msg=str(k))
    pl.close()


def test_emulated_unwrap_of_complex_field_and_mask():
    from tests.emu_harness import EmuPlan
    ny, nx = 64, 64
    pl = EmuPlan(nx, ny)
    fs = fields(ny, nx, 11)
    y, x = np.mgrid[0:ny, 0:nx]
    envelope = np.exp(-((x - 32.0) ** 2 + (y - 30.0) ** 2) / 60.0)
    psi = np.stack([fs['smooth'] * envelope, fs['vortices'] * envelope])
    got = pl.unwrap_phase(psi)
    masked = pl.unwrap_phase(psi, mask=True)
    for c in range(2):
        w = np.angle(psi[c])
        want_inc = oracle_unwrap(w, return_increments=True)[1]
        np.testing.assert_array_equal(increments(got[c], w), want_inc)
        dens = np.abs(psi[c]) ** 2
        keep = dens >= dens.max() * 1e-6
        assert (~keep).any() and keep.any()
        np.testing.assert_array_equal(masked[c][keep], got[c][keep])
        assert not masked[c][~keep].any()
    pl.close()


def test_emulated_unwrap_complex64_plan():
    from tests.emu_harness import EmuPlan
    pl = EmuPlan(64, 32, dtype=np.complex64)
    f = fields(32, 64, 2)['smooth'].astype(np.complex64)
    got = pl.unwrap_phase(f[None])[0]
    w = np.arctan2(f.imag.astype(np.float64), f.real.astype(np.float64))
    np.testing.assert_array_equal(increments(got, w), oracle_unwrap(w, return_increments=True)[1])
    pl.close()


@pytest.mark.parametrize('case,run', ENERGY_ROBUST)
def test_emulated_energy_with_unwrapping_vs_oracle(case, run):
    from tests.emu_harness import plan_from_problem
    z = np.load(os.path.join(GOLDEN, case + '.npz'))
    pre = f'r{run}_'
    prob = orc.Problem.from_golden(z, pre)
    pl = plan_from_problem(prob, str(z[pre + 'mode']), float(z[pre + 'dt']))
    psik = z[pre + 'psik_final']
    want = orc.energy(prob, psik, unwrap=oracle_unwrap)
    got = pl.energy(psik, 2 * prob.kL * prob.is_coupling, 2)[0]
    np.testing.assert_allclose(got, want, rtol=1e-10)
    if case != 'ground_64':           # everywhere else the unwrapping changes the reference's number
        assert abs(want[1] - z[pre + 'energy_identity_unwrap'][1]) > 1e-3 * abs(want[1])
    pl.close()


@pytest.mark.parametrize('case,run', [('cgrad_64', 0), ('nocoupl_64', 0)])
def test_emulated_energy_of_every_step_with_unwrapping(case, run):
    """sgpe_full_steps_energy(unwrap_mode 2): eng_expect in the reference's definition (phase unwrapped) after every
    full step against the oracle stepping beside it; state and populations as without the tracking."""
    from tests.emu_harness import plan_from_problem
    z = np.load(os.path.join(GOLDEN, case + '.npz'))
    pre = f'r{run}_'
    prob = orc.Problem.from_golden(z, pre)
    mode, dt, n = str(z[pre + 'mode']), float(z[pre + 'dt']), int(z[pre + 'n_steps'])
    kl = 2 * prob.kL * prob.is_coupling
    pl = plan_from_problem(prob, mode, dt)
    pops, eng = pl.full_steps_energy(n, kl, 2)
    np.testing.assert_allclose(pops[0], z[pre + 'pops_vals'], rtol=1e-12)
    np.testing.assert_allclose(np.abs(pl.store()[0] - z[pre + 'psik_final']).max(), 0, atol=1e-12 * np.abs(z[pre + 'psik_final']).max())
    o = orc.OraclePropagator(prob, dt, mode)
    for i in range(n):
        o.full_step()
        want = orc.energy(prob, o.psik, unwrap=oracle_unwrap)
        if i == n - 1:
            np.testing.assert_allclose(eng[0, i], want, rtol=1e-10, err_msg=f'step {i}')
        # (E_pot and E_int do not depend on the phase; E_kin of an intermediate state may sit on a 2 pi decision)
        np.testing.assert_allclose(eng[0, i, 2:], want[2:], rtol=1e-10, err_msg=f'step {i}')
    pl.close()


def test_unwrap_rejects_bad_arguments():
    from tests.emu_harness import EmuPlan
    from spinor_gpe_b200._capi import SgpeError
    pl = EmuPlan(32, 32)
    ang = np.zeros((1, 32, 32))
    with pytest.raises(SgpeError):
        pl.unwrap_phase(ang, mask=True)                           # a mask needs the complex field
    with pytest.raises(SgpeError):
        pl.energy(np.ones((2, 32, 32), dtype=complex), 0.0, 3)    # grid not set / bad mode
    pl.close()


# ----------------------------------------------------------------------------- CUDA library (B200)
@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(64, 64), (128, 256), (512, 512)])
def test_gpu_unwrap_equals_oracle_on_same_angles(shape):
    """Same wrapped angles in -> the same integer field out, bit for bit: keys are computed with explicit
    round-to-nearest operations, the radix sort is stable, the union-find follows the published merge rules."""
    import torch
    from spinor_gpe_b200.plan import Plan
    ny, nx = shape
    pl = Plan(nx, ny)
    fs = fields(ny, nx, 3 * ny + nx)
    ang = np.stack([np.angle(f) for f in fs.values()])
    got = pl.unwrap_phase(torch.from_numpy(ang)).cpu().numpy()
    pl.set_option('unwrap_sort', 1)                               # edges sorted on the host: same order, same result
    got_host_sort = pl.unwrap_phase(torch.from_numpy(ang)).cpu().numpy()
    for k, name in enumerate(fs):
        want = oracle_unwrap(ang[k])
        np.testing.assert_array_equal(got[k], want, err_msg=name)
        np.testing.assert_array_equal(got_host_sort[k], want, err_msg=name)
    pl.close()


@pytest.mark.gpu
def test_gpu_unwrap_of_complex_field_and_tensor_tools():
    import torch
    from spinor_gpe_b200 import tensor_tools as tt
    ny, nx = 128, 128
    fs = fields(ny, nx, 21)
    y, x = np.mgrid[0:ny, 0:nx]
    envelope = np.exp(-((x - 60.0) ** 2 + (y - 66.0) ** 2) / 300.0)
    psi = [fs['smooth'] * envelope, fs['smooth'].conj() * envelope]
    dens = [np.abs(p) ** 2 for p in psi]
    # NumPy in / NumPy out (the reference's host-side call, tensor_tools.py:528-539), computed through the library
    got = tt.phase(psi, uwrap=True, dens=dens)
    # CUDA tensors in / CUDA tensors out (beyond the reference)
    got_t = tt.phase([torch.as_tensor(p).cuda() for p in psi], uwrap=True,
                     dens=[torch.as_tensor(d).cuda() for d in dens])
    for c in range(2):
        w = np.angle(psi[c])
        want = oracle_unwrap(w)
        want[dens[c] < dens[c].max() * 1e-6] = 0
        np.testing.assert_array_equal(got[c], want)
        # device-side atan2 may differ from libm in the last bit: same integers, angles to rounding
        keep = dens[c] >= dens[c].max() * 1e-6
        np.testing.assert_allclose(got_t[c].cpu().numpy()[keep], want[keep], rtol=0, atol=1e-12)
        assert not got_t[c].cpu().numpy()[~keep].any()


@pytest.mark.gpu
@pytest.mark.parametrize('case,run', ENERGY_ROBUST)
def test_gpu_energy_with_unwrapping_vs_oracle(case, run):
    """eng_expect as the reference evaluates it (phase unwrapped) against the oracle, tolerance 1e-9 (north_star)."""
    import torch
    from tests.test_gpu_parity import GOLDEN_SPECS, build_case
    from spinor_gpe_b200 import TensorPropagator
    z = np.load(os.path.join(GOLDEN, case + '.npz'))
    pre = f'r{run}_'
    prob = orc.Problem.from_golden(z, pre)
    psik = z[pre + 'psik_final']
    want = orc.energy(prob, psik, unwrap=oracle_unwrap)
    spec = GOLDEN_SPECS[case]
    ps = build_case(spec)
    assert not any(step[0] == 'call' for step in spec['runs'])     # operators are the same for every run of these cases
    prop = TensorPropagator(ps, float(z[pre + 'dt']), 1, 'cuda', time=str(z[pre + 'mode']))
    assert prop.unwrap == 'herraez'
    got = prop.eng_expect([torch.as_tensor(psik[0]), torch.as_tensor(psik[1])])
    np.testing.assert_allclose(got, want, rtol=1e-9)
    got_none = prop.eng_expect([torch.as_tensor(psik[0]), torch.as_tensor(psik[1])], unwrap='none')
    np.testing.assert_allclose(got_none, z[pre + 'energy_identity_unwrap'], rtol=1e-9)


@pytest.mark.gpu
def test_gpu_unwrap_at_full_size_properties():
    """2048^2 (BASELINE configs[2] mesh): result - input is an exact integer field, neighbours continuous."""
    import torch
    from spinor_gpe_b200.plan import Plan
    n = 2048
    pl = Plan(n, n)
    y, x = torch.meshgrid(torch.arange(n, dtype=torch.float64, device='cuda'),
                          torch.arange(n, dtype=torch.float64, device='cuda'), indexing='ij')
    true = 0.011 * x + 0.007 * y + 6.0 * torch.sin(x / 97.0) * torch.cos(y / 131.0)
    psi = torch.polar(1.0 + 0.2 * torch.cos(x / 50.0), true)
    out = pl.unwrap_phase(psi[None])[0]
    k = (out - torch.angle(psi)) / TWO_PI
    assert float((k - torch.round(k)).abs().max()) < 1e-9
    d = (out - true) / TWO_PI
    assert float((d - torch.round(d[0, 0])).abs().max()) < 1e-9
    pl.close()


@pytest.mark.gpu
def test_gpu_device_merging_equals_host_merging_at_full_size():
    """2048^2, noise + vortices + a smooth ramp: the device-built spanning tree with the anchor pass gives the integer
    field of the all-host merging (option unwrap_merge = 1) bit for bit, global offset included."""
    import torch
    from spinor_gpe_b200.plan import Plan
    n = 2048
    pl = Plan(n, n)
    g = torch.Generator(device='cuda').manual_seed(7)
    y, x = torch.meshgrid(torch.arange(n, dtype=torch.float64, device='cuda'),
                          torch.arange(n, dtype=torch.float64, device='cuda'), indexing='ij')
    noise = torch.complex(torch.randn(n, n, generator=g, device='cuda', dtype=torch.float64),
                          torch.randn(n, n, generator=g, device='cuda', dtype=torch.float64))
    vort = torch.complex(x - 0.31 * n, y - 0.37 * n) * torch.complex(x - 0.62 * n, -(y - 0.71 * n))
    envelope = torch.exp(-((x - n / 2) ** 2 + (y - n / 2) ** 2) / (0.05 * n * n))
    planes = torch.stack([noise, vort * torch.exp(0.02j * x) + 30.0 * noise,
                          torch.polar(envelope, 0.011 * x + 0.007 * y), torch.ones_like(noise)])
    got = pl.unwrap_phase(planes)             # 4 planes: spanning tree and anchor bisection on the device
    pl.set_option('unwrap_anchor', 0)         # anchor pass over the tree edges on the host
    assert torch.equal(pl.unwrap_phase(planes), got)
    pl.set_option('unwrap_anchor', 1)
    pl.set_option('unwrap_tail', 0)           # bisection all the way down
    assert torch.equal(pl.unwrap_phase(planes), got)
    pl.set_option('unwrap_merge', 1)
    host = pl.unwrap_phase(planes)
    assert torch.equal(got, host)
    k = (got - torch.angle(planes)) / TWO_PI
    assert float((k - torch.round(k)).abs().max()) < 1e-9
    assert int(torch.round(k[1]).max() - torch.round(k[1]).min()) >= 2
    pl.close()


@pytest.mark.gpu
@pytest.mark.parametrize('case,run', [('cgrad_64', 0), ('nocoupl_64', 0)])
def test_gpu_energy_of_every_step_in_the_reference_definition(case, run):
    """PSpinor.imaginary / real(track_energy='herraez'): eng_expect with the phase unwrapped as the reference does,
    after EVERY full step (PropResult.eng_history), against the oracle stepping beside it; the final state and the
    populations are those of the golden run."""
    from tests.test_gpu_parity import GOLDEN_SPECS, build_case
    z = np.load(os.path.join(GOLDEN, case + '.npz'))
    pre = f'r{run}_'
    prob = orc.Problem.from_golden(z, pre)
    spec = GOLDEN_SPECS[case]
    ps = build_case(spec)
    mode, dt, n = str(z[pre + 'mode']), float(z[pre + 'dt']), int(z[pre + 'n_steps'])
    fn = ps.imaginary if mode == 'imag' else ps.real
    res, prop = fn(dt, n, 'cuda', track_energy='herraez')
    assert np.abs(np.array(res.psik) - z[pre + 'psik_final']).max() < 1e-10 * np.abs(z[pre + 'psik_final']).max()
    np.testing.assert_allclose(res.pops['vals'], z[pre + 'pops_vals'], rtol=1e-9)
    o = orc.OraclePropagator(prob, dt, mode)
    for i in range(n):
        o.full_step()
        want = orc.energy(prob, o.psik, unwrap=oracle_unwrap)
        np.testing.assert_allclose(res.eng_history[i, 2:], want[2:], rtol=1e-9, err_msg=f'step {i}')
        if i == n - 1:
            np.testing.assert_allclose(res.eng_history[i], want, rtol=1e-9)
    np.testing.assert_allclose(res.eng_history[-1], res.eng_final, rtol=1e-9)
