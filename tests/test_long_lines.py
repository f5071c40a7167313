"""Meshes with more than 4096 points along a line on ONE device: ``slab.LongLinePlan`` (a one-rank ``SlabPropagator``
without a process group — four-step lines, digit-transposed k order) behind ``TensorPropagator``.

CPU: the kernel sources in emulation with the four-step split forced on small meshes, against the oracle.
GPU: the public API (``PSpinor.imaginary / real``) with forced splits and with genuinely long lines (8192 points)."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import spinor_oracle as orc
from oracle.unwrap_oracle import unwrap_phase as oracle_unwrap


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def make_ps(mesh, raman=True, atom_num=1e4):
    from spinor_gpe_b200 import PSpinor
    w0 = 2 * np.pi * 50
    tmp = os.path.join(tempfile.mkdtemp(prefix='sgpe_ll_'), 'run') + os.sep
    if raman:
        ps = PSpinor(tmp, overwrite=True, atom_num=atom_num, omeg={'x': w0, 'y': 1.5 * w0, 'z': 40 * w0},
                     g_sc={'uu': 1, 'dd': 0.98, 'ud': 1.02}, r_sizes=(16, 12), mesh_points=mesh)
        ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
        ps.shift_momentum(scale=0.7, frac=(0.3, 0.7))
        ps.coupling_uniform(1.5 * ps.EL_recoil)
        ps.detuning_grad(-3.0)
        ps.rot_coupling = False
    else:
        ps = PSpinor(tmp, overwrite=True, atom_num=1e2, omeg={'x': w0, 'y': w0, 'z': 40 * w0},
                     g_sc={'uu': 1, 'dd': 1, 'ud': 1.04}, r_sizes=(8, 8), mesh_points=mesh)
        ps.coupling_setup(wavel=790.1e-9, kin_shift=False)
    return ps


def problem_of(ps):
    return orc.Problem(ps.psik, ps.kin_eng_spin, ps.pot_eng_spin, ps.coupling, ps.space['dr'], ps.space['dv_r'],
                       ps.space['dv_k'], [ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud']], ps.atom_num,
                       x=ps.space['x'], kL=ps.kL_recoil, is_coupling=ps.is_coupling, rot_coupling=ps.rot_coupling)


def emu_plan(ps, dt, mode, splits):
    from spinor_gpe_b200.slab import LongLinePlan
    from tests.emu_harness import emu_lib
    nx, ny = int(ps.space['mesh_points'][0]), int(ps.space['mesh_points'][1])
    mk = lambda: [torch.zeros(2 * nx * ny, dtype=torch.complex128)]       # noqa: E731
    return LongLinePlan(ps, dt, mode, 'cpu', 'c128', plan_kwargs={'_lib': emu_lib()}, split_x=splits[0],
                        split_y=splits[1], exchange_buffers={'k': mk(), 'r': mk()})


@pytest.mark.parametrize('mode,splits,mesh', [('real', (32, None), (1024, 64)), ('imag', (None, 32), (64, 1024)),
                                               ('imag', (None, None), (128, 64))])
def test_emulated_long_line_plan_against_oracle(mode, splits, mesh):
    ps = make_ps(mesh)
    dt, n = (1 / 50, 2) if mode == 'imag' else (1 / 2000, 2)
    prob = problem_of(ps)
    want = orc.OraclePropagator(prob, dt, mode).run(n)
    pl = emu_plan(ps, dt, mode, splits)
    pops = torch.zeros((1, n + 1, 2), dtype=torch.float64)
    pl.full_steps(n, pops, first=0)
    assert rel(pl.store()[0].numpy(), want['psik']) < 1e-12
    np.testing.assert_allclose(pops[0, :n].numpy(), want['pops_vals'], rtol=1e-12)
    assert rel(pl.real_space()[0].numpy(), want['psi']) < 1e-12          # ttools.ifft_2d of the state
    # the state survives reading it and its real-space image: one more step == the oracle's n + 1 steps
    pl.full_steps(1, pops, first=n)
    want2 = orc.OraclePropagator(prob, dt, mode).run(n + 1)
    assert rel(pl.store()[0].numpy(), want2['psik']) < 1e-12
    np.testing.assert_allclose(pops[0].numpy(), want2['pops_vals'], rtol=1e-12)
    # load() replaces the state (tensor and NumPy input), a sub-step continues from it
    o = orc.OraclePropagator(prob, dt, mode)
    o.single_step(o.ops_out)
    for given in (torch.as_tensor(np.array(ps.psik)), np.array(ps.psik)):
        pl.load(given)
        pl.single_step(pl.substeps()[0])
        assert rel(pl.store()[0].numpy(), o.psik.numpy()) < 1e-13
    pl.close()


@pytest.mark.parametrize('splits,mesh', [((32, None), (1024, 64)), ((None, None), (64, 64))])
def test_emulated_long_line_energy(splits, mesh):
    """eng_expect on the row plan of the long-line machinery, wrapped and unwrapped phase, vs the oracle."""
    ps = make_ps(mesh, raman=False)
    prob = problem_of(ps)
    want = orc.OraclePropagator(prob, 1 / 50, 'imag').run(3)
    pl = emu_plan(ps, 1 / 50, 'imag', splits)
    pl.full_steps(3)
    kl = 2 * ps.kL_recoil
    np.testing.assert_allclose(pl.energy(None, kl, 'none')[0].numpy(), orc.energy(prob, want['psik']), rtol=1e-10)
    np.testing.assert_allclose(pl.energy(None, kl, 'herraez')[0].numpy(),
                               orc.energy(prob, want['psik'], unwrap=oracle_unwrap), rtol=1e-10)
    # the energy of ANOTHER state (not normalised to the atom number) leaves the current one alone
    other = 0.5 * np.array(ps.psik)
    np.testing.assert_allclose(pl.energy(torch.as_tensor(other), kl, 'none')[0].numpy(), orc.energy(prob, other),
                               rtol=1e-10)
    assert rel(pl.store()[0].numpy(), want['psik']) < 1e-12
    pl.close()


# ----------------------------------------------------------------------------- CUDA (B200), public API
@pytest.mark.gpu
@pytest.mark.parametrize('mode,mesh,long_lines', [('real', (1024, 64), {'split_x': 32}),
                                                   ('imag', (64, 1024), {'split_y': 32}),
                                                   ('imag', (1024, 1024), {'split_x': 32, 'split_y': 32}),
                                                   ('real', (8192, 64), None),
                                                   ('imag', (64, 8192), None)])
def test_gpu_long_lines_through_public_api(mode, mesh, long_lines):
    """PSpinor.imaginary()/real() on meshes that need the four-step lines (forced on small meshes, genuine at 8192
    points) against the oracle: psi_k, psi, populations within the north_star tolerances."""
    ps = make_ps(mesh)
    dt, n = (1 / 50, 3) if mode == 'imag' else (1 / 2000, 3)
    want = orc.OraclePropagator(problem_of(ps), dt, mode).run(n, n_samples=3)
    res, prop = (ps.imaginary if mode == 'imag' else ps.real)(dt, n, 'cuda', is_sampling=True, n_samples=3,
                                                              long_lines=long_lines, unwrap='none')
    assert prop._long
    assert rel(np.array(res.psik), want['psik']) < 1e-10
    assert rel(np.array(res.psi), want['psi']) < 1e-10
    np.testing.assert_allclose(res.pops['vals'], want['pops_vals'], rtol=1e-9)
    with np.load(res.sampled_path) as smp:
        assert rel(smp['psiks'], want['sampled_psiks']) < 1e-10
    # single sub-steps and the lazily materialised psik
    o = orc.OraclePropagator(problem_of(ps), dt, mode)     # ps.psik is now the final state of the run above
    o.single_step(o.ops_in)
    from spinor_gpe_b200 import TensorPropagator
    prop = TensorPropagator(ps, dt, n, 'cuda', time=mode, long_lines=long_lines)
    prop.single_step(prop.dt_in, prop.eng_in)
    assert rel(np.array([p.cpu().numpy() for p in prop.psik]), o.psik.numpy()) < 1e-10


@pytest.mark.gpu
def test_gpu_long_lines_energy_and_ground_state():
    """Ground state on a 8192 x 64 mesh: energy (wrapped and unwrapped phase) against the oracle."""
    ps = make_ps((8192, 64), raman=False)
    prob = problem_of(ps)
    want = orc.OraclePropagator(prob, 1 / 50, 'imag').run(4)
    res, prop = ps.imaginary(1 / 50, 4, 'cuda')
    assert prop._long and prop.unwrap == 'herraez'
    assert rel(np.array(res.psik), want['psik']) < 1e-10
    np.testing.assert_allclose(prop.eng_expect(None, unwrap='none'), want['energy'], rtol=1e-9)
    np.testing.assert_allclose(res.eng_final, orc.energy(prob, want['psik'], unwrap=oracle_unwrap), rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize('mesh,long_lines', [((1024, 64), {'split_x': 32}), ((8192, 64), None)])
def test_gpu_long_lines_energy_tracking_and_spectral_kinetic(mesh, long_lines):
    """Per-step energy (PropResult.eng_history) and the spectral kinetic energy on four-step lines, against the oracle
    and NumPy."""
    ps = make_ps(mesh)
    prob = problem_of(ps)
    n = 3
    o = orc.OraclePropagator(prob, 1 / 50, 'imag')
    want = []
    for _ in range(n):
        o.full_step()
        want.append(orc.energy(prob, o.psik))
    res, prop = ps.imaginary(1 / 50, n, 'cuda', long_lines=long_lines, unwrap='none', track_energy=True)
    assert prop._long
    assert rel(np.array(res.psik), o.psik.numpy()) < 1e-10
    np.testing.assert_allclose(res.eng_history[:, 2:], np.array(want)[:, 2:], rtol=1e-9)
    np.testing.assert_allclose(res.eng_history[-1], res.eng_final, rtol=1e-9)
    kin = [(np.asarray(prob.kin[c]) * np.abs(res.psik[c]) ** 2).sum() * prob.dv_k for c in range(2)]
    np.testing.assert_allclose(prop.kin_expect_spectral(), kin, rtol=1e-9)
