"""GPU parity tests (run on the B200 with ``-m gpu``): the CUDA path, called through the C ABI
(spinor_gpe_b200.plan.Plan / the TensorPropagator drop-in), against

* the golden fixtures produced by the unmodified reference (tests/golden),
* the CPU oracle on seeded inputs at sizes it finishes in seconds,
* size-independent properties at BASELINE.json's full sizes (2048^2, 4096^2).

Tolerances are the north_star's: psi rel-L2 <= 1e-10 (complex128) / 1e-5 (complex64); energy and atom
number <= 1e-9 relative."""
import glob
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import spinor_oracle as orc
from oracle.gen_golden import CASES as GOLDEN_SPECS
from tests.test_pspinor_setup import build_case

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, '*.npz'))
               if 'tensor_tools' not in p)
ENERGY_OK = {('cgrad_64', 0), ('cgrad_64', 1), ('ground_64', 0), ('nocoupl_64', 0), ('raman_64x32', 1)}
TOL_PSI, TOL_PSI_C64, TOL_SCALAR = 1e-10, 1e-5, 1e-9


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def problem_of(ps):
    return orc.Problem(ps.psik, ps.kin_eng_spin, ps.pot_eng_spin, ps.coupling, ps.space['dr'], ps.space['dv_r'],
                       ps.space['dv_k'], [ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud']], ps.atom_num,
                       x=ps.space['x'], kL=ps.kL_recoil, is_coupling=ps.is_coupling, rot_coupling=ps.rot_coupling)


def make_ps(mesh, **kw):
    from spinor_gpe_b200 import PSpinor
    w0 = 2 * np.pi * 50
    args = dict(atom_num=1e3, omeg={'x': w0, 'y': w0, 'z': 40 * w0}, g_sc={'uu': 1, 'dd': 1, 'ud': 1.04},
                pop_frac=(0.5, 0.5), r_sizes=(8, 8), mesh_points=mesh)
    args.update(kw)
    tmp = os.path.join(tempfile.mkdtemp(prefix='sgpe_g_'), 'run') + os.sep
    return PSpinor(tmp, overwrite=True, **args)


def test_native_library_is_the_one_running():
    from spinor_gpe_b200 import _lib
    assert b'sm_100a' in _lib.lib().sgpe_version()
    loaded = open('/proc/self/maps').read()
    assert 'libsgpe.so' in loaded


# ------------------------------------------------------------------ golden fixtures (the reference itself)
@pytest.mark.parametrize('case', CASES)
def test_golden_through_public_api(case):
    """Replays every run of the fixture through PSpinor.imaginary()/real() -> TensorPropagator."""
    z = np.load(os.path.join(GOLDEN, case + '.npz'))
    spec = GOLDEN_SPECS[case]
    ps = build_case(spec)
    run_idx = 0
    for run in spec['runs']:
        if run[0] == 'call':
            from oracle.gen_golden import _resolve
            getattr(ps, run[1])(*_resolve(ps, run[2]))
            continue
        if run[0] == 'zero':                       # e.g. the trap switched off between two runs
            setattr(ps, run[1], np.zeros_like(getattr(ps, run[1])))
            continue
        mode, dt, n = run
        pre = f'r{run_idx}_'
        from spinor_gpe_b200 import TensorPropagator
        prop = TensorPropagator(ps, dt, n, 'cuda', time=mode)
        prop.single_step(prop.dt_out, prop.eng_out)
        assert rel(np.array([p.cpu().numpy() for p in prop.psik]), z[pre + 'psik_single_out']) < TOL_PSI
        prop = TensorPropagator(ps, dt, n, 'cuda', time=mode)
        prop.single_step(prop.dt_in, prop.eng_in)
        assert rel(np.array([p.cpu().numpy() for p in prop.psik]), z[pre + 'psik_single_in']) < TOL_PSI
        prop = TensorPropagator(ps, dt, n, 'cuda', time=mode)
        prop.full_step()
        assert rel(np.array([p.cpu().numpy() for p in prop.psik]), z[pre + 'psik_full1']) < TOL_PSI

        fn = ps.imaginary if mode == 'imag' else ps.real
        res, prop = fn(dt, n, 'cuda', is_sampling=True, n_samples=2)
        assert rel(np.array(res.psik), z[pre + 'psik_final']) < TOL_PSI
        assert rel(np.array(res.psi), z[pre + 'psi_final']) < TOL_PSI
        np.testing.assert_allclose(res.pops['vals'], z[pre + 'pops_vals'], rtol=TOL_SCALAR)
        np.testing.assert_allclose(res.pops['vals'].sum(axis=1), z[pre + 'pops_vals'].sum(axis=1), rtol=TOL_SCALAR)
        np.testing.assert_array_equal(res.pops['times'], z[pre + 'pops_times'])
        with np.load(res.sampled_path) as smp:
            assert rel(smp['psiks'], z[pre + 'sampled_psiks']) < TOL_PSI
            np.testing.assert_array_equal(smp['times'], z[pre + 'sampled_times'])
        assert os.path.basename(res.sampled_path) == f"psik_sampled{run_idx + 1}-{ps.paths['folder']}.npz"
        # the fixture's energy was generated with the identity in place of skimage's unwrap_phase (not installed
        # where the reference ran); the unwrapped variant is checked in tests/test_unwrap.py.  E_pot and E_int
        # (tensor_propagator.py:313-318) depend on the densities only: pinned on EVERY run.  E_kin / E_tot carry the
        # finite-difference phase gradient, meaningful only where the wrapped phase is conditioned (ENERGY_OK).
        got_e = prop.eng_expect(None, unwrap='none')
        np.testing.assert_allclose(got_e[2:], z[pre + 'energy_identity_unwrap'][2:], rtol=TOL_SCALAR)
        if (case, run_idx) in ENERGY_OK:
            np.testing.assert_allclose(got_e, z[pre + 'energy_identity_unwrap'], rtol=TOL_SCALAR)
        run_idx += 1


def test_tensor_tools_vectors_on_gpu():
    from spinor_gpe_b200 import tensor_tools as tt
    z = np.load(os.path.join(GOLDEN, 'tensor_tools_vectors.npz'))
    for tag in 'ab':
        psi = [torch.as_tensor(p).cuda() for p in z[f'{tag}_psi']]
        dr = z[f'{tag}_dr']
        got = lambda lst: np.array([t.cpu().numpy() for t in lst])   # noqa: E731
        assert rel(got(tt.fft_2d(psi, dr)), z[f'{tag}_fft2']) < 1e-13
        assert rel(got(tt.ifft_2d(psi, dr)), z[f'{tag}_ifft2']) < 1e-13
        for ax in (0, 1):
            assert rel(got(tt.fft_1d(psi, dr, ax)), z[f'{tag}_fft1_ax{ax}']) < 1e-13
            assert rel(got(tt.ifft_1d(psi, dr, ax)), z[f'{tag}_ifft1_ax{ax}']) < 1e-13
        pn, dn = tt.norm(psi, 0.125, 1234.5)
        assert rel(got(pn), z[f'{tag}_norm_psi']) < 1e-13
        assert rel(got(dn), z[f'{tag}_norm_dens']) < 1e-13
        np.testing.assert_allclose(tt.calc_pops(psi, 0.125), z[f'{tag}_pops'], rtol=1e-12)


def test_gradients_and_phase_of_cuda_tensors():
    """ttools.grad / grad_sq / phase on CUDA tensors (the reference raises for tensors, tensor_tools.py:343-345, 533-536):
    the device stencils of sgpe_gradient against np.gradient on the same data, NumPy inputs untouched."""
    from spinor_gpe_b200 import tensor_tools as tt
    rng = np.random.default_rng(5)
    for shape, dr in (((64, 128), (0.25, 0.5)), ((2048, 256), (0.01, 0.02)), ((30, 50), (1.0, 3.0))):
        psi = [rng.standard_normal(shape) + 1j * rng.standard_normal(shape) for _ in range(2)]
        want = tt.grad(psi, dr)                                  # NumPy branch = the reference's
        got = tt.grad([torch.as_tensor(p).cuda() for p in psi], dr)
        for gc, wc in zip(got, want):
            for g, w in zip(gc, wc):
                assert g.is_cuda and rel(g.cpu().numpy(), w) < 1e-14
        dens = [np.abs(p) ** 2 for p in psi]
        got_sq = tt.grad_sq([torch.as_tensor(np.sqrt(d)).cuda() for d in dens], dr)
        want_sq = tt.grad_sq([np.sqrt(d) for d in dens], dr)
        for g, w in zip(got_sq, want_sq):
            assert rel(g.cpu().numpy(), w) < 1e-13
        ph = tt.phase([torch.as_tensor(p).cuda() for p in psi], uwrap=False, dens=[torch.as_tensor(d).cuda() for d in dens])
        ph_want = tt.phase(psi, uwrap=False, dens=dens)
        for g, w in zip(ph, ph_want):
            np.testing.assert_allclose(g.cpu().numpy(), w, rtol=0, atol=1e-15)
    f32 = torch.as_tensor(rng.standard_normal((64, 64)).astype(np.float32)).cuda()
    g32 = tt.grad_comp(f32, (0.5, 0.5))
    assert g32[0].dtype == torch.float32 and rel(g32[1].cpu().numpy(), np.gradient(f32.cpu().numpy(), 0.5, 0.5)[1]) < 1e-6
    with pytest.raises(RuntimeError):
        tt.grad_comp(torch.zeros(8, 8), (1, 1))                  # no CPU fallback for tensors


def test_reference_fft_invariants_on_gpu():
    """Port of the reference's own tests (spinor_gpe/tests/fft_func_tests.py:27-472): all-ones grids of
    128..1024 points, delta_r=(1,1): round trip, 1-D x 1-D == 2-D, Parseval with vol_elem 4 pi^2 / N."""
    from spinor_gpe_b200 import tensor_tools as tt
    eps = 10 * 2.2e-16
    for n in (128, 256, 512, 1024):
        ones = [torch.ones((n, n), dtype=torch.complex128, device='cuda') for _ in range(2)]
        dr = (1, 1)
        back = tt.ifft_2d(tt.fft_2d(ones, dr), dr)
        assert max(float((b - o).abs().max()) for b, o in zip(back, ones)) < eps
        for first, second in ((0, 1), (1, 0)):
            two = tt.fft_1d(tt.fft_1d(ones, dr, first), dr, second)
            ref = tt.fft_2d(ones, dr)
            assert max(float((a - b).abs().max()) for a, b in zip(two, ref)) < eps * n * n
            two = tt.ifft_1d(tt.ifft_1d(ones, dr, first), dr, second)
            ref = tt.ifft_2d(ones, dr)
            assert max(float((a - b).abs().max()) for a, b in zip(two, ref)) < eps * n * n
        n_r = tt.calc_atoms(ones, 1.0)
        n_k = tt.calc_atoms(tt.fft_2d(ones, dr), 4 * np.pi ** 2 / (n * n))
        assert abs(n_r - n_k) < eps * n * n


# ------------------------------------------------------------------ oracle at larger sizes
SEEDED = [
    # mesh (nx, ny), mode, dt, steps, coupling kind, rotating frame, kin_shift
    ((256, 256), 'imag', 1 / 50, 10, 'zero', True, False),
    ((512, 256), 'real', 1 / 2000, 8, 'uniform', False, True),
    ((256, 1024), 'imag', 1 / 50, 6, 'dense', True, True),
    ((1024, 1024), 'real', 1 / 5000, 4, 'uniform', False, True),
    ((2048, 128), 'real', 1 / 2000, 4, 'dense', False, True),
    ((128, 2048), 'imag', 1 / 50, 4, 'none', True, False),
    ((4096, 64), 'imag', 1 / 50, 3, 'uniform', True, True),
    ((64, 4096), 'real', 1 / 2000, 3, 'zero', True, False),
]


@pytest.mark.parametrize('separable', [True, False])
@pytest.mark.parametrize('mesh,mode,dt,n,cpl,rot,kshift', SEEDED)
def test_against_oracle(mesh, mode, dt, n, cpl, rot, kshift, separable):
    ps = make_ps(mesh, atom_num=1e4, r_sizes=(16, 16), g_sc={'uu': 1, 'dd': 0.98, 'ud': 1.02})
    if cpl != 'none':
        ps.coupling_setup(wavel=790.1e-9, kin_shift=kshift)
        if kshift:
            ps.shift_momentum(scale=0.7, frac=(0.3, 0.7))
        if cpl == 'uniform':
            ps.coupling_uniform(1.5 * ps.EL_recoil)
        elif cpl == 'dense':
            ps.coupling_grad(slope=0.3, offset=2.0, axis=1)
        ps.detuning_grad(-3.0)
    ps.rot_coupling = rot
    rng = np.random.default_rng(99999)          # break the symmetry of the TF state with seeded noise
    noise = [1 + 0.05 * (rng.standard_normal(p.shape) + 1j * rng.standard_normal(p.shape)) for p in ps.psik]
    ps.psik = [p * q for p, q in zip(ps.psik, noise)]
    want = orc.OraclePropagator(problem_of(ps), dt, mode).run(n)
    res, prop = (ps.imaginary if mode == 'imag' else ps.real)(dt, n, 'cuda', separable=separable)
    assert prop.separable == {'kin': separable, 'pot': separable}
    assert rel(np.array(res.psik), want['psik']) < TOL_PSI
    np.testing.assert_allclose(res.pops['vals'], want['pops_vals'], rtol=TOL_SCALAR)
    assert abs(res.pops['vals'][-1].sum() / ps.atom_num - 1) < TOL_SCALAR


def test_headline_config_against_oracle():
    """BASELINE configs[2] itself — 2048^2 complex128, imaginary time, the benchmark's parameters
    (benchmarks/benchmark_prop.py:50-67) — against the oracle: psi_k, per-step populations, atom number and the
    energy of the final state (E_pot / E_int always; E_kin / E_tot for this smooth rotating-frame ground state)."""
    import bench
    ps = bench.build_problem(2048)
    n = 3
    want = orc.OraclePropagator(problem_of(ps), 1 / 50, 'imag').run(n)
    res, prop = ps.imaginary(1 / 50, n, 'cuda', unwrap='none')
    assert prop.separable == {'kin': True, 'pot': True}
    assert rel(np.array(res.psik), want['psik']) < TOL_PSI
    assert rel(np.array(res.psi), want['psi']) < TOL_PSI
    np.testing.assert_allclose(res.pops['vals'], want['pops_vals'], rtol=TOL_SCALAR)
    assert abs(res.pops['vals'][-1].sum() / ps.atom_num - 1) < TOL_SCALAR
    np.testing.assert_allclose(res.eng_final[2:], want['energy'][2:], rtol=TOL_SCALAR)
    np.testing.assert_allclose(res.eng_final, want['energy'], rtol=TOL_SCALAR)
    # the energy of every step on the same configuration (the "energy tracking" of configs[2]): six steps, i.e. the polar
    # side chain (persistent inverse column pass, polar row pass, shuffle stencils) launched normally AND replayed from the
    # captured graph; step 3 against the oracle, the last step against the stand-alone evaluation of the final state
    ps3 = bench.build_problem(2048)
    res3, prop3 = ps3.imaginary(1 / 50, 6, 'cuda', unwrap='none', track_energy=True)
    np.testing.assert_allclose(res3.eng_history[2], want['energy'], rtol=TOL_SCALAR)
    np.testing.assert_allclose(res3.pops['vals'][:3], want['pops_vals'], rtol=TOL_SCALAR)
    np.testing.assert_allclose(res3.eng_history[-1], res3.eng_final, rtol=TOL_SCALAR)
    assert np.all(np.diff(res3.eng_history[:, 0]) < 0)          # imaginary time: the energy goes down every step
    # the general (dense-operator) kernels on the same configuration
    ps2 = bench.build_problem(2048)
    res2, prop2 = ps2.imaginary(1 / 50, n, 'cuda', unwrap='none', separable=False)
    assert prop2.separable == {'kin': False, 'pot': False}
    assert rel(np.array(res2.psik), want['psik']) < TOL_PSI
    np.testing.assert_allclose(res2.pops['vals'], want['pops_vals'], rtol=TOL_SCALAR)


def test_config2_raman_kick_real_time_against_oracle():
    """BASELINE configs[1] as written (examples/3_raman_rabi.py:126-183 at 1024^2): Raman set-up with the
    spin-dependent kinetic shift, momentum kick, laboratory-frame coupling phase, a short imaginary-time relaxation,
    then real-time Rabi flopping under a uniform coupling — 12 steps against the oracle."""
    ps = make_ps((1024, 1024), atom_num=1e4, g_sc={'uu': 1, 'dd': 1, 'ud': 0.0}, pop_frac=(1.0, 0.0),
                 r_sizes=(16, 16))
    ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
    ps.shift_momentum(scale=1.0, frac=(0, 1.0))
    ps.rot_coupling = False
    ps.rand_seed = 99999
    o = orc.OraclePropagator(problem_of(ps), 1 / 50, 'imag')
    relax = o.run(4)
    res0, _ = ps.imaginary(1 / 50, 4, 'cuda', unwrap='none')
    assert rel(np.array(res0.psik), relax['psik']) < TOL_PSI
    np.testing.assert_allclose(res0.pops['vals'], relax['pops_vals'], rtol=TOL_SCALAR, atol=1e-9 * ps.atom_num)
    ps.coupling_uniform(1.0 * ps.EL_recoil)
    n = 12
    want = orc.OraclePropagator(problem_of(ps), 1 / 5000, 'real').run(n, n_samples=3)
    res1, prop1 = ps.real(1 / 5000, n, 'cuda', is_sampling=True, n_samples=3, unwrap='none')
    assert rel(np.array(res1.psik), want['psik']) < TOL_PSI
    assert rel(np.array(res1.psi), want['psi']) < TOL_PSI
    np.testing.assert_allclose(res1.pops['vals'], want['pops_vals'], rtol=TOL_SCALAR, atol=1e-9 * ps.atom_num)
    np.testing.assert_allclose(res1.pops['vals'].sum(axis=1), ps.atom_num, rtol=TOL_SCALAR)
    with np.load(res1.sampled_path) as smp:
        assert rel(smp['psiks'], want['sampled_psiks']) < TOL_PSI
    np.testing.assert_allclose(res1.eng_final[2:], want['energy'][2:], rtol=TOL_SCALAR)
    assert res1.pops['vals'][-1, 1] > 1e-3 * ps.atom_num       # the coupling does transfer population


def test_user_grids_of_any_memory_layout():
    """Operator grids set by the user as Fortran-ordered / transposed-view arrays (non-separable potential, dense
    coupling, spin-dependent detuning) reach the kernels as row-major float64: same result as the oracle, which reads
    them through NumPy indexing.  The plan keeps its own references to the device copies."""
    ps = make_ps((128, 256), atom_num=1e4, r_sizes=(16, 16))
    ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
    x, y = ps.space['x_mesh'], ps.space['y_mesh']
    pot = np.asfortranarray(0.5 * (x ** 2 + y ** 2) + 0.3 * np.sin(x) * np.cos(0.7 * y) + 0.05 * x * y)
    assert not pot.flags['C_CONTIGUOUS']
    ps.pot_eng = pot
    ps.detuning = np.ascontiguousarray((0.4 * x * y).T).T                 # transposed view, non-separable
    ps.coupling = np.asfortranarray(1.5 * ps.EL_recoil * (1 + 0.2 * np.tanh(x * y / 10)))
    ps.rot_coupling = False
    want = orc.OraclePropagator(problem_of(ps), 1 / 2000, 'real').run(4)
    res, prop = ps.real(1 / 2000, 4, 'cuda', unwrap='none')
    assert prop.separable['pot'] is False
    prop.pot_eng_spin = None                                              # the plan holds its own references
    assert rel(np.array(res.psik), want['psik']) < TOL_PSI
    np.testing.assert_allclose(res.pops['vals'], want['pops_vals'], rtol=TOL_SCALAR)
    np.testing.assert_allclose(res.eng_final[2:], want['energy'][2:], rtol=TOL_SCALAR)
    np.testing.assert_allclose(prop.eng_expect(None, unwrap='none')[2:], want['energy'][2:], rtol=TOL_SCALAR)


def test_coupling_energy_without_coupling_setup():
    """ps.coupling set without coupling_setup(): single_step skips the coupling operator (is_coupling False,
    tensor_propagator.py:252) but eng_expect still adds the coupling energy from self.coupling (:319-321)."""
    ps = make_ps((128, 128), atom_num=1e3, pop_frac=(0.7, 0.3))
    assert not ps.is_coupling
    x, y = ps.space['x_mesh'], ps.space['y_mesh']
    for cpl in (np.full_like(x, 0.8), 0.8 + 0.1 * x):
        ps.coupling = cpl
        want = orc.OraclePropagator(problem_of(ps), 1 / 50, 'imag').run(3)
        psik0 = [p.copy() for p in ps.psik]
        res, prop = ps.imaginary(1 / 50, 3, 'cuda', unwrap='none')
        ps.psik = psik0
        assert rel(np.array(res.psik), want['psik']) < TOL_PSI
        e_cpl_want = want['energy'][0] - sum(want['energy'][1:])
        e_cpl_got = res.eng_final[0] - sum(res.eng_final[1:])
        assert abs(e_cpl_want) > 1e-3 * abs(want['energy'][0])
        np.testing.assert_allclose(e_cpl_got, e_cpl_want, rtol=1e-7)          # a difference of large sums
        np.testing.assert_allclose(res.eng_final, want['energy'], rtol=TOL_SCALAR)


def test_bitwise_reproducible():
    """The same propagation repeated gives bit-identical states (fixed-order reductions, no float atomics)."""
    from spinor_gpe_b200 import TensorPropagator
    ps = make_ps((512, 256), atom_num=1e4, r_sizes=(16, 16))
    ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
    ps.coupling_uniform(1.5 * ps.EL_recoil)
    ps.rot_coupling = False
    outs = []
    for rep in range(6):
        prop = TensorPropagator(ps, 1 / 2000, 8, 'cuda', time='real')
        if rep == 5:
            prop._plan.set_option('prefetch', 0)
        prop._plan.full_steps(8)
        outs.append(torch.stack(prop.psik).clone())
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


@pytest.mark.parametrize('mesh,mode,dt,cpl,precision', [((256, 256), 'imag', 1 / 50, 'zero', 'c128'),
                                                        ((128, 1024), 'real', 1 / 2000, 'uniform', 'c128'),
                                                        ((64, 2048), 'imag', 1 / 50, 'dense', 'c128'),
                                                        ((64, 2048), 'real', 1 / 2000, 'uniform', 'c64'),
                                                        ((32, 4096), 'imag', 1 / 50, 'zero', 'c128')])
def test_two_barrier_groups_per_column_tile(mesh, mode, dt, cpl, precision):
    """col_tile = 3 (two independent barrier groups per column tile) against the oracle and the default kernel;
    repeated runs are bit-identical."""
    from spinor_gpe_b200 import TensorPropagator
    ps = make_ps(mesh, atom_num=1e4, r_sizes=(16, 16), g_sc={'uu': 1, 'dd': 0.98, 'ud': 1.02})
    ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
    if cpl == 'uniform':
        ps.coupling_uniform(1.5 * ps.EL_recoil)
    elif cpl == 'dense':
        ps.coupling_grad(slope=0.3, offset=2.0, axis=1)
    ps.rot_coupling = cpl != 'uniform'
    rng = np.random.default_rng(4242)
    ps.psik = [p * (1 + 0.05 * (rng.standard_normal(p.shape) + 1j * rng.standard_normal(p.shape))) for p in ps.psik]
    want = orc.OraclePropagator(problem_of(ps), dt, mode).run(3)
    outs = []
    for tile in (3, 3, 0):
        prop = TensorPropagator(ps, dt, 3, 'cuda', time=mode, precision=precision)
        prop._plan.set_option('col_tile', tile)
        pops = torch.zeros((1, 3, 2), dtype=torch.float64, device='cuda')
        prop._plan.full_steps(3, pops)
        outs.append((torch.stack(prop.psik).clone(), pops.cpu().numpy()[0]))
    tol = TOL_PSI if precision == 'c128' else TOL_PSI_C64
    assert rel(outs[0][0].cpu().numpy(), want['psik']) < tol
    np.testing.assert_allclose(outs[0][1], want['pops_vals'], rtol=TOL_SCALAR if precision == 'c128' else 1e-5)
    assert torch.equal(outs[0][0], outs[1][0])
    assert rel(outs[0][0].cpu().numpy(), outs[2][0].cpu().numpy()) < (1e-13 if precision == 'c128' else 1e-5)


@pytest.mark.parametrize('mesh,mode,dt,cpl,precision', [((256, 256), 'imag', 1 / 50, 'zero', 'c128'),
                                                        ((128, 1024), 'real', 1 / 2000, 'uniform', 'c128'),
                                                        ((512, 2048), 'imag', 1 / 50, 'dense', 'c128'),
                                                        ((64, 2048), 'real', 1 / 2000, 'uniform', 'c64'),
                                                        ((32, 4096), 'imag', 1 / 50, 'zero', 'c128'),
                                                        ((2048, 512), 'imag', 1 / 50, 'zero', 'c64')])
@pytest.mark.parametrize('kernel', [2, 3, 4, 5, 6, 7])
def test_persistent_column_pass(mesh, mode, dt, cpl, precision, kernel):
    """col_kernel = 2 / 3 (persistent column-pass CTAs, TMA-staged tiles; 2: split inverse exchange, 3: staging behind
    the inverse transform) against the oracle and the
    one-tile-per-CTA kernel; repeated runs are bit-identical; per-step energy tracking through it as well."""
    from spinor_gpe_b200 import TensorPropagator
    ps = make_ps(mesh, atom_num=1e4, r_sizes=(16, 16), g_sc={'uu': 1, 'dd': 0.98, 'ud': 1.02})
    ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
    if cpl == 'uniform':
        ps.coupling_uniform(1.5 * ps.EL_recoil)
    elif cpl == 'dense':
        ps.coupling_grad(slope=0.3, offset=2.0, axis=1)
    ps.rot_coupling = cpl != 'uniform'
    rng = np.random.default_rng(4243)
    ps.psik = [p * (1 + 0.05 * (rng.standard_normal(p.shape) + 1j * rng.standard_normal(p.shape))) for p in ps.psik]
    want = orc.OraclePropagator(problem_of(ps), dt, mode).run(3)
    outs = []
    for kern in (kernel, kernel, 1):
        prop = TensorPropagator(ps, dt, 3, 'cuda', time=mode, precision=precision)
        prop._plan.set_option('col_kernel', kern)
        pops = torch.zeros((1, 3, 2), dtype=torch.float64, device='cuda')
        eng = torch.zeros((1, 3, 4), dtype=torch.float64, device='cuda')
        prop._plan.full_steps(3, pops, energy=eng, kl_term=2 * ps.kL_recoil)
        outs.append((torch.stack(prop.psik).clone(), pops.cpu().numpy()[0], eng.cpu().numpy()[0]))
    tol = TOL_PSI if precision == 'c128' else TOL_PSI_C64
    assert rel(outs[0][0].cpu().numpy(), want['psik']) < tol
    np.testing.assert_allclose(outs[0][1], want['pops_vals'], rtol=TOL_SCALAR if precision == 'c128' else 1e-5)
    assert torch.equal(outs[0][0], outs[1][0])
    assert rel(outs[0][0].cpu().numpy(), outs[2][0].cpu().numpy()) < (1e-13 if precision == 'c128' else 1e-5)
    np.testing.assert_allclose(outs[0][2][:, 2:], outs[2][2][:, 2:], rtol=1e-9 if precision == 'c128' else 1e-4)


@pytest.mark.parametrize('mesh,mode,dt,batch', [((256, 256), 'imag', 1 / 50, 1), ((128, 64), 'real', 1 / 2000, 3),
                                                 ((1024, 512), 'imag', 1 / 50, 1)])
def test_graph_replay_matches_plain_launches(mesh, mode, dt, batch):
    """Option graph = 1: from the third step on sgpe_full_steps replays ONE captured steady-state step (six kernel
    nodes, the populations slot read from a device-side counter).  Bit-identical state and populations, also across
    several calls on the same plan (slot offsets) and after an operator change (re-capture); oracle parity."""
    from spinor_gpe_b200 import _capi
    from spinor_gpe_b200._separable import split_separable
    from spinor_gpe_b200.plan import Plan
    ps = make_ps(mesh, atom_num=1e4, r_sizes=(16, 16), g_sc={'uu': 1, 'dd': 0.98, 'ud': 1.02})
    ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
    rng = np.random.default_rng(11)
    ps.psik = [p * (1 + 0.05 * (rng.standard_normal(p.shape) + 1j * rng.standard_normal(p.shape))) for p in ps.psik]
    n1, n2 = 7, 5
    outs = []
    for graph in (1, 0):
        pl = Plan(mesh[0], mesh[1], batch)
        pl.set_grid(ps.space['dr'][0], ps.space['dr'][1], ps.space['dv_r'], ps.space['dv_k'], ps.atom_num)
        pl.set_interactions(ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud'])
        pl.set_kinetic_separable(*split_separable(np.array(ps.kin_eng_spin)))
        pl.set_potential_separable(*split_separable(np.array(ps.pot_eng_spin)))
        pl.set_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=np.linspace(0.5, 1.5, batch) * ps.EL_recoil)
        pl.set_time(mode, dt)
        pl.set_option('graph', graph)
        pl.load(np.stack([np.array(ps.psik)] * batch))
        pops = torch.zeros((batch, n1 + 2 * n2, 2), dtype=torch.float64, device='cuda')
        l0 = pl.launch_count()
        pl.full_steps(n1, pops, first=0)
        pl.full_steps(n2, pops, first=n1)                      # same graph, other slots
        pl.set_interactions(ps.g_sc['uu'], ps.g_sc['dd'], 0.5 * ps.g_sc['ud'])     # invalidates the captured step
        pl.full_steps(n2, pops, first=n1 + n2)
        # the same replay with the energy side chain (nine kernel nodes, energy slot from a second device-side counter)
        eng = torch.zeros((batch, n1 + n2, 4), dtype=torch.float64, device='cuda')
        pops_e = torch.zeros((batch, n1 + n2, 2), dtype=torch.float64, device='cuda')
        pl.full_steps(n1, pops_e, first=0, energy=eng, kl_term=2 * ps.kL_recoil)
        pl.full_steps(n2, pops_e, first=n1, energy=eng, kl_term=2 * ps.kL_recoil)
        outs.append((pl.store().clone(), pops.clone(), pl.launch_count() - l0, eng.clone(), pops_e.clone()))
        pl.close()
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][3], outs[1][3]) and torch.equal(outs[0][4], outs[1][4])
    assert float(outs[0][1].abs().min()) > 0                   # every slot was written
    assert float(outs[0][3][:, :, 0].abs().min()) > 0 and float(outs[0][4].abs().min()) > 0
    if batch == 1:
        ps.coupling_uniform(0.5 * ps.EL_recoil)
        o = orc.OraclePropagator(problem_of(ps), dt, mode)
        want = o.run(n1 + n2)
        np.testing.assert_allclose(outs[0][1][0, :n1 + n2].cpu().numpy(), want['pops_vals'], rtol=TOL_SCALAR)


@pytest.mark.parametrize('mesh,mode,dt,cpl', [((96, 80), 'real', 1 / 2000, 'uniform'), ((30, 50), 'imag', 1 / 50, 'dense'),
                                               ((1000, 600), 'imag', 1 / 50, 'zero'), ((16, 8), 'real', 1 / 2000, 'none'),
                                               ((750, 4), 'imag', 1 / 50, 'uniform'), ((24, 1458), 'real', 1 / 2000, 'zero')])
@pytest.mark.parametrize('separable', [True, False])
def test_generic_mesh_sizes_against_oracle(mesh, mode, dt, cpl, separable):
    """Even mesh sizes that are not powers of two (prime factors 2, 3, 5, 7) and powers of two below 32 — what the
    reference accepts (pspinor.py:331-332 asserts even sizes) — through the public API: psi_k, populations, energy,
    stand-alone transforms."""
    from spinor_gpe_b200 import tensor_tools as tt
    ps = make_ps(mesh, atom_num=1e4, r_sizes=(16, 16), g_sc={'uu': 1, 'dd': 0.98, 'ud': 1.02})
    if cpl != 'none':
        ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
        if cpl == 'uniform':
            ps.coupling_uniform(1.5 * ps.EL_recoil)
        elif cpl == 'dense':
            ps.coupling_grad(slope=0.3, offset=2.0, axis=1)
        ps.detuning_grad(-3.0)
    ps.rot_coupling = cpl != 'uniform'
    rng = np.random.default_rng(31)
    ps.psik = [p * (1 + 0.05 * (rng.standard_normal(p.shape) + 1j * rng.standard_normal(p.shape))) for p in ps.psik]
    want = orc.OraclePropagator(problem_of(ps), dt, mode).run(4)
    res, prop = (ps.imaginary if mode == 'imag' else ps.real)(dt, 4, 'cuda', separable=separable, unwrap='none')
    assert rel(np.array(res.psik), want['psik']) < TOL_PSI
    assert rel(np.array(res.psi), want['psi']) < TOL_PSI
    np.testing.assert_allclose(res.pops['vals'], want['pops_vals'], rtol=TOL_SCALAR)
    np.testing.assert_allclose(res.eng_final[2:], want['energy'][2:], rtol=TOL_SCALAR)
    psi = [torch.as_tensor(p).cuda() for p in want['psi']]
    back = tt.ifft_2d(tt.fft_2d(psi, ps.space['dr']), ps.space['dr'])
    assert rel(np.array([b.cpu().numpy() for b in back]), want['psi']) < 1e-13
    assert rel(np.array([b.cpu().numpy() for b in tt.fft_2d(psi, ps.space['dr'])]), want['psik']) < 1e-12


def test_single_step_with_custom_operator_tables():
    """single_step(t_step, eng) with caller-built operator tables (tensor_propagator.py:224-271 takes any `eng` dict):
    the reference's own tables reproduce the fused step; modified tables follow the oracle fed the same tables."""
    from spinor_gpe_b200 import TensorPropagator
    from spinor_gpe_b200 import tensor_tools as tt
    ps = make_ps((128, 64), atom_num=1e4, r_sizes=(16, 16))
    ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
    ps.coupling_uniform(1.5 * ps.EL_recoil)
    ps.rot_coupling = False
    prop = TensorPropagator(ps, 1 / 2000, 1, 'cuda', time='real')
    eng = {'kin': tt.evolution_op(prop.dt_out / 2, prop.kin_eng_spin), 'pot': tt.evolution_op(prop.dt_out, prop.pot_eng_spin),
           'coupl': tt.coupling_op(prop.dt_out, prop.coupling / 2, prop.expon)}
    prop.single_step(prop.dt_out, eng)
    got = np.array([p.cpu().numpy() for p in prop.psik])
    prop2 = TensorPropagator(ps, 1 / 2000, 1, 'cuda', time='real')
    prop2.single_step(prop2.dt_out, prop2.eng_out)
    assert rel(got, np.array([p.cpu().numpy() for p in prop2.psik])) < 1e-12
    o = orc.OraclePropagator(problem_of(ps), 1 / 2000, 'real')
    o.single_step(o.ops_out)
    assert rel(got, o.psik.numpy()) < TOL_PSI
    # tables that are no evolution operators of the propagator's grids (an absorbing boundary folded into `pot`)
    damp = torch.as_tensor(np.exp(-0.01 * (ps.space['x_mesh'] ** 2 + ps.space['y_mesh'] ** 2) / 16 ** 2), device='cuda')
    eng['pot'] = [p * damp for p in eng['pot']]
    prop3 = TensorPropagator(ps, 1 / 2000, 1, 'cuda', time='real')
    prop3.single_step(prop3.dt_out, eng)
    o = orc.OraclePropagator(problem_of(ps), 1 / 2000, 'real')
    o.ops_out['pot'] = o.ops_out['pot'] * damp.cpu()
    o.single_step(o.ops_out)
    assert rel(np.array([p.cpu().numpy() for p in prop3.psik]), o.psik.numpy()) < TOL_PSI


def test_complex64_against_oracle():
    ps = make_ps((512, 512))
    ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
    ps.coupling_uniform(1.0 * ps.EL_recoil)
    want = orc.OraclePropagator(problem_of(ps), 1 / 50, 'imag').run(5)
    res, _ = ps.imaginary(1 / 50, 5, 'cuda', precision='c64')
    assert rel(np.array(res.psik), want['psik']) < TOL_PSI_C64
    np.testing.assert_allclose(res.pops['vals'], want['pops_vals'], rtol=1e-5)


def test_energy_of_ground_state_against_oracle():
    """Energy parity where it is pinned (wrapped phase; smooth ground state, rotating frame)."""
    ps = make_ps((256, 256), atom_num=1e2)
    ps.coupling_setup(wavel=790.1e-9, kin_shift=False)
    prob = problem_of(ps)
    o = orc.OraclePropagator(prob, 1 / 50, 'imag')
    want = o.run(30)
    res, prop = ps.imaginary(1 / 50, 30, 'cuda')
    np.testing.assert_allclose(prop.eng_expect(None, unwrap='none'), want['energy'], rtol=TOL_SCALAR)
    # nothing to unwrap in a smooth ground state: the reference's (unwrapped) number is the same
    np.testing.assert_allclose(res.eng_final, want['energy'], rtol=TOL_SCALAR)
    # 'local' unwrapping agrees with the wrapped phase when there is nothing to unwrap
    np.testing.assert_allclose(prop.eng_expect(None, unwrap='local'), want['energy'], rtol=1e-6)


@pytest.mark.parametrize('mode,dt,cpl', [('imag', 1 / 50, 'zero'), ('real', 1 / 2000, 'uniform'), ('real', 1 / 2000, 'dense')])
def test_energy_tracking_every_step(mode, dt, cpl):
    """PSpinor.imaginary/real(track_energy=True): PropResult.eng_history against the oracle's eng_expect after every
    full step (wrapped phase), final state and populations unaffected."""
    ps = make_ps((256, 128), atom_num=1e4, r_sizes=(16, 16), g_sc={'uu': 1, 'dd': 0.98, 'ud': 1.02},
                 pop_frac=(0.6, 0.4), phase_factor=1j)
    ps.coupling_setup(wavel=790.1e-9, kin_shift=(cpl != 'zero'))
    if cpl == 'uniform':
        ps.coupling_uniform(1.5 * ps.EL_recoil)
    elif cpl == 'dense':
        ps.coupling_grad(slope=0.3, offset=2.0, axis=1)
    rng = np.random.default_rng(7)
    ps.psik = [p * (1 + 0.05 * (rng.standard_normal(p.shape) + 1j * rng.standard_normal(p.shape))) for p in ps.psik]
    prob = problem_of(ps)
    n = 5
    o = orc.OraclePropagator(prob, dt, mode)
    want = []
    for _ in range(n):
        o.full_step()
        want.append(orc.energy(prob, o.psik))
    res, prop = (ps.imaginary if mode == 'imag' else ps.real)(dt, n, 'cuda', track_energy=True, unwrap='none')
    assert rel(np.array(res.psik), o.psik.numpy()) < TOL_PSI
    np.testing.assert_allclose(res.eng_history, np.array(want), rtol=TOL_SCALAR)
    np.testing.assert_allclose(res.eng_history[-1], res.eng_final, rtol=TOL_SCALAR)


def test_spectral_kinetic_energy():
    """TensorPropagator.kin_expect_spectral against NumPy on the same state (dense and separable operators)."""
    ps = make_ps((256, 128), atom_num=1e4, r_sizes=(16, 16))
    ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
    ps.coupling_uniform(1.5 * ps.EL_recoil)
    for sep in (True, False):
        res, prop = ps_copy_run(ps, sep)
        want = [(np.asarray(ps.kin_eng_spin[c]) * np.abs(res.psik[c]) ** 2).sum() * ps.space['dv_k'] for c in range(2)]
        np.testing.assert_allclose(prop.kin_expect_spectral(), want, rtol=TOL_SCALAR)
        np.testing.assert_allclose(prop.kin_expect_spectral(res.psik), want, rtol=TOL_SCALAR)


def test_batched_sweep_matches_individual_runs():
    """Config-4 style: several trajectories (coupling x detuning sweep) in one plan == one by one."""
    from spinor_gpe_b200 import _capi
    from spinor_gpe_b200.plan import Plan
    base = make_ps((128, 128), atom_num=1e4, r_sizes=(16, 16), g_sc={'uu': 1, 'dd': 0.995, 'ud': 0.995})
    base.coupling_setup(wavel=804e-9, kin_shift=True)
    base.shift_momentum(scale=0.6, frac=(0.5, 0.5))
    omegas = [c * base.EL_recoil for c in (0.5, 2.0, 5.0)]
    slopes = [-12.0, 0.0, 12.0]
    pots, wants = [], []
    for om, sl in zip(omegas, slopes):
        base.coupling_uniform(om)
        base.detuning_grad(sl)
        pots.append(np.array(base.pot_eng_spin))
        wants.append(orc.OraclePropagator(problem_of(base), 1 / 50, 'imag').run(4))
    B = len(omegas)
    pl = Plan(128, 128, B)
    pl.set_grid(base.space['dr'][0], base.space['dr'][1], base.space['dv_r'], base.space['dv_k'], base.atom_num)
    pl.set_interactions(base.g_sc['uu'], base.g_sc['dd'], base.g_sc['ud'])
    pl.set_kinetic(base.kin_eng_spin[0], base.kin_eng_spin[1])
    pot = np.stack(pots)                                   # (B, 2, ny, nx)
    pl.set_potential(pot[:, 0].copy(), pot[:, 1].copy(), batched=True)
    pl.set_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=np.array(omegas))
    pl.set_time('imag', 1 / 50)
    pl.load(np.stack([np.array(base.psik)] * B))
    pops = torch.zeros((B, 4, 2), dtype=torch.float64, device='cuda')
    pl.full_steps(4, pops)
    out = pl.store().cpu().numpy()
    for b in range(B):
        assert rel(out[b], wants[b]['psik']) < TOL_PSI
        np.testing.assert_allclose(pops[b].cpu().numpy(), wants[b]['pops_vals'], rtol=TOL_SCALAR)


def test_host_buffer_entry_point():
    from spinor_gpe_b200 import _capi
    from spinor_gpe_b200.plan import Plan
    ps = make_ps((256, 256))
    want = orc.OraclePropagator(problem_of(ps), 1 / 50, 'imag').run(3)
    pl = Plan(256, 256, 1)
    pl.set_grid(ps.space['dr'][0], ps.space['dr'][1], ps.space['dv_r'], ps.space['dv_k'], ps.atom_num)
    pl.set_interactions(ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud'])
    pl.set_kinetic(ps.kin_eng_spin[0], ps.kin_eng_spin[1])
    pl.set_potential(ps.pot_eng_spin[0], ps.pot_eng_spin[1], shared=True)
    pl.set_coupling(_capi.SGPE_COUPLING_NONE)
    pl.set_time('imag', 1 / 50)
    out, pops = pl.run_host(np.array(ps.psik)[None], 3)
    assert rel(out[0].numpy(), want['psik']) < TOL_PSI
    np.testing.assert_allclose(pops[0].numpy(), want['pops_vals'], rtol=TOL_SCALAR)


# ------------------------------------------------------------------ properties at full size
@pytest.mark.parametrize('n', [2048, 4096])
def test_full_size_fft_roundtrip_and_parseval(n):
    from spinor_gpe_b200.plan import Plan
    g = torch.Generator(device='cuda').manual_seed(99999)
    psi = torch.randn((1, 2, n, n), dtype=torch.float64, device='cuda', generator=g) \
        + 1j * torch.randn((1, 2, n, n), dtype=torch.float64, device='cuda', generator=g)
    pl = Plan(n, n, 1)
    dx = 16.0 / n
    dvk = (2 * np.pi / 16.0) ** 2
    pl.set_grid(dx, dx, dx * dx, dvk, 100.0)
    psik = pl.fft2d(psi)
    back = pl.fft2d(psik, inverse=True)
    assert float((back - psi).abs().max()) < 1e-12
    n_r = pl.sumsq(psi).sum().item() * dx * dx
    n_k = pl.sumsq(psik).sum().item() * dvk
    assert abs(n_r / n_k - 1) < 1e-12
    # linearity of the transform
    a, b = psi, torch.flip(psi, dims=(-1,))
    lhs = pl.fft2d(2.0 * a + b)
    rhs = 2.0 * psik + pl.fft2d(b)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) < 1e-13


def test_full_size_propagation_properties():
    """2048^2 complex128 (the headline mesh): real-time evolution conserves the atom number and, with the
    interactions off, a harmonic-oscillator eigenstate only acquires a phase; imaginary time renormalises
    to N every step; the stationary state of imaginary time is a fixed point of further steps."""
    ps = make_ps((2048, 2048), atom_num=1e2)
    ps.coupling_setup(wavel=790.1e-9, kin_shift=False)
    res, prop = ps.imaginary(1 / 50, 10, 'cuda')
    np.testing.assert_allclose(res.pops['vals'].sum(axis=1), ps.atom_num, rtol=1e-12)
    np.testing.assert_allclose(res.pops['vals'][:, 0], res.pops['vals'][:, 1], rtol=1e-9)   # symmetric spinor
    res_r, _ = ps.real(1 / 5000, 10, 'cuda')
    np.testing.assert_allclose(res_r.pops['vals'].sum(axis=1), ps.atom_num, rtol=1e-12)

    # non-interacting oscillator ground state exp(-(x^2+y^2)/2): |psi| invariant under real-time steps
    ps0 = make_ps((2048, 2048), atom_num=1e2, g_sc={'uu': 0.0, 'dd': 0.0, 'ud': 0.0})
    x, y = ps0.space['x_mesh'], ps0.space['y_mesh']
    gauss = np.exp(-(x ** 2 + y ** 2) / 2).astype(complex)
    from spinor_gpe_b200 import tensor_tools as tt
    psi, _ = tt.norm([gauss, gauss.copy()], ps0.space['dv_r'], ps0.atom_num)
    ps0.psi, ps0.psik = psi, tt.fft_2d(psi, ps0.space['dr'])
    res0, _ = ps0.real(1 / 1000, 10, 'cuda')
    dens0 = np.abs(psi[0]) ** 2
    assert np.abs(res0.dens[0] - dens0).max() / dens0.max() < 1e-6      # splitting error O(dt^4)


def test_run_sweep_on_gpu():
    """sweep.run_sweep (single process): batched trajectories == one-by-one oracle runs."""
    from spinor_gpe_b200.sweep import detuning_coupling_grid, run_sweep
    base = make_ps((128, 128), atom_num=1e4, r_sizes=(16, 16), g_sc={'uu': 1, 'dd': 0.995, 'ud': 0.995})
    base.coupling_setup(wavel=804e-9, kin_shift=True)
    base.shift_momentum(scale=0.6, frac=(0.5, 0.5))
    trajs = detuning_coupling_grid(base, [0.5 * base.EL_recoil, 5 * base.EL_recoil], [-12.0, 12.0])
    out = run_sweep(base, trajs, 1 / 50, 4, time='imag', device='cuda', batch=3, keep_states=True)
    for i, tr in enumerate(trajs):
        prob = orc.Problem(base.psik, base.kin_eng_spin, tr.pot, np.full_like(base.pot_eng, tr.omega),
                           base.space['dr'], base.space['dv_r'], base.space['dv_k'],
                           [base.g_sc['uu'], base.g_sc['dd'], base.g_sc['ud']], base.atom_num, x=base.space['x'],
                           kL=base.kL_recoil, is_coupling=True, rot_coupling=True)
        want = orc.OraclePropagator(prob, 1 / 50, 'imag').run(4)
        assert rel(out['psik'][i], want['psik']) < TOL_PSI
        np.testing.assert_allclose(out['pops'][i], want['pops_vals'], rtol=TOL_SCALAR)


def _slab_worker(rank, world, port, outdir):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from spinor_gpe_b200.slab import SlabPropagator
        ps = make_ps((512, 256), atom_num=1e4, r_sizes=(16, 16))
        ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
        ps.coupling_uniform(1.5 * ps.EL_recoil)
        ps.rot_coupling = False
        want = orc.OraclePropagator(problem_of(ps), 1 / 2000, 'real').run(4)
        sp = SlabPropagator(ps, 1 / 2000, time='real', device=f'cuda:{rank}')
        pops = torch.zeros((4, 2), dtype=torch.float64, device=f'cuda:{rank}')
        sp.full_steps(4, pops)
        got = sp.gather_psik().cpu().numpy()
        assert rel(got, want['psik']) < TOL_PSI
        np.testing.assert_allclose(pops.cpu().numpy(), want['pops_vals'], rtol=TOL_SCALAR)
    finally:
        dist.destroy_process_group()


def test_slab_two_gpus():
    """Slab-decomposed propagation over 2 GPUs with the NCCL all-to-all (skipped on a 1-GPU box)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_slab_worker, args=(2, port, tempfile.mkdtemp(prefix='sgpe_slab_')), nprocs=2, join=True)


def test_fine_mesh_imaginary_time_factor_tables():
    """k_max^2 dt / 4 >> 709 (exp overflow range): the separable factor tables must stay finite and agree with
    the dense path and the oracle (regression: anchoring the split at a grid corner overflowed at 4096 points)."""
    ps = make_ps((4096, 64), atom_num=1e3, r_sizes=(2, 2))
    ps.coupling_setup(wavel=790.1e-9, kin_shift=False)
    want = orc.OraclePropagator(problem_of(ps), 1 / 50, 'imag').run(2)
    for sep in (True, False):
        res, prop = ps_copy_run(ps, sep)
        assert np.isfinite(np.array(res.psik)).all()
        assert rel(np.array(res.psik), want['psik']) < TOL_PSI
        np.testing.assert_allclose(res.pops['vals'], want['pops_vals'], rtol=TOL_SCALAR)


def ps_copy_run(ps, separable):
    import copy
    q = copy.copy(ps)
    q.psik = [p.copy() for p in ps.psik]
    q.psi = [p.copy() for p in ps.psi]
    return q.imaginary(1 / 50, 2, 'cuda', separable=separable)
