"""PSpinor clone vs the reference's set-up (fixtures from oracle/gen_golden.py).  CPU only."""
import os
import tempfile

import numpy as np
import pytest

from oracle.gen_golden import CASES, _resolve, snapshot_inputs
from spinor_gpe_b200 import PSpinor
from spinor_gpe_b200 import tensor_tools as tt

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def build_case(spec):
    tmp = os.path.join(tempfile.mkdtemp(prefix='sgpe_t_'), 'run') + os.sep
    ps = PSpinor(tmp, overwrite=True, **spec['ctor'])
    for meth, arg in spec['setup']:
        arg = _resolve(ps, arg)
        getattr(ps, meth)(**arg) if isinstance(arg, dict) else getattr(ps, meth)(*arg)
    for k, v in spec['attrs'].items():
        setattr(ps, k, v)
    return ps


@pytest.mark.parametrize('name', sorted(CASES))
def test_setup_matches_reference(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    ps = build_case(CASES[name])
    sc = np.array([ps.a_x, ps.a_sc, ps.chem_pot, ps.rad_tf, ps.time_scale, ps.kL_recoil, ps.EL_recoil])
    np.testing.assert_allclose(sc, z['setup_scalars'], rtol=1e-15)
    for k, v in snapshot_inputs(ps).items():
        np.testing.assert_allclose(np.asarray(v, dtype=complex), np.asarray(z['r0_in_' + k], dtype=complex),
                                   rtol=1e-13, atol=1e-13, err_msg=k)
    assert os.path.exists(ps.paths['trial'] + 'tf_wf-' + ps.paths['folder'] + '.npz')


def test_data_path_rules():
    base = tempfile.mkdtemp(prefix='sgpe_t_')
    p = os.path.join(base, 'proj', 'trial') + os.sep
    ps = PSpinor(p, mesh_points=(32, 32))
    assert ps.paths['folder'] == 'trial' and os.path.isdir(ps.paths['code'])
    with pytest.raises(FileExistsError):
        PSpinor(p, mesh_points=(32, 32))
    PSpinor(p, mesh_points=(32, 32), overwrite=True)
    with pytest.raises(AssertionError):
        PSpinor(os.path.join(base, 'odd') + os.sep, mesh_points=(33, 32))
    with pytest.raises(AssertionError):
        PSpinor(os.path.join(base, 'pf') + os.sep, mesh_points=(32, 32), pop_frac=(0.5, 0.6))


def test_numpy_tensor_tools_vectors():
    z = np.load(os.path.join(GOLDEN, 'tensor_tools_vectors.npz'))
    for tag in 'ab':
        psi = list(z[f'{tag}_psi'])
        dr = z[f'{tag}_dr']
        np.testing.assert_allclose(np.array(tt.fft_2d(psi, dr)), z[f'{tag}_fft2'], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(np.array(tt.ifft_2d(psi, dr)), z[f'{tag}_ifft2'], rtol=1e-12, atol=1e-12)
        for ax in (0, 1):
            np.testing.assert_allclose(np.array(tt.fft_1d(psi, dr, ax)), z[f'{tag}_fft1_ax{ax}'], rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(np.array(tt.ifft_1d(psi, dr, ax)), z[f'{tag}_ifft1_ax{ax}'], rtol=1e-12, atol=1e-12)
        pn, dn = tt.norm(psi, 0.125, 1234.5)
        np.testing.assert_allclose(np.array(pn), z[f'{tag}_norm_psi'], rtol=1e-14)
        np.testing.assert_allclose(np.array(dn), z[f'{tag}_norm_dens'], rtol=1e-14)
        np.testing.assert_allclose(tt.calc_pops(psi, 0.125), z[f'{tag}_pops'], rtol=1e-14)


def test_next_available_path_and_propresult():
    from spinor_gpe_b200.tensor_propagator import next_available_path
    from spinor_gpe_b200 import PropResult
    d = tempfile.mkdtemp(prefix='sgpe_t_')
    stem = os.path.join(d, 'psik_sampled')
    assert next_available_path(stem, 'run', '.npz') == stem + '1-run.npz'
    open(stem + '1-run.npz', 'w').close()
    assert next_available_path(stem, 'run', '.npz') == stem + '2-run.npz'
    rng = np.random.default_rng(0)
    psi = [rng.standard_normal((32, 32)) + 1j * rng.standard_normal((32, 32)) for _ in range(2)]
    res = PropResult(psi, tt.fft_2d(psi), [0, 0, 0, 0], {'times': np.zeros(1), 'vals': np.zeros((1, 2))})
    assert 0 <= res.calc_separation() <= 1
    assert res.rebin(res.dens, (16, 16))[0].shape == (16, 16)
    with pytest.raises(NotImplementedError):
        res.plot_eng()
