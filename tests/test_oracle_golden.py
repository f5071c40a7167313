"""Pins the CPU oracle (oracle/spinor_oracle.py) to fixtures produced by the unmodified reference
(oracle/gen_golden.py → tests/golden/*.npz).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import spinor_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, '*.npz'))
               if 'tensor_tools' not in p)


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def runs_of(case):
    z = np.load(os.path.join(GOLDEN, case + '.npz'))
    return z, [f'r{i}_' for i in range(int(z['n_runs']))]


def test_cases_present():
    assert set(CASES) >= {'ground_64', 'raman_64x32', 'dgrad_32x64', 'nocoupl_64', 'cgrad_64'}


@pytest.mark.parametrize('case', CASES)
def test_single_and_full_step(case):
    z, runs = runs_of(case)
    for pre in runs:
        prob = orc.Problem.from_golden(z, pre)
        mode, dt = str(z[pre + 'mode']), float(z[pre + 'dt'])
        o = orc.OraclePropagator(prob, dt, mode)
        o.single_step(o.ops_out)
        assert rel_l2(o.psik.numpy(), z[pre + 'psik_single_out']) < 1e-13
        o = orc.OraclePropagator(prob, dt, mode)
        o.single_step(o.ops_in)
        assert rel_l2(o.psik.numpy(), z[pre + 'psik_single_in']) < 1e-13
        o = orc.OraclePropagator(prob, dt, mode)
        o.full_step()
        assert rel_l2(o.psik.numpy(), z[pre + 'psik_full1']) < 1e-13


@pytest.mark.parametrize('case', CASES)
def test_prop_loop(case):
    z, runs = runs_of(case)
    for pre in runs:
        prob = orc.Problem.from_golden(z, pre)
        mode, dt, n = str(z[pre + 'mode']), float(z[pre + 'dt']), int(z[pre + 'n_steps'])
        out = orc.OraclePropagator(prob, dt, mode).run(n, n_samples=2)
        assert rel_l2(out['psik'], z[pre + 'psik_final']) < 1e-12
        assert rel_l2(out['psi'], z[pre + 'psi_final']) < 1e-12
        np.testing.assert_allclose(out['pops_vals'], z[pre + 'pops_vals'], rtol=1e-12)
        np.testing.assert_allclose(out['pops_times'], z[pre + 'pops_times'], rtol=0, atol=0)
        assert rel_l2(out['sampled_psiks'], z[pre + 'sampled_psiks']) < 1e-12
        np.testing.assert_allclose(out['sampled_times'], z[pre + 'sampled_times'], rtol=0, atol=0)
        # energy: pinned for the identity unwrap only (see oracle header)
        np.testing.assert_allclose(out['energy'], z[pre + 'energy_identity_unwrap'], rtol=1e-10)


def test_tensor_tools_vectors():
    z = np.load(os.path.join(GOLDEN, 'tensor_tools_vectors.npz'))
    for tag in 'ab':
        psi = torch.as_tensor(z[f'{tag}_psi'])
        dr = z[f'{tag}_dr']
        assert rel_l2(orc.fft2(psi, dr).numpy(), z[f'{tag}_fft2']) < 1e-14
        assert rel_l2(orc.ifft2(psi, dr).numpy(), z[f'{tag}_ifft2']) < 1e-14
        for ax in (0, 1):
            assert rel_l2(orc.fft1(psi, dr, ax).numpy(), z[f'{tag}_fft1_ax{ax}']) < 1e-14
            assert rel_l2(orc.ifft1(psi, dr, ax).numpy(), z[f'{tag}_ifft1_ax{ax}']) < 1e-14
        pn, dn = orc.normalise(psi, 0.125, 1234.5)
        assert rel_l2(pn.numpy(), z[f'{tag}_norm_psi']) < 1e-15
        assert rel_l2(dn.numpy(), z[f'{tag}_norm_dens']) < 1e-15
        np.testing.assert_allclose(orc.populations(psi, 0.125), z[f'{tag}_pops'], rtol=1e-14)


def test_reference_fft_invariants():
    """The invariants the reference's own tests pin (tests/fft_func_tests.py:27-472): round trip,
    1-D x 1-D == 2-D, Parseval with vol_elem = 4 pi^2 / N, on all-ones grids."""
    eps = 10 * 2.2e-16
    for n in (128, 256):
        ones = torch.ones((2, n, n), dtype=torch.complex128)
        dr = (1.0, 1.0)
        back = orc.ifft2(orc.fft2(ones, dr), dr)
        assert float((back - ones).abs().max()) < eps
        two = orc.fft1(orc.fft1(ones, dr, 0), dr, 1)
        assert float((two - orc.fft2(ones, dr)).abs().max()) < eps * n * n
        n_r = sum(orc.populations(ones, 1.0))
        n_k = sum(orc.populations(orc.fft2(ones, dr), 4 * np.pi ** 2 / (n * n)))
        assert abs(n_r - n_k) < eps * n * n
