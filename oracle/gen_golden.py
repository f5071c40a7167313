"""Generate golden fixtures by running the UNMODIFIED reference (read-only, /root/reference).

TEST INFRASTRUCTURE ONLY.  This script can only run in the build container (the reference tree is
not shipped to the GPU box).  It imports ``spinor_gpe`` from ``/root/reference`` with ``matplotlib``
and ``skimage`` stubbed (neither is installed; the stubs are never reached by ψ / population
arithmetic — ``skimage.restoration.unwrap_phase`` is reached only by ``eng_expect`` and is stubbed by
the identity, so the stored energies are pinned for the *identity-unwrap* restatement only, see
DESIGN.md), drives ``PSpinor`` / ``TensorPropagator`` on CPU for a set of small cases and stores
inputs + outputs as ``tests/golden/<case>.npz``.

Usage:  python oracle/gen_golden.py [--out tests/golden]
"""
import argparse
import os
import shutil
import sys
import tempfile
import types

import numpy as np

REF_ROOT = '/root/reference'


def _install_stubs():
    """Stub the two imports the reference needs but this image lacks."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Any:
        def __getattr__(self, _):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    rest = mod('skimage.restoration', unwrap_phase=lambda a, *args, **kw: np.array(a, copy=True))
    mod('skimage', restoration=rest)
    def lazy(name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Any()

    plt = mod('matplotlib.pyplot', __getattr__=lazy)
    gs = mod('matplotlib.gridspec', __getattr__=lazy)
    an = mod('matplotlib.animation', __getattr__=lazy)
    mod('matplotlib', pyplot=plt, gridspec=gs, animation=an, __getattr__=lazy)


def import_reference():
    import torch  # noqa: F401  (import before the stubs so torch never sees them)
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import torch
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    from spinor_gpe.pspinor import pspinor as ref_spin
    from spinor_gpe.pspinor import tensor_propagator as ref_tprop
    from spinor_gpe.pspinor import tensor_tools as ref_tt
    ref_tprop.tqdm = lambda it, *a, **k: it      # silence the progress bar
    return ref_spin, ref_tprop, ref_tt


W0 = 2 * np.pi * 50

# Each case: constructor kwargs, the setup calls (name, kwargs) applied in order, then runs.
CASES = {
    # BASELINE config 1 / example 1 (ground state, Omega = 0 but is_coupling=True), shrunk to 64x64
    'ground_64': dict(
        ctor=dict(atom_num=1e2, omeg={'x': W0, 'y': W0, 'z': 40 * W0},
                  g_sc={'uu': 1, 'dd': 1, 'ud': 1.04}, pop_frac=(0.5, 0.5), r_sizes=(8, 8),
                  mesh_points=(64, 64)),
        setup=[('coupling_setup', dict(wavel=790.1e-9, kin_shift=False))],
        attrs=dict(rand_seed=99999),
        runs=[('imag', 1 / 50, 20)]),
    # BASELINE config 2 / example 3 (Raman, non-rotated frame, momentum kick), shrunk; non-square
    'raman_64x32': dict(
        ctor=dict(atom_num=1e4, omeg={'x': W0, 'y': W0, 'z': 40 * W0},
                  g_sc={'uu': 1, 'dd': 1, 'ud': 0.0}, pop_frac=(1.0, 0.0), r_sizes=(16, 16),
                  mesh_points=(64, 32)),
        setup=[('coupling_setup', dict(wavel=790.1e-9, kin_shift=True)),
               ('shift_momentum', dict(scale=1.0, frac=(0, 1.0)))],
        attrs=dict(rand_seed=99999, rot_coupling=False),
        runs=[('imag', 1 / 50, 6),
              ('call', 'coupling_uniform', ('EL', 1.0)),
              ('real', 1 / 5000, 12)]),
    # BASELINE config 4 / example 4 (detuning gradient, rotated frame), shrunk; non-square other way
    'dgrad_32x64': dict(
        ctor=dict(atom_num=1e4, omeg={'x': W0, 'y': W0, 'z': 40 * W0},
                  g_sc={'uu': 1, 'dd': 0.995, 'ud': 0.995}, pop_frac=(0.5, 0.5), r_sizes=(16, 16),
                  mesh_points=(32, 64)),
        setup=[('coupling_setup', dict(wavel=804e-9, kin_shift=True)),
               ('shift_momentum', dict(scale=0.6, frac=(0.5, 0.5))),
               ('coupling_uniform', ('EL', 5.0)),
               ('detuning_grad', dict(slope=-12))],
        attrs=dict(rand_seed=99999, rot_coupling=True),
        runs=[('imag', 1 / 50, 10)]),
    # No coupling at all (is_coupling False), anisotropic trap, real time (example 2 flavour)
    'nocoupl_64': dict(
        ctor=dict(atom_num=1e3, omeg={'x': W0, 'y': 2 * W0, 'z': 40 * W0},
                  g_sc=None, pop_frac=(0.7, 0.3), r_sizes=(12, 12), mesh_points=(64, 64),
                  phase_factor=1j),
        setup=[],
        attrs=dict(),
        runs=[('real', 1 / 400, 8)]),
    # Dense, spatially varying coupling (coupling_grad) + uniform detuning, non-rotated frame, real
    'cgrad_64': dict(
        ctor=dict(atom_num=5e3, omeg={'x': W0, 'y': W0, 'z': 40 * W0},
                  g_sc={'uu': 1, 'dd': 0.98, 'ud': 1.01}, pop_frac=(0.5, 0.5), r_sizes=(12, 12),
                  mesh_points=(64, 64)),
        setup=[('coupling_setup', dict(wavel=790.1e-9, scale=0.5, kin_shift=True)),
               ('coupling_grad', dict(slope=0.4, offset=3.0, axis=0)),
               ('detuning_uniform', dict(value=0.8))],
        attrs=dict(rot_coupling=False),
        runs=[('real', 1 / 1000, 8), ('imag', 1 / 100, 4)]),
    # Vortices imprinted on both components (seed_vortices, pspinor.py:680-745), no Raman coupling, real time:
    # phase singularities, zeros of the density inside the cloud
    'vortex_64': dict(
        ctor=dict(atom_num=2e3, omeg={'x': W0, 'y': W0, 'z': 40 * W0},
                  g_sc={'uu': 1, 'dd': 0.97, 'ud': 0.9}, pop_frac=(0.5, 0.5), r_sizes=(10, 10),
                  mesh_points=(64, 64)),
        setup=[('seed_vortices', dict(positions=[[1.5, 0.5], [-2.0, 1.0], [0.3, -2.2]], windings=[1, -1, 2]))],
        attrs=dict(),
        runs=[('real', 1 / 400, 8)]),
    # Time of flight (example 2): relax in an anisotropic trap, switch the trap off, expand in real time
    'tof_32x64': dict(
        ctor=dict(atom_num=1e4, omeg={'x': W0, 'y': 4 * W0, 'z': 40 * W0},
                  g_sc={'uu': 1, 'dd': 1, 'ud': 0.5}, pop_frac=(0.5, 0.5), r_sizes=(16, 16),
                  mesh_points=(32, 64)),
        setup=[('coupling_setup', dict(wavel=790.1e-9))],
        attrs=dict(rand_seed=99999),
        runs=[('imag', 1 / 50, 6),
              ('zero', 'pot_eng'),
              ('real', 1 / 500, 8)]),
}


def _resolve(ps, arg):
    if isinstance(arg, tuple) and len(arg) == 2 and arg[0] == 'EL':
        return (arg[1] * ps.EL_recoil,)
    return arg


def snapshot_inputs(ps):
    """Everything TensorPropagator.__init__ reads from the PSpinor (tensor_propagator.py:93-126)."""
    return dict(
        psik=np.array(ps.psik), kin=np.array(ps.kin_eng_spin), pot=np.array(ps.pot_eng_spin),
        coupling=np.array(ps.coupling), dr=np.array(ps.space['dr']), dk=np.array(ps.space['dk']),
        dv_r=float(ps.space['dv_r']), dv_k=float(ps.space['dv_k']),
        x=np.array(ps.space['x']), y=np.array(ps.space['y']),
        g=np.array([ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud']], dtype=np.float64),
        atom_num=float(ps.atom_num), kL=float(ps.kL_recoil), is_coupling=bool(ps.is_coupling),
        rot_coupling=bool(ps.rot_coupling))


def run_case(name, spec, ref_spin, ref_tprop, ref_tt, out_dir):
    tmp = tempfile.mkdtemp(prefix='sgpe_golden_') + '/'
    shutil.rmtree(tmp)
    ctor = dict(spec['ctor'])
    ps = ref_spin.PSpinor(tmp, overwrite=True, **ctor)
    for meth, arg in spec['setup']:
        arg = _resolve(ps, arg)
        if isinstance(arg, dict):
            getattr(ps, meth)(**arg)
        else:
            getattr(ps, meth)(*arg)
    for k, v in spec['attrs'].items():
        setattr(ps, k, v)

    out = {}
    # PSpinor-level scalars (for the PSpinor clone test)
    out['setup_scalars'] = np.array([ps.a_x, ps.a_sc, ps.chem_pot, ps.rad_tf, ps.time_scale,
                                     ps.kL_recoil, ps.EL_recoil], dtype=np.float64)
    out['setup_psi'] = np.array(ps.psi)
    run_idx = 0
    for run in spec['runs']:
        if run[0] == 'call':
            arg = _resolve(ps, run[2])
            getattr(ps, run[1])(*arg)
            continue
        if run[0] == 'zero':                       # e.g. ps.pot_eng = zeros: the trap is switched off
            setattr(ps, run[1], np.zeros_like(getattr(ps, run[1])))
            continue
        mode, dt, n = run
        pre = f'r{run_idx}_'
        for k, v in snapshot_inputs(ps).items():
            out[pre + 'in_' + k] = v
        out[pre + 'mode'] = np.array(mode)
        out[pre + 'dt'] = np.array(dt)
        out[pre + 'n_steps'] = np.array(n)
        # (1) one outer single_step, (2) one full_step — on a scratch propagator
        prop = ref_tprop.TensorPropagator(ps, dt, n, 'cpu', time=mode)
        prop.single_step(prop.dt_out, prop.eng_out)
        out[pre + 'psik_single_out'] = np.array(ref_tt.to_numpy(prop.psik))
        prop = ref_tprop.TensorPropagator(ps, dt, n, 'cpu', time=mode)
        prop.single_step(prop.dt_in, prop.eng_in)
        out[pre + 'psik_single_in'] = np.array(ref_tt.to_numpy(prop.psik))
        prop = ref_tprop.TensorPropagator(ps, dt, n, 'cpu', time=mode)
        prop.full_step()
        out[pre + 'psik_full1'] = np.array(ref_tt.to_numpy(prop.psik))
        # (3) the public entry point: n steps, with sampling
        n_samples = 2
        fn = ps.imaginary if mode == 'imag' else ps.real
        res, prop = fn(dt, n, 'cpu', is_sampling=True, n_samples=n_samples)
        out[pre + 'psik_final'] = np.array(res.psik)
        out[pre + 'psi_final'] = np.array(res.psi)
        out[pre + 'pops_vals'] = np.array(res.pops['vals'])
        out[pre + 'pops_times'] = np.array(res.pops['times'])
        out[pre + 'energy_identity_unwrap'] = np.array(res.eng_final, dtype=np.float64)
        with np.load(res.sampled_path) as smp:
            out[pre + 'sampled_psiks'] = np.array(smp['psiks'])
            out[pre + 'sampled_times'] = np.array(smp['times'])
        out[pre + 'sampled_name'] = np.array(os.path.basename(res.sampled_path))
        run_idx += 1
    out['n_runs'] = np.array(run_idx)
    shutil.rmtree(tmp, ignore_errors=True)
    path = os.path.join(out_dir, name + '.npz')
    np.savez_compressed(path, **out)
    print(f'{name}: {run_idx} runs -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)')


def fft_vectors(ref_tt, out_dir):
    """Random-input vectors for the transform/normalisation helpers (tensor_tools.py:130-311, 466-484)."""
    rng = np.random.default_rng(99999)
    out = {}
    for tag, (ny, nx) in {'a': (32, 64), 'b': (64, 32)}.items():
        psi = [rng.standard_normal((ny, nx)) + 1j * rng.standard_normal((ny, nx)) for _ in range(2)]
        dr = np.array([0.25, 0.5])
        out[f'{tag}_psi'] = np.array(psi)
        out[f'{tag}_dr'] = dr
        out[f'{tag}_fft2'] = np.array(ref_tt.fft_2d(psi, dr))
        out[f'{tag}_ifft2'] = np.array(ref_tt.ifft_2d(psi, dr))
        for ax in (0, 1):
            out[f'{tag}_fft1_ax{ax}'] = np.array(ref_tt.fft_1d(psi, dr, axis=ax))
            out[f'{tag}_ifft1_ax{ax}'] = np.array(ref_tt.ifft_1d(psi, dr, axis=ax))
        pn, dn = ref_tt.norm(psi, 0.125, 1234.5)
        out[f'{tag}_norm_psi'] = np.array(pn)
        out[f'{tag}_norm_dens'] = np.array(dn)
        out[f'{tag}_pops'] = np.array(ref_tt.calc_pops(psi, 0.125))
    path = os.path.join(out_dir, 'tensor_tools_vectors.npz')
    np.savez_compressed(path, **out)
    print(f'tensor_tools vectors -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden'))
    ap.add_argument('--only', default=None)
    args = ap.parse_args()
    out_dir = os.path.abspath(args.out)
    os.makedirs(out_dir, exist_ok=True)
    ref_spin, ref_tprop, ref_tt = import_reference()
    for name, spec in CASES.items():
        if args.only and name != args.only:
            continue
        run_case(name, spec, ref_spin, ref_tprop, ref_tt, out_dir)
    if not args.only:
        fft_vectors(ref_tt, out_dir)


if __name__ == '__main__':
    main()
