"""CPU oracle for the split-step propagator hot path — TEST INFRASTRUCTURE, NOT A PRODUCT PATH.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs
may import this module.  The shipped package (``spinor_gpe_b200``) never imports it and has no CPU
fallback: without the CUDA extension it raises.

What it is: a from-scratch restatement, in torch-CPU float64/complex128 (the reference's own CPU
arithmetic: MKL FFT + ATen element-wise), of the algorithm in

* ``spinor_gpe/pspinor/tensor_propagator.py`` — ``__init__`` :96-149 (sub-steps, operators),
  ``full_step`` :214-222, ``single_step`` :224-271, ``prop_loop`` :173-211, ``eng_expect`` :273-324
* ``spinor_gpe/pspinor/tensor_tools.py`` — ``fft_2d`` :201-228, ``ifft_2d`` :231-258,
  ``fft_1d``/``ifft_1d`` :130-198, ``norm`` :261-311, ``density``/``norm_sq`` :392-441,
  ``calc_pops`` :466-484, ``phase_comp`` :514-539, ``evolution_op`` :546-560, ``coupling_op`` :563-591

State is a single tensor of shape (2, Ny, Nx) instead of the reference's list of two tensors.

Pinning: ``tests/test_oracle_golden.py`` checks every function here against fixtures produced by the
unmodified reference run on CPU in the build container (``oracle/gen_golden.py`` →
``tests/golden/*.npz``).  ψ, populations, FFT helpers and ``norm`` are pinned.  ``energy`` is pinned only
for the identity phase-unwrap (``skimage.restoration.unwrap_phase`` — scikit-image 0.16.2 per the
reference's requirements.txt:25 — is not installed here, so the general unwrap is PARITY UNPINNED;
see DESIGN.md §oracle).
"""
import math

import numpy as np
import torch

C128 = torch.complex128
F64 = torch.float64
MAGIC_GAMMA = 1.0 / (2.0 + 2.0 ** (1.0 / 3.0))      # tensor_propagator.py:101


# ----------------------------------------------------------------------------- transforms
def _scale2(dr):
    return float(dr[0]) * float(dr[1]) / (2.0 * math.pi)          # tensor_tools.py:218, 248


def fft2(psi, dr):
    """tensor_tools.py:218-226 — fftn, × dx·dy/2π, fftshift (over the last two dims)."""
    return torch.fft.fftshift(torch.fft.fftn(psi, dim=(-2, -1)) * _scale2(dr), dim=(-2, -1))


def ifft2(psik, dr):
    """tensor_tools.py:248-256 — ifftshift, ifftn, ÷ dx·dy/2π."""
    return torch.fft.ifftn(torch.fft.ifftshift(psik, dim=(-2, -1)), dim=(-2, -1)) / _scale2(dr)


def fft1(psi, dr, axis):
    """tensor_tools.py:150-162 — ``axis`` 0 is x (array dim -1), 1 is y (array dim -2)."""
    dim = -1 if axis == 0 else -2
    s = float(dr[axis]) / math.sqrt(2.0 * math.pi)
    return torch.fft.fftshift(torch.fft.fft(psi, dim=dim) * s, dim=dim)


def ifft1(psik, dr, axis):
    """tensor_tools.py:187-196."""
    dim = -1 if axis == 0 else -2
    s = float(dr[axis]) / math.sqrt(2.0 * math.pi)
    return torch.fft.ifft(torch.fft.ifftshift(psik, dim=dim), dim=dim) / s


# ----------------------------------------------------------------------------- reductions
def density(psi):
    """tensor_tools.py:437 — abs(ψ)**2 per component."""
    return torch.abs(psi) ** 2


def normalise(psi, vol, atom_num):
    """tensor_tools.py:289-305 — returns (ψ/√nf, n/nf) with nf = Σ(n0+n1)·vol/N."""
    dens = density(psi)
    nf = float(torch.sum(dens[0] + dens[1]) * vol / atom_num)
    return psi / math.sqrt(nf), dens / nf


def populations(psi, vol):
    """tensor_tools.py:481-482."""
    dens = density(psi)
    return [float(dens[0].sum() * vol), float(dens[1].sum() * vol)]


# ----------------------------------------------------------------------------- operators
def evolution(t, energy):
    """tensor_tools.py:556-558 — exp(-i·E·t); ``t`` may be complex (imaginary time)."""
    return torch.exp(-1.0j * energy * t)


def coupling_matrix(t, coupling, expon):
    """tensor_tools.py:586-590 — 2×2 list of (Ny,Nx) complex tensors."""
    arg = coupling * t / 2
    c = torch.cos(arg)
    s = -1.0j * torch.sin(arg)
    return [[c, s * torch.exp(-1.0j * expon)], [s * torch.exp(1.0j * expon), c]]


class Problem:
    """The inputs TensorPropagator.__init__ takes from a PSpinor (tensor_propagator.py:93-129)."""

    def __init__(self, psik, kin, pot, coupling, dr, dv_r, dv_k, g, atom_num, x=None, kL=1.0,
                 is_coupling=False, rot_coupling=True):
        self.psik = torch.as_tensor(np.asarray(psik), dtype=C128).clone()
        self.kin = torch.as_tensor(np.asarray(kin), dtype=F64)
        self.pot = torch.as_tensor(np.asarray(pot), dtype=F64)
        self.coupling = torch.as_tensor(np.asarray(coupling), dtype=F64)
        self.dr = (float(dr[0]), float(dr[1]))
        self.dv_r = float(dv_r)
        self.dv_k = float(dv_k)
        self.g_uu, self.g_dd, self.g_ud = (float(v) for v in g)
        self.atom_num = float(atom_num)
        self.kL = float(kL)
        self.is_coupling = bool(is_coupling)
        ny, nx = self.psik.shape[-2:]
        if rot_coupling or x is None:
            self.expon = torch.zeros((), dtype=F64)                       # :126-127
        else:
            xs = torch.as_tensor(np.asarray(x), dtype=F64)
            self.expon = (2 * self.kL * xs).reshape(1, nx).expand(ny, nx)  # :129 (x_mesh varies along dim -1)

    @classmethod
    def from_golden(cls, z, pre):
        g = lambda k: z[pre + 'in_' + k]  # noqa: E731
        return cls(g('psik'), g('kin'), g('pot'), g('coupling'), g('dr'), float(g('dv_r')),
                   float(g('dv_k')), g('g'), float(g('atom_num')), x=g('x'), kL=float(g('kL')),
                   is_coupling=bool(g('is_coupling')), rot_coupling=bool(g('rot_coupling')))


class OraclePropagator:
    """Restatement of TensorPropagator (operators precomputed as in :138-149)."""

    def __init__(self, prob, t_step, time='imag'):
        self.p = prob
        self.psik = prob.psik.clone()
        self.t_step = -1.0j * t_step if time == 'imag' else t_step             # :96-99
        self.dt_out = self.t_step * MAGIC_GAMMA                                # :102
        self.dt_in = self.t_step * (1 - 2 * MAGIC_GAMMA)                       # :103
        self.ops_out = self._ops(self.dt_out, outer=True)
        self.ops_in = self._ops(self.dt_in, outer=False)

    def _ops(self, dt, outer):
        p = self.p
        if outer:      # :142-143  coupling_op(dt_out, coupling / 2, expon)
            cpl = coupling_matrix(dt, p.coupling / 2, p.expon)
        else:          # :148-149  coupling_op(dt_in / 2, coupling, expon)
            cpl = coupling_matrix(dt / 2, p.coupling, p.expon)
        return dict(dt=dt, kin=evolution(dt / 2, p.kin), pot=evolution(dt, p.pot), coupl=cpl)

    def _apply_coupling(self, psi, cpl):
        """:253-254 — ψ'_r = Σ_c C[r][c]·ψ_c."""
        return torch.stack([cpl[0][0] * psi[0] + cpl[0][1] * psi[1],
                            cpl[1][0] * psi[0] + cpl[1][1] * psi[1]])

    def single_step(self, ops):
        """:242-271."""
        p = self.p
        psik = ops['kin'] * self.psik
        psi = ifft2(psik, p.dr)
        psi, dens = normalise(psi, p.dv_r, p.atom_num)
        e_int = torch.stack([p.g_uu * dens[0] + p.g_ud * dens[1],
                             p.g_dd * dens[1] + p.g_ud * dens[0]])
        int_op = evolution(ops['dt'] / 2, e_int)
        psi = int_op * psi
        if p.is_coupling:
            psi = self._apply_coupling(psi, ops['coupl'])
        psi = ops['pot'] * psi
        if p.is_coupling:
            psi = self._apply_coupling(psi, ops['coupl'])
        psi = int_op * psi                     # same operator: the density is not recomputed (:262-267)
        psik = ops['kin'] * fft2(psi, p.dr)
        self.psik, _ = normalise(psik, p.dv_k, p.atom_num)

    def full_step(self):
        """:220-222."""
        self.single_step(self.ops_out)
        self.single_step(self.ops_in)
        self.single_step(self.ops_out)

    def run(self, n_steps, n_samples=0):
        """prop_loop :173-211 (without file output).  Returns dict(psik, psi, pops_vals, pops_times,
        sampled_psiks, sampled_times, energy)."""
        vals = np.empty((n_steps, 2))
        times = np.linspace(0, n_steps * abs(self.t_step), n_steps)
        samples, rate = [], (n_steps / n_samples if n_samples else None)
        for i in range(n_steps):
            if n_samples and i % rate == 0:
                samples.append(self.psik.numpy().copy())          # BEFORE the step (:186-189)
            self.full_step()
            vals[i] = populations(self.psik, self.p.dv_k)
        out = dict(psik=self.psik.numpy().copy(), psi=ifft2(self.psik, self.p.dr).numpy(),
                   pops_vals=vals, pops_times=times, energy=energy(self.p, self.psik))
        if n_samples:
            out['sampled_psiks'] = np.array(samples)
            out['sampled_times'] = np.linspace(0, n_steps * abs(self.t_step), n_samples)
        return out


# ----------------------------------------------------------------------------- energy
def _masked_phase(psi, dens, unwrap=None):
    """tensor_tools.py:528-539 — angle → unwrap → zero where n < 1e-6·max(n)."""
    ang = np.angle(psi)
    if unwrap is not None:
        ang = unwrap(ang)
    ang = np.array(ang, copy=True)
    ang[dens < dens.max() * 1e-6] = 0
    return ang


def energy(prob, psik, unwrap=None):
    """tensor_propagator.py:298-324 — returns [E_tot, E_kin, E_pot, E_int] as raw grid sums.

    ``unwrap``: callable replacing skimage.restoration.unwrap_phase; None = identity (the only
    variant pinned by the golden fixtures).  Quirks kept: np.gradient's first returned array is
    d/d(axis 0) with spacing dr[0] and is what the reference calls "x"; no volume element; the
    interaction term carries no 1/2; the coupling term ignores the Raman phase.
    """
    # eng_expect works on NumPy arrays (:296-300), i.e. with numpy.fft, not torch.fft.  Mirrored here
    # because the wrapped phase is ill-conditioned wherever ψ is (nearly) real and negative.
    pk = np.asarray(psik.numpy() if isinstance(psik, torch.Tensor) else psik)
    psi = np.array([np.fft.ifftn(np.fft.ifftshift(c)) / _scale2(prob.dr) for c in pk])
    dr = np.array(prob.dr)
    dens = np.abs(psi) ** 2
    root = np.sqrt(dens)
    kin = 0.0
    for c in range(2):
        ph = _masked_phase(psi[c], dens[c], unwrap)
        g0, g1 = np.gradient(ph, *dr)
        r0, r1 = np.gradient(root[c], *dr)
        kin = kin + (r0 ** 2 + r1 ** 2) + dens[c] * (g0 ** 2 + g1 ** 2) \
            + dens[c] * g0 * (2 * prob.kL * prob.is_coupling)
    kin = kin / 2
    pot = dens[0] * prob.pot[0].numpy() + dens[1] * prob.pot[1].numpy()
    inter = prob.g_uu * dens[0] ** 2 + prob.g_dd * dens[1] ** 2 + prob.g_ud * dens[0] * dens[1]
    coupl = (np.conj(psi[0]) * psi[1] + np.conj(psi[1]) * psi[0]) * prob.coupling.numpy() / 2
    total = float(np.real((kin + pot + inter + coupl).sum()))
    return [total, float(np.real(kin).sum()), float(np.real(pot).sum()), float(np.real(inter).sum())]
