"""Physical constants (SI) used to scale the simulation — same values as the reference's
spinor_gpe/constants.py:14-27 so that a_x, a_sc, g_sc and kL_recoil agree bit for bit."""
import scipy.constants as _sc

h = _sc.h
hbar = _sc.hbar
c = _sc.c
eps0 = _sc.epsilon_0
a0 = _sc.physical_constants['Bohr radius'][0]
e = _sc.elementary_charge

# Rubidium-87 (mass and scattering length as the reference rounds them)
Rb87 = {
    'm': 87 * 1.66e-27,          # kg
    'D2': 780.1e-9,              # m
    'a_sc': 100.4 * a0,          # m
}
Rb87['g'] = 4 * _sc.pi * hbar ** 2 * Rb87['a_sc'] / Rb87['m']
Rb87['k_r'] = 2 * _sc.pi / Rb87['D2']
Rb87['E_r'] = (h / (2 * Rb87['m'])) * (1 / Rb87['D2']) ** 2
