"""sgpe_exp (csrc/kernels.cuh): the branch-free exp() of the per-point imaginary-time factors.  The function's
source text is compiled with g++ as is and compared with long-double exp over the argument range of a propagator
factor; the GPU parity tests then cover it in place."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
#include <cmath>
#include <cstring>
#define SGPE_DI static inline
static inline long long __double_as_longlong(double x) { long long r; std::memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; std::memcpy(&r, &x, 8); return r; }
%s
extern "C" void run(const double* x, double* y, long n) { for (long i = 0; i < n; i++) y[i] = sgpe_exp(x[i]); }
extern "C" void run_small(const double* x, double* y, long n) { for (long i = 0; i < n; i++) y[i] = sgpe_exp_small(x[i]); }
'''


def _build():
    src = open(os.path.join(ROOT, 'spinor_gpe_b200', 'csrc', 'kernels.cuh')).read()
    a = src.index('SGPE_DI double sgpe_exp(double x) {')
    b = src.index('}\n', a) + 2
    a2 = src.index('SGPE_DI double sgpe_exp_small(double x) {')
    b2 = src.index('}\n', a2) + 2
    tmp = tempfile.mkdtemp(prefix='sgpe_exp_')
    with open(os.path.join(tmp, 'e.cpp'), 'w') as f:
        f.write(HARNESS % (src[a:b] + src[a2:b2]))
    so = os.path.join(tmp, 'e.so')
    subprocess.run(['g++', '-O2', '-shared', '-fPIC', '-o', so, os.path.join(tmp, 'e.cpp')], check=True)
    lib = ctypes.CDLL(so)
    lib.run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    lib.run_small.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    return lib


def test_sgpe_exp_within_one_ulp():
    lib = _build()
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(-2, 2, 400000), rng.uniform(-690, 690, 400000), -rng.uniform(0, 1e-3, 200000),
                        np.array([0.0, -0.0, 1.0, -1.0, np.log(2) / 2, -np.log(2) / 2, 1e-300, -1e-300])])
    y = np.empty_like(x)
    lib.run(x.ctypes.data, y.ctypes.data, x.size)
    want = np.exp(x.astype(np.longdouble))
    rel = np.abs((y.astype(np.longdouble) - want) / want).astype(np.float64)
    assert rel.max() < 2.0 ** -52, rel.max()             # < 1 ulp
    assert y[x == 0.0].tolist() == [1.0, 1.0]


def test_sgpe_exp_out_of_range_saturates():
    lib = _build()
    x = np.array([-800.0, -5000.0, 800.0, 5000.0])
    y = np.empty_like(x)
    lib.run(x.ctypes.data, y.ctypes.data, x.size)
    assert np.all(np.isfinite(y)) and np.all(y[:2] < 1e-290) and np.all(y[:2] > 0) and np.all(y[2:] > 1e290)


def test_sgpe_exp_small_within_one_ulp():
    """The degree-9 polynomial the row pass uses for interaction factors with |argument| <= 1/16."""
    lib = _build()
    rng = np.random.default_rng(8)
    x = np.concatenate([rng.uniform(-0.0625, 0.0625, 500000), -rng.uniform(0, 1e-6, 1000), np.array([0.0, 0.0625, -0.0625])])
    y = np.empty_like(x)
    lib.run_small(x.ctypes.data, y.ctypes.data, x.size)
    want = np.exp(x.astype(np.longdouble))
    rel = np.abs((y.astype(np.longdouble) - want) / want).astype(np.float64)
    assert rel.max() < 2.0 ** -52, rel.max()


ATAN_HARNESS = r'''
#include <cmath>
#include <cstring>
#define SGPE_DI static inline
#define __device__
static inline double seed(double d) { double r = 1.0 / d; unsigned long long u; std::memcpy(&u, &r, 8); u &= 0xffffffff00000000ull; std::memcpy(&r, &u, 8); return r; }
#define SGPE_RCP_SEED(d) seed(d)
#define SGPE_RSQRT_SEED(d) seed(std::sqrt(d))
#define SGPE_D2I_RN(x) ((int)nearbyint(x))
%s
extern "C" void run_sqrt(const double* x, double* out, long n) { for (long i = 0; i < n; i++) out[i] = sgpe_sqrt(x[i]); }
extern "C" void run(const double* y, const double* x, double* out, long n) { for (long i = 0; i < n; i++) out[i] = sgpe_atan2(y[i], x[i]); }
'''


def test_sgpe_atan2_within_two_ulp():
    """sgpe_atan2 (the polar store of per-step energy tracking): source text compiled with g++, the 20-bit hardware
    reciprocal seed modelled by truncation, against long-double atan2 over all quadrants, the axes and tiny magnitudes."""
    src = open(os.path.join(ROOT, 'spinor_gpe_b200', 'csrc', 'kernels.cuh')).read()
    a = src.index('SGPE_DI double sgpe_rcp(double d) {')
    b = src.index('}\n', src.index('SGPE_DI double sgpe_atan2(double y, double x) {')) + 2
    assert src.index('SGPE_DI double sgpe_sqrt(double n) {') > a
    tmp = tempfile.mkdtemp(prefix='sgpe_atan_')
    with open(os.path.join(tmp, 'a.cpp'), 'w') as f:
        f.write(ATAN_HARNESS % src[a:b])
    so = os.path.join(tmp, 'a.so')
    subprocess.run(['g++', '-O2', '-shared', '-fPIC', '-o', so, os.path.join(tmp, 'a.cpp')], check=True)
    lib = ctypes.CDLL(so)
    lib.run.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_long]
    rng = np.random.default_rng(11)
    n = 400000
    ang = rng.uniform(-np.pi, np.pi, n)
    mag = 10.0 ** rng.uniform(-200, 200, n)
    x = np.concatenate([mag * np.cos(ang), rng.normal(size=n), [1.0, -1.0, 0.0, 0.0, 1.0, -1.0, 1e-300, 3.0, -3.0]])
    y = np.concatenate([mag * np.sin(ang), rng.normal(size=n), [0.0, 0.0, 1.0, -1.0, 1.0, -1.0, 1e-300, -0.0, -0.0]])
    out = np.empty_like(x)
    lib.run(y.ctypes.data, x.ctypes.data, out.ctypes.data, x.size)
    want = np.arctan2(y.astype(np.longdouble), x.astype(np.longdouble))
    ok = np.hypot(x, y) > 1e-279
    err = np.abs(out[ok].astype(np.longdouble) - want[ok]).astype(np.float64)
    ulp = np.spacing(np.abs(want[ok]).astype(np.float64))
    assert (err / ulp).max() < 2.0, (err / ulp).max()
    assert out[~ok].tolist() == [0.0] * int((~ok).sum())
    assert out[-2] == 0.0 and np.signbit(out[-2]) and out[-1] == -np.pi       # atan2(-0, 3) = -0, atan2(-0, -3) = -pi
    # sgpe_sqrt of the same store
    lib.run_sqrt.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_long]
    v = np.concatenate([10.0 ** rng.uniform(-280, 300, n), rng.uniform(0, 4, n), [0.0, 1.0, 4.0, 1e-300]])
    got = np.empty_like(v)
    lib.run_sqrt(v.ctypes.data, got.ctypes.data, v.size)
    ref = np.sqrt(v.astype(np.longdouble))
    big = v > 1e-289
    assert (np.abs(got[big].astype(np.longdouble) - ref[big]).astype(np.float64) / np.spacing(np.sqrt(v[big]))).max() < 1.0
    assert got[~big].tolist() == [0.0] * int((~big).sum())
