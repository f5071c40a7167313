// unwrap.cuh — device side of the two-dimensional phase unwrapping the energy expectation needs.
//
// The reference unwraps the phase of each real-space component with skimage.restoration.unwrap_phase
// (tensor_tools.py:531, from TensorPropagator.eng_expect, tensor_propagator.py:304): the algorithm of Herraez,
// Burton, Lalor and Gdeisat, Appl. Opt. 41, 7437 (2002) — "sorting by reliability following a non-continuous path".
// Its data-parallel parts run here, one thread per pixel:
//   unwrap_angle_pass  : phi = atan2(Im psi, Re psi)                                     (np.angle, :529)
//   unwrap_reliab_pass : reliability = H^2 + V^2 + D1^2 + D2^2 of the wrapped second differences; border pixels get
//                        9999999 ("least reliable"; scikit-image adds rand() there, which only decides how the four
//                        corner pixels attach — every other border pixel joins through its interior neighbour)
//   unwrap_edge_pass   : one sort key (the reliability sum of the two pixels, as order-preserving integer bits) and
//                        one payload (edge id and wrap count) per horizontal / vertical pixel pair
//   (radix sort of the edges by key: cub, in sgpe_api.cu)
//   unwrap_apply_pass  : phi + 2 pi * increment, optionally zeroed where the density is below 1e-6 of its maximum
//                        (tensor_tools.py:538)
// The region merging between sort and apply is inherently sequential (every merge depends on all the earlier ones)
// and runs on the host: an offset-carrying union-find in sgpe_api.cu.
//
// The arithmetic that decides the ORDER of the edges is written with explicit round-to-nearest multiplies and adds
// (no FMA contraction) so that keys are bit-identical to a plain C evaluation of the same expressions.
#pragma once

#include "fft_core.cuh"

namespace sgpe {

constexpr double kUnwrapPi = 3.141592653589793;
constexpr double kUnwrapTwoPi = 6.283185307179586;
constexpr double kUnwrapBorder = 9999999.0;

SGPE_DI double unwrap_wrap(double d) {
    if (d > kUnwrapPi) return d - kUnwrapTwoPi;
    if (d < -kUnwrapPi) return d + kUnwrapTwoPi;
    return d;
}
// multiples of 2 pi the right / lower pixel of an edge needs relative to the left / upper one, negated
SGPE_DI int unwrap_find_wrap(double left, double right) {
    const double d = left - right;
    if (d > kUnwrapPi) return -1;
    if (d < -kUnwrapPi) return 1;
    return 0;
}
SGPE_DI double unwrap_second_diff_sq(double before, double centre, double after) {
    const double s = unwrap_wrap(before - centre) - unwrap_wrap(centre - after);
    return __dmul_rn(s, s);
}

template <typename T>
__global__ void __launch_bounds__(256) unwrap_angle_pass(const typename cx_of<T>::type* psi, long long total,
                                                         double* phi) {
    typedef typename cx_of<T>::type C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const C z = psi[i];
        phi[i] = atan2((double)z.y, (double)z.x);
    }
}

// per-plane maximum of |psi|^2 as the bits of a non-negative double (integer order == floating-point order);
// grid (blocks, planes), maxbits zeroed by the caller
template <typename T>
__global__ void __launch_bounds__(256) unwrap_maxdens_pass(const typename cx_of<T>::type* psi, long long plane,
                                                           unsigned long long* maxbits) {
    typedef typename cx_of<T>::type C;
    SGPE_DYN_SMEM(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw);
    const C* p = psi + (long long)blockIdx.y * plane;
    double mx = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < plane;
         i += (long long)gridDim.x * blockDim.x) {
        const C z = p[i];
        const double d = (double)z.x * z.x + (double)z.y * z.y;
        mx = d > mx ? d : mx;
    }
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] = red[threadIdx.x] > red[threadIdx.x + s] ? red[threadIdx.x] : red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicMax(&maxbits[blockIdx.y], (unsigned long long)__double_as_longlong(red[0]));
}

__global__ void __launch_bounds__(256) unwrap_reliab_pass(const double* phi, int nx, int ny, double* rel) {
    const long long plane = (long long)nx * ny;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < plane;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / nx), j = (int)(idx - (long long)i * nx);
        double r = kUnwrapBorder;
        if (i > 0 && i < ny - 1 && j > 0 && j < nx - 1) {
            const double* w = phi + idx;
            const double c = w[0];
            const double h = unwrap_second_diff_sq(w[-1], c, w[1]);
            const double v = unwrap_second_diff_sq(w[-nx], c, w[nx]);
            const double d1 = unwrap_second_diff_sq(w[-nx - 1], c, w[nx + 1]);
            const double d2 = unwrap_second_diff_sq(w[-nx + 1], c, w[nx - 1]);
            r = __dadd_rn(__dadd_rn(__dadd_rn(h, v), d1), d2);
        }
        rel[idx] = r;
    }
}

// Edge e < ny (nx - 1): the horizontal pair (i, j)-(i, j + 1), e = i (nx - 1) + j; then the vertical pairs
// (i, j)-(i + 1, j), e = ny (nx - 1) + i nx + j.  Payload = e << 2 | (wrap count + 1).
__global__ void __launch_bounds__(256) unwrap_edge_pass(const double* phi, const double* rel, int nx, int ny,
                                                        unsigned long long* keys, unsigned* vals) {
    const long long plane = (long long)nx * ny;
    const long long n_horizontal = (long long)ny * (nx - 1);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < plane;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / nx), j = (int)(idx - (long long)i * nx);
        const double c = phi[idx], rc = rel[idx];
        if (j < nx - 1) {
            const long long e = (long long)i * (nx - 1) + j;
            keys[e] = (unsigned long long)__double_as_longlong(__dadd_rn(rc, rel[idx + 1]));
            vals[e] = ((unsigned)e << 2) | (unsigned)(unwrap_find_wrap(c, phi[idx + 1]) + 1);
        }
        if (i < ny - 1) {
            const long long e = n_horizontal + idx;
            keys[e] = (unsigned long long)__double_as_longlong(__dadd_rn(rc, rel[idx + nx]));
            vals[e] = ((unsigned)e << 2) | (unsigned)(unwrap_find_wrap(c, phi[idx + nx]) + 1);
        }
    }
}

// out = phi + 2 pi inc; with psi and maxbits given, 0 where |psi|^2 < 1e-6 max|psi|^2 of the plane
template <typename T>
__global__ void __launch_bounds__(256) unwrap_apply_pass(const typename cx_of<T>::type* psi, const double* phi,
                                                         const int* inc, const unsigned long long* maxbits,
                                                         long long plane, long long total, double* out) {
    typedef typename cx_of<T>::type C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        double v = __dadd_rn(phi[i], __dmul_rn(kUnwrapTwoPi, (double)inc[i]));
        if (psi != nullptr) {
            const C z = psi[i];
            const double n = (double)z.x * z.x + (double)z.y * z.y;
            const double thr = __longlong_as_double((long long)maxbits[i / plane]) * 1e-6;
            if (n < thr) v = 0.0;
        }
        out[i] = v;
    }
}

}  // namespace sgpe
