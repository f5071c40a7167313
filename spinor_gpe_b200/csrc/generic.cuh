// generic.cuh — the propagator on meshes the register-resident kernels of kernels.cuh do not cover: any EVEN number of
// points per axis whose prime factors are 2, 3, 5 and 7 (the reference asserts even sizes, pspinor.py:331-332, and hands
// them to cuFFT / MKL), including powers of two below 32.
//
// Same algorithm and state conventions as the fused passes, one operation per kernel:
//   gen_fft_pass    : batched 1-D Stockham transform along x or y, runtime radix plan (stages of 4 / 2 / 3 / 5 / 7), the
//                     lines of a CTA ping-pong between two shared-memory images; un-normalised in both directions
//   gen_kspace_pass : v <- v FA (S = sum |v|^2), v <- v FB (T = sum |v|^2) with the two-stage fixed-order reduction,
//                     totals and populations of the column pass (reference tensor_propagator.py:242, 270-271, 194)
//   gen_rspace_pass : normalise, I C P C I per pixel (tensor_propagator.py:244-267)
//   gen_scale_sign  : the (-1)^(x+y) shift signs and scales of the stand-alone transforms (tensor_tools.py:218-256)
// sgpe_api.cu strings them together behind the same run_col / run_row calls, so stepping, junction handling, stand-alone
// transforms and the energy expectation work unchanged.  These meshes are launch- and latency-bound by nature (a
// 2048-point line is a power of two); nothing here is tuned beyond coalesced accesses.
#pragma once

#include "kernels.cuh"

namespace sgpe {

#define SGPE_GEN_MAX_STAGES 16

struct GenFftArgs {
    const void* in; void* out;
    int n;                         // transform length
    int nlines;                    // lines per plane: along x: ny rows; along y: nx columns
    long long elem_stride;         // distance between consecutive elements of a line (1 along x, nx along y)
    long long line_stride;         // distance between consecutive lines (nx along x, 1 along y)
    long long plane;               // elements per (trajectory, component) plane
    int nplanes;                   // batch * 2
    int lpc;                       // lines per CTA
    int dir;                       // -1 forward, +1 inverse
    int nstages; int radix[SGPE_GEN_MAX_STAGES];
};

// exp(dir * 2 pi i * num / den), evaluated in double whatever the plan's precision
SGPE_DI double2 gen_root(int dir, long long num, long long den) {
    double s, c;
    sincospi(2.0 * (double)(num % den) / (double)den, &s, &c);
    double2 w; w.x = c; w.y = (double)dir * s;
    return w;
}

template <typename T>
__global__ void __launch_bounds__(256) gen_fft_pass(GenFftArgs a) {
    typedef typename cx_of<T>::type C;
    SGPE_DYN_SMEM(smem_raw);
    C* buf0 = reinterpret_cast<C*>(smem_raw);
    C* buf1 = buf0 + (size_t)a.lpc * a.n;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int groups = (a.nlines + a.lpc - 1) / a.lpc;
    const int pl = blockIdx.x / groups, l0 = (blockIdx.x % groups) * a.lpc;
    const int nl = (a.nlines - l0 < a.lpc) ? (a.nlines - l0) : a.lpc;
    const C* in = static_cast<const C*>(a.in) + (long long)pl * a.plane;
    C* out = static_cast<C*>(a.out) + (long long)pl * a.plane;
    const bool contiguous = (a.elem_stride == 1);
    // load: along x the elements of a line are contiguous, along y the lines of the CTA are adjacent in memory
    for (int e = tid; e < nl * a.n; e += nthr) {
        const int l = contiguous ? e / a.n : e % nl, i = contiguous ? e % a.n : e / nl;
        buf0[l * a.n + i] = in[(long long)(l0 + l) * a.line_stride + (long long)i * a.elem_stride];
    }
    __syncthreads();
    C* src = buf0; C* dst = buf1;
    int ns = 1;
    for (int st = 0; st < a.nstages; st++) {
        const int r = a.radix[st];
        const int nb = a.n / r;                       // butterflies per line
        for (int e = tid; e < nl * nb; e += nthr) {
            const int l = e / nb, jb = e % nb;
            const int k = jb % ns;
            const C* x = src + l * a.n;
            C* y = dst + l * a.n + (jb - k) * r + k;
            double2 v[7];
            for (int t = 0; t < r; t++) {
                const C z = x[jb + t * nb];
                double2 zz; zz.x = (double)z.x; zz.y = (double)z.y;
                v[t] = (t == 0 || k == 0) ? zz : cmul(zz, gen_root(a.dir, (long long)t * k, (long long)ns * r));
            }
            for (int q = 0; q < r; q++) {             // DFT_r, O(r^2): r <= 7
                double2 acc = v[0];
                for (int t = 1; t < r; t++) acc = cadd(acc, cmul(v[t], gen_root(a.dir, (long long)t * q, r)));
                C o; o.x = (T)acc.x; o.y = (T)acc.y;
                y[q * ns] = o;
            }
        }
        __syncthreads();
        C* tmp = src; src = dst; dst = tmp;
        ns *= r;
    }
    for (int e = tid; e < nl * a.n; e += nthr) {
        const int l = contiguous ? e / a.n : e % nl, i = contiguous ? e % a.n : e / nl;
        out[(long long)(l0 + l) * a.line_stride + (long long)i * a.elem_stride] = src[l * a.n + i];
    }
}

// out = in * scale * (-1)^(sx x + sy y)
template <typename T> struct GenScaleArgs {
    typedef typename cx_of<T>::type C;
    const C* in; C* out; int nx; long long total; int sign_x, sign_y; double scale;
    const double* scale_tot; double scale_num;      // optional: also * sqrt(scale_num / (tot[b][1] + tot[b][2]))
    long long per_batch;
};
template <typename T>
__global__ void __launch_bounds__(256) gen_scale_sign(GenScaleArgs<T> a) {
    typedef typename cx_of<T>::type C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / a.nx;
        const int x = (int)(i - row * a.nx);
        double s = a.scale;
        if (a.scale_tot != nullptr) {
            const double* tot = a.scale_tot + 4 * (i / a.per_batch);
            s *= sqrt(a.scale_num / (tot[1] + tot[2]));
        }
        if (((a.sign_x ? x : 0) + (a.sign_y ? (int)(row & 1) : 0)) & 1) s = -s;
        a.out[i] = cscale(a.in[i], (T)s);
    }
}

// k-space factors and sums on the whole state (ColArgs as the column pass takes them; tiles = CTAs of a trajectory)
template <typename T, int TM>
__global__ void __launch_bounds__(256) gen_kspace_pass(ColArgs<T> a) {
    typedef typename cx_of<T>::type C;
    SGPE_DYN_SMEM(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw);
    const int b = blockIdx.y, nblk = gridDim.x, tid = threadIdx.x;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};           // S0, T0, S1, T1
    const bool any_k = a.has_a || a.has_b;
    for (int comp = 0; comp < 2; comp++) {
        const long long base = ((long long)b * 2 + comp) * a.plane;
        const double* kin = (comp == 0 ? a.kin0 : a.kin1) + (long long)b * a.kin_bstride;
        for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < a.plane; i += (long long)nblk * blockDim.x) {
            const int y = (int)(i / a.nx), x = (int)(i - (long long)y * a.nx);
            C v = a.in[base + i];
            if (a.has_a) {
                const C f = a.kin_mode == 0 ? evo<TM, T, C>(__ldg(&kin[i]), a.ka_re, a.ka_im)
                    : combine_factor<TM>(__ldg(&a.xa[(long long)b * a.sepx_bstride + (long long)comp * a.nx + x]),
                                         __ldg(&a.ya[(long long)b * a.sepy_bstride + (long long)comp * a.ny + y]));
                v = mul_factor<TM>(v, f);
                acc[2 * comp] += (double)v.x * v.x + (double)v.y * v.y;
                if (a.aux != nullptr) a.aux[base + i] = v;
            }
            if (a.has_b) {
                const C f = a.kin_mode == 0 ? evo<TM, T, C>(__ldg(&kin[i]), a.kb_re, a.kb_im)
                    : combine_factor<TM>(__ldg(&a.xb[(long long)b * a.sepx_bstride + (long long)comp * a.nx + x]),
                                         __ldg(&a.yb[(long long)b * a.sepy_bstride + (long long)comp * a.ny + y]));
                v = mul_factor<TM>(v, f);
                acc[2 * comp + 1] += (double)v.x * v.x + (double)v.y * v.y;
            }
            a.out[base + i] = v;
        }
    }
    if (!any_k) return;
    if (!a.has_b) { acc[1] = acc[0]; acc[3] = acc[2]; }
    if (!a.has_a) { acc[0] = acc[1]; acc[2] = acc[3]; }
    cta_reduce<4>(acc, red);
    if (tid == 0) {
        double* p = a.partials + ((long long)b * nblk + blockIdx.x) * 4;
        p[0] = acc[0]; p[1] = acc[1]; p[2] = acc[2]; p[3] = acc[3];
        __threadfence();
        red[0] = (atomicAdd(&a.counter[b], 1u) == (unsigned)(nblk - 1)) ? 1.0 : 0.0;
    }
    __syncthreads();
    const bool last = red[0] != 0.0;
    __syncthreads();
    if (last) {       // fixed-order fold: bit-reproducible
        __threadfence();
        double t4[4] = {0.0, 0.0, 0.0, 0.0};
        const double* p = a.partials + (long long)b * nblk * 4;
        for (int t = tid; t < nblk; t += blockDim.x) {
#pragma unroll
            for (int q = 0; q < 4; q++) t4[q] += __ldcg(&p[4 * t + q]);
        }
        cta_reduce<4>(t4, red);
        if (tid == 0) {
            double* tot = a.totals + (long long)b * 4;
            tot[0] = t4[1] + t4[3];
            tot[1] = t4[0];
            tot[2] = t4[2];
            int slot = a.pops_slot;
            if (a.slot_ctr != nullptr) { slot = a.slot_ctr[b]; a.slot_ctr[b] = slot + 1; }
            if (a.pops != nullptr && slot >= 0) {
                double* pp = a.pops + (long long)b * a.pops_bstride + 2LL * slot;
                const double inv = a.atom_num / (t4[0] + t4[2]);
                pp[0] = t4[0] * inv; pp[1] = t4[2] * inv;
            }
            a.counter[b] = 0u;
        }
    }
}

// real-space operators per pixel (the point-wise section of row_pass on the whole state)
template <typename T, int TM>
__global__ void __launch_bounds__(256) gen_rspace_pass(RowArgs<T> a) {
    typedef typename cx_of<T>::type C;
    const int b = blockIdx.y;
    const T alpha = (T)sqrt(a.norm_c / a.totals[(long long)b * 4]);
    const bool same_pot = (a.pot0 == a.pot1);
    T cu_diag = (T)1, cu_s = (T)0;
    if (a.cpl_mode == 1) {
        C one; one.x = (T)1; one.y = (T)0;
        C t01, t10;
        coupling_entries<TM, T, C>(a.omega_b[b] * a.tc, one, cu_diag, t01, t10);
        cu_s = (TM == TM_REAL) ? -t01.y : -t01.x;
    }
    const long long base0 = (long long)b * 2 * a.plane, base1 = base0 + a.plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.plane; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / a.nx), x = (int)(i - (long long)y * a.nx);
        C p = cscale(a.in[base0 + i], alpha), q = cscale(a.in[base1 + i], alpha);
        const double n0 = (double)p.x * p.x + (double)p.y * p.y;
        const double n1 = (double)q.x * q.x + (double)q.y * q.y;
        const C i0 = evo<TM, T, C>(a.g_uu * n0 + a.g_ud * n1, a.ti_re, a.ti_im);
        const C i1 = evo<TM, T, C>(a.g_dd * n1 + a.g_ud * n0, a.ti_re, a.ti_im);
        p = mul_factor<TM>(p, i0); q = mul_factor<TM>(q, i1);
        T diag = (T)1; C o01, o10;
        o01.x = o01.y = o10.x = o10.y = (T)0;
        if (a.cpl_mode) {
            C ph; ph.x = (T)1; ph.y = (T)0;
            if (a.eiphi != nullptr) ph = __ldg(&a.eiphi[x]);
            if (a.cpl_mode == 1) {
                diag = cu_diag;
                if (TM == TM_REAL) {
                    o01.x = -cu_s * ph.y; o01.y = -cu_s * ph.x; o10.x = cu_s * ph.y; o10.y = -cu_s * ph.x;
                } else {
                    o01.x = -cu_s * ph.x; o01.y = cu_s * ph.y; o10.x = -cu_s * ph.x; o10.y = -cu_s * ph.y;
                }
            } else {
                coupling_entries<TM, T, C>(__ldg(&a.coupling[(long long)b * a.cpl_bstride + i]) * a.tc, ph, diag, o01, o10);
            }
            const C p2 = cadd(cscale(p, diag), cmul(o01, q));
            const C q2 = cadd(cmul(o10, p), cscale(q, diag));
            p = p2; q = q2;
        }
        C f0, f1;
        if (a.pot_mode == 0) {
            const long long pi = (long long)b * a.pot_bstride + i;
            f0 = evo<TM, T, C>(__ldg(&a.pot0[pi]), a.tp_re, a.tp_im);
            f1 = same_pot ? f0 : evo<TM, T, C>(__ldg(&a.pot1[pi]), a.tp_re, a.tp_im);
        } else {
            const long long ox = (long long)b * a.sepx_bstride + x, oy = (long long)b * a.sepy_bstride + y;
            f0 = combine_factor<TM>(__ldg(&a.px[ox]), __ldg(&a.py[oy]));
            f1 = combine_factor<TM>(__ldg(&a.px[ox + a.nx]), __ldg(&a.py[oy + a.ny]));
        }
        p = mul_factor<TM>(p, f0); q = mul_factor<TM>(q, f1);
        if (a.cpl_mode) {
            const C p2 = cadd(cscale(p, diag), cmul(o01, q));
            const C q2 = cadd(cmul(o10, p), cscale(q, diag));
            p = p2; q = q2;
        }
        a.out[base0 + i] = mul_factor<TM>(p, i0);
        a.out[base1 + i] = mul_factor<TM>(q, i1);
    }
}

}  // namespace sgpe
