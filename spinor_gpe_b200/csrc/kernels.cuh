// kernels.cuh — the two fused passes of the split-step propagator, plus small helpers.
//
//   col_pass : for W adjacent columns of one component:  [FFT_y] -> [K_a] -> (S = sum |.|^2) -> [K_b]
//              -> (T = sum |.|^2) -> [iFFT_y].   K_a is the trailing kinetic half-step of single step s,
//              K_b the leading one of single step s+1 (reference tensor_propagator.py:270 and :242), the
//              sums are the two global reductions of ttools.norm (tensor_tools.py:303; the real-space
//              one is taken in k-space via Parseval) and ttools.calc_pops (:482).
//   row_pass : for one row of both components: [iFFT_x] -> normalise -> I -> C -> P -> C -> I -> [FFT_x]
//              (reference tensor_propagator.py:243-269).
//
// The fftshift/ifftshift of ttools.fft_2d/ifft_2d (tensor_tools.py:226, 255) never appear: for even N
// they are the sign (-1)^(i+j) on the real-space side, every real-space operator is diagonal in (i,j),
// and the sign cancels between the inverse and the forward transform.  The stand-alone transforms
// apply it explicitly (sign_in / sign_out).
#pragma once

#include "fft_core.cuh"

namespace sgpe {

enum { TM_REAL = 0, TM_IMAG = 1 };

// exp(x) for the imaginary-time factors evaluated per grid point (the non-linear term; dense operator grids).
// Same scheme as the library routine — x = k ln2 + r, |r| <= ln2/2, degree-11 minimax polynomial (3e-18), 2^k
// through the exponent field — without its out-of-range branches: k is clamped to [-1021, 1022], so every result
// that is a normal double is computed as usual and those that would be subnormal / overflow come out as ~1e-308 /
// ~1e+308 instead of (nearly) 0 / inf — a factor of that size multiplies nothing that matters.  Max error 1 ulp
// against exp() (tests/test_fast_exp.py); roughly half the instructions.
SGPE_DI double sgpe_exp(double x) {
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);          // 1.5 * 2^52: low word = round(x/ln2)
    int k = (int)(unsigned)(__double_as_longlong(t) & 0xffffffffLL);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, -6.93147180369123816490e-01, x);
    r = fma(kf, -1.90821492927058770002e-10, r);
    double p = 2.5110049204818658e-08;
    p = fma(p, r, 2.763265472252779e-07);
    p = fma(p, r, 2.755724088722987e-06);
    p = fma(p, r, 2.4801485441561313e-05);
    p = fma(p, r, 0.00019841269890076403);
    p = fma(p, r, 0.0013888888952352863);
    p = fma(p, r, 0.008333333333319589);
    p = fma(p, r, 0.04166666666648795);
    p = fma(p, r, 0.1666666666666668);
    p = fma(p, r, 0.5000000000000019);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    k = k < -1021 ? -1021 : (k > 1022 ? 1022 : k);
    return __longlong_as_double(__double_as_longlong(p) + ((long long)k << 52));
}

// exp(x) for |x| <= 1/16: Taylor polynomial of degree 9 (truncation 2.5e-19 relative), no range reduction, nine FMAs
// against the ~20 floating-point and integer operations of sgpe_exp.  What the non-linear factor of an imaginary-time
// sub-step asks for when g n dt / 2 is small (the benchmark's: <= 0.03) — two of them per pixel in the row pass.
SGPE_DI double sgpe_exp_small(double x) {
    double p = 2.7557319223985893e-06;            // 1 / 9!
    p = fma(p, x, 2.4801587301587302e-05);        // 1 / 8!
    p = fma(p, x, 1.9841269841269841e-04);        // 1 / 7!
    p = fma(p, x, 1.3888888888888889e-03);        // 1 / 6!
    p = fma(p, x, 8.3333333333333332e-03);        // 1 / 5!
    p = fma(p, x, 4.1666666666666664e-02);        // 1 / 4!
    p = fma(p, x, 1.6666666666666666e-01);        // 1 / 3!
    p = fma(p, x, 0.5);
    p = fma(p, x, 1.0);
    p = fma(p, x, 1.0);
    return p;
}
// the two interaction factors of a pixel in imaginary time (double precision): the short polynomial when both arguments
// are small, sgpe_exp otherwise (both within one ulp of exp)
SGPE_DI void sgpe_exp_pair(double x0, double x1, double& e0, double& e1) {
    if (fabs(x0) <= 0.0625 && fabs(x1) <= 0.0625) { e0 = sgpe_exp_small(x0); e1 = sgpe_exp_small(x1); }
    else { e0 = sgpe_exp(x0); e1 = sgpe_exp(x1); }
}

// reciprocal to (nearly) full precision without the library's special cases: hardware seed (~20 bits) + two Newton steps
SGPE_DI double sgpe_rcp(double d) {
    double r = SGPE_RCP_SEED(d);
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// sqrt(n) for the same store: hardware seed of 1/sqrt (~20 bits), two coupled Newton steps on g ~ sqrt(n), h ~ 1/(2 sqrt(n)),
// and a final residual correction (< 1 ulp in tests/test_fast_exp.py); n below 1e-290 (always masked) gives 0
SGPE_DI double sgpe_sqrt(double n) {
    const double y = SGPE_RSQRT_SEED(n);
    double g = n * y, h = 0.5 * y;
    double e = fma(-h, g, 0.5);
    g = fma(g, e, g); h = fma(h, e, h);
    e = fma(-h, g, 0.5);
    g = fma(g, e, g); h = fma(h, e, h);
    g = fma(fma(-g, g, n), h, g);
    return n > 1e-290 ? g : 0.0;
}

// atan2(y, x) for the polar store of the energy tracking (two per pixel and step).  t = min / max of the magnitudes lies
// in [0, 1]; with c = i / 16 the nearest sixteenth, atan t = atan c + atan s, s = (min - c max) / (max + c min),
// |s| <= 1/32, so ONE division and a degree-11 odd Taylor polynomial (truncation 3e-21) replace the library's
// division + degree-~40 polynomial: about half the FP64 instructions, < 2 ulp (tests/test_fast_exp.py).
// Magnitudes below 1e-280 (always masked: n < 1e-6 max n) return 0.
__device__ const double sgpe_atan_tab[17] = {0.0, 0.06241880999595735, 0.12435499454676144, 0.18534794999569476, 0.24497866312686414, 0.3028848683749714, 0.35877067027057225, 0.4124104415973873, 0.4636476090008061, 0.5123894603107377, 0.5585993153435624, 0.6022873461349642, 0.6435011087932844, 0.6823165548747481, 0.7188299996216245, 0.7531512809621944, 0.7853981633974483};
SGPE_DI double sgpe_atan2(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const bool swap = ay > ax;
    const double mx = swap ? ay : ax, mn = swap ? ax : ay;
    const double t0 = mn * SGPE_RCP_SEED(mx);
    int i = SGPE_D2I_RN(t0 * 16.0);
    i = i < 0 ? 0 : (i > 16 ? 16 : i);
    const double c = (double)i * 0.0625;
    const double num = fma(-c, mx, mn), den = fma(c, mn, mx);
    const double rd = sgpe_rcp(den);
    double q = num * rd;
    q = fma(fma(-den, q, num), rd, q);
    const double s2 = q * q;
    double p = -1.0 / 11.0;
    p = fma(p, s2, 1.0 / 9.0);
    p = fma(p, s2, -1.0 / 7.0);
    p = fma(p, s2, 1.0 / 5.0);
    p = fma(p, s2, -1.0 / 3.0);
    double a = fma(q * s2, p, q) + sgpe_atan_tab[i];
    if (swap) a = 1.5707963267948966 - a;
    if (x < 0.0) a = 3.141592653589793 - a;
    if (!(mx > 1e-280)) a = 0.0;
    return copysign(a, y);
}

// |x|^2 as a double for the norm sums: complex128 in double; complex64 squared in single precision (two FP32 operations
// and one conversion instead of two conversions and two FP64 operations per element - the FP64 pipe of the B200 has a
// quarter of the FP32 lanes) and ACCUMULATED in double: the sum over the mesh keeps its 1e-7 / sqrt(points) accuracy.
SGPE_DI double abs_sq(double2 x) { return x.x * x.x + x.y * x.y; }
SGPE_DI double abs_sq(float2 x) { return (double)fmaf(x.x, x.x, x.y * x.y); }

// exp(-i * e * tau), tau = (tr, ti):  real time tau = (dt, 0);  imaginary time tau = (0, -dt)
template <int TM, typename T, typename C> SGPE_DI C evo(double e, double tr, double ti) {
    C r;
    if constexpr (sizeof(T) == 4) {
        // complex64 plans: the argument is formed in double, the transcendental runs on the FP32 pipe
        if (TM == TM_REAL) {
            float s, c;
            sincosf((float)(e * tr), &s, &c);
            r.x = c; r.y = -s;
        } else {
            r.x = expf((float)(e * ti)); r.y = 0.f;
        }
    } else {
        if (TM == TM_REAL) {
            double s, c;
            sincos(e * tr, &s, &c);
            r.x = (T)c; r.y = (T)(-s);
        } else {
            r.x = (T)sgpe_exp(e * ti); r.y = (T)0;
        }
    }
    return r;
}

// conjugate E values in place (the inverse transform runs the forward code between two conjugations, so that
// each pass contains ONE copy of the FFT code, executed by a two-iteration loop: instruction-cache footprint)
template <int E, typename C> SGPE_DI void conj_all(C (&v)[E]) {
#pragma unroll
    for (int m = 0; m < E; m++) v[m].y = -v[m].y;
}

// ---------------------------------------------------------------------------------------------
// Fused exchange of the slab (multi-GPU) mode.  Instead of "store locally, pack, NCCL all-to-all, unpack", the
// LAST pass of each direction stores every element straight into the buffer of the rank that owns it in the
// next direction (peer device memory mapped through CUDA IPC, NVLink / NVSwitch): the transfer overlaps the
// transforms tile by tile and the pack / unpack passes disappear.  All arrays stay row-major:
//   row slab of rank r : [2][Ny/P][Nx]   (rows  Y in [r Ny/P, (r+1) Ny/P))
//   k slab   of rank q : [2][Ny][Nx/P]   (columns X in [q Nx/P, (q+1) Nx/P))
// mode 1 (written by the row direction): element (Y_local, X) -> rank X / seg, k slab;
// mode 2 (written by the k direction)  : element (Y, X_local) -> rank Y / seg, row slab.
#define SGPE_MAX_PEERS 16
template <typename C> struct Scatter {
    C* peer[SGPE_MAX_PEERS];
    int mode;              // 0: off (plain store to `out`)
    int seg;               // mode 1: Nx/P columns per rank;  mode 2: Ny/P rows per rank
    int drow;              // row length of the destination arrays (mode 1: Nx/P, mode 2: Nx)
    int base;              // mode 1: first global row of this rank;  mode 2: first global column of this rank
    long long dplane;      // component stride of the destination arrays
};
template <typename C> SGPE_DI C* scatter_ptr(const Scatter<C>& s, int comp, int Y, int X) {
    if (s.mode == 1) {
        const int q = X / s.seg;
        return s.peer[q] + (long long)comp * s.dplane + (long long)(Y + s.base) * s.drow + (X - q * s.seg);
    }
    const int q = Y / s.seg;
    return s.peer[q] + (long long)comp * s.dplane + (long long)(Y - q * s.seg) * s.drow + (X + s.base);
}

template <typename T> struct ColArgs {
    typedef typename cx_of<T>::type C;
    const C* in;  C* out;          // [B][2][ny][nx]
    const C* tw;                   // per-stage twiddle tables for ny: [radix-8 plan (ny)][radix-16 plan (ny)]
    int nx, ny; long long plane;
    int do_fwd, do_inv;
    int prefetch_ahead;            // > 0: prefetch into L2 the tile this many CTAs ahead (the next one on this SM)
    // k-space factors: v <- v*FA (then S = sum|v|^2), v <- v*FB (then T = sum|v|^2)
    int has_a, has_b;
    int kin_mode;                  // 0: dense kin grids, factors evaluated here; 1: separable tables
    const double* kin0; const double* kin1; long long kin_bstride;   // dense: [ny][nx], stored (shifted) k order
    double ka_re, ka_im, kb_re, kb_im;                               // dense: time arguments of FA / FB
    const C* xa; const C* ya; const C* xb; const C* yb;              // separable: [2][nx] / [2][ny] factor tables
    long long sepx_bstride, sepy_bstride;
    int sign_in, sign_out; double scale_out;     // (-1)^y on load / store, output scale (stand-alone 1-D use)
    double* partials;              // [B][ntiles][2]
    unsigned* counter;             // [B]
    double* totals;                // [B][4] : T, S0, S1, -
    double* pops; long long pops_bstride; int pops_slot;   // [B][n][2], slot < 0: don't record
    int* slot_ctr;                 // [B] or null.  Non-null (CUDA-graph replay of the steady-state step: node arguments
                                   // are frozen): the slot is read from slot_ctr[b] and post-incremented by the fold
    double atom_num;
    unsigned long long* dbg;       // dev tool: per-CTA phase timestamps [nCTA][8] (null in production; generic kernel only)
    const void* tile_map;          // host pointer to the SgpeTileMap of `in` (persistent kernel; read by the launcher only)
    int kernel_sel;                // 0: one tile per CTA, 1 / 2: persistent CTAs with asynchronously staged tiles (col_pass_p)
    unsigned long long* zero2;     // optional [B][2] words cleared by the first CTA (the density maxima the NEXT pass folds)
    C* aux;                        // optional second output [B][2][ny][nx]: the state right after FA — at a full-step
                                   // junction that is the (un-normalised) k-space state of the step boundary, which
                                   // per-step energy tracking transforms back on the side (sgpe_full_steps_energy)
};

// multiply by a factor: in imaginary time every factor is real (only .x is meaningful)
template <int TM, typename C> SGPE_DI C mul_factor(C v, C f) {
    if (TM == TM_REAL) return cmul(v, f);
    return cscale(v, f.x);
}
template <int TM, typename C> SGPE_DI C combine_factor(C f, C g) {
    if (TM == TM_REAL) return cmul(f, g);
    C r; r.x = f.x * g.x; r.y = 0; return r;
}

// FAST = 1: the steady-state junction (forward + factors + inverse, separable tables, no sign / scale) with the
// other branches compiled out.  FAST = 2: the same plus the store of the boundary state to `aux`.
// G > 1: the CTA's tile of G * W adjacent columns is worked on by G independent barrier groups of W columns each
// (own shared-memory image, own named barrier, own partial-sum slot).  Where the register file has room for one
// CTA only (complex128 at 2048 points: 512 threads x 128 registers), the groups give the SM two instruction streams
// that drift out of phase — one group's butterflies overlap the other's exchange — while the CTA as a whole still
// touches G * W * sizeof(C) = 64 contiguous bytes of every row.
template <typename T, int N, int E, int W, int TM, int FAST, int G = 1>
__global__ void __launch_bounds__(G * W * N / E, (G * W * N / E <= 256) ? 2 : 1) col_pass(ColArgs<T> a) {
    typedef typename cx_of<T>::type C;
    constexpr int NT = N / E;
    constexpr int TG = W * NT;                 // threads per group
    SGPE_DYN_SMEM(smem_raw);
    const int g = (G == 1) ? 0 : (int)threadIdx.x / TG;
    C* sm = reinterpret_cast<C*>(smem_raw) + (size_t)g * N * W;
    double* red = reinterpret_cast<double*>(smem_raw + sizeof(C) * (size_t)N * W * G) + g * (32 * 4);

    const int tid = (G == 1) ? (int)threadIdx.x : (int)threadIdx.x - g * TG;
    const int c = tid % W, j = tid / W;
    const int tiles_per_comp = a.nx / (W * G);
    const int ntiles = 2 * tiles_per_comp;
    const int nslots = ntiles * G;             // partial-sum slots: one per group
    const int tile = blockIdx.x;
    const int slot = tile * G + g;
    const int comp = tile / tiles_per_comp;
    const int col = (tile % tiles_per_comp) * (W * G) + g * W + c;
    const GroupBar gbar = {1 + g, TG};
    const int b = blockIdx.y;
    const bool do_fwd = FAST ? true : (a.do_fwd != 0), do_inv = FAST ? true : (a.do_inv != 0);
    if (!FAST && a.zero2 != nullptr && blockIdx.x == 0 && threadIdx.x < 2) a.zero2[2 * b + threadIdx.x] = 0ull;
    const int kin_mode = FAST ? 1 : a.kin_mode;
    const int sign_in = FAST ? 0 : a.sign_in, sign_out = FAST ? 0 : a.sign_out;
    const long long off = ((long long)b * 2 + comp) * a.plane + col;

#define SGPE_CMARK(k) do { if (!FAST && a.dbg != nullptr && tid == 0 && g == 0) a.dbg[(long long)blockIdx.x * 8 + (k)] = SGPE_GLOBALTIMER(); } while (0)
    if (!FAST && a.dbg != nullptr && tid == 0 && g == 0) a.dbg[(long long)blockIdx.x * 8 + 7] = SGPE_SMID();
    SGPE_CMARK(0);
    C v[1][E];
#pragma unroll
    for (int m = 0; m < E; m++) v[0][m] = SGPE_LD_STREAM(&a.in[off + (long long)(j + m * NT) * a.nx]);
    if (a.prefetch_ahead > 0 && c == 0) {
        const int nt = tile + a.prefetch_ahead;
        if (nt < ntiles) {
            const long long noff = ((long long)b * 2 + nt / tiles_per_comp) * a.plane + (nt % tiles_per_comp) * (W * G) + g * W;
#pragma unroll
            for (int m = 0; m < E; m++) SGPE_PREFETCH_L2(&a.in[noff + (long long)(j + m * NT) * a.nx]);
        }
    }
    if (sign_in) {
#pragma unroll
        for (int m = 0; m < E; m++)
            if ((j + m * NT) & 1) { v[0][m].x = -v[0][m].x; v[0][m].y = -v[0][m].y; }
    }

    C* const sms[1] = {sm};
    if (!FAST && a.dbg != nullptr) { if (v[0][0].x == (T)1.2345e300 || v[0][E - 1].y == (T)1.2345e300) a.dbg[6] = 1; SGPE_CMARK(1); }
    // (the two-iteration-loop trick of row_pass was measured slower here: 16 elements per thread, more spills)
    if (do_fwd) {
        if constexpr (G == 1) cta_fft<T, N, E, -1, W, 1>(v, j, c, sms, a.tw + (E == 16 ? N : 0));
        else group_fft<T, N, E, -1, W>(v, j, c, sms, a.tw + (E == 16 ? N : 0), gbar);
    }
    SGPE_CMARK(2);
    double acc[2] = {0.0, 0.0};   // S (after FA), T (after FB)
    const bool any_k = a.has_a || a.has_b;

    if (any_k) {
        if (kin_mode == 0) {
            const double* kin = (comp == 0 ? a.kin0 : a.kin1) + (long long)b * a.kin_bstride + col;
            double e[E];
#pragma unroll
            for (int m = 0; m < E; m++) e[m] = __ldg(&kin[(long long)(j + m * NT) * a.nx]);
#pragma unroll
            for (int m = 0; m < E; m++) {
                C x = v[0][m];
                if (a.has_a) {
                    x = mul_factor<TM>(x, evo<TM, T, C>(e[m], a.ka_re, a.ka_im));
                    acc[0] += abs_sq(x);
                    if (!FAST && a.aux != nullptr) SGPE_ST_STREAM(&a.aux[off + (long long)(j + m * NT) * a.nx], x);
                }
                if (a.has_b) {
                    x = mul_factor<TM>(x, evo<TM, T, C>(e[m], a.kb_re, a.kb_im));
                    acc[1] += abs_sq(x);
                }
                v[0][m] = x;
            }
        } else {
            const long long ox = (long long)b * a.sepx_bstride + (long long)comp * a.nx + col;
            const long long oy = (long long)b * a.sepy_bstride + (long long)comp * a.ny + j;
            C fxa, fxb;
            fxa.x = (T)1; fxa.y = (T)0; fxb = fxa;
            if (a.has_a) fxa = __ldg(&a.xa[ox]);
            if (a.has_b) fxb = __ldg(&a.xb[ox]);
#pragma unroll
            for (int m = 0; m < E; m++) {
                C x = v[0][m];
                if (a.has_a) {
                    x = mul_factor<TM>(x, combine_factor<TM>(fxa, __ldg(&a.ya[oy + m * NT])));
                    acc[0] += abs_sq(x);
                    if ((FAST == 2) || (!FAST && a.aux != nullptr))
                        SGPE_ST_STREAM(&a.aux[off + (long long)(j + m * NT) * a.nx], x);
                }
                if (a.has_b) {
                    x = mul_factor<TM>(x, combine_factor<TM>(fxb, __ldg(&a.yb[oy + m * NT])));
                    acc[1] += abs_sq(x);
                }
                v[0][m] = x;
            }
        }
        if (!a.has_b) acc[1] = acc[0];
        if (!a.has_a) acc[0] = acc[1];
    }

    // The tile's partial sums are published BEFORE the inverse transform: the barriers of the reduction and the
    // round trip of the ticket to L2 then hide behind the transform, and the CTA retires right after its stores
    // (with one CTA per SM the epilogue is dead time for the whole SM).
    unsigned ticket = 0u;
    if (any_k) {
        if constexpr (G == 1) cta_reduce<2>(acc, red);
        else group_reduce<2>(acc, red, tid, TG, gbar);
        if (tid == 0) {
            double* p = a.partials + ((long long)b * nslots + slot) * 2;
            p[0] = acc[0]; p[1] = acc[1];
            __threadfence();
            ticket = atomicAdd(&a.counter[b], 1u);
        }
    }

    SGPE_CMARK(3);
    if (do_inv) {
        if constexpr (G == 1) cta_fft<T, N, E, +1, W, 1>(v, j, c, sms, a.tw + (E == 16 ? N : 0));
        else group_fft<T, N, E, +1, W>(v, j, c, sms, a.tw + (E == 16 ? N : 0), gbar);
    }
    SGPE_CMARK(4);

    if (!FAST && (sign_out || a.scale_out != 1.0)) {
        const T sc = (T)a.scale_out;
#pragma unroll
        for (int m = 0; m < E; m++) {
            const T s = (sign_out && ((j + m * NT) & 1)) ? -sc : sc;
            v[0][m] = cscale(v[0][m], s);
        }
    }
#pragma unroll
    for (int m = 0; m < E; m++) SGPE_ST_STREAM(&a.out[off + (long long)(j + m * NT) * a.nx], v[0][m]);
    SGPE_CMARK(5);
#undef SGPE_CMARK

    if (any_k) {
        if (tid == 0) red[0] = (ticket == (unsigned)(nslots - 1)) ? 1.0 : 0.0;
        if constexpr (G == 1) __syncthreads(); else gbar.sync();
        const bool last = red[0] != 0.0;      // (the fold's first write to `red` comes after a barrier of its own)
        if (last) {       // the last tile (group) of this trajectory folds the partials in a fixed order
            __threadfence();
            double t4[4] = {0.0, 0.0, 0.0, 0.0};     // S0, T0, S1, T1
            const double* p = a.partials + (long long)b * nslots * 2;
            for (int t = tid; t < nslots; t += TG) {
                const int cp = (t >= nslots / 2) ? 2 : 0;
                t4[cp + 0] += __ldcg(&p[2 * t]);
                t4[cp + 1] += __ldcg(&p[2 * t + 1]);
            }
            if constexpr (G == 1) cta_reduce<4>(t4, red);
            else group_reduce<4>(t4, red, tid, TG, gbar);
            if (tid == 0) {
                double* tot = a.totals + (long long)b * 4;
                tot[0] = t4[1] + t4[3];
                tot[1] = t4[0];
                tot[2] = t4[2];
                int slot = a.pops_slot;
                if (a.slot_ctr != nullptr) { slot = a.slot_ctr[b]; a.slot_ctr[b] = slot + 1; }
                if (a.pops != nullptr && slot >= 0) {
                    // calc_pops of the normalised psi_k: N * S_c / (S_0 + S_1)   (tensor_tools.py:482)
                    double* pp = a.pops + (long long)b * a.pops_bstride + 2LL * slot;
                    const double inv = a.atom_num / (t4[0] + t4[2]);
                    pp[0] = t4[0] * inv; pp[1] = t4[2] * inv;
                }
                a.counter[b] = 0u;
            }
        }
    }
}

// the twiddle source of a kernel: the plan's tables in global memory, or the kernel's copy in shared memory
template <typename C, int TWS> struct TwSource {
    typedef const C* type;
    SGPE_DI static type make(const C* global, const C*) { return global; }
};
template <typename C> struct TwSource<C, 1> {
    typedef SmemTable<C> type;
    SGPE_DI static type make(const C*, const C* shared) { type t; t.p = shared; return t; }
};

// k-space factors of one thread's E points of a column: v <- v FA (S += |v|^2, optional copy to aux), v <- v FB
// (T += |v|^2); which of the two exist is a compile-time choice here (branches inside the unrolled loop keep the
// compiler from batching the table loads: every load then waits out its own L2 round trip).
template <typename T, int E, int NT, int TM, bool HAS_A, bool HAS_B, bool AUX, typename C>
SGPE_DI void k_factors(C (&v)[E], C fxa, C fxb, const C* __restrict__ ya, const C* __restrict__ yb, C* aux, long long aux_stride,
                       double (&acc)[2]) {
    constexpr int CH = E < 4 ? E : 4;          // table loads in batches of four: latency overlapped, few registers
#pragma unroll
    for (int m0 = 0; m0 < E; m0 += CH) {
        C fa[CH], fb[CH];
#pragma unroll
        for (int q = 0; q < CH; q++) {
            if (HAS_A) fa[q] = __ldg(&ya[(m0 + q) * NT]);
            if (HAS_B) fb[q] = __ldg(&yb[(m0 + q) * NT]);
        }
#pragma unroll
        for (int q = 0; q < CH; q++) {
            const int m = m0 + q;
            C x = v[m];
            if (HAS_A) {
                x = mul_factor<TM>(x, combine_factor<TM>(fxa, fa[q]));
                acc[0] += abs_sq(x);
                if (AUX) SGPE_ST_STREAM(&aux[(long long)m * aux_stride], x);
            }
            if (HAS_B) {
                x = mul_factor<TM>(x, combine_factor<TM>(fxb, fb[q]));
                acc[1] += abs_sq(x);
            }
            v[m] = x;
        }
    }
    if (!HAS_B) acc[1] = acc[0];
    if (!HAS_A) acc[0] = acc[1];
}

// imaginary time, the y-dependent factor tables of the launch held in shared memory (real factors): no L2 round trips
// in the K phase of a persistent CTA whose shared-memory footprint leaves no L1 to cache the tables in
template <typename T, int E, int NT, bool HAS_A, bool HAS_B, bool AUX, typename C>
SGPE_DI void k_factors_smem_imag(C (&v)[E], T fxa, T fxb, const T* ya, const T* yb, C* aux, long long aux_stride,
                                 double (&acc)[2]) {
#pragma unroll
    for (int m = 0; m < E; m++) {
        C x = v[m];
        if (HAS_A) {
            x = cscale(x, fxa * ya[m * NT]);
            acc[0] += abs_sq(x);
            if (AUX) SGPE_ST_STREAM(&aux[(long long)m * aux_stride], x);
        }
        if (HAS_B) {
            x = cscale(x, fxb * yb[m * NT]);
            acc[1] += abs_sq(x);
        }
        v[m] = x;
    }
    if (!HAS_B) acc[1] = acc[0];
    if (!HAS_A) acc[0] = acc[1];
}

// the same with dense kinetic grids: the factors exp(-i kin tau) are evaluated per point (general operators)
template <typename T, int E, int TM, bool HAS_A, bool HAS_B, bool AUX, typename C>
SGPE_DI void k_factors_dense(C (&v)[E], const double* __restrict__ kin, long long kin_stride, double ka_re, double ka_im,
                             double kb_re, double kb_im, C* aux, long long aux_stride, double (&acc)[2]) {
    constexpr int CH = E < 4 ? E : 4;
#pragma unroll
    for (int m0 = 0; m0 < E; m0 += CH) {
        double e[CH];
#pragma unroll
        for (int q = 0; q < CH; q++) e[q] = __ldg(&kin[(long long)(m0 + q) * kin_stride]);
#pragma unroll
        for (int q = 0; q < CH; q++) {
            const int m = m0 + q;
            C x = v[m];
            if (HAS_A) {
                x = mul_factor<TM>(x, evo<TM, T, C>(e[q], ka_re, ka_im));
                acc[0] += abs_sq(x);
                if (AUX) SGPE_ST_STREAM(&aux[(long long)m * aux_stride], x);
            }
            if (HAS_B) {
                x = mul_factor<TM>(x, evo<TM, T, C>(e[q], kb_re, kb_im));
                acc[1] += abs_sq(x);
            }
            v[m] = x;
        }
    }
    if (!HAS_B) acc[1] = acc[0];
    if (!HAS_A) acc[0] = acc[1];
}

// Persistent column pass of the steady-state junction (what FAST = 1 / 2 compute) with the NEXT tile staged
// asynchronously: one CTA per SM slot walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ...; TMA (cp.async.bulk.tensor,
// N / 256 boxes of 256 rows x 64 bytes, completion on an mbarrier) lands the CTA's next tile in shared memory while the
// current one is still being worked on, so the global-load latency and the CTA turnaround of the one-tile-per-CTA
// kernel (4 + 1 us of a 17 us tile at 2048 points, profiles/r02_timeline.txt) leave the critical path.
//   S = complex image of a tile: TMA landing zone and exchange buffer.
//   XSPLIT = 1: the inverse transform exchanges through a second, REAL image X (re and im one after the other, half
//               the size): S is free - and refilled - right after the forward transform (a window of ~10 us);
//   XSPLIT = 0: both transforms exchange through S; it is refilled behind the last exchange read of the inverse
//               transform, while the last butterflies run and the tile is stored (a window of ~3 us, shared memory and
//               L1 as in the one-tile-per-CTA kernel).
// The partial sums of a tile are stored right away, the ticket (fence + atomic) is taken ONCE per CTA after its last
// tile; the CTA that completes the count folds the partials in a fixed order as before.
// (resident CTAs asked of the compiler: complex64 tiles of half the width are meant to run two to an SM - 512 threads at
// 64 registers; without the bound ptxas takes 128 and only one fits.  SGPE_F32_COL_BLOCKS = 2 asks for the two.)
#ifndef SGPE_F32_COL_BLOCKS
#define SGPE_F32_COL_BLOCKS 1
#endif
template <typename T, int N, int E, int W, int TM, int XSPLIT, int TWS, int KM = 1>
__global__ void __launch_bounds__(W * N / E, (W * N / E <= 256 || (SGPE_F32_COL_BLOCKS == 2 && sizeof(T) == 4 && W * N / E <= 512 && W * N * 2 * sizeof(T) <= 64 * 1024)) ? 2 : 1)
col_pass_p(const SGPE_GRID_CONSTANT SgpeTileMap tmap, ColArgs<T> a) {
    typedef typename cx_of<T>::type C;
    constexpr int NT = N / E;
    constexpr int R0 = E;
    constexpr int BOXR = N < 256 ? N : 256, NBOX = N / BOXR;
    constexpr unsigned TILE_BYTES = (unsigned)(N * W * sizeof(C));
    constexpr int LAST_NS = LastStage<N, E>::value;
    SGPE_DYN_SMEM_128(smem_p);
    C* const S = reinterpret_cast<C*>(smem_p);
    T* const X = reinterpret_cast<T*>(smem_p + TILE_BYTES);
    // TWS = 1: the twiddle tables of this plan (N entries) live in shared memory for the life of the CTA.
    // TWS = 2 (imaginary time, factor tables): the same bytes hold the y-dependent k factors of the launch instead,
    // real parts only, [FA | FB] of the component the CTA is working on
    C* const TW = reinterpret_cast<C*>(smem_p + TILE_BYTES + (XSPLIT ? TILE_BYTES / 2 : 0));
    T* const KT = reinterpret_cast<T*>(TW);
    static_assert(TWS != 2 || (TM == TM_IMAG && KM == 1), "shared-memory k factors: imaginary time, factor tables");
    double* const red = reinterpret_cast<double*>(smem_p + TILE_BYTES + (XSPLIT ? TILE_BYTES / 2 : 0) + (TWS ? N * sizeof(C) : 0));
    SgpeMbar* const mbar = reinterpret_cast<SgpeMbar*>(red + 32 * 4);
    int* const flag = reinterpret_cast<int*>(mbar + 1);

    const int b = blockIdx.y;
    const int tiles_per_comp = a.nx / W;
    const int ntiles = 2 * tiles_per_comp;
    // KM = 2: inverse transform only (the boundary state of per-step energy tracking on its way back to real space):
    // the staged tile goes from S to the registers, S is refilled at once, the transform exchanges through X
    static_assert(KM != 2 || XSPLIT == 1, "the inverse-only pass exchanges through the real image");
    const bool any_k = KM != 2 && (a.has_a || a.has_b);

    // stage tile `t` of this trajectory into S (one elected thread) 
    auto stage = [&](int t) {
        sgpe_mbar_expect_tx(mbar, TILE_BYTES);
#pragma unroll 1
        for (int q = 0; q < NBOX; q++)
            sgpe_tma_load_2d(S + (size_t)q * BOXR * W, &tmap, 2 * (t % tiles_per_comp) * W,
                             (b * 2 + t / tiles_per_comp) * a.ny + q * BOXR, mbar, BOXR, (int)(W * sizeof(C)));
    };

    if (threadIdx.x == 0) sgpe_mbar_init(mbar, 1);
    if (KM == 2 && a.zero2 != nullptr && blockIdx.x == 0 && threadIdx.x < 2) a.zero2[2 * b + threadIdx.x] = 0ull;
    if (TWS == 1) {
        const C* src = a.tw + (E == 16 ? N : 0);
        for (int i = threadIdx.x; i < N; i += W * NT) TW[i] = __ldg(&src[i]);
    }
    __syncthreads();
    if (threadIdx.x == 0) stage(blockIdx.x);
    int done = 0;
    int kt_comp = -1;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, done++) {
        // (thread coordinates are re-derived per tile: nothing that depends on them stays live across the loop)
        const int tid = threadIdx.x;
        const int c = tid % W, j = tid / W;
        typename TwSource<C, TWS == 1>::type tw = TwSource<C, TWS == 1>::make(a.tw + (E == 16 ? N : 0), TW);
        const int comp = tile / tiles_per_comp;
        // the x-dependent factors of this tile's W columns: fetched by W threads NOW, read by everybody in the K phase
        // (behind the barriers of the forward transform) - as plain loads at the head of the K phase they stalled the
        // whole CTA for an L2 round trip per tile
        C* const FX = reinterpret_cast<C*>(mbar + 16);           // [2][W], behind the barrier word, flag and stamps
        if (KM == 1 && any_k && tid < 2 * W) {
            const int which = tid / W, cc = tid % W;
            const long long oxx = (long long)b * a.sepx_bstride + (long long)comp * a.nx + (tile % tiles_per_comp) * W + cc;
            C one; one.x = (T)1; one.y = (T)0;
            FX[which * W + cc] = which == 0 ? (a.has_a ? __ldg(&a.xa[oxx]) : one) : (a.has_b ? __ldg(&a.xb[oxx]) : one);
        }
        if (TWS == 2 && any_k && comp != kt_comp) {
            // (the barriers of the forward transform order these stores before the K phase reads them, those of the
            // previous tile's inverse transform ordered its reads before them)
            kt_comp = comp;
            const long long oyb = (long long)b * a.sepy_bstride + (long long)comp * a.ny;
            for (int i = tid; i < N; i += W * NT) {
                KT[i] = a.has_a ? __ldg(&a.ya[oyb + i]).x : (T)1;
                KT[N + i] = a.has_b ? __ldg(&a.yb[oyb + i]).x : (T)1;
            }
        }
        const int col = (tile % tiles_per_comp) * W + c;
        const long long off = ((long long)b * 2 + comp) * a.plane + col;

#ifdef SGPE_TIMELINE     // dev build only (tools/timeline.py): phase timestamps of every tile
        // (stamps go to shared memory and are flushed once per tile: no 64-bit address stays live across the pass)
        unsigned long long* const marks = reinterpret_cast<unsigned long long*>(flag + 2);
#define SGPE_PMARK(k) do { if (threadIdx.x == 0) marks[k] = SGPE_GLOBALTIMER(); } while (0)
#else
#define SGPE_PMARK(k) do { } while (0)
#endif
        SGPE_PMARK(0);
        sgpe_mbar_wait(mbar, (unsigned)(done & 1));
        SGPE_PMARK(1);
        C v[1][E];
#pragma unroll
        for (int m = 0; m < E; m++) v[0][m] = S[(j + m * NT) * W + c];

        // forward transform through S (first stage by hand: S must be read by everybody before it is overwritten)
        C* const sms[1] = {S};
        if constexpr (KM != 2) {
            stage_compute<T, N, E, -1, 1>(v[0], j, tw);
            __syncthreads();
            if constexpr (R0 < N) {
                stage_store<T, N, E, W, 1>(v[0], j, c, S);
                __syncthreads();
                stage_load<T, N, E, W>(v[0], j, c, S);
                __syncthreads();
                cta_fft_from<T, N, E, -1, W, 1, R0>(v, j, c, sms, tw, CtaBar());
            }
        } else {
            __syncthreads();              // everybody has read the staged tile
        }
        if (XSPLIT && tid == 0 && tile + (int)gridDim.x < ntiles) {     // S is free: its last readers are behind a barrier
            sgpe_fence_proxy_async();
            stage(tile + (int)gridDim.x);
        }

        SGPE_PMARK(2);
        double acc[2] = {0.0, 0.0};   // S (after FA), T (after FB)
        if (any_k) {
            const long long ox = (long long)b * a.sepx_bstride + (long long)comp * a.nx + col;
            const long long oy = (long long)b * a.sepy_bstride + (long long)comp * a.ny + j;
            C fxa, fxb;
            fxa.x = (T)1; fxa.y = (T)0; fxb = fxa;
            if (KM != 0) { fxa = FX[c]; fxb = FX[W + c]; }
            (void)ox;
            C* const aux = a.aux != nullptr ? a.aux + off + (long long)j * a.nx : nullptr;
            const long long aux_stride = (long long)NT * a.nx;
            if (KM == 0) {                  // dense kinetic grids, stored (shifted) k order like the state
                const double* kin = (comp == 0 ? a.kin0 : a.kin1) + (long long)b * a.kin_bstride + col + (long long)j * a.nx;
                if (a.has_a && a.has_b) {
                    if (aux != nullptr) k_factors_dense<T, E, TM, true, true, true>(v[0], kin, aux_stride, a.ka_re, a.ka_im, a.kb_re, a.kb_im, aux, aux_stride, acc);
                    else k_factors_dense<T, E, TM, true, true, false>(v[0], kin, aux_stride, a.ka_re, a.ka_im, a.kb_re, a.kb_im, aux, aux_stride, acc);
                } else if (a.has_a) {
                    if (aux != nullptr) k_factors_dense<T, E, TM, true, false, true>(v[0], kin, aux_stride, a.ka_re, a.ka_im, a.kb_re, a.kb_im, aux, aux_stride, acc);
                    else k_factors_dense<T, E, TM, true, false, false>(v[0], kin, aux_stride, a.ka_re, a.ka_im, a.kb_re, a.kb_im, aux, aux_stride, acc);
                } else {
                    k_factors_dense<T, E, TM, false, true, false>(v[0], kin, aux_stride, a.ka_re, a.ka_im, a.kb_re, a.kb_im, aux, aux_stride, acc);
                }
            } else if (TWS == 2) {
                const T* const ka = KT + j; const T* const kb = KT + N + j;
                if (a.has_a && a.has_b) {
                    if (aux != nullptr) k_factors_smem_imag<T, E, NT, true, true, true>(v[0], fxa.x, fxb.x, ka, kb, aux, aux_stride, acc);
                    else k_factors_smem_imag<T, E, NT, true, true, false>(v[0], fxa.x, fxb.x, ka, kb, aux, aux_stride, acc);
                } else if (a.has_a) {
                    if (aux != nullptr) k_factors_smem_imag<T, E, NT, true, false, true>(v[0], fxa.x, fxb.x, ka, kb, aux, aux_stride, acc);
                    else k_factors_smem_imag<T, E, NT, true, false, false>(v[0], fxa.x, fxb.x, ka, kb, aux, aux_stride, acc);
                } else {
                    k_factors_smem_imag<T, E, NT, false, true, false>(v[0], fxa.x, fxb.x, ka, kb, aux, aux_stride, acc);
                }
            } else if (a.has_a && a.has_b) {
                if (aux != nullptr) k_factors<T, E, NT, TM, true, true, true>(v[0], fxa, fxb, a.ya + oy, a.yb + oy, aux, aux_stride, acc);
                else k_factors<T, E, NT, TM, true, true, false>(v[0], fxa, fxb, a.ya + oy, a.yb + oy, aux, aux_stride, acc);
            } else if (a.has_a) {
                if (aux != nullptr) k_factors<T, E, NT, TM, true, false, true>(v[0], fxa, fxb, a.ya + oy, a.yb, aux, aux_stride, acc);
                else k_factors<T, E, NT, TM, true, false, false>(v[0], fxa, fxb, a.ya + oy, a.yb, aux, aux_stride, acc);
            } else {
                k_factors<T, E, NT, TM, false, true, false>(v[0], fxa, fxb, a.ya, a.yb + oy, aux, aux_stride, acc);
            }
            cta_reduce<2>(acc, red);
            if (tid == 0) {
                double* p = a.partials + ((long long)b * ntiles + tile) * 2;
                p[0] = acc[0]; p[1] = acc[1];
            }
        }

        SGPE_PMARK(3);
        if constexpr (XSPLIT) {
            T* const xs[1] = {X};
            cta_fft_split_from<T, N, E, +1, W, 1, 1>(v, j, c, xs, tw, CtaBar());
        } else {
            // (fresh copies of the thread coordinates: the exchange addresses of the two transforms are NOT shared, which
            // would keep 32 of them live across the pass and spill)
            int j2 = j, c2 = c;
            SGPE_OPAQUE(j2); SGPE_OPAQUE(c2);
            cta_fft_head<T, N, E, +1, W, 1, 1>(v, j2, c2, sms, tw, CtaBar());
            if (tid == 0 && tile + (int)gridDim.x < ntiles) {           // behind the last exchange read: S is free
                sgpe_fence_proxy_async();
                stage(tile + (int)gridDim.x);
            }
            stage_compute<T, N, E, +1, LAST_NS>(v[0], j2, tw);
        }
        SGPE_PMARK(4);
#pragma unroll
        for (int m = 0; m < E; m++) SGPE_ST_STREAM(&a.out[off + (long long)(j + m * NT) * a.nx], v[0][m]);
        SGPE_PMARK(5);
#ifdef SGPE_TIMELINE
        if (a.dbg != nullptr && threadIdx.x == 0) {
            unsigned long long* d = a.dbg + ((long long)b * ntiles + tile) * 8;
            for (int k = 0; k < 6; k++) d[k] = marks[k];
            d[7] = SGPE_SMID();
        }
#endif
#undef SGPE_PMARK
    }

    if (any_k) {
        const int tid = threadIdx.x;
        if (tid == 0) {
            __threadfence();
            const unsigned before = atomicAdd(&a.counter[b], (unsigned)done);
            *flag = (before + (unsigned)done == (unsigned)ntiles) ? 1 : 0;
        }
        __syncthreads();
        if (*flag) {      // the CTA that completed the count folds the partials in a fixed order
            __threadfence();
            double t4[4] = {0.0, 0.0, 0.0, 0.0};     // S0, T0, S1, T1
            const double* p = a.partials + (long long)b * ntiles * 2;
            for (int t = tid; t < ntiles; t += W * NT) {
                const int cp = (t >= ntiles / 2) ? 2 : 0;
                t4[cp + 0] += __ldcg(&p[2 * t]);
                t4[cp + 1] += __ldcg(&p[2 * t + 1]);
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
}

// Persistent column pass, two barrier groups per CTA.  The staged tile of W columns (64-byte rows, one TMA landing zone
// S as in col_pass_p) is worked on by TWO independent groups of half the threads, W / 2 columns each, with their own
// named barrier, their own real exchange image (both transforms use the split re / im exchange) and their own
// partial-sum slot.  The groups drift out of phase - the scheduler prefers the higher warp ids, so one group's
// butterflies run while the other waits at its exchange - which overlaps the FP64 pipe with the shared-memory pipe
// inside one SM; the register file (one tile of complex128 data) has no room for a second CTA to do that.  The tile is
// loaded by TMA whatever the group width, so the groups' 32-byte row halves cost nothing on the load side.
// The second group to have read tile t out of S stages tile t + gridDim.x (shared-memory ticket).
template <typename T, int N, int E, int W, int TM>
__global__ void __launch_bounds__(W * N / E, 1)
col_pass_pg(const SGPE_GRID_CONSTANT SgpeTileMap tmap, ColArgs<T> a) {
    typedef typename cx_of<T>::type C;
    constexpr int NT = N / E;
    constexpr int WG = W / 2;                   // columns per group
    constexpr int TG = WG * NT;                 // threads per group
    constexpr int BOXR = N < 256 ? N : 256, NBOX = N / BOXR;
    constexpr unsigned TILE_BYTES = (unsigned)(N * W * sizeof(C));
    SGPE_DYN_SMEM_128(smem_p);
    C* const S = reinterpret_cast<C*>(smem_p);
    double* const red0 = reinterpret_cast<double*>(smem_p + TILE_BYTES + TILE_BYTES / 2);
    SgpeMbar* const mbar = reinterpret_cast<SgpeMbar*>(red0 + 2 * 32 * 4);
    unsigned* const ticket = reinterpret_cast<unsigned*>(mbar + 1);
    int* const flag = reinterpret_cast<int*>(ticket + 1);           // [2]

    const int b = blockIdx.y;
    const int tiles_per_comp = a.nx / W;
    const int ntiles = 2 * tiles_per_comp;
    const int nslots = 2 * ntiles;
    const bool any_k = a.has_a || a.has_b;

    auto stage = [&](int t) {
        sgpe_mbar_expect_tx(mbar, TILE_BYTES);
#pragma unroll 1
        for (int q = 0; q < NBOX; q++)
            sgpe_tma_load_2d(S + (size_t)q * BOXR * W, &tmap, 2 * (t % tiles_per_comp) * W,
                             (b * 2 + t / tiles_per_comp) * a.ny + q * BOXR, mbar, BOXR, (int)(W * sizeof(C)));
    };

    if (threadIdx.x == 0) { sgpe_mbar_init(mbar, 1); *ticket = 0u; }
    __syncthreads();
    if (threadIdx.x == 0) stage(blockIdx.x);
    const int g = (int)threadIdx.x / TG;
    const GroupBar gbar = {1 + g, TG};
    T* const X = reinterpret_cast<T*>(smem_p + TILE_BYTES) + (size_t)g * N * WG;
    double* const red = red0 + g * (32 * 4);
    int done = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, done++) {
        const int tid = (int)threadIdx.x - g * TG;
        const int c = tid % WG, j = tid / WG;
        const C* __restrict__ tw = a.tw + (E == 16 ? N : 0);
        const int comp = tile / tiles_per_comp;
        const int col = (tile % tiles_per_comp) * W + g * WG + c;
        const long long off = ((long long)b * 2 + comp) * a.plane + col;

        sgpe_mbar_wait(mbar, (unsigned)(done & 1));
        C v[1][E];
#pragma unroll
        for (int m = 0; m < E; m++) v[0][m] = S[(j + m * NT) * W + g * WG + c];
        stage_compute<T, N, E, -1, 1>(v[0], j, tw);
        gbar.sync();                        // this group has its half of the tile in registers
        if (tid == 0) {
            const unsigned t0 = atomicAdd(ticket, 1u);
            if ((t0 & 1u) && tile + (int)gridDim.x < ntiles) {      // the other group was first: S is free
                sgpe_fence_proxy_async();
                stage(tile + (int)gridDim.x);
            }
        }
        T* const xs[1] = {X};
        {   // forward transform: the stage-1 butterflies are done, continue with the exchange
            constexpr int R0 = E;
            if constexpr (R0 < N) {
                stage_store_part<T, N, E, WG, 1, 0>(v[0], j, c, X);
                gbar.sync();
                stage_load_part<T, N, E, WG, 0>(v[0], j, c, X);
                gbar.sync();
                stage_store_part<T, N, E, WG, 1, 1>(v[0], j, c, X);
                gbar.sync();
                stage_load_part<T, N, E, WG, 1>(v[0], j, c, X);
                cta_fft_split_from<T, N, E, -1, WG, 1, R0>(v, j, c, xs, tw, gbar);
            }
        }

        double acc[2] = {0.0, 0.0};   // S (after FA), T (after FB)
        if (any_k) {
            const long long ox = (long long)b * a.sepx_bstride + (long long)comp * a.nx + col;
            const long long oy = (long long)b * a.sepy_bstride + (long long)comp * a.ny + j;
            C fxa, fxb;
            fxa.x = (T)1; fxa.y = (T)0; fxb = fxa;
            if (a.has_a) fxa = __ldg(&a.xa[ox]);
            if (a.has_b) fxb = __ldg(&a.xb[ox]);
            C* const aux = a.aux != nullptr ? a.aux + off + (long long)j * a.nx : nullptr;
            const long long aux_stride = (long long)NT * a.nx;
            if (a.has_a && a.has_b) {
                if (aux != nullptr) k_factors<T, E, NT, TM, true, true, true>(v[0], fxa, fxb, a.ya + oy, a.yb + oy, aux, aux_stride, acc);
                else k_factors<T, E, NT, TM, true, true, false>(v[0], fxa, fxb, a.ya + oy, a.yb + oy, aux, aux_stride, acc);
            } else if (a.has_a) {
                if (aux != nullptr) k_factors<T, E, NT, TM, true, false, true>(v[0], fxa, fxb, a.ya + oy, a.yb, aux, aux_stride, acc);
                else k_factors<T, E, NT, TM, true, false, false>(v[0], fxa, fxb, a.ya + oy, a.yb, aux, aux_stride, acc);
            } else {
                k_factors<T, E, NT, TM, false, true, false>(v[0], fxa, fxb, a.ya, a.yb + oy, aux, aux_stride, acc);
            }
            group_reduce<2>(acc, red, tid, TG, gbar);
            if (tid == 0) {
                double* p = a.partials + ((long long)b * nslots + 2 * tile + g) * 2;
                p[0] = acc[0]; p[1] = acc[1];
            }
        }

        cta_fft_split_from<T, N, E, +1, WG, 1, 1>(v, j, c, xs, tw, gbar);
#pragma unroll
        for (int m = 0; m < E; m++) SGPE_ST_STREAM(&a.out[off + (long long)(j + m * NT) * a.nx], v[0][m]);
    }

    if (any_k) {
        const int tid = (int)threadIdx.x - g * TG;
        if (tid == 0) {
            __threadfence();
            const unsigned before = atomicAdd(&a.counter[b], (unsigned)done);
            flag[g] = (before + (unsigned)done == (unsigned)nslots) ? 1 : 0;
        }
        gbar.sync();
        if (flag[g]) {    // the group that completed the count folds the partials in a fixed order
            __threadfence();
            double t4[4] = {0.0, 0.0, 0.0, 0.0};     // S0, T0, S1, T1
            const double* p = a.partials + (long long)b * nslots * 2;
            for (int t = tid; t < nslots; t += TG) {
                const int cp = (t >= nslots / 2) ? 2 : 0;
                t4[cp + 0] += __ldcg(&p[2 * t]);
                t4[cp + 1] += __ldcg(&p[2 * t + 1]);
            }
            group_reduce<4>(t4, red, tid, TG, gbar);
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
}

template <typename T> struct RowArgs {
    typedef typename cx_of<T>::type C;
    const C* in;  C* out;          // [B][2][ny][nx]
    const C* tw;                   // per-stage twiddle tables for nx (same layout)
    int nx, ny; long long plane;
    int do_inv, do_pw, do_fwd;
    int prefetch_ahead;            // > 0: L2 prefetch of the rows this many CTAs ahead
    int resident; int stagger_ns;  // de-phase the CTAs that share an SM: the 2nd arrival on each SM starts late
    unsigned* sm_slots;            // [>= #SMs] arrival counters, zeroed before the launch
    unsigned long long* dbg;       // dev tool: per-CTA phase timestamps [nCTA][8] (null in production)
    int sign_in, sign_out; double scale_out;   // sign bit 1: (-1)^x, bit 2: (-1)^y
    const double* pot0; const double* pot1; long long pot_bstride;     // [ny][nx]
    int pot_mode;                  // 0: dense grids, 1: separable factor tables px[2][nx], py[2][ny]
    const C* px; const C* py; long long sepx_bstride, sepy_bstride;
    int cpl_mode;                  // 0: none, 1: uniform (omega_b[b]), 2: dense (coupling[ny][nx])
    const double* coupling; long long cpl_bstride;
    const double* omega_b;         // [B]
    const C* eiphi;                // [nx] exp(+i*expon) or null (rotating frame)
    double g_uu, g_dd, g_ud;
    double ti_re, ti_im;           // interaction time argument (dt_sub / 2)
    double tp_re, tp_im;           // potential time argument (dt_sub)
    double tc;                     // coupling angle per unit Omega (|dt_sub| / 4)
    const double* totals;          // [B][4], T at [0]
    double norm_c;                 // N_atoms / (dv_r * nx * ny)
    Scatter<C> sc;                 // slab mode: fused exchange on the store (non-FAST kernels only)
    // non-FAST kernels, stand-alone inverse of an un-normalised k-space state (per-step energy tracking): the output
    // is also multiplied by sqrt(scale_num / (scale_tot[b][1] + scale_tot[b][2])) — ttools.norm with the sums the
    // junction pass left on the device — and the per-component maxima of |out|^2 are folded into maxbits[b][2]
    // (bits of non-negative doubles: integer max == floating-point max, order independent; zeroed by the caller)
    const double* scale_tot; double scale_num;
    unsigned long long* maxbits;
    int polar;                     // store (sqrt(re^2 + im^2), atan2(im, re)) instead of (re, im): input of the energy stencils
};

// (|z|, arg z) packed as a complex number
template <typename T, typename C> SGPE_DI C to_polar(C z) {
    C o;
    if constexpr (sizeof(T) == 8) { o.x = sgpe_sqrt(z.x * z.x + z.y * z.y); o.y = sgpe_atan2(z.y, z.x); }
    else { o.x = (T)sqrt(z.x * z.x + z.y * z.y); o.y = (T)atan2(z.y, z.x); }
    return o;
}

// 2x2 coupling operator (reference tensor_tools.py:586-590) for theta = Omega*tc and exp(i phi) = ph
template <int TM, typename T, typename C>
SGPE_DI void coupling_entries(double theta, C ph, T& diag, C& off01, C& off10) {
    if (TM == TM_REAL) {
        double s, c;
        sincos(theta, &s, &c);
        diag = (T)c;
        const T sn = (T)s;
        off01.x = -sn * ph.y; off01.y = -sn * ph.x;     // -i sin(theta) e^{-i phi}
        off10.x =  sn * ph.y; off10.y = -sn * ph.x;     // -i sin(theta) e^{+i phi}
    } else {
        diag = (T)cosh(theta);
        const T sh = (T)sinh(theta);
        off01.x = -sh * ph.x; off01.y =  sh * ph.y;     // -sinh(theta) e^{-i phi}
        off10.x = -sh * ph.x; off10.y = -sh * ph.y;     // -sinh(theta) e^{+i phi}
    }
}

// FAST = 1 is the specialisation for the common full pass (inverse + point-wise + forward, no coupling,
// separable potential, no sign / scale): same code with the other branches compiled out, which roughly halves
// the instruction footprint (the generic kernel does not fit the instruction cache: ncu stall_no_inst).
// resident CTAs asked of the compiler: complex128 -> 128 registers per thread (64 hold the two components' data),
// complex64 -> 64 registers per thread (twice the CTAs per SM: the FP32 passes are latency-, not register-bound)
#ifndef SGPE_F32_THREADS_PER_SM
#define SGPE_F32_THREADS_PER_SM 768
#endif
#ifndef SGPE_INVP_BLOCKS
#define SGPE_INVP_BLOCKS 2
#endif
template <typename T, int E = 8> constexpr int row_min_blocks(int threads, int fast = 0) {
    // (complex64 with 16 elements per thread holds 64 data registers: 512 threads, 128 registers each)
    // (SGPE_INVP_BLOCKS: resident CTAs asked for the inverse-only polar pass, whose point-wise phase holds no operators)
    return sizeof(T) == 8 ? (threads <= 256 ? ((fast == 3 && threads == 256) ? SGPE_INVP_BLOCKS : 2) : 1)
         : (E == 16 ? (threads <= 512 ? 512 / threads : 1)
                    : (threads <= SGPE_F32_THREADS_PER_SM ? SGPE_F32_THREADS_PER_SM / threads : 1));
}
template <typename T, int N, int E, int RPC, int TM, int FAST>
__global__ void __launch_bounds__(RPC * N / E, row_min_blocks<T, E>(RPC * N / E, FAST)) row_pass(RowArgs<T> a) {
    typedef typename cx_of<T>::type C;
    constexpr int NT = N / E;
    SGPE_DYN_SMEM(smem_raw);
    C* smem = reinterpret_cast<C*>(smem_raw);

    const int tid = threadIdx.x;
    const int r = tid / NT, j = tid % NT;
    const int y = blockIdx.x * RPC + r;
    const int b = blockIdx.y;
    // FAST = 2: the same specialisation for DENSE potential grids (factors evaluated per point, no coupling)
    // FAST = 3: the stand-alone inverse of per-step energy tracking (inverse transform only, device-side scale, density
    // maxima, sign on the store, POLAR output: (sqrt(n), atan2) instead of (re, im) - see RowArgs::polar)
    constexpr bool GEN = (FAST == 0), INVP = (FAST == 3);
    const int cpl_mode = GEN ? a.cpl_mode : 0, pot_mode = FAST == 1 ? 1 : (FAST == 2 ? 0 : a.pot_mode);
    const int sign_in = GEN ? a.sign_in : 0, sign_out = (GEN || INVP) ? a.sign_out : 0;
    const bool do_inv = GEN ? (a.do_inv != 0) : true, do_pw = GEN ? (a.do_pw != 0) : !INVP;
    const bool do_fwd = GEN ? (a.do_fwd != 0) : !INVP;
    // everything the point-wise phase needs from memory is pulled into L1 before the transforms start
    // (prefetches cost no registers; holding the values across the FFT made the kernel spill)
    if (do_pw) {
        SGPE_PREFETCH_L1(&a.totals[(long long)b * 4]);
        if (pot_mode == 1) {
            const long long oy = (long long)b * a.sepy_bstride + y;
            SGPE_PREFETCH_L1(&a.py[oy]);
            SGPE_PREFETCH_L1(&a.py[oy + a.ny]);
        }
    }
    if (GEN && a.stagger_ns > 0) {
        // co-resident CTAs launched together run in lock-step (same phase -> they fight for the same unit);
        // delaying the second arrival on every SM by ~half a CTA lifetime interleaves their phases for good
        if (tid == 0 && (atomicAdd(&a.sm_slots[SGPE_SMID()], 1u) == 1u)) SGPE_NANOSLEEP((unsigned)a.stagger_ns);
        __syncthreads();
    }
    const long long off0 = ((long long)b * 2) * a.plane + (long long)y * a.nx;
    const long long off1 = off0 + a.plane;
#define SGPE_MARK(k) do { if (GEN && a.dbg != nullptr && tid == 0) a.dbg[(long long)blockIdx.x * 8 + (k)] = SGPE_GLOBALTIMER(); } while (0)
    if (GEN && a.dbg != nullptr && tid == 0) a.dbg[(long long)blockIdx.x * 8 + 7] = SGPE_SMID();
    SGPE_MARK(0);

    C v[2][E];
#pragma unroll
    for (int m = 0; m < E; m++) {
        v[0][m] = SGPE_LD_STREAM(&a.in[off0 + j + m * NT]);
        v[1][m] = SGPE_LD_STREAM(&a.in[off1 + j + m * NT]);
    }
    if (a.prefetch_ahead > 0) {
        const int ny2 = (blockIdx.x + a.prefetch_ahead) * RPC + r;
        if (ny2 < a.ny) {
            constexpr int PER_LINE = 128 / (int)sizeof(C);      // elements per 128-byte line
            const long long n0 = ((long long)b * 2) * a.plane + (long long)ny2 * a.nx;
            for (int x = j * PER_LINE; x < N; x += NT * PER_LINE) {
                SGPE_PREFETCH_L2(&a.in[n0 + x]);
                SGPE_PREFETCH_L2(&a.in[n0 + a.plane + x]);
            }
        }
    }
    if (sign_in) {
#pragma unroll
        for (int m = 0; m < E; m++) {
            if ((((sign_in & 1) ? (j + m * NT) : 0) + ((sign_in & 2) ? y : 0)) & 1) {
                v[0][m].x = -v[0][m].x; v[0][m].y = -v[0][m].y;
                v[1][m].x = -v[1][m].x; v[1][m].y = -v[1][m].y;
            }
        }
    }
    C* const sms[2] = {smem + (size_t)(2 * r) * N, smem + (size_t)(2 * r + 1) * N};
    if (GEN && a.dbg != nullptr) { if (v[0][0].x == (T)1.2345e300 || v[1][E - 1].y == (T)1.2345e300) a.dbg[6] = 1; SGPE_MARK(1); }

#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
    if (pass == 0 ? do_inv : do_fwd) {
        if (pass == 0) { conj_all(v[0]); conj_all(v[1]); }
        cta_fft<T, N, E, -1, 1, 2>(v, j, 0, sms, a.tw + (E == 16 ? N : 0));
        if (pass == 0) { conj_all(v[0]); conj_all(v[1]); }
    }
    if (pass == 1) break;
    SGPE_MARK(2);

    if (do_pw) {
        const T alpha = (T)sqrt(a.norm_c / a.totals[(long long)b * 4]);
        C py0, py1;
        py0.x = (T)1; py0.y = (T)0; py1 = py0;
        if (pot_mode == 1) {
            const long long oy = (long long)b * a.sepy_bstride + y;
            py0 = __ldg(&a.py[oy]);
            py1 = __ldg(&a.py[oy + a.ny]);
        }
        const long long prow = (long long)b * a.pot_bstride + (long long)y * a.nx;
        const bool same_pot = (a.pot0 == a.pot1);
#ifndef SGPE_NO_FUSED_IPI
        if constexpr (FAST == 1 || FAST == 2) {
            // Without coupling the operators of a sub-step commute: I(dt/2) P(dt) I(dt/2) (tensor_propagator.py:249-267) is ONE
            // diagonal factor per component, alpha * P * exp(-i e_int dt) with e_int evaluated on the un-scaled |v|^2 and
            // alpha^2 folded into the couplings - per pixel pair 60 instead of 72 FP64 instructions in imaginary time (one
            // transcendental per component as before: I^2 = exp of the doubled argument), rounding-level differences only.
            const double a2 = (double)alpha * (double)alpha;
            const double c_uu = a.g_uu * a2, c_dd = a.g_dd * a2, c_ud = a.g_ud * a2;
            const double t2r = 2.0 * a.ti_re, t2i = 2.0 * a.ti_im;
            const C apy0 = cscale(py0, alpha), apy1 = cscale(py1, alpha);
#pragma unroll
            for (int m = 0; m < E; m++) {
                const int x = j + m * NT;
                const C p = v[0][m], q = v[1][m];
                C e0, e1;
                if constexpr (sizeof(T) == 4 && TM == TM_IMAG) {
                    // complex64, imaginary time: the decay exponent in single precision throughout (its rounding, 1e-7 of
                    // an exponent of order 1e-2, is far below the state's own precision; a PHASE of real time keeps the
                    // double-precision argument) - the FP64 pipe and the conversions leave the issue-bound pass
                    const float m0 = fmaf(p.x, p.x, p.y * p.y), m1 = fmaf(q.x, q.x, q.y * q.y);
                    e0.x = expf(fmaf((float)(c_uu * t2i), m0, (float)(c_ud * t2i) * m1)); e0.y = 0.f;
                    e1.x = expf(fmaf((float)(c_dd * t2i), m1, (float)(c_ud * t2i) * m0)); e1.y = 0.f;
                } else {
                    const double m0 = (double)p.x * p.x + (double)p.y * p.y;
                    const double m1 = (double)q.x * q.x + (double)q.y * q.y;
                    e0 = evo<TM, T, C>(c_uu * m0 + c_ud * m1, t2r, t2i);
                    e1 = evo<TM, T, C>(c_dd * m1 + c_ud * m0, t2r, t2i);
                }
                C f0, f1;
                if (FAST == 2) {
                    f0 = cscale(evo<TM, T, C>(__ldg(&a.pot0[prow + x]), a.tp_re, a.tp_im), alpha);
                    f1 = same_pot ? f0 : cscale(evo<TM, T, C>(__ldg(&a.pot1[prow + x]), a.tp_re, a.tp_im), alpha);
                } else {
                    const long long ox = (long long)b * a.sepx_bstride + x;
                    f0 = combine_factor<TM>(__ldg(&a.px[ox]), apy0);
                    f1 = combine_factor<TM>(__ldg(&a.px[ox + a.nx]), apy1);
                }
                v[0][m] = mul_factor<TM>(p, combine_factor<TM>(f0, e0));
                v[1][m] = mul_factor<TM>(q, combine_factor<TM>(f1, e1));
            }
        } else {
#else
        {
#endif
        T cu_diag = (T)1, cu_s = (T)0;          // uniform coupling: cos/sin (cosh/sinh) once per thread
        if (cpl_mode == 1) {
            C one; one.x = (T)1; one.y = (T)0;
            C t01, t10;
            coupling_entries<TM, T, C>(a.omega_b[b] * a.tc, one, cu_diag, t01, t10);
            cu_s = (TM == TM_REAL) ? -t01.y : -t01.x;       // sin(theta) resp. sinh(theta)
        }
#pragma unroll
        for (int m = 0; m < E; m++) {
            const int x = j + m * NT;
            C p = cscale(v[0][m], alpha), q = cscale(v[1][m], alpha);
            const double n0 = (double)p.x * p.x + (double)p.y * p.y;
            const double n1 = (double)q.x * q.x + (double)q.y * q.y;
            // (a short polynomial for small arguments - sgpe_exp_pair - was measured SLOWER here: 107 vs 103 us, the second
            // code path costs more in registers and instruction cache than the nine saved FMAs per factor return)
            const C i0 = evo<TM, T, C>(a.g_uu * n0 + a.g_ud * n1, a.ti_re, a.ti_im);
            const C i1 = evo<TM, T, C>(a.g_dd * n1 + a.g_ud * n0, a.ti_re, a.ti_im);
            p = mul_factor<TM>(p, i0); q = mul_factor<TM>(q, i1);
            T diag = (T)1; C o01, o10;
            if (cpl_mode) {
                C ph; ph.x = (T)1; ph.y = (T)0;
                if (a.eiphi != nullptr) ph = __ldg(&a.eiphi[x]);
                if (cpl_mode == 1) {
                    diag = cu_diag;
                    if (TM == TM_REAL) {
                        o01.x = -cu_s * ph.y; o01.y = -cu_s * ph.x; o10.x = cu_s * ph.y; o10.y = -cu_s * ph.x;
                    } else {
                        o01.x = -cu_s * ph.x; o01.y = cu_s * ph.y; o10.x = -cu_s * ph.x; o10.y = -cu_s * ph.y;
                    }
                } else {
                    const double om = __ldg(&a.coupling[(long long)b * a.cpl_bstride + (long long)y * a.nx + x]);
                    coupling_entries<TM, T, C>(om * a.tc, ph, diag, o01, o10);
                }
                const C p2 = cadd(cscale(p, diag), cmul(o01, q));
                const C q2 = cadd(cmul(o10, p), cscale(q, diag));
                p = p2; q = q2;
            }
            C f0, f1;
            if (pot_mode == 0) {
                f0 = evo<TM, T, C>(__ldg(&a.pot0[prow + x]), a.tp_re, a.tp_im);
                f1 = same_pot ? f0 : evo<TM, T, C>(__ldg(&a.pot1[prow + x]), a.tp_re, a.tp_im);
            } else {
                const long long ox = (long long)b * a.sepx_bstride + x;
                f0 = combine_factor<TM>(__ldg(&a.px[ox]), py0);
                f1 = combine_factor<TM>(__ldg(&a.px[ox + a.nx]), py1);
            }
            p = mul_factor<TM>(p, f0); q = mul_factor<TM>(q, f1);
            if (cpl_mode) {
                const C p2 = cadd(cscale(p, diag), cmul(o01, q));
                const C q2 = cadd(cmul(o10, p), cscale(q, diag));
                p = p2; q = q2;
            }
            v[0][m] = mul_factor<TM>(p, i0); v[1][m] = mul_factor<TM>(q, i1);
        }
        }   // (sequential I C P C I)
    }

    SGPE_MARK(3);
    }   // pass loop
    SGPE_MARK(4);

    T sc = (GEN || INVP) ? (T)a.scale_out : (T)1;
    if (INVP || (GEN && a.scale_tot != nullptr)) {
        const double* tot = a.scale_tot + (long long)b * 4;
        sc = (T)((double)sc * sqrt(a.scale_num / (tot[1] + tot[2])));
    }
    if (INVP || (GEN && a.maxbits != nullptr)) {
        double mx0 = 0.0, mx1 = 0.0;
        const double s2 = (double)sc * (double)sc;
#pragma unroll
        for (int m = 0; m < E; m++) {
            const double d0 = ((double)v[0][m].x * v[0][m].x + (double)v[0][m].y * v[0][m].y) * s2;
            const double d1 = ((double)v[1][m].x * v[1][m].x + (double)v[1][m].y * v[1][m].y) * s2;
            mx0 = d0 > mx0 ? d0 : mx0; mx1 = d1 > mx1 ? d1 : mx1;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double t0 = __shfl_xor_sync(0xffffffffu, mx0, o), t1 = __shfl_xor_sync(0xffffffffu, mx1, o);
            mx0 = t0 > mx0 ? t0 : mx0; mx1 = t1 > mx1 ? t1 : mx1;
        }
        if ((tid & 31) == 0) {
            atomicMax(&a.maxbits[2 * b], (unsigned long long)__double_as_longlong(mx0));
            atomicMax(&a.maxbits[2 * b + 1], (unsigned long long)__double_as_longlong(mx1));
        }
    }
    if (GEN && a.sc.mode) {       // fused exchange: each run of `seg` columns goes to the rank that owns it
#pragma unroll
        for (int m = 0; m < E; m++) {
            SGPE_ST_STREAM(scatter_ptr(a.sc, 0, y, j + m * NT), v[0][m]);
            SGPE_ST_STREAM(scatter_ptr(a.sc, 1, y, j + m * NT), v[1][m]);
        }
    } else if (INVP || (GEN && a.polar)) {
        // polar store for the energy functional: the square roots and arctangents of eng_expect (tensor_propagator.py:
        // 306-311) are evaluated HERE, sixteen independent chains per thread beside the other CTA's memory phases,
        // instead of one dependent chain per pixel in the stencil pass
#pragma unroll
        for (int m = 0; m < E; m++) {
            T s = sc;
            if ((((sign_out & 1) ? (j + m * NT) : 0) + ((sign_out & 2) ? y : 0)) & 1) s = -s;
            SGPE_ST_STREAM(&a.out[off0 + j + m * NT], to_polar<T>(cscale(v[0][m], s)));
            SGPE_ST_STREAM(&a.out[off1 + j + m * NT], to_polar<T>(cscale(v[1][m], s)));
        }
    } else {
#pragma unroll
    for (int m = 0; m < E; m++) {
        T s = sc;
        if ((((sign_out & 1) ? (j + m * NT) : 0) + ((sign_out & 2) ? y : 0)) & 1) s = -s;
        SGPE_ST_STREAM(&a.out[off0 + j + m * NT], cscale(v[0][m], s));
        SGPE_ST_STREAM(&a.out[off1 + j + m * NT], cscale(v[1][m], s));
    }
    }
    SGPE_MARK(5);
#undef SGPE_MARK
}

// Row pass, split variant: one component per thread (twice the threads, half the registers of row_pass).
// The two threads that own the same pixels of the two components trade what the point-wise operators
// need through shared memory: the densities (for I) and, with coupling, the wavefunction (for C).
template <typename T, int N, int E, int RPC, int TM>
__global__ void __launch_bounds__(RPC * 2 * N / E, (RPC * 2 * N / E <= 512) ? 2 : 1) row_pass_split(RowArgs<T> a) {
    typedef typename cx_of<T>::type C;
    constexpr int NT = N / E;
    SGPE_DYN_SMEM(smem_raw);
    C* smem = reinterpret_cast<C*>(smem_raw);

    const int tid = threadIdx.x;
    const int r = tid / (2 * NT), comp = (tid / NT) & 1, j = tid % NT;
    const int y = blockIdx.x * RPC + r;
    const int b = blockIdx.y;
    const long long off = ((long long)b * 2 + comp) * a.plane + (long long)y * a.nx;

    C v[1][E];
#pragma unroll
    for (int m = 0; m < E; m++) v[0][m] = a.in[off + j + m * NT];
    if (a.sign_in) {
#pragma unroll
        for (int m = 0; m < E; m++) {
            if ((((a.sign_in & 1) ? (j + m * NT) : 0) + ((a.sign_in & 2) ? y : 0)) & 1) {
                v[0][m].x = -v[0][m].x; v[0][m].y = -v[0][m].y;
            }
        }
    }
    C* mine = smem + (size_t)(2 * r + comp) * N;
    C* other = smem + (size_t)(2 * r + (1 - comp)) * N;
    C* const sms[1] = {mine};

    if (a.do_inv) cta_fft<T, N, E, +1, 1, 1>(v, j, 0, sms, a.tw + (E == 16 ? N : 0));

    if (a.do_pw) {
        const T alpha = (T)sqrt(a.norm_c / a.totals[(long long)b * 4]);
        const double g_self = (comp == 0) ? a.g_uu : a.g_dd;
        double* dmine = reinterpret_cast<double*>(mine);
        const double* dother = reinterpret_cast<const double*>(other);
        double n_own[E];
#pragma unroll
        for (int m = 0; m < E; m++) {
            v[0][m] = cscale(v[0][m], alpha);
            n_own[m] = (double)v[0][m].x * v[0][m].x + (double)v[0][m].y * v[0][m].y;
            dmine[j + m * NT] = n_own[m];
        }
        __syncthreads();
        C iop[E];
#pragma unroll
        for (int m = 0; m < E; m++) {
            iop[m] = evo<TM, T, C>(g_self * n_own[m] + a.g_ud * dother[j + m * NT], a.ti_re, a.ti_im);
            v[0][m] = mul_factor<TM>(v[0][m], iop[m]);
        }
        __syncthreads();

        T cu_diag = (T)1, cu_s = (T)0;
        if (a.cpl_mode == 1) {
            C one; one.x = (T)1; one.y = (T)0;
            C t01, t10;
            coupling_entries<TM, T, C>(a.omega_b[b] * a.tc, one, cu_diag, t01, t10);
            cu_s = (TM == TM_REAL) ? -t01.y : -t01.x;
        }
        // this thread's row of the 2x2 coupling operator: new = diag*own + off*partner
        auto couple = [&]() {
#pragma unroll
            for (int m = 0; m < E; m++) mine[j + m * NT] = v[0][m];
            __syncthreads();
#pragma unroll
            for (int m = 0; m < E; m++) {
                const int x = j + m * NT;
                const C q = other[x];
                C ph; ph.x = (T)1; ph.y = (T)0;
                if (a.eiphi != nullptr) ph = __ldg(&a.eiphi[x]);
                T diag; C o01, o10;
                if (a.cpl_mode == 1) {
                    diag = cu_diag;
                    if (TM == TM_REAL) {
                        o01.x = -cu_s * ph.y; o01.y = -cu_s * ph.x; o10.x = cu_s * ph.y; o10.y = -cu_s * ph.x;
                    } else {
                        o01.x = -cu_s * ph.x; o01.y = cu_s * ph.y; o10.x = -cu_s * ph.x; o10.y = -cu_s * ph.y;
                    }
                } else {
                    const double om = __ldg(&a.coupling[(long long)b * a.cpl_bstride + (long long)y * a.nx + x]);
                    coupling_entries<TM, T, C>(om * a.tc, ph, diag, o01, o10);
                }
                v[0][m] = cadd(cscale(v[0][m], diag), cmul(comp == 0 ? o01 : o10, q));
            }
            __syncthreads();
        };
        if (a.cpl_mode) couple();

        if (a.pot_mode == 0) {
            const double* pot = (comp == 0 ? a.pot0 : a.pot1) + (long long)b * a.pot_bstride + (long long)y * a.nx;
#pragma unroll
            for (int m = 0; m < E; m++)
                v[0][m] = mul_factor<TM>(v[0][m], evo<TM, T, C>(__ldg(&pot[j + m * NT]), a.tp_re, a.tp_im));
        } else {
            const C py = __ldg(&a.py[(long long)b * a.sepy_bstride + (long long)comp * a.ny + y]);
            const C* px = a.px + (long long)b * a.sepx_bstride + (long long)comp * a.nx;
#pragma unroll
            for (int m = 0; m < E; m++)
                v[0][m] = mul_factor<TM>(v[0][m], combine_factor<TM>(__ldg(&px[j + m * NT]), py));
        }
        if (a.cpl_mode) couple();
#pragma unroll
        for (int m = 0; m < E; m++) v[0][m] = mul_factor<TM>(v[0][m], iop[m]);
    }

    if (a.do_fwd) cta_fft<T, N, E, -1, 1, 1>(v, j, 0, sms, a.tw + (E == 16 ? N : 0));

    const T sc = (T)a.scale_out;
#pragma unroll
    for (int m = 0; m < E; m++) {
        T s = sc;
        if ((((a.sign_out & 1) ? (j + m * NT) : 0) + ((a.sign_out & 2) ? y : 0)) & 1) s = -s;
        a.out[off + j + m * NT] = cscale(v[0][m], s);
    }
}

// ---------------------------------------------------------------------------------------------
// Slab (distributed) mode.  The grid is split by rows over P ranks; a 2-D transform is x-lines locally,
// an all-to-all transpose, y-lines locally.  In the transposed layout the k-space junction
// (FFT_y, K_a, sums, K_b, sums, iFFT_y) runs on CONTIGUOUS lines: kline_pass is the column pass of the
// single-GPU path re-expressed on rows (both components of one line per CTA, like row_pass).
// Array seen by this kernel: [2][nlines][N]; `nx` of the args = N (line length), `ny` = nlines.
template <typename T> struct KLineArgs {
    typedef typename cx_of<T>::type C;
    const C* in; C* out; const C* tw;
    int nx, ny; long long plane;
    int do_fwd, do_inv, has_a, has_b;
    int kin_mode;                                    // 0 dense (grids given line-major), 1 separable
    const double* kin0; const double* kin1;          // dense: [nlines][N]
    double ka_re, ka_im, kb_re, kb_im;
    const C* la; const C* pa; const C* lb; const C* pb;   // separable: per-line [2][nlines/group], per-position [2][group*N]
    int group;                                       // sub-lines per long line (1 normally): line table index =
                                                     // line/group, position table index = (line%group)*N + pos
    double* partials; unsigned* counter; double* sums;    // sums: [3] = T, S0, S1 of the LOCAL slab
    Scatter<C> sc;                                   // fused exchange on the store (mode 1: row direction)
    // window of the slab handled by this launch (chunked pipelining of the exchange, sgpe_slab_window): the lines
    // [line0, line0 + nblk * RPC); the grid may be smaller than nblk (persistent CTAs walking the window), which is
    // how a scatter launch leaves most of every SM to the passes of the next chunk running beside it
    int line0, nblk;
    int wlines, max_ctas;                            // host side: lines in the window, cap on the grid (0 = none)
};

template <typename T, int N, int E, int RPC, int TM>
__global__ void __launch_bounds__(RPC * N / E, (RPC * N / E <= 256) ? 2 : 1) kline_pass(KLineArgs<T> a) {
    typedef typename cx_of<T>::type C;
    constexpr int NT = N / E;
    SGPE_DYN_SMEM(smem_raw);
    C* smem = reinterpret_cast<C*>(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw + sizeof(C) * (size_t)RPC * 2 * N);

    const int tid = threadIdx.x;
    const int r = tid / NT, j = tid % NT;
    C* const sms[2] = {smem + (size_t)(2 * r) * N, smem + (size_t)(2 * r + 1) * N};
#pragma unroll 1
    for (int vb = blockIdx.x; vb < a.nblk; vb += gridDim.x) {
    const int line = a.line0 + vb * RPC + r;
    const long long off0 = (long long)line * a.nx, off1 = off0 + a.plane;

    C v[2][E];
#pragma unroll
    for (int m = 0; m < E; m++) {
        v[0][m] = a.in[off0 + j + m * NT];
        v[1][m] = a.in[off1 + j + m * NT];
    }
    double acc[4] = {0.0, 0.0, 0.0, 0.0};      // S0, T0, S1, T1
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
    if (pass == 0 ? a.do_fwd : a.do_inv) {
        if (pass == 1) { conj_all(v[0]); conj_all(v[1]); }
        cta_fft<T, N, E, -1, 1, 2>(v, j, 0, sms, a.tw + (E == 16 ? N : 0));
        if (pass == 1) { conj_all(v[0]); conj_all(v[1]); }
    }
    if (pass == 1) break;
    if (a.has_a || a.has_b) {
#pragma unroll
        for (int comp = 0; comp < 2; comp++) {
            C la, lb;
            la.x = (T)1; la.y = (T)0; lb = la;
            const int lgrp = line / a.group;
            const long long pbase = (long long)comp * a.group * a.nx + (long long)(line % a.group) * a.nx;
            if (a.kin_mode == 1) {
                if (a.has_a) la = __ldg(&a.la[(long long)comp * (a.ny / a.group) + lgrp]);
                if (a.has_b) lb = __ldg(&a.lb[(long long)comp * (a.ny / a.group) + lgrp]);
            }
            const double* kin = (comp == 0 ? a.kin0 : a.kin1);
#pragma unroll
            for (int m = 0; m < E; m++) {
                const int pos = j + m * NT;
                C x = v[comp][m];
                if (a.has_a) {
                    const C f = (a.kin_mode == 0) ? evo<TM, T, C>(__ldg(&kin[off0 + pos]), a.ka_re, a.ka_im)
                                                  : combine_factor<TM>(la, __ldg(&a.pa[pbase + pos]));
                    x = mul_factor<TM>(x, f);
                    acc[2 * comp] += abs_sq(x);
                }
                if (a.has_b) {
                    const C f = (a.kin_mode == 0) ? evo<TM, T, C>(__ldg(&kin[off0 + pos]), a.kb_re, a.kb_im)
                                                  : combine_factor<TM>(lb, __ldg(&a.pb[pbase + pos]));
                    x = mul_factor<TM>(x, f);
                    acc[2 * comp + 1] += abs_sq(x);
                }
                v[comp][m] = x;
            }
        }
        if (!a.has_b) { acc[1] = acc[0]; acc[3] = acc[2]; }
        if (!a.has_a) { acc[0] = acc[1]; acc[2] = acc[3]; }
    }
    }   // pass loop
    if (a.sc.mode) {      // long line = line / group (local row), position along it = (line % group) * N + pos
        const int Y = line / a.group, X0 = (line % a.group) * a.nx + j;
#pragma unroll
        for (int m = 0; m < E; m++) {
            SGPE_ST_STREAM(scatter_ptr(a.sc, 0, Y, X0 + m * NT), v[0][m]);
            SGPE_ST_STREAM(scatter_ptr(a.sc, 1, Y, X0 + m * NT), v[1][m]);
        }
    } else {
#pragma unroll
    for (int m = 0; m < E; m++) {
        a.out[off0 + j + m * NT] = v[0][m];
        a.out[off1 + j + m * NT] = v[1][m];
    }
    }
    if (a.has_a || a.has_b) {
        const int nblk = a.nblk;
        cta_reduce<4>(acc, red);
        if (tid == 0) {
            double* p = a.partials + (long long)vb * 4;
            p[0] = acc[0]; p[1] = acc[1]; p[2] = acc[2]; p[3] = acc[3];
            __threadfence();
            red[0] = (atomicAdd(&a.counter[0], 1u) == (unsigned)(nblk - 1)) ? 1.0 : 0.0;
        }
        __syncthreads();
        const bool last = red[0] != 0.0;
        __syncthreads();
        if (last) {
            __threadfence();
            double t4[4] = {0.0, 0.0, 0.0, 0.0};
            for (int t = tid; t < nblk; t += blockDim.x) {
#pragma unroll
                for (int q = 0; q < 4; q++) t4[q] += __ldcg(&a.partials[4LL * t + q]);
            }
            cta_reduce<4>(t4, red);
            if (tid == 0) {
                a.sums[0] = t4[1] + t4[3]; a.sums[1] = t4[0]; a.sums[2] = t4[2];
                a.counter[0] = 0u;
            }
        }
    }
    __syncthreads();      // the next window block reuses the exchange images
    }   // window loop
}

// ---------------------------------------------------------------------------------------------
// Lines longer than one CTA can hold (> 4096 points): four-step transform.  A line of N = N1*N2 points is the
// matrix A[n1][n2] (n = n1*N2 + n2).  forward:  FFT over n1 (stride N2)  ->  * w_N^(k1 n2)  ->  FFT over n2
// (contiguous), result X[k1 + N1 k2] left at position k1*N2 + k2 ("digit-transposed" order, which the
// propagator never needs to undo: operator tables are permuted instead).  inverse: the mirror image.
// mid_pass is the STRIDED half with the real-space operators fused in, for both components of one line:
//     [* conj w] -> [iFFT over k1] -> [normalise, I C P C I at x = n1*N2 + n2] -> [FFT over n1] -> [* w]
// i.e. for long lines the row pass becomes  contiguous-iFFT | mid_pass | contiguous-FFT  (kline_pass does the
// contiguous halves), and the k-space junction  mid_pass(fwd) | kline_pass(FFT K iFFT) | mid_pass(inv).
// Array: [2][nlines][N1][N2].  One CTA = W adjacent n2 of one line; thread (c, j) holds n1 = j + m*N1/E.
template <typename T> struct MidArgs {
    typedef typename cx_of<T>::type C;
    RowArgs<T> r;                  // operators / flags of the row pass (r.nx = N1*N2, r.ny = nlines, r.tw = N1 tables)
    int n2;                        // contiguous dimension of the matrix view
    int pre_tw, post_tw;           // multiply by conj(w_N^(k1 n2)) before the inverse / by w_N^(k1 n2) after the forward
    const C* tw4;                  // [N1][N2] four-step twiddles exp(-2 pi i k1 n2 / N)
    int inner;                     // 1: array [2][nlines][N1][N2].  > 1: k slab [2][N1][N2][inner] (row-major slab, the
                                   // lines run down the columns): r.ny = 1, n2 = N2 * inner, twiddle column = n2 / inner
    // window of the slab handled by this launch (sgpe_slab_window).  inner == 1: lines y0 .. y0 + gridDim.y.
    // inner > 1: columns [x0, x0 + wtiles * W) of every n2 digit: virtual block vb -> digit vb / wtiles, tile
    // vb % wtiles; nvb virtual blocks walked by gridDim.x (possibly fewer, persistent) CTAs.
    int y0, x0, wtiles, wstride, nvb;
    int wcount, max_ctas;                            // host side: lines / columns in the window, cap on the grid
};

template <typename T, int N1, int E, int W, int TM>
__global__ void __launch_bounds__(W * N1 / E, (W * N1 / E <= 128) ? 4 : ((W * N1 / E <= 256) ? 2 : 1)) mid_pass(MidArgs<T> ma) {
    typedef typename cx_of<T>::type C;
    const RowArgs<T>& a = ma.r;
    constexpr int NT = N1 / E;
    SGPE_DYN_SMEM(smem_raw);
    C* smem = reinterpret_cast<C*>(smem_raw);
    const int tid = threadIdx.x;
    const int c = tid % W, j = tid / W;
    C* const sms[2] = {smem, smem + (size_t)N1 * W};
#pragma unroll 1
    for (int vb = blockIdx.x; vb < ma.nvb; vb += gridDim.x) {
    const int n2 = (vb / ma.wtiles) * ma.wstride + ma.x0 + (vb % ma.wtiles) * W + c;
    const int y = ma.y0 + blockIdx.y;               // line (local row)
    const int b = 0;
    const long long line0 = (long long)y * a.nx + n2, line1 = line0 + a.plane;

    C v[2][E];
#pragma unroll
    for (int m = 0; m < E; m++) {
        const long long o = (long long)(j + m * NT) * ma.n2;
        v[0][m] = SGPE_LD_STREAM(&a.in[line0 + o]);
        v[1][m] = SGPE_LD_STREAM(&a.in[line1 + o]);
    }
    const int tw_col = n2 / ma.inner, tw_row = ma.n2 / ma.inner;     // inner == 1: (n2, N2)
    if (ma.pre_tw) {
#pragma unroll
        for (int m = 0; m < E; m++) {
            const C w = __ldg(&ma.tw4[(long long)(j + m * NT) * tw_row + tw_col]);
            v[0][m] = cmulc(v[0][m], w); v[1][m] = cmulc(v[1][m], w);
        }
    }
    if (a.do_inv) cta_fft<T, N1, E, +1, W, 2>(v, j, c, sms, a.tw + (E == 16 ? N1 : 0));

    if (a.do_pw) {
        const int cpl_mode = a.cpl_mode, pot_mode = a.pot_mode;
        const T alpha = (T)sqrt(a.norm_c / a.totals[0]);
        C py0, py1;
        py0.x = (T)1; py0.y = (T)0; py1 = py0;
        if (pot_mode == 1) { py0 = __ldg(&a.py[y]); py1 = __ldg(&a.py[y + a.ny]); }
        const long long prow = (long long)y * a.nx;
        const bool same_pot = (a.pot0 == a.pot1);
        T cu_diag = (T)1, cu_s = (T)0;
        if (cpl_mode == 1) {
            C one; one.x = (T)1; one.y = (T)0;
            C t01, t10;
            coupling_entries<TM, T, C>(a.omega_b[b] * a.tc, one, cu_diag, t01, t10);
            cu_s = (TM == TM_REAL) ? -t01.y : -t01.x;
        }
#pragma unroll
        for (int m = 0; m < E; m++) {
            const int x = (j + m * NT) * ma.n2 + n2;           // natural position along the long line
            C p = cscale(v[0][m], alpha), q = cscale(v[1][m], alpha);
            const double d0 = (double)p.x * p.x + (double)p.y * p.y;
            const double d1 = (double)q.x * q.x + (double)q.y * q.y;
            const C i0 = evo<TM, T, C>(a.g_uu * d0 + a.g_ud * d1, a.ti_re, a.ti_im);
            const C i1 = evo<TM, T, C>(a.g_dd * d1 + a.g_ud * d0, a.ti_re, a.ti_im);
            p = mul_factor<TM>(p, i0); q = mul_factor<TM>(q, i1);
            T diag = (T)1; C o01, o10;
            if (cpl_mode) {
                C ph; ph.x = (T)1; ph.y = (T)0;
                if (a.eiphi != nullptr) ph = __ldg(&a.eiphi[x]);
                if (cpl_mode == 1) {
                    diag = cu_diag;
                    if (TM == TM_REAL) {
                        o01.x = -cu_s * ph.y; o01.y = -cu_s * ph.x; o10.x = cu_s * ph.y; o10.y = -cu_s * ph.x;
                    } else {
                        o01.x = -cu_s * ph.x; o01.y = cu_s * ph.y; o10.x = -cu_s * ph.x; o10.y = -cu_s * ph.y;
                    }
                } else {
                    coupling_entries<TM, T, C>(__ldg(&a.coupling[prow + x]) * a.tc, ph, diag, o01, o10);
                }
                const C p2 = cadd(cscale(p, diag), cmul(o01, q));
                const C q2 = cadd(cmul(o10, p), cscale(q, diag));
                p = p2; q = q2;
            }
            C f0, f1;
            if (pot_mode == 0) {
                f0 = evo<TM, T, C>(__ldg(&a.pot0[prow + x]), a.tp_re, a.tp_im);
                f1 = same_pot ? f0 : evo<TM, T, C>(__ldg(&a.pot1[prow + x]), a.tp_re, a.tp_im);
            } else {
                f0 = combine_factor<TM>(__ldg(&a.px[x]), py0);
                f1 = combine_factor<TM>(__ldg(&a.px[x + a.nx]), py1);
            }
            p = mul_factor<TM>(p, f0); q = mul_factor<TM>(q, f1);
            if (cpl_mode) {
                const C p2 = cadd(cscale(p, diag), cmul(o01, q));
                const C q2 = cadd(cmul(o10, p), cscale(q, diag));
                p = p2; q = q2;
            }
            v[0][m] = mul_factor<TM>(p, i0); v[1][m] = mul_factor<TM>(q, i1);
        }
    }

    if (a.do_fwd) cta_fft<T, N1, E, -1, W, 2>(v, j, c, sms, a.tw + (E == 16 ? N1 : 0));
    if (ma.post_tw) {
#pragma unroll
        for (int m = 0; m < E; m++) {
            const C w = __ldg(&ma.tw4[(long long)(j + m * NT) * tw_row + tw_col]);
            v[0][m] = cmul(v[0][m], w); v[1][m] = cmul(v[1][m], w);
        }
    }
    if (a.sc.mode) {      // k slab -> row slabs: natural row Y = n1 * N2 + n2-digit, local column n2 % inner
        const int X = n2 - tw_col * ma.inner;
#pragma unroll
        for (int m = 0; m < E; m++) {
            const int Y = (j + m * NT) * tw_row + tw_col;
            SGPE_ST_STREAM(scatter_ptr(a.sc, 0, Y, X), v[0][m]);
            SGPE_ST_STREAM(scatter_ptr(a.sc, 1, Y, X), v[1][m]);
        }
    } else {
#pragma unroll
    for (int m = 0; m < E; m++) {
        const long long o = (long long)(j + m * NT) * ma.n2;
        SGPE_ST_STREAM(&a.out[line0 + o], v[0][m]);
        SGPE_ST_STREAM(&a.out[line1 + o], v[1][m]);
    }
    }
    __syncthreads();      // the next window block reuses the exchange images
    }   // window loop
}

// k-space junction of the slab mode on the ROW-MAJOR k slab [2][G][N][inner] (the fused-exchange layout): for W
// adjacent columns of one component and one group g (G = n1 of a four-step split, 1 otherwise):
//   [FFT over N, stride inner] -> K_a -> sums -> K_b -> sums -> [iFFT]     (col_pass of the single-GPU path with a
// group index, slab-wide sums and the scatter store).  Kinetic operator: separable tables pa/pb [2][G*N] (position
// along the line) x la/lb [2][inner] (column), or dense grids [G*N][inner] evaluated here.
template <typename T> struct KColArgs {
    typedef typename cx_of<T>::type C;
    const C* in; C* out; const C* tw;
    int inner, groups; long long plane;               // plane = G * N * inner
    int do_fwd, do_inv, has_a, has_b, kin_mode;
    const double* kin0; const double* kin1;
    double ka_re, ka_im, kb_re, kb_im;
    const C* la; const C* pa; const C* lb; const C* pb;
    double* partials; unsigned* counter; double* sums;
    Scatter<C> sc;                                    // mode 2 (k direction -> row slabs)
    int x0;                                           // first column of the window (gridDim.x * W columns)
    int wcount;                                       // host side: columns in the window
};

template <typename T, int N, int E, int W, int TM>
__global__ void __launch_bounds__(W * N / E) kcol_pass(KColArgs<T> a) {
    typedef typename cx_of<T>::type C;
    constexpr int NT = N / E;
    SGPE_DYN_SMEM(smem_raw);
    C* sm = reinterpret_cast<C*>(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw + sizeof(C) * (size_t)N * W);
    const int tid = threadIdx.x;
    const int c = tid % W, j = tid / W;
    const int col = a.x0 + blockIdx.x * W + c;
    const int g = blockIdx.y, comp = blockIdx.z;
    const long long off = (long long)comp * a.plane + (long long)g * N * a.inner + col;

    C v[1][E];
#pragma unroll
    for (int m = 0; m < E; m++) v[0][m] = SGPE_LD_STREAM(&a.in[off + (long long)(j + m * NT) * a.inner]);
    C* const sms[1] = {sm};
    if (a.do_fwd) cta_fft<T, N, E, -1, W, 1>(v, j, c, sms, a.tw + (E == 16 ? N : 0));
    double acc[2] = {0.0, 0.0};       // S (after K_a), T (after K_b) of this component
    const bool any_k = a.has_a || a.has_b;
    if (any_k) {
        C la, lb;
        la.x = (T)1; la.y = (T)0; lb = la;
        if (a.kin_mode == 1) {
            if (a.has_a) la = __ldg(&a.la[(long long)comp * a.inner + col]);
            if (a.has_b) lb = __ldg(&a.lb[(long long)comp * a.inner + col]);
        }
        const double* kin = (comp == 0 ? a.kin0 : a.kin1);
        const long long pbase = (long long)comp * a.groups * N + (long long)g * N;
#pragma unroll
        for (int m = 0; m < E; m++) {
            const int pos = j + m * NT;
            C x = v[0][m];
            if (a.has_a) {
                const C f = (a.kin_mode == 0)
                    ? evo<TM, T, C>(__ldg(&kin[((long long)g * N + pos) * a.inner + col]), a.ka_re, a.ka_im)
                    : combine_factor<TM>(la, __ldg(&a.pa[pbase + pos]));
                x = mul_factor<TM>(x, f);
                acc[0] += abs_sq(x);
            }
            if (a.has_b) {
                const C f = (a.kin_mode == 0)
                    ? evo<TM, T, C>(__ldg(&kin[((long long)g * N + pos) * a.inner + col]), a.kb_re, a.kb_im)
                    : combine_factor<TM>(lb, __ldg(&a.pb[pbase + pos]));
                x = mul_factor<TM>(x, f);
                acc[1] += abs_sq(x);
            }
            v[0][m] = x;
        }
        if (!a.has_b) acc[1] = acc[0];
        if (!a.has_a) acc[0] = acc[1];
    }
    if (a.do_inv) cta_fft<T, N, E, +1, W, 1>(v, j, c, sms, a.tw + (E == 16 ? N : 0));

    if (a.sc.mode) {          // only without a group split: the line index is the natural row
#pragma unroll
        for (int m = 0; m < E; m++) SGPE_ST_STREAM(scatter_ptr(a.sc, comp, j + m * NT, col), v[0][m]);
    } else {
#pragma unroll
        for (int m = 0; m < E; m++) SGPE_ST_STREAM(&a.out[off + (long long)(j + m * NT) * a.inner], v[0][m]);
    }

    if (any_k) {
        const int nblk = gridDim.x * gridDim.y * gridDim.z;
        const int blk = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;     // component-major
        cta_reduce<2>(acc, red);
        if (tid == 0) {
            double* p = a.partials + 2LL * blk;
            p[0] = acc[0]; p[1] = acc[1];
            __threadfence();
            red[0] = (atomicAdd(&a.counter[0], 1u) == (unsigned)(nblk - 1)) ? 1.0 : 0.0;
        }
        __syncthreads();
        const bool last = red[0] != 0.0;
        __syncthreads();
        if (last) {       // fixed-order fold: bit-reproducible
            __threadfence();
            double t4[4] = {0.0, 0.0, 0.0, 0.0};     // S0, T0, S1, T1
            const int half = nblk / 2;
            for (int t = tid; t < nblk; t += blockDim.x) {
                const int cp = (t >= half) ? 2 : 0;
                t4[cp + 0] += __ldcg(&a.partials[2LL * t]);
                t4[cp + 1] += __ldcg(&a.partials[2LL * t + 1]);
            }
            cta_reduce<4>(t4, red);
            if (tid == 0) {
                a.sums[0] = t4[1] + t4[3]; a.sums[1] = t4[0]; a.sums[2] = t4[2];
                a.counter[0] = 0u;
            }
        }
    }
}

// pack for the all-to-all: in [2][A][P*Bw] -> out [P][2][A][Bw]   (chunk q of every line goes to rank q)
template <typename T> struct PackArgs {
    typedef typename cx_of<T>::type C;
    const C* in; C* out; int A, P, Bw;
};
template <typename T>
__global__ void __launch_bounds__(256) slab_pack(PackArgs<T> a) {
    typedef typename cx_of<T>::type C;
    const long long per_line = (long long)a.P * a.Bw;
    const long long total = 2LL * a.A * per_line;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long w = i % a.Bw, q = (i / a.Bw) % a.P, l = (i / per_line) % a.A, c = i / (per_line * a.A);
        a.out[((q * 2 + c) * a.A + l) * a.Bw + w] = a.in[i];
    }
}
// unpack + transpose after the all-to-all: in [P][2][Bh][Bw] -> out [2][Bw][P*Bh], out[c][w][p*Bh+h] = in[p][c][h][w]
template <typename T> struct UnpackArgs {
    typedef typename cx_of<T>::type C;
    const C* in; C* out; int P, Bh, Bw;
};
template <typename T>
__global__ void __launch_bounds__(256) slab_unpack_transpose(UnpackArgs<T> a) {
    typedef typename cx_of<T>::type C;
    SGPE_DYN_SMEM(smem_raw);
    C* tile = reinterpret_cast<C*>(smem_raw);            // 32 x 33
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int tiles_w = a.Bw / 32, tiles_h = a.Bh / 32;
    const long long ntiles = (long long)a.P * 2 * tiles_h * tiles_w;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tw_ = (int)(t % tiles_w), th = (int)((t / tiles_w) % tiles_h);
        const int c = (int)((t / ((long long)tiles_w * tiles_h)) % 2), p = (int)(t / ((long long)tiles_w * tiles_h * 2));
        const C* src = a.in + (((long long)p * 2 + c) * a.Bh + th * 32) * a.Bw + tw_ * 32;
#pragma unroll
        for (int k = 0; k < 4; k++) tile[(ty + 8 * k) * 33 + tx] = src[(long long)(ty + 8 * k) * a.Bw + tx];
        __syncthreads();
        C* dst = a.out + ((long long)c * a.Bw + tw_ * 32) * ((long long)a.P * a.Bh) + (long long)p * a.Bh + th * 32;
#pragma unroll
        for (int k = 0; k < 4; k++) dst[(long long)(ty + 8 * k) * ((long long)a.P * a.Bh) + tx] = tile[tx * 33 + ty + 8 * k];
        __syncthreads();
    }
}

// out[i] = value, i < n (the slot counters of a graph replay)
template <typename I>
__global__ void __launch_bounds__(128) fill_value(I* out, I value, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = value;
}

// factor table: out[i] = exp(-i * e[i] * tau)   (separable operators: 1-D energy vectors -> 1-D factor tables)
template <typename T> struct ExpTableArgs {
    typedef typename cx_of<T>::type C;
    const double* e; C* out; long long n; double tr, ti; int tm;
};
template <typename T>
__global__ void __launch_bounds__(256) exp_table(ExpTableArgs<T> a) {
    typedef typename cx_of<T>::type C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x)
    {   // one-off tables: always evaluated in double, rounded once to the plan's precision
        const double2 w = (a.tm == TM_REAL) ? evo<TM_REAL, double, double2>(a.e[i], a.tr, a.ti)
                                            : evo<TM_IMAG, double, double2>(a.e[i], a.tr, a.ti);
        C o; o.x = (T)w.x; o.y = (T)w.y;
        a.out[i] = o;
    }
}

// out = in * sqrt(N / (dv * (S0 + S1)))  — the trailing ttools.norm of single_step
// (tensor_propagator.py:271) applied when the k-space state is materialised.
template <typename T> struct ScaleArgs {
    typedef typename cx_of<T>::type C;
    const C* in; C* out; long long per_batch;   // elements per trajectory (2*ny*nx)
    const double* totals; double atom_over_dv;
};
template <typename T>
__global__ void __launch_bounds__(256) scale_by_norm(ScaleArgs<T> a) {
    typedef typename cx_of<T>::type C;
    const int b = blockIdx.y;
    const double* tot = a.totals + (long long)b * 4;
    const T beta = (T)sqrt(a.atom_over_dv / (tot[1] + tot[2]));
    const long long base = (long long)b * a.per_batch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.per_batch;
         i += (long long)gridDim.x * blockDim.x)
        a.out[base + i] = cscale(a.in[base + i], beta);
}

// per-component sums of |psi|^2 (ttools.calc_pops / norm_sq, tensor_tools.py:437, 482): partials then
// the last CTA folds them in a fixed order -> totals[b][1], totals[b][2] (and [0] = their sum).
template <typename T> struct SumsqArgs {
    typedef typename cx_of<T>::type C;
    const C* in; long long plane;
    double* partials; unsigned* counter; double* totals;
    double* out2;                  // optional [B][2] compact copy of the per-component sums
};
template <typename T>
__global__ void __launch_bounds__(256) sumsq_pass(SumsqArgs<T> a) {
    typedef typename cx_of<T>::type C;
    SGPE_DYN_SMEM(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw);
    const int b = blockIdx.y, nblk = gridDim.x, tid = threadIdx.x;
    double acc[2] = {0.0, 0.0};
    for (int comp = 0; comp < 2; comp++) {
        const C* p = a.in + ((long long)b * 2 + comp) * a.plane;
        for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < a.plane; i += (long long)nblk * blockDim.x) {
            const C z = p[i];
            acc[comp] += (double)z.x * z.x + (double)z.y * z.y;
        }
    }
    cta_reduce<2>(acc, red);
    if (tid == 0) {
        double* p = a.partials + ((long long)b * nblk + blockIdx.x) * 2;
        p[0] = acc[0]; p[1] = acc[1];
        __threadfence();
        const unsigned ticket = atomicAdd(&a.counter[b], 1u);
        red[0] = (ticket == (unsigned)(nblk - 1)) ? 1.0 : 0.0;
    }
    __syncthreads();
    const bool last = red[0] != 0.0;
    __syncthreads();
    if (last) {
        __threadfence();
        double t2[2] = {0.0, 0.0};
        const double* p = a.partials + (long long)b * nblk * 2;
        for (int t = tid; t < nblk; t += blockDim.x) { t2[0] += __ldcg(&p[2 * t]); t2[1] += __ldcg(&p[2 * t + 1]); }
        cta_reduce<2>(t2, red);
        if (tid == 0) {
            double* tot = a.totals + (long long)b * 4;
            tot[0] = t2[0] + t2[1]; tot[1] = t2[0]; tot[2] = t2[1];
            if (a.out2 != nullptr) { a.out2[2 * b] = t2[0]; a.out2[2 * b + 1] = t2[1]; }
            a.counter[b] = 0u;
        }
    }
}

// Spectral kinetic energy  sum_k kin_c(k) |psi_k,c|^2  per component (times `scale` = dv_k): the k-space counterpart of
// the finite-difference + unwrapped-phase expression of eng_expect (tensor_propagator.py:306-311) that needs no phase
// at all.  kin_c is the propagator's own kinetic grid (dense, or separable kin_x[c][kx] + kin_y[c][ky]),
// Raman shift and the reference's "- min" offset (pspinor.py:496-501) included.  Same two-stage fixed-order reduction
// as sumsq_pass.
template <typename T> struct KineticArgs {
    typedef typename cx_of<T>::type C;
    const C* in; int nx, ny; long long plane;
    int kin_mode; const double* kin0; const double* kin1; long long kin_bstride;
    const double* kin_x; const double* kin_y; long long kinx_bstride, kiny_bstride;
    double scale;
    double* partials; unsigned* counter; double* out;      // out [B][2]
};
template <typename T>
__global__ void __launch_bounds__(256) kinetic_pass(KineticArgs<T> a) {
    typedef typename cx_of<T>::type C;
    SGPE_DYN_SMEM(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw);
    const int b = blockIdx.y, nblk = gridDim.x, tid = threadIdx.x;
    double acc[2] = {0.0, 0.0};
    for (int comp = 0; comp < 2; comp++) {
        const C* p = a.in + ((long long)b * 2 + comp) * a.plane;
        const double* kd = (comp == 0 ? a.kin0 : a.kin1) + (long long)b * a.kin_bstride;
        const double* kx = a.kin_x + (long long)b * a.kinx_bstride + (long long)comp * a.nx;
        const double* ky = a.kin_y + (long long)b * a.kiny_bstride + (long long)comp * a.ny;
        for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < a.plane; i += (long long)nblk * blockDim.x) {
            const C z = p[i];
            double e;
            if (a.kin_mode == 0) e = __ldg(&kd[i]);
            else { const int iy = (int)(i / a.nx); e = __ldg(&kx[i - (long long)iy * a.nx]) + __ldg(&ky[iy]); }
            acc[comp] += e * ((double)z.x * z.x + (double)z.y * z.y);
        }
    }
    cta_reduce<2>(acc, red);
    if (tid == 0) {
        double* p = a.partials + ((long long)b * nblk + blockIdx.x) * 2;
        p[0] = acc[0]; p[1] = acc[1];
        __threadfence();
        const unsigned ticket = atomicAdd(&a.counter[b], 1u);
        red[0] = (ticket == (unsigned)(nblk - 1)) ? 1.0 : 0.0;
    }
    __syncthreads();
    const bool last = red[0] != 0.0;
    __syncthreads();
    if (last) {
        __threadfence();
        double t2[2] = {0.0, 0.0};
        const double* p = a.partials + (long long)b * nblk * 2;
        for (int t = tid; t < nblk; t += blockDim.x) { t2[0] += __ldcg(&p[2 * t]); t2[1] += __ldcg(&p[2 * t + 1]); }
        cta_reduce<2>(t2, red);
        if (tid == 0) {
            a.out[2 * b] = t2[0] * a.scale; a.out[2 * b + 1] = t2[1] * a.scale;
            a.counter[b] = 0u;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Energy expectation (reference TensorPropagator.eng_expect, tensor_propagator.py:298-324) on the
// real-space state.  Pass 1: per-component max density (the mask threshold of phase_comp,
// tensor_tools.py:538).  Pass 2: per pixel the np.gradient stencils (2nd-order central, 1st-order at
// the edges, tensor_tools.py:342) of sqrt(n) and of the masked phase, the potential / interaction /
// coupling terms, summed over the grid with no volume element (:322-324).  Reference quirks kept:
// the first np.gradient output (d/d axis 0, i.e. along y) uses spacing dr[0] and is what the Raman
// term multiplies; the interaction term has no 1/2; the coupling term ignores the Raman phase.
template <typename T> struct MaxDensArgs {
    typedef typename cx_of<T>::type C;
    const C* psi; long long plane;
    double* partials; unsigned* counter; double* maxdens;   // maxdens [B][2]
};
template <typename T>
__global__ void __launch_bounds__(256) maxdens_pass(MaxDensArgs<T> a) {
    typedef typename cx_of<T>::type C;
    SGPE_DYN_SMEM(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw);
    const int b = blockIdx.y, nblk = gridDim.x, tid = threadIdx.x;
    double mx[2] = {0.0, 0.0};
    for (int comp = 0; comp < 2; comp++) {
        const C* p = a.psi + ((long long)b * 2 + comp) * a.plane;
        for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < a.plane; i += (long long)nblk * blockDim.x) {
            const C z = p[i];
            const double d = (double)z.x * z.x + (double)z.y * z.y;
            mx[comp] = d > mx[comp] ? d : mx[comp];
        }
    }
    // max-reduce through shared memory (order independent)
    red[tid * 2] = mx[0]; red[tid * 2 + 1] = mx[1];
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (tid < s) {
            red[tid * 2] = red[tid * 2] > red[(tid + s) * 2] ? red[tid * 2] : red[(tid + s) * 2];
            red[tid * 2 + 1] = red[tid * 2 + 1] > red[(tid + s) * 2 + 1] ? red[tid * 2 + 1] : red[(tid + s) * 2 + 1];
        }
        __syncthreads();
    }
    if (tid == 0) {
        double* p = a.partials + ((long long)b * nblk + blockIdx.x) * 2;
        p[0] = red[0]; p[1] = red[1];
        __threadfence();
        red[2] = (atomicAdd(&a.counter[b], 1u) == (unsigned)(nblk - 1)) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (red[2] != 0.0 && tid == 0) {
        __threadfence();
        const double* p = a.partials + (long long)b * nblk * 2;
        double m0 = 0.0, m1 = 0.0;
        for (int t = 0; t < nblk; t++) {
            const double v0 = __ldcg(&p[2 * t]), v1 = __ldcg(&p[2 * t + 1]);
            m0 = v0 > m0 ? v0 : m0; m1 = v1 > m1 ? v1 : m1;
        }
        a.maxdens[2 * b] = m0; a.maxdens[2 * b + 1] = m1;
        a.counter[b] = 0u;
    }
}

template <typename T> struct EnergyArgs {
    typedef typename cx_of<T>::type C;
    const C* psi; int nx, ny; long long plane;
    const double* pot0; const double* pot1; long long pot_bstride;
    int pot_mode; const double* pot_x; const double* pot_y; long long potx_bstride, poty_bstride;   // separable
    int cpl_mode; const double* coupling; long long cpl_bstride; const double* omega_b;
    double g_uu, g_dd, g_ud;
    double kl2;                 // 2 * kL_recoil * is_coupling
    double inv_h0, inv_h1;      // 1/dr[0] (used along axis 0 = y!) and 1/dr[1] (axis 1 = x)
    int unwrap_mode;            // 0: identity (wrapped phase as is), 1: local wrapped differences,
                                // 2: unwrapped field = wrapped phase + 2 pi * inc (unwrap.cuh)
    const int* inc;             // mode 2: [B][2][ny][nx] multiples of 2 pi from the region merging
    const double* maxdens;
    double* partials; unsigned* counter; double* out;        // out [b * out_bstride + {0..3}]: total, kin, pot, int
    long long out_bstride;
    int rows;                   // streaming kernel: rows per CTA band
    int polar;                  // streaming kernel: `psi` holds (|psi|, arg psi) (RowArgs::polar)
    int* slot_ctr;              // [B] or null.  Non-null (graph replay): the result goes to out[b * out_bstride + 4 * slot + q]
                                // with the slot read from slot_ctr[b] and post-incremented by the fold
};

SGPE_DI double sgpe_wrap_pi(double d) {
    const double pi = 3.14159265358979323846;
    if (d > pi) d -= 2 * pi;
    if (d < -pi) d += 2 * pi;
    return d;
}
// np.gradient along one axis from the values at i-1, i, i+1 (clamped loads at the edges)
SGPE_DI double sgpe_grad3(double fm, double f0, double fp, int i, int n, double inv_h) {
    if (i == 0) return (fp - f0) * inv_h;
    if (i == n - 1) return (f0 - fm) * inv_h;
    return (fp - fm) * (0.5 * inv_h);
}
SGPE_DI double sgpe_grad3_wrapped(double fm, double f0, double fp, int i, int n, double inv_h) {
    if (i == 0) return sgpe_wrap_pi(fp - f0) * inv_h;
    if (i == n - 1) return sgpe_wrap_pi(f0 - fm) * inv_h;
    return (sgpe_wrap_pi(fp - f0) + sgpe_wrap_pi(f0 - fm)) * (0.5 * inv_h);
}

// ---------------------------------------------------------------------------------------------
// np.gradient of one (ny, nx) field (reference ttools.grad_comp, tensor_tools.py:331-350, which hands NumPy arrays to
// np.gradient and raises for tensors): second-order central differences inside, first-order one-sided at the edges.
// nc = 1 real field, nc = 2 complex field seen as interleaved (re, im) - np.gradient differentiates both parts alike.
// g0 = d / d(axis 0) with spacing 1 / inv_h0, g1 = d / d(axis 1) with spacing 1 / inv_h1.
template <typename R>
__global__ void __launch_bounds__(256) gradient_pass(const R* f, int ny, int nx, int nc, double inv_h0, double inv_h1, R* g0, R* g1) {
    const long long row = (long long)nx * nc, n = (long long)ny * row;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / row), j = (int)((idx - (long long)i * row) / nc);
        const double c = (double)f[idx];
        const double up = (double)f[i > 0 ? idx - row : idx], dn = (double)f[i < ny - 1 ? idx + row : idx];
        const double lf = (double)f[j > 0 ? idx - nc : idx], rt = (double)f[j < nx - 1 ? idx + nc : idx];
        g0[idx] = (R)sgpe_grad3(up, c, dn, i, ny, inv_h0);
        g1[idx] = (R)sgpe_grad3(lf, c, rt, j, nx, inv_h1);
    }
}

// Each CTA walks 32 x 8 pixel tiles.  sqrt(n) and the masked phase are evaluated ONCE per pixel of the tile plus its
// one-pixel halo (340 points for 256 outputs) into shared memory and the stencils read them from there: the
// square roots and arctangents were 5x redundant when every pixel evaluated its own neighbours (FP64-bound).
template <typename T>
__global__ void __launch_bounds__(256) energy_pass(EnergyArgs<T> a) {
    typedef typename cx_of<T>::type C;
    constexpr int TX = 32, TY = 8, HX = TX + 2, HY = TY + 2;
    SGPE_DYN_SMEM(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw);                 // [32 * 4]
    double* s_r = red + 32 * 4;                                        // [HY][HX] sqrt(n)
    double* s_ph = s_r + HX * HY;                                      // [HY][HX] masked phase
    const int b = blockIdx.y, nblk = gridDim.x, tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
    const int tiles_x = (a.nx + TX - 1) / TX, tiles_y = (a.ny + TY - 1) / TY;     // ragged last tiles on generic meshes
    const long long ntiles = (long long)tiles_x * tiles_y;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long t = blockIdx.x; t < ntiles; t += nblk) {
        const int i0 = (int)(t / tiles_x) * TY, j0 = (int)(t % tiles_x) * TX;
        const bool inside = (i0 + ty < a.ny) && (j0 + tx < a.nx);
        const int i = inside ? i0 + ty : 0;          // axis 0 (y)
        const int j = inside ? j0 + tx : 0;          // axis 1 (x)
        double kin = 0.0, dens[2];
        C ctr[2];
        for (int comp = 0; comp < 2; comp++) {
            const C* p = a.psi + ((long long)b * 2 + comp) * a.plane;
            const int* inc = a.unwrap_mode == 2 ? a.inc + ((long long)b * 2 + comp) * a.plane : nullptr;
            const double thr = a.maxdens[2 * b + comp] * 1e-6;
            for (int q = tid; q < HX * HY; q += 256) {
                const int ly = q / HX, lx = q - ly * HX;
                int gi = i0 + ly - 1, gj = j0 + lx - 1;                 // clamped: the edge stencils never use them
                gi = gi < 0 ? 0 : (gi > a.ny - 1 ? a.ny - 1 : gi);
                gj = gj < 0 ? 0 : (gj > a.nx - 1 ? a.nx - 1 : gj);
                const C z = p[(long long)gi * a.nx + gj];
                const double n = (double)z.x * z.x + (double)z.y * z.y;
                s_r[q] = sqrt(n);
                double ph = 0.0;
                if (!(n < thr)) {
                    ph = atan2((double)z.y, (double)z.x);
                    if (inc != nullptr) ph = __dadd_rn(ph, __dmul_rn(6.283185307179586, (double)inc[(long long)gi * a.nx + gj]));
                }
                s_ph[q] = ph;
            }
            __syncthreads();
            const int c0 = (ty + 1) * HX + (tx + 1);
            const C z0 = p[(long long)i * a.nx + j];
            const double n0 = (double)z0.x * z0.x + (double)z0.y * z0.y;
            const double r0 = sgpe_grad3(s_r[c0 - HX], s_r[c0], s_r[c0 + HX], i, a.ny, a.inv_h0);
            const double r1 = sgpe_grad3(s_r[c0 - 1], s_r[c0], s_r[c0 + 1], j, a.nx, a.inv_h1);
            double g0, g1;
            if (a.unwrap_mode != 1) {
                g0 = sgpe_grad3(s_ph[c0 - HX], s_ph[c0], s_ph[c0 + HX], i, a.ny, a.inv_h0);
                g1 = sgpe_grad3(s_ph[c0 - 1], s_ph[c0], s_ph[c0 + 1], j, a.nx, a.inv_h1);
            } else {
                g0 = sgpe_grad3_wrapped(s_ph[c0 - HX], s_ph[c0], s_ph[c0 + HX], i, a.ny, a.inv_h0);
                g1 = sgpe_grad3_wrapped(s_ph[c0 - 1], s_ph[c0], s_ph[c0 + 1], j, a.nx, a.inv_h1);
            }
            kin += (r0 * r0 + r1 * r1) + n0 * (g0 * g0 + g1 * g1) + n0 * g0 * a.kl2;
            dens[comp] = n0;
            ctr[comp] = z0;
            __syncthreads();
        }
        kin *= 0.5;
        const long long pix = (long long)i * a.nx + j;
        double v0, v1;
        if (a.pot_mode == 0) {
            v0 = __ldg(&a.pot0[(long long)b * a.pot_bstride + pix]);
            v1 = __ldg(&a.pot1[(long long)b * a.pot_bstride + pix]);
        } else {
            const double* px = a.pot_x + (long long)b * a.potx_bstride;
            const double* py = a.pot_y + (long long)b * a.poty_bstride;
            v0 = __ldg(&px[j]) + __ldg(&py[i]);
            v1 = __ldg(&px[a.nx + j]) + __ldg(&py[a.ny + i]);
        }
        const double pot = dens[0] * v0 + dens[1] * v1;
        const double inter = a.g_uu * dens[0] * dens[0] + a.g_dd * dens[1] * dens[1] + a.g_ud * dens[0] * dens[1];
        double om = 0.0;
        if (a.cpl_mode == 1) om = a.omega_b[b];
        else if (a.cpl_mode == 2) om = __ldg(&a.coupling[(long long)b * a.cpl_bstride + pix]);
        const double coupl = ((double)ctr[0].x * ctr[1].x + (double)ctr[0].y * ctr[1].y) * om;   // Re(conj(p0) p1) * Omega
        if (inside) { acc[0] += kin + pot + inter + coupl; acc[1] += kin; acc[2] += pot; acc[3] += inter; }
    }
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
    if (last) {
        __threadfence();
        double t4[4] = {0.0, 0.0, 0.0, 0.0};
        const double* p = a.partials + (long long)b * nblk * 4;
        for (int t = tid; t < nblk; t += blockDim.x) {
#pragma unroll
            for (int q = 0; q < 4; q++) t4[q] += __ldcg(&p[4 * t + q]);
        }
        cta_reduce<4>(t4, red);
        if (tid == 0) {
            long long slot = 0;
            if (a.slot_ctr != nullptr) { slot = a.slot_ctr[b]; a.slot_ctr[b] = (int)slot + 1; }
#pragma unroll
            for (int q = 0; q < 4; q++) a.out[(long long)b * a.out_bstride + 4 * slot + q] = t4[q];
            a.counter[b] = 0u;
        }
    }
}

// The same functional, streaming: a CTA owns a band of `rows` rows x 256 columns and walks down the rows with a
// three-row window in registers; sqrt(n) and the masked phase of a pixel are evaluated ONCE (the tiled kernel above pays a
// 33 % halo and two barriers per tile and component, and was latency-bound at 160 us for 2048^2), the x-neighbours come
// from a double-buffered shared-memory row (one barrier per row), the y-neighbours are the thread's own previous rows.
// Loads are full coalesced row segments, the next row is fetched while the current one is worked on.
// POLAR: `psi` holds (sqrt(n), atan2(im, re)) per pixel - what row_pass<..., FAST = 3> stores - and the pass is left with the
// mask, the stencils and the sums.
template <typename T, bool POLAR>
__global__ void __launch_bounds__(256, 2) energy_stream_pass(EnergyArgs<T> a) {
    typedef typename cx_of<T>::type C;
    constexpr int BX = 256, RW = BX + 2;
    SGPE_DYN_SMEM(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw);                 // [32 * 4]
    double* s_row = red + 32 * 4;                                      // [2 buffers][2 components][r, ph][RW]
    const int b = blockIdx.y, nblk = gridDim.x, tid = threadIdx.x;
    const int nxb = (a.nx + BX - 1) / BX;
    const int x0 = (int)(blockIdx.x % nxb) * BX, y0 = (int)(blockIdx.x / nxb) * a.rows;
    const int y1 = (y0 + a.rows < a.ny) ? y0 + a.rows : a.ny;
    const int x = x0 + tid;
    const bool col_ok = x < a.nx;
    const int xc = col_ok ? x : a.nx - 1;
    // the two halo columns of the band are looked after by threads 0 (left) and 1 (right)
    const int xh = tid == 0 ? (x0 > 0 ? x0 - 1 : 0) : ((x0 + BX < a.nx) ? x0 + BX : a.nx - 1);
    const C* p0 = a.psi + ((long long)b * 2) * a.plane;
    const C* p1 = p0 + a.plane;
    const int* inc0 = a.unwrap_mode == 2 ? a.inc + ((long long)b * 2) * a.plane : nullptr;
    const int* inc1 = a.unwrap_mode == 2 ? inc0 + a.plane : nullptr;
    const double thr0 = a.maxdens[2 * b] * 1e-6, thr1 = a.maxdens[2 * b + 1] * 1e-6;
    double om_u = 0.0;
    if (a.cpl_mode == 1) om_u = a.omega_b[b];

    auto prep = [&](C z, double thr, const int* inc, long long pix, double& r, double& ph) {
        double n;
        if (POLAR) { r = (double)z.x; n = r * r; }
        else { n = (double)z.x * z.x + (double)z.y * z.y; r = sqrt(n); }
        ph = 0.0;
        if (!(n < thr)) {
            ph = POLAR ? (double)z.y : atan2((double)z.y, (double)z.x);
            if (inc != nullptr) ph = __dadd_rn(ph, __dmul_rn(6.283185307179586, (double)inc[pix]));
        }
    };

    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    // window: index 0 = two rows back, 1 = previous row; per component r, ph; of the previous row also the x-derivatives,
    // the density and the wavefunction itself
    double wr[2][2], wph[2][2], gxr[2], gxph[2];
    C zprev[2];
    wr[0][0] = wr[0][1] = wr[1][0] = wr[1][1] = 0.0; wph[0][0] = wph[0][1] = wph[1][0] = wph[1][1] = 0.0;
    gxr[0] = gxr[1] = gxph[0] = gxph[1] = 0.0;
    zprev[0].x = zprev[0].y = zprev[1].x = zprev[1].y = (T)0;

    int yl = y0 - 1 < 0 ? 0 : y0 - 1;
    C zn0 = p0[(long long)yl * a.nx + xc], zn1 = p1[(long long)yl * a.nx + xc];
    C hn0 = zn0, hn1 = zn1;                              // threads 0 / 1: the halo pixels of the next row
    if (tid < 2) { hn0 = p0[(long long)yl * a.nx + xh]; hn1 = p1[(long long)yl * a.nx + xh]; }
    int it = 0;
    for (int yy = y0 - 1; yy <= y1; yy++, it++) {
        const C z0 = zn0, z1 = zn1, h0 = hn0, h1 = hn1;
        const int ycur = yy < 0 ? 0 : (yy > a.ny - 1 ? a.ny - 1 : yy);
        if (yy < y1) {                                   // fetch the next row while this one is worked on
            const int yn = yy + 1 > a.ny - 1 ? a.ny - 1 : yy + 1;
            zn0 = p0[(long long)yn * a.nx + xc]; zn1 = p1[(long long)yn * a.nx + xc];
            if (tid < 2) { hn0 = p0[(long long)yn * a.nx + xh]; hn1 = p1[(long long)yn * a.nx + xh]; }
            // (one row ahead in registers does not cover the DRAM latency: rows further down are pulled into L2)
            const int yf = yy + 6;
            if (yf <= y1 && yf < a.ny && (tid & 1) == 0) {
                SGPE_PREFETCH_L2(&p0[(long long)yf * a.nx + xc]);
                SGPE_PREFETCH_L2(&p1[(long long)yf * a.nx + xc]);
            }
        }
        const long long pix = (long long)ycur * a.nx + xc;
        double r[2], ph[2];
        prep(z0, thr0, inc0, pix, r[0], ph[0]);
        prep(z1, thr1, inc1, pix, r[1], ph[1]);
        double* buf = s_row + (it & 1) * (4 * RW);
        buf[0 * RW + tid + 1] = r[0]; buf[1 * RW + tid + 1] = ph[0];
        buf[2 * RW + tid + 1] = r[1]; buf[3 * RW + tid + 1] = ph[1];
        if (tid < 2) {
            const long long hp = (long long)ycur * a.nx + xh;
            double hr, hph;
            const int slot = tid == 0 ? 0 : RW - 1;
            prep(h0, thr0, inc0, hp, hr, hph); buf[0 * RW + slot] = hr; buf[1 * RW + slot] = hph;
            prep(h1, thr1, inc1, hp, hr, hph); buf[2 * RW + slot] = hr; buf[3 * RW + slot] = hph;
        }
        __syncthreads();
        // finish the previous row: its y-derivatives need this row
        if (yy - 1 >= y0 && yy - 1 < y1 && col_ok) {
            const int i = yy - 1;
            double kin = 0.0, dens[2];
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const double gyr = sgpe_grad3(wr[0][c], wr[1][c], r[c], i, a.ny, a.inv_h0);
                const double gyph = a.unwrap_mode != 1 ? sgpe_grad3(wph[0][c], wph[1][c], ph[c], i, a.ny, a.inv_h0)
                                                       : sgpe_grad3_wrapped(wph[0][c], wph[1][c], ph[c], i, a.ny, a.inv_h0);
                const double n0 = POLAR ? (double)zprev[c].x * zprev[c].x
                                        : (double)zprev[c].x * zprev[c].x + (double)zprev[c].y * zprev[c].y;
                kin += (gyr * gyr + gxr[c] * gxr[c]) + n0 * (gyph * gyph + gxph[c] * gxph[c]) + n0 * gyph * a.kl2;
                dens[c] = n0;
            }
            kin *= 0.5;
            const long long pp = (long long)i * a.nx + x;
            double v0, v1;
            if (a.pot_mode == 0) {
                v0 = __ldg(&a.pot0[(long long)b * a.pot_bstride + pp]);
                v1 = __ldg(&a.pot1[(long long)b * a.pot_bstride + pp]);
            } else {
                const double* px = a.pot_x + (long long)b * a.potx_bstride;
                const double* py = a.pot_y + (long long)b * a.poty_bstride;
                v0 = __ldg(&px[x]) + __ldg(&py[i]);
                v1 = __ldg(&px[a.nx + x]) + __ldg(&py[a.ny + i]);
            }
            const double pot = dens[0] * v0 + dens[1] * v1;
            const double inter = a.g_uu * dens[0] * dens[0] + a.g_dd * dens[1] * dens[1] + a.g_ud * dens[0] * dens[1];
            double om = om_u;
            if (a.cpl_mode == 2) om = __ldg(&a.coupling[(long long)b * a.cpl_bstride + pp]);
            double coupl = 0.0;                          // Re(conj(p0) p1) * Omega
            if (!POLAR) coupl = ((double)zprev[0].x * zprev[1].x + (double)zprev[0].y * zprev[1].y) * om;
            else if (om != 0.0) coupl = (double)zprev[0].x * zprev[1].x * cos((double)zprev[0].y - (double)zprev[1].y) * om;
            acc[0] += kin + pot + inter + coupl; acc[1] += kin; acc[2] += pot; acc[3] += inter;
        }
        // x-derivatives of this row (from the shared row), then shift the window
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const double rm = buf[(2 * c) * RW + tid], rp = buf[(2 * c) * RW + tid + 2];
            const double pm = buf[(2 * c + 1) * RW + tid], pq = buf[(2 * c + 1) * RW + tid + 2];
            gxr[c] = sgpe_grad3(rm, r[c], rp, xc, a.nx, a.inv_h1);
            gxph[c] = a.unwrap_mode != 1 ? sgpe_grad3(pm, ph[c], pq, xc, a.nx, a.inv_h1)
                                         : sgpe_grad3_wrapped(pm, ph[c], pq, xc, a.nx, a.inv_h1);
            wr[0][c] = wr[1][c]; wph[0][c] = wph[1][c];
            wr[1][c] = r[c]; wph[1][c] = ph[c];
        }
        zprev[0] = z0; zprev[1] = z1;
    }
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
    if (last) {
        __threadfence();
        double t4[4] = {0.0, 0.0, 0.0, 0.0};
        const double* p = a.partials + (long long)b * nblk * 4;
        for (int t = tid; t < nblk; t += blockDim.x) {
#pragma unroll
            for (int q = 0; q < 4; q++) t4[q] += __ldcg(&p[4 * t + q]);
        }
        cta_reduce<4>(t4, red);
        if (tid == 0) {
            long long slot = 0;
            if (a.slot_ctr != nullptr) { slot = a.slot_ctr[b]; a.slot_ctr[b] = (int)slot + 1; }
#pragma unroll
            for (int q = 0; q < 4; q++) a.out[(long long)b * a.out_bstride + 4 * slot + q] = t4[q];
            a.counter[b] = 0u;
        }
    }
}

// The stencil pass for POLAR input, without shared memory or CTA barriers: a WARP owns a band of `rows` rows x 30 columns
// (lanes 0 and 31 carry the halo columns), the x-neighbours come from warp shuffles, the y-neighbours are the lane's own
// previous rows, FOUR rows are in flight per lane (a ring of named registers: rotating them by copies would wait for
// every load) and every warp of the SM runs on its own.  Loads beyond the mesh are clamped to the edge pixel, which
// turns np.gradient's one-sided edge differences into the SAME expression (f[+1] - f[-1]) * c with c = 1/h instead of
// 1/(2h): no edge branches or selects.  WRAPPED: differences of the phase taken modulo 2 pi (unwrap_mode 1).
// (unwrap_mode 2 - a field of 2 pi multiples from the region merging - goes through energy_stream_pass.)
template <typename T, bool WRAPPED>
__global__ void __launch_bounds__(256, 2) energy_polar_pass(EnergyArgs<T> a) {
    typedef typename cx_of<T>::type C;
    SGPE_DYN_SMEM(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw);                 // [32 * 4]
    const int b = blockIdx.y, nblk = gridDim.x, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int ncb = (a.nx + 29) / 30, ncb8 = (ncb + 7) / 8;
    const int cb = (int)(blockIdx.x % ncb8) * 8 + warp;
    // (the bands are handed out from the bottom of the mesh upwards: the rows the preceding pass wrote last are still in L2)
    const int y0 = (int)((gridDim.x - 1 - blockIdx.x) / ncb8) * a.rows;
    const int y1 = (y0 + a.rows < a.ny) ? y0 + a.rows : a.ny;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    {   // (no branch around the body: a warp past the last column band works on the last one and adds nothing, so that
        // the compiler sees the shuffles in convergent code; y0 < ny by the construction of the grid)
        const int xr = (cb < ncb ? cb : ncb - 1) * 30 - 1 + lane;
        const int x = xr < 0 ? 0 : (xr > a.nx - 1 ? a.nx - 1 : xr);
        const bool out_ok = lane >= 1 && lane <= 30 && xr < a.nx && cb < ncb;
        const double cx = (x == 0 || x == a.nx - 1) ? a.inv_h1 : 0.5 * a.inv_h1;
        const C* p0 = a.psi + ((long long)b * 2) * a.plane + x;
        const double thr0 = a.maxdens[2 * b] * 1e-6, thr1 = a.maxdens[2 * b + 1] * 1e-6;
        double om_u = 0.0;
        if (a.cpl_mode == 1) om_u = a.omega_b[b];
        const bool dense_pot = a.pot_mode == 0, dense_cpl = a.cpl_mode == 2;
        double vx0 = 0.0, vx1 = 0.0;
        // (every batch / column offset is folded into these pointers once: nothing but the row index is left for the loop)
        const double* const py0 = a.pot_y + (long long)b * a.poty_bstride;
        const double* const py1 = py0 + a.ny;
        const double* const pd0 = dense_pot ? a.pot0 + (long long)b * a.pot_bstride + x : nullptr;
        const double* const pd1 = dense_pot ? a.pot1 + (long long)b * a.pot_bstride + x : nullptr;
        const double* const cpd = dense_cpl ? a.coupling + (long long)b * a.cpl_bstride + x : nullptr;
        const int nx = a.nx, ny = a.ny;
        const double inv_h0 = a.inv_h0, half_h0 = 0.5 * a.inv_h0, kl2 = a.kl2, g_uu = a.g_uu, g_dd = a.g_dd, g_ud = a.g_ud;
        if (!dense_pot) {
            const double* px = a.pot_x + (long long)b * a.potx_bstride;
            vx0 = __ldg(&px[x]); vx1 = __ldg(&px[a.nx + x]);
        }
        // window: w?0 = two rows back, w?1 = previous row (per component |psi| and masked phase); of the previous row also
        // the x-derivatives and the raw phases (coupling term)
        double wr0[2] = {0.0, 0.0}, wr1[2] = {0.0, 0.0}, wp0[2] = {0.0, 0.0}, wp1[2] = {0.0, 0.0};
        double gxr[2] = {0.0, 0.0}, gxph[2] = {0.0, 0.0}, rawp[2] = {0.0, 0.0};
        struct Row { C z0, z1; };
        // running fetch pointers: row fy clamped to the mesh (rows past the band are fetched and never used: no branch)
        int fy = y0 - 1;
        const C* f0 = p0 + (long long)(fy < 0 ? 0 : fy) * a.nx;
        const long long plane = a.plane;
        auto fetch = [&](Row& q) {
            q.z0 = SGPE_LD_STREAM(&f0[0]); q.z1 = SGPE_LD_STREAM(&f0[plane]);
            if (fy >= 0 && fy < ny - 1) f0 += nx;
            fy++;
        };
        auto process = [&](int yy, const Row& q) {
            double r[2], ph[2];
            r[0] = (double)q.z0.x; r[1] = (double)q.z1.x;
            ph[0] = (r[0] * r[0] < thr0) ? 0.0 : (double)q.z0.y;
            ph[1] = (r[1] * r[1] < thr1) ? 0.0 : (double)q.z1.y;
            const int i = yy - 1;                     // the previous row is finished now: its y-derivatives need this row
            if (i >= y0 && i < y1 && out_ok) {
                const double cy = (i == 0 || i == ny - 1) ? inv_h0 : half_h0;
                double kin = 0.0, dens[2];
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const double gyr = (r[c] - wr0[c]) * cy;
                    const double gyph = WRAPPED ? (sgpe_wrap_pi(ph[c] - wp1[c]) + sgpe_wrap_pi(wp1[c] - wp0[c])) * cy
                                                : (ph[c] - wp0[c]) * cy;
                    const double n0 = wr1[c] * wr1[c];
                    kin += (gyr * gyr + gxr[c] * gxr[c]) + n0 * (gyph * gyph + gxph[c] * gxph[c]) + n0 * gyph * kl2;
                    dens[c] = n0;
                }
                kin *= 0.5;
                double v0, v1;
                if (dense_pot) {
                    const long long pp = (long long)i * nx;
                    v0 = __ldg(&pd0[pp]); v1 = __ldg(&pd1[pp]);
                } else {
                    v0 = vx0 + __ldg(&py0[i]); v1 = vx1 + __ldg(&py1[i]);
                }
                const double pot = dens[0] * v0 + dens[1] * v1;
                const double inter = g_uu * dens[0] * dens[0] + g_dd * dens[1] * dens[1] + g_ud * dens[0] * dens[1];
                double om = om_u;
                if (dense_cpl) om = __ldg(&cpd[(long long)i * nx]);
                double coupl = 0.0;                      // Re(conj(p0) p1) * Omega
                if (om != 0.0) coupl = wr1[0] * wr1[1] * cos(rawp[0] - rawp[1]) * om;
                acc[0] += kin + pot + inter + coupl; acc[1] += kin; acc[2] += pot; acc[3] += inter;
            }
            // x-derivatives of this row from the neighbouring lanes, then shift the window
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const double rm = __shfl_up_sync(0xffffffffu, r[c], 1), rp = __shfl_down_sync(0xffffffffu, r[c], 1);
                const double pm = __shfl_up_sync(0xffffffffu, ph[c], 1), pq = __shfl_down_sync(0xffffffffu, ph[c], 1);
                gxr[c] = (rp - rm) * cx;
                gxph[c] = WRAPPED ? (sgpe_wrap_pi(pq - ph[c]) + sgpe_wrap_pi(ph[c] - pm)) * cx : (pq - pm) * cx;
                wr0[c] = wr1[c]; wp0[c] = wp1[c];
                wr1[c] = r[c]; wp1[c] = ph[c];
            }
            rawp[0] = (double)q.z0.y; rawp[1] = (double)q.z1.y;
        };
        Row qa, qb, qc, qd;
        fetch(qa); fetch(qb); fetch(qc); fetch(qd);
#pragma unroll 1
        for (int yy = y0 - 1; yy <= y1; yy += 4) {     // (rows past y1 are processed into nothing: i >= y1)
            process(yy, qa);     fetch(qa);
            process(yy + 1, qb); fetch(qb);
            process(yy + 2, qc); fetch(qc);
            process(yy + 3, qd); fetch(qd);
        }
    }
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
    if (last) {
        __threadfence();
        double t4[4] = {0.0, 0.0, 0.0, 0.0};
        const double* p = a.partials + (long long)b * nblk * 4;
        for (int t = tid; t < nblk; t += blockDim.x) {
#pragma unroll
            for (int q = 0; q < 4; q++) t4[q] += __ldcg(&p[4 * t + q]);
        }
        cta_reduce<4>(t4, red);
        if (tid == 0) {
            long long slot = 0;
            if (a.slot_ctr != nullptr) { slot = a.slot_ctr[b]; a.slot_ctr[b] = (int)slot + 1; }
#pragma unroll
            for (int q = 0; q < 4; q++) a.out[(long long)b * a.out_bstride + 4 * slot + q] = t4[q];
            a.counter[b] = 0u;
        }
    }
}

}  // namespace sgpe
