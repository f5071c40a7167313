// fft_core.cuh — register-resident Stockham FFT building blocks for one CTA.
//
// Data model: a "line" is one 1-D transform of length N.  A line is owned by NT = N/E threads; thread j
// of the line holds the E elements  x[j + m*NT], m = 0..E-1  in registers — the same element set before
// and after every stage, and before/after the whole transform (natural order in, natural order out).
// Between stages the CTA exchanges data through shared memory (one write + one read of the line).
// Shared memory is XOR-swizzled so that both the strided stage writes and the unit-stride reads are
// bank-conflict free for 16-byte (complex128) and 8-byte (complex64) elements; W lines can be
// interleaved element-wise ([n][c] layout, c fastest) which is what the column pass uses to turn W
// adjacent columns into 64/128-byte global-memory segments.
//
// Stage s (radix R = min(E, N/Ns), Ns = product of the previous radices) is the textbook Stockham
// autosort step: butterfly jb in [0, N/R) reads x[jb + t*N/R], multiplies by w_{Ns*R}^{t*(jb mod Ns)},
// does a DFT_R and writes y[(jb/Ns)*Ns*R + (jb mod Ns) + t*Ns].  A thread runs E/R butterflies
// (jb = j + b*NT), whose inputs are exactly its E registers.
//
// This replaces the library calls torch.fft.fftn / ifftn at reference tensor_tools.py:225, 256.
#pragma once

#include "emu_or_cuda.h"

namespace sgpe {

template <typename T> struct cx_of;
template <> struct cx_of<double> { typedef double2 type; };
template <> struct cx_of<float>  { typedef float2 type; };

#define SGPE_DI __device__ __forceinline__

template <typename C> SGPE_DI C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> SGPE_DI C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
// a * b
template <typename C> SGPE_DI C cmul(C a, C b) {
    C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
// a * conj(b)
template <typename C> SGPE_DI C cmulc(C a, C b) {
    C r; r.x = a.x * b.x + a.y * b.y; r.y = a.y * b.x - a.x * b.y; return r;
}
template <typename C, typename T> SGPE_DI C cscale(C a, T s) { a.x *= s; a.y *= s; return a; }
#if !defined(SGPE_EMU) && !defined(SGPE_NO_F32X2)
// complex64: Blackwell's packed FP32x2 pipe (FADD2 / FMUL2 / FFMA2) adds, subtracts and scales a complex number in ONE
// instruction (the operand negation of FADD2 covers both halves).  The complex64 passes are issue-bound (ncu: issue slots
// 50-60 % busy at 4 warps per scheduler, FP32 pipe 27-43 %), and two thirds of their FP32 instructions are the additions of
// the butterflies.  Same IEEE results as the scalar forms.
SGPE_DI float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
SGPE_DI float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
SGPE_DI float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
#endif
// multiply by DIR*i  (DIR = -1: forward transform, -i;  DIR = +1: inverse, +i)
template <int DIR, typename C> SGPE_DI C mul_i(C a) {
    C r;
    if (DIR < 0) { r.x = a.y; r.y = -a.x; } else { r.x = -a.y; r.y = a.x; }
    return r;
}

// ---- constant twiddles exp(DIR * 2*pi*i * M / 16)
template <typename T> struct W16 {
    // cos / sin of 2*pi*m/16, m = 0..4
    static constexpr T C1 = (T)0.92387953251128675613;   // cos(pi/8)
    static constexpr T S1 = (T)0.38268343236508977173;   // sin(pi/8)
    static constexpr T H  = (T)0.70710678118654752440;   // sqrt(1/2)
};

template <int DIR, int M, typename T, typename C> SGPE_DI C mul_w16(C a) {
    constexpr int m = ((M % 16) + 16) % 16;
    if constexpr (m == 0) { return a; }
    else if constexpr (m == 4) { return mul_i<DIR>(a); }
    else if constexpr (m == 8) { C r; r.x = -a.x; r.y = -a.y; return r; }
    else if constexpr (m == 12) { return mul_i<-DIR>(a); }
    else if constexpr (m == 2) {   // (H, DIR*H)
        C r; r.x = W16<T>::H * (a.x - (T)DIR * a.y); r.y = W16<T>::H * (a.y + (T)DIR * a.x); return r;
    }
    else if constexpr (m == 6) {   // (-H, DIR*H)
        C r; r.x = W16<T>::H * (-a.x - (T)DIR * a.y); r.y = W16<T>::H * ((T)DIR * a.x - a.y); return r;
    }
    else if constexpr (m == 10) {  // (-H, -DIR*H)
        C r; r.x = W16<T>::H * ((T)DIR * a.y - a.x); r.y = W16<T>::H * (-a.y - (T)DIR * a.x); return r;
    }
    else if constexpr (m == 14) {  // (H, -DIR*H)
        C r; r.x = W16<T>::H * (a.x + (T)DIR * a.y); r.y = W16<T>::H * (a.y - (T)DIR * a.x); return r;
    }
    else {
        // odd m: (wr, DIR*wi) with wr = cos(2 pi m/16), wi = sin(2 pi m/16)
        constexpr T wr = (m == 1 || m == 15) ? W16<T>::C1 : (m == 3 || m == 13) ? W16<T>::S1
                       : (m == 5 || m == 11) ? -W16<T>::S1 : -W16<T>::C1;
        constexpr T wi0 = (m == 1 || m == 7) ? W16<T>::S1 : (m == 3 || m == 5) ? W16<T>::C1
                        : (m == 9 || m == 15) ? -W16<T>::S1 : -W16<T>::C1;
        constexpr T wi = (T)DIR * wi0;
        C r; r.x = a.x * wr - a.y * wi; r.y = a.x * wi + a.y * wr; return r;
    }
}

// ---- small in-register DFTs, natural order in and out
template <int DIR, typename C> SGPE_DI void dft2(C& a0, C& a1) {
    C t = a0; a0 = cadd(t, a1); a1 = csub(t, a1);
}
template <int DIR, typename C> SGPE_DI void dft4(C& a0, C& a1, C& a2, C& a3) {
    C t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_i<DIR>(csub(a1, a3));
    a0 = cadd(t0, t2); a2 = csub(t0, t2); a1 = cadd(t1, t3); a3 = csub(t1, t3);
}
template <int DIR, typename T, typename C> SGPE_DI void dft8(C (&a)[8]) {
    C e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
    C o0 = a[1], o1 = a[3], o2 = a[5], o3 = a[7];
    dft4<DIR>(e0, e1, e2, e3);
    dft4<DIR>(o0, o1, o2, o3);
    o1 = mul_w16<DIR, 2, T>(o1);
    o2 = mul_w16<DIR, 4, T>(o2);
    o3 = mul_w16<DIR, 6, T>(o3);
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
    a[1] = cadd(e1, o1); a[5] = csub(e1, o1);
    a[2] = cadd(e2, o2); a[6] = csub(e2, o2);
    a[3] = cadd(e3, o3); a[7] = csub(e3, o3);
}
template <int DIR, typename T, typename C> SGPE_DI void dft16(C (&a)[16]) {
    // n = 4*n1 + n2 ; k = k1 + 4*k2
    dft4<DIR>(a[0], a[4], a[8],  a[12]);
    dft4<DIR>(a[1], a[5], a[9],  a[13]);
    dft4<DIR>(a[2], a[6], a[10], a[14]);
    dft4<DIR>(a[3], a[7], a[11], a[15]);
    // a[4*k1 + n2] *= w16^(n2*k1)
    a[5]  = mul_w16<DIR, 1, T>(a[5]);  a[6]  = mul_w16<DIR, 2, T>(a[6]);  a[7]  = mul_w16<DIR, 3, T>(a[7]);
    a[9]  = mul_w16<DIR, 2, T>(a[9]);  a[10] = mul_w16<DIR, 4, T>(a[10]); a[11] = mul_w16<DIR, 6, T>(a[11]);
    a[13] = mul_w16<DIR, 3, T>(a[13]); a[14] = mul_w16<DIR, 6, T>(a[14]); a[15] = mul_w16<DIR, 9, T>(a[15]);
    dft4<DIR>(a[0],  a[1],  a[2],  a[3]);
    dft4<DIR>(a[4],  a[5],  a[6],  a[7]);
    dft4<DIR>(a[8],  a[9],  a[10], a[11]);
    dft4<DIR>(a[12], a[13], a[14], a[15]);
    // a[4*k1 + k2] holds X[k1 + 4*k2]  -> transpose the 4x4
    C t;
    t = a[1];  a[1]  = a[4];  a[4]  = t;
    t = a[2];  a[2]  = a[8];  a[8]  = t;
    t = a[3];  a[3]  = a[12]; a[12] = t;
    t = a[6];  a[6]  = a[9];  a[9]  = t;
    t = a[7];  a[7]  = a[13]; a[13] = t;
    t = a[11]; a[11] = a[14]; a[14] = t;
}

template <int R, int DIR, typename T, int STRIDE, typename C> SGPE_DI void dft_strided(C* p) {
    if constexpr (R == 2) {
        dft2<DIR>(p[0], p[STRIDE]);
    } else if constexpr (R == 4) {
        dft4<DIR>(p[0], p[STRIDE], p[2 * STRIDE], p[3 * STRIDE]);
    } else if constexpr (R == 8) {
        C a[8];
#pragma unroll
        for (int t = 0; t < 8; t++) a[t] = p[t * STRIDE];
        dft8<DIR, T>(a);
#pragma unroll
        for (int t = 0; t < 8; t++) p[t * STRIDE] = a[t];
    } else {
        static_assert(R == 16, "radix must be 2, 4, 8 or 16");
        C a[16];
#pragma unroll
        for (int t = 0; t < 16; t++) a[t] = p[t * STRIDE];
        dft16<DIR, T>(a);
#pragma unroll
        for (int t = 0; t < 16; t++) p[t * STRIDE] = a[t];
    }
}

// ---- geometry of one length-N transform held E elements per thread
template <typename T, int N, int E, int W>
struct LineGeom {
    static_assert((N & (N - 1)) == 0 && (E & (E - 1)) == 0 && E <= N, "powers of two");
    typedef typename cx_of<T>::type C;
    static constexpr int NT = N / E;                           // threads per line
    static constexpr int R0 = E;                               // first-stage radix (E <= N)
    static constexpr int PHASE = 128 / (int)sizeof(C);         // lanes served per shared-memory wavefront
    static constexpr int MASK = (PHASE / W > 1) ? (PHASE / W - 1) : 0;
    static constexpr int log2c(int v) { return v <= 1 ? 0 : 1 + log2c(v >> 1); }
    static constexpr int SH = log2c(R0);
    SGPE_DI static int swz(int n) { return n ^ ((n >> SH) & MASK); }
    static constexpr int radix(int Ns) { return (N / Ns) < E ? (N / Ns) : E; }
};

// Where a stage reads its twiddles: global memory through the read-only path (plain pointer), or a copy of the tables
// in shared memory (SmemTable: persistent kernels whose shared-memory footprint leaves no L1 to cache them in).
template <typename C> struct SmemTable {
    const C* p;
    SGPE_DI SmemTable operator+(int o) const { SmemTable r; r.p = p + o; return r; }
};
template <typename C> SGPE_DI C tw_at(const C* __restrict__ p, int i) { return __ldg(&p[i]); }
template <typename C> SGPE_DI C tw_at(SmemTable<C> t, int i) { return t.p[i]; }
// twiddle + butterflies of the stage whose previous-radix product is Ns, on one line's registers
template <typename T, int N, int E, int DIR, int Ns, typename C, typename TWP>
SGPE_DI void stage_compute(C (&v)[E], int j, TWP tw) {
    constexpr int R = (N / Ns) < E ? (N / Ns) : E;
    constexpr int NB = E / R;
    constexpr int NT = N / E;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        if constexpr (Ns > 1) {
            // per-stage table, [t-1][k] with k fastest (consecutive lanes read consecutive entries); the
            // stage with previous-radix product Ns starts at entry Ns - E (see sgpe_api.cu: upload_twiddles)
            const int k = (j + b * NT) & (Ns - 1);
#pragma unroll
            for (int t = 1; t < R; t++) {
                const C w = tw_at<C>(tw, (Ns - E) + k + (t - 1) * Ns);
                v[b + t * NB] = (DIR < 0) ? cmul(v[b + t * NB], w) : cmulc(v[b + t * NB], w);
            }
        }
        dft_strided<R, DIR, T, NB>(&v[b]);
    }
}

// scatter the stage outputs into the line's shared-memory image (element n of column c at swz(n)*W + c)
template <typename T, int N, int E, int W, int Ns, typename C>
SGPE_DI void stage_store(const C (&v)[E], int j, int c, C* sm) {
    typedef LineGeom<T, N, E, W> G;
    constexpr int R = (N / Ns) < E ? (N / Ns) : E;
    constexpr int NB = E / R;
    constexpr int NT = N / E;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int jb = j + b * NT;
        const int k = jb & (Ns - 1);
        const int j0 = (jb - k) * R + k;
#pragma unroll
        for (int t = 0; t < R; t++) sm[G::swz(j0 + t * Ns) * W + c] = v[b + t * NB];
    }
}

template <typename T, int N, int E, int W, typename C>
SGPE_DI void stage_load(C (&v)[E], int j, int c, const C* sm) {
    typedef LineGeom<T, N, E, W> G;
    constexpr int NT = N / E;
#pragma unroll
    for (int m = 0; m < E; m++) v[m] = sm[G::swz(j + m * NT) * W + c];
}

// Who synchronises at the exchange points of a transform: the whole CTA, or one of several independent thread
// groups of a CTA (named barrier `id` over `count` threads, whole warps).  Groups of one CTA drift out of phase, so
// one group's butterflies overlap the other's shared-memory exchange — what two resident CTAs would give where the
// register file only has room for one.
struct CtaBar {
    SGPE_DI void sync() const { __syncthreads(); }
};
struct GroupBar {
    int id, count;
    SGPE_DI void sync() const { SGPE_NAMED_BAR(id, count); }
};

// Full transform of L lines per thread (each with its own shared-memory image); all threads of the
// CTA (or of the barrier group) must call this together (it contains barriers).
template <typename T, int N, int E, int DIR, int W, int L, int Ns, typename C, typename B, typename TWP>
SGPE_DI void cta_fft_from(C (&v)[L][E], int j, int c, C* const (&sm)[L], TWP tw, const B& bar) {
    constexpr int R = (N / Ns) < E ? (N / Ns) : E;
#pragma unroll
    for (int l = 0; l < L; l++) stage_compute<T, N, E, DIR, Ns>(v[l], j, tw);
    if constexpr (Ns * R < N) {
#pragma unroll
        for (int l = 0; l < L; l++) stage_store<T, N, E, W, Ns>(v[l], j, c, sm[l]);
        bar.sync();
#pragma unroll
        for (int l = 0; l < L; l++) stage_load<T, N, E, W>(v[l], j, c, sm[l]);
        bar.sync();
        cta_fft_from<T, N, E, DIR, W, L, Ns * R>(v, j, c, sm, tw, bar);
    }
}

// product of the radices before the LAST stage of the length-N, E-per-thread plan (1 when there is a single stage)
template <int N, int E, int Ns = 1> struct LastStage {
    static constexpr int R = (N / Ns) < E ? (N / Ns) : E;
    static constexpr int value = LastStage<N, E, (Ns * R >= N) ? 0 : Ns * R>::value == 0 ? Ns : LastStage<N, E, (Ns * R >= N) ? 0 : Ns * R>::value;
};
template <int N, int E> struct LastStage<N, E, 0> { static constexpr int value = 0; };
// every stage from Ns on EXCEPT the butterflies of the last one: returns behind the barrier that follows the last
// exchange read, i.e. at the point from which the shared-memory images are free again
template <typename T, int N, int E, int DIR, int W, int L, int Ns, typename C, typename B, typename TWP>
SGPE_DI void cta_fft_head(C (&v)[L][E], int j, int c, C* const (&sm)[L], TWP tw, const B& bar) {
    constexpr int R = (N / Ns) < E ? (N / Ns) : E;
    if constexpr (Ns * R < N) {
#pragma unroll
        for (int l = 0; l < L; l++) stage_compute<T, N, E, DIR, Ns>(v[l], j, tw);
#pragma unroll
        for (int l = 0; l < L; l++) stage_store<T, N, E, W, Ns>(v[l], j, c, sm[l]);
        bar.sync();
#pragma unroll
        for (int l = 0; l < L; l++) stage_load<T, N, E, W>(v[l], j, c, sm[l]);
        bar.sync();
        cta_fft_head<T, N, E, DIR, W, L, Ns * R>(v, j, c, sm, tw, bar);
    }
}

// Two lines per thread: software-pipelined exchange.  Line A's shared-memory read latency is covered by
// line B's stores and vice versa, and each barrier serves one line's write->read and the other line's
// read->write hazard, so the barrier count is unchanged (2 per stage) while the post-barrier bubble shrinks.
// Precondition of a round: A's stage-Ns outputs are stored (not yet visible), B's are still in registers.
template <typename T, int N, int E, int DIR, int W, int Ns, typename C, typename TWP>
SGPE_DI void cta_fft2_round(C (&v)[2][E], int j, int c, C* const (&sm)[2], TWP tw) {
    constexpr int R = (N / Ns) < E ? (N / Ns) : E;
    constexpr int Ns2 = Ns * R;
    constexpr int R2 = (N / Ns2) < E ? (N / Ns2) : E;
    constexpr bool more = (Ns2 * R2 < N);
    __syncthreads();
    stage_load<T, N, E, W>(v[0], j, c, sm[0]);
    stage_store<T, N, E, W, Ns>(v[1], j, c, sm[1]);
    stage_compute<T, N, E, DIR, Ns2>(v[0], j, tw);
    __syncthreads();
    stage_load<T, N, E, W>(v[1], j, c, sm[1]);
    if constexpr (more) stage_store<T, N, E, W, Ns2>(v[0], j, c, sm[0]);
    stage_compute<T, N, E, DIR, Ns2>(v[1], j, tw);
    if constexpr (more) cta_fft2_round<T, N, E, DIR, W, Ns2>(v, j, c, sm, tw);
}

template <typename T, int N, int E, int DIR, int W, int L, typename C, typename TWP>
SGPE_DI void cta_fft(C (&v)[L][E], int j, int c, C* const (&sm)[L], TWP tw) {
    if constexpr (L == 2 && E < N) {
        // (a following cta_fft of the same two lines may start right away: it touches line A's image, whose
        // readers are past the last barrier, before its own first barrier, and line B's only after it)
        stage_compute<T, N, E, DIR, 1>(v[0], j, tw);
        stage_store<T, N, E, W, 1>(v[0], j, c, sm[0]);
        stage_compute<T, N, E, DIR, 1>(v[1], j, tw);
        cta_fft2_round<T, N, E, DIR, W, 1>(v, j, c, sm, tw);
    } else {
        cta_fft_from<T, N, E, DIR, W, L, 1>(v, j, c, sm, tw, CtaBar());
    }
}
// one line per thread, synchronising over a barrier group
template <typename T, int N, int E, int DIR, int W, typename C, typename B, typename TWP>
SGPE_DI void group_fft(C (&v)[1][E], int j, int c, C* const (&sm)[1], TWP tw, const B& bar) {
    cta_fft_from<T, N, E, DIR, W, 1, 1>(v, j, c, sm, tw, bar);
}

// ---- split exchange: the real and the imaginary parts of a line travel through ONE real-valued image of the line,
// one after the other (twice the barriers, half the shared memory of the complex image).  What the persistent passes
// use for their second transform, so that the complex image can hold the asynchronously staged NEXT tile meanwhile.
template <typename T, int N, int E, int W>
struct SplitGeom {
    static constexpr int PHASE = 128 / (int)sizeof(T);          // 8-byte (double) / 4-byte (float) elements per wavefront
    static constexpr int MASK = (PHASE / W > 1) ? (PHASE / W - 1) : 0;
    static constexpr int SH = LineGeom<T, N, E, W>::SH;
    SGPE_DI static int swz(int n) { return n ^ ((n >> SH) & MASK); }
};
template <typename T, int N, int E, int W, int Ns, int PART, typename C>
SGPE_DI void stage_store_part(const C (&v)[E], int j, int c, T* sm) {
    typedef SplitGeom<T, N, E, W> G;
    constexpr int R = (N / Ns) < E ? (N / Ns) : E;
    constexpr int NB = E / R;
    constexpr int NT = N / E;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int jb = j + b * NT;
        const int k = jb & (Ns - 1);
        const int j0 = (jb - k) * R + k;
#pragma unroll
        for (int t = 0; t < R; t++) sm[G::swz(j0 + t * Ns) * W + c] = PART ? v[b + t * NB].y : v[b + t * NB].x;
    }
}
template <typename T, int N, int E, int W, int PART, typename C>
SGPE_DI void stage_load_part(C (&v)[E], int j, int c, const T* sm) {
    typedef SplitGeom<T, N, E, W> G;
    constexpr int NT = N / E;
#pragma unroll
    for (int m = 0; m < E; m++) {
        if (PART) v[m].y = sm[G::swz(j + m * NT) * W + c];
        else v[m].x = sm[G::swz(j + m * NT) * W + c];
    }
}
// full transform of L lines per thread through real images sm[l] (N * W reals each); ends right after the last
// butterflies (the images' readers are behind a barrier only when another exchange followed)
template <typename T, int N, int E, int DIR, int W, int L, int Ns, typename C, typename B, typename TWP>
SGPE_DI void cta_fft_split_from(C (&v)[L][E], int j, int c, T* const (&sm)[L], TWP tw, const B& bar) {
    constexpr int R = (N / Ns) < E ? (N / Ns) : E;
#pragma unroll
    for (int l = 0; l < L; l++) stage_compute<T, N, E, DIR, Ns>(v[l], j, tw);
    if constexpr (Ns * R < N) {
        bar.sync();                         // the previous readers of the images are done
#pragma unroll
        for (int l = 0; l < L; l++) stage_store_part<T, N, E, W, Ns, 0>(v[l], j, c, sm[l]);
        bar.sync();
#pragma unroll
        for (int l = 0; l < L; l++) stage_load_part<T, N, E, W, 0>(v[l], j, c, sm[l]);
        bar.sync();
#pragma unroll
        for (int l = 0; l < L; l++) stage_store_part<T, N, E, W, Ns, 1>(v[l], j, c, sm[l]);
        bar.sync();
#pragma unroll
        for (int l = 0; l < L; l++) stage_load_part<T, N, E, W, 1>(v[l], j, c, sm[l]);
        cta_fft_split_from<T, N, E, DIR, W, L, Ns * R>(v, j, c, sm, tw, bar);
    }
}

// ---- deterministic CTA reduction of NV doubles (warp shuffle, then shared memory in warp order)
template <int NV>
SGPE_DI void cta_reduce(double (&val)[NV], double* red /* >= 32*NV doubles of shared memory */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double x = val[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        val[i] = x;
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) red[warp * NV + i] = val[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double s = 0.0;
        for (int w = 0; w < nwarps; w++) s += red[w * NV + i];
        val[i] = s;       // every thread holds the CTA total
    }
    __syncthreads();
}

// the same over one barrier group of a CTA: `tid` / `nthreads` are relative to the group, `red` is the group's own
template <int NV, typename B>
SGPE_DI void group_reduce(double (&val)[NV], double* red, int tid, int nthreads, const B& bar) {
    const int lane = tid & 31, warp = tid >> 5;
    const int nwarps = (nthreads + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double x = val[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        val[i] = x;
    }
    bar.sync();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) red[warp * NV + i] = val[i];
    }
    bar.sync();
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double s = 0.0;
        for (int w = 0; w < nwarps; w++) s += red[w * NV + i];
        val[i] = s;
    }
    bar.sync();
}

}  // namespace sgpe
