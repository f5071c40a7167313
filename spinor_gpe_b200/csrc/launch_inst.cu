// launch_inst.cu — compiled once per transform length (-DSGPE_N=<N>): instantiates the row and column
// passes for that length in both precisions and both time modes and exports plain launch functions.
#include "kernels.cuh"
#include "launch.h"

#ifndef SGPE_N
#error "compile with -DSGPE_N=<transform length>"
#endif

namespace sgpe {

template <typename T, int N> struct RowCfg {
    static constexpr int E = 8;
    static constexpr int NT = N / E;
    static constexpr int RPC = (NT >= 128) ? 1 : (128 / NT);          // >= 128 threads per CTA
    static constexpr int THREADS = RPC * NT;
    static constexpr size_t SMEM = (size_t)RPC * 2 * N * sizeof(typename cx_of<T>::type);
};

template <typename T, int N> struct ColCfg {
    static constexpr int CB = (int)sizeof(typename cx_of<T>::type);
    static constexpr int E = (N >= 256) ? 16 : 8;
    static constexpr int NT = N / E;
    static constexpr int W0 = 64 / CB;                                 // 64-byte global segments
    static constexpr int WCAP = (128 * 1024) / (N * CB);               // tile <= 128 KiB of shared memory
    static constexpr int W1 = (W0 < WCAP) ? W0 : WCAP;
    static constexpr int W = (W1 * NT < 32) ? (32 / NT) : W1;          // at least one full warp
    static constexpr int THREADS = W * NT;
    static constexpr size_t SMEM = (size_t)N * W * CB + 32 * 4 * sizeof(double);
};

template <typename K> static void allow_smem(K kern, size_t bytes) {
#ifndef SGPE_EMU
    if (bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
#else
    (void)kern; (void)bytes;
#endif
}

template <typename T, int N, int TM>
static int launch_row_t(const RowArgs<T>& a, int batch, cudaStream_t st) {
    typedef RowCfg<T, N> Cfg;
    if (a.ny % Cfg::RPC != 0) return -2;
    static bool once = false;
    if (!once) { allow_smem(row_pass<T, N, Cfg::E, Cfg::RPC, TM>, Cfg::SMEM); once = true; }
    dim3 grid(a.ny / Cfg::RPC, batch), block(Cfg::THREADS);
    SGPE_LAUNCH((row_pass<T, N, Cfg::E, Cfg::RPC, TM>), grid, block, Cfg::SMEM, st, a);
    return 0;
}

template <typename T, int N, int TM, int W>
static int launch_col_w(const ColArgs<T>& a, int batch, cudaStream_t st) {
    typedef ColCfg<T, N> Cfg;
    constexpr size_t smem = (size_t)N * W * Cfg::CB + 32 * 4 * sizeof(double);
    if (a.nx % W != 0) return -2;
    static bool once = false;
    if (!once) { allow_smem(col_pass<T, N, Cfg::E, W, TM>, smem); once = true; }
    dim3 grid(2 * a.nx / W, batch), block(W * Cfg::NT);
    SGPE_LAUNCH((col_pass<T, N, Cfg::E, W, TM>), grid, block, smem, st, a);
    return 0;
}

// wsel: 0 = default tile width (64-byte segments), 2 = narrow tiles (half the shared memory per CTA so
// that two CTAs share an SM and their load / compute / store phases overlap)
template <typename T, int N, int TM>
static int launch_col_t(const ColArgs<T>& a, int batch, int wsel, cudaStream_t st) {
    typedef ColCfg<T, N> Cfg;
    constexpr int WN = Cfg::W / 2;
    if constexpr (WN >= 1 && WN * Cfg::NT >= 32) {
        if (wsel == 2) return launch_col_w<T, N, TM, WN>(a, batch, st);
    }
    return launch_col_w<T, N, TM, Cfg::W>(a, batch, st);
}

#define SGPE_CAT2(a, b) a##b
#define SGPE_CAT(a, b) SGPE_CAT2(a, b)

int SGPE_CAT(launch_row_, SGPE_N)(int dtype, int tm, const void* args, int batch, cudaStream_t st) {
    if (dtype == 0) {
        const RowArgs<double>& a = *static_cast<const RowArgs<double>*>(args);
        return tm == TM_REAL ? launch_row_t<double, SGPE_N, TM_REAL>(a, batch, st)
                             : launch_row_t<double, SGPE_N, TM_IMAG>(a, batch, st);
    }
    const RowArgs<float>& a = *static_cast<const RowArgs<float>*>(args);
    return tm == TM_REAL ? launch_row_t<float, SGPE_N, TM_REAL>(a, batch, st)
                         : launch_row_t<float, SGPE_N, TM_IMAG>(a, batch, st);
}

int SGPE_CAT(launch_col_, SGPE_N)(int dtype, int tm, const void* args, int batch, int wsel, cudaStream_t st) {
    if (dtype == 0) {
        const ColArgs<double>& a = *static_cast<const ColArgs<double>*>(args);
        return tm == TM_REAL ? launch_col_t<double, SGPE_N, TM_REAL>(a, batch, wsel, st)
                             : launch_col_t<double, SGPE_N, TM_IMAG>(a, batch, wsel, st);
    }
    const ColArgs<float>& a = *static_cast<const ColArgs<float>*>(args);
    return tm == TM_REAL ? launch_col_t<float, SGPE_N, TM_REAL>(a, batch, wsel, st)
                         : launch_col_t<float, SGPE_N, TM_IMAG>(a, batch, wsel, st);
}

int SGPE_CAT(col_tile_width_, SGPE_N)(int dtype) {
    return dtype == 0 ? ColCfg<double, SGPE_N>::W : ColCfg<float, SGPE_N>::W;
}

}  // namespace sgpe
