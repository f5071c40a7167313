// launch_inst.cu — compiled once per transform length (-DSGPE_N=<N>): instantiates the row and column
// passes for that length in both precisions and both time modes and exports plain launch functions.
#include "kernels.cuh"
#include "launch.h"

#ifndef SGPE_N
#error "compile with -DSGPE_N=<transform length>"
#endif

namespace sgpe {

// elements per thread of the row / line passes: 8 (three exchanges at 2048 points); complex64 in imaginary time runs 16
// (two exchanges, 64 data registers, four CTAs per SM): 55.9 -> 50.2 us at 2048^2, while with the sincos of real time in
// the point-wise phase the same change loses 4 % (profiles/r02_variants.md)
template <typename T, int N, int TM = TM_REAL> struct RowCfg {
    static constexpr int E = (sizeof(T) == 4 && N >= 256 && TM == TM_IMAG) ? 16 : 8;
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

// CTAs of this kernel resident on the whole device = how far ahead (in block index) the next work item
// of an SM lies; used for the L2 prefetch of the next tile
template <typename K> static int resident_ctas(K kern, int threads, size_t smem) {
#ifndef SGPE_EMU
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess) return 0;
    return per_sm * sms;
#else
    (void)kern; (void)threads; (void)smem;
    return 0;
#endif
}

template <typename K> static void allow_smem(K kern, size_t bytes) {
#ifndef SGPE_EMU
    if (bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
#else
    (void)kern; (void)bytes;
#endif
}

template <typename T, int N, int TM>
static int launch_row_split_t(const RowArgs<T>& a, int batch, cudaStream_t st) {
    constexpr int E = 8, NT = N / E;
    constexpr int RPC = (2 * NT >= 256) ? 1 : (256 / (2 * NT));
    constexpr size_t smem = (size_t)RPC * 2 * N * sizeof(typename cx_of<T>::type);
    if (a.ny % RPC != 0) return -2;
    static bool once = false;
    if (!once) { allow_smem(row_pass_split<T, N, E, RPC, TM>, smem); once = true; }
    dim3 grid(a.ny / RPC, batch), block(RPC * 2 * NT);
    SGPE_LAUNCH((row_pass_split<T, N, E, RPC, TM>), grid, block, smem, st, a);
    return 0;
}

template <typename T, int N, int TM>
static int launch_row_t(const RowArgs<T>& a, int batch, cudaStream_t st) {
    typedef RowCfg<T, N, TM> Cfg;
    if (a.ny % Cfg::RPC != 0) return -2;
    static bool once = false;
    static int ahead = 0;
    if (!once) {
        allow_smem(row_pass<T, N, Cfg::E, Cfg::RPC, TM, 0>, Cfg::SMEM);
        allow_smem(row_pass<T, N, Cfg::E, Cfg::RPC, TM, 1>, Cfg::SMEM);
        allow_smem(row_pass<T, N, Cfg::E, Cfg::RPC, TM, 2>, Cfg::SMEM);
        allow_smem(row_pass<T, N, Cfg::E, Cfg::RPC, TM, 3>, Cfg::SMEM);
        ahead = resident_ctas(row_pass<T, N, Cfg::E, Cfg::RPC, TM, 1>, Cfg::THREADS, Cfg::SMEM);
        once = true;
    }
    dim3 grid(a.ny / Cfg::RPC, batch), block(Cfg::THREADS);
    RowArgs<T> a2 = a;
    if (a2.prefetch_ahead) a2.prefetch_ahead = ahead;
    a2.resident = ahead;
    const bool fast = a.do_inv && a.do_pw && a.do_fwd && a.cpl_mode == 0 && !a.sign_in &&
                      !a.sign_out && a.scale_out == 1.0 && a.stagger_ns == 0 && a.dbg == nullptr && a.sc.mode == 0 &&
                      a.scale_tot == nullptr && a.maxbits == nullptr && !a.polar;
    // the stand-alone inverse of per-step energy tracking: device-side scale, density maxima, polar store
    const bool invp = a.do_inv && !a.do_pw && !a.do_fwd && !a.sign_in && a.stagger_ns == 0 && a.dbg == nullptr &&
                      a.sc.mode == 0 && a.scale_tot != nullptr && a.maxbits != nullptr && a.polar;
    if (invp) {
        SGPE_LAUNCH((row_pass<T, N, Cfg::E, Cfg::RPC, TM, 3>), grid, block, Cfg::SMEM, st, a2);
    } else if (fast && a.pot_mode == 1) {
        SGPE_LAUNCH((row_pass<T, N, Cfg::E, Cfg::RPC, TM, 1>), grid, block, Cfg::SMEM, st, a2);
    } else if (fast && a.pot_mode == 0) {
        SGPE_LAUNCH((row_pass<T, N, Cfg::E, Cfg::RPC, TM, 2>), grid, block, Cfg::SMEM, st, a2);
    } else {
        SGPE_LAUNCH((row_pass<T, N, Cfg::E, Cfg::RPC, TM, 0>), grid, block, Cfg::SMEM, st, a2);
    }
    return 0;
}

template <typename T, int N, int TM, int W, int E>
static int launch_col_w(const ColArgs<T>& a, int batch, cudaStream_t st) {
    typedef ColCfg<T, N> Cfg;
    constexpr size_t smem = (size_t)N * W * Cfg::CB + 32 * 4 * sizeof(double);
    if (a.nx % W != 0) return -2;
    static bool once = false;
    static int ahead = 0;
    if (!once) {
        allow_smem(col_pass<T, N, E, W, TM, 0>, smem);
        allow_smem(col_pass<T, N, E, W, TM, 1>, smem);
        allow_smem(col_pass<T, N, E, W, TM, 2>, smem);
        ahead = resident_ctas(col_pass<T, N, E, W, TM, 1>, W * (N / E), smem);
        once = true;
    }
    dim3 grid(2 * a.nx / W, batch), block(W * (N / E));
    ColArgs<T> a2 = a;
    if (a2.prefetch_ahead) a2.prefetch_ahead = ahead;
    const bool fast = a.do_fwd && a.do_inv && a.kin_mode == 1 && !a.sign_in && !a.sign_out && a.scale_out == 1.0 &&
                      a.dbg == nullptr;
    if (fast && a.aux != nullptr && a.has_a) {
        SGPE_LAUNCH((col_pass<T, N, E, W, TM, 2>), grid, block, smem, st, a2);
    } else if (fast) {
        SGPE_LAUNCH((col_pass<T, N, E, W, TM, 1>), grid, block, smem, st, a2);
    } else {
        SGPE_LAUNCH((col_pass<T, N, E, W, TM, 0>), grid, block, smem, st, a2);
    }
    return 0;
}

// persistent column pass with TMA-staged tiles (col_pass_p): one CTA per resident slot, grid capped at the tile count
template <typename T, int N, int TM, int W, int E, int XSPLIT, int TWS, int KM = 1>
static int launch_col_p(const ColArgs<T>& a, int batch, cudaStream_t st) {
    typedef ColCfg<T, N> Cfg;
    constexpr size_t tile = (size_t)N * W * Cfg::CB;
    constexpr size_t smem = tile + (XSPLIT ? tile / 2 : 0) + (TWS ? (size_t)N * Cfg::CB : 0) + 32 * 4 * sizeof(double) + 512;
    if (a.nx % W != 0 || a.tile_map == nullptr) return -2;
    if (smem > 227 * 1024) return -2;
    static bool once = false;
    static int resident = 0;
    if (!once) {
        allow_smem(col_pass_p<T, N, E, W, TM, XSPLIT, TWS, KM>, smem);
        resident = resident_ctas(col_pass_p<T, N, E, W, TM, XSPLIT, TWS, KM>, W * (N / E), smem);
        once = true;
    }
    const int ntiles = 2 * a.nx / W;
    int ctas = resident > 0 ? resident / (batch < 1 ? 1 : batch) : ntiles;      // the batch shares the device
    if (ctas < 1) ctas = 1;
    if (ctas > ntiles) ctas = ntiles;
    dim3 grid(ctas, batch), block(W * (N / E));
    const SgpeTileMap& map = *static_cast<const SgpeTileMap*>(a.tile_map);
    SGPE_LAUNCH((col_pass_p<T, N, E, W, TM, XSPLIT, TWS, KM>), grid, block, smem, st, map, a);
    return 0;
}

// persistent column pass, two barrier groups per CTA (col_pass_pg)
template <typename T, int N, int TM, int W, int E>
static int launch_col_pg(const ColArgs<T>& a, int batch, cudaStream_t st) {
    typedef ColCfg<T, N> Cfg;
    constexpr size_t tile = (size_t)N * W * Cfg::CB;
    constexpr size_t smem = tile + tile / 2 + 2 * 32 * 4 * sizeof(double) + 64;
    if (a.nx % W != 0 || a.tile_map == nullptr) return -2;
    static bool once = false;
    static int resident = 0;
    if (!once) {
        allow_smem(col_pass_pg<T, N, E, W, TM>, smem);
        resident = resident_ctas(col_pass_pg<T, N, E, W, TM>, W * (N / E), smem);
        once = true;
    }
    const int ntiles = 2 * a.nx / W;
    int ctas = resident > 0 ? resident / (batch < 1 ? 1 : batch) : ntiles;
    if (ctas < 1) ctas = 1;
    if (ctas > ntiles) ctas = ntiles;
    dim3 grid(ctas, batch), block(W * (N / E));
    const SgpeTileMap& map = *static_cast<const SgpeTileMap*>(a.tile_map);
    SGPE_LAUNCH((col_pass_pg<T, N, E, W, TM>), grid, block, smem, st, map, a);
    return 0;
}

// G barrier groups of W columns each in one CTA (col_pass<..., G>): same tile, same global segments, G instruction
// streams per SM
template <typename T, int N, int TM, int W, int E, int G>
static int launch_col_g(const ColArgs<T>& a, int batch, cudaStream_t st) {
    typedef ColCfg<T, N> Cfg;
    constexpr size_t smem = (size_t)N * W * G * Cfg::CB + (size_t)G * 32 * 4 * sizeof(double);
    if (a.nx % (W * G) != 0) return -2;
    static bool once = false;
    static int ahead = 0;
    if (!once) {
        allow_smem(col_pass<T, N, E, W, TM, 0, G>, smem);
        allow_smem(col_pass<T, N, E, W, TM, 1, G>, smem);
        ahead = resident_ctas(col_pass<T, N, E, W, TM, 1, G>, G * W * (N / E), smem);
        once = true;
    }
    dim3 grid(2 * a.nx / (W * G), batch), block(G * W * (N / E));
    ColArgs<T> a2 = a;
    if (a2.prefetch_ahead) a2.prefetch_ahead = ahead;
    const bool fast = a.do_fwd && a.do_inv && a.kin_mode == 1 && !a.sign_in && !a.sign_out && a.scale_out == 1.0 &&
                      a.aux == nullptr && a.dbg == nullptr;      // (the boundary-state store exists in the generic kernel only)
    if (fast) {
        SGPE_LAUNCH((col_pass<T, N, E, W, TM, 1, G>), grid, block, smem, st, a2);
    } else {
        SGPE_LAUNCH((col_pass<T, N, E, W, TM, 0, G>), grid, block, smem, st, a2);
    }
    return 0;
}

// wsel: 0 = default tile width (64-byte segments), 2 = narrow tiles (half the shared memory per CTA so
// that two CTAs share an SM and their load / compute / store phases overlap), 3 = default tile worked on by two
// independent barrier groups of half the width each
template <typename T, int N, int TM>
static int launch_col_t(const ColArgs<T>& a, int batch, int wsel, cudaStream_t st) {
    typedef ColCfg<T, N> Cfg;
    if constexpr (Cfg::W % 2 == 0 && (Cfg::W / 2) * Cfg::NT >= 32) {
        if (wsel == 3) return launch_col_g<T, N, TM, Cfg::W / 2, Cfg::E, 2>(a, batch, st);
    }
#ifdef SGPE_EXPERIMENTAL
    // measured slower on B200 (profiles/r01_variants.md); kept for experiments and the emulation tests
    constexpr int WN = Cfg::W / 2;
    if constexpr (WN >= 1 && WN * Cfg::NT >= 32) {
        if (wsel == 2) return launch_col_w<T, N, TM, WN, Cfg::E>(a, batch, st);
    }
    // wsel 8: radix-8 variant (8 elements per thread, twice the threads, one more exchange per transform)
    if constexpr (Cfg::E == 16 && Cfg::W * (N / 8) <= 1024) {
        if (wsel == 8) return launch_col_w<T, N, TM, Cfg::W, 8>(a, batch, st);
    }
#else
    if (wsel != 0 && wsel != 3) return -3;
#endif
    if constexpr (Cfg::E == 16) {
        // the persistent kernel covers the steady-state junction: forward + factors + inverse, separable tables
        const bool fast = a.do_fwd && a.do_inv && !a.sign_in && !a.sign_out && a.scale_out == 1.0 && a.in == a.out;
        // the stand-alone inverse of per-step energy tracking: persistent CTAs, staged tiles, no forward / K phase
        const bool inv_only = !a.do_fwd && a.do_inv && !a.has_a && !a.has_b && !a.sign_in && !a.sign_out &&
                              a.scale_out == 1.0 && a.in == a.out && a.tile_map != nullptr && wsel == 0;
        if (inv_only && a.kernel_sel >= 1) {
            int rci = -2;
            if constexpr (Cfg::W % 2 == 0 && (Cfg::W / 2) * Cfg::NT >= 32) {
                if (a.kernel_sel == 5) rci = launch_col_p<T, N, TM, Cfg::W / 2, Cfg::E, 1, 0, 2>(a, batch, st);
            }
            if (a.kernel_sel != 5) rci = launch_col_p<T, N, TM, Cfg::W, Cfg::E, 1, 0, 2>(a, batch, st);
            if (rci != -2) return rci;
        }
        // (the first persistent variant also evaluates dense kinetic grids; the others take factor tables only)
        if (a.kernel_sel >= 2 && a.kin_mode != 1) return launch_col_w<T, N, TM, Cfg::W, Cfg::E>(a, batch, st);
        if (a.kernel_sel == 1 && fast && a.tile_map != nullptr && wsel == 0)
            return a.kin_mode == 1 ? launch_col_p<T, N, TM, Cfg::W, Cfg::E, 1, 0, 1>(a, batch, st)
                                   : launch_col_p<T, N, TM, Cfg::W, Cfg::E, 1, 0, 0>(a, batch, st);
        if (a.kernel_sel == 2 && fast && a.tile_map != nullptr && wsel == 0)
            return launch_col_p<T, N, TM, Cfg::W, Cfg::E, 0, 0>(a, batch, st);
        if (a.kernel_sel == 4 && fast && a.tile_map != nullptr && wsel == 0) {
            int rc4 = launch_col_p<T, N, TM, Cfg::W, Cfg::E, 1, 1>(a, batch, st);
            return rc4 == -2 ? launch_col_p<T, N, TM, Cfg::W, Cfg::E, 1, 0>(a, batch, st) : rc4;
        }
        if constexpr (TM == TM_IMAG) {       // k factors of the launch in shared memory (imaginary time)
            if (a.kernel_sel == 6 && fast && a.tile_map != nullptr && wsel == 0) {
                int rc6 = launch_col_p<T, N, TM, Cfg::W, Cfg::E, 1, 2>(a, batch, st);
                return rc6 == -2 ? launch_col_p<T, N, TM, Cfg::W, Cfg::E, 1, 0>(a, batch, st) : rc6;
            }
        } else {
            if (a.kernel_sel == 6 && fast && a.tile_map != nullptr && wsel == 0)
                return launch_col_p<T, N, TM, Cfg::W, Cfg::E, 1, 0>(a, batch, st);
        }
        if constexpr (Cfg::W % 2 == 0 && (Cfg::W / 2) * Cfg::NT >= 32) {
            if (a.kernel_sel == 5 && fast && a.tile_map != nullptr && wsel == 0)
                return launch_col_p<T, N, TM, Cfg::W / 2, Cfg::E, 1, 0>(a, batch, st);
        }
        if constexpr (Cfg::W % 2 == 0 && (Cfg::W / 2) * Cfg::NT >= 32) {
            if (a.kernel_sel == 3 && fast && a.tile_map != nullptr && wsel == 0)
                return launch_col_pg<T, N, TM, Cfg::W, Cfg::E>(a, batch, st);
        }
    }
    return launch_col_w<T, N, TM, Cfg::W, Cfg::E>(a, batch, st);
}

template <typename T, int N, int TM>
static int launch_kline_t(const KLineArgs<T>& a, cudaStream_t st) {
    typedef RowCfg<T, N> Cfg;
    constexpr size_t smem = Cfg::SMEM + 32 * 4 * sizeof(double);
    if (a.ny % Cfg::RPC != 0 || a.wlines % Cfg::RPC != 0 || a.line0 % Cfg::RPC != 0) return -2;
    static bool once = false;
    if (!once) { allow_smem(kline_pass<T, N, Cfg::E, Cfg::RPC, TM>, smem); once = true; }
    KLineArgs<T> a2 = a;
    a2.nblk = a.wlines / Cfg::RPC;
    const int ctas = (a.max_ctas > 0 && a.max_ctas < a2.nblk) ? a.max_ctas : a2.nblk;
    dim3 grid(ctas), block(Cfg::THREADS);
    SGPE_LAUNCH((kline_pass<T, N, Cfg::E, Cfg::RPC, TM>), grid, block, smem, st, a2);
    return 0;
}

// k-space junction on the row-major k slab (fused-exchange layout): W adjacent columns of one component
template <typename T, int N> struct KColCfg {
    static constexpr int CB = (int)sizeof(typename cx_of<T>::type);
    static constexpr int E = (N >= 128) ? 16 : 8;                      // 128 = 16 x 8: one exchange per transform
    static constexpr int NT = N / E;
    static constexpr int WT = (256 / NT > 32) ? 32 : (256 / NT);       // aim at 256 threads, tiles <= 32 columns
    static constexpr int W = (ColCfg<T, N>::W > WT) ? ColCfg<T, N>::W : WT;
    static constexpr size_t SMEM = (size_t)N * W * CB + 32 * 4 * sizeof(double);
};
template <typename T, int N, int TM>
static int launch_kcol_t(const KColArgs<T>& a, cudaStream_t st) {
    typedef KColCfg<T, N> Cfg;
    if (a.inner % Cfg::W != 0 || a.wcount % Cfg::W != 0 || a.x0 % Cfg::W != 0) return -2;
    static bool once = false;
    if (!once) { allow_smem(kcol_pass<T, N, Cfg::E, Cfg::W, TM>, Cfg::SMEM); once = true; }
    dim3 grid(a.wcount / Cfg::W, a.groups, 2), block(Cfg::W * Cfg::NT);
    SGPE_LAUNCH((kcol_pass<T, N, Cfg::E, Cfg::W, TM>), grid, block, Cfg::SMEM, st, a);
    return 0;
}

// strided half of the four-step transform (only lengths that fit: 2 * N1 * W elements of shared memory)
template <typename T, int N, int TM>
static int launch_mid_t(const MidArgs<T>& a, int n2_tiles_unused, cudaStream_t st) {
    (void)n2_tiles_unused;
    if constexpr (N <= 1024) {
        constexpr int E = 8, NT = N / E;
        constexpr int CB = (int)sizeof(typename cx_of<T>::type);
        constexpr int W0 = (NT <= 16) ? (128 / NT) : 4;
        constexpr int W = (CB == 8 && W0 < 8) ? 8 : W0;          // complex64: keep 64-byte segments
        constexpr size_t smem = (size_t)2 * N * W * CB;
        if (a.n2 % W != 0) return -2;
        static bool once = false;
        if (!once) { allow_smem(mid_pass<T, N, E, W, TM>, smem); once = true; }
        MidArgs<T> a2 = a;
        dim3 grid(1, 1), block(W * NT);
        if (a.inner == 1) {       // lines y0 .. y0 + wcount, every line fully
            a2.x0 = 0; a2.wtiles = a.n2 / W; a2.wstride = 0; a2.nvb = a.n2 / W;
            grid.x = a2.nvb; grid.y = a.wcount;
        } else {                  // columns [x0, x0 + wcount) of every n2 digit of the column slab
            if (a.wcount % W != 0 || a.x0 % W != 0) return -2;
            a2.y0 = 0; a2.wtiles = a.wcount / W; a2.wstride = a.inner; a2.nvb = (a.n2 / a.inner) * a2.wtiles;
            grid.x = (a.max_ctas > 0 && a.max_ctas < a2.nvb) ? a.max_ctas : a2.nvb;
        }
        SGPE_LAUNCH((mid_pass<T, N, E, W, TM>), grid, block, smem, st, a2);
        return 0;
    } else {
        return -2;
    }
}

#define SGPE_CAT2(a, b) a##b
#define SGPE_CAT(a, b) SGPE_CAT2(a, b)

int SGPE_CAT(launch_row_, SGPE_N)(int dtype, int tm, const void* args, int batch, int mode, cudaStream_t st) {
    if (mode == 1) {      // split variant: one component per thread
#ifndef SGPE_EXPERIMENTAL
        return -3;
#else
        if (dtype == 0) {
            const RowArgs<double>& a = *static_cast<const RowArgs<double>*>(args);
            return tm == TM_REAL ? launch_row_split_t<double, SGPE_N, TM_REAL>(a, batch, st)
                                 : launch_row_split_t<double, SGPE_N, TM_IMAG>(a, batch, st);
        }
        const RowArgs<float>& a = *static_cast<const RowArgs<float>*>(args);
        return tm == TM_REAL ? launch_row_split_t<float, SGPE_N, TM_REAL>(a, batch, st)
                             : launch_row_split_t<float, SGPE_N, TM_IMAG>(a, batch, st);
#endif
    }
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

int SGPE_CAT(launch_kline_, SGPE_N)(int dtype, int tm, const void* args, cudaStream_t st) {
    if (dtype == 0) {
        const KLineArgs<double>& a = *static_cast<const KLineArgs<double>*>(args);
        return tm == TM_REAL ? launch_kline_t<double, SGPE_N, TM_REAL>(a, st) : launch_kline_t<double, SGPE_N, TM_IMAG>(a, st);
    }
    const KLineArgs<float>& a = *static_cast<const KLineArgs<float>*>(args);
    return tm == TM_REAL ? launch_kline_t<float, SGPE_N, TM_REAL>(a, st) : launch_kline_t<float, SGPE_N, TM_IMAG>(a, st);
}

int SGPE_CAT(launch_kcol_, SGPE_N)(int dtype, int tm, const void* args, cudaStream_t st) {
    if (dtype == 0) {
        const KColArgs<double>& a = *static_cast<const KColArgs<double>*>(args);
        return tm == TM_REAL ? launch_kcol_t<double, SGPE_N, TM_REAL>(a, st) : launch_kcol_t<double, SGPE_N, TM_IMAG>(a, st);
    }
    const KColArgs<float>& a = *static_cast<const KColArgs<float>*>(args);
    return tm == TM_REAL ? launch_kcol_t<float, SGPE_N, TM_REAL>(a, st) : launch_kcol_t<float, SGPE_N, TM_IMAG>(a, st);
}
int SGPE_CAT(kcol_tile_width_, SGPE_N)(int dtype) {
    return dtype == 0 ? KColCfg<double, SGPE_N>::W : KColCfg<float, SGPE_N>::W;
}

int SGPE_CAT(launch_mid_, SGPE_N)(int dtype, int tm, const void* args, cudaStream_t st) {
    if (dtype == 0) {
        const MidArgs<double>& a = *static_cast<const MidArgs<double>*>(args);
        return tm == TM_REAL ? launch_mid_t<double, SGPE_N, TM_REAL>(a, 0, st) : launch_mid_t<double, SGPE_N, TM_IMAG>(a, 0, st);
    }
    const MidArgs<float>& a = *static_cast<const MidArgs<float>*>(args);
    return tm == TM_REAL ? launch_mid_t<float, SGPE_N, TM_REAL>(a, 0, st) : launch_mid_t<float, SGPE_N, TM_IMAG>(a, 0, st);
}

int SGPE_CAT(col_tile_width_, SGPE_N)(int dtype) {
    return dtype == 0 ? ColCfg<double, SGPE_N>::W : ColCfg<float, SGPE_N>::W;
}

}  // namespace sgpe
