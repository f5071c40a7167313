// launch.h — per-length launchers (defined in launch_inst.cu, one object per length).
#pragma once
#include "emu_or_cuda.h"

namespace sgpe {

#define SGPE_FOR_EACH_N(X) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096)

#define SGPE_DECL(N)                                                                               \
    int launch_row_##N(int dtype, int tm, const void* args, int batch, int mode, cudaStream_t st); \
    int launch_col_##N(int dtype, int tm, const void* args, int batch, int wsel, cudaStream_t st); \
    int launch_kline_##N(int dtype, int tm, const void* args, cudaStream_t st);                    \
    int launch_mid_##N(int dtype, int tm, const void* args, cudaStream_t st);                      \
    int launch_kcol_##N(int dtype, int tm, const void* args, cudaStream_t st);                     \
    int kcol_tile_width_##N(int dtype);                                                            \
    int col_tile_width_##N(int dtype);
SGPE_FOR_EACH_N(SGPE_DECL)
#undef SGPE_DECL

inline bool supported_length(int n) {
#define SGPE_CASE(N) if (n == N) return true;
    SGPE_FOR_EACH_N(SGPE_CASE)
#undef SGPE_CASE
    return false;
}
inline int launch_row(int n, int dtype, int tm, const void* args, int batch, int mode, cudaStream_t st) {
#define SGPE_CASE(N) if (n == N) return launch_row_##N(dtype, tm, args, batch, mode, st);
    SGPE_FOR_EACH_N(SGPE_CASE)
#undef SGPE_CASE
    return -1;
}
inline int launch_col(int n, int dtype, int tm, const void* args, int batch, int wsel, cudaStream_t st) {
#define SGPE_CASE(N) if (n == N) return launch_col_##N(dtype, tm, args, batch, wsel, st);
    SGPE_FOR_EACH_N(SGPE_CASE)
#undef SGPE_CASE
    return -1;
}
inline int launch_kline(int n, int dtype, int tm, const void* args, cudaStream_t st) {
#define SGPE_CASE(N) if (n == N) return launch_kline_##N(dtype, tm, args, st);
    SGPE_FOR_EACH_N(SGPE_CASE)
#undef SGPE_CASE
    return -1;
}
inline int launch_mid(int n, int dtype, int tm, const void* args, cudaStream_t st) {
#define SGPE_CASE(N) if (n == N) return launch_mid_##N(dtype, tm, args, st);
    SGPE_FOR_EACH_N(SGPE_CASE)
#undef SGPE_CASE
    return -1;
}
inline int launch_kcol(int n, int dtype, int tm, const void* args, cudaStream_t st) {
#define SGPE_CASE(N) if (n == N) return launch_kcol_##N(dtype, tm, args, st);
    SGPE_FOR_EACH_N(SGPE_CASE)
#undef SGPE_CASE
    return -1;
}
inline int kcol_tile_width(int n, int dtype) {
#define SGPE_CASE(N) if (n == N) return kcol_tile_width_##N(dtype);
    SGPE_FOR_EACH_N(SGPE_CASE)
#undef SGPE_CASE
    return -1;
}
inline int col_tile_width(int n, int dtype) {
#define SGPE_CASE(N) if (n == N) return col_tile_width_##N(dtype);
    SGPE_FOR_EACH_N(SGPE_CASE)
#undef SGPE_CASE
    return -1;
}

}  // namespace sgpe
