// emu_or_cuda.h — the product build includes the CUDA runtime.  A test-only build (tests/emu, macro
// SGPE_EMU) swaps in a single-threaded fibre model of a CTA so the very same kernel sources can be
// checked on a CPU-only machine.  The emulated library is never loaded by the spinor_gpe_b200 package.
#pragma once
#ifdef SGPE_EMU
#include "cuda_emu.h"
#define SGPE_PREFETCH_L2(ptr) ((void)(ptr))
#define SGPE_NANOSLEEP(ns) ((void)(ns))
#define SGPE_DYN_SMEM(name) unsigned char* name = ::emu::dyn_smem()
#define SGPE_LAUNCH(kern, grid, block, smem, stream, ...) \
    ::emu::launch((grid), (block), (smem), [&]() { kern(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
#define SGPE_NANOSLEEP(ns) __nanosleep(ns)
#define SGPE_PREFETCH_L2(ptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr))
#define SGPE_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define SGPE_LAUNCH(kern, grid, block, smem, stream, ...) \
    kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
