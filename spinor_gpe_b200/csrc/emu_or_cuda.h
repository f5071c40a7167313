// emu_or_cuda.h — the product build includes the CUDA runtime.  A test-only build (tests/emu, macro
// SGPE_EMU) swaps in a single-threaded fibre model of a CTA so the very same kernel sources can be
// checked on a CPU-only machine.  The emulated library is never loaded by the spinor_gpe_b200 package.
#pragma once
#ifdef SGPE_EMU
#include "cuda_emu.h"
#define SGPE_PREFETCH_L2(ptr) ((void)(ptr))
#define SGPE_PREFETCH_L1(ptr) ((void)(ptr))
#define SGPE_NANOSLEEP(ns) ((void)(ns))
#define SGPE_SMID() 0u
#define SGPE_LD_STREAM(p) (*(p))
#define SGPE_ST_STREAM(p, v) (*(p) = (v))
#define SGPE_GLOBALTIMER() 0ull
#define SGPE_NAMED_BAR(id, count) ::emu::named_barrier((id), (count))
#define SGPE_DYN_SMEM(name) unsigned char* name = ::emu::dyn_smem()
#define SGPE_LAUNCH(kern, grid, block, smem, stream, ...) \
    ::emu::launch((grid), (block), (smem), [&]() { kern(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
#define SGPE_NANOSLEEP(ns) __nanosleep(ns)
__device__ __forceinline__ unsigned sgpe_smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define SGPE_SMID() sgpe_smid()
// the state is touched once per pass: keep it out of L1 (the twiddle / factor tables live there)
#define SGPE_LD_STREAM(p) __ldcs(p)
#define SGPE_ST_STREAM(p, v) __stcs((p), (v))
__device__ __forceinline__ unsigned long long sgpe_gtimer() { unsigned long long r; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(r)); return r; }
#define SGPE_GLOBALTIMER() sgpe_gtimer()
// barrier `id` (1..15; 0 is __syncthreads) over `count` threads, a multiple of the warp size
#define SGPE_NAMED_BAR(id, count) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory")
#define SGPE_PREFETCH_L1(ptr) asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr))
#define SGPE_PREFETCH_L2(ptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr))
#define SGPE_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define SGPE_LAUNCH(kern, grid, block, smem, stream, ...) \
    kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
