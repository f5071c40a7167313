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
// ---- asynchronous tile staging (TMA + mbarrier on the device): the emulation copies at issue time
#define SGPE_GRID_CONSTANT
#define SGPE_TID_X() ((int)threadIdx.x)
#define SGPE_OPAQUE(x) ((void)(x))
#define SGPE_DYN_SMEM_128(name) unsigned char* name = ::emu::dyn_smem()
// (the hardware reciprocal seed carries ~20 bits: the emulation truncates 1 / d to the upper word as well)
inline double sgpe_rcp_seed_emu(double d) { double r = 1.0 / d; unsigned long long u; memcpy(&u, &r, 8); u &= 0xffffffff00000000ull; memcpy(&r, &u, 8); return r; }
#define SGPE_RCP_SEED(d) sgpe_rcp_seed_emu(d)
inline double sgpe_rsqrt_seed_emu(double d) { double r = 1.0 / sqrt(d); unsigned long long u; memcpy(&u, &r, 8); u &= 0xffffffff00000000ull; memcpy(&r, &u, 8); return r; }
#define SGPE_RSQRT_SEED(d) sgpe_rsqrt_seed_emu(d)
#define SGPE_D2I_RN(x) ((int)nearbyint(x))
struct SgpeTileMap { const unsigned char* base; long long row_stride; int elem_bytes; };
struct SgpeMbar { int pending; unsigned phase; };
inline void sgpe_mbar_init(SgpeMbar* b, unsigned) { b->pending = 0; b->phase = 0; }
inline void sgpe_mbar_expect_tx(SgpeMbar* b, unsigned bytes) { b->pending += (int)bytes; }
inline void sgpe_mbar_complete(SgpeMbar* b, unsigned bytes) { b->pending -= (int)bytes; if (b->pending == 0) b->phase ^= 1u; }
inline void sgpe_mbar_wait(SgpeMbar* b, unsigned parity) { while ((b->phase & 1u) == parity) ::emu::yield(); }
inline void sgpe_fence_proxy_async() {}
// box of `rows` x `row_bytes` whose first element is (c0 [elements], c1 [row]) of the mapped array
inline void sgpe_tma_load_2d(void* dst, const SgpeTileMap* m, int c0, int c1, SgpeMbar* bar, int rows, int row_bytes) {
    for (int r = 0; r < rows; r++)
        memcpy((unsigned char*)dst + (size_t)r * row_bytes,
               m->base + (long long)(c1 + r) * m->row_stride + (long long)c0 * m->elem_bytes, (size_t)row_bytes);
    sgpe_mbar_complete(bar, (unsigned)(rows * row_bytes));
}
inline void sgpe_bulk_load(void* dst, const void* src, unsigned bytes, SgpeMbar* bar) {
    memcpy(dst, src, bytes);
    sgpe_mbar_complete(bar, bytes);
}
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
// ---- asynchronous tile staging: TMA (cp.async.bulk[.tensor]) global -> shared, completion on an mbarrier.
// The tile descriptor is a CUtensorMap encoded on the host (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint, no
// link-time dependency on libcuda) and passed as a __grid_constant__ kernel parameter.
#include <cuda.h>
#define SGPE_GRID_CONSTANT __grid_constant__
// threadIdx.x through a volatile read: the value (and everything derived from it) is re-materialised where it is asked
// for instead of being kept in registers across a persistent loop
__device__ __forceinline__ int sgpe_tid_x() { int r; asm volatile("mov.u32 %0, %%tid.x;" : "=r"(r)); return r; }
#define SGPE_TID_X() sgpe_tid_x()
// the compiler forgets what it knows about an integer (no common sub-expressions across this point)
#define SGPE_OPAQUE(x) asm volatile("" : "+r"(x))
#define SGPE_DYN_SMEM_128(name) extern __shared__ __align__(128) unsigned char name[]
__device__ __forceinline__ double sgpe_rcp_seed(double d) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d)); return r; }
#define SGPE_RCP_SEED(d) sgpe_rcp_seed(d)
__device__ __forceinline__ double sgpe_rsqrt_seed(double d) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d)); return r; }
#define SGPE_RSQRT_SEED(d) sgpe_rsqrt_seed(d)
#define SGPE_D2I_RN(x) __double2int_rn(x)
struct alignas(64) SgpeTileMap { CUtensorMap m; };
typedef unsigned long long SgpeMbar;
__device__ __forceinline__ unsigned sgpe_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sgpe_mbar_init(SgpeMbar* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sgpe_smem_u32(b)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void sgpe_mbar_expect_tx(SgpeMbar* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sgpe_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sgpe_mbar_wait(SgpeMbar* b, unsigned parity) {
    unsigned ok = 0;
    const unsigned addr = sgpe_smem_u32(b);
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void sgpe_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sgpe_tma_load_2d(void* dst, const SgpeTileMap* m, int c0, int c1, SgpeMbar* bar, int, int) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sgpe_smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(m)), "r"(sgpe_smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void sgpe_bulk_load(void* dst, const void* src, unsigned bytes, SgpeMbar* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sgpe_smem_u32(dst)), "l"(src), "r"(bytes), "r"(sgpe_smem_u32(bar)) : "memory");
}
#endif
