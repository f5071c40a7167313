// cuda_emu.h — TEST INFRASTRUCTURE ONLY.  A minimal, deterministic model of the CUDA execution
// constructs the sgpe kernels use, so that the unmodified kernel sources can be compiled with g++ and
// run on a CPU-only box: every thread of a CTA is a ucontext fibre, __syncthreads() and warp shuffles
// are cooperative yield points, CTAs run one after another.  It models semantics, not performance.
// Nothing in the product (spinor_gpe_b200) includes or links this.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct alignas(16) double2 { double x, y; };
struct alignas(8) float2 { float x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

namespace emu {

constexpr size_t kStack = 256 * 1024;

struct State {
    uint3 tid, bid;
    dim3 bdim, gdim;
    std::vector<unsigned char> smem;
    // scheduler
    ucontext_t sched;
    std::vector<ucontext_t> fib;
    std::vector<std::vector<unsigned char>> stacks;
    std::vector<char> done;
    int cur = 0, nthreads = 0, alive = 0;
    int bar_arrived = 0;
    unsigned bar_gen = 0;
    int nbar_arrived[16] = {0};
    unsigned nbar_gen[16] = {0};
    std::vector<int> warp_arrived;
    std::vector<unsigned> warp_gen;
    std::vector<double> shfl_d[2];
    std::vector<unsigned long long> shfl_u[2];
    std::function<void()> body;
};
inline State& S() { static State s; return s; }

inline unsigned char* dyn_smem() { return S().smem.data(); }

inline void yield() {
    State& s = S();
    int me = s.cur;
    swapcontext(&s.fib[me], &s.sched);
}
inline void set_ids(int t) {
    State& s = S();
    s.cur = t;
    s.tid.x = t % s.bdim.x;
    s.tid.y = (t / s.bdim.x) % s.bdim.y;
    s.tid.z = t / (s.bdim.x * s.bdim.y);
}
inline void fibre_entry() {
    State& s = S();
    s.body();
    s.done[s.cur] = 1;
    s.alive--;
    swapcontext(&s.fib[s.cur], &s.sched);
}
inline void syncthreads() {
    State& s = S();
    unsigned g = s.bar_gen;
    if (++s.bar_arrived == s.nthreads) { s.bar_arrived = 0; s.bar_gen++; return; }
    while (s.bar_gen == g) yield();
}
// bar.sync id, count: the first `count` arrivals at barrier `id` form one generation
inline void named_barrier(int id, int count) {
    State& s = S();
    if (id < 0 || id >= 16) { fprintf(stderr, "emu: barrier id %d\n", id); abort(); }
    unsigned g = s.nbar_gen[id];
    if (++s.nbar_arrived[id] == count) { s.nbar_arrived[id] = 0; s.nbar_gen[id]++; return; }
    while (s.nbar_gen[id] == g) yield();
}
inline void warp_sync() {
    State& s = S();
    int w = s.cur / 32;
    int lanes = std::min(32, s.nthreads - w * 32);
    unsigned g = s.warp_gen[w];
    if (++s.warp_arrived[w] == lanes) { s.warp_arrived[w] = 0; s.warp_gen[w]++; return; }
    while (s.warp_gen[w] == g) yield();
}
inline double shfl_xor(double v, int m) {
    State& s = S();
    int w = s.cur / 32;
    int p = s.warp_gen[w] & 1;
    s.shfl_d[p][s.cur] = v;
    warp_sync();
    return s.shfl_d[p][s.cur ^ m];
}

// value of lane `src` of the caller's warp (every lane of the warp calls)
inline double shfl_from(double v, int src) {
    State& s = S();
    int w = s.cur / 32;
    int p = s.warp_gen[w] & 1;
    s.shfl_d[p][s.cur] = v;
    warp_sync();
    return s.shfl_d[p][w * 32 + src];
}

template <typename F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F f) {
    State& s = S();
    s.bdim = block; s.gdim = grid;
    s.nthreads = block.x * block.y * block.z;
    if (s.nthreads % 32 != 0) { fprintf(stderr, "emu: block size %d not a multiple of 32\n", s.nthreads); abort(); }
    s.smem.assign(smem_bytes + 64, 0);
    s.fib.resize(s.nthreads);
    if ((int)s.stacks.size() < s.nthreads) s.stacks.resize(s.nthreads);
    for (auto& st : s.stacks) if (st.size() != kStack) st.resize(kStack);
    s.done.assign(s.nthreads, 0);
    s.warp_arrived.assign((s.nthreads + 31) / 32, 0);
    s.warp_gen.assign((s.nthreads + 31) / 32, 0);
    for (int p = 0; p < 2; p++) { s.shfl_d[p].assign(s.nthreads, 0.0); }
    s.body = f;
    for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++) {
        s.bid.x = bx; s.bid.y = by; s.bid.z = bz;
        s.alive = s.nthreads; s.bar_arrived = 0;
        for (int q = 0; q < 16; q++) s.nbar_arrived[q] = 0;
        std::fill(s.done.begin(), s.done.end(), 0);
        std::fill(s.warp_arrived.begin(), s.warp_arrived.end(), 0);
        for (int t = 0; t < s.nthreads; t++) {
            getcontext(&s.fib[t]);
            s.fib[t].uc_stack.ss_sp = s.stacks[t].data();
            s.fib[t].uc_stack.ss_size = kStack;
            s.fib[t].uc_link = &s.sched;
            makecontext(&s.fib[t], (void (*)())fibre_entry, 0);
        }
        while (s.alive > 0) {
            for (int t = 0; t < s.nthreads; t++) {
                if (s.done[t]) continue;
                set_ids(t);
                swapcontext(&s.sched, &s.fib[t]);
            }
        }
    }
}
}  // namespace emu

#define threadIdx (::emu::S().tid)
#define blockIdx (::emu::S().bid)
#define blockDim (::emu::S().bdim)
#define gridDim (::emu::S().gdim)
inline void __syncthreads() { ::emu::syncthreads(); }
inline void __threadfence() {}
inline double __shfl_xor_sync(unsigned, double v, int m) { return ::emu::shfl_xor(v, m); }
inline double __shfl_up_sync(unsigned, double v, int d) { int l = ::emu::S().cur & 31; return ::emu::shfl_from(v, l >= d ? l - d : l); }
inline double __shfl_down_sync(unsigned, double v, int d) { int l = ::emu::S().cur & 31; return ::emu::shfl_from(v, l + d <= 31 ? l + d : l); }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline long long __double_as_longlong(double x) { long long r; std::memcpy(&r, &x, 8); return r; }
inline double __longlong_as_double(long long x) { double r; std::memcpy(&r, &x, 8); return r; }
template <typename T> inline T __ldcg(const T* p) { return *p; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline unsigned atomicMin(unsigned* p, unsigned v) { unsigned o = *p; if (v < o) *p = v; return o; }
inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
inline unsigned atomicCAS(unsigned* p, unsigned expect, unsigned v) { unsigned o = *p; if (o == expect) *p = v; return o; }
inline unsigned atomicExch(unsigned* p, unsigned v) { unsigned o = *p; *p = v; return o; }
inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline void sincos(double x, double* s, double* c) { *s = std::sin(x); *c = std::cos(x); }
inline void sincosf(float x, float* s, float* c) { *s = std::sin(x); *c = std::cos(x); }
inline void sincospi(double x, double* s, double* c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline int __double2loint(double x) { unsigned long long u; memcpy(&u, &x, 8); return (int)(unsigned)(u & 0xffffffffull); }
inline int __double2hiint(double x) { unsigned long long u; memcpy(&u, &x, 8); return (int)(unsigned)(u >> 32); }
inline double __hiloint2double(int hi, int lo) {
    unsigned long long u = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo; double x; memcpy(&x, &u, 8); return x;
}

// runtime shims
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu error"; }
