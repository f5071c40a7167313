// stage_probe.cu — dev tool: what the asynchronous copy engines sustain per SM on the column-tile access pattern
// (rows of 64 / 128 / 32 bytes, 32 KiB apart) of a [2][2048][2048] complex128 state.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o stage_probe stage_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct alignas(64) Map { CUtensorMap m; };
__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(n) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned ok = 0;
    while (!ok)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(su32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load(void* dst, const Map* m, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(su32(dst)), "l"((unsigned long long)m), "r"(su32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store(const Map* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"((unsigned long long)m), "r"(su32(src)), "r"(c0), "r"(c1) : "memory");
}

// mode 0: TMA loads only.  mode 1: TMA load + TMA store of the same tile (in-place copy).  mode 2: TMA load, plain
// 16-byte register stores from shared memory by the threads.  STAGES half-/full-tiles in flight.
template <int STAGES>
__global__ void __launch_bounds__(512, 1) tma_tiles(const __grid_constant__ Map map, int ntiles, int tiles_per_plane,
                                                    int rows, int row_bytes, int box_rows, int mode, double2* out, int nx) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar[STAGES];
    const int tile_bytes = rows * row_bytes;
    const int nbox = rows / box_rows;
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES; s++) mbar_init(&bar[s], 1);
    __syncthreads();
    auto issue = [&](int tile, int s) {
        mbar_expect(&bar[s], tile_bytes);
        const int plane = tile / tiles_per_plane, t = tile % tiles_per_plane;
        for (int q = 0; q < nbox; q++)
            tma_load(smem + (size_t)s * tile_bytes + (size_t)q * box_rows * row_bytes, &map, t * (row_bytes / 8),
                     plane * rows + q * box_rows, &bar[s]);
    };
    int issued = 0;
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES; s++) {
            const int tile = blockIdx.x + s * gridDim.x;
            if (tile < ntiles) issue(tile, s);
        }
    unsigned phase[STAGES] = {0};
    int k = 0;
    double sink = 0.0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, k++) {
        const int s = k % STAGES;
        mbar_wait(&bar[s], phase[s]);
        phase[s] ^= 1u;
        const unsigned char* src = smem + (size_t)s * tile_bytes;
        if (mode == 1) {
            if (threadIdx.x == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const int plane = tile / tiles_per_plane, t = tile % tiles_per_plane;
                for (int q = 0; q < nbox; q++)
                    tma_store(&map, src + (size_t)q * box_rows * row_bytes, t * (row_bytes / 8), plane * rows + q * box_rows);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncthreads();
        } else if (mode == 2) {
            const int plane = tile / tiles_per_plane, t = tile % tiles_per_plane;
            const int per_row = row_bytes / 16;
            const double2* sv = reinterpret_cast<const double2*>(src);
            for (int e = threadIdx.x; e < rows * per_row; e += blockDim.x) {
                const int r = e / per_row, c = e % per_row;
                __stcs(&out[((long long)plane * rows + r) * nx + t * per_row + c], sv[e]);
            }
            __syncthreads();
        } else {
            sink += (double)src[threadIdx.x];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const int nt = tile + STAGES * gridDim.x;
            if (nt < ntiles) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(nt, s); }
        }
        (void)issued;
    }
    if (sink == 12345.678) out[0].x = sink;
}

// LDGSTS variant: 512 threads, 16 bytes each, 16 per thread (tile of 2048 rows x 64 bytes), two half-tiles in flight
__global__ void __launch_bounds__(512, 1) ldgsts_tiles(const double2* in, int ntiles, int tiles_per_plane, int nx, double2* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    double2* S = reinterpret_cast<double2*>(smem);
    const int c = threadIdx.x % 4, j = threadIdx.x / 4;
    double sink = 0.0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int plane = tile / tiles_per_plane, t = tile % tiles_per_plane;
        const double2* base = in + (long long)plane * 2048 * nx + t * 4 + c;
        for (int m = 0; m < 16; m++) {
            const int r = j + m * 128;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(su32(&S[r * 4 + c])), "l"(base + (long long)r * nx) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        sink += S[threadIdx.x].x;
        __syncthreads();
    }
    if (sink == 12345.678) out[0].x = sink;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int nx = 2048, ny = 2048, planes = 2;
    const size_t bytes = (size_t)planes * ny * nx * 16;
    double2 *a, *b;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 0, bytes)); CK(cudaMemset(b, 0, bytes));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn encode = (EncodeFn)fn;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    struct Cfg { int row_bytes, rows, box_rows, stages, mode; const char* name; };
    const Cfg cfgs[] = {
        {64, 2048, 256, 1, 0, "TMA load  64 B rows, 128 KiB tile, 1 in flight"},
        {64, 1024, 256, 2, 0, "TMA load  64 B rows,  64 KiB half tiles, 2 in flight"},
        {64, 512, 256, 4, 0, "TMA load  64 B rows,  32 KiB quarter tiles, 4 in flight"},
        {128, 1024, 256, 1, 0, "TMA load 128 B rows, 128 KiB tile, 1 in flight"},
        {128, 512, 256, 2, 0, "TMA load 128 B rows,  64 KiB half tiles, 2 in flight"},
        {32, 2048, 256, 2, 0, "TMA load  32 B rows,  64 KiB tiles, 2 in flight"},
        {64, 1024, 256, 2, 1, "TMA load + TMA store 64 B rows, half tiles, 2 in flight"},
        {64, 1024, 256, 2, 2, "TMA load + register stores 64 B rows, half tiles, 2 in flight"},
        {64, 2048, 64, 1, 0, "TMA load  64 B rows, boxes of 64 rows, 1 in flight"},
    };
    for (const Cfg& c : cfgs) {
        Map map;
        const cuuint64_t dims[2] = {2ull * nx, (cuuint64_t)planes * ny};
        const cuuint64_t strides[1] = {(cuuint64_t)nx * 16};
        const cuuint32_t box[2] = {(cuuint32_t)(c.row_bytes / 8), (cuuint32_t)c.box_rows};
        const cuuint32_t es[2] = {1, 1};
        CUresult rc = encode(&map.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, a, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); return 1; }
        // tiles: the array is cut into (ny / rows) row blocks x (nx * 16 / row_bytes) column tiles per plane... keep it simple:
        // a "plane" for the kernel = one block of `rows` rows
        const int tiles_per_plane = nx * 16 / c.row_bytes;
        const int nplanes = planes * ny / c.rows;
        const int ntiles = tiles_per_plane * nplanes;
        const size_t smem = (size_t)c.stages * c.rows * c.row_bytes;
        auto run = [&](auto kern) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            for (int rep = 0; rep < 3; rep++) {
                if (rep == 2) cudaEventRecord(e0);
                kern<<<sms, 512, smem>>>(map, ntiles, tiles_per_plane, c.rows, c.row_bytes, c.box_rows, c.mode, b, nx);
            }
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("%-66s %8.1f us  %7.1f GB/s read\n", c.name, ms * 1e3, bytes / (ms * 1e-3) / 1e9);
        };
        if (c.stages == 1) run(tma_tiles<1>);
        else if (c.stages == 2) run(tma_tiles<2>);
        else run(tma_tiles<4>);
    }
    {
        const size_t smem = 2048 * 64;
        CK(cudaFuncSetAttribute(ldgsts_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int rep = 0; rep < 3; rep++) {
            if (rep == 2) cudaEventRecord(e0);
            ldgsts_tiles<<<sms, 512, smem>>>(a, 2 * nx / 4, nx / 4, nx, b);
        }
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-66s %8.1f us  %7.1f GB/s read\n", "LDGSTS 16 B per thread, 64 B rows, 128 KiB tile, 1 in flight", ms * 1e3, bytes / (ms * 1e-3) / 1e9);
    }
    return 0;
}
