// Micro-benchmark (B200): how fast can ONE SM push the backward stash out, and what does that do to concurrent TMA loads?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Ipapr_b200/csrc tools/ubench_store.cu -o gpurun_out/ubench_store
// The stack kernel's training mode writes 64 KB per job per CTA with TMA bulk stores (31 B/cycle/SM, and weight loads slow
// down behind them: profiles/r01_stack_kernel_study.md).  This measures the alternative: coalesced st.global.v4 from the 16
// epilogue warps (512 contiguous bytes per warp instruction, the layout of a no-swizzle MN-major operand), alone and
// against the weight-load stream.  Destination either a small per-CTA ring (stays in L2) or a large buffer (HBM).
//   mode bit0: TMA bulk stores (one thread, `depth` 16 KB pieces in flight)      bit1: st.global.v4 from 16 warps
//        bit2: TMA bulk loads, two 32 KB in flight, from a 2 MB L2-resident source
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"

using namespace papr;

struct Result { long long store_cycles, load_cycles; long long store_bytes, load_bytes; };

__global__ void __launch_bounds__(640, 1) bench(Result *res, uint8_t *dst, size_t dst_per_cta, const uint8_t *src, int iters, int mode, int depth)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *tile = smem;                      // 64 KB "activation tile"
    uint8_t *scratch = smem + 65536;           // 64 KB load landing zone
    uint64_t *bar = (uint64_t *)(smem + 131072);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) ((uint32_t *)tile)[i] = i * 2654435761u;
    fence_proxy_async();
    __syncthreads();
    uint8_t *my = dst + (size_t)blockIdx.x * dst_per_cta;
    const size_t n_slots = dst_per_cta / 65536;
    if (warp == 0 && lane == 0 && (mode & 4)) {
        const long long t0 = clock64();
        uint32_t ph[2] = {0, 0};
        for (int i = 0; i < 2; ++i) { mbar_arrive_expect_tx(&bar[i], 32768); bulk_g2s(scratch + i * 32768, src + ((size_t)(blockIdx.x * 7 + i) % 64) * 32768, 32768, &bar[i]); }
        for (int it = 2; it < 2 * iters; ++it) {
            const int b = it & 1;
            mbar_wait(&bar[b], ph[b]); ph[b] ^= 1;
            mbar_arrive_expect_tx(&bar[b], 32768);
            bulk_g2s(scratch + b * 32768, src + ((size_t)(blockIdx.x * 7 + it) % 64) * 32768, 32768, &bar[b]);
        }
        mbar_wait(&bar[0], ph[0]); mbar_wait(&bar[1], ph[1]);
        res[blockIdx.x].load_cycles = clock64() - t0;
        res[blockIdx.x].load_bytes = (long long)2 * iters * 32768;
    } else if (warp == 2 && lane == 0 && (mode & 1)) {
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            uint8_t *d = my + (size_t)(it % n_slots) * 65536;
            for (int g = 0; g < 4; ++g) {
                bulk_s2g(d + g * 16384, tile + g * 16384, 16384);
                bulk_commit();
                if (depth <= 1) bulk_wait_read<0>(); else if (depth == 2) bulk_wait_read<1>(); else if (depth == 3) bulk_wait_read<2>(); else bulk_wait_read<3>();
            }
        }
        bulk_wait<0>();
        res[blockIdx.x].store_cycles = clock64() - t0;
        res[blockIdx.x].store_bytes = (long long)iters * 65536;
    } else if (warp >= 4 && (mode & 2)) {
        // warp (g, quad): 64-column group g, rows quad*32 + lane -- as the epilogue; chunk-major destination:
        // slot = [g][chunk 0..7][row 0..127][16 B]  -> a warp instruction writes 32 rows x 16 B = 512 contiguous bytes
        const int ew = warp - 4, g = ew >> 2, quad = ew & 3, row = quad * 32 + lane;
        uint4 v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = *(const uint4 *)(tile + g * 16384 + row * 128 + c * 16);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            uint8_t *d = my + (size_t)(it % n_slots) * 65536 + g * 16384 + row * 16;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                v[c].x += it;
                asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(d + c * 2048), "r"(v[c].x), "r"(v[c].y), "r"(v[c].z), "r"(v[c].w) : "memory");
            }
            if (depth >= 100) {      // pace like the epilogue: ~1,750 cycles of other work per job would go here
                const long long t1 = clock64();
                while (clock64() - t1 < depth) { }
            }
        }
        __threadfence();
        if (ew == 0 && lane == 0) {
            res[blockIdx.x].store_cycles = clock64() - t0;
            res[blockIdx.x].store_bytes = (long long)iters * 65536;
        }
    }
}

int main(int argc, char **argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    const int smem = 131072 + 1024 + 1024;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    Result *d; cudaMalloc(&d, 148 * sizeof(Result));
    uint8_t *src; cudaMalloc(&src, 64 * 32768); cudaMemset(src, 1, 64 * 32768);
    const size_t big = (size_t)148 * 2000 * 65536;       // 19.4 GB: every iteration a fresh slot -> HBM writes
    uint8_t *dst; if (cudaMalloc(&dst, big) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    struct { const char *name; int mode, depth; size_t per_cta; } cases[] = {
        {"TMA stores, L2 ring (8 slots), depth 3", 1, 3, 8 * 65536},
        {"TMA stores, HBM, depth 3", 1, 3, (size_t)2000 * 65536},
        {"st.global.v4 x16 warps, L2 ring", 2, 0, 8 * 65536},
        {"st.global.v4 x16 warps, HBM", 2, 0, (size_t)2000 * 65536},
        {"st.global.v4 paced 1500 cyc/job, HBM", 2, 1500, (size_t)2000 * 65536},
        {"TMA loads alone", 4, 0, 8 * 65536},
        {"TMA stores + loads, L2 ring", 5, 3, 8 * 65536},
        {"TMA stores + loads, HBM", 5, 3, (size_t)2000 * 65536},
        {"st.global + loads, L2 ring", 6, 0, 8 * 65536},
        {"st.global + loads, HBM", 6, 0, (size_t)2000 * 65536},
        {"st.global paced 1500 + loads, HBM", 6, 1500, (size_t)2000 * 65536},
        {"st.global paced 2000 + loads, HBM", 6, 2000, (size_t)2000 * 65536},
    };
    for (auto &c : cases) {
        cudaMemset(d, 0, 148 * sizeof(Result));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        bench<<<148, 640, smem>>>(d, dst, c.per_cta, src, 50, c.mode, c.depth);      // warm-up
        cudaEventRecord(e0);
        bench<<<148, 640, smem>>>(d, dst, c.per_cta, src, iters, c.mode, c.depth);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        Result h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double sb = 0, sc = 0, lb = 0, lc = 0;
        for (int i = 0; i < 148; ++i) { sb += h[i].store_bytes; sc += h[i].store_cycles; lb += h[i].load_bytes; lc += h[i].load_cycles; }
        printf("%-42s %8.3f ms | stores %6.1f B/cyc/SM (%7.1f GB/s chip) | loads %6.1f B/cyc/SM\n", c.name, ms,
               sc > 0 ? sb / sc : 0.0, sb / ms / 1e6, lc > 0 ? lb / lc : 0.0);
    }
    return 0;
}
