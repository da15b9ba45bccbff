// Micro-benchmark (B200): tcgen05.ld drain rate by shape / warp count, alone and against a concurrent tcgen05.mma stream.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Ipapr_b200/csrc tools/ubench_tmem.cu -o gpurun_out/ubench_tmem
// It answers one design question for stack.cu: is the 128 x 256 fp32 accumulator drain bound by TMEM read bandwidth, by
// the shape of the load, or by contention with the tensor pipe?
#include "tc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace papr;

enum { S32x32_32 = 0, S32x32_64, S32x32_16, S16x256_8, S16x256_16, S16x128_16, S16x64_32, S32x32_32_NOWAIT, S_STS128, S_LDS128 };

template <int SHAPE> __device__ __forceinline__ uint32_t tld(uint32_t taddr)
{
    // returns xor of a few registers so the load cannot be dropped; bytes moved per call: see bytes_per_call()
    uint32_t acc = 0;
    if constexpr (SHAPE == S_STS128) {          // taddr = shared address of this thread's 128-byte row; 8 x 16 B, swizzled like stack.cu
        const uint32_t sw = (taddr >> 7) & 7;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(taddr + ((c ^ sw) << 4)), "r"(taddr) : "memory");
        return 0;
    } else if constexpr (SHAPE == S_LDS128) {
        const uint32_t sw = (taddr >> 7) & 7;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t a, b, cc, d;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(cc), "=r"(d) : "r"(taddr + ((c ^ sw) << 4)) : "memory");
            acc ^= a ^ d;
        }
        return acc;
    } else
    if constexpr (SHAPE == S32x32_32 || SHAPE == S32x32_32_NOWAIT) {
        uint32_t v[32];
        tmem_ld32(taddr, v);
        if (SHAPE == S32x32_32) tmem_ld_wait();
        acc = v[0] ^ v[31];
    } else if constexpr (SHAPE == S32x32_16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        tmem_ld_wait();
        acc = v[0] ^ v[15];
    } else if constexpr (SHAPE == S32x32_64) {
        uint32_t v[64];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
            "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]),
              "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]),
              "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]),
              "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
            : "r"(taddr));
        tmem_ld_wait();
        acc = v[0] ^ v[63];
    } else if constexpr (SHAPE == S16x256_8 || SHAPE == S16x128_16 || SHAPE == S16x64_32) {
        uint32_t v[32];
#define LD32(SH)                                                                                                                   \
    asm volatile("tcgen05.ld.sync.aligned." SH ".b32 "                                                                              \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), \
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), \
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                              \
                 : "r"(taddr))
        if constexpr (SHAPE == S16x256_8) LD32("16x256b.x8");
        else if constexpr (SHAPE == S16x128_16) LD32("16x128b.x16");
        else LD32("16x64b.x32");
#undef LD32
        tmem_ld_wait();
        acc = v[0] ^ v[31];
    } else if constexpr (SHAPE == S16x256_16) {
        uint32_t v[64];
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x16.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
            "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]),
              "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]),
              "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]),
              "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
            : "r"(taddr));
        tmem_ld_wait();
        acc = v[0] ^ v[63];
    }
    return acc;
}

__host__ __device__ constexpr int bytes_per_call(int shape)
{
    return shape == S32x32_64 || shape == S16x256_16 ? 8192 : shape == S32x32_16 ? 2048 : 4096;
}

struct Result { long long ld_cycles, mma_cycles; unsigned sink; };

// warps 0..NW-1 drain TMEM columns [256, 512) repeatedly; warp NW issues `mma_batches` batches of 16 x (M128 N256 K16)
// into columns [0, 256) when with_mma; both are timed with clock64
template <int SHAPE, int NW> __global__ void __launch_bounds__((NW + 1) * 32, 1) bench(Result *out, int iters, int mma_batches, int with_ld)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (64 + 128) * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    long long t0 = 0, t1 = 0;
    unsigned sink = 0;
    if (warp < NW) {
        if (with_ld) {
            const int quad = warp & 3;
            const int grp = warp >> 2;                       // column group, as in stack.cu
            constexpr int cols = SHAPE == S32x32_64 ? 64 : SHAPE == S32x32_16 ? 16 : SHAPE == S16x256_16 ? 128 : SHAPE == S16x256_8 ? 64
                                 : SHAPE == S16x128_16 ? 64 : SHAPE == S16x64_32 ? 64 : 32;
            t0 = clock64();
            if constexpr (SHAPE == S_STS128 || SHAPE == S_LDS128) {
                // scratch rows beyond the MMA operands: 32 KB region at smem + 192 KB, one 128-byte row per thread (mod 256 rows)
                const uint32_t base = smem_u32(smem + 192 * 1024) + ((threadIdx.x & 255) << 7);
                for (int it = 0; it < iters; ++it) sink ^= tld<SHAPE>(base);
            } else
            for (int it = 0; it < iters; ++it) {
                const uint32_t col = 256 + ((grp * 64 + it * cols) & 255 & ~(cols - 1));
                uint32_t lane_off = (uint32_t)(quad * 32) << 16;
                if (SHAPE >= S16x256_8 && SHAPE <= S16x64_32 && (it & 1)) lane_off += 16u << 16;
                sink ^= tld<SHAPE>(tb + lane_off + (col > 512 - cols ? 512 - cols : col));
            }
            tmem_ld_wait();
            t1 = clock64();
        }
    } else if (lane == 0 && mma_batches > 0) {
        const uint32_t idesc = umma_idesc(128, 256, false, false);
        const uint64_t ad = umma_desc(smem_u32(smem), 16, 1024);
        const uint64_t bd = umma_desc(smem_u32(smem + 65536), 16, 1024);
        t0 = clock64();
        for (int b = 0; b < mma_batches; ++b) {
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tb, ad + kb * 1024 + 2 * k, bd + kb * 2048 + 2 * k, idesc, (kb | k) ? 1u : 0u);
            umma_commit(&bar[b & 1]);
            if (b > 0) mbar_wait(&bar[(b - 1) & 1], (uint32_t)(((b - 1) >> 1) & 1));      // one batch always queued behind
        }
        mbar_wait(&bar[(mma_batches - 1) & 1], (uint32_t)(((mma_batches - 1) >> 1) & 1));
        t1 = clock64();
    }
    if (lane == 0 && blockIdx.x == 0) {
        if (warp == 0) out->ld_cycles = t1 - t0;
        if (warp == NW) out->mma_cycles = t1 - t0;
    }
    if (sink == 0x12345678u && blockIdx.x == 0) out->sink = sink;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <int SHAPE, int NW> void run(const char *name, Result *d)
{
    const int iters = 2048, batches = 256;
    const int smem = (64 + 128 + 32) * 1024 + 1024;
    cudaFuncSetAttribute(bench<SHAPE, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    Result h[3];
    for (int mode = 0; mode < 3; ++mode) {      // 0: ld alone, 1: mma alone, 2: both
        cudaMemset(d, 0, sizeof(Result));
        bench<SHAPE, NW><<<148, (NW + 1) * 32, smem>>>(d, iters, mode == 0 ? 0 : batches, mode != 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s NW=%d mode %d: %s\n", name, NW, mode, cudaGetErrorString(e)); return; }
        cudaMemcpy(&h[mode], d, sizeof(Result), cudaMemcpyDeviceToHost);
    }
    const double bytes = (double)iters * bytes_per_call(SHAPE) * NW;
    printf("%-14s warps=%2d | ld alone %6.1f B/cyc/SM (128x256 fp32 drain = %5.0f cyc) | mma alone %5.0f cyc/batch | together: ld %6.1f B/cyc, "
           "mma %5.0f cyc/batch\n",
           name, NW, bytes / h[0].ld_cycles, 131072.0 / (bytes / h[0].ld_cycles), (double)h[1].mma_cycles / batches,
           bytes / h[2].ld_cycles, (double)h[2].mma_cycles / batches);
}


// ---- 2-CTA (cta_group::2) MMA stream, M = 256 across the pair, as stack.cu issues it; optional concurrent traffic:
// bit0 tcgen05.ld by 16 warps, bit1 st.shared by the same warps, bit2 TMA bulk loads (L2 -> smem scratch) in both CTAs
struct Result2 { long long mma_cycles, tma_cycles, ld_cycles; };
template <int CTAS> __global__ void __launch_bounds__(18 * 32, 1) bench2(Result2 *out, const uint8_t *gsrc, int mma_batches, int traffic, int iters, uint32_t getenv_ssz)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[2], tbar[2], done_bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0;
    for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        // random data: two bf16 in [-2, 2) with random mantissas (sign | exponent 0x3f/0x3e.. | mantissa)
        const uint32_t r = (h & 0x807f807fu) | 0x3f003f00u | ((h >> 3) & 0x00800080u);
        ((uint32_t *)smem)[i] = (traffic & 8) ? r : 0x3c003c00u;
    }
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&tbar[0], 1); mbar_init(&tbar[1], 1); mbar_init(&done_bar, 1); fence_barrier_init(); }
    if (warp == 0) { if (CTAS == 2) tmem_alloc2(&tmem_slot, 512); else tmem_alloc(&tmem_slot, 512); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (CTAS == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    long long t0 = 0, t1 = 0;
    unsigned sink = 0;
    if (warp < 16) {
        if (traffic & 16) mbar_wait(&done_bar, 0);          // all 512 threads poll one mbarrier, as stack.cu's epilogue warps do
        if (traffic & 3) {
            const int quad = warp & 3, grp = warp >> 2;
            const uint32_t srow = smem_u32(smem + 128 * 1024) + (uint32_t)(((grp & 1) * 128 + quad * 32 + lane) << 7);
            t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                if (traffic & 1) sink ^= tld<S32x32_32>(tb + ((uint32_t)(quad * 32) << 16) + 256 + grp * 64 + (it & 1) * 32);
                if (traffic & 2) {
                    const uint32_t sw = (srow >> 7) & 7;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(srow + (((it & 1) * 4 + c) ^ sw) * 16), "r"(sink) : "memory");
                }
            }
            t1 = clock64();
            if (warp == 0 && lane == 0 && blockIdx.x == 0) out->ld_cycles = t1 - t0;
        }
    } else if (warp == 16) {
        if (lane == 0 && rank == 0 && mma_batches > 0) {
            const uint32_t idesc = umma_idesc(CTAS == 2 ? 256 : 128, 256, false, false);
            const uint64_t ad = umma_desc(smem_u32(smem), 16, 1024);
            const uint64_t bd = umma_desc(smem_u32(smem + 65536), 16, 1024);
            const uint32_t bstep = CTAS == 2 ? 1024 : 2048;      // 16-byte units per 64-wide K block of B (N/2 or N rows)
            t0 = clock64();
            for (int b = 0; b < mma_batches; ++b) {
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (CTAS == 2) umma2_bf16(tb, ad + kb * 1024 + 2 * k, bd + kb * bstep + 2 * k, idesc, (kb | k) ? 1u : 0u);
                        else umma_bf16(tb, ad + kb * 1024 + 2 * k, bd + kb * bstep + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    }
                if (CTAS == 2) umma2_commit(&bar[b & 1]); else umma_commit(&bar[b & 1]);
                if (b > 0) mbar_wait(&bar[(b - 1) & 1], (uint32_t)(((b - 1) >> 1) & 1));
            }
            mbar_wait(&bar[(mma_batches - 1) & 1], (uint32_t)(((mma_batches - 1) >> 1) & 1));
            t1 = clock64();
            if (blockIdx.x == 0) out->mma_cycles = t1 - t0;
        }
        if (lane == 0) mbar_arrive(&done_bar);
    } else if (warp == 17) {
        if (lane == 0 && (traffic & 32)) {
            uint8_t *srcs = smem + (CTAS == 2 ? 128 : 192) * 1024;
            uint8_t *dstg = const_cast<uint8_t *>(gsrc) + (size_t)(blockIdx.x % 16) * (1 << 20);
            t0 = clock64();
            const uint32_t ssz = getenv_ssz;
            for (int it = 0; it < iters / 4; ++it) {
                for (uint32_t o = 0; o < 16384; o += ssz) bulk_s2g(dstg + (size_t)(it & 31) * 16384 + o, srcs + (it & 1) * 16384 + o, ssz);
                bulk_commit();
                bulk_wait_read<2>();
            }
            bulk_wait<0>();
            t1 = clock64();
            if (blockIdx.x == 0) out->tma_cycles = t1 - t0;
        }
        if (lane == 0 && (traffic & 4)) {
            // two 32 KB loads in flight into the scratch region (CTAS==1 shares it with B's upper half only when N rows = 256:
            // scratch sits at 128 KB for the pair kernel and at 192 KB - 64 KB = 128 KB .. see host smem size)
            uint8_t *dst = smem + (CTAS == 2 ? 128 : 192) * 1024;
            const uint8_t *src = gsrc + (size_t)(blockIdx.x % 16) * (1 << 20);
            t0 = clock64();
            for (int it = 0; it < iters / 4; ++it) {
                const int sb = it & 1;
                if (it >= 2) mbar_wait(&tbar[sb], (uint32_t)(((it - 2) >> 1) & 1));
                constexpr uint32_t CH = CTAS == 2 ? 32768 : 16384;
                mbar_arrive_expect_tx(&tbar[sb], CH);
                bulk_g2s(dst + sb * CH, src + (size_t)(it & 15) * CH, CH / 2, &tbar[sb]);
                bulk_g2s(dst + sb * CH + CH / 2, src + (size_t)(it & 15) * CH + CH / 2, CH / 2, &tbar[sb]);
            }
            const int n = iters / 4;
            if (n >= 2) mbar_wait(&tbar[(n - 2) & 1], (uint32_t)(((n - 2) >> 1) & 1));
            if (n >= 1) mbar_wait(&tbar[(n - 1) & 1], (uint32_t)(((n - 1) >> 1) & 1));
            t1 = clock64();
            if (blockIdx.x == 0) out->tma_cycles = t1 - t0;
        }
    }
        if (warp == 0 && lane == 0 && (traffic & 64)) {
            // two 32 KB loads in flight into the scratch region (CTAS==1 shares it with B's upper half only when N rows = 256:
            // scratch sits at 128 KB for the pair kernel and at 192 KB - 64 KB = 128 KB .. see host smem size)
            uint8_t *dst = smem + (CTAS == 2 ? 128 : 192) * 1024;
            const uint8_t *src = gsrc + (size_t)(blockIdx.x % 16) * (1 << 20);
            t0 = clock64();
            for (int it = 0; it < iters / 4; ++it) {
                const int sb = it & 1;
                if (it >= 2) mbar_wait(&tbar[sb], (uint32_t)(((it - 2) >> 1) & 1));
                constexpr uint32_t CH = CTAS == 2 ? 32768 : 16384;
                mbar_arrive_expect_tx(&tbar[sb], CH);
                bulk_g2s(dst + sb * CH, src + (size_t)(it & 15) * CH, CH / 2, &tbar[sb]);
                bulk_g2s(dst + sb * CH + CH / 2, src + (size_t)(it & 15) * CH + CH / 2, CH / 2, &tbar[sb]);
            }
            const int n = iters / 4;
            if (n >= 2) mbar_wait(&tbar[(n - 2) & 1], (uint32_t)(((n - 2) >> 1) & 1));
            if (n >= 1) mbar_wait(&tbar[(n - 1) & 1], (uint32_t)(((n - 1) >> 1) & 1));
            t1 = clock64();
            if (blockIdx.x == 0) out->ld_cycles = t1 - t0;
        }
    
    if (sink == 0x12345678u && blockIdx.x == 0) out->ld_cycles = sink;
    tc_fence_before();
    __syncthreads();
    if (CTAS == 2) cluster_sync_all();
    if (warp == 0) { if (CTAS == 2) tmem_dealloc2(tb, 512); else tmem_dealloc(tb, 512); }
}

template <int CTAS> void run2(Result2 *d, const uint8_t *gsrc)
{
    const int smem = (CTAS == 2 ? 192 : 224) * 1024 + 1024 + 1024;     // pair: A 64 + B/2 64 + scratch 64; single: A 64 + B 128 + scratch 32(+)
    cudaFuncSetAttribute(bench2<CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int batches = 256, iters = 2048;
    for (int traffic : {4, 32, 64, 96}) {
        cudaMemset(d, 0, sizeof(Result2));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(getenv("GRID") ? atoi(getenv("GRID")) : 148); cfg.blockDim = dim3(18 * 32); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, bench2<CTAS>, d, gsrc, batches, traffic, iters, (uint32_t)(getenv("SSZ") ? atoi(getenv("SSZ")) : 16384));
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("bench2<%d> traffic %d: %s\n", CTAS, traffic, cudaGetErrorString(e)); return; }
        Result2 h;
        cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("cta_group::%d M=%d N=256 K=256 batch | traffic: %s%s%s%s| mma %5.0f cyc/batch", CTAS, 128 * CTAS, traffic & 1 ? "tmem-ld " : "",
               traffic & 2 ? "sts " : "", traffic & 4 ? "tma-load " : (traffic & 32 ? "tma-STORE " : ""), traffic & 16 ? "512-thread-mbarrier-poll " : (traffic & 8 ? "RANDOM-DATA " : "const-data "), (double)h.mma_cycles / batches);
        if (traffic & 64) printf(" | tma-load(warp0) %5.1f B/cyc/SM", (double)(iters / 4) * (CTAS == 2 ? 32768 : 16384) / h.ld_cycles);
        if (traffic & 4) printf(" | tma %5.1f B/cyc/SM", (double)(iters / 4) * (CTAS == 2 ? 32768 : 16384) / h.tma_cycles);
        else if (traffic & 32) printf(" | tma-store %5.1f B/cyc/SM", (double)(iters / 4) * 16384 / h.tma_cycles);
        if (traffic & 3) printf(" | ld/sts loop %5.0f cyc per 128x256 tile", (double)h.ld_cycles / iters * 2);
        printf("\n");
    }
}

int main()
{
    { Result2 *d2; uint8_t *g; cudaMalloc(&d2, sizeof(Result2)); cudaMalloc(&g, 16 << 20); cudaMemset(g, 0x3c, 16 << 20);
      run2<2>(d2, g); run2<1>(d2, g); if (getenv("ONLY2")) return 0; }
    Result *d;
    cudaMalloc(&d, sizeof(Result));
#define RUN(S, NW) run<S, NW>(#S, d)
    RUN(S32x32_32, 4); RUN(S32x32_32, 8); RUN(S32x32_32, 16);
    RUN(S32x32_32_NOWAIT, 4); RUN(S32x32_32_NOWAIT, 8); RUN(S32x32_32_NOWAIT, 16);
    RUN(S32x32_16, 4); RUN(S32x32_16, 16);
    RUN(S32x32_64, 4); RUN(S32x32_64, 8);
    RUN(S16x256_8, 4); RUN(S16x256_8, 8); RUN(S16x256_8, 16);
    RUN(S16x256_16, 4); RUN(S16x256_16, 8);
    RUN(S16x128_16, 4); RUN(S16x128_16, 16);
    RUN(S16x64_32, 4); RUN(S16x64_32, 16);
    RUN(S_STS128, 4); RUN(S_STS128, 8); RUN(S_STS128, 16);
    RUN(S_LDS128, 4); RUN(S_LDS128, 8); RUN(S_LDS128, 16);
    return 0;
}
