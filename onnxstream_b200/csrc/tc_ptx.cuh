// tc_ptx.cuh -- inline-PTX wrappers shared by the tcgen05 kernels (gemm_tcgen05.cu: one CTA per tile; gemm2_tcgen05.cu: CTA pairs).
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cstdio>

namespace tcptx {

// ---- PTX wrappers ---------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    long long t0 = 0;
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        // watchdog: a protocol bug must surface as a launch failure, never as a hung GPU (~2 s at 2 GHz)
        long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000LL) { printf("tcgen05 kernel: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}

__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// The producer / MMA warps run WARP-UNIFORM control flow and issue from an elect.sync-guarded region.  (Inside an
// `if (lane == 0)` region ptxas cannot keep the operands of UTCHMMA / UTMALDG in uniform registers and wraps every one of them
// in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop: ~1000 issue cycles per k-block, measured with ncu source sampling.)
__device__ __forceinline__ void tma_load_3d_s(uint32_t smem_addr, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_addr), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// one lane of a converged warp, chosen by the hardware (elect.sync): ptxas knows the guarded region runs on exactly one lane
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], single CTA, fp16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued MMAs have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (see cute/arch/mma_sm100_desc.hpp SmemDescriptor): SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;   // LayoutType::SWIZZLE_128B
    return d;
}


// ---- GroupNorm statistics of a GEMM / conv output, gathered in the producing kernel's epilogue -----------------------------------
// The consumer GroupNorm needs sum(x) and sum(x^2) per group of `cpg` consecutive channels over ALL rows (pixels).  In the epilogue a
// lane owns one ROW of the tile (TMEM lane = row), so per-column sums need a transpose: the warp parks its 32 x 32 chunk of fp16-rounded
// outputs in a private 2 KiB shared buffer (row pitch 64 B, 16-byte units XOR-swizzled by (row >> 1) & 3: conflict-free writes), then lane
// l walks column l down the 32 rows (one wavefront per row), and adds its column's (sum, sum of squares) to the CTA's per-group
// shared accumulators.  After the tile the accumulators are flushed to the global fp64 statistics (the reference accumulates in double,
// src/onnxstream.cpp:4788-5055) -- fp32 only ever sums <= 128 rows x cpg values.
constexpr int GN_MAX_GROUPS = 64;

// h: this lane's row, 32 fp16 values already rounded as they are stored (rows outside the problem must be passed as zeros)
__device__ __forceinline__ void gn_stats_chunk(uint32_t wbuf, const uint32_t (&h)[16], int n_base, int n_end, int cpg, float* cta_stats, int lane)
{
    const uint32_t row = wbuf + (uint32_t)lane * 64u;
    const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
    for (uint32_t u = 0; u < 4; u++)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((u ^ sw) << 4)), "r"(h[4 * u]), "r"(h[4 * u + 1]), "r"(h[4 * u + 2]), "r"(h[4 * u + 3]) : "memory");
    __syncwarp();
    float s = 0.f, ss = 0.f;
    const uint32_t unit = (uint32_t)lane >> 3, within = ((uint32_t)lane & 7u) * 2u;
#pragma unroll
    for (uint32_t r = 0; r < 32; r++) {
        unsigned short raw;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(raw) : "r"(wbuf + r * 64u + ((unit ^ ((r >> 1) & 3u)) << 4) + within));
        const float v = __half2float(__ushort_as_half(raw));
        s += v; ss = fmaf(v, v, ss);
    }
    // lane l holds column n_base + l; the lanes of one group are contiguous: segmented suffix sums by shuffle (5 steps) leave each
    // group's total in its first lane, which alone touches the shared accumulator -- 32 colliding float atomics per chunk measured
    // +10 us per launch (profiles/r02_ab_epilogue.txt)
    const int n = n_base + lane;
    const int g = n / cpg;
    if (n >= n_end) { s = 0.f; ss = 0.f; }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float ts = __shfl_down_sync(0xffffffffu, s, off), tss = __shfl_down_sync(0xffffffffu, ss, off);
        if (lane + off < 32 && (n + off) / cpg == g) { s += ts; ss += tss; }
    }
    const bool head = lane == 0 || (n - 1) / cpg != g;
    if (head && n < n_end) {
        atomicAdd(&cta_stats[2 * g], s);
        atomicAdd(&cta_stats[2 * g + 1], ss);
    }
    __syncwarp();
}

// after a tile: thread t of the epilogue group moves accumulator t to the global fp64 statistics and re-arms it
__device__ __forceinline__ void gn_stats_flush(float* cta_stats, double* gstats, int groups, int t)
{
    if (t < 2 * groups) {
        const float v = atomicExch(&cta_stats[t], 0.f);
        if (v != 0.f) atomicAdd(&gstats[t], (double)v);
    }
}

}  // namespace tcptx
