// gemm_tcgen05.cu -- the tensor-core path for fp16 MatMul / Gemm / Conv / attention GEMMs on sm_100a.
//
// One persistent, warp-specialised kernel:
//   warp 0      TMA producer: cp.async.bulk.tensor tiles of A and B into 128B-swizzled shared memory, mbarrier-tracked
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (kind::f16, M=128, N=128, K=16) into TMEM
//   warps 2..5  epilogue: tcgen05.ld the fp32 accumulator, add bias / residual, round to fp16, store
// The accumulator is double-buffered in TMEM (2 x 128 columns) so the epilogue of tile i overlaps the main loop of
// tile i+1; shared memory holds a STAGES-deep ring of (A,B) k-blocks.
//
// A is always K-major ([M,K] activations).  B is either MN-major ([K,N] row-major: ONNX MatMul weights, the
// pre-transposed K of the attention pattern, V) or K-major ([N,K] row-major: OHWI conv weights).  A convolution is the
// same kernel with the A tiles fetched as 3-D boxes of the NHWC input -- one box per filter tap and 64-channel block,
// out-of-bounds (padding) elements zero-filled by TMA -- i.e. an implicit GEMM with no im2col buffer.
//
// Replaces: XnnPack::matrix_multiply / matrix_multiply_dynamic / convolution for T = uint16_t
// (src/onnxstream.cpp:929-1215, 1292-1534) and the cuBLAS offload CublasOps::OpFullyConnected::run (src/onnxstream.cpp:308-352).

#include "common.cuh"
#include "workspace.h"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 128;
constexpr int BLOCK_K = 64;            // 64 fp16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES_DEEP = 6;          // long K loops: one CTA per SM, deep TMA ring
constexpr int STAGES_SHORT = 3;         // short K loops: half the shared memory so two CTAs share an SM and hide each other's prologue / epilogue
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;   // 256
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;  // 16 KiB
constexpr int smem_bytes_for(int stages) { return stages * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/; }
constexpr int NUM_THREADS = 192;       // 6 warps

struct TcParams {
    int M, N, K;                 // GEMM view of the problem (conv: M = Ho*Wo, K = Cin per tap)
    int batch;
    int m_tiles, n_tiles;
    int bn;                      // N extent of a tile (64 / 80 / 96 / 128): chosen per problem to fill the 148 SMs
    int b_kmajor;                // 1: B is [N,K] row-major
    int a_swap, b_swap;          // tensor map has (batch, row) order swapped because the batch stride is the smaller one (per-head views)
    // conv geometry (taps == 1 for a plain GEMM)
    int taps, kw, pad_top, pad_left, Wo, Ho, bw, bh, tiles_x;
    int k_blocks_per_tap;
    int stride;                  // conv stride (TMA traversal stride on W and H)
    int short_k;                 // host hint: few k-blocks per CTA -> 3-stage ring, two CTAs per SM
    int split_k;                 // > 1: each tile's k-blocks are divided among split_k CTAs, fp32 partials go to `ws`
    float* ws;                   // split-K workspace [split][batch][M][N] fp32
    int* counters;               // split-K arrival counters, one per output tile (self-resetting)
    // grouped launch (groups > 1): `batch` problems share A and the shape; problem g has its own B map (map_b, map_b1, map_b2) and output
    int groups;
    __half* C1;
    __half* C2;
    // output
    __half* C;
    const __half* bias;
    const __half* residual;
    long long stride_c;          // elements between batches
    long long ldc;               // elements between output rows (== N for a dense C)
};

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
        else if (now - t0 > 4000000000LL) { printf("tc_gemm_kernel: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
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

// instruction descriptor for kind::f16: fp16 x fp16 -> fp32, A K-major, B K- or MN-major
__device__ __forceinline__ uint32_t make_idesc(int b_mn_major, int bn)
{
    uint32_t d = 0;
    d |= 1u << 4;                              // c_format = F32
    d |= 0u << 7;                              // a_format = F16
    d |= 0u << 10;                             // b_format = F16
    d |= 0u << 15;                             // a_major  = K
    d |= (uint32_t)(b_mn_major ? 1 : 0) << 16; // b_major
    d |= (uint32_t)(bn >> 3) << 17;            // n_dim
    d |= (uint32_t)(BLOCK_M >> 4) << 24;       // m_dim
    return d;
}

// ---- the kernel -----------------------------------------------------------------------------------------------

template <int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, STAGES <= STAGES_SHORT ? 2 : 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_b1,
               const __grid_constant__ CUtensorMap map_b2, const TcParams p)
{
    osb_pdl_trigger_entry();   // let the next kernel's CTAs be scheduled as ours drain; it waits for our completion before touching memory
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
    uint64_t* bars = (uint64_t*)(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
    uint64_t* full = bars;                       // [STAGES]
    uint64_t* empty = bars + STAGES;             // [STAGES]
    uint64_t* acc_full = bars + 2 * STAGES;      // [ACC_STAGES]
    uint64_t* acc_empty = acc_full + ACC_STAGES; // [ACC_STAGES]
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + ACC_STAGES);
    volatile int* split_flag = (volatile int*)(tmem_slot + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < ACC_STAGES; i++) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4 * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    osb_pdl_wait();      // everything above (barrier init, TMEM alloc, descriptor prefetch) overlapped the previous kernel's tail

    const int tiles_per_batch = p.m_tiles * p.n_tiles;
    const int total_tiles = tiles_per_batch * p.batch * p.split_k;
    const int k_blocks_all = p.taps * p.k_blocks_per_tap;
    const int kb_per_split = (k_blocks_all + p.split_k - 1) / p.split_k;   // host guarantees (split_k - 1) * kb_per_split < k_blocks_all

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform control flow, leader-predicated issue) =====================
        {
            const uint32_t sa0 = smem_u32(smem_a), sb0 = smem_u32(smem_b);
            const uint32_t tx_bytes = A_STAGE_BYTES + (p.b_kmajor ? p.bn * (BLOCK_K * 2) : ((p.bn + 63) / 64) * (B_STAGE_BYTES / 2));
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int sp = tile % p.split_k, t2 = tile / p.split_k;
                int b = t2 / tiles_per_batch, r = t2 % tiles_per_batch;
                int mt = r % p.m_tiles, nt = r / p.m_tiles;
                int n0 = nt * p.bn;
                int y0 = 0, x0 = 0, m0 = mt * BLOCK_M;
                if (p.taps > 1 || p.bh > 0) { y0 = (mt / p.tiles_x) * p.bh; x0 = (mt % p.tiles_x) * p.bw; }
                // grouped launch: the batch index selects the B map; every operand is addressed at batch coordinate 0
                const CUtensorMap* mbp = &map_b;
                if (p.groups > 1) { if (b == 1) mbp = &map_b1; else if (b == 2) mbp = &map_b2; b = 0; }
                int kb_lo = sp * kb_per_split, kb_hi = min(kb_lo + kb_per_split, k_blocks_all);
                // (tap, channel block) walk incrementally: no divisions inside the k loop
                int tap = kb_lo / p.k_blocks_per_tap, kcb = kb_lo % p.k_blocks_per_tap;
                int ky = tap / p.kw, kx = tap % p.kw;
                const int ax = x0 * p.stride - p.pad_left, ay = y0 * p.stride - p.pad_top;
                for (int kb = kb_lo; kb < kb_hi; kb++) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&full[stage], tx_bytes);
                        const int kc = kcb * BLOCK_K;
                        const uint32_t sa = sa0 + stage * A_STAGE_BYTES;
                        const uint32_t sb = sb0 + stage * B_STAGE_BYTES;
                        if (p.bh > 0) tma_load_3d_s(sa, &map_a, &full[stage], kc, ax + kx, ay + ky);
                        else if (p.a_swap) tma_load_3d_s(sa, &map_a, &full[stage], kc, b, m0);
                        else tma_load_3d_s(sa, &map_a, &full[stage], kc, m0, b);
                        const int kglob = tap * p.K + kc;   // K index into B (conv: taps are concatenated along K)
                        if (p.b_kmajor) {
                            if (p.b_swap) tma_load_3d_s(sb, mbp, &full[stage], kglob, b, n0);
                            else tma_load_3d_s(sb, mbp, &full[stage], kglob, n0, b);
                        } else {
                            if (p.b_swap) {
                                tma_load_3d_s(sb, mbp, &full[stage], n0, b, kglob);
                                if (p.bn > 64) tma_load_3d_s(sb + B_STAGE_BYTES / 2, mbp, &full[stage], n0 + 64, b, kglob);
                            } else {
                                tma_load_3d_s(sb, mbp, &full[stage], n0, kglob, b);
                                if (p.bn > 64) tma_load_3d_s(sb + B_STAGE_BYTES / 2, mbp, &full[stage], n0 + 64, kglob, b);
                            }
                        }
                    }
                    __syncwarp();
                    if (++kcb == p.k_blocks_per_tap) { kcb = 0; tap++; if (++kx == p.kw) { kx = 0; ky++; } }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform control flow, leader-predicated issue) =====================
        const uint32_t idesc = make_idesc(p.b_kmajor ? 0 : 1, p.bn);
        // Descriptor templates: everything but the 14-bit start address is loop-invariant.
        //   A, K-major SW128: 8-row groups 1024 B apart; K advances 32 B inside the swizzle row.
        //   B, K-major: same.  B, MN-major SW128: two 64-column atoms 8192 B apart (LBO), 8-row k-groups 1024 B apart (SBO);
        //   K advances 16 rows = 2048 B.
        const uint64_t adesc0 = make_smem_desc(smem_u32(smem_a), 16, 1024);
        const uint64_t bdesc0 = p.b_kmajor ? make_smem_desc(smem_u32(smem_b), 16, 1024) : make_smem_desc(smem_u32(smem_b), B_STAGE_BYTES / 2, 1024);
        const uint32_t b_kstep = p.b_kmajor ? (UMMA_K * 2) >> 4 : (UMMA_K * 128) >> 4;   // descriptor address units (16 B)
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(&acc_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
            const int sp = tile % p.split_k;
            const int kb_lo = sp * kb_per_split, kb_hi = min(kb_lo + kb_per_split, k_blocks_all);
            for (int kb = kb_lo; kb < kb_hi; kb++) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint64_t adesc = adesc0 + (uint64_t)(stage * (A_STAGE_BYTES >> 4));
                const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (B_STAGE_BYTES >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; k++)
                        umma_f16(tmem_d, adesc + (uint64_t)(k * ((UMMA_K * 2) >> 4)), bdesc + (uint64_t)(k * b_kstep), idesc,
                                 (kb != kb_lo || k != 0) ? 1u : 0u);
                    umma_commit(&empty[stage]);                          // frees the smem slot when these MMAs retire
                    if (kb == kb_hi - 1) umma_commit(&acc_full[acc]);    // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
        osb_pdl_trigger_late();   // every MMA of this CTA is issued: the next kernel may start its prologue on SMs that drain
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                 // TMEM lane quadrant this warp may access
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int sp = tile % p.split_k, t2 = tile / p.split_k;
            int b = t2 / tiles_per_batch, r = t2 % tiles_per_batch;
            int mt = r % p.m_tiles, nt = r / p.m_tiles;
            int n0 = nt * p.bn;
            const int n_end = min(p.N, n0 + p.bn);
            int row_in_tile = q * 32 + lane;
            long long out_row;   // row index into C (conv: output pixel index)
            bool row_ok;
            if (p.bh > 0) {
                int y = (mt / p.tiles_x) * p.bh + row_in_tile / p.bw, x = (mt % p.tiles_x) * p.bw + row_in_tile % p.bw;
                row_ok = y < p.Ho && x < p.Wo;
                out_row = (long long)y * p.Wo + x;
            } else {
                int m = mt * BLOCK_M + row_in_tile;
                row_ok = m < p.M;
                out_row = m;
            }
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            __half* cbase = p.C;
            if (p.groups > 1) cbase = b == 0 ? p.C : (b == 1 ? p.C1 : p.C2);    // stride_c == 0 in a grouped launch
            __half* crow = cbase + (long long)b * p.stride_c + out_row * p.ldc;
            const __half* rrow = p.residual ? p.residual + (long long)b * p.stride_c + out_row * p.ldc : nullptr;
            float* wrow = p.split_k > 1 ? p.ws + (((long long)sp * p.batch + b) * p.M + out_row) * p.N : nullptr;
            const bool vec_ok = (p.N & 7) == 0;
#pragma unroll 1
            for (int c = 0; c < p.bn; c += 32) {
                if (n0 + c >= n_end) break;     // warp-uniform
                uint32_t v[32];
                uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c);
                tmem_ld_32x32b_x32(taddr, v);
                if (wrow) {
                    // split-K: raw fp32 partials; the last CTA to arrive for this tile reduces them (below)
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            int n = n0 + c + j;
                            if (n + 3 < n_end && (p.N & 3) == 0) *reinterpret_cast<uint4*>(wrow + n) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                            else for (int t = 0; t < 4; t++) if (n + t < n_end) wrow[n + t] = __uint_as_float(v[j + t]);
                        }
                    }
                } else if (row_ok && !vec_ok) {
                    for (int j = 0; j < 32; j++) {
                        int n = n0 + c + j;
                        if (n >= n_end) break;
                        float f = __uint_as_float(v[j]);
                        if (p.bias) f += __half2float(p.bias[n]);
                        if (rrow) f += __half2float(rrow[n]);
                        crow[n] = __float2half_rn(f);
                    }
                } else if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        int n = n0 + c + j;
                        if (n >= n_end) break;  // N % 8 == 0 and bn % 16 == 0
                        float f[8];
#pragma unroll
                        for (int t = 0; t < 8; t++) f[t] = __uint_as_float(v[j + t]);
                        if (p.bias) {
                            Vec<__half, 8> bv = load_vec<__half, 8>(p.bias + n);
#pragma unroll
                            for (int t = 0; t < 8; t++) f[t] += __half2float(bv.v[t]);
                        }
                        if (rrow) {
                            Vec<__half, 8> rv = load_vec<__half, 8>(rrow + n);
#pragma unroll
                            for (int t = 0; t < 8; t++) f[t] += __half2float(rv.v[t]);
                        }
                        Vec<__half, 8> o;
#pragma unroll
                        for (int t = 0; t < 8; t++) o.v[t] = __float2half_rn(f[t]);
                        store_vec<__half, 8>(crow + n, o);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            if (wrow && p.counters) {
                // Parallel in-kernel split-K reduction (no second launch).  All split_k CTAs of a tile are co-resident (the host
                // keeps tiles * split_k <= resident CTA slots), so they can rendezvous on a counter: publish partials -> arrive ->
                // wait for everyone -> each CTA reduces its own slice of the tile's rows in fp32, adds bias / residual, rounds once.
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (warp == 2 && lane == 0) {
                    atomicAdd(&p.counters[2 * t2], 1);
                    long long t0 = clock64();
                    while (true) {
                        int seen;
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.counters + 2 * t2) : "memory");
                        if (seen >= p.split_k) break;
                        if (clock64() - t0 > 4000000000LL) { printf("tc_gemm_kernel: split-K rendezvous timed out (block %d)\n", blockIdx.x); __trap(); }
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                __threadfence();
                {
                    const int rows_per = (BLOCK_M + p.split_k - 1) / p.split_k;
                    const int r_lo = sp * rows_per, r_hi = min(r_lo + rows_per, BLOCK_M);
                    const long long plane = (long long)p.batch * p.M * p.N;
                    const int nvec = (n_end - n0) >> 2;      // N % 4 == 0 (host guarantees it on this path)
                    for (int rr = r_lo + q; rr < r_hi; rr += 4) {   // one tile row per epilogue warp and pass; lanes sweep the columns
                        long long orow_idx; bool ok;
                        if (p.bh > 0) {
                            int y = (mt / p.tiles_x) * p.bh + rr / p.bw, x = (mt % p.tiles_x) * p.bw + rr % p.bw;
                            ok = y < p.Ho && x < p.Wo; orow_idx = (long long)y * p.Wo + x;
                        } else { int m = mt * BLOCK_M + rr; ok = m < p.M; orow_idx = m; }
                        if (!ok) continue;
                        const float* src0 = p.ws + ((long long)b * p.M + orow_idx) * p.N;
                        __half* dst = p.C + (long long)b * p.stride_c + orow_idx * p.ldc;
                        const __half* res = p.residual ? p.residual + (long long)b * p.stride_c + orow_idx * p.ldc : nullptr;
                        for (int v = lane; v < nvec; v += 32) {
                            int n = n0 + (v << 2);
                            float4 a = __ldcg(reinterpret_cast<const float4*>(src0 + n));
                            for (int sidx = 1; sidx < p.split_k; sidx++) {
                                float4 t = __ldcg(reinterpret_cast<const float4*>(src0 + sidx * plane + n));
                                a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                            }
                            if (p.bias) { a.x += __half2float(p.bias[n]); a.y += __half2float(p.bias[n + 1]); a.z += __half2float(p.bias[n + 2]); a.w += __half2float(p.bias[n + 3]); }
                            if (res) { Vec<__half, 4> r4 = load_vec<__half, 4>(res + n); a.x += __half2float(r4.v[0]); a.y += __half2float(r4.v[1]); a.z += __half2float(r4.v[2]); a.w += __half2float(r4.v[3]); }
                            Vec<__half, 4> o4;
                            o4.v[0] = __float2half_rn(a.x); o4.v[1] = __float2half_rn(a.y); o4.v[2] = __float2half_rn(a.z); o4.v[3] = __float2half_rn(a.w);
                            store_vec<__half, 4>(dst + n, o4);
                        }
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (warp == 2 && lane == 0) {
                    int done = atomicAdd(&p.counters[2 * t2 + 1], 1);
                    if (done == p.split_k - 1) { p.counters[2 * t2] = 0; p.counters[2 * t2 + 1] = 0; __threadfence(); }   // re-arm for the next launch
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// split-K second pass: out[row][n] = fp16(sum_s ws[s][row][n] + bias[n] + residual[row][n]); rows = batch * M
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, __half* __restrict__ out, const __half* __restrict__ bias,
                                     const __half* __restrict__ residual, long long rows, int N, int splits)
{
    osb_pdl_prologue();
    // N % 4 == 0: one float4 of every split plane per thread, fully coalesced
    const long long total4 = rows * N / 4;
    const long long plane4 = total4;
    const float4* w4 = reinterpret_cast<const float4*>(ws);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        float4 a = w4[i];
        for (int s = 1; s < splits; s++) { float4 t = w4[(long long)s * plane4 + i]; a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
        long long e = i * 4;
        int n = (int)(e % N);
        if (bias) { a.x += __half2float(bias[n]); a.y += __half2float(bias[n + 1]); a.z += __half2float(bias[n + 2]); a.w += __half2float(bias[n + 3]); }
        if (residual) {
            Vec<__half, 4> r = load_vec<__half, 4>(residual + e);
            a.x += __half2float(r.v[0]); a.y += __half2float(r.v[1]); a.z += __half2float(r.v[2]); a.w += __half2float(r.v[3]);
        }
        Vec<__half, 4> o;
        o.v[0] = __float2half_rn(a.x); o.v[1] = __float2half_rn(a.y); o.v[2] = __float2half_rn(a.z); o.v[3] = __float2half_rn(a.w);
        store_vec<__half, 4>(out + e, o);
    }
}

// ---- host side -------------------------------------------------------------------------------------------------
constexpr size_t WS_MAX = OSB_WS_SPLITK_BYTES;   // fixed-capacity per-stream workspace (workspace.h): never re-allocated, graph-safe

// pick a split factor: fill the SMs when the tile count is small, keep >= 2 k-blocks per split, stay inside the workspace
int choose_split(int tiles, int k_blocks, size_t out_elems, cudaStream_t st, OsbWorkspace** ws_out)
{
    static const int forced = [] { const char* e = getenv("OSB_TC_SPLIT"); return e ? atoi(e) : 0; }();   // tuning experiments only
    // measured over every tc shape of the SD 1.5 UNet (r01 sweep): below ~32 k-blocks the second launch (the reduce) costs more
    // than the idle SMs do
    *ws_out = nullptr;
    if ((tiles >= 100 && forced <= 0) || k_blocks < 4 || (k_blocks < 32 && forced <= 0)) return 1;
    int split = forced > 0 ? forced : 148 / tiles;
    split = std::min(split, k_blocks / 2);
    while (split > 1 && (size_t)split * out_elems * 4 > WS_MAX) split--;
    if (split <= 1) return 1;
    int kb_per = (k_blocks + split - 1) / split;
    split = (k_blocks + kb_per - 1) / kb_per;          // no empty splits: every CTA must run at least one k-block
    if (split <= 1) return 1;
    OsbWorkspace* ws = osb_workspace(st, OSB_WS_SPLITK);
    if (!ws) return 1;                                  // capturing before any eager run, or out of memory: run unsplit
    *ws_out = ws;
    return split;
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    });
    return fn;
}

// rank-3 fp16 tensor map with 128B swizzle; dims/strides innermost first
bool make_map(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes, uint64_t s2_bytes,
              uint32_t b0, uint32_t b1, uint32_t b2, uint32_t traversal_stride = 1)
{
    auto enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = { d0, d1, d2 };
    cuuint64_t strides[2] = { s1_bytes, s2_bytes };
    cuuint32_t box[3] = { b0, b1, b2 };
    cuuint32_t estr[3] = { 1, traversal_stride, traversal_stride };   // strided conv: every s-th pixel of the box span
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// rank-3 map over (inner, row, batch) that keeps the global strides ascending: when the batch stride is the smaller one
// (per-head slices of a [T, heads*d] buffer) the two outer dimensions are swapped and *swapped is set.
bool make_map_rb(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t batch, uint64_t row_stride_bytes, uint64_t batch_stride_bytes,
                 uint32_t box_inner, uint32_t box_rows, int* swapped)
{
    if (batch > 1 && batch_stride_bytes < row_stride_bytes) {
        *swapped = 1;
        return make_map(map, base, inner, batch, rows, batch_stride_bytes, row_stride_bytes, box_inner, 1, box_rows);
    }
    *swapped = 0;
    return make_map(map, base, inner, rows, batch, row_stride_bytes, batch_stride_bytes, box_inner, box_rows, 1);
}

int num_sms()
{
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

// optional per-launch timing (bench.py's roofline leg): CUDA events on the launching stream around every launch
struct ProfRec { cudaEvent_t a, b; double flops, bytes; int M, N, K, taps, batch, split, conv; };
bool g_prof = false;
std::vector<ProfRec> g_prof_list;

// Few k-blocks per CTA => latency-bound: prefer the 3-stage variant (2 CTAs / SM).  OSB_TC_SHORT=0 disables it (A/B runs).
int short_k_hint(int k_blocks, int split)
{
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("OSB_TC_SHORT"); enabled = (e && e[0] == '0') ? 0 : 1; }
    int per_cta = (k_blocks + split - 1) / std::max(split, 1);
    return enabled && per_cta <= 12 ? 1 : 0;
}

bool inkernel_reduce()
{
    static int v = -1;
    // measured on B200 (SD1.5 UNet step): the rendezvous costs more than the 4 us reduce kernel it saves (9.17 vs 8.00 ms per
    // step), so the separate vectorised reduce kernel is the default; OSB_TC_INKERNEL_REDUCE=1 selects the in-kernel variant.
    if (v < 0) { const char* e = getenv("OSB_TC_INKERNEL_REDUCE"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

int launch(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, cudaStream_t st, const CUtensorMap* mb1p = nullptr, const CUtensorMap* mb2p = nullptr)
{
    const CUtensorMap& mb1 = mb1p ? *mb1p : mb;
    const CUtensorMap& mb2 = mb2p ? *mb2p : mb;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<STAGES_DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_for(STAGES_DEEP));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_gemm_kernel<STAGES_SHORT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_for(STAGES_SHORT));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    int total = p.m_tiles * p.n_tiles * p.batch * p.split_k;
    const bool short_k = p.short_k != 0;
    int grid = std::min(total, num_sms() * (short_k ? 2 : 1));
    ProfRec rec{};
    if (g_prof) {
        cudaEventCreate(&rec.a); cudaEventCreate(&rec.b);
        double M = p.M, N = p.N, Kt = (double)p.K * p.taps, B = p.batch;
        rec.flops = 2.0 * M * N * Kt * B;
        // algorithmic bytes: A once (conv: the input image once), B once, C once (+ residual / bias reads)
        double a_bytes = (p.bh > 0 ? M * p.K : M * Kt) * 2.0 * B;
        rec.bytes = a_bytes + N * Kt * 2.0 * (p.bh > 0 ? 1.0 : B) + M * N * 2.0 * B * (p.residual ? 2.0 : 1.0) + (p.bias ? N * 2.0 : 0.0);
        rec.M = p.M; rec.N = p.N; rec.K = p.K; rec.taps = p.taps; rec.batch = p.batch; rec.split = p.split_k; rec.conv = p.bh > 0;
        cudaEventRecord(rec.a, st);
    }
    if (short_k) osb_launch((tc_gemm_kernel<STAGES_SHORT>), grid, NUM_THREADS, (size_t)smem_bytes_for(STAGES_SHORT), st, ma, mb, mb1, mb2, p);
    else osb_launch((tc_gemm_kernel<STAGES_DEEP>), grid, NUM_THREADS, (size_t)smem_bytes_for(STAGES_DEEP), st, ma, mb, mb1, mb2, p);
    if (p.split_k > 1 && !p.counters) {
        launched(1);
        long long total4 = (long long)p.batch * p.M * p.N / 4;
        int rgrid = (int)std::min<long long>((total4 + 255) / 256, 148 * 8);
        osb_launch((splitk_reduce_kernel), rgrid, 256, 0, st, (const float*)p.ws, p.C, p.bias, p.residual, (long long)p.batch * p.M, p.N, p.split_k);
        if (g_prof) { cudaEventRecord(rec.b, st); g_prof_list.push_back(rec); }
        return launched(0);
    }
    if (g_prof) { cudaEventRecord(rec.b, st); g_prof_list.push_back(rec); }
    return launched(1);
}

// Tile width: fewest waves over the 148 SMs, then least padded work (e.g. N = 320 at M = 4096: 80 -> 128 tiles in one wave
// with no padding, instead of 96 tiles of 128 with 1/6 of every third tile wasted).
int env_int(const char* name)
{
    const char* e = getenv(name);
    return e ? atoi(e) : 0;
}

int choose_bn(int64_t m_tiles, int64_t N, int64_t batch)
{
    static const int forced = env_int("OSB_TC_BN");     // tuning experiments only
    if (forced == 128 || forced == 96 || forced == 80 || forced == 64) return (N <= 64 && forced > 64) ? 64 : forced;
    // Measured (r01): the mainloop is bound by L2->SM bytes per k-block, (128 + bn) * 128 B, so the widest tile wins whenever the
    // K loop is long; split-K (not a narrower tile) supplies parallelism.  Narrow tiles only where N itself is narrow.
    (void)m_tiles; (void)batch;
    if (N <= 64) return 64;
    if (N <= 80) return 80;
    if (N <= 96) return 96;
    return 128;
}

inline uint32_t next_pow2(uint32_t v) { uint32_t r = 1; while (r < v) r <<= 1; return r; }

}  // namespace

extern "C" void osb_tc_profile(int enable)
{
    for (auto& r : g_prof_list) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof_list.clear();
    g_prof = enable != 0;
}

// Sums over the launches recorded since osb_tc_profile(1): out = { launches, total ms, total flops, total algorithmic bytes }
extern "C" int osb_tc_profile_read(double* out4)
{
    double ms = 0, fl = 0, by = 0;
    for (auto& r : g_prof_list) {
        if (cudaEventSynchronize(r.b) != cudaSuccess) return -1;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return -1;
        ms += t; fl += r.flops; by += r.bytes;
    }
    out4[0] = (double)g_prof_list.size(); out4[1] = ms; out4[2] = fl; out4[3] = by;
    return 0;
}

// One text line per recorded launch: "M N K taps batch split conv ms gflop"
extern "C" int osb_tc_profile_dump(char* buf, int cap)
{
    int off = 0;
    for (auto& r : g_prof_list) {
        float t = 0.f;
        if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return -1;
        int n = snprintf(buf + off, cap - off, "%d %d %d %d %d %d %d %.4f %.3f\n", r.M, r.N, r.K, r.taps, r.batch, r.split, r.conv, t, r.flops * 1e-9);
        if (n < 0 || off + n >= cap) break;
        off += n;
    }
    return off;
}

bool osb_tc_gemm_ok(int64_t M, int64_t N, int64_t K, int bt, const void* A, const void* B, const void* C, int64_t sa, int64_t sb, int64_t sc,
                    int64_t lda, int64_t ldb, int64_t ldc)
{
    if ((lda % 8) || (ldb % 8) || (ldc % 8)) return false;
    if (M < 32 || N < 1 || K < 8) return false;
    if (K % 8) return false;
    if (N % 8) return false;   // ragged N only through the conv entry (K-major B, scalar epilogue)
    if (M > (1 << 30) || N > (1 << 30) || K > (1 << 30)) return false;
    if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) return false;
    if ((sa % 8) || (sb % 8) || (sc % 8)) return false;
    (void)bt;
    return get_encode() != nullptr || A == nullptr;
}

int osb_tc_gemm_launch(const void* A, const void* B, void* C, const void* bias, const void* residual, int64_t batch, int64_t M, int64_t N, int64_t K,
                       int64_t sa, int64_t sb, int64_t sc, int bt, cudaStream_t st, int64_t lda, int64_t ldb, int64_t ldc)
{
    if (lda <= 0) lda = K;
    if (ldb <= 0) ldb = bt ? K : N;
    if (ldc <= 0) ldc = N;
    CUtensorMap ma, mb;
    // A: [batch][M][K]; a shared operand (stride 0) is presented as batch extent 1 and the batch coordinate ignored
    uint64_t abatch = sa ? (uint64_t)batch : 1, bbatch = sb ? (uint64_t)batch : 1;
    if (batch > 1 && (!sa || !sb)) {
        // shared operands across the batch: fold the batch into per-batch launches (rare: 2-D weights with n > 1)
        for (int64_t i = 0; i < batch; i++) {
            int r = osb_tc_gemm_launch((const __half*)A + i * sa, (const __half*)B + i * sb, (__half*)C + i * sc, bias,
                                       residual ? (const __half*)residual + i * sc : nullptr, 1, M, N, K, M * lda, bt ? N * ldb : K * ldb, sc, bt, st, lda, ldb, ldc);
            if (r) return r;
        }
        return 0;
    }
    int a_swap = 0, b_swap = 0;
    if (!make_map_rb(&ma, A, (uint64_t)K, (uint64_t)M, abatch, (uint64_t)lda * 2, (uint64_t)(sa ? sa : M * lda) * 2, BLOCK_K, BLOCK_M, &a_swap)) return (int)cudaErrorInvalidValue;
    int64_t m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    int bn = choose_bn(m_tiles, N, batch);
    bool okb = bt ? make_map_rb(&mb, B, (uint64_t)K, (uint64_t)N, bbatch, (uint64_t)ldb * 2, (uint64_t)(sb ? sb : N * ldb) * 2, BLOCK_K, (uint32_t)bn, &b_swap)
                  : make_map_rb(&mb, B, (uint64_t)N, (uint64_t)K, bbatch, (uint64_t)ldb * 2, (uint64_t)(sb ? sb : K * ldb) * 2, 64, BLOCK_K, &b_swap);
    if (!okb) return (int)cudaErrorInvalidValue;
    TcParams p{};
    p.M = (int)M; p.N = (int)N; p.K = (int)K; p.batch = (int)batch;
    p.bn = bn;
    p.m_tiles = (int)m_tiles; p.n_tiles = (int)((N + bn - 1) / bn);
    p.b_kmajor = bt ? 1 : 0;
    p.a_swap = a_swap; p.b_swap = b_swap;
    p.taps = 1; p.kw = 1; p.bh = 0; p.bw = 0; p.tiles_x = 1;
    p.k_blocks_per_tap = (int)((K + BLOCK_K - 1) / BLOCK_K);
    p.stride = 1;
    p.C = (__half*)C; p.bias = (const __half*)bias; p.residual = (const __half*)residual; p.stride_c = sc; p.ldc = ldc;
    OsbWorkspace* wsp = nullptr;
    p.split_k = (ldc == N && (sc == M * N || batch == 1)) ? choose_split(p.m_tiles * p.n_tiles * p.batch, p.k_blocks_per_tap, (size_t)batch * M * N, st, &wsp) : 1;
    p.ws = wsp ? wsp->splitk : nullptr;
    // in-kernel rendezvous reduction needs every CTA resident at once and float4-aligned rows; otherwise the reduce kernel runs
    p.counters = (p.split_k > 1 && p.N % 4 == 0 && p.ldc % 4 == 0 && (long long)p.m_tiles * p.n_tiles * p.batch * p.split_k <= num_sms() && inkernel_reduce() && wsp) ? wsp->splitk_counters : nullptr;
    p.short_k = short_k_hint(p.taps * p.k_blocks_per_tap, p.split_k);
    return launch(ma, mb, p, st);
}

// `groups` (2 or 3) GEMMs C_g = A * B_g that share A and the shape, in ONE launch (the q/k/v projections of an attention block):
// three times the tiles per launch, one dependency instead of three.  No bias / residual / split-K.
int osb_tc_gemm_grouped_launch(const void* A, const void* const* B, void* const* C, int groups, int64_t M, int64_t N, int64_t K, int bt, cudaStream_t st,
                               int64_t lda, int64_t ldb, int64_t ldc)
{
    if (groups < 2 || groups > 3) return (int)cudaErrorInvalidValue;
    CUtensorMap ma, mb[3];
    int a_swap = 0, b_swap = 0;
    if (!make_map_rb(&ma, A, (uint64_t)K, (uint64_t)M, 1, (uint64_t)lda * 2, (uint64_t)(M * lda) * 2, BLOCK_K, BLOCK_M, &a_swap)) return (int)cudaErrorInvalidValue;
    int64_t m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    int bn = choose_bn(m_tiles, N, groups);
    for (int g = 0; g < groups; g++) {
        int sw = 0;
        bool ok = bt ? make_map_rb(&mb[g], B[g], (uint64_t)K, (uint64_t)N, 1, (uint64_t)ldb * 2, (uint64_t)(N * ldb) * 2, BLOCK_K, (uint32_t)bn, &sw)
                     : make_map_rb(&mb[g], B[g], (uint64_t)N, (uint64_t)K, 1, (uint64_t)ldb * 2, (uint64_t)(K * ldb) * 2, 64, BLOCK_K, &sw);
        if (!ok || (g > 0 && sw != b_swap)) return (int)cudaErrorInvalidValue;
        b_swap = sw;
    }
    TcParams p{};
    p.M = (int)M; p.N = (int)N; p.K = (int)K; p.batch = groups; p.groups = groups;
    p.bn = bn;
    p.m_tiles = (int)m_tiles; p.n_tiles = (int)((N + bn - 1) / bn);
    p.b_kmajor = bt ? 1 : 0;
    p.a_swap = a_swap; p.b_swap = b_swap;
    p.taps = 1; p.kw = 1; p.bh = 0; p.bw = 0; p.tiles_x = 1;
    p.k_blocks_per_tap = (int)((K + BLOCK_K - 1) / BLOCK_K);
    p.stride = 1;
    p.C = (__half*)C[0]; p.C1 = (__half*)C[1]; p.C2 = (__half*)(groups > 2 ? C[2] : C[1]);
    p.bias = nullptr; p.residual = nullptr; p.stride_c = 0; p.ldc = ldc;
    p.split_k = 1; p.ws = nullptr; p.counters = nullptr;
    p.short_k = short_k_hint(p.taps * p.k_blocks_per_tap, 1);
    return launch(ma, mb[0], p, st, &mb[1], groups > 2 ? &mb[2] : &mb[1]);
}

bool osb_tc_conv_ok(int64_t H, int64_t W, int64_t Cin, int64_t Cout, int kh, int kw, int stride, const void* x, const void* w, const void* y)
{
    if (stride < 1 || stride > 2) return false;
    if (Cin % 8 || Cin < 16) return false;     // TMA needs 16-byte pixel strides; tiny-Cin stems stay on the CUDA-core kernel
    if (H * W < 64) return false;
    if (kh > 7 || kw > 7) return false;
    if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) return false;
    return get_encode() != nullptr;
}

int osb_tc_conv_launch(const void* x, const void* w, const void* bias, const void* residual, void* y, int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                       int kh, int kw, int stride, int pad_top, int pad_left, int64_t Ho, int64_t Wo, cudaStream_t st)
{
    uint32_t bw = std::min<uint32_t>(128, next_pow2((uint32_t)Wo)), bh = 128 / bw;
    CUtensorMap ma, mb;
    // A: NHWC input as (C, W, H); one box = bh rows x bw pixels x 64 channels, zero-filled outside the image.  With a
    // traversal stride s the box spans bw*s x bh*s input pixels and TMA delivers every s-th one.
    if (!make_map(&ma, x, (uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)Cin * 2, (uint64_t)W * Cin * 2, BLOCK_K, bw * stride, bh * stride, (uint32_t)stride))
        return (int)cudaErrorInvalidValue;
    // B: OHWI weights = [Cout][kh*kw*Cin], K-major
    int64_t Ktot = (int64_t)kh * kw * Cin;
    int64_t tiles_x_ = (Wo + bw - 1) / bw, m_tiles_ = tiles_x_ * ((Ho + bh - 1) / bh);
    int bn = choose_bn(m_tiles_, Cout, 1);
    if (!make_map(&mb, w, (uint64_t)Ktot, (uint64_t)Cout, 1, (uint64_t)Ktot * 2, (uint64_t)Ktot * Cout * 2, BLOCK_K, (uint32_t)bn, 1)) return (int)cudaErrorInvalidValue;
    TcParams p{};
    p.bn = bn;
    p.M = (int)(Ho * Wo); p.N = (int)Cout; p.K = (int)Cin; p.batch = 1;
    p.tiles_x = (int)((Wo + bw - 1) / bw);
    p.m_tiles = p.tiles_x * (int)((Ho + bh - 1) / bh);
    p.n_tiles = (int)((Cout + bn - 1) / bn);
    p.b_kmajor = 1;
    p.taps = kh * kw; p.kw = kw; p.pad_top = pad_top; p.pad_left = pad_left; p.Wo = (int)Wo; p.Ho = (int)Ho; p.bw = (int)bw; p.bh = (int)bh;
    p.k_blocks_per_tap = (int)((Cin + BLOCK_K - 1) / BLOCK_K);
    p.stride = stride;
    p.C = (__half*)y; p.bias = (const __half*)bias; p.residual = (const __half*)residual; p.stride_c = 0; p.ldc = Cout;
    // the split-K reduce paths move float4 / half4 vectors: ragged Cout (conv_out, 3 or 4 channels) runs unsplit
    OsbWorkspace* wsp = nullptr;
    p.split_k = (Cout % 4 == 0) ? choose_split(p.m_tiles * p.n_tiles, p.taps * p.k_blocks_per_tap, (size_t)Ho * Wo * Cout, st, &wsp) : 1;
    p.ws = wsp ? wsp->splitk : nullptr;
    // in-kernel rendezvous reduction needs every CTA resident at once and float4-aligned rows; otherwise the reduce kernel runs
    p.counters = (p.split_k > 1 && p.N % 4 == 0 && p.ldc % 4 == 0 && (long long)p.m_tiles * p.n_tiles * p.batch * p.split_k <= num_sms() && inkernel_reduce() && wsp) ? wsp->splitk_counters : nullptr;
    p.short_k = short_k_hint(p.taps * p.k_blocks_per_tap, p.split_k);
    return launch(ma, mb, p, st);
}
