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
#include "tc_ptx.cuh"
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
constexpr int GN_SMEM_BYTES = 4 * 2048 + 2 * tcptx::GN_MAX_GROUPS * 4;   // GroupNorm statistics: 4 warp-private transposition buffers + the CTA accumulators
constexpr int smem_bytes_for(int stages, bool extras = true) { return stages * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/ + (extras ? GN_SMEM_BYTES : 0); }
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
    const __half* bias2;         // second per-column addend (the time-embedding row a resnet adds to conv1's output), or null
    const __half* residual;
    double* gn_stats;            // != null: per-group (sum, sum of squares) of the stored output, for the GroupNorm that consumes it
    int gn_cpg, gn_groups;       // channels per group, number of groups (N == gn_cpg * gn_groups)
    int gn_debug;                // bisecting aid (OSB_GN_DEBUG): 1 = skip the per-chunk gathering, 2 = skip the per-tile flush, 3 = both
    int bf16;                    // operands are bfloat16 (the fp32 path: bf16 triple-split operands, see osb_tc_gemm_f32x)
    int f32_out;                 // raw fp32 accumulators go to `ws` even when split_k == 1; the reduce kernel writes fp32 C (+ fp32 bias / residual)
    long long stride_c;          // elements between batches
    long long ldc;               // elements between output rows (== N for a dense C)
};

using namespace tcptx;

// instruction descriptor for kind::f16: fp16 x fp16 -> fp32, A K-major, B K- or MN-major
__device__ __forceinline__ uint32_t make_idesc(int b_mn_major, int bn, int bf16 = 0)
{
    uint32_t d = 0;
    d |= 1u << 4;                              // c_format = F32
    d |= (bf16 ? 1u : 0u) << 7;                // a_format = F16 / BF16
    d |= (bf16 ? 1u : 0u) << 10;               // b_format = F16 / BF16
    d |= 0u << 15;                             // a_major  = K
    d |= (uint32_t)(b_mn_major ? 1 : 0) << 16; // b_major
    d |= (uint32_t)(bn >> 3) << 17;            // n_dim
    d |= (uint32_t)(BLOCK_M >> 4) << 24;       // m_dim
    return d;
}

// ---- the kernel -----------------------------------------------------------------------------------------------

// EXTRAS = the epilogue also adds `bias2` and gathers GroupNorm statistics.  A separate instantiation: the extra code costs ~14 registers
// and ~0.3 ms over the 253 tensor-core launches of a UNet step when it rides along in every launch (measured, profiles/r02_ab_epilogue.txt).
template <int STAGES, bool EXTRAS>
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
    uint8_t* gn_buf = (uint8_t*)bars + 256;                 // [4][2048] warp-private chunk transposition buffers
    float* gn_acc = (float*)(gn_buf + 4 * 2048);            // [2 * GN_MAX_GROUPS] per-CTA (sum, sum of squares) accumulators

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (EXTRAS && p.gn_stats && threadIdx.x < 2 * GN_MAX_GROUPS) gn_acc[threadIdx.x] = 0.f;

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
        const uint32_t idesc = make_idesc(p.b_kmajor ? 0 : 1, p.bn, p.bf16);
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
            float* wrow = (p.split_k > 1 || p.f32_out) ? p.ws + (((long long)sp * p.batch + b) * p.M + out_row) * p.N : nullptr;
            const bool vec_ok = (p.N & 7) == 0;
            // In-kernel split-K, "last finisher reduces" (no second launch, no co-residency assumption): every CTA of a tile takes a ticket
            // AFTER its main loop.  Tickets 0 .. S-2 publish their fp32 partial and leave; the holder of ticket S-1 -- by construction every
            // other CTA of the tile has finished its main loop and is in, or past, an epilogue that never blocks -- waits for their `done`
            // signals, then runs the normal epilogue on its own TMEM accumulator plus the S-1 partials (L2-hot), re-arming the counters.
            bool last_finisher = false;
            const float* others = nullptr;       // plane 0 of the partial workspace, this row
            if (wrow && p.counters) {
                if (warp == 2 && lane == 0) *split_flag = atomicAdd(&p.counters[2 * t2], 1);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                last_finisher = *split_flag == p.split_k - 1;
                asm volatile("bar.sync 1, 128;" ::: "memory");          // everyone has read the flag before the next tile's ticket overwrites it
                if (last_finisher) {
                    if (warp == 2 && lane == 0) {
                        long long t0 = clock64();
                        while (true) {
                            int seen;
                            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.counters + 2 * t2 + 1) : "memory");
                            if (seen >= p.split_k - 1) break;
                            if (clock64() - t0 > 4000000000LL) { printf("tc_gemm_kernel: split-K partials never arrived (block %d)\n", blockIdx.x); __trap(); }
                        }
                        p.counters[2 * t2] = 0; p.counters[2 * t2 + 1] = 0;       // re-arm for the next launch (nobody else touches them any more)
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    __threadfence();
                    others = p.ws + ((long long)b * p.M + out_row) * p.N;
                    wrow = nullptr;                                             // take the normal (final) epilogue below
                }
            }
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
                        if (others) {
                            const long long plane = (long long)p.batch * p.M * p.N;
                            for (int sidx = 0; sidx < p.split_k; sidx++) {
                                if (sidx == sp) continue;
                                const float4 a0 = __ldcg(reinterpret_cast<const float4*>(others + sidx * plane + n));
                                const float4 a1 = __ldcg(reinterpret_cast<const float4*>(others + sidx * plane + n + 4));
                                f[0] += a0.x; f[1] += a0.y; f[2] += a0.z; f[3] += a0.w; f[4] += a1.x; f[5] += a1.y; f[6] += a1.z; f[7] += a1.w;
                            }
                        }
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
                        if (EXTRAS && p.bias2) {
                            Vec<__half, 8> bv = load_vec<__half, 8>(p.bias2 + n);
#pragma unroll
                            for (int t = 0; t < 8; t++) f[t] += __half2float(bv.v[t]);
                        }
                        Vec<__half, 8> o;
#pragma unroll
                        for (int t = 0; t < 8; t++) o.v[t] = __float2half_rn(f[t]);
                        store_vec<__half, 8>(crow + n, o);
                        if (EXTRAS) {
#pragma unroll
                            for (int t = 0; t < 4; t++) v[(j >> 1) + t] = *reinterpret_cast<uint32_t*>(&o.v[2 * t]);   // keep the rounded values for the statistics
                        }
                    }
                }
                if (EXTRAS && p.gn_stats && !wrow && !(p.gn_debug & 1)) {
                    // rows outside the problem (and units past N) contribute zeros
                    uint32_t h[16];
#pragma unroll
                    for (int t = 0; t < 16; t++) h[t] = (row_ok && vec_ok && n0 + c + 2 * t < n_end) ? v[t] : 0u;
                    gn_stats_chunk(smem_u32(gn_buf) + (uint32_t)q * 2048u, h, n0 + c, n_end, p.gn_cpg, gn_acc, lane);
                }
            }
            if (EXTRAS && p.gn_stats && !wrow && !(p.gn_debug & 2)) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                gn_stats_flush(gn_acc, p.gn_stats, p.gn_groups, (int)threadIdx.x - 64);
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            if (wrow && p.counters) {
                // a partial CTA: publish, signal `done`, leave
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (warp == 2 && lane == 0) { asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p.counters + 2 * t2 + 1) : "memory"); }
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

// split-K second pass: out[row][n] = fp16(sum_s ws[s][row][n] + bias[n] + bias2[n] + residual[row][n]); rows = batch * M.
// Optionally gathers the GroupNorm statistics of the result (per-block shared accumulators -> global fp64).
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, __half* __restrict__ out, const __half* __restrict__ bias, const __half* __restrict__ bias2,
                                     const __half* __restrict__ residual, long long rows, int N, int splits, double* __restrict__ gn_stats, int gn_cpg, int gn_groups)
{
    osb_pdl_prologue();
    __shared__ float acc[2 * GN_MAX_GROUPS];
    if (gn_stats) { if (threadIdx.x < 2 * GN_MAX_GROUPS) acc[threadIdx.x] = 0.f; __syncthreads(); }
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
        if (bias2) { a.x += __half2float(bias2[n]); a.y += __half2float(bias2[n + 1]); a.z += __half2float(bias2[n + 2]); a.w += __half2float(bias2[n + 3]); }
        if (residual) {
            Vec<__half, 4> r = load_vec<__half, 4>(residual + e);
            a.x += __half2float(r.v[0]); a.y += __half2float(r.v[1]); a.z += __half2float(r.v[2]); a.w += __half2float(r.v[3]);
        }
        Vec<__half, 4> o;
        o.v[0] = __float2half_rn(a.x); o.v[1] = __float2half_rn(a.y); o.v[2] = __float2half_rn(a.z); o.v[3] = __float2half_rn(a.w);
        store_vec<__half, 4>(out + e, o);
        if (gn_stats) {
            // the 4 columns of a thread lie in one group when cpg % 4 == 0 (host guarantees it)
            float x0 = __half2float(o.v[0]), x1 = __half2float(o.v[1]), x2 = __half2float(o.v[2]), x3 = __half2float(o.v[3]);
            const int g = n / gn_cpg;
            atomicAdd(&acc[2 * g], (x0 + x1) + (x2 + x3));
            atomicAdd(&acc[2 * g + 1], fmaf(x0, x0, x1 * x1) + fmaf(x2, x2, x3 * x3));
        }
    }
    if (gn_stats) {
        __syncthreads();
        if (threadIdx.x < 2 * gn_groups) { float v = acc[threadIdx.x]; if (v != 0.f) atomicAdd(&gn_stats[threadIdx.x], (double)v); }
    }
}

// fp32-output variant (the bf16 triple-split path): out[row][n] = sum_s ws[s][row][n] + bias[n] + residual[row][n], everything fp32, any N
__global__ void splitk_reduce_f32_kernel(const float* __restrict__ ws, float* __restrict__ out, const float* __restrict__ bias, const float* __restrict__ residual,
                                         long long rows, int N, int splits)
{
    osb_pdl_prologue();
    const long long total = rows * N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float a = ws[i];
        for (int s = 1; s < splits; s++) a += ws[(long long)s * total + i];
        if (bias) a += bias[(int)(i % N)];
        if (residual) a += residual[i];
        out[i] = a;
    }
}

#include "gemm_pair.cuh"
#include "gemm_i8.cuh"

// ---- host side -------------------------------------------------------------------------------------------------
constexpr size_t WS_MAX = OSB_WS_SPLITK_BYTES;   // fixed-capacity per-stream workspace (workspace.h): never re-allocated, graph-safe

bool inkernel_reduce();

// pick a split factor: fill the SMs when the tile count is small, keep >= 2 k-blocks per split, stay inside the workspace
int choose_split(int tiles, int k_blocks, size_t out_elems, cudaStream_t st, OsbWorkspace** ws_out)
{
    static const int forced = [] { const char* e = getenv("OSB_TC_SPLIT"); return e ? atoi(e) : 0; }();   // tuning experiments only
    // measured over every tc shape of the SD 1.5 UNet (r01 sweep): below ~32 k-blocks the second launch (the reduce) costs more
    // than the idle SMs do
    *ws_out = nullptr;
    static const int min_kb = [] { const char* e = getenv("OSB_TC_SPLIT_MINKB"); int v = e ? atoi(e) : 0; return v > 0 ? v : 32; }();
    if ((tiles >= 100 && forced <= 0) || k_blocks < 4 || (k_blocks < min_kb && forced <= 0)) return 1;
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
              uint32_t b0, uint32_t b1, uint32_t b2, uint32_t traversal_stride = 1, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT16)
{
    auto enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = { d0, d1, d2 };
    cuuint64_t strides[2] = { s1_bytes, s2_bytes };
    cuuint32_t box[3] = { b0, b1, b2 };
    cuuint32_t estr[3] = { 1, traversal_stride, traversal_stride };   // strided conv: every s-th pixel of the box span
    CUresult r = enc(map, dtype, 3, const_cast<void*>(base), dims, strides, box, estr,
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
    // Opt-in (OSB_TC_INKERNEL_REDUCE=1).  Round 1: all CTAs of a tile rendezvous and each reduces a slice -- 9.17 vs 8.00 ms per UNet step.
    // Round 2: "last finisher reduces" (deadlock-free without co-residency, see the epilogue) -- correct, but the one CTA that finishes last
    // pulls (S-1) fp32 planes of its tile through a single SM's L2 port (S up to 14: ~850 KB, ~9 us): 7.34 vs 5.38 ms per step with the
    // separate reduce kernel, which spreads the same bytes over every SM.  The vectorised reduce kernel stays the default.
    if (v < 0) { const char* e = getenv("OSB_TC_INKERNEL_REDUCE"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

void prof_begin(ProfRec& rec, const TcParams& p, cudaStream_t st)
{
    cudaEventCreate(&rec.a); cudaEventCreate(&rec.b);
    // the fp32 path (bf16 triple split) runs a 6x longer K: ALGORITHMIC work is the fp32 problem's (K / 6, 4-byte elements)
    const double kdiv = p.bf16 ? 6.0 : 1.0, es = p.bf16 ? 4.0 : 2.0;
    double M = p.M, N = p.N, Kt = (double)p.K * p.taps / kdiv, B = p.batch;
    rec.flops = 2.0 * M * N * Kt * B;
    // algorithmic bytes: A once (conv: the input image once), B once, C once (+ residual / bias reads)
    double a_bytes = (p.bh > 0 ? M * (p.K / kdiv) : M * Kt) * es * B;
    rec.bytes = a_bytes + N * Kt * es * (p.bh > 0 ? 1.0 : B) + M * N * es * B * (p.residual ? 2.0 : 1.0) + (p.bias ? N * es : 0.0);
    rec.M = p.M; rec.N = p.N; rec.K = p.K; rec.taps = p.taps; rec.batch = p.batch; rec.split = p.split_k; rec.conv = p.bh > 0;
    cudaEventRecord(rec.a, st);
}

bool carveout_max()
{
    static const bool v = [] { const char* e = getenv("OSB_SMEM_CARVEOUT"); return e && e[0] == '1'; }();
    return v;
}

int launch(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, cudaStream_t st, const CUtensorMap* mb1p = nullptr, const CUtensorMap* mb2p = nullptr)
{
    const CUtensorMap& mb1 = mb1p ? *mb1p : mb;
    const CUtensorMap& mb2 = mb2p ? *mb2p : mb;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<STAGES_DEEP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_for(STAGES_DEEP, false));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_gemm_kernel<STAGES_SHORT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_for(STAGES_SHORT, false));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_gemm_kernel<STAGES_DEEP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_for(STAGES_DEEP, true));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_gemm_kernel<STAGES_SHORT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_for(STAGES_SHORT, true));
        if (e != cudaSuccess) return (int)e;
        if (carveout_max()) {
            // one shared-memory carve-out for every tensor-core kernel: consecutive kernels that need different carve-outs make the SMs
            // drain and reconfigure (the EXTRAS instantiations need > 196 KiB per SM, the lean ones do not)
            cudaFuncSetAttribute(tc_gemm_kernel<STAGES_DEEP, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(tc_gemm_kernel<STAGES_SHORT, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(tc_gemm_kernel<STAGES_DEEP, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(tc_gemm_kernel<STAGES_SHORT, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
        attr_set = true;
    }
    int total = p.m_tiles * p.n_tiles * p.batch * p.split_k;
    const bool short_k = p.short_k != 0;
    int grid = std::min(total, num_sms() * (short_k ? 2 : 1));
    ProfRec rec{};
    if (g_prof) prof_begin(rec, p, st);
    const bool extras = p.bias2 != nullptr || p.gn_stats != nullptr;
    if (short_k && extras) osb_launch((tc_gemm_kernel<STAGES_SHORT, true>), grid, NUM_THREADS, (size_t)smem_bytes_for(STAGES_SHORT, true), st, ma, mb, mb1, mb2, p);
    else if (short_k) osb_launch((tc_gemm_kernel<STAGES_SHORT, false>), grid, NUM_THREADS, (size_t)smem_bytes_for(STAGES_SHORT, false), st, ma, mb, mb1, mb2, p);
    else if (extras) osb_launch((tc_gemm_kernel<STAGES_DEEP, true>), grid, NUM_THREADS, (size_t)smem_bytes_for(STAGES_DEEP, true), st, ma, mb, mb1, mb2, p);
    else osb_launch((tc_gemm_kernel<STAGES_DEEP, false>), grid, NUM_THREADS, (size_t)smem_bytes_for(STAGES_DEEP, false), st, ma, mb, mb1, mb2, p);
    if (p.f32_out) {
        launched(1);
        const long long total = (long long)p.batch * p.M * p.N;
        int rgrid = (int)std::min<long long>((total + 255) / 256, 148 * 8);
        osb_launch((splitk_reduce_f32_kernel), rgrid, 256, 0, st, (const float*)p.ws, (float*)p.C, (const float*)p.bias, (const float*)p.residual, (long long)p.batch * p.M, p.N, p.split_k);
        if (g_prof) { cudaEventRecord(rec.b, st); g_prof_list.push_back(rec); }
        return launched(0);
    }
    if (p.split_k > 1 && !p.counters) {
        launched(1);
        long long total4 = (long long)p.batch * p.M * p.N / 4;
        int rgrid = (int)std::min<long long>((total4 + 255) / 256, 148 * 8);
        osb_launch((splitk_reduce_kernel), rgrid, 256, 0, st, (const float*)p.ws, p.C, p.bias, p.bias2, p.residual, (long long)p.batch * p.M, p.N, p.split_k,
                   p.gn_stats, p.gn_cpg, p.gn_groups);
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


// ---- CTA-pair kernel: host side -------------------------------------------------------------------------------------
// Cost model (r01 finding: these kernels are bound by L2->SM bytes per CTA and k-block, and by waves): a launch costs
// waves x bytes-per-CTA-per-k-block.  single: 128 x bn tiles over 148 SMs, (128 + bn) x 128 B; pair: 256 x bn tiles over 74 SM
// pairs, (128 + bn / 2) x 128 B per CTA.  OSB_TC_PAIR=0 disables the pair kernel, =2 forces it wherever it is eligible.
int g_pair_mode = -1;
int pair_mode()
{
    if (g_pair_mode < 0) { const char* e = getenv("OSB_TC_PAIR"); g_pair_mode = e ? atoi(e) : 1; }
    return g_pair_mode;
}

int choose_pair_bn(int64_t m_tiles, int64_t N, int64_t batch, bool b_kmajor, double* cost)
{
    static const int forced = env_int("OSB_TC_PAIR_BN");
    const int cands_k[4] = { 64, 128, 192, 256 }, cands_mn[2] = { 128, 256 };
    const int* cands = b_kmajor ? cands_k : cands_mn;
    const int nc = b_kmajor ? 4 : 2;
    const int64_t m_pairs = (m_tiles + 1) / 2;
    const int pairs_hw = num_sms() / 2;
    int best = 0; double best_cost = 0;
    for (int i = 0; i < nc; i++) {
        int bn = cands[i];
        if (forced && bn != forced) continue;
        if (i > 0 && bn - 64 >= N) break;                    // wider than the problem
        int64_t tiles = m_pairs * ((N + bn - 1) / bn) * batch;
        int64_t waves = (tiles + pairs_hw - 1) / pairs_hw;
        double c = (double)waves * (16384.0 + 64.0 * bn);
        if (!best || c < best_cost) { best = bn; best_cost = c; }
    }
    *cost = best_cost;
    return best;
}

double single_cost(int64_t m_tiles, int64_t N, int64_t batch, int bn)
{
    int64_t tiles = m_tiles * ((N + bn - 1) / bn) * batch;
    int64_t waves = (tiles + num_sms() - 1) / num_sms();
    return (double)waves * (16384.0 + 128.0 * bn);
}

// pair kernel or single-CTA kernel?  `k_blocks` = length of the K loop, `tiles1` = tiles of the single-CTA decomposition
bool use_pair(int64_t m_tiles, int64_t N, int64_t batch, bool b_kmajor, int k_blocks, int bn1, int* bn_pair)
{
    const int mode = pair_mode();
    if (mode == 0 || m_tiles < 2 || (N % 8)) return false;
    double cp = 0;
    int bn = choose_pair_bn(m_tiles, N, batch, b_kmajor, &cp);
    if (!bn) return false;
    *bn_pair = bn;
    if (mode == 2) return true;
    const int64_t tiles1 = m_tiles * ((N + bn1 - 1) / bn1) * batch;
    if (tiles1 < 100 && k_blocks >= 32) return false;      // long K over few tiles: the split-K path of the single-CTA kernel
    if (k_blocks < 8) return false;                        // launch-latency bound: the lighter prologue wins
    // The pair kernel pays for its cluster launch, two cluster barriers and the staged TMA-store epilogue: on problems the single-CTA
    // kernel finishes in ONE wave it measured slower in the UNet step (profiles/r02_launches_step*.csv: 96-tile conv 20.5 vs 16.1 us),
    // on multi-wave problems it wins by up to 1.5x (64x64 640->640 conv: 41 vs 60 us; 8192^3: 1327 vs 842 TF/s).
    if (tiles1 <= num_sms()) return false;
    return cp < 0.9 * single_cost(m_tiles, N, batch, bn1);
}

int launch_pair(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const TcParams& p, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(pairk::tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pairk::P_SMEM);
        if (e != cudaSuccess) return (int)e;
        if (carveout_max()) cudaFuncSetAttribute(pairk::tc_pair_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    const int m_pairs = (p.m_tiles + 1) / 2;
    const int total = m_pairs * p.n_tiles * p.batch;
    const int pairs = std::min(total, num_sms() / 2);
    ProfRec rec{};
    if (g_prof) prof_begin(rec, p, st);
    osb_launch((pairk::tc_pair_kernel), 2 * pairs, pairk::P_THREADS, (size_t)pairk::P_SMEM, st, ma, mb, mc, p);
    if (g_prof) { cudaEventRecord(rec.b, st); g_prof_list.push_back(rec); }
    return launched(1);
}

inline uint32_t next_pow2(uint32_t v) { uint32_t r = 1; while (r < v) r <<= 1; return r; }

}  // namespace

// 0: single-CTA kernel only, 1: cost model (default), 2: pair kernel wherever eligible (tests / A-B runs)
extern "C" void osb_tc_set_pair_mode(int mode) { g_pair_mode = mode; }

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


// ---- W8A8 (kind::i8) entry points ------------------------------------------------------------------------------------
// TMA needs 16-byte global strides: K % 16 (K-major operands), N % 16 (MN-major weights and the 16-byte output vectors)
extern "C" int osb_qu8_tc_gemm_ok(int64_t M, int64_t N, int64_t K, const void* A, const void* B, const void* C)
{
    if (M < 32 || N < 16 || K < 16 || (N % 16) || (K % 16)) return 0;
    if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) return 0;
    static const int off = env_int("OSB_QU8_TC_OFF");
    return get_encode() != nullptr && !off;
}

extern "C" int osb_qu8_tc_conv_ok(int64_t Cin, int64_t Cout, int64_t Ho, int64_t Wo, int kh, int kw, int stride, const void* x, const void* w, const void* y)
{
    if ((Cin % 16) || (Cout % 16) || Ho * Wo < 64 || kh > 7 || kw > 7 || stride < 1 || stride > 2) return 0;
    if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) return 0;
    static const int off = env_int("OSB_QU8_TC_OFF");
    return get_encode() != nullptr && !off;
}

extern "C" int osb_rowsum_u8(const void* x, void* out, int64_t rows, int64_t cols, void* stream)
{
    if (rows * cols == 0) return 0;
    osb_launch((i8k::rowsum_u8_kernel), (unsigned)std::min<int64_t>((rows + 7) / 8, 148 * 8), 256, 0, (cudaStream_t)stream, (const uint8_t*)x, (int32_t*)out, (long long)rows, (long long)cols);
    return launched();
}

extern "C" int osb_colsum_u8(const void* w, void* out, int64_t K, int64_t N, void* stream)
{
    if (K * N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)N * 4, st);
    if (e != cudaSuccess) return (int)e;
    const int64_t gx = (N + 255) / 256;
    const int64_t gy = std::max<int64_t>(1, std::min<int64_t>((K + 63) / 64, (148 * 4 + gx - 1) / gx));
    const int64_t k_per = (K + gy - 1) / gy;
    osb_launch((i8k::colsum_u8_kernel), dim3((unsigned)gx, (unsigned)((K + k_per - 1) / k_per)), 256, 0, st, (const uint8_t*)w, (int32_t*)out, (long long)K, (long long)N, (long long)k_per);
    return launched();
}

extern "C" int osb_pad_sum_u8(const void* x, void* xp, void* psum, int64_t H, int64_t W, int64_t C, int64_t Hp, int64_t Wp, int pad_top, int pad_left, int zx, void* stream)
{
    if (Hp * Wp * C == 0) return 0;
    if (C % 16) return (int)cudaErrorInvalidValue;
    osb_launch((i8k::pad_sum_u8_kernel), (unsigned)std::min<int64_t>((Hp * Wp + 7) / 8, 148 * 16), 256, 0, (cudaStream_t)stream, (const uint8_t*)x, (uint8_t*)xp, (int32_t*)psum,
               (int)H, (int)W, (int)C, (int)Hp, (int)Wp, pad_top, pad_left, zx);
    return launched();
}

static int launch_i8(const CUtensorMap& ma, const CUtensorMap& mb, const i8k::I8Params& p, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(i8k::tc_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, i8k::I8_SMEM);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int total = p.m_tiles * p.n_tiles;
    osb_launch((i8k::tc_i8_kernel), std::min(total, num_sms()), NUM_THREADS, (size_t)i8k::I8_SMEM, st, ma, mb, p);
    return launched(1);
}

// C[M,N] (uint8) = requant((A - zx) (B - zw) + bias): A [M,K] uint8, B [K,N] (bt = 0, ONNX MatMul) or [N,K] (bt = 1); rsum = rowsum_x[M],
// csum = colsum_w[N] (osb_rowsum_u8 / osb_colsum_u8)
extern "C" int osb_qu8_tc_gemm(const void* A, const void* B, void* C, const void* bias, const void* rsum, const void* csum, int64_t M, int64_t N, int64_t K, int bt,
                               int zx, float sx, int zw, float sw, int zy, float sy, void* stream)
{
    CUtensorMap ma, mb;
    const int bn = N <= 64 ? 64 : 128;
    if (!make_map(&ma, A, (uint64_t)K, (uint64_t)M, 1, (uint64_t)K, (uint64_t)M * K, i8k::I8_BLOCK_K, BLOCK_M, 1, 1, CU_TENSOR_MAP_DATA_TYPE_UINT8)) return (int)cudaErrorInvalidValue;
    bool ok = bt ? make_map(&mb, B, (uint64_t)K, (uint64_t)N, 1, (uint64_t)K, (uint64_t)N * K, i8k::I8_BLOCK_K, (uint32_t)bn, 1, 1, CU_TENSOR_MAP_DATA_TYPE_UINT8)
                 : make_map(&mb, B, (uint64_t)N, (uint64_t)K, 1, (uint64_t)N, (uint64_t)N * K, 128, i8k::I8_BLOCK_K, 1, 1, CU_TENSOR_MAP_DATA_TYPE_UINT8);
    if (!ok) return (int)cudaErrorInvalidValue;
    i8k::I8Params p{};
    p.M = (int)M; p.N = (int)N; p.K = (int)K; p.bn = bn;
    p.m_tiles = (int)((M + BLOCK_M - 1) / BLOCK_M); p.n_tiles = (int)((N + bn - 1) / bn);
    p.b_kmajor = bt ? 1 : 0;
    p.taps = 1; p.kw = 1; p.bh = 0; p.bw = 0; p.tiles_x = 1; p.stride = 1;
    p.k_blocks_per_tap = (int)((K + i8k::I8_BLOCK_K - 1) / i8k::I8_BLOCK_K);
    p.rsum = (const int32_t*)rsum; p.csum = (const int32_t*)csum; p.bias = (const int32_t*)bias;
    p.zx = zx; p.zw = zw; p.zy = zy; p.kzz = (int)(K * zx * zw); p.requant = sx * sw / sy;
    p.C = (uint8_t*)C; p.ldc = N;
    return launch_i8(ma, mb, p, (cudaStream_t)stream);
}

// y[Ho,Wo,Cout] (uint8): xp = the zero-point-padded NHWC image [Hp,Wp,Cin] with its per-pixel channel sums psum (osb_pad_sum_u8), w = OHWI
extern "C" int osb_qu8_tc_conv(const void* xp, const void* psum, const void* w, const void* bias, const void* csum, void* y, int64_t Hp, int64_t Wp, int64_t Cin, int64_t Cout,
                               int kh, int kw, int stride, int64_t Ho, int64_t Wo, int zx, float sx, int zw, float sw, int zy, float sy, void* stream)
{
    const uint32_t bw = std::min<uint32_t>(128, next_pow2((uint32_t)Wo)), bh = 128 / bw;
    CUtensorMap ma, mb;
    const int64_t Ktot = (int64_t)kh * kw * Cin;
    const int bn = Cout <= 64 ? 64 : 128;
    if (!make_map(&ma, xp, (uint64_t)Cin, (uint64_t)Wp, (uint64_t)Hp, (uint64_t)Cin, (uint64_t)Wp * Cin, i8k::I8_BLOCK_K, bw * stride, bh * stride, (uint32_t)stride, CU_TENSOR_MAP_DATA_TYPE_UINT8))
        return (int)cudaErrorInvalidValue;
    if (!make_map(&mb, w, (uint64_t)Ktot, (uint64_t)Cout, 1, (uint64_t)Ktot, (uint64_t)Ktot * Cout, i8k::I8_BLOCK_K, (uint32_t)bn, 1, 1, CU_TENSOR_MAP_DATA_TYPE_UINT8)) return (int)cudaErrorInvalidValue;
    i8k::I8Params p{};
    p.bn = bn;
    p.M = (int)(Ho * Wo); p.N = (int)Cout; p.K = (int)Cin;
    p.tiles_x = (int)((Wo + bw - 1) / bw);
    p.m_tiles = p.tiles_x * (int)((Ho + bh - 1) / bh);
    p.n_tiles = (int)((Cout + bn - 1) / bn);
    p.b_kmajor = 1;
    p.taps = kh * kw; p.kw = kw; p.Wo = (int)Wo; p.Ho = (int)Ho; p.bw = (int)bw; p.bh = (int)bh; p.stride = stride; p.Wp = (int)Wp;
    p.k_blocks_per_tap = (int)((Cin + i8k::I8_BLOCK_K - 1) / i8k::I8_BLOCK_K);
    p.rsum = (const int32_t*)psum; p.csum = (const int32_t*)csum; p.bias = (const int32_t*)bias;
    p.zx = zx; p.zw = zw; p.zy = zy; p.kzz = (int)(Ktot * zx * zw); p.requant = sx * sw / sy;
    p.C = (uint8_t*)y; p.ldc = Cout;
    return launch_i8(ma, mb, p, (cudaStream_t)stream);
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

// fp32 problems on the tensor cores (osb_tc_gemm_f32x / osb_tc_conv_f32x below): the launch builders run with this flag set -- operands are
// bf16 triple-split expansions, C / bias / residual are fp32, no CTA-pair tiles, accumulators leave through the fp32 workspace
static thread_local int g_f32x = 0;

// finish a TcParams for the fp32 path; false = the fp32 partial planes do not fit the fixed workspace
static bool f32x_params(TcParams& p, OsbWorkspace* wsp, cudaStream_t st)
{
    p.bf16 = 1; p.f32_out = 1; p.counters = nullptr;
    p.bias2 = nullptr; p.gn_stats = nullptr;
    if ((size_t)p.split_k * p.batch * p.M * p.N * 4 > WS_MAX) p.split_k = 1;
    if ((size_t)p.batch * p.M * p.N * 4 > WS_MAX) return false;
    if (!wsp) wsp = osb_workspace(st, OSB_WS_SPLITK);
    if (!wsp) return false;
    p.ws = wsp->splitk;
    return true;
}

int osb_tc_gemm_launch(const void* A, const void* B, void* C, const void* bias, const void* residual, int64_t batch, int64_t M, int64_t N, int64_t K,
                       int64_t sa, int64_t sb, int64_t sc, int bt, cudaStream_t st, int64_t lda, int64_t ldb, int64_t ldc)
{
    if (g_f32x && batch != 1) return (int)cudaErrorNotSupported;
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
    {
        // CTA-pair kernel (256 x bn tiles, TMA-store epilogue) where its cost model wins: dense C only
        int bnp = 0;
        const int k_blocks = (int)((K + BLOCK_K - 1) / BLOCK_K);
        if (!g_f32x && ldc == N && (batch == 1 || sc == M * N) && use_pair(m_tiles, N, batch, bt != 0, k_blocks, bn, &bnp)) {
            CUtensorMap mc;
            int c_swap = 0;
            bool ok = bt ? make_map_rb(&mb, B, (uint64_t)K, (uint64_t)N, bbatch, (uint64_t)ldb * 2, (uint64_t)(sb ? sb : N * ldb) * 2, BLOCK_K, (uint32_t)(bnp / 2), &b_swap)
                         : make_map_rb(&mb, B, (uint64_t)N, (uint64_t)K, bbatch, (uint64_t)ldb * 2, (uint64_t)(sb ? sb : K * ldb) * 2, 64, BLOCK_K, &b_swap);
            ok = ok && make_map_rb(&mc, C, (uint64_t)N, (uint64_t)M, (uint64_t)batch, (uint64_t)ldc * 2, (uint64_t)(batch > 1 ? sc : M * ldc) * 2, 64, 32, &c_swap) && !c_swap;
            if (ok) {
                TcParams p{};
                p.M = (int)M; p.N = (int)N; p.K = (int)K; p.batch = (int)batch;
                p.bn = bnp; p.m_tiles = (int)m_tiles; p.n_tiles = (int)((N + bnp - 1) / bnp);
                p.b_kmajor = bt ? 1 : 0; p.a_swap = a_swap; p.b_swap = b_swap;
                p.taps = 1; p.kw = 1; p.bh = 0; p.bw = 0; p.tiles_x = 1; p.k_blocks_per_tap = k_blocks; p.stride = 1;
                p.C = (__half*)C; p.bias = (const __half*)bias; p.residual = (const __half*)residual; p.stride_c = sc; p.ldc = ldc; p.split_k = 1;
                return launch_pair(ma, mb, mc, p, st);
            }
        }
    }
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
    p.counters = (p.split_k > 1 && p.N % 8 == 0 && p.ldc % 8 == 0 && (long long)p.m_tiles * p.n_tiles * p.batch <= 2048 && inkernel_reduce() && wsp) ? wsp->splitk_counters : nullptr;
    if (g_f32x && !f32x_params(p, wsp, st)) return (int)cudaErrorNotSupported;
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

// bias2: second per-channel addend or null.  gn_stats != null (and Cout % gn_groups == 0): the kernel adds the per-group (sum, sum of
// squares) of the stored output to gn_stats[2 * groups] (fp64) and sets *gn_done = 1 -- when the chosen decomposition cannot (ragged
// Cout, split-K with a group width that is not a multiple of 4, in-kernel reduce) *gn_done stays 0 and the caller computes them itself.
int osb_tc_conv_launch(const void* x, const void* w, const void* bias, const void* residual, void* y, int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                       int kh, int kw, int stride, int pad_top, int pad_left, int64_t Ho, int64_t Wo, cudaStream_t st,
                       const void* bias2, double* gn_stats, int gn_groups, int* gn_done)
{
    if (gn_done) *gn_done = 0;
    if (gn_stats && (gn_groups < 1 || gn_groups > tcptx::GN_MAX_GROUPS || Cout % gn_groups || Cout % 8)) gn_stats = nullptr;
    const int gn_cpg = gn_stats ? (int)(Cout / gn_groups) : 0;
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
    {
        // CTA-pair kernel: each CTA of the pair takes one 128-pixel box of the same tiling; output through a (Cout, Wo, Ho) store map
        int bnp = 0;
        const int k_blocks = kh * kw * (int)((Cin + BLOCK_K - 1) / BLOCK_K);
        if (!g_f32x && Cout % 8 == 0 && use_pair(m_tiles_, Cout, 1, true, k_blocks, bn, &bnp)) {
            CUtensorMap mc;
            const uint32_t box_w = std::min<uint32_t>(bw, 32), box_h = 32 / box_w;
            bool ok = make_map(&mb, w, (uint64_t)Ktot, (uint64_t)Cout, 1, (uint64_t)Ktot * 2, (uint64_t)Ktot * Cout * 2, BLOCK_K, (uint32_t)(bnp / 2), 1) &&
                      make_map(&mc, y, (uint64_t)Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)Cout * 2, (uint64_t)Wo * Cout * 2, 64, box_w, box_h);
            if (ok) {
                TcParams p{};
                p.bn = bnp;
                p.M = (int)(Ho * Wo); p.N = (int)Cout; p.K = (int)Cin; p.batch = 1;
                p.tiles_x = (int)tiles_x_; p.m_tiles = (int)m_tiles_; p.n_tiles = (int)((Cout + bnp - 1) / bnp);
                p.b_kmajor = 1;
                p.taps = kh * kw; p.kw = kw; p.pad_top = pad_top; p.pad_left = pad_left; p.Wo = (int)Wo; p.Ho = (int)Ho; p.bw = (int)bw; p.bh = (int)bh;
                p.k_blocks_per_tap = (int)((Cin + BLOCK_K - 1) / BLOCK_K);
                p.stride = stride;
                p.C = (__half*)y; p.bias = (const __half*)bias; p.residual = (const __half*)residual; p.stride_c = 0; p.ldc = Cout; p.split_k = 1;
                p.bias2 = (const __half*)bias2; p.gn_stats = gn_stats; p.gn_cpg = gn_cpg; p.gn_groups = gn_groups;
                if (gn_stats && gn_done) *gn_done = 1;
                return launch_pair(ma, mb, mc, p, st);
            }
        }
    }
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
    p.counters = (p.split_k > 1 && p.N % 8 == 0 && p.ldc % 8 == 0 && (long long)p.m_tiles * p.n_tiles * p.batch <= 2048 && inkernel_reduce() && wsp) ? wsp->splitk_counters : nullptr;
    p.short_k = short_k_hint(p.taps * p.k_blocks_per_tap, p.split_k);
    p.bias2 = (const __half*)bias2;
    // statistics: the tile epilogue (unsplit) or the reduce kernel (split-K; needs 4 consecutive columns inside one group)
    const bool stats_ok = gn_stats && (p.split_k == 1 || p.counters || gn_cpg % 4 == 0);     // final epilogue (unsplit / last finisher) or the reduce kernel
    if (stats_ok) { p.gn_stats = gn_stats; p.gn_cpg = gn_cpg; p.gn_groups = gn_groups; if (gn_done) *gn_done = 1; }
    { static const int dbg = env_int("OSB_GN_DEBUG"); p.gn_debug = dbg; }
    if (g_f32x) {
        if (!f32x_params(p, wsp, st)) return (int)cudaErrorNotSupported;
        p.short_k = short_k_hint(p.taps * p.k_blocks_per_tap, p.split_k);
    }
    return launch(ma, mb, p, st);
}

// ---- fp32 GEMM / conv on the tensor cores: bf16 triple split ------------------------------------------------------------------------
// x = h + m + l with h = bf16(x), m = bf16(x - h), l = bf16(x - h - m) carries 24 mantissa bits; a.b ~= ah.bh + ah.bm + am.bh + ah.bl + al.bh + am.bm
// (the dropped terms are below 2^-24 relative).  The six products are ONE tensor-core contraction over a 6x longer K: the caller expands
// A to [h|h|m|h|l|m] and B to [h|m|h|l|h|m] along K (osb_bf16x3_expand_*), products are exact in fp32 and accumulate in the fp32 TMEM
// accumulator.  A, B: bf16 expansions (K6 = 6 K); C, bias, residual: fp32; C dense [M][N].  cudaErrorNotSupported: run the CUDA-core kernel.
// shape predicates of the fp32 tensor-core path (before the caller spends time expanding operands)
extern "C" int osb_tc_gemm_f32x_ok(int64_t M, int64_t N, int64_t K)
{
    static const bool on = [] { const char* e = getenv("OSB_F32_TC"); return !(e && e[0] == '0'); }();
    return on && M >= 32 && N >= 8 && N % 8 == 0 && K >= 8 && (6 * K) % 8 == 0 && (size_t)M * N * 4 <= WS_MAX && get_encode() != nullptr ? 1 : 0;
}
extern "C" int osb_tc_conv_f32x_ok(int64_t H, int64_t W, int64_t Cin, int64_t Cout, int kh, int kw, int stride, int64_t Ho, int64_t Wo)
{
    static const bool on = [] { const char* e = getenv("OSB_F32_TC"); return !(e && e[0] == '0'); }();
    return on && (6 * Cin) % 8 == 0 && 6 * Cin >= 16 && stride >= 1 && stride <= 2 && H * W >= 64 && kh <= 7 && kw <= 7 && (size_t)Ho * Wo * Cout * 4 <= WS_MAX &&
           get_encode() != nullptr ? 1 : 0;
}

extern "C" int osb_tc_gemm_f32x(const void* A6, const void* B6, void* C, const void* bias, const void* residual, int64_t M, int64_t N, int64_t K6, int bt, void* stream)
{
    const int64_t ldb = bt ? K6 : N;
    // (ldc is a float pitch here: only 16-byte pointer alignment of C matters to the workspace reduce)
    if (!osb_tc_gemm_ok(M, N, K6, bt, A6, B6, nullptr, 0, 0, 0, K6, ldb, 8) || (N % 8)) return (int)cudaErrorNotSupported;
    g_f32x = 1;
    int r = osb_tc_gemm_launch(A6, B6, C, bias, residual, 1, M, N, K6, 0, 0, M * N, bt, (cudaStream_t)stream, K6, ldb, N);
    g_f32x = 0;
    return r;
}

extern "C" int osb_tc_conv_f32x(const void* x6, const void* w6, const void* bias, const void* residual, void* y, int64_t H, int64_t W, int64_t Cin6, int64_t Cout,
                     int kh, int kw, int stride, int pad_top, int pad_left, int64_t Ho, int64_t Wo, void* stream)
{
    if (!osb_tc_conv_ok(H, W, Cin6, Cout, kh, kw, stride, x6, w6, nullptr)) return (int)cudaErrorNotSupported;
    g_f32x = 1;
    int r = osb_tc_conv_launch(x6, w6, bias, residual, y, H, W, Cin6, Cout, kh, kw, stride, pad_top, pad_left, Ho, Wo, (cudaStream_t)stream, nullptr, nullptr, 0, nullptr);
    g_f32x = 0;
    return r;
}
