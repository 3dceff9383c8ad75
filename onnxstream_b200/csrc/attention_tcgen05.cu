// attention_tcgen05.cu -- fused flash-style attention for sm_100a: softmax(Q K^T * scale) V in ONE kernel, score tile
// in TMEM, never in HBM.  Replaces the reference's sliced AttentionFusedOps loop (src/onnxstream.cpp:6696-6929: per head and
// per Q-slice a MatMul, a Mul, a Softmax and a MatMul, each through XNNPACK with the [Tq/parts, Tk] score tile round-tripping
// through two aux buffers, src/onnxstream.cpp:6798-6799).
//
// One CTA per (head, 128-query tile), 192 threads:
//   warp 0      TMA producer: Q tile once, then a (K tile, V tile) pair per 128 keys into a 2-stage ring
//   warp 1      MMA issuer: S[j] = Q K_j^T (tcgen05.mma, 128x128x64, S double-buffered in TMEM) and O += P_j V_j (128x64x128)
//   warps 2..5  softmax: thread r owns query row r -- tcgen05.ld its S row, online max / sum in fp32 (exp2 with the scale
//               folded in), rescales its O row in TMEM when the running max moves (tcgen05.ld / tcgen05.st), writes P as fp16
//               into 128B-swizzled shared memory (the A operand of the second MMA), finally O / sum -> fp16 -> global
// Q, K, V are read in place from the [T, heads*d] projection buffers through strided tensor maps and O is written in the
// merged [T, heads*d] layout, so the exported graph's head split / merge costs nothing.  d <= 64, d % 8 == 0.

#include "common.cuh"
#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <mutex>

namespace {

constexpr int BQ = 128;                 // queries per CTA
constexpr int BKV = 128;                // keys per tile
constexpr int BD = 64;                  // head dim padded to one 128-byte swizzle row
constexpr int KV_STAGES = 4;            // K/V ring: a stage is freed by PV(j) and needed again by QK(j + KV_STAGES); with 2 stages the TMA round trip was exposed on every tile
constexpr int Q_BYTES = BQ * BD * 2;    // 16 KiB
constexpr int K_BYTES = BKV * BD * 2;   // 16 KiB
constexpr int V_BYTES = BKV * BD * 2;   // 16 KiB (128 key rows x 64 columns)
constexpr int P_BYTES = BQ * BKV * 2;   // 32 KiB (two 64-key k-blocks)
constexpr int XCHG_BYTES = 2 * 2 * BQ * 4;   // half-row maxima / sums exchanged between the two warps of a row quadrant (double-buffered)
constexpr int FA_SMEM = Q_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + 2 * P_BYTES + 1024 + 256 + XCHG_BYTES;
constexpr int FA_THREADS = 320;          // warp 0 TMA, warp 1 MMA, warps 2..9 softmax (two per TMEM lane quadrant)
constexpr int TMEM_COLS_FA = 512;       // S0 [0,128) S1 [128,256) O [256,320)
constexpr int O_COL = 256;

struct FaParams {
    int T, Tk, d, heads;
    int q_tiles, kv_tiles;
    float scale_log2;        // scale * log2(e)
    float tau;               // lazy-rescaling threshold in log2 units (0 = exact running maximum)
    __half* out;             // [T, ldo] merged layout, head h at column h*d
    long long ldo;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t addr = smem_u32(bar), done = 0;
    long long t0 = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000LL) { printf("flash_attention_kernel: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 2^x on the SFU, flush-to-zero: one MUFU.EX2 (exp2f() adds a range check and two scaling multiplies for denormal results,
// which a probability that is about to be rounded to fp16 does not need).  ex2(-inf) = +0.
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// one lane of a converged warp, chosen by the hardware: ptxas knows the guarded region runs on exactly one lane
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
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
// issue only: several loads can be in flight before one tcgen05.wait::ld
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r)
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
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr),
          "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;   // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint32_t idesc_f16(int n, int b_mn_major)
{
    return (1u << 4) | ((uint32_t)(b_mn_major ? 1 : 0) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(FA_THREADS, 1)
flash_attention_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                       const FaParams p)
{
    osb_pdl_trigger_entry();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + Q_BYTES;
    uint8_t* sV = sK + KV_STAGES * K_BYTES;
    uint8_t* sP = sV + KV_STAGES * V_BYTES;
    uint64_t* bars = (uint64_t*)(sP + 2 * P_BYTES);   // P is double-buffered: softmax of tile j+1 never waits for the PV MMA of tile j
    uint64_t* q_full = bars;                 // [1]
    uint64_t* kv_full = bars + 1;                      // [KV_STAGES]
    uint64_t* kv_empty = kv_full + KV_STAGES;          // [KV_STAGES]
    uint64_t* s_full = kv_empty + KV_STAGES;           // [2]
    uint64_t* s_empty = s_full + 2;                    // [2]
    uint64_t* p_full = s_empty + 2;                    // [2]
    uint64_t* pv_done = p_full + 2;                    // [2]
    uint32_t* tmem_slot = (uint32_t*)(pv_done + 2);
    float* xchg = (float*)((uint8_t*)bars + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, head = blockIdx.y;
    const int q0 = qt * BQ;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < KV_STAGES; i++) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 256); mbar_init(&p_full[i], 256); mbar_init(&pv_done[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS_FA) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    osb_pdl_wait();

    const int n_kv = p.kv_tiles;

    // Producer and MMA warps run warp-uniform control flow and issue from elect.sync-guarded regions: inside an `if (lane == 0)`
    // region ptxas wraps every UTCHMMA / UTMALDG in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop (measured: the issue loop, not
    // the tensor pipe or the loads, bounded the kernel).
    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(q_full, Q_BYTES);
            tma_load_3d(sQ, &map_q, q_full, 0, head, q0);
        }
        __syncwarp();
        for (int j = 0; j < n_kv; j++) {
            const int st = j % KV_STAGES;
            mbar_wait(&kv_empty[st], ((j / KV_STAGES) & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&kv_full[st], K_BYTES + V_BYTES);
                tma_load_3d(sK + st * K_BYTES, &map_k, &kv_full[st], 0, head, j * BKV);
                tma_load_3d(sV + st * V_BYTES, &map_v, &kv_full[st], 0, head, j * BKV);
                tma_load_3d(sV + st * V_BYTES + V_BYTES / 2, &map_v, &kv_full[st], 0, head, j * BKV + 64);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        const uint32_t idesc_s = idesc_f16(BKV, 0);     // S = Q K^T: A K-major, B K-major, N = 128 keys
        const uint32_t idesc_o = idesc_f16(BD, 1);      // O += P V : A K-major (P), B MN-major (V rows = keys), N = 64
        mbar_wait(q_full, 0);
        tc_fence_after();
        // descriptor templates; only the 14-bit start-address field (16-byte units) moves
        const uint64_t qdesc0 = smem_desc(smem_u32(sQ), 16, 1024);
        const uint64_t kdesc0 = smem_desc(smem_u32(sK), 16, 1024);
        const uint64_t pdesc0 = smem_desc(smem_u32(sP), 16, 1024);
        const uint64_t vdesc0 = smem_desc(smem_u32(sV), V_BYTES, 1024);
        for (int j = 0; j <= n_kv; j++) {
            if (j < n_kv) {
                const int st = j & 1, ks = j % KV_STAGES;
                mbar_wait(&kv_full[ks], (j / KV_STAGES) & 1);
                mbar_wait(&s_empty[st], ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint64_t kdesc = kdesc0 + (uint64_t)(ks * (K_BYTES >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BD / 16; k++)
                        umma_f16(tmem_base + (uint32_t)(st * BKV), qdesc0 + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), idesc_s, k != 0);
                    umma_commit(&s_full[st]);
                }
                __syncwarp();
            }
            if (j >= 1) {
                const int jj = j - 1, st = jj & 1;
                mbar_wait(&p_full[st], (jj >> 1) & 1);
                tc_fence_after();
                const int ks = jj % KV_STAGES;
                const uint64_t vdesc = vdesc0 + (uint64_t)(ks * (V_BYTES >> 4));
                const uint64_t pdesc = pdesc0 + (uint64_t)(st * (P_BYTES >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BKV / 16; k++) {
                        // P: two 64-key k-blocks 16 KiB apart, 32 B per 16-key step inside a block; V: 16 key rows = 2048 B per step
                        umma_f16(tmem_base + O_COL, pdesc + (uint64_t)(((k >> 2) * (P_BYTES / 2) + (k & 3) * 32) >> 4), vdesc + (uint64_t)(k * (2048 >> 4)),
                                 idesc_o, (jj != 0 || k != 0) ? 1u : 0u);
                    }
                    umma_commit(&kv_empty[ks]);
                    umma_commit(&pv_done[st]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== softmax / correction / epilogue (warps 2..9) =====================
        // Two warps per TMEM lane quadrant: warp w and w + 4 own the same 32 query rows and split the 128 keys of a tile (and the
        // 64 columns of O) in halves, so every SM sub-partition has two softmax warps to interleave (one warp alone left the SFU
        // idle while it waited on TMEM loads / fences).  Per tile the pair exchanges its half-row maxima through shared memory
        // (one 64-thread named barrier); the row sums stay per half (same running maximum) and are added once at the end.
        const int qd = warp & 3;
        const int half = (warp - 2) >> 2;             // 0: keys [0,64) of the tile, O columns [0,32); 1: the other halves
        const int row = qd * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        const int bar_id = 1 + qd;                    // named barrier of this pair (0 is __syncthreads)
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_kv; j++) {
            int st = j & 1;
            mbar_wait(&s_full[st], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t s_addr = tmem_base + lane_addr + (uint32_t)(st * BKV + half * 64);
            const int key0 = j * BKV + half * 64;
            uint32_t sv[64];
            tmem_ld32_nowait(s_addr, sv);
            tmem_ld32_nowait(s_addr + 32, sv + 32);
            tmem_ld_wait();
            const bool tail = key0 + 64 > p.Tk;           // only the last tile has padding keys (K rows zero-filled by TMA)
            // Two uniform code paths: the full-tile path carries no per-element predicates (measured: the predicated single path
            // spent ~20 issue slots per score, the SFU needs 8).
            float mt = -INFINITY;
            if (!tail) {
#pragma unroll
                for (int t = 0; t < 64; t++) mt = fmaxf(mt, __uint_as_float(sv[t]));
            } else {
#pragma unroll
                for (int t = 0; t < 64; t++) if (key0 + t < p.Tk) mt = fmaxf(mt, __uint_as_float(sv[t]));
            }
            // exchange the half-row maxima (double-buffered by tile parity: one barrier per tile is enough)
            float* xw = xchg + ((j & 1) * 2 + half) * BQ;
            float* xr = xchg + ((j & 1) * 2 + (half ^ 1)) * BQ;
            xw[row] = mt;
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            mt = fmaxf(mt, xr[row]);
            // Lazy rescaling: the stale maximum is kept while the new one exceeds it by at most tau (log2 units) -- P <= 2^tau stays well
            // inside fp16, O and l accumulate in fp32 -- so most tiles skip the O rescale and with it the wait on the previous PV MMA.
            // tau = 0 is the exact running maximum (the default until the threshold is validated on hardware: OSB_FLASH_TAU).
            const float m_cand = fmaxf(m_run, mt * p.scale_log2);     // scale_log2 > 0
            const float m_new = (m_cand - m_run > p.tau) ? m_cand : m_run;
            const float alpha = ex2_approx(m_run - m_new);            // 0 on the first tile (m_run = -inf)
            const float neg_m = -m_new;
            // p = 2^(s*scale*log2e - m_new): one FFMA + one MUFU.EX2 per score, packed to fp16 as produced
            uint32_t pk[32];
            float lsum0 = 0.f, lsum1 = 0.f;
            if (!tail) {
#pragma unroll
                for (int t = 0; t < 64; t += 2) {
                    float p0 = ex2_approx(fmaf(__uint_as_float(sv[t]), p.scale_log2, neg_m));
                    float p1 = ex2_approx(fmaf(__uint_as_float(sv[t + 1]), p.scale_log2, neg_m));
                    lsum0 += p0; lsum1 += p1;
                    __half2 h2 = __floats2half2_rn(p0, p1);
                    pk[t >> 1] = *reinterpret_cast<uint32_t*>(&h2);
                }
            } else {
#pragma unroll
                for (int t = 0; t < 64; t += 2) {
                    float p0 = key0 + t < p.Tk ? ex2_approx(fmaf(__uint_as_float(sv[t]), p.scale_log2, neg_m)) : 0.f;
                    float p1 = key0 + t + 1 < p.Tk ? ex2_approx(fmaf(__uint_as_float(sv[t + 1]), p.scale_log2, neg_m)) : 0.f;
                    lsum0 += p0; lsum1 += p1;
                    __half2 h2 = __floats2half2_rn(p0, p1);
                    pk[t >> 1] = *reinterpret_cast<uint32_t*>(&h2);
                }
            }
            const float lsum = lsum0 + lsum1;
            tc_fence_before();
            mbar_arrive(&s_empty[st]);                   // S[st] may be overwritten by QK^T of tile j+2
            l_run = l_run * alpha + lsum;                // partial sum over this warp's key halves
            m_run = m_new;
            // O may only be rescaled once PV(j-1) has retired; the P buffer of this parity is free once PV(j-2) has (the tensor pipe
            // retires in order).  Rescaling is skipped when no row of the warp moved its maximum -- the common case after a few
            // tiles, and then the softmax of tile j never waits for the MMA of tile j-1 (both warps of a pair see the same alpha).
            if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
                mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
                tc_fence_after();
                uint32_t o[32];
                tmem_ld32(tmem_base + lane_addr + O_COL + half * 32, o);
#pragma unroll
                for (int t = 0; t < 32; t++) o[t] = __float_as_uint(__uint_as_float(o[t]) * alpha);
                tmem_st32(tmem_base + lane_addr + O_COL + half * 32, o);
            } else if (j > 1) {
                mbar_wait(&pv_done[j & 1], ((j - 2) >> 1) & 1);
            }
            // P half-row -> shared memory, K-major SWIZZLE_128B: k-block = this warp's key half, 16-byte chunk index XOR (row % 8)
            {
                const uint32_t prow = smem_u32(sP) + st * P_BYTES + row * 128 + half * (P_BYTES / 2);
#pragma unroll
                for (int c8 = 0; c8 < 8; c8++) {
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + ((c8 ^ (row & 7)) << 4)),
                                 "r"(pk[c8 * 4]), "r"(pk[c8 * 4 + 1]), "r"(pk[c8 * 4 + 2]), "r"(pk[c8 * 4 + 3]) : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the async proxy (UMMA)
            tc_fence_before();
            mbar_arrive(&p_full[st]);
        }
        // epilogue: O / l -> fp16 -> out[q, head*d + c]; l = sum of the pair's partial sums (same running maximum)
        {
            float* xw = xchg + ((n_kv & 1) * 2 + half) * BQ;
            float* xr = xchg + ((n_kv & 1) * 2 + (half ^ 1)) * BQ;
            xw[row] = l_run;
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            l_run += xr[row];
        }
        mbar_wait(&pv_done[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1);
        tc_fence_after();
        const int qrow = q0 + row;
        const float inv = 1.f / l_run;
        __half* orow = p.out + (long long)qrow * p.ldo + (long long)head * p.d;
        const int c = half * 32;
        if (c < p.d) {
            uint32_t o[32];
            tmem_ld32(tmem_base + lane_addr + O_COL + c, o);
            if (qrow < p.T) {
#pragma unroll
                for (int t = 0; t < 32; t += 8) {
                    if (c + t >= p.d) break;       // d % 8 == 0
                    Vec<__half, 8> w;
#pragma unroll
                    for (int u = 0; u < 8; u++) w.v[u] = __float2half_rn(__uint_as_float(o[t + u]) * inv);
                    store_vec<__half, 8>(orow + c + t, w);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS_FA) : "memory");
    }
}

PFN_cuTensorMapEncodeTiled_v12000 fa_encode()
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

// [rows, heads*d] projection viewed as (d, heads, rows): strides (d*2, ld*2) bytes; box (64, 1, box_rows)
bool head_map(CUtensorMap* map, const void* base, int d, int heads, int64_t rows, int64_t ld, uint32_t box_rows)
{
    auto enc = fa_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = { (cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)rows };
    cuuint64_t strides[2] = { (cuuint64_t)d * 2, (cuuint64_t)ld * 2 };
    cuuint32_t box[3] = { 64, 1, box_rows };
    cuuint32_t estr[3] = { 1, 1, 1 };
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

extern "C" int osb_flash_attention_ok(int64_t T, int64_t Tk, int64_t d, int dtype)
{
    return dtype == OSB_F16 && d >= 8 && d <= 64 && d % 8 == 0 && T >= 64 && Tk >= 1 && fa_encode() != nullptr;
}

// q [T, heads*d] (row stride ldq), k / v [Tk, heads*d] (row strides ldk / ldv), out [T, heads*d] (row stride ldo); fp16.
extern "C" int osb_flash_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out, int64_t ldo,
                                   int64_t heads, int64_t T, int64_t Tk, int64_t d, float scale, void* stream)
{
    if (heads * T == 0) return 0;
    if (d > 64 || d % 8 || (ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return (int)cudaErrorInvalidValue;
    if ((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) != 0) return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap mq, mk, mv;
    if (!head_map(&mq, q, (int)d, (int)heads, T, ldq, BQ)) return (int)cudaErrorInvalidValue;
    if (!head_map(&mk, k, (int)d, (int)heads, Tk, ldk, BKV)) return (int)cudaErrorInvalidValue;
    if (!head_map(&mv, v, (int)d, (int)heads, Tk, ldv, 64)) return (int)cudaErrorInvalidValue;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(flash_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    FaParams p{};
    p.T = (int)T; p.Tk = (int)Tk; p.d = (int)d; p.heads = (int)heads;
    p.q_tiles = (int)((T + BQ - 1) / BQ); p.kv_tiles = (int)((Tk + BKV - 1) / BKV);
    p.scale_log2 = scale * 1.4426950408889634f;
    static const float tau_env = [] { const char* e = getenv("OSB_FLASH_TAU"); float v = e ? (float)atof(e) : 0.f; return v < 0.f ? 0.f : (v > 12.f ? 12.f : v); }();
    p.tau = tau_env;
    p.out = (__half*)out; p.ldo = ldo;
    dim3 grid((unsigned)p.q_tiles, (unsigned)heads);
    osb_launch((flash_attention_kernel), grid, FA_THREADS, (size_t)FA_SMEM, st, mq, mk, mv, p);
    return launched(1);
}
