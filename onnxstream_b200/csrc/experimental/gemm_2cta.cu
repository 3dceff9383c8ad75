// gemm_2cta.cu -- EXPERIMENTAL, NOT PART OF THE PRODUCT BUILD (onnxstream_b200/build.py does not compile it).
//
// Round-2 groundwork for DESIGN.md section 6 item 1: a `tcgen05.mma.cta_group::2` GEMM, C[M,N] = A[M,K] * B[N,K]^T (fp16 in, fp32
// accumulate, fp16 out), with a CTA PAIR working on one 256 x 256 output tile.  Why: round 1 measured that a 128 x 128 tile per CTA
// moves (128 + 128) * 128 B per k-block for 128 * 128 * 64 MACs and that the chip-wide L2->SM rate (~6300 B/clk) caps such a kernel
// near 900 TF/s (885 measured on 8192^3).  In pair mode each CTA still loads 16 KiB of A (its 128 rows) + 16 KiB of B (its 128 of the
// 256 columns) per k-block but the pair retires 256 x 256 x 64 MACs: half the L2->SM bytes per FLOP.
//
// Structure (follows the "canonical Blackwell GEMM" of /opt/skills/guides/blackwell_cuda_programming.md):
//   cluster (2,1,1); warp 0 = TMA producer in BOTH CTAs (cp.async.bulk.tensor...cta_group::2 signalling the LEADER's full barrier),
//   warp 1 = MMA issuer in the LEADER only (UMMA M = 256, N = 256, K = 16; commit multicast to both CTAs' barriers),
//   warps 2..5 = epilogue in both CTAs (each CTA owns the 128 accumulator rows that live in its own TMEM).
// Status: compiles for sm_100a (ptxas accepts every instruction form); NEVER RUN -- the round-1 GPU budget was spent.  The check
// harness is scripts/exp_gemm_2cta_check.py.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC
//        -o build/libexp_gemm_2cta.so onnxstream_b200/csrc/experimental/gemm_2cta.cu -lcuda  (the script does it).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace {

constexpr int TILE_M = 256, TILE_N = 256, BLOCK_K = 64, UMMA_K = 16;
constexpr int CTA_M = 128;                       // accumulator rows per CTA
constexpr int CTA_N_LOAD = TILE_N / 2;           // B columns each CTA loads
constexpr int STAGES = 6;
constexpr int A_BYTES = CTA_M * BLOCK_K * 2;     // 16 KiB
constexpr int B_BYTES = CTA_N_LOAD * BLOCK_K * 2;   // 16 KiB
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * TILE_N;   // 512: the whole TMEM of each SM
constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 + 512;
constexpr int THREADS = 192;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;      // clears the CTA-pair peer bit of a shared::cluster address: "the leader's copy"

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t addr = smem_u32(bar), done = 0;
    long long t0 = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000LL) { printf("gemm_2cta: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
// local arrive + expected transaction bytes (leader's producer)
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta)
{
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
// 2-D tile load executed by both CTAs of the pair; the transaction bytes are credited to the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_dst), "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once all previously issued MMAs retired) on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
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
// K-major SWIZZLE_128B operand descriptor (version 1): 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)(16 >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc_f16_m256_n256()
{
    uint32_t d = 0;
    d |= 1u << 4;                          // D = f32
    d |= (uint32_t)(TILE_N >> 3) << 17;    // N
    d |= (uint32_t)(TILE_M >> 4) << 24;    // M = 256 (pair)
    return d;                              // A, B = f16, both K-major
}

struct Params { int M, N, K; __half* C; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Params p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_BYTES;
    uint64_t* bars = (uint64_t*)(smem + STAGES * (A_BYTES + B_BYTES));
    uint64_t* full = bars;                         // [STAGES]   (used in the leader; both CTAs' loads complete_tx on the leader's copy)
    uint64_t* empty = bars + STAGES;               // [STAGES]   (one per CTA: its own producer waits on it)
    uint64_t* acc_full = bars + 2 * STAGES;        // [ACC_STAGES] (one per CTA: its own epilogue waits on it)
    uint64_t* acc_empty = acc_full + ACC_STAGES;   // [ACC_STAGES] (leader: counts the epilogue threads of BOTH CTAs)
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }     // full: one arrival per CTA of the pair
        for (int i = 0; i < ACC_STAGES; i++) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * 4 * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        // both CTAs execute the pair allocation (one warp each)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync();          // barrier inits and TMEM allocation of BOTH CTAs are visible before any cross-CTA signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_m = p.M / TILE_M, tiles_n = p.N / TILE_N;
    const int total_tiles = tiles_m * tiles_n;
    const int k_blocks = p.K / BLOCK_K;
    const int pair = blockIdx.x >> 1, pairs = gridDim.x >> 1;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        const uint32_t sa0 = smem_u32(smem_a), sb0 = smem_u32(smem_b);
        int stage = 0; uint32_t phase = 0;
        for (int tile = pair; tile < total_tiles; tile += pairs) {
            const int mt = tile % tiles_m, nt = tile / tiles_m;
            const int m0 = mt * TILE_M + (int)rank * CTA_M;          // this CTA's 128 rows of A
            const int n0 = nt * TILE_N + (int)rank * CTA_N_LOAD;     // this CTA's 128 of the 256 columns of B
            for (int kb = 0; kb < k_blocks; kb++) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(&full[stage], 2 * (A_BYTES + B_BYTES));      // both CTAs' bytes land on this barrier
                    tma_load_2d_2sm(sa0 + stage * A_BYTES, &map_a, &full[stage], kb * BLOCK_K, m0);
                    tma_load_2d_2sm(sb0 + stage * B_BYTES, &map_b, &full[stage], kb * BLOCK_K, n0);
                    if (!leader) mbar_arrive_cluster(&full[stage], 0);                      // second arrival, no bytes of its own to announce
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            const uint32_t idesc = idesc_f16_m256_n256();
            const uint64_t adesc0 = smem_desc(smem_u32(smem_a)), bdesc0 = smem_desc(smem_u32(smem_b));
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = pair; tile < total_tiles; tile += pairs) {
                mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TILE_N);
                for (int kb = 0; kb < k_blocks; kb++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = adesc0 + (uint64_t)(stage * (A_BYTES >> 4));
                    const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (B_BYTES >> 4));
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; k++)
                            umma_f16_2sm(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                        umma_commit_2sm(&empty[stage], 0b11);                        // frees the slot in BOTH CTAs
                        if (kb == k_blocks - 1) umma_commit_2sm(&acc_full[acc], 0b11);   // wakes the epilogue of BOTH CTAs
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5, both CTAs): this CTA's 128 rows x 256 columns =====================
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = pair; tile < total_tiles; tile += pairs) {
            const int mt = tile % tiles_m, nt = tile / tiles_m;
            const long long row = (long long)mt * TILE_M + (long long)rank * CTA_M + q * 32 + lane;
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            __half* crow = p.C + row * p.N + (long long)nt * TILE_N;
#pragma unroll 1
            for (int c = 0; c < TILE_N; c += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TILE_N + c), v);
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    __half h[8];
#pragma unroll
                    for (int t = 0; t < 8; t++) h[t] = __float2half_rn(__uint_as_float(v[j + t]));
                    *reinterpret_cast<uint4*>(crow + c + j) = *reinterpret_cast<const uint4*>(h);
                }
            }
            tc_fence_before();
            mbar_arrive_cluster(&acc_empty[acc], 0);     // every epilogue thread of the pair reports to the leader's barrier
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    cluster_sync();          // nobody frees TMEM (or exits, taking its shared memory away) while the peer may still signal it
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)ptr;
    }
    return fn;
}

bool make_map_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows)
{
    auto enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = { inner, rows };
    cuuint64_t strides[1] = { inner * 2 };
    cuuint32_t box[2] = { box_inner, box_rows };
    cuuint32_t estr[2] = { 1, 1 };
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// C[M,N] = A[M,K] * B[N,K]^T, fp16 row-major operands, M % 256 == 0, N % 256 == 0, K % 64 == 0.  Returns a cudaError_t as int.
extern "C" int osb_exp_gemm_2cta(const void* A, const void* B, void* C, long long M, long long N, long long K, void* stream)
{
    if (M % TILE_M || N % TILE_N || K % BLOCK_K || M <= 0 || N <= 0 || K <= 0) return (int)cudaErrorInvalidValue;
    CUtensorMap ma, mb;
    if (!make_map_2d(&ma, A, (uint64_t)K, (uint64_t)M, BLOCK_K, CTA_M)) return (int)cudaErrorInvalidValue;
    if (!make_map_2d(&mb, B, (uint64_t)K, (uint64_t)N, BLOCK_K, CTA_N_LOAD)) return (int)cudaErrorInvalidValue;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(gemm_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long tiles = (M / TILE_M) * (N / TILE_N);
    int pairs = (int)(tiles < sms / 2 ? tiles : sms / 2);
    Params p{ (int)M, (int)N, (int)K, (__half*)C };
    gemm_2cta_kernel<<<dim3(2 * pairs), dim3(THREADS), SMEM_BYTES, (cudaStream_t)stream>>>(ma, mb, p);
    return (int)cudaGetLastError();
}
