// gemm_pair.cuh -- the CTA-PAIR variant of the tensor-core GEMM / implicit-GEMM conv (included by gemm_tcgen05.cu, same TcParams).
//
// Two CTAs of a (2,1,1) cluster -- two SMs of one TPC -- work on ONE 256 x bn output tile with `tcgen05.mma.cta_group::2`:
//   * each CTA loads its own 128 rows of A (16 KiB per k-block) and HALF of the B tile (bn/2 columns), so a pair moves
//     2 * (16 KiB + bn * 64 B) per k-block for 256 x bn x 64 MACs: half the L2->SM bytes per FLOP of the 128 x 128 single-CTA tile,
//     which is what bounded that kernel on large problems (885 TF/s on 8192^3; the pair skeleton reached 1433);
//   * the leader CTA's MMA warp issues UMMA M = 256, N = bn, K = 16; each CTA's TMEM receives its own 128 accumulator rows;
//   * `tcgen05.commit ... multicast::cluster` frees the shared-memory slot / publishes the accumulator in BOTH CTAs;
//   * epilogue: TMEM -> registers (+bias, +residual, fp16) -> 128B-swizzled shared staging -> `cp.async.bulk.tensor` STORE
//     (32-row x 64-column boxes, one per epilogue warp and panel, clipped by TMA at the M / N edges).
// Replaces the same reference code as gemm_tcgen05.cu (XnnPack::convolution / matrix_multiply for fp16, src/onnxstream.cpp:929-1534;
// CublasOps::OpFullyConnected::run, src/onnxstream.cpp:308-352) on the shapes where the pair tile wins (host heuristic below).

namespace pairk {

constexpr int P_STAGES = 5;
constexpr int P_A_BYTES = 128 * 64 * 2;          // 16 KiB: this CTA's 128 rows x 64 k
constexpr int P_B_BYTES = 128 * 64 * 2;          // up to 128 of the tile's 256 columns x 64 k
constexpr int P_STG_WARP = 2 * 4096;             // per epilogue warp: two 32-row x 128-byte staging panels
constexpr int P_SMEM = P_STAGES * (P_A_BYTES + P_B_BYTES) + 4 * P_STG_WARP + 1024 /*align*/ + 512 /*barriers*/ + GN_SMEM_BYTES;
constexpr int P_THREADS = 192;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;      // clears the pair-peer bit of a shared::cluster address: "the leader's copy"

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta)
{
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
// tile load executed by both CTAs of the pair; the transaction bytes are credited to the LEADER's barrier
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_dst), "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
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
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
tc_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_c, const TcParams p)
{
    osb_pdl_trigger_entry();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + P_STAGES * P_A_BYTES;
    uint8_t* smem_stg = smem + P_STAGES * (P_A_BYTES + P_B_BYTES);                 // 1024-byte aligned: the store swizzle is address-based
    uint64_t* bars = (uint64_t*)(smem_stg + 4 * P_STG_WARP);
    uint64_t* full = bars;                         // [P_STAGES]  used in the leader: both CTAs' loads complete_tx on the leader's copy
    uint64_t* empty = bars + P_STAGES;             // [P_STAGES]  one per CTA: its own producer waits on it
    uint64_t* acc_full = bars + 2 * P_STAGES;      // [2]         one per CTA: its own epilogue waits on it
    uint64_t* acc_empty = acc_full + 2;            // [2]         leader: one arrival per epilogue warp of BOTH CTAs
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
    uint8_t* gn_buf = (uint8_t*)bars + 512;                 // [4][2048] warp-private chunk transposition buffers (GroupNorm statistics)
    float* gn_acc = (float*)(gn_buf + 4 * 2048);            // [2 * GN_MAX_GROUPS]
    if (p.gn_stats && threadIdx.x < 2 * GN_MAX_GROUPS) gn_acc[threadIdx.x] = 0.f;   // ordered before use by the cluster barrier below

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;
    const uint32_t tmem_cols = p.bn > 128 ? 512u : 256u;      // two accumulator stages of bn columns, power of two

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < P_STAGES; i++) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }     // full: one arrival per CTA of the pair
        for (int i = 0; i < 2; i++) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        // both CTAs execute the pair allocation (one warp each)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();      // barrier inits and TMEM allocation of BOTH CTAs are visible before any cross-CTA signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    osb_pdl_wait();

    const int m_pairs = (p.m_tiles + 1) >> 1;
    const int tiles_per_batch = m_pairs * p.n_tiles;
    const int total_tiles = tiles_per_batch * p.batch;
    const int k_blocks = p.taps * p.k_blocks_per_tap;
    const int pair = blockIdx.x >> 1, pairs = gridDim.x >> 1;
    const int half_n = p.bn >> 1;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs; warp-uniform control flow, elected issue) =====================
        const uint32_t sa0 = smem_u32(smem_a), sb0 = smem_u32(smem_b);
        const uint32_t b_half_bytes = p.b_kmajor ? (uint32_t)half_n * (BLOCK_K * 2) : (uint32_t)(half_n / 64) * 8192u;
        const uint32_t tx_bytes = 2u * (P_A_BYTES + b_half_bytes);
        int stage = 0; uint32_t phase = 0;
        for (int tile = pair; tile < total_tiles; tile += pairs) {
            const int b = tile / tiles_per_batch, r = tile % tiles_per_batch;
            const int mp = r % m_pairs, nt = r / m_pairs;
            const int mt = 2 * mp + (int)rank;                       // this CTA's 128-row tile (may lie past the end: TMA zero-fills)
            const int n0 = nt * p.bn + (int)rank * half_n;           // this CTA's half of the B columns
            int y0 = 0, x0 = 0;
            const int m0 = mt * BLOCK_M;
            if (p.bh > 0) { y0 = (mt / p.tiles_x) * p.bh; x0 = (mt % p.tiles_x) * p.bw; }
            int tap = 0, kcb = 0, ky = 0, kx = 0;
            const int ax = x0 * p.stride - p.pad_left, ay = y0 * p.stride - p.pad_top;
            for (int kb = 0; kb < k_blocks; kb++) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(&full[stage], tx_bytes);       // both CTAs' bytes land on this barrier
                    const int kc = kcb * BLOCK_K;
                    const uint32_t sa = sa0 + stage * P_A_BYTES, sb = sb0 + stage * P_B_BYTES;
                    if (p.bh > 0) tma_load_3d_2sm(sa, &map_a, &full[stage], kc, ax + kx, ay + ky);
                    else if (p.a_swap) tma_load_3d_2sm(sa, &map_a, &full[stage], kc, b, m0);
                    else tma_load_3d_2sm(sa, &map_a, &full[stage], kc, m0, b);
                    const int kglob = tap * p.K + kc;
                    if (p.b_kmajor) {
                        if (p.b_swap) tma_load_3d_2sm(sb, &map_b, &full[stage], kglob, b, n0);
                        else tma_load_3d_2sm(sb, &map_b, &full[stage], kglob, n0, b);
                    } else {
                        for (int a = 0; a < half_n / 64; a++) {
                            if (p.b_swap) tma_load_3d_2sm(sb + a * 8192, &map_b, &full[stage], n0 + 64 * a, b, kglob);
                            else tma_load_3d_2sm(sb + a * 8192, &map_b, &full[stage], n0 + 64 * a, kglob, b);
                        }
                    }
                    if (!leader) mbar_arrive_cluster(&full[stage], 0);        // second arrival, no bytes of its own to announce
                }
                __syncwarp();
                if (++kcb == p.k_blocks_per_tap) { kcb = 0; tap++; if (++kx == p.kw) { kx = 0; ky++; } }
                if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            uint32_t idesc = 0;
            idesc |= 1u << 4;                                    // D = f32
            idesc |= (uint32_t)(p.b_kmajor ? 0 : 1) << 16;       // B major
            idesc |= (uint32_t)(p.bn >> 3) << 17;                // N
            idesc |= (uint32_t)(256 >> 4) << 24;                 // M = 256 (the pair)
            const uint64_t adesc0 = make_smem_desc(smem_u32(smem_a), 16, 1024);
            const uint64_t bdesc0 = p.b_kmajor ? make_smem_desc(smem_u32(smem_b), 16, 1024) : make_smem_desc(smem_u32(smem_b), 8192, 1024);
            const uint32_t b_kstep = p.b_kmajor ? (UMMA_K * 2) >> 4 : (UMMA_K * 128) >> 4;
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = pair; tile < total_tiles; tile += pairs) {
                mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.bn);
                for (int kb = 0; kb < k_blocks; kb++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = adesc0 + (uint64_t)(stage * (P_A_BYTES >> 4));
                    const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (P_B_BYTES >> 4));
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; k++)
                            umma_f16_2sm(tmem_d, adesc + (uint64_t)(k * ((UMMA_K * 2) >> 4)), bdesc + (uint64_t)(k * b_kstep), idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit_2sm(&empty[stage], 0b11);                          // frees the slot in BOTH CTAs
                        if (kb == k_blocks - 1) umma_commit_2sm(&acc_full[acc], 0b11); // wakes the epilogue of BOTH CTAs
                    }
                    __syncwarp();
                    if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        osb_pdl_trigger_late();
    } else {
        // ===================== epilogue (warps 2..5, both CTAs): this CTA's 128 rows x bn columns =====================
        const int q = warp & 3;
        const uint32_t stg0 = smem_u32(smem_stg) + (uint32_t)q * P_STG_WARP;
        int acc = 0; uint32_t acc_phase = 0;
        int panels_issued = 0;
        for (int tile = pair; tile < total_tiles; tile += pairs) {
            const int b = tile / tiles_per_batch, r = tile % tiles_per_batch;
            const int mp = r % m_pairs, nt = r / m_pairs;
            const int mt = 2 * mp + (int)rank;
            const int n0 = nt * p.bn;
            const int n_end = min(p.N, n0 + p.bn);
            const int row_in_tile = q * 32 + lane;
            long long out_row; bool row_ok;
            int sx = 0, sy = 0;              // store-box origin of this warp (conv: pixel coordinates; GEMM: row index in sy)
            if (p.bh > 0) {
                const int ty = (mt / p.tiles_x) * p.bh, tx = (mt % p.tiles_x) * p.bw;
                const int y = ty + row_in_tile / p.bw, x = tx + row_in_tile % p.bw;
                row_ok = mt < p.m_tiles && y < p.Ho && x < p.Wo;
                out_row = (long long)y * p.Wo + x;
                sx = tx + (q * 32) % p.bw; sy = ty + (q * 32) / p.bw;
            } else {
                const int m = mt * BLOCK_M + row_in_tile;
                row_ok = m < p.M;
                out_row = m;
                sy = mt * BLOCK_M + q * 32;
            }
            const __half* rrow = p.residual ? p.residual + (long long)b * p.stride_c + out_row * p.ldc : nullptr;
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            if (mt < p.m_tiles) {
#pragma unroll 1
                for (int c = 0; c < p.bn; c += 32) {
                    if (n0 + c >= n_end) break;     // warp-uniform
                    const int half = (c >> 5) & 1, panel = c >> 6;
                    const uint32_t buf = stg0 + (uint32_t)((panels_issued & 1) * 4096);
                    if (half == 0 && panels_issued >= 2) {      // the store that read this buffer two panels ago must have drained it
                        if (lane == 0) bulk_wait_read<1>();
                        __syncwarp();
                    }
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.bn + c), v);
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const int n = n0 + c + j;
                        float f[8];
#pragma unroll
                        for (int t = 0; t < 8; t++) f[t] = __uint_as_float(v[j + t]);
                        if (n < n_end) {           // N % 8 == 0: a unit is entirely inside or outside
                            if (p.bias) {
                                Vec<__half, 8> bv = load_vec<__half, 8>(p.bias + n);
#pragma unroll
                                for (int t = 0; t < 8; t++) f[t] += __half2float(bv.v[t]);
                            }
                            if (p.bias2) {
                                Vec<__half, 8> bv = load_vec<__half, 8>(p.bias2 + n);
#pragma unroll
                                for (int t = 0; t < 8; t++) f[t] += __half2float(bv.v[t]);
                            }
                            if (rrow && row_ok) {
                                Vec<__half, 8> rv = load_vec<__half, 8>(rrow + n);
#pragma unroll
                                for (int t = 0; t < 8; t++) f[t] += __half2float(rv.v[t]);
                            }
                        }
                        Vec<__half, 8> o;
#pragma unroll
                        for (int t = 0; t < 8; t++) o.v[t] = __float2half_rn(f[t]);
#pragma unroll
                        for (int t = 0; t < 4; t++) v[(j >> 1) + t] = (row_ok && n < n_end) ? *reinterpret_cast<uint32_t*>(&o.v[2 * t]) : 0u;   // rounded values, for the statistics
                        // staging panel = 32 rows x 128 bytes in the TMA SWIZZLE_128B layout: 16-byte unit u of row r lives at u ^ (r & 7)
                        const uint32_t unit = (uint32_t)(half * 4 + (j >> 3));
                        const uint32_t addr = buf + (uint32_t)lane * 128u + ((unit ^ ((uint32_t)lane & 7u)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(*reinterpret_cast<uint32_t*>(&o.v[0])), "r"(*reinterpret_cast<uint32_t*>(&o.v[2])),
                                     "r"(*reinterpret_cast<uint32_t*>(&o.v[4])), "r"(*reinterpret_cast<uint32_t*>(&o.v[6])) : "memory");
                    }
                    if (p.gn_stats) {
                        uint32_t h[16];
#pragma unroll
                        for (int t = 0; t < 16; t++) h[t] = v[t];
                        gn_stats_chunk(smem_u32(gn_buf) + (uint32_t)q * 2048u, h, n0 + c, n_end, p.gn_cpg, gn_acc, lane);
                    }
                    const bool last_of_panel = half == 1 || c + 32 >= p.bn || n0 + c + 32 >= n_end;
                    if (last_of_panel) {
                        fence_async_smem();       // generic-proxy writes -> visible to the async proxy (TMA)
                        __syncwarp();
                        if (lane == 0) {
                            // columns past N and rows past M / the image are clipped by the store; a half-filled last panel only ever
                            // occurs at the N edge (bn % 64 == 0), where the unwritten half is out of bounds
                            if (p.bh > 0) tma_store_3d(&map_c, buf, n0 + panel * 64, sx, sy);
                            else tma_store_3d(&map_c, buf, n0 + panel * 64, sy, b);
                            bulk_commit();
                        }
                        panels_issued++;
                    }
                }
            }
            if (p.gn_stats) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                gn_stats_flush(gn_acc, p.gn_stats, p.gn_groups, (int)threadIdx.x - 64);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&acc_empty[acc], 0);     // one arrival per epilogue warp of the pair, on the leader's barrier
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) bulk_wait_all();      // staging memory must outlive the stores that read it
    }

    tc_fence_before();
    cluster_sync_all();          // nobody frees TMEM (or exits, taking its shared memory away) while the peer may still signal it
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

}  // namespace pairk
