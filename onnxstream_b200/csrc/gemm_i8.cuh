// gemm_i8.cuh -- W8A8 on the tensor cores: `tcgen05.mma.kind::i8`, uint8 x uint8 -> int32 in TMEM (included by gemm_tcgen05.cu).
//
// Replaces XnnPack::matrix_multiply<uint8_t,int32_t> / convolution for T = uint8_t (src/onnxstream.cpp:1035-1215, 1292-1534; the reference
// drives XNNPACK's qu8 GEMM / conv with the parameters of Qu8MatMulData / Qu8ConvData, 1025-1033) for the shapes TMA can address
// (16-byte strides); other shapes stay on the CUDA-core igemm_kernel (kernels_gemm.cu), same arithmetic.
//
// Arithmetic (bit-exact with XNNPACK's "fp32" requantisation, pinned in tests against the reference run):
//     acc[m,n] = sum_k (x[m,k] - zx) * (w[k,n] - zw) + bias[n]                    (int32)
//     y[m,n]   = clamp(lrintf(acc * (sx * sw / sy)), 0 - zy, 255 - zy) + zy       (uint8)
// The tensor core multiplies the RAW unsigned bytes: raw[m,n] = sum_k x*w, and the epilogue applies
//     acc = raw - zw * rowsum_x[m] - zx * colsum_w[n] + K * zx * zw + bias[n]
// with rowsum_x / colsum_w from small byte-sum kernels.  A convolution pads with the INPUT ZERO POINT in XNNPACK, while TMA fills
// out-of-bounds elements with 0: the host therefore pads the NHWC image once (value zx, kernel in this file) and the conv runs
// un-padded on it, so every tap of every output pixel is a real byte and rowsum_x[m] is the sum of kh*kw per-pixel channel sums.
//
// Structure = tc_gemm_kernel (warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue, 2 accumulator stages in TMEM), with
// 128-byte k-blocks = 128 elements, UMMA_K = 32.  A is K-major; B is K-major (OHWI conv weights) or MN-major ([K,N] MatMul weights --
// MN-major is valid for 8-bit integer operands, cute/arch/mma_sm100_desc.hpp InstrDescriptor::b_major_).

namespace i8k {

constexpr int I8_BLOCK_K = 128;        // elements = bytes: one 128B swizzle row
constexpr int I8_UMMA_K = 32;
constexpr int I8_STAGES = 4;
constexpr int I8_SMEM = I8_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 + 256;

struct I8Params {
    int M, N, K;                 // GEMM view (conv: M = Ho*Wo, K = Cin per tap)
    int m_tiles, n_tiles, bn;
    int b_kmajor;
    int taps, kw, Wo, Ho, bw, bh, tiles_x, k_blocks_per_tap, stride;
    int Wp;                      // conv: width of the zero-point-padded image (rows of `psum`)
    const int32_t* rsum;         // GEMM: rowsum_x[M];  conv: per-pixel channel sums of the padded image [Hp * Wp]
    const int32_t* csum;         // colsum_w[N]
    const int32_t* bias;         // int32 bias[N] or null
    int zx, zw, zy;
    int kzz;                     // K_total * zx * zw
    float requant;               // sx * sw / sy
    uint8_t* C;
    long long ldc;
};

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_i8_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const I8Params p)
{
    osb_pdl_trigger_entry();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + I8_STAGES * A_STAGE_BYTES;
    uint64_t* bars = (uint64_t*)(smem + I8_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
    uint64_t* full = bars;
    uint64_t* empty = bars + I8_STAGES;
    uint64_t* acc_full = bars + 2 * I8_STAGES;
    uint64_t* acc_empty = acc_full + ACC_STAGES;
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < I8_STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
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
    osb_pdl_wait();

    const int total_tiles = p.m_tiles * p.n_tiles;
    const int k_blocks = p.taps * p.k_blocks_per_tap;

    if (warp == 0) {
        // ===================== TMA producer =====================
        const uint32_t sa0 = smem_u32(smem_a), sb0 = smem_u32(smem_b);
        const uint32_t tx_bytes = A_STAGE_BYTES + (p.b_kmajor ? (uint32_t)p.bn * I8_BLOCK_K : (uint32_t)B_STAGE_BYTES);
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile % p.m_tiles, nt = tile / p.m_tiles;
            const int n0 = nt * p.bn, m0 = mt * BLOCK_M;
            int y0 = 0, x0 = 0;
            if (p.bh > 0) { y0 = (mt / p.tiles_x) * p.bh; x0 = (mt % p.tiles_x) * p.bw; }
            int tap = 0, kcb = 0, ky = 0, kx = 0;
            const int ax = x0 * p.stride, ay = y0 * p.stride;       // the image is already padded: no negative coordinates
            for (int kb = 0; kb < k_blocks; kb++) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&full[stage], tx_bytes);
                    const int kc = kcb * I8_BLOCK_K;
                    const uint32_t sa = sa0 + stage * A_STAGE_BYTES, sb = sb0 + stage * B_STAGE_BYTES;
                    if (p.bh > 0) tma_load_3d_s(sa, &map_a, &full[stage], kc, ax + kx, ay + ky);
                    else tma_load_3d_s(sa, &map_a, &full[stage], kc, m0, 0);
                    const int kglob = tap * p.K + kc;
                    if (p.b_kmajor) tma_load_3d_s(sb, &map_b, &full[stage], kglob, n0, 0);
                    else tma_load_3d_s(sb, &map_b, &full[stage], n0, kglob, 0);      // one box: 128 columns (128 B) x 128 k-rows
                }
                __syncwarp();
                if (++kcb == p.k_blocks_per_tap) { kcb = 0; tap++; if (++kx == p.kw) { kx = 0; ky++; } }
                if (++stage == I8_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        uint32_t idesc = 0;
        idesc |= 2u << 4;                                    // D = S32
        idesc |= 0u << 7;                                    // A = unsigned 8 bit
        idesc |= 0u << 10;                                   // B = unsigned 8 bit
        idesc |= (uint32_t)(p.b_kmajor ? 0 : 1) << 16;       // B major
        idesc |= (uint32_t)(p.bn >> 3) << 17;
        idesc |= (uint32_t)(BLOCK_M >> 4) << 24;
        const uint64_t adesc0 = make_smem_desc(smem_u32(smem_a), 16, 1024);
        const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_b), 16, 1024);     // MN-major: a single 128-column atom, 8-row k-groups 1024 B apart
        const uint32_t a_kstep = I8_UMMA_K >> 4;                                  // 32 bytes along K
        const uint32_t b_kstep = p.b_kmajor ? (I8_UMMA_K >> 4) : ((I8_UMMA_K * 128) >> 4);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(&acc_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
            for (int kb = 0; kb < k_blocks; kb++) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint64_t adesc = adesc0 + (uint64_t)(stage * (A_STAGE_BYTES >> 4));
                const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (B_STAGE_BYTES >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < I8_BLOCK_K / I8_UMMA_K; k++)
                        umma_i8(tmem_d, adesc + (uint64_t)(k * a_kstep), bdesc + (uint64_t)(k * b_kstep), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&empty[stage]);
                    if (kb == k_blocks - 1) umma_commit(&acc_full[acc]);
                }
                __syncwarp();
                if (++stage == I8_STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
        osb_pdl_trigger_late();
    } else {
        // ===================== epilogue: zero-point corrections + fp32 requantisation =====================
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile % p.m_tiles, nt = tile / p.m_tiles;
            const int n0 = nt * p.bn;
            const int n_end = min(p.N, n0 + p.bn);
            const int row_in_tile = q * 32 + lane;
            long long out_row; bool row_ok; int rs = 0;
            if (p.bh > 0) {
                const int y = (mt / p.tiles_x) * p.bh + row_in_tile / p.bw, x = (mt % p.tiles_x) * p.bw + row_in_tile % p.bw;
                row_ok = y < p.Ho && x < p.Wo;
                out_row = (long long)y * p.Wo + x;
                if (row_ok) {
                    const int kh = p.taps / p.kw;
                    for (int ky = 0; ky < kh; ky++)
                        for (int kx = 0; kx < p.kw; kx++) rs += p.rsum[(long long)(y * p.stride + ky) * p.Wp + (x * p.stride + kx)];
                }
            } else {
                const int m = mt * BLOCK_M + row_in_tile;
                row_ok = m < p.M;
                out_row = m;
                if (row_ok) rs = p.rsum[m];
            }
            const int row_term = p.kzz - p.zw * rs;
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            uint8_t* crow = p.C + out_row * p.ldc;
#pragma unroll 1
            for (int c = 0; c < p.bn; c += 32) {
                if (n0 + c >= n_end) break;
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c), v);
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 32; j += 16) {
                        const int n = n0 + c + j;
                        if (n >= n_end) break;            // N % 16 == 0
                        uint32_t packed[4];
#pragma unroll
                        for (int t = 0; t < 16; t++) {
                            int a = (int)v[j + t] + row_term - p.zx * __ldg(p.csum + n + t);
                            if (p.bias) a += __ldg(p.bias + n + t);
                            float f = (float)a * p.requant;
                            f = fminf(fmaxf(f, (float)(0 - p.zy)), (float)(255 - p.zy));
                            const uint32_t o = (uint32_t)((int)rintf(f) + p.zy) & 0xFFu;
                            if ((t & 3) == 0) packed[t >> 2] = o; else packed[t >> 2] |= o << (8 * (t & 3));
                        }
                        *reinterpret_cast<uint4*>(crow + n) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- byte-sum helpers -----------------------------------------------------------------------------------------------

// out[r] = sum of the `cols` bytes of row r (row sums of A, column sums of K-major weights [N][K]); one warp per row
__global__ void rowsum_u8_kernel(const uint8_t* __restrict__ x, int32_t* __restrict__ out, long long rows, long long cols)
{
    osb_pdl_prologue();
    const int lane = threadIdx.x & 31;
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (blockDim.x >> 5)) {
        const uint8_t* row = x + r * cols;
        int s = 0;
        long long c = 0;
        if (((uintptr_t)row & 15) == 0)
            for (c = (long long)lane * 16; c + 16 <= cols; c += 512) {
                const uint4 v = *reinterpret_cast<const uint4*>(row + c);
                s = (int)__dp4a((unsigned)v.x, 0x01010101u, (unsigned)s); s = (int)__dp4a((unsigned)v.y, 0x01010101u, (unsigned)s); s = (int)__dp4a((unsigned)v.z, 0x01010101u, (unsigned)s); s = (int)__dp4a((unsigned)v.w, 0x01010101u, (unsigned)s);
            }
        // tail (and the unaligned case): bytes not covered by the vector loop
        const long long done = ((uintptr_t)row & 15) == 0 ? (cols / 16) * 16 : 0;
        for (long long t = done + lane; t < cols; t += 32) s += row[t];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[r] = s;
    }
}

// out[n] = sum over k of w[k][n] for an MN-major weight matrix [K][N] (ONNX MatMul layout): thread per column, K split over blockIdx.y
__global__ void colsum_u8_kernel(const uint8_t* __restrict__ w, int32_t* __restrict__ out, long long K, long long N, long long k_per)
{
    osb_pdl_prologue();
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const long long k0 = (long long)blockIdx.y * k_per, k1 = min(K, k0 + k_per);
    int s = 0;
    for (long long k = k0; k < k1; k++) s += w[k * N + n];
    atomicAdd(&out[n], s);
}

// Pads an NHWC uint8 image with the input zero point (what XNNPACK's qu8 convolution does, and TMA's zero fill cannot) and writes the
// per-pixel channel sums of the padded image: xp[Hp][Wp][C], psum[Hp * Wp].  One warp per padded pixel.
__global__ void pad_sum_u8_kernel(const uint8_t* __restrict__ x, uint8_t* __restrict__ xp, int32_t* __restrict__ psum, int H, int W, int C, int Hp, int Wp,
                                  int pad_top, int pad_left, int zx)
{
    osb_pdl_prologue();
    const int lane = threadIdx.x & 31;
    const long long pixels = (long long)Hp * Wp;
    for (long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pix < pixels; pix += (long long)gridDim.x * (blockDim.x >> 5)) {
        const int yp = (int)(pix / Wp), xq = (int)(pix % Wp);
        const int y = yp - pad_top, xx = xq - pad_left;
        uint8_t* dst = xp + pix * C;
        int s = 0;
        if (y >= 0 && y < H && xx >= 0 && xx < W) {
            const uint8_t* src = x + ((long long)y * W + xx) * C;
            for (int c = lane * 16; c < C; c += 512) {          // C % 16 == 0
                const uint4 v = *reinterpret_cast<const uint4*>(src + c);
                *reinterpret_cast<uint4*>(dst + c) = v;
                s = (int)__dp4a((unsigned)v.x, 0x01010101u, (unsigned)s); s = (int)__dp4a((unsigned)v.y, 0x01010101u, (unsigned)s); s = (int)__dp4a((unsigned)v.z, 0x01010101u, (unsigned)s); s = (int)__dp4a((unsigned)v.w, 0x01010101u, (unsigned)s);
            }
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        } else {
            const uint32_t z4 = (uint32_t)zx * 0x01010101u;
            for (int c = lane * 16; c < C; c += 512) *reinterpret_cast<uint4*>(dst + c) = make_uint4(z4, z4, z4, z4);
            s = zx * C;
        }
        if (lane == 0) psum[pix] = s;
    }
}

}  // namespace i8k
