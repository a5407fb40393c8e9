// Dense layer GEMM on the 5th-generation tensor cores:  C[M,N] = act(A[M,K] * Bt[N,K]^T + bias[N])
//
// Used for the fully connected layers of the AlexNet hash head (lib/architecture.py:363-382: fc6 9216->4096,
// fc7 4096->4096, fc8 = lib/ops.py:287-302 `linear` 4096->HASH_DIM).  fp32 operands are consumed as TF32 by
// tcgen05.mma (kind::tf32), accumulated in fp32 in tensor memory.
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0      : TMA producer  -- cp.async.bulk.tensor.2d of a 128x32 A tile and a BNx32 B tile (128-byte rows,
//                 SWIZZLE_128B) into a 4-stage shared-memory ring, completion on mbarriers
//   warp 1      : allocates BN TMEM columns, issues tcgen05.mma (one elected thread; 4 MMAs of K=8 per stage),
//                 tcgen05.commit releases the stage / signals the epilogue
//   warps 2..5  : epilogue -- tcgen05.ld 32 lanes x 32 columns, + bias, ReLU, float4 stores
#include "umma.cuh"

#include <climits>
#include <cstdlib>

namespace hg {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 32;  // fp32 elements = 128 bytes = one swizzle atom row
constexpr int kGemmStages = 4;
constexpr int kGemmThreads = 192;

// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = BN
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_c),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int BN, bool RELU>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const float* __restrict__ bias,
                 float* __restrict__ C, int M, int N, int K, int ldc)
{
    extern __shared__ __align__(1024) uint8_t gsm[];
    constexpr uint32_t A_BYTES = kGemmBM * kGemmBK * 4;  // 16 KB
    constexpr uint32_t B_BYTES = BN * kGemmBK * 4;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    // the dynamic shared memory base is only 16-byte aligned by contract: round up to the 1024 bytes SWIZZLE_128B needs
    const uint32_t base = (smem_u32(gsm) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t full_bar[kGemmStages], empty_bar[kGemmStages], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * BN;
    const int nkb = K / kGemmBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGemmStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {  // TMEM allocation: one warp, power-of-two column count >= 32
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kGemmStages;
                mbar_wait(&empty_bar[s], (uint32_t)(((kb / kGemmStages) & 1) ^ 1));
                mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_BYTES;
                tma_load_2d(sa, &tmap_a, smem_u32(&full_bar[s]), kb * kGemmBK, m0);
                tma_load_2d(sb, &tmap_b, smem_u32(&full_bar[s]), kb * kGemmBK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(kGemmBM, BN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kGemmStages;
                mbar_wait(&full_bar[s], (uint32_t)((kb / kGemmStages) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_BYTES;
                const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sb);
#pragma unroll
                for (int k = 0; k < kGemmBK / 8; ++k)  // UMMA_K = 8 tf32 = 32 bytes: advance the start address by 2 (x16 B)
                    umma_tf32(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                umma_commit(&empty_bar[s]);  // the stage may be refilled once these MMAs have read it
            }
            umma_commit(&tmem_full_bar);     // accumulator complete
        }
    } else {
        // epilogue warps 2..5: TMEM lane quarter = warp % 4
        const int quarter = warp & 3;
        mbar_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + quarter * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < M) {
                float* crow = C + (size_t)row * ldc + n0 + c0;
                const bool vec = ((reinterpret_cast<uintptr_t>(crow) & 15) == 0) && (n0 + c0 + 32 <= N);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float v[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int n = n0 + c0 + j + t;
                        float x = __uint_as_float(r[j + t]) + ((bias != nullptr && n < N) ? __ldg(bias + n) : 0.0f);
                        v[t] = RELU ? fmaxf(x, 0.0f) : x;
                    }
                    if (vec) {
                        *reinterpret_cast<float4*>(crow + j) = make_float4(v[0], v[1], v[2], v[3]);
                    } else {
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (n0 + c0 + j + t < N) crow[j + t] = v[t];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

// ================================================================================================================
// Convolution as an IMPLICIT GEMM on the tensor cores (lib/architecture.py:253-351):
//   out[m, co] = relu(bias[co] + sum_k A[m, k] * Wt[co, k]),  m = (crop, oy, ox),  k = (ky, kx, ci) of one channel group
// The im2col matrix A is never written: the four producer warps gather each 128 x 32 A tile straight from the NHWC
// activations (16-byte loads, 8 lanes cover the 128 bytes of a row so a warp reads four whole rows) and store it into the
// shared-memory stage in the SWIZZLE_128B K-major layout tcgen05.mma expects (chunk ^= row & 7), fence.proxy.async,
// mbarrier arrive.  The weights (B operand, [Cout_g, Kpad] K-major) arrive by TMA as in gemm_tf32_kernel; the producer
// warps then turn into the epilogue warps (bias + ReLU, NHWC stores).
// ================================================================================================================
struct ConvGemmArgs {
    const float* in;    // [N, H, W, C] NHWC (C = stored channels, multiple of 4)
    const float* bias;  // [Cog] of this group
    float* out;         // [M, ldc], already offset to the group's first output channel
    int64_t M;
    int H, W, C, c0, Cg, KH, KW, stride, pad, Ho, Wo, Kpad, Cog, ldc;
    int relu;  // 1: ReLU after the bias (convolutions, fc6, fc7); 0: linear (fc8)
};

__host__ __device__ constexpr int conv_tmem_cols(int BN) { return BN <= 64 ? 64 : (BN <= 128 ? 128 : 256); }
// X3 = error-compensated TF32 ("3xTF32"): every fp32 operand is split into hi (its upper 19 bits, exactly a TF32 number) and
// lo = x - hi; hi*hi + lo*hi + hi*lo accumulated in the fp32 accumulator reproduces the fp32 product to ~2^-21, so the
// tensor-core convolution matches the fp32 graph like the CUDA-core one does.  Stage = [A hi][A lo][B hi][B lo].
// MT = M tiles (128 rows each, one accumulator each) per CTA.  With MT = 2 a weight (B) tile fetched from L2 serves 256 output
// rows: the weight tiles are two thirds of the kernel's L2 traffic at conv2 (every 128-row CTA re-reads the whole 1.2 MB
// filter bank of its group: 9 of 13 GB), and the kernel is bound by exactly that traffic.
__host__ __device__ constexpr int conv_stages(int BN, bool X3, int MT) { return X3 ? (MT == 2 ? 2 : (BN <= 128 ? 3 : 2)) : 3; }

template <int BN, bool X3, int MT>
__global__ void __launch_bounds__(kGemmThreads, (!X3 && BN <= 128 && MT == 1 ? 2 : 1))
conv_gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_blo, ConvGemmArgs a)
{
    extern __shared__ __align__(1024) uint8_t csm[];
    constexpr int kConvStages = conv_stages(BN, X3, MT);
    constexpr uint32_t A_BYTES = kGemmBM * kGemmBK * 4;  // 16 KB
    constexpr uint32_t B_BYTES = BN * kGemmBK * 4;
    constexpr uint32_t A_TILE = (X3 ? 2 : 1) * A_BYTES;   // [hi | lo] of one M tile
    constexpr uint32_t A_ALL = MT * A_TILE;
    constexpr uint32_t STAGE_BYTES = A_ALL + (X3 ? 2 : 1) * B_BYTES;
    constexpr uint32_t TCOLS = conv_tmem_cols(BN);           // accumulator columns of one M tile
    const uint32_t base = (smem_u32(csm) + 1023u) & ~1023u;
    uint8_t* const base_ptr = csm + (base - smem_u32(csm));
    int2* const tab = reinterpret_cast<int2*>(base_ptr + kConvStages * STAGE_BYTES);  // per 16-byte K chunk: {input offset, ky | kx << 16}
    __shared__ __align__(8) uint64_t full_bar[kConvStages], empty_bar[kConvStages], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m0 = (int64_t)blockIdx.y * (kGemmBM * MT);
    const int n0 = blockIdx.x * BN;
    const int nkb = a.Kpad / kGemmBK;
    const int K = a.KH * a.KW * a.Cg;

    for (int j = threadIdx.x; j < a.Kpad / 4; j += kGemmThreads) {
        const int k = j * 4;
        if (k < K) {
            const int ci = k % a.Cg, kx = (k / a.Cg) % a.KW, ky = k / (a.Cg * a.KW);
            tab[j] = make_int2((ky * a.W + kx) * a.C + ci, ky | (kx << 16));
        } else {
            tab[j] = make_int2(-1, 0);
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kConvStages; ++s) { mbar_init(&full_bar[s], 1 + 4); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)(MT * TCOLS)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {  // weights by TMA
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kConvStages;
                mbar_wait(&empty_bar[s], (uint32_t)(((kb / kConvStages) & 1) ^ 1));
                mbar_arrive_expect_tx(&full_bar[s], (X3 ? 2 : 1) * B_BYTES);
                tma_load_2d(base + s * STAGE_BYTES + A_ALL, &tmap_b, smem_u32(&full_bar[s]), kb * kGemmBK, n0);
                if (X3) tma_load_2d(base + s * STAGE_BYTES + A_ALL + B_BYTES, &tmap_blo, smem_u32(&full_bar[s]), kb * kGemmBK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(kGemmBM, BN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kConvStages;
                mbar_wait(&full_bar[s], (uint32_t)((kb / kConvStages) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_ALL;
                const uint64_t bdesc = umma_desc_sw128(sb);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const uint64_t adesc = umma_desc_sw128(sa + mt * A_TILE);
                    const uint32_t acc = tmem_base + (uint32_t)(mt * TCOLS);
#pragma unroll
                    for (int k = 0; k < kGemmBK / 8; ++k) {
                        umma_tf32(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                        if (X3) {
                            const uint64_t alo = umma_desc_sw128(sa + mt * A_TILE + A_BYTES), blo = umma_desc_sw128(sb + B_BYTES);
                            umma_tf32(acc, alo + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
                            umma_tf32(acc, adesc + (uint64_t)(2 * k), blo + (uint64_t)(2 * k), idesc, 1u);
                        }
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&tmem_full_bar);
        }
    } else {
        // ---- producer: gather the A tiles (thread g: 16-byte chunk g & 7 of rows (g >> 3) + 16 i) ----
        const int g = threadIdx.x - 64;
        const int chunk = g & 7;
        int rbase[MT][8];  // input offset of the window origin of row i of M tile mt (floats), INT_MIN: row beyond M
        int ryx[MT][8];    // (iy0 + 1024) | (ix0 + 1024) << 16
        uint32_t soff[8];  // byte offset of my chunk inside an A tile
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = (g >> 3) + 16 * i;
            soff[i] = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4));
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int64_t m = m0 + mt * kGemmBM + r;
                if (m < a.M) {
                    const int ox = (int)(m % a.Wo), oy = (int)((m / a.Wo) % a.Ho);
                    const int64_t n = m / ((int64_t)a.Wo * a.Ho);
                    const int iy0 = oy * a.stride - a.pad, ix0 = ox * a.stride - a.pad;
                    rbase[mt][i] = (int)(((n * a.H + iy0) * a.W + ix0) * a.C + a.c0);
                    ryx[mt][i] = (iy0 + 1024) | ((ix0 + 1024) << 16);
                } else {
                    rbase[mt][i] = INT_MIN;
                    ryx[mt][i] = 0;
                }
            }
        }
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kConvStages;
            const int2 t = tab[kb * 8 + chunk];
            const int ky = t.y & 0xffff, kx = t.y >> 16;
            float4 v[MT][8];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    v[mt][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int iy = (ryx[mt][i] & 0xffff) - 1024 + ky, ix = (ryx[mt][i] >> 16) - 1024 + kx;
                    if (t.x >= 0 && rbase[mt][i] != INT_MIN && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
                        v[mt][i] = __ldg(reinterpret_cast<const float4*>(a.in + rbase[mt][i] + t.x));
                }
            mbar_wait(&empty_bar[s], (uint32_t)(((kb / kConvStages) & 1) ^ 1));  // the loads above are already in flight
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                uint8_t* sa = base_ptr + s * STAGE_BYTES + mt * A_TILE;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (X3) {
                        float4 hi, lo;
                        hi.x = __uint_as_float(__float_as_uint(v[mt][i].x) & 0xFFFFE000u); lo.x = v[mt][i].x - hi.x;
                        hi.y = __uint_as_float(__float_as_uint(v[mt][i].y) & 0xFFFFE000u); lo.y = v[mt][i].y - hi.y;
                        hi.z = __uint_as_float(__float_as_uint(v[mt][i].z) & 0xFFFFE000u); lo.z = v[mt][i].z - hi.z;
                        hi.w = __uint_as_float(__float_as_uint(v[mt][i].w) & 0xFFFFE000u); lo.w = v[mt][i].w - hi.w;
                        *reinterpret_cast<float4*>(sa + soff[i]) = hi;
                        *reinterpret_cast<float4*>(sa + A_BYTES + soff[i]) = lo;
                    } else {
                        *reinterpret_cast<float4*>(sa + soff[i]) = v[mt][i];
                    }
                }
            }
            fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[s])) : "memory");
            }
        }
        // ---- epilogue: TMEM lane quarter = warp % 4 ----
        const int quarter = warp & 3;
        mbar_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
        const int64_t row = m0 + mt * kGemmBM + quarter * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * TCOLS + c0);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < a.M && n0 + c0 < a.Cog) {
                float* crow = a.out + (size_t)row * a.ldc + n0 + c0;
                const bool vec = ((reinterpret_cast<uintptr_t>(crow) & 15) == 0) && (n0 + c0 + 32 <= a.Cog);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float o[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int n = n0 + c0 + j + t;
                        const float x = __uint_as_float(r[j + t]) + (n < a.Cog ? __ldg(a.bias + n) : 0.0f);
                        o[t] = a.relu ? fmaxf(x, 0.0f) : x;
                    }
                    if ((BN % 32) != 0 && c0 + j >= BN) break;  // BN = 144: the last 32-column chunk is half a tile wide (BN % 4 == 0)
                    if (vec) {
                        *reinterpret_cast<float4*>(crow + j) = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (n0 + c0 + j + t < a.Cog) crow[j + t] = o[t];
                    }
                }
            }
        }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(MT * TCOLS)) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        (void)cudaGetLastError();
    }
    return fn;
}

// [rows, K] fp32 row-major (row stride ld elements) -> 2-D tensor map, box = 32 x box_rows, 128-byte swizzle, OOB -> 0
static int make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(HG_ECUDA, "gemm_tf32: cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HG_ECUDA, "gemm_tf32: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return HG_OK;
}

template <int BN, bool X3, int MT>
static int launch_conv_gemm_mt(const CUtensorMap& tb, const CUtensorMap& tblo, const ConvGemmArgs& a, cudaStream_t st)
{
    const size_t smem = (size_t)conv_stages(BN, X3, MT) * (X3 ? 2 : 1) * (MT * kGemmBM + BN) * kGemmBK * 4 + (size_t)(a.Kpad / 4) * sizeof(int2) + 1024;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        HG_CUDA_TRY(cudaFuncSetAttribute(conv_gemm_tf32_kernel<BN, X3, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((unsigned)ceil_div(a.Cog, BN), (unsigned)ceil_div(a.M, kGemmBM * MT));
    conv_gemm_tf32_kernel<BN, X3, MT><<<grid, kGemmThreads, smem, st>>>(tb, tblo, a);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

// two M tiles per CTA for the error-compensated mode (BN <= 128: 2 stages x 96 KB; BN = 192 would need 2 x 112 KB plus the tap
// table) when there are enough rows; HG_CONV_MT=1 keeps one tile per CTA
template <int BN, bool X3>
static int launch_conv_gemm(const CUtensorMap& tb, const CUtensorMap& tblo, const ConvGemmArgs& a, cudaStream_t st)
{
    static const int mt_env = []() { const char* v = getenv("HG_CONV_MT"); return (v && *v) ? atoi(v) : 0; }();
    if constexpr (X3 && BN <= 144) {
        // (a grid that cannot even fill a quarter of the SMs -- fc8: 64 outputs, M = 1280 -> 5 CTAs of 256 rows -- keeps 128-row CTAs)
        const int sms = device_facts().sm_count > 0 ? device_facts().sm_count : 148;
        const bool tiny = ceil_div(a.Cog, BN) * ceil_div(a.M, 2 * kGemmBM) * 4 < sms;
        if (mt_env != 1 && a.M >= 2 * kGemmBM && (mt_env == 2 || !tiny)) return launch_conv_gemm_mt<BN, X3, 2>(tb, tblo, a, st);
    }
    return launch_conv_gemm_mt<BN, X3, 1>(tb, tblo, a, st);
}

// one channel group of a convolution layer; wt = this group's weights [Cog, Kpad] K-major (hg_conv_weight_pack: upper 19
// bits of every weight), wt_lo = the remainders for the error-compensated mode (NULL: plain TF32)
int conv_gemm_tf32(const float* in, const float* wt, const float* wt_lo, const float* bias, float* out, int64_t M, int H, int W, int C, int c0,
                   int Cg, int KH, int KW, int stride, int pad, int Ho, int Wo, int Kpad, int Cog, int ldc, cudaStream_t st, int relu)
{
    if ((Cg % 4) || (C % 4) || (c0 % 4) || (reinterpret_cast<uintptr_t>(in) & 15) || (Kpad % kGemmBK))
        return fail(HG_EINVAL, "conv_gemm_tf32: channels must be multiples of 4, Kpad a multiple of %d", kGemmBK);
    if ((int64_t)M / ((int64_t)Ho * Wo) * H * W * C >= (int64_t(1) << 31)) return fail(HG_EINVAL, "conv_gemm_tf32: input too large for 32-bit offsets");
    ConvGemmArgs a{};
    a.in = in; a.bias = bias; a.out = out; a.M = M; a.H = H; a.W = W; a.C = C; a.c0 = c0; a.Cg = Cg; a.KH = KH; a.KW = KW; a.stride = stride;
    a.pad = pad; a.Ho = Ho; a.Wo = Wo; a.Kpad = Kpad; a.Cog = Cog; a.ldc = ldc; a.relu = relu;
    int BN = (Cog % 128 == 0) ? 128 : ((Cog % 192 == 0) ? 192 : ((Cog % 96 == 0) ? 96 : (Cog <= 64 ? 64 : 128)));
    if (wt_lo && BN == 128 && M >= 2 * kGemmBM) {
        // One 256-row CTA per SM: the launch takes ceil(CTAs / SMs) rounds, each ~BN long.  fc6 / fc7 (4096 outputs, M = 1280): 32 x 5 =
        // 160 CTAs of 128 columns are two rounds, the second one 8 % full; 29 x 5 = 145 CTAs of 144 columns (the last tile 64 wide)
        // are ONE round -- fc6 + fc7 1.20 -> 0.60 ms.  Convolutions (thousands of CTAs) keep 128.
        const int sms = device_facts().sm_count > 0 ? device_facts().sm_count : 148;
        const int64_t my = ceil_div(M, 2 * kGemmBM);
        const int64_t cost128 = ceil_div(ceil_div(Cog, 128) * my, sms) * 128, cost144 = ceil_div(ceil_div(Cog, 144) * my, sms) * 144;
        if (cost144 < cost128) BN = 144;
    }
    CUtensorMap tb, tblo;
    int rc;
    if ((rc = make_map(&tb, wt, Cog, Kpad, Kpad, BN)) != HG_OK) return rc;
    if ((rc = make_map(&tblo, wt_lo ? wt_lo : wt, Cog, Kpad, Kpad, BN)) != HG_OK) return rc;
    if (wt_lo) {
        switch (BN) {
            case 64: return launch_conv_gemm<64, true>(tb, tblo, a, st);
            case 96: return launch_conv_gemm<96, true>(tb, tblo, a, st);
            case 144: return launch_conv_gemm<144, true>(tb, tblo, a, st);
            case 192: return launch_conv_gemm<192, true>(tb, tblo, a, st);
            default: return launch_conv_gemm<128, true>(tb, tblo, a, st);
        }
    }
    switch (BN) {
        case 64: return launch_conv_gemm<64, false>(tb, tblo, a, st);
        case 96: return launch_conv_gemm<96, false>(tb, tblo, a, st);
        case 192: return launch_conv_gemm<192, false>(tb, tblo, a, st);
        default: return launch_conv_gemm<128, false>(tb, tblo, a, st);
    }
}

template <int BN, bool RELU>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const float* bias, float* C, int M, int N, int K, int ldc, cudaStream_t st)
{
    const size_t smem = (size_t)kGemmStages * (kGemmBM + BN) * kGemmBK * 4 + 1024;
    static thread_local bool configured = false;
    if (!configured) {
        HG_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32_kernel<BN, RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, kGemmBM));
    gemm_tf32_kernel<BN, RELU><<<grid, kGemmThreads, smem, st>>>(ta, tb, bias, C, M, N, K, ldc);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

int gemm_tf32(const float* A, int64_t lda, const float* Bt, int64_t ldb, const float* bias, float* C, int64_t ldc, int M, int N, int K, int relu,
              cudaStream_t st)
{
    if (M <= 0 || N <= 0) return HG_OK;
    if (K <= 0 || (K % kGemmBK) != 0) return fail(HG_EINVAL, "gemm_tf32: K=%d must be a positive multiple of %d", K, kGemmBK);
    if (!A || !Bt || !C) return fail(HG_EINVAL, "gemm_tf32: NULL pointer");
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(Bt) & 15) || (lda % 4) || (ldb % 4))
        return fail(HG_EINVAL, "gemm_tf32: A / Bt must be 16-byte aligned with row strides that are multiples of 4 floats");
    const int BN = N <= 64 ? 64 : 128;
    CUtensorMap ta, tb;
    int rc;
    if ((rc = make_map(&ta, A, M, K, lda, kGemmBM)) != HG_OK) return rc;
    if ((rc = make_map(&tb, Bt, N, K, ldb, BN)) != HG_OK) return rc;
    if (BN == 64) return relu ? launch_gemm<64, true>(ta, tb, bias, C, M, N, K, (int)ldc, st) : launch_gemm<64, false>(ta, tb, bias, C, M, N, K, (int)ldc, st);
    return relu ? launch_gemm<128, true>(ta, tb, bias, C, M, N, K, (int)ldc, st) : launch_gemm<128, false>(ta, tb, bias, C, M, N, K, (int)ldc, st);
}

}  // namespace hg

extern "C" int hg_gemm_tf32(const float* d_a, int64_t lda, const float* d_bt, int64_t ldb, const float* d_bias, float* d_c, int64_t ldc, int M,
                            int N, int K, int relu, void* stream)
{
    if (!hg::device_facts().ok) return hg::fail(HG_ECUDA, "hg_gemm_tf32: no CUDA device");
    return hg::gemm_tf32(d_a, lda, d_bt, ldb, d_bias, d_c, ldc, M, N, K, relu, (cudaStream_t)stream);
}
