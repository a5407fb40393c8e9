// Tensor-core variant of the hot kernel: the all-pairs contraction of lib/metric.py:13 as an EXACT int8 GEMM.
//
// On {-1,+1} codes  ip = sum_k q_k * db_k = b - 2 * d_H  (SURVEY A.3), so a row is a candidate (d_H <= T_q) iff
// ip - (b - 2 T_q) >= 0.  The codes are expanded once to int8 (+1 / -1, zero padding) and contracted with
// tcgen05.mma kind::i8 (int32 accumulation in tensor memory: exact), which moves the contraction off the POPC pipe
// that bounds select_kernel.  The per-query threshold rides in the GEMM as one extra K step
// (A columns hold 2 T_q - b split over two int8 values, B columns hold 1), so the accumulator's SIGN is the answer
// and the epilogue never subtracts:
//
//   per CTA: 256 queries (two 128-row A operands + their threshold columns, loaded once by TMA) x a PAIR of adjacent
//            database splits (a 128-row B tile = 64 rows of each split, so every bin keeps one writer)
//   warp 0      TMA producer: per tile the int8 rows (B operand, swizzled) and the packed rows (code + label words,
//               for the rare path) of both splits into a ring of stages
//   warp 1      TMEM allocator (256 columns = 2 query halves x 128; two CTAs share an SM) and MMA issuer; the two
//               query halves hand their accumulators back and forth independently
//   warps 2..17 epilogue: thread <-> query (TMEM lane), warp <-> (32 queries, one split = 64 accumulator columns).
//               tcgen05.ld ... .pack::16b brings two accumulators per
//               register (|ip'| <= 384 fits 16 bits); one PRMT with sign replication turns two registers into four
//               0x00/0xFF bytes (ALU pipe), one multiply-add drops them into the hit mask (FMA pipe, hit_mask32): 0.25 + 0.25
//               integer op per pair on two different pipes.  The accumulator is
//               then handed back to the MMA warp and the ~R/Ndb hits of the tile are walked in row order: distance
//               recomputed from the packed words (POPC), relevance from the label words, entry appended to the
//               thread's private bin exactly as select_kernel does.
// The PRMT / multiply-add gather leaves accumulator column 4i+k of a 32-column group at mask bit 8k+i; expand_db_kernel stores
// the int8 database rows of every 32-row group in the inverse order, so that mask bit j IS row j of the group.
// The bins, thresholds, AP kernel and exactness guard are shared with the POPC path (rank.cu).
// select_q_kernel (below) is the same contraction with a QUEUED epilogue: the hot kernel of C4 (sparse top-R, 33..64-bit codes).
#include "umma.cuh"

#include <algorithm>
#include <cstdlib>

namespace hg {

constexpr int kUmmaEpiWarps = 16;
constexpr int kUmmaThreads = 32 * (2 + kUmmaEpiWarps);
// KP = int8 bytes per code row: 32 (b <= 32), 64 (b <= 64), 128 (b <= 128), 256 (b <= 256: two 128-byte K blocks).
template <int KP> struct UmmaCfg {
    static constexpr int S = (KP <= 64 ? 4 : 3);              // ring depth: two CTAs must fit one SM (one CTA for KP = 256)
    static constexpr int KB = (KP > 128 ? KP / 128 : 1);      // K blocks of one operand row
    static constexpr int KW = KP / KB;                        // bytes of a K block = swizzle width (32 / 64 / 128)
    static constexpr int CTAS = (KP > 128 ? 1 : 2);           // resident CTAs per SM
    static constexpr int QW = (KP > 128 ? 8 : 4);             // code words a query keeps in registers
    static constexpr uint32_t ROWS_BYTES = 128u * (KP > 128 ? 12u : 8u) * 4u;  // packed-row bytes per stage (Wr <= 8, 12 for b > 128)
};
constexpr int kUmmaHalfRows = 64;  // database rows per split and tile (a B tile = two of these)
constexpr int kUmmaXBytes = 32;    // threshold K step: one UMMA_K of int8 columns, SWIZZLE_32B rows

__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N)
{
    // D = S32 (2), A = B = signed int8 (1), K-major, dense
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_c),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// same on a precomputed shared-space address (keeps the address arithmetic out of the tile loop)
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// one non-blocking probe
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// one probe on a precomputed shared-space address
__device__ __forceinline__ bool mbar_try_a(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "HG_WAITA:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra HG_DONEA;\n"
        "bra HG_WAITA;\n"
        "HG_DONEA:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// 16 code bits starting at bit0 of a packed row -> 16 int8 values: +1 where the bit is set, -1 where it is clear,
// 0 beyond bit b
__device__ __forceinline__ uint4 expand16(const uint32_t* __restrict__ row, int bit0, int b)
{
    uint32_t bits = 0;
    if (bit0 < b) bits = (__ldg(row + (bit0 >> 5)) >> (bit0 & 31)) & 0xFFFFu;
    uint32_t w[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint32_t nib = (bits >> (4 * g)) & 0xFu;
        const uint32_t spread = (nib * 0x00204081u) & 0x01010101u;  // bit i of the nibble -> byte i
        uint32_t v = (spread * 0xFEu) ^ 0xFFFFFFFFu;                // 1 -> 0x01, 0 -> 0xFF
        const int valid = b - (bit0 + 4 * g);                       // bytes of this word that are real code bits
        if (valid <= 0) v = 0;
        else if (valid < 4) v &= (1u << (8 * valid)) - 1u;
        w[g] = v;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// queries: packed code words -> int8 rows [n, KP] in query order
__global__ void __launch_bounds__(256) expand_q_kernel(const uint32_t* __restrict__ rows, int64_t n, int b, int Wr, int KP, uint8_t* __restrict__ out)
{
    const int cpr = KP / 16;  // 16-byte chunks per row
    const int64_t total = n * cpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / cpr;
        const int c = (int)(i - row * cpr);
        *reinterpret_cast<uint4*>(out + row * KP + 16 * c) = expand16(rows + row * Wr, 16 * c, b);
    }
}

// database rows [lo, hi) -> int8 rows; position p of a 32-row group holds row 8 (p & 3) + (p >> 2) of that group
// (see the header: mask bit j of the epilogue is then row j).  Positions up to the end of the last group are
// written; rows >= ndb become zero rows (the epilogue masks them out).
__global__ void __launch_bounds__(256) expand_db_kernel(const uint32_t* __restrict__ rows, int64_t lo, int64_t hi_pad, int64_t ndb, int b, int Wr, int KP,
                                                        uint8_t* __restrict__ out)
{
    const int cpr = KP / 16;
    const int64_t total = (hi_pad - lo) * cpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pos = lo + i / cpr;
        const int c = (int)(i % cpr);
        const int p = (int)(pos & 31);
        const int64_t src = (pos - p) + 8 * (p & 3) + (p >> 2);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (src < ndb) v = expand16(rows + src * Wr, 16 * c, b);
        *reinterpret_cast<uint4*>(out + pos * KP + 16 * c) = v;
    }
}

// threshold columns: qx[slot] = 32 int8, the first four sum to 2 T - b = -(b - 2 T), each part within [-64, 64] for
// b <= 256 (never-hit rows: 4 x -128 = -512 < -b); bx = 128 rows of (1, 1, 1, 1, 0, ...).
// ip' = ip + (2 T - b) >= 0  <=>  d_H <= T.
__global__ void __launch_bounds__(256) thr_columns_kernel(const int* __restrict__ thr, int64_t nq, int64_t nq_pad, int b, uint8_t* __restrict__ qx,
                                                          uint8_t* __restrict__ bx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq_pad) {
        uint32_t word = 0x80808080u;
        if (i < nq) {
            const int T = thr[i];
            if (T >= 0) {
                const int v = 2 * T - b;  // in [-b, b]; floor((v + k) / 4), k = 0..3, sum to v
                word = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) word |= (uint32_t)(((v + k) >> 2) & 0xFF) << (8 * k);
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(qx + i * 32);
        dst[0] = make_uint4(word, 0, 0, 0);
        dst[1] = make_uint4(0, 0, 0, 0);
    }
    if (i < 128) {
        uint4* dst = reinterpret_cast<uint4*>(bx + i * 32);
        dst[0] = make_uint4(0x01010101u, 0, 0, 0);
        dst[1] = make_uint4(0, 0, 0, 0);
    }
}

// tcgen05.ld of 32 accumulator columns as 16 registers: register j = (column 2j+1 low half) << 16 | (column 2j low half)
__device__ __forceinline__ void tmem_ld_32cols_pack16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// bytes 1 and 3 of a and of b, each replaced by 8 copies of its sign bit (PRMT sign-replicate mode)
__device__ __forceinline__ uint32_t sign_bytes(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// 32 accumulator columns -> 32-bit mask of the NON-NEGATIVE ones, the hits (bit 8k+i = column 4i+k).
// p_i = the four sign-replicated bytes (0xFF = negative) of columns 4i .. 4i+3.  The eight words are merged on the FMA pipe
// instead of the (binding) ALU pipe: a byte of ones is 2^8 - 1, so p_i = 255 * M_i with M_i the 0x01-per-byte word of the
// negative columns, and X = sum_i p_i << i = 255 * M (mod 2^32) with M the mask of the negative columns.  255^-1 mod 2^32 is
// -0x01010101, hence M = -X * 0x01010101 and the hit mask ~M = -M - 1 = X * 0x01010101 - 1: eight multiply-adds in all.
__device__ __forceinline__ uint32_t hit_mask32(uint32_t taddr, uint32_t sel)
{
    uint32_t r[16];
    tmem_ld_32cols_pack16(taddr, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t x = sign_bytes(r[0], r[1], sel);
#pragma unroll
    for (int i = 1; i < 8; ++i) x += sign_bytes(r[2 * i], r[2 * i + 1], sel) << i;
    return x * 0x01010101u - 1u;
}

// One CTA = 256 queries x TWO adjacent database splits (bins).  A 128-row B tile holds 64 rows of the first split
// (accumulator columns 0..63) and 64 rows of the second (columns 64..127); every epilogue warp owns 32 queries x one
// split, so 16 epilogue warps per CTA (32 per SM) share the CUDA-core work and each bin still has a single writer
// that sees its rows in ascending order.
// MODE fixes the packed-row shape at compile time: 1 = two code words in 4-word rows (32 < b <= 64, L <= 64: C4),
// 2 = four code words in 8-word rows (96 < b <= 128: C5), 0 = read it from the arguments.
// KP = 256 (b > 128): an operand row is two 128-byte K blocks, stored block after block ([block][row][128 B], each block
// SWIZZLE_128B) and contracted by 2 x 4 K steps; one CTA per SM.  KP = 32 (b <= 32): SWIZZLE_32B rows, one K step; 2-word
// packed rows travel as 16-byte row PAIRS (a.rows_paired) because a TMA box row must be a multiple of 16 bytes.
template <int KP, int MODE>
__global__ void __launch_bounds__(kUmmaThreads, UmmaCfg<KP>::CTAS)
select_umma_kernel(const __grid_constant__ CUtensorMap tmap_q8, const __grid_constant__ CUtensorMap tmap_db8,
                   const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_qx,
                   const __grid_constant__ CUtensorMap tmap_bx, UmmaSelectArgs a)
{
    constexpr int S = UmmaCfg<KP>::S;
    constexpr int KB = UmmaCfg<KP>::KB, KW = UmmaCfg<KP>::KW, QW = UmmaCfg<KP>::QW;
    constexpr uint32_t A_BYTES = 2 * 128 * KP;
    constexpr uint32_t AX_BYTES = 2 * 128 * kUmmaXBytes;
    constexpr uint32_t BX_BYTES = 128 * kUmmaXBytes;
    constexpr uint32_t FIXED_BYTES = A_BYTES + AX_BYTES + BX_BYTES;
    constexpr uint32_t B_BYTES = 128 * KP;
    constexpr uint32_t STAGE_BYTES = B_BYTES + UmmaCfg<KP>::ROWS_BYTES;
    extern __shared__ __align__(1024) uint8_t usm[];
    const uint32_t base = (smem_u32(usm) + 1023u) & ~1023u;
    uint8_t* const base_ptr = usm + (base - smem_u32(usm));
    __shared__ __align__(8) uint64_t bars[2 * S + 5];  // one array: every barrier is (one opaque base register) + constant
    __shared__ uint32_t tmem_base_slot;
    uint64_t& a_full = bars[0];
    uint64_t* const full_bar = bars + 1;
    uint64_t* const empty_bar = bars + 1 + S;
    uint64_t* const tmem_full = bars + 2 * S + 1;   // [2]: one accumulator (128 queries x 128 rows) per query half,
    uint64_t* const tmem_empty = bars + 2 * S + 3;  // [2]  handed back and forth independently

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q0 = (int64_t)blockIdx.x * 256;
    const int split_a = a.split0 + 2 * (int)blockIdx.y;  // this CTA's bins: split_a and split_a + 1
    const int64_t rowa = (int64_t)split_a * a.SL;
    const int64_t rows_first = max((int64_t)0, min(a.SL, a.ndb - rowa));  // the first split is never shorter than the second
    const int ntiles = (int)((rows_first + kUmmaHalfRows - 1) / kUmmaHalfRows);
    const int WrK = MODE == 1 ? 4 : (MODE == 2 ? 8 : a.Wr);
    const uint32_t half_rows_bytes = (uint32_t)kUmmaHalfRows * WrK * 4;

    if (threadIdx.x == 0) {
        mbar_init(&a_full, 1);
        for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kUmmaEpiWarps); }
        for (int h = 0; h < 2; ++h) { mbar_init(&tmem_full[h], 1); mbar_init(&tmem_empty[h], kUmmaEpiWarps / 2); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&a_full, FIXED_BYTES);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int kb = 0; kb < KB; ++kb)  // operand of query half hh: [K block][128 rows][KW bytes]
                    tma_load_2d(base + (uint32_t)((hh * KB + kb) * 128 * KW), &tmap_q8, smem_u32(&a_full), kb * KW, (int)q0 + hh * 128);
            tma_load_2d(base + A_BYTES, &tmap_qx, smem_u32(&a_full), 0, (int)q0);
            tma_load_2d(base + A_BYTES + 128 * kUmmaXBytes, &tmap_qx, smem_u32(&a_full), 0, (int)q0 + 128);
            tma_load_2d(base + A_BYTES + AX_BYTES, &tmap_bx, smem_u32(&a_full), 0, 0);
            tma_load_2d(base + A_BYTES + AX_BYTES + 64 * kUmmaXBytes, &tmap_bx, smem_u32(&a_full), 0, 64);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % S;
                mbar_wait(&empty_bar[s], (uint32_t)(((t / S) & 1) ^ 1));
                mbar_arrive_expect_tx(&full_bar[s], B_BYTES + 2 * half_rows_bytes);
                const uint32_t sb = base + FIXED_BYTES + s * STAGE_BYTES;
                const int ra = (int)(rowa + (int64_t)t * kUmmaHalfRows);  // rows of the first split; the second one starts SL rows later
                const int rb = (int)(rowa + a.SL + (int64_t)t * kUmmaHalfRows);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {  // B tile: [K block][64 rows of split a | 64 rows of split b][KW bytes]
                    tma_load_2d(sb + (uint32_t)(kb * 128 * KW), &tmap_db8, smem_u32(&full_bar[s]), kb * KW, ra);
                    tma_load_2d(sb + (uint32_t)(kb * 128 * KW + 64 * KW), &tmap_db8, smem_u32(&full_bar[s]), kb * KW, rb);
                }
                const int rsh = a.rows_paired ? 1 : 0;  // 2-word rows are boxed as 16-byte pairs: ra, rb are even (tile multiples)
                tma_load_2d(sb + B_BYTES, &tmap_rows, smem_u32(&full_bar[s]), 0, ra >> rsh);
                tma_load_2d(sb + B_BYTES + half_rows_bytes, &tmap_rows, smem_u32(&full_bar[s]), 0, rb >> rsh);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(128, 128);
            mbar_wait(&a_full, 0);
            const uint64_t bxdesc = umma_desc_kmajor(base + A_BYTES + AX_BYTES, kUmmaXBytes);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % S;
                mbar_wait(&full_bar[s], (uint32_t)((t / S) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_addr = base + FIXED_BYTES + s * STAGE_BYTES;
                // each query half as soon as ITS eight epilogue warps have read the previous tile's accumulator
                uint32_t todo = 3u;
                while (todo) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (!(todo & (1u << h))) continue;
                        if (todo == (1u << h)) mbar_wait(&tmem_empty[h], (uint32_t)((t & 1) ^ 1));
                        else if (!mbar_try(&tmem_empty[h], (uint32_t)((t & 1) ^ 1))) continue;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                        for (int kb = 0; kb < KB; ++kb) {
                            const uint64_t adesc = umma_desc_kmajor(base + (uint32_t)((h * KB + kb) * 128 * KW), KW);
                            const uint64_t bdesc = umma_desc_kmajor(b_addr + (uint32_t)(kb * 128 * KW), KW);
#pragma unroll
                            for (int k = 0; k < KW / 32; ++k)  // UMMA_K = 32 int8 = 32 bytes: start address advances by 2 (x16 B)
                                umma_i8(tmem_base + (uint32_t)(h * 128), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                        }
                        // threshold K step: ip' = ip + (2 T_q - b)
                        umma_i8(tmem_base + (uint32_t)(h * 128), umma_desc_kmajor(base + A_BYTES + h * 128 * kUmmaXBytes, kUmmaXBytes), bxdesc, idesc, 1u);
                        umma_commit(&tmem_full[h]);
                        todo &= ~(1u << h);
                    }
                }
            }
        }
    } else {
        // ---- epilogue: thread <-> (query, split) ---------------------------------------------------------
        const int e = warp - 2;
        const int quarter = warp & 3;    // TMEM lane quarter this warp may read
        const int h = (e >> 2) & 1;      // which 128-query half (A operand)
        const int ch = e >> 3;           // which split of the pair (accumulator columns 64 ch .. 64 ch + 63)
        const int64_t slot = q0 + h * 128 + quarter * 32 + lane;
        const int split = split_a + ch;
        const int64_t row0 = (int64_t)split * a.SL;
        const int64_t nrows = max((int64_t)0, min(a.SL, a.ndb - row0));
        const bool valid = slot < a.nq && split < a.P;
        const int W = MODE == 1 ? 2 : (MODE == 2 ? 4 : a.W), LW = a.LW, Wr = WrK;
        const uint32_t sel = a.prmt_sel;
        uint32_t qw[QW], ql[4] = {0, 0, 0, 0};
#pragma unroll
        for (int w = 0; w < QW; ++w) qw[w] = 0;
        uint32_t pos = 0, nback = 0, start = 0, end = 0;  // front stack (d < T) grows up from start, back stack (d == T) down from end
        int Tq = -1;
        int64_t bin = -1;
        if (valid) {
#pragma unroll
            for (int w = 0; w < QW; ++w)
                if (w < W) qw[w] = a.q_rows[slot * Wr + w];
            for (int w = 0; w < LW && w < 4; ++w) ql[w] = a.q_rows[slot * Wr + W + w];
            Tq = a.thr[slot];
            bin = slot * a.P + split;
            start = (uint32_t)(bin * (int64_t)a.cap);
            end = start + a.cap;
            pos = start;
        }
        const uint32_t live = (valid && Tq >= 0) ? 0xFFFFFFFFu : 0u;
        uint32_t* const lists = a.lists;
        const uint32_t tmem_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(h * 128 + ch * 64);
        uint32_t bar0 = smem_u32(&bars[0]);
        asm volatile("" : "+r"(bar0));  // keep the shared-window address in a register instead of re-deriving it per tile
        const uint32_t full_a = bar0 + 8u, empty_a = bar0 + 8u * (1 + S);
        const uint32_t tfull_a = bar0 + 8u * (2 * S + 1 + h), tempty_a = bar0 + 8u * (2 * S + 3 + h);
        const int nrows32 = (int)nrows;                 // <= 2^21
        const int nfull = nrows32 / kUmmaHalfRows;      // tiles in which all 64 rows of my split exist
        int s = 0;
        uint32_t ph = 0, tph = 0;                       // ring-stage parity, accumulator parity
        const uint8_t* stage_rows = base_ptr + FIXED_BYTES + B_BYTES + ch * half_rows_bytes;
        for (int t = 0; t < ntiles; ++t) {
            mbar_wait_a(full_a + 8u * s, ph);   // packed rows of tile t have landed (observed by this thread)
            mbar_wait_a(tfull_a, tph);          // MMAs of tile t are complete
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t* srows = reinterpret_cast<const uint32_t*>(stage_rows + s * STAGE_BYTES);
            const uint32_t m0 = hit_mask32(tmem_row, sel);
            const uint32_t m1 = hit_mask32(tmem_row + 32u, sel);
            // the accumulators are consumed: let the MMA warp start the next tile while the hits are written out
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_a(tempty_a);
            uint32_t h0 = m0 & live, h1 = m1 & live;
            if (t >= nfull) {  // last tile(s) of the split: drop the rows that do not exist
                const int left = nrows32 - t * kUmmaHalfRows;
                h0 &= left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
                h1 &= left <= 32 ? 0u : ((1u << (left - 32)) - 1u);
            }
            // rare path: about R/Ndb of the pairs, ascending row order (mask bit j of word g = row 32 g + j of the tile half)
            while (h0 | h1) {
                const bool first = h0 != 0u;
                const uint32_t hm = first ? h0 : h1;
                const int rl = (first ? 0 : 32) + (__ffs((int)hm) - 1);
                const uint32_t cleared = hm & (hm - 1u);
                if (first) h0 = cleared; else h1 = cleared;
                const uint32_t* prow = srows + rl * Wr;
                if (KP == 32 && a.rows_paired && row0 + (int64_t)t * kUmmaHalfRows + rl == a.ndb - 1)
                    prow = a.db_rows + (a.ndb - 1) * Wr;  // odd database size: the last row has no pair partner in the TMA box
                int d = 0;
                uint32_t m = 0;
                if (MODE == 1) {
                    const uint4 pr = *reinterpret_cast<const uint4*>(prow);
                    d = __popc(qw[0] ^ pr.x) + __popc(qw[1] ^ pr.y);
                    m = (ql[0] & pr.z) | (ql[1] & pr.w);
                } else if (MODE == 2) {
                    const uint4 pc = *reinterpret_cast<const uint4*>(prow), pl = *reinterpret_cast<const uint4*>(prow + 4);
                    d = __popc(qw[0] ^ pc.x) + __popc(qw[1] ^ pc.y) + __popc(qw[2] ^ pc.z) + __popc(qw[3] ^ pc.w);
                    m = (ql[0] & pl.x) | (ql[1] & pl.y) | (ql[2] & pl.z) | (ql[3] & pl.w);
                } else if (Wr == 4) {  // one 16-byte load brings the code words and the label word(s): W = 1 (+ <= 3 label words), 2 (+ <= 2) or 3 (+ 1)
                    const uint4 pr = *reinterpret_cast<const uint4*>(prow);
                    d = __popc(qw[0] ^ pr.x);
                    if (W == 1) { m = (ql[0] & pr.y) | (ql[1] & pr.z) | (ql[2] & pr.w); }
                    else if (W == 3) { d += __popc(qw[1] ^ pr.y) + __popc(qw[2] ^ pr.z); m = ql[0] & pr.w; }
                    else { d += __popc(qw[1] ^ pr.y); m = (ql[0] & pr.z) | (ql[1] & pr.w); }
                } else {
#pragma unroll
                    for (int w = 0; w < QW; ++w)
                        if (w < W) d += __popc(qw[w] ^ prow[w]);
#pragma unroll
                    for (int w = 0; w < 4; ++w)
                        if (w < LW) m |= ql[w] & prow[W + w];
                }
                const bool eq = d == Tq;
                if (pos + nback < end)
                    lists[eq ? end - 1u - nback : pos] = ((uint32_t)d * (1u << kIdxBits) + (uint32_t)(t * kUmmaHalfRows + rl)) | (m ? 0x80000000u : 0u);
                if (eq) nback += 1; else pos += 1;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(empty_a + 8u * s);  // shared-memory stage free for the producer
            tph ^= 1u;
            if (++s == S) { s = 0; ph ^= 1u; }
        }
        if (bin >= 0) { a.bin_cnt[bin] = pos - start; a.bin_cnt0[bin] = nback; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// ---- queued mode: the list-appending epilogue without its 23 % lane occupancy -----------------------------------------------
// select_umma_kernel walks the hits of a tile before it touches the next one, so every warp-tile costs max-over-lanes hit
// steps (2.5 at C4 for 0.57 hits per lane).  Here a lane only PARKS its non-zero mask words in a private shared-memory FIFO
// (word + tile tag) and per tile the warp consumes ONE parked hit per lane, in two halves around the accumulator hand-off:
//   take()     before the wait: pop / pick the lowest row of the word being consumed, REQUEST its packed row from global memory
//              (L2-resident; hits may be consumed many tiles late, so the staged tile is gone -- the ring therefore holds only
//              the int8 operands and is released by the MMA warp itself, tcgen05.commit);
//   consume()  after the masks of the tile are built and the accumulator is handed back: distance, relevance, append.
// The row's L2 latency hides behind the wait and the mask building (loading it inside the hit step put ~700 cycles on the
// hand-off path: no gain over the tile-walking kernel), the lanes no longer wait for each other at tile boundaries, and a
// lane's backlog is smoothed over the following tiles.  A FIFO about to fill, or a long wait with many lanes holding work,
// triggers extra (blocking) steps.  Measured at C4: 1.306 -> 1.125 ms, 1085 M -> 792 M warp instructions.  Bins, entry format,
// order (one writer per bin, ascending rows) and the AP kernel are those of select_umma_kernel.
template <int KP> struct QCfg {
    static constexpr int S = (KP == 32 ? 8 : (KP == 64 ? 4 : 3));  // ring depth
    static constexpr int D = (KP == 128 ? 4 : 8);                  // FIFO entries per lane (power of two)
};

template <int KP, int MODE>
__global__ void __launch_bounds__(kUmmaThreads, UmmaCfg<KP>::CTAS)
select_q_kernel(const __grid_constant__ CUtensorMap tmap_q8, const __grid_constant__ CUtensorMap tmap_db8, const __grid_constant__ CUtensorMap tmap_qx,
                const __grid_constant__ CUtensorMap tmap_bx, UmmaSelectArgs a)
{
    constexpr int S = QCfg<KP>::S, D = QCfg<KP>::D;
    constexpr int NBUF = 1;          // accumulators per query half (two 64-column accumulators per half -- 128 x 64 MMAs -- measured: 1.35 vs 1.14 ms)
    constexpr uint32_t ACC = 128u;   // tensor-memory columns of one 128 x 128 accumulator
    constexpr int KB = UmmaCfg<KP>::KB, KW = UmmaCfg<KP>::KW, QW = UmmaCfg<KP>::QW;
    constexpr uint32_t A_BYTES = 2 * 128 * KP;
    constexpr uint32_t AX_BYTES = 2 * 128 * kUmmaXBytes;
    constexpr uint32_t BX_BYTES = 128 * kUmmaXBytes;
    constexpr uint32_t FIXED_BYTES = A_BYTES + AX_BYTES + BX_BYTES;
    constexpr uint32_t STAGE_BYTES = 128 * KP;
    extern __shared__ __align__(1024) uint8_t usm[];
    const uint32_t base = (smem_u32(usm) + 1023u) & ~1023u;
    uint8_t* const base_ptr = usm + (base - smem_u32(usm));
    __shared__ __align__(8) uint64_t bars[2 * S + 1 + 4 * NBUF];
    __shared__ uint32_t tmem_base_slot;
    uint64_t& a_full = bars[0];
    uint64_t* const full_bar = bars + 1;
    uint64_t* const empty_bar = bars + 1 + S;
    uint64_t* const tmem_full = bars + 2 * S + 1;              // [2 halves][NBUF]
    uint64_t* const tmem_empty = bars + 2 * S + 1 + 2 * NBUF;  // [2 halves][NBUF]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q0 = (int64_t)blockIdx.x * 256;
    const int split_a = a.split0 + 2 * (int)blockIdx.y;
    const int64_t rowa = (int64_t)split_a * a.SL;
    const int64_t rows_first = max((int64_t)0, min(a.SL, a.ndb - rowa));
    const int ntiles = (int)((rows_first + kUmmaHalfRows - 1) / kUmmaHalfRows);

    if (threadIdx.x == 0) {
        mbar_init(&a_full, 1);
        for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int h = 0; h < 2 * NBUF; ++h) { mbar_init(&tmem_full[h], 1); mbar_init(&tmem_empty[h], kUmmaEpiWarps / 2); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&a_full, FIXED_BYTES);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d(base + (uint32_t)((hh * KB + kb) * 128 * KW), &tmap_q8, smem_u32(&a_full), kb * KW, (int)q0 + hh * 128);
            tma_load_2d(base + A_BYTES, &tmap_qx, smem_u32(&a_full), 0, (int)q0);
            tma_load_2d(base + A_BYTES + 128 * kUmmaXBytes, &tmap_qx, smem_u32(&a_full), 0, (int)q0 + 128);
            tma_load_2d(base + A_BYTES + AX_BYTES, &tmap_bx, smem_u32(&a_full), 0, 0);
            tma_load_2d(base + A_BYTES + AX_BYTES + 64 * kUmmaXBytes, &tmap_bx, smem_u32(&a_full), 0, 64);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % S;
                mbar_wait(&empty_bar[s], (uint32_t)(((t / S) & 1) ^ 1));
                mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                const uint32_t sb = base + FIXED_BYTES + s * STAGE_BYTES;
                const int ra = (int)(rowa + (int64_t)t * kUmmaHalfRows);
                const int rb = (int)(rowa + a.SL + (int64_t)t * kUmmaHalfRows);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    tma_load_2d(sb + (uint32_t)(kb * 128 * KW), &tmap_db8, smem_u32(&full_bar[s]), kb * KW, ra);
                    tma_load_2d(sb + (uint32_t)(kb * 128 * KW + 64 * KW), &tmap_db8, smem_u32(&full_bar[s]), kb * KW, rb);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(128, 128);
            mbar_wait(&a_full, 0);
            const uint64_t bxdesc = umma_desc_kmajor(base + A_BYTES + AX_BYTES, kUmmaXBytes);
            auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t acc) { umma_i8(d, ad, bd, idesc, acc); };
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % S;
                const int pb = 0;                 // accumulator buffer of this tile (NBUF = 1)
                const uint32_t use = (uint32_t)t;  // how often that buffer was filled before
                mbar_wait(&full_bar[s], (uint32_t)((t / S) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_addr = base + FIXED_BYTES + s * STAGE_BYTES;
                uint32_t todo = 3u;
                while (todo) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (!(todo & (1u << h))) continue;
                        uint64_t* const te = &tmem_empty[h * NBUF + pb];
                        if (todo == (1u << h)) mbar_wait(te, (use & 1u) ^ 1u);
                        else if (!mbar_try(te, (use & 1u) ^ 1u)) continue;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t dcol = tmem_base + (uint32_t)(h * NBUF + pb) * ACC;
#pragma unroll
                        for (int kb = 0; kb < KB; ++kb) {
                            const uint64_t adesc = umma_desc_kmajor(base + (uint32_t)((h * KB + kb) * 128 * KW), KW);
                            const uint64_t bdesc = umma_desc_kmajor(b_addr + (uint32_t)(kb * 128 * KW), KW);
#pragma unroll
                            for (int k = 0; k < KW / 32; ++k)
                                mma(dcol, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), (uint32_t)((kb | k) != 0));
                        }
                        mma(dcol, umma_desc_kmajor(base + A_BYTES + h * 128 * kUmmaXBytes, kUmmaXBytes), bxdesc, 1u);
                        umma_commit(&tmem_full[h * NBUF + pb]);
                        todo &= ~(1u << h);
                    }
                }
                umma_commit(&empty_bar[s]);  // the operands of this stage are free once both MMAs retire
            }
        }
    } else {
        // ---- epilogue: thread <-> (query, split) = one bin -------------------------------------------------------
        const int e = warp - 2;
        const int quarter = warp & 3;
        const int h = (e >> 2) & 1;
        const int ch = e >> 3;
        const int64_t slot = q0 + h * 128 + quarter * 32 + lane;
        const int split = split_a + ch;
        const int64_t row0 = (int64_t)split * a.SL;
        const int64_t nrows = max((int64_t)0, min(a.SL, a.ndb - row0));
        const bool valid = slot < a.nq && split < a.P;
        const int W = MODE == 1 ? 2 : (MODE == 2 ? 4 : a.W), LW = a.LW, Wr = MODE == 1 ? 4 : (MODE == 2 ? 8 : a.Wr);
        const uint32_t sel = a.prmt_sel;
        uint32_t qw[QW], ql[4] = {0, 0, 0, 0};
#pragma unroll
        for (int w = 0; w < QW; ++w) qw[w] = 0;
        uint32_t pos = 0, bpos = 0, start = 0, end = 0;  // front stack (d < T): next slot pos, grows up from start; back stack (d == T): next slot bpos, grows down from end - 1
        int Tq = -1;
        int64_t bin = -1;
        if (valid) {
#pragma unroll
            for (int w = 0; w < QW; ++w)
                if (w < W) qw[w] = a.q_rows[slot * Wr + w];
            for (int w = 0; w < LW && w < 4; ++w) ql[w] = a.q_rows[slot * Wr + W + w];
            Tq = a.thr[slot];
            bin = slot * a.P + split;
            start = (uint32_t)(bin * (int64_t)a.cap);
            end = start + a.cap;
            pos = start;
            bpos = end - 1u;
        }
        // (no lane mask: queries beyond nq and queries without a threshold carry never-hit threshold columns, rows beyond the split are masked below)
        uint32_t* const lists = a.lists;
        const uint32_t* const split_rows = a.db_rows + (valid ? row0 : 0) * Wr;  // packed rows of my split (global, L2-resident)
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)ch * (ACC / 2u);
        uint32_t bar0 = smem_u32(&bars[0]);
        asm volatile("" : "+r"(bar0));
        const uint32_t tfull_0 = bar0 + 8u * (2 * S + 1 + h * NBUF), tempty_0 = bar0 + 8u * (2 * S + 1 + 2 * NBUF + h * NBUF);
        const int nrows32 = (int)nrows;
        const int nfull = nrows32 / kUmmaHalfRows;
        // private FIFO of parked mask words: entry i of my lane at fifo[i * 32] = (word, 2 * tile + word index)
        // (rd, wr count in bytes of one ring row: 32 lanes x 8 bytes)
        uint8_t* const fifo = base_ptr + FIXED_BYTES + S * STAGE_BYTES + ((size_t)e * (D * 32) + lane) * sizeof(uint2);
        constexpr uint32_t ROW = 32u * (uint32_t)sizeof(uint2), RING = (uint32_t)(D - 1) * ROW;
        uint32_t rd = 0, wr = 0, cur = 0, curtag = 0;
        const uint32_t FULLM = 0xFFFFFFFFu;
        const int min_lanes = a.drain_lanes;

        // A hit is consumed in two halves so that the L2 latency of its packed row is never waited for:
        //   take():    the lowest row of the word being consumed (rows ascend within a word, words ascend in the FIFO) -> request the row
        //   consume(): distance, relevance, append -- one call later, when the row has long arrived
        bool pv = false;       // a requested row is pending
        uint32_t prl = 0;      // its row inside the split
        uint4 pa = make_uint4(0, 0, 0, 0), pb = make_uint4(0, 0, 0, 0);  // its packed words (MODE 1: pa, MODE 2: pa + pb)
        auto take = [&]() {
            if (cur == 0u && rd != wr) {
                const uint2 en = *reinterpret_cast<const uint2*>(fifo + (rd & RING));
                cur = en.x; curtag = en.y;
                rd += ROW;
            }
            if (cur != 0u) {
                prl = curtag * 32u + (uint32_t)(__ffs((int)cur) - 1);
                cur &= cur - 1u;
                pv = true;
                const uint32_t* prow = split_rows + (size_t)prl * Wr;
                if (MODE == 1) pa = __ldg(reinterpret_cast<const uint4*>(prow));
                if (MODE == 2) { pa = __ldg(reinterpret_cast<const uint4*>(prow)); pb = __ldg(reinterpret_cast<const uint4*>(prow + 4)); }
            }
        };
        auto consume = [&]() {
            if (pv) {
                int d = 0;
                uint32_t m = 0;
                if (MODE == 1) {
                    d = __popc(qw[0] ^ pa.x) + __popc(qw[1] ^ pa.y);
                    m = (ql[0] & pa.z) | (ql[1] & pa.w);
                } else if (MODE == 2) {
                    d = __popc(qw[0] ^ pa.x) + __popc(qw[1] ^ pa.y) + __popc(qw[2] ^ pa.z) + __popc(qw[3] ^ pa.w);
                    m = (ql[0] & pb.x) | (ql[1] & pb.y) | (ql[2] & pb.z) | (ql[3] & pb.w);
                } else {
                    const uint32_t* prow = split_rows + (size_t)prl * Wr;
#pragma unroll
                    for (int w = 0; w < QW; ++w)
                        if (w < W) d += __popc(qw[w] ^ __ldg(prow + w));
#pragma unroll
                    for (int w = 0; w < 4; ++w)
                        if (w < LW) m |= ql[w] & __ldg(prow + W + w);
                }
                const bool eq = d == Tq;
                const uint32_t idx = eq ? bpos : pos;
                if ((int32_t)(bpos - pos) >= 0) lists[idx] = ((uint32_t)d * (1u << kIdxBits) + prl) | (m ? 0x80000000u : 0u);  // room left: pos <= bpos
                if (eq) bpos -= 1; else pos += 1;
                pv = false;
            }
        };

        for (int t = 0; t < ntiles; ++t) {
            const uint32_t pb = 0u;
            const uint32_t tph = (uint32_t)t & 1u;
            const uint32_t tfull_a = tfull_0 + 8u * pb, tempty_a = tempty_0 + 8u * pb;
            const uint32_t tmem_row = tmem_lane + (uint32_t)(h * NBUF + (int)pb) * ACC;
            // request the next parked hit of every lane, then wait for the accumulator of tile t; a long wait with a large backlog
            // (>= min_lanes lanes with work) is spent on further hits
            take();  // (nothing is pending here: the previous iteration consumed what it had requested)
            while (!__all_sync(FULLM, mbar_try_a(tfull_a, tph))) {
                const uint32_t pend = __ballot_sync(FULLM, (cur != 0u) | (rd != wr));
                if (__popc(pend) >= min_lanes) { consume(); take(); }
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t m0 = hit_mask32(tmem_row, sel);
            const uint32_t m1 = hit_mask32(tmem_row + 32u, sel);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_a(tempty_a);
            uint32_t h0 = m0, h1 = m1;
            if (t >= nfull) {
                const int left = nrows32 - t * kUmmaHalfRows;
                h0 &= left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
                h1 &= left <= 32 ? 0u : ((1u << (left - 32)) - 1u);
            }
            consume();  // the row requested before the wait
            // room for two more words in every FIFO
            while (__any_sync(FULLM, wr - rd > (uint32_t)(D - 2) * ROW)) { take(); consume(); }
            if (h0) { *reinterpret_cast<uint2*>(fifo + (wr & RING)) = make_uint2(h0, 2u * (uint32_t)t); wr += ROW; }
            if (h1) { *reinterpret_cast<uint2*>(fifo + (wr & RING)) = make_uint2(h1, 2u * (uint32_t)t + 1u); wr += ROW; }
        }
        consume();
        while (__any_sync(FULLM, (cur != 0u) | (rd != wr))) { take(); consume(); }
        if (bin >= 0) { a.bin_cnt[bin] = pos - start; a.bin_cnt0[bin] = end - 1u - bpos; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
// [rows, row_bytes] bytes, boxes of box_rows x box_bytes (box_bytes = the swizzle width: 32 / 64 / 128; row_bytes = 256
// is read as two 128-byte column blocks)
static int make_map_u8(CUtensorMap* map, const uint8_t* ptr, int64_t rows, int row_bytes, int box_rows, int box_bytes = 0)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(HG_ECUDA, "select_umma: cuTensorMapEncodeTiled is not available from the driver");
    if (box_bytes == 0) box_bytes = row_bytes;
    cuuint64_t gdim[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_bytes, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = box_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (box_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HG_ECUDA, "select_umma: cuTensorMapEncodeTiled(u8, %d B rows) failed (%d)", row_bytes, (int)r);
    return HG_OK;
}

// packed rows [rows, Wr] words, boxes of 64 rows; 2-word rows (8 bytes, below the 16-byte minimum of a box row) are
// described as rows/2 PAIRS of 4 words, boxes of 32 pairs (an odd last row is read from global memory by the kernel)
static int make_map_rows(CUtensorMap* map, const uint32_t* ptr, int64_t rows, int Wr)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(HG_ECUDA, "select_umma: cuTensorMapEncodeTiled is not available from the driver");
    int box_rows = kUmmaHalfRows;
    if (Wr == 2) { Wr = 4; rows = std::max<int64_t>(1, rows / 2); box_rows = kUmmaHalfRows / 2; }
    cuuint64_t gdim[2] = {(cuuint64_t)Wr, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)Wr * 4};
    cuuint32_t box[2] = {(cuuint32_t)Wr, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HG_ECUDA, "select_umma: cuTensorMapEncodeTiled(rows) failed (%d)", (int)r);
    return HG_OK;
}

int umma_select_kp(int b, int Wr)
{
    // int8 row bytes; 0 = this shape stays on the POPC kernel (a packed-row stride that TMA cannot box: 3- and 5-word rows,
    // i.e. b <= 32 with more than 32 labels ... 64 or more than 96)
    if (b <= 0 || b > 256) return 0;
    if (b <= 32) return (Wr == 2 || Wr == 4) ? 32 : 0;
    if (b <= 128) return (Wr == 4 || Wr == 8) ? (b <= 64 ? 64 : 128) : 0;
    return Wr == 12 ? 256 : 0;
}

static unsigned grid_for(int64_t threads_needed)
{
    const int sms = device_facts().sm_count > 0 ? device_facts().sm_count : 148;
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(threads_needed, 256), (int64_t)sms * 16));
}

int umma_expand_q(const uint32_t* rows, int64_t n, int b, int Wr, int KP, uint8_t* out, cudaStream_t st)
{
    if (n <= 0) return HG_OK;
    expand_q_kernel<<<grid_for(n * (KP / 16)), 256, 0, st>>>(rows, n, b, Wr, KP, out);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

int umma_expand_db(const uint32_t* db_rows, int64_t lo, int64_t hi, int64_t ndb, int b, int Wr, int KP, uint8_t* db8, cudaStream_t st)
{
    if (hi <= lo) return HG_OK;
    if ((lo & 31) || ((hi & 31) && hi != ndb)) return fail(HG_EINVAL, "select_umma: database chunks must be whole 32-row groups");
    const int64_t hi_pad = round_up(hi, 32);
    expand_db_kernel<<<grid_for((hi_pad - lo) * (KP / 16)), 256, 0, st>>>(db_rows, lo, hi_pad, ndb, b, Wr, KP, db8);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

int umma_thr_columns(const int* thr, int64_t nq, int b, uint8_t* qx, uint8_t* bx, cudaStream_t st)
{
    const int64_t nq_pad = std::max<int64_t>(round_up(nq, 256), 128);
    thr_columns_kernel<<<(unsigned)ceil_div(nq_pad, 256), 256, 0, st>>>(thr, nq, nq_pad, b, qx, bx);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

template <int KP, int MODE>
static int launch_umma(const CUtensorMap& tq, const CUtensorMap& tdb, const CUtensorMap& trows, const CUtensorMap& tqx, const CUtensorMap& tbx,
                       const UmmaSelectArgs& a, cudaStream_t st)
{
    const size_t smem = (size_t)2 * 128 * KP + (size_t)3 * 128 * kUmmaXBytes + (size_t)UmmaCfg<KP>::S * (128 * KP + UmmaCfg<KP>::ROWS_BYTES) + 1024;
    static thread_local bool configured = false;
    if (!configured) {
        HG_CUDA_TRY(cudaFuncSetAttribute(select_umma_kernel<KP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((unsigned)ceil_div(a.nq, 256), (unsigned)ceil_div(a.n_splits, 2));
    select_umma_kernel<KP, MODE><<<grid, kUmmaThreads, smem, st>>>(tq, tdb, trows, tqx, tbx, a);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

template <int KP, int MODE>
static int launch_q(const CUtensorMap& tq, const CUtensorMap& tdb, const CUtensorMap& tqx, const CUtensorMap& tbx, const UmmaSelectArgs& a, cudaStream_t st)
{
    const size_t smem = (size_t)2 * 128 * KP + (size_t)3 * 128 * kUmmaXBytes + (size_t)QCfg<KP>::S * (128 * KP) + (size_t)kUmmaEpiWarps * QCfg<KP>::D * 32 * sizeof(uint2) + 1024;
    static thread_local bool configured = false;
    if (!configured) {
        HG_CUDA_TRY(cudaFuncSetAttribute(select_q_kernel<KP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((unsigned)ceil_div(a.nq, 256), (unsigned)ceil_div(a.n_splits, 2));
    select_q_kernel<KP, MODE><<<grid, kUmmaThreads, smem, st>>>(tq, tdb, tqx, tbx, a);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

int umma_select_launch(const UmmaSelectArgs& a_in, cudaStream_t st)
{
    UmmaSelectArgs a = a_in;
    if (a.prmt_sel == 0) {
        a.prmt_sel = 0xFDB9u;  // bytes 1, 3 of the first register, bytes 1, 3 of the second, sign-replicated
        const char* v = getenv("HG_UMMA_PRMT");  // probe override (hex)
        if (v && *v) a.prmt_sel = (uint32_t)strtoul(v, nullptr, 16);
    }
    CUtensorMap tq, tdb, trows, tqx, tbx;
    int rc;
    if (a.split0 & 1) return fail(HG_EINVAL, "select_umma: a launch must start at an even split");
    const int kw = a.KP > 128 ? 128 : a.KP;
    a.rows_paired = a.Wr == 2 ? 1 : 0;
    if ((rc = make_map_u8(&tq, a.q8, a.nq, a.KP, 128, kw)) != HG_OK) return rc;
    if ((rc = make_map_u8(&tdb, a.db8, round_up(a.ndb, 32), a.KP, kUmmaHalfRows, kw)) != HG_OK) return rc;
    if ((rc = make_map_u8(&tqx, a.qx, std::max<int64_t>(round_up(a.nq, 256), 128), kUmmaXBytes, 128)) != HG_OK) return rc;
    if ((rc = make_map_u8(&tbx, a.bx, 128, kUmmaXBytes, kUmmaHalfRows)) != HG_OK) return rc;
    if (a.queued && a.KP == 64 && a.W == 2 && a.Wr == 4 && (a.SL & 63) == 0) {  // parked mask words, hits consumed asynchronously (select_q_kernel)
        const char* v = getenv("HG_DRAIN_LANES");
        a.drain_lanes = (v && *v) ? atoi(v) : 12;
        return launch_q<64, 1>(tq, tdb, tqx, tbx, a, st);
    }
    if ((rc = make_map_rows(&trows, a.db_rows, a.ndb, a.Wr)) != HG_OK) return rc;
    if (a.KP == 32) return launch_umma<32, 0>(tq, tdb, trows, tqx, tbx, a, st);
    if (a.KP == 64) return (a.W == 2 && a.Wr == 4) ? launch_umma<64, 1>(tq, tdb, trows, tqx, tbx, a, st) : launch_umma<64, 0>(tq, tdb, trows, tqx, tbx, a, st);
    if (a.KP == 128) return (a.W == 4 && a.Wr == 8) ? launch_umma<128, 2>(tq, tdb, trows, tqx, tbx, a, st) : launch_umma<128, 0>(tq, tdb, trows, tqx, tbx, a, st);
    return launch_umma<256, 0>(tq, tdb, trows, tqx, tbx, a, st);
}


// ---- int8 tcgen05 peak (roofline denominator of the tensor view) -----------------------------------------------------
// One CTA per SM keeps its operands resident in shared memory (A 128 x 128 B, B 256 x 128 B, SWIZZLE_128B K-major) and one
// thread issues `iters` x 4 back-to-back tcgen05.mma kind::i8 (M = 128, N = 256, K = 32) alternating between two
// accumulators in tensor memory: nothing but the tensor pipe is exercised.
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters, uint32_t seed, uint32_t* __restrict__ sink)
{
    extern __shared__ __align__(1024) uint8_t psm[];
    const uint32_t base = (smem_u32(psm) + 1023u) & ~1023u;
    uint8_t* const ptr = psm + (base - smem_u32(psm));
    __shared__ __align__(8) uint64_t done_bar;
    __shared__ uint32_t tmem_slot;
    constexpr uint32_t A_BYTES = 128 * 128, B_BYTES = 256 * 128;
    for (uint32_t i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) {
        uint32_t x = (i + seed) * 2654435761u;
        x ^= x >> 15;
        reinterpret_cast<uint32_t*>(ptr)[i] = (x & 0x01010101u) * 0xFEu ^ 0xFFFFFFFFu;  // bytes of +1 / -1
    }
    fence_proxy_async();
    if (threadIdx.x == 0) { mbar_init(&done_bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = umma_idesc_i8(128, 256);
        const uint64_t adesc = umma_desc_kmajor(base, 128), bdesc = umma_desc_kmajor(base + A_BYTES, 128);
        for (int it = 0; it < iters; ++it) {
            const uint32_t acc = tmem + (uint32_t)((it & 1) * 256);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_i8(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)(it > 1 || k != 0));
        }
        umma_commit(&done_bar);
        mbar_wait(&done_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t r[16];
        tmem_ld_32cols_pack16(tmem, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (sink && r[0] == 0x12345678u && r[5] == 0x9ABCDEFu) sink[blockIdx.x] = r[1];  // keeps the accumulators observable
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

int i8_peak(double* ops_per_s, double* ms_out, int iters, cudaStream_t st)
{
    const int sms = device_facts().sm_count > 0 ? device_facts().sm_count : 148;
    if (iters <= 0) iters = 8192;
    const size_t smem = 128 * 128 + 256 * 128 + 1024;
    HG_CUDA_TRY(cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t* sink = nullptr;
    HG_CUDA_TRY(cudaMalloc(&sink, sizeof(uint32_t) * sms));
    cudaEvent_t e0, e1;
    HG_CUDA_TRY(cudaEventCreate(&e0));
    HG_CUDA_TRY(cudaEventCreate(&e1));
    float best = 1e30f;
    count_launch(4);
    i8_peak_kernel<<<sms, 128, smem, st>>>(iters / 4 + 1, 1u, sink);  // warm-up
    for (int rep = 0; rep < 3; ++rep) {
        HG_CUDA_TRY(cudaEventRecord(e0, st));
        i8_peak_kernel<<<sms, 128, smem, st>>>(iters, 7u + rep, sink);
        HG_CUDA_TRY(cudaEventRecord(e1, st));
        HG_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        HG_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    HG_CUDA_TRY(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *ops_per_s = 2.0 * 128.0 * 256.0 * 32.0 * 4.0 * (double)iters * (double)sms / ((double)best * 1e-3);
    if (ms_out) *ms_out = best;
    return HG_OK;
}

}  // namespace hg

extern "C" int hg_i8_peak(double* ops_per_s, double* ms_out, int iters, void* stream)
{
    if (!ops_per_s) return hg::fail(HG_EINVAL, "hg_i8_peak: bad arguments");
    if (!hg::device_facts().ok) return hg::fail(HG_ECUDA, "hg_i8_peak: no CUDA device");
    return hg::i8_peak(ops_per_s, ms_out, iters, (cudaStream_t)stream);
}

