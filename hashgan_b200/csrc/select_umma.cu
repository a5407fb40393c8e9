// Tensor-core variant of the hot kernel: the all-pairs contraction of lib/metric.py:13 as an EXACT int8 GEMM.
//
// On {-1,+1} codes  ip = sum_k q_k * db_k = b - 2 * d_H  (SURVEY A.3), so a row is a candidate (d_H <= T_q) iff
// ip >= b - 2 T_q.  The codes are expanded once to int8 (+1 / -1, zero padding) and contracted with
// tcgen05.mma kind::i8 (int32 accumulation in tensor memory: exact), which moves the contraction off the POPC pipe
// that bounds select_kernel.  The epilogue then only has to threshold the accumulators:
//
//   per CTA: 256 queries (two 128-row A operands, loaded once by TMA) x one database split
//   warp 0      TMA producer: per 128-row database tile the int8 tile (B operand, swizzled) and the packed rows
//               (code + label words, for the rare path) into a 4-stage ring
//   warp 1      TMEM allocator (256 columns = 2 query halves x 128; two CTAs share an SM) and MMA issuer
//   warps 2..9  epilogue: thread <-> query (TMEM lane).  tcgen05.ld 32 columns at a time; two integer ops per pair
//               build four 32-bit hit masks (sub + funnel shift of the sign bit), then the accumulator is handed
//               back to the MMA warp; the ~R/Ndb hits of the tile are walked in row order: distance recomputed
//               from the packed words (POPC), relevance from the label words, entry appended to the thread's
//               private bin exactly as select_kernel does.
// The bins, thresholds, AP kernel and exactness guard are shared with the POPC path (rank.cu).
#include "umma.cuh"

namespace hg {

constexpr int kUmmaThreads = 320;
template <int KP> struct UmmaCfg { static constexpr int S = (KP == 64 ? 4 : 3); };  // ring depth: two CTAs must fit one SM
constexpr int kUmmaTileRows = 128;
constexpr uint32_t kUmmaRowsMax = 128 * 8 * 4;  // packed-row bytes per stage (Wr <= 8)

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

// packed code words -> int8 rows [n, KP]: +1 where the bit is set, -1 where it is clear, 0 beyond bit b
__global__ void __launch_bounds__(256) expand_codes_kernel(const uint32_t* __restrict__ rows, int64_t n, int b, int Wr, int KP, uint8_t* __restrict__ out)
{
    const int cpr = KP / 16;  // 16-byte chunks per row
    const int64_t total = n * cpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / cpr;
        const int c = (int)(i - row * cpr);
        const int bit0 = 16 * c;
        uint32_t bits = 0;
        if (bit0 < b) bits = (__ldg(rows + row * Wr + (bit0 >> 5)) >> (bit0 & 31)) & 0xFFFFu;
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
        *reinterpret_cast<uint4*>(out + row * KP + 16 * c) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

template <int KP>
__global__ void __launch_bounds__(kUmmaThreads, 2)
select_umma_kernel(const __grid_constant__ CUtensorMap tmap_q8, const __grid_constant__ CUtensorMap tmap_db8,
                   const __grid_constant__ CUtensorMap tmap_rows, UmmaSelectArgs a)
{
    constexpr int S = UmmaCfg<KP>::S;
    constexpr uint32_t A_BYTES = 2 * 128 * KP;
    constexpr uint32_t B_BYTES = kUmmaTileRows * KP;
    constexpr uint32_t STAGE_BYTES = B_BYTES + kUmmaRowsMax;
    extern __shared__ __align__(1024) uint8_t usm[];
    const uint32_t base = (smem_u32(usm) + 1023u) & ~1023u;
    uint8_t* const base_ptr = usm + (base - smem_u32(usm));
    __shared__ __align__(8) uint64_t a_full, full_bar[S], empty_bar[S], tmem_full, tmem_empty;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q0 = (int64_t)blockIdx.x * 256;
    const int split = a.split0 + (int)blockIdx.y;
    const int64_t row0 = (int64_t)split * a.SL;
    const int64_t row1 = min(row0 + a.SL, a.ndb);
    const int64_t nrows = row1 > row0 ? row1 - row0 : 0;
    const int ntiles = (int)((nrows + kUmmaTileRows - 1) / kUmmaTileRows);
    const uint32_t rows_bytes = (uint32_t)kUmmaTileRows * a.Wr * 4;

    if (threadIdx.x == 0) {
        mbar_init(&a_full, 1);
        for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 8); }
        mbar_init(&tmem_full, 1);
        mbar_init(&tmem_empty, 8);
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
            mbar_arrive_expect_tx(&a_full, A_BYTES);
            tma_load_2d(base, &tmap_q8, smem_u32(&a_full), 0, (int)q0);
            tma_load_2d(base + 128 * KP, &tmap_q8, smem_u32(&a_full), 0, (int)q0 + 128);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % S;
                mbar_wait(&empty_bar[s], (uint32_t)(((t / S) & 1) ^ 1));
                mbar_arrive_expect_tx(&full_bar[s], B_BYTES + rows_bytes);
                const uint32_t sb = base + A_BYTES + s * STAGE_BYTES;
                const int r = (int)(row0 + (int64_t)t * kUmmaTileRows);
                tma_load_2d(sb, &tmap_db8, smem_u32(&full_bar[s]), 0, r);
                tma_load_2d(sb + B_BYTES, &tmap_rows, smem_u32(&full_bar[s]), 0, r);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(128, 128);
            mbar_wait(&a_full, 0);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % S;
                mbar_wait(&full_bar[s], (uint32_t)((t / S) & 1));
                mbar_wait(&tmem_empty, (uint32_t)((t & 1) ^ 1));  // the epilogue has read the previous tile's accumulators
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t bdesc = umma_desc_kmajor(base + A_BYTES + s * STAGE_BYTES, KP);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint64_t adesc = umma_desc_kmajor(base + h * 128 * KP, KP);
#pragma unroll
                    for (int k = 0; k < KP / 32; ++k)  // UMMA_K = 32 int8 = 32 bytes: start address advances by 2 (x16 B)
                        umma_i8(tmem_base + (uint32_t)(h * 128), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
                }
                umma_commit(&tmem_full);
            }
        }
    } else {
        // ---- epilogue: thread <-> query ----------------------------------------------------------------
        const int e = warp - 2;
        const int h = e >> 2;            // which 128-query half (A operand)
        const int quarter = warp & 3;    // TMEM lane quarter this warp may read
        const int64_t slot = q0 + h * 128 + quarter * 32 + lane;
        const bool valid = slot < a.nq;
        const int W = a.W, LW = a.LW, Wr = a.Wr;
        uint32_t qw[4] = {0, 0, 0, 0}, ql[4] = {0, 0, 0, 0};
        int thr_ip = 1 << 20;
        uint32_t pos = 0, nback = 0, start = 0, end = 0;  // front stack (d < T) grows up from start, back stack (d == T) down from end
        int Tq = -1;
        int64_t bin = -1;
        if (valid) {
            for (int w = 0; w < W && w < 4; ++w) qw[w] = a.q_rows[slot * Wr + w];
            for (int w = 0; w < LW && w < 4; ++w) ql[w] = a.q_rows[slot * Wr + W + w];
            const int T = a.thr[slot];
            Tq = T;
            if (T >= 0) thr_ip = a.b - 2 * T;
            bin = slot * a.P + split;
            start = (uint32_t)(bin * (int64_t)a.cap);
            end = start + a.cap;
            pos = start;
        }
        uint32_t* const lists = a.lists;
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % S;
            mbar_wait(&full_bar[s], (uint32_t)((t / S) & 1));  // packed rows of tile t have landed (observed by this thread)
            mbar_wait(&tmem_full, (uint32_t)(t & 1));           // MMAs of tile t are complete
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t* srows = reinterpret_cast<const uint32_t*>(base_ptr + A_BYTES + s * STAGE_BYTES + B_BYTES);
            const int tile_rows = (int)min((int64_t)kUmmaTileRows, nrows - (int64_t)t * kUmmaTileRows);
            uint32_t hits[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(h * 128 + g * 32);
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
                // bit (31 - j) of `miss` = sign(ip_j - thr) = 1 when row j is NOT a candidate
                uint32_t miss = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) miss = __funnelshift_l((uint32_t)((int)r[j] - thr_ip), miss, 1);
                const int vr = tile_rows - g * 32;  // rows of this 32-column group that exist in the database
                hits[g] = vr >= 32 ? ~miss : (vr <= 0 ? 0u : (~miss & (0xFFFFFFFFu << (32 - vr))));
            }
            // the accumulators are consumed: let the MMA warp start the next tile while the hits are written out
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty);
            // rare path: about R/Ndb of the pairs, ascending row order
            while (hits[0] | hits[1] | hits[2] | hits[3]) {
                const int g = hits[0] ? 0 : (hits[1] ? 1 : (hits[2] ? 2 : 3));
                const uint32_t hm = hits[0] ? hits[0] : (hits[1] ? hits[1] : (hits[2] ? hits[2] : hits[3]));
                const int j = __clz(hm);
                const uint32_t cleared = hm & ~(0x80000000u >> j);
                if (g == 0) hits[0] = cleared; else if (g == 1) hits[1] = cleared; else if (g == 2) hits[2] = cleared; else hits[3] = cleared;
                const int rl = g * 32 + j;
                const uint32_t* prow = srows + rl * Wr;
                int d = 0;
                uint32_t m = 0;
                if (Wr == 4) {  // one 16-byte load brings the code words and the label word(s): W = 2 (+ <= 2 label words) or W = 3 (+ 1)
                    const uint4 pr = *reinterpret_cast<const uint4*>(prow);
                    d = __popc(qw[0] ^ pr.x) + __popc(qw[1] ^ pr.y);
                    if (W == 3) { d += __popc(qw[2] ^ pr.z); m = ql[0] & pr.w; }
                    else { m = (ql[0] & pr.z) | (ql[1] & pr.w); }
                } else {
#pragma unroll
                    for (int w = 0; w < 4; ++w)
                        if (w < W) d += __popc(qw[w] ^ prow[w]);
#pragma unroll
                    for (int w = 0; w < 4; ++w)
                        if (w < LW) m |= ql[w] & prow[W + w];
                }
                const bool eq = d == Tq;
                if (pos + nback < end)
                    lists[eq ? end - 1u - nback : pos] = ((uint32_t)d * (1u << kIdxBits) + (uint32_t)(t * kUmmaTileRows + rl)) | (m ? 0x80000000u : 0u);
                if (eq) nback += 1; else pos += 1;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);  // shared-memory stage free for the producer
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

// ---- host side ----------------------------------------------------------------------------------------------
static int make_map_u8(CUtensorMap* map, const uint8_t* ptr, int64_t rows, int KP)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(HG_ECUDA, "select_umma: cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[2] = {(cuuint64_t)KP, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)KP};
    cuuint32_t box[2] = {(cuuint32_t)KP, (cuuint32_t)kUmmaTileRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    KP == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HG_ECUDA, "select_umma: cuTensorMapEncodeTiled(u8) failed (%d)", (int)r);
    return HG_OK;
}

static int make_map_rows(CUtensorMap* map, const uint32_t* ptr, int64_t rows, int Wr)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(HG_ECUDA, "select_umma: cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[2] = {(cuuint64_t)Wr, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)Wr * 4};
    cuuint32_t box[2] = {(cuuint32_t)Wr, (cuuint32_t)kUmmaTileRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HG_ECUDA, "select_umma: cuTensorMapEncodeTiled(rows) failed (%d)", (int)r);
    return HG_OK;
}

int umma_select_kp(int b, int Wr)
{
    // int8 row bytes; 0 = this shape stays on the POPC kernel (b <= 32: one POPC per pair is already cheap;
    // b > 128 or a packed-row stride that TMA cannot box: not implemented)
    if (b <= 32 || b > 128 || (Wr != 4 && Wr != 8)) return 0;
    return b <= 64 ? 64 : 128;
}

int umma_expand(const uint32_t* rows, int64_t n, int b, int Wr, int KP, uint8_t* out, cudaStream_t st)
{
    if (n <= 0) return HG_OK;
    const int sms = device_facts().sm_count > 0 ? device_facts().sm_count : 148;
    const int64_t total = n * (KP / 16);
    const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, 256), (int64_t)sms * 16));
    expand_codes_kernel<<<(unsigned)blocks, 256, 0, st>>>(rows, n, b, Wr, KP, out);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

template <int KP>
static int launch_umma(const CUtensorMap& tq, const CUtensorMap& tdb, const CUtensorMap& trows, const UmmaSelectArgs& a, cudaStream_t st)
{
    const size_t smem = (size_t)2 * 128 * KP + (size_t)UmmaCfg<KP>::S * (kUmmaTileRows * KP + kUmmaRowsMax) + 1024;
    static thread_local bool configured = false;
    if (!configured) {
        HG_CUDA_TRY(cudaFuncSetAttribute(select_umma_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((unsigned)ceil_div(a.nq, 256), (unsigned)a.n_splits);
    select_umma_kernel<KP><<<grid, kUmmaThreads, smem, st>>>(tq, tdb, trows, a);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

int umma_select_launch(const UmmaSelectArgs& a, cudaStream_t st)
{
    CUtensorMap tq, tdb, trows;
    int rc;
    if ((rc = make_map_u8(&tq, a.q8, a.nq, a.KP)) != HG_OK) return rc;
    if ((rc = make_map_u8(&tdb, a.db8, a.ndb, a.KP)) != HG_OK) return rc;
    if ((rc = make_map_rows(&trows, a.db_rows, a.ndb, a.Wr)) != HG_OK) return rc;
    return a.KP == 128 ? launch_umma<128>(tq, tdb, trows, a, st) : launch_umma<64>(tq, tdb, trows, a, st);
}

}  // namespace hg
