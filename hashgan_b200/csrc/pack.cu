// sign + bit-pack of features and 0/1 labels into PACKED ROWS:  [ W code words | LW label words | zero pad ]
// (row stride hg_row_words(b, L) uint32).  HBM-bound streaming kernels: coalesced loads, one warp ballot per
// 32 elements, direct word stores when a row of bits is word-aligned, else one atomicOr per word segment.
//
// New stage relative to the reference, which ranks raw tanh outputs (main.py:155-157 -> lib/metric.py:13);
// the label bits replace the per-query gather/compare of lib/metric.py:17-19.
#include "common.cuh"

#include <algorithm>

namespace hg {

template <typename T, bool LABEL>
__device__ __forceinline__ bool to_bit(T v, int& any_bad)
{
    if (LABEL) {
        any_bad |= (v != (T)0 && v != (T)1);
        return v == (T)1;
    }
    return v > (T)0;
}

// Aligned features: cols % 32 == 0, contiguous input rows (ld == cols) and a 16-byte aligned base, so 32 consecutive
// flat elements are exactly one output word.  Every lane loads 16 bytes (4 features) per request, four requests
// (2 KB per warp) in flight; the 8 lanes that share an output word OR their nibbles together with three shuffles.
__global__ void __launch_bounds__(256) pack_sign_aligned_kernel(const float* __restrict__ in, int64_t total, int words_per_row,
                                                                 uint32_t* __restrict__ out, int row_words)
{
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const float4* __restrict__ in4 = reinterpret_cast<const float4*>(in);
    const int64_t total4 = total >> 2;  // total % 32 == 0
    for (int64_t base = warp_id * (32 * U); base < total4; base += n_warps * (32 * U)) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = base + u * 32 + lane;
            v[u] = (e < total4) ? __ldg(in4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t nib = (v[u].x > 0.0f ? 1u : 0u) | (v[u].y > 0.0f ? 2u : 0u) | (v[u].z > 0.0f ? 4u : 0u) | (v[u].w > 0.0f ? 8u : 0u);
            uint32_t word = nib << (4 * (lane & 7));
            word |= __shfl_xor_sync(0xffffffffu, word, 1);
            word |= __shfl_xor_sync(0xffffffffu, word, 2);
            word |= __shfl_xor_sync(0xffffffffu, word, 4);
            const int64_t e = base + u * 32 + lane;  // float4 index; output word = e / 8
            if ((lane & 7) == 0 && e < total4) {
                const int64_t fw = e >> 3;
                const int64_t row = fw / words_per_row;
                out[row * row_words + (fw - row * words_per_row)] = word;
            }
        }
    }
}

// Labels, contiguous rows of L integers: one warp packs 32 rows at a time.  The 32 L values are read with coalesced
// loads (all in flight), each lane turns its values into bits of a shared-memory word, then every lane stores the
// label words of one row.
template <typename T>
__global__ void __launch_bounds__(256) pack_labels_kernel(const T* __restrict__ in, int64_t n, int L, uint32_t* __restrict__ out, int row_words,
                                                           int col0, int LW, int* __restrict__ bad)
{
    extern __shared__ uint32_t lab_sm[];  // [warps][32 * LW]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* sm = lab_sm + wib * 32 * LW;
    const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const uint32_t inv = 0xFFFFFFFFu / (uint32_t)L + 1u;  // e / L for e < 32 * L <= 32 * 1024 via mulhi
    int any_bad = 0;
    for (int64_t r0 = warp_id * 32; r0 < n; r0 += n_warps * 32) {
        const int rows = (int)min((int64_t)32, n - r0);
        const int count = rows * L;
        for (int i = lane; i < 32 * LW; i += 32) sm[i] = 0;
        __syncwarp();
        const T* src = in + r0 * L;
        for (int e0 = 0; e0 < count; e0 += 32 * 8) {
            T v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * 32 + lane;
                v[u] = e < count ? src[e] : (T)0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * 32 + lane;
                any_bad |= (v[u] != (T)0 && v[u] != (T)1);
                if (v[u] == (T)1) {
                    const int r = L == 1 ? e : (int)__umulhi((uint32_t)e, inv);
                    const int c = e - r * L;
                    atomicOr(&sm[r * LW + (c >> 5)], 1u << (c & 31));
                }
            }
        }
        __syncwarp();
        if (lane < rows)
            for (int w = 0; w < LW; ++w) out[(r0 + lane) * row_words + col0 + w] = sm[lane * LW + w];
        __syncwarp();
    }
    if (bad != nullptr) {
        any_bad = __any_sync(0xffffffffu, any_bad);
        if (any_bad && lane == 0) atomicOr(bad, 1);
    }
}

// General case (ragged rows or strided input).  One thread per element in flat order; a run of bits that lives
// in one output word is OR-ed in by its first lane.  The destination words must have been zeroed.
template <typename T, bool LABEL>
__global__ void __launch_bounds__(256) pack_bits_kernel(const T* __restrict__ in, int64_t n, int cols, int64_t ld,
                                                         uint32_t* __restrict__ out, int row_words, int col0, int* __restrict__ bad)
{
    const int lane = threadIdx.x & 31;
    const int64_t total = n * (int64_t)cols;
    const int64_t warp_stride = (int64_t)gridDim.x * (blockDim.x >> 5) * 32;
    int any_bad = 0;
    for (int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < total; base += warp_stride) {
        const int64_t e = base + lane;
        const bool in_range = e < total;
        int64_t row = 0;
        int col = 0;
        bool bit = false;
        if (in_range) {
            row = e / cols;
            col = (int)(e - row * cols);
            bit = to_bit<T, LABEL>(in[row * ld + col], any_bad);
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, bit);
        if (in_range && (lane == 0 || (col & 31) == 0)) {
            int len = 32 - (col & 31);
            if (cols - col < len) len = cols - col;
            if (32 - lane < len) len = 32 - lane;
            const uint32_t mask = (len >= 32) ? 0xffffffffu : ((1u << len) - 1u);
            const uint32_t seg = (ballot >> lane) & mask;
            if (seg) atomicOr(&out[row * row_words + col0 + (col >> 5)], seg << (col & 31));
        }
    }
    if (LABEL && bad != nullptr) {
        any_bad = __any_sync(0xffffffffu, any_bad);
        if (any_bad && lane == 0) atomicOr(bad, 1);
    }
}

// ---- pack fused with its exchange: the all-gather of packed rows as peer-memory stores -------------------------------
// Multi-GPU evaluation row-shards the database, packs locally and all-gathers the packed rows (sharding.py).
// Here the pack kernel IS the all-gather: every packed row is stored straight into the database buffer of every rank
// (NVLink / NVSwitch peer pointers of a symmetric allocation; dst[r] already points at this rank's first row inside rank
// r's buffer), so no separate collective runs -- one signal-pad barrier afterwards publishes the rows.
// One warp packs 32 rows: code words by 16-byte loads + shuffles, label words through shared memory, then whole rows
// (pads zeroed) are written with coalesced stores, destination by destination.
constexpr int kMaxPeers = 16;
struct PeerDst { uint32_t* p[kMaxPeers]; };

template <typename T>
__global__ void __launch_bounds__(256) pack_rows_push_kernel(const float* __restrict__ feat, const T* __restrict__ lab, int64_t n, int b, int L,
                                                             int W, int LW, int Wr, const __grid_constant__ PeerDst dst, int n_dst,
                                                             int* __restrict__ bad)
{
    extern __shared__ uint32_t push_sm[];  // [warps][32 * Wr]: the 32 packed rows of the warp
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* rows = push_sm + wib * 32 * Wr;
    const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int cw = b >> 5;              // code words that carry bits (b % 32 == 0)
    const int f4_per_row = b >> 2;      // float4 per feature row
    const uint32_t inv = 0xFFFFFFFFu / (uint32_t)L + 1u;
    int any_bad = 0;
    for (int64_t r0 = warp_id * 32; r0 < n; r0 += n_warps * 32) {
        const int nrows = (int)min((int64_t)32, n - r0);
        for (int i = lane; i < 32 * Wr; i += 32) rows[i] = 0;
        __syncwarp();
        // code words: the warp's rows are one contiguous block of nrows * b floats
        const float4* src4 = reinterpret_cast<const float4*>(feat + r0 * b);
        const int total4 = nrows * f4_per_row;
        for (int e0 = 0; e0 < total4; e0 += 32 * 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * 32 + lane;
                v[u] = e < total4 ? __ldg(src4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t nib = (v[u].x > 0.0f ? 1u : 0u) | (v[u].y > 0.0f ? 2u : 0u) | (v[u].z > 0.0f ? 4u : 0u) | (v[u].w > 0.0f ? 8u : 0u);
                uint32_t word = nib << (4 * (lane & 7));
                word |= __shfl_xor_sync(0xffffffffu, word, 1);
                word |= __shfl_xor_sync(0xffffffffu, word, 2);
                word |= __shfl_xor_sync(0xffffffffu, word, 4);
                const int e = e0 + u * 32 + lane;
                if ((lane & 7) == 0 && e < total4) {
                    const int fw = e >> 3;  // flat code word of the block
                    const int row = fw / cw;
                    rows[row * Wr + (fw - row * cw)] = word;
                }
            }
        }
        // label words
        if (lab != nullptr) {
            const T* lsrc = lab + r0 * L;
            const int count = nrows * L;
            for (int e0 = 0; e0 < count; e0 += 32 * 8) {
                T v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = e0 + u * 32 + lane;
                    v[u] = e < count ? lsrc[e] : (T)0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = e0 + u * 32 + lane;
                    any_bad |= (v[u] != (T)0 && v[u] != (T)1);
                    if (v[u] == (T)1) {
                        const int r = L == 1 ? e : (int)__umulhi((uint32_t)e, inv);
                        const int c = e - r * L;
                        atomicOr(&rows[r * Wr + W + (c >> 5)], 1u << (c & 31));
                    }
                }
            }
        }
        __syncwarp();
        // push: nrows * Wr contiguous words per destination
        const int words = nrows * Wr;
        for (int d = 0; d < n_dst; ++d) {
            uint32_t* out = dst.p[d] + r0 * Wr;
            if ((Wr & 3) == 0) {
                for (int i = lane * 4; i < words; i += 128) *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(rows + i);
            } else {
                for (int i = lane; i < words; i += 32) out[i] = rows[i];
            }
        }
        __syncwarp();
    }
    if (bad != nullptr) {
        any_bad = __any_sync(0xffffffffu, any_bad);
        if (any_bad && lane == 0) atomicOr(bad, 1);
    }
}

template <typename T>
static int launch_push(const float* feat, const void* lab, int64_t n, int b, int L, int W, int LW, int Wr, const PeerDst& dst, int n_dst, int* bad,
                       cudaStream_t st)
{
    const size_t smem = (size_t)8 * 32 * Wr * sizeof(uint32_t);
    const int sms = device_facts().sm_count > 0 ? device_facts().sm_count : 148;
    const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>(ceil_div(ceil_div(n, 32), 8), (int64_t)sms * 8));
    pack_rows_push_kernel<T><<<(unsigned)blocks, 256, smem, st>>>(feat, (const T*)lab, n, b, L, W, LW, Wr, dst, n_dst, bad);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

static int64_t grid_for(int64_t warps_needed, int mult)
{
    const int sms = device_facts().sm_count > 0 ? device_facts().sm_count : 148;
    const int64_t max_blocks = (int64_t)sms * 8 * mult;  // 8 resident CTAs of 256 threads per SM
    int64_t blocks = ceil_div(warps_needed, 8);
    return blocks > max_blocks ? max_blocks : (blocks < 1 ? 1 : blocks);
}

template <typename T>
static int launch_labels(const void* lab, int64_t n, int L, uint32_t* rows, int row_words, int col0, int* bad, cudaStream_t st)
{
    const int LW = (L + 31) / 32;
    const size_t smem = (size_t)8 * 32 * LW * sizeof(uint32_t);
    pack_labels_kernel<T><<<(unsigned)grid_for(ceil_div(n, 32), 1), 256, smem, st>>>((const T*)lab, n, L, rows, row_words, col0, LW, bad);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

}  // namespace hg

extern "C" int hg_code_words(int b)
{
    if (b <= 0 || b > HG_MAX_BITS) return 0;
    const int W = (b + 31) / 32;
    return W <= 4 ? W : 8;
}

extern "C" int hg_label_words(int L)
{
    if (L <= 0 || L > HG_MAX_LABELS) return 0;
    return (L + 31) / 32;
}

extern "C" int hg_row_words(int b, int L)
{
    const int W = hg_code_words(b), LW = hg_label_words(L);
    if (W == 0 || LW == 0) return 0;
    const int align = (W == 2) ? 2 : ((W == 4 || W == 8) ? 4 : 1);  // vector loads of the code words
    int Wr = ((W + LW + align - 1) / align) * align;
    // the tensor-core select stages packed rows by TMA: 16- or 32-byte rows.  Hash lengths of 33..128 bits with more than
    // 32 labels (e.g. 64-bit codes on NUS-WIDE's 81 labels) get 8-word rows instead of 5..7.
    if (W >= 2 && W <= 4 && LW <= 4 && W + LW > 4) Wr = 8;
    return Wr;
}

extern "C" int hg_pack_rows(const float* d_feat, int64_t ld, const void* d_lab, int lab_elem_bytes, int64_t n, int b, int L,
                            uint32_t* d_rows, int* d_bad, void* stream)
{
    const int W = hg_code_words(b), LW = hg_label_words(L), Wr = hg_row_words(b, L);
    if (W == 0) return hg::fail(HG_EINVAL, "hg_pack_rows: unsupported hash length b=%d (1..%d)", b, HG_MAX_BITS);
    if (LW == 0) return hg::fail(HG_EINVAL, "hg_pack_rows: unsupported label width L=%d (1..%d)", L, HG_MAX_LABELS);
    if (n < 0 || ld < b) return hg::fail(HG_EINVAL, "hg_pack_rows: bad n=%lld / ld=%lld", (long long)n, (long long)ld);
    if (n == 0) return HG_OK;
    if (!d_feat || !d_rows) return hg::fail(HG_EINVAL, "hg_pack_rows: NULL pointer");
    if (d_lab && lab_elem_bytes != 8 && lab_elem_bytes != 4 && lab_elem_bytes != 1)
        return hg::fail(HG_EINVAL, "hg_pack_rows: lab_elem_bytes must be 8, 4 or 1 (got %d)", lab_elem_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    HG_CUDA_TRY(cudaMemsetAsync(d_rows, 0, sizeof(uint32_t) * (size_t)n * Wr, st));
    const int64_t total = n * (int64_t)b;
    if ((b & 31) == 0 && ld == b && (reinterpret_cast<uintptr_t>(d_feat) & 15) == 0) {
        hg::pack_sign_aligned_kernel<<<(unsigned)hg::grid_for(hg::ceil_div(total, 128), 1), 256, 0, st>>>(d_feat, total, b / 32, d_rows, Wr);
    } else {
        hg::pack_bits_kernel<float, false><<<(unsigned)hg::grid_for(hg::ceil_div(total, 32), 4), 256, 0, st>>>(d_feat, n, b, ld, d_rows, Wr, 0, nullptr);
    }
    hg::count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    if (!d_lab) return HG_OK;
    switch (lab_elem_bytes) {
        case 8: return hg::launch_labels<long long>(d_lab, n, L, d_rows, Wr, W, d_bad, st);
        case 4: return hg::launch_labels<int>(d_lab, n, L, d_rows, Wr, W, d_bad, st);
        default: return hg::launch_labels<signed char>(d_lab, n, L, d_rows, Wr, W, d_bad, st);
    }
}

extern "C" int hg_pack_rows_push(const float* d_feat, int64_t ld, const void* d_lab, int lab_elem_bytes, int64_t n, int b, int L,
                                 uint32_t* const* h_dst_rows, int n_dst, int* d_bad, void* stream)
{
    const int W = hg_code_words(b), LW = hg_label_words(L), Wr = hg_row_words(b, L);
    if (W == 0 || LW == 0) return hg::fail(HG_EINVAL, "hg_pack_rows_push: unsupported b=%d / L=%d", b, L);
    if (n < 0 || !h_dst_rows || n_dst < 1 || n_dst > hg::kMaxPeers) return hg::fail(HG_EINVAL, "hg_pack_rows_push: 1..%d destinations", hg::kMaxPeers);
    if ((b & 31) != 0 || ld != b || (reinterpret_cast<uintptr_t>(d_feat) & 15) != 0)
        return hg::fail(HG_ERANGE, "hg_pack_rows_push: needs b %% 32 == 0 and contiguous 16-byte aligned feature rows (use hg_pack_rows + all-gather)");
    if (d_lab && lab_elem_bytes != 8 && lab_elem_bytes != 4 && lab_elem_bytes != 1)
        return hg::fail(HG_EINVAL, "hg_pack_rows_push: lab_elem_bytes must be 8, 4 or 1 (got %d)", lab_elem_bytes);
    if (n == 0) return HG_OK;
    hg::PeerDst dst{};
    for (int d = 0; d < n_dst; ++d) {
        if (!h_dst_rows[d] || (reinterpret_cast<uintptr_t>(h_dst_rows[d]) & ((Wr & 3) == 0 ? 15 : 3)) != 0)
            return hg::fail(HG_EINVAL, "hg_pack_rows_push: destination %d is NULL or misaligned", d);
        dst.p[d] = h_dst_rows[d];
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (d_lab ? lab_elem_bytes : 8) {
        case 8: return hg::launch_push<long long>(d_feat, d_lab, n, b, L, W, LW, Wr, dst, n_dst, d_bad, st);
        case 4: return hg::launch_push<int>(d_feat, d_lab, n, b, L, W, LW, Wr, dst, n_dst, d_bad, st);
        default: return hg::launch_push<signed char>(d_feat, d_lab, n, b, L, W, LW, Wr, dst, n_dst, d_bad, st);
    }
}
