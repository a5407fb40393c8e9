// sign + bit-pack of features, and 0/1 label matrices -> bit rows.  HBM-bound streaming kernels:
// coalesced loads, one warp ballot per 32 elements, direct word stores (row length % 32 == 0) or
// one atomicOr per row/word segment (ragged rows such as b = 48 or L = 10).
//
// Replaces (new stage, the reference ranks raw tanh outputs): main.py:155-157 -> lib/metric.py:13,
// and the label gather/compare of lib/metric.py:17-19.
#include "common.cuh"

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

// Dense case: rows are contiguous (ld == cols) and cols % 32 == 0 and the output row is exactly cols/32
// words, so flat element e lands in flat word e/32.  Four independent 128-byte warp loads in flight.
template <typename T, bool LABEL>
__global__ void __launch_bounds__(256) pack_bits_dense_kernel(const T* __restrict__ in, int64_t total, uint32_t* __restrict__ out,
                                                               int* __restrict__ bad)
{
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    int any_bad = 0;
    for (int64_t base = warp_id * (32 * U); base < total; base += n_warps * (32 * U)) {
        T v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = base + u * 32 + lane;
            v[u] = (e < total) ? in[e] : (T)0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t ballot = __ballot_sync(0xffffffffu, to_bit<T, LABEL>(v[u], any_bad));
            const int64_t e0 = base + u * 32;
            if (lane == 0 && e0 < total) out[e0 >> 5] = ballot;
        }
    }
    if (LABEL && bad != nullptr) {
        any_bad = __any_sync(0xffffffffu, any_bad);
        if (any_bad && lane == 0) atomicOr(bad, 1);
    }
}

// General case (ragged rows, padded output rows or strided input).  One thread per element in flat
// order; a run of bits that lives in one output word is written with one atomicOr by its first lane.
template <typename T, bool LABEL>
__global__ void __launch_bounds__(256) pack_bits_kernel(const T* __restrict__ in, int64_t n, int cols, int64_t ld, int wpr,
                                                         uint32_t* __restrict__ out, int* __restrict__ bad)
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
            if (seg) atomicOr(&out[row * wpr + (col >> 5)], seg << (col & 31));
        }
    }
    if (LABEL && bad != nullptr) {
        any_bad = __any_sync(0xffffffffu, any_bad);
        if (any_bad && lane == 0) atomicOr(bad, 1);
    }
}

template <typename T, bool LABEL>
static int launch_pack(const void* in, int64_t n, int cols, int64_t ld, int wpr, uint32_t* out, int* bad, cudaStream_t st)
{
    if (n == 0) return HG_OK;
    const int64_t total = n * (int64_t)cols;
    const int threads = 256;
    const int sms = device_facts().sm_count > 0 ? device_facts().sm_count : 148;
    const int64_t max_blocks = (int64_t)sms * 8;  // 8 resident CTAs of 256 threads per SM
    const bool dense = ((cols & 31) == 0) && (wpr * 32 == cols) && (ld == cols);
    if (dense) {
        int64_t blocks = ceil_div(ceil_div(total, 32 * 4), threads / 32);
        if (blocks > max_blocks) blocks = max_blocks;
        pack_bits_dense_kernel<T, LABEL><<<(unsigned)blocks, threads, 0, st>>>((const T*)in, total, out, bad);
    } else {
        HG_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(uint32_t) * (size_t)n * wpr, st));
        int64_t blocks = ceil_div(ceil_div(total, 32), threads / 32);
        if (blocks > max_blocks * 4) blocks = max_blocks * 4;
        pack_bits_kernel<T, LABEL><<<(unsigned)blocks, threads, 0, st>>>((const T*)in, n, cols, ld, wpr, out, bad);
    }
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

extern "C" int hg_pack_sign_f32(const float* d_feat, int64_t n, int b, int64_t ld, uint32_t* d_codes, void* stream)
{
    const int wpr = hg_code_words(b);
    if (wpr == 0) return hg::fail(HG_EINVAL, "hg_pack_sign_f32: unsupported hash length b=%d (1..%d)", b, HG_MAX_BITS);
    if (n < 0 || ld < b) return hg::fail(HG_EINVAL, "hg_pack_sign_f32: bad n=%lld / ld=%lld", (long long)n, (long long)ld);
    if (n > 0 && (!d_feat || !d_codes)) return hg::fail(HG_EINVAL, "hg_pack_sign_f32: NULL pointer");
    return hg::launch_pack<float, false>(d_feat, n, b, ld, wpr, d_codes, nullptr, (cudaStream_t)stream);
}

extern "C" int hg_pack_labels(const void* d_lab, int elem_bytes, int64_t n, int L, uint32_t* d_packed, int* d_bad, void* stream)
{
    const int wpr = hg_label_words(L);
    if (wpr == 0) return hg::fail(HG_EINVAL, "hg_pack_labels: unsupported label width L=%d", L);
    if (n < 0) return hg::fail(HG_EINVAL, "hg_pack_labels: negative n");
    if (n > 0 && (!d_lab || !d_packed)) return hg::fail(HG_EINVAL, "hg_pack_labels: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    switch (elem_bytes) {
        case 8: return hg::launch_pack<long long, true>(d_lab, n, L, L, wpr, d_packed, d_bad, st);
        case 4: return hg::launch_pack<int, true>(d_lab, n, L, L, wpr, d_packed, d_bad, st);
        case 1: return hg::launch_pack<signed char, true>(d_lab, n, L, L, wpr, d_packed, d_bad, st);
        default: return hg::fail(HG_EINVAL, "hg_pack_labels: elem_bytes must be 8, 4 or 1 (got %d)", elem_bytes);
    }
}
