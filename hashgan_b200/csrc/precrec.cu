// precision@R / recall@R next to mAP@R.  The reference reports mAP only (lib/metric.py:12-24); the two
// companions are defined on the same ranking and the same relevance test (lib/metric.py:17-19: a database row is relevant
// when it shares at least one positive label with the query):
//     precision@R = rel / R                      rel = relevant rows inside the top-R (lib/metric.py:20, hg_hamming_map's d_rel)
//     recall@R    = rel / total                  total = relevant rows in the WHOLE database
// `total` is the only quantity the candidate walk cannot deliver, so this file computes it: one pass over the label words
// of the packed rows (all pairs, but 1 AND + 1 compare + 1 add per pair per label word and no code words).
#include "common.cuh"

#include <algorithm>

namespace hg {

constexpr int kRelThreads = 128;
constexpr int kRelTile = 2048;  // database rows per shared-memory tile (label words only)

// thread <-> query (label words in registers); the label words of a tile of database rows are staged in shared memory and
// read as warp broadcasts, four rows per 16-byte load when the label fits one word
template <int LW>
__global__ void __launch_bounds__(kRelThreads) relevant_totals_kernel(const uint32_t* __restrict__ q_rows, int64_t nq, const uint32_t* __restrict__ db_rows,
                                                                      int64_t ndb, int Wr, int W, int64_t rows_per_cta, uint32_t* __restrict__ total)
{
    __shared__ __align__(16) uint32_t lab[kRelTile * LW];
    const int64_t q = (int64_t)blockIdx.x * kRelThreads + threadIdx.x;
    const bool valid = q < nq;
    uint32_t ql[LW];
#pragma unroll
    for (int w = 0; w < LW; ++w) ql[w] = valid ? q_rows[q * Wr + W + w] : 0u;
    const int64_t lo = (int64_t)blockIdx.y * rows_per_cta, hi = min(ndb, lo + rows_per_cta);
    uint32_t cnt = 0;
    for (int64_t r0 = lo; r0 < hi; r0 += kRelTile) {
        const int rows = (int)min((int64_t)kRelTile, hi - r0);
        __syncthreads();
        for (int i = threadIdx.x; i < rows * LW; i += kRelThreads) lab[i] = __ldg(db_rows + (r0 + i / LW) * Wr + W + (i % LW));
        for (int i = rows * LW + threadIdx.x; i < ((rows * LW + 3) & ~3); i += kRelThreads) lab[i] = 0u;  // pad the last 16-byte group
        __syncthreads();
        if (LW == 1) {
            const uint4* l4 = reinterpret_cast<const uint4*>(lab);
#pragma unroll 4
            for (int j = 0; j < (rows + 3) / 4; ++j) {
                const uint4 v = l4[j];
                cnt += ((v.x & ql[0]) != 0u) + ((v.y & ql[0]) != 0u) + ((v.z & ql[0]) != 0u) + ((v.w & ql[0]) != 0u);
            }
        } else {
#pragma unroll 2
            for (int j = 0; j < rows; ++j) {
                uint32_t m = 0;
#pragma unroll
                for (int w = 0; w < LW; ++w) m |= lab[j * LW + w] & ql[w];
                cnt += (m != 0u);
            }
        }
    }
    if (valid && cnt) atomicAdd(&total[q], cnt);
}

}  // namespace hg

extern "C" int hg_relevant_totals(const uint32_t* d_q_rows, int64_t nq, const uint32_t* d_db_rows, int64_t ndb, int b, int L, uint32_t* d_total,
                                  void* stream)
{
    const int W = hg_code_words(b), LW = hg_label_words(L), Wr = hg_row_words(b, L);
    if (W == 0 || LW == 0) return hg::fail(HG_EINVAL, "hg_relevant_totals: unsupported b=%d / L=%d", b, L);
    if (nq < 0 || ndb < 0) return hg::fail(HG_EINVAL, "hg_relevant_totals: negative size");
    if (nq == 0) return HG_OK;
    if (!d_q_rows || !d_total || (ndb > 0 && !d_db_rows)) return hg::fail(HG_EINVAL, "hg_relevant_totals: NULL pointer");
    if (!hg::device_facts().ok) return hg::fail(HG_ECUDA, "hg_relevant_totals: no CUDA device");
    cudaStream_t st = (cudaStream_t)stream;
    HG_CUDA_TRY(cudaMemsetAsync(d_total, 0, sizeof(uint32_t) * (size_t)nq, st));
    if (ndb == 0) return HG_OK;
    const int sms = hg::device_facts().sm_count > 0 ? hg::device_facts().sm_count : 148;
    const int64_t qblocks = hg::ceil_div(nq, hg::kRelThreads);
    // enough CTAs for ~8 per SM; every CTA walks whole tiles
    int64_t chunks = std::max<int64_t>(1, std::min<int64_t>(hg::ceil_div(ndb, hg::kRelTile), hg::ceil_div((int64_t)sms * 8, qblocks)));
    const int64_t rows_per_cta = hg::round_up(hg::ceil_div(ndb, chunks), hg::kRelTile);
    chunks = hg::ceil_div(ndb, rows_per_cta);
    dim3 grid((unsigned)qblocks, (unsigned)chunks);
    switch (LW) {
        case 1: hg::relevant_totals_kernel<1><<<grid, hg::kRelThreads, 0, st>>>(d_q_rows, nq, d_db_rows, ndb, Wr, W, rows_per_cta, d_total); break;
        case 2: hg::relevant_totals_kernel<2><<<grid, hg::kRelThreads, 0, st>>>(d_q_rows, nq, d_db_rows, ndb, Wr, W, rows_per_cta, d_total); break;
        case 3: hg::relevant_totals_kernel<3><<<grid, hg::kRelThreads, 0, st>>>(d_q_rows, nq, d_db_rows, ndb, Wr, W, rows_per_cta, d_total); break;
        case 4: hg::relevant_totals_kernel<4><<<grid, hg::kRelThreads, 0, st>>>(d_q_rows, nq, d_db_rows, ndb, Wr, W, rows_per_cta, d_total); break;
        default: return hg::fail(HG_EINVAL, "hg_relevant_totals: unsupported label word count %d", LW);
    }
    hg::count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}
