// Shared helpers for libhashgan_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>

#include "../../include/hashgan_b200.h"

namespace hg {

// ---- error plumbing --------------------------------------------------------------------------
char* last_error_buf();
inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}
#define HG_CUDA_TRY(expr)                                                                            \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return hg::fail(HG_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// every kernel launch of the library is counted (bench.py reports it as gpu_launches)
void count_launch(int n = 1);

// optional per-phase CUDA-event timing of hg_hamming_map (flag HG_FLAG_TIMING)
enum Phase { kPhaseSample = 0, kPhaseThreshold, kPhaseExpand, kPhaseSelect, kPhaseAp, kPhaseExact, kNumPhases };
struct PhaseTimer {
    cudaEvent_t ev[kNumPhases + 1] = {};
    bool created = false, armed = false;
    int ensure();
    void mark(int i, cudaStream_t st) { if (armed) cudaEventRecord(ev[i], st); }
};
PhaseTimer& phase_timer();

// chunked host pipeline (api.cu <-> rank.cu): database rows arrive chunk by chunk (whole splits per chunk)
struct MapChunks {
    static constexpr int kMax = 64;
    int K = 0;
    int64_t row_lo[kMax], row_hi[kMax];
    // threshold sample drawn over the WHOLE database (the caller copies and packs these rows before chunk 0): n_seg
    // segments of seg_rows rows, segment i = database rows [i * seg_stride, i * seg_stride + seg_rows); 0 = no sample
    // (the plan ranks the whole database for its estimate, or the exact path was forced)
    int64_t sample_n_seg = 0, sample_seg_stride = 0;
    int sample_seg_rows = 0;
    const uint32_t* sample_packed = nullptr;  // [sample_n_seg * sample_seg_rows, Wr] packed rows, valid on the stream before prepare(0)
};
typedef int (*PrepareRowsFn)(void* user, int chunk, int64_t row_lo, int64_t row_hi, cudaStream_t st);
int plan_chunks(int64_t nq, int64_t ndb, int b, int L, int64_t R, int k_req, MapChunks* out, size_t* ws_bytes);
int hamming_map_chunked(const uint32_t* q_rows, int64_t nq, const uint32_t* db_rows, int64_t ndb, int b, int L, int64_t R, unsigned flags,
                        double* d_ap, void* ws, size_t ws_bytes, cudaStream_t st, const MapChunks* chunks, PrepareRowsFn prepare, void* user);

// tensor-core select (select_umma.cu), driven from rank.cu
struct UmmaSelectArgs {
    const uint32_t* q_rows;
    const uint32_t* db_rows;
    int64_t nq, ndb;
    int b, W, LW, Wr, KP;
    const int* thr;
    int P, split0, n_splits;
    int64_t SL;
    uint32_t* lists;
    uint32_t cap;
    uint32_t* bin_cnt;   // candidates closer than the threshold distance, stored from the start of the bin
    uint32_t* bin_cnt0;  // candidates at the threshold distance, stored from the end of the bin downwards
    const uint8_t* q8;   // [nq, KP] int8 codes
    const uint8_t* db8;  // [round_up(ndb, 32), KP], rows of every 32-row group permuted (select_umma.cu)
    const uint8_t* qx;   // [round_up(nq, 256), 32] threshold columns of the A operand
    const uint8_t* bx;   // [128, 32] threshold columns of the B operand (constant)
    int queued;          // 1 = select_q_kernel where it applies (parked mask words), 0 = select_umma_kernel (hits walked tile by tile)
    int drain_lanes;     // queued mode: lanes that must have parked work before a waiting warp spends a hit step
    uint32_t prmt_sel;   // PRMT selector of the epilogue's sign gather
    int rows_paired;     // set by the launcher: 2-word packed rows are staged as 16-byte row pairs
};
int umma_select_kp(int b, int Wr);  // int8 row bytes (64 / 128), 0 = shape not supported by the tensor-core path
int umma_expand_q(const uint32_t* rows, int64_t n, int b, int Wr, int KP, uint8_t* out, cudaStream_t st);
// database rows [lo, hi) of db_rows (whole 32-row groups, or up to the end of the database) -> db8
int umma_expand_db(const uint32_t* db_rows, int64_t lo, int64_t hi, int64_t ndb, int b, int Wr, int KP, uint8_t* db8, cudaStream_t st);
int umma_thr_columns(const int* thr, int64_t nq, int b, uint8_t* qx, uint8_t* bx, cudaStream_t st);
int umma_select_launch(const UmmaSelectArgs& a, cudaStream_t st);

struct DeviceFacts {
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t l2_bytes = 0;
    bool ok = false;
};
const DeviceFacts& device_facts();

// ---- entry layout of a candidate ---------------------------------------------------------------
// [0,21) row inside the db split, [21,31) Hamming distance, bit 31 relevance (set by the AP kernel).
constexpr int kIdxBits = 21;
constexpr uint32_t kIdxMask = (1u << kIdxBits) - 1;
constexpr uint32_t kDistMask = 0x3FFu;
constexpr int64_t kMaxSplitRows = int64_t(1) << kIdxBits;

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---- PTX wrappers: mbarrier + 1-D bulk TMA (cp.async.bulk) -------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "HG_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra HG_DONE;\n"
        "bra HG_WAIT;\n"
        "HG_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (TMA, SASS UBLKCP); bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif

}  // namespace hg
