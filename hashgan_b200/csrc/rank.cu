// Hamming ranking + mAP@R on bit-packed codes: the metric hot path (lib/metric.py:12-23).
//
//   lib/metric.py:13  ips = np.dot(query.output, database.output.T)   -> d_H = popc(q ^ db)   (ip = b - 2 d_H on +-1 codes)
//   lib/metric.py:14  ids = np.argsort(-ips, 1)                       -> exact (d_H asc, db row asc) ranking of the top R
//   lib/metric.py:16-23 relevance / cumsum / AP                       -> integer prefix counts, fp64 divides
//
// Design: keys take only b+1 values, so ranking is a counting sort, never a comparison
// sort, and the [Nq, Ndb] matrix is never materialised.
//   1. hist_kernel on a strided SAMPLE of the database -> per-query distance histogram.
//   2. thr_kernel: per-query threshold T_q such that count(d <= T_q) >= R with high probability.
//   3. select_kernel (the hot kernel): ONE pass over all pairs.  Thread <-> query (QT queries in
//      registers), database tiles (packed rows: code words + label words) staged into shared memory by 1-D
//      bulk TMA and read as warp broadcasts; per pair: W x (XOR, POPC), adds, one compare; pairs with
//      d <= T_q (about R/Ndb of them) are appended, with their relevance bit, to the thread's PRIVATE bin
//      (query, db split) -> entries are in database-row order with no atomics.
//   4. ap_kernel: G threads per query walk its bins sequentially in row order; private per-distance running
//      counters give every candidate its exact rank and relevant-prefix count -> AP in fp64; optional ids/dist.
//   5. Exactness guard: a query whose candidate count fell short of R, or whose bin overflowed, is put on
//      a fail list and redone by the two-pass exact path (full per-split histograms -> exact d*, exact bin
//      offsets/quotas -> select_kernel<EXACT> -> ap_kernel).  All launches are unconditional and sized for
//      the worst case; CTAs beyond the fail count exit at once, so the call stays asynchronous.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

namespace hg {

// ================================================================================================
// Plan: all sizes derived from (nq, ndb, b, L, R) and the SM count, shared by the workspace query and
// the launcher.
// ================================================================================================
struct Plan {
    int b = 0, L = 0, W = 0, LW = 0, Wr = 0, QT = 0, TQ = 0, TILE = 0, P = 0, nqt = 0;
    int umma_kp = 0;  // > 0: the fast-path select runs on the tensor cores (select_umma.cu), int8 rows of this many bytes
    bool dense = false;  // R is a large part of the database: no selection at all, dense_ap_kernel walks the packed rows
    bool queued = false; // select_q_kernel (mask words parked, hits consumed asynchronously) instead of select_umma_kernel
    int64_t nq = 0, ndb = 0, R = 0, SL = 0;
    uint32_t cap = 0;
    // sample pass
    int64_t n_seg = 0, seg_stride = 0, sample_rows = 0;
    int seg_per_chunk = 0, n_chunks = 0;
    // workspace byte offsets
    size_t off_ctrl = 0, off_thr = 0, off_thr2 = 0, off_fail = 0, off_wide = 0, off_hist_s = 0, off_bin_cnt = 0, off_bin_cnt0 = 0, off_bin_cnt2 = 0,
           off_bin_off2 = 0, off_bin_cap2 = 0, off_quota2 = 0, off_hist2 = 0, off_lists = 0, off_q8 = 0, off_db8 = 0, off_qx = 0, off_bx = 0, total = 0;
    bool ok = false;
};

constexpr int kSelectThreads = 128;
constexpr int kSelectCtasPerSm = 16;
constexpr int kUmmaWaves = 8;
constexpr int kHistThreads = 128;
constexpr int kApWarps = 8;  // exact_plan_kernel: warps per CTA
constexpr int64_t kSampleTarget = 16384;
constexpr float kSampleZ = 4.0f;

// rows per shared-memory tile: about 8 KB per pipeline stage
static int tile_rows_for(int Wr) { return Wr <= 2 ? 1024 : (Wr <= 4 ? 512 : (Wr <= 8 ? 256 : 128)); }

static int env_int(const char* name, int fallback)
{
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : fallback;
}

static Plan make_plan(int64_t nq, int64_t ndb, int b, int L, int64_t R, int ctas_mult = 1)
{
    Plan p;
    p.W = hg_code_words(b);
    p.LW = hg_label_words(L);
    p.Wr = hg_row_words(b, L);
    if (p.W == 0 || p.LW == 0 || nq <= 0 || ndb <= 0 || R <= 0 || R > ndb || ndb >= (int64_t(1) << 31) || nq >= (int64_t(1) << 31))
        return p;
    p.b = b; p.L = L; p.nq = nq; p.ndb = ndb; p.R = R;
    {
        // every second row or more is in the top-R: selecting costs more than it saves (HG_DENSE=0 / 1 forces the choice)
        const int dv = env_int("HG_DENSE", -1);
        p.dense = dv < 0 ? (R * 2 >= ndb && ndb / 32 <= kMaxSplitRows) : (dv != 0 && ndb / 32 <= kMaxSplitRows);
    }
    const int sms = device_facts().sm_count > 0 ? device_facts().sm_count : 148;
    p.QT = nq >= 1024 ? 2 : 1;  // measured on B200 (C4): QT=2 with ~16 CTAs/SM beats QT=4 (scripts/tune_select.py)
    {
        const int qt = env_int("HG_SELECT_QT", 0);  // tuning override
        if (qt == 1 || qt == 2 || qt == 4) p.QT = qt;
    }
    p.TQ = kSelectThreads * p.QT;
    p.nqt = (int)ceil_div(nq, p.TQ);
    p.TILE = tile_rows_for(p.Wr);
    // db splits: one wave of CTAs that are all resident (grid ~ SMs x CTAs/SM), split length a multiple of the
    // tile, at most 2^21 rows
    {
        const char* be = getenv("HG_SELECT_BACKEND");
        if (!(be && be[0] == 'p')) p.umma_kp = umma_select_kp(b, p.Wr);  // "popc" forces the POPC kernel
        // b <= 32 costs the POPC kernel ONE word-op per pair, and a top-R that is a large part of the database (C1: R = Ndb)
        // makes every pair a candidate -- the tensor-core kernel's advantage is the cheap rejection of non-candidates, so it
        // only takes short codes when the top-R is sparse
        if (p.umma_kp == 32 && R * 8 > ndb && !(be && be[0] == 'u')) p.umma_kp = 0;
    }
    int64_t target_ctas = (int64_t)sms * env_int("HG_SELECT_CTAS_PER_SM", kSelectCtasPerSm) * ctas_mult;
    int64_t units = p.nqt;
    if (p.umma_kp) {  // one 320-thread CTA per SM holds all 512 TMEM columns: size the grid in waves of 148
        target_ctas = (int64_t)sms * env_int("HG_UMMA_WAVES", kUmmaWaves) * ctas_mult;
        units = ceil_div(nq, 256);
    }
    int64_t P0 = std::max<int64_t>(1, ceil_div(target_ctas, units));
    int64_t SL = round_up(ceil_div(ndb, P0), p.TILE);
    if (p.umma_kp && ctas_mult == 1 && env_int("HG_UMMA_WAVES", 0) == 0) {
        // One CTA ranks 256 queries against a PAIR of adjacent splits; two CTAs are resident per SM and every CTA costs
        // (rows of its pair + a fixed start-up), so the launch takes ceil(CTAs / slots) rounds of the longest pair.
        // Pick the pair count with the shortest makespan instead of a fixed wave count: a grid of 4.05 waves would
        // leave the last 0.05 wave running alone.
        const int64_t slots = (int64_t)sms * (p.umma_kp > 128 ? 1 : 2);
        const int64_t startup_rows = 768;  // A-operand load + TMEM allocation + pipeline fill, in database rows
        double best = 1e300;
        for (int64_t cand = 1; cand <= 2 * P0 + 8; ++cand) {
            const int64_t pair = std::max<int64_t>(2 * p.TILE, round_up(ceil_div(ndb, cand), 2 * p.TILE));
            if (pair / 2 > kMaxSplitRows) continue;
            const int64_t np = ceil_div(ndb, pair);
            const int64_t rounds = ceil_div(units * np, slots);
            const double cost = (double)rounds * (double)(pair + startup_rows);
            if (cost < best * 0.995) { best = cost; SL = pair / 2; }  // prefer fewer splits unless clearly better
        }
    }
    {
        const int want = env_int("HG_SPLITS", 0);  // tuning override: number of database splits
        if (want > 0) SL = round_up(ceil_div(ndb, want), p.umma_kp ? 2 * p.TILE : p.TILE) ;
    }
    SL = std::min<int64_t>(SL, kMaxSplitRows);
    SL = std::max<int64_t>(SL, p.TILE);
    p.SL = SL;
    p.P = (int)ceil_div(ndb, SL);
    // candidate budget per query, spread evenly over the bins
    const int64_t all_rows = (int64_t)p.P * SL;
    int64_t capq = std::min<int64_t>(all_rows, std::max<int64_t>(6 * R, R + 16384));
    {
        // The budget is spread EVENLY over the bins, but a database stored class by class (the caller defines the row order,
        // lib/dataloader.py:93-94 only shuffles inside evaluate()) concentrates a query's candidates in the few splits that
        // hold its class.  Address space is cheap (only the entries really written cost bandwidth): up to 8x the budget while
        // the list area stays below 6 GB (C4: 5x), so that a class covering >= 4..6 % of the rows still fits its bins; anything
        // more concentrated takes the exact path (correct, slower; hg_hamming_map_stats reports the count).
        const int64_t mult = std::max<int64_t>(1, std::min<int64_t>(8, (int64_t)((6.0 * (double)(1ull << 30)) / (4.0 * (double)nq * (double)capq))));
        capq = std::min<int64_t>(all_rows, capq * mult);
    }
    int64_t cap = round_up(ceil_div(capq, p.P), 8);
    cap = std::min<int64_t>(cap, SL);
    while (cap * p.P < R) cap += 8;  // the exact path reuses the list area and needs R entries per query
    p.cap = (uint32_t)cap;
    {
        // The queued select (select_q_kernel) pays off while hits are sparse: at most ~2 per lane and 64-row tile (C4: 0.57); a dense
        // top-R (C2: 5.8) keeps the kernel that walks the hits from the staged tile.  HG_SELECT_MODE=lists / queue forces the choice.
        const char* mode = getenv("HG_SELECT_MODE");
        const bool sparse = 64.0 * 1.8 * (double)R <= 2.0 * (double)ndb;
        p.queued = p.umma_kp == 64 && p.W == 2 && p.Wr == 4 && (mode && mode[0] == 'l' ? false : (mode && mode[0] == 'q' ? true : sparse));
    }
    if ((double)nq * p.P * (double)cap >= 4294967295.0 || (double)nq * (double)R >= 4294967295.0)
        return p;  // bins are addressed with 32-bit word offsets: the caller must split the query batch
    // sample: kSampleTarget rows in TILE-row segments spread evenly; the whole db when it is small
    const int64_t tiles_total = ceil_div(ndb, p.TILE);
    int64_t want_seg = std::max<int64_t>(1, kSampleTarget / p.TILE);
    if (tiles_total <= 2 * want_seg || R * 4 >= ndb) {
        p.n_seg = tiles_total;
        p.seg_stride = p.TILE;
        p.sample_rows = ndb;
    } else {
        p.n_seg = want_seg;
        p.seg_stride = (tiles_total / want_seg) * p.TILE;
        p.sample_rows = want_seg * p.TILE;  // every sampled segment is a full tile by construction
    }
    {
        const int64_t hist_qt = ceil_div(nq, kHistThreads);
        int64_t chunks = std::max<int64_t>(1, std::min<int64_t>(p.n_seg, ceil_div((int64_t)sms * env_int("HG_HIST_CTAS_PER_SM", 4), hist_qt)));
        p.seg_per_chunk = (int)ceil_div(p.n_seg, chunks);
        p.n_chunks = (int)ceil_div(p.n_seg, p.seg_per_chunk);
    }
    // workspace
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t bins = (size_t)nq * p.P;
    p.off_ctrl = take(256);
    p.off_thr = take(sizeof(int) * nq);
    p.off_thr2 = take(sizeof(int) * nq);
    p.off_fail = take(sizeof(int) * nq);
    p.off_wide = take(sizeof(int) * nq);
    p.off_hist_s = take(sizeof(uint32_t) * (size_t)nq * (b + 1));
    p.off_bin_cnt = take(sizeof(uint32_t) * bins);
    p.off_bin_cnt0 = take(sizeof(uint32_t) * bins);
    p.off_bin_cnt2 = take(sizeof(uint32_t) * bins);
    p.off_bin_off2 = take(sizeof(uint32_t) * bins);
    p.off_bin_cap2 = take(sizeof(uint32_t) * bins);
    p.off_quota2 = take(sizeof(uint32_t) * bins);
    p.off_hist2 = take(sizeof(uint32_t) * bins * (b + 1));
    p.off_lists = take(sizeof(uint32_t) * bins * p.cap);
    if (p.umma_kp) {
        p.off_q8 = take((size_t)nq * p.umma_kp);
        p.off_db8 = take((size_t)round_up(ndb, 128) * p.umma_kp);
        p.off_qx = take((size_t)std::max<int64_t>(round_up(nq, 256), 128) * 32);
        p.off_bx = take((size_t)128 * 32);
    }
    p.total = off;
    p.ok = true;
    return p;
}

// ================================================================================================
// Device helpers
// ================================================================================================
// code words of one packed row (row = pointer to its first word; alignment guaranteed by hg_row_words)
template <int W>
__device__ __forceinline__ void load_code(const uint32_t* __restrict__ row, uint32_t (&v)[W])
{
    if constexpr (W == 1) {
        v[0] = row[0];
    } else if constexpr (W == 2) {
        const uint2 t = *reinterpret_cast<const uint2*>(row);
        v[0] = t.x; v[1] = t.y;
    } else if constexpr (W == 4) {
        const uint4 t = *reinterpret_cast<const uint4*>(row);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (W == 8) {
        const uint4 t0 = reinterpret_cast<const uint4*>(row)[0];
        const uint4 t1 = reinterpret_cast<const uint4*>(row)[1];
        v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w;
        v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
    } else {
#pragma unroll
        for (int w = 0; w < W; ++w) v[w] = row[w];
    }
}

template <int W>
__device__ __forceinline__ int hamming(const uint32_t (&a)[W], const uint32_t (&b)[W])
{
    int d = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) d += __popc(a[w] ^ b[w]);
    return d;
}

// ================================================================================================
// 1. Histogram kernel (sample pass and exact path).  Thread <-> one query; its histogram is a private
//    column of shared memory (bank = thread), so updates need no atomics.
// ================================================================================================
struct HistParams {
    const uint32_t* q_rows;
    const uint32_t* db_rows;
    int64_t nq, ndb;
    int b, Wr;
    const int* n_active;  // indirect mode (exact path): number of listed queries
    const int* qlist;     // indirect mode: query ids
    int64_t seg_stride, n_seg;
    int seg_rows, seg_per_chunk;
    uint32_t* out;   // [(slot * out_chunks + chunk') * (b+1) + d]
    int out_chunks;  // 1 -> all chunks accumulate into one histogram per query
};

// NARROW: 16-bit counters (a CTA sees fewer than 65536 rows: the sample pass) -- half the shared memory, so nine instead of
// five CTAs are resident per SM in this latency-bound kernel; updated by plain load/add/store (there is no 16-bit atomic).
template <int W, bool NARROW>
__global__ void __launch_bounds__(kHistThreads) hist_kernel(HistParams p)
{
    extern __shared__ __align__(16) uint32_t smem[];
    using CT = typename std::conditional<NARROW, uint16_t, uint32_t>::type;
    const int nbins = p.b + 1;
    const int Wr = p.Wr;
    uint32_t* tile = smem;                                               // [seg_rows * Wr], 16-byte aligned
    CT* h = reinterpret_cast<CT*>(smem + (size_t)p.seg_rows * Wr);       // [nbins][kHistThreads]
    const int tid = threadIdx.x;
    const int64_t n_act = p.qlist ? (int64_t)*p.n_active : p.nq;
    const int64_t slot0 = (int64_t)blockIdx.x * kHistThreads;
    if (slot0 >= n_act) return;
    const int64_t slot = slot0 + tid;
    const bool valid = slot < n_act;
    const int64_t q = valid ? (p.qlist ? (int64_t)p.qlist[slot] : slot) : 0;
    uint32_t qw[W];
#pragma unroll
    for (int w = 0; w < W; ++w) qw[w] = valid ? p.q_rows[q * Wr + w] : 0u;
    for (int i = tid; i < nbins * kHistThreads; i += kHistThreads) h[i] = 0;
    const int chunk = blockIdx.y;
    for (int g = 0; g < p.seg_per_chunk; ++g) {
        const int64_t seg = (int64_t)chunk * p.seg_per_chunk + g;
        if (seg >= p.n_seg) break;
        const int64_t row0 = seg * p.seg_stride;
        if (row0 >= p.ndb) break;
        const int rows = (int)min((int64_t)p.seg_rows, p.ndb - row0);
        __syncthreads();  // previous tile fully consumed (and h zeroed on the first trip)
        const uint4* src = reinterpret_cast<const uint4*>(p.db_rows + row0 * Wr);  // row0 is a tile multiple: 16-byte aligned
        const int n16 = (rows * Wr) >> 2;
        for (int i = tid; i < n16; i += kHistThreads) reinterpret_cast<uint4*>(tile)[i] = __ldg(src + i);
        for (int i = (n16 << 2) + tid; i < rows * Wr; i += kHistThreads) tile[i] = __ldg(p.db_rows + row0 * Wr + i);
        __syncthreads();
        if (valid) {
#pragma unroll 4
            for (int j = 0; j < rows; ++j) {
                uint32_t v[W];
                load_code<W>(tile + j * Wr, v);
                const int d = hamming<W>(qw, v);
                if (NARROW) {
                    h[d * kHistThreads + tid] = (CT)(h[d * kHistThreads + tid] + 1);
                } else {
                    // a reduction without a result (ATOMS.POPC.INC with no destination): the column is private to this thread, the
                    // point is that nothing waits for the value -- a plain `+= 1` chains load -> add -> store per row.  Measured
                    // on B200 (C4 sample pass): 0.164 -> 0.139 ms.  (The AP kernel's counters need the old value back; there the
                    // atomic with a result is 2x SLOWER than load/add/store -- measured 0.43 -> 0.85 ms -- so it keeps the plain form.)
                    atomicAdd(reinterpret_cast<uint32_t*>(h) + d * kHistThreads + tid, 1u);
                }
            }
        }
    }
    __syncthreads();
    // flush: consecutive threads -> consecutive distances of one query (coalesced atomics)
    const int oc = p.out_chunks > 1 ? chunk : 0;
    for (int i = tid; i < nbins * kHistThreads; i += kHistThreads) {
        const int ql = i / nbins, d = i - ql * nbins;
        if (slot0 + ql < n_act) {
            const uint32_t v = h[d * kHistThreads + ql];
            if (v) atomicAdd(&p.out[((slot0 + ql) * p.out_chunks + oc) * nbins + d], v);
        }
    }
}

// ================================================================================================
// 2. Threshold from the sampled histogram.
// ================================================================================================
// smallest distance whose cumulative count reaches `need` (b when it never does); one warp per query, coalesced bins
__device__ __forceinline__ int first_reaching(const uint32_t* __restrict__ h, int b, double need)
{
    const int lane = threadIdx.x & 31;
    double carry = 0.0;
    for (int d0 = 0; d0 <= b; d0 += 32) {
        const int d = d0 + lane;
        double v = d <= b ? (double)h[d] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, d <= b && carry + v >= need);
        if (hit) return d0 + __ffs(hit) - 1;
        carry += __shfl_sync(0xffffffffu, v, 31);
    }
    return b;
}

__global__ void __launch_bounds__(256) thr_kernel(const uint32_t* __restrict__ hist_s, int64_t nq, int b, int64_t sample_rows, int64_t ndb, int64_t R,
                                                  float z, int force_exact, int* __restrict__ thr)
{
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    if (force_exact) { if ((threadIdx.x & 31) == 0) thr[q] = -1; return; }
    double need;
    if (sample_rows >= ndb) {
        need = (double)R;  // the "sample" is the whole database: exact
    } else {
        const double p0 = (double)R / (double)ndb;
        const double mu = p0 * (double)sample_rows;
        need = ceil(mu + (double)z * sqrt(mu * (1.0 - p0)) + 2.0);
    }
    const int T = first_reaching(hist_s + q * (b + 1), b, need);
    if ((threadIdx.x & 31) == 0) thr[q] = T;
}

// ================================================================================================
// 3. The hot kernel: all-pairs XOR/POPC + threshold select into private, row-ordered bins.
// ================================================================================================
struct SelectParams {
    const uint32_t* q_rows;
    const uint32_t* db_rows;
    int64_t nq, ndb;
    const int* thr;       // [nq] by query id
    const int* n_active;  // EXACT: fail count
    const int* qlist;     // EXACT: fail list
    int P, Wr, LW, TILE;
    int split0;  // first database split of this launch (chunked host pipeline), grid.y splits from here
    int64_t SL, R;
    uint32_t* lists;
    uint32_t cap;              // fast path: bin = q*P + s at lists + bin*cap
    const uint32_t* bin_off2;  // EXACT: bin = f*P + s at lists + f*R + bin_off2[bin]
    const uint32_t* bin_cap2;
    const uint32_t* quota2;
    uint32_t* bin_cnt;   // out: candidates with d < T seen per bin (EXACT: all candidates); front + back may exceed the capacity -> overflow
    uint32_t* bin_cnt0;  // out (fast path): candidates with d == T seen per bin, stored from the end of the bin downwards
};

// row stride when the label fits one word (the common case: L <= 32) -- a compile-time constant so that a whole
// packed row (code words + label word) is fetched with the widest shared-memory loads
template <int W> struct RowLW1 { static constexpr int Wr = (W == 1 ? 2 : (W == 2 ? 4 : (W == 3 ? 4 : (W == 4 ? 8 : 12)))); };

// code words + first label word of one packed row in shared memory
template <int W>
__device__ __forceinline__ void load_row_lw1(const uint32_t* __restrict__ row, uint32_t (&v)[W], uint32_t& lab)
{
    if constexpr (W == 1) {
        const uint2 t = *reinterpret_cast<const uint2*>(row);
        v[0] = t.x; lab = t.y;
    } else if constexpr (W == 2) {
        const uint4 t = *reinterpret_cast<const uint4*>(row);
        v[0] = t.x; v[1] = t.y; lab = t.z;
    } else if constexpr (W == 3) {
        const uint4 t = *reinterpret_cast<const uint4*>(row);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; lab = t.w;
    } else {
        load_code<W>(row, v);
        lab = row[W];
    }
}

// LW1 = true : L <= 32, row stride RowLW1<W>::Wr, the relevance bit costs one AND on registers in the rare path
// LW1 = false: any label width, runtime row stride, label words read from the tile in the rare path
template <int W, int QT, bool EXACT, bool LW1>
__global__ void __launch_bounds__(kSelectThreads) select_kernel(SelectParams p)
{
    constexpr int NT = kSelectThreads;
    constexpr int TQ = NT * QT;
    constexpr int U = (W <= 2 ? 4 : 2);
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ __align__(8) uint64_t s_full[2];
    const int Wr = LW1 ? RowLW1<W>::Wr : p.Wr;
    const int LW = LW1 ? 1 : p.LW;
    const int TILE = p.TILE;
    const int tile_words = TILE * Wr;          // smem: [2][tile_words] database tiles, then (generic labels) [LW][TQ]
    uint32_t* s_qlab = smem + 2 * tile_words;  // only used when !LW1

    const int tid = threadIdx.x;
    const int split = p.split0 + (int)blockIdx.y;
    const int64_t n_act = EXACT ? (int64_t)*p.n_active : p.nq;
    const int64_t slot0 = (int64_t)blockIdx.x * TQ;
    if (slot0 >= n_act) return;

    const int64_t row0 = (int64_t)split * p.SL;
    const int64_t row1 = min(row0 + p.SL, p.ndb);
    const int64_t nrows = row1 > row0 ? row1 - row0 : 0;
    const int ntiles = (int)((nrows + TILE - 1) / TILE);

    // per-query state: code words, label word, threshold, and the private bin as 32-bit word offsets into p.lists
    // (make_plan keeps the whole list area below 2^32 entries): pos counts every candidate seen, stores stop at end.
    uint32_t qw[QT][W];
    uint32_t qlab[QT];
    int T[QT];
    uint32_t pos[QT], start[QT], end[QT];
    uint32_t neq[QT], quota[QT];
    int64_t bin[QT];
#pragma unroll
    for (int k = 0; k < QT; ++k) {
        const int64_t slot = slot0 + (int64_t)k * NT + tid;
        const bool valid = slot < n_act;
        const int64_t q = valid ? (EXACT ? (int64_t)p.qlist[slot] : slot) : 0;
#pragma unroll
        for (int w = 0; w < W; ++w) qw[k][w] = valid ? p.q_rows[q * Wr + w] : 0u;
        qlab[k] = valid ? p.q_rows[q * Wr + W] : 0u;
        if (!LW1)
            for (int w = 0; w < LW; ++w) s_qlab[w * TQ + k * NT + tid] = valid ? p.q_rows[q * Wr + W + w] : 0u;
        T[k] = valid ? p.thr[q] : -1;
        bin[k] = valid ? slot * p.P + split : -1;
        neq[k] = 0;
        if (EXACT) {
            start[k] = valid ? (uint32_t)(slot * p.R + (int64_t)p.bin_off2[bin[k]]) : 0u;
            end[k] = start[k] + (valid ? p.bin_cap2[bin[k]] : 0u);
            quota[k] = valid ? p.quota2[bin[k]] : 0u;
        } else {
            start[k] = valid ? (uint32_t)(bin[k] * (int64_t)p.cap) : 0u;
            end[k] = start[k] + (valid ? p.cap : 0u);
            quota[k] = 0xffffffffu;
        }
        pos[k] = start[k];
    }
    uint32_t* const lists = p.lists;

    if (tid == 0) {
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        fence_barrier_init();
    }
    __syncthreads();

    const uint32_t* src0 = p.db_rows + row0 * Wr;
    const uint32_t tile_bytes = (uint32_t)tile_words * 4;
    auto tile_rows = [&](int t) -> int { return (int)min((int64_t)TILE, nrows - (int64_t)t * TILE); };
    auto issue = [&](int t) {
        const int buf = t & 1;
        const int rows = tile_rows(t);
        const uint32_t* src = src0 + (int64_t)t * tile_words;
        if (rows == TILE) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&s_full[buf], tile_bytes);
                tma_load_1d(smem + buf * tile_words, src, tile_bytes, &s_full[buf]);
            }
        } else {
            uint32_t* dst = smem + buf * tile_words;
            for (int i = tid; i < rows * Wr; i += NT) dst[i] = __ldg(src + i);
        }
    };
    if (ntiles > 0) issue(0);
    if (ntiles > 1) issue(1);
    __syncthreads();  // ragged tiles written with plain stores become visible

    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const int rows = tile_rows(t);
        if (rows == TILE) mbar_wait(&s_full[buf], (uint32_t)((t >> 1) & 1));
        const uint32_t* tile = smem + buf * tile_words;
        const uint32_t lbase = (uint32_t)t * TILE;

        // rare path (about R/Ndb of the pairs): append (relevance, distance, row) to this thread's private bin
        auto emit = [&](int k, int d, uint32_t lab, const uint32_t* row, uint32_t lidx) {
            bool ok = true;
            if (EXACT) {
                if (d == T[k]) { ok = neq[k] < quota[k]; neq[k]++; }
            }
            if (ok) {
                // fast path: rows at the threshold distance (d == T) are stacked downwards from the end of the bin, closer
                // rows upwards from its start, so the AP kernel can stream the d == T class and stop at the quota
                const bool eq = !EXACT && d == T[k];
                const uint32_t front = pos[k], back = EXACT ? 0u : neq[k];
                if (front + back < end[k]) {
                    uint32_t m = lab & qlab[k];
                    if (!LW1) {
#pragma unroll 1
                        for (int w = 1; w < LW; ++w) m |= row[W + w] & s_qlab[w * TQ + k * NT + tid];
                    }
                    lists[eq ? end[k] - 1u - back : front] = ((uint32_t)d * (1u << kIdxBits) + lidx) | (m ? 0x80000000u : 0u);
                }
                if (eq) neq[k] = back + 1; else pos[k] = front + 1;
            }
        };

        int j = 0;
        for (; j + U <= rows; j += U) {
            uint32_t v[U][W], lab[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (LW1) load_row_lw1<W>(tile + (j + u) * Wr, v[u], lab[u]);
                else { load_code<W>(tile + (j + u) * Wr, v[u]); lab[u] = 0; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int k = 0; k < QT; ++k) {
                    const int d = hamming<W>(qw[k], v[u]);
                    if (d <= T[k]) {
                        const uint32_t* row = tile + (j + u) * Wr;
                        emit(k, d, LW1 ? lab[u] : row[W], row, lbase + j + u);
                    }
                }
            }
        }
        for (; j < rows; ++j) {
            const uint32_t* row = tile + j * Wr;
            uint32_t v[W];
            load_code<W>(row, v);
#pragma unroll
            for (int k = 0; k < QT; ++k) {
                const int d = hamming<W>(qw[k], v);
                if (d <= T[k]) emit(k, d, row[W], row, lbase + j);
            }
        }
        __syncthreads();  // everyone is done with this buffer
        if (t + 2 < ntiles) issue(t + 2);
        // a ragged tile (only ever the last one) is ordered by the __syncthreads of the next trip
    }

#pragma unroll
    for (int k = 0; k < QT; ++k)
        if (bin[k] >= 0) {
            p.bin_cnt[bin[k]] = pos[k] - start[k];
            if (!EXACT) p.bin_cnt0[bin[k]] = neq[k];
        }
}

// ================================================================================================
// 4. AP kernel.  G threads per query, each walks a contiguous range of the query's bins SEQUENTIALLY (bins and
//    entries are in database-row order).  Per-thread private counters per distance (a shared-memory column per
//    thread, bank == thread: no atomics, no conflicts):
//      phase A  count entries / relevant entries per distance in my range;
//      phase A2 per distance: exclusive scan over the G threads of the query + running totals over distances
//               -> my counters now hold "rank before my first entry of that distance" (N_d + earlier ranges);
//      phase B  walk again: rank = ++N[d], cum = (M[d] += relevant); an entry is in the top-R iff rank <= R;
//               relevant entries add cum / rank (fp64); optional ids / dist output at position rank-1.
//    WINDOW = true : 32 counters per thread indexed by (distance & 31) -- exact whenever the candidate distances of
//               a query span fewer than 32 values (always, on hash codes: the top-R tail of a binomial is a few
//               sigma wide); 256 B of shared memory per thread -> 6 CTAs per SM.  Queries with a wider span are
//               put on the wide list and redone by the WINDOW = false variant (b+1 counters per thread).
// ================================================================================================
struct ApParams {
    int64_t nq;
    const int* n_active;  // indirect modes: number of listed queries
    const int* qlist;     // indirect modes: query ids (wide list or fail list)
    int bins_by_slot;     // exact path: bins/offsets are indexed by the position in qlist, not by the query id
    int P, b, G;
    int64_t SL, R;
    const uint32_t* lists;
    uint32_t cap;
    const uint32_t* bin_off2;  // exact path
    const uint32_t* bin_cap2;  // exact path
    const uint32_t* bin_cnt;
    const uint32_t* bin_cnt0;  // fast path: d == T class stacked from the end of each bin (null: single class)
    const int* thr;
    double* ap;
    uint32_t* ids;
    uint16_t* dist;
    int32_t* rel;
    int* fail_list;  // queries to redo exactly (fast path only)
    int* n_fail;
    int* wide_list;  // queries whose distance span does not fit the window (WINDOW only)
    int* n_wide;
    int no_fallback;
    int halves;  // the lanes of a query share half bins instead of whole bins (set by the launcher)
};

// Error-free accumulation (Knuth TwoSum): hi + lo carries the AP sum to ~2^-100, so the rounded result does not depend
// on how the candidates of a query are split over threads (G, P) -- query chunking and the launch shape stay invisible.
__device__ __forceinline__ void dd_add(double& hi, double& lo, double x)
{
    const double s = __dadd_rn(hi, x);
    const double bb = __dsub_rn(s, hi);
    lo = __dadd_rn(lo, __dadd_rn(__dsub_rn(hi, __dsub_rn(s, bb)), __dsub_rn(x, bb)));
    hi = s;
}

// precision-at-rank term cum / rank of lib/metric.py:21 as cum * (1 / rank): the correctly rounded reciprocal of an
// integer is cheaper than the IEEE divide; the product is within one ulp (1.1e-16 relative) of the quotient.
__device__ __forceinline__ double ap_term(uint32_t cum, uint32_t rank) { return __dmul_rn((double)cum, __drcp_rn((double)rank)); }

template <bool WINDOW>
__global__ void __launch_bounds__(128) ap_kernel(ApParams p)
{
    extern __shared__ __align__(16) uint32_t smem[];
    // WINDOW: the two 16-bit counters of a distance share ONE shared-memory word, relevant count << 16 | count (ranks below 65536:
    // a query with more closer-than-threshold candidates goes to the wide list) -- half the shared memory per thread, so more CTAs
    // are resident in this latency-bound kernel, and one load / add / store per candidate and one shuffle per scan step instead of two
    using CT = uint32_t;
    const int NTB = blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int nb = p.b + 1;
    const int ncol = WINDOW ? 32 : nb;               // counters per thread and kind
    CT* cN = reinterpret_cast<CT*>(smem) + tid;      // cN[c * NTB]: count per distance  -> rank base
    CT* cM = cN + (size_t)ncol * NTB;                // cM[c * NTB]: relevant per distance -> relevant base (!WINDOW only)
    const int G = p.G;
    const bool exact = p.bins_by_slot != 0;
    const int64_t n_act = p.qlist ? (int64_t)*p.n_active : p.nq;
    const int64_t slot = ((int64_t)blockIdx.x * NTB + tid) / G;
    const int g = tid % G;  // groups are aligned inside a warp because G divides 32
    const uint32_t FULL = 0xffffffffu;
    if ((int64_t)blockIdx.x * NTB / G >= n_act) return;  // whole CTA beyond the active queries
    bool live = slot < n_act;
    const int64_t q = live ? (p.qlist ? (int64_t)p.qlist[slot] : slot) : 0;
    const int64_t binbase = (exact ? slot : q) * p.P;
    const int T = live ? p.thr[q] : -1;
    const int s0 = (int)((int64_t)g * p.P / G), s1 = (int)((int64_t)(g + 1) * p.P / G);

    for (int c = 0; c < ncol; ++c) {
        cN[c * NTB] = 0;
        if (!WINDOW) cM[c * NTB] = 0;
    }

    // ---- totals / overflow over the query's bins (group reduction) -------------------------
    unsigned long long total = 0, total1 = 0;  // all candidates / candidates closer than the threshold distance
    int ovf = 0;
    if (live) {
        for (int s = s0; s < s1; ++s) {
            const uint32_t c1 = p.bin_cnt[binbase + s];
            const uint32_t c0 = p.bin_cnt0 ? p.bin_cnt0[binbase + s] : 0u;
            const uint32_t capb = exact ? p.bin_cap2[binbase + s] : p.cap;
            ovf |= ((unsigned long long)c1 + c0 > capb);
            total += (unsigned long long)c1 + c0;
            total1 += c1;
        }
    }
    for (int o = 1; o < G; o <<= 1) {
        total += __shfl_xor_sync(FULL, total, o);
        total1 += __shfl_xor_sync(FULL, total1, o);
        ovf |= __shfl_xor_sync(FULL, ovf, o);
    }
    if (live && (ovf || total < (unsigned long long)p.R || T < 0)) {
        if (g == 0) {
            if (p.fail_list != nullptr && !p.no_fallback) {
                const int i = atomicAdd(p.n_fail, 1);
                p.fail_list[i] = (int)q;
            } else {
                p.ap[q] = -1.0;  // diagnostics only: never reached with the fallback enabled
                if (p.n_fail) atomicAdd(p.n_fail, 1);
            }
        }
        live = false;
    }

    auto bin_ptr = [&](int s) -> const uint32_t* {
        return exact ? p.lists + slot * p.R + (int64_t)p.bin_off2[binbase + s] : p.lists + (binbase + s) * (int64_t)p.cap;
    };
    auto col = [&](uint32_t d) -> uint32_t { return (WINDOW ? (d & 31u) : d) * (uint32_t)NTB; };

    // walks my bins in row order; 32 bytes of entries in flight per thread while the previous 32 are consumed
    // The lanes share HALF bins (first / second half of a bin's entries, split at a multiple of 8 entries so that the 16-byte loads
    // stay aligned): with P bins over G lanes a lane owned 1 or 2 bins (C4: 44 bins over 32 lanes, 69 % balance), now 2 or 3 halves.
    // (only where that improves the balance: p.halves, chosen by the launcher)
    const int hsh = p.halves ? 1 : 0;
    const int h0 = (int)((int64_t)g * (p.P << hsh) / G), h1 = (int)((int64_t)(g + 1) * (p.P << hsh) / G);
    auto walk = [&](auto&& fn) {
        for (int hb = h0; hb < h1; ++hb) {
            const int s = hb >> hsh;
            const uint32_t cnt = p.bin_cnt[binbase + s];
            const uint32_t mid = hsh ? ((cnt >> 1) & ~7u) : 0u;
            const uint32_t c = (hsh && !(hb & 1)) ? mid : cnt;  // entries [e, c) of the bin
            uint32_t e = (hsh && (hb & 1)) ? mid : 0u;
            const uint32_t* lp = bin_ptr(s);
            const int64_t row0 = (int64_t)s * p.SL;
            if ((reinterpret_cast<uintptr_t>(lp) & 15) == 0 && c >= e + 8) {
                const uint4* vp = reinterpret_cast<const uint4*>(lp);
                uint4 a0 = __ldg(vp + (e >> 2)), a1 = __ldg(vp + (e >> 2) + 1);
                for (; e + 8 <= c; e += 8) {
                    uint4 n0 = a0, n1 = a1;
                    if (e + 16 <= c) { n0 = __ldg(vp + (e >> 2) + 2); n1 = __ldg(vp + (e >> 2) + 3); }
                    fn(a0.x, row0); fn(a0.y, row0); fn(a0.z, row0); fn(a0.w, row0);
                    fn(a1.x, row0); fn(a1.y, row0); fn(a1.z, row0); fn(a1.w, row0);
                    a0 = n0; a1 = n1;
                }
            }
            for (; e < c; ++e) fn(__ldg(lp + e), row0);
        }
    };

    // ---- phase A: private histograms (+ distance span for the window variant) -------------------
    uint32_t dmin = 0xffffffffu, dmax = 0;
    if (live) {
        walk([&](uint32_t ent, int64_t) {
            const uint32_t d = (ent >> kIdxBits) & kDistMask;
            const uint32_t c = col(d);
            if (WINDOW) {
                cN[c] += 1u + ((ent >> 15) & 0x10000u);
                dmin = min(dmin, d); dmax = max(dmax, d);
            } else {
                cN[c] += 1u;
                cM[c] += ent >> 31;
            }
        });
    }
    if (WINDOW) {
        for (int o = 1; o < G; o <<= 1) {
            dmin = min(dmin, __shfl_xor_sync(FULL, dmin, o));
            dmax = max(dmax, __shfl_xor_sync(FULL, dmax, o));
        }
        if (live && (dmax - dmin >= 32u || total1 >= 65536ull)) {
            if (g == 0) {
                const int i = atomicAdd(p.n_wide, 1);
                p.wide_list[i] = (int)q;
            }
            live = false;
        }
        if (!live) dmin = 0;
    }

    // ---- phase A2: rank bases, distances in ascending order (all 32 lanes take part in the shuffles) ----
    if (WINDOW) {  // both halves of the packed word at once: live queries keep every partial sum below 65536
        // only the distances that occur: dmin .. dmax of the queries of this warp (typically a dozen of the 32 window slots)
        const int span = (int)__reduce_max_sync(FULL, live ? dmax - dmin + 1u : 0u);
        uint32_t cn = 0;
        for (int i = 0; i < span; ++i) {
            const uint32_t c = col(dmin + (uint32_t)i);
            const uint32_t v = cN[c];
            uint32_t iv = v;
            for (int o = 1; o < G; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, iv, o, G);
                if (g >= o) iv += t;
            }
            const uint32_t tot = __shfl_sync(FULL, iv, G - 1, G);
            cN[c] = cn + (iv - v);
            cn += tot;
        }
    } else {
        uint32_t cn = 0, cm = 0;
        for (int i = 0; i < ncol; ++i) {
            const uint32_t c = col((uint32_t)i);
            const uint32_t vN = cN[c], vM = cM[c];
            uint32_t iN = vN, iM = vM;
            for (int o = 1; o < G; o <<= 1) {
                const uint32_t tN = __shfl_up_sync(FULL, iN, o, G);
                const uint32_t tM = __shfl_up_sync(FULL, iM, o, G);
                if (g >= o) { iN += tN; iM += tM; }
            }
            const uint32_t totN = __shfl_sync(FULL, iN, G - 1, G);
            const uint32_t totM = __shfl_sync(FULL, iM, G - 1, G);
            cN[c] = cn + (iN - vN);
            cM[c] = cm + (iM - vM);
            cn += totN; cm += totM;
        }
    }

    // ---- phase B: ranks, relevant prefix counts, AP ------------------------------------------------
    double acc = 0.0, acc_lo = 0.0;
    int relc = 0;
    const uint32_t R32 = (uint32_t)p.R;
    if (live) {
        walk([&](uint32_t ent, int64_t row0) {
            const uint32_t d = (ent >> kIdxBits) & kDistMask;
            const uint32_t m = ent >> 31;
            const uint32_t c = col(d);
            uint32_t rank, cum;
            if (WINDOW) {
                const uint32_t v = cN[c] + 1u + (m << 16);
                cN[c] = v;
                rank = v & 0xFFFFu; cum = v >> 16;
            } else {
                rank = cN[c] + 1u; cum = cM[c] + m;
                cN[c] = rank; cM[c] = cum;
            }
            if (rank <= R32) {
                if (p.ids) p.ids[q * p.R + (rank - 1u)] = (uint32_t)(row0 + (ent & kIdxMask));
                if (p.dist) p.dist[q * p.R + (rank - 1u)] = (uint16_t)d;
                if (m) {
                    dd_add(acc, acc_lo, ap_term(cum, rank));
                    relc += 1;
                }
            }
        });
    }
    // relevant rows among the closer-than-threshold class (all of them are inside the top-R when total1 < R)
    for (int o = 1; o < G; o <<= 1) relc += __shfl_xor_sync(FULL, relc, o);

    // ---- part 2: the d == T class, streamed in row order until the top-R is full ---------------------------------
    int relc0 = 0;
    if (p.bin_cnt0 != nullptr) {
        const bool need0 = live && total1 < (unsigned long long)p.R;
        const uint32_t quota = need0 ? (uint32_t)(p.R - (int64_t)total1) : 0u;
        const uint32_t base_rank = (uint32_t)total1, base_cum = (uint32_t)relc;
        uint32_t n_done = 0, m_done = 0;
        const uint32_t gmask = G >= 32 ? FULL : ((1u << G) - 1u);
        const int gshift = lane - g;  // first lane of my group
        // the class counts of G bins at a time (one coalesced load per lane, handed round by shuffle) instead of one dependent load per
        // bin inside this serial loop (1250 queries over 237 splits, the N = 8 strong-scaling share: 0.122 -> 0.116 ms)
        for (int sb = 0; sb < p.P; sb += G) {
        if (!__any_sync(FULL, need0 && n_done < quota)) break;  // every query of the warp has its top-R
        const uint32_t cblk = (need0 && sb + g < p.P) ? p.bin_cnt0[binbase + sb + g] : 0u;
        const int nblk = min(G, p.P - sb);
        for (int j = 0; j < nblk; ++j) {
            const int s = sb + j;
            const uint32_t c0 = __shfl_sync(FULL, cblk, gshift + j);
            const uint32_t* back = need0 ? bin_ptr(s) + (p.cap - 1u) : nullptr;  // entry i of the class (row order) lives at back - i
            const int64_t row0 = (int64_t)s * p.SL;
            for (uint32_t i0 = 0;; i0 += (uint32_t)G) {
                const bool more = i0 < c0 && n_done < quota;
                if (!__any_sync(FULL, more)) break;
                const uint32_t i = i0 + (uint32_t)g;
                const bool act = more && i < c0;
                const uint32_t ent = act ? __ldg(back - i) : 0u;
                const uint32_t relbit = act ? (ent >> 31) : 0u;
                const uint32_t gb = (__ballot_sync(FULL, relbit != 0u) >> gshift) & gmask;  // relevance bits of my group
                if (more) {
                    const uint32_t rank = base_rank + n_done + (uint32_t)g + 1u;
                    if (act && rank <= R32) {
                        if (p.ids) p.ids[q * p.R + (rank - 1u)] = (uint32_t)(row0 + (ent & kIdxMask));
                        if (p.dist) p.dist[q * p.R + (rank - 1u)] = (uint16_t)((ent >> kIdxBits) & kDistMask);
                        if (relbit) {
                            const uint32_t cum = base_cum + m_done + (uint32_t)__popc(gb & ((1u << g) - 1u)) + 1u;
                            dd_add(acc, acc_lo, ap_term(cum, rank));
                            relc0 += 1;
                        }
                    }
                    n_done += min((uint32_t)G, c0 - i0);
                    m_done += (uint32_t)__popc(gb);
                }
            }
        }
        }
    }
    // fixed-order reduction over the G lanes of the query (deterministic)
    for (int o = 1; o < G; o <<= 1) {
        const double ohi = __shfl_xor_sync(FULL, acc, o), olo = __shfl_xor_sync(FULL, acc_lo, o);
        dd_add(acc, acc_lo, ohi);
        acc_lo = __dadd_rn(acc_lo, olo);
        relc0 += __shfl_xor_sync(FULL, relc0, o);
    }
    relc += relc0;
    acc = __dadd_rn(acc, acc_lo);
    if (live && g == 0) {
        p.ap[q] = relc ? acc / (double)relc : __longlong_as_double(0x7ff8000000000000LL);
        if (p.rel) p.rel[q] = relc;
    }
}

// ================================================================================================
// 4b. Dense top-R (R a large part of the database; cifar_evaluation.yaml ranks the WHOLE database: MAP_R == DB_SIZE).
//     Every row is a candidate, so nothing is selected and nothing is written: the AP walk of ap_kernel runs directly over
//     the packed rows, the Hamming distance and the relevance bit of a row are recomputed in both passes (one XOR + POPC per
//     word) instead of being stored as 54 M candidate entries and read back twice (C1: select 0.46 ms + AP 0.52 ms).
//     One CTA per query, thread t owns the t-th contiguous row range (thread order == row order, as in ap_kernel);
//     private per-distance counters -> exact (distance, row) ranks; rows ranked beyond R contribute nothing.
// ================================================================================================
struct DenseApParams {
    const uint32_t* q_rows;
    const uint32_t* db_rows;
    int64_t nq, ndb, R;
    int b, LW, Wr, G;
    double* ap;
    uint32_t* ids;
    uint16_t* dist;
    int32_t* rel;
    const int* n_active;  // indirect mode (exact path of a few failed queries): number of listed queries
    const int* qlist;     // indirect mode: query ids; the grid is sized for the largest list, CTAs beyond it exit at once
};

// one CTA = one query; thread t owns the t-th of blockDim.x contiguous row ranges.
// LW1 (label fits one word: the row stride is a compile-time constant): a thread fetches four rows at a time (two or four
// 16-byte loads in flight; one row for the 8- and 12-word rows) -- the threads of a warp stream 32 different ranges, so every load instruction touches 32
// cache lines and costs 32 tag look-ups; row-at-a-time loads made the kernel L1-tag bound (1.0 ms at C1).
template <int W, bool LW1>
__global__ void __launch_bounds__(256) dense_ap_kernel(DenseApParams p)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int NTB = blockDim.x, NW = NTB >> 5;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = p.b + 1;
    uint32_t* cN = smem + tid;                      // cN[d * NTB]: rows at distance d -> rank base
    uint32_t* cM = smem + (size_t)nb * NTB + tid;   // cM[d * NTB]: relevant rows at distance d -> relevant base
    uint2* wtot = reinterpret_cast<uint2*>(smem + (size_t)2 * nb * NTB);  // [nb][NW]: per-warp totals of a distance
    const uint32_t FULL = 0xffffffffu;
    if (p.qlist != nullptr && (int)blockIdx.x >= *p.n_active) return;  // whole CTA
    const int64_t q = p.qlist ? (int64_t)p.qlist[blockIdx.x] : (int64_t)blockIdx.x;
    const int Wr = LW1 ? RowLW1<W>::Wr : p.Wr, LW = LW1 ? 1 : p.LW;
    constexpr int RB = LW1 ? (RowLW1<W>::Wr <= 4 ? 4 : 1) : 1;  // rows per batch: 4 (32 bytes of 2-word rows, 64 bytes of 4-word rows: measured best at C1 / C1_64) or 1
    uint32_t qw[W], ql[4] = {0, 0, 0, 0};
#pragma unroll
    for (int w = 0; w < W; ++w) qw[w] = p.q_rows[q * Wr + w];
    for (int w = 0; w < LW && w < 4; ++w) ql[w] = p.q_rows[q * Wr + W + w];
    for (int c = 0; c < nb; ++c) { cN[c * NTB] = 0; cM[c * NTB] = 0; }
    // ranges start at multiples of RB rows so that the batches are 16-byte aligned
    const int64_t lo = ((int64_t)tid * p.ndb / NTB) / RB * RB;
    const int64_t hi = tid + 1 == NTB ? p.ndb : ((int64_t)(tid + 1) * p.ndb / NTB) / RB * RB;

    auto words_entry = [&](const uint32_t* w, uint32_t& d, uint32_t& m) {  // w: the row's words (registers when LW1)
        uint32_t v[W];
#pragma unroll
        for (int i = 0; i < W; ++i) v[i] = w[i];
        d = (uint32_t)hamming<W>(qw, v);
        uint32_t mm = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < LW) mm |= ql[i] & w[W + i];
        m = mm ? 1u : 0u;
    };
    // walks my rows in order; fn(row, d, m)
    auto walk = [&](auto&& fn) {
        int64_t r = lo;
        if constexpr (LW1) {
            constexpr int WrC = RowLW1<W>::Wr;
            constexpr int NV = WrC * RB / 4;  // 16-byte loads per batch (4 or 2; 3 for the 12-word rows of b > 128)
            for (; r + RB <= hi; r += RB) {
                uint32_t buf[NV * 4];
                const uint4* src = reinterpret_cast<const uint4*>(p.db_rows + r * WrC);
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const uint4 t = __ldg(src + i);
                    buf[4 * i] = t.x; buf[4 * i + 1] = t.y; buf[4 * i + 2] = t.z; buf[4 * i + 3] = t.w;
                }
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    uint32_t d, m;
                    words_entry(buf + j * WrC, d, m);
                    fn(r + j, d, m);
                }
            }
        }
        for (; r < hi; ++r) {
            uint32_t wbuf[12];
            const uint32_t* row = p.db_rows + r * Wr;
            for (int i = 0; i < W + LW; ++i) wbuf[i] = __ldg(row + i);
            uint32_t d, m;
            words_entry(wbuf, d, m);
            fn(r, d, m);
        }
    };

    // ---- pass A: private histograms ----
    walk([&](int64_t, uint32_t d, uint32_t m) {
        cN[d * NTB] += 1u;
        cM[d * NTB] += m;
    });
    // ---- rank bases: distances in ascending order, threads of the query in row order.  Per distance: exclusive scan inside
    //      the warp (kept in the counter), warp totals through shared memory, then the totals of the closer distances ----
    for (int i = 0; i < nb; ++i) {
        const uint32_t vN = cN[i * NTB], vM = cM[i * NTB];
        uint32_t iN = vN, iM = vM;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t tN = __shfl_up_sync(FULL, iN, o);
            const uint32_t tM = __shfl_up_sync(FULL, iM, o);
            if (lane >= o) { iN += tN; iM += tM; }
        }
        cN[i * NTB] = iN - vN;
        cM[i * NTB] = iM - vM;
        if (lane == 31) wtot[i * NW + warp] = make_uint2(iN, iM);
    }
    __syncthreads();
    {
        uint32_t cn = 0, cm = 0;
        for (int i = 0; i < nb; ++i) {
            uint32_t bN = cn, bM = cm;
            for (int w = 0; w < NW; ++w) {
                const uint2 t = wtot[i * NW + w];
                if (w < warp) { bN += t.x; bM += t.y; }
                cn += t.x; cm += t.y;
            }
            cN[i * NTB] += bN;
            cM[i * NTB] += bM;
        }
    }
    // ---- pass B: ranks, relevant prefix counts, AP ----
    double acc = 0.0, acc_lo = 0.0;
    int relc = 0;
    const uint32_t R32 = (uint32_t)p.R;
    walk([&](int64_t r, uint32_t d, uint32_t m) {
        const uint32_t rank = cN[d * NTB] + 1u;
        const uint32_t cum = cM[d * NTB] + m;
        cN[d * NTB] = rank;
        cM[d * NTB] = cum;
        if (rank <= R32) {
            if (p.ids) p.ids[q * p.R + (rank - 1u)] = (uint32_t)r;
            if (p.dist) p.dist[q * p.R + (rank - 1u)] = (uint16_t)d;
            if (m) {
                dd_add(acc, acc_lo, ap_term(cum, rank));
                relc += 1;
            }
        }
    });
    // fixed-order reduction (deterministic): lanes by xor tree, then the warps in order
    for (int o = 1; o < 32; o <<= 1) {
        const double ohi = __shfl_xor_sync(FULL, acc, o), olo = __shfl_xor_sync(FULL, acc_lo, o);
        dd_add(acc, acc_lo, ohi);
        acc_lo = __dadd_rn(acc_lo, olo);
        relc += __shfl_xor_sync(FULL, relc, o);
    }
    __syncthreads();  // wtot is reused for the warp results
    double* red = reinterpret_cast<double*>(wtot);  // [NW][3]
    if (lane == 0) { red[warp * 3] = acc; red[warp * 3 + 1] = acc_lo; red[warp * 3 + 2] = (double)relc; }
    __syncthreads();
    if (tid == 0) {
        double hi_ = 0.0, lo_ = 0.0, rc = 0.0;
        for (int w = 0; w < NW; ++w) { dd_add(hi_, lo_, red[w * 3]); lo_ = __dadd_rn(lo_, red[w * 3 + 1]); rc += red[w * 3 + 2]; }
        const double sum = __dadd_rn(hi_, lo_);
        p.ap[q] = rc > 0.0 ? sum / rc : __longlong_as_double(0x7ff8000000000000LL);
        if (p.rel) p.rel[q] = (int32_t)rc;
    }
}

template <int W>
static int launch_dense_ap(DenseApParams dp, cudaStream_t st)
{
    const int nb = dp.b + 1;
    // threads per query: as many as the shared-memory counters allow (<= 256), but at least a few dozen rows per thread
    int threads = 256;
    auto smem_for = [&](int t) { return (size_t)t * 2 * nb * sizeof(uint32_t) + std::max<size_t>((size_t)nb * (t / 32) * sizeof(uint2), (size_t)(t / 32) * 3 * sizeof(double)) + 16; };
    // (a short list of failed queries leaves most SMs empty anyway: take all the threads one CTA can have)
    const size_t smem_cap = dp.qlist ? 200 * 1024 : 100 * 1024;
    while (threads > 32 && (smem_for(threads) > smem_cap || dp.ndb / threads < 32)) threads >>= 1;
    const size_t smem = smem_for(threads);
    dp.G = threads;
    if (dp.LW == 1) {
        static thread_local size_t configured = 0;
        if (smem > 48 * 1024 && smem > configured) {
            HG_CUDA_TRY(cudaFuncSetAttribute(dense_ap_kernel<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        dense_ap_kernel<W, true><<<(unsigned)dp.nq, threads, smem, st>>>(dp);
    } else {
        static thread_local size_t configured = 0;
        if (smem > 48 * 1024 && smem > configured) {
            HG_CUDA_TRY(cudaFuncSetAttribute(dense_ap_kernel<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        dense_ap_kernel<W, false><<<(unsigned)dp.nq, threads, smem, st>>>(dp);
    }
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

// ================================================================================================
// 5. Exact path helpers.
// ================================================================================================
// The failed queries of a call go ONE of two ways, decided on the device (the call stays asynchronous): up to `dense_max` of
// them are ranked by dense_ap_kernel, one CTA per query walking the whole database twice -- the per-split machinery below is
// thread-per-query and leaves the GPU idle for a handful of queries (26 failed queries of a class-sorted 1M-row database:
// 5.1 ms, against 0.3 ms for the walk); larger lists (HG_FLAG_FORCE_EXACT on a big batch, degenerate codes) keep the
// per-split path, whose database tiles are shared by 128 queries.  ctrl[4] / ctrl[5] = the count each path sees.
// The routing costs no launch of its own: zero_hist2_kernel (first kernel of the per-split path) derives its count from ctrl[0] and
// publishes both counts for the kernels behind it.
__global__ void zero_hist2_kernel(uint32_t* __restrict__ hist2, int* __restrict__ ctrl, int dense_max, int64_t per_query)
{
    const int n_all = ctrl[0];
    const int small = n_all <= dense_max ? n_all : 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctrl[4] = small; ctrl[5] = n_all - small; }
    const int64_t n = (int64_t)(n_all - small) * per_query;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) hist2[i] = 0;
}

struct ExactPlanParams {
    const uint32_t* hist2;  // [f][s][d]
    const int* n_fail;
    const int* fail_list;
    int P, b;
    int64_t R;
    int* thr2;  // by query id
    uint32_t* bin_off2;
    uint32_t* bin_cap2;
    uint32_t* quota2;
};

__global__ void __launch_bounds__(kApWarps * 32) exact_plan_kernel(ExactPlanParams p)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = p.b + 1;
    unsigned long long* tot = reinterpret_cast<unsigned long long*>(smem) + (size_t)warp * nb;
    const int64_t f = (int64_t)blockIdx.x * kApWarps + warp;
    if (f >= (int64_t)*p.n_fail) return;
    const int q = p.fail_list[f];
    const uint32_t* h2 = p.hist2 + f * p.P * nb;
    const uint32_t FULL = 0xffffffffu;
    for (int d = lane; d < nb; d += 32) {
        unsigned long long s = 0;
        for (int sp = 0; sp < p.P; ++sp) s += h2[(int64_t)sp * nb + d];
        tot[d] = s;
    }
    __syncwarp();
    int dstar = p.b;
    unsigned long long below = 0;
    if (lane == 0) {
        unsigned long long cum = 0;
        for (int d = 0; d < nb; ++d) {
            if (cum + tot[d] >= (unsigned long long)p.R) { dstar = d; below = cum; break; }
            cum += tot[d];
        }
        p.thr2[q] = dstar;
    }
    dstar = __shfl_sync(FULL, dstar, 0);
    below = __shfl_sync(FULL, below, 0);
    const unsigned long long quota = (unsigned long long)p.R - below;
    unsigned long long run_eq = 0, run_off = 0;
    for (int sb = 0; sb < p.P; sb += 32) {
        const int s = sb + lane;
        unsigned long long lt = 0, eq = 0;
        if (s < p.P) {
            const uint32_t* hs = h2 + (int64_t)s * nb;
            for (int d = 0; d < dstar; ++d) lt += hs[d];
            eq = hs[dstar];
        }
        // exclusive scans over the 32 splits of this trip
        unsigned long long eq_incl = eq;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(FULL, eq_incl, o);
            if (lane >= o) eq_incl += t;
        }
        const unsigned long long eq_before = run_eq + eq_incl - eq;
        unsigned long long take = 0;
        if (quota > eq_before) take = min(eq, quota - eq_before);
        const unsigned long long capb = lt + take;
        unsigned long long cap_incl = capb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(FULL, cap_incl, o);
            if (lane >= o) cap_incl += t;
        }
        if (s < p.P) {
            p.bin_off2[f * p.P + s] = (uint32_t)(run_off + cap_incl - capb);
            p.bin_cap2[f * p.P + s] = (uint32_t)capb;
            p.quota2[f * p.P + s] = (uint32_t)take;
        }
        run_eq += __shfl_sync(FULL, eq_incl, 31);
        run_off += __shfl_sync(FULL, cap_incl, 31);
    }
}

// ================================================================================================
// Launchers
// ================================================================================================
template <int W>
static int launch_hist(const HistParams& hp, int64_t n_slots_max, int n_chunks, cudaStream_t st)
{
    // 16-bit counters (half the shared memory, nine resident CTAs instead of five) were measured on B200: 0.146-0.166 ms for
    // the C4 sample pass against 0.138 ms with 32-bit result-less atomics -- the load/add/store chain costs more than the
    // occupancy buys.  HG_HIST_NARROW=1 selects them for experiments.
    const bool narrow = env_int("HG_HIST_NARROW", 0) != 0 && (int64_t)hp.seg_per_chunk * hp.seg_rows < 65536;
    const size_t smem = (narrow ? sizeof(uint16_t) : sizeof(uint32_t)) * (size_t)(hp.b + 1) * kHistThreads + sizeof(uint32_t) * (size_t)hp.seg_rows * hp.Wr;
    dim3 grid((unsigned)ceil_div(n_slots_max, kHistThreads), (unsigned)n_chunks);
    if (narrow) {
        static thread_local size_t configured = 0;
        if (smem > configured) {
            HG_CUDA_TRY(cudaFuncSetAttribute(hist_kernel<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        hist_kernel<W, true><<<grid, kHistThreads, smem, st>>>(hp);
    } else {
        static thread_local size_t configured = 0;
        if (smem > configured) {
            HG_CUDA_TRY(cudaFuncSetAttribute(hist_kernel<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        hist_kernel<W, false><<<grid, kHistThreads, smem, st>>>(hp);
    }
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

template <int W, bool EXACT, bool LW1>
static int launch_select_q(const SelectParams& sp, const Plan& pl, int n_splits, cudaStream_t st)
{
    dim3 grid((unsigned)pl.nqt, (unsigned)n_splits);
    // <= 16 KB of tiles (+ <= 8 KB of query label words in the generic-label variant): no opt-in needed
    const size_t smem = sizeof(uint32_t) * (2 * (size_t)pl.TILE * pl.Wr + (LW1 ? 0 : (size_t)pl.LW * pl.TQ));
    switch (pl.QT) {
        case 4: select_kernel<W, 4, EXACT, LW1><<<grid, kSelectThreads, smem, st>>>(sp); break;
        case 2: select_kernel<W, 2, EXACT, LW1><<<grid, kSelectThreads, smem, st>>>(sp); break;
        default: select_kernel<W, 1, EXACT, LW1><<<grid, kSelectThreads, smem, st>>>(sp); break;
    }
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

template <int W, bool EXACT>
static int launch_select_w(const SelectParams& sp, const Plan& pl, int n_splits, cudaStream_t st)
{
    return pl.LW == 1 ? launch_select_q<W, EXACT, true>(sp, pl, n_splits, st) : launch_select_q<W, EXACT, false>(sp, pl, n_splits, st);
}

template <bool WINDOW>
static int launch_ap(ApParams ap, int64_t n_slots_max, cudaStream_t st)
{
    const int ncol = WINDOW ? 32 : ap.b + 1;
    // one shared-memory column of 2*ncol counters per thread (16-bit in the window variant)
    const size_t csize = WINDOW ? sizeof(uint16_t) : sizeof(uint32_t);
    int threads = 128;
    while (threads > 32 && (size_t)threads * 2 * ncol * csize > 200 * 1024) threads >>= 1;
    const size_t smem = (size_t)threads * 2 * ncol * csize;
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        HG_CUDA_TRY(cudaFuncSetAttribute(ap_kernel<WINDOW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    // threads per query: enough threads to fill the GPU (resident CTAs limited by shared memory), power of two <= 32, <= P
    const int sms = device_facts().sm_count > 0 ? device_facts().sm_count : 148;
    const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (220 * 1024) / (smem + 1024)));
    const int64_t want_threads = (int64_t)sms * ctas_per_sm * threads * 2;
    int G = 1;
    while (G < 32 && G * 2 <= ap.P && (int64_t)n_slots_max * G < want_threads) G <<= 1;
    G = env_int("HG_AP_G", G);
    if (G < 1 || G > 32 || (G & (G - 1)) || G > ap.P) G = 1;
    ap.G = G;
    {
        // whole bins: a lane owns floor or ceil(P / G) of them; halves pay off when that leaves the lanes unevenly loaded (C4: 44 bins
        // over 32 lanes = 69 % -> 88 halves = 92 %; C5: 88 bins = 92 % either way, where the halves only add loop overhead)
        const double whole = (double)ap.P / ((double)G * (double)ceil_div(ap.P, G));
        const double half = (double)(2 * ap.P) / ((double)G * (double)ceil_div(2 * ap.P, G));
        ap.halves = half > whole + 0.1 ? 1 : 0;
    }
    const int64_t blocks = ceil_div(n_slots_max * G, threads);
    ap_kernel<WINDOW><<<(unsigned)blocks, threads, smem, st>>>(ap);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

template <int W>
static int run_map(const Plan& pl, const uint32_t* q_rows, const uint32_t* db_rows, unsigned flags, double* d_ap, uint32_t* d_ids, uint16_t* d_dist, int32_t* d_rel, char* ws, cudaStream_t st,
                   const MapChunks* chunks = nullptr, PrepareRowsFn prepare = nullptr, void* user = nullptr)
{
    // Chunked mode (host pipeline): the database rows of chunk k become valid when prepare(k) has been enqueued
    // on `st`; thresholds are estimated from the caller's sample block (segments spread over the whole database, copied
    // ahead of chunk 0), select runs chunk by chunk behind the copies.
    int K = (chunks && prepare) ? chunks->K : 1;
    int64_t sample_stride = pl.seg_stride, sample_nseg = pl.n_seg, sample_rows = pl.sample_rows;
    int sample_spc = pl.seg_per_chunk, sample_chunks = pl.n_chunks;
    int prepared = 0;
    auto prepare_upto = [&](int k_end) -> int {
        for (; prepared < k_end; ++prepared) {
            int prc = prepare(user, prepared, chunks->row_lo[prepared], chunks->row_hi[prepared], st);
            if (prc != HG_OK) return prc;
        }
        return HG_OK;
    };
    const uint32_t* hist_rows = db_rows;   // rows the threshold sample is read from
    int64_t hist_ndb = pl.ndb;
    if (K > 1) {
        if (pl.sample_rows >= pl.ndb || (flags & HG_FLAG_FORCE_EXACT)) {
            int prc = prepare_upto(K);  // the estimate needs the whole database: no overlap possible
            if (prc != HG_OK) return prc;
            K = 1;
        } else if (chunks->sample_packed) {
            // the caller gathered the plan's sample segments (spread over the WHOLE database, lib/dataloader.py:93-94 leaves the
            // row order to the caller: a class-sorted database must not bias the thresholds) into one contiguous block
            hist_rows = chunks->sample_packed;
            hist_ndb = chunks->sample_n_seg * chunks->sample_seg_rows;
            sample_nseg = chunks->sample_n_seg;
            sample_stride = chunks->sample_seg_rows;  // contiguous segments
            sample_rows = hist_ndb;
            int prc = prepare_upto(1);
            if (prc != HG_OK) return prc;
        } else {
            // no sample block (HG_HOST_SAMPLE=chunk0): estimate from chunk 0 alone -- fine for a shuffled database, biased for a sorted one
            const int64_t avail_tiles = (chunks->row_hi[0] - chunks->row_lo[0]) / pl.TILE;  // chunk 0 = whole splits = whole tiles
            sample_nseg = std::min<int64_t>(pl.n_seg, std::max<int64_t>(1, avail_tiles));
            sample_stride = std::max<int64_t>(1, avail_tiles / sample_nseg) * pl.TILE;
            sample_rows = sample_nseg * pl.TILE;
            sample_spc = (int)ceil_div(sample_nseg, std::max<int64_t>(1, std::min<int64_t>(sample_nseg, pl.n_chunks)));
            sample_chunks = (int)ceil_div(sample_nseg, sample_spc);
            hist_ndb = chunks->row_hi[0];
            int prc = prepare_upto(1);
            if (prc != HG_OK) return prc;
        }
    } else if (chunks && prepare) {
        int prc = prepare_upto(chunks->K);
        if (prc != HG_OK) return prc;
    }

    int* ctrl = reinterpret_cast<int*>(ws + pl.off_ctrl);
    int* thr = reinterpret_cast<int*>(ws + pl.off_thr);
    int* thr2 = reinterpret_cast<int*>(ws + pl.off_thr2);
    int* fail_list = reinterpret_cast<int*>(ws + pl.off_fail);
    int* wide_list = reinterpret_cast<int*>(ws + pl.off_wide);
    uint32_t* hist_s = reinterpret_cast<uint32_t*>(ws + pl.off_hist_s);
    uint32_t* bin_cnt = reinterpret_cast<uint32_t*>(ws + pl.off_bin_cnt);
    uint32_t* bin_cnt0 = reinterpret_cast<uint32_t*>(ws + pl.off_bin_cnt0);
    uint32_t* bin_cnt2 = reinterpret_cast<uint32_t*>(ws + pl.off_bin_cnt2);
    uint32_t* bin_off2 = reinterpret_cast<uint32_t*>(ws + pl.off_bin_off2);
    uint32_t* bin_cap2 = reinterpret_cast<uint32_t*>(ws + pl.off_bin_cap2);
    uint32_t* quota2 = reinterpret_cast<uint32_t*>(ws + pl.off_quota2);
    uint32_t* hist2 = reinterpret_cast<uint32_t*>(ws + pl.off_hist2);
    uint32_t* lists = reinterpret_cast<uint32_t*>(ws + pl.off_lists);
    int* n_fail = ctrl;
    const bool force_exact = (flags & HG_FLAG_FORCE_EXACT) != 0;
    const bool no_fallback = (flags & HG_FLAG_NO_FALLBACK) != 0;
    const int nb = pl.b + 1;

    PhaseTimer& timer = phase_timer();
    timer.armed = false;
    if (flags & HG_FLAG_TIMING) {
        int trc = timer.ensure();
        if (trc != HG_OK) return trc;
        timer.armed = true;
    }
    HG_CUDA_TRY(cudaMemsetAsync(ctrl, 0, 256, st));
    int rc;
    timer.mark(kPhaseSample, st);
    if (pl.dense && !force_exact) {
        if (chunks && prepare && (rc = prepare_upto(chunks->K)) != HG_OK) return rc;  // the walk needs the whole database
        timer.mark(kPhaseThreshold, st);
        timer.mark(kPhaseExpand, st);
        timer.mark(kPhaseSelect, st);
        timer.mark(kPhaseAp, st);
        DenseApParams dp{};
        dp.q_rows = q_rows; dp.db_rows = db_rows; dp.nq = pl.nq; dp.ndb = pl.ndb; dp.R = pl.R; dp.b = pl.b; dp.LW = pl.LW; dp.Wr = pl.Wr;
        dp.ap = d_ap; dp.ids = d_ids; dp.dist = d_dist; dp.rel = d_rel;
        if ((rc = launch_dense_ap<W>(dp, st)) != HG_OK) return rc;
        timer.mark(kPhaseExact, st);
        timer.mark(kNumPhases, st);
        return HG_OK;
    }
    if (!force_exact) {
        // 1. sampled histogram
        HG_CUDA_TRY(cudaMemsetAsync(hist_s, 0, sizeof(uint32_t) * (size_t)pl.nq * nb, st));
        HistParams hp{};
        hp.q_rows = q_rows; hp.db_rows = hist_rows; hp.nq = pl.nq; hp.ndb = hist_ndb; hp.b = pl.b; hp.Wr = pl.Wr;
        hp.n_active = nullptr; hp.qlist = nullptr;
        hp.seg_stride = sample_stride; hp.n_seg = sample_nseg; hp.seg_rows = pl.TILE; hp.seg_per_chunk = sample_spc;
        hp.out = hist_s; hp.out_chunks = 1;
        if ((rc = launch_hist<W>(hp, pl.nq, sample_chunks, st)) != HG_OK) return rc;
    }
    // 2. thresholds
    timer.mark(kPhaseThreshold, st);
    thr_kernel<<<(unsigned)ceil_div(pl.nq * 32, 256), 256, 0, st>>>(hist_s, pl.nq, pl.b, sample_rows, pl.ndb, pl.R, kSampleZ,
                                                                    force_exact ? 1 : 0, thr);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    timer.mark(kPhaseExpand, st);
    bool select_marked = false;
    if (!force_exact) {
        // 3. single-pass select
        SelectParams sp{};
        sp.q_rows = q_rows; sp.db_rows = db_rows; sp.nq = pl.nq; sp.ndb = pl.ndb; sp.Wr = pl.Wr; sp.LW = pl.LW; sp.TILE = pl.TILE; sp.thr = thr;
        sp.n_active = nullptr; sp.qlist = nullptr; sp.P = pl.P; sp.SL = pl.SL; sp.R = pl.R;
        sp.lists = lists; sp.cap = pl.cap; sp.bin_off2 = nullptr; sp.bin_cap2 = nullptr; sp.quota2 = nullptr;
        sp.bin_cnt = bin_cnt; sp.bin_cnt0 = bin_cnt0;
        UmmaSelectArgs ua{};
        uint8_t* q8 = reinterpret_cast<uint8_t*>(ws + pl.off_q8);
        uint8_t* db8 = reinterpret_cast<uint8_t*>(ws + pl.off_db8);
        if (pl.umma_kp) {
            ua.q_rows = q_rows; ua.db_rows = db_rows; ua.nq = pl.nq; ua.ndb = pl.ndb; ua.b = pl.b; ua.W = pl.W; ua.LW = pl.LW; ua.Wr = pl.Wr;
            ua.KP = pl.umma_kp; ua.thr = thr; ua.P = pl.P; ua.SL = pl.SL; ua.lists = lists; ua.cap = pl.cap; ua.bin_cnt = bin_cnt; ua.bin_cnt0 = bin_cnt0;
            ua.q8 = q8; ua.db8 = db8;
            ua.queued = pl.queued ? 1 : 0;
            ua.qx = reinterpret_cast<uint8_t*>(ws + pl.off_qx); ua.bx = reinterpret_cast<uint8_t*>(ws + pl.off_bx);
            if ((rc = umma_expand_q(q_rows, pl.nq, pl.b, pl.Wr, pl.umma_kp, q8, st)) != HG_OK) return rc;
            if ((rc = umma_thr_columns(thr, pl.nq, pl.b, const_cast<uint8_t*>(ua.qx), const_cast<uint8_t*>(ua.bx), st)) != HG_OK) return rc;
        }
        if (K > 1 || !pl.umma_kp) { timer.mark(kPhaseSelect, st); select_marked = true; }  // chunked: expansion is interleaved with select
        // one launch per chunk of whole splits (a single chunk unless the host pipeline feeds the database piecewise)
        for (int k = 0; k < K; ++k) {
            int64_t lo = 0, hi = pl.ndb;
            if (K > 1) {
                if ((rc = prepare_upto(k + 1)) != HG_OK) return rc;
                lo = chunks->row_lo[k]; hi = chunks->row_hi[k];
            }
            const int split0 = (int)(lo / pl.SL), n_splits = (int)ceil_div(hi - lo, pl.SL);
            if (pl.umma_kp) {
                if ((rc = umma_expand_db(db_rows, lo, hi, pl.ndb, pl.b, pl.Wr, pl.umma_kp, db8, st)) != HG_OK) return rc;
                if (!select_marked) { timer.mark(kPhaseSelect, st); select_marked = true; }
                ua.split0 = split0; ua.n_splits = n_splits;
                if ((rc = umma_select_launch(ua, st)) != HG_OK) return rc;
            } else {
                sp.split0 = split0;
                if ((rc = launch_select_w<W, false>(sp, pl, n_splits, st)) != HG_OK) return rc;
            }
        }
    } else {
        HG_CUDA_TRY(cudaMemsetAsync(bin_cnt, 0, sizeof(uint32_t) * (size_t)pl.nq * pl.P, st));
        HG_CUDA_TRY(cudaMemsetAsync(bin_cnt0, 0, sizeof(uint32_t) * (size_t)pl.nq * pl.P, st));
    }
    if (!select_marked) timer.mark(kPhaseSelect, st);
    // 4. AP (queries that cannot be answered exactly from their candidates go to the fail list)
    timer.mark(kPhaseAp, st);
    {
        ApParams ap{};
        ap.nq = pl.nq; ap.n_active = nullptr; ap.qlist = nullptr; ap.bins_by_slot = 0; ap.P = pl.P; ap.b = pl.b; ap.SL = pl.SL; ap.R = pl.R;
        ap.lists = lists; ap.cap = pl.cap; ap.bin_off2 = nullptr; ap.bin_cap2 = nullptr; ap.bin_cnt = bin_cnt; ap.bin_cnt0 = bin_cnt0; ap.thr = thr;
        ap.ap = d_ap; ap.ids = d_ids; ap.dist = d_dist; ap.rel = d_rel;
        ap.fail_list = fail_list; ap.n_fail = n_fail; ap.no_fallback = no_fallback ? 1 : 0;
        ap.wide_list = wide_list; ap.n_wide = ctrl + 2;
        if ((rc = launch_ap<true>(ap, pl.nq, st)) != HG_OK) return rc;
        // queries whose candidate distances span >= 32 values: same bins, full-width counters
        ap.n_active = ctrl + 2; ap.qlist = wide_list; ap.wide_list = nullptr; ap.n_wide = nullptr;
        if ((rc = launch_ap<false>(ap, pl.nq, st)) != HG_OK) return rc;
    }
    timer.mark(kPhaseExact, st);
    if (no_fallback) { timer.mark(kNumPhases, st); return HG_OK; }
    // 5. exact path for the fail list (grids sized for nq, CTAs beyond the fail count exit immediately)
    {
        const int sms = device_facts().sm_count > 0 ? device_facts().sm_count : 148;
        // up to 512 failed queries: one dense walk each (zero_hist2_kernel publishes the routing); HG_EXACT_DENSE_MAX=0 keeps every list on the per-split path
        const int dense_max = (pl.ndb / 32 <= kMaxSplitRows) ? (int)std::min<int64_t>(pl.nq, env_int("HG_EXACT_DENSE_MAX", 512)) : 0;
        zero_hist2_kernel<<<sms * 4, 256, 0, st>>>(hist2, ctrl, dense_max, (int64_t)pl.P * nb);
        count_launch();
        HG_CUDA_TRY(cudaGetLastError());
        if (dense_max > 0) {
            DenseApParams dp{};
            dp.q_rows = q_rows; dp.db_rows = db_rows; dp.nq = dense_max; dp.ndb = pl.ndb; dp.R = pl.R; dp.b = pl.b; dp.LW = pl.LW; dp.Wr = pl.Wr;
            dp.ap = d_ap; dp.ids = d_ids; dp.dist = d_dist; dp.rel = d_rel;
            dp.n_active = ctrl + 4; dp.qlist = fail_list;
            if ((rc = launch_dense_ap<W>(dp, st)) != HG_OK) return rc;
        }
        n_fail = ctrl + 5;  // what is left for the per-split path
        HistParams hp{};
        hp.q_rows = q_rows; hp.db_rows = db_rows; hp.nq = pl.nq; hp.ndb = pl.ndb; hp.b = pl.b; hp.Wr = pl.Wr;
        hp.n_active = n_fail; hp.qlist = fail_list;
        hp.seg_stride = pl.TILE; hp.n_seg = ceil_div(pl.ndb, pl.TILE); hp.seg_rows = pl.TILE; hp.seg_per_chunk = (int)(pl.SL / pl.TILE);
        hp.out = hist2; hp.out_chunks = pl.P;
        if ((rc = launch_hist<W>(hp, pl.nq, pl.P, st)) != HG_OK) return rc;

        ExactPlanParams ep{};
        ep.hist2 = hist2; ep.n_fail = n_fail; ep.fail_list = fail_list; ep.P = pl.P; ep.b = pl.b; ep.R = pl.R;
        ep.thr2 = thr2; ep.bin_off2 = bin_off2; ep.bin_cap2 = bin_cap2; ep.quota2 = quota2;
        const size_t smem = sizeof(unsigned long long) * (size_t)nb * kApWarps;
        exact_plan_kernel<<<(unsigned)ceil_div(pl.nq, kApWarps), kApWarps * 32, smem, st>>>(ep);
        count_launch();
        HG_CUDA_TRY(cudaGetLastError());

        SelectParams sp{};
        sp.q_rows = q_rows; sp.db_rows = db_rows; sp.nq = pl.nq; sp.ndb = pl.ndb; sp.Wr = pl.Wr; sp.LW = pl.LW; sp.TILE = pl.TILE; sp.thr = thr2;
        sp.n_active = n_fail; sp.qlist = fail_list; sp.P = pl.P; sp.SL = pl.SL; sp.R = pl.R;
        sp.lists = lists; sp.cap = 0; sp.bin_off2 = bin_off2; sp.bin_cap2 = bin_cap2; sp.quota2 = quota2;
        sp.bin_cnt = bin_cnt2; sp.bin_cnt0 = nullptr; sp.split0 = 0;
        if ((rc = launch_select_w<W, true>(sp, pl, pl.P, st)) != HG_OK) return rc;

        ApParams ap{};
        ap.nq = pl.nq; ap.n_active = n_fail; ap.qlist = fail_list; ap.bins_by_slot = 1; ap.P = pl.P; ap.b = pl.b; ap.SL = pl.SL; ap.R = pl.R;
        ap.lists = lists; ap.cap = 0; ap.bin_off2 = bin_off2; ap.bin_cap2 = bin_cap2; ap.bin_cnt = bin_cnt2; ap.bin_cnt0 = nullptr; ap.thr = thr2;
        ap.ap = d_ap; ap.ids = d_ids; ap.dist = d_dist; ap.rel = d_rel;
        ap.fail_list = nullptr; ap.n_fail = ctrl + 1; ap.no_fallback = 0; ap.wide_list = nullptr; ap.n_wide = nullptr;
        if ((rc = launch_ap<false>(ap, pl.nq, st)) != HG_OK) return rc;
    }
    timer.mark(kNumPhases, st);
    return HG_OK;
}

// ---- internal entry points used by the pipelined host path (api.cu) ---------------------------------------
// The chunked path launches select once per chunk: it uses kChunkCtasMult x more (smaller) splits so that every
// launch still fills the SMs.
constexpr int kChunkCtasMult = 4;

int plan_chunks(int64_t nq, int64_t ndb, int b, int L, int64_t R, int k_req, MapChunks* out, size_t* ws_bytes)
{
    const Plan pl = make_plan(nq, ndb, b, L, R, kChunkCtasMult);
    if (!pl.ok || !out) return fail(HG_EINVAL, "plan_chunks: sizes out of range");
    if (ws_bytes) *ws_bytes = pl.total;
    int K = std::max(1, std::min(std::min(k_req, pl.P / 2), MapChunks::kMax));
    // Chunk k is ranked while chunk k + 1 is still on the PCIe bus, so what is exposed at the end is the work on the LAST
    // chunk: the last two chunks are short (0.6 and 0.25 of a regular one).  Whole splits per chunk, even boundaries
    // because the tensor-core select ranks splits in pairs.
    double wsum = 0.0, w[MapChunks::kMax];
    for (int k = 0; k < K; ++k) { w[k] = (K >= 4 && k == K - 1) ? 0.25 : ((K >= 4 && k == K - 2) ? 0.6 : 1.0); wsum += w[k]; }
    int64_t s_prev = 0;
    double acc = 0.0;
    int kk = 0;
    for (int k = 0; k < K; ++k) {
        acc += w[k];
        int64_t s_hi = k + 1 == K ? (int64_t)pl.P : ((int64_t)llround(acc / wsum * (double)pl.P) & ~int64_t(1));
        s_hi = std::min<int64_t>(std::max<int64_t>(s_hi, s_prev), pl.P);
        if (s_hi == s_prev) continue;  // rounding left nothing for this chunk
        out->row_lo[kk] = s_prev * pl.SL;
        out->row_hi[kk] = std::min<int64_t>(s_hi * pl.SL, ndb);
        s_prev = s_hi;
        ++kk;
    }
    out->K = std::max(1, kk);
    out->sample_n_seg = 0; out->sample_seg_stride = 0; out->sample_seg_rows = 0; out->sample_packed = nullptr;
    if (out->K > 1 && pl.sample_rows < pl.ndb) {  // every sampled segment is a full tile (make_plan)
        out->sample_n_seg = pl.n_seg;
        out->sample_seg_stride = pl.seg_stride;
        out->sample_seg_rows = pl.TILE;
    }
    return HG_OK;
}

int hamming_map_chunked(const uint32_t* q_rows, int64_t nq, const uint32_t* db_rows, int64_t ndb, int b, int L, int64_t R, unsigned flags,
                        double* d_ap, void* ws, size_t ws_bytes, cudaStream_t st, const MapChunks* chunks, PrepareRowsFn prepare, void* user)
{
    const Plan pl = make_plan(nq, ndb, b, L, R, kChunkCtasMult);
    if (!pl.ok) return fail(HG_EINVAL, "hamming_map_chunked: sizes out of range");
    if (ws_bytes < pl.total) return fail(HG_ENOMEM, "hamming_map_chunked: workspace %zu B < required %zu B", ws_bytes, pl.total);
    char* w = static_cast<char*>(ws);
    switch (pl.W) {
        case 1: return run_map<1>(pl, q_rows, db_rows, flags, d_ap, nullptr, nullptr, nullptr, w, st, chunks, prepare, user);
        case 2: return run_map<2>(pl, q_rows, db_rows, flags, d_ap, nullptr, nullptr, nullptr, w, st, chunks, prepare, user);
        case 3: return run_map<3>(pl, q_rows, db_rows, flags, d_ap, nullptr, nullptr, nullptr, w, st, chunks, prepare, user);
        case 4: return run_map<4>(pl, q_rows, db_rows, flags, d_ap, nullptr, nullptr, nullptr, w, st, chunks, prepare, user);
        case 8: return run_map<8>(pl, q_rows, db_rows, flags, d_ap, nullptr, nullptr, nullptr, w, st, chunks, prepare, user);
        default: return fail(HG_EINVAL, "hamming_map_chunked: unsupported word count %d", pl.W);
    }
}

}  // namespace hg

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" size_t hg_hamming_map_workspace_bytes(int64_t nq, int64_t ndb, int b, int L, int64_t R)
{
    const hg::Plan pl = hg::make_plan(nq, ndb, b, L, R);
    return pl.ok ? pl.total : 0;
}

extern "C" int hg_hamming_map(const uint32_t* d_q_rows, int64_t nq, const uint32_t* d_db_rows, int64_t ndb, int b, int L, int64_t R,
                              unsigned flags, double* d_ap, uint32_t* d_ids, uint16_t* d_dist, int32_t* d_rel, void* d_workspace,
                              size_t workspace_bytes, void* stream)
{
    if (nq == 0) return HG_OK;
    if (nq < 0 || ndb < 0) return hg::fail(HG_EINVAL, "hg_hamming_map: negative size");
    if (R <= 0) return hg::fail(HG_EINVAL, "hg_hamming_map: R must be positive (got %lld)", (long long)R);
    if (R > ndb) return hg::fail(HG_ERANGE, "hg_hamming_map: R=%lld exceeds the database size %lld", (long long)R, (long long)ndb);
    if (hg_code_words(b) == 0) return hg::fail(HG_EINVAL, "hg_hamming_map: unsupported hash length b=%d (1..%d)", b, HG_MAX_BITS);
    if (hg_label_words(L) == 0) return hg::fail(HG_EINVAL, "hg_hamming_map: unsupported label width L=%d (1..%d)", L, HG_MAX_LABELS);
    if (!d_q_rows || !d_db_rows || !d_ap || !d_workspace) return hg::fail(HG_EINVAL, "hg_hamming_map: NULL pointer");
    if ((reinterpret_cast<uintptr_t>(d_db_rows) & 15) || (reinterpret_cast<uintptr_t>(d_workspace) & 255))
        return hg::fail(HG_EINVAL, "hg_hamming_map: d_db_rows must be 16-byte and the workspace 256-byte aligned");
    if (!hg::device_facts().ok) return hg::fail(HG_ECUDA, "hg_hamming_map: no CUDA device");
    const hg::Plan pl = hg::make_plan(nq, ndb, b, L, R);
    if (!pl.ok) return hg::fail(HG_EINVAL, "hg_hamming_map: sizes out of range (nq=%lld ndb=%lld)", (long long)nq, (long long)ndb);
    if (workspace_bytes < pl.total)
        return hg::fail(HG_ENOMEM, "hg_hamming_map: workspace %zu B < required %zu B", workspace_bytes, pl.total);
    char* ws = static_cast<char*>(d_workspace);
    cudaStream_t st = (cudaStream_t)stream;
    switch (pl.W) {
        case 1: return hg::run_map<1>(pl, d_q_rows, d_db_rows, flags, d_ap, d_ids, d_dist, d_rel, ws, st);
        case 2: return hg::run_map<2>(pl, d_q_rows, d_db_rows, flags, d_ap, d_ids, d_dist, d_rel, ws, st);
        case 3: return hg::run_map<3>(pl, d_q_rows, d_db_rows, flags, d_ap, d_ids, d_dist, d_rel, ws, st);
        case 4: return hg::run_map<4>(pl, d_q_rows, d_db_rows, flags, d_ap, d_ids, d_dist, d_rel, ws, st);
        case 8: return hg::run_map<8>(pl, d_q_rows, d_db_rows, flags, d_ap, d_ids, d_dist, d_rel, ws, st);
        default: return hg::fail(HG_EINVAL, "hg_hamming_map: unsupported word count %d", pl.W);
    }
}

extern "C" int hg_select_backend(int b, int L)
{
    const int Wr = hg_row_words(b, L);
    if (Wr == 0) return -1;
    const char* be = getenv("HG_SELECT_BACKEND");
    if (be && be[0] == 'p') return 0;
    return hg::umma_select_kp(b, Wr);
}

extern "C" int hg_select_backend_for(int64_t nq, int64_t ndb, int b, int L, int64_t R)
{
    const hg::Plan pl = hg::make_plan(nq, ndb, b, L, R);
    if (!pl.ok) return -1;
    return pl.dense ? 1 : pl.umma_kp;
}

extern "C" int hg_select_queued_for(int64_t nq, int64_t ndb, int b, int L, int64_t R)
{
    const hg::Plan pl = hg::make_plan(nq, ndb, b, L, R);
    if (!pl.ok) return -1;
    return (!pl.dense && pl.umma_kp && pl.queued) ? 1 : 0;
}

extern "C" int hg_hamming_map_stats(const void* d_workspace, size_t workspace_bytes, int64_t nq, int64_t ndb, int b, int L, int64_t R,
                                    int64_t out[8], void* stream)
{
    const hg::Plan pl = hg::make_plan(nq, ndb, b, L, R);
    if (!pl.ok || !d_workspace || workspace_bytes < pl.total || !out) return hg::fail(HG_EINVAL, "hg_hamming_map_stats: bad arguments");
    int ctrl[4] = {0, 0, 0, 0};
    HG_CUDA_TRY(cudaMemcpyAsync(ctrl, static_cast<const char*>(d_workspace) + pl.off_ctrl, sizeof(ctrl), cudaMemcpyDeviceToHost,
                                (cudaStream_t)stream));
    HG_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    out[0] = ctrl[0];
    out[1] = pl.P;
    out[2] = pl.SL;
    out[3] = pl.cap;
    out[4] = pl.TQ;
    out[5] = pl.sample_rows;
    out[6] = ctrl[1];
    out[7] = ctrl[2];
    return HG_OK;
}
