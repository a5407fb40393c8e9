// Misc C-ABI entry points: version / errors / device facts, lib/metric.py:24 (mean over kept queries),
// the host-buffer end-to-end call (== MAPs(R).get_maps_by_feature, main.py:164) and the POPC-pipe
// microbenchmark that gives the Hamming kernel its roofline denominator.
#include "common.cuh"

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace hg {

char* last_error_buf()
{
    static thread_local char buf[512] = {0};
    return buf;
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int PhaseTimer::ensure()
{
    if (!created) {
        for (int i = 0; i <= kNumPhases; ++i) HG_CUDA_TRY(cudaEventCreate(&ev[i]));
        created = true;
    }
    return HG_OK;
}
PhaseTimer& phase_timer()
{
    static thread_local PhaseTimer t;
    return t;
}

const DeviceFacts& device_facts()
{
    static DeviceFacts facts[64];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        static DeviceFacts none;
        (void)cudaGetLastError();
        return none;
    }
    std::lock_guard<std::mutex> lock(mu);
    DeviceFacts& f = facts[dev];
    if (!f.ok) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) f.sm_count = v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess) f.cc_major = v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev) == cudaSuccess) f.cc_minor = v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev) == cudaSuccess) f.l2_bytes = (size_t)v;
        f.ok = f.sm_count > 0;
        (void)cudaGetLastError();
    }
    return f;
}

// NumPy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE), so that the
// mean over the kept queries is bit-identical to the reference's np.mean (lib/metric.py:24).
static double pairwise_sum(const double* a, int64_t n)
{
    if (n < 8) {
        double r = 0.0;  // numpy starts from -0.0; identical for our non-negative inputs except n == 0
        for (int64_t i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
}

// -------------------------------------------------------------------------------------------------
// POPC-pipe microbenchmark
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) popc_peak_kernel(uint32_t* __restrict__ out, int iters, uint32_t seed)
{
    constexpr int K = 8;
    uint32_t x[K], acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        x[k] = seed * (2654435761u * (threadIdx.x + 1 + k * 977u)) + blockIdx.x * 40503u;
        acc[k] = 0;
    }
    uint32_t y = seed ^ (threadIdx.x * 2246822519u);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += __popc(x[k] ^ y);
        y = y * 1664525u + 1013904223u;
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += acc[k];
    if (s == 0xdeadbeefu) out[0] = s;  // keeps the chain alive without a store on the timed path
}

// -------------------------------------------------------------------------------------------------
// Cached device arena for the host-buffer entry point
// -------------------------------------------------------------------------------------------------
struct Arena {
    void* ptr = nullptr;
    size_t bytes = 0;
    double* pinned_ap = nullptr;
    size_t pinned_n = 0;
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t copy_stream = nullptr;  // H2D of the database chunks, overlapped with packing + ranking
    cudaEvent_t chunk_ev[MapChunks::kMax] = {};
    cudaEvent_t sample_ev = nullptr;     // the threshold sample block has landed
    bool events = false;
    int device = -1;
};
static Arena g_arena;
static std::mutex g_arena_mu;

static int arena_reserve(Arena& a, size_t bytes, size_t n_ap)
{
    int dev = 0;
    HG_CUDA_TRY(cudaGetDevice(&dev));
    if (a.device != dev) {
        if (a.ptr) cudaFree(a.ptr);
        if (a.pinned_ap) cudaFreeHost(a.pinned_ap);
        if (a.stream) cudaStreamDestroy(a.stream);
        if (a.copy_stream) cudaStreamDestroy(a.copy_stream);
        if (a.events) { for (auto& e : a.chunk_ev) cudaEventDestroy(e); cudaEventDestroy(a.sample_ev); }
        a = Arena();
        a.device = dev;
    }
    if (!a.stream) HG_CUDA_TRY(cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking));
    if (!a.copy_stream) HG_CUDA_TRY(cudaStreamCreateWithFlags(&a.copy_stream, cudaStreamNonBlocking));
    if (!a.events) {
        for (auto& e : a.chunk_ev) HG_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        HG_CUDA_TRY(cudaEventCreateWithFlags(&a.sample_ev, cudaEventDisableTiming));
        a.events = true;
    }
    if (bytes > a.bytes) {
        if (a.ptr) HG_CUDA_TRY(cudaFree(a.ptr));
        a.ptr = nullptr; a.bytes = 0;
        const size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&a.ptr, want);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            return fail(HG_ENOMEM, "hg_maps_by_feature_host: cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        }
        a.bytes = want;
    }
    if (n_ap > a.pinned_n) {
        if (a.pinned_ap) HG_CUDA_TRY(cudaFreeHost(a.pinned_ap));
        a.pinned_ap = nullptr; a.pinned_n = 0;
        HG_CUDA_TRY(cudaMallocHost(&a.pinned_ap, sizeof(double) * n_ap));
        a.pinned_n = n_ap;
    }
    return HG_OK;
}

}  // namespace hg

extern "C" int hg_version(void) { return 100; }  // 0.1.0

// CRC-32C (Castagnoli) of a host buffer, continuing from `crc`: the checksum TensorFlow checkpoints carry per table
// block and per tensor (hashgan_b200/tf_checkpoint.py verifies them on read).  Plain host code, slicing-by-4 tables.
extern "C" uint32_t hg_crc32c(const void* data, size_t n, uint32_t crc)
{
    static uint32_t tab[4][256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
            tab[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int t = 1; t < 4; ++t) tab[t][i] = (tab[t - 1][i] >> 8) ^ tab[0][tab[t - 1][i] & 0xFFu];
        ready = true;
    }
    const unsigned char* p = static_cast<const unsigned char*>(data);
    uint32_t c = crc ^ 0xFFFFFFFFu;
    while (n >= 4) {
        c ^= (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        c = tab[3][c & 0xFFu] ^ tab[2][(c >> 8) & 0xFFu] ^ tab[1][(c >> 16) & 0xFFu] ^ tab[0][c >> 24];
        p += 4; n -= 4;
    }
    while (n--) c = tab[0][(c ^ *p++) & 0xFFu] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

extern "C" int64_t hg_launch_count(int reset)
{
    const long long v = hg::g_launches.load(std::memory_order_relaxed);
    if (reset) hg::g_launches.store(0, std::memory_order_relaxed);
    return (int64_t)v;
}

extern "C" int hg_hamming_map_phase_ms(float out[6])
{
    hg::PhaseTimer& t = hg::phase_timer();
    if (!out || !t.created || !t.armed) return hg::fail(HG_EINVAL, "hg_hamming_map_phase_ms: no timed hg_hamming_map call on this thread");
    HG_CUDA_TRY(cudaEventSynchronize(t.ev[hg::kNumPhases]));
    for (int i = 0; i < hg::kNumPhases; ++i) HG_CUDA_TRY(cudaEventElapsedTime(&out[i], t.ev[i], t.ev[i + 1]));
    return HG_OK;
}

extern "C" const char* hg_last_error(void) { return hg::last_error_buf(); }

extern "C" int hg_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes)
{
    const hg::DeviceFacts& f = hg::device_facts();
    if (!f.ok) return hg::fail(HG_ECUDA, "hg_device_info: no usable CUDA device");
    if (sm_count) *sm_count = f.sm_count;
    if (cc_major) *cc_major = f.cc_major;
    if (cc_minor) *cc_minor = f.cc_minor;
    if (l2_bytes) *l2_bytes = f.l2_bytes;
    return HG_OK;
}

// Host-only half of hg_mean_ap, exported so the CPU test-suite can pin it against np.mean.
extern "C" int hg_mean_ap_host(const double* h_ap, int64_t nq, double* map_out, int64_t* n_used)
{
    if (nq < 0 || (nq > 0 && !h_ap) || !map_out) return hg::fail(HG_EINVAL, "hg_mean_ap_host: bad arguments");
    std::vector<double> kept;
    kept.reserve((size_t)nq);
    for (int64_t i = 0; i < nq; ++i)
        if (!std::isnan(h_ap[i])) kept.push_back(h_ap[i]);
    if (n_used) *n_used = (int64_t)kept.size();
    *map_out = kept.empty() ? std::nan("") : hg::pairwise_sum(kept.data(), (int64_t)kept.size()) / (double)kept.size();
    return HG_OK;
}

extern "C" int hg_mean_ap(const double* d_ap, int64_t nq, double* map_out, int64_t* n_used, void* stream)
{
    if (nq < 0 || (nq > 0 && !d_ap) || !map_out) return hg::fail(HG_EINVAL, "hg_mean_ap: bad arguments");
    std::vector<double> host((size_t)nq);
    if (nq > 0) {
        HG_CUDA_TRY(cudaMemcpyAsync(host.data(), d_ap, sizeof(double) * (size_t)nq, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        HG_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    }
    return hg_mean_ap_host(host.data(), nq, map_out, n_used);
}

extern "C" int hg_popc_peak(double* wordops_per_s, double* ms_out, int iters, void* stream)
{
    if (!wordops_per_s || iters <= 0) return hg::fail(HG_EINVAL, "hg_popc_peak: bad arguments");
    const hg::DeviceFacts& f = hg::device_facts();
    if (!f.ok) return hg::fail(HG_ECUDA, "hg_popc_peak: no CUDA device");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* d_out = nullptr;
    HG_CUDA_TRY(cudaMalloc(&d_out, 256));
    cudaEvent_t e0, e1;
    HG_CUDA_TRY(cudaEventCreate(&e0));
    HG_CUDA_TRY(cudaEventCreate(&e1));
    const int blocks = f.sm_count * 8, threads = 256;
    hg::count_launch(2);
    hg::popc_peak_kernel<<<blocks, threads, 0, st>>>(d_out, iters / 8 + 1, 12345u);  // warm-up
    HG_CUDA_TRY(cudaEventRecord(e0, st));
    hg::popc_peak_kernel<<<blocks, threads, 0, st>>>(d_out, iters, 98765u);
    HG_CUDA_TRY(cudaEventRecord(e1, st));
    HG_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    HG_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    HG_CUDA_TRY(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    const double ops = (double)blocks * threads * (double)iters * 8.0;
    *wordops_per_s = ops / ((double)ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return HG_OK;
}

extern "C" int hg_release_cached(void)
{
    std::lock_guard<std::mutex> lock(hg::g_arena_mu);
    hg::Arena& a = hg::g_arena;
    if (a.ptr) cudaFree(a.ptr);
    if (a.pinned_ap) cudaFreeHost(a.pinned_ap);
    if (a.stream) cudaStreamDestroy(a.stream);
    if (a.copy_stream) cudaStreamDestroy(a.copy_stream);
    if (a.events) { for (auto& e : a.chunk_ev) cudaEventDestroy(e); cudaEventDestroy(a.sample_ev); }
    a = hg::Arena();
    (void)cudaGetLastError();
    return HG_OK;
}

extern "C" int hg_maps_by_feature_host(const float* h_db_feat, const void* h_db_lab, int64_t ndb, const float* h_q_feat,
                                       const void* h_q_lab, int64_t nq, int b, int L, int lab_elem_bytes, int64_t R, unsigned flags,
                                       double* map_out, double* h_ap_out)
{
    if (!map_out) return hg::fail(HG_EINVAL, "hg_maps_by_feature_host: map_out is NULL");
    if (nq < 0 || ndb < 0) return hg::fail(HG_EINVAL, "hg_maps_by_feature_host: negative size");
    if (R <= 0) return hg::fail(HG_EINVAL, "hg_maps_by_feature_host: R must be positive");
    if (R > ndb) return hg::fail(HG_ERANGE, "hg_maps_by_feature_host: R=%lld exceeds the database size %lld", (long long)R, (long long)ndb);
    if (nq == 0) { *map_out = std::nan(""); return HG_OK; }
    const int W = hg_code_words(b), LW = hg_label_words(L), Wr = hg_row_words(b, L);
    if (W == 0) return hg::fail(HG_EINVAL, "hg_maps_by_feature_host: unsupported hash length b=%d", b);
    if (LW == 0) return hg::fail(HG_EINVAL, "hg_maps_by_feature_host: unsupported label width L=%d", L);
    if (lab_elem_bytes != 8 && lab_elem_bytes != 4 && lab_elem_bytes != 1)
        return hg::fail(HG_EINVAL, "hg_maps_by_feature_host: lab_elem_bytes must be 8, 4 or 1");
    if (!h_db_feat || !h_db_lab || !h_q_feat || !h_q_lab) return hg::fail(HG_EINVAL, "hg_maps_by_feature_host: NULL pointer");
    hg::MapChunks chunks;
    size_t ws_bytes = 0;
    const char* kenv = getenv("HG_HOST_CHUNKS");
    if (hg::plan_chunks(nq, ndb, b, L, R, (kenv && *kenv) ? atoi(kenv) : 8, &chunks, &ws_bytes) != HG_OK || ws_bytes == 0)
        return hg::fail(HG_EINVAL, "hg_maps_by_feature_host: sizes out of range (split the query batch)");

    std::lock_guard<std::mutex> lock(hg::g_arena_mu);
    hg::Arena& a = hg::g_arena;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_dbf = take(sizeof(float) * (size_t)ndb * b);
    const size_t o_qf = take(sizeof(float) * (size_t)nq * b);
    const size_t o_dbl = take((size_t)lab_elem_bytes * ndb * L);
    const size_t o_ql = take((size_t)lab_elem_bytes * nq * L);
    const size_t o_dbr = take(sizeof(uint32_t) * (size_t)ndb * Wr);
    const size_t o_qr = take(sizeof(uint32_t) * (size_t)nq * Wr);
    const size_t o_ap = take(sizeof(double) * (size_t)nq);
    const size_t o_bad = take(256);
    {
        const char* senv = getenv("HG_HOST_SAMPLE");  // "chunk0": thresholds from the first chunk only (diagnostics)
        if (senv && senv[0] == 'c') chunks.sample_n_seg = 0;
    }
    const int64_t n_sample = chunks.sample_n_seg * chunks.sample_seg_rows;
    const size_t o_sf = take(sizeof(float) * (size_t)n_sample * b);
    const size_t o_sr = take(sizeof(uint32_t) * (size_t)n_sample * Wr);
    const size_t o_ws = take(ws_bytes);
    int rc = hg::arena_reserve(a, off, (size_t)nq + 8);
    if (rc != HG_OK) return rc;
    char* base = static_cast<char*>(a.ptr);
    cudaStream_t st = a.stream;
    float* d_dbf = reinterpret_cast<float*>(base + o_dbf);
    float* d_qf = reinterpret_cast<float*>(base + o_qf);
    uint32_t* d_dbr = reinterpret_cast<uint32_t*>(base + o_dbr);
    uint32_t* d_qr = reinterpret_cast<uint32_t*>(base + o_qr);
    double* d_ap = reinterpret_cast<double*>(base + o_ap);
    int* d_bad = reinterpret_cast<int*>(base + o_bad);
    // an error return must not leave copies from the caller's buffers in flight (they may be freed right after the call)
    struct Drain {
        hg::Arena& a;
        bool armed;
        ~Drain() { if (armed) { cudaStreamSynchronize(a.copy_stream); cudaStreamSynchronize(a.stream); (void)cudaGetLastError(); } }
    } drain{a, true};

    HG_CUDA_TRY(cudaMemsetAsync(d_bad, 0, 256, st));
    HG_CUDA_TRY(cudaMemcpyAsync(d_qf, h_q_feat, sizeof(float) * (size_t)nq * b, cudaMemcpyHostToDevice, st));
    HG_CUDA_TRY(cudaMemcpyAsync(base + o_ql, h_q_lab, (size_t)lab_elem_bytes * nq * L, cudaMemcpyHostToDevice, st));
    if ((rc = hg_pack_rows(d_qf, b, base + o_ql, lab_elem_bytes, nq, b, L, d_qr, d_bad, st)) != HG_OK) return rc;

    // Database: chunked H2D on the copy stream, each chunk packed and ranked on the compute stream as soon as it
    // has landed, so the PCIe transfer (339 MB at C4) overlaps the select kernel instead of preceding it.
    struct Pipe {
        hg::Arena* a; hg::MapChunks ch; bool issued[hg::MapChunks::kMax]; bool pinned;
        const float* h_feat; const char* h_lab; float* d_feat; char* d_lab; uint32_t* d_rows; int* d_bad;
        int b, L, Wr, lab_bytes;
        int issue(int k) {
            if (issued[k]) return HG_OK;
            const int64_t lo = ch.row_lo[k], n = ch.row_hi[k] - lo;
            HG_CUDA_TRY(cudaMemcpyAsync(d_feat + lo * b, h_feat + lo * b, sizeof(float) * (size_t)n * b, cudaMemcpyHostToDevice, a->copy_stream));
            HG_CUDA_TRY(cudaMemcpyAsync(d_lab + (size_t)lo * L * lab_bytes, h_lab + (size_t)lo * L * lab_bytes, (size_t)lab_bytes * n * L,
                                        cudaMemcpyHostToDevice, a->copy_stream));
            HG_CUDA_TRY(cudaEventRecord(a->chunk_ev[k], a->copy_stream));
            issued[k] = true;
            return HG_OK;
        }
        static int prepare(void* user, int k, int64_t lo, int64_t hi, cudaStream_t cst) {
            Pipe* p = static_cast<Pipe*>(user);
            int prc = p->issue(k);
            if (prc != HG_OK) return prc;
            // pageable host memory makes cudaMemcpyAsync block the host: issue one chunk ahead at most, after the
            // previous chunk's kernels are queued; pinned memory was queued up front
            HG_CUDA_TRY(cudaStreamWaitEvent(cst, p->a->chunk_ev[k], 0));
            return hg_pack_rows(p->d_feat + lo * p->b, p->b, p->d_lab + (size_t)lo * p->L * p->lab_bytes, p->lab_bytes, hi - lo, p->b, p->L,
                                p->d_rows + lo * p->Wr, p->d_bad, cst);
        }
    } pipe;
    pipe.a = &a; pipe.h_feat = h_db_feat; pipe.h_lab = static_cast<const char*>(h_db_lab); pipe.d_feat = d_dbf; pipe.d_lab = base + o_dbl;
    pipe.d_rows = d_dbr; pipe.d_bad = d_bad; pipe.b = b; pipe.L = L; pipe.Wr = Wr; pipe.lab_bytes = lab_elem_bytes;
    for (bool& f : pipe.issued) f = false;
    pipe.ch = chunks;
    {
        cudaPointerAttributes pa{}, pb{};
        const bool ok = cudaPointerGetAttributes(&pa, h_db_feat) == cudaSuccess && cudaPointerGetAttributes(&pb, h_db_lab) == cudaSuccess;
        (void)cudaGetLastError();
        pipe.pinned = ok && pa.type == cudaMemoryTypeHost && pb.type == cudaMemoryTypeHost;
    }
    // the copy stream must not overtake the previous call's reads of these buffers: both streams were drained by
    // the synchronize at the end of that call
    if (n_sample > 0) {
        // threshold sample: the plan's segments, spread over the whole database, gathered by ONE strided copy (C4: 16k rows =
        // 4 MB) and packed (only the code words matter for the estimate).  The copy goes FIRST on the copy stream, ahead of the
        // chunks: queued on the compute stream it waits for the query pack kernel, the chunk copies overtake it in the H2D
        // queue and nothing is ranked before the whole database has landed (measured: 9.3 ms instead of 7.1 ms per call).
        float* d_sf = reinterpret_cast<float*>(base + o_sf);
        uint32_t* d_sr = reinterpret_cast<uint32_t*>(base + o_sr);
        const size_t seg_bytes = sizeof(float) * (size_t)chunks.sample_seg_rows * b;
        HG_CUDA_TRY(cudaMemcpy2DAsync(d_sf, seg_bytes, h_db_feat, sizeof(float) * (size_t)chunks.sample_seg_stride * b, seg_bytes,
                                      (size_t)chunks.sample_n_seg, cudaMemcpyHostToDevice, a.copy_stream));
        HG_CUDA_TRY(cudaEventRecord(a.sample_ev, a.copy_stream));
        HG_CUDA_TRY(cudaStreamWaitEvent(st, a.sample_ev, 0));
        if ((rc = hg_pack_rows(d_sf, b, nullptr, lab_elem_bytes, n_sample, b, L, d_sr, nullptr, st)) != HG_OK) return rc;
        pipe.ch.sample_packed = d_sr;
    }
    if (pipe.pinned)
        for (int k = 0; k < pipe.ch.K; ++k)
            if ((rc = pipe.issue(k)) != HG_OK) return rc;
    if ((rc = hg::hamming_map_chunked(d_qr, nq, d_dbr, ndb, b, L, R, flags, d_ap, base + o_ws, ws_bytes, st, &pipe.ch, &Pipe::prepare, &pipe)) != HG_OK)
        return rc;
    HG_CUDA_TRY(cudaMemcpyAsync(a.pinned_ap, d_ap, sizeof(double) * (size_t)nq, cudaMemcpyDeviceToHost, st));
    int* h_bad = reinterpret_cast<int*>(a.pinned_ap + nq);
    HG_CUDA_TRY(cudaMemcpyAsync(h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    HG_CUDA_TRY(cudaStreamSynchronize(st));
    drain.armed = false;  // every chunk event was waited on by `st`: both streams are idle
    if (*h_bad) return hg::fail(HG_ELABEL, "hg_maps_by_feature_host: labels must be 0/1");
    if (h_ap_out) memcpy(h_ap_out, a.pinned_ap, sizeof(double) * (size_t)nq);
    return hg_mean_ap_host(a.pinned_ap, nq, map_out, nullptr);
}
