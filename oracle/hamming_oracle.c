/* CPU oracle (plain C) for the binarised retrieval metric -- TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this
 * library.  It is never linked into, nor called from, the product (libhashgan_b200.so).
 *
 * What it restates: thuml/HashGAN lib/metric.py:12-24 (MAPs.get_maps_by_feature) on
 * {-1,+1}^b inputs, in the packed-bit domain:
 *   lib/metric.py:13  ips = dot(q, db.T)            ->  ip = b - 2*popcount(q ^ db)
 *   lib/metric.py:14  ids = argsort(-ips, 1)        ->  stable counting sort by Hamming
 *                                                     distance == np.argsort(kind='stable')
 *                                                     == order (distance asc, db row asc)
 *   lib/metric.py:17-19 imatch                      ->  (q_label_bits & db_label_bits) != 0
 *   lib/metric.py:20-23 rel, px, AP                 ->  same integers, same fp64 divides
 * Parity is pinned: tests/test_oracle.py checks this file against oracle/maps_oracle.py
 * (NumPy restatement), the unmodified reference (where /root/reference is mounted) and the
 * committed golden vectors under tests/golden/ (made by oracle/gen_golden.py from the
 * unmodified reference).
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC -o oracle/_build/libhamming_oracle.so oracle/hamming_oracle.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HGO_MAX_BITS 4096

/* bit j of word w = feat[i, 32w+j] > 0 ; pad bits are zero. */
int hgo_pack_sign_f32(const float* feat, int64_t n, int b, uint32_t* codes)
{
    if (!feat || !codes || n < 0 || b <= 0) return 1;
    const int W = (b + 31) / 32;
    for (int64_t i = 0; i < n; ++i) {
        for (int w = 0; w < W; ++w) {
            uint32_t word = 0;
            for (int j = 0; j < 32; ++j) {
                int c = 32 * w + j;
                if (c < b && feat[i * (int64_t)b + c] > 0.0f) word |= (1u << j);
            }
            codes[i * (int64_t)W + w] = word;
        }
    }
    return 0;
}

/* bit j of word w = (lab[i, 32w+j] == 1).  The reference compares db.label == label where the
 * query's zeros were rewritten to -1 (lib/metric.py:17-19), so a match needs a shared 1. */
int hgo_pack_labels_i64(const int64_t* lab, int64_t n, int L, uint32_t* packed)
{
    if (!lab || !packed || n < 0 || L <= 0) return 1;
    const int LW = (L + 31) / 32;
    for (int64_t i = 0; i < n; ++i) {
        for (int w = 0; w < LW; ++w) {
            uint32_t word = 0;
            for (int j = 0; j < 32; ++j) {
                int c = 32 * w + j;
                if (c < L && lab[i * (int64_t)L + c] == 1) word |= (1u << j);
            }
            packed[i * (int64_t)LW + w] = word;
        }
    }
    return 0;
}

static inline int hamming_words(const uint32_t* a, const uint32_t* b, int W)
{
    int d = 0;
    for (int w = 0; w < W; ++w) d += __builtin_popcount(a[w] ^ b[w]);
    return d;
}

/* Per-query AP@R (NaN where no relevant item is in the top-R), optional rel / ids / dist.
 * ids/dist are [nq, R] in rank order.  Returns 0 on success; 2 if R > ndb (the reference
 * raises ValueError there, lib/metric.py:21 broadcast). */
int hgo_hamming_map(const uint32_t* q_codes, const uint32_t* q_lab, int64_t nq,
                    const uint32_t* db_codes, const uint32_t* db_lab, int64_t ndb,
                    int b, int L, int64_t R,
                    double* ap, int64_t* rel_out, uint32_t* ids_out, uint16_t* dist_out,
                    int nthreads)
{
    if (b <= 0 || b > HGO_MAX_BITS || L <= 0 || nq < 0 || ndb < 0 || R <= 0) return 1;
    if (R > ndb) return 2;
    const int W = (b + 31) / 32, LW = (L + 31) / 32;
    int fail = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel
    {
        uint16_t* dist = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)(ndb ? ndb : 1));
        uint32_t* order = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)R);
        int64_t* start = (int64_t*)malloc(sizeof(int64_t) * (size_t)(b + 2));
        if (!dist || !order || !start) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for schedule(dynamic, 4)
            for (int64_t q = 0; q < nq; ++q) {
                const uint32_t* qc = q_codes + q * W;
                const uint32_t* ql = q_lab + q * LW;
                memset(start, 0, sizeof(int64_t) * (size_t)(b + 2));
                for (int64_t i = 0; i < ndb; ++i) {
                    int d = hamming_words(qc, db_codes + i * W, W);
                    dist[i] = (uint16_t)d;
                    start[d + 1]++;
                }
                for (int d = 0; d <= b; ++d) start[d + 1] += start[d];
                /* stable scatter of the rows whose rank falls below R */
                for (int64_t i = 0; i < ndb; ++i) {
                    int64_t pos = start[dist[i]]++;
                    if (pos < R) order[pos] = (uint32_t)i;
                }
                int64_t rel = 0;
                double acc = 0.0;
                for (int64_t r = 0; r < R; ++r) {
                    const uint32_t* dl = db_lab + (int64_t)order[r] * LW;
                    uint32_t m = 0;
                    for (int w = 0; w < LW; ++w) m |= (ql[w] & dl[w]);
                    if (m) {
                        rel++;
                        acc += (double)rel / (double)(r + 1);
                    }
                    if (ids_out) ids_out[q * R + r] = order[r];
                    if (dist_out) dist_out[q * R + r] = dist[order[r]];
                }
                if (rel_out) rel_out[q] = rel;
                ap[q] = rel ? acc / (double)rel : NAN;
            }
        }
        free(dist);
        free(order);
        free(start);
    }
    return fail ? 3 : 0;
}
