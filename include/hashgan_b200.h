/* hashgan_b200.h -- C ABI of libhashgan_b200.so (sm_100a).
 *
 * Drop-in boundary for the retrieval-evaluation hot path of thuml/HashGAN.  The reference has no
 * FFI of its own (it is pure Python); each entry point below cites the reference code it replaces.
 * All pointers named d_* are DEVICE pointers (e.g. torch.Tensor.data_ptr()), h_* are HOST pointers.
 * `stream` is a cudaStream_t passed as void* (NULL = default stream).  Every function returns
 * 0 on success and a non-zero HG_E* code otherwise; hg_last_error() gives the message of the
 * last failure on the calling thread.  No function allocates device memory behind the caller's back
 * except hg_maps_by_feature_host (documented there).  Functions are thread-compatible: concurrent
 * calls must use distinct workspaces.
 */
#ifndef HASHGAN_B200_H
#define HASHGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HG_OK 0
#define HG_EINVAL 1      /* bad argument (NULL pointer, size, alignment, unsupported b / L) */
#define HG_ERANGE 2      /* R > ndb: the reference raises ValueError here (lib/metric.py:21 broadcast) */
#define HG_ENOMEM 3      /* workspace too small / allocation failed */
#define HG_ECUDA 4       /* CUDA runtime error, see hg_last_error() */
#define HG_ELABEL 5      /* a label was not 0/1 (lib/metric.py:17-19 is only defined for 0/1 labels) */

#define HG_MAX_BITS 256  /* hash length b supported by the kernels (reference default 64, lib/config.py:10) */
#define HG_MAX_LABELS 128  /* label width L (CIFAR-10: 10, NUS-WIDE: 81, COCO: 80) */

/* flags for hg_hamming_map */
#define HG_FLAG_FORCE_EXACT 1u  /* skip the sampled-threshold fast pass; every query takes the two-pass exact path */
#define HG_FLAG_NO_FALLBACK 2u  /* (diagnostics) do not run the exact path; failed queries get AP = -1 */
#define HG_FLAG_TIMING 4u       /* record CUDA events around the phases, read back with hg_hamming_map_phase_ms */

int hg_version(void);

/* CRC-32C (Castagnoli) of a HOST buffer, continuing from `crc` (0 to start).  Used by the TensorFlow-checkpoint reader
 * (hashgan_b200/tf_checkpoint.py) that replaces tf.train.Saver.restore, main.py:187-195.  hg_crc32c("123456789") = 0xE3069283. */
uint32_t hg_crc32c(const void* data, size_t n, uint32_t crc);
const char* hg_last_error(void);

/* Device facts used for grid sizing (multiples of the SM count). */
int hg_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes);

/* Packed-row layout shared by queries and database:
 *     row = [ W code words | LW label words | zero pad ]   (uint32, row stride hg_row_words(b, L))
 * W  = hg_code_words(b)  = ceil(b/32) for b <= 128, 8 for 128 < b <= 256 (zero pad words); 0 if unsupported.
 * LW = hg_label_words(L) = ceil(L/32), L <= HG_MAX_LABELS; 0 if unsupported.
 * The stride rounds W+LW up so that the code words can be fetched with one 64/128-bit load; 33..128-bit codes get
 * 4- or 8-word rows (16 / 32 bytes) so that the tensor-core select can stage them by TMA. */
int hg_code_words(int b);
int hg_label_words(int L);
int hg_row_words(int b, int L);

/* sign + bit-pack of features and 0/1 labels into packed rows (d_rows: [n, hg_row_words(b, L)] uint32).
 * New in the build: the reference never binarises (main.py:155-157 hands tanh outputs straight to
 * lib/metric.py:13); on {-1,+1} inputs ip = b - 2*d_H, so ranking by inner product (lib/metric.py:13-14) ==
 * ranking by Hamming distance of these words.  Code bit j of word w = (feat[i, 32w+j] > 0).
 * The label bits replace the per-query gather/compare of lib/metric.py:17-19: imatch = any_l(db.label[l] ==
 * label[l]) with the query's zeros rewritten to -1  <=>  (q_label_bits & db_label_bits) != 0 for 0/1 labels.
 * d_feat: [n, ld] float32 row-major, first b columns used.  d_lab: [n, L] int64 (the dtype np.array(list-of-int)
 * gives, lib/dataloader.py:45,74), int32 or int8/uint8 (lab_elem_bytes = 8, 4, 1); may be NULL (label bits zero).
 * d_bad (optional, int32[1]) is OR-ed with 1 if any label is not 0/1. */
int hg_pack_rows(const float* d_feat, int64_t ld, const void* d_lab, int lab_elem_bytes, int64_t n, int b, int L,
                 uint32_t* d_rows, int* d_bad, void* stream);

/* hg_pack_rows fused with the exchange step of the multi-GPU path (one all-gather of packed rows): the rows
 * are stored straight into n_dst destination buffers -- this rank's own database buffer and its peers', reached through
 * NVLink / NVSwitch peer pointers of a symmetric allocation.  h_dst_rows is a HOST array of n_dst (<= 16) DEVICE pointers,
 * each already offset to this rank's first row inside that destination, 16-byte aligned.  The caller publishes the rows
 * with a cross-GPU barrier afterwards (hashgan_b200/sharding.py: SymmetricRows).  Needs b % 32 == 0 and contiguous
 * feature rows; returns HG_ERANGE otherwise (use hg_pack_rows + an all-gather then).  The reference has no multi-GPU
 * path (main.py:263). */
int hg_pack_rows_push(const float* d_feat, int64_t ld, const void* d_lab, int lab_elem_bytes, int64_t n, int b, int L,
                      uint32_t* const* h_dst_rows, int n_dst, int* d_bad, void* stream);

/* Workspace (bytes) hg_hamming_map needs for these sizes; 0 on invalid arguments. */
size_t hg_hamming_map_workspace_bytes(int64_t nq, int64_t ndb, int b, int L, int64_t R);

/* The metric hot path, lib/metric.py:12-23 for all queries at once:
 *   lib/metric.py:13-14  all-pairs inner product + argsort   -> XOR/POPC Hamming distance + exact
 *                        (distance asc, database row asc) ranking == np.argsort(kind='stable')
 *   lib/metric.py:16-23  per-query relevance, cumsum, AP     -> integer prefix counts + fp64 divides
 * d_ap[q] = AP@R of query q, NaN where the top-R holds no relevant row (the reference skips those
 * queries, lib/metric.py:22-23).  Optional outputs (may be NULL): d_ids [nq, R] uint32 database rows in
 * rank order, d_dist [nq, R] uint16 their Hamming distances, d_rel [nq] int32 relevant count in top-R.
 * d_q_rows / d_db_rows are packed rows as produced by hg_pack_rows; d_db_rows must be 16-byte aligned.
 * Asynchronous on `stream`; the workspace must stay alive until the stream has drained. */
int hg_hamming_map(const uint32_t* d_q_rows, int64_t nq, const uint32_t* d_db_rows, int64_t ndb,
                   int b, int L, int64_t R, unsigned flags,
                   double* d_ap, uint32_t* d_ids, uint16_t* d_dist, int32_t* d_rel,
                   void* d_workspace, size_t workspace_bytes, void* stream);

/* Diagnostics of the last hg_hamming_map run on this workspace (host copy, synchronises `stream`):
 * out[0] = queries that needed the exact two-pass path, out[1] = db splits P, out[2] = rows per split,
 * out[3] = entries per candidate bin, out[4] = queries per CTA tile, out[5] = sample rows used,
 * out[6] = queries the exact path could not answer (always 0), out[7] = queries whose candidate distances
 * spanned >= 32 values (redone by the full-width AP kernel). */
int hg_hamming_map_stats(const void* d_workspace, size_t workspace_bytes, int64_t nq, int64_t ndb, int b, int L,
                         int64_t R, int64_t out[8], void* stream);

/* Device time (ms, CUDA events on the call's stream) of the phases of the last hg_hamming_map call made by this
 * thread with HG_FLAG_TIMING: out[0] sampled histogram, out[1] thresholds, out[2] int8 expansion of the codes
 * (tensor-core backend only), out[3] select (the all-pairs kernel), out[4] AP, out[5] exact-path chain.
 * Synchronises on the last event. */
int hg_hamming_map_phase_ms(float out[6]);

/* Which all-pairs kernel the row shape (b, L) can run on: 0 = select_kernel only (XOR/POPC on the integer pipe),
 * 32 / 64 / 128 / 256 = select_umma_kernel (exact int8 tcgen05.mma; the value is the int8 row width: b <= 32, <= 64, <= 128,
 * <= 256); -1 = unsupported shape.  The environment variable HG_SELECT_BACKEND=popc forces 0. */
int hg_select_backend(int b, int L);

/* The kernel hg_hamming_map really uses for a whole problem (same values, plus 1): short codes (b <= 32) keep the POPC kernel
 * when the top-R is a large part of the database (R * 8 > ndb), where every pair is a candidate and the tensor-core kernel's
 * cheap rejection buys nothing (HG_SELECT_BACKEND=umma lifts that rule); and 1 = no selection at all: with R * 2 >= ndb
 * (cifar_evaluation.yaml ranks the whole database, MAP_R == DB_SIZE) dense_ap_kernel walks the packed rows directly
 * (HG_DENSE=0 / 1 forces the choice). */
int hg_select_backend_for(int64_t nq, int64_t ndb, int b, int L, int64_t R);

/* 1 when that tensor-core select is the QUEUED kernel (select_q_kernel: non-zero hit-mask words are parked in per-lane
 * shared-memory FIFOs and consumed one hit per lane and tile, around the wait for the next accumulator), 0 when the hits
 * are walked tile by tile (select_umma_kernel).  The plan queues while the top-R is sparse (about two hits per lane and
 * 64-row tile or fewer) on two-word codes in four-word rows (C4); HG_SELECT_MODE=queue / lists forces the choice. */
int hg_select_queued_for(int64_t nq, int64_t ndb, int b, int L, int64_t R);

/* Real-valued ranking mode -- the reference's literal behaviour on un-binarised features:
 *   lib/metric.py:13  ips = np.dot(query.output, database.output.T)   fp32 inner products (FMA, increasing k)
 *   lib/metric.py:14  np.argsort(-ips, 1)[:, :R]                       exact top-R by (ip descending, database row ascending)
 *   lib/metric.py:16-23  per-query relevance / cumsum / AP             label words of the packed rows (hg_pack_rows)
 * d_q_feat [nq, b] / d_db_feat [ndb, b] fp32 row-major; d_q_rows / d_db_rows packed rows of the same inputs (only their label
 * words are read).  d_ap [nq] (NaN where no relevant row is in the top-R); optional d_ids [nq, R] (database rows in rank
 * order), d_ips [nq, R] (their inner products), d_rel [nq].  The workspace holds the keys of one query chunk
 * (4 B per pair; hg_ip_map_workspace_bytes sizes it for chunks of up to 592 queries, any size from one query up works).
 * Exact by construction (radix select + stable radix sort): no sampling, no fallback.  Asynchronous on `stream`. */
size_t hg_ip_map_workspace_bytes(int64_t nq, int64_t ndb, int b, int L, int64_t R);
int hg_ip_map(const float* d_q_feat, const uint32_t* d_q_rows, int64_t nq, const float* d_db_feat, const uint32_t* d_db_rows,
              int64_t ndb, int b, int L, int64_t R, double* d_ap, uint32_t* d_ids, float* d_ips, int32_t* d_rel,
              void* d_workspace, size_t workspace_bytes, void* stream);

/* Number of kernels this library has launched in this process so far (reset != 0 zeroes the counter). */
int64_t hg_launch_count(int reset);

/* lib/metric.py:24 on the device result: mean over non-NaN entries, computed on the host in fp64 from a
 * D2H copy of d_ap (synchronises `stream`).  *map_out = NaN when every query was skipped. */
int hg_mean_ap(const double* d_ap, int64_t nq, double* map_out, int64_t* n_used, void* stream);

/* Relevant rows of the WHOLE database per query: d_total[q] = #{ rows r : labels(q) & labels(r) != 0 } -- the relevance test of
 * lib/metric.py:17-19 applied to every row instead of the top-R.  With hg_hamming_map's d_rel (relevant rows inside the top-R,
 * lib/metric.py:20) it gives the companions of mAP@R on the same ranking:
 *     precision@R = rel / R          recall@R = rel / total.
 * Packed rows as produced by hg_pack_rows (only the label words are read).  d_total: [nq] uint32. */
int hg_relevant_totals(const uint32_t* d_q_rows, int64_t nq, const uint32_t* d_db_rows, int64_t ndb, int b, int L, uint32_t* d_total,
                       void* stream);

/* Host-only half of hg_mean_ap: lib/metric.py:24 (`np.mean(np.array(apx))`) on a HOST vector of per-query APs --
 * NaN entries are the queries the reference skips (lib/metric.py:22-23), the kept ones are summed with NumPy's
 * pairwise scheme so the result is bit-identical to np.mean.  Needs no GPU (tests/test_abi.py pins it against NumPy). */
int hg_mean_ap_host(const double* h_ap, int64_t nq, double* map_out, int64_t* n_used);

/* End-to-end convenience with HOST buffers == MAPs(R).get_maps_by_feature(database, query)
 * (lib/metric.py:12-24, call site main.py:164).  Copies features/labels H2D (pinned staging, chunked and
 * overlapped with packing), ranks, copies the per-query AP back.  Device memory is taken from a cached
 * per-process arena that grows on demand (freed by hg_release_cached()).  h_ap_out may be NULL.
 * lab_elem_bytes: 8 / 4 / 1. */
int hg_maps_by_feature_host(const float* h_db_feat, const void* h_db_lab, int64_t ndb,
                            const float* h_q_feat, const void* h_q_lab, int64_t nq,
                            int b, int L, int lab_elem_bytes, int64_t R, unsigned flags,
                            double* map_out, double* h_ap_out);
int hg_release_cached(void);

/* Dense layer on the tensor cores (tcgen05.mma kind::tf32, fp32 accumulation in TMEM):
 *     C[M,N] = act(A[M,K] * Bt[N,K]^T + bias[N]),  act = ReLU when relu != 0
 * Replaces tf.matmul + tf.nn.bias_add (+ tf.nn.relu) of the fully connected layers, lib/architecture.py:363-377
 * (fc6, fc7) and lib/ops.py:287-302 (`linear`, fc8).  A and Bt are fp32 row-major with row strides lda / ldb
 * (elements, multiples of 4; 16-byte aligned bases); Bt is the weight matrix TRANSPOSED to [N, K].  K % 32 == 0.
 * d_bias may be NULL.  Asynchronous on `stream`. */
int hg_gemm_tf32(const float* d_a, int64_t lda, const float* d_bt, int64_t ldb, const float* d_bias, float* d_c, int64_t ldc,
                 int M, int N, int K, int relu, void* stream);

/* AlexNet hash head, stage='val' (lib/architecture.py:196-392 with main.py:144-148 and lib/util.py:12-21 in front):
 * uint8 images -> crop-averaged tanh outputs in (-1, 1), the `.output` that main.py:155-157 hands to the metric.
 * Weight tensors are fp32 device pointers in the reference's own layouts and names (lib/params.py registry):
 *   conv_w[i] HWIO: discriminator.conv1..5.weights = [11,11,3,96] [5,5,48,256] [3,3,256,384] [3,3,192,384] [3,3,192,256]
 *   conv_b[i]     : discriminator.conv1..5.biases
 *   fc6_wt / fc7_wt / fc8_wt: discriminator.fc6.weights [9216,4096], discriminator.fc7.weights [4096,4096],
 *                   discriminator.ACGANOutput.W [4096,HASH_DIM] -- each TRANSPOSED to [N, K] (hg_transpose_f32), so the
 *                   tensor-core GEMM reads both operands K-major; fc6 rows are in (h, w, c) order as in the reference
 *   fc6_b / fc7_b / fc8_b: discriminator.fc6.biases, discriminator.fc7.biases, discriminator.ACGANOutput.b */
typedef struct HgAlexNetWeights {
    const float* conv_w[5];
    const float* conv_b[5];
    const float* fc6_wt; const float* fc6_b;
    const float* fc7_wt; const float* fc7_b;
    const float* fc8_wt; const float* fc8_b;
    const float* conv_wt[5];  /* tensor-core convolutions: hg_conv_weight_pack of conv_w[i] */
    const float* fc_wt3[3];   /* optional, HG_ENC_CONV_TF32X3: hg_conv_weight_pack(W, 1, 1, K, N, 1) of the fc6 / fc7 / fc8 matrices
                               * [K, N]; when present the dense layers also run error-compensated (fp32-grade end to end) */
    const float* conv1_fused; /* optional, HG_ENC_FUSED_STAGE1: hg_conv1_fused_pack of conv_w[0] for images of conv1_fused_wh pixels */
    int conv1_fused_wh;
} HgAlexNetWeights;

#define HG_ENC_LRN 1u        /* local response normalisation after pool1/pool2: on iff TRAIN.WGAN_SCALE == 0 (architecture.py:268,294) */
#define HG_ENC_CONV_TF32 2u  /* opt-in: conv1-5 as implicit GEMM on tcgen05 with plain TF32 operands (6x faster than the fp32 CUDA cores, ~1e-3 relative error) */
#define HG_ENC_FUSED_STAGE1 32u /* normalise + resize + 10-crop + mean + conv1 + ReLU + pool1 (+ LRN) as ONE kernel on effective filters
                                 * (csrc/encoder_stage1.cu): needs conv1_fused packed for this wh (32 or 64) */
#define HG_ENC_TIMING 16u    /* record CUDA events between the stages of this call (diagnostics): hg_alexnet_phase_ms */
#define HG_ENC_CONV_TF32X3 8u /* conv1-5 (and fc6-8 when fc_wt3 is set) as implicit GEMM on tcgen05 with error-compensated TF32 (hi/lo split, 3 MMAs): fp32-grade accuracy */

/* Workspace bytes for a batch of n images (10 n crops) with these flags. */
size_t hg_alexnet_workspace_bytes(int n, unsigned flags);

/* HWIO convolution weights [KH, KW, Cg, Cout] -> per-group K-major [groups][Cout/groups][Kpad] (channels per group padded
 * to a multiple of 4, Kpad = KH*KW*Cg4 rounded up to 32, zero padded): the B operand of the tensor-core convolution.  d_out holds 2 * Cout * Kpad floats: first the upper 19 bits of
 * every weight (exactly TF32), then the remainders w - hi used by the error-compensated mode (HG_ENC_CONV_TF32X3). */
int hg_conv_weight_pack(const float* d_w_hwio, int KH, int KW, int Cg, int Cout, int groups, float* d_out, void* stream);

/* Effective conv1 filters of the fused first stage (HG_ENC_FUSED_STAGE1).  Resize (lib/util.py:19), crop offset, flip
 * (lib/architecture.py:215-244) and conv1 (:253-258) are linear and compose exactly: per crop type (corner / centre x plain /
 * flipped) and per phase of the output pixel in the source grid, an 11 x 11 filter over the 256-pixel image becomes a
 * WIN x WIN filter (4 for wh = 32, 5 for wh = 64) over the SOURCE image.  d_conv1_hwio: discriminator.conv1.weights
 * [11, 11, 3, 96]; d_out: hg_conv1_fused_floats(wh) floats (0 = this image size is not supported: use the unfused path). */
size_t hg_conv1_fused_floats(int wh);
int hg_conv1_fused_pack(const float* d_conv1_hwio, int wh, float* d_out, void* stream);

/* d_images: uint8 [n, 3, wh, wh], RGB planes -- the loader's flattened batch (lib/dataloader.py:110-113), wh <= 256.
 * d_out: float32 [n, hash_dim].  Deterministic mode only: no de-quantisation noise (main.py:147) and no eval-time
 * dropout (architecture.py:369,377).  fc6-8 run on tcgen05 tensor cores (TF32), conv1-5 on the CUDA cores (fp32). */
int hg_alexnet_encode(const uint8_t* d_images, int n, int wh, const HgAlexNetWeights* w, int hash_dim, unsigned flags,
                      float* d_out, void* d_workspace, size_t workspace_bytes, void* stream);

/* The reference's STOCHASTIC evaluation graph: Model.normalize adds U(0, 1/128) de-quantisation noise to every pixel also at
 * eval (main.py:147) and tf.nn.dropout(x, 0.5) after fc6 / fc7 has no stage guard (lib/architecture.py:369,377), so the
 * reference's eval outputs are random.  hg_alexnet_encode is the expectation-free deterministic graph (no noise, no dropout);
 * this entry point reproduces the stochastic one with a counter-based generator: draw(seed, stream, index) =
 * splitmix64-finaliser(seed ^ stream + 0x9E3779B97F4A7C15 * (index + 1)), noise = top 24 bits * 2^-24 / 128 per source pixel
 * (index = flat position in d_images), dropout keeps an activation iff the top bit is set (index = crop_row * 4096 + unit;
 * crop rows in the crop-major order of lib/architecture.py:242-244).  hashgan_b200/encoder.py holds the same generator in
 * NumPy, so tests feed identical draws to the fp32 oracle.  seed != 0. */
int hg_alexnet_encode_stochastic(const uint8_t* d_images, int n, int wh, const HgAlexNetWeights* w, int hash_dim, unsigned flags,
                                 float* d_out, void* d_workspace, size_t workspace_bytes, uint64_t seed, void* stream);

/* out[c][r] = in[r][c] (fp32, [rows, cols] -> [cols, rows]); used once per model to transpose the fc weights. */
int hg_transpose_f32(const float* d_in, int rows, int cols, float* d_out, void* stream);

/* Integer-pipe microbenchmark: measured XOR+POPC word-ops per second of this GPU (the binding roofline
 * of the Hamming kernel).  Runs `iters` dependent-free popc chains on every SM. */
int hg_popc_peak(double* wordops_per_s, double* ms, int iters, void* stream);

/* Tensor-pipe microbenchmark: measured int8 tcgen05.mma throughput (multiply + add = 2 ops) of this GPU -- every SM issues
 * `iters` x 4 back-to-back tcgen05.mma kind::i8 (M 128, N 256, K 32) on resident shared-memory tiles; best of 3 launches
 * (iters <= 0: 8192).  The denominator of bench.py's roofline.tensor for the tensor-core select kernel. */
int hg_i8_peak(double* ops_per_s, double* ms, int iters, void* stream);

/* Milliseconds of the stages of the last hg_alexnet_encode[_stochastic] call made with HG_ENC_TIMING on this thread
 * (synchronises on its last event): out = { crops (main.py:144-148, lib/util.py:12-21, architecture.py:215-249), conv1-5,
 * max-pool + LRN, fc6-8 (+ dropout), tanh + crop mean }.  With HG_ENC_FUSED_STAGE1 out[0] is the fused first stage
 * (crops + conv1 + pool1 + LRN1) and out[1] / out[2] hold conv2-5 / pool2, LRN2, pool5 only. */
int hg_alexnet_phase_ms(float out[5]);

#ifdef __cplusplus
}
#endif
#endif /* HASHGAN_B200_H */
