// Fused first stage of the AlexNet hash-head forward: uint8 images -> pool1 (+LRN) output, one kernel.
//
// What it replaces (reference call sites):
//   main.py:144-148               normalize 2x/256 - 1 (+ U(0,1/128) de-quantisation noise in the stochastic mode)
//   lib/util.py:12-21             (x+1)*255.99/2, NCHW -> NHWC, tf.image.resize_bilinear wh -> 256 (TF1 legacy sampling)
//   lib/architecture.py:215-249   10 crops 227x227 (5 of the flipped image, 5 plain), minus the channel mean
//   lib/architecture.py:253-258   conv1 11x11 stride 4 VALID 3 -> 96, bias, ReLU
//   lib/architecture.py:261-271   max pool 3x3 stride 2, LRN (when TRAIN.WGAN_SCALE == 0)
//
// The separate kernels materialise the crop tensor [10n,227,227,4] (1 GB per 128-image batch) and conv1's output
// [10n,55,55,96] (1.5 GB) and spend 6.6 of the 18 ms of a batch there (conv1 gathers 121 taps per output pixel).  But every
// value conv1 reads is a bilinear interpolation of a handful of SOURCE pixels: with s = 256 / wh (8 for CIFAR's 32x32), the
// 11 taps of a filter row starting at position Y of the 256-image touch source rows floor(Y/s) .. floor((Y+10)/s)+1 -- at most
// WIN = 4 rows (5 for wh = 64) -- with weights that depend only on Y mod s.  Resize, crop offset, flip (= the filter mirrored
// in x) and conv1 are all linear, so they compose EXACTLY into an effective WIN x WIN x 3 -> 96 filter per phase
// (Y mod s, X mod s): K = 48 instead of 363 multiply-adds per output, no crop tensor, no 256-image.  The bottom/right clamp of
// the legacy bilinear (min(ceil(f), wh-1)) is a replicated border of the source image.  The mean is subtracted from the source
// pixels (bilinear weights sum to one).  hg_conv1_fused_pack builds the effective filters once per model in fp64; the kernel
// runs the small contraction in fp32 on the CUDA cores (exact algebra: it differs from the unfused path only by fp32
// summation order), pools the 55x55 map row by row in shared memory and applies the LRN before anything is written.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

namespace hg {

constexpr int kS1Threads = 352;  // 11 warps: 336 = 4 pixel parities x 7 pixel groups x 12 channel groups compute, all pool
constexpr int kS1Compute = 336;
constexpr int kS1Cout = 96;
constexpr int kS1Types = 4;      // crop type = (centre crop ? 2 : 0) + (flipped ? 1 : 0)

__host__ __device__ constexpr int s1_win(int wh) { return wh == 32 ? 4 : 5; }          // source rows a filter row can touch
__host__ __device__ constexpr int s1_sets_per_axis(int wh) { return wh == 32 ? 2 : 1; }  // phases per axis inside one crop type

// same helpers as encoder.cu (kept bit-identical: the stochastic draws are pinned by tests/test_encoder_host.py)
__host__ __device__ __forceinline__ uint64_t s1_mix(uint64_t seed, uint64_t idx)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
constexpr uint64_t kS1StreamNoise = 0xD1B54A32D192ED03ull;

// ---- effective filters ------------------------------------------------------------------------------------------------
// out[type][set = a * NP + b][k = (dy * WIN + dx) * 3 + ch][o]; phase of the set: py = py0(type) + 4a, px = px0(type) + 4b
// with py0 = 0 (corner crops: offsets 0 / 28) or 2 (centre crop: offset 14) and px0 the same for plain crops, and
// (245 - ox) mod 4 = 1 (corner) / 3 (centre) for flipped crops, whose taps run right to left: tap position u = 10 - tx.
__global__ void __launch_bounds__(256) conv1_fused_pack_kernel(const float* __restrict__ w_hwio, int wh, float* __restrict__ out)
{
    const int WIN = s1_win(wh), NP = s1_sets_per_axis(wh), s = 256 / wh, K = WIN * WIN * 3;
    const int total = kS1Types * NP * NP * K * kS1Cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int o = i % kS1Cout;
        const int k = (i / kS1Cout) % K;
        const int set = (i / (kS1Cout * K)) % (NP * NP);
        const int type = i / (kS1Cout * K * NP * NP);
        const int ch = k % 3, dx = (k / 3) % WIN, dy = k / (3 * WIN);
        const bool centre = (type & 2) != 0, flip = (type & 1) != 0;
        const int py = (centre ? 2 : 0) + 4 * (set / NP);
        const int px = (flip ? (centre ? 3 : 1) : (centre ? 2 : 0)) + 4 * (set % NP);
        double acc = 0.0;
        for (int ty = 0; ty < 11; ++ty) {
            const int Py = py + ty, ay = Py / s;
            const double fy = (double)(Py % s) / (double)s;
            const double wy = (dy == ay ? 1.0 - fy : 0.0) + (dy == ay + 1 ? fy : 0.0);
            if (wy == 0.0) continue;
            for (int u = 0; u < 11; ++u) {
                const int Px = px + u, ax = Px / s;
                const double fx = (double)(Px % s) / (double)s;
                const double wx = (dx == ax ? 1.0 - fx : 0.0) + (dx == ax + 1 ? fx : 0.0);
                if (wx == 0.0) continue;
                const int tx = flip ? 10 - u : u;
                acc += wy * wx * (double)w_hwio[((ty * 11 + tx) * 3 + ch) * kS1Cout + o];
            }
        }
        out[i] = (float)acc;
    }
}

struct Stage1Params {
    const unsigned char* img;  // [n, 3, wh, wh]
    int n, wh;
    const float* wfused;       // hg_conv1_fused_pack
    const float* bias;         // [96]
    float* out;                // [10 n, 27, 27, 96]
    uint64_t seed;             // 0 = deterministic (no de-quantisation noise)
    int lrn;
};

// One CTA = one crop of one image (crop-major rows: nn = crop * n + image, lib/architecture.py:242-244).
// Per step m the CTA computes conv rows 2m and 2m+1 (all 55 columns, 96 channels) into shared memory: thread <-> (pixel
// parity (row, column), 4 pixels of that parity, 8 output channels) -- pixels of one parity share the phase, hence the filter
// set.  Then pooled row m-1 is completed with conv row 2m, normalised and written, and pooled row m is started.
// WSMEM: the filter sets of the crop type live in shared memory (74 KB for wh = 32: one CTA per SM); otherwise they are read
// through L1 (they are hot: every CTA of the SM uses one of 4 types) and two CTAs share an SM (the register file allows no more: 32 accumulators per thread).
template <int WIN, bool WSMEM>
__global__ void __launch_bounds__(kS1Threads, WSMEM ? 1 : 2) conv1_stage_kernel(Stage1Params p)
{
    constexpr int K = WIN * WIN * 3;
    extern __shared__ __align__(16) float s1sm[];
    const int wh = p.wh, sw = wh + 2, s = 256 / wh, NP = s / 4;
    float* const Wsm = s1sm;                                     // [NP*NP][K][96] (WSMEM only)
    float* const Ssm = Wsm + (WSMEM ? NP * NP * K * kS1Cout : 0); // [sw][sw][4] source pixels (scaled, mean subtracted, replicated border; RGB + pad: one 16-byte load per pixel)
    float* const rowbuf = Ssm + sw * sw * 4;                      // [2][56][96] conv rows 2m, 2m+1 (bias + ReLU applied)
    float* const pool = rowbuf + 2 * 56 * kS1Cout;                // [27][96] pooled row under construction
    const int tid = threadIdx.x;
    const int nn = blockIdx.x, crop = nn / p.n, b = nn % p.n;
    const int kk = crop % 5;
    const int oy = (kk == 1 || kk == 2) ? 28 : (kk == 4 ? 14 : 0);  // (0,0) (28,28) (28,0) (0,28) (14,14)
    const int ox = (kk == 1 || kk == 3) ? 28 : (kk == 4 ? 14 : 0);
    const bool flip = crop < 5;                                     // crops 0..4 come from the left-right flipped image
    const int type = (kk == 4 ? 2 : 0) + (flip ? 1 : 0);

    {
        if (WSMEM) {
            const float4* src = reinterpret_cast<const float4*>(p.wfused + (size_t)type * NP * NP * K * kS1Cout);
            float4* dst = reinterpret_cast<float4*>(Wsm);
            for (int i = tid; i < NP * NP * K * kS1Cout / 4; i += kS1Threads) dst[i] = __ldg(src + i);
        }
        const float mean[3] = {103.939f, 116.779f, 123.68f};
        for (int i = tid; i < sw * sw * 4; i += kS1Threads) {
            const int ch = i & 3, j = (i >> 2) % sw, r = (i >> 2) / sw;
            if (ch == 3) { Ssm[i] = 0.0f; continue; }
            const int64_t flat = ((int64_t)b * 3 + ch) * wh * wh + min(r, wh - 1) * wh + min(j, wh - 1);
            float noise = 0.0f;  // main.py:147: one draw per source pixel, shared by the 10 crops (same index as prep_crops_kernel)
            if (p.seed) noise = (float)(s1_mix(p.seed ^ kS1StreamNoise, (uint64_t)flat) >> 40) * (1.0f / 16777216.0f) * (1.0f / 128.0f);
            const float x = 2.0f * (float)p.img[flat] / 256.0f - 1.0f + noise;   // main.py:146-147
            Ssm[i] = (x + 1.0f) * 255.99f / 2.0f - mean[ch];                        // lib/util.py:13, lib/architecture.py:247-249
        }
    }
    __syncthreads();

    // compute role
    const int cg = tid % 12, g = (tid / 12) % 7, sub = tid / 84;  // sub = 2 * row parity + column parity (tid < 336)
    const int sr = sub >> 1, sc = sub & 1, ch0 = cg * 8;
    float bias[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bias[i] = tid < kS1Compute ? __ldg(p.bias + ch0 + i) : 0.0f;
    // columns of my 4 pixels and their window origins (constant over the rows)
    int col[4], xoff[4], setx = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        col[q] = 2 * (4 * g + q) + sc;
        const int c = min(col[q], 54);
        const int xb = flip ? 245 - ox - 4 * c : ox + 4 * c;  // position of tap u = 0 in the 256-wide image
        xoff[q] = (xb / s) * 4;
        if (q == 0) setx = (xb % s) >> 2;                       // same for my 4 pixels (their columns differ by multiples of 2: 8 positions)
    }

    float* const outp = p.out + (size_t)nn * 27 * 27 * kS1Cout;
    for (int m = 0; m < 28; ++m) {
        const int r = 2 * m + sr;
        if (tid < kS1Compute && r <= 54) {
            const int yb = oy + 4 * r;
            const float* wp = (WSMEM ? Wsm : p.wfused + (size_t)type * NP * NP * K * kS1Cout) + (size_t)(((yb % s) >> 2) * NP + setx) * K * kS1Cout + ch0;
            const float* sp = Ssm + (yb / s) * sw * 4;
            float acc[4][8];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[q][i] = 0.0f;
#pragma unroll 1
            for (int dy = 0; dy < WIN; ++dy) {
                const float* srow = sp + dy * sw * 4;
#pragma unroll
                for (int dx = 0; dx < WIN; ++dx) {
                    float a[4][3];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 v = *reinterpret_cast<const float4*>(srow + xoff[q] + dx * 4);
                        a[q][0] = v.x; a[q][1] = v.y; a[q][2] = v.z;
                    }
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const float* wk = wp + ((dy * WIN + dx) * 3 + ch) * kS1Cout;
                        const float4 w0 = WSMEM ? *reinterpret_cast<const float4*>(wk) : __ldg(reinterpret_cast<const float4*>(wk));
                        const float4 w1 = WSMEM ? *reinterpret_cast<const float4*>(wk + 4) : __ldg(reinterpret_cast<const float4*>(wk + 4));
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float x = a[q][ch];
                            acc[q][0] = fmaf(x, w0.x, acc[q][0]); acc[q][1] = fmaf(x, w0.y, acc[q][1]);
                            acc[q][2] = fmaf(x, w0.z, acc[q][2]); acc[q][3] = fmaf(x, w0.w, acc[q][3]);
                            acc[q][4] = fmaf(x, w1.x, acc[q][4]); acc[q][5] = fmaf(x, w1.y, acc[q][5]);
                            acc[q][6] = fmaf(x, w1.z, acc[q][6]); acc[q][7] = fmaf(x, w1.w, acc[q][7]);
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (col[q] <= 54) {
                    float4* dst = reinterpret_cast<float4*>(rowbuf + (sr * 56 + col[q]) * kS1Cout + ch0);
                    dst[0] = make_float4(fmaxf(acc[q][0] + bias[0], 0.f), fmaxf(acc[q][1] + bias[1], 0.f), fmaxf(acc[q][2] + bias[2], 0.f), fmaxf(acc[q][3] + bias[3], 0.f));
                    dst[1] = make_float4(fmaxf(acc[q][4] + bias[4], 0.f), fmaxf(acc[q][5] + bias[5], 0.f), fmaxf(acc[q][6] + bias[6], 0.f), fmaxf(acc[q][7] + bias[7], 0.f));
                }
            }
        }
        __syncthreads();
        // pooled row m-1 = max(rows 2m-2, 2m-1 (already in pool[]), row 2m)
        if (m >= 1) {
            for (int i = tid; i < 27 * kS1Cout; i += kS1Threads) {
                const int pc = i / kS1Cout, ch = i % kS1Cout;
                const float* r0 = rowbuf + (2 * pc) * kS1Cout + ch;
                pool[i] = fmaxf(pool[i], fmaxf(r0[0], fmaxf(r0[kS1Cout], r0[2 * kS1Cout])));
            }
            __syncthreads();
            float* orow = outp + (size_t)(m - 1) * 27 * kS1Cout;
            for (int i = tid; i < 27 * kS1Cout; i += kS1Threads) {
                float v = pool[i];
                if (p.lrn) {  // tf.nn.local_response_normalization(depth_radius=2, bias=1, alpha=2e-5, beta=0.75): same arithmetic as lrn_kernel
                    const int ch = i % kS1Cout;
                    float sum = 0.0f;
#pragma unroll
                    for (int d = -2; d <= 2; ++d) {
                        const int cc = ch + d;
                        if (cc >= 0 && cc < kS1Cout) { const float t = pool[i + d]; sum = fmaf(t, t, sum); }
                    }
                    v = v * powf(1.0f + 2e-5f * sum, -0.75f);
                }
                orow[i] = v;
            }
            __syncthreads();
        }
        if (m <= 26) {  // pooled row m starts with conv rows 2m and 2m+1
            for (int i = tid; i < 27 * kS1Cout; i += kS1Threads) {
                const int pc = i / kS1Cout, ch = i % kS1Cout;
                const float* r0 = rowbuf + (2 * pc) * kS1Cout + ch;
                const float* r1 = r0 + 56 * kS1Cout;
                pool[i] = fmaxf(fmaxf(r0[0], fmaxf(r0[kS1Cout], r0[2 * kS1Cout])), fmaxf(r1[0], fmaxf(r1[kS1Cout], r1[2 * kS1Cout])));
            }
        }
        __syncthreads();  // rowbuf is overwritten by the next step
    }
}

static size_t s1_smem_bytes(int wh, bool wsmem)
{
    const int WIN = s1_win(wh), NP = s1_sets_per_axis(wh), K = WIN * WIN * 3, sw = wh + 2;
    return sizeof(float) * ((wsmem ? (size_t)NP * NP * K * kS1Cout : 0) + (size_t)sw * sw * 4 + 2 * 56 * kS1Cout + 27 * kS1Cout);
}

bool stage1_supported(int wh) { return wh == 32 || wh == 64; }

size_t stage1_weight_floats(int wh)
{
    if (!stage1_supported(wh)) return 0;
    const int WIN = s1_win(wh), NP = s1_sets_per_axis(wh);
    return (size_t)kS1Types * NP * NP * WIN * WIN * 3 * kS1Cout;
}

int stage1_launch(const unsigned char* img, int n, int wh, const float* wfused, const float* bias, float* out, uint64_t seed, bool lrn, cudaStream_t st)
{
    if (!stage1_supported(wh)) return fail(HG_EINVAL, "fused conv1 stage: image size %d is not supported (32 or 64)", wh);
    Stage1Params p{img, n, wh, wfused, bias, out, seed, lrn ? 1 : 0};
    static const bool wsmem = []() { const char* v = getenv("HG_STAGE1_WSMEM"); return v && *v == '1'; }();  // default: filters through L1, 2 CTAs per SM
    const size_t smem = s1_smem_bytes(wh, wsmem);
    auto launch = [&](auto kernel) -> int {
        HG_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<10 * n, kS1Threads, smem, st>>>(p);
        return HG_OK;
    };
    int rc;
    if (wh == 32) rc = wsmem ? launch(conv1_stage_kernel<4, true>) : launch(conv1_stage_kernel<4, false>);
    else rc = wsmem ? launch(conv1_stage_kernel<5, true>) : launch(conv1_stage_kernel<5, false>);
    if (rc != HG_OK) return rc;
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

}  // namespace hg

extern "C" size_t hg_conv1_fused_floats(int wh) { return hg::stage1_weight_floats(wh); }

extern "C" int hg_conv1_fused_pack(const float* d_conv1_hwio, int wh, float* d_out, void* stream)
{
    if (!d_conv1_hwio || !d_out) return hg::fail(HG_EINVAL, "hg_conv1_fused_pack: NULL pointer");
    if (!hg::stage1_supported(wh)) return hg::fail(HG_ERANGE, "hg_conv1_fused_pack: image size %d is not supported (32 or 64)", wh);
    const size_t total = hg::stage1_weight_floats(wh);
    hg::conv1_fused_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_conv1_hwio, wh, d_out);
    hg::count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}
