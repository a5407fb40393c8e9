// AlexNet hash-head forward (stage='val'): uint8 images -> [n, HASH_DIM] float32 in (-1, 1).
//
// Restates, as hand-written CUDA, the evaluation graph of the reference:
//   main.py:144-148               normalize: 2*x/256 - 1   (de-quantisation noise off: deterministic mode)
//   lib/util.py:12-21             (x+1)*255.99/2, NCHW -> NHWC, tf.image.resize_bilinear -> 256x256 (TF1 legacy sampling)
//   lib/architecture.py:215-249   10 crops 227x227 (5 of the left-right flipped image, 5 plain), minus the channel mean
//   lib/architecture.py:253-359   conv1..conv5 (+bias, ReLU), 3x3/2 max pools, LRN when TRAIN.WGAN_SCALE == 0
//   lib/architecture.py:363-382   fc6, fc7 (+bias, ReLU; eval-time dropout off in deterministic mode), fc8 = lib/ops.py:183-304
//   lib/architecture.py:386-389   tanh, mean over the 10 crops
// Layout: activations NHWC fp32 (as in the reference), conv weights HWIO (as stored by the reference's .npy / ckpt),
// fc weights transposed to [N, K] once so that the tensor-core GEMM reads both operands K-major.
// fc6-8 run on tcgen05 (gemm_tf32.cu); conv1-5 are fp32 implicit-GEMM kernels on the CUDA cores (tensor-core conv
// is a "next" row).
#include "common.cuh"

namespace hg {

int gemm_tf32(const float* A, int64_t lda, const float* Bt, int64_t ldb, const float* bias, float* C, int64_t ldc, int M, int N, int K, int relu,
              cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------
// K0-K4: normalize + scale + legacy bilinear resize to 256x256 + 10-crop (+flip) + mean subtraction, fused.
// in : uint8 [n, 3, wh, wh] (RGB planes, the loader's flattened layout, lib/dataloader.py:110-113)
// out: float [10n, 227, 227, 3]; crop block k holds rows k*n .. (k+1)*n (lib/architecture.py:242-244)
// ---------------------------------------------------------------------------------------------------------
// counter-based generator of the stochastic mode (splitmix64 finaliser of seed + index): the same function in
// hashgan_b200/encoder.py reproduces every draw on the host, so the stochastic path is testable against the oracle
__host__ __device__ __forceinline__ uint64_t hg_mix(uint64_t seed, uint64_t idx)
{
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
constexpr uint64_t kStreamNoise = 0xD1B54A32D192ED03ull, kStreamDrop6 = 0xA24BAED4963EE407ull, kStreamDrop7 = 0x9FB21C651E98DF25ull;

__device__ __forceinline__ float scaled_pixel(unsigned char v, float noise)
{
    const float x = 2.0f * (float)v / 256.0f - 1.0f + noise;  // main.py:146-147
    return (x + 1.0f) * 255.99f / 2.0f;                       // lib/util.py:13
}
// de-quantisation noise of main.py:147, tf.random_uniform([B, 3*wh*wh], 0, 1/128): one draw per source pixel (shared by the 10 crops)
__device__ __forceinline__ float pixel_noise(uint64_t seed, int64_t flat)
{
    if (seed == 0) return 0.0f;
    return (float)(hg_mix(seed ^ kStreamNoise, (uint64_t)flat) >> 40) * (1.0f / 16777216.0f) * (1.0f / 128.0f);
}

// tf.nn.dropout(x, 0.5) of lib/architecture.py:369,377 (active at eval in the reference: no stage guard): keep with p = 0.5, scale by 2
__global__ void __launch_bounds__(256) dropout_half_kernel(float* __restrict__ x, int64_t n, uint64_t seed)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = (hg_mix(seed, (uint64_t)i) >> 63) ? 2.0f * x[i] : 0.0f;
}

__global__ void __launch_bounds__(256) prep_crops_kernel(const unsigned char* __restrict__ img, int n, int wh, int cs, float* __restrict__ out,
                                                         uint64_t seed)
{
    const int64_t total = (int64_t)10 * n * 227 * 227;
    const float scale = (float)wh / 256.0f;  // tf.image.resize_bilinear, align_corners=False, no half-pixel centres
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % 227);
        const int y = (int)((i / 227) % 227);
        const int64_t nn = i / (227 * 227);
        const int k = (int)(nn / n), b = (int)(nn % n);
        const int kk = k % 5;
        const int oy = (kk == 1 || kk == 2) ? 28 : (kk == 4 ? 14 : 0);  // (0,0) (28,28) (28,0) (0,28) (14,14)
        const int ox = (kk == 1 || kk == 3) ? 28 : (kk == 4 ? 14 : 0);
        const int Y = oy + y;
        const int X = (k < 5) ? 255 - (ox + x) : ox + x;  // crops 0..4 are taken from the left-right flipped image
        const float fy = (float)Y * scale, fx = (float)X * scale;
        const int y0 = (int)floorf(fy), x0 = (int)floorf(fx);
        const int y1 = min((int)ceilf(fy), wh - 1), x1 = min((int)ceilf(fx), wh - 1);
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        const float mean[3] = {103.939f, 116.779f, 123.68f};
        float* o = out + i * cs;  // cs = 3, or 4 (4th channel zero) for the tensor-core conv1
        if (cs == 4) o[3] = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int64_t plane = ((int64_t)b * 3 + c) * wh * wh;
            const unsigned char* p = img + plane;
            const float tl = scaled_pixel(p[y0 * wh + x0], pixel_noise(seed, plane + y0 * wh + x0));
            const float tr = scaled_pixel(p[y0 * wh + x1], pixel_noise(seed, plane + y0 * wh + x1));
            const float bl = scaled_pixel(p[y1 * wh + x0], pixel_noise(seed, plane + y1 * wh + x0));
            const float br = scaled_pixel(p[y1 * wh + x1], pixel_noise(seed, plane + y1 * wh + x1));
            const float top = tl + (tr - tl) * lx;
            const float bot = bl + (br - bl) * lx;
            o[c] = (top + (bot - top) * ly) - mean[c];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K5/K8/K10-12: convolution as an implicit GEMM on the CUDA cores, fused bias + ReLU.
//   in  [N, H, W, C]  (NHWC), weights [KH, KW, C/groups, Cout] (HWIO), out [N, Ho, Wo, Cout]
//   tile: 64 output pixels x BN output channels of one group, K chunks of 16; 256 threads, 4 x (BN/16) outputs each.
// ---------------------------------------------------------------------------------------------------------
struct ConvParams {
    const float* in;
    const float* w;
    const float* bias;
    float* out;
    int N, H, W, C, KH, KW, stride, pad, Ho, Wo, Cout, groups;
};

template <int BN, bool VEC>
__global__ void __launch_bounds__(256) conv_relu_kernel(ConvParams p)
{
    constexpr int BM = 64, BK = 16, TN = BN / 16;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const int g = blockIdx.z;
    const int Cg = p.C / p.groups, Cog = p.Cout / p.groups;
    const int K = p.KH * p.KW * Cg;
    const int64_t M = (int64_t)p.N * p.Ho * p.Wo;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tm = tid / 16, tn = tid % 16;  // 16 x 16 threads: 4 pixels x TN channels each

    // A-tile loader: thread -> pixel (tid / 4), 4 consecutive k starting at (tid % 4) * 4
    const int a_m = tid >> 2, a_k = (tid & 3) * 4;
    const int64_t am = m0 + a_m;
    const bool a_valid = am < M;
    int a_n = 0, a_oy = 0, a_ox = 0;
    if (a_valid) {
        a_ox = (int)(am % p.Wo);
        a_oy = (int)((am / p.Wo) % p.Ho);
        a_n = (int)(am / ((int64_t)p.Wo * p.Ho));
    }
    const float* in_n = p.in + (int64_t)a_n * p.H * p.W * p.C + g * Cg;
    // B-tile loader: thread -> k row (tid / 16), 4 * (BN/64) ... handled as TN floats at column tn*TN
    const int b_k = tid >> 4, b_n = (tid & 15) * TN;

    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        // ---- gather A ----
        float av[4] = {0.f, 0.f, 0.f, 0.f};
        if (a_valid) {
            const int k = k0 + a_k;
            if (VEC) {  // Cg % 4 == 0: the 4 values share (ky, kx) and are contiguous in memory
                if (k < K) {
                    const int ci = k % Cg, kx = (k / Cg) % p.KW, ky = k / (Cg * p.KW);
                    const int iy = a_oy * p.stride + ky - p.pad, ix = a_ox * p.stride + kx - p.pad;
                    if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(in_n + ((int64_t)iy * p.W + ix) * p.C + ci));
                        av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
                    }
                }
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int kk = k + t;
                    if (kk < K) {
                        const int ci = kk % Cg, kx = (kk / Cg) % p.KW, ky = kk / (Cg * p.KW);
                        const int iy = a_oy * p.stride + ky - p.pad, ix = a_ox * p.stride + kx - p.pad;
                        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) av[t] = __ldg(in_n + ((int64_t)iy * p.W + ix) * p.C + ci);
                    }
                }
            }
        }
        // ---- load B ----
        float bv[TN];
        {
            const int k = k0 + b_k;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = n0 + b_n + j;
                bv[j] = (k < K && n < Cog) ? __ldg(p.w + (int64_t)k * p.Cout + g * Cog + n) : 0.0f;
            }
        }
        __syncthreads();  // previous chunk consumed
#pragma unroll
        for (int t = 0; t < 4; ++t) As[a_k + t][a_m] = av[t];
#pragma unroll
        for (int j = 0; j < TN; ++j) Bs[b_k][b_n + j] = bv[j];
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][tm * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float b[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tn * TN + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    // ---- epilogue: bias + ReLU, NHWC store ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + tm * 4 + i;
        if (m >= M) continue;
        float* o = p.out + m * p.Cout + g * Cog;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tn * TN + j;
            if (n < Cog) o[n] = fmaxf(acc[i][j] + __ldg(p.bias + g * Cog + n), 0.0f);
        }
    }
}

// K6/K9: 3x3 stride-2 VALID max pool, NHWC
__global__ void __launch_bounds__(256) maxpool3s2_kernel(const float* __restrict__ in, int N, int H, int W, int C, int Ho, int Wo, float* __restrict__ out)
{
    const int64_t total = (int64_t)N * Ho * Wo * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int ox = (int)((i / C) % Wo);
        const int oy = (int)((i / ((int64_t)C * Wo)) % Ho);
        const int64_t n = i / ((int64_t)C * Wo * Ho);
        const float* p = in + ((n * H + oy * 2) * W + ox * 2) * C + c;
        float m = -3.402823466e38f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) m = fmaxf(m, __ldg(p + ((int64_t)dy * W + dx) * C));
        out[i] = m;
    }
}

// K6 + K7 fused: 3x3 / 2 max pool followed by the LRN over the channel axis (lib/architecture.py:291-309).  One thread per
// channel (blockDim.x == C), a CTA walks pooled pixels: the pooled vector of a pixel goes through shared memory so that the
// LRN reads its four channel neighbours from there -- the pooled tensor never travels to HBM and back (pool2 + LRN2 of a
// 128-image batch: 0.93 ms as two kernels).  Same arithmetic as maxpool3s2_kernel + lrn_kernel.
__global__ void __launch_bounds__(256) maxpool3s2_lrn_kernel(const float* __restrict__ in, int N, int H, int W, int C, int Ho, int Wo, float* __restrict__ out)
{
    __shared__ float pooled[2][256 + 4];
    const int c = threadIdx.x;
    const int64_t pixels = (int64_t)N * Ho * Wo;
    int buf = 0;
    if (c < 2) { pooled[0][c] = pooled[1][c] = 0.0f; pooled[0][C + 2 + c] = pooled[1][C + 2 + c] = 0.0f; }  // zero halo: channels -2, -1, C, C+1
    for (int64_t px = blockIdx.x; px < pixels; px += gridDim.x, buf ^= 1) {
        const int ox = (int)(px % Wo);
        const int oy = (int)((px / Wo) % Ho);
        const int64_t n = px / ((int64_t)Wo * Ho);
        const float* p = in + ((n * H + oy * 2) * W + ox * 2) * C + c;
        float m = -3.402823466e38f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) m = fmaxf(m, __ldg(p + ((int64_t)dy * W + dx) * C));
        pooled[buf][c + 2] = m;
        __syncthreads();  // double buffered: the next pixel writes the other buffer, one barrier per pixel is enough
        float s = 0.0f;
#pragma unroll
        for (int d = 0; d <= 4; ++d) { const float v = pooled[buf][c + d]; s = fmaf(v, v, s); }
        out[px * C + c] = m * powf(1.0f + 2e-5f * s, -0.75f);
    }
}

// The same, one WARP per pooled pixel (C == 256: a lane owns 8 consecutive channels as two float4): 18 independent 16-byte loads
// per lane instead of nine 4-byte loads and a block barrier per pixel, the four channel neighbours of the LRN window come from the
// adjacent lanes by shuffle.  Same arithmetic, same summation order as maxpool3s2_lrn_kernel (bit-identical output).
__global__ void __launch_bounds__(256) maxpool3s2_lrn_c256_kernel(const float* __restrict__ in, int N, int H, int W, int Ho, int Wo, float* __restrict__ out)
{
    constexpr int C = 256;
    const int lane = threadIdx.x & 31;
    const int64_t pixels = (int64_t)N * Ho * Wo;
    const int64_t px = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (px >= pixels) return;
    const int ox = (int)(px % Wo);
    const int oy = (int)((px / Wo) % Ho);
    const int64_t n = px / ((int64_t)Wo * Ho);
    const float4* p = reinterpret_cast<const float4*>(in + ((n * H + oy * 2) * W + ox * 2) * C) + 2 * lane;
    float4 v[9][2];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const float4* q = p + ((int64_t)dy * W + dx) * (C / 4);
            v[dy * 3 + dx][0] = __ldg(q);
            v[dy * 3 + dx][1] = __ldg(q + 1);
        }
    float e[12];
#pragma unroll
    for (int j = 0; j < 8; ++j) e[2 + j] = -3.402823466e38f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        e[2] = fmaxf(e[2], v[t][0].x); e[3] = fmaxf(e[3], v[t][0].y); e[4] = fmaxf(e[4], v[t][0].z); e[5] = fmaxf(e[5], v[t][0].w);
        e[6] = fmaxf(e[6], v[t][1].x); e[7] = fmaxf(e[7], v[t][1].y); e[8] = fmaxf(e[8], v[t][1].z); e[9] = fmaxf(e[9], v[t][1].w);
    }
    const unsigned FULL = 0xffffffffu;
    const float u6 = __shfl_up_sync(FULL, e[8], 1), u7 = __shfl_up_sync(FULL, e[9], 1);
    const float d0 = __shfl_down_sync(FULL, e[2], 1), d1 = __shfl_down_sync(FULL, e[3], 1);
    e[0] = lane == 0 ? 0.0f : u6; e[1] = lane == 0 ? 0.0f : u7;      // channels -2, -1 do not exist
    e[10] = lane == 31 ? 0.0f : d0; e[11] = lane == 31 ? 0.0f : d1;  // channels C, C + 1
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float s = 0.0f;
#pragma unroll
        for (int d = 0; d <= 4; ++d) s = fmaf(e[j + d], e[j + d], s);
        r[j] = e[2 + j] * powf(1.0f + 2e-5f * s, -0.75f);
    }
    float4* o = reinterpret_cast<float4*>(out + px * C) + 2 * lane;
    o[0] = make_float4(r[0], r[1], r[2], r[3]);
    o[1] = make_float4(r[4], r[5], r[6], r[7]);
}

// 3x3 / stride 2 max pooling, four channels per thread (C % 4 == 0, 16-byte aligned tensors)
__global__ void __launch_bounds__(256) maxpool3s2_v4_kernel(const float* __restrict__ in, int N, int H, int W, int C, int Ho, int Wo, float* __restrict__ out)
{
    const int C4 = C >> 2;
    const int64_t total = (int64_t)N * Ho * Wo * C4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        const int64_t px = i / C4;
        const int ox = (int)(px % Wo);
        const int oy = (int)((px / Wo) % Ho);
        const int64_t n = px / ((int64_t)Wo * Ho);
        const float4* p = reinterpret_cast<const float4*>(in + ((n * H + oy * 2) * W + ox * 2) * C) + c4;
        float4 m = make_float4(-3.402823466e38f, -3.402823466e38f, -3.402823466e38f, -3.402823466e38f);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float4 v = __ldg(p + ((int64_t)dy * W + dx) * C4);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        reinterpret_cast<float4*>(out)[i] = m;
    }
}

// K7: tf.nn.local_response_normalization(depth_radius=2, bias=1, alpha=2e-5, beta=0.75) over the channel axis
__global__ void __launch_bounds__(256) lrn_kernel(const float* __restrict__ in, int64_t pixels, int C, float* __restrict__ out)
{
    const int64_t total = pixels * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const float* p = in + (i - c);
        float s = 0.0f;
#pragma unroll
        for (int d = -2; d <= 2; ++d) {
            const int cc = c + d;
            if (cc >= 0 && cc < C) { const float v = __ldg(p + cc); s = fmaf(v, v, s); }
        }
        out[i] = __ldg(p + c) * powf(1.0f + 2e-5f * s, -0.75f);
    }
}

// K16: tanh, then the mean over the 10 crop blocks (lib/architecture.py:386-389)
__global__ void __launch_bounds__(256) tanh_crop_mean_kernel(const float* __restrict__ fc8, int n, int b, float* __restrict__ out)
{
    const int total = n * b;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < 10; ++k) s += tanhf(__ldg(fc8 + (size_t)k * total + i));
        out[i] = s / 10.0f;
    }
}

__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out)
{
    __shared__ float tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int r = by + j, c = bx + tx;
        tile[j][tx] = (r < rows && c < cols) ? in[(size_t)r * cols + c] : 0.0f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = bx + j, r = by + tx;  // out[c][r]
        if (c < cols && r < rows) out[(size_t)c * rows + r] = tile[tx][j];
    }
}

// ---- tensor-core convolution: implicit GEMM (gemm_tf32.cu) -------------------------------------------------------------
// HWIO weights [KH*KW*Cg, Cout] -> per group K-major [groups][Cog][Kpad] (zero padded) for the GEMM's B operand
// (input channels per group are padded from Cg to Cgp -- conv1: 3 -> 4 -- so that every tap is a whole number of float4)
__global__ void __launch_bounds__(256) conv_weight_pack_kernel(const float* __restrict__ w, int taps, int Cg, int Cgp, int Cout, int Kpad,
                                                                float* __restrict__ out)
{
    const int64_t total = (int64_t)Cout * Kpad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % Kpad);
        const int co = (int)(i / Kpad);  // = g * Cog + n
        const int tap = k / Cgp, ci = k % Cgp;
        const float v = (tap < taps && ci < Cg) ? w[((int64_t)tap * Cg + ci) * Cout + co] : 0.0f;
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);  // exactly a TF32 number
        out[i] = hi;
        out[total + i] = v - hi;  // remainder, used by the error-compensated (3xTF32) convolution
    }
}

static bool env_flag(const char* name)
{
    const char* v = getenv(name);
    return v && *v && *v != '0';
}

static unsigned grid_1d(int64_t total, int threads)
{
    const int sms = device_facts().sm_count > 0 ? device_facts().sm_count : 148;
    const int64_t want = ceil_div(total, threads);
    const int64_t cap = (int64_t)sms * 16;
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, cap));
}

static int launch_conv(const float* in, const float* w, const float* bias, float* out, int N, int H, int W, int C, int KH, int KW, int stride,
                       int pad, int Cout, int groups, cudaStream_t st)
{
    ConvParams p;
    p.in = in; p.w = w; p.bias = bias; p.out = out; p.N = N; p.H = H; p.W = W; p.C = C; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad;
    p.Ho = (H + 2 * pad - KH) / stride + 1; p.Wo = (W + 2 * pad - KW) / stride + 1; p.Cout = Cout; p.groups = groups;
    const int Cog = Cout / groups, Cg = C / groups;
    const int64_t M = (int64_t)N * p.Ho * p.Wo;
    const bool vec = (Cg % 4 == 0) && (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
    const int BN = (Cog % 128 == 0) ? 128 : ((Cog % 96 == 0) ? 96 : 64);
    dim3 grid((unsigned)ceil_div(M, 64), (unsigned)ceil_div(Cog, BN), (unsigned)groups);
    if (BN == 128) { if (vec) conv_relu_kernel<128, true><<<grid, 256, 0, st>>>(p); else conv_relu_kernel<128, false><<<grid, 256, 0, st>>>(p); }
    else if (BN == 96) { if (vec) conv_relu_kernel<96, true><<<grid, 256, 0, st>>>(p); else conv_relu_kernel<96, false><<<grid, 256, 0, st>>>(p); }
    else { if (vec) conv_relu_kernel<64, true><<<grid, 256, 0, st>>>(p); else conv_relu_kernel<64, false><<<grid, 256, 0, st>>>(p); }
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

// encoder_stage1.cu
bool stage1_supported(int wh);
int stage1_launch(const unsigned char* img, int n, int wh, const float* wfused, const float* bias, float* out, uint64_t seed, bool lrn, cudaStream_t st);

static int conv_cgp(int Cg) { return (Cg + 3) & ~3; }  // channels per group as the tensor-core path lays them out
static int conv_kpad(int KH, int KW, int Cg) { return (int)round_up((int64_t)KH * KW * conv_cgp(Cg), 32); }

int conv_gemm_tf32(const float* in, const float* wt, const float* wt_lo, const float* bias, float* out, int64_t M, int H, int W, int C, int c0,
                   int Cg, int KH, int KW, int stride, int pad, int Ho, int Wo, int Kpad, int Cog, int ldc, cudaStream_t st, int relu = 1);  // gemm_tf32.cu
// convolution on the tensor cores: one implicit GEMM per channel group (+bias, ReLU) straight into the NHWC output;
// x3 = error-compensated TF32 (weights pre-split by hg_conv_weight_pack: [hi | lo])
static int launch_conv_tf32(const float* in, const float* wt, const float* bias, float* out, int N, int H, int W, int C, int KH, int KW, int stride,
                            int pad, int Cout, int groups, bool x3, cudaStream_t st)
{
    const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
    const int Cg = C / groups, Cog = Cout / groups;  // C is the stored (padded) channel count: conv1 reads 4-channel crops
    const int Kpad = conv_kpad(KH, KW, Cg);
    const int64_t M = (int64_t)N * Ho * Wo;
    if (M >= (int64_t(1) << 31)) return fail(HG_EINVAL, "conv_tf32: batch too large");
    const float* wt_lo = x3 ? wt + (size_t)Cout * Kpad : nullptr;
    for (int g = 0; g < groups; ++g) {
        int rc = conv_gemm_tf32(in, wt + (size_t)g * Cog * Kpad, wt_lo ? wt_lo + (size_t)g * Cog * Kpad : nullptr, bias + g * Cog, out + g * Cog, M, H, W,
                                C, g * Cg, Cg, KH, KW, stride, pad, Ho, Wo, Kpad, Cog, Cout, st);
        if (rc != HG_OK) return rc;
    }
    return HG_OK;
}

// bytes of ONE of the two ping-pong activation buffers: the largest activation, conv1's output [10n,55,55,96]
// (crops [10n,227,227,3] and conv2's output [10n,27,27,256] are smaller)
static size_t buf_bytes(int n) { return (((size_t)n * 10 * 55 * 55 * 96 * sizeof(float)) + 255) & ~size_t(255); }

}  // namespace hg


extern "C" size_t hg_alexnet_workspace_bytes(int n, unsigned flags)
{
    if (n <= 0) return 0;
    (void)flags;
    return 2 * hg::buf_bytes(n);
}

extern "C" int hg_conv_weight_pack(const float* d_w_hwio, int KH, int KW, int Cg, int Cout, int groups, float* d_out, void* stream)
{
    if (!d_w_hwio || !d_out || KH <= 0 || KW <= 0 || Cg <= 0 || Cout <= 0 || groups <= 0 || Cout % groups) return hg::fail(HG_EINVAL, "hg_conv_weight_pack: bad arguments");
    const int Kpad = hg::conv_kpad(KH, KW, Cg);
    hg::conv_weight_pack_kernel<<<hg::grid_1d((int64_t)Cout * Kpad, 256), 256, 0, (cudaStream_t)stream>>>(d_w_hwio, KH * KW, Cg, hg::conv_cgp(Cg), Cout, Kpad, d_out);
    hg::count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

extern "C" int hg_transpose_f32(const float* d_in, int rows, int cols, float* d_out, void* stream)
{
    if (rows <= 0 || cols <= 0 || !d_in || !d_out) return hg::fail(HG_EINVAL, "hg_transpose_f32: bad arguments");
    dim3 grid((unsigned)hg::ceil_div(cols, 32), (unsigned)hg::ceil_div(rows, 32));
    hg::transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, rows, cols, d_out);
    hg::count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    return HG_OK;
}

namespace hg {
// HG_ENC_TIMING: events at the stage boundaries of one encode call; stage k of the call is credited to a category
struct EncTimer {
    static constexpr int kMaxMarks = 40;
    cudaEvent_t ev[kMaxMarks] = {};
    int cat[kMaxMarks] = {};
    int n = 0;
    bool created = false, armed = false;
    int ensure()
    {
        if (!created) {
            for (int i = 0; i < kMaxMarks; ++i) HG_CUDA_TRY(cudaEventCreate(&ev[i]));
            created = true;
        }
        return HG_OK;
    }
    // everything enqueued since the previous mark belongs to category c (first call: c is ignored)
    void mark(int c, cudaStream_t st)
    {
        if (armed && n < kMaxMarks) { cat[n] = c; cudaEventRecord(ev[n], st); ++n; }
    }
};
static EncTimer& enc_timer()
{
    static thread_local EncTimer t;
    return t;
}
enum { kEncPrep = 0, kEncConv, kEncPool, kEncDense, kEncTail };
}  // namespace hg

extern "C" int hg_alexnet_phase_ms(float out[5])
{
    hg::EncTimer& t = hg::enc_timer();
    if (!out || !t.created || !t.armed || t.n < 2) return hg::fail(HG_EINVAL, "hg_alexnet_phase_ms: no timed hg_alexnet_encode call on this thread");
    HG_CUDA_TRY(cudaEventSynchronize(t.ev[t.n - 1]));
    for (int i = 0; i < 5; ++i) out[i] = 0.f;
    for (int i = 1; i < t.n; ++i) {
        float ms = 0.f;
        HG_CUDA_TRY(cudaEventElapsedTime(&ms, t.ev[i - 1], t.ev[i]));
        out[t.cat[i]] += ms;
    }
    return HG_OK;
}

static int alexnet_encode_impl(const uint8_t* d_images, int n, int wh, const HgAlexNetWeights* w, int hash_dim, unsigned flags, float* d_out,
                               void* d_workspace, size_t workspace_bytes, void* stream, uint64_t seed)
{
    using namespace hg;
    int rc0;
    if (n == 0) return HG_OK;
    if (n < 0 || wh <= 0 || wh > 256) return fail(HG_EINVAL, "hg_alexnet_encode: bad n=%d / wh=%d", n, wh);
    if (hash_dim <= 0 || hash_dim > 256) return fail(HG_EINVAL, "hg_alexnet_encode: unsupported HASH_DIM=%d (1..256)", hash_dim);
    if (!d_images || !w || !d_out || !d_workspace) return fail(HG_EINVAL, "hg_alexnet_encode: NULL pointer");
    if (workspace_bytes < hg_alexnet_workspace_bytes(n, flags)) return fail(HG_ENOMEM, "hg_alexnet_encode: workspace too small");
    for (int i = 0; i < 5; ++i)
        if (!w->conv_w[i] || !w->conv_b[i]) return fail(HG_EINVAL, "hg_alexnet_encode: conv%d weights missing", i + 1);
    if (!w->fc6_wt || !w->fc6_b || !w->fc7_wt || !w->fc7_b || !w->fc8_wt || !w->fc8_b) return fail(HG_EINVAL, "hg_alexnet_encode: fc weights missing");
    if (flags & ~(unsigned)(HG_ENC_LRN | HG_ENC_CONV_TF32 | HG_ENC_CONV_TF32X3 | HG_ENC_TIMING | HG_ENC_FUSED_STAGE1)) return fail(HG_EINVAL, "hg_alexnet_encode: unknown flag");
    const bool x3 = (flags & HG_ENC_CONV_TF32X3) != 0;
    const bool tc = x3 || (flags & HG_ENC_CONV_TF32) != 0;
    if (tc)
        for (int i = 0; i < 5; ++i)
            if (!w->conv_wt[i]) return fail(HG_EINVAL, "hg_alexnet_encode: HG_ENC_CONV_TF32 needs conv_wt[%d] (hg_conv_weight_pack)", i);
    cudaStream_t st = (cudaStream_t)stream;
    EncTimer& tm = enc_timer();
    tm.armed = false;
    tm.n = 0;
    if (flags & HG_ENC_TIMING) {
        if ((rc0 = tm.ensure()) != HG_OK) return rc0;
        tm.armed = true;
    }
    tm.mark(kEncPrep, st);
    const bool lrn = (flags & HG_ENC_LRN) != 0;
    float* A = static_cast<float*>(d_workspace);
    float* B = reinterpret_cast<float*>(static_cast<char*>(d_workspace) + buf_bytes(n));
    const int N = 10 * n;
    int rc;
    // one convolution layer: fp32 on the CUDA cores (default, parity with the fp32 oracle to ~1e-5) or TF32 on tcgen05
    auto conv = [&](int i, const float* src, float* dst, int H, int C, int KH, int stride, int pad, int Cout, int groups) -> int {
        return tc ? launch_conv_tf32(src, w->conv_wt[i], w->conv_b[i], dst, N, H, H, (C + 3) & ~3, KH, KH, stride, pad, Cout, groups, x3, st)
                  : launch_conv(src, w->conv_w[i], w->conv_b[i], dst, N, H, H, C, KH, KH, stride, pad, Cout, groups, st);
    };
    const bool fused1 = (flags & HG_ENC_FUSED_STAGE1) != 0;
    if (fused1 && (!w->conv1_fused || w->conv1_fused_wh != wh || !stage1_supported(wh)))
        return fail(HG_EINVAL, "hg_alexnet_encode: HG_ENC_FUSED_STAGE1 needs conv1_fused packed for wh=%d (hg_conv1_fused_pack; have wh=%d)", wh,
                    w->conv1_fused ? w->conv1_fused_wh : 0);
    float* cur = A;
    float* other = B;
    if (fused1) {
        // images -> pool1 (+LRN) output [N,27,27,96] in one kernel: no crop tensor, no conv1 output tensor
        if ((rc = stage1_launch(d_images, n, wh, w->conv1_fused, w->conv_b[0], A, seed, lrn, st)) != HG_OK) return rc;
        tm.mark(kEncPrep, st);  // the fused stage is reported in the slot of the kernels it replaces first
    } else {
        // crops -> A
        prep_crops_kernel<<<grid_1d((int64_t)N * 227 * 227, 256), 256, 0, st>>>(d_images, n, wh, tc ? 4 : 3, A, seed);
        count_launch();
        HG_CUDA_TRY(cudaGetLastError());
        tm.mark(kEncPrep, st);
        // conv1 11x11/4 VALID 3->96 : A -> B [N,55,55,96]
        if ((rc = conv(0, A, B, 227, 3, 11, 4, 0, 96, 1)) != HG_OK) return rc;
        tm.mark(kEncConv, st);
        // pool1 : B -> A [N,27,27,96]
        maxpool3s2_kernel<<<grid_1d((int64_t)N * 27 * 27 * 96, 256), 256, 0, st>>>(B, N, 55, 55, 96, 27, 27, A);
        count_launch();
        if (lrn) {
            lrn_kernel<<<grid_1d((int64_t)N * 27 * 27 * 96, 256), 256, 0, st>>>(cur, (int64_t)N * 27 * 27, 96, other);
            count_launch();
            std::swap(cur, other);
        }
    }
    HG_CUDA_TRY(cudaGetLastError());
    tm.mark(kEncPool, st);
    // conv2 5x5 SAME, 2 groups 48->128 : -> [N,27,27,256]
    if ((rc = conv(1, cur, other, 27, 96, 5, 1, 2, 256, 2)) != HG_OK) return rc;
    tm.mark(kEncConv, st);
    std::swap(cur, other);
    if (lrn) {  // pool2 + LRN2 in one kernel
        const int sms = device_facts().sm_count > 0 ? device_facts().sm_count : 148;
        if (env_flag("HG_POOL_LRN_BLOCK"))  // the block-per-pixel kernel (kept for comparison)
            maxpool3s2_lrn_kernel<<<(unsigned)std::min<int64_t>((int64_t)N * 13 * 13, (int64_t)sms * 16), 256, 0, st>>>(cur, N, 27, 27, 256, 13, 13, other);
        else
            maxpool3s2_lrn_c256_kernel<<<(unsigned)ceil_div((int64_t)N * 13 * 13, 8), 256, 0, st>>>(cur, N, 27, 27, 13, 13, other);
        count_launch();
        std::swap(cur, other);
    } else {
        maxpool3s2_kernel<<<grid_1d((int64_t)N * 13 * 13 * 256, 256), 256, 0, st>>>(cur, N, 27, 27, 256, 13, 13, other);
        count_launch();
        std::swap(cur, other);
    }
    HG_CUDA_TRY(cudaGetLastError());
    tm.mark(kEncPool, st);
    // conv3 3x3 SAME 256->384, conv4 3x3 SAME 2 groups 192->192, conv5 3x3 SAME 2 groups 192->128
    if ((rc = conv(2, cur, other, 13, 256, 3, 1, 1, 384, 1)) != HG_OK) return rc;
    std::swap(cur, other);
    if ((rc = conv(3, cur, other, 13, 384, 3, 1, 1, 384, 2)) != HG_OK) return rc;
    std::swap(cur, other);
    if ((rc = conv(4, cur, other, 13, 384, 3, 1, 1, 256, 2)) != HG_OK) return rc;
    tm.mark(kEncConv, st);
    std::swap(cur, other);
    // pool5 -> [N,6,6,256] == [N, 9216] in (h, w, c) order, the row order of the fc6 weights
    maxpool3s2_v4_kernel<<<grid_1d((int64_t)N * 6 * 6 * 64, 256), 256, 0, st>>>(cur, N, 13, 13, 256, 6, 6, other);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    tm.mark(kEncPool, st);
    std::swap(cur, other);
    // fc6, fc7 (+ReLU), fc8 on the tensor cores.  In the error-compensated mode a dense layer is the 1x1 "convolution"
    // of the implicit-GEMM kernel (H = W = 1, C = K): its producer warps split the activations into hi/lo on the fly.
    const bool fc3 = x3 && w->fc_wt3[0] && w->fc_wt3[1] && w->fc_wt3[2];
    auto dense = [&](const float* src, float* dst, const float* wt3, const float* wt, const float* bias, int K, int Nout, int relu) -> int {
        if (fc3) return conv_gemm_tf32(src, wt3, wt3 + (size_t)Nout * K, bias, dst, N, 1, 1, K, 0, K, 1, 1, 1, 0, 1, 1, K, Nout, Nout, st, relu);
        return gemm_tf32(src, K, wt, K, bias, dst, Nout, N, Nout, K, relu, st);
    };
    if ((rc = dense(cur, other, w->fc_wt3[0], w->fc6_wt, w->fc6_b, 9216, 4096, 1)) != HG_OK) return rc;
    std::swap(cur, other);
    if (seed) {
        dropout_half_kernel<<<grid_1d((int64_t)N * 4096, 256), 256, 0, st>>>(cur, (int64_t)N * 4096, seed ^ kStreamDrop6);
        count_launch();
    }
    if ((rc = dense(cur, other, w->fc_wt3[1], w->fc7_wt, w->fc7_b, 4096, 4096, 1)) != HG_OK) return rc;
    std::swap(cur, other);
    if (seed) {
        dropout_half_kernel<<<grid_1d((int64_t)N * 4096, 256), 256, 0, st>>>(cur, (int64_t)N * 4096, seed ^ kStreamDrop7);
        count_launch();
    }
    if ((rc = dense(cur, other, w->fc_wt3[2], w->fc8_wt, w->fc8_b, 4096, hash_dim, 0)) != HG_OK) return rc;
    std::swap(cur, other);
    tm.mark(kEncDense, st);
    tanh_crop_mean_kernel<<<grid_1d((int64_t)n * hash_dim, 256), 256, 0, st>>>(cur, n, hash_dim, d_out);
    count_launch();
    HG_CUDA_TRY(cudaGetLastError());
    tm.mark(kEncTail, st);
    return HG_OK;
}

extern "C" int hg_alexnet_encode(const uint8_t* d_images, int n, int wh, const HgAlexNetWeights* w, int hash_dim, unsigned flags, float* d_out,
                                 void* d_workspace, size_t workspace_bytes, void* stream)
{
    return alexnet_encode_impl(d_images, n, wh, w, hash_dim, flags, d_out, d_workspace, workspace_bytes, stream, 0);
}

extern "C" int hg_alexnet_encode_stochastic(const uint8_t* d_images, int n, int wh, const HgAlexNetWeights* w, int hash_dim, unsigned flags,
                                            float* d_out, void* d_workspace, size_t workspace_bytes, uint64_t seed, void* stream)
{
    if (seed == 0) return hg::fail(HG_EINVAL, "hg_alexnet_encode_stochastic: seed must be non-zero (0 is the deterministic mode)");
    return alexnet_encode_impl(d_images, n, wh, w, hash_dim, flags, d_out, d_workspace, workspace_bytes, stream, seed);
}
