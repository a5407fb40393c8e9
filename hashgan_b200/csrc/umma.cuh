// tcgen05 / TMA helpers shared by the tensor-core kernels (gemm_tf32.cu, select_umma.cu).
#pragma once
#include "common.cuh"

#include <cuda.h>

namespace hg {

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t mbar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(mbar), "r"(c0), "r"(c1)
                 : "memory");
}

// shared-memory matrix descriptor: K-major operand, 128-byte rows, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell), bits [46,48)
    d |= (uint64_t)2 << 61;                         // layout type SWIZZLE_128B, bits [61,64)
    return d;
}

__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// shared-memory matrix descriptor of a K-major operand whose rows are `row_bytes` (32, 64 or 128) wide and swizzled
// with the matching mode; 8-row groups are 8*row_bytes apart
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, int row_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6)) << 61;  // SWIZZLE_128B = 2, SWIZZLE_64B = 4, SWIZZLE_32B = 6
    return d;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();  // cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no link-time dependency on libcuda)

}  // namespace hg
