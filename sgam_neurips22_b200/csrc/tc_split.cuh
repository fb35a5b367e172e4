// fp32 pair -> packed split-bf16 (hi, lo) words, shared by the GEMM epilogue and the operand producers.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(a), h1 = __float2bfloat16_rn(b);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(a - __bfloat162float(h0)), l1 = __float2bfloat16_rn(b - __bfloat162float(h1));
    hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}

