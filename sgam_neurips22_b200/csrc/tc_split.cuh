// fp32 pair -> packed split-bf16 (hi, lo) words, shared by the GEMM epilogue and the operand producers.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

// hi = bf16(x) (round to nearest even), lo = bf16(x - hi); element a in the low half-word, b in the high one.
// One packed conversion per pair (cvt.rn.bf16x2.f32 -> F2FP.BF16.F32.PACK_AB) and integer re-expansion of hi: 6
// instructions per pair instead of 10 with scalar conversions + byte permutes; the results are bit-identical.
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - ah, b - bh);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
