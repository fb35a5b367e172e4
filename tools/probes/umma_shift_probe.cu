// Probe: can a K-major SWIZZLE_128B UMMA operand start at a row that is NOT a multiple of 8 (start address not 1024-byte
// aligned)?  A 3x3 convolution's three horizontal taps are the same pixel rows shifted by one 128-byte row; if the descriptor
// can express the shift, one TMA load serves three taps.  Variants: base_offset field (bits 49..51) = 0, or = (addr >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I sgam_neurips22_b200/csrc -o umma_shift_probe tools/probes/umma_shift_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "tc_common.cuh"

using namespace tc;

constexpr int M = 128, N = 64, K = 64, ROWS_B = N + 8;

__device__ __forceinline__ uint32_t swz(int row, int col) {      // byte offset of bf16 element (row, col) in a 1024-B aligned SW128 tile
    const int chunk = (col * 2) >> 4;
    return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4) + ((col * 2) & 15));
}

__global__ void __launch_bounds__(128, 1)
probe(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D, int shift, int use_base_offset) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sa = smem, *sb = smem + 16384;
    __shared__ __align__(8) uint64_t done_bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < M * K; e += 128) { const int r = e / K, c = e % K; *(__nv_bfloat16 *)(sa + swz(r, c)) = A[e]; }
    for (int e = tid; e < ROWS_B * K; e += 128) { const int r = e / K, c = e % K; *(__nv_bfloat16 *)(sb + swz(r, c)) = B[e]; }
    if (tid == 0) { mbar_init(&done_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_smem;
    if (tid == 0) {
        constexpr uint32_t idesc = make_idesc(M, N);
        const uint8_t *bstart = sb + shift * 128;
        uint64_t adesc = make_smem_desc<128>(sa), bdesc = make_smem_desc<128>(bstart);
        if (use_base_offset) bdesc |= (uint64_t)((smem_u32(bstart) >> 7) & 7) << 49;
        for (int k = 0; k < K / 16; ++k) umma_bf16(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
        umma_commit(&done_bar);
    }
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64) : "memory"); }
}

int main() {
    std::vector<__nv_bfloat16> hA(M * K), hB(ROWS_B * K);
    std::vector<float> fA(M * K), fB(ROWS_B * K);
    srand(1);
    for (int i = 0; i < M * K; ++i) { fA[i] = (float)(rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(fA[i]); }
    for (int i = 0; i < ROWS_B * K; ++i) { fB[i] = (float)(rand() % 17 - 8) / 8.0f; hB[i] = __float2bfloat16(fB[i]); }
    __nv_bfloat16 *dA, *dB; float *dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
    std::vector<float> hD(M * N);
    for (int ubo = 0; ubo < 2; ++ubo)
        for (int shift = 0; shift < 8; ++shift) {
            cudaMemset(dD, 0, M * N * 4);
            probe<<<1, 128, 40960>>>(dA, dB, dD, shift, ubo);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("base_offset=%d shift=%d: CUDA error %s\n", ubo, shift, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0.0; int bad = 0;
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0.0;
                    for (int k = 0; k < K; ++k) ref += (double)fA[m * K + k] * (double)fB[(n + shift) * K + k];
                    const double err = fabs(ref - (double)hD[m * N + n]);
                    if (err > maxerr) maxerr = err;
                    if (err > 1e-3) ++bad;
                }
            printf("base_offset field %s  row shift %d: max |err| %.3g  wrong elements %d / %d  %s\n", ubo ? "set " : "zero", shift, maxerr, bad, M * N,
                   bad ? "MISMATCH" : "ok");
        }
    return 0;
}
