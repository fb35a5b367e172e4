// Stage (ii): VectorQuantizer2 nearest neighbour (quantize.py:275-319, 344-381 with topk=1).
//
// d[t,e] = (|z_t|^2 + |e|^2) - 2 z_t.e ; idx[t] = first arg-min ; z_q[t] = E[idx[t]].
// Canonical arithmetic shared bit-for-bit with oracle_vq_nearest(): every dot product and squared norm
// is ONE sequential fp32 fma chain over k = 0..D-1, so indices (including exact ties) are reproducible.
//
// Kernel: 64 tokens x 128 codes per CTA, codebook / latent tiles staged k-major in shared memory
// (32-wide k chunks), 4x8 register micro-tile per thread, per-token arg-min by 16-lane shuffle reduction,
// one 64-bit atomicMin per (token, CTA) on a packed (orderable distance, index) key.
#include "common.cuh"

namespace {

constexpr int TT = 64, TC = 128, BK = 32, NT = 256;
constexpr int ZS = TT + 4, ES = TC + 4;   // padded row strides (floats), keep 16-byte alignment

__global__ void __launch_bounds__(NT)
vq_distance_argmin_kernel(const float *__restrict__ z, const float *__restrict__ E, int T, int n_e, int D,
                          unsigned long long *__restrict__ best) {
    __shared__ __align__(16) float Zs[BK][ZS];
    __shared__ __align__(16) float Es[BK][ES];
    __shared__ float zz_s[TT], ee_s[TC];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int c0 = blockIdx.x * TC, t0 = blockIdx.y * TT;
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.0f;
    float nrm = 0.0f;
    // norm ownership: thread (ty<8, tx) owns code  cn = (ty>>2)*64 + tx*4 + (ty&3)
    //                 thread (ty>=8, tx<8... ) owns token tn = (ty-8)*8 + tx   (tx < 8)
    const bool own_e = ty < 8, own_z = (ty >= 8) && (tx < 8);
    const int cn = (ty >> 2) * 64 + tx * 4 + (ty & 3), tn = (ty - 8) * 8 + tx;

    for (int k0 = 0; k0 < D; k0 += BK) {
        // stage tiles (k-major): 64x32 latent floats = 512 float4, 128x32 codebook floats = 1024 float4
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int f = tid + r * NT, row = f >> 3, kq = f & 7;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t0 + row < T) v = __ldg(reinterpret_cast<const float4 *>(z + (size_t)(t0 + row) * D + k0) + kq);
            Zs[kq * 4 + 0][row] = v.x; Zs[kq * 4 + 1][row] = v.y; Zs[kq * 4 + 2][row] = v.z; Zs[kq * 4 + 3][row] = v.w;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int f = tid + r * NT, row = f >> 3, kq = f & 7;
            float4 v = __ldg(reinterpret_cast<const float4 *>(E + (size_t)(c0 + row) * D + k0) + kq);
            Es[kq * 4 + 0][row] = v.x; Es[kq * 4 + 1][row] = v.y; Es[kq * 4 + 2][row] = v.z; Es[kq * 4 + 3][row] = v.w;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < BK; ++k) {
            const float4 zv = *reinterpret_cast<const float4 *>(&Zs[k][ty * 4]);
            const float4 e0 = *reinterpret_cast<const float4 *>(&Es[k][tx * 4]);
            const float4 e1 = *reinterpret_cast<const float4 *>(&Es[k][64 + tx * 4]);
            const float zr[4] = {zv.x, zv.y, zv.z, zv.w};
            const float er[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = __fmaf_rn(zr[a], er[b], acc[a][b]);
            if (own_e) { const float v = Es[k][cn]; nrm = __fmaf_rn(v, v, nrm); }
            else if (own_z) { const float v = Zs[k][tn]; nrm = __fmaf_rn(v, v, nrm); }
        }
        __syncthreads();
    }
    if (own_e) ee_s[cn] = nrm;
    else if (own_z) zz_s[tn] = nrm;
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int t = ty * 4 + a;
        const float zz = zz_s[t];
        unsigned long long key = ~0ull;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int c = (b >> 2) * 64 + tx * 4 + (b & 3);
            const float d = __fadd_rn(__fsub_rn(__fadd_rn(zz, ee_s[c]), __fmul_rn(2.0f, acc[a][b])), 0.0f);   // quantize.py:285-287 (+0: -0 -> +0)
            const unsigned long long kk = ((unsigned long long)float_orderable(d) << 32) | (unsigned)(c0 + c);
            key = kk < key ? kk : key;
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
        }
        if (tx == 0 && t0 + t < T) atomicMin(best + t0 + t, key);
    }
}

__global__ void __launch_bounds__(64)
vq_gather_kernel(const unsigned long long *__restrict__ best, const float *__restrict__ E, int D,
                 int64_t *__restrict__ idx, float *__restrict__ z_q, float *__restrict__ dmin) {
    const int t = blockIdx.x;
    const unsigned long long key = best[t];
    const unsigned e = (unsigned)(key & 0xffffffffull);
    if (threadIdx.x == 0) {
        idx[t] = (int64_t)e;
        if (dmin) dmin[t] = orderable_float((uint32_t)(key >> 32));
    }
    const float4 *src = reinterpret_cast<const float4 *>(E + (size_t)e * D);
    float4 *dst = reinterpret_cast<float4 *>(z_q + (size_t)t * D);
    for (int k = threadIdx.x; k < D / 4; k += 64) dst[k] = __ldg(src + k);    // quantize.py:292 / :368
}

}  // namespace

extern "C" size_t sgam_vq_workspace_bytes(int T) { return (size_t)T * sizeof(unsigned long long); }

extern "C" int sgam_vq_nearest(const float *z, const float *codebook, int T, int n_e, int D, void *best, int64_t *idx,
                               float *z_q, float *dmin, void *stream) {
    SGAM_REQUIRE(z && codebook && best && idx && z_q, "vq_nearest: null pointer");
    SGAM_REQUIRE(T > 0 && n_e > 0 && D > 0, "vq_nearest: bad shape T=%d n_e=%d D=%d", T, n_e, D);
    SGAM_REQUIRE(D % BK == 0 && n_e % TC == 0, "vq_nearest: needs D %% %d == 0 and n_e %% %d == 0 (got D=%d n_e=%d)", BK, TC, D, n_e);
    cudaStream_t s = (cudaStream_t)stream;
    SGAM_CUDA_OK(cudaMemsetAsync(best, 0xff, sgam_vq_workspace_bytes(T), s));
    dim3 grid(n_e / TC, cdiv(T, TT));
    vq_distance_argmin_kernel<<<grid, NT, 0, s>>>(z, codebook, T, n_e, D, (unsigned long long *)best);
    SGAM_LAUNCH_OK();
    vq_gather_kernel<<<T, 64, 0, s>>>((const unsigned long long *)best, codebook, D, idx, z_q, dmin);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}
