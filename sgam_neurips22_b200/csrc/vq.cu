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

// canonical squared norms: one sequential fma chain per row (the oracle's order)
__global__ void __launch_bounds__(128)
vq_norms_kernel(const float *__restrict__ x, float *__restrict__ out, int rows, int D) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float4 *src = reinterpret_cast<const float4 *>(x + (size_t)r * D);
    float s = 0.0f;
    for (int k = 0; k < D / 4; ++k) {
        const float4 v = __ldg(src + k);
        s = __fmaf_rn(v.x, v.x, s); s = __fmaf_rn(v.y, v.y, s); s = __fmaf_rn(v.z, v.z, s); s = __fmaf_rn(v.w, v.w, s);
    }
    out[r] = s;
}

// Exact phase of the tensor-core search.  One CTA per token: find the tiles whose approximate minimum is within
// `slack` of the best one (the canonical arg-min is provably among them when slack >= 2 * max |d_approx - d_canonical|),
// re-evaluate their 128 codes each in the canonical fp32 order, take the first arg-min, gather the code vector.
constexpr int VQ_TILE = 128;
__global__ void __launch_bounds__(VQ_TILE)
vq_refine_kernel(const float *__restrict__ z, const float *__restrict__ E, const float *__restrict__ zz, const float *__restrict__ ee,
                 const float *__restrict__ tilemin, float ee_max, float slack_rel, int n_tiles, int D,
                 int64_t *__restrict__ idx, float *__restrict__ z_q, float *__restrict__ dmin) {
    extern __shared__ float zs[];                     // [D] token, then the candidate list
    __shared__ float red_f[4];
    __shared__ unsigned long long red_k[4];
    __shared__ int n_cand;
    int *cand = reinterpret_cast<int *>(zs + D);      // [n_tiles]
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < D; k += VQ_TILE) zs[k] = z[(size_t)t * D + k];
    if (tid == 0) n_cand = 0;
    float m = INFINITY;
    for (int i = tid; i < n_tiles; i += VQ_TILE) m = fminf(m, tilemin[(size_t)t * n_tiles + i]);
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red_f[warp] = m;
    __syncthreads();
    m = fminf(fminf(red_f[0], red_f[1]), fminf(red_f[2], red_f[3]));
    const float zzt = zz[t];
    const float thr = m + slack_rel * (zzt + ee_max);
    for (int i = tid; i < n_tiles; i += VQ_TILE)
        if (tilemin[(size_t)t * n_tiles + i] <= thr) cand[atomicAdd(&n_cand, 1)] = i;
    __syncthreads();
    unsigned long long key = ~0ull;
    const int nc = n_cand;
    for (int ci = 0; ci < nc; ++ci) {
        const int c = cand[ci] * VQ_TILE + tid;
        const float4 *ev = reinterpret_cast<const float4 *>(E + (size_t)c * D);
        float dot = 0.0f;
        for (int k = 0; k < D / 4; ++k) {
            const float4 e4 = __ldg(ev + k);
            dot = __fmaf_rn(zs[4 * k], e4.x, dot); dot = __fmaf_rn(zs[4 * k + 1], e4.y, dot);
            dot = __fmaf_rn(zs[4 * k + 2], e4.z, dot); dot = __fmaf_rn(zs[4 * k + 3], e4.w, dot);
        }
        const float d = __fadd_rn(__fsub_rn(__fadd_rn(zzt, ee[c]), __fmul_rn(2.0f, dot)), 0.0f);   // quantize.py:285-287
        const unsigned long long kk = ((unsigned long long)float_orderable(d) << 32) | (unsigned)c;
        key = kk < key ? kk : key;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other < key ? other : key;
    }
    if (lane == 0) red_k[warp] = key;
    __syncthreads();
    key = red_k[0];
    for (int w = 1; w < 4; ++w) key = red_k[w] < key ? red_k[w] : key;
    const unsigned e = (unsigned)(key & 0xffffffffull);
    if (tid == 0) {
        idx[t] = (int64_t)e;
        if (dmin) dmin[t] = orderable_float((uint32_t)(key >> 32));
    }
    for (int k = tid; k < D; k += VQ_TILE) z_q[(size_t)t * D + k] = __ldg(E + (size_t)e * D + k);      // quantize.py:292 / :368
}

}  // namespace

int sgam_vq_tilemin_launch(const void *z_hi, const void *z_lo, const void *e_hi, const void *e_lo, const float *zz, const float *ee,
                           float *tilemin, int T, int n_e, int D, cudaStream_t stream);     // net_tc.cu

extern "C" int sgam_vq_norms(const float *x, float *out, int rows, int D, void *stream) {
    SGAM_REQUIRE(x && out && rows > 0 && D > 0 && D % 4 == 0, "vq_norms: bad arguments");
    vq_norms_kernel<<<cdiv(rows, 128), 128, 0, (cudaStream_t)stream>>>(x, out, rows, D);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}

extern "C" size_t sgam_vq_tc_workspace_bytes(int T, int n_e) { return ((size_t)T * (n_e / VQ_TILE) + (size_t)T) * sizeof(float); }

extern "C" int sgam_vq_nearest_tc(const float *z, const void *z_hi, const void *z_lo, const float *codebook, const void *e_hi,
                                  const void *e_lo, const float *ee, float ee_max, int T, int n_e, int D, void *workspace,
                                  int64_t *idx, float *z_q, float *dmin, void *stream) {
    SGAM_REQUIRE(z && z_hi && z_lo && codebook && e_hi && e_lo && ee && workspace && idx && z_q, "vq_nearest_tc: null pointer");
    SGAM_REQUIRE(T > 0 && n_e > 0 && D % 64 == 0 && n_e % VQ_TILE == 0, "vq_nearest_tc: needs D %% 64 == 0 and n_e %% 128 == 0 (D=%d n_e=%d)", D, n_e);
    cudaStream_t s = (cudaStream_t)stream;
    float *tilemin = (float *)workspace, *zz = tilemin + (size_t)T * (n_e / VQ_TILE);
    vq_norms_kernel<<<cdiv(T, 128), 128, 0, s>>>(z, zz, T, D);
    SGAM_LAUNCH_OK();
    int rc = sgam_vq_tilemin_launch(z_hi, z_lo, e_hi, e_lo, zz, ee, tilemin, T, n_e, D, s);
    if (rc) return rc;
    // |d_tc - d_canonical| <= 2 * (2^-15 + 256 * 2^-24) * |z||e| <= 1e-4 * (|z|^2 + |e|^2)/2 ; slack = 2 * that bound
    const int n_tiles = n_e / VQ_TILE;
    const size_t smem = (size_t)D * sizeof(float) + (size_t)n_tiles * sizeof(int);
    vq_refine_kernel<<<T, VQ_TILE, smem, s>>>(z, codebook, zz, ee, tilemin, ee_max, 1e-4f, n_tiles, D, idx, z_q, dmin);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}

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
