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
                          unsigned long long *__restrict__ best, float *__restrict__ dmat) {
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
            if (dmat && t0 + t < T) dmat[(size_t)(t0 + t) * n_e + c0 + c] = d;      // top-k sampling needs every distance
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

// Exact phase of the tensor-core search.  One CTA per token: find the 32-code sub-tiles whose approximate minimum is within
// `slack` of the best one (the canonical arg-min is provably among them when slack >= 2 * max |d_approx - d_canonical|),
// re-evaluate their codes in the canonical fp32 order -- one warp per candidate sub-tile, lane = code -- take the first
// arg-min, gather the code vector.  (Round 1 kept one minimum per 128-code GEMM tile and re-evaluated 128 codes per
// candidate with one thread per code: 4x the exact distances and 128 KB of codebook per candidate through L1.)
constexpr int VQ_TILE = 128;                      // column tile of the distance GEMM
constexpr int VQ_SUB = 32;                        // codes per sub-tile minimum
__global__ void __launch_bounds__(VQ_TILE)
vq_refine_kernel(const float *__restrict__ z, const float *__restrict__ E, const float *__restrict__ zz, const float *__restrict__ ee,
                 const float *__restrict__ submin, float ee_max, float slack_rel, int n_sub, int D,
                 int64_t *__restrict__ idx, float *__restrict__ z_q, float *__restrict__ dmin) {
    extern __shared__ float zs[];                     // [D] token, then the candidate list
    __shared__ float red_f[4];
    __shared__ unsigned long long red_k[4];
    __shared__ int n_cand;
    int *cand = reinterpret_cast<int *>(zs + D);      // [n_sub]
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < D; k += VQ_TILE) zs[k] = z[(size_t)t * D + k];
    if (tid == 0) n_cand = 0;
    float m = INFINITY;
    for (int i = tid; i < n_sub; i += VQ_TILE) m = fminf(m, submin[(size_t)t * n_sub + i]);
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red_f[warp] = m;
    __syncthreads();
    m = fminf(fminf(red_f[0], red_f[1]), fminf(red_f[2], red_f[3]));
    const float zzt = zz[t];
    const float thr = m + slack_rel * (zzt + ee_max);
    // NaN-safe: a NaN / inf latent makes every minimum (and thr) NaN or inf; `!(x > thr)` then keeps EVERY sub-tile, so
    // such a row is re-evaluated exhaustively and gets exactly the canonical kernel's answer instead of an empty list
    for (int i = tid; i < n_sub; i += VQ_TILE)
        if (!(submin[(size_t)t * n_sub + i] > thr)) cand[atomicAdd(&n_cand, 1)] = i;
    __syncthreads();
    unsigned long long key = ~0ull;
    const int nc = n_cand;
    for (int ci = warp; ci < nc; ci += VQ_TILE / 32) {                 // the list order varies, the minimum over it does not
        const int c = cand[ci] * VQ_SUB + lane;
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
    unsigned e = (unsigned)(key & 0xffffffffull);
    if (e >= (unsigned)(n_sub * VQ_SUB)) e = 0;        // cannot happen with the NaN-safe candidate test; never index out of the codebook
    if (tid == 0) {
        idx[t] = (int64_t)e;
        if (dmin) dmin[t] = orderable_float((uint32_t)(key >> 32));
    }
    for (int k = tid; k < D; k += VQ_TILE) z_q[(size_t)t * D + k] = __ldg(E + (size_t)e * D + k);      // quantize.py:292 / :368
}

// counter-based uniform in [0,1): splitmix64 of (seed, token, sample)
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned t, unsigned s) {
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)t * 0x100000001B3ull + s + 1);
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (float)(x >> 40) * (1.0f / 16777216.0f);
}

// quantize.py:344-381 get_multiple_codewords for topk > 1: per token the k nearest codes (ascending distance, first
// index on ties), p = softmax(-d_topk), S multinomial draws with replacement, tokens whose down-sampled extrapolation
// mask is 0 pinned to the nearest code.  row0_probs = 1 reproduces the reference, which draws every token from the
// probabilities of token 0 (quantize.py:358).  One CTA per token; k <= 32.
constexpr int MAXK = 32;
__global__ void __launch_bounds__(256)
vq_topk_sample_kernel(const float *__restrict__ dmat, const float *__restrict__ E, const uint8_t *__restrict__ mask,
                      long long mask_bstride, int mask_w, int fy, int fx, int lat_w, int tokens_per_image, int n_e, int D, int K, int S,
                      unsigned long long seed, int row0_probs, int64_t *__restrict__ topk_idx, float *__restrict__ topk_p,
                      int64_t *__restrict__ idx, float *__restrict__ z_q) {
    __shared__ unsigned long long red[8];
    __shared__ unsigned long long chosen[MAXK];
    __shared__ float prob[MAXK];
    __shared__ int pick[64];
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *row = dmat + (size_t)t * n_e;
    unsigned long long last = 0;                                   // keys are > 0 for finite d >= -inf ... select strictly increasing keys
    for (int r = 0; r < K; ++r) {
        unsigned long long key = ~0ull;
        for (int c = tid; c < n_e; c += 256) {
            const unsigned long long kk = ((unsigned long long)float_orderable(row[c]) << 32) | (unsigned)c;
            if ((r == 0 || kk > last) && kk < key) key = kk;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
        }
        if (lane == 0) red[warp] = key;
        __syncthreads();
        key = red[0];
        for (int w = 1; w < 8; ++w) key = red[w] < key ? red[w] : key;
        if (tid == 0) chosen[r] = key;
        last = key;
        __syncthreads();
    }
    if (tid == 0) {                                                // softmax(-d) over the K candidates (quantize.py:354-355)
        const float d0 = orderable_float((uint32_t)(chosen[0] >> 32));
        float sum = 0.f;
        for (int r = 0; r < K; ++r) { prob[r] = expf(-(orderable_float((uint32_t)(chosen[r] >> 32)) - d0)); sum += prob[r]; }
        for (int r = 0; r < K; ++r) prob[r] /= sum;
    }
    __syncthreads();
    if (tid < K) {
        topk_idx[(size_t)t * K + tid] = (int64_t)(chosen[tid] & 0xffffffffull);
        topk_p[(size_t)t * K + tid] = prob[tid];
    }
    // the sampling distribution: own row, or row 0 of the same call (reference behaviour) -- row 0's probabilities are
    // produced by block 0; to stay single-pass every block recomputes them from row 0's distances when asked to.
    __shared__ float prob0[MAXK];
    if (row0_probs && t != 0) {
        const float *row0 = dmat;                                  // token 0
        unsigned long long l0 = 0;
        for (int r = 0; r < K; ++r) {
            unsigned long long key = ~0ull;
            for (int c = tid; c < n_e; c += 256) {
                const unsigned long long kk = ((unsigned long long)float_orderable(row0[c]) << 32) | (unsigned)c;
                if ((r == 0 || kk > l0) && kk < key) key = kk;
            }
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other < key ? other : key;
            }
            if (lane == 0) red[warp] = key;
            __syncthreads();
            key = red[0];
            for (int w = 1; w < 8; ++w) key = red[w] < key ? red[w] : key;
            if (tid == 0) prob0[r] = orderable_float((uint32_t)(key >> 32));
            l0 = key;
            __syncthreads();
        }
        if (tid == 0) {
            const float d0 = prob0[0];
            float sum = 0.f;
            for (int r = 0; r < K; ++r) { prob0[r] = expf(-(prob0[r] - d0)); sum += prob0[r]; }
            for (int r = 0; r < K; ++r) prob0[r] /= sum;
        }
    } else if (tid < K) {
        prob0[tid] = prob[tid];
    }
    __syncthreads();
    // mask down-sampling: nearest = top-left pixel of each block (quantize.py:345)
    const int img = t / tokens_per_image, tl = t - img * tokens_per_image;
    const int ly = tl / lat_w, lx = tl - ly * lat_w;
    const bool extrapolated = mask ? (mask[(size_t)img * mask_bstride + (size_t)(ly * fy) * mask_w + lx * fx] != 0) : true;
    if (tid < S) {
        int r = 0;
        if (extrapolated) {                                        // pinned tokens keep the nearest code (:364-367)
            const float u = uniform01(seed, (unsigned)t, (unsigned)tid);
            float c = 0.f;
            r = K - 1;
            for (int k = 0; k < K; ++k) { c += prob0[k]; if (u < c) { r = k; break; } }
        }
        pick[tid] = r;
        idx[(size_t)t * S + tid] = (int64_t)(chosen[r] & 0xffffffffull);
    }
    __syncthreads();
    for (int e = tid; e < S * (D / 4); e += 256) {
        const int sidx = e / (D / 4), k = e - sidx * (D / 4);
        const unsigned code = (unsigned)(chosen[pick[sidx]] & 0xffffffffull);
        reinterpret_cast<float4 *>(z_q + ((size_t)t * S + sidx) * D)[k] = __ldg(reinterpret_cast<const float4 *>(E + (size_t)code * D) + k);
    }
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

extern "C" size_t sgam_vq_tc_workspace_bytes(int T, int n_e) { return ((size_t)T * (n_e / VQ_SUB) + (size_t)T) * sizeof(float); }

extern "C" int sgam_vq_nearest_tc(const float *z, const void *z_hi, const void *z_lo, const float *codebook, const void *e_hi,
                                  const void *e_lo, const float *ee, float ee_max, int T, int n_e, int D, void *workspace,
                                  int64_t *idx, float *z_q, float *dmin, void *stream) {
    SGAM_REQUIRE(z && z_hi && z_lo && codebook && e_hi && e_lo && ee && workspace && idx && z_q, "vq_nearest_tc: null pointer");
    SGAM_REQUIRE(T > 0 && n_e > 0 && D % 64 == 0 && n_e % VQ_TILE == 0, "vq_nearest_tc: needs D %% 64 == 0 and n_e %% 128 == 0 (D=%d n_e=%d)", D, n_e);
    cudaStream_t s = (cudaStream_t)stream;
    float *tilemin = (float *)workspace, *zz = tilemin + (size_t)T * (n_e / VQ_SUB);
    vq_norms_kernel<<<cdiv(T, 128), 128, 0, s>>>(z, zz, T, D);
    SGAM_LAUNCH_OK();
    int rc = sgam_vq_tilemin_launch(z_hi, z_lo, e_hi, e_lo, zz, ee, tilemin, T, n_e, D, s);
    if (rc) return rc;
    // |d_tc - d_canonical| <= 2 * (2^-15 + 256 * 2^-24) * |z||e| <= 1e-4 * (|z|^2 + |e|^2)/2 ; slack = 2 * that bound
    const int n_sub = n_e / VQ_SUB;
    const size_t smem = (size_t)D * sizeof(float) + (size_t)n_sub * sizeof(int);
    vq_refine_kernel<<<T, VQ_TILE, smem, s>>>(z, codebook, zz, ee, tilemin, ee_max, 1e-4f, n_sub, D, idx, z_q, dmin);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}

extern "C" size_t sgam_vq_topk_workspace_bytes(int T, int n_e) { return (size_t)T * n_e * sizeof(float) + (size_t)T * sizeof(unsigned long long); }

extern "C" int sgam_vq_topk_sample(const float *z, const float *codebook, const uint8_t *mask, long long mask_bstride, int mask_w,
                                   int fy, int fx, int lat_w, int tokens_per_image, int T, int n_e, int D, int topk, int samples,
                                   unsigned long long seed, int row0_probs, void *workspace, int64_t *topk_idx, float *topk_p,
                                   int64_t *idx, float *z_q, void *stream) {
    SGAM_REQUIRE(z && codebook && workspace && topk_idx && topk_p && idx && z_q, "vq_topk_sample: null pointer");
    SGAM_REQUIRE(T > 0 && D % BK == 0 && n_e % TC == 0, "vq_topk_sample: needs D %% %d == 0 and n_e %% %d == 0", BK, TC);
    SGAM_REQUIRE(topk >= 1 && topk <= MAXK && topk <= n_e && samples >= 1 && samples <= 64, "vq_topk_sample: topk in [1,%d], samples in [1,64]", MAXK);
    cudaStream_t s = (cudaStream_t)stream;
    float *dmat = (float *)workspace;
    unsigned long long *best = (unsigned long long *)(dmat + (size_t)T * n_e);
    SGAM_CUDA_OK(cudaMemsetAsync(best, 0xff, (size_t)T * sizeof(unsigned long long), s));
    dim3 grid(n_e / TC, cdiv(T, TT));
    vq_distance_argmin_kernel<<<grid, NT, 0, s>>>(z, codebook, T, n_e, D, best, dmat);
    SGAM_LAUNCH_OK();
    vq_topk_sample_kernel<<<T, 256, 0, s>>>(dmat, codebook, mask, mask_bstride, mask_w, fy, fx, lat_w, tokens_per_image, n_e, D, topk,
                                            samples, seed, row0_probs, topk_idx, topk_p, idx, z_q);
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
    vq_distance_argmin_kernel<<<grid, NT, 0, s>>>(z, codebook, T, n_e, D, (unsigned long long *)best, nullptr);
    SGAM_LAUNCH_OK();
    vq_gather_kernel<<<T, 64, 0, s>>>((const unsigned long long *)best, codebook, D, idx, z_q, dmin);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}
