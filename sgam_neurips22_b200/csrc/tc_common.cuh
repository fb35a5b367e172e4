// Shared by the 1-CTA (net_tc.cu) and 2-CTA (net_tc2.cu) tcgen05 GEMM kernels: PTX wrappers (mbarrier, TMA, tcgen05),
// UMMA descriptors, the parameter block and the epilogue helpers.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_split.cuh"

namespace tc {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done;
}
// Bounded wait: a protocol bug (a missing arrive, a wrong parity, a tx-count that never completes) used to be a silent
// hang of the whole GPU box.  After SGAM_MBAR_TIMEOUT_NS of wall time (%globaltimer, sampled every 4096 polls, so the
// fast path costs nothing) the waiter reports which barrier / parity / role was stuck and traps: the launch fails with
// cudaErrorLaunchFailure and the host sees an error instead of a time-out.
#ifndef SGAM_MBAR_TIMEOUT_NS
#define SGAM_MBAR_TIMEOUT_NS 4000000000ull
#endif
static __device__ __noinline__ void mbar_timeout_report(uint32_t bar_addr, uint32_t parity) {
    printf("sgam: mbarrier wait timed out: block %d thread %d (warp %d) barrier smem 0x%x parity %u\n", (int)blockIdx.x,
           (int)threadIdx.x, (int)(threadIdx.x >> 5), bar_addr, parity);
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        if (mbar_try_wait(bar, parity)) return;
        if ((spins & 0xfffu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > SGAM_MBAR_TIMEOUT_NS) mbar_timeout_report(smem_u32(bar), parity);
        }
    }
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int NTHREADS = 128>
__device__ __forceinline__ void epilogue_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 in, fp32 accumulate), single CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (= 1, unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (8 rows of ROW_BYTES) | [46,48) version = 1
//   [61,64) layout: SWIZZLE_128B = 2 (128-byte rows), SWIZZLE_64B = 4 (64-byte rows)
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t make_smem_desc(const void *tile) {
    static_assert(ROW_BYTES == 128 || ROW_BYTES == 64, "K-major rows are one swizzle span");
    const uint64_t addr = (uint64_t)((smem_u32(tile) & 0x3FFFFu) >> 4);
    return addr | (1ull << 16) | ((uint64_t)((8 * ROW_BYTES) >> 4) << 32) | (1ull << 46) | ((ROW_BYTES == 128 ? 2ull : 4ull) << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b format BF16 (1) @7/@10,
// K-major A and B (0) @15/@16, N >> 3 @17, M >> 4 @24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ the GEMM kernel
constexpr int TC_THREADS = 192;

struct TcParams {
    // tiling of M: m-tile -> (b, ty, tx); rows of the tile are (ly, lx) with lx < BW, ly < BH, BW*BH = 128*MT
    int tiles_x, tiles_y, BW, BH;
    int tiles_m, tiles_n;       // persistent schedule: tile id = m_tile * tiles_n + n_tile (the n-tiles of one pixel block
                                // run on neighbouring SMs at the same time, so the A box is fetched from HBM once)
    int Ho, Wo;                 // output pixel grid per batch element (plain GEMM: Ho = 1, Wo = M)
    int taps, ks, pad, stride;  // conv: ks*ks taps, left/top pad, stride; plain GEMM: taps = 1, ks = 1, pad = 0, stride = 1
    int kblocks_per_tap;        // ceil(C / BK)
    int N;                      // total output columns (row stride of D)
    int n_valid;                // columns actually stored (< N only for the zero-padded small-Cout head)
    int out_nchw;               // store D as [b][n][oy][ox] (decoder conv_out) instead of row-major [m][n]
    int nsplit;                 // 3 = hi*hi + hi*lo + lo*hi ; 1 = hi*hi only
    int a_batched, b_batched;   // whether the operand has a batch dimension (else coordinate 0)
    long long d_batch_stride;   // elements between batch slices of D / R
    float alpha;
    const float *bias_n, *bias_m, *R;
    float *D;                   // fp32 output (or null)
    __nv_bfloat16 *D_hi, *D_lo; // split-bf16 output (or null)
    float *stats;               // or null: GroupNorm partial sums per (batch, 128-pixel block) [B][blocks][32][2]
    int cpg;                    // channels per group = N / 32 when stats != null
    // VQ mode (vq_tilemin != null): no D; the epilogue forms d = (zz[m] + ee[n]) - 2*acc and keeps the per-row minimum
    // over the tile's columns: vq_tilemin[m * tiles_n + n_tile]
    const float *vq_zz, *vq_ee;
    float *vq_tilemin;
    // split-K (ksplit > 1): tile id gains a k-split index (fastest); each split accumulates kb_per_split k-blocks and
    // stores its raw fp32 partial tile to splitk_ws[split][B*M][N]; splitk_reduce_kernel adds them in split order.
    int ksplit, kb_per_split;
    float *splitk_ws;
    long long split_stride;
    // Sub-pixel form of Upsample (nearest x2, then 3x3 conv; diffusionmodules/model.py:49-52): output parity (py, px) is a
    // 2x2 convolution of the LOW-resolution tensor with pre-summed taps.  shift_x / shift_y move the tap window
    // (parity 0 reads rows {i-1, i}, parity 1 rows {i, i+1}); up = 1 scatters tile pixel (oy, ox) to output pixel
    // (2 oy + py, 2 ox + px) of the 2Ho x 2Wo tensor; the GroupNorm partial sums of the four parity launches share one
    // buffer: stat_tiles slots per image (0 = this launch's own count), this launch's first slot = stat_tile0.
    int shift_x, shift_y, up, py, px, stat_tiles, stat_tile0;
    // Fused q / k / v projection of an AttnBlock (diffusionmodules/model.py:158-175): one GEMM with N = 3C columns.  Columns
    // below vt_col0 (q | k) go row-major to D_hi / D_lo with row pitch ldd (0 = N); columns from vt_col0 on (v) are stored
    // TRANSPOSED as split bf16 V^T [b][column - vt_col0][vt_T tokens] -- the K-major B operand of the P.V product.
    int ldd, vt_col0;
    long long vt_T;
    __nv_bfloat16 *vt_hi, *vt_lo;
};

// per-warp GroupNorm partial sums of one 32-column chunk: CPG channels per group, rows = lanes
template <int CPG>
__device__ __forceinline__ void chunk_group_sums(const float (&o)[32], bool row_ok, int lane, float *dst /* [32/CPG][2] */) {
    constexpr int G = 32 / CPG;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int j = 0; j < CPG; ++j) { const float v = row_ok ? o[g * CPG + j] : 0.f; s += v; q = fmaf(v, v, q); }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, off); q += __shfl_xor_sync(0xffffffffu, q, off); }
        if (lane == 0) { dst[2 * g] = s; dst[2 * g + 1] = q; }
    }
}

// Per-warp staging for coalesced output stores: 32 rows x 32 fp32 columns, rows padded to 36 floats (16-byte aligned,
// conflict-free for 128-bit accesses by quarter-warps), plus the global element offset of each of the warp's rows.
constexpr int STG_STRIDE = 36;
template <int EW>
struct EpiStageT {
    float tile[EW][32 * STG_STRIDE];
    long long rowoff[EW][32];               // row_off of the warp's row i, or -1 when the row is outside the image
};
using EpiStage = EpiStageT<4>;

// Epilogue of one 128-row accumulator: thread (q, lane) owns box row r (= TMEM lane 32q + lane) and walks the BN
// columns in 32-wide chunks: VQ tile minimum, split-K partial store, or alpha / bias / residual + fp32 / split-bf16 /
// NCHW stores + fused GroupNorm partial statistics.  Must be called by all four epilogue warps (named barrier 1).
// A TMEM lane is an output ROW, so a thread holds 32 consecutive columns of its own row; storing them directly makes
// every warp-wide store touch 32 rows x 16 bytes (32 half-used sectors; the LSU, not HBM, then bounds the small-K
// GEMMs).  The finished chunk therefore goes through the warp's staging tile and is written back row-contiguously:
// 8 lanes x 16 B = one 128-byte line of a row (fp32), 4 lanes x 16 B = one 64-byte segment (bf16 planes).
// EW epilogue warps (4 or 8): with 8, warps w and w + 4 share a TMEM lane quadrant (the same 32 rows) and take alternate 32-column
// chunks -- a single warp per scheduler runs the ~400 dependent instructions of a chunk at a fraction of the issue rate, which
// bounds the short-K (1x1) layers by their epilogue.  q below is the warp's index w in 0 .. EW-1 (staging tile, statistics slot).
template <int BN, int EW = 4>
__device__ __forceinline__ void epilogue_rows(const TcParams &p, uint32_t tmem_acc, int r, int b, int oy0, int ox0, int n0, int n_tile,
                                      int ksp, long long slot, float (*stat_s)[BN / 4 * 2], EpiStageT<EW> &es, int q, int lane) {
    constexpr int C_STEP = 32 * (EW / 4);
    const int c_first = (q >> 2) * 32;
    const int ly = r / p.BW, lx = r - ly * p.BW;
    const int oy = oy0 + ly, ox = ox0 + lx;
    const bool row_ok = (oy < p.Ho) && (ox < p.Wo);
    const long long m = p.up ? (long long)(2 * oy + p.py) * (2 * p.Wo) + (2 * ox + p.px)      // sub-pixel scatter
                             : (long long)oy * p.Wo + ox;  // row within the batch slice
    const long long row_off = (long long)b * p.d_batch_stride + m * (p.ldd ? p.ldd : p.N);
    const float bm = (p.bias_m && row_ok) ? __ldg(p.bias_m + m) : 0.0f;
    float *stg = es.tile[q];
    __syncwarp();
    es.rowoff[q][lane] = row_ok ? row_off : -1;
    __syncwarp();
    if (p.vq_tilemin) {                               // codebook search: per-row minima of the approximate distances, one per
        const float zz = row_ok ? __ldg(p.vq_zz + m) : 0.0f;                     // 32-code sub-tile (the refinement re-evaluates whole
        float *dst = p.vq_tilemin + (m * p.tiles_n + n_tile) * (BN / 32);        // sub-tiles: finer ones = fewer exact distances)
#pragma unroll 1
        for (int c0 = c_first; c0 < BN; c0 += C_STEP) {
            uint32_t v[32];
            tmem_ld32(tmem_acc + (uint32_t)c0, v);
            float best = INFINITY;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 e4 = __ldg(reinterpret_cast<const float4 *>(p.vq_ee + n0 + c0 + j));
                best = fminf(best, fminf(fminf((zz + e4.x) - 2.0f * __uint_as_float(v[j]), (zz + e4.y) - 2.0f * __uint_as_float(v[j + 1])),
                                         fminf((zz + e4.z) - 2.0f * __uint_as_float(v[j + 2]), (zz + e4.w) - 2.0f * __uint_as_float(v[j + 3]))));
            }
            if (row_ok) dst[c0 / 32] = best;
        }
        return;
    }
    if (p.ksplit > 1) {                               // raw partial sums; bias / residual / statistics happen in the reduce kernel
        float *dst = p.splitk_ws + (long long)ksp * p.split_stride + row_off + n0;
#pragma unroll 1
        for (int c0 = c_first; c0 < BN; c0 += C_STEP) {
            uint32_t v[32];
            tmem_ld32(tmem_acc + (uint32_t)c0, v);
            if (row_ok && n0 + c0 < p.n_valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4 *>(dst + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                            __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
        }
        return;
    }
#pragma unroll 1
    for (int c0 = c_first; c0 < BN; c0 += C_STEP) {
        uint32_t v[32];
        tmem_ld32(tmem_acc + (uint32_t)c0, v);
        const int n = n0 + c0;
        if (n >= p.n_valid) continue;
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = p.alpha * __uint_as_float(v[j]) + bm;
        if (p.out_nchw || n + 32 > p.n_valid) {      // ragged / NCHW tail (the 4-channel head): scalar stores, coalesced over ox
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (n + j < p.n_valid) {
                        float val = o[j] + (p.bias_n ? __ldg(p.bias_n + n + j) : 0.0f);
                        if (p.out_nchw) p.D[(((long long)b * p.n_valid + n + j) * p.Ho + oy) * p.Wo + ox] = val;
                        else p.D[row_off + n + j] = val + (p.R ? __ldg(p.R + row_off + n + j) : 0.0f);
                    }
                }
            }
            continue;
        }
        if (p.bias_n) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 bv = __ldg(reinterpret_cast<const float4 *>(p.bias_n + n + j));
                o[j] += bv.x; o[j + 1] += bv.y; o[j + 2] += bv.z; o[j + 3] += bv.w;
            }
        }
        if (p.R && row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 rv = __ldg(reinterpret_cast<const float4 *>(p.R + row_off + n + j));
                o[j] += rv.x; o[j + 1] += rv.y; o[j + 2] += rv.z; o[j + 3] += rv.w;
            }
        }
        if (p.vt_hi && n >= p.vt_col0) {              // v columns: transpose the 32 x 32 chunk through the staging tile
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4 *>(stg + lane * STG_STRIDE + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            __syncwarp();
            float col[32];                            // this lane's channel (n + lane) at the warp's 32 tokens
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) col[rr] = stg[rr * STG_STRIDE + lane];
            __syncwarp();
            const unsigned okmask = __ballot_sync(0xffffffffu, row_ok);
            const long long m0 = __shfl_sync(0xffffffffu, m, 0);
            const bool contiguous = (okmask == 0xffffffffu) && (__shfl_sync(0xffffffffu, m, 31) == m0 + 31) && ((m0 & 7) == 0);
            const long long base = ((long long)b * (p.n_valid - p.vt_col0) + (n - p.vt_col0 + lane)) * p.vt_T;
            if (contiguous) {                         // 32 consecutive tokens: 64 contiguous bytes per plane and lane
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) split2(col[j], col[j + 1], hi[j / 2], lo[j / 2]);
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    *reinterpret_cast<uint4 *>(p.vt_hi + base + m0 + 2 * j) = make_uint4(hi[j], hi[j + 1], hi[j + 2], hi[j + 3]);
                    *reinterpret_cast<uint4 *>(p.vt_lo + base + m0 + 2 * j) = make_uint4(lo[j], lo[j + 1], lo[j + 2], lo[j + 3]);
                }
            } else {
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) {       // (unrolled: col[] must stay in registers)
                    const long long mr = __shfl_sync(0xffffffffu, m, rr);
                    if ((okmask >> rr) & 1u) {
                        const __nv_bfloat16 h = __float2bfloat16_rn(col[rr]);
                        p.vt_hi[base + mr] = h;
                        p.vt_lo[base + mr] = __float2bfloat16_rn(col[rr] - __bfloat162float(h));
                    }
                }
            }
            continue;
        }
        if (p.stats) {                                // GroupNorm statistics of the finished output, per warp
            float *dst = &stat_s[q][(c0 / p.cpg) * 2];
            if (p.cpg == 4) chunk_group_sums<4>(o, row_ok, lane, dst);
            else if (p.cpg == 8) chunk_group_sums<8>(o, row_ok, lane, dst);
            else chunk_group_sums<16>(o, row_ok, lane, dst);
        }
        if (p.D) {                                    // fp32: stage the chunk, then 4 rows x 128 contiguous bytes per store
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4 *>(stg + lane * STG_STRIDE + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + (lane >> 3), cc = (lane & 7) * 4;
                const long long off = es.rowoff[q][rr];
                const float4 val = *reinterpret_cast<const float4 *>(stg + rr * STG_STRIDE + cc);
                if (off >= 0) *reinterpret_cast<float4 *>(p.D + off + n + cc) = val;
            }
            __syncwarp();
        }
        if (p.D_hi) {                                 // split bf16: hi words in floats [0,16), lo words in [16,32) of the row
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) split2(o[j], o[j + 1], hi[j / 2], lo[j / 2]);
            uint32_t *stg_u = reinterpret_cast<uint32_t *>(stg);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                *reinterpret_cast<uint4 *>(stg_u + lane * STG_STRIDE + j) = make_uint4(hi[j], hi[j + 1], hi[j + 2], hi[j + 3]);
                *reinterpret_cast<uint4 *>(stg_u + lane * STG_STRIDE + 16 + j) = make_uint4(lo[j], lo[j + 1], lo[j + 2], lo[j + 3]);
            }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int rr = it * 8 + (lane >> 2), cc = (lane & 3) * 4;           // cc in 32-bit words = 2 bf16 each
                const long long off = es.rowoff[q][rr];
                const uint4 h4 = *reinterpret_cast<const uint4 *>(stg_u + rr * STG_STRIDE + cc);
                const uint4 l4 = *reinterpret_cast<const uint4 *>(stg_u + rr * STG_STRIDE + 16 + cc);
                if (off >= 0) {
                    *reinterpret_cast<uint4 *>(p.D_hi + off + n + 2 * cc) = h4;
                    *reinterpret_cast<uint4 *>(p.D_lo + off + n + 2 * cc) = l4;
                }
            }
            __syncwarp();
        }
    }
    if (p.stats) {
        epilogue_bar_sync<32 * EW>();
        const int e = threadIdx.x - 64;                     // 0 .. 32 EW - 1
        const int nvals = (BN / p.cpg) * 2;
        if (e < nvals) {
            const int h = (((e >> 1) * p.cpg) >> 5) % (EW / 4) * 4;       // the four warps that processed this group's chunk
            const float v = (stat_s[h][e] + stat_s[h + 1][e]) + (stat_s[h + 2][e] + stat_s[h + 3][e]);
            const int g = n0 / p.cpg + (e >> 1);
            p.stats[(slot * 32 + g) * 2 + (e & 1)] = v;
        }
        epilogue_bar_sync<32 * EW>();
    }
}

}  // namespace tc
