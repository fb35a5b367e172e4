// Stage (iii), 128-channel convolutions with SWAPPED operands (sm_100a).
//
// The 128 -> 128 channel 3x3 convolutions at 256x256 / 128x128 are 40 % of the network's tensor time, and with pixels
// on the M side (tc_gemm_kernel / tc_gemm2_kernel) their MMA is 128 x 128 x 16: it reads 8 KB of operands from shared
// memory in 64 clk = 128 B/clk, the SM's whole shared-memory bandwidth, so the tensor pipe idles at ~62 % while TMA
// fills the ring (ncu).  The N = 256 GEMMs of the same network run at 84-92 %.  This kernel puts
// the OUTPUT CHANNELS on the M side and 256 OUTPUT PIXELS on the N side,
//
//     D^T[c, p] = sum_k W[c, k] * A[p, k]        (c < 128 channels, p < 256 pixels of the tile)
//
// so the instruction is 128 x 256 x 16: 12 KB of operands per 128 clk = 96 B/clk.  Both operands are K-major swizzled
// tiles exactly as before (the pixel tile is still ONE TMA box of the NHWC tensor shifted by the filter tap; only the
// descriptor roles are exchanged), and the split-bf16 products are W_hi*A_hi + W_hi*A_lo + W_lo*A_hi.
//
// The accumulator arrives transposed -- TMEM lane = channel, column = pixel -- which suits the NHWC output: for a fixed
// pixel the 32 lanes of a warp hold 32 consecutive channels, so every store (and residual load) of the epilogue is one
// 128-byte line, no staging.  GroupNorm statistics: a group is 4 neighbouring lanes, each lane sums its channel over
// the 128 pixels of a block in registers, two shuffles finish the group -- no shared memory, no cross-warp step.
//
// CTA = 6 warps as in tc_gemm_kernel (TMA producer, MMA issuer + TMEM allocation, 4 epilogue warps), persistent, two
// 256-column accumulators in TMEM (all 512 columns) so that the epilogue of tile i overlaps the main loop of tile i+1.
#include <stdlib.h>
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace {

using namespace tc;

template <int BK, int STAGES>
struct Tc3Cfg {
    static constexpr int ROW_BYTES = BK * 2;
    static constexpr int W_PLANE = 128 * ROW_BYTES;      // weight tile: 128 output channels (M side), one of (hi, lo)
    static constexpr int P_PLANE = 256 * ROW_BYTES;      // pixel tile: 256 output pixels (N side)
    static constexpr int STAGE_BYTES = 2 * W_PLANE + 2 * P_PLANE;
    static constexpr int TMEM_COLS = 512;                // two 256-column fp32 accumulators
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024;
    static_assert(SMEM <= 227 * 1024, "stage ring exceeds shared memory");
};

// 256-pixel boxes: BW2 x BH2 = 256 output pixels; the two 128-pixel halves of a box are the blocks the GroupNorm
// partial-sum layout is defined on (sgam_tc_gn_partial_floats): neighbours in x when the image is >= 256 wide, in y otherwise
struct SwapGeom {
    int BW2, BH2, tiles_x2, tiles_y2, wide, tiles128_x, tiles128;
};

// Epilogue of one 128-channel x 256-pixel accumulator (shared by the two swapped-operand kernels): warp q owns TMEM lanes
// 32q.. = output channels n0 + 32q + lane; the columns are the tile's pixels.
__device__ __forceinline__ void swap_epilogue_tile(const TcParams &p, const SwapGeom &g, uint32_t tmem_acc, int q, int lane, int n0,
                                                   int tx, int ty, int b) {
    const int cpg = p.cpg;                                          // channels per GroupNorm group (4 for 128 channels)
    const int c = n0 + 32 * q + lane;                       // this thread's output channel
    const float bn = p.bias_n ? __ldg(p.bias_n + c) : 0.0f;
    float s_acc = 0.f, q_acc = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_acc + (uint32_t)c0, v);
        const int ly = c0 / g.BW2, lx = c0 - ly * g.BW2;    // BW2 % 32 == 0: the 32 pixels of a chunk share a row
        const int oy = ty * g.BH2 + ly, ox = tx * g.BW2 + lx;
        const long long m = p.up ? (long long)(2 * oy + p.py) * (2 * p.Wo) + (2 * ox + p.px) : (long long)oy * p.Wo + ox;
        const long long off = (long long)b * p.d_batch_stride + m * p.N + c;
        const long long xstep = p.up ? 2LL * p.N : (long long)p.N;     // neighbouring tile pixels: every other output pixel when up
        float o[32];
        if (p.R) {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __ldg(p.R + off + (long long)j * xstep);
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] += p.alpha * __uint_as_float(v[j]) + bn;
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = p.alpha * __uint_as_float(v[j]) + bn;
        }
        if (p.D) {
#pragma unroll
            for (int j = 0; j < 32; ++j) p.D[off + (long long)j * xstep] = o[j]; // lanes = 32 consecutive channels: one 128-byte line
        }
        if (p.D_hi) {                       // split-bf16 planes for a consumer that is another tensor-core conv (Downsample /
#pragma unroll                                      // sub-pixel Upsample): 64 contiguous bytes per plane and pixel
            for (int j = 0; j < 32; ++j) {
                const __nv_bfloat16 h = __float2bfloat16_rn(o[j]);
                p.D_hi[off + (long long)j * xstep] = h;
                p.D_lo[off + (long long)j * xstep] = __float2bfloat16_rn(o[j] - __bfloat162float(h));
            }
        }
        if (p.stats) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { s_acc += o[j]; q_acc = fmaf(o[j], o[j], q_acc); }
            if ((c0 & 127) == 96) {                          // a 128-pixel block is complete
                float s = s_acc, qq = q_acc;
                for (int d = 1; d < cpg; d <<= 1) { s += __shfl_xor_sync(0xffffffffu, s, d); qq += __shfl_xor_sync(0xffffffffu, qq, d); }
                if ((lane & (cpg - 1)) == 0) {
                    const int half = c0 >> 7;
                    const int tx128 = g.wide ? 2 * tx + half : tx, ty128 = g.wide ? ty : 2 * ty + half;
                    const long long per_img = p.stat_tiles ? p.stat_tiles : g.tiles128;
                    const long long slot = (long long)b * per_img + p.stat_tile0 + (long long)ty128 * g.tiles128_x + tx128;
                    const int grp = c / cpg;
                    p.stats[(slot * 32 + grp) * 2] = s;
                    p.stats[(slot * 32 + grp) * 2 + 1] = qq;
                }
                s_acc = 0.f; q_acc = 0.f;
            }
        }
    }
}

template <int BK, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_swap_kernel(const __grid_constant__ CUtensorMap mapP_hi, const __grid_constant__ CUtensorMap mapP_lo,
                    const __grid_constant__ CUtensorMap mapW_hi, const __grid_constant__ CUtensorMap mapW_lo, const TcParams p,
                    const SwapGeom g) {
    using Cfg = Tc3Cfg<BK, STAGES>;
    constexpr int ROW_BYTES = Cfg::ROW_BYTES, W_PLANE = Cfg::W_PLANE, P_PLANE = Cfg::P_PLANE, STAGE_BYTES = Cfg::STAGE_BYTES;
    SGAM_PDL_TRIGGER();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = p.taps * p.kblocks_per_tap;
    const int total_tiles = p.tiles_m * p.tiles_n;          // pixel boxes x 128-channel tiles

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    SGAM_PDL_WAIT();

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint32_t tx_bytes = (p.nsplit == 3) ? STAGE_BYTES : (W_PLANE + P_PLANE);
            int kbg = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n0 = (tile % p.tiles_n) * 128;
                int t = tile / p.tiles_n;
                const int tx = t % g.tiles_x2; t /= g.tiles_x2;
                const int ty = t % g.tiles_y2; const int b = t / g.tiles_y2;
                for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
                    const int s = kbg % STAGES, it = kbg / STAGES;
                    mbar_wait(&empty_bar[s], (it & 1) ^ 1);
                    uint8_t *st = smem + (size_t)s * STAGE_BYTES;
                    const int tap = kb / p.kblocks_per_tap, kc = kb - tap * p.kblocks_per_tap;
                    const int kh = tap / p.ks, kw = tap - kh * p.ks;
                    const int cx = tx * g.BW2 + kw - p.pad + p.shift_x, cy = ty * g.BH2 + kh - p.pad + p.shift_y;
                    mbar_expect_tx(&full_bar[s], tx_bytes);
                    tma_load_3d(st, &mapW_hi, &full_bar[s], kb * BK, n0, 0);
                    tma_load_4d(st + 2 * W_PLANE, &mapP_hi, &full_bar[s], kc * BK, cx, cy, b);
                    if (p.nsplit == 3) {
                        tma_load_3d(st + W_PLANE, &mapW_lo, &full_bar[s], kb * BK, n0, 0);
                        tma_load_4d(st + 2 * W_PLANE + P_PLANE, &mapP_lo, &full_bar[s], kc * BK, cx, cy, b);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: D^T (128 channels x 256 pixels) += W (128 x 16) . A^T (16 x 256) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, 256);
            int kbg = 0, li = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
                const int acc = li & 1;
                mbar_wait(&tmem_empty_bar[acc], ((li >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256);
                for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
                    const int s = kbg % STAGES, it = kbg / STAGES;
                    mbar_wait(&full_bar[s], it & 1);
                    tc_fence_after();
                    uint8_t *st = smem + (size_t)s * STAGE_BYTES;
                    const uint64_t w_hi = make_smem_desc<ROW_BYTES>(st), w_lo = make_smem_desc<ROW_BYTES>(st + W_PLANE);
                    const uint64_t a_hi = make_smem_desc<ROW_BYTES>(st + 2 * W_PLANE), a_lo = make_smem_desc<ROW_BYTES>(st + 2 * W_PLANE + P_PLANE);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t off = (uint64_t)(k * 2);
                        umma_bf16(tmem_d, w_hi + off, a_hi + off, idesc, (kb || k) ? 1u : 0u);
                        if (p.nsplit == 3) {
                            umma_bf16(tmem_d, w_hi + off, a_lo + off, idesc, 1u);
                            umma_bf16(tmem_d, w_lo + off, a_hi + off, idesc, 1u);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full_bar[acc]);
            }
        }
    } else {
        // ===== epilogue: warp q owns TMEM lanes 32q.. = output channels n0 + 32q + lane; columns are the box's pixels =====
        const int q = warp & 3;
        int li = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
            const int acc = li & 1;
            const int n0 = (tile % p.tiles_n) * 128;
            int t = tile / p.tiles_n;
            const int tx = t % g.tiles_x2; t /= g.tiles_x2;
            const int ty = t % g.tiles_y2; const int b = t / g.tiles_y2;
            mbar_wait(&tmem_full_bar[acc], (li >> 1) & 1);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
            swap_epilogue_tile(p, g, tmem_acc, q, lane, n0, tx, ty, b);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
    }
}

// ---- 3x3 convolutions, one pixel-row load for the three horizontal taps ---------------------------------------------------
// ncu on the kernel above (128 -> 128 channels at 256 x 256, 36 % of the step): 3.9 GB cross the L2 -> SM crossbar per launch at
// 11.8 TB/s -- the chip-wide L2 throughput cap -- while the tensor pipe is 81 % busy: the layer is bound by operand traffic, and
// two thirds of it are the nine tap-shifted copies of the SAME pixels.  The three taps of one filter row read one image row shifted
// by one pixel, i.e. by one 128-byte row of the K-major operand, and a SWIZZLE_128B UMMA descriptor may start at ANY 128-byte row
// (the swizzle is a function of the absolute shared-memory address; tools/probes/umma_shift_probe.cu: exact for shifts 0..7 with
// the descriptor's base-offset field left 0).  So this kernel loads, per filter row kh and 64-channel chunk, the tile's pixel row
// with a one-pixel halo on each side ONCE ((BWT + 2) x 64 channels, a 256-pixel box + a 2-pixel box; TMA's zero fill is the
// padding) and issues the MMAs of kw = 0, 1, 2 with the B descriptor advanced by kw rows; only the 32 KB weight tiles stream per
// tap.  Per 256-pixel tile: 6 x 66 KB of pixels + 18 x 32 KB of weights = 0.97 MB instead of 1.73 MB.
// RPT image rows per tile: 1 (W % 256 == 0) or 2 (W == 128: two N = 128 MMAs per step, one per image row).
template <int RPT>
struct RowCfg {
    static constexpr int BWT = 256 / RPT;                                          // tile pixels per image row
    static constexpr int ROW_PLANE = ((BWT + 2) * 128 + 1023) / 1024 * 1024;       // one image row + halo, one plane (hi or lo)
    static constexpr int P_PLANE = RPT * ROW_PLANE;
    static constexpr int P_STAGE = 2 * P_PLANE;
    static constexpr int W_PLANE = 128 * 128;                                      // 128 output channels x 64 input channels
    static constexpr int W_STAGE = 2 * W_PLANE;
    static constexpr int PS = 2, WS = 2;                                           // ring depths (pixel rows, weight tiles)
    static constexpr uint32_t P_TX = 2u * RPT * (BWT + 2) * 128;                   // bytes TMA delivers per pixel stage (hi + lo)
    static constexpr size_t SMEM = (size_t)PS * P_STAGE + (size_t)WS * W_STAGE + 1024;
    static_assert(SMEM <= 227 * 1024, "rings exceed shared memory");
};

template <int RPT>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_swaprow_kernel(const __grid_constant__ CUtensorMap mapP_hi, const __grid_constant__ CUtensorMap mapP_lo,
                       const __grid_constant__ CUtensorMap mapT_hi, const __grid_constant__ CUtensorMap mapT_lo,
                       const __grid_constant__ CUtensorMap mapW_hi, const __grid_constant__ CUtensorMap mapW_lo, const TcParams p,
                       const SwapGeom g, const int Cin) {
    using Cfg = RowCfg<RPT>;
    constexpr int BWT = Cfg::BWT, ROW_PLANE = Cfg::ROW_PLANE, P_PLANE = Cfg::P_PLANE, P_STAGE = Cfg::P_STAGE;
    constexpr int W_PLANE = Cfg::W_PLANE, W_STAGE = Cfg::W_STAGE, PS = Cfg::PS, WS = Cfg::WS;
    SGAM_PDL_TRIGGER();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *p_ring = smem, *w_ring = smem + (size_t)PS * P_STAGE;
    __shared__ __align__(8) uint64_t p_full[PS], p_empty[PS], w_full[WS], w_empty[WS], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KC = Cin / 64;                                    // 64-channel chunks
    const int total_tiles = p.tiles_m * p.tiles_n;

    if (threadIdx.x == 0) {
        for (int s = 0; s < PS; ++s) { mbar_init(&p_full[s], 1); mbar_init(&p_empty[s], 1); }
        for (int s = 0; s < WS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    SGAM_PDL_WAIT();

    if (warp == 0) {
        // ===== TMA producer: one pixel stage per (filter row, channel chunk), three weight stages behind it =====
        if (lane == 0) {
            const bool three = p.nsplit == 3;
            int pc = 0, wc = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n0 = (tile % p.tiles_n) * 128;
                int t = tile / p.tiles_n;
                const int tx = t % g.tiles_x2; t /= g.tiles_x2;
                const int ty = t % g.tiles_y2; const int b = t / g.tiles_y2;
                const int cx = tx * BWT - 1;
                for (int kh = 0; kh < 3; ++kh) {
                    for (int kc = 0; kc < KC; ++kc, ++pc) {
                        const int ps = pc % PS;
                        mbar_wait(&p_empty[ps], ((pc / PS) & 1) ^ 1);
                        uint8_t *pst = p_ring + (size_t)ps * P_STAGE;
                        mbar_expect_tx(&p_full[ps], three ? Cfg::P_TX : Cfg::P_TX / 2);
#pragma unroll
                        for (int r = 0; r < RPT; ++r) {
                            const int cy = ty * RPT + r + kh - 1;
                            uint8_t *row = pst + r * ROW_PLANE;
                            tma_load_4d(row, &mapP_hi, &p_full[ps], kc * 64, cx, cy, b);
                            tma_load_4d(row + BWT * 128, &mapT_hi, &p_full[ps], kc * 64, cx + BWT, cy, b);
                            if (three) {
                                tma_load_4d(row + P_PLANE, &mapP_lo, &p_full[ps], kc * 64, cx, cy, b);
                                tma_load_4d(row + P_PLANE + BWT * 128, &mapT_lo, &p_full[ps], kc * 64, cx + BWT, cy, b);
                            }
                        }
                        for (int kw = 0; kw < 3; ++kw, ++wc) {
                            const int ws = wc % WS;
                            mbar_wait(&w_empty[ws], ((wc / WS) & 1) ^ 1);
                            uint8_t *wst = w_ring + (size_t)ws * W_STAGE;
                            const int kcol = (kh * 3 + kw) * Cin + kc * 64;
                            mbar_expect_tx(&w_full[ws], three ? W_STAGE : W_PLANE);
                            tma_load_3d(wst, &mapW_hi, &w_full[ws], kcol, n0, 0);
                            if (three) tma_load_3d(wst + W_PLANE, &mapW_lo, &w_full[ws], kcol, n0, 0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: D^T (128 channels x BWT pixels per image row) += W (128 x 16) . P^T (16 x BWT), P advanced by kw rows =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, BWT);
            const bool three = p.nsplit == 3;
            int pc = 0, wc = 0, li = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
                const int acc = li & 1;
                mbar_wait(&tmem_empty_bar[acc], ((li >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int kh = 0; kh < 3; ++kh) {
                    for (int kc = 0; kc < KC; ++kc, ++pc) {
                        const int ps = pc % PS;
                        mbar_wait(&p_full[ps], (pc / PS) & 1);
                        tc_fence_after();
                        const uint8_t *pst = p_ring + (size_t)ps * P_STAGE;
                        for (int kw = 0; kw < 3; ++kw, ++wc) {
                            const int ws = wc % WS;
                            mbar_wait(&w_full[ws], (wc / WS) & 1);
                            tc_fence_after();
                            const uint8_t *wst = w_ring + (size_t)ws * W_STAGE;
                            const uint64_t w_hi = make_smem_desc<128>(wst), w_lo = make_smem_desc<128>(wst + W_PLANE);
                            const uint32_t fresh = (kh | kc | kw) ? 1u : 0u;
#pragma unroll
                            for (int r = 0; r < RPT; ++r) {
                                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256 + r * BWT);
                                const uint64_t a_hi = make_smem_desc<128>(pst + r * ROW_PLANE + kw * 128);
                                const uint64_t a_lo = make_smem_desc<128>(pst + P_PLANE + r * ROW_PLANE + kw * 128);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint64_t off = (uint64_t)(k * 2);
                                    umma_bf16(tmem_d, w_hi + off, a_hi + off, idesc, (fresh | (uint32_t)k) ? 1u : 0u);
                                    if (three) {
                                        umma_bf16(tmem_d, w_hi + off, a_lo + off, idesc, 1u);
                                        umma_bf16(tmem_d, w_lo + off, a_hi + off, idesc, 1u);
                                    }
                                }
                            }
                            umma_commit(&w_empty[ws]);
                        }
                        umma_commit(&p_empty[ps]);
                    }
                }
                umma_commit(&tmem_full_bar[acc]);
            }
        }
    } else {
        const int q = warp & 3;
        int li = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
            const int acc = li & 1;
            const int n0 = (tile % p.tiles_n) * 128;
            int t = tile / p.tiles_n;
            const int tx = t % g.tiles_x2; t /= g.tiles_x2;
            const int ty = t % g.tiles_y2; const int b = t / g.tiles_y2;
            mbar_wait(&tmem_full_bar[acc], (li >> 1) & 1);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
            swap_epilogue_tile(p, g, tmem_acc, q, lane, n0, tx, ty, b);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

int swaprow_mode() {           // SGAM_TC_SWAPROW=0: every swapped-operand conv on tc_gemm_swap_kernel (nine pixel boxes per chunk)
    static int v = -1;
    if (v < 0) { const char *e = getenv("SGAM_TC_SWAPROW"); v = e ? atoi(e) : 1; }
    return v;
}

template <int RPT>
int launch_swaprow(const void *x_hi, const void *x_lo, const CUtensorMap &w_hi, const CUtensorMap &w_lo, const TcParams &p, const SwapGeom &g,
                   int B, int H, int W, int Cin, cudaStream_t s) {
    using Cfg = RowCfg<RPT>;
    CUtensorMap p_hi, p_lo, t_hi, t_lo;
    const long long adims[4] = {Cin, W, H, B};
    const int abox[4] = {64, Cfg::BWT, 1, 1}, tbox[4] = {64, 2, 1, 1};
    int rc;
    if ((rc = make_map(&p_hi, x_hi, 4, adims, abox)) || (rc = make_map(&p_lo, x_lo, 4, adims, abox)) ||
        (rc = make_map(&t_hi, x_hi, 4, adims, tbox)) || (rc = make_map(&t_lo, x_lo, 4, adims, tbox)))
        return rc;
    static bool configured = false;
    if (!configured) {
        SGAM_CUDA_OK(cudaFuncSetAttribute(tc_gemm_swaprow_kernel<RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        configured = true;
    }
    const int total = p.tiles_m * p.tiles_n;
    const int grid = total < sm_count_cached() ? total : sm_count_cached();
    SGAM_PDL_LAUNCH(SGAM_PDL_GEMM1, tc_gemm_swaprow_kernel<RPT>, grid, TC_THREADS, Cfg::SMEM, s, p_hi, p_lo, t_hi, t_lo, w_hi, w_lo, p, g, Cin);
    return SGAM_OK;
}

// ---- GroupNorm + swish inside the operand path of the row-reuse kernel ---------------------------------------------------------
// The separate GroupNorm-apply pass of a 128-channel 256 x 256 layer moves 537 MB through HBM (94 us) to produce the split-bf16
// operand the conv then reads.  With one pixel-row load per filter row every input element enters shared memory only three times, so
// the transform can live here: TMA lands the fp32 rows ((256 + 2) pixels x 64 channels = 66 KB, exactly the size of the hi + lo
// planes), eight transform warps read them into registers, apply (x - mean) * rstd * gamma + beta, swish and the bf16 split -- the same
// expressions as gn_apply_split_kernel, so the operand bits are identical -- and write the two K-major SWIZZLE_128B planes IN PLACE
// (all reads of a stage finish before the first write: named barrier); pixels outside the image become exact zeros AFTER the
// activation, as the conv's padding requires.  W % 256 == 0 (one image row per tile) only.
// STATUS: correct (bit-identical to the two-kernel path, tests/test_gpu_tc.py) but SLOWER -- 531 us against 94 + 295 us at 8 x 256 x 256
// (442 us with the arithmetic removed): the N = 256 MMAs already read 96 B/clk of operands from shared memory and TMA writes another
// 36 B/clk; the transform's 29 B/clk of loads / stores push the sum past the SM's 128 B/clk, and with 66 KB stages only two fit, so
// TMA -> transform -> MMA serialise per buffer.  Opt-in (SGAM_FUSED_GNCONV=1); profiles/r2_rejected_experiments.txt.
constexpr int GNC_TRANSFORM_WARPS = 8;
constexpr int GNC_THREADS = TC_THREADS + 32 * GNC_TRANSFORM_WARPS;

struct GnOperand {
    const float *meanrstd, *gamma, *beta;   // [B][32][2], [Cin], [Cin]
    int cpg, H, W;                          // channels per group (Cin / 32, a multiple of 4); image extent
    int probe;                              // development probe (SGAM_GNCONV_PROBE): 1 = no swish, 2 = no arithmetic at all
};

__global__ void __launch_bounds__(GNC_THREADS, 1)
tc_gemm_gnconv_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapXT,
                      const __grid_constant__ CUtensorMap mapW_hi, const __grid_constant__ CUtensorMap mapW_lo, const TcParams p,
                      const SwapGeom g, const int Cin, const GnOperand gn) {
    using Cfg = RowCfg<1>;
    constexpr int BWT = 256, P_PLANE = Cfg::P_PLANE, P_STAGE = Cfg::P_STAGE, W_PLANE = Cfg::W_PLANE, W_STAGE = Cfg::W_STAGE;
    constexpr int PS = Cfg::PS, WS = Cfg::WS;
    constexpr uint32_t X_TX = (BWT + 2) * 64 * 4;               // fp32 bytes per pixel stage
    static_assert(X_TX <= P_STAGE, "the fp32 rows must fit the stage they are transformed in");
    SGAM_PDL_TRIGGER();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *p_ring = smem, *w_ring = smem + (size_t)PS * P_STAGE;
    __shared__ __align__(8) uint64_t x_full[PS], p_ready[PS], p_empty[PS], w_full[WS], w_empty[WS], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KC = Cin / 64;
    const int total_tiles = p.tiles_m * p.tiles_n;

    if (threadIdx.x == 0) {
        for (int s = 0; s < PS; ++s) { mbar_init(&x_full[s], 1); mbar_init(&p_ready[s], GNC_TRANSFORM_WARPS); mbar_init(&p_empty[s], 1); }
        for (int s = 0; s < WS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    SGAM_PDL_WAIT();

    if (warp == 0) {
        // ===== TMA producer: fp32 pixel rows (one stage per filter row and channel chunk), three weight tiles behind each =====
        if (lane == 0) {
            const bool three = p.nsplit == 3;
            int pc = 0, wc = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n0 = (tile % p.tiles_n) * 128;
                int t = tile / p.tiles_n;
                const int tx = t % g.tiles_x2; t /= g.tiles_x2;
                const int ty = t % g.tiles_y2; const int b = t / g.tiles_y2;
                const int cx = tx * BWT - 1;
                for (int kh = 0; kh < 3; ++kh) {
                    for (int kc = 0; kc < KC; ++kc, ++pc) {
                        const int ps = pc % PS;
                        mbar_wait(&p_empty[ps], ((pc / PS) & 1) ^ 1);
                        uint8_t *pst = p_ring + (size_t)ps * P_STAGE;
                        mbar_expect_tx(&x_full[ps], X_TX);
                        tma_load_4d(pst, &mapX, &x_full[ps], kc * 64, cx, ty + kh - 1, b);
                        tma_load_4d(pst + BWT * 256, &mapXT, &x_full[ps], kc * 64, cx + BWT, ty + kh - 1, b);
                        for (int kw = 0; kw < 3; ++kw, ++wc) {
                            const int ws = wc % WS;
                            mbar_wait(&w_empty[ws], ((wc / WS) & 1) ^ 1);
                            uint8_t *wst = w_ring + (size_t)ws * W_STAGE;
                            const int kcol = (kh * 3 + kw) * Cin + kc * 64;
                            mbar_expect_tx(&w_full[ws], three ? W_STAGE : W_PLANE);
                            tma_load_3d(wst, &mapW_hi, &w_full[ws], kcol, n0, 0);
                            if (three) tma_load_3d(wst + W_PLANE, &mapW_lo, &w_full[ws], kcol, n0, 0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (as tc_gemm_swaprow_kernel<1>, but a pixel stage is ready when the transform warps have written it) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, BWT);
            const bool three = p.nsplit == 3;
            int pc = 0, wc = 0, li = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
                const int acc = li & 1;
                mbar_wait(&tmem_empty_bar[acc], ((li >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256);
                for (int kh = 0; kh < 3; ++kh) {
                    for (int kc = 0; kc < KC; ++kc, ++pc) {
                        const int ps = pc % PS;
                        mbar_wait(&p_ready[ps], (pc / PS) & 1);
                        tc_fence_after();
                        const uint8_t *pst = p_ring + (size_t)ps * P_STAGE;
                        for (int kw = 0; kw < 3; ++kw, ++wc) {
                            const int ws = wc % WS;
                            mbar_wait(&w_full[ws], (wc / WS) & 1);
                            tc_fence_after();
                            const uint8_t *wst = w_ring + (size_t)ws * W_STAGE;
                            const uint64_t w_hi = make_smem_desc<128>(wst), w_lo = make_smem_desc<128>(wst + W_PLANE);
                            const uint64_t a_hi = make_smem_desc<128>(pst + kw * 128), a_lo = make_smem_desc<128>(pst + P_PLANE + kw * 128);
                            const uint32_t fresh = (kh | kc | kw) ? 1u : 0u;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t off = (uint64_t)(k * 2);
                                umma_bf16(tmem_d, w_hi + off, a_hi + off, idesc, (fresh | (uint32_t)k) ? 1u : 0u);
                                if (three) {
                                    umma_bf16(tmem_d, w_hi + off, a_lo + off, idesc, 1u);
                                    umma_bf16(tmem_d, w_lo + off, a_hi + off, idesc, 1u);
                                }
                            }
                            umma_commit(&w_empty[ws]);
                        }
                        umma_commit(&p_empty[ps]);
                    }
                }
                umma_commit(&tmem_full_bar[acc]);
            }
        }
    } else if (warp < 6) {
        const int q = warp & 3;
        int li = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
            const int acc = li & 1;
            const int n0 = (tile % p.tiles_n) * 128;
            int t = tile / p.tiles_n;
            const int tx = t % g.tiles_x2; t /= g.tiles_x2;
            const int ty = t % g.tiles_y2; const int b = t / g.tiles_y2;
            mbar_wait(&tmem_full_bar[acc], (li >> 1) & 1);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
            swap_epilogue_tile(p, g, tmem_acc, q, lane, n0, tx, ty, b);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
    } else {
        // ===== transform warps: fp32 rows -> GroupNorm + swish -> split-bf16 planes, in place =====
        const int tt = threadIdx.x - TC_THREADS;                    // 0 .. 255
        const int quad = tt & 15;                                   // this thread's channel quad inside the 64-channel chunk
        constexpr int NQ = (BWT + 2) * 16;                          // float4 elements per stage
        constexpr int PER = (NQ + 255) / 256;                       // 17
        int pc = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int t = tile / p.tiles_n;
            const int tx = t % g.tiles_x2; t /= g.tiles_x2;
            const int ty = t % g.tiles_y2; const int b = t / g.tiles_y2;
            const int cx = tx * BWT - 1;
            for (int kh = 0; kh < 3; ++kh) {
                const int cy = ty + kh - 1;
                const bool row_in = cy >= 0 && cy < gn.H;
                for (int kc = 0; kc < KC; ++kc, ++pc) {
                    const int ps = pc % PS;
                    const int c0 = kc * 64 + quad * 4;
                    const int grp = c0 / gn.cpg;
                    const float mu = __ldg(gn.meanrstd + (b * 32 + grp) * 2), rs = __ldg(gn.meanrstd + (b * 32 + grp) * 2 + 1);
                    const float4 ga = __ldg(reinterpret_cast<const float4 *>(gn.gamma + c0)), be = __ldg(reinterpret_cast<const float4 *>(gn.beta + c0));
                    uint8_t *pst = p_ring + (size_t)ps * P_STAGE;
                    mbar_wait(&x_full[ps], (pc / PS) & 1);          // (one poller per warp + __syncwarp measured slower)
                    uint2 hi[PER], lo[PER];
#pragma unroll
                    for (int j = 0; j < PER; ++j) {
                        const int i = tt + 256 * j;                 // float4 index: pixel i / 16, quad i % 16 (= quad)
                        if (i < NQ) {
                            const int px = i >> 4;
                            const float4 v = *reinterpret_cast<const float4 *>(pst + (size_t)i * 16);
                            float o[4] = {(v.x - mu) * rs * ga.x + be.x, (v.y - mu) * rs * ga.y + be.y,
                                          (v.z - mu) * rs * ga.z + be.z, (v.w - mu) * rs * ga.w + be.w};
                            if (gn.probe == 2) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
                            if (gn.probe == 0) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) o[k] = __fdividef(o[k], 1.0f + __expf(-o[k]));
                            }
                            const int gx = cx + px;
                            if (!(row_in && gx >= 0 && gx < gn.W)) { o[0] = 0.0f; o[1] = 0.0f; o[2] = 0.0f; o[3] = 0.0f; }
                            split2(o[0], o[1], hi[j].x, lo[j].x);
                            split2(o[2], o[3], hi[j].y, lo[j].y);
                        }
                    }
                    asm volatile("bar.sync 2, 256;" ::: "memory");         // every fp32 value of the stage is in registers
#pragma unroll
                    for (int j = 0; j < PER; ++j) {
                        const int i = tt + 256 * j;
                        if (i < NQ) {
                            const int px = i >> 4;
                            // K-major SWIZZLE_128B row px: 16-byte chunk (quad / 2) ^ (row & 7), 8 bytes at (quad & 1) * 8
                            const uint32_t off = (uint32_t)px * 128u + ((((uint32_t)quad >> 1) ^ ((uint32_t)px & 7u)) << 4) + (((uint32_t)quad & 1u) << 3);
                            *reinterpret_cast<uint2 *>(pst + off) = hi[j];
                            *reinterpret_cast<uint2 *>(pst + P_PLANE + off) = lo[j];
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&p_ready[ps]);
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

int gnconv_mode() {            // SGAM_TC_GNCONV=0: GroupNorm apply as a separate pass everywhere
    static int v = -1;
    if (v < 0) { const char *e = getenv("SGAM_TC_GNCONV"); v = e ? atoi(e) : 1; }
    return v;
}

int swap_mode() {              // 0 = off, 64 = BK 64 / 2 stages, 32 = BK 32 / 4 stages (SGAM_TC_SWAP)
    static int v = -1;
    if (v < 0) { const char *e = getenv("SGAM_TC_SWAP"); v = e ? atoi(e) : 64; if (v != 0 && v != 32 && v != 64) v = 64; }
    return v;
}

template <int BK, int STAGES>
int launch_swap(const CUtensorMap &p_hi, const CUtensorMap &p_lo, const CUtensorMap &w_hi, const CUtensorMap &w_lo, const TcParams &p,
                const SwapGeom &g, cudaStream_t s) {
    using Cfg = Tc3Cfg<BK, STAGES>;
    static bool configured = false;
    if (!configured) {
        SGAM_CUDA_OK(cudaFuncSetAttribute(tc_gemm_swap_kernel<BK, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        configured = true;
    }
    const int total = p.tiles_m * p.tiles_n;
    const int grid = total < sm_count_cached() ? total : sm_count_cached();
    SGAM_PDL_LAUNCH(SGAM_PDL_GEMM1, (tc_gemm_swap_kernel<BK, STAGES>), grid, TC_THREADS, Cfg::SMEM, s, p_hi, p_lo, w_hi, w_lo, p, g);
    return SGAM_OK;
}

}  // namespace

namespace tc {

// Does the swapped-operand kernel apply?  128 output channels, stride 1, plain fp32 NHWC output, an image that tiles
// exactly into 256-pixel boxes whose rows are a multiple of 32 pixels, and enough boxes to fill the machine.
bool swap_applicable(int B, int Ho, int Wo, int Cin, int Cout, int stride, bool fp32_nhwc_only) {
    static int max_c = -1;                       // SGAM_TC_SWAP_MAXC: largest channel count routed here (experiment knob)
    if (max_c < 0) { const char *e = getenv("SGAM_TC_SWAP_MAXC"); max_c = e ? atoi(e) : 128; }
    if (!swap_mode() || Cout % 128 || Cout > max_c || stride != 1 || !fp32_nhwc_only || Cin % 64) return false;
    const bool exact = (Wo >= 256) ? (Wo % 256 == 0) : (Wo >= 32 && (Wo & (Wo - 1)) == 0 && Ho % (256 / Wo) == 0);
    if (!exact) return false;
    return (long long)B * Ho * Wo / 256 * (Cout / 128) >= sm_count_cached();
}

int launch_conv_swap(const void *x_hi, const void *x_lo, const void *w_hi, const void *w_lo, TcParams p, int B, int H, int W, int Cin,
                     int Cout, int ksize, cudaStream_t s) {
    const int mode = swap_mode(), BK = mode == 32 ? 32 : 64;
    SwapGeom g;
    g.BW2 = W >= 256 ? 256 : W; g.BH2 = 256 / g.BW2; g.tiles_x2 = W / g.BW2; g.tiles_y2 = H / g.BH2;
    g.wide = W >= 256; g.tiles128_x = W >= 128 ? W / 128 : 1;
    { const int BW = W >= 128 ? 128 : W, BH = 128 / BW; g.tiles128 = cdiv(W, BW) * cdiv(H, BH); }
    CUtensorMap p_hi, p_lo, wm_hi, wm_lo;
    const long long adims[4] = {Cin, W, H, B};
    const int abox[4] = {BK, g.BW2, g.BH2, 1};
    const long long bdims[3] = {(long long)ksize * ksize * Cin, Cout, 1};
    const int bbox[3] = {BK, 128, 1};
    int rc;
    if ((rc = make_map(&p_hi, x_hi, 4, adims, abox)) || (rc = make_map(&p_lo, x_lo, 4, adims, abox)) ||
        (rc = make_map(&wm_hi, w_hi, 3, bdims, bbox)) || (rc = make_map(&wm_lo, w_lo, 3, bdims, bbox)))
        return rc;
    p.kblocks_per_tap = Cin / BK;
    p.tiles_m = g.tiles_x2 * g.tiles_y2 * B;
    p.tiles_n = Cout / 128;
    p.ksplit = 1;
    // plain 3x3 convolutions: one pixel-row load for the three horizontal taps (the sub-pixel Upsample's 2x2 parity convs stay here)
    if (ksize == 3 && !p.up && p.shift_x == 0 && p.shift_y == 0 && swaprow_mode() && (W % 256 == 0 || W == 128)) {
        CUtensorMap w64_hi, w64_lo;
        const int wbox[3] = {64, 128, 1};
        if ((rc = make_map(&w64_hi, w_hi, 3, bdims, wbox)) || (rc = make_map(&w64_lo, w_lo, 3, bdims, wbox))) return rc;
        if (W == 128) return launch_swaprow<2>(x_hi, x_lo, w64_hi, w64_lo, p, g, B, H, W, Cin, s);
        return launch_swaprow<1>(x_hi, x_lo, w64_hi, w64_lo, p, g, B, H, W, Cin, s);
    }
    if (BK == 32) return launch_swap<32, 4>(p_hi, p_lo, wm_hi, wm_lo, p, g, s);
    return launch_swap<64, 2>(p_hi, p_lo, wm_hi, wm_lo, p, g, s);
}

// Does the fused GroupNorm + conv kernel apply?  A plain 3x3 / stride-1 conv the swapped-operand row kernel takes, one image row per
// tile (W % 256 == 0), channel groups that are whole channel quads.
bool gnconv_applicable(int B, int H, int W, int Cin, int Cout) {
    return gnconv_mode() && swaprow_mode() && W % 256 == 0 && Cin % 64 == 0 && (Cin / 32) % 4 == 0 && swap_applicable(B, H, W, Cin, Cout, 1, true);
}

int launch_gnconv(const float *x, const float *meanrstd, const float *gamma, const float *beta, const void *w_hi, const void *w_lo, TcParams p,
                  int B, int H, int W, int Cin, int Cout, cudaStream_t s) {
    SwapGeom g;
    g.BW2 = 256; g.BH2 = 1; g.tiles_x2 = W / 256; g.tiles_y2 = H;
    g.wide = 1; g.tiles128_x = W / 128;
    g.tiles128 = cdiv(W, 128) * H;
    CUtensorMap mx, mxt, wm_hi, wm_lo;
    const long long adims[4] = {Cin, W, H, B};
    const int abox[4] = {64, 256, 1, 1}, tbox[4] = {64, 2, 1, 1};
    const long long bdims[3] = {9LL * Cin, Cout, 1};
    const int wbox[3] = {64, 128, 1};
    int rc;
    if ((rc = make_map_f32(&mx, x, 4, adims, abox)) || (rc = make_map_f32(&mxt, x, 4, adims, tbox)) ||
        (rc = make_map(&wm_hi, w_hi, 3, bdims, wbox)) || (rc = make_map(&wm_lo, w_lo, 3, bdims, wbox)))
        return rc;
    p.kblocks_per_tap = Cin / 64;
    p.tiles_m = g.tiles_x2 * g.tiles_y2 * B;
    p.tiles_n = Cout / 128;
    p.ksplit = 1;
    static int probe = -1;
    if (probe < 0) { const char *e = getenv("SGAM_GNCONV_PROBE"); probe = e ? atoi(e) : 0; }
    GnOperand gn{meanrstd, gamma, beta, Cin / 32, H, W, probe};
    static bool configured = false;
    if (!configured) {
        SGAM_CUDA_OK(cudaFuncSetAttribute(tc_gemm_gnconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RowCfg<1>::SMEM));
        configured = true;
    }
    const int total = p.tiles_m * p.tiles_n;
    const int grid = total < sm_count_cached() ? total : sm_count_cached();
    SGAM_PDL_LAUNCH(SGAM_PDL_GEMM1, tc_gemm_gnconv_kernel, grid, GNC_THREADS, RowCfg<1>::SMEM, s, mx, mxt, wm_hi, wm_lo, p, g, Cin, gn);
    return SGAM_OK;
}

}  // namespace tc
