// 2-CTA (cta_group::2) variant of the tcgen05 implicit-GEMM kernel (see net_tc.cu for the operand staging scheme).
//
// Why: in the 1-CTA kernel every 128x128x16 MMA reads 4 KB of A and 4 KB of B from shared memory in 64 clk
// (= 128 B/clk, the SMEM bandwidth of an SM) while TMA refills the ring at another 85 B/clk, which caps the tensor
// pipe at ~60 %.  Here a cluster of two CTAs computes a 256-row x BN tile: each CTA stages its own
// 128 pixel rows of A and only HALF of the weight tile; the pair's MMA (M = 256, issued by the leader CTA) reads B
// from both shared memories.  Per SM that is 96 B/clk of operand reads and 62 B/clk of TMA fill for BN = 128, and the
// kernel becomes MMA-bound for BN = 256.
//
// Protocol (r = rank in the pair, leader = rank 0):
//   full[s]        lives in the leader; 1 arrival (the leader's producer: arrive.expect_tx of BOTH CTAs' bytes) + the
//                  complete_tx of both CTAs' TMA loads (cp.async.bulk.tensor ... .cta_group::2 onto the leader's barrier).
//   empty[s]       one per CTA; the leader's tcgen05.commit multicasts the arrival to both.
//   tmem_full[a]   one per CTA; multicast commit after the last k-block of a tile.
//   tmem_empty[a]  lives in the leader; 2 EW arrivals (EW = 4 or 8 epilogue warps x 2 CTAs, remote for r = 1).
#include "tc_common.cuh"
#include "tc_host.cuh"
#include "tc_pair.cuh"

using namespace tc;

namespace {

template <int BN, int STAGES, int EW = 4>
struct Tc2Cfg {
    static constexpr int THREADS = 64 + 32 * EW;              // producer warp, MMA warp, EW epilogue warps
    static constexpr int A_PLANE = 128 * 128;                 // 128 rows x 64 bf16
    static constexpr int B_PLANE = (BN / 2) * 128;            // this CTA's half of the weight tile
    static constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;
    static constexpr int TMEM_COLS = 2 * BN;                  // double-buffered 128 x BN accumulator per CTA
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024;
    static_assert(TMEM_COLS <= 512 && SMEM <= 227 * 1024, "resources");
};

struct Pair {                   // geometry of the 256-row pair tile
    int tiles_x2, tiles_y2;     // pair tiles per image
    int BW2, BH2;               // pair tile extent in pixels (BW2 * BH2 = 256)
    int dx, dy;                 // offset of rank 1's 128-pixel box inside the pair tile
};

template <int BN, int STAGES, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo, const TcParams p, const Pair g) {
    using Cfg = Tc2Cfg<BN, STAGES, EW>;
    constexpr int A_PLANE = Cfg::A_PLANE, B_PLANE = Cfg::B_PLANE, STAGE_BYTES = Cfg::STAGE_BYTES, BK = 64;
    SGAM_PDL_TRIGGER();                              // the next kernel may start launching; it waits for our completion itself
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float stat_s[EW][BN / 4 * 2];
    __shared__ __align__(16) EpiStageT<EW> epi_stage; // per-warp transpose tiles of the coalesced epilogue stores

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int num_kb = p.taps * p.kblocks_per_tap;
    const int total_tiles = p.tiles_m * p.tiles_n;              // pair tiles x column tiles

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 2 * EW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                          // barriers initialised and TMEM allocated in both CTAs
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    SGAM_PDL_WAIT();                                 // barriers / TMEM are set up; operands of the preceding kernel from here on

    if (warp == 0) {
        // ===== TMA producer (one thread per CTA): own A rows, own half of B; completion lands on the leader's barrier =====
        if (lane == 0) {
            const uint32_t tx_bytes = (p.nsplit == 3) ? STAGE_BYTES : (A_PLANE + B_PLANE);
            int kbg = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
                const int n0 = (tile % p.tiles_n) * BN + (int)rank * (BN / 2);
                int t = tile / p.tiles_n;
                const int tx = t % g.tiles_x2; t /= g.tiles_x2;
                const int ty = t % g.tiles_y2; const int b = t / g.tiles_y2;
                const int ox0 = tx * g.BW2 + (int)rank * g.dx, oy0 = ty * g.BH2 + (int)rank * g.dy;
                const int ab = p.a_batched ? b : 0, bb = p.b_batched ? b : 0;
                for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
                    const int s = kbg % STAGES, it = kbg / STAGES;
                    mbar_wait(&empty_bar[s], (it & 1) ^ 1);
                    uint8_t *st = smem + (size_t)s * STAGE_BYTES;
                    const uint32_t full0 = map_to_cta(&full_bar[s], 0);
                    const int tap = kb / p.kblocks_per_tap, kc = kb - tap * p.kblocks_per_tap;
                    const int kh = tap / p.ks, kw = tap - kh * p.ks;
                    const int cx = ox0 * p.stride + kw - p.pad + p.shift_x, cy = oy0 * p.stride + kh - p.pad + p.shift_y;
                    if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * tx_bytes);   // the leader alone expects both CTAs' bytes (the
                    tma2_load_4d(st, &mapA_hi, full0, kc * BK, cx, cy, ab);      // tx-count may go negative until it arrives)
                    tma2_load_3d(st + 2 * A_PLANE, &mapB_hi, full0, kb * BK, n0, bb);
                    if (p.nsplit == 3) {
                        tma2_load_4d(st + A_PLANE, &mapA_lo, full0, kc * BK, cx, cy, ab);
                        tma2_load_3d(st + 2 * A_PLANE + B_PLANE, &mapB_lo, full0, kb * BK, n0, bb);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the LEADER CTA drives both tensor cores =====
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc(256, BN);
            int kbg = 0, li = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++li) {
                const int acc = li & 1;
                mbar_wait(&tmem_empty_bar[acc], ((li >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < num_kb; ++kb, ++kbg) {
                    const int s = kbg % STAGES, it = kbg / STAGES;
                    mbar_wait(&full_bar[s], it & 1);
                    tc_fence_after();
                    uint8_t *st = smem + (size_t)s * STAGE_BYTES;
                    const uint64_t a_hi = make_smem_desc<128>(st), a_lo = make_smem_desc<128>(st + A_PLANE);
                    const uint64_t b_hi = make_smem_desc<128>(st + 2 * A_PLANE), b_lo = make_smem_desc<128>(st + 2 * A_PLANE + B_PLANE);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t off = (uint64_t)(k * 2);
                        umma2_bf16(tmem_d, a_hi + off, b_hi + off, idesc, (kb | k) ? 1u : 0u);
                        if (p.nsplit == 3) {
                            umma2_bf16(tmem_d, a_hi + off, b_lo + off, idesc, 1u);
                            umma2_bf16(tmem_d, a_lo + off, b_hi + off, idesc, 1u);
                        }
                    }
                    umma2_commit_multicast(&empty_bar[s]);        // frees this stage in both CTAs
                }
                umma2_commit_multicast(&tmem_full_bar[acc]);      // both epilogues may read their 128 rows
            }
        }
    } else {
        // ===== epilogue (both CTAs): 128 rows of the pair tile each; TMEM lane quadrant q = warp % 4, warp index w =====
        const int q = warp & 3, w = EW == 4 ? q : warp - 2;
        int li = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++li) {
            const int acc = li & 1;
            const int n_tile = tile % p.tiles_n, n0 = n_tile * BN;
            int t = tile / p.tiles_n;
            const int m2 = t % (g.tiles_x2 * g.tiles_y2);
            const int tx = t % g.tiles_x2; t /= g.tiles_x2;
            const int ty = t % g.tiles_y2; const int b = t / g.tiles_y2;
            const int ox0 = tx * g.BW2 + (int)rank * g.dx, oy0 = ty * g.BH2 + (int)rank * g.dy;
            mbar_wait(&tmem_full_bar[acc], (li >> 1) & 1);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            const long long per_img = p.stat_tiles ? p.stat_tiles : 2LL * g.tiles_x2 * g.tiles_y2;
            const long long slot = (long long)b * per_img + p.stat_tile0 + (long long)m2 * 2 + rank;
            epilogue_rows<BN, EW>(p, tmem_acc, q * 32 + lane, b, oy0, ox0, n0, n_tile, 0, slot, stat_s, epi_stage, w, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(map_to_cta(&tmem_empty_bar[acc], 0));
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                          // nobody exits while the peer may still signal / read it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
    }
}

template <int BN, int STAGES, int EW>
int launch2_cfg(const CUtensorMap &a_hi, const CUtensorMap &a_lo, const CUtensorMap &b_hi, const CUtensorMap &b_lo, const TcParams &p,
                const Pair &g, cudaStream_t s) {
    using Cfg = Tc2Cfg<BN, STAGES, EW>;
    static bool configured = false;
    if (!configured) {
        SGAM_CUDA_OK(cudaFuncSetAttribute(tc_gemm2_kernel<BN, STAGES, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        configured = true;
    }
    const int total = p.tiles_m * p.tiles_n, max_clusters = sm_count_cached() / 2;
    const int clusters = total < max_clusters ? total : max_clusters;
    SGAM_PDL_LAUNCH(SGAM_PDL_GEMM2, (tc_gemm2_kernel<BN, STAGES, EW>), 2 * clusters, Cfg::THREADS, Cfg::SMEM, s, a_hi, a_lo, b_hi, b_lo, p, g);
    return SGAM_OK;
}

}  // namespace

namespace tc {

// Does the 2-CTA kernel apply?  The image must tile exactly into 256-pixel pair tiles, the column count into BN, the
// output must be a plain row-major fp32 / split-bf16 tensor, and there must be enough pair tiles to fill the machine.
bool tc2_applicable(int B, int Ho, int Wo, int N, int stride, int out_nchw, int n_valid, int *BN_out) {
    if (out_nchw || n_valid != N || N % 128) return false;
    const bool exact = (Wo >= 256) ? (Wo % 256 == 0) : (Wo > 0 && 256 % Wo == 0 && (Wo & (Wo - 1)) == 0 && Ho % (256 / Wo) == 0 && Wo >= 2);
    if (!exact) return false;
    if (Wo < 256 && (256 / Wo) < 2) return false;
    const long long pair_tiles = (long long)B * ((long long)Ho * Wo / 256);
    const int BN = (N % 256 == 0 && pair_tiles * (N / 256) >= sm_count_cached() / 2) ? 256 : 128;
    if (pair_tiles * (N / BN) < sm_count_cached() / 2) return false;       // under-filled grids: 1-CTA kernel (64-wide tiles, split-K)
    (void)stride;
    *BN_out = BN;
    return true;
}

// p carries everything except the tiling; a/b maps must have been built with box rows 128 (A) and BN/2 (B).
int launch_tc2(int BN, const CUtensorMap &a_hi, const CUtensorMap &a_lo, const CUtensorMap &b_hi, const CUtensorMap &b_lo, TcParams p,
               int B, int Ho, int Wo, int N, cudaStream_t s) {
    Pair g;
    if (Wo >= 256) { g.BW2 = 256; g.BH2 = 1; g.dx = 128; g.dy = 0; p.BW = 128; p.BH = 1; }
    else { g.BW2 = Wo; g.BH2 = 256 / Wo; g.dx = 0; g.dy = g.BH2 / 2; p.BW = Wo; p.BH = g.BH2 / 2; }
    g.tiles_x2 = Wo / g.BW2; g.tiles_y2 = Ho / g.BH2;
    p.tiles_x = g.tiles_x2; p.tiles_y = g.tiles_y2;
    p.tiles_m = g.tiles_x2 * g.tiles_y2 * B;
    p.tiles_n = N / BN;
    p.ksplit = 1; p.kb_per_split = p.taps * p.kblocks_per_tap;
    // short K loops (1x1 convs, the fused q/k/v projection: <= 8 k-blocks) are bound by the epilogue: eight epilogue warps, and
    // one ring stage less to make room for their staging tiles (SGAM_TC_EPI8=0: the four-warp kernels everywhere)
    static int epi8 = -1;
    if (epi8 < 0) { const char *e = getenv("SGAM_TC_EPI8"); epi8 = e ? atoi(e) : 1; }
    if (epi8 && p.kb_per_split <= 8) {
        if (BN == 256) return launch2_cfg<256, 2, 8>(a_hi, a_lo, b_hi, b_lo, p, g, s);
        return launch2_cfg<128, 3, 8>(a_hi, a_lo, b_hi, b_lo, p, g, s);
    }
    if (BN == 256) return launch2_cfg<256, 3, 4>(a_hi, a_lo, b_hi, b_lo, p, g, s);
    return launch2_cfg<128, 4, 4>(a_hi, a_lo, b_hi, b_lo, p, g, s);
}

}  // namespace tc
