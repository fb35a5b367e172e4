// Stage (iii): fused single-head spatial self-attention of the 256-channel AttnBlocks (sm_100a).
//
//   reference: sgam/generative_sensing_module/modules/diffusionmodules/model.py:168-192
//       w_ = bmm(q^T, k) * C^-0.5 ; w_ = softmax(w_, dim=keys) ; h_ = bmm(v, w_^T)
//
// Round 1 ran this as three passes (QK^T GEMM -> [B,T,T] fp32 scores in HBM -> softmax pass -> split-bf16
// probabilities in HBM -> PV GEMM): ~2 GB of DRAM traffic per 4096-token block at 8 trajectories against ~134 MB of
// q / k / v / o, and a 1 GiB score matrix per block per frame at 512 x 512.  Here ONE kernel keeps the whole
// score tile on the SM ("flash" schedule): the scores S of 256 queries x 128 keys are a TMEM accumulator, the online
// softmax runs in registers, the probabilities P go back into the SAME tensor-memory columns as packed split-bf16 pairs
// and feed the P.V product as the MMA's A operand straight from TMEM, and the output accumulator O (256 queries x 256
// channels, fp32) stays in TMEM until the last key tile.  Nothing of size T x T ever exists.
//
// Precision: as everywhere in the tensor-core path, every fp32 product is three bf16 MMAs,
//   S = Qh.Kh + Qh.Kl + Ql.Kh          O += Ph.Vh + Ph.Vl + Pl.Vh        (x = xh + xl, xh = bf16(x), xl = bf16(x - xh))
// with fp32 accumulation; softmax statistics (running maximum, row sum) are fp32.
//
// Work decomposition: a cluster of two CTAs (cta_group::2) owns 256 consecutive queries of one image and streams all
// T keys in tiles of 128.  Per CTA: its 128 query rows of Q stay resident in shared memory for the whole tile
// (128 KB: 4 channel blocks x {hi, lo} x 16 KB, K-major SWIZZLE_128B); K and V^T arrive through a ring of six 16 KB
// slots -- a K slot is {hi, lo} of this CTA's HALF of the key tile for one 64-channel block, a V slot is one plane of
// this CTA's HALF of the channels for 64 keys -- so every operand byte is fetched once per pair and the pair's MMAs
// (M = 256) read the B operand from both shared memories.  TMEM per CTA: S0 | S1 | O = 128 + 128 + 256 columns.
//
// Roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only) + TMEM allocation, warps 2-5 softmax /
// correction / epilogue (thread = query row = TMEM lane).  Issue order of the leader:
//   S(0) S(1) | PV(0) S(2) | PV(1) S(3) | ...   -- S(j+2) overwrites the columns of P(j) and is ordered behind PV(j) by the
// in-order tensor pipe; the softmax of tile j overlaps PV(j-1) and S(j+1).
// The running maximum is "lazy": O and the row sum are rescaled only when a row's maximum grows by more than 2^8 over the
// reference it was accumulated with (P stays <= 256, exact in fp32 / split bf16), which makes the TMEM round trip of
// the correction rare; it happens behind pv_done(j-1), i.e. with no P.V product in flight.
//
// Barriers (parities derive from running counters, never from local phase bits):
//   full[s] / empty[s]   slot ring; full lives in the leader (expect_tx of both CTAs' bytes), empty is multicast
//   q_full / q_empty     the resident Q tile
//   s_full[b]            (per CTA, multicast commit) S(j) complete, b = j & 1
//   p_full[b]            (leader, 8 arrivals) all softmax warps of both CTAs stored P(j) (+ any O correction)
//   pv_done              (per CTA, multicast commit) PV(j) complete
//   o_free               (leader, 8 arrivals) the epilogue has read O; the next tile's PV(0) may overwrite it
#include "tc_common.cuh"
#include "tc_host.cuh"
#include "tc_pair.cuh"

using namespace tc;

namespace {

constexpr int AT_D = 256;                          // channels = head dimension
constexpr int AT_KB = AT_D / 64;                   // 64-channel blocks of the QK^T reduction
constexpr int AT_BN = 128;                         // keys per tile
constexpr int AT_SLOT = 16384;
constexpr int AT_SLOTS = 6;
constexpr int AT_Q_BYTES = AT_KB * 2 * 16384;      // resident Q: 4 x {hi, lo} x [128 rows x 128 B]
constexpr size_t AT_SMEM = (size_t)AT_Q_BYTES + (size_t)AT_SLOTS * AT_SLOT + 1024;
constexpr uint32_t AT_S0 = 0, AT_O = 256;          // TMEM columns: S0 = [0,128), S1 = [128,256), O = [256,512)
constexpr float AT_TAU = 8.0f;                     // lazy-rescale threshold (log2 units)
static_assert(AT_SMEM <= 227 * 1024 - 512, "shared memory budget");

struct AttnParams {
    int T, tiles_per_img, total_tiles, n_iter;     // total_tiles = work items = query tiles x key splits; n_iter = key tiles per item
    int nsp;                                       // key splits per query tile (1 = the item sees every key and writes o itself)
    float c1;                                      // C^-0.5 * log2(e): exp(s * scale - m) = exp2(s * c1 - m * log2 e)
    __nv_bfloat16 *o_hi, *o_lo;                    // [B, T, 256]
    float *o_part, *ml_part;                       // nsp > 1: un-normalised O [items][256][256] and (reference max, row sum) [items][256][2]
};

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
// tcgen05.wait::ld with the destination registers as in/out operands: nothing that consumes them can be scheduled above it
__device__ __forceinline__ void tmem_ld_fence(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
           "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
           "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
           "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
           "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// P chunk: 32 scores of this thread's row -> probabilities (reference maximum m, log2 domain) -> row-sum contribution
// and 16 + 16 packed split-bf16 words written into the tensor-memory columns of P_hi / P_lo
__device__ __forceinline__ float softmax_chunk(const uint32_t (&v)[32], float c1, float m, uint32_t t_hi, uint32_t t_lo) {
    uint32_t hi[16], lo[16];
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(v[j]), c1, -m));
        const float p1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), c1, -m));
        sum += p0 + p1;
        split2(p0, p1, hi[j / 2], lo[j / 2]);
    }
    tmem_st16(t_hi, hi);
    tmem_st16(t_lo, lo);
    return sum;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap mapQ_hi, const __grid_constant__ CUtensorMap mapQ_lo,
                const __grid_constant__ CUtensorMap mapK_hi, const __grid_constant__ CUtensorMap mapK_lo,
                const __grid_constant__ CUtensorMap mapV_hi, const __grid_constant__ CUtensorMap mapV_lo, const AttnParams p) {
    SGAM_PDL_TRIGGER();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *q_smem = smem;                                   // [kb][hi | lo][128 rows x 128 B]
    uint8_t *ring = smem + AT_Q_BYTES;
    __shared__ __align__(8) uint64_t full_bar[AT_SLOTS], empty_bar[AT_SLOTS], q_full, q_empty, s_full[2], p_full[2], pv_done, o_free;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int n = p.n_iter;

    if (threadIdx.x == 0) {
        for (int s = 0; s < AT_SLOTS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&q_full, 1); mbar_init(&q_empty, 1); mbar_init(&pv_done, 1); mbar_init(&o_free, 8);
        for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&p_full[b], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    SGAM_PDL_WAIT();

    if (warp == 0) {
        // ===== TMA producer (one thread per CTA): own Q rows, own half of every K / V tile; completion on the leader's barriers =====
        if (lane == 0) {
            uint32_t sc = 0;                                   // slots produced so far (ring position)
            int li = 0;
            auto acquire = [&]() -> uint8_t * {                // next slot, free; returns its address (barrier = full_bar[sc % SLOTS])
                const uint32_t s = sc % AT_SLOTS, use = sc / AT_SLOTS;
                mbar_wait(&empty_bar[s], (use & 1) ^ 1);
                if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * AT_SLOT);
                return ring + (size_t)s * AT_SLOT;
            };
            for (int item = cluster_id; item < p.total_tiles; item += num_clusters, ++li) {
                const int tile = item / p.nsp, j0 = (item - tile * p.nsp) * n;      // this item's first key tile
                const int b = tile / p.tiles_per_img, row0 = (tile - b * p.tiles_per_img) * 256 + (int)rank * 128;
                auto load_k = [&](int jj) {                    // keys [128 j + 64 rank, +64), channel block kb: {hi 8 KB | lo 8 KB}
                    const int j = j0 + jj;
                    for (int kb = 0; kb < AT_KB; ++kb) {
                        uint8_t *slot = acquire();
                        const uint32_t bar = map_to_cta(&full_bar[sc % AT_SLOTS], 0);
                        tma2_load_3d(slot, &mapK_hi, bar, kb * 64, j * AT_BN + (int)rank * 64, b);
                        tma2_load_3d(slot + AT_SLOT / 2, &mapK_lo, bar, kb * 64, j * AT_BN + (int)rank * 64, b);
                        ++sc;
                    }
                };
                auto load_v = [&](int jj) {                    // V^T rows (channels) [128 rank, +128), keys [128 j + 64 kb, +64): hi slot, lo slot
                    const int j = j0 + jj;
                    for (int kb = 0; kb < AT_BN / 64; ++kb) {
                        uint8_t *slot = acquire();
                        tma2_load_3d(slot, &mapV_hi, map_to_cta(&full_bar[sc % AT_SLOTS], 0), j * AT_BN + kb * 64, (int)rank * 128, b);
                        ++sc;
                        slot = acquire();
                        tma2_load_3d(slot, &mapV_lo, map_to_cta(&full_bar[sc % AT_SLOTS], 0), j * AT_BN + kb * 64, (int)rank * 128, b);
                        ++sc;
                    }
                };
                mbar_wait(&q_empty, (li & 1) ^ 1);             // the previous tile's last S product has read Q
                if (rank == 0) mbar_expect_tx(&q_full, 2 * AT_Q_BYTES);
                const uint32_t qbar = map_to_cta(&q_full, 0);
                for (int kb = 0; kb < AT_KB; ++kb) {
                    tma2_load_3d(q_smem + kb * 32768, &mapQ_hi, qbar, kb * 64, row0, b);
                    tma2_load_3d(q_smem + kb * 32768 + 16384, &mapQ_lo, qbar, kb * 64, row0, b);
                }
                load_k(0);
                if (n > 1) load_k(1);
                for (int j = 0; j < n; ++j) {
                    load_v(j);
                    if (j + 2 < n) load_k(j + 2);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the LEADER drives both tensor cores =====
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc_s = make_idesc(256, AT_BN), idesc_pv = make_idesc(256, AT_D);
            uint32_t sc = 0;                                   // slots consumed
            int g = 0, li = 0;                                 // key tiles issued before this query tile; query tiles done
            for (int item = cluster_id; item < p.total_tiles; item += num_clusters, ++li) {
                auto issue_s = [&](int j) {                    // S(j) = Q . K_j^T into S[(g + j) & 1]
                    const uint32_t tmem_s = tmem_base + AT_S0 + (uint32_t)(((g + j) & 1) * AT_BN);
                    for (int kb = 0; kb < AT_KB; ++kb) {
                        const uint32_t s = sc % AT_SLOTS, use = sc / AT_SLOTS;
                        mbar_wait(&full_bar[s], use & 1);
                        tc_fence_after();
                        const uint8_t *slot = ring + (size_t)s * AT_SLOT;
                        const uint64_t q_hi = make_smem_desc<128>(q_smem + kb * 32768), q_lo = make_smem_desc<128>(q_smem + kb * 32768 + 16384);
                        const uint64_t k_hi = make_smem_desc<128>(slot), k_lo = make_smem_desc<128>(slot + AT_SLOT / 2);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t off = (uint64_t)(ks * 2);
                            umma2_bf16(tmem_s, q_hi + off, k_hi + off, idesc_s, (kb | ks) ? 1u : 0u);
                            umma2_bf16(tmem_s, q_hi + off, k_lo + off, idesc_s, 1u);
                            umma2_bf16(tmem_s, q_lo + off, k_hi + off, idesc_s, 1u);
                        }
                        umma2_commit_multicast(&empty_bar[s]);
                        ++sc;
                    }
                    if (j == n - 1) umma2_commit_multicast(&q_empty);      // every product that reads this tile's Q has been issued
                    umma2_commit_multicast(&s_full[(g + j) & 1]);
                };
                auto issue_pv = [&](int j) {                   // O (+)= P(j) . V_j, P read from the tensor-memory columns of S[(g + j) & 1]
                    const int gi = g + j;
                    mbar_wait(&p_full[gi & 1], (gi >> 1) & 1);
                    if (j == 0 && li > 0) mbar_wait(&o_free, (li - 1) & 1);   // the previous tile's epilogue has drained O
                    tc_fence_after();
                    const uint32_t tmem_p = tmem_base + AT_S0 + (uint32_t)((gi & 1) * AT_BN), tmem_o = tmem_base + AT_O;
                    for (int kb = 0; kb < AT_BN / 64; ++kb) {
                        const uint32_t s_hi = sc % AT_SLOTS, s_lo = (sc + 1) % AT_SLOTS;
                        mbar_wait(&full_bar[s_hi], (sc / AT_SLOTS) & 1);
                        mbar_wait(&full_bar[s_lo], ((sc + 1) / AT_SLOTS) & 1);
                        tc_fence_after();
                        const uint64_t v_hi = make_smem_desc<128>(ring + (size_t)s_hi * AT_SLOT), v_lo = make_smem_desc<128>(ring + (size_t)s_lo * AT_SLOT);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t off = (uint64_t)(ks * 2);
                            const uint32_t a_hi = tmem_p + (uint32_t)(kb * 32 + ks * 8), a_lo = a_hi + 64;     // 16 keys = 8 packed columns
                            umma2_bf16_ts(tmem_o, a_hi, v_hi + off, idesc_pv, (j | kb | ks) ? 1u : 0u);
                            umma2_bf16_ts(tmem_o, a_hi, v_lo + off, idesc_pv, 1u);
                            umma2_bf16_ts(tmem_o, a_lo, v_hi + off, idesc_pv, 1u);
                        }
                        umma2_commit_multicast(&empty_bar[s_hi]);
                        umma2_commit_multicast(&empty_bar[s_lo]);
                        sc += 2;
                    }
                    umma2_commit_multicast(&pv_done);
                };
                mbar_wait(&q_full, li & 1);
                tc_fence_after();
                issue_s(0);
                if (n > 1) issue_s(1);
                for (int j = 0; j < n; ++j) {
                    issue_pv(j);
                    if (j + 2 < n) issue_s(j + 2);
                }
                g += n;
            }
        }
    } else {
        // ===== softmax / correction / epilogue (both CTAs): thread = query row = TMEM lane =====
        const int q = warp & 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const float c1 = p.c1;
        int g = 0, li = 0;
        for (int item = cluster_id; item < p.total_tiles; item += num_clusters, ++li) {
            const int tile = item / p.nsp;
            const int b = tile / p.tiles_per_img;
            const long long row = (long long)b * p.T + (long long)(tile - b * p.tiles_per_img) * 256 + (int)rank * 128 + q * 32 + lane;
            float m_used = 0.0f, l = 0.0f;                    // reference maximum (log2 domain) and row sum of exp2(s - m_used)
            const uint32_t t_o = tmem_base + lane_base + AT_O;
            for (int j = 0; j < n; ++j) {
                const int gi = g + j, bsel = gi & 1;
                const uint32_t t_s = tmem_base + lane_base + AT_S0 + (uint32_t)(bsel * AT_BN);
                mbar_wait(&s_full[bsel], (gi >> 1) & 1);
                tc_fence_after();
                uint32_t v0[32], v1[32], v2[32], v3[32];
                tmem_ld32_nowait(t_s, v0); tmem_ld32_nowait(t_s + 32, v1); tmem_ld32_nowait(t_s + 64, v2); tmem_ld32_nowait(t_s + 96, v3);
                tmem_ld_fence(v0); tmem_ld_fence(v1); tmem_ld_fence(v2); tmem_ld_fence(v3);
                float mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    mx = fmaxf(mx, fmaxf(fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])), fmaxf(__uint_as_float(v2[i]), __uint_as_float(v3[i]))));
                mx *= c1;                                      // c1 > 0
                bool need = false;
                if (j == 0) m_used = mx;                       // O is (re)initialised by PV(0): nothing to correct
                else need = mx > m_used + AT_TAU;
                // pv_done is observed EXACTLY once per key tile and always before this warp's p_full arrive: a parity wait
                // is only sound while the waiter is at most one phase behind and never ahead.  PV(j-2) is complete here
                // (S(j) was issued behind it), PV(j) cannot be issued before the arrive below.
                bool seen_pv = (j == 0);
                if (__any_sync(0xffffffffu, need)) {           // rare: rescale O and the row sum to the new reference maximum
                    const float m_new = need ? mx : m_used;
                    const float f = ex2_approx(m_used - m_new);     // exactly 1 for the rows that keep their reference
                    mbar_wait(&pv_done, (gi - 1) & 1);         // PV(j-1) complete: no P.V product is in flight
                    seen_pv = true;
                    tc_fence_after();
#pragma unroll 1
                    for (int c = 0; c < AT_D; c += 32) {
                        uint32_t o[32];
                        tmem_ld32(t_o + (uint32_t)c, o);
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                        tmem_st32(t_o + (uint32_t)c, o);
                    }
                    l *= f;
                    m_used = m_new;
                }
                // P(j) over S(j)'s own columns: keys 2c, 2c+1 -> packed column c of P_hi ([0,64)) / P_lo ([64,128))
                l += softmax_chunk(v0, c1, m_used, t_s, t_s + 64);
                l += softmax_chunk(v1, c1, m_used, t_s + 16, t_s + 80);
                l += softmax_chunk(v2, c1, m_used, t_s + 32, t_s + 96);
                l += softmax_chunk(v3, c1, m_used, t_s + 48, t_s + 112);
                tmem_st_wait();
                if (!seen_pv) mbar_wait(&pv_done, (gi - 1) & 1);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_cta(&p_full[bsel], 0));
            }
            // ---- epilogue: O / l -> split bf16 [B, T, 256]; a thread writes 64 contiguous bytes per plane and chunk ----
            mbar_wait(&pv_done, (g + n - 1) & 1);
            tc_fence_after();
            if (p.nsp > 1) {
                // key-split item: un-normalised O and (m, l) to the workspace; attn_combine_kernel merges the splits
                const long long prow = (long long)item * 256 + (int)rank * 128 + q * 32 + lane;
                float *dp = p.o_part + prow * AT_D;
#pragma unroll 1
                for (int c = 0; c < AT_D; c += 32) {
                    uint32_t o[32];
                    tmem_ld32(t_o + (uint32_t)c, o);
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4 *>(dp + c + i) = make_float4(__uint_as_float(o[i]), __uint_as_float(o[i + 1]), __uint_as_float(o[i + 2]), __uint_as_float(o[i + 3]));
                }
                *reinterpret_cast<float2 *>(p.ml_part + prow * 2) = make_float2(m_used, l);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_cta(&o_free, 0));
                g += n;
                continue;
            }
            const float inv = 1.0f / l;
            __nv_bfloat16 *dh = p.o_hi + row * AT_D, *dl = p.o_lo + row * AT_D;
#pragma unroll 1
            for (int c = 0; c < AT_D; c += 32) {
                uint32_t o[32], hi[16], lo[16];
                tmem_ld32(t_o + (uint32_t)c, o);
#pragma unroll
                for (int i = 0; i < 32; i += 2) split2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv, hi[i / 2], lo[i / 2]);
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    *reinterpret_cast<uint4 *>(dh + c + 2 * i) = make_uint4(hi[i], hi[i + 1], hi[i + 2], hi[i + 3]);
                    *reinterpret_cast<uint4 *>(dl + c + 2 * i) = make_uint4(lo[i], lo[i + 1], lo[i + 2], lo[i + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(map_to_cta(&o_free, 0));
            g += n;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                      // nobody exits while the peer may still signal / read it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// Merge of the key splits (nsp > 1): o = sum_s 2^(m_s - M) O_s / sum_s 2^(m_s - M) l_s with M = max_s m_s, split to bf16 planes.
// One warp per query row, 8 channels per lane; partial row of split s of query tile t: ((t * nsp + s) * 256 + row in tile).
__global__ void __launch_bounds__(256)
attn_combine_kernel(const float *__restrict__ o_part, const float *__restrict__ ml_part, int nsp, long long rows,
                    __nv_bfloat16 *__restrict__ o_hi, __nv_bfloat16 *__restrict__ o_lo) {
    SGAM_PDL_PROLOGUE();
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const long long tile = row >> 8, r = row & 255;
    float M = -INFINITY;
    for (int s = 0; s < nsp; ++s) M = fmaxf(M, __ldg(ml_part + ((tile * nsp + s) * 256 + r) * 2));
    float L = 0.0f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < nsp; ++s) {
        const long long pr = (tile * nsp + s) * 256 + r;
        const float2 ml = __ldg(reinterpret_cast<const float2 *>(ml_part + pr * 2));
        const float w = ex2_approx(ml.x - M);
        L = fmaf(w, ml.y, L);
        const float4 a = __ldg(reinterpret_cast<const float4 *>(o_part + pr * AT_D) + 2 * lane), b4 = __ldg(reinterpret_cast<const float4 *>(o_part + pr * AT_D) + 2 * lane + 1);
        acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
        acc[4] = fmaf(w, b4.x, acc[4]); acc[5] = fmaf(w, b4.y, acc[5]); acc[6] = fmaf(w, b4.z, acc[6]); acc[7] = fmaf(w, b4.w, acc[7]);
    }
    const float inv = 1.0f / L;
    uint32_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) split2(acc[2 * k] * inv, acc[2 * k + 1] * inv, h[k], l[k]);
    reinterpret_cast<uint4 *>(o_hi + row * AT_D)[lane] = make_uint4(h[0], h[1], h[2], h[3]);
    reinterpret_cast<uint4 *>(o_lo + row * AT_D)[lane] = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace

// 1 if the fused kernel handles this shape: 256 channels, a token count that tiles into 256-query pair tiles.
extern "C" int sgam_attention_tc_supported(int B, int T, int C) { return B > 0 && C == AT_D && T >= 256 && T % 256 == 0; }

// Key splits the fused kernel wants for this batch: 1 when the query tiles alone fill the SM pairs, otherwise the largest of
// {2, 4, 8} that divides the key-tile count, leaves >= 2 key tiles per item and does not overshoot the machine.
extern "C" int sgam_attention_tc_splits(int B, int T) {
    const long long tiles = (long long)B * (T / 256);
    const int n_iter = T / AT_BN, pairs = sm_count_cached() / 2;
    if (tiles * 3 >= pairs * 2) return 1;
    int best = 1;
    for (int s = 2; s <= 8; s *= 2)
        if (n_iter % s == 0 && n_iter / s >= 2 && tiles * s <= pairs + pairs / 4) best = s;
    return best;
}
extern "C" size_t sgam_attention_tc_workspace_bytes(int B, int T, int kv_splits) {
    return kv_splits > 1 ? (size_t)kv_splits * B * T * (AT_D + 2) * sizeof(float) : 0;
}

extern "C" int sgam_attention_tc(const void *q_hi, const void *q_lo, const void *k_hi, const void *k_lo, const void *vt_hi,
                                 const void *vt_lo, void *o_hi, void *o_lo, int B, int T, int C, float scale, int kv_splits,
                                 void *workspace, int ld_qk, void *stream) {
    SGAM_REQUIRE(q_hi && q_lo && k_hi && k_lo && vt_hi && vt_lo && o_hi && o_lo, "attention_tc: null pointer");
    SGAM_REQUIRE(sgam_attention_tc_supported(B, T, C), "attention_tc: needs C == 256 and T %% 256 == 0 (B=%d T=%d C=%d)", B, T, C);
    SGAM_REQUIRE(scale > 0.0f, "attention_tc: scale must be positive");
    SGAM_REQUIRE(kv_splits >= 1 && (T / AT_BN) % kv_splits == 0 && (kv_splits == 1 || workspace), "attention_tc: kv_splits %d must divide the %d key tiles (and needs a workspace)", kv_splits, T / AT_BN);
    SGAM_REQUIRE(ld_qk == 0 || (ld_qk >= C && ld_qk % 8 == 0), "attention_tc: ld_qk %d must be 0 (dense) or a multiple of 8 >= C", ld_qk);
    CUtensorMap mq_hi, mq_lo, mk_hi, mk_lo, mv_hi, mv_lo;
    const long long qk_dims[3] = {C, T, B}, v_dims[3] = {T, C, B};
    const int q_box[3] = {64, 128, 1}, k_box[3] = {64, 64, 1}, v_box[3] = {64, 128, 1};
    int rc;
    if ((rc = make_map(&mq_hi, q_hi, 3, qk_dims, q_box, nullptr, ld_qk)) || (rc = make_map(&mq_lo, q_lo, 3, qk_dims, q_box, nullptr, ld_qk)) ||
        (rc = make_map(&mk_hi, k_hi, 3, qk_dims, k_box, nullptr, ld_qk)) || (rc = make_map(&mk_lo, k_lo, 3, qk_dims, k_box, nullptr, ld_qk)) ||
        (rc = make_map(&mv_hi, vt_hi, 3, v_dims, v_box)) || (rc = make_map(&mv_lo, vt_lo, 3, v_dims, v_box)))
        return rc;
    AttnParams p;
    p.T = T; p.tiles_per_img = T / 256; p.nsp = kv_splits; p.total_tiles = B * (T / 256) * kv_splits; p.n_iter = T / AT_BN / kv_splits;
    p.c1 = scale * 1.4426950408889634f;
    p.o_hi = (__nv_bfloat16 *)o_hi; p.o_lo = (__nv_bfloat16 *)o_lo;
    p.o_part = (float *)workspace;
    p.ml_part = p.o_part ? p.o_part + (size_t)kv_splits * B * T * AT_D : nullptr;
    static bool configured = false;
    if (!configured) {
        SGAM_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_SMEM));
        configured = true;
    }
    const int max_clusters = sm_count_cached() / 2;
    const int clusters = p.total_tiles < max_clusters ? p.total_tiles : max_clusters;
    SGAM_PDL_LAUNCH(SGAM_PDL_GEMM2, attn_fwd_kernel, 2 * clusters, TC_THREADS, AT_SMEM, (cudaStream_t)stream, mq_hi, mq_lo, mk_hi, mk_lo, mv_hi,
                    mv_lo, p);
    if (kv_splits > 1) {
        const long long rows = (long long)B * T;
        SGAM_PDL_LAUNCH(SGAM_PDL_MISC, attn_combine_kernel, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, p.o_part, p.ml_part, kv_splits, rows,
                        p.o_hi, p.o_lo);
    }
    return SGAM_OK;
}
