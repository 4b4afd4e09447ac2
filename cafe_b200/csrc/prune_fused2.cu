// prune_fused2.cu — K2, second generation of the fused persistent pruning kernel (the default path).
//
// Replaces, per objective evaluation, the reference's F x (2n-2) calls of square_matrix_multiply
// (libtree/birthdeath.c:163-182) under compute_internal_node_likelihood (cafe/cafe_tree.c:226-271),
// initialize_leaf_likelihoods (:191-211) and compute_posterior's root reduction (cafe/lambda.cpp:657-689).
//
// Why a second kernel: prune_fused.cu runs 3 independent CTAs per SM.  The SM's warp arbiter serves the
// highest warp slot first, so one of the three CTAs is starved of the DMMA pipe and finishes ~20 % after the
// other two (profiles/r1_fused_v1_cta_timeline.txt), and every CTA leaves the pipe for its epilogues (global
// gathers), leaf-pair products and proxy fences.  Here ONE CTA per SM owns the SM, 12 warps:
//
//   warps 4..11  DMMA consumers, 2 M-groups x 4 N-warps (warp tile 48 families x 32 sizes), CTA tile 96 families x 128
//                sizes: exactly two DMMA warps per SM sub-partition, one of each group.  All eight consume the same
//                shared-memory ring (2 stages of 2 K blocks), so the matrix tile (B) is fetched once per CTA.  Consumers
//                touch shared memory only: no global loads, no global stores, no membar.  Fragment loads go through
//                32-bit shared addresses with immediate offsets (a ring-stage boundary is ~30 integer instructions).
//   warp 0       TMA producer (one lane): child vectors (A, 96 x 16 sizes) and matrix K-blocks (B, 128 rows x 16 sizes;
//                the B maps end at the op's last size, so TMA zero-fills the rows of sizes that do not exist).
//   warps 1..2   leaf-pair gatherers: a node whose two children are leaves has the vector M_a[.][c_a] * M_b[.][c_b]
//                (cafe_tree.c:204-210); the gatherers write these vectors one pair of tiles ahead into scratch slots of
//                their own, the parent's GEMM streams them like any other vector.  One progress counter per gatherer.
//   warp 3       epilogue manager: stages the sibling factor of the next pass in the 96 KB C tile (TMA for a stored
//                partial product, cp.async row gathers for a leaf sibling), and writes the finished tile back with
//                TMA stores.  Consumers multiply in place (C = acc * C).  All fences live in this warp.
//
// Round-2 history (profiles/r2_k2_experiments.md): per-group C halves with the groups one ring stage apart (3.82 ms) and a
// register epilogue straight to global memory (5.0 ms) lost against 3.72 ms for this structure; the lean stage boundary, the
// sigma permutation of the matrix-tile rows (no lane-dependent access order in the C tile), compile-time epilogue
// instantiations and the clipped B maps then took it to 3.58 ms at configs[1] and 132.8 ms (0.93 of cuBLAS DGEMM) at
// configs[2].  k_prune_fused2<.., WIN = true> is the windowed instantiation for the conditional distribution and the p-values:
// per-family column windows, and every role ends a tile's K loops and output passes at the tile's largest window and runs the
// root op only over the passes somebody reads (tile_win_*); its callers order the families so that a tile is homogeneous and
// every CTA gets the same mix (fused2_tile_slots).  The score instantiation carries none of this.
//
// Bit-for-bit the same arithmetic as prune_fused.cu / prune.cu (same DMMA order over K, one rounding per product).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <type_traits>

#include "common.cuh"

// Timing ablations (results are garbage), compiled in only for A/B builds (tools/k2_variants.sh "-DCAFE_K2_ABLATE"): the hot loops
// of the product build carry no test of Params::dbg.  CAFE_GPU_DBG bits: 1 no epilogue work, 2 no epilogue at all, 4 no store.
// -DCAFE_K2_NOSYNC: no ring at all (K loops on whatever is in shared memory, no producer).
#ifdef CAFE_K2_ABLATE
#define K2_DBG(bits) ((P.dbg & (bits)) != 0)
#else
#define K2_DBG(bits) false
#endif
#ifdef CAFE_K2_NOSYNC
#define K2_DBG_NOSYNC true
#else
#define K2_DBG_NOSYNC false
#endif

namespace fused2 {

constexpr int GM = 2;                  // M-groups (consumer warpgroups): exactly two DMMA warps per SM sub-partition
constexpr int HM = 48;                 // families per group
constexpr int TILE_M = GM * HM;        // 96 families per CTA tile
constexpr int TN = 128;                // output sizes per pass
constexpr int BK = 16;                 // sizes per K block (16 doubles = 128 B = one swizzle row)
constexpr int NSTAGE = 2;                // ring stages, each KB_PER_STAGE K blocks
constexpr int KB_PER_STAGE = 2;
constexpr int A_BYTES = TILE_M * 128;  // 12 KB
constexpr int B_BYTES = TN * 128;      // 16 KB
constexpr int SUB_BYTES = A_BYTES + B_BYTES;        // one K block: A | B
constexpr int STAGE_BYTES = KB_PER_STAGE * SUB_BYTES;
constexpr int C_BOX_BYTES = TILE_M * 128;           // one box: 96 families x 16 sizes
constexpr int C_BOXES = TN / BK;                    // 8
constexpr int C_BYTES = C_BOXES * C_BOX_BYTES;      // 96 KB
constexpr int N_CONSUMER_WARPS = 4 * GM;
constexpr int THREADS = (N_CONSUMER_WARPS + 4) * 32;
constexpr int MB = HM / 8, NB = 4, WCOLS = NB * 8;
constexpr int N_GATHER_WARPS = 2;
static_assert(N_GATHER_WARPS == 2, "producer_main waits for exactly two gatherer counters");
// The helper warps are warps 0..3 and the DMMA warps 4..11: the sub-partition arbiter prefers the highest warp id, so the
// rarely-ready helpers never take an issue slot from a DMMA warp that is ready.
constexpr int N_AUX_WARPS = 4;
// register re-partition of the 384 x 168 launch allocation: 8*32*192 + 4*32*120 = 64512
constexpr int REGS_CONSUMER = 192, REGS_AUX = 120;

struct Op {              // one GEMM of the post-order schedule (a tree edge below an internal node)
    int is_root;
    int key;             // matrix of the GEMM child's branch
    int a_kind;          // 0: child vector in slot in_slot, 1: child is a leaf pair (a1, a2), its vector in leaf-pair slot in_slot
    int in_slot, out_slot;
    int leaf_a1, key_a1, leaf_a2, key_a2;
    int other_kind;      // 0 none, 1 leaf sibling, 2 multiply into out_slot
    int leaf_o, key_o;
};

struct Params {
    const Op* ops;
    int n_ops, n_slots, n_cherry;   // per CTA and tile: n_slots node-vector slots; per CTA, pair parity and tile: n_cherry leaf-pair slots
    int F, F_pad;
    int W, R, root_min;
    int Sp, Vp;
    int n_mblocks;               // ceil(F / 8)
    double* scratch;             // [grid][cta_rows][Vp]
    const double* MT;            // [D][Sp][Sp] transposed matrices
    const int* counts;           // [n_leaves][leaf_stride]: observed (or simulated) size of family f at leaf k = counts[k * leaf_stride + f]
    size_t leaf_stride;
    // Windowed mode (conditional distribution / p-values): per-family column window, cafe_family.c:250-254 and
    // conditional_distribution.cpp:29.  Sizes above colmax[f] are exact zeros in every node vector of family f (the reference
    // does not compute them) and a leaf whose size lies above the window has factor 0.  nullptr: the full range for everybody.
    const int* colmax;           // nullable, [F_pad]
    const int* root_pick;        // nullable, [F_pad]: root SIZE whose likelihood goes to L0_out[f] (root range {s} of the distribution)
    double* L0_out;              // [F] with root_pick
    const int* root_need;        // nullable, [F_pad]: root rows 0 .. root_need[f]-1 are all a caller reads of family f (p-values: the forced root range)
    const double* logprior;      // [R]
    const double* prior_mant;    // [R] prior = mant * 2^exp, mant in [1,2)  (host frexp; exp = -2^30 where the prior is 0)
    const int* prior_exp;        // [R]
    double* logpost;             // [F_pad]; nullptr: no posterior reduction at the root (windowed mode)
    double* maxlik;
    int* argmax;
    double* Lroot_out;           // nullable, [F][R]
    int dbg;                     // timing ablations, honoured by -DCAFE_K2_ABLATE builds only (see K2_DBG above)
    long long* cta_times;        // nullable debug: [grid][4] = smid, start ns, end ns, 8-family blocks
    long long* warp_prof;        // nullable debug: CTA 0, [16 warps][8] cycle sums (see consumer_main / producer_main)
    long long* timeline;         // nullable debug: CTA 0, warps 0 and 4 (one sub-partition): [2][1024 items][4] clock stamps
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// Same for the helper warps, which are never latency critical: a long suspend-time hint keeps them asleep in hardware instead
// of re-polling every few dozen cycles next to the DMMA warps of their SM sub-partition.
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(20000u)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"((uint64_t)map), "r"(c0), "r"(c1), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy accesses <-> async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16-byte / 8-byte asynchronous copies with zero fill of the bytes beyond src_bytes
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival once all cp.async of this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int OPFLAGS_CAP = 1024;
struct Ctl {
    uint64_t full[NSTAGE];       // producer expect_tx
    uint64_t empty[NSTAGE];      // 8 consumer warps
    uint64_t c_ready;            // 32 lanes of the epilogue manager (+ TMA bytes)
    uint64_t c_done;             // 8 consumer warps
    volatile int done[2];        // finished (stored, visible) ops per tile of the pair
    volatile int cherry_count[N_GATHER_WARPS];  // leaf-pair vectors finished by each gatherer warp (its half of the rows), in (pair, op, tile)
                                 // order.  One counter per warp: a sum would let one warp's lead stand in for the other's lag (seen at
                                 // pair 0 on a cold context, where the producer is right behind the gatherers)
    int rowoff_o[TILE_M];        // epilogue manager: count * Sp of the leaf sibling, per tile row (-1: leaf outside the window)
    int colmax[2][TILE_M];       // windowed mode: column window of the rows of the two tiles of the pair
    int pick[2][TILE_M];         // windowed mode: root row index to extract (-1 none)
    unsigned char opflags[OPFLAGS_CAP];  // per op: bit0 is_root, bit1 a_kind, bits 2-3 other_kind (what the consumers need)
    double red_ml[GM][4][HM];    // root reduction across the 4 N-warps of a group
    double red_mp[GM][4][HM];
    int red_am[GM][4][HM];
};
__device__ __forceinline__ void group_bar(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

// This CTA's families: a contiguous range of 8-family blocks cut into an even number of tiles of <= 12 blocks;
// tiles are processed two at a time (the op sequence of one interleaved with the other's) so that a vector is never
// streamed right after it was stored.  Inside a tile, group g owns mbv(g) consecutive blocks.
struct TilePlan {
    int mb_lo, n_mb, n_tiles, n_pairs;
    __device__ TilePlan(const Params& P) {
        const int G = gridDim.x, c = blockIdx.x;
        mb_lo = (int)((long long)P.n_mblocks * c / G);
        n_mb = (int)((long long)P.n_mblocks * (c + 1) / G) - mb_lo;
        n_tiles = (n_mb + TILE_M / 8 - 1) / (TILE_M / 8);
        if (n_mb >= 2 && (n_tiles & 1)) ++n_tiles;
        n_pairs = (n_tiles + 1) / 2;
    }
    // Tile sizes are even wherever possible (both groups of a tile then hold the same number of blocks and finish every ring
    // stage together): all tiles get the even base size e, the first ones 2 more, and one tile the odd block if n_mb is odd.
    __device__ bool tile(int t, int& mb0, int& m) const {
        if (t >= n_tiles) return false;
        const int e = (n_mb / n_tiles) & ~1;
        const int r = n_mb - e * n_tiles;   // < 2 * n_tiles
        const int n2 = r >> 1;              // tiles with e + 2 blocks
        auto size = [&](int i) { return e + (i < n2 ? 2 : 0) + ((r & 1) && i == n2 ? 1 : 0); };
        mb0 = mb_lo + e * t + 2 * min(t, n2) + ((r & 1) && t > n2 ? 1 : 0);
        m = size(t);
        return true;
    }
    __device__ static int mbv(int m, int g) { return (m + GM - 1 - g) / GM; }
    __device__ static int pre(int m, int g) { return g == 0 ? 0 : mbv(m, 0); }
    // family of tile row r (clamped into [0, F) so that gathers of unused rows stay in bounds)
    // family of tile row r, or -1 for a row without a family
    __device__ static int family_or_neg(int mb0, int m, int r, int F) {
        const int g = r / HM, lr = r - g * HM;
        const int f = (mb0 + pre(m, g)) * 8 + lr;
        return (lr < mbv(m, g) * 8 && f < F) ? f : -1;
    }
    __device__ static int family(int mb0, int m, int r, int F) {
        const int g = r / HM, lr = r - g * HM;
        const int f = (mb0 + pre(m, g)) * 8 + lr;
        return (lr < mbv(m, g) * 8 && f < F) ? f : (F - 1);
    }
};

// Scratch rows of one CTA: [2 tiles][n_slots] node vectors, then [2 pair parities][2 tiles][n_cherry] leaf-pair vectors, 96 rows each.
__device__ __forceinline__ int cta_rows(const Params& P) { return (2 * P.n_slots + 4 * P.n_cherry) * TILE_M; }
__device__ __forceinline__ int cherry_row(const Params& P, int pair, int h, int c) {
    return (2 * P.n_slots + ((pair & 1) * 2 + h) * P.n_cherry + c) * TILE_M;
}

__device__ __forceinline__ void advance(uint32_t& stage, uint32_t& phase) {
    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
}

// Windowed mode: 1 + the largest column window among the families of a tile.  No node vector of the tile has a non-zero size at or
// above it (Params::colmax), so every K loop of the tile ends there and output sizes from there on are never computed: the
// producer, the epilogue manager and the consumers all derive the tile's pass and K-block counts from this one number.  The rows
// of the conditional distribution are ordered by root size, so the families of a tile have nearly the same window.
// The same goes for the root rows: of the R root sizes the conditional distribution reads ONE per family (Params::root_pick)
// and the p-values the first root_need[f]; the root op of a tile runs only the 128-row passes [rch_lo, rch_hi) that hold a row
// somebody reads.
struct TileWin {
    int wmax;            // K extent and, below the root, number of output sizes
    int rch_lo, rch_hi;  // passes of the root op
};
__device__ __forceinline__ void tilewin_row(const Params& P, int f, int& w, int& lo, int& hi) {
    w = max(w, __ldg(P.colmax + f));
    if (P.root_pick) { const int r = __ldg(P.root_pick + f) - P.root_min; lo = min(lo, r); hi = max(hi, r); }
    else if (P.root_need) { lo = 0; hi = max(hi, __ldg(P.root_need + f) - 1); }
}
__device__ __forceinline__ TileWin tilewin_finish(const Params& P, int w, int lo, int hi) {
    TileWin t;
    t.wmax = min(P.W, w + 1);
    const int n_chunks = (P.R + TN - 1) / TN;
    t.rch_lo = 0; t.rch_hi = n_chunks;
    if (P.root_pick || P.root_need) {
        lo = max(0, min(lo, P.R - 1)); hi = max(lo, min(hi, P.R - 1));
        t.rch_lo = lo / TN; t.rch_hi = hi / TN + 1;
    }
    return t;
}
__device__ __forceinline__ TileWin tile_win_lane(const Params& P, int mb0, int m) {
    int w = 0, lo = 0x7fffffff, hi = 0;
    for (int r = 0; r < TILE_M; ++r) {
        const int f = TilePlan::family_or_neg(mb0, m, r, P.F);
        if (f >= 0) tilewin_row(P, f, w, lo, hi);
    }
    return tilewin_finish(P, w, lo, hi);
}
__device__ __forceinline__ TileWin tile_win_warp(const Params& P, int mb0, int m, int lane) {
    int w = 0, lo = 0x7fffffff, hi = 0;
    for (int r = lane; r < TILE_M; r += 32) {
        const int f = TilePlan::family_or_neg(mb0, m, r, P.F);
        if (f >= 0) tilewin_row(P, f, w, lo, hi);
    }
    w = __reduce_max_sync(0xffffffffu, w); lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
    return tilewin_finish(P, w, lo, hi);
}

// ================================ TMA producer (one lane) ================================ ================================
template <bool PROF, bool WIN>
__device__ __forceinline__ void producer_main(const CUtensorMap* tmA, const CUtensorMap* tmB, const CUtensorMap* tmBroot, const Params& P,
                                              unsigned char* stage_base, Ctl* ctl) {
    const TilePlan plan(P);
    const int scratch_row0 = blockIdx.x * cta_rows(P);
    const int n_kblocks_full = (P.W + BK - 1) / BK;
    if (K2_DBG_NOSYNC) return;
    uint32_t stage = 0, phase = 0;
    int ops_done_base = 0;
    const bool prof = PROF && P.warp_prof != nullptr && blockIdx.x == 0;
    long long t_wait_done = 0, t_wait_empty = 0;
    const long long t_begin = prof ? clock64() : 0;
    int cherry_units = 0;  // leaf-pair vectors needed so far, in the gatherers' order (pair, op, tile)
    int p_item = 0;
    for (int pair = 0; pair < plan.n_pairs; ++pair) {
        int wmax_h[2] = {P.W, P.W}, rlo_h[2] = {0, 0}, rhi_h[2] = {0, 0};
        if (WIN) {
            for (int h = 0; h < 2; ++h) {
                int mb0, m;
                if (plan.tile(2 * pair + h, mb0, m)) { const TileWin tw = tile_win_lane(P, mb0, m); wmax_h[h] = tw.wmax; rlo_h[h] = tw.rch_lo; rhi_h[h] = tw.rch_hi; }
            }
        }
        for (int oi = 0; oi < P.n_ops; ++oi) {
            const Op op = P.ops[oi];
            const int r0 = op.is_root ? P.root_min : 0;
            const int nrows_full = op.is_root ? P.R : P.W;
            const int n_chunks_full = (nrows_full + TN - 1) / TN;
            for (int h = 0; h < 2; ++h) {
                if (2 * pair + h >= plan.n_tiles) continue;
                const int n_chunks = WIN ? (op.is_root ? rhi_h[h] : (wmax_h[h] + TN - 1) / TN) : n_chunks_full;
                const int ch_lo = (WIN && op.is_root) ? rlo_h[h] : 0;
                const int n_kblocks = WIN ? (wmax_h[h] + BK - 1) / BK : n_kblocks_full;
                const long long t_op = prof ? clock64() : 0;
                if (op.a_kind == 0) {
                    // the vector to stream was stored by an earlier op of this tile: wait until it is visible
                    const long long t0 = prof ? clock64() : 0;
                    while (!K2_DBG(2) && ctl->done[h] < ops_done_base + oi) { __nanosleep(20); }
                    __threadfence_block();
                    fence_proxy_async();
                    if (prof) t_wait_done += clock64() - t0;
                } else {
                    // a leaf-pair vector: written by the two gatherers, normally a whole pair of tiles ahead
                    ++cherry_units;
                    while (ctl->cherry_count[0] < cherry_units || ctl->cherry_count[1] < cherry_units) { __nanosleep(20); }
                    __threadfence_block();
                    fence_proxy_async();
                }
                const int a_row = scratch_row0 + (op.a_kind == 0 ? (h * P.n_slots + op.in_slot) * TILE_M : cherry_row(P, pair, h, op.in_slot));
                for (int ch = ch_lo; ch < n_chunks; ++ch) {
                    const long long t_item = prof ? clock64() : 0;
                    long long t_first = 0;
                    for (int kb = 0; kb < n_kblocks; kb += KB_PER_STAGE) {
                        const long long t0 = prof ? clock64() : 0;
                        mbar_wait_sleepy(&ctl->empty[stage], phase ^ 1);
                        if (prof) { t_wait_empty += clock64() - t0; if (kb == 0) t_first = clock64(); }
                        const int nsub = min(KB_PER_STAGE, n_kblocks - kb);
                        mbar_arrive_expect_tx(&ctl->full[stage], nsub * SUB_BYTES);
                        for (int j = 0; j < nsub; ++j) {
                            unsigned char* sA = stage_base + stage * STAGE_BYTES + j * SUB_BYTES;
                            tma_load_2d(sA, tmA, (kb + j) * BK, a_row, &ctl->full[stage]);
                            tma_load_3d(sA + A_BYTES, op.is_root ? tmBroot : tmB, (kb + j) * BK, r0 + ch * TN, op.key, &ctl->full[stage]);
                        }
                        advance(stage, phase);
                    }
                    if (prof && P.timeline && p_item < 1024) {
                        long long* tl = P.timeline + ((size_t)2 * 1024 + p_item) * 4;
                        tl[0] = t_op; tl[1] = t_item; tl[2] = t_first; tl[3] = clock64();
                    }
                    ++p_item;
                }
            }
        }
        ops_done_base += P.n_ops;
    }
    if (prof) {
        long long* o = P.warp_prof + N_CONSUMER_WARPS * 8;
        o[0] = clock64() - t_begin; o[1] = t_wait_done; o[2] = t_wait_empty;
    }
}

// ================================ leaf-pair gatherers (2 warps) ================================
// A node whose two children are leaves has the vector L[j] = M_a[j][count_a] * M_b[j][count_b] (cafe_tree.c:204-210 twice, then
// the product of :266-270): two gathered rows of the transposed matrices.  The gatherers write these vectors for the NEXT pair
// of tiles into dedicated scratch slots while the DMMA warps work on the current pair - a whole pair of tiles (milliseconds)
// of slack, no coupling to the ring, one product per element.  The parent's GEMM then streams the slot like any other vector.
// Gatherer gi owns the tile rows [48 gi, 48 gi + 48); a warp handles two rows at a time, lanes along the sizes.
template <bool WIN>
__device__ __forceinline__ void gatherer_main(const Params& P, double* scratch, Ctl* ctl, int gi) {
    const TilePlan plan(P);
    const int lane = threadIdx.x & 31;
    if (P.n_cherry == 0 || K2_DBG_NOSYNC) return;
    double* cta_scratch = scratch + (size_t)blockIdx.x * cta_rows(P) * P.Vp;
    const int n_pieces = (P.W + 1) / 2;  // 16-byte pieces of a vector that hold a size < W
    int units_done = 0;
    for (int pair = 0; pair < plan.n_pairs; ++pair) {
        // slot set (pair & 1) was last read by pair - 2: wait until pair - 1 has completed an op (then pair - 2 is over)
        if (pair >= 2) {
            const int need = (pair - 1) * P.n_ops + 1;
            while (!K2_DBG(2) && ctl->done[0] < need) { __nanosleep(500); }
            __threadfence_block();
        }
        for (int oi = 0; oi < P.n_ops; ++oi) {
            const Op op = P.ops[oi];
            if (op.a_kind != 1) continue;
            const double* __restrict__ MTa = P.MT + (size_t)op.key_a1 * P.Sp * P.Sp;
            const double* __restrict__ MTb = P.MT + (size_t)op.key_a2 * P.Sp * P.Sp;
            for (int h = 0; h < 2; ++h) {
                int mb0, m;
                if (!plan.tile(2 * pair + h, mb0, m)) continue;
                double* out = cta_scratch + (size_t)cherry_row(P, pair, h, op.in_slot) * P.Vp;
                for (int r0 = (TILE_M / 2) * gi; r0 < (TILE_M / 2) * (gi + 1); r0 += 2) {
                    const double2* pa[2]; const double2* pb[2]; double2* po[2]; bool ok[2]; int wlim[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int f = TilePlan::family_or_neg(mb0, m, r0 + u, P.F);
                        ok[u] = f >= 0;
                        const int fc = ok[u] ? f : 0;
                        const int ca = __ldg(P.counts + (size_t)op.leaf_a1 * P.leaf_stride + fc), cb = __ldg(P.counts + (size_t)op.leaf_a2 * P.leaf_stride + fc);
                        // sizes >= W stay exact zeros (the vector has length W although the matrices are wider when S > W); with a
                        // per-family window also the sizes above it, and the whole vector when a leaf lies outside the window
                        wlim[u] = P.W;
                        if (WIN) {
                            const int cm = __ldg(P.colmax + fc);
                            wlim[u] = (ca <= cm && cb <= cm) ? min(P.W, cm + 1) : 0;
                        }
                        pa[u] = reinterpret_cast<const double2*>(MTa + (size_t)ca * P.Sp);
                        pb[u] = reinterpret_cast<const double2*>(MTb + (size_t)cb * P.Sp);
                        po[u] = reinterpret_cast<double2*>(out + (size_t)(r0 + u) * P.Vp);
                    }
                    for (int p0 = 0; p0 < n_pieces; p0 += 128) {  // 4 pieces per lane and row in flight
                        double2 x[2][4], y[2][4];
#pragma unroll
                        for (int u = 0; u < 2; ++u)
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int pc = p0 + k * 32 + lane;
                                x[u][k] = make_double2(0.0, 0.0); y[u][k] = x[u][k];
                                if (ok[u] && pc < n_pieces) { x[u][k] = __ldg(pa[u] + pc); y[u][k] = __ldg(pb[u] + pc); }
                            }
#pragma unroll
                        for (int u = 0; u < 2; ++u)
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int pc = p0 + k * 32 + lane;
                                if (ok[u] && pc < n_pieces) {
                                    const double lo = (2 * pc < wlim[u]) ? __dmul_rn(x[u][k].x, y[u][k].x) : 0.0;
                                    const double hi = (2 * pc + 1 < wlim[u]) ? __dmul_rn(x[u][k].y, y[u][k].y) : 0.0;
                                    po[u][pc] = make_double2(lo, hi);
                                }
                            }
                    }
                }
                __syncwarp();
                ++units_done;
                if (lane == 0) {
                    __threadfence();  // the vector is read by the producer's TMA loads
                    ctl->cherry_count[gi] = units_done;
                }
            }
        }
    }
}

// ================================ epilogue manager (1 warp) ================================ ================================
template <bool PROF, bool WIN>
__device__ __forceinline__ void cmanager_main(const CUtensorMap* tmA, const Params& P, unsigned char* Cbuf, Ctl* ctl) {
    const TilePlan plan(P);
    const int lane = threadIdx.x & 31;
    const int scratch_row0 = blockIdx.x * cta_rows(P);
    const uint32_t sC = smem_u32(Cbuf);
    if (K2_DBG(2)) return;
    uint32_t item = 0;
    int ops_done_base = 0;
    const bool prof = PROF && P.warp_prof != nullptr && blockIdx.x == 0;
    long long t_prep = 0, t_wait_cdone = 0, t_store = 0;
    const long long t_begin = prof ? clock64() : 0;
    for (int pair = 0; pair < plan.n_pairs; ++pair) {
        int wmax_h[2] = {P.W, P.W}, rlo_h[2] = {0, 0}, rhi_h[2] = {0, 0};
        if (WIN) {
            // windowed mode: the windows / root picks of the rows of this pair's tiles (every consumer has left the previous pair:
            // its last pass was handed back through c_done before this point)
            __syncwarp();
            for (int h = 0; h < 2; ++h) {
                int mb0, m;
                if (!plan.tile(2 * pair + h, mb0, m)) continue;
                for (int r = lane; r < TILE_M; r += 32) {
                    const int f = TilePlan::family(mb0, m, r, P.F);
                    ctl->colmax[h][r] = __ldg(P.colmax + f);
                    ctl->pick[h][r] = P.root_pick ? __ldg(P.root_pick + f) - P.root_min : -1;
                }
                const TileWin tw = tile_win_warp(P, mb0, m, lane);
                wmax_h[h] = tw.wmax; rlo_h[h] = tw.rch_lo; rhi_h[h] = tw.rch_hi;
            }
            __syncwarp();
        }
        for (int oi = 0; oi < P.n_ops; ++oi) {
            const Op op = P.ops[oi];
            const int r0 = op.is_root ? P.root_min : 0;
            const bool reduce_now = op.is_root && op.other_kind != 0;
            const double* __restrict__ MTo = P.MT + (size_t)op.key_o * P.Sp * P.Sp;
            const int nrows_full = op.is_root ? P.R : P.W;
            const int n_chunks_full = (nrows_full + TN - 1) / TN;
            for (int h = 0; h < 2; ++h) {
                int mb0, m;
                if (!plan.tile(2 * pair + h, mb0, m)) continue;
                const int nrows = (WIN && !op.is_root) ? wmax_h[h] : nrows_full;  // windowed: sizes from the tile's largest window on are not computed
                const int n_chunks = WIN ? (op.is_root ? rhi_h[h] : (nrows + TN - 1) / TN) : n_chunks_full;  // one past the last pass
                const int ch_lo = (WIN && op.is_root) ? rlo_h[h] : 0;
                if (op.other_kind == 1) {
                    __syncwarp();
                    for (int r = lane; r < TILE_M; r += 32) {
                        const int f = TilePlan::family(mb0, m, r, P.F);
                        const int cnt = __ldg(P.counts + (size_t)op.leaf_o * P.leaf_stride + f);
                        // a leaf outside the family's window has factor 0 (the one-hot entry is not part of the vector)
                        ctl->rowoff_o[r] = (WIN && cnt > __ldg(P.colmax + f)) ? -1 : cnt * P.Sp;
                    }
                    __syncwarp();
                }
                const int out_row = scratch_row0 + (h * P.n_slots + op.out_slot) * TILE_M;
                for (int ch = ch_lo; ch < n_chunks; ++ch) {
                    const int nbx = min(C_BOXES, (P.Vp - ch * TN) / BK);  // boxes of this pass inside the vector
                    // sizes of the pass the consumers own: whole 8-size blocks up to nrows (consumer_main).  What lies beyond, up to
                    // the end of the stored boxes, must be zeros in the slot - staged here, never touched by the consumers.
                    const int ncons = min(TN, ((nrows - ch * TN + 7) >> 3) << 3);
                    const long long tc0 = prof ? clock64() : 0;
                    // ---- stage the sibling factor of this pass in the C tile ----
                    if (op.other_kind == 2) {
                        if (lane == 0) {
                            mbar_arrive_expect_tx(&ctl->c_ready, nbx * C_BOX_BYTES);
                            for (int b = 0; b < nbx; ++b)
                                tma_load_2d(Cbuf + b * C_BOX_BYTES, tmA, ch * TN + b * BK, out_row, &ctl->c_ready);
                        } else {
                            mbar_arrive(&ctl->c_ready);
                        }
                    } else if (op.other_kind == 1) {
                        // leaf sibling (cafe_tree.c:204-210): row `count` of the transposed matrix, sizes r0 + pass
                        const int gc0 = r0 + ch * TN;
                        if ((r0 & 1) == 0) {
                            const int cj = lane & 7, rs = lane >> 3;
                            for (int b = 0; b < nbx; ++b) {
                                const int col = gc0 + b * BK + 2 * cj;
                                const bool in = col < P.Sp && b * BK + 2 * cj < ncons;
                                const double* src0 = MTo + (in ? col : 0);
#pragma unroll 4
                                for (int i = 0; i < TILE_M / 4; ++i) {
                                    const int r = rs + 4 * i;
                                    const int ro = ctl->rowoff_o[r];
                                    cp_async16(sC + b * C_BOX_BYTES + r * 128 + ((cj ^ (r & 7)) << 4), src0 + max(ro, 0), (in && ro >= 0) ? 16 : 0);
                                }
                            }
                        } else {  // odd first size (the root range starts at 1): rows are only 8-byte aligned
                            const int cl = lane & 15, rs = lane >> 4;
                            for (int b = 0; b < nbx; ++b) {
                                const int col = gc0 + b * BK + cl;
                                const bool in = col < P.Sp && b * BK + cl < ncons;
                                const double* src0 = MTo + (in ? col : 0);
#pragma unroll 4
                                for (int i = 0; i < TILE_M / 2; ++i) {
                                    const int r = rs + 2 * i;
                                    const int ro = ctl->rowoff_o[r];
                                    cp_async8(sC + b * C_BOX_BYTES + r * 128 + (((cl >> 1) ^ (r & 7)) << 4) + ((cl & 1) << 3),
                                              src0 + max(ro, 0), (in && ro >= 0) ? 8 : 0);
                                }
                            }
                        }
                        cp_async_arrive_noinc(&ctl->c_ready);
                    } else {
                        if (ncons < nbx * BK) {  // no sibling factor to stage: only the zeros beyond the consumers' blocks
                            const int nz = nbx * BK - ncons;
                            for (int idx = lane; idx < TILE_M * nz; idx += 32) {
                                const int r = idx / nz, c = ncons + idx % nz;
                                *reinterpret_cast<double*>(Cbuf + (c >> 4) * C_BOX_BYTES + r * 128 + ((((c & 15) >> 1) ^ (r & 7)) << 4) + ((c & 1) << 3)) = 0.0;
                            }
                        }
                        mbar_arrive(&ctl->c_ready);
                    }
                    // ---- the consumers multiply in place ----
                    const long long tc1 = prof ? clock64() : 0;
                    mbar_wait_sleepy(&ctl->c_done, item & 1);
                    const long long tc2 = prof ? clock64() : 0;
                    if (!reduce_now) {
                        // ... then the tile goes back to the scratch slot
                        if (lane == 0 && !K2_DBG(4)) {
                            fence_proxy_async_smem();  // consumer writes (generic proxy, acquired above) -> TMA store (async proxy)
                            for (int b = 0; b < nbx; ++b) tma_store_2d(tmA, ch * TN + b * BK, out_row, Cbuf + b * C_BOX_BYTES);
                            bulk_commit();
                            bulk_wait_all();  // the C tile is free again and the slot is written
                        }
                    } else {
                        // ... or, at the root (reduced by the consumers), is only copied out on request
                        const int ncols = min(TN, nrows - ch * TN);
                        if (WIN && P.root_pick) {  // the distribution's root range {s}: one likelihood per simulated family
                            for (int r = lane; r < TILE_M; r += 32) {
                                const int f = TilePlan::family_or_neg(mb0, m, r, P.F);
                                const int c = ctl->pick[h][r] - ch * TN;
                                if (f >= 0 && c >= 0 && c < ncols)
                                    P.L0_out[f] = *reinterpret_cast<const double*>(Cbuf + (c >> 4) * C_BOX_BYTES + r * 128 + ((((c & 15) >> 1) ^ (r & 7)) << 4) + ((c & 1) << 3));
                            }
                        }
                        if (P.Lroot_out) {  // get_likelihoods (cafe_tree.c:325-329): rows copied out, lanes along the sizes
                            for (int r = 0; r < TILE_M; ++r) {
                                const int f = TilePlan::family_or_neg(mb0, m, r, P.F);
                                if (f < 0) continue;
                                for (int c = lane; c < ncols; c += 32)
                                    P.Lroot_out[(size_t)f * P.R + ch * TN + c] =
                                        *reinterpret_cast<const double*>(Cbuf + (c >> 4) * C_BOX_BYTES + r * 128 + ((((c & 15) >> 1) ^ (r & 7)) << 4) + ((c & 1) << 3));
                            }
                        }
                    }
                    __syncwarp();
                    if (ch == n_chunks - 1 && lane == 0) {
                        __threadfence();
                        ctl->done[h] = ops_done_base + oi + 1;
                    }
                    ++item;
                    if (prof) { const long long tc3 = clock64(); t_prep += tc1 - tc0; t_wait_cdone += tc2 - tc1; t_store += tc3 - tc2; }
                }
            }
        }
        ops_done_base += P.n_ops;
    }
    if (prof && lane == 0) {
        long long* o = P.warp_prof + (N_CONSUMER_WARPS + 3) * 8;
        o[0] = clock64() - t_begin; o[1] = t_prep; o[2] = t_wait_cdone; o[3] = t_store; o[4] = item;
    }
}

// ================================ warps 4..11: DMMA consumers ================================
// Shared-memory accesses of the K loops go through 32-bit shared addresses with compile-time offsets (LDS [R + imm]): a lane keeps
// eight address registers per ring stage (A and B, one per k4-step of a K block: the swizzle term differs per step) and nothing
// else is recomputed at a stage boundary.  Written with generic pointers the compiler rematerialised the whole address chain
// (thread id, shared window base, alignment, swizzle) at every stage to stay inside the register budget: ~70 instructions with
// special-register and constant-bank latencies between the last DMMA of a stage and the first of the next, for both DMMA warps
// of a sub-partition at once (profiles/r2_k2_experiments.md).
template <int IMM>
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(IMM));
    return v;
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// non-blocking probe of an mbarrier phase: issued a few k4-steps before the result is needed
__device__ __forceinline__ uint32_t mbar_test_u32(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// one arrival per warp, by an elected lane (no thread-id read, no predicate register to keep alive across the K loop)
__device__ __forceinline__ void mbar_arrive_elect_u32(uint32_t bar) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n"
        "}\n" ::"r"(bar) : "memory");
}
// a value the compiler must keep in a register instead of recomputing it from special registers at every use
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

// Fragments of one k4-step: NBV B fragments and one A fragment per 8-family block, OFF = byte offset of the K block in the stage.
// NBV: 8-size blocks of this warp in the pass (4, or 3 in a pass whose blocks do not divide by 4, see consumer_main)
template <int MBV, int NBV, int OFF>
__device__ __forceinline__ void load_frags(double (&fa)[MB], double (&fb)[NB], uint32_t pa, uint32_t pb) {
#define CAFE_LDB(nb_) if (nb_ < NBV) fb[nb_] = lds_f64<OFF + nb_ * 1024>(pb);
    CAFE_LDB(0) CAFE_LDB(1) CAFE_LDB(2) CAFE_LDB(3)
#undef CAFE_LDB
#define CAFE_LDA(mb_) if (mb_ < MBV) fa[mb_] = lds_f64<OFF + mb_ * 1024>(pa);
    CAFE_LDA(0) CAFE_LDA(1) CAFE_LDA(2) CAFE_LDA(3) CAFE_LDA(4) CAFE_LDA(5)
#undef CAFE_LDA
}
template <int MBV, int NBV>
__device__ __forceinline__ void mma_frags(double (&acc)[MB][NB][2], const double (&fa)[MB], const double (&fb)[NB]) {
#pragma unroll
    for (int mb = 0; mb < MBV; ++mb)
#pragma unroll
        for (int nb = 0; nb < NBV; ++nb) dmma_884(acc[mb][nb][0], acc[mb][nb][1], fa[mb], fb[nb]);
}

// K loop of one pass: n_kblocks K blocks of 16 sizes (the last one with tail_steps k4-steps), KB_PER_STAGE per ring stage.
// MBV == 0: this warp has no work in the tile, it only keeps the ring moving.
// blk0: first 8-size block of this warp inside the pass's 128 sizes.
// The fragments of the next step are fetched before the DMMAs of the current one; during the last step of a stage the first
// fragments of the next stage (its barrier is probed two steps earlier: the ~100-cycle latency of the probe then never sits
// between two DMMAs), so that a stage boundary costs the arrive on the empty barrier and eight address updates.
template <int MBV, int NBV>
__device__ __forceinline__ void gemm_kblocks(double (&acc)[MB][NB][2], uint32_t ring, uint32_t bars, uint32_t& stage,
                                             uint32_t& phase, int n_kblocks, int tail_steps, int grp, int blk0, int pg, int sg, int q,
                                             bool prof, long long& t_wait_full) {
    static_assert(KB_PER_STAGE == 2 && NSTAGE == 2, "the stage loop below is written for two stages of two K blocks");
    // bars: shared address of Ctl::full[0]; full[s] at +8 s, empty[s] at +16 + 8 s
    const int n_stages = (n_kblocks + KB_PER_STAGE - 1) / KB_PER_STAGE;
    if (MBV == 0) {
        for (int st = 0; st < n_stages; ++st) {
            if (!K2_DBG_NOSYNC) mbar_wait_u32(bars + 8 * stage, phase);
            __syncwarp();
            if (!K2_DBG_NOSYNC) mbar_arrive_elect_u32(bars + 16 + 8 * stage);
            stage ^= 1; phase ^= (stage == 0);
        }
        return;
    }
    const int hi = q >> 1;
    const uint32_t lane_off = ring + stage * STAGE_BYTES + ((q & 1) << 3);
    uint32_t pa[4], pb[4];  // this lane's A / B fragment addresses in the current stage, per k4-step of a K block
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // tile row pg (A: families, mma_row_perm) / sg (B: sizes, sigma) of every 8-row block
        pa[kk] = lane_off + pg * 128 + (((2 * kk + hi) ^ pg) << 4) + grp * (HM * 128);
        pb[kk] = lane_off + sg * 128 + (((2 * kk + hi) ^ sg) << 4) + A_BYTES + blk0 * 8 * 128;
    }
    int delta = stage ? -STAGE_BYTES : STAGE_BYTES;   // to the other stage
    uint32_t fbar_next = bars + 8 * (stage ^ 1), ebar = bars + 16 + 8 * stage;  // Ctl is 1024-byte aligned: ^ 8 switches the stage
    const int n_full = (tail_steps == 4) ? n_kblocks : n_kblocks - 1;  // K blocks with all four steps

    double fa[2][MB], fb[2][NB];
    if (!K2_DBG_NOSYNC) {
        const long long t0 = prof ? clock64() : 0;
        mbar_wait_u32(bars + 8 * stage, phase);
        if (prof) t_wait_full += clock64() - t0;
    }
    load_frags<MBV, NBV, 0>(fa[0], fb[0], pa[0], pb[0]);
    for (int st = 0; st < n_stages; ++st) {
        const int kb0 = st * KB_PER_STAGE;
        if (kb0 + 2 <= n_full) {
            // two full K blocks; then the first fragments of the next stage (if any)
            const bool has_next = st + 1 < n_stages;
            load_frags<MBV, NBV, 0>(fa[1], fb[1], pa[1], pb[1]);          mma_frags<MBV, NBV>(acc, fa[0], fb[0]);
            load_frags<MBV, NBV, 0>(fa[0], fb[0], pa[2], pb[2]);          mma_frags<MBV, NBV>(acc, fa[1], fb[1]);
            load_frags<MBV, NBV, 0>(fa[1], fb[1], pa[3], pb[3]);          mma_frags<MBV, NBV>(acc, fa[0], fb[0]);
            load_frags<MBV, NBV, SUB_BYTES>(fa[0], fb[0], pa[0], pb[0]);  mma_frags<MBV, NBV>(acc, fa[1], fb[1]);
            load_frags<MBV, NBV, SUB_BYTES>(fa[1], fb[1], pa[1], pb[1]);  mma_frags<MBV, NBV>(acc, fa[0], fb[0]);
            uint32_t ready = 1;
            if (has_next && !K2_DBG_NOSYNC) ready = mbar_test_u32(fbar_next, phase ^ stage);
            load_frags<MBV, NBV, SUB_BYTES>(fa[0], fb[0], pa[2], pb[2]);  mma_frags<MBV, NBV>(acc, fa[1], fb[1]);
            load_frags<MBV, NBV, SUB_BYTES>(fa[1], fb[1], pa[3], pb[3]);  mma_frags<MBV, NBV>(acc, fa[0], fb[0]);
            if (has_next) {
                if (!ready) {
                    const long long t0 = prof ? clock64() : 0;
                    mbar_wait_u32(fbar_next, phase ^ stage);
                    if (prof) t_wait_full += clock64() - t0;
                }
                load_frags<MBV, NBV, 0>(fa[0], fb[0], pa[0] + delta, pb[0] + delta);
            }
            mma_frags<MBV, NBV>(acc, fa[1], fb[1]);
        } else {
            // the last stage of the pass: [full block] [partial block], either may be missing
            int kb = kb0;
            if (kb < n_full) {
                const bool more = kb + 1 < n_kblocks;  // a partial block follows in this stage
                load_frags<MBV, NBV, 0>(fa[1], fb[1], pa[1], pb[1]);  mma_frags<MBV, NBV>(acc, fa[0], fb[0]);
                load_frags<MBV, NBV, 0>(fa[0], fb[0], pa[2], pb[2]);  mma_frags<MBV, NBV>(acc, fa[1], fb[1]);
                load_frags<MBV, NBV, 0>(fa[1], fb[1], pa[3], pb[3]);  mma_frags<MBV, NBV>(acc, fa[0], fb[0]);
                if (more) load_frags<MBV, NBV, SUB_BYTES>(fa[0], fb[0], pa[0], pb[0]);
                mma_frags<MBV, NBV>(acc, fa[1], fb[1]);
                ++kb;
            }
            if (kb < n_kblocks && kb >= n_full) {
                // the partial block sits at offset 0 of the stage when it is alone, after the full block otherwise
                if (kb == kb0) {
                    mma_frags<MBV, NBV>(acc, fa[0], fb[0]);
                    if (tail_steps > 1) { load_frags<MBV, NBV, 0>(fa[0], fb[0], pa[1], pb[1]); mma_frags<MBV, NBV>(acc, fa[0], fb[0]); }
                    if (tail_steps > 2) { load_frags<MBV, NBV, 0>(fa[0], fb[0], pa[2], pb[2]); mma_frags<MBV, NBV>(acc, fa[0], fb[0]); }
                } else {
                    mma_frags<MBV, NBV>(acc, fa[0], fb[0]);
                    if (tail_steps > 1) { load_frags<MBV, NBV, SUB_BYTES>(fa[0], fb[0], pa[1], pb[1]); mma_frags<MBV, NBV>(acc, fa[0], fb[0]); }
                    if (tail_steps > 2) { load_frags<MBV, NBV, SUB_BYTES>(fa[0], fb[0], pa[2], pb[2]); mma_frags<MBV, NBV>(acc, fa[0], fb[0]); }
                }
            }
        }
        // every lane's fragment loads of this stage have completed: the warp-wide DMMAs that consume them have been issued
        if (!K2_DBG_NOSYNC) mbar_arrive_elect_u32(ebar);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) { pa[kk] += delta; pb[kk] += delta; }
        delta = -delta; fbar_next ^= 8; ebar ^= 8;
        phase ^= stage; stage ^= 1;
    }
}

template <bool PROF, bool WIN>
__device__ __forceinline__ void consumer_main(const Params& P, unsigned char* stage_base, unsigned char* Cbuf, Ctl* ctl) {
    const TilePlan plan(P);
    const int warp = (threadIdx.x >> 5) - N_AUX_WARPS, lane = threadIdx.x & 31;  // consumer warp 0..7
    const int grp = warp >> 2, nw = warp & 3;
    const int g = lane >> 2, q = lane & 3;
    const int pg = mma_row_perm(g);
    // Rows of the matrix tile (B operand = output sizes) use a second permutation, sigma = {0,2,5,7,6,4,3,1}: like mma_row_perm it
    // puts the four rows of a half-warp into four different 32-byte bank pairs (conflict-free fragment loads), and in addition the
    // sizes sigma(2q), q = 0..3, are distinct mod 4 (and so are sigma(2q+1)): with the 128B swizzle the lanes' first (second)
    // accumulator elements then cover all 32 banks in the C tile by themselves.  mma_row_perm alone needed a lane-dependent access
    // order there, i.e. two selects per element in the epilogue.
    const int sg = (0x13467520 >> (4 * g)) & 7;
    const int pcA = (0x13467520 >> (8 * q)) & 7, pcB = (0x13467520 >> (8 * q + 4)) & 7;  // sizes of this lane's accumulator pair in an 8-block
    const int n_kblocks_full = (P.W + BK - 1) / BK;
    const int tail_steps_full = ((P.W - (n_kblocks_full - 1) * BK) + 3) >> 2;  // k4-steps of the last K block (1..4)

    // Byte offsets of this lane's two accumulator columns inside a C box, for even / odd 8-size blocks (128B swizzle).
    int coff[2][2];
#pragma unroll
    for (int par = 0; par < 2; ++par) {
        coff[0][par] = pg * 128 + ((pcA & 1) << 3) + ((((par << 2) | (pcA >> 1)) ^ pg) << 4);
        coff[1][par] = pg * 128 + ((pcB & 1) << 3) + ((((par << 2) | (pcB >> 1)) ^ pg) << 4);
    }
    unsigned char* cgrp = Cbuf + (grp * HM) * 128;  // a lane's element of block b (of the pass) and 8-family block mb: box b >> 1, parity b & 1

    // debug profile (CTA 0): cycles waiting for ring stages / in K loops / waiting for the C tile / in epilogues
    const bool prof = PROF && P.warp_prof != nullptr && blockIdx.x == 0;
    long long t_wait_full = 0, t_kloop = 0, t_wait_c = 0, t_epi = 0, t_epi_root = 0, t_kloop_cherry = 0;
    const long long t_begin = prof ? clock64() : 0;

    uint32_t stage = 0, phase = 0, item = 0;
    const uint32_t ring_u32 = opaque_u32(smem_u32(stage_base)), bars_u32 = opaque_u32(smem_u32(&ctl->full[0]));
    for (int pair = 0; pair < plan.n_pairs; ++pair) {
        // the two tiles of the pair: this group's 8-family blocks (everything else about a tile concerns the helper warps)
        int mbv_h[2], f0_h[2], wmax_h[2] = {0, 0}, rlo_h[2] = {0, 0}, rhi_h[2] = {0, 0};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int mb0 = 0, m = 0;
            const bool have = plan.tile(2 * pair + h, mb0, m);
            mbv_h[h] = have ? TilePlan::mbv(m, grp) : -1;
            f0_h[h] = (mb0 + TilePlan::pre(m, grp)) * 8;
            if (WIN && have) { const TileWin tw = tile_win_warp(P, mb0, m, lane); wmax_h[h] = tw.wmax; rlo_h[h] = tw.rch_lo; rhi_h[h] = tw.rch_hi; }
        }
        for (int oi = 0; oi < P.n_ops; ++oi) {
            const int flags = ctl->opflags[oi];
            const bool is_root = flags & 1;
            const int other_kind = (flags >> 2) & 3;
            const bool reduce_now = is_root && other_kind != 0;
            const int nrows_full = is_root ? P.R : P.W;
            const int n_chunks_full = (nrows_full + TN - 1) / TN;
            for (int h = 0; h < 2; ++h) {
                const int mbv_t = h ? mbv_h[1] : mbv_h[0], f0_t = h ? f0_h[1] : f0_h[0];
                if (mbv_t < 0) continue;
                // windowed mode: the tile's own K extent and, below the root, output sizes (tile_wmax_warp)
                const int wmax_t = h ? wmax_h[1] : wmax_h[0];
                const int nrows = (WIN && !is_root) ? wmax_t : nrows_full;
                const int n_chunks = WIN ? (is_root ? (h ? rhi_h[1] : rhi_h[0]) : (nrows + TN - 1) / TN) : n_chunks_full;  // one past the last pass
                const int ch_lo = (WIN && is_root) ? (h ? rlo_h[1] : rlo_h[0]) : 0;
                const int n_kblocks = WIN ? (wmax_t + BK - 1) / BK : n_kblocks_full;
                const int tail_steps = WIN ? ((wmax_t - (n_kblocks - 1) * BK) + 3) >> 2 : tail_steps_full;
                // running root reduction of one family row of this group, owned by the group's first HM threads
                double run_ml = -1.0, run_mp = -INFINITY; int run_am = 0x7fffffff;
                for (int ch = ch_lo; ch < n_chunks; ++ch) {
                    // The 8-size blocks of the pass that hold a size < nrows go to the four N-warps of the group: four each in a
                    // full pass.  A pass with fewer blocks (the last one: 13 at W = 481) is split evenly instead, and group 1
                    // takes them in an order rotated by two warps - warp nw of both groups runs on SM sub-partition nw, so the
                    // larger shares of the two groups land on different sub-partitions (4+3, 3+3, 3+4, 3+3 blocks instead of 4+4 x 4).
                    const int nbl = min(TN / 8, (nrows - ch * TN + 7) >> 3);
                    int blk0 = 4 * nw, cnt = 4;
                    if (nbl < TN / 8) {
                        const int wr = (grp == 0) ? nw : ((nw + 2) & 3);
                        const int base = nbl >> 2, rem = nbl & 3;
                        cnt = base + (wr < rem ? 1 : 0);
                        blk0 = wr * base + min(wr, rem);
                    }
                    const int n0 = ch * TN + blk0 * 8;  // first output size of this warp
                    const int mbw = (cnt > 0) ? mbv_t : 0;
                    double acc[MB][NB][2];
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;

                    const long long tk0 = prof ? clock64() : 0;
#define CAFE_K(MBV_, NBV_) gemm_kblocks<MBV_, NBV_>(acc, ring_u32, bars_u32, stage, phase, n_kblocks, tail_steps, grp, blk0, pg, sg, q, prof, t_wait_full);
                    if (cnt <= 3) {
                        switch (mbw) {
                            case 6: CAFE_K(6, 3) break;
                            case 5: CAFE_K(5, 3) break;
                            case 4: CAFE_K(4, 3) break;
                            case 3: CAFE_K(3, 3) break;
                            case 2: CAFE_K(2, 3) break;
                            case 1: CAFE_K(1, 3) break;
                            default: CAFE_K(0, 4) break;
                        }
                    } else {
                        switch (mbw) {
                            case 6: CAFE_K(6, 4) break;
                            case 5: CAFE_K(5, 4) break;
                            case 4: CAFE_K(4, 4) break;
                            case 3: CAFE_K(3, 4) break;
                            case 2: CAFE_K(2, 4) break;
                            case 1: CAFE_K(1, 4) break;
                            default: CAFE_K(0, 4) break;
                        }
                    }
#undef CAFE_K
                    const long long tk1 = prof ? clock64() : 0;

                    // ---------------- epilogue of this pass: C = acc * C in shared memory ----------------
                    if (!K2_DBG(2)) mbar_wait(&ctl->c_ready, item & 1);
                    const long long tk2 = prof ? clock64() : 0;
                    if (K2_DBG(3)) {  // keep the accumulators alive
                        double sum = 0.0;
#pragma unroll
                        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb) sum += acc[mb][nb][0] + acc[mb][nb][1];
                        if (sum == 1.2345e-300) P.logpost[0] = sum;
                    }
                    // Three instantiations.  FAST (the common case): the warp owns the four blocks 4 nw .. 4 nw + 3 of the pass and
                    // all of its 32 sizes lie below the limit - block indices are compile-time constants, no guards, no selects;
                    // with or without a sibling factor (FAC).  Otherwise the general form: runtime block indices, per-element limit.
                    auto multiply_in_place = [&](auto fast_c, auto fac_c) {
                        constexpr bool FAST = decltype(fast_c)::value, FAC = decltype(fac_c)::value;
#pragma unroll
                        for (int mb = 0; mb < MB; ++mb) {
                            if (mb < mbw) {
                                // all eight factors of this 8-family block first, then the products, then the stores
                                // (shared-memory pointers may alias for the compiler: written out explicitly)
                                double fac[NB][2];
                                unsigned char* cel[NB][2];  // this lane's two elements of every block
#pragma unroll
                                for (int nb = 0; nb < NB; ++nb) {
                                    const int boxi = FAST ? (2 * nw + (nb >> 1)) : ((blk0 + nb) >> 1);
                                    const int par = FAST ? (nb & 1) : ((blk0 + nb) & 1);
                                    unsigned char* box = cgrp + boxi * C_BOX_BYTES + mb * 1024;
                                    cel[nb][0] = box + (par ? coff[0][1] : coff[0][0]);
                                    cel[nb][1] = box + (par ? coff[1][1] : coff[1][0]);
                                    fac[nb][0] = 1.0; fac[nb][1] = 1.0;
                                    if (FAST ? FAC : (other_kind != 0 && nb < cnt)) {
                                        fac[nb][0] = *reinterpret_cast<const volatile double*>(cel[nb][0]);
                                        fac[nb][1] = *reinterpret_cast<const volatile double*>(cel[nb][1]);
                                    }
                                }
                                double out[NB][2];
                                // sizes >= nrows are exact zeros already (the matrix tile's rows end at nrows: zero fill, see the tensor
                                // maps).  Windowed mode: below the root also the sizes above the family's own window must be zeros
                                // (the reference never computes them)
                                const int lim = (WIN && !is_root) ? min(nrows, ctl->colmax[h][grp * HM + mb * 8 + pg] + 1) : nrows;
#pragma unroll
                                for (int nb = 0; nb < NB; ++nb) {
                                    if (FAST) {  // a product with 1.0 is the value itself
                                        out[nb][0] = FAC ? __dmul_rn(acc[mb][nb][0], fac[nb][0]) : acc[mb][nb][0];
                                        out[nb][1] = FAC ? __dmul_rn(acc[mb][nb][1], fac[nb][1]) : acc[mb][nb][1];
                                    } else if (WIN) {
                                        out[nb][0] = (n0 + nb * 8 + pcA < lim) ? __dmul_rn(acc[mb][nb][0], fac[nb][0]) : 0.0;
                                        out[nb][1] = (n0 + nb * 8 + pcB < lim) ? __dmul_rn(acc[mb][nb][1], fac[nb][1]) : 0.0;
                                    } else {
                                        out[nb][0] = __dmul_rn(acc[mb][nb][0], fac[nb][0]);
                                        out[nb][1] = __dmul_rn(acc[mb][nb][1], fac[nb][1]);
                                    }
                                }
#pragma unroll
                                for (int nb = 0; nb < NB; ++nb) {
                                    if (FAST || nb < cnt) {
                                        *reinterpret_cast<volatile double*>(cel[nb][0]) = out[nb][0];
                                        *reinterpret_cast<volatile double*>(cel[nb][1]) = out[nb][1];
                                    }
                                }
                            }
                        }
                    };
                    if (!K2_DBG(3)) {
                        #ifdef K2_NO_FAST
                        const bool fast = false;
#else
                        const bool fast = cnt == 4 && blk0 == 4 * nw && !(WIN && !is_root);
#endif
                        if (fast && other_kind != 0) multiply_in_place(std::true_type{}, std::true_type{});
                        else if (fast) multiply_in_place(std::true_type{}, std::false_type{});
                        else multiply_in_place(std::false_type{}, std::false_type{});
                    }
                    // no proxy fence here (MEMBAR.ALL.CTA would drain every store of the warp with the DMMA pipe idle): the arrive
                    // below releases the writes, the epilogue manager acquires them and fences before its TMA store
                    if (reduce_now && P.logpost && !K2_DBG(3)) {
                        // root: L[i] = acc * other; max / first argmax of L and max of log L + log prior (lambda.cpp:670-686).
                        // log is monotonic, so among this lane's eight sizes of a family only the one with the largest product
                        // L * prior can carry the maximum: the products are compared exactly as (exponent sum, mantissa product)
                        // - no underflow - and ONE log is taken per lane and family instead of eight (two products closer than an
                        // ulp could swap; their log sums then differ by ~1e-16).  Comparisons on bit patterns: non-negative doubles
                        // order like integers, and DSETP would queue behind the DMMAs of the other group on the fp64 pipe.
                        // L is read back from the C tile (each lane reads what it just wrote): the accumulators are dead by now, which
                        // keeps this rarely executed block out of the register budget of the K loops.
                        double pr_m[NB][2], pr_lp[NB][2]; int pr_e[NB][2];  // priors of this lane's 8 sizes, the same for every family
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const int i = min(n0 + nb * 8 + (hh ? pcB : pcA), nrows - 1);
                                pr_m[nb][hh] = __ldg(P.prior_mant + i); pr_e[nb][hh] = __ldg(P.prior_exp + i); pr_lp[nb][hh] = __ldg(P.logprior + i);
                            }
                        }
#pragma unroll
                        for (int mb = 0; mb < MB; ++mb) {
                            long long ml = -1, best_f = 0; double mp = -INFINITY, best_v = 0.0, best_lp = 0.0; int am = 0x7fffffff, best_e = -0x7fffffff;
                            if (mb < mbw) {
#pragma unroll
                                for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
                                    for (int hh = 0; hh < 2; ++hh) {
                                        const int i = n0 + nb * 8 + (hh ? pcB : pcA);
                                        if (i < nrows && nb < cnt) {
                                            const int b = blk0 + nb;
                                            const double v = *reinterpret_cast<const volatile double*>(cgrp + (b >> 1) * C_BOX_BYTES + mb * 1024 + ((b & 1) ? coff[hh][1] : coff[hh][0]));
                                            const long long vb = __double_as_longlong(v);
                                            if (vb > ml || (vb == ml && i < am)) { ml = vb; am = i; }
                                            if (vb > 0) {
                                                double vs = v;
                                                int hi32 = __double2hiint(vs), e = (hi32 >> 20) & 0x7ff;
                                                if (e == 0) { vs = __dmul_rn(vs, 0x1p200); hi32 = __double2hiint(vs); e = ((hi32 >> 20) & 0x7ff) - 200; }
                                                // mantissa product in [1,4): its own exponent bit joins the exponent sum, its fraction breaks ties
                                                const long long pb = __double_as_longlong(__dmul_rn(__hiloint2double((hi32 & 0x000fffff) | 0x3ff00000, __double2loint(vs)), pr_m[nb][hh]));
                                                e += pr_e[nb][hh] + (int)(pb >> 52);
                                                const long long frac = pb & 0x000fffffffffffffLL;
                                                if (e > best_e || (e == best_e && frac > best_f)) { best_e = e; best_f = frac; best_v = v; best_lp = pr_lp[nb][hh]; }
                                            }
                                        }
                                    }
                                }
                                if (best_e != -0x7fffffff) mp = log(best_v) + best_lp;
                            }
                            // the 4 lanes of a quad hold the same family row
#pragma unroll
                            for (int off = 1; off <= 2; off <<= 1) {
                                const long long oml = __shfl_xor_sync(0xffffffffu, ml, off); const int oam = __shfl_xor_sync(0xffffffffu, am, off);
                                const double omp = __shfl_xor_sync(0xffffffffu, mp, off);
                                if (oml > ml || (oml == ml && oam < am)) { ml = oml; am = oam; }
                                if (omp > mp) mp = omp;
                            }
                            const int row = mb * 8 + pg;
                            if (q == 0) { ctl->red_ml[grp][nw][row] = __longlong_as_double(ml); ctl->red_mp[grp][nw][row] = mp; ctl->red_am[grp][nw][row] = am; }
                        }
                    }
                    const long long tk2b = prof ? clock64() : 0;
                    __syncwarp();
                    if (lane == 0 && !K2_DBG(2)) mbar_arrive(&ctl->c_done);
                    ++item;
                    if (prof && P.timeline && nw == 0 && lane == 0 && item <= 1024) {
                        long long* tl = P.timeline + ((size_t)grp * 1024 + (item - 1)) * 4;  // block 2 of the timeline: the producer, see producer_main
                        tl[0] = tk0; tl[1] = tk1; tl[2] = tk2; tl[3] = tk2b;
                    }
                    if (prof) { t_kloop += tk1 - tk0; t_wait_c += tk2 - tk1; t_epi += tk2b - tk2; if (flags & 2) t_kloop_cherry += tk1 - tk0; if (reduce_now) t_epi_root += tk2b - tk2; }
                    if (reduce_now && P.logpost && !K2_DBG(3)) {
                        group_bar(grp);
                        if (nw * 32 + lane < HM) {  // the first HM threads of the group own one family row each
                            const int row = nw * 32 + lane;
                            for (int w = 0; w < 4; ++w) {
                                const double oml = ctl->red_ml[grp][w][row], omp = ctl->red_mp[grp][w][row]; const int oam = ctl->red_am[grp][w][row];
                                if (oml > run_ml || (oml == run_ml && oam < run_am)) { run_ml = oml; run_am = oam; }
                                if (omp > run_mp) run_mp = omp;
                            }
                            const int f = f0_t + row;
                            if (ch == n_chunks - 1 && row < mbv_t * 8 && f < P.F) {
                                // max_j exp(log L + log prior) == exp(max_j(log L + log prior)); its log is the family's term
                                P.logpost[f] = log(exp(run_mp)); P.maxlik[f] = run_ml; P.argmax[f] = run_am;
                            }
                        }
                        group_bar(grp);
                    }
                }
            }
        }
    }
    if (prof && lane == 0) {
        long long* o = P.warp_prof + warp * 8;
        o[0] = clock64() - t_begin; o[1] = t_kloop; o[2] = t_wait_full; o[3] = t_wait_c; o[4] = t_epi; o[5] = t_epi_root; o[6] = 0; o[7] = t_kloop_cherry;
    }
}

// WIN: windowed mode (Params::colmax set): per-family column windows and root picks, for the conditional distribution and the
// p-values.  A separate instantiation, so that the score path carries none of it (measured: 0.8 % of a launch otherwise).
template <bool PROF, bool WIN>
__global__ void __launch_bounds__(THREADS, 1)
k_prune_fused2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmBroot, const Params P) {
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B tiles must start on a 1024-byte boundary of the shared window
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* stage_base = smem;                       // NSTAGE x KB_PER_STAGE x (A | B)
    unsigned char* Cbuf = smem + NSTAGE * STAGE_BYTES;      // 8 boxes of 96 x 16
    Ctl* ctl = reinterpret_cast<Ctl*>(Cbuf + C_BYTES);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], N_CONSUMER_WARPS); }
        mbar_init(&ctl->c_ready, 32);
        mbar_init(&ctl->c_done, N_CONSUMER_WARPS);
        ctl->done[0] = ctl->done[1] = 0; ctl->cherry_count[0] = ctl->cherry_count[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < P.n_ops; i += THREADS) {
        const Op o = P.ops[i];
        ctl->opflags[i] = (unsigned char)((o.is_root ? 1 : 0) | (o.a_kind << 1) | (o.other_kind << 2));
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5;
    if (warp < N_AUX_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_AUX));
        if (warp == 0) { if ((threadIdx.x & 31) == 0) producer_main<PROF, WIN>(&tmA, &tmB, &tmBroot, P, stage_base, ctl); }
        else if (warp == 3) cmanager_main<PROF, WIN>(&tmA, P, Cbuf, ctl);
        else gatherer_main<WIN>(P, P.scratch, ctl, warp - 1);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_CONSUMER));
        long long t_start = 0;
        if (P.cta_times && threadIdx.x == N_AUX_WARPS * 32) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
        consumer_main<PROF, WIN>(P, stage_base, Cbuf, ctl);
        if (P.cta_times && threadIdx.x == N_AUX_WARPS * 32) {
            long long t_end; unsigned smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const TilePlan plan(P);
            long long* o = P.cta_times + (size_t)blockIdx.x * 4;
            o[0] = smid; o[1] = t_start; o[2] = t_end; o[3] = plan.n_mb;
        }
    }
}

// Error-model leaves (cafe_tree.c:196-203): the leaf vector is row `observed` of the error matrix, so the leaf's factor is
// factor[i] = sum_j M[i][j] * E[observed][j] (birthdeath.c:172-180).  Built once per evaluation as ONE more transposed matrix
// per such leaf, MTE[observed][i] - same terms, same ascending-j order, one rounding per product and per sum, true sizes above
// colmax skipped as the matvec's column window does - the fused kernel then gathers a row of it exactly like a row of MT.
__global__ void __launch_bounds__(256)
k_err_leaf_matrix(const double* __restrict__ MTsrc, const int* __restrict__ rowptr, const int* __restrict__ col,
                  const double* __restrict__ val, int dim, int Sp, int colmax, double* __restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x, o = blockIdx.y;
    if (i >= Sp) return;
    double s = 0.0;
    if (o < dim) {
        for (int k = rowptr[o]; k < rowptr[o + 1]; ++k) {
            const int j = col[k];
            if (j <= colmax) s = __dadd_rn(s, __dmul_rn(MTsrc[(size_t)j * Sp + i], val[k]));
        }
    }
    out[(size_t)o * Sp + i] = s;
}

}  // namespace fused2

// =================================================================================================
// host side
// =================================================================================================
typedef CUresult (*PFN_encodeTiled2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled2 get_encode_fn2() {
    static PFN_encodeTiled2 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled2)p;
    }
    return fn;
}

struct Fused2State {
    fused2::Op* d_ops = nullptr; int ops_cap = 0;
    double* d_scratch = nullptr; size_t scratch_cap = 0;
    bool attr_set = false;
    std::vector<fused2::Op> ops;  // schedule of the current launch
    int n_slots = 0, n_cherry = 0;
};
static Fused2State& fstate2(cafe_gpu_ctx* ctx) {
    if (!ctx->fused2_state) ctx->fused2_state = new Fused2State();
    return *static_cast<Fused2State*>(ctx->fused2_state);
}
void fused2_release(cafe_gpu_ctx* ctx) {
    if (!ctx->fused2_state) return;
    Fused2State* s = static_cast<Fused2State*>(ctx->fused2_state);
    cudaFree(s->d_ops); cudaFree(s->d_scratch);
    delete s;
    ctx->fused2_state = nullptr;
}

static bool fused2_tree_supported(const cafe_gpu_ctx* ctx) {
    if (get_encode_fn2() == nullptr) return false;
    if (ctx->n_leaves < 3) return false;                 // the root of a two-leaf tree is itself a leaf pair
    if (ctx->n_nodes > fused2::OPFLAGS_CAP) return false;
    return true;
}
bool fused2_supported(const cafe_gpu_ctx* ctx) {
    if (!fused2_tree_supported(ctx)) return false;
    if (ctx->max_count >= ctx->W) return false;          // a one-hot leaf outside the matvec columns needs the guarded path
    return true;
}
// Windowed jobs (conditional distribution, p-values): leaves outside a family's window are handled by the window itself.  An
// error-model leaf's factor is sum_{j <= window} M[i][j] E[observed][j]: the per-leaf matrix of k_err_leaf_matrix sums over ALL
// columns, which is the same number as long as no non-zero entry of row `observed` lies above the window.  Every window reaches
// at least 50 sizes above the family's largest count (cafe_family.c:253, conditional_distribution.cpp:29: max + MAX(50, max/5),
// capped by the global range where the sums agree anyway), so error models whose rows reach less than 50 sizes above their
// diagonal - every model the reference's reader and `esterror` produce is a band of a few sizes - take the fused kernel; a wider
// one falls back to the per-node kernels.
// Host mirror of TilePlan (the kernel's static split of the 8-family blocks over CTAs and tiles): the family positions
// [first, first + count) of every tile of a launch over F families, as {tile index inside its CTA, first, count}.  Used by
// callers that choose the ORDER of the families (run_pvalues sorts them by window); a mismatch with TilePlan would cost speed,
// never correctness - a family's result does not depend on its position.
void fused2_tile_slots(const cafe_gpu_ctx* ctx, int F, std::vector<std::array<int, 3>>& slots) {
    slots.clear();
    const int n_mblocks = (F + 7) / 8;
    const int G = std::max(1, std::min(ctx->sm_count, (n_mblocks + 1) / 2));
    const int per_tile = fused2::TILE_M / 8;
    for (int c = 0; c < G; ++c) {
        const int mb_lo = (int)((long long)n_mblocks * c / G);
        const int n_mb = (int)((long long)n_mblocks * (c + 1) / G) - mb_lo;
        if (n_mb <= 0) continue;
        int n_tiles = (n_mb + per_tile - 1) / per_tile;
        if (n_mb >= 2 && (n_tiles & 1)) ++n_tiles;
        const int e = (n_mb / n_tiles) & ~1, r = n_mb - e * n_tiles, n2 = r >> 1;
        for (int t = 0; t < n_tiles; ++t) {
            const int m = e + (t < n2 ? 2 : 0) + ((r & 1) && t == n2 ? 1 : 0);
            const int mb0 = mb_lo + e * t + 2 * std::min(t, n2) + ((r & 1) && t > n2 ? 1 : 0);
            const int first = mb0 * 8, count = std::min(F, (mb0 + m) * 8) - first;
            if (count > 0) slots.push_back({t, first, count});
        }
    }
}

bool fused2_windowed_supported(const cafe_gpu_ctx* ctx) {
    if (!fused2_tree_supported(ctx)) return false;
    for (int e : ctx->leaf_err)
        if (e >= 0 && ctx->errs[e].max_up >= 50) return false;
    return true;
}

// Post-order schedule of the GEMMs.  A node whose two children are leaves ("cherry") never gets a vector slot; the needier
// internal child is evaluated first (Sethi–Ullman), its slot is released as soon as its parent's GEMM is issued.
static void build_schedule2(const cafe_gpu_ctx* ctx, Fused2State& st) {
    using fused2::Op;
    const int n = ctx->n_nodes;
    // matrix index of a leaf's factor rows: its branch's key, or its own error-model matrix behind the keys (launch_prune_fused2)
    auto leaf_key = [&](int leaf_node) {
        const int k = leaf_node / 2;
        const bool err = !ctx->leaf_err.empty() && ctx->leaf_err[k] >= 0;
        return err ? (int)ctx->mat_cap + k : ctx->node_key[leaf_node];
    };
    auto is_leaf = [&](int v) { return ctx->left[v] < 0; };
    auto is_cherry = [&](int v) { return !is_leaf(v) && is_leaf(ctx->left[v]) && is_leaf(ctx->right[v]); };
    auto is_virtual = [&](int v) { return is_leaf(v) || is_cherry(v); };
    std::vector<int> need(n, 0);
    std::function<int(int)> calc = [&](int v) -> int {
        if (is_virtual(v)) return need[v] = 0;
        int a = calc(ctx->left[v]), b = calc(ctx->right[v]);
        int hi = std::max(a, b), lo = std::min(a, b);
        int k = std::max(hi, lo + (hi > 0 ? 1 : 0));
        int live_children = (a > 0) + (b > 0);
        return need[v] = std::max(k, live_children + 1);
    };
    calc(ctx->root);

    st.ops.clear();
    std::vector<int> free_slots;
    int n_slots = 0;
    auto alloc = [&]() {
        if (!free_slots.empty()) { int s = free_slots.back(); free_slots.pop_back(); return s; }
        return n_slots++;
    };
    int n_cherry = 0;
    auto gemm_over = [&](Op& op, int child, int slot) {  // the GEMM operand: a stored vector or a leaf pair
        op.key = ctx->node_key[child];
        if (is_cherry(child)) {
            const int a = ctx->left[child], b = ctx->right[child];
            op.a_kind = 1; op.in_slot = n_cherry++;
            op.leaf_a1 = a / 2; op.key_a1 = leaf_key(a);
            op.leaf_a2 = b / 2; op.key_a2 = leaf_key(b);
        } else {
            op.a_kind = 0; op.in_slot = slot;
        }
    };
    std::function<int(int)> eval = [&](int v) -> int {  // returns the slot of v's vector (v internal, not a cherry)
        const int a = ctx->left[v], b = ctx->right[v];
        Op op{};
        op.is_root = (v == ctx->root);
        if (is_leaf(a) != is_leaf(b)) {
            const int gch = is_leaf(a) ? b : a, l = is_leaf(a) ? a : b;
            const int sg = is_cherry(gch) ? -1 : eval(gch);
            op.out_slot = alloc();
            gemm_over(op, gch, sg);
            op.other_kind = 1; op.leaf_o = l / 2; op.key_o = leaf_key(l);
            st.ops.push_back(op);
            if (sg >= 0) free_slots.push_back(sg);
            return op.out_slot;
        }
        // two internal children
        const int first = need[a] >= need[b] ? a : b, second = (first == a) ? b : a;
        const int s1 = is_cherry(first) ? -1 : eval(first);
        const int s2 = is_cherry(second) ? -1 : eval(second);
        op.out_slot = alloc();
        gemm_over(op, first, s1);
        op.other_kind = 0;
        st.ops.push_back(op);
        if (s1 >= 0) free_slots.push_back(s1);
        Op op2{};
        op2.is_root = op.is_root; op2.out_slot = op.out_slot;
        gemm_over(op2, second, s2);
        op2.other_kind = 2;
        st.ops.push_back(op2);
        if (s2 >= 0) free_slots.push_back(s2);
        return op.out_slot;
    };
    eval(ctx->root);
    st.n_slots = std::max(1, n_slots);
    st.n_cherry = n_cherry;
}

int launch_prune_fused2(cafe_gpu_ctx* ctx, double* d_Lroot_out) {
    Fused2Job job;
    job.counts = ctx->d_counts; job.leaf_stride = (size_t)ctx->F_pad; job.F = ctx->F; job.F_pad = ctx->F_pad;
    job.root_r0 = ctx->root_min; job.root_rows = ctx->R;
    job.posterior = true; job.d_Lroot_out = d_Lroot_out;
    return launch_prune_fused2_job(ctx, job);
}

int launch_prune_fused2_job(cafe_gpu_ctx* ctx, const Fused2Job& job) {
    using namespace fused2;
    Fused2State& st = fstate2(ctx);
    PFN_encodeTiled2 encode = get_encode_fn2();
    if (!encode) CAFE_FAIL(ctx, CAFE_GPU_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available");

    // ---- schedule (a few hundred host instructions; the keys of the branches change with every rate vector) ----
    build_schedule2(ctx, st);
    if ((int)st.ops.size() > st.ops_cap) {
        cudaFree(st.d_ops); st.d_ops = nullptr;
        st.ops_cap = std::max<int>((int)st.ops.size(), 2 * ctx->n_nodes);
        CAFE_CK(ctx, cudaMalloc(&st.d_ops, st.ops_cap * sizeof(Op)));
    }
    CAFE_CK(ctx, cudaMemcpyAsync(st.d_ops, st.ops.data(), st.ops.size() * sizeof(Op), cudaMemcpyHostToDevice, ctx->stream));

    // ---- error-model leaves: one extra transposed matrix each, behind the keys ----
    for (int k = 0; k < ctx->n_leaves; ++k) {
        const int e = ctx->leaf_err.empty() ? -1 : ctx->leaf_err[k];
        if (e < 0) continue;
        const ErrModelDev& E = ctx->errs[e];
        const size_t mat = (size_t)ctx->Sp * ctx->Sp;
        dim3 g((ctx->Sp + 255) / 256, ctx->Sp);
        k_err_leaf_matrix<<<g, 256, 0, ctx->stream>>>(ctx->d_MT + (size_t)ctx->node_key[2 * k] * mat, E.d_rowptr, E.d_col, E.d_val, E.dim,
                                                      ctx->Sp, ctx->W - 1, ctx->d_MT + ((size_t)ctx->mat_cap + k) * mat);
        ctx->launches++;
    }
    CAFE_CK(ctx, cudaGetLastError());

    // ---- geometry: one CTA per SM, every CTA at least two 8-family blocks ----
    const int n_mblocks = (job.F + 7) / 8;
    const int grid = std::max(1, std::min(ctx->sm_count, (n_mblocks + 1) / 2));
    const size_t cta_rows = (size_t)(2 * st.n_slots + 4 * st.n_cherry) * TILE_M;
    const size_t scratch_doubles = (size_t)grid * cta_rows * ctx->Vp;
    if (scratch_doubles > st.scratch_cap) {
        cudaFree(st.d_scratch); st.d_scratch = nullptr;
        CAFE_CK(ctx, cudaMalloc(&st.d_scratch, scratch_doubles * sizeof(double)));
        CAFE_CK(ctx, cudaMemsetAsync(st.d_scratch, 0, scratch_doubles * sizeof(double), ctx->stream));
        st.scratch_cap = scratch_doubles;
    }

    // ---- tensor maps (SWIZZLE_128B, zero OOB fill) ----
    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[2] = {(cuuint64_t)ctx->Vp, (cuuint64_t)grid * cta_rows};
        cuuint64_t strides[1] = {(cuuint64_t)ctx->Vp * sizeof(double)};
        cuuint32_t box[2] = {BK, TILE_M};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, st.d_scratch, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) CAFE_FAIL(ctx, CAFE_GPU_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: " + std::to_string((int)r));
    }
    // The matrix tile: rows = output sizes.  The row extent of the map is the number of sizes the op produces (W below the root,
    // root_min + R at the root), not the padded matrix: TMA fills the rows beyond it with zeros, so the accumulators of sizes that
    // do not exist are exact zeros (matrix rows in [W, S) are not zero when S > W) and the epilogue needs no per-element limit.
    auto encode_B = [&](CUtensorMap* tm, int rows) -> CUresult {
        cuuint64_t dims[3] = {(cuuint64_t)ctx->Sp, (cuuint64_t)std::min(rows, ctx->Sp), (cuuint64_t)ctx->mat_cap};
        cuuint64_t strides[2] = {(cuuint64_t)ctx->Sp * sizeof(double), (cuuint64_t)ctx->Sp * ctx->Sp * sizeof(double)};
        cuuint32_t box[3] = {BK, TN, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, ctx->d_M, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUtensorMap tmBroot;
    {
        CUresult r = encode_B(&tmB, ctx->W);
        if (r == CUDA_SUCCESS) r = encode_B(&tmBroot, job.root_r0 + job.root_rows);
        if (r != CUDA_SUCCESS) CAFE_FAIL(ctx, CAFE_GPU_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: " + std::to_string((int)r));
    }

    Params P{};
    P.ops = st.d_ops; P.n_ops = (int)st.ops.size(); P.n_slots = st.n_slots; P.n_cherry = st.n_cherry; P.F = job.F; P.F_pad = job.F_pad;
    P.scratch = st.d_scratch;
    P.W = ctx->W; P.R = job.root_rows; P.root_min = job.root_r0; P.Sp = ctx->Sp; P.Vp = ctx->Vp; P.n_mblocks = n_mblocks;
    P.MT = ctx->d_MT; P.counts = job.counts; P.leaf_stride = job.leaf_stride;
    P.colmax = job.d_colmax; P.root_pick = job.d_root_pick; P.L0_out = job.d_L0_out; P.root_need = job.d_root_need;
    if (job.posterior) {
        P.logprior = ctx->d_logprior; P.prior_mant = ctx->d_prior_mant; P.prior_exp = ctx->d_prior_exp;
        P.logpost = ctx->d_logpost; P.maxlik = ctx->d_maxlik; P.argmax = ctx->d_argmax;
    }
    P.Lroot_out = job.d_Lroot_out;
    if (job.root_rows > ctx->Vp || job.root_r0 + job.root_rows > ctx->Sp)
        CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "fused pruning: root rows exceed the vector / matrix");

    if (const char* d = std::getenv("CAFE_GPU_DBG")) P.dbg = std::atoi(d);
    const size_t smem_bytes = (size_t)NSTAGE * STAGE_BYTES + C_BYTES + sizeof(Ctl) + 1024;
    if (!st.attr_set) {
        CAFE_CK(ctx, cudaFuncSetAttribute(k_prune_fused2<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        CAFE_CK(ctx, cudaFuncSetAttribute(k_prune_fused2<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        CAFE_CK(ctx, cudaFuncSetAttribute(k_prune_fused2<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        st.attr_set = true;
    }
    const char* trace_path = std::getenv("CAFE_GPU_TRACE");
    long long* d_trace = nullptr;
    if (trace_path) {
        CAFE_CK(ctx, cudaMalloc(&d_trace, ((size_t)grid * 4 + 128 + 12288) * sizeof(long long)));
        CAFE_CK(ctx, cudaMemsetAsync(d_trace, 0, ((size_t)grid * 4 + 128 + 12288) * sizeof(long long), ctx->stream));
        P.cta_times = d_trace;
        P.warp_prof = d_trace + (size_t)grid * 4;
        P.timeline = P.warp_prof + 128;
    }
    const bool windowed = job.d_colmax != nullptr;
    if (job.d_root_pick && !windowed) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "fused pruning: a root pick needs the per-family windows");
    if (windowed) k_prune_fused2<false, true><<<grid, THREADS, smem_bytes, ctx->stream>>>(tmA, tmB, tmBroot, P);
    else if (trace_path) k_prune_fused2<true, false><<<grid, THREADS, smem_bytes, ctx->stream>>>(tmA, tmB, tmBroot, P);
    else k_prune_fused2<false, false><<<grid, THREADS, smem_bytes, ctx->stream>>>(tmA, tmB, tmBroot, P);
    ctx->launches++;
    CAFE_CK(ctx, cudaGetLastError());
    if (trace_path) {  // debug only: synchronous dump "cta <i> <smid> <start ns> <end ns> <8-family blocks>"
        std::vector<long long> h((size_t)grid * 4 + 128 + 12288);
        CAFE_CK(ctx, cudaMemcpyAsync(h.data(), d_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(d_trace);
        if (FILE* fp = std::fopen(trace_path, "w")) {
            for (int c = 0; c < grid; ++c) std::fprintf(fp, "cta %d %lld %lld %lld %lld\n", c, h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
            // consumers: total, K loops, wait ring, wait C tile, epilogues, root epilogues, -, K loops of leaf-pair items | producer: total, wait done, wait empty
            // epilogue manager: total, prep, wait consumers, store, items            (cycles, CTA 0)
            for (int w = 0; w < 16; ++w) {
                const long long* o = &h[(size_t)grid * 4 + w * 8];
                std::fprintf(fp, "warp %d %lld %lld %lld %lld %lld %lld %lld %lld\n", w, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
            }
            // per pass of warps 0 and 4 (the two DMMA warps of sub-partition 0): K loop start, K loop end, C tile ready, epilogue end
            for (int g2 = 0; g2 < 3; ++g2)
                for (int it = 0; it < 1024; ++it) {
                    const long long* o = &h[(size_t)grid * 4 + 128 + ((size_t)g2 * 1024 + it) * 4];
                    if (o[0]) std::fprintf(fp, "tl %d %d %lld %lld %lld %lld\n", g2, it, o[0], o[1], o[2], o[3]);
                }
            std::fclose(fp);
        }
    }
    return CAFE_GPU_OK;
}
