// prune_fused.cu — K2, fused: ONE persistent kernel walks the whole species tree for every family.
//
// Replaces, per objective evaluation, the reference's F x (2n-2) calls of square_matrix_multiply
// (libtree/birthdeath.c:163-182) under compute_internal_node_likelihood (cafe/cafe_tree.c:226-271),
// initialize_leaf_likelihoods (:191-211) and compute_posterior's root reduction
// (cafe/lambda.cpp:657-689).
//
// Shape of the work: per internal tree edge a batched matvec = GEMM
//      Out[f][i] = sum_j M_edge[r0+i][j] * L_child[f][j]         (f: families, i,j: family sizes)
// executed on the fp64 tensor pipe (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4; tcgen05 has no fp64 kind).
//
// Kernel structure (one CTA per SM, persistent):
//   * warp 8      : TMA producer. Streams K-blocks of the child vectors (A, 64 families x 16 sizes) and of
//                   the transition matrix (B, 256 rows x 16 cols) into a 4-stage shared-memory ring with
//                   cp.async.bulk.tensor + mbarrier complete_tx; SWIZZLE_128B makes the DMMA fragment
//                   loads conflict-free (see mma_row_perm in common.cuh).
//   * warps 0..7  : DMMA consumers, all along N: warp tile 64 families x 32 sizes (64 accumulators / lane),
//                   CTA tile 64 x 256 => the whole likelihood vector of config 2 (W=251) in one pass.
//   * epilogue    : child product fused — the sibling's factor is either a column gather from the
//                   transposed matrix (leaf edges, cafe_tree.c:204-210) or the partial product already
//                   stored; at the root the posterior max/argmax reduction of lambda.cpp:670-686.
//   * a CTA owns a contiguous range of 8-family blocks and keeps the node vectors of its families in a
//     private scratch region (L2 resident); it interleaves TWO 64-family half-tiles so that the
//     producer never waits for the epilogue of the vector it has to stream next.
//
// HBM/L2 picture per objective evaluation (config 2): matrices 20 x 0.5 MB are re-read from L2 by every
// half-tile (~0.7 GB/SM-second of L2 traffic, far below the L2 cap); node vectors never leave L2;
// HBM traffic is the counts in (4 B per leaf per family) and 20 B per family out.  The binding roof is
// the fp64 pipe.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace fused {

// Three CTAs per SM (3 DMMA warps per sub-partition: one warp alone reaches only ~73 % of the DMMA issue rate,
// three reach ~97 %, tools/fp64_peak.cu): while one is in an epilogue (sibling-factor gathers from L2, stores) the other keeps the
// DMMA pipe busy.  Each CTA: 1 DMMA warpgroup (4 warps, one per SM sub-partition) + 1 producer warpgroup.
constexpr int HM = 32;            // families per half-tile
constexpr int TN = 128;           // output sizes per N pass
constexpr int BK = 16;            // sizes per K block (16 doubles = 128 B = one swizzle row)
constexpr int NSTAGE = 3;
constexpr int A_BYTES = HM * 128;   // 4 KB
constexpr int B_BYTES = TN * 128;   // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int N_CONSUMER_WARPS = 4;
constexpr int THREADS = (N_CONSUMER_WARPS + 4) * 32;  // DMMA warpgroup + producer warpgroup (only its first lane works)
constexpr int CTAS_PER_SM = 3;
// setmaxnreg split of the launch allocation (80 regs x 256 threads): 4*136 + 4*24 = 8*80
constexpr int REGS_PRODUCER = 24, REGS_CONSUMER = 136;
constexpr int MB = HM / 8;        // 4 m-blocks per half-tile
constexpr int NB = 4;             // n-blocks per warp (32 sizes)
constexpr int WCOLS = NB * 8;     // sizes per warp
constexpr int MAX_SLOTS = 16;

struct Op {              // one node of the post-order schedule
    int kind;            // 0: both children are leaves, 1: GEMM over an internal child
    int is_root;
    int key;             // matrix of the GEMM child's branch
    int in_slot, out_slot;
    int other_kind;      // 0 none, 1 leaf sibling, 2 multiply into out_slot
    int leaf_a, key_a;   // leaf sibling (kind 1) / first leaf (kind 0)
    int leaf_b, key_b;   // second leaf (kind 0)
};

struct LeafErr {         // sparse error rows of one leaf (nullptr = one-hot leaf)
    const int* rowptr;
    const int* col;
    const double* val;
};

struct Params {
    const Op* ops;
    int n_ops;
    int n_slots;
    int F, F_pad;
    int W, R, root_min;
    int Sp, Vp;
    int n_mblocks;               // ceil(F / 8)
    int counts_in_window;        // every observed size < W (then a one-hot leaf never falls outside the matvec columns)
    const double* MT;            // [D][Sp][Sp] transposed matrices
    const int* counts;           // [n_leaves][F_pad]
    const LeafErr* leaf_err;     // [n_leaves]
    const double* logprior;      // [R]
    double* scratch;             // [grid][2][n_slots][HM][Vp]
    double* logpost;             // [F_pad]
    double* maxlik;
    int* argmax;
    double* Lroot_out;           // nullable, [F][R]
    long long* trace;            // nullable debug trace of CTA 0: [event][warp][4] clock64 stamps (CAFE_GPU_TRACE)
    long long* cta_times;        // nullable debug: [grid][4] = smid, start ns, end ns, 8-family blocks
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
// generic-proxy writes (st.global of a node vector) -> async-proxy reads (the TMA that streams it next)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(N_CONSUMER_WARPS * 32) : "memory"); }

// error-model leaf (cafe_tree.c:196-203): sparse combination of matrix columns, ascending true size.  Cold path.
__device__ __noinline__ double leaf_factor_err(const double* __restrict__ MT, const int* rowptr, const int* col, const double* val,
                                               int Sp, int count, int colmax, int r) {
    double s = 0.0;
    for (int k = rowptr[count]; k < rowptr[count + 1]; ++k) {
        const int j = col[k];
        if (j <= colmax) s = __dadd_rn(s, __dmul_rn(MT[(size_t)j * Sp + r], val[k]));
    }
    return s;
}
__device__ __forceinline__ double leaf_factor(const double* __restrict__ MT, const LeafErr& E, int Sp, int count, int colmax, int r) {
    if (E.rowptr == nullptr) return (count <= colmax) ? MT[(size_t)count * Sp + r] : 0.0;
    return leaf_factor_err(MT, E.rowptr, E.col, E.val, Sp, count, colmax, r);
}

struct SharedCtl {
    uint64_t full[NSTAGE];
    uint64_t empty[NSTAGE];
    volatile int done[2];                 // completed ops per half-tile (consumer -> producer)
    double red_ml[N_CONSUMER_WARPS][HM];  // root reduction scratch
    double red_mp[N_CONSUMER_WARPS][HM];
    int red_am[N_CONSUMER_WARPS][HM];
    double run_ml[HM], run_mp[HM];
    int run_am[HM];
};

// this CTA's families: a contiguous range of 8-family blocks, cut into pairs of half-tiles
struct TilePlan {
    int mb_lo, n_mb, n_pairs;
    __device__ TilePlan(const Params& P) {
        const int G = gridDim.x, c = blockIdx.x;
        mb_lo = (int)((long long)P.n_mblocks * c / G);
        n_mb = (int)((long long)P.n_mblocks * (c + 1) / G) - mb_lo;
        n_pairs = (n_mb + 2 * MB - 1) / (2 * MB);
    }
    // half-tile h of pair `pair`: number of valid 8-family blocks and first family
    __device__ void half(int pair, int h, int& mb_valid, int& f0) const {
        const int p_lo = mb_lo + (int)((long long)n_mb * pair / n_pairs), p_hi = mb_lo + (int)((long long)n_mb * (pair + 1) / n_pairs);
        const int m0 = (p_hi - p_lo + 1) / 2;
        mb_valid = h ? (p_hi - p_lo) / 2 : m0;
        f0 = (h ? p_lo + m0 : p_lo) * 8;
    }
};

// ================================ TMA producer (one lane) ================================
__device__ __forceinline__ void producer_main(const CUtensorMap* tmA, const CUtensorMap* tmB, const Params& P,
                                              unsigned char* stage_base, SharedCtl* ctl) {
    const TilePlan plan(P);
    const int scratch_row0 = blockIdx.x * 2 * P.n_slots * HM;  // row of this CTA in the A tensor map
    const int n_kblocks = (P.W + BK - 1) / BK;
    uint32_t stage = 0, phase = 0;
    int ops_done_base = 0;
    for (int pair = 0; pair < plan.n_pairs; ++pair) {
        for (int oi = 0; oi < P.n_ops; ++oi) {
            const Op op = P.ops[oi];
            if (op.kind != 1) continue;
            const int r0 = op.is_root ? P.root_min : 0;
            const int nrows = op.is_root ? P.R : P.W;
            const int n_chunks = (nrows + TN - 1) / TN;
            for (int h = 0; h < 2; ++h) {
                int mb_valid, f0;
                plan.half(pair, h, mb_valid, f0);
                if (mb_valid == 0) continue;
                // the vector to stream was written by an earlier op of this half-tile: wait for it
                while (ctl->done[h] < ops_done_base + oi) { __nanosleep(20); }
                __threadfence_block();
                const int a_row = scratch_row0 + (h * P.n_slots + op.in_slot) * HM;
                for (int ch = 0; ch < n_chunks; ++ch) {
                    for (int kb = 0; kb < n_kblocks; ++kb) {
                        mbar_wait(&ctl->empty[stage], phase ^ 1);
                        unsigned char* sA = stage_base + stage * STAGE_BYTES;
                        mbar_arrive_expect_tx(&ctl->full[stage], STAGE_BYTES);
                        tma_load_2d(sA, tmA, kb * BK, a_row, &ctl->full[stage]);
                        tma_load_3d(sA + A_BYTES, tmB, kb * BK, r0 + ch * TN, op.key, &ctl->full[stage]);
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        ops_done_base += P.n_ops;
    }
}

// ================================ DMMA consumers (8 warps) ================================
// One k4-step of the warp tile: 4 B fragments, then per 8-family block one A fragment and 4 DMMAs.
// No predicates inside: MBV is a compile-time count, so the DMMA stream is straight-line code.
template <int MBV>
__device__ __forceinline__ void kstep(double (&acc)[MB][NB][2], const unsigned char* sA, const unsigned char* sB, int off) {
    double b[NB];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) b[nb] = *reinterpret_cast<const double*>(sB + nb * 1024 + off);
#pragma unroll
    for (int mb = 0; mb < MBV; ++mb) {
        const double a = *reinterpret_cast<const double*>(sA + mb * 1024 + off);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) dmma_884(acc[mb][nb][0], acc[mb][nb][1], a, b[nb]);
    }
}

// K loop of one N pass: consume n_kblocks ring stages.
template <int MBV>
__device__ __forceinline__ void gemm_kblocks(double (&acc)[MB][NB][2], unsigned char* stage_base, SharedCtl* ctl, uint32_t& stage,
                                             uint32_t& phase, int n_kblocks, int tail_steps, int warp, int lane, int pg, int q,
                                             bool warp_has_columns) {
    const int off0 = pg * 128 + ((q & 1) << 3);
    const int hi = q >> 1;
    for (int kb = 0; kb < n_kblocks; ++kb) {
        mbar_wait(&ctl->full[stage], phase);
        const unsigned char* sA = stage_base + stage * STAGE_BYTES;
        const unsigned char* sB = sA + A_BYTES + warp * WCOLS * 128;
        if (warp_has_columns) {
            if (kb + 1 < n_kblocks || tail_steps == 4) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) kstep<MBV>(acc, sA, sB, off0 + (((2 * kk + hi) ^ pg) << 4));
            } else {
                for (int kk = 0; kk < tail_steps; ++kk) kstep<MBV>(acc, sA, sB, off0 + (((2 * kk + hi) ^ pg) << 4));
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->empty[stage]);
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
    }
}

__device__ __forceinline__ void consumer_main(const Params& P, unsigned char* stage_base, SharedCtl* ctl) {
    const TilePlan plan(P);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const int pg = mma_row_perm(g);
    const int pc0 = mma_row_perm(2 * q), pc1 = mma_row_perm(2 * q + 1);
    const size_t half_stride = (size_t)P.n_slots * HM * P.Vp;  // doubles per half-tile scratch
    double* my_scratch = P.scratch + (size_t)blockIdx.x * 2 * half_stride;
    const int n_kblocks = (P.W + BK - 1) / BK;
    const int tail_steps = ((P.W - (n_kblocks - 1) * BK) + 3) >> 2;  // k4-steps of the last K block (1..4)

    uint32_t stage = 0, phase = 0;
    int ops_done_base = 0;
    for (int pair = 0; pair < plan.n_pairs; ++pair) {
        for (int oi = 0; oi < P.n_ops; ++oi) {
            const Op op = P.ops[oi];
            const int r0 = op.is_root ? P.root_min : 0;
            const int nrows = op.is_root ? P.R : P.W;
            const int n_chunks = (nrows + TN - 1) / TN;
            for (int h = 0; h < 2; ++h) {
                int mb_valid, f0;
                plan.half(pair, h, mb_valid, f0);
                if (mb_valid == 0) continue;
                double* out = my_scratch + h * half_stride + (size_t)op.out_slot * HM * P.Vp;
                const bool tracing = P.trace != nullptr && blockIdx.x == 0 && lane == 0;
                long long* tr = tracing ? P.trace + ((size_t)((pair * P.n_ops + oi) * 2 + h) * N_CONSUMER_WARPS + warp) * 4 : nullptr;
                if (tracing) { tr[0] = clock64(); tr[1] = tr[0]; }

                if (op.kind == 0) {
                    // ---- both children are leaves: product of two gathered columns, warp per family ----
                    const double* __restrict__ MTa = P.MT + (size_t)op.key_a * P.Sp * P.Sp;
                    const double* __restrict__ MTb = P.MT + (size_t)op.key_b * P.Sp * P.Sp;
                    const LeafErr Ea = P.leaf_err[op.leaf_a], Eb = P.leaf_err[op.leaf_b];
                    const bool lp_fast = !op.is_root && Ea.rowptr == nullptr && Eb.rowptr == nullptr && P.counts_in_window;
                    if (lp_fast) {
                        // common case (one-hot leaves): two gathered matrix columns per family, two families in flight per
                        // warp => 32 independent 8-byte loads per lane before the first store
                        const int n_rows = mb_valid * 8;
                        for (int row = warp * 2; row < n_rows; row += 2 * N_CONSUMER_WARPS) {
                            const double* pa[2]; const double* pb[2];
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const int f = min(f0 + row + u, P.F - 1);
                                const int ca = __ldg(P.counts + (size_t)op.leaf_a * P.F_pad + f), cb = __ldg(P.counts + (size_t)op.leaf_b * P.F_pad + f);
                                pa[u] = MTa + (size_t)ca * P.Sp + lane;
                                pb[u] = MTb + (size_t)cb * P.Sp + lane;
                            }
                            for (int ib = 0; ib < P.Vp; ib += 256) {
                                double va[2][8], vb[2][8];
#pragma unroll
                                for (int u = 0; u < 2; ++u)
#pragma unroll
                                    for (int c8 = 0; c8 < 8; ++c8) {
                                        const int i = ib + c8 * 32;  // + lane folded into the pointers; Sp >= Vp, zero padded beyond W
                                        const bool in = i + lane < P.Vp;
                                        va[u][c8] = in ? __ldg(pa[u] + i) : 0.0;
                                        vb[u][c8] = in ? __ldg(pb[u] + i) : 0.0;
                                    }
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    if (row + u < n_rows) {
                                        double* o = out + (size_t)(row + u) * P.Vp + lane;
#pragma unroll
                                        for (int c8 = 0; c8 < 8; ++c8) {
                                            const int i = ib + c8 * 32;
                                            // sizes >= W must stay exact zeros (the matrices are wider than W when S > W)
                                            if (i + lane < P.Vp) o[i] = (i + lane < P.W) ? va[u][c8] * vb[u][c8] : 0.0;
                                        }
                                    }
                                }
                            }
                        }
                    } else
                    for (int row = warp; row < mb_valid * 8; row += N_CONSUMER_WARPS) {
                        const int f = f0 + row;
                        double* o = out + (size_t)row * P.Vp;
                        if (f < P.F) {
                            const int ca = P.counts[(size_t)op.leaf_a * P.F_pad + f], cb = P.counts[(size_t)op.leaf_b * P.F_pad + f];
                            double ml = -1.0, mp = -INFINITY; int am = 0x7fffffff;
                            for (int ib = 0; ib < P.Vp; ib += 256) {  // 8 independent column groups per batch
                                double va[8], vb[8];
#pragma unroll
                                for (int u = 0; u < 8; ++u) {
                                    const int i = ib + u * 32 + lane;
                                    va[u] = 0.0; vb[u] = 0.0;
                                    if (i < nrows) {
                                        va[u] = leaf_factor(MTa, Ea, P.Sp, ca, P.W - 1, r0 + i);
                                        vb[u] = leaf_factor(MTb, Eb, P.Sp, cb, P.W - 1, r0 + i);
                                    }
                                }
#pragma unroll
                                for (int u = 0; u < 8; ++u) {
                                    const int i = ib + u * 32 + lane;
                                    const double v = va[u] * vb[u];
                                    if (!op.is_root) { if (i < P.Vp) o[i] = v; }
                                    else if (i < nrows) {
                                        if (P.Lroot_out) P.Lroot_out[(size_t)f * P.R + i] = v;
                                        if (v > ml) { ml = v; am = i; }
                                        const double x = log(v) + P.logprior[i];
                                        if (x > mp) mp = x;
                                    }
                                }
                            }
                            if (op.is_root) {  // two-leaf tree: the root itself is a leaf pair
#pragma unroll
                                for (int off = 16; off > 0; off >>= 1) {
                                    double oml = __shfl_xor_sync(0xffffffffu, ml, off); int oam = __shfl_xor_sync(0xffffffffu, am, off);
                                    double omp = __shfl_xor_sync(0xffffffffu, mp, off);
                                    if (oml > ml || (oml == ml && oam < am)) { ml = oml; am = oam; }
                                    if (omp > mp) mp = omp;
                                }
                                if (lane == 0) { P.logpost[f] = log(exp(mp)); P.maxlik[f] = ml; P.argmax[f] = am; }
                            }
                        } else if (!op.is_root) {
                            for (int i = lane; i < P.Vp; i += 32) o[i] = 0.0;
                        }
                    }
                } else {
                    // ---- GEMM over the internal child, the whole vector in passes of 256 sizes ----
                    const double* __restrict__ MTl = P.MT + (size_t)(op.other_kind == 1 ? op.key_a : 0) * P.Sp * P.Sp;
                    LeafErr El{nullptr, nullptr, nullptr};
                    if (op.other_kind == 1) El = P.leaf_err[op.leaf_a];
                    const bool reduce_now = op.is_root && op.other_kind != 0;
                    if (reduce_now && threadIdx.x < HM) { ctl->run_ml[threadIdx.x] = -1.0; ctl->run_mp[threadIdx.x] = -INFINITY; ctl->run_am[threadIdx.x] = 0x7fffffff; }
                    // observed sizes of the leaf sibling for this lane's 8 family rows (latency hidden by the K loop)
                    int cnt[MB];
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb) {
                        const int f = f0 + mb * 8 + pg;
                        cnt[mb] = (op.other_kind == 1 && mb < mb_valid && f < P.F) ? __ldg(P.counts + (size_t)op.leaf_a * P.F_pad + f) : 0;
                    }

                    for (int ch = 0; ch < n_chunks; ++ch) {
                        const int n0 = ch * TN + warp * WCOLS;  // first output size of this warp
                        const bool warp_has_columns = n0 < nrows;
                        double acc[MB][NB][2];
#pragma unroll
                        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;

                        switch (mb_valid) {
                            case 4: gemm_kblocks<4>(acc, stage_base, ctl, stage, phase, n_kblocks, tail_steps, warp, lane, pg, q, warp_has_columns); break;
                            case 3: gemm_kblocks<3>(acc, stage_base, ctl, stage, phase, n_kblocks, tail_steps, warp, lane, pg, q, warp_has_columns); break;
                            case 2: gemm_kblocks<2>(acc, stage_base, ctl, stage, phase, n_kblocks, tail_steps, warp, lane, pg, q, warp_has_columns); break;
                            default: gemm_kblocks<1>(acc, stage_base, ctl, stage, phase, n_kblocks, tail_steps, warp, lane, pg, q, warp_has_columns); break;
                        }

                        if (tracing) tr[1] = clock64();
                        // ---------------- epilogue of this pass ----------------
                        // Unguarded path: sizes in [nrows, Vp) come out as exact zeros by themselves (matrix rows/columns
                        // beyond S are zero padding, stored partials there are zero), rows beyond F are private garbage.
                        const bool fast = (n0 + WCOLS <= P.Vp) && El.rowptr == nullptr && P.counts_in_window;
                        if (!reduce_now && fast) {
                            // common case: every element of the warp tile exists; sibling factors are fetched in
                            // batches of 16 independent loads (two 8-family blocks) before they are consumed
#pragma unroll
                            for (int mb = 0; mb < MB; mb += 2) {
                                if (mb < mb_valid) {
                                    double fac[2][NB][2];
#pragma unroll
                                    for (int u = 0; u < 2; ++u) {
                                        const int row = (mb + u) * 8 + pg;
                                        const double* src = (op.other_kind == 1) ? MTl + (size_t)cnt[mb + u] * P.Sp + r0 + n0
                                                                                 : out + (size_t)row * P.Vp + n0;
#pragma unroll
                                        for (int nb = 0; nb < NB; ++nb) {
                                            fac[u][nb][0] = 1.0; fac[u][nb][1] = 1.0;
                                            if (op.other_kind != 0 && mb + u < mb_valid) {
                                                fac[u][nb][0] = src[nb * 8 + pc0];
                                                fac[u][nb][1] = src[nb * 8 + pc1];
                                            }
                                        }
                                    }
#pragma unroll
                                    for (int u = 0; u < 2; ++u) {
                                        if (mb + u < mb_valid) {
                                            double* o = out + (size_t)((mb + u) * 8 + pg) * P.Vp + n0;
#pragma unroll
                                            for (int nb = 0; nb < NB; ++nb) {
                                                // sizes >= nrows must stay exact zeros: matrix rows in [W, S) are not zero when S > W
                                                o[nb * 8 + pc0] = (n0 + nb * 8 + pc0 < nrows) ? acc[mb + u][nb][0] * fac[u][nb][0] : 0.0;
                                                o[nb * 8 + pc1] = (n0 + nb * 8 + pc1 < nrows) ? acc[mb + u][nb][1] * fac[u][nb][1] : 0.0;
                                            }
                                        }
                                    }
                                }
                            }
                        } else if (!reduce_now) {  // edge tiles / error-model leaves: fully guarded, loads still batched
#pragma unroll
                            for (int mb = 0; mb < MB; ++mb) {
                                if (mb < mb_valid) {
                                    const int row = mb * 8 + pg, f = f0 + row;
                                    double* o = out + (size_t)row * P.Vp;
                                    double fac[NB][2];
#pragma unroll
                                    for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
                                        for (int hh = 0; hh < 2; ++hh) {
                                            const int i = n0 + nb * 8 + (hh ? pc1 : pc0);
                                            fac[nb][hh] = 0.0;
                                            if (i < nrows && f < P.F) {
                                                fac[nb][hh] = 1.0;
                                                if (op.other_kind == 1) fac[nb][hh] = leaf_factor(MTl, El, P.Sp, cnt[mb], P.W - 1, r0 + i);
                                                else if (op.other_kind == 2) fac[nb][hh] = o[i];
                                            }
                                        }
                                    }
#pragma unroll
                                    for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
                                        for (int hh = 0; hh < 2; ++hh) {
                                            const int i = n0 + nb * 8 + (hh ? pc1 : pc0);
                                            if (i < P.Vp) o[i] = (i < nrows && f < P.F) ? acc[mb][nb][hh] * fac[nb][hh] : 0.0;
                                        }
                                    }
                                }
                            }
                        } else {
                            // root: L[i] = acc * other; max/argmax of L and max of log L + log prior (lambda.cpp:670-686)
#pragma unroll
                            for (int mb = 0; mb < MB; ++mb) {
                                double ml = -1.0, mp = -INFINITY; int am = 0x7fffffff;
                                const int row = mb * 8 + pg, f = f0 + row;
                                if (mb < mb_valid && f < P.F) {
                                    const double* o = out + (size_t)row * P.Vp;
                                    double fac[NB][2];
#pragma unroll
                                    for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
                                        for (int hh = 0; hh < 2; ++hh) {
                                            const int i = n0 + nb * 8 + (hh ? pc1 : pc0);
                                            fac[nb][hh] = 0.0;
                                            if (i < nrows) fac[nb][hh] = (op.other_kind == 1) ? leaf_factor(MTl, El, P.Sp, cnt[mb], P.W - 1, r0 + i) : o[i];
                                        }
                                    }
#pragma unroll
                                    for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
                                        for (int hh = 0; hh < 2; ++hh) {
                                            const int i = n0 + nb * 8 + (hh ? pc1 : pc0);
                                            if (i < nrows) {
                                                const double v = acc[mb][nb][hh] * fac[nb][hh];
                                                if (P.Lroot_out) P.Lroot_out[(size_t)f * P.R + i] = v;
                                                if (v > ml || (v == ml && i < am)) { ml = v; am = i; }
                                                const double x = log(v) + P.logprior[i];
                                                if (x > mp) mp = x;
                                            }
                                        }
                                    }
                                }
                                // the 4 lanes of a quad hold the same family row
#pragma unroll
                                for (int off = 1; off <= 2; off <<= 1) {
                                    double oml = __shfl_xor_sync(0xffffffffu, ml, off); int oam = __shfl_xor_sync(0xffffffffu, am, off);
                                    double omp = __shfl_xor_sync(0xffffffffu, mp, off);
                                    if (oml > ml || (oml == ml && oam < am)) { ml = oml; am = oam; }
                                    if (omp > mp) mp = omp;
                                }
                                if (q == 0) { ctl->red_ml[warp][row] = ml; ctl->red_mp[warp][row] = mp; ctl->red_am[warp][row] = am; }
                            }
                            consumer_bar();
                            if (threadIdx.x < HM) {
                                const int row = threadIdx.x;
                                double ml = ctl->run_ml[row], mp = ctl->run_mp[row]; int am = ctl->run_am[row];
                                for (int w = 0; w < N_CONSUMER_WARPS; ++w) {
                                    double oml = ctl->red_ml[w][row], omp = ctl->red_mp[w][row]; int oam = ctl->red_am[w][row];
                                    if (oml > ml || (oml == ml && oam < am)) { ml = oml; am = oam; }
                                    if (omp > mp) mp = omp;
                                }
                                ctl->run_ml[row] = ml; ctl->run_mp[row] = mp; ctl->run_am[row] = am;
                                const int f = f0 + row;
                                if (ch == n_chunks - 1 && row < mb_valid * 8 && f < P.F) {
                                    // max_j exp(log L + log prior) == exp(max_j(log L + log prior)); its log is the family's term
                                    P.logpost[f] = log(exp(mp)); P.maxlik[f] = ml; P.argmax[f] = am;
                                }
                            }
                            consumer_bar();
                        }
                    }
                }
                // publish: this half-tile finished op `oi` (its vector may be streamed by TMA from now on)
                if (tracing) tr[2] = clock64();
                fence_proxy_async();
                consumer_bar();
                if (tracing) tr[3] = clock64();
                if (threadIdx.x == 0) { __threadfence_block(); ctl->done[h] = ops_done_base + oi + 1; }
            }
        }
        ops_done_base += P.n_ops;
    }
}

__global__ void __launch_bounds__(THREADS, CTAS_PER_SM)
k_prune_fused(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params P) {
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B tiles must start on a 1024-byte boundary of the shared window
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* stage_base = smem;  // NSTAGE x (A | B), each 1024-aligned
    SharedCtl* ctl = reinterpret_cast<SharedCtl*>(smem + NSTAGE * STAGE_BYTES);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], N_CONSUMER_WARPS); }
        ctl->done[0] = ctl->done[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Register re-partition (Hopper+/Blackwell idiom): the producer warpgroup gives its registers back, the two
    // DMMA warpgroups take them — 64 fp64 accumulators per lane do not fit the uniform 168-register split.
    // The two roles never share control flow after this point.
    if (threadIdx.x >= N_CONSUMER_WARPS * 32) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
        if (threadIdx.x == N_CONSUMER_WARPS * 32) producer_main(&tmA, &tmB, P, stage_base, ctl);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_CONSUMER));
        long long t_start = 0;
        if (P.cta_times && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
        consumer_main(P, stage_base, ctl);
        if (P.cta_times && threadIdx.x == 0) {
            long long t_end; unsigned smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const TilePlan plan(P);
            long long* o = P.cta_times + (size_t)blockIdx.x * 4;
            o[0] = smid; o[1] = t_start; o[2] = t_end; o[3] = plan.n_mb;
        }
    }
}

}  // namespace fused

// =================================================================================================
// host side
// =================================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

struct FusedState {
    fused::Op* d_ops = nullptr; int ops_cap = 0;
    fused::LeafErr* d_leaf_err = nullptr; int leaf_cap = 0;
    double* d_scratch = nullptr; size_t scratch_cap = 0;
    int grid = 0;
    bool attr_set = false;
};
static FusedState& fstate(cafe_gpu_ctx* ctx) {
    if (!ctx->fused_state) ctx->fused_state = new FusedState();
    return *static_cast<FusedState*>(ctx->fused_state);
}
void fused_release(cafe_gpu_ctx* ctx) {
    if (!ctx->fused_state) return;
    FusedState* s = static_cast<FusedState*>(ctx->fused_state);
    cudaFree(s->d_ops); cudaFree(s->d_leaf_err); cudaFree(s->d_scratch);
    delete s;
    ctx->fused_state = nullptr;
}

bool fused_supported(const cafe_gpu_ctx* ctx) {
    return ctx->n_slots <= fused::MAX_SLOTS && get_encode_fn() != nullptr;
}

int launch_prune_fused(cafe_gpu_ctx* ctx, double* d_Lroot_out) {
    using namespace fused;
    FusedState& st = fstate(ctx);
    PFN_encodeTiled encode = get_encode_fn();
    if (!encode) CAFE_FAIL(ctx, CAFE_GPU_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available");

    // ---- schedule and per-leaf error rows ----
    std::vector<Op> ops(ctx->ops.size());
    for (size_t i = 0; i < ops.size(); ++i) {
        const PruneOp& s = ctx->ops[i];
        Op o{};
        o.kind = s.gemm_child < 0 ? 0 : 1; o.is_root = s.is_root; o.key = s.gemm_key; o.in_slot = s.in_slot; o.out_slot = s.out_slot;
        o.other_kind = s.other_kind; o.leaf_a = s.leaf_a < 0 ? 0 : s.leaf_a; o.key_a = s.leaf_a_key < 0 ? 0 : s.leaf_a_key;
        o.leaf_b = s.leaf_b < 0 ? 0 : s.leaf_b; o.key_b = s.leaf_b_key < 0 ? 0 : s.leaf_b_key;
        ops[i] = o;
    }
    if ((int)ops.size() > st.ops_cap) {
        cudaFree(st.d_ops); st.d_ops = nullptr;
        st.ops_cap = std::max<int>((int)ops.size(), 2 * ctx->n_nodes);
        CAFE_CK(ctx, cudaMalloc(&st.d_ops, st.ops_cap * sizeof(Op)));
    }
    CAFE_CK(ctx, cudaMemcpyAsync(st.d_ops, ops.data(), ops.size() * sizeof(Op), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<LeafErr> le(ctx->n_leaves, LeafErr{nullptr, nullptr, nullptr});
    for (int k = 0; k < ctx->n_leaves; ++k) {
        int e = ctx->leaf_err.empty() ? -1 : ctx->leaf_err[k];
        if (e >= 0) le[k] = LeafErr{ctx->errs[e].d_rowptr, ctx->errs[e].d_col, ctx->errs[e].d_val};
    }
    if (ctx->n_leaves > st.leaf_cap) {
        cudaFree(st.d_leaf_err); st.d_leaf_err = nullptr;
        st.leaf_cap = ctx->n_leaves;
        CAFE_CK(ctx, cudaMalloc(&st.d_leaf_err, st.leaf_cap * sizeof(LeafErr)));
    }
    CAFE_CK(ctx, cudaMemcpyAsync(st.d_leaf_err, le.data(), le.size() * sizeof(LeafErr), cudaMemcpyHostToDevice, ctx->stream));

    // ---- geometry ----
    const int n_mblocks = (ctx->F + 7) / 8;
    const int grid = std::max(1, std::min(CTAS_PER_SM * ctx->sm_count, (n_mblocks + 2 * MB - 1) / (2 * MB)));
    const size_t scratch_doubles = (size_t)grid * 2 * ctx->n_slots * HM * ctx->Vp;
    if (scratch_doubles > st.scratch_cap) {
        cudaFree(st.d_scratch); st.d_scratch = nullptr;
        CAFE_CK(ctx, cudaMalloc(&st.d_scratch, scratch_doubles * sizeof(double)));
        CAFE_CK(ctx, cudaMemsetAsync(st.d_scratch, 0, scratch_doubles * sizeof(double), ctx->stream));
        st.scratch_cap = scratch_doubles;
    }

    // ---- tensor maps (SWIZZLE_128B, zero OOB fill) ----
    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[2] = {(cuuint64_t)ctx->Vp, (cuuint64_t)grid * 2 * ctx->n_slots * HM};
        cuuint64_t strides[1] = {(cuuint64_t)ctx->Vp * sizeof(double)};
        cuuint32_t box[2] = {BK, HM};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, st.d_scratch, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) CAFE_FAIL(ctx, CAFE_GPU_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: " + std::to_string((int)r));
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)ctx->Sp, (cuuint64_t)ctx->Sp, (cuuint64_t)ctx->mat_cap};
        cuuint64_t strides[2] = {(cuuint64_t)ctx->Sp * sizeof(double), (cuuint64_t)ctx->Sp * ctx->Sp * sizeof(double)};
        cuuint32_t box[3] = {BK, TN, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, ctx->d_M, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) CAFE_FAIL(ctx, CAFE_GPU_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: " + std::to_string((int)r));
    }

    Params P{};
    P.ops = st.d_ops; P.n_ops = (int)ops.size(); P.n_slots = ctx->n_slots; P.F = ctx->F; P.F_pad = ctx->F_pad;
    P.counts_in_window = ctx->max_count < ctx->W ? 1 : 0;
    P.W = ctx->W; P.R = ctx->R; P.root_min = ctx->root_min; P.Sp = ctx->Sp; P.Vp = ctx->Vp; P.n_mblocks = n_mblocks;
    P.MT = ctx->d_MT; P.counts = ctx->d_counts; P.leaf_err = st.d_leaf_err; P.logprior = ctx->d_logprior;
    P.scratch = st.d_scratch; P.logpost = ctx->d_logpost; P.maxlik = ctx->d_maxlik; P.argmax = ctx->d_argmax; P.Lroot_out = d_Lroot_out;

    const size_t smem_bytes = (size_t)NSTAGE * STAGE_BYTES + sizeof(SharedCtl) + 1024;
    if (!st.attr_set) {
        CAFE_CK(ctx, cudaFuncSetAttribute(k_prune_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        st.attr_set = true;
    }
    const char* trace_path = std::getenv("CAFE_GPU_TRACE");
    long long* d_trace = nullptr;
    size_t trace_n = 0;
    if (trace_path) {
        const int pairs_max = ((n_mblocks + grid - 1) / grid + 2 * MB - 1) / (2 * MB) + 1;
        trace_n = (size_t)pairs_max * ops.size() * 2 * N_CONSUMER_WARPS * 4 + (size_t)grid * 4;
        CAFE_CK(ctx, cudaMalloc(&d_trace, trace_n * sizeof(long long)));
        CAFE_CK(ctx, cudaMemsetAsync(d_trace, 0, trace_n * sizeof(long long), ctx->stream));
        P.trace = d_trace;
        P.cta_times = d_trace + trace_n - (size_t)grid * 4;
    }
    k_prune_fused<<<grid, THREADS, smem_bytes, ctx->stream>>>(tmA, tmB, P);
    ctx->launches++;
    CAFE_CK(ctx, cudaGetLastError());
    if (trace_path) {  // debug only: synchronous dump "event warp t0 t_kloop_end t_epilogue_end t_barrier_end kind"
        std::vector<long long> h(trace_n);
        CAFE_CK(ctx, cudaMemcpyAsync(h.data(), d_trace, trace_n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CAFE_CK(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(d_trace);
        if (FILE* fp = std::fopen(trace_path, "w")) {
            for (int c = 0; c < grid; ++c) {
                const long long* t = &h[trace_n - (size_t)grid * 4 + (size_t)c * 4];
                std::fprintf(fp, "cta %d %lld %lld %lld %lld\n", c, t[0], t[1], t[2], t[3]);
            }
            for (size_t e = 0; e < (trace_n - (size_t)grid * 4) / (N_CONSUMER_WARPS * 4); ++e)
                for (int w = 0; w < N_CONSUMER_WARPS; ++w) {
                    const long long* t = &h[(e * N_CONSUMER_WARPS + w) * 4];
                    if (t[0]) std::fprintf(fp, "%zu %d %lld %lld %lld %lld %d\n", e, w, t[0], t[1], t[2], t[3], ops[(e / 2) % ops.size()].kind);
                }
            std::fclose(fp);
        }
    }
    return CAFE_GPU_OK;
}
