// ozaki_tile_bench.cu — split-precision (Ozaki) fp64 GEMM on the 5th-generation tensor cores, as a measured prototype.
//
// BASELINE.json's north star names "tcgen05 tensor cores ... fp32/split where tolerance allows".  tcgen05.mma has no fp64 kind,
// so an fp64 result from these tensor cores means slicing both operands into 8-bit integers, one exact `kind::i8` MMA per pair
// of slices with int32 accumulators in TMEM, and an fp64 recombination.  This tool is ONE edge GEMM of K2 in that form,
//      Out[f][i] = sum_j B[i][j] * A[f][j]      (A: node vectors, families x sizes; B: transition matrix, cafe_tree.c:226-271 /
//                                                birthdeath.c:163-182)
// with the operands of the real workload (tools/ozaki_tile_bench.py feeds the leaf-pair vectors and a transition matrix of
// BASELINE configs[1]), so that throughput and error are measured, not estimated (profiles/r2_ozaki_study.md has the estimate).
// It is NOT on the product path (the DMMA kernel of csrc/prune_fused2.cu is); nothing under cafe_b200/ links it.
//
// Scheme.  All entries are >= 0 (probabilities), so slices are UNSIGNED 8-bit digits: per row, x = v * 2^-e with e the exponent
// of the row maximum, x = sum_s d_s 256^-(s+1), S digits kept.  Slice products with s + t < S are accumulated by weight class
// w = s + t in S int32 accumulators (exact: at most S * 256 * 255^2 < 2^27 each), recombined in two int64 halves and two
// int64 -> fp64 conversions per output, scaled by 2^(eA[f] + eB[i]).
//
// Kernel (one CTA per 128-family x 64-size output tile, K = 256 sizes = 2 chunks of 128 B):
//   warp 4 lane 0  TMA producer: all S slices of the B tile (S x 2 x 8 KB) once, A slices (16 KB = 128 rows x 128 B, SWIZZLE_128B)
//                  through a ring, order (chunk, slice s)
//   warp 5 lane 0  tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 64, K = 32 per instruction: for every A stage the slices
//                  t <= S-1-s of B, four k-steps each, accumulator w = s + t at TMEM columns [64 w, 64 w + 64); tcgen05.commit
//                  frees the stage / signals the epilogue
//   warps 0..3     epilogue: tcgen05.ld (32x32b) of the S accumulators, integer recombination, fp64 scaling, store
// TMEM: S x 64 <= 512 columns, which is what limits the tile to N = 64 for S = 7, 8.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC -o tools/libozaki_tile.so tools/ozaki_tile_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {

constexpr int TM = 128, TN = 64, KTOT = 256, KCH = 128;      // tile, total K, K bytes per chunk (one 128B swizzle row)
constexpr int NCH = KTOT / KCH;
constexpr int A_STAGE_BYTES = TM * KCH;                       // 16 KB
constexpr int B_SLICE_BYTES = TN * KCH;                       // 8 KB per (slice, chunk)
constexpr int NSTAGE = 5;
constexpr int MAXS = 8;
constexpr int THREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
// tcgen05.commit: the mbarrier receives one arrival once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 B, groups of 8 rows 1024 B apart (cute/arch/mma_sm100_desc.hpp
// SmemDescriptor: start address >> 4 in [0,14), leading byte offset [16,30) (unused for swizzled K-major), stride byte offset
// [32,46), version 1 at [46,48), layout type 2 = SWIZZLE_128B at [61,64)).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (same header, InstrDescriptor): c_format S32 = 2 at [4,6), a/b format UINT8 = 0 at [7,10) / [10,13),
// both K-major, N >> 3 at [17,23), M >> 4 at [24,29).
constexpr uint32_t IDESC_U8_128x64 = (2u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}

struct Ctl {
    uint64_t full_a[NSTAGE], empty_a[NSTAGE], full_b, acc_done;
    uint32_t tmem_base;
};

// mode 0: everything; 1: MMAs and feed only (epilogue waits, loads nothing, stores nothing); 2: no MMAs (feed + epilogue)
template <int S>
__global__ void __launch_bounds__(THREADS, 1)
k_ozaki_tile(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const int* __restrict__ eA,
             const int* __restrict__ eB, double* __restrict__ out, int N, int mode) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sB = smem;                                   // [chunk][slice][64 rows x 128 B]
    unsigned char* sA = smem + NCH * S * B_SLICE_BYTES;         // [stage][128 rows x 128 B]
    Ctl* ctl = reinterpret_cast<Ctl*>(sA + NSTAGE * A_STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f0 = blockIdx.x * TM, n0 = blockIdx.y * TN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&ctl->full_a[s], 1); mbar_init(&ctl->empty_a[s], 1); }
        mbar_init(&ctl->full_b, 1); mbar_init(&ctl->acc_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {  // TMEM: all 512 columns (one CTA per SM: the shared memory request guarantees it)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = ctl->tmem_base;

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(&ctl->full_b, NCH * S * B_SLICE_BYTES);
            for (int c = 0; c < NCH; ++c)
                for (int t = 0; t < S; ++t) tma_load_3d(sB + (c * S + t) * B_SLICE_BYTES, &tmB, c * KCH, n0, t, &ctl->full_b);
            uint32_t stage = 0, phase = 0;
            for (int c = 0; c < NCH; ++c)
                for (int s = 0; s < S; ++s) {
                    mbar_wait(&ctl->empty_a[stage], phase ^ 1);
                    mbar_expect_tx(&ctl->full_a[stage], A_STAGE_BYTES);
                    tma_load_3d(sA + stage * A_STAGE_BYTES, &tmA, c * KCH, f0, s, &ctl->full_a[stage]);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            mbar_wait(&ctl->full_b, 0);
            uint32_t stage = 0, phase = 0;
            for (int c = 0; c < NCH; ++c)
                for (int s = 0; s < S; ++s) {
                    mbar_wait(&ctl->full_a[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (mode != 2) {
                        const uint64_t da = umma_desc_sw128(smem_u32(sA + stage * A_STAGE_BYTES));
                        for (int t = 0; t + s < S; ++t) {
                            const uint64_t db = umma_desc_sw128(smem_u32(sB + (c * S + t) * B_SLICE_BYTES));
#pragma unroll
                            for (int k = 0; k < KCH / 32; ++k)  // 32 bytes further along K inside the swizzled row: start address + 2
                                umma_i8(tmem + (uint32_t)(s + t) * TN, da + 2 * k, db + 2 * k, IDESC_U8_128x64, (c | s | k) != 0);
                        }
                    }
                    umma_commit(&ctl->empty_a[stage]);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            umma_commit(&ctl->acc_done);
        }
    } else {
        // epilogue: thread = TMEM lane = family row of the tile
        mbar_wait(&ctl->acc_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (mode != 1) {
            const int row = warp * 32 + lane;
            const int ea = eA[f0 + row];
            double* orow = out + (size_t)(f0 + row) * N + n0;
            const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
            for (int j0 = 0; j0 < TN; j0 += 8) {
                uint32_t acc[S][8];
#pragma unroll
                for (int w = 0; w < S; ++w) tmem_ld8(tlane + (uint32_t)(w * TN + j0), acc[w]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // value = sum_w acc_w 256^-(w+2): classes 0..3 and 4..S-1 in two exact int64 sums
                    long long hi = 0, lo = 0;
#pragma unroll
                    for (int w = 0; w < S; ++w) {
                        if (w < 4) hi += (long long)acc[w][j] << (8 * (3 - w));
                        else lo += (long long)acc[w][j] << (8 * (S - 1 - w));
                    }
                    double v = (double)hi * 0x1p-40 + (double)lo * (1.0 / (double)(1ull << (8 * (S + 1) - 32)) * 0x1p-32);
                    orow[j0 + j] = scalbn(v, ea + __ldg(eB + n0 + j0 + j));
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// Slicing: one warp per row of 256 doubles.  e = exponent of the row maximum (row = 2^e * x, x in [0,1)), S base-256 digits of x.
template <int S>
__global__ void k_slice(const double* __restrict__ X, int rows, int rows_pad, unsigned char* __restrict__ planes, int* __restrict__ e_out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows_pad) return;
    double v[8];
    double mx = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[i] = row < rows ? X[(size_t)row * KTOT + lane * 8 + i] : 0.0;
        mx = fmax(mx, v[i]);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);  // mx = m * 2^e, m in [0.5, 1)
    if (lane == 0) e_out[row] = e;
    unsigned long long packed[S];
#pragma unroll
    for (int s = 0; s < S; ++s) packed[s] = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double x = scalbn(v[i], -e);  // exact, in [0, 1)
#pragma unroll
        for (int s = 0; s < S; ++s) {
            x *= 256.0;
            const double d = floor(x);
            x -= d;
            packed[s] |= (unsigned long long)(unsigned int)d << (8 * i);
        }
    }
#pragma unroll
    for (int s = 0; s < S; ++s)
        *reinterpret_cast<unsigned long long*>(planes + ((size_t)s * rows_pad + row) * KTOT + lane * 8) = packed[s];
}

typedef CUresult (*PFN_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <int S>
int run(const double* A, const double* B, int F, int N, double* out, float* ms, int iters, int mode) {
    const int Fp = (F + TM - 1) / TM * TM, Np = (N + TN - 1) / TN * TN;
    double *dA, *dB, *dOut;
    unsigned char *pA, *pB;
    int *eA, *eB;
    CK(cudaMalloc(&dA, (size_t)F * KTOT * 8)); CK(cudaMalloc(&dB, (size_t)N * KTOT * 8)); CK(cudaMalloc(&dOut, (size_t)Fp * Np * 8));
    CK(cudaMalloc(&pA, (size_t)S * Fp * KTOT)); CK(cudaMalloc(&pB, (size_t)S * Np * KTOT));
    CK(cudaMalloc(&eA, Fp * 4)); CK(cudaMalloc(&eB, Np * 4));
    CK(cudaMemcpy(dA, A, (size_t)F * KTOT * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B, (size_t)N * KTOT * 8, cudaMemcpyHostToDevice));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    PFN_encode encode = (PFN_encode)fn;
    CUtensorMap tmA, tmB;
    auto mk = [&](CUtensorMap* tm, void* base, int rows, int box_rows) {
        cuuint64_t dims[3] = {(cuuint64_t)KTOT, (cuuint64_t)rows, (cuuint64_t)S};
        cuuint64_t strides[2] = {(cuuint64_t)KTOT, (cuuint64_t)KTOT * rows};
        cuuint32_t box[3] = {KCH, (cuuint32_t)box_rows, 1};
        cuuint32_t es[3] = {1, 1, 1};
        return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    if (mk(&tmA, pA, Fp, TM) != CUDA_SUCCESS || mk(&tmB, pB, Np, TN) != CUDA_SUCCESS) { std::fprintf(stderr, "tensor map failed\n"); return 1; }
    const size_t smem = (size_t)NCH * S * B_SLICE_BYTES + (size_t)NSTAGE * A_STAGE_BYTES + sizeof(Ctl) + 1024;
    CK(cudaFuncSetAttribute(k_ozaki_tile<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    float t_slice = 0, t_mma = 0;
    for (int it = 0; it < iters + 2; ++it) {
        CK(cudaEventRecord(e0));
        k_slice<S><<<(Fp + 7) / 8, 256>>>(dA, F, Fp, pA, eA);
        k_slice<S><<<(Np + 7) / 8, 256>>>(dB, N, Np, pB, eB);
        CK(cudaEventRecord(e1));
        k_ozaki_tile<S><<<dim3(Fp / TM, Np / TN), THREADS, smem>>>(tmA, tmB, eA, eB, dOut, Np, mode);
        CK(cudaEventRecord(e2));
        CK(cudaEventSynchronize(e2));
        CK(cudaGetLastError());
        float a, b;
        CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e1, e2));
        if (it >= 2) { t_slice += a; t_mma += b; }
    }
    ms[0] = t_slice / iters; ms[1] = t_mma / iters;
    CK(cudaMemcpy2D(out, (size_t)N * 8, dOut, (size_t)Np * 8, (size_t)N * 8, F, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dOut); cudaFree(pA); cudaFree(pB); cudaFree(eA); cudaFree(eB);
    return 0;
}

}  // namespace

// A: F x 256 fp64 (row-major, entries >= 0), B: N x 256 fp64 (row i = output size i), out: F x N fp64.
// ms[0] = slicing of both operands, ms[1] = the tcgen05 kernel (mean over `iters` launches after two warm-ups).
extern "C" int ozaki_tile_gemm(const double* A, const double* B, int F, int N, int S, double* out, float* ms, int iters, int mode) {
    if (S == 7) return run<7>(A, B, F, N, out, ms, iters, mode);
    if (S == 8) return run<8>(A, B, F, N, out, ms, iters, mode);
    if (S == 6) return run<6>(A, B, F, N, out, ms, iters, mode);
    if (S == 4) return run<4>(A, B, F, N, out, ms, iters, mode);
    return 2;
}
