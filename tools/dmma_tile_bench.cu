// dmma_tile_bench.cu — how fast can DMMA.8x8x4 run when its operands come from shared memory the way K2 feeds them?
// One CTA per SM, WPS warps per SM sub-partition, each warp owns an (8*MBV) x 32 tile and loops over "stages" of 4 k4-steps
// with register double buffering of the fragments (the structure of gemm_kblocks in cafe_b200/csrc/prune_fused2.cu), optionally
// with an mbarrier wait + arrive per stage.  Prints one JSON line per variant.  Build: nvcc -arch=sm_100a -O3 -o dmma_tile_bench dmma_tile_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ int perm(int g) { return 2 * (g & 3) + (g >> 2); }

constexpr int NB = 4;

template <int MBV, bool CHERRY = false>
__device__ __forceinline__ void load_frags(double (&fa)[MBV], double (&fb)[NB], const unsigned char* sA, const unsigned char* sB, int off) {
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) fb[nb] = *reinterpret_cast<const double*>(sB + nb * 1024 + off);
#pragma unroll
    for (int mb = 0; mb < MBV; ++mb) {
        fa[mb] = *reinterpret_cast<const double*>(sA + mb * 1024 + off);
        if (CHERRY) fa[mb] = __dmul_rn(fa[mb], *reinterpret_cast<const double*>(sA + 8192 + mb * 1024 + off));
    }
}

// MODE 0: registers only (no shared-memory loads in the loop); 1: shared-memory fragments; 2: + mbarrier wait/arrive per stage
// (each warp has its own barrier, arrives itself: cost of the instructions, not of waiting for a producer)
template <int MBV, int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_tile(double* out, int stages) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3, pg = perm(g);
    // 3 "stages" of A (96 rows x 128 B) and B (128 rows x 128 B)
    for (int i = threadIdx.x; i < 3 * (96 + 128) * 16; i += blockDim.x) reinterpret_cast<double*>(smem)[i] = 1e-3 * (i % 97);
    if (threadIdx.x < 16) mbar_init(&bars[threadIdx.x], 1);
    __syncthreads();
    double acc[MBV][NB][2];
#pragma unroll
    for (int mb = 0; mb < MBV; ++mb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    const int off0 = pg * 128 + ((q & 1) << 3), hi = q >> 1;
    const int a_off = ((warp >> 2) * MBV * 8 % 96) * 128, b_off = 96 * 128 + (warp & 3) * 32 * 128;
    double fa[2][MBV], fb[2][NB];
    int st = 0;
    uint32_t phase = 0;
    load_frags<MBV, MODE == 3>(fa[0], fb[0], smem + a_off, smem + b_off, off0 + ((hi ^ pg) << 4));
    for (int s = 0; s < stages; ++s) {
        const unsigned char* base = smem + st * (224 * 128);
        int nst = st + 1 == 3 ? 0 : st + 1;
        const unsigned char* nbase = smem + nst * (224 * 128);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            if (MODE >= 1) {
                if (kk < 3) load_frags<MBV, MODE == 3>(fa[(kk + 1) & 1], fb[(kk + 1) & 1], base + a_off, base + b_off, off0 + (((2 * (kk + 1) + hi) ^ pg) << 4));
                else {
                    if (MODE == 2 || MODE == 3) { if (lane == 0) mbar_arrive(&bars[warp]); mbar_wait(&bars[warp], phase); phase ^= 1; }
                    load_frags<MBV, MODE == 3>(fa[0], fb[0], nbase + a_off, nbase + b_off, off0 + ((hi ^ pg) << 4));
                }
            }
#pragma unroll
            for (int mb = 0; mb < MBV; ++mb)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], fa[MODE >= 1 ? (kk & 1) : 0][mb], fb[MODE >= 1 ? (kk & 1) : 0][nb]);
        }
        st = nst;
    }
    double sum = 0;
#pragma unroll
    for (int mb = 0; mb < MBV; ++mb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) sum += acc[mb][nb][0] + acc[mb][nb][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

template <int MBV, int MODE, int MAXT = 512>
static void run(int sms, int wps, double* out) {
    const int stages = 4096;
    const size_t smem = 3 * 224 * 128;
    CK(cudaFuncSetAttribute(k_tile<MBV, MODE, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        CK(cudaEventRecord(e0));
        k_tile<MBV, MODE, MAXT><<<sms, wps * 4 * 32, smem>>>(out, stages);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double fl = 2.0 * 256 * (MBV * NB * 4.0) * stages * (wps * 4.0) * sms;
    printf("{\"variant\": \"tile\", \"mbv\": %d, \"mode\": %d, \"warps_per_smsp\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", MBV, MODE, wps, best, fl / best * 1e-9);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, sms);
    double* out; CK(cudaMalloc(&out, sizeof(double) * 148 * 512));
    // 256 threads max => up to 255 registers per thread: the 48x32 warp tile (MBV 6) of prune_fused2.cu without spills
    for (int wps = 1; wps <= 2; ++wps) {
        run<4, 1, 256>(sms, wps, out); run<4, 2, 256>(sms, wps, out); run<4, 3, 256>(sms, wps, out);
        run<5, 2, 256>(sms, wps, out);
        run<6, 0, 256>(sms, wps, out); run<6, 1, 256>(sms, wps, out); run<6, 2, 256>(sms, wps, out); run<6, 3, 256>(sms, wps, out);
        run<8, 1, 256>(sms, wps, out); run<8, 2, 256>(sms, wps, out);
    }
    for (int wps = 3; wps <= 4; ++wps) { run<4, 2>(sms, wps, out); run<4, 3>(sms, wps, out); run<3, 2>(sms, wps, out); run<3, 3>(sms, wps, out); }
    return 0;
}
