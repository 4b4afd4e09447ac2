// fp64 issue-rate microbenchmark for sm_100a: DFMA, DMMA (mma.sync m8n8k4 f64) and a mix.
// Prints one JSON object per variant. Used to fix the fp64 roofline denominator.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(512) k_dmma(double* out, int iters, double a, double b) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix: per loop NM dmma + NF dfma on independent accumulators
template <int NM, int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double a, double b) {
    double c0[NM], c1[NM], f[NF];
#pragma unroll
    for (int i = 0; i < NM; ++i) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (NM > NF ? NM : NF); ++i) {
            if (i < NM) dmma884(c0[i], c1[i], a, b);
            if (i < NF) f[i] = fma(f[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; ++i) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_it(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * 148 * 64 * 1024));
    const int iters = 4096;
    for (int bps = 1; bps <= 4; bps *= 2) {       // blocks (256 thr) per SM
        int grid = sms * bps;
        {
            float ms = time_it([&] { k_dfma<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 16 * iters * 256.0 * grid;
            printf("{\"variant\": \"dfma16\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", bps * 8, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&] { k_dmma<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 256 * 16 * iters * 8.0 * grid;   // 256 FMA per warp-instr
            printf("{\"variant\": \"dmma884x16\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", bps * 8, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&] { k_dmma<4><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 256 * 4 * iters * 8.0 * grid;
            printf("{\"variant\": \"dmma884x4\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", bps * 8, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&] { k_mix<8, 8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * iters * grid * (256.0 * 8 * 8 + 8 * 256.0);
            printf("{\"variant\": \"mix_8dmma_8dfma\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", bps * 8, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&] { k_mix<8, 64><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * (iters / 4) * grid * (256.0 * 8 * 8 + 64 * 256.0);
            printf("{\"variant\": \"mix_8dmma_64dfma\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", bps * 8, ms, fl / ms * 1e-9);
        }
    }
    // warps per SM sweep for DMMA with 32 independent accumulators (1 warp per SM sub-partition = 4 warps/SM)
    for (int wps : {4, 8, 12, 16}) {
        float ms = time_it([&] { k_dmma<32><<<sms, wps * 32>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = 2.0 * 256 * 32 * iters * (double)wps * sms;
        printf("{\"variant\": \"dmma884x32_wps\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", wps, ms, fl / ms * 1e-9);
    }
    // single-warp dependent-chain latency of DMMA
    {
        float ms = time_it([&] { k_dmma<1><<<1, 32>>>(out, 1 << 16, 1.0000001, 1e-9); }, 3);
        printf("{\"variant\": \"dmma_latency_chain\", \"ns_per_dmma\": %.3f}\n", ms * 1e6 / (1 << 16));
        ms = time_it([&] { k_dfma<1><<<1, 32>>>(out, 1 << 16, 1.0000001, 1e-9); }, 3);
        printf("{\"variant\": \"dfma_latency_chain\", \"ns_per_dfma\": %.3f}\n", ms * 1e6 / (1 << 16));
    }
    return 0;
}
