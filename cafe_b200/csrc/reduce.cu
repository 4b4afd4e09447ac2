// reduce.cu — K3: score reduction  sum_f mult_f * log(max posterior_f)  and the first zero-likelihood
// family (get_posterior, cafe/lambda.cpp:691-724).  Deterministic: one CTA, fixed assignment of
// families to threads, fixed-shape tree reduction — the same inputs always give the same bits.
// Output is device resident (out[0] = partial score, out[1] = min first_index of a zero family as a
// double, +inf if none) so that a multi-GPU caller can all-reduce it without a host round trip.
#include "common.cuh"

namespace {
constexpr int RED_THREADS = 1024;

__global__ void __launch_bounds__(RED_THREADS)
k_score_reduce(const double* __restrict__ logpost, const double* __restrict__ maxlik, const int* __restrict__ mult,
               const int* __restrict__ first, int F, double* __restrict__ out) {
    __shared__ double s_sum[RED_THREADS];
    __shared__ double s_min[RED_THREADS];
    double sum = 0.0, mn = INFINITY;
    for (int f = threadIdx.x; f < F; f += RED_THREADS) {
        if (maxlik[f] == 0.0) {           // lambda.cpp:715 — the family that makes the reference throw
            mn = fmin(mn, (double)first[f]);
        } else {
            sum += (double)mult[f] * logpost[f];
        }
    }
    s_sum[threadIdx.x] = sum;
    s_min[threadIdx.x] = mn;
    __syncthreads();
    for (int off = RED_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            s_sum[threadIdx.x] += s_sum[threadIdx.x + off];
            s_min[threadIdx.x] = fmin(s_min[threadIdx.x], s_min[threadIdx.x + off]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = s_sum[0]; out[1] = s_min[0]; }
}
}  // namespace

int launch_score_reduce(cafe_gpu_ctx* ctx, double* d_out2) {
    k_score_reduce<<<1, RED_THREADS, 0, ctx->stream>>>(ctx->d_logpost, ctx->d_maxlik, ctx->d_mult, ctx->d_first,
                                                       ctx->F, d_out2);
    ctx->launches++;
    CAFE_CK(ctx, cudaGetLastError());
    return CAFE_GPU_OK;
}
