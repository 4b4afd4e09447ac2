// comm.cu — multi-GPU plumbing of the likelihood path, behind the C-ABI.
//
// Families shard naturally (get_posterior has no cross-family state except the running sum and the first zero family,
// cafe/lambda.cpp:698-722), and the D distinct transition matrices are the same on every rank.  One objective evaluation
// with a communicator of `world` ranks is therefore
//     K1 on this rank's ceil(D / world) keys                                   (reset_birthdeath_cache, cafe_main.c:319)
//  -> ncclAllGather of d_M in place + local transposes of the received keys    (exchange 1: world-1 / world of D*Sp*Sp*8 bytes in)
//  -> K2 + K3 on this rank's families                                          (get_posterior, lambda.cpp:691-724)
//  -> ncclAllGather of {partial score, first zero family} (16 B per rank) and a sum in rank order on every rank
//                                                                              (exchange 2: the "allreduce of the scalar")
// all on the context's stream, no host round trip in between.  The gather-then-ordered-sum gives every rank the same bits and
// carries the sum and the min in one collective (ncclAllReduce would need two, and its summation order is NCCL's).
//
// Two ways to get a communicator, same code path afterwards:
//   * one process per GPU (torchrun, MPI, ...): rank 0 calls cafe_gpu_comm_unique_id, the launcher broadcasts the 128 bytes,
//     every rank calls cafe_gpu_comm_init(ctx, id, rank, world);
//   * one process, several devices (the C++ host library, CAFE_GPUS=...): cafe_gpu_create_multi makes one context per device
//     (ncclCommInitAll) and returns the leader; collectives of the local ranks are issued inside ncclGroupStart/End.
//
// libnccl.so.2 is opened lazily with dlopen: a process that never asks for a communicator never loads it, and a process that
// already has it (torch's bundled copy) shares that copy.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return; }
        auto sym = [&](const char* name) -> void* {
            void* p = dlsym(api.handle, name);
            if (!p && api.error.empty()) api.error = std::string("libnccl: missing symbol ") + name;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    return api;
}

#define CAFE_NCCL(ctx, expr)                                                                                          \
    do {                                                                                                              \
        ncclResult_t r__ = (expr);                                                                                    \
        if (r__ != ncclSuccess) {                                                                                     \
            (ctx)->err = std::string(#expr) + ": " + nccl().GetErrorString(r__);                                      \
            return CAFE_GPU_ERR_CUDA;                                                                                 \
        }                                                                                                             \
    } while (0)

// out[0] = sum of the partial scores in rank order, out[1] = min of the first zero-family indices (+inf if none)
__global__ void k_finish_score(const double* __restrict__ all, int world, double* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0, z = INFINITY;
    for (int r = 0; r < world; ++r) { s += all[2 * r]; z = fmin(z, all[2 * r + 1]); }
    out[0] = s; out[1] = z;
}

int attach(cafe_gpu_ctx* ctx, ncclComm_t comm, int rank, int world) {
    ctx->nccl_comm = comm; ctx->comm_rank = rank; ctx->comm_world = world;
    ctx->shard_rank = rank; ctx->shard_world = world;
    ctx->matrices_valid = false; ctx->results_valid = false;
    cudaFree(ctx->d_score_all); ctx->d_score_all = nullptr;
    CAFE_CK(ctx, cudaMalloc(&ctx->d_score_all, (size_t)2 * world * sizeof(double)));
    return CAFE_GPU_OK;
}

}  // namespace

void comm_release(cafe_gpu_ctx* ctx) {
    if (ctx->nccl_comm && nccl().CommDestroy) nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr; ctx->comm_rank = 0; ctx->comm_world = 1;
}

// Exchange 1.  Every local context has built the keys [key_lo, key_hi) of its rank into chunk `rank` of d_M (chunks of
// keys_per_rank matrices, cafe_gpu_build_matrices); the in-place all-gather fills the other chunks, a local transpose kernel
// then writes their MT copies (half the NVLink bytes of gathering M and MT).
int comm_exchange_matrices(std::vector<cafe_gpu_ctx*>& L) {
    NcclApi& N = nccl();
    cafe_gpu_ctx* lead = L[0];
    if (!N.error.empty()) CAFE_FAIL(lead, CAFE_GPU_ERR_UNSUPPORTED, N.error);
    for (cafe_gpu_ctx* c : L) {
        if (!c->nccl_comm) CAFE_FAIL(lead, CAFE_GPU_ERR_STATE, "exchange: context has no communicator");
        cudaSetDevice(c->device);
        if (c->timing) CAFE_CK(c, cudaEventRecord(c->evt(c->ring_x, EV_XCHG_BEGIN), c->stream));
    }
    if (L.size() > 1) CAFE_NCCL(lead, N.GroupStart());
    for (cafe_gpu_ctx* c : L) {
        const size_t chunk = (size_t)c->keys_per_rank * c->Sp * c->Sp;
        cudaSetDevice(c->device);
        CAFE_NCCL(lead, N.AllGather(c->d_M + (size_t)c->comm_rank * chunk, c->d_M, chunk, ncclDouble, (ncclComm_t)c->nccl_comm, c->stream));
        c->launches++;
    }
    if (L.size() > 1) CAFE_NCCL(lead, N.GroupEnd());
    for (cafe_gpu_ctx* c : L) {
        cudaSetDevice(c->device);
        int rc = launch_transpose_keys(c, 0, c->key_lo, c->key_hi, (int)c->keys.size());
        if (rc) { lead->err = c->err; return rc; }
        if (c->timing) { CAFE_CK(c, cudaEventRecord(c->evt(c->ring_x, EV_XCHG_END), c->stream)); c->ring_x++; }
        c->matrices_need_exchange = false;
        c->matrices_valid = true;
    }
    return CAFE_GPU_OK;
}

// Exchange 2.  d_score of every rank -> d_score_all on every rank -> d_score_final = {ordered sum, min}.
int comm_reduce_scores(std::vector<cafe_gpu_ctx*>& L) {
    NcclApi& N = nccl();
    cafe_gpu_ctx* lead = L[0];
    if (!N.error.empty()) CAFE_FAIL(lead, CAFE_GPU_ERR_UNSUPPORTED, N.error);
    for (cafe_gpu_ctx* c : L) {
        cudaSetDevice(c->device);
        if (c->timing) CAFE_CK(c, cudaEventRecord(c->evt(c->ring_r, EV_RED_BEGIN), c->stream));
    }
    if (L.size() > 1) CAFE_NCCL(lead, N.GroupStart());
    for (cafe_gpu_ctx* c : L) {
        cudaSetDevice(c->device);
        CAFE_NCCL(lead, N.AllGather(c->d_score, c->d_score_all, 2, ncclDouble, (ncclComm_t)c->nccl_comm, c->stream));
        c->launches++;
    }
    if (L.size() > 1) CAFE_NCCL(lead, N.GroupEnd());
    for (cafe_gpu_ctx* c : L) {
        cudaSetDevice(c->device);
        k_finish_score<<<1, 32, 0, c->stream>>>(c->d_score_all, c->comm_world, c->d_score_final);
        c->launches++;
        CAFE_CK(c, cudaGetLastError());
        if (c->timing) { CAFE_CK(c, cudaEventRecord(c->evt(c->ring_r, EV_RED_END), c->stream)); c->ring_r++; }
    }
    return CAFE_GPU_OK;
}

extern "C" {

int cafe_gpu_comm_unique_id(void* id_out, int id_bytes) {
    if (!id_out || id_bytes < (int)sizeof(ncclUniqueId)) return CAFE_GPU_ERR_ARG;
    NcclApi& N = nccl();
    if (!N.error.empty()) return CAFE_GPU_ERR_UNSUPPORTED;
    ncclUniqueId id;
    if (N.GetUniqueId(&id) != ncclSuccess) return CAFE_GPU_ERR_CUDA;
    std::memset(id_out, 0, id_bytes);
    std::memcpy(id_out, &id, sizeof(id));
    return CAFE_GPU_OK;
}

int cafe_gpu_comm_init(cafe_gpu_ctx* ctx, const void* id, int id_bytes, int rank, int world) {
    if (!ctx || !id || id_bytes < (int)sizeof(ncclUniqueId)) return CAFE_GPU_ERR_ARG;
    if (world < 1 || rank < 0 || rank >= world) CAFE_FAIL(ctx, CAFE_GPU_ERR_ARG, "comm_init: need 0 <= rank < world");
    if (ctx->leader || !ctx->peers.empty()) CAFE_FAIL(ctx, CAFE_GPU_ERR_STATE, "comm_init: context belongs to cafe_gpu_create_multi");
    NcclApi& N = nccl();
    if (!N.error.empty()) CAFE_FAIL(ctx, CAFE_GPU_ERR_UNSUPPORTED, N.error);
    cudaSetDevice(ctx->device);
    comm_release(ctx);
    if (world == 1) return attach(ctx, nullptr, 0, 1);
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    CAFE_NCCL(ctx, N.CommInitRank(&comm, world, uid, rank));
    return attach(ctx, comm, rank, world);
}

int cafe_gpu_comm_size(const cafe_gpu_ctx* ctx) { return ctx ? ctx->comm_world : 0; }
int cafe_gpu_comm_rank(const cafe_gpu_ctx* ctx) { return ctx ? ctx->comm_rank : -1; }

int cafe_gpu_create_multi(cafe_gpu_ctx** out, const int* devices, int n_devices) {
    if (!out || n_devices < 1) return CAFE_GPU_ERR_ARG;
    *out = nullptr;
    std::vector<int> dev(n_devices);
    for (int i = 0; i < n_devices; ++i) dev[i] = devices ? devices[i] : i;
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (dev[i] == dev[j]) return CAFE_GPU_ERR_ARG;  // NCCL needs distinct devices within one communicator
    std::vector<cafe_gpu_ctx*> L(n_devices, nullptr);
    auto fail = [&](int rc) { for (cafe_gpu_ctx* c : L) if (c) { c->peers.clear(); cafe_gpu_destroy(c); } return rc; };
    for (int i = 0; i < n_devices; ++i) {
        int rc = cafe_gpu_create(&L[i], dev[i]);
        if (rc) return fail(rc);
    }
    if (n_devices > 1) {
        NcclApi& N = nccl();
        if (!N.error.empty()) return fail(CAFE_GPU_ERR_UNSUPPORTED);
        std::vector<ncclComm_t> comms(n_devices);
        if (N.CommInitAll(comms.data(), n_devices, dev.data()) != ncclSuccess) return fail(CAFE_GPU_ERR_CUDA);
        for (int i = 0; i < n_devices; ++i) {
            cudaSetDevice(dev[i]);
            int rc = attach(L[i], comms[i], i, n_devices);
            if (rc) return fail(rc);
            if (i > 0) { L[i]->leader = L[0]; L[0]->peers.push_back(L[i]); }
        }
    }
    cudaSetDevice(dev[0]);
    *out = L[0];
    return CAFE_GPU_OK;
}

int cafe_gpu_num_devices(const cafe_gpu_ctx* ctx) { return ctx ? (int)ctx->peers.size() + 1 : 0; }

}  // extern "C"
