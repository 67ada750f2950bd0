// comm.cu — the one exchange step of the data-parallel nn.rs training step: a sum all-reduce of the
// weight / bias gradients across the GPUs of one box, NCCL over NVLink 5 / NVSwitch.
//
// libnccl is dlopen'ed at first use (the torch-bundled libnccl.so.2 when the process already loaded it, else the
// system one), so that libsliced_b200.so itself loads on a box without NCCL or without a GPU.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess = 0 };
enum { ncclInt32 = 2, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    if (api.handle || api.ok) return api;
    const char* names[] = {getenv("SLICED_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return api;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce;
    return api;
}

}  // namespace

#define SL_NCCL(ctx, expr)                                                                                   \
    do {                                                                                                     \
        ncclResult_t r__ = (expr);                                                                           \
        if (r__ != ncclSuccess)                                                                              \
            return sl_set_error((ctx), SL_ERR_NCCL, "%s: %s -> %s", __func__, #expr,                         \
                                nccl().GetErrorString ? nccl().GetErrorString(r__) : "nccl error");          \
    } while (0)

extern "C" {

int sl_comm_unique_id(void* id_out_128) {
    if (!id_out_128) return sl_set_error(nullptr, SL_ERR_INVALID_ARG, "sl_comm_unique_id: NULL");
    if (!nccl().ok) return sl_set_error(nullptr, SL_ERR_NCCL, "sl_comm_unique_id: libnccl.so.2 not found");
    ncclUniqueId id;
    SL_NCCL(nullptr, nccl().GetUniqueId(&id));
    memcpy(id_out_128, &id, sizeof(id));
    return SL_OK;
}

int sl_comm_init_rank(sl_ctx* ctx, int nranks, int rank, const void* id_128) {
    SL_REQUIRE(ctx, ctx && id_128, "NULL argument");
    SL_REQUIRE(ctx, nranks >= 1 && rank >= 0 && rank < nranks, "bad rank");
    if (!nccl().ok) return sl_set_error(ctx, SL_ERR_NCCL, "sl_comm_init_rank: libnccl.so.2 not found");
    if (ctx->nccl_comm) return sl_set_error(ctx, SL_ERR_INVALID_ARG, "sl_comm_init_rank: this context already has a communicator (sl_comm_destroy first)");
    SL_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_128, sizeof(id));
    ncclComm_t comm = nullptr;
    SL_NCCL(ctx, nccl().CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    return SL_OK;
}

static int nccl_dtype(int dtype) { return dtype == SL_F32 ? ncclFloat32 : (dtype == SL_F64 ? ncclFloat64 : (dtype == SL_I32 ? ncclInt32 : -1)); }

int sl_allreduce_sum(sl_ctx* ctx, int dtype, void* buf, size_t n) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, buf);
    SL_REQUIRE(ctx, nccl_dtype(dtype) >= 0, "bad dtype");
    if (n == 0) return SL_OK;
    SL_REQUIRE(ctx, buf != nullptr, "NULL buffer");
    if (!ctx->nccl_comm) {
        if (ctx->nranks <= 1) return SL_OK;  // a world of one: the sum over ranks is the buffer itself
        return sl_set_error(ctx, SL_ERR_NCCL, "sl_allreduce_sum: communicator not initialised");
    }
    if (getenv("SLICED_DP_NOCOMM")) return SL_OK;   // diagnosis only (tools/dp_sweep.py): the step without its exchange
    const int dt = nccl_dtype(dtype);
    SL_NCCL(ctx, nccl().AllReduce(buf, buf, n, dt, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return SL_OK;
}

// Overlapped exchange: the all-reduce runs on the context's communication stream, ordered after everything issued so far on
// the compute stream (so the gradient segment is complete), while the compute stream carries on with the rest of backward.
int sl_allreduce_sum_async(sl_ctx* ctx, int dtype, void* buf, size_t n) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, buf);
    SL_REQUIRE(ctx, nccl_dtype(dtype) >= 0, "bad dtype");
    if (n == 0) return SL_OK;
    SL_REQUIRE(ctx, buf != nullptr, "NULL buffer");
    if (!ctx->nccl_comm) {
        if (ctx->nranks <= 1) return SL_OK;
        return sl_set_error(ctx, SL_ERR_NCCL, "sl_allreduce_sum_async: communicator not initialised");
    }
    if (getenv("SLICED_DP_NOCOMM")) return SL_OK;
    if (!ctx->comm_stream) {
        SL_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
        SL_CUDA(ctx, cudaEventCreateWithFlags(&ctx->comm_ready, cudaEventDisableTiming));
        SL_CUDA(ctx, cudaEventCreateWithFlags(&ctx->comm_done, cudaEventDisableTiming));
    }
    SL_CUDA(ctx, cudaEventRecord(ctx->comm_ready, ctx->stream));
    SL_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_ready, 0));
    const int dt = nccl_dtype(dtype);
    SL_NCCL(ctx, nccl().AllReduce(buf, buf, n, dt, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->comm_stream));
    ctx->comm_pending = true;
    if (ctx->comm_issued >= ctx->comm_events.size()) {
        cudaEvent_t e = nullptr;
        SL_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->comm_events.push_back(e);
    }
    SL_CUDA(ctx, cudaEventRecord(ctx->comm_events[ctx->comm_issued++], ctx->comm_stream));
    return SL_OK;
}

int sl_comm_nranks(sl_ctx* ctx) { return ctx && ctx->nccl_comm ? ctx->nranks : 1; }

// Number of overlapped exchanges issued since the last sl_comm_wait.
int sl_comm_issued(sl_ctx* ctx) { return ctx ? (int)ctx->comm_issued : 0; }

// The compute stream waits for the FIRST n exchanges issued since the last sl_comm_wait (they complete in issue order), so work
// that only needs those — the SGD update of a layer whose gradients have arrived — can run while later exchanges are in flight.
int sl_comm_wait_n(sl_ctx* ctx, int n) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    if (n <= 0 || !ctx->comm_stream || ctx->comm_issued == 0) return SL_OK;
    const size_t i = (size_t)n < ctx->comm_issued ? (size_t)n : ctx->comm_issued;
    sl_ctx::ProfRec pr{};
    const bool prof = ctx->profiling && ctx->profiling_all;
    if (prof) {   // in-situ breakdown: time the compute stream spends idle waiting for these exchanges
        cudaEventCreate(&pr.a);
        cudaEventCreate(&pr.b);
        pr.name = "(exposed gradient exchange: compute stream waiting in sl_comm_wait_n)";
        cudaEventRecord(pr.a, ctx->stream);
    }
    SL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->comm_events[i - 1], 0));
    if (prof) {
        cudaEventRecord(pr.b, ctx->stream);
        ctx->prof.push_back(pr);
    }
    return SL_OK;
}

// The compute stream waits for every exchange issued with sl_allreduce_sum_async.
int sl_comm_wait(sl_ctx* ctx) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    ctx->comm_issued = 0;
    if (!ctx->comm_stream || !ctx->comm_pending) return SL_OK;
    SL_CUDA(ctx, cudaEventRecord(ctx->comm_done, ctx->comm_stream));
    // in-situ breakdown (sl_ctx_profile_report): how long the compute stream sits idle here = the EXPOSED part of the exchange
    sl_ctx::ProfRec pr{};
    const bool prof = ctx->profiling && ctx->profiling_all;
    if (prof) {
        cudaEventCreate(&pr.a);
        cudaEventCreate(&pr.b);
        pr.name = "(exposed gradient exchange: compute stream waiting in sl_comm_wait)";
        cudaEventRecord(pr.a, ctx->stream);
    }
    SL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->comm_done, 0));
    if (prof) {
        cudaEventRecord(pr.b, ctx->stream);
        ctx->prof.push_back(pr);
    }
    ctx->comm_pending = false;
    return SL_OK;
}

int sl_comm_destroy(sl_ctx* ctx) {
    if (ctx && ctx->comm_stream) {
        cudaStreamSynchronize(ctx->comm_stream);
        cudaStreamDestroy(ctx->comm_stream);
        cudaEventDestroy(ctx->comm_ready);
        cudaEventDestroy(ctx->comm_done);
        for (auto e : ctx->comm_events) cudaEventDestroy(e);
        ctx->comm_events.clear();
        ctx->comm_issued = 0;
        ctx->comm_stream = nullptr;
    }
    if (!ctx || !ctx->nccl_comm) return SL_OK;
    if (nccl().ok) nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
    return SL_OK;
}

}  // extern "C"
