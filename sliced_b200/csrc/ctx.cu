// ctx.cu — context, memory and transfer entry points of the C ABI (include/sliced_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

static thread_local std::string g_noctx_error;

int sl_set_error(sl_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    else g_noctx_error = buf;
    return code;
}

// true while ctx->stream is being captured into a CUDA graph (sl_graph_begin .. _end): nothing may synchronise, allocate or free
static bool capturing(sl_ctx* ctx) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return st != cudaStreamCaptureStatusNone;
}

static int grow(sl_ctx* ctx, void** p, size_t* cur, size_t bytes, void** out) {
    if (bytes > *cur) {
        if (capturing(ctx))
            return sl_set_error(ctx, SL_ERR_INVALID_ARG, "scratch would have to grow (%zu -> %zu bytes) during graph capture: run the sequence once eagerly first", *cur, bytes);
        // grow-only; in-flight kernels using the old block are ordered before the free on this stream
        if (*p) {
            SL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            SL_CUDA(ctx, cudaFree(*p));
            *p = nullptr;
            *cur = 0;
        }
        size_t want = bytes + (bytes >> 3);
        want = (want + 255) & ~size_t(255);
        SL_CUDA(ctx, cudaMalloc(p, want));
        *cur = want;
    }
    *out = *p;
    return SL_OK;
}

int sl_ws_reserve(sl_ctx* ctx, size_t bytes, void** out) { return grow(ctx, &ctx->ws, &ctx->ws_bytes, bytes, out); }
int sl_ws2_reserve(sl_ctx* ctx, size_t bytes, void** out) { return grow(ctx, &ctx->ws2, &ctx->ws2_bytes, bytes, out); }

extern "C" {

int sl_abi_version(void) { return SL_ABI_VERSION; }

int sl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int ctx_create_common(int device, void* stream, bool borrow, sl_ctx** out_ctx) {
    if (!out_ctx) return sl_set_error(nullptr, SL_ERR_INVALID_ARG, "sl_ctx_create: out_ctx is NULL");
    *out_ctx = nullptr;
    int n = sl_device_count();
    if (n <= 0) return sl_set_error(nullptr, SL_ERR_NO_DEVICE, "sl_ctx_create: no CUDA device visible (this library has no CPU fallback)");
    if (device < 0 || device >= n) return sl_set_error(nullptr, SL_ERR_INVALID_ARG, "sl_ctx_create: device %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return sl_set_error(nullptr, SL_ERR_CUDA, "sl_ctx_create: cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return sl_set_error(nullptr, SL_ERR_UNSUPPORTED, "sl_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                            device, prop.major, prop.minor);
    sl_ctx* ctx = new sl_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete ctx;
        return sl_set_error(nullptr, SL_ERR_CUDA, "sl_ctx_create: cudaSetDevice(%d) failed", device);
    }
    if (borrow) {
        ctx->stream = (cudaStream_t)stream;
        ctx->owns_stream = false;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return sl_set_error(nullptr, SL_ERR_CUDA, "sl_ctx_create: cudaStreamCreate failed");
        }
        ctx->owns_stream = true;
    }
    const char* env = getenv("SLICED_GEMM_MODE");
    if (env) {
        if (!strcmp(env, "tf32")) ctx->gemm_mode = SL_GEMM_TF32;
        else if (!strcmp(env, "simt")) ctx->gemm_mode = SL_GEMM_SIMT;
        else if (!strcmp(env, "3xf16")) ctx->gemm_mode = SL_GEMM_3XF16;
        else if (!strcmp(env, "3xtf32")) ctx->gemm_mode = SL_GEMM_3XTF32;
    }
    *out_ctx = ctx;
    return SL_OK;
}

int sl_ctx_create(int device, sl_ctx** out_ctx) { return ctx_create_common(device, nullptr, false, out_ctx); }
int sl_ctx_create_on_stream(int device, void* cuda_stream, sl_ctx** out_ctx) {
    return ctx_create_common(device, cuda_stream, true, out_ctx);
}

int sl_comm_destroy(sl_ctx* ctx);

int sl_ctx_destroy(sl_ctx* ctx) {
    if (!ctx) return SL_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    sl_comm_destroy(ctx);
    for (auto& e : ctx->plane_cache) {
        if (e.hi) cudaFree(e.hi);
        if (e.lo) cudaFree(e.lo);
    }
    for (auto& e : ctx->rowplane_cache) {
        if (e.hi) cudaFree(e.hi);
        if (e.lo) cudaFree(e.lo);
        if (e.inv) cudaFree(e.inv);
    }
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->ws2) cudaFree(ctx->ws2);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
        cudaEventDestroy(ctx->copy_done);
        cudaEventDestroy(ctx->compute_done);
    }
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return SL_OK;
}

const char* sl_last_error_string(sl_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_noctx_error.c_str(); }
void* sl_ctx_stream(sl_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int sl_ctx_device(sl_ctx* ctx) { return ctx ? ctx->device : -1; }
uint64_t sl_ctx_launch_count(sl_ctx* ctx) { return ctx ? ctx->launches : 0; }

int sl_ctx_set_gemm_mode(sl_ctx* ctx, int mode) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    SL_REQUIRE(ctx, mode >= SL_GEMM_3XTF32 && mode <= SL_GEMM_3XF16, "bad mode");
    ctx->gemm_mode = mode;
    return SL_OK;
}

int sl_ctx_profile_begin(sl_ctx* ctx) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    for (auto& r : ctx->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    ctx->prof.clear();
    ctx->profiling = true;
    return SL_OK;
}

// Per-kernel report of everything launched since sl_ctx_profile_begin (call BEFORE sl_ctx_profile_end): one line
// "name,launches,total_ms" per kernel, sorted by time.  Enables timing of every launch from the next _begin on.
int sl_ctx_profile_report(sl_ctx* ctx, char* out, size_t cap) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    ctx->profiling_all = true;
    if (!out || cap == 0) return SL_OK;
    out[0] = 0;
    SL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<std::pair<std::string, std::pair<int, double>>> agg;
    for (auto& r : ctx->prof) {
        float t = 0;
        if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) continue;
        const std::string nm = r.name ? r.name : "gemm_mma";
        bool found = false;
        for (auto& a : agg)
            if (a.first == nm) { a.second.first++; a.second.second += t; found = true; break; }
        if (!found) agg.push_back({nm, {1, (double)t}});
    }
    for (size_t i = 0; i < agg.size(); ++i)
        for (size_t j = i + 1; j < agg.size(); ++j)
            if (agg[j].second.second > agg[i].second.second) std::swap(agg[i], agg[j]);
    size_t pos = 0;
    for (auto& a : agg) {
        int n = snprintf(out + pos, cap - pos, "%s,%d,%.4f\n", a.first.c_str(), a.second.first, a.second.second);
        if (n < 0 || (size_t)n >= cap - pos) break;
        pos += n;
    }
    return SL_OK;
}

int sl_ctx_profile_end(sl_ctx* ctx, uint64_t* n_launches, double* total_ms, double* total_flops) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    ctx->profiling = false;
    SL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double ms = 0, fl = 0;
    uint64_t n_mma = 0;
    for (auto& r : ctx->prof) {
        float t = 0;
        SL_CUDA(ctx, cudaEventElapsedTime(&t, r.a, r.b));
        if (r.flops > 0) {   // the MMA kernels (the other records exist only in profiling_all mode)
            ms += t;
            fl += r.flops;
            n_mma++;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    if (n_launches) *n_launches = n_mma;
    if (total_ms) *total_ms = ms;
    if (total_flops) *total_flops = fl;
    ctx->prof.clear();
    return SL_OK;
}

int sl_malloc(sl_ctx* ctx, size_t bytes, void** out_dptr) {
    SL_REQUIRE(ctx, ctx && out_dptr, "NULL argument");
    *out_dptr = nullptr;
    if (bytes == 0) return SL_OK;
    if (capturing(ctx)) return sl_set_error(ctx, SL_ERR_INVALID_ARG, "sl_malloc during graph capture (the captured sequence must be allocation-free)");
    SL_CUDA(ctx, cudaSetDevice(ctx->device));
    SL_CUDA(ctx, cudaMalloc(out_dptr, bytes));
    return SL_OK;
}

int sl_free(sl_ctx* ctx, void* dptr) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    if (!dptr) return SL_OK;
    if (capturing(ctx)) return sl_set_error(ctx, SL_ERR_INVALID_ARG, "sl_free during graph capture (the captured sequence must be allocation-free)");
    SL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->comm_stream) SL_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));   // the buffer may still be in an overlapped exchange
    if (ctx->copy_stream) SL_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    SL_CUDA(ctx, cudaFree(dptr));
    return SL_OK;
}

int sl_host_alloc(sl_ctx* ctx, size_t bytes, void** out_hptr) {
    SL_REQUIRE(ctx, ctx && out_hptr, "NULL argument");
    SL_CUDA(ctx, cudaHostAlloc(out_hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return SL_OK;
}

int sl_host_free(sl_ctx* ctx, void* hptr) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    if (hptr) SL_CUDA(ctx, cudaFreeHost(hptr));
    return SL_OK;
}

int sl_write(sl_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, dst_dev);
    if (bytes == 0) return SL_OK;
    SL_REQUIRE(ctx, dst_dev && src_host, "NULL pointer");
    SL_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SL_OK;
}

static int ensure_copy_stream(sl_ctx* ctx) {
    if (!ctx->copy_stream) {
        SL_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        SL_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
        SL_CUDA(ctx, cudaEventCreateWithFlags(&ctx->compute_done, cudaEventDisableTiming));
    }
    return SL_OK;
}

int sl_write_prefetch(sl_ctx* ctx, void* dst_dev, const void* src_host_pinned, size_t bytes) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, dst_dev);
    if (bytes == 0) return SL_OK;
    SL_REQUIRE(ctx, dst_dev && src_host_pinned, "NULL pointer");
    int rc = ensure_copy_stream(ctx);
    if (rc != SL_OK) return rc;
    SL_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host_pinned, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    return SL_OK;
}

int sl_prefetch_wait(sl_ctx* ctx) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    if (!ctx->copy_stream) return SL_OK;
    SL_CUDA(ctx, cudaEventRecord(ctx->copy_done, ctx->copy_stream));
    SL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
    return SL_OK;
}

int sl_prefetch_release(sl_ctx* ctx) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    int rc = ensure_copy_stream(ctx);
    if (rc != SL_OK) return rc;
    SL_CUDA(ctx, cudaEventRecord(ctx->compute_done, ctx->stream));
    SL_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->compute_done, 0));
    return SL_OK;
}

int sl_read(sl_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    if (bytes == 0) return SL_OK;
    SL_REQUIRE(ctx, dst_host && src_dev, "NULL pointer");
    SL_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SL_OK;
}

// device -> (pinned) host without blocking: the copy is ordered on the ctx stream like any op; the data is on the host after the
// next sl_sync / sl_read.  A training loop reads each step's loss this way without stalling the launch of the next step.
int sl_read_async(sl_ctx* ctx, void* dst_host_pinned, const void* src_dev, size_t bytes) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    if (bytes == 0) return SL_OK;
    SL_REQUIRE(ctx, dst_host_pinned && src_dev, "NULL pointer");
    SL_CUDA(ctx, cudaMemcpyAsync(dst_host_pinned, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return SL_OK;
}

int sl_copy(sl_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, dst_dev);
    if (bytes == 0) return SL_OK;
    SL_REQUIRE(ctx, dst_dev && src_dev, "NULL pointer");
    SL_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return SL_OK;
}

int sl_clear(sl_ctx* ctx, void* dst_dev, size_t bytes) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    sl_note_writes(ctx, dst_dev);
    if (bytes == 0) return SL_OK;
    SL_REQUIRE(ctx, dst_dev, "NULL pointer");
    SL_CUDA(ctx, cudaMemsetAsync(dst_dev, 0, bytes, ctx->stream));
    return SL_OK;
}

int sl_sync(sl_ctx* ctx) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    SL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SL_OK;
}

// ---- CUDA-graph capture / replay of a call sequence (the custos `Lazy` + `run()` seam, examples/sine_net.rs:178-233)
int sl_graph_begin(sl_ctx* ctx) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    SL_REQUIRE(ctx, !ctx->profiling, "profiling and graph capture are exclusive");
    SL_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    return SL_OK;
}

int sl_graph_end(sl_ctx* ctx, void** out_graph_exec) {
    SL_REQUIRE(ctx, ctx && out_graph_exec, "NULL argument");
    *out_graph_exec = nullptr;
    cudaGraph_t g = nullptr;
    SL_CUDA(ctx, cudaStreamEndCapture(ctx->stream, &g));
    cudaGraphExec_t ge = nullptr;
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return sl_set_error(ctx, SL_ERR_CUDA, "sl_graph_end: cudaGraphInstantiate -> %s", cudaGetErrorString(e));
    *out_graph_exec = ge;
    return SL_OK;
}

int sl_graph_launch(sl_ctx* ctx, void* graph_exec) {
    SL_REQUIRE(ctx, ctx && graph_exec, "NULL argument");
    SL_CUDA(ctx, cudaGraphLaunch((cudaGraphExec_t)graph_exec, ctx->stream));
    ctx->launches++;
    return SL_OK;
}

int sl_graph_destroy(sl_ctx* ctx, void* graph_exec) {
    SL_REQUIRE(ctx, ctx != nullptr, "ctx is NULL");
    if (graph_exec) SL_CUDA(ctx, cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
    return SL_OK;
}

}  // extern "C"
