// api.cu -- context, device arrays, ABI self-description and NCCL plumbing of libarmon_b200.so.
#include "common.cuh"

#include <cstdarg>
#include <cstring>

namespace {
thread_local char g_last_error[1024] = "";
constexpr size_t SCRATCH_ELEMS = 2 + 2 * 65536;   // reduction scratch: 2 results + two per-row partial arrays
constexpr size_t STAGE_F32_ELEMS = 8u << 20;      // 32 MB of Float32 staging per chunk

__global__ void k_widen(double *dst, const float *src, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = (double)src[i];
}
__global__ void k_narrow(float *dst, const double *src, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = __double2float_rn(src[i]);
}
}   // namespace

void armon_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int armon_ctx_activate(armon_ctx *ctx)
{
    ARMON_CUDA(cudaSetDevice(ctx->device));
    return ARMON_OK;
}

extern "C" {

int armon_b200_abi_version(void) { return ARMON_B200_ABI_VERSION; }
int armon_flt_size(void) { return (int)sizeof(double); }
int armon_idx_size(void) { return (int)sizeof(int64_t); }
const char *armon_last_error(void) { return g_last_error; }

int armon_device_count(int *count)
{
    ARMON_CHECK_ARG(count != nullptr, "null count");
    *count = 0;
    cudaError_t err = cudaGetDeviceCount(count);
    if (err != cudaSuccess) {
        *count = 0;
        armon_set_error("no CUDA device: %s", cudaGetErrorString(err));
        return ARMON_ERR_NO_DEVICE;
    }
    return ARMON_OK;
}

int armon_ctx_create(int device, armon_ctx **out)
{
    ARMON_CHECK_ARG(out != nullptr, "null context pointer");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        armon_set_error("no CUDA device visible: the B200 backend has no CPU fallback");
        return ARMON_ERR_NO_DEVICE;
    }
    ARMON_CHECK_ARG(device >= 0 && device < count, "device ordinal out of range");
    ARMON_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ARMON_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        armon_set_error("device %d (%s) is sm_%d%d: this library only carries sm_100a code", device, prop.name,
                        prop.major, prop.minor);
        return ARMON_ERR_NO_DEVICE;
    }
    armon_ctx *ctx = new armon_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ARMON_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ARMON_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    ctx->scratch_elems = SCRATCH_ELEMS;
    ARMON_CUDA(cudaMalloc(&ctx->scratch, ctx->scratch_elems * sizeof(double)));
    ARMON_CUDA(cudaMallocHost(&ctx->pinned, 64 * sizeof(double)));
    *out = ctx;
    return ARMON_OK;
}

int armon_ctx_destroy(armon_ctx *ctx)
{
    if (!ctx) return ARMON_OK;
    cudaSetDevice(ctx->device);
    // ncclCommDestroy is collective in effect (it waits for the peers' proxies to disconnect): it belongs to
    // armon_ctx_comm_destroy, which every rank calls at the same point of the program.  A context dropped with a live
    // communicator (garbage collection, error paths: not synchronised across ranks) aborts it instead of blocking.
    if (ctx->comm) { ncclCommAbort(ctx->comm); ctx->comm = nullptr; }
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->stage_f32) cudaFree(ctx->stage_f32);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    delete ctx;
    return ARMON_OK;
}

int armon_ctx_sync(armon_ctx *ctx)
{
    ARMON_CHECK_ARG(ctx != nullptr, "null context");
    if (int rc = armon_ctx_activate(ctx)) return rc;
    ARMON_CUDA(cudaStreamSynchronize(ctx->comm_stream));
    ARMON_CUDA(cudaStreamSynchronize(ctx->stream));
    return ARMON_OK;
}

int armon_device_memory_info(armon_ctx *ctx, uint64_t *free_bytes, uint64_t *total_bytes)
{
    ARMON_CHECK_ARG(ctx && free_bytes && total_bytes, "null argument");
    if (int rc = armon_ctx_activate(ctx)) return rc;
    size_t f = 0, t = 0;
    ARMON_CUDA(cudaMemGetInfo(&f, &t));
    *free_bytes = f;
    *total_bytes = t;
    return ARMON_OK;
}

int armon_device_name(armon_ctx *ctx, char *buf, int len)
{
    ARMON_CHECK_ARG(ctx && buf && len > 0, "null argument");
    cudaDeviceProp prop;
    ARMON_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
    snprintf(buf, (size_t)len, "%s (sm_%d%d, %d SMs, device %d)", prop.name, prop.major, prop.minor,
             prop.multiProcessorCount, ctx->device);
    return ARMON_OK;
}

int armon_ctx_launch_count(armon_ctx *ctx, uint64_t *count)
{
    ARMON_CHECK_ARG(ctx && count, "null argument");
    *count = ctx->launches;
    return ARMON_OK;
}

int armon_alloc(armon_ctx *ctx, uint64_t n_elems, double **dptr)
{
    ARMON_CHECK_ARG(ctx && dptr, "null argument");
    if (int rc = armon_ctx_activate(ctx)) return rc;
    *dptr = nullptr;
    ARMON_CUDA(cudaMalloc(dptr, (n_elems ? n_elems : 1) * sizeof(double)));
    return ARMON_OK;
}

int armon_free(armon_ctx *ctx, double *dptr)
{
    ARMON_CHECK_ARG(ctx != nullptr, "null context");
    if (int rc = armon_ctx_activate(ctx)) return rc;
    if (dptr) ARMON_CUDA(cudaFree(dptr));
    return ARMON_OK;
}

int armon_copy_h2d(armon_ctx *ctx, double *dst_dev, const double *src_host, uint64_t n_elems)
{
    ARMON_CHECK_ARG(ctx && dst_dev && src_host, "null argument");
    if (int rc = armon_ctx_activate(ctx)) return rc;
    ARMON_CUDA(cudaMemcpyAsync(dst_dev, src_host, n_elems * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ARMON_CUDA(cudaStreamSynchronize(ctx->stream));   // the host buffer may be pageable / reused right away
    return ARMON_OK;
}

int armon_copy_d2h(armon_ctx *ctx, double *dst_host, const double *src_dev, uint64_t n_elems)
{
    ARMON_CHECK_ARG(ctx && dst_host && src_dev, "null argument");
    if (int rc = armon_ctx_activate(ctx)) return rc;
    ARMON_CUDA(cudaMemcpyAsync(dst_host, src_dev, n_elems * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ARMON_CUDA(cudaStreamSynchronize(ctx->stream));
    return ARMON_OK;
}

int armon_copy_d2d(armon_ctx *ctx, double *dst_dev, const double *src_dev, uint64_t n_elems)
{
    ARMON_CHECK_ARG(ctx && dst_dev && src_dev, "null argument");
    if (int rc = armon_ctx_activate(ctx)) return rc;
    ARMON_CUDA(cudaMemcpyAsync(dst_dev, src_dev, n_elems * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return ARMON_OK;
}

int armon_copy_h2d_f32(armon_ctx *ctx, double *dst_dev, const float *src_host, uint64_t n_elems)
{
    ARMON_CHECK_ARG(ctx && dst_dev && src_host, "null argument");
    if (int rc = armon_ctx_activate(ctx)) return rc;
    if (!ctx->stage_f32) ARMON_CUDA(cudaMalloc(&ctx->stage_f32, STAGE_F32_ELEMS * sizeof(float)));
    for (uint64_t off = 0; off < n_elems; off += STAGE_F32_ELEMS) {
        const size_t n = (size_t)(n_elems - off < STAGE_F32_ELEMS ? n_elems - off : STAGE_F32_ELEMS);
        ARMON_CUDA(cudaMemcpyAsync(ctx->stage_f32, src_host + off, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        k_widen<<<148 * 8, 256, 0, ctx->stream>>>(dst_dev + off, ctx->stage_f32, n);
        ARMON_LAUNCH_CHECK(ctx);
        ARMON_CUDA(cudaStreamSynchronize(ctx->stream));   // the staging buffer and the host buffer may be reused right away
    }
    return ARMON_OK;
}

int armon_copy_d2h_f32(armon_ctx *ctx, float *dst_host, const double *src_dev, uint64_t n_elems)
{
    ARMON_CHECK_ARG(ctx && dst_host && src_dev, "null argument");
    if (int rc = armon_ctx_activate(ctx)) return rc;
    if (!ctx->stage_f32) ARMON_CUDA(cudaMalloc(&ctx->stage_f32, STAGE_F32_ELEMS * sizeof(float)));
    for (uint64_t off = 0; off < n_elems; off += STAGE_F32_ELEMS) {
        const size_t n = (size_t)(n_elems - off < STAGE_F32_ELEMS ? n_elems - off : STAGE_F32_ELEMS);
        k_narrow<<<148 * 8, 256, 0, ctx->stream>>>(ctx->stage_f32, src_dev + off, n);
        ARMON_LAUNCH_CHECK(ctx);
        ARMON_CUDA(cudaMemcpyAsync(dst_host + off, ctx->stage_f32, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        ARMON_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return ARMON_OK;
}

// ---- NCCL plumbing ------------------------------------------------------------------------------------
int armon_comm_unique_id(char id[128])
{
    ARMON_CHECK_ARG(id != nullptr, "null id");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId uid;
    ARMON_NCCL(ncclGetUniqueId(&uid));
    memcpy(id, &uid, sizeof(uid));
    return ARMON_OK;
}

int armon_ctx_comm_init(armon_ctx *ctx, const char id[128], int rank, int nranks)
{
    ARMON_CHECK_ARG(ctx && id, "null argument");
    ARMON_CHECK_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "rank / nranks");
    if (int rc = armon_ctx_activate(ctx)) return rc;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ARMON_NCCL(ncclCommInitRank(&ctx->comm, nranks, uid, rank));
    ctx->rank = rank;
    ctx->nranks = nranks;
    return ARMON_OK;
}

int armon_ctx_comm_destroy(armon_ctx *ctx)
{
    ARMON_CHECK_ARG(ctx != nullptr, "null context");
    if (ctx->comm) {
        if (int rc = armon_ctx_activate(ctx)) return rc;
        ARMON_CUDA(cudaStreamSynchronize(ctx->stream));
        ARMON_NCCL(ncclCommDestroy(ctx->comm));
        ctx->comm = nullptr;
        ctx->rank = 0;
        ctx->nranks = 1;
    }
    return ARMON_OK;
}

}   // extern "C"
