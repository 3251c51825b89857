// lpx_comm.cu — incumbent sharing between the one-process-per-GPU ranks of a node-sharded
// Branch & Bound: an 8-byte max-allreduce and a small all-gather over
// NCCL (NVLink 5 / NVSwitch inside one box).  NCCL is loaded at run time with dlopen so that
// liblpx.so has no link-time dependency on it; inside a PyTorch process the already loaded
// libnccl.so.2 is reused.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "lpx_common.cuh"
#include "lpx_runtime.hpp"

namespace lpx {
namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) =
        nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_api;
ncclComm_t g_comm = nullptr;
int g_world = 1, g_rank = 0;
void* g_dev = nullptr;
size_t g_dev_bytes = 0;

int load_api() {
    if (g_api.handle) return LPX_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error(std::string("cannot load libnccl.so.2: ") + dlerror());
        return LPX_E_NCCL;
    }
    g_api.handle = h;
    g_api.GetUniqueId = (decltype(g_api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_api.CommInitRank = (decltype(g_api.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_api.CommDestroy = (decltype(g_api.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_api.AllReduce = (decltype(g_api.AllReduce))dlsym(h, "ncclAllReduce");
    g_api.AllGather = (decltype(g_api.AllGather))dlsym(h, "ncclAllGather");
    g_api.GetErrorString = (decltype(g_api.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_api.GetUniqueId || !g_api.CommInitRank || !g_api.CommDestroy || !g_api.AllReduce || !g_api.AllGather) {
        set_error("libnccl.so.2 lacks a required symbol");
        g_api = NcclApi();
        return LPX_E_NCCL;
    }
    return LPX_OK;
}

int nccl_fail(ncclResult_t r, const char* what) {
    set_error(std::string("NCCL error in ") + what + ": " + (g_api.GetErrorString ? g_api.GetErrorString(r) : "?"));
    return LPX_E_NCCL;
}

int ensure_buf(size_t bytes) {
    if (g_dev_bytes >= bytes) return LPX_OK;
    if (g_dev) cudaFree(g_dev);
    g_dev = nullptr;
    g_dev_bytes = 0;
    LPX_CUDA(cudaMalloc(&g_dev, bytes));
    g_dev_bytes = bytes;
    return LPX_OK;
}

}  // namespace

// Exact merge of per-rank result buffers on the device: every 8-byte word is non-zero on at most one
// rank, so an integer sum reproduces the owner's bits (no rounding, no sign-of-zero loss).
int comm_merge_u64(unsigned long long* dev, size_t words, cudaStream_t s) {
    if (!g_comm || words == 0) return LPX_OK;
    ncclResult_t r = g_api.AllReduce(dev, dev, words, ncclUint64, ncclSum, g_comm, s);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce(u64 sum)");
    return LPX_OK;
}
// All-gather of device buffers on stream s (bytes per rank; recv holds world * bytes).  One rank: a copy.
int comm_allgather_dev(const void* send, void* recv, size_t bytes, cudaStream_t s) {
    if (!g_comm) {
        LPX_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, s));
        return LPX_OK;
    }
    ncclResult_t r = g_api.AllGather(send, recv, bytes, ncclChar, g_comm, s);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllGather");
    return LPX_OK;
}
int comm_world() { return g_world; }
int comm_rank() { return g_rank; }
}  // namespace lpx

using namespace lpx;

extern "C" {

int lpx_comm_unique_id(void* id128) {
    if (!id128) return LPX_E_BAD_ARGS;
    int rc = load_api();
    if (rc != LPX_OK) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ncclResult_t r = g_api.GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
    std::memcpy(id128, &id, 128);
    return LPX_OK;
}

int lpx_comm_init(int world, int rank, const void* id128) {
    if (world < 1 || rank < 0 || rank >= world || !id128) {
        set_error("lpx_comm_init: bad arguments");
        return LPX_E_BAD_ARGS;
    }
    int rc = ensure_device();
    if (rc != LPX_OK) return rc;
    if ((rc = load_api()) != LPX_OK) return rc;
    if (g_comm) lpx_comm_destroy();
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclResult_t r = g_api.CommInitRank(&g_comm, world, id, rank);
    if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitRank");
    g_world = world;
    g_rank = rank;
    return LPX_OK;
}

int lpx_comm_allreduce_max(double* values, int count) {
    if (!values || count < 1) return LPX_E_BAD_ARGS;
    if (!g_comm) return LPX_OK;  // single rank: identity
    int rc = ensure_buf((size_t)count * 8);
    if (rc != LPX_OK) return rc;
    cudaStream_t s = rt().stream;
    LPX_CUDA(cudaMemcpyAsync(g_dev, values, (size_t)count * 8, cudaMemcpyHostToDevice, s));
    ncclResult_t r = g_api.AllReduce(g_dev, g_dev, (size_t)count, ncclDouble, ncclMax, g_comm, s);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
    LPX_CUDA(cudaMemcpyAsync(values, g_dev, (size_t)count * 8, cudaMemcpyDeviceToHost, s));
    LPX_CUDA(cudaStreamSynchronize(s));
    return LPX_OK;
}

int lpx_comm_allgather(const void* send, void* recv, size_t bytes_per_rank) {
    if (!send || !recv || bytes_per_rank == 0) return LPX_E_BAD_ARGS;
    if (!g_comm) {
        std::memcpy(recv, send, bytes_per_rank);
        return LPX_OK;
    }
    int rc = ensure_buf(bytes_per_rank * (size_t)(g_world + 1));
    if (rc != LPX_OK) return rc;
    cudaStream_t s = rt().stream;
    char* sendbuf = (char*)g_dev;
    char* recvbuf = sendbuf + bytes_per_rank;
    LPX_CUDA(cudaMemcpyAsync(sendbuf, send, bytes_per_rank, cudaMemcpyHostToDevice, s));
    ncclResult_t r = g_api.AllGather(sendbuf, recvbuf, bytes_per_rank, ncclChar, g_comm, s);
    if (r != ncclSuccess) return nccl_fail(r, "ncclAllGather");
    LPX_CUDA(cudaMemcpyAsync(recv, recvbuf, bytes_per_rank * (size_t)g_world, cudaMemcpyDeviceToHost, s));
    LPX_CUDA(cudaStreamSynchronize(s));
    return LPX_OK;
}

void lpx_comm_destroy(void) {
    if (g_comm && g_api.CommDestroy) g_api.CommDestroy(g_comm);
    g_comm = nullptr;
    g_world = 1;
    g_rank = 0;
    if (g_dev) cudaFree(g_dev);
    g_dev = nullptr;
    g_dev_bytes = 0;
}

int lpx_comm_world(void) { return g_world; }
int lpx_comm_rank(void) { return g_rank; }

}  // extern "C"
