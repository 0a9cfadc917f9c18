// The multi-GPU communicator behind the C ABI: one process per GPU, one ncclComm_t per AvnContext (SURVEY.md 8b "one AvnCtx spans N devices,
// NCCL communicator created by the library"; VERDICT r1 "the multi-GPU data plane lives in Python").  NCCL is bound at run time with dlopen —
// the library has no link-time dependency on it, loads on a box without NCCL, and shares the copy a host process has already loaded (a
// PyTorch host: torch's bundled libnccl.so.2; a Rust/Bevy host: the system one).  AVN_NCCL_LIB overrides the name.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "context.hpp"

namespace avn {
namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string error;

    bool load() {
        if (handle) return true;
        const char* names[] = {getenv("AVN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
            error = dlerror();
        }
        if (!handle) return false;
#define AVN_SYM(field, sym)                                                        \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, sym));                 \
    if (!field) { error = std::string("missing symbol ") + sym; handle = nullptr; return false; }
        AVN_SYM(GetUniqueId, "ncclGetUniqueId");
        AVN_SYM(CommInitRank, "ncclCommInitRank");
        AVN_SYM(CommDestroy, "ncclCommDestroy");
        AVN_SYM(AllGather, "ncclAllGather");
        AVN_SYM(AllReduce, "ncclAllReduce");
        AVN_SYM(GetErrorString, "ncclGetErrorString");
        AVN_SYM(GetVersion, "ncclGetVersion");
#undef AVN_SYM
        return true;
    }
};

NcclApi& nccl() {
    static NcclApi api;
    return api;
}
std::mutex g_nccl_mutex;

class Comm final : public CommBase {
   public:
    Comm(cudaStream_t stream, ErrorSink* err) : stream_(stream), err_(err) {}
    ~Comm() override { destroy(); }

    AvnStatus unique_id(void* out) override {
        if (!out) return err_->fail(AVN_ERR_INVALID_ARGUMENT, "out_id is required");
        std::lock_guard<std::mutex> lk(g_nccl_mutex);
        if (!nccl().load()) return err_->fail(AVN_ERR_NCCL, "NCCL not available: %s", nccl().error.c_str());
        ncclUniqueId id;
        ncclResult_t r = nccl().GetUniqueId(&id);
        if (r != ncclSuccess) return err_->fail(AVN_ERR_NCCL, "ncclGetUniqueId: %s", nccl().GetErrorString(r));
        static_assert(sizeof id == AVN_COMM_ID_BYTES, "ncclUniqueId size");
        memcpy(out, &id, sizeof id);
        return AVN_OK;
    }
    AvnStatus init(uint32_t rank, uint32_t world, const void* id_bytes) override {
        if (world == 0 || rank >= world) return err_->fail(AVN_ERR_INVALID_ARGUMENT, "comm: rank %u / world %u", rank, world);
        destroy();
        rank_ = int(rank);
        world_ = int(world);
        if (world == 1) return AVN_OK;   // a communicator of one: every collective is a local copy
        if (!id_bytes) return err_->fail(AVN_ERR_INVALID_ARGUMENT, "comm: the unique id is required for world > 1");
        {
            std::lock_guard<std::mutex> lk(g_nccl_mutex);
            if (!nccl().load()) return err_->fail(AVN_ERR_NCCL, "NCCL not available: %s", nccl().error.c_str());
        }
        ncclUniqueId id;
        memcpy(&id, id_bytes, sizeof id);
        ncclResult_t r = nccl().CommInitRank(&comm_, int(world), id, int(rank));
        if (r != ncclSuccess) {
            comm_ = nullptr;
            return err_->fail(AVN_ERR_NCCL, "ncclCommInitRank(rank %u of %u): %s", rank, world, nccl().GetErrorString(r));
        }
        return AVN_OK;
    }
    void destroy() {
        if (comm_) {
            cudaStreamSynchronize(stream_);
            nccl().CommDestroy(comm_);
            comm_ = nullptr;
        }
        world_ = 1;
        rank_ = 0;
    }
    AvnStatus shutdown() override { destroy(); return AVN_OK; }
    int rank() const override { return rank_; }
    int world() const override { return world_; }

    AvnStatus all_gather(const void* send, void* recv, size_t bytes) override {
        if (bytes == 0) return AVN_OK;
        if (!send || !recv) return err_->fail(AVN_ERR_INVALID_ARGUMENT, "all_gather: device pointers are required");
        if (world_ == 1) {
            if (send != recv) AVN_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, stream_));
            return AVN_OK;
        }
        if (!comm_) return err_->fail(AVN_ERR_NCCL, "all_gather before avn_comm_init");
        ncclResult_t r = nccl().AllGather(send, recv, bytes, ncclUint8, comm_, stream_);
        if (r != ncclSuccess) return err_->fail(AVN_ERR_NCCL, "ncclAllGather(%zu bytes): %s", bytes, nccl().GetErrorString(r));
        return AVN_OK;
    }
    AvnStatus all_reduce_max_i32(int* dev, size_t count) override {
        if (world_ == 1 || count == 0) return AVN_OK;
        if (!comm_) return err_->fail(AVN_ERR_NCCL, "all_reduce before avn_comm_init");
        ncclResult_t r = nccl().AllReduce(dev, dev, count, ncclInt32, ncclMax, comm_, stream_);
        if (r != ncclSuccess) return err_->fail(AVN_ERR_NCCL, "ncclAllReduce: %s", nccl().GetErrorString(r));
        return AVN_OK;
    }

   private:
    cudaStream_t stream_;
    ErrorSink* err_;
    ncclComm_t comm_ = nullptr;
    int rank_ = 0, world_ = 1;
};

}  // namespace

CommBase* make_comm(cudaStream_t stream, ErrorSink* err) { return new Comm(stream, err); }

}  // namespace avn
