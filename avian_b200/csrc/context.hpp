// Host-side plumbing shared by the ABI translation units: error reporting, grow-only device buffers.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/avian_b200.h"

namespace avn {

struct ErrorSink {
    std::string msg;
    AvnStatus fail(AvnStatus code, const char* fmt, ...) {
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        msg = buf;
        return code;
    }
};

#define AVN_CUDA(expr)                                                                                        \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            return err_->fail(_e == cudaErrorMemoryAllocation ? AVN_ERR_OUT_OF_MEMORY : AVN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, \
                              cudaGetErrorString(_e), __FILE__, __LINE__);                                   \
    } while (0)

// grow-only device allocation
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

// the context's communicator (comm.cu): NCCL bound at run time; a communicator of one needs no NCCL at all
struct CommBase {
    virtual ~CommBase() {}
    virtual AvnStatus unique_id(void* out_id) = 0;
    virtual AvnStatus init(uint32_t rank, uint32_t world, const void* id) = 0;
    virtual AvnStatus shutdown() = 0;
    virtual int rank() const = 0;
    virtual int world() const = 0;
    virtual AvnStatus all_gather(const void* send_dev, void* recv_dev, size_t bytes_per_rank) = 0;   // on the context's stream
    virtual AvnStatus all_reduce_max_i32(int* dev, size_t count) = 0;
};
CommBase* make_comm(cudaStream_t stream, ErrorSink* err);

struct ContactsBase;
struct SolverBase {
    virtual ~SolverBase() {}
    virtual AvnStatus upload(const AvnStepParams* prm, AvnBodyColumns* bodies, AvnManifoldColumns* manifolds, AvnJointSet* joints) = 0;
    virtual AvnStatus upload_edges(const AvnStepParams* prm, AvnBodyColumns* bodies, AvnEdgeManifolds* manifolds, AvnJointSet* joints) = 0;
    // the same with the edge-indexed columns already on the device (ContactsBase::view / outputs): only the graph columns are uploaded
    virtual AvnStatus upload_graph(const AvnStepParams* prm, AvnBodyColumns* bodies, const AvnEdgeManifolds* graph, ContactsBase* contacts, AvnJointSet* joints) = 0;
    // the same with the colour-major list on the device as well (ContactsBase::graph_view): nothing of the constraints crosses the bus
    virtual AvnStatus upload_resident(const AvnStepParams* prm, AvnBodyColumns* bodies, ContactsBase* contacts, AvnJointSet* joints) = 0;
    virtual AvnStatus run() = 0;
    virtual AvnStatus run_range(uint32_t first, uint32_t count, uint32_t flags) = 0;
    virtual AvnStatus set_boundary(const AvnBoundary* bnd) = 0;
    virtual AvnStatus boundary_snapshot() = 0;
    virtual AvnStatus boundary_pack(void* device_table) = 0;
    virtual AvnStatus boundary_apply(const void* device_gathered) = 0;
    // the whole partitioned stage of one rank: launches substep by substep with the boundary exchange over `comm` in between
    virtual AvnStatus step_partitioned(CommBase* comm) = 0;
    virtual int needs_restitution() const = 0;
    // start the host-to-device copy of the next upload's body columns on a second stream (overlaps whatever runs before the solver stage)
    virtual AvnStatus prefetch_bodies(AvnBodyColumns* bodies, uint32_t flags) = 0;
    virtual AvnStatus download() = 0;
    virtual void timings(AvnTimings* t) const = 0;
};
// the new pairs of the last broad-phase run where the run left them (device memory); count is known on the host after the run settled
struct DevicePairs {
    uint64_t count = 0;
    const uint32_t* c1 = nullptr; const uint32_t* c2 = nullptr; const uint32_t* b1 = nullptr; const uint32_t* b2 = nullptr;
    const uint8_t* flags = nullptr;
};

struct BroadphaseBase {
    virtual ~BroadphaseBase() {}
    virtual AvnStatus upload(AvnAabbColumns* aabbs) = 0;
    virtual AvnStatus run() = 0;
    virtual AvnStatus download(AvnPairList* out) = 0;
    // device-resident pipeline: the pairs stay on the device (the contact store takes them from there); only the persistent order and the
    // pair count go to the host
    virtual AvnStatus device_pairs(DevicePairs* out) = 0;
    virtual AvnStatus download_order(uint64_t* out_pair_count) = 0;
    // ContactGraph::pair_set kept by the contact store on the device: used as the "existing pairs" set of every later upload that
    // brings no host key list (table == NULL switches back)
    virtual void set_existing_device(const uint64_t* table, uint64_t mask) = 0;
    virtual void timings(AvnTimings* t) const = 0;
};

struct AabbBase {
    virtual ~AabbBase() {}
    virtual AvnStatus update(const AvnAabbParams* prm, AvnColliderColumns* colliders) = 0;
};
AabbBase* make_aabb_updater(uint32_t scalar_bits, cudaStream_t stream, ErrorSink* err);

struct NarrowBase {
    virtual ~NarrowBase() {}
    virtual AvnStatus run(const AvnNarrowParams* prm, const AvnNarrowInput* in, AvnRawManifolds* out) = 0;
};
NarrowBase* make_narrow(uint32_t scalar_bits, cudaStream_t stream, ErrorSink* err);

struct ContactsBase {
    virtual ~ContactsBase() {}
    virtual AvnStatus reserve(uint32_t capacity) = 0;
    virtual AvnStatus add(uint32_t n, const uint32_t* ids, const uint32_t* c1, const uint32_t* c2, const uint32_t* b1, const uint32_t* b2) = 0;
    virtual AvnStatus remove(uint32_t n, const uint32_t* ids) = 0;
    virtual AvnStatus narrow_phase(const AvnNarrowParams* prm, const AvnNarrowInput* in, uint32_t match_contacts, double length_unit, uint8_t* out_count,
                                   uint8_t* out_disjoint) = 0;
    virtual AvnStatus view(AvnEdgeManifolds* out) = 0;                       // device pointers of the edge-indexed columns
    virtual void outputs(void** ws_n, void** ws_t, void** nimp) = 0;         // device pointers store_contact_impulses writes
    virtual uint32_t capacity() const = 0;
    virtual AvnStatus download_impulses(void* ws_n, void* ws_t, void* nimp) = 0;
    // ---- the ContactGraph + ConstraintGraph on the device (contacts.cu)
    virtual AvnStatus configure(const AvnContactGraphConfig* cfg) = 0;
    // start the host-to-device copy of step()'s collider / body columns on a second stream (before the broad phase is waited for)
    virtual AvnStatus prefetch_inputs(const AvnNarrowParams* prm, const AvnNarrowInput* in, uint32_t match_contacts, double length_unit, uint32_t flags) = 0;
    virtual AvnStatus step(const AvnNarrowParams* prm, const AvnNarrowInput* in, uint32_t match_contacts, double length_unit, const DevicePairs* new_pairs,
                           AvnContactStep* out) = 0;
    struct ResidentGraph {          // device pointers of the colour-major list the last step() built
        uint32_t count = 0, any_restitution = 0;
        uint32_t color_offsets[AVN_GRAPH_COLOR_COUNT + 1] = {};
        const uint32_t* edge = nullptr; const int32_t* body1 = nullptr; const int32_t* body2 = nullptr;
        const void* friction = nullptr; const void* restitution = nullptr;
    };
    virtual AvnStatus graph_view(ResidentGraph* out) = 0;
    virtual void pair_set(const uint64_t** table, uint64_t* mask) = 0;
    virtual AvnStatus download_graph(uint32_t capacity, uint32_t* c1, uint32_t* c2, uint8_t* live, uint8_t* touching, int8_t* colour, uint32_t* edge_list) = 0;
    // ---- persistent simulation islands + sleeping decisions (contacts.cu)
    virtual AvnStatus islands_configure(const AvnIslandsConfig* cfg) = 0;
    virtual AvnStatus islands_step(AvnIslandsStep* step) = 0;
};
ContactsBase* make_contacts(uint32_t scalar_bits, cudaStream_t stream, ErrorSink* err);

SolverBase* make_solver(uint32_t scalar_bits, cudaStream_t stream, ErrorSink* err, uint32_t cfg_flags, int device);
BroadphaseBase* make_broadphase(uint32_t scalar_bits, cudaStream_t stream, ErrorSink* err, int device);

}  // namespace avn
