// extern "C" surface of libavian_b200.so (declared in include/avian_b200.h).
#include <cstring>
#include <memory>
#include <mutex>
#include <new>

#include "context.hpp"
#include "joint_schedule.hpp"

struct AvnContext {
    int device = 0;
    uint32_t scalar_bits = 32;
    cudaStream_t stream = nullptr;
    avn::ErrorSink err;
    std::unique_ptr<avn::SolverBase> solver;
    std::unique_ptr<avn::BroadphaseBase> broadphase;
    std::unique_ptr<avn::AabbBase> aabbs;
    std::unique_ptr<avn::NarrowBase> narrow;
    std::unique_ptr<avn::ContactsBase> contacts;
    std::unique_ptr<avn::CommBase> comm;
    AvnTimings last{};
};

namespace {
std::string g_create_error;
std::mutex g_create_mutex;

AvnStatus create_fail(AvnStatus code, const std::string& msg) {
    std::lock_guard<std::mutex> lk(g_create_mutex);
    g_create_error = msg;
    return code;
}
bool bind(AvnContext* ctx) { return cudaSetDevice(ctx->device) == cudaSuccess; }

// "nothing throws or aborts across the ABI": every entry point that reaches C++ code which may allocate (std::vector, std::string, new) runs
// inside this guard; an exception becomes a status code and a message.
template <class F>
AvnStatus guarded(AvnContext* ctx, F&& body) noexcept {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    try {
        if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
        return body();
    } catch (const std::bad_alloc&) {
        try { return ctx->err.fail(AVN_ERR_OUT_OF_MEMORY, "host allocation failed"); } catch (...) { return AVN_ERR_OUT_OF_MEMORY; }
    } catch (const std::exception& e) {
        try { return ctx->err.fail(AVN_ERR_INVALID_ARGUMENT, "internal error: %s", e.what()); } catch (...) { return AVN_ERR_INVALID_ARGUMENT; }
    } catch (...) {
        return AVN_ERR_INVALID_ARGUMENT;
    }
}
thread_local char t_create_error[512];
}  // namespace

extern "C" {

uint32_t avn_abi_version(void) { return AVN_ABI_VERSION; }

AvnStatus avn_create(const AvnConfig* config, AvnContext** out_ctx) {
    if (!config || !out_ctx) return create_fail(AVN_ERR_INVALID_ARGUMENT, "config and out_ctx are required");
    *out_ctx = nullptr;
    cudaStream_t stream = nullptr;
    try {
        if (config->abi_version != AVN_ABI_VERSION) return create_fail(AVN_ERR_INVALID_ARGUMENT, "ABI version mismatch");
        if (config->scalar_bits != 32 && config->scalar_bits != 64) return create_fail(AVN_ERR_INVALID_ARGUMENT, "scalar_bits must be 32 or 64");
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            return create_fail(AVN_ERR_CUDA, std::string("no usable CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e));
        if (config->device < 0 || config->device >= count) return create_fail(AVN_ERR_INVALID_ARGUMENT, "device ordinal out of range");
        if ((e = cudaSetDevice(config->device)) != cudaSuccess) return create_fail(AVN_ERR_CUDA, cudaGetErrorString(e));
        cudaDeviceProp prop;
        if ((e = cudaGetDeviceProperties(&prop, config->device)) != cudaSuccess) return create_fail(AVN_ERR_CUDA, cudaGetErrorString(e));
        if (prop.major != 10)
            return create_fail(AVN_ERR_UNSUPPORTED, std::string("kernels are built for sm_100a only; device is ") + prop.name + " (sm_" +
                                                        std::to_string(prop.major) + std::to_string(prop.minor) + ")");
        auto ctx = std::make_unique<AvnContext>();
        ctx->device = config->device;
        ctx->scalar_bits = config->scalar_bits;
        if ((e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)) != cudaSuccess) return create_fail(AVN_ERR_CUDA, cudaGetErrorString(e));
        ctx->stream = stream;
        ctx->solver.reset(avn::make_solver(config->scalar_bits, ctx->stream, &ctx->err, config->flags, config->device));
        ctx->broadphase.reset(avn::make_broadphase(config->scalar_bits, ctx->stream, &ctx->err, config->device));
        ctx->aabbs.reset(avn::make_aabb_updater(config->scalar_bits, ctx->stream, &ctx->err));
        ctx->narrow.reset(avn::make_narrow(config->scalar_bits, ctx->stream, &ctx->err));
        ctx->contacts.reset(avn::make_contacts(config->scalar_bits, ctx->stream, &ctx->err));
        ctx->comm.reset(avn::make_comm(ctx->stream, &ctx->err));
        if (!ctx->solver || !ctx->broadphase || !ctx->aabbs || !ctx->narrow || !ctx->contacts) {
            ctx.reset();   // the members hold the stream: release them before it goes
            cudaStreamDestroy(stream);
            return create_fail(AVN_ERR_UNSUPPORTED, "scalar type not available");
        }
        *out_ctx = ctx.release();
        return AVN_OK;
    } catch (...) {
        if (stream) cudaStreamDestroy(stream);
        try { return create_fail(AVN_ERR_OUT_OF_MEMORY, "host allocation failed in avn_create"); } catch (...) { return AVN_ERR_OUT_OF_MEMORY; }
    }
}

void avn_destroy(AvnContext* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->comm.reset();
    ctx->solver.reset();
    ctx->broadphase.reset();
    ctx->aabbs.reset();
    ctx->narrow.reset();
    ctx->contacts.reset();
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* avn_last_error(const AvnContext* ctx) {
    if (ctx) return ctx->err.msg.c_str();
    std::lock_guard<std::mutex> lk(g_create_mutex);   // copied under the lock: the returned pointer stays valid for this thread
    std::strncpy(t_create_error, g_create_error.c_str(), sizeof t_create_error - 1);
    t_create_error[sizeof t_create_error - 1] = 0;
    return t_create_error;
}

AvnStatus avn_alloc_pinned(AvnContext* ctx, size_t bytes, void** out_ptr) {
    if (!ctx || !out_ptr) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    cudaError_t e = cudaHostAlloc(out_ptr, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) return ctx->err.fail(AVN_ERR_OUT_OF_MEMORY, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    return AVN_OK;
}

AvnStatus avn_free_pinned(AvnContext* ctx, void* ptr) {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!ptr) return AVN_OK;
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) return ctx->err.fail(AVN_ERR_CUDA, "cudaFreeHost: %s", cudaGetErrorString(e));
    return AVN_OK;
}

AvnStatus avn_solver_upload(AvnContext* ctx, const AvnStepParams* params, AvnBodyColumns* bodies, AvnManifoldColumns* manifolds, AvnJointSet* joints) {
    return guarded(ctx, [&] { return ctx->solver->upload(params, bodies, manifolds, joints); });
}
AvnStatus avn_solver_run(AvnContext* ctx) {
    return guarded(ctx, [&] { return ctx->solver->run(); });
}
AvnStatus avn_solver_upload_edges(AvnContext* ctx, const AvnStepParams* params, AvnBodyColumns* bodies, AvnEdgeManifolds* manifolds, AvnJointSet* joints) {
    return guarded(ctx, [&] { return ctx->solver->upload_edges(params, bodies, manifolds, joints); });
}
AvnStatus avn_solver_run_range(AvnContext* ctx, uint32_t first_substep, uint32_t substep_count, uint32_t run_flags) {
    return guarded(ctx, [&] { return ctx->solver->run_range(first_substep, substep_count, run_flags); });
}
AvnStatus avn_solver_set_boundary(AvnContext* ctx, const AvnBoundary* boundary) {
    return guarded(ctx, [&] { return ctx->solver->set_boundary(boundary); });
}
AvnStatus avn_solver_boundary_snapshot(AvnContext* ctx) {
    return guarded(ctx, [&] { return ctx->solver->boundary_snapshot(); });
}
AvnStatus avn_solver_boundary_pack(AvnContext* ctx, void* device_table) {
    return guarded(ctx, [&] { return ctx->solver->boundary_pack(device_table); });
}
AvnStatus avn_solver_boundary_apply(AvnContext* ctx, const void* device_gathered) {
    return guarded(ctx, [&] { return ctx->solver->boundary_apply(device_gathered); });
}
AvnStatus avn_solver_needs_restitution(AvnContext* ctx, int* out_nonzero) {
    if (!ctx || !out_nonzero) return AVN_ERR_INVALID_ARGUMENT;
    *out_nonzero = ctx->solver->needs_restitution();
    return AVN_OK;
}
AvnStatus avn_solver_step_partitioned(AvnContext* ctx) {
    return guarded(ctx, [&] { return ctx->solver->step_partitioned(ctx->comm.get()); });
}
AvnStatus avn_comm_unique_id(AvnContext* ctx, void* out_id) {
    return guarded(ctx, [&] { return ctx->comm->unique_id(out_id); });
}
AvnStatus avn_comm_init(AvnContext* ctx, uint32_t rank, uint32_t world, const void* unique_id) {
    return guarded(ctx, [&] { return ctx->comm->init(rank, world, unique_id); });
}
AvnStatus avn_comm_destroy(AvnContext* ctx) {
    return guarded(ctx, [&] { return ctx->comm->shutdown(); });
}
AvnStatus avn_comm_all_gather(AvnContext* ctx, const void* send_device, void* recv_device, size_t bytes_per_rank) {
    return guarded(ctx, [&] { return ctx->comm->all_gather(send_device, recv_device, bytes_per_rank); });
}
AvnStatus avn_get_stream(AvnContext* ctx, void** out_stream) {
    if (!ctx || !out_stream) return AVN_ERR_INVALID_ARGUMENT;
    *out_stream = (void*)ctx->stream;
    return AVN_OK;
}
AvnStatus avn_solver_download(AvnContext* ctx) {
    return guarded(ctx, [&] {
        AvnStatus st = ctx->solver->download();
        ctx->solver->timings(&ctx->last);
        return st;
    });
}
AvnStatus avn_solver_step(AvnContext* ctx, const AvnStepParams* params, AvnBodyColumns* bodies, AvnManifoldColumns* manifolds, AvnJointSet* joints) {
    AvnStatus st = avn_solver_upload(ctx, params, bodies, manifolds, joints);
    if (st != AVN_OK) return st;
    if ((st = avn_solver_run(ctx)) != AVN_OK) return st;
    return avn_solver_download(ctx);
}

AvnStatus avn_broadphase_upload(AvnContext* ctx, AvnAabbColumns* aabbs) {
    return guarded(ctx, [&] { return ctx->broadphase->upload(aabbs); });
}
AvnStatus avn_broadphase_run(AvnContext* ctx) {
    return guarded(ctx, [&] { return ctx->broadphase->run(); });
}
AvnStatus avn_broadphase_download(AvnContext* ctx, AvnPairList* out_pairs) {
    return guarded(ctx, [&] {
        AvnStatus st = ctx->broadphase->download(out_pairs);
        ctx->broadphase->timings(&ctx->last);
        return st;
    });
}
AvnStatus avn_broadphase(AvnContext* ctx, AvnAabbColumns* aabbs, AvnPairList* out_pairs) {
    AvnStatus st = avn_broadphase_upload(ctx, aabbs);
    if (st != AVN_OK) return st;
    if ((st = avn_broadphase_run(ctx)) != AVN_OK) return st;
    return avn_broadphase_download(ctx, out_pairs);
}

AvnStatus avn_update_aabbs(AvnContext* ctx, const AvnAabbParams* params, AvnColliderColumns* colliders) {
    return guarded(ctx, [&] { return ctx->aabbs->update(params, colliders); });
}

AvnStatus avn_narrow_phase(AvnContext* ctx, const AvnNarrowParams* params, const AvnNarrowInput* input, AvnRawManifolds* out) {
    return guarded(ctx, [&] { return ctx->narrow->run(params, input, out); });
}

AvnStatus avn_contacts_reserve(AvnContext* ctx, uint32_t capacity) { return guarded(ctx, [&] { return ctx->contacts->reserve(capacity); }); }
AvnStatus avn_contacts_add(AvnContext* ctx, uint32_t n, const uint32_t* ids, const uint32_t* collider1, const uint32_t* collider2, const uint32_t* body1,
                           const uint32_t* body2) {
    return guarded(ctx, [&] { return ctx->contacts->add(n, ids, collider1, collider2, body1, body2); });
}
AvnStatus avn_contacts_remove(AvnContext* ctx, uint32_t n, const uint32_t* ids) { return guarded(ctx, [&] { return ctx->contacts->remove(n, ids); }); }
AvnStatus avn_contacts_narrow_phase(AvnContext* ctx, const AvnNarrowParams* params, const AvnNarrowInput* input, uint32_t match_contacts, double length_unit,
                                    uint8_t* out_point_count, uint8_t* out_disjoint) {
    return guarded(ctx, [&] { return ctx->contacts->narrow_phase(params, input, match_contacts, length_unit, out_point_count, out_disjoint); });
}
AvnStatus avn_solver_upload_graph(AvnContext* ctx, const AvnStepParams* params, AvnBodyColumns* bodies, const AvnEdgeManifolds* graph, AvnJointSet* joints) {
    return guarded(ctx, [&] { return ctx->solver->upload_graph(params, bodies, graph, ctx->contacts.get(), joints); });
}
AvnStatus avn_contacts_configure(AvnContext* ctx, const AvnContactGraphConfig* config) { return guarded(ctx, [&] { return ctx->contacts->configure(config); }); }
AvnStatus avn_contacts_step(AvnContext* ctx, const AvnNarrowParams* params, const AvnNarrowInput* input, uint32_t match_contacts, double length_unit, uint32_t flags,
                            AvnContactStep* out) {
    return guarded(ctx, [&] {
        avn::DevicePairs pairs;
        const bool take = (flags & AVN_CONTACTS_TAKE_BROADPHASE_PAIRS) != 0;
        // this step's collider / body columns start moving to the device before the broad phase is waited for
        if (params && input && out) {
            AvnStatus st = ctx->contacts->prefetch_inputs(params, input, match_contacts, length_unit, flags);
            if (st != AVN_OK) return st;
        }
        if (take) {
            AvnStatus st = ctx->broadphase->device_pairs(&pairs);
            if (st != AVN_OK) return st;
        }
        AvnStatus st = ctx->contacts->step(params, input, match_contacts, length_unit, take ? &pairs : nullptr, out);
        if (st != AVN_OK) return st;
        // ContactGraph::pair_set stays on the device: the next broad phase filters against it
        const uint64_t* table = nullptr;
        uint64_t mask = 0;
        ctx->contacts->pair_set(&table, &mask);
        ctx->broadphase->set_existing_device(table, mask);
        return AVN_OK;
    });
}
AvnStatus avn_solver_upload_resident(AvnContext* ctx, const AvnStepParams* params, AvnBodyColumns* bodies, AvnJointSet* joints) {
    return guarded(ctx, [&] { return ctx->solver->upload_resident(params, bodies, ctx->contacts.get(), joints); });
}
AvnStatus avn_solver_prefetch_bodies(AvnContext* ctx, AvnBodyColumns* bodies, uint32_t flags) {
    return guarded(ctx, [&] { return ctx->solver->prefetch_bodies(bodies, flags); });
}
AvnStatus avn_broadphase_download_order(AvnContext* ctx, uint64_t* out_pair_count) { return guarded(ctx, [&] { return ctx->broadphase->download_order(out_pair_count); }); }
AvnStatus avn_contacts_download_graph(AvnContext* ctx, uint32_t capacity, uint32_t* collider1, uint32_t* collider2, uint8_t* live, uint8_t* touching, int8_t* colour,
                                      uint32_t* edge_list) {
    return guarded(ctx, [&] { return ctx->contacts->download_graph(capacity, collider1, collider2, live, touching, colour, edge_list); });
}
AvnStatus avn_islands_configure(AvnContext* ctx, const AvnIslandsConfig* config) { return guarded(ctx, [&] { return ctx->contacts->islands_configure(config); }); }
AvnStatus avn_islands_step(AvnContext* ctx, AvnIslandsStep* step) { return guarded(ctx, [&] { return ctx->contacts->islands_step(step); }); }
AvnStatus avn_contacts_download_impulses(AvnContext* ctx, void* warm_start_normal, void* warm_start_tangent, void* normal_impulse) {
    return guarded(ctx, [&] { return ctx->contacts->download_impulses(warm_start_normal, warm_start_tangent, normal_impulse); });
}

AvnStatus avn_get_timings(const AvnContext* ctx, AvnTimings* out) {
    if (!ctx || !out) return AVN_ERR_INVALID_ARGUMENT;
    *out = ctx->last;
    return AVN_OK;
}

AvnStatus avn_joint_levels(const AvnBodyColumns* bodies, const AvnJointSet* joints, uint32_t* out_level, uint32_t* out_level_count) {
    if (!bodies || !joints) return AVN_ERR_INVALID_ARGUMENT;
    try {
    avn::JointSchedule sch;
    std::string error;
    AvnStatus st = avn::build_joint_schedule(*bodies, *joints, sch, error);
    if (st != AVN_OK) return create_fail(st, error);
    if (out_level)
        for (size_t g = 0; g < sch.level_of_global.size(); ++g) out_level[g] = uint32_t(sch.level_of_global[g]);
    if (out_level_count) *out_level_count = uint32_t(sch.n_levels);
    return AVN_OK;
    } catch (...) {
        return AVN_ERR_OUT_OF_MEMORY;
    }
}

}  // extern "C"
