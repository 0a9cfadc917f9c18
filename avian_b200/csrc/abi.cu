// extern "C" surface of libavian_b200.so (declared in include/avian_b200.h).
#include <memory>
#include <mutex>

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
}  // namespace

extern "C" {

uint32_t avn_abi_version(void) { return AVN_ABI_VERSION; }

AvnStatus avn_create(const AvnConfig* config, AvnContext** out_ctx) {
    if (!config || !out_ctx) return create_fail(AVN_ERR_INVALID_ARGUMENT, "config and out_ctx are required");
    *out_ctx = nullptr;
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
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return create_fail(AVN_ERR_CUDA, cudaGetErrorString(e));
    ctx->solver.reset(avn::make_solver(config->scalar_bits, ctx->stream, &ctx->err, config->flags, config->device));
    ctx->broadphase.reset(avn::make_broadphase(config->scalar_bits, ctx->stream, &ctx->err, config->device));
    ctx->aabbs.reset(avn::make_aabb_updater(config->scalar_bits, ctx->stream, &ctx->err));
    ctx->narrow.reset(avn::make_narrow(config->scalar_bits, ctx->stream, &ctx->err));
    ctx->contacts.reset(avn::make_contacts(config->scalar_bits, ctx->stream, &ctx->err));
    if (!ctx->solver || !ctx->broadphase || !ctx->aabbs || !ctx->narrow || !ctx->contacts) return create_fail(AVN_ERR_UNSUPPORTED, "scalar type not available");
    *out_ctx = ctx.release();
    return AVN_OK;
}

void avn_destroy(AvnContext* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
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
    std::lock_guard<std::mutex> lk(g_create_mutex);
    return g_create_error.c_str();
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
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    return ctx->solver->upload(params, bodies, manifolds, joints);
}
AvnStatus avn_solver_run(AvnContext* ctx) {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    return ctx->solver->run();
}
AvnStatus avn_solver_upload_edges(AvnContext* ctx, const AvnStepParams* params, AvnBodyColumns* bodies, AvnEdgeManifolds* manifolds, AvnJointSet* joints) {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    return ctx->solver->upload_edges(params, bodies, manifolds, joints);
}
AvnStatus avn_solver_run_range(AvnContext* ctx, uint32_t first_substep, uint32_t substep_count, uint32_t run_flags) {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    return ctx->solver->run_range(first_substep, substep_count, run_flags);
}
AvnStatus avn_solver_set_boundary(AvnContext* ctx, const AvnBoundary* boundary) {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    return ctx->solver->set_boundary(boundary);
}
AvnStatus avn_solver_boundary_snapshot(AvnContext* ctx) {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    return ctx->solver->boundary_snapshot();
}
AvnStatus avn_solver_boundary_pack(AvnContext* ctx, void* device_table) {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    return ctx->solver->boundary_pack(device_table);
}
AvnStatus avn_solver_boundary_apply(AvnContext* ctx, const void* device_gathered) {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    return ctx->solver->boundary_apply(device_gathered);
}
AvnStatus avn_solver_needs_restitution(AvnContext* ctx, int* out_nonzero) {
    if (!ctx || !out_nonzero) return AVN_ERR_INVALID_ARGUMENT;
    *out_nonzero = ctx->solver->needs_restitution();
    return AVN_OK;
}
AvnStatus avn_get_stream(AvnContext* ctx, void** out_stream) {
    if (!ctx || !out_stream) return AVN_ERR_INVALID_ARGUMENT;
    *out_stream = (void*)ctx->stream;
    return AVN_OK;
}
AvnStatus avn_solver_download(AvnContext* ctx) {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    AvnStatus st = ctx->solver->download();
    ctx->solver->timings(&ctx->last);
    return st;
}
AvnStatus avn_solver_step(AvnContext* ctx, const AvnStepParams* params, AvnBodyColumns* bodies, AvnManifoldColumns* manifolds, AvnJointSet* joints) {
    AvnStatus st = avn_solver_upload(ctx, params, bodies, manifolds, joints);
    if (st != AVN_OK) return st;
    if ((st = avn_solver_run(ctx)) != AVN_OK) return st;
    return avn_solver_download(ctx);
}

AvnStatus avn_broadphase_upload(AvnContext* ctx, AvnAabbColumns* aabbs) {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    return ctx->broadphase->upload(aabbs);
}
AvnStatus avn_broadphase_run(AvnContext* ctx) {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    return ctx->broadphase->run();
}
AvnStatus avn_broadphase_download(AvnContext* ctx, AvnPairList* out_pairs) {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    AvnStatus st = ctx->broadphase->download(out_pairs);
    ctx->broadphase->timings(&ctx->last);
    return st;
}
AvnStatus avn_broadphase(AvnContext* ctx, AvnAabbColumns* aabbs, AvnPairList* out_pairs) {
    AvnStatus st = avn_broadphase_upload(ctx, aabbs);
    if (st != AVN_OK) return st;
    if ((st = avn_broadphase_run(ctx)) != AVN_OK) return st;
    return avn_broadphase_download(ctx, out_pairs);
}

AvnStatus avn_update_aabbs(AvnContext* ctx, const AvnAabbParams* params, AvnColliderColumns* colliders) {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    return ctx->aabbs->update(params, colliders);
}

AvnStatus avn_narrow_phase(AvnContext* ctx, const AvnNarrowParams* params, const AvnNarrowInput* input, AvnRawManifolds* out) {
    if (!ctx) return AVN_ERR_INVALID_ARGUMENT;
    if (!bind(ctx)) return ctx->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed");
    return ctx->narrow->run(params, input, out);
}

#define AVN_ENTER(ctx)                                                             \
    if (!(ctx)) return AVN_ERR_INVALID_ARGUMENT;                                   \
    if (!bind(ctx)) return (ctx)->err.fail(AVN_ERR_CUDA, "cudaSetDevice failed")
AvnStatus avn_contacts_reserve(AvnContext* ctx, uint32_t capacity) { AVN_ENTER(ctx); return ctx->contacts->reserve(capacity); }
AvnStatus avn_contacts_add(AvnContext* ctx, uint32_t n, const uint32_t* ids, const uint32_t* collider1, const uint32_t* collider2, const uint32_t* body1,
                           const uint32_t* body2) {
    AVN_ENTER(ctx);
    return ctx->contacts->add(n, ids, collider1, collider2, body1, body2);
}
AvnStatus avn_contacts_remove(AvnContext* ctx, uint32_t n, const uint32_t* ids) { AVN_ENTER(ctx); return ctx->contacts->remove(n, ids); }
AvnStatus avn_contacts_narrow_phase(AvnContext* ctx, const AvnNarrowParams* params, const AvnNarrowInput* input, uint32_t match_contacts, double length_unit,
                                    uint8_t* out_point_count, uint8_t* out_disjoint) {
    AVN_ENTER(ctx);
    return ctx->contacts->narrow_phase(params, input, match_contacts, length_unit, out_point_count, out_disjoint);
}
AvnStatus avn_solver_upload_graph(AvnContext* ctx, const AvnStepParams* params, AvnBodyColumns* bodies, const AvnEdgeManifolds* graph, AvnJointSet* joints) {
    AVN_ENTER(ctx);
    return ctx->solver->upload_graph(params, bodies, graph, ctx->contacts.get(), joints);
}
AvnStatus avn_contacts_download_impulses(AvnContext* ctx, void* warm_start_normal, void* warm_start_tangent, void* normal_impulse) {
    AVN_ENTER(ctx);
    return ctx->contacts->download_impulses(warm_start_normal, warm_start_tangent, normal_impulse);
}

AvnStatus avn_get_timings(const AvnContext* ctx, AvnTimings* out) {
    if (!ctx || !out) return AVN_ERR_INVALID_ARGUMENT;
    *out = ctx->last;
    return AVN_OK;
}

AvnStatus avn_joint_levels(const AvnBodyColumns* bodies, const AvnJointSet* joints, uint32_t* out_level, uint32_t* out_level_count) {
    if (!bodies || !joints) return AVN_ERR_INVALID_ARGUMENT;
    avn::JointSchedule sch;
    std::string error;
    AvnStatus st = avn::build_joint_schedule(*bodies, *joints, sch, error);
    if (st != AVN_OK) return create_fail(st, error);
    if (out_level)
        for (size_t g = 0; g < sch.level_of_global.size(); ++g) out_level[g] = uint32_t(sch.level_of_global[g]);
    if (out_level_count) *out_level_count = uint32_t(sch.n_levels);
    return AVN_OK;
}

}  // extern "C"
