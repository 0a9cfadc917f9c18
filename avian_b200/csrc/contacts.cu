// Device-resident contact edges (protocol: plugins.ResidentWorld).
// One row per ContactId (contact_graph.rs:521-631 assigns them; the host keeps that graph).  A row holds the pair (colliders, bodies), the
// manifold the last narrow phase found (4 point slots, column scalar type: what the solver reads through avn_solver_upload_graph), the
// unrounded anchors of that manifold (double: what the next step's match_contacts compares) and the warm-start impulses (in = what the
// next solve starts from, written by the matching; out = what the last solve left, written by store_contact_impulses).
// Per step only point counts and disjoint flags go to the host (2 B per row) and the edge list of the constraint graph comes back.
#include "context.hpp"
#include "contact_rows.hpp"

namespace avn {
namespace {

__global__ void edge_add_kernel(int n, const uint32_t* __restrict__ ids, const uint32_t* __restrict__ c1, const uint32_t* __restrict__ c2,
                                const uint32_t* __restrict__ b1, const uint32_t* __restrict__ b2, uint32_t* rc1, uint32_t* rc2, uint32_t* rb1, uint32_t* rb2,
                                uint8_t* live, uint8_t* count, uint8_t* prev_count) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t e = ids[k];
    rc1[e] = c1[k]; rc2[e] = c2[k]; rb1[e] = b1[k]; rb2[e] = b2[k];
    live[e] = 1;        // a ContactId handed to a new pair starts without history
    count[e] = 0;
    prev_count[e] = 0;
}
__global__ void edge_remove_kernel(int n, const uint32_t* __restrict__ ids, uint8_t* live, uint8_t* count, uint8_t* prev_count) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t e = ids[k];
    live[e] = 0; count[e] = 0; prev_count[e] = 0;
}

template <class S>
__global__ void __launch_bounds__(128) narrow_edges_kernel(const __grid_constant__ NarrowEdgeArgs<S> a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < a.r.E) narrow_edge_row<S>(a, e);   // csrc/contact_rows.hpp: the same function the CPU tests run
}

template <class S>
class Contacts final : public ContactsBase {
   public:
    Contacts(cudaStream_t stream, ErrorSink* err) : stream_(stream), err_(err) {}

    AvnStatus reserve(uint32_t capacity) override {
        if (capacity <= E_) return AVN_OK;
        const size_t n = capacity;
        struct Col { DevBuf* buf; size_t bytes_per_row; };
        Col cols[] = {{&c1_, 4}, {&c2_, 4}, {&b1_, 4}, {&b2_, 4}, {&live_, 1}, {&count_, 1}, {&disjoint_, 1}, {&normal_, 3 * sizeof(S)}, {&a1_, 12 * sizeof(S)},
                      {&a2_, 12 * sizeof(S)}, {&pen_, 4 * sizeof(S)}, {&ns_, 4 * sizeof(S)}, {&prev_count_, 1}, {&prev_a1_, 12 * sizeof(double)},
                      {&prev_a2_, 12 * sizeof(double)}, {&ws_n_in_, 4 * sizeof(S)}, {&ws_t_in_, 8 * sizeof(S)}, {&ws_n_out_, 4 * sizeof(S)},
                      {&ws_t_out_, 8 * sizeof(S)}, {&nimp_in_, 4 * sizeof(S)}, {&nimp_out_, 4 * sizeof(S)}};
        for (Col& c : cols) {   // grow, keep the old rows, zero the new ones
            void* fresh = nullptr;
            AVN_CUDA(cudaMalloc(&fresh, n * c.bytes_per_row));
            AVN_CUDA(cudaMemsetAsync(fresh, 0, n * c.bytes_per_row, stream_));
            if (c.buf->p && E_) AVN_CUDA(cudaMemcpyAsync(fresh, c.buf->p, size_t(E_) * c.bytes_per_row, cudaMemcpyDeviceToDevice, stream_));
            AVN_CUDA(cudaStreamSynchronize(stream_));
            if (c.buf->p) cudaFree(c.buf->p);
            c.buf->p = fresh;
            c.buf->cap = n * c.bytes_per_row;
        }
        E_ = capacity;
        return AVN_OK;
    }

    AvnStatus add(uint32_t n, const uint32_t* ids, const uint32_t* c1, const uint32_t* c2, const uint32_t* b1, const uint32_t* b2) override {
        if (n == 0) return AVN_OK;
        if (!ids || !c1 || !c2 || !b1 || !b2) return err_->fail(AVN_ERR_INVALID_ARGUMENT, "contacts_add: every array is required");
        for (uint32_t k = 0; k < n; ++k)
            if (ids[k] >= E_) return err_->fail(AVN_ERR_CAPACITY, "contacts_add: id %u >= capacity %u (avn_contacts_reserve first)", ids[k], E_);
        AVN_CUDA(stage_.ensure(size_t(5) * n * 4));
        uint32_t* s = stage_.as<uint32_t>();
        const uint32_t* src[5] = {ids, c1, c2, b1, b2};
        for (int c = 0; c < 5; ++c) AVN_CUDA(cudaMemcpyAsync(s + size_t(c) * n, src[c], size_t(n) * 4, cudaMemcpyHostToDevice, stream_));
        edge_add_kernel<<<(n + 255) / 256, 256, 0, stream_>>>(int(n), s, s + n, s + 2 * size_t(n), s + 3 * size_t(n), s + 4 * size_t(n), c1_.as<uint32_t>(),
                                                              c2_.as<uint32_t>(), b1_.as<uint32_t>(), b2_.as<uint32_t>(), live_.as<uint8_t>(), count_.as<uint8_t>(),
                                                              prev_count_.as<uint8_t>());
        AVN_CUDA(cudaGetLastError());
        AVN_CUDA(cudaStreamSynchronize(stream_));   // the host arrays may be reused by the caller
        return AVN_OK;
    }

    AvnStatus remove(uint32_t n, const uint32_t* ids) override {
        if (n == 0) return AVN_OK;
        if (!ids) return err_->fail(AVN_ERR_INVALID_ARGUMENT, "contacts_remove: ids are required");
        for (uint32_t k = 0; k < n; ++k)
            if (ids[k] >= E_) return err_->fail(AVN_ERR_INVALID_ARGUMENT, "contacts_remove: id %u >= capacity %u", ids[k], E_);
        AVN_CUDA(stage_.ensure(size_t(n) * 4));
        AVN_CUDA(cudaMemcpyAsync(stage_.p, ids, size_t(n) * 4, cudaMemcpyHostToDevice, stream_));
        edge_remove_kernel<<<(n + 255) / 256, 256, 0, stream_>>>(int(n), stage_.as<uint32_t>(), live_.as<uint8_t>(), count_.as<uint8_t>(), prev_count_.as<uint8_t>());
        AVN_CUDA(cudaGetLastError());
        AVN_CUDA(cudaStreamSynchronize(stream_));
        return AVN_OK;
    }

    AvnStatus narrow_phase(const AvnNarrowParams* prm, const AvnNarrowInput* in, uint32_t match_contacts, double length_unit, uint8_t* out_count,
                           uint8_t* out_disjoint) override {
        if (!prm || !in || !out_count || !out_disjoint) return err_->fail(AVN_ERR_INVALID_ARGUMENT, "contacts_narrow_phase: params, input and outputs are required");
        if (E_ == 0) return AVN_OK;
        const size_t C = in->collider_count, B = in->body_count;
        if (!in->dims || !in->position || !in->rotation || !in->linear_velocity || !in->angular_velocity || !in->aabb_min || !in->aabb_max)
            return err_->fail(AVN_ERR_INVALID_ARGUMENT, "contacts_narrow_phase: dims, position, rotation, velocities and AABBs are required");
        NarrowEdgeArgs<S> a{};
        a.r = rows();
        AvnStatus st;
#define UPC(buf, host, cnt, T, dst) if ((st = up<T>(buf, host, cnt, &dst)) != AVN_OK) return st
        UPC(i_shape_, in->shape, C, uint8_t, a.shape);
        UPC(i_dims_, in->dims, 3 * C, S, a.dims);
        UPC(i_pos_, in->position, 3 * C, S, a.pos);
        UPC(i_rot_, in->rotation, 4 * C, S, a.rot);
        UPC(i_lv_, in->linear_velocity, 3 * B, S, a.lv);
        UPC(i_av_, in->angular_velocity, 3 * B, S, a.av);
        UPC(i_amin_, in->aabb_min, 3 * C, S, a.amin);
        UPC(i_amax_, in->aabb_max, 3 * C, S, a.amax);
#undef UPC
        a.dt = prm->dt;
        a.tol = prm->contact_tolerance;
        a.thr2 = (0.1 * length_unit) * (0.1 * length_unit);
        a.match = match_contacts ? 1 : 0;
        narrow_edges_kernel<S><<<(E_ + 127) / 128, 128, 0, stream_>>>(a);
        AVN_CUDA(cudaGetLastError());
        AVN_CUDA(cudaMemcpyAsync(out_count, count_.p, E_, cudaMemcpyDeviceToHost, stream_));
        AVN_CUDA(cudaMemcpyAsync(out_disjoint, disjoint_.p, E_, cudaMemcpyDeviceToHost, stream_));
        AVN_CUDA(cudaStreamSynchronize(stream_));
        return AVN_OK;
    }

    AvnStatus view(AvnEdgeManifolds* out) override {   // DEVICE pointers: the source of avn_solver_upload_graph
        if (!out) return err_->fail(AVN_ERR_INVALID_ARGUMENT, "out is required");
        out->edge_capacity = E_;
        out->point_count = count_.as<uint8_t>();
        out->normal = normal_.p; out->anchor1 = a1_.p; out->anchor2 = a2_.p; out->penetration = pen_.p; out->normal_speed = ns_.p;
        out->warm_start_normal_impulse = ws_n_in_.p;
        out->warm_start_tangent_impulse = ws_t_in_.p;
        out->normal_impulse = nimp_in_.p;
        return AVN_OK;
    }
    void outputs(void** ws_n, void** ws_t, void** nimp) override { *ws_n = ws_n_out_.p; *ws_t = ws_t_out_.p; *nimp = nimp_out_.p; }
    uint32_t capacity() const override { return E_; }

    AvnStatus download_impulses(void* ws_n, void* ws_t, void* nimp) override {   // tests / debugging: the solver's outputs per edge
        if (E_ == 0) return AVN_OK;
        if (ws_n) AVN_CUDA(cudaMemcpyAsync(ws_n, ws_n_out_.p, size_t(E_) * 4 * sizeof(S), cudaMemcpyDeviceToHost, stream_));
        if (ws_t) AVN_CUDA(cudaMemcpyAsync(ws_t, ws_t_out_.p, size_t(E_) * 8 * sizeof(S), cudaMemcpyDeviceToHost, stream_));
        if (nimp) AVN_CUDA(cudaMemcpyAsync(nimp, nimp_out_.p, size_t(E_) * 4 * sizeof(S), cudaMemcpyDeviceToHost, stream_));
        AVN_CUDA(cudaStreamSynchronize(stream_));
        return AVN_OK;
    }

   private:
    EdgeRows<S> rows() {
        EdgeRows<S> r{};
        r.E = int(E_);
        r.c1 = c1_.as<uint32_t>(); r.c2 = c2_.as<uint32_t>(); r.b1 = b1_.as<uint32_t>(); r.b2 = b2_.as<uint32_t>(); r.live = live_.as<uint8_t>();
        r.count = count_.as<uint8_t>(); r.disjoint = disjoint_.as<uint8_t>(); r.normal = normal_.as<S>(); r.a1 = a1_.as<S>(); r.a2 = a2_.as<S>();
        r.pen = pen_.as<S>(); r.ns = ns_.as<S>(); r.prev_count = prev_count_.as<uint8_t>(); r.prev_a1 = prev_a1_.as<double>(); r.prev_a2 = prev_a2_.as<double>();
        r.ws_n_in = ws_n_in_.as<S>(); r.ws_t_in = ws_t_in_.as<S>(); r.ws_n_out = ws_n_out_.as<S>(); r.ws_t_out = ws_t_out_.as<S>();
        r.nimp_in = nimp_in_.as<S>(); r.nimp_out = nimp_out_.as<S>();
        return r;
    }
    template <class T> AvnStatus up(DevBuf& buf, const void* host, size_t count, const T** dev) {
        *dev = nullptr;
        if (!host || count == 0) return AVN_OK;
        AVN_CUDA(buf.ensure(count * sizeof(T)));
        AVN_CUDA(cudaMemcpyAsync(buf.p, host, count * sizeof(T), cudaMemcpyHostToDevice, stream_));
        *dev = buf.as<T>();
        return AVN_OK;
    }
    cudaStream_t stream_;
    ErrorSink* err_;
    uint32_t E_ = 0;
    DevBuf c1_, c2_, b1_, b2_, live_, count_, disjoint_, normal_, a1_, a2_, pen_, ns_, prev_count_, prev_a1_, prev_a2_, ws_n_in_, ws_t_in_, ws_n_out_, ws_t_out_,
        nimp_in_, nimp_out_, stage_;
    DevBuf i_shape_, i_dims_, i_pos_, i_rot_, i_lv_, i_av_, i_amin_, i_amax_;
};

}  // namespace

ContactsBase* make_contacts(uint32_t scalar_bits, cudaStream_t stream, ErrorSink* err) {
    if (scalar_bits == 32) return new Contacts<float>(stream, err);
    if (scalar_bits == 64) return new Contacts<double>(stream, err);
    return nullptr;
}

}  // namespace avn
