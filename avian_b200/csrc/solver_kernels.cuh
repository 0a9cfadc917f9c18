// Kernels of the solver stage.
//
// Two launch strategies over the SAME per-item device routines (solver_dev.cuh / joints_dev.cuh):
//   * step_megakernel: ONE persistent cooperative kernel per physics step.  The grid is sized to exactly fill
//     the 148 SMs (occupancy x SM count); every phase of the step (prepare, each graph colour of each pass of each
//     substep, each joint level, finalize) is a grid-stride loop followed by a grid-wide barrier.  A 100k-cube
//     step has ~300-600 dependent phases of only 10^4..10^5 independent items each, so the step is bound by
//     phase latency; removing ~500 kernel launches and keeping the working set hot in the 126 MB L2 between phases
//     is what the B200 wants.
//   * phase kernels: one launch per phase; the same arithmetic, used for profiling single phases under ncu, for
//     the roofline measurement of the solver-iteration kernel, and as the fallback when a cooperative launch is
//     refused.
#pragma once
#include <cooperative_groups.h>

#include "joints_dev.cuh"

namespace avn {
namespace cg = cooperative_groups;

constexpr int MEGA_BLOCK = 128, MEGA_BLOCKS_PER_SM = 3;

enum PhaseOp {
    OP_PREPARE_BODY = 0, OP_PREPARE_CONSTRAINT, OP_PREPARE_JOINT, OP_INTEGRATE_VEL, OP_INTEGRATE_POS, OP_WARM, OP_SOLVE_BIAS,
    OP_RELAX, OP_RESTITUTION, OP_SOLVE_JOINT, OP_PROJECT_VEL, OP_DAMP_JOINT, OP_WRITEBACK_BODY, OP_STORE_IMPULSE, OP_JOINT_FORCE
};

template <class S, int OP>
__device__ __forceinline__ void run_item(const DevSolver<S>& d, int i) {
    if (OP == OP_PREPARE_BODY) prepare_body_item(d, i);
    else if (OP == OP_PREPARE_CONSTRAINT) prepare_constraint_item(d, i);
    else if (OP == OP_PREPARE_JOINT) prepare_joint_item(d, i);
    else if (OP == OP_INTEGRATE_VEL) integrate_velocity_item(d, i);
    else if (OP == OP_INTEGRATE_POS) { integrate_position_item(d, i); if (d.J > 0) store_pre_solve_item(d, i); }
    else if (OP == OP_WARM) contact_item<S, PASS_WARM>(d, i);
    else if (OP == OP_SOLVE_BIAS) contact_item<S, PASS_SOLVE_BIAS>(d, i);
    else if (OP == OP_RELAX) contact_item<S, PASS_RELAX>(d, i);
    else if (OP == OP_RESTITUTION) contact_item<S, PASS_RESTITUTION>(d, i);
    else if (OP == OP_SOLVE_JOINT) solve_joint_item(d, i);
    else if (OP == OP_PROJECT_VEL) project_velocity_item(d, i);
    else if (OP == OP_DAMP_JOINT) damp_joint_item(d, i);
    else if (OP == OP_WRITEBACK_BODY) writeback_body_item(d, i);
    else if (OP == OP_STORE_IMPULSE) store_impulse_item(d, i);
    else if (OP == OP_JOINT_FORCE) joint_force_item(d, i);
}

// one launch per phase: items [begin, begin+count).  `serial` = the overflow colour: one thread, list order.
template <class S, int OP>
__global__ void __launch_bounds__(256) phase_kernel(const __grid_constant__ DevSolver<S> d, int begin, int count, int serial) {
    if (serial) {
        if (blockIdx.x == 0 && threadIdx.x == 0)
            for (int i = 0; i < count; ++i) run_item<S, OP>(d, begin + i);
        return;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) run_item<S, OP>(d, begin + i);
}

// __noinline__: each phase keeps its own register allocation instead of the union of all phases
template <class S, int OP>
__device__ __noinline__ void grid_phase(const DevSolver<S>& d, int begin, int count) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) run_item<S, OP>(d, begin + i);
}

template <class S, int OP>
__device__ __noinline__ void grid_serial(const DevSolver<S>& d, int begin, int count) {
    for (int i = 0; i < count; ++i) run_item<S, OP>(d, begin + i);
}

// all graph colours of one contact pass, reference order: overflow colour serially first, then colours 0..22
// (solver/plugin.rs:461-479, 553-572, 643-668)
template <class S, int OP>
__device__ __forceinline__ void grid_contact_pass(const DevSolver<S>& d, cg::grid_group& grid) {
    const int ov = d.color_off[AVN_COLOR_OVERFLOW], ovn = d.color_off[AVN_COLOR_OVERFLOW + 1] - ov;
    if (ovn > 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) grid_serial<S, OP>(d, ov, ovn);
        grid.sync();
    }
    for (int c = 0; c < AVN_COLOR_OVERFLOW; ++c) {
        const int b = d.color_off[c], n = d.color_off[c + 1] - b;
        if (n <= 0) continue;
        grid_phase<S, OP>(d, b, n);
        grid.sync();
    }
}

template <class S>
__global__ void __launch_bounds__(MEGA_BLOCK, MEGA_BLOCKS_PER_SM) step_megakernel(const __grid_constant__ DevSolver<S> d) {
    cg::grid_group grid = cg::this_grid();
    // ---- prepare
    grid_phase<S, OP_PREPARE_BODY>(d, 0, d.B + 1);
    grid.sync();
    grid_phase<S, OP_PREPARE_CONSTRAINT>(d, 0, d.M);
    grid_phase<S, OP_PREPARE_JOINT>(d, 0, d.J);
    grid.sync();
    // ---- run_substep_schedule (solver/schedule.rs:194-213)
    for (int sub = 0; sub < d.substeps; ++sub) {
        grid_phase<S, OP_INTEGRATE_VEL>(d, 0, d.B);
        grid.sync();
        if (d.M > 0) {
            grid_contact_pass<S, OP_WARM>(d, grid);
            for (int it = 0; it < d.iters; ++it) grid_contact_pass<S, OP_SOLVE_BIAS>(d, grid);
        }
        grid_phase<S, OP_INTEGRATE_POS>(d, 0, d.B);
        grid.sync();
        if (d.M > 0) grid_contact_pass<S, OP_RELAX>(d, grid);
        if (d.J > 0) {
            for (int l = 0; l < d.n_levels; ++l) {
                const int b = d.level_off[l], n = d.level_off[l + 1] - b;
                grid_phase<S, OP_SOLVE_JOINT>(d, b, n);
                grid.sync();
            }
            grid_phase<S, OP_PROJECT_VEL>(d, 0, d.B);
            grid.sync();
            if (d.any_joint_damping) {
                for (int l = 0; l < d.n_levels; ++l) {
                    const int b = d.level_off[l], n = d.level_off[l + 1] - b;
                    grid_phase<S, OP_DAMP_JOINT>(d, b, n);
                    grid.sync();
                }
            }
        }
    }
    // ---- restitution, writeback, store impulses
    if (d.M > 0 && *d.any_restitution) grid_contact_pass<S, OP_RESTITUTION>(d, grid);
    grid_phase<S, OP_WRITEBACK_BODY>(d, 0, d.B);
    grid_phase<S, OP_STORE_IMPULSE>(d, 0, d.M);
    grid_phase<S, OP_JOINT_FORCE>(d, 0, d.J);
}

}  // namespace avn
