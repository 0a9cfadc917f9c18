/*
 * avian_b200.h — C ABI of libavian_b200.so, the B200-native replacement for the avian3d substep hot path.
 *
 * The reference (avianphysics/avian @ 5bef382) has no FFI: its hot path is three Bevy plugins
 * (`IntegratorPlugin`, `BroadPhasePlugin`, `SolverPlugin` + `XpbdSolverPlugin`).  A thin Rust shim
 *  snapshots the ECS component columns those plugins read into the column structs
 * below once per physics step, calls the entry points here, and scatters the results back.
 *
 * Each entry point cites the reference system(s) it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns an AvnStatus (0 = ok, < 0 = error); nothing throws or aborts across the ABI;
 *     avn_last_error() returns a human-readable message for the last failing call on that context.
 *   - the caller owns all host buffers; they must stay valid until the call returns.  Buffers obtained
 *     from avn_alloc_pinned() are page-locked, which makes the host<->device copies asynchronous DMA.
 *   - the library owns all device memory.  One call in flight per context; a context is not thread-safe,
 *     but any host thread may call (the context binds its CUDA device on entry).
 *   - "scalar" columns are `float` when the context was created with scalar_bits = 32 and `double` when
 *     scalar_bits = 64 (reference features `f32` / `f64`, crates/avian3d/Cargo.toml:14-77).
 *   - Vec3 columns are packed [n][3], quaternions are packed [n][4] in glam order x,y,z,w, symmetric 3x3
 *     matrices are packed [n][6] = m00,m01,m02,m11,m12,m22 (glam_matrix_extras::SymmetricMat3).
 *   - there is NO CPU fallback: if no CUDA device is usable avn_create() fails with AVN_ERR_CUDA.
 */
#ifndef AVIAN_B200_H
#define AVIAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVN_ABI_VERSION 1u

/* src/dynamics/solver/constraint_graph.rs:39-48 */
#define AVN_GRAPH_COLOR_COUNT 24
#define AVN_COLOR_OVERFLOW 23
#define AVN_DYNAMIC_COLOR_COUNT 20
/* ContactManifold::prune_points keeps at most 4 points in 3D (src/collision/contact_types/mod.rs:478-566). */
#define AVN_MAX_MANIFOLD_POINTS 4
/* A contact/joint body that has no solver body and no column entry (static, zero velocity). */
#define AVN_NO_BODY (-1)

typedef enum AvnStatus {
    AVN_OK = 0,
    AVN_ERR_INVALID_ARGUMENT = -1,
    AVN_ERR_CUDA = -2,
    AVN_ERR_OUT_OF_MEMORY = -3,
    AVN_ERR_UNSUPPORTED = -4,
    AVN_ERR_CAPACITY = -5,
    AVN_ERR_NCCL = -6
} AvnStatus;

/* RigidBody (src/dynamics/rigid_body/mod.rs). Static entries are optional: they carry pose for joints and
 * LinearVelocity for contact tangents but are SolverBody::DUMMY inside the solver (solver_body/mod.rs:93-104). */
typedef enum AvnBodyKind { AVN_BODY_DYNAMIC = 0, AVN_BODY_KINEMATIC = 1, AVN_BODY_STATIC = 2 } AvnBodyKind;

/* LockedAxes bit layout (src/dynamics/rigid_body/locked_axes.rs; same bits as SolverBodyFlags, solver_body/mod.rs:133-146). */
#define AVN_LOCK_TRANSLATION_X 0x20u
#define AVN_LOCK_TRANSLATION_Y 0x10u
#define AVN_LOCK_TRANSLATION_Z 0x08u
#define AVN_LOCK_ROTATION_X 0x04u
#define AVN_LOCK_ROTATION_Y 0x02u
#define AVN_LOCK_ROTATION_Z 0x01u

/* integration_flags bits: marker components integrator/mod.rs:168-195 */
#define AVN_CUSTOM_VELOCITY_INTEGRATION 0x1u
#define AVN_CUSTOM_POSITION_INTEGRATION 0x2u

/* AabbIntervalFlags, src/collision/broad_phase.rs:187-196 */
#define AVN_AABB_IS_INACTIVE 0x01u
#define AVN_AABB_CONTACT_EVENTS 0x02u
#define AVN_AABB_GENERATE_CONSTRAINTS 0x04u
#define AVN_AABB_CUSTOM_FILTER 0x08u
#define AVN_AABB_MODIFY_CONTACTS 0x10u
/* Multi-GPU x-slab partition: the interval is a copy of one owned by the next slab ("halo").  It is only ever the LATER
 * element of a pair: the sweep never starts from it, so pairs between two halo intervals are left to the slab that owns them.  Not a
 * reference flag; single-GPU callers never set it. */
#define AVN_AABB_HALO 0x80u
/* An interval that reaches far beyond its own slab (a ground slab) is not swept by its owner alone: it carries AVN_AABB_SPLIT_I in
 * every slab it reaches and pairs, as the earlier element, only with intervals that slab owns (never with AVN_AABB_HALO ones).  In
 * the slabs that do not own it it also carries AVN_AABB_NOT_J: it is never the later element there. */
#define AVN_AABB_SPLIT_I 0x40u
#define AVN_AABB_NOT_J 0x20u

/* Flag bits of an emitted pair (what collect_collision_pairs stores on ContactEdge / ContactPair,
 * broad_phase.rs:443-468) plus NEEDS_HOOK: the shim must still call CollisionHooks::filter_pairs for it
 * (broad_phase.rs:431-439), in list order, and drop the pair when the hook says no. */
#define AVN_PAIR_CONTACT_EVENTS 0x01u
#define AVN_PAIR_MODIFY_CONTACTS 0x02u
#define AVN_PAIR_GENERATE_CONSTRAINTS 0x04u
#define AVN_PAIR_NEEDS_HOOK 0x08u

typedef struct AvnContext AvnContext;

typedef struct AvnConfig {
    uint32_t abi_version;     /* AVN_ABI_VERSION */
    int32_t device;           /* CUDA device ordinal */
    uint32_t scalar_bits;     /* 32 (feature f32) or 64 (feature f64) */
    uint32_t flags;           /* AVN_CFG_* */
} AvnConfig;

/* Evaluate sin/cos of Quat::from_scaled_axis in double and round (bit-compatible with a correctly rounded
 * libm; default).  Without it the f32 path uses the CUDA sinf/cosf (<= 1 ulp). */
#define AVN_CFG_FAST_TRIG 0x1u

/* ---- step parameters: resources read by the three plugins -------------------------------------------- */
typedef struct AvnStepParams {
    double dt;                /* Time<Physics>::delta_secs_f64()  (src/schedule/mod.rs:247-260) */
    double h;                 /* Time<Substeps>::delta_secs_f64() (solver/schedule.rs:195-200)  */
    uint32_t substeps;        /* SubstepCount (solver/schedule.rs:187-191) */
    uint32_t restitution_iterations; /* SolverConfig (solver/plugin.rs:291-302) */
    double gravity[3];        /* Gravity (integrator/mod.rs:150-162) */
    double contact_damping_ratio;
    double contact_frequency_factor;
    double max_overlap_solve_speed;
    double warm_start_coefficient;
    double restitution_threshold;
    double length_unit;       /* PhysicsLengthUnit (solver/plugin.rs:200-207) */
    uint32_t match_contacts;  /* NarrowPhaseConfig::match_contacts: warm starting enabled (plugin.rs:432) */
    uint32_t solver_iterations; /* EXTENSION, reference semantics = 1: repeats the biased solve pass */
} AvnStepParams;

/* ---- bodies: Appendix B of SURVEY.md; queries at solver_body/plugin.rs:174-185, integrator/mod.rs:261-268 */
typedef struct AvnBodyColumns {
    uint32_t count;
    uint32_t _pad;
    const uint8_t* kind;              /* AvnBodyKind */
    void* position;                   /* [n][3] in/out  Position */
    void* rotation;                   /* [n][4] in/out  Rotation */
    void* linear_velocity;            /* [n][3] in/out  LinearVelocity */
    void* angular_velocity;           /* [n][3] in/out  AngularVelocity */
    const void* inverse_mass;         /* [n]     ComputedMass::inverse() */
    const void* inverse_inertia_local;/* [n][6]  ComputedAngularInertia::inverse() (local frame) */
    const void* center_of_mass;       /* [n][3]  ComputedCenterOfMass (local); NULL = zero */
    const uint8_t* locked_axes;       /* NULL = none */
    const int8_t* dominance;          /* NULL = 0 */
    const void* linear_damping;       /* [n] NULL = 0 */
    const void* angular_damping;      /* [n] NULL = 0 */
    const void* gravity_scale;        /* [n] NULL = 1 */
    const void* linear_acceleration;  /* [n][3] VelocityIntegrationData::linear_increment as written by ForcePlugin
                                         (an acceleration until UpdateVelocityIncrements); NULL = 0 */
    const void* angular_acceleration; /* [n][3] NULL = 0 */
    const void* max_linear_speed;     /* [n] MaxLinearSpeed; NULL or +inf = absent */
    const void* max_angular_speed;    /* [n] MaxAngularSpeed; NULL or +inf = absent */
    const uint8_t* integration_flags; /* NULL = 0 */
} AvnBodyColumns;

/* ---- contact manifolds, grouped by graph colour (ConstraintGraph.colors[c].manifold_handles order,
 *      solver/plugin.rs:389-434; ContactManifold/ContactPoint at contact_types/mod.rs:342-378,603-660) ---- */
typedef struct AvnManifoldColumns {
    uint32_t count;                                     /* M */
    uint32_t point_count;                               /* P = point_offsets[M] */
    uint32_t color_offsets[AVN_GRAPH_COLOR_COUNT + 1];  /* colour c owns manifolds [off[c], off[c+1]) */
    const int32_t* body1;             /* [M] index into AvnBodyColumns or AVN_NO_BODY */
    const int32_t* body2;
    const void* normal;               /* [M][3] */
    const void* friction;             /* [M] */
    const void* restitution;          /* [M] */
    const void* tangent_velocity;     /* [M][3] NULL = 0 */
    const uint32_t* point_offsets;    /* [M+1], at most AVN_MAX_MANIFOLD_POINTS per manifold */
    const void* anchor1;              /* [P][3] */
    const void* anchor2;              /* [P][3] */
    const void* penetration;          /* [P] */
    const void* normal_speed;         /* [P] */
    void* warm_start_normal_impulse;  /* [P]    in/out (store_contact_impulses, plugin.rs:741-750) */
    void* warm_start_tangent_impulse; /* [P][2] in/out */
    void* normal_impulse;             /* [P]    out: ContactPoint::normal_impulse (total) */
} AvnManifoldColumns;

/* ---- joints: five typed arrays in the reference's solve order (xpbd/plugin.rs:58-86), each in ECS table
 *      order.  Frames are the already-localised ones (update_local_frames in joints/{fixed,revolute,...}.rs). ------------------ */
typedef enum AvnJointType {
    AVN_JOINT_FIXED = 0,
    AVN_JOINT_REVOLUTE = 1,
    AVN_JOINT_SPHERICAL = 2,
    AVN_JOINT_PRISMATIC = 3,
    AVN_JOINT_DISTANCE = 4,
    AVN_JOINT_TYPE_COUNT = 5
} AvnJointType;

typedef struct AvnJointColumns {
    uint32_t count;
    uint32_t _pad;
    const int32_t* body1;             /* [J] index into AvnBodyColumns (static bodies need an entry: pose is read) */
    const int32_t* body2;
    const void* local_anchor1;        /* [J][3] */
    const void* local_anchor2;        /* [J][3] */
    const void* local_basis1;         /* [J][4] (unused by Distance); NULL = identity */
    const void* local_basis2;
    const void* axis;                 /* [J][3] hinge_axis / twist_axis / slider_axis; NULL = type default (Z / Y / X) */
    const uint8_t* limit_enabled;     /* [J] bit0: angle_limit | swing_limit | prismatic limits ; bit1: twist_limit.
                                         Distance joints always use limit (min,max). NULL = 0 */
    const void* limit_min;            /* [J] */
    const void* limit_max;            /* [J] */
    const void* limit2_min;           /* [J] spherical twist_limit */
    const void* limit2_max;
    /* compliances, NULL = 0.  meaning per type:
     *   fixed:     c0 point      c1 angle
     *   revolute:  c0 point      c1 align     c2 limit
     *   spherical: c0 point      c1 swing     c2 twist
     *   prismatic: c0 align(pos) c1 angle     c2 limit (unused by the reference solve, kept for layout)
     *   distance:  c0 compliance */
    const void* compliance0;
    const void* compliance1;
    const void* compliance2;
    const uint8_t* damping_enabled;   /* [J] JointDamping present; NULL = none */
    const void* damping_linear;       /* [J] */
    const void* damping_angular;      /* [J] */
    void* force;                      /* [J][3] out: JointForces::force  (xpbd/plugin.rs:242-260); NULL = skip */
    void* torque;                     /* [J][3] out */
} AvnJointColumns;

typedef struct AvnJointSet {
    AvnJointColumns types[AVN_JOINT_TYPE_COUNT];
} AvnJointSet;

/* ---- broad phase: AabbIntervals (broad_phase.rs:176-202), in the PERSISTENT interval order (previous
 *      frame's sorted order, then newly added colliders appended, broad_phase.rs:296-315) -------------- */
typedef struct AvnAabbColumns {
    uint32_t count;                   /* C */
    uint32_t retained_count;          /* out (written by upload and again by download): intervals that stay in the list = entries of order_out.
                                         Intervals whose AABB has a NaN or infinite component are dropped, as update_aabb_intervals' retain
                                         does (broad_phase.rs:243-245): they form no pairs and are absent from order_out */
    const uint32_t* collider;         /* [C] Entity::index() of the collider */
    const uint32_t* body;             /* [C] Entity::index() of ColliderOf::body */
    const void* aabb_min;             /* [C][3] ColliderAabb::min */
    const void* aabb_max;             /* [C][3] */
    const uint32_t* memberships;      /* [C] CollisionLayers; NULL = 1 (default layer) */
    const uint32_t* filters;          /* [C] NULL = 0xFFFFFFFF */
    const uint8_t* flags;             /* [C] AVN_AABB_* */
    uint32_t* order_out;              /* [C] out: new persistent order, as indices into these columns (retained_count entries); NULL = skip */
    /* pairs already in ContactGraph::pair_set (contact_graph.rs:95): PairKey u64, any order */
    const uint64_t* existing_pairs;
    uint64_t existing_pair_count;
    /* body pairs (PairKey of body Entity::index()) whose joints disable collision (broad_phase.rs:423-428) */
    const uint64_t* joint_disabled_body_pairs;
    uint64_t joint_disabled_pair_count;
} AvnAabbColumns;

typedef struct AvnPairList {
    uint64_t capacity;                /* in: elements available in each array below */
    uint64_t count;                   /* out: pairs found (may exceed capacity -> AVN_ERR_CAPACITY, arrays hold the prefix) */
    uint32_t* collider1;              /* [capacity] out: Entity::index() (i before j in the sorted order) */
    uint32_t* collider2;
    uint32_t* body1;
    uint32_t* body2;
    uint8_t* flags;                   /* AVN_PAIR_* */
} AvnPairList;

/* Device-time per phase in milliseconds of the last avn_solver_step / avn_broadphase call, named after
 * SolverDiagnostics (src/dynamics/solver/diagnostics.rs:13-39) and CollisionDiagnostics (collision/diagnostics.rs:13-20). */
typedef struct AvnTimings {
    float h2d_ms;
    float prepare_ms;          /* prepare_solver_bodies + prepare_joints + prepare_constraints + update_velocity_increments */
    float substep_loop_ms;     /* integrate_velocities .. joint damping, all substeps */
    float finalize_ms;         /* apply_restitution + finalize + store_impulses */
    float d2h_ms;
    float broad_phase_ms;
    float total_ms;
    uint32_t kernel_launches;  /* kernels launched by the last call */
    uint32_t contact_constraint_count;
    uint32_t joint_levels;
    uint32_t active_colors;
    uint32_t launch_mode;      /* AVN_LAUNCH_*: how the last solver stage was launched (a refused cooperative launch falls back to phases) */
} AvnTimings;
#define AVN_LAUNCH_PHASES 0u        /* one kernel launch per phase (~500 per step) */
#define AVN_LAUNCH_MEGA_BARRIER 1u  /* one persistent cooperative kernel, grid barriers between colours */
#define AVN_LAUNCH_MEGA_WAVE 2u     /* one persistent cooperative kernel, per-body event counters instead of barriers */
#define AVN_LAUNCH_MEGA_ISLANDS 3u  /* one persistent cooperative kernel, one warp per simulation island through the whole substep loop (many small islands) */

/* lifecycle ------------------------------------------------------------------------------------------- */
AvnStatus avn_create(const AvnConfig* config, AvnContext** out_ctx);
void avn_destroy(AvnContext* ctx);
const char* avn_last_error(const AvnContext* ctx); /* ctx may be NULL: error of the last failed avn_create */
uint32_t avn_abi_version(void);

/* pinned host memory for the column buffers */
AvnStatus avn_alloc_pinned(AvnContext* ctx, size_t bytes, void** out_ptr);
AvnStatus avn_free_pinned(AvnContext* ctx, void* ptr);

/*
 * One full solver stage of a physics step.  Replaces, in order:
 *   prepare_solver_bodies (solver_body/plugin.rs:173-251), prepare_xpbd_joint<T> x5 (xpbd/plugin.rs:125-142),
 *   update_contact_softness + prepare_contact_constraints (solver/plugin.rs:326-448),
 *   pre_process_velocity_increments (integrator/mod.rs:260-313),
 *   run_substep_schedule (solver/schedule.rs:194-213): integrate_velocities, clamp_velocities, warm_start,
 *     solve_contacts<true>, integrate_positions, solve_contacts<false>, solve_xpbd_joint<T> x5,
 *     project_linear/angular_velocity, joint_damping<T> x5,
 *   solve_restitution (solver/plugin.rs:630-718), writeback_solver_bodies (solver_body/plugin.rs:255-284),
 *   writeback_joint_forces (xpbd/plugin.rs:242-260), store_contact_impulses (solver/plugin.rs:722-755).
 * Host buffers in, host buffers out (copies are inside the call).  manifolds / joints may be NULL.
 */
AvnStatus avn_solver_step(AvnContext* ctx, const AvnStepParams* params, AvnBodyColumns* bodies,
                          AvnManifoldColumns* manifolds, AvnJointSet* joints);

/* The same stage split in three so a caller can keep the snapshot resident in HBM:
 *   upload (H2D only) -> run (kernels only, repeatable: every run restarts from the uploaded snapshot)
 *   -> download (D2H of the last run's results into the column buffers given to upload). */
AvnStatus avn_solver_upload(AvnContext* ctx, const AvnStepParams* params, AvnBodyColumns* bodies,
                            AvnManifoldColumns* manifolds, AvnJointSet* joints);
AvnStatus avn_solver_run(AvnContext* ctx);
AvnStatus avn_solver_download(AvnContext* ctx);

/* The same stage fed from EDGE-INDEXED manifold storage: the layout avn_narrow_phase writes (4 point slots per contact edge, indexed by
 * ContactId) plus the constraint graph as a colour-major list of edge ids.  prepare_contact_constraints reads manifold m through
 * edge[m]; store_contact_impulses writes the impulses back to the edge's slots.  This is the input form of a device-resident pipeline:
 * the geometry never has to be compacted or sent through the host, only the edge list changes hands. */
typedef struct AvnEdgeManifolds {
    uint32_t count;                                     /* M: manifolds in the constraint graph */
    uint32_t edge_capacity;                             /* E: rows of the edge-indexed columns */
    uint32_t color_offsets[AVN_GRAPH_COLOR_COUNT + 1];  /* colour c owns edge[off[c], off[c+1]) */
    const uint32_t* edge;             /* [M] ContactId of manifold m */
    const int32_t* body1;             /* [M] */
    const int32_t* body2;
    const void* friction;             /* [M] */
    const void* restitution;          /* [M] */
    const uint8_t* point_count;       /* [E] 0..4 */
    const void* normal;               /* [E][3] */
    const void* anchor1;              /* [E][4][3] */
    const void* anchor2;
    const void* penetration;          /* [E][4] */
    const void* normal_speed;         /* [E][4] */
    void* warm_start_normal_impulse;  /* [E][4]    in/out */
    void* warm_start_tangent_impulse; /* [E][4][2] in/out */
    void* normal_impulse;             /* [E][4]    out */
} AvnEdgeManifolds;
AvnStatus avn_solver_upload_edges(AvnContext* ctx, const AvnStepParams* params, AvnBodyColumns* bodies, AvnEdgeManifolds* manifolds, AvnJointSet* joints);

/* ---- one coupled scene over several GPUs: the x-slab partition (BASELINE north_star "single all-gather of boundary
 *      state per substep where the scene spans GPUs").  Not a reference interface: the reference is single-process. -------------
 * Each rank uploads its own bodies and constraints plus copies ("ghosts") of the remote bodies its constraints touch.  A body held
 * by more than one rank is a BOUNDARY body; it has one slot in a table every rank agrees on.  Per substep each rank launches
 * avn_solver_run_range for that substep, packs for every boundary body it holds the velocity change its own constraints caused
 * (relative to the velocity right after integrate_velocities, which every holder computes identically) and, if it owns the body,
 * the body's delta_position / delta_rotation, into its own packed table (one record per held boundary body); the tables are all-gathered (NCCL, by the caller, on the stream avn_get_stream
 * returns); avn_solver_boundary_apply sets v = v_ref + sum over ranks in rank order of their changes and takes the owner's deltas.
 * Impulses therefore cross a cut once per substep instead of once per constraint: results match the single-GPU step to solver
 * tolerance, not to 1e-5; a scene whose constraints do not cross a cut is reproduced bit for bit. */
#define AVN_RUN_PREPARE 0x1u      /* prepare bodies / constraints: must be part of the first launch after an upload */
#define AVN_RUN_RESTITUTION 0x2u  /* solve_restitution after the substeps of this launch */
#define AVN_RUN_FINALIZE 0x4u     /* writeback_solver_bodies + store_contact_impulses: must be part of the last launch */

typedef struct AvnBoundary {
    uint32_t count;               /* boundary bodies held by this rank: record k of this rank's table belongs to body[k] */
    uint32_t record_count;        /* records per rank's table = the largest `count` of any rank (all-gather needs equal sizes) */
    uint32_t rank, world;
    const int32_t* body;          /* [count] index into the uploaded AvnBodyColumns */
    const int32_t* source;        /* [count][world] record of body[k] in rank r's table, or -1 when rank r does not hold it */
    const int32_t* owner_rank;    /* [count] the rank whose delta_position / delta_rotation are authoritative */
} AvnBoundary;
/* scalars per record of the exchange table (4 rows of 4): a table is record_count * AVN_BOUNDARY_RECORD_SCALARS scalars,
 * the gathered tables world times that, rank-major.  Only held bodies travel: the table is as large as the busiest rank's list. */
#define AVN_BOUNDARY_RECORD_SCALARS 16

AvnStatus avn_solver_run_range(AvnContext* ctx, uint32_t first_substep, uint32_t substep_count, uint32_t run_flags);
AvnStatus avn_solver_set_boundary(AvnContext* ctx, const AvnBoundary* boundary);   /* after upload; NULL / count 0 clears it */
AvnStatus avn_solver_boundary_snapshot(AvnContext* ctx);                           /* v_ref = current velocity (before a restitution launch) */
AvnStatus avn_solver_boundary_pack(AvnContext* ctx, void* device_table);           /* DEVICE pointer */
AvnStatus avn_solver_boundary_apply(AvnContext* ctx, const void* device_gathered); /* DEVICE pointer */
AvnStatus avn_solver_needs_restitution(AvnContext* ctx, int* out_nonzero);         /* any uploaded restitution coefficient != 0 */
/* the context's CUDA stream (a cudaStream_t) so that the caller's collective can be ordered with the launches above */
AvnStatus avn_get_stream(AvnContext* ctx, void** out_stream);

/* ---- the communicator: one process per GPU, one AvnContext per process, one NCCL communicator per context.  NCCL is bound
 *      at run time (dlopen "libnccl.so.2", or the name in AVN_NCCL_LIB): a host that never calls avn_comm_init with world > 1 needs no NCCL.
 *      One rank calls avn_comm_unique_id and hands the AVN_COMM_ID_BYTES bytes to the others over any channel the host has (a file, a socket,
 *      MPI, torch.distributed's store); then EVERY rank calls avn_comm_init with the same bytes (collective).  Errors: AVN_ERR_NCCL. */
#define AVN_COMM_ID_BYTES 128
AvnStatus avn_comm_unique_id(AvnContext* ctx, void* out_id /* [AVN_COMM_ID_BYTES] */);
AvnStatus avn_comm_init(AvnContext* ctx, uint32_t rank, uint32_t world, const void* unique_id /* NULL allowed when world == 1 */);
AvnStatus avn_comm_destroy(AvnContext* ctx);
/* all-gather of bytes_per_rank bytes of DEVICE memory from every rank into recv_device (world * bytes_per_rank, rank-major), enqueued on the
 * context's stream (pair lists of the slab broad phase, results of the owned rows) */
AvnStatus avn_comm_all_gather(AvnContext* ctx, const void* send_device, void* recv_device, size_t bytes_per_rank);
/*
 * The whole partitioned solver stage of this rank, inside the library: after avn_solver_upload (this rank's share) and
 * avn_solver_set_boundary, runs for every substep  avn_solver_run_range -> pack -> ncclAllGather of the packed tables -> apply  on the
 * context's stream, the restitution launch + one more exchange when any rank uploaded a non-zero coefficient (agreed with one all-reduce),
 * and the finalize launch.  Follow with avn_solver_download.  With a communicator of one rank (or no boundary) it equals avn_solver_run.
 * The exchange tables live in the library; nothing crosses the host.
 */
AvnStatus avn_solver_step_partitioned(AvnContext* ctx);

/*
 * Sweep-and-prune pair generation.  Replaces collect_collision_pairs / sweep_and_prune (broad_phase.rs:343-474)
 * over the intervals that update_aabb_intervals / add_new_aabb_intervals maintain (broad_phase.rs:214-315).
 * The emitted list is exactly the sequence of ContactGraph::add_edge_and_key_with calls the reference makes
 * (same pairs, same order); pairs flagged AVN_PAIR_NEEDS_HOOK still need the host filter_pairs callback.
 */
AvnStatus avn_broadphase(AvnContext* ctx, AvnAabbColumns* aabbs, AvnPairList* out_pairs);
/* split form: the AvnAabbColumns struct itself (not only its buffers) must stay valid until avn_broadphase_download returns */
AvnStatus avn_broadphase_upload(AvnContext* ctx, AvnAabbColumns* aabbs);
AvnStatus avn_broadphase_run(AvnContext* ctx);
AvnStatus avn_broadphase_download(AvnContext* ctx, AvnPairList* out_pairs);

/* ---- collider AABBs (SURVEY.md 8f "next #2"): update_aabb for the shapes the device knows ------------------------------- */
typedef enum AvnShape { AVN_SHAPE_CUBOID = 0, AVN_SHAPE_SPHERE = 1 } AvnShape;

typedef struct AvnAabbParams {
    double dt;                         /* Time::delta (full step, collider/backend.rs:536) */
    double contact_tolerance;          /* PhysicsLengthUnit * NarrowPhaseConfig::contact_tolerance (default 0.005) */
    double default_speculative_margin; /* PhysicsLengthUnit * NarrowPhaseConfig::default_speculative_margin (default Scalar::MAX: pass +inf) */
} AvnAabbParams;

typedef struct AvnColliderColumns {
    uint32_t count;
    uint32_t _pad;
    const uint8_t* shape;              /* [C] AvnShape */
    const void* dims;                  /* [C][3] cuboid half extents / sphere radius in [0] (scaled shape) */
    const void* position;              /* [C][3] collider Position */
    const void* rotation;              /* [C][4] collider Rotation */
    const void* linear_velocity;       /* [C][3] the velocity update_aabb uses: the collider's own LinearVelocity, or its body's velocity at
                                          the collider offset (backend.rs:560-580); NULL = 0 */
    const void* angular_velocity;      /* [C][3] NULL = 0 */
    const void* collision_margin;      /* [C] CollisionMargin; NULL = 0 */
    const void* speculative_margin;    /* [C] SpeculativeMargin (+inf for SweptCcd); NULL = the default */
    void* aabb_min;                    /* [C][3] out: ColliderAabb::min */
    void* aabb_max;                    /* [C][3] out */
} AvnColliderColumns;

/*
 * Replaces update_aabb::<Collider> (src/collision/collider/backend.rs:498-625) for cuboid and sphere colliders: swept AABB from the
 * current pose to the pose after dt (rotation advanced by Quat::from_scaled_axis + fast_renormalize, translation clamped to the
 * speculative margin), grown by contact_tolerance + collision margin.
 */
AvnStatus avn_update_aabbs(AvnContext* ctx, const AvnAabbParams* params, AvnColliderColumns* colliders);

/* ---- contact manifolds (SURVEY.md 8f "next #1", geometry stage): one manifold of at most 4 points per contact pair of cuboid / sphere
 *      colliders.  Stands where NarrowPhase::update calls contact_manifolds (narrow_phase/system_param.rs:437-830,
 *      collider/parry/contact_query.rs:156-261); the arithmetic is this repository's generator (csrc/narrow_math.hpp — parry3d is not
 *      vendored), shared with the host fixture.  Matching, the touching state machine and the constraint graph stay on the host. -------- */
typedef struct AvnNarrowParams {
    double dt;                         /* Time::delta (narrow_phase/mod.rs:289) */
    double contact_tolerance;          /* PhysicsLengthUnit * NarrowPhaseConfig::contact_tolerance */
} AvnNarrowParams;

typedef struct AvnNarrowInput {
    uint32_t pair_count, collider_count, body_count, _pad;
    const uint32_t* collider1;         /* [pairs] row of the collider columns below (ascending ContactId order is the caller's business) */
    const uint32_t* collider2;
    const uint32_t* body1;             /* [pairs] row of the body velocity columns */
    const uint32_t* body2;
    const uint8_t* shape;              /* [C] AvnShape; NULL = cuboid */
    const void* dims;                  /* [C][3] cuboid half extents / sphere radius in [0] */
    const void* position;              /* [C][3] collider Position (collider at the body origin, centre of mass at the origin) */
    const void* rotation;              /* [C][4] */
    const void* linear_velocity;       /* [B][3] */
    const void* angular_velocity;      /* [B][3] */
    const void* aabb_min;              /* [C][3] optional: pairs whose AABBs are disjoint are reported in `disjoint` and skipped */
    const void* aabb_max;
} AvnNarrowInput;

typedef struct AvnRawManifolds {      /* fixed stride: 4 point slots per pair, unused slots zero */
    uint8_t* point_count;              /* [pairs] 0..4 (0 = not touching within the speculative margin) */
    uint8_t* disjoint;                 /* [pairs] optional */
    void* normal;                      /* [pairs][3] from collider1 to collider2 */
    void* anchor1;                     /* [pairs][4][3] */
    void* anchor2;
    void* penetration;                 /* [pairs][4] */
    void* normal_speed;                /* [pairs][4] */
} AvnRawManifolds;

AvnStatus avn_narrow_phase(AvnContext* ctx, const AvnNarrowParams* params, const AvnNarrowInput* input, AvnRawManifolds* out);

/* ---- device-resident contact edges: the contact pairs, their manifolds and warm-start impulses stay on the device between steps; the host
 *      keeps the ContactGraph and the ConstraintGraph (contact_graph.rs, constraint_graph.rs) and exchanges a few bytes per edge with the
 *      device (protocol and its CPU specification: avian_b200/plugins.py ResidentWorld).  Row = ContactId. ---------------------------------- */
AvnStatus avn_contacts_reserve(AvnContext* ctx, uint32_t capacity);                  /* rows; grows, keeps the existing rows */
AvnStatus avn_contacts_add(AvnContext* ctx, uint32_t n, const uint32_t* ids, const uint32_t* collider1, const uint32_t* collider2,
                           const uint32_t* body1, const uint32_t* body2);            /* ContactGraph::add_edge: the row starts without history */
AvnStatus avn_contacts_remove(AvnContext* ctx, uint32_t n, const uint32_t* ids);     /* ContactGraph::remove_edge */
/* Geometry + match_contacts for every live row (input: only the collider / body columns of AvnNarrowInput; the pair arrays are ignored).
 * out_point_count / out_disjoint: [capacity] host arrays — all the host needs for the touching state machine and the graphs. */
AvnStatus avn_contacts_narrow_phase(AvnContext* ctx, const AvnNarrowParams* params, const AvnNarrowInput* input, uint32_t match_contacts,
                                    double length_unit, uint8_t* out_point_count, uint8_t* out_disjoint);
/* The solver stage reading its manifolds from the resident rows: `graph` carries only count, color_offsets, edge, body1, body2, friction,
 * restitution (host); store_contact_impulses writes into the rows.  Then avn_solver_run / avn_solver_download (bodies only) as usual. */
/* graph->edge == NULL: the constraint graph has not changed since the previous avn_solver_upload_graph (no contact started or stopped
 * touching) — the list stays on the device and only the bodies are uploaded; count and color_offsets must repeat the previous call's. */
AvnStatus avn_solver_upload_graph(AvnContext* ctx, const AvnStepParams* params, AvnBodyColumns* bodies, const AvnEdgeManifolds* graph, AvnJointSet* joints);
/* the impulses of the rows as the last solve left them (tests, tools): [capacity][4], [capacity][4][2], [capacity][4]; any may be NULL */
AvnStatus avn_contacts_download_impulses(AvnContext* ctx, void* warm_start_normal, void* warm_start_tangent, void* normal_impulse);

/* ---- the ContactGraph and the ConstraintGraph on the device (SURVEY.md 8f "next #3"): after this nothing of the contact pipeline lives on
 *      the host.  Replaces, for the pairs the device narrow phase covers,
 *        ContactGraph::add_edge_and_key_with          (collision/contact_types/contact_graph.rs:521-565; ids: data_structures/id_pool.rs:43-52)
 *        the status-change loop of NarrowPhase::update (collision/narrow_phase/system_param.rs:136-389: removal of separated pairs,
 *                                                       started / stopped touching)
 *        ConstraintGraph::push_manifold / pop_manifold (dynamics/solver/constraint_graph.rs:163-296)
 *      with the reference's results: the same ContactId for every pair, the same colour for every manifold (the greedy colouring is
 *      order dependent; the device reproduces the sequential order with a dependency wavefront, csrc/contacts.cu).  The order of the
 *      manifolds INSIDE a colour is ascending ContactId instead of the reference's swap_remove order — it does not influence the solve;
 *      the overflow colour, which is solved serially, keeps the reference's list order. ------------------------------------------------ */
typedef struct AvnContactGraphConfig {
    uint32_t body_count, collider_count;
    const uint8_t* body_kind;          /* [B] AvnBodyKind: static bodies never enter a colour's body set */
    const double* friction;            /* [C] per collider; a pair's coefficient is the mean of its colliders' (NULL = 0.5) */
    const double* restitution;         /* [C] (NULL = 0) */
} AvnContactGraphConfig;
AvnStatus avn_contacts_configure(AvnContext* ctx, const AvnContactGraphConfig* config);

typedef struct AvnContactStep {
    uint32_t rows_high_water;          /* ContactIds in use: [0, rows_high_water) */
    uint32_t rows_live;                /* pairs in the ContactGraph after this step */
    uint32_t pairs_added, pairs_removed, started_touching, stopped_touching;
    uint32_t manifold_count;           /* manifolds in the ConstraintGraph = length of the colour-major list */
    uint32_t colouring_rounds;         /* dependency levels the greedy colouring needed this step */
    uint32_t any_restitution;
    uint32_t _pad;
    uint32_t color_offsets[AVN_GRAPH_COLOR_COUNT + 1];
} AvnContactStep;
#define AVN_CONTACTS_TAKE_BROADPHASE_PAIRS 0x1u   /* add the new pairs of the context's last avn_broadphase_run (read in device memory) */
#define AVN_CONTACTS_SHAPES_UNCHANGED 0x2u        /* input->shape and input->dims equal the previous call's: not copied again */
/* One step of the contact pipeline on the device: (new pairs ->) rows, geometry + match_contacts for every live row, the status loop, the
 * graphs, the colour-major list.  input: the collider / body columns of AvnNarrowInput (pair arrays ignored).  From the first call on the
 * contact store's pair set is the broad phase's "existing pairs" set (AvnAabbColumns::existing_pairs may stay NULL). */
AvnStatus avn_contacts_step(AvnContext* ctx, const AvnNarrowParams* params, const AvnNarrowInput* input, uint32_t match_contacts, double length_unit,
                            uint32_t flags, AvnContactStep* out);
/* The solver stage fed entirely from the device: manifolds from the contact rows, the constraint graph from the last avn_contacts_step.
 * Then avn_solver_run / avn_solver_download (bodies only) as usual. */
AvnStatus avn_solver_upload_resident(AvnContext* ctx, const AvnStepParams* params, AvnBodyColumns* bodies, AvnJointSet* joints);
/* Optional, any pipeline: start copying the body columns of the NEXT avn_solver_upload* / avn_solver_step call to the device now, on a second
 * stream, so that the copy overlaps the stages that run before the solver (broad phase, contact pipeline).  The next upload must be given the
 * same column pointers (otherwise it simply copies again); the host columns must not change in between.
 * AVN_BODIES_STATIC_UNCHANGED: the columns that describe the bodies (kind, locked_axes, dominance, integration_flags, inverse_mass,
 * inverse_inertia_local, center_of_mass, dampings, gravity_scale, max speeds) equal those of the previous upload of the same number of
 * bodies; only position, rotation, velocities and accelerations are copied. */
#define AVN_BODIES_STATIC_UNCHANGED 0x1u
AvnStatus avn_solver_prefetch_bodies(AvnContext* ctx, AvnBodyColumns* bodies, uint32_t flags);
/* After avn_broadphase_run in the device-resident pipeline: waits for the run, writes order_out / retained_count of the uploaded columns and
 * returns the number of new pairs; the pairs themselves stay on the device for avn_contacts_step. */
AvnStatus avn_broadphase_download_order(AvnContext* ctx, uint64_t* out_pair_count);
/* the graphs as the device holds them (tests, tools): per row [capacity] colliders, live / touching flags, colour (-1 = not in the
 * constraint graph); edge_list [manifold_count] = the colour-major list.  Any pointer may be NULL. */
AvnStatus avn_contacts_download_graph(AvnContext* ctx, uint32_t capacity, uint32_t* collider1, uint32_t* collider2, uint8_t* live, uint8_t* touching,
                                      int8_t* colour, uint32_t* edge_list);

/* ---- simulation islands and sleeping on the device (SURVEY.md 8f "next #4").  Replaces the bookkeeping and the decisions of
 *        PhysicsIslands::add_contact / remove_contact / add_joint / merge_islands / split_island   (dynamics/solver/islands/mod.rs:513-1270)
 *        update_sleeping_states, wake_islands_with_sleeping_disabled, sleep_islands                 (dynamics/solver/islands/sleeping.rs:164-292)
 *      Islands are PERSISTENT like the reference's: merged when a touching, constraint-generating contact (or a joint) links two of them,
 *      marked (constraints_removed) when such a contact goes, and split lazily — one island per step, the one holding the sleepiest body
 *      that wants to sleep — by recomputing its connected components.  The contact events come from the last avn_contacts_step of this
 *      context (the rows on the device); the body velocities are the solver's results of the same step.  Output: island label and Sleeping
 *      flag per body; APPLYING a decision (taking a sleeping body's constraints out of the step, `Sleeping` component) stays with the host
 *      shim, like the reference's SleepIslands / WakeIslands commands.
 *      One deviation, stated: the reference tracks the split candidate as an island id that a merge can retire (when the candidate is the
 *      smaller island of the merge); here the candidate is the sleepiest BODY, and the island that holds it one step later is split. -------- */
typedef struct AvnIslandsConfig {
    uint32_t body_count, joint_count;
    const uint8_t* body_kind;            /* [B] AvnBodyKind: static bodies have no island */
    const float* sleep_threshold_linear; /* [B] SleepThreshold::linear  (NULL = 0.15; negative = never sleeps) */
    const float* sleep_threshold_angular;/* [B] SleepThreshold::angular (NULL = 0.15) */
    const uint8_t* sleeping_disabled;    /* [B] SleepingDisabled marker (NULL = none) */
    const uint32_t* joint_body1;         /* [J] bodies linked by joints (PhysicsIslands::add_joint) */
    const uint32_t* joint_body2;
    float time_to_sleep;                 /* TimeToSleep (default 0.5 s) */
    float length_unit;                   /* PhysicsLengthUnit */
} AvnIslandsConfig;
/* Islands start as one per non-static body, linked by the joints and — when the contact store already holds touching pairs — by those. */
AvnStatus avn_islands_configure(AvnContext* ctx, const AvnIslandsConfig* config);

typedef struct AvnIslandsStep {
    float delta_secs;                    /* in: Time::delta_secs of the step */
    uint32_t _pad;
    const void* linear_velocity;         /* in: [B][3] SolverBody velocities after the solve (column scalar type) */
    const void* angular_velocity;        /* in: [B][3] */
    const uint8_t* wake;                 /* in: [B] optional: bodies the application touched (wake_on_changed): their islands wake up */
    uint32_t* island;                    /* out: [B] island label = smallest body index of the island (0xFFFFFFFF for static bodies); NULL = skip */
    uint8_t* sleeping;                   /* out: [B] 1 = the body's island sleeps; NULL = skip */
    float* sleep_timer;                  /* out: [B] SleepTimer; NULL = skip */
    uint32_t island_count, sleeping_islands, islands_put_to_sleep, islands_woken, split_bodies, merges;   /* out */
} AvnIslandsStep;
/* Once per step, after avn_contacts_step and the solver stage of the same step (it consumes that step's contact events). */
AvnStatus avn_islands_step(AvnContext* ctx, AvnIslandsStep* step);

AvnStatus avn_get_timings(const AvnContext* ctx, AvnTimings* out);

/*
 * Host-only helper (needs no context and no GPU): the order-preserving level schedule the solver stage uses for
 * joints.  The reference solves joints serially in type order then table order (xpbd/plugin.rs:58-86, 145-189);
 * out_level[g] (g = index in that global order) is the parallel phase joint g runs in: joints of one level touch
 * disjoint written bodies, and every joint runs after all earlier joints it shares a written body with.
 */
AvnStatus avn_joint_levels(const AvnBodyColumns* bodies, const AvnJointSet* joints, uint32_t* out_level, uint32_t* out_level_count);

#ifdef __cplusplus
}
#endif
#endif /* AVIAN_B200_H */
