/* cpb200.h -- C ABI of the B200 step engine (libcpb200.so).
 *
 * This is the drop-in boundary of the hot path: everything cpSpaceStep()
 * (reference src/cpSpaceStep.c:335-445) does between "the host object graph is
 * final" and "body / arbiter state is readable again" happens behind these
 * entry points, as hand-written sm_100a kernels.  The C99 host layer
 * (chipmunk2d_b200/host/, which implements the public Chipmunk2D API of
 * include/chipmunk/chipmunk.h) is the only intended caller; INTEGRATION.md shows
 * the equivalent binding inside the reference's own cpSpace.c / cpSpaceStep.c.
 *
 * Plain C: pointers + counts, no C++ or torch types.  All floating point is
 * IEEE double (cpFloat = double, chipmunk_types.h:55-68).  There is NO CPU
 * fallback: every call fails loudly (non-zero return + cpb200_last_error())
 * when no CUDA device is usable.
 *
 * Each entry point cites the reference interface it replaces.
 */
#ifndef CPB200_H
#define CPB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CPB200_API __attribute__((visibility("default")))
#else
#define CPB200_API
#endif

typedef struct cpb200_world cpb200_world;

enum { CPB200_BODY_DYNAMIC = 0, CPB200_BODY_KINEMATIC = 1, CPB200_BODY_STATIC = 2 };
enum { CPB200_SHAPE_CIRCLE = 0, CPB200_SHAPE_SEGMENT = 1, CPB200_SHAPE_POLY = 2 };
enum {
	CPB200_JOINT_PIN = 0, CPB200_JOINT_SLIDE = 1, CPB200_JOINT_PIVOT = 2, CPB200_JOINT_GROOVE = 3,
	CPB200_JOINT_DAMPED_SPRING = 4, CPB200_JOINT_DAMPED_ROTARY_SPRING = 5, CPB200_JOINT_ROTARY_LIMIT = 6,
	CPB200_JOINT_RATCHET = 7, CPB200_JOINT_GEAR = 8, CPB200_JOINT_SIMPLE_MOTOR = 9
};
/* arbiter states, numerically equal to enum cpArbiterState (chipmunk_structs.h:83-95) */
enum {
	CPB200_ARB_FIRST_COLLISION = 0, CPB200_ARB_NORMAL = 1, CPB200_ARB_IGNORE = 2,
	CPB200_ARB_CACHED = 3, CPB200_ARB_INVALIDATED = 4
};

/* Per-space parameters: the fields of struct cpSpace read by the step
 * (chipmunk_structs.h:396-412; setters cpSpace.h:82-124). A world holds
 * n_spaces independent spaces (1 for a plain cpSpace, many for batched use). */
typedef struct cpb200_space_params {
	double gravity[2];
	double damping;               /* per-second damping; the engine applies pow(damping, dt) computed on the host (cpSpaceStep.c:399) */
	double idle_speed_threshold;
	double sleep_time_threshold;  /* INFINITY disables sleeping (cpSpace.c:155) */
	double collision_slop;
	double collision_bias;
	uint32_t collision_persistence;
	int32_t iterations;
} cpb200_space_params;

/* struct cpBody (chipmunk_structs.h:35-81) minus pointers. */
typedef struct cpb200_body_desc {
	double p[2], v[2], f[2];
	double a, w, t;
	double rot[2];                /* transform.a, transform.b = the host's cos(a), sin(a) (cpBody.c:347-357) */
	double m, i;                  /* INFINITY for kinematic/static (cpBody.c:156-169) */
	double cog[2];
	double v_bias[2], w_bias;
	double idle_time;             /* sleeping.idleTime */
	int32_t type;                 /* CPB200_BODY_* */
	int32_t space;                /* index of the owning space within the world */
	int32_t sleeping;             /* 1 = body is asleep (sleeping.root != NULL) */
	int32_t sleep_group;          /* bodies asleep together share a group id (component root); -1 when awake */
	int32_t custom;               /* CPB200_BODY_HOST_POSITION | CPB200_BODY_HOST_VELOCITY: integrated by a host callback */
	int32_t pad_;
} cpb200_body_desc;
/* Bodies with a user-supplied position_func / velocity_func (cpBody.c:482-491): the device skips its own integrator for
 * them (K1 / K9) and the host layer runs the callback on its mirror at the point of the step where the reference calls it
 * (cpSpaceStep.c:362-367, 398-404), through the split step. */
#define CPB200_BODY_HOST_POSITION 1
#define CPB200_BODY_HOST_VELOCITY 2

/* struct cpShape + cpCircleShape / cpSegmentShape / cpPolyShape (chipmunk_structs.h:177-236). */
typedef struct cpb200_shape_desc {
	int32_t type;                 /* CPB200_SHAPE_* */
	int32_t body;                 /* index into the uploaded body array */
	uint32_t hashid;              /* shape->hashid (cpSpace.c:432): stable id used in contact hashes and arbiter keys */
	int32_t sensor;
	uint32_t categories, mask;
	uint64_t group;
	uint64_t collision_type;
	double e, u;
	double surface_v[2];
	double r;
	double a[2], b[2];            /* circle: a = c (offset); segment: a, b (body-local end points) */
	double a_tangent[2], b_tangent[2];
	int32_t n_verts, vert_offset; /* poly: range in the uploaded vertex array (hull order, body-local) */
} cpb200_shape_desc;

/* struct cpConstraint + the joint structs (chipmunk_structs.h:250-382). */
typedef struct cpb200_joint_desc {
	int32_t type;                 /* CPB200_JOINT_* */
	int32_t a, b;                 /* body indices */
	int32_t collide_bodies;
	double max_force, error_bias, max_bias;
	double anchor_a[2], anchor_b[2];
	double prm[4];                /* same packing as cpb_scene_joint.prm */
	double acc[2];                /* accumulated impulse carried between steps (jnAcc / jAcc) */
} cpb200_joint_desc;

/* State read back per body: what cpBodyUpdatePosition/Velocity + the solver
 * changed (cpBody.c:493-522).  80 bytes. */
typedef struct cpb200_body_state {
	double p[2], v[2];
	double a, w;
	double rot[2];                /* cos a, sin a: the device's transform rotation (cpBody.c:347-357) */
	double idle_time;
	int32_t sleeping;
	int32_t sleep_group;
} cpb200_body_state;

/* One active or cached arbiter (struct cpArbiter + struct cpContact,
 * chipmunk_structs.h:101-145).  r1/r2 are body-relative like the reference's. */
typedef struct cpb200_arbiter {
	int32_t shape_a, shape_b;     /* uploaded shape indices, in cpCollide order (cpCollision.c:701-726) */
	int32_t body_a, body_b;
	int32_t count;                /* 0 for cached / sensor / rejected arbiters */
	int32_t state;                /* CPB200_ARB_* */
	uint32_t stamp;
	int32_t active;               /* 1 = solved this step (pushed to space->arbiters, cpSpaceStep.c:274); 2 = dormant: kept with
	                               * its contacts while its bodies sleep (cpSpaceComponent.c:94-105); 0 = inactive */
	int32_t record, pad;          /* index of the device record (cpb200_world_edit_arbiters) */
	double n[2];
	double e, u;
	double surface_vr[2];
	struct {
		double r1[2], r2[2];
		double n_mass, t_mass, bounce, bias;
		double jn_acc, jt_acc, j_bias;
		uint64_t hash;
	} contacts[2];
} cpb200_arbiter;

/* Per-joint solver state after a step (for cpConstraintGetImpulse and warm starting). */
typedef struct cpb200_joint_state {
	double acc[2];                /* jnAcc | jAcc(x,y) */
	double impulse;               /* what klass->getImpulse would return */
	double aux;                   /* ratchet: current angle */
} cpb200_joint_state;

/* Step statistics (new).  Reduced over spaces on the device. */
typedef struct cpb200_stats {
	uint64_t steps;
	uint32_t n_bodies, n_awake, n_shapes, n_joints;
	uint32_t n_pairs;             /* survivors of the QueryReject rules (cpSpaceStep.c:219-232) */
	uint32_t n_arbiters;          /* active arbiters */
	uint32_t n_contacts;          /* contacts in active arbiters */
	uint32_t n_cached;            /* arbiters kept in the cache but inactive */
	uint32_t n_colours;           /* colours used by the constraint graph this step */
	uint32_t overflow;            /* non-zero: a device buffer was too small; results of that step are invalid */
	double kinetic_energy;        /* sum over dynamic bodies of m*v^2 + i*w^2 (cpBody.c:581-588, not halved) */
	double max_penetration;       /* max over active contacts of -dist (>= 0) */
	uint32_t n_row_solves;        /* world-wide solver: contact rows x iterations visited in the last step ... */
	uint32_t n_row_idle;          /* ... of which this many left both bodies bit-identical (clamped impulses): their scatters were skipped */
} cpb200_stats;

/* ---- lifetime ------------------------------------------------------------------ */

/* Replaces the allocation half of cpSpaceInit (cpSpace.c:119-178).  `device` is the CUDA
 * ordinal; n_spaces >= 1.  Returns NULL on failure (see cpb200_last_error). */
CPB200_API cpb200_world *cpb200_world_create(int device, int n_spaces);
/* Replaces cpSpaceDestroy (cpSpace.c:188-229). */
CPB200_API void cpb200_world_destroy(cpb200_world *w);
/* Thread-local description of the last failure; never NULL. */
CPB200_API const char *cpb200_last_error(void);
/* 1 when the library was built for sm_100a and a CUDA device is present. */
CPB200_API int cpb200_device_available(void);

/* ---- upload (host object graph -> SoA device buffers) -------------------------- */

/* Space property setters (cpSpace.h:82-124). */
CPB200_API int cpb200_world_set_space_params(cpb200_world *w, int space, const cpb200_space_params *p);
/* Replaces the body/shape/constraint registries filled by cpSpaceAddBody/AddShape/
 * AddConstraint (cpSpace.c:417-474).  Each call REPLACES the whole set; arbiters are
 * keyed by shape hashid and survive re-uploads.  Buffers are HOST memory. */
CPB200_API int cpb200_world_set_bodies(cpb200_world *w, int n, const cpb200_body_desc *bodies);
CPB200_API int cpb200_world_set_shapes(cpb200_world *w, int n, const cpb200_shape_desc *shapes, int n_verts, const double *verts_xy);
CPB200_API int cpb200_world_set_joints(cpb200_world *w, int n, const cpb200_joint_desc *joints);
/* Overwrite kinematic state of bodies [first, first+n) without touching anything else
 * (cpBodySetPosition/Velocity/Angle... between steps, cpBody.c:374-467). */
CPB200_API int cpb200_world_update_bodies(cpb200_world *w, int first, int n, const cpb200_body_desc *bodies);
/* Forces only -- the usual per-step host input (cpBodySetForce / cpBodyApplyForceAt*, cpBody.c:420-424,
 * 534-543): fxyt[n][3] = f.x f.y torque for bodies [first, first+n). */
CPB200_API int cpb200_world_set_body_forces(cpb200_world *w, int first, int n, const double *fxyt);
/* Page-locked host memory for the per-step exchange buffers (forces in, body states out): a transfer from/to
 * such a buffer is one DMA at full link speed, without the driver's bounce copy.  NULL on failure. */
CPB200_API void *cpb200_host_alloc(size_t bytes);
CPB200_API void cpb200_host_free(void *p);
/* cpBodyActivate on bodies that are awake (cpSpaceComponent.c:113-119): their idle timers restart.  Waking a
 * SLEEPING body changes the contact graph and goes through cpb200_world_update_bodies instead. */
CPB200_API int cpb200_world_touch_bodies(cpb200_world *w, int n, const int32_t *indices);
/* Pre-size device buffers (pairs, arbiters) for at least this many; 0 keeps the default. */
CPB200_API int cpb200_world_reserve(cpb200_world *w, int max_pairs, int max_arbiters);

/* ---- the step ------------------------------------------------------------------- */

/* ---- structural edits by range (cpSpaceAddBody / AddShape / AddConstraint, cpSpace.c:417-474) ----
 * Objects are appended behind the existing ones: nothing that is already on the device moves, is re-uploaded or loses
 * its cached arbiters, colours or warm-start state.  New shapes name bodies by their final index (append bodies first);
 * vert_offset of new polygons counts from the first appended vertex.  The object arrays carry slack (a quarter of their
 * size); a call that does not fit returns 1 WITHOUT changing anything and the caller falls back to the full
 * cpb200_world_set_bodies / set_shapes / set_joints sequence, which allocates new slack.  Bound I/O buffers
 * (cpb200_world_bind_io) are unbound by cpb200_world_append_bodies. */
CPB200_API int cpb200_world_append_bodies(cpb200_world *w, int n, const cpb200_body_desc *bodies);
CPB200_API int cpb200_world_append_shapes(cpb200_world *w, int n, const cpb200_shape_desc *shapes, int n_verts, const double *verts_xy);
CPB200_API int cpb200_world_append_joints(cpb200_world *w, int n, const cpb200_joint_desc *joints);
/* cpSpaceRemoveShape / RemoveBody / RemoveConstraint (cpSpace.c:476-571) in place: the LAST object of the class moves
 * into the hole -- a host that keeps its registries dense the same way (swap with last) keeps slot == index -- and every
 * index that named the moved object is re-pointed on the device (arbiter records, shapes' body indices, joints, sleeping
 * groups).  Arbiter records of a removed shape die.  A body can only be removed once no shape or joint names it.
 * Returns 1 (nothing changed) when the engine's bookkeeping cannot follow; the caller then re-uploads. */
CPB200_API int cpb200_world_remove_shape(cpb200_world *w, int index);
CPB200_API int cpb200_world_remove_body(cpb200_world *w, int index);
CPB200_API int cpb200_world_remove_joint(cpb200_world *w, int index);

/* cpSpaceStep (cpSpaceStep.c:335-445) for every space of the world.  Asynchronous:
 * returns once the kernels are enqueued on the world's stream. */
CPB200_API int cpb200_world_step(cpb200_world *w, double dt);
/* Per-step host I/O bound to the world and overlapped with the step.  While bound, EVERY step (cpb200_world_step or the
 * split step) reads the bodies' forces from forces_fxyt[n_bodies][3] (f.x f.y torque; what cpBodySetForce / SetTorque
 * leave in the cpBody, cpBody.c:420-424) and writes state_out[2][n_bodies][3] = (p.x p.y a) of every body, then
 * (v.x v.y w) of every body.  The copies run on a side stream: the forces are only consumed by the velocity update
 * (cpBodyUpdateVelocity, cpSpaceStep.c:398-404), so they travel during the collision phase; positions are final after
 * cpBodyUpdatePosition (cpSpaceStep.c:362-367), so they travel during the rest of the step; only the velocities are
 * copied after the solver.  Both buffers must be page-locked (cpb200_host_alloc); either may be NULL; (NULL, NULL)
 * unbinds, and so does cpb200_world_set_bodies.  The caller may write forces_fxyt and read state_out between
 * cpb200_world_sync and the next step. */
CPB200_API int cpb200_world_bind_io(cpb200_world *w, const double *forces_fxyt, double *state_out);
/* A step whose launch sequence repeats (same dt, no upload in between, no profiling) is captured into a CUDA graph on
 * its second occurrence and replayed with a single launch afterwards.  Results are identical either way
 * (tests/test_gpu_graph.py); the hook turns the replay off (0) or on (1; default unless CPB200_NO_GRAPH is set).
 * out2 = (graphs captured, steps replayed). */
CPB200_API int cpb200_world_set_graph(cpb200_world *w, int on);
CPB200_API int cpb200_world_get_graph_stats(cpb200_world *w, unsigned long long *out2);
/* Empty, or why capturing this world's step failed (the world then launches kernel by kernel for good). */
CPB200_API const char *cpb200_world_graph_error(cpb200_world *w);
/* The same step in two halves for spaces with collision handlers (cpSpaceStep.c:234-290): after
 * cpb200_world_step_collide the records touched this step can be read with cpb200_world_get_arbiters, the
 * handlers' decisions are written back with cpb200_world_edit_arbiters, cpb200_world_step_finish runs
 * islands, cache ageing, prestep, velocity integration and the solver. */
CPB200_API int cpb200_world_step_collide(cpb200_world *w, double dt);
CPB200_API int cpb200_world_step_finish(cpb200_world *w);
#define CPB200_EDIT_IGNORE   1u   /* begin() returned false / cpArbiterIgnore: ignored until the shapes separate (cpArbiter.c:46-50) */
#define CPB200_EDIT_REJECT   2u   /* preSolve() returned false: not solved this step (cpSpaceStep.c:275-285) */
#define CPB200_EDIT_MATERIAL 4u   /* cpArbiterSetRestitution / SetFriction / SetSurfaceVelocity (cpArbiter.c:97-143) */
#define CPB200_EDIT_CONTACTS 8u   /* cpArbiterSetContactPointSet (cpArbiter.c:181-209): n, r1, r2 of the existing contacts */
typedef struct cpb200_arbiter_edit {
	int32_t record; uint32_t flags;
	double e, u, surface_vr[2];
	double n[2], r1[2][2], r2[2][2];
} cpb200_arbiter_edit;
CPB200_API int cpb200_world_edit_arbiters(cpb200_world *w, int n, const cpb200_arbiter_edit *edits);
/* n steps timed with CUDA events recorded on the world's stream; *ms = device time in milliseconds. */
CPB200_API int cpb200_world_time_steps(cpb200_world *w, double dt, int n, float *ms);
/* Number of kernels this library has launched in this process (all worlds). */
CPB200_API unsigned long long cpb200_launch_count(void);
/* Block until all enqueued steps have finished; returns non-zero on a device error or
 * buffer overflow. */
CPB200_API int cpb200_world_sync(cpb200_world *w);

/* ---- read-back ------------------------------------------------------------------ */

/* Body state after the last step (host buffer of n entries; implies a sync). */
CPB200_API int cpb200_world_get_bodies(cpb200_world *w, int first, int n, cpb200_body_state *out);
/* Pending bias velocities: out[n][3] = v_bias.x, v_bias.y, w_bias.  The solver of the last step wrote them and
 * the NEXT step's position update consumes them (cpBodyUpdatePosition, cpBody.c:511-522), so a host that
 * re-uploads bodies between two steps (cpb200_world_update_bodies / set_bodies) reads them first and hands them
 * back in cpb200_body_desc.v_bias / w_bias -- in the reference they simply stay in the cpBody (implies a sync). */
CPB200_API int cpb200_world_get_body_bias(cpb200_world *w, int first, int n, double *out);
/* Cached shape AABBs (cpShapeGetBB, cpShape.h:124): out[n][4] = l b r t. */
CPB200_API int cpb200_world_get_shape_bbs(cpb200_world *w, int first, int n, double *out);
/* Arbiters of the last step (space->arbiters + cachedArbiters, cpSpaceStep.c:250-288).
 * Returns the number available; writes at most cap.  active_only != 0 keeps only solved ones. */
CPB200_API int cpb200_world_get_arbiters(cpb200_world *w, int cap, cpb200_arbiter *out, int active_only);
CPB200_API int cpb200_world_get_joints(cpb200_world *w, int first, int n, cpb200_joint_state *out);
CPB200_API int cpb200_world_get_stats(cpb200_world *w, cpb200_stats *out);

/* ---- validation hooks (used by the parity tests; not needed by a normal host) ---- */

/* Overlapping-pair set of the last step as sorted (min<<32|max) of uploaded shape indices
 * (SURVEY.md 8a a6/a7 contract).  Returns the count; writes at most cap. */
CPB200_API long cpb200_world_get_pairs(cpb200_world *w, long cap, uint64_t *out);
/* Solver order: 0 = graph-coloured parallel Gauss-Seidel (default);
 * 1 = serial, arbiters in the order given to cpb200_world_set_arbiter_order (or by
 * ascending key when none was given), then joints in upload order -- the reference's
 * sequential order (cpSpaceStep.c:418-427), used for tight one-step parity. */
CPB200_API int cpb200_world_set_solver_mode(cpb200_world *w, int mode);
/* order[n] = (shape index a)<<32 | (shape index b) in the sequence the reference pushed
 * its arbiters; pairs not listed are solved afterwards in key order. Applies to the next step only. */
CPB200_API int cpb200_world_set_arbiter_order(cpb200_world *w, int n, const uint64_t *order);
/* Force the persistent coloured solver to run on `blocks` CTAs (0 = size automatically).  Results do
 * not depend on the grid size; the determinism test runs 1 CTA against the full grid. */
CPB200_API int cpb200_world_set_solver_grid(cpb200_world *w, int blocks);
/* Serial mode only: order[n] = joint indices in the sequence the reference holds them in
 * space->constraints (sleeping reorders that array); joints not listed follow in upload order.
 * Applies to the next step only. */
CPB200_API int cpb200_world_set_joint_order(cpb200_world *w, int n, const int32_t *order);
/* ---- host callbacks in the middle of a step (custom integrators / spring force functions; SURVEY.md 8b slow path) ----
 * All three are meant for the gaps of the split step: positions are final after cpb200_world_step_collide, velocities
 * are integrated (for bodies without CPB200_BODY_HOST_VELOCITY) by cpb200_world_step_presolve. */
/* state of the listed bodies (any order) */
CPB200_API int cpb200_world_get_bodies_indexed(cpb200_world *w, int n, const int32_t *indices, cpb200_body_state *out);
/* vxyw[n][3] = v.x v.y w of the listed bodies (what a custom velocity_func left in the cpBody); their forces are cleared */
CPB200_API int cpb200_world_set_body_velocities_indexed(cpb200_world *w, int n, const int32_t *indices, const double *vxyw);
/* f[n] = the value springForceFunc(spring, dist) / springTorqueFunc(spring, relativeAngle) returned for the listed damped
 * (rotary) springs: the NEXT prestep uses it instead of the default linear law (cpDampedSpring.c:24-27, 50;
 * cpDampedRotarySpring.c:24-27, 48).  Applies to one prestep only. */
CPB200_API int cpb200_world_set_spring_forces(cpb200_world *w, int n, const int32_t *joints, const double *f);

/* ---- the PRODUCTION solver order, pinned against the oracle (tests/test_gpu_production_order.py) ----
 * A split step can also stop right before the solver: step_collide -> step_presolve (islands, cache ageing,
 * prestep, velocity integration: everything of cpSpaceStep.c:335-404) -> [read the solver's inputs] ->
 * step_finish (K10 colouring + K11 warm start / iterations, cpSpaceStep.c:406-427). */
CPB200_API int cpb200_world_step_presolve(cpb200_world *w);
/* Solver inputs / outputs as the kernels hold them: out[n][8] = v.x v.y w m_inv v_bias.x v_bias.y w_bias i_inv. */
CPB200_API int cpb200_world_get_body_solver_state(cpb200_world *w, int first, int n, double *out);
/* out[n][CPB200_JOINT_SOLVER_ROW] = type a b live max_force max_bias r1.xy r2.xy n.xy nMass|iSum|clamp k[4] bias.xy
 * jAcc.xy aux0 (spring target_vrn | ratchet angle) aux1 (spring v_coef) prm[4] 0 */
#define CPB200_JOINT_SOLVER_ROW 28
CPB200_API int cpb200_world_get_joint_solver_state(cpb200_world *w, int first, int n, double *out);
/* Which coloured kernel family runs: 0 automatic (default), 1 world-wide persistent kernel with L2-cached rows,
 * 2 the same with streamed rows, 3 space-local (one CTA per space; fails at step time if the world's layout
 * does not allow it).  The arithmetic is the same; the tests pin every family. */
CPB200_API int cpb200_world_set_solver_variant(cpb200_world *w, int variant);
/* The sequence in which the last step's coloured solver visited its constraints, colour phase after colour
 * phase (space after space for the space-local family): item >= 0 = arbiter record index (cpb200_arbiter.record),
 * item < 0 = joint -(item + 1).  Constraints of one phase share no dynamic body, so replaying the sequence
 * with a sequential solver (cpSpaceStep.c:406-427) must give the parallel result bit for bit.
 * Returns the count; writes at most cap. */
CPB200_API long cpb200_world_get_solver_order(cpb200_world *w, long cap, int64_t *out);
/* out[2][65]: begin of every colour in the solver's row list, then in its joint list (world-wide coloured step). */
CPB200_API int cpb200_world_get_colour_starts(cpb200_world *w, int32_t *out);
/* 0 serial, 1 world-wide coloured, 2 space-local: what the last step ran. */
CPB200_API int cpb200_world_get_solver_path(cpb200_world *w);
/* Run only the narrowphase on one uploaded shape pair with current world caches
 * (cpShapesCollide, cpShape.c:259-283): out = count n.x n.y (pA.xy pB.xy dist) x2. */
CPB200_API int cpb200_world_collide_pair(cpb200_world *w, int shape_a, int shape_b, double *out13);
/* ---- space queries (cpSpaceQuery.c:24-246) against the world cache of the last step ---------------------
 * One data-parallel scan over all shapes per query (filter reject -> AABB reject -> exact per-shape test of
 * cpShape.c:298-455 / cpPolyShape.c:66-145).  `space` restricts a batched world to one space (-1 = all).
 * All functions return the number of hits found (which may exceed `cap`; only the first `cap` records, in
 * ascending shape index, are written) or -1 on error. */
typedef struct cpb200_filter { uint64_t group; uint32_t categories, mask; } cpb200_filter;
/* point query: d = signed distance, g = gradient | segment query: d = alpha in [0,1], g = surface normal */
typedef struct cpb200_query_hit { int32_t shape; int32_t pad; double point[2]; double d; double g[2]; } cpb200_query_hit;
/* cpSpacePointQuery (cpSpaceQuery.c:33-59): every shape with distance < max_distance, sensors included. */
CPB200_API int cpb200_world_point_query(cpb200_world *w, int space, const double point[2], double max_distance, const cpb200_filter *filter, int cap, cpb200_query_hit *out);
/* cpSpacePointQueryNearest (cpSpaceQuery.c:61-101): the closest non-sensor shape; returns 0 or 1. */
CPB200_API int cpb200_world_point_query_nearest(cpb200_world *w, int space, const double point[2], double max_distance, const cpb200_filter *filter, cpb200_query_hit *out);
/* cpSpaceSegmentQuery (cpSpaceQuery.c:113-147): every shape the swept circle of `radius` hits, sensors included. */
CPB200_API int cpb200_world_segment_query(cpb200_world *w, int space, const double a[2], const double b[2], double radius, const cpb200_filter *filter, int cap, cpb200_query_hit *out);
/* cpSpaceSegmentQueryFirst (cpSpaceQuery.c:149-188): the first non-sensor hit with alpha < 1; returns 0 or 1. */
CPB200_API int cpb200_world_segment_query_first(cpb200_world *w, int space, const double a[2], const double b[2], double radius, const cpb200_filter *filter, cpb200_query_hit *out);
/* cpSpaceBBQuery (cpSpaceQuery.c:190-221): shapes whose cached AABB intersects bb = (l, b, r, t). */
CPB200_API int cpb200_world_bb_query(cpb200_world *w, int space, const double bb[4], const cpb200_filter *filter, int cap, int32_t *shapes);
/* cpShapePointQuery / cpShapeSegmentQuery (cpShape.c:223-258) on one uploaded shape: no filter, no range. */
CPB200_API int cpb200_world_shape_point_query(cpb200_world *w, int shape, const double point[2], cpb200_query_hit *out);
CPB200_API int cpb200_world_shape_segment_query(cpb200_world *w, int shape, const double a[2], const double b[2], double radius, cpb200_query_hit *out);
/* cpSpaceShapeQuery (cpSpaceQuery.c:223-246): narrowphase of a caller-supplied shape, given in world space as
 * cacheData would leave it, against every shape whose AABB it overlaps.  `self` = its index if it is part of
 * the world (never reported), else -1.  Polygon: verts_normals = count x (v.x v.y n.x n.y), world space. */
typedef struct cpb200_query_shape {
	int32_t type, count, self, pad;
	double r;
	double a[2], b[2], n[2];         /* circle: a = centre | segment: a, b, n */
	double rot[2];                   /* owning body's rotation */
	double a_tangent[2], b_tangent[2];
	double bb[4];
	cpb200_filter filter;
} cpb200_query_shape;
typedef struct cpb200_shape_hit { int32_t shape, count; double normal[2]; double points[2][5]; /* pointA.xy pointB.xy distance */ } cpb200_shape_hit;
CPB200_API int cpb200_world_shape_query(cpb200_world *w, int space, const cpb200_query_shape *q, const double *verts_normals, int cap, cpb200_shape_hit *out);

/* Per-stage device time of the last step in microseconds (CUDA events), names via
 * cpb200_stage_name(i).  Returns the number of stages. */
CPB200_API int cpb200_world_get_stage_times(cpb200_world *w, int cap, float *usec);
CPB200_API const char *cpb200_stage_name(int i);
/* Inside the persistent colour+solve kernel of the last step (device globaltimer): microseconds spent
 * colouring, building rows, warm starting, iterating; out6[4] = colouring rounds, out6[5] = constraints that
 * had to be (re)coloured (the rest kept last step's colour). */
CPB200_API int cpb200_world_get_solver_profile(cpb200_world *w, double *out6);
/* Enable (1) / disable (0) per-stage event timing (adds syncs; off by default). */
CPB200_API int cpb200_world_set_profiling(cpb200_world *w, int on);

/* Self-tests of the device-wide primitives under the broadphase (radix sort of (u64 key, i32
 * value) pairs by the low `bits` key bits; exclusive scan), host buffers in and out. */
CPB200_API int cpb200_debug_sort_pairs(int device, int n, int bits, uint64_t *keys, int32_t *vals);
CPB200_API int cpb200_debug_exclusive_scan(int device, int n, uint32_t *data);

#ifdef __cplusplus
}
#endif
#endif
