// cpb_world.h -- SoA device layout of a world (the replacement for the reference's
// cpBody / cpShape / cpArbiter / cpContact / cpConstraint AoS structs,
// chipmunk_structs.h:35-382).  All views are plain structs of device pointers that are
// passed to kernels by value.
#pragma once
#include "cpb_math.h"
#include "../../include/cpb200.h"

// ---- per-space step constants (host computes the pow()s with libm: SURVEY.md 8c) ----
struct DSpace {
	V2 gravity;
	double damping_dt;      // pow(space->damping, dt)              cpSpaceStep.c:399
	double bias_coef;       // 1 - pow(collisionBias, dt)            cpSpaceStep.c:384
	double slop;
	double idle_speed;      // idleSpeedThreshold
	double sleep_threshold; // sleepTimeThreshold
	uint32_t persistence;
	int32_t iterations;
};

// ---- bodies (cpBody, chipmunk_structs.h:35-81) ----
struct DBodies {
	int n;
	V2 *pos;        // p
	double *ang;    // a
	V2 *rot;        // (cos a, sin a)  = transform.a, transform.b
	V2 *txy;        // transform.tx, transform.ty
	V2 *cog;
	double4 *V;     // (v.x, v.y, w, m_inv)           solver-hot: one 32-byte sector
	double4 *VB;    // (v_bias.x, v_bias.y, w_bias, i_inv)
	V2 *MI;         // (m_inv, i_inv)
	V2 *M;          // (m, i)
	V2 *force;      // f
	double *torque; // t
	double *idle;   // sleeping.idleTime
	int *type;      // CPB200_BODY_*
	int *space;
	int *sleeping;  // 0 awake, 1 asleep
	int *sgroup;    // sleeping component id (root body index) or -1
	int *custom;    // CPB200_BODY_HOST_POSITION | CPB200_BODY_HOST_VELOCITY: a host callback integrates this body
};

// ---- shapes (cpShape + cpCircleShape/cpSegmentShape/cpPolyShape) ----
struct DShapes {
	int n, nv;
	int *type, *body;
	uint32_t *hashid;
	uint32_t *hlocal;     // hashid minus the lowest hashid of the shape's space (colouring priorities)
	int *sensor;
	uint32_t *cat, *mask;
	uint64_t *group, *ctype;
	double *e, *u, *r;
	V2 *surfv;
	V2 *la, *lb, *ln;     // body-local: circle c | segment a, b, n
	V2 *atan, *btan;      // segment neighbour tangents
	int *pcount, *poff;   // poly vertex range
	V2 *lpv, *lpn;        // [nv] body-local hull vertices / edge normals (planes[count+i], cpPolyShape.c:147-165)
	// world-space cache written by k_shape_cache (cacheData, cpShape.c:291-296,378-405; cpPolyShape.c:39-64)
	V2 *wa, *wb, *wn;     // circle tc | segment ta, tb, tn
	V2 *wpv, *wpn;        // [nv]
	double4 *bb;          // (l, b, r, t)
	// packed copies for the narrowphase's random gathers (one 32-byte sector instead of 3-6 scattered words)
	double4 *mat;         // (e, u, surface_v.x, surface_v.y), static
	double4 *circ;        // circles, one 64-byte line per shape, rewritten by k_shape_cache: [2s] = (tc.x, tc.y, r, bits) with
	                      // bits = body index (29) | has surface velocity << 29 | body not dynamic << 30 | sensor << 31 | hashid << 32;
	                      // [2s+1] = (body p.x, p.y, e, u).  The circle-circle narrowphase gathers sector 0 of both shapes for the
	                      // test and, on a hit, sector 1 of the same lines: two DRAM bursts per pair instead of ten scattered sectors
	uint2 *ids;           // (hashid, hlocal)
	double4 *filt;        // bit patterns (body | type << 32, categories | mask << 32, group, 0), static: everything QueryReject and
	                      // the pair classes read about a shape in one sector (k_bvh_pairs' flush)
};

// ---- arbiters + contacts (cpArbiter / cpContact, chipmunk_structs.h:101-145) ----
// Double buffered: the previous step's records stay readable for warm starting
// (the role of the reference's contact-buffer ring, cpSpaceStep.c:109-183).
struct DArbs {
	int cap;
	int *count_ptr;        // device counter: number of records
	uint64_t *key;         // (min hashid)<<32 | (max hashid): the unordered shape pair (cpSpaceStep.c:249-251)
	int *sa, *sb;          // shape indices in cpCollide order (a.type <= b.type)
	int *ba, *bb;          // body indices
	int *cnt;              // contact count (kept while CACHED so hashes can still be matched)
	int *state;            // CPB200_ARB_*
	uint32_t *stamp;
	int *active;           // solved this step
	int *seen;             // touched by the following step's collision phase
	uint32_t *gjkid;       // cpCollisionID warm start for GJK
	V2 *n;
	double *e, *u;
	V2 *svr;               // surface_vr
	// contacts: contact k of record i at CIDX(A, i, k) = k*cap + i (two planes: records with one contact, the
	// common case, then touch dense memory; prev and cur always share cap)
	V2 *r1, *r2;
	double *nmass, *tmass, *bounce, *bias;
	double *jn, *jt, *jb;  // jnAcc, jtAcc, jBias
	uint64_t *hash;
	int *colour;           // colour assigned this step (-1 = not in the solver)
	uint64_t *pri;         // colouring priority: hash of the space-local shape pair (same in a batch as alone)
	int *hint;             // last step's colour if the arbiter was solved then, else -1
	// what the NEXT step's collision phase needs from a record, packed into one 64-byte line by k_pack_warm:
	// warm[2i] = (jn0, jt0, jn1, jt1), warm[2i+1] = (hash0, hash1, gjkid<<32 | colour<<24 | active<<16 | cnt<<8 | state, -)
	double4 *warm;
};

#define CIDX(A, i, k) ((k)*(A).cap + (i))

// open-addressing table: shape-pair key -> arbiter record index (replaces cpHashSet cachedArbiters)
struct DTable {
	uint32_t mask;         // allocated capacity - 1 (capacity is a power of two)
	ulonglong2 *slots;     // x = key (0 = empty), y = record index: one 16-byte access per probe
	uint32_t *dmask;       // device word: the mask in use = smallest power of two >= 2 x this step's records, minus 1
	                       // (k_table_clear): the table is rebuilt every step, so it is only as large -- to clear and to
	                       // scatter over -- as the step needs, not as the capacity allows
};

// ---- colour-sorted solver rows (one per active arbiter), rebuilt every step ----
struct DRows {
	int cap;
	int *arb;              // back pointer to the arbiter record
	int *ba, *bb;
	int *cnt;
	V2 *n, *svr;
	double *u;
	// contact k of row r at [k*cap + r]
	V2 *r1, *r2;
	double *nmass, *tmass, *bounce, *bias;
	double *jn, *jt, *jb;
	// Packed view for the space-local solver (k_sl_rows / k_sl_solve), carved from separate storage: five vectors
	// per row instead of sixteen scalar planes.  That solver is latency- and instruction-bound (rows come from
	// L1/L2, one warp per space), where fewer, wider requests win 20%; the world-wide solver streams its rows from
	// HBM with a full warp per 256 contiguous bytes, where the scalar planes measured 12% FASTER than 32-byte
	// per-thread vectors and a packed impulse vector doubles the write traffic.
	int4 *hdr;             // (body a, body b, count [negative: first collision], arbiter record)
	double4 *nsv;          // (n.x, n.y, surface_vr.x, surface_vr.y)
	double4 *r12;          // [k*cap + r] (r1.x, r1.y, r2.x, r2.y)
	double4 *mass;         // [k*cap + r] (nMass, tMass, bias, bounce)
	double4 *imp;          // [k*cap + r] (jnAcc, jtAcc, jBias, friction u)
};

// ---- joints (cpConstraint + joint structs, chipmunk_structs.h:250-382) ----
struct DJoints {
	int n;
	int *type, *a, *b;
	double *max_force, *max_bias, *bias_coef;   // bias_coef = 1 - pow(errorBias, dt), host libm (chipmunk_private.h:264-268)
	V2 *anchor_a, *anchor_b;
	double4 *prm;
	// solver state
	V2 *r1, *r2, *nrm;     // nrm doubles as grv_tn for groove joints
	double *nmass;         // nMass | iSum | clamp (groove)
	double4 *k;            // pivot / groove mass tensor (k11 k12 k21 k22 as cpMat2x2 a b c d)
	V2 *bias;              // scalar joints use .x
	V2 *acc;               // jnAcc / jAcc
	double *aux0, *aux1;   // spring: target_vrn, v_coef | ratchet: angle
	V2 *jspring;           // (value, 1.0) = the spring force / torque a host callback returned for the next prestep; (., 0) = default law
	int *colour;
	int *row;              // colour-sorted order: row -> joint index
	uint64_t *pri;         // colouring priority: hash of the space-local joint index
	int *hint;             // last step's colour, -1 if none
};

// ---- broadphase scratch (LBVH over all shapes, rebuilt every step) ----
struct DBvh {
	int n;                 // number of leaves (= shapes)
	uint64_t *keys;        // sorted morton keys
	int *leaf_shape;       // sorted position -> shape index
	int *left, *right;     // [n-1] children; >= n-1 means leaf (value - (n-1))
	int *parent;           // [2n-1]
	double4 *nbb;          // [2n-1] node bounds; leaves at n-1+i
	int2 *nsp;             // [2n-1] node space-id range
	int *flags;            // [n-1] refit arrival counters
	double *bounds;        // [4] world bounds l b r t
	int *nskip;            // [2n-1] subtree skip key: last leaf position if every leaf below is active, INT_MAX otherwise
	// traversal layout, packed after the refit: one dependent load per level
	double4 *cbox;         // [2(n-1)] boxes of the two children of internal node i at [2i], [2i+1]
	int4 *cinfo;           // [n-1] (left, right, left skip key, right skip key)
	int4 *cspace;          // [n-1] (left space min, max, right space min, max)
	// nodes (leaves or internal) whose own leaf range lies inside one refit window while their parent's does not: where the
	// shared-memory refit hands over to the pass through L2.  Listed by k_bvh_build, valid as long as the topology is.
	int *top_list;         // [n]
	int *top_count;        // [1]
};

// pair lists by class: 0 circle-circle, 1 circle-segment, 2 everything that needs GJK
struct DPairs {
	int cap;
	int *count;            // [4]: pairs per class, [3] = raw candidates
	int *a[3], *b[3];      // shape indices, a.type <= b.type
	int2 *cand;            // [cap] leaf hits of the tree traversal, before QueryReject (k_bvh_pairs -> k_pair_filter)
};

// device-side step counters / flags
struct DCounters {
	int n_pairs[3];
	int n_contacts;
	int n_active;          // active arbiters
	int n_colours;
	int overflow;          // bit 0 pairs, 1 arbiters, 2 table, 3 bvh stack, 4 colours
	int colour_remaining[2];
	int colour_rounds;
	int n_overflow_colour;
	int n_cached;
	uint32_t stamp;        // space->stamp (cpSpaceStep.c:349): advanced on the device by k_reset_step, so a captured step graph replays it
	int no_gjk_stage;      // experiment switch (env CPB200_NO_GJK_STAGE): k_collide<2> reads polygon vertices from global memory
	int n_row_solves;      // rows x iterations the world-wide solver visited this step ...
	int n_row_idle;        // ... and how many of those changed neither body (clamped impulses: all four scatters skipped)
	unsigned bvh_visits, bvh_queries;   // tree nodes visited / query leaves, sampled (every 16th CTA of k_bvh_pairs): the host's measure of how far an aged topology has degraded
};

#define CPB_MAX_COLOURS 64
