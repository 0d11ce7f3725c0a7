/* cp_oracle.h -- TEST INFRASTRUCTURE: plain-C restatement of the cpSpaceStep stage algorithms.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use anything under
 * oracle/; the product (chipmunk2d_b200/) never links, imports or executes it.
 *
 * Each function restates one stage of the reference on flat arrays and cites the reference lines it
 * follows.  Parity is PINNED: tests/test_cpu_oracle_restatement.py replays data captured from the
 * unmodified reference (oracle/_ref) through these functions and requires bit-identical results
 * (same compiler, -ffp-contract=off).  The restatement is sequential and single threaded.
 */
#ifndef CP_ORACLE_H
#define CP_ORACLE_H
#include <stdint.h>

typedef struct cpo_vec { double x, y; } cpo_vec;

/* world-space view of one shape for the narrowphase (what cacheData leaves behind) */
typedef struct cpo_shape {
	int type;              /* 0 circle, 1 segment, 2 poly */
	int count;             /* poly vertex count */
	uint64_t hashid;
	cpo_vec a, b, n;       /* circle: a = tc | segment: ta, tb, tn */
	double r;
	double bb[4];          /* l b r t */
	const double *planes;  /* poly: count x (v.x v.y n.x n.y), world space */
	cpo_vec rot;           /* owning body's rotation */
	cpo_vec a_tangent, b_tangent;
} cpo_shape;

typedef struct cpo_manifold {
	int count;
	cpo_vec n;
	cpo_vec p1[2], p2[2];  /* absolute surface points on a / b */
	uint64_t hash[2];
	uint32_t id;
} cpo_manifold;

/* solver view of a body */
typedef struct cpo_body {
	cpo_vec p, v, v_bias, cog, f;
	double a, w, w_bias, t;
	double m_inv, i_inv;
	int type;              /* 0 dynamic, 1 kinematic, 2 static */
} cpo_body;

typedef struct cpo_contact { cpo_vec r1, r2; double nMass, tMass, bounce, jnAcc, jtAcc, jBias, bias; } cpo_contact;
typedef struct cpo_arbiter {
	int body_a, body_b, count, first_collision;
	cpo_vec n, surface_vr;
	double e, u;
	cpo_contact contacts[2];
} cpo_arbiter;

typedef struct cpo_joint {
	int type;              /* CPB_JOINT_* numbering of scenes/cpb_scene.h */
	int a, b;
	double maxForce, errorBias, maxBias;
	cpo_vec anchorA, anchorB;
	double prm[4];
	/* solver state */
	cpo_vec r1, r2, n, bias2, jAcc2;
	double nMass, bias, jnAcc, k[4], target_vrn, v_coef, iSum, clamp;
} cpo_joint;

int cp_oracle_version(void);
int cpo_sizeof(int what);

/* K1 / K9 */
void cpo_body_update_position(cpo_body *b, double dt, double transform6[6]);
void cpo_body_update_velocity(cpo_body *b, cpo_vec gravity, double damping, double dt);
/* K2 */
void cpo_cache_circle(cpo_vec c, double r, const double T[6], cpo_shape *out);
void cpo_cache_segment(cpo_vec a, cpo_vec b, cpo_vec n, double r, const double T[6], cpo_shape *out);
void cpo_cache_poly(int count, const double *local_planes, double r, const double T[6], double *world_planes, cpo_shape *out);
/* K3+K4: brute-force overlapping-pair set over cached AABBs with the QueryReject rules */
long cpo_pairs(int n_shapes, const double *bb4, const int *body, const int *body_active, const uint64_t *group,
	const uint32_t *categories, const uint32_t *mask, int n_nocollide, const uint64_t *nocollide_body_pairs, long cap, uint64_t *out);
/* K5 */
void cpo_collide(const cpo_shape *a, const cpo_shape *b, cpo_manifold *out);
/* K8 / K11 for contacts */
void cpo_arbiter_prestep(cpo_arbiter *arb, const cpo_body *bodies, double dt, double slop, double bias_coef);
void cpo_arbiter_apply_cached(cpo_arbiter *arb, cpo_body *bodies, double dt_coef);
void cpo_arbiter_apply_impulse(cpo_arbiter *arb, cpo_body *bodies);
/* K8 / K11 for all ten joint classes */
void cpo_joint_prestep(cpo_joint *j, cpo_body *bodies, const double *transforms6, double dt);
void cpo_joint_apply_cached(cpo_joint *j, cpo_body *bodies, double dt_coef);
void cpo_joint_apply_impulse(cpo_joint *j, cpo_body *bodies, double dt);
/* the solver loop of cpSpaceStep in the given order */
void cpo_solve(int n_arb, cpo_arbiter *arbs, int n_joints, cpo_joint *joints, cpo_body *bodies, int iterations, double dt, double dt_coef);
/* the same loop over one merged sequence: item >= 0 = arbiter index, item < 0 = joint -(item + 1) (a coloured order replayed) */
void cpo_solve_sequence(long n_items, const int64_t *items, cpo_arbiter *arbs, cpo_joint *joints, cpo_body *bodies, int iterations, double dt, double dt_coef);

#endif
