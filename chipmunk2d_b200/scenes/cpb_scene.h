/* cpb_scene.h -- flat, library-independent description of a Chipmunk2D space.
 *
 * A "scene blob" is what the parity tests and the benchmark exchange between
 * the reference build (oracle/_ref) and the B200 build: the reference's demo
 * code (demo/Bench.c, PyramidStack.c, Chains.c) is run once inside oracle/_ref,
 * the resulting cpSpace is flattened into this format by oracle/ref_probe.c,
 * and cpb_scene_load() (scene_io.c, public C API only) re-creates the identical
 * space in either library.  Synthetic scenes (BASELINE.json configs 3-5) are
 * written straight into this format by chipmunk2d_b200/scenes.py.
 *
 * Layout of a blob: cpb_scene_header, then bodies[], shapes[], verts[], joints[].
 * All records are plain little-endian C structs with no pointers.
 */
#ifndef CPB_SCENE_H
#define CPB_SCENE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPB_SCENE_MAGIC 0x31454e4543535043ull /* "CPSCENE1" */

enum { CPB_BODY_DYNAMIC = 0, CPB_BODY_KINEMATIC = 1, CPB_BODY_STATIC = 2 };
enum { CPB_SHAPE_CIRCLE = 0, CPB_SHAPE_SEGMENT = 1, CPB_SHAPE_POLY = 2 };
enum {
	CPB_JOINT_PIN = 0, CPB_JOINT_SLIDE = 1, CPB_JOINT_PIVOT = 2, CPB_JOINT_GROOVE = 3,
	CPB_JOINT_DAMPED_SPRING = 4, CPB_JOINT_DAMPED_ROTARY_SPRING = 5, CPB_JOINT_ROTARY_LIMIT = 6,
	CPB_JOINT_RATCHET = 7, CPB_JOINT_GEAR = 8, CPB_JOINT_SIMPLE_MOTOR = 9
};

typedef struct cpb_scene_header {
	uint64_t magic;
	int32_t n_bodies, n_shapes, n_verts, n_joints;
	/* space parameters (cpSpace.h:82-124 setters) */
	int32_t iterations;
	uint32_t collision_persistence;
	double gravity[2];
	double damping;
	double idle_speed_threshold;
	double sleep_time_threshold;   /* INFINITY = sleeping disabled */
	double collision_slop;
	double collision_bias;
	double timestep;               /* the demo's dt; informational */
} cpb_scene_header;

typedef struct cpb_scene_body {
	int32_t type;                  /* CPB_BODY_* */
	int32_t is_space_static;       /* 1 = this is cpSpaceGetStaticBody(space) */
	double m, i;
	double cog[2];
	double p[2], v[2], f[2];
	double a, w, t;
} cpb_scene_body;

typedef struct cpb_scene_shape {
	int32_t type;                  /* CPB_SHAPE_* */
	int32_t body;                  /* index into bodies[] */
	int32_t sensor;
	int32_t n_verts, vert_offset;  /* poly only: untransformed hull verts in verts[] */
	uint32_t categories, mask;
	int32_t _pad;
	uint64_t group;
	uint64_t collision_type;
	double e, u;
	double surface_v[2];
	double r;                      /* circle radius / segment radius / poly radius */
	double a[2], b[2];             /* circle: a = offset c; segment: endpoints a, b */
	double a_tangent[2], b_tangent[2]; /* segment neighbour tangents (cpShape.c:538-546) */
	double mass;                   /* shape massInfo.m (0 = body-level mass) */
} cpb_scene_shape;

typedef struct cpb_scene_joint {
	int32_t type;                  /* CPB_JOINT_* */
	int32_t a, b;                  /* body indices */
	int32_t collide_bodies;
	double max_force, error_bias, max_bias;
	double anchor_a[2], anchor_b[2];
	/* type-specific parameters:
	 *  pin: dist | slide: min,max | groove: grv_a(x,y)=anchor_a, grv_b = prm[0..1]
	 *  damped spring: restLength, stiffness, damping
	 *  damped rotary spring: restAngle, stiffness, damping
	 *  rotary limit: min,max | ratchet: angle(current), phase, ratchet
	 *  gear: phase, ratio | simple motor: rate */
	double prm[4];
	/* warm-start state carried with the scene (all zero for a fresh scene) */
	double acc[2];
} cpb_scene_joint;

static inline size_t cpb_scene_bytes(const cpb_scene_header *h){
	return sizeof(cpb_scene_header) + (size_t)h->n_bodies*sizeof(cpb_scene_body)
		+ (size_t)h->n_shapes*sizeof(cpb_scene_shape) + (size_t)h->n_verts*2*sizeof(double)
		+ (size_t)h->n_joints*sizeof(cpb_scene_joint);
}
static inline const cpb_scene_body *cpb_scene_bodies(const cpb_scene_header *h){
	return (const cpb_scene_body *)(h + 1);
}
static inline const cpb_scene_shape *cpb_scene_shapes(const cpb_scene_header *h){
	return (const cpb_scene_shape *)(cpb_scene_bodies(h) + h->n_bodies);
}
static inline const double *cpb_scene_verts(const cpb_scene_header *h){
	return (const double *)(cpb_scene_shapes(h) + h->n_shapes);
}
static inline const cpb_scene_joint *cpb_scene_joints(const cpb_scene_header *h){
	return (const cpb_scene_joint *)(cpb_scene_verts(h) + 2*(size_t)h->n_verts);
}

#ifdef __cplusplus
}
#endif
#endif
