/* ref_probe.c -- TEST INFRASTRUCTURE (oracle side).  Not part of the product.
 *
 * Compiled ONLY into oracle/_ref/libchipmunk_ref.so together with the unmodified
 * reference sources taken from where they lie under /root/reference (never
 * copied into this repository).  It reaches into the reference's private
 * structs (chipmunk_private.h / chipmunk_structs.h) to expose, as flat arrays,
 * the intermediate state the parity tests compare the CUDA path against:
 *
 *   - the demo scenes of BASELINE.json configs 1/2/5 built by the reference's own
 *     demo code (demo/Bench.c:462-480 bench_list, demo/PyramidStack.c:32-82,
 *     demo/Chains.c:56-134), flattened into a cpb_scene blob;
 *   - per-body solver state incl. v_bias/w_bias/idleTime (chipmunk_structs.h:35-81);
 *   - per-shape world cache + AABB (chipmunk_structs.h:177-236);
 *   - space->arbiters in solver order with every cpContact field
 *     (chipmunk_structs.h:101-145);
 *   - the overlapping-pair set: brute force over cached shape->bb with the
 *     membership and rejection rules of cpSpaceStep.c:204-232;
 *   - joint solver state (chipmunk_structs.h:250-382).
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>

#include "chipmunk/chipmunk_private.h"
#include "chipmunk/chipmunk_unsafe.h"
#include "ChipmunkDemo.h"
#include "cpb_scene.h"

#define REFP_EXPORT __attribute__((visibility("default")))
#define TAG(i) ((cpDataPointer)(uintptr_t)((i) + 1))
#define UNTAG(p) ((int)((uintptr_t)(p)) - 1)

/* ---- globals the demo sources expect (demo/ChipmunkDemo.h:54-68) ---- */
int ChipmunkDemoTicks = 0;
double ChipmunkDemoTime = 0.0;
cpVect ChipmunkDemoKeyboard = {0, 0};
cpVect ChipmunkDemoMouse = {0, 0};
cpBool ChipmunkDemoRightClick = cpFalse;
cpBool ChipmunkDemoRightDown = cpFalse;
char const *ChipmunkDemoMessageString = NULL;
#define REFP_GRABBABLE_MASK_BIT (1u<<31)
cpShapeFilter GRAB_FILTER = {CP_NO_GROUP, REFP_GRABBABLE_MASK_BIT, REFP_GRABBABLE_MASK_BIT};
cpShapeFilter NOT_GRABBABLE_FILTER = {CP_NO_GROUP, ~REFP_GRABBABLE_MASK_BIT, ~REFP_GRABBABLE_MASK_BIT};
void ChipmunkDemoPrintString(char const *fmt, ...){ (void)fmt; }
void ChipmunkDemoDefaultDrawImpl(cpSpace *space){ (void)space; }
void ChipmunkDemoFreeSpaceChildren(cpSpace *space){ (void)space; }

extern ChipmunkDemo bench_list[];
extern int bench_count;
extern ChipmunkDemo PyramidStack;
extern ChipmunkDemo Chains;

static ChipmunkDemo *find_demo(const char *name)
{
	if(strcmp(name, "PyramidStack") == 0) return &PyramidStack;
	if(strcmp(name, "Chains") == 0) return &Chains;
	for(int i = 0; i < bench_count; i++){
		const char *n = strstr(bench_list[i].name, "- ");
		n = (n ? n + 2 : bench_list[i].name);
		if(strcmp(n, name) == 0) return &bench_list[i];
	}
	return NULL;
}

REFP_EXPORT int refp_demo_count(void){ return bench_count + 2; }
REFP_EXPORT const char *refp_demo_name(int i)
{
	if(i < bench_count){
		const char *n = strstr(bench_list[i].name, "- ");
		return (n ? n + 2 : bench_list[i].name);
	}
	return (i == bench_count ? "PyramidStack" : "Chains");
}

/* srand(45073) is what RunDemo does before every initFunc (demo/ChipmunkDemo.c:378). */
REFP_EXPORT cpSpace *refp_demo_build(const char *name)
{
	ChipmunkDemo *demo = find_demo(name);
	if(!demo) return NULL;
	srand(45073);
	return demo->initFunc();
}

REFP_EXPORT double refp_demo_timestep(const char *name)
{
	ChipmunkDemo *demo = find_demo(name);
	return (demo ? demo->timestep : 0.0);
}

/* ---- flatten a cpSpace into a cpb_scene blob ---- */

typedef struct collector { void **arr; int n, cap; } collector;
static void coll_push(collector *c, void *p)
{
	if(c->n == c->cap){ c->cap = (c->cap ? 2*c->cap : 1024); c->arr = (void **)realloc(c->arr, sizeof(void *)*(size_t)c->cap); }
	c->arr[c->n++] = p;
}
static void coll_shape(cpShape *s, void *c){ coll_push((collector *)c, s); }
static void coll_constraint(cpConstraint *k, void *c){ coll_push((collector *)c, k); }
static void coll_body(cpBody *b, void *c){ coll_push((collector *)c, b); }

static int cmp_shape_hashid(const void *a, const void *b)
{
	cpHashValue ha = (*(cpShape *const *)a)->hashid, hb = (*(cpShape *const *)b)->hashid;
	return (ha < hb ? -1 : (ha > hb ? 1 : 0));
}

static int joint_type(cpConstraint *c)
{
	if(cpConstraintIsPinJoint(c)) return CPB_JOINT_PIN;
	if(cpConstraintIsSlideJoint(c)) return CPB_JOINT_SLIDE;
	if(cpConstraintIsPivotJoint(c)) return CPB_JOINT_PIVOT;
	if(cpConstraintIsGrooveJoint(c)) return CPB_JOINT_GROOVE;
	if(cpConstraintIsDampedSpring(c)) return CPB_JOINT_DAMPED_SPRING;
	if(cpConstraintIsDampedRotarySpring(c)) return CPB_JOINT_DAMPED_ROTARY_SPRING;
	if(cpConstraintIsRotaryLimitJoint(c)) return CPB_JOINT_ROTARY_LIMIT;
	if(cpConstraintIsRatchetJoint(c)) return CPB_JOINT_RATCHET;
	if(cpConstraintIsGearJoint(c)) return CPB_JOINT_GEAR;
	if(cpConstraintIsSimpleMotor(c)) return CPB_JOINT_SIMPLE_MOTOR;
	return -1;
}

static void put2(double *d, cpVect v){ d[0] = v.x; d[1] = v.y; }

/* Writes the blob into buf if cap is large enough; always returns the size needed.
 * Side effect: tags userData of every body/shape/constraint with its scene index + 1
 * (the Bench/PyramidStack/Chains demos do not use userData). */
REFP_EXPORT size_t refp_scene_dump(cpSpace *space, double timestep, void *buf, size_t cap)
{
	collector bodies = {0}, shapes = {0}, joints = {0};
	/* body 0 is always the space's built-in static body */
	coll_push(&bodies, cpSpaceGetStaticBody(space));
	cpSpaceEachBody(space, coll_body, &bodies);
	cpSpaceEachShape(space, coll_shape, &shapes);
	cpSpaceEachConstraint(space, coll_constraint, &joints);
	qsort(shapes.arr, (size_t)shapes.n, sizeof(void *), cmp_shape_hashid);

	for(int i = 0; i < bodies.n; i++) ((cpBody *)bodies.arr[i])->userData = TAG(i);

	int n_verts = 0;
	for(int i = 0; i < shapes.n; i++){
		cpShape *s = (cpShape *)shapes.arr[i];
		if(s->klass->type == CP_POLY_SHAPE) n_verts += ((cpPolyShape *)s)->count;
	}

	cpb_scene_header h;
	memset(&h, 0, sizeof(h));
	h.magic = CPB_SCENE_MAGIC;
	h.n_bodies = bodies.n; h.n_shapes = shapes.n; h.n_verts = n_verts; h.n_joints = joints.n;
	h.iterations = space->iterations;
	h.collision_persistence = space->collisionPersistence;
	put2(h.gravity, space->gravity);
	h.damping = space->damping;
	h.idle_speed_threshold = space->idleSpeedThreshold;
	h.sleep_time_threshold = space->sleepTimeThreshold;
	h.collision_slop = space->collisionSlop;
	h.collision_bias = space->collisionBias;
	h.timestep = timestep;

	size_t need = cpb_scene_bytes(&h);
	if(buf && cap >= need){
		memset(buf, 0, need);
		memcpy(buf, &h, sizeof(h));
		cpb_scene_header *hh = (cpb_scene_header *)buf;
		cpb_scene_body *ob = (cpb_scene_body *)cpb_scene_bodies(hh);
		cpb_scene_shape *os = (cpb_scene_shape *)cpb_scene_shapes(hh);
		double *ov = (double *)cpb_scene_verts(hh);
		cpb_scene_joint *oj = (cpb_scene_joint *)cpb_scene_joints(hh);

		for(int i = 0; i < bodies.n; i++){
			cpBody *b = (cpBody *)bodies.arr[i];
			cpBodyType t = cpBodyGetType(b);
			ob[i].type = (t == CP_BODY_TYPE_STATIC ? CPB_BODY_STATIC : (t == CP_BODY_TYPE_KINEMATIC ? CPB_BODY_KINEMATIC : CPB_BODY_DYNAMIC));
			ob[i].is_space_static = (i == 0);
			ob[i].m = b->m; ob[i].i = b->i;
			put2(ob[i].cog, b->cog); put2(ob[i].p, b->p); put2(ob[i].v, b->v); put2(ob[i].f, b->f);
			ob[i].a = b->a; ob[i].w = b->w; ob[i].t = b->t;
		}

		int voff = 0;
		for(int i = 0; i < shapes.n; i++){
			cpShape *s = (cpShape *)shapes.arr[i];
			s->userData = TAG(i);
			cpb_scene_shape *o = &os[i];
			o->body = UNTAG(s->body->userData);
			o->sensor = s->sensor;
			o->categories = s->filter.categories; o->mask = s->filter.mask; o->group = (uint64_t)s->filter.group;
			o->collision_type = (uint64_t)s->type;
			o->e = s->e; o->u = s->u; put2(o->surface_v, s->surfaceV);
			o->mass = s->massInfo.m;
			switch(s->klass->type){
			case CP_CIRCLE_SHAPE: {
				cpCircleShape *c = (cpCircleShape *)s;
				o->type = CPB_SHAPE_CIRCLE; o->r = c->r; put2(o->a, c->c);
				break;
			}
			case CP_SEGMENT_SHAPE: {
				cpSegmentShape *g = (cpSegmentShape *)s;
				o->type = CPB_SHAPE_SEGMENT; o->r = g->r; put2(o->a, g->a); put2(o->b, g->b);
				put2(o->a_tangent, g->a_tangent); put2(o->b_tangent, g->b_tangent);
				break;
			}
			case CP_POLY_SHAPE: {
				cpPolyShape *p = (cpPolyShape *)s;
				o->type = CPB_SHAPE_POLY; o->r = p->r; o->n_verts = p->count; o->vert_offset = voff;
				for(int k = 0; k < p->count; k++) put2(ov + 2*(size_t)(voff + k), p->planes[p->count + k].v0);
				voff += p->count;
				break;
			}
			default: break;
			}
		}

		for(int i = 0; i < joints.n; i++){
			cpConstraint *c = (cpConstraint *)joints.arr[i];
			c->userData = TAG(i);
			cpb_scene_joint *o = &oj[i];
			o->type = joint_type(c);
			o->a = UNTAG(c->a->userData); o->b = UNTAG(c->b->userData);
			o->collide_bodies = c->collideBodies;
			o->max_force = c->maxForce; o->error_bias = c->errorBias; o->max_bias = c->maxBias;
			switch(o->type){
			case CPB_JOINT_PIN: { cpPinJoint *j = (cpPinJoint *)c; put2(o->anchor_a, j->anchorA); put2(o->anchor_b, j->anchorB); o->prm[0] = j->dist; o->acc[0] = j->jnAcc; break; }
			case CPB_JOINT_SLIDE: { cpSlideJoint *j = (cpSlideJoint *)c; put2(o->anchor_a, j->anchorA); put2(o->anchor_b, j->anchorB); o->prm[0] = j->min; o->prm[1] = j->max; o->acc[0] = j->jnAcc; break; }
			case CPB_JOINT_PIVOT: { cpPivotJoint *j = (cpPivotJoint *)c; put2(o->anchor_a, j->anchorA); put2(o->anchor_b, j->anchorB); put2(o->acc, j->jAcc); break; }
			case CPB_JOINT_GROOVE: { cpGrooveJoint *j = (cpGrooveJoint *)c; put2(o->anchor_a, j->grv_a); put2(o->prm, j->grv_b); put2(o->anchor_b, j->anchorB); put2(o->acc, j->jAcc); break; }
			case CPB_JOINT_DAMPED_SPRING: { cpDampedSpring *j = (cpDampedSpring *)c; put2(o->anchor_a, j->anchorA); put2(o->anchor_b, j->anchorB); o->prm[0] = j->restLength; o->prm[1] = j->stiffness; o->prm[2] = j->damping; break; }
			case CPB_JOINT_DAMPED_ROTARY_SPRING: { cpDampedRotarySpring *j = (cpDampedRotarySpring *)c; o->prm[0] = j->restAngle; o->prm[1] = j->stiffness; o->prm[2] = j->damping; break; }
			case CPB_JOINT_ROTARY_LIMIT: { cpRotaryLimitJoint *j = (cpRotaryLimitJoint *)c; o->prm[0] = j->min; o->prm[1] = j->max; o->acc[0] = j->jAcc; break; }
			case CPB_JOINT_RATCHET: { cpRatchetJoint *j = (cpRatchetJoint *)c; o->prm[0] = j->angle; o->prm[1] = j->phase; o->prm[2] = j->ratchet; o->acc[0] = j->jAcc; break; }
			case CPB_JOINT_GEAR: { cpGearJoint *j = (cpGearJoint *)c; o->prm[0] = j->phase; o->prm[1] = j->ratio; o->acc[0] = j->jAcc; break; }
			case CPB_JOINT_SIMPLE_MOTOR: { cpSimpleMotor *j = (cpSimpleMotor *)c; o->prm[0] = j->rate; o->acc[0] = j->jAcc; break; }
			default: break;
			}
		}
	}

	free(bodies.arr); free(shapes.arr); free(joints.arr);
	return need;
}

/* ---- private state read-back ---- */

#define REFP_BODY_ROW 24
typedef struct row_dump { double *out; int n; } row_dump;

static void body_row(cpBody *b, void *ctx)
{
	row_dump *d = (row_dump *)ctx;
	int i = UNTAG(b->userData);
	if(i < 0 || i >= d->n) return;
	double *o = d->out + (size_t)i*REFP_BODY_ROW;
	o[0] = b->p.x; o[1] = b->p.y; o[2] = b->v.x; o[3] = b->v.y; o[4] = b->a; o[5] = b->w;
	o[6] = b->v_bias.x; o[7] = b->v_bias.y; o[8] = b->w_bias;
	o[9] = b->f.x; o[10] = b->f.y; o[11] = b->t;
	o[12] = b->transform.a; o[13] = b->transform.b; o[14] = b->transform.c; o[15] = b->transform.d;
	o[16] = b->transform.tx; o[17] = b->transform.ty;
	o[18] = b->sleeping.idleTime; o[19] = (double)(b->sleeping.root != NULL);
	o[20] = b->m_inv; o[21] = b->i_inv; o[22] = b->cog.x; o[23] = b->cog.y;
}

/* out[n][24]: p v a w v_bias w_bias f t transform(a b c d tx ty) idleTime sleeping m_inv i_inv cog */
REFP_EXPORT void refp_get_bodies(cpSpace *space, int n, double *out)
{
	row_dump d = {out, n};
	body_row(cpSpaceGetStaticBody(space), &d);
	cpSpaceEachBody(space, body_row, &d);
}

#define REFP_SHAPE_ROW 10
static void shape_row(cpShape *s, void *ctx)
{
	row_dump *d = (row_dump *)ctx;
	int i = UNTAG(s->userData);
	if(i < 0 || i >= d->n) return;
	double *o = d->out + (size_t)i*REFP_SHAPE_ROW;
	o[0] = s->bb.l; o[1] = s->bb.b; o[2] = s->bb.r; o[3] = s->bb.t;
	switch(s->klass->type){
	case CP_CIRCLE_SHAPE: { cpCircleShape *c = (cpCircleShape *)s; o[4] = c->tc.x; o[5] = c->tc.y; break; }
	case CP_SEGMENT_SHAPE: { cpSegmentShape *g = (cpSegmentShape *)s; o[4] = g->ta.x; o[5] = g->ta.y; o[6] = g->tb.x; o[7] = g->tb.y; o[8] = g->tn.x; o[9] = g->tn.y; break; }
	default: break;
	}
}

/* out[n][10]: bb(l b r t) then circle tc / segment ta tb tn (poly planes: refp_get_poly_planes) */
REFP_EXPORT void refp_get_shapes(cpSpace *space, int n, double *out)
{
	row_dump d = {out, n};
	cpSpaceEachShape(space, shape_row, &d);
}

typedef struct plane_dump { double *out; const int *vert_offset; int n; } plane_dump;
static void plane_row(cpShape *s, void *ctx)
{
	plane_dump *d = (plane_dump *)ctx;
	int i = UNTAG(s->userData);
	if(i < 0 || i >= d->n || s->klass->type != CP_POLY_SHAPE) return;
	cpPolyShape *p = (cpPolyShape *)s;
	double *o = d->out + 4*(size_t)d->vert_offset[i];
	for(int k = 0; k < p->count; k++){
		o[4*k + 0] = p->planes[k].v0.x; o[4*k + 1] = p->planes[k].v0.y;
		o[4*k + 2] = p->planes[k].n.x;  o[4*k + 3] = p->planes[k].n.y;
	}
}

/* out[n_verts][4] = world-space (v0.x v0.y n.x n.y) at each poly's vert_offset (scene blob order) */
REFP_EXPORT void refp_get_poly_planes(cpSpace *space, int n_shapes, const int *vert_offset, double *out)
{
	plane_dump d = {out, vert_offset, n_shapes};
	cpSpaceEachShape(space, plane_row, &d);
}

/* space->arbiters in SOLVER ORDER (cpSpaceStep.c:274,418-421).
 * row[36] = shapeA shapeB count state n.x n.y e u surface_vr.x surface_vr.y stamp swapped
 *           then per contact (12 each, x2): r1.x r1.y r2.x r2.y nMass tMass bounce jnAcc jtAcc jBias bias hash_lo32
 * (the upper 32 hash bits are returned separately in hash_hi if non-NULL: [2*i + k]) */
#define REFP_ARB_ROW 36
REFP_EXPORT int refp_get_arbiters(cpSpace *space, int cap, double *out, unsigned int *hash_hi)
{
	cpArray *arbs = space->arbiters;
	for(int i = 0; i < arbs->num && i < cap; i++){
		cpArbiter *arb = (cpArbiter *)arbs->arr[i];
		double *o = out + (size_t)i*REFP_ARB_ROW;
		memset(o, 0, sizeof(double)*REFP_ARB_ROW);
		o[0] = UNTAG(arb->a->userData); o[1] = UNTAG(arb->b->userData);
		o[2] = arb->count; o[3] = arb->state; o[4] = arb->n.x; o[5] = arb->n.y;
		o[6] = arb->e; o[7] = arb->u; o[8] = arb->surface_vr.x; o[9] = arb->surface_vr.y;
		o[10] = arb->stamp; o[11] = arb->swapped;
		for(int k = 0; k < arb->count && k < 2; k++){
			struct cpContact *c = &arb->contacts[k];
			double *q = o + 12 + 12*k;
			q[0] = c->r1.x; q[1] = c->r1.y; q[2] = c->r2.x; q[3] = c->r2.y;
			q[4] = c->nMass; q[5] = c->tMass; q[6] = c->bounce;
			q[7] = c->jnAcc; q[8] = c->jtAcc; q[9] = c->jBias; q[10] = c->bias;
			q[11] = (double)(unsigned int)(c->hash & 0xffffffffu);
			if(hash_hi) hash_hi[2*i + k] = (unsigned int)(c->hash >> 32);
		}
	}
	return arbs->num;
}

/* Joint solver state in scene order.
 * row[12] = r1.x r1.y r2.x r2.y n.x n.y nMass bias(.x) bias.y jAcc(.x) jAcc.y impulse */
#define REFP_JOINT_ROW 12
typedef struct joint_dump { double *out; int n; } joint_dump;
static void joint_row(cpConstraint *c, void *ctx)
{
	joint_dump *d = (joint_dump *)ctx;
	int i = UNTAG(c->userData);
	if(i < 0 || i >= d->n) return;
	double *o = d->out + (size_t)i*REFP_JOINT_ROW;
	memset(o, 0, sizeof(double)*REFP_JOINT_ROW);
	switch(joint_type(c)){
	case CPB_JOINT_PIN: { cpPinJoint *j = (cpPinJoint *)c; o[0]=j->r1.x;o[1]=j->r1.y;o[2]=j->r2.x;o[3]=j->r2.y;o[4]=j->n.x;o[5]=j->n.y;o[6]=j->nMass;o[7]=j->bias;o[9]=j->jnAcc; break; }
	case CPB_JOINT_SLIDE: { cpSlideJoint *j = (cpSlideJoint *)c; o[0]=j->r1.x;o[1]=j->r1.y;o[2]=j->r2.x;o[3]=j->r2.y;o[4]=j->n.x;o[5]=j->n.y;o[6]=j->nMass;o[7]=j->bias;o[9]=j->jnAcc; break; }
	case CPB_JOINT_PIVOT: { cpPivotJoint *j = (cpPivotJoint *)c; o[0]=j->r1.x;o[1]=j->r1.y;o[2]=j->r2.x;o[3]=j->r2.y;o[7]=j->bias.x;o[8]=j->bias.y;o[9]=j->jAcc.x;o[10]=j->jAcc.y; break; }
	case CPB_JOINT_DAMPED_SPRING: { cpDampedSpring *j = (cpDampedSpring *)c; o[0]=j->r1.x;o[1]=j->r1.y;o[2]=j->r2.x;o[3]=j->r2.y;o[4]=j->n.x;o[5]=j->n.y;o[6]=j->nMass;o[7]=j->v_coef;o[8]=j->target_vrn;o[9]=j->jAcc; break; }
	case CPB_JOINT_GEAR: { cpGearJoint *j = (cpGearJoint *)c; o[6]=j->iSum;o[7]=j->bias;o[9]=j->jAcc; break; }
	case CPB_JOINT_ROTARY_LIMIT: { cpRotaryLimitJoint *j = (cpRotaryLimitJoint *)c; o[6]=j->iSum;o[7]=j->bias;o[9]=j->jAcc; break; }
	case CPB_JOINT_RATCHET: { cpRatchetJoint *j = (cpRatchetJoint *)c; o[6]=j->iSum;o[7]=j->bias;o[9]=j->jAcc;o[8]=j->angle; break; }
	case CPB_JOINT_SIMPLE_MOTOR: { cpSimpleMotor *j = (cpSimpleMotor *)c; o[6]=j->iSum;o[9]=j->jAcc; break; }
	case CPB_JOINT_DAMPED_ROTARY_SPRING: { cpDampedRotarySpring *j = (cpDampedRotarySpring *)c; o[6]=j->iSum;o[7]=j->w_coef;o[8]=j->target_wrn;o[9]=j->jAcc; break; }
	case CPB_JOINT_GROOVE: { cpGrooveJoint *j = (cpGrooveJoint *)c; o[0]=j->r1.x;o[1]=j->r1.y;o[2]=j->r2.x;o[3]=j->r2.y;o[4]=j->grv_tn.x;o[5]=j->grv_tn.y;o[6]=j->clamp;o[7]=j->bias.x;o[8]=j->bias.y;o[9]=j->jAcc.x;o[10]=j->jAcc.y; break; }
	default: break;
	}
	o[11] = cpConstraintGetImpulse(c);
}

REFP_EXPORT void refp_get_joints(cpSpace *space, int n, double *out)
{
	joint_dump d = {out, n};
	cpSpaceEachConstraint(space, joint_row, &d);
}

/* ---- the overlapping-pair set ----
 * A ranges over shapes of awake non-static bodies, B over all shapes; a pair is
 * emitted once (as min<<32|max of the scene shape indices) when it survives the
 * QueryReject rules of cpSpaceStep.c:219-232.  `asleep` (per body, scene order,
 * may be NULL) gives the sleeping flags as they were at collision time. */
static int reject_constraint(cpBody *a, cpBody *b)
{
	CP_BODY_FOREACH_CONSTRAINT(a, c){
		if(!c->collideBodies && ((c->a == a && c->b == b) || (c->a == b && c->b == a))) return 1;
	}
	return 0;
}

static int cmp_u64(const void *a, const void *b)
{
	uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
	return (x < y ? -1 : (x > y ? 1 : 0));
}

static int cmp_shape_left(const void *a, const void *b)
{
	cpFloat x = (*(cpShape *const *)a)->bb.l, y = (*(cpShape *const *)b)->bb.l;
	return (x < y ? -1 : (x > y ? 1 : 0));
}

/* Every pair of shapes is examined whose closed x-intervals overlap (shapes sorted by bb.l, inner loop until
 * b.l > a.r): the same set an all-pairs double loop finds, in O(n log n + candidates) -- a 50 000-circle pile is
 * 2.5e9 box tests per step otherwise. */
REFP_EXPORT long refp_pairs_bruteforce(cpSpace *space, const unsigned char *asleep, long cap, uint64_t *out)
{
	collector shapes = {0};
	cpSpaceEachShape(space, coll_shape, &shapes);
	qsort(shapes.arr, (size_t)shapes.n, sizeof(void *), cmp_shape_left);
	unsigned char *active = (unsigned char *)malloc((size_t)shapes.n + 1);
	for(int i = 0; i < shapes.n; i++){
		cpShape *a = (cpShape *)shapes.arr[i];
		active[i] = (cpBodyGetType(a->body) != CP_BODY_TYPE_STATIC) && !(asleep ? asleep[UNTAG(a->body->userData)] : cpBodyIsSleeping(a->body));
	}
	long n = 0;
	for(int i = 0; i < shapes.n; i++){
		cpShape *a = (cpShape *)shapes.arr[i];
		int ia = UNTAG(a->userData);
		for(int j = i + 1; j < shapes.n; j++){
			cpShape *b = (cpShape *)shapes.arr[j];
			if(b->bb.l > a->bb.r) break;
			int ib = UNTAG(b->userData);
			if(!active[i] && !active[j]) continue;
			if(!cpBBIntersects(a->bb, b->bb)) continue;
			if(a->body == b->body) continue;
			if(cpShapeFilterReject(a->filter, b->filter)) continue;
			if(reject_constraint(a->body, b->body)) continue;
			if(n < cap){
				uint64_t lo = (uint64_t)(ia < ib ? ia : ib), hi = (uint64_t)(ia < ib ? ib : ia);
				out[n] = (lo << 32) | hi;
			}
			n++;
		}
	}
	free(active);
	free(shapes.arr);
	if(n <= cap) qsort(out, (size_t)n, sizeof(uint64_t), cmp_u64);
	return n;
}

/* Overwrite kinematic state of every tagged body (for feeding the same state to both
 * libraries mid-simulation).  in[n][6] = p.x p.y v.x v.y a w ; static bodies are skipped. */
REFP_EXPORT void refp_step(cpSpace *space, double dt, int n)
{
	for(int i = 0; i < n; i++) cpSpaceStep(space, dt);
}

/* space->constraints in SOLVER ORDER (cpSpaceStep.c:423-426) as scene joint indices.  Sleeping
 * and waking reorder this array (cpArrayDeleteObj moves the last element into the hole). */
REFP_EXPORT int refp_get_constraint_order(cpSpace *space, int cap, int *out)
{
	cpArray *cons = space->constraints;
	for(int i = 0; i < cons->num && i < cap; i++) out[i] = UNTAG(((cpConstraint *)cons->arr[i])->userData);
	return cons->num;
}

REFP_EXPORT int refp_space_counts(cpSpace *space, int *out)
{
	out[0] = space->dynamicBodies->num;
	out[1] = space->staticBodies->num;
	out[2] = space->arbiters->num;
	out[3] = space->constraints->num;
	out[4] = (int)space->stamp;
	out[5] = space->sleepingComponents->num;
	int contacts = 0;
	for(int i = 0; i < space->arbiters->num; i++) contacts += ((cpArbiter *)space->arbiters->arr[i])->count;
	out[6] = contacts;
	return 7;
}

/* ---- solver-order hook: let the UNMODIFIED reference solve one step in a caller-given order ----
 * cpSpaceStep walks space->arbiters and space->constraints in array order (cpSpaceStep.c:406-427).  Both arrays
 * are complete once the collision phase and cpSpaceProcessComponents are over, and the first thing the reference
 * calls after that through a public function pointer is every dynamic body's velocity_func (cpSpaceStep.c:398-404).
 * refp_install_order_hook points that pointer of every body at a wrapper which, once per step, permutes the two
 * arrays into the order set by refp_set_solver_order and then calls the stock cpBodyUpdateVelocity.  No reference
 * arithmetic changes: a Gauss-Seidel sweep is run over the same constraints in another order -- the order the
 * device's coloured solver used -- so that one production step can be compared with the reference itself. */
typedef struct order_key { uint64_t key; int rank; uint64_t hash0; } order_key;
static struct {
	cpSpace *space;
	int pending;
	int n_arb; order_key *arb;      /* sorted by key */
	int n_con; int *con_rank;       /* con_rank[scene joint index] */
	int n_con_cap;
	int applied, unmatched;
} g_order;

static int cmp_order_key(const void *a, const void *b)
{
	uint64_t x = ((const order_key *)a)->key, y = ((const order_key *)b)->key;
	return (x < y ? -1 : (x > y ? 1 : 0));
}

typedef struct ranked_ptr { void *p; long rank; int idx; } ranked_ptr;
static int cmp_ranked(const void *a, const void *b)
{
	const ranked_ptr *x = (const ranked_ptr *)a, *y = (const ranked_ptr *)b;
	if(x->rank != y->rank) return (x->rank < y->rank ? -1 : 1);
	return (x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0));
}

static void order_apply(cpSpace *space)
{
	cpArray *arbs = space->arbiters;
	ranked_ptr *tmp = (ranked_ptr *)malloc(sizeof(ranked_ptr)*(size_t)(arbs->num + space->constraints->num + 1));
	g_order.unmatched = 0;
	for(int i = 0; i < arbs->num; i++){
		cpArbiter *arb = (cpArbiter *)arbs->arr[i];
		uint64_t ta = (uint64_t)(uint32_t)UNTAG(arb->a->userData), tb = (uint64_t)(uint32_t)UNTAG(arb->b->userData);
		order_key probe; probe.key = (ta < tb ? (ta << 32) | tb : (tb << 32) | ta);
		order_key *hit = (order_key *)bsearch(&probe, g_order.arb, (size_t)g_order.n_arb, sizeof(order_key), cmp_order_key);
		tmp[i].p = arb; tmp[i].idx = i; tmp[i].rank = (hit ? hit->rank : 0x7fffffffL);
		if(!hit) g_order.unmatched++;
		/* visit the two contacts in the caller's order as well */
		if(hit && arb->count == 2 && arb->contacts[0].hash != hit->hash0 && arb->contacts[1].hash == hit->hash0){
			struct cpContact c = arb->contacts[0]; arb->contacts[0] = arb->contacts[1]; arb->contacts[1] = c;
		}
	}
	qsort(tmp, (size_t)arbs->num, sizeof(ranked_ptr), cmp_ranked);
	for(int i = 0; i < arbs->num; i++) arbs->arr[i] = tmp[i].p;
	cpArray *cons = space->constraints;
	for(int i = 0; i < cons->num; i++){
		int tag = UNTAG(((cpConstraint *)cons->arr[i])->userData);
		tmp[i].p = cons->arr[i]; tmp[i].idx = i;
		tmp[i].rank = (tag >= 0 && tag < g_order.n_con && g_order.con_rank[tag] >= 0 ? g_order.con_rank[tag] : 0x7fffffffL);
	}
	qsort(tmp, (size_t)cons->num, sizeof(ranked_ptr), cmp_ranked);
	for(int i = 0; i < cons->num; i++) cons->arr[i] = tmp[i].p;
	free(tmp);
	g_order.applied++;
}

static void hook_velocity(cpBody *body, cpVect gravity, cpFloat damping, cpFloat dt)
{
	if(g_order.pending && body->space == g_order.space){ g_order.pending = 0; order_apply(body->space); }
	cpBodyUpdateVelocity(body, gravity, damping, dt);
}

static void hook_body(cpBody *b, void *unused){ (void)unused; if(b->velocity_func == cpBodyUpdateVelocity) b->velocity_func = hook_velocity; }

REFP_EXPORT void refp_install_order_hook(cpSpace *space){ cpSpaceEachBody(space, hook_body, NULL); }

/* pairs[n_arb] = (shape tag a)<<32 | (shape tag b) in the wanted solver order, hash0[n_arb] = hash of the contact to
 * visit first (0 = leave the contact order alone); cons[n_con] = scene joint indices in the wanted order.
 * Applies to the NEXT cpSpaceStep of `space` only.  Arbiters / constraints not listed keep their relative order
 * after the listed ones. */
REFP_EXPORT void refp_set_solver_order(cpSpace *space, int n_arb, const uint64_t *pairs, const uint64_t *hash0, int n_con, const int *cons, int n_joints_total)
{
	g_order.space = space;
	g_order.arb = (order_key *)realloc(g_order.arb, sizeof(order_key)*(size_t)(n_arb + 1));
	for(int i = 0; i < n_arb; i++){
		uint64_t a = pairs[i] >> 32, b = pairs[i] & 0xffffffffu;
		g_order.arb[i].key = (a < b ? (a << 32) | b : (b << 32) | a);
		g_order.arb[i].rank = i;
		g_order.arb[i].hash0 = (hash0 ? hash0[i] : 0);
	}
	g_order.n_arb = n_arb;
	qsort(g_order.arb, (size_t)n_arb, sizeof(order_key), cmp_order_key);
	g_order.con_rank = (int *)realloc(g_order.con_rank, sizeof(int)*(size_t)(n_joints_total + 1));
	for(int i = 0; i < n_joints_total; i++) g_order.con_rank[i] = -1;
	for(int i = 0; i < n_con; i++) if(cons[i] >= 0 && cons[i] < n_joints_total) g_order.con_rank[cons[i]] = i;
	g_order.n_con = n_joints_total;
	g_order.pending = 1;
}

/* out[0] = times the order was applied, out[1] = arbiters of the last application that the caller had not listed */
REFP_EXPORT void refp_order_hook_stats(int *out){ out[0] = g_order.applied; out[1] = g_order.unmatched; }
