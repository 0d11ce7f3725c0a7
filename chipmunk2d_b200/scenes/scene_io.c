/* scene_io.c -- instantiate a scene blob (cpb_scene.h) through the PUBLIC
 * Chipmunk2D C API only, and read results back through public getters.
 *
 * This file is compiled twice: once against the unmodified reference
 * (oracle/_ref/libscene_ref.so, headers from /root/reference/include) and once
 * against the B200 drop-in (chipmunk2d_b200/lib/libscene_b200.so, headers from
 * include/).  It therefore is also the proof that the drop-in boundary holds:
 * the same translation unit links and runs against both libraries.
 *
 * Every body/shape/constraint gets userData = (index + 1) so that state can be
 * reported in scene order regardless of the library's internal ordering.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <time.h>

#include "chipmunk/chipmunk.h"
#include "chipmunk/cpHastySpace.h"
#include "cpb_scene.h"

#ifndef CPB_EXPORT
#define CPB_EXPORT __attribute__((visibility("default")))
#endif

#define TAG(i) ((cpDataPointer)(uintptr_t)((i) + 1))
#define UNTAG(p) ((int)((uintptr_t)(p)) - 1)

static cpVect V(const double *d){ return cpv(d[0], d[1]); }

CPB_EXPORT cpSpace *
cpb_scene_load(const void *blob, int hasty, int hasty_threads)
{
	const cpb_scene_header *h = (const cpb_scene_header *)blob;
	if(h->magic != CPB_SCENE_MAGIC){
		fprintf(stderr, "cpb_scene_load: bad magic\n");
		return NULL;
	}
	const cpb_scene_body *sb = cpb_scene_bodies(h);
	const cpb_scene_shape *ss = cpb_scene_shapes(h);
	const double *sv = cpb_scene_verts(h);
	const cpb_scene_joint *sj = cpb_scene_joints(h);

	cpSpace *space;
	if(hasty){
		space = cpHastySpaceNew();
		cpHastySpaceSetThreads(space, (unsigned long)hasty_threads);
	} else {
		space = cpSpaceNew();
	}
	cpSpaceSetIterations(space, h->iterations);
	cpSpaceSetGravity(space, V(h->gravity));
	cpSpaceSetDamping(space, h->damping);
	cpSpaceSetIdleSpeedThreshold(space, h->idle_speed_threshold);
	cpSpaceSetSleepTimeThreshold(space, h->sleep_time_threshold);
	cpSpaceSetCollisionSlop(space, h->collision_slop);
	cpSpaceSetCollisionBias(space, h->collision_bias);
	cpSpaceSetCollisionPersistence(space, h->collision_persistence);

	cpBody **bodies = (cpBody **)calloc((size_t)h->n_bodies + 1, sizeof(cpBody *));
	for(int i = 0; i < h->n_bodies; i++){
		const cpb_scene_body *b = &sb[i];
		cpBody *body;
		if(b->is_space_static){
			body = cpSpaceGetStaticBody(space);
		} else if(b->type == CPB_BODY_STATIC){
			body = cpSpaceAddBody(space, cpBodyNewStatic());
		} else if(b->type == CPB_BODY_KINEMATIC){
			body = cpSpaceAddBody(space, cpBodyNewKinematic());
		} else {
			body = cpSpaceAddBody(space, cpBodyNew(b->m, b->i));
			if(b->cog[0] != 0.0 || b->cog[1] != 0.0) cpBodySetCenterOfGravity(body, V(b->cog));
		}
		if(!b->is_space_static || b->p[0] != 0.0 || b->p[1] != 0.0) cpBodySetPosition(body, V(b->p));
		if(b->a != 0.0) cpBodySetAngle(body, b->a);
		if(b->type != CPB_BODY_STATIC){
			cpBodySetVelocity(body, V(b->v));
			cpBodySetAngularVelocity(body, b->w);
		}
		if(b->type == CPB_BODY_DYNAMIC){
			cpBodySetForce(body, V(b->f));
			cpBodySetTorque(body, b->t);
		}
		cpBodySetUserData(body, TAG(i));
		bodies[i] = body;
	}

	for(int i = 0; i < h->n_shapes; i++){
		const cpb_scene_shape *s = &ss[i];
		cpBody *body = bodies[s->body];
		cpShape *shape = NULL;
		switch(s->type){
		case CPB_SHAPE_CIRCLE:
			shape = cpCircleShapeNew(body, s->r, V(s->a));
			break;
		case CPB_SHAPE_SEGMENT:
			shape = cpSegmentShapeNew(body, V(s->a), V(s->b), s->r);
			/* tangents are stored directly; reproduce them through the public setter
			 * only when non-zero (cpSegmentShapeSetNeighbors derives them from points) */
			break;
		case CPB_SHAPE_POLY: {
			cpVect *verts = (cpVect *)malloc(sizeof(cpVect)*(size_t)s->n_verts);
			for(int k = 0; k < s->n_verts; k++) verts[k] = V(sv + 2*((size_t)s->vert_offset + k));
			shape = cpPolyShapeNewRaw(body, s->n_verts, verts, s->r);
			free(verts);
			break;
		}
		default:
			fprintf(stderr, "cpb_scene_load: bad shape type %d\n", s->type);
			abort();
		}
		cpShapeSetElasticity(shape, s->e);
		cpShapeSetFriction(shape, s->u);
		cpShapeSetSurfaceVelocity(shape, V(s->surface_v));
		cpShapeSetSensor(shape, (cpBool)s->sensor);
		cpShapeSetCollisionType(shape, (cpCollisionType)s->collision_type);
		cpShapeFilter f = {(cpGroup)s->group, s->categories, s->mask};
		cpShapeSetFilter(shape, f);
		if(s->mass > 0.0) cpShapeSetMass(shape, s->mass);
		cpShapeSetUserData(shape, TAG(i));
		cpSpaceAddShape(space, shape);
	}

	for(int i = 0; i < h->n_joints; i++){
		const cpb_scene_joint *j = &sj[i];
		cpBody *a = bodies[j->a], *b = bodies[j->b];
		cpConstraint *c = NULL;
		switch(j->type){
		case CPB_JOINT_PIN:
			c = cpPinJointNew(a, b, V(j->anchor_a), V(j->anchor_b));
			cpPinJointSetDist(c, j->prm[0]);
			break;
		case CPB_JOINT_SLIDE:
			c = cpSlideJointNew(a, b, V(j->anchor_a), V(j->anchor_b), j->prm[0], j->prm[1]);
			break;
		case CPB_JOINT_PIVOT:
			c = cpPivotJointNew2(a, b, V(j->anchor_a), V(j->anchor_b));
			break;
		case CPB_JOINT_GROOVE:
			c = cpGrooveJointNew(a, b, V(j->anchor_a), V(j->prm), V(j->anchor_b));
			break;
		case CPB_JOINT_DAMPED_SPRING:
			c = cpDampedSpringNew(a, b, V(j->anchor_a), V(j->anchor_b), j->prm[0], j->prm[1], j->prm[2]);
			break;
		case CPB_JOINT_DAMPED_ROTARY_SPRING:
			c = cpDampedRotarySpringNew(a, b, j->prm[0], j->prm[1], j->prm[2]);
			break;
		case CPB_JOINT_ROTARY_LIMIT:
			c = cpRotaryLimitJointNew(a, b, j->prm[0], j->prm[1]);
			break;
		case CPB_JOINT_RATCHET:
			c = cpRatchetJointNew(a, b, j->prm[1], j->prm[2]);
			cpRatchetJointSetAngle(c, j->prm[0]);
			break;
		case CPB_JOINT_GEAR:
			c = cpGearJointNew(a, b, j->prm[0], j->prm[1]);
			break;
		case CPB_JOINT_SIMPLE_MOTOR:
			c = cpSimpleMotorNew(a, b, j->prm[0]);
			break;
		default:
			fprintf(stderr, "cpb_scene_load: bad joint type %d\n", j->type);
			abort();
		}
		cpConstraintSetMaxForce(c, j->max_force);
		cpConstraintSetErrorBias(c, j->error_bias);
		cpConstraintSetMaxBias(c, j->max_bias);
		cpConstraintSetCollideBodies(c, (cpBool)j->collide_bodies);
		cpConstraintSetUserData(c, TAG(i));
		cpSpaceAddConstraint(space, c);
	}

	free(bodies);
	return space;
}

/* ---- teardown (the space never owns its children: cpSpace.c:188-229) ---- */

typedef struct ptr_list { void **arr; int n, cap; } ptr_list;
static void push(ptr_list *l, void *p){
	if(l->n == l->cap){ l->cap = l->cap ? 2*l->cap : 1024; l->arr = (void **)realloc(l->arr, sizeof(void *)*(size_t)l->cap); }
	l->arr[l->n++] = p;
}
static void collect_body(cpBody *b, void *l){ push((ptr_list *)l, b); }
static void collect_shape(cpShape *s, void *l){ push((ptr_list *)l, s); }
static void collect_constraint(cpConstraint *c, void *l){ push((ptr_list *)l, c); }

CPB_EXPORT void
cpb_scene_free(cpSpace *space, int hasty)
{
	ptr_list bodies = {0}, shapes = {0}, constraints = {0};
	cpSpaceEachShape(space, collect_shape, &shapes);
	cpSpaceEachConstraint(space, collect_constraint, &constraints);
	cpSpaceEachBody(space, collect_body, &bodies);
	for(int i = 0; i < constraints.n; i++){ cpSpaceRemoveConstraint(space, (cpConstraint *)constraints.arr[i]); cpConstraintFree((cpConstraint *)constraints.arr[i]); }
	for(int i = 0; i < shapes.n; i++){ cpSpaceRemoveShape(space, (cpShape *)shapes.arr[i]); cpShapeFree((cpShape *)shapes.arr[i]); }
	for(int i = 0; i < bodies.n; i++){ cpSpaceRemoveBody(space, (cpBody *)bodies.arr[i]); cpBodyFree((cpBody *)bodies.arr[i]); }
	free(bodies.arr); free(shapes.arr); free(constraints.arr);
	if(hasty) cpHastySpaceFree(space); else cpSpaceFree(space);
}

/* ---- stepping / timing ---- */

CPB_EXPORT void
cpb_scene_step(cpSpace *space, double dt, int n, int hasty)
{
	for(int i = 0; i < n; i++){
		if(hasty) cpHastySpaceStep(space, dt); else cpSpaceStep(space, dt);
	}
}

/* Wall seconds for n steps, measured the way demo/ChipmunkDemo.c:521-538 does
 * (monotonic clock around the step loop only). */
CPB_EXPORT double
cpb_scene_time_steps(cpSpace *space, double dt, int n, int hasty)
{
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	cpb_scene_step(space, dt, n, hasty);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9*(double)(t1.tv_nsec - t0.tv_nsec);
}

/* ---- state read-back through public getters ---- */

#define CPB_BODY_STATE_DOUBLES 10
typedef struct body_dump { double *out; int n; } body_dump;

static void dump_body(cpBody *body, void *ctx){
	body_dump *d = (body_dump *)ctx;
	int i = UNTAG(cpBodyGetUserData(body));
	if(i < 0 || i >= d->n) return;
	double *o = d->out + (size_t)i*CPB_BODY_STATE_DOUBLES;
	cpVect p = cpBodyGetPosition(body), v = cpBodyGetVelocity(body), rot = cpBodyGetRotation(body);
	o[0] = p.x; o[1] = p.y; o[2] = v.x; o[3] = v.y;
	o[4] = cpBodyGetAngle(body); o[5] = cpBodyGetAngularVelocity(body);
	o[6] = rot.x; o[7] = rot.y;
	o[8] = (double)cpBodyIsSleeping(body);
	o[9] = cpBodyKineticEnergy(body);
}

/* out[n][10] = p.x p.y v.x v.y a w rot.x rot.y sleeping KE, rows in scene order.
 * Rows of bodies the space does not iterate (the built-in static body) stay untouched. */
CPB_EXPORT void
cpb_scene_get_bodies(cpSpace *space, int n, double *out)
{
	body_dump d = {out, n};
	cpSpaceEachBody(space, dump_body, &d);
}

#define CPB_SHAPE_STATE_DOUBLES 4
typedef struct shape_dump { double *out; int n; } shape_dump;
static void dump_shape(cpShape *shape, void *ctx){
	shape_dump *d = (shape_dump *)ctx;
	int i = UNTAG(cpShapeGetUserData(shape));
	if(i < 0 || i >= d->n) return;
	cpBB bb = cpShapeGetBB(shape);
	double *o = d->out + (size_t)i*CPB_SHAPE_STATE_DOUBLES;
	o[0] = bb.l; o[1] = bb.b; o[2] = bb.r; o[3] = bb.t;
}

/* out[n][4] = cached AABB (l b r t) of every shape, scene order. */
CPB_EXPORT void
cpb_scene_get_shape_bbs(cpSpace *space, int n, double *out)
{
	shape_dump d = {out, n};
	cpSpaceEachShape(space, dump_shape, &d);
}

/* Contact graph through cpBodyEachArbiter + the cpArbiter accessors
 * (cpArbiter.h:32-143).  One row per (body, arbiter) visit with body == arbiter's
 * first body, so every arbiter appears once.
 * row[16] = shapeA shapeB count normal.x normal.y
 *           (pointA.x pointA.y pointB.x pointB.y distance) x2  firstContact */
#define CPB_ARB_ROW 16
typedef struct arb_dump { double *out; int cap, n; } arb_dump;

static void dump_arbiter(cpBody *body, cpArbiter *arb, void *ctx){
	arb_dump *d = (arb_dump *)ctx;
	cpBody *ba, *bb;
	cpArbiterGetBodies(arb, &ba, &bb);
	/* each arbiter is threaded on both bodies; report it from the side that the
	 * accessor calls "a" after the per-body swap, and only for the lower-tagged
	 * dynamic owner to avoid duplicates */
	cpShape *sa, *sb;
	cpArbiterGetShapes(arb, &sa, &sb);
	int ia = UNTAG(cpShapeGetUserData(sa)), ib = UNTAG(cpShapeGetUserData(sb));
	int other_dynamic = (cpBodyGetType(bb) == CP_BODY_TYPE_DYNAMIC);
	if(other_dynamic && UNTAG(cpBodyGetUserData(bb)) < UNTAG(cpBodyGetUserData(ba))) return;
	if(d->n >= d->cap){ d->n++; return; }
	double *o = d->out + (size_t)d->n*CPB_ARB_ROW;
	memset(o, 0, sizeof(double)*CPB_ARB_ROW);
	cpContactPointSet set = cpArbiterGetContactPointSet(arb);
	o[0] = ia; o[1] = ib; o[2] = set.count; o[3] = set.normal.x; o[4] = set.normal.y;
	for(int k = 0; k < set.count && k < 2; k++){
		o[5 + 5*k + 0] = set.points[k].pointA.x; o[5 + 5*k + 1] = set.points[k].pointA.y;
		o[5 + 5*k + 2] = set.points[k].pointB.x; o[5 + 5*k + 3] = set.points[k].pointB.y;
		o[5 + 5*k + 4] = set.points[k].distance;
	}
	o[15] = (double)cpArbiterIsFirstContact(arb);
	d->n++;
	(void)body;
}

static void dump_body_arbiters(cpBody *body, void *ctx){
	if(cpBodyGetType(body) != CP_BODY_TYPE_DYNAMIC) return;
	cpBodyEachArbiter(body, dump_arbiter, ctx);
}

/* Returns the number of arbiters found (may exceed cap; only cap rows are written). */
CPB_EXPORT int
cpb_scene_get_arbiters(cpSpace *space, int cap, double *out)
{
	arb_dump d = {out, cap, 0};
	cpSpaceEachBody(space, dump_body_arbiters, &d);
	return d.n;
}

/* Narrowphase of one shape pair through the public cpShapesCollide (cpShape.c:259-283).
 * out[13] = count normal.x normal.y (pointA.xy pointB.xy distance) x2 */
typedef struct find_shape { int want; cpShape *found; } find_shape;
static void find_shape_cb(cpShape *s, void *ctx){
	find_shape *f = (find_shape *)ctx;
	if(UNTAG(cpShapeGetUserData(s)) == f->want) f->found = s;
}

CPB_EXPORT int
cpb_scene_shapes_collide(cpSpace *space, int ia, int ib, double *out)
{
	find_shape fa = {ia, NULL}, fb = {ib, NULL};
	cpSpaceEachShape(space, find_shape_cb, &fa);
	cpSpaceEachShape(space, find_shape_cb, &fb);
	if(!fa.found || !fb.found) return -1;
	cpContactPointSet set = cpShapesCollide(fa.found, fb.found);
	memset(out, 0, sizeof(double)*13);
	out[0] = set.count; out[1] = set.normal.x; out[2] = set.normal.y;
	for(int k = 0; k < set.count && k < 2; k++){
		out[3 + 5*k + 0] = set.points[k].pointA.x; out[3 + 5*k + 1] = set.points[k].pointA.y;
		out[3 + 5*k + 2] = set.points[k].pointB.x; out[3 + 5*k + 3] = set.points[k].pointB.y;
		out[3 + 5*k + 4] = set.points[k].distance;
	}
	return set.count;
}

/* ---- end-to-end loop for the benchmark: host buffers in, host buffers out, every step ----
 * Per step: write an external force into every dynamic body (host -> library), cpSpaceStep, read every
 * body's position back into out_xy[n][2] (library -> host).  Returns wall seconds for n_steps. */
typedef struct e2e_ctx { double fx, fy; double *out; int n; } e2e_ctx;
static void e2e_push(cpBody *body, void *ctx){
	e2e_ctx *c = (e2e_ctx *)ctx;
	if(cpBodyGetType(body) == CP_BODY_TYPE_DYNAMIC) cpBodySetForce(body, cpv(c->fx, c->fy));
}
static void e2e_pull(cpBody *body, void *ctx){
	e2e_ctx *c = (e2e_ctx *)ctx;
	int i = UNTAG(cpBodyGetUserData(body));
	if(i < 0 || i >= c->n) return;
	cpVect p = cpBodyGetPosition(body);
	c->out[2*(size_t)i] = p.x; c->out[2*(size_t)i + 1] = p.y;
}

CPB_EXPORT double
cpb_scene_e2e_steps(cpSpace *space, double dt, int n_steps, int n_bodies, double *out_xy, double fx, double fy, int hasty)
{
	struct timespec t0, t1;
	e2e_ctx ctx = {fx, fy, out_xy, n_bodies};
	clock_gettime(CLOCK_MONOTONIC, &t0);
	double t_push = 0.0, t_step = 0.0, t_pull = 0.0, t_fetch = 0.0;
	const int prof = (getenv("CPB_E2E_PROFILE") != NULL);
	for(int s = 0; s < n_steps; s++){
		struct timespec a, b, c, d;
		if(prof) clock_gettime(CLOCK_MONOTONIC, &a);
		cpSpaceEachBody(space, e2e_push, &ctx);
		if(prof) clock_gettime(CLOCK_MONOTONIC, &b);
		if(hasty) cpHastySpaceStep(space, dt); else cpSpaceStep(space, dt);
		if(prof){
			/* the first getter waits for the step, downloads the body state and refreshes the mirrors: time it apart */
			struct timespec c2;
			clock_gettime(CLOCK_MONOTONIC, &c);
			(void)cpBodyGetPosition(cpSpaceGetStaticBody(space));
			clock_gettime(CLOCK_MONOTONIC, &c2);
			t_fetch += (double)(c2.tv_sec - c.tv_sec) + 1e-9*(double)(c2.tv_nsec - c.tv_nsec);
		}
		cpSpaceEachBody(space, e2e_pull, &ctx);
		if(prof){
			clock_gettime(CLOCK_MONOTONIC, &d);
			t_push += (double)(b.tv_sec - a.tv_sec) + 1e-9*(double)(b.tv_nsec - a.tv_nsec);
			t_step += (double)(c.tv_sec - b.tv_sec) + 1e-9*(double)(c.tv_nsec - b.tv_nsec);
			t_pull += (double)(d.tv_sec - c.tv_sec) + 1e-9*(double)(d.tv_nsec - c.tv_nsec);
		}
	}
	if(prof) fprintf(stderr, "e2e per step: push %.2f ms  cpSpaceStep (upload + enqueue) %.2f ms  pull (wait + download + getters) %.2f ms, of which wait + download + mirror refresh %.2f ms\n",
		1e3*t_push/n_steps, 1e3*t_step/n_steps, 1e3*t_pull/n_steps, 1e3*t_fetch/n_steps);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9*(double)(t1.tv_nsec - t0.tv_nsec);
}

/* ---- space queries through the public API (cpSpaceQuery.c), same code against either library ----
 * Hits are reported by shape tag; rows are sorted by tag because the visiting order of the reference
 * depends on its tree shape.  Row layouts: point [tag px py distance gx gy], segment [tag px py nx ny alpha],
 * bb [tag], shape query [tag count nx ny (pA.xy pB.xy dist) x2]. */
typedef struct query_rows { double *out; int stride, cap, n; } query_rows;
static double *query_row(query_rows *q){ double *r = (q->n < q->cap ? q->out + (size_t)q->stride*q->n : NULL); q->n++; return r; }
static int cmp_row(const void *a, const void *b){ double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }
static void point_cb(cpShape *s, cpVect p, cpFloat d, cpVect g, void *ctx){
	double *r = query_row((query_rows *)ctx);
	if(r){ r[0] = UNTAG(cpShapeGetUserData(s)); r[1] = p.x; r[2] = p.y; r[3] = d; r[4] = g.x; r[5] = g.y; }
}
static void segment_cb(cpShape *s, cpVect p, cpVect n, cpFloat alpha, void *ctx){
	double *r = query_row((query_rows *)ctx);
	if(r){ r[0] = UNTAG(cpShapeGetUserData(s)); r[1] = p.x; r[2] = p.y; r[3] = n.x; r[4] = n.y; r[5] = alpha; }
}
static void bb_cb(cpShape *s, void *ctx){
	double *r = query_row((query_rows *)ctx);
	if(r) r[0] = UNTAG(cpShapeGetUserData(s));
}
static void shape_cb(cpShape *s, cpContactPointSet *set, void *ctx){
	double *r = query_row((query_rows *)ctx);
	if(!r) return;
	memset(r, 0, sizeof(double)*14);
	r[0] = UNTAG(cpShapeGetUserData(s)); r[1] = set->count; r[2] = set->normal.x; r[3] = set->normal.y;
	for(int k = 0; k < set->count && k < 2; k++){
		r[4 + 5*k] = set->points[k].pointA.x; r[5 + 5*k] = set->points[k].pointA.y;
		r[6 + 5*k] = set->points[k].pointB.x; r[7 + 5*k] = set->points[k].pointB.y; r[8 + 5*k] = set->points[k].distance;
	}
}
static int finish_rows(query_rows *q){ if(q->n <= q->cap) qsort(q->out, (size_t)q->n, sizeof(double)*(size_t)q->stride, cmp_row); return q->n; }

CPB_EXPORT int
cpb_scene_point_query(cpSpace *space, double x, double y, double max_dist, uint64_t group, uint32_t cat, uint32_t mask, int cap, double *out6)
{
	query_rows q = {out6, 6, cap, 0};
	cpSpacePointQuery(space, cpv(x, y), max_dist, cpShapeFilterNew((cpGroup)group, cat, mask), point_cb, &q);
	return finish_rows(&q);
}

/* out6 = [tag px py distance gx gy]; returns 1 if a shape was found */
CPB_EXPORT int
cpb_scene_point_query_nearest(cpSpace *space, double x, double y, double max_dist, uint64_t group, uint32_t cat, uint32_t mask, double *out6)
{
	cpPointQueryInfo info;
	cpShape *s = cpSpacePointQueryNearest(space, cpv(x, y), max_dist, cpShapeFilterNew((cpGroup)group, cat, mask), &info);
	out6[0] = (s ? UNTAG(cpShapeGetUserData(s)) : -1); out6[1] = info.point.x; out6[2] = info.point.y; out6[3] = info.distance;
	out6[4] = info.gradient.x; out6[5] = info.gradient.y;
	return s != NULL;
}

CPB_EXPORT int
cpb_scene_segment_query(cpSpace *space, double ax, double ay, double bx, double by, double radius, uint64_t group, uint32_t cat, uint32_t mask, int cap, double *out6)
{
	query_rows q = {out6, 6, cap, 0};
	cpSpaceSegmentQuery(space, cpv(ax, ay), cpv(bx, by), radius, cpShapeFilterNew((cpGroup)group, cat, mask), segment_cb, &q);
	return finish_rows(&q);
}

CPB_EXPORT int
cpb_scene_segment_query_first(cpSpace *space, double ax, double ay, double bx, double by, double radius, uint64_t group, uint32_t cat, uint32_t mask, double *out6)
{
	cpSegmentQueryInfo info;
	cpShape *s = cpSpaceSegmentQueryFirst(space, cpv(ax, ay), cpv(bx, by), radius, cpShapeFilterNew((cpGroup)group, cat, mask), &info);
	out6[0] = (s ? UNTAG(cpShapeGetUserData(s)) : -1); out6[1] = info.point.x; out6[2] = info.point.y; out6[3] = info.normal.x; out6[4] = info.normal.y;
	out6[5] = info.alpha;
	return s != NULL;
}

CPB_EXPORT int
cpb_scene_bb_query(cpSpace *space, double l, double b, double r, double t, uint64_t group, uint32_t cat, uint32_t mask, int cap, double *out1)
{
	query_rows q = {out1, 1, cap, 0};
	cpSpaceBBQuery(space, cpBBNew(l, b, r, t), cpShapeFilterNew((cpGroup)group, cat, mask), bb_cb, &q);
	return finish_rows(&q);
}

/* per-shape queries on the shape with the given tag: out6 as above; returns distance hit flag */
CPB_EXPORT int
cpb_scene_shape_point_query(cpSpace *space, int tag, double x, double y, double *out6)
{
	find_shape f = {tag, NULL};
	cpSpaceEachShape(space, find_shape_cb, &f);
	if(!f.found) return -1;
	cpPointQueryInfo info;
	cpShapePointQuery(f.found, cpv(x, y), &info);
	out6[0] = tag; out6[1] = info.point.x; out6[2] = info.point.y; out6[3] = info.distance; out6[4] = info.gradient.x; out6[5] = info.gradient.y;
	return 1;
}

CPB_EXPORT int
cpb_scene_shape_segment_query(cpSpace *space, int tag, double ax, double ay, double bx, double by, double radius, double *out6)
{
	find_shape f = {tag, NULL};
	cpSpaceEachShape(space, find_shape_cb, &f);
	if(!f.found) return -1;
	cpSegmentQueryInfo info;
	cpBool hit = cpShapeSegmentQuery(f.found, cpv(ax, ay), cpv(bx, by), radius, &info);
	out6[0] = (hit ? tag : -1); out6[1] = info.point.x; out6[2] = info.point.y; out6[3] = info.normal.x; out6[4] = info.normal.y; out6[5] = info.alpha;
	return hit;
}

/* cpSpaceShapeQuery with a probe that is NOT part of the space: kind 0 = circle(radius) at (x, y),
 * 1 = fat segment from (x, y) to (x + w, y + h) with `radius`, 2 = box w x h with bevel `radius`, rotated by angle.
 * Returns the hit count (rows sorted by tag); *any = the function's return value. */
CPB_EXPORT int
cpb_scene_shape_query(cpSpace *space, int kind, double x, double y, double angle, double w, double h, double radius, int cap, double *out14, int *any)
{
	cpBody *body = cpBodyNewKinematic();
	cpBodySetPosition(body, cpv(x, y));
	cpBodySetAngle(body, angle);
	cpShape *probe = (kind == 0 ? cpCircleShapeNew(body, radius, cpvzero)
	               : kind == 1 ? cpSegmentShapeNew(body, cpvzero, cpv(w, h), radius)
	                           : cpBoxShapeNew(body, w, h, radius));
	query_rows q = {out14, 14, cap, 0};
	cpBool r = cpSpaceShapeQuery(space, probe, shape_cb, &q);
	if(any) *any = r;
	cpShapeFree(probe);
	cpBodyFree(body);
	return finish_rows(&q);
}

/* ---- collision-handler semantics (cpSpaceStep.c:257-285, cpArbiter.c:46-50, 97-143), same code against
 * either library: one dynamic ball over static geometry, so the solver order cannot matter and the
 * trajectories of the two libraries agree to rounding.
 * out[12] = nBegin nPreSolve nPostSolve nSeparate  p.x p.y v.x v.y w  maxHeightAfterFirstContact  firstContactStep
 *           callbacks that found the arbiter user data set in begin (scenario 10) */
typedef struct hs_ctx { int which; int nBegin, nPre, nPost, nSep; int step, firstStep; int nData; } hs_ctx;
static cpBool hs_begin(cpArbiter *arb, cpSpace *space, void *data){
	hs_ctx *c = (hs_ctx *)data;
	c->nBegin++;
	if(c->firstStep < 0) c->firstStep = c->step;
	if(c->which == 2) return cpFalse;                 /* ignored until the shapes separate */
	if(c->which == 10) cpArbiterSetUserData(arb, data);   /* must still be there in every later callback of this pair */
	return cpTrue;
}
static cpBool hs_presolve(cpArbiter *arb, cpSpace *space, void *data){
	hs_ctx *c = (hs_ctx *)data;
	c->nPre++;
	if(c->which == 10 && cpArbiterGetUserData(arb) == data) c->nData++;
	switch(c->which){
	case 1: return cpFalse;                            /* never solved: the ball falls through */
	case 3: cpArbiterSetRestitution(arb, 1.0); break;  /* bouncy although both shapes have e = 0 */
	case 4: cpArbiterSetSurfaceVelocity(arb, cpv(50.0, 0.0)); cpArbiterSetFriction(arb, 1.0); break;   /* conveyor */
	case 6: if(cpArbiterGetNormal(arb).y > 0.0) return cpArbiterIgnore(arb); break;   /* one-way platform (demo/OneWay.c) */
	case 7: if(c->nPre == 10) return cpArbiterIgnore(arb); break;
	default: break;
	}
	return cpTrue;
}
static void hs_postsolve(cpArbiter *arb, cpSpace *space, void *data){ hs_ctx *c = (hs_ctx *)data; c->nPost++; if(c->which == 10 && cpArbiterGetUserData(arb) == data) c->nData++; }
static void hs_separate(cpArbiter *arb, cpSpace *space, void *data){ ((hs_ctx *)data)->nSep++; }

CPB_EXPORT int
cpb_scene_handler_scenario(int which, int n_steps, double *out12)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetIterations(space, 10);
	cpSpaceSetGravity(space, cpv(0.0, -100.0));
	cpSpaceSetCollisionSlop(space, 0.5);
	cpBody *sb = cpSpaceGetStaticBody(space);
	cpShape *ground = cpSpaceAddShape(space, cpSegmentShapeNew(sb, cpv(-400.0, 0.0), cpv(400.0, 0.0), 0.0));
	cpShapeSetElasticity(ground, 0.0); cpShapeSetFriction(ground, (which == 4 ? 0.0 : 0.7));
	cpShapeSetCollisionType(ground, 1);
	if(which == 5) cpShapeSetSensor(ground, cpTrue);
	cpShape *floor2 = NULL;
	if(which == 6){
		/* the platform is `ground`; a second floor far below catches nothing within the run */
		floor2 = cpSpaceAddShape(space, cpSegmentShapeNew(sb, cpv(-400.0, -1000.0), cpv(400.0, -1000.0), 0.0));
		cpShapeSetCollisionType(floor2, 3);
	}
	cpBody *ball = cpSpaceAddBody(space, cpBodyNew(1.0, cpMomentForCircle(1.0, 0.0, 10.0, cpvzero)));
	cpShape *bs = cpSpaceAddShape(space, cpCircleShapeNew(ball, 10.0, cpvzero));
	cpShapeSetElasticity(bs, 0.0); cpShapeSetFriction(bs, 0.7);
	cpShapeSetCollisionType(bs, 2);
	if(which == 6){ cpBodySetPosition(ball, cpv(0.0, -40.0)); cpBodySetVelocity(ball, cpv(0.0, 160.0)); }
	else cpBodySetPosition(ball, cpv(0.0, 50.0));

	hs_ctx ctx = {which, 0, 0, 0, 0, 0, -1, 0};
	cpCollisionHandler *h = (which == 8 ? cpSpaceAddWildcardHandler(space, 2)
	                       : which == 9 ? cpSpaceAddDefaultCollisionHandler(space)
	                                    : cpSpaceAddCollisionHandler(space, 1, 2));
	h->beginFunc = hs_begin; h->preSolveFunc = hs_presolve; h->postSolveFunc = hs_postsolve; h->separateFunc = hs_separate;
	h->userData = &ctx;

	double maxh = -1e300;
	for(int s = 0; s < n_steps; s++){
		ctx.step = s;
		cpSpaceStep(space, 1.0/60.0);
		if(ctx.firstStep >= 0 && s > ctx.firstStep + 2){ double y = cpBodyGetPosition(ball).y; if(y > maxh) maxh = y; }
	}
	cpVect p = cpBodyGetPosition(ball), v = cpBodyGetVelocity(ball);
	out12[0] = ctx.nBegin; out12[1] = ctx.nPre; out12[2] = ctx.nPost; out12[3] = ctx.nSep;
	out12[4] = p.x; out12[5] = p.y; out12[6] = v.x; out12[7] = v.y; out12[8] = cpBodyGetAngularVelocity(ball);
	out12[9] = (maxh > -1e299 ? maxh : 0.0); out12[10] = ctx.firstStep; out12[11] = ctx.nData;
	cpSpaceRemoveShape(space, bs); cpSpaceRemoveBody(space, ball); cpSpaceRemoveShape(space, ground);
	if(floor2){ cpSpaceRemoveShape(space, floor2); cpShapeFree(floor2); }
	cpShapeFree(bs); cpBodyFree(ball); cpShapeFree(ground);
	cpSpaceFree(space);
	return 0;
}
