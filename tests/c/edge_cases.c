/* edge_cases.c -- corner cases of the step path through the public C API only, so the same file runs
 * against the unmodified reference (oracle/_ref) and against the B200 drop-in and the outputs are compared
 * line by line (tests/test_gpu_edges.py).  The cases are the ones the reference's own suite and demos lean
 * on (xcode/ObjectiveChipmunkTests/SpaceTest.m, BodyTest.m, ShapeTest.m): empty and
 * degenerate spaces, coincident shapes (cpCollision.c:530-534 fallback normal), sensors, filters, kinematic
 * movers, structural edits and teleports between steps, explicit sleep / wake.
 *
 * Output: one line per observation, "<case> <E|A|I> <values...>".  E = must be bit-identical (values printed
 * with %a), A = approximate (1e-9 relative; the value depends on sin/cos or on more than one contact),
 * I = an invariant the program checked itself (prints 1 when it holds).
 * Every scenario keeps at most ONE contact per dynamic body (or none), so the order in which a solver
 * visits the arbiters cannot change the result. */
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "chipmunk/chipmunk.h"

static void body_line(const char *name, const char *kind, cpBody *b)
{
	cpVect p = cpBodyGetPosition(b), v = cpBodyGetVelocity(b);
	printf("%s %s %a %a %a %a %a %a\n", name, kind, p.x, p.y, v.x, v.y, cpBodyGetAngle(b), cpBodyGetAngularVelocity(b));
}

static cpBody *add_ball(cpSpace *space, cpVect pos, cpFloat r, cpFloat m, cpShape **shape_out)
{
	cpBody *b = cpSpaceAddBody(space, cpBodyNew(m, cpMomentForCircle(m, 0, r, cpvzero)));
	cpBodySetPosition(b, pos);
	cpShape *s = cpSpaceAddShape(space, cpCircleShapeNew(b, r, cpvzero));
	cpShapeSetFriction(s, 0.7); cpShapeSetElasticity(s, 0.2);
	if(shape_out) *shape_out = s;
	return b;
}

/* 1. a space with nothing in it steps (cpSpaceStep.c:335-445 with every array empty) */
static void empty_space(void)
{
	cpSpace *space = cpSpaceNew();
	for(int i = 0; i < 3; i++) cpSpaceStep(space, 1.0/60.0);
	printf("empty_space E %a\n", cpSpaceGetCurrentTimeStep(space));
	cpSpaceStep(space, 0.0);   /* dt == 0 returns at once (cpSpaceStep.c:338) */
	printf("empty_space_dt0 E %a\n", cpSpaceGetCurrentTimeStep(space));
	cpSpaceFree(space);
}

/* 2. bodies without shapes: integrators only, with damping and per-body force / torque */
static void bodies_without_shapes(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetGravity(space, cpv(0.5, -9.75));
	cpSpaceSetDamping(space, 0.9);
	cpBody *b[3];
	for(int i = 0; i < 3; i++){
		b[i] = cpSpaceAddBody(space, cpBodyNew(1.0 + i, 2.0 + i));
		cpBodySetPosition(b[i], cpv(10.0*i, 1.0));
		cpBodySetVelocity(b[i], cpv(1.0 - i, 0.25*i));
		cpBodySetAngularVelocity(b[i], 0.5*i);
	}
	for(int s = 0; s < 5; s++){
		cpBodySetForce(b[1], cpv(3.0, 4.0));
		cpBodySetTorque(b[2], -1.5);
		cpSpaceStep(space, 1.0/60.0);
	}
	for(int i = 0; i < 3; i++){ body_line("bodies_without_shapes", "E", b[i]); }
	printf("bodies_without_shapes_force E %a %a\n", cpBodyGetForce(b[1]).x, cpBodyGetTorque(b[2]));
	for(int i = 0; i < 3; i++){ cpSpaceRemoveBody(space, b[i]); cpBodyFree(b[i]); }
	cpSpaceFree(space);
}

/* 3. static geometry only, then the first dynamic body arrives */
static void static_only_then_one_ball(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetGravity(space, cpv(0, -100));
	cpShape *g = cpSpaceAddShape(space, cpSegmentShapeNew(cpSpaceGetStaticBody(space), cpv(-50, 0), cpv(50, 0), 0));
	cpShapeSetFriction(g, 1.0);
	cpSpaceStep(space, 1.0/60.0);
	cpSpaceStep(space, 1.0/60.0);
	cpShape *bs; cpBody *ball = add_ball(space, cpv(0, 5.2), 5.0, 1.0, &bs);
	for(int s = 0; s < 20; s++) cpSpaceStep(space, 1.0/60.0);
	body_line("static_only_then_one_ball", "E", ball);
	cpSpaceRemoveShape(space, bs); cpSpaceRemoveBody(space, ball); cpSpaceRemoveShape(space, g);
	cpShapeFree(bs); cpBodyFree(ball); cpShapeFree(g);
	cpSpaceFree(space);
}

/* 4. two circles on exactly the same point: distance 0, the normal falls back to (1, 0) FROM THE FIRST SHAPE OF
 * THE PAIR TO THE SECOND.  Which shape the reference's index reports first depends on the shape of its BBTree
 * (cpBBTree.c:425-480) and differs for cpSpaceHash, so the SIGN of the push is implementation-defined; the
 * magnitudes, the opposite directions and everything else are not. */
static void coincident_circles(void)
{
	cpSpace *space = cpSpaceNew();
	cpShape *sa, *sb;
	cpBody *a = add_ball(space, cpv(3, 4), 2.0, 1.0, &sa), *b = add_ball(space, cpv(3, 4), 1.0, 2.0, &sb);
	cpContactPointSet set = cpShapesCollide(sa, sb);
	printf("coincident_circles_set E %d %a %a %a\n", set.count, set.normal.x, set.normal.y, set.points[0].distance);
	for(int s = 0; s < 3; s++) cpSpaceStep(space, 1.0/60.0);
	cpVect pa = cpBodyGetPosition(a), pb = cpBodyGetPosition(b), va = cpBodyGetVelocity(a), vb = cpBodyGetVelocity(b);
	printf("coincident_circles A %a %a %a %a %a %a\n", fabs(pa.x - 3.0), fabs(pb.x - 3.0), pa.y, pb.y, fabs(va.x), fabs(vb.x));
	printf("coincident_circles_opposite I %d %d\n", (pa.x - 3.0)*(pb.x - 3.0) < 0.0, va.y == 0.0 && vb.y == 0.0);
	cpSpaceRemoveShape(space, sa); cpSpaceRemoveShape(space, sb); cpSpaceRemoveBody(space, a); cpSpaceRemoveBody(space, b);
	cpShapeFree(sa); cpShapeFree(sb); cpBodyFree(a); cpBodyFree(b);
	cpSpaceFree(space);
}

/* 5. sensors and filters produce no response; a sensor still reports its overlap to cpShapesCollide */
static void sensors_and_filters(void)
{
	cpSpace *space = cpSpaceNew();
	cpShape *s[4];
	cpBody *b0 = add_ball(space, cpv(0, 0), 2.0, 1.0, &s[0]), *b1 = add_ball(space, cpv(1, 0), 2.0, 1.0, &s[1]);
	cpBody *b2 = add_ball(space, cpv(20, 0), 2.0, 1.0, &s[2]), *b3 = add_ball(space, cpv(21, 0), 2.0, 1.0, &s[3]);
	cpShapeSetSensor(s[1], cpTrue);
	cpShapeSetFilter(s[2], cpShapeFilterNew(7, CP_ALL_CATEGORIES, CP_ALL_CATEGORIES));
	cpShapeSetFilter(s[3], cpShapeFilterNew(7, CP_ALL_CATEGORIES, CP_ALL_CATEGORIES));
	for(int k = 0; k < 3; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("sensors_and_filters", "E", b0); body_line("sensors_and_filters", "E", b1);
	body_line("sensors_and_filters", "E", b2); body_line("sensors_and_filters", "E", b3);
	printf("sensors_and_filters_set E %d\n", cpShapesCollide(s[0], s[1]).count);
	/* category / mask rejection */
	cpShapeSetSensor(s[1], cpFalse);
	cpShapeSetFilter(s[0], cpShapeFilterNew(CP_NO_GROUP, 1, 2));
	cpShapeSetFilter(s[1], cpShapeFilterNew(CP_NO_GROUP, 1, 2));
	for(int k = 0; k < 3; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("sensors_and_filters_mask", "E", b0); body_line("sensors_and_filters_mask", "E", b1);
	cpBody *bs[4] = {b0, b1, b2, b3};
	for(int i = 0; i < 4; i++){ cpSpaceRemoveShape(space, s[i]); cpSpaceRemoveBody(space, bs[i]); cpShapeFree(s[i]); cpBodyFree(bs[i]); }
	cpSpaceFree(space);
}

/* 6. a kinematic pusher (infinite mass, moved by its velocity) shoves a free ball */
static void kinematic_pusher(void)
{
	cpSpace *space = cpSpaceNew();
	cpBody *kin = cpSpaceAddBody(space, cpBodyNewKinematic());
	cpBodySetPosition(kin, cpv(-10, 0));
	cpBodySetVelocity(kin, cpv(30, 0));
	cpShape *ks = cpSpaceAddShape(space, cpCircleShapeNew(kin, 3.0, cpvzero));
	cpShape *bs; cpBody *ball = add_ball(space, cpv(0, 0), 2.0, 1.0, &bs);
	for(int s = 0; s < 30; s++) cpSpaceStep(space, 1.0/60.0);
	body_line("kinematic_pusher", "E", kin); body_line("kinematic_pusher", "E", ball);
	cpSpaceRemoveShape(space, ks); cpSpaceRemoveShape(space, bs); cpSpaceRemoveBody(space, kin); cpSpaceRemoveBody(space, ball);
	cpShapeFree(ks); cpShapeFree(bs); cpBodyFree(kin); cpBodyFree(ball);
	cpSpaceFree(space);
}

/* 7. structural edits and teleports between steps: three balls, each on its own ground segment, far apart */
static void edits_between_steps(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetGravity(space, cpv(0, -100));
	cpShape *g[3], *s[3]; cpBody *b[3];
	for(int i = 0; i < 3; i++){
		g[i] = cpSpaceAddShape(space, cpSegmentShapeNew(cpSpaceGetStaticBody(space), cpv(100.0*i - 20, 0), cpv(100.0*i + 20, 0), 1.0));
		cpShapeSetFriction(g[i], 1.0);
	}
	for(int i = 0; i < 2; i++) b[i] = add_ball(space, cpv(100.0*i, 8.0), 5.0, 1.0 + i, &s[i]);
	for(int k = 0; k < 15; k++) cpSpaceStep(space, 1.0/60.0);
	/* add a third ball, teleport the first, give the second a kick */
	b[2] = add_ball(space, cpv(200.0, 6.5), 5.0, 3.0, &s[2]);
	cpBodySetPosition(b[0], cpv(5.0, 12.0)); cpSpaceReindexShapesForBody(space, b[0]);
	cpBodySetVelocity(b[1], cpv(4.0, 0.0));
	for(int k = 0; k < 15; k++) cpSpaceStep(space, 1.0/60.0);
	for(int i = 0; i < 3; i++) body_line("edits_between_steps_a", "E", b[i]);
	/* remove the second ball and its ground; the others keep their cached contacts */
	cpSpaceRemoveShape(space, s[1]); cpSpaceRemoveBody(space, b[1]); cpSpaceRemoveShape(space, g[1]);
	for(int k = 0; k < 15; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("edits_between_steps_b", "E", b[0]); body_line("edits_between_steps_b", "E", b[2]);
	printf("edits_between_steps_contains E %d %d %d\n", cpSpaceContainsBody(space, b[1]), cpSpaceContainsShape(space, s[0]), cpSpaceContainsShape(space, g[1]));
	cpShapeFree(s[1]); cpBodyFree(b[1]); cpShapeFree(g[1]);
	for(int i = 0; i < 3; i += 2){ cpSpaceRemoveShape(space, s[i]); cpSpaceRemoveBody(space, b[i]); cpSpaceRemoveShape(space, g[i]); cpShapeFree(s[i]); cpBodyFree(b[i]); cpShapeFree(g[i]); }
	cpSpaceFree(space);
}

/* 8. explicit sleep and wake (cpSpaceComponent.c:309-349): a sleeping body ignores gravity until activated */
static void explicit_sleep(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetGravity(space, cpv(0, -10));
	cpSpaceSetSleepTimeThreshold(space, 0.5);
	cpShape *s0, *s1; cpBody *a = add_ball(space, cpv(0, 50), 1.0, 1.0, &s0), *b = add_ball(space, cpv(30, 50), 1.0, 1.0, &s1);
	cpSpaceStep(space, 1.0/60.0);
	cpBodySleep(a);
	for(int k = 0; k < 10; k++) cpSpaceStep(space, 1.0/60.0);
	printf("explicit_sleep_flags E %d %d\n", cpBodyIsSleeping(a), cpBodyIsSleeping(b));
	body_line("explicit_sleep_asleep", "E", a); body_line("explicit_sleep_asleep", "E", b);
	cpBodyActivate(a);
	for(int k = 0; k < 10; k++) cpSpaceStep(space, 1.0/60.0);
	printf("explicit_sleep_flags2 E %d %d\n", cpBodyIsSleeping(a), cpBodyIsSleeping(b));
	body_line("explicit_sleep_awake", "E", a); body_line("explicit_sleep_awake", "E", b);
	cpSpaceRemoveShape(space, s0); cpSpaceRemoveShape(space, s1); cpSpaceRemoveBody(space, a); cpSpaceRemoveBody(space, b);
	cpShapeFree(s0); cpShapeFree(s1); cpBodyFree(a); cpBodyFree(b);
	cpSpaceFree(space);
}

/* 9. one body with far more contacts than a colour mask has bits: 90 small balls overlap one heavy ball in free
 * space.  Every arbiter shares the big ball, so the result depends on the solver's order: check invariants
 * (momentum is conserved by every impulse pair; everything stays finite; the overlap is being pushed out). */
static void hub_with_many_contacts(void)
{
	enum { N = 90 };
	cpSpace *space = cpSpaceNew();
	cpSpaceSetIterations(space, 10);
	cpShape *hs; cpBody *hub = add_ball(space, cpv(0, 0), 50.0, 100.0, &hs);
	cpShape *s[N]; cpBody *b[N];
	for(int i = 0; i < N; i++){
		cpFloat ang = 2.0*3.14159265358979323846*i/N;
		b[i] = add_ball(space, cpv(50.5*cos(ang), 50.5*sin(ang)), 1.0, 1.0, &s[i]);   /* 0.5 deep into the hub, apart from each other */
		cpShapeSetFriction(s[i], 0.0);
	}
	cpShapeSetFriction(hs, 0.0);
	for(int k = 0; k < 20; k++) cpSpaceStep(space, 1.0/60.0);
	cpVect mom = cpvmult(cpBodyGetVelocity(hub), cpBodyGetMass(hub));
	int finite = isfinite(cpBodyGetPosition(hub).x) && isfinite(cpBodyGetPosition(hub).y);
	cpFloat min_dist = INFINITY;
	for(int i = 0; i < N; i++){
		mom = cpvadd(mom, cpvmult(cpBodyGetVelocity(b[i]), cpBodyGetMass(b[i])));
		cpVect p = cpBodyGetPosition(b[i]);
		finite = finite && isfinite(p.x) && isfinite(p.y);
		min_dist = cpfmin(min_dist, cpvdist(p, cpBodyGetPosition(hub)));
	}
	printf("hub_with_many_contacts I %d %d %d\n", finite, cpvlength(mom) < 1e-9, min_dist > 50.5 && min_dist < 51.5);
	for(int i = 0; i < N; i++){ cpSpaceRemoveShape(space, s[i]); cpSpaceRemoveBody(space, b[i]); cpShapeFree(s[i]); cpBodyFree(b[i]); }
	cpSpaceRemoveShape(space, hs); cpSpaceRemoveBody(space, hub); cpShapeFree(hs); cpBodyFree(hub);
	cpSpaceFree(space);
}

/* 10. a box resting on a segment: two contacts on one body, rotation stays ~0 (A: sin/cos of a tiny angle) */
static void resting_box(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetGravity(space, cpv(0, -100));
	cpShape *g = cpSpaceAddShape(space, cpSegmentShapeNew(cpSpaceGetStaticBody(space), cpv(-50, 0), cpv(50, 0), 0));
	cpShapeSetFriction(g, 1.0);
	cpBody *box = cpSpaceAddBody(space, cpBodyNew(2.0, cpMomentForBox(2.0, 10, 10)));
	cpBodySetPosition(box, cpv(0, 5.3));
	cpShape *bs = cpSpaceAddShape(space, cpBoxShapeNew(box, 10, 10, 0.0));
	cpShapeSetFriction(bs, 0.6);
	for(int k = 0; k < 60; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("resting_box", "A", box);
	cpSpaceRemoveShape(space, bs); cpSpaceRemoveBody(space, box); cpSpaceRemoveShape(space, g);
	cpShapeFree(bs); cpBodyFree(box); cpShapeFree(g);
	cpSpaceFree(space);
}

/* ---- part 2: parameters and objects edited while the simulation runs (one contact / joint per body) ---- */

static cpSpace *ground_space(cpShape **ground)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetGravity(space, cpv(0, -100));
	*ground = cpSpaceAddShape(space, cpSegmentShapeNew(cpSpaceGetStaticBody(space), cpv(-200, 0), cpv(200, 0), 0));
	cpShapeSetFriction(*ground, 1.0); cpShapeSetElasticity(*ground, 1.0);
	return space;
}

/* 11. the time step changes every step: warm starting scales by dt/prev_dt (cpSpaceStep.c:407) */
static void changing_dt(void)
{
	cpShape *g, *bs; cpSpace *space = ground_space(&g);
	cpBody *ball = add_ball(space, cpv(0, 5.5), 5.0, 1.0, &bs);
	const cpFloat dts[4] = {1.0/60.0, 1.0/120.0, 1.0/30.0, 1.0/90.0};
	for(int k = 0; k < 40; k++) cpSpaceStep(space, dts[k & 3]);
	body_line("changing_dt", "E", ball);
	cpSpaceRemoveShape(space, bs); cpSpaceRemoveBody(space, ball); cpSpaceRemoveShape(space, g);
	cpShapeFree(bs); cpBodyFree(ball); cpShapeFree(g); cpSpaceFree(space);
}

/* 12. space parameters edited between steps: gravity, damping, iterations, slop, bias, persistence */
static void space_parameters_midrun(void)
{
	cpShape *g, *bs; cpSpace *space = ground_space(&g);
	cpBody *ball = add_ball(space, cpv(0, 6.0), 5.0, 1.0, &bs);
	for(int k = 0; k < 10; k++) cpSpaceStep(space, 1.0/60.0);
	cpSpaceSetGravity(space, cpv(20, -50)); cpSpaceSetDamping(space, 0.8); cpSpaceSetIterations(space, 3);
	for(int k = 0; k < 10; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("space_parameters_midrun_a", "E", ball);
	cpSpaceSetCollisionSlop(space, 0.5); cpSpaceSetCollisionBias(space, 0.01); cpSpaceSetCollisionPersistence(space, 1);
	for(int k = 0; k < 10; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("space_parameters_midrun_b", "E", ball);
	cpSpaceRemoveShape(space, bs); cpSpaceRemoveBody(space, ball); cpSpaceRemoveShape(space, g);
	cpShapeFree(bs); cpBodyFree(ball); cpShapeFree(g); cpSpaceFree(space);
}

/* 13. body edits between steps: mass, impulse at a point, force at a point, type to static and back */
static void body_edits_midrun(void)
{
	cpShape *g, *s0, *s1; cpSpace *space = ground_space(&g);
	cpBody *a = add_ball(space, cpv(-50, 5.5), 5.0, 1.0, &s0), *b = add_ball(space, cpv(50, 5.5), 5.0, 2.0, &s1);
	for(int k = 0; k < 10; k++) cpSpaceStep(space, 1.0/60.0);
	cpBodySetMass(a, 4.0); cpBodySetMoment(a, cpMomentForCircle(4.0, 0, 5.0, cpvzero));
	cpBodyApplyImpulseAtWorldPoint(b, cpv(3.0, 0.0), cpvadd(cpBodyGetPosition(b), cpv(0, 5.0)));
	for(int k = 0; k < 10; k++){
		cpBodyApplyForceAtLocalPoint(a, cpv(5.0, 0.0), cpv(0.0, 2.0));
		cpSpaceStep(space, 1.0/60.0);
	}
	body_line("body_edits_midrun_a", "E", a); body_line("body_edits_midrun_a", "A", b);
	cpBodySetType(b, CP_BODY_TYPE_STATIC);
	for(int k = 0; k < 5; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("body_edits_midrun_static", "A", b);
	cpBodySetType(b, CP_BODY_TYPE_DYNAMIC);
	cpBodySetMass(b, 2.0); cpBodySetMoment(b, cpMomentForCircle(2.0, 0, 5.0, cpvzero));
	for(int k = 0; k < 10; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("body_edits_midrun_b", "E", a); body_line("body_edits_midrun_b", "A", b);
	cpSpaceRemoveShape(space, s0); cpSpaceRemoveShape(space, s1); cpSpaceRemoveBody(space, a); cpSpaceRemoveBody(space, b); cpSpaceRemoveShape(space, g);
	cpShapeFree(s0); cpShapeFree(s1); cpBodyFree(a); cpBodyFree(b); cpShapeFree(g); cpSpaceFree(space);
}

/* 14. shape edits between steps: friction / elasticity / surface velocity, unsafe radius change, remove + re-add */
static void shape_edits_midrun(void)
{
	cpShape *g, *bs; cpSpace *space = ground_space(&g);
	cpBody *ball = add_ball(space, cpv(0, 5.5), 5.0, 1.0, &bs);
	for(int k = 0; k < 10; k++) cpSpaceStep(space, 1.0/60.0);
	cpShapeSetSurfaceVelocity(g, cpv(15.0, 0.0)); cpShapeSetFriction(bs, 0.3);
	for(int k = 0; k < 10; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("shape_edits_midrun_conveyor", "E", ball);
	cpCircleShapeSetRadius(bs, 6.0);
	for(int k = 0; k < 10; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("shape_edits_midrun_radius", "E", ball);
	cpSpaceRemoveShape(space, bs); cpSpaceAddShape(space, bs);    /* a new hashid: no warm start from the old arbiter */
	for(int k = 0; k < 10; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("shape_edits_midrun_readd", "E", ball);
	cpSegmentShapeSetEndpoints(g, cpv(-200, -2), cpv(200, -2)); cpSpaceReindexStatic(space);
	for(int k = 0; k < 10; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("shape_edits_midrun_ground", "E", ball);
	cpSpaceRemoveShape(space, bs); cpSpaceRemoveBody(space, ball); cpSpaceRemoveShape(space, g);
	cpShapeFree(bs); cpBodyFree(ball); cpShapeFree(g); cpSpaceFree(space);
}

/* 15. a perfectly elastic drop: bounce uses the velocities from before the velocity update (cpArbiter.c:436) */
static void elastic_bounce(void)
{
	cpShape *g, *bs; cpSpace *space = ground_space(&g);
	cpBody *ball = add_ball(space, cpv(0, 30.0), 5.0, 1.0, &bs);
	cpShapeSetElasticity(bs, 1.0);
	for(int k = 0; k < 90; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("elastic_bounce", "E", ball);
	cpSpaceRemoveShape(space, bs); cpSpaceRemoveBody(space, ball); cpSpaceRemoveShape(space, g);
	cpShapeFree(bs); cpBodyFree(ball); cpShapeFree(g); cpSpaceFree(space);
}

/* 16. joints added, edited and removed while running: every body carries one joint to the static body */
static void joints_midrun(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetGravity(space, cpv(0, -100));
	cpBody *st = cpSpaceGetStaticBody(space), *b[4];
	for(int i = 0; i < 4; i++){
		b[i] = cpSpaceAddBody(space, cpBodyNew(1.0 + i, 10.0 + i));
		cpBodySetPosition(b[i], cpv(100.0*i + 10.0, 0.0));
	}
	cpConstraint *pin = cpSpaceAddConstraint(space, cpPinJointNew(st, b[0], cpv(0, 0), cpv(0, 0)));
	cpConstraint *spring = cpSpaceAddConstraint(space, cpDampedSpringNew(st, b[1], cpv(100, 0), cpv(0, 0), 5.0, 40.0, 0.5));
	for(int k = 0; k < 20; k++) cpSpaceStep(space, 1.0/60.0);
	cpConstraint *pivot = cpSpaceAddConstraint(space, cpPivotJointNew(st, b[2], cpv(200, 0)));
	cpConstraint *slide = cpSpaceAddConstraint(space, cpSlideJointNew(st, b[3], cpv(300, 0), cpv(0, 0), 5.0, 15.0));
	cpDampedSpringSetStiffness(spring, 80.0); cpPinJointSetDist(pin, 12.0);
	for(int k = 0; k < 20; k++) cpSpaceStep(space, 1.0/60.0);
	for(int i = 0; i < 4; i++) body_line("joints_midrun_a", "A", b[i]);
	printf("joints_midrun_impulse A %a %a %a %a\n", cpConstraintGetImpulse(pin), cpConstraintGetImpulse(spring), cpConstraintGetImpulse(pivot), cpConstraintGetImpulse(slide));
	cpSpaceRemoveConstraint(space, pin); cpConstraintSetMaxForce(pivot, 50.0); cpConstraintSetMaxBias(slide, 1.0); cpConstraintSetErrorBias(slide, 0.5);
	for(int k = 0; k < 20; k++) cpSpaceStep(space, 1.0/60.0);
	for(int i = 0; i < 4; i++) body_line("joints_midrun_b", "A", b[i]);
	cpSpaceRemoveConstraint(space, spring); cpSpaceRemoveConstraint(space, pivot); cpSpaceRemoveConstraint(space, slide);
	cpConstraintFree(pin); cpConstraintFree(spring); cpConstraintFree(pivot); cpConstraintFree(slide);
	for(int i = 0; i < 4; i++){ cpSpaceRemoveBody(space, b[i]); cpBodyFree(b[i]); }
	cpSpaceFree(space);
}

/* 17. a body removed and added back later keeps what its cpBody holds; one that sleeps through an edit stays asleep */
static void remove_and_return(void)
{
	cpShape *g, *s0, *s1; cpSpace *space = ground_space(&g);
	cpSpaceSetSleepTimeThreshold(space, 0.2);
	cpBody *a = add_ball(space, cpv(-50, 5.5), 5.0, 1.0, &s0), *b = add_ball(space, cpv(50, 5.5), 5.0, 2.0, &s1);
	for(int k = 0; k < 60; k++) cpSpaceStep(space, 1.0/60.0);
	printf("remove_and_return_sleeping E %d %d\n", cpBodyIsSleeping(a), cpBodyIsSleeping(b));
	cpSpaceRemoveShape(space, s0); cpSpaceRemoveBody(space, a);
	for(int k = 0; k < 5; k++) cpSpaceStep(space, 1.0/60.0);
	printf("remove_and_return_sleeping2 E %d\n", cpBodyIsSleeping(b));
	body_line("remove_and_return_out", "E", a);
	cpSpaceAddBody(space, a); cpSpaceAddShape(space, s0);
	cpBodySetVelocity(a, cpv(0, 20));
	for(int k = 0; k < 30; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("remove_and_return_back", "E", a); body_line("remove_and_return_back", "E", b);
	printf("remove_and_return_sleeping3 E %d %d\n", cpBodyIsSleeping(a), cpBodyIsSleeping(b));
	cpSpaceRemoveShape(space, s0); cpSpaceRemoveShape(space, s1); cpSpaceRemoveBody(space, a); cpSpaceRemoveBody(space, b); cpSpaceRemoveShape(space, g);
	cpShapeFree(s0); cpShapeFree(s1); cpBodyFree(a); cpBodyFree(b); cpShapeFree(g); cpSpaceFree(space);
}

/* ---- part 3: polygons and fat segments (one arbiter with two contacts per body: still order-free) ---- */

/* 18. a capsule (fat segment) dropped on the ground segment: segment-segment narrowphase, then a kick */
static void capsule_on_ground(void)
{
	cpShape *g; cpSpace *space = ground_space(&g);
	cpBody *cap = cpSpaceAddBody(space, cpBodyNew(2.0, cpMomentForSegment(2.0, cpv(-8, 0), cpv(8, 0), 2.0)));
	cpBodySetPosition(cap, cpv(0, 4.0));
	cpShape *cs = cpSpaceAddShape(space, cpSegmentShapeNew(cap, cpv(-8, 0), cpv(8, 0), 2.0));
	cpShapeSetFriction(cs, 0.5);
	for(int k = 0; k < 40; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("capsule_on_ground_rest", "A", cap);
	cpBodyApplyImpulseAtLocalPoint(cap, cpv(6.0, 0.0), cpv(0.0, 0.0));
	for(int k = 0; k < 20; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("capsule_on_ground_kick", "A", cap);
	cpSpaceRemoveShape(space, cs); cpSpaceRemoveBody(space, cap); cpSpaceRemoveShape(space, g);
	cpShapeFree(cs); cpBodyFree(cap); cpShapeFree(g); cpSpaceFree(space);
}

/* 19. a box whose rounding radius, angle and position are edited while it rests */
static void box_edits_midrun(void)
{
	cpShape *g; cpSpace *space = ground_space(&g);
	cpBody *box = cpSpaceAddBody(space, cpBodyNew(2.0, cpMomentForBox(2.0, 10, 6)));
	cpBodySetPosition(box, cpv(0, 3.2));
	cpShape *bs = cpSpaceAddShape(space, cpBoxShapeNew(box, 10, 6, 0.0));
	cpShapeSetFriction(bs, 0.6);
	for(int k = 0; k < 30; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("box_edits_midrun_rest", "A", box);
	cpPolyShapeSetRadius(bs, 0.5);
	for(int k = 0; k < 20; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("box_edits_midrun_radius", "A", box);
	cpBodySetAngle(box, 0.3); cpBodySetPosition(box, cpv(30.0, 9.0)); cpSpaceReindexShapesForBody(space, box);
	for(int k = 0; k < 12; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("box_edits_midrun_tilted", "A", box);
	cpContactPointSet set = cpShapesCollide(bs, g);
	printf("box_edits_midrun_set A %d %a %a\n", set.count, set.normal.x, set.normal.y);
	cpSpaceRemoveShape(space, bs); cpSpaceRemoveBody(space, box); cpSpaceRemoveShape(space, g);
	cpShapeFree(bs); cpBodyFree(box); cpShapeFree(g); cpSpaceFree(space);
}

/* 20. a static platform (its own static body) is moved under a resting ball and reindexed */
static void moved_static_platform(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetGravity(space, cpv(0, -100));
	cpBody *plat = cpSpaceAddBody(space, cpBodyNewStatic());
	cpBodySetPosition(plat, cpv(0, 0));
	cpShape *ps = cpSpaceAddShape(space, cpSegmentShapeNew(plat, cpv(-30, 0), cpv(30, 0), 1.0));
	cpShapeSetFriction(ps, 1.0);
	cpShape *bs; cpBody *ball = add_ball(space, cpv(0, 6.3), 5.0, 1.0, &bs);
	for(int k = 0; k < 20; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("moved_static_platform_rest", "E", ball);
	cpBodySetPosition(plat, cpv(0, -3.0)); cpSpaceReindexShapesForBody(space, plat);
	for(int k = 0; k < 20; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("moved_static_platform_lowered", "E", ball);
	cpBB bb = cpShapeGetBB(ps);
	printf("moved_static_platform_bb E %a %a %a %a\n", bb.l, bb.b, bb.r, bb.t);
	cpSpaceRemoveShape(space, bs); cpSpaceRemoveBody(space, ball); cpSpaceRemoveShape(space, ps); cpSpaceRemoveBody(space, plat);
	cpShapeFree(bs); cpBodyFree(ball); cpShapeFree(ps); cpBodyFree(plat); cpSpaceFree(space);
}

/* ---- part 4: sleeping with joints, idle timers, queries right after an edit ---- */

/* 21. two free balls joined by a slide joint fall asleep together and wake together */
static void sleeping_pair_with_joint(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetSleepTimeThreshold(space, 0.2);
	cpShape *s0, *s1; cpBody *a = add_ball(space, cpv(0, 0), 2.0, 1.0, &s0), *b = add_ball(space, cpv(10, 0), 2.0, 3.0, &s1);
	cpConstraint *slide = cpSpaceAddConstraint(space, cpSlideJointNew(a, b, cpvzero, cpvzero, 5.0, 15.0));
	for(int k = 0; k < 30; k++) cpSpaceStep(space, 1.0/60.0);
	printf("sleeping_pair_with_joint_flags E %d %d\n", cpBodyIsSleeping(a), cpBodyIsSleeping(b));
	cpBodyApplyImpulseAtLocalPoint(a, cpv(-8.0, 1.0), cpvzero);
	printf("sleeping_pair_with_joint_woken E %d %d\n", cpBodyIsSleeping(a), cpBodyIsSleeping(b));
	for(int k = 0; k < 60; k++) cpSpaceStep(space, 1.0/60.0);
	body_line("sleeping_pair_with_joint", "E", a); body_line("sleeping_pair_with_joint", "E", b);
	printf("sleeping_pair_with_joint_impulse E %a\n", cpConstraintGetImpulse(slide));
	cpSpaceRemoveConstraint(space, slide); cpConstraintFree(slide);
	cpSpaceRemoveShape(space, s0); cpSpaceRemoveShape(space, s1); cpSpaceRemoveBody(space, a); cpSpaceRemoveBody(space, b);
	cpShapeFree(s0); cpShapeFree(s1); cpBodyFree(a); cpBodyFree(b); cpSpaceFree(space);
}

/* 22. cpBodyActivate restarts the idle timer: a resting ball that is poked every 20 steps stays awake, then sleeps */
static void idle_timer_reset(void)
{
	cpShape *g, *bs; cpSpace *space = ground_space(&g);
	cpSpaceSetSleepTimeThreshold(space, 0.5);
	cpSpaceSetIdleSpeedThreshold(space, 1.0);
	cpBody *ball = add_ball(space, cpv(0, 5.0), 5.0, 1.0, &bs);
	int flags[6];
	for(int k = 0; k < 120; k++){
		cpSpaceStep(space, 1.0/60.0);
		if(k % 20 == 19){ flags[k/20] = cpBodyIsSleeping(ball); if(k < 70) cpBodyActivate(ball); }
	}
	printf("idle_timer_reset E %d %d %d %d %d %d\n", flags[0], flags[1], flags[2], flags[3], flags[4], flags[5]);
	body_line("idle_timer_reset", "E", ball);
	cpSpaceRemoveShape(space, bs); cpSpaceRemoveBody(space, ball); cpSpaceRemoveShape(space, g);
	cpShapeFree(bs); cpBodyFree(ball); cpShapeFree(g); cpSpaceFree(space);
}

/* 23. queries see a teleport at once (no step in between), after cpSpaceReindexShapesForBody */
static void query_after_edit(void)
{
	cpShape *g, *bs; cpSpace *space = ground_space(&g);
	cpBody *ball = add_ball(space, cpv(0, 5.5), 5.0, 1.0, &bs);
	for(int k = 0; k < 5; k++) cpSpaceStep(space, 1.0/60.0);
	cpBodySetPosition(ball, cpv(40.0, 20.0)); cpSpaceReindexShapesForBody(space, ball);
	cpPointQueryInfo info;
	cpShape *hit = cpSpacePointQueryNearest(space, cpv(40.0, 30.0), 100.0, CP_SHAPE_FILTER_ALL, &info);
	printf("query_after_edit_point E %d %a %a %a\n", hit == bs, info.distance, info.point.x, info.point.y);
	cpSegmentQueryInfo seg;
	hit = cpSpaceSegmentQueryFirst(space, cpv(40.0, 50.0), cpv(40.0, -50.0), 0.0, CP_SHAPE_FILTER_ALL, &seg);
	printf("query_after_edit_segment E %d %a %a %a\n", hit == bs, seg.alpha, seg.point.y, seg.normal.y);
	cpBB bb = cpShapeGetBB(bs);
	printf("query_after_edit_bb E %a %a %a %a\n", bb.l, bb.b, bb.r, bb.t);
	cpSpaceRemoveShape(space, bs); cpSpaceRemoveBody(space, ball); cpSpaceRemoveShape(space, g);
	cpShapeFree(bs); cpBodyFree(ball); cpShapeFree(g); cpSpaceFree(space);
}

/* ---- part 5: the cpArbiter accessors and derived body getters after a step ---- */
static void print_arbiter(cpBody *body, cpArbiter *arb, void *tag)
{
	cpVect n = cpArbiterGetNormal(arb), pa = cpArbiterGetPointA(arb, 0), pb = cpArbiterGetPointB(arb, 0), j = cpArbiterTotalImpulse(arb);
	cpShape *a, *b; cpArbiterGetShapes(arb, &a, &b);
	cpBody *ba, *bb; cpArbiterGetBodies(arb, &ba, &bb);
	printf("%s E %d %d %d %d %a %a %a %a %a %a %a\n", (const char *)tag, cpArbiterGetCount(arb), cpArbiterIsFirstContact(arb), ba == body, cpShapeGetBody(a) == ba,
		n.x, n.y, cpArbiterGetDepth(arb, 0), pa.x, pa.y, pb.x, pb.y);
	printf("%s_impulse E %a %a %a %a %a %a\n", (const char *)tag, j.x, j.y, cpArbiterTotalKE(arb), cpArbiterGetRestitution(arb), cpArbiterGetFriction(arb), cpArbiterGetSurfaceVelocity(arb).x);
	cpContactPointSet set = cpArbiterGetContactPointSet(arb);
	printf("%s_set E %d %a %a %a %a %a\n", (const char *)tag, set.count, set.normal.x, set.normal.y, set.points[0].pointA.y, set.points[0].pointB.y, set.points[0].distance);
}

/* 24. a ball sliding on a conveyor ground: every accessor of its one arbiter, on the first contact and later */
static void arbiter_accessors(void)
{
	cpShape *g, *bs; cpSpace *space = ground_space(&g);
	cpShapeSetSurfaceVelocity(g, cpv(5.0, 0.0)); cpShapeSetElasticity(g, 0.5);
	cpBody *ball = add_ball(space, cpv(0, 4.95), 5.0, 2.0, &bs);
	cpBodySetVelocity(ball, cpv(3.0, -1.0));
	cpSpaceStep(space, 1.0/60.0);
	cpBodyEachArbiter(ball, print_arbiter, (void *)"arbiter_accessors_first");
	for(int k = 0; k < 9; k++) cpSpaceStep(space, 1.0/60.0);
	cpBodyEachArbiter(ball, print_arbiter, (void *)"arbiter_accessors_later");
	cpVect vw = cpBodyGetVelocityAtWorldPoint(ball, cpvadd(cpBodyGetPosition(ball), cpv(0, -5))), lw = cpBodyLocalToWorld(ball, cpv(1, 2)), wl = cpBodyWorldToLocal(ball, cpv(1, 2));
	printf("arbiter_accessors_body A %a %a %a %a %a %a %a\n", cpBodyKineticEnergy(ball), vw.x, vw.y, lw.x, lw.y, wl.x, wl.y);
	cpSpaceRemoveShape(space, bs); cpSpaceRemoveBody(space, ball); cpSpaceRemoveShape(space, g);
	cpShapeFree(bs); cpBodyFree(ball); cpShapeFree(g); cpSpaceFree(space);
}

/* 25. demo/Planet.c: a custom velocity function (gravity towards the origin, cpBodyUpdateVelocity with it) on bodies that
 * orbit a static planet and never touch anything; one body keeps the default integrator in the same space */
static void planet_gravity(cpBody *body, cpVect gravity, cpFloat damping, cpFloat dt)
{
	cpVect p = cpBodyGetPosition(body);
	cpFloat sqdist = cpvlengthsq(p);
	cpVect g = cpvmult(p, -5.0e6/(sqdist*cpfsqrt(sqdist)));
	cpBodyUpdateVelocity(body, g, damping, dt);
}

static void custom_velocity_func(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetGravity(space, cpv(0, -10));
	cpSpaceSetDamping(space, 0.99);
	cpBody *b[4];
	for(int i = 0; i < 4; i++){
		cpFloat r = 200.0 + 60.0*i;
		b[i] = cpSpaceAddBody(space, cpBodyNew(1.0 + i, 10.0 + i));
		cpBodySetPosition(b[i], cpv(r, 10.0*i));
		cpBodySetVelocity(b[i], cpv(0.0, cpfsqrt(5.0e6/r)));
		cpBodySetAngularVelocity(b[i], 0.1*i);
		if(i < 3) cpBodySetVelocityUpdateFunc(b[i], planet_gravity);
		cpShape *s = cpSpaceAddShape(space, cpCircleShapeNew(b[i], 4.0, cpvzero));
		cpShapeSetFriction(s, 0.5);
	}
	for(int k = 0; k < 40; k++){
		if(k == 10) cpBodySetForce(b[1], cpv(30.0, -20.0));     /* consumed by the custom function's cpBodyUpdateVelocity */
		if(k == 20) cpBodySetVelocityUpdateFunc(b[2], cpBodyUpdateVelocity);   /* back to the default */
		cpSpaceStep(space, 1.0/60.0);
		if(k % 13 == 0 || k == 39) for(int i = 0; i < 4; i++){ char n[48]; sprintf(n, "custom_velocity_%d_%d", k, i); body_line(n, "A", b[i]); }
	}
	cpSpaceFree(space);
}

/* 26. a custom position function (wraps x into [0, 100) after the default update) next to default bodies */
static void wrap_position(cpBody *body, cpFloat dt)
{
	cpBodyUpdatePosition(body, dt);
	cpVect p = cpBodyGetPosition(body);
	if(p.x >= 100.0) cpBodySetPosition(body, cpv(p.x - 100.0, p.y));
}

static void custom_position_func(void)
{
	cpShape *g; cpSpace *space = ground_space(&g);
	cpBody *a = add_ball(space, cpv(90.0, 30.0), 5.0, 1.0, NULL), *b = add_ball(space, cpv(20.0, 4.95), 5.0, 1.0, NULL);
	cpBodySetVelocity(a, cpv(120.0, 0.0));
	cpBodySetPositionUpdateFunc(a, wrap_position);
	for(int k = 0; k < 30; k++){
		cpSpaceStep(space, 1.0/60.0);
		if(k % 7 == 0 || k == 29){ char n[48]; sprintf(n, "custom_position_%d_a", k); body_line(n, "A", a); sprintf(n, "custom_position_%d_b", k); body_line(n, "A", b); }
	}
	cpSpaceFree(space);
}

/* 27. demo/Springies.c: a clamped spring force function, and a custom torque function on a rotary spring */
static cpFloat clamped_spring_force(cpConstraint *spring, cpFloat dist)
{
	cpFloat clamp = 20.0;
	return cpfclamp(cpDampedSpringGetRestLength(spring) - dist, -clamp, clamp)*cpDampedSpringGetStiffness(spring);
}
static cpFloat cubic_spring_torque(cpConstraint *spring, cpFloat relativeAngle)
{
	cpFloat d = relativeAngle - cpDampedRotarySpringGetRestAngle(spring);
	return d*d*d*cpDampedRotarySpringGetStiffness(spring);
}

static void custom_spring_funcs(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetGravity(space, cpv(0, -30));
	cpBody *a = cpSpaceAddBody(space, cpBodyNew(1.0, 20.0)), *b = cpSpaceAddBody(space, cpBodyNew(2.0, 30.0));
	cpBody *c = cpSpaceAddBody(space, cpBodyNew(1.5, 25.0)), *d = cpSpaceAddBody(space, cpBodyNew(1.0, 15.0));
	cpBodySetPosition(a, cpv(0, 0)); cpBodySetPosition(b, cpv(90, 10)); cpBodySetPosition(c, cpv(0, 200)); cpBodySetPosition(d, cpv(50, 200));
	cpBodySetAngle(c, 0.9); cpBodySetAngularVelocity(d, 2.0);
	cpConstraint *s1 = cpSpaceAddConstraint(space, cpDampedSpringNew(a, b, cpv(1, 0), cpv(-1, 0), 40.0, 15.0, 0.8));
	cpDampedSpringSetSpringForceFunc(s1, clamped_spring_force);
	cpConstraint *s2 = cpSpaceAddConstraint(space, cpDampedRotarySpringNew(c, d, 0.2, 120.0, 4.0));
	cpDampedRotarySpringSetSpringTorqueFunc(s2, cubic_spring_torque);
	/* default law in the same space; through c's centre of gravity it only touches c's linear velocity, the rotary spring
	 * only its angular velocity: the two commute exactly, so the solver's order cannot show */
	cpConstraint *s3 = cpSpaceAddConstraint(space, cpDampedSpringNew(cpSpaceGetStaticBody(space), c, cpv(0, 320), cpv(0, 0), 100.0, 8.0, 0.3));
	for(int k = 0; k < 45; k++){
		cpSpaceStep(space, 1.0/60.0);
		if(k % 11 == 0 || k == 44){
			char n[48];
			sprintf(n, "custom_spring_%d_a", k); body_line(n, "A", a); sprintf(n, "custom_spring_%d_b", k); body_line(n, "A", b);
			sprintf(n, "custom_spring_%d_c", k); body_line(n, "A", c); sprintf(n, "custom_spring_%d_d", k); body_line(n, "A", d);
			printf("custom_spring_%d_impulse A %a %a %a\n", k, cpConstraintGetImpulse(s1), cpConstraintGetImpulse(s2), cpConstraintGetImpulse(s3));
		}
	}
	cpSpaceFree(space);
}

/* 28. two touching shapes removed inside ONE post-step callback: separate must fire for every arbiter of both
 * (cpSpaceRemoveShape -> cpSpaceFilterArbiters, cpSpace.c:482-511), and cpBodyEachArbiter keeps working between edits */
static int g_separates = 0, g_each = 0;
static void count_separate(cpArbiter *arb, cpSpace *space, cpDataPointer data){ g_separates++; }
static void count_each(cpBody *body, cpArbiter *arb, void *data){ g_each++; }
static cpShape *g_rm[2]; static cpBody *g_rb[2];
static void remove_both(cpSpace *space, void *key, void *data)
{
	cpSpaceRemoveShape(space, g_rm[0]);
	g_each = 0; cpBodyEachArbiter(g_rb[1], count_each, NULL);      /* the other ball still sees its arbiters after the first edit */
	printf("remove_two_between E %a %a\n", (double)g_separates, (double)g_each);
	cpSpaceRemoveShape(space, g_rm[1]);
	printf("remove_two_after E %a\n", (double)g_separates);
}
static void remove_two_in_one_callback(void)
{
	cpShape *g; cpSpace *space = ground_space(&g);
	cpCollisionHandler *h = cpSpaceAddDefaultCollisionHandler(space);
	h->separateFunc = count_separate;
	g_rb[0] = add_ball(space, cpv(0.0, 4.95), 5.0, 1.0, &g_rm[0]);
	g_rb[1] = add_ball(space, cpv(9.9, 4.95), 5.0, 1.0, &g_rm[1]);
	g_separates = 0;
	for(int k = 0; k < 5; k++) cpSpaceStep(space, 1.0/60.0);
	g_each = 0; cpBodyEachArbiter(g_rb[0], count_each, NULL);
	printf("remove_two_before E %a %a\n", (double)g_separates, (double)g_each);
	cpSpaceAddPostStepCallback(space, remove_both, (void *)&g_rm, NULL);
	cpSpaceStep(space, 1.0/60.0);
	cpSpaceStep(space, 1.0/60.0);
	printf("remove_two_end E %a\n", (double)g_separates);
	cpSpaceFree(space);
}

/* 29. a constraint preSolve callback (cpSpaceStep.c:389-396) that re-parameterises its joint every step: the new rate
 * must act in the SAME step, and the callback sees this step's positions (it runs after the position update) */
static int g_presolve_calls = 0; static double g_seen_angle = 0.0;
static void motor_presolve(cpConstraint *c, cpSpace *space)
{
	g_presolve_calls++;
	g_seen_angle = cpBodyGetAngle(cpConstraintGetBodyB(c));
	cpSimpleMotorSetRate(c, (g_presolve_calls % 2) ? 3.0 : -1.5);
	if(g_presolve_calls == 6) cpConstraintSetMaxForce(c, 50.0);
}
static void constraint_presolve_callback(void)
{
	cpSpace *space = cpSpaceNew();
	cpBody *wheel = cpSpaceAddBody(space, cpBodyNew(2.0, 8.0));
	cpBodySetPosition(wheel, cpv(3, 4));
	cpConstraint *m = cpSpaceAddConstraint(space, cpSimpleMotorNew(cpSpaceGetStaticBody(space), wheel, 1.0));
	cpConstraintSetPreSolveFunc(m, motor_presolve);
	g_presolve_calls = 0;
	for(int k = 0; k < 10; k++){
		cpSpaceStep(space, 1.0/60.0);
		printf("constraint_presolve_%d A %a %a %a %a\n", k, (double)g_presolve_calls, g_seen_angle, cpBodyGetAngularVelocity(wheel), cpConstraintGetImpulse(m));
	}
	cpSpaceFree(space);
}

int main(void)
{
	empty_space();
	bodies_without_shapes();
	static_only_then_one_ball();
	coincident_circles();
	sensors_and_filters();
	kinematic_pusher();
	edits_between_steps();
	explicit_sleep();
	hub_with_many_contacts();
	resting_box();
	changing_dt();
	space_parameters_midrun();
	body_edits_midrun();
	shape_edits_midrun();
	elastic_bounce();
	joints_midrun();
	remove_and_return();
	capsule_on_ground();
	box_edits_midrun();
	moved_static_platform();
	sleeping_pair_with_joint();
	idle_timer_reset();
	query_after_edit();
	arbiter_accessors();
	custom_velocity_func();
	custom_position_func();
	custom_spring_funcs();
	remove_two_in_one_callback();
	constraint_presolve_callback();
	return 0;
}
