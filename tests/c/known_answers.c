/* known_answers.c -- the reference's own known-answer tests for the step path, restated in C
 * (the originals are Objective-C XCTest cases that cannot be built here):
 *   time_stepping        xcode/ObjectiveChipmunkTests/BodyTest.m:147-185   (exact integrator values)
 *   basic_simulation     SpaceTest.m:252-272                                (rest heights within 1.1*slop)
 *   init_step_free       SpaceTest.m:295-302
 *   initial_sleeping     SpaceTest.m:274-293
 *   handlers_move        CallbacksTest.m:61-115 (separateByRemove = false)
 *   handlers_remove      CallbacksTest.m:61-115 (separateByRemove = true)
 *   post_step_removal    CallbacksTest.m:257-294
 *   shapes_index         SpaceTest.m:229-250 (static vs dynamic membership by body type)
 * Uses only the public C API, so the same file is linked against the unmodified reference
 * (oracle/_ref) and against the B200 drop-in.  Prints one "name PASS|FAIL" line per case;
 * exit status = number of failures. */
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "chipmunk/chipmunk.h"

static int failures = 0;
#define CHECK(name, cond) do { int ok_ = (cond); printf("%s %s\n", name, ok_ ? "PASS" : "FAIL"); if(!ok_) failures++; } while(0)

static void time_stepping(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetGravity(space, cpv(0, -1));
	cpBody *dyn = cpSpaceAddBody(space, cpBodyNew(1.0, 1.0));
	cpBody *sta = cpSpaceAddBody(space, cpBodyNewStatic());
	cpBody *kin = cpSpaceAddBody(space, cpBodyNewKinematic());
	cpBodySetVelocity(kin, cpv(1, 0));
	cpSpaceStep(space, 1.0);
	int ok = cpveql(cpBodyGetPosition(dyn), cpvzero) && cpveql(cpBodyGetVelocity(dyn), cpv(0, -1))
		&& cpveql(cpBodyGetPosition(sta), cpvzero) && cpveql(cpBodyGetVelocity(sta), cpvzero)
		&& cpveql(cpBodyGetPosition(kin), cpv(1, 0)) && cpveql(cpBodyGetVelocity(kin), cpv(1, 0));
	cpSpaceStep(space, 1.0);
	ok = ok && cpveql(cpBodyGetPosition(dyn), cpv(0, -1)) && cpveql(cpBodyGetVelocity(dyn), cpv(0, -2))
		&& cpveql(cpBodyGetPosition(sta), cpvzero) && cpveql(cpBodyGetVelocity(sta), cpvzero)
		&& cpveql(cpBodyGetPosition(kin), cpv(2, 0)) && cpveql(cpBodyGetVelocity(kin), cpv(1, 0));
	CHECK("time_stepping", ok);
	cpSpaceRemoveBody(space, dyn); cpSpaceRemoveBody(space, sta); cpSpaceRemoveBody(space, kin);
	cpBodyFree(dyn); cpBodyFree(sta); cpBodyFree(kin);
	cpSpaceFree(space);
}

static cpShape *add_bound(cpSpace *space, cpVect a, cpVect b, cpFloat thickness)
{
	cpShape *s = cpSpaceAddShape(space, cpSegmentShapeNew(cpSpaceGetStaticBody(space), a, b, thickness));
	cpShapeSetElasticity(s, 1.0); cpShapeSetFriction(s, 1.0);
	return s;
}

static void basic_simulation(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetGravity(space, cpv(0, -100));
	/* ChipmunkSpace addBounds: four segments of the given radius placed just outside the box */
	cpFloat l = -50 - 1, b = 0 - 1, r = 50 + 1, t = 100 + 1;
	add_bound(space, cpv(l, b), cpv(l, t), 1); add_bound(space, cpv(l, t), cpv(r, t), 1);
	add_bound(space, cpv(r, t), cpv(r, b), 1); add_bound(space, cpv(r, b), cpv(l, b), 1);
	cpBody *ball = cpSpaceAddBody(space, cpBodyNew(1, cpMomentForCircle(1, 0, 1, cpvzero)));
	cpBodySetPosition(ball, cpv(-10, 10));
	cpSpaceAddShape(space, cpCircleShapeNew(ball, 1, cpvzero));
	cpBody *box = cpSpaceAddBody(space, cpBodyNew(1, cpMomentForBox(1, 2, 2)));
	cpBodySetPosition(box, cpv(10, 10));
	cpSpaceAddShape(space, cpBoxShapeNew(box, 2, 2, 0));
	for(int i = 0; i < 100; i++) cpSpaceStep(space, 0.01);
	cpFloat slop = cpSpaceGetCollisionSlop(space);
	CHECK("basic_simulation", cpfabs(cpBodyGetPosition(ball).y - 1) < 1.1*slop && cpfabs(cpBodyGetPosition(box).y - 1) < 1.1*slop);
	cpSpaceFree(space);
}

static void init_step_free(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceStep(space, 1);
	cpSpaceFree(space);
	CHECK("init_step_free", 1);
}

static void initial_sleeping(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetSleepTimeThreshold(space, 10.0);
	cpBody *b1 = cpSpaceAddBody(space, cpBodyNew(1.0, 1.0));
	cpSpaceAddShape(space, cpCircleShapeNew(b1, 1.0, cpvzero));
	cpBody *b2 = cpSpaceAddBody(space, cpBodyNew(1.0, 1.0));
	cpSpaceAddShape(space, cpCircleShapeNew(b2, 1.0, cpvzero));
	cpBodySleep(b1);
	cpSpaceStep(space, 1.0);
	CHECK("initial_sleeping", !cpBodyIsSleeping(b1) && !cpBodyIsSleeping(b2));
	cpSpaceFree(space);
}

static cpBool Begin(cpArbiter *arb, cpSpace *space, char *s){ strcat(s, "Begin-"); return cpTrue; }
static cpBool PreSolve(cpArbiter *arb, cpSpace *space, char *s){ strcat(s, "PreSolve-"); return cpTrue; }
static void PostSolve(cpArbiter *arb, cpSpace *space, char *s){ strcat(s, "PostSolve-"); }
static void Separate(cpArbiter *arb, cpSpace *space, char *s){ strcat(s, "Separate-"); }

static void handlers(int separateByRemove)
{
	char str[512] = "";
	cpSpace *space = cpSpaceNew();
	cpSpaceSetCollisionBias(space, 1.0);
	cpFloat radius = 5;
	cpBody *body1 = cpSpaceAddBody(space, cpBodyNew(1, 1));
	cpBodySetPosition(body1, cpv(0*radius*1.5, 0));
	cpSpaceAddShape(space, cpCircleShapeNew(body1, radius, cpvzero));
	cpBody *body2 = cpSpaceAddBody(space, cpBodyNew(1, 1));
	cpBodySetPosition(body2, cpv(1*radius*1.5, 0));
	cpShape *shape2 = cpSpaceAddShape(space, cpCircleShapeNew(body2, radius, cpvzero));
	cpCollisionHandler *handler = cpSpaceAddCollisionHandler(space, 0, 0);
	handler->beginFunc = (cpCollisionBeginFunc)Begin;
	handler->preSolveFunc = (cpCollisionPreSolveFunc)PreSolve;
	handler->postSolveFunc = (cpCollisionPostSolveFunc)PostSolve;
	handler->separateFunc = (cpCollisionSeparateFunc)Separate;
	handler->userData = str;
	cpSpaceStep(space, 0.1);
	int ok = (strcmp(str, "Begin-PreSolve-PostSolve-") == 0);
	cpSpaceStep(space, 0.1);
	ok = ok && (strcmp(str, "Begin-PreSolve-PostSolve-PreSolve-PostSolve-") == 0);
	if(separateByRemove){
		cpSpaceRemoveShape(space, shape2);
	} else {
		cpBodySetPosition(body2, cpv(100, 100));
		cpSpaceStep(space, 0.1);
	}
	ok = ok && (strcmp(str, "Begin-PreSolve-PostSolve-PreSolve-PostSolve-Separate-") == 0);
	cpSpaceStep(space, 0.1);
	CHECK(separateByRemove ? "handlers_remove" : "handlers_move", ok);
	if(!ok) printf("  callback string: %s\n", str);
	cpSpaceFree(space);
}

static void remove_bar(cpSpace *space, void *key, void *data){ cpSpaceRemoveShape(space, (cpShape *)key); }
static cpBool ball_hits_bar(cpArbiter *arb, cpSpace *space, void *data)
{
	CP_ARBITER_GET_SHAPES(arb, ballShape, barShape);
	cpSpaceAddPostStepCallback(space, remove_bar, barShape, NULL);
	return cpTrue;
}

static void post_step_removal(void)
{
	cpSpace *space = cpSpaceNew();
	cpSpaceSetGravity(space, cpv(0, -100));
	cpCollisionHandler *handler = cpSpaceAddCollisionHandler(space, 1, 2);
	handler->beginFunc = ball_hits_bar;
	cpSpaceAddShape(space, cpSegmentShapeNew(cpSpaceGetStaticBody(space), cpv(-10, 0), cpv(10, 0), 1));
	cpShape *bar = cpSpaceAddShape(space, cpSegmentShapeNew(cpSpaceGetStaticBody(space), cpv(-10, 2), cpv(10, 2), 1));
	cpShapeSetCollisionType(bar, 2);
	cpBody *ball = cpSpaceAddBody(space, cpBodyNew(1, cpMomentForCircle(1, 0, 1, cpvzero)));
	cpBodySetPosition(ball, cpv(0, 10));
	cpShape *shape = cpSpaceAddShape(space, cpCircleShapeNew(ball, 1, cpvzero));
	cpShapeSetCollisionType(shape, 1);
	for(int i = 0; i < 100; i++) cpSpaceStep(space, 0.01);
	CHECK("post_step_removal", cpfabs(cpBodyGetPosition(ball).y - 2.0) < 1.1*cpSpaceGetCollisionSlop(space));
	cpSpaceFree(space);
}

static void count_shape(cpShape *s, void *n){ (*(int *)n)++; }
static void shapes_index(void)
{
	/* a shape on a static body and one on a dynamic body are both iterated; body types are reported */
	cpSpace *space = cpSpaceNew();
	cpBody *dyn = cpSpaceAddBody(space, cpBodyNew(1, 1));
	cpSpaceAddShape(space, cpCircleShapeNew(dyn, 1, cpvzero));
	cpSpaceAddShape(space, cpCircleShapeNew(cpSpaceGetStaticBody(space), 1, cpv(10, 0)));
	int n = 0;
	cpSpaceEachShape(space, count_shape, &n);
	CHECK("shapes_index", n == 2 && cpBodyGetType(dyn) == CP_BODY_TYPE_DYNAMIC && cpBodyGetType(cpSpaceGetStaticBody(space)) == CP_BODY_TYPE_STATIC);
	cpSpaceFree(space);
}

int main(void)
{
	time_stepping();
	basic_simulation();
	init_step_free();
	initial_sleeping();
	handlers(0);
	handlers(1);
	post_step_removal();
	shapes_index();
	return failures;
}
