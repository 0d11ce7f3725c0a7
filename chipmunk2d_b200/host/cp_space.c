/* cp_space.c -- cpSpace / cpHastySpace on top of the C ABI (include/cpb200.h).
 *
 * Replaces reference src/cpSpace.c (lifecycle, registries, handlers) and the host half of
 * src/cpSpaceStep.c / src/cpHastySpace.c: cpSpaceStep() uploads whatever changed on the host,
 * enqueues ONE device step (every stage of cpSpaceStep.c:335-445 is a CUDA kernel) and marks the
 * mirrors stale.  User callbacks run on the host after the device step from downloaded state:
 *   - constraint preSolve before the step, postSolve after it (Chains' breakable joints);
 *   - collision handlers begin / preSolve / postSolve / separate in that order per arbiter,
 *     observational (their return values cannot retro-actively filter the finished solve);
 *   - post-step callbacks exactly like cpSpaceUnlock(space, true) (cpSpaceStep.c:84-103).
 */
#include <string.h>
#include <stdio.h>

#include "cp_host.h"
#include <omp.h>
#include <stdlib.h>

static int g_default_device = -1;

void cpSpaceSetDefaultDeviceB200(int device){ g_default_device = device; }

static int
default_device(void)
{
	if(g_default_device >= 0) return g_default_device;
	const char *env = getenv("CPB200_DEVICE");
	return (env ? atoi(env) : 0);
}

static void *
grow(void *arr, int *cap, int need, size_t elt)
{
	if(need <= *cap) return arr;
	int ncap = (*cap ? *cap : 16);
	while(ncap < need) ncap *= 2;
	arr = cprealloc(arr, elt*(size_t)ncap);
	*cap = ncap;
	return arr;
}

/* ---- lifecycle (cpSpace.c:105-229) ---- */
static cpBool handler_true(cpArbiter *arb, cpSpace *space, cpDataPointer data){ (void)arb; (void)space; (void)data; return cpTrue; }
static void handler_nothing(cpArbiter *arb, cpSpace *space, cpDataPointer data){ (void)arb; (void)space; (void)data; }
static const cpCollisionHandler handler_do_nothing = {CP_WILDCARD_COLLISION_TYPE, CP_WILDCARD_COLLISION_TYPE, handler_true, handler_true, handler_nothing, handler_nothing, NULL};

cpSpace *cpSpaceAlloc(void){ return (cpSpace *)cpcalloc(1, sizeof(cpSpace)); }

cpSpace *
cpSpaceInit(cpSpace *space)
{
	memset(space, 0, sizeof(*space));
	space->iterations = 10;
	space->gravity = cpvzero;
	space->damping = 1.0;
	/* the reference spells these defaults with float literals (cpSpace.c:136-137): the values are the
	 * float-rounded ones, 0.100000001490116 and 0.899999976158142^60 */
	space->collisionSlop = (cpFloat)0.1f;
	space->collisionBias = cpfpow((cpFloat)(1.0f - 0.1f), 60.0);
	space->collisionPersistence = 3;
	space->idleSpeedThreshold = 0.0;
	space->sleepTimeThreshold = INFINITY;
	space->device = default_device();
	memcpy(&space->defaultHandler, &handler_do_nothing, sizeof(cpCollisionHandler));

	cpBody *staticBody = cpBodyInit(&space->_staticBody, 0.0, 0.0);
	cpBodySetType(staticBody, CP_BODY_TYPE_STATIC);
	space->staticBody = staticBody;
	space->bodies = (cpBody **)grow(NULL, &space->capBodies, 16, sizeof(cpBody *));
	space->bodies[0] = staticBody;
	space->nBodies = 1;
	staticBody->space = space;
	staticBody->index = 0;
	space->topologyDirty = cpTrue;
	space->paramsDirty = cpTrue;
	space->noAppend = (getenv("CPB200_NO_APPEND") != NULL);
	return space;
}

cpSpace *cpSpaceNew(void){ return cpSpaceInit(cpSpaceAlloc()); }

void
cpSpaceDestroy(cpSpace *space)
{
	/* like the reference, the space never owns bodies/shapes/constraints (cpSpace.c:188-229) */
	if(space->world) cpSpaceFetchBodiesB200(space);   /* the bodies outlive the space: leave them their last state */
	for(int i = 0; i < space->nBodies; i++){ space->bodies[i]->space = NULL; space->bodies[i]->index = -1; }
	for(int i = 0; i < space->nShapes; i++){ space->shapes[i]->space = NULL; space->shapes[i]->index = -1; }
	for(int i = 0; i < space->nConstraints; i++){ space->constraints[i]->space = NULL; space->constraints[i]->index = -1; }
	if(space->xferStates) cpb200_host_free(space->xferStates);
	if(space->xferForces) cpb200_host_free(space->xferForces);
	space->xferStates = space->xferForces = NULL; space->xferStatesBytes = space->xferForcesBytes = 0;
	if(space->world) cpb200_world_destroy(space->world);
	space->world = NULL;
	cpfree(space->bodies); cpfree(space->shapes); cpfree(space->constraints);
	cpfree(space->handlers); cpfree(space->postStep); cpfree(space->arbs); cpfree(space->arbData); space->arbData = NULL;
	space->bodies = NULL; space->shapes = NULL; space->constraints = NULL;
	space->handlers = NULL; space->postStep = NULL; space->arbs = NULL;
}

void cpSpaceFree(cpSpace *space){ if(space){ cpSpaceDestroy(space); cpfree(space); } }

/* ---- properties (cpSpace.c:231-381) ---- */
#define SPACE_PROP(ctype, Name, field) \
	ctype cpSpaceGet##Name(const cpSpace *space){ return space->field; } \
	void cpSpaceSet##Name(cpSpace *space, ctype value){ space->field = value; space->paramsDirty = cpTrue; }
SPACE_PROP(cpVect, Gravity, gravity)
SPACE_PROP(cpFloat, IdleSpeedThreshold, idleSpeedThreshold)
SPACE_PROP(cpFloat, SleepTimeThreshold, sleepTimeThreshold)
SPACE_PROP(cpTimestamp, CollisionPersistence, collisionPersistence)
SPACE_PROP(cpDataPointer, UserData, userData)
int cpSpaceGetIterations(const cpSpace *space){ return space->iterations; }
void cpSpaceSetIterations(cpSpace *space, int iterations){ cpAssertHard(iterations > 0, "Iterations must be positive and non-zero."); space->iterations = iterations; space->paramsDirty = cpTrue; }
cpFloat cpSpaceGetDamping(const cpSpace *space){ return space->damping; }
void cpSpaceSetDamping(cpSpace *space, cpFloat damping){ cpAssertHard(damping >= 0.0, "Damping must be positive."); space->damping = damping; space->paramsDirty = cpTrue; }
cpFloat cpSpaceGetCollisionSlop(const cpSpace *space){ return space->collisionSlop; }
void cpSpaceSetCollisionSlop(cpSpace *space, cpFloat collisionSlop){ space->collisionSlop = collisionSlop; space->paramsDirty = cpTrue; }
cpFloat cpSpaceGetCollisionBias(const cpSpace *space){ return space->collisionBias; }
void cpSpaceSetCollisionBias(cpSpace *space, cpFloat collisionBias){ space->collisionBias = collisionBias; space->paramsDirty = cpTrue; }
cpBody *cpSpaceGetStaticBody(const cpSpace *space){ return space->staticBody; }
cpFloat cpSpaceGetCurrentTimeStep(const cpSpace *space){ return space->curr_dt; }
cpBool cpSpaceIsLocked(cpSpace *space){ return (space->locked > 0); }
void cpSpaceMarkTopologyDirty(cpSpace *space){ space->topologyDirty = cpTrue; }
/* a re-parameterised object forces the full re-upload only if the device already holds it: one that was added since the
 * last sync is still waiting for its (first) upload as part of an appended range */
/* (mass, moment, type, centre of gravity and integrator changes travel with the ordinary body upload; parameter changes of
 * shapes / constraints re-upload that object class only -- a per-frame cpSimpleMotorSetRate must not re-upload the world) */
void cpSpaceMarkBodyDirtyB200(cpBody *body){ cpSpace *sp = body->space; if(sp && !(sp->appendDirty && !sp->topologyDirty && body->index >= sp->nBodiesOnDevice)) sp->bodiesDirty = cpTrue; }
void cpSpaceMarkShapeDirtyB200(cpShape *shape){ cpSpace *sp = shape->space; if(sp && !(sp->appendDirty && !sp->topologyDirty && shape->index >= sp->nShapesOnDevice)) sp->shapesDirty = cpTrue; }
void cpSpaceMarkConstraintDirtyB200(cpConstraint *c){ cpSpace *sp = c->space; if(sp && !(sp->appendDirty && !sp->topologyDirty && c->index >= sp->nConstraintsOnDevice)) sp->jointsDirty = cpTrue; }
/* an addition keeps the device's objects where they are if nothing else is pending */
static cpBool
can_append(const cpSpace *space)
{
	return space->world != NULL && !space->topologyDirty && !space->shapesDirty && !space->jointsDirty && !space->noAppend && space->locked == 0;
}
void cpSpaceSetSolverModeB200(cpSpace *space, int mode){ space->solverMode = mode; space->paramsDirty = cpTrue; }

/* ---- collision handlers (cpSpace.c:383-413) ---- */
/* DefaultBegin .. DefaultSeparate (cpSpace.c:63-90): dispatch to the wildcard handlers of both shapes */
static cpBool default_begin(cpArbiter *arb, cpSpace *space, cpDataPointer data){
	cpBool retA = cpArbiterCallWildcardBeginA(arb, space);
	cpBool retB = cpArbiterCallWildcardBeginB(arb, space);
	return retA && retB;
}
static cpBool default_presolve(cpArbiter *arb, cpSpace *space, cpDataPointer data){
	cpBool retA = cpArbiterCallWildcardPreSolveA(arb, space);
	cpBool retB = cpArbiterCallWildcardPreSolveB(arb, space);
	return retA && retB;
}
static void default_postsolve(cpArbiter *arb, cpSpace *space, cpDataPointer data){
	cpArbiterCallWildcardPostSolveA(arb, space);
	cpArbiterCallWildcardPostSolveB(arb, space);
}
static void default_separate(cpArbiter *arb, cpSpace *space, cpDataPointer data){
	cpArbiterCallWildcardSeparateA(arb, space);
	cpArbiterCallWildcardSeparateB(arb, space);
}

static cpCollisionHandler *
add_handler(cpSpace *space, cpCollisionType a, cpCollisionType b)
{
	for(int i = 0; i < space->nHandlers; i++){
		cpCollisionHandler *h = &space->handlers[i];
		if((h->typeA == a && h->typeB == b) || (h->typeA == b && h->typeB == a)) return h;
	}
	cpAssertHard(space->nHandlers < 256, "Too many collision handlers (the handler table is fixed so that returned pointers stay valid).");
	if(!space->handlers){ space->handlers = (cpCollisionHandler *)cpcalloc(256, sizeof(cpCollisionHandler)); space->capHandlers = 256; }
	/* a pair handler's unset callbacks fall through to the wildcard handlers of both types
	 * (cpSpace.c:398-403); a wildcard handler's own defaults accept everything (cpSpace.c:405-413) */
	cpCollisionHandler pair = {a, b, default_begin, default_presolve, default_postsolve, default_separate, NULL};
	cpCollisionHandler wild = {a, b, handler_true, handler_true, handler_nothing, handler_nothing, NULL};
	memcpy(&space->handlers[space->nHandlers], (b == CP_WILDCARD_COLLISION_TYPE ? &wild : &pair), sizeof(pair));
	return &space->handlers[space->nHandlers++];
}

/* cpSpaceUseWildcardDefaultHandler (cpSpace.c:382-390) */
static void
use_wildcard_default_handler(cpSpace *space)
{
	if(!space->usesWildcards){
		space->usesWildcards = cpTrue;
		cpCollisionHandler def = {CP_WILDCARD_COLLISION_TYPE, CP_WILDCARD_COLLISION_TYPE, default_begin, default_presolve, default_postsolve, default_separate, NULL};
		memcpy(&space->defaultHandler, &def, sizeof(def));
	}
	space->hasDefaultHandler = cpTrue;   /* every pair may now reach a user callback */
}

cpCollisionHandler *
cpSpaceAddDefaultCollisionHandler(cpSpace *space)
{
	use_wildcard_default_handler(space);
	return &space->defaultHandler;
}

cpCollisionHandler *cpSpaceAddCollisionHandler(cpSpace *space, cpCollisionType a, cpCollisionType b){ return add_handler(space, a, b); }

cpCollisionHandler *
cpSpaceAddWildcardHandler(cpSpace *space, cpCollisionType type)
{
	use_wildcard_default_handler(space);
	return add_handler(space, type, CP_WILDCARD_COLLISION_TYPE);
}

cpCollisionHandler *
cpSpaceLookupWildcardB200(cpSpace *space, cpCollisionType type)
{
	for(int i = 0; i < space->nHandlers; i++){
		cpCollisionHandler *h = &space->handlers[i];
		if(h->typeA == type && h->typeB == CP_WILDCARD_COLLISION_TYPE) return h;
	}
	return NULL;
}

static cpCollisionHandler *
lookup_handler(cpSpace *space, cpCollisionType a, cpCollisionType b)
{
	for(int i = 0; i < space->nHandlers; i++){
		cpCollisionHandler *h = &space->handlers[i];
		if(h->typeB == CP_WILDCARD_COLLISION_TYPE) continue;
		if((h->typeA == a && h->typeB == b) || (h->typeA == b && h->typeB == a)) return h;
	}
	return &space->defaultHandler;
}

/* ---- registries (cpSpace.c:415-571) ---- */
cpBody *
cpSpaceAddBody(cpSpace *space, cpBody *body)
{
	cpAssertHard(body->space != space, "You have already added this body to this space. You must not add it a second time.");
	cpAssertHard(!body->space, "You have already added this body to another space. You cannot add it to a second.");
	cpAssertSpaceUnlocked(space);
	const cpBool append = can_append(space);
	if(!append) cpSpaceFetchBodiesB200(space);          /* mirrors current before everything is re-uploaded from them */
	body->fetchStamp = space->fetchStamp;   /* a newcomer has no record in the last download */
	space->bodies = (cpBody **)grow(space->bodies, &space->capBodies, space->nBodies + 1, sizeof(cpBody *));
	body->index = space->nBodies;
	space->bodies[space->nBodies++] = body;
	body->space = space;
	if(append) space->appendDirty = cpTrue; else space->topologyDirty = cpTrue;
	if(body->position_func != cpBodyUpdatePosition || body->velocity_func != cpBodyUpdateVelocity) space->anyCustom = cpTrue;
	return body;
}

cpShape *
cpSpaceAddShape(cpSpace *space, cpShape *shape)
{
	cpAssertHard(shape->space != space, "You have already added this shape to this space. You must not add it a second time.");
	cpAssertHard(!shape->space, "You have already added this shape to another space. You cannot add it to a second.");
	cpAssertHard(shape->body, "The shape's body is not defined.");
	cpAssertHard(shape->body->space == space, "The shape's body must be added to the space before the shape.");
	cpAssertSpaceUnlocked(space);
	cpBody *body = shape->body;
	/* appended in place when nothing else is pending; a shape whose mass changes a body that is already on the device
	 * (cpBodyAddShape -> cpBodyAccumulateMassFromShapes) marks the topology dirty by itself */
	const cpBool append = can_append(space);
	if(append) space->appendDirty = cpTrue;
	cpBodyActivate(body);
	cpBodyAddShape(body, shape);
	shape->hashid = space->shapeIDCounter++;
	cpBodySyncForRead(body);
	cpShapeUpdate(shape, body->transform);
	space->shapes = (cpShape **)grow(space->shapes, &space->capShapes, space->nShapes + 1, sizeof(cpShape *));
	shape->index = space->nShapes;
	space->shapes[space->nShapes++] = shape;
	shape->space = space;
	if(!append) space->topologyDirty = cpTrue;
	return shape;
}

cpConstraint *
cpSpaceAddConstraint(cpSpace *space, cpConstraint *constraint)
{
	cpAssertHard(constraint->space != space, "You have already added this constraint to this space. You must not add it a second time.");
	cpAssertHard(!constraint->space, "You have already added this constraint to another space. You cannot add it to a second.");
	cpAssertSpaceUnlocked(space);
	cpBody *a = constraint->a, *b = constraint->b;
	cpAssertHard(a != NULL && b != NULL, "Constraint is attached to a NULL body.");
	cpAssertHard(a->space == space && b->space == space, "The constraint's bodies must be added to the space before the constraint.");
	cpBodyActivate(a);
	cpBodyActivate(b);
	space->constraints = (cpConstraint **)grow(space->constraints, &space->capConstraints, space->nConstraints + 1, sizeof(cpConstraint *));
	constraint->index = space->nConstraints;
	space->constraints[space->nConstraints++] = constraint;
	cpBodyAddConstraint(a, constraint);
	cpBodyAddConstraint(b, constraint);
	constraint->space = space;
	if(can_append(space)) space->appendDirty = cpTrue; else space->topologyDirty = cpTrue;
	if(constraint->forceFunc) space->anyCustom = cpTrue;
	return constraint;
}

/* f4: a removal is done in place on the device (the engine moves its last object into the hole, exactly like the
 * registries below) when nothing else is pending; additions that have not travelled yet go first, so that "last" means
 * the same object on both sides.  Returns cpFalse if the removal has to take the full re-upload instead. */
static cpBool append_to_device(cpSpace *space);
static cpBool
flush_appends_for_removal(cpSpace *space)
{
	if(!can_append(space)) return cpFalse;
	if(space->appendDirty){
		if(!append_to_device(space)){ space->topologyDirty = cpTrue; return cpFalse; }
		space->appendDirty = cpFalse;
	}
	return cpTrue;
}

/* before a structural edit the mirrors must hold the device's latest state, because the edit forces a
 * full re-upload from them */
static void
sync_before_edit(cpSpace *space)
{
	cpSpaceFetchBodiesB200(space);
	cpSpaceFetchBiasB200(space);
	if(space->jointStale) cpSpaceFetchJointsB200(space);
}

/* The arbiter mirrors of the last step stay in use across host-side edits (cpBodyEachArbiter, the getters, separate
 * callbacks of later removals): before a removal compacts the host arrays they are fetched while device records can still
 * be mapped to host objects, and the mirrors that touch the removed shape are dropped afterwards. */
static void
drop_arbiters_of_shape(cpSpace *space, cpShape *shape)
{
	int m = 0;
	for(int i = 0; i < space->nBodies; i++) space->bodies[i]->firstArb = -1;
	for(int k = 0; k < space->nArbs; k++){
		cpArbiter *arb = &space->arbs[k];
		if(arb->a == shape || arb->b == shape) continue;
		if(m != k) space->arbs[m] = *arb;
		arb = &space->arbs[m];
		arb->next_a = arb->next_b = -1;
		/* (the same rule cpSpaceFetchArbitersB200 threads by) */
		if(arb->active || (arb->count > 0 && arb->state != CP_ARBITER_STATE_CACHED && (arb->body_a->sleepRoot || arb->body_b->sleepRoot))){
			arb->next_a = arb->body_a->firstArb; arb->body_a->firstArb = m;
			arb->next_b = arb->body_b->firstArb; arb->body_b->firstArb = m;
		}
		m++;
	}
	space->nArbs = m;
}

void
cpSpaceRemoveShape(cpSpace *space, cpShape *shape)
{
	cpAssertHard(cpSpaceContainsShape(space, shape), "Cannot remove a shape that was not added to the space. (Removed twice maybe?)");
	cpAssertSpaceUnlocked(space);
	if(space->world && shape->index < space->nShapesOnDevice){
		/* arbiters of the removed shape separate now (cpSpaceFilterArbiters, cpSpace.c:482-511) -- on EVERY removal, also the
		 * second one inside the same post-step callback */
		if(space->arbStale) cpSpaceFetchArbitersB200(space);
		if(space->nHandlers > 0 || space->hasDefaultHandler){
			space->locked++;
			for(int k = 0; k < space->nArbs; k++){
				cpArbiter *arb = &space->arbs[k];
				if((arb->a == shape || arb->b == shape) && arb->state != CP_ARBITER_STATE_CACHED && arb->stamp == space->stamp){
					arb->state = CP_ARBITER_STATE_INVALIDATED;
					arb->handler->separateFunc(arb, space, arb->handler->userData);
				}
			}
			space->locked--;
		}
		drop_arbiters_of_shape(space, shape);
	}
	cpBool inPlace = cpFalse;
	if(flush_appends_for_removal(space) && shape->index < space->nShapesOnDevice){
		int rc = cpb200_world_remove_shape(space->world, shape->index);
		if(rc < 0) cpEngineError("shape removal");
		inPlace = (rc == 0);
	}
	if(!inPlace) sync_before_edit(space);
	cpBody *body = shape->body;
	cpBodyActivate(body);
	cpBodyRemoveShape(body, shape);
	int i = shape->index, last = --space->nShapes;
	if(i != last){ space->shapes[i] = space->shapes[last]; space->shapes[i]->index = i; }
	shape->space = NULL;
	shape->index = -1;
	if(inPlace){ space->nShapesOnDevice--; space->bbStale = cpTrue; }
	else { space->topologyDirty = cpTrue; space->shapeIndexDirty = cpTrue; }
}

void
cpSpaceRemoveBody(cpSpace *space, cpBody *body)
{
	cpAssertHard(body != cpSpaceGetStaticBody(space), "Cannot remove the designated static body for the space.");
	cpAssertHard(cpSpaceContainsBody(space, body), "Cannot remove a body that was not added to the space. (Removed twice maybe?)");
	cpAssertSpaceUnlocked(space);
	cpBool inPlace = cpFalse;
	if(body->shapeList == NULL && body->constraintList == NULL && flush_appends_for_removal(space) && body->index < space->nBodiesOnDevice){
		/* every mirror takes its record of the last download first: the records are addressed by the OLD slots */
		cpSpaceFetchBodiesB200(space);
		cpSpaceFetchBiasB200(space);
		int rc = cpb200_world_remove_body(space->world, body->index);
		if(rc < 0) cpEngineError("body removal");
		inPlace = (rc == 0);
	}
	if(!inPlace) sync_before_edit(space);
	cpBodyActivate(body);
	int i = body->index, last = --space->nBodies;
	if(i != last){ space->bodies[i] = space->bodies[last]; space->bodies[i]->index = i; }
	body->space = NULL;
	body->index = -1;
	body->sleepRoot = NULL;
	if(inPlace) space->nBodiesOnDevice--; else space->topologyDirty = cpTrue;
}

void
cpSpaceRemoveConstraint(cpSpace *space, cpConstraint *constraint)
{
	cpAssertHard(cpSpaceContainsConstraint(space, constraint), "Cannot remove a constraint that was not added to the space. (Removed twice maybe?)");
	cpAssertSpaceUnlocked(space);
	cpBool inPlace = cpFalse;
	if(flush_appends_for_removal(space) && constraint->index < space->nConstraintsOnDevice){
		if(space->jointStale) cpSpaceFetchJointsB200(space);     /* the removed joint keeps its last impulse for cpConstraintGetImpulse */
		int rc = cpb200_world_remove_joint(space->world, constraint->index);
		if(rc < 0) cpEngineError("constraint removal");
		inPlace = (rc == 0);
	}
	if(!inPlace) sync_before_edit(space);
	cpBodyActivate(constraint->a);
	cpBodyActivate(constraint->b);
	cpBodyRemoveConstraint(constraint->a, constraint);
	cpBodyRemoveConstraint(constraint->b, constraint);
	int i = constraint->index, last = --space->nConstraints;
	if(i != last){ space->constraints[i] = space->constraints[last]; space->constraints[i]->index = i; }
	constraint->space = NULL;
	constraint->index = -1;
	if(inPlace) space->nConstraintsOnDevice--;
	else { space->jointIndexDirty = cpTrue;   /* host slots no longer match device indices until the next upload */ space->topologyDirty = cpTrue; }
}

cpBool cpSpaceContainsShape(cpSpace *space, cpShape *shape){ return (shape->space == space); }
cpBool cpSpaceContainsBody(cpSpace *space, cpBody *body){ return (body->space == space); }
cpBool cpSpaceContainsConstraint(cpSpace *space, cpConstraint *constraint){ return (constraint->space == space); }

/* ---- iteration (cpSpace.c:573-651) ---- */
void
cpSpaceEachBody(cpSpace *space, cpSpaceBodyIteratorFunc func, void *data)
{
	space->locked++;
	/* the callback usually touches the first lines of each body, which the parallel mirror update left in other
	 * cores' caches: ask for them a few bodies ahead */
	cpBody **bodies = space->bodies;
	const int n = space->nBodies;
	for(int i = 1; i < n; i++){
		if(i + 8 < n){ const char *nx = (const char *)bodies[i + 8]; __builtin_prefetch(nx, 1); __builtin_prefetch(nx + 64, 0); }
		func(bodies[i], data);
	}
	space->locked--;
}

void
cpSpaceEachShape(cpSpace *space, cpSpaceShapeIteratorFunc func, void *data)
{
	space->locked++;
	for(int i = 0; i < space->nShapes; i++) func(space->shapes[i], data);
	space->locked--;
}

void
cpSpaceEachConstraint(cpSpace *space, cpSpaceConstraintIteratorFunc func, void *data)
{
	space->locked++;
	for(int i = 0; i < space->nConstraints; i++) func(space->constraints[i], data);
	space->locked--;
}

/* the device rebuilds its broadphase from scratch every step; "reindexing" is a re-upload */
void cpSpaceReindexStatic(cpSpace *space){ space->topologyDirty = cpTrue; }
void cpSpaceReindexShape(cpSpace *space, cpShape *shape){ (void)shape; space->topologyDirty = cpTrue; }
void cpSpaceReindexShapesForBody(cpSpace *space, cpBody *body){ (void)body; space->topologyDirty = cpTrue; }
void cpSpaceUseSpatialHash(cpSpace *space, cpFloat dim, int count){ (void)space; (void)dim; (void)count; }

cpBool
cpSpaceAddPostStepCallback(cpSpace *space, cpPostStepFunc func, void *key, void *data)
{
	for(int i = 0; i < space->nPostStep; i++){ if(space->postStep[i].key == key) return cpFalse; }
	space->postStep = (cpPostStepCallback *)grow(space->postStep, &space->capPostStep, space->nPostStep + 1, sizeof(cpPostStepCallback));
	cpPostStepCallback cb = {func, key, data};
	space->postStep[space->nPostStep++] = cb;
	return cpTrue;
}

static void
run_post_step_callbacks(cpSpace *space)
{
	if(space->locked != 0 || space->skipPostStep) return;
	space->skipPostStep = cpTrue;
	for(int i = 0; i < space->nPostStep; i++){
		cpPostStepCallback cb = space->postStep[i];
		space->postStep[i].func = NULL;
		if(cb.func) cb.func(space, cb.key, cb.data);
	}
	space->nPostStep = 0;
	space->skipPostStep = cpFalse;
}

/* ---- host <-> device ---- */
static void
ensure_world(cpSpace *space)
{
	if(space->world) return;
	space->world = cpb200_world_create(space->device, 1);
	if(!space->world) cpEngineError("cpb200_world_create");
	/* The device buffers default to 16 pair candidates and 8 arbiter records per shape; a scene that needs more fails
	 * loudly at its next read-back (cpb200_world_get_bodies / get_arbiters report the overflow) and can raise them here. */
	const char *rp = getenv("CPB200_RESERVE_PAIRS"), *ra = getenv("CPB200_RESERVE_ARBITERS");
	if(rp || ra) cpb200_world_reserve(space->world, rp ? atoi(rp) : 0, ra ? atoi(ra) : 0);
}

static void
fill_body_desc(cpb200_body_desc *d, const cpBody *b)
{
	memset(d, 0, sizeof(*d));
	d->p[0] = b->p.x; d->p[1] = b->p.y; d->v[0] = b->v.x; d->v[1] = b->v.y; d->f[0] = b->f.x; d->f[1] = b->f.y;
	d->a = b->a; d->w = b->w; d->t = b->t;
	d->rot[0] = b->transform.a; d->rot[1] = b->transform.b;
	d->m = b->m; d->i = b->i;
	d->cog[0] = b->cog.x; d->cog[1] = b->cog.y;
	d->v_bias[0] = b->v_bias.x; d->v_bias[1] = b->v_bias.y; d->w_bias = b->w_bias;
	d->idle_time = b->idleTime;
	d->type = (b->idleTime == INFINITY ? CPB200_BODY_STATIC : (b->m == INFINITY ? CPB200_BODY_KINEMATIC : CPB200_BODY_DYNAMIC));
	d->space = 0;
	d->sleeping = (b->sleepRoot != NULL);
	d->sleep_group = (b->sleepRoot ? b->sleepRoot->index : -1);
	d->custom = (b->position_func != cpBodyUpdatePosition ? CPB200_BODY_HOST_POSITION : 0) | (b->velocity_func != cpBodyUpdateVelocity ? CPB200_BODY_HOST_VELOCITY : 0);
}

static void
upload_bodies(cpSpace *space, cpBool full)
{
	int n = space->nBodies;
	cpb200_body_desc *descs = (cpb200_body_desc *)cpcalloc((size_t)n, sizeof(cpb200_body_desc));
	for(int i = 0; i < n; i++) fill_body_desc(&descs[i], space->bodies[i]);
	int rc = (full ? cpb200_world_set_bodies(space->world, n, descs) : cpb200_world_update_bodies(space->world, 0, n, descs));
	cpfree(descs);
	if(rc) cpEngineError("body upload");
	space->nBodiesOnDevice = n;
	space->biasStale = cpFalse;
}

/* grow-only page-locked exchange buffer */
static void *
xfer_buffer(void **buf, size_t *have, size_t need)
{
	if(*have < need){
		if(*buf) cpb200_host_free(*buf);
		size_t bytes = need + need/4 + 4096;
		*buf = cpb200_host_alloc(bytes);
		cpAssertHard(*buf != NULL, "Could not allocate the page-locked exchange buffer.");
		*have = bytes;
	}
	return *buf;
}

/* Host loops over all bodies (mirror <-> exchange buffer) are memory bound; large spaces split them over
 * the host cores.  cpHastySpaceSetThreads() caps the team (the reference's knob for its threaded solver). */
#define CP_PARALLEL_MIN_BODIES 20000
static int
host_threads(const cpSpace *space, int n)
{
	if(n < CP_PARALLEL_MIN_BODIES) return 1;
	int t = omp_get_num_procs();
	if(t > 16) t = 16;
	/* several processes on one host (one per GPU) share the cores: CPB200_HOST_THREADS caps the team */
	const char *cap = getenv("CPB200_HOST_THREADS");
	if(cap && atoi(cap) > 0 && atoi(cap) < t) t = atoi(cap);
	if(space->hasty && space->hastyThreads > 0 && (int)space->hastyThreads < t) t = (int)space->hastyThreads;
	return (t < 1 ? 1 : t);
}

static void
upload_forces(cpSpace *space)
{
	int n = space->nBodies;
	double *f = (double *)xfer_buffer(&space->xferForces, &space->xferForcesBytes, (size_t)n*3*sizeof(double));
	cpBody **bodies = space->bodies;
	#pragma omp parallel for schedule(static) num_threads(host_threads(space, n))
	for(int i = 0; i < n; i++){ const cpBody *b = bodies[i]; f[3*i] = b->f.x; f[3*i + 1] = b->f.y; f[3*i + 2] = b->t; }
	if(cpb200_world_set_body_forces(space->world, 0, n, f)) cpEngineError("force upload");
}

/* one shape as the C ABI takes it; polygon vertices are appended at *voff of `verts` */
static void
fill_shape_desc(cpSpace *space, cpShape *s, cpb200_shape_desc *d, double *verts, int *voff)
{
	d->type = s->klass;
	cpAssertHard(s->body->space == space, "A shape's body was removed from the space while the shape is still in it.");
	d->body = s->body->index;
	d->hashid = (uint32_t)s->hashid;
	d->sensor = s->sensor;
	d->categories = s->filter.categories; d->mask = s->filter.mask; d->group = (uint64_t)s->filter.group;
	d->collision_type = (uint64_t)s->type;
	d->e = s->e; d->u = s->u;
	d->surface_v[0] = s->surfaceV.x; d->surface_v[1] = s->surfaceV.y;
	switch(s->klass){
	case CP_CIRCLE_SHAPE: { cpCircleShape *c = (cpCircleShape *)s; d->r = c->r; d->a[0] = c->c.x; d->a[1] = c->c.y; break; }
	case CP_SEGMENT_SHAPE: {
		cpSegmentShape *g = (cpSegmentShape *)s;
		d->r = g->r; d->a[0] = g->a.x; d->a[1] = g->a.y; d->b[0] = g->b.x; d->b[1] = g->b.y;
		d->a_tangent[0] = g->a_tangent.x; d->a_tangent[1] = g->a_tangent.y; d->b_tangent[0] = g->b_tangent.x; d->b_tangent[1] = g->b_tangent.y;
		break;
	}
	default: {
		cpPolyShape *p = (cpPolyShape *)s;
		d->r = p->r; d->n_verts = p->count; d->vert_offset = *voff;
		for(int k = 0; k < p->count; k++){ verts[2*(*voff + k)] = p->verts[k].x; verts[2*(*voff + k) + 1] = p->verts[k].y; }
		*voff += p->count;
		break;
	}
	}
}

/* shapes [first, nShapes): the whole set (first = 0, cpb200_world_set_shapes) or the appended tail */
static int
upload_shape_range(cpSpace *space, int first)
{
	int n = space->nShapes - first, nv = 0;
	for(int i = first; i < space->nShapes; i++){ if(space->shapes[i]->klass == CP_POLY_SHAPE) nv += ((cpPolyShape *)space->shapes[i])->count; }
	cpb200_shape_desc *descs = (cpb200_shape_desc *)cpcalloc((n > 0 ? (size_t)n : 1), sizeof(cpb200_shape_desc));
	double *verts = (double *)cpcalloc((size_t)(nv ? nv : 1), 2*sizeof(double));
	int voff = 0;
	for(int i = 0; i < n; i++) fill_shape_desc(space, space->shapes[first + i], &descs[i], verts, &voff);
	int rc = (first == 0 ? cpb200_world_set_shapes(space->world, n, descs, nv, verts) : cpb200_world_append_shapes(space->world, n, descs, nv, verts));
	cpfree(descs); cpfree(verts);
	if(rc < 0) cpEngineError("shape upload");
	if(rc == 0){ space->nShapesOnDevice = space->nShapes; space->nVertsOnDevice = (first == 0 ? nv : space->nVertsOnDevice + nv); }
	return rc;
}

static void upload_shapes(cpSpace *space){ upload_shape_range(space, 0); }

/* constraints [first, nConstraints): the whole set (first = 0) or the appended tail */
static int
upload_joint_range(cpSpace *space, int first)
{
	int n = space->nConstraints - first;
	cpb200_joint_desc *descs = (cpb200_joint_desc *)cpcalloc((size_t)(n > 0 ? n : 1), sizeof(cpb200_joint_desc));
	for(int i = 0; i < n; i++){
		cpConstraint *c = space->constraints[first + i];
		cpb200_joint_desc *d = &descs[i];
		cpAssertHard(c->a->space == space && c->b->space == space, "A constraint's body was removed from the space while the constraint is still in it.");
		d->type = c->klass; d->a = c->a->index; d->b = c->b->index;
		d->collide_bodies = c->collideBodies;
		d->max_force = c->maxForce; d->error_bias = c->errorBias; d->max_bias = c->maxBias;
		d->anchor_a[0] = c->anchorA.x; d->anchor_a[1] = c->anchorA.y; d->anchor_b[0] = c->anchorB.x; d->anchor_b[1] = c->anchorB.y;
		for(int k = 0; k < 4; k++) d->prm[k] = c->prm[k];
		d->acc[0] = c->acc.x; d->acc[1] = c->acc.y;
	}
	int rc = (first == 0 ? cpb200_world_set_joints(space->world, n, descs) : cpb200_world_append_joints(space->world, n, descs));
	cpfree(descs);
	if(rc < 0) cpEngineError("joint upload");
	if(rc == 0){ space->nConstraintsOnDevice = space->nConstraints; if(first == 0) space->jointIndexDirty = cpFalse; }
	return rc;
}

static void upload_joints(cpSpace *space){ upload_joint_range(space, 0); }

/* bodies [first, nBodies) appended behind the ones the device holds */
static int
append_bodies(cpSpace *space, int first)
{
	int n = space->nBodies - first;
	cpb200_body_desc *descs = (cpb200_body_desc *)cpcalloc((size_t)(n > 0 ? n : 1), sizeof(cpb200_body_desc));
	for(int i = 0; i < n; i++) fill_body_desc(&descs[i], space->bodies[first + i]);
	int rc = cpb200_world_append_bodies(space->world, n, descs);
	cpfree(descs);
	if(rc < 0) cpEngineError("body upload");
	if(rc == 0) space->nBodiesOnDevice = space->nBodies;
	return rc;
}

/* Everything added since the last sync, as appended ranges.  Returns cpFalse if some array on the device has no room
 * left (nothing is lost: the caller takes the full re-upload, which allocates new slack). */
static cpBool
append_to_device(cpSpace *space)
{
	if(space->nBodies > space->nBodiesOnDevice && append_bodies(space, space->nBodiesOnDevice) != 0) return cpFalse;
	if(space->nShapes > space->nShapesOnDevice && upload_shape_range(space, space->nShapesOnDevice) != 0) return cpFalse;
	if(space->nConstraints > space->nConstraintsOnDevice && upload_joint_range(space, space->nConstraintsOnDevice) != 0) return cpFalse;
	return cpTrue;
}

static void
upload_params(cpSpace *space)
{
	cpb200_space_params p;
	memset(&p, 0, sizeof(p));
	p.gravity[0] = space->gravity.x; p.gravity[1] = space->gravity.y;
	p.damping = space->damping;
	p.idle_speed_threshold = space->idleSpeedThreshold;
	p.sleep_time_threshold = space->sleepTimeThreshold;
	p.collision_slop = space->collisionSlop;
	p.collision_bias = space->collisionBias;
	p.collision_persistence = space->collisionPersistence;
	p.iterations = space->iterations;
	if(cpb200_world_set_space_params(space->world, 0, &p)) cpEngineError("parameter upload");
	if(cpb200_world_set_solver_mode(space->world, space->solverMode)) cpEngineError("solver mode");
}

static void
sync_to_device(cpSpace *space)
{
	ensure_world(space);
	if(space->appendDirty && !space->topologyDirty){
		/* f4: additions travel as appended ranges; nothing that is on the device is touched */
		if(!append_to_device(space)) space->topologyDirty = cpTrue;
	}
	space->appendDirty = cpFalse;
	if(space->topologyDirty){
		cpSpaceFetchBodiesB200(space);
		cpSpaceFetchBiasB200(space);
		if(space->jointStale) cpSpaceFetchJointsB200(space);
		upload_bodies(space, cpTrue);
		upload_shapes(space);
		upload_joints(space);
		space->topologyDirty = cpFalse;
		space->bodiesDirty = cpFalse;
		space->forcesDirty = cpFalse;
		space->shapesDirty = space->jointsDirty = cpFalse;
		space->shapeIndexDirty = cpFalse;
	} else if(space->shapesDirty || space->jointsDirty){
		/* parameters of shapes / constraints only (host slots still equal device indices: removals set topologyDirty) */
		if(space->bodiesDirty){ cpSpaceFetchBodiesB200(space); cpSpaceFetchBiasB200(space); upload_bodies(space, cpFalse); }
		else if(space->forcesDirty){ cpSpaceUnpackAllB200(space); upload_forces(space); }
		if(space->shapesDirty) upload_shapes(space);
		if(space->jointsDirty){ if(space->jointStale) cpSpaceFetchJointsB200(space); upload_joints(space); }
		space->bodiesDirty = space->forcesDirty = space->shapesDirty = space->jointsDirty = cpFalse;
	} else if(space->bodiesDirty){
		cpSpaceFetchBodiesB200(space);
		cpSpaceFetchBiasB200(space);
		upload_bodies(space, cpFalse);
		space->bodiesDirty = cpFalse;
		space->forcesDirty = cpFalse;
	} else if(space->forcesDirty){
		cpSpaceUnpackAllB200(space);   /* forces the last step consumed are zeroed in the mirrors first */
		upload_forces(space);
		space->forcesDirty = cpFalse;
	}
	if(space->touchDirty){
		/* bodies activated while awake (cpBodySetForce etc. call cpBodyActivate): restart their idle timers */
		int n = space->nBodies, m = 0;
		int32_t *idx = (int32_t *)cpcalloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
		for(int i = 0; i < n; i++){ cpBody *b = space->bodies[i]; if(b->idleReset){ b->idleReset = cpFalse; idx[m++] = i; } }
		if(cpb200_world_touch_bodies(space->world, m, idx)) cpEngineError("idle timer reset");
		cpfree(idx);
		space->touchDirty = cpFalse;
	}
	if(space->paramsDirty){ upload_params(space); space->paramsDirty = cpFalse; }
}

/* make the device world current with every host-side edit (used by the query entry points) */
void cpSpacePrepareDeviceB200(cpSpace *space){ sync_to_device(space); }

/* one transfer: the state of every body into the space's page-locked buffer */
void
cpSpaceDownloadBodiesB200(cpSpace *space)
{
	if(!space->hostStale || !space->world){ space->hostStale = cpFalse; return; }
	space->hostStale = cpFalse;
	int n = (space->nBodies < space->nBodiesOnDevice ? space->nBodies : space->nBodiesOnDevice);   /* bodies added since the last sync are not there yet */
	cpb200_body_state *st = (cpb200_body_state *)xfer_buffer(&space->xferStates, &space->xferStatesBytes, (size_t)n*sizeof(cpb200_body_state));
	if(cpb200_world_get_bodies(space->world, 0, n, st)) cpEngineError("body download");
	space->fetchStamp++;
	space->someMirrorsStale = cpTrue;
}

/* this body's record of the last download into its mirror */
void
cpBodyUnpackB200(cpBody *b)
{
	cpSpace *space = b->space;
	b->fetchStamp = space->fetchStamp;
	if(b->idleTime == INFINITY || space->xferStates == NULL || b->index < 0 || b->index >= space->nBodiesOnDevice) return; /* static bodies never change on the device; a body added since the last sync is not there yet */
	const cpb200_body_state *s = (const cpb200_body_state *)space->xferStates + b->index;
	b->p = cpv(s->p[0], s->p[1]);
	b->v = cpv(s->v[0], s->v[1]);
	b->a = s->a;
	b->w = s->w;
	cpVect rot = cpv(s->rot[0], s->rot[1]), c = b->cog;
	b->transform = cpTransformNewTranspose(
		rot.x, -rot.y, b->p.x - (c.x*rot.x - c.y*rot.y),
		rot.y,  rot.x, b->p.y - (c.x*rot.y + c.y*rot.x));
	b->idleTime = s->idle_time;
	b->sleepRoot = (s->sleeping && s->sleep_group >= 0 && s->sleep_group < space->nBodies ? space->bodies[s->sleep_group] : NULL);
	if(!s->sleeping){
		/* the step consumed the forces and the bias velocities (cpBody.c:505-507, 518-519); a force set after
		 * the step went through cpBodyActivate, i.e. through this function, BEFORE it was written */
		if(b->m != INFINITY){ b->f = cpvzero; b->t = 0.0; }
		b->v_bias = cpvzero; b->w_bias = 0.0;
	}
}

/* every mirror current (before anything walks all bodies: uploads, structural edits, arbiter threading) */
void
cpSpaceUnpackAllB200(cpSpace *space)
{
	if(!space->someMirrorsStale) return;
	int n = space->nBodies;
	const unsigned stamp = space->fetchStamp;
	cpBody **bodies = space->bodies;
	#pragma omp parallel for schedule(static) num_threads(host_threads(space, n))
	for(int i = 0; i < n; i++){ if(bodies[i]->fetchStamp != stamp) cpBodyUnpackB200(bodies[i]); }
	space->someMirrorsStale = cpFalse;
}

void
cpSpaceFetchBodiesB200(cpSpace *space)
{
	if(space->hostStale) cpSpaceDownloadBodiesB200(space);
	cpSpaceUnpackAllB200(space);
}

void cpSpaceSyncB200(cpSpace *space){ cpSpaceFetchBodiesB200(space); }

/* The bias velocities the last step's solver left for the next position update (cpBody.c:511-522) live only on
 * the device; the mirrors hold zero after a download.  Before the host re-uploads bodies between two steps
 * (an edited body, a structural change) it takes them back, so that the upload does not cancel the pending
 * penetration correction -- in the reference they simply stay in the cpBody across such edits.  Must run while
 * host slots still equal device indices, i.e. before a removal compacts space->bodies. */
void
cpSpaceFetchBiasB200(cpSpace *space)
{
	if(!space->biasStale || !space->world) return;
	cpSpaceFetchBodiesB200(space);
	space->biasStale = cpFalse;
	int n = space->nBodiesOnDevice;
	if(n > space->nBodies) n = space->nBodies;
	if(n <= 0) return;
	double *vb = (double *)cpcalloc((size_t)n, 3*sizeof(double));
	if(cpb200_world_get_body_bias(space->world, 0, n, vb)) cpEngineError("bias velocity download");
	for(int i = 0; i < n; i++){ cpBody *b = space->bodies[i]; b->v_bias = cpv(vb[3*i], vb[3*i + 1]); b->w_bias = vb[3*i + 2]; }
	cpfree(vb);
}

void
cpSpaceFetchBBsB200(cpSpace *space)
{
	space->bbStale = cpFalse;
	if(!space->world || space->nShapes == 0 || space->topologyDirty) return;
	int n = (space->nShapes < space->nShapesOnDevice ? space->nShapes : space->nShapesOnDevice);   /* shapes added since the last sync keep the box cpShapeUpdate gave them */
	if(n <= 0) return;
	double *bbs = (double *)cpcalloc((size_t)n, 4*sizeof(double));
	if(cpb200_world_get_shape_bbs(space->world, 0, n, bbs)) cpEngineError("AABB download");
	for(int i = 0; i < n; i++) space->shapes[i]->bb = cpBBNew(bbs[4*i], bbs[4*i + 1], bbs[4*i + 2], bbs[4*i + 3]);
	cpfree(bbs);
}

void
cpSpaceFetchJointsB200(cpSpace *space)
{
	space->jointStale = cpFalse;
	/* Host slots equal device indices until a removal compacts space->constraints (removals fetch first, see
	 * sync_before_edit); constraints added since the last upload sit behind the device's count.  A pending
	 * parameter edit or addition (topologyDirty) must NOT skip this fetch: the re-upload that follows would hand
	 * the device stale accumulated impulses and the joints would lose their warm start. */
	if(!space->world || space->jointIndexDirty) return;
	int n = (space->nConstraintsOnDevice < space->nConstraints ? space->nConstraintsOnDevice : space->nConstraints);
	if(n <= 0) return;
	cpb200_joint_state *st = (cpb200_joint_state *)cpcalloc((size_t)n, sizeof(cpb200_joint_state));
	if(cpb200_world_get_joints(space->world, 0, n, st)) cpEngineError("joint download");
	for(int i = 0; i < n; i++){
		cpConstraint *c = space->constraints[i];
		c->acc = cpv(st[i].acc[0], st[i].acc[1]);
		c->impulse = st[i].impulse;
		if(c->klass == CPB200_JOINT_RATCHET) c->prm[0] = st[i].aux;
	}
	cpfree(st);
}

void
cpSpaceFetchArbitersB200(cpSpace *space)
{
	space->arbStale = cpFalse;
	if(space->shapeIndexDirty) return;   /* device records cannot be mapped to host objects any more: the mirrors fetched before the removal stay */
	/* remember the user data of the outgoing mirrors */
	space->nArbData = 0;
	for(int i = 0; i < space->nArbs; i++){
		cpArbiter *arb = &space->arbs[i];
		if(arb->data == NULL || arb->state == CP_ARBITER_STATE_CACHED) continue;   /* separated pairs drop theirs */
		space->arbData = (struct cpArbData *)grow(space->arbData, &space->capArbData, space->nArbData + 1, sizeof(struct cpArbData));
		cpHashValue ha = arb->a->hashid, hb = arb->b->hashid;
		struct cpArbData *d = &space->arbData[space->nArbData++];
		d->lo = (ha < hb ? ha : hb); d->hi = (ha < hb ? hb : ha); d->data = arb->data;
	}
	for(int i = 0; i < space->nBodies; i++) space->bodies[i]->firstArb = -1;
	space->nArbs = 0;
	if(!space->world) return;
	/* everything in the cache: active, dormant (sleeping) and cached-for-persistence records */
	int n = cpb200_world_get_arbiters(space->world, 0, NULL, 0);
	if(n < 0) cpEngineError("arbiter download");
	if(n == 0) return;
	cpb200_arbiter *recs = (cpb200_arbiter *)cpcalloc((size_t)n, sizeof(cpb200_arbiter));
	n = cpb200_world_get_arbiters(space->world, n, recs, 0);
	if(n < 0) cpEngineError("arbiter download");
	space->arbs = (cpArbiter *)grow(space->arbs, &space->capArbs, n, sizeof(cpArbiter));
	cpSpaceFetchBodiesB200(space);
	for(int i = 0; i < n; i++){
		const cpb200_arbiter *r = &recs[i];
		if(r->shape_a < 0 || r->shape_a >= space->nShapesOnDevice || r->shape_a >= space->nShapes || r->shape_b < 0 || r->shape_b >= space->nShapesOnDevice || r->shape_b >= space->nShapes) continue;
		cpArbiter *arb = &space->arbs[space->nArbs];
		memset(arb, 0, sizeof(*arb));
		arb->space = space;
		arb->a = space->shapes[r->shape_a]; arb->b = space->shapes[r->shape_b];
		arb->body_a = arb->a->body; arb->body_b = arb->b->body;
		arb->e = r->e; arb->u = r->u;
		arb->surface_vr = cpv(r->surface_vr[0], r->surface_vr[1]);
		arb->n = cpv(r->n[0], r->n[1]);
		arb->count = r->count;
		arb->state = r->state;
		arb->stamp = r->stamp;
		arb->active = r->active;
		arb->record = r->record;
		if(space->nArbData > 0){
			cpHashValue ha = arb->a->hashid, hb = arb->b->hashid, lo = (ha < hb ? ha : hb), hi = (ha < hb ? hb : ha);
			for(int k = 0; k < space->nArbData; k++){ if(space->arbData[k].lo == lo && space->arbData[k].hi == hi){ arb->data = space->arbData[k].data; break; } }
		}
		for(int k = 0; k < 2; k++){
			struct cpContact *c = &arb->contacts[k];
			c->r1 = cpv(r->contacts[k].r1[0], r->contacts[k].r1[1]);
			c->r2 = cpv(r->contacts[k].r2[0], r->contacts[k].r2[1]);
			c->nMass = r->contacts[k].n_mass; c->tMass = r->contacts[k].t_mass;
			c->bounce = r->contacts[k].bounce; c->bias = r->contacts[k].bias;
			c->jnAcc = r->contacts[k].jn_acc; c->jtAcc = r->contacts[k].jt_acc; c->jBias = r->contacts[k].j_bias;
			c->hash = (cpHashValue)r->contacts[k].hash;
		}
		cpCollisionHandler *h = lookup_handler(space, arb->a->type, arb->b->type);
		arb->handler = h;
		/* cpArbiterUpdate (cpArbiter.c:402-404) */
		arb->swapped = (arb->a->type != h->typeA && h->typeA != CP_WILDCARD_COLLISION_TYPE);
		arb->next_a = arb->next_b = -1;
		/* thread onto both bodies: only arbiters that take part in the contact graph (active this step,
		 * or dormant ones of sleeping bodies: r->count is kept for those) */
		if(r->active || (r->count > 0 && r->state != CP_ARBITER_STATE_CACHED && (arb->body_a->sleepRoot || arb->body_b->sleepRoot))){
			int k = space->nArbs;
			arb->next_a = arb->body_a->firstArb; arb->body_a->firstArb = k;
			arb->next_b = arb->body_b->firstArb; arb->body_b->firstArb = k;
		}
		space->nArbs++;
	}
	cpfree(recs);
}

void
cpSpaceCollidePairB200(cpSpace *space, const cpShape *a, const cpShape *b, cpContactPointSet *out)
{
	sync_to_device(space);
	double buf[13];
	if(cpb200_world_collide_pair(space->world, a->index, b->index, buf) < 0) cpEngineError("cpShapesCollide");
	out->count = (int)buf[0];
	out->normal = cpv(buf[1], buf[2]);
	for(int k = 0; k < out->count && k < CP_MAX_CONTACTS_PER_ARBITER; k++){
		out->points[k].pointA = cpv(buf[3 + 5*k], buf[4 + 5*k]);
		out->points[k].pointB = cpv(buf[5 + 5*k], buf[6 + 5*k]);
		out->points[k].distance = buf[7 + 5*k];
	}
}

/* ---- callbacks that need downloaded state ---- */
static cpBool
space_has_collision_callbacks(const cpSpace *space)
{
	return (space->nHandlers > 0 || space->hasDefaultHandler);
}

/* Between the two halves of the step (the point where the reference runs them, cpSpaceStep.c:257-285): the
 * device has produced this step's arbiters; begin/preSolve decide which of them are solved and may change
 * their material or contacts.  The decisions go back as a list of edits. */
static void
run_begin_presolve_callbacks(cpSpace *space)
{
	cpSpaceFetchArbitersB200(space);
	int nEdits = 0;
	cpb200_arbiter_edit *edits = (cpb200_arbiter_edit *)cpcalloc((size_t)(space->nArbs > 0 ? space->nArbs : 1), sizeof(cpb200_arbiter_edit));
	for(int i = 0; i < space->nArbs; i++){
		cpArbiter *arb = &space->arbs[i];
		cpCollisionHandler *h = arb->handler;
		if(arb->stamp != space->stamp) continue;
		const cpArbiter before = *arb;
		uint32_t flags = 0;
		if(arb->state == CP_ARBITER_STATE_FIRST_COLLISION && !h->beginFunc(arb, space, h->userData)) cpArbiterIgnore(arb);
		if(arb->state != CP_ARBITER_STATE_IGNORE && !h->preSolveFunc(arb, space, h->userData)) flags |= CPB200_EDIT_REJECT;
		if(arb->state == CP_ARBITER_STATE_IGNORE && before.state != CP_ARBITER_STATE_IGNORE) flags |= CPB200_EDIT_IGNORE;
		if(arb->e != before.e || arb->u != before.u || !cpveql(arb->surface_vr, before.surface_vr)) flags |= CPB200_EDIT_MATERIAL;
		if(!cpveql(arb->n, before.n)) flags |= CPB200_EDIT_CONTACTS;
		for(int k = 0; k < arb->count && k < CP_MAX_CONTACTS_PER_ARBITER; k++){
			if(!cpveql(arb->contacts[k].r1, before.contacts[k].r1) || !cpveql(arb->contacts[k].r2, before.contacts[k].r2)) flags |= CPB200_EDIT_CONTACTS;
		}
		if(!flags) continue;
		cpb200_arbiter_edit *ed = &edits[nEdits++];
		ed->record = arb->record; ed->flags = flags;
		ed->e = arb->e; ed->u = arb->u; ed->surface_vr[0] = arb->surface_vr.x; ed->surface_vr[1] = arb->surface_vr.y;
		ed->n[0] = arb->n.x; ed->n[1] = arb->n.y;
		for(int k = 0; k < CP_MAX_CONTACTS_PER_ARBITER; k++){
			ed->r1[k][0] = arb->contacts[k].r1.x; ed->r1[k][1] = arb->contacts[k].r1.y;
			ed->r2[k][0] = arb->contacts[k].r2.x; ed->r2[k][1] = arb->contacts[k].r2.y;
		}
	}
	if(cpb200_world_edit_arbiters(space->world, nEdits, edits)) cpEngineError("arbiter edits");
	cpfree(edits);
}

/* After the solver: separate for the pairs that stopped touching this step (cpSpaceStep.c:309-314), then
 * postSolve for everything that was solved (cpSpaceStep.c:438-443). */
static void
run_separate_postsolve_callbacks(cpSpace *space)
{
	cpSpaceFetchArbitersB200(space);
	for(int i = 0; i < space->nArbs; i++){
		cpArbiter *arb = &space->arbs[i];
		cpCollisionHandler *h = arb->handler;
		/* became cached this step <=> last touched on the previous stamp */
		if(arb->state == CP_ARBITER_STATE_CACHED && arb->stamp + 1 == space->stamp) h->separateFunc(arb, space, h->userData);
	}
	for(int i = 0; i < space->nArbs; i++){
		cpArbiter *arb = &space->arbs[i];
		cpCollisionHandler *h = arb->handler;
		if(arb->state != CP_ARBITER_STATE_CACHED && arb->stamp == space->stamp && arb->active == 1) h->postSolveFunc(arb, space, h->userData);
	}
}

/* ---- user callbacks inside the step (slow path) ----
 * Bodies with a custom position_func / velocity_func and springs with a custom force function are few (demo/Planet.c,
 * demo/Springies.c); the step of such a space is split at the points where the reference makes those calls
 * (cpSpaceStep.c:362-367 position, cpDampedSpring.c:50 inside the constraint prestep, cpSpaceStep.c:398-404 velocity)
 * and only the flagged objects make the round trip through the host. */
typedef struct cpCustomWork { int nPos, nVel, nSpring; int32_t *pos, *vel, *spring; } cpCustomWork;

static void
custom_collect(cpSpace *space, cpCustomWork *cw)
{
	memset(cw, 0, sizeof(*cw));
	if(!space->anyCustom) return;
	cw->pos = (int32_t *)cpcalloc((size_t)space->nBodies + 1, sizeof(int32_t));
	cw->vel = (int32_t *)cpcalloc((size_t)space->nBodies + 1, sizeof(int32_t));
	cw->spring = (int32_t *)cpcalloc((size_t)space->nConstraints + 1, sizeof(int32_t));
	for(int i = 0; i < space->nBodies; i++){
		cpBody *b = space->bodies[i];
		if(b->idleTime == INFINITY) continue;                                     /* static: never integrated (cpSpace.c:447) */
		if(b->position_func != cpBodyUpdatePosition) cw->pos[cw->nPos++] = i;
		if(b->velocity_func != cpBodyUpdateVelocity && b->m != INFINITY) cw->vel[cw->nVel++] = i;
	}
	for(int i = 0; i < space->nConstraints; i++){
		cpConstraint *c = space->constraints[i];
		if(c->forceFunc && (c->klass == CPB200_JOINT_DAMPED_SPRING || c->klass == CPB200_JOINT_DAMPED_ROTARY_SPRING)) cw->spring[cw->nSpring++] = i;
	}
}

static void custom_free(cpCustomWork *cw){ cpfree(cw->pos); cpfree(cw->vel); cpfree(cw->spring); }

/* Every mirror current with the device (the user's callback may read any body), the listed bodies' forces untouched:
 * an ordinary download assumes the step has consumed them (cpBodyUnpackB200), here the step is still in progress. */
static void
custom_fetch_bodies(cpSpace *space, int n, const int32_t *idx)
{
	if(n == 0) return;
	double *keep = (double *)cpcalloc((size_t)n, 3*sizeof(double));
	for(int k = 0; k < n; k++){ const cpBody *b = space->bodies[idx[k]]; keep[3*k] = b->f.x; keep[3*k + 1] = b->f.y; keep[3*k + 2] = b->t; }
	space->hostStale = cpTrue;
	cpSpaceFetchBodiesB200(space);
	for(int k = 0; k < n; k++){ cpBody *b = space->bodies[idx[k]]; b->f = cpv(keep[3*k], keep[3*k + 1]); b->t = keep[3*k + 2]; }
	cpfree(keep);
}

/* cpSpaceStep.c:362-367: position_func of every awake body, before anything else */
static void
custom_positions(cpSpace *space, cpCustomWork *cw, cpFloat dt)
{
	if(cw->nPos == 0) return;
	if(space->world){
		/* the callbacks integrate the mirrors: those must hold the device's latest state, pending bias velocities included */
		cpSpaceFetchBodiesB200(space);
		cpSpaceFetchBiasB200(space);
	}
	space->locked++;
	for(int k = 0; k < cw->nPos; k++){
		cpBody *b = space->bodies[cw->pos[k]];
		if(b->sleepRoot) continue;
		b->position_func(b, dt);
	}
	space->locked--;
	space->bodiesDirty = cpTrue;     /* the integrated mirrors travel with the ordinary body upload */
}

/* cpDampedSpring.c:36-52 / cpDampedRotarySpring.c:36-48: the argument the reference hands to the user's function */
static void
custom_springs(cpSpace *space, cpCustomWork *cw)
{
	if(cw->nSpring == 0) return;
	int nb = 0;
	int32_t *bidx = (int32_t *)cpcalloc((size_t)2*cw->nSpring, sizeof(int32_t));
	for(int k = 0; k < cw->nSpring; k++){ cpConstraint *c = space->constraints[cw->spring[k]]; bidx[nb++] = c->a->index; bidx[nb++] = c->b->index; }
	custom_fetch_bodies(space, nb, bidx);           /* positions after this step's position update */
	cpfree(bidx);
	double *f = (double *)cpcalloc((size_t)cw->nSpring, sizeof(double));
	space->locked++;
	for(int k = 0; k < cw->nSpring; k++){
		cpConstraint *c = space->constraints[cw->spring[k]];
		cpBody *a = c->a, *b = c->b;
		if(c->klass == CPB200_JOINT_DAMPED_SPRING){
			cpVect r1 = cpTransformVect(a->transform, cpvsub(c->anchorA, a->cog));
			cpVect r2 = cpTransformVect(b->transform, cpvsub(c->anchorB, b->cog));
			cpFloat dist = cpvlength(cpvsub(cpvadd(b->p, r2), cpvadd(a->p, r1)));
			f[k] = ((cpDampedSpringForceFunc)c->forceFunc)(c, dist);
		} else {
			f[k] = ((cpDampedRotarySpringTorqueFunc)c->forceFunc)(c, a->a - b->a);
		}
	}
	space->locked--;
	if(cpb200_world_set_spring_forces(space->world, cw->nSpring, cw->spring, f)) cpEngineError("spring force upload");
	cpfree(f);
}

/* cpSpaceStep.c:398-404: velocity_func(body, gravity, damping^dt, dt) between the prestep and the solver */
static void
custom_velocities(cpSpace *space, cpCustomWork *cw, cpFloat dt)
{
	if(cw->nVel == 0) return;
	custom_fetch_bodies(space, cw->nVel, cw->vel);
	const cpFloat damping = cpfpow(space->damping, dt);
	double *v = (double *)cpcalloc((size_t)cw->nVel, 3*sizeof(double));
	const cpBool dirty = space->bodiesDirty, forces = space->forcesDirty;
	space->locked++;
	for(int k = 0; k < cw->nVel; k++){
		cpBody *b = space->bodies[cw->vel[k]];
		if(!b->sleepRoot) b->velocity_func(b, space->gravity, damping, dt);
		v[3*k] = b->v.x; v[3*k + 1] = b->v.y; v[3*k + 2] = b->w;
	}
	space->locked--;
	space->bodiesDirty = dirty; space->forcesDirty = forces;   /* the result goes straight to the device below */
	if(cpb200_world_set_body_velocities_indexed(space->world, cw->nVel, cw->vel, v)) cpEngineError("velocity upload (custom integrator)");
	cpfree(v);
}

/* ---- the step ---- */
static void
step_once(cpSpace *space, cpFloat dt, cpBool callbacks)
{
	/* constraint preSolve callbacks run after the collision phase and the islands pass, before the prestep
	 * (cpSpaceStep.c:389-396): a space that has one steps in parts */
	cpBool constraintPreSolve = cpFalse;
	if(callbacks){ for(int i = 0; i < space->nConstraints; i++){ if(space->constraints[i]->preSolve){ constraintPreSolve = cpTrue; break; } } }
	cpCustomWork cw;
	custom_collect(space, &cw);
	custom_positions(space, &cw, dt);
	sync_to_device(space);
	const cpBool handlers = space_has_collision_callbacks(space);
	const cpBool midstep = (cw.nSpring > 0 || cw.nVel > 0 || constraintPreSolve);
	if(handlers || midstep){
		/* split step: the handlers' return values and edits take effect in THIS step, like the reference */
		if(cpb200_world_step_collide(space->world, dt)) cpEngineError("cpSpaceStep (collision phase)");
		space->stamp++;
		space->curr_dt = dt;
		space->hostStale = cpTrue; space->bbStale = cpTrue; space->arbStale = cpTrue;
		if(handlers){
			space->locked++;
			run_begin_presolve_callbacks(space);
			space->locked--;
		}
		if(midstep){
			if(constraintPreSolve){
				space->locked++;
				for(int i = 0; i < space->nConstraints; i++){
					cpConstraint *c = space->constraints[i];
					if(c->preSolve) c->preSolve(c, space);
				}
				space->locked--;
				/* what the callback changed (cpSimpleMotorSetRate, cpConstraintSetMaxForce, ...) must reach this step's prestep */
				if(space->jointsDirty){ if(space->jointStale) cpSpaceFetchJointsB200(space); upload_joints(space); space->jointsDirty = cpFalse; }
				if(space->bodiesDirty){ cpSpaceFetchBodiesB200(space); upload_bodies(space, cpFalse); space->bodiesDirty = space->forcesDirty = cpFalse; }
				else if(space->forcesDirty){ cpSpaceUnpackAllB200(space); upload_forces(space); space->forcesDirty = cpFalse; }
			}
			custom_springs(space, &cw);
			if(cpb200_world_step_presolve(space->world)) cpEngineError("cpSpaceStep (prestep phase)");
			custom_velocities(space, &cw, dt);
		}
		if(cpb200_world_step_finish(space->world)) cpEngineError("cpSpaceStep (solver phase)");
	} else {
		if(cpb200_world_step(space->world, dt)) cpEngineError("cpSpaceStep");
		space->stamp++;
		space->curr_dt = dt;
	}
	if(space->anyCustom) custom_free(&cw);
	space->hostStale = cpTrue;
	space->biasStale = cpTrue;
	space->bbStale = cpTrue;
	space->arbStale = cpTrue;
	space->jointStale = cpTrue;
	if(callbacks || handlers){
		space->locked++;
		cpBool any = cpFalse;
		for(int i = 0; i < space->nConstraints; i++){ if(space->constraints[i]->postSolve){ any = cpTrue; break; } }
		if(any){
			cpSpaceFetchJointsB200(space);
			for(int i = 0; i < space->nConstraints; i++){
				cpConstraint *c = space->constraints[i];
				if(c->postSolve) c->postSolve(c, space);
			}
		}
		if(handlers) run_separate_postsolve_callbacks(space);
		space->locked--;
		run_post_step_callbacks(space);
	}
}

void
cpSpaceStep(cpSpace *space, cpFloat dt)
{
	if(dt == 0.0) return; /* cpSpaceStep.c:339 */
	cpAssertHard(space->locked == 0, "cpSpaceStep() cannot be called from inside a callback of the same space.");
	step_once(space, dt, cpTrue);
}

void
cpSpaceStepManyB200(cpSpace *space, cpFloat dt, int n)
{
	if(dt == 0.0) return;
	for(int i = 0; i < n; i++) step_once(space, dt, (i == n - 1));
}

/* ---- cpHastySpace (reference cpHastySpace.c:513-700) ---- */
cpSpace *
cpHastySpaceNew(void)
{
	cpSpace *space = cpSpaceNew();
	space->hasty = cpTrue;
	space->hastyThreads = 1;
	return space;
}

void cpHastySpaceFree(cpSpace *space){ cpSpaceFree(space); }

void
cpHastySpaceSetThreads(cpSpace *space, unsigned long threads)
{
	/* recorded only: the solver already runs on all SMs of the device */
	space->hastyThreads = (threads == 0 ? 1 : threads);
}

unsigned long cpHastySpaceGetThreads(cpSpace *space){ return space->hastyThreads; }
void cpHastySpaceStep(cpSpace *space, cpFloat dt){ cpSpaceStep(space, dt); }
