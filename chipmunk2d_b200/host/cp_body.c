/* cp_body.c -- host mirror of rigid bodies (public API of reference cpBody.h:49-187).
 *
 * The mirror holds what the user set and what was last downloaded from the device.  Getters call
 * cpBodySyncForRead() first, setters additionally mark the space so the new state is uploaded
 * before the next step.  Semantics follow reference src/cpBody.c (type encoding :136-201, mass
 * accumulation :206-239, setters :241-467, integrators :493-522) and the sleep API of
 * src/cpSpaceComponent.c:113-349.
 */
#include <stdlib.h>
#include <string.h>
#include "cp_host.h"

/* 64-byte aligned, so that the struct's "lines" (cp_host.h) are cache lines; released with cpfree like any other object */
cpBody *
cpBodyAlloc(void)
{
	void *mem = NULL;
	if(posix_memalign(&mem, 64, (sizeof(cpBody) + 63) & ~(size_t)63)) return NULL;
	memset(mem, 0, sizeof(cpBody));
	return (cpBody *)mem;
}

static void
touch(cpBody *body)
{
	/* a host-side edit of kinematic state: fetch first so unrelated fields are not rolled back */
	cpBodySyncForRead(body);
	if(body->space) body->space->bodiesDirty = cpTrue;
}

static void
touch_force(cpBody *body)
{
	/* forces are consumed by the step and never written back by the device: no fetch needed */
	if(body->space) body->space->forcesDirty = cpTrue;
}

void
cpBodySetTransformInternal(cpBody *body, cpVect p, cpFloat a)
{
	/* SetTransform (cpBody.c:347-357) */
	cpVect rot = cpvforangle(a);
	cpVect c = body->cog;
	body->transform = cpTransformNewTranspose(
		rot.x, -rot.y, p.x - (c.x*rot.x - c.y*rot.y),
		rot.y,  rot.x, p.y - (c.x*rot.y + c.y*rot.x));
}

cpBody *
cpBodyInit(cpBody *body, cpFloat mass, cpFloat moment)
{
	body->space = NULL;
	body->shapeList = NULL;
	body->constraintList = NULL;
	body->velocity_func = cpBodyUpdateVelocity;
	body->position_func = cpBodyUpdatePosition;
	body->sleepRoot = NULL;
	body->idleTime = 0.0;
	body->p = cpvzero; body->v = cpvzero; body->f = cpvzero;
	body->w = 0.0; body->t = 0.0;
	body->v_bias = cpvzero; body->w_bias = 0.0;
	body->cog = cpvzero;
	body->userData = NULL;
	body->index = -1;
	body->firstArb = -1;
	body->m = body->i = 0.0;
	body->m_inv = body->i_inv = INFINITY;
	cpBodySetMass(body, mass);
	cpBodySetMoment(body, moment);
	cpBodySetAngle(body, 0.0);
	return body;
}

cpBody *cpBodyNew(cpFloat mass, cpFloat moment){ return cpBodyInit(cpBodyAlloc(), mass, moment); }

cpBody *
cpBodyNewKinematic(void)
{
	cpBody *body = cpBodyNew(0.0, 0.0);
	cpBodySetType(body, CP_BODY_TYPE_KINEMATIC);
	return body;
}

cpBody *
cpBodyNewStatic(void)
{
	cpBody *body = cpBodyNew(0.0, 0.0);
	cpBodySetType(body, CP_BODY_TYPE_STATIC);
	return body;
}

void cpBodyDestroy(cpBody *body){ (void)body; }
void cpBodyFree(cpBody *body){ if(body){ cpBodyDestroy(body); cpfree(body); } }

cpBool
cpBodyIsSleeping(const cpBody *body)
{
	cpBodySyncForRead(body);
	return (body->sleepRoot != NULL);
}

cpBodyType
cpBodyGetType(cpBody *body)
{
	if(body->idleTime == INFINITY) return CP_BODY_TYPE_STATIC;
	if(body->m == INFINITY) return CP_BODY_TYPE_KINEMATIC;
	return CP_BODY_TYPE_DYNAMIC;
}

void
cpBodySetType(cpBody *body, cpBodyType type)
{
	cpBodySyncForRead(body);
	cpBodyType oldType = cpBodyGetType(body);
	if(oldType == type) return;
	body->idleTime = (type == CP_BODY_TYPE_STATIC ? (cpFloat)INFINITY : 0.0);
	if(type == CP_BODY_TYPE_DYNAMIC){
		body->m = body->i = 0.0;
		body->m_inv = body->i_inv = INFINITY;
		cpBodyAccumulateMassFromShapes(body);
	} else {
		body->m = body->i = INFINITY;
		body->m_inv = body->i_inv = 0.0;
		body->v = cpvzero;
		body->w = 0.0;
	}
	if(body->space){
		cpAssertSpaceUnlocked(body->space);
		if(oldType != CP_BODY_TYPE_STATIC) cpBodyActivate(body);
		cpSpaceMarkBodyDirtyB200(body);
	}
}

void
cpBodyAccumulateMassFromShapes(cpBody *body)
{
	if(body == NULL || cpBodyGetType(body) != CP_BODY_TYPE_DYNAMIC) return;
	body->m = body->i = 0.0;
	body->cog = cpvzero;
	cpVect pos = cpBodyGetPosition(body);
	for(cpShape *shape = body->shapeList; shape; shape = shape->next){
		struct cpShapeMassInfo *info = &shape->massInfo;
		cpFloat m = info->m;
		if(m > 0.0){
			cpFloat msum = body->m + m;
			body->i += m*info->i + cpvdistsq(body->cog, info->cog)*(m*body->m)/msum;
			body->cog = cpvlerp(body->cog, info->cog, m/msum);
			body->m = msum;
		}
	}
	body->m_inv = 1.0/body->m;
	body->i_inv = 1.0/body->i;
	cpBodySetPosition(body, pos);
	if(body->space) cpSpaceMarkBodyDirtyB200(body);
}

cpSpace *cpBodyGetSpace(const cpBody *body){ return body->space; }
cpFloat cpBodyGetMass(const cpBody *body){ return body->m; }

void
cpBodySetMass(cpBody *body, cpFloat mass)
{
	cpAssertHard(cpBodyGetType(body) == CP_BODY_TYPE_DYNAMIC, "You cannot set the mass of kinematic or static bodies.");
	cpAssertHard(0.0 <= mass && mass < INFINITY, "Mass must be positive and finite.");
	cpBodyActivate(body);
	body->m = mass;
	body->m_inv = (mass == 0.0 ? (cpFloat)INFINITY : 1.0/mass);
	if(body->space) cpSpaceMarkBodyDirtyB200(body);
}

cpFloat cpBodyGetMoment(const cpBody *body){ return body->i; }

void
cpBodySetMoment(cpBody *body, cpFloat moment)
{
	cpAssertHard(moment >= 0.0, "Moment of Inertia must be positive.");
	cpBodyActivate(body);
	body->i = moment;
	body->i_inv = (moment == 0.0 ? (cpFloat)INFINITY : 1.0/moment);
	if(body->space) cpSpaceMarkBodyDirtyB200(body);
}

cpVect cpBodyGetRotation(const cpBody *body){ cpBodySyncForRead(body); return cpv(body->transform.a, body->transform.b); }

/* shape / constraint lists (cpBody.c:287-342) */
void
cpBodyAddShape(cpBody *body, cpShape *shape)
{
	cpShape *next = body->shapeList;
	if(next) next->prev = shape;
	shape->next = next;
	shape->prev = NULL;
	body->shapeList = shape;
	if(shape->massInfo.m > 0.0) cpBodyAccumulateMassFromShapes(body);
}

void
cpBodyRemoveShape(cpBody *body, cpShape *shape)
{
	cpShape *prev = shape->prev, *next = shape->next;
	if(prev) prev->next = next; else body->shapeList = next;
	if(next) next->prev = prev;
	shape->prev = NULL;
	shape->next = NULL;
	if(cpBodyGetType(body) == CP_BODY_TYPE_DYNAMIC && shape->massInfo.m > 0.0) cpBodyAccumulateMassFromShapes(body);
}

void
cpBodyAddConstraint(cpBody *body, cpConstraint *constraint)
{
	if(constraint->a == body) constraint->next_a = body->constraintList; else constraint->next_b = body->constraintList;
	body->constraintList = constraint;
}

void
cpBodyRemoveConstraint(cpBody *body, cpConstraint *constraint)
{
	cpConstraint **link = &body->constraintList;
	while(*link && *link != constraint) link = ((*link)->a == body ? &(*link)->next_a : &(*link)->next_b);
	if(*link) *link = cpConstraintNext(constraint, body);
	if(constraint->a == body) constraint->next_a = NULL; else constraint->next_b = NULL;
}

cpVect cpBodyGetPosition(const cpBody *body){ cpBodySyncForRead(body); return cpTransformPoint(body->transform, cpvzero); }

void
cpBodySetPosition(cpBody *body, cpVect position)
{
	cpBodyActivate(body);
	touch(body);
	cpVect p = body->p = cpvadd(cpTransformVect(body->transform, body->cog), position);
	cpBodySetTransformInternal(body, p, body->a);
}

cpVect cpBodyGetCenterOfGravity(const cpBody *body){ return body->cog; }

void
cpBodySetCenterOfGravity(cpBody *body, cpVect cog)
{
	cpBodyActivate(body);
	touch(body);
	body->cog = cog;
	if(body->space) cpSpaceMarkBodyDirtyB200(body);
}

cpVect cpBodyGetVelocity(const cpBody *body){ cpBodySyncForRead(body); return body->v; }
void cpBodySetVelocity(cpBody *body, cpVect velocity){ cpBodyActivate(body); touch(body); body->v = velocity; }
cpVect cpBodyGetForce(const cpBody *body){ cpBodySyncForRead(body); return body->f; }
void cpBodySetForce(cpBody *body, cpVect force){ cpBodyActivate(body); touch_force(body); body->f = force; }
cpFloat cpBodyGetAngle(const cpBody *body){ cpBodySyncForRead(body); return body->a; }

void
cpBodySetAngle(cpBody *body, cpFloat angle)
{
	cpBodyActivate(body);
	touch(body);
	body->a = angle;
	cpBodySetTransformInternal(body, body->p, angle);
}

cpFloat cpBodyGetAngularVelocity(const cpBody *body){ cpBodySyncForRead(body); return body->w; }
void cpBodySetAngularVelocity(cpBody *body, cpFloat angularVelocity){ cpBodyActivate(body); touch(body); body->w = angularVelocity; }
cpFloat cpBodyGetTorque(const cpBody *body){ cpBodySyncForRead(body); return body->t; }
void cpBodySetTorque(cpBody *body, cpFloat torque){ cpBodyActivate(body); touch_force(body); body->t = torque; }
cpDataPointer cpBodyGetUserData(const cpBody *body){ return body->userData; }
void cpBodySetUserData(cpBody *body, cpDataPointer userData){ body->userData = userData; }

/* The device integrates every body with the reference's default integrators (K1 / K9).  A body with a user-supplied
 * integrator (cpBody.c:482-491; demo/Planet.c) is skipped by those kernels: the host layer runs the callback on the
 * body's mirror at the point of the step where the reference calls it and moves the result to the device (cpSpaceStep's
 * slow path in cp_space.c).  Everything else about the body -- collisions, contacts, the solver -- stays on the device. */
void
cpBodySetVelocityUpdateFunc(cpBody *body, cpBodyVelocityFunc velocityFunc)
{
	body->velocity_func = (velocityFunc ? velocityFunc : cpBodyUpdateVelocity);
	if(body->space){ if(body->velocity_func != cpBodyUpdateVelocity) body->space->anyCustom = cpTrue; cpSpaceMarkBodyDirtyB200(body); }
}

void
cpBodySetPositionUpdateFunc(cpBody *body, cpBodyPositionFunc positionFunc)
{
	body->position_func = (positionFunc ? positionFunc : cpBodyUpdatePosition);
	if(body->space){ if(body->position_func != cpBodyUpdatePosition) body->space->anyCustom = cpTrue; cpSpaceMarkBodyDirtyB200(body); }
}

/* Host versions of the integrators for bodies the user steps by hand (cpBody.c:493-522). */
void
cpBodyUpdateVelocity(cpBody *body, cpVect gravity, cpFloat damping, cpFloat dt)
{
	if(cpBodyGetType(body) == CP_BODY_TYPE_KINEMATIC) return;
	touch(body);
	body->v = cpvadd(cpvmult(body->v, damping), cpvmult(cpvadd(gravity, cpvmult(body->f, body->m_inv)), dt));
	body->w = body->w*damping + body->t*body->i_inv*dt;
	body->f = cpvzero;
	body->t = 0.0;
}

void
cpBodyUpdatePosition(cpBody *body, cpFloat dt)
{
	touch(body);
	cpVect p = body->p = cpvadd(body->p, cpvmult(cpvadd(body->v, body->v_bias), dt));
	cpFloat a = body->a = body->a + (body->w + body->w_bias)*dt;
	cpBodySetTransformInternal(body, p, a);
	body->v_bias = cpvzero;
	body->w_bias = 0.0;
}

cpVect cpBodyLocalToWorld(const cpBody *body, const cpVect point){ cpBodySyncForRead(body); return cpTransformPoint(body->transform, point); }
cpVect cpBodyWorldToLocal(const cpBody *body, const cpVect point){ cpBodySyncForRead(body); return cpTransformPoint(cpTransformRigidInverse(body->transform), point); }

void
cpBodyApplyForceAtWorldPoint(cpBody *body, cpVect force, cpVect point)
{
	cpBodyActivate(body);
	cpBodySyncForRead(body);
	touch_force(body);
	body->f = cpvadd(body->f, force);
	cpVect r = cpvsub(point, cpTransformPoint(body->transform, body->cog));
	body->t += cpvcross(r, force);
}

void
cpBodyApplyForceAtLocalPoint(cpBody *body, cpVect force, cpVect point)
{
	cpBodySyncForRead(body);
	cpBodyApplyForceAtWorldPoint(body, cpTransformVect(body->transform, force), cpTransformPoint(body->transform, point));
}

void
cpBodyApplyImpulseAtWorldPoint(cpBody *body, cpVect impulse, cpVect point)
{
	cpBodyActivate(body);
	touch(body);
	cpVect r = cpvsub(point, cpTransformPoint(body->transform, body->cog));
	body->v = cpvadd(body->v, cpvmult(impulse, body->m_inv));
	body->w += body->i_inv*cpvcross(r, impulse);
}

void
cpBodyApplyImpulseAtLocalPoint(cpBody *body, cpVect impulse, cpVect point)
{
	cpBodySyncForRead(body);
	cpBodyApplyImpulseAtWorldPoint(body, cpTransformVect(body->transform, impulse), cpTransformPoint(body->transform, point));
}

cpVect
cpBodyGetVelocityAtLocalPoint(const cpBody *body, cpVect point)
{
	cpBodySyncForRead(body);
	cpVect r = cpTransformVect(body->transform, cpvsub(point, body->cog));
	return cpvadd(body->v, cpvmult(cpvperp(r), body->w));
}

cpVect
cpBodyGetVelocityAtWorldPoint(const cpBody *body, cpVect point)
{
	cpBodySyncForRead(body);
	cpVect r = cpvsub(point, cpTransformPoint(body->transform, body->cog));
	return cpvadd(body->v, cpvmult(cpvperp(r), body->w));
}

cpFloat
cpBodyKineticEnergy(const cpBody *body)
{
	cpBodySyncForRead(body);
	cpFloat vsq = cpvdot(body->v, body->v);
	cpFloat wsq = body->w*body->w;
	return (vsq ? vsq*body->m : 0.0) + (wsq ? wsq*body->i : 0.0);
}

void
cpBodyEachShape(cpBody *body, cpBodyShapeIteratorFunc func, void *data)
{
	cpShape *shape = body->shapeList;
	while(shape){
		cpShape *next = shape->next;
		func(body, shape, data);
		shape = next;
	}
}

void
cpBodyEachConstraint(cpBody *body, cpBodyConstraintIteratorFunc func, void *data)
{
	cpConstraint *constraint = body->constraintList;
	while(constraint){
		cpConstraint *next = cpConstraintNext(constraint, body);
		func(body, constraint, data);
		constraint = next;
	}
}

/* Contact graph iteration (cpBody.c:612-626): the arbiters are fetched from the device on demand
 * and threaded per body by cpSpaceFetchArbitersB200. */
void
cpBodyEachArbiter(cpBody *body, cpBodyArbiterIteratorFunc func, void *data)
{
	cpSpace *space = body->space;
	if(!space) return;
	if(space->arbStale) cpSpaceFetchArbitersB200(space);
	int i = body->firstArb;
	while(i >= 0){
		cpArbiter *arb = &space->arbs[i];
		int next = (arb->body_a == body ? arb->next_a : arb->next_b);
		cpBool swapped = arb->swapped;
		arb->swapped = (body == arb->body_b);
		func(body, arb, data);
		arb->swapped = swapped;
		i = next;
	}
}

/* ---- sleeping API (cpSpaceComponent.c:113-153, 309-349) ---- */
void
cpBodyActivate(cpBody *body)
{
	if(body == NULL || cpBodyGetType(body) != CP_BODY_TYPE_DYNAMIC) return;
	cpSpace *space = body->space;
	if(space == NULL){ body->idleTime = 0.0; return; }
	cpBodySyncForRead(body);
	if(body->idleTime != 0.0){
		/* an awake body only needs its idle timer restarted on the device: a 4-byte index, not a re-upload */
		body->idleTime = 0.0;
		if(body->sleepRoot) space->bodiesDirty = cpTrue; else { body->idleReset = cpTrue; space->touchDirty = cpTrue; }
	}
	cpBody *root = body->sleepRoot;
	if(root){
		/* wake the whole sleeping component */
		for(int i = 0; i < space->nBodies; i++){
			cpBody *other = space->bodies[i];
			if(other->sleepRoot == root){ other->sleepRoot = NULL; other->idleTime = 0.0; }
		}
		space->bodiesDirty = cpTrue;
	}
}

void
cpBodyActivateStatic(cpBody *body, cpShape *filter)
{
	cpAssertHard(cpBodyGetType(body) == CP_BODY_TYPE_STATIC, "cpBodyActivateStatic() called on a non-static body.");
	cpSpace *space = body->space;
	if(!space) return;
	if(space->arbStale) cpSpaceFetchArbitersB200(space);
	for(int i = body->firstArb; i >= 0;){
		cpArbiter *arb = &space->arbs[i];
		if(!filter || filter == arb->a || filter == arb->b) cpBodyActivate(arb->body_a == body ? arb->body_b : arb->body_a);
		i = (arb->body_a == body ? arb->next_a : arb->next_b);
	}
}

void cpBodySleep(cpBody *body){ cpBodySleepWithGroup(body, NULL); }

void
cpBodySleepWithGroup(cpBody *body, cpBody *group)
{
	cpAssertHard(cpBodyGetType(body) == CP_BODY_TYPE_DYNAMIC, "Non-dynamic bodies cannot be put to sleep.");
	cpSpace *space = body->space;
	cpAssertHard(space, "Cannot put a body to sleep that has not been added to a space.");
	cpAssertHard(!cpSpaceIsLocked(space), "Bodies cannot be put to sleep during a query or a call to cpSpaceStep(). Put these calls into a post-step callback.");
	cpAssertHard(cpSpaceGetSleepTimeThreshold(space) < INFINITY, "Sleeping is not enabled on the space. You cannot sleep a body without setting a sleep time threshold on the space.");
	cpAssertHard(group == NULL || cpBodyIsSleeping(group), "Cannot use a non-sleeping body as a group identifier.");
	cpBodySyncForRead(body);
	if(body->sleepRoot){
		cpAssertHard(body->sleepRoot == (group ? group->sleepRoot : body->sleepRoot), "The body is already sleeping and it's group cannot be reassigned.");
		return;
	}
	body->sleepRoot = (group ? group->sleepRoot : body);
	body->idleTime = 0.0;
	space->bodiesDirty = cpTrue;
	/* shapes of a sleeping body keep their last cached AABB; make sure it is current */
	space->topologyDirty = cpTrue;
}
