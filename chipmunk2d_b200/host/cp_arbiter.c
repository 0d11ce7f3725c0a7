/* cp_arbiter.c -- cpArbiter accessors (public API of reference cpArbiter.h:32-143, implemented in
 * src/cpArbiter.c:54-247) over the arbiter records downloaded from the device.
 *
 * A cpArbiter* handed out by cpBodyEachArbiter or a collision handler is valid until the next
 * cpSpaceStep, like the reference's (cpSpaceStep.c:187-202).  Setters that change the outcome of
 * the solve (cpArbiterSet*, cpArbiterIgnore) only make sense inside begin/preSolve callbacks; those run
 * between the two halves of a split device step (cp_space.c: run_begin_presolve_callbacks), where the
 * changes made to the mirror are diffed and sent to the device before the solver runs.
 */
#include "cp_host.h"

cpFloat cpArbiterGetRestitution(const cpArbiter *arb){ return arb->e; }
void cpArbiterSetRestitution(cpArbiter *arb, cpFloat restitution){ arb->e = restitution; }
cpFloat cpArbiterGetFriction(const cpArbiter *arb){ return arb->u; }
void cpArbiterSetFriction(cpArbiter *arb, cpFloat friction){ arb->u = friction; }

cpVect
cpArbiterGetSurfaceVelocity(cpArbiter *arb)
{
	return cpvmult(arb->surface_vr, arb->swapped ? -1.0 : 1.0);
}

void
cpArbiterSetSurfaceVelocity(cpArbiter *arb, cpVect vr)
{
	arb->surface_vr = cpvmult(vr, arb->swapped ? -1.0 : 1.0);
}

cpDataPointer cpArbiterGetUserData(const cpArbiter *arb){ return arb->data; }
void cpArbiterSetUserData(cpArbiter *arb, cpDataPointer userData){ arb->data = userData; }

int
cpArbiterGetCount(const cpArbiter *arb)
{
	/* zero while cached / invalidated, i.e. inside separate() (cpArbiter.c:64-68) */
	return (arb->state < CP_ARBITER_STATE_CACHED ? arb->count : 0);
}

cpVect
cpArbiterTotalImpulse(const cpArbiter *arb)
{
	cpVect n = arb->n, sum = cpvzero;
	for(int i = 0, count = cpArbiterGetCount(arb); i < count; i++){
		const struct cpContact *con = &arb->contacts[i];
		sum = cpvadd(sum, cpvrotate(n, cpv(con->jnAcc, con->jtAcc)));
	}
	return (arb->swapped ? sum : cpvneg(sum));
}

cpFloat
cpArbiterTotalKE(const cpArbiter *arb)
{
	cpFloat eCoef = (1.0 - arb->e)/(1.0 + arb->e);
	cpFloat sum = 0.0;
	for(int i = 0, count = cpArbiterGetCount(arb); i < count; i++){
		const struct cpContact *con = &arb->contacts[i];
		cpFloat jnAcc = con->jnAcc, jtAcc = con->jtAcc;
		sum += eCoef*jnAcc*jnAcc/con->nMass + jtAcc*jtAcc/con->tMass;
	}
	return sum;
}

cpBool
cpArbiterIgnore(cpArbiter *arb)
{
	arb->state = CP_ARBITER_STATE_IGNORE;
	return cpFalse;
}

void
cpArbiterGetShapes(const cpArbiter *arb, cpShape **a, cpShape **b)
{
	if(arb->swapped){ (*a) = arb->b; (*b) = arb->a; } else { (*a) = arb->a; (*b) = arb->b; }
}

void
cpArbiterGetBodies(const cpArbiter *arb, cpBody **a, cpBody **b)
{
	cpShape *shape_a, *shape_b;
	cpArbiterGetShapes(arb, &shape_a, &shape_b);
	(*a) = shape_a->body;
	(*b) = shape_b->body;
}

cpBool cpArbiterIsFirstContact(const cpArbiter *arb){ return arb->state == CP_ARBITER_STATE_FIRST_COLLISION; }
cpBool cpArbiterIsRemoval(const cpArbiter *arb){ return arb->state == CP_ARBITER_STATE_INVALIDATED; }

cpVect
cpArbiterGetNormal(const cpArbiter *arb)
{
	return cpvmult(arb->n, arb->swapped ? -1.0 : 1.0);
}

static cpVect body_p(const cpBody *body){ cpBodySyncForRead(body); return body->p; }

cpVect
cpArbiterGetPointA(const cpArbiter *arb, int i)
{
	cpAssertHard(0 <= i && i < cpArbiterGetCount(arb), "Index error: The specified contact index is invalid for this arbiter");
	return cpvadd(body_p(arb->body_a), arb->contacts[i].r1);
}

cpVect
cpArbiterGetPointB(const cpArbiter *arb, int i)
{
	cpAssertHard(0 <= i && i < cpArbiterGetCount(arb), "Index error: The specified contact index is invalid for this arbiter");
	return cpvadd(body_p(arb->body_b), arb->contacts[i].r2);
}

cpFloat
cpArbiterGetDepth(const cpArbiter *arb, int i)
{
	cpAssertHard(0 <= i && i < cpArbiterGetCount(arb), "Index error: The specified contact index is invalid for this arbiter");
	const struct cpContact *con = &arb->contacts[i];
	return cpvdot(cpvadd(cpvsub(con->r2, con->r1), cpvsub(body_p(arb->body_b), body_p(arb->body_a))), arb->n);
}

cpContactPointSet
cpArbiterGetContactPointSet(const cpArbiter *arb)
{
	cpContactPointSet set;
	set.count = cpArbiterGetCount(arb);
	cpBool swapped = arb->swapped;
	cpVect n = arb->n;
	set.normal = (swapped ? cpvneg(n) : n);
	cpVect pa = body_p(arb->body_a), pb = body_p(arb->body_b);
	for(int i = 0; i < set.count; i++){
		cpVect p1 = cpvadd(pa, arb->contacts[i].r1);
		cpVect p2 = cpvadd(pb, arb->contacts[i].r2);
		set.points[i].pointA = (swapped ? p2 : p1);
		set.points[i].pointB = (swapped ? p1 : p2);
		set.points[i].distance = cpvdot(cpvsub(p2, p1), n);
	}
	return set;
}

void
cpArbiterSetContactPointSet(cpArbiter *arb, cpContactPointSet *set)
{
	int count = set->count;
	cpAssertHard(count == arb->count, "The number of contact points cannot be changed.");
	cpBool swapped = arb->swapped;
	arb->n = (swapped ? cpvneg(set->normal) : set->normal);
	cpVect pa = body_p(arb->body_a), pb = body_p(arb->body_b);
	for(int i = 0; i < count; i++){
		cpVect p1 = set->points[i].pointA, p2 = set->points[i].pointB;
		arb->contacts[i].r1 = cpvsub(swapped ? p2 : p1, pa);
		arb->contacts[i].r2 = cpvsub(swapped ? p1 : p2, pb);
	}
}

/* wildcard dispatch (cpArbiter.c:249-313): handlers registered through cpSpaceAddWildcardHandler are
 * looked up by the space when it runs the callbacks; the per-arbiter trampolines keep their signatures */
extern cpCollisionHandler *cpSpaceLookupWildcardB200(cpSpace *space, cpCollisionType type);

static cpCollisionHandler *wild(cpArbiter *arb, cpSpace *space, int second)
{
	cpShape *a, *b;
	cpArbiterGetShapes(arb, &a, &b);
	return cpSpaceLookupWildcardB200(space, second ? b->type : a->type);
}

cpBool cpArbiterCallWildcardBeginA(cpArbiter *arb, cpSpace *space){ cpCollisionHandler *h = wild(arb, space, 0); return h ? h->beginFunc(arb, space, h->userData) : cpTrue; }
cpBool
cpArbiterCallWildcardBeginB(cpArbiter *arb, cpSpace *space)
{
	cpCollisionHandler *h = wild(arb, space, 1);
	if(!h) return cpTrue;
	arb->swapped = !arb->swapped;
	cpBool r = h->beginFunc(arb, space, h->userData);
	arb->swapped = !arb->swapped;
	return r;
}
cpBool cpArbiterCallWildcardPreSolveA(cpArbiter *arb, cpSpace *space){ cpCollisionHandler *h = wild(arb, space, 0); return h ? h->preSolveFunc(arb, space, h->userData) : cpTrue; }
cpBool
cpArbiterCallWildcardPreSolveB(cpArbiter *arb, cpSpace *space)
{
	cpCollisionHandler *h = wild(arb, space, 1);
	if(!h) return cpTrue;
	arb->swapped = !arb->swapped;
	cpBool r = h->preSolveFunc(arb, space, h->userData);
	arb->swapped = !arb->swapped;
	return r;
}
void cpArbiterCallWildcardPostSolveA(cpArbiter *arb, cpSpace *space){ cpCollisionHandler *h = wild(arb, space, 0); if(h) h->postSolveFunc(arb, space, h->userData); }
void
cpArbiterCallWildcardPostSolveB(cpArbiter *arb, cpSpace *space)
{
	cpCollisionHandler *h = wild(arb, space, 1);
	if(!h) return;
	arb->swapped = !arb->swapped;
	h->postSolveFunc(arb, space, h->userData);
	arb->swapped = !arb->swapped;
}
void cpArbiterCallWildcardSeparateA(cpArbiter *arb, cpSpace *space){ cpCollisionHandler *h = wild(arb, space, 0); if(h) h->separateFunc(arb, space, h->userData); }
void
cpArbiterCallWildcardSeparateB(cpArbiter *arb, cpSpace *space)
{
	cpCollisionHandler *h = wild(arb, space, 1);
	if(!h) return;
	arb->swapped = !arb->swapped;
	h->separateFunc(arb, space, h->userData);
	arb->swapped = !arb->swapped;
}
