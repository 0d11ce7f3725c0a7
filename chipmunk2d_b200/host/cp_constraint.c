/* cp_constraint.c -- host mirror of the ten joint classes (public API of reference cpConstraint.h
 * and cp{Pin,Slide,Pivot,Groove}Joint.h, cpDamped{,Rotary}Spring.h, cpRotaryLimitJoint.h,
 * cpRatchetJoint.h, cpGearJoint.h, cpSimpleMotor.h).
 *
 * Only parameters live here; preStep / applyCachedImpulse / applyImpulse of every class run on the
 * device (csrc/k_joint.cuh).  All classes share struct cpConstraint: anchors + prm[4] use the packing
 * of cpb200_joint_desc.  Accessor families are generated with the macros below.
 */
#include "cp_host.h"

static void
joint_dirty(cpConstraint *c)
{
	cpBodyActivate(c->a);
	cpBodyActivate(c->b);
	if(c->space) cpSpaceMarkConstraintDirtyB200(c);
}

static cpConstraint *
joint_init(cpConstraint *c, int klass, cpBody *a, cpBody *b)
{
	/* cpConstraintInit (cpConstraint.c:38-57) */
	c->klass = klass;
	c->a = a;
	c->b = b;
	c->space = NULL;
	c->next_a = NULL;
	c->next_b = NULL;
	c->maxForce = (cpFloat)INFINITY;
	c->errorBias = cpfpow((cpFloat)(1.0f - 0.1f), 60.0);   /* float literals in the reference (cpConstraint.c:50) */
	c->maxBias = (cpFloat)INFINITY;
	c->collideBodies = cpTrue;
	c->preSolve = NULL;
	c->postSolve = NULL;
	c->userData = NULL;
	c->anchorA = cpvzero; c->anchorB = cpvzero;
	c->prm[0] = c->prm[1] = c->prm[2] = c->prm[3] = 0.0;
	c->acc = cpvzero;
	c->impulse = 0.0;
	c->forceFunc = NULL;
	c->index = -1;
	return c;
}

void cpConstraintDestroy(cpConstraint *constraint){ (void)constraint; }
void cpConstraintFree(cpConstraint *constraint){ if(constraint){ cpConstraintDestroy(constraint); cpfree(constraint); } }
cpSpace *cpConstraintGetSpace(const cpConstraint *constraint){ return constraint->space; }
cpBody *cpConstraintGetBodyA(const cpConstraint *constraint){ return constraint->a; }
cpBody *cpConstraintGetBodyB(const cpConstraint *constraint){ return constraint->b; }
cpFloat cpConstraintGetMaxForce(const cpConstraint *constraint){ return constraint->maxForce; }
void cpConstraintSetMaxForce(cpConstraint *constraint, cpFloat maxForce){ cpAssertHard(maxForce >= 0.0, "maxForce must be positive."); constraint->maxForce = maxForce; joint_dirty(constraint); }
cpFloat cpConstraintGetErrorBias(const cpConstraint *constraint){ return constraint->errorBias; }
void cpConstraintSetErrorBias(cpConstraint *constraint, cpFloat errorBias){ cpAssertHard(errorBias >= 0.0, "errorBias must be positive."); constraint->errorBias = errorBias; joint_dirty(constraint); }
cpFloat cpConstraintGetMaxBias(const cpConstraint *constraint){ return constraint->maxBias; }
void cpConstraintSetMaxBias(cpConstraint *constraint, cpFloat maxBias){ cpAssertHard(maxBias >= 0.0, "maxBias must be positive."); constraint->maxBias = maxBias; joint_dirty(constraint); }
cpBool cpConstraintGetCollideBodies(const cpConstraint *constraint){ return constraint->collideBodies; }
void cpConstraintSetCollideBodies(cpConstraint *constraint, cpBool collideBodies){ constraint->collideBodies = collideBodies; joint_dirty(constraint); }
cpConstraintPreSolveFunc cpConstraintGetPreSolveFunc(const cpConstraint *constraint){ return constraint->preSolve; }
void cpConstraintSetPreSolveFunc(cpConstraint *constraint, cpConstraintPreSolveFunc preSolveFunc){ constraint->preSolve = preSolveFunc; }
cpConstraintPostSolveFunc cpConstraintGetPostSolveFunc(const cpConstraint *constraint){ return constraint->postSolve; }
void cpConstraintSetPostSolveFunc(cpConstraint *constraint, cpConstraintPostSolveFunc postSolveFunc){ constraint->postSolve = postSolveFunc; }
cpDataPointer cpConstraintGetUserData(const cpConstraint *constraint){ return constraint->userData; }
void cpConstraintSetUserData(cpConstraint *constraint, cpDataPointer userData){ constraint->userData = userData; }

cpFloat
cpConstraintGetImpulse(cpConstraint *constraint)
{
	cpSpace *space = constraint->space;
	if(space && space->jointStale) cpSpaceFetchJointsB200(space);
	return constraint->impulse;
}

#define JOINT_CLASS(Type, KLASS) \
	cpBool cpConstraintIs##Type(const cpConstraint *constraint){ return constraint->klass == KLASS; } \
	cp##Type *cp##Type##Alloc(void){ return (cp##Type *)cpcalloc(1, sizeof(cp##Type)); }
#define JOINT_PROP(Type, KLASS, ctype, Name, field) \
	ctype cp##Type##Get##Name(const cpConstraint *constraint){ cpAssertHard(constraint->klass == KLASS, "Constraint is not a " #Type "."); return constraint->field; } \
	void cp##Type##Set##Name(cpConstraint *constraint, ctype value){ cpAssertHard(constraint->klass == KLASS, "Constraint is not a " #Type "."); constraint->field = value; joint_dirty(constraint); }

/* ---- pin (cpPinJoint.c:95-171): prm[0] = dist ---- */
JOINT_CLASS(PinJoint, CPB200_JOINT_PIN)
cpPinJoint *
cpPinJointInit(cpPinJoint *joint, cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB)
{
	cpConstraint *c = joint_init((cpConstraint *)joint, CPB200_JOINT_PIN, a, b);
	c->anchorA = anchorA;
	c->anchorB = anchorB;
	cpVect p1 = (a ? cpBodyLocalToWorld(a, anchorA) : anchorA);
	cpVect p2 = (b ? cpBodyLocalToWorld(b, anchorB) : anchorB);
	c->prm[0] = cpvlength(cpvsub(p2, p1));
	cpAssertWarn(c->prm[0] > 0.0, "You created a 0 length pin joint. A pivot joint will be much more stable.");
	return joint;
}
cpConstraint *cpPinJointNew(cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB){ return (cpConstraint *)cpPinJointInit(cpPinJointAlloc(), a, b, anchorA, anchorB); }
JOINT_PROP(PinJoint, CPB200_JOINT_PIN, cpVect, AnchorA, anchorA)
JOINT_PROP(PinJoint, CPB200_JOINT_PIN, cpVect, AnchorB, anchorB)
JOINT_PROP(PinJoint, CPB200_JOINT_PIN, cpFloat, Dist, prm[0])

/* ---- slide (cpSlideJoint.c:108-195): prm[0] = min, prm[1] = max ---- */
JOINT_CLASS(SlideJoint, CPB200_JOINT_SLIDE)
cpSlideJoint *
cpSlideJointInit(cpSlideJoint *joint, cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB, cpFloat min, cpFloat max)
{
	cpConstraint *c = joint_init((cpConstraint *)joint, CPB200_JOINT_SLIDE, a, b);
	c->anchorA = anchorA; c->anchorB = anchorB;
	c->prm[0] = min; c->prm[1] = max;
	return joint;
}
cpConstraint *cpSlideJointNew(cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB, cpFloat min, cpFloat max){ return (cpConstraint *)cpSlideJointInit(cpSlideJointAlloc(), a, b, anchorA, anchorB, min, max); }
JOINT_PROP(SlideJoint, CPB200_JOINT_SLIDE, cpVect, AnchorA, anchorA)
JOINT_PROP(SlideJoint, CPB200_JOINT_SLIDE, cpVect, AnchorB, anchorB)
JOINT_PROP(SlideJoint, CPB200_JOINT_SLIDE, cpFloat, Min, prm[0])
JOINT_PROP(SlideJoint, CPB200_JOINT_SLIDE, cpFloat, Max, prm[1])

/* ---- pivot (cpPivotJoint.c:88-152) ---- */
JOINT_CLASS(PivotJoint, CPB200_JOINT_PIVOT)
cpPivotJoint *
cpPivotJointInit(cpPivotJoint *joint, cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB)
{
	cpConstraint *c = joint_init((cpConstraint *)joint, CPB200_JOINT_PIVOT, a, b);
	c->anchorA = anchorA; c->anchorB = anchorB;
	return joint;
}
cpConstraint *cpPivotJointNew2(cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB){ return (cpConstraint *)cpPivotJointInit(cpPivotJointAlloc(), a, b, anchorA, anchorB); }
cpConstraint *
cpPivotJointNew(cpBody *a, cpBody *b, cpVect pivot)
{
	cpVect anchorA = (a ? cpBodyWorldToLocal(a, pivot) : pivot);
	cpVect anchorB = (b ? cpBodyWorldToLocal(b, pivot) : pivot);
	return cpPivotJointNew2(a, b, anchorA, anchorB);
}
JOINT_PROP(PivotJoint, CPB200_JOINT_PIVOT, cpVect, AnchorA, anchorA)
JOINT_PROP(PivotJoint, CPB200_JOINT_PIVOT, cpVect, AnchorB, anchorB)

/* ---- groove (cpGrooveJoint.c:118-197): anchorA = grv_a, prm[0..1] = grv_b ---- */
JOINT_CLASS(GrooveJoint, CPB200_JOINT_GROOVE)
cpGrooveJoint *
cpGrooveJointInit(cpGrooveJoint *joint, cpBody *a, cpBody *b, cpVect groove_a, cpVect groove_b, cpVect anchorB)
{
	cpConstraint *c = joint_init((cpConstraint *)joint, CPB200_JOINT_GROOVE, a, b);
	c->anchorA = groove_a;
	c->prm[0] = groove_b.x; c->prm[1] = groove_b.y;
	c->anchorB = anchorB;
	return joint;
}
cpConstraint *cpGrooveJointNew(cpBody *a, cpBody *b, cpVect groove_a, cpVect groove_b, cpVect anchorB){ return (cpConstraint *)cpGrooveJointInit(cpGrooveJointAlloc(), a, b, groove_a, groove_b, anchorB); }
JOINT_PROP(GrooveJoint, CPB200_JOINT_GROOVE, cpVect, GrooveA, anchorA)
JOINT_PROP(GrooveJoint, CPB200_JOINT_GROOVE, cpVect, AnchorB, anchorB)
cpVect cpGrooveJointGetGrooveB(const cpConstraint *constraint){ cpAssertHard(constraint->klass == CPB200_JOINT_GROOVE, "Constraint is not a groove joint."); return cpv(constraint->prm[0], constraint->prm[1]); }
void cpGrooveJointSetGrooveB(cpConstraint *constraint, cpVect value){ cpAssertHard(constraint->klass == CPB200_JOINT_GROOVE, "Constraint is not a groove joint."); constraint->prm[0] = value.x; constraint->prm[1] = value.y; joint_dirty(constraint); }

/* ---- damped spring (cpDampedSpring.c:96-216): prm = restLength, stiffness, damping ---- */
JOINT_CLASS(DampedSpring, CPB200_JOINT_DAMPED_SPRING)
cpDampedSpring *
cpDampedSpringInit(cpDampedSpring *spring, cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB, cpFloat restLength, cpFloat stiffness, cpFloat damping)
{
	cpConstraint *c = joint_init((cpConstraint *)spring, CPB200_JOINT_DAMPED_SPRING, a, b);
	c->anchorA = anchorA; c->anchorB = anchorB;
	c->prm[0] = restLength; c->prm[1] = stiffness; c->prm[2] = damping;
	return spring;
}
cpConstraint *cpDampedSpringNew(cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB, cpFloat restLength, cpFloat stiffness, cpFloat damping){ return (cpConstraint *)cpDampedSpringInit(cpDampedSpringAlloc(), a, b, anchorA, anchorB, restLength, stiffness, damping); }
JOINT_PROP(DampedSpring, CPB200_JOINT_DAMPED_SPRING, cpVect, AnchorA, anchorA)
JOINT_PROP(DampedSpring, CPB200_JOINT_DAMPED_SPRING, cpVect, AnchorB, anchorB)
JOINT_PROP(DampedSpring, CPB200_JOINT_DAMPED_SPRING, cpFloat, RestLength, prm[0])
JOINT_PROP(DampedSpring, CPB200_JOINT_DAMPED_SPRING, cpFloat, Stiffness, prm[1])
JOINT_PROP(DampedSpring, CPB200_JOINT_DAMPED_SPRING, cpFloat, Damping, prm[2])
cpDampedSpringForceFunc cpDampedSpringGetSpringForceFunc(const cpConstraint *constraint){ return (cpDampedSpringForceFunc)constraint->forceFunc; }
void
cpDampedSpringSetSpringForceFunc(cpConstraint *constraint, cpDampedSpringForceFunc springForceFunc)
{
	/* the device evaluates the default linear spring (cpDampedSpring.c:24-27); a user function is called by the host
	 * layer between the collision phase and the prestep (slow path of cpSpaceStep in cp_space.c; demo/Springies.c) */
	constraint->forceFunc = (void *)springForceFunc;
	if(springForceFunc && constraint->space) constraint->space->anyCustom = cpTrue;
}

/* ---- damped rotary spring (cpDampedRotarySpring.c:90-178): prm = restAngle, stiffness, damping ---- */
JOINT_CLASS(DampedRotarySpring, CPB200_JOINT_DAMPED_ROTARY_SPRING)
cpDampedRotarySpring *
cpDampedRotarySpringInit(cpDampedRotarySpring *spring, cpBody *a, cpBody *b, cpFloat restAngle, cpFloat stiffness, cpFloat damping)
{
	cpConstraint *c = joint_init((cpConstraint *)spring, CPB200_JOINT_DAMPED_ROTARY_SPRING, a, b);
	c->prm[0] = restAngle; c->prm[1] = stiffness; c->prm[2] = damping;
	return spring;
}
cpConstraint *cpDampedRotarySpringNew(cpBody *a, cpBody *b, cpFloat restAngle, cpFloat stiffness, cpFloat damping){ return (cpConstraint *)cpDampedRotarySpringInit(cpDampedRotarySpringAlloc(), a, b, restAngle, stiffness, damping); }
JOINT_PROP(DampedRotarySpring, CPB200_JOINT_DAMPED_ROTARY_SPRING, cpFloat, RestAngle, prm[0])
JOINT_PROP(DampedRotarySpring, CPB200_JOINT_DAMPED_ROTARY_SPRING, cpFloat, Stiffness, prm[1])
JOINT_PROP(DampedRotarySpring, CPB200_JOINT_DAMPED_ROTARY_SPRING, cpFloat, Damping, prm[2])
cpDampedRotarySpringTorqueFunc cpDampedRotarySpringGetSpringTorqueFunc(const cpConstraint *constraint){ return (cpDampedRotarySpringTorqueFunc)constraint->forceFunc; }
void
cpDampedRotarySpringSetSpringTorqueFunc(cpConstraint *constraint, cpDampedRotarySpringTorqueFunc springTorqueFunc)
{
	constraint->forceFunc = (void *)springTorqueFunc;
	if(springTorqueFunc && constraint->space) constraint->space->anyCustom = cpTrue;
}

/* ---- rotary limit (cpRotaryLimitJoint.c:104-160): prm = min, max ---- */
JOINT_CLASS(RotaryLimitJoint, CPB200_JOINT_ROTARY_LIMIT)
cpRotaryLimitJoint *
cpRotaryLimitJointInit(cpRotaryLimitJoint *joint, cpBody *a, cpBody *b, cpFloat min, cpFloat max)
{
	cpConstraint *c = joint_init((cpConstraint *)joint, CPB200_JOINT_ROTARY_LIMIT, a, b);
	c->prm[0] = min; c->prm[1] = max;
	return joint;
}
cpConstraint *cpRotaryLimitJointNew(cpBody *a, cpBody *b, cpFloat min, cpFloat max){ return (cpConstraint *)cpRotaryLimitJointInit(cpRotaryLimitJointAlloc(), a, b, min, max); }
JOINT_PROP(RotaryLimitJoint, CPB200_JOINT_ROTARY_LIMIT, cpFloat, Min, prm[0])
JOINT_PROP(RotaryLimitJoint, CPB200_JOINT_ROTARY_LIMIT, cpFloat, Max, prm[1])

/* ---- ratchet (cpRatchetJoint.c:107-179): prm = angle, phase, ratchet ---- */
JOINT_CLASS(RatchetJoint, CPB200_JOINT_RATCHET)
cpRatchetJoint *
cpRatchetJointInit(cpRatchetJoint *joint, cpBody *a, cpBody *b, cpFloat phase, cpFloat ratchet)
{
	cpConstraint *c = joint_init((cpConstraint *)joint, CPB200_JOINT_RATCHET, a, b);
	/* angle starts at the current relative angle (cpRatchetJoint.c:118-119) */
	c->prm[0] = (b ? cpBodyGetAngle(b) : 0.0) - (a ? cpBodyGetAngle(a) : 0.0);
	c->prm[1] = phase; c->prm[2] = ratchet;
	return joint;
}
cpConstraint *cpRatchetJointNew(cpBody *a, cpBody *b, cpFloat phase, cpFloat ratchet){ return (cpConstraint *)cpRatchetJointInit(cpRatchetJointAlloc(), a, b, phase, ratchet); }
cpFloat
cpRatchetJointGetAngle(const cpConstraint *constraint)
{
	cpAssertHard(constraint->klass == CPB200_JOINT_RATCHET, "Constraint is not a ratchet joint.");
	if(constraint->space && constraint->space->jointStale) cpSpaceFetchJointsB200(constraint->space);
	return constraint->prm[0];
}
void cpRatchetJointSetAngle(cpConstraint *constraint, cpFloat value){ cpAssertHard(constraint->klass == CPB200_JOINT_RATCHET, "Constraint is not a ratchet joint."); constraint->prm[0] = value; joint_dirty(constraint); }
JOINT_PROP(RatchetJoint, CPB200_JOINT_RATCHET, cpFloat, Phase, prm[1])
JOINT_PROP(RatchetJoint, CPB200_JOINT_RATCHET, cpFloat, Ratchet, prm[2])

/* ---- gear (cpGearJoint.c:86-145): prm = phase, ratio ---- */
JOINT_CLASS(GearJoint, CPB200_JOINT_GEAR)
cpGearJoint *
cpGearJointInit(cpGearJoint *joint, cpBody *a, cpBody *b, cpFloat phase, cpFloat ratio)
{
	cpConstraint *c = joint_init((cpConstraint *)joint, CPB200_JOINT_GEAR, a, b);
	c->prm[0] = phase; c->prm[1] = ratio;
	return joint;
}
cpConstraint *cpGearJointNew(cpBody *a, cpBody *b, cpFloat phase, cpFloat ratio){ return (cpConstraint *)cpGearJointInit(cpGearJointAlloc(), a, b, phase, ratio); }
JOINT_PROP(GearJoint, CPB200_JOINT_GEAR, cpFloat, Phase, prm[0])
JOINT_PROP(GearJoint, CPB200_JOINT_GEAR, cpFloat, Ratio, prm[1])

/* ---- simple motor (cpSimpleMotor.c:81-123): prm[0] = rate ---- */
JOINT_CLASS(SimpleMotor, CPB200_JOINT_SIMPLE_MOTOR)
cpSimpleMotor *
cpSimpleMotorInit(cpSimpleMotor *joint, cpBody *a, cpBody *b, cpFloat rate)
{
	cpConstraint *c = joint_init((cpConstraint *)joint, CPB200_JOINT_SIMPLE_MOTOR, a, b);
	c->prm[0] = rate;
	return joint;
}
cpConstraint *cpSimpleMotorNew(cpBody *a, cpBody *b, cpFloat rate){ return (cpConstraint *)cpSimpleMotorInit(cpSimpleMotorAlloc(), a, b, rate); }
JOINT_PROP(SimpleMotor, CPB200_JOINT_SIMPLE_MOTOR, cpFloat, Rate, prm[0])
