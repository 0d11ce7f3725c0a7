/* chipmunk/chipmunk.h -- public C API of the B200 drop-in for Chipmunk2D 7.0.3.
 *
 * One consolidated header written for this build.  It declares, with identical names,
 * argument order and by-value C ABI (cpFloat = double, cpVect / cpBB / cpTransform passed by
 * value), the part of the reference's public API that sits on or next to the cpSpaceStep hot
 * path (reference headers under include/chipmunk; SURVEY.md 8b "minimum set"):
 *   spaces, bodies, circle/segment/poly shapes, the ten joint classes, arbiters, collision
 *   handlers, post-step callbacks, iterators, cpHastySpace (chipmunk/cpHastySpace.h), space and
 *   shape queries (point / segment / bb / shape, as device scans), custom body integrators and
 *   spring force functions (host callbacks through a split step), the moment/area/hull helpers
 *   and the inline cpVect / cpBB / cpTransform math.
 * Not provided (out of the hot-path scope): spatial-index classes, debug draw,
 * autogeometry (cpMarch/cpPolyline), struct layouts (chipmunk_structs.h).
 *
 * A cpSpace created here lives on a B200: cpSpaceStep() runs every stage of the step as CUDA
 * kernels through the C ABI in include/cpb200.h.  Object state read through the getters below
 * is fetched from the device lazily.
 */
#ifndef CHIPMUNK_B200_H
#define CHIPMUNK_B200_H

#include <stdlib.h>
#include <stdint.h>
#include <math.h>
#include <float.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CP_EXPORT __attribute__((visibility("default")))
#else
#define CP_EXPORT
#endif

/* ---- assertions (reference chipmunk.h:41-58) ---- */
CP_EXPORT void cpMessage(const char *condition, const char *file, int line, int isError, int isHardError, const char *message, ...);
#ifdef NDEBUG
#define cpAssertWarn(cond, ...)
#define cpAssertSoft(cond, ...)
#else
#define cpAssertSoft(cond, ...) if(!(cond)){cpMessage(#cond, __FILE__, __LINE__, 1, 0, __VA_ARGS__); abort();}
#define cpAssertWarn(cond, ...) if(!(cond)) cpMessage(#cond, __FILE__, __LINE__, 0, 0, __VA_ARGS__)
#endif
#define cpAssertHard(cond, ...) if(!(cond)){cpMessage(#cond, __FILE__, __LINE__, 1, 1, __VA_ARGS__); abort();}

#ifndef cpcalloc
#define cpcalloc calloc
#endif
#ifndef cprealloc
#define cprealloc realloc
#endif
#ifndef cpfree
#define cpfree free
#endif

/* ---- scalar types (reference chipmunk_types.h) ---- */
typedef double cpFloat;
#define CP_USE_DOUBLES 1
#define cpfsqrt sqrt
#define cpfsin sin
#define cpfcos cos
#define cpfacos acos
#define cpfatan2 atan2
#define cpfmod fmod
#define cpfexp exp
#define cpfpow pow
#define cpffloor floor
#define cpfceil ceil
#define CPFLOAT_MIN DBL_MIN
#ifndef INFINITY
#define INFINITY (HUGE_VAL)
#endif
#define CP_PI ((cpFloat)3.14159265358979323846264338327950288)

typedef uintptr_t cpHashValue;
typedef uint32_t cpCollisionID;
typedef unsigned char cpBool;
#define cpTrue 1
#define cpFalse 0
typedef void *cpDataPointer;
typedef uintptr_t cpCollisionType;
typedef uintptr_t cpGroup;
typedef unsigned int cpBitmask;
typedef unsigned int cpTimestamp;
#define CP_NO_GROUP ((cpGroup)0)
#define CP_ALL_CATEGORIES (~(cpBitmask)0)
#define CP_WILDCARD_COLLISION_TYPE (~(cpCollisionType)0)

static inline cpFloat cpfmax(cpFloat a, cpFloat b){ return (a > b) ? a : b; }
static inline cpFloat cpfmin(cpFloat a, cpFloat b){ return (a < b) ? a : b; }
static inline cpFloat cpfabs(cpFloat f){ return (f < 0) ? -f : f; }
static inline cpFloat cpfclamp(cpFloat f, cpFloat lo, cpFloat hi){ return cpfmin(cpfmax(f, lo), hi); }
static inline cpFloat cpfclamp01(cpFloat f){ return cpfmax(0.0, cpfmin(f, 1.0)); }
static inline cpFloat cpflerp(cpFloat f1, cpFloat f2, cpFloat t){ return f1*(1.0 - t) + f2*t; }
static inline cpFloat cpflerpconst(cpFloat f1, cpFloat f2, cpFloat d){ return f1 + cpfclamp(f2 - f1, -d, d); }

typedef struct cpVect { cpFloat x, y; } cpVect;
typedef struct cpTransform { cpFloat a, b, c, d, tx, ty; } cpTransform;
typedef struct cpMat2x2 { cpFloat a, b, c, d; } cpMat2x2;
typedef struct cpBB { cpFloat l, b, r, t; } cpBB;

/* ---- opaque objects ---- */
typedef struct cpBody cpBody;
typedef struct cpShape cpShape;
typedef struct cpCircleShape cpCircleShape;
typedef struct cpSegmentShape cpSegmentShape;
typedef struct cpPolyShape cpPolyShape;
typedef struct cpConstraint cpConstraint;
typedef struct cpPinJoint cpPinJoint;
typedef struct cpSlideJoint cpSlideJoint;
typedef struct cpPivotJoint cpPivotJoint;
typedef struct cpGrooveJoint cpGrooveJoint;
typedef struct cpDampedSpring cpDampedSpring;
typedef struct cpDampedRotarySpring cpDampedRotarySpring;
typedef struct cpRotaryLimitJoint cpRotaryLimitJoint;
typedef struct cpRatchetJoint cpRatchetJoint;
typedef struct cpGearJoint cpGearJoint;
typedef struct cpSimpleMotor cpSimpleMotor;
typedef struct cpCollisionHandler cpCollisionHandler;
typedef struct cpContactPointSet cpContactPointSet;
typedef struct cpArbiter cpArbiter;
typedef struct cpSpace cpSpace;

/* ---- cpVect (reference cpVect.h) ---- */
static const cpVect cpvzero = {0.0, 0.0};
static inline cpVect cpv(const cpFloat x, const cpFloat y){ cpVect v = {x, y}; return v; }
static inline cpBool cpveql(const cpVect a, const cpVect b){ return (a.x == b.x && a.y == b.y); }
static inline cpVect cpvadd(const cpVect a, const cpVect b){ return cpv(a.x + b.x, a.y + b.y); }
static inline cpVect cpvsub(const cpVect a, const cpVect b){ return cpv(a.x - b.x, a.y - b.y); }
static inline cpVect cpvneg(const cpVect v){ return cpv(-v.x, -v.y); }
static inline cpVect cpvmult(const cpVect v, const cpFloat s){ return cpv(v.x*s, v.y*s); }
static inline cpFloat cpvdot(const cpVect a, const cpVect b){ return a.x*b.x + a.y*b.y; }
static inline cpFloat cpvcross(const cpVect a, const cpVect b){ return a.x*b.y - a.y*b.x; }
static inline cpVect cpvperp(const cpVect v){ return cpv(-v.y, v.x); }
static inline cpVect cpvrperp(const cpVect v){ return cpv(v.y, -v.x); }
static inline cpVect cpvproject(const cpVect a, const cpVect b){ return cpvmult(b, cpvdot(a, b)/cpvdot(b, b)); }
static inline cpVect cpvforangle(const cpFloat a){ return cpv(cpfcos(a), cpfsin(a)); }
static inline cpFloat cpvtoangle(const cpVect v){ return cpfatan2(v.y, v.x); }
static inline cpVect cpvrotate(const cpVect a, const cpVect b){ return cpv(a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x); }
static inline cpVect cpvunrotate(const cpVect a, const cpVect b){ return cpv(a.x*b.x + a.y*b.y, a.y*b.x - a.x*b.y); }
static inline cpFloat cpvlengthsq(const cpVect v){ return cpvdot(v, v); }
static inline cpFloat cpvlength(const cpVect v){ return cpfsqrt(cpvdot(v, v)); }
static inline cpVect cpvlerp(const cpVect a, const cpVect b, const cpFloat t){ return cpvadd(cpvmult(a, 1.0 - t), cpvmult(b, t)); }
static inline cpVect cpvnormalize(const cpVect v){ return cpvmult(v, 1.0/(cpvlength(v) + CPFLOAT_MIN)); }
static inline cpVect cpvslerp(const cpVect a, const cpVect b, const cpFloat t)
{
	cpFloat dot = cpvdot(cpvnormalize(a), cpvnormalize(b));
	cpFloat omega = cpfacos(cpfclamp(dot, -1.0, 1.0));
	if(omega < 1e-3) return cpvlerp(a, b, t);
	cpFloat denom = 1.0/cpfsin(omega);
	return cpvadd(cpvmult(a, cpfsin((1.0 - t)*omega)*denom), cpvmult(b, cpfsin(t*omega)*denom));
}
static inline cpVect cpvslerpconst(const cpVect a, const cpVect b, const cpFloat ang)
{
	cpFloat dot = cpvdot(cpvnormalize(a), cpvnormalize(b));
	cpFloat omega = cpfacos(cpfclamp(dot, -1.0, 1.0));
	return cpvslerp(a, b, cpfmin(ang, omega)/omega);
}
static inline cpVect cpvclamp(const cpVect v, const cpFloat len){ return (cpvdot(v, v) > len*len) ? cpvmult(cpvnormalize(v), len) : v; }
static inline cpVect cpvlerpconst(cpVect a, cpVect b, cpFloat d){ return cpvadd(a, cpvclamp(cpvsub(b, a), d)); }
static inline cpFloat cpvdist(const cpVect a, const cpVect b){ return cpvlength(cpvsub(a, b)); }
static inline cpFloat cpvdistsq(const cpVect a, const cpVect b){ return cpvlengthsq(cpvsub(a, b)); }
static inline cpBool cpvnear(const cpVect a, const cpVect b, const cpFloat dist){ return cpvdistsq(a, b) < dist*dist; }

static inline cpMat2x2 cpMat2x2New(cpFloat a, cpFloat b, cpFloat c, cpFloat d){ cpMat2x2 m = {a, b, c, d}; return m; }
static inline cpVect cpMat2x2Transform(cpMat2x2 m, cpVect v){ return cpv(v.x*m.a + v.y*m.b, v.x*m.c + v.y*m.d); }

/* ---- cpBB (reference cpBB.h) ---- */
static inline cpBB cpBBNew(const cpFloat l, const cpFloat b, const cpFloat r, const cpFloat t){ cpBB bb = {l, b, r, t}; return bb; }
static inline cpBB cpBBNewForExtents(const cpVect c, const cpFloat hw, const cpFloat hh){ return cpBBNew(c.x - hw, c.y - hh, c.x + hw, c.y + hh); }
static inline cpBB cpBBNewForCircle(const cpVect p, const cpFloat r){ return cpBBNewForExtents(p, r, r); }
static inline cpBool cpBBIntersects(const cpBB a, const cpBB b){ return (a.l <= b.r && b.l <= a.r && a.b <= b.t && b.b <= a.t); }
static inline cpBool cpBBContainsBB(const cpBB bb, const cpBB other){ return (bb.l <= other.l && bb.r >= other.r && bb.b <= other.b && bb.t >= other.t); }
static inline cpBool cpBBContainsVect(const cpBB bb, const cpVect v){ return (bb.l <= v.x && bb.r >= v.x && bb.b <= v.y && bb.t >= v.y); }
static inline cpBB cpBBMerge(const cpBB a, const cpBB b){ return cpBBNew(cpfmin(a.l, b.l), cpfmin(a.b, b.b), cpfmax(a.r, b.r), cpfmax(a.t, b.t)); }
static inline cpBB cpBBExpand(const cpBB bb, const cpVect v){ return cpBBNew(cpfmin(bb.l, v.x), cpfmin(bb.b, v.y), cpfmax(bb.r, v.x), cpfmax(bb.t, v.y)); }
static inline cpVect cpBBCenter(cpBB bb){ return cpvlerp(cpv(bb.l, bb.b), cpv(bb.r, bb.t), 0.5); }
static inline cpFloat cpBBArea(cpBB bb){ return (bb.r - bb.l)*(bb.t - bb.b); }
static inline cpFloat cpBBMergedArea(cpBB a, cpBB b){ return (cpfmax(a.r, b.r) - cpfmin(a.l, b.l))*(cpfmax(a.t, b.t) - cpfmin(a.b, b.b)); }
static inline cpFloat cpBBSegmentQuery(cpBB bb, cpVect a, cpVect b)
{
	cpVect delta = cpvsub(b, a);
	cpFloat tmin = -INFINITY, tmax = INFINITY;
	if(delta.x == 0.0){
		if(a.x < bb.l || bb.r < a.x) return INFINITY;
	} else {
		cpFloat t1 = (bb.l - a.x)/delta.x, t2 = (bb.r - a.x)/delta.x;
		tmin = cpfmax(tmin, cpfmin(t1, t2));
		tmax = cpfmin(tmax, cpfmax(t1, t2));
	}
	if(delta.y == 0.0){
		if(a.y < bb.b || bb.t < a.y) return INFINITY;
	} else {
		cpFloat t1 = (bb.b - a.y)/delta.y, t2 = (bb.t - a.y)/delta.y;
		tmin = cpfmax(tmin, cpfmin(t1, t2));
		tmax = cpfmin(tmax, cpfmax(t1, t2));
	}
	if(tmin <= tmax && 0.0 <= tmax && tmin <= 1.0) return cpfmax(tmin, 0.0);
	return INFINITY;
}
static inline cpBool cpBBIntersectsSegment(cpBB bb, cpVect a, cpVect b){ return (cpBBSegmentQuery(bb, a, b) != INFINITY); }
static inline cpVect cpBBClampVect(const cpBB bb, const cpVect v){ return cpv(cpfclamp(v.x, bb.l, bb.r), cpfclamp(v.y, bb.b, bb.t)); }
static inline cpVect cpBBWrapVect(const cpBB bb, const cpVect v)
{
	cpFloat dx = cpfabs(bb.r - bb.l), modx = cpfmod(v.x - bb.l, dx), x = (modx > 0.0) ? modx : modx + dx;
	cpFloat dy = cpfabs(bb.t - bb.b), mody = cpfmod(v.y - bb.b, dy), y = (mody > 0.0) ? mody : mody + dy;
	return cpv(x + bb.l, y + bb.b);
}
static inline cpBB cpBBOffset(const cpBB bb, const cpVect v){ return cpBBNew(bb.l + v.x, bb.b + v.y, bb.r + v.x, bb.t + v.y); }

/* ---- cpTransform (reference cpTransform.h) ---- */
static const cpTransform cpTransformIdentity = {1.0, 0.0, 0.0, 1.0, 0.0, 0.0};
static inline cpTransform cpTransformNew(cpFloat a, cpFloat b, cpFloat c, cpFloat d, cpFloat tx, cpFloat ty){ cpTransform t = {a, b, c, d, tx, ty}; return t; }
static inline cpTransform cpTransformNewTranspose(cpFloat a, cpFloat c, cpFloat tx, cpFloat b, cpFloat d, cpFloat ty){ cpTransform t = {a, b, c, d, tx, ty}; return t; }
static inline cpTransform cpTransformInverse(cpTransform t)
{
	cpFloat inv_det = 1.0/(t.a*t.d - t.c*t.b);
	return cpTransformNewTranspose(t.d*inv_det, -t.c*inv_det, (t.c*t.ty - t.tx*t.d)*inv_det, -t.b*inv_det, t.a*inv_det, (t.tx*t.b - t.a*t.ty)*inv_det);
}
static inline cpTransform cpTransformMult(cpTransform t1, cpTransform t2)
{
	return cpTransformNewTranspose(
		t1.a*t2.a + t1.c*t2.b, t1.a*t2.c + t1.c*t2.d, t1.a*t2.tx + t1.c*t2.ty + t1.tx,
		t1.b*t2.a + t1.d*t2.b, t1.b*t2.c + t1.d*t2.d, t1.b*t2.tx + t1.d*t2.ty + t1.ty);
}
static inline cpVect cpTransformPoint(cpTransform t, cpVect p){ return cpv(t.a*p.x + t.c*p.y + t.tx, t.b*p.x + t.d*p.y + t.ty); }
static inline cpVect cpTransformVect(cpTransform t, cpVect v){ return cpv(t.a*v.x + t.c*v.y, t.b*v.x + t.d*v.y); }
static inline cpBB cpTransformbBB(cpTransform t, cpBB bb)
{
	cpVect center = cpBBCenter(bb);
	cpFloat hw = (bb.r - bb.l)*0.5, hh = (bb.t - bb.b)*0.5;
	cpFloat a = t.a*hw, b = t.c*hh, d = t.b*hw, e = t.d*hh;
	cpFloat hw_max = cpfmax(cpfabs(a + b), cpfabs(a - b));
	cpFloat hh_max = cpfmax(cpfabs(d + e), cpfabs(d - e));
	return cpBBNewForExtents(cpTransformPoint(t, center), hw_max, hh_max);
}
static inline cpTransform cpTransformTranslate(cpVect translate){ return cpTransformNewTranspose(1.0, 0.0, translate.x, 0.0, 1.0, translate.y); }
static inline cpTransform cpTransformScale(cpFloat sx, cpFloat sy){ return cpTransformNewTranspose(sx, 0.0, 0.0, 0.0, sy, 0.0); }
static inline cpTransform cpTransformRotate(cpFloat radians){ cpVect rot = cpvforangle(radians); return cpTransformNewTranspose(rot.x, -rot.y, 0.0, rot.y, rot.x, 0.0); }
static inline cpTransform cpTransformRigid(cpVect translate, cpFloat radians){ cpVect rot = cpvforangle(radians); return cpTransformNewTranspose(rot.x, -rot.y, translate.x, rot.y, rot.x, translate.y); }
static inline cpTransform cpTransformRigidInverse(cpTransform t){ return cpTransformNewTranspose(t.d, -t.c, (t.c*t.ty - t.tx*t.d), -t.b, t.a, (t.tx*t.b - t.a*t.ty)); }
static inline cpTransform cpTransformWrap(cpTransform outer, cpTransform inner){ return cpTransformMult(cpTransformInverse(outer), cpTransformMult(inner, outer)); }
static inline cpTransform cpTransformWrapInverse(cpTransform outer, cpTransform inner){ return cpTransformMult(outer, cpTransformMult(inner, cpTransformInverse(outer))); }
static inline cpTransform cpTransformOrtho(cpBB bb)
{
	return cpTransformNewTranspose(2.0/(bb.r - bb.l), 0.0, -(bb.r + bb.l)/(bb.r - bb.l), 0.0, 2.0/(bb.t - bb.b), -(bb.t + bb.b)/(bb.t - bb.b));
}
static inline cpTransform cpTransformBoneScale(cpVect v0, cpVect v1){ cpVect d = cpvsub(v1, v0); return cpTransformNewTranspose(d.x, -d.y, v0.x, d.y, d.x, v0.y); }
static inline cpTransform cpTransformAxialScale(cpVect axis, cpVect pivot, cpFloat scale)
{
	cpFloat A = axis.x*axis.y*(scale - 1.0);
	cpFloat B = cpvdot(axis, pivot)*(1.0 - scale);
	return cpTransformNewTranspose(scale*axis.x*axis.x + axis.y*axis.y, A, axis.x*B, A, axis.x*axis.x + scale*axis.y*axis.y, axis.y*B);
}

/* ---- version / helpers (reference chipmunk.h:128-215, chipmunk.c) ---- */
#define CP_VERSION_MAJOR 7
#define CP_VERSION_MINOR 0
#define CP_VERSION_RELEASE 3
CP_EXPORT extern const char *cpVersionString;
CP_EXPORT cpFloat cpMomentForCircle(cpFloat m, cpFloat r1, cpFloat r2, cpVect offset);
CP_EXPORT cpFloat cpAreaForCircle(cpFloat r1, cpFloat r2);
CP_EXPORT cpFloat cpMomentForSegment(cpFloat m, cpVect a, cpVect b, cpFloat radius);
CP_EXPORT cpFloat cpAreaForSegment(cpVect a, cpVect b, cpFloat radius);
CP_EXPORT cpFloat cpMomentForPoly(cpFloat m, int count, const cpVect *verts, cpVect offset, cpFloat radius);
CP_EXPORT cpFloat cpAreaForPoly(const int count, const cpVect *verts, cpFloat radius);
CP_EXPORT cpVect cpCentroidForPoly(const int count, const cpVect *verts);
CP_EXPORT cpFloat cpMomentForBox(cpFloat m, cpFloat width, cpFloat height);
CP_EXPORT cpFloat cpMomentForBox2(cpFloat m, cpBB box);
CP_EXPORT int cpConvexHull(int count, const cpVect *verts, cpVect *result, int *first, cpFloat tol);
#define CP_CONVEX_HULL(__count__, __verts__, __count_var__, __verts_var__) \
cpVect *__verts_var__ = (cpVect *)alloca(__count__*sizeof(cpVect)); \
int __count_var__ = cpConvexHull(__count__, __verts__, __verts_var__, NULL, 0.0);

/* ---- bodies (reference cpBody.h) ---- */
typedef enum cpBodyType { CP_BODY_TYPE_DYNAMIC, CP_BODY_TYPE_KINEMATIC, CP_BODY_TYPE_STATIC } cpBodyType;
typedef void (*cpBodyVelocityFunc)(cpBody *body, cpVect gravity, cpFloat damping, cpFloat dt);
typedef void (*cpBodyPositionFunc)(cpBody *body, cpFloat dt);
typedef void (*cpBodyShapeIteratorFunc)(cpBody *body, cpShape *shape, void *data);
typedef void (*cpBodyConstraintIteratorFunc)(cpBody *body, cpConstraint *constraint, void *data);
typedef void (*cpBodyArbiterIteratorFunc)(cpBody *body, cpArbiter *arbiter, void *data);

CP_EXPORT cpBody *cpBodyAlloc(void);
CP_EXPORT cpBody *cpBodyInit(cpBody *body, cpFloat mass, cpFloat moment);
CP_EXPORT cpBody *cpBodyNew(cpFloat mass, cpFloat moment);
CP_EXPORT cpBody *cpBodyNewKinematic(void);
CP_EXPORT cpBody *cpBodyNewStatic(void);
CP_EXPORT void cpBodyDestroy(cpBody *body);
CP_EXPORT void cpBodyFree(cpBody *body);
CP_EXPORT void cpBodyActivate(cpBody *body);
CP_EXPORT void cpBodyActivateStatic(cpBody *body, cpShape *filter);
CP_EXPORT void cpBodySleep(cpBody *body);
CP_EXPORT void cpBodySleepWithGroup(cpBody *body, cpBody *group);
CP_EXPORT cpBool cpBodyIsSleeping(const cpBody *body);
CP_EXPORT cpBodyType cpBodyGetType(cpBody *body);
CP_EXPORT void cpBodySetType(cpBody *body, cpBodyType type);
CP_EXPORT cpSpace *cpBodyGetSpace(const cpBody *body);
CP_EXPORT cpFloat cpBodyGetMass(const cpBody *body);
CP_EXPORT void cpBodySetMass(cpBody *body, cpFloat m);
CP_EXPORT cpFloat cpBodyGetMoment(const cpBody *body);
CP_EXPORT void cpBodySetMoment(cpBody *body, cpFloat i);
CP_EXPORT cpVect cpBodyGetPosition(const cpBody *body);
CP_EXPORT void cpBodySetPosition(cpBody *body, cpVect pos);
CP_EXPORT cpVect cpBodyGetCenterOfGravity(const cpBody *body);
CP_EXPORT void cpBodySetCenterOfGravity(cpBody *body, cpVect cog);
CP_EXPORT cpVect cpBodyGetVelocity(const cpBody *body);
CP_EXPORT void cpBodySetVelocity(cpBody *body, cpVect velocity);
CP_EXPORT cpVect cpBodyGetForce(const cpBody *body);
CP_EXPORT void cpBodySetForce(cpBody *body, cpVect force);
CP_EXPORT cpFloat cpBodyGetAngle(const cpBody *body);
CP_EXPORT void cpBodySetAngle(cpBody *body, cpFloat a);
CP_EXPORT cpFloat cpBodyGetAngularVelocity(const cpBody *body);
CP_EXPORT void cpBodySetAngularVelocity(cpBody *body, cpFloat angularVelocity);
CP_EXPORT cpFloat cpBodyGetTorque(const cpBody *body);
CP_EXPORT void cpBodySetTorque(cpBody *body, cpFloat torque);
CP_EXPORT cpVect cpBodyGetRotation(const cpBody *body);
CP_EXPORT cpDataPointer cpBodyGetUserData(const cpBody *body);
CP_EXPORT void cpBodySetUserData(cpBody *body, cpDataPointer userData);
CP_EXPORT void cpBodySetVelocityUpdateFunc(cpBody *body, cpBodyVelocityFunc velocityFunc);
CP_EXPORT void cpBodySetPositionUpdateFunc(cpBody *body, cpBodyPositionFunc positionFunc);
CP_EXPORT void cpBodyUpdateVelocity(cpBody *body, cpVect gravity, cpFloat damping, cpFloat dt);
CP_EXPORT void cpBodyUpdatePosition(cpBody *body, cpFloat dt);
CP_EXPORT cpVect cpBodyLocalToWorld(const cpBody *body, const cpVect point);
CP_EXPORT cpVect cpBodyWorldToLocal(const cpBody *body, const cpVect point);
CP_EXPORT void cpBodyApplyForceAtWorldPoint(cpBody *body, cpVect force, cpVect point);
CP_EXPORT void cpBodyApplyForceAtLocalPoint(cpBody *body, cpVect force, cpVect point);
CP_EXPORT void cpBodyApplyImpulseAtWorldPoint(cpBody *body, cpVect impulse, cpVect point);
CP_EXPORT void cpBodyApplyImpulseAtLocalPoint(cpBody *body, cpVect impulse, cpVect point);
CP_EXPORT cpVect cpBodyGetVelocityAtWorldPoint(const cpBody *body, cpVect point);
CP_EXPORT cpVect cpBodyGetVelocityAtLocalPoint(const cpBody *body, cpVect point);
CP_EXPORT cpFloat cpBodyKineticEnergy(const cpBody *body);
CP_EXPORT void cpBodyEachShape(cpBody *body, cpBodyShapeIteratorFunc func, void *data);
CP_EXPORT void cpBodyEachConstraint(cpBody *body, cpBodyConstraintIteratorFunc func, void *data);
CP_EXPORT void cpBodyEachArbiter(cpBody *body, cpBodyArbiterIteratorFunc func, void *data);

/* ---- shapes (reference cpShape.h, cpPolyShape.h, chipmunk_unsafe.h) ---- */
typedef struct cpShapeFilter { cpGroup group; cpBitmask categories; cpBitmask mask; } cpShapeFilter;
static const cpShapeFilter CP_SHAPE_FILTER_ALL = {CP_NO_GROUP, CP_ALL_CATEGORIES, CP_ALL_CATEGORIES};
static const cpShapeFilter CP_SHAPE_FILTER_NONE = {CP_NO_GROUP, ~CP_ALL_CATEGORIES, ~CP_ALL_CATEGORIES};
static inline cpShapeFilter cpShapeFilterNew(cpGroup group, cpBitmask categories, cpBitmask mask){ cpShapeFilter f = {group, categories, mask}; return f; }

#define CP_MAX_CONTACTS_PER_ARBITER 2
struct cpContactPointSet {
	int count;
	cpVect normal;
	struct { cpVect pointA, pointB; cpFloat distance; } points[CP_MAX_CONTACTS_PER_ARBITER];
};

CP_EXPORT void cpShapeDestroy(cpShape *shape);
CP_EXPORT void cpShapeFree(cpShape *shape);
CP_EXPORT cpBB cpShapeCacheBB(cpShape *shape);
CP_EXPORT cpBB cpShapeUpdate(cpShape *shape, cpTransform transform);
CP_EXPORT cpContactPointSet cpShapesCollide(const cpShape *a, const cpShape *b);
/* nearest-point / segment queries on one shape (reference cpShape.h:27-49, 88-97) */
typedef struct cpPointQueryInfo { const cpShape *shape; cpVect point; cpFloat distance; cpVect gradient; } cpPointQueryInfo;
typedef struct cpSegmentQueryInfo { const cpShape *shape; cpVect point; cpVect normal; cpFloat alpha; } cpSegmentQueryInfo;
CP_EXPORT cpFloat cpShapePointQuery(const cpShape *shape, cpVect p, cpPointQueryInfo *out);
CP_EXPORT cpBool cpShapeSegmentQuery(const cpShape *shape, cpVect a, cpVect b, cpFloat radius, cpSegmentQueryInfo *info);
CP_EXPORT cpSpace *cpShapeGetSpace(const cpShape *shape);
CP_EXPORT cpBody *cpShapeGetBody(const cpShape *shape);
CP_EXPORT void cpShapeSetBody(cpShape *shape, cpBody *body);
CP_EXPORT cpFloat cpShapeGetMass(cpShape *shape);
CP_EXPORT void cpShapeSetMass(cpShape *shape, cpFloat mass);
CP_EXPORT cpFloat cpShapeGetDensity(cpShape *shape);
CP_EXPORT void cpShapeSetDensity(cpShape *shape, cpFloat density);
CP_EXPORT cpFloat cpShapeGetMoment(cpShape *shape);
CP_EXPORT cpFloat cpShapeGetArea(cpShape *shape);
CP_EXPORT cpVect cpShapeGetCenterOfGravity(cpShape *shape);
CP_EXPORT cpBB cpShapeGetBB(const cpShape *shape);
CP_EXPORT cpBool cpShapeGetSensor(const cpShape *shape);
CP_EXPORT void cpShapeSetSensor(cpShape *shape, cpBool sensor);
CP_EXPORT cpFloat cpShapeGetElasticity(const cpShape *shape);
CP_EXPORT void cpShapeSetElasticity(cpShape *shape, cpFloat elasticity);
CP_EXPORT cpFloat cpShapeGetFriction(const cpShape *shape);
CP_EXPORT void cpShapeSetFriction(cpShape *shape, cpFloat friction);
CP_EXPORT cpVect cpShapeGetSurfaceVelocity(const cpShape *shape);
CP_EXPORT void cpShapeSetSurfaceVelocity(cpShape *shape, cpVect surfaceVelocity);
CP_EXPORT cpDataPointer cpShapeGetUserData(const cpShape *shape);
CP_EXPORT void cpShapeSetUserData(cpShape *shape, cpDataPointer userData);
CP_EXPORT cpCollisionType cpShapeGetCollisionType(const cpShape *shape);
CP_EXPORT void cpShapeSetCollisionType(cpShape *shape, cpCollisionType collisionType);
CP_EXPORT cpShapeFilter cpShapeGetFilter(const cpShape *shape);
CP_EXPORT void cpShapeSetFilter(cpShape *shape, cpShapeFilter filter);

CP_EXPORT cpCircleShape *cpCircleShapeAlloc(void);
CP_EXPORT cpCircleShape *cpCircleShapeInit(cpCircleShape *circle, cpBody *body, cpFloat radius, cpVect offset);
CP_EXPORT cpShape *cpCircleShapeNew(cpBody *body, cpFloat radius, cpVect offset);
CP_EXPORT cpVect cpCircleShapeGetOffset(const cpShape *shape);
CP_EXPORT cpFloat cpCircleShapeGetRadius(const cpShape *shape);
CP_EXPORT cpSegmentShape *cpSegmentShapeAlloc(void);
CP_EXPORT cpSegmentShape *cpSegmentShapeInit(cpSegmentShape *seg, cpBody *body, cpVect a, cpVect b, cpFloat radius);
CP_EXPORT cpShape *cpSegmentShapeNew(cpBody *body, cpVect a, cpVect b, cpFloat radius);
CP_EXPORT void cpSegmentShapeSetNeighbors(cpShape *shape, cpVect prev, cpVect next);
CP_EXPORT cpVect cpSegmentShapeGetA(const cpShape *shape);
CP_EXPORT cpVect cpSegmentShapeGetB(const cpShape *shape);
CP_EXPORT cpVect cpSegmentShapeGetNormal(const cpShape *shape);
CP_EXPORT cpFloat cpSegmentShapeGetRadius(const cpShape *shape);
CP_EXPORT cpPolyShape *cpPolyShapeAlloc(void);
CP_EXPORT cpPolyShape *cpPolyShapeInit(cpPolyShape *poly, cpBody *body, int count, const cpVect *verts, cpTransform transform, cpFloat radius);
CP_EXPORT cpPolyShape *cpPolyShapeInitRaw(cpPolyShape *poly, cpBody *body, int count, const cpVect *verts, cpFloat radius);
CP_EXPORT cpShape *cpPolyShapeNew(cpBody *body, int count, const cpVect *verts, cpTransform transform, cpFloat radius);
CP_EXPORT cpShape *cpPolyShapeNewRaw(cpBody *body, int count, const cpVect *verts, cpFloat radius);
CP_EXPORT cpPolyShape *cpBoxShapeInit(cpPolyShape *poly, cpBody *body, cpFloat width, cpFloat height, cpFloat radius);
CP_EXPORT cpPolyShape *cpBoxShapeInit2(cpPolyShape *poly, cpBody *body, cpBB box, cpFloat radius);
CP_EXPORT cpShape *cpBoxShapeNew(cpBody *body, cpFloat width, cpFloat height, cpFloat radius);
CP_EXPORT cpShape *cpBoxShapeNew2(cpBody *body, cpBB box, cpFloat radius);
CP_EXPORT int cpPolyShapeGetCount(const cpShape *shape);
CP_EXPORT cpVect cpPolyShapeGetVert(const cpShape *shape, int index);
CP_EXPORT cpFloat cpPolyShapeGetRadius(const cpShape *shape);
/* chipmunk_unsafe.h */
CP_EXPORT void cpCircleShapeSetRadius(cpShape *shape, cpFloat radius);
CP_EXPORT void cpCircleShapeSetOffset(cpShape *shape, cpVect offset);
CP_EXPORT void cpSegmentShapeSetEndpoints(cpShape *shape, cpVect a, cpVect b);
CP_EXPORT void cpSegmentShapeSetRadius(cpShape *shape, cpFloat radius);
CP_EXPORT void cpPolyShapeSetVerts(cpShape *shape, int count, cpVect *verts, cpTransform transform);
CP_EXPORT void cpPolyShapeSetVertsRaw(cpShape *shape, int count, cpVect *verts);
CP_EXPORT void cpPolyShapeSetRadius(cpShape *shape, cpFloat radius);

/* ---- constraints (reference cpConstraint.h and the joint headers) ---- */
typedef void (*cpConstraintPreSolveFunc)(cpConstraint *constraint, cpSpace *space);
typedef void (*cpConstraintPostSolveFunc)(cpConstraint *constraint, cpSpace *space);
typedef cpFloat (*cpDampedSpringForceFunc)(cpConstraint *spring, cpFloat dist);
typedef cpFloat (*cpDampedRotarySpringTorqueFunc)(struct cpConstraint *spring, cpFloat relativeAngle);

CP_EXPORT void cpConstraintDestroy(cpConstraint *constraint);
CP_EXPORT void cpConstraintFree(cpConstraint *constraint);
CP_EXPORT cpSpace *cpConstraintGetSpace(const cpConstraint *constraint);
CP_EXPORT cpBody *cpConstraintGetBodyA(const cpConstraint *constraint);
CP_EXPORT cpBody *cpConstraintGetBodyB(const cpConstraint *constraint);
CP_EXPORT cpFloat cpConstraintGetMaxForce(const cpConstraint *constraint);
CP_EXPORT void cpConstraintSetMaxForce(cpConstraint *constraint, cpFloat maxForce);
CP_EXPORT cpFloat cpConstraintGetErrorBias(const cpConstraint *constraint);
CP_EXPORT void cpConstraintSetErrorBias(cpConstraint *constraint, cpFloat errorBias);
CP_EXPORT cpFloat cpConstraintGetMaxBias(const cpConstraint *constraint);
CP_EXPORT void cpConstraintSetMaxBias(cpConstraint *constraint, cpFloat maxBias);
CP_EXPORT cpBool cpConstraintGetCollideBodies(const cpConstraint *constraint);
CP_EXPORT void cpConstraintSetCollideBodies(cpConstraint *constraint, cpBool collideBodies);
CP_EXPORT cpConstraintPreSolveFunc cpConstraintGetPreSolveFunc(const cpConstraint *constraint);
CP_EXPORT void cpConstraintSetPreSolveFunc(cpConstraint *constraint, cpConstraintPreSolveFunc preSolveFunc);
CP_EXPORT cpConstraintPostSolveFunc cpConstraintGetPostSolveFunc(const cpConstraint *constraint);
CP_EXPORT void cpConstraintSetPostSolveFunc(cpConstraint *constraint, cpConstraintPostSolveFunc postSolveFunc);
CP_EXPORT cpDataPointer cpConstraintGetUserData(const cpConstraint *constraint);
CP_EXPORT void cpConstraintSetUserData(cpConstraint *constraint, cpDataPointer userData);
CP_EXPORT cpFloat cpConstraintGetImpulse(cpConstraint *constraint);

#define CP_JOINT_COMMON(Type) \
	CP_EXPORT cpBool cpConstraintIs##Type(const cpConstraint *constraint); \
	CP_EXPORT cp##Type *cp##Type##Alloc(void);
#define CP_JOINT_PROP(Type, ctype, Name) \
	CP_EXPORT ctype cp##Type##Get##Name(const cpConstraint *constraint); \
	CP_EXPORT void cp##Type##Set##Name(cpConstraint *constraint, ctype value);

CP_JOINT_COMMON(PinJoint)
CP_EXPORT cpPinJoint *cpPinJointInit(cpPinJoint *joint, cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB);
CP_EXPORT cpConstraint *cpPinJointNew(cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB);
CP_JOINT_PROP(PinJoint, cpVect, AnchorA) CP_JOINT_PROP(PinJoint, cpVect, AnchorB) CP_JOINT_PROP(PinJoint, cpFloat, Dist)

CP_JOINT_COMMON(SlideJoint)
CP_EXPORT cpSlideJoint *cpSlideJointInit(cpSlideJoint *joint, cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB, cpFloat min, cpFloat max);
CP_EXPORT cpConstraint *cpSlideJointNew(cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB, cpFloat min, cpFloat max);
CP_JOINT_PROP(SlideJoint, cpVect, AnchorA) CP_JOINT_PROP(SlideJoint, cpVect, AnchorB) CP_JOINT_PROP(SlideJoint, cpFloat, Min) CP_JOINT_PROP(SlideJoint, cpFloat, Max)

CP_JOINT_COMMON(PivotJoint)
CP_EXPORT cpPivotJoint *cpPivotJointInit(cpPivotJoint *joint, cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB);
CP_EXPORT cpConstraint *cpPivotJointNew(cpBody *a, cpBody *b, cpVect pivot);
CP_EXPORT cpConstraint *cpPivotJointNew2(cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB);
CP_JOINT_PROP(PivotJoint, cpVect, AnchorA) CP_JOINT_PROP(PivotJoint, cpVect, AnchorB)

CP_JOINT_COMMON(GrooveJoint)
CP_EXPORT cpGrooveJoint *cpGrooveJointInit(cpGrooveJoint *joint, cpBody *a, cpBody *b, cpVect groove_a, cpVect groove_b, cpVect anchorB);
CP_EXPORT cpConstraint *cpGrooveJointNew(cpBody *a, cpBody *b, cpVect groove_a, cpVect groove_b, cpVect anchorB);
CP_JOINT_PROP(GrooveJoint, cpVect, GrooveA) CP_JOINT_PROP(GrooveJoint, cpVect, GrooveB) CP_JOINT_PROP(GrooveJoint, cpVect, AnchorB)

CP_JOINT_COMMON(DampedSpring)
CP_EXPORT cpDampedSpring *cpDampedSpringInit(cpDampedSpring *joint, cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB, cpFloat restLength, cpFloat stiffness, cpFloat damping);
CP_EXPORT cpConstraint *cpDampedSpringNew(cpBody *a, cpBody *b, cpVect anchorA, cpVect anchorB, cpFloat restLength, cpFloat stiffness, cpFloat damping);
CP_JOINT_PROP(DampedSpring, cpVect, AnchorA) CP_JOINT_PROP(DampedSpring, cpVect, AnchorB) CP_JOINT_PROP(DampedSpring, cpFloat, RestLength)
CP_JOINT_PROP(DampedSpring, cpFloat, Stiffness) CP_JOINT_PROP(DampedSpring, cpFloat, Damping) CP_JOINT_PROP(DampedSpring, cpDampedSpringForceFunc, SpringForceFunc)

CP_JOINT_COMMON(DampedRotarySpring)
CP_EXPORT cpDampedRotarySpring *cpDampedRotarySpringInit(cpDampedRotarySpring *joint, cpBody *a, cpBody *b, cpFloat restAngle, cpFloat stiffness, cpFloat damping);
CP_EXPORT cpConstraint *cpDampedRotarySpringNew(cpBody *a, cpBody *b, cpFloat restAngle, cpFloat stiffness, cpFloat damping);
CP_JOINT_PROP(DampedRotarySpring, cpFloat, RestAngle) CP_JOINT_PROP(DampedRotarySpring, cpFloat, Stiffness) CP_JOINT_PROP(DampedRotarySpring, cpFloat, Damping)
CP_JOINT_PROP(DampedRotarySpring, cpDampedRotarySpringTorqueFunc, SpringTorqueFunc)

CP_JOINT_COMMON(RotaryLimitJoint)
CP_EXPORT cpRotaryLimitJoint *cpRotaryLimitJointInit(cpRotaryLimitJoint *joint, cpBody *a, cpBody *b, cpFloat min, cpFloat max);
CP_EXPORT cpConstraint *cpRotaryLimitJointNew(cpBody *a, cpBody *b, cpFloat min, cpFloat max);
CP_JOINT_PROP(RotaryLimitJoint, cpFloat, Min) CP_JOINT_PROP(RotaryLimitJoint, cpFloat, Max)

CP_JOINT_COMMON(RatchetJoint)
CP_EXPORT cpRatchetJoint *cpRatchetJointInit(cpRatchetJoint *joint, cpBody *a, cpBody *b, cpFloat phase, cpFloat ratchet);
CP_EXPORT cpConstraint *cpRatchetJointNew(cpBody *a, cpBody *b, cpFloat phase, cpFloat ratchet);
CP_JOINT_PROP(RatchetJoint, cpFloat, Angle) CP_JOINT_PROP(RatchetJoint, cpFloat, Phase) CP_JOINT_PROP(RatchetJoint, cpFloat, Ratchet)

CP_JOINT_COMMON(GearJoint)
CP_EXPORT cpGearJoint *cpGearJointInit(cpGearJoint *joint, cpBody *a, cpBody *b, cpFloat phase, cpFloat ratio);
CP_EXPORT cpConstraint *cpGearJointNew(cpBody *a, cpBody *b, cpFloat phase, cpFloat ratio);
CP_JOINT_PROP(GearJoint, cpFloat, Phase) CP_JOINT_PROP(GearJoint, cpFloat, Ratio)

CP_JOINT_COMMON(SimpleMotor)
CP_EXPORT cpSimpleMotor *cpSimpleMotorInit(cpSimpleMotor *joint, cpBody *a, cpBody *b, cpFloat rate);
CP_EXPORT cpConstraint *cpSimpleMotorNew(cpBody *a, cpBody *b, cpFloat rate);
CP_JOINT_PROP(SimpleMotor, cpFloat, Rate)

/* ---- arbiters (reference cpArbiter.h) ---- */
CP_EXPORT cpFloat cpArbiterGetRestitution(const cpArbiter *arb);
CP_EXPORT void cpArbiterSetRestitution(cpArbiter *arb, cpFloat restitution);
CP_EXPORT cpFloat cpArbiterGetFriction(const cpArbiter *arb);
CP_EXPORT void cpArbiterSetFriction(cpArbiter *arb, cpFloat friction);
CP_EXPORT cpVect cpArbiterGetSurfaceVelocity(cpArbiter *arb);
CP_EXPORT void cpArbiterSetSurfaceVelocity(cpArbiter *arb, cpVect vr);
CP_EXPORT cpDataPointer cpArbiterGetUserData(const cpArbiter *arb);
CP_EXPORT void cpArbiterSetUserData(cpArbiter *arb, cpDataPointer userData);
CP_EXPORT cpVect cpArbiterTotalImpulse(const cpArbiter *arb);
CP_EXPORT cpFloat cpArbiterTotalKE(const cpArbiter *arb);
CP_EXPORT cpBool cpArbiterIgnore(cpArbiter *arb);
CP_EXPORT void cpArbiterGetShapes(const cpArbiter *arb, cpShape **a, cpShape **b);
#define CP_ARBITER_GET_SHAPES(__arb__, __a__, __b__) cpShape *__a__, *__b__; cpArbiterGetShapes(__arb__, &__a__, &__b__);
CP_EXPORT void cpArbiterGetBodies(const cpArbiter *arb, cpBody **a, cpBody **b);
#define CP_ARBITER_GET_BODIES(__arb__, __a__, __b__) cpBody *__a__, *__b__; cpArbiterGetBodies(__arb__, &__a__, &__b__);
CP_EXPORT cpContactPointSet cpArbiterGetContactPointSet(const cpArbiter *arb);
CP_EXPORT void cpArbiterSetContactPointSet(cpArbiter *arb, cpContactPointSet *set);
CP_EXPORT cpBool cpArbiterIsFirstContact(const cpArbiter *arb);
CP_EXPORT cpBool cpArbiterIsRemoval(const cpArbiter *arb);
CP_EXPORT int cpArbiterGetCount(const cpArbiter *arb);
CP_EXPORT cpVect cpArbiterGetNormal(const cpArbiter *arb);
CP_EXPORT cpVect cpArbiterGetPointA(const cpArbiter *arb, int i);
CP_EXPORT cpVect cpArbiterGetPointB(const cpArbiter *arb, int i);
CP_EXPORT cpFloat cpArbiterGetDepth(const cpArbiter *arb, int i);
CP_EXPORT cpBool cpArbiterCallWildcardBeginA(cpArbiter *arb, cpSpace *space);
CP_EXPORT cpBool cpArbiterCallWildcardBeginB(cpArbiter *arb, cpSpace *space);
CP_EXPORT cpBool cpArbiterCallWildcardPreSolveA(cpArbiter *arb, cpSpace *space);
CP_EXPORT cpBool cpArbiterCallWildcardPreSolveB(cpArbiter *arb, cpSpace *space);
CP_EXPORT void cpArbiterCallWildcardPostSolveA(cpArbiter *arb, cpSpace *space);
CP_EXPORT void cpArbiterCallWildcardPostSolveB(cpArbiter *arb, cpSpace *space);
CP_EXPORT void cpArbiterCallWildcardSeparateA(cpArbiter *arb, cpSpace *space);
CP_EXPORT void cpArbiterCallWildcardSeparateB(cpArbiter *arb, cpSpace *space);

/* ---- spaces (reference cpSpace.h) ---- */
typedef cpBool (*cpCollisionBeginFunc)(cpArbiter *arb, cpSpace *space, cpDataPointer userData);
typedef cpBool (*cpCollisionPreSolveFunc)(cpArbiter *arb, cpSpace *space, cpDataPointer userData);
typedef void (*cpCollisionPostSolveFunc)(cpArbiter *arb, cpSpace *space, cpDataPointer userData);
typedef void (*cpCollisionSeparateFunc)(cpArbiter *arb, cpSpace *space, cpDataPointer userData);
struct cpCollisionHandler {
	const cpCollisionType typeA;
	const cpCollisionType typeB;
	cpCollisionBeginFunc beginFunc;
	cpCollisionPreSolveFunc preSolveFunc;
	cpCollisionPostSolveFunc postSolveFunc;
	cpCollisionSeparateFunc separateFunc;
	cpDataPointer userData;
};
typedef void (*cpPostStepFunc)(cpSpace *space, void *key, void *data);
typedef void (*cpSpaceBodyIteratorFunc)(cpBody *body, void *data);
typedef void (*cpSpaceShapeIteratorFunc)(cpShape *shape, void *data);
typedef void (*cpSpaceConstraintIteratorFunc)(cpConstraint *constraint, void *data);

CP_EXPORT cpSpace *cpSpaceAlloc(void);
CP_EXPORT cpSpace *cpSpaceInit(cpSpace *space);
CP_EXPORT cpSpace *cpSpaceNew(void);
CP_EXPORT void cpSpaceDestroy(cpSpace *space);
CP_EXPORT void cpSpaceFree(cpSpace *space);
#define CP_SPACE_PROP(ctype, Name) \
	CP_EXPORT ctype cpSpaceGet##Name(const cpSpace *space); \
	CP_EXPORT void cpSpaceSet##Name(cpSpace *space, ctype value);
CP_SPACE_PROP(int, Iterations) CP_SPACE_PROP(cpVect, Gravity) CP_SPACE_PROP(cpFloat, Damping)
CP_SPACE_PROP(cpFloat, IdleSpeedThreshold) CP_SPACE_PROP(cpFloat, SleepTimeThreshold) CP_SPACE_PROP(cpFloat, CollisionSlop)
CP_SPACE_PROP(cpFloat, CollisionBias) CP_SPACE_PROP(cpTimestamp, CollisionPersistence) CP_SPACE_PROP(cpDataPointer, UserData)
CP_EXPORT cpBody *cpSpaceGetStaticBody(const cpSpace *space);
CP_EXPORT cpFloat cpSpaceGetCurrentTimeStep(const cpSpace *space);
CP_EXPORT cpBool cpSpaceIsLocked(cpSpace *space);
CP_EXPORT cpCollisionHandler *cpSpaceAddDefaultCollisionHandler(cpSpace *space);
CP_EXPORT cpCollisionHandler *cpSpaceAddCollisionHandler(cpSpace *space, cpCollisionType a, cpCollisionType b);
CP_EXPORT cpCollisionHandler *cpSpaceAddWildcardHandler(cpSpace *space, cpCollisionType type);
CP_EXPORT cpShape *cpSpaceAddShape(cpSpace *space, cpShape *shape);
CP_EXPORT cpBody *cpSpaceAddBody(cpSpace *space, cpBody *body);
CP_EXPORT cpConstraint *cpSpaceAddConstraint(cpSpace *space, cpConstraint *constraint);
CP_EXPORT void cpSpaceRemoveShape(cpSpace *space, cpShape *shape);
CP_EXPORT void cpSpaceRemoveBody(cpSpace *space, cpBody *body);
CP_EXPORT void cpSpaceRemoveConstraint(cpSpace *space, cpConstraint *constraint);
CP_EXPORT cpBool cpSpaceContainsShape(cpSpace *space, cpShape *shape);
CP_EXPORT cpBool cpSpaceContainsBody(cpSpace *space, cpBody *body);
CP_EXPORT cpBool cpSpaceContainsConstraint(cpSpace *space, cpConstraint *constraint);
CP_EXPORT cpBool cpSpaceAddPostStepCallback(cpSpace *space, cpPostStepFunc func, void *key, void *data);
CP_EXPORT void cpSpaceEachBody(cpSpace *space, cpSpaceBodyIteratorFunc func, void *data);
CP_EXPORT void cpSpaceEachShape(cpSpace *space, cpSpaceShapeIteratorFunc func, void *data);
CP_EXPORT void cpSpaceEachConstraint(cpSpace *space, cpSpaceConstraintIteratorFunc func, void *data);
/* space queries (reference cpSpace.h:191-222, cpSpaceQuery.c): run as data-parallel scans on the device */
typedef void (*cpSpacePointQueryFunc)(cpShape *shape, cpVect point, cpFloat distance, cpVect gradient, void *data);
typedef void (*cpSpaceSegmentQueryFunc)(cpShape *shape, cpVect point, cpVect normal, cpFloat alpha, void *data);
typedef void (*cpSpaceBBQueryFunc)(cpShape *shape, void *data);
typedef void (*cpSpaceShapeQueryFunc)(cpShape *shape, cpContactPointSet *points, void *data);
CP_EXPORT void cpSpacePointQuery(cpSpace *space, cpVect point, cpFloat maxDistance, cpShapeFilter filter, cpSpacePointQueryFunc func, void *data);
CP_EXPORT cpShape *cpSpacePointQueryNearest(cpSpace *space, cpVect point, cpFloat maxDistance, cpShapeFilter filter, cpPointQueryInfo *out);
CP_EXPORT void cpSpaceSegmentQuery(cpSpace *space, cpVect start, cpVect end, cpFloat radius, cpShapeFilter filter, cpSpaceSegmentQueryFunc func, void *data);
CP_EXPORT cpShape *cpSpaceSegmentQueryFirst(cpSpace *space, cpVect start, cpVect end, cpFloat radius, cpShapeFilter filter, cpSegmentQueryInfo *out);
CP_EXPORT void cpSpaceBBQuery(cpSpace *space, cpBB bb, cpShapeFilter filter, cpSpaceBBQueryFunc func, void *data);
CP_EXPORT cpBool cpSpaceShapeQuery(cpSpace *space, cpShape *shape, cpSpaceShapeQueryFunc func, void *data);
CP_EXPORT void cpSpaceReindexStatic(cpSpace *space);
CP_EXPORT void cpSpaceReindexShape(cpSpace *space, cpShape *shape);
CP_EXPORT void cpSpaceReindexShapesForBody(cpSpace *space, cpBody *body);
CP_EXPORT void cpSpaceUseSpatialHash(cpSpace *space, cpFloat dim, int count);
CP_EXPORT void cpSpaceStep(cpSpace *space, cpFloat dt);

/* ---- B200 extensions (not in the reference) ---- */
/* CUDA device a space will be created on by the next cpSpaceNew / cpHastySpaceNew (default 0,
 * or the CPB200_DEVICE environment variable). */
CP_EXPORT void cpSpaceSetDefaultDeviceB200(int device);
/* Select the solver order: 0 graph-coloured parallel (default), 1 serial validation order. */
CP_EXPORT void cpSpaceSetSolverModeB200(cpSpace *space, int mode);
/* Step without leaving the device n times (no host round trip in between; callbacks and
 * post-step callbacks run once at the end). */
CP_EXPORT void cpSpaceStepManyB200(cpSpace *space, cpFloat dt, int n);
/* Force a download of the body state (normally lazy: the first getter after a step does it). */
CP_EXPORT void cpSpaceSyncB200(cpSpace *space);

#ifdef __cplusplus
}
#endif
#endif
