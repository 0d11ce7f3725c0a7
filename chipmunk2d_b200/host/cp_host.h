/* cp_host.h -- private object model of the C99 host layer.
 *
 * The host keeps plain C mirrors of the user's objects (what the reference keeps in
 * chipmunk_structs.h); the simulation state itself lives in SoA device buffers behind the
 * C ABI of include/cpb200.h.  Coherence protocol (SURVEY.md 8b "host-visible state"):
 *   - structural edits and parameter setters mark the space dirty; cpSpaceStep uploads before
 *     stepping (full re-upload for structure, body-state upload for kinematic setters);
 *   - after a step the mirrors are stale; the first getter downloads body state, shape AABBs,
 *     arbiters or joint impulses on demand.
 */
#ifndef CP_HOST_H
#define CP_HOST_H

#include "chipmunk/chipmunk.h"
#include "chipmunk/cpHastySpace.h"
#include "cpb200.h"

enum { CP_CIRCLE_SHAPE = 0, CP_SEGMENT_SHAPE = 1, CP_POLY_SHAPE = 2 };

/* numerically equal to CPB200_ARB_* and to the reference's enum cpArbiterState */
enum {
	CP_ARBITER_STATE_FIRST_COLLISION = 0, CP_ARBITER_STATE_NORMAL = 1, CP_ARBITER_STATE_IGNORE = 2,
	CP_ARBITER_STATE_CACHED = 3, CP_ARBITER_STATE_INVALIDATED = 4
};

struct cpBody {
	/* Field order is the access pattern of the per-body API loops a host runs every step (cpBodySetForce /
	 * cpBodyApplyForce* on every body, then cpBodyGetPosition / GetVelocity on every body): what those calls
	 * touch sits in the first two cache lines, so a million-body loop streams 128 B per body, not the whole struct. */
	cpSpace *space;
	cpFloat idleTime;          /* INFINITY <=> static (cpBody.c:136-146) */
	cpBody *sleepRoot;         /* non-NULL <=> asleep; all members of a sleeping component share it */
	cpFloat m;
	cpVect f;
	cpDataPointer userData;    /* (callbacks of cpSpaceEachBody usually start by asking for it) */
	cpBool idleReset;          /* activated while awake since the last upload: the device restarts its idle timer */
	unsigned fetchStamp;       /* == space->fetchStamp <=> this mirror holds the state of the last download */
	/* second line: what the getters read */
	cpTransform transform;
	cpVect p;
	/* the rest */
	cpFloat t;
	cpVect v;
	cpFloat a, w;
	cpFloat m_inv, i, i_inv;
	cpVect cog;
	cpVect v_bias;
	cpFloat w_bias;
	cpBodyVelocityFunc velocity_func;
	cpBodyPositionFunc position_func;
	cpShape *shapeList;
	cpConstraint *constraintList;
	int index;                 /* slot in space->bodies == device body index, -1 when not in a space */
	int firstArb;              /* head of this body's arbiter list in space->arbs, -1 = none */
};

struct cpShapeMassInfo { cpFloat m, i; cpVect cog; cpFloat area; };

struct cpShape {
	int klass;                 /* CP_*_SHAPE */
	cpSpace *space;
	cpBody *body;
	struct cpShapeMassInfo massInfo;
	cpBB bb;
	cpBool sensor;
	cpFloat e, u;
	cpVect surfaceV;
	cpDataPointer userData;
	cpCollisionType type;
	cpShapeFilter filter;
	cpShape *next, *prev;      /* body's shape list */
	cpHashValue hashid;
	int index;                 /* slot in space->shapes == device shape index */
};

struct cpCircleShape { cpShape shape; cpVect c, tc; cpFloat r; };
struct cpSegmentShape { cpShape shape; cpVect a, b, n, ta, tb, tn; cpFloat r; cpVect a_tangent, b_tangent; };
struct cpPolyShape { cpShape shape; cpFloat r; int count; cpVect *verts, *normals, *tverts, *tnormals; };

/* every joint class shares one record; cpPinJoint etc. are this struct under another tag */
struct cpConstraint {
	int klass;                 /* CPB200_JOINT_* */
	cpSpace *space;
	cpBody *a, *b;
	cpConstraint *next_a, *next_b;
	cpFloat maxForce, errorBias, maxBias;
	cpBool collideBodies;
	cpConstraintPreSolveFunc preSolve;
	cpConstraintPostSolveFunc postSolve;
	cpDataPointer userData;
	cpVect anchorA, anchorB;   /* groove: anchorA = grv_a */
	cpFloat prm[4];            /* packing of cpb200_joint_desc.prm */
	cpVect acc;                /* accumulated impulse mirror */
	cpFloat impulse;           /* getImpulse mirror */
	void *forceFunc;           /* spring force / torque function (default = NULL) */
	int index;
};
struct cpPinJoint { cpConstraint constraint; };
struct cpSlideJoint { cpConstraint constraint; };
struct cpPivotJoint { cpConstraint constraint; };
struct cpGrooveJoint { cpConstraint constraint; };
struct cpDampedSpring { cpConstraint constraint; };
struct cpDampedRotarySpring { cpConstraint constraint; };
struct cpRotaryLimitJoint { cpConstraint constraint; };
struct cpRatchetJoint { cpConstraint constraint; };
struct cpGearJoint { cpConstraint constraint; };
struct cpSimpleMotor { cpConstraint constraint; };

struct cpContact { cpVect r1, r2; cpFloat nMass, tMass, bounce, jnAcc, jtAcc, jBias, bias; cpHashValue hash; };

struct cpArbiter {
	cpSpace *space;
	cpFloat e, u;
	cpVect surface_vr;
	cpDataPointer data;
	cpShape *a, *b;
	cpBody *body_a, *body_b;
	int count;
	struct cpContact contacts[CP_MAX_CONTACTS_PER_ARBITER];
	cpVect n;
	cpCollisionHandler *handler;
	cpBool swapped;
	cpTimestamp stamp;
	int state;
	int active;
	int next_a, next_b;        /* per-body lists (indices into space->arbs) */
	int record;                /* device record (edits between the two halves of a step) */
};

typedef struct cpPostStepCallback { cpPostStepFunc func; void *key; void *data; } cpPostStepCallback;

struct cpSpace {
	int iterations;
	cpVect gravity;
	cpFloat damping;
	cpFloat idleSpeedThreshold, sleepTimeThreshold;
	cpFloat collisionSlop, collisionBias;
	cpTimestamp collisionPersistence;
	cpDataPointer userData;
	cpTimestamp stamp;
	cpFloat curr_dt;
	int locked;
	cpHashValue shapeIDCounter;

	cpBody *staticBody;
	cpBody _staticBody;
	cpBody **bodies; int nBodies, capBodies;              /* bodies[0] is always staticBody */
	cpShape **shapes; int nShapes, capShapes;
	cpConstraint **constraints; int nConstraints, capConstraints;

	cpCollisionHandler *handlers; int nHandlers, capHandlers;
	cpCollisionHandler defaultHandler;
	cpBool usesWildcards, hasDefaultHandler;
	cpPostStepCallback *postStep; int nPostStep, capPostStep;
	cpBool skipPostStep;

	/* device side */
	cpb200_world *world;
	int device;
	int solverMode;
	cpBool topologyDirty;      /* bodies/shapes/joints added, removed or re-parameterised */
	cpBool bodiesDirty;        /* kinematic state of some body changed on the host */
	cpBool forcesDirty;        /* only forces / torques changed: uploaded as 24 bytes per body */
	cpBool touchDirty;         /* some awake body was activated: its idle timer restarts on the device */
	cpBool paramsDirty;
	cpBool hostStale;          /* device has newer body state than the last download */
	cpBool biasStale;          /* the device holds bias velocities of the last step that the mirrors do not */
	int nBodiesOnDevice;       /* body count of the last upload (host slot == device index below it) */
	int nConstraintsOnDevice;  /* the same for constraints ... */
	int nShapesOnDevice, nVertsOnDevice;
	/* bodies / shapes / constraints added since the last sync sit behind the uploaded ones (slots >= n...OnDevice) and
	 * reach the device as appended ranges (cpb200_world_append_*): nothing else is re-uploaded.  Any removal or edit of
	 * an object that IS on the device sets topologyDirty instead (full re-upload). */
	cpBool anyCustom;          /* some body / spring of this space has (had) a user callback: steps scan for them (slow path) */
	cpBool appendDirty;
	cpBool noAppend;           /* env CPB200_NO_APPEND: always take the full re-upload (comparison / validation) */
	cpBool jointIndexDirty;    /* ... until a removal compacts the host array (cleared by the next upload) */
	cpBool shapeIndexDirty;    /* a removal compacted space->shapes / bodies: device records can no longer be mapped to host objects,
	                              the arbiter mirrors fetched BEFORE that removal stay in use until the next step */
	cpBool jointsDirty;        /* only constraint parameters changed: the joints are re-uploaded, nothing else */
	cpBool shapesDirty;        /* only shape parameters changed: the shapes are re-uploaded, nothing else */
	unsigned fetchStamp;       /* bumped by every download; bodies unpack their record on first access */
	cpBool someMirrorsStale;   /* a download happened and not every body has unpacked its record yet */
	cpBool bbStale, arbStale, jointStale;
	cpArbiter *arbs; int nArbs, capArbs;
	/* user data attached to arbiters (cpArbiterSetUserData) survives the re-download of the mirrors: kept by
	 * unordered shape pair until the pair separates, like the reference keeps it in the cached arbiter */
	struct cpArbData { cpHashValue lo, hi; cpDataPointer data; } *arbData; int nArbData, capArbData;

	/* per-step exchange buffers, page-locked, grown on demand and kept for the life of the space */
	void *xferStates; size_t xferStatesBytes;
	void *xferForces; size_t xferForcesBytes;

	cpBool hasty;
	unsigned long hastyThreads;
};

/* internal helpers */
void cpSpacePrepareDeviceB200(cpSpace *space);
void cpSpaceFetchBodiesB200(cpSpace *space);
void cpSpaceFetchBiasB200(cpSpace *space);
void cpSpaceFetchArbitersB200(cpSpace *space);
void cpSpaceFetchJointsB200(cpSpace *space);
void cpSpaceFetchBBsB200(cpSpace *space);
void cpSpaceMarkTopologyDirty(cpSpace *space);
void cpSpaceMarkBodyDirtyB200(cpBody *body);
void cpSpaceMarkShapeDirtyB200(cpShape *shape);
void cpSpaceMarkConstraintDirtyB200(cpConstraint *c);
void cpBodySetTransformInternal(cpBody *body, cpVect p, cpFloat a);
void cpBodyAccumulateMassFromShapes(cpBody *body);
void cpBodyAddShape(cpBody *body, cpShape *shape);
void cpBodyRemoveShape(cpBody *body, cpShape *shape);
void cpBodyAddConstraint(cpBody *body, cpConstraint *constraint);
void cpBodyRemoveConstraint(cpBody *body, cpConstraint *constraint);
void cpEngineError(const char *what);

static inline cpConstraint *cpConstraintNext(cpConstraint *node, cpBody *body){ return (node->a == body ? node->next_a : node->next_b); }

/* make sure the mirrors of `body` are current before the host reads or edits them */
void cpSpaceDownloadBodiesB200(cpSpace *space);
void cpBodyUnpackB200(cpBody *body);
void cpSpaceUnpackAllB200(cpSpace *space);
/* The body state comes back from the device in one transfer and every mirror takes its record right away, split
 * over the host cores (cpSpaceUnpackAllB200).  Unpacking lazily, per body on first access, measured 20% SLOWER
 * end to end at 1M bodies: it serialises a million 25 ns unpacks onto the caller's thread. */
static inline void cpBodySyncForRead(const cpBody *body){
	cpSpace *space = body->space;
	if(space && space->hostStale) cpSpaceFetchBodiesB200(space);
}

#define cpAssertSpaceUnlocked(space) \
	cpAssertHard(!(space)->locked, \
		"This operation cannot be done safely during a call to cpSpaceStep() or during a query. " \
		"Put these calls into a post-step callback.");

#endif
