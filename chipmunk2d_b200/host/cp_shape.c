/* cp_shape.c -- host mirror of collision shapes (public API of reference cpShape.h:78-197,
 * cpPolyShape.h, chipmunk_unsafe.h:47-60).
 *
 * Geometry is immutable on the device between uploads, so every mutator marks the owning space's
 * topology dirty.  The host also keeps a world-space cache (cpShapeCacheBB / cpShapeUpdate,
 * reference cpShape.c:210-220) computed with the same formulas as the device kernel K2; after a
 * step cpShapeGetBB returns the device's AABBs (downloaded on demand).
 */
#include <string.h>

#include "cp_host.h"

static void
shape_dirty(cpShape *shape)
{
	if(shape->space) cpSpaceMarkShapeDirtyB200(shape);
}

static cpShape *
shape_init(cpShape *shape, int klass, cpBody *body, struct cpShapeMassInfo massInfo)
{
	shape->klass = klass;
	shape->body = body;
	shape->massInfo = massInfo;
	shape->sensor = 0;
	shape->e = 0.0;
	shape->u = 0.0;
	shape->surfaceV = cpvzero;
	shape->type = 0;
	shape->filter.group = CP_NO_GROUP;
	shape->filter.categories = CP_ALL_CATEGORIES;
	shape->filter.mask = CP_ALL_CATEGORIES;
	shape->userData = NULL;
	shape->space = NULL;
	shape->next = NULL;
	shape->prev = NULL;
	shape->index = -1;
	return shape;
}

void
cpShapeDestroy(cpShape *shape)
{
	if(shape && shape->klass == CP_POLY_SHAPE){
		cpPolyShape *poly = (cpPolyShape *)shape;
		cpfree(poly->verts); poly->verts = NULL;
	}
}

void cpShapeFree(cpShape *shape){ if(shape){ cpShapeDestroy(shape); cpfree(shape); } }
cpSpace *cpShapeGetSpace(const cpShape *shape){ return shape->space; }
cpBody *cpShapeGetBody(const cpShape *shape){ return shape->body; }

void
cpShapeSetBody(cpShape *shape, cpBody *body)
{
	cpAssertHard(shape->space == NULL, "You cannot change the body on an active shape. You must remove the shape from the space before changing the body.");
	shape->body = body;
}

cpFloat cpShapeGetMass(cpShape *shape){ return shape->massInfo.m; }

void
cpShapeSetMass(cpShape *shape, cpFloat mass)
{
	cpBody *body = shape->body;
	cpBodyActivate(body);
	shape->massInfo.m = mass;
	cpBodyAccumulateMassFromShapes(body);
}

cpFloat cpShapeGetDensity(cpShape *shape){ return shape->massInfo.m/shape->massInfo.area; }
void cpShapeSetDensity(cpShape *shape, cpFloat density){ cpShapeSetMass(shape, density*shape->massInfo.area); }
cpFloat cpShapeGetMoment(cpShape *shape){ return shape->massInfo.m*shape->massInfo.i; }
cpFloat cpShapeGetArea(cpShape *shape){ return shape->massInfo.area; }
cpVect cpShapeGetCenterOfGravity(cpShape *shape){ return shape->massInfo.cog; }

cpBB
cpShapeGetBB(const cpShape *shape)
{
	if(shape->space && shape->space->bbStale) cpSpaceFetchBBsB200(shape->space);
	return shape->bb;
}

cpBool cpShapeGetSensor(const cpShape *shape){ return shape->sensor; }
void cpShapeSetSensor(cpShape *shape, cpBool sensor){ cpBodyActivate(shape->body); shape->sensor = sensor; shape_dirty(shape); }
cpFloat cpShapeGetElasticity(const cpShape *shape){ return shape->e; }
void cpShapeSetElasticity(cpShape *shape, cpFloat elasticity){ cpAssertHard(elasticity >= 0.0, "Elasticity must be positive."); cpBodyActivate(shape->body); shape->e = elasticity; shape_dirty(shape); }
cpFloat cpShapeGetFriction(const cpShape *shape){ return shape->u; }
void cpShapeSetFriction(cpShape *shape, cpFloat friction){ cpAssertHard(friction >= 0.0, "Friction must be postive."); cpBodyActivate(shape->body); shape->u = friction; shape_dirty(shape); }
cpVect cpShapeGetSurfaceVelocity(const cpShape *shape){ return shape->surfaceV; }
void cpShapeSetSurfaceVelocity(cpShape *shape, cpVect surfaceVelocity){ cpBodyActivate(shape->body); shape->surfaceV = surfaceVelocity; shape_dirty(shape); }
cpDataPointer cpShapeGetUserData(const cpShape *shape){ return shape->userData; }
void cpShapeSetUserData(cpShape *shape, cpDataPointer userData){ shape->userData = userData; }
cpCollisionType cpShapeGetCollisionType(const cpShape *shape){ return shape->type; }
void cpShapeSetCollisionType(cpShape *shape, cpCollisionType collisionType){ cpBodyActivate(shape->body); shape->type = collisionType; shape_dirty(shape); }
cpShapeFilter cpShapeGetFilter(const cpShape *shape){ return shape->filter; }
void cpShapeSetFilter(cpShape *shape, cpShapeFilter filter){ cpBodyActivate(shape->body); shape->filter = filter; shape_dirty(shape); }

/* ---- world-space cache (cacheData: cpShape.c:291-296, 378-405; cpPolyShape.c:39-64) ---- */
cpBB
cpShapeUpdate(cpShape *shape, cpTransform transform)
{
	switch(shape->klass){
	case CP_CIRCLE_SHAPE: {
		cpCircleShape *circle = (cpCircleShape *)shape;
		cpVect c = circle->tc = cpTransformPoint(transform, circle->c);
		return (shape->bb = cpBBNewForCircle(c, circle->r));
	}
	case CP_SEGMENT_SHAPE: {
		cpSegmentShape *seg = (cpSegmentShape *)shape;
		seg->ta = cpTransformPoint(transform, seg->a);
		seg->tb = cpTransformPoint(transform, seg->b);
		seg->tn = cpTransformVect(transform, seg->n);
		cpFloat l, r, b, t;
		if(seg->ta.x < seg->tb.x){ l = seg->ta.x; r = seg->tb.x; } else { l = seg->tb.x; r = seg->ta.x; }
		if(seg->ta.y < seg->tb.y){ b = seg->ta.y; t = seg->tb.y; } else { b = seg->tb.y; t = seg->ta.y; }
		cpFloat rad = seg->r;
		return (shape->bb = cpBBNew(l - rad, b - rad, r + rad, t + rad));
	}
	default: {
		cpPolyShape *poly = (cpPolyShape *)shape;
		cpFloat l = (cpFloat)INFINITY, r = -(cpFloat)INFINITY, b = (cpFloat)INFINITY, t = -(cpFloat)INFINITY;
		for(int i = 0; i < poly->count; i++){
			cpVect v = poly->tverts[i] = cpTransformPoint(transform, poly->verts[i]);
			poly->tnormals[i] = cpTransformVect(transform, poly->normals[i]);
			l = cpfmin(l, v.x); r = cpfmax(r, v.x);
			b = cpfmin(b, v.y); t = cpfmax(t, v.y);
		}
		cpFloat radius = poly->r;
		return (shape->bb = cpBBNew(l - radius, b - radius, r + radius, t + radius));
	}
	}
}

cpBB
cpShapeCacheBB(cpShape *shape)
{
	cpBodySyncForRead(shape->body);
	return cpShapeUpdate(shape, shape->body->transform);
}

/* Narrowphase of two shapes (cpShape.c:259-283).  Runs on the device (K5) for shapes that live in
 * the same space; there is no host narrowphase. */
cpContactPointSet
cpShapesCollide(const cpShape *a, const cpShape *b)
{
	cpContactPointSet set;
	memset(&set, 0, sizeof(set));
	cpAssertHard(a->space && a->space == b->space, "cpShapesCollide() on the B200 path needs both shapes in the same space (the narrowphase runs on the device).");
	extern void cpSpaceCollidePairB200(cpSpace *space, const cpShape *a, const cpShape *b, cpContactPointSet *out);
	cpSpaceCollidePairB200(a->space, a, b, &set);
	return set;
}

/* ---- circles (cpShape.c:285-336) ---- */
cpCircleShape *cpCircleShapeAlloc(void){ return (cpCircleShape *)cpcalloc(1, sizeof(cpCircleShape)); }

static struct cpShapeMassInfo
circle_mass_info(cpFloat mass, cpFloat radius, cpVect center)
{
	struct cpShapeMassInfo info = {mass, cpMomentForCircle(1.0, 0.0, radius, cpvzero), center, cpAreaForCircle(0.0, radius)};
	return info;
}

cpCircleShape *
cpCircleShapeInit(cpCircleShape *circle, cpBody *body, cpFloat radius, cpVect offset)
{
	circle->c = offset;
	circle->r = radius;
	shape_init((cpShape *)circle, CP_CIRCLE_SHAPE, body, circle_mass_info(0.0, radius, offset));
	return circle;
}

cpShape *cpCircleShapeNew(cpBody *body, cpFloat radius, cpVect offset){ return (cpShape *)cpCircleShapeInit(cpCircleShapeAlloc(), body, radius, offset); }
cpVect cpCircleShapeGetOffset(const cpShape *shape){ cpAssertHard(shape->klass == CP_CIRCLE_SHAPE, "Shape is not a circle shape."); return ((cpCircleShape *)shape)->c; }
cpFloat cpCircleShapeGetRadius(const cpShape *shape){ cpAssertHard(shape->klass == CP_CIRCLE_SHAPE, "Shape is not a circle shape."); return ((cpCircleShape *)shape)->r; }

void
cpCircleShapeSetRadius(cpShape *shape, cpFloat radius)
{
	cpAssertHard(shape->klass == CP_CIRCLE_SHAPE, "Shape is not a circle shape.");
	cpCircleShape *circle = (cpCircleShape *)shape;
	circle->r = radius;
	cpFloat mass = shape->massInfo.m;
	shape->massInfo = circle_mass_info(mass, circle->r, circle->c);
	if(mass > 0.0) cpBodyAccumulateMassFromShapes(shape->body);
	shape_dirty(shape);
}

void
cpCircleShapeSetOffset(cpShape *shape, cpVect offset)
{
	cpAssertHard(shape->klass == CP_CIRCLE_SHAPE, "Shape is not a circle shape.");
	cpCircleShape *circle = (cpCircleShape *)shape;
	circle->c = offset;
	cpFloat mass = shape->massInfo.m;
	shape->massInfo = circle_mass_info(mass, circle->r, circle->c);
	if(mass > 0.0) cpBodyAccumulateMassFromShapes(shape->body);
	shape_dirty(shape);
}

/* ---- segments (cpShape.c:338-580) ---- */
cpSegmentShape *cpSegmentShapeAlloc(void){ return (cpSegmentShape *)cpcalloc(1, sizeof(cpSegmentShape)); }

static struct cpShapeMassInfo
segment_mass_info(cpFloat mass, cpVect a, cpVect b, cpFloat r)
{
	struct cpShapeMassInfo info = {mass, cpMomentForBox(1.0, cpvdist(a, b) + 2.0*r, 2.0*r), cpvlerp(a, b, 0.5), cpAreaForSegment(a, b, r)};
	return info;
}

cpSegmentShape *
cpSegmentShapeInit(cpSegmentShape *seg, cpBody *body, cpVect a, cpVect b, cpFloat r)
{
	seg->a = a;
	seg->b = b;
	seg->n = cpvrperp(cpvnormalize(cpvsub(b, a)));
	seg->r = r;
	seg->a_tangent = cpvzero;
	seg->b_tangent = cpvzero;
	shape_init((cpShape *)seg, CP_SEGMENT_SHAPE, body, segment_mass_info(0.0, a, b, r));
	return seg;
}

cpShape *cpSegmentShapeNew(cpBody *body, cpVect a, cpVect b, cpFloat r){ return (cpShape *)cpSegmentShapeInit(cpSegmentShapeAlloc(), body, a, b, r); }
cpVect cpSegmentShapeGetA(const cpShape *shape){ cpAssertHard(shape->klass == CP_SEGMENT_SHAPE, "Shape is not a segment shape."); return ((cpSegmentShape *)shape)->a; }
cpVect cpSegmentShapeGetB(const cpShape *shape){ cpAssertHard(shape->klass == CP_SEGMENT_SHAPE, "Shape is not a segment shape."); return ((cpSegmentShape *)shape)->b; }
cpVect cpSegmentShapeGetNormal(const cpShape *shape){ cpAssertHard(shape->klass == CP_SEGMENT_SHAPE, "Shape is not a segment shape."); return ((cpSegmentShape *)shape)->n; }
cpFloat cpSegmentShapeGetRadius(const cpShape *shape){ cpAssertHard(shape->klass == CP_SEGMENT_SHAPE, "Shape is not a segment shape."); return ((cpSegmentShape *)shape)->r; }

void
cpSegmentShapeSetNeighbors(cpShape *shape, cpVect prev, cpVect next)
{
	cpAssertHard(shape->klass == CP_SEGMENT_SHAPE, "Shape is not a segment shape.");
	cpSegmentShape *seg = (cpSegmentShape *)shape;
	seg->a_tangent = cpvsub(prev, seg->a);
	seg->b_tangent = cpvsub(next, seg->b);
	shape_dirty(shape);
}

void
cpSegmentShapeSetEndpoints(cpShape *shape, cpVect a, cpVect b)
{
	cpAssertHard(shape->klass == CP_SEGMENT_SHAPE, "Shape is not a segment shape.");
	cpSegmentShape *seg = (cpSegmentShape *)shape;
	seg->a = a;
	seg->b = b;
	seg->n = cpvperp(cpvnormalize(cpvsub(b, a)));
	cpFloat mass = shape->massInfo.m;
	shape->massInfo = segment_mass_info(mass, seg->a, seg->b, seg->r);
	if(mass > 0.0) cpBodyAccumulateMassFromShapes(shape->body);
	shape_dirty(shape);
}

void
cpSegmentShapeSetRadius(cpShape *shape, cpFloat radius)
{
	cpAssertHard(shape->klass == CP_SEGMENT_SHAPE, "Shape is not a segment shape.");
	cpSegmentShape *seg = (cpSegmentShape *)shape;
	seg->r = radius;
	cpFloat mass = shape->massInfo.m;
	shape->massInfo = segment_mass_info(mass, seg->a, seg->b, seg->r);
	if(mass > 0.0) cpBodyAccumulateMassFromShapes(shape->body);
	shape_dirty(shape);
}

/* ---- polygons (cpPolyShape.c:147-324) ---- */
cpPolyShape *cpPolyShapeAlloc(void){ return (cpPolyShape *)cpcalloc(1, sizeof(cpPolyShape)); }

static void
poly_set_verts(cpPolyShape *poly, int count, const cpVect *verts)
{
	cpfree(poly->verts);
	poly->count = count;
	poly->verts = (cpVect *)cpcalloc(4*(size_t)(count > 0 ? count : 1), sizeof(cpVect));
	poly->normals = poly->verts + count;
	poly->tverts = poly->verts + 2*count;
	poly->tnormals = poly->verts + 3*count;
	for(int i = 0; i < count; i++){
		/* plane i holds vertex i and the outward normal of the edge (i-1 -> i) (cpPolyShape.c:157-164) */
		cpVect a = verts[(i - 1 + count)%count], b = verts[i];
		poly->verts[i] = b;
		poly->normals[i] = cpvnormalize(cpvrperp(cpvsub(b, a)));
	}
}

static struct cpShapeMassInfo
poly_mass_info(cpFloat mass, int count, const cpVect *verts, cpFloat radius)
{
	cpVect centroid = cpCentroidForPoly(count, verts);
	struct cpShapeMassInfo info = {mass, cpMomentForPoly(1.0, count, verts, cpvneg(centroid), radius), centroid, cpAreaForPoly(count, verts, radius)};
	return info;
}

cpPolyShape *
cpPolyShapeInitRaw(cpPolyShape *poly, cpBody *body, int count, const cpVect *verts, cpFloat radius)
{
	shape_init((cpShape *)poly, CP_POLY_SHAPE, body, poly_mass_info(0.0, count, verts, radius));
	poly->verts = NULL;
	poly_set_verts(poly, count, verts);
	poly->r = radius;
	return poly;
}

cpPolyShape *
cpPolyShapeInit(cpPolyShape *poly, cpBody *body, int count, const cpVect *verts, cpTransform transform, cpFloat radius)
{
	cpVect *hull = (cpVect *)cpcalloc((size_t)(count > 0 ? count : 1), sizeof(cpVect));
	for(int i = 0; i < count; i++) hull[i] = cpTransformPoint(transform, verts[i]);
	int hullCount = cpConvexHull(count, hull, hull, NULL, 0.0);
	cpPolyShapeInitRaw(poly, body, hullCount, hull, radius);
	cpfree(hull);
	return poly;
}

cpShape *cpPolyShapeNew(cpBody *body, int count, const cpVect *verts, cpTransform transform, cpFloat radius){ return (cpShape *)cpPolyShapeInit(cpPolyShapeAlloc(), body, count, verts, transform, radius); }
cpShape *cpPolyShapeNewRaw(cpBody *body, int count, const cpVect *verts, cpFloat radius){ return (cpShape *)cpPolyShapeInitRaw(cpPolyShapeAlloc(), body, count, verts, radius); }

cpPolyShape *
cpBoxShapeInit2(cpPolyShape *poly, cpBody *body, cpBB box, cpFloat radius)
{
	cpVect verts[4] = {cpv(box.r, box.b), cpv(box.r, box.t), cpv(box.l, box.t), cpv(box.l, box.b)};
	return cpPolyShapeInitRaw(poly, body, 4, verts, radius);
}

cpPolyShape *
cpBoxShapeInit(cpPolyShape *poly, cpBody *body, cpFloat width, cpFloat height, cpFloat radius)
{
	cpFloat hw = width/2.0, hh = height/2.0;
	return cpBoxShapeInit2(poly, body, cpBBNew(-hw, -hh, hw, hh), radius);
}

cpShape *cpBoxShapeNew(cpBody *body, cpFloat width, cpFloat height, cpFloat radius){ return (cpShape *)cpBoxShapeInit(cpPolyShapeAlloc(), body, width, height, radius); }
cpShape *cpBoxShapeNew2(cpBody *body, cpBB box, cpFloat radius){ return (cpShape *)cpBoxShapeInit2(cpPolyShapeAlloc(), body, box, radius); }

int cpPolyShapeGetCount(const cpShape *shape){ cpAssertHard(shape->klass == CP_POLY_SHAPE, "Shape is not a poly shape."); return ((cpPolyShape *)shape)->count; }

cpVect
cpPolyShapeGetVert(const cpShape *shape, int i)
{
	cpAssertHard(shape->klass == CP_POLY_SHAPE, "Shape is not a poly shape.");
	int count = cpPolyShapeGetCount(shape);
	cpAssertHard(0 <= i && i < count, "Index out of range.");
	return ((cpPolyShape *)shape)->verts[i];
}

cpFloat cpPolyShapeGetRadius(const cpShape *shape){ cpAssertHard(shape->klass == CP_POLY_SHAPE, "Shape is not a poly shape."); return ((cpPolyShape *)shape)->r; }

void
cpPolyShapeSetVertsRaw(cpShape *shape, int count, cpVect *verts)
{
	cpAssertHard(shape->klass == CP_POLY_SHAPE, "Shape is not a poly shape.");
	cpPolyShape *poly = (cpPolyShape *)shape;
	poly_set_verts(poly, count, verts);
	cpFloat mass = shape->massInfo.m;
	shape->massInfo = poly_mass_info(mass, count, verts, poly->r);
	if(mass > 0.0) cpBodyAccumulateMassFromShapes(shape->body);
	shape_dirty(shape);
}

void
cpPolyShapeSetVerts(cpShape *shape, int count, cpVect *verts, cpTransform transform)
{
	cpVect *hull = (cpVect *)cpcalloc((size_t)(count > 0 ? count : 1), sizeof(cpVect));
	for(int i = 0; i < count; i++) hull[i] = cpTransformPoint(transform, verts[i]);
	int hullCount = cpConvexHull(count, hull, hull, NULL, 0.0);
	cpPolyShapeSetVertsRaw(shape, hullCount, hull);
	cpfree(hull);
}

void
cpPolyShapeSetRadius(cpShape *shape, cpFloat radius)
{
	cpAssertHard(shape->klass == CP_POLY_SHAPE, "Shape is not a poly shape.");
	((cpPolyShape *)shape)->r = radius;
	shape_dirty(shape);
}
