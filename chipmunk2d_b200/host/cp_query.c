/* cp_query.c -- space and shape queries of the public API (reference src/cpSpaceQuery.c:24-246,
 * cpShape.c:223-258) on top of the C ABI's query scans.
 *
 * Every query is one data-parallel pass over the device's world-space shape cache; the host only marshals
 * arguments, downloads the hit records (ascending shape index) and runs the user's callback per hit with
 * the space locked, like the reference.  Queries see the state of the last step plus any host-side edits
 * made since (they are uploaded first).  There is no host geometry path: shapes must be part of a space,
 * except the probe shape of cpSpaceShapeQuery, whose world-space form is computed here with cpShapeUpdate.
 */
#include <string.h>

#include "cp_host.h"

static cpb200_filter
to_filter(cpShapeFilter f)
{
	cpb200_filter o;
	o.group = (uint64_t)f.group; o.categories = f.categories; o.mask = f.mask;
	return o;
}

static cpSpace *
query_space(const cpShape *shape, const char *what)
{
	cpAssertHard(shape->space != NULL, what);
	cpSpacePrepareDeviceB200(shape->space);
	return shape->space;
}

cpFloat
cpShapePointQuery(const cpShape *shape, cpVect p, cpPointQueryInfo *info)
{
	cpPointQueryInfo blank = {NULL, cpvzero, (cpFloat)INFINITY, cpvzero};
	if(info) (*info) = blank; else info = &blank;
	cpSpace *space = query_space(shape, "cpShapePointQuery() on the B200 path needs a shape that was added to a space (the test runs on the device).");
	double pt[2] = {p.x, p.y};
	cpb200_query_hit h;
	if(cpb200_world_shape_point_query(space->world, shape->index, pt, &h) != 1) cpEngineError("cpShapePointQuery");
	info->shape = shape; info->point = cpv(h.point[0], h.point[1]); info->distance = h.d; info->gradient = cpv(h.g[0], h.g[1]);
	return info->distance;
}

cpBool
cpShapeSegmentQuery(const cpShape *shape, cpVect a, cpVect b, cpFloat radius, cpSegmentQueryInfo *info)
{
	cpSegmentQueryInfo blank = {NULL, b, cpvzero, 1.0};
	if(info) (*info) = blank; else info = &blank;
	cpSpace *space = query_space(shape, "cpShapeSegmentQuery() on the B200 path needs a shape that was added to a space (the test runs on the device).");
	double pa[2] = {a.x, a.y}, pb[2] = {b.x, b.y};
	cpb200_query_hit h;
	int n = cpb200_world_shape_segment_query(space->world, shape->index, pa, pb, radius, &h);
	if(n < 0) cpEngineError("cpShapeSegmentQuery");
	if(n == 1){ info->shape = shape; info->point = cpv(h.point[0], h.point[1]); info->normal = cpv(h.g[0], h.g[1]); info->alpha = h.d; }
	return (info->shape != NULL);
}

/* growing download of all hits of one scan */
static cpb200_query_hit *
fetch_hits(cpSpace *space, int kind, const double *a, const double *b, cpFloat radius, cpShapeFilter filter, int *count)
{
	cpb200_filter f = to_filter(filter);
	int cap = 256;
	for(;;){
		cpb200_query_hit *hits = (cpb200_query_hit *)cpcalloc((size_t)cap, sizeof(cpb200_query_hit));
		int n = (kind == 0 ? cpb200_world_point_query(space->world, 0, a, radius, &f, cap, hits)
		                   : cpb200_world_segment_query(space->world, 0, a, b, radius, &f, cap, hits));
		if(n < 0) cpEngineError("space query");
		if(n <= cap){ *count = n; return hits; }
		cpfree(hits);
		cap = n;
	}
}

void
cpSpacePointQuery(cpSpace *space, cpVect point, cpFloat maxDistance, cpShapeFilter filter, cpSpacePointQueryFunc func, void *data)
{
	cpSpacePrepareDeviceB200(space);
	double p[2] = {point.x, point.y};
	int n = 0;
	cpb200_query_hit *hits = fetch_hits(space, 0, p, NULL, maxDistance, filter, &n);
	space->locked++;
	for(int i = 0; i < n; i++) func(space->shapes[hits[i].shape], cpv(hits[i].point[0], hits[i].point[1]), hits[i].d, cpv(hits[i].g[0], hits[i].g[1]), data);
	space->locked--;
	cpfree(hits);
}

cpShape *
cpSpacePointQueryNearest(cpSpace *space, cpVect point, cpFloat maxDistance, cpShapeFilter filter, cpPointQueryInfo *out)
{
	cpPointQueryInfo info = {NULL, cpvzero, maxDistance, cpvzero};
	if(out) (*out) = info; else out = &info;
	cpSpacePrepareDeviceB200(space);
	double p[2] = {point.x, point.y};
	cpb200_filter f = to_filter(filter);
	cpb200_query_hit h;
	int n = cpb200_world_point_query_nearest(space->world, 0, p, maxDistance, &f, &h);
	if(n < 0) cpEngineError("cpSpacePointQueryNearest");
	if(n == 1){ out->shape = space->shapes[h.shape]; out->point = cpv(h.point[0], h.point[1]); out->distance = h.d; out->gradient = cpv(h.g[0], h.g[1]); }
	return (cpShape *)out->shape;
}

void
cpSpaceSegmentQuery(cpSpace *space, cpVect start, cpVect end, cpFloat radius, cpShapeFilter filter, cpSpaceSegmentQueryFunc func, void *data)
{
	cpSpacePrepareDeviceB200(space);
	double a[2] = {start.x, start.y}, b[2] = {end.x, end.y};
	int n = 0;
	cpb200_query_hit *hits = fetch_hits(space, 1, a, b, radius, filter, &n);
	space->locked++;
	for(int i = 0; i < n; i++) func(space->shapes[hits[i].shape], cpv(hits[i].point[0], hits[i].point[1]), cpv(hits[i].g[0], hits[i].g[1]), hits[i].d, data);
	space->locked--;
	cpfree(hits);
}

cpShape *
cpSpaceSegmentQueryFirst(cpSpace *space, cpVect start, cpVect end, cpFloat radius, cpShapeFilter filter, cpSegmentQueryInfo *out)
{
	cpSegmentQueryInfo info = {NULL, end, cpvzero, 1.0};
	if(out) (*out) = info; else out = &info;
	cpSpacePrepareDeviceB200(space);
	double a[2] = {start.x, start.y}, b[2] = {end.x, end.y};
	cpb200_filter f = to_filter(filter);
	cpb200_query_hit h;
	int n = cpb200_world_segment_query_first(space->world, 0, a, b, radius, &f, &h);
	if(n < 0) cpEngineError("cpSpaceSegmentQueryFirst");
	if(n == 1){ out->shape = space->shapes[h.shape]; out->point = cpv(h.point[0], h.point[1]); out->normal = cpv(h.g[0], h.g[1]); out->alpha = h.d; }
	return (cpShape *)out->shape;
}

void
cpSpaceBBQuery(cpSpace *space, cpBB bb, cpShapeFilter filter, cpSpaceBBQueryFunc func, void *data)
{
	cpSpacePrepareDeviceB200(space);
	double box[4] = {bb.l, bb.b, bb.r, bb.t};
	cpb200_filter f = to_filter(filter);
	int cap = 256, n = 0;
	int32_t *ids = NULL;
	for(;;){
		ids = (int32_t *)cpcalloc((size_t)cap, sizeof(int32_t));
		n = cpb200_world_bb_query(space->world, 0, box, &f, cap, ids);
		if(n < 0) cpEngineError("cpSpaceBBQuery");
		if(n <= cap) break;
		cpfree(ids);
		cap = n;
	}
	space->locked++;
	for(int i = 0; i < n; i++) func(space->shapes[ids[i]], data);
	space->locked--;
	cpfree(ids);
}

cpBool
cpSpaceShapeQuery(cpSpace *space, cpShape *shape, cpSpaceShapeQueryFunc func, void *data)
{
	cpSpacePrepareDeviceB200(space);
	cpBody *body = shape->body;
	/* world-space form of the probe, exactly what cacheData leaves behind (cpSpaceQuery.c:235-236) */
	cpBB bb = (body ? cpShapeCacheBB(shape) : shape->bb);
	cpb200_query_shape q;
	memset(&q, 0, sizeof(q));
	q.type = shape->klass;
	q.self = (shape->space == space ? shape->index : -1);
	q.bb[0] = bb.l; q.bb[1] = bb.b; q.bb[2] = bb.r; q.bb[3] = bb.t;
	q.filter = to_filter(shape->filter);
	cpVect rot = (body ? cpBodyGetRotation(body) : cpv(1.0, 0.0));
	q.rot[0] = rot.x; q.rot[1] = rot.y;
	double *planes = NULL;
	switch(shape->klass){
	case CP_CIRCLE_SHAPE: { cpCircleShape *c = (cpCircleShape *)shape; q.r = c->r; q.a[0] = c->tc.x; q.a[1] = c->tc.y; break; }
	case CP_SEGMENT_SHAPE: {
		cpSegmentShape *s = (cpSegmentShape *)shape;
		q.r = s->r; q.a[0] = s->ta.x; q.a[1] = s->ta.y; q.b[0] = s->tb.x; q.b[1] = s->tb.y; q.n[0] = s->tn.x; q.n[1] = s->tn.y;
		q.a_tangent[0] = s->a_tangent.x; q.a_tangent[1] = s->a_tangent.y; q.b_tangent[0] = s->b_tangent.x; q.b_tangent[1] = s->b_tangent.y;
		break;
	}
	default: {
		cpPolyShape *p = (cpPolyShape *)shape;
		q.r = p->r; q.count = p->count;
		planes = (double *)cpcalloc((size_t)p->count, 4*sizeof(double));
		for(int i = 0; i < p->count; i++){ planes[4*i] = p->tverts[i].x; planes[4*i + 1] = p->tverts[i].y; planes[4*i + 2] = p->tnormals[i].x; planes[4*i + 3] = p->tnormals[i].y; }
		break;
	}
	}
	int cap = 64, n = 0;
	cpb200_shape_hit *hits = NULL;
	for(;;){
		hits = (cpb200_shape_hit *)cpcalloc((size_t)cap, sizeof(cpb200_shape_hit));
		n = cpb200_world_shape_query(space->world, 0, &q, planes, cap, hits);
		if(n < 0) cpEngineError("cpSpaceShapeQuery");
		if(n <= cap) break;
		cpfree(hits);
		cap = n;
	}
	cpBool any = cpFalse;
	space->locked++;
	for(int i = 0; i < n; i++){
		cpShape *other = space->shapes[hits[i].shape];
		cpContactPointSet set;
		memset(&set, 0, sizeof(set));
		set.count = hits[i].count;
		set.normal = cpv(hits[i].normal[0], hits[i].normal[1]);
		for(int k = 0; k < set.count && k < CP_MAX_CONTACTS_PER_ARBITER; k++){
			set.points[k].pointA = cpv(hits[i].points[k][0], hits[i].points[k][1]);
			set.points[k].pointB = cpv(hits[i].points[k][2], hits[i].points[k][3]);
			set.points[k].distance = hits[i].points[k][4];
		}
		if(func) func(other, &set, data);
		/* the reference keeps the flag of the LAST callback (cpSpaceQuery.c:218); any non-sensor contact counts here */
		if(!(shape->sensor || other->sensor)) any = cpTrue;
	}
	space->locked--;
	cpfree(hits);
	if(planes) cpfree(planes);
	return any;
}
