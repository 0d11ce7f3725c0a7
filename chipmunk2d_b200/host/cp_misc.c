/* cp_misc.c -- assertion sink, mass/area helpers and the convex hull
 * (reference src/chipmunk.c:31-274).  Setup-time code: runs on the host. */
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "cp_host.h"

const char *cpVersionString = "7.0.3";

void
cpMessage(const char *condition, const char *file, int line, int isError, int isHardError, const char *message, ...)
{
	fprintf(stderr, (isError ? "Aborting due to Chipmunk error: " : "Chipmunk warning: "));
	va_list vargs;
	va_start(vargs, message);
	vfprintf(stderr, message, vargs);
	va_end(vargs);
	fprintf(stderr, "\n\tFailed condition: %s\n\tSource:%s:%d\n", condition, file, line);
	(void)isHardError;
}

/* a failing device call is a hard error: there is no CPU path to fall back to */
void
cpEngineError(const char *what)
{
	fprintf(stderr, "Aborting due to Chipmunk error: the B200 step engine failed in %s: %s\n", what, cpb200_last_error());
	abort();
}

cpFloat cpMomentForCircle(cpFloat m, cpFloat r1, cpFloat r2, cpVect offset){ return m*(0.5*(r1*r1 + r2*r2) + cpvlengthsq(offset)); }
cpFloat cpAreaForCircle(cpFloat r1, cpFloat r2){ return (cpFloat)CP_PI*cpfabs(r1*r1 - r2*r2); }

cpFloat
cpMomentForSegment(cpFloat m, cpVect a, cpVect b, cpFloat r)
{
	cpVect offset = cpvlerp(a, b, 0.5);
	cpFloat length = cpvdist(b, a) + 2.0*r;
	return m*((length*length + 4.0*r*r)/12.0 + cpvlengthsq(offset));
}

cpFloat cpAreaForSegment(cpVect a, cpVect b, cpFloat r){ return r*((cpFloat)CP_PI*r + 2.0*cpvdist(a, b)); }

cpFloat
cpMomentForPoly(cpFloat m, int count, const cpVect *verts, cpVect offset, cpFloat r)
{
	(void)r;
	if(count == 2) return cpMomentForSegment(m, verts[0], verts[1], 0.0);
	cpFloat num = 0.0, den = 0.0;
	for(int i = 0; i < count; i++){
		cpVect v1 = cpvadd(verts[i], offset);
		cpVect v2 = cpvadd(verts[(i + 1)%count], offset);
		cpFloat a = cpvcross(v2, v1);
		cpFloat b = cpvdot(v1, v1) + cpvdot(v1, v2) + cpvdot(v2, v2);
		num += a*b;
		den += a;
	}
	return (m*num)/(6.0*den);
}

cpFloat
cpAreaForPoly(const int count, const cpVect *verts, cpFloat r)
{
	cpFloat area = 0.0, perimeter = 0.0;
	for(int i = 0; i < count; i++){
		cpVect v1 = verts[i], v2 = verts[(i + 1)%count];
		area += cpvcross(v1, v2);
		perimeter += cpvdist(v1, v2);
	}
	return r*(CP_PI*cpfabs(r) + perimeter) + area/2.0;
}

cpVect
cpCentroidForPoly(const int count, const cpVect *verts)
{
	cpFloat sum = 0.0;
	cpVect vsum = cpvzero;
	for(int i = 0; i < count; i++){
		cpVect v1 = verts[i], v2 = verts[(i + 1)%count];
		cpFloat cross = cpvcross(v1, v2);
		sum += cross;
		vsum = cpvadd(vsum, cpvmult(cpvadd(v1, v2), cross));
	}
	return cpvmult(vsum, 1.0/(3.0*sum));
}

cpFloat cpMomentForBox(cpFloat m, cpFloat width, cpFloat height){ return m*(width*width + height*height)/12.0; }

cpFloat
cpMomentForBox2(cpFloat m, cpBB box)
{
	cpFloat width = box.r - box.l, height = box.t - box.b;
	cpVect offset = cpvmult(cpv(box.l + box.r, box.b + box.t), 0.5);
	return cpMomentForBox(m, width, height) + m*cpvlengthsq(offset);
}

/* Convex hull with the reference's output convention (src/chipmunk.c:250-274): counter-clockwise,
 * starting at the minimum-x (then minimum-y) vertex, points within tol of an edge dropped.
 * Implemented as a monotone chain over the lexicographically sorted points. */
static int
hull_cmp(const void *pa, const void *pb)
{
	const cpVect *a = (const cpVect *)pa, *b = (const cpVect *)pb;
	if(a->x != b->x) return (a->x < b->x ? -1 : 1);
	if(a->y != b->y) return (a->y < b->y ? -1 : 1);
	return 0;
}

static cpBool
hull_keeps(cpVect o, cpVect a, cpVect b, cpFloat tol)
{
	/* a stays on the chain o -> a -> b only for a strict left turn (by more than tol) */
	cpVect d = cpvsub(b, o);
	return cpvcross(d, cpvsub(a, o)) < -tol*cpvlength(d);
}

int
cpConvexHull(int count, const cpVect *verts, cpVect *result, int *first, cpFloat tol)
{
	if(count <= 0){ if(first) *first = 0; return 0; }
	int start = 0;
	for(int i = 1; i < count; i++){
		if(verts[i].x < verts[start].x || (verts[i].x == verts[start].x && verts[i].y < verts[start].y)) start = i;
	}
	if(first) *first = start;

	cpVect *pts = (cpVect *)cpcalloc((size_t)count, sizeof(cpVect));
	memcpy(pts, verts, sizeof(cpVect)*(size_t)count);
	qsort(pts, (size_t)count, sizeof(cpVect), hull_cmp);
	int n = 0;
	for(int i = 0; i < count; i++){ if(n == 0 || !cpveql(pts[i], pts[n - 1])) pts[n++] = pts[i]; }
	if(n == 1){ result[0] = pts[0]; cpfree(pts); return 1; }

	cpVect *hull = (cpVect *)cpcalloc((size_t)(2*n + 2), sizeof(cpVect));
	int k = 0;
	for(int i = 0; i < n; i++){                        /* lower chain, left to right */
		while(k >= 2 && !hull_keeps(hull[k - 2], hull[k - 1], pts[i], tol)) k--;
		hull[k++] = pts[i];
	}
	for(int i = n - 2, lower = k + 1; i >= 0; i--){    /* upper chain, right to left */
		while(k >= lower && !hull_keeps(hull[k - 2], hull[k - 1], pts[i], tol)) k--;
		hull[k++] = pts[i];
	}
	k--;                                               /* the start point was appended twice */
	if(k < 1) k = 1;
	memcpy(result, hull, sizeof(cpVect)*(size_t)k);
	cpfree(hull); cpfree(pts);
	return k;
}
