// k_query.cuh -- space queries against the shapes' world-space cache (cpSpaceQuery.c:24-246).
//
// The reference walks its two BBTrees with the query's bounding box and runs the exact per-shape test on
// every leaf it reaches; which shapes are REPORTED is decided by the exact test alone (point: distance <
// maxDistance, segment: a hit with alpha in [0,1], bb: cpBBIntersects on the cached AABB).  A single query
// against a million shapes is a 20 us data-parallel scan on this machine, cheaper than any tree descent by
// one thread, so every query kind is one thread per shape over the cache that k_shape_cache leaves behind:
//   filter reject (cpShapeFilterReject, chipmunk_private.h:144-155) -> AABB reject -> exact test -> append
// "Nearest"/"first" variants reduce to the best hit on the device (ordered-key atomicMin, then the winner
// writes its record), so only one record crosses the bus.
// Per-shape tests restate cpShape.c:298-330, 405-455, cpPolyShape.c:66-145 and CircleSegmentQuery
// (chipmunk_private.h:121-142) with the reference's operation order (bit-identical on the same cache).
#pragma once
#include "cpb_world.h"

#define CPB_MAGIC_EPSILON 1e-5

struct QFilter { uint64_t group; uint32_t categories, mask; };

// hit record shared by the point query (d = distance, g = gradient) and the segment query (d = alpha, g = normal)
struct QHit { int shape; int pad; double px, py, d, gx, gy; };

struct QParams {
	int kind;            // 0 point, 1 segment, 2 bb
	int space;           // only shapes of this space (-1 = any)
	int only_shape;      // restrict to one shape (-1 = all); no filter / range test then (cpShapePointQuery)
	int skip_sensors;
	V2 a, b;             // point: a | segment: a -> b | bb: (a.x, a.y) = (l, b), (b.x, b.y) = (r, t)
	double radius;       // point: maxDistance | segment: radius
	QFilter filter;
};

CPB_DEVICE bool q_filter_reject(const DShapes &S, int s, const QFilter &f){
	uint64_t g = S.group[s];
	return (g != 0 && g == f.group) || (S.cat[s] & f.mask) == 0 || (f.categories & S.mask[s]) == 0;
}

// cpClosetPointOnSegment (chipmunk.h:183-188)
CPB_DEVICE V2 q_closest_on_segment(V2 p, V2 a, V2 b){
	V2 delta = vsub(a, b);
	double t = fclamp01_cp(vdot(delta, vsub(p, b))/vlensq(delta));
	return vadd(b, vmul(delta, t));
}

CPB_DEVICE void q_point_circle(V2 tc, double r, V2 p, QHit &h){
	V2 delta = vsub(p, tc);
	double d = vlen(delta);
	double r_over_d = (d > 0.0 ? r/d : r);
	V2 pt = vadd(tc, vmul(delta, r_over_d));
	h.px = pt.x; h.py = pt.y; h.d = d - r;
	V2 g = (d > CPB_MAGIC_EPSILON ? vmul(delta, 1.0/d) : v2(0.0, 1.0));
	h.gx = g.x; h.gy = g.y;
}

CPB_DEVICE void q_point_segment(V2 ta, V2 tb, V2 n_local, double r, V2 p, QHit &h){
	V2 closest = q_closest_on_segment(p, ta, tb);
	V2 delta = vsub(p, closest);
	double d = vlen(delta);
	V2 g = vmul(delta, 1.0/d);
	V2 pt = (d ? vadd(closest, vmul(g, r)) : closest);
	h.px = pt.x; h.py = pt.y; h.d = d - r;
	// the reference falls back to seg->n, the body-local normal (cpShape.c:421)
	V2 gr = (d > CPB_MAGIC_EPSILON ? g : n_local);
	h.gx = gr.x; h.gy = gr.y;
}

CPB_DEVICE void q_point_poly(const V2 *pv, const V2 *pn, int count, double r, V2 p, QHit &h){
	V2 v0 = pv[count - 1];
	double minDist = INFINITY;
	V2 closestPoint = v2(0, 0), closestNormal = v2(0, 0);
	bool outside = false;
	for(int i = 0; i < count; i++){
		V2 v1 = pv[i];
		outside = outside || (vdot(pn[i], vsub(p, v1)) > 0.0);
		V2 closest = q_closest_on_segment(p, v0, v1);
		double dist = vlen(vsub(p, closest));
		if(dist < minDist){ minDist = dist; closestPoint = closest; closestNormal = pn[i]; }
		v0 = v1;
	}
	double dist = (outside ? minDist : -minDist);
	V2 g = vmul(vsub(p, closestPoint), 1.0/dist);
	V2 pt = vadd(closestPoint, vmul(g, r));
	h.px = pt.x; h.py = pt.y; h.d = dist - r;
	V2 gr = (minDist > CPB_MAGIC_EPSILON ? g : closestNormal);
	h.gx = gr.x; h.gy = gr.y;
}

CPB_DEVICE void q_shape_point(const DShapes &S, int s, V2 p, QHit &h){
	int type = S.type[s];
	h.shape = s; h.pad = 0;
	if(type == CPB200_SHAPE_CIRCLE) q_point_circle(S.wa[s], S.r[s], p, h);
	else if(type == CPB200_SHAPE_SEGMENT) q_point_segment(S.wa[s], S.wb[s], S.ln[s], S.r[s], p, h);
	else q_point_poly(S.wpv + S.poff[s], S.wpn + S.poff[s], S.pcount[s], S.r[s], p, h);
}

// segment-query result while it is being built: hit == false means info->shape == NULL
struct QSeg { bool hit; V2 point, normal; double alpha; };

CPB_DEVICE void q_circle_segment(V2 center, double r1, V2 a, V2 b, double r2, QSeg &info){
	V2 da = vsub(a, center), db = vsub(b, center);
	double rsum = r1 + r2;
	double qa = vdot(da, da) - 2.0*vdot(da, db) + vdot(db, db);
	double qb = vdot(da, db) - vdot(da, da);
	double det = qb*qb - qa*(vdot(da, da) - rsum*rsum);
	if(det >= 0.0){
		double t = (-qb - sqrt(det))/(qa);
		if(0.0 <= t && t <= 1.0){
			V2 n = vnormalize(vlerp(da, db, t));
			info.hit = true;
			info.point = vsub(vlerp(a, b, t), vmul(n, r2));
			info.normal = n;
			info.alpha = t;
		}
	}
}

CPB_DEVICE void q_segment_segment(V2 ta, V2 tb, V2 tn, double sr, V2 a, V2 b, double r2, QSeg &info){
	V2 n = tn;
	double d = vdot(vsub(ta, a), n);
	double r = sr + r2;
	V2 flipped_n = (d > 0.0 ? vneg(n) : n);
	V2 seg_offset = vsub(vmul(flipped_n, r), a);
	V2 seg_a = vadd(ta, seg_offset), seg_b = vadd(tb, seg_offset);
	V2 delta = vsub(b, a);
	if(vcross(delta, seg_a)*vcross(delta, seg_b) <= 0.0){
		double d_offset = d + (d > 0.0 ? -r : r);
		double ad = -d_offset;
		double bd = vdot(delta, n) - d_offset;
		if(ad*bd < 0.0){
			double t = ad/(ad - bd);
			info.hit = true;
			info.point = vsub(vlerp(a, b, t), vmul(flipped_n, r2));
			info.normal = flipped_n;
			info.alpha = t;
		}
	} else if(r != 0.0){
		QSeg i1 = {false, b, v2(0, 0), 1.0}, i2 = {false, b, v2(0, 0), 1.0};
		q_circle_segment(ta, sr, a, b, r2, i1);
		q_circle_segment(tb, sr, a, b, r2, i2);
		info = (i1.alpha < i2.alpha ? i1 : i2);
	}
}

CPB_DEVICE void q_segment_poly(const V2 *pv, const V2 *pn, int count, double r, V2 a, V2 b, double r2, QSeg &info){
	double rsum = r + r2;
	for(int i = 0; i < count; i++){
		V2 n = pn[i];
		double an = vdot(a, n);
		double d = an - vdot(pv[i], n) - rsum;
		if(d < 0.0) continue;
		double bn = vdot(b, n);
		double t = d/fmax_cp(an - bn, DBL_MIN);
		if(t < 0.0 || 1.0 < t) continue;
		V2 point = vlerp(a, b, t);
		double dt = vcross(n, point);
		double dtMin = vcross(n, pv[(i - 1 + count)%count]);
		double dtMax = vcross(n, pv[i]);
		if(dtMin <= dt && dt <= dtMax){
			info.hit = true;
			info.point = vsub(vlerp(a, b, t), vmul(n, r2));
			info.normal = n;
			info.alpha = t;
		}
	}
	if(rsum > 0.0){
		for(int i = 0; i < count; i++){
			QSeg c = {false, b, v2(0, 0), 1.0};
			q_circle_segment(pv[i], r, a, b, r2, c);
			if(c.alpha < info.alpha) info = c;
		}
	}
}

// cpShapeSegmentQuery (cpShape.c:237-258): a start point already within `radius` of the shape is a hit at alpha 0
CPB_DEVICE bool q_shape_segment(const DShapes &S, int s, V2 a, V2 b, double radius, QHit &h){
	QSeg info = {false, b, v2(0, 0), 1.0};
	QHit nearest;
	q_shape_point(S, s, a, nearest);
	if(nearest.d <= radius){
		info.hit = true;
		info.alpha = 0.0;
		info.normal = vnormalize(vsub(a, v2(nearest.px, nearest.py)));
	} else {
		int type = S.type[s];
		if(type == CPB200_SHAPE_CIRCLE) q_circle_segment(S.wa[s], S.r[s], a, b, radius, info);
		else if(type == CPB200_SHAPE_SEGMENT) q_segment_segment(S.wa[s], S.wb[s], S.wn[s], S.r[s], a, b, radius, info);
		else q_segment_poly(S.wpv + S.poff[s], S.wpn + S.poff[s], S.pcount[s], S.r[s], a, b, radius, info);
	}
	h.shape = s; h.pad = 0;
	h.px = info.point.x; h.py = info.point.y; h.d = info.alpha; h.gx = info.normal.x; h.gy = info.normal.y;
	return info.hit;
}

// one candidate shape against the query; true = it is a hit
CPB_DEVICE bool q_test(const DShapes &S, const DBodies &B, const QParams &Q, int s, QHit &h){
	if(Q.only_shape >= 0){
		if(s != Q.only_shape) return false;
		if(Q.kind == 0){ q_shape_point(S, s, Q.a, h); return true; }
		return q_shape_segment(S, s, Q.a, Q.b, Q.radius, h);
	}
	if(Q.space >= 0 && B.space[S.body[s]] != Q.space) return false;
	if(q_filter_reject(S, s, Q.filter)) return false;
	if(Q.skip_sensors && S.sensor[s]) return false;
	double4 bb = S.bb[s];
	if(Q.kind == 0){
		// cheap reject: distance < maxDistance implies the AABB reaches into the query's box
		double md = fmax_cp(Q.radius, 0.0);
		if(!(Q.a.x - md <= bb.z && bb.x <= Q.a.x + md && Q.a.y - md <= bb.w && bb.y <= Q.a.y + md)) return false;
		q_shape_point(S, s, Q.a, h);
		return h.d < Q.radius;
	} else if(Q.kind == 1){
		return q_shape_segment(S, s, Q.a, Q.b, Q.radius, h);
	}
	// cpBBIntersects(query, shape->bb) (cpBB.h:59-62)
	h.shape = s; h.pad = 0; h.px = h.py = h.d = h.gx = h.gy = 0.0;
	return (Q.a.x <= bb.z && bb.x <= Q.b.x && Q.a.y <= bb.w && bb.y <= Q.b.y);
}

// all hits, appended in no particular order (the host sorts them by shape index)
__global__ void k_query_all(DShapes S, DBodies B, QParams Q, QHit *out, int cap, int *count)
{
	int rounded = ((S.n + 31)/32)*32;
	for(int s = CPB_TID; s < rounded; s += CPB_NTHREADS){
		QHit h;
		bool hit = (s < S.n) && q_test(S, B, Q, s, h);
		int slot = cpb_warp_append(count, hit);
		if(hit && slot < cap) out[slot] = h;
	}
}

// order-preserving map double -> uint64 (smaller double = smaller key)
CPB_DEVICE unsigned long long q_ordered(double d){
	unsigned long long b = (unsigned long long)__double_as_longlong(d);
	return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// best hit (smallest distance / alpha; ties -> lowest shape index): pass 1 finds the key, pass 2 the shape, pass 3 writes
__global__ void k_query_best(DShapes S, DBodies B, QParams Q, int pass, unsigned long long *best_key, int *best_shape, QHit *out)
{
	for(int s = CPB_TID; s < S.n; s += CPB_NTHREADS){
		if(pass == 2 && s != *best_shape) continue;
		QHit h;
		if(!q_test(S, B, Q, s, h)) continue;
		unsigned long long key = q_ordered(h.d);
		if(pass == 0) atomicMin(best_key, key);
		else if(pass == 1){ if(key == *best_key) atomicMin(best_shape, s); }
		else out[0] = h;
	}
}

// ---- cpSpaceShapeQuery (cpSpaceQuery.c:203-246): a caller-supplied shape against every shape of the space ----
// The query shape need not live in the world: the host passes its world-space description (what cacheData
// would have produced) and, for a polygon, device copies of its vertices / edge normals.
struct QShape {
	int type, count;
	double r;
	V2 a, b, n, rot, atan, btan;
	double4 bb;
	const V2 *pv, *pn;
	int self;            // index of the query shape if it is part of the world (never reported), else -1
};

struct QShapeHit { int shape, count; double nx, ny; double pts[2][5]; };   // pts[k] = pointA.xy, pointB.xy, distance

__global__ void k_query_shape(DShapes S, DBodies B, QShape q, QFilter filter, int space, QShapeHit *out, int cap, int *count)
{
	int rounded = ((S.n + 31)/32)*32;
	for(int s = CPB_TID; s < rounded; s += CPB_NTHREADS){
		bool hit = false;
		QShapeHit h;
		if(s < S.n && s != q.self && (space < 0 || B.space[S.body[s]] == space) && !q_filter_reject(S, s, filter)){
			double4 bb = S.bb[s];
			if(q.bb.x <= bb.z && bb.x <= q.bb.z && q.bb.y <= bb.w && bb.y <= q.bb.w){
				NShape a;
				a.type = q.type; a.count = q.count; a.hashid = 0; a.a = q.a; a.b = q.b; a.n = q.n; a.r = q.r;
				a.bbc = v2((q.bb.x + q.bb.z)*0.5, (q.bb.y + q.bb.w)*0.5);
				a.pv = q.pv; a.pn = q.pn; a.sv = 0; a.rot = q.rot; a.atan = q.atan; a.btan = q.btan;
				NShape b = load_nshape(S, B, s);
				// cpShapesCollide (cpShape.c:259-283): cpCollide wants a.type <= b.type; swap back afterwards
				bool swapped = (a.type > b.type);
				Manifold m; m.id = 0; m.count = 0;
				if(swapped) collide_shapes(b, a, m); else collide_shapes(a, b, m);
				if(m.count > 0){
					hit = true;
					V2 n = swapped ? vneg(m.n) : m.n;
					h.shape = s; h.count = m.count; h.nx = n.x; h.ny = n.y;
					for(int k = 0; k < 2; k++) for(int c = 0; c < 5; c++) h.pts[k][c] = 0.0;
					for(int k = 0; k < m.count; k++){
						V2 p1 = m.p1[k], p2 = m.p2[k];
						V2 A_ = swapped ? p2 : p1, B_ = swapped ? p1 : p2;
						h.pts[k][0] = A_.x; h.pts[k][1] = A_.y; h.pts[k][2] = B_.x; h.pts[k][3] = B_.y;
						h.pts[k][4] = vdot(vsub(p2, p1), n);
					}
				}
			}
		}
		int slot = cpb_warp_append(count, hit);
		if(hit && slot < cap) out[slot] = h;
	}
}
