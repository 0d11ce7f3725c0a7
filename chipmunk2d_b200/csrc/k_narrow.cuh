// k_narrow.cuh -- K5: per-pair narrowphase (device functions).
//
// B200 restatement of cpCollide and its callees (reference src/cpCollision.c:44-726,
// src/cpRobust.c:4-13).  One thread owns one shape pair.  The reference's two
// tail-recursive routines become bounded loops:
//   * GJK  (cpCollision.c:348-392, 416-472): a 1-simplex walk on the Minkowski difference;
//   * EPA  (cpCollision.c:270-343): the expanding hull is kept as an array of 16-bit
//     vertex-index pairs (33 x 2 B per thread) instead of 56-byte MinkowskiPoints --
//     the points are recomputed from the cached world vertices, which is bit-identical
//     to the reference's stored copies and keeps the thread's local memory tiny.
// All predicates keep the reference's association (cpRobust.c must not be re-associated;
// the file is built with -fmad=false).
#pragma once
#include "cpb_math.h"

#define CPB_MAX_GJK_ITERATIONS 30
#define CPB_MAX_EPA_ITERATIONS 30
#define CPB_EPA_HULL_MAX 34

// World-space view of one shape for the narrowphase.
struct NShape {
	int type;            // CPB200_SHAPE_*
	int count;           // poly vertex count
	uint32_t hashid;
	V2 a, b, n;          // circle: a = tc.  segment: a = ta, b = tb, n = tn
	double r;
	V2 bbc;              // cpBBCenter(shape->bb) (cpBB.h:97-100)
	const V2 *pv;        // poly: world vertices  (planes[i].v0)
	const V2 *pn;        // poly: world normals   (planes[i].n)
	const double *sv;    // poly: the same vertices staged in shared memory by k_collide<2> (NULL: read pv), see nshape_vert
	V2 rot;              // owning body's rotation (cpBodyGetRotation)
	V2 atan, btan;       // segment neighbour tangents (body-local)
};

struct Manifold {
	int count;
	V2 n;
	V2 p1[2], p2[2];     // absolute surface points on a / b (cpCollision.c:44-56)
	uint64_t hash[2];
	uint32_t id;         // GJK warm-start id (cpCollisionID)
};

CPB_DEVICE void push_contact(Manifold &m, V2 p1, V2 p2, uint64_t hash){
	m.p1[m.count] = p1; m.p2[m.count] = p2; m.hash[m.count] = hash; m.count++;
}

// ---- robust predicates (cpRobust.c:4-13) ----
CPB_DEVICE bool check_point_greater(V2 a, V2 b, V2 c){
	return (b.y - a.y)*(a.x + b.x - 2*c.x) > (b.x - a.x)*(a.y + b.y - 2*c.y);
}
CPB_DEVICE bool check_axis(V2 v0, V2 v1, V2 p, V2 n){
	return vdot(p, n) <= fmax_cp(vdot(v0, n), vdot(v1, n));
}

// GJK and EPA evaluate support points over and over (cpCollision.c:62-78: a scan of every vertex per call, several
// calls per iteration, for both shapes): k_collide<2> copies the world vertices of the pair's polygons into shared
// memory once -- vertex i of thread t at sv[(2i + c)*CPB_GJK_STAGE_STRIDE], c = 0 (x), 1 (y), so that a warp reading
// "its" vertex i touches consecutive words -- and every later read comes from there.  Polygons with more than
// CPB_GJK_STAGE_VERTS vertices, the one-pair hook and the shape queries keep the global pointer.
#define CPB_GJK_STAGE_VERTS 8
#define CPB_GJK_STAGE_STRIDE 128     // = threads per CTA of k_collide
CPB_DEVICE V2 nshape_vert(const NShape &s, int i){
#ifndef CPB_EMU
	if(s.sv) return v2(s.sv[(2*i)*CPB_GJK_STAGE_STRIDE], s.sv[(2*i + 1)*CPB_GJK_STAGE_STRIDE]);
#endif
	return s.pv[i];
}

// ---- support points (cpCollision.c:62-117) ----
struct SupportPoint { V2 p; uint32_t index; };

CPB_DEVICE int poly_support_index(const NShape &s, V2 n){
	double max = -INFINITY;
	int index = 0;
	for(int i = 0; i < s.count; i++){
		V2 v = nshape_vert(s, i);
		double d = vdot(v, n);
		if(d > max){ max = d; index = i; }
	}
	return index;
}

CPB_DEVICE SupportPoint support_point(const NShape &s, V2 n){
	SupportPoint sp;
	if(s.type == 0){ sp.p = s.a; sp.index = 0; }
	else if(s.type == 1){
		if(vdot(s.a, n) > vdot(s.b, n)){ sp.p = s.a; sp.index = 0; } else { sp.p = s.b; sp.index = 1; }
	} else {
		int i = poly_support_index(s, n);
		sp.p = nshape_vert(s, i); sp.index = (uint32_t)i;
	}
	return sp;
}

// ShapePoint (cpCollision.c:395-413)
CPB_DEVICE SupportPoint shape_point(const NShape &s, int i){
	SupportPoint sp;
	if(s.type == 0){ sp.p = s.a; sp.index = 0; }
	else if(s.type == 1){ sp.p = (i == 0 ? s.a : s.b); sp.index = (uint32_t)i; }
	else { int index = (i < s.count ? i : 0); sp.p = nshape_vert(s, index); sp.index = (uint32_t)index; }
	return sp;
}

// MinkowskiPoint (cpCollision.c:119-134)
struct MPoint { V2 a, b, ab; uint32_t id; };

CPB_DEVICE MPoint mpoint(SupportPoint a, SupportPoint b){
	MPoint m; m.a = a.p; m.b = b.p; m.ab = vsub(b.p, a.p); m.id = (a.index & 0xFF)<<8 | (b.index & 0xFF);
	return m;
}
CPB_DEVICE MPoint support(const NShape &s1, const NShape &s2, V2 n){
	return mpoint(support_point(s1, vneg(n)), support_point(s2, n));
}
// rebuild a Minkowski point from its 16-bit id (see file header)
CPB_DEVICE MPoint mpoint_from_id(const NShape &s1, const NShape &s2, uint32_t id){
	return mpoint(shape_point(s1, (int)((id >> 8) & 0xFF)), shape_point(s2, (int)(id & 0xFF)));
}

// ---- closest points (cpCollision.c:197-266) ----
CPB_DEVICE double closest_t(V2 a, V2 b){
	V2 delta = vsub(b, a);
	return -fclamp_cp(vdot(delta, vadd(a, b))/(vlensq(delta) + DBL_MIN), -1.0, 1.0);
}
CPB_DEVICE V2 lerp_t(V2 a, V2 b, double t){
	double ht = 0.5*t;
	return vadd(vmul(a, 0.5 - ht), vmul(b, 0.5 + ht));
}
struct ClosestPoints { V2 a, b, n; double d; uint32_t id; };

CPB_DEVICE ClosestPoints closest_points_new(const MPoint &v0, const MPoint &v1){
	double t = closest_t(v0.ab, v1.ab);
	V2 p = lerp_t(v0.ab, v1.ab, t);
	V2 pa = lerp_t(v0.a, v1.a, t);
	V2 pb = lerp_t(v0.b, v1.b, t);
	uint32_t id = (v0.id & 0xFFFF)<<16 | (v1.id & 0xFFFF);
	V2 delta = vsub(v1.ab, v0.ab);
	V2 n = vnormalize(vrperp(delta));
	double d = vdot(n, p);
	ClosestPoints out;
	out.a = pa; out.b = pb; out.id = id;
	if(d <= 0.0 || (-1.0 < t && t < 1.0)){
		out.n = n; out.d = d;
	} else {
		double d2 = vlen(p);
		out.n = vmul(p, 1.0/(d2 + DBL_MIN)); out.d = d2;
	}
	return out;
}
CPB_DEVICE double closest_dist(V2 v0, V2 v1){
	return vlensq(lerp_t(v0, v1, closest_t(v0, v1)));
}

// ---- EPA (cpCollision.c:270-343) as a loop over an id-hull ----
CPB_DEVICE ClosestPoints epa(const NShape &s1, const NShape &s2, const MPoint &e0, const MPoint &e1, const MPoint &e2){
	uint16_t hullA[CPB_EPA_HULL_MAX], hullB[CPB_EPA_HULL_MAX];
	uint16_t *hull = hullA, *hull2 = hullB;
	hull[0] = (uint16_t)e0.id; hull[1] = (uint16_t)e1.id; hull[2] = (uint16_t)e2.id;
	int count = 3;
	for(int iteration = 1; ; iteration++){
		int mini = 0;
		double minDist = INFINITY;
		{
			V2 prev = mpoint_from_id(s1, s2, hull[count - 1]).ab;
			for(int j = 0, i = count - 1; j < count; i = j, j++){
				V2 cur = mpoint_from_id(s1, s2, hull[j]).ab;
				double d = closest_dist(prev, cur);
				if(d < minDist){ minDist = d; mini = i; }
				prev = cur;
			}
		}
		MPoint v0 = mpoint_from_id(s1, s2, hull[mini]);
		MPoint v1 = mpoint_from_id(s1, s2, hull[(mini + 1)%count]);
		MPoint p = support(s1, s2, vperp(vsub(v1.ab, v0.ab)));
		bool duplicate = (p.id == v0.id || p.id == v1.id);
		if(!duplicate && check_point_greater(v0.ab, v1.ab, p.ab) && iteration < CPB_MAX_EPA_ITERATIONS){
			int count2 = 1;
			hull2[0] = (uint16_t)p.id;
			V2 h0 = p.ab;
			for(int i = 0; i < count; i++){
				int index = (mini + 1 + i)%count;
				V2 h1 = mpoint_from_id(s1, s2, hull[index]).ab;
				V2 h2 = (i + 1 < count ? mpoint_from_id(s1, s2, hull[(index + 1)%count]).ab : p.ab);
				if(check_point_greater(h0, h2, h1)){
					hull2[count2] = hull[index];
					count2++;
					h0 = h1;
				}
			}
			uint16_t *tmp = hull; hull = hull2; hull2 = tmp;
			count = count2;
		} else {
			return closest_points_new(v0, v1);
		}
	}
}

// ---- GJK (cpCollision.c:348-392, 416-472) ----
CPB_DEVICE ClosestPoints gjk(const NShape &s1, const NShape &s2, uint32_t *id){
	MPoint v0, v1;
	if(*id){
		v0 = mpoint(shape_point(s1, (int)((*id >> 24) & 0xFF)), shape_point(s2, (int)((*id >> 16) & 0xFF)));
		v1 = mpoint(shape_point(s1, (int)((*id >>  8) & 0xFF)), shape_point(s2, (int)((*id      ) & 0xFF)));
	} else {
		V2 axis = vperp(vsub(s1.bbc, s2.bbc));
		v0 = support(s1, s2, axis);
		v1 = support(s1, s2, vneg(axis));
	}
	ClosestPoints pts;
	int iteration = 1;
	for(;;){
		if(iteration > CPB_MAX_GJK_ITERATIONS){ pts = closest_points_new(v0, v1); break; }
		if(check_point_greater(v1.ab, v0.ab, v2(0.0, 0.0))){
			// origin is behind the axis: flip, same iteration (cpCollision.c:356-358)
			MPoint t = v0; v0 = v1; v1 = t;
			continue;
		}
		double t = closest_t(v0.ab, v1.ab);
		V2 n = (-1.0 < t && t < 1.0 ? vperp(vsub(v1.ab, v0.ab)) : vneg(lerp_t(v0.ab, v1.ab, t)));
		MPoint p = support(s1, s2, n);
		if(check_point_greater(p.ab, v0.ab, v2(0.0, 0.0)) && check_point_greater(v1.ab, p.ab, v2(0.0, 0.0))){
			pts = epa(s1, s2, v0, p, v1);
			break;
		}
		if(check_axis(v0.ab, v1.ab, p.ab, n)){ pts = closest_points_new(v0, v1); break; }
		if(closest_dist(v0.ab, p.ab) < closest_dist(p.ab, v1.ab)){ v1 = p; } else { v0 = p; }
		iteration++;
	}
	*id = pts.id;
	return pts;
}

// ---- support edges + clipping (cpCollision.c:150-195, 477-518) ----
struct EdgePoint { V2 p; uint64_t hash; };
struct Edge { EdgePoint a, b; double r; V2 n; };

CPB_DEVICE Edge support_edge_poly(const NShape &s, V2 n){
	int count = s.count;
	int i1 = poly_support_index(s, n);
	int i0 = (i1 - 1 + count)%count;
	int i2 = (i1 + 1)%count;
	uint64_t hashid = s.hashid;
	Edge e;
	if(vdot(n, s.pn[i1]) > vdot(n, s.pn[i2])){
		e.a.p = nshape_vert(s, i0); e.a.hash = hash_pair(hashid, (uint64_t)i0);
		e.b.p = nshape_vert(s, i1); e.b.hash = hash_pair(hashid, (uint64_t)i1);
		e.r = s.r; e.n = s.pn[i1];
	} else {
		e.a.p = nshape_vert(s, i1); e.a.hash = hash_pair(hashid, (uint64_t)i1);
		e.b.p = nshape_vert(s, i2); e.b.hash = hash_pair(hashid, (uint64_t)i2);
		e.r = s.r; e.n = s.pn[i2];
	}
	return e;
}
CPB_DEVICE Edge support_edge_segment(const NShape &s, V2 n){
	uint64_t hashid = s.hashid;
	Edge e;
	if(vdot(s.n, n) > 0.0){
		e.a.p = s.a; e.a.hash = hash_pair(hashid, 0);
		e.b.p = s.b; e.b.hash = hash_pair(hashid, 1);
		e.r = s.r; e.n = s.n;
	} else {
		e.a.p = s.b; e.a.hash = hash_pair(hashid, 1);
		e.b.p = s.a; e.b.hash = hash_pair(hashid, 0);
		e.r = s.r; e.n = vneg(s.n);
	}
	return e;
}

CPB_DEVICE void contact_points(const Edge &e1, const Edge &e2, const ClosestPoints &points, Manifold &m){
	double mindist = e1.r + e2.r;
	if(points.d <= mindist){
		V2 n = m.n = points.n;
		double d_e1_a = vcross(e1.a.p, n);
		double d_e1_b = vcross(e1.b.p, n);
		double d_e2_a = vcross(e2.a.p, n);
		double d_e2_b = vcross(e2.b.p, n);
		double e1_denom = 1.0/(d_e1_b - d_e1_a + DBL_MIN);
		double e2_denom = 1.0/(d_e2_b - d_e2_a + DBL_MIN);
		{
			V2 p1 = vadd(vmul(n,  e1.r), vlerp(e1.a.p, e1.b.p, fclamp01_cp((d_e2_b - d_e1_a)*e1_denom)));
			V2 p2 = vadd(vmul(n, -e2.r), vlerp(e2.a.p, e2.b.p, fclamp01_cp((d_e1_a - d_e2_a)*e2_denom)));
			double dist = vdot(vsub(p2, p1), n);
			if(dist <= 0.0) push_contact(m, p1, p2, hash_pair(e1.a.hash, e2.b.hash));
		}{
			V2 p1 = vadd(vmul(n,  e1.r), vlerp(e1.a.p, e1.b.p, fclamp01_cp((d_e2_a - d_e1_a)*e1_denom)));
			V2 p2 = vadd(vmul(n, -e2.r), vlerp(e2.a.p, e2.b.p, fclamp01_cp((d_e1_b - d_e2_a)*e2_denom)));
			double dist = vdot(vsub(p2, p1), n);
			if(dist <= 0.0) push_contact(m, p1, p2, hash_pair(e1.b.hash, e2.a.hash));
		}
	}
}

// ---- the six collision functions (cpCollision.c:525-679) ----
CPB_DEVICE void circle_to_circle(const NShape &c1, const NShape &c2, Manifold &m){
	double mindist = c1.r + c2.r;
	V2 delta = vsub(c2.a, c1.a);
	double distsq = vlensq(delta);
	if(distsq < mindist*mindist){
		double dist = sqrt(distsq);
		V2 n = m.n = (dist ? vmul(delta, 1.0/dist) : v2(1.0, 0.0));
		push_contact(m, vadd(c1.a, vmul(n, c1.r)), vadd(c2.a, vmul(n, -c2.r)), 0);
	}
}

CPB_DEVICE void circle_to_segment(const NShape &circle, const NShape &seg, Manifold &m){
	V2 seg_a = seg.a, seg_b = seg.b, center = circle.a;
	V2 seg_delta = vsub(seg_b, seg_a);
	double closest_t_ = fclamp01_cp(vdot(seg_delta, vsub(center, seg_a))/vlensq(seg_delta));
	V2 closest = vadd(seg_a, vmul(seg_delta, closest_t_));
	double mindist = circle.r + seg.r;
	V2 delta = vsub(closest, center);
	double distsq = vlensq(delta);
	if(distsq < mindist*mindist){
		double dist = sqrt(distsq);
		V2 n = m.n = (dist ? vmul(delta, 1.0/dist) : seg.n);
		V2 rot = seg.rot;
		if(
			(closest_t_ != 0.0 || vdot(n, vrotate(seg.atan, rot)) >= 0.0) &&
			(closest_t_ != 1.0 || vdot(n, vrotate(seg.btan, rot)) >= 0.0)
		){
			push_contact(m, vadd(center, vmul(n, circle.r)), vadd(closest, vmul(n, -seg.r)), 0);
		}
	}
}

CPB_DEVICE void segment_to_segment(const NShape &seg1, const NShape &seg2, Manifold &m){
	ClosestPoints points = gjk(seg1, seg2, &m.id);
	V2 n = points.n;
	V2 rot1 = seg1.rot, rot2 = seg2.rot;
	if(
		points.d <= (seg1.r + seg2.r) && (
			(!veql(points.a, seg1.a) || vdot(n, vrotate(seg1.atan, rot1)) <= 0.0) &&
			(!veql(points.a, seg1.b) || vdot(n, vrotate(seg1.btan, rot1)) <= 0.0) &&
			(!veql(points.b, seg2.a) || vdot(n, vrotate(seg2.atan, rot2)) >= 0.0) &&
			(!veql(points.b, seg2.b) || vdot(n, vrotate(seg2.btan, rot2)) >= 0.0)
		)
	){
		contact_points(support_edge_segment(seg1, n), support_edge_segment(seg2, vneg(n)), points, m);
	}
}

CPB_DEVICE void poly_to_poly(const NShape &p1, const NShape &p2, Manifold &m){
	ClosestPoints points = gjk(p1, p2, &m.id);
	if(points.d - p1.r - p2.r <= 0.0){
		contact_points(support_edge_poly(p1, points.n), support_edge_poly(p2, vneg(points.n)), points, m);
	}
}

CPB_DEVICE void segment_to_poly(const NShape &seg, const NShape &poly, Manifold &m){
	ClosestPoints points = gjk(seg, poly, &m.id);
	V2 n = points.n;
	V2 rot = seg.rot;
	if(
		points.d - seg.r - poly.r <= 0.0 && (
			(!veql(points.a, seg.a) || vdot(n, vrotate(seg.atan, rot)) <= 0.0) &&
			(!veql(points.a, seg.b) || vdot(n, vrotate(seg.btan, rot)) <= 0.0)
		)
	){
		contact_points(support_edge_segment(seg, n), support_edge_poly(poly, vneg(n)), points, m);
	}
}

CPB_DEVICE void circle_to_poly(const NShape &circle, const NShape &poly, Manifold &m){
	ClosestPoints points = gjk(circle, poly, &m.id);
	if(points.d <= circle.r + poly.r){
		V2 n = m.n = points.n;
		push_contact(m, vadd(points.a, vmul(n, circle.r)), vadd(points.b, vmul(n, -poly.r)), 0);
	}
}

// cpCollide dispatch (cpCollision.c:688-726).  Requires a.type <= b.type.
CPB_DEVICE void collide_shapes(const NShape &a, const NShape &b, Manifold &m){
	m.count = 0; m.n = v2(0.0, 0.0);
	int code = a.type + b.type*3;
	switch(code){
		case 0: circle_to_circle(a, b, m); break;
		case 3: circle_to_segment(a, b, m); break;
		case 4: segment_to_segment(a, b, m); break;
		case 6: circle_to_poly(a, b, m); break;
		case 7: segment_to_poly(a, b, m); break;
		case 8: poly_to_poly(a, b, m); break;
		default: break;
	}
}
