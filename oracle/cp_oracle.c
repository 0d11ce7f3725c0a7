/* cp_oracle.c -- TEST INFRASTRUCTURE: plain-C restatement of the cpSpaceStep stage algorithms
 * (see cp_oracle.h).  Not part of the product; never linked or imported by chipmunk2d_b200/.
 * Pinned bit-for-bit against the unmodified reference by tests/test_cpu_oracle_restatement.py. */
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdlib.h>
#include "cp_oracle.h"

int cp_oracle_version(void){ return 3; }

/* ---- vector helpers with the reference's association (cpVect.h:58-206) ---- */
static cpo_vec V(double x, double y){ cpo_vec v = {x, y}; return v; }
static cpo_vec add(cpo_vec a, cpo_vec b){ return V(a.x + b.x, a.y + b.y); }
static cpo_vec sub(cpo_vec a, cpo_vec b){ return V(a.x - b.x, a.y - b.y); }
static cpo_vec neg(cpo_vec a){ return V(-a.x, -a.y); }
static cpo_vec mul(cpo_vec a, double s){ return V(a.x*s, a.y*s); }
static double dot(cpo_vec a, cpo_vec b){ return a.x*b.x + a.y*b.y; }
static double cross(cpo_vec a, cpo_vec b){ return a.x*b.y - a.y*b.x; }
static cpo_vec perp(cpo_vec a){ return V(-a.y, a.x); }
static cpo_vec rperp(cpo_vec a){ return V(a.y, -a.x); }
static cpo_vec rotate(cpo_vec a, cpo_vec b){ return V(a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x); }
static double lensq(cpo_vec a){ return dot(a, a); }
static double len(cpo_vec a){ return sqrt(dot(a, a)); }
static cpo_vec lerp(cpo_vec a, cpo_vec b, double t){ return add(mul(a, 1.0 - t), mul(b, t)); }
static cpo_vec normalize(cpo_vec a){ return mul(a, 1.0/(len(a) + DBL_MIN)); }
static int eql(cpo_vec a, cpo_vec b){ return a.x == b.x && a.y == b.y; }
static double fmax_(double a, double b){ return (a > b) ? a : b; }   /* chipmunk_types.h:119-146 */
static double fmin_(double a, double b){ return (a < b) ? a : b; }
static double fabs_(double a){ return (a < 0) ? -a : a; }
static double clamp(double f, double lo, double hi){ return fmin_(fmax_(f, lo), hi); }
static double clamp01(double f){ return fmax_(0.0, fmin_(f, 1.0)); }
static cpo_vec vclamp(cpo_vec v, double l){ return (dot(v, v) > l*l) ? mul(normalize(v), l) : v; }
/* cpTransform {a b c d tx ty}: cpTransformPoint / cpTransformVect (cpTransform.h:73-84) */
static cpo_vec tpoint(const double T[6], cpo_vec p){ return V(T[0]*p.x + T[2]*p.y + T[4], T[1]*p.x + T[3]*p.y + T[5]); }
static cpo_vec tvect(const double T[6], cpo_vec v){ return V(T[0]*v.x + T[2]*v.y, T[1]*v.x + T[3]*v.y); }
#define HASH_COEF 3344921057ull
static uint64_t hash_pair(uint64_t a, uint64_t b){ return (a*HASH_COEF) ^ (b*HASH_COEF); } /* chipmunk_private.h:28-29 */

/* ---- K1: cpBodyUpdatePosition + SetTransform (cpBody.c:511-522, 347-357) ---- */
void cpo_body_update_position(cpo_body *b, double dt, double T[6])
{
	cpo_vec p = b->p = add(b->p, mul(add(b->v, b->v_bias), dt));
	double a = b->a = b->a + (b->w + b->w_bias)*dt;
	cpo_vec rot = V(cos(a), sin(a)), c = b->cog;
	T[0] = rot.x; T[1] = rot.y; T[2] = -rot.y; T[3] = rot.x;
	T[4] = p.x - (c.x*rot.x - c.y*rot.y);
	T[5] = p.y - (c.x*rot.y + c.y*rot.x);
	b->v_bias = V(0, 0);
	b->w_bias = 0.0;
}

/* ---- K9: cpBodyUpdateVelocity (cpBody.c:493-509) ---- */
void cpo_body_update_velocity(cpo_body *b, cpo_vec gravity, double damping, double dt)
{
	if(b->type == 1) return;
	b->v = add(mul(b->v, damping), mul(add(gravity, mul(b->f, b->m_inv)), dt));
	b->w = b->w*damping + b->t*b->i_inv*dt;
	b->f = V(0, 0);
	b->t = 0.0;
}

/* ---- K2: cacheData (cpShape.c:291-296, 378-405; cpPolyShape.c:39-64) ---- */
void cpo_cache_circle(cpo_vec c, double r, const double T[6], cpo_shape *out)
{
	cpo_vec tc = tpoint(T, c);
	out->type = 0; out->a = tc; out->r = r;
	out->bb[0] = tc.x - r; out->bb[1] = tc.y - r; out->bb[2] = tc.x + r; out->bb[3] = tc.y + r;
}

void cpo_cache_segment(cpo_vec a, cpo_vec b, cpo_vec n, double rad, const double T[6], cpo_shape *out)
{
	cpo_vec ta = tpoint(T, a), tb = tpoint(T, b);
	out->type = 1; out->a = ta; out->b = tb; out->n = tvect(T, n); out->r = rad;
	double l, r, bt, t;
	if(ta.x < tb.x){ l = ta.x; r = tb.x; } else { l = tb.x; r = ta.x; }
	if(ta.y < tb.y){ bt = ta.y; t = tb.y; } else { bt = tb.y; t = ta.y; }
	out->bb[0] = l - rad; out->bb[1] = bt - rad; out->bb[2] = r + rad; out->bb[3] = t + rad;
}

void cpo_cache_poly(int count, const double *src, double radius, const double T[6], double *dst, cpo_shape *out)
{
	double l = INFINITY, r = -INFINITY, b = INFINITY, t = -INFINITY;
	for(int i = 0; i < count; i++){
		cpo_vec v = tpoint(T, V(src[4*i], src[4*i + 1]));
		cpo_vec n = tvect(T, V(src[4*i + 2], src[4*i + 3]));
		dst[4*i] = v.x; dst[4*i + 1] = v.y; dst[4*i + 2] = n.x; dst[4*i + 3] = n.y;
		l = fmin_(l, v.x); r = fmax_(r, v.x); b = fmin_(b, v.y); t = fmax_(t, v.y);
	}
	out->type = 2; out->count = count; out->planes = dst; out->r = radius;
	out->bb[0] = l - radius; out->bb[1] = b - radius; out->bb[2] = r + radius; out->bb[3] = t + radius;
}

/* ---- K3 + K4: the overlapping-pair set by brute force (QueryReject, cpSpaceStep.c:204-232;
 * cpBBIntersects cpBB.h:58-61; cpShapeFilterReject chipmunk_private.h:144-155).  A ranges over
 * shapes of awake non-static bodies, B over all shapes. ---- */
static int cmp_u64(const void *a, const void *b){ uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return (x < y ? -1 : (x > y)); }

long cpo_pairs(int n, const double *bb, const int *body, const int *body_active, const uint64_t *group, const uint32_t *cat, const uint32_t *mask,
	int n_nc, const uint64_t *nc, long cap, uint64_t *out)
{
	long count = 0;
	for(int i = 0; i < n; i++){
		for(int j = i + 1; j < n; j++){
			if(!body_active[body[i]] && !body_active[body[j]]) continue;
			const double *a = bb + 4*i, *b = bb + 4*j;
			if(!(a[0] <= b[2] && b[0] <= a[2] && a[1] <= b[3] && b[1] <= a[3])) continue;
			if(body[i] == body[j]) continue;
			if(group[i] != 0 && group[i] == group[j]) continue;
			if((cat[i] & mask[j]) == 0 || (cat[j] & mask[i]) == 0) continue;
			uint64_t lo = (uint64_t)(body[i] < body[j] ? body[i] : body[j]), hi = (uint64_t)(body[i] < body[j] ? body[j] : body[i]);
			uint64_t key = (lo << 32) | hi;
			int rejected = 0;
			for(int k = 0; k < n_nc; k++){ if(nc[k] == key){ rejected = 1; break; } }
			if(rejected) continue;
			if(count < cap) out[count] = ((uint64_t)i << 32) | (uint64_t)j;
			count++;
		}
	}
	if(count <= cap) qsort(out, (size_t)count, sizeof(uint64_t), cmp_u64);
	return count;
}

/* ---- K5: narrowphase (cpCollision.c) ---- */
typedef struct support_point { cpo_vec p; uint32_t index; } support_point;
typedef struct mink_point { cpo_vec a, b, ab; uint32_t id; } mink_point;
typedef struct closest_points { cpo_vec a, b, n; double d; uint32_t id; } closest_points;

static cpo_vec plane_v(const cpo_shape *s, int i){ return V(s->planes[4*i], s->planes[4*i + 1]); }
static cpo_vec plane_n(const cpo_shape *s, int i){ return V(s->planes[4*i + 2], s->planes[4*i + 3]); }

/* PolySupportPointIndex (cpCollision.c:62-78): first maximum wins */
static int poly_support_index(const cpo_shape *s, cpo_vec n)
{
	double max = -INFINITY; int index = 0;
	for(int i = 0; i < s->count; i++){ double d = dot(plane_v(s, i), n); if(d > max){ max = d; index = i; } }
	return index;
}

/* Circle/Segment/PolySupportPoint (cpCollision.c:95-117) */
static support_point support_of(const cpo_shape *s, cpo_vec n)
{
	support_point sp;
	if(s->type == 0){ sp.p = s->a; sp.index = 0; }
	else if(s->type == 1){ if(dot(s->a, n) > dot(s->b, n)){ sp.p = s->a; sp.index = 0; } else { sp.p = s->b; sp.index = 1; } }
	else { int i = poly_support_index(s, n); sp.p = plane_v(s, i); sp.index = (uint32_t)i; }
	return sp;
}

/* ShapePoint (cpCollision.c:395-413) */
static support_point shape_point(const cpo_shape *s, int i)
{
	support_point sp;
	if(s->type == 0){ sp.p = s->a; sp.index = 0; }
	else if(s->type == 1){ sp.p = (i == 0 ? s->a : s->b); sp.index = (uint32_t)i; }
	else { int k = (i < s->count ? i : 0); sp.p = plane_v(s, k); sp.index = (uint32_t)k; }
	return sp;
}

/* MinkowskiPointNew / Support (cpCollision.c:119-148) */
static mink_point mink(support_point a, support_point b){ mink_point m = {a.p, b.p, sub(b.p, a.p), (a.index & 0xFF)<<8 | (b.index & 0xFF)}; return m; }
static mink_point support(const cpo_shape *s1, const cpo_shape *s2, cpo_vec n){ return mink(support_of(s1, neg(n)), support_of(s2, n)); }

/* cpCheckPointGreater / cpCheckAxis (cpRobust.c:4-13) */
static int point_greater(cpo_vec a, cpo_vec b, cpo_vec c){ return (b.y - a.y)*(a.x + b.x - 2*c.x) > (b.x - a.x)*(a.y + b.y - 2*c.y); }
static int check_axis(cpo_vec v0, cpo_vec v1, cpo_vec p, cpo_vec n){ return dot(p, n) <= fmax_(dot(v0, n), dot(v1, n)); }

/* ClosestT / LerpT / ClosestPointsNew / ClosestDist (cpCollision.c:197-266) */
static double closest_t(cpo_vec a, cpo_vec b){ cpo_vec d = sub(b, a); return -clamp(dot(d, add(a, b))/(lensq(d) + DBL_MIN), -1.0, 1.0); }
static cpo_vec lerp_t(cpo_vec a, cpo_vec b, double t){ double ht = 0.5*t; return add(mul(a, 0.5 - ht), mul(b, 0.5 + ht)); }
static double closest_dist(cpo_vec v0, cpo_vec v1){ return lensq(lerp_t(v0, v1, closest_t(v0, v1))); }

static closest_points closest_points_new(mink_point v0, mink_point v1)
{
	double t = closest_t(v0.ab, v1.ab);
	cpo_vec p = lerp_t(v0.ab, v1.ab, t);
	cpo_vec pa = lerp_t(v0.a, v1.a, t), pb = lerp_t(v0.b, v1.b, t);
	uint32_t id = (v0.id & 0xFFFF)<<16 | (v1.id & 0xFFFF);
	cpo_vec n = normalize(rperp(sub(v1.ab, v0.ab)));
	double d = dot(n, p);
	closest_points out = {pa, pb, n, d, id};
	if(!(d <= 0.0 || (-1.0 < t && t < 1.0))){
		double d2 = len(p);
		out.n = mul(p, 1.0/(d2 + DBL_MIN)); out.d = d2;
	}
	return out;
}

/* EPARecurse / EPA (cpCollision.c:270-343), the tail recursion unrolled into a loop over two hull buffers */
static closest_points epa(const cpo_shape *s1, const cpo_shape *s2, mink_point e0, mink_point e1, mink_point e2)
{
	mink_point bufA[40], bufB[40];
	mink_point *hull = bufA, *hull2 = bufB;
	hull[0] = e0; hull[1] = e1; hull[2] = e2;
	int count = 3;
	for(int iteration = 1; ; iteration++){
		int mini = 0; double minDist = INFINITY;
		for(int j = 0, i = count - 1; j < count; i = j, j++){
			double d = closest_dist(hull[i].ab, hull[j].ab);
			if(d < minDist){ minDist = d; mini = i; }
		}
		mink_point v0 = hull[mini], v1 = hull[(mini + 1)%count];
		mink_point p = support(s1, s2, perp(sub(v1.ab, v0.ab)));
		int duplicate = (p.id == v0.id || p.id == v1.id);
		if(!duplicate && point_greater(v0.ab, v1.ab, p.ab) && iteration < 30){
			int count2 = 1;
			hull2[0] = p;
			for(int i = 0; i < count; i++){
				int index = (mini + 1 + i)%count;
				cpo_vec h0 = hull2[count2 - 1].ab, h1 = hull[index].ab;
				cpo_vec h2 = (i + 1 < count ? hull[(index + 1)%count] : p).ab;
				if(point_greater(h0, h2, h1)) hull2[count2++] = hull[index];
			}
			mink_point *tmp = hull; hull = hull2; hull2 = tmp;
			count = count2;
		} else {
			return closest_points_new(v0, v1);
		}
	}
}

/* GJKRecurse / GJK (cpCollision.c:348-392, 416-472) */
static closest_points gjk(const cpo_shape *s1, const cpo_shape *s2, uint32_t *id)
{
	mink_point v0, v1;
	if(*id){
		v0 = mink(shape_point(s1, (int)((*id>>24)&0xFF)), shape_point(s2, (int)((*id>>16)&0xFF)));
		v1 = mink(shape_point(s1, (int)((*id>> 8)&0xFF)), shape_point(s2, (int)((*id    )&0xFF)));
	} else {
		cpo_vec c1 = lerp(V(s1->bb[0], s1->bb[1]), V(s1->bb[2], s1->bb[3]), 0.5);
		cpo_vec c2 = lerp(V(s2->bb[0], s2->bb[1]), V(s2->bb[2], s2->bb[3]), 0.5);
		cpo_vec axis = perp(sub(c1, c2));
		v0 = support(s1, s2, axis);
		v1 = support(s1, s2, neg(axis));
	}
	closest_points pts;
	for(int iteration = 1; ; ){
		if(iteration > 30){ pts = closest_points_new(v0, v1); break; }
		if(point_greater(v1.ab, v0.ab, V(0, 0))){ mink_point t = v0; v0 = v1; v1 = t; continue; }
		double t = closest_t(v0.ab, v1.ab);
		cpo_vec n = (-1.0 < t && t < 1.0 ? perp(sub(v1.ab, v0.ab)) : neg(lerp_t(v0.ab, v1.ab, t)));
		mink_point p = support(s1, s2, n);
		if(point_greater(p.ab, v0.ab, V(0, 0)) && point_greater(v1.ab, p.ab, V(0, 0))){ pts = epa(s1, s2, v0, p, v1); break; }
		if(check_axis(v0.ab, v1.ab, p.ab, n)){ pts = closest_points_new(v0, v1); break; }
		if(closest_dist(v0.ab, p.ab) < closest_dist(p.ab, v1.ab)) v1 = p; else v0 = p;
		iteration++;
	}
	*id = pts.id;
	return pts;
}

/* SupportEdgeForPoly / SupportEdgeForSegment / ContactPoints (cpCollision.c:150-195, 477-518) */
typedef struct edge { cpo_vec ap, bp; uint64_t ah, bh; double r; cpo_vec n; } edge;

static edge edge_for_poly(const cpo_shape *s, cpo_vec n)
{
	int count = s->count, i1 = poly_support_index(s, n);
	int i0 = (i1 - 1 + count)%count, i2 = (i1 + 1)%count;
	edge e;
	if(dot(n, plane_n(s, i1)) > dot(n, plane_n(s, i2))){
		e.ap = plane_v(s, i0); e.ah = hash_pair(s->hashid, (uint64_t)i0); e.bp = plane_v(s, i1); e.bh = hash_pair(s->hashid, (uint64_t)i1); e.n = plane_n(s, i1);
	} else {
		e.ap = plane_v(s, i1); e.ah = hash_pair(s->hashid, (uint64_t)i1); e.bp = plane_v(s, i2); e.bh = hash_pair(s->hashid, (uint64_t)i2); e.n = plane_n(s, i2);
	}
	e.r = s->r;
	return e;
}

static edge edge_for_segment(const cpo_shape *s, cpo_vec n)
{
	edge e;
	if(dot(s->n, n) > 0.0){ e.ap = s->a; e.ah = hash_pair(s->hashid, 0); e.bp = s->b; e.bh = hash_pair(s->hashid, 1); e.n = s->n; }
	else { e.ap = s->b; e.ah = hash_pair(s->hashid, 1); e.bp = s->a; e.bh = hash_pair(s->hashid, 0); e.n = neg(s->n); }
	e.r = s->r;
	return e;
}

static void push(cpo_manifold *m, cpo_vec p1, cpo_vec p2, uint64_t hash){ m->p1[m->count] = p1; m->p2[m->count] = p2; m->hash[m->count] = hash; m->count++; }

static void contact_points(edge e1, edge e2, closest_points pts, cpo_manifold *m)
{
	if(pts.d <= e1.r + e2.r){
		cpo_vec n = m->n = pts.n;
		double d_e1_a = cross(e1.ap, n), d_e1_b = cross(e1.bp, n), d_e2_a = cross(e2.ap, n), d_e2_b = cross(e2.bp, n);
		double e1_denom = 1.0/(d_e1_b - d_e1_a + DBL_MIN), e2_denom = 1.0/(d_e2_b - d_e2_a + DBL_MIN);
		{
			cpo_vec p1 = add(mul(n,  e1.r), lerp(e1.ap, e1.bp, clamp01((d_e2_b - d_e1_a)*e1_denom)));
			cpo_vec p2 = add(mul(n, -e2.r), lerp(e2.ap, e2.bp, clamp01((d_e1_a - d_e2_a)*e2_denom)));
			if(dot(sub(p2, p1), n) <= 0.0) push(m, p1, p2, hash_pair(e1.ah, e2.bh));
		}{
			cpo_vec p1 = add(mul(n,  e1.r), lerp(e1.ap, e1.bp, clamp01((d_e2_a - d_e1_a)*e1_denom)));
			cpo_vec p2 = add(mul(n, -e2.r), lerp(e2.ap, e2.bp, clamp01((d_e1_b - d_e2_a)*e2_denom)));
			if(dot(sub(p2, p1), n) <= 0.0) push(m, p1, p2, hash_pair(e1.bh, e2.ah));
		}
	}
}

/* cpCollide and the six collision functions (cpCollision.c:525-726); requires a->type <= b->type */
void cpo_collide(const cpo_shape *a, const cpo_shape *b, cpo_manifold *m)
{
	m->count = 0; m->n = V(0, 0);
	int code = a->type + 3*b->type;
	if(code == 0){                                   /* CircleToCircle :525-537 */
		double mindist = a->r + b->r;
		cpo_vec delta = sub(b->a, a->a);
		double distsq = lensq(delta);
		if(distsq < mindist*mindist){
			double dist = sqrt(distsq);
			cpo_vec n = m->n = (dist ? mul(delta, 1.0/dist) : V(1.0, 0.0));
			push(m, add(a->a, mul(n, a->r)), add(b->a, mul(n, -b->r)), 0);
		}
	} else if(code == 3){                            /* CircleToSegment :539-569 */
		cpo_vec seg_delta = sub(b->b, b->a);
		double ct = clamp01(dot(seg_delta, sub(a->a, b->a))/lensq(seg_delta));
		cpo_vec closest = add(b->a, mul(seg_delta, ct));
		double mindist = a->r + b->r;
		cpo_vec delta = sub(closest, a->a);
		double distsq = lensq(delta);
		if(distsq < mindist*mindist){
			double dist = sqrt(distsq);
			cpo_vec n = m->n = (dist ? mul(delta, 1.0/dist) : b->n);
			if((ct != 0.0 || dot(n, rotate(b->a_tangent, b->rot)) >= 0.0) && (ct != 1.0 || dot(n, rotate(b->b_tangent, b->rot)) >= 0.0)){
				push(m, add(a->a, mul(n, a->r)), add(closest, mul(n, -b->r)), 0);
			}
		}
	} else if(code == 4){                            /* SegmentToSegment :571-604 */
		closest_points pts = gjk(a, b, &m->id);
		cpo_vec n = pts.n;
		if(pts.d <= (a->r + b->r) &&
			(!eql(pts.a, a->a) || dot(n, rotate(a->a_tangent, a->rot)) <= 0.0) && (!eql(pts.a, a->b) || dot(n, rotate(a->b_tangent, a->rot)) <= 0.0) &&
			(!eql(pts.b, b->a) || dot(n, rotate(b->a_tangent, b->rot)) >= 0.0) && (!eql(pts.b, b->b) || dot(n, rotate(b->b_tangent, b->rot)) >= 0.0)){
			contact_points(edge_for_segment(a, n), edge_for_segment(b, neg(n)), pts, m);
		}
	} else if(code == 6){                            /* CircleToPoly :661-679 */
		closest_points pts = gjk(a, b, &m->id);
		if(pts.d <= a->r + b->r){
			cpo_vec n = m->n = pts.n;
			push(m, add(pts.a, mul(n, a->r)), add(pts.b, mul(n, -b->r)), 0);
		}
	} else if(code == 7){                            /* SegmentToPoly :629-659 */
		closest_points pts = gjk(a, b, &m->id);
		cpo_vec n = pts.n;
		if(pts.d - a->r - b->r <= 0.0 &&
			(!eql(pts.a, a->a) || dot(n, rotate(a->a_tangent, a->rot)) <= 0.0) && (!eql(pts.a, a->b) || dot(n, rotate(a->b_tangent, a->rot)) <= 0.0)){
			contact_points(edge_for_segment(a, n), edge_for_poly(b, neg(n)), pts, m);
		}
	} else if(code == 8){                            /* PolyToPoly :606-627 */
		closest_points pts = gjk(a, b, &m->id);
		if(pts.d - a->r - b->r <= 0.0) contact_points(edge_for_poly(a, pts.n), edge_for_poly(b, neg(pts.n)), pts, m);
	}
}

/* ---- solver helpers (chipmunk_private.h:172-268) ---- */
static cpo_vec relative_velocity(const cpo_body *a, const cpo_body *b, cpo_vec r1, cpo_vec r2)
{
	cpo_vec v1 = add(a->v, mul(perp(r1), a->w)), v2 = add(b->v, mul(perp(r2), b->w));
	return sub(v2, v1);
}
static void apply_impulse(cpo_body *b, cpo_vec j, cpo_vec r){ b->v = add(b->v, mul(j, b->m_inv)); b->w += b->i_inv*cross(r, j); }
static void apply_impulses(cpo_body *a, cpo_body *b, cpo_vec r1, cpo_vec r2, cpo_vec j){ apply_impulse(a, neg(j), r1); apply_impulse(b, j, r2); }
static void apply_bias_impulse(cpo_body *b, cpo_vec j, cpo_vec r){ b->v_bias = add(b->v_bias, mul(j, b->m_inv)); b->w_bias += b->i_inv*cross(r, j); }
static double k_scalar_body(const cpo_body *b, cpo_vec r, cpo_vec n){ double rcn = cross(r, n); return b->m_inv + b->i_inv*rcn*rcn; }
static double k_scalar(const cpo_body *a, const cpo_body *b, cpo_vec r1, cpo_vec r2, cpo_vec n){ return k_scalar_body(a, r1, n) + k_scalar_body(b, r2, n); }

/* ---- K8: cpArbiterPreStep (cpArbiter.c:416-439) ---- */
void cpo_arbiter_prestep(cpo_arbiter *arb, const cpo_body *bodies, double dt, double slop, double bias)
{
	const cpo_body *a = &bodies[arb->body_a], *b = &bodies[arb->body_b];
	cpo_vec n = arb->n, body_delta = sub(b->p, a->p);
	for(int i = 0; i < arb->count; i++){
		cpo_contact *con = &arb->contacts[i];
		con->nMass = 1.0/k_scalar(a, b, con->r1, con->r2, n);
		con->tMass = 1.0/k_scalar(a, b, con->r1, con->r2, perp(n));
		double dist = dot(add(sub(con->r2, con->r1), body_delta), n);
		con->bias = -bias*fmin_(0.0, dist + slop)/dt;
		con->jBias = 0.0;
		con->bounce = dot(relative_velocity(a, b, con->r1, con->r2), n)*arb->e;
	}
}

/* ---- cpArbiterApplyCachedImpulse (cpArbiter.c:441-455) ---- */
void cpo_arbiter_apply_cached(cpo_arbiter *arb, cpo_body *bodies, double dt_coef)
{
	if(arb->first_collision) return;
	cpo_body *a = &bodies[arb->body_a], *b = &bodies[arb->body_b];
	for(int i = 0; i < arb->count; i++){
		cpo_contact *con = &arb->contacts[i];
		cpo_vec j = rotate(arb->n, V(con->jnAcc, con->jtAcc));
		apply_impulses(a, b, con->r1, con->r2, mul(j, dt_coef));
	}
}

/* ---- cpArbiterApplyImpulse (cpArbiter.c:459-498) ---- */
void cpo_arbiter_apply_impulse(cpo_arbiter *arb, cpo_body *bodies)
{
	cpo_body *a = &bodies[arb->body_a], *b = &bodies[arb->body_b];
	cpo_vec n = arb->n, surface_vr = arb->surface_vr;
	double friction = arb->u;
	for(int i = 0; i < arb->count; i++){
		cpo_contact *con = &arb->contacts[i];
		double nMass = con->nMass;
		cpo_vec r1 = con->r1, r2 = con->r2;
		cpo_vec vb1 = add(a->v_bias, mul(perp(r1), a->w_bias));
		cpo_vec vb2 = add(b->v_bias, mul(perp(r2), b->w_bias));
		cpo_vec vr = add(relative_velocity(a, b, r1, r2), surface_vr);
		double vbn = dot(sub(vb2, vb1), n), vrn = dot(vr, n), vrt = dot(vr, perp(n));
		double jbn = (con->bias - vbn)*nMass, jbnOld = con->jBias;
		con->jBias = fmax_(jbnOld + jbn, 0.0);
		double jn = -(con->bounce + vrn)*nMass, jnOld = con->jnAcc;
		con->jnAcc = fmax_(jnOld + jn, 0.0);
		double jtMax = friction*con->jnAcc, jt = -vrt*con->tMass, jtOld = con->jtAcc;
		con->jtAcc = clamp(jtOld + jt, -jtMax, jtMax);
		cpo_vec jb = mul(n, con->jBias - jbnOld);
		apply_bias_impulse(a, neg(jb), r1); apply_bias_impulse(b, jb, r2);
		apply_impulses(a, b, r1, r2, rotate(n, V(con->jnAcc - jnOld, con->jtAcc - jtOld)));
	}
}

/* ---- joints: pin cpPinJoint.c:24-75, slide cpSlideJoint.c:24-89, pivot cpPivotJoint.c:24-70,
 * groove cpGrooveJoint.c:24-98, damped spring cpDampedSpring.c:29-77, damped rotary spring
 * cpDampedRotarySpring.c:29-72, rotary limit cpRotaryLimitJoint.c:24-86, ratchet cpRatchetJoint.c:24-89,
 * gear cpGearJoint.c:24-69, simple motor cpSimpleMotor.c:24-65 ---- */
/* k_tensor (chipmunk_private.h:228-262), inverted, as cpMat2x2 (a b c d) */
static void k_tensor(const cpo_body *a, const cpo_body *b, cpo_vec r1, cpo_vec r2, double k[4])
{
	double m_sum = a->m_inv + b->m_inv;
	double k11 = m_sum, k12 = 0.0, k21 = 0.0, k22 = m_sum;
	double a_i = a->i_inv, r1xsq = r1.x*r1.x*a_i, r1ysq = r1.y*r1.y*a_i, r1nxy = -r1.x*r1.y*a_i;
	k11 += r1ysq; k12 += r1nxy; k21 += r1nxy; k22 += r1xsq;
	double b_i = b->i_inv, r2xsq = r2.x*r2.x*b_i, r2ysq = r2.y*r2.y*b_i, r2nxy = -r2.x*r2.y*b_i;
	k11 += r2ysq; k12 += r2nxy; k21 += r2nxy; k22 += r2xsq;
	double det_inv = 1.0/(k11*k22 - k12*k21);
	k[0] = k22*det_inv; k[1] = -k12*det_inv; k[2] = -k21*det_inv; k[3] = k11*det_inv;
}
static double bias_coef(double errorBias, double dt){ return 1.0 - pow(errorBias, dt); }

void cpo_joint_prestep(cpo_joint *j, cpo_body *bodies, const double *T6, double dt)
{
	cpo_body *a = &bodies[j->a], *b = &bodies[j->b];
	const double *Ta = T6 + 6*j->a, *Tb = T6 + 6*j->b;
	double maxBias = j->maxBias;
	if(j->type == 0 || j->type == 1 || j->type == 2 || j->type == 4){
		j->r1 = tvect(Ta, sub(j->anchorA, a->cog));
		j->r2 = tvect(Tb, sub(j->anchorB, b->cog));
	}
	cpo_vec delta = sub(add(b->p, j->r2), add(a->p, j->r1));
	double dist = len(delta);
	switch(j->type){
	case 0: /* pin */
		j->n = mul(delta, 1.0/(dist ? dist : (double)INFINITY));
		j->nMass = 1.0/k_scalar(a, b, j->r1, j->r2, j->n);
		j->bias = clamp(-bias_coef(j->errorBias, dt)*(dist - j->prm[0])/dt, -maxBias, maxBias);
		break;
	case 1: { /* slide */
		double pdist = 0.0;
		if(dist > j->prm[1]){ pdist = dist - j->prm[1]; j->n = normalize(delta); }
		else if(dist < j->prm[0]){ pdist = j->prm[0] - dist; j->n = neg(normalize(delta)); }
		else { j->n = V(0, 0); j->jnAcc = 0.0; }
		j->nMass = 1.0/k_scalar(a, b, j->r1, j->r2, j->n);
		j->bias = clamp(-bias_coef(j->errorBias, dt)*pdist/dt, -maxBias, maxBias);
		break;
	}
	case 2: /* pivot */
		k_tensor(a, b, j->r1, j->r2, j->k);
		j->bias2 = vclamp(mul(delta, -bias_coef(j->errorBias, dt)/dt), maxBias);
		break;
	case 3: { /* groove: anchorA = grv_a, prm[0..1] = grv_b; grv_n = perp(normalize(grv_b - grv_a)) (cpGrooveJoint.c:128) */
		cpo_vec grv_b = V(j->prm[0], j->prm[1]);
		cpo_vec grv_n = perp(normalize(sub(grv_b, j->anchorA)));
		cpo_vec ta = tpoint(Ta, j->anchorA), tb = tpoint(Ta, grv_b);
		cpo_vec n = tvect(Ta, grv_n);
		double d = dot(ta, n);
		j->n = n;
		j->r2 = tvect(Tb, sub(j->anchorB, b->cog));
		double td = cross(add(b->p, j->r2), n);
		if(td <= cross(ta, n)){ j->clamp = 1.0; j->r1 = sub(ta, a->p); }
		else if(td >= cross(tb, n)){ j->clamp = -1.0; j->r1 = sub(tb, a->p); }
		else { j->clamp = 0.0; j->r1 = sub(add(mul(perp(n), -td), mul(n, d)), a->p); }
		k_tensor(a, b, j->r1, j->r2, j->k);
		cpo_vec gdelta = sub(add(b->p, j->r2), add(a->p, j->r1));
		j->bias2 = vclamp(mul(gdelta, -bias_coef(j->errorBias, dt)/dt), maxBias);
		break;
	}
	case 4: { /* damped spring: applies its spring impulse here (cpDampedSpring.c:49-52) */
		j->n = mul(delta, 1.0/(dist ? dist : (double)INFINITY));
		double k = k_scalar(a, b, j->r1, j->r2, j->n);
		j->nMass = 1.0/k;
		j->target_vrn = 0.0;
		j->v_coef = 1.0 - exp(-j->prm[2]*dt*k);
		double f_spring = (j->prm[0] - dist)*j->prm[1];
		double j_spring = j->jnAcc = f_spring*dt;
		apply_impulses(a, b, j->r1, j->r2, mul(j->n, j_spring));
		break;
	}
	case 5: { /* damped rotary spring: prm = restAngle, stiffness, damping; applies its torque here */
		double moment = a->i_inv + b->i_inv;
		j->iSum = 1.0/moment;
		j->v_coef = 1.0 - exp(-j->prm[2]*dt*moment);
		j->target_vrn = 0.0;
		double j_spring = ((a->a - b->a) - j->prm[0])*j->prm[1]*dt;
		j->jnAcc = j_spring;
		a->w -= j_spring*a->i_inv;
		b->w += j_spring*b->i_inv;
		break;
	}
	case 6: { /* rotary limit: prm = min, max */
		double rdist = b->a - a->a, pdist = 0.0;
		if(rdist > j->prm[1]) pdist = j->prm[1] - rdist; else if(rdist < j->prm[0]) pdist = j->prm[0] - rdist;
		j->iSum = 1.0/(a->i_inv + b->i_inv);
		j->bias = clamp(-bias_coef(j->errorBias, dt)*pdist/dt, -maxBias, maxBias);
		if(!j->bias) j->jnAcc = 0.0;
		break;
	}
	case 7: { /* ratchet: prm = angle (state), phase, ratchet */
		double angle = j->prm[0], phase = j->prm[1], ratchet = j->prm[2];
		double rdelta = b->a - a->a, diff = angle - rdelta, pdist = 0.0;
		if(diff*ratchet > 0.0) pdist = diff;
		else j->prm[0] = floor((rdelta - phase)/ratchet)*ratchet + phase;
		j->iSum = 1.0/(a->i_inv + b->i_inv);
		j->bias = clamp(-bias_coef(j->errorBias, dt)*pdist/dt, -maxBias, maxBias);
		if(!j->bias) j->jnAcc = 0.0;
		break;
	}
	case 8: { /* gear: prm = phase, ratio */
		double ratio = j->prm[1], ratio_inv = 1.0/ratio;
		j->iSum = 1.0/(a->i_inv*ratio_inv + ratio*b->i_inv);
		j->bias = clamp(-bias_coef(j->errorBias, dt)*(b->a*ratio - a->a - j->prm[0])/dt, -maxBias, maxBias);
		break;
	}
	case 9: /* simple motor: prm = rate */
		j->iSum = 1.0/(a->i_inv + b->i_inv);
		break;
	default: break;
	}
}

void cpo_joint_apply_cached(cpo_joint *j, cpo_body *bodies, double dt_coef)
{
	cpo_body *a = &bodies[j->a], *b = &bodies[j->b];
	switch(j->type){
	case 0: case 1: apply_impulses(a, b, j->r1, j->r2, mul(j->n, j->jnAcc*dt_coef)); break;
	case 2: case 3: apply_impulses(a, b, j->r1, j->r2, mul(j->jAcc2, dt_coef)); break;
	case 6: case 7: case 9: { double jj = j->jnAcc*dt_coef; a->w -= jj*a->i_inv; b->w += jj*b->i_inv; break; }
	case 8: { double jj = j->jnAcc*dt_coef; a->w -= jj*a->i_inv*(1.0/j->prm[1]); b->w += jj*b->i_inv; break; }
	default: break;
	}
}

void cpo_joint_apply_impulse(cpo_joint *j, cpo_body *bodies, double dt)
{
	cpo_body *a = &bodies[j->a], *b = &bodies[j->b];
	switch(j->type){
	case 0: {
		double vrn = dot(relative_velocity(a, b, j->r1, j->r2), j->n);
		double jnMax = j->maxForce*dt;
		double jn = (j->bias - vrn)*j->nMass, jnOld = j->jnAcc;
		j->jnAcc = clamp(jnOld + jn, -jnMax, jnMax);
		apply_impulses(a, b, j->r1, j->r2, mul(j->n, j->jnAcc - jnOld));
		break;
	}
	case 1: {
		if(eql(j->n, V(0, 0))) return;
		double vrn = dot(relative_velocity(a, b, j->r1, j->r2), j->n);
		double jn = (j->bias - vrn)*j->nMass, jnOld = j->jnAcc;
		j->jnAcc = clamp(jnOld + jn, -j->maxForce*dt, 0.0);
		apply_impulses(a, b, j->r1, j->r2, mul(j->n, j->jnAcc - jnOld));
		break;
	}
	case 2: {
		cpo_vec vr = relative_velocity(a, b, j->r1, j->r2);
		cpo_vec d = sub(j->bias2, vr);
		cpo_vec jj = V(d.x*j->k[0] + d.y*j->k[1], d.x*j->k[2] + d.y*j->k[3]);
		cpo_vec jOld = j->jAcc2;
		j->jAcc2 = vclamp(add(jOld, jj), j->maxForce*dt);
		apply_impulses(a, b, j->r1, j->r2, sub(j->jAcc2, jOld));
		break;
	}
	case 4: {
		double vrn = dot(relative_velocity(a, b, j->r1, j->r2), j->n);
		double v_damp = (j->target_vrn - vrn)*j->v_coef;
		j->target_vrn = vrn + v_damp;
		double j_damp = v_damp*j->nMass;
		j->jnAcc += j_damp;
		apply_impulses(a, b, j->r1, j->r2, mul(j->n, j_damp));
		break;
	}
	case 3: {
		cpo_vec vr = relative_velocity(a, b, j->r1, j->r2);
		cpo_vec d = sub(j->bias2, vr);
		cpo_vec jj = V(d.x*j->k[0] + d.y*j->k[1], d.x*j->k[2] + d.y*j->k[3]);
		cpo_vec jOld = j->jAcc2, jn = add(jOld, jj), n = j->n;
		/* grooveConstrain (cpGrooveJoint.c:73-78); cpvproject (cpVect.h:98-101) */
		cpo_vec jClamp = (j->clamp*cross(jn, n) > 0.0) ? jn : mul(n, dot(jn, n)/dot(n, n));
		j->jAcc2 = vclamp(jClamp, j->maxForce*dt);
		apply_impulses(a, b, j->r1, j->r2, sub(j->jAcc2, jOld));
		break;
	}
	case 5: {
		double wrn = a->w - b->w;
		double w_damp = (j->target_vrn - wrn)*j->v_coef;
		j->target_vrn = wrn + w_damp;
		double j_damp = w_damp*j->iSum;
		j->jnAcc += j_damp;
		a->w += j_damp*a->i_inv; b->w -= j_damp*b->i_inv;
		break;
	}
	case 6: {
		if(!j->bias) return;
		double wr = b->w - a->w, jMax = j->maxForce*dt;
		double jj = -(j->bias + wr)*j->iSum, jOld = j->jnAcc;
		if(j->bias < 0.0) j->jnAcc = clamp(jOld + jj, 0.0, jMax); else j->jnAcc = clamp(jOld + jj, -jMax, 0.0);
		jj = j->jnAcc - jOld;
		a->w -= jj*a->i_inv; b->w += jj*b->i_inv;
		break;
	}
	case 7: {
		if(!j->bias) return;
		double wr = b->w - a->w, ratchet = j->prm[2], jMax = j->maxForce*dt;
		double jj = -(j->bias + wr)*j->iSum, jOld = j->jnAcc;
		j->jnAcc = clamp((jOld + jj)*ratchet, 0.0, jMax*fabs_(ratchet))/ratchet;
		jj = j->jnAcc - jOld;
		a->w -= jj*a->i_inv; b->w += jj*b->i_inv;
		break;
	}
	case 9: {
		double wr = b->w - a->w + j->prm[0], jMax = j->maxForce*dt;
		double jj = -wr*j->iSum, jOld = j->jnAcc;
		j->jnAcc = clamp(jOld + jj, -jMax, jMax);
		jj = j->jnAcc - jOld;
		a->w -= jj*a->i_inv; b->w += jj*b->i_inv;
		break;
	}
	case 8: {
		double ratio = j->prm[1], ratio_inv = 1.0/ratio;
		double wr = b->w*ratio - a->w, jMax = j->maxForce*dt;
		double jj = (j->bias - wr)*j->iSum, jOld = j->jnAcc;
		j->jnAcc = clamp(jOld + jj, -jMax, jMax);
		jj = j->jnAcc - jOld;
		a->w -= jj*a->i_inv*ratio_inv; b->w += jj*b->i_inv;
		break;
	}
	default: break;
	}
}

/* ---- the solver section of cpSpaceStep (cpSpaceStep.c:406-427): cached impulses, then `iterations`
 * sweeps over the arbiters followed by the constraints, in array order ---- */
void cpo_solve(int n_arb, cpo_arbiter *arbs, int n_joints, cpo_joint *joints, cpo_body *bodies, int iterations, double dt, double dt_coef)
{
	for(int i = 0; i < n_arb; i++) cpo_arbiter_apply_cached(&arbs[i], bodies, dt_coef);
	for(int i = 0; i < n_joints; i++) cpo_joint_apply_cached(&joints[i], bodies, dt_coef);
	for(int it = 0; it < iterations; it++){
		for(int i = 0; i < n_arb; i++) cpo_arbiter_apply_impulse(&arbs[i], bodies);
		for(int i = 0; i < n_joints; i++) cpo_joint_apply_impulse(&joints[i], bodies, dt);
	}
}

/* ---- the same solver section for an arbitrary interleaving of arbiters and joints: items[q] >= 0 names
 * arbiter items[q], items[q] < 0 names joint -(items[q] + 1).  This is what a graph-coloured solver does when its
 * colours are replayed one after the other (constraints of one colour share no dynamic body, so their relative
 * order inside the colour cannot matter): cached impulses in sequence order, then `iterations` sweeps in
 * sequence order (cpSpaceStep.c:406-427 with the two arrays merged into one sequence). ---- */
void cpo_solve_sequence(long n_items, const int64_t *items, cpo_arbiter *arbs, cpo_joint *joints, cpo_body *bodies, int iterations, double dt, double dt_coef)
{
	for(long q = 0; q < n_items; q++){
		if(items[q] >= 0) cpo_arbiter_apply_cached(&arbs[items[q]], bodies, dt_coef);
		else cpo_joint_apply_cached(&joints[-(items[q] + 1)], bodies, dt_coef);
	}
	for(int it = 0; it < iterations; it++){
		for(long q = 0; q < n_items; q++){
			if(items[q] >= 0) cpo_arbiter_apply_impulse(&arbs[items[q]], bodies);
			else cpo_joint_apply_impulse(&joints[-(items[q] + 1)], bodies, dt);
		}
	}
}

/* struct sizes for the numpy mirrors in tests/replay.py: 0 cpo_body, 1 cpo_arbiter, 2 cpo_joint */
int cpo_sizeof(int what)
{
	switch(what){ case 0: return (int)sizeof(cpo_body); case 1: return (int)sizeof(cpo_arbiter); case 2: return (int)sizeof(cpo_joint); default: return -1; }
}
