// k_body.cuh -- K1 (integrate position + transform), K2 (shape world cache + AABB),
// K9 (integrate velocity).  One thread per body / per shape; SoA fp64; -fmad=false.
#pragma once
#include "cpb_world.h"

// K1: cpBodyUpdatePosition + SetTransform (cpBody.c:511-522, 347-357) for every body of the
// reference's dynamicBodies array: awake dynamic AND kinematic bodies (cpSpace.c:447).
// (also clears the per-body scratch words of this step's graph colouring)
__global__ void k_integrate_pos(DBodies B, double dt, unsigned long long *claim, unsigned long long *bmask)
{
	int i = CPB_TID;
	if(i >= B.n) return;
	claim[i] = 0ull; bmask[i] = 0ull;
	if(B.type[i] == CPB200_BODY_STATIC || B.sleeping[i]) return;
	if(B.custom[i] & CPB200_BODY_HOST_POSITION) return;   // the host ran this body's position_func and uploaded the result
	double4 V = B.V[i], VB = B.VB[i];
	V2 p = vadd(B.pos[i], vmul(vadd(v2(V.x, V.y), v2(VB.x, VB.y)), dt));
	double a = B.ang[i] + (V.z + VB.z)*dt;
	double s, c;
	sincos(a, &s, &c);
	V2 rot = v2(c, s);
	V2 cg = B.cog[i];
	B.pos[i] = p;
	B.ang[i] = a;
	B.rot[i] = rot;
	B.txy[i] = v2(p.x - (cg.x*rot.x - cg.y*rot.y), p.y - (cg.x*rot.y + cg.y*rot.x));
	B.VB[i] = make_double4(0.0, 0.0, 0.0, VB.w);   // lane w carries i_inv for the solver
}

// Translation part of SetTransform for bodies whose rotation came from the host
// (initial upload: the host's libm cos/sin are kept so static geometry is bit-identical).
__global__ void k_body_transform(DBodies B, int first, int n)
{
	int i = first + CPB_TID;
	if(i >= first + n || i >= B.n) return;
	V2 p = B.pos[i], rot = B.rot[i], cg = B.cog[i];
	B.txy[i] = v2(p.x - (cg.x*rot.x - cg.y*rot.y), p.y - (cg.x*rot.y + cg.y*rot.x));
}

// K2: cpShapeUpdateFunc -> cacheData (cpSpaceStep.c:329-333; cpShape.c:291-296, 378-405;
// cpPolyShape.c:39-64).  all != 0 recaches every shape (upload); otherwise only shapes in
// the reference's dynamic index: those on awake, non-static bodies.
__global__ void k_shape_cache(DShapes S, DBodies B, int all)
{
	int s = CPB_TID;
	if(s >= S.n) return;
	int b = S.body[s];
	if(!all && (B.type[b] == CPB200_BODY_STATIC || B.sleeping[b])) return;
	Xf T; T.rot = B.rot[b]; T.t = B.txy[b];
	double rad = S.r[s];
	int type = S.type[s];
	if(type == CPB200_SHAPE_CIRCLE){
		V2 c = xf_point(T, S.la[s]);
		S.wa[s] = c;
		S.bb[s] = make_double4(c.x - rad, c.y - rad, c.x + rad, c.y + rad);
		// the packed line of the circle-circle narrowphase (cpb_world.h)
		const double4 mt = S.mat[s];
		const V2 bp = B.pos[b];
		const unsigned lo = ((unsigned)b & 0x1fffffffu) | ((mt.z != 0.0 || mt.w != 0.0) ? 0x20000000u : 0u) |
			(B.type[b] != CPB200_BODY_DYNAMIC ? 0x40000000u : 0u) | (S.sensor[s] ? 0x80000000u : 0u);
		const unsigned long long bits = (unsigned long long)lo | ((unsigned long long)S.ids[s].x << 32);
		S.circ[2*(size_t)s] = make_double4(c.x, c.y, rad, __longlong_as_double((long long)bits));
		S.circ[2*(size_t)s + 1] = make_double4(bp.x, bp.y, mt.x, mt.y);
	} else if(type == CPB200_SHAPE_SEGMENT){
		V2 ta = xf_point(T, S.la[s]);
		V2 tb = xf_point(T, S.lb[s]);
		S.wa[s] = ta; S.wb[s] = tb; S.wn[s] = xf_vect(T, S.ln[s]);
		double l, r, bt, t;
		if(ta.x < tb.x){ l = ta.x; r = tb.x; } else { l = tb.x; r = ta.x; }
		if(ta.y < tb.y){ bt = ta.y; t = tb.y; } else { bt = tb.y; t = ta.y; }
		S.bb[s] = make_double4(l - rad, bt - rad, r + rad, t + rad);
	} else {
		int off = S.poff[s], cnt = S.pcount[s];
		double l = INFINITY, r = -INFINITY, bt = INFINITY, t = -INFINITY;
		for(int k = 0; k < cnt; k++){
			V2 v = xf_point(T, S.lpv[off + k]);
			V2 n = xf_vect(T, S.lpn[off + k]);
			S.wpv[off + k] = v;
			S.wpn[off + k] = n;
			l = fmin_cp(l, v.x); r = fmax_cp(r, v.x);
			bt = fmin_cp(bt, v.y); t = fmax_cp(t, v.y);
		}
		S.bb[s] = make_double4(l - rad, bt - rad, r + rad, t + rad);
	}
}

// K9: cpBodyUpdateVelocity (cpBody.c:493-509); kinematic bodies are skipped, forces reset.
// Also clears the two per-body words of the colouring that follows (one launch instead of two memsets).
__global__ void k_integrate_vel(DBodies B, const DSpace *__restrict__ spaces, double dt)
{
	int i = CPB_TID;
	if(i >= B.n) return;
	if(B.type[i] != CPB200_BODY_DYNAMIC || B.sleeping[i]) return;
	if(B.custom[i] & CPB200_BODY_HOST_VELOCITY) return;   // the host runs this body's velocity_func between prestep and solver
	DSpace sp = spaces[B.space[i]];
	double4 V = B.V[i];
	V2 mi = B.MI[i];
	V2 f = B.force[i];
	double damping = sp.damping_dt;
	V2 v = vadd(vmul(v2(V.x, V.y), damping), vmul(vadd(sp.gravity, vmul(f, mi.x)), dt));
	double w = V.z*damping + B.torque[i]*mi.y*dt;
	B.V[i] = make_double4(v.x, v.y, w, V.w);         // lane w carries m_inv for the solver
	B.force[i] = v2(0.0, 0.0);
	B.torque[i] = 0.0;
}
