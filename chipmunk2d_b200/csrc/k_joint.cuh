// k_joint.cuh -- joints: preStep / applyCachedImpulse / applyImpulse of the reference's ten
// constraint classes as device functions (thread per joint in preStep; called from the
// coloured or the serial solver for the impulse phases).
//   pin cpPinJoint.c:24-75 | slide cpSlideJoint.c:24-89 | pivot cpPivotJoint.c:24-70
//   groove cpGrooveJoint.c:24-98 | damped spring cpDampedSpring.c:29-77
//   damped rotary spring cpDampedRotarySpring.c:29-72 | rotary limit cpRotaryLimitJoint.c:24-86
//   ratchet cpRatchetJoint.c:24-89 | gear cpGearJoint.c:24-69 | simple motor cpSimpleMotor.c:24-65
// bias_coef() = 1 - pow(errorBias, dt) (chipmunk_private.h:264-268) is computed on the host with
// libm and uploaded per joint, so only exp() (spring damping) is evaluated by the device libm.
#pragma once
#include "cpb_world.h"

// velocity of one body as the solver sees it: V = (v.x, v.y, w, -)
CPB_DEVICE void apply_impulse(double4 &V, V2 mi, V2 j, V2 r){
	// apply_impulse (chipmunk_private.h:185-189)
	V.x = V.x + j.x*mi.x; V.y = V.y + j.y*mi.x;
	V.z += mi.y*vcross(r, j);
}
CPB_DEVICE void apply_impulses(double4 &Va, double4 &Vb, V2 mia, V2 mib, V2 r1, V2 r2, V2 j){
	apply_impulse(Va, mia, vneg(j), r1);
	apply_impulse(Vb, mib, j, r2);
}
CPB_DEVICE V2 relative_velocity(const double4 &Va, const double4 &Vb, V2 r1, V2 r2){
	V2 v1 = vadd(v2(Va.x, Va.y), vmul(vperp(r1), Va.z));
	V2 v2_ = vadd(v2(Vb.x, Vb.y), vmul(vperp(r2), Vb.z));
	return vsub(v2_, v1);
}
CPB_DEVICE double k_scalar2(V2 mia, V2 mib, V2 r1, V2 r2, V2 n){
	double rcn1 = vcross(r1, n), rcn2 = vcross(r2, n);
	return (mia.x + mia.y*rcn1*rcn1) + (mib.x + mib.y*rcn2*rcn2);
}
// k_tensor (chipmunk_private.h:228-262); result packed as cpMat2x2 (a b c d)
CPB_DEVICE double4 k_tensor(V2 mia, V2 mib, V2 r1, V2 r2){
	double m_sum = mia.x + mib.x;
	double k11 = m_sum, k12 = 0.0, k21 = 0.0, k22 = m_sum;
	double a_i_inv = mia.y;
	double r1xsq =  r1.x * r1.x * a_i_inv;
	double r1ysq =  r1.y * r1.y * a_i_inv;
	double r1nxy = -r1.x * r1.y * a_i_inv;
	k11 += r1ysq; k12 += r1nxy; k21 += r1nxy; k22 += r1xsq;
	double b_i_inv = mib.y;
	double r2xsq =  r2.x * r2.x * b_i_inv;
	double r2ysq =  r2.y * r2.y * b_i_inv;
	double r2nxy = -r2.x * r2.y * b_i_inv;
	k11 += r2ysq; k12 += r2nxy; k21 += r2nxy; k22 += r2xsq;
	double det = k11*k22 - k12*k21;
	double det_inv = 1.0/det;
	return make_double4(k22*det_inv, -k12*det_inv, -k21*det_inv, k11*det_inv);
}
CPB_DEVICE V2 mat_transform(double4 m, V2 v){ return v2(v.x*m.x + v.y*m.y, v.x*m.z + v.y*m.w); }

#ifndef CPB_EMU
__device__ __forceinline__ void atomic_add_d(double *p, double v){ atomicAdd(p, v); }
#else
static inline void atomic_add_d(double *p, double v){ *p += v; }
#endif

// preStep of every joint class.  Spring classes apply their spring impulse to the bodies here
// (cpDampedSpring.c:49-52, cpDampedRotarySpring.c:47-52); bodies shared by several springs are
// updated with fp64 atomics.
__global__ void k_joint_prestep(DJoints J, DBodies B, double dt)
{
	int j = CPB_TID;
	if(j >= J.n) return;
	int a = J.a[j], b = J.b[j];
	J.hint[j] = J.colour[j];   // last step's colour (or a negative marker) seeds this step's colouring
	if(B.sleeping[a] || B.sleeping[b]){
		// the reference removes joints of sleeping bodies from space->constraints (cpSpaceComponent.c:107-110)
		if((B.sleeping[a] || B.type[a] == CPB200_BODY_STATIC) && (B.sleeping[b] || B.type[b] == CPB200_BODY_STATIC)){ J.colour[j] = -2; return; }
	}
	J.colour[j] = -1;
	int type = J.type[j];
	Xf Ta, Tb; Ta.rot = B.rot[a]; Ta.t = B.txy[a]; Tb.rot = B.rot[b]; Tb.t = B.txy[b];
	V2 pa = B.pos[a], pb = B.pos[b];
	V2 mia = B.MI[a], mib = B.MI[b];
	double max_bias = J.max_bias[j];
	double bcoef = J.bias_coef[j];
	double4 prm = J.prm[j];
	switch(type){
	case CPB200_JOINT_PIN: {
		V2 r1 = xf_vect(Ta, vsub(J.anchor_a[j], B.cog[a])), r2 = xf_vect(Tb, vsub(J.anchor_b[j], B.cog[b]));
		V2 delta = vsub(vadd(pb, r2), vadd(pa, r1));
		double dist = vlen(delta);
		V2 n = vmul(delta, 1.0/(dist ? dist : (double)INFINITY));
		J.r1[j] = r1; J.r2[j] = r2; J.nrm[j] = n;
		J.nmass[j] = 1.0/k_scalar2(mia, mib, r1, r2, n);
		J.bias[j] = v2(fclamp_cp(-bcoef*(dist - prm.x)/dt, -max_bias, max_bias), 0.0);
		break;
	}
	case CPB200_JOINT_SLIDE: {
		V2 r1 = xf_vect(Ta, vsub(J.anchor_a[j], B.cog[a])), r2 = xf_vect(Tb, vsub(J.anchor_b[j], B.cog[b]));
		V2 delta = vsub(vadd(pb, r2), vadd(pa, r1));
		double dist = vlen(delta);
		double pdist = 0.0;
		V2 n;
		if(dist > prm.y){ pdist = dist - prm.y; n = vnormalize(delta); }
		else if(dist < prm.x){ pdist = prm.x - dist; n = vneg(vnormalize(delta)); }
		else { n = v2(0.0, 0.0); J.acc[j] = v2(0.0, 0.0); }
		J.r1[j] = r1; J.r2[j] = r2; J.nrm[j] = n;
		J.nmass[j] = 1.0/k_scalar2(mia, mib, r1, r2, n);
		J.bias[j] = v2(fclamp_cp(-bcoef*pdist/dt, -max_bias, max_bias), 0.0);
		break;
	}
	case CPB200_JOINT_PIVOT: {
		V2 r1 = xf_vect(Ta, vsub(J.anchor_a[j], B.cog[a])), r2 = xf_vect(Tb, vsub(J.anchor_b[j], B.cog[b]));
		J.r1[j] = r1; J.r2[j] = r2;
		J.k[j] = k_tensor(mia, mib, r1, r2);
		V2 delta = vsub(vadd(pb, r2), vadd(pa, r1));
		J.bias[j] = vclamp(vmul(delta, -bcoef/dt), max_bias);
		break;
	}
	case CPB200_JOINT_GROOVE: {
		// anchor_a = grv_a, prm.xy = grv_b, prm.zw = grv_n
		V2 ta = xf_point(Ta, J.anchor_a[j]);
		V2 tb = xf_point(Ta, v2(prm.x, prm.y));
		V2 n = xf_vect(Ta, v2(prm.z, prm.w));
		double d = vdot(ta, n);
		V2 r2 = xf_vect(Tb, vsub(J.anchor_b[j], B.cog[b]));
		double td = vcross(vadd(pb, r2), n);
		V2 r1; double clamp;
		if(td <= vcross(ta, n)){ clamp = 1.0; r1 = vsub(ta, pa); }
		else if(td >= vcross(tb, n)){ clamp = -1.0; r1 = vsub(tb, pa); }
		else { clamp = 0.0; r1 = vsub(vadd(vmul(vperp(n), -td), vmul(n, d)), pa); }
		J.nrm[j] = n; J.r1[j] = r1; J.r2[j] = r2; J.nmass[j] = clamp;
		J.k[j] = k_tensor(mia, mib, r1, r2);
		V2 delta = vsub(vadd(pb, r2), vadd(pa, r1));
		J.bias[j] = vclamp(vmul(delta, -bcoef/dt), max_bias);
		break;
	}
	case CPB200_JOINT_DAMPED_SPRING: {
		// prm = restLength, stiffness, damping
		V2 r1 = xf_vect(Ta, vsub(J.anchor_a[j], B.cog[a])), r2 = xf_vect(Tb, vsub(J.anchor_b[j], B.cog[b]));
		V2 delta = vsub(vadd(pb, r2), vadd(pa, r1));
		double dist = vlen(delta);
		V2 n = vmul(delta, 1.0/(dist ? dist : (double)INFINITY));
		double k = k_scalar2(mia, mib, r1, r2, n);
		J.r1[j] = r1; J.r2[j] = r2; J.nrm[j] = n;
		J.nmass[j] = 1.0/k;
		J.aux0[j] = 0.0;                                  // target_vrn
		J.aux1[j] = 1.0 - exp(-prm.z*dt*k);               // v_coef
		V2 fc = J.jspring[j];                            // a host springForceFunc's answer for this step, if any
		double f_spring = (fc.y != 0.0 ? fc.x : (prm.x - dist)*prm.y);   // defaultSpringForce (cpDampedSpring.c:24-27)
		J.jspring[j] = v2(0.0, 0.0);
		double j_spring = f_spring*dt;
		J.acc[j] = v2(j_spring, 0.0);
		V2 imp = vmul(n, j_spring);
		// apply_impulses(a, b, r1, r2, n*j_spring)
		V2 ni = vneg(imp);
		if(mia.x != 0.0 || mia.y != 0.0){
			atomic_add_d(&B.V[a].x, ni.x*mia.x); atomic_add_d(&B.V[a].y, ni.y*mia.x); atomic_add_d(&B.V[a].z, mia.y*vcross(r1, ni));
		}
		if(mib.x != 0.0 || mib.y != 0.0){
			atomic_add_d(&B.V[b].x, imp.x*mib.x); atomic_add_d(&B.V[b].y, imp.y*mib.x); atomic_add_d(&B.V[b].z, mib.y*vcross(r2, imp));
		}
		break;
	}
	case CPB200_JOINT_DAMPED_ROTARY_SPRING: {
		// prm = restAngle, stiffness, damping
		double moment = mia.y + mib.y;
		J.nmass[j] = 1.0/moment;                          // iSum
		J.aux1[j] = 1.0 - exp(-prm.z*dt*moment);          // w_coef
		J.aux0[j] = 0.0;                                  // target_wrn
		V2 fc = J.jspring[j];                            // a host springTorqueFunc's answer for this step, if any
		J.jspring[j] = v2(0.0, 0.0);
		double j_spring = (fc.y != 0.0 ? fc.x : ((B.ang[a] - B.ang[b]) - prm.x)*prm.y)*dt;
		J.acc[j] = v2(j_spring, 0.0);
		if(mia.y != 0.0) atomic_add_d(&B.V[a].z, -(j_spring*mia.y));
		if(mib.y != 0.0) atomic_add_d(&B.V[b].z, j_spring*mib.y);
		break;
	}
	case CPB200_JOINT_ROTARY_LIMIT: {
		double dist = B.ang[b] - B.ang[a];
		double pdist = 0.0;
		if(dist > prm.y) pdist = prm.y - dist; else if(dist < prm.x) pdist = prm.x - dist;
		J.nmass[j] = 1.0/(mia.y + mib.y);
		double bias = fclamp_cp(-bcoef*pdist/dt, -max_bias, max_bias);
		J.bias[j] = v2(bias, 0.0);
		if(!bias) J.acc[j] = v2(0.0, 0.0);
		break;
	}
	case CPB200_JOINT_RATCHET: {
		// aux0 = angle, prm.y = phase, prm.z = ratchet
		double angle = J.aux0[j], phase = prm.y, ratchet = prm.z;
		double delta = B.ang[b] - B.ang[a];
		double diff = angle - delta;
		double pdist = 0.0;
		if(diff*ratchet > 0.0) pdist = diff;
		else J.aux0[j] = floor((delta - phase)/ratchet)*ratchet + phase;
		J.nmass[j] = 1.0/(mia.y + mib.y);
		double bias = fclamp_cp(-bcoef*pdist/dt, -max_bias, max_bias);
		J.bias[j] = v2(bias, 0.0);
		if(!bias) J.acc[j] = v2(0.0, 0.0);
		break;
	}
	case CPB200_JOINT_GEAR: {
		// prm = phase, ratio ; ratio_inv = 1/ratio
		double ratio = prm.y, ratio_inv = 1.0/ratio;
		J.nmass[j] = 1.0/(mia.y*ratio_inv + ratio*mib.y);
		J.bias[j] = v2(fclamp_cp(-bcoef*(B.ang[b]*ratio - B.ang[a] - prm.x)/dt, -max_bias, max_bias), 0.0);
		break;
	}
	case CPB200_JOINT_SIMPLE_MOTOR: {
		J.nmass[j] = 1.0/(mia.y + mib.y);
		break;
	}
	default: break;
	}
}

// applyCachedImpulse of joint j on register copies of the two bodies' velocities.
CPB_DEVICE void joint_apply_cached(const DJoints &J, int j, double4 &Va, double4 &Vb, V2 mia, V2 mib, double dt_coef)
{
	switch(J.type[j]){
	case CPB200_JOINT_PIN: case CPB200_JOINT_SLIDE:
		apply_impulses(Va, Vb, mia, mib, J.r1[j], J.r2[j], vmul(J.nrm[j], J.acc[j].x*dt_coef));
		break;
	case CPB200_JOINT_PIVOT: case CPB200_JOINT_GROOVE:
		apply_impulses(Va, Vb, mia, mib, J.r1[j], J.r2[j], vmul(J.acc[j], dt_coef));
		break;
	case CPB200_JOINT_GEAR: {
		double jj = J.acc[j].x*dt_coef;
		double ratio_inv = 1.0/J.prm[j].y;
		Va.z -= jj*mia.y*ratio_inv; Vb.z += jj*mib.y;
		break;
	}
	case CPB200_JOINT_ROTARY_LIMIT: case CPB200_JOINT_RATCHET: case CPB200_JOINT_SIMPLE_MOTOR: {
		double jj = J.acc[j].x*dt_coef;
		Va.z -= jj*mia.y; Vb.z += jj*mib.y;
		break;
	}
	default: break; // springs: no-op
	}
}

// applyImpulse of joint j.
CPB_DEVICE void joint_apply(const DJoints &J, int j, double4 &Va, double4 &Vb, V2 mia, V2 mib, double dt)
{
	double max_force = J.max_force[j];
	switch(J.type[j]){
	case CPB200_JOINT_PIN: {
		V2 n = J.nrm[j], r1 = J.r1[j], r2 = J.r2[j];
		double vrn = vdot(relative_velocity(Va, Vb, r1, r2), n);
		double jnMax = max_force*dt;
		double jn = (J.bias[j].x - vrn)*J.nmass[j];
		double jnOld = J.acc[j].x;
		double jnAcc = fclamp_cp(jnOld + jn, -jnMax, jnMax);
		J.acc[j] = v2(jnAcc, 0.0);
		jn = jnAcc - jnOld;
		apply_impulses(Va, Vb, mia, mib, r1, r2, vmul(n, jn));
		break;
	}
	case CPB200_JOINT_SLIDE: {
		V2 n = J.nrm[j];
		if(n.x == 0.0 && n.y == 0.0) return;
		V2 r1 = J.r1[j], r2 = J.r2[j];
		V2 vr = relative_velocity(Va, Vb, r1, r2);
		double vrn = vdot(vr, n);
		double jn = (J.bias[j].x - vrn)*J.nmass[j];
		double jnOld = J.acc[j].x;
		double jnAcc = fclamp_cp(jnOld + jn, -max_force*dt, 0.0);
		J.acc[j] = v2(jnAcc, 0.0);
		jn = jnAcc - jnOld;
		apply_impulses(Va, Vb, mia, mib, r1, r2, vmul(n, jn));
		break;
	}
	case CPB200_JOINT_PIVOT: {
		V2 r1 = J.r1[j], r2 = J.r2[j];
		V2 vr = relative_velocity(Va, Vb, r1, r2);
		V2 jj = mat_transform(J.k[j], vsub(J.bias[j], vr));
		V2 jOld = J.acc[j];
		V2 jAcc = vclamp(vadd(jOld, jj), max_force*dt);
		J.acc[j] = jAcc;
		jj = vsub(jAcc, jOld);
		apply_impulses(Va, Vb, mia, mib, r1, r2, jj);
		break;
	}
	case CPB200_JOINT_GROOVE: {
		V2 r1 = J.r1[j], r2 = J.r2[j];
		V2 vr = relative_velocity(Va, Vb, r1, r2);
		V2 jj = mat_transform(J.k[j], vsub(J.bias[j], vr));
		V2 jOld = J.acc[j];
		V2 jn = vadd(jOld, jj);
		V2 n = J.nrm[j];
		// grooveConstrain (cpGrooveJoint.c:73-78); cpvproject (cpVect.h:98-101)
		V2 jClamp = (J.nmass[j]*vcross(jn, n) > 0.0) ? jn : vmul(n, vdot(jn, n)/vdot(n, n));
		V2 jAcc = vclamp(jClamp, max_force*dt);
		J.acc[j] = jAcc;
		jj = vsub(jAcc, jOld);
		apply_impulses(Va, Vb, mia, mib, r1, r2, jj);
		break;
	}
	case CPB200_JOINT_DAMPED_SPRING: {
		V2 n = J.nrm[j], r1 = J.r1[j], r2 = J.r2[j];
		double vrn = vdot(relative_velocity(Va, Vb, r1, r2), n);
		double v_damp = (J.aux0[j] - vrn)*J.aux1[j];
		J.aux0[j] = vrn + v_damp;
		double j_damp = v_damp*J.nmass[j];
		J.acc[j] = v2(J.acc[j].x + j_damp, 0.0);
		apply_impulses(Va, Vb, mia, mib, r1, r2, vmul(n, j_damp));
		break;
	}
	case CPB200_JOINT_DAMPED_ROTARY_SPRING: {
		double wrn = Va.z - Vb.z;
		double w_damp = (J.aux0[j] - wrn)*J.aux1[j];
		J.aux0[j] = wrn + w_damp;
		double j_damp = w_damp*J.nmass[j];
		J.acc[j] = v2(J.acc[j].x + j_damp, 0.0);
		Va.z += j_damp*mia.y; Vb.z -= j_damp*mib.y;
		break;
	}
	case CPB200_JOINT_ROTARY_LIMIT: {
		double bias = J.bias[j].x;
		if(!bias) return;
		double wr = Vb.z - Va.z;
		double jMax = max_force*dt;
		double jj = -(bias + wr)*J.nmass[j];
		double jOld = J.acc[j].x;
		double jAcc = (bias < 0.0) ? fclamp_cp(jOld + jj, 0.0, jMax) : fclamp_cp(jOld + jj, -jMax, 0.0);
		J.acc[j] = v2(jAcc, 0.0);
		jj = jAcc - jOld;
		Va.z -= jj*mia.y; Vb.z += jj*mib.y;
		break;
	}
	case CPB200_JOINT_RATCHET: {
		double bias = J.bias[j].x;
		if(!bias) return;
		double wr = Vb.z - Va.z;
		double ratchet = J.prm[j].z;
		double jMax = max_force*dt;
		double jj = -(bias + wr)*J.nmass[j];
		double jOld = J.acc[j].x;
		double jAcc = fclamp_cp((jOld + jj)*ratchet, 0.0, jMax*fabs_cp(ratchet))/ratchet;
		J.acc[j] = v2(jAcc, 0.0);
		jj = jAcc - jOld;
		Va.z -= jj*mia.y; Vb.z += jj*mib.y;
		break;
	}
	case CPB200_JOINT_GEAR: {
		double ratio = J.prm[j].y, ratio_inv = 1.0/ratio;
		double wr = Vb.z*ratio - Va.z;
		double jMax = max_force*dt;
		double jj = (J.bias[j].x - wr)*J.nmass[j];
		double jOld = J.acc[j].x;
		double jAcc = fclamp_cp(jOld + jj, -jMax, jMax);
		J.acc[j] = v2(jAcc, 0.0);
		jj = jAcc - jOld;
		Va.z -= jj*mia.y*ratio_inv; Vb.z += jj*mib.y;
		break;
	}
	case CPB200_JOINT_SIMPLE_MOTOR: {
		double wr = Vb.z - Va.z + J.prm[j].x;
		double jMax = max_force*dt;
		double jj = -wr*J.nmass[j];
		double jOld = J.acc[j].x;
		double jAcc = fclamp_cp(jOld + jj, -jMax, jMax);
		J.acc[j] = v2(jAcc, 0.0);
		jj = jAcc - jOld;
		Va.z -= jj*mia.y; Vb.z += jj*mib.y;
		break;
	}
	default: break;
	}
}

// getImpulse of each class (e.g. cpPinJoint.c:77-81): |jnAcc| or |jAcc|
CPB_DEVICE double joint_impulse(const DJoints &J, int j){
	switch(J.type[j]){
	case CPB200_JOINT_PIVOT: case CPB200_JOINT_GROOVE: return vlen(J.acc[j]);
	case CPB200_JOINT_DAMPED_SPRING: case CPB200_JOINT_DAMPED_ROTARY_SPRING: return J.acc[j].x;
	default: return fabs_cp(J.acc[j].x);
	}
}
