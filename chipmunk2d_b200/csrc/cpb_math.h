// cpb_math.h -- device vector math with the reference's exact operation order.
//
// The reference's inline helpers (cpVect.h:58-206, chipmunk_types.h:119-146,
// cpTransform.h:73-84) fix how every product and sum is associated; the kernels
// are compiled with -fmad=false and use these helpers so that, apart from sin/cos
// and exp, device results are bit-identical to the reference built with
// -ffp-contract=off.  cpfmax/cpfmin are branchy ternaries (NaN behaviour differs
// from fmax/fmin), reproduced as such.
#pragma once
#include "cpb_rt.h"
#include <float.h>

typedef double2 V2;

// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256).  A double4 read through the compiler's default
// path is split into two 128-bit requests, i.e. two L1/L2 transactions for the same 32-byte sector; the hot
// random gathers (body velocity sectors, BVH child boxes, packed shape sectors) use these instead.
#ifndef CPB_EMU
__device__ __forceinline__ double4 ld4_cg(const double4 *p){   // through L2 (data written by other SMs)
	double4 v;
	asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
	return v;
}
__device__ __forceinline__ double4 ld4_nc(const double4 *p){   // read-only for the lifetime of the kernel
	double4 v;
	asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
	return v;
}
__device__ __forceinline__ double4 ld4_cs(const double4 *p){   // streamed once: evict-first
	double4 v;
	asm volatile("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
	return v;
}
__device__ __forceinline__ void st4_cs(double4 *p, double4 v){
	asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ double4 ld4_ca(const double4 *p){   // default caching (L1 + L2)
	double4 v;
	asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
	return v;
}
__device__ __forceinline__ void st4_wb(double4 *p, double4 v){
	asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ void st4_cg(double4 *p, double4 v){
	asm volatile("st.global.cg.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
#else
static inline double4 ld4_cg(const double4 *p){ return *p; }
static inline double4 ld4_nc(const double4 *p){ return *p; }
static inline void st4_cg(double4 *p, double4 v){ *p = v; }
static inline double4 ld4_cs(const double4 *p){ return *p; }
static inline void st4_cs(double4 *p, double4 v){ *p = v; }
static inline double4 ld4_ca(const double4 *p){ return *p; }
static inline void st4_wb(double4 *p, double4 v){ *p = v; }
#endif

CPB_HD V2 v2(double x, double y){ return make_double2(x, y); }
CPB_HD V2 vadd(V2 a, V2 b){ return v2(a.x + b.x, a.y + b.y); }
CPB_HD V2 vsub(V2 a, V2 b){ return v2(a.x - b.x, a.y - b.y); }
CPB_HD V2 vneg(V2 a){ return v2(-a.x, -a.y); }
CPB_HD V2 vmul(V2 a, double s){ return v2(a.x*s, a.y*s); }
CPB_HD double vdot(V2 a, V2 b){ return a.x*b.x + a.y*b.y; }
CPB_HD double vcross(V2 a, V2 b){ return a.x*b.y - a.y*b.x; }
CPB_HD V2 vperp(V2 a){ return v2(-a.y, a.x); }
CPB_HD V2 vrperp(V2 a){ return v2(a.y, -a.x); }
CPB_HD V2 vrotate(V2 a, V2 b){ return v2(a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x); }
CPB_HD double vlensq(V2 a){ return vdot(a, a); }
CPB_HD double vlen(V2 a){ return sqrt(vdot(a, a)); }
CPB_HD V2 vlerp(V2 a, V2 b, double t){ return vadd(vmul(a, 1.0 - t), vmul(b, t)); }
CPB_HD V2 vnormalize(V2 a){ return vmul(a, 1.0/(vlen(a) + DBL_MIN)); }
CPB_HD bool veql(V2 a, V2 b){ return a.x == b.x && a.y == b.y; }
CPB_HD double fmax_cp(double a, double b){ return (a > b) ? a : b; }
CPB_HD double fmin_cp(double a, double b){ return (a < b) ? a : b; }
CPB_HD double fabs_cp(double f){ return (f < 0) ? -f : f; }
CPB_HD double fclamp_cp(double f, double lo, double hi){ return fmin_cp(fmax_cp(f, lo), hi); }
CPB_HD double fclamp01_cp(double f){ return fmax_cp(0.0, fmin_cp(f, 1.0)); }
CPB_HD V2 vclamp(V2 v, double len){ return (vdot(v, v) > len*len) ? vmul(vnormalize(v), len) : v; }

// Rigid transform stored as rotation (cos, sin) + translation:
// cpTransform {a=rot.x, b=rot.y, c=-rot.y, d=rot.x, tx, ty} (cpBody.c:347-357).
struct Xf { V2 rot; V2 t; };
// cpTransformPoint: (a*x + c*y + tx, b*x + d*y + ty)
CPB_HD V2 xf_point(Xf T, V2 p){ return v2(T.rot.x*p.x + (-T.rot.y)*p.y + T.t.x, T.rot.y*p.x + T.rot.x*p.y + T.t.y); }
// cpTransformVect: (a*x + c*y, b*x + d*y)
CPB_HD V2 xf_vect(Xf T, V2 v){ return v2(T.rot.x*v.x + (-T.rot.y)*v.y, T.rot.y*v.x + T.rot.x*v.y); }

// CP_HASH_PAIR (chipmunk_private.h:28-29) on 64-bit uintptr_t
#define CPB_HASH_COEF 3344921057ull
CPB_HD uint64_t hash_pair(uint64_t a, uint64_t b){ return (a*CPB_HASH_COEF) ^ (b*CPB_HASH_COEF); }

// 64-bit mix (splitmix64 finaliser) for table slots and colouring priorities
CPB_HD uint64_t mix64(uint64_t x){
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31; return x;
}
