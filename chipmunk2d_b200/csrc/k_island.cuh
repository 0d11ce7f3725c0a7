// k_island.cuh -- K7: islands and sleeping.
//
// Replaces cpSpaceProcessComponents (cpSpaceComponent.c:220-307) and the activate/deactivate
// machinery around it (cpSpaceComponent.c:28-153).  The reference threads arbiters onto
// per-body linked lists and flood-fills components with a recursive DFS; here components are
// found with a lock-free union-find over the constraint edges (active arbiters + joints between
// awake dynamic bodies), in five small kernels:
//   idle timers -> wake marks (a sleeping body touched by an active arbiter wakes its whole
//   sleeping component; kinematic contact keeps bodies awake) -> union-find -> a component
//   falls asleep iff every member's idleTime >= sleepTimeThreshold -> arbiters whose bodies all
//   rest become dormant (kept with their contacts, not solved; cpSpaceComponent.c:94-105) until
//   k_arb_carry sees one of their bodies awake again and re-activates them
//   (cpSpaceComponent.c:45-73).
// Sleeping shapes simply stop being "active" leaves of the LBVH (= moved to the static index).
#pragma once
#include "cpb_world.h"

struct DIslands {
	int *parent;        // union-find parent per body
	int *wake;          // per body: component rooted here must wake
	int *comp_active;   // per body (root): some member is not idle long enough
	int *woken;         // per body: woke up this step
	int *touch;         // per body: idle long enough itself, but joined to an awake body that is not
	int *any_woken;     // [1] some body woke up this step
	int *flags;         // [2][4] per step parity: (some awake body has idled long enough, some body sleeps, some body is kinematic, -)
	                    // written by k_sleep_idle; in a live pile all three are 0 and the passes over the arbiters return at once
};
#define ISL_IDLE 0
#define ISL_SLEEPING 1
#define ISL_KINEMATIC 2

CPB_DEVICE bool space_sleeps(const DSpace &sp){ return sp.sleep_threshold != INFINITY; }

// idle timers (cpSpaceComponent.c:237-250)
__global__ void k_sleep_idle(DBodies B, DIslands I, const DSpace *__restrict__ spaces, double dt, int par)
{
	int i = CPB_TID;
	if(i >= B.n) return;
	I.parent[i] = i; I.wake[i] = 0; I.comp_active[i] = 0; I.woken[i] = 0; I.touch[i] = 0;
	if(i == 0){ *I.any_woken = 0; int *nf = I.flags + 4*(par ^ 1); nf[0] = nf[1] = nf[2] = nf[3] = 0; }   // the NEXT step's flags
	int *fl = I.flags + 4*par;
	if(B.type[i] == CPB200_BODY_KINEMATIC && !fl[ISL_KINEMATIC]) fl[ISL_KINEMATIC] = 1;
	if(B.sleeping[i] && !fl[ISL_SLEEPING]) fl[ISL_SLEEPING] = 1;
	if(B.type[i] != CPB200_BODY_DYNAMIC || B.sleeping[i]) return;
	DSpace sp = spaces[B.space[i]];
	if(!space_sleeps(sp)) return;
	double dv = sp.idle_speed;
	double dvsq = (dv ? dv*dv : vlensq(sp.gravity)*dt*dt);
	V2 M = B.M[i];
	double4 V = B.V[i];
	double keThreshold = (dvsq ? M.x*dvsq : 0.0);
	double vsq = V.x*V.x + V.y*V.y, wsq = V.z*V.z;
	double ke = (vsq ? vsq*M.x : 0.0) + (wsq ? wsq*M.y : 0.0);
	const double idle = (ke > keThreshold ? 0.0 : B.idle[i] + dt);
	B.idle[i] = idle;
	if(idle >= sp.sleep_threshold && !fl[ISL_IDLE]) fl[ISL_IDLE] = 1;
}

// wake marks from this step's active arbiters and from joints (cpSpaceComponent.c:253-278)
__global__ void k_sleep_wake_mark(DBodies B, DIslands I, DArbs A, DJoints J, const DSpace *__restrict__ spaces, int par)
{
	// only a sleeping body can be woken and only a kinematic one resets its partner's timer
	if(!I.flags[4*par + ISL_SLEEPING] && !I.flags[4*par + ISL_KINEMATIC]) return;
	int nA = *A.count_ptr; if(nA > A.cap) nA = A.cap;
	int total = nA + J.n;
	for(int c = CPB_TID; c < total; c += CPB_NTHREADS){
		int a, b;
		if(c < nA){ if(A.active[c] != 1) continue; a = A.ba[c]; b = A.bb[c]; }
		else { a = J.a[c - nA]; b = J.b[c - nA]; }
		if(!space_sleeps(spaces[B.space[a]])) continue;
		bool arb = (c < nA);
		int ta = B.type[a], tb = B.type[b];
		// cpBodyActivate(a) if b is kinematic or a sleeps; same for b
		if(ta == CPB200_BODY_DYNAMIC && (tb == CPB200_BODY_KINEMATIC || (arb && B.sleeping[a]))){
			B.idle[a] = 0.0;
			if(B.sleeping[a]) I.wake[B.sgroup[a]] = 1;
		}
		if(tb == CPB200_BODY_DYNAMIC && (ta == CPB200_BODY_KINEMATIC || (arb && B.sleeping[b]))){
			B.idle[b] = 0.0;
			if(B.sleeping[b]) I.wake[B.sgroup[b]] = 1;
		}
	}
}

__global__ void k_sleep_wake_apply(DBodies B, DIslands I, int par)
{
	if(!I.flags[4*par + ISL_SLEEPING]) return;
	int i = CPB_TID;
	if(i >= B.n || !B.sleeping[i]) return;
	int g = B.sgroup[i];
	if(g >= 0 && I.wake[g]){
		B.sleeping[i] = 0; B.sgroup[i] = -1; B.idle[i] = 0.0;
		I.woken[i] = 1;
		*I.any_woken = 1;
	}
}

CPB_DEVICE int uf_find(int *parent, int x){
	// find with path halving: every visited node is re-pointed at its grandparent.  Racing
	// writers only ever store an ancestor, so the forest stays valid without atomics.
	volatile int *vp = (volatile int *)parent;
	int p = vp[x];
	while(p != x){
		int g = vp[p];
		if(g != p) vp[x] = g;
		x = p; p = g;
	}
	return x;
}

// bodies touching a just-woken body get their idle timer reset (cpSpaceComponent.c:145-151); a pass of its own so
// that the union pass below reads settled idle timers.  Returns at once in the usual step where nothing woke up.
__global__ void k_sleep_woken_reset(DBodies B, DIslands I, DArbs A, DJoints J)
{
	if(!*I.any_woken) return;
	int nA = *A.count_ptr; if(nA > A.cap) nA = A.cap;
	int total = nA + J.n;
	for(int c = CPB_TID; c < total; c += CPB_NTHREADS){
		int a, b;
		if(c < nA){ if(A.active[c] != 1) continue; a = A.ba[c]; b = A.bb[c]; }
		else { a = J.a[c - nA]; b = J.b[c - nA]; }
		if(I.woken[a] && B.type[b] == CPB200_BODY_DYNAMIC) B.idle[b] = 0.0;
		if(I.woken[b] && B.type[a] == CPB200_BODY_DYNAMIC) B.idle[a] = 0.0;
	}
}

// Union over the edges between awake dynamic bodies that have BOTH idled long enough.  A component can only fall
// asleep if every member has (ComponentActive, cpSpaceComponent.c:210-218), i.e. iff it is a component of this
// "idle subgraph" with no edge to a body that has not: such edges only mark their idle endpoint.  While a pile
// is still settling no body qualifies and the pass is a streaming read; the full union-find over millions of
// edges of one giant component runs only in the step in which that component actually falls asleep.
__global__ void k_sleep_union(DBodies B, DIslands I, DArbs A, DJoints J, const DSpace *__restrict__ spaces, int par)
{
	if(!I.flags[4*par + ISL_IDLE]) return;      // nobody has idled long enough: no component can fall asleep this step
	int nA = *A.count_ptr; if(nA > A.cap) nA = A.cap;
	int total = nA + J.n;
	for(int c = CPB_TID; c < total; c += CPB_NTHREADS){
		int a, b;
		if(c < nA){ if(A.active[c] != 1) continue; a = A.ba[c]; b = A.bb[c]; }
		else { a = J.a[c - nA]; b = J.b[c - nA]; }
		if(B.type[a] != CPB200_BODY_DYNAMIC || B.type[b] != CPB200_BODY_DYNAMIC) continue;
		if(B.sleeping[a] || B.sleeping[b]) continue;
		const double thr = spaces[B.space[a]].sleep_threshold;
		const bool ia = (B.idle[a] >= thr), ib = (B.idle[b] >= thr);
		if(!(ia && ib)){
			if(ia) I.touch[a] = 1;
			if(ib) I.touch[b] = 1;
			continue;
		}
		for(;;){
			int ra = uf_find(I.parent, a), rb = uf_find(I.parent, b);
			if(ra == rb) break;
			// randomised linking: the root with the smaller hash goes under the other one, which keeps
			// the expected tree depth logarithmic whatever the body numbering (index-ordered linking
			// degenerates into chains as long as a row of the pile)
			uint64_t ha = mix64((uint64_t)ra), hb = mix64((uint64_t)rb);
			if(ha > hb || (ha == hb && ra < rb)){ int t = ra; ra = rb; rb = t; }
			if(atomicCAS(&I.parent[ra], ra, rb) == ra) break;
		}
	}
}

__global__ void k_sleep_components(DBodies B, DIslands I, const DSpace *__restrict__ spaces, int par)
{
	if(!I.flags[4*par + ISL_IDLE]) return;
	int i = CPB_TID;
	if(i >= B.n) return;
	if(B.type[i] != CPB200_BODY_DYNAMIC || B.sleeping[i]) return;
	int r = uf_find(I.parent, i);
	DSpace sp = spaces[B.space[i]];
	// ComponentActive (cpSpaceComponent.c:210-218)
	if(!space_sleeps(sp) || B.idle[i] < sp.sleep_threshold || I.touch[i]) I.comp_active[r] = 1;
}

__global__ void k_sleep_apply(DBodies B, DIslands I, int par)
{
	if(!I.flags[4*par + ISL_IDLE]) return;      // (and k_sleep_components did not mark anything)
	int i = CPB_TID;
	if(i >= B.n) return;
	if(B.type[i] != CPB200_BODY_DYNAMIC || B.sleeping[i]) return;
	// look the root up again instead of caching it in parent[i]: another thread's path halving may
	// still overwrite parent[i] with a (valid, but non-root) ancestor it read earlier
	int r = uf_find(I.parent, i);
	if(!I.comp_active[r]){ B.sleeping[i] = 1; B.sgroup[i] = r; }
}

// arbiters whose bodies all rest leave the solver but keep their contacts (cpSpaceComponent.c:94-105)
__global__ void k_sleep_arbs(DBodies B, DIslands I, DArbs A, DCounters *C, int par)
{
	// an arbiter goes dormant only if a body of it sleeps: one that slept before this step or one that fell asleep in it
	if(!I.flags[4*par + ISL_SLEEPING] && !I.flags[4*par + ISL_IDLE]) return;
	int nA = *A.count_ptr; if(nA > A.cap) nA = A.cap;
	for(int i = CPB_TID; i < nA; i += CPB_NTHREADS){
		if(A.active[i] != 1) continue;
		int a = A.ba[i], b = A.bb[i];
		bool ra = (B.type[a] == CPB200_BODY_STATIC) || B.sleeping[a];
		bool rb = (B.type[b] == CPB200_BODY_STATIC) || B.sleeping[b];
		if(ra && rb){
			A.active[i] = 2;
			atomicAdd(&C->n_active, -1); atomicAdd(&C->n_contacts, -A.cnt[i]);
		}
	}
}

static int islands_step(DIslands &I, DBodies &B, DShapes &S, DJoints &J, DArbs &Ap, DArbs &Ac, DTable &Tc, const DSpace *spaces, double dt, int par, DCounters *C, int sm_count, cudaStream_t st)
{
	(void)S; (void)Ap; (void)Tc;
	if(B.n == 0) return 0;
	int gb = cpb_div_up(B.n, 256);
	int ge = std::min(cpb_div_up(Ac.cap + J.n + 1, 256), sm_count*8);
	LAUNCH(k_sleep_idle, gb, 256, st, B, I, spaces, dt, par);
	LAUNCH(k_sleep_wake_mark, ge, 256, st, B, I, Ac, J, spaces, par);
	LAUNCH(k_sleep_wake_apply, gb, 256, st, B, I, par);
	LAUNCH(k_sleep_woken_reset, ge, 256, st, B, I, Ac, J);
	LAUNCH(k_sleep_union, ge, 256, st, B, I, Ac, J, spaces, par);
	LAUNCH(k_sleep_components, gb, 256, st, B, I, spaces, par);
	LAUNCH(k_sleep_apply, gb, 256, st, B, I, par);
	LAUNCH(k_sleep_arbs, ge, 256, st, B, I, Ac, C, par);
	return 0;
}
