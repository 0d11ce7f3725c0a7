// k_solve.cuh -- K10 (graph colouring) + K11 (warm start + Gauss-Seidel iterations).
//
// The reference walks space->arbiters then space->constraints sequentially, `iterations` times
// (cpSpaceStep.c:406-427), every constraint reading and writing the velocities of its two
// bodies in place.  Here the constraint graph (nodes = bodies with finite mass, edges = active
// arbiters and joints) is edge-coloured every step so that constraints of one colour touch
// disjoint dynamic bodies; a colour is then solved by independent threads and colours are
// separated by grid-wide barriers inside ONE persistent cooperative kernel.
//
// Colouring: Jones-Plassmann rounds.  Every uncoloured constraint bids for its dynamic bodies
// with atomicMax(round | hash(stable key)); a constraint that holds both bodies takes the lowest
// colour free in both bodies' 64-bit colour masks.  Priorities depend only on the shape-pair key
// / joint index, so the colouring (and therefore the solve) is deterministic run to run.
// Static and kinematic bodies (m_inv = i_inv = 0) are never written and impose no conflict
// (SURVEY.md hard part 8).  Colour 63 is an overflow bucket solved by a single thread.
//
// Serial mode (validation): one thread replays the reference's exact sequential order.
#pragma once
#include "cpb_world.h"
#include "k_joint.cuh"
#include "prims.cuh"

#define CPB_OVERFLOW_COLOUR (CPB_MAX_COLOURS - 1)

#define CPB_MAX_COLOUR_ROUNDS 200

struct DColour {
	unsigned long long *claim;   // [n_bodies]
	unsigned long long *bmask;   // [n_bodies]
	int *ccount, *cstart, *ccursor;   // [CPB_MAX_COLOURS + 1] arbiters per colour
	int *jcount, *jstart, *jcursor;   // [CPB_MAX_COLOURS + 1] joints per colour
	int *wl[2];                  // worklists of still uncoloured constraints (ping-pong per round)
	int *wl_n;                   // [CPB_MAX_COLOUR_ROUNDS + 2] worklist length entering each round
	const uint2 *ids;            // DShapes::ids (hashid, hlocal): priorities of records that carry the ~0 sentinel
	unsigned long long *prof;    // [8] globaltimer stamps of the persistent kernel (start, coloured, rows built, warm start done, end) + rounds
};

// Words that other CTAs update while the persistent kernel runs are read through L2 (ld.cg) and
// updated with atomics, never through a possibly stale L1 line.
#ifndef CPB_EMU
__device__ __forceinline__ unsigned long long ld_u64(const unsigned long long *p){ return __ldcg(p); }
__device__ __forceinline__ double ld_f64(const double *p){ return __ldcg(p); }
#else
static inline unsigned long long ld_u64(const unsigned long long *p){ return *p; }
static inline double ld_f64(const double *p){ return *p; }
#endif

CPB_DEVICE bool body_is_dynamic(const DBodies &B, int b){ V2 mi = B.MI[b]; return mi.x != 0.0 || mi.y != 0.0; }

// unified constraint index c: [0, nA) arbiter records, [nA, nA + nJ) joints
CPB_DEVICE bool cons_fetch(const DArbs &A, const DJoints &J, int nA, int c, int &a, int &b, uint64_t &pri, int &col){
	if(c < nA){
		if(A.active[c] != 1) return false;
		a = A.ba[c]; b = A.bb[c]; col = A.colour[c];
		pri = A.pri[c];
	} else {
		int j = c - nA;
		col = J.colour[j];
		if(col == -2) return false;
		a = J.a[j]; b = J.b[j];
		pri = J.pri[j];
	}
	return true;
}

// priority of a worklist constraint; k_collide<0> leaves ~0 instead of gathering the space-local shape ids for
// every circle pair (only records that must be coloured afresh ever need them)
CPB_DEVICE uint64_t cons_pri(const DArbs &A, const DColour &K, int nA, int c, uint64_t pri){
	if(c < nA && pri == ~0ull){
		uint32_t la = K.ids[A.sa[c]].y, lb = K.ids[A.sb[c]].y;
		uint64_t lo = (la < lb ? la : lb), hi = (la < lb ? lb : la);
		pri = mix64((lo << 32) | hi) >> 8;
	}
	return pri;
}

// per-colour histogram: block-local in shared memory inside the persistent kernel (one global atomic per
// colour per CTA instead of one per constraint on a handful of hot addresses)
CPB_DEVICE void hist_add(int *shist, int *ghist, int colour){
	if(shist) atomicAdd(&shist[colour], 1); else atomicAdd(&ghist[colour], 1);
}

CPB_DEVICE void cons_set_colour(const DArbs &A, const DJoints &J, const DColour &K, int *shist, int nA, int c, int colour){
	if(c < nA){ A.colour[c] = colour; hist_add(shist, K.ccount, colour); }
	else { J.colour[c - nA] = colour; hist_add(shist ? shist + CPB_MAX_COLOURS : NULL, K.jcount, colour); }
}

// Seed: a constraint that was solved last step keeps last step's colour (all of those were mutually
// conflict-free then and still join the same two bodies, so they cannot collide with each other now);
// everything else -- new contacts, re-activated ones, or all of them after a re-upload -- goes on the
// worklist for the Jones-Plassmann rounds.  In a settled pile the worklist is a fraction of a percent.
CPB_DEVICE void colour_seed(const DBodies &B, const DArbs &A, const DJoints &J, const DColour &K, int *shist, int nA, int use_hints, int tid, int nth){
	int total = nA + J.n;
	int rounded = ((total + 31)/32)*32;
	for(int c = tid; c < rounded; c += nth){
		int a = 0, b = 0, col = 0; uint64_t pri = 0;
		bool live = (c < total) && cons_fetch(A, J, nA, c, a, b, pri, col);
		bool queued = false;
		if(live){
			int hint = (c < nA ? A.hint[c] : J.hint[c - nA]);
			if(use_hints && hint >= 0 && hint < CPB_OVERFLOW_COLOUR){
				unsigned long long bit = 1ull << hint;
				if(body_is_dynamic(B, a)) atomicOr(&K.bmask[a], bit);
				if(body_is_dynamic(B, b)) atomicOr(&K.bmask[b], bit);
				cons_set_colour(A, J, K, shist, nA, c, hint);
			} else queued = true;
		}
		int slot = cpb_warp_append(&K.wl_n[0], queued);
		if(queued) K.wl[0][slot] = c;
	}
}

CPB_DEVICE void colour_phase_a(const DBodies &B, const DArbs &A, const DJoints &J, const DColour &K, int nA, int round, int tid, int nth){
	const int *wl = K.wl[round & 1];
	int n = *((volatile int *)&K.wl_n[round]);
	for(int k = tid; k < n; k += nth){
		int c = wl[k];
		int a, b, col; uint64_t pri;
		cons_fetch(A, J, nA, c, a, b, pri, col);
		pri = cons_pri(A, K, nA, c, pri);
		unsigned long long bid = ((unsigned long long)(round + 1) << 56) | pri;
		if(body_is_dynamic(B, a)) atomicMax(&K.claim[a], bid);
		if(body_is_dynamic(B, b)) atomicMax(&K.claim[b], bid);
	}
}

CPB_DEVICE void colour_phase_b(const DBodies &B, const DArbs &A, const DJoints &J, const DColour &K, int *shist, int nA, int round, int tid, int nth){
	const int *wl = K.wl[round & 1];
	int *next = K.wl[(round + 1) & 1];
	int n = *((volatile int *)&K.wl_n[round]);
	int rounded = ((n + 31)/32)*32;
	for(int k = tid; k < rounded; k += nth){
		bool lost = false;
		int c = 0;
		if(k < n){
			c = wl[k];
			int a, b, col; uint64_t pri;
			cons_fetch(A, J, nA, c, a, b, pri, col);
			pri = cons_pri(A, K, nA, c, pri);
			unsigned long long bid = ((unsigned long long)(round + 1) << 56) | pri;
			bool da = body_is_dynamic(B, a), db = body_is_dynamic(B, b);
			bool win = (!da || ld_u64(&K.claim[a]) == bid) && (!db || ld_u64(&K.claim[b]) == bid);
			if(win){
				unsigned long long m = (da ? ld_u64(&K.bmask[a]) : 0ull) | (db ? ld_u64(&K.bmask[b]) : 0ull);
				unsigned long long freebits = ~m & ((1ull << CPB_OVERFLOW_COLOUR) - 1ull);
				int colour = (freebits ? __ffsll((long long)freebits) - 1 : CPB_OVERFLOW_COLOUR);
				if(colour != CPB_OVERFLOW_COLOUR){
					unsigned long long bit = 1ull << colour;
					if(da) atomicOr(&K.bmask[a], bit);
					if(db) atomicOr(&K.bmask[b], bit);
				}
				cons_set_colour(A, J, K, shist, nA, c, colour);
			} else lost = true;
		}
		int slot = cpb_warp_append(&K.wl_n[round + 1], lost);
		if(lost) next[slot] = c;
	}
}

// Whatever is still on the worklist when the rounds run out goes to the overflow bucket (solved by one thread after the
// regular colours of every pass): slower, never dropped.
CPB_DEVICE void colour_leftover(const DArbs &A, const DJoints &J, const DColour &K, int *shist, int nA, int round, int tid, int nth){
	const int *wl = K.wl[round & 1];
	int n = *((volatile int *)&K.wl_n[round]);
	for(int k = tid; k < n; k += nth) cons_set_colour(A, J, K, shist, nA, wl[k], CPB_OVERFLOW_COLOUR);
}

// exclusive prefix of the per-colour counts (single thread; 64 entries) + number of colours in use
CPB_DEVICE void colour_starts(const DColour &K, DCounters *C){
	int ra = 0, rj = 0, ncol = 0;
	for(int c = 0; c <= CPB_MAX_COLOURS; c++){
		int na = (c < CPB_MAX_COLOURS ? K.ccount[c] : 0), nj = (c < CPB_MAX_COLOURS ? K.jcount[c] : 0);
		K.cstart[c] = ra; K.jstart[c] = rj;
		K.ccursor[c] = 0; K.jcursor[c] = 0;
		ra += na; rj += nj;
		if(na + nj > 0) ncol = c + 1;
	}
	C->n_colours = ncol;
}

// A production step folds K8 (cpArbiterPreStep) into the row build: the row build reads every active record anyway, so
// nMass / tMass / bias / bounce go straight from registers into the row instead of through five record arrays and back
// (k_arb_prestep's round trip: 5 doubles per contact written, then gathered again 0.3 ms later).  on == 0: the numbers
// are in the records already (split steps -- collision handlers, validation hooks --, the serial solver, the emulator).
struct DPrestep { const DSpace *spaces; double dt; int on; };

CPB_DEVICE void write_row(const DBodies &B, const DPrestep &P, const DArbs &A, const DRows &R, int i, int r){
	if(r >= R.cap) return;
	// gather everything first, then scatter: the record and row arrays may alias as far as the compiler knows, so
	// interleaved copies would serialise into load -> store -> load chains (one memory latency per field)
	const int ba = A.ba[i], bb = A.bb[i], cnt = A.cnt[i], state = A.state[i];
	const V2 n = A.n[i], svr = A.svr[i];
	const double u = A.u[i];
	const int s0 = CIDX(A, i, 0), s1 = CIDX(A, i, 1);
	const V2 r1a = A.r1[s0], r2a = A.r2[s0];
	double nma, tma, boa, bia, jba;
	const double jna = A.jn[s0], jta = A.jt[s0];
	V2 r1b = r1a, r2b = r2a;
	double nmb = 0.0, tmb = 0.0, bob = 0.0, bib = 0.0, jnb = 0.0, jtb = 0.0, jbb = 0.0;
	if(cnt == 2){ r1b = A.r1[s1]; r2b = A.r2[s1]; jnb = A.jn[s1]; jtb = A.jt[s1]; }
	if(P.on){
		const DSpace sp = P.spaces[B.space[ba]];
		const PrestepBodies pb = prestep_bodies(B, ba, bb);
		const double e = A.e[i];
		prestep_contact(pb, n, e, sp.bias_coef, sp.slop, P.dt, r1a, r2a, nma, tma, bia, boa);
		jba = 0.0;
		if(cnt == 2) prestep_contact(pb, n, e, sp.bias_coef, sp.slop, P.dt, r1b, r2b, nmb, tmb, bib, bob);
	} else {
		nma = A.nmass[s0]; tma = A.tmass[s0]; boa = A.bounce[s0]; bia = A.bias[s0]; jba = A.jb[s0];
		if(cnt == 2){ nmb = A.nmass[s1]; tmb = A.tmass[s1]; bob = A.bounce[s1]; bib = A.bias[s1]; jbb = A.jb[s1]; }
	}
	R.arb[r] = i; R.ba[r] = ba; R.bb[r] = bb;
	// first-collision arbiters skip the warm start (cpArbiter.c:444): flag in the sign of cnt
	R.cnt[r] = (state == CPB200_ARB_FIRST_COLLISION ? -cnt : cnt);
	R.n[r] = n; R.svr[r] = svr; R.u[r] = u;
	if(cnt >= 1){
		R.r1[r] = r1a; R.r2[r] = r2a;
		R.nmass[r] = nma; R.tmass[r] = tma; R.bounce[r] = boa; R.bias[r] = bia;
		R.jn[r] = jna; R.jt[r] = jta; R.jb[r] = jba;
	}
	if(cnt == 2){
		const int d = R.cap + r;
		R.r1[d] = r1b; R.r2[d] = r2b;
		R.nmass[d] = nmb; R.tmass[d] = tmb; R.bounce[d] = bob; R.bias[d] = bib;
		R.jn[d] = jnb; R.jt[d] = jtb; R.jb[d] = jbb;
	}
}

// the same row in the packed layout of the space-local solver
CPB_DEVICE void write_row_packed(const DBodies &B, const DPrestep &P, const DArbs &A, const DRows &R, int i, int r){
	if(r >= R.cap) return;
	const int ba = A.ba[i], bb = A.bb[i], cnt = A.cnt[i], state = A.state[i];
	const V2 n = A.n[i], svr = A.svr[i];
	const double u = A.u[i];
	const int s0 = CIDX(A, i, 0), s1 = CIDX(A, i, 1);
	const V2 r1a = A.r1[s0], r2a = A.r2[s0];
	double nma, tma, boa, bia, jba;
	const double jna = A.jn[s0], jta = A.jt[s0];
	V2 r1b = r1a, r2b = r2a;
	double nmb = 0.0, tmb = 0.0, bob = 0.0, bib = 0.0, jnb = 0.0, jtb = 0.0, jbb = 0.0;
	if(cnt == 2){ r1b = A.r1[s1]; r2b = A.r2[s1]; jnb = A.jn[s1]; jtb = A.jt[s1]; }
	if(P.on){
		const DSpace sp = P.spaces[B.space[ba]];
		const PrestepBodies pb = prestep_bodies(B, ba, bb);
		const double e = A.e[i];
		prestep_contact(pb, n, e, sp.bias_coef, sp.slop, P.dt, r1a, r2a, nma, tma, bia, boa);
		jba = 0.0;
		if(cnt == 2) prestep_contact(pb, n, e, sp.bias_coef, sp.slop, P.dt, r1b, r2b, nmb, tmb, bib, bob);
	} else {
		nma = A.nmass[s0]; tma = A.tmass[s0]; boa = A.bounce[s0]; bia = A.bias[s0]; jba = A.jb[s0];
		if(cnt == 2){ nmb = A.nmass[s1]; tmb = A.tmass[s1]; bob = A.bounce[s1]; bib = A.bias[s1]; jbb = A.jb[s1]; }
	}
	R.hdr[r] = make_int4(ba, bb, (state == CPB200_ARB_FIRST_COLLISION ? -cnt : cnt), i);
	R.nsv[r] = make_double4(n.x, n.y, svr.x, svr.y);
	if(cnt >= 1){
		R.r12[r] = make_double4(r1a.x, r1a.y, r2a.x, r2a.y);
		R.mass[r] = make_double4(nma, tma, bia, boa);
		R.imp[r] = make_double4(jna, jta, jba, u);
	}
	if(cnt == 2){
		const int d = R.cap + r;
		R.r12[d] = make_double4(r1b.x, r1b.y, r2b.x, r2b.y);
		R.mass[d] = make_double4(nmb, tmb, bib, bob);
		R.imp[d] = make_double4(jnb, jtb, jbb, u);
	}
}

// scatter arbiter records into colour-sorted SoA rows (the solver's coalesced working set).
// scnt/sbase: per-CTA shared scratch [CPB_MAX_COLOURS] (NULL in the emulation build): a CTA counts its
// rows per colour, reserves one contiguous range per colour with a single global atomic, then hands the
// slots out with shared-memory atomics.
CPB_DEVICE void build_rows(const DBodies &B, const DPrestep &P, const DArbs &A, const DJoints &J, const DRows &R, const DColour &K, int *scnt, int *sbase, int nA, int tid, int nth, bool chunked){
#ifndef CPB_EMU
	if(scnt){
		// every CTA takes ONE contiguous chunk of the records (they were appended in pair-list order, i.e. along the
		// broadphase's Morton order): the rows it reserves per colour then come from one neighbourhood of the scene, and a warp
		// of the iteration loop gathers body sectors that lie close together.  (Strided over the whole record range, 256
		// records at a time, the same loop interleaved ~40 distant neighbourhoods inside every CTA's slice of a colour.)
		const int per = (nA + (int)gridDim.x - 1)/(int)gridDim.x;
		const int i0 = (chunked ? (int)blockIdx.x*per + (int)threadIdx.x : tid), i1 = (chunked ? min(nA, ((int)blockIdx.x + 1)*per) : nA);
		const int istep = (chunked ? (int)blockDim.x : nth);
		for(int i = i0; i < i1; i += istep){
			if(A.active[i] != 1) continue;
			int col = A.colour[i];
			if(col >= 0) atomicAdd(&scnt[col], 1);
		}
		__syncthreads();
		if(threadIdx.x < CPB_MAX_COLOURS){
			int n = scnt[threadIdx.x];
			sbase[threadIdx.x] = (n ? atomicAdd(&K.ccursor[threadIdx.x], n) : 0);
			scnt[threadIdx.x] = 0;
		}
		__syncthreads();
		for(int i = i0; i < i1; i += istep){
			if(A.active[i] != 1){
				// (the stand-alone K8 kernel's other duty, cpSpaceStep.c:283)
				if(P.on && A.active[i] == 0 && A.state[i] == CPB200_ARB_FIRST_COLLISION) A.state[i] = CPB200_ARB_NORMAL;
				continue;
			}
			int col = A.colour[i];
			if(col < 0) continue;
			write_row(B, P, A, R, i, K.cstart[col] + sbase[col] + atomicAdd(&scnt[col], 1));
		}
		__syncthreads();
		if(threadIdx.x < CPB_MAX_COLOURS) scnt[threadIdx.x] = 0;
	} else
#endif
	{
		for(int i = tid; i < nA; i += nth){
			if(A.active[i] != 1){
				if(P.on && A.active[i] == 0 && A.state[i] == CPB200_ARB_FIRST_COLLISION) A.state[i] = CPB200_ARB_NORMAL;
				continue;
			}
			int col = A.colour[i];
			if(col < 0) continue;
			write_row(B, P, A, R, i, K.cstart[col] + atomicAdd(&K.ccursor[col], 1));
		}
	}
	for(int j = tid; j < J.n; j += nth){
		int col = J.colour[j];
		if(col < 0) continue;
		J.row[K.jstart[col] + atomicAdd(&K.jcursor[col], 1)] = j;
	}
}

// ---- contact math shared by the coloured and the serial solver ----
// cpArbiterApplyCachedImpulse for one contact (cpArbiter.c:441-455)
CPB_DEVICE void contact_apply_cached(double4 &Va, double4 &Vb, V2 mia, V2 mib, V2 n, V2 r1, V2 r2, double jn, double jt, double dt_coef){
	V2 j = vrotate(n, v2(jn, jt));
	apply_impulses(Va, Vb, mia, mib, r1, r2, vmul(j, dt_coef));
}

// apply_impulse (chipmunk_private.h:185-189) that also reports whether any bit of the body's sector changed: the
// solver skips the scatter of an untouched sector, and comparing each component as it is replaced costs no registers
// (a before/after snapshot of the four sectors cost 24)
CPB_DEVICE void apply_impulse_track(double4 &V, V2 mi, V2 j, V2 r, bool &changed){
	const double x = V.x + j.x*mi.x, y = V.y + j.y*mi.x;
	double z = V.z; z += mi.y*vcross(r, j);
	changed = changed || __double_as_longlong(x) != __double_as_longlong(V.x) || __double_as_longlong(y) != __double_as_longlong(V.y) || __double_as_longlong(z) != __double_as_longlong(V.z);
	V.x = x; V.y = y; V.z = z;
}
CPB_DEVICE void apply_impulses_track(double4 &Va, double4 &Vb, V2 mia, V2 mib, V2 r1, V2 r2, V2 j, bool &cha, bool &chb){
	apply_impulse_track(Va, mia, vneg(j), r1, cha);
	apply_impulse_track(Vb, mib, j, r2, chb);
}

// which of a row's four body sectors the solve changed
struct RowDirty { bool va, vb, vba, vbb; };

// cpArbiterApplyImpulse for one contact (cpArbiter.c:459-498)
CPB_DEVICE void contact_apply(double4 &Va, double4 &Vb, double4 &VBa, double4 &VBb, V2 mia, V2 mib,
	V2 n, V2 surface_vr, double friction, V2 r1, V2 r2, double nMass, double tMass, double bias, double bounce,
	double &jnAcc, double &jtAcc, double &jBias, RowDirty &dirty)
{
	V2 vb1 = vadd(v2(VBa.x, VBa.y), vmul(vperp(r1), VBa.z));
	V2 vb2 = vadd(v2(VBb.x, VBb.y), vmul(vperp(r2), VBb.z));
	V2 vr = vadd(relative_velocity(Va, Vb, r1, r2), surface_vr);

	double vbn = vdot(vsub(vb2, vb1), n);
	double vrn = vdot(vr, n);
	double vrt = vdot(vr, vperp(n));

	double jbn = (bias - vbn)*nMass;
	double jbnOld = jBias;
	jBias = fmax_cp(jbnOld + jbn, 0.0);

	double jn = -(bounce + vrn)*nMass;
	double jnOld = jnAcc;
	jnAcc = fmax_cp(jnOld + jn, 0.0);

	double jtMax = friction*jnAcc;
	double jt = -vrt*tMass;
	double jtOld = jtAcc;
	jtAcc = fclamp_cp(jtOld + jt, -jtMax, jtMax);

	apply_impulses_track(VBa, VBb, mia, mib, r1, r2, vmul(n, jBias - jbnOld), dirty.vba, dirty.vbb);
	apply_impulses_track(Va, Vb, mia, mib, r1, r2, vrotate(n, v2(jnAcc - jnOld, jtAcc - jtOld)), dirty.va, dirty.vb);
}

// Body velocities are gathered/scattered once per colour phase and are written by other SMs in
// the previous phase: go through L2 (ld.cg / st.cg), never through the non-coherent L1.
CPB_DEVICE double4 ld_vel(const double4 *p){ return ld4_cg(p); }
CPB_DEVICE void st_vel(double4 *p, double4 v){ st4_cg(p, v); }

// Where the solver keeps the two velocity sectors of a body: in global memory behind L2 (one world-wide
// constraint graph, any CTA may touch any body) or in the shared memory of the CTA that owns the whole space
// (space-local solver below; bodies of a space are contiguous, index = body - first body of the space).
#ifdef CPB_EMU
#define CPB_MEMBER inline
#else
#define CPB_MEMBER __device__ __forceinline__
#endif
// STREAM_ROWS: rows larger than the L2 can keep next to the velocities go past it (evict-first, see ROW_LD);
// smaller row sets (mid-size scenes) use the default policy and are served from L2 from the second pass on.
template<bool STREAM_ROWS> struct VelGlobalT {
	static const bool STREAM = STREAM_ROWS;
	double4 *V, *VB;
	CPB_MEMBER double4 ldV(int b) const { return ld_vel(&V[b]); }
	CPB_MEMBER double4 ldVB(int b) const { return ld_vel(&VB[b]); }
	CPB_MEMBER void stV(int b, double4 v) const { st_vel(&V[b], v); }
	CPB_MEMBER void stVB(int b, double4 v) const { st_vel(&VB[b], v); }
};
typedef VelGlobalT<true> VelGlobal;
struct VelShared {
	static const bool STREAM = false;    // a space's rows are re-read by the same SM every pass: cache them
	double4 *V, *VB; int b0;
	CPB_MEMBER double4 ldV(int b) const { return V[b - b0]; }
	CPB_MEMBER double4 ldVB(int b) const { return VB[b - b0]; }
	CPB_MEMBER void stV(int b, double4 v) const { V[b - b0] = v; }
	CPB_MEMBER void stVB(int b, double4 v) const { VB[b - b0] = v; }
};

// Solver rows are streamed once per pass and never reused before the next pass evicts them: load/store them
// with the evict-first policy (ld.global.cs / st.global.cs) so the ~80 MB of body velocities that every
// pass gathers again stays resident in the 126 MB L2.  (A run-time createpolicy/L2::cache_hint variant that
// keeps small row sets L2-resident measured 9% slower on both the 1M pile and the batched spaces.)
#if !defined(CPB_EMU) && !defined(CPB_NO_ROW_STREAM)
#define ROW_LD(p) __ldcs(p)
#define ROW_ST(p, v) __stcs((p), (v))
#else
#define ROW_LD(p) (*(p))
#define ROW_ST(p, v) (*(p) = (v))
#endif

#ifndef CPB_EMU
template<bool STREAM, class T> __device__ __forceinline__ T row_ld(const T *p){ if(STREAM) return __ldcs(p); else return *p; }
template<bool STREAM, class T> __device__ __forceinline__ void row_st(T *p, T v){ if(STREAM) __stcs(p, v); else *p = v; }
#else
template<bool STREAM, class T> static inline T row_ld(const T *p){ return *p; }
template<bool STREAM, class T> static inline void row_st(T *p, T v){ *p = v; }
#endif

CPB_DEVICE bool same_bits(double a, double b){ return __double_as_longlong(a) == __double_as_longlong(b); }
CPB_DEVICE bool same_bits3(double4 a, double4 b){ return same_bits(a.x, b.x) && same_bits(a.y, b.y) && same_bits(a.z, b.z); }

// Optional L2 prefetch of a row's first-contact fields one row ahead (-DCPB_ROW_PREFETCH).  Measured on the
// 1M pile: 13% SLOWER iterations -- the solver is bound by request throughput of the memory system, not by
// the latency of an individual row, so extra requests hurt.  Kept as an experiment switch only.
#if !defined(CPB_EMU) && defined(CPB_ROW_PREFETCH)
__device__ __forceinline__ void pf_l2(const void *p){ asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__device__ __forceinline__ void row_prefetch(const DRows &R, int r){
	pf_l2(&R.n[r]); pf_l2(&R.svr[r]); pf_l2(&R.u[r]); pf_l2(&R.r1[r]); pf_l2(&R.r2[r]);
	pf_l2(&R.nmass[r]); pf_l2(&R.tmass[r]); pf_l2(&R.bias[r]); pf_l2(&R.bounce[r]);
	pf_l2(&R.jn[r]); pf_l2(&R.jt[r]); pf_l2(&R.jb[r]);
}
#else
CPB_DEVICE void row_prefetch(const DRows &, int){ }
#endif
// The first row a thread solves in the NEXT colour phase: its constants and impulses belong to this thread alone, so they
// can be pulled into L2 while the grid barrier is still closed (one prefetch burst per thread and phase, unlike the
// per-row variant above); after the barrier only the velocity gathers are left to wait for.
#ifndef CPB_EMU
__device__ __forceinline__ void row_prefetch_phase(const DRows &R, int r){
	asm volatile("prefetch.global.L2 [%0];" :: "l"(&R.n[r])); asm volatile("prefetch.global.L2 [%0];" :: "l"(&R.r1[r])); asm volatile("prefetch.global.L2 [%0];" :: "l"(&R.r2[r]));
	asm volatile("prefetch.global.L2 [%0];" :: "l"(&R.nmass[r])); asm volatile("prefetch.global.L2 [%0];" :: "l"(&R.tmass[r])); asm volatile("prefetch.global.L2 [%0];" :: "l"(&R.bias[r]));
	asm volatile("prefetch.global.L2 [%0];" :: "l"(&R.jn[r])); asm volatile("prefetch.global.L2 [%0];" :: "l"(&R.jt[r])); asm volatile("prefetch.global.L2 [%0];" :: "l"(&R.jb[r]));
}
#else
static inline void row_prefetch_phase(const DRows &, int){ }
#endif

// one contact of a row: constants + accumulated impulses
template<bool STREAM> struct RowContact {
	V2 r1, r2; double nmass, tmass, bias, bounce, jn, jt, jb;
	CPB_MEMBER void load(const DRows &R, int d){
		r1 = row_ld<STREAM>(&R.r1[d]); r2 = row_ld<STREAM>(&R.r2[d]);
		nmass = row_ld<STREAM>(&R.nmass[d]); tmass = row_ld<STREAM>(&R.tmass[d]); bias = row_ld<STREAM>(&R.bias[d]); bounce = row_ld<STREAM>(&R.bounce[d]);
		jn = row_ld<STREAM>(&R.jn[d]); jt = row_ld<STREAM>(&R.jt[d]); jb = row_ld<STREAM>(&R.jb[d]);
	}
	CPB_MEMBER void solve(const DRows &R, int d, double4 &Va, double4 &Vb, double4 &VBa, double4 &VBb, V2 mia, V2 mib, V2 n, V2 svr, double u, RowDirty &dirty){
#ifndef CPB_NO_SKIP_SAME
		const double jn0 = jn, jt0 = jt, jb0 = jb;
#endif
		contact_apply(Va, Vb, VBa, VBb, mia, mib, n, svr, u, r1, r2, nmass, tmass, bias, bounce, jn, jt, jb, dirty);
#ifndef CPB_NO_SKIP_SAME
		if(!same_bits(jn, jn0)) row_st<STREAM>(&R.jn[d], jn);
		if(!same_bits(jt, jt0)) row_st<STREAM>(&R.jt[d], jt);
		if(!same_bits(jb, jb0)) row_st<STREAM>(&R.jb[d], jb);
#else
		row_st<STREAM>(&R.jn[d], jn); row_st<STREAM>(&R.jt[d], jt); row_st<STREAM>(&R.jb[d], jb);
#endif
	}
};

// one colour-sorted row: mode 0 = warm start, 1 = iteration
// (ba, bb, cnt) are passed in: the persistent kernel fetches them for a thread's next row while the previous
// phase is still draining, so that after the barrier the velocity gathers do not wait for an index load
// returns true if the row changed a body (iteration passes; warm start always counts as changed)
template<class VS> CPB_DEVICE bool solve_row_idx(const VS &vs, const DBodies &B, const DRows &R, int r, int ba, int bb, int cnt, int mode, double dt_coef){
	bool first = (cnt < 0);
	if(first) cnt = -cnt;
	if(mode == 0 && first) return true;
	double4 Va = vs.ldV(ba), Vb = vs.ldV(bb);
	V2 n = row_ld<VS::STREAM>(&R.n[r]);
	if(mode == 0){
		V2 mia = B.MI[ba], mib = B.MI[bb];
		bool dyn_a = (mia.x != 0.0 || mia.y != 0.0), dyn_b = (mib.x != 0.0 || mib.y != 0.0);
		for(int k = 0; k < cnt; k++){
			int d = k*R.cap + r;
			contact_apply_cached(Va, Vb, mia, mib, n, row_ld<VS::STREAM>(&R.r1[d]), row_ld<VS::STREAM>(&R.r2[d]), row_ld<VS::STREAM>(&R.jn[d]), row_ld<VS::STREAM>(&R.jt[d]), dt_coef);
		}
		if(dyn_a) vs.stV(ba, Va);
		if(dyn_b) vs.stV(bb, Vb);
		return true;
	}
	double4 VBa = vs.ldVB(ba), VBb = vs.ldVB(bb);
	// (m_inv, i_inv) ride in the spare lanes of the two velocity sectors
	V2 mia = v2(Va.w, VBa.w), mib = v2(Vb.w, VBb.w);
	bool dyn_a = (mia.x != 0.0 || mia.y != 0.0), dyn_b = (mib.x != 0.0 || mib.y != 0.0);
	V2 svr = row_ld<VS::STREAM>(&R.svr[r]);
	double u = row_ld<VS::STREAM>(&R.u[r]);
	// A contact that does not push this iteration (clamped impulses) leaves both bodies exactly as they
	// were: skip the scatter then.  Bitwise comparison, so the stored state is identical either way.
	RowDirty dirty = {false, false, false, false};
	for(int k = 0; k < cnt; k++){
		RowContact<VS::STREAM> c;
		c.load(R, k*R.cap + r);
		c.solve(R, k*R.cap + r, Va, Vb, VBa, VBb, mia, mib, n, svr, u, dirty);
	}
#ifndef CPB_NO_SKIP_SAME
	if(dyn_a){ if(dirty.va) vs.stV(ba, Va); if(dirty.vba) vs.stVB(ba, VBa); }
	if(dyn_b){ if(dirty.vb) vs.stV(bb, Vb); if(dirty.vbb) vs.stVB(bb, VBb); }
#else
	if(dyn_a){ vs.stV(ba, Va); vs.stVB(ba, VBa); }
	if(dyn_b){ vs.stV(bb, Vb); vs.stVB(bb, VBb); }
#endif
	return dirty.va || dirty.vb || dirty.vba || dirty.vbb;
}

CPB_DEVICE void solve_row(const DBodies &B, const DRows &R, int r, int mode, double dt_coef){
	int ba = ROW_LD(&R.ba[r]), bb = ROW_LD(&R.bb[r]);
	int cnt = ROW_LD(&R.cnt[r]);
	VelGlobal vs = {B.V, B.VB};
	solve_row_idx(vs, B, R, r, ba, bb, cnt, mode, dt_coef);
}

template<class VS> CPB_DEVICE void solve_joint_idx(const VS &vs, const DBodies &B, const DJoints &J, int j, int a, int b, int mode, double dt, double dt_coef){
	V2 mia = B.MI[a], mib = B.MI[b];
	double4 Va = vs.ldV(a), Vb = vs.ldV(b);
	if(mode == 0) joint_apply_cached(J, j, Va, Vb, mia, mib, dt_coef);
	else joint_apply(J, j, Va, Vb, mia, mib, dt);
	if(mia.x != 0.0 || mia.y != 0.0) vs.stV(a, Va);
	if(mib.x != 0.0 || mib.y != 0.0) vs.stV(b, Vb);
}

CPB_DEVICE void solve_joint(const DBodies &B, const DJoints &J, int j, int mode, double dt, double dt_coef){
	VelGlobal vs = {B.V, B.VB};
	solve_joint_idx(vs, B, J, j, J.a[j], J.b[j], mode, dt, dt_coef);
}

// all rows + joints of one colour
CPB_DEVICE void solve_colour(const DBodies &B, const DRows &R, const DJoints &J, const DColour &K, int colour, int mode, double dt, double dt_coef, int tid, int nth){
	int r0 = K.cstart[colour], r1 = K.cstart[colour + 1];
	for(int r = r0 + tid; r < r1; r += nth) solve_row(B, R, r, mode, dt_coef);
	int j0 = K.jstart[colour], j1 = K.jstart[colour + 1];
	for(int q = j0 + tid; q < j1; q += nth) solve_joint(B, J, J.row[q], mode, dt, dt_coef);
}

// overflow bucket: sequential, one thread
template<bool JOINTS> CPB_DEVICE void solve_overflow(const DBodies &B, const DRows &R, const DJoints &J, const DColour &K, int mode, double dt, double dt_coef){
	int r0 = K.cstart[CPB_OVERFLOW_COLOUR], r1 = K.cstart[CPB_OVERFLOW_COLOUR + 1];
	for(int r = r0; r < r1; r++) solve_row(B, R, r, mode, dt_coef);
	if(!JOINTS) return;
	int j0 = K.jstart[CPB_OVERFLOW_COLOUR], j1 = K.jstart[CPB_OVERFLOW_COLOUR + 1];
	for(int q = j0; q < j1; q++) solve_joint(B, J, J.row[q], mode, dt, dt_coef);
}

// copy the accumulated impulses back into the persistent arbiter records
CPB_DEVICE void rows_writeback(const DArbs &A, const DRows &R, int n_rows, int tid, int nth){
	for(int r = tid; r < n_rows; r += nth){
		int i = R.arb[r];
		int cnt = R.cnt[r]; if(cnt < 0) cnt = -cnt;
		// loads first, stores after (the compiler must assume the arrays alias and would chain them otherwise)
		const double jn0 = ld_f64(&R.jn[r]), jt0 = ld_f64(&R.jt[r]), jb0 = ld_f64(&R.jb[r]);
		double jn1 = 0.0, jt1 = 0.0, jb1 = 0.0;
		if(cnt == 2){ const int d = R.cap + r; jn1 = ld_f64(&R.jn[d]); jt1 = ld_f64(&R.jt[d]); jb1 = ld_f64(&R.jb[d]); }
		if(cnt >= 1){ const int s = CIDX(A, i, 0); A.jn[s] = jn0; A.jt[s] = jt0; A.jb[s] = jb0; }
		if(cnt == 2){ const int s = CIDX(A, i, 1); A.jn[s] = jn1; A.jt[s] = jt1; A.jb[s] = jb1; }
	}
}

// One row of the packed layout against shared-memory velocities (space-local solver): every load of the row --
// both contacts -- is issued up front (one memory latency per row; the compiler cannot hoist the second contact's
// loads over the first one's stores itself), the impulse vector is written back only if it changed.
CPB_DEVICE void solve_row_packed(const VelShared &vs, const DRows &R, int r, int ba, int bb, int cnt, int mode, double dt_coef){
	const bool first = (cnt < 0);
	if(first) cnt = -cnt;
	if(mode == 0 && first) return;
	const double4 nsv = ld4_ca(&R.nsv[r]);
	const double4 r12a = ld4_ca(&R.r12[r]), impa = ld4_ca(&R.imp[r]);
	double4 r12b = r12a, impb = impa, massa = impa, massb = impa;
	if(cnt == 2){ r12b = ld4_ca(&R.r12[R.cap + r]); impb = ld4_ca(&R.imp[R.cap + r]); }
	if(mode != 0){ massa = ld4_ca(&R.mass[r]); if(cnt == 2) massb = ld4_ca(&R.mass[R.cap + r]); }
	double4 Va = vs.ldV(ba), Vb = vs.ldV(bb);
	double4 VBa = vs.ldVB(ba), VBb = vs.ldVB(bb);
	// (m_inv, i_inv) ride in the spare lanes of the two velocity sectors
	const V2 mia = v2(Va.w, VBa.w), mib = v2(Vb.w, VBb.w);
	const bool dyn_a = (mia.x != 0.0 || mia.y != 0.0), dyn_b = (mib.x != 0.0 || mib.y != 0.0);
	const V2 n = v2(nsv.x, nsv.y);
	if(mode == 0){
		contact_apply_cached(Va, Vb, mia, mib, n, v2(r12a.x, r12a.y), v2(r12a.z, r12a.w), impa.x, impa.y, dt_coef);
		if(cnt == 2) contact_apply_cached(Va, Vb, mia, mib, n, v2(r12b.x, r12b.y), v2(r12b.z, r12b.w), impb.x, impb.y, dt_coef);
		if(dyn_a) vs.stV(ba, Va);
		if(dyn_b) vs.stV(bb, Vb);
		return;
	}
	const V2 svr = v2(nsv.z, nsv.w);
	const double u = impa.w;
	RowDirty dirty = {false, false, false, false};
	{
		double jn = impa.x, jt = impa.y, jb = impa.z;
		contact_apply(Va, Vb, VBa, VBb, mia, mib, n, svr, u, v2(r12a.x, r12a.y), v2(r12a.z, r12a.w), massa.x, massa.y, massa.z, massa.w, jn, jt, jb, dirty);
		if(!(same_bits(jn, impa.x) && same_bits(jt, impa.y) && same_bits(jb, impa.z))) st4_wb(&R.imp[r], make_double4(jn, jt, jb, impa.w));
	}
	if(cnt == 2){
		double jn = impb.x, jt = impb.y, jb = impb.z;
		contact_apply(Va, Vb, VBa, VBb, mia, mib, n, svr, u, v2(r12b.x, r12b.y), v2(r12b.z, r12b.w), massb.x, massb.y, massb.z, massb.w, jn, jt, jb, dirty);
		if(!(same_bits(jn, impb.x) && same_bits(jt, impb.y) && same_bits(jb, impb.z))) st4_wb(&R.imp[R.cap + r], make_double4(jn, jt, jb, impb.w));
	}
	if(dyn_a){ vs.stV(ba, Va); vs.stVB(ba, VBa); }
	if(dyn_b){ vs.stV(bb, Vb); vs.stVB(bb, VBb); }
}

// ---- space-local solver: batched worlds of many small spaces -------------------------------------------------
// Spaces never interact, so a space that fits one CTA does not need the grid: its bodies' velocity sectors
// live in shared memory for the whole solve and the colours are separated by __syncthreads (no L2 round trip,
// no waiting for the slowest CTA of the device).  The colouring is still the world-wide one above; the rows
// are sorted by (space, colour) instead of colour:
//   k_colour_solve (SL mode)  colours, then histograms the buckets  start[space*64 + colour]
//   exclusive scan            bucket begins
//   k_sl_rows                 rows / joint list in bucket order; the atomic cursor IS the bucket word, which
//                             therefore ends up holding the bucket END = the next bucket's begin
//   k_sl_solve                one CTA per space: warm start + iterations + write-back
// Bucket layout: [n_spaces*64 arbiter buckets][1 sentinel, always empty][n_spaces*64 joint buckets]; after the
// scan the sentinel holds the number of rows, the base of the joint list.
struct DSpaceLocal {
	int n_spaces;
	const int *body0, *nbody;    // [n_spaces] contiguous body range of each space
	uint32_t *start;             // [2*n_spaces*CPB_MAX_COLOURS + 2]; NULL = mode off
};
CPB_DEVICE int sl_arb_bucket(int space, int colour){ return space*CPB_MAX_COLOURS + colour; }
CPB_DEVICE int sl_joint_bucket(const DSpaceLocal &SL, int space, int colour){ return SL.n_spaces*CPB_MAX_COLOURS + 1 + space*CPB_MAX_COLOURS + colour; }

CPB_DEVICE void sl_count(const DBodies &B, const DArbs &A, const DJoints &J, const DSpaceLocal &SL, int nA, int tid, int nth){
	for(int i = tid; i < nA; i += nth){
		if(A.active[i] != 1) continue;
		int col = A.colour[i];
		if(col >= 0) atomicAdd(&SL.start[sl_arb_bucket(B.space[A.ba[i]], col)], 1u);
	}
	for(int j = tid; j < J.n; j += nth){
		int col = J.colour[j];
		if(col >= 0) atomicAdd(&SL.start[sl_joint_bucket(SL, B.space[J.a[j]], col)], 1u);
	}
}

__global__ void k_sl_rows(DBodies B, DArbs A, DJoints J, DRows R, DSpaceLocal SL, DPrestep P)
{
	int nA = *A.count_ptr; if(nA > A.cap) nA = A.cap;
	const int tid = CPB_TID, nth = CPB_NTHREADS;
	const int jbase = (int)SL.start[SL.n_spaces*CPB_MAX_COLOURS];   // sentinel: never incremented
	for(int i = tid; i < nA; i += nth){
		if(A.active[i] != 1){
			if(P.on && A.active[i] == 0 && A.state[i] == CPB200_ARB_FIRST_COLLISION) A.state[i] = CPB200_ARB_NORMAL;
			continue;
		}
		int col = A.colour[i];
		if(col < 0) continue;
		write_row_packed(B, P, A, R, i, (int)atomicAdd(&SL.start[sl_arb_bucket(B.space[A.ba[i]], col)], 1u));
	}
	for(int j = tid; j < J.n; j += nth){
		int col = J.colour[j];
		if(col < 0) continue;
		J.row[(int)atomicAdd(&SL.start[sl_joint_bucket(SL, B.space[J.a[j]], col)], 1u) - jbase] = j;
	}
}

#ifndef CPB_EMU
// JOINTS = false: instantiation for spaces without constraints (the ten joint classes cost ~40 registers; without
// them more one-warp CTAs fit an SM).  `spaces`: the spaces this launch handles, one CTA each.
template<bool JOINTS>
__global__ void k_sl_solve(DBodies B, DArbs A, DJoints J, DRows R, DSpaceLocal SL, const int *__restrict__ spaces, int iterations, double dt, double dt_coef)
{
	extern __shared__ double4 s_vel[];
	__shared__ int s_a[CPB_MAX_COLOURS + 1], s_j[CPB_MAX_COLOURS + 1];   // bucket begins ([c]) / ends ([c + 1])
	__shared__ int s_cols[CPB_MAX_COLOURS], s_ncols;
	const int s = spaces[blockIdx.x], t = threadIdx.x, nt = blockDim.x;
	const int b0 = SL.body0[s], nb = SL.nbody[s];
	const int jbase = (int)SL.start[SL.n_spaces*CPB_MAX_COLOURS];
	for(int c = t; c <= CPB_MAX_COLOURS; c += nt){
		int ka = sl_arb_bucket(s, c) - 1, kj = sl_joint_bucket(SL, s, c) - 1;
		s_a[c] = (ka >= 0 ? (int)SL.start[ka] : 0);
		s_j[c] = (JOINTS ? (int)SL.start[kj] - jbase : 0);
	}
	VelShared vs = {s_vel, s_vel + nb, b0};
	for(int b = t; b < nb; b += nt){ s_vel[b] = B.V[b0 + b]; s_vel[nb + b] = B.VB[b0 + b]; }
	__syncthreads();
	if(t == 0){
		int n = 0;
		for(int c = 0; c < CPB_MAX_COLOURS; c++) if(s_a[c + 1] > s_a[c] || s_j[c + 1] > s_j[c]) s_cols[n++] = c;
		s_ncols = n;
	}
	__syncthreads();
	const int ncols = s_ncols;
	// indices of this thread's first row / joint of the next phase are fetched one phase ahead
	int pr = -1, pba = 0, pbb = 0, pcnt = 0, pq = -1, pj = 0, pja = 0, pjb = 0;
	#define SL_PREFETCH(c_) do { \
		pr = s_a[c_] + t; if(pr < s_a[(c_) + 1]){ const int4 h_ = R.hdr[pr]; pba = h_.x; pbb = h_.y; pcnt = h_.z; } else pr = -1; \
		pq = s_j[c_] + (nt - 1 - t); if(JOINTS && pq < s_j[(c_) + 1]){ pj = J.row[pq]; pja = J.a[pj]; pjb = J.b[pj]; } else pq = -1; } while(0)
	if(ncols > 0) SL_PREFETCH(s_cols[0]);
	for(int pass = 0; pass <= iterations; pass++){
		const int mode = (pass == 0 ? 0 : 1);
		for(int ci = 0; ci < ncols; ci++){
			const int c = s_cols[ci];
			const int cn = s_cols[ci + 1 < ncols ? ci + 1 : 0];
			if(c == CPB_OVERFLOW_COLOUR){
				if(t == 0){
					for(int r = s_a[c]; r < s_a[c + 1]; r++){ const int4 h_ = R.hdr[r]; solve_row_packed(vs, R, r, h_.x, h_.y, h_.z, mode, dt_coef); }
					if(JOINTS) for(int q = s_j[c]; q < s_j[c + 1]; q++){ int j = J.row[q]; solve_joint_idx(vs, B, J, j, J.a[j], J.b[j], mode, dt, dt_coef); }
				}
				SL_PREFETCH(cn);
			} else {
				int r = pr, ba = pba, bb = pbb, cnt = pcnt, q = pq, j = pj, ja = pja, jb = pjb;
				SL_PREFETCH(cn);   // constant data: issue before this phase's work so the latency overlaps it
				const int r1 = s_a[c + 1], q1 = s_j[c + 1];
				if(r >= 0){
					solve_row_packed(vs, R, r, ba, bb, cnt, mode, dt_coef);
					for(r += nt; r < r1; r += nt){ const int4 h_ = R.hdr[r]; solve_row_packed(vs, R, r, h_.x, h_.y, h_.z, mode, dt_coef); }
				}
				// joints from the far end of the CTA: in small colours a thread gets a row or a joint, not both
				if(JOINTS && q >= 0){
					solve_joint_idx(vs, B, J, j, ja, jb, mode, dt, dt_coef);
					for(q += nt; q < q1; q += nt){ int j2 = J.row[q]; solve_joint_idx(vs, B, J, j2, J.a[j2], J.b[j2], mode, dt, dt_coef); }
				}
			}
			__syncthreads();
		}
	}
	#undef SL_PREFETCH
	for(int b = t; b < nb; b += nt){ B.V[b0 + b] = s_vel[b]; B.VB[b0 + b] = s_vel[nb + b]; }
	// accumulated impulses back into the arbiter records
	for(int r = s_a[0] + t; r < s_a[CPB_MAX_COLOURS]; r += nt){
		const int4 h = R.hdr[r];
		const int i = h.w, cnt = (h.z < 0 ? -h.z : h.z);
		const double4 i0 = R.imp[r];
		double4 i1 = i0;
		if(cnt == 2) i1 = R.imp[R.cap + r];
		if(cnt >= 1){ const int sidx = CIDX(A, i, 0); A.jn[sidx] = i0.x; A.jt[sidx] = i0.y; A.jb[sidx] = i0.z; }
		if(cnt == 2){ const int sidx = CIDX(A, i, 1); A.jn[sidx] = i1.x; A.jt[sidx] = i1.y; A.jb[sidx] = i1.z; }
	}
}
#endif

#ifndef CPB_EMU
// Grid-wide barrier for the persistent kernel (launched cooperatively, so all CTAs are resident).
// bar[0] = arrival count, bar[1] = generation.  Thread 0 of every CTA arrives with one atomic and
// spins on the generation word; the gpu-scope fences on both sides order the phase's global
// writes before every later read (the same pattern cooperative_groups::grid_group::sync uses),
// at one L2 round trip instead of a library call per colour.
__device__ __forceinline__ unsigned bar_ld_acquire(const unsigned *p){ unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned bar_arrive(unsigned *p){ unsigned o; asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(o) : "l"(p) : "memory"); return o; }
// One monotonically increasing arrival counter per kernel (zeroed by k_reset_step): barrier number g is open once the
// count reaches g x nblocks.  The arrival is a release (everything this CTA wrote in the phase, ordered before it by the
// block barrier), the spin an acquire load: no stand-alone fences, no generation word, no reset by the last arriver.
// Measured on B200, 296 CTAs: 1.16 us per barrier against 2.35 us for fence + atomicAdd + generation word + fence
// (tools/micro/barrier_bench.cu); a two-level variant (groups of 16 CTAs) was slower than either.
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned nblocks, unsigned &target)
{
	__syncthreads();
	if(nblocks == 1) return;   // a single CTA (small scenes): the block barrier already orders its global accesses
	if(threadIdx.x == 0){
		target += nblocks;
		bar_arrive(bar);
		while((int)(bar_ld_acquire(bar) - target) < 0){ }
	}
	__syncthreads();
}
#define GRID_SYNC() grid_barrier(bar + (PHASE == 2 ? 32 : 0), gridDim.x, bar_target)
__device__ __forceinline__ unsigned long long global_ns(){ unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PROF(i) do { if(tid == 0) K.prof[i] = global_ns(); } while(0)

// K10 + K11 in one persistent launch.
#ifndef CPB_SOLVE_MIN_BLOCKS
#define CPB_SOLVE_MIN_BLOCKS 2
#endif

// SPACE_LOCAL: colour only, then histogram the (space, colour) buckets for the space-local solver below.
// JOINTS = false: instantiation for worlds without constraints (the ten joint classes cost ~20 registers of the
// iteration loop's budget; a pile of circles never needs them).
// PHASE 1 = colouring + row build, PHASE 2 = warm start + iterations + write-back: two launches, because the register
// allocation of a kernel is the maximum over its phases -- the row build holds a whole arbiter record in registers
// (gather, then scatter) and would cap the occupancy of the iteration loop, which is the part that needs warps to
// hide its gathers.  (PHASE 0 = both in one launch, kept for comparison.)
template<bool SPACE_LOCAL, bool STREAM_ROWS, bool JOINTS, int PHASE, int MINB> __global__ void __launch_bounds__(256, MINB) k_colour_solve(DBodies B, DArbs A, DJoints J, DRows R, DColour K, DCounters *C, unsigned *bar, DSpaceLocal SL, int use_hints, int iterations, double dt, double dt_coef, DPrestep P)
{
	__shared__ int s_hist[2*CPB_MAX_COLOURS];
	__shared__ int s_base[CPB_MAX_COLOURS];
	const int tid = CPB_TID, nth = CPB_NTHREADS;
	int nA = *A.count_ptr; if(nA > A.cap) nA = A.cap;
	unsigned bar_target = 0;
	if(PHASE != 2){
	if(threadIdx.x < 2*CPB_MAX_COLOURS) s_hist[threadIdx.x] = 0;
	__syncthreads();
	// flush this CTA's colour histogram (arbiters | joints) into the global one
	#define HIST_FLUSH() do { __syncthreads(); if(threadIdx.x < 2*CPB_MAX_COLOURS){ int v_ = s_hist[threadIdx.x]; \
		if(v_){ atomicAdd(threadIdx.x < CPB_MAX_COLOURS ? &K.ccount[threadIdx.x] : &K.jcount[threadIdx.x - CPB_MAX_COLOURS], v_); s_hist[threadIdx.x] = 0; } } } while(0)

	PROF(0);
	// K10: colouring -- keep last step's colours, then Jones-Plassmann rounds over the worklist
	colour_seed(B, A, J, K, s_hist, nA, use_hints & 1, tid, nth);
	HIST_FLUSH();
	GRID_SYNC();
	int rounds_done = 0;
	for(int round = 0; round < CPB_MAX_COLOUR_ROUNDS; round++){
		if(*((volatile int *)&K.wl_n[round]) == 0) break;
		rounds_done = round + 1;
		colour_phase_a(B, A, J, K, nA, round, tid, nth);
		GRID_SYNC();
		colour_phase_b(B, A, J, K, s_hist, nA, round, tid, nth);
		HIST_FLUSH();
		GRID_SYNC();
	}
	if(rounds_done == CPB_MAX_COLOUR_ROUNDS && *((volatile int *)&K.wl_n[CPB_MAX_COLOUR_ROUNDS]) != 0){
		colour_leftover(A, J, K, s_hist, nA, CPB_MAX_COLOUR_ROUNDS, tid, nth);
		HIST_FLUSH();
		GRID_SYNC();
	}
	PROF(1);
	if(tid == 0){ colour_starts(K, C); K.prof[5] = (unsigned long long)rounds_done; K.prof[6] = (unsigned long long)K.wl_n[0]; }
	if(SPACE_LOCAL){
		// space-local mode: this launch only colours; count the (space, colour) buckets for k_sl_rows
		sl_count(B, A, J, SL, nA, tid, nth);
		if(tid == 0){ K.prof[2] = K.prof[3] = K.prof[4] = global_ns(); }
		return;
	}
	GRID_SYNC();
	build_rows(B, P, A, J, R, K, s_hist, s_base, nA, tid, nth, (use_hints & 4) == 0);   // bit 2 of use_hints: experiment switch, strided records
	if(PHASE == 1){ __syncthreads(); PROF(2); return; }
	GRID_SYNC();
	PROF(2);
	}

	// K11: warm start then iterations, colour by colour.  A thread's rows of a colour are r0 + tid + k*nth and
	// its joints q0 + (nth-1-tid) + k*nth (from the other end of the grid, so that in small colours a thread
	// has a row or a joint, not both).  The indices of the first row/joint of the NEXT phase are fetched
	// before the barrier: once it opens, the velocity gathers can issue at once.
	__shared__ int s_cstart[CPB_MAX_COLOURS + 1], s_jstart[CPB_MAX_COLOURS + 1];
	if(threadIdx.x <= CPB_MAX_COLOURS){ s_cstart[threadIdx.x] = K.cstart[threadIdx.x]; s_jstart[threadIdx.x] = K.jstart[threadIdx.x]; }
	__syncthreads();
	int ncol = *((volatile int *)&C->n_colours);
	int nreg = (ncol > CPB_OVERFLOW_COLOUR ? CPB_OVERFLOW_COLOUR : ncol);
	bool has_overflow = (ncol > CPB_OVERFLOW_COLOUR);
	const int jtid = nth - 1 - tid;
	const VelGlobalT<STREAM_ROWS> vg = {B.V, B.VB};
	const bool PHASE_PREFETCH = (STREAM_ROWS && (use_hints & 2) == 0);   // bit 1 of use_hints: experiment switch, prefetch off
	int n_solves = 0, n_idle = 0;                // step statistics: row visits of the iteration passes / visits that changed nothing
	int pr = -1, pba = 0, pbb = 0, pcnt = 0;     // prefetched row
	int pq = -1, pj = 0, pja = 0, pjb = 0;       // prefetched joint
	#define PREFETCH_PHASE(c_) do { \
		pr = s_cstart[c_] + tid; if(pr < s_cstart[(c_) + 1]){ pba = row_ld<STREAM_ROWS>(&R.ba[pr]); pbb = row_ld<STREAM_ROWS>(&R.bb[pr]); pcnt = row_ld<STREAM_ROWS>(&R.cnt[pr]); if(PHASE_PREFETCH) row_prefetch_phase(R, pr); } else pr = -1; \
		pq = s_jstart[c_] + jtid; if(JOINTS && pq < s_jstart[(c_) + 1]){ pj = J.row[pq]; pja = J.a[pj]; pjb = J.b[pj]; } else pq = -1; } while(0)
	if(nreg > 0) PREFETCH_PHASE(0);
	for(int pass = 0; pass <= iterations; pass++){
		int mode = (pass == 0 ? 0 : 1);
		for(int c = 0; c < nreg; c++){
			const int r1 = s_cstart[c + 1], j1 = s_jstart[c + 1];
			while(pr >= 0){
				int r = pr, ba = pba, bb = pbb, cnt = pcnt;
				pr += nth;
				if(pr < r1){ pba = row_ld<STREAM_ROWS>(&R.ba[pr]); pbb = row_ld<STREAM_ROWS>(&R.bb[pr]); pcnt = row_ld<STREAM_ROWS>(&R.cnt[pr]); row_prefetch(R, pr); } else pr = -1;
				const bool moved = solve_row_idx(vg, B, R, r, ba, bb, cnt, mode, dt_coef);
				if(mode){ n_solves++; n_idle += (moved ? 0 : 1); }
			}
			while(JOINTS && pq >= 0){
				int j = pj, a = pja, b = pjb;
				pq += nth;
				if(pq < j1){ pj = J.row[pq]; pja = J.a[pj]; pjb = J.b[pj]; } else pq = -1;
				solve_joint_idx(vg, B, J, j, a, b, mode, dt, dt_coef);
			}
			int cn = (c + 1 < nreg ? c + 1 : 0);
			if(c + 1 < nreg || pass < iterations) PREFETCH_PHASE(cn);
			GRID_SYNC();
		}
		if(has_overflow){
			if(tid == 0) solve_overflow<JOINTS>(B, R, J, K, mode, dt, dt_coef);
			GRID_SYNC();
		}
		if(pass == 0) PROF(3);
	}
	#undef PREFETCH_PHASE
	{
		n_solves = __reduce_add_sync(0xffffffffu, n_solves); n_idle = __reduce_add_sync(0xffffffffu, n_idle);
		if((threadIdx.x & 31) == 0 && n_solves){ atomicAdd(&C->n_row_solves, n_solves); atomicAdd(&C->n_row_idle, n_idle); }
	}
	PROF(4);
	int n_rows = K.cstart[CPB_MAX_COLOURS];
	rows_writeback(A, R, n_rows, tid, nth);
}
#endif

// kernels used by the emulation build (multi-launch variant of the same phases)
__global__ void k_colour_seed(DBodies B, DArbs A, DJoints J, DColour K, int use_hints){
	int nA = *A.count_ptr; if(nA > A.cap) nA = A.cap;
	colour_seed(B, A, J, K, (int *)NULL, nA, use_hints & 1, CPB_TID, CPB_NTHREADS);
}
__global__ void k_colour_a(DBodies B, DArbs A, DJoints J, DColour K, int round){
	int nA = *A.count_ptr; if(nA > A.cap) nA = A.cap;
	colour_phase_a(B, A, J, K, nA, round, CPB_TID, CPB_NTHREADS);
}
__global__ void k_colour_b(DBodies B, DArbs A, DJoints J, DColour K, int round){
	int nA = *A.count_ptr; if(nA > A.cap) nA = A.cap;
	colour_phase_b(B, A, J, K, (int *)NULL, nA, round, CPB_TID, CPB_NTHREADS);
}
__global__ void k_colour_finish(DArbs A, DJoints J, DRows R, DColour K, DCounters *C, int stage){
	int nA = *A.count_ptr; if(nA > A.cap) nA = A.cap;
	if(stage == 2){ colour_leftover(A, J, K, (int *)NULL, nA, CPB_MAX_COLOUR_ROUNDS, CPB_TID, CPB_NTHREADS); return; }
	if(stage == 0){ if(CPB_TID == 0) colour_starts(K, C); }
	else { DPrestep P = {NULL, 0.0, 0}; DBodies B = {}; build_rows(B, P, A, J, R, K, (int *)NULL, (int *)NULL, nA, CPB_TID, CPB_NTHREADS, true); }
}
__global__ void k_solve_colour(DBodies B, DRows R, DJoints J, DColour K, int colour, int mode, double dt, double dt_coef){
	if(colour == CPB_OVERFLOW_COLOUR){ if(CPB_TID == 0) solve_overflow<true>(B, R, J, K, mode, dt, dt_coef); }
	else solve_colour(B, R, J, K, colour, mode, dt, dt_coef, CPB_TID, CPB_NTHREADS);
}
__global__ void k_rows_writeback(DArbs A, DRows R, DColour K){
	rows_writeback(A, R, K.cstart[CPB_MAX_COLOURS], CPB_TID, CPB_NTHREADS);
}

// ---- serial validation mode: the reference's exact order (cpSpaceStep.c:406-427) ----
// order[n_order] lists arbiter record indices; joints follow in upload order.
__global__ void k_solve_serial(DBodies B, DArbs A, DJoints J, const int *__restrict__ order, int n_order, const int *__restrict__ jorder, int n_jorder, int iterations, double dt, double dt_coef)
{
	if(CPB_TID != 0) return;
	// joints named by the caller first (J.row doubles as a "listed" marker here), the rest in upload order
	for(int j = 0; j < J.n; j++) J.row[j] = 0;
	for(int q = 0; q < n_jorder; q++){ int j = jorder[q]; if(j >= 0 && j < J.n) J.row[j] = 1; }
	for(int pass = 0; pass <= iterations; pass++){
		for(int q = 0; q < n_order; q++){
			int i = order[q];
			if(i < 0) continue;
			bool reversed = (i & 0x40000000) != 0;
			i &= 0x3fffffff;
			if(A.active[i] != 1) continue;
			if(pass == 0 && A.state[i] == CPB200_ARB_FIRST_COLLISION) continue;
			int ba = A.ba[i], bb = A.bb[i];
			V2 mia = B.MI[ba], mib = B.MI[bb];
			double4 Va = B.V[ba], Vb = B.V[bb], VBa = B.VB[ba], VBb = B.VB[bb];
			V2 n = A.n[i];
			int cnt = A.cnt[i];
			RowDirty dirty = {false, false, false, false};
			for(int kk = 0; kk < cnt; kk++){
				int k = (reversed ? cnt - 1 - kk : kk);
				int s = CIDX(A, i, k);
				if(pass == 0) contact_apply_cached(Va, Vb, mia, mib, n, A.r1[s], A.r2[s], A.jn[s], A.jt[s], dt_coef);
				else contact_apply(Va, Vb, VBa, VBb, mia, mib, n, A.svr[i], A.u[i], A.r1[s], A.r2[s], A.nmass[s], A.tmass[s], A.bias[s], A.bounce[s], A.jn[s], A.jt[s], A.jb[s], dirty);
			}
			if(mia.x != 0.0 || mia.y != 0.0){ B.V[ba] = Va; B.VB[ba] = VBa; }
			if(mib.x != 0.0 || mib.y != 0.0){ B.V[bb] = Vb; B.VB[bb] = VBb; }
		}
		for(int q = 0; q < n_jorder; q++){
			int j = jorder[q];
			if(j < 0 || j >= J.n || J.colour[j] == -2) continue;
			solve_joint(B, J, j, (pass == 0 ? 0 : 1), dt, dt_coef);
		}
		for(int j = 0; j < J.n; j++){
			if(J.colour[j] == -2 || J.row[j]) continue;
			solve_joint(B, J, j, (pass == 0 ? 0 : 1), dt, dt_coef);
		}
	}
}
