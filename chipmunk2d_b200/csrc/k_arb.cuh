// k_arb.cuh -- K5 wrapper + K6 (arbiter cache) + K8 (contact prestep).
//
// The reference handles one candidate pair at a time inside the broadphase callback
// cpSpaceCollideShapes (cpSpaceStep.c:235-289): narrowphase, find-or-create the arbiter in
// the cpHashSet keyed by the shape pair, cpArbiterUpdate (warm-start matching by contact
// hash, cpArbiter.c:356-414), push to space->arbiters.  Here one thread does the same for one
// pair against a device hash table:
//   * prev table/records: last step's arbiters (double buffer = the reference's contact
//     buffer ring, cpSpaceStep.c:109-183);
//   * cur table/records: filled this step; unmatched prev records are aged or carried over by
//     k_arb_carry (= cpSpaceArbiterSetFilter, cpSpaceStep.c:292-325).
#pragma once
#include "cpb_world.h"
#include "k_narrow.cuh"
#include "prims.cuh"

CPB_DEVICE NShape load_nshape(const DShapes &S, const DBodies &B, int s){
	NShape o;
	o.type = S.type[s];
	o.hashid = S.hashid[s];
	o.r = S.r[s];
	double4 bb = S.bb[s];
	o.bbc = vlerp(v2(bb.x, bb.y), v2(bb.z, bb.w), 0.5);
	o.count = 0; o.pv = 0; o.pn = 0; o.sv = 0;
	o.a = v2(0, 0); o.b = v2(0, 0); o.n = v2(0, 0);
	o.rot = v2(1, 0); o.atan = v2(0, 0); o.btan = v2(0, 0);
	if(o.type == CPB200_SHAPE_CIRCLE){
		o.a = S.wa[s];
	} else if(o.type == CPB200_SHAPE_SEGMENT){
		o.a = S.wa[s]; o.b = S.wb[s]; o.n = S.wn[s];
		o.rot = B.rot[S.body[s]];
		o.atan = S.atan[s]; o.btan = S.btan[s];
	} else {
		o.count = S.pcount[s];
		o.pv = S.wpv + S.poff[s];
		o.pn = S.wpn + S.poff[s];
	}
	return o;
}

CPB_DEVICE uint64_t arb_key(uint32_t ha, uint32_t hb){
	uint64_t lo = (ha < hb ? ha : hb), hi = (ha < hb ? hb : ha);
	return (lo << 32) | hi;
}

#ifndef CPB_EMU
__device__ __forceinline__ ulonglong2 ld_slot(const ulonglong2 *p){   // one 128-bit load per probe (the tables are read-only while probed)
	ulonglong2 v;
	asm volatile("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
	return v;
}
#else
static inline ulonglong2 ld_slot(const ulonglong2 *p){ return *p; }
#endif

CPB_DEVICE int table_find(const DTable &T, uint32_t mask, uint64_t key){   // mask = *T.dmask, read once per kernel
	uint32_t slot = (uint32_t)mix64(key) & mask;
	for(uint32_t probe = 0; probe <= mask; probe++){
		ulonglong2 e = ld_slot(&T.slots[slot]);
		if(e.x == key) return (int)e.y;
		if(e.x == 0) return -1;
		slot = (slot + 1) & mask;
	}
	return -1;
}

CPB_DEVICE bool table_insert(const DTable &T, uint32_t mask, uint64_t key, int val){
	uint32_t slot = (uint32_t)mix64(key) & mask;
	for(uint32_t probe = 0; probe <= mask; probe++){
		unsigned long long old = atomicCAS(&T.slots[slot].x, 0ull, (unsigned long long)key);
		if(old == 0ull || old == (unsigned long long)key){ T.slots[slot].y = (unsigned long long)val; return true; }
		slot = (slot + 1) & mask;
	}
	return false;
}

// The table of this step's records is built in one pass AFTER the collision phase and the cache filter have
// appended them (k_collide, k_arb_carry): inside those kernels the insert's compare-and-swap sat at the end of a
// chain of dependent memory accesses per pair (a quarter of k_collide's stall samples); here every record is an
// independent thread.  The table is sized to the records of the step: clear (which picks the size) -> build.
__global__ void k_table_clear(DArbs cur, DTable T)
{
	int n = *cur.count_ptr; if(n > cur.cap) n = cur.cap;
	uint32_t want = 64;
	while(want < 2u*(uint32_t)n && want - 1u < T.mask) want <<= 1;
	const uint32_t mask = want - 1u;
	if(CPB_TID == 0) *T.dmask = mask;       // read by k_table_build and by the next step's lookups
	for(size_t i = CPB_TID; i <= (size_t)mask; i += CPB_NTHREADS) T.slots[i] = make_ulonglong2(0ull, 0ull);
}

__global__ void k_table_build(DArbs cur, DTable T, DCounters *C)
{
	int n = *cur.count_ptr; if(n > cur.cap) n = cur.cap;
	const uint32_t mask = *T.dmask;
	for(int i = CPB_TID; i < n; i += CPB_NTHREADS){
		uint64_t key = cur.key[i];
		if(key == ~0ull || key == 0ull) continue;
		if(!table_insert(T, mask, key, i)) atomicOr((unsigned *)&C->overflow, 4u);
	}
}

// small worlds (launch-bound): both passes in one CTA
#ifndef CPB_EMU
__global__ void __launch_bounds__(1024) k_table_small(DArbs cur, DTable T, DCounters *C)
{
	int n = *cur.count_ptr; if(n > cur.cap) n = cur.cap;
	uint32_t want = 64;
	while(want < 2u*(uint32_t)n && want - 1u < T.mask) want <<= 1;
	const uint32_t mask = want - 1u;
	if(CPB_TID == 0) *T.dmask = mask;
	for(size_t i = CPB_TID; i <= (size_t)mask; i += CPB_NTHREADS) T.slots[i] = make_ulonglong2(0ull, 0ull);
	__syncthreads();
	for(int i = CPB_TID; i < n; i += CPB_NTHREADS){
		uint64_t key = cur.key[i];
		if(key == ~0ull || key == 0ull) continue;
		if(!table_insert(T, mask, key, i)) atomicOr((unsigned *)&C->overflow, 4u);
	}
}
#endif

// K5 + K6 for one pair class (CLS 0 circle-circle, 1 circle-segment, 2 GJK family).
// P lists shape pairs with a.type <= b.type.  Block-stride loop over a device-side count.
#ifndef CPB_COLLIDE_CTAS
#define CPB_COLLIDE_CTAS 8
#endif
// The GJK/EPA family must leave room for CPB_COLLIDE_GJK_CTAS resident CTAs per SM (4 -> at most 128 registers; left
// alone the compiler takes 233 and a third of the warps: +12 % on the polygon scenes).  The circle kernels are
// faster with the registers they ask for (capped to 80 they spill: +8 % on the 1 M pile).
#ifndef CPB_COLLIDE_GJK_CTAS
#define CPB_COLLIDE_GJK_CTAS 4
#endif
template <int CLS>
__global__ void __launch_bounds__(128, (CLS == 2 ? CPB_COLLIDE_GJK_CTAS : 1)) k_collide(DShapes S, DBodies B, const int *__restrict__ pa, const int *__restrict__ pb, const int *__restrict__ pcount, int pcap,
	DArbs prev, DTable prev_table, DArbs cur, DCounters *C)
{
	int np = *pcount; if(np > pcap) np = pcap;
	const uint32_t stamp = C->stamp;
#ifndef CPB_EMU
	// per-thread staging of both polygons' vertices (GJK class only): 2 shapes x 8 vertices x (x, y) x 128 threads = 32 KB
	__shared__ double s_verts[CLS == 2 ? 2*2*CPB_GJK_STAGE_VERTS*CPB_GJK_STAGE_STRIDE : 1];
	const bool stage_verts = (CLS == 2 && C->no_gjk_stage == 0);     // experiment switch CPB200_NO_GJK_STAGE
#endif
	const uint32_t pmask = *prev_table.dmask;
	int my_active = 0, my_contacts = 0;     // step statistics: summed per thread, flushed once per warp after the loop
	for(int base = blockIdx.x*blockDim.x; base < np; base += gridDim.x*blockDim.x){
		int i = base + threadIdx.x;
		bool have = false;
		Manifold m; m.count = 0; m.id = 0; m.n = v2(0, 0);
		int sa = 0, sb = 0, pi = -1, ba = 0, bb = 0;
		bool sensor = false;
		uint2 ida = {0u, 0u}, idb = {0u, 0u};
		uint64_t key = 0;
		double4 w0 = make_double4(0, 0, 0, 0), w1 = w0;
		// gathers the record needs (body positions, materials, body types) are issued as soon as the body indices
		// are known, together with the table probe: one memory latency for all of them instead of a chain
		V2 pa_ = v2(0, 0), pb_ = v2(0, 0);
		double4 ma = w0, mb = w0;
		int type_a = 0, type_b = 0;
		if(i < np){
			sa = pa[i]; sb = pb[i];
			if(CLS == 0){
				// circle-circle: everything the test needs sits in sector 0 of one packed line per shape, everything a
				// hit needs beyond that in sector 1 of the same lines (cpb_world.h)
				double4 ca = ld4_nc(&S.circ[2*(size_t)sa]), cb = ld4_nc(&S.circ[2*(size_t)sb]);
				NShape a, b;
				a.type = 0; a.a = v2(ca.x, ca.y); a.r = ca.z; b.type = 0; b.a = v2(cb.x, cb.y); b.r = cb.z;
				const unsigned long long wa_ = (unsigned long long)__double_as_longlong(ca.w), wb_ = (unsigned long long)__double_as_longlong(cb.w);
				ba = (int)(wa_ & 0x1fffffffull); bb = (int)(wb_ & 0x1fffffffull); sensor = (((wa_ | wb_) >> 31) & 1ull) != 0;
				ida.x = (uint32_t)(wa_ >> 32); idb.x = (uint32_t)(wb_ >> 32);
				key = arb_key(ida.x, idb.x);
				circle_to_circle(a, b, m);
				if(m.count > 0){
					double4 da = ld4_nc(&S.circ[2*(size_t)sa + 1]), db = ld4_nc(&S.circ[2*(size_t)sb + 1]);
					pi = table_find(prev_table, pmask, key);
					pa_ = v2(da.x, da.y); pb_ = v2(db.x, db.y);
					ma = make_double4(da.z, da.w, 0.0, 0.0); mb = make_double4(db.z, db.w, 0.0, 0.0);
					type_a = ((wa_ >> 30) & 1ull) ? CPB200_BODY_STATIC : CPB200_BODY_DYNAMIC;   // only "dynamic or not" matters below
					type_b = ((wb_ >> 30) & 1ull) ? CPB200_BODY_STATIC : CPB200_BODY_DYNAMIC;
					if(((wa_ | wb_) >> 29) & 1ull){ ma = ld4_nc(&S.mat[sa]); mb = ld4_nc(&S.mat[sb]); }    // surface velocities are rare
				}
				if(pi >= 0){ w0 = ld4_nc(&prev.warm[2*pi]); w1 = ld4_nc(&prev.warm[2*pi + 1]); }
			} else {
				ida = S.ids[sa]; idb = S.ids[sb];
				key = arb_key(ida.x, idb.x);
				ba = S.body[sa]; bb = S.body[sb]; sensor = (S.sensor[sa] || S.sensor[sb]);
				pi = table_find(prev_table, pmask, key);
				if(pi >= 0){ w0 = ld4_nc(&prev.warm[2*pi]); w1 = ld4_nc(&prev.warm[2*pi + 1]); }
				m.id = (pi >= 0 ? (uint32_t)((unsigned long long)__double_as_longlong(w1.z) >> 32) : 0u);
				NShape a = load_nshape(S, B, sa), b = load_nshape(S, B, sb);
#ifndef CPB_EMU
				if(CLS == 2 && stage_verts){
					// north star stage 3: the polygons' world vertices go to shared memory once per pair (nshape_vert)
					if(a.type == CPB200_SHAPE_POLY && a.count <= CPB_GJK_STAGE_VERTS){
						double *sv = s_verts + threadIdx.x;
						for(int k = 0; k < a.count; k++){ V2 v = a.pv[k]; sv[(2*k)*CPB_GJK_STAGE_STRIDE] = v.x; sv[(2*k + 1)*CPB_GJK_STAGE_STRIDE] = v.y; }
						a.sv = sv;
					}
					if(b.type == CPB200_SHAPE_POLY && b.count <= CPB_GJK_STAGE_VERTS){
						double *sv = s_verts + 2*CPB_GJK_STAGE_VERTS*CPB_GJK_STAGE_STRIDE + threadIdx.x;
						for(int k = 0; k < b.count; k++){ V2 v = b.pv[k]; sv[(2*k)*CPB_GJK_STAGE_STRIDE] = v.x; sv[(2*k + 1)*CPB_GJK_STAGE_STRIDE] = v.y; }
						b.sv = sv;
					}
				}
#endif
				if(CLS == 1) circle_to_segment(a, b, m);
				else collide_shapes(a, b, m);
				// (not before the narrowphase: GJK/EPA needs the registers, and many box pairs do not touch)
				if(m.count > 0){
					pa_ = B.pos[ba]; pb_ = B.pos[bb]; type_a = B.type[ba]; type_b = B.type[bb];
					ma = ld4_nc(&S.mat[sa]); mb = ld4_nc(&S.mat[sb]);
				}
			}
			have = (m.count > 0);
		}
		int slot = cpb_warp_append(cur.count_ptr, have);
		if(!have) continue;
		if(slot >= cur.cap){ atomicOr((unsigned *)&C->overflow, 2u); continue; }

		// cpArbiterUpdate (cpArbiter.c:356-414); the previous record's fields come from its packed line
		int state = CPB200_ARB_FIRST_COLLISION;
		int pcnt = 0, pactive = 0, pcolour = -1;
		uint64_t phash[2] = {0, 0};
		double pjn[2] = {0.0, 0.0}, pjt[2] = {0.0, 0.0};
		if(pi >= 0){
			unsigned meta = (unsigned)(unsigned long long)__double_as_longlong(w1.z);
			int ps = (int)(meta & 0xffu);
			pcnt = (int)((meta >> 8) & 0xffu); pactive = (int)((meta >> 16) & 0xffu); pcolour = (int)(signed char)(meta >> 24);
			phash[0] = (uint64_t)__double_as_longlong(w1.x); phash[1] = (uint64_t)__double_as_longlong(w1.y);
			pjn[0] = w0.x; pjt[0] = w0.y; pjn[1] = w0.z; pjt[1] = w0.w;
			// CACHED -> FIRST_COLLISION (cpArbiter.c:412-413); IGNORE is sticky until separation
			state = (ps == CPB200_ARB_CACHED ? CPB200_ARB_FIRST_COLLISION : (ps == CPB200_ARB_IGNORE ? CPB200_ARB_IGNORE : CPB200_ARB_NORMAL));
		}
		// active <=> pushed to space->arbiters (cpSpaceStep.c:261-274); the default handler accepts everything
		const bool both_inf = (type_a != CPB200_BODY_DYNAMIC) && (type_b != CPB200_BODY_DYNAMIC);
		const bool active = (state != CPB200_ARB_IGNORE) && !sensor && !both_inf;
		// a rejected first contact still shows FIRST_COLLISION to the begin handler; k_arb_prestep downgrades it
		// to NORMAL afterwards (cpSpaceStep.c:283)
		if(!active && state != CPB200_ARB_IGNORE && state != CPB200_ARB_FIRST_COLLISION) state = CPB200_ARB_NORMAL;
		const V2 svr0 = vsub(v2(mb.z, mb.w), v2(ma.z, ma.w));
		if(pi >= 0) prev.seen[pi] = 1;
		for(int k = 0; k < m.count; k++){
			double jn = 0.0, jt = 0.0;
			for(int j = 0; j < pcnt; j++){
				if(m.hash[k] == phash[j]){ jn = pjn[j]; jt = pjt[j]; }
			}
			int c = CIDX(cur, slot, k);
			cur.r1[c] = vsub(m.p1[k], pa_);
			cur.r2[c] = vsub(m.p2[k], pb_);
			cur.jn[c] = jn; cur.jt[c] = jt; cur.jb[c] = 0.0;
			cur.hash[c] = m.hash[k];
			cur.nmass[c] = 0.0; cur.tmass[c] = 0.0; cur.bounce[c] = 0.0; cur.bias[c] = 0.0;
		}
		cur.key[slot] = key;
		cur.sa[slot] = sa; cur.sb[slot] = sb; cur.ba[slot] = ba; cur.bb[slot] = bb;
		cur.cnt[slot] = m.count;
		cur.n[slot] = m.n;
		cur.gjkid[slot] = m.id;
		cur.e[slot] = ma.x*mb.x;
		cur.u[slot] = ma.y*mb.y;
		cur.svr[slot] = vsub(svr0, vmul(m.n, vdot(svr0, m.n)));
		cur.stamp[slot] = stamp;
		cur.seen[slot] = 0;
		cur.colour[slot] = -1;
		// colouring priority: hash of the space-local shape pair; circle pairs never loaded the local ids, the
		// colouring resolves the sentinel for the few records it has to colour afresh (arb_pri, k_solve.cuh)
		cur.pri[slot] = (CLS == 0 ? ~0ull : mix64(arb_key(ida.y, idb.y)) >> 8);
		cur.hint[slot] = (pi >= 0 && pactive == 1 ? pcolour : -1);
		cur.active[slot] = active ? 1 : 0;
		cur.state[slot] = state;
		if(active){ my_active += 1; my_contacts += m.count; }
	}
#ifndef CPB_EMU
	{
		// (inside the loop these were two atomics per warp and trip on two neighbouring hot words: a quarter of the
		// kernel's time on the 1 M pile); every lane comes through here exactly once, after its last trip
		unsigned lanes = __activemask();
		int na = __reduce_add_sync(lanes, my_active), nc = __reduce_add_sync(lanes, my_contacts);
		if((threadIdx.x & 31) == (unsigned)(__ffs(lanes) - 1) && na){ atomicAdd(&C->n_active, na); atomicAdd(&C->n_contacts, nc); }
	}
#else
	if(my_active){ atomicAdd(&C->n_active, my_active); atomicAdd(&C->n_contacts, my_contacts); }
#endif
}

// Packs what the collision phase reads from last step's records into one 64-byte line per record (the
// lookups are random: one DRAM burst instead of eight scattered sectors).
__global__ void k_pack_warm(DArbs A)
{
	int n = *A.count_ptr; if(n > A.cap) n = A.cap;
	for(int i = CPB_TID; i < n; i += CPB_NTHREADS){
		unsigned long long meta = ((unsigned long long)A.gjkid[i] << 32) | ((unsigned long long)(unsigned char)(signed char)A.colour[i] << 24) |
			((unsigned long long)(A.active[i] & 0xff) << 16) | ((unsigned long long)(A.cnt[i] & 0xff) << 8) | (unsigned long long)(A.state[i] & 0xff);
		const int c0 = CIDX(A, i, 0), c1 = CIDX(A, i, 1);
		A.warm[2*i] = make_double4(A.jn[c0], A.jt[c0], A.jn[c1], A.jt[c1]);
		A.warm[2*i + 1] = make_double4(__longlong_as_double((long long)A.hash[c0]), __longlong_as_double((long long)A.hash[c1]), __longlong_as_double((long long)meta), 0.0);
	}
}

// cpSpaceArbiterSetFilter (cpSpaceStep.c:292-325) over the previous step's records that the
// collision phase did not touch.
__global__ void k_arb_carry(DBodies B, DArbs prev, DArbs cur, const DSpace *__restrict__ spaces, DCounters *C)
{
	int n_prev = *prev.count_ptr; if(n_prev > prev.cap) n_prev = prev.cap;
	const uint32_t stamp = C->stamp;
	for(int base = blockIdx.x*blockDim.x; base < n_prev; base += gridDim.x*blockDim.x){
		int i = base + threadIdx.x;
		bool keep = false;
		int new_state = 0, new_active = 0;
		if(i < n_prev && !prev.seen[i] && prev.key[i] != ~0ull){ // ~0 = a shape of this record was removed
			int ba = prev.ba[i], bb = prev.bb[i];
			bool a_rest = (B.type[ba] == CPB200_BODY_STATIC) || B.sleeping[ba];
			bool b_rest = (B.type[bb] == CPB200_BODY_STATIC) || B.sleeping[bb];
			new_state = prev.state[i];
			new_active = (prev.active[i] == 2 ? 2 : 0);
			if(a_rest && b_rest){
				keep = true; // preserved untouched (cpSpaceStep.c:302-307)
			} else if(prev.active[i] == 2){
				// a body of a dormant arbiter woke up: back into the solver with its saved contacts
				// and a fresh stamp (cpSpaceActivateBody, cpSpaceComponent.c:45-73)
				keep = true; new_active = 1;
			} else {
				uint32_t ticks = stamp - prev.stamp[i];
				if(ticks >= 1 && new_state != CPB200_ARB_CACHED) new_state = CPB200_ARB_CACHED;
				keep = (ticks < spaces[B.space[ba]].persistence);
			}
		}
		int slot = cpb_warp_append(cur.count_ptr, keep);
		if(!keep) continue;
		if(slot >= cur.cap){ atomicOr((unsigned *)&C->overflow, 2u); continue; }
		// gather the whole record first, then scatter it: copied field by field the loads would queue up behind the
		// stores (the compiler must assume prev and cur alias), one memory latency per field
		const uint64_t key = prev.key[i];
		const int sa = prev.sa[i], sb = prev.sb[i], ba_ = prev.ba[i], bb_ = prev.bb[i], cnt = prev.cnt[i];
		const uint32_t pstamp = prev.stamp[i], gjkid = prev.gjkid[i];
		const V2 n = prev.n[i], svr = prev.svr[i];
		const double e = prev.e[i], u = prev.u[i];
		const uint64_t pri = prev.pri[i];
		V2 r1[2], r2[2]; double nm[2], tm[2], bo[2], bi[2], jn[2], jt[2], jb[2]; uint64_t hs[2];
#pragma unroll
		for(int k = 0; k < 2; k++){
			int p = CIDX(prev, i, k);
			r1[k] = prev.r1[p]; r2[k] = prev.r2[p];
			nm[k] = prev.nmass[p]; tm[k] = prev.tmass[p]; bo[k] = prev.bounce[p]; bi[k] = prev.bias[p];
			jn[k] = prev.jn[p]; jt[k] = prev.jt[p]; jb[k] = prev.jb[p];
			hs[k] = prev.hash[p];
		}
		cur.key[slot] = key;
		cur.sa[slot] = sa; cur.sb[slot] = sb; cur.ba[slot] = ba_; cur.bb[slot] = bb_;
		cur.cnt[slot] = cnt;
		cur.state[slot] = new_state;
		cur.stamp[slot] = (new_active == 1 ? stamp : pstamp);
		cur.active[slot] = new_active;
		cur.seen[slot] = 0;
		cur.gjkid[slot] = gjkid;
		cur.n[slot] = n; cur.e[slot] = e; cur.u[slot] = u; cur.svr[slot] = svr;
		cur.colour[slot] = -1;
		cur.pri[slot] = pri;
		cur.hint[slot] = -1;
#pragma unroll
		for(int k = 0; k < 2; k++){
			int c = CIDX(cur, slot, k);
			cur.r1[c] = r1[k]; cur.r2[c] = r2[k];
			cur.nmass[c] = nm[k]; cur.tmass[c] = tm[k]; cur.bounce[c] = bo[k]; cur.bias[c] = bi[k];
			cur.jn[c] = jn[k]; cur.jt[c] = jt[k]; cur.jb[c] = jb[k];
			cur.hash[c] = hs[k];
		}
		if(new_active == 1){ atomicAdd(&C->n_active, 1); atomicAdd(&C->n_contacts, cnt); }
		else atomicAdd(&C->n_cached, 1);
	}
}

// K8: cpArbiterPreStep (cpArbiter.c:416-439) with k_scalar (chipmunk_private.h:205-226).
CPB_DEVICE double k_scalar_body(V2 mi, V2 r, V2 n){
	double rcn = vcross(r, n);
	return mi.x + mi.y*rcn*rcn;
}

// cpArbiterPreStep for one contact (cpArbiter.c:416-439).  Shared by the stand-alone K8 kernel and by the solver's row
// build (which computes the same numbers on the fly in a production step): one body of code, one operation order.
struct PrestepBodies { V2 mia, mib; double4 Va, Vb; V2 body_delta; };
CPB_DEVICE PrestepBodies prestep_bodies(const DBodies &B, int ba, int bb){
	PrestepBodies pb;
	pb.mia = B.MI[ba]; pb.mib = B.MI[bb];
	pb.Va = B.V[ba]; pb.Vb = B.V[bb];
	pb.body_delta = vsub(B.pos[bb], B.pos[ba]);
	return pb;
}
CPB_DEVICE void prestep_contact(const PrestepBodies &pb, V2 n_, double e, double bias_coef, double slop, double dt, V2 r1, V2 r2,
	double &nmass, double &tmass, double &bias, double &bounce)
{
	nmass = 1.0/(k_scalar_body(pb.mia, r1, n_) + k_scalar_body(pb.mib, r2, n_));
	V2 t = vperp(n_);
	tmass = 1.0/(k_scalar_body(pb.mia, r1, t) + k_scalar_body(pb.mib, r2, t));
	double dist = vdot(vadd(vsub(r2, r1), pb.body_delta), n_);
	bias = -bias_coef*fmin_cp(0.0, dist + slop)/dt;
	// normal_relative_velocity (chipmunk_private.h:172-183)
	V2 v1 = vadd(v2(pb.Va.x, pb.Va.y), vmul(vperp(r1), pb.Va.z));
	V2 v2_ = vadd(v2(pb.Vb.x, pb.Vb.y), vmul(vperp(r2), pb.Vb.z));
	bounce = vdot(vsub(v2_, v1), n_)*e;
}

// derived_only: refresh nMass / tMass / bias of the active records for a read-back after a production step (whose row
// build computed them on the fly and never stored them); bounce needs the velocities from before the step and the bias
// impulse belongs to the solver by then: both are left alone.
__global__ void k_arb_prestep(DBodies B, DArbs A, const DSpace *__restrict__ spaces, double dt, int derived_only)
{
	int n = *A.count_ptr; if(n > A.cap) n = A.cap;
	for(int i = CPB_TID; i < n; i += CPB_NTHREADS){
	if(A.active[i] != 1){
		if(!derived_only && A.active[i] == 0 && A.state[i] == CPB200_ARB_FIRST_COLLISION) A.state[i] = CPB200_ARB_NORMAL;   // cpSpaceStep.c:283
		continue;
	}
	int ba = A.ba[i], bb = A.bb[i];
	DSpace sp = spaces[B.space[ba]];
	V2 n_ = A.n[i];
	const PrestepBodies pb = prestep_bodies(B, ba, bb);
	double e = A.e[i];
	int cnt = A.cnt[i];
	for(int k = 0; k < cnt; k++){
		int c = CIDX(A, i, k);
		double nm, tm, bi, bo;
		prestep_contact(pb, n_, e, sp.bias_coef, sp.slop, dt, A.r1[c], A.r2[c], nm, tm, bi, bo);
		A.nmass[c] = nm; A.tmass[c] = tm; A.bias[c] = bi;
		if(!derived_only){ A.jb[c] = 0.0; A.bounce[c] = bo; }
	}
	}
}

// validation hook: narrowphase of one pair (cpShapesCollide, cpShape.c:259-283)
__global__ void k_collide_one(DShapes S, DBodies B, int sa, int sb, double *out)
{
	if(CPB_TID != 0) return;
	bool swapped = false;
	if(S.type[sa] > S.type[sb]){ int t = sa; sa = sb; sb = t; swapped = true; }
	NShape a = load_nshape(S, B, sa), b = load_nshape(S, B, sb);
	Manifold m; m.id = 0;
	collide_shapes(a, b, m);
	for(int k = 0; k < 13; k++) out[k] = 0.0;
	out[0] = m.count;
	V2 n = swapped ? vneg(m.n) : m.n;
	out[1] = n.x; out[2] = n.y;
	for(int k = 0; k < m.count; k++){
		V2 p1 = m.p1[k], p2 = m.p2[k];
		V2 A_ = swapped ? p2 : p1, B_ = swapped ? p1 : p2;
		out[3 + 5*k + 0] = A_.x; out[3 + 5*k + 1] = A_.y; out[3 + 5*k + 2] = B_.x; out[3 + 5*k + 3] = B_.y;
		out[3 + 5*k + 4] = vdot(vsub(p2, p1), n);
	}
}
