// k_broad.cuh -- K3 + K4: broadphase and pair filter.
//
// Replaces cpBBTreeReindexQuery / cpSpaceHashReindexQuery / cpSweep1DReindexQuery
// (cpBBTree.c:638-654, cpSpaceHash.c:448-457, cpSweep1D.c:211-233) with a linear BVH
// that is rebuilt from scratch every step:
//   world bounds -> 32-bit Morton code of each AABB centre (space id in the high key
//   bits so batched spaces never mix) -> radix sort (prims.cuh) -> Karras radix tree
//   -> bottom-up AABB refit -> one stack traversal per active leaf.
// Candidate leaves go straight through the QueryReject rules (cpSpaceStep.c:204-232):
// closed-interval AABB test (exact: leaf boxes ARE the cached shape->bb), same body,
// cpShapeFilterReject, no-collide joints.  Membership follows SURVEY.md 8a a6/a7: A ranges
// over shapes of awake non-static bodies, B over everything; each unordered pair is
// emitted exactly once, binned by narrowphase class.
#pragma once
#include "cpb_world.h"
#include "prims.cuh"

// Deferred nodes of the depth-first traversal: at most one per level.  The depth of the radix tree is bounded by the number of
// distinct prefix lengths of its keys: 32 Morton bits + the space-id bits (<= 12 for 4096 spaces) + the index bits that break
// ties between equal keys (<= 31) = 75 < 96, so the overflow flag (bit 3) cannot be raised by any input; it stays as a guard.
#define CPB_BVH_STACK 96

CPB_DEVICE bool shape_is_active(const DBodies &B, int body){
	return B.type[body] != CPB200_BODY_STATIC && !B.sleeping[body];
}

// ---- world bounds: min/max over every shape AABB ----
#ifndef CPB_EMU
__device__ __forceinline__ void atomic_min_double(double *addr, double v){
	// order-preserving for IEEE doubles through a signed/unsigned split
	if(v >= 0.0) atomicMin((long long *)addr, __double_as_longlong(v));
	else atomicMax((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ void atomic_max_double(double *addr, double v){
	if(v >= 0.0) atomicMax((long long *)addr, __double_as_longlong(v));
	else atomicMin((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}
#else
static inline void atomic_min_double(double *addr, double v){ if(v < *addr) *addr = v; }
static inline void atomic_max_double(double *addr, double v){ if(v > *addr) *addr = v; }
#endif

__global__ void k_bounds_init(double *bounds, int *top_count)
{
	if(CPB_TID == 0){ bounds[0] = INFINITY; bounds[1] = INFINITY; bounds[2] = -INFINITY; bounds[3] = -INFINITY; *top_count = 0; }
}

__global__ void k_bounds(DShapes S, double *bounds)
{
	int i = CPB_TID;
	double l = INFINITY, b = INFINITY, r = -INFINITY, t = -INFINITY;
	for(int s = i; s < S.n; s += CPB_NTHREADS){
		double4 bb = S.bb[s];
		V2 c = v2((bb.x + bb.z)*0.5, (bb.y + bb.w)*0.5);
		if(c.x == c.x && c.y == c.y && fabs(c.x) != INFINITY && fabs(c.y) != INFINITY){
			l = fmin(l, c.x); r = fmax(r, c.x); b = fmin(b, c.y); t = fmax(t, c.y);
		}
	}
#ifndef CPB_EMU
	for(int d = 16; d > 0; d >>= 1){
		l = fmin(l, __shfl_xor_sync(0xffffffffu, l, d)); b = fmin(b, __shfl_xor_sync(0xffffffffu, b, d));
		r = fmax(r, __shfl_xor_sync(0xffffffffu, r, d)); t = fmax(t, __shfl_xor_sync(0xffffffffu, t, d));
	}
	// one set of atomics per CTA, not per warp: the four words share a sector and every warp of the grid hit it
	__shared__ double s_red[32][4];
	const int wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	if((threadIdx.x & 31) == 0){ s_red[wid][0] = l; s_red[wid][1] = b; s_red[wid][2] = r; s_red[wid][3] = t; }
	__syncthreads();
	if(threadIdx.x != 0) return;
	for(int k = 1; k < nw; k++){ l = fmin(l, s_red[k][0]); b = fmin(b, s_red[k][1]); r = fmax(r, s_red[k][2]); t = fmax(t, s_red[k][3]); }
#endif
	if(l <= r){
		atomic_min_double(&bounds[0], l); atomic_min_double(&bounds[1], b);
		atomic_max_double(&bounds[2], r); atomic_max_double(&bounds[3], t);
	}
}

CPB_DEVICE uint32_t spread16(uint32_t x){
	x &= 0xffffu;
	x = (x | (x << 8)) & 0x00ff00ffu;
	x = (x | (x << 4)) & 0x0f0f0f0fu;
	x = (x | (x << 2)) & 0x33333333u;
	x = (x | (x << 1)) & 0x55555555u;
	return x;
}

__global__ void k_morton(DShapes S, DBodies B, const double *__restrict__ bounds, uint64_t *keys, int *vals, int drop_bits, int *refit_flags)
{
	int s = CPB_TID;
	if(s >= S.n) return;
	refit_flags[s] = 0;   // arrival counters of k_bvh_refit (saves a memset launch)
	double4 bb = S.bb[s];
	double cx = (bb.x + bb.z)*0.5, cy = (bb.y + bb.w)*0.5;
	double w = bounds[2] - bounds[0], h = bounds[3] - bounds[1];
	double fx = (w > 0.0 ? (cx - bounds[0])/w : 0.0), fy = (h > 0.0 ? (cy - bounds[1])/h : 0.0);
	fx = fmin(fmax(fx, 0.0), 1.0); fy = fmin(fmax(fy, 0.0), 1.0);
	if(!(fx == fx)) fx = 0.0;
	if(!(fy == fy)) fy = 0.0;
	uint32_t qx = (uint32_t)(fx*65535.0), qy = (uint32_t)(fy*65535.0);
	uint32_t m = spread16(qx) | (spread16(qy) << 1);
	// only as many Morton bits as the shape count needs (ties keep the upload order, Karras' index fallback handles them):
	// the radix sort then skips the dropped low digits
	m &= ~((1u << drop_bits) - 1u);
	keys[s] = ((uint64_t)(uint32_t)B.space[S.body[s]] << 32) | (uint64_t)m;
	vals[s] = s;
}

// ---- Karras 2012 radix tree over the sorted keys ----
CPB_DEVICE int bvh_delta(const uint64_t *__restrict__ keys, int n, int i, int j){
	if(j < 0 || j >= n) return -1;
	uint64_t a = keys[i], b = keys[j];
	if(a == b) return 64 + __clz(i ^ j);
	return __clzll((long long)(a ^ b));
}

#define CPB_REFIT_WIN 256
#define CPB_BVH_LOCAL 0x40000000
__global__ void k_bvh_build(DBvh T)
{
	int i = CPB_TID;
	int n = T.n;
	if(i >= n - 1) return;
	const uint64_t *keys = T.keys;
	int d = (bvh_delta(keys, n, i, i + 1) - bvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
	int dmin = bvh_delta(keys, n, i, i - d);
	int lmax = 2;
	while(bvh_delta(keys, n, i, i + lmax*d) > dmin) lmax *= 2;
	int l = 0;
	for(int t = lmax/2; t >= 1; t /= 2){
		if(bvh_delta(keys, n, i, i + (l + t)*d) > dmin) l += t;
	}
	int j = i + l*d;
	int dnode = bvh_delta(keys, n, i, j);
	int s = 0;
	int t = l;
	do {
		t = (t + 1)/2;
		if(bvh_delta(keys, n, i, i + (s + t)*d) > dnode) s += t;
	} while(t > 1);
	int gamma = i + s*d + (d < 0 ? d : 0);
	int lo = (i < j ? i : j), hi = (i < j ? j : i);
	int left = (lo == gamma ? (n - 1) + gamma : gamma);
	int right = (hi == gamma + 1 ? (n - 1) + gamma + 1 : gamma + 1);
	T.left[i] = left; T.right[i] = right;
	// the parent word also says whether that parent's leaf range [lo, hi] lies inside one window of CPB_REFIT_WIN leaves:
	// such nodes (255 of 256) are refitted from shared memory by the CTA that owns the window (k_bvh_refit_fused)
	const bool local = (lo/CPB_REFIT_WIN == hi/CPB_REFIT_WIN);
	const int up = i | (local ? CPB_BVH_LOCAL : 0);
	T.parent[left] = up; T.parent[right] = up;
	if(i == 0) T.parent[0] = -1;
	// children that are the top of a window-local subtree (or a lone leaf) under a node that is not: k_bvh_refit_top starts there
	if(!local){
		if(lo/CPB_REFIT_WIN == gamma/CPB_REFIT_WIN){ int k = atomicAdd(T.top_count, 1); if(k < n) T.top_list[k] = left; }
		if((gamma + 1)/CPB_REFIT_WIN == hi/CPB_REFIT_WIN){ int k = atomicAdd(T.top_count, 1); if(k < n) T.top_list[k] = right; }
	}
}

// reset_flags: a step that keeps the tree's topology (no k_morton pass) clears the refit's arrival counters here
__global__ void k_bvh_leaves(DBvh T, DShapes S, DBodies B, int reset_flags)
{
	int i = CPB_TID;
	if(i >= T.n) return;
	if(reset_flags) T.flags[i] = 0;
	int s = T.leaf_shape[i];
	T.nbb[(T.n - 1) + i] = S.bb[s];
	int body = S.body[s];
	int sp = B.space[body];
	T.nsp[(T.n - 1) + i] = make_int2(sp, sp);
	// a query leaf i only needs partners at Morton positions > i unless the partner is inactive
	// (static / sleeping shapes never query): an all-active subtree ending at or before i is skipped
	T.nskip[(T.n - 1) + i] = (shape_is_active(B, body) ? i : 0x7fffffff);
}

#ifndef CPB_EMU
__device__ __forceinline__ double4 ld_cg4(const double4 *p){
	double2 lo = __ldcg((const double2 *)p), hi = __ldcg((const double2 *)p + 1);
	return make_double4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ int2 ld_cg_i2(const int2 *p){ return __ldcg(p); }
#else
static inline double4 ld_cg4(const double4 *p){ return *p; }
static inline int2 ld_cg_i2(const int2 *p){ return *p; }
#endif
#ifndef CPB_EMU
__device__ __forceinline__ int ld_cg_i(const int *p){ return __ldcg(p); }
#else
static inline int ld_cg_i(const int *p){ return *p; }
#endif

// pack the children's boxes / skip keys / space ranges next to their parent (traversal layout).  A separate pass:
// written from inside the refit's upward walk it lengthened that latency chain by more than this kernel costs.
__global__ void k_bvh_pack(DBvh T)
{
	int i = CPB_TID;
	if(i >= T.n - 1) return;
	int l = T.left[i], r = T.right[i];
	T.cbox[2*i] = T.nbb[l]; T.cbox[2*i + 1] = T.nbb[r];
	T.cinfo[i] = make_int4(l, r, T.nskip[l], T.nskip[r]);
	int2 a = T.nsp[l], b = T.nsp[r];
	T.cspace[i] = make_int4(a.x, a.y, b.x, b.y);
}

__global__ void k_bvh_refit(DBvh T)
{
	int i = CPB_TID;
	int n = T.n;
	if(i >= n || n < 2) return;
	int cur = T.parent[(n - 1) + i];
	while(cur >= 0){
		cur &= (CPB_BVH_LOCAL - 1);
		// one acquire-release atomic instead of fence + atomic + fence: the first child to arrive releases the box it wrote,
		// the second acquires it (the stand-alone fences were most of this kernel's 20-level latency chain)
#ifndef CPB_EMU
		int old;
		asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(old) : "l"(&T.flags[cur]) : "memory");
#else
		int old = atomicAdd(&T.flags[cur], 1);
#endif
		if(old == 0) return; // first child to arrive: the sibling's thread finishes this node
		int l = T.left[cur], r = T.right[cur];
		// children were written by other threads (possibly other SMs): read through L2
		double4 a = ld_cg4(&T.nbb[l]);
		double4 b = ld_cg4(&T.nbb[r]);
		T.nbb[cur] = make_double4(fmin(a.x, b.x), fmin(a.y, b.y), fmax(a.z, b.z), fmax(a.w, b.w));
		int2 sa = ld_cg_i2(&T.nsp[l]), sb = ld_cg_i2(&T.nsp[r]);
		T.nsp[cur] = make_int2(sa.x < sb.x ? sa.x : sb.x, sa.y > sb.y ? sa.y : sb.y);
		int ka = ld_cg_i(&T.nskip[l]), kb = ld_cg_i(&T.nskip[r]);
		T.nskip[cur] = (ka > kb ? ka : kb);
		cur = T.parent[cur];
	}
}

#ifndef CPB_EMU
// Leaves + refit + traversal layout in one pass (the production path; the three kernels above are what the emulation build
// runs).  A CTA owns a window of CPB_REFIT_WIN consecutive leaves.  Every internal node whose leaf range lies inside the
// window -- its id does too, an internal node's id is one end of its range -- is refitted from shared memory: arrival
// counters, child boxes and child links never leave the SM, so the bottom eight levels of the tree (255 of 256 nodes)
// cost shared-memory latencies instead of an L2 atomic and two L2 loads per level.  The upper levels are a second, small
// launch (k_bvh_refit_top) with the acquire-release protocol of k_bvh_refit.  Whoever
// completes a node holds both children in registers and writes the node's traversal record (k_bvh_pack's job) right
// there; node boxes go to global memory only where a non-local parent will read them.  The second arrival clears the
// global counter it used, so no pass has to reset them for the next step.
__device__ __forceinline__ void bvh_node_finish(const DBvh &T, int cur, int l, int r, double4 a, double4 b, int2 sa, int2 sb, int ka, int kb,
	double4 &box, int2 &sp, int &skip)
{
	T.cbox[2*cur] = a; T.cbox[2*cur + 1] = b;
	T.cinfo[cur] = make_int4(l, r, ka, kb);
	T.cspace[cur] = make_int4(sa.x, sa.y, sb.x, sb.y);
	box = make_double4(fmin(a.x, b.x), fmin(a.y, b.y), fmax(a.z, b.z), fmax(a.w, b.w));
	sp = make_int2(sa.x < sb.x ? sa.x : sb.x, sa.y > sb.y ? sa.y : sb.y);
	skip = (ka > kb ? ka : kb);
}

__global__ void __launch_bounds__(CPB_REFIT_WIN) k_bvh_refit_fused(DBvh T, DShapes S, DBodies B)
{
	__shared__ double4 s_box[2*CPB_REFIT_WIN];      // [0, WIN) internal nodes by id - base, [WIN, 2 WIN) leaves by position - base
	__shared__ int2 s_sp[2*CPB_REFIT_WIN];
	__shared__ int s_skip[2*CPB_REFIT_WIN];
	__shared__ int s_left[CPB_REFIT_WIN], s_right[CPB_REFIT_WIN], s_parent[CPB_REFIT_WIN], s_flag[CPB_REFIT_WIN];
	const int n = T.n, t = threadIdx.x, base = blockIdx.x*CPB_REFIT_WIN, i = base + t;
	s_flag[t] = 0;
	if(i < n - 1){ s_left[t] = T.left[i]; s_right[t] = T.right[i]; s_parent[t] = T.parent[i]; }
	__syncthreads();
	if(i >= n || n < 2) return;
	// the leaf (k_bvh_leaves)
	double4 box; int2 sp; int skip;
	{
		const int s = T.leaf_shape[i];
		box = S.bb[s];
		const int body = S.body[s];
		const int space = B.space[body];
		sp = make_int2(space, space);
		skip = (shape_is_active(B, body) ? i : 0x7fffffff);
	}
	int me = (n - 1) + i;
	s_box[CPB_REFIT_WIN + t] = box; s_sp[CPB_REFIT_WIN + t] = sp; s_skip[CPB_REFIT_WIN + t] = skip;
	int up = T.parent[me];
	// inside the window: shared memory only
	while(up >= 0 && (up & CPB_BVH_LOCAL)){
		const int cur = up & (CPB_BVH_LOCAL - 1), slot = cur - base;
		__threadfence_block();
		if(atomicAdd(&s_flag[slot], 1) == 0) return;   // first child to arrive: the sibling's thread finishes this node
		__threadfence_block();
		const int l = s_left[slot], r = s_right[slot];
		const int il = (l >= n - 1 ? CPB_REFIT_WIN + (l - (n - 1)) - base : l - base);
		const int ir = (r >= n - 1 ? CPB_REFIT_WIN + (r - (n - 1)) - base : r - base);
		const double4 a = s_box[il], b = s_box[ir];
		const int2 sa = s_sp[il], sb = s_sp[ir];
		const int ka = s_skip[il], kb = s_skip[ir];
		bvh_node_finish(T, cur, l, r, a, b, sa, sb, ka, kb, box, sp, skip);
		s_box[slot] = box; s_sp[slot] = sp; s_skip[slot] = skip;
		me = cur;
		up = s_parent[slot];
	}
	// the node this thread completed last is the child of a node outside the window (or the root): publish it for k_bvh_refit_top
	if(up >= 0){ T.nbb[me] = box; T.nsp[me] = sp; T.nskip[me] = skip; }
}

// The levels above the windows: one thread per window-top node (T.top_list, a few thousand at a million leaves -- all
// resident at once, where the leaf-parallel kernel kept every CTA alive for as long as its one surviving thread climbed).
__global__ void __launch_bounds__(128) k_bvh_refit_top(DBvh T)
{
	const int count = min(*T.top_count, T.n);
	for(int k = CPB_TID; k < count; k += CPB_NTHREADS){
		int up = T.parent[T.top_list[k]];
		while(up >= 0){
			const int cur = up & (CPB_BVH_LOCAL - 1);
			// the links are topology (constant while the tree is refitted): fetched beside the arrival atomic, not behind it --
			// two dependent L2 round trips per level (atomic, child data) instead of four
			const int l = __ldg(&T.left[cur]), r = __ldg(&T.right[cur]), next = __ldg(&T.parent[cur]);
			int old;
			asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(old) : "l"(&T.flags[cur]) : "memory");
			if(old == 0) break;      // first child to arrive: the sibling's thread finishes this node
			T.flags[cur] = 0;        // both children are in: the counter is ready for the next step
			const double4 a = ld_cg4(&T.nbb[l]), b = ld_cg4(&T.nbb[r]);
			const int2 sa = ld_cg_i2(&T.nsp[l]), sb = ld_cg_i2(&T.nsp[r]);
			const int ka = ld_cg_i(&T.nskip[l]), kb = ld_cg_i(&T.nskip[r]);
			double4 box; int2 sp; int skip;
			bvh_node_finish(T, cur, l, r, a, b, sa, sb, ka, kb, box, sp, skip);
			T.nbb[cur] = box; T.nsp[cur] = sp; T.nskip[cur] = skip;
			up = next;
		}
	}
}
#endif

// ---- K4 rules ----
CPB_DEVICE bool bb_intersects(double4 a, double4 b){
	// cpBBIntersects (cpBB.h:58-61): closed intervals
	return (a.x <= b.z && b.x <= a.z && a.y <= b.w && b.y <= a.w);
}

// Body pairs joined by a constraint with collideBodies == false (QueryRejectConstraint, cpSpaceStep.c:204-217):
// an open-addressing set of (lo << 32 | hi) + 1 keys built by the host, `mask` = capacity - 1.  One probe
// (one L2 access) in the common case -- the binary search over the sorted list it replaces was 17 dependent
// loads per candidate pair and the whole cost of the pair kernels on the Chains spaces.
CPB_DEVICE bool nocollide_lookup(const uint64_t *__restrict__ set, int mask, int ba, int bb){
	uint64_t lo = (uint64_t)(uint32_t)(ba < bb ? ba : bb), hi = (uint64_t)(uint32_t)(ba < bb ? bb : ba);
	uint64_t key = ((lo << 32) | hi) + 1ull;
	uint32_t slot = (uint32_t)mix64(key) & (uint32_t)mask;
	for(;;){
		uint64_t k = set[slot];
		if(k == key) return true;
		if(k == 0) return false;
		slot = (slot + 1) & (uint32_t)mask;
	}
}

CPB_DEVICE bool query_reject(const DShapes &S, int sa, int sb, const uint64_t *nocollide, int n_nocollide){
	int ba = S.body[sa], bb = S.body[sb];
	if(ba == bb) return true;
	// cpShapeFilterReject (chipmunk_private.h:144-155)
	uint64_t ga = S.group[sa], gb = S.group[sb];
	if(ga != 0 && ga == gb) return true;
	if((S.cat[sa] & S.mask[sb]) == 0 || (S.cat[sb] & S.mask[sa]) == 0) return true;
	// QueryRejectConstraint (cpSpaceStep.c:204-217)
	if(n_nocollide > 0 && nocollide_lookup(nocollide, n_nocollide, ba, bb)) return true;
	return false;
}

CPB_DEVICE void emit_pair(const DShapes &S, const DPairs &P, int *overflow, int sa, int sb){
	int ta = S.type[sa], tb = S.type[sb];
	// cpCollide orders by shape type (cpCollision.c:706-710); equal types: lower index first
	if(ta > tb || (ta == tb && sa > sb)){ int t = sa; sa = sb; sb = t; t = ta; ta = tb; tb = t; }
	int cls = (tb == 0 ? 0 : (ta == 0 && tb == 1 ? 1 : 2));
	// warp-ballot compaction: one atomic per class per converged warp
	int slot = -1;
	for(int c = 0; c < 3; c++){
		int sl = cpb_warp_append(&P.count[c], cls == c);
		if(cls == c) slot = sl;
	}
	if(slot < P.cap){ P.a[cls][slot] = sa; P.b[cls][slot] = sb; }
	else atomicOr((unsigned *)overflow, 1u);
}

// One thread per leaf in Morton order (neighbouring threads walk neighbouring paths).
// The warp stays converged around the traversal: every iteration each unfinished lane visits ONE node and only
// appends the leaf hits it finds to a per-warp candidate list in shared memory; when the list runs full (and at
// the end) the warp deals the candidates out one per lane and runs the expensive part -- filter gathers, constraint
// lookup, class lists -- converged, 32 pairs per ballot and global atomic.  (Inline, that part ran once per hit
// with a handful of active lanes and waited for a global atomic each time.)
#define CPB_PAIR_CAND 192     // per warp; a visit appends at most 2 per lane -> flush above CPB_PAIR_CAND - 64
#ifndef CPB_EMU
// The traversal only collects leaf hits: when a warp's list runs full (and at the end) it is copied to the global
// candidate list with one reservation.  QueryReject and the class lists are a pass of their own (k_pair_filter):
// done inside the traversal, that part needed registers the traversal cannot afford (occupancy is what hides the
// tree's latency) and put a returning atomic on a hot word behind every 32 hits.
__device__ __forceinline__ void pairs_flush(int2 *cand, int *count, int lane, const DPairs &P, int *overflow)
{
	__syncwarp();
	const int n = *count;
	int base = 0;
	if(lane == 0 && n) base = atomicAdd(&P.count[3], n);
	base = __shfl_sync(0xffffffffu, base, 0);
	for(int k = lane; k < n; k += 32){
		if(base + k < P.cap) P.cand[base + k] = cand[k];
		else atomicOr((unsigned *)overflow, 1u);
	}
	__syncwarp();
	if(lane == 0) *count = 0;
	__syncwarp();
}

// K4 over the candidate list: a warp takes 32 x CPB_FILTER_ROUNDS candidates at a time; every lane filters and
// classifies its candidates first -- one packed sector per shape (DShapes::filt), no atomic between the rounds, so
// the gathers of all rounds are in flight together -- then ONE reservation per pair class for the whole chunk
// (the list counters are single hot words), then the writes in list order.
#define CPB_FILTER_ROUNDS 4
__global__ void __launch_bounds__(128) k_pair_filter(DShapes S, DPairs P, const uint64_t *__restrict__ nocollide, int n_nocollide, int *overflow)
{
	int n = P.count[3]; if(n > P.cap) n = P.cap;
	const int lane = threadIdx.x & 31;
	const int warp = (blockIdx.x*blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x*blockDim.x) >> 5;
	const int chunk = 32*CPB_FILTER_ROUNDS;
	for(int first = warp*chunk; first < n; first += n_warps*chunk){
		int sa[CPB_FILTER_ROUNDS], sb[CPB_FILTER_ROUNDS], cl[CPB_FILTER_ROUNDS];
		double4 fa[CPB_FILTER_ROUNDS], fb[CPB_FILTER_ROUNDS];
#pragma unroll
		for(int t = 0; t < CPB_FILTER_ROUNDS; t++){
			const int k = first + lane + 32*t;
			cl[t] = -1; sa[t] = 0; sb[t] = 0;
			fa[t] = make_double4(0, 0, 0, 0); fb[t] = fa[t];
			if(k < n){ int2 c = P.cand[k]; sa[t] = c.x; sb[t] = c.y; cl[t] = 3; fa[t] = ld4_nc(&S.filt[c.x]); fb[t] = ld4_nc(&S.filt[c.y]); }
		}
		unsigned v0[CPB_FILTER_ROUNDS], v1[CPB_FILTER_ROUNDS], v2[CPB_FILTER_ROUNDS];
		int tot0 = 0, tot1 = 0, tot2 = 0;
#pragma unroll
		for(int t = 0; t < CPB_FILTER_ROUNDS; t++){
			if(cl[t] == 3){
				const unsigned long long xa = (unsigned long long)__double_as_longlong(fa[t].x), ya = (unsigned long long)__double_as_longlong(fa[t].y);
				const unsigned long long xb = (unsigned long long)__double_as_longlong(fb[t].x), yb = (unsigned long long)__double_as_longlong(fb[t].y);
				const unsigned long long ga = (unsigned long long)__double_as_longlong(fa[t].z), gb = (unsigned long long)__double_as_longlong(fb[t].z);
				const int ba = (int)(uint32_t)xa, bb = (int)(uint32_t)xb;
				int ta = (int)(xa >> 32), tb = (int)(xb >> 32);
				// QueryReject (cpSpaceStep.c:204-232): same body, same non-zero group, category / mask, a joint between the bodies
				bool rej = (ba == bb) || (ga != 0 && ga == gb) || (((uint32_t)ya & (uint32_t)(yb >> 32)) == 0) || (((uint32_t)yb & (uint32_t)(ya >> 32)) == 0);
				if(!rej && n_nocollide > 0) rej = nocollide_lookup(nocollide, n_nocollide, ba, bb);
				if(rej) cl[t] = -1;
				else {
					// cpCollide orders by shape type (cpCollision.c:706-710); equal types: lower index first
					if(ta > tb || (ta == tb && sa[t] > sb[t])){ int x = sa[t]; sa[t] = sb[t]; sb[t] = x; x = ta; ta = tb; tb = x; }
					cl[t] = (tb == 0 ? 0 : (ta == 0 && tb == 1 ? 1 : 2));
				}
			}
			v0[t] = __ballot_sync(0xffffffffu, cl[t] == 0); v1[t] = __ballot_sync(0xffffffffu, cl[t] == 1); v2[t] = __ballot_sync(0xffffffffu, cl[t] == 2);
			tot0 += __popc(v0[t]); tot1 += __popc(v1[t]); tot2 += __popc(v2[t]);
		}
		int base = 0;
		if(lane == 0 && tot0) base = atomicAdd(&P.count[0], tot0);
		if(lane == 1 && tot1) base = atomicAdd(&P.count[1], tot1);
		if(lane == 2 && tot2) base = atomicAdd(&P.count[2], tot2);
		int run0 = __shfl_sync(0xffffffffu, base, 0), run1 = __shfl_sync(0xffffffffu, base, 1), run2 = __shfl_sync(0xffffffffu, base, 2);
		const unsigned lt = (1u << lane) - 1u;
#pragma unroll
		for(int t = 0; t < CPB_FILTER_ROUNDS; t++){
			if(cl[t] >= 0){
				const int c = cl[t];
				const int slot = (c == 0 ? run0 + __popc(v0[t] & lt) : (c == 1 ? run1 + __popc(v1[t] & lt) : run2 + __popc(v2[t] & lt)));
				if(slot < P.cap){ P.a[c][slot] = sa[t]; P.b[c][slot] = sb[t]; }
				else atomicOr((unsigned *)overflow, 1u);
			}
			run0 += __popc(v0[t]); run1 += __popc(v1[t]); run2 += __popc(v2[t]);
		}
	}
}
#endif

__global__ void __launch_bounds__(128) k_bvh_pairs(DBvh T, DShapes S, DBodies B, DPairs P, const uint64_t *__restrict__ nocollide, int n_nocollide, int multi_space, int *overflow, unsigned *visits)
{
	int i = CPB_TID;
	int n = T.n;
#ifndef CPB_EMU
	__shared__ int2 s_cand[4][CPB_PAIR_CAND];
	__shared__ int s_n[4];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if(lane == 0) s_n[wid] = 0;
	__syncwarp();
	bool done = (i >= n || n < 2);
#else
	if(i >= n || n < 2) return;
	bool done = false;
#endif
	int si = 0, qsp = 0;
	double4 q = make_double4(0, 0, 0, 0);
	if(!done){
		si = T.leaf_shape[i];
		int bi = S.body[si];
		if(!shape_is_active(B, bi)) done = true;
		else { q = ld4_nc(&S.bb[si]); qsp = B.space[bi]; }
	}
#ifdef CPB_EMU
	if(done) return;
#endif
	int stack[CPB_BVH_STACK];
	int sp = 0;
	int node = 0;
#ifndef CPB_EMU
	const bool sampled = ((blockIdx.x & 15) == 0);
	unsigned n_visit = 0, n_query = (done ? 0u : 1u);
#endif
	for(;;){
#ifndef CPB_EMU
		if(!__any_sync(0xffffffffu, !done)) break;
#endif
		if(!done){
#ifndef CPB_EMU
			n_visit++;
#endif
			int4 ci = T.cinfo[node];
			double4 box[2] = {ld4_nc(&T.cbox[2*node]), ld4_nc(&T.cbox[2*node + 1])};
			int4 cs = make_int4(0, 0, 0, 0);
			if(multi_space) cs = T.cspace[node];
			int next = -1;
#pragma unroll
			for(int c = 0; c < 2; c++){
				int ch = (c ? ci.y : ci.x), skip = (c ? ci.w : ci.z);
				if(skip <= i) continue;   // every leaf below is active and at/before i: those leaves report the pair
				if(!bb_intersects(q, box[c])) continue;
				if(multi_space){ int lo = (c ? cs.z : cs.x), hi = (c ? cs.w : cs.y); if(qsp < lo || qsp > hi) continue; }
				if(ch >= n - 1){
					int sj = T.leaf_shape[ch - (n - 1)];
#ifndef CPB_EMU
					s_cand[wid][atomicAdd(&s_n[wid], 1)] = make_int2(si, sj);
#else
					if(!query_reject(S, si, sj, nocollide, n_nocollide)) emit_pair(S, P, overflow, si, sj);
#endif
				} else {
					if(next < 0) next = ch;
					else if(sp < CPB_BVH_STACK) stack[sp++] = ch;
					else atomicOr((unsigned *)overflow, 8u);
				}
			}
			if(next >= 0) node = next;
			else if(sp == 0) done = true;
			else node = stack[--sp];
		}
#ifndef CPB_EMU
		__syncwarp();
		if(s_n[wid] > CPB_PAIR_CAND - 64) pairs_flush(s_cand[wid], &s_n[wid], lane, P, overflow);
#else
		if(done) break;
#endif
	}
#ifndef CPB_EMU
	pairs_flush(s_cand[wid], &s_n[wid], lane, P, overflow);
	if(sampled){
		// tree quality sample for the host (DCounters::bvh_visits / bvh_queries)
		n_visit = __reduce_add_sync(0xffffffffu, n_visit); n_query = __reduce_add_sync(0xffffffffu, n_query);
		if(lane == 0 && n_query){ atomicAdd(&visits[0], n_visit); atomicAdd(&visits[1], n_query); }
	}
#endif
}

// ---- space-local broadphase: batched worlds of many small spaces -------------------------------------------
// A space of a few hundred shapes does not need a tree: one CTA per space stages the space's AABBs (and which
// of them are active) in shared memory and every active shape tests all others -- 10^4 box tests per space,
// no bounds / Morton / sort / build / refit / traversal launches.  Same membership rule as k_bvh_pairs (a pair
// needs one active shape; two active shapes are reported by the higher index), same QueryReject filter, same
// warp-ballot append: the pair SET is identical (tests compare it with the reference bit for bit).
struct DSpaceShapes { const int *shape0, *nshape; };   // contiguous shape range of each space

#ifndef CPB_EMU
// Two phases per space.  (1) all-pairs AABB tests from shared memory: a hit only appends (i, j) to a candidate list
// in shared memory (cheap, so the divergence of "my shape overlaps shape j" costs little).  (2) the candidates
// are dealt out to the threads one each: the expensive part -- filter gathers, constraint lookup, list append --
// runs with full, converged warps, 32 pairs per ballot and global atomic.  (Doing (2) inside (1) ran the
// expensive path once per hit with 14 of 32 lanes active on average and waited for a global atomic each time.)
// (cand_cap candidates per CTA, sized by the host: 2048 for the batched demo spaces, 4 x shapes for one larger space)
__global__ void k_sl_pairs(DSpaceShapes SS, DShapes S, DBodies B, DPairs P, const uint64_t *__restrict__ nocollide, int n_nocollide, int *overflow, int max_nshape, int cand_cap)
{
	extern __shared__ double4 s_bb[];                 // [max_nshape] AABBs, then [max_nshape] ints: active flags, then [cand_cap] candidates
	__shared__ int s_ncand;
	const int sp = blockIdx.x, s0 = SS.shape0[sp], n = SS.nshape[sp];
	int *s_act = (int *)(s_bb + max_nshape);
	unsigned *s_cand = (unsigned *)(s_act + max_nshape);
	const int CPB_SLP_CAND = cand_cap;
	if(threadIdx.x == 0) s_ncand = 0;
	for(int k = threadIdx.x; k < n; k += blockDim.x){
		s_bb[k] = S.bb[s0 + k];
		s_act[k] = shape_is_active(B, S.body[s0 + k]) ? 1 : 0;
	}
	__syncthreads();
	for(int i = threadIdx.x; i < n; i += blockDim.x){
		if(!s_act[i]) continue;
		const double4 q = s_bb[i];
		for(int j = 0; j < n; j++){
			if(j == i || (s_act[j] && j < i) || !bb_intersects(q, s_bb[j])) continue;
			int slot = atomicAdd(&s_ncand, 1);
			if(slot < CPB_SLP_CAND) s_cand[slot] = ((unsigned)i << 16) | (unsigned)j;
			else if(!query_reject(S, s0 + i, s0 + j, nocollide, n_nocollide)) emit_pair(S, P, overflow, s0 + i, s0 + j);   // list full
		}
	}
	__syncthreads();
	const int nc = (s_ncand < CPB_SLP_CAND ? s_ncand : CPB_SLP_CAND);
	for(int k = threadIdx.x; k < ((nc + 31) & ~31); k += blockDim.x){
		bool hit = (k < nc);
		int sa = 0, sb = 0;
		if(hit){
			unsigned c = s_cand[k];
			sa = s0 + (int)(c >> 16); sb = s0 + (int)(c & 0xffffu);
			hit = !query_reject(S, sa, sb, nocollide, n_nocollide);
		}
		if(__any_sync(__activemask(), hit)){ if(hit) emit_pair(S, P, overflow, sa, sb); }
	}
}
#endif
