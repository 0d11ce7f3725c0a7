// world.cu -- the C ABI of include/cpb200.h: device memory management, uploads, the step
// schedule (one CUDA stream; K1..K11 in the order of cpSpaceStep, cpSpaceStep.c:335-445) and
// read-back.  See DESIGN.md for the data layout and per-kernel roofline notes.
#include <stdarg.h>
#include <vector>
#include <algorithm>

#include "cpb_rt.h"
#include "cpb_math.h"
#include "cpb_world.h"
#include "prims.cuh"
#include "k_body.cuh"
#include "k_broad.cuh"
#include "k_arb.cuh"
#include "k_joint.cuh"
#include "k_solve.cuh"
#include "k_island.cuh"
#include "k_query.cuh"

#ifdef CPB_EMU
emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
#else
unsigned long long g_cpb_launches = 0;
#endif
#define CPB_ROW_BYTES_EST 200                 // bytes of one solver row (one contact; two-contact rows are rarer)
#ifndef CPB_L2_RESIDENT_BYTES
#define CPB_L2_RESIDENT_BYTES (80u << 20)     // of the 126 MB L2: what a pass may occupy and still be served from it
#endif
#ifndef CPB_CARRY_BLOCK
#define CPB_CARRY_BLOCK 128
#endif
#ifndef CPB_CARRY_CTAS
#define CPB_CARRY_CTAS 16
#endif
#define CPB_SL_MAX_SHAPES 512          // a space up to this many shapes is broadphased all-pairs in one CTA (k_sl_pairs).  (2048 was tried for
                                       // the 1000-body Bench scenes: one CTA needs 166 us for their 1.1 M box tests, the LBVH's 17 launches 80 us.)
#define CPB_SL_MAX_SMEM (200*1024)   // shared memory a space's velocity sectors may take in k_sl_solve

// ------------------------------------------------------------------ errors
static thread_local char g_error[512] = "";

void cpb_set_error(const char *fmt, ...)
{
	va_list ap; va_start(ap, fmt);
	vsnprintf(g_error, sizeof(g_error), fmt, ap);
	va_end(ap);
}

extern "C" const char *cpb200_last_error(void){ return g_error; }

extern "C" void *cpb200_host_alloc(size_t bytes)
{
#ifdef CPB_EMU
	return malloc(bytes);
#else
	void *p = NULL;
	if(cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess){ cudaGetLastError(); return NULL; }
	return p;
#endif
}

extern "C" void cpb200_host_free(void *p)
{
#ifdef CPB_EMU
	free(p);
#else
	if(p) cudaFreeHost(p);
#endif
}

extern "C" int cpb200_device_available(void)
{
#ifdef CPB_EMU
	return 1;
#else
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess){ cudaGetLastError(); return 0; }
	return n > 0;
#endif
}

// ------------------------------------------------------------------ stages (profiling)
enum {
	ST_INTEGRATE_POS, ST_SHAPE_CACHE, ST_BVH_KEYS, ST_BVH_SORT, ST_BVH_BUILD, ST_BVH_PAIRS,
	ST_COLLIDE, ST_ISLANDS, ST_CARRY, ST_PRESTEP, ST_COLOUR, ST_INTEGRATE_VEL, ST_SOLVE, ST_COUNT
};
static const char *g_stage_names[ST_COUNT] = {
	"integrate_pos", "shape_cache", "bvh_keys", "bvh_sort", "bvh_build", "bvh_pairs",
	"collide", "islands", "arbiter_carry", "prestep", "colour_rows", "integrate_vel", "solve"
};
extern "C" const char *cpb200_stage_name(int i){ return (i >= 0 && i < ST_COUNT) ? g_stage_names[i] : ""; }

// ------------------------------------------------------------------ world
struct AllocGroup {
	std::vector<void *> ptrs;
	unsigned long long gen = 0;   // bumped by every allocation and release: part of the step graph's signature, so a graph never
	                              // outlives a buffer it names (a freed and re-allocated array may or may not get its old address)
	void release(){ for(void *p : ptrs) cudaFree(p); ptrs.clear(); gen++; }
};

struct cpb200_world {
	int device;
	cudaStream_t stream;
	int n_spaces;
	int sm_count;
	int coop_blocks;        // co-resident CTAs of the colouring kernel
	int solve_minb;         // resident CTAs per SM the iteration kernel is compiled for (2: 128 registers, 3: 80 registers)

	std::vector<cpb200_space_params> sp;
	DSpace *d_spaces;
	bool spaces_dirty;
	double spaces_dt;

	DBodies B; AllocGroup gB;
	DShapes S; AllocGroup gS;
	DJoints J; AllocGroup gJ;
	std::vector<double> joint_error_bias;
	double joints_dt;
	uint64_t *d_nocollide; int n_nocollide;

	DArbs A[2]; DTable T[2]; AllocGroup gA;
	int cur;
	DRows R;
	DColour K; AllocGroup gK;
	DBvh bvh; AllocGroup gV;
	uint64_t *keys_b; int *vals_b; uint32_t *sort_tmp;
	DPairs P; AllocGroup gP;
	DIslands I; AllocGroup gI;
	DCounters *C;
	DCounters *hC;          // pinned
	int cap_pairs, cap_arbs;
	// object arrays are allocated with slack so that cpb200_world_append_* can add objects in place (f4)
	int cap_bodies, cap_shapes, cap_verts, cap_joints;
	std::vector<uint32_t> space_base;     // lowest shape hashid of every space (hlocal = hashid - base)
	std::vector<int> joint_base;          // first joint index of every space (colouring priorities)
	std::vector<uint64_t> nocollide_keys; // sorted body pairs joined by a constraint with collideBodies == 0
	std::vector<uint64_t> joint_nocollide; // per joint: its body-pair key + 1 if collideBodies == 0, else 0
	int user_cap_pairs, user_cap_arbs;

	uint32_t stamp;
	double curr_dt;
	uint64_t steps;
	bool cache_dirty;
	bool any_sleep_enabled;

	int solver_mode;
	int *d_order; int order_cap;
	uint64_t *d_user_order; int n_user_order; int user_order_cap;
	int *d_joint_order; int n_joint_order; int joint_order_cap;

	bool profiling;
	cudaEvent_t ev[ST_COUNT + 1];
	float stage_us[ST_COUNT];

	void *d_stage; size_t stage_bytes;   // device staging for host <-> SoA conversion kernels
	unsigned *d_barrier;    // grid barrier words of the persistent solver
	bool mid_step;          // between cpb200_world_step_collide and cpb200_world_step_finish
	bool mid_solve;         // validation hook: between cpb200_world_step_presolve and cpb200_world_step_finish
	int solver_variant;     // validation hook: 0 automatic, 1 world-wide + cached rows, 2 world-wide + streamed rows, 3 space-local
	int last_solver_path;   // what the last step ran: 0 serial, 1 world-wide coloured, 2 space-local
	// A step whose launch sequence is fixed (no collision handlers, no profiling, same dt, same buffers) is captured once
	// into a CUDA graph per arbiter-buffer parity and replayed with one launch: a 1000-body scene is ~35 kernels of a few
	// microseconds each, and their launch gaps were most of its step time.  Every counter a kernel needs lives on the device.
	// (one graph per arbiter-buffer parity and per kind of broadphase step: tree rebuilt / tree topology kept)
	struct StepGraph { cudaGraphExec_t exec; unsigned long long sig; int launches; int solver_path; } graph[4];
	unsigned long long graph_last_sig[4];  // signature of the previous step of the same slot (a graph is captured when it repeats)
	// The LBVH's topology (leaf order + Karras hierarchy) is kept for bvh_period steps and only refitted in between: a
	// refitted tree is an exact bounding hierarchy whatever its age, so the pair set does not depend on this; only the
	// traversal gets dearer as the boxes of an old topology spread.  Structural edits rebuild at once.
	bool bvh_valid; int bvh_age, bvh_period;
	bool arb_derived_stale[2];  // per arbiter buffer: the last step computed nMass / tMass / bias inside its row build; a read-back recomputes them
	int pack_ctas;              // CTAs per SM of the side-stream k_pack_warm (env CPB200_PACK_CTAS)
	bool refit_unfused;         // env CPB200_REFIT_UNFUSED (measurement switch): k_bvh_leaves + k_bvh_refit + k_bvh_pack instead of the fused kernel
	bool bvh_no_valve;          // env CPB200_BVH_NO_VALVE (measurement switch)
	double bvh_fresh_visits;    // visits per query right after a rebuild (0 = not measured): cpb200_world_sync rebuilds early when an aged tree needs 1.5x that
	bool graph_enabled;
	// per-step host I/O bound to the world (cpb200_world_bind_io): page-locked host buffers, device staging, a side stream
	const double *io_src; double *io_sink;
	double *d_io_force, *d_io_pos, *d_io_vel; int io_cap;
	cudaStream_t stream_io; cudaEvent_t ev_io_begin, ev_io_forces, ev_io_pos, ev_io_join;
	char graph_error[384];              // why capturing failed (the world then keeps launching kernel by kernel)
	unsigned long long graph_replays, graph_captures;
	double step_dt, step_dt_coef; int step_iterations;
	bool hints_valid;       // last step's colours may seed this step's colouring
	bool no_hints;          // validation hook (env CPB200_NO_HINTS): colour from scratch every step
	bool no_phase_prefetch; // experiment switch (env CPB200_NO_PHASE_PREFETCH)
	bool rows_strided;      // experiment switch (env CPB200_ROWS_STRIDED): round 1's strided record assignment in the row build
	int wl_cap; AllocGroup gW;
	int force_blocks;       // validation hook: fixed persistent grid size (0 = automatic)
	int last_active;        // active arbiters seen at the last host read-back (grid sizing hint)
	// space-local solver (k_sl_solve): usable when every space owns one contiguous, small body range
	std::vector<int> body_space;   // host copy of the bodies' space index
	bool sl_dirty, sl_ok, sl_disabled;
	int sl_max_nbody;
	DSpaceLocal SL; AllocGroup gSL;
	std::vector<int> shape_body;   // host copy of the shapes' body index
	DSpaceShapes SS; bool sl_shapes_ok; int sl_max_nshape;
	std::vector<int> joint_body;   // host copy: body a of every joint (its space has constraints)
	int *d_sl_plain, *d_sl_jointed; int n_sl_plain, n_sl_jointed;
	cudaEvent_t ev_fork_a, ev_join_a;   // k_pack_warm runs beside the broadphase (stream2)
	cudaStream_t stream2; cudaEvent_t ev_fork, ev_join;   // the two k_sl_solve launches touch disjoint spaces: run them side by side   // spaces without / with joints (k_sl_solve launches)
	uint32_t *sl_tmp;
	void *d_query; size_t query_bytes;   // device buffer for query hits (+ counters in its first 64 bytes)
	double *d_scratch;      // small scratch (collide_one output, stats)
	double *h_scratch;      // pinned
};

template <typename T>
static int dalloc(AllocGroup &g, T *&p, size_t n)
{
	void *q = NULL;
	if(n == 0) n = 1;
	cudaError_t e = cudaMalloc(&q, sizeof(T)*n);
	if(e != cudaSuccess){ cpb_set_error("cudaMalloc(%zu bytes) failed: %s", sizeof(T)*n, cudaGetErrorString(e)); p = NULL; return -1; }
	cudaMemsetAsync(q, 0, sizeof(T)*n, 0);
	g.ptrs.push_back(q);
	g.gen++;
	p = (T *)q;
	return 0;
}
#define DA(g, p, n) do { if(dalloc(g, p, (size_t)(n))) return -1; } while(0)

template <typename T>
static int upload(cpb200_world *w, T *dst, const std::vector<T> &src)
{
	if(src.empty()) return 0;
	CPB_CHECK(cudaMemcpyAsync(dst, src.data(), sizeof(T)*src.size(), cudaMemcpyHostToDevice, w->stream));
	CPB_CHECK(cudaStreamSynchronize(w->stream));
	return 0;
}

// the same without the wait: for a batch of small uploads from vectors that outlive the batch (one sync at its end)
template <typename T>
static int upload_nowait(cpb200_world *w, T *dst, const std::vector<T> &src)
{
	if(src.empty()) return 0;
	CPB_CHECK(cudaMemcpyAsync(dst, src.data(), sizeof(T)*src.size(), cudaMemcpyHostToDevice, w->stream));
	return 0;
}

template <typename T>
static int download(cpb200_world *w, std::vector<T> &dst, const T *src, size_t n)
{
	dst.resize(n);
	if(n == 0) return 0;
	CPB_CHECK(cudaMemcpyAsync(dst.data(), src, sizeof(T)*n, cudaMemcpyDeviceToHost, w->stream));
	return 0;
}

// per-contact arrays: plane k of n records lives at src + k*cap; dst = [plane 0 | plane 1]
template <typename T>
static int download2(cpb200_world *w, std::vector<T> &dst, const T *src, size_t n, int cap)
{
	dst.resize(2*n);
	if(n == 0) return 0;
	CPB_CHECK(cudaMemcpyAsync(dst.data(), src, sizeof(T)*n, cudaMemcpyDeviceToHost, w->stream));
	CPB_CHECK(cudaMemcpyAsync(dst.data() + n, src + cap, sizeof(T)*n, cudaMemcpyDeviceToHost, w->stream));
	return 0;
}

static inline int grid_for(int n, int block){ int g = cpb_div_up(n > 0 ? n : 1, block); return g; }

static int world_sync(cpb200_world *w)
{
	CPB_CHECK(cudaStreamSynchronize(w->stream));
	CPB_CHECK(cudaGetLastError());
	return 0;
}

static int alloc_arbs(cpb200_world *w, int cap);
static int counters_check(cpb200_world *w);
static int alloc_pairs(cpb200_world *w, int cap);

extern "C" cpb200_world *cpb200_world_create(int device, int n_spaces)
{
	if(n_spaces < 1){ cpb_set_error("n_spaces must be >= 1"); return NULL; }
#ifndef CPB_EMU
	int ndev = 0;
	if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0){
		cudaGetLastError();
		cpb_set_error("no CUDA device available: libcpb200 has no CPU fallback");
		return NULL;
	}
	if(device < 0 || device >= ndev){ cpb_set_error("CUDA device %d out of range (%d present)", device, ndev); return NULL; }
	if(cudaSetDevice(device) != cudaSuccess){ cpb_set_error("cudaSetDevice(%d) failed", device); return NULL; }
#endif
	cpb200_world *w = new cpb200_world();
	w->device = device;
	w->n_spaces = n_spaces;
	w->stream = 0;
	cudaStreamCreate(&w->stream);
	cudaStreamCreate(&w->stream2);
	cudaEventCreateWithFlags(&w->ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&w->ev_join, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&w->ev_fork_a, cudaEventDisableTiming); cudaEventCreateWithFlags(&w->ev_join_a, cudaEventDisableTiming);
	w->sm_count = 148;
	w->coop_blocks = 148; w->solve_minb = 2;
#ifndef CPB_EMU
	{
		cudaDeviceProp prop;
		if(cudaGetDeviceProperties(&prop, device) == cudaSuccess) w->sm_count = prop.multiProcessorCount;
		int per_sm = 1;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_colour_solve<false, true, true, 1, 2>, 256, 0);
		if(per_sm < 1) per_sm = 1;
		if(per_sm > 4) per_sm = 4;
		w->coop_blocks = w->sm_count*per_sm;
		{ const char *e = getenv("CPB200_SOLVE_MINB"); w->solve_minb = (e && atoi(e) == 3 ? 3 : 2); }
		cudaFuncSetAttribute(k_sl_solve<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CPB_SL_MAX_SMEM);
		cudaFuncSetAttribute(k_sl_solve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CPB_SL_MAX_SMEM);
		cudaFuncSetAttribute(k_sl_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, CPB_SL_MAX_SHAPES*(int)(sizeof(double4) + sizeof(int) + 4*sizeof(unsigned)));
	}
#endif
	cpb200_space_params def;
	memset(&def, 0, sizeof(def));
	def.damping = 1.0; def.sleep_time_threshold = INFINITY; def.collision_slop = 0.1;
	def.collision_bias = pow(1.0 - 0.1, 60.0); def.collision_persistence = 3; def.iterations = 10;
	w->sp.assign((size_t)n_spaces, def);
	w->d_spaces = NULL; w->spaces_dirty = true; w->spaces_dt = 0.0;
	memset(&w->B, 0, sizeof(w->B)); memset(&w->S, 0, sizeof(w->S)); memset(&w->J, 0, sizeof(w->J));
	memset(w->A, 0, sizeof(w->A)); memset(w->T, 0, sizeof(w->T)); memset(&w->R, 0, sizeof(w->R));
	memset(&w->K, 0, sizeof(w->K)); memset(&w->bvh, 0, sizeof(w->bvh)); memset(&w->P, 0, sizeof(w->P));
	memset(&w->I, 0, sizeof(w->I));
	w->joints_dt = 0.0; w->d_nocollide = NULL; w->n_nocollide = 0;
	w->cur = 0; w->keys_b = NULL; w->vals_b = NULL; w->sort_tmp = NULL;
	w->cap_pairs = 0; w->cap_arbs = 0; w->user_cap_pairs = 0; w->user_cap_arbs = 0;
	w->cap_bodies = w->cap_shapes = w->cap_verts = w->cap_joints = 0;
	w->stamp = 0; w->curr_dt = 0.0; w->steps = 0; w->cache_dirty = true; w->any_sleep_enabled = false;
	w->solver_mode = 0; w->d_order = NULL; w->order_cap = 0; w->d_user_order = NULL; w->n_user_order = 0; w->user_order_cap = 0;
	w->d_joint_order = NULL; w->n_joint_order = 0; w->joint_order_cap = 0;
	w->profiling = false;
	for(int i = 0; i <= ST_COUNT; i++) cudaEventCreate(&w->ev[i]);
	memset(w->stage_us, 0, sizeof(w->stage_us));
	void *p = NULL;
	cudaMalloc(&p, sizeof(DSpace)*(size_t)n_spaces); w->d_spaces = (DSpace *)p;
	cudaMalloc(&p, sizeof(DCounters)); w->C = (DCounters *)p; cudaMemsetAsync(w->C, 0, sizeof(DCounters), w->stream);
	if(w->C && getenv("CPB200_NO_GJK_STAGE")){ int one = 1; cudaMemcpyAsync(&w->C->no_gjk_stage, &one, sizeof(int), cudaMemcpyHostToDevice, w->stream); cudaStreamSynchronize(w->stream); }
	cudaMallocHost(&p, sizeof(DCounters)); w->hC = (DCounters *)p; memset(w->hC, 0, sizeof(DCounters));
	cudaMalloc(&p, sizeof(unsigned)*64); w->d_barrier = (unsigned *)p; cudaMemsetAsync(w->d_barrier, 0, sizeof(unsigned)*64, w->stream);
	w->no_hints = (getenv("CPB200_NO_HINTS") != NULL); w->no_phase_prefetch = (getenv("CPB200_NO_PHASE_PREFETCH") != NULL); w->rows_strided = (getenv("CPB200_ROWS_STRIDED") != NULL);
	w->d_query = NULL; w->query_bytes = 0;
	w->mid_solve = false; w->solver_variant = 0; w->last_solver_path = 0;
	memset(w->graph, 0, sizeof(w->graph)); memset(w->graph_last_sig, 0, sizeof(w->graph_last_sig)); w->graph_replays = w->graph_captures = 0;
	w->arb_derived_stale[0] = w->arb_derived_stale[1] = false;
	w->bvh_valid = false; w->bvh_age = 0; w->bvh_period = 8; w->bvh_fresh_visits = 0.0; w->bvh_no_valve = (getenv("CPB200_BVH_NO_VALVE") != NULL); w->refit_unfused = (getenv("CPB200_REFIT_UNFUSED") != NULL);
	w->pack_ctas = 3; { const char *e = getenv("CPB200_PACK_CTAS"); if(e && atoi(e) >= 1) w->pack_ctas = atoi(e); }
	{ const char *e = getenv("CPB200_BVH_PERIOD"); if(e && atoi(e) >= 1) w->bvh_period = atoi(e); }
	w->graph_enabled = (getenv("CPB200_NO_GRAPH") == NULL); w->graph_error[0] = 0;
	w->io_src = NULL; w->io_sink = NULL; w->d_io_force = w->d_io_pos = w->d_io_vel = NULL; w->io_cap = 0;
	cudaStreamCreate(&w->stream_io);
	cudaEventCreateWithFlags(&w->ev_io_begin, cudaEventDisableTiming); cudaEventCreateWithFlags(&w->ev_io_forces, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&w->ev_io_pos, cudaEventDisableTiming); cudaEventCreateWithFlags(&w->ev_io_join, cudaEventDisableTiming);
	w->mid_step = false; w->step_dt = 0.0; w->step_dt_coef = 0.0; w->step_iterations = 0;
	w->sl_dirty = true; w->bvh_valid = false; w->sl_ok = false; w->sl_disabled = (getenv("CPB200_NO_SPACE_LOCAL") != NULL); w->sl_max_nbody = 0;
	memset(&w->SL, 0, sizeof(w->SL)); w->sl_tmp = NULL;
	memset(&w->SS, 0, sizeof(w->SS)); w->sl_shapes_ok = false; w->sl_max_nshape = 0;
	w->d_sl_plain = w->d_sl_jointed = NULL; w->n_sl_plain = w->n_sl_jointed = 0;
	w->last_active = 0; w->force_blocks = 0; w->hints_valid = false; w->wl_cap = 0; w->d_stage = NULL; w->stage_bytes = 0;
	cudaMalloc(&p, sizeof(double)*64); w->d_scratch = (double *)p;
	cudaMallocHost(&p, sizeof(double)*64); w->h_scratch = (double *)p;
	if(!w->d_spaces || !w->C || !w->hC){ cpb_set_error("device allocation failed"); delete w; return NULL; }
	return w;
}

extern "C" void cpb200_world_destroy(cpb200_world *w)
{
	if(!w) return;
	cudaSetDevice(w->device);
	cudaStreamSynchronize(w->stream);
#ifndef CPB_EMU
	for(int k = 0; k < 4; k++) if(w->graph[k].exec) cudaGraphExecDestroy(w->graph[k].exec);
#endif
	w->gB.release(); w->gS.release(); w->gJ.release(); w->gA.release(); w->gK.release(); w->gV.release(); w->gP.release(); w->gI.release(); w->gW.release();
	cudaFree(w->d_barrier);
	if(w->d_stage) cudaFree(w->d_stage);
	cudaFree(w->d_spaces); cudaFree(w->C); cudaFreeHost(w->hC); cudaFree(w->d_scratch); cudaFreeHost(w->h_scratch);
	if(w->d_query) cudaFree(w->d_query);
	if(w->d_order) cudaFree(w->d_order);
	if(w->d_user_order) cudaFree(w->d_user_order);
	if(w->d_joint_order) cudaFree(w->d_joint_order);
	if(w->d_nocollide) cudaFree(w->d_nocollide);
	for(int i = 0; i <= ST_COUNT; i++) cudaEventDestroy(w->ev[i]);
	cudaStreamDestroy(w->stream2); cudaEventDestroy(w->ev_fork); cudaEventDestroy(w->ev_join); cudaEventDestroy(w->ev_fork_a); cudaEventDestroy(w->ev_join_a);
	cudaStreamDestroy(w->stream_io); cudaEventDestroy(w->ev_io_begin); cudaEventDestroy(w->ev_io_forces); cudaEventDestroy(w->ev_io_pos); cudaEventDestroy(w->ev_io_join);
	if(w->d_io_force) cudaFree(w->d_io_force);
	if(w->d_io_pos) cudaFree(w->d_io_pos);
	if(w->d_io_vel) cudaFree(w->d_io_vel);
	cudaStreamDestroy(w->stream);
	delete w;
}

extern "C" int cpb200_world_set_space_params(cpb200_world *w, int space, const cpb200_space_params *p)
{
	if(!w || space < 0 || space >= w->n_spaces){ cpb_set_error("bad space index"); return -1; }
	w->sp[(size_t)space] = *p;
	w->spaces_dirty = true;
	return 0;
}

static int refresh_spaces(cpb200_world *w, double dt)
{
	if(!w->spaces_dirty && w->spaces_dt == dt) return 0;
	std::vector<DSpace> h((size_t)w->n_spaces);
	w->any_sleep_enabled = false;
	for(int i = 0; i < w->n_spaces; i++){
		const cpb200_space_params &p = w->sp[(size_t)i];
		DSpace &d = h[(size_t)i];
		d.gravity = v2(p.gravity[0], p.gravity[1]);
		d.damping_dt = pow(p.damping, dt);                 // cpSpaceStep.c:399
		d.bias_coef = 1.0 - pow(p.collision_bias, dt);     // cpSpaceStep.c:384
		d.slop = p.collision_slop;
		d.idle_speed = p.idle_speed_threshold;
		d.sleep_threshold = p.sleep_time_threshold;
		d.persistence = p.collision_persistence;
		d.iterations = p.iterations;
		if(p.sleep_time_threshold != INFINITY) w->any_sleep_enabled = true;
	}
	CPB_CHECK(cudaMemcpyAsync(w->d_spaces, h.data(), sizeof(DSpace)*h.size(), cudaMemcpyHostToDevice, w->stream));
	CPB_CHECK(cudaStreamSynchronize(w->stream));
	w->spaces_dirty = false; w->spaces_dt = dt;
	return 0;
}

// ------------------------------------------------------------------ uploads
// Host records travel as ONE copy into a device staging buffer and are scattered into / gathered from the
// SoA arrays by a kernel (no per-field host loops, no per-field cudaMemcpy).
static int stage_reserve(cpb200_world *w, size_t bytes)
{
	if(bytes <= w->stage_bytes) return 0;
	if(w->d_stage) cudaFree(w->d_stage);
	w->d_stage = NULL; w->stage_bytes = 0;
	void *p = NULL;
	CPB_CHECK(cudaMalloc(&p, bytes + bytes/4 + 256));
	w->d_stage = p; w->stage_bytes = bytes + bytes/4 + 256;
	return 0;
}

__global__ void k_unpack_bodies(DBodies B, const cpb200_body_desc *__restrict__ src, int first, int n, int n_spaces, int *bad)
{
	int k = CPB_TID;
	if(k >= n) return;
	cpb200_body_desc d = src[k];
	int i = first + k;
	if(d.space < 0 || d.space >= n_spaces){ *bad = 1; d.space = 0; }
	B.pos[i] = v2(d.p[0], d.p[1]); B.ang[i] = d.a; B.rot[i] = v2(d.rot[0], d.rot[1]); B.cog[i] = v2(d.cog[0], d.cog[1]);
	bool dyn = (d.type == CPB200_BODY_DYNAMIC);
	// the spare lanes of the two velocity sectors carry (m_inv, i_inv): the solver then touches two
	// 32-byte sectors per body instead of three
	B.V[i] = make_double4(d.v[0], d.v[1], d.w, dyn ? 1.0/d.m : 0.0); B.VB[i] = make_double4(d.v_bias[0], d.v_bias[1], d.w_bias, dyn ? 1.0/d.i : 0.0);
	// m_inv / i_inv exactly as cpBodySetMass / SetMoment compute them (cpBody.c:246-270): 1/m, 0 for infinite mass
	B.MI[i] = v2(dyn ? 1.0/d.m : 0.0, dyn ? 1.0/d.i : 0.0);
	B.M[i] = v2(d.m, d.i);
	B.force[i] = v2(d.f[0], d.f[1]); B.torque[i] = d.t; B.idle[i] = d.idle_time;
	B.type[i] = d.type; B.space[i] = d.space; B.sleeping[i] = d.sleeping; B.sgroup[i] = d.sleep_group; B.custom[i] = d.custom;
	// translation part of SetTransform from the host-supplied rotation (cpBody.c:347-357)
	V2 p = v2(d.p[0], d.p[1]), rot = v2(d.rot[0], d.rot[1]), cg = v2(d.cog[0], d.cog[1]);
	B.txy[i] = v2(p.x - (cg.x*rot.x - cg.y*rot.y), p.y - (cg.x*rot.y + cg.y*rot.x));
}

__global__ void k_pack_body_state(DBodies B, cpb200_body_state *__restrict__ dst, int first, int n)
{
	int k = CPB_TID;
	if(k >= n) return;
	int i = first + k;
	cpb200_body_state o;
	V2 p = B.pos[i], rot = B.rot[i]; double4 V = B.V[i];
	o.p[0] = p.x; o.p[1] = p.y; o.v[0] = V.x; o.v[1] = V.y; o.a = B.ang[i]; o.w = V.z;
	o.rot[0] = rot.x; o.rot[1] = rot.y; o.idle_time = B.idle[i]; o.sleeping = B.sleeping[i]; o.sleep_group = B.sgroup[i];
	dst[k] = o;
}

__global__ void k_pack_body_bias(DBodies B, double *__restrict__ dst, int first, int n)
{
	int k = CPB_TID;
	if(k >= n) return;
	double4 VB = B.VB[first + k];
	dst[3*k] = VB.x; dst[3*k + 1] = VB.y; dst[3*k + 2] = VB.z;
}

__global__ void k_set_forces(DBodies B, const double *__restrict__ fxyt, int first, int n)
{
	int k = CPB_TID;
	if(k >= n) return;
	B.force[first + k] = v2(fxyt[3*k], fxyt[3*k + 1]);
	B.torque[first + k] = fxyt[3*k + 2];
}

extern "C" int cpb200_world_set_bodies(cpb200_world *w, int n, const cpb200_body_desc *bodies)
{
	if(!w || n < 0){ cpb_set_error("bad arguments"); return -1; }
	cudaSetDevice(w->device);
	if(world_sync(w)) return -1;
	w->gB.release();
	DBodies &B = w->B;
	B.n = n;
	const int cap = n + n/4 + 256;
	w->cap_bodies = cap;
	DA(w->gB, B.pos, cap); DA(w->gB, B.ang, cap); DA(w->gB, B.rot, cap); DA(w->gB, B.txy, cap); DA(w->gB, B.cog, cap);
	DA(w->gB, B.V, cap); DA(w->gB, B.VB, cap); DA(w->gB, B.MI, cap); DA(w->gB, B.M, cap); DA(w->gB, B.force, cap);
	DA(w->gB, B.torque, cap); DA(w->gB, B.idle, cap); DA(w->gB, B.type, cap); DA(w->gB, B.space, cap);
	DA(w->gB, B.sleeping, cap); DA(w->gB, B.sgroup, cap); DA(w->gB, B.custom, cap);
	w->gK.release();
	DA(w->gK, w->K.claim, cap); DA(w->gK, w->K.bmask, cap);
	DA(w->gK, w->K.ccount, CPB_MAX_COLOURS + 1); DA(w->gK, w->K.cstart, CPB_MAX_COLOURS + 1); DA(w->gK, w->K.ccursor, CPB_MAX_COLOURS + 1);
	DA(w->gK, w->K.jcount, CPB_MAX_COLOURS + 1); DA(w->gK, w->K.jstart, CPB_MAX_COLOURS + 1); DA(w->gK, w->K.jcursor, CPB_MAX_COLOURS + 1);
	DA(w->gK, w->K.wl_n, CPB_MAX_COLOUR_ROUNDS + 2); DA(w->gK, w->K.prof, 8);
	w->hints_valid = false; // body types / masses may have changed: colour from scratch once
	w->io_src = NULL; w->io_sink = NULL;   // bound host buffers were sized for the old body count: bind again
	w->body_space.clear(); w->sl_dirty = true; w->bvh_valid = false;
	w->gI.release();
	DA(w->gI, w->I.parent, cap); DA(w->gI, w->I.wake, cap); DA(w->gI, w->I.comp_active, cap); DA(w->gI, w->I.woken, cap); DA(w->gI, w->I.touch, cap); DA(w->gI, w->I.any_woken, 4); DA(w->gI, w->I.flags, 8);
	int r = cpb200_world_update_bodies(w, 0, n, bodies);
	w->cache_dirty = true;
	return r;
}

static int upload_body_range(cpb200_world *w, int first, int n, const cpb200_body_desc *bodies, bool invalidate_hints);

extern "C" int cpb200_world_update_bodies(cpb200_world *w, int first, int n, const cpb200_body_desc *bodies)
{
	if(!w || first < 0 || n < 0 || first + n > w->B.n){ cpb_set_error("body range out of bounds"); return -1; }
	return upload_body_range(w, first, n, bodies, true);
}

/* f4: bodies appended behind the existing ones; nothing that is already on the device moves or is re-uploaded.
 * Returns 1 (and changes nothing) when the arrays' slack is used up: the caller then re-uploads with cpb200_world_set_bodies,
 * which allocates new slack. */
extern "C" int cpb200_world_append_bodies(cpb200_world *w, int n, const cpb200_body_desc *bodies)
{
	if(!w || n < 0 || (n > 0 && !bodies)){ cpb_set_error("bad arguments"); return -1; }
	if(n == 0) return 0;
	if(w->mid_step || w->mid_solve){ cpb_set_error("structural edit inside a split step"); return -1; }
	if(w->B.n + n > w->cap_bodies || (int)w->body_space.size() != w->B.n) return 1;
	const int first = w->B.n;
	w->B.n += n;
	w->body_space.resize((size_t)w->B.n, -1);
	w->io_src = NULL; w->io_sink = NULL;
	int rc = upload_body_range(w, first, n, bodies, false);
	if(rc){ w->B.n = first; w->body_space.resize((size_t)first); }
	return rc;
}

static int upload_body_range(cpb200_world *w, int first, int n, const cpb200_body_desc *bodies, bool invalidate_hints)
{
	if(n == 0) return 0;
	cudaSetDevice(w->device);
	size_t bytes = sizeof(cpb200_body_desc)*(size_t)n;
	if((int)w->body_space.size() != w->B.n){ w->body_space.assign((size_t)w->B.n, 0); w->sl_dirty = true; w->bvh_valid = false; }
	for(int i = 0; i < n; i++){
		if(w->body_space[(size_t)(first + i)] != (int)bodies[i].space){ w->body_space[(size_t)(first + i)] = (int)bodies[i].space; w->sl_dirty = true; w->bvh_valid = false; }
	}
	if(stage_reserve(w, bytes + 64)) return -1;
	int *bad = (int *)((char *)w->d_stage + ((bytes + 15) & ~(size_t)15));
	CPB_CHECK(cudaMemsetAsync(bad, 0, sizeof(int), w->stream));
	CPB_CHECK(cudaMemcpyAsync(w->d_stage, bodies, bytes, cudaMemcpyHostToDevice, w->stream));
	LAUNCH(k_unpack_bodies, grid_for(n, 128), 128, w->stream, w->B, (const cpb200_body_desc *)w->d_stage, first, n, w->n_spaces, bad);
	int h_bad = 0;
	CPB_CHECK(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
	w->cache_dirty = true;
	// a body may have changed between dynamic and static / kinematic: two constraints that kept last step's colour
	// could then share a body that is now written -- colour from scratch once (appended bodies have no constraints yet)
	if(invalidate_hints) w->hints_valid = false;
	if(world_sync(w)) return -1;
	if(h_bad){ cpb_set_error("a body names a space index outside [0, %d)", w->n_spaces); return -1; }
	return 0;
}

/* Forces only: the common per-step host input (cpBodySetForce / cpBodyApplyForce*, cpBody.c:420-424, 534-543).
 * fxyt[n][3] = f.x f.y torque. */
extern "C" int cpb200_world_set_body_forces(cpb200_world *w, int first, int n, const double *fxyt)
{
	if(!w || first < 0 || n < 0 || first + n > w->B.n){ cpb_set_error("body range out of bounds"); return -1; }
	if(n == 0) return 0;
	cudaSetDevice(w->device);
	size_t bytes = sizeof(double)*3*(size_t)n;
	if(stage_reserve(w, bytes)) return -1;
	CPB_CHECK(cudaMemcpyAsync(w->d_stage, fxyt, bytes, cudaMemcpyHostToDevice, w->stream));
	LAUNCH(k_set_forces, grid_for(n, 256), 256, w->stream, w->B, (const double *)w->d_stage, first, n);
	return world_sync(w);
}

__global__ void k_touch_bodies(DBodies B, const int *__restrict__ idx, int n)
{
	int k = CPB_TID;
	if(k >= n) return;
	int i = idx[k];
	if(i >= 0 && i < B.n && !B.sleeping[i]) B.idle[i] = 0.0;
}

extern "C" int cpb200_world_touch_bodies(cpb200_world *w, int n, const int32_t *indices)
{
	if(!w || n < 0 || (n > 0 && !indices)){ cpb_set_error("bad arguments"); return -1; }
	if(n == 0) return 0;
	cudaSetDevice(w->device);
	size_t bytes = sizeof(int32_t)*(size_t)n;
	if(stage_reserve(w, bytes)) return -1;
	CPB_CHECK(cudaMemcpyAsync(w->d_stage, indices, bytes, cudaMemcpyHostToDevice, w->stream));
	LAUNCH(k_touch_bodies, grid_for(n, 256), 256, w->stream, w->B, (const int *)w->d_stage, n);
	return world_sync(w);
}

// ---- per-step host I/O overlapped with the step (cpb200_world_bind_io) ----
__global__ void k_pack_pos(DBodies B, double *__restrict__ dst)
{
	int i = CPB_TID;
	if(i >= B.n) return;
	V2 p = B.pos[i];
	dst[3*(size_t)i] = p.x; dst[3*(size_t)i + 1] = p.y; dst[3*(size_t)i + 2] = B.ang[i];
}

__global__ void k_pack_vel(DBodies B, double *__restrict__ dst)
{
	int i = CPB_TID;
	if(i >= B.n) return;
	double4 V = B.V[i];
	dst[3*(size_t)i] = V.x; dst[3*(size_t)i + 1] = V.y; dst[3*(size_t)i + 2] = V.z;
}

static bool host_ptr_is_pinned(const void *p)
{
#ifdef CPB_EMU
	return true;
#else
	cudaPointerAttributes a;
	if(cudaPointerGetAttributes(&a, p) != cudaSuccess){ cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeHost;
#endif
}

extern "C" int cpb200_world_bind_io(cpb200_world *w, const double *forces_fxyt, double *state_out)
{
	if(!w){ cpb_set_error("null world"); return -1; }
	if(w->mid_step || w->mid_solve){ cpb_set_error("cpb200_world_bind_io inside a split step"); return -1; }
	cudaSetDevice(w->device);
	if(world_sync(w)) return -1;
	if((forces_fxyt && !host_ptr_is_pinned(forces_fxyt)) || (state_out && !host_ptr_is_pinned(state_out))){
		cpb_set_error("cpb200_world_bind_io needs page-locked host buffers (cpb200_host_alloc): the copies run beside the step's kernels");
		return -1;
	}
	if((forces_fxyt || state_out) && w->io_cap < w->B.n){
		if(w->d_io_force) cudaFree(w->d_io_force);
		if(w->d_io_pos) cudaFree(w->d_io_pos);
		if(w->d_io_vel) cudaFree(w->d_io_vel);
		w->d_io_force = w->d_io_pos = w->d_io_vel = NULL; w->io_cap = 0;
		void *p = NULL; size_t bytes = sizeof(double)*3*(size_t)std::max(w->B.n, 1);
		CPB_CHECK(cudaMalloc(&p, bytes)); w->d_io_force = (double *)p;
		CPB_CHECK(cudaMalloc(&p, bytes)); w->d_io_pos = (double *)p;
		CPB_CHECK(cudaMalloc(&p, bytes)); w->d_io_vel = (double *)p;
		w->io_cap = w->B.n;
	}
	w->io_src = forces_fxyt; w->io_sink = state_out;
	return 0;
}

// ---- host callbacks in the middle of a step (custom integrators, spring force functions) ----
__global__ void k_pack_body_state_indexed(DBodies B, cpb200_body_state *__restrict__ dst, const int *__restrict__ idx, int n)
{
	int k = CPB_TID;
	if(k >= n) return;
	int i = idx[k];
	cpb200_body_state o;
	memset(&o, 0, sizeof(o));
	if(i >= 0 && i < B.n){
		V2 p = B.pos[i], rot = B.rot[i]; double4 V = B.V[i];
		o.p[0] = p.x; o.p[1] = p.y; o.v[0] = V.x; o.v[1] = V.y; o.a = B.ang[i]; o.w = V.z;
		o.rot[0] = rot.x; o.rot[1] = rot.y; o.idle_time = B.idle[i]; o.sleeping = B.sleeping[i]; o.sleep_group = B.sgroup[i];
	}
	dst[k] = o;
}

__global__ void k_set_velocities_indexed(DBodies B, const int *__restrict__ idx, const double *__restrict__ vxyw, int n)
{
	int k = CPB_TID;
	if(k >= n) return;
	int i = idx[k];
	if(i < 0 || i >= B.n) return;
	double4 V = B.V[i];
	B.V[i] = make_double4(vxyw[3*k], vxyw[3*k + 1], vxyw[3*k + 2], V.w);
	B.force[i] = v2(0.0, 0.0); B.torque[i] = 0.0;
}

__global__ void k_set_spring_forces(DJoints J, const int *__restrict__ idx, const double *__restrict__ f, int n)
{
	int k = CPB_TID;
	if(k >= n) return;
	int j = idx[k];
	if(j >= 0 && j < J.n) J.jspring[j] = v2(f[k], 1.0);
}

// [indices | payload] in one staging buffer: returns the device payload pointer
static int stage_indexed(cpb200_world *w, int n, const int32_t *indices, const void *payload, size_t payload_bytes, const int **d_idx, void **d_payload)
{
	size_t ib = (sizeof(int32_t)*(size_t)n + 63) & ~(size_t)63;
	if(stage_reserve(w, ib + payload_bytes + 64)) return -1;
	CPB_CHECK(cudaMemcpyAsync(w->d_stage, indices, sizeof(int32_t)*(size_t)n, cudaMemcpyHostToDevice, w->stream));
	*d_idx = (const int *)w->d_stage;
	*d_payload = (char *)w->d_stage + ib;
	if(payload && payload_bytes) CPB_CHECK(cudaMemcpyAsync(*d_payload, payload, payload_bytes, cudaMemcpyHostToDevice, w->stream));
	return 0;
}

extern "C" int cpb200_world_get_bodies_indexed(cpb200_world *w, int n, const int32_t *indices, cpb200_body_state *out)
{
	if(!w || n < 0 || (n > 0 && (!indices || !out))){ cpb_set_error("bad arguments"); return -1; }
	if(n == 0) return 0;
	cudaSetDevice(w->device);
	const int *d_idx; void *d_out;
	if(stage_indexed(w, n, indices, NULL, sizeof(cpb200_body_state)*(size_t)n, &d_idx, &d_out)) return -1;
	LAUNCH(k_pack_body_state_indexed, grid_for(n, 128), 128, w->stream, w->B, (cpb200_body_state *)d_out, d_idx, n);
	CPB_CHECK(cudaMemcpyAsync(out, d_out, sizeof(cpb200_body_state)*(size_t)n, cudaMemcpyDeviceToHost, w->stream));
	return world_sync(w);
}

extern "C" int cpb200_world_set_body_velocities_indexed(cpb200_world *w, int n, const int32_t *indices, const double *vxyw)
{
	if(!w || n < 0 || (n > 0 && (!indices || !vxyw))){ cpb_set_error("bad arguments"); return -1; }
	if(n == 0) return 0;
	cudaSetDevice(w->device);
	const int *d_idx; void *d_v;
	if(stage_indexed(w, n, indices, vxyw, sizeof(double)*3*(size_t)n, &d_idx, &d_v)) return -1;
	LAUNCH(k_set_velocities_indexed, grid_for(n, 128), 128, w->stream, w->B, d_idx, (const double *)d_v, n);
	return world_sync(w);
}

extern "C" int cpb200_world_set_spring_forces(cpb200_world *w, int n, const int32_t *joints, const double *f)
{
	if(!w || n < 0 || (n > 0 && (!joints || !f))){ cpb_set_error("bad arguments"); return -1; }
	if(n == 0) return 0;
	cudaSetDevice(w->device);
	const int *d_idx; void *d_f;
	if(stage_indexed(w, n, joints, f, sizeof(double)*(size_t)n, &d_idx, &d_f)) return -1;
	LAUNCH(k_set_spring_forces, grid_for(n, 128), 128, w->stream, w->J, d_idx, (const double *)d_f, n);
	return world_sync(w);
}

extern "C" int cpb200_world_reserve(cpb200_world *w, int max_pairs, int max_arbiters)
{
	if(!w) return -1;
	w->user_cap_pairs = max_pairs; w->user_cap_arbs = max_arbiters;
	return 0;
}

static int ensure_worklists(cpb200_world *w, int need)
{
	if(need <= w->wl_cap) return 0;
	w->gW.release();
	DA(w->gW, w->K.wl[0], need); DA(w->gW, w->K.wl[1], need);
	w->wl_cap = need;
	return 0;
}

static int alloc_pairs(cpb200_world *w, int cap)
{
	w->gP.release();
	w->P.cap = cap;
	DA(w->gP, w->P.count, 4);
	for(int c = 0; c < 3; c++){ DA(w->gP, w->P.a[c], cap); DA(w->gP, w->P.b[c], cap); }
	DA(w->gP, w->P.cand, cap);
	w->cap_pairs = cap;
	return 0;
}

static int alloc_arbs(cpb200_world *w, int cap)
{
	w->gA.release();
	uint32_t tcap = 64;
	while(tcap < (uint32_t)cap*2u) tcap <<= 1;
	for(int k = 0; k < 2; k++){
		DArbs &A = w->A[k];
		A.cap = cap;
		DA(w->gA, A.count_ptr, 4);
		DA(w->gA, A.key, cap); DA(w->gA, A.sa, cap); DA(w->gA, A.sb, cap); DA(w->gA, A.ba, cap); DA(w->gA, A.bb, cap);
		DA(w->gA, A.cnt, cap); DA(w->gA, A.state, cap); DA(w->gA, A.stamp, cap); DA(w->gA, A.active, cap); DA(w->gA, A.seen, cap);
		DA(w->gA, A.gjkid, cap); DA(w->gA, A.n, cap); DA(w->gA, A.e, cap); DA(w->gA, A.u, cap); DA(w->gA, A.svr, cap);
		DA(w->gA, A.r1, 2*(size_t)cap); DA(w->gA, A.r2, 2*(size_t)cap);
		DA(w->gA, A.nmass, 2*(size_t)cap); DA(w->gA, A.tmass, 2*(size_t)cap); DA(w->gA, A.bounce, 2*(size_t)cap); DA(w->gA, A.bias, 2*(size_t)cap);
		DA(w->gA, A.jn, 2*(size_t)cap); DA(w->gA, A.jt, 2*(size_t)cap); DA(w->gA, A.jb, 2*(size_t)cap); DA(w->gA, A.hash, 2*(size_t)cap);
		DA(w->gA, A.colour, cap); DA(w->gA, A.pri, cap); DA(w->gA, A.hint, cap); DA(w->gA, A.warm, 2*(size_t)cap);
		DTable &T = w->T[k];
		T.mask = tcap - 1;
		DA(w->gA, T.slots, tcap);
		DA(w->gA, T.dmask, 4);
		CPB_CHECK(cudaMemcpyAsync(T.dmask, &T.mask, sizeof(uint32_t), cudaMemcpyHostToDevice, 0));
		CPB_CHECK(cudaStreamSynchronize(0));
	}
	DRows &R = w->R;
	R.cap = cap;
	DA(w->gA, R.arb, cap); DA(w->gA, R.ba, cap); DA(w->gA, R.bb, cap); DA(w->gA, R.cnt, cap);
	DA(w->gA, R.n, cap); DA(w->gA, R.svr, cap); DA(w->gA, R.u, cap);
	DA(w->gA, R.r1, 2*(size_t)cap); DA(w->gA, R.r2, 2*(size_t)cap);
	DA(w->gA, R.nmass, 2*(size_t)cap); DA(w->gA, R.tmass, 2*(size_t)cap); DA(w->gA, R.bounce, 2*(size_t)cap); DA(w->gA, R.bias, 2*(size_t)cap);
	DA(w->gA, R.jn, 2*(size_t)cap); DA(w->gA, R.jt, 2*(size_t)cap); DA(w->gA, R.jb, 2*(size_t)cap);
	DA(w->gA, R.hdr, cap); DA(w->gA, R.nsv, cap);
	DA(w->gA, R.r12, 2*(size_t)cap); DA(w->gA, R.mass, 2*(size_t)cap); DA(w->gA, R.imp, 2*(size_t)cap);
	w->cap_arbs = cap;
	w->cur = 0;
	return 0;
}

extern "C" int cpb200_world_set_shapes(cpb200_world *w, int n, const cpb200_shape_desc *shapes, int n_verts, const double *verts_xy)
{
	if(!w || n < 0 || n_verts < 0){ cpb_set_error("bad arguments"); return -1; }
	cudaSetDevice(w->device);
	if(world_sync(w)) return -1;
	size_t N = (size_t)n, NV = (size_t)n_verts;
	// arbiter records refer to shapes/bodies by index: remember which hashid each old index had so the
	// cached records can be re-pointed after this re-upload (warm-start data survives structural edits)
	std::vector<uint32_t> old_hashid;
	if(w->steps > 0 && w->S.n > 0 && w->cap_arbs > 0){
		if(download(w, old_hashid, w->S.hashid, (size_t)w->S.n) || world_sync(w)) return -1;
	}
	std::vector<int> type(N), body(N), sensor(N), pcount(N), poff(N);
	std::vector<uint32_t> hashid(N), cat(N), mask(N), hlocal(N);
	std::vector<int> body_space;
	if(download(w, body_space, w->B.space, (size_t)w->B.n) || world_sync(w)) return -1;
	std::vector<uint32_t> space_base((size_t)w->n_spaces, 0xffffffffu);
	std::vector<uint64_t> group(N), ctype(N);
	std::vector<double> e(N), u(N), r(N);
	std::vector<V2> surfv(N), la(N), lb(N), ln(N), atan_(N), btan_(N), lpv(NV), lpn(NV);
	for(size_t i = 0; i < N; i++){
		const cpb200_shape_desc &d = shapes[i];
		if(d.body < 0 || d.body >= w->B.n){ cpb_set_error("shape %zu: body index %d out of range (upload bodies first)", i, d.body); return -1; }
		type[i] = d.type; body[i] = d.body; sensor[i] = d.sensor; hashid[i] = d.hashid; cat[i] = d.categories; mask[i] = d.mask;
		group[i] = d.group; ctype[i] = d.collision_type; e[i] = d.e; u[i] = d.u; r[i] = d.r;
		surfv[i] = v2(d.surface_v[0], d.surface_v[1]);
		la[i] = v2(d.a[0], d.a[1]); lb[i] = v2(d.b[0], d.b[1]);
		atan_[i] = v2(d.a_tangent[0], d.a_tangent[1]); btan_[i] = v2(d.b_tangent[0], d.b_tangent[1]);
		ln[i] = v2(0, 0); pcount[i] = 0; poff[i] = 0;
		if(d.type == CPB200_SHAPE_SEGMENT){
			// cpSegmentShapeInit: n = cpvrperp(cpvnormalize(cpvsub(b, a))) (cpShape.c:498)
			ln[i] = vrperp(vnormalize(vsub(lb[i], la[i])));
		} else if(d.type == CPB200_SHAPE_POLY){
			if(d.n_verts < 1 || d.vert_offset < 0 || d.vert_offset + d.n_verts > n_verts){ cpb_set_error("shape %zu: vertex range out of bounds", i); return -1; }
			if(d.n_verts > 255){ cpb_set_error("shape %zu: polygons are limited to 255 vertices (GJK vertex ids are 8 bit, cpCollision.c:129-134)", i); return -1; }
			pcount[i] = d.n_verts; poff[i] = d.vert_offset;
			// SetVerts (cpPolyShape.c:147-165): plane i holds vertex i and the normal of edge (i-1 -> i)
			for(int k = 0; k < d.n_verts; k++){
				const double *va = verts_xy + 2*(size_t)(d.vert_offset + (k - 1 + d.n_verts)%d.n_verts);
				const double *vb = verts_xy + 2*(size_t)(d.vert_offset + k);
				V2 a = v2(va[0], va[1]), b = v2(vb[0], vb[1]);
				lpv[(size_t)d.vert_offset + k] = b;
				lpn[(size_t)d.vert_offset + k] = vnormalize(vrperp(vsub(b, a)));
			}
		}
	}
	for(size_t i = 0; i < N; i++){ uint32_t &b = space_base[(size_t)body_space[(size_t)body[i]]]; b = std::min(b, hashid[i]); }
	for(size_t i = 0; i < N; i++) hlocal[i] = hashid[i] - space_base[(size_t)body_space[(size_t)body[i]]];
	std::vector<double4> mat(N); std::vector<uint2> ids(N);
	std::vector<double4> filt(N);
	for(size_t i = 0; i < N; i++){
		const unsigned long long x = (unsigned long long)(uint32_t)body[i] | ((unsigned long long)(uint32_t)type[i] << 32);
		const unsigned long long y = (unsigned long long)cat[i] | ((unsigned long long)mask[i] << 32), z = group[i];
		memcpy(&filt[i].x, &x, 8); memcpy(&filt[i].y, &y, 8); memcpy(&filt[i].z, &z, 8); filt[i].w = 0.0;
	}
	for(size_t i = 0; i < N; i++){ mat[i] = make_double4(e[i], u[i], surfv[i].x, surfv[i].y); ids[i].x = hashid[i]; ids[i].y = hlocal[i]; }
	w->gS.release();
	DShapes &S = w->S;
	S.n = n; S.nv = n_verts;
	const int cap = n + n/4 + 256, capv = n_verts + n_verts/4 + 1024;
	w->cap_shapes = cap; w->cap_verts = capv;
	w->space_base = space_base;
	DA(w->gS, S.type, cap); DA(w->gS, S.body, cap); DA(w->gS, S.hashid, cap); DA(w->gS, S.hlocal, cap); DA(w->gS, S.sensor, cap); DA(w->gS, S.cat, cap); DA(w->gS, S.mask, cap);
	DA(w->gS, S.group, cap); DA(w->gS, S.ctype, cap); DA(w->gS, S.e, cap); DA(w->gS, S.u, cap); DA(w->gS, S.r, cap); DA(w->gS, S.surfv, cap);
	DA(w->gS, S.la, cap); DA(w->gS, S.lb, cap); DA(w->gS, S.ln, cap); DA(w->gS, S.atan, cap); DA(w->gS, S.btan, cap);
	DA(w->gS, S.pcount, cap); DA(w->gS, S.poff, cap); DA(w->gS, S.lpv, capv); DA(w->gS, S.lpn, capv);
	DA(w->gS, S.mat, cap); DA(w->gS, S.circ, 2*(size_t)cap); DA(w->gS, S.ids, cap); DA(w->gS, S.filt, cap);
	DA(w->gS, S.wa, cap); DA(w->gS, S.wb, cap); DA(w->gS, S.wn, cap); DA(w->gS, S.wpv, capv); DA(w->gS, S.wpn, capv); DA(w->gS, S.bb, cap);
	if(upload(w, S.type, type) || upload(w, S.body, body) || upload(w, S.hashid, hashid) || upload(w, S.hlocal, hlocal) || upload(w, S.sensor, sensor) || upload(w, S.cat, cat) ||
	   upload(w, S.mask, mask) || upload(w, S.group, group) || upload(w, S.ctype, ctype) || upload(w, S.e, e) || upload(w, S.u, u) || upload(w, S.r, r) ||
	   upload(w, S.surfv, surfv) || upload(w, S.la, la) || upload(w, S.lb, lb) || upload(w, S.ln, ln) || upload(w, S.atan, atan_) || upload(w, S.btan, btan_) ||
	   upload(w, S.mat, mat) || upload(w, S.ids, ids) || upload(w, S.filt, filt) || upload(w, S.pcount, pcount) || upload(w, S.poff, poff) || upload(w, S.lpv, lpv) || upload(w, S.lpn, lpn)) return -1;

	w->shape_body = body; w->sl_dirty = true; w->bvh_valid = false;

	// broadphase scratch (sized for the capacity: appended shapes need no reallocation)
	w->gV.release();
	DBvh &T = w->bvh;
	T.n = n;
	int nn = cap;
	DA(w->gV, T.keys, nn); DA(w->gV, T.leaf_shape, nn); DA(w->gV, T.left, nn); DA(w->gV, T.right, nn); DA(w->gV, T.parent, 2*nn);
	DA(w->gV, T.nbb, 2*nn); DA(w->gV, T.nsp, 2*nn); DA(w->gV, T.flags, nn); DA(w->gV, T.bounds, 4); DA(w->gV, T.top_list, nn); DA(w->gV, T.top_count, 4);
	DA(w->gV, T.nskip, 2*nn); DA(w->gV, T.cbox, 2*nn); DA(w->gV, T.cinfo, nn); DA(w->gV, T.cspace, nn);
	DA(w->gV, w->keys_b, nn); DA(w->gV, w->vals_b, nn);
	DA(w->gV, w->sort_tmp, cpb_sort_tmp_elems(nn) + 16);

	int want_pairs = std::max(w->user_cap_pairs, 16*cap + 1024);
	int want_arbs = std::max(w->user_cap_arbs, 8*cap + 1024);
	if(want_pairs > w->cap_pairs){ if(alloc_pairs(w, want_pairs)) return -1; }
	if(want_arbs > w->cap_arbs){
		// growing drops the cached arbiters (warm-start data); callers that care reserve up front
		if(alloc_arbs(w, want_arbs)) return -1;
	} else if(!old_hashid.empty()){
		DArbs &A = w->A[w->cur];
		int n_rec = 0;
		CPB_CHECK(cudaMemcpyAsync(&n_rec, A.count_ptr, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
		if(world_sync(w)) return -1;
		n_rec = std::min(n_rec, A.cap);
		if(n_rec > 0){
			// new index of every live hashid: a sorted (hashid, index) list, searched per record (hashids come from a counter
			// that only grows, so a table indexed by hashid would cost as much as every shape the space ever had)
			std::vector<std::pair<uint32_t, int> > by_hash(N);
			for(size_t i = 0; i < N; i++) by_hash[i] = std::make_pair(hashid[i], (int)i);
			std::sort(by_hash.begin(), by_hash.end());
			auto find_new = [&](uint32_t h) -> int {
				auto it = std::lower_bound(by_hash.begin(), by_hash.end(), std::make_pair(h, -1));
				return (it != by_hash.end() && it->first == h ? it->second : -1);
			};
			std::vector<int> sa, sb, ba, bb; std::vector<uint64_t> key;
			if(download(w, sa, A.sa, (size_t)n_rec) || download(w, sb, A.sb, (size_t)n_rec) || download(w, key, A.key, (size_t)n_rec) || world_sync(w)) return -1;
			ba.resize((size_t)n_rec); bb.resize((size_t)n_rec);
			for(int i = 0; i < n_rec; i++){
				int na = -1, nb = -1;
				if(sa[i] >= 0 && (size_t)sa[i] < old_hashid.size()) na = find_new(old_hashid[sa[i]]);
				if(sb[i] >= 0 && (size_t)sb[i] < old_hashid.size()) nb = find_new(old_hashid[sb[i]]);
				if(na < 0 || nb < 0){ key[i] = ~0ull; sa[i] = sb[i] = 0; ba[i] = bb[i] = 0; } // a shape was removed: record dies in the next carry pass
				else { sa[i] = na; sb[i] = nb; ba[i] = body[(size_t)na]; bb[i] = body[(size_t)nb]; }
			}
			if(upload(w, A.sa, sa) || upload(w, A.sb, sb) || upload(w, A.ba, ba) || upload(w, A.bb, bb) || upload(w, A.key, key)) return -1;
		}
	}
	w->cache_dirty = true;
	w->hints_valid = false;
	return world_sync(w);
}

/* f4: shapes appended behind the existing ones.  vert_offset of the new polygons counts from the first appended vertex.
 * Returns 1 (nothing changed) when the slack of the shape / vertex arrays is used up or a new hashid lies below its
 * space's lowest one: the caller then re-uploads everything with cpb200_world_set_shapes. */
extern "C" int cpb200_world_append_shapes(cpb200_world *w, int n, const cpb200_shape_desc *shapes, int n_verts, const double *verts_xy)
{
	if(!w || n < 0 || n_verts < 0 || (n > 0 && !shapes)){ cpb_set_error("bad arguments"); return -1; }
	if(n == 0) return 0;
	if(w->mid_step || w->mid_solve){ cpb_set_error("structural edit inside a split step"); return -1; }
	DShapes &S = w->S;
	if(S.n + n > w->cap_shapes || S.nv + n_verts > w->cap_verts || (int)w->shape_body.size() != S.n || (int)w->body_space.size() != w->B.n ||
	   (int)w->space_base.size() != w->n_spaces) return 1;
	cudaSetDevice(w->device);
	const size_t N = (size_t)n, NV = (size_t)n_verts;
	const int s0 = S.n, v0 = S.nv;
	std::vector<int> type(N), body(N), sensor(N), pcount(N), poff(N);
	std::vector<uint32_t> hashid(N), cat(N), mask(N), hlocal(N);
	std::vector<uint64_t> group(N), ctype(N);
	std::vector<double> e(N), u(N), r(N);
	std::vector<V2> surfv(N), la(N), lb(N), ln(N), atan_(N), btan_(N), lpv(NV), lpn(NV);
	std::vector<double4> mat(N), filt(N); std::vector<uint2> ids(N);
	std::vector<uint32_t> base = w->space_base;
	for(size_t i = 0; i < N; i++){
		const cpb200_shape_desc &d = shapes[i];
		if(d.body < 0 || d.body >= w->B.n){ cpb_set_error("appended shape %zu: body index %d out of range (append bodies first)", i, d.body); return -1; }
		type[i] = d.type; body[i] = d.body; sensor[i] = d.sensor; hashid[i] = d.hashid; cat[i] = d.categories; mask[i] = d.mask;
		group[i] = d.group; ctype[i] = d.collision_type; e[i] = d.e; u[i] = d.u; r[i] = d.r;
		surfv[i] = v2(d.surface_v[0], d.surface_v[1]);
		la[i] = v2(d.a[0], d.a[1]); lb[i] = v2(d.b[0], d.b[1]);
		atan_[i] = v2(d.a_tangent[0], d.a_tangent[1]); btan_[i] = v2(d.b_tangent[0], d.b_tangent[1]);
		ln[i] = v2(0, 0); pcount[i] = 0; poff[i] = 0;
		if(d.type == CPB200_SHAPE_SEGMENT) ln[i] = vrperp(vnormalize(vsub(lb[i], la[i])));          // cpShape.c:498
		else if(d.type == CPB200_SHAPE_POLY){
			if(d.n_verts < 1 || d.vert_offset < 0 || d.vert_offset + d.n_verts > n_verts){ cpb_set_error("appended shape %zu: vertex range out of bounds", i); return -1; }
			if(d.n_verts > 255){ cpb_set_error("appended shape %zu: polygons are limited to 255 vertices", i); return -1; }
			pcount[i] = d.n_verts; poff[i] = v0 + d.vert_offset;
			for(int k = 0; k < d.n_verts; k++){                                                        // cpPolyShape.c:147-165
				const double *va = verts_xy + 2*(size_t)(d.vert_offset + (k - 1 + d.n_verts)%d.n_verts);
				const double *vb = verts_xy + 2*(size_t)(d.vert_offset + k);
				V2 a = v2(va[0], va[1]), b = v2(vb[0], vb[1]);
				lpv[(size_t)d.vert_offset + k] = b;
				lpn[(size_t)d.vert_offset + k] = vnormalize(vrperp(vsub(b, a)));
			}
		}
		const int sp = w->body_space[(size_t)d.body];
		if(sp < 0 || sp >= w->n_spaces) return 1;
		uint32_t &sb = base[(size_t)sp];
		if(sb == 0xffffffffu) sb = hashid[i];
		if(hashid[i] < sb) return 1;          // would shift every hlocal of the space
		hlocal[i] = hashid[i] - sb;
		const unsigned long long x = (unsigned long long)(uint32_t)body[i] | ((unsigned long long)(uint32_t)type[i] << 32);
		const unsigned long long y = (unsigned long long)cat[i] | ((unsigned long long)mask[i] << 32), z = group[i];
		memcpy(&filt[i].x, &x, 8); memcpy(&filt[i].y, &y, 8); memcpy(&filt[i].z, &z, 8); filt[i].w = 0.0;
		mat[i] = make_double4(e[i], u[i], surfv[i].x, surfv[i].y); ids[i].x = hashid[i]; ids[i].y = hlocal[i];
	}
	if(world_sync(w)) return -1;
	if(upload_nowait(w, S.type + s0, type) || upload_nowait(w, S.body + s0, body) || upload_nowait(w, S.hashid + s0, hashid) || upload_nowait(w, S.hlocal + s0, hlocal) || upload_nowait(w, S.sensor + s0, sensor) ||
	   upload_nowait(w, S.cat + s0, cat) || upload_nowait(w, S.mask + s0, mask) || upload_nowait(w, S.group + s0, group) || upload_nowait(w, S.ctype + s0, ctype) || upload_nowait(w, S.e + s0, e) || upload_nowait(w, S.u + s0, u) ||
	   upload_nowait(w, S.r + s0, r) || upload_nowait(w, S.surfv + s0, surfv) || upload_nowait(w, S.la + s0, la) || upload_nowait(w, S.lb + s0, lb) || upload_nowait(w, S.ln + s0, ln) || upload_nowait(w, S.atan + s0, atan_) ||
	   upload_nowait(w, S.btan + s0, btan_) || upload_nowait(w, S.mat + s0, mat) || upload_nowait(w, S.ids + s0, ids) || upload_nowait(w, S.filt + s0, filt) || upload_nowait(w, S.pcount + s0, pcount) ||
	   upload_nowait(w, S.poff + s0, poff) || upload_nowait(w, S.lpv + v0, lpv) || upload_nowait(w, S.lpn + v0, lpn)) return -1;
	if(world_sync(w)) return -1;       // the vectors above are locals
	w->space_base = base;
	S.n += n; S.nv += n_verts;
	w->bvh.n = S.n;
	w->shape_body.insert(w->shape_body.end(), body.begin(), body.end());
	w->sl_dirty = true; w->bvh_valid = false;
	w->cache_dirty = true;      // the world cache of the new shapes (and nothing else changes: same bodies, same transforms)
	return 0;
}

// the device set of body pairs joined by a constraint with collideBodies == 0 (QueryRejectConstraint, cpSpaceStep.c:204-217)
static int upload_nocollide(cpb200_world *w, const std::vector<uint64_t> &nocollide)
{
	if(w->d_nocollide){ cudaFree(w->d_nocollide); w->d_nocollide = NULL; }
	w->n_nocollide = 0;   // = capacity - 1 of the device set (0: no pairs)
	w->nocollide_keys = nocollide;
	if(nocollide.empty()) return 0;
	size_t capn = 16;
	while(capn < 2*nocollide.size()) capn <<= 1;
	std::vector<uint64_t> set(capn, 0ull);
	for(uint64_t k : nocollide){
		uint64_t key = k + 1ull;
		size_t slot = (size_t)((uint32_t)mix64(key) & (uint32_t)(capn - 1));
		while(set[slot] != 0ull) slot = (slot + 1) & (capn - 1);
		set[slot] = key;
	}
	void *p = NULL;
	CPB_CHECK(cudaMalloc(&p, sizeof(uint64_t)*capn));
	w->d_nocollide = (uint64_t *)p;
	w->n_nocollide = (int)(capn - 1);
	CPB_CHECK(cudaMemcpyAsync(w->d_nocollide, set.data(), sizeof(uint64_t)*capn, cudaMemcpyHostToDevice, w->stream));
	CPB_CHECK(cudaStreamSynchronize(w->stream));   // `set` is a local
	return 0;
}

extern "C" int cpb200_world_set_joints(cpb200_world *w, int n, const cpb200_joint_desc *joints)
{
	if(!w || n < 0){ cpb_set_error("bad arguments"); return -1; }
	cudaSetDevice(w->device);
	if(world_sync(w)) return -1;
	size_t N = (size_t)n;
	std::vector<int> type(N), a(N), b(N);
	std::vector<double> max_force(N), max_bias(N), aux0(N, 0.0);
	std::vector<V2> anchor_a(N), anchor_b(N), acc(N);
	std::vector<double4> prm(N);
	std::vector<uint64_t> nocollide, jpri(N), jnc(N);
	std::vector<int> body_space;
	if(download(w, body_space, w->B.space, (size_t)w->B.n) || world_sync(w)) return -1;
	std::vector<int> joint_base((size_t)w->n_spaces, -1);
	w->joint_error_bias.resize(N);
	for(size_t i = 0; i < N; i++){
		const cpb200_joint_desc &d = joints[i];
		if(d.a < 0 || d.a >= w->B.n || d.b < 0 || d.b >= w->B.n){ cpb_set_error("joint %zu: body index out of range", i); return -1; }
		type[i] = d.type; a[i] = d.a; b[i] = d.b;
		max_force[i] = d.max_force; max_bias[i] = d.max_bias; w->joint_error_bias[i] = d.error_bias;
		anchor_a[i] = v2(d.anchor_a[0], d.anchor_a[1]); anchor_b[i] = v2(d.anchor_b[0], d.anchor_b[1]);
		prm[i] = make_double4(d.prm[0], d.prm[1], d.prm[2], d.prm[3]);
		acc[i] = v2(d.acc[0], d.acc[1]);
		{
			int &jb = joint_base[(size_t)body_space[(size_t)d.a]];
			if(jb < 0) jb = (int)i;
			jpri[i] = mix64(0x9e3779b97f4a7c15ull ^ (uint64_t)((int)i - jb)) >> 8;
		}
		if(d.type == CPB200_JOINT_RATCHET) aux0[i] = d.prm[0];
		if(d.type == CPB200_JOINT_GROOVE){
			// cpGrooveJointInit: grv_n = cpvperp(cpvnormalize(cpvsub(groove_b, groove_a))) (cpGrooveJoint.c:128)
			V2 gn = vperp(vnormalize(vsub(v2(d.prm[0], d.prm[1]), anchor_a[i])));
			prm[i].z = gn.x; prm[i].w = gn.y;
		}
		jnc[i] = 0;
		if(!d.collide_bodies){
			uint64_t lo = (uint64_t)(uint32_t)std::min(d.a, d.b), hi = (uint64_t)(uint32_t)std::max(d.a, d.b);
			nocollide.push_back((lo << 32) | hi);
			jnc[i] = ((lo << 32) | hi) + 1ull;
		}
	}
	w->joint_nocollide = jnc;
	std::sort(nocollide.begin(), nocollide.end());
	nocollide.erase(std::unique(nocollide.begin(), nocollide.end()), nocollide.end());
	w->gJ.release();
	DJoints &J = w->J;
	J.n = n;
	const int cap = n + n/4 + 256;
	w->cap_joints = cap;
	w->joint_base = joint_base;
	DA(w->gJ, J.type, cap); DA(w->gJ, J.a, cap); DA(w->gJ, J.b, cap); DA(w->gJ, J.max_force, cap); DA(w->gJ, J.max_bias, cap); DA(w->gJ, J.bias_coef, cap);
	DA(w->gJ, J.anchor_a, cap); DA(w->gJ, J.anchor_b, cap); DA(w->gJ, J.prm, cap);
	DA(w->gJ, J.r1, cap); DA(w->gJ, J.r2, cap); DA(w->gJ, J.nrm, cap); DA(w->gJ, J.nmass, cap); DA(w->gJ, J.k, cap); DA(w->gJ, J.bias, cap); DA(w->gJ, J.acc, cap);
	DA(w->gJ, J.aux0, cap); DA(w->gJ, J.aux1, cap); DA(w->gJ, J.jspring, cap); DA(w->gJ, J.colour, cap); DA(w->gJ, J.row, cap); DA(w->gJ, J.pri, cap); DA(w->gJ, J.hint, cap);
	if(upload(w, J.type, type) || upload(w, J.a, a) || upload(w, J.b, b) || upload(w, J.max_force, max_force) || upload(w, J.max_bias, max_bias) ||
	   upload(w, J.anchor_a, anchor_a) || upload(w, J.anchor_b, anchor_b) || upload(w, J.prm, prm) || upload(w, J.acc, acc) || upload(w, J.aux0, aux0) || upload(w, J.pri, jpri)) return -1;
	if(upload_nocollide(w, nocollide)) return -1;
	w->joint_body = a; w->sl_dirty = true; w->bvh_valid = false;
	w->joints_dt = 0.0; // force bias_coef refresh
	w->hints_valid = false;
	return world_sync(w);
}

/* f4: joints appended behind the existing ones (their accumulated impulses start from joints[i].acc).  Returns 1 when
 * the slack is used up: the caller then re-uploads with cpb200_world_set_joints. */
extern "C" int cpb200_world_append_joints(cpb200_world *w, int n, const cpb200_joint_desc *joints)
{
	if(!w || n < 0 || (n > 0 && !joints)){ cpb_set_error("bad arguments"); return -1; }
	if(n == 0) return 0;
	if(w->mid_step || w->mid_solve){ cpb_set_error("structural edit inside a split step"); return -1; }
	DJoints &J = w->J;
	if(J.n + n > w->cap_joints || (int)w->joint_base.size() != w->n_spaces || (int)w->body_space.size() != w->B.n || w->joint_error_bias.size() != (size_t)J.n || w->joint_nocollide.size() != (size_t)J.n) return 1;
	cudaSetDevice(w->device);
	const size_t N = (size_t)n;
	const int j0 = J.n;
	std::vector<int> type(N), a(N), b(N), colour(N, -1), hint(N, -1);
	std::vector<double> max_force(N), max_bias(N), aux0(N, 0.0);
	std::vector<V2> anchor_a(N), anchor_b(N), acc(N);
	std::vector<double4> prm(N);
	std::vector<uint64_t> jpri(N), nocollide = w->nocollide_keys, jnc_new;
	std::vector<int> base = w->joint_base;
	bool new_nocollide = false;
	for(size_t i = 0; i < N; i++){
		const cpb200_joint_desc &d = joints[i];
		if(d.a < 0 || d.a >= w->B.n || d.b < 0 || d.b >= w->B.n){ cpb_set_error("appended joint %zu: body index out of range", i); return -1; }
		type[i] = d.type; a[i] = d.a; b[i] = d.b;
		max_force[i] = d.max_force; max_bias[i] = d.max_bias;
		anchor_a[i] = v2(d.anchor_a[0], d.anchor_a[1]); anchor_b[i] = v2(d.anchor_b[0], d.anchor_b[1]);
		prm[i] = make_double4(d.prm[0], d.prm[1], d.prm[2], d.prm[3]);
		acc[i] = v2(d.acc[0], d.acc[1]);
		int &jb = base[(size_t)w->body_space[(size_t)d.a]];
		if(jb < 0) jb = j0 + (int)i;
		jpri[i] = mix64(0x9e3779b97f4a7c15ull ^ (uint64_t)(j0 + (int)i - jb)) >> 8;
		if(d.type == CPB200_JOINT_RATCHET) aux0[i] = d.prm[0];
		if(d.type == CPB200_JOINT_GROOVE){
			V2 gn = vperp(vnormalize(vsub(v2(d.prm[0], d.prm[1]), anchor_a[i])));          // cpGrooveJoint.c:128
			prm[i].z = gn.x; prm[i].w = gn.y;
		}
		jnc_new.push_back(0);
		if(!d.collide_bodies){
			uint64_t lo = (uint64_t)(uint32_t)std::min(d.a, d.b), hi = (uint64_t)(uint32_t)std::max(d.a, d.b);
			nocollide.push_back((lo << 32) | hi); new_nocollide = true;
			jnc_new.back() = ((lo << 32) | hi) + 1ull;
		}
	}
	if(world_sync(w)) return -1;
	// (colour / hint = -1: a new joint has no colour to keep; the colouring puts it on its worklist)
	if(upload_nowait(w, J.type + j0, type) || upload_nowait(w, J.a + j0, a) || upload_nowait(w, J.b + j0, b) || upload_nowait(w, J.max_force + j0, max_force) || upload_nowait(w, J.max_bias + j0, max_bias) ||
	   upload_nowait(w, J.anchor_a + j0, anchor_a) || upload_nowait(w, J.anchor_b + j0, anchor_b) || upload_nowait(w, J.prm + j0, prm) || upload_nowait(w, J.acc + j0, acc) || upload_nowait(w, J.aux0 + j0, aux0) ||
	   upload_nowait(w, J.pri + j0, jpri) || upload_nowait(w, J.colour + j0, colour) || upload_nowait(w, J.hint + j0, hint)) return -1;
	if(world_sync(w)) return -1;       // the vectors above are locals
	if(new_nocollide){
		std::sort(nocollide.begin(), nocollide.end());
		nocollide.erase(std::unique(nocollide.begin(), nocollide.end()), nocollide.end());
		if(upload_nocollide(w, nocollide)) return -1;
	}
	for(size_t i = 0; i < N; i++) w->joint_error_bias.push_back(joints[i].error_bias);
	w->joint_nocollide.insert(w->joint_nocollide.end(), jnc_new.begin(), jnc_new.end());
	w->joint_base = base;
	J.n += n;
	w->joint_body.insert(w->joint_body.end(), a.begin(), a.end());
	w->sl_dirty = true; w->bvh_valid = false;
	w->joints_dt = 0.0;   // bias coefficients of all joints are refreshed at the next step (one small upload)
	return world_sync(w);
}

// ---- f4: removal in place.  The last object of the class moves into the hole (the host registries do the same swap, so
// host slot == device index stays true); everything that names the moved object by index is re-pointed on the device.
__global__ void k_move_shape(DShapes S, int dst, int src)
{
	if(CPB_TID != 0) return;
	#define MV(a) S.a[dst] = S.a[src]
	MV(type); MV(body); MV(hashid); MV(hlocal); MV(sensor); MV(cat); MV(mask); MV(group); MV(ctype); MV(e); MV(u); MV(r); MV(surfv);
	MV(la); MV(lb); MV(ln); MV(atan); MV(btan); MV(pcount); MV(poff); MV(wa); MV(wb); MV(wn); MV(bb); MV(mat); MV(ids); MV(filt);
	#undef MV
	S.circ[2*(size_t)dst] = S.circ[2*(size_t)src]; S.circ[2*(size_t)dst + 1] = S.circ[2*(size_t)src + 1];
}

// records of the removed shape die (their key can never be looked up again: hashids are not reused), records of the
// moved shape follow it
__global__ void k_arb_remap_shape(DArbs A, int removed, int moved)
{
	int n = *A.count_ptr; if(n > A.cap) n = A.cap;
	for(int i = CPB_TID; i < n; i += CPB_NTHREADS){
		int sa = A.sa[i], sb = A.sb[i];
		if(sa == removed || sb == removed){ A.key[i] = ~0ull; A.active[i] = 0; A.sa[i] = 0; A.sb[i] = 0; A.ba[i] = 0; A.bb[i] = 0; continue; }
		if(sa == moved) A.sa[i] = removed;
		if(sb == moved) A.sb[i] = removed;
	}
}

extern "C" int cpb200_world_remove_shape(cpb200_world *w, int index)
{
	if(!w || index < 0 || index >= w->S.n){ cpb_set_error("shape index out of range"); return -1; }
	if(w->mid_step || w->mid_solve){ cpb_set_error("structural edit inside a split step"); return -1; }
	if((int)w->shape_body.size() != w->S.n) return 1;
	cudaSetDevice(w->device);
	const int last = w->S.n - 1;
	if(index != last) LAUNCH(k_move_shape, 1, 32, w->stream, w->S, index, last);
	if(w->cap_arbs > 0) LAUNCH(k_arb_remap_shape, std::min(grid_for(w->A[w->cur].cap, 256), w->sm_count*8), 256, w->stream, w->A[w->cur], index, (index != last ? last : -1));
	w->shape_body[(size_t)index] = w->shape_body[(size_t)last];
	w->shape_body.pop_back();
	w->S.n = last; w->bvh.n = last;
	w->sl_dirty = true; w->bvh_valid = false;
	return world_sync(w);
}

__global__ void k_move_joint(DJoints J, int dst, int src, unsigned long long pri)
{
	if(CPB_TID != 0) return;
	#define MV(a) J.a[dst] = J.a[src]
	J.type[dst] = J.type[src]; J.a[dst] = J.a[src]; J.b[dst] = J.b[src];
	MV(max_force); MV(max_bias); MV(bias_coef); MV(anchor_a); MV(anchor_b); MV(prm); MV(r1); MV(r2); MV(nrm); MV(nmass); MV(k); MV(bias); MV(acc);
	MV(aux0); MV(aux1); MV(jspring); MV(colour); MV(hint);
	#undef MV
	J.pri[dst] = pri;     // priorities must stay unique among the live joints: the one of its new index
}

extern "C" int cpb200_world_remove_joint(cpb200_world *w, int index)
{
	if(!w || index < 0 || index >= w->J.n){ cpb_set_error("joint index out of range"); return -1; }
	if(w->mid_step || w->mid_solve){ cpb_set_error("structural edit inside a split step"); return -1; }
	if((int)w->joint_body.size() != w->J.n || w->joint_error_bias.size() != (size_t)w->J.n || w->joint_nocollide.size() != (size_t)w->J.n ||
	   (int)w->joint_base.size() != w->n_spaces || (int)w->body_space.size() != w->B.n) return 1;
	cudaSetDevice(w->device);
	const int last = w->J.n - 1;
	if(index != last){
		int jb = w->joint_base[(size_t)w->body_space[(size_t)w->joint_body[(size_t)last]]];
		if(jb < 0) jb = 0;
		LAUNCH(k_move_joint, 1, 32, w->stream, w->J, index, last, (unsigned long long)(mix64(0x9e3779b97f4a7c15ull ^ (uint64_t)(index - jb)) >> 8));
	}
	const bool had_nocollide = (w->joint_nocollide[(size_t)index] != 0);
	w->joint_body[(size_t)index] = w->joint_body[(size_t)last]; w->joint_body.pop_back();
	w->joint_error_bias[(size_t)index] = w->joint_error_bias[(size_t)last]; w->joint_error_bias.pop_back();
	w->joint_nocollide[(size_t)index] = w->joint_nocollide[(size_t)last]; w->joint_nocollide.pop_back();
	w->J.n = last;
	if(had_nocollide){
		std::vector<uint64_t> keys;
		for(uint64_t k : w->joint_nocollide) if(k) keys.push_back(k - 1);
		std::sort(keys.begin(), keys.end());
		keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
		if(upload_nocollide(w, keys)) return -1;
	}
	w->sl_dirty = true; w->bvh_valid = false;
	// (last step's colours of the remaining constraints stay valid: removing an edge cannot create a conflict)
	return world_sync(w);
}

__global__ void k_move_body(DBodies B, int dst, int src)
{
	if(CPB_TID != 0) return;
	#define MV(a) B.a[dst] = B.a[src]
	MV(pos); MV(ang); MV(rot); MV(txy); MV(cog); MV(V); MV(VB); MV(MI); MV(M); MV(force); MV(torque); MV(idle); MV(type); MV(space); MV(sleeping); MV(sgroup); MV(custom);
	#undef MV
}

// every index that named body `moved` now names `dst`; sleeping groups are named by their root body as well
__global__ void k_remap_body(DBodies B, DShapes S, DJoints J, DArbs A, int dst, int moved)
{
	const int tid = CPB_TID, nth = CPB_NTHREADS;
	for(int i = tid; i < B.n; i += nth){ if(B.sgroup[i] == moved) B.sgroup[i] = dst; }
	for(int s = tid; s < S.n; s += nth){
		if(S.body[s] != moved) continue;
		S.body[s] = dst;
		unsigned long long x = (unsigned long long)__double_as_longlong(S.filt[s].x);
		x = (x & 0xffffffff00000000ull) | (unsigned long long)(uint32_t)dst;
		S.filt[s].x = __longlong_as_double((long long)x);
	}
	for(int j = tid; j < J.n; j += nth){ if(J.a[j] == moved) J.a[j] = dst; if(J.b[j] == moved) J.b[j] = dst; }
	int n = *A.count_ptr; if(n > A.cap) n = A.cap;
	for(int i = tid; i < n; i += nth){ if(A.ba[i] == moved) A.ba[i] = dst; if(A.bb[i] == moved) A.bb[i] = dst; }
}

/* The body must not be named by any shape or joint any more (remove those first); returns 1 otherwise, or when the
 * engine's host-side bookkeeping cannot follow (the caller then re-uploads). */
extern "C" int cpb200_world_remove_body(cpb200_world *w, int index)
{
	if(!w || index < 0 || index >= w->B.n){ cpb_set_error("body index out of range"); return -1; }
	if(w->mid_step || w->mid_solve){ cpb_set_error("structural edit inside a split step"); return -1; }
	if((int)w->body_space.size() != w->B.n || (int)w->shape_body.size() != w->S.n || (int)w->joint_body.size() != w->J.n) return 1;
	for(int b : w->shape_body) if(b == index) return 1;
	for(int b : w->joint_body) if(b == index) return 1;
	cudaSetDevice(w->device);
	std::vector<int> jb;
	if(w->J.n){ if(download(w, jb, w->J.b, (size_t)w->J.n) || world_sync(w)) return -1; for(int b : jb) if(b == index) return 1; }
	const int last = w->B.n - 1;
	if(index != last){
		LAUNCH(k_move_body, 1, 32, w->stream, w->B, index, last);
		w->B.n = last;      // (the remap below must not walk over the vacated slot)
		DArbs dummy = w->A[w->cur];
		if(w->cap_arbs == 0){ cpb_set_error("internal: arbiter buffers missing"); return -1; }
		LAUNCH(k_remap_body, std::min(grid_for(std::max(std::max(w->B.n, w->S.n), dummy.cap), 256), w->sm_count*8), 256, w->stream, w->B, w->S, w->J, dummy, index, last);
		for(int &b : w->shape_body) if(b == last) b = index;
		for(int &b : w->joint_body) if(b == last) b = index;
		w->body_space[(size_t)index] = w->body_space[(size_t)last];
	}
	w->B.n = last;
	w->body_space.pop_back();
	w->io_src = NULL; w->io_sink = NULL;
	w->sl_dirty = true; w->bvh_valid = false;
	w->cache_dirty = true;      // the packed circle lines carry the body index: rewritten for every shape by the next cache pass
	return world_sync(w);
}

static int refresh_joint_bias(cpb200_world *w, double dt)
{
	if(w->J.n == 0 || w->joints_dt == dt) return 0;
	std::vector<double> bc(w->joint_error_bias.size());
	double last_eb = NAN, last = 0.0;
	for(size_t i = 0; i < bc.size(); i++){
		double eb = w->joint_error_bias[i];
		if(!(eb == last_eb)){ last_eb = eb; last = 1.0 - pow(eb, dt); } // bias_coef (chipmunk_private.h:264-268)
		bc[i] = last;
	}
	if(upload(w, w->J.bias_coef, bc)) return -1;
	w->joints_dt = dt;
	return 0;
}

// ------------------------------------------------------------------ the step
#define STAGE_BEGIN(w) do { if((w)->profiling) cudaEventRecord((w)->ev[0], (w)->stream); } while(0)
#define STAGE_END(w, id) do { if((w)->profiling){ cudaEventRecord((w)->ev[(id) + 1], (w)->stream); } } while(0)

// start-of-step bookkeeping in one launch (the small per-step clears used to be five memsets)
__global__ void k_reset_step(DCounters *C, int *pair_count, int *cur_count, int *ccount, int *jcount, int *wl_n, unsigned *bar)
{
	if(CPB_TID == 0){ bar[0] = 0u; bar[32] = 0u; }   // arrival counters of the two cooperative solver kernels
	if(ccount){
		for(int k = CPB_TID; k <= CPB_MAX_COLOURS; k += CPB_NTHREADS){ ccount[k] = 0; jcount[k] = 0; }
		for(int k = CPB_TID; k < CPB_MAX_COLOUR_ROUNDS + 2; k += CPB_NTHREADS) wl_n[k] = 0;
	}
	if(CPB_TID != 0) return;
	C->stamp++;            // cpSpaceStep.c:349
	C->n_pairs[0] = C->n_pairs[1] = C->n_pairs[2] = 0;
	C->n_contacts = 0; C->n_active = 0; C->n_colours = 0; C->n_cached = 0; C->n_row_solves = 0; C->n_row_idle = 0; C->bvh_visits = 0; C->bvh_queries = 0;
	C->colour_remaining[0] = C->colour_remaining[1] = 0; C->colour_rounds = 0; C->n_overflow_colour = 0;
	pair_count[0] = pair_count[1] = pair_count[2] = pair_count[3] = 0;
	*cur_count = 0;
}

__global__ void k_finish_step(DCounters *C, const int *pair_count)
{
	if(CPB_TID != 0) return;
	C->n_pairs[0] = pair_count[0]; C->n_pairs[1] = pair_count[1]; C->n_pairs[2] = pair_count[2];
}

// map the user's (shape a, shape b) order list onto arbiter record indices
__global__ void k_build_order(DShapes S, DArbs A, DTable T, const uint64_t *__restrict__ user, int n_user, int *order, int *n_order_out)
{
	if(CPB_TID != 0) return;
	int nA = *A.count_ptr; if(nA > A.cap) nA = A.cap;
	for(int i = 0; i < nA; i++) A.seen[i] = 0;
	int n = 0;
	for(int q = 0; q < n_user; q++){
		int sa = (int)(user[q] >> 32), sb = (int)(user[q] & 0xffffffffu);
		if(sa < 0 || sa >= S.n || sb < 0 || sb >= S.n) continue;
		int idx = table_find(T, *T.dmask, arb_key(S.hashid[sa], S.hashid[sb]));
		if(idx >= 0 && A.active[idx] == 1 && !A.seen[idx]){
			// bit 30 = the caller's a/b orientation is the reverse of ours: contacts are then visited in
			// reverse, which is the order the reference's ContactPoints produced them in (cpCollision.c:477-518)
			order[n++] = idx | (A.sa[idx] != sa ? 0x40000000 : 0); A.seen[idx] = 1;
		}
	}
	for(int i = 0; i < nA; i++){
		if(A.active[i] == 1 && !A.seen[i]) order[n++] = i;
		A.seen[i] = 0;
	}
	*n_order_out = n;
}

// Space-local solver layout: every space must own one contiguous body range small enough for the shared
// memory of a CTA (two 32-byte velocity sectors per body).
static int sl_refresh(cpb200_world *w)
{
	w->sl_dirty = false; w->sl_ok = false; w->sl_shapes_ok = false;
	w->gSL.release(); memset(&w->SL, 0, sizeof(w->SL)); w->sl_tmp = NULL;
	const int ns = w->n_spaces, nb = w->B.n;
	if(w->sl_disabled || nb == 0 || (int)w->body_space.size() != nb) return 0;
	if((size_t)cpb_div_up(nb, ns)*64 > CPB_SL_MAX_SMEM) return 0;   // the average space is already too large for a CTA: skip the scan (1 M bodies)
	std::vector<int> first((size_t)ns, -1), count((size_t)ns, 0);
	for(int i = 0; i < nb; i++){
		int sp = w->body_space[(size_t)i];
		if(sp < 0 || sp >= ns) return 0;
		if(first[(size_t)sp] < 0) first[(size_t)sp] = i;
		else if(first[(size_t)sp] + count[(size_t)sp] != i) return 0;   // not contiguous
		count[(size_t)sp]++;
	}
	int mx = 0;
	for(int sp = 0; sp < ns; sp++){ if(first[(size_t)sp] < 0) first[(size_t)sp] = 0; mx = std::max(mx, count[(size_t)sp]); }
	if((size_t)mx*64 > CPB_SL_MAX_SMEM) return 0;
	int *d_first = NULL, *d_count = NULL; uint32_t *d_start = NULL;
	size_t nbuckets = 2*(size_t)ns*CPB_MAX_COLOURS + 2;
	DA(w->gSL, d_first, ns); DA(w->gSL, d_count, ns); DA(w->gSL, d_start, nbuckets);
	DA(w->gSL, w->sl_tmp, cpb_scan_tmp_elems((int)nbuckets) + 16);
	CPB_CHECK(cudaMemcpyAsync(d_first, first.data(), sizeof(int)*(size_t)ns, cudaMemcpyHostToDevice, w->stream));
	CPB_CHECK(cudaMemcpyAsync(d_count, count.data(), sizeof(int)*(size_t)ns, cudaMemcpyHostToDevice, w->stream));
	CPB_CHECK(cudaStreamSynchronize(w->stream));
	w->SL.n_spaces = ns; w->SL.body0 = d_first; w->SL.nbody = d_count; w->SL.start = d_start;
	w->sl_max_nbody = mx;
	{
		std::vector<char> jointed((size_t)ns, 0);
		for(int b : w->joint_body){ if(b >= 0 && b < nb) jointed[(size_t)w->body_space[(size_t)b]] = 1; }
		std::vector<int> plain, withj;
		for(int sp = 0; sp < ns; sp++) (jointed[(size_t)sp] ? withj : plain).push_back(sp);
		int *d_a = NULL, *d_b = NULL;
		DA(w->gSL, d_a, ns); DA(w->gSL, d_b, ns);
		if(!plain.empty()) CPB_CHECK(cudaMemcpyAsync(d_a, plain.data(), sizeof(int)*plain.size(), cudaMemcpyHostToDevice, w->stream));
		if(!withj.empty()) CPB_CHECK(cudaMemcpyAsync(d_b, withj.data(), sizeof(int)*withj.size(), cudaMemcpyHostToDevice, w->stream));
		CPB_CHECK(cudaStreamSynchronize(w->stream));
		w->d_sl_plain = d_a; w->n_sl_plain = (int)plain.size(); w->d_sl_jointed = d_b; w->n_sl_jointed = (int)withj.size();
	}
	w->sl_ok = true;
	// contiguous shape range per space (space-local broadphase)
	const int nsh = w->S.n;
	if((int)w->shape_body.size() == nsh && nsh > 0){
		std::vector<int> sfirst((size_t)ns, -1), scount((size_t)ns, 0);
		bool ok = true;
		for(int i = 0; i < nsh && ok; i++){
			int b = w->shape_body[(size_t)i];
			if(b < 0 || b >= nb){ ok = false; break; }
			int sp = w->body_space[(size_t)b];
			if(sfirst[(size_t)sp] < 0) sfirst[(size_t)sp] = i;
			else if(sfirst[(size_t)sp] + scount[(size_t)sp] != i) ok = false;
			scount[(size_t)sp]++;
		}
		if(ok){
			int smx = 0;
			for(int sp = 0; sp < ns; sp++){ if(sfirst[(size_t)sp] < 0) sfirst[(size_t)sp] = 0; smx = std::max(smx, scount[(size_t)sp]); }
			int *d_s0 = NULL, *d_sn = NULL;
			DA(w->gSL, d_s0, ns); DA(w->gSL, d_sn, ns);
			CPB_CHECK(cudaMemcpyAsync(d_s0, sfirst.data(), sizeof(int)*(size_t)ns, cudaMemcpyHostToDevice, w->stream));
			CPB_CHECK(cudaMemcpyAsync(d_sn, scount.data(), sizeof(int)*(size_t)ns, cudaMemcpyHostToDevice, w->stream));
			CPB_CHECK(cudaStreamSynchronize(w->stream));
			w->SS.shape0 = d_s0; w->SS.nshape = d_sn;
			w->sl_max_nshape = smx;
			w->sl_shapes_ok = true;
		}
	}
	return 0;
}

// The step in two halves.  Phase A: positions, shape cache, broadphase, narrowphase + arbiter update (everything
// up to the point where the reference calls the begin/preSolve collision handlers, cpSpaceStep.c:234-290).
// Phase B: islands, cache ageing, prestep, velocities, solver.  cpb200_world_step runs both back to back;
// the host layer of a space WITH collision handlers calls them separately and edits arbiters in between.
// does this step rebuild the LBVH (bounds, Morton keys, sort, hierarchy) or only refit the one it has?
static bool bvh_in_use(const cpb200_world *w)
{
#ifndef CPB_EMU
	if(w->sl_shapes_ok && w->sl_max_nshape <= CPB_SL_MAX_SHAPES && w->S.n >= 2) return false;
#endif
	return w->S.n >= 2;
}
static bool bvh_rebuild_due(const cpb200_world *w){ return !w->bvh_valid || w->bvh_age >= w->bvh_period; }

static int step_phase_a(cpb200_world *w, double dt)
{
	cudaSetDevice(w->device);
	cudaStream_t st = w->stream;
	DBodies &B = w->B; DShapes &S = w->S;
	if(S.n > 0 && w->cap_arbs == 0){ cpb_set_error("internal: arbiter buffers missing"); return -1; }
	if(w->cap_arbs == 0){ if(alloc_arbs(w, 1024) || alloc_pairs(w, 1024)) return -1; }

	w->stamp++;
	double prev_dt = w->curr_dt;
	w->curr_dt = dt;
	double dt_coef = (prev_dt == 0.0 ? 0.0 : dt/prev_dt); // cpSpaceStep.c:407
	if(refresh_spaces(w, dt) || refresh_joint_bias(w, dt)) return -1;

	int iterations = 0;
	for(int i = 0; i < w->n_spaces; i++) iterations = std::max(iterations, w->sp[(size_t)i].iterations);

	if(w->cache_dirty){
		if(S.n) LAUNCH(k_shape_cache, grid_for(S.n, 128), 128, st, S, B, 1);
		w->cache_dirty = false;
	}

	// swap arbiter buffers: last step's records become "prev"
	int prv = w->cur; w->cur ^= 1;
	DArbs &Ap = w->A[prv]; DArbs &Ac = w->A[w->cur];
	DTable &Tp = w->T[prv];
	w->arb_derived_stale[w->cur] = false;
	LAUNCH(k_reset_step, 1, 32, st, w->C, w->P.count, Ac.count_ptr, w->K.ccount, w->K.jcount, w->K.wl_n, w->d_barrier);

	const int nb = B.n, ns = S.n;
	const int wide = w->sm_count*8;

	STAGE_BEGIN(w);
	const bool io = ((w->io_src || w->io_sink) && nb > 0 && w->io_cap >= nb);
	if(io){
		// host I/O of this step on a side stream: the forces travel while the collision phase runs (K9 consumes them),
		// the new positions -- final after K1 -- travel while the rest of the step runs
		CPB_CHECK(cudaEventRecord(w->ev_io_begin, st)); CPB_CHECK(cudaStreamWaitEvent(w->stream_io, w->ev_io_begin, 0));
		if(w->io_src){
			CPB_CHECK(cudaMemcpyAsync(w->d_io_force, w->io_src, sizeof(double)*3*(size_t)nb, cudaMemcpyHostToDevice, w->stream_io));
			CPB_CHECK(cudaEventRecord(w->ev_io_forces, w->stream_io));
		}
	}
	if(nb) LAUNCH(k_integrate_pos, grid_for(nb, 256), 256, st, B, dt, w->K.claim, w->K.bmask);
	if(io && w->io_sink){
		LAUNCH(k_pack_pos, grid_for(nb, 256), 256, st, B, w->d_io_pos);
		CPB_CHECK(cudaEventRecord(w->ev_io_pos, st)); CPB_CHECK(cudaStreamWaitEvent(w->stream_io, w->ev_io_pos, 0));
		CPB_CHECK(cudaMemcpyAsync(w->io_sink, w->d_io_pos, sizeof(double)*3*(size_t)nb, cudaMemcpyDeviceToHost, w->stream_io));
	}
	STAGE_END(w, ST_INTEGRATE_POS);
	if(ns) LAUNCH(k_shape_cache, grid_for(ns, 128), 128, st, S, B, 0);
	STAGE_END(w, ST_SHAPE_CACHE);

	// The warm-start lines of last step's records (what the narrowphase looks up) depend on nothing this step does: the
	// pass that packs them runs on a second stream, three CTAs per SM, beside the tree build / refit, and joins in front
	// of the collide kernels.  Worth 2 % of the 1 M step together with the shared-memory refit.  Measured alternatives:
	// beside the integrators (bandwidth-bound themselves) it gains nothing; beside the old refit (a 20-level chain of L2
	// atomics) anything wider than one CTA per SM slowed that chain down by what it saved; beside the traversal it costs
	// more (0.24 -> 0.34 ms) than it saves.
#ifndef CPB_EMU
	CPB_CHECK(cudaEventRecord(w->ev_fork_a, st)); CPB_CHECK(cudaStreamWaitEvent(w->stream2, w->ev_fork_a, 0));
	LAUNCH(k_pack_warm, std::min(grid_for(Ap.cap, 256), w->sm_count*w->pack_ctas), 256, w->stream2, Ap);
	CPB_CHECK(cudaEventRecord(w->ev_join_a, w->stream2));
#endif


	// K3: space-local all-pairs broadphase for worlds of small spaces, else the LBVH
	if(w->sl_dirty && sl_refresh(w)) return -1;
#ifndef CPB_EMU
	const bool sl_broad = w->sl_shapes_ok && w->sl_max_nshape <= CPB_SL_MAX_SHAPES && ns >= 2;
#else
	const bool sl_broad = false;
#endif
	if(sl_broad){
#ifndef CPB_EMU
		STAGE_END(w, ST_BVH_KEYS); STAGE_END(w, ST_BVH_SORT); STAGE_END(w, ST_BVH_BUILD);
		int threads = 32; while(threads < 1024 && threads < w->sl_max_nshape) threads *= 2;
		if(w->n_spaces > 1 && threads > 256) threads = 256;   // many spaces: more resident CTAs beat wider ones
		const int cand_cap = std::max(2048, 4*w->sl_max_nshape);
		size_t smem = (size_t)w->sl_max_nshape*(sizeof(double4) + sizeof(int)) + sizeof(unsigned)*(size_t)cand_cap;
		LAUNCH_SMEM(k_sl_pairs, w->n_spaces, threads, smem, st, w->SS, S, B, w->P, (const uint64_t *)w->d_nocollide, w->n_nocollide, &w->C->overflow, w->sl_max_nshape, cand_cap);
		STAGE_END(w, ST_BVH_PAIRS);
#endif
	} else if(ns >= 2){
		DBvh &T = w->bvh;
		const bool rebuild = bvh_rebuild_due(w);
		if(rebuild){
			LAUNCH(k_bounds_init, 1, 32, st, T.bounds, T.top_count);
			LAUNCH(k_bounds, std::min(grid_for(ns, 256), wide), 256, st, S, T.bounds);
			// Morton bits: log2(shapes) + 4 (sixteen cells per shape), in whole radix digits
			int want_bits = 4; while((1 << (want_bits - 4)) < ns && want_bits < 32) want_bits++;
			want_bits = std::min(32, std::max(16, (want_bits + 7) & ~7));
			const int drop_bits = 32 - want_bits;
			LAUNCH(k_morton, grid_for(ns, 256), 256, st, S, B, (const double *)T.bounds, T.keys, T.leaf_shape, drop_bits, T.flags);
			STAGE_END(w, ST_BVH_KEYS);
			int space_bits = 0; while((1 << space_bits) < w->n_spaces) space_bits++;
			int bits = 32 + space_bits;
			int where = cpb_radix_sort(T.keys, T.leaf_shape, w->keys_b, w->vals_b, ns, bits, w->sort_tmp, st, drop_bits);
			if(where){
				if(w->bvh_period > 1){
					// the leaf order outlives this step (and the graphs that replay the steps in between name T.leaf_shape): bring it home
					CPB_CHECK(cudaMemcpyAsync(T.keys, w->keys_b, sizeof(uint64_t)*(size_t)ns, cudaMemcpyDeviceToDevice, st));
					CPB_CHECK(cudaMemcpyAsync(T.leaf_shape, w->vals_b, sizeof(int)*(size_t)ns, cudaMemcpyDeviceToDevice, st));
				} else { std::swap(T.keys, w->keys_b); std::swap(T.leaf_shape, w->vals_b); }
			}
			STAGE_END(w, ST_BVH_SORT);
			LAUNCH(k_bvh_build, grid_for(ns - 1, 256), 256, st, T);
			w->bvh_valid = true; w->bvh_age = 1;
		} else {
			STAGE_END(w, ST_BVH_KEYS); STAGE_END(w, ST_BVH_SORT);
			w->bvh_age++;
		}
#ifndef CPB_EMU
		if(!w->refit_unfused){
			LAUNCH(k_bvh_refit_fused, grid_for(ns, CPB_REFIT_WIN), CPB_REFIT_WIN, st, T, S, B);
			LAUNCH(k_bvh_refit_top, std::min(grid_for(ns/16 + 128, 128), wide), 128, st, T);
		} else
#endif
		{
			LAUNCH(k_bvh_leaves, grid_for(ns, 256), 256, st, T, S, B, 1);
			LAUNCH(k_bvh_refit, grid_for(ns, 256), 256, st, T);
			LAUNCH(k_bvh_pack, grid_for(ns - 1, 256), 256, st, T);
		}
		STAGE_END(w, ST_BVH_BUILD);
		LAUNCH(k_bvh_pairs, grid_for(ns, 128), 128, st, T, S, B, w->P, (const uint64_t *)w->d_nocollide, w->n_nocollide, (int)(w->n_spaces > 1), &w->C->overflow, &w->C->bvh_visits);
#ifndef CPB_EMU
		LAUNCH(k_pair_filter, std::min(grid_for(w->P.cap, 128*CPB_FILTER_ROUNDS), wide), 128, st, S, w->P, (const uint64_t *)w->d_nocollide, w->n_nocollide, &w->C->overflow);
#endif
		STAGE_END(w, ST_BVH_PAIRS);
	} else {
		STAGE_END(w, ST_BVH_KEYS); STAGE_END(w, ST_BVH_SORT); STAGE_END(w, ST_BVH_BUILD); STAGE_END(w, ST_BVH_PAIRS);
	}

	// K5 + K6
	{
		int g = std::min(grid_for(w->P.cap, 128), w->sm_count*CPB_COLLIDE_CTAS);
#ifndef CPB_EMU
		CPB_CHECK(cudaStreamWaitEvent(st, w->ev_join_a, 0));
#else
		LAUNCH(k_pack_warm, std::min(grid_for(Ap.cap, 256), wide), 256, st, Ap);
#endif
		LAUNCH(k_collide<0>, g, 128, st, S, B, (const int *)w->P.a[0], (const int *)w->P.b[0], (const int *)&w->P.count[0], w->P.cap, Ap, Tp, Ac, w->C);
		LAUNCH(k_collide<1>, g, 128, st, S, B, (const int *)w->P.a[1], (const int *)w->P.b[1], (const int *)&w->P.count[1], w->P.cap, Ap, Tp, Ac, w->C);
		LAUNCH(k_collide<2>, g, 128, st, S, B, (const int *)w->P.a[2], (const int *)w->P.b[2], (const int *)&w->P.count[2], w->P.cap, Ap, Tp, Ac, w->C);
	}
	STAGE_END(w, ST_COLLIDE);
	w->step_dt = dt; w->step_dt_coef = dt_coef; w->step_iterations = iterations;
	w->mid_step = true;
	return 0;
}

static int step_phase_b2(cpb200_world *w, bool fused);

// K9 (with the forces of a bound host buffer)
static int launch_integrate_vel(cpb200_world *w)
{
	cudaStream_t st = w->stream;
	const int nb = w->B.n;
	if(w->io_src && nb > 0 && w->io_cap >= nb){
		CPB_CHECK(cudaStreamWaitEvent(st, w->ev_io_forces, 0));
		LAUNCH(k_set_forces, grid_for(nb, 256), 256, st, w->B, (const double *)w->d_io_force, 0, nb);
	}
	if(nb) LAUNCH(k_integrate_vel, grid_for(nb, 256), 256, st, w->B, (const DSpace *)w->d_spaces, w->step_dt);
	return 0;
}

// Phase B up to the solver: islands, cache ageing + table, prestep, velocity integration.
// fused (a production step that nothing interrupts between the collision phase and the solver): the arbiters' prestep
// is folded into the solver's row build and the velocity integration follows it -- cpArbiterPreStep must see the
// velocities from before cpBodyUpdateVelocity (bounce, cpArbiter.c:436) -- see step_phase_b2.
static int step_phase_b1(cpb200_world *w, bool fused)
{
	cudaSetDevice(w->device);
	cudaStream_t st = w->stream;
	DBodies &B = w->B; DShapes &S = w->S; DJoints &J = w->J;
	(void)S; (void)J;
	double dt = w->step_dt;
	const int prv = w->cur ^ 1;
	DArbs &Ap = w->A[prv]; DArbs &Ac = w->A[w->cur];
	DTable &Tc = w->T[w->cur];
	const int wide = w->sm_count*8;
	w->mid_step = false;

	// K7: islands / sleeping (cpSpaceProcessComponents) -- before the cache filter, like the reference
	if(w->any_sleep_enabled){
		if(islands_step(w->I, B, S, J, Ap, Ac, Tc, w->d_spaces, dt, w->cur & 1, w->C, w->sm_count, st)) return -1;
	}
	STAGE_END(w, ST_ISLANDS);

	{
		int g = std::min(grid_for(Ap.cap, CPB_CARRY_BLOCK), w->sm_count*CPB_CARRY_CTAS);
		LAUNCH(k_arb_carry, g, CPB_CARRY_BLOCK, st, B, Ap, Ac, (const DSpace *)w->d_spaces, w->C);
		// the table of this step's records (next step's warm-start lookups), sized to what the step produced
#ifndef CPB_EMU
		if(Ac.cap <= 65536) LAUNCH(k_table_small, 1, 1024, st, Ac, Tc, w->C);
		else
#endif
		{
			LAUNCH(k_table_clear, std::min(grid_for(Ac.cap, 128), wide), 256, st, Ac, Tc);
			LAUNCH(k_table_build, std::min(grid_for(Ac.cap, 256), wide*2), 256, st, Ac, Tc, w->C);
		}
	}
	STAGE_END(w, ST_CARRY);

	// K8
	{
		int g = std::min(grid_for(Ac.cap, 128), wide);
		if(!fused) LAUNCH(k_arb_prestep, g, 128, st, B, Ac, (const DSpace *)w->d_spaces, dt, 0);
		if(J.n) LAUNCH(k_joint_prestep, grid_for(J.n, 128), 128, st, J, B, dt);
	}
	w->arb_derived_stale[w->cur] = fused;
	STAGE_END(w, ST_PRESTEP);
	if(!fused){
		// K9 (stage order of the profile: colour_rows closes empty here, the colouring is timed with the solver)
		STAGE_END(w, ST_COLOUR);
		if(launch_integrate_vel(w)) return -1;
		STAGE_END(w, ST_INTEGRATE_VEL);
	}
	w->mid_solve = true;
	return 0;
}

static int step_phase_b(cpb200_world *w, bool fused)
{
#ifdef CPB_EMU
	fused = false;
#endif
	if(w->solver_mode == 1) fused = false;   // the serial validation solver reads the records
	if(step_phase_b1(w, fused)) return -1;
	return step_phase_b2(w, fused);
}

// How the coloured solver is launched this step: grid of the persistent kernel, kernel family, CTA width of the
// space-local solver.  Depends on host-side estimates only, so the step-graph signature can include it.
struct SolvePlan { int blocks, iter_blocks; bool space_local, stream_rows; int threads; };
static SolvePlan plan_solver(cpb200_world *w)
{
	SolvePlan p;
	const int nb = w->B.n;
	// size the persistent grid to the work: ~256 rows of one colour per CTA, never more than
	// what is co-resident (148 SMs x 2 CTAs of 256 threads)
	int est_cons = std::max(w->last_active, nb) + w->J.n;
	// (rounded up to three significant bits: the constraint count moves a little every step, and a plan that followed it
	// exactly would change the step graph's signature every time the host reads the counters back)
	{ int sh = 0; while((est_cons >> sh) > 15) sh++; est_cons = (((est_cons >> sh) + (sh ? 1 : 0)) << sh); }
	p.blocks = std::max(1, std::min(w->coop_blocks, cpb_div_up(est_cons + 1, 256)));
	if(est_cons <= 4096) p.blocks = 1;   // small scenes: one CTA, colours separated by __syncthreads only
	p.iter_blocks = std::max(1, std::min(w->sm_count*(w->solve_minb == 3 ? 3 : 2), cpb_div_up(est_cons + 1, 256)));
	if(est_cons <= 4096) p.iter_blocks = 1;
	if(w->force_blocks > 0){ p.blocks = std::min(w->coop_blocks, w->force_blocks); p.iter_blocks = std::min(w->sm_count*2, w->force_blocks); }
	// Space-local path: worlds whose spaces are each a small contiguous body range (batched demo
	// spaces, or one small scene).  A forced grid (validation hook) keeps the world-wide solver.
	p.space_local = w->sl_ok && w->force_blocks == 0 && est_cons/w->n_spaces <= 4096;
	if(w->solver_variant == 1 || w->solver_variant == 2) p.space_local = false;
	if(w->solver_variant == 3) p.space_local = true;
	// rows + velocity sectors of one pass: stream the rows past the L2 only if they would not fit next to the velocities
	p.stream_rows = ((size_t)est_cons*(size_t)CPB_ROW_BYTES_EST + (size_t)nb*64 > (size_t)CPB_L2_RESIDENT_BYTES);
	if(w->solver_variant == 1) p.stream_rows = false;
	if(w->solver_variant == 2) p.stream_rows = true;
	// CTA width of k_sl_solve: about one row per thread in an average colour, a warp at least
	int per_space = est_cons/w->n_spaces;
	p.threads = 32; while(p.threads < 256 && p.threads*8 < per_space) p.threads *= 2;
	return p;
}

// K10 + K11 and the end of the step
static int step_phase_b2(cpb200_world *w, bool fused)
{
	cudaSetDevice(w->device);
	cudaStream_t st = w->stream;
	DBodies &B = w->B; DShapes &S = w->S; DJoints &J = w->J;
	double dt = w->step_dt, dt_coef = w->step_dt_coef;
	int iterations = w->step_iterations;
	DArbs &Ac = w->A[w->cur];
	DTable &Tc = w->T[w->cur];
	const int nb = B.n;
	const int wide = w->sm_count*8;
	w->mid_solve = false;

	if(w->solver_mode == 1){
		int need = Ac.cap;
		if(w->order_cap < need){
			if(w->d_order) cudaFree(w->d_order);
			void *p = NULL; CPB_CHECK(cudaMalloc(&p, sizeof(int)*(size_t)(need + 4))); w->d_order = (int *)p; w->order_cap = need;
		}
		LAUNCH(k_build_order, 1, 32, st, S, Ac, Tc, (const uint64_t *)w->d_user_order, w->n_user_order, w->d_order, w->d_order + need);
		CPB_CHECK(cudaMemcpyAsync(w->h_scratch, w->d_order + need, sizeof(int), cudaMemcpyDeviceToHost, st));
		CPB_CHECK(cudaStreamSynchronize(st));
		int n_order = *(int *)w->h_scratch;
		LAUNCH(k_solve_serial, 1, 32, st, B, Ac, J, (const int *)w->d_order, n_order, (const int *)w->d_joint_order, w->n_joint_order, iterations, dt, dt_coef);
		w->n_user_order = 0; w->n_joint_order = 0;
		w->last_solver_path = 0;
	} else {
		DColour &K = w->K;
		K.ids = S.ids;
		// (claim / bmask were cleared by k_integrate_vel, the colour histograms and worklist lengths by k_reset_step)
		if(ensure_worklists(w, Ac.cap + J.n + 64)) return -1;
		int use_hints = (w->hints_valid && !w->no_hints ? 1 : 0);
		if(w->no_phase_prefetch) use_hints |= 2;
		if(w->rows_strided) use_hints |= 4;
		w->hints_valid = true;
#ifndef CPB_EMU
		{
			DCounters *C = w->C; DRows R = w->R; unsigned *bar = w->d_barrier;
			if(w->sl_dirty && sl_refresh(w)) return -1;
			if(w->solver_variant == 3 && !w->sl_ok){ cpb_set_error("solver variant 3 (space-local) needs every space to own one contiguous body range that fits a CTA's shared memory"); return -1; }
			const SolvePlan plan = plan_solver(w);
			const int blocks = plan.blocks;
			const bool space_local = plan.space_local, stream_rows = plan.stream_rows;
			w->last_solver_path = (space_local ? 2 : 1);
			DSpaceLocal SL = w->SL;
			if(!space_local) SL.start = NULL;
			size_t nbuckets = 2*(size_t)w->n_spaces*CPB_MAX_COLOURS + 2;
			if(space_local) cudaMemsetAsync(SL.start, 0, sizeof(uint32_t)*nbuckets, st);
			DPrestep P = {(const DSpace *)w->d_spaces, dt, fused ? 1 : 0};
			void *args[] = {&B, &Ac, &J, &R, &K, &C, &bar, &SL, &use_hints, &iterations, &dt, &dt_coef, &P};
			const bool joints = (J.n > 0);
			// colouring (+ row build) and the iteration loop are separate cooperative launches: each gets its own register budget
			void *k_colour = space_local ? (void *)k_colour_solve<true, true, true, 1, 2> : (void *)k_colour_solve<false, true, true, 1, 2>;
			CPB_CHECK(cudaLaunchCooperativeKernel(k_colour, dim3(blocks), dim3(256), args, 0, st));
			g_cpb_launches++;
			if(fused && !space_local){
				STAGE_END(w, ST_COLOUR);
				if(launch_integrate_vel(w)) return -1;
				STAGE_END(w, ST_INTEGRATE_VEL);
			}
			if(!space_local){
				const int m3 = (w->solve_minb == 3 ? 1 : 0);
				static void *const k_iterate[2][2][2] = {   // [stream_rows][joints][min blocks 2 | 3]
					{{(void *)k_colour_solve<false, false, false, 2, 2>, (void *)k_colour_solve<false, false, false, 2, 3>},
					 {(void *)k_colour_solve<false, false, true, 2, 2>, (void *)k_colour_solve<false, false, true, 2, 3>}},
					{{(void *)k_colour_solve<false, true, false, 2, 2>, (void *)k_colour_solve<false, true, false, 2, 3>},
					 {(void *)k_colour_solve<false, true, true, 2, 2>, (void *)k_colour_solve<false, true, true, 2, 3>}}};
				const int iter_blocks = std::min(plan.iter_blocks, w->sm_count*(m3 ? 3 : 2));
				CPB_CHECK(cudaLaunchCooperativeKernel(k_iterate[stream_rows ? 1 : 0][joints ? 1 : 0][m3], dim3(iter_blocks), dim3(256), args, 0, st));
				g_cpb_launches++;
			}
			if(space_local){
				cpb_exclusive_scan(SL.start, SL.start, (int)nbuckets, w->sl_tmp, st);
				int g = std::min(grid_for(Ac.cap + J.n, 256), wide);
				LAUNCH(k_sl_rows, g, 256, st, B, Ac, J, R, SL, P);
				if(fused){
					STAGE_END(w, ST_COLOUR);
					if(launch_integrate_vel(w)) return -1;
					STAGE_END(w, ST_INTEGRATE_VEL);
				}
				const int threads = plan.threads;
				size_t smem = (size_t)w->sl_max_nbody*64;
				const bool fork = (w->n_sl_plain > 0 && w->n_sl_jointed > 0);
				cudaStream_t sj = (fork ? w->stream2 : st);
				if(fork){ CPB_CHECK(cudaEventRecord(w->ev_fork, st)); CPB_CHECK(cudaStreamWaitEvent(sj, w->ev_fork, 0)); }
				if(w->n_sl_plain) LAUNCH_SMEM(k_sl_solve<false>, w->n_sl_plain, threads, smem, st, B, Ac, J, R, SL, (const int *)w->d_sl_plain, iterations, dt, dt_coef);
				if(w->n_sl_jointed) LAUNCH_SMEM(k_sl_solve<true>, w->n_sl_jointed, threads, smem, sj, B, Ac, J, R, SL, (const int *)w->d_sl_jointed, iterations, dt, dt_coef);
				if(fork){ CPB_CHECK(cudaEventRecord(w->ev_join, sj)); CPB_CHECK(cudaStreamWaitEvent(st, w->ev_join, 0)); }
			}
		}
#else
		{
			LAUNCH(k_colour_seed, 4, 64, st, B, Ac, J, K, use_hints);
			for(int round = 0; round < CPB_MAX_COLOUR_ROUNDS; round++){
				if(K.wl_n[round] == 0) break;
				LAUNCH(k_colour_a, 4, 64, st, B, Ac, J, K, round);
				LAUNCH(k_colour_b, 4, 64, st, B, Ac, J, K, round);
			}
			if(K.wl_n[CPB_MAX_COLOUR_ROUNDS] != 0) LAUNCH(k_colour_finish, 4, 64, st, Ac, J, w->R, K, w->C, 2);
			LAUNCH(k_colour_finish, 1, 32, st, Ac, J, w->R, K, w->C, 0);
			LAUNCH(k_colour_finish, 4, 64, st, Ac, J, w->R, K, w->C, 1);
			int ncol = w->C->n_colours;
			for(int pass = 0; pass <= iterations; pass++){
				for(int c = 0; c < ncol; c++) LAUNCH(k_solve_colour, 4, 64, st, B, w->R, J, K, c, (pass == 0 ? 0 : 1), dt, dt_coef);
			}
			LAUNCH(k_rows_writeback, 4, 64, st, Ac, w->R, K);
			w->last_solver_path = 1;
		}
#endif
	}
	STAGE_END(w, ST_SOLVE);
	LAUNCH(k_finish_step, 1, 32, st, w->C, (const int *)w->P.count);
	if((w->io_src || w->io_sink) && nb > 0 && w->io_cap >= nb){
		if(w->io_sink){
			LAUNCH(k_pack_vel, grid_for(nb, 256), 256, st, B, w->d_io_vel);
			CPB_CHECK(cudaMemcpyAsync(w->io_sink + 3*(size_t)nb, w->d_io_vel, sizeof(double)*3*(size_t)nb, cudaMemcpyDeviceToHost, st));
		}
		CPB_CHECK(cudaEventRecord(w->ev_io_join, w->stream_io)); CPB_CHECK(cudaStreamWaitEvent(st, w->ev_io_join, 0));
	}

	w->steps++;
	if(w->profiling){
		CPB_CHECK(cudaStreamSynchronize(st));
		for(int i = 0; i < ST_COUNT; i++){
			float ms = 0.f;
			cudaEventElapsedTime(&ms, w->ev[i], w->ev[i + 1]);
			w->stage_us[i] = ms*1000.f;
		}
	}
	CPB_CHECK(cudaGetLastError());
	return 0;
}

#ifndef CPB_EMU
// ---- the step as a CUDA graph ----
static unsigned long long sig_mix(unsigned long long h, unsigned long long v){ h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); return h*0x100000001b3ull; }
static unsigned long long sig_ptr(unsigned long long h, const void *p){ return sig_mix(h, (unsigned long long)(uintptr_t)p); }

// Everything the launch sequence of a production step depends on: buffer addresses and sizes (kernel arguments are
// captured by value), dt and the iteration count, and the launch plan of the solver.
static unsigned long long step_signature(cpb200_world *w, double dt, int iterations, const SolvePlan &plan, bool rebuild)
{
	unsigned long long h = 0xcbf29ce484222325ull, d;
	memcpy(&d, &dt, 8);
	h = sig_mix(h, rebuild ? 2ull : 1ull); h = sig_mix(h, (unsigned long long)(w->bvh_period > 1 ? 1 : 0));
	// every group of device arrays: their generation counters (the pointers hashed below are a sample, not all of them)
	h = sig_mix(h, w->gB.gen); h = sig_mix(h, w->gS.gen); h = sig_mix(h, w->gJ.gen); h = sig_mix(h, w->gA.gen); h = sig_mix(h, w->gK.gen);
	h = sig_mix(h, w->gV.gen); h = sig_mix(h, w->gP.gen); h = sig_mix(h, w->gI.gen); h = sig_mix(h, w->gW.gen); h = sig_mix(h, w->gSL.gen);
	h = sig_mix(h, d); h = sig_mix(h, (unsigned long long)iterations);
	h = sig_mix(h, (unsigned long long)w->B.n); h = sig_mix(h, (unsigned long long)w->S.n); h = sig_mix(h, (unsigned long long)w->S.nv); h = sig_mix(h, (unsigned long long)w->J.n);
	h = sig_mix(h, (unsigned long long)w->cap_arbs); h = sig_mix(h, (unsigned long long)w->cap_pairs); h = sig_mix(h, (unsigned long long)w->n_spaces);
	h = sig_ptr(h, w->B.pos); h = sig_ptr(h, w->S.type); h = sig_ptr(h, w->J.type); h = sig_ptr(h, w->A[0].key); h = sig_ptr(h, w->P.cand);
	// (the radix sort ping-pongs between two key buffers and the host swaps its pointers when the sorted keys end up in the
	// second one; no step reads the previous step's keys, so a replayed graph may use them in either role: hash the pair)
	h = sig_ptr(h, std::min((const void *)w->bvh.keys, (const void *)w->keys_b)); h = sig_ptr(h, std::max((const void *)w->bvh.keys, (const void *)w->keys_b)); h = sig_ptr(h, w->K.wl[0]); h = sig_ptr(h, w->d_nocollide); h = sig_ptr(h, w->SL.start); h = sig_ptr(h, w->I.parent);
	h = sig_mix(h, (unsigned long long)w->n_nocollide);
	h = sig_ptr(h, w->io_src); h = sig_ptr(h, w->io_sink); h = sig_ptr(h, w->d_io_pos);
	h = sig_mix(h, (unsigned long long)((w->hints_valid && !w->no_hints) ? 1 : 0));
	h = sig_mix(h, (unsigned long long)(w->any_sleep_enabled ? 1 : 0));
	h = sig_mix(h, (unsigned long long)plan.iter_blocks);
	h = sig_mix(h, (unsigned long long)plan.blocks); h = sig_mix(h, (unsigned long long)((plan.space_local ? 1 : 0) | (plan.stream_rows ? 2 : 0))); h = sig_mix(h, (unsigned long long)plan.threads);
	h = sig_mix(h, (unsigned long long)w->sl_max_nbody); h = sig_mix(h, (unsigned long long)w->sl_max_nshape); h = sig_mix(h, (unsigned long long)((w->sl_shapes_ok ? 1 : 0) | (w->sl_ok ? 2 : 0)));
	h = sig_mix(h, (unsigned long long)w->n_sl_plain); h = sig_mix(h, (unsigned long long)w->n_sl_jointed);
	return h ? h : 1ull;
}

// host-side effects of step_phase_a + step_phase_b, for a replayed graph
static void step_bookkeeping(cpb200_world *w, double dt, int iterations, int solver_path, int launches, bool rebuild)
{
	if(bvh_in_use(w)){ if(rebuild){ w->bvh_valid = true; w->bvh_age = 1; } else w->bvh_age++; }
	w->stamp++;
	w->curr_dt = dt;
	w->cur ^= 1;
	w->arb_derived_stale[w->cur] = true;   // graphs hold production steps only (prestep folded into the row build)
	w->step_dt = dt; w->step_dt_coef = 1.0; w->step_iterations = iterations;
	w->mid_step = false; w->mid_solve = false;
	w->hints_valid = true;
	w->last_solver_path = solver_path;
	w->steps++;
	g_cpb_launches += (unsigned long long)launches;
}

static int step_graphed(cpb200_world *w, double dt)
{
	cudaSetDevice(w->device);
	cudaStream_t st = w->stream;
	int iterations = 0;
	for(int i = 0; i < w->n_spaces; i++) iterations = std::max(iterations, w->sp[(size_t)i].iterations);
	// A step qualifies when nothing on the host side has to happen inside it: same dt as the last step (dt_coef = 1, the
	// per-space and per-joint pow() terms are current), caches and layouts current, scratch buffers large enough.
	bool steady = (w->cap_arbs > 0 && !w->cache_dirty && w->curr_dt == dt && !w->spaces_dirty && w->spaces_dt == dt &&
	               (w->J.n == 0 || w->joints_dt == dt) && !w->sl_dirty && w->wl_cap >= w->A[0].cap + w->J.n + 64 &&
	               !(w->solver_variant == 3 && !w->sl_ok) && w->B.n > 0);
	if(!steady){
		memset(w->graph_last_sig, 0, sizeof(w->graph_last_sig));
		if(step_phase_a(w, dt)) return -1;
		return step_phase_b(w, true);
	}
	const SolvePlan plan = plan_solver(w);
	const bool rebuild = bvh_in_use(w) && bvh_rebuild_due(w);
	const unsigned long long sig = step_signature(w, dt, iterations, plan, rebuild);
	const int slot = (w->cur & 1)*2 + (rebuild ? 1 : 0);
	cpb200_world::StepGraph &G = w->graph[slot];
	if(G.exec && G.sig == sig){
		step_bookkeeping(w, dt, iterations, G.solver_path, G.launches, rebuild);
		CPB_CHECK(cudaGraphLaunch(G.exec, st));
		w->graph_replays++;
		return 0;
	}
	if(w->graph_last_sig[slot] != sig){
		// first step with this signature: run it as it is; if the next one of this slot looks the same it is captured
		w->graph_last_sig[slot] = sig;
		if(step_phase_a(w, dt)) return -1;
		return step_phase_b(w, true);
	}
	// capture: the step functions enqueue into the capturing stream; nothing executes until the graph is launched
	const uint32_t s_stamp = w->stamp; const double s_curr_dt = w->curr_dt; const int s_cur = w->cur; const uint64_t s_steps = w->steps;
	const bool s_hints = w->hints_valid; const unsigned long long s_launches = g_cpb_launches; const int s_path = w->last_solver_path;
	const bool s_bvh_valid = w->bvh_valid; const int s_bvh_age = w->bvh_age;
	if(G.exec){ cudaGraphExecDestroy(G.exec); G.exec = NULL; }
	cudaGraph_t graph = NULL;
	int rc = -1;
	cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
	const char *where = "cudaStreamBeginCapture";
	if(ce == cudaSuccess){
		rc = step_phase_a(w, dt);
		if(!rc) rc = step_phase_b(w, true);
		if(rc){ where = "enqueue under capture"; snprintf(w->graph_error, sizeof(w->graph_error), "%s", cpb200_last_error()); }
		ce = cudaStreamEndCapture(st, &graph);
		if(ce != cudaSuccess || !graph){ if(!rc) where = "cudaStreamEndCapture"; rc = -1; }
	}
	cudaGraphExec_t exec = NULL;
	if(!rc){ ce = cudaGraphInstantiate(&exec, graph, 0); if(ce != cudaSuccess){ rc = -1; where = "cudaGraphInstantiate"; } }
	if(graph) cudaGraphDestroy(graph);
	if(rc){
		size_t len = strlen(w->graph_error);
		snprintf(w->graph_error + len, sizeof(w->graph_error) - len, " [%s: %s]", where, cudaGetErrorString(ce));
		// this world's step cannot be captured (driver limitation): undo the host bookkeeping and run it the ordinary way from now on
		cudaGetLastError();
		w->stamp = s_stamp; w->curr_dt = s_curr_dt; w->cur = s_cur; w->steps = s_steps; w->hints_valid = s_hints; g_cpb_launches = s_launches;
		w->last_solver_path = s_path; w->mid_step = false; w->mid_solve = false; w->bvh_valid = s_bvh_valid; w->bvh_age = s_bvh_age;
		w->graph_enabled = false;
		if(step_phase_a(w, dt)) return -1;
		return step_phase_b(w, true);
	}
	G.exec = exec; G.sig = sig; G.launches = (int)(g_cpb_launches - s_launches); G.solver_path = w->last_solver_path;
	w->graph_captures++;
	CPB_CHECK(cudaGraphLaunch(G.exec, st));
	return 0;
}
#endif

extern "C" int cpb200_world_step(cpb200_world *w, double dt)
{
	if(!w){ cpb_set_error("null world"); return -1; }
	if(dt == 0.0) return 0; // cpSpaceStep.c:339
	if(w->mid_step || w->mid_solve){ cpb_set_error("cpb200_world_step while a split step is open (call cpb200_world_step_finish)"); return -1; }
#ifndef CPB_EMU
	if(w->graph_enabled && w->solver_mode == 0 && !w->profiling) return step_graphed(w, dt);
#endif
	if(step_phase_a(w, dt)) return -1;
	return step_phase_b(w, true);
}

/* Validation hook: replay the captured step graph (1, default unless CPB200_NO_GRAPH is set) or launch every kernel (0). */
extern "C" int cpb200_world_set_graph(cpb200_world *w, int on)
{
	if(!w) return -1;
	w->graph_enabled = (on != 0);
	return 0;
}

extern "C" const char *cpb200_world_graph_error(cpb200_world *w){ return w ? w->graph_error : ""; }

extern "C" int cpb200_world_get_graph_stats(cpb200_world *w, unsigned long long *out2)
{
	if(!w || !out2) return -1;
	out2[0] = w->graph_captures; out2[1] = w->graph_replays;
	return 0;
}

extern "C" int cpb200_world_step_collide(cpb200_world *w, double dt)
{
	if(!w){ cpb_set_error("null world"); return -1; }
	if(dt == 0.0){ cpb_set_error("split step with dt == 0"); return -1; }
	if(w->mid_step || w->mid_solve){ cpb_set_error("a split step is already open"); return -1; }
	return step_phase_a(w, dt);
}

extern "C" int cpb200_world_step_finish(cpb200_world *w)
{
	if(!w){ cpb_set_error("null world"); return -1; }
	if(w->mid_solve) return step_phase_b2(w, false);
	if(!w->mid_step){ cpb_set_error("cpb200_world_step_finish without cpb200_world_step_collide"); return -1; }
	return step_phase_b(w, false);
}

extern "C" int cpb200_world_step_presolve(cpb200_world *w)
{
	if(!w){ cpb_set_error("null world"); return -1; }
	if(!w->mid_step){ cpb_set_error("cpb200_world_step_presolve without cpb200_world_step_collide"); return -1; }
	return step_phase_b1(w, false);
}

// Host decisions of the begin/preSolve handlers applied to this step's records (cpSpaceStep.c:257-285,
// cpArbiter.c:46-50, 97-143).
__global__ void k_arb_edit(DArbs A, DCounters *C, const cpb200_arbiter_edit *edits, int n)
{
	int e = CPB_TID;
	if(e >= n) return;
	cpb200_arbiter_edit ed = edits[e];
	int i = ed.record;
	int nA = *A.count_ptr; if(nA > A.cap) nA = A.cap;
	if(i < 0 || i >= nA) return;
	if(ed.flags & CPB200_EDIT_MATERIAL){ A.e[i] = ed.e; A.u[i] = ed.u; A.svr[i] = make_double2(ed.surface_vr[0], ed.surface_vr[1]); }
	if(ed.flags & CPB200_EDIT_CONTACTS){
		A.n[i] = make_double2(ed.n[0], ed.n[1]);
		for(int k = 0; k < A.cnt[i] && k < 2; k++){
			A.r1[CIDX(A, i, k)] = make_double2(ed.r1[k][0], ed.r1[k][1]);
			A.r2[CIDX(A, i, k)] = make_double2(ed.r2[k][0], ed.r2[k][1]);
		}
	}
	if(ed.flags & (CPB200_EDIT_IGNORE | CPB200_EDIT_REJECT)){
		if(A.active[i] == 1){ atomicAdd(&C->n_active, -1); atomicAdd(&C->n_contacts, -A.cnt[i]); }
		A.active[i] = 0;
		// the reference drops the contacts of a rejected arbiter (arb->count = 0): nothing to warm start from
		A.cnt[i] = 0;
		if(ed.flags & CPB200_EDIT_IGNORE) A.state[i] = CPB200_ARB_IGNORE;
		else if(A.state[i] != CPB200_ARB_IGNORE) A.state[i] = CPB200_ARB_NORMAL;
	}
}

extern "C" int cpb200_world_edit_arbiters(cpb200_world *w, int n, const cpb200_arbiter_edit *edits)
{
	if(!w || n < 0 || (n > 0 && !edits)){ cpb_set_error("bad arguments"); return -1; }
	if(!w->mid_step){ cpb_set_error("arbiters can only be edited between cpb200_world_step_collide and cpb200_world_step_finish"); return -1; }
	if(n == 0) return 0;
	cudaSetDevice(w->device);
	size_t bytes = sizeof(cpb200_arbiter_edit)*(size_t)n;
	if(stage_reserve(w, bytes)) return -1;
	CPB_CHECK(cudaMemcpyAsync(w->d_stage, edits, bytes, cudaMemcpyHostToDevice, w->stream));
	LAUNCH(k_arb_edit, grid_for(n, 128), 128, w->stream, w->A[w->cur], w->C, (const cpb200_arbiter_edit *)w->d_stage, n);
	return world_sync(w);
}

extern "C" unsigned long long cpb200_launch_count(void)
{
#ifdef CPB_EMU
	return 0;
#else
	return g_cpb_launches;
#endif
}

// n steps bracketed by CUDA events on the world's own stream (what bench.py times)
extern "C" int cpb200_world_time_steps(cpb200_world *w, double dt, int n, float *ms)
{
	if(!w || n < 0 || !ms){ cpb_set_error("bad arguments"); return -1; }
	cudaSetDevice(w->device);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	CPB_CHECK(cudaStreamSynchronize(w->stream));
	cudaEventRecord(e0, w->stream);
	int rc = 0;
	for(int i = 0; i < n && rc == 0; i++) rc = cpb200_world_step(w, dt);
	cudaEventRecord(e1, w->stream);
	cudaEventSynchronize(e1);
	*ms = 0.f;
	cudaEventElapsedTime(ms, e0, e1);
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	if(rc) return rc;
	return cpb200_world_sync(w);
}

// hC holds a fresh copy of the device counters: refresh the grid-sizing hint, turn an overflow flag into an error
static int counters_check(cpb200_world *w)
{
	w->last_active = w->hC->n_active;
	// the counters are those of the last step: how many nodes its tree traversal visited per query.  Measured on a fresh
	// tree it is the yardstick; an aged topology that needs half as much again is rebuilt at the next step.
	if(w->bvh_valid && w->hC->bvh_queries > 64 && !w->bvh_no_valve){
		const double v = (double)w->hC->bvh_visits/(double)w->hC->bvh_queries;
		if(w->bvh_age <= 1) w->bvh_fresh_visits = v;
		else if(w->bvh_fresh_visits > 0.0 && v > 1.5*w->bvh_fresh_visits){ w->bvh_valid = false; w->bvh_fresh_visits = 0.0; }
	}
	if(w->hC->overflow){
		cpb_set_error("device buffer overflow (flags 0x%x: 1 pairs, 2 arbiters, 4 table, 8 bvh stack): collisions of the last step were dropped; "
			"call cpb200_world_reserve with larger capacities (in use: %d pair candidates, %d arbiter records)", w->hC->overflow, w->cap_pairs, w->cap_arbs);
		return -2;
	}
	return 0;
}

extern "C" int cpb200_world_sync(cpb200_world *w)
{
	if(!w) return -1;
	cudaSetDevice(w->device);
	if(world_sync(w)) return -1;
	CPB_CHECK(cudaMemcpyAsync(w->hC, w->C, sizeof(DCounters), cudaMemcpyDeviceToHost, w->stream));
	CPB_CHECK(cudaStreamSynchronize(w->stream));
	return counters_check(w);
}

// ------------------------------------------------------------------ read-back

extern "C" int cpb200_world_get_bodies(cpb200_world *w, int first, int n, cpb200_body_state *out)
{
	if(!w || first < 0 || n < 0 || first + n > w->B.n){ cpb_set_error("body range out of bounds"); return -1; }
	if(n == 0) return 0;
	cudaSetDevice(w->device);
	size_t bytes = sizeof(cpb200_body_state)*(size_t)n;
	if(stage_reserve(w, bytes)) return -1;
	LAUNCH(k_pack_body_state, grid_for(n, 128), 128, w->stream, w->B, (cpb200_body_state *)w->d_stage, first, n);
	CPB_CHECK(cudaMemcpyAsync(out, w->d_stage, bytes, cudaMemcpyDeviceToHost, w->stream));
	// the step counters ride along (64 bytes): a host that only ever reads bodies back -- the Chipmunk API layer -- still
	// learns about exhausted pair / arbiter buffers, and the solver's grid sizing gets a fresh constraint count
	CPB_CHECK(cudaMemcpyAsync(w->hC, w->C, sizeof(DCounters), cudaMemcpyDeviceToHost, w->stream));
	if(world_sync(w)) return -1;
	return counters_check(w);
}

extern "C" int cpb200_world_get_body_bias(cpb200_world *w, int first, int n, double *out)
{
	if(!w || first < 0 || n < 0 || first + n > w->B.n){ cpb_set_error("body range out of bounds"); return -1; }
	if(n == 0) return 0;
	cudaSetDevice(w->device);
	size_t bytes = 3*sizeof(double)*(size_t)n;
	if(stage_reserve(w, bytes)) return -1;
	LAUNCH(k_pack_body_bias, grid_for(n, 128), 128, w->stream, w->B, (double *)w->d_stage, first, n);
	CPB_CHECK(cudaMemcpyAsync(out, w->d_stage, bytes, cudaMemcpyDeviceToHost, w->stream));
	return world_sync(w);
}

static int ensure_cache(cpb200_world *w)
{
	if(w->cache_dirty){
		if(w->S.n) LAUNCH(k_shape_cache, grid_for(w->S.n, 128), 128, w->stream, w->S, w->B, 1);
		w->cache_dirty = false;
	}
	return 0;
}

extern "C" int cpb200_world_get_shape_bbs(cpb200_world *w, int first, int n, double *out)
{
	if(!w || first < 0 || n < 0 || first + n > w->S.n){ cpb_set_error("shape range out of bounds"); return -1; }
	cudaSetDevice(w->device);
	ensure_cache(w);
	if(n == 0) return 0;
	CPB_CHECK(cudaMemcpyAsync(out, w->S.bb + first, sizeof(double4)*(size_t)n, cudaMemcpyDeviceToHost, w->stream));
	return world_sync(w);
}

extern "C" int cpb200_world_get_arbiters(cpb200_world *w, int cap, cpb200_arbiter *out, int active_only)
{
	if(!w){ cpb_set_error("null world"); return -1; }
	cudaSetDevice(w->device);
	if(w->cap_arbs == 0) return 0;
	DArbs &A = w->A[w->cur];
	int n = 0;
	CPB_CHECK(cudaMemcpyAsync(&n, A.count_ptr, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
	CPB_CHECK(cudaMemcpyAsync(w->hC, w->C, sizeof(DCounters), cudaMemcpyDeviceToHost, w->stream));
	if(world_sync(w)) return -1;
	if(counters_check(w)) return -2;   // a host that reads arbiters (collision handlers) must not see a truncated list silently
	if(n > A.cap) n = A.cap;
	if(w->arb_derived_stale[w->cur] && out && n > 0){
		// the last step was a production step: nMass / tMass / bias lived in the solver's rows only (step_phase_b1)
		LAUNCH(k_arb_prestep, std::min(grid_for(A.cap, 128), w->sm_count*8), 128, w->stream, w->B, A, (const DSpace *)w->d_spaces, w->step_dt, 1);
		w->arb_derived_stale[w->cur] = false;
	}
	size_t N = (size_t)n;
	std::vector<int> sa, sb, ba, bb, cnt, state, active; std::vector<uint32_t> stamp; std::vector<V2> nn, svr, r1, r2;
	std::vector<double> e, u, nmass, tmass, bounce, bias, jn, jt, jb; std::vector<uint64_t> hash;
	if(download(w, sa, A.sa, N) || download(w, sb, A.sb, N) || download(w, ba, A.ba, N) || download(w, bb, A.bb, N) || download(w, cnt, A.cnt, N) ||
	   download(w, state, A.state, N) || download(w, active, A.active, N) || download(w, stamp, A.stamp, N) || download(w, nn, A.n, N) ||
	   download(w, svr, A.svr, N) || download(w, e, A.e, N) || download(w, u, A.u, N) || download2(w, r1, A.r1, N, A.cap) || download2(w, r2, A.r2, N, A.cap) ||
	   download2(w, nmass, A.nmass, N, A.cap) || download2(w, tmass, A.tmass, N, A.cap) || download2(w, bounce, A.bounce, N, A.cap) || download2(w, bias, A.bias, N, A.cap) ||
	   download2(w, jn, A.jn, N, A.cap) || download2(w, jt, A.jt, N, A.cap) || download2(w, jb, A.jb, N, A.cap) || download2(w, hash, A.hash, N, A.cap)) return -1;
	if(world_sync(w)) return -1;
	int m = 0;
	for(size_t i = 0; i < N; i++){
		if(active_only && active[i] != 1) continue;
		if(m < cap && out){
			cpb200_arbiter &o = out[m];
			memset(&o, 0, sizeof(o));
			o.shape_a = sa[i]; o.shape_b = sb[i]; o.body_a = ba[i]; o.body_b = bb[i];
			o.count = (active[i] || w->mid_step ? cnt[i] : 0); o.state = state[i]; o.stamp = stamp[i]; o.active = active[i]; o.record = (int32_t)i;
			o.n[0] = nn[i].x; o.n[1] = nn[i].y; o.e = e[i]; o.u = u[i]; o.surface_vr[0] = svr[i].x; o.surface_vr[1] = svr[i].y;
			for(int k = 0; k < 2; k++){
				size_t c = (size_t)k*N + i;
				o.contacts[k].r1[0] = r1[c].x; o.contacts[k].r1[1] = r1[c].y; o.contacts[k].r2[0] = r2[c].x; o.contacts[k].r2[1] = r2[c].y;
				o.contacts[k].n_mass = nmass[c]; o.contacts[k].t_mass = tmass[c]; o.contacts[k].bounce = bounce[c]; o.contacts[k].bias = bias[c];
				o.contacts[k].jn_acc = jn[c]; o.contacts[k].jt_acc = jt[c]; o.contacts[k].j_bias = jb[c]; o.contacts[k].hash = hash[c];
			}
		}
		m++;
	}
	return m;
}

__global__ void k_joint_state(DJoints J, cpb200_joint_state *out, int first, int n)
{
	int i = CPB_TID;
	if(i >= n) return;
	int j = first + i;
	out[i].acc[0] = J.acc[j].x; out[i].acc[1] = J.acc[j].y;
	out[i].impulse = joint_impulse(J, j);
	out[i].aux = J.aux0[j];
}

extern "C" int cpb200_world_get_joints(cpb200_world *w, int first, int n, cpb200_joint_state *out)
{
	if(!w || first < 0 || n < 0 || first + n > w->J.n){ cpb_set_error("joint range out of bounds"); return -1; }
	if(n == 0) return 0;
	cudaSetDevice(w->device);
	void *p = NULL;
	CPB_CHECK(cudaMalloc(&p, sizeof(cpb200_joint_state)*(size_t)n));
	LAUNCH(k_joint_state, grid_for(n, 128), 128, w->stream, w->J, (cpb200_joint_state *)p, first, n);
	cudaError_t e = cudaMemcpyAsync(out, p, sizeof(cpb200_joint_state)*(size_t)n, cudaMemcpyDeviceToHost, w->stream);
	int r = world_sync(w);
	cudaFree(p);
	if(e != cudaSuccess){ cpb_set_error("joint read-back failed"); return -1; }
	return r;
}

__global__ void k_stats(DBodies B, DArbs A, double *out, unsigned *awake)
{
	double ke = 0.0, pen = 0.0;
	unsigned aw = 0;
	for(int i = CPB_TID; i < B.n; i += CPB_NTHREADS){
		if(B.type[i] != CPB200_BODY_DYNAMIC) continue;
		if(!B.sleeping[i]) aw++;
		double4 V = B.V[i]; V2 M = B.M[i];
		double vsq = V.x*V.x + V.y*V.y, wsq = V.z*V.z;
		ke += (vsq ? vsq*M.x : 0.0) + (wsq ? wsq*M.y : 0.0); // cpBodyKineticEnergy (cpBody.c:581-588)
	}
	int nA = *A.count_ptr; if(nA > A.cap) nA = A.cap;
	for(int i = CPB_TID; i < nA; i += CPB_NTHREADS){
		if(A.active[i] != 1) continue;
		V2 n = A.n[i];
		V2 delta = vsub(B.pos[A.bb[i]], B.pos[A.ba[i]]);
		for(int k = 0; k < A.cnt[i]; k++){
			double dist = vdot(vadd(vsub(A.r2[CIDX(A, i, k)], A.r1[CIDX(A, i, k)]), delta), n);
			if(-dist > pen) pen = -dist;
		}
	}
	if(ke != 0.0) atomic_add_d(&out[0], ke);
	if(pen > 0.0) atomic_max_double(&out[1], pen);
	if(aw) atomicAdd(awake, aw);
}

extern "C" int cpb200_world_get_stats(cpb200_world *w, cpb200_stats *out)
{
	if(!w || !out){ cpb_set_error("bad arguments"); return -1; }
	cudaSetDevice(w->device);
	memset(out, 0, sizeof(*out));
	cudaMemsetAsync(w->d_scratch, 0, sizeof(double)*4, w->stream);
	if(w->cap_arbs){
		LAUNCH(k_stats, std::min(grid_for(std::max(w->B.n, 1), 256), w->sm_count*4), 256, w->stream, w->B, w->A[w->cur], w->d_scratch, (unsigned *)(w->d_scratch + 2));
	}
	CPB_CHECK(cudaMemcpyAsync(w->h_scratch, w->d_scratch, sizeof(double)*4, cudaMemcpyDeviceToHost, w->stream));
	CPB_CHECK(cudaMemcpyAsync(w->hC, w->C, sizeof(DCounters), cudaMemcpyDeviceToHost, w->stream));
	if(world_sync(w)) return -1;
	out->steps = w->steps;
	out->n_bodies = (uint32_t)w->B.n; out->n_shapes = (uint32_t)w->S.n; out->n_joints = (uint32_t)w->J.n;
	out->n_awake = *(unsigned *)(w->h_scratch + 2);
	out->n_pairs = (uint32_t)(w->hC->n_pairs[0] + w->hC->n_pairs[1] + w->hC->n_pairs[2]);
	w->last_active = w->hC->n_active;
	out->n_arbiters = (uint32_t)w->hC->n_active; out->n_contacts = (uint32_t)w->hC->n_contacts;
	out->n_cached = (uint32_t)w->hC->n_cached; out->n_colours = (uint32_t)w->hC->n_colours; out->overflow = (uint32_t)w->hC->overflow;
	out->kinetic_energy = w->h_scratch[0]; out->max_penetration = w->h_scratch[1];
	out->n_row_solves = (uint32_t)w->hC->n_row_solves; out->n_row_idle = (uint32_t)w->hC->n_row_idle;
	return 0;
}

// ------------------------------------------------------------------ validation hooks
extern "C" long cpb200_world_get_pairs(cpb200_world *w, long cap, uint64_t *out)
{
	if(!w){ cpb_set_error("null world"); return -1; }
	cudaSetDevice(w->device);
	if(w->cap_pairs == 0) return 0;
	int cnt[3] = {0, 0, 0};
	CPB_CHECK(cudaMemcpyAsync(cnt, w->P.count, sizeof(int)*3, cudaMemcpyDeviceToHost, w->stream));
	if(world_sync(w)) return -1;
	std::vector<uint64_t> keys;
	for(int c = 0; c < 3; c++){
		size_t n = (size_t)std::min(cnt[c], w->P.cap);
		std::vector<int> a, b;
		if(download(w, a, w->P.a[c], n) || download(w, b, w->P.b[c], n)) return -1;
		if(world_sync(w)) return -1;
		for(size_t i = 0; i < n; i++){
			uint64_t lo = (uint64_t)(uint32_t)std::min(a[i], b[i]), hi = (uint64_t)(uint32_t)std::max(a[i], b[i]);
			keys.push_back((lo << 32) | hi);
		}
	}
	std::sort(keys.begin(), keys.end());
	for(size_t i = 0; i < keys.size() && (long)i < cap; i++) out[i] = keys[i];
	return (long)keys.size();
}

extern "C" int cpb200_world_set_solver_mode(cpb200_world *w, int mode)
{
	if(!w || (mode != 0 && mode != 1)){ cpb_set_error("solver mode must be 0 or 1"); return -1; }
	w->solver_mode = mode;
	return 0;
}

extern "C" int cpb200_world_set_arbiter_order(cpb200_world *w, int n, const uint64_t *order)
{
	if(!w || n < 0){ cpb_set_error("bad arguments"); return -1; }
	cudaSetDevice(w->device);
	if(n > w->user_order_cap){
		if(w->d_user_order) cudaFree(w->d_user_order);
		void *p = NULL; CPB_CHECK(cudaMalloc(&p, sizeof(uint64_t)*(size_t)n)); w->d_user_order = (uint64_t *)p; w->user_order_cap = n;
	}
	if(n) CPB_CHECK(cudaMemcpyAsync(w->d_user_order, order, sizeof(uint64_t)*(size_t)n, cudaMemcpyHostToDevice, w->stream));
	w->n_user_order = n;
	return world_sync(w);
}

extern "C" int cpb200_world_set_solver_grid(cpb200_world *w, int blocks)
{
	if(!w || blocks < 0){ cpb_set_error("bad arguments"); return -1; }
	w->force_blocks = blocks;
	return 0;
}

extern "C" int cpb200_world_set_joint_order(cpb200_world *w, int n, const int32_t *order)
{
	if(!w || n < 0){ cpb_set_error("bad arguments"); return -1; }
	cudaSetDevice(w->device);
	if(n > w->joint_order_cap){
		if(w->d_joint_order) cudaFree(w->d_joint_order);
		void *p = NULL; CPB_CHECK(cudaMalloc(&p, sizeof(int)*(size_t)n)); w->d_joint_order = (int *)p; w->joint_order_cap = n;
	}
	if(n) CPB_CHECK(cudaMemcpyAsync(w->d_joint_order, order, sizeof(int)*(size_t)n, cudaMemcpyHostToDevice, w->stream));
	w->n_joint_order = n;
	return world_sync(w);
}

// ---- validation hooks for the PRODUCTION solver order (tests/test_gpu_production_order.py) ----
extern "C" int cpb200_world_set_solver_variant(cpb200_world *w, int variant)
{
	if(!w || variant < 0 || variant > 3){ cpb_set_error("solver variant must be 0..3"); return -1; }
	w->solver_variant = variant;
	return 0;
}

__global__ void k_pack_body_solver_state(DBodies B, double *__restrict__ dst, int first, int n)
{
	int k = CPB_TID;
	if(k >= n) return;
	double4 V = B.V[first + k], VB = B.VB[first + k];
	double *o = dst + 8*(size_t)k;
	o[0] = V.x; o[1] = V.y; o[2] = V.z; o[3] = V.w; o[4] = VB.x; o[5] = VB.y; o[6] = VB.z; o[7] = VB.w;
}

extern "C" int cpb200_world_get_body_solver_state(cpb200_world *w, int first, int n, double *out)
{
	if(!w || first < 0 || n < 0 || first + n > w->B.n){ cpb_set_error("body range out of bounds"); return -1; }
	if(n == 0) return 0;
	cudaSetDevice(w->device);
	size_t bytes = 8*sizeof(double)*(size_t)n;
	if(stage_reserve(w, bytes)) return -1;
	LAUNCH(k_pack_body_solver_state, grid_for(n, 128), 128, w->stream, w->B, (double *)w->d_stage, first, n);
	CPB_CHECK(cudaMemcpyAsync(out, w->d_stage, bytes, cudaMemcpyDeviceToHost, w->stream));
	return world_sync(w);
}

__global__ void k_pack_joint_solver_state(DJoints J, double *__restrict__ dst, int first, int n)
{
	int k = CPB_TID;
	if(k >= n) return;
	int j = first + k;
	double *o = dst + CPB200_JOINT_SOLVER_ROW*(size_t)k;
	double4 kk = J.k[j], prm = J.prm[j];
	o[0] = J.type[j]; o[1] = J.a[j]; o[2] = J.b[j]; o[3] = (J.colour[j] == -2 ? 0.0 : 1.0);
	o[4] = J.max_force[j]; o[5] = J.max_bias[j];
	o[6] = J.r1[j].x; o[7] = J.r1[j].y; o[8] = J.r2[j].x; o[9] = J.r2[j].y; o[10] = J.nrm[j].x; o[11] = J.nrm[j].y;
	o[12] = J.nmass[j]; o[13] = kk.x; o[14] = kk.y; o[15] = kk.z; o[16] = kk.w;
	o[17] = J.bias[j].x; o[18] = J.bias[j].y; o[19] = J.acc[j].x; o[20] = J.acc[j].y; o[21] = J.aux0[j]; o[22] = J.aux1[j];
	o[23] = prm.x; o[24] = prm.y; o[25] = prm.z; o[26] = prm.w; o[27] = 0.0;
}

extern "C" int cpb200_world_get_joint_solver_state(cpb200_world *w, int first, int n, double *out)
{
	if(!w || first < 0 || n < 0 || first + n > w->J.n){ cpb_set_error("joint range out of bounds"); return -1; }
	if(n == 0) return 0;
	cudaSetDevice(w->device);
	size_t bytes = CPB200_JOINT_SOLVER_ROW*sizeof(double)*(size_t)n;
	if(stage_reserve(w, bytes)) return -1;
	LAUNCH(k_pack_joint_solver_state, grid_for(n, 128), 128, w->stream, w->J, (double *)w->d_stage, first, n);
	CPB_CHECK(cudaMemcpyAsync(out, w->d_stage, bytes, cudaMemcpyDeviceToHost, w->stream));
	return world_sync(w);
}

// The sequence in which the last step's production solver visited its constraints, colour phase by colour phase
// (within a phase the constraints share no dynamic body and ran in parallel): item >= 0 = arbiter record index,
// item < 0 = joint -(item + 1).  Space-local path: space after space (spaces never interact).
extern "C" long cpb200_world_get_solver_order(cpb200_world *w, long cap, int64_t *out)
{
	if(!w){ cpb_set_error("null world"); return -1; }
	if(w->last_solver_path == 0){ cpb_set_error("the last step did not run the coloured solver"); return -1; }
	cudaSetDevice(w->device);
	if(world_sync(w)) return -1;
	std::vector<int64_t> seq;
	std::vector<int> jrow;
	if(download(w, jrow, w->J.row, (size_t)w->J.n) || world_sync(w)) return -1;
	if(w->last_solver_path == 1){
		std::vector<int> cstart, jstart, arb;
		if(download(w, cstart, w->K.cstart, CPB_MAX_COLOURS + 1) || download(w, jstart, w->K.jstart, CPB_MAX_COLOURS + 1) || world_sync(w)) return -1;
		int n_rows = std::min(cstart[CPB_MAX_COLOURS], w->R.cap);
		if(download(w, arb, w->R.arb, (size_t)n_rows) || world_sync(w)) return -1;
		for(int c = 0; c < CPB_MAX_COLOURS; c++){
			for(int r = cstart[c]; r < cstart[c + 1] && r < n_rows; r++) seq.push_back(arb[(size_t)r]);
			for(int q = jstart[c]; q < jstart[c + 1] && q < w->J.n; q++) seq.push_back(-(int64_t)jrow[(size_t)q] - 1);
		}
	} else {
		const int ns = w->n_spaces;
		size_t nbuckets = 2*(size_t)ns*CPB_MAX_COLOURS + 2;
		std::vector<uint32_t> start; std::vector<int4> hdr;
		if(download(w, start, (const uint32_t *)w->SL.start, nbuckets) || world_sync(w)) return -1;
		// after k_sl_rows every bucket word holds the END of its bucket = the begin of the next one
		const int jbase = (int)start[(size_t)ns*CPB_MAX_COLOURS];
		int n_rows = std::min(jbase, w->R.cap);
		if(download(w, hdr, (const int4 *)w->R.hdr, (size_t)n_rows) || world_sync(w)) return -1;
		for(int sp = 0; sp < ns; sp++){
			for(int c = 0; c < CPB_MAX_COLOURS; c++){
				int ka = sp*CPB_MAX_COLOURS + c, kj = ns*CPB_MAX_COLOURS + 1 + sp*CPB_MAX_COLOURS + c;
				int r0 = (ka > 0 ? (int)start[(size_t)ka - 1] : 0), r1 = (int)start[(size_t)ka];
				int q0 = (int)start[(size_t)kj - 1] - jbase, q1 = (int)start[(size_t)kj] - jbase;
				for(int r = r0; r < r1 && r < n_rows; r++) seq.push_back(hdr[(size_t)r].w);
				for(int q = q0; q < q1 && q < w->J.n; q++) seq.push_back(-(int64_t)jrow[(size_t)q] - 1);
			}
		}
	}
	for(size_t i = 0; i < seq.size() && (long)i < cap && out; i++) out[i] = seq[i];
	return (long)seq.size();
}

extern "C" int cpb200_world_get_solver_path(cpb200_world *w){ return w ? w->last_solver_path : -1; }

/* out[2][CPB_MAX_COLOURS + 1]: begin of every colour in the row list / in the joint list of the last world-wide coloured step */
extern "C" int cpb200_world_get_colour_starts(cpb200_world *w, int32_t *out)
{
	if(!w || !out || !w->K.cstart){ cpb_set_error("no colouring yet"); return -1; }
	cudaSetDevice(w->device);
	CPB_CHECK(cudaMemcpyAsync(out, w->K.cstart, sizeof(int)*(CPB_MAX_COLOURS + 1), cudaMemcpyDeviceToHost, w->stream));
	CPB_CHECK(cudaMemcpyAsync(out + CPB_MAX_COLOURS + 1, w->K.jstart, sizeof(int)*(CPB_MAX_COLOURS + 1), cudaMemcpyDeviceToHost, w->stream));
	return world_sync(w);
}

extern "C" int cpb200_world_collide_pair(cpb200_world *w, int shape_a, int shape_b, double *out13)
{
	if(!w || shape_a < 0 || shape_b < 0 || shape_a >= w->S.n || shape_b >= w->S.n){ cpb_set_error("shape index out of range"); return -1; }
	cudaSetDevice(w->device);
	ensure_cache(w);
	LAUNCH(k_collide_one, 1, 32, w->stream, w->S, w->B, shape_a, shape_b, w->d_scratch + 8);
	CPB_CHECK(cudaMemcpyAsync(w->h_scratch + 8, w->d_scratch + 8, sizeof(double)*13, cudaMemcpyDeviceToHost, w->stream));
	if(world_sync(w)) return -1;
	memcpy(out13, w->h_scratch + 8, sizeof(double)*13);
	return (int)out13[0];
}

// ------------------------------------------------------------------ space queries
static int query_reserve(cpb200_world *w, size_t bytes)
{
	bytes += 64;
	if(w->query_bytes >= bytes) return 0;
	if(w->d_query) cudaFree(w->d_query);
	w->d_query = NULL; w->query_bytes = 0;
	size_t want = bytes + bytes/2 + 4096;
	CPB_CHECK(cudaMalloc(&w->d_query, want));
	w->query_bytes = want;
	return 0;
}

static QFilter make_filter(const cpb200_filter *f)
{
	QFilter q; q.group = 0; q.categories = ~0u; q.mask = ~0u;     // CP_SHAPE_FILTER_ALL
	if(f){ q.group = f->group; q.categories = f->categories; q.mask = f->mask; }
	return q;
}

static void copy_hit(cpb200_query_hit *o, const QHit &h)
{
	o->shape = h.shape; o->pad = 0; o->point[0] = h.px; o->point[1] = h.py; o->d = h.d; o->g[0] = h.gx; o->g[1] = h.gy;
}

// all hits of one scan, sorted by shape index; retried once with a larger buffer if the first guess was short
static int query_all(cpb200_world *w, const QParams &Q, std::vector<QHit> &hits)
{
	cudaSetDevice(w->device);
	ensure_cache(w);
	hits.clear();
	if(w->S.n == 0) return 0;
	int cap = 1024;
	for(int attempt = 0; attempt < 2; attempt++){
		if(query_reserve(w, sizeof(QHit)*(size_t)cap)) return -1;
		int *count = (int *)w->d_query;
		QHit *out = (QHit *)((char *)w->d_query + 64);
		CPB_CHECK(cudaMemsetAsync(count, 0, sizeof(int), w->stream));
		LAUNCH(k_query_all, std::min(grid_for(w->S.n, 256), w->sm_count*8), 256, w->stream, w->S, w->B, Q, out, cap, count);
		int n = 0;
		CPB_CHECK(cudaMemcpyAsync(&n, count, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
		if(world_sync(w)) return -1;
		if(n <= cap){
			hits.resize((size_t)n);
			if(n){ CPB_CHECK(cudaMemcpyAsync(hits.data(), out, sizeof(QHit)*(size_t)n, cudaMemcpyDeviceToHost, w->stream)); if(world_sync(w)) return -1; }
			std::sort(hits.begin(), hits.end(), [](const QHit &a, const QHit &b){ return a.shape < b.shape; });
			return n;
		}
		cap = n;
	}
	cpb_set_error("query buffer could not be sized");
	return -1;
}

static int query_best(cpb200_world *w, const QParams &Q, cpb200_query_hit *out)
{
	cudaSetDevice(w->device);
	ensure_cache(w);
	if(w->S.n == 0) return 0;
	if(query_reserve(w, sizeof(QHit))) return -1;
	unsigned long long *key = (unsigned long long *)w->d_query;
	int *best = (int *)((char *)w->d_query + 8);
	QHit *rec = (QHit *)((char *)w->d_query + 64);
	CPB_CHECK(cudaMemsetAsync(key, 0xff, 8, w->stream));
	int big = 0x7fffffff;
	CPB_CHECK(cudaMemcpyAsync(best, &big, sizeof(int), cudaMemcpyHostToDevice, w->stream));
	int g = std::min(grid_for(w->S.n, 256), w->sm_count*8);
	for(int pass = 0; pass < 3; pass++) LAUNCH(k_query_best, g, 256, w->stream, w->S, w->B, Q, pass, key, best, rec);
	int h_best = 0; QHit h;
	CPB_CHECK(cudaMemcpyAsync(&h_best, best, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
	CPB_CHECK(cudaMemcpyAsync(&h, rec, sizeof(QHit), cudaMemcpyDeviceToHost, w->stream));
	if(world_sync(w)) return -1;
	if(h_best == big) return 0;
	if(out) copy_hit(out, h);
	return 1;
}

static QParams make_query(int kind, int space, int only_shape, int skip_sensors, const double a[2], const double b[2], double radius, const cpb200_filter *f)
{
	QParams Q;
	Q.kind = kind; Q.space = space; Q.only_shape = only_shape; Q.skip_sensors = skip_sensors;
	Q.a = make_double2(a[0], a[1]); Q.b = (b ? make_double2(b[0], b[1]) : make_double2(0.0, 0.0));
	Q.radius = radius; Q.filter = make_filter(f);
	return Q;
}

extern "C" int cpb200_world_point_query(cpb200_world *w, int space, const double point[2], double max_distance, const cpb200_filter *filter, int cap, cpb200_query_hit *out)
{
	if(!w || !point){ cpb_set_error("bad arguments"); return -1; }
	std::vector<QHit> hits;
	int n = query_all(w, make_query(0, space, -1, 0, point, NULL, max_distance, filter), hits);
	for(int i = 0; i < n && i < cap && out; i++) copy_hit(&out[i], hits[(size_t)i]);
	return n;
}

extern "C" int cpb200_world_point_query_nearest(cpb200_world *w, int space, const double point[2], double max_distance, const cpb200_filter *filter, cpb200_query_hit *out)
{
	if(!w || !point){ cpb_set_error("bad arguments"); return -1; }
	return query_best(w, make_query(0, space, -1, 1, point, NULL, max_distance, filter), out);
}

extern "C" int cpb200_world_segment_query(cpb200_world *w, int space, const double a[2], const double b[2], double radius, const cpb200_filter *filter, int cap, cpb200_query_hit *out)
{
	if(!w || !a || !b){ cpb_set_error("bad arguments"); return -1; }
	std::vector<QHit> hits;
	int n = query_all(w, make_query(1, space, -1, 0, a, b, radius, filter), hits);
	for(int i = 0; i < n && i < cap && out; i++) copy_hit(&out[i], hits[(size_t)i]);
	return n;
}

extern "C" int cpb200_world_segment_query_first(cpb200_world *w, int space, const double a[2], const double b[2], double radius, const cpb200_filter *filter, cpb200_query_hit *out)
{
	if(!w || !a || !b){ cpb_set_error("bad arguments"); return -1; }
	cpb200_query_hit h;
	int n = query_best(w, make_query(1, space, -1, 1, a, b, radius, filter), &h);
	if(n == 1 && !(h.d < 1.0)) n = 0;     // info.alpha < out->alpha with out->alpha = 1 (cpSpaceQuery.c:158)
	if(n == 1 && out) *out = h;
	return n;
}

extern "C" int cpb200_world_bb_query(cpb200_world *w, int space, const double bb[4], const cpb200_filter *filter, int cap, int32_t *shapes)
{
	if(!w || !bb){ cpb_set_error("bad arguments"); return -1; }
	std::vector<QHit> hits;
	int n = query_all(w, make_query(2, space, -1, 0, bb, bb + 2, 0.0, filter), hits);
	for(int i = 0; i < n && i < cap && shapes; i++) shapes[i] = hits[(size_t)i].shape;
	return n;
}

extern "C" int cpb200_world_shape_point_query(cpb200_world *w, int shape, const double point[2], cpb200_query_hit *out)
{
	if(!w || !point || shape < 0 || shape >= w->S.n){ cpb_set_error("shape index out of range"); return -1; }
	std::vector<QHit> hits;
	int n = query_all(w, make_query(0, -1, shape, 0, point, NULL, 0.0, NULL), hits);
	if(n == 1 && out) copy_hit(out, hits[0]);
	return n;
}

extern "C" int cpb200_world_shape_segment_query(cpb200_world *w, int shape, const double a[2], const double b[2], double radius, cpb200_query_hit *out)
{
	if(!w || !a || !b || shape < 0 || shape >= w->S.n){ cpb_set_error("shape index out of range"); return -1; }
	std::vector<QHit> hits;
	int n = query_all(w, make_query(1, -1, shape, 0, a, b, radius, NULL), hits);
	if(n == 1 && out) copy_hit(out, hits[0]);
	return n;
}

extern "C" int cpb200_world_shape_query(cpb200_world *w, int space, const cpb200_query_shape *q, const double *verts_normals, int cap, cpb200_shape_hit *out)
{
	if(!w || !q || (q->type == CPB200_SHAPE_POLY && (!verts_normals || q->count < 1))){ cpb_set_error("bad arguments"); return -1; }
	cudaSetDevice(w->device);
	ensure_cache(w);
	if(w->S.n == 0) return 0;
	QShape qs;
	qs.type = q->type; qs.count = (q->type == CPB200_SHAPE_POLY ? q->count : 0); qs.r = q->r; qs.self = q->self;
	qs.a = make_double2(q->a[0], q->a[1]); qs.b = make_double2(q->b[0], q->b[1]); qs.n = make_double2(q->n[0], q->n[1]);
	qs.rot = make_double2(q->rot[0], q->rot[1]);
	qs.atan = make_double2(q->a_tangent[0], q->a_tangent[1]); qs.btan = make_double2(q->b_tangent[0], q->b_tangent[1]);
	qs.bb = make_double4(q->bb[0], q->bb[1], q->bb[2], q->bb[3]);
	qs.pv = NULL; qs.pn = NULL;
	QFilter filter = make_filter(&q->filter);
	size_t poly_bytes = sizeof(V2)*2*(size_t)qs.count;
	poly_bytes = (poly_bytes + 63) & ~(size_t)63;
	int hcap = std::max(cap, 256);
	std::vector<QShapeHit> hits;
	for(int attempt = 0; attempt < 2; attempt++){
		if(query_reserve(w, poly_bytes + sizeof(QShapeHit)*(size_t)hcap)) return -1;
		int *count = (int *)w->d_query;
		char *base = (char *)w->d_query + 64;
		if(qs.count){
			std::vector<V2> pv((size_t)qs.count), pn((size_t)qs.count);
			for(int i = 0; i < qs.count; i++){ pv[(size_t)i] = make_double2(verts_normals[4*i], verts_normals[4*i + 1]); pn[(size_t)i] = make_double2(verts_normals[4*i + 2], verts_normals[4*i + 3]); }
			CPB_CHECK(cudaMemcpyAsync(base, pv.data(), sizeof(V2)*(size_t)qs.count, cudaMemcpyHostToDevice, w->stream));
			CPB_CHECK(cudaMemcpyAsync(base + sizeof(V2)*(size_t)qs.count, pn.data(), sizeof(V2)*(size_t)qs.count, cudaMemcpyHostToDevice, w->stream));
			if(world_sync(w)) return -1;     // pv / pn are stack vectors
			qs.pv = (const V2 *)base; qs.pn = (const V2 *)(base + sizeof(V2)*(size_t)qs.count);
		}
		QShapeHit *dout = (QShapeHit *)(base + poly_bytes);
		CPB_CHECK(cudaMemsetAsync(count, 0, sizeof(int), w->stream));
		LAUNCH(k_query_shape, std::min(grid_for(w->S.n, 128), w->sm_count*8), 128, w->stream, w->S, w->B, qs, filter, space, dout, hcap, count);
		int n = 0;
		CPB_CHECK(cudaMemcpyAsync(&n, count, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
		if(world_sync(w)) return -1;
		if(n <= hcap){
			hits.resize((size_t)n);
			if(n){ CPB_CHECK(cudaMemcpyAsync(hits.data(), dout, sizeof(QShapeHit)*(size_t)n, cudaMemcpyDeviceToHost, w->stream)); if(world_sync(w)) return -1; }
			std::sort(hits.begin(), hits.end(), [](const QShapeHit &a, const QShapeHit &b){ return a.shape < b.shape; });
			for(int i = 0; i < n && i < cap && out; i++){
				const QShapeHit &h = hits[(size_t)i];
				out[i].shape = h.shape; out[i].count = h.count; out[i].normal[0] = h.nx; out[i].normal[1] = h.ny;
				memcpy(out[i].points, h.pts, sizeof(h.pts));
			}
			return n;
		}
		hcap = n;
	}
	cpb_set_error("query buffer could not be sized");
	return -1;
}

extern "C" int cpb200_world_get_stage_times(cpb200_world *w, int cap, float *usec)
{
	if(!w) return -1;
	for(int i = 0; i < ST_COUNT && i < cap; i++) usec[i] = w->stage_us[i];
	return ST_COUNT;
}

extern "C" int cpb200_world_get_solver_profile(cpb200_world *w, double *usec5 /* [6] */)
{
	if(!w || !w->K.prof){ cpb_set_error("no solver profile"); return -1; }
	unsigned long long t[8];
	CPB_CHECK(cudaMemcpyAsync(t, w->K.prof, sizeof(t), cudaMemcpyDeviceToHost, w->stream));
	if(world_sync(w)) return -1;
	for(int i = 0; i < 4; i++) usec5[i] = (double)(t[i + 1] - t[i])*1e-3;
	usec5[4] = (double)t[5];
	usec5[5] = (double)t[6];
	return 0;
}

extern "C" int cpb200_world_set_profiling(cpb200_world *w, int on)
{
	if(!w) return -1;
	w->profiling = (on != 0);
	return 0;
}

// ------------------------------------------------------------------ primitive self-tests
// (exercised by tests/test_gpu_prims.py; host buffers in, host buffers out)
extern "C" int cpb200_debug_sort_pairs(int device, int n, int bits, uint64_t *keys, int32_t *vals)
{
	if(n < 0) return -1;
	cudaSetDevice(device);
	uint64_t *ka = NULL, *kb = NULL; int *va = NULL, *vb = NULL; uint32_t *tmp = NULL;
	size_t N = (size_t)(n > 0 ? n : 1);
	void *p = NULL;
	CPB_CHECK(cudaMalloc(&p, 8*N)); ka = (uint64_t *)p;
	CPB_CHECK(cudaMalloc(&p, 8*N)); kb = (uint64_t *)p;
	CPB_CHECK(cudaMalloc(&p, 4*N)); va = (int *)p;
	CPB_CHECK(cudaMalloc(&p, 4*N)); vb = (int *)p;
	CPB_CHECK(cudaMalloc(&p, 4*(cpb_sort_tmp_elems(n) + 16))); tmp = (uint32_t *)p;
	CPB_CHECK(cudaMemcpy(ka, keys, 8*(size_t)n, cudaMemcpyHostToDevice));
	CPB_CHECK(cudaMemcpy(va, vals, 4*(size_t)n, cudaMemcpyHostToDevice));
	int where = cpb_radix_sort(ka, va, kb, vb, n, bits, tmp, 0);
	CPB_CHECK(cudaDeviceSynchronize());
	CPB_CHECK(cudaMemcpy(keys, where ? kb : ka, 8*(size_t)n, cudaMemcpyDeviceToHost));
	CPB_CHECK(cudaMemcpy(vals, where ? vb : va, 4*(size_t)n, cudaMemcpyDeviceToHost));
	cudaFree(ka); cudaFree(kb); cudaFree(va); cudaFree(vb); cudaFree(tmp);
	CPB_CHECK(cudaGetLastError());
	return 0;
}

extern "C" int cpb200_debug_exclusive_scan(int device, int n, uint32_t *data)
{
	if(n < 0) return -1;
	cudaSetDevice(device);
	uint32_t *d = NULL, *tmp = NULL;
	void *p = NULL;
	CPB_CHECK(cudaMalloc(&p, 4*(size_t)(n > 0 ? n : 1))); d = (uint32_t *)p;
	CPB_CHECK(cudaMalloc(&p, 4*(cpb_scan_tmp_elems(n) + 16))); tmp = (uint32_t *)p;
	CPB_CHECK(cudaMemcpy(d, data, 4*(size_t)n, cudaMemcpyHostToDevice));
	cpb_exclusive_scan(d, d, n, tmp, 0);
	CPB_CHECK(cudaDeviceSynchronize());
	CPB_CHECK(cudaMemcpy(data, d, 4*(size_t)n, cudaMemcpyDeviceToHost));
	cudaFree(d); cudaFree(tmp);
	CPB_CHECK(cudaGetLastError());
	return 0;
}
