// gpx_wide.cu — the tick of ONE large world (hundreds to hundreds of thousands of bodies): BASELINE config 4.
//
// Same JPH_PhysicsSystem_Update(system, dt, 2, jobSystem) contract as gpx_tick.cu (engine/src/physics/MapPhysics.c:
// 105-108), same device functions (gpx_solver.cuh) and therefore the same fp32 results per contact, but a different
// decomposition: state lives in global memory (L2-resident at 100k bodies) and every stage is one thread per item.
//
//   kw_begin      thread/body      SoA -> work record (first sub-step), forces, inertia, AABB, largest extents
//   kw_keys       thread/body      sort key = (row along z, lower x bound, body)
//   radix sort    (gpx_bvh.cu)     bodies ordered by row, then by the lower x bound of their boxes (bitonic below 4096)
//   kw_gather     thread/slot      sorted boxes packed into two float4 streams
//   kw_sweep      thread/body      sort-and-sweep broadphase: walk forward in the own row and through the reachable
//                                  part of the next row, exact box test, layer matrix -> pair list (atomic append)
//   kw_pairs      thread/pair      box-box / sphere contact manifolds
//   kw_static     thread/body      LBVH candidates (cached), box/sphere-vs-triangle manifolds grouped by normal
//   kw_link       thread/manifold  warm-start lookup in a hash table of the previous sub-step's manifolds, incidence
//                                  lists per dynamic body, colouring priority, union-find over dynamic-dynamic contacts
//   kw_isl_*      thread/item      simulation islands (connected components of the contact graph): manifold count per
//                                  island; islands of at most 32 manifolds packed into 32-slot windows
//   kw_island     warp/window      small islands inside one warp: the same colouring, set-up, warm start, velocity rows
//                                  and integration, with warp barriers in place of grid-wide ones and the bodies'
//                                  velocities in shared memory (a stack of boxes is an island; config 4 has 10 000)
//   kw_colour     cooperative      the remaining (large) islands: Jones-Plassmann colouring with hashed priorities
//                                  (result depends on the contact graph only, not on list order), per-colour lists
//   kw_solve      cooperative      large islands: set-up; warm start; 10 x (colour by colour) velocity rows; integrate
//                                  (also free and kinematic bodies); 2 x (colour by colour) position rows — grid-wide
//                                  barriers between colours, or block barriers in one block when the islands are few
//   kw_island_pos warp/window      the small islands' position rows, after the kinematic bodies have moved
//   kw_finish     thread/item      hash table for the next sub-step; wake marks; last sub-step: work records -> SoA
//   kw_sleep_*    thread/item      once per tick: islands of the awake bodies, sleep test, whole islands to sleep
//
// No float atomics and no order-dependent reductions: within a colour no two manifolds share a dynamic body, so the
// result is a pure function of the contact set.  Islands do not share dynamic bodies either, and the colouring of a
// manifold depends only on its own island, so solving an island inside a warp gives the bits the grid-wide phases give.
#include <cooperative_groups.h>
#include <cstdlib>

#include "gpx_solver.cuh"

namespace cg = cooperative_groups;

namespace gpx {

constexpr int WIDE_MAXADJ = 32;      // manifolds incident to one dynamic body (a 128-byte row of the incidence table)
constexpr int WIDE_MAXCOL = 64;
constexpr uint32_t WT = 128;         // threads per block of the per-item kernels
constexpr uint32_t NARROW_T = 64;    // threads per block of the narrowphase kernels (shared polygon scratch)

enum WideCounter { WC_NMAN = 0, WC_NPREV, WC_NCOL, WC_ERR, WC_UNCOLOURED, WC_NACTIVE, WC_MAXEXT_X, WC_MAXEXT_Z, WC_COLCNT = 8,
				   WC_COLOFF = WC_COLCNT + WIDE_MAXCOL, WC_COLCUR = WC_COLOFF + WIDE_MAXCOL + 1, WC_NISL = WC_COLCUR + WIDE_MAXCOL,
				   WC_ISLCUR, WC_NBIG, WC_NMED, WC_MEDCUR, WC_UNSORTED, WC_COUNT };
constexpr uint32_t MEDIUM_MAX = 1024;    // manifolds of an island solved by one block (kw_island_block); more: the cooperative kernels
constexpr uint32_t MEDIUM_T = 128;       // threads of that block
constexpr uint32_t SINGLE_BLOCK_MAX = 4096;  // manifolds of large islands up to which ONE block colours and solves them
constexpr uint32_t ISLAND_MAX = 32;      // manifolds of an island solved inside one warp (one per lane)
constexpr uint32_t ISLAND_WARPS = 4;     // warps (islands) per block of kw_island
constexpr uint32_t ROOT_SMALL = 0x80000000u;
constexpr int SENSOR_COLOUR = -0x40000000;  // a touching pair with a sensor in it: a contact event, never solved (far from the
                                           // -4 - colour codes kw_island leaves behind)

// Everything a velocity phase needs about one manifold, in colour order: the phases stream these records (coalesced)
// instead of chasing list -> manifold -> bodies -> parked constants.
struct __align__(16) SolveRec
{
	Con con;
	Rows rows;
	uint32_t mi;
};
static_assert(sizeof(SolveRec) % 16 == 0, "solver records are copied in 16-byte pieces");

struct WideDevice
{
	uint32_t nb = 0, n_pad = 0, cap_m = 0, hsize = 0, isl_slots = 0;
	SBody *bodies = nullptr;
	unsigned long long *keys = nullptr, *keys_tmp = nullptr;  // keys_tmp / sort_hist: radix sort scratch
	// the order of the last sort (body per sorted position) and this sub-step's keys by body: bodies move little between
	// sub-steps, so the keys laid out in the last order are sorted by two passes of tile sorts almost every time, and the
	// radix sort behind them returns at once
	uint32_t *order = nullptr, *sort_fallbacks = nullptr;
	unsigned long long *bkeys = nullptr;
	uint32_t *sort_hist = nullptr;
	float4 *boxlo = nullptr, *boxhi = nullptr;
	SMan *man[2] = {nullptr, nullptr};
	uint32_t *ord[2] = {nullptr, nullptr};  // ordinal of a manifold among those with the same (a, b)
	int cur = 0;
	SolveRec *recs = nullptr;
	uint32_t *counters = nullptr;
	unsigned long long *hkeys = nullptr;
	uint32_t *hvals = nullptr;
	uint32_t *adj = nullptr, *adj_n = nullptr;
	uint32_t *prio = nullptr;
	int *pending = nullptr;
	uint32_t *col_list = nullptr;
	// islands: union-find parent per body, flattened root (| ROOT_SMALL when the body is solved by kw_island), manifold
	// count / list offset / fill cursor per root, roots of the small islands, their manifold lists
	uint32_t *parent = nullptr, *root_of = nullptr, *isl_cnt = nullptr, *isl_off = nullptr, *isl_cur = nullptr;
	uint32_t *isl_man = nullptr, *big_list = nullptr, *med_list = nullptr, *med_coloff = nullptr;
	uint32_t med_slots = 0;  // most medium islands a world can hold
	unsigned char *can_sleep = nullptr;  // per island root: every body of the island is a sleep candidate
	// contact events (gpx_events_enable): sorted keys of the touching pairs, scratch for the diff against the last tick
	uint32_t n_ev = 0;
	unsigned long long *ev_keys = nullptr, *ev_tmp = nullptr, *ev_cur = nullptr;
	uint32_t *ev_hist = nullptr, *ev_flag = nullptr, *ev_pos = nullptr, *ev_ncur = nullptr;
	int coop_grid_colour = 0, coop_grid_solve = 0;
	// the tick as a CUDA graph: the launch sequence depends on nothing the host sees during a tick, so it is captured
	// once per signature (arguments, sub-steps, buffer parity, sleep / event passes) and replayed
	cudaGraphExec_t graph = nullptr;
	unsigned char graph_sig[1024];
	size_t graph_sig_n = 0;
	uint64_t graph_launches = 0;
	bool graph_off = false;
};

struct WideArgs
{
	BodyStore bs;
	SBody *bodies;
	unsigned long long *keys;
	float4 *boxlo, *boxhi;
	SMan *man, *prev;
	uint32_t *ord, *prev_ord;
	uint32_t *order, *sort_fallbacks;
	unsigned long long *bkeys;  // this sub-step's keys in body order; `keys` is the sorted array
	SolveRec *recs;
	uint32_t *cnt;
	unsigned long long *hkeys;
	uint32_t *hvals;
	uint32_t *adj, *adj_n;
	uint32_t *prio;
	int *pending;
	uint32_t *col_list;
	uint32_t *parent, *root_of, *isl_cnt, *isl_off, *isl_cur, *isl_man, *big_list;
	uint32_t *med_list, *med_coloff;  // roots of the medium islands; 65 colour offsets per medium island
	uint32_t med_max;                 // MEDIUM_MAX, or 0 with GPX_WIDE_NO_ISLANDS
	uint32_t med_slots;               // capacity of med_list
	uint4 *cand;
	StaticView sv;
	uint32_t nb, n_pad, cap_m, hmask;
	uint32_t cap, rows_per_world;  // bodies per world; rows of the sweep each world owns (4096 for a single world)
	uint32_t isl_max;  // islands up to this many manifolds are solved inside a warp (ISLAND_MAX; 0 sends everything to the phased kernels)
	uint32_t vel_steps, pos_steps;
	float gx, gy, gz, h;
	int first, last;
};

__device__ __forceinline__ uint32_t sortable(float f)
{
	uint32_t b = __float_as_uint(f);
	return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

__device__ __forceinline__ unsigned long long man_key(uint32_t a, uint32_t b, uint32_t ord)
{
	return ((unsigned long long)(a + 1u) << 40) | ((unsigned long long)b << 8) | (unsigned long long)(ord & 0xFFu);
}
__device__ __forceinline__ uint32_t key_slot(unsigned long long k, uint32_t mask)
{
	k ^= k >> 33;
	k *= 0xFF51AFD7ED558CCDull;
	k ^= k >> 33;
	k *= 0xC4CEB9FE1A85EC53ull;
	k ^= k >> 33;
	return (uint32_t)k & mask;
}
// Colouring priority: a fixed hash of the manifold's identity (the CPU restatement of the tick uses the same constants)
__device__ __forceinline__ uint32_t man_prio(uint32_t a, uint32_t b, uint32_t ord)
{
	uint32_t h = (a * 0x9E3779B1u) ^ ((b + ord * 0x7F4A7C15u) * 0x85EBCA77u);
	h ^= h >> 15;
	h *= 0x2C1B3C6Du;
	h ^= h >> 12;
	h *= 0x297A2D39u;
	h ^= h >> 15;
	return h;
}

// ---------------------------------------------------------------------------------------------------- per body

// Sort keys of slots without a box (dead, empty shape, padding): above every live key, and still carrying the slot, so
// that the sorted array stays a permutation of the slots.
__device__ __forceinline__ unsigned long long dead_key(uint32_t i) { return (~0ull << 20) | (unsigned long long)i; }
__device__ __forceinline__ bool is_dead_key(unsigned long long k) { return (k >> 20) == (~0ull >> 20); }

// one body of kw_begin; `ex` / `ez` return the extents of its box along x and z (0 when it has none)
__device__ __forceinline__ void begin_body(const WideArgs &a, uint32_t i, float &ex, float &ez)
{
	if (i >= a.nb)
	{
		a.bkeys[i] = dead_key(i);
		return;
	}
	SBody &b = a.bodies[i];
	if (a.first)
	{
		const float4 p = a.bs.pos[i], q = a.bs.quat[i], l = a.bs.lin[i], w = a.bs.ang[i];
		const float4 p0 = a.bs.prop0[i], p1 = a.bs.prop1[i], p2 = a.bs.prop2[i];
		b.x = V(p);
		b.q = Q(q);
		b.v = V(l);
		b.w = V(w);
		b.inv_mass = p0.x;
		b.inv_i = V(p0.y, p0.z, p0.w);
		b.he = V(p1);
		b.friction = p1.w;
		b.lin_damp = p2.x;
		b.ang_damp = p2.y;
		b.grav = p2.z;
		b.restitution = p2.w;
		uint32_t f0 = a.bs.flags[i];
		if ((f0 & BF_ALIVE) && ((f0 >> BF_MOTION_SHIFT) & 3u) == GPX_MOTION_KINEMATIC &&
			(l.x != 0.0f || l.y != 0.0f || l.z != 0.0f || w.x != 0.0f || w.y != 0.0f || w.z != 0.0f))
			f0 |= BF_KIN_MOVING;
		b.flags = f0;
	}
	a.adj_n[i] = 0;
	a.parent[i] = i;
	a.isl_cnt[i] = 0;
	a.isl_cur[i] = 0;
	const uint32_t f = b.flags;
	if (!(f & BF_ALIVE))
	{
		a.bkeys[i] = dead_key(i);
		return;
	}
	const float h = a.h;
	if (is_dynamic(f))
	{
		const uint32_t dofs = dofs_of(f);
		const v3 gravity = V(a.gx, a.gy, a.gz);
		b.v = b.v + (gravity * (h * b.grav));
		b.v = b.v * fmaxf(0.0f, 1.0f - (b.lin_damp * h));
		b.w = b.w * fmaxf(0.0f, 1.0f - (b.ang_damp * h));
		b.v = clamp_len(mask_lin(dofs, b.v), MAX_LINEAR_VELOCITY);
		v3 ww = b.w;
		if (!(dofs & 8u)) ww.x = 0.0f;
		if (!(dofs & 16u)) ww.y = 0.0f;
		if (!(dofs & 32u)) ww.z = 0.0f;
		b.w = clamp_len(ww, MAX_ANGULAR_VELOCITY);
	}
	body_world_inertia(b);
	body_aabb(b);
	if (shape_of(f) == GPX_SHAPE_EMPTY)
	{
		a.bkeys[i] = dead_key(i);
		return;
	}
	a.bkeys[i] = 0ull;  // filled by kw_keys once the largest extents are known
	ex = fmaxf(b.hi.x - b.lo.x, 0.0f);
	ez = fmaxf(b.hi.z - b.lo.z, 0.0f);
}

__global__ void __launch_bounds__(WT) kw_begin(WideArgs a)
{
	const uint32_t i = blockIdx.x * WT + threadIdx.x;
	float ex = 0.0f, ez = 0.0f;
	if (i < a.n_pad) begin_body(a, i, ex, ez);
	// the largest extents: one atomic per warp, not one per body (10^5 atomics on two words of one sector serialise);
	// non-negative floats order like their bit patterns
	const uint32_t mx = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(ex));
	const uint32_t mz = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(ez));
	if ((threadIdx.x & 31u) == 0u)
	{
		if (mx) atomicMax(&a.cnt[WC_MAXEXT_X], mx);
		if (mz) atomicMax(&a.cnt[WC_MAXEXT_Z], mz);
	}
}

// Sort key: [row: 12 bits][lower x bound, order-preserving: 32 bits][body: 20 bits].  Rows are slabs along z at least
// as thick as the largest body (plus the contact margin), so two bodies that overlap in z sit in the same or in
// adjacent rows; within a row the bodies are ordered by the lower x bound of their boxes.
constexpr int WIDE_ROWS = 4096;
__device__ __forceinline__ float row_thickness(const WideArgs &a)
{
	return (__uint_as_float(a.cnt[WC_MAXEXT_Z]) + (4.0f * SPECULATIVE_DISTANCE)) + 1.0e-3f;
}
// Every world owns `rows` consecutive rows (all 4096 when there is one world); bodies beyond a world's slab range share its
// first / last row, which costs candidates, not correctness.  Worlds beyond the row count share rows too: the sweep's
// exact test rejects pairs of different worlds.
__device__ __forceinline__ uint32_t row_of(float lo_z, float thickness, uint32_t world, uint32_t rows)
{
	const float r = floorf(lo_z / thickness) + (float)(rows / 2u);
	const uint32_t local = (uint32_t)fminf(fmaxf(r, 0.0f), (float)(rows - 1u));
	return min(world * rows + local, (uint32_t)WIDE_ROWS - 1u);
}
__device__ __forceinline__ unsigned long long sweep_key(uint32_t row, float lo_x, uint32_t body)
{
	return ((unsigned long long)row << 52) | ((unsigned long long)sortable(lo_x) << 20) | (unsigned long long)body;
}

__global__ void __launch_bounds__(WT) kw_keys(WideArgs a)
{
	const uint32_t i = blockIdx.x * WT + threadIdx.x;
	if (i >= a.nb || is_dead_key(a.bkeys[i])) return;
	const SBody &b = a.bodies[i];
	a.bkeys[i] = sweep_key(row_of(b.lo.z, row_thickness(a), i / a.cap, a.rows_per_world), b.lo.x, i);
}

// ---- the sort, given the order of the last one
// Sort of one tile of SORT_TILE keys in shared memory; tile t covers [offset + t * SORT_TILE, ...).  The first pass (`order`
// given) reads this sub-step's keys in the order the last sort left the bodies in and sorts each aligned tile — most are
// sorted as they come, which one comparison per key finds out; the rest take a bitonic sort.  The second pass runs over the
// tiles shifted by half a tile: both halves of such a window are sorted by then, so it only has to merge them, and only
// where the two keys in the middle are out of order.  Together they sort any array whose keys are less than half a tile
// from their place; kw_sorted_check says whether that was enough.
constexpr uint32_t SORT_TILE = 2048;
__device__ __forceinline__ void tile_exchange(unsigned long long *s, uint32_t i, uint32_t l, bool ascending)
{
	const unsigned long long x = s[i], y = s[l];
	if ((x > y) == ascending)
	{
		s[i] = y;
		s[l] = x;
	}
}
__global__ void __launch_bounds__(SORT_TILE / 2) kw_tile_sort(unsigned long long *keys, const unsigned long long *bkeys,
															  const uint32_t *order, uint32_t n, uint32_t offset)
{
	__shared__ unsigned long long s[SORT_TILE];
	const uint32_t base = offset + blockIdx.x * SORT_TILE, t = threadIdx.x;
	for (uint32_t k = t; k < SORT_TILE; k += SORT_TILE / 2)
		s[k] = base + k < n ? (order ? bkeys[order[base + k]] : keys[base + k]) : ~0ull;
	__syncthreads();
	const bool merge_only = order == nullptr;
	bool unsorted;
	if (merge_only)
		unsorted = s[SORT_TILE / 2 - 1] > s[SORT_TILE / 2];  // the same answer in every thread
	else
		unsorted = __syncthreads_or((s[2 * t] > s[2 * t + 1]) || (2 * t + 2 < SORT_TILE && s[2 * t + 1] > s[2 * t + 2]));
	if (unsorted)
	{
		if (merge_only)
		{
			// bitonic merge of two ascending halves: mirror step, then the half-cleaners
			tile_exchange(s, t, SORT_TILE - 1u - t, true);
			__syncthreads();
			for (uint32_t j = SORT_TILE >> 2; j > 0; j >>= 1)
			{
				const uint32_t i = ((t & ~(j - 1u)) << 1) | (t & (j - 1u));
				tile_exchange(s, i, i | j, true);
				__syncthreads();
			}
		}
		else
			for (uint32_t k = 2; k <= SORT_TILE; k <<= 1)
				for (uint32_t j = k >> 1; j > 0; j >>= 1)
				{
					const uint32_t i = ((t & ~(j - 1u)) << 1) | (t & (j - 1u));
					tile_exchange(s, i, i | j, (i & k) == 0u);
					__syncthreads();
				}
	}
	else if (merge_only)
		return;  // in place already
	for (uint32_t k = t; k < SORT_TILE; k += SORT_TILE / 2)
		if (base + k < n) keys[base + k] = s[k];
}

__global__ void __launch_bounds__(WT) kw_sorted_check(WideArgs a)
{
	const uint32_t p = blockIdx.x * WT + threadIdx.x;
	if (p + 1 < a.n_pad && a.keys[p] > a.keys[p + 1])
		if (atomicExch(&a.cnt[WC_UNSORTED], 1u) == 0u) a.sort_fallbacks[0]++;
}

__global__ void kw_iota(uint32_t *v, uint32_t n)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) v[i] = i;
}

__global__ void __launch_bounds__(WT) kw_gather(WideArgs a)
{
	const uint32_t p = blockIdx.x * WT + threadIdx.x;
	if (p >= a.n_pad) return;
	const unsigned long long k = a.keys[p];
	const uint32_t i = (uint32_t)(k & 0xFFFFFull);
	a.order[p] = i;
	if (is_dead_key(k))
	{
		a.boxlo[p] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu));
		return;
	}
	const SBody &b = a.bodies[i];
	a.boxlo[p] = F4(b.lo, __uint_as_float(i));
	a.boxhi[p] = F4(b.hi, __uint_as_float(b.flags));
}

// Sort-and-sweep over rows: body p walks forward in its own row while boxes can still overlap on x, then through
// the part of the next row that can overlap it (found by binary search on the sort keys).  The exact test is the
// ensemble kernel's, evaluated with the lower-numbered body first, so the pair set equals an all-pairs test.
__device__ __forceinline__ void sweep_test(WideArgs &a, float4 lo_p, float4 hi_p, uint32_t ip, uint32_t fp, uint32_t q)
{
	const float4 lo_q = a.boxlo[q], hi_q = a.boxhi[q];
	const uint32_t iq = __float_as_uint(lo_q.w), fq = __float_as_uint(hi_q.w);
	if (ip / a.cap != iq / a.cap) return;  // bodies of different worlds never meet
	// at least one awake dynamic body — or a moving kinematic body reaching a sleeper, which only wakes it
	if (!is_dynamic(fp) && !is_dynamic(fq) &&
		!(((fp & BF_KIN_MOVING) && (fq & BF_ASLEEP)) || ((fq & BF_KIN_MOVING) && (fp & BF_ASLEEP))))
		return;
	if (!layers_collide(layer_of(fp), layer_of(fq))) return;
	const bool p_first = ip < iq;
	const v3 alo = p_first ? V(lo_p) : V(lo_q), ahi = p_first ? V(hi_p) : V(hi_q);
	const v3 blo = p_first ? V(lo_q) : V(lo_p), bhi = p_first ? V(hi_q) : V(hi_p);
	if (!aabb_overlap(alo, ahi, blo, bhi, SPECULATIVE_DISTANCE)) return;
	// one slot per pair, one atomic per group of lanes that found a pair together (slot order never matters: every
	// later decision goes by body indices)
	const unsigned found = __activemask();
	const unsigned lane = threadIdx.x & 31u;
	const int leader = __ffs((int)found) - 1;
	uint32_t k = 0;
	if ((int)lane == leader) k = atomicAdd(&a.cnt[WC_NMAN], (uint32_t)__popc(found));
	k = __shfl_sync(found, k, leader) + (uint32_t)__popc(found & ((1u << lane) - 1u));
	if (k >= a.cap_m)
	{
		atomicOr(&a.cnt[WC_ERR], (uint32_t)GPX_ERR_CONTACT_CONSTRAINTS_FULL);
		return;
	}
	SMan &m = a.man[k];
	m.a = p_first ? ip : iq;
	m.b = p_first ? iq : ip;
	m.np = 0;
	m.tri = 0;
	m.colour = -2;
	a.ord[k] = 0;
}

__global__ void __launch_bounds__(WT) kw_sweep(WideArgs a)
{
	const uint32_t p = blockIdx.x * WT + threadIdx.x;
	if (p >= a.n_pad) return;
	const unsigned long long key_p = a.keys[p];
	if (is_dead_key(key_p)) return;
	const float4 lo_p = a.boxlo[p], hi_p = a.boxhi[p];
	const uint32_t ip = __float_as_uint(lo_p.w), fp = __float_as_uint(hi_p.w);
	const uint32_t row = (uint32_t)(key_p >> 52);
	const float reach = (hi_p.x + SPECULATIVE_DISTANCE) + SPECULATIVE_DISTANCE;  // a little beyond the exact test's reach
	// own row: later bodies only (earlier ones find p themselves)
	for (uint32_t q = p + 1; q < a.n_pad; q++)
	{
		const unsigned long long kq = a.keys[q];
		if (is_dead_key(kq) || (uint32_t)(kq >> 52) != row) break;
		if (a.boxlo[q].x > reach) break;
		sweep_test(a, lo_p, hi_p, ip, fp, q);
	}
	// next row: everything whose box can reach back to lo_p.x
	if (row + 1 >= (uint32_t)WIDE_ROWS) return;
	const float back = (lo_p.x - __uint_as_float(a.cnt[WC_MAXEXT_X])) - (4.0f * SPECULATIVE_DISTANCE);
	const unsigned long long want = sweep_key(row + 1, back, 0u);
	uint32_t lo = p + 1, hi = a.n_pad;  // first q with keys[q] >= want (dead keys sort last, so the array is sorted)
	while (lo < hi)
	{
		const uint32_t mid = (lo + hi) >> 1;
		if (a.keys[mid] < want) lo = mid + 1;
		else hi = mid;
	}
	for (uint32_t q = lo; q < a.n_pad; q++)
	{
		const unsigned long long kq = a.keys[q];
		if (is_dead_key(kq) || (uint32_t)(kq >> 52) != row + 1) break;
		if (a.boxlo[q].x > reach) break;
		sweep_test(a, lo_p, hi_p, ip, fp, q);
	}
}

// ---------------------------------------------------------------------------------------------------- narrowphase

__global__ void __launch_bounds__(NARROW_T) kw_pairs(WideArgs a)
{
	__shared__ Scratch scratch[NARROW_T];
	const uint32_t k = blockIdx.x * NARROW_T + threadIdx.x;
	const uint32_t n = min(a.cnt[WC_NMAN], a.cap_m);  // only pairs have been appended so far
	if (k >= n) return;
	SMan &m = a.man[k];
	pair_contact(a.bodies[m.a], a.bodies[m.b], scratch[threadIdx.x], m);
	const uint32_t fa = a.bodies[m.a].flags, fb = a.bodies[m.b].flags;
	if (((fa | fb) & BF_ASLEEP) && m.np > 0 && !((fa | fb) & BF_SENSOR))
	{
		// a contact with an active body wakes a sleeper (applied by kw_finish: awake from the next sub-step on)
		if ((fa & BF_ASLEEP) && is_active_body(fb)) atomicOr(&a.bodies[m.a].flags, BF_WAKE_MARK);
		if ((fb & BF_ASLEEP) && is_active_body(fa)) atomicOr(&a.bodies[m.b].flags, BF_WAKE_MARK);
		if (!is_dynamic(fa) && !is_dynamic(fb)) m.np = 0;  // kinematic against sleeper: nothing to solve
	}
}

__global__ void __launch_bounds__(NARROW_T) kw_static(WideArgs a)
{
	__shared__ Scratch scratch[NARROW_T];
	const uint32_t i = blockIdx.x * NARROW_T + threadIdx.x;
	if (i >= a.nb) return;
	const SBody &A = a.bodies[i];
	const uint32_t fa = A.flags;
	if (!(fa & BF_ALIVE) || shape_of(fa) == GPX_SHAPE_EMPTY) return;
	const uint32_t la = layer_of(fa);
	if (!is_dynamic(fa) || (fa & BF_SENSOR) || !(la == 1 || la == 2)) return;
	StaticSlot slots[MAX_STATIC_PER_BODY];
	uint32_t err = 0;
	const int nslots = body_static_contacts(a.sv, a.cand + 8ull * i, A, scratch[threadIdx.x], slots, err);
	if (err) atomicOr(&a.cnt[WC_ERR], err);
	if (nslots == 0) return;
	const uint32_t base = atomicAdd(&a.cnt[WC_NMAN], (uint32_t)nslots);
	for (int s = 0; s < nslots; s++)
	{
		const uint32_t k = base + s;
		if (k >= a.cap_m)
		{
			atomicOr(&a.cnt[WC_ERR], (uint32_t)GPX_ERR_CONTACT_CONSTRAINTS_FULL);
			break;
		}
		SMan &m = a.man[k];
		m.a = i;
		m.b = STATIC_BODY_BASE + slots[s].sbody;
		m.tri = slots[s].tri;
		m.colour = -2;
		m.n = slots[s].n;
		m.friction = slots[s].friction;
		m.restitution = A.restitution;
		store_points(m, A, nullptr, slots[s].np, slots[s].p1, slots[s].p2);
		uint32_t o = 0;
		for (int t = 0; t < s; t++) o += slots[t].sbody == slots[s].sbody ? 1u : 0u;
		a.ord[k] = o;
	}
}

// ---------------------------------------------------------------------------------------------------- linking

__device__ __forceinline__ int hash_find(const WideArgs &a, unsigned long long key)
{
	uint32_t s = key_slot(key, a.hmask);
	for (uint32_t probe = 0; probe <= a.hmask; probe++)
	{
		const unsigned long long k = a.hkeys[s];
		if (k == key) return (int)a.hvals[s];
		if (k == 0ull) return -1;
		s = (s + 1u) & a.hmask;
	}
	return -1;
}

// Union-find over bodies, lock-free: a root is only ever hooked under a smaller index, so the root of a finished
// component is its smallest body whatever the interleaving.
__device__ __forceinline__ uint32_t isl_find(uint32_t *parent, uint32_t x)
{
	uint32_t p = *(volatile uint32_t *)&parent[x];
	while (p != x)
	{
		const uint32_t g = *(volatile uint32_t *)&parent[p];
		if (g != p) atomicMin(&parent[x], g);  // path halving; only ever lowers an entry towards its root
		x = p;
		p = g;
	}
	return x;
}
__device__ __forceinline__ void isl_unite(uint32_t *parent, uint32_t x, uint32_t y)
{
	for (;;)
	{
		x = isl_find(parent, x);
		y = isl_find(parent, y);
		if (x == y) return;
		if (x > y)
		{
			const uint32_t t = x;
			x = y;
			y = t;
		}
		if (atomicCAS(&parent[y], y, x) == y) return;
	}
}

__global__ void __launch_bounds__(WT) kw_link(WideArgs a)
{
	const uint32_t mi = blockIdx.x * WT + threadIdx.x;
	const uint32_t n = min(a.cnt[WC_NMAN], a.cap_m);
	if (mi >= n) return;
	SMan &m = a.man[mi];
	if (m.np == 0)
	{
		m.colour = -2;
		return;
	}
	if (m.b < STATIC_BODY_BASE && ((a.bodies[m.a].flags | a.bodies[m.b].flags) & BF_SENSOR))
	{
		m.colour = SENSOR_COLOUR;  // sensor overlaps are events, not contacts
		return;
	}
	// warm start: previous manifolds of the same body pair, in creation order (ordinal 0, 1, ...)
	bool got_cf = false;
	for (uint32_t o = 0; o < (uint32_t)MAX_SLOTS; o++)
	{
		const int j = hash_find(a, man_key(m.a, m.b, o));
		if (j < 0) break;
		const SMan &old = a.prev[j];
		if (old.tri != m.tri) continue;  // another slot of the same static body (the wall next to the floor)
		for (int p = 0; p < m.np; p++)
		{
			if (m.ln[p] != 0.0f) continue;
			for (int k = 0; k < old.np; k++)
				if (len2(m.p1l[p] - old.p1l[k]) < PRESERVE_LAMBDA_MAX_DIST_SQ &&
					len2(m.p2l[p] - old.p2l[k]) < PRESERVE_LAMBDA_MAX_DIST_SQ)
				{
					m.ln[p] = old.ln[k];
					// the friction impulse of the manifold comes from the first old manifold a point is found in
					if (!got_cf)
					{
						got_cf = true;
						m.cf[0] = old.cf[0];
						m.cf[1] = old.cf[1];
						m.cf[2] = old.cf[2];
					}
					break;
				}
		}
		if (m.b < STATIC_BODY_BASE) break;  // body pairs have exactly one manifold
	}
	// incidence lists of the dynamic bodies (what the colouring walks)
	const bool a_dyn = is_dynamic(a.bodies[m.a].flags);
	const bool b_dyn = m.b < STATIC_BODY_BASE && is_dynamic(a.bodies[m.b].flags);
	if (a_dyn)
	{
		const uint32_t k = atomicAdd(&a.adj_n[m.a], 1u);
		if (k < (uint32_t)WIDE_MAXADJ) a.adj[m.a * WIDE_MAXADJ + k] = mi;
		else atomicOr(&a.cnt[WC_ERR], (uint32_t)GPX_ERR_CONTACT_CONSTRAINTS_FULL);
	}
	if (b_dyn)
	{
		const uint32_t k = atomicAdd(&a.adj_n[m.b], 1u);
		if (k < (uint32_t)WIDE_MAXADJ) a.adj[m.b * WIDE_MAXADJ + k] = mi;
		else atomicOr(&a.cnt[WC_ERR], (uint32_t)GPX_ERR_CONTACT_CONSTRAINTS_FULL);
	}
	// priorities come from the bodies' indices inside their world, so a world colours the same wherever it sits
	a.prio[mi] = man_prio(m.a % a.cap, m.b < STATIC_BODY_BASE ? m.b % a.cap : m.b, a.ord[mi]);
	m.colour = -1;
	if (a_dyn && b_dyn) isl_unite(a.parent, m.a, m.b);
}


// ---------------------------------------------------------------------------------------------------- islands

// the dynamic end of a manifold names its island (both ends of a dynamic-dynamic contact share one root)
__device__ __forceinline__ uint32_t man_island_body(const WideArgs &a, const SMan &m)
{
	return is_dynamic(a.bodies[m.a].flags) ? m.a : m.b;
}

__global__ void __launch_bounds__(WT) kw_isl_count(WideArgs a)
{
	const uint32_t mi = blockIdx.x * WT + threadIdx.x;
	const uint32_t n = min(a.cnt[WC_NMAN], a.cap_m);
	if (mi >= n) return;
	const SMan &m = a.man[mi];
	if (m.np == 0 || m.colour == SENSOR_COLOUR) return;
	const uint32_t r = isl_find(a.parent, man_island_body(a, m));
	a.pending[mi] = (int)r;
	atomicAdd(&a.isl_cnt[r], 1u);
}

// Small islands are packed into 32-slot windows of isl_man, one window per warp of kw_island: an island never
// straddles a window, several short ones share one.  Each block packs the roots among its bodies and reserves whole
// windows with one atomic; unused slots keep the 0xFFFFFFFF the list was cleared to.
__global__ void __launch_bounds__(WT) kw_isl_place(WideArgs a)
{
	__shared__ uint32_t scnt[WT], soff[WT], sbase;
	const uint32_t i = blockIdx.x * WT + threadIdx.x;
	uint32_t c = 0;
	if (i < a.nb)
	{
		const uint32_t f = a.bodies[i].flags;
		uint32_t r = i;
		bool small = false;
		if ((f & BF_ALIVE) && is_dynamic(f) && a.adj_n[i] > 0)
		{
			r = isl_find(a.parent, i);
			small = a.isl_cnt[r] <= max(a.isl_max, a.med_max);  // solved (and integrated) outside the cooperative kernels
		}
		a.root_of[i] = small ? (r | ROOT_SMALL) : r;
		c = a.isl_cnt[i];  // non-zero only at roots
		if (c > a.isl_max)
		{
			if (c <= a.med_max)
			{
				// a medium island: its own range, counted from the top of the large-island arrays, and a block
				a.isl_off[i] = atomicAdd(&a.cnt[WC_MEDCUR], c);
				a.med_list[atomicAdd(&a.cnt[WC_NMED], 1u)] = i;
			}
			c = 0;
		}
	}
	scnt[threadIdx.x] = c;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		uint32_t pos = 0, islands = 0;
		for (uint32_t k = 0; k < WT; k++)
		{
			const uint32_t ck = scnt[k];
			if (!ck) continue;
			if ((pos & 31u) + ck > 32u) pos = (pos + 31u) & ~31u;
			soff[k] = pos;
			pos += ck;
			islands++;
		}
		const uint32_t total = (pos + 31u) & ~31u;
		sbase = total ? atomicAdd(&a.cnt[WC_ISLCUR], total) : 0u;
		if (islands) atomicAdd(&a.cnt[WC_NISL], islands);
	}
	__syncthreads();
	if (c) a.isl_off[i] = sbase + soff[threadIdx.x];
}

__global__ void __launch_bounds__(WT) kw_isl_fill(WideArgs a)
{
	const uint32_t mi = blockIdx.x * WT + threadIdx.x;
	const uint32_t n = min(a.cnt[WC_NMAN], a.cap_m);
	bool big = false;
	if (mi < n)
	{
		SMan &m = a.man[mi];
		if (m.np > 0 && m.colour != SENSOR_COLOUR)
		{
			const uint32_t r = (uint32_t)a.pending[mi];
			const uint32_t cnt_r = a.isl_cnt[r];
			if (cnt_r <= a.isl_max)
			{
				a.isl_man[a.isl_off[r] + atomicAdd(&a.isl_cur[r], 1u)] = mi;
				m.colour = -3;  // not the cooperative kernels' business
			}
			else if (cnt_r <= a.med_max)
				a.big_list[(a.cap_m - a.isl_off[r] - cnt_r) + atomicAdd(&a.isl_cur[r], 1u)] = mi;  // stays -1: kw_island_block colours it
			else
				big = true;
		}
	}
	// what is left for kw_colour / kw_solve: one list slot each, one atomic per warp
	const uint32_t votes = __ballot_sync(0xFFFFFFFFu, big);
	if (!votes) return;
	const uint32_t lane = threadIdx.x & 31u;
	uint32_t base = 0;
	if (lane == 0)
	{
		base = atomicAdd(&a.cnt[WC_NBIG], (uint32_t)__popc(votes));
		atomicAdd(&a.cnt[WC_UNCOLOURED], (uint32_t)__popc(votes));
	}
	base = __shfl_sync(0xFFFFFFFFu, base, 0);
	if (big) a.big_list[base + (uint32_t)__popc(votes & ((1u << lane) - 1u))] = mi;
}

// One warp = one 32-slot window of isl_man = a few whole islands, one manifold per lane.  The sequence per manifold
// is exactly kw_colour + kw_solve's; only the barriers are warp barriers.  Lanes of different islands never refer to
// each other (neighbour masks stay inside an island), so they simply share the colour phases.
__global__ void __launch_bounds__(ISLAND_WARPS * 32) kw_island(WideArgs a)
{
	__shared__ float vel[ISLAND_WARPS][64][6];
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t win = blockIdx.x * ISLAND_WARPS + (threadIdx.x >> 5);
	if (win * 32u >= a.cnt[WC_ISLCUR]) return;
	const uint32_t mi = a.isl_man[win * 32u + lane];
	const bool have = mi != 0xFFFFFFFFu;
	const uint32_t FULL = 0xFFFFFFFFu;
	Rows R;  // this lane's manifold: rows in registers for the whole solve

	// --- neighbours (manifolds sharing a dynamic body) as a lane mask
	if (have) a.pending[mi] = (int)lane;
	__syncwarp();
	uint32_t nbr = 0, ma = 0, mb = 0, mord = 0, mprio = 0;
	bool a_dyn = false, b_dyn = false;
	if (have)
	{
		const SMan &m = a.man[mi];
		ma = m.a;
		mb = m.b;
		mord = a.ord[mi];
		mprio = a.prio[mi];
		a_dyn = is_dynamic(a.bodies[ma].flags);
		b_dyn = mb < STATIC_BODY_BASE && is_dynamic(a.bodies[mb].flags);
		const uint32_t ends[2] = {ma, mb};
		const bool dyn[2] = {a_dyn, b_dyn};
		for (int e = 0; e < 2; e++)
		{
			if (!dyn[e]) continue;
			const uint32_t cntb = min(a.adj_n[ends[e]], (uint32_t)WIDE_MAXADJ);
			for (uint32_t k = 0; k < cntb; k++)
			{
				const uint32_t other = a.adj[ends[e] * WIDE_MAXADJ + k];
				if (other != mi) nbr |= 1u << (uint32_t)a.pending[other];
			}
		}
	}
	// lanes anybody has as a neighbour: the only ones worth broadcasting from
	uint32_t any_nbr = nbr;
	for (int o = 16; o > 0; o >>= 1) any_nbr |= __shfl_xor_sync(FULL, any_nbr, o);
	// --- which neighbours outrank this manifold (same total order as `outranks`)
	uint32_t higher = 0;
	for (uint32_t rest = any_nbr; rest; rest &= rest - 1u)
	{
		const uint32_t j = (uint32_t)__ffs((int)rest) - 1u;
		const uint32_t pj = __shfl_sync(FULL, mprio, j), aj = __shfl_sync(FULL, ma, j), bj = __shfl_sync(FULL, mb, j),
					   oj = __shfl_sync(FULL, mord, j);
		if (!((nbr >> j) & 1u)) continue;
		bool up;
		if (pj != mprio) up = pj > mprio;
		else if (aj != ma) up = aj > ma;
		else if (bj != mb) up = bj > mb;
		else up = oj > mord;
		if (up) higher |= 1u << j;
	}
	// --- Jones-Plassmann rounds
	int colour = have ? -1 : -2;
	for (int round = 0; round < 64; round++)
	{
		const uint32_t uncol = __ballot_sync(FULL, colour == -1);
		if (!uncol) break;
		unsigned long long used = 0ull;
		for (uint32_t rest = any_nbr; rest; rest &= rest - 1u)
		{
			const uint32_t j = (uint32_t)__ffs((int)rest) - 1u;
			const int cj = __shfl_sync(FULL, colour, j);
			if (((nbr >> j) & 1u) && cj >= 0) used |= 1ull << cj;
		}
		if (colour == -1 && !(higher & uncol))
		{
			int d = __ffsll((long long)~used) - 1;
			if (d < 0 || d >= WIDE_MAXCOL)
			{
				d = WIDE_MAXCOL - 1;
				atomicOr(&a.cnt[WC_ERR], (uint32_t)GPX_ERR_CONTACT_CONSTRAINTS_FULL);
			}
			colour = d;
		}
	}
	int ncol = colour + 1;
	for (int o = 16; o > 0; o >>= 1) ncol = max(ncol, __shfl_xor_sync(FULL, ncol, o));

	// --- set-up
	const float h = a.h;
	Con c;
	if (have)
	{
		build_rows(c, R, a.man[mi], a.bodies, h);
	}
	// --- velocities of the island's dynamic bodies live in shared memory while the rows run.  A body belongs to the
	// lane that holds the first manifold of its incidence list (slot 2 * lane + end); the other lanes look the slot up.
	float(*sv)[6] = vel[threadIdx.x >> 5];
	const bool own_a = have && a_dyn && a.adj[ma * WIDE_MAXADJ] == mi;
	const bool own_b = have && b_dyn && a.adj[mb * WIDE_MAXADJ] == mi;
	if (own_a)
	{
		const SBody &b = a.bodies[ma];
		a.isl_cur[ma] = 2u * lane;
		float *d = sv[2u * lane];
		d[0] = b.v.x; d[1] = b.v.y; d[2] = b.v.z; d[3] = b.w.x; d[4] = b.w.y; d[5] = b.w.z;
	}
	if (own_b)
	{
		const SBody &b = a.bodies[mb];
		a.isl_cur[mb] = 2u * lane + 1u;
		float *d = sv[2u * lane + 1u];
		d[0] = b.v.x; d[1] = b.v.y; d[2] = b.v.z; d[3] = b.w.x; d[4] = b.w.y; d[5] = b.w.z;
	}
	__syncwarp();
	// ends that do not move under impulses (kinematic, or no second body) keep their velocity in registers
	Vel fixed;
	fixed.va = fixed.wa = fixed.vb = fixed.wb = V(0.0f, 0.0f, 0.0f);
	float *pa = nullptr, *pb = nullptr;
	if (have)
	{
		load_vel(c, a.bodies, fixed);
		if (a_dyn) pa = sv[a.isl_cur[ma]];
		if (b_dyn) pb = sv[a.isl_cur[mb]];
	}
	// --- warm start (it == 0), then the velocity iterations, colour by colour
	for (uint32_t it = 0; it <= a.vel_steps; it++)
		for (int col = 0; col < ncol; col++)
		{
			if (colour == col)
			{
				Vel u = fixed;
				if (pa)
				{
					u.va = V(pa[0], pa[1], pa[2]);
					u.wa = V(pa[3], pa[4], pa[5]);
				}
				if (pb)
				{
					u.vb = V(pb[0], pb[1], pb[2]);
					u.wb = V(pb[3], pb[4], pb[5]);
				}
				if (it == 0)
					warm_start(c, R, R.f, u);
				else
					solve_velocity(c, R, R.f, u, it - 1u);
				if (pa)
				{
					pa[0] = u.va.x; pa[1] = u.va.y; pa[2] = u.va.z; pa[3] = u.wa.x; pa[4] = u.wa.y; pa[5] = u.wa.z;
				}
				if (pb)
				{
					pb[0] = u.vb.x; pb[1] = u.vb.y; pb[2] = u.vb.z; pb[3] = u.wb.x; pb[4] = u.wb.y; pb[5] = u.wb.z;
				}
			}
			__syncwarp();
		}
	if (have)
	{
		save_impulses(a.man[mi], R);
		// --- velocities back, and integrate: every dynamic body of the island once, by the lane that owns it
		const uint32_t ends[2] = {ma, mb};
		const bool own[2] = {own_a, own_b};
		for (int e = 0; e < 2; e++)
		{
			if (!own[e]) continue;
			SBody &b = a.bodies[ends[e]];
			const float *d = sv[2u * lane + (uint32_t)e];
			const v3 v = V(d[0], d[1], d[2]), w = V(d[3], d[4], d[5]);
			b.v = v;
			b.w = w;
			b.x = b.x + (v * h);
			b.q = qstep(b.q, w * h);
		}
	}
	// the position rows run in kw_island_pos, after the kinematic bodies have moved (kw_solve integrates them)
	if (have) a.man[mi].colour = -4 - colour;
}

// Position iterations of the small islands (one window per warp), colours as left by kw_island.
__global__ void __launch_bounds__(ISLAND_WARPS * 32) kw_island_pos(WideArgs a)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t win = blockIdx.x * ISLAND_WARPS + (threadIdx.x >> 5);
	if (win * 32u >= a.cnt[WC_ISLCUR]) return;
	const uint32_t mi = a.isl_man[win * 32u + lane];
	const bool have = mi != 0xFFFFFFFFu;
	const int colour = have ? -4 - a.man[mi].colour : -2;
	int ncol = colour + 1;
	for (int o = 16; o > 0; o >>= 1) ncol = max(ncol, __shfl_xor_sync(0xFFFFFFFFu, ncol, o));
	for (uint32_t it = 0; it < a.pos_steps; it++)
		for (int col = 0; col < ncol; col++)
		{
			if (colour == col) solve_position(a.man[mi], a.bodies);
			__syncwarp();
		}
}

// total order on manifolds for the colouring: priority, then identity
__device__ __forceinline__ bool outranks(const WideArgs &a, uint32_t x, uint32_t y)
{
	const uint32_t px = a.prio[x], py = a.prio[y];
	if (px != py) return px > py;
	const SMan &mx = a.man[x], &my = a.man[y];
	if (mx.a != my.a) return mx.a > my.a;
	if (mx.b != my.b) return mx.b > my.b;
	return a.ord[x] > a.ord[y];
}

// One Jones-Plassmann decision: the colour manifold `mi` may take now (it outranks all its uncoloured neighbours), or -1.
__device__ __forceinline__ int jp_decide(const WideArgs &a, uint32_t mi)
{
	const SMan &m = a.man[mi];
	if (m.colour != -1) return -1;
	unsigned long long used = 0ull;
	const uint32_t ends[2] = {m.a, m.b};
	for (int e = 0; e < 2; e++)
	{
		const uint32_t body = ends[e];
		if (body >= STATIC_BODY_BASE || !is_dynamic(a.bodies[body].flags)) continue;
		const uint32_t cntb = min(a.adj_n[body], (uint32_t)WIDE_MAXADJ);
		for (uint32_t j = 0; j < cntb; j++)
		{
			const uint32_t other = a.adj[body * WIDE_MAXADJ + j];
			if (other == mi) continue;
			const int oc = *(volatile int *)&a.man[other].colour;
			if (oc == -1)
			{
				if (outranks(a, other, mi)) return -1;
			}
			else if (oc >= 0)
				used |= 1ull << oc;
		}
	}
	int decision = __ffsll((long long)~used) - 1;
	if (decision < 0 || decision >= WIDE_MAXCOL)
	{
		decision = WIDE_MAXCOL - 1;
		atomicOr(&a.cnt[WC_ERR], (uint32_t)GPX_ERR_CONTACT_CONSTRAINTS_FULL);
	}
	return decision;
}

// Barrier of the cooperative kernels.  When the large islands are few, ONE block does the phased work and a block
// barrier is all it needs; the other blocks only help with the grid-wide integration pass of kw_solve.
struct PhaseSync
{
	cg::grid_group grid;
	bool single;
	__device__ __forceinline__ void operator()() const
	{
		if (single) __syncthreads();
		else grid.sync();
	}
};

// Jones-Plassmann: in each round every uncoloured manifold that outranks all its uncoloured neighbours takes the
// smallest colour none of its coloured neighbours has.  Decisions of a round only read colours of earlier rounds.
// Works on the manifolds of the large islands (big_list); the small ones were coloured inside kw_island.
__global__ void __launch_bounds__(256) kw_colour(WideArgs a)
{
	const uint32_t n = a.cnt[WC_NBIG];
	PhaseSync bar{cg::this_grid(), n <= SINGLE_BLOCK_MAX};
	if (bar.single && blockIdx.x != 0) return;
	const uint32_t tid = bar.single ? threadIdx.x : blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t stride = bar.single ? blockDim.x : gridDim.x * blockDim.x;
	for (int round = 0; round < 4096; round++)
	{
		// everyone reads the counter between the barrier that ended the previous round and the next one, and nobody
		// changes it before that next barrier, so all threads take the same branch
		if (*(volatile uint32_t *)&a.cnt[WC_UNCOLOURED] == 0) break;
		for (uint32_t k = tid; k < n; k += stride)
		{
			const uint32_t mi = a.big_list[k];
			const int decision = jp_decide(a, mi);
			a.pending[mi] = decision;
		}
		bar();
		uint32_t done = 0;
		for (uint32_t k = tid; k < n; k += stride)
		{
			const uint32_t mi = a.big_list[k];
			const int d = a.pending[mi];
			if (d >= 0)
			{
				a.man[mi].colour = d;
				atomicAdd(&a.cnt[WC_COLCNT + d], 1u);
				done++;
			}
		}
		if (done) atomicSub(&a.cnt[WC_UNCOLOURED], done);
		__threadfence();
		bar();
	}
	bar();
	// per-colour lists
	if (tid == 0)
	{
		uint32_t off = 0, ncol = 0;
		for (int c = 0; c < WIDE_MAXCOL; c++)
		{
			a.cnt[WC_COLOFF + c] = off;
			a.cnt[WC_COLCUR + c] = off;
			off += a.cnt[WC_COLCNT + c];
			if (a.cnt[WC_COLCNT + c]) ncol = (uint32_t)c + 1u;
		}
		a.cnt[WC_COLOFF + WIDE_MAXCOL] = off;
		a.cnt[WC_NCOL] = ncol;
		a.cnt[WC_NACTIVE] = off;
		__threadfence();
	}
	bar();
	for (uint32_t k = tid; k < n; k += stride)
	{
		const uint32_t mi = a.big_list[k];
		const int c = a.man[mi].colour;
		if (c >= 0) a.col_list[atomicAdd(&a.cnt[WC_COLCUR + c], 1u)] = mi;
	}
}

// ---------------------------------------------------------------------------------------------------- solve

// One colour phase over the solver records [lo, hi): each warp stages its 32 consecutive records in shared memory with
// coalesced 16-byte copies, solves from there (warm start or one velocity iteration) and writes back only the impulses.
__device__ __forceinline__ void solve_colour_range(const WideArgs &a, SolveRec *recs, uint32_t lo, uint32_t hi, uint32_t tid, uint32_t stride,
												   SolveRec *stage, uint32_t lane, uint32_t it)
{
	for (uint32_t k0 = lo + (tid & ~31u); k0 < hi; k0 += stride)
	{
		const uint32_t cnt32 = min(32u, hi - k0);
		const float4 *src = reinterpret_cast<const float4 *>(recs + k0);
		float4 *dst = reinterpret_cast<float4 *>(stage + (threadIdx.x & ~31u));
		const uint32_t n16 = cnt32 * (uint32_t)(sizeof(SolveRec) / 16);
		for (uint32_t q = lane; q < n16; q += 32u) dst[q] = __ldcg(&src[q]);
		__syncwarp();
		if (lane < cnt32)
		{
			SolveRec &r = stage[threadIdx.x];
			const Con c = r.con;
			Vel u;
			load_vel(c, a.bodies, u);
			if (it == 0)
				warm_start(c, r.rows, r.rows.f, u);
			else
				solve_velocity(c, r.rows, r.rows.f, u, it - 1u);
			store_vel(c, a.bodies, u);
			SolveRec &g = recs[k0 + lane];
#pragma unroll
			for (int p = 0; p < 4; p++) g.rows.ln[p] = r.rows.ln[p];
#pragma unroll
			for (int p = 0; p < 3; p++) g.rows.cf[p] = r.rows.cf[p];
		}
		__syncwarp();
	}
}


__global__ void __launch_bounds__(256) kw_solve(WideArgs a)
{
	cg::grid_group grid = cg::this_grid();
	extern __shared__ __align__(16) unsigned char stage_raw[];
	SolveRec *stage = reinterpret_cast<SolveRec *>(stage_raw);
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t nact = a.cnt[WC_NACTIVE];
	PhaseSync bar{grid, nact <= SINGLE_BLOCK_MAX};
	const bool phased = !bar.single || blockIdx.x == 0;  // this block takes part in the colour phases
	const uint32_t tid = bar.single ? threadIdx.x : blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t stride = bar.single ? blockDim.x : gridDim.x * blockDim.x;
	const int ncol = (int)a.cnt[WC_NCOL];
	const float h = a.h;
	if (phased)
	{
		// set-up: lever arms, effective masses, bias -> one solver record per active manifold, in colour order
		for (uint32_t k = tid; k < nact; k += stride)
		{
			const uint32_t mi = a.col_list[k];
			SMan &m = a.man[mi];
			SolveRec &r = a.recs[k];
			Con c;
			build_rows(c, r.rows, m, a.bodies, h);
			r.con = c;
			r.mi = mi;
		}
		bar();
		// it == 0: warm start; then the velocity iterations.  Colour by colour: no two manifolds of a colour share a
		// dynamic body, so every body is written by at most one thread per phase.
		for (uint32_t it = 0; it <= a.vel_steps; it++)
			for (int col = 0; col < ncol; col++)
			{
				const uint32_t lo = a.cnt[WC_COLOFF + col], hi = a.cnt[WC_COLOFF + col + 1];
				solve_colour_range(a, a.recs, lo, hi, tid, stride, stage, lane, it);
				bar();
			}
		// accumulated impulses back into the manifolds (next sub-step's warm start reads them there)
		for (uint32_t k = tid; k < nact; k += stride)
		{
			const SolveRec &r = a.recs[k];
			save_impulses(a.man[r.mi], r.rows);
		}
	}
	grid.sync();
	// integrate: everything that moves and was not integrated inside a small island — the large islands, bodies
	// without contacts, kinematic bodies.  After every set-up (kw_island ran earlier), before every position row.
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.nb; i += gridDim.x * blockDim.x)
	{
		SBody &b = a.bodies[i];
		if (!(b.flags & BF_ALIVE) || ((b.flags >> BF_MOTION_SHIFT) & 3u) == GPX_MOTION_STATIC || (b.flags & BF_ASLEEP)) continue;
		if (a.root_of[i] & ROOT_SMALL) continue;  // integrated by kw_island
		b.x = b.x + (b.v * h);
		b.q = qstep(b.q, b.w * h);
	}
	grid.sync();
	if (!phased) return;
	// position iterations
	for (uint32_t it = 0; it < a.pos_steps; it++)
		for (int col = 0; col < ncol; col++)
		{
			const uint32_t lo = a.cnt[WC_COLOFF + col], hi = a.cnt[WC_COLOFF + col + 1];
			for (uint32_t k = lo + tid; k < hi; k += stride) solve_position(a.man[a.col_list[k]], a.bodies);
			bar();
		}
}

// ---- medium islands (33 .. MEDIUM_MAX manifolds: a pile of a few dozen boxes): ONE block per island runs what kw_colour
// and kw_solve run grid-wide — the same Jones-Plassmann decisions, the same records and colour phases — with block
// barriers.  An island owns the range [base, base + count) of big_list / col_list / recs, counted from the top of those
// arrays (the cooperative kernels fill them from the bottom; every manifold is in exactly one class).
__global__ void __launch_bounds__(MEDIUM_T) kw_island_block(WideArgs a)
{
	extern __shared__ __align__(16) unsigned char stage_raw[];
	SolveRec *stage = reinterpret_cast<SolveRec *>(stage_raw);
	__shared__ uint32_t s_colcnt[WIDE_MAXCOL], s_coloff[WIDE_MAXCOL + 1], s_cur[WIDE_MAXCOL], s_uncol, s_ncol;
	const uint32_t tid = threadIdx.x, lane = threadIdx.x & 31u;
	const uint32_t n_med = min(a.cnt[WC_NMED], a.med_slots);
	for (uint32_t isl = blockIdx.x; isl < n_med; isl += gridDim.x)
	{
	const uint32_t root = a.med_list[isl];
	const uint32_t count = a.isl_cnt[root];
	const uint32_t base = a.cap_m - a.isl_off[root] - count;
	__syncthreads();  // the previous island's shared counters are done with
	for (uint32_t c = tid; c < (uint32_t)WIDE_MAXCOL; c += MEDIUM_T) s_colcnt[c] = 0;
	if (tid == 0) s_uncol = count;
	__syncthreads();
	// --- colouring
	for (int round = 0; round < 4096; round++)
	{
		if (*(volatile uint32_t *)&s_uncol == 0) break;
		for (uint32_t k = tid; k < count; k += MEDIUM_T)
		{
			const uint32_t mi = a.big_list[base + k];
			a.pending[mi] = jp_decide(a, mi);
		}
		__syncthreads();
		uint32_t done = 0;
		for (uint32_t k = tid; k < count; k += MEDIUM_T)
		{
			const uint32_t mi = a.big_list[base + k];
			const int d = a.pending[mi];
			if (d >= 0)
			{
				a.man[mi].colour = d;
				atomicAdd(&s_colcnt[d], 1u);
				done++;
			}
		}
		if (done) atomicSub(&s_uncol, done);
		__threadfence_block();
		__syncthreads();
	}
	if (tid == 0)
	{
		uint32_t off = 0, ncol = 0;
		for (int c = 0; c < WIDE_MAXCOL; c++)
		{
			s_coloff[c] = off;
			s_cur[c] = off;
			off += s_colcnt[c];
			if (s_colcnt[c]) ncol = (uint32_t)c + 1u;
		}
		s_coloff[WIDE_MAXCOL] = off;
		s_ncol = ncol;
	}
	__syncthreads();
	for (uint32_t k = tid; k < count; k += MEDIUM_T)
	{
		const uint32_t mi = a.big_list[base + k];
		a.col_list[base + atomicAdd(&s_cur[a.man[mi].colour], 1u)] = mi;
	}
	for (uint32_t c = tid; c <= (uint32_t)WIDE_MAXCOL; c += MEDIUM_T) a.med_coloff[isl * (WIDE_MAXCOL + 1) + c] = s_coloff[c];
	__threadfence_block();
	__syncthreads();
	// --- set-up: one solver record per manifold, in colour order
	const float h = a.h;
	const int ncol = (int)s_ncol;
	for (uint32_t k = tid; k < count; k += MEDIUM_T)
	{
		const uint32_t mi = a.col_list[base + k];
		SMan &m = a.man[mi];
		SolveRec &r = a.recs[base + k];
		Con c;
		build_rows(c, r.rows, m, a.bodies, h);
		r.con = c;
		r.mi = mi;
	}
	__threadfence_block();
	__syncthreads();
	// --- warm start, then the velocity iterations, colour by colour
	for (uint32_t it = 0; it <= a.vel_steps; it++)
		for (int col = 0; col < ncol; col++)
		{
			solve_colour_range(a, a.recs + base, s_coloff[col], s_coloff[col + 1], tid, MEDIUM_T, stage, lane, it);
			__threadfence_block();
			__syncthreads();
		}
	// --- impulses back into the manifolds; integrate the island's dynamic bodies (each by the thread that holds the
	// first manifold of its incidence list)
	for (uint32_t k = tid; k < count; k += MEDIUM_T)
	{
		const SolveRec &r = a.recs[base + k];
		SMan &m = a.man[r.mi];
		save_impulses(m, r.rows);
		const uint32_t ends[2] = {m.a, m.b};
		for (int e = 0; e < 2; e++)
		{
			const uint32_t body = ends[e];
			if (body >= STATIC_BODY_BASE || !is_dynamic(a.bodies[body].flags) || a.adj[body * WIDE_MAXADJ] != r.mi) continue;
			SBody &b = a.bodies[body];
			b.x = b.x + (b.v * h);
			b.q = qstep(b.q, b.w * h);
		}
	}
	}
}

// position iterations of the medium islands, after the kinematic bodies have moved (kw_solve integrates them)
__global__ void __launch_bounds__(MEDIUM_T) kw_island_block_pos(WideArgs a)
{
	const uint32_t n_med = min(a.cnt[WC_NMED], a.med_slots);
	for (uint32_t isl = blockIdx.x; isl < n_med; isl += gridDim.x)
	{
	const uint32_t root = a.med_list[isl];
	const uint32_t count = a.isl_cnt[root];
	const uint32_t base = a.cap_m - a.isl_off[root] - count;
	const uint32_t *coloff = a.med_coloff + isl * (WIDE_MAXCOL + 1);
	int ncol = 0;
	for (int c = 0; c < WIDE_MAXCOL; c++)
		if (coloff[c + 1] > coloff[c]) ncol = c + 1;
	for (uint32_t it = 0; it < a.pos_steps; it++)
		for (int col = 0; col < ncol; col++)
		{
			for (uint32_t k = coloff[col] + threadIdx.x; k < coloff[col + 1]; k += MEDIUM_T) solve_position(a.man[a.col_list[base + k]], a.bodies);
			__threadfence_block();
			__syncthreads();
		}
	}
}

__global__ void __launch_bounds__(WT) kw_finish(WideArgs a)
{
	const uint32_t t = blockIdx.x * WT + threadIdx.x;
	const uint32_t n = min(a.cnt[WC_NMAN], a.cap_m);
	if (t < n)
	{
		const SMan &m = a.man[t];
		if (m.np > 0 && m.colour != SENSOR_COLOUR)
		{
			const unsigned long long key = man_key(m.a, m.b, a.ord[t]);
			uint32_t s = key_slot(key, a.hmask);
			for (uint32_t probe = 0; probe <= a.hmask; probe++)
			{
				const unsigned long long old = atomicCAS(&a.hkeys[s], 0ull, key);
				if (old == 0ull || old == key)
				{
					a.hvals[s] = t;
					break;
				}
				s = (s + 1u) & a.hmask;
			}
		}
	}
	if (t < a.nb)
	{
		SBody &b = a.bodies[t];
		bool woke = false;
		if (b.flags & BF_WAKE_MARK)
		{
			b.flags &= ~(BF_WAKE_MARK | BF_ASLEEP);
			a.bs.sleep_t[t] = -1.0f;
			woke = true;
		}
		if (a.last && (b.flags & BF_ALIVE))
		{
			a.bs.pos[t] = F4(b.x, 0.0f);
			a.bs.quat[t] = make_float4(b.q.x, b.q.y, b.q.z, b.q.w);
			a.bs.lin[t] = F4(b.v, 0.0f);
			a.bs.ang[t] = F4(b.w, 0.0f);
		}
		if (woke) a.bs.flags[t] = b.flags & ~BF_KIN_MOVING;
	}
}


// ---- sleeping, once per tick after the last sub-step (see k_sleep in gpx_tick.cu for the rule): islands of the awake
// dynamic bodies from the last sub-step's contacts, the per-body test, then whole islands of candidates go to sleep
__global__ void __launch_bounds__(WT) kw_sleep_init(WideArgs a, unsigned char *can)
{
	const uint32_t i = blockIdx.x * WT + threadIdx.x;
	if (i >= a.nb) return;
	a.parent[i] = i;
	can[i] = 1;
}

__global__ void __launch_bounds__(WT) kw_sleep_link(WideArgs a)
{
	const uint32_t mi = blockIdx.x * WT + threadIdx.x;
	if (mi >= min(a.cnt[WC_NMAN], a.cap_m)) return;
	const SMan &m = a.man[mi];
	if (m.np == 0 || m.b >= STATIC_BODY_BASE || m.colour == SENSOR_COLOUR) return;
	if (is_dynamic(a.bodies[m.a].flags) && is_dynamic(a.bodies[m.b].flags)) isl_unite(a.parent, m.a, m.b);
}

__global__ void __launch_bounds__(WT) kw_sleep_test(WideArgs a, unsigned char *can, float dt)
{
	const uint32_t i = blockIdx.x * WT + threadIdx.x;
	if (i >= a.nb) return;
	const uint32_t f = a.bodies[i].flags;
	if (!(f & BF_ALIVE) || !is_dynamic(f)) return;
	if (!sleep_test_body(a.bs, i, f, dt)) can[isl_find(a.parent, i)] = 0;
}

__global__ void __launch_bounds__(WT) kw_sleep_apply(WideArgs a, const unsigned char *can)
{
	const uint32_t i = blockIdx.x * WT + threadIdx.x;
	if (i >= a.nb) return;
	const uint32_t f = a.bodies[i].flags;
	if (!(f & BF_ALIVE) || !is_dynamic(f) || !can[isl_find(a.parent, i)]) return;
	a.bs.flags[i] = (f | BF_ASLEEP) & ~(BF_KIN_MOVING | BF_WAKE_MARK);
	a.bs.lin[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	a.bs.ang[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}


// ---- contact events of a wide world (gpx_events_enable), once per tick: the keys (a << 32 | b) of every touching pair
// of the last sub-step — solver manifolds, sensor overlaps, the character's contacts — are sorted and made unique, then
// compared with the previous tick's list by binary search.  Output order as in the ensemble kernel: added and
// persisted pairs sorted by key, then the removed ones sorted by key.
constexpr uint32_t EV_CHARACTER = 0x200000u;  // + world: the character's id inside the event keys (bodies < 2^20 < this < static ids)
struct EvArgs
{
	const SMan *man;
	const uint32_t *cnt;
	uint32_t cap_m, n_ev;
	uint32_t worlds, cap;  // an ensemble of wide worlds: body indices are world * cap + local
	const unsigned long long *ch_keys;
	const uint32_t *ch_nkeys;
	unsigned long long *keys, *cur, *prev;
	uint32_t *flag, *pos, *ncur, *nprev, *count;
	uint4 *out;
};

__global__ void __launch_bounds__(WT) kw_ev_keys(EvArgs e)
{
	const uint32_t i = blockIdx.x * WT + threadIdx.x;
	if (i >= e.n_ev) return;
	unsigned long long key = ~0ull;
	if (i < e.cap_m)
	{
		if (i < min(e.cnt[WC_NMAN], e.cap_m) && e.man[i].np > 0) key = ((unsigned long long)e.man[i].a << 32) | e.man[i].b;
	}
	else if (e.ch_keys)
	{
		// the characters' contacts, one block of 64 per world: (other << 32 | character).  With the body index made global
		// and the character's pseudo id replaced by EV_CHARACTER + world — still between the bodies' ids and the static
		// ones, so the order within a world is the one a single world has — the keys of different worlds stay apart.
		const uint32_t j = i - e.cap_m, wi = j / CHARACTER_MAX_CONTACTS, slot = j % CHARACTER_MAX_CONTACTS;
		if (wi < e.worlds && slot < min(e.ch_nkeys[wi], (uint32_t)CHARACTER_MAX_CONTACTS))
		{
			const unsigned long long k = e.ch_keys[(size_t)wi * CHARACTER_MAX_CONTACTS + slot];
			// (body, character) for bodies, (character, static id) for the map's meshes
			const uint32_t ka = (uint32_t)(k >> 32), kb = (uint32_t)(k & 0xFFFFFFFFull);
			if (ka == CHARACTER_BODY_ID)
				key = ((unsigned long long)(EV_CHARACTER + wi) << 32) | kb;
			else
				key = ((unsigned long long)(ka + wi * e.cap) << 32) | (EV_CHARACTER + wi);
		}
	}
	e.keys[i] = key;
}

// a key of the lists above as the record the host reads: ids local to the world, the world in .w
__device__ __forceinline__ uint4 ev_record(const EvArgs &e, unsigned long long k, uint32_t kind)
{
	uint32_t a = (uint32_t)(k >> 32), b = (uint32_t)(k & 0xFFFFFFFFull), world = 0;
	if (a >= EV_CHARACTER && a < STATIC_BODY_BASE)
	{
		world = a - EV_CHARACTER;  // (character, static id)
		a = CHARACTER_BODY_ID;
	}
	else
	{
		world = a / e.cap;
		a %= e.cap;
		if (b >= EV_CHARACTER && b < STATIC_BODY_BASE) b = CHARACTER_BODY_ID;  // (body, character)
		else if (b < STATIC_BODY_BASE) b %= e.cap;
	}
	return make_uint4(a, b, kind, world);
}

__global__ void __launch_bounds__(WT) kw_ev_unique(EvArgs e)
{
	const uint32_t i = blockIdx.x * WT + threadIdx.x;
	if (i >= e.n_ev) return;
	const unsigned long long k = e.keys[i];
	const uint32_t f = (k != ~0ull && (i == 0 || e.keys[i - 1] != k)) ? 1u : 0u;
	e.flag[i] = f;
	e.pos[i] = f;
}

__global__ void __launch_bounds__(WT) kw_ev_compact(EvArgs e)
{
	const uint32_t i = blockIdx.x * WT + threadIdx.x;
	if (i >= e.n_ev) return;
	if (e.flag[i]) e.cur[e.pos[i]] = e.keys[i];
	if (i == e.n_ev - 1) e.ncur[0] = e.pos[i] + e.flag[i];
}

__device__ __forceinline__ bool ev_contains(const unsigned long long *s, uint32_t n, unsigned long long k)
{
	uint32_t lo = 0, hi = n;
	while (lo < hi)
	{
		const uint32_t mid = (lo + hi) >> 1;
		if (s[mid] < k) lo = mid + 1;
		else hi = mid;
	}
	return lo < n && s[lo] == k;
}

// pass 0: added / persisted records, and which previous keys are gone (flag for the second compaction)
__global__ void __launch_bounds__(WT) kw_ev_diff(EvArgs e)
{
	const uint32_t i = blockIdx.x * WT + threadIdx.x;
	if (i >= e.n_ev) return;
	const uint32_t ncur = e.ncur[0], nprev = e.nprev[0];
	if (i < ncur)
	{
		const unsigned long long k = e.cur[i];
		e.out[i] = ev_record(e, k, ev_contains(e.prev, nprev, k) ? 2u : 1u);
	}
	const uint32_t gone = (i < nprev && !ev_contains(e.cur, ncur, e.prev[i])) ? 1u : 0u;
	e.flag[i] = gone;
	e.pos[i] = gone;
}

__global__ void __launch_bounds__(WT) kw_ev_removed(EvArgs e)
{
	const uint32_t i = blockIdx.x * WT + threadIdx.x;
	if (i >= e.n_ev) return;
	const uint32_t ncur = e.ncur[0];
	if (e.flag[i])
	{
		const unsigned long long k = e.prev[i];
		e.out[ncur + e.pos[i]] = ev_record(e, k, 3u);
	}
	if (i == e.n_ev - 1) e.count[0] = ncur + e.pos[i] + e.flag[i];
}

// the current list becomes the previous one (after kw_ev_removed has read the old one)
__global__ void __launch_bounds__(WT) kw_ev_roll(EvArgs e)
{
	const uint32_t i = blockIdx.x * WT + threadIdx.x;
	if (i >= e.n_ev) return;
	const uint32_t ncur = e.ncur[0];
	if (i < ncur) e.prev[i] = e.cur[i];
	if (i == 0) e.nprev[0] = ncur;
}

// ---------------------------------------------------------------------------------------------------- host

template <typename T>
static bool walloc(T **p, size_t n)
{
	if (cudaMalloc(p, sizeof(T) * n) != cudaSuccess) return false;
	return cudaMemset(*p, 0, sizeof(T) * n) == cudaSuccess;
}

int wide_create(gpx_world *w)
{
	WideDevice *d = new WideDevice();
	w->wide = d;
	d->nb = w->W * w->cap;  // all worlds' bodies in one array, world-major (the layout of the body store)
	d->n_pad = next_pow2(d->nb);
	d->cap_m = w->W * w->cap_m;
	d->hsize = next_pow2(2u * d->cap_m);
	// worst case of the window packing: every window half empty, plus one open window per kw_isl_place block
	d->isl_slots = 2u * d->cap_m + 32u * ((d->nb + WT - 1) / WT);
	d->med_slots = d->cap_m / (ISLAND_MAX + 1u) + 1u;  // a medium island has at least 33 manifolds
	bool ok = walloc(&d->bodies, d->nb) && walloc(&d->keys, d->n_pad) && walloc(&d->boxlo, d->n_pad) &&
			  walloc(&d->boxhi, d->n_pad) && walloc(&d->man[0], d->cap_m) && walloc(&d->man[1], d->cap_m) &&
			  walloc(&d->ord[0], d->cap_m) && walloc(&d->ord[1], d->cap_m) && walloc(&d->recs, d->cap_m) &&
			  walloc(&d->counters, (size_t)WC_COUNT) && walloc(&d->hkeys, d->hsize) && walloc(&d->hvals, d->hsize) &&
			  walloc(&d->adj, (size_t)d->nb * WIDE_MAXADJ) && walloc(&d->adj_n, d->nb) && walloc(&d->prio, d->cap_m) &&
			  walloc(&d->pending, d->cap_m) && walloc(&d->col_list, d->cap_m) && walloc(&d->parent, d->nb) &&
			  walloc(&d->root_of, d->nb) && walloc(&d->isl_cnt, d->nb) && walloc(&d->isl_off, d->nb) && walloc(&d->isl_cur, d->nb) &&
			  walloc(&d->isl_man, (size_t)d->isl_slots) && walloc(&d->big_list, d->cap_m) && walloc(&d->med_list, (size_t)d->med_slots) &&
			  walloc(&d->med_coloff, (size_t)d->med_slots * (WIDE_MAXCOL + 1)) &&
			  walloc(&d->keys_tmp, d->n_pad) && walloc(&d->sort_hist, (size_t)256 * (d->n_pad / 1024u + 1u)) &&
			  walloc(&d->can_sleep, d->nb) && walloc(&d->order, d->n_pad) && walloc(&d->bkeys, d->n_pad) &&
			  walloc(&d->sort_fallbacks, 1);
	if (ok) kw_iota<<<(d->n_pad + 255u) / 256u, 256>>>(d->order, d->n_pad);
	if (!ok)
	{
		set_error("wide_create", cudaGetLastError());
		return GPX_ERR_CUDA;
	}
	int sms = 148, per_sm = 1;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
	GPX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kw_colour, 256, 0));
	d->coop_grid_colour = sms * (per_sm > 0 ? per_sm : 1);
	GPX_CUDA(cudaFuncSetAttribute(kw_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(256 * sizeof(SolveRec))));
	GPX_CUDA(cudaFuncSetAttribute(kw_island_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MEDIUM_T * sizeof(SolveRec))));
	GPX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kw_solve, 256, 256 * sizeof(SolveRec)));
	d->coop_grid_solve = sms * (per_sm > 0 ? per_sm : 1);
	return GPX_OK;
}

void wide_destroy(gpx_world *w)
{
	WideDevice *d = w->wide;
	if (!d) return;
	cudaFree(d->bodies); cudaFree(d->keys); cudaFree(d->boxlo); cudaFree(d->boxhi); cudaFree(d->man[0]); cudaFree(d->man[1]);
	if (d->graph) cudaGraphExecDestroy(d->graph);
	cudaFree(d->order); cudaFree(d->bkeys); cudaFree(d->sort_fallbacks);
	cudaFree(d->ord[0]); cudaFree(d->ord[1]); cudaFree(d->recs); cudaFree(d->counters); cudaFree(d->hkeys); cudaFree(d->hvals);
	cudaFree(d->adj); cudaFree(d->adj_n); cudaFree(d->prio); cudaFree(d->pending); cudaFree(d->col_list);
	cudaFree(d->parent); cudaFree(d->root_of); cudaFree(d->isl_cnt); cudaFree(d->isl_off); cudaFree(d->isl_cur);
	cudaFree(d->isl_man); cudaFree(d->big_list); cudaFree(d->med_list); cudaFree(d->med_coloff); cudaFree(d->keys_tmp); cudaFree(d->sort_hist); cudaFree(d->can_sleep);
	cudaFree(d->ev_keys); cudaFree(d->ev_tmp); cudaFree(d->ev_cur); cudaFree(d->ev_hist); cudaFree(d->ev_flag); cudaFree(d->ev_pos);
	cudaFree(d->ev_ncur);
	delete d;
	w->wide = nullptr;
}

uint32_t wide_event_capacity(const gpx_world *w) { return w->wide ? w->wide->n_ev : 0u; }

int wide_events_enable(gpx_world *w, bool enable)
{
	WideDevice *d = w->wide;
	if (enable && !d->ev_keys)
	{
		d->n_ev = ((d->cap_m + w->W * CHARACTER_MAX_CONTACTS + 1023u) / 1024u) * 1024u;
		const bool ok = walloc(&d->ev_keys, d->n_ev) && walloc(&d->ev_tmp, d->n_ev) && walloc(&d->ev_cur, d->n_ev) &&
						walloc(&d->ev_hist, (size_t)256 * (d->n_ev / 1024u + 1u)) && walloc(&d->ev_flag, d->n_ev) &&
						walloc(&d->ev_pos, d->n_ev) && walloc(&d->ev_ncur, 1) && walloc(&w->d_ev_prev, d->n_ev) &&
						walloc(&w->d_ev_nprev, 1) && walloc(&w->d_ev_count, 1) && walloc(&w->d_ev_out, 2 * (size_t)d->n_ev);
		if (!ok)
		{
			set_error("wide_events_enable", cudaGetLastError());
			return GPX_ERR_CUDA;
		}
	}
	else if (!enable && d->ev_keys)
	{
		cudaFree(d->ev_keys); cudaFree(d->ev_tmp); cudaFree(d->ev_cur); cudaFree(d->ev_hist); cudaFree(d->ev_flag);
		cudaFree(d->ev_pos); cudaFree(d->ev_ncur);
		cudaFree(w->d_ev_prev); cudaFree(w->d_ev_nprev); cudaFree(w->d_ev_count); cudaFree(w->d_ev_out);
		d->ev_keys = d->ev_tmp = d->ev_cur = nullptr;
		d->ev_hist = d->ev_flag = d->ev_pos = d->ev_ncur = nullptr;
		w->d_ev_prev = nullptr;
		w->d_ev_nprev = w->d_ev_count = nullptr;
		w->d_ev_out = nullptr;
	}
	return GPX_OK;
}

static int wide_events(gpx_world *w, const SMan *man)
{
	WideDevice *d = w->wide;
	cudaStream_t st = w->stream;
	EvArgs e;
	e.man = man;
	e.cnt = d->counters;
	e.cap_m = d->cap_m;
	e.n_ev = d->n_ev;
	e.worlds = w->W;
	e.cap = w->cap;
	e.ch_keys = w->d_ch_keys;
	e.ch_nkeys = w->d_ch_nkeys;
	e.keys = d->ev_keys;
	e.cur = d->ev_cur;
	e.prev = w->d_ev_prev;
	e.flag = d->ev_flag;
	e.pos = d->ev_pos;
	e.ncur = d->ev_ncur;
	e.nprev = w->d_ev_nprev;
	e.count = w->d_ev_count;
	e.out = w->d_ev_out;
	const uint32_t g = d->n_ev / WT;
	kw_ev_keys<<<g, WT, 0, st>>>(e);
	count_launch();
	radix_sort_u64(d->ev_keys, d->ev_tmp, d->ev_hist, d->n_ev, 0u, st);
	kw_ev_unique<<<g, WT, 0, st>>>(e);
	exclusive_scan_u32(d->ev_pos, d->n_ev, st);
	kw_ev_compact<<<g, WT, 0, st>>>(e);
	kw_ev_diff<<<g, WT, 0, st>>>(e);
	exclusive_scan_u32(d->ev_pos, d->n_ev, st);
	kw_ev_removed<<<g, WT, 0, st>>>(e);
	kw_ev_roll<<<g, WT, 0, st>>>(e);
	count_launch(5);
	GPX_CUDA(cudaGetLastError());
	return GPX_OK;
}

static int coop_launch(const void *fn, int grid, WideArgs &a, cudaStream_t st, size_t smem = 0)
{
	void *params[] = {&a};
	GPX_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(256), params, smem, st));
	count_launch();
	return GPX_OK;
}

int wide_counters(gpx_world *w, uint32_t *out8)
{
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	GPX_CUDA(cudaMemcpy(out8, w->wide->counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	GPX_CUDA(cudaMemcpy(out8 + WC_NPREV, w->wide->counters + WC_NISL, sizeof(uint32_t), cudaMemcpyDeviceToHost));
	GPX_CUDA(cudaMemcpy(out8 + WC_UNCOLOURED, w->wide->counters + WC_NMED, sizeof(uint32_t), cudaMemcpyDeviceToHost));
	GPX_CUDA(cudaMemcpy(out8 + WC_MAXEXT_X, w->wide->sort_fallbacks, sizeof(uint32_t), cudaMemcpyDeviceToHost));  // sub-steps that needed the radix passes
	return GPX_OK;
}

__global__ void kw_or_word(uint32_t *dst, const uint32_t *src) { *dst |= *src; }

// gpx_read_stats: the manifolds of the last sub-step that hold contact points, counted into their world's entry
__global__ void kw_count_manifolds(const SMan *man, const uint32_t *cnt, uint32_t cap_m, uint32_t cap, gpx_world_stats *stats)
{
	const uint32_t n = min(cnt[WC_NMAN], cap_m);
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
		if (man[i].np > 0) atomicAdd(&stats[man[i].a / cap].manifolds, 1u);
}

int wide_stats(gpx_world *w)
{
	const WideDevice *d = w->wide;
	kw_count_manifolds<<<296, 256, 0, w->stream>>>(d->man[d->cur ^ 1], d->counters, d->cap_m, w->cap, w->d_stats);
	count_launch();
	GPX_CUDA(cudaGetLastError());
	return GPX_OK;
}

static int record_wide_tick(gpx_world *w, WideArgs &a, float dt, int substeps);

int launch_wide_tick(gpx_world *w, float dt, int substeps)
{
	WideDevice *d = w->wide;
	cudaStream_t st = w->stream;
	if (substeps < 1) substeps = 1;
	WideArgs a;
	memset(&a, 0, sizeof(a));
	a.bs = w->bs;
	a.bodies = d->bodies;
	a.keys = d->keys;
	a.order = d->order;
	a.sort_fallbacks = d->sort_fallbacks;
	a.bkeys = d->bkeys;
	a.boxlo = d->boxlo;
	a.boxhi = d->boxhi;
	a.recs = d->recs;
	a.cnt = d->counters;
	a.hkeys = d->hkeys;
	a.hvals = d->hvals;
	a.adj = d->adj;
	a.adj_n = d->adj_n;
	a.prio = d->prio;
	a.pending = d->pending;
	a.col_list = d->col_list;
	a.parent = d->parent;
	a.root_of = d->root_of;
	a.isl_cnt = d->isl_cnt;
	a.isl_off = d->isl_off;
	a.isl_cur = d->isl_cur;
	a.isl_man = d->isl_man;
	a.big_list = d->big_list;
	a.med_list = d->med_list;
	a.med_coloff = d->med_coloff;
	a.med_slots = d->med_slots;
	a.cand = w->d_cand;
	a.sv.nodes = w->sd.nodes;
	a.sv.tris = w->sd.tri;
	a.sv.n_nodes = w->sd.n_nodes;
	a.nb = d->nb;
	a.n_pad = d->n_pad;
	a.cap_m = d->cap_m;
	a.hmask = d->hsize - 1u;
	a.isl_max = getenv("GPX_WIDE_NO_ISLANDS") ? 0u : ISLAND_MAX;
	a.med_max = (getenv("GPX_WIDE_NO_ISLANDS") || getenv("GPX_WIDE_NO_BLOCKS")) ? 0u : MEDIUM_MAX;
	a.cap = w->cap;
	a.rows_per_world = (uint32_t)WIDE_ROWS / (w->W < (uint32_t)WIDE_ROWS ? w->W : (uint32_t)WIDE_ROWS);
	a.vel_steps = w->cfg.velocity_steps ? w->cfg.velocity_steps : 10u;
	a.pos_steps = w->cfg.position_steps ? w->cfg.position_steps : 2u;
	a.gx = w->cfg.gravity[0];
	a.gy = w->cfg.gravity[1];
	a.gz = w->cfg.gravity[2];
	a.h = dt / (float)substeps;
	static const bool no_graph = getenv("GPX_WIDE_NO_GRAPH") != nullptr;
	if (no_graph || d->graph_off) return record_wide_tick(w, a, dt, substeps);
	// signature of the launch sequence: every kernel argument and everything that selects launches
	struct { int substeps, cur; float dt; bool sleep; const void *ev, *ch_keys, *ch_nkeys, *ev_prev, *ev_nprev, *ev_count; } extra;
	memset(&extra, 0, sizeof(extra));
	extra.substeps = substeps;
	extra.cur = d->cur;
	extra.dt = dt;
	extra.sleep = w->sleep_enabled;
	extra.ev = w->d_ev_out;
	extra.ch_keys = w->d_ch_keys;
	extra.ch_nkeys = w->d_ch_nkeys;
	extra.ev_prev = w->d_ev_prev;
	extra.ev_nprev = w->d_ev_nprev;
	extra.ev_count = w->d_ev_count;
	unsigned char sig[sizeof(a) + sizeof(extra)];
	static_assert(sizeof(sig) <= sizeof(d->graph_sig), "graph signature buffer");
	memcpy(sig, &a, sizeof(a));
	memcpy(sig + sizeof(a), &extra, sizeof(extra));
	if (!d->graph || d->graph_sig_n != sizeof(sig) || memcmp(sig, d->graph_sig, sizeof(sig)) != 0)
	{
		if (d->graph) cudaGraphExecDestroy(d->graph);
		d->graph = nullptr;
		const int cur0 = d->cur;
		const uint64_t l0 = gpx_launch_count();
		cudaGraph_t g = nullptr;
		bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
		const int rc = ok ? record_wide_tick(w, a, dt, substeps) : GPX_ERR_CUDA;
		if (ok) ok = cudaStreamEndCapture(st, &g) == cudaSuccess && rc == GPX_OK && g;
		if (ok) ok = cudaGraphInstantiate(&d->graph, g, 0) == cudaSuccess;
		if (g) cudaGraphDestroy(g);
		d->graph_launches = gpx_launch_count() - l0;
		d->cur = cur0;
		if (!ok)
		{
			// no graphs for this world from here on: plain launches
			cudaGetLastError();
			d->graph = nullptr;
			d->graph_off = true;
			return record_wide_tick(w, a, dt, substeps);
		}
		memcpy(d->graph_sig, sig, sizeof(sig));
		d->graph_sig_n = sizeof(sig);
	}
	else
		count_launch(d->graph_launches);
	GPX_CUDA(cudaGraphLaunch(d->graph, st));
	d->cur ^= substeps & 1;
	return GPX_OK;
}

static int record_wide_tick(gpx_world *w, WideArgs &a, float dt, int substeps)
{
	WideDevice *d = w->wide;
	cudaStream_t st = w->stream;
	// every tick reports its own errors (Jolt's Update returns that update's result): the error word starts clean
	GPX_CUDA(cudaMemsetAsync(d->counters + WC_ERR, 0, sizeof(uint32_t), st));
	const uint32_t gb = (d->n_pad + WT - 1) / WT, gm = (d->cap_m + WT - 1) / WT;
	for (int sub = 0; sub < substeps; sub++)
	{
		a.first = sub == 0;
		a.last = sub == substeps - 1;
		a.man = d->man[d->cur];
		a.prev = d->man[d->cur ^ 1];
		a.ord = d->ord[d->cur];
		a.prev_ord = d->ord[d->cur ^ 1];
		// counters: everything but the error word and the previous count starts from zero
		GPX_CUDA(cudaMemsetAsync(d->counters + WC_NMAN, 0, sizeof(uint32_t), st));
		GPX_CUDA(cudaMemsetAsync(d->counters + WC_UNCOLOURED, 0, sizeof(uint32_t) * (WC_COUNT - WC_UNCOLOURED), st));
		kw_begin<<<gb, WT, 0, st>>>(a);
		kw_keys<<<gb, WT, 0, st>>>(a);
		count_launch(2);
		if (d->n_pad >= 4096u)
		{
			static const bool no_coherence = getenv("GPX_WIDE_NO_COHERENT_SORT") != nullptr;
			if (!no_coherence)
			{
				// the keys in the last order, two passes of tile sorts; the radix sort (all 64 bits: the input is in no
				// particular order below bit 20) runs only if that was not enough
				kw_tile_sort<<<d->n_pad / SORT_TILE, SORT_TILE / 2, 0, st>>>(d->keys, d->bkeys, d->order, d->n_pad, 0u);
				kw_tile_sort<<<d->n_pad / SORT_TILE - 1u, SORT_TILE / 2, 0, st>>>(d->keys, nullptr, nullptr, d->n_pad, SORT_TILE / 2);
				kw_sorted_check<<<gb, WT, 0, st>>>(a);
				count_launch(3);
				radix_sort_u64(d->keys, d->keys_tmp, d->sort_hist, d->n_pad, 0u, st, d->counters + WC_UNSORTED);
			}
			else
			{
				// keys arrive in body order with the body index in the low 20 bits: a stable sort of the bits above it is
				// the full sort
				GPX_CUDA(cudaMemcpyAsync(d->keys, d->bkeys, sizeof(unsigned long long) * d->n_pad, cudaMemcpyDeviceToDevice, st));
				radix_sort_u64(d->keys, d->keys_tmp, d->sort_hist, d->n_pad, 20u, st);
			}
		}
		else
		{
			GPX_CUDA(cudaMemcpyAsync(d->keys, d->bkeys, sizeof(unsigned long long) * d->n_pad, cudaMemcpyDeviceToDevice, st));
			bitonic_sort_u64(d->keys, d->n_pad, st);
		}
		kw_gather<<<gb, WT, 0, st>>>(a);
		kw_sweep<<<gb, WT, 0, st>>>(a);
		kw_pairs<<<(d->cap_m + NARROW_T - 1) / NARROW_T, NARROW_T, 0, st>>>(a);
		kw_static<<<(d->nb + NARROW_T - 1) / NARROW_T, NARROW_T, 0, st>>>(a);
		kw_link<<<gm, WT, 0, st>>>(a);
		GPX_CUDA(cudaMemsetAsync(d->isl_man, 0xFF, sizeof(uint32_t) * d->isl_slots, st));
		kw_isl_count<<<gm, WT, 0, st>>>(a);
		kw_isl_place<<<(d->nb + WT - 1) / WT, WT, 0, st>>>(a);
		kw_isl_fill<<<gm, WT, 0, st>>>(a);
		// one warp per 32-slot window; warps beyond the windows in use leave at once
		const uint32_t gi = (d->isl_slots / 32u + ISLAND_WARPS - 1) / ISLAND_WARPS;
		kw_island<<<gi, ISLAND_WARPS * 32, 0, st>>>(a);
		// a fixed grid of blocks walks the medium islands, one island per block at a time
		const uint32_t gmed = d->med_slots < (uint32_t)d->coop_grid_colour * 2u ? d->med_slots : (uint32_t)d->coop_grid_colour * 2u;
		kw_island_block<<<gmed, MEDIUM_T, MEDIUM_T * sizeof(SolveRec), st>>>(a);
		count_launch(10);
		int rc;
		if ((rc = coop_launch((const void *)kw_colour, d->coop_grid_colour, a, st)) != GPX_OK) return rc;
		if ((rc = coop_launch((const void *)kw_solve, d->coop_grid_solve, a, st, 256 * sizeof(SolveRec))) != GPX_OK) return rc;
		kw_island_pos<<<gi, ISLAND_WARPS * 32, 0, st>>>(a);
		kw_island_block_pos<<<gmed, MEDIUM_T, 0, st>>>(a);
		count_launch(2);
		// this sub-step's manifolds become the next one's warm-start table
		GPX_CUDA(cudaMemsetAsync(d->hkeys, 0, sizeof(unsigned long long) * d->hsize, st));
		kw_finish<<<max(gm, gb), WT, 0, st>>>(a);
		count_launch();
		d->cur ^= 1;
	}
	if (w->sleep_enabled)
	{
		// the manifolds of the last sub-step sit in the buffer `cur` just left
		a.man = d->man[d->cur ^ 1];
		const uint32_t gn = (d->nb + WT - 1) / WT;
		kw_sleep_init<<<gn, WT, 0, st>>>(a, d->can_sleep);
		kw_sleep_link<<<gm, WT, 0, st>>>(a);
		kw_sleep_test<<<gn, WT, 0, st>>>(a, d->can_sleep, dt);
		kw_sleep_apply<<<gn, WT, 0, st>>>(a, d->can_sleep);
		count_launch(4);
	}
	if (w->d_ev_out)
	{
		int rc = wide_events(w, d->man[d->cur ^ 1]);
		if (rc != GPX_OK) return rc;
	}
	// this tick's error word into the world's error slot (d_err[0]); d_err[1] keeps every error the world ever had (stats)
	GPX_CUDA(cudaMemcpyAsync(w->d_err, d->counters + WC_ERR, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
	kw_or_word<<<1, 1, 0, st>>>(w->d_err + 1, d->counters + WC_ERR);
	count_launch();
	GPX_CUDA(cudaGetLastError());
	return GPX_OK;
}

}  // namespace gpx
