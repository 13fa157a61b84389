// gpx_tick.cu — the fixed-timestep physics tick, one fused kernel per tick for an ensemble of small worlds.
//
// Replaces JPH_PhysicsSystem_Update(system, dt, collisionSteps = 2, jobSystem) as the reference calls it from
// MapFixedUpdate (engine/src/physics/MapPhysics.c:105-108).  One TILE of lanes owns one world; the world's bodies and
// contact manifolds live in shared memory for the whole tick, so HBM sees each body exactly once in and once out per
// tick (structure-of-arrays body store, float4 loads/stores, lane = body).  The warm-start cache of the previous
// sub-step lives in global memory (L2-resident) so that the shared-memory footprint of a world stays under 8 KB and
// every world of a 4096-world ensemble is resident at once (one wave, ~28 worlds per SM).
//
// Phases of a sub-step (barriers are tile-wide; a tile never spans warps):
//   1 lane/body     gravity, damping, velocity clamp, world inverse inertia, AABB          ("integrate velocities")
//   2 lane/body     LBVH box query -> candidate triangles; box/sphere vs triangle SAT + face clipping;
//                   manifolds grouped by normal per static body                           ("narrowphase vs map")
//   3 lane/body     all-pairs AABB sweep over the world's <= 64 bodies -> ordered pair list ("broadphase")
//   4 lane/pair     box-box / sphere contact manifolds, one thread per pair               ("narrowphase")
//   5 lane/manifold match against the previous sub-step's cache, carry impulses           ("warm start")
//   6 lane 0        greedy graph colouring in canonical manifold order
//   7 lane/manifold constraint rows held in REGISTERS for the whole velocity solve; per colour the lane pulls its two
//                   bodies' velocities from shared memory, runs its rows, pushes them back  ("coloured Gauss-Seidel")
//   8 lane/body     integrate positions and rotations                                      ("integrate")
//   9 lane/manifold per colour: 2 x Baumgarte position solve
// The solve order (colour, manifold index) and every arithmetic expression are fixed, so results are reproducible
// and identical for any TILE width.
#include <cooperative_groups.h>

#include <atomic>
#include <cstdlib>

#include "gpx_solver.cuh"

namespace cg = cooperative_groups;

namespace gpx {

constexpr uint32_t MAX_BUSY_WORLDS = 512;  // grid of the 32-lane launch; further busy worlds stay with the narrow launch

struct TickArgs
{
	BodyStore bs;
	ManifoldCache mc;
	const float4 *nodes;
	const float4 *tris;
	uint32_t n_nodes;
	uint32_t *err;  // [0] OR of all worlds' errors, [1 + world] per world
	float4 *con_park;  // 31 float4 per manifold slot: parked solver rows of worlds with more manifolds than lanes
	uint4 *cand;  // per body 8 x uint4: {count, -, -, -}, {fat lo xyz, -}, {fat hi xyz, -}... see cand_* below
	// contact events (gpx_events_enable): per world the sorted touching pairs of the previous tick and this tick's events
	const unsigned long long *ch_keys;  // the player character's contacts (gpx_char.cu), 64 per world, or nullptr
	const uint32_t *ch_nkeys;
	unsigned long long *ev_prev;
	uint32_t *ev_nprev;
	uint4 *ev_out;
	uint32_t *ev_count;
	const uint32_t *busy_list, *busy_count;  // worlds routed to the 32-lane launch (that launch only)
	const uint8_t *busy_flag;                // narrow launch: skip these worlds
	// the same three for the NEXT tick, written by whichever launch runs a world (nullptr: no split for this world set)
	uint32_t *next_list, *next_count;
	uint8_t *next_flag;
	const uint8_t *cur_flag;                 // this tick's flags as seen by BOTH launches (busy_flag is the narrow launch's)
	uint32_t next_above;                     // a world that ends the tick with more manifolds than this is "busy"
	unsigned long long *phase_cycles;  // optional (gpx_debug_phase_cycles): SM cycles per phase summed over tiles' lane 0
	uint32_t jitter;                   // debugging (GPX_DEBUG_JITTER): lanes leave every barrier after a random delay
	float4 *host_pos, *host_quat;      // the host's back mirror buffer (mapped pinned memory), or nullptr
	uint8_t *mirror_fresh;             // per world: which of the host's two mirror buffers hold its current state
	uint32_t mirror_back;              // the buffer host_pos points into (0 / 1)
	TickParams p;
};

enum Phase { PH_LOAD, PH_FORCES, PH_STATIC, PH_PAIRS, PH_MATCH, PH_COLOUR, PH_SETUP, PH_WARM, PH_VELOCITY, PH_INTEGRATE,
			 PH_POSITION, PH_CACHE, PH_STORE, PH_COUNT };

struct PhaseClock
{
	unsigned long long *out;
	long long t;
	__device__ __forceinline__ void start(unsigned long long *o, int lane)
	{
		out = lane == 0 ? o : nullptr;
		if (out) t = clock64();
	}
	__device__ __forceinline__ void mark(int phase)
	{
		if (out)
		{
			long long n = clock64();
			atomicAdd(&out[phase], (unsigned long long)(n - t));
			t = n;
		}
	}
};

// One Scratch per lane while contacts are generated; once they are, the same bytes hold the keys of the previous
// sub-step's manifolds (a, b, np) and the list of active manifolds.
__host__ __device__ inline size_t world_scratch_bytes(uint32_t tile, uint32_t cap_m)
{
	// also: (cap_m + 64) 8-byte keys for the contact-event pass at the end of the tick
	size_t a = (sizeof(Scratch) > sizeof(FricRows) ? sizeof(Scratch) : sizeof(FricRows)) * tile, b = sizeof(uint32_t) * 4 * cap_m,
		   c = 8u * ((size_t)cap_m + CHARACTER_MAX_CONTACTS);
	if (c > b) b = c;
	return ((a > b ? a : b) + 15u) & ~(size_t)15u;
}

__host__ __device__ inline size_t world_smem_bytes(uint32_t tile, uint32_t cap, uint32_t cap_m)
{
	size_t b = 0;
	b += sizeof(unsigned long long) * cap;  // per-body pair masks / colour sets (first: 8-byte aligned)
	b += sizeof(SBody) * cap;
	b += sizeof(SMan) * cap_m;
	b += sizeof(uint32_t) * 2 * cap_m;      // pair list (a | slot << 16, b)
	b += sizeof(uint32_t) * cap_m;          // active manifolds in canonical order
	b += world_scratch_bytes(tile, cap_m) + 16;  // narrowphase polygon scratch, later the cached keys, then the friction rows (16-byte aligned)
	b += sizeof(uint32_t) * 2 * cap;        // per-body counts, bases
	b += sizeof(uint32_t) * 8;              // header
	return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ bool sensor_pair(const SMan &m, const SBody *bodies)
{
	return m.b < STATIC_BODY_BASE && ((bodies[m.a].flags | bodies[m.b].flags) & BF_SENSOR) != 0;
}

// Velocity solve with one lane per active manifold: the manifold's rows are built once into REGISTERS, then warm start
// and the iterations; per colour a lane pulls its bodies' velocities from shared memory, runs its rows and pushes them
// back.
template <int TILE, typename Tile>
__device__ __forceinline__ void solve_one_per_lane(Tile &tile, int lane, SMan *man, const uint32_t *act, uint32_t nact,
												   SBody *bodies, FricRows &F, int ncol, uint32_t vel_steps, float h, PhaseClock &pc)
{
	const bool mine = (uint32_t)lane < nact;
	SMan &m = man[mine ? act[lane] : 0];
	Con c;
	Rows R;
	int colour = -1;
	tile.sync();  // the cached keys that share `F`'s bytes are dead from here on
	if (mine)
	{
		build_rows(c, R, m, bodies, h);
		F = R.f;
		colour = m.colour;
	}
	// every lane has read the velocities its rows' restitution bias is made of before any lane's warm start changes them
	// (the set-up of ALL manifolds precedes the first impulse)
	tile.sync();
	pc.mark(PH_SETUP);
	for (int col = 0; col < ncol; col++)
	{
		if (colour == col)
		{
			Vel u;
			load_vel(c, bodies, u);
			warm_start(c, R, F, u);
			store_vel(c, bodies, u);
		}
		tile.sync();
	}
	pc.mark(PH_WARM);
	for (uint32_t it = 0; it < vel_steps; it++)
		for (int col = 0; col < ncol; col++)
		{
			if (colour == col)
			{
				Vel u;
				load_vel(c, bodies, u);
				solve_velocity(c, R, F, u, it);
				store_vel(c, bodies, u);
			}
			tile.sync();
		}
	if (mine) save_impulses(m, R);
}

// Routing of the next tick, decided where the manifold count is known (the end of this one): next_list[0 .. n) = worlds
// for the 32-lane launch, next_flag[world] = 1 for exactly those.  The order of the list does not matter.
// A world that needed the wide launch stays on it for BUSY_STICKY more ticks: a tipping column hovers around the limit,
// and every tick it spends on a narrow tile with one manifold too many holds up the whole wave.  The flag byte is the
// number of ticks left.
constexpr uint32_t BUSY_STICKY = 30;
__device__ __forceinline__ void route_next(const TickArgs &a, uint32_t world, uint32_t count)
{
	if (!a.next_flag) return;
	const uint32_t left = a.cur_flag[world];
	uint32_t f = count > a.next_above ? BUSY_STICKY : (left > 1u ? left - 1u : 0u);
	if (count == 0) f = 0;  // an idle world has no tick to run
	if (f)
	{
		const uint32_t k = atomicAdd(a.next_count, 1u);
		if (k < MAX_BUSY_WORLDS)
			a.next_list[k] = world;
		else
			f = 0;
	}
	a.next_flag[world] = (uint8_t)f;
}

// A world's lanes: a tile of a warp (8 / 16 / 32 lanes, barriers are warp barriers) or, for worlds of a few dozen bodies and
// a hundred-odd manifolds, a whole block (TILE = 256: one manifold per thread still holds — with 32 lanes such a world
// revisited five manifolds per lane and colour phase and parked their rows in L2).
struct BlockTile
{
	__device__ __forceinline__ int thread_rank() const { return (int)threadIdx.x; }
	__device__ __forceinline__ void sync() const { __syncthreads(); }
	__device__ __forceinline__ bool any(bool p) const { return __syncthreads_or(p ? 1 : 0) != 0; }
};

// (Builds with -DGPX_JITTER only.)  A world's lanes with a debugging switch on their barriers: with `jitter` set, every lane leaves every barrier after its
// own pseudo-random delay (up to a microsecond), so lanes enter each phase staggered.  Code that needs a barrier it does
// not have then fails under tests/fuzz_parity.py instead of once in fifty seeds; results must not change.
template <typename Tile>
struct JitterTile
{
	Tile t;
	uint32_t jitter;
	__device__ __forceinline__ int thread_rank() const { return (int)t.thread_rank(); }
	__device__ __forceinline__ void delay() const
	{
		if (!jitter) return;
		uint32_t h = (threadIdx.x * 2654435761u) ^ (uint32_t)clock64() ^ jitter;
		h ^= h >> 13;
		h *= 0x5BD1E995u;
		h ^= h >> 15;
		const long long until = clock64() + (long long)(h & 2047u);
		while (clock64() < until) {}
	}
	__device__ __forceinline__ void sync() const
	{
		t.sync();
		delay();
	}
	__device__ __forceinline__ bool any(bool p) const
	{
		const bool r = t.any(p);
		delay();
		return r;
	}
};

template <int TILE>
__device__ __forceinline__ auto make_tile()
{
	if constexpr (TILE <= 32)
		return cg::tiled_partition<TILE>(cg::this_thread_block());
	else
		return BlockTile{};
}

template <int TILE>
__global__ void __launch_bounds__(TILE > 128 ? TILE : 128) k_tick(TickArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
#ifdef GPX_JITTER  // make -C csrc clean && make -C csrc NVFLAGS_EXTRA=-DGPX_JITTER, then run with GPX_DEBUG_JITTER=1
	auto raw_tile = make_tile<TILE>();
	JitterTile<decltype(raw_tile)> tile{raw_tile, a.jitter};
#else
	auto tile = make_tile<TILE>();
#endif
	const int lane = tile.thread_rank();
	const uint32_t tiles_per_block = blockDim.x / TILE;
	const uint32_t slot = blockIdx.x * tiles_per_block + threadIdx.x / TILE;
	const uint32_t cap = a.p.cap, cap_m = a.p.cap_m;
	// Worlds whose last tick ended with more manifolds than a narrow tile has lanes (a toppled column) are handed to a
	// second launch of this kernel with 32 lanes per world, so they do not hold up the wave: a.busy_list != nullptr
	// selects that launch; a.split_above > 0 makes the narrow launch skip them.
	uint32_t world = slot;
	if (a.busy_list)
	{
		if (slot >= min(*a.busy_count, MAX_BUSY_WORLDS)) return;
		world = a.busy_list[slot];
	}
	else
	{
		if (world >= a.p.worlds) return;  // whole tile exits together
		if (a.busy_flag && a.busy_flag[world]) return;
	}

	unsigned char *base = smem_raw + world_smem_bytes(TILE, cap, cap_m) * (threadIdx.x / TILE);
	unsigned long long *pmask = reinterpret_cast<unsigned long long *>(base);
	SBody *bodies = reinterpret_cast<SBody *>(pmask + cap);
	SMan *man = reinterpret_cast<SMan *>(bodies + cap);
	uint32_t *pair_a = reinterpret_cast<uint32_t *>(man + cap_m);
	uint32_t *pair_b = pair_a + cap_m;
	uint32_t *act = pair_b + cap_m;
	// One region, three lives per sub-step: a polygon Scratch per lane while contacts are generated; the cached keys
	// during the warm-start match; the friction rows of the lane's manifold during the velocity solve.
	unsigned char *scratch_raw = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(act + cap_m) + 15u) & ~(uintptr_t)15u);
	Scratch &scratch = reinterpret_cast<Scratch *>(scratch_raw)[lane];
	FricRows &fric = reinterpret_cast<FricRows *>(scratch_raw)[lane];
	uint32_t *pkey_a = reinterpret_cast<uint32_t *>(scratch_raw);
	uint32_t *pkey_b = pkey_a + cap_m;
	uint32_t *pkey_np = pkey_b + cap_m;
	uint32_t *pkey_tri = pkey_np + cap_m;
	uint32_t *cnt_static = reinterpret_cast<uint32_t *>(scratch_raw + world_scratch_bytes(TILE, cap_m));
	uint32_t *slot_base = cnt_static + cap;
	uint32_t *hdr = slot_base + cap;  // 0 nman, 1 nprev, 2 ncol, 3 err, 4 npairs, 5 nact

	PhaseClock pc;
	pc.start(a.phase_cycles, lane);
	const uint32_t g0 = world * cap;
	bool any_active = false;
	// ---- load: HBM -> shared, lane = body, 16-byte vector loads
	for (uint32_t i = lane; i < cap; i += TILE)
	{
		SBody &b = bodies[i];
		const float4 p = a.bs.pos[g0 + i], q = a.bs.quat[g0 + i], l = a.bs.lin[g0 + i], w = a.bs.ang[g0 + i];
		const float4 p0 = a.bs.prop0[g0 + i], p1 = a.bs.prop1[g0 + i], p2 = a.bs.prop2[g0 + i];
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
		uint32_t f = a.bs.flags[g0 + i];
		if ((f & BF_ALIVE) && ((f >> BF_MOTION_SHIFT) & 3u) == GPX_MOTION_KINEMATIC &&
			(l.x != 0.0f || l.y != 0.0f || l.z != 0.0f || w.x != 0.0f || w.y != 0.0f || w.z != 0.0f))
			f |= BF_KIN_MOVING;
		b.flags = f;
		if ((f & BF_ALIVE) && is_active_body(f)) any_active = true;
	}
	// A world in which everything sleeps (or nothing can move) has no tick to run.  Its bodies stay as they are; the
	// contact cache empties, as it would if the tick ran (sleepers generate no contacts), so a wake-up starts cold
	// whether or not this shortcut was taken.  (With contact events on, the pass below still has to report the pairs.)
	if (!a.ev_out && !tile.any(any_active))
	{
		if (lane == 0)
		{
			a.mc.count[world] = 0;
			route_next(a, world, 0);
		}
		// the mirror's two buffers alternate: this one may be a tick older than the world's last change — once
		if (a.host_pos && !((a.mirror_fresh[world] >> a.mirror_back) & 1u))
		{
			for (uint32_t i = lane; i < cap; i += TILE)
			{
				a.host_pos[g0 + i] = a.bs.pos[g0 + i];
				a.host_quat[g0 + i] = a.bs.quat[g0 + i];
			}
			tile.sync();
			if (lane == 0) a.mirror_fresh[world] |= (uint8_t)(1u << a.mirror_back);
		}
		return;
	}
	const uint32_t m0 = world * cap_m;
	if (lane == 0)
	{
		hdr[1] = min(a.mc.count[world], cap_m);
		hdr[3] = 0;
		hdr[6] = hdr[7] = 0;  // sleepers touched by an active body in this sub-step (bit = body)
	}
	tile.sync();
	pc.mark(PH_LOAD);

	const float h = a.p.h;
	const v3 gravity = V(a.p.gx, a.p.gy, a.p.gz);

	for (int sub = 0; sub < a.p.substeps; sub++)
	{
		// ---- 1: forces, inertia, bounds
		for (uint32_t i = lane; i < cap; i += TILE)
		{
			SBody &b = bodies[i];
			if (!(b.flags & BF_ALIVE)) continue;
			if (is_dynamic(b.flags))
			{
				const uint32_t dofs = dofs_of(b.flags);
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
		}
		tile.sync();
		pc.mark(PH_FORCES);

		// ---- 2 + 3: per body: contacts with the static map (kept in registers/local until slots are known) and the
		// mask of higher-numbered bodies whose boxes overlap
		for (uint32_t i0 = 0; i0 < cap; i0 += TILE)
		{
			const uint32_t i = i0 + lane;
			StaticSlot slots[MAX_STATIC_PER_BODY];
			int nslots = 0;
			unsigned long long mask = 0ull;
			uint32_t err = 0;
			if (i < cap && (bodies[i].flags & BF_ALIVE) && shape_of(bodies[i].flags) != GPX_SHAPE_EMPTY)
			{
				const SBody &A = bodies[i];
				const uint32_t fa = A.flags;
				const uint32_t la = layer_of(fa);
				if (is_dynamic(fa) && !(fa & BF_SENSOR) && (la == 1 || la == 2))
				{
					const StaticView sv = {a.nodes, a.tris, a.n_nodes};
					nslots = body_static_contacts(sv, a.cand + 8ull * (g0 + i), A, scratch, slots, err);
				}
				for (uint32_t j = i + 1; j < cap; j++)
				{
					const SBody &B = bodies[j];
					const uint32_t fb = B.flags;
					if (!(fb & BF_ALIVE) || shape_of(fb) == GPX_SHAPE_EMPTY) continue;
					// at least one awake dynamic body — or a moving kinematic body reaching a sleeper, which only wakes it
					if (!is_dynamic(fa) && !is_dynamic(fb) &&
						!(((fa & BF_KIN_MOVING) && (fb & BF_ASLEEP)) || ((fb & BF_KIN_MOVING) && (fa & BF_ASLEEP))))
						continue;
					if (!layers_collide(la, layer_of(fb))) continue;
					if (!aabb_overlap(A.lo, A.hi, B.lo, B.hi, SPECULATIVE_DISTANCE)) continue;
					mask |= 1ull << j;
				}
			}
			if (i < cap)
			{
				cnt_static[i] = (uint32_t)nslots;
				pmask[i] = mask;
			}
			if (err) atomicOr(&hdr[3], err);
			tile.sync();
			// slots for this chunk of bodies.  A block-wide tile holds all bodies in one chunk: every body lane sums the counts
			// of the bodies before it (at most 63) and writes its own pairs — the same slots and the same order as the serial
			// prefix below, which is what a tile of a warp uses for its handful of bodies.
			if constexpr (TILE > 32)
			{
				if (lane == 0) hdr[4] = 0;
				tile.sync();
				if (i < cap)
				{
					uint32_t n = 0, np = 0;
					for (uint32_t k = 0; k < i; k++)
					{
						const uint32_t pk = (uint32_t)__popcll(pmask[k]);
						n += cnt_static[k] + pk;
						np += pk;
					}
					slot_base[i] = n;
					n += cnt_static[i];
					uint32_t stored = 0;
					unsigned long long pm = pmask[i];
					while (pm)
					{
						const int j = __ffsll((long long)pm) - 1;
						pm &= pm - 1;
						if (n < cap_m && np < cap_m)
						{
							pair_a[np] = i | (n << 16);
							pair_b[np] = (uint32_t)j;
							stored++;
						}
						np++;
						n++;
					}
					if (stored) atomicAdd(&hdr[4], stored);
					if (i == cap - 1)
					{
						if (n > cap_m) atomicOr(&hdr[3], (uint32_t)GPX_ERR_CONTACT_CONSTRAINTS_FULL);
						hdr[0] = min(n, cap_m);
					}
				}
			}
			else if (lane == 0)
			{
				uint32_t n = i0 == 0 ? 0 : hdr[0];
				uint32_t np = i0 == 0 ? 0 : hdr[4];
				for (uint32_t k = i0; k < min(i0 + TILE, cap); k++)
				{
					slot_base[k] = n;
					n += cnt_static[k];
					unsigned long long pm = pmask[k];
					while (pm)
					{
						int j = __ffsll((long long)pm) - 1;
						pm &= pm - 1;
						if (n < cap_m && np < cap_m)
						{
							pair_a[np] = k | (n << 16);
							pair_b[np] = (uint32_t)j;
							np++;
						}
						n++;
					}
				}
				if (n > cap_m) hdr[3] |= GPX_ERR_CONTACT_CONSTRAINTS_FULL;
				hdr[0] = min(n, cap_m);
				hdr[4] = np;
			}
			tile.sync();
			if (i < cap)
			{
				const SBody &A = bodies[i];
				for (int s = 0; s < nslots; s++)
				{
					const uint32_t slot = slot_base[i] + s;
					if (slot >= cap_m) break;
					SMan &m = man[slot];
					m.a = i;
					m.b = STATIC_BODY_BASE + slots[s].sbody;
					m.tri = slots[s].tri;
					m.n = slots[s].n;
					m.friction = slots[s].friction;
					m.restitution = A.restitution;
					store_points(m, A, nullptr, slots[s].np, slots[s].p1, slots[s].p2);
				}
			}
		}
		tile.sync();
		pc.mark(PH_STATIC);

		// ---- 4: body-body contact manifolds, one lane per candidate pair
		const uint32_t npairs = hdr[4];
		for (uint32_t pi = lane; pi < npairs; pi += TILE)
		{
			const uint32_t ia = pair_a[pi] & 0xFFFFu, slot = pair_a[pi] >> 16, ib = pair_b[pi];
			const SBody &A = bodies[ia];
			const SBody &B = bodies[ib];
			SMan &m = man[slot];
			m.a = ia;
			m.b = ib;
			m.np = 0;
			m.tri = 0;
			pair_contact(A, B, scratch, m);
			const uint32_t fa = A.flags, fb = B.flags;
			if (((fa | fb) & BF_ASLEEP) && m.np > 0 && !((fa | fb) & BF_SENSOR))
			{
				// a contact with an active body wakes a sleeper; it takes part in the solve from the next sub-step on
				if ((fa & BF_ASLEEP) && is_active_body(fb)) atomicOr(&hdr[6 + (ia >> 5)], 1u << (ia & 31u));
				if ((fb & BF_ASLEEP) && is_active_body(fa)) atomicOr(&hdr[6 + (ib >> 5)], 1u << (ib & 31u));
				if (!is_dynamic(fa) && !is_dynamic(fb)) m.np = 0;  // kinematic against sleeper: nothing to solve
			}
		}
		tile.sync();
		pc.mark(PH_PAIRS);

		// ---- 5: carry impulses from the previous sub-step's manifolds (keys in shared memory, records in L2)
		const uint32_t nman = hdr[0], nprev = hdr[1];
		for (uint32_t i = lane; i < nprev; i += TILE)
		{
			const uint4 k = __ldcg(&a.mc.key[m0 + i]);
			pkey_a[i] = k.x;
			pkey_b[i] = k.y;
			pkey_np[i] = k.z;
			pkey_tri[i] = k.w;
		}
		tile.sync();
		for (uint32_t mi0 = 0; mi0 < nman; mi0 += TILE)
		{
			// stage 1 (shared memory only, no early exits): first cached record with this manifold's key, and how many
			const uint32_t mi = mi0 + lane;
			SMan &m = man[mi < nman ? mi : 0];
			const bool live = mi < nman && m.np > 0 && !sensor_pair(m, bodies);
			int j0 = -1, nmatch = 0;
			for (uint32_t j = 0; j < nprev; j++)
			{
				const bool hit = live && pkey_a[j] == m.a && pkey_b[j] == m.b && pkey_tri[j] == m.tri;
				if (hit && j0 < 0) j0 = (int)j;
				nmatch += hit ? 1 : 0;
			}
			// stage 2: every lane with a match pulls its record at once (one overlapped batch of L2 reads per tile)
			bool got_cf = false;
			for (int j = j0; j >= 0 && nmatch > 0; )
			{
				const uint32_t onp = pkey_np[j];
				float4 c1[4], c2[4];
#pragma unroll
				for (int k = 0; k < 4; k++)
				{
					c1[k] = __ldcg(&a.mc.p1[4 * (m0 + j) + k]);
					c2[k] = __ldcg(&a.mc.p2[4 * (m0 + j) + k]);
				}
#pragma unroll
				for (int p = 0; p < 4; p++)
				{
					if (p >= m.np) continue;
					if (m.ln[p] != 0.0f) continue;
					const v3 a1 = m.p1l[p], a2 = m.p2l[p];
					bool done = false;
#pragma unroll
					for (int k = 0; k < 4; k++)
					{
						const bool ok = !done && (uint32_t)k < onp && len2(a1 - V(c1[k])) < PRESERVE_LAMBDA_MAX_DIST_SQ &&
										len2(a2 - V(c2[k])) < PRESERVE_LAMBDA_MAX_DIST_SQ;
						if (ok)
						{
							m.ln[p] = c1[k].w;
							// the friction impulse of the manifold comes from the first old manifold a point is found in
							if (!got_cf)
							{
								got_cf = true;
								m.cf[0] = c2[0].w;
								m.cf[1] = c2[1].w;
								m.cf[2] = c2[2].w;
							}
							done = true;
						}
					}
				}
				// further records with the same key (several manifolds of one body against one static body): rare
				nmatch--;
				int jn = -1;
				if (nmatch > 0)
					for (uint32_t jj = (uint32_t)j + 1; jj < nprev; jj++)
						if (pkey_a[jj] == m.a && pkey_b[jj] == m.b && pkey_tri[jj] == m.tri)
						{
							jn = (int)jj;
							break;
						}
				j = jn;
			}
		}
		tile.sync();  // the cached keys are overwritten below; a lane of another warp may still be matching against them
		pc.mark(PH_MATCH);
		// ---- 6: greedy colouring in canonical order (only dynamic bodies constrain a colour); active list.  The lanes
		// classify the manifolds into one word each (solved or not, the slots of its dynamic bodies); lane 0 then walks
		// those words, so its chain per manifold is one load, two colour sets, a first-fit and the stores — it used to
		// chase manifold -> bodies -> flags through shared memory for every one of them.
		{
			uint32_t *desc = pkey_a;  // the cached keys are dead from here on
			for (uint32_t mi = lane; mi < nman; mi += TILE)
			{
				SMan &m = man[mi];
				uint32_t d = 0;
				if (m.np == 0)
					m.colour = -1;
				else if (sensor_pair(m, bodies))
					m.colour = -3;  // touching, reported as a contact event, never solved
				else
				{
					const bool a_dyn = is_dynamic(bodies[m.a].flags);
					const bool b_dyn = m.b < STATIC_BODY_BASE && is_dynamic(bodies[m.b].flags);
					d = 0x80000000u | (a_dyn ? m.a : 0xFFu) | ((b_dyn ? m.b : 0xFFu) << 8);
				}
				desc[mi] = d;
			}
			tile.sync();
			if (lane == 0)
			{
				unsigned long long *used = pmask;  // reuse: per-body colour sets
				for (uint32_t k = 0; k < cap; k++) used[k] = 0ull;
				int ncol = 0;
				uint32_t nact = 0;
				for (uint32_t mi = 0; mi < nman; mi++)
				{
					const uint32_t d = desc[mi];
					if (!d) continue;
					const uint32_t ia = d & 0xFFu, ib = (d >> 8) & 0xFFu;
					unsigned long long u = 0ull;
					if (ia != 0xFFu) u |= used[ia];
					if (ib != 0xFFu) u |= used[ib];
					const int c = ~u ? min(63, __ffsll((long long)~u) - 1) : 63;  // first colour neither body has yet
					man[mi].colour = c;
					if (ia != 0xFFu) used[ia] |= 1ull << c;
					if (ib != 0xFFu) used[ib] |= 1ull << c;
					if (c + 1 > ncol) ncol = c + 1;
					act[nact++] = mi;
				}
				hdr[2] = (uint32_t)ncol;
				hdr[5] = nact;
			}
		}
		tile.sync();
		const int ncol = (int)hdr[2];
		const uint32_t nact = hdr[5];
		pc.mark(PH_COLOUR);

		// ---- 7: set-up, warm start, velocity iterations; within a colour no two manifolds share a dynamic body
		if (nact <= (uint32_t)TILE)
			solve_one_per_lane<TILE>(tile, lane, man, act, nact, bodies, fric, ncol, a.p.vel_steps, h, pc);
		else
		{
			// more manifolds than lanes: a lane revisits several manifolds, so their rows are parked in an L2-resident
			// scratch record and pulled back each visit
			float4 *park = a.con_park + 31ull * m0;
			for (uint32_t k = lane; k < nact; k += TILE)
			{
				Con c;
				Rows R;
				build_rows(c, R, man[act[k]], bodies, h);
				park_rows(R, park + 31ull * act[k]);
			}
			__threadfence_block();
			tile.sync();
			for (uint32_t it = 0; it <= a.p.vel_steps; it++)
				for (int col = 0; col < ncol; col++)
				{
					for (uint32_t k = lane; k < nact; k += TILE)
					{
						SMan &m = man[act[k]];
						if (m.colour != col) continue;
						Con c;
						Rows R;
						con_header(c, m, bodies);
						unpark_rows(R, park + 31ull * act[k]);
						Vel u;
						load_vel(c, bodies, u);
						if (it == 0)
							warm_start(c, R, R.f, u);
						else
							solve_velocity(c, R, R.f, u, it - 1u);
						store_vel(c, bodies, u);
						// only the accumulated impulses change: float4 14 (ln) and 15 (cf) of the parked record
						float4 *g = park + 31ull * act[k];
						__stcg(&g[14], make_float4(R.ln[0], R.ln[1], R.ln[2], R.ln[3]));
						__stcg(&g[15], make_float4(R.cf[0], R.cf[1], R.cf[2], 0.0f));
						if (it == a.p.vel_steps) save_impulses(m, R);
					}
					tile.sync();
				}
		}
		tile.sync();
		pc.mark(PH_VELOCITY);

		// ---- 8: integrate
		for (uint32_t i = lane; i < cap; i += TILE)
		{
			SBody &b = bodies[i];
			if (!(b.flags & BF_ALIVE) || ((b.flags >> BF_MOTION_SHIFT) & 3u) == GPX_MOTION_STATIC || (b.flags & BF_ASLEEP)) continue;
			b.x = b.x + (b.v * h);
			b.q = qstep(b.q, b.w * h);
		}
		tile.sync();
		pc.mark(PH_INTEGRATE);

		// ---- 9: position iterations
		for (uint32_t it = 0; it < a.p.pos_steps; it++)
		{
			bool moved = false;
			for (int c = 0; c < ncol; c++)
			{
				for (uint32_t k = lane; k < nact; k += TILE)
					if (man[act[k]].colour == c) moved = solve_position(man[act[k]], bodies) || moved;
				tile.sync();
			}
			// an iteration that found every point inside the slop changed nothing: the remaining ones would find the same
			if (!tile.any(moved)) break;
		}

		pc.mark(PH_POSITION);
		// ---- this sub-step's manifolds become the warm-start cache (compacted, canonical order kept)
		for (uint32_t k = lane; k < nact; k += TILE)
		{
			const SMan &m = man[act[k]];
			__stcg(&a.mc.key[m0 + k], make_uint4(m.a, m.b, (uint32_t)m.np, m.tri));
#pragma unroll
			for (int p = 0; p < 4; p++)
			{
				__stcg(&a.mc.p1[4 * (m0 + k) + p], F4(m.p1l[p], m.ln[p]));
				__stcg(&a.mc.p2[4 * (m0 + k) + p], F4(m.p2l[p], p < 3 ? m.cf[p] : 0.0f));
			}
		}
		if (lane == 0) hdr[1] = nact;
		if (hdr[6] | hdr[7])
		{
			// wake: awake from the next sub-step on; the sleep test starts over
			for (uint32_t i = lane; i < cap; i += TILE)
				if ((hdr[6 + (i >> 5)] >> (i & 31u)) & 1u)
				{
					bodies[i].flags &= ~BF_ASLEEP;
					a.bs.flags[g0 + i] = bodies[i].flags & ~BF_KIN_MOVING;
					a.bs.sleep_t[g0 + i] = -1.0f;
				}
			tile.sync();
			if (lane == 0) hdr[6] = hdr[7] = 0;
		}
		__threadfence_block();
		tile.sync();
		pc.mark(PH_CACHE);
	}

	// ---- contact events: touching pairs after the last sub-step (solver manifolds + sensor overlaps) against the
	// previous tick's; canonical order = added/persisted sorted by (a, b), then removed sorted by (a, b)
	if (a.ev_out && lane == 0)
	{
		unsigned long long *keys = reinterpret_cast<unsigned long long *>(scratch_raw);  // free after the last sub-step
		const uint32_t nman = hdr[0];
		uint32_t n = 0;
		for (uint32_t mi = 0; mi < nman; mi++)
		{
			if (man[mi].np == 0) continue;
			const unsigned long long key = ((unsigned long long)man[mi].a << 32) | man[mi].b;
			uint32_t k = n;
			bool dup = false;
			while (k > 0 && keys[k - 1] >= key)
			{
				if (keys[k - 1] == key)
				{
					dup = true;
					break;
				}
				k--;
			}
			if (dup) continue;
			for (uint32_t t = n; t > k; t--) keys[t] = keys[t - 1];
			keys[k] = key;
			n++;
		}
		// the player character's contacts join the same list (pseudo body id CHARACTER_BODY_ID)
		const uint32_t nch = a.ch_keys ? min(a.ch_nkeys[world], CHARACTER_MAX_CONTACTS) : 0u;
		for (uint32_t ci = 0; ci < nch; ci++)
		{
			const unsigned long long key = a.ch_keys[(size_t)world * CHARACTER_MAX_CONTACTS + ci];
			uint32_t k = n;
			bool dup = false;
			while (k > 0 && keys[k - 1] >= key)
			{
				if (keys[k - 1] == key)
				{
					dup = true;
					break;
				}
				k--;
			}
			if (dup) continue;
			for (uint32_t t = n; t > k; t--) keys[t] = keys[t - 1];
			keys[k] = key;
			n++;
		}
		const uint32_t ev_stride = cap_m + CHARACTER_MAX_CONTACTS;
		unsigned long long *prev = a.ev_prev + (size_t)world * ev_stride;
		const uint32_t np = a.ev_nprev[world];
		uint4 *out = a.ev_out + (size_t)world * 2u * ev_stride;
		uint32_t e = 0, j = 0;
		for (uint32_t i = 0; i < n; i++)
		{
			while (j < np && prev[j] < keys[i]) j++;
			const bool persisted = j < np && prev[j] == keys[i];
			out[e++] = make_uint4((uint32_t)(keys[i] >> 32), (uint32_t)(keys[i] & 0xFFFFFFFFull), persisted ? 2u : 1u, world);
		}
		uint32_t i = 0;
		for (j = 0; j < np; j++)
		{
			while (i < n && keys[i] < prev[j]) i++;
			if (i < n && keys[i] == prev[j]) continue;
			out[e++] = make_uint4((uint32_t)(prev[j] >> 32), (uint32_t)(prev[j] & 0xFFFFFFFFull), 3u, world);
		}
		for (uint32_t t = 0; t < n; t++) prev[t] = keys[t];
		a.ev_nprev[world] = n;
		a.ev_count[world] = e;
	}
	tile.sync();

	// ---- store: shared -> HBM
	for (uint32_t i = lane; i < cap; i += TILE)
	{
		const SBody &b = bodies[i];
		if (!(b.flags & BF_ALIVE)) continue;
		a.bs.pos[g0 + i] = F4(b.x, 0.0f);
		a.bs.quat[g0 + i] = make_float4(b.q.x, b.q.y, b.q.z, b.q.w);
		a.bs.lin[g0 + i] = F4(b.v, 0.0f);
		a.bs.ang[g0 + i] = F4(b.w, 0.0f);
		if (a.host_pos)
		{
			// straight into the host's mirror: worlds finish at different times, so by the end of the launch most of
			// this has already crossed the bus and the read-back copy is not needed
			a.host_pos[g0 + i] = F4(b.x, 0.0f);
			a.host_quat[g0 + i] = make_float4(b.q.x, b.q.y, b.q.z, b.q.w);
		}
	}
	if (lane == 0)
	{
		if (a.mirror_fresh) a.mirror_fresh[world] = a.host_pos ? (uint8_t)(1u << a.mirror_back) : (uint8_t)0;
		a.mc.count[world] = hdr[1];
		route_next(a, world, hdr[1]);
		if (hdr[3])
		{
			a.err[1 + world] |= hdr[3];
			atomicOr(&a.err[0], hdr[3]);
		}
	}
	pc.mark(PH_STORE);
}

__global__ void k_apply_commands(BodyStore bs, const BodyCommand *__restrict__ cmd, uint32_t n)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const BodyCommand c = cmd[i];
	if (c.mask & 1u) bs.pos[c.index] = c.pos;
	if (c.mask & 2u) bs.quat[c.index] = c.quat;
	if (c.mask & 4u) bs.lin[c.index] = c.lin;
	if (c.mask & 8u) bs.ang[c.index] = c.ang;
	if (c.mask & 16u)
	{
		bs.prop0[c.index] = c.prop0;
		bs.prop1[c.index] = c.prop1;
		bs.prop2[c.index] = c.prop2;
		bs.flags[c.index] = c.flags;
	}
	if ((c.mask & 48u) == 32u)
		bs.flags[c.index] = (bs.flags[c.index] & ~(0xFFu << BF_RAYFLAG_SHIFT)) | (c.flags & (0xFFu << BF_RAYFLAG_SHIFT));
	if (c.mask & (16u | 64u))
	{
		// a new body, or one the host activated (a non-zero velocity, SetPosition(.., Activate)): awake, test restarts
		if (!(c.mask & 16u)) bs.flags[c.index] &= ~BF_ASLEEP;
		bs.sleep_t[c.index] = -1.0f;
	}
}


// ---- sleeping: Jolt's sleep test, once per tick (0.03 m/s for 0.5 s).  Three test points per body
// (centre of mass, the extents along the two larger local axes) each live in a sphere that grows to hold them; a
// radius above 15 mm restarts the test, 0.5 s without a restart makes the body a candidate, and an island — awake
// dynamic bodies joined by the contacts of the last sub-step — goes to sleep when all its bodies are candidates.
// One warp per world; union-find over at most 64 bodies in shared memory.
constexpr int SLEEP_WARPS = 4;

__global__ void __launch_bounds__(SLEEP_WARPS * 32) k_sleep(BodyStore bs, ManifoldCache mc, uint32_t worlds, uint32_t cap, uint32_t cap_m,
															float dt)
{
	__shared__ unsigned char parent_s[SLEEP_WARPS][64], can_s[SLEEP_WARPS][64];
	const uint32_t lane = threadIdx.x & 31u, wi = threadIdx.x >> 5;
	const uint32_t world = blockIdx.x * SLEEP_WARPS + wi;
	if (world >= worlds) return;
	unsigned char *parent = parent_s[wi], *can = can_s[wi];
	const uint32_t g0 = world * cap, m0 = world * cap_m;
	bool any = false;
	for (uint32_t i = lane; i < cap; i += 32u)
	{
		parent[i] = (unsigned char)i;
		can[i] = 1;
		const uint32_t f = bs.flags[g0 + i];
		if ((f & BF_ALIVE) && is_dynamic(f)) any = true;
	}
	if (!__any_sync(0xFFFFFFFFu, any)) return;
	__syncwarp();
	// islands: the smaller index becomes the root.  The cache holds the solved manifolds of the last sub-step.
	const uint32_t nman = min(mc.count[world], cap_m);
	for (uint32_t k0 = 0; k0 < nman; k0 += 32u)
	{
		const uint32_t k = k0 + lane;
		uint4 key = make_uint4(0u, STATIC_BODY_BASE, 0u, 0u);
		if (k < nman) key = __ldcg(&mc.key[m0 + k]);
		bool link = key.y < STATIC_BODY_BASE && is_dynamic(bs.flags[g0 + key.x]) && is_dynamic(bs.flags[g0 + key.y]);
		// unions one at a time, in list order (a handful per world)
		uint32_t todo = __ballot_sync(0xFFFFFFFFu, link);
		while (todo)
		{
			const int src = __ffs((int)todo) - 1;
			todo &= todo - 1u;
			const uint32_t xa = __shfl_sync(0xFFFFFFFFu, key.x, src), xb = __shfl_sync(0xFFFFFFFFu, key.y, src);
			if (lane == 0)
			{
				uint32_t ra = xa, rb = xb;
				while (parent[ra] != ra) ra = parent[ra];
				while (parent[rb] != rb) rb = parent[rb];
				if (ra < rb) parent[rb] = (unsigned char)ra;
				else if (rb < ra) parent[ra] = (unsigned char)rb;
			}
			__syncwarp();
		}
	}
	__syncwarp();
	for (uint32_t i = lane; i < cap; i += 32u)
	{
		const uint32_t f = bs.flags[g0 + i];
		if (!(f & BF_ALIVE) || !is_dynamic(f)) continue;
		const bool candidate = sleep_test_body(bs, g0 + i, f, dt);
		if (!candidate)
		{
			uint32_t r = i;
			while (parent[r] != r) r = parent[r];
			can[r] = 0;
		}
	}
	__syncwarp();
	for (uint32_t i = lane; i < cap; i += 32u)
	{
		const uint32_t f = bs.flags[g0 + i];
		if (!(f & BF_ALIVE) || !is_dynamic(f)) continue;
		uint32_t r = i;
		while (parent[r] != r) r = parent[r];
		if (!can[r]) continue;
		bs.flags[g0 + i] = f | BF_ASLEEP;
		bs.lin[g0 + i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		bs.ang[g0 + i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	}
}

// per-world summary for the end-of-run gather: one warp per world
__global__ void k_stats(BodyStore bs, ManifoldCache mc, const uint32_t *err, uint32_t worlds, uint32_t cap, uint32_t ticks,
						gpx_world_stats *out)
{
	const uint32_t world = (blockIdx.x * blockDim.x + threadIdx.x) / 32u;
	const uint32_t lane = threadIdx.x & 31u;
	if (world >= worlds) return;
	float ke = 0.0f, vmax = 0.0f;
	uint32_t awake = 0;
	unsigned long long sum = 0ull;
	for (uint32_t i = lane; i < cap; i += 32)
	{
		const uint32_t g = world * cap + i;
		const uint32_t f = bs.flags[g];
		if (!(f & BF_ALIVE)) continue;
		const float4 p = bs.pos[g], q = bs.quat[g], l = bs.lin[g], p0 = bs.prop0[g];
		// order-independent checksum of the exact position/rotation bits
		unsigned long long hsh = 1469598103934665603ull * (i + 1);
		const uint32_t w[7] = {__float_as_uint(p.x), __float_as_uint(p.y), __float_as_uint(p.z), __float_as_uint(q.x),
							   __float_as_uint(q.y), __float_as_uint(q.z), __float_as_uint(q.w)};
#pragma unroll
		for (int k = 0; k < 7; k++) hsh = (hsh ^ w[k]) * 1099511628211ull;
		sum += hsh;
		if (is_dynamic(f))
		{
			awake++;
			const float v2 = ((l.x * l.x) + (l.y * l.y)) + (l.z * l.z);
			if (p0.x > 0.0f) ke += (0.5f / p0.x) * v2;
			vmax = fmaxf(vmax, sqrtf(v2));
		}
	}
	for (int o = 16; o > 0; o >>= 1)
	{
		ke += __shfl_xor_sync(0xFFFFFFFFu, ke, o);
		vmax = fmaxf(vmax, __shfl_xor_sync(0xFFFFFFFFu, vmax, o));
		awake += __shfl_xor_sync(0xFFFFFFFFu, awake, o);
		sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
	}
	if (lane == 0)
	{
		gpx_world_stats s;
		s.kinetic_energy = ke;
		s.max_speed = vmax;
		s.awake_bodies = awake;
		s.manifolds = mc.count[world];
		s.position_checksum = sum;
		s.ticks = ticks;
		s.error = err[1 + world];
		out[world] = s;
	}
}

template <int TILE>
static int launch_tick_t(gpx_world *w, const TickArgs &a, cudaStream_t stream, uint32_t grid_worlds)
{
	// one warp per block (32 / TILE worlds): the finest granularity for spreading 4096 worlds over 148 SMs in ONE wave
	const size_t per_world = world_smem_bytes(TILE, w->cap, w->cap_m);
	const size_t budget = 200u * 1024u;
	if (per_world > budget) return GPX_ERR_CAPACITY;
	uint32_t wpb = TILE <= 32 ? 32 / TILE : 1;
	while (wpb > 1 && per_world * wpb > budget) wpb >>= 1;
	const uint32_t threads = wpb * TILE;
	const size_t smem = per_world * wpb;
	// per instantiation AND per device (function attributes are per device; one process may hold worlds on several, each
	// stepping on its own thread)
	static std::atomic<size_t> configured[GPX_MAX_DEVICES];
	const int device = w->device;
	const int dev = device >= 0 && device < GPX_MAX_DEVICES ? device : 0;
	if (device != dev || smem > configured[dev].load(std::memory_order_acquire))
	{
		GPX_CUDA(cudaFuncSetAttribute(k_tick<TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		int carve = 100;
		if (const char *e = getenv("GPX_CARVEOUT")) carve = atoi(e);  // tuning knob for profiling runs
		GPX_CUDA(cudaFuncSetAttribute(k_tick<TILE>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
		size_t old = configured[dev].load(std::memory_order_relaxed);
		while (old < smem && !configured[dev].compare_exchange_weak(old, smem, std::memory_order_release)) {}
	}
	const uint32_t grid = (grid_worlds + wpb - 1) / wpb;
	k_tick<TILE><<<grid, threads, smem, stream>>>(a);
	count_launch();
	GPX_CUDA(cudaGetLastError());
	return GPX_OK;
}

int launch_tick(gpx_world *w, float dt, int substeps)
{
	TickArgs a;
	a.bs = w->bs;
	a.mc = w->mc;
	a.nodes = w->sd.nodes;
	a.tris = w->sd.tri;
	a.n_nodes = w->sd.n_nodes;
	a.err = w->d_err;
	a.phase_cycles = w->d_phase;
	a.cand = w->d_cand;
	a.ch_keys = w->d_ch_keys;
	a.ch_nkeys = w->d_ch_nkeys;
	a.ev_prev = w->d_ev_prev;
	a.ev_nprev = w->d_ev_nprev;
	a.ev_out = w->d_ev_out;
	a.ev_count = w->d_ev_count;
	a.con_park = w->d_park;
	a.jitter = getenv("GPX_DEBUG_JITTER") ? 0x9E3779B9u + w->ticks : 0u;
	{
		// the host's back mirror buffer (the one gpx_sync_transforms will publish next)
		const uint32_t back = (w->mirror_gen.load(std::memory_order_relaxed) + 1u) & 1u;
		a.host_pos = w->mb_dev[0] && w->mb_dev[1] && w->synced_since_step ? w->mb_dev[back] : nullptr;
		a.host_quat = a.host_pos ? a.host_pos + (size_t)w->W * w->cap : nullptr;
		a.mirror_fresh = w->d_mirror_fresh;
		a.mirror_back = back;
	}
	a.p.worlds = w->W;
	a.p.cap = w->cap;
	a.p.cap_m = w->cap_m;
	a.p.vel_steps = w->cfg.velocity_steps ? w->cfg.velocity_steps : 10u;
	a.p.pos_steps = w->cfg.position_steps ? w->cfg.position_steps : 2u;
	a.p.gx = w->cfg.gravity[0];
	a.p.gy = w->cfg.gravity[1];
	a.p.gz = w->cfg.gravity[2];
	if (substeps < 1) substeps = 1;
	a.p.substeps = substeps;
	a.p.h = dt / (float)substeps;
	a.busy_list = a.busy_count = nullptr;
	a.busy_flag = nullptr;
	a.next_list = a.next_count = nullptr;
	a.next_flag = nullptr;
	a.cur_flag = nullptr;
	a.next_above = 0;
	// The widest tile that still puts every world on the machine at once: the tick is a chain of dependent instructions
	// per world, and a world that has a warp to itself (no other world's branches in its instruction stream) runs that
	// chain fastest.  148 SMs x 7 resident one-warp blocks: up to 1036 worlds get 32 lanes each (one launch, no routing),
	// up to 2072 get 16, more get 8 (four worlds per warp).
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
	const uint32_t one_wave = (uint32_t)sms * 7u;
	const char *force_tile = getenv("GPX_TILE");  // experiments and tests: 8 / 16 / 32 (read per launch)
	const uint32_t forced = force_tile ? (uint32_t)atoi(force_tile) : 0u;
	// worlds of more than 32 bodies: a block per world (one manifold per thread up to 256 manifolds)
	const bool no_block = getenv("GPX_NO_BLOCK_TILE") != nullptr;  // read per launch: tests switch it
	if (w->cap > 32 && !no_block && world_smem_bytes(256, w->cap, w->cap_m) <= 200u * 1024u) return launch_tick_t<256>(w, a, w->stream, w->W);
	if (w->cap > 16 || forced == 32u || (!forced && w->W <= one_wave)) return launch_tick_t<32>(w, a, w->stream, w->W);
	// Ensembles of small worlds: the narrow launch takes every world whose previous tick fitted its lanes, a 32-lane
	// launch on a second stream takes the rest.  Each world wrote its own routing at the end of its previous tick
	// (route_next; two sets of list / count / flags, used alternately), so each world runs exactly once and no
	// classification kernel sits in front of the two launches.
	const uint32_t tile = forced == 8u || forced == 16u ? (w->cap <= 8 ? forced : 16u) : ((w->cap <= 8 && w->W > 2u * one_wave) ? 8u : 16u);
	int rc;
	const uint32_t cur = w->busy_cur, nxt = cur ^ 1u;
	w->busy_cur = nxt;
	GPX_CUDA(cudaMemsetAsync(w->d_busy_n + nxt, 0, sizeof(uint32_t), w->stream));
	a.next_list = w->d_busy + (size_t)nxt * w->W;
	a.next_count = w->d_busy_n + nxt;
	a.next_flag = w->d_busy_flag + (size_t)nxt * w->W;
	a.next_above = getenv("GPX_DEBUG_NO_ROUTE") ? 0x7FFFFFFFu : tile;  // debugging: nobody is ever handed to the 32-lane launch
	a.cur_flag = w->d_busy_flag + (size_t)cur * w->W;
	GPX_CUDA(cudaEventRecord(w->ev_fork, w->stream));
	GPX_CUDA(cudaStreamWaitEvent(w->stream2, w->ev_fork, 0));
	TickArgs b = a;
	if (getenv("GPX_PHASE_BUSY_ONLY")) a.phase_cycles = nullptr;  // debugging: phase counters of the 32-lane launch alone
	b.busy_list = w->d_busy + (size_t)cur * w->W;
	b.busy_count = w->d_busy_n + cur;
	if ((rc = launch_tick_t<32>(w, b, w->stream2, w->W < MAX_BUSY_WORLDS ? w->W : MAX_BUSY_WORLDS)) != GPX_OK) return rc;
	GPX_CUDA(cudaEventRecord(w->ev_join, w->stream2));
	a.busy_flag = w->d_busy_flag + (size_t)cur * w->W;
	rc = tile == 8u ? launch_tick_t<8>(w, a, w->stream, w->W) : launch_tick_t<16>(w, a, w->stream, w->W);
	GPX_CUDA(cudaStreamWaitEvent(w->stream, w->ev_join, 0));
	return rc;
}

int launch_sleep_test(gpx_world *w, float dt)
{
	k_sleep<<<(w->W + SLEEP_WARPS - 1) / SLEEP_WARPS, SLEEP_WARPS * 32, 0, w->stream>>>(w->bs, w->mc, w->W, w->cap, w->cap_m, dt);
	count_launch();
	GPX_CUDA(cudaGetLastError());
	return GPX_OK;
}

int launch_apply_commands(gpx_world *w, const BodyCommand *d_cmd, uint32_t n)
{
	if (n == 0) return GPX_OK;
	k_apply_commands<<<(n + 127) / 128, 128, 0, w->stream>>>(w->bs, d_cmd, n);
	count_launch();
	GPX_CUDA(cudaGetLastError());
	return GPX_OK;
}

int launch_stats(gpx_world *w)
{
	const uint32_t threads = 128;
	const uint32_t grid = (w->W * 32u + threads - 1) / threads;
	k_stats<<<grid, threads, 0, w->stream>>>(w->bs, w->mc, w->d_err, w->W, w->cap, w->ticks, w->d_stats);
	count_launch();
	GPX_CUDA(cudaGetLastError());
	return GPX_OK;
}

}  // namespace gpx
