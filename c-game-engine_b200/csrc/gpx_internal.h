// gpx_internal.h — host-side world object and kernel entry points shared by the libgpx translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/gpx.h"

namespace gpx {

constexpr uint32_t STATIC_BODY_BASE = 0x400000u;  // ids >= this name static collision meshes (shared by all worlds)
constexpr float RAY_MISS_FRACTION = 2.0f;
constexpr uint32_t CHARACTER_BODY_ID = 0x3FFFFFu;  // pseudo body id of a world's player character in contact events
constexpr uint32_t CHARACTER_MAX_CONTACTS = 64;
constexpr int GPX_MAX_DEVICES = 64;  // per-device bookkeeping of function attributes
constexpr float BVH_PAD = 1.0e-3f;  // node boxes are padded; leaves are re-tested exactly

// body flag word
constexpr uint32_t BF_ALIVE = 1u;
constexpr uint32_t BF_SHAPE_SHIFT = 1;   // 3 bits
constexpr uint32_t BF_MOTION_SHIFT = 4;  // 2 bits
constexpr uint32_t BF_LAYER_SHIFT = 6;   // 2 bits
constexpr uint32_t BF_SENSOR = 1u << 8;
constexpr uint32_t BF_DOF_SHIFT = 9;     // 6 bits
constexpr uint32_t BF_ALLOW_SLEEP = 1u << 15;
constexpr uint32_t BF_RAYFLAG_SHIFT = 16;  // 8 bits
constexpr uint32_t BF_ASLEEP = 1u << 24;      // a dynamic body that is asleep: static for everything the tick does
constexpr uint32_t BF_KIN_MOVING = 1u << 25;  // work records only: a kinematic body with a non-zero velocity
constexpr uint32_t BF_WAKE_MARK = 1u << 26;   // work records only (wide worlds): a sleeper touched by an active body

// Structure-of-arrays body store in HBM; index = world * cap + slot.  Every array is 16-byte vectorised.
struct BodyStore
{
	float4 *pos;    // xyz, w unused
	float4 *quat;   // xyzw
	float4 *lin;    // linear velocity xyz
	float4 *ang;    // angular velocity xyz
	float4 *prop0;  // inv mass, local inverse inertia diagonal xyz
	float4 *prop1;  // half extents xyz (sphere: radius in x), friction
	float4 *prop2;  // linear damping, angular damping, gravity factor, restitution
	uint32_t *flags;
	float4 *sleep_c;  // 3 per body: sleep-test sphere (centre xyz, radius w)
	float *sleep_t;   // time the test points stayed inside their spheres; < 0: spheres not set
};

// Contact cache carried between sub-steps and ticks (warm starting); index = world * cap_m + m
struct ManifoldCache
{
	uint4 *key;       // a, b, np, 0
	float4 *p1;       // 4 per manifold: local point on a, w = normal lambda
	float4 *p2;       // 4 per manifold: local point on b (static: world); w of the first three = the manifold's friction
	                  // impulse (tangent 1, tangent 2, twist)
	uint32_t *count;  // per world
};

struct StaticDevice
{
	uint32_t n_tris = 0, n_nodes = 0;
	float4 *tri = nullptr;    // 4 float4 per triangle in LBVH order: (a, orig index) (b, static body) (c, friction) (n, ray flags)
	float4 *nodes = nullptr;  // 4 float4 per internal node: c0 xy bounds, c1 xy bounds, both z bounds, child indices
	// the rays' tree: same record formats over split references (large triangles cut into several leaves that all
	// point at the original triangle); aliases the tree above when nothing was split
	uint32_t n_ray_leaves = 0, n_ray_nodes = 0;
	float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};  // bounds of the static geometry (rays: coherence bins)
	uint32_t ray_depth = 64;  // deepest leaf of the rays' tree (64: unknown, a radix tree's bound)
	float4 *ray_tri = nullptr, *ray_nodes = nullptr;
};
constexpr uint32_t RAY_TREE_MAX_HOST_BUILD = 1u << 18;  // triangles up to which the rays' tree gets its SAH topology on the host
constexpr uint32_t RAY_TREE_MAX_LEAVES = 1000;  // (2 * 1000 - 1) * 64 B = 128 KB in one SM's shared memory; measured optimum
                                                // on shapes.gmap (850: 3.60, 1000: 3.83, 1250: 3.67 G rays/s)

// player character of one world (gpx_char.cu)
struct CharDev
{
	float px, py, pz, hh;
	float vx, vy, vz, r;
	float gnx, gny, gnz, cos_slope;
	uint32_t alive, ground, ground_body, pad;
};

struct BodyCommand  // host -> device write, applied by k_apply_commands before the next step
{
	uint32_t index;
	uint32_t mask;  // 1 pos, 2 quat, 4 lin, 8 ang, 16 props+flags, 32 ray-flag byte of the flag word
	uint32_t flags;
	uint32_t pad;
	float4 pos, quat, lin, ang, prop0, prop1, prop2;
};

struct StaticBodyHost
{
	gpx_transform xfm;
	float friction;
	uint32_t ray_flags;
	uint64_t user_data;
	uint32_t first, count;
};

struct TickParams
{
	uint32_t worlds, cap, cap_m;
	uint32_t vel_steps, pos_steps;
	float gx, gy, gz;
	float h;  // sub-step
	int substeps;
};

}  // namespace gpx

namespace gpx { struct WideDevice; }

struct gpx_world
{
	gpx_world_config cfg{};
	int device = 0;
	cudaStream_t stream = nullptr, stream2 = nullptr;  // stream2: the 32-lane launch for busy worlds (gpx_tick.cu)
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	// gpx_raycast_batch_async: the hits travel back on their own stream while the tick runs on `stream`
	cudaStream_t stream_copy = nullptr;
	cudaEvent_t ev_rays_done = nullptr, ev_hits_done = nullptr;
	bool hits_pending = false;
	// gpx_raycast_batch on large batches: chunks pipelined over three streams (H2D | kernel | D2H), one event pair per chunk
	cudaStream_t stream_h2d = nullptr;
	cudaEvent_t ev_pipe[32] = {};
	uint32_t *d_busy = nullptr, *d_busy_n = nullptr;
	uint8_t *d_busy_flag = nullptr;
	uint32_t busy_cur = 0;  // which of the two routing sets (list / count / flags) the coming tick reads
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	uint32_t W = 0, cap = 0, cap_m = 0;

	// static soup (host staging, world space) + device LBVH
	std::vector<float> h_tris;
	std::vector<uint32_t> h_tri_body;
	std::vector<gpx::StaticBodyHost> sbodies;
	bool static_dirty = false;
	gpx::StaticDevice sd;

	// bodies
	gpx::BodyStore bs{};
	gpx::ManifoldCache mc{};
	std::vector<uint32_t> free_hint;    // per world: no free slot below this index
	std::vector<uint32_t> h_flags;      // host shadow of the flag word (slot allocation, getters)
	std::vector<uint64_t> h_user_data;  // Actor* per body
	std::vector<gpx::BodyCommand> pending;
	gpx::BodyCommand *d_cmd = nullptr;
	size_t d_cmd_cap = 0;
	std::mutex mu;  // guards pending/h_flags: create/destroy/set arrive from several engine threads

	// host mirror (pinned) served to getters
	// transforms are double-buffered: a readback fills the back pair, then mirror_gen is bumped (front = mirror_gen & 1), so
	// getters on other threads (render, LOD) never see a half-written tick
	float4 *mb_pos[2] = {nullptr, nullptr}, *mb_quat[2] = {nullptr, nullptr};
	// the same two buffers as the device sees them (nullptr: not mapped).  The ensemble tick writes each world's final
	// positions and orientations straight into the back buffer as the world finishes, so gpx_sync_transforms has only the
	// error word left to fetch; mirror_direct says the back buffer holds the state of the last thing enqueued.
	float4 *mb_dev[2] = {nullptr, nullptr};
	bool mirror_direct = false;
	// direct writes cost the tick a few microseconds of bus traffic, so they are only asked for when the caller has been
	// reading every tick back (the last step was followed by a synchronisation)
	bool synced_since_step = false;
	uint8_t *d_mirror_fresh = nullptr;  // per world: bit b = mirror buffer b holds the world's current state (kernel-written)
	std::atomic<uint32_t> mirror_gen{0};
	float4 *m_lin = nullptr, *m_ang = nullptr;
	uint32_t *d_err = nullptr;  // [0] = OR of per-world tick errors
	uint32_t *m_err = nullptr;  // pinned
	gpx_world_stats *d_stats = nullptr;
	// contact events (gpx_events_enable)
	unsigned long long *d_ev_prev = nullptr;
	uint32_t *d_ev_nprev = nullptr, *d_ev_count = nullptr;
	uint4 *d_ev_out = nullptr;
	std::vector<uint4> h_ev_out;
	std::vector<uint32_t> h_ev_count;
	// player characters (gpx_char.cu): one per world, contact keys for the event pass
	gpx::CharDev *d_ch = nullptr;
	unsigned long long *d_ch_keys = nullptr;
	uint32_t *d_ch_nkeys = nullptr;
	std::vector<gpx::CharDev> h_ch;
	float4 *d_park = nullptr;  // 31 float4 per manifold slot (gpx_tick.cu, worlds with more manifolds than lanes)
	gpx::WideDevice *wide = nullptr;  // non-null: this world runs the wide-world kernels
	uint4 *d_cand = nullptr;  // static-candidate cache, 8 x uint4 per body (gpx_tick.cu)
	unsigned long long *d_phase = nullptr;  // 16 counters, allocated by gpx_debug_phase_cycles(enable)
	uint32_t ticks = 0;
	bool sleep_enabled = false;  // some body was created with allow_sleeping: run the per-tick sleep test

	// ray staging
	void *d_rays = nullptr, *d_hits = nullptr;
	size_t ray_cap = 0;
	void *d_capq = nullptr, *d_capo = nullptr;  // staging of gpx_overlap_capsule_batch, grown on demand
	size_t capq_cap = 0;
};

namespace gpx {
// gpx_bvh.cu
int build_static(gpx_world *w);
void bitonic_sort_u64(unsigned long long *d_keys, uint32_t n_pad, cudaStream_t st);
void exclusive_scan_u32(uint32_t *d, uint32_t n, cudaStream_t st);
void radix_sort_u64(unsigned long long *d_keys, unsigned long long *tmp, uint32_t *ghist, uint32_t n, uint32_t first_bit,
					cudaStream_t st, const uint32_t *only_if = nullptr);
uint32_t next_pow2(uint32_t v);
// gpx_wide.cu: the tick of ONE large world (more than 64 bodies), state in global memory
struct WideDevice;
int wide_create(gpx_world *w);
void wide_destroy(gpx_world *w);
int launch_wide_tick(gpx_world *w, float dt, int substeps);
int wide_events_enable(gpx_world *w, bool enable);
uint32_t wide_event_capacity(const gpx_world *w);
int wide_counters(gpx_world *w, uint32_t *out8);
int wide_stats(gpx_world *w);  // after launch_stats: adds the wide world's manifold counts to d_stats
// gpx_rays.cu
int launch_raycast(gpx_world *w, const void *d_rays, uint64_t n, void *d_hits);
int launch_spherecast(gpx_world *w, const void *d_casts, uint64_t n, void *d_hits);
// gpx_char.cu
int launch_character(gpx_world *w, float dt, const gpx_character_update_settings *cfg);
int launch_overlap_capsules(gpx_world *w, const void *d_queries, uint64_t n, void *d_out);
// gpx_tick.cu
int launch_tick(gpx_world *w, float dt, int substeps);
int launch_sleep_test(gpx_world *w, float dt);
int launch_sleep_test(gpx_world *w, float dt);
int launch_apply_commands(gpx_world *w, const BodyCommand *d_cmd, uint32_t n);
int launch_stats(gpx_world *w);
// gpx_api.cu
void set_error(const char *what, cudaError_t e);
void count_launch(uint64_t n = 1);
}  // namespace gpx

#define GPX_CUDA(call)                                         \
	do                                                         \
	{                                                          \
		cudaError_t e_ = (call);                               \
		if (e_ != cudaSuccess)                                 \
		{                                                      \
			gpx::set_error(#call, e_);                         \
			return GPX_ERR_CUDA;                               \
		}                                                      \
	} while (0)
