// gpx_api.cu — the C ABI of include/gpx.h: world/body management, host mirror, command queue.
// Host logic only; every compute entry point ends in a kernel launch from the other translation units.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>

#include <zlib.h>

#include "gpx_internal.h"
#include "gpx_math.cuh"

namespace gpx {

static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};
static int g_device = -1;

static inline float4 *front_pos(const gpx_world *w) { return w->mb_pos[w->mirror_gen.load(std::memory_order_acquire) & 1u]; }
static inline float4 *front_quat(const gpx_world *w) { return w->mb_quat[w->mirror_gen.load(std::memory_order_acquire) & 1u]; }

void set_error(const char *what, cudaError_t e)
{
	g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
}
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static uint32_t pack_flags(const gpx_body_desc &d)
{
	uint32_t f = BF_ALIVE;
	f |= (d.shape & 7u) << BF_SHAPE_SHIFT;
	f |= (d.motion_type & 3u) << BF_MOTION_SHIFT;
	f |= (d.layer & 3u) << BF_LAYER_SHIFT;
	if (d.is_sensor) f |= BF_SENSOR;
	f |= ((d.allowed_dofs ? d.allowed_dofs : 63u) & 63u) << BF_DOF_SHIFT;
	if (d.allow_sleeping) f |= BF_ALLOW_SLEEP;
	f |= (d.ray_flags & 0xFFu) << BF_RAYFLAG_SHIFT;
	return f;
}

// JPH_MassProperties override + CalculateInertia (game/src/actor/prop/Physbox.c:27-32): the shape's inertia scaled
// to the requested mass; density 1000 when no mass is given.
static void mass_properties(const gpx_body_desc &d, float &inv_mass, float inv_i[3])
{
	inv_mass = 0.0f;
	inv_i[0] = inv_i[1] = inv_i[2] = 0.0f;
	if (d.motion_type != GPX_MOTION_DYNAMIC || d.shape == GPX_SHAPE_EMPTY) return;
	float m, ix, iy, iz;
	if (d.shape == GPX_SHAPE_BOX)
	{
		float sx = 2.0f * d.half_extents[0], sy = 2.0f * d.half_extents[1], sz = 2.0f * d.half_extents[2];
		float vol = (sx * sy) * sz;
		m = d.mass > 0.0f ? d.mass : 1000.0f * vol;
		float k = m / 12.0f;
		ix = k * ((sy * sy) + (sz * sz));
		iy = k * ((sx * sx) + (sz * sz));
		iz = k * ((sx * sx) + (sy * sy));
	}
	else
	{
		float r = d.half_extents[0];
		float vol = (4.18879020f * r) * (r * r);
		m = d.mass > 0.0f ? d.mass : 1000.0f * vol;
		ix = iy = iz = (0.4f * m) * (r * r);
	}
	inv_mass = 1.0f / m;
	inv_i[0] = 1.0f / ix;
	inv_i[1] = 1.0f / iy;
	inv_i[2] = 1.0f / iz;
}

static BodyCommand command_from_desc(uint32_t index, const gpx_body_desc &d)
{
	BodyCommand c;
	memset(&c, 0, sizeof(c));
	c.index = index;
	c.mask = 31u;
	c.flags = pack_flags(d);
	c.pos = make_float4(d.position[0], d.position[1], d.position[2], 0.0f);
	q4 q;
	q.x = d.rotation[0]; q.y = d.rotation[1]; q.z = d.rotation[2]; q.w = d.rotation[3];
	q = qnormalize(q);
	c.quat = make_float4(q.x, q.y, q.z, q.w);
	c.lin = make_float4(d.linear_velocity[0], d.linear_velocity[1], d.linear_velocity[2], 0.0f);
	c.ang = make_float4(d.angular_velocity[0], d.angular_velocity[1], d.angular_velocity[2], 0.0f);
	float im, ii[3];
	mass_properties(d, im, ii);
	c.prop0 = make_float4(im, ii[0], ii[1], ii[2]);
	c.prop1 = make_float4(d.half_extents[0], d.half_extents[1], d.half_extents[2], d.friction);
	c.prop2 = make_float4(d.linear_damping, d.angular_damping, d.gravity_factor, d.restitution);
	return c;
}

// Push queued host writes to the device (one H2D + one scatter kernel).  Caller holds w->mu.
static int flush_commands(gpx_world *w)
{
	if (w->pending.empty()) return GPX_OK;
	// The scatter kernel applies commands in parallel, so several writes to one body (destroy + create of a reused slot,
	// two setters) are folded into one command here, later fields winning, to keep the order the engine issued them in.
	{
		std::unordered_map<uint32_t, size_t> at;
		std::vector<BodyCommand> merged;
		merged.reserve(w->pending.size());
		for (const BodyCommand &c : w->pending)
		{
			auto it = at.find(c.index);
			if (it == at.end())
			{
				at.emplace(c.index, merged.size());
				merged.push_back(c);
				continue;
			}
			BodyCommand &e = merged[it->second];
			if (c.mask & 1u) e.pos = c.pos;
			if (c.mask & 2u) e.quat = c.quat;
			if (c.mask & 4u) e.lin = c.lin;
			if (c.mask & 8u) e.ang = c.ang;
			if (c.mask & 16u)
			{
				e.prop0 = c.prop0;
				e.prop1 = c.prop1;
				e.prop2 = c.prop2;
				e.flags = c.flags;
			}
			if (c.mask & 32u) e.flags = (e.flags & ~(0xFFu << BF_RAYFLAG_SHIFT)) | (c.flags & (0xFFu << BF_RAYFLAG_SHIFT));
			// (64 = wake carries no payload)
			e.mask |= c.mask;
		}
		w->pending.swap(merged);
	}
	const size_t n = w->pending.size();
	if (n > w->d_cmd_cap)
	{
		if (w->d_cmd) cudaFree(w->d_cmd);
		w->d_cmd_cap = n * 2;
		GPX_CUDA(cudaMalloc(&w->d_cmd, sizeof(BodyCommand) * w->d_cmd_cap));
	}
	GPX_CUDA(cudaMemcpyAsync(w->d_cmd, w->pending.data(), sizeof(BodyCommand) * n, cudaMemcpyHostToDevice, w->stream));
	int rc = launch_apply_commands(w, w->d_cmd, (uint32_t)n);
	w->mirror_direct = false;  // the device state has moved on from what the last tick wrote into the mirror
	// the staging vector is pageable: wait so it can be reused
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	w->pending.clear();
	return rc;
}

template <typename T>
static int dalloc(T **p, size_t n)
{
	GPX_CUDA(cudaMalloc(p, sizeof(T) * n));
	GPX_CUDA(cudaMemset(*p, 0, sizeof(T) * n));
	return GPX_OK;
}

}  // namespace gpx

namespace {
struct Cursor
{
	const uint8_t *d;
	uint64_t n, o = 0;
	bool bad = false;
	int depth = 0;  // nesting of arrays / key-value lists while skipping a parameter (hostile files recurse for ever)
	template <typename T>
	T get()
	{
		T v{};
		if (o + sizeof(T) > n) { bad = true; return v; }
		memcpy(&v, d + o, sizeof(T));
		o += sizeof(T);
		return v;
	}
	void skip(uint64_t k)
	{
		if (o + k > n || o + k < o) bad = true;
		else o += k;
	}
	void skip_string() { uint64_t l = get<uint64_t>(); if (!bad) skip(l); }
	void skip_param();
	void skip_kvlist()
	{
		uint64_t k = get<uint64_t>();
		for (uint64_t i = 0; i < k && !bad; i++) { skip_string(); skip_param(); }
	}
};
void Cursor::skip_param()
{
	if (++depth > 32)
	{
		bad = true;
		depth--;
		return;
	}
	struct Leave { int &d; ~Leave() { d--; } } leave{depth};
	uint8_t t = get<uint8_t>();
	switch (t)
	{
		case 0: case 3: skip(1); break;            /* byte, bool */
		case 1: case 2: skip(4); break;            /* int, float */
		case 4: skip_string(); break;
		case 6: skip(16); break;                   /* colour */
		case 7: skip_kvlist(); break;
		case 8: { uint64_t k = get<uint64_t>(); for (uint64_t i = 0; i < k && !bad; i++) skip_param(); break; }
		case 9: skip(8); break;
		case 10: skip(8); break;
		case 11: skip(12); break;
		default: break;
	}
}
}  // namespace

using namespace gpx;

extern "C" {

int gpx_init(int device)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0)
	{
		set_error("gpx_init: no CUDA device (this library has no CPU path)", e == cudaSuccess ? cudaErrorNoDevice : e);
		return -GPX_ERR_CUDA;
	}
	if (device < 0 || device >= n) return -GPX_ERR_INVALID_ARG;
	e = cudaSetDevice(device);
	if (e != cudaSuccess)
	{
		set_error("cudaSetDevice", e);
		return -GPX_ERR_CUDA;
	}
	g_device = device;
	return GPX_ABI_VERSION;
}

void gpx_shutdown(void) { g_device = -1; }

const char *gpx_last_error(void) { return g_last_error.c_str(); }

uint64_t gpx_launch_count(void) { return g_launches.load(); }

gpx_world *gpx_world_create(const gpx_world_config *cfg)
{
	if (!cfg || cfg->worlds == 0 || cfg->max_bodies_per_world == 0) return nullptr;
	// up to 64 bodies per world: ensembles (one tile of lanes per world, gpx_tick.cu).  More: ONE wide world (gpx_wide.cu)
	const bool wide = cfg->max_bodies_per_world > 64 || (cfg->flags & GPX_WORLD_WIDE);
	if (wide && (uint64_t)cfg->worlds * cfg->max_bodies_per_world > (1ull << 20)) return nullptr;
	if (cudaSetDevice(cfg->device) != cudaSuccess) return nullptr;
	gpx_world *w = new gpx_world();
	w->cfg = *cfg;
	w->device = cfg->device;
	w->W = cfg->worlds;
	w->cap = cfg->max_bodies_per_world;
	w->cap_m = cfg->max_manifolds_per_world ? cfg->max_manifolds_per_world
										 : (wide ? w->cap * 8u : (w->cap * 3u < 16u ? 16u : w->cap * 3u));
	if (w->cap_m & 1u) w->cap_m++;
	const size_t nb = (size_t)w->W * w->cap, nm = (size_t)w->W * w->cap_m;
	bool ok = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) == cudaSuccess;
	ok = ok && cudaEventCreate(&w->ev0) == cudaSuccess && cudaEventCreate(&w->ev1) == cudaSuccess;
	ok = ok && cudaStreamCreateWithFlags(&w->stream_copy, cudaStreamNonBlocking) == cudaSuccess &&
		 cudaEventCreateWithFlags(&w->ev_rays_done, cudaEventDisableTiming) == cudaSuccess &&
		 cudaEventCreateWithFlags(&w->ev_hits_done, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaStreamCreateWithFlags(&w->stream2, cudaStreamNonBlocking) == cudaSuccess &&
		 cudaEventCreateWithFlags(&w->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
		 cudaEventCreateWithFlags(&w->ev_join, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && dalloc(&w->d_busy, 2 * (size_t)cfg->worlds) == GPX_OK && dalloc(&w->d_busy_n, 2) == GPX_OK &&
		 dalloc(&w->d_busy_flag, 2 * (size_t)cfg->worlds) == GPX_OK;  // two routing sets, see launch_tick
	// positions, orientations and the error words share ONE allocation (and so do their host mirrors): what
	// gpx_sync_transforms reads back every tick is a single contiguous copy
	{
		const size_t pose_bytes = 2 * sizeof(float4) * nb + sizeof(uint32_t) * ((size_t)w->W + 1);
		unsigned char *pose = nullptr;
		ok = ok && cudaMalloc(&pose, pose_bytes) == cudaSuccess && cudaMemset(pose, 0, pose_bytes) == cudaSuccess;
		w->bs.pos = reinterpret_cast<float4 *>(pose);
		if (ok)
		{
			w->bs.quat = w->bs.pos + nb;
			w->d_err = reinterpret_cast<uint32_t *>(pose + 2 * sizeof(float4) * nb);
		}
	}
	ok = ok && dalloc(&w->bs.lin, nb) == GPX_OK &&
		 dalloc(&w->bs.ang, nb) == GPX_OK && dalloc(&w->bs.prop0, nb) == GPX_OK && dalloc(&w->bs.prop1, nb) == GPX_OK &&
		 dalloc(&w->bs.prop2, nb) == GPX_OK && dalloc(&w->bs.flags, nb) == GPX_OK && dalloc(&w->bs.sleep_c, 3 * nb) == GPX_OK &&
		 dalloc(&w->bs.sleep_t, nb) == GPX_OK;
	ok = ok && dalloc(&w->mc.count, (size_t)w->W) == GPX_OK;
	if (!wide)
		ok = ok && dalloc(&w->mc.key, nm) == GPX_OK && dalloc(&w->mc.p1, nm * 4) == GPX_OK && dalloc(&w->mc.p2, nm * 4) == GPX_OK &&
			 dalloc(&w->d_park, nm * 31) == GPX_OK;
	else
		ok = ok && wide_create(w) == GPX_OK;
	ok = ok && cudaMalloc(&w->d_cand, sizeof(uint4) * 8 * nb) == cudaSuccess &&
		 cudaMemset(w->d_cand, 0xFF, sizeof(uint4) * 8 * nb) == cudaSuccess;
	ok = ok && dalloc(&w->d_stats, (size_t)w->W) == GPX_OK;
	ok = ok && cudaMallocHost(&w->mb_pos[0], sizeof(float4) * (2 * nb + 1)) == cudaSuccess &&
		 cudaMallocHost(&w->mb_pos[1], sizeof(float4) * (2 * nb + 1)) == cudaSuccess &&
		 cudaMallocHost(&w->m_lin, sizeof(float4) * nb) == cudaSuccess &&
		 cudaMallocHost(&w->m_ang, sizeof(float4) * nb) == cudaSuccess &&
		 cudaMallocHost(&w->m_err, sizeof(uint32_t) * 4 + sizeof(gpx_ray) + sizeof(gpx_hit)) == cudaSuccess;  // + one ray, one hit
	if (!ok)
	{
		set_error("gpx_world_create", cudaGetLastError());
		gpx_world_destroy(w);
		return nullptr;
	}
	for (int k = 0; k < 2; k++)
	{
		memset(w->mb_pos[k], 0, sizeof(float4) * (2 * nb + 1));
		w->mb_quat[k] = w->mb_pos[k] + nb;  // the error word follows at [2 * nb]
		if (k == 0 && !w->d_mirror_fresh && cudaMalloc(&w->d_mirror_fresh, w->W) == cudaSuccess) cudaMemset(w->d_mirror_fresh, 0, w->W);
		void *dev = nullptr;
		if (!getenv("GPX_NO_DIRECT_MIRROR") && w->d_mirror_fresh && cudaHostGetDevicePointer(&dev, w->mb_pos[k], 0) == cudaSuccess)
			w->mb_dev[k] = reinterpret_cast<float4 *>(dev);
		else
			cudaGetLastError();
	}
	memset(w->m_lin, 0, sizeof(float4) * nb);
	memset(w->m_ang, 0, sizeof(float4) * nb);
	w->m_err[0] = 0;
	w->free_hint.assign(w->W, 0u);
	w->h_flags.assign(nb, 0u);
	w->h_user_data.assign(nb, 0ull);
	return w;
}

void gpx_world_destroy(gpx_world *w)
{
	if (!w) return;
	cudaSetDevice(w->device);
	if (w->stream) cudaStreamSynchronize(w->stream);
	if (w->stream_copy) cudaStreamSynchronize(w->stream_copy);
	wide_destroy(w);
	cudaFree(w->bs.pos); /* + quat, d_err: one allocation */ cudaFree(w->bs.lin); cudaFree(w->bs.ang);
	cudaFree(w->bs.prop0); cudaFree(w->bs.prop1); cudaFree(w->bs.prop2); cudaFree(w->bs.flags);
	cudaFree(w->bs.sleep_c); cudaFree(w->bs.sleep_t);
	cudaFree(w->mc.key); cudaFree(w->mc.p1); cudaFree(w->mc.p2); cudaFree(w->mc.count);
	cudaFree(w->d_stats); cudaFree(w->d_cmd); if (w->sd.ray_tri != w->sd.tri) cudaFree(w->sd.ray_tri);
	if (w->sd.ray_nodes != w->sd.nodes) cudaFree(w->sd.ray_nodes);
	cudaFree(w->sd.tri); cudaFree(w->sd.nodes);
	cudaFree(w->d_rays); cudaFree(w->d_hits); cudaFree(w->d_capq); cudaFree(w->d_capo); cudaFree(w->d_phase); cudaFree(w->d_cand); cudaFree(w->d_park);
	cudaFree(w->d_ev_prev); cudaFree(w->d_ev_nprev); cudaFree(w->d_ev_count); cudaFree(w->d_ev_out);
	cudaFree(w->d_ch); cudaFree(w->d_ch_keys); cudaFree(w->d_ch_nkeys);
	cudaFree(w->d_mirror_fresh);
	cudaFreeHost(w->mb_pos[0]); cudaFreeHost(w->mb_pos[1]); cudaFreeHost(w->m_lin); cudaFreeHost(w->m_ang); cudaFreeHost(w->m_err);
	if (w->ev0) cudaEventDestroy(w->ev0);
	if (w->ev1) cudaEventDestroy(w->ev1);
	if (w->ev_fork) cudaEventDestroy(w->ev_fork);
	if (w->ev_join) cudaEventDestroy(w->ev_join);
	if (w->stream2) cudaStreamDestroy(w->stream2);
	if (w->ev_rays_done) cudaEventDestroy(w->ev_rays_done);
	if (w->ev_hits_done) cudaEventDestroy(w->ev_hits_done);
	if (w->stream_copy) cudaStreamDestroy(w->stream_copy);
	if (w->stream_h2d) cudaStreamDestroy(w->stream_h2d);
	for (cudaEvent_t e : w->ev_pipe)
		if (e) cudaEventDestroy(e);
	cudaFree(w->d_busy); cudaFree(w->d_busy_n); cudaFree(w->d_busy_flag);
	if (w->stream) cudaStreamDestroy(w->stream);
	delete w;
}

/* ---- static geometry */

int gpx_static_add_mesh(gpx_world *w, const gpx_transform *xfm, const float *tris, uint64_t ntris, float friction,
						uint64_t user_data, uint32_t *out_body)
{
	if (!w || !xfm || (!tris && ntris)) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	if (w->cfg.max_static_triangles && w->h_tri_body.size() + ntris > w->cfg.max_static_triangles) return GPX_ERR_CAPACITY;
	StaticBodyHost sb;
	sb.xfm = *xfm;
	sb.friction = friction;
	sb.ray_flags = GPX_BODY_BLOCKS_LASERS;  // map geometry has no actor: lasers stop on it (Laser.c:74-85)
	sb.user_data = user_data;
	sb.first = (uint32_t)w->h_tri_body.size();
	sb.count = (uint32_t)ntris;
	const uint32_t k = (uint32_t)w->sbodies.size();
	w->sbodies.push_back(sb);
	q4 q;
	q.x = xfm->rotation[0]; q.y = xfm->rotation[1]; q.z = xfm->rotation[2]; q.w = xfm->rotation[3];
	const v3 p = V(xfm->position[0], xfm->position[1], xfm->position[2]);
	for (uint64_t i = 0; i < ntris; i++)
		for (int v = 0; v < 3; v++)
		{
			const float *t = tris + i * 9 + v * 3;
			v3 pw = qrot(q, V(t[0], t[1], t[2])) + p;
			w->h_tris.push_back(pw.x);
			w->h_tris.push_back(pw.y);
			w->h_tris.push_back(pw.z);
		}
	w->h_tri_body.insert(w->h_tri_body.end(), ntris, k);
	w->static_dirty = true;
	if (out_body) *out_body = STATIC_BODY_BASE + k;
	return GPX_OK;
}

int gpx_static_commit(gpx_world *w)
{
	if (!w) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	return build_static(w);
}

int gpx_static_remove_mesh(gpx_world *w, uint32_t body)
{
	if (!w || body < STATIC_BODY_BASE || body - STATIC_BODY_BASE >= w->sbodies.size()) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	StaticBodyHost &sb = w->sbodies[body - STATIC_BODY_BASE];
	if (sb.count == 0) return GPX_OK;
	const uint32_t first = sb.first, count = sb.count;
	w->h_tris.erase(w->h_tris.begin() + 9ull * first, w->h_tris.begin() + 9ull * (first + count));
	w->h_tri_body.erase(w->h_tri_body.begin() + first, w->h_tri_body.begin() + first + count);
	for (StaticBodyHost &o : w->sbodies)
		if (o.first > first) o.first -= count;
	sb.count = 0;
	w->static_dirty = true;
	return GPX_OK;
}

int gpx_static_info(const gpx_world *w, uint32_t *n_tris, uint32_t *n_nodes, uint32_t *n_bodies)
{
	if (!w) return GPX_ERR_INVALID_ARG;
	if (n_tris) *n_tris = w->sd.n_tris;
	if (n_nodes) *n_nodes = w->sd.n_nodes;
	if (n_bodies) *n_bodies = (uint32_t)w->sbodies.size();
	return GPX_OK;
}

/* Collision section of a decompressed .gmap (engine/src/assets/MapLoader.c:54-273): skip sky/strings/actors/models,
 * then numCollisionMeshes x { pos, subShapeCount x { numTris x 9 f32 } }; each mesh becomes one static body with
 * friction 4.25 (MapLoader.c:263). */

static int load_gmap_body(gpx_world *w, const uint8_t *body, uint64_t size);

int gpx_static_load_gmap(gpx_world *w, const uint8_t *body, uint64_t size)
{
	if (!w || !body) return -GPX_ERR_INVALID_ARG;
	// nothing may unwind through the C ABI: allocation failures on a corrupt file become an error code
	try { return load_gmap_body(w, body, size); }
	catch (...) { return -GPX_ERR_INVALID_ARG; }
}

static int load_gmap_body(gpx_world *w, const uint8_t *body, uint64_t size)
{
	Cursor c{body, size};
	if (c.get<uint8_t>()) c.skip_string();
	c.skip_string();
	c.skip_string();
	const uint64_t n_actors = c.get<uint64_t>();
	for (uint64_t i = 0; i < n_actors && !c.bad; i++)
	{
		c.skip_string();
		c.skip(24);
		const uint64_t n_conn = c.get<uint64_t>();
		for (uint64_t j = 0; j < n_conn && !c.bad; j++)
		{
			c.skip_string(); c.skip_string(); c.skip_string();
			if (c.get<uint8_t>()) c.skip_param();
			c.skip(8);
		}
		c.skip_kvlist();
	}
	const uint64_t n_models = c.get<uint64_t>();
	for (uint64_t i = 0; i < n_models && !c.bad; i++)
	{
		c.skip_string();
		c.skip((uint64_t)c.get<uint32_t>() * 28u);
		c.skip((uint64_t)c.get<uint32_t>() * 4u);
	}
	const uint64_t n_meshes = c.get<uint64_t>();
	int added = 0;
	std::vector<float> tris;
	for (uint64_t i = 0; i < n_meshes && !c.bad; i++)
	{
		gpx_transform xfm;
		memset(&xfm, 0, sizeof(xfm));
		xfm.rotation[3] = 1.0f;
		xfm.position[0] = c.get<float>();
		xfm.position[1] = c.get<float>();
		xfm.position[2] = c.get<float>();
		const uint64_t n_sub = c.get<uint64_t>();
		if (n_sub == 0) continue;  // MapLoader.c:213-216
		tris.clear();
		for (uint64_t j = 0; j < n_sub && !c.bad; j++)
		{
			const uint64_t nt = c.get<uint64_t>();
			if (c.bad || nt > (c.n - c.o) / 36u) { c.bad = true; break; }  // (no nt * 36: it wraps for a hostile count)
			const size_t at = tris.size();
			tris.resize(at + nt * 9u);
			memcpy(tris.data() + at, c.d + c.o, nt * 36u);
			c.skip(nt * 36u);
		}
		if (c.bad) break;
		int rc = gpx_static_add_mesh(w, &xfm, tris.data(), tris.size() / 9u, 4.25f, 0, nullptr);
		if (rc != GPX_OK) return -rc;
		added++;
	}
	if (c.bad) return -GPX_ERR_INVALID_ARG;
	return added;
}

/* The asset container around a map (engine/src/assets/AssetReader.c:150-257, AssetReader.h:15-17): 23-byte header
 * <u32 magic "GAME"><u8 version = 2><u8 type><u8 typeVersion><u64 rawSize><u64 gzSize>, then one gzip member.  The
 * same size checks as DecompressAsset; the decompressed body goes to gpx_static_load_gmap. */
// The asset container (AssetReader.c:150-257): header checks as DecompressAsset, then the one gzip member.
static int inflate_container(const uint8_t *blob, uint64_t size, uint8_t want_type, std::vector<uint8_t> &body)
{
	const uint64_t HEADER = 23;
	if (!blob || size < HEADER) return GPX_ERR_INVALID_ARG;
	uint32_t magic;
	memcpy(&magic, blob, 4);
	if (magic != 0x454D4147u || blob[4] != 2 || (want_type != 0xFF && blob[5] != want_type)) return GPX_ERR_INVALID_ARG;
	uint64_t raw_size, gz_size;
	memcpy(&raw_size, blob + 7, 8);
	memcpy(&gz_size, blob + 15, 8);
	if (size - HEADER != gz_size || raw_size >= (1ull << 32)) return GPX_ERR_INVALID_ARG;
	try { body.resize((size_t)raw_size); }
	catch (...) { return GPX_ERR_INVALID_ARG; }
	z_stream zs;
	memset(&zs, 0, sizeof(zs));
	if (inflateInit2(&zs, MAX_WBITS | 16) != Z_OK) return GPX_ERR_INVALID_ARG;
	zs.next_in = const_cast<Bytef *>(blob + HEADER);
	zs.avail_in = (uInt)gz_size;
	zs.next_out = body.data();
	zs.avail_out = (uInt)raw_size;
	const int zrc = inflate(&zs, Z_FINISH);
	const uint64_t got = zs.total_out;
	inflateEnd(&zs);
	return zrc == Z_STREAM_END && got == raw_size ? GPX_OK : GPX_ERR_INVALID_ARG;
}

int gpx_static_load_gmap_container(gpx_world *w, const uint8_t *blob, uint64_t size)
{
	if (!w || !blob) return -GPX_ERR_INVALID_ARG;
	std::vector<uint8_t> body;
	if (inflate_container(blob, size, 0xFF, body) != GPX_OK) return -GPX_ERR_INVALID_ARG;
	return gpx_static_load_gmap(w, body.data(), body.size());
}

// Walks a decompressed .gmdl up to its collision section (ModelLoader.c:66-151); the cursor is left on the hull /
// triangle counts.
static bool gmdl_seek_collision(Cursor &c, gpx_model_collision &m)
{
	const uint32_t n_mat = c.get<uint32_t>(), n_slot = c.get<uint32_t>(), n_skin = c.get<uint32_t>(), n_lod = c.get<uint32_t>();
	m.collision_type = c.get<uint8_t>();
	for (uint32_t i = 0; i < n_mat && !c.bad; i++)
	{
		c.skip_string();
		c.skip(16 + 4);  // colour, shader
	}
	c.skip(4ull * n_slot * n_skin);
	for (uint32_t i = 0; i < n_lod && !c.bad; i++)
	{
		c.skip(8);  // lod distance, squared distance
		const uint64_t nv = c.get<uint64_t>();
		if (c.bad || nv > (c.n - c.o) / 48u) { c.bad = true; break; }
		c.skip(nv * 48u);  // ModelVertex
		c.skip(4);         // total index count
		std::vector<uint32_t> counts;
		for (uint32_t j = 0; j < n_slot && !c.bad; j++) counts.push_back(c.get<uint32_t>());
		for (uint32_t cnt : counts) c.skip(4ull * cnt);
	}
	for (int k = 0; k < 3; k++) m.bb_origin[k] = c.get<float>();
	for (int k = 0; k < 3; k++) m.bb_extents[k] = c.get<float>();
	return !c.bad && m.collision_type <= 2;
}

int gpx_model_load_gmdl(const uint8_t *body, uint64_t size, float tolerance, gpx_model_collision *out)
{
	if (!body || !out) return GPX_ERR_INVALID_ARG;
	try
	{
		gpx_model_collision m;
		memset(&m, 0, sizeof(m));
		Cursor c{body, size};
		if (!gmdl_seek_collision(c, m)) return GPX_ERR_INVALID_ARG;
		m.exact = 1;
		if (m.collision_type == 2)
		{
			const uint64_t nh = c.get<uint64_t>();
			if (c.bad || nh > (c.n - c.o) / 20u) return GPX_ERR_INVALID_ARG;
			m.n_hulls = (uint32_t)nh;
			for (uint64_t h = 0; h < nh; h++)
			{
				const uint64_t np = c.get<uint64_t>();
				float off[3];
				for (int k = 0; k < 3; k++) off[k] = c.get<float>();
				if (c.bad || np > (c.n - c.o) / 12u) return GPX_ERR_INVALID_ARG;
				if (h < GPX_MODEL_MAX_HULLS)
				{
					std::vector<float> pts(3 * (size_t)np);
					memcpy(pts.data(), c.d + c.o, 12 * (size_t)np);
					m.hull_points[h] = np;
					if (np < 3 || gpx_shape_from_hull(pts.data(), np, tolerance, &m.hull[h]) != GPX_OK) return GPX_ERR_INVALID_ARG;
					for (int k = 0; k < 3; k++) m.hull[h].center[k] += off[k];  // AddShape2(settings, &offset, ...), ModelLoader.c:336
					if (!m.hull[h].exact) m.exact = 0;
				}
				else
					m.exact = 0;
				c.skip(12 * np);
			}
		}
		else if (m.collision_type == 1)
		{
			m.n_triangles = c.get<uint64_t>();
			if (c.bad || m.n_triangles > (c.n - c.o) / 36u) return GPX_ERR_INVALID_ARG;
		}
		if (c.bad) return GPX_ERR_INVALID_ARG;
		*out = m;
		return GPX_OK;
	}
	catch (...) { return GPX_ERR_INVALID_ARG; }
}

int gpx_model_load_gmdl_container(const uint8_t *blob, uint64_t size, float tolerance, gpx_model_collision *out)
{
	std::vector<uint8_t> body;
	const int rc = inflate_container(blob, size, 0xFF, body);
	if (rc != GPX_OK) return rc;
	return gpx_model_load_gmdl(body.data(), body.size(), tolerance, out);
}

int gpx_static_add_gmdl(gpx_world *w, const gpx_transform *xfm, const uint8_t *body, uint64_t size, float friction, uint32_t ray_flags)
{
	if (!w || !xfm || !body) return -GPX_ERR_INVALID_ARG;
	try
	{
		gpx_model_collision m;
		memset(&m, 0, sizeof(m));
		Cursor c{body, size};
		if (!gmdl_seek_collision(c, m) || m.collision_type != 1) return -GPX_ERR_INVALID_ARG;
		const uint64_t nt = c.get<uint64_t>();
		if (c.bad || nt == 0 || nt > (c.n - c.o) / 36u) return -GPX_ERR_INVALID_ARG;
		std::vector<float> tris(9 * (size_t)nt);
		memcpy(tris.data(), c.d + c.o, 36 * (size_t)nt);
		uint32_t sbody = 0;
		const int rc = gpx_static_add_mesh(w, xfm, tris.data(), nt, friction, 0, &sbody);
		if (rc != GPX_OK) return -rc;
		if (ray_flags != GPX_BODY_BLOCKS_LASERS) gpx_body_set_ray_flags(w, 0, sbody, ray_flags);
		return (int)(sbody - STATIC_BODY_BASE);
	}
	catch (...) { return -GPX_ERR_INVALID_ARG; }
}

int gpx_static_load_gmap_file(gpx_world *w, const char *path)
{
	if (!w || !path) return -GPX_ERR_INVALID_ARG;
	FILE *f = fopen(path, "rb");
	if (!f) return -GPX_ERR_INVALID_ARG;
	fseek(f, 0, SEEK_END);
	const long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	std::vector<uint8_t> blob(n > 0 ? (size_t)n : 0);
	const size_t got = blob.empty() ? 0 : fread(blob.data(), 1, blob.size(), f);
	fclose(f);
	if (got != blob.size()) return -GPX_ERR_INVALID_ARG;
	return gpx_static_load_gmap_container(w, blob.data(), blob.size());
}

/* JPH_ConvexHullShape_Create as the model loader uses it (engine/src/assets/ModelLoader.c:324-341: one hull per
 * ModelConvexHull of a .gmdl, ModelLoader.c:154-183).  The body store knows boxes and spheres, so a hull is
 * classified: all points at the same distance from the centroid (to 2 %) -> SPHERE (orb.gmdl, 32 514 points);
 * every point on or just inside the surface of its bounding box, with points near all eight corners -> BOX
 * (cube.gmdl: a bevelled 0.4 m cube); anything else -> its bounding box, reported as approximate. */
int gpx_shape_from_hull(const float *points, uint64_t n, float tolerance, gpx_hull_shape *out)
{
	if (!points || n < 4 || !out || !(tolerance >= 0.0f)) return GPX_ERR_INVALID_ARG;
	double lo[3] = {1e30, 1e30, 1e30}, hi[3] = {-1e30, -1e30, -1e30}, c[3] = {0, 0, 0};
	for (uint64_t i = 0; i < n; i++)
		for (int k = 0; k < 3; k++)
		{
			const double v = points[3 * i + k];
			lo[k] = v < lo[k] ? v : lo[k];
			hi[k] = v > hi[k] ? v : hi[k];
			c[k] += v;
		}
	for (int k = 0; k < 3; k++) c[k] /= (double)n;
	double rmin = 1e30, rmax = 0.0;
	for (uint64_t i = 0; i < n; i++)
	{
		double r2 = 0.0;
		for (int k = 0; k < 3; k++)
		{
			const double d = points[3 * i + k] - c[k];
			r2 += d * d;
		}
		const double r = sqrt(r2);
		rmin = r < rmin ? r : rmin;
		rmax = r > rmax ? r : rmax;
	}
	memset(out, 0, sizeof(*out));
	double he[3], bc[3];
	for (int k = 0; k < 3; k++)
	{
		he[k] = 0.5 * (hi[k] - lo[k]);
		bc[k] = 0.5 * (hi[k] + lo[k]);
	}
	/* radii agree to 2 % and the points reach that radius along every axis (a cube's corner points also share one radius) */
	const bool round_box = fabs(he[0] - rmax) <= 0.03 * rmax && fabs(he[1] - rmax) <= 0.03 * rmax && fabs(he[2] - rmax) <= 0.03 * rmax;
	if (rmax - rmin <= 0.02 * rmax && rmax - rmin <= 2.0 * tolerance && round_box)
	{
		out->shape = GPX_SHAPE_SPHERE;
		out->half_extents[0] = (float)(0.5 * (rmax + rmin));
		for (int k = 0; k < 3; k++) out->center[k] = (float)c[k];
		out->exact = 1;
		return GPX_OK;
	}
	/* box: all points near the surface (a bevel cuts at most `tolerance` off edges and corners), every corner region hit */
	bool on_surface = true;
	uint32_t corners = 0;
	for (uint64_t i = 0; i < n; i++)
	{
		double gap = 1e30;
		uint32_t oct = 0;
		bool near_corner = true;
		for (int k = 0; k < 3; k++)
		{
			const double d = points[3 * i + k] - bc[k];
			const double g = he[k] - fabs(d);
			gap = g < gap ? g : gap;
			if (d > 0) oct |= 1u << k;
			if (g > 4.0 * tolerance) near_corner = false;
		}
		if (gap > tolerance) on_surface = false;
		if (near_corner) corners |= 1u << oct;
	}
	out->shape = GPX_SHAPE_BOX;
	for (int k = 0; k < 3; k++)
	{
		out->half_extents[k] = (float)he[k];
		out->center[k] = (float)bc[k];
	}
	out->exact = (on_surface && corners == 0xFFu) ? 1u : 0u;
	return GPX_OK;
}

/* ---- bodies */

static inline bool valid_slot(const gpx_world *w, uint32_t world, uint32_t body)
{
	return w && world < w->W && body < w->cap;
}

uint32_t gpx_body_create(gpx_world *w, uint32_t world, const gpx_body_desc *desc)
{
	if (!w || !desc || world >= w->W) return GPX_INVALID_BODY;
	std::lock_guard<std::mutex> lk(w->mu);
	const size_t base = (size_t)world * w->cap;
	// lowest free slot; everything below the hint is known to be taken (destroy lowers it)
	for (uint32_t i = w->free_hint[world]; i < w->cap; i++)
		if (!(w->h_flags[base + i] & BF_ALIVE))
		{
			w->free_hint[world] = i + 1;
			BodyCommand c = command_from_desc((uint32_t)(base + i), *desc);
			if (desc->allow_sleeping && desc->motion_type == GPX_MOTION_DYNAMIC) w->sleep_enabled = true;
			w->h_flags[base + i] = c.flags;
			w->h_user_data[base + i] = desc->user_data;
			front_pos(w)[base + i] = c.pos;
			front_quat(w)[base + i] = c.quat;
			w->m_lin[base + i] = c.lin;
			w->m_ang[base + i] = c.ang;
			w->pending.push_back(c);
			return i;
		}
	return GPX_INVALID_BODY;
}

int gpx_body_create_all(gpx_world *w, const gpx_body_desc *descs, uint32_t count, const float *linvel, const float *angvel,
						uint32_t *out_ids)
{
	if (!w || !descs) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	// same free slots in every world: take them from world 0 and require the same layout everywhere
	std::vector<uint32_t> ids;
	for (uint32_t i = 0; i < w->cap && ids.size() < count; i++)
		if (!(w->h_flags[i] & BF_ALIVE)) ids.push_back(i);
	if (ids.size() < count) return GPX_ERR_CAPACITY;
	w->pending.reserve(w->pending.size() + (size_t)w->W * count);
	for (uint32_t k = 0; k < count; k++)
	{
		BodyCommand proto = command_from_desc(0, descs[k]);
		if (descs[k].allow_sleeping && descs[k].motion_type == GPX_MOTION_DYNAMIC) w->sleep_enabled = true;
		for (uint32_t wi = 0; wi < w->W; wi++)
		{
			const size_t g = (size_t)wi * w->cap + ids[k];
			if (w->h_flags[g] & BF_ALIVE) return GPX_ERR_CAPACITY;
			BodyCommand c = proto;
			c.index = (uint32_t)g;
			if (linvel)
			{
				const float *v = linvel + ((size_t)wi * count + k) * 3;
				c.lin = make_float4(v[0], v[1], v[2], 0.0f);
			}
			if (angvel)
			{
				const float *v = angvel + ((size_t)wi * count + k) * 3;
				c.ang = make_float4(v[0], v[1], v[2], 0.0f);
			}
			w->h_flags[g] = c.flags;
			w->h_user_data[g] = descs[k].user_data;
			front_pos(w)[g] = c.pos;
			front_quat(w)[g] = c.quat;
			w->m_lin[g] = c.lin;
			w->m_ang[g] = c.ang;
			w->pending.push_back(c);
		}
		if (out_ids) out_ids[k] = ids[k];
	}
	return GPX_OK;
}

int gpx_body_destroy(gpx_world *w, uint32_t world, uint32_t body)
{
	if (!valid_slot(w, world, body)) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	const size_t g = (size_t)world * w->cap + body;
	if (!(w->h_flags[g] & BF_ALIVE)) return GPX_ERR_INVALID_ARG;
	w->h_flags[g] = 0;
	w->h_user_data[g] = 0;
	if (body < w->free_hint[world]) w->free_hint[world] = body;
	BodyCommand c;
	memset(&c, 0, sizeof(c));
	c.index = (uint32_t)g;
	c.mask = 31u;
	c.quat = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
	w->pending.push_back(c);
	return GPX_OK;
}

static int queue_write(gpx_world *w, uint32_t world, uint32_t body, uint32_t mask, const float *a, const float *b)
{
	if (!valid_slot(w, world, body)) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	const size_t g = (size_t)world * w->cap + body;
	if (!(w->h_flags[g] & BF_ALIVE)) return GPX_ERR_INVALID_ARG;
	BodyCommand c;
	memset(&c, 0, sizeof(c));
	c.index = (uint32_t)g;
	c.mask = mask;
	if (mask & 1u) front_pos(w)[g] = c.pos = make_float4(a[0], a[1], a[2], 0.0f);
	if (mask & 2u)
	{
		q4 q;
		q.x = a[0]; q.y = a[1]; q.z = a[2]; q.w = a[3];
		q = qnormalize(q);
		front_quat(w)[g] = c.quat = make_float4(q.x, q.y, q.z, q.w);
	}
	if (mask & 4u) w->m_lin[g] = c.lin = make_float4(a[0], a[1], a[2], 0.0f);
	if (mask & 8u)
	{
		const float *s = (mask & 4u) ? b : a;
		w->m_ang[g] = c.ang = make_float4(s[0], s[1], s[2], 0.0f);
	}
	w->pending.push_back(c);
	return GPX_OK;
}

// 64 = wake: a non-zero velocity activates a sleeping body (BodyInterface::SetLinearVelocity), and so does
// JPH_Activation_Activate on SetPosition / SetRotation
static inline uint32_t wake_if(bool on) { return on ? 64u : 0u; }
static inline bool nonzero3(const float *v) { return v[0] != 0.0f || v[1] != 0.0f || v[2] != 0.0f; }

int gpx_body_set_linear_velocity(gpx_world *w, uint32_t world, uint32_t body, const float v[3])
{
	return v ? queue_write(w, world, body, 4u | wake_if(nonzero3(v)), v, nullptr) : GPX_ERR_INVALID_ARG;
}
int gpx_body_set_linear_and_angular_velocity(gpx_world *w, uint32_t world, uint32_t body, const float v[3], const float av[3])
{
	return (v && av) ? queue_write(w, world, body, 12u | wake_if(nonzero3(v) || nonzero3(av)), v, av) : GPX_ERR_INVALID_ARG;
}
int gpx_body_set_position(gpx_world *w, uint32_t world, uint32_t body, const float p[3], int activate)
{
	return p ? queue_write(w, world, body, 1u | wake_if(activate != 0), p, nullptr) : GPX_ERR_INVALID_ARG;
}
int gpx_body_set_rotation(gpx_world *w, uint32_t world, uint32_t body, const float q[4], int activate)
{
	return q ? queue_write(w, world, body, 2u | wake_if(activate != 0), q, nullptr) : GPX_ERR_INVALID_ARG;
}
int gpx_body_wake(gpx_world *w, uint32_t world, uint32_t body)
{
	return queue_write(w, world, body, 64u, nullptr, nullptr);
}
int gpx_read_sleeping(gpx_world *w, uint8_t *out, uint64_t capacity)
{
	if (!w || !out) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	const size_t nb = (size_t)w->W * w->cap;
	if (capacity < nb) return GPX_ERR_CAPACITY;
	int rc;
	if ((rc = flush_commands(w)) != GPX_OK) return rc;
	std::vector<uint32_t> f(nb);
	GPX_CUDA(cudaMemcpyAsync(f.data(), w->bs.flags, sizeof(uint32_t) * nb, cudaMemcpyDeviceToHost, w->stream));
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	for (size_t g = 0; g < nb; g++) out[g] = (f[g] & BF_ALIVE) && (f[g] & BF_ASLEEP) ? 1 : 0;
	return GPX_OK;
}

int gpx_body_set_ray_flags(gpx_world *w, uint32_t world, uint32_t body, uint32_t ray_flags)
{
	if (w && body >= STATIC_BODY_BASE && body - STATIC_BODY_BASE < w->sbodies.size())
	{
		std::lock_guard<std::mutex> lk(w->mu);
		StaticBodyHost &sb = w->sbodies[body - STATIC_BODY_BASE];
		if (sb.ray_flags != (ray_flags & 0xFFu))
		{
			sb.ray_flags = ray_flags & 0xFFu;
			w->static_dirty = true;  // the per-body table travels with the static upload
		}
		return GPX_OK;
	}
	if (!valid_slot(w, world, body)) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	const size_t g = (size_t)world * w->cap + body;
	if (!(w->h_flags[g] & BF_ALIVE)) return GPX_ERR_INVALID_ARG;
	const uint32_t f = (w->h_flags[g] & ~(0xFFu << BF_RAYFLAG_SHIFT)) | ((ray_flags & 0xFFu) << BF_RAYFLAG_SHIFT);
	if (f == w->h_flags[g]) return GPX_OK;
	w->h_flags[g] = f;
	BodyCommand c;
	memset(&c, 0, sizeof(c));
	c.index = (uint32_t)g;
	c.mask = 32u;
	c.flags = f;
	w->pending.push_back(c);
	return GPX_OK;
}

int gpx_body_get_transform(const gpx_world *w, uint32_t world, uint32_t body, gpx_transform *out)
{
	if (!out) return GPX_ERR_INVALID_ARG;
	if (w && world < w->W && body >= STATIC_BODY_BASE && body - STATIC_BODY_BASE < w->sbodies.size())
	{
		*out = w->sbodies[body - STATIC_BODY_BASE].xfm;
		return GPX_OK;
	}
	if (!valid_slot(w, world, body)) return GPX_ERR_INVALID_ARG;
	const size_t g = (size_t)world * w->cap + body;
	float4 p, q;
	for (;;)
	{
		// wait-free unless a whole readback completes in between (then read the new front)
		const uint32_t gen = w->mirror_gen.load(std::memory_order_acquire);
		p = w->mb_pos[gen & 1u][g];
		q = w->mb_quat[gen & 1u][g];
		if (w->mirror_gen.load(std::memory_order_acquire) == gen) break;
	}
	out->position[0] = p.x; out->position[1] = p.y; out->position[2] = p.z;
	out->rotation[0] = q.x; out->rotation[1] = q.y; out->rotation[2] = q.z; out->rotation[3] = q.w;
	return GPX_OK;
}

int gpx_body_get_world_matrix(const gpx_world *w, uint32_t world, uint32_t body, float m[16])
{
	gpx_transform t;
	int rc = gpx_body_get_transform(w, world, body, &t);
	if (rc != GPX_OK) return rc;
	q4 q;
	q.x = t.rotation[0]; q.y = t.rotation[1]; q.z = t.rotation[2]; q.w = t.rotation[3];
	m33 R = qmat(q);
	/* column-major 4x4, as cglm/Jolt Mat44 (engine/src/graphics/RenderingHelpers.c:101-126) */
	m[0] = R.c0.x; m[1] = R.c0.y; m[2] = R.c0.z; m[3] = 0.0f;
	m[4] = R.c1.x; m[5] = R.c1.y; m[6] = R.c1.z; m[7] = 0.0f;
	m[8] = R.c2.x; m[9] = R.c2.y; m[10] = R.c2.z; m[11] = 0.0f;
	m[12] = t.position[0]; m[13] = t.position[1]; m[14] = t.position[2]; m[15] = 1.0f;
	return GPX_OK;
}

int gpx_body_get_velocity(const gpx_world *w, uint32_t world, uint32_t body, float v[3], float av[3])
{
	if (!valid_slot(w, world, body)) return GPX_ERR_INVALID_ARG;
	const size_t g = (size_t)world * w->cap + body;
	if (v) { v[0] = w->m_lin[g].x; v[1] = w->m_lin[g].y; v[2] = w->m_lin[g].z; }
	if (av) { av[0] = w->m_ang[g].x; av[1] = w->m_ang[g].y; av[2] = w->m_ang[g].z; }
	return GPX_OK;
}

uint64_t gpx_body_get_user_data(const gpx_world *w, uint32_t world, uint32_t body)
{
	if (w && body >= STATIC_BODY_BASE && body - STATIC_BODY_BASE < w->sbodies.size())
		return w->sbodies[body - STATIC_BODY_BASE].user_data;
	if (!valid_slot(w, world, body)) return 0;
	return w->h_user_data[(size_t)world * w->cap + body];
}

int gpx_body_is_active(const gpx_world *w, uint32_t world, uint32_t body)
{
	if (!valid_slot(w, world, body)) return 0;
	return (w->h_flags[(size_t)world * w->cap + body] & BF_ALIVE) ? 1 : 0;
}

/* ---- tick */

int gpx_step(gpx_world *w, float dt, int collision_steps)
{
	if (!w || !(dt > 0.0f)) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	int rc;
	if (w->static_dirty && (rc = build_static(w)) != GPX_OK) return rc;
	if ((rc = flush_commands(w)) != GPX_OK) return rc;
	if ((rc = (w->wide ? launch_wide_tick(w, dt, collision_steps) : launch_tick(w, dt, collision_steps))) != GPX_OK) return rc;
	w->mirror_direct = !w->wide && w->mb_dev[0] && w->mb_dev[1] && w->synced_since_step;  // what launch_tick decided
	w->synced_since_step = false;
	// the sleep test runs once per tick, and only in worlds that hold a body allowed to sleep (wide worlds run theirs
	// at the end of launch_wide_tick)
	if (w->sleep_enabled && !w->wide && (rc = launch_sleep_test(w, dt)) != GPX_OK) return rc;
	w->ticks++;
	return (int)w->m_err[0];
}

// The world's stream waits for a hits copy still travelling on the copy stream (callers hold w->mu)
static int join_hits(gpx_world *w)
{
	if (w->hits_pending)
	{
		GPX_CUDA(cudaStreamWaitEvent(w->stream, w->ev_hits_done, 0));
		w->hits_pending = false;
	}
	return GPX_OK;
}

int gpx_sync_transforms(gpx_world *w)
{
	if (!w) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	int rc;
	if ((rc = flush_commands(w)) != GPX_OK) return rc;
	if ((rc = join_hits(w)) != GPX_OK) return rc;
	const size_t nb = (size_t)w->W * w->cap;
	const uint32_t gen = w->mirror_gen.load(std::memory_order_relaxed), back = (gen + 1u) & 1u;
	if (w->mirror_direct)
		// the tick wrote positions and orientations into the back buffer itself: only the error word is left
		GPX_CUDA(cudaMemcpyAsync(w->mb_pos[back] + 2 * nb, w->bs.pos + 2 * nb, sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
	else
		// positions | orientations | error word: one allocation on each side, one copy
		GPX_CUDA(cudaMemcpyAsync(w->mb_pos[back], w->bs.pos, 2 * sizeof(float4) * nb + sizeof(uint32_t), cudaMemcpyDeviceToHost,
								 w->stream));
	// the error word has been read: the ticks after this synchronisation report their own errors (Jolt's Update returns
	// each update's result); the per-world words d_err[1 + world] stay sticky for gpx_read_stats
	GPX_CUDA(cudaMemsetAsync(w->d_err, 0, sizeof(uint32_t), w->stream));
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	w->m_err[0] = *reinterpret_cast<const uint32_t *>(w->mb_pos[back] + 2 * nb);
	w->mirror_gen.store(gen + 1u, std::memory_order_release);  // the tick just read back becomes the front
	w->mirror_direct = false;                                  // the next tick writes the other buffer
	w->synced_since_step = true;
	return (int)w->m_err[0];
}

int gpx_read_transforms(gpx_world *w, gpx_transform *out, uint64_t capacity)
{
	if (!w || !out) return GPX_ERR_INVALID_ARG;
	const size_t nb = (size_t)w->W * w->cap;
	if (capacity < nb) return GPX_ERR_CAPACITY;
	int rc = gpx_sync_transforms(w);
	for (size_t g = 0; g < nb; g++)
	{
		const float4 p = front_pos(w)[g], q = front_quat(w)[g];
		out[g].position[0] = p.x; out[g].position[1] = p.y; out[g].position[2] = p.z;
		out[g].rotation[0] = q.x; out[g].rotation[1] = q.y; out[g].rotation[2] = q.z; out[g].rotation[3] = q.w;
	}
	return rc;
}

int gpx_read_velocities(gpx_world *w, float *out, uint64_t capacity)
{
	if (!w || !out) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	const size_t nb = (size_t)w->W * w->cap;
	if (capacity < nb) return GPX_ERR_CAPACITY;
	int rc;
	if ((rc = flush_commands(w)) != GPX_OK) return rc;
	GPX_CUDA(cudaMemcpyAsync(w->m_lin, w->bs.lin, sizeof(float4) * nb, cudaMemcpyDeviceToHost, w->stream));
	GPX_CUDA(cudaMemcpyAsync(w->m_ang, w->bs.ang, sizeof(float4) * nb, cudaMemcpyDeviceToHost, w->stream));
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	for (size_t g = 0; g < nb; g++)
	{
		out[6 * g + 0] = w->m_lin[g].x; out[6 * g + 1] = w->m_lin[g].y; out[6 * g + 2] = w->m_lin[g].z;
		out[6 * g + 3] = w->m_ang[g].x; out[6 * g + 4] = w->m_ang[g].y; out[6 * g + 5] = w->m_ang[g].z;
	}
	return GPX_OK;
}

int gpx_read_stats(gpx_world *w, gpx_world_stats *out)
{
	if (!w || !out) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	int rc;
	if ((rc = flush_commands(w)) != GPX_OK) return rc;
	if ((rc = launch_stats(w)) != GPX_OK) return rc;
	if (w->wide && (rc = wide_stats(w)) != GPX_OK) return rc;
	GPX_CUDA(cudaMemcpyAsync(out, w->d_stats, sizeof(gpx_world_stats) * w->W, cudaMemcpyDeviceToHost, w->stream));
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	return GPX_OK;
}

/* ---- rays */

// callers hold w->mu
static int raycast_device_locked(gpx_world *w, const void *d_rays, uint64_t n, void *d_hits)
{
	cudaSetDevice(w->device);
	int rc;
	if (w->static_dirty && (rc = build_static(w)) != GPX_OK) return rc;
	if ((rc = flush_commands(w)) != GPX_OK) return rc;
	return launch_raycast(w, d_rays, n, d_hits);
}

int gpx_raycast_batch_device(gpx_world *w, const void *d_rays, uint64_t n, void *d_hits)
{
	if (!w || (n && (!d_rays || !d_hits))) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	return raycast_device_locked(w, d_rays, n, d_hits);
}

// callers hold w->mu; the staging buffers hold n rays
static int raycast_pipelined(gpx_world *w, const gpx_ray *rays, uint64_t n, gpx_hit *hits, uint64_t chunks)
{
	int rc;
	if (w->static_dirty && (rc = build_static(w)) != GPX_OK) return rc;
	if ((rc = flush_commands(w)) != GPX_OK) return rc;
	if (!w->stream_h2d) GPX_CUDA(cudaStreamCreateWithFlags(&w->stream_h2d, cudaStreamNonBlocking));
	for (uint64_t k = 0; k < 2 * chunks + 1; k++)
		if (!w->ev_pipe[k]) GPX_CUDA(cudaEventCreateWithFlags(&w->ev_pipe[k], cudaEventDisableTiming));
	// the copy stream starts after whatever the world's stream still does with the staging buffers
	GPX_CUDA(cudaEventRecord(w->ev_pipe[2 * chunks], w->stream));
	GPX_CUDA(cudaStreamWaitEvent(w->stream_h2d, w->ev_pipe[2 * chunks], 0));
	const uint64_t per = ((n + chunks - 1) / chunks + 1023u) & ~1023ull;
	gpx_ray *d_rays = static_cast<gpx_ray *>(w->d_rays);
	gpx_hit *d_hits = static_cast<gpx_hit *>(w->d_hits);
	uint64_t c = 0;
	for (uint64_t off = 0; off < n; off += per, c++)
	{
		const uint64_t cnt = n - off < per ? n - off : per;
		GPX_CUDA(cudaMemcpyAsync(d_rays + off, rays + off, sizeof(gpx_ray) * cnt, cudaMemcpyHostToDevice, w->stream_h2d));
		GPX_CUDA(cudaEventRecord(w->ev_pipe[2 * c], w->stream_h2d));
		GPX_CUDA(cudaStreamWaitEvent(w->stream, w->ev_pipe[2 * c], 0));
		if ((rc = launch_raycast(w, d_rays + off, cnt, d_hits + off)) != GPX_OK) return rc;
		GPX_CUDA(cudaEventRecord(w->ev_pipe[2 * c + 1], w->stream));
		GPX_CUDA(cudaStreamWaitEvent(w->stream_copy, w->ev_pipe[2 * c + 1], 0));
		GPX_CUDA(cudaMemcpyAsync(hits + off, d_hits + off, sizeof(gpx_hit) * cnt, cudaMemcpyDeviceToHost, w->stream_copy));
	}
	GPX_CUDA(cudaEventRecord(w->ev_hits_done, w->stream_copy));
	GPX_CUDA(cudaStreamWaitEvent(w->stream, w->ev_hits_done, 0));
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	return GPX_OK;
}

// One batch through the world's staging buffers.  The world's mutex is held from staging the rays to the last use of
// the shared buffers (two threads casting on one world would otherwise overwrite each other's rays or hits, or free
// buffers the other is about to launch with); the synchronous variant therefore also waits inside.
// the device's view of a host buffer that is pinned and mapped (cudaMallocHost / cudaHostRegister), or nullptr
static void *mapped_alias(const void *host)
{
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, host) != cudaSuccess)
	{
		cudaGetLastError();
		return nullptr;
	}
	return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

static int raycast_enqueue_locked(gpx_world *w, const gpx_ray *rays, uint64_t n, gpx_hit *hits, bool async);
static int raycast_enqueue(gpx_world *w, const gpx_ray *rays, uint64_t n, gpx_hit *hits, bool async)
{
	std::lock_guard<std::mutex> lk(w->mu);
	return raycast_enqueue_locked(w, rays, n, hits, async);
}

static int raycast_enqueue_locked(gpx_world *w, const gpx_ray *rays, uint64_t n, gpx_hit *hits, bool async)
{
	{
		cudaSetDevice(w->device);
		// Pinned, mapped host buffers: the kernel reads the rays and writes the hits across the bus itself — no staging
		// copies, one launch; the bus carries both directions at once while the kernel traces.
		static const bool no_zero_copy = getenv("GPX_NO_ZERO_COPY_RAYS") != nullptr;
		void *zr = no_zero_copy ? nullptr : mapped_alias(rays), *zh = zr ? mapped_alias(hits) : nullptr;
		if (zr && zh)
		{
			int rc = raycast_device_locked(w, zr, n, zh);
			if (rc != GPX_OK || async) return rc;  // async: valid once the world's stream has passed this point
			GPX_CUDA(cudaStreamSynchronize(w->stream));
			return GPX_OK;
		}
		int jr = join_hits(w);  // the previous batch's hits leave d_hits before this one's kernel writes it
		if (jr != GPX_OK) return jr;
		if (n > w->ray_cap)
		{
			GPX_CUDA(cudaStreamSynchronize(w->stream));  // a batch still in flight may be using the old buffers
			cudaFree(w->d_rays);
			cudaFree(w->d_hits);
			w->d_rays = w->d_hits = nullptr;
			w->ray_cap = 0;
			GPX_CUDA(cudaMalloc(&w->d_rays, sizeof(gpx_ray) * n));
			GPX_CUDA(cudaMalloc(&w->d_hits, sizeof(gpx_hit) * n));
			w->ray_cap = n;
		}
	}
	// A large synchronous batch is PCIe-bound (48 bytes per ray cross the bus): cut it into chunks so that chunk c's
	// kernel and chunk c-1's hits copy run under chunk c+1's rays copy.
	constexpr uint64_t PIPE_MIN = 1ull << 18, PIPE_CHUNKS = 8;
	if (!async && n >= PIPE_MIN) return raycast_pipelined(w, rays, n, hits, PIPE_CHUNKS);
	GPX_CUDA(cudaMemcpyAsync(w->d_rays, rays, sizeof(gpx_ray) * n, cudaMemcpyHostToDevice, w->stream));
	int rc = raycast_device_locked(w, w->d_rays, n, w->d_hits);
	if (rc != GPX_OK) return rc;
	if (!async)
	{
		GPX_CUDA(cudaMemcpyAsync(hits, w->d_hits, sizeof(gpx_hit) * n, cudaMemcpyDeviceToHost, w->stream));
		GPX_CUDA(cudaStreamSynchronize(w->stream));
		return GPX_OK;
	}
	// async: the copy back overlaps whatever the caller enqueues next (the tick); the world's stream picks it up again
	// at the next gpx_sync_transforms / gpx_device_sync / ray batch
	GPX_CUDA(cudaEventRecord(w->ev_rays_done, w->stream));
	GPX_CUDA(cudaStreamWaitEvent(w->stream_copy, w->ev_rays_done, 0));
	GPX_CUDA(cudaMemcpyAsync(hits, w->d_hits, sizeof(gpx_hit) * n, cudaMemcpyDeviceToHost, w->stream_copy));
	GPX_CUDA(cudaEventRecord(w->ev_hits_done, w->stream_copy));
	w->hits_pending = true;
	return GPX_OK;
}

int gpx_raycast_batch(gpx_world *w, const gpx_ray *rays, uint64_t n, gpx_hit *hits)
{
	if (!w || (n && (!rays || !hits))) return GPX_ERR_INVALID_ARG;
	if (n == 0) return GPX_OK;
	return raycast_enqueue(w, rays, n, hits, false);
}

int gpx_raycast_batch_async(gpx_world *w, const gpx_ray *rays, uint64_t n, gpx_hit *hits)
{
	if (!w || (n && (!rays || !hits))) return GPX_ERR_INVALID_ARG;
	if (n == 0) return GPX_OK;
	return raycast_enqueue(w, rays, n, hits, true);
}

int gpx_raycast_transform(gpx_world *w, uint32_t world, const gpx_transform *origin, float max_distance, uint32_t mask,
						  gpx_hit *out)
{
	if (!w || !origin || !out || world >= w->W) return GPX_ERR_INVALID_ARG;
	gpx_ray r;
	q4 q;
	q.x = origin->rotation[0]; q.y = origin->rotation[1]; q.z = origin->rotation[2]; q.w = origin->rotation[3];
	const v3 d = qrot(q, V(0.0f, 0.0f, -1.0f));  // forward = local -Z (PlayerPhysics.c:360, LaserEmitter.c:105-110)
	r.origin[0] = origin->position[0]; r.origin[1] = origin->position[1]; r.origin[2] = origin->position[2];
	r.tmax = max_distance;
	r.dir[0] = d.x; r.dir[1] = d.y; r.dir[2] = d.z;
	r.mask = (mask & 0xFFFFu) | (world << 16);
	// the engine's one ray per call (crosshair, a laser): through the world's pinned scratch record, so that it takes the
	// copy-free path — one launch and one synchronisation
	std::lock_guard<std::mutex> lk(w->mu);
	if (!w->m_err) return GPX_ERR_INVALID_ARG;
	gpx_ray *pr = reinterpret_cast<gpx_ray *>(w->m_err + 4);
	gpx_hit *ph = reinterpret_cast<gpx_hit *>(w->m_err + 4 + sizeof(gpx_ray) / sizeof(uint32_t));
	*pr = r;
	const int rc = raycast_enqueue_locked(w, pr, 1, ph, false);
	if (rc == GPX_OK) *out = *ph;
	return rc;
}

int gpx_spherecast_batch(gpx_world *w, const gpx_sphere_cast *casts, uint64_t n, gpx_cast_hit *hits)
{
	if (!w || (n && (!casts || !hits))) return GPX_ERR_INVALID_ARG;
	if (n == 0) return GPX_OK;
	static_assert(sizeof(gpx_sphere_cast) == 48 && sizeof(gpx_cast_hit) == 32, "three / two float4 per record");
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	int rc;
	if (w->static_dirty && (rc = build_static(w)) != GPX_OK) return rc;
	if ((rc = flush_commands(w)) != GPX_OK) return rc;
	// staging shared with the capsule queries: 48 B in, 32 B out per cast
	if (2 * n > w->capq_cap)
	{
		GPX_CUDA(cudaStreamSynchronize(w->stream));
		cudaFree(w->d_capq);
		cudaFree(w->d_capo);
		w->d_capq = w->d_capo = nullptr;
		w->capq_cap = 0;
		GPX_CUDA(cudaMalloc(&w->d_capq, 64ull * n));
		GPX_CUDA(cudaMalloc(&w->d_capo, 64ull * n));
		w->capq_cap = 2 * n;
	}
	GPX_CUDA(cudaMemcpyAsync(w->d_capq, casts, 48ull * n, cudaMemcpyHostToDevice, w->stream));
	if ((rc = launch_spherecast(w, w->d_capq, n, w->d_capo)) != GPX_OK) return rc;
	GPX_CUDA(cudaMemcpyAsync(hits, w->d_capo, 32ull * n, cudaMemcpyDeviceToHost, w->stream));
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	return GPX_OK;
}

int gpx_overlap_capsule_batch(gpx_world *w, const gpx_capsule_query *queries, uint64_t n, gpx_overlap *out)
{
	if (!w || (n && (!queries || !out))) return GPX_ERR_INVALID_ARG;
	if (n == 0) return GPX_OK;
	static_assert(sizeof(gpx_capsule_query) == 32 && sizeof(gpx_overlap) == 32, "two float4 per record");
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	int rc;
	if (w->static_dirty && (rc = build_static(w)) != GPX_OK) return rc;
	if ((rc = flush_commands(w)) != GPX_OK) return rc;
	if (n > w->capq_cap)
	{
		GPX_CUDA(cudaStreamSynchronize(w->stream));
		cudaFree(w->d_capq);
		cudaFree(w->d_capo);
		w->d_capq = w->d_capo = nullptr;
		w->capq_cap = 0;
		GPX_CUDA(cudaMalloc(&w->d_capq, 32ull * n));
		GPX_CUDA(cudaMalloc(&w->d_capo, 32ull * n));
		w->capq_cap = n;
	}
	GPX_CUDA(cudaMemcpyAsync(w->d_capq, queries, 32ull * n, cudaMemcpyHostToDevice, w->stream));
	if ((rc = launch_overlap_capsules(w, w->d_capq, n, w->d_capo)) != GPX_OK) return rc;
	GPX_CUDA(cudaMemcpyAsync(out, w->d_capo, 32ull * n, cudaMemcpyDeviceToHost, w->stream));
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	return GPX_OK;
}

/* ---- harness helpers */

void *gpx_device_alloc(uint64_t bytes)
{
	void *p = nullptr;
	if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
	return p;
}
void gpx_device_free(void *p) { cudaFree(p); }
void *gpx_host_alloc(uint64_t bytes)
{
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
	return p;
}
void gpx_host_free(void *p) { cudaFreeHost(p); }
int gpx_memcpy_h2d(void *dst, const void *src, uint64_t bytes)
{
	GPX_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
	return GPX_OK;
}
int gpx_memcpy_d2h(void *dst, const void *src, uint64_t bytes)
{
	GPX_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
	return GPX_OK;
}
int gpx_device_sync(gpx_world *w)
{
	if (!w) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	int rc = join_hits(w);
	if (rc != GPX_OK) return rc;
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	return GPX_OK;
}
/* ---- player character */

static int char_upload(gpx_world *w, uint32_t world)
{
	GPX_CUDA(cudaMemcpyAsync(w->d_ch + world, &w->h_ch[world], sizeof(CharDev), cudaMemcpyHostToDevice, w->stream));
	GPX_CUDA(cudaStreamSynchronize(w->stream));  // h_ch is pageable
	return GPX_OK;
}

int gpx_character_create(gpx_world *w, uint32_t world, const gpx_character_desc *d)
{
	if (!w || !d || world >= w->W) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	if (!w->d_ch)
	{
		int rc;
		if ((rc = dalloc(&w->d_ch, (size_t)w->W)) != GPX_OK ||
			(rc = dalloc(&w->d_ch_keys, (size_t)w->W * CHARACTER_MAX_CONTACTS)) != GPX_OK ||
			(rc = dalloc(&w->d_ch_nkeys, (size_t)w->W)) != GPX_OK)
			return rc;
		w->h_ch.assign(w->W, CharDev{});
	}
	CharDev &c = w->h_ch[world];
	memset(&c, 0, sizeof(c));
	c.px = d->position[0]; c.py = d->position[1]; c.pz = d->position[2];
	c.hh = d->half_height;
	c.r = d->radius;
	c.cos_slope = cosf(d->max_slope_deg * 0.0174532925f);
	c.gny = 1.0f;
	c.alive = 1;
	c.ground = 3;  /* JPH_GroundState_InAir until the first update */
	c.ground_body = GPX_INVALID_BODY;
	return char_upload(w, world);
}

int gpx_character_destroy(gpx_world *w, uint32_t world)
{
	if (!w || world >= w->W || !w->d_ch) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	memset(&w->h_ch[world], 0, sizeof(CharDev));
	GPX_CUDA(cudaMemsetAsync(w->d_ch_nkeys + world, 0, sizeof(uint32_t), w->stream));
	return char_upload(w, world);
}

static int char_fetch(gpx_world *w, uint32_t world)
{
	GPX_CUDA(cudaMemcpyAsync(&w->h_ch[world], w->d_ch + world, sizeof(CharDev), cudaMemcpyDeviceToHost, w->stream));
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	return GPX_OK;
}

int gpx_character_set_linear_velocity(gpx_world *w, uint32_t world, const float v[3])
{
	if (!w || world >= w->W || !w->d_ch || !v) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	// three floats into the record, in stream order; no read-back (the source is pageable: staged before the call returns)
	GPX_CUDA(cudaMemcpyAsync(&w->d_ch[world].vx, v, 3 * sizeof(float), cudaMemcpyHostToDevice, w->stream));
	return GPX_OK;
}

int gpx_character_set_position(gpx_world *w, uint32_t world, const float p[3])
{
	if (!w || world >= w->W || !w->d_ch || !p) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	GPX_CUDA(cudaMemcpyAsync(&w->d_ch[world].px, p, 3 * sizeof(float), cudaMemcpyHostToDevice, w->stream));
	return GPX_OK;
}

int gpx_character_update(gpx_world *w, float dt) { return gpx_character_update_ex(w, dt, nullptr); }

int gpx_character_update_ex(gpx_world *w, float dt, const gpx_character_update_settings *settings)
{
	if (!w || !w->d_ch || !(dt > 0.0f)) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	int rc;
	if (w->static_dirty && (rc = build_static(w)) != GPX_OK) return rc;
	if ((rc = flush_commands(w)) != GPX_OK) return rc;
	return launch_character(w, dt, settings);
}

int gpx_character_get(gpx_world *w, uint32_t world, gpx_character_state *out)
{
	if (!w || world >= w->W || !w->d_ch || !out) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	int rc = char_fetch(w, world);
	if (rc != GPX_OK) return rc;
	const CharDev &c = w->h_ch[world];
	out->position[0] = c.px; out->position[1] = c.py; out->position[2] = c.pz;
	out->linear_velocity[0] = c.vx; out->linear_velocity[1] = c.vy; out->linear_velocity[2] = c.vz;
	out->ground_normal[0] = c.gnx; out->ground_normal[1] = c.gny; out->ground_normal[2] = c.gnz;
	out->ground_state = c.ground;
	out->ground_body = c.ground_body;
	return c.alive ? GPX_OK : GPX_ERR_INVALID_ARG;
}

int gpx_character_contacts(gpx_world *w, uint32_t world, uint32_t *others, uint32_t capacity, uint32_t *count)
{
	if (!w || world >= w->W || !w->d_ch || !count || (capacity && !others)) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	unsigned long long keys[CHARACTER_MAX_CONTACTS];
	uint32_t n = 0;
	GPX_CUDA(cudaMemcpyAsync(&n, w->d_ch_nkeys + world, sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
	GPX_CUDA(cudaMemcpyAsync(keys, w->d_ch_keys + (size_t)world * CHARACTER_MAX_CONTACTS, sizeof(keys), cudaMemcpyDeviceToHost,
							 w->stream));
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	if (n > CHARACTER_MAX_CONTACTS) n = CHARACTER_MAX_CONTACTS;
	std::sort(keys, keys + n);
	n = (uint32_t)(std::unique(keys, keys + n) - keys);
	for (uint32_t i = 0; i < n && i < capacity; i++)
	{
		const uint32_t a = (uint32_t)(keys[i] >> 32), b = (uint32_t)(keys[i] & 0xFFFFFFFFull);
		others[i] = a == CHARACTER_BODY_ID ? b : a;
	}
	*count = n;
	return GPX_OK;
}

int gpx_events_enable(gpx_world *w, int enable)
{
	if (!w) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	if (w->wide) return wide_events_enable(w, enable != 0);
	if (enable && !w->d_ev_out)
	{
		const size_t nm = (size_t)w->W * (w->cap_m + CHARACTER_MAX_CONTACTS);
		int rc;
		if ((rc = dalloc(&w->d_ev_prev, nm)) != GPX_OK || (rc = dalloc(&w->d_ev_nprev, (size_t)w->W)) != GPX_OK ||
			(rc = dalloc(&w->d_ev_count, (size_t)w->W)) != GPX_OK || (rc = dalloc(&w->d_ev_out, 2 * nm)) != GPX_OK)
			return rc;
	}
	else if (!enable && w->d_ev_out)
	{
		cudaFree(w->d_ev_prev); cudaFree(w->d_ev_nprev); cudaFree(w->d_ev_count); cudaFree(w->d_ev_out);
		w->d_ev_prev = nullptr;
		w->d_ev_nprev = w->d_ev_count = nullptr;
		w->d_ev_out = nullptr;
	}
	return GPX_OK;
}

int gpx_poll_events(gpx_world *w, gpx_contact_event *out, uint64_t capacity, uint64_t *count)
{
	if (!w || !count || (capacity && !out)) return GPX_ERR_INVALID_ARG;
	*count = 0;
	std::lock_guard<std::mutex> lk(w->mu);
	if (!w->d_ev_out) return GPX_ERR_INVALID_ARG;
	cudaSetDevice(w->device);
	if (w->wide)
	{
		// one world: fetch the count, then only that many records
		uint32_t n = 0;
		GPX_CUDA(cudaMemcpyAsync(&n, w->d_ev_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
		GPX_CUDA(cudaStreamSynchronize(w->stream));
		const uint32_t cap_ev = 2u * wide_event_capacity(w);
		if (n > cap_ev) n = cap_ev;
		w->h_ev_out.resize(n);
		if (n) GPX_CUDA(cudaMemcpyAsync(w->h_ev_out.data(), w->d_ev_out, sizeof(uint4) * n, cudaMemcpyDeviceToHost, w->stream));
		GPX_CUDA(cudaStreamSynchronize(w->stream));
		// the device lists all added / persisted pairs, then all removed ones, each sorted by global body index; the host's
		// format is world after world, each with its added / persisted and then its removed pairs: a stable bucketing
		std::vector<uint32_t> at((size_t)w->W + 1, 0u);
		for (uint32_t k = 0; k < n; k++)
			if (w->h_ev_out[k].w < w->W) at[w->h_ev_out[k].w + 1]++;
		for (uint32_t wi = 0; wi < w->W; wi++) at[wi + 1] += at[wi];
		for (uint32_t k = 0; k < n; k++)
		{
			const uint4 ev = w->h_ev_out[k];
			if (ev.w >= w->W) continue;
			const uint32_t dst = at[ev.w]++;
			if (dst >= capacity) continue;
			out[dst].world = ev.w;
			out[dst].body_a = ev.x;
			out[dst].body_b = ev.y;
			out[dst].kind = ev.z;
		}
		*count = n;
		return n > capacity ? GPX_ERR_CAPACITY : GPX_OK;
	}
	const size_t per = 2u * ((size_t)w->cap_m + CHARACTER_MAX_CONTACTS);
	w->h_ev_count.resize(w->W);
	w->h_ev_out.resize((size_t)w->W * per);
	GPX_CUDA(cudaMemcpyAsync(w->h_ev_count.data(), w->d_ev_count, sizeof(uint32_t) * w->W, cudaMemcpyDeviceToHost, w->stream));
	GPX_CUDA(cudaMemcpyAsync(w->h_ev_out.data(), w->d_ev_out, sizeof(uint4) * w->h_ev_out.size(), cudaMemcpyDeviceToHost, w->stream));
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	uint64_t n = 0;
	for (uint32_t wi = 0; wi < w->W; wi++)
		for (uint32_t k = 0; k < w->h_ev_count[wi] && k < per; k++, n++)
			if (n < capacity)
			{
				const uint4 e = w->h_ev_out[wi * per + k];
				out[n].world = wi;
				out[n].body_a = e.x;
				out[n].body_b = e.y;
				out[n].kind = e.z;
			}
	*count = n;
	return n > capacity ? GPX_ERR_CAPACITY : GPX_OK;
}

int gpx_debug_wide_counters(gpx_world *w, uint32_t *out8)
{
	if (!w || !w->wide || !out8) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	return wide_counters(w, out8);
}

int gpx_debug_phase_cycles(gpx_world *w, int enable, uint64_t *out16)
{
	if (!w) return GPX_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(w->mu);
	cudaSetDevice(w->device);
	GPX_CUDA(cudaStreamSynchronize(w->stream));
	if (w->d_phase && out16) GPX_CUDA(cudaMemcpy(out16, w->d_phase, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
	if (enable && !w->d_phase) GPX_CUDA(cudaMalloc(&w->d_phase, 16 * sizeof(uint64_t)));
	if (w->d_phase) GPX_CUDA(cudaMemset(w->d_phase, 0, 16 * sizeof(uint64_t)));
	if (!enable && w->d_phase)
	{
		cudaFree(w->d_phase);
		w->d_phase = nullptr;
	}
	return GPX_OK;
}
int gpx_timer_begin(gpx_world *w)
{
	if (!w) return GPX_ERR_INVALID_ARG;
	GPX_CUDA(cudaEventRecord(w->ev0, w->stream));
	return GPX_OK;
}
float gpx_timer_end(gpx_world *w)
{
	if (!w) return -1.0f;
	if (cudaEventRecord(w->ev1, w->stream) != cudaSuccess) return -1.0f;
	if (cudaEventSynchronize(w->ev1) != cudaSuccess) return -1.0f;
	float ms = -1.0f;
	cudaEventElapsedTime(&ms, w->ev0, w->ev1);
	return ms;
}

}  // extern "C"
