// joltc_gpx.cpp — the joltc subset NBT22/c-game-engine calls (include/joltc_gpx.h), forwarded to libgpx's C ABI.
//
// Host code only: object bookkeeping (ref-counted shapes, settings objects, id -> body records), the evaluation of the
// engine's layer callbacks into bit masks, the character listener dispatch, and the host quaternion helpers.  All
// simulation and ray work happens in libgpx.so on the device; when no device is usable JPH_Init returns false and
// JPH_PhysicsSystem_Create returns NULL — there is no CPU path behind these entry points.
#include <algorithm>
#include "../../include/joltc_gpx.h"
#include "../../include/gpx.h"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace {

constexpr uint32_t N_LAYERS = 4;     // enum ObjectLayers, engine/include/engine/physics/Physics.h:36-42
constexpr uint32_t N_BP_LAYERS = 2;  // enum BroadPhaseLayers, Physics.h:44-51
constexpr uint32_t STATIC_BASE = 0x400000u;
// Public body ids carry a sequence number in the top byte, like Jolt's (index in the low 23 bits): a slot that is reused
// gets a new id, and calls made with the old one find nothing.
constexpr uint32_t ID_INDEX_MASK = 0x007FFFFFu;
constexpr uint32_t ID_SEQ_SHIFT = 24;
constexpr float MIN_HALF_EXTENT = 0.01f;  // a flat 4-point hull (ActorWall.c:20-49) becomes a 2 cm slab

void complain(const char *what) { fprintf(stderr, "joltc_gpx: %s\n", what); }

struct v3 { float x, y, z; };
inline v3 V(const Vector3 &a) { return {a.x, a.y, a.z}; }
inline v3 operator+(v3 a, v3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline v3 operator-(v3 a, v3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline v3 operator*(v3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline v3 cross(v3 a, v3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline v3 qrot(const JPH_Quat &q, v3 v)
{
	// v + 2 w (u x v) + 2 u x (u x v)
	const v3 u = {q.x, q.y, q.z};
	const v3 t = cross(u, v) * 2.0f;
	return v + t * q.w + cross(u, t);
}
inline JPH_Quat qmul(const JPH_Quat &a, const JPH_Quat &b)
{
	return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
			a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline bool is_identity(const JPH_Quat &q) { return fabsf(q.x) < 1e-6f && fabsf(q.y) < 1e-6f && fabsf(q.z) < 1e-6f; }

}  // namespace

/* ---- object definitions -------------------------------------------------------------------------------------------- */

enum ShapeKind { SK_EMPTY, SK_BOX, SK_HULL, SK_CYLINDER, SK_CAPSULE, SK_MESH, SK_COMPOUND };

struct JPH_Shape
{
	std::atomic<int> refs{1};
	ShapeKind kind = SK_EMPTY;
	v3 half{0, 0, 0};   // box half extents; cylinder/capsule: x = radius, y = half height
	float convex_radius = 0.0f;
	std::vector<float> points;  // hull: xyz...; mesh: 9 floats per triangle
	struct Child { v3 pos; JPH_Quat rot; JPH_Shape *shape; };
	std::vector<Child> children;
};

struct JPH_ShapeSettings
{
	ShapeKind kind = SK_EMPTY;
	std::vector<float> tris;
	std::vector<JPH_Shape::Child> children;
	~JPH_ShapeSettings()
	{
		for (auto &c : children) JPH_Shape_Destroy(c.shape);
	}
};

struct JPH_BodyCreationSettings
{
	JPH_Shape *shape = nullptr;
	Transform xfm{};
	JPH_MotionType motion = JPH_MotionType_Static;
	JPH_ObjectLayer layer = 0;
	uint64_t user_data = 0;
	float friction = 0.2f;  // Jolt default
	bool sensor = false;
	float mass = 0.0f;
	uint32_t dofs = JPH_AllowedDOFs_All;
};

struct JPH_JobSystem { int token; };
struct JPH_BroadPhaseLayerInterface { uint8_t bp_of_layer[N_LAYERS]; };
struct JPH_ObjectLayerPairFilter { uint8_t collide[N_LAYERS]; /* bit j of [i] */ };
struct JPH_ObjectVsBroadPhaseLayerFilter { uint8_t collide[N_LAYERS]; /* bit b of [i] */ };
struct JPH_BroadPhaseLayerFilter { uint32_t bp_mask; };
struct JPH_ObjectLayerFilter { uint32_t layer_mask; };
struct JPH_BodyFilter { JPH_BodyFilter_Impl impl; };
struct JPH_ShapeFilter { int unused; };
struct JPH_CharacterContactListener { JPH_CharacterContactListener_Impl impl; };
struct JPH_DebugRenderer { int unused; };
struct JPH_BodyDrawFilter { int unused; };

struct JPH_BodyInterface { JPH_PhysicsSystem *sys; };
struct JPH_NarrowPhaseQuery { JPH_PhysicsSystem *sys; };
struct JPH_Body { JPH_PhysicsSystem *sys; JPH_BodyID id; };

struct BodyRecord
{
	v3 center{0, 0, 0};  // offset of the device primitive's centre from the body origin, body space
	JPH_ObjectLayer layer = 0;
	uint64_t user_data = 0;
	JPH_Shape *shape = nullptr;  // the body's reference
	int ray_flag = -1;           // last value of the body filter pushed to the device, -1 = never evaluated
};

struct JPH_CharacterVirtual
{
	JPH_PhysicsSystem *sys = nullptr;
	uint64_t user_data = 0;
	JPH_CharacterContactListener *listener = nullptr;
	JPH_Quat rotation{0, 0, 0, 1};
	bool stale = true;  // cached state older than the last update / set
	gpx_character_state state{};
	std::vector<uint32_t> touching;  // device ids the character touched after the previous ExtendedUpdate (listener diff)
};

struct JPH_PhysicsSystem
{
	gpx_world *w = nullptr;
	JPH_BodyInterface bi{this};
	JPH_NarrowPhaseQuery npq{this};
	uint8_t bp_of_layer[N_LAYERS] = {0, 1, 1, 0};
	std::mutex mu;  // guards `bodies`
	std::unordered_map<JPH_BodyID, BodyRecord> bodies;        // keyed by public id
	std::unordered_map<uint32_t, uint32_t> slot_seq;           // device id -> uses so far
	std::unordered_map<uint32_t, JPH_BodyID> public_of;        // device id -> public id of the body living there
	JPH_BodyID to_public(uint32_t raw)
	{
		std::lock_guard<std::mutex> lk(mu);
		auto it = public_of.find(raw);
		return it == public_of.end() ? raw : it->second;
	}
	JPH_CharacterVirtual *character = nullptr;
	std::vector<gpx_contact_event> events;
	float gravity[3] = {0.0f, -9.81f, 0.0f};
	uint32_t max_bodies = 64, max_manifolds = 0;
	bool build_world();
};

static std::atomic<int> g_init{0};

bool JPH_PhysicsSystem::build_world()
{
	if (w) return true;
	gpx_world_config cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.worlds = 1;
	cfg.max_bodies_per_world = max_bodies;
	cfg.max_manifolds_per_world = max_manifolds;
	cfg.gravity[0] = gravity[0]; cfg.gravity[1] = gravity[1]; cfg.gravity[2] = gravity[2];
	const char *dev = getenv("GPX_DEVICE");
	cfg.device = dev ? atoi(dev) : 0;
	w = gpx_world_create(&cfg);
	if (!w)
	{
		fprintf(stderr, "joltc_gpx: gpx_world_create failed: %s\n", gpx_last_error());
		return false;
	}
	return true;
}

/* ---- host math ------------------------------------------------------------------------------------------------------ */

extern "C" {

const Vector3 Vector3_Zero = {0.0f, 0.0f, 0.0f}, Vector3_AxisX = {1.0f, 0.0f, 0.0f}, Vector3_AxisY = {0.0f, 1.0f, 0.0f},
			  Vector3_AxisZ = {0.0f, 0.0f, 1.0f};
const JPH_Quat JPH_Quat_Identity = {0.0f, 0.0f, 0.0f, 1.0f};

void Vector3_Add(const Vector3 *a, const Vector3 *b, Vector3 *out) { *out = {a->x + b->x, a->y + b->y, a->z + b->z}; }
void Vector3_Subtract(const Vector3 *a, const Vector3 *b, Vector3 *out) { *out = {a->x - b->x, a->y - b->y, a->z - b->z}; }
void Vector3_MultiplyScalar(const Vector3 *a, float s, Vector3 *out) { *out = {a->x * s, a->y * s, a->z * s}; }
float Vector3_LengthSquared(const Vector3 *a) { return a->x * a->x + a->y * a->y + a->z * a->z; }
float Vector3_Length(const Vector3 *a) { return sqrtf(Vector3_LengthSquared(a)); }
void Vector3_Normalized(const Vector3 *a, Vector3 *out)
{
	const float l = Vector3_Length(a);
	*out = {a->x / l, a->y / l, a->z / l};
}

void JPH_Quat_Rotation(const Vector3 *axis, float angle, JPH_Quat *out)
{
	const float s = sinf(0.5f * angle), c = cosf(0.5f * angle);
	*out = {axis->x * s, axis->y * s, axis->z * s, c};
}
void JPH_Quat_Rotate(const JPH_Quat *q, const Vector3 *v, Vector3 *out)
{
	const v3 r = qrot(*q, V(*v));
	*out = {r.x, r.y, r.z};
}
void JPH_Quat_RotateAxisZ(const JPH_Quat *q, Vector3 *out)
{
	const v3 r = qrot(*q, {0.0f, 0.0f, 1.0f});
	*out = {r.x, r.y, r.z};
}
// Twist angle about `axis`: 2 atan((q.xyz . axis) / q.w), pi when w == 0.
float JPH_Quat_GetRotationAngle(const JPH_Quat *q, const Vector3 *axis)
{
	if (q->w == 0.0f) return 3.14159265358979323846f;
	return 2.0f * atanf((q->x * axis->x + q->y * axis->y + q->z * axis->z) / q->w);
}
void JPH_Quat_Normalized(const JPH_Quat *q, JPH_Quat *out)
{
	const float l = sqrtf(q->x * q->x + q->y * q->y + q->z * q->z + q->w * q->w);
	*out = {q->x / l, q->y / l, q->z / l, q->w / l};
}
void JPH_Quat_Multiply(const JPH_Quat *a, const JPH_Quat *b, JPH_Quat *out) { *out = qmul(*a, *b); }
void JPH_Quat_Lerp(const JPH_Quat *from, const JPH_Quat *to, float t, JPH_Quat *out)
{
	const float s = 1.0f - t;
	*out = {s * from->x + t * to->x, s * from->y + t * to->y, s * from->z + t * to->z, s * from->w + t * to->w};
}
// Rotation about X, then Y, then Z: q = qz * qy * qx.
void JPH_Quat_FromEulerAngles(const Vector3 *a, JPH_Quat *out)
{
	const float cx = cosf(0.5f * a->x), sx = sinf(0.5f * a->x);
	const float cy = cosf(0.5f * a->y), sy = sinf(0.5f * a->y);
	const float cz = cosf(0.5f * a->z), sz = sinf(0.5f * a->z);
	*out = {cz * sx * cy - sz * cx * sy, cz * cx * sy + sz * sx * cy, sz * cx * cy - cz * sx * sy, cz * cx * cy + sz * sx * sy};
}
void JPH_Quat_GetEulerAngles(const JPH_Quat *q, Vector3 *out)
{
	const float y_sq = q->y * q->y;
	const float t0 = 2.0f * (q->w * q->x + q->y * q->z), t1 = 1.0f - 2.0f * (q->x * q->x + y_sq);
	float t2 = 2.0f * (q->w * q->y - q->z * q->x);
	t2 = t2 > 1.0f ? 1.0f : (t2 < -1.0f ? -1.0f : t2);
	const float t3 = 2.0f * (q->w * q->z + q->x * q->y), t4 = 1.0f - 2.0f * (y_sq + q->z * q->z);
	*out = {atan2f(t0, t1), asinf(t2), atan2f(t3, t4)};
}

/* ---- lifecycle ------------------------------------------------------------------------------------------------------- */

bool JPH_Init(void)
{
	const char *dev = getenv("GPX_DEVICE");
	const int rc = gpx_init(dev ? atoi(dev) : 0);
	if (rc < 0)
	{
		fprintf(stderr, "joltc_gpx: no usable CUDA device (%s)\n", gpx_last_error());
		return false;
	}
	g_init.store(1);
	return true;
}
void JPH_Shutdown(void)
{
	g_init.store(0);
	gpx_shutdown();
}
JPH_JobSystem *JPH_JobSystemThreadPool_Create(const JPH_JobSystemThreadPoolConfig *) { return new JPH_JobSystem{1}; }
void JPH_JobSystem_Destroy(JPH_JobSystem *js) { delete js; }

/* ---- layer tables ----------------------------------------------------------------------------------------------------- */

JPH_BroadPhaseLayerInterface *JPH_BroadPhaseLayerInterface_Create(uint32_t n, const JPH_BroadPhaseLayerInterface_Impl *impl)
{
	if (!impl || !impl->GetBroadPhaseLayer || n != N_BP_LAYERS)
	{
		complain("BroadPhaseLayerInterface: expected the engine's 2 broad-phase layers and a GetBroadPhaseLayer callback");
		return nullptr;
	}
	auto *o = new JPH_BroadPhaseLayerInterface;
	for (uint32_t l = 0; l < N_LAYERS; l++) o->bp_of_layer[l] = impl->GetBroadPhaseLayer(l);
	return o;
}
JPH_ObjectLayerPairFilter *JPH_ObjectLayerPairFilter_Create(const JPH_ObjectLayerPairFilter_Impl *impl)
{
	if (!impl || !impl->ShouldCollide) return nullptr;
	auto *o = new JPH_ObjectLayerPairFilter;
	for (uint32_t a = 0; a < N_LAYERS; a++)
	{
		o->collide[a] = 0;
		for (uint32_t b = 0; b < N_LAYERS; b++)
			if (impl->ShouldCollide(a, b)) o->collide[a] |= (uint8_t)(1u << b);
	}
	return o;
}
JPH_ObjectVsBroadPhaseLayerFilter *JPH_ObjectVsBroadPhaseLayerFilter_Create(const JPH_ObjectVsBroadPhaseLayerFilter_Impl *impl)
{
	if (!impl || !impl->ShouldCollide) return nullptr;
	auto *o = new JPH_ObjectVsBroadPhaseLayerFilter;
	for (uint32_t a = 0; a < N_LAYERS; a++)
	{
		o->collide[a] = 0;
		for (uint32_t b = 0; b < N_BP_LAYERS; b++)
			if (impl->ShouldCollide(a, (JPH_BroadPhaseLayer)b)) o->collide[a] |= (uint8_t)(1u << b);
	}
	return o;
}

/* ---- world --------------------------------------------------------------------------------------------------------------- */

JPH_PhysicsSystem *JPH_PhysicsSystem_Create(const JPH_PhysicsSystemSettings *s)
{
	if (!s || !g_init.load())
	{
		complain("PhysicsSystem_Create before a successful JPH_Init");
		return nullptr;
	}
	// The device's collision matrix is fixed to the engine's (Physics.c:35-60): DYNAMIC and PLAYER collide with STATIC,
	// DYNAMIC and SENSOR; STATIC and SENSOR initiate nothing; STATIC,SENSOR -> bp 0 and DYNAMIC,PLAYER -> bp 1.
	static const uint8_t want_pair[N_LAYERS] = {0x0, 0xB, 0xB, 0x0};
	static const uint8_t want_bp[N_LAYERS] = {0, 1, 1, 0};
	static const uint8_t want_vs_bp[N_LAYERS] = {0x0, 0x3, 0x3, 0x0};
	if (s->objectLayerPairFilter && memcmp(s->objectLayerPairFilter->collide, want_pair, N_LAYERS) != 0)
	{
		complain("ObjectLayerPairFilter differs from the collision matrix compiled into the device kernels");
		return nullptr;
	}
	if (s->broadPhaseLayerInterface && memcmp(s->broadPhaseLayerInterface->bp_of_layer, want_bp, N_LAYERS) != 0)
	{
		complain("BroadPhaseLayerInterface differs from the layer map compiled into the device kernels");
		return nullptr;
	}
	if (s->objectVsBroadPhaseLayerFilter && memcmp(s->objectVsBroadPhaseLayerFilter->collide, want_vs_bp, N_LAYERS) != 0)
	{
		complain("ObjectVsBroadPhaseLayerFilter differs from the layer map compiled into the device kernels");
		return nullptr;
	}
	auto *sys = new JPH_PhysicsSystem;
	const char *env = getenv("GPX_MAX_BODIES");
	// joltc's default when the engine leaves maxBodies at 0 (Physics.c:91-97 does): 10240 bodies, served by the wide-world
	// kernels; a map known to stay within 64 bodies can ask for the single fused kernel with maxBodies / GPX_MAX_BODIES <= 64
	sys->max_bodies = env ? (uint32_t)atoi(env) : (s->maxBodies ? s->maxBodies : 10240u);
	// The engine's MAX_CONTACT_CONSTRAINTS (16384) is a pool bound, not a need: the default sizing (3 per body) applies
	// unless the limit is smaller.
	sys->max_manifolds = 0;
	if (s->maxContactConstraints && s->maxContactConstraints < 3u * sys->max_bodies) sys->max_manifolds = s->maxContactConstraints;
	// The world itself is created with the first call that needs it, so that SetGravity (Physics.c:99) still lands in
	// its configuration.
	return sys;
}

void JPH_PhysicsSystem_Destroy(JPH_PhysicsSystem *sys)
{
	if (!sys) return;
	for (auto &kv : sys->bodies)
		if (kv.second.shape) JPH_Shape_Destroy(kv.second.shape);
	if (sys->character) sys->character->sys = nullptr;
	if (sys->w) gpx_world_destroy(sys->w);
	delete sys;
}

void JPH_PhysicsSystem_SetGravity(JPH_PhysicsSystem *sys, const Vector3 *g)
{
	if (!sys || !g) return;
	if (sys->w)
	{
		if (g->x != sys->gravity[0] || g->y != sys->gravity[1] || g->z != sys->gravity[2])
			complain("SetGravity after the first body is ignored (gravity is part of the world configuration)");
		return;
	}
	sys->gravity[0] = g->x; sys->gravity[1] = g->y; sys->gravity[2] = g->z;
}

JPH_BodyInterface *JPH_PhysicsSystem_GetBodyInterface(const JPH_PhysicsSystem *sys)
{
	return sys ? const_cast<JPH_BodyInterface *>(&sys->bi) : nullptr;
}
const JPH_NarrowPhaseQuery *JPH_PhysicsSystem_GetNarrowPhaseQuery(const JPH_PhysicsSystem *sys) { return sys ? &sys->npq : nullptr; }

void JPH_PhysicsSystem_OptimizeBroadPhase(JPH_PhysicsSystem *sys)
{
	if (sys && sys->build_world() && gpx_static_commit(sys->w) != GPX_OK)
		fprintf(stderr, "joltc_gpx: static commit failed: %s\n", gpx_last_error());
}

static void deliver_character_events(JPH_PhysicsSystem *sys)
{
	JPH_CharacterVirtual *ch = sys->character;
	if (!ch) return;
	uint64_t n = 0;
	sys->events.resize(256);
	if (gpx_poll_events(sys->w, sys->events.data(), sys->events.size(), &n) != GPX_OK) return;
	if (n > sys->events.size())
	{
		sys->events.resize(n);
		gpx_poll_events(sys->w, sys->events.data(), sys->events.size(), &n);
	}
	// The character listener's callbacks already ran inside JPH_CharacterVirtual_ExtendedUpdate (character_callbacks
	// below), where Jolt runs them; the tick's event list, which repeats the character's contacts next to the bodies',
	// is only drained here.  (The engine registers no body-body ContactListener.)
	ch->stale = true;
}

// CharacterContactListener callbacks of one ExtendedUpdate (PlayerPhysics.c:89-152), on the caller's thread and with no
// lock held — handlers create and destroy bodies (Coin.c:85), and a body destroyed here is gone before the
// JPH_PhysicsSystem_Update of the same MapFixedUpdate, as with Jolt.  Order: contacts that exist now (added or
// persisted) by id, bodies before map meshes, then the ones that ended.
static void character_callbacks(JPH_CharacterVirtual *ch)
{
	JPH_PhysicsSystem *sys = ch->sys;
	uint32_t now[64], n = 0;
	if (gpx_character_contacts(sys->w, 0, now, 64, &n) != GPX_OK) return;
	if (n > 64) n = 64;
	std::vector<uint32_t> before;
	before.swap(ch->touching);
	ch->touching.assign(now, now + n);
	if (!ch->listener) return;
	const JPH_CharacterContactListener_Impl &cb = ch->listener->impl;
	gpx_character_state cs;
	memset(&cs, 0, sizeof(cs));
	gpx_character_get(sys->w, 0, &cs);
	const JPH_RVec3 pos = {cs.position[0], cs.position[1], cs.position[2]};
	const Vector3 nrm = {cs.ground_normal[0], cs.ground_normal[1], cs.ground_normal[2]};
	auto rank = [](uint32_t id) { return id >= GPX_STATIC_BODY_BASE ? (1ull << 32) | id : (uint64_t)id; };
	for (uint32_t i = 0; i < n; i++)
	{
		const bool was = std::find(before.begin(), before.end(), now[i]) != before.end();
		const JPH_BodyID other = sys->to_public(now[i]);
		if (cb.OnContactValidate && !cb.OnContactValidate(ch, other, 0)) continue;
		JPH_CharacterContactSettings io = {true, true};
		if (!was && cb.OnContactAdded) cb.OnContactAdded(ch, other, 0, &pos, &nrm, &io);
		if (was && cb.OnContactPersisted) cb.OnContactPersisted(ch, other, 0, &pos, &nrm, &io);
	}
	std::sort(before.begin(), before.end(), [&](uint32_t a, uint32_t b) { return rank(a) < rank(b); });
	for (uint32_t id : before)
		if (std::find(now, now + n, id) == now + n && cb.OnContactRemoved) cb.OnContactRemoved(ch, sys->to_public(id), 0);
}

JPH_PhysicsUpdateError JPH_PhysicsSystem_Update(JPH_PhysicsSystem *sys, float dt, int collisionSteps, JPH_JobSystem *)
{
	if (!sys || !sys->build_world()) return JPH_PhysicsUpdateError_ContactConstraintsFull;
	int rc = gpx_step(sys->w, dt, collisionSteps);
	if (rc == GPX_OK) rc = gpx_sync_transforms(sys->w);
	if (rc != GPX_OK)
	{
		fprintf(stderr, "joltc_gpx: tick failed with %d (%s)\n", rc, gpx_last_error());
		// tick errors are Jolt's own bits (1, 2, 4); library errors (CUDA, capacity, arguments) have no Jolt counterpart and
		// are reported as "contact constraints full", the engine treats any non-zero value as fatal (MapPhysics.c:109-113)
		const bool tick_error = rc > 0 && rc < 8;
		return (JPH_PhysicsUpdateError)(tick_error ? rc : JPH_PhysicsUpdateError_ContactConstraintsFull);
	}
	deliver_character_events(sys);
	return JPH_PhysicsUpdateError_None;
}

/* ---- shapes ---------------------------------------------------------------------------------------------------------------- */

JPH_BoxShape *JPH_BoxShape_Create(const Vector3 *half, float convexRadius)
{
	if (!half) return nullptr;
	auto *s = new JPH_Shape;
	s->kind = SK_BOX;
	s->half = V(*half);
	s->convex_radius = convexRadius;
	return s;
}
JPH_ConvexHullShape *JPH_ConvexHullShape_Create(const Vector3 *points, uint32_t count, float maxConvexRadius)
{
	if (!points || count < 3) return nullptr;
	auto *s = new JPH_Shape;
	s->kind = SK_HULL;
	s->convex_radius = maxConvexRadius;
	s->points.assign(&points[0].x, &points[0].x + 3ull * count);
	return s;
}
JPH_CylinderShape *JPH_CylinderShape_Create(float halfHeight, float radius)
{
	auto *s = new JPH_Shape;
	s->kind = SK_CYLINDER;
	s->half = {radius, halfHeight, radius};
	return s;
}
JPH_CapsuleShape *JPH_CapsuleShape_Create(float halfHeight, float radius)
{
	auto *s = new JPH_Shape;
	s->kind = SK_CAPSULE;
	s->half = {radius, halfHeight, radius};
	return s;
}
JPH_MeshShapeSettings *JPH_MeshShapeSettings_Create(const JPH_Triangle *tris, uint32_t count)
{
	if (!tris && count) return nullptr;
	auto *s = new JPH_ShapeSettings;
	s->kind = SK_MESH;
	s->tris.reserve(9ull * count);
	for (uint32_t i = 0; i < count; i++)
	{
		const Vector3 *v[3] = {&tris[i].v1, &tris[i].v2, &tris[i].v3};
		for (int k = 0; k < 3; k++)
		{
			s->tris.push_back(v[k]->x);
			s->tris.push_back(v[k]->y);
			s->tris.push_back(v[k]->z);
		}
	}
	return s;
}
JPH_MeshShape *JPH_MeshShapeSettings_CreateShape(const JPH_MeshShapeSettings *settings)
{
	if (!settings || settings->kind != SK_MESH) return nullptr;
	auto *s = new JPH_Shape;
	s->kind = SK_MESH;
	s->points = settings->tris;
	return s;
}
JPH_StaticCompoundShapeSettings *JPH_StaticCompoundShapeSettings_Create(void)
{
	auto *s = new JPH_ShapeSettings;
	s->kind = SK_COMPOUND;
	return s;
}
void JPH_CompoundShapeSettings_AddShape2(JPH_CompoundShapeSettings *settings, const Vector3 *position, const JPH_Quat *rotation,
										 const JPH_Shape *shape, uint32_t)
{
	if (!settings || settings->kind != SK_COMPOUND || !shape) return;
	auto *sh = const_cast<JPH_Shape *>(shape);
	sh->refs.fetch_add(1);
	settings->children.push_back({position ? V(*position) : v3{0, 0, 0}, rotation ? *rotation : JPH_Quat_Identity, sh});
}
JPH_StaticCompoundShape *JPH_StaticCompoundShape_Create(const JPH_StaticCompoundShapeSettings *settings)
{
	if (!settings || settings->kind != SK_COMPOUND) return nullptr;
	auto *s = new JPH_Shape;
	s->kind = SK_COMPOUND;
	s->children = settings->children;
	for (auto &c : s->children) c.shape->refs.fetch_add(1);
	return s;
}
JPH_EmptyShapeSettings *JPH_EmptyShapeSettings_Create(const Vector3 *)
{
	auto *s = new JPH_ShapeSettings;
	s->kind = SK_EMPTY;
	return s;
}
void JPH_ShapeSettings_Destroy(JPH_ShapeSettings *settings) { delete settings; }
void JPH_Shape_Destroy(JPH_Shape *shape)
{
	if (!shape) return;
	if (shape->refs.fetch_sub(1) == 1)
	{
		for (auto &c : shape->children) JPH_Shape_Destroy(c.shape);
		delete shape;
	}
}

}  // extern "C"

/* ---- lowering a shape tree onto what the device holds -------------------------------------------------------------- */

namespace {

struct Lowered
{
	uint32_t shape = GPX_SHAPE_EMPTY;  // device primitive, or mesh (tris non-empty)
	v3 half{0, 0, 0};
	v3 center{0, 0, 0};
	bool exact = true;
	std::vector<float> tris;  // body space
	std::vector<float> hull_points;  // body space, all convex children
	uint32_t convex_parts = 0;
};

void collect(const JPH_Shape *s, v3 pos, const JPH_Quat &rot, Lowered &out)
{
	switch (s->kind)
	{
		case SK_EMPTY: break;
		case SK_MESH:
			for (size_t i = 0; i + 2 < s->points.size(); i += 3)
			{
				const v3 p = qrot(rot, {s->points[i], s->points[i + 1], s->points[i + 2]}) + pos;
				out.tris.push_back(p.x); out.tris.push_back(p.y); out.tris.push_back(p.z);
			}
			break;
		case SK_COMPOUND:
			for (const auto &c : s->children) collect(c.shape, qrot(rot, c.pos) + pos, qmul(rot, c.rot), out);
			break;
		case SK_HULL:
			out.convex_parts++;
			for (size_t i = 0; i + 2 < s->points.size(); i += 3)
			{
				const v3 p = qrot(rot, {s->points[i], s->points[i + 1], s->points[i + 2]}) + pos;
				out.hull_points.push_back(p.x); out.hull_points.push_back(p.y); out.hull_points.push_back(p.z);
			}
			break;
		case SK_BOX:
		case SK_CYLINDER:
		case SK_CAPSULE:
		{
			// the 8 corners of the bounding box stand for the primitive; a lone, unrotated box stays exact (below)
			out.convex_parts++;
			for (int k = 0; k < 8; k++)
			{
				const v3 c = {(k & 1) ? s->half.x : -s->half.x, (k & 2) ? s->half.y : -s->half.y, (k & 4) ? s->half.z : -s->half.z};
				const v3 p = qrot(rot, c) + pos;
				out.hull_points.push_back(p.x); out.hull_points.push_back(p.y); out.hull_points.push_back(p.z);
			}
			if (s->kind != SK_BOX || !is_identity(rot)) out.exact = false;
			break;
		}
	}
}

bool lower(const JPH_Shape *s, Lowered &out)
{
	collect(s, {0, 0, 0}, JPH_Quat_Identity, out);
	if (!out.tris.empty())
	{
		if (out.convex_parts) out.exact = false;  // convex parts next to triangles are dropped
		return true;
	}
	if (out.convex_parts == 0) return true;  // empty
	gpx_hull_shape hs;
	if (gpx_shape_from_hull(out.hull_points.data(), out.hull_points.size() / 3, 0.02f, &hs) != GPX_OK) return false;
	out.shape = hs.shape;
	out.half = {hs.half_extents[0], hs.half_extents[1], hs.half_extents[2]};
	out.center = {hs.center[0], hs.center[1], hs.center[2]};
	if (!hs.exact || out.convex_parts > 1) out.exact = false;
	if (hs.shape == GPX_SHAPE_BOX)
	{
		float *h[3] = {&out.half.x, &out.half.y, &out.half.z};
		for (float *e : h)
			if (*e < MIN_HALF_EXTENT) *e = MIN_HALF_EXTENT;
	}
	return true;
}

JPH_BodyID make_body(JPH_PhysicsSystem *sys, const JPH_BodyCreationSettings *st)
{
	if (!sys->build_world()) return JPH_BodyId_InvalidBodyID;
	Lowered lo;
	if (st->shape && !lower(st->shape, lo)) return JPH_BodyId_InvalidBodyID;
	BodyRecord rec;
	rec.layer = st->layer;
	rec.user_data = st->user_data;
	JPH_BodyID id;
	if (!lo.tris.empty())
	{
		if (st->motion != JPH_MotionType_Static)
		{
			complain("a triangle mesh on a moving body is not representable; body not created");
			return JPH_BodyId_InvalidBodyID;
		}
		if (st->layer != 0) complain("mesh bodies live on OBJECT_LAYER_STATIC; the requested layer is ignored");
		gpx_transform x;
		memcpy(x.position, &st->xfm.position, sizeof(float) * 3);
		memcpy(x.rotation, &st->xfm.rotation, sizeof(float) * 4);
		uint32_t out = GPX_INVALID_BODY;
		if (gpx_static_add_mesh(sys->w, &x, lo.tris.data(), lo.tris.size() / 9, st->friction, st->user_data, &out) != GPX_OK)
			return JPH_BodyId_InvalidBodyID;
		id = out;
	}
	else
	{
		gpx_body_desc d;
		memset(&d, 0, sizeof(d));
		d.shape = lo.shape;
		d.half_extents[0] = lo.half.x; d.half_extents[1] = lo.half.y; d.half_extents[2] = lo.half.z;
		d.convex_radius = st->shape ? st->shape->convex_radius : 0.0f;
		const v3 c = qrot(st->xfm.rotation, lo.center) + V(st->xfm.position);
		d.position[0] = c.x; d.position[1] = c.y; d.position[2] = c.z;
		memcpy(d.rotation, &st->xfm.rotation, sizeof(float) * 4);
		d.motion_type = (uint32_t)st->motion;
		d.layer = st->layer;
		d.mass = st->mass;
		d.friction = st->friction;
		d.restitution = 0.0f;
		d.linear_damping = 0.05f;
		d.angular_damping = 0.05f;
		d.gravity_factor = 1.0f;
		d.is_sensor = st->sensor ? 1u : 0u;
		d.allowed_dofs = st->dofs;
		d.allow_sleeping = 1;
		d.ray_flags = st->user_data == 0 ? GPX_BODY_BLOCKS_LASERS : 0u;  // refined by the body filter before filtered casts
		d.user_data = st->user_data;
		id = gpx_body_create(sys->w, 0, &d);
		if (id == GPX_INVALID_BODY)
		{
			complain("CreateAndAddBody: no free body slot (JPH_PhysicsSystemSettings.maxBodies / GPX_MAX_BODIES too small)");
			return JPH_BodyId_InvalidBodyID;
		}
		rec.center = lo.center;
	}
	if (st->shape)
	{
		rec.shape = st->shape;
		st->shape->refs.fetch_add(1);
	}
	std::lock_guard<std::mutex> lk(sys->mu);
	const uint32_t seq = sys->slot_seq[id]++ % 255u;  // 255 is left out so no id can equal JPH_BodyId_InvalidBodyID
	const JPH_BodyID pub = id | (seq << ID_SEQ_SHIFT);
	sys->bodies[pub] = rec;
	sys->public_of[id] = pub;
	return pub;
}

// The device slot behind a public id, if that body is still alive and lives in a body slot (not the static soup).
bool live_slot(JPH_PhysicsSystem *sys, JPH_BodyID id, uint32_t &raw)
{
	raw = id & ID_INDEX_MASK;
	if (!sys->w || raw >= STATIC_BASE) return false;
	std::lock_guard<std::mutex> lk(sys->mu);
	return sys->bodies.find(id) != sys->bodies.end();
}

// Body origin and rotation from the device primitive's centre.
bool body_pose(JPH_PhysicsSystem *sys, JPH_BodyID id, v3 &pos, JPH_Quat &rot)
{
	v3 center{0, 0, 0};
	{
		std::lock_guard<std::mutex> lk(sys->mu);
		auto it = sys->bodies.find(id);
		if (it == sys->bodies.end() || !sys->w) return false;
		center = it->second.center;
	}
	// The mirror is written by Update's readback and, in between, by gpx_body_create and the setters themselves.
	gpx_transform t;
	if (gpx_body_get_transform(sys->w, 0, id & ID_INDEX_MASK, &t) != GPX_OK) return false;
	rot = {t.rotation[0], t.rotation[1], t.rotation[2], t.rotation[3]};
	pos = v3{t.position[0], t.position[1], t.position[2]} - qrot(rot, center);
	return true;
}

}  // namespace

extern "C" {

int JPH_GPX_ShapeIsExact(const JPH_Shape *shape)
{
	if (!shape) return 0;
	Lowered lo;
	return lower(shape, lo) && lo.exact ? 1 : 0;
}

/* ---- body creation settings ------------------------------------------------------------------------------------------ */

JPH_BodyCreationSettings *JPH_BodyCreationSettings_Create2_GAME(const JPH_Shape *shape, const Transform *xfm, JPH_MotionType motion,
																JPH_ObjectLayer layer, void *userData)
{
	if (!xfm) return nullptr;
	auto *s = new JPH_BodyCreationSettings;
	s->shape = const_cast<JPH_Shape *>(shape);
	if (s->shape) s->shape->refs.fetch_add(1);
	s->xfm = *xfm;
	s->motion = motion;
	s->layer = layer;
	s->user_data = (uint64_t)(uintptr_t)userData;
	return s;
}
JPH_BodyCreationSettings *JPH_BodyCreationSettings_Create_GAME(const JPH_ShapeSettings *shapeSettings, const Transform *xfm, JPH_MotionType motion,
															   JPH_ObjectLayer layer, void *userData)
{
	JPH_Shape *shape = nullptr;
	if (shapeSettings && shapeSettings->kind == SK_MESH) shape = JPH_MeshShapeSettings_CreateShape(shapeSettings);
	else if (shapeSettings && shapeSettings->kind == SK_COMPOUND) shape = JPH_StaticCompoundShape_Create(shapeSettings);
	else
	{
		shape = new JPH_Shape;  // empty (Actor.c:153, Laser.c:114)
		shape->kind = SK_EMPTY;
	}
	JPH_BodyCreationSettings *s = JPH_BodyCreationSettings_Create2_GAME(shape, xfm, motion, layer, userData);
	JPH_Shape_Destroy(shape);
	return s;
}
void JPH_BodyCreationSettings_Destroy(JPH_BodyCreationSettings *s)
{
	if (!s) return;
	JPH_Shape_Destroy(s->shape);
	delete s;
}
void JPH_BodyCreationSettings_SetFriction(JPH_BodyCreationSettings *s, float friction) { if (s) s->friction = friction; }
void JPH_BodyCreationSettings_SetIsSensor(JPH_BodyCreationSettings *s, bool sensor) { if (s) s->sensor = sensor; }
void JPH_BodyCreationSettings_SetMassPropertiesOverride(JPH_BodyCreationSettings *s, const JPH_MassProperties *mp)
{
	if (s && mp) s->mass = mp->mass;
}
void JPH_BodyCreationSettings_SetOverrideMassProperties(JPH_BodyCreationSettings *s, JPH_OverrideMassProperties mode)
{
	// CalculateInertia (the only mode the reference uses): keep the given mass, scale the shape's inertia to it.
	if (s && mode == JPH_OverrideMassProperties_CalculateMassAndInertia) s->mass = 0.0f;
}
void JPH_BodyCreationSettings_SetAllowedDOFs(JPH_BodyCreationSettings *s, JPH_AllowedDOFs dofs) { if (s) s->dofs = (uint32_t)dofs; }

/* ---- body interface ----------------------------------------------------------------------------------------------------- */

JPH_BodyID JPH_BodyInterface_CreateAndAddBody(JPH_BodyInterface *bi, const JPH_BodyCreationSettings *settings, JPH_Activation)
{
	if (!bi || !bi->sys || !settings) return JPH_BodyId_InvalidBodyID;
	return make_body(bi->sys, settings);
}

void JPH_BodyInterface_RemoveAndDestroyBody(JPH_BodyInterface *bi, JPH_BodyID id)
{
	if (!bi || !bi->sys || !bi->sys->w) return;
	JPH_PhysicsSystem *sys = bi->sys;
	JPH_Shape *shape = nullptr;
	{
		std::lock_guard<std::mutex> lk(sys->mu);
		auto it = sys->bodies.find(id);
		if (it == sys->bodies.end()) return;
		shape = it->second.shape;
		sys->bodies.erase(it);  // public_of keeps the last id of the slot: a late 'removed' event still names this body
	}
	const uint32_t raw = id & ID_INDEX_MASK;
	if (raw >= STATIC_BASE) gpx_static_remove_mesh(sys->w, raw);
	else gpx_body_destroy(sys->w, 0, raw);
	JPH_Shape_Destroy(shape);
}

void JPH_BodyInterface_GetPosition(JPH_BodyInterface *bi, JPH_BodyID id, JPH_RVec3 *out)
{
	v3 p;
	JPH_Quat q;
	if (out && bi && bi->sys && body_pose(bi->sys, id, p, q)) *out = {p.x, p.y, p.z};
}
void JPH_BodyInterface_GetRotation(JPH_BodyInterface *bi, JPH_BodyID id, JPH_Quat *out)
{
	v3 p;
	JPH_Quat q;
	if (out && bi && bi->sys && body_pose(bi->sys, id, p, q)) *out = q;
}
void JPH_BodyInterface_GetPositionAndRotation(JPH_BodyInterface *bi, JPH_BodyID id, JPH_RVec3 *position, JPH_Quat *rotation)
{
	v3 p;
	JPH_Quat q;
	if (!bi || !bi->sys || !body_pose(bi->sys, id, p, q)) return;
	if (position) *position = {p.x, p.y, p.z};
	if (rotation) *rotation = q;
}
void JPH_BodyInterface_GetWorldTransform(JPH_BodyInterface *bi, JPH_BodyID id, JPH_RMat44 *out)
{
	v3 p;
	JPH_Quat q;
	if (!out || !bi || !bi->sys || !body_pose(bi->sys, id, p, q)) return;
	const v3 c0 = qrot(q, {1, 0, 0}), c1 = qrot(q, {0, 1, 0}), c2 = qrot(q, {0, 0, 1});
	const float m[16] = {c0.x, c0.y, c0.z, 0.0f, c1.x, c1.y, c1.z, 0.0f, c2.x, c2.y, c2.z, 0.0f, p.x, p.y, p.z, 1.0f};
	memcpy(out->m, m, sizeof(m));
}
uint64_t JPH_BodyInterface_GetUserData(JPH_BodyInterface *bi, JPH_BodyID id)
{
	if (!bi || !bi->sys) return 0;
	std::lock_guard<std::mutex> lk(bi->sys->mu);
	auto it = bi->sys->bodies.find(id);
	return it == bi->sys->bodies.end() ? 0 : it->second.user_data;
}
void JPH_BodyInterface_SetLinearVelocity(JPH_BodyInterface *bi, JPH_BodyID id, const Vector3 *v)
{
	uint32_t raw;
	if (bi && bi->sys && v && live_slot(bi->sys, id, raw)) gpx_body_set_linear_velocity(bi->sys->w, 0, raw, &v->x);
}
void JPH_BodyInterface_SetLinearAndAngularVelocity(JPH_BodyInterface *bi, JPH_BodyID id, const Vector3 *lin, const Vector3 *ang)
{
	uint32_t raw;
	if (bi && bi->sys && lin && ang && live_slot(bi->sys, id, raw))
		gpx_body_set_linear_and_angular_velocity(bi->sys->w, 0, raw, &lin->x, &ang->x);
}
void JPH_BodyInterface_SetPosition(JPH_BodyInterface *bi, JPH_BodyID id, const JPH_RVec3 *position, JPH_Activation activation)
{
	if (!bi || !bi->sys || !bi->sys->w || !position || (id & ID_INDEX_MASK) >= STATIC_BASE) return;
	v3 p;
	JPH_Quat q;
	v3 center{0, 0, 0};
	{
		std::lock_guard<std::mutex> lk(bi->sys->mu);
		auto it = bi->sys->bodies.find(id);
		if (it == bi->sys->bodies.end()) return;
		center = it->second.center;
	}
	if (!body_pose(bi->sys, id, p, q)) return;
	const v3 c = V(*position) + qrot(q, center);
	gpx_body_set_position(bi->sys->w, 0, id & ID_INDEX_MASK, &c.x, activation == JPH_Activation_Activate);
}
void JPH_BodyInterface_SetRotation(JPH_BodyInterface *bi, JPH_BodyID id, const JPH_Quat *rotation, JPH_Activation activation)
{
	if (!bi || !bi->sys || !bi->sys->w || !rotation || (id & ID_INDEX_MASK) >= STATIC_BASE) return;
	v3 p;
	JPH_Quat q;
	v3 center{0, 0, 0};
	{
		std::lock_guard<std::mutex> lk(bi->sys->mu);
		auto it = bi->sys->bodies.find(id);
		if (it == bi->sys->bodies.end()) return;
		center = it->second.center;
	}
	if (!body_pose(bi->sys, id, p, q)) return;
	gpx_body_set_rotation(bi->sys->w, 0, id & ID_INDEX_MASK, &rotation->x, activation == JPH_Activation_Activate);
	if (center.x != 0.0f || center.y != 0.0f || center.z != 0.0f)
	{
		// the body turns about its origin, so an off-centre primitive moves
		JPH_Quat nq;
		JPH_Quat_Normalized(rotation, &nq);
		const v3 c = p + qrot(nq, center);
		gpx_body_set_position(bi->sys->w, 0, id & ID_INDEX_MASK, &c.x, activation == JPH_Activation_Activate);
	}
}
uint64_t JPH_Body_GetUserData(const JPH_Body *body) { return body ? JPH_BodyInterface_GetUserData(&body->sys->bi, body->id) : 0; }
JPH_ObjectLayer JPH_Body_GetObjectLayer(const JPH_Body *body)
{
	if (!body) return 0;
	std::lock_guard<std::mutex> lk(body->sys->mu);
	auto it = body->sys->bodies.find(body->id);
	return it == body->sys->bodies.end() ? 0 : it->second.layer;
}

/* ---- ray queries ---------------------------------------------------------------------------------------------------------- */

JPH_BroadPhaseLayerFilter *JPH_BroadPhaseLayerFilter_Create(const JPH_BroadPhaseLayerFilter_Impl *impl)
{
	auto *f = new JPH_BroadPhaseLayerFilter{0};
	for (uint32_t b = 0; b < N_BP_LAYERS; b++)
		if (!impl || !impl->ShouldCollide || impl->ShouldCollide((JPH_BroadPhaseLayer)b)) f->bp_mask |= 1u << b;
	return f;
}
void JPH_BroadPhaseLayerFilter_Destroy(JPH_BroadPhaseLayerFilter *f) { delete f; }
JPH_ObjectLayerFilter *JPH_ObjectLayerFilter_Create(const JPH_ObjectLayerFilter_Impl *impl)
{
	auto *f = new JPH_ObjectLayerFilter{0};
	for (uint32_t l = 0; l < N_LAYERS; l++)
		if (!impl || !impl->ShouldCollide || impl->ShouldCollide(l)) f->layer_mask |= 1u << l;
	return f;
}
void JPH_ObjectLayerFilter_Destroy(JPH_ObjectLayerFilter *f) { delete f; }
JPH_BodyFilter *JPH_BodyFilter_Create(const JPH_BodyFilter_Impl *impl)
{
	auto *f = new JPH_BodyFilter;
	memset(&f->impl, 0, sizeof(f->impl));
	if (impl) f->impl = *impl;
	return f;
}
void JPH_BodyFilter_Destroy(JPH_BodyFilter *f) { delete f; }
JPH_ShapeFilter *JPH_ShapeFilter_Create(const JPH_ShapeFilter_Impl *) { return new JPH_ShapeFilter{0}; }
void JPH_ShapeFilter_Destroy(JPH_ShapeFilter *f) { delete f; }

}  // extern "C"

namespace {

// A layer passes when its object-layer callback said yes AND the broad-phase layer it lives in passes.
uint32_t ray_mask(const JPH_PhysicsSystem *sys, const JPH_BroadPhaseLayerFilter *bp, const JPH_ObjectLayerFilter *ol)
{
	uint32_t m = 0;
	for (uint32_t l = 0; l < N_LAYERS; l++)
	{
		const bool bp_ok = !bp || (bp->bp_mask >> sys->bp_of_layer[l]) & 1u;
		const bool ol_ok = !ol || (ol->layer_mask >> l) & 1u;
		if (bp_ok && ol_ok) m |= 1u << l;
	}
	return m;
}

// Laser.c:74-85: the filter reads actor->flags through the body's user data, so it is asked again for every live body
// and only changes travel to the device.
void refresh_body_filter(JPH_PhysicsSystem *sys, const JPH_BodyFilter *f)
{
	std::vector<JPH_BodyID> ids;
	{
		std::lock_guard<std::mutex> lk(sys->mu);
		ids.reserve(sys->bodies.size());
		for (auto &kv : sys->bodies) ids.push_back(kv.first);
	}
	for (JPH_BodyID id : ids)
	{
		bool pass = true;
		if (f->impl.ShouldCollide) pass = f->impl.ShouldCollide(id);
		else if (f->impl.ShouldCollideLocked)
		{
			JPH_Body b{sys, id};
			pass = f->impl.ShouldCollideLocked(&b);
		}
		std::lock_guard<std::mutex> lk(sys->mu);
		auto it = sys->bodies.find(id);
		if (it == sys->bodies.end() || it->second.ray_flag == (int)pass) continue;
		it->second.ray_flag = (int)pass;
		gpx_body_set_ray_flags(sys->w, 0, id & ID_INDEX_MASK, pass ? GPX_BODY_BLOCKS_LASERS : 0u);
	}
}

bool cast(JPH_PhysicsSystem *sys, const Transform *origin, float maxDistance, uint32_t mask, JPH_RayCastResult *result, Vector3 *offset)
{
	gpx_transform t;
	memcpy(t.position, &origin->position, sizeof(float) * 3);
	memcpy(t.rotation, &origin->rotation, sizeof(float) * 4);
	gpx_hit h;
	if (gpx_raycast_transform(sys->w, 0, &t, maxDistance, mask, &h) != GPX_OK) return false;
	if (h.body == GPX_INVALID_BODY) return false;
	if (result)
	{
		result->bodyID = sys->to_public(h.body);
		result->fraction = h.fraction;
		result->subShapeID2 = h.face;
	}
	if (offset)
	{
		const v3 d = qrot(origin->rotation, {0.0f, 0.0f, -1.0f}) * (h.fraction * maxDistance);
		*offset = {d.x, d.y, d.z};
	}
	return true;
}

}  // namespace

extern "C" {

bool JPH_NarrowPhaseQuery_CastRay_GAME(const JPH_NarrowPhaseQuery *query, const Transform *origin, float maxDistance, JPH_RayCastResult *result,
									   const JPH_BroadPhaseLayerFilter *bp, const JPH_ObjectLayerFilter *ol)
{
	if (!query || !query->sys || !origin || !query->sys->build_world()) return false;
	return cast(query->sys, origin, maxDistance, ray_mask(query->sys, bp, ol), result, nullptr);
}

bool JPH_NarrowPhaseQuery_CastRay2_GAME(const JPH_NarrowPhaseQuery *query, JPH_BodyInterface *bi, JPH_BodyID originBody, float maxDistance,
										JPH_RayCastResult *result, Vector3 *hitPointOffset, const JPH_BroadPhaseLayerFilter *bp,
										const JPH_ObjectLayerFilter *ol, const JPH_BodyFilter *bodyFilter)
{
	if (!query || !query->sys || !bi || !query->sys->build_world()) return false;
	JPH_PhysicsSystem *sys = query->sys;
	Transform origin;
	v3 p;
	if (!body_pose(sys, originBody, p, origin.rotation)) return false;
	origin.position = {p.x, p.y, p.z};
	uint32_t mask = ray_mask(sys, bp, ol);
	if (bodyFilter && (bodyFilter->impl.ShouldCollide || bodyFilter->impl.ShouldCollideLocked))
	{
		refresh_body_filter(sys, bodyFilter);
		mask |= GPX_RAYMASK_REQUIRE_BLOCKS_LASERS;
	}
	return cast(sys, &origin, maxDistance, mask, result, hitPointOffset);
}

/* ---- player character ----------------------------------------------------------------------------------------------------- */

void JPH_CharacterVirtualSettings_Init(JPH_CharacterVirtualSettings *s)
{
	if (!s) return;
	if (s->base.up.x == 0.0f && s->base.up.y == 0.0f && s->base.up.z == 0.0f) s->base.up = Vector3_AxisY;
	if (s->base.supportingVolume.normal.x == 0.0f && s->base.supportingVolume.normal.y == 0.0f && s->base.supportingVolume.normal.z == 0.0f)
	{
		s->base.supportingVolume.normal = Vector3_AxisY;
		s->base.supportingVolume.distance = -1.0e10f;
	}
	if (s->base.maxSlopeAngle == 0.0f) s->base.maxSlopeAngle = 50.0f * 3.14159265358979323846f / 180.0f;
	if (s->mass == 0.0f) s->mass = 70.0f;
	if (s->maxStrength == 0.0f) s->maxStrength = 100.0f;
	if (s->predictiveContactDistance == 0.0f) s->predictiveContactDistance = 0.1f;
	if (s->maxCollisionIterations == 0) s->maxCollisionIterations = 5;
	if (s->maxConstraintIterations == 0) s->maxConstraintIterations = 15;
	if (s->minTimeRemaining == 0.0f) s->minTimeRemaining = 1.0e-4f;
	if (s->collisionTolerance == 0.0f) s->collisionTolerance = 1.0e-3f;
	if (s->characterPadding == 0.0f) s->characterPadding = 0.02f;
	if (s->maxNumHits == 0) s->maxNumHits = 256;
	if (s->hitReductionCosMaxAngle == 0.0f) s->hitReductionCosMaxAngle = 0.999f;
	if (s->penetrationRecoverySpeed == 0.0f) s->penetrationRecoverySpeed = 1.0f;
}

JPH_CharacterVirtual *JPH_CharacterVirtual_Create(const JPH_CharacterVirtualSettings *s, const JPH_RVec3 *position, const JPH_Quat *rotation,
												  uint64_t userData, JPH_PhysicsSystem *sys)
{
	if (!s || !position || !sys || !s->base.shape || !sys->build_world()) return nullptr;
	if (sys->character)
	{
		complain("one character per physics system");
		return nullptr;
	}
	if (s->base.shape->kind != SK_CAPSULE)
	{
		complain("the character shape must be a capsule (PlayerPhysics.c:176)");
		return nullptr;
	}
	gpx_character_desc d;
	memset(&d, 0, sizeof(d));
	d.half_height = s->base.shape->half.y;
	d.radius = s->base.shape->half.x;
	d.max_slope_deg = s->base.maxSlopeAngle * 180.0f / 3.14159265358979323846f;
	d.mass = s->mass;
	d.position[0] = position->x; d.position[1] = position->y; d.position[2] = position->z;
	if (gpx_character_create(sys->w, 0, &d) != GPX_OK)
	{
		fprintf(stderr, "joltc_gpx: character create failed: %s\n", gpx_last_error());
		return nullptr;
	}
	if (gpx_events_enable(sys->w, 1) != GPX_OK) complain("contact events unavailable: listener callbacks will not fire");
	auto *ch = new JPH_CharacterVirtual;
	ch->sys = sys;
	ch->user_data = userData;
	if (rotation) ch->rotation = *rotation;
	sys->character = ch;
	return ch;
}
void JPH_CharacterVirtual_Destroy(JPH_CharacterVirtual *ch)
{
	if (!ch) return;
	if (ch->sys)
	{
		gpx_character_destroy(ch->sys->w, 0);
		ch->sys->character = nullptr;
	}
	delete ch;
}
void JPH_CharacterVirtual_SetUserData(JPH_CharacterVirtual *ch, uint64_t userData) { if (ch) ch->user_data = userData; }
uint64_t JPH_CharacterVirtual_GetUserData(const JPH_CharacterVirtual *ch) { return ch ? ch->user_data : 0; }
void JPH_CharacterVirtual_SetListener(JPH_CharacterVirtual *ch, JPH_CharacterContactListener *listener) { if (ch) ch->listener = listener; }
void JPH_CharacterVirtual_SetPosition(JPH_CharacterVirtual *ch, const JPH_RVec3 *p)
{
	if (!ch || !ch->sys || !p) return;
	gpx_character_set_position(ch->sys->w, 0, &p->x);
	ch->stale = true;
}
void JPH_CharacterVirtual_SetRotation(JPH_CharacterVirtual *ch, const JPH_Quat *q) { if (ch && q) ch->rotation = *q; }

static const gpx_character_state *character_state(const JPH_CharacterVirtual *cch)
{
	auto *ch = const_cast<JPH_CharacterVirtual *>(cch);
	if (!ch || !ch->sys) return nullptr;
	if (ch->stale)
	{
		if (gpx_character_get(ch->sys->w, 0, &ch->state) != GPX_OK) return nullptr;
		ch->stale = false;
	}
	return &ch->state;
}
void JPH_CharacterVirtual_GetPosition(const JPH_CharacterVirtual *ch, JPH_RVec3 *out)
{
	const gpx_character_state *s = character_state(ch);
	if (s && out) *out = {s->position[0], s->position[1], s->position[2]};
}
void JPH_CharacterVirtual_GetLinearVelocity(const JPH_CharacterVirtual *ch, Vector3 *out)
{
	const gpx_character_state *s = character_state(ch);
	if (s && out) *out = {s->linear_velocity[0], s->linear_velocity[1], s->linear_velocity[2]};
}
void JPH_CharacterVirtual_SetLinearVelocity(JPH_CharacterVirtual *ch, const Vector3 *v)
{
	if (!ch || !ch->sys || !v) return;
	gpx_character_set_linear_velocity(ch->sys->w, 0, &v->x);
	ch->stale = true;
}
JPH_GroundState JPH_CharacterBase_GetGroundState(const JPH_CharacterBase *ch)
{
	const gpx_character_state *s = character_state(ch);
	return s ? (JPH_GroundState)s->ground_state : JPH_GroundState_InAir;
}
void JPH_CharacterVirtual_ExtendedUpdate(JPH_CharacterVirtual *ch, float dt, const JPH_ExtendedUpdateSettings *es, JPH_ObjectLayer,
										 const JPH_PhysicsSystem *, const JPH_BodyFilter *, const JPH_ShapeFilter *)
{
	if (!ch || !ch->sys) return;
	// the step vectors are vertical in the engine (PlayerPhysics.c:440-441): down is -y, up is +y
	gpx_character_update_settings cfg;
	memset(&cfg, 0, sizeof(cfg));
	if (es)
	{
		cfg.stick_to_floor_step_down = es->stickToFloorStepDown.y < 0.0f ? -es->stickToFloorStepDown.y : 0.0f;
		cfg.walk_stairs_step_up = es->walkStairsStepUp.y > 0.0f ? es->walkStairsStepUp.y : 0.0f;
		cfg.walk_stairs_min_step_forward = es->walkStairsMinStepForward;
		cfg.walk_stairs_step_forward_test = es->walkStairsStepForwardTest;
		cfg.walk_stairs_cos_angle_forward_contact = es->walkStairsCosAngleForwardContact;
	}
	if (gpx_character_update_ex(ch->sys->w, dt, &cfg) != GPX_OK) fprintf(stderr, "joltc_gpx: character update failed: %s\n", gpx_last_error());
	ch->stale = true;
	character_callbacks(ch);
}
JPH_CharacterContactListener *JPH_CharacterContactListener_Create(const JPH_CharacterContactListener_Impl *impl)
{
	auto *l = new JPH_CharacterContactListener;
	memset(&l->impl, 0, sizeof(l->impl));
	if (impl) l->impl = *impl;
	return l;
}
void JPH_CharacterContactListener_Destroy(JPH_CharacterContactListener *l) { delete l; }

/* ---- debug drawing: accepted, draws nothing -------------------------------------------------------------------------- */

JPH_DebugRenderer *JPH_DebugRenderer_Create(void *) { return new JPH_DebugRenderer{0}; }
void JPH_DebugRenderer_Destroy(JPH_DebugRenderer *r) { delete r; }
void JPH_DebugRenderer_SetImpl(const JPH_DebugRenderer_Impl *) {}
JPH_BodyDrawFilter *JPH_BodyDrawFilter_Create(void *) { return new JPH_BodyDrawFilter{0}; }
void JPH_BodyDrawFilter_Destroy(JPH_BodyDrawFilter *f) { delete f; }
void JPH_BodyDrawFilter_SetImpl(const JPH_BodyDrawFilter_Impl *) {}
void JPH_PhysicsSystem_DrawBodies(const JPH_PhysicsSystem *, const JPH_DrawSettings *, JPH_DebugRenderer *, const JPH_BodyDrawFilter *) {}

}  // extern "C"
