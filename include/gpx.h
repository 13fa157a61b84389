/*
 * gpx.h — C ABI of the B200 physics-tick / ray-query backend (libgpx.so).
 *
 * This is the boundary the engine's physics wrapper binds instead of joltc.  Plain C: opaque handle,
 * POD structs, pointers + sizes, int error codes; no C++ or torch types.  Every entry point names the
 * reference interface it replaces (paths relative to the reference tree).
 *
 * One `gpx_world` is an ENSEMBLE of `worlds` independent world instances that share one static map
 * (collision triangles + LBVH).  The engine uses worlds == 1; parameter sweeps use thousands.
 * There is no CPU path: every call that computes launches sm_100a kernels and fails with
 * GPX_ERR_CUDA if the device is unusable.
 */
#ifndef GPX_H
#define GPX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPX_ABI_VERSION 1

/* ---- error codes.  0..4 mirror JPH_PhysicsUpdateError as consumed at engine/src/physics/MapPhysics.c:105-113
 *      (non-zero is fatal in the engine). */
enum gpx_error
{
	GPX_OK = 0,
	GPX_ERR_MANIFOLD_CACHE_FULL = 1,
	GPX_ERR_BODY_PAIR_CACHE_FULL = 2,
	GPX_ERR_CONTACT_CONSTRAINTS_FULL = 4, /* more manifolds than max_manifolds_per_world — or, in a world of more than 64
	                                       * bodies, more than 32 manifolds on one dynamic body; the tick goes on without
	                                       * the contacts that did not fit */
	GPX_ERR_INVALID_ARG = 16,
	GPX_ERR_CAPACITY = 17,
	GPX_ERR_CUDA = 32,
};

/* ---- enums with the reference's values */
enum gpx_motion_type /* JPH_MotionType */
{
	GPX_MOTION_STATIC = 0,
	GPX_MOTION_KINEMATIC = 1,
	GPX_MOTION_DYNAMIC = 2,
};

enum gpx_object_layer /* engine/include/engine/physics/Physics.h:36-42 */
{
	GPX_LAYER_STATIC = 0,
	GPX_LAYER_DYNAMIC = 1,
	GPX_LAYER_PLAYER = 2,
	GPX_LAYER_SENSOR = 3,
};

enum gpx_shape_type /* shape vocabulary of SURVEY §8 row a12 */
{
	GPX_SHAPE_EMPTY = 0,  /* JPH_EmptyShapeSettings_Create  (Actor.c:153, Laser.c:114) */
	GPX_SHAPE_BOX = 1,    /* JPH_BoxShape_Create            (ModelLoader.c:152, Trigger.c:37, Coin.c:43) */
	GPX_SHAPE_SPHERE = 2, /* hull recognised as a sphere    (orb.gmdl, ModelLoader.c:330) */
};

enum gpx_allowed_dofs /* JPH_AllowedDOFs (TestActor.c:42-46) */
{
	GPX_DOF_TX = 1, GPX_DOF_TY = 2, GPX_DOF_TZ = 4,
	GPX_DOF_RX = 8, GPX_DOF_RY = 16, GPX_DOF_RZ = 32,
	GPX_DOF_ALL = 63,
};

/* Ray layer masks: bit i = object layer i accepted.  Replaces the BroadPhaseLayerFilter/ObjectLayerFilter callback
 * pairs, which in the reference are pure functions of the layer (PlayerPhysics.c:55-77, Laser.c:40-72). */
#define GPX_RAYMASK_STATIC (1u << GPX_LAYER_STATIC)
#define GPX_RAYMASK_STATIC_DYNAMIC ((1u << GPX_LAYER_STATIC) | (1u << GPX_LAYER_DYNAMIC))
/* Extra ray flag: only hit bodies whose `ray_flags` has GPX_BODY_BLOCKS_LASERS (the laser BodyFilter, Laser.c:74-85). */
#define GPX_RAYMASK_REQUIRE_BLOCKS_LASERS (1u << 8)

#define GPX_BODY_BLOCKS_LASERS 1u /* "no actor, or actor->flags has ACTOR_FLAG_CAN_BLOCK_LASERS" evaluated at create time */

#define GPX_INVALID_BODY 0xFFFFFFFFu /* JPH_BodyId_InvalidBodyID */
#define GPX_INVALID_FACE 0xFFFFFFFFu

typedef struct gpx_world gpx_world;

/* gpx_world_config.flags.  GPX_WORLD_WIDE runs the wide-world kernels (global sort-and-sweep, thread per pair, islands)
 * for an ENSEMBLE of worlds too — what worlds of dozens of bodies in contact want (a 64-box pile has ~150 manifolds, far
 * more than the lanes the fused ensemble kernel gives a world).  Implied by max_bodies_per_world > 64.  worlds *
 * max_bodies_per_world <= 2^20.  Contact events are available for worlds == 1 only in this mode. */
#define GPX_WORLD_WIDE 1u

/* JPH_PhysicsSystemSettings (engine/src/physics/Physics.c:89-100) + the Jolt defaults the tick uses */
typedef struct gpx_world_config
{
	uint32_t worlds;                  /* number of independent world instances sharing the static map */
	uint32_t max_bodies_per_world;    /* body slots per world: <= 64 for the fused ensemble kernel; more selects GPX_WORLD_WIDE */
	uint32_t max_manifolds_per_world; /* contact manifolds per world and sub-step; 0 = default */
	uint32_t max_static_triangles;    /* capacity of the shared static triangle soup */
	float gravity[3];                 /* JPH_PhysicsSystem_SetGravity (Physics.c:99) */
	int32_t device;                   /* CUDA device ordinal */
	uint32_t velocity_steps;          /* 0 = 10 */
	uint32_t position_steps;          /* 0 = 2  */
	uint32_t flags;                   /* GPX_WORLD_* */
} gpx_world_config;

/* JPH_BodyCreationSettings as the reference fills it (Create2_GAME + setters) */
typedef struct gpx_body_desc
{
	uint32_t shape;          /* gpx_shape_type */
	float half_extents[3];   /* box half extents; sphere: x = radius */
	float convex_radius;     /* carried for parity with JPH_BoxShape_Create(.., r); contacts use the full extents */
	float position[3];
	float rotation[4];       /* x y z w */
	float linear_velocity[3];
	float angular_velocity[3];
	uint32_t motion_type;    /* gpx_motion_type */
	uint32_t layer;          /* gpx_object_layer */
	float mass;              /* > 0: SetMassPropertiesOverride + CalculateInertia (Physbox.c:27-32); 0: density 1000 */
	float friction;          /* Jolt default 0.2; map geometry 4.25 (MapLoader.c:263) */
	float restitution;       /* Jolt default 0 */
	float linear_damping;    /* Jolt default 0.05 */
	float angular_damping;   /* Jolt default 0.05 */
	float gravity_factor;    /* Jolt default 1 */
	uint32_t is_sensor;      /* JPH_BodyCreationSettings_SetIsSensor (Trigger.c:44) */
	uint32_t allowed_dofs;   /* gpx_allowed_dofs; 0 = all */
	uint32_t allow_sleeping; /* Jolt default 1; 0 keeps the body awake for ever */
	uint32_t ray_flags;      /* GPX_BODY_* */
	uint64_t user_data;      /* Actor* */
} gpx_body_desc;

/* Transform (joltc/Math/Transform.h as used at PlayerPhysics.c:305, Actor.c:32) */
typedef struct gpx_transform
{
	float position[3];
	float rotation[4];
} gpx_transform;

typedef struct gpx_ray /* 32 B */
{
	float origin[3];
	float tmax;      /* maxDistance: 10 (PlayerPhysics.c:24,305) or 50 (Laser.c:110) */
	float dir[3];    /* unit direction */
	uint32_t mask;   /* GPX_RAYMASK_*; high 16 bits = world index */
} gpx_ray;

typedef struct gpx_hit /* 16 B — JPH_RayCastResult {bodyID, fraction, subShapeID2} */
{
	float fraction;  /* t / tmax in [0,1]; 2.0f on miss */
	uint32_t body;   /* body id, GPX_INVALID_BODY on miss */
	uint32_t face;   /* canonical face id: index of the triangle in upload order; box face 0..5; sphere 0 */
	uint32_t world;  /* world index the ray ran in */
} gpx_hit;

typedef struct gpx_world_stats /* 32 B, gathered across GPUs at the end of a run */
{
	float kinetic_energy;
	float max_speed;
	uint32_t awake_bodies;
	uint32_t manifolds;
	uint64_t position_checksum;
	uint32_t ticks;
	uint32_t error;
} gpx_world_stats;

/* ---- lifecycle ------------------------------------------------------------------------------------------- */

/* JPH_Init (Physics.c:75).  Returns GPX_ABI_VERSION on success, negative gpx_error otherwise. */
int gpx_init(int device);
/* JPH_Shutdown (Physics.c:86) */
void gpx_shutdown(void);
/* Last CUDA/driver error string for this thread ("" if none). */
const char *gpx_last_error(void);

/* JPH_PhysicsSystem_Create + SetGravity (Physics.c:89-100) */
gpx_world *gpx_world_create(const gpx_world_config *cfg);
/* JPH_PhysicsSystem_Destroy (Physics.c:105) */
void gpx_world_destroy(gpx_world *w);

/* ---- static map geometry ------------------------------------------------------------------------------------ */

/* One collision mesh = one static body: MeshShape -> StaticCompound -> CreateAndAddBody(Static, LAYER_STATIC)
 * (MapLoader.c:206-271; static model actors StaticModel.c:21-72).  `tris` = ntris*9 floats relative to `xfm`.
 * Shared by every world of the ensemble.  Returns the static body id in *out_body. */
int gpx_static_add_mesh(gpx_world *w, const gpx_transform *xfm, const float *tris, uint64_t ntris, float friction,
						uint64_t user_data, uint32_t *out_body);
/* JPH_BodyInterface_RemoveAndDestroyBody for a static mesh body (Map.c:113, Actor.c:68): its triangles leave the soup at
 * the next commit.  The body id is not reused; canonical face ids of later meshes move down. */
int gpx_static_remove_mesh(gpx_world *w, uint32_t body);
/* JPH_PhysicsSystem_OptimizeBroadPhase (MapLoader.c:273): uploads the soup and (re)builds the LBVH on device. */
int gpx_static_commit(gpx_world *w);
/* Parse a decompressed .gmap collision section straight into the static soup (MapLoader.c:200-273).
 * `body` points at the whole decompressed map; returns number of static bodies added or negative error. */
int gpx_static_load_gmap(gpx_world *w, const uint8_t *body, uint64_t size);
/* The same from the asset container (23-byte header + gzip member; engine/src/assets/AssetReader.c:150-257), in memory
 * or on disk (LoadAsset, AssetReader.c:271-305).  Same return value. */
int gpx_static_load_gmap_container(gpx_world *w, const uint8_t *blob, uint64_t size);
int gpx_static_load_gmap_file(gpx_world *w, const char *path);

/* ---- shapes ------------------------------------------------------------------------------------------------------ */

/* JPH_ConvexHullShape_Create(points, n, radius) as the model loader calls it (engine/src/assets/ModelLoader.c:324-341).
 * The body store represents boxes and spheres; a convex hull is classified into one of them.  `exact` = 1 when the
 * hull IS that primitive up to `tolerance` (cube.gmdl -> its 0.4 m box, orb.gmdl -> radius 0.4 sphere), 0 when the
 * result is only the hull's bounding box (leafy.gmdl) — the same stand-in the engine uses for `collision = 1` models
 * (ModelLoader.c:152).  Host-side; needs no device. */
typedef struct gpx_hull_shape
{
	uint32_t shape;        /* GPX_SHAPE_BOX or GPX_SHAPE_SPHERE */
	float half_extents[3]; /* sphere: x = radius */
	float center[3];       /* offset of the primitive's centre from the hull's origin */
	uint32_t exact;
} gpx_hull_shape;
int gpx_shape_from_hull(const float *points, uint64_t n, float tolerance, gpx_hull_shape *out);

/* The collision section of a decompressed .gmdl (engine/src/assets/ModelLoader.c:54-211: materials, skins and LODs are
 * skipped, then the bounding box, then either `numHulls x { numPoints, offset, points }` for dynamic models or
 * `numTriangles x 9 f32` for static ones).  Dynamic hulls are classified like gpx_shape_from_hull (with the hull's offset
 * added to `center`); a static model's triangles can be added to a world with gpx_static_add_gmdl.  Host-side; needs no
 * device.  Returns GPX_OK or GPX_ERR_INVALID_ARG for a malformed file. */
#define GPX_MODEL_MAX_HULLS 8
typedef struct gpx_model_collision
{
	uint32_t collision_type;   /* 0 none, 1 static (triangle mesh), 2 dynamic (convex hulls) */
	float bb_origin[3], bb_extents[3]; /* ModelDefinition::boundingBoxOrigin / Extents: the shape `collision = 1` actors use */
	uint32_t n_hulls;          /* hulls in the file (only the first GPX_MODEL_MAX_HULLS are described below) */
	uint64_t n_triangles;      /* static models */
	uint64_t hull_points[GPX_MODEL_MAX_HULLS];
	gpx_hull_shape hull[GPX_MODEL_MAX_HULLS];
	uint32_t exact;            /* 1 when every hull is exactly a box or a sphere */
} gpx_model_collision;
int gpx_model_load_gmdl(const uint8_t *body, uint64_t size, float tolerance, gpx_model_collision *out);
/* the same from the asset container (23-byte header + gzip member, AssetReader.c:150-257) */
int gpx_model_load_gmdl_container(const uint8_t *blob, uint64_t size, float tolerance, gpx_model_collision *out);

/* A static model's triangle mesh (CreateStaticModelShape, ModelLoader.c:345-351: JPH_MeshShapeSettings_Create over the
 * file's triangles) as one static body at `xfm`, e.g. the laser emitters of test.gmap.  Returns the static body index or
 * a negative error; gpx_static_commit afterwards. */
int gpx_static_add_gmdl(gpx_world *w, const gpx_transform *xfm, const uint8_t *body, uint64_t size, float friction,
						uint32_t ray_flags);

/* ---- bodies --------------------------------------------------------------------------------------------------- */

/* JPH_BodyInterface_CreateAndAddBody (17 call sites).  Returns body id or GPX_INVALID_BODY. */
uint32_t gpx_body_create(gpx_world *w, uint32_t world, const gpx_body_desc *desc);
/* Same body list instantiated in every world; per-world initial velocities optional
 * (`linvel`/`angvel` = worlds*count*3 floats or NULL).  `out_ids` (count) receives the ids, equal in all worlds. */
int gpx_body_create_all(gpx_world *w, const gpx_body_desc *descs, uint32_t count, const float *linvel,
						const float *angvel, uint32_t *out_ids);
/* JPH_BodyInterface_RemoveAndDestroyBody (Actor.c:68, Map.c:113, Door.c:207) */
int gpx_body_destroy(gpx_world *w, uint32_t world, uint32_t body);

/* Setters, applied at the start of the next step (Door.c:59-96, PlayerPhysics.c:377-384, ActorWall.c:70) */
int gpx_body_set_linear_velocity(gpx_world *w, uint32_t world, uint32_t body, const float v[3]);
int gpx_body_set_linear_and_angular_velocity(gpx_world *w, uint32_t world, uint32_t body, const float v[3],
											 const float av[3]);
int gpx_body_set_position(gpx_world *w, uint32_t world, uint32_t body, const float p[3], int activate);
int gpx_body_set_rotation(gpx_world *w, uint32_t world, uint32_t body, const float q[4], int activate);

/* JPH_BodyInterface_ActivateBody: wakes a sleeping body (also done by a non-zero velocity and by activate != 0 above). */
int gpx_body_wake(gpx_world *w, uint32_t world, uint32_t body);
/* Sleeping (the sleep test of JPH_PhysicsSystem_Update): bodies created with allow_sleeping whose three
 * test points stay within 15 mm (0.03 m/s x 0.5 s) for 0.5 s become candidates; an island of candidates goes to sleep
 * (velocities zeroed, static for the tick) until an active body touches it or the host wakes it.  A world in which
 * nothing is active costs no tick (ensemble worlds; a wide world still runs its broadphase).  Evaluated once per tick.
 * `out` receives worlds * max_bodies bytes: 1 = asleep. */
int gpx_read_sleeping(gpx_world *w, uint8_t *out, uint64_t capacity);
/* Re-evaluates the laser BodyFilter for one body (Laser.c:74-85 reads actor->flags, which may change after create). */
int gpx_body_set_ray_flags(gpx_world *w, uint32_t world, uint32_t body, uint32_t ray_flags);

/* Getters served from the host mirror refreshed by gpx_sync_transforms (rows a9: GetPosition/Rotation/
 * PositionAndRotation/WorldTransform/UserData) */
int gpx_body_get_transform(const gpx_world *w, uint32_t world, uint32_t body, gpx_transform *out);
int gpx_body_get_world_matrix(const gpx_world *w, uint32_t world, uint32_t body, float out_mat4[16]);
int gpx_body_get_velocity(const gpx_world *w, uint32_t world, uint32_t body, float v[3], float av[3]);
uint64_t gpx_body_get_user_data(const gpx_world *w, uint32_t world, uint32_t body);
int gpx_body_is_active(const gpx_world *w, uint32_t world, uint32_t body);

/* ---- the tick ---------------------------------------------------------------------------------------------------- */

/* JPH_PhysicsSystem_Update(system, dt, collisionSteps, jobSystem) (MapPhysics.c:105-108).
 * Asynchronous on the world's stream; returns the error of the PREVIOUS completed step merged with launch errors.
 * gpx_sync_* waits. */
int gpx_step(gpx_world *w, float dt, int collision_steps);
/* Wait for the stream; D2H the transform mirror (pinned) so getters are wait-free.  Returns the tick's error code. */
int gpx_sync_transforms(gpx_world *w);
/* Bulk readback: worlds*max_bodies transforms (28 B each) / velocities (24 B each) into caller memory. */
int gpx_read_transforms(gpx_world *w, gpx_transform *out, uint64_t capacity);
int gpx_read_velocities(gpx_world *w, float *out_lin_ang6, uint64_t capacity);
/* Per-world stats computed on device, `out` has `worlds` entries (host). */
int gpx_read_stats(gpx_world *w, gpx_world_stats *out);

/* ---- the step loop ---------------------------------------------------------------------------------------------------- */

/* PhysicsThreadMain and its controls (engine/src/subsystem/threads/PhysicsThread.c:59-159, PhysicsThread.h:10-36),
 * without SDL: one thread calls `function(state, delta)` at 60 Hz while holding the tick mutex.  `delta` is the wall
 * time of the previous tick, idle time included, in units of 1/60 s, clamped to 6 (the 10 ticks/s floor of
 * engine/include/engine/physics/Physics.h:12-22); MapFixedUpdate turns it into dt = delta / 60 (MapPhysics.c:72).
 * Host code only.  One loop per process, as in the engine. */
typedef void (*gpx_fixed_update_fn)(void *state, double delta);               /* GameStateFixedUpdateFunction */
typedef void (*gpx_input_event_fn)(void *state, const void *event, uint64_t size);
/* PhysicsThreadInit: starts the thread; it idles at 60 Hz (counting frames) until a function is set. */
int gpx_thread_init(void *state);
/* PhysicsThreadSetFunction: waits for the running iteration to pick up its function, resets the frame counter. */
void gpx_thread_set_function(gpx_fixed_update_fn function);
/* PhysicsThreadQueueInputEvent: the bytes are copied and handed to the input handler at the start of the next tick. */
void gpx_thread_queue_input_event(const void *event, uint64_t size);
void gpx_thread_set_input_handler(gpx_input_event_fn handler);
/* PhysicsThreadTerminate: posts quit, joins. */
void gpx_thread_terminate(void);
/* PhysicsThreadLockTickMutex / UnlockTickMutex: exclude the fixed update (GlobalState.c:179-192 changes maps under it). */
void gpx_thread_lock_tick_mutex(void);
void gpx_thread_unlock_tick_mutex(void);
/* Not in the engine: pinned != 0 makes every delta exactly 1 and drops the sleep — headless, reproducible runs. */
void gpx_thread_set_pinned_delta(int pinned);
uint64_t gpx_thread_frame(void);        /* GlobalState.physicsFrame */
uint64_t gpx_thread_last_tick_ns(void); /* what TickGraphUpdate is fed (PhysicsThread.c:109) */

/* ---- player character -------------------------------------------------------------------------------------------- */

/* JPH_CharacterVirtual as the engine uses it (engine/src/physics/PlayerPhysics.c:173-194): a capsule that is not a
 * body of the world; one per world instance.  Its contacts are reported by gpx_poll_events with the pseudo body id
 * GPX_CHARACTER_BODY (the CharacterContactListener callbacks, PlayerPhysics.c:89-152). */
#define GPX_CHARACTER_BODY 0x3FFFFFu
/* ids at or above this name the static collision meshes, in the order they were added */
#define GPX_STATIC_BODY_BASE 0x400000u
enum gpx_ground_state /* JPH_GroundState */
{
	GPX_GROUND_ON_GROUND = 0,
	GPX_GROUND_ON_STEEP_GROUND = 1,
	GPX_GROUND_NOT_SUPPORTED = 2,
	GPX_GROUND_IN_AIR = 3,
};
typedef struct gpx_character_desc /* JPH_CharacterVirtualSettings + JPH_CapsuleShape_Create(halfHeight, radius) */
{
	float half_height;   /* 0.2  (PlayerPhysics.c:176) */
	float radius;        /* 0.25 */
	float max_slope_deg; /* MAX_WALKABLE_SLOPE = 50 */
	float mass;          /* 10; carried.  The push on dynamic bodies is limited by the character's strength (Jolt default 100 N) */
	float position[3];
} gpx_character_desc;
typedef struct gpx_character_state
{
	float position[3];        /* JPH_CharacterVirtual_GetPosition (MapPhysics.c:77) */
	float linear_velocity[3]; /* JPH_CharacterVirtual_GetLinearVelocity (PlayerPhysics.c:287) */
	float ground_normal[3];
	uint32_t ground_state;    /* JPH_CharacterBase_GetGroundState (PlayerPhysics.c:284) */
	uint32_t ground_body;     /* body or static mesh stood on, GPX_INVALID_BODY in the air */
} gpx_character_state;
/* JPH_CharacterVirtual_Create / _Destroy */
int gpx_character_create(gpx_world *w, uint32_t world, const gpx_character_desc *desc);
int gpx_character_destroy(gpx_world *w, uint32_t world);
/* JPH_CharacterVirtual_SetLinearVelocity (PlayerPhysics.c:294) / _SetPosition (PlayerPhysics.c:198) */
int gpx_character_set_linear_velocity(gpx_world *w, uint32_t world, const float v[3]);
int gpx_character_set_position(gpx_world *w, uint32_t world, const float p[3]);
/* JPH_CharacterVirtual_ExtendedUpdate (PlayerPhysics.c:447) for the character of every world: move by velocity * dt,
 * collide and slide against the map and the solid bodies (dynamic ones are pushed and woken), ground state, stick to the
 * floor.  Call before gpx_step, as
 * MapFixedUpdate does (MapPhysics.c:74 then :105).  Asynchronous on the world's stream. */
int gpx_character_update(gpx_world *w, float dt);
/* The same with JPH_ExtendedUpdateSettings as the engine fills it (PlayerPhysics.c:439-446): after the move, a character
 * that stood on walkable ground and is now in the air without moving up is set down on a floor found within
 * stick_to_floor_step_down; a character on the ground whose horizontal move was cut short by something too steep that
 * faces the motion (within the angle of walk_stairs_cos_angle_forward_contact) tries the rest of the move lifted by
 * walk_stairs_step_up — at least walk_stairs_min_step_forward — and stands on walkable floor found within the step height
 * below; if it lands on the edge of the step, walkable floor walk_stairs_step_forward_test further on decides.  All zero
 * = gpx_character_update. */
typedef struct gpx_character_update_settings
{
	float stick_to_floor_step_down;              /* 0.25 */
	float walk_stairs_step_up;                   /* 0.25 */
	float walk_stairs_min_step_forward;          /* 0.02 */
	float walk_stairs_step_forward_test;         /* 0.15 */
	float walk_stairs_cos_angle_forward_contact; /* cos 75 deg */
} gpx_character_update_settings;
int gpx_character_update_ex(gpx_world *w, float dt, const gpx_character_update_settings *settings);
/* Waits for the stream and reads the character back. */
int gpx_character_get(gpx_world *w, uint32_t world, gpx_character_state *out);
/* What the character touches after the last gpx_character_update: body ids first (ascending), then static meshes
 * (GPX_STATIC_BODY_BASE + k, ascending), sensors included — the order in which the tick's contact events list them.
 * Waits for the stream.  This is what lets a host deliver the CharacterContactListener callbacks of
 * PlayerPhysics.c:89-152 from inside its ExtendedUpdate, as Jolt does, instead of after the next gpx_step.  *count gets
 * the number of contacts even when it exceeds `capacity`. */
int gpx_character_contacts(gpx_world *w, uint32_t world, uint32_t *others, uint32_t capacity, uint32_t *count);

/* ---- contact events ------------------------------------------------------------------------------------------------ */

/* The listener the engine registers (JPH_CharacterContactListener: OnContactAdded / Persisted / Removed ->
 * ActorDefinition::OnPlayerContact*, engine/src/physics/PlayerPhysics.c:89-152) becomes a polled list: after a tick,
 * every pair of bodies that touches (solver contacts and sensor overlaps alike; body_b >= 0x400000 names a static
 * map mesh) is reported as added, persisted or removed relative to the previous tick.  Order per world: added and
 * persisted pairs sorted by (body_a, body_b), then removed pairs sorted the same way. */
enum gpx_event_kind { GPX_EVENT_ADDED = 1, GPX_EVENT_PERSISTED = 2, GPX_EVENT_REMOVED = 3 };
typedef struct gpx_contact_event
{
	uint32_t world;
	uint32_t body_a; /* lower body id */
	uint32_t body_b;
	uint32_t kind;   /* gpx_event_kind */
} gpx_contact_event;
/* Events cost one pass per tick on the device and are off by default.  (A wide world sorts its touching pairs once per
 * tick for them.) */
int gpx_events_enable(gpx_world *w, int enable);
/* Waits for the stream and returns the events of the last completed tick; *count is the number available. */
int gpx_poll_events(gpx_world *w, gpx_contact_event *out, uint64_t capacity, uint64_t *count);

/* ---- ray queries -------------------------------------------------------------------------------------------------- */

/* Batched closest-hit rays (JPH_NarrowPhaseQuery_CastRay_GAME / CastRay2_GAME; PlayerPhysics.c:305, Laser.c:142).
 * Host buffers; copies are part of the call. */
int gpx_raycast_batch(gpx_world *w, const gpx_ray *rays, uint64_t n, gpx_hit *hits);
/* Same, but only enqueued on the world's stream: `rays` and `hits` must be page-locked (gpx_host_alloc) and stay
 * untouched until the next gpx_sync_transforms / gpx_device_sync, which is when `hits` becomes valid.  Lets a tick's
 * laser batch, the step and the transform readback share ONE synchronisation. */
int gpx_raycast_batch_async(gpx_world *w, const gpx_ray *rays, uint64_t n, gpx_hit *hits);
/* Same with device-resident buffers (cudaMalloc'd by the caller or by gpx_device_alloc), async on the world's stream. */
int gpx_raycast_batch_device(gpx_world *w, const void *d_rays, uint64_t n, void *d_hits);
/* Engine-style single ray: origin transform, direction = transform's local -Z. */
int gpx_raycast_transform(gpx_world *w, uint32_t world, const gpx_transform *origin, float max_distance, uint32_t mask,
						  gpx_hit *out);

/* ---- shape queries ------------------------------------------------------------------------------------------------------ */

/* Batched sphere casts (shape casts): a sphere of `radius` moved from `origin` along the unit `dir` for up to `tmax`, first
 * contact with the static map and the bodies the layer mask admits (same mask bits, body filter and world index as
 * gpx_ray).  fraction = t / tmax, 0 when the sphere overlaps something at the start, 2.0 on a miss; normal = unit vector
 * from the contact point to the sphere's centre at that moment; face as in gpx_hit (box: the face the normal leans to).
 * Host buffers; copies and the wait are part of the call. */
typedef struct gpx_sphere_cast /* 48 B */
{
	float origin[3];
	float tmax;
	float dir[3];
	uint32_t mask;
	float radius;
	float pad[3];
} gpx_sphere_cast;
typedef struct gpx_cast_hit /* 32 B */
{
	float fraction;
	uint32_t body;
	uint32_t face;
	uint32_t world;
	float normal[3];
	float pad;
} gpx_cast_hit;
int gpx_spherecast_batch(gpx_world *w, const gpx_sphere_cast *casts, uint64_t n, gpx_cast_hit *hits);

/* Batched capsule overlap: the collide-shape query JPH_CharacterVirtual_ExtendedUpdate is built from
 * (engine/src/physics/PlayerPhysics.c:447), for callers that bring their own upright capsules.  Each query reports the
 * deepest penetration against the static map and the solid bodies (layers STATIC and DYNAMIC, no sensors) of its world:
 * depth > 0, the unit normal that pushes the capsule out, and the body (>= 0x400000: a static mesh).  depth == 0 and
 * body == GPX_INVALID_BODY when nothing is touched.  Ties go to the lower triangle index, then the lower body id. */
typedef struct gpx_capsule_query /* 32 B */
{
	float center[3];
	float half_height; /* of the cylinder part; the axis is +y */
	float radius;
	uint32_t world;
	uint32_t reserved[2];
} gpx_capsule_query;
typedef struct gpx_overlap /* 32 B */
{
	float depth;
	float normal[3];
	uint32_t body;
	uint32_t world;
	uint32_t reserved[2];
} gpx_overlap;
int gpx_overlap_capsule_batch(gpx_world *w, const gpx_capsule_query *queries, uint64_t n, gpx_overlap *out);

/* ---- device helpers for harnesses --------------------------------------------------------------------------------- */
void *gpx_device_alloc(uint64_t bytes);
void gpx_device_free(void *p);
/* Page-locked host memory for ray/hit staging buffers (so the copies inside gpx_raycast_batch run at PCIe rate). */
void *gpx_host_alloc(uint64_t bytes);
void gpx_host_free(void *p);
int gpx_memcpy_h2d(void *dst, const void *src, uint64_t bytes);
int gpx_memcpy_d2h(void *dst, const void *src, uint64_t bytes);
int gpx_device_sync(gpx_world *w);
/* cudaEvent timing on the world's stream: begin/end bracket, returns elapsed ms of the last bracket. */
int gpx_timer_begin(gpx_world *w);
float gpx_timer_end(gpx_world *w);
/* Number of kernels this library launched since gpx_init (for bench.py's gpu_launches). */
uint64_t gpx_launch_count(void);
/* Profiling aid: SM cycles spent per tick phase, summed over every world's lane 0 since the last call
 * (enable != 0 arms the counters; out16 may be NULL).  Order: load, forces, static narrowphase, pair narrowphase,
 * warm-start match, colouring, constraint set-up, warm start, velocity iterations, integrate, position iterations,
 * cache write, store. */
int gpx_debug_phase_cycles(gpx_world *w, int enable, uint64_t *out16);
/* Profiling aid for a wide world: counters of the last sub-step — manifold slots used, islands solved inside one warp,
 * colours of the large islands, sticky error word, islands solved by one block each, manifolds of the large islands,
 * largest body extent in x and z as float bits. */
int gpx_debug_wide_counters(gpx_world *w, uint32_t *out8);
/* Static LBVH introspection for tests: node count and triangle count after commit. */
int gpx_static_info(const gpx_world *w, uint32_t *n_tris, uint32_t *n_nodes, uint32_t *n_bodies);

#ifdef __cplusplus
}
#endif
#endif /* GPX_H */
