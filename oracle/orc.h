/*
 * orc.h — CPU oracle for the physics tick and ray queries.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or
 * call this.  The product (libgpx.so) never links it and has no CPU path.
 *
 * PARITY UNPINNED: the reference delegates this path to joltc @226bd1e0afa8552917ffd56bca03ccb5c14ecd27 ->
 * JoltPhysics, which is not vendored in the reference tree and cannot be fetched or built here; the
 * reference has no tests or golden vectors.  This file therefore restates Jolt's PUBLISHED tick structure
 * (sub-steps; gravity+damping; speculative contacts; manifold clipping between supporting faces; 4-point
 * manifold reduction; warm-started sequential impulses, friction before non-penetration; integrate; Baumgarte
 * position pass) with Jolt's documented default constants, anchored on the reference's call sites:
 *   tick order / dt / 2 collision steps ..... engine/src/physics/MapPhysics.c:58-119
 *   world, layers, gravity ................... engine/src/physics/Physics.c:20-100, Physics.h:12-51
 *   static map upload, friction 4.25 ......... engine/src/assets/MapLoader.c:200-273
 *   body parameters .......................... game/src/actor/prop/Physbox.c:19-38, engine/src/actor/Trigger.c:33-50 ...
 *   rays and their filters ................... engine/src/physics/PlayerPhysics.c:55-86,297-315, game/src/actor/prop/Laser.c:40-158
 * It is pinned instead by analytic known answers (tests/test_oracle.py): free fall with damping, resting contact height,
 * Moller-Trumbore and swept-sphere known hits, momentum conservation, Coulomb stopping distances, the incline law,
 * restitution, and the dissipation a
 * contact solver owes its user — a kicked 8-box column loses its kinetic energy monotonically (one-second windows), comes
 * to rest at the stacked heights and, with sleeping allowed, is asleep within five seconds.
 *
 * Two deliberate departures from Jolt's solver, both needed for that last property with 10 velocity iterations (DESIGN.md
 * section 3 has the measurements): friction is solved per MANIFOLD (two tangent rows through the centroid of the contact
 * points plus one twist row about the normal, limited by friction x the manifold's total normal impulse) instead of per
 * contact point, and the non-penetration rows of a manifold are swept forwards in even iterations and backwards in odd
 * ones instead of always in the same order.  Manifolds against a static mesh are cached per (body, mesh, triangle that
 * opened the manifold's slot), the nearest equivalent here of Jolt's sub-shape-pair key.
 *
 * Plain C, single precision, compiled with -ffp-contract=off so that every expression rounds exactly like the
 * CUDA build (-fmad=false).  Algorithms here are deliberately the naive ones (all-pairs broadphase — above 256 bodies
 * its candidates come from a sort-and-sweep that yields the same pairs in the same order —, brute-force triangle loops,
 * sequential constraint order): the GPU path must reproduce their results, not their structure.
 */
#ifndef ORC_H
#define ORC_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_STATIC_BODY_BASE 0x400000u
#define ORC_INVALID 0xFFFFFFFFu

enum { ORC_SHAPE_EMPTY = 0, ORC_SHAPE_BOX = 1, ORC_SHAPE_SPHERE = 2 };
enum { ORC_MOTION_STATIC = 0, ORC_MOTION_KINEMATIC = 1, ORC_MOTION_DYNAMIC = 2 };

typedef struct orc_ray { float origin[3]; float tmax; float dir[3]; uint32_t mask; } orc_ray;
typedef struct orc_hit { float fraction; uint32_t body; uint32_t face; uint32_t world; } orc_hit;

typedef struct orc_body_desc
{
	uint32_t shape;
	float half_extents[3];
	float convex_radius;
	float position[3];
	float rotation[4];
	float linear_velocity[3];
	float angular_velocity[3];
	uint32_t motion_type;
	uint32_t layer;
	float mass;
	float friction;
	float restitution;
	float linear_damping;
	float angular_damping;
	float gravity_factor;
	uint32_t is_sensor;
	uint32_t allowed_dofs;
	uint32_t allow_sleeping;
	uint32_t ray_flags;
	uint64_t user_data;
} orc_body_desc;

typedef struct orc_world orc_world;

orc_world *orc_world_create(uint32_t max_bodies, uint32_t max_manifolds, const float gravity[3],
							uint32_t velocity_steps, uint32_t position_steps);
void orc_world_destroy(orc_world *w);
/* 0 (default): greedy colouring in canonical order, as the ensemble kernel; 1: hashed-priority rounds, as the wide-world kernels */
void orc_world_set_mode(orc_world *w, int mode);
/* static collision mesh = one static body; tris relative to (pos, rot) */
uint32_t orc_static_add_mesh(orc_world *w, const float pos[3], const float rot[4], const float *tris, uint64_t ntris,
							 float friction);
void orc_static_commit(orc_world *w);
uint32_t orc_body_create(orc_world *w, const orc_body_desc *d);
void orc_body_destroy(orc_world *w, uint32_t id);
void orc_body_set_velocity(orc_world *w, uint32_t id, const float v[3], const float av[3]);
void orc_body_set_position(orc_world *w, uint32_t id, const float p[3]);
void orc_body_wake(orc_world *w, uint32_t id);
uint32_t orc_body_asleep(const orc_world *w, uint32_t id);
void orc_body_set_ray_flags(orc_world *w, uint32_t id, uint32_t ray_flags);
/* one tick = collision_steps sub-steps of dt/collision_steps; returns 0 or an error code (4 = contact constraints full) */
int orc_step(orc_world *w, float dt, int collision_steps);
/* the same tick over host threads (wide mode only; any other world takes orc_step): bit-identical to orc_step */
int orc_step_mt(orc_world *w, float dt, int collision_steps);
/* state access: out = 7 floats pos+quat, 6 floats lin+ang */
void orc_body_get(const orc_world *w, uint32_t id, float *xf7, float *vel6);
uint32_t orc_body_active(const orc_world *w, uint32_t id);
uint32_t orc_manifold_count(const orc_world *w);
/* contact events of the last step: triples (a, b, kind) */
uint32_t orc_events(const orc_world *w, uint32_t *out, uint32_t cap);
/* player character: a capsule moved by discrete collide-and-slide before the tick (see the character section of orc.c);
 * ground: 0 on ground, 1 on steep ground, 3 in air (JPH_GroundState values); its contacts appear in orc_events with
 * the pseudo body id 0x3FFFFF */
void orc_character_create(orc_world *w, const float pos[3], float half_height, float radius, float max_slope_deg);
void orc_character_destroy(orc_world *w);
void orc_character_set_velocity(orc_world *w, const float v[3]);
void orc_character_set_position(orc_world *w, const float p[3]);
void orc_character_update(orc_world *w, float dt);
/* JPH_ExtendedUpdateSettings as the engine fills it (PlayerPhysics.c:439-446); all zero = plain update */
typedef struct
{
	float stick_to_floor_step_down;  /* 0.25 */
	float walk_stairs_step_up;       /* 0.25 */
	float walk_stairs_min_step_forward;            /* 0.02 */
	float walk_stairs_step_forward_test;           /* 0.15 */
	float walk_stairs_cos_angle_forward_contact;   /* cos 75 deg */
} orc_character_settings;
void orc_character_update_ex(orc_world *w, float dt, const orc_character_settings *cfg);
float orc_overlap_capsule(const orc_world *w, const float center[3], float half_height, float radius, float normal[3],
						  uint32_t *body);
void orc_character_get(const orc_world *w, float pos[3], float vel[3], uint32_t *ground, uint32_t *ground_body);
/* ids the character touches after the last orc_character_update: bodies ascending, then static meshes ascending */
uint32_t orc_character_contacts(const orc_world *w, uint32_t *others, uint32_t cap);
/* Sphere casts (shape casts): a sphere of `radius` moved from `origin` along the unit `dir` for up to `tmax`; the first
 * contact with the static triangles and the bodies the layer mask admits.  fraction = t / tmax (0 when the sphere
 * already overlaps something at the start), normal = from the contact point to the sphere's centre at that moment. */
typedef struct orc_sphere_cast { float origin[3]; float tmax; float dir[3]; uint32_t mask; float radius; float pad[3]; } orc_sphere_cast;
typedef struct orc_cast_hit { float fraction; uint32_t body; uint32_t face; uint32_t world; float normal[3]; float pad; } orc_cast_hit;
void orc_spherecast(const orc_world *w, const orc_sphere_cast *casts, uint64_t n, orc_cast_hit *hits);
/* closest-hit rays, brute force over every static triangle and every body */
void orc_raycast(const orc_world *w, const orc_ray *rays, uint64_t n, orc_hit *hits);
uint32_t orc_static_triangles(const orc_world *w, float *out9, uint32_t *out_body, uint32_t cap);
/* threads used by the *_mt helpers below (OpenMP) */
int orc_max_threads(void);
void orc_raycast_mt(const orc_world *w, const orc_ray *rays, uint64_t n, orc_hit *hits);
/* step many independent worlds (array of handles) in parallel over host threads */
int orc_step_many(orc_world **ws, uint32_t n, float dt, int collision_steps, int ticks);

#ifdef __cplusplus
}
#endif
#endif
