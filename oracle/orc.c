/* orc.c — CPU oracle (TEST INFRASTRUCTURE ONLY; header comment in orc.h states scope, sources and "parity unpinned").
 *
 * Structure of one sub-step (restating Jolt's published PhysicsSystem::Update step as the reference invokes it with
 * collisionSteps = 2, engine/src/physics/MapPhysics.c:105-108):
 *   1 apply gravity + damping, clamp velocities            (dynamic bodies)
 *   2 find contacts: all-pairs AABB + brute-force triangle loop, SAT + supporting-face clipping, <=4 points
 *   3 carry impulses from the previous sub-step's manifolds (warm start)
 *   4 colour manifolds greedily in canonical order; solve in (colour, index) order
 *   5 velocity_steps x (manifold friction rows, then the points' non-penetration rows in alternating order)
 *   6 integrate positions / rotations
 *   7 position_steps x Baumgarte position correction
 */
#include "orc.h"
#include "orc_math.h"

#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ---- constants (Jolt PhysicsSettings defaults, SURVEY §8c "upstream, unverified") */
#define SPECULATIVE_DISTANCE 0.02f
#define PENETRATION_SLOP 0.02f
#define BAUMGARTE 0.2f
#define MAX_PENETRATION_DISTANCE 0.2f
#define MAX_LINEAR_VELOCITY 500.0f
#define MAX_ANGULAR_VELOCITY 47.1238898f /* 0.25 * pi * 60 */
#define MIN_VELOCITY_FOR_RESTITUTION 1.0f
#define NORMAL_COS_MAX_DELTA 0.99619470f /* cos 5 deg: manifolds of one body pair merge below this angle */
#define PRESERVE_LAMBDA_MAX_DIST_SQ 1.0e-4f
#define FACE_AXIS_TOL 1.0e-4f
#define EDGE_AXIS_TOL 2.0e-3f
#define RAY_MISS_FRACTION 2.0f
#define MAX_POLY 16
#define MAX_SLOTS 4 /* manifolds per (body, static body) pair */

typedef struct
{
	int alive;
	uint32_t shape;
	v3 he;
	v3 x;
	q4 q;
	v3 v, w;
	uint32_t motion, layer;
	float inv_mass;
	v3 inv_inertia; /* local diagonal */
	float friction, restitution, lin_damp, ang_damp, grav_factor;
	uint32_t sensor, dofs, allow_sleep, ray_flags;
	uint64_t user_data;
	/* per sub-step solver view */
	float im;   /* inverse mass, 0 unless dynamic */
	float M[6]; /* world inverse inertia xx xy xz yy yz zz */
	/* sleeping (see the sleeping section): asleep bodies behave as static until something active touches them */
	int asleep, wake_mark;
	float sleep_t;  /* time the test points have stayed inside their spheres; < 0: spheres not set yet */
	v3 sleep_c[3];
	float sleep_r[3];
} body_t;

/* "dynamic" for everything the tick does: a dynamic body that is awake */
static inline int is_dyn(const body_t *b) { return b->motion == ORC_MOTION_DYNAMIC && !b->asleep; }
/* bodies that find contacts and wake sleepers: awake dynamic ones and kinematic ones that move */
static inline int is_active(const body_t *b)
{
	if (is_dyn(b)) return 1;
	return b->motion == ORC_MOTION_KINEMATIC &&
		   (b->v.x != 0.0f || b->v.y != 0.0f || b->v.z != 0.0f || b->w.x != 0.0f || b->w.y != 0.0f || b->w.z != 0.0f);
}

/* One constraint row: the velocity it measures is Jv = (axis.va + a1.wa) - (axis.vb + a2.wb); an impulse d along it
 * changes va by -lA*d, wa by -I1*d, vb by +lB*d, wb by +I2*d.  em = 1 / (J M^-1 J^T). */
typedef struct
{
	v3 a1, a2; /* r1 x axis, r2 x axis (twist row: the axis itself) */
	v3 I1, I2; /* world inverse inertia times a1 / a2 */
	float em;
} row_t;

/* The rows of one manifold, rebuilt every sub-step from the bodies' poses: a non-penetration row per contact point, and
 * for the manifold as a whole two friction rows through the centroid of the contact points plus one twist row about the
 * normal, limited by the friction coefficient times the sum of the normal impulses (times the patch radius for twist). */
typedef struct
{
	row_t n[4];
	float bias[4];
	row_t t[2];
	row_t w;
	float rp;              /* patch radius: RMS distance of the contact points from their centroid */
	v3 nA, nB, tA[2], tB[2]; /* axis times inverse mass, locked translation axes zeroed */
} rows_t;

typedef struct
{
	uint32_t a, b; /* a: slot body (dynamic side); b: slot body or ORC_STATIC_BODY_BASE + k */
	v3 n;          /* world normal, a -> b */
	float depth;   /* deepest penetration seen when merging */
	int np;
	v3 p1l[4], p2l[4]; /* contact points in the local frames of a and b (static: world) */
	float ln[4];       /* accumulated normal impulse per point */
	float cf[3];       /* accumulated friction impulse of the manifold: tangent 1, tangent 2 (at the centroid), twist about n */
	/* solver scratch */
	float friction, restitution;
	v3 t1, t2;
	rows_t rows;
	int colour;
	uint32_t ord;  /* ordinal among the manifolds of the same (a, b): 0 for body pairs, slot index for static bodies */
	uint32_t tri;  /* static bodies: the triangle that opened this manifold's slot — what keeps the warm start of a box's
	                * floor manifold and of its wall manifold (same body pair, contact points a hair apart) from feeding each other */
	uint32_t prio; /* colouring priority (mode 1) */
} manifold_t;

typedef struct
{
	v3 v0, e1, e2; /* world space: first vertex and the two edges from it */
	v3 vb, vc;     /* the other two vertices as uploaded */
	v3 n;          /* unit normal */
	v3 lo, hi;
	uint32_t body; /* static body index k */
} tri_t;

typedef struct
{
	float friction;
	uint32_t ray_flags;
	uint32_t first, count;
} sbody_t;

struct orc_world
{
	uint32_t free_hint;  /* body slots below it are all alive */
	uint32_t max_bodies, max_manifolds, vel_steps, pos_steps;
	v3 gravity;
	body_t *bodies;
	tri_t *tris;
	uint32_t ntris, cap_tris;
	sbody_t *sbodies;
	uint32_t nsbodies, cap_sbodies;
	manifold_t *man, *prev;
	uint32_t nman, nprev;
	uint32_t *order;
	/* touching pairs for contact events: keys (a << 32 | b), sorted; sensors are kept apart from the solver's manifolds */
	uint64_t *sens, *ev_cur, *ev_prev;
	uint32_t nsens, nev_cur, nev_prev;
	uint32_t *events; /* triples a, b, kind (1 added, 2 persisted, 3 removed) of the last tick */
	uint32_t nevents;
	/* player character (capsule), see the character section */
	int ch_alive;
	v3 ch_x, ch_v, ch_ground_n;
	float ch_hh, ch_r, ch_cos_slope;
	uint32_t ch_ground, ch_ground_body;
	uint64_t ch_keys[64];
	uint32_t ch_nkeys;
	int mode; /* 0: greedy colouring in canonical order (ensembles); 1: hashed-priority rounds (wide worlds) */
};

/* ------------------------------------------------------------------------------------------ world management */

orc_world *orc_world_create(uint32_t max_bodies, uint32_t max_manifolds, const float gravity[3], uint32_t vs,
							uint32_t ps)
{
	orc_world *w = (orc_world *)calloc(1, sizeof(orc_world));
	w->max_bodies = max_bodies;
	w->max_manifolds = max_manifolds ? max_manifolds : max_bodies * 8u;
	w->vel_steps = vs ? vs : 10u;
	w->pos_steps = ps ? ps : 2u;
	w->gravity = V(gravity[0], gravity[1], gravity[2]);
	w->bodies = (body_t *)calloc(max_bodies, sizeof(body_t));
	w->man = (manifold_t *)calloc(w->max_manifolds, sizeof(manifold_t));
	w->prev = (manifold_t *)calloc(w->max_manifolds, sizeof(manifold_t));
	w->order = (uint32_t *)calloc(w->max_manifolds, sizeof(uint32_t));
	w->sens = (uint64_t *)calloc(w->max_manifolds, sizeof(uint64_t));
	w->ev_cur = (uint64_t *)calloc(2 * w->max_manifolds + 64, sizeof(uint64_t));
	w->ev_prev = (uint64_t *)calloc(2 * w->max_manifolds + 64, sizeof(uint64_t));
	w->events = (uint32_t *)calloc(12 * w->max_manifolds + 6 * 64, sizeof(uint32_t));
	return w;
}

void orc_world_set_mode(orc_world *w, int mode) { w->mode = mode; }

void orc_world_destroy(orc_world *w)
{
	if (!w) return;
	free(w->bodies);
	free(w->tris);
	free(w->sbodies);
	free(w->man);
	free(w->prev);
	free(w->order);
	free(w->sens);
	free(w->ev_cur);
	free(w->ev_prev);
	free(w->events);
	free(w);
}

uint32_t orc_static_add_mesh(orc_world *w, const float pos[3], const float rot[4], const float *tris, uint64_t ntris,
							 float friction)
{
	if (w->nsbodies == w->cap_sbodies)
	{
		w->cap_sbodies = w->cap_sbodies ? w->cap_sbodies * 2 : 16;
		w->sbodies = (sbody_t *)realloc(w->sbodies, w->cap_sbodies * sizeof(sbody_t));
	}
	if (w->ntris + ntris > w->cap_tris)
	{
		w->cap_tris = (uint32_t)(w->ntris + ntris) * 2;
		w->tris = (tri_t *)realloc(w->tris, w->cap_tris * sizeof(tri_t));
	}
	const uint32_t k = w->nsbodies++;
	sbody_t *sb = &w->sbodies[k];
	sb->friction = friction;
	sb->ray_flags = 1;
	sb->first = w->ntris;
	sb->count = (uint32_t)ntris;
	const q4 q = {rot[0], rot[1], rot[2], rot[3]};
	const v3 p = V(pos[0], pos[1], pos[2]);
	for (uint64_t i = 0; i < ntris; i++)
	{
		const float *t = tris + i * 9;
		v3 a = vadd(qrot(q, V(t[0], t[1], t[2])), p);
		v3 b = vadd(qrot(q, V(t[3], t[4], t[5])), p);
		v3 c = vadd(qrot(q, V(t[6], t[7], t[8])), p);
		tri_t *T = &w->tris[w->ntris++];
		T->v0 = a;
		T->vb = b;
		T->vc = c;
		T->e1 = vsub(b, a);
		T->e2 = vsub(c, a);
		v3 n = vcross(T->e1, T->e2);
		float l = vlen(n);
		T->n = l > 0.0f ? vscale(n, 1.0f / l) : V(0, 1, 0);
		T->lo = V(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)));
		T->hi = V(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)));
		T->body = k;
	}
	return ORC_STATIC_BODY_BASE + k;
}

void orc_static_commit(orc_world *w) { (void)w; }

uint32_t orc_static_triangles(const orc_world *w, float *out9, uint32_t *out_body, uint32_t cap)
{
	for (uint32_t i = 0; i < w->ntris && i < cap; i++)
	{
		const tri_t *T = &w->tris[i];
		v3 b = T->vb, c = T->vc;
		float *o = out9 + 9 * i;
		o[0] = T->v0.x; o[1] = T->v0.y; o[2] = T->v0.z;
		o[3] = b.x; o[4] = b.y; o[5] = b.z;
		o[6] = c.x; o[7] = c.y; o[8] = c.z;
		if (out_body) out_body[i] = ORC_STATIC_BODY_BASE + T->body;
	}
	return w->ntris;
}

uint32_t orc_body_create(orc_world *w, const orc_body_desc *d)
{
	/* the lowest free slot; every slot below free_hint is known to be taken */
	uint32_t id = ORC_INVALID;
	for (uint32_t i = w->free_hint; i < w->max_bodies; i++)
		if (!w->bodies[i].alive)
		{
			id = i;
			break;
		}
	if (id == ORC_INVALID) return id;
	w->free_hint = id + 1;
	body_t *b = &w->bodies[id];
	memset(b, 0, sizeof(*b));
	b->alive = 1;
	b->shape = d->shape;
	b->he = V(d->half_extents[0], d->half_extents[1], d->half_extents[2]);
	b->x = V(d->position[0], d->position[1], d->position[2]);
	q4 q = {d->rotation[0], d->rotation[1], d->rotation[2], d->rotation[3]};
	b->q = qnormalize(q);
	b->v = V(d->linear_velocity[0], d->linear_velocity[1], d->linear_velocity[2]);
	b->w = V(d->angular_velocity[0], d->angular_velocity[1], d->angular_velocity[2]);
	b->motion = d->motion_type;
	b->layer = d->layer;
	b->friction = d->friction;
	b->restitution = d->restitution;
	b->lin_damp = d->linear_damping;
	b->ang_damp = d->angular_damping;
	b->grav_factor = d->gravity_factor;
	b->sensor = d->is_sensor;
	b->dofs = d->allowed_dofs ? d->allowed_dofs : 63u;
	b->allow_sleep = d->allow_sleeping;
	b->sleep_t = -1.0f;
	b->ray_flags = d->ray_flags;
	b->user_data = d->user_data;
	if (b->motion == ORC_MOTION_DYNAMIC && b->shape != ORC_SHAPE_EMPTY)
	{
		/* mass override + CalculateInertia (Physbox.c:27-32): inertia of the shape at density 1000 scaled to the mass */
		float m, ix, iy, iz;
		if (b->shape == ORC_SHAPE_BOX)
		{
			float sx = 2.0f * b->he.x, sy = 2.0f * b->he.y, sz = 2.0f * b->he.z;
			float vol = (sx * sy) * sz;
			m = d->mass > 0.0f ? d->mass : 1000.0f * vol;
			float k = m / 12.0f;
			ix = k * ((sy * sy) + (sz * sz));
			iy = k * ((sx * sx) + (sz * sz));
			iz = k * ((sx * sx) + (sy * sy));
		}
		else
		{
			float r = b->he.x;
			float vol = (4.18879020f * r) * (r * r);
			m = d->mass > 0.0f ? d->mass : 1000.0f * vol;
			ix = iy = iz = (0.4f * m) * (r * r);
		}
		b->inv_mass = 1.0f / m;
		b->inv_inertia = V(1.0f / ix, 1.0f / iy, 1.0f / iz);
	}
	return id;
}

void orc_body_destroy(orc_world *w, uint32_t id)
{
	if (id >= w->max_bodies) return;
	w->bodies[id].alive = 0;
	if (id < w->free_hint) w->free_hint = id;
}

static void wake_body(body_t *b)
{
	b->asleep = 0;
	b->wake_mark = 0;
	b->sleep_t = -1.0f;
}

/* JPH_BodyInterface_ActivateBody, and what SetPosition(.., JPH_Activation_Activate) implies */
void orc_body_wake(orc_world *w, uint32_t id)
{
	if (id < w->max_bodies) wake_body(&w->bodies[id]);
}

uint32_t orc_body_asleep(const orc_world *w, uint32_t id) { return id < w->max_bodies && w->bodies[id].asleep; }

void orc_body_set_velocity(orc_world *w, uint32_t id, const float v[3], const float av[3])
{
	if (id >= w->max_bodies) return;
	if (v) w->bodies[id].v = V(v[0], v[1], v[2]);
	if (av) w->bodies[id].w = V(av[0], av[1], av[2]);
	/* a non-zero velocity activates the body (BodyInterface::SetLinearVelocity) */
	if ((v && (v[0] != 0.0f || v[1] != 0.0f || v[2] != 0.0f)) || (av && (av[0] != 0.0f || av[1] != 0.0f || av[2] != 0.0f)))
		wake_body(&w->bodies[id]);
}

/* JPH_BodyInterface_SetPosition (Door.c:82-96) and the re-evaluated laser body filter (Laser.c:74-85) */
void orc_body_set_position(orc_world *w, uint32_t id, const float p[3])
{
	if (id >= w->max_bodies || !p) return;
	w->bodies[id].x = V(p[0], p[1], p[2]);
}

void orc_body_set_ray_flags(orc_world *w, uint32_t id, uint32_t ray_flags)
{
	if (id >= w->max_bodies) return;
	w->bodies[id].ray_flags = ray_flags;
}

void orc_body_get(const orc_world *w, uint32_t id, float *xf7, float *vel6)
{
	const body_t *b = &w->bodies[id];
	if (xf7)
	{
		xf7[0] = b->x.x; xf7[1] = b->x.y; xf7[2] = b->x.z;
		xf7[3] = b->q.x; xf7[4] = b->q.y; xf7[5] = b->q.z; xf7[6] = b->q.w;
	}
	if (vel6)
	{
		vel6[0] = b->v.x; vel6[1] = b->v.y; vel6[2] = b->v.z;
		vel6[3] = b->w.x; vel6[4] = b->w.y; vel6[5] = b->w.z;
	}
}

uint32_t orc_body_active(const orc_world *w, uint32_t id) { return id < w->max_bodies && w->bodies[id].alive; }
uint32_t orc_manifold_count(const orc_world *w) { return w->nprev; }
uint32_t orc_events(const orc_world *w, uint32_t *out, uint32_t cap)
{
	uint32_t n = w->nevents < cap ? w->nevents : cap;
	if (out) memcpy(out, w->events, (size_t)n * 3 * sizeof(uint32_t));
	return w->nevents;
}

/* ------------------------------------------------------------------------------------------ rays */

/* Two-sided Moller-Trumbore; accepts t in [0, tmax]. */
static int ray_tri(v3 o, v3 d, float tmax, const tri_t *T, float *tout)
{
	v3 p = vcross(d, T->e2);
	float det = vdot(T->e1, p);
	if (fabsf(det) < 1.0e-12f) return 0;
	float inv = 1.0f / det;
	v3 tv = vsub(o, T->v0);
	float u = vdot(tv, p) * inv;
	if (u < 0.0f || u > 1.0f) return 0;
	v3 q = vcross(tv, T->e1);
	float v = vdot(d, q) * inv;
	if (v < 0.0f || (u + v) > 1.0f) return 0;
	float t = vdot(T->e2, q) * inv;
	if (t < 0.0f || t > tmax) return 0;
	*tout = t;
	return 1;
}

static int ray_box(v3 o, v3 d, float tmax, const body_t *b, float *tout, uint32_t *face)
{
	m33 R = qmat(b->q);
	v3 lo = mtmul(&R, vsub(o, b->x));
	v3 ld = mtmul(&R, d);
	float tn = -3.0e38f, tf = 3.0e38f;
	uint32_t fn = 0;
	for (int k = 0; k < 3; k++)
	{
		float ok = vget(lo, k), dk = vget(ld, k), hk = vget(b->he, k);
		if (dk == 0.0f)
		{
			if (ok < -hk || ok > hk) return 0;
			continue;
		}
		float inv = 1.0f / dk;
		float t1 = (-hk - ok) * inv, t2 = (hk - ok) * inv;
		uint32_t f1 = (uint32_t)(2 * k), f2 = (uint32_t)(2 * k + 1); /* face 2k = -axis, 2k+1 = +axis */
		if (t1 > t2)
		{
			float tt = t1; t1 = t2; t2 = tt;
			f1 = f2;
		}
		if (t1 > tn) { tn = t1; fn = f1; }
		if (t2 < tf) tf = t2;
	}
	if (tn > tf || tf < 0.0f) return 0;
	float t = tn < 0.0f ? 0.0f : tn;
	if (t > tmax) return 0;
	*tout = t;
	*face = fn;
	return 1;
}

static int ray_sphere(v3 o, v3 d, float tmax, const body_t *b, float *tout)
{
	v3 m = vsub(o, b->x);
	float r = b->he.x;
	float bq = vdot(m, d);
	float c = vdot(m, m) - (r * r);
	if (c > 0.0f && bq > 0.0f) return 0;
	float disc = (bq * bq) - c;
	if (disc < 0.0f) return 0;
	float t = -bq - sqrtf(disc);
	if (t < 0.0f) t = 0.0f;
	if (t > tmax) return 0;
	*tout = t;
	return 1;
}

static void raycast_one(const orc_world *w, const orc_ray *r, orc_hit *h)
{
	const v3 o = V(r->origin[0], r->origin[1], r->origin[2]);
	const v3 d = V(r->dir[0], r->dir[1], r->dir[2]);
	const uint32_t layers = r->mask & 0xFu;
	const int need_flag = (r->mask & (1u << 8)) != 0;
	float best = 3.0e38f;
	uint32_t bbody = ORC_INVALID, bface = ORC_INVALID;
	if (layers & 1u)
		for (uint32_t i = 0; i < w->ntris; i++)
		{
			const tri_t *T = &w->tris[i];
			if (need_flag && !(w->sbodies[T->body].ray_flags & 1u)) continue;
			float t;
			if (ray_tri(o, d, r->tmax, T, &t) && t < best)
			{
				best = t;
				bbody = ORC_STATIC_BODY_BASE + T->body;
				bface = i;
			}
		}
	for (uint32_t i = 0; i < w->max_bodies; i++)
	{
		const body_t *b = &w->bodies[i];
		if (!b->alive || b->shape == ORC_SHAPE_EMPTY) continue;
		if (!((layers >> b->layer) & 1u)) continue;
		if (need_flag && !(b->ray_flags & 1u)) continue;
		float t;
		uint32_t f = 0;
		int hit = b->shape == ORC_SHAPE_BOX ? ray_box(o, d, r->tmax, b, &t, &f) : ray_sphere(o, d, r->tmax, b, &t);
		if (hit && t < best)
		{
			best = t;
			bbody = i;
			bface = f;
		}
	}
	h->world = r->mask >> 16;
	if (bbody == ORC_INVALID)
	{
		h->fraction = RAY_MISS_FRACTION;
		h->body = ORC_INVALID;
		h->face = ORC_INVALID;
	}
	else
	{
		h->fraction = best / r->tmax;
		h->body = bbody;
		h->face = bface;
	}
}

void orc_raycast(const orc_world *w, const orc_ray *rays, uint64_t n, orc_hit *hits)
{
	for (uint64_t i = 0; i < n; i++) raycast_one(w, &rays[i], &hits[i]);
}

/* ---- host-thread fan-out for the timed CPU baseline (plain pthreads, static partition) */
typedef struct
{
	void (*fn)(void *ctx, int64_t lo, int64_t hi);
	void *ctx;
	int64_t lo, hi;
} job_t;

/* ------------------------------------------------------------------------------------------ sphere casts
 * First contact of a sphere swept along a ray with a triangle = the earliest of: the sphere's lowest point reaching the
 * triangle's plane inside the triangle; the centre's ray entering a cylinder of the sphere's radius around an edge; the
 * centre's ray entering a sphere of that radius around a vertex.  Boxes are the same with six faces, twelve edges and
 * eight corners in the box's frame. */

static v3 closest_on_tri(v3 p, v3 a, v3 b, v3 c);

/* the centre's ray (o, unit d) against the cylinder of radius r around the segment p0 -> p1 */
/* Both quadratics are solved from an origin advanced to just outside the feature's bounding sphere: with the cast's own
 * origin, tens of metres away, the terms that cancel are a million times the radius squared and fp32 leaves nothing of the
 * answer (false hits with "normals" of length 1.2).  A solution whose contact point is not at the radius is refused. */
#define SWEEP_RADIUS_TOL 0.02f
static int sweep_edge(v3 o, v3 d, float tmax, float r, v3 p0, v3 p1, float *best, v3 *n)
{
	const v3 ed = vsub(p1, p0);
	const float ee = vdot(ed, ed);
	const float t0 = fmaxf(0.0f, vdot(vsub(vmadd(p0, ed, 0.5f), o), d) - ((0.5f * sqrtf(ee)) + r));
	const v3 o2 = vmadd(o, d, t0), m = vsub(o2, p0);
	const float md = vdot(m, ed), dd = vdot(d, ed);
	const float a = ee - (dd * dd);
	if (!(a > (1.0e-5f * ee))) return 0; /* along the edge: the spheres around its ends cover it */
	const float k = vdot(m, m) - (r * r);
	const float c = (ee * k) - (md * md);
	const float b = (ee * vdot(m, d)) - (dd * md);
	const float disc = (b * b) - (a * c);
	if (disc < 0.0f) return 0;
	const float t = (-b - sqrtf(disc)) / a, tt = t0 + t;
	if (!(t >= 0.0f && tt <= tmax && tt < *best)) return 0;
	const float s = md + (t * dd);
	if (s < 0.0f || s > ee) return 0;
	const v3 q = vsub(vmadd(o2, d, t), vmadd(p0, ed, s / ee));
	if (fabsf(vlen2(q) - (r * r)) > (SWEEP_RADIUS_TOL * (r * r))) return 0;
	*best = tt;
	*n = vscale(q, 1.0f / r);
	return 1;
}

static int sweep_vertex(v3 o, v3 d, float tmax, float r, v3 p, float *best, v3 *n)
{
	const float t0 = fmaxf(0.0f, vdot(vsub(p, o), d) - r);
	const v3 o2 = vmadd(o, d, t0), m = vsub(o2, p);
	const float b = vdot(m, d), c = vdot(m, m) - (r * r);
	const float disc = (b * b) - c;
	if (disc < 0.0f) return 0;
	const float t = -b - sqrtf(disc), tt = t0 + t;
	if (!(t >= 0.0f && tt <= tmax && tt < *best)) return 0;
	const v3 q = vsub(vmadd(o2, d, t), p);
	if (fabsf(vlen2(q) - (r * r)) > (SWEEP_RADIUS_TOL * (r * r))) return 0;
	*best = tt;
	*n = vscale(q, 1.0f / r);
	return 1;
}

static int sweep_sphere_tri(v3 o, v3 d, float tmax, float r, v3 a, v3 b, v3 c, float *tout, v3 *nout)
{
	/* overlapping at the start */
	const v3 cp = closest_on_tri(o, a, b, c);
	const v3 dv = vsub(o, cp);
	const float d2 = vlen2(dv);
	if (d2 <= (r * r))
	{
		*tout = 0.0f;
		*nout = d2 > 1.0e-12f ? vscale(dv, 1.0f / sqrtf(d2)) : vneg(d);
		return 1;
	}
	const v3 e1 = vsub(b, a), e2 = vsub(c, a);
	const v3 nn = vcross(e1, e2);
	const float l2 = vlen2(nn);
	float best = 3.0e38f;
	v3 bn = V(0, 0, 0);
	if (l2 > 1.0e-20f)
	{
		v3 n = vscale(nn, 1.0f / sqrtf(l2));
		float s0 = vdot(n, vsub(o, a)), nd = vdot(n, d);
		if (s0 < 0.0f)
		{
			n = vneg(n);
			s0 = -s0;
			nd = -nd;
		}
		if (nd < 0.0f && s0 > r)
		{
			const float t = (r - s0) / nd;
			if (t <= tmax)
			{
				/* where the sphere touches the plane; inside the triangle? */
				const v3 p = vsub(vmadd(o, d, t), vscale(n, r));
				const v3 ca = vcross(vsub(b, a), vsub(p, a)), cb = vcross(vsub(c, b), vsub(p, b)), cc = vcross(vsub(a, c), vsub(p, c));
				const float sa = vdot(ca, nn), sb = vdot(cb, nn), sc = vdot(cc, nn);
				if (sa >= 0.0f && sb >= 0.0f && sc >= 0.0f)
				{
					*tout = t;
					*nout = n;
					return 1;
				}
			}
		}
	}
	if (!(r > 0.0f)) return 0; /* a ray only meets the face */
	int hit = 0;
	hit |= sweep_edge(o, d, tmax, r, a, b, &best, &bn);
	hit |= sweep_edge(o, d, tmax, r, b, c, &best, &bn);
	hit |= sweep_edge(o, d, tmax, r, c, a, &best, &bn);
	hit |= sweep_vertex(o, d, tmax, r, a, &best, &bn);
	hit |= sweep_vertex(o, d, tmax, r, b, &best, &bn);
	hit |= sweep_vertex(o, d, tmax, r, c, &best, &bn);
	if (!hit) return 0;
	*tout = best;
	*nout = bn;
	return 1;
}

static int sweep_sphere_sphere(v3 o, v3 d, float tmax, float r, v3 x, float R, float *tout, v3 *nout)
{
	const float rr = r + R;
	const v3 m = vsub(o, x);
	const float mm = vdot(m, m);
	if (mm <= (rr * rr))
	{
		*tout = 0.0f;
		*nout = mm > 1.0e-12f ? vscale(m, 1.0f / sqrtf(mm)) : vneg(d);
		return 1;
	}
	float best = 3.0e38f;
	v3 n;
	if (!sweep_vertex(o, d, tmax, rr, x, &best, &n)) return 0;
	*tout = best;
	*nout = n;
	return 1;
}

/* in the box's frame; the normal goes back to the world; face = the box face the normal leans to (2k: -axis, 2k+1: +axis) */
static int sweep_sphere_box(v3 o, v3 d, float tmax, float r, const body_t *B, float *tout, v3 *nout, uint32_t *face)
{
	const m33 R = qmat(B->q);
	const v3 lo = mtmul(&R, vsub(o, B->x)), ld = mtmul(&R, d), he = B->he;
	float best = 3.0e38f;
	v3 bn = V(0, 0, 0);
	int hit = 0;
	const v3 cl = V(fminf(fmaxf(lo.x, -he.x), he.x), fminf(fmaxf(lo.y, -he.y), he.y), fminf(fmaxf(lo.z, -he.z), he.z));
	const v3 dv = vsub(lo, cl);
	const float d2 = vlen2(dv);
	if (d2 <= (r * r))
	{
		best = 0.0f;
		bn = d2 > 1.0e-12f ? vscale(dv, 1.0f / sqrtf(d2)) : vneg(ld);
		hit = 1;
	}
	else
	{
		/* faces: the plane he_k + r on the side the centre comes from, hit inside the face's rectangle */
		for (int k = 0; k < 3; k++)
		{
			const float ok = vget(lo, k), dk = vget(ld, k), hk = vget(he, k);
			const float side = ok >= 0.0f ? 1.0f : -1.0f;
			if ((dk * side) >= 0.0f || (ok * side) <= (hk + r)) continue; /* moving away, or not outside this slab */
			const float t = (((hk + r) * side) - ok) / dk;
			if (!(t >= 0.0f && t <= tmax && t < best)) continue;
			const v3 q = vmadd(lo, ld, t);
			const int u = (k + 1) % 3, v = (k + 2) % 3;
			if (fabsf(vget(q, u)) <= vget(he, u) && fabsf(vget(q, v)) <= vget(he, v))
			{
				best = t;
				bn = V(k == 0 ? side : 0.0f, k == 1 ? side : 0.0f, k == 2 ? side : 0.0f);
				hit = 1;
			}
		}
		/* twelve edges (four along each axis), eight corners (a ray, radius 0, only meets the faces) */
		for (int k = 0; k < 3 && r > 0.0f; k++)
			for (int su = -1; su <= 1; su += 2)
				for (int sv = -1; sv <= 1; sv += 2)
				{
					const int u = (k + 1) % 3, v = (k + 2) % 3;
					float p0[3], p1[3];
					p0[k] = -vget(he, k); p1[k] = vget(he, k);
					p0[u] = p1[u] = (float)su * vget(he, u);
					p0[v] = p1[v] = (float)sv * vget(he, v);
					hit |= sweep_edge(lo, ld, tmax, r, V(p0[0], p0[1], p0[2]), V(p1[0], p1[1], p1[2]), &best, &bn);
				}
		for (int sx = -1; sx <= 1 && r > 0.0f; sx += 2)
			for (int sy = -1; sy <= 1; sy += 2)
				for (int sz = -1; sz <= 1; sz += 2)
					hit |= sweep_vertex(lo, ld, tmax, r, V((float)sx * he.x, (float)sy * he.y, (float)sz * he.z), &best, &bn);
	}
	if (!hit) return 0;
	const float ax = fabsf(bn.x), ay = fabsf(bn.y), az = fabsf(bn.z);
	int k = 0;
	if (ay > ax) k = 1;
	if (az > (k == 0 ? ax : ay)) k = 2;
	*face = (uint32_t)(2 * k + (vget(bn, k) > 0.0f ? 1 : 0));
	*tout = best;
	*nout = mmul(&R, bn);
	return 1;
}

static void spherecast_one(const orc_world *w, const orc_sphere_cast *q, orc_cast_hit *h)
{
	const v3 o = V(q->origin[0], q->origin[1], q->origin[2]);
	const v3 d = V(q->dir[0], q->dir[1], q->dir[2]);
	const uint32_t layers = q->mask & 0xFu;
	const int need_flag = (q->mask & (1u << 8)) != 0;
	float best = 3.0e38f;
	uint32_t bbody = ORC_INVALID, bface = ORC_INVALID;
	v3 bn = V(0, 0, 0);
	if (layers & 1u)
		for (uint32_t i = 0; i < w->ntris; i++)
		{
			const tri_t *T = &w->tris[i];
			if (need_flag && !(w->sbodies[T->body].ray_flags & 1u)) continue;
			float t;
			v3 n;
			if (sweep_sphere_tri(o, d, q->tmax, q->radius, T->v0, T->vb, T->vc, &t, &n) && t < best)
			{
				best = t;
				bn = n;
				bbody = ORC_STATIC_BODY_BASE + T->body;
				bface = i;
			}
		}
	for (uint32_t i = 0; i < w->max_bodies; i++)
	{
		const body_t *b = &w->bodies[i];
		if (!b->alive || b->shape == ORC_SHAPE_EMPTY) continue;
		if (!((layers >> b->layer) & 1u)) continue;
		if (need_flag && !(b->ray_flags & 1u)) continue;
		float t;
		v3 n;
		uint32_t f = 0;
		int hit = b->shape == ORC_SHAPE_BOX ? sweep_sphere_box(o, d, q->tmax, q->radius, b, &t, &n, &f)
											: sweep_sphere_sphere(o, d, q->tmax, q->radius, b->x, b->he.x, &t, &n);
		if (hit && t < best)
		{
			best = t;
			bn = n;
			bbody = i;
			bface = f;
		}
	}
	memset(h, 0, sizeof(*h));
	h->world = q->mask >> 16;
	if (bbody == ORC_INVALID)
	{
		h->fraction = RAY_MISS_FRACTION;
		h->body = ORC_INVALID;
		h->face = ORC_INVALID;
	}
	else
	{
		h->fraction = best / q->tmax;
		h->body = bbody;
		h->face = bface;
		h->normal[0] = bn.x;
		h->normal[1] = bn.y;
		h->normal[2] = bn.z;
	}
}

void orc_spherecast(const orc_world *w, const orc_sphere_cast *casts, uint64_t n, orc_cast_hit *hits)
{
	for (uint64_t i = 0; i < n; i++) spherecast_one(w, &casts[i], &hits[i]);
}

static void *job_main(void *arg)
{
	job_t *j = (job_t *)arg;
	j->fn(j->ctx, j->lo, j->hi);
	return NULL;
}

int orc_max_threads(void)
{
	long n = sysconf(_SC_NPROCESSORS_ONLN);
	return n < 1 ? 1 : (n > 256 ? 256 : (int)n);
}

static void parallel_for(int64_t n, void (*fn)(void *, int64_t, int64_t), void *ctx)
{
	int nt = orc_max_threads();
	if ((int64_t)nt > n) nt = n > 0 ? (int)n : 1;
	pthread_t th[256];
	job_t jobs[256];
	for (int t = 0; t < nt; t++)
	{
		jobs[t].fn = fn;
		jobs[t].ctx = ctx;
		jobs[t].lo = n * t / nt;
		jobs[t].hi = n * (t + 1) / nt;
		if (t > 0) pthread_create(&th[t], NULL, job_main, &jobs[t]);
	}
	job_main(&jobs[0]);
	for (int t = 1; t < nt; t++) pthread_join(th[t], NULL);
}

typedef struct { const orc_world *w; const orc_ray *rays; orc_hit *hits; } rayjob_t;
static void ray_range(void *ctx, int64_t lo, int64_t hi)
{
	rayjob_t *r = (rayjob_t *)ctx;
	for (int64_t i = lo; i < hi; i++) raycast_one(r->w, &r->rays[i], &r->hits[i]);
}

void orc_raycast_mt(const orc_world *w, const orc_ray *rays, uint64_t n, orc_hit *hits)
{
	rayjob_t r = {w, rays, hits};
	parallel_for((int64_t)n, ray_range, &r);
}

/* ------------------------------------------------------------------------------------------ narrowphase */

typedef struct
{
	v3 n;        /* a -> b */
	float depth; /* -separation along n */
	int np;
	v3 p1[MAX_POLY], p2[MAX_POLY]; /* world points on a and on b */
} hit_t;

static void body_aabb(const body_t *b, v3 *lo, v3 *hi)
{
	v3 e;
	if (b->shape == ORC_SHAPE_BOX)
	{
		m33 R = qmat(b->q);
		e.x = ((fabsf(R.c0.x) * b->he.x) + (fabsf(R.c1.x) * b->he.y)) + (fabsf(R.c2.x) * b->he.z);
		e.y = ((fabsf(R.c0.y) * b->he.x) + (fabsf(R.c1.y) * b->he.y)) + (fabsf(R.c2.y) * b->he.z);
		e.z = ((fabsf(R.c0.z) * b->he.x) + (fabsf(R.c1.z) * b->he.y)) + (fabsf(R.c2.z) * b->he.z);
	}
	else
		e = V(b->he.x, b->he.x, b->he.x);
	*lo = vsub(b->x, e);
	*hi = vadd(b->x, e);
}

static int aabb_overlap(v3 alo, v3 ahi, v3 blo, v3 bhi, float m)
{
	return (alo.x - m) <= bhi.x && blo.x <= (ahi.x + m) && (alo.y - m) <= bhi.y && blo.y <= (ahi.y + m) &&
		   (alo.z - m) <= bhi.z && blo.z <= (ahi.z + m);
}

/* Sutherland-Hodgman: keep the side where (p - origin).normal >= 0 */
static int clip_plane(const v3 *in, int n, v3 origin, v3 normal, v3 *out)
{
	int m = 0;
	if (n == 0) return 0;
	v3 e1 = in[n - 1];
	float prev = vdot(vsub(origin, e1), normal);
	int prev_in = prev < 0.0f;
	for (int i = 0; i < n; i++)
	{
		v3 e2 = in[i];
		float num = vdot(vsub(origin, e2), normal);
		int cur_in = num < 0.0f;
		if (cur_in != prev_in)
		{
			v3 e12 = vsub(e2, e1);
			float den = vdot(e12, normal);
			if (den != 0.0f)
			{
				if (m < MAX_POLY) out[m++] = vadd(e1, vscale(e12, prev / den));
			}
			else
				cur_in = prev_in;
		}
		if (cur_in && m < MAX_POLY) out[m++] = e2;
		prev = num;
		prev_in = cur_in;
		e1 = e2;
	}
	return m;
}

/* Clip face2 against the side planes of face1 (planes through face1's edges, parallel to `axis`), then keep the
 * points within max_sep of face1's plane (outward unit normal n1) and project them onto it. */
static void manifold_between_faces(const v3 *f1, int n1v, v3 n1, const v3 *f2, int n2v, v3 axis, float max_sep,
								   hit_t *h)
{
	v3 bufa[MAX_POLY], bufb[MAX_POLY];
	v3 *src = bufa, *dst = bufb;
	int n = n2v;
	for (int i = 0; i < n2v; i++) src[i] = f2[i];
	v3 cen = f1[0];
	for (int i = 1; i < n1v; i++) cen = vadd(cen, f1[i]);
	cen = vscale(cen, 1.0f / (float)n1v);
	for (int i = 0; i < n1v && n > 0; i++)
	{
		v3 a = f1[i], b = f1[(i + 1) % n1v];
		v3 pn = vcross(axis, vsub(b, a));
		if (vdot(vsub(cen, a), pn) < 0.0f) pn = vneg(pn);
		n = clip_plane(src, n, a, pn, dst);
		v3 *t = src; src = dst; dst = t;
	}
	h->np = 0;
	for (int i = 0; i < n; i++)
	{
		float dist = vdot(vsub(src[i], f1[0]), n1);
		if (dist <= max_sep)
		{
			h->p2[h->np] = src[i];
			h->p1[h->np] = vsub(src[i], vscale(n1, dist));
			h->np++;
		}
	}
}

/* supporting face of a box in world direction dir; also returns its outward unit normal */
static void box_face(const body_t *b, const m33 *R, v3 dir, v3 *out4, v3 *nout)
{
	v3 l = mtmul(R, dir);
	float ax = fabsf(l.x), ay = fabsf(l.y), az = fabsf(l.z);
	int k = 0;
	if (ay > ax) k = 1;
	if (az > (k == 0 ? ax : ay)) k = 2;
	float s = vget(l, k) < 0.0f ? -1.0f : 1.0f;
	int u = (k + 1) % 3, v = (k + 2) % 3;
	const v3 *cols = &R->c0;
	v3 ck = vscale(cols[k], s * vget(b->he, k));
	v3 cu = vscale(cols[u], vget(b->he, u));
	v3 cv = vscale(cols[v], vget(b->he, v));
	v3 c = vadd(b->x, ck);
	out4[0] = vadd(vadd(c, cu), cv);
	out4[1] = vadd(vsub(c, cu), cv);
	out4[2] = vsub(vsub(c, cu), cv);
	out4[3] = vsub(vadd(c, cu), cv);
	*nout = vscale(cols[k], s);
}

static float box_radius(const body_t *b, const m33 *R, v3 L)
{
	return ((b->he.x * fabsf(vdot(R->c0, L))) + (b->he.y * fabsf(vdot(R->c1, L)))) + (b->he.z * fabsf(vdot(R->c2, L)));
}

/* support vertex of a box in direction dir */
static v3 box_support(const body_t *b, const m33 *R, v3 dir)
{
	v3 l = mtmul(R, dir);
	v3 s = V(l.x < 0.0f ? -b->he.x : b->he.x, l.y < 0.0f ? -b->he.y : b->he.y, l.z < 0.0f ? -b->he.z : b->he.z);
	return vadd(b->x, mmul(R, s));
}

/* closest points between segments p1+s*d1 (s in [-h1,h1]) and p2+t*d2 (t in [-h2,h2]); d1,d2 unit */
static void closest_on_edges(v3 p1, v3 d1, float h1, v3 p2, v3 d2, float h2, v3 *c1, v3 *c2)
{
	v3 r = vsub(p1, p2);
	float b = vdot(d1, d2);
	float c = vdot(d1, r);
	float f = vdot(d2, r);
	float den = 1.0f - (b * b);
	float s = 0.0f;
	if (den > 1.0e-6f) s = ((b * f) - c) / den;
	if (s < -h1) s = -h1;
	if (s > h1) s = h1;
	float t = (b * s) + f;
	if (t < -h2) t = -h2;
	if (t > h2) t = h2;
	s = (b * t) - c;
	if (s < -h1) s = -h1;
	if (s > h1) s = h1;
	*c1 = vadd(p1, vscale(d1, s));
	*c2 = vadd(p2, vscale(d2, t));
}

static int collide_box_box(const body_t *A, const body_t *B, float max_sep, hit_t *h)
{
	m33 RA = qmat(A->q), RB = qmat(B->q);
	const v3 *ca = &RA.c0, *cb = &RB.c0;
	v3 d = vsub(B->x, A->x);
	float best = -3.0e38f;
	v3 bn = V(0, 1, 0);
	int kind = 0, ei = 0, ej = 0; /* kind 0: A face, 1: B face, 2: edge */
	for (int i = 0; i < 3; i++)
	{
		v3 L = ca[i];
		float dl = vdot(d, L);
		float sep = fabsf(dl) - (vget(A->he, i) + box_radius(B, &RB, L));
		if (sep > max_sep) return 0;
		if (i == 0 || sep > best + FACE_AXIS_TOL)
		{
			best = sep;
			bn = dl < 0.0f ? vneg(L) : L;
			kind = 0;
		}
	}
	for (int i = 0; i < 3; i++)
	{
		v3 L = cb[i];
		float dl = vdot(d, L);
		float sep = fabsf(dl) - (box_radius(A, &RA, L) + vget(B->he, i));
		if (sep > max_sep) return 0;
		if (sep > best + FACE_AXIS_TOL)
		{
			best = sep;
			bn = dl < 0.0f ? vneg(L) : L;
			kind = 1;
		}
	}
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++)
		{
			v3 L = vcross(ca[i], cb[j]);
			float l2 = vlen2(L);
			if (l2 < 1.0e-6f) continue;
			L = vscale(L, 1.0f / sqrtf(l2));
			float dl = vdot(d, L);
			float sep = fabsf(dl) - (box_radius(A, &RA, L) + box_radius(B, &RB, L));
			if (sep > max_sep) return 0;
			if (sep > best + EDGE_AXIS_TOL)
			{
				best = sep;
				bn = dl < 0.0f ? vneg(L) : L;
				kind = 2;
				ei = i;
				ej = j;
			}
		}
	h->n = bn;
	h->depth = -best;
	v3 fa[4], fb[4], na, nb;
	box_face(A, &RA, bn, fa, &na);
	box_face(B, &RB, vneg(bn), fb, &nb);
	manifold_between_faces(fa, 4, na, fb, 4, bn, max_sep, h);
	if (h->np == 0)
	{
		if (kind == 2)
		{
			v3 pa = A->x, pb = B->x;
			for (int k = 0; k < 3; k++)
			{
				if (k != ei) pa = vadd(pa, vscale(ca[k], vdot(bn, ca[k]) < 0.0f ? -vget(A->he, k) : vget(A->he, k)));
				if (k != ej) pb = vsub(pb, vscale(cb[k], vdot(bn, cb[k]) < 0.0f ? -vget(B->he, k) : vget(B->he, k)));
			}
			closest_on_edges(pa, ca[ei], vget(A->he, ei), pb, cb[ej], vget(B->he, ej), &h->p1[0], &h->p2[0]);
		}
		else if (kind == 0)
		{
			h->p2[0] = box_support(B, &RB, vneg(bn));
			h->p1[0] = vadd(h->p2[0], vscale(bn, -best));
		}
		else
		{
			h->p1[0] = box_support(A, &RA, bn);
			h->p2[0] = vadd(h->p1[0], vscale(bn, best));
		}
		h->np = 1;
	}
	return 1;
}

static int collide_box_tri(const body_t *A, const tri_t *T, float max_sep, hit_t *h)
{
	m33 R = qmat(A->q);
	const v3 *ca = &R.c0;
	v3 tv[3];
	tv[0] = T->v0;
	tv[1] = T->vb;
	tv[2] = T->vc;
	float best;
	v3 bn;
	int kind = 0, ei = 0, ej = 0;
	{
		/* triangle plane */
		float r = box_radius(A, &R, T->n);
		float cp = vdot(A->x, T->n), tp = vdot(T->v0, T->n);
		float sp = tp - (cp + r), sm = (cp - r) - tp;
		if (sp > sm) { best = sp; bn = T->n; }
		else { best = sm; bn = vneg(T->n); }
		if (best > max_sep) return 0;
	}
	for (int i = 0; i < 3; i++)
	{
		v3 L = ca[i];
		float p0 = vdot(tv[0], L), p1 = vdot(tv[1], L), p2 = vdot(tv[2], L);
		float tmin = fminf(p0, fminf(p1, p2)), tmax = fmaxf(p0, fmaxf(p1, p2));
		float cp = vdot(A->x, L), r = vget(A->he, i);
		float sp = tmin - (cp + r), sm = (cp - r) - tmax;
		float sep = sp > sm ? sp : sm;
		if (sep > max_sep) return 0;
		if (sep > best + FACE_AXIS_TOL)
		{
			best = sep;
			bn = sp > sm ? L : vneg(L);
			kind = 1;
		}
	}
	v3 te[3];
	te[0] = T->e1;
	te[1] = vsub(tv[2], tv[1]);
	te[2] = vsub(tv[0], tv[2]);
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++)
		{
			v3 L = vcross(ca[i], te[j]);
			float l2 = vlen2(L);
			if (l2 < 1.0e-8f * vlen2(te[j])) continue;
			L = vscale(L, 1.0f / sqrtf(l2));
			float p0 = vdot(tv[0], L), p1 = vdot(tv[1], L), p2 = vdot(tv[2], L);
			float tmin = fminf(p0, fminf(p1, p2)), tmax = fmaxf(p0, fmaxf(p1, p2));
			float cp = vdot(A->x, L), r = box_radius(A, &R, L);
			float sp = tmin - (cp + r), sm = (cp - r) - tmax;
			float sep = sp > sm ? sp : sm;
			if (sep > max_sep) return 0;
			if (sep > best + EDGE_AXIS_TOL)
			{
				best = sep;
				bn = sp > sm ? L : vneg(L);
				kind = 2;
				ei = i;
				ej = j;
			}
		}
	h->n = bn;
	h->depth = -best;
	v3 fa[4], na;
	box_face(A, &R, bn, fa, &na);
	manifold_between_faces(fa, 4, na, tv, 3, bn, max_sep, h);
	if (h->np == 0)
	{
		if (kind == 2)
		{
			v3 pa = A->x;
			for (int k = 0; k < 3; k++)
				if (k != ei) pa = vadd(pa, vscale(ca[k], vdot(bn, ca[k]) < 0.0f ? -vget(A->he, k) : vget(A->he, k)));
			float el = vlen(te[ej]);
			v3 ed = vscale(te[ej], 1.0f / el);
			v3 mid = vadd(tv[ej], vscale(te[ej], 0.5f));
			closest_on_edges(pa, ca[ei], vget(A->he, ei), mid, ed, 0.5f * el, &h->p1[0], &h->p2[0]);
		}
		else
		{
			h->p1[0] = box_support(A, &R, bn);
			h->p2[0] = vadd(h->p1[0], vscale(bn, best));
		}
		h->np = 1;
	}
	return 1;
}

/* closest point on triangle (Ericson 5.1.5) */
static v3 closest_on_tri(v3 p, v3 a, v3 b, v3 c)
{
	v3 ab = vsub(b, a), ac = vsub(c, a), ap = vsub(p, a);
	float d1 = vdot(ab, ap), d2 = vdot(ac, ap);
	if (d1 <= 0.0f && d2 <= 0.0f) return a;
	v3 bp = vsub(p, b);
	float d3 = vdot(ab, bp), d4 = vdot(ac, bp);
	if (d3 >= 0.0f && d4 <= d3) return b;
	float vc = (d1 * d4) - (d3 * d2);
	if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) return vadd(a, vscale(ab, d1 / (d1 - d3)));
	v3 cp = vsub(p, c);
	float d5 = vdot(ab, cp), d6 = vdot(ac, cp);
	if (d6 >= 0.0f && d5 <= d6) return c;
	float vb = (d5 * d2) - (d1 * d6);
	if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) return vadd(a, vscale(ac, d2 / (d2 - d6)));
	float va = (d3 * d6) - (d5 * d4);
	if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f)
		return vadd(b, vscale(vsub(c, b), (d4 - d3) / ((d4 - d3) + (d5 - d6))));
	float den = 1.0f / ((va + vb) + vc);
	float v = vb * den, w = vc * den;
	return vadd(vadd(a, vscale(ab, v)), vscale(ac, w));
}

static int collide_sphere_tri(const body_t *A, const tri_t *T, float max_sep, hit_t *h)
{
	v3 c = closest_on_tri(A->x, T->v0, T->vb, T->vc);
	v3 d = vsub(c, A->x);
	float dist = vlen(d);
	float r = A->he.x;
	if ((dist - r) > max_sep) return 0;
	h->n = dist > 1.0e-9f ? vscale(d, 1.0f / dist) : vneg(T->n);
	h->depth = r - dist;
	h->np = 1;
	h->p1[0] = vadd(A->x, vscale(h->n, r));
	h->p2[0] = c;
	return 1;
}

static int collide_sphere_sphere(const body_t *A, const body_t *B, float max_sep, hit_t *h)
{
	v3 d = vsub(B->x, A->x);
	float dist = vlen(d);
	float ra = A->he.x, rb = B->he.x;
	if ((dist - (ra + rb)) > max_sep) return 0;
	h->n = dist > 1.0e-9f ? vscale(d, 1.0f / dist) : V(0, 1, 0);
	h->depth = (ra + rb) - dist;
	h->np = 1;
	h->p1[0] = vadd(A->x, vscale(h->n, ra));
	h->p2[0] = vsub(B->x, vscale(h->n, rb));
	return 1;
}

/* sphere S vs box X.  Normal returned points from the sphere to the box. */
static int collide_sphere_box(const body_t *S, const body_t *X, float max_sep, hit_t *h)
{
	m33 R = qmat(X->q);
	v3 l = mtmul(&R, vsub(S->x, X->x));
	v3 cl = V(fminf(fmaxf(l.x, -X->he.x), X->he.x), fminf(fmaxf(l.y, -X->he.y), X->he.y),
			  fminf(fmaxf(l.z, -X->he.z), X->he.z));
	v3 dl = vsub(cl, l);
	float dist = vlen(dl);
	float r = S->he.x;
	v3 nl;
	v3 pb;
	float depth;
	if (dist > 1.0e-9f)
	{
		if ((dist - r) > max_sep) return 0;
		nl = vscale(dl, 1.0f / dist);
		pb = cl;
		depth = r - dist;
	}
	else
	{
		/* centre inside the box: push out through the nearest face */
		float dx = X->he.x - fabsf(l.x), dy = X->he.y - fabsf(l.y), dz = X->he.z - fabsf(l.z);
		int k = 0;
		float dm = dx;
		if (dy < dm) { dm = dy; k = 1; }
		if (dz < dm) { dm = dz; k = 2; }
		float s = vget(l, k) < 0.0f ? -1.0f : 1.0f;
		nl = V(k == 0 ? -s : 0.0f, k == 1 ? -s : 0.0f, k == 2 ? -s : 0.0f);
		pb = l;
		if (k == 0) pb.x = s * X->he.x;
		if (k == 1) pb.y = s * X->he.y;
		if (k == 2) pb.z = s * X->he.z;
		depth = r + dm;
	}
	h->n = mmul(&R, nl);
	h->depth = depth;
	h->np = 1;
	h->p1[0] = vadd(S->x, vscale(h->n, r));
	h->p2[0] = vadd(X->x, mmul(&R, pb));
	return 1;
}

/* Reduce to <= 4 points keeping the largest, deepest spread (restating Jolt's PruneContactPoints). */
static void prune_points(v3 xa, v3 axis, int *np, v3 *p1, v3 *p2)
{
	int n = *np;
	if (n <= 4) return;
	v3 proj[2 * MAX_POLY];
	float dsq[2 * MAX_POLY];
	for (int i = 0; i < n; i++)
	{
		v3 v1 = vsub(p1[i], xa);
		proj[i] = vsub(v1, vscale(axis, vdot(v1, axis)));
		dsq[i] = fmaxf(1.0e-6f, vlen2(vsub(p2[i], p1[i])));
	}
	int i1 = 0;
	float best = -1.0f;
	for (int i = 0; i < n; i++)
	{
		float v = fmaxf(1.0e-6f, vlen2(proj[i])) * dsq[i];
		if (v > best) { best = v; i1 = i; }
	}
	int i2 = -1;
	best = -1.0f;
	for (int i = 0; i < n; i++)
		if (i != i1)
		{
			float v = fmaxf(1.0e-6f, vlen2(vsub(proj[i], proj[i1]))) * dsq[i];
			if (v > best) { best = v; i2 = i; }
		}
	int i3 = -1, i4 = -1;
	float mn = 0.0f, mx = 0.0f;
	v3 perp = vcross(vsub(proj[i2], proj[i1]), axis);
	for (int i = 0; i < n; i++)
		if (i != i1 && i != i2)
		{
			float v = vdot(perp, vsub(proj[i], proj[i1]));
			if (v < mn) { mn = v; i3 = i; }
			else if (v > mx) { mx = v; i4 = i; }
		}
	v3 o1[4], o2[4];
	int m = 0;
	o1[m] = p1[i1]; o2[m++] = p2[i1];
	if (i3 >= 0) { o1[m] = p1[i3]; o2[m++] = p2[i3]; }
	o1[m] = p1[i2]; o2[m++] = p2[i2];
	if (i4 >= 0) { o1[m] = p1[i4]; o2[m++] = p2[i4]; }
	for (int i = 0; i < m; i++) { p1[i] = o1[i]; p2[i] = o2[i]; }
	*np = m;
}

/* ------------------------------------------------------------------------------------------ contact generation */

static int layers_collide(const body_t *a, const body_t *b)
{
	/* ObjectLayerShouldCollide, both orders (engine/src/physics/Physics.c:35-52): DYNAMIC/PLAYER vs STATIC/DYNAMIC/SENSOR */
	int a_init = a->layer == 1 || a->layer == 2, b_init = b->layer == 1 || b->layer == 2;
	int a_tgt = a->layer == 0 || a->layer == 1 || a->layer == 3, b_tgt = b->layer == 0 || b->layer == 1 || b->layer == 3;
	return (a_init && b_tgt) || (b_init && a_tgt);
}

static manifold_t *push_manifold(orc_world *w, int *err)
{
	if (w->nman >= w->max_manifolds)
	{
		*err = 4;
		return NULL;
	}
	manifold_t *m = &w->man[w->nman++];
	memset(m, 0, sizeof(*m));
	return m;
}

static void store_points(manifold_t *m, const body_t *A, const body_t *B, int np, const v3 *p1, const v3 *p2)
{
	m33 RA = qmat(A->q);
	m->np = np;
	for (int i = 0; i < np; i++)
	{
		m->p1l[i] = mtmul(&RA, vsub(p1[i], A->x));
		if (B)
		{
			m33 RB = qmat(B->q);
			m->p2l[i] = mtmul(&RB, vsub(p2[i], B->x));
		}
		else
			m->p2l[i] = p2[i];
	}
}

/* Candidate partners for large worlds.  The pair loop of find_contacts visits (i, j > i) in ascending order and applies
 * every filter itself; for more than SWEEP_MIN_BODIES slots the j's it visits come from this sort-and-sweep over the
 * x axis instead of from i+1..max_bodies.  The list is a superset of the overlapping pairs in the same (i, j) order, so
 * the contacts found — and their order — are those of the all-pairs loop. */
#define SWEEP_MIN_BODIES 256
typedef struct { float lo; uint32_t id; } sweep_key_t;
static int sweep_key_cmp(const void *a, const void *b)
{
	const sweep_key_t *x = (const sweep_key_t *)a, *y = (const sweep_key_t *)b;
	return x->lo < y->lo ? -1 : x->lo > y->lo ? 1 : x->id < y->id ? -1 : x->id > y->id;
}
static int u64_cmp(const void *a, const void *b)
{
	const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
	return x < y ? -1 : x > y;
}
/* the sweep proper over sorted positions [lo, hi) of slice k: pairs (min << 32 | max) into the slice's own list */
typedef struct
{
	uint32_t n, nslices;
	float m;
	const sweep_key_t *keys;
	const v3 *slo, *shi;
	uint64_t **pairs;
	size_t *cnt;
} sweep_t;
static void sweep_slices(void *ctx, int64_t k0, int64_t k1)
{
	sweep_t *S = (sweep_t *)ctx;
	const float m = S->m;
	for (int64_t k = k0; k < k1; k++)
	{
		const uint32_t p0 = (uint32_t)((uint64_t)S->n * k / S->nslices), p1 = (uint32_t)((uint64_t)S->n * (k + 1) / S->nslices);
		size_t cap = 4 * (size_t)(p1 - p0) + 64, cnt = 0;
		uint64_t *pairs = (uint64_t *)malloc(sizeof(uint64_t) * cap);
		for (uint32_t p = p0; p < p1; p++)
		{
			const uint32_t a = S->keys[p].id;
			const v3 alo = S->slo[p], ahi = S->shi[p];
			const float reach = ahi.x + m;
			for (uint32_t q = p + 1; q < S->n && S->slo[q].x <= reach; q++)
			{
				if (alo.z - m > S->shi[q].z || S->slo[q].z - m > ahi.z || alo.y - m > S->shi[q].y || S->slo[q].y - m > ahi.y) continue;
				const uint32_t b = S->keys[q].id;
				if (cnt == cap)
				{
					cap *= 2;
					pairs = (uint64_t *)realloc(pairs, sizeof(uint64_t) * cap);
				}
				pairs[cnt++] = a < b ? ((uint64_t)a << 32) | b : ((uint64_t)b << 32) | a;
			}
		}
		S->pairs[k] = pairs;
		S->cnt[k] = cnt;
	}
}
static void pool_for(int64_t n, void (*fn)(void *, int64_t, int64_t), void *ctx, int64_t grain);

/* returns the sorted candidate pairs (i << 32 | j), *first[i] = index of body i's first pair, first[max_bodies] = count;
 * threaded: the sweep runs in slices over the host threads (orc_step_mt) */
static uint64_t *sweep_candidates(const orc_world *w, uint32_t **first_out, int threaded)
{
	const uint32_t nb = w->max_bodies;
	v3 *lo = (v3 *)malloc(sizeof(v3) * nb), *hi = (v3 *)malloc(sizeof(v3) * nb);
	sweep_key_t *keys = (sweep_key_t *)malloc(sizeof(sweep_key_t) * nb);
	uint32_t n = 0;
	for (uint32_t i = 0; i < nb; i++)
	{
		const body_t *B = &w->bodies[i];
		if (!B->alive || B->shape == ORC_SHAPE_EMPTY) continue;
		body_aabb(B, &lo[i], &hi[i]);
		keys[n].lo = lo[i].x;
		keys[n++].id = i;
	}
	qsort(keys, n, sizeof(sweep_key_t), sweep_key_cmp);
	/* the boxes once more in sorted order: the sweep walks them front to back */
	v3 *slo = (v3 *)malloc(sizeof(v3) * (n + 1)), *shi = (v3 *)malloc(sizeof(v3) * (n + 1));
	for (uint32_t p = 0; p < n; p++)
	{
		slo[p] = lo[keys[p].id];
		shi[p] = hi[keys[p].id];
	}
	enum { MAX_SLICES = 256 };
	uint64_t *slice_pairs[MAX_SLICES];
	size_t slice_cnt[MAX_SLICES];
	sweep_t S = {n, threaded ? MAX_SLICES : 1, 2.0f * SPECULATIVE_DISTANCE + 1e-3f, keys, slo, shi, slice_pairs, slice_cnt};
	if (threaded) pool_for(S.nslices, sweep_slices, &S, 2);
	else sweep_slices(&S, 0, 1);
	size_t cnt = 0;
	for (uint32_t k = 0; k < S.nslices; k++) cnt += slice_cnt[k];
	uint64_t *pairs = (uint64_t *)malloc(sizeof(uint64_t) * (cnt + 1));
	cnt = 0;
	for (uint32_t k = 0; k < S.nslices; k++)
	{
		memcpy(pairs + cnt, slice_pairs[k], sizeof(uint64_t) * slice_cnt[k]);
		cnt += slice_cnt[k];
		free(slice_pairs[k]);
	}
	qsort(pairs, cnt, sizeof(uint64_t), u64_cmp);
	uint32_t *first = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)nb + 1));
	size_t k = 0;
	for (uint32_t i = 0; i <= nb; i++)
	{
		while (k < cnt && (uint32_t)(pairs[k] >> 32) < i) k++;
		first[i] = (uint32_t)k;
	}
	free(slo);
	free(shi);
	free(lo);
	free(hi);
	free(keys);
	*first_out = first;
	return pairs;
}

/* contacts whose first body lies in [i0, i1), appended to w->man / w->sens */
static void find_contacts_among(orc_world *w, int *err, const uint64_t *cand, const uint32_t *cand_first, uint32_t i0, uint32_t i1)
{
	for (uint32_t i = i0; i < i1; i++)
	{
		body_t *A = &w->bodies[i];
		if (!A->alive || A->shape == ORC_SHAPE_EMPTY) continue;
		v3 alo, ahi;
		body_aabb(A, &alo, &ahi);
		/* (a) against the static triangle soup: dynamic bodies of layers that collide with STATIC */
		if (is_dyn(A) && !A->sensor && (A->layer == 1 || A->layer == 2))
		{
			/* per static body: up to MAX_SLOTS manifolds grouped by normal */
			uint32_t cur_body = ORC_INVALID;
			v3 sp1[MAX_SLOTS][8], sp2[MAX_SLOTS][8];
			int snp[MAX_SLOTS];
			manifold_t *slots[MAX_SLOTS];
			int nslots = 0;
			for (uint32_t t = 0; t <= w->ntris; t++)
			{
				const tri_t *T = t < w->ntris ? &w->tris[t] : NULL;
				if (!T || T->body != cur_body)
				{
					/* flush slots of the previous static body */
					for (int s = 0; s < nslots; s++) store_points(slots[s], A, NULL, snp[s], sp1[s], sp2[s]);
					nslots = 0;
					if (!T) break;
					cur_body = T->body;
				}
				if (!aabb_overlap(alo, ahi, T->lo, T->hi, SPECULATIVE_DISTANCE)) continue;
				hit_t h;
				int hit = A->shape == ORC_SHAPE_BOX ? collide_box_tri(A, T, SPECULATIVE_DISTANCE, &h)
													: collide_sphere_tri(A, T, SPECULATIVE_DISTANCE, &h);
				if (!hit) continue;
				prune_points(A->x, h.n, &h.np, h.p1, h.p2);
				int s = -1;
				for (int k = 0; k < nslots; k++)
					if (vdot(slots[k]->n, h.n) >= NORMAL_COS_MAX_DELTA)
					{
						s = k;
						break;
					}
				if (s < 0)
				{
					if (nslots == MAX_SLOTS) continue;
					manifold_t *m = push_manifold(w, err);
					if (!m) return;
					m->a = i;
					m->b = ORC_STATIC_BODY_BASE + T->body;
					m->n = h.n;
					m->depth = h.depth;
					m->friction = sqrtf(A->friction * w->sbodies[T->body].friction);
					m->restitution = A->restitution;
					s = nslots++;
					slots[s] = m;
					snp[s] = 0;
					m->ord = (uint32_t)s;
					m->tri = t;
				}
				else if (h.depth > slots[s]->depth)
				{
					slots[s]->depth = h.depth;
					slots[s]->n = h.n;
				}
				for (int k = 0; k < h.np; k++)
				{
					sp1[s][snp[s]] = h.p1[k];
					sp2[s][snp[s]] = h.p2[k];
					snp[s]++;
				}
				prune_points(A->x, slots[s]->n, &snp[s], sp1[s], sp2[s]);
			}
		}
		/* (b) against higher-numbered slot bodies */
		const uint32_t npartners = cand ? cand_first[i + 1] - cand_first[i] : w->max_bodies - (i + 1);
		for (uint32_t q = 0; q < npartners; q++)
		{
			const uint32_t j = cand ? (uint32_t)cand[cand_first[i] + q] : i + 1 + q;
			body_t *B = &w->bodies[j];
			if (!B->alive || B->shape == ORC_SHAPE_EMPTY) continue;
			/* at least one awake dynamic body — or a moving kinematic body reaching a sleeper, which only wakes it */
			const int solved = is_dyn(A) || is_dyn(B);
			if (!solved && !((is_active(A) && B->asleep) || (is_active(B) && A->asleep))) continue;
			if (!layers_collide(A, B)) continue;
			v3 blo, bhi;
			body_aabb(B, &blo, &bhi);
			if (!aabb_overlap(alo, ahi, blo, bhi, SPECULATIVE_DISTANCE)) continue;
			hit_t h;
			int hit;
			if (A->shape == ORC_SHAPE_BOX && B->shape == ORC_SHAPE_BOX)
				hit = collide_box_box(A, B, SPECULATIVE_DISTANCE, &h);
			else if (A->shape == ORC_SHAPE_SPHERE && B->shape == ORC_SHAPE_SPHERE)
				hit = collide_sphere_sphere(A, B, SPECULATIVE_DISTANCE, &h);
			else if (A->shape == ORC_SHAPE_SPHERE)
				hit = collide_sphere_box(A, B, SPECULATIVE_DISTANCE, &h);
			else
			{
				hit = collide_sphere_box(B, A, SPECULATIVE_DISTANCE, &h);
				if (hit)
				{
					h.n = vneg(h.n);
					v3 t = h.p1[0];
					h.p1[0] = h.p2[0];
					h.p2[0] = t;
				}
			}
			if (!hit) continue;
			if (A->sensor || B->sensor)
			{
				/* sensors produce events only (not part of the solve) */
				if (w->nsens < w->max_manifolds) w->sens[w->nsens++] = ((uint64_t)i << 32) | j;
				continue;
			}
			prune_points(A->x, h.n, &h.np, h.p1, h.p2);
			/* a contact with an active body wakes a sleeper; it takes part in the solve from the next sub-step on */
			if (h.np > 0)
			{
				if (A->asleep && is_active(B)) A->wake_mark = 1;
				if (B->asleep && is_active(A)) B->wake_mark = 1;
			}
			if (!solved) continue;
			manifold_t *m = push_manifold(w, err);
			if (!m) return;
			m->a = i;
			m->b = j;
			m->n = h.n;
			m->depth = h.depth;
			m->friction = sqrtf(A->friction * B->friction);
			m->restitution = fmaxf(A->restitution, B->restitution);
			store_points(m, A, B, h.np, h.p1, h.p2);
		}
	}
}

static void find_contacts(orc_world *w, int *err)
{
	uint32_t *cand_first = NULL;
	uint64_t *cand = w->max_bodies > SWEEP_MIN_BODIES ? sweep_candidates(w, &cand_first, 0) : NULL;
	w->nman = 0;
	w->nsens = 0;
	find_contacts_among(w, err, cand, cand_first, 0, w->max_bodies);
	free(cand);
	free(cand_first);
}

/* ------------------------------------------------------------------------------------------ solver */

static void warm_start_match_range(orc_world *w, const uint32_t *run, uint32_t lo, uint32_t hi);
static void warm_start_match(orc_world *w)
{
	/* both lists are in canonical order (body a ascending), so the old manifolds of body a are one contiguous run */
	uint32_t *run = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)w->max_bodies + 1));
	for (uint32_t a = 0, j = 0; a <= w->max_bodies; a++)
	{
		while (j < w->nprev && w->prev[j].a < a) j++;
		run[a] = j;
	}
	warm_start_match_range(w, run, 0, w->nman);
	free(run);
}

static void warm_start_match_range(orc_world *w, const uint32_t *run, uint32_t lo, uint32_t hi)
{
	for (uint32_t i = lo; i < hi; i++)
	{
		manifold_t *m = &w->man[i];
		int got_cf = 0;
		for (uint32_t j = run[m->a]; j < run[m->a + 1]; j++)
		{
			const manifold_t *o = &w->prev[j];
			if (o->b != m->b || o->tri != m->tri) continue;
			for (int p = 0; p < m->np; p++)
			{
				if (m->ln[p] != 0.0f) continue;
				for (int k = 0; k < o->np; k++)
					if (vlen2(vsub(m->p1l[p], o->p1l[k])) < PRESERVE_LAMBDA_MAX_DIST_SQ &&
						vlen2(vsub(m->p2l[p], o->p2l[k])) < PRESERVE_LAMBDA_MAX_DIST_SQ)
					{
						m->ln[p] = o->ln[k];
						/* the friction impulse of the manifold comes from the first old manifold a point is found in */
						if (!got_cf)
						{
							got_cf = 1;
							m->cf[0] = o->cf[0];
							m->cf[1] = o->cf[1];
							m->cf[2] = o->cf[2];
						}
						break;
					}
			}
		}
	}
}

static int colour_manifolds(orc_world *w)
{
	/* greedy first-fit in canonical order; only dynamic bodies constrain the colour */
	uint32_t nb = w->max_bodies;
	uint64_t *used = (uint64_t *)calloc(nb, sizeof(uint64_t));
	int ncol = 0;
	for (uint32_t i = 0; i < w->nman; i++)
	{
		manifold_t *m = &w->man[i];
		uint64_t u = 0;
		int a_dyn = is_dyn(&w->bodies[m->a]);
		int b_dyn = m->b < ORC_STATIC_BODY_BASE && is_dyn(&w->bodies[m->b]);
		if (a_dyn) u |= used[m->a];
		if (b_dyn) u |= used[m->b];
		int c = 0;
		while (c < 63 && ((u >> c) & 1u)) c++;
		m->colour = c;
		if (a_dyn) used[m->a] |= 1ull << c;
		if (b_dyn) used[m->b] |= 1ull << c;
		if (c + 1 > ncol) ncol = c + 1;
	}
	free(used);
	uint32_t k = 0;
	for (int c = 0; c < ncol; c++)
		for (uint32_t i = 0; i < w->nman; i++)
			if (w->man[i].colour == c) w->order[k++] = i;
	return ncol;
}

/* Mode 1 (wide worlds): Jones-Plassmann colouring with hashed priorities.  In each round every uncoloured manifold
 * that outranks all its uncoloured neighbours (manifolds sharing a dynamic body) takes the smallest colour none of
 * its already-coloured neighbours has; decisions of a round only see colours of earlier rounds.  The result is a
 * function of the contact graph alone, so a parallel evaluation gives the same colours. */
static uint32_t man_prio(uint32_t a, uint32_t b, uint32_t ord)
{
	uint32_t h = (a * 0x9E3779B1u) ^ ((b + ord * 0x7F4A7C15u) * 0x85EBCA77u);
	h ^= h >> 15;
	h *= 0x2C1B3C6Du;
	h ^= h >> 12;
	h *= 0x297A2D39u;
	h ^= h >> 15;
	return h;
}

static int outranks(const manifold_t *x, const manifold_t *y)
{
	if (x->prio != y->prio) return x->prio > y->prio;
	if (x->a != y->a) return x->a > y->a;
	if (x->b != y->b) return x->b > y->b;
	return x->ord > y->ord;
}

static int colour_manifolds_jp(orc_world *w)
{
	const uint32_t n = w->nman, nb = w->max_bodies;
	/* incidence lists of the dynamic bodies */
	uint32_t *cnt = (uint32_t *)calloc(nb + 1, sizeof(uint32_t));
	for (uint32_t i = 0; i < n; i++)
	{
		manifold_t *m = &w->man[i];
		m->prio = man_prio(m->a, m->b, m->ord);
		m->colour = -1;
		if (is_dyn(&w->bodies[m->a])) cnt[m->a + 1]++;
		if (m->b < ORC_STATIC_BODY_BASE && is_dyn(&w->bodies[m->b])) cnt[m->b + 1]++;
	}
	for (uint32_t i = 0; i < nb; i++) cnt[i + 1] += cnt[i];
	uint32_t *adj = (uint32_t *)malloc((cnt[nb] + 1) * sizeof(uint32_t));
	uint32_t *cur = (uint32_t *)malloc((nb + 1) * sizeof(uint32_t));
	memcpy(cur, cnt, (nb + 1) * sizeof(uint32_t));
	for (uint32_t i = 0; i < n; i++)
	{
		manifold_t *m = &w->man[i];
		if (is_dyn(&w->bodies[m->a])) adj[cur[m->a]++] = i;
		if (m->b < ORC_STATIC_BODY_BASE && is_dyn(&w->bodies[m->b])) adj[cur[m->b]++] = i;
	}
	int *pending = (int *)malloc((n + 1) * sizeof(int));
	uint32_t left = n;
	int ncol = 0;
	while (left)
	{
		for (uint32_t i = 0; i < n; i++)
		{
			const manifold_t *m = &w->man[i];
			pending[i] = -1;
			if (m->colour != -1) continue;
			int top = 1;
			uint64_t used = 0;
			const uint32_t ends[2] = {m->a, m->b};
			for (int e = 0; e < 2 && top; e++)
			{
				uint32_t body = ends[e];
				if (body >= ORC_STATIC_BODY_BASE || !is_dyn(&w->bodies[body])) continue;
				for (uint32_t k = cnt[body]; k < cnt[body + 1]; k++)
				{
					uint32_t other = adj[k];
					if (other == i) continue;
					int oc = w->man[other].colour;
					if (oc == -1)
					{
						if (outranks(&w->man[other], m))
						{
							top = 0;
							break;
						}
					}
					else
						used |= 1ull << oc;
				}
			}
			if (top)
			{
				int c = 0;
				while (c < 63 && ((used >> c) & 1u)) c++;
				pending[i] = c;
			}
		}
		for (uint32_t i = 0; i < n; i++)
			if (pending[i] >= 0)
			{
				w->man[i].colour = pending[i];
				if (pending[i] + 1 > ncol) ncol = pending[i] + 1;
				left--;
			}
	}
	free(pending);
	free(cur);
	free(adj);
	free(cnt);
	uint32_t k = 0;
	for (int c = 0; c < ncol; c++)
		for (uint32_t i = 0; i < n; i++)
			if (w->man[i].colour == c) w->order[k++] = i;
	return ncol;
}

/* World-space inverse inertia R diag(inv_i) R^T as 6 unique entries (mirrored), locked rotation axes zeroed.
 * Refreshed once per sub-step (after forces) and at the start of each manifold's position pass. */
static void body_world_inertia(body_t *b)
{
	if (!is_dyn(b))
	{
		memset(b->M, 0, sizeof(b->M));
		b->im = 0.0f;
		return;
	}
	m33 R = qmat(b->q);
	v3 s0 = vscale(R.c0, b->inv_inertia.x), s1 = vscale(R.c1, b->inv_inertia.y), s2 = vscale(R.c2, b->inv_inertia.z);
	float xx = ((s0.x * R.c0.x) + (s1.x * R.c1.x)) + (s2.x * R.c2.x);
	float xy = ((s0.x * R.c0.y) + (s1.x * R.c1.y)) + (s2.x * R.c2.y);
	float xz = ((s0.x * R.c0.z) + (s1.x * R.c1.z)) + (s2.x * R.c2.z);
	float yy = ((s0.y * R.c0.y) + (s1.y * R.c1.y)) + (s2.y * R.c2.y);
	float yz = ((s0.y * R.c0.z) + (s1.y * R.c1.z)) + (s2.y * R.c2.z);
	float zz = ((s0.z * R.c0.z) + (s1.z * R.c1.z)) + (s2.z * R.c2.z);
	const int lx = !(b->dofs & 8u), ly = !(b->dofs & 16u), lz = !(b->dofs & 32u);
	b->M[0] = lx ? 0.0f : xx;
	b->M[1] = (lx || ly) ? 0.0f : xy;
	b->M[2] = (lx || lz) ? 0.0f : xz;
	b->M[3] = ly ? 0.0f : yy;
	b->M[4] = (ly || lz) ? 0.0f : yz;
	b->M[5] = lz ? 0.0f : zz;
	b->im = b->inv_mass;
}

static v3 sym_mul(const float *M, v3 v)
{
	return V(fmaf(M[2], v.z, fmaf(M[1], v.y, M[0] * v.x)), fmaf(M[4], v.z, fmaf(M[3], v.y, M[1] * v.x)),
			 fmaf(M[5], v.z, fmaf(M[4], v.y, M[2] * v.x)));
}

static v3 mask_lin(uint32_t dofs, v3 a)
{
	if (!(dofs & 1u)) a.x = 0.0f;
	if (!(dofs & 2u)) a.y = 0.0f;
	if (!(dofs & 4u)) a.z = 0.0f;
	return a;
}

static const float ZERO_M[6] = {0, 0, 0, 0, 0, 0};

/* 1 / (J M^-1 J^T) for a contact axis */
static float eff_mass(float ima, const float *MA, float imb, const float *MB, v3 r1, v3 r2, v3 axis)
{
	v3 r1xa = vcross(r1, axis), r2xa = vcross(r2, axis);
	float k = ((ima + imb) + vdot(r1xa, sym_mul(MA, r1xa))) + vdot(r2xa, sym_mul(MB, r2xa));
	return k > 0.0f ? 1.0f / k : 0.0f;
}

/* v - t * s */
static inline v3 vmsub(v3 v, v3 t, float s) { return V(fmaf(-t.x, s, v.x), fmaf(-t.y, s, v.y), fmaf(-t.z, s, v.z)); }

/* the two bodies' velocities as the rows see them (b static: zero) */
typedef struct { v3 va, wa, vb, wb; } vel_t;

static vel_t load_vel(const body_t *A, const body_t *B)
{
	vel_t u;
	u.va = A->v;
	u.wa = A->w;
	u.vb = B ? B->v : V(0, 0, 0);
	u.wb = B ? B->w : V(0, 0, 0);
	return u;
}

static void store_vel(body_t *A, body_t *B, const vel_t *u)
{
	if (is_dyn(A))
	{
		A->v = u->va;
		A->w = u->wa;
	}
	if (B && is_dyn(B))
	{
		B->v = u->vb;
		B->w = u->wb;
	}
}

static row_t row_setup(float ima, const float *MA, float imb, const float *MB, v3 r1, v3 r2, v3 axis)
{
	row_t r;
	r.a1 = vcross(r1, axis);
	r.a2 = vcross(r2, axis);
	r.I1 = sym_mul(MA, r.a1);
	r.I2 = sym_mul(MB, r.a2);
	float k = ((ima + imb) + vdot(r.a1, r.I1)) + vdot(r.a2, r.I2);
	r.em = k > 0.0f ? 1.0f / k : 0.0f;
	return r;
}

static inline float row_jv(const row_t *r, v3 axis, const vel_t *u)
{
	return (vdot(axis, u->va) + vdot(r->a1, u->wa)) - (vdot(axis, u->vb) + vdot(r->a2, u->wb));
}

static inline void row_apply(const row_t *r, v3 lA, v3 lB, float d, vel_t *u)
{
	u->va = vmsub(u->va, lA, d);
	u->wa = vmsub(u->wa, r->I1, d);
	u->vb = vmadd(u->vb, lB, d);
	u->wb = vmadd(u->wb, r->I2, d);
}

/* twist about the normal: angular parts only */
static inline float twist_jv(v3 n, const vel_t *u) { return vdot(n, u->wa) - vdot(n, u->wb); }
static inline void twist_apply(const row_t *r, float d, vel_t *u)
{
	u->wa = vmsub(u->wa, r->I1, d);
	u->wb = vmadd(u->wb, r->I2, d);
}

/* per sub-step set-up (reads body state only): the manifold's rows and the speculative / restitution bias */
static void setup_manifold(orc_world *w, manifold_t *m, float h)
{
	body_t *A = &w->bodies[m->a];
	body_t *B = m->b < ORC_STATIC_BODY_BASE ? &w->bodies[m->b] : NULL;
	const float imb = B ? B->im : 0.0f;
	const float *MB = B ? B->M : ZERO_M;
	rows_t *R = &m->rows;
	m->t1 = vperp(m->n);
	m->t2 = vcross(m->n, m->t1);
	R->nA = mask_lin(A->dofs, vscale(m->n, A->im));
	R->tA[0] = mask_lin(A->dofs, vscale(m->t1, A->im));
	R->tA[1] = mask_lin(A->dofs, vscale(m->t2, A->im));
	R->nB = B ? mask_lin(B->dofs, vscale(m->n, imb)) : V(0, 0, 0);
	R->tB[0] = B ? mask_lin(B->dofs, vscale(m->t1, imb)) : V(0, 0, 0);
	R->tB[1] = B ? mask_lin(B->dofs, vscale(m->t2, imb)) : V(0, 0, 0);
	const vel_t u = load_vel(A, B);
	v3 mid[4];
	v3 csum = V(0, 0, 0);
	for (int k = 0; k < m->np; k++)
	{
		v3 p1 = vadd(A->x, qrot(A->q, m->p1l[k]));
		v3 p2 = B ? vadd(B->x, qrot(B->q, m->p2l[k])) : m->p2l[k];
		mid[k] = vscale(vadd(p1, p2), 0.5f);
		csum = vadd(csum, mid[k]);
		v3 r1 = vsub(mid[k], A->x);
		v3 r2 = B ? vsub(mid[k], B->x) : V(0, 0, 0);
		R->n[k] = row_setup(A->im, A->M, imb, MB, r1, r2, m->n);
		float pen = vdot(vsub(p1, p2), m->n);
		float bias = fmaxf(0.0f, -pen / h);
		if (m->restitution > 0.0f)
		{
			float nv = -row_jv(&R->n[k], m->n, &u); /* separating speed of b relative to a */
			if (nv < -MIN_VELOCITY_FOR_RESTITUTION) bias = m->restitution * nv;
		}
		R->bias[k] = bias;
	}
	/* rows of points the manifold does not have are zero: they measure nothing and apply nothing */
	for (int k = m->np; k < 4; k++)
	{
		memset(&R->n[k], 0, sizeof(R->n[k]));
		R->bias[k] = 0.0f;
		m->ln[k] = 0.0f;
	}
	if (m->np == 0) return;
	const float inv_np = 1.0f / (float)m->np;
	const v3 c = vscale(csum, inv_np);
	float s = 0.0f;
	for (int k = 0; k < m->np; k++) s = s + vlen2(vsub(mid[k], c));
	R->rp = sqrtf(s * inv_np);
	const v3 rc1 = vsub(c, A->x), rc2 = B ? vsub(c, B->x) : V(0, 0, 0);
	R->t[0] = row_setup(A->im, A->M, imb, MB, rc1, rc2, m->t1);
	R->t[1] = row_setup(A->im, A->M, imb, MB, rc1, rc2, m->t2);
	R->w.a1 = m->n;
	R->w.a2 = m->n;
	R->w.I1 = sym_mul(A->M, m->n);
	R->w.I2 = sym_mul(MB, m->n);
	float kw = vdot(m->n, R->w.I1) + vdot(m->n, R->w.I2);
	R->w.em = kw > 0.0f ? 1.0f / kw : 0.0f;
}

/* re-apply the impulses carried over from the previous sub-step */
static void warm_start(orc_world *w, manifold_t *m)
{
	body_t *A = &w->bodies[m->a];
	body_t *B = m->b < ORC_STATIC_BODY_BASE ? &w->bodies[m->b] : NULL;
	const rows_t *R = &m->rows;
	if (m->np == 0) return;
	vel_t u = load_vel(A, B);
	for (int k = 0; k < 4; k++) row_apply(&R->n[k], R->nA, R->nB, m->ln[k], &u);
	row_apply(&R->t[0], R->tA[0], R->tB[0], m->cf[0], &u);
	row_apply(&R->t[1], R->tA[1], R->tB[1], m->cf[1], &u);
	twist_apply(&R->w, m->cf[2], &u);
	store_vel(A, B, &u);
}

/* One velocity iteration of a manifold.  Friction first (non-penetration is more important, so it goes last): the two
 * tangent rows through the centroid share one limit, friction * (sum of the normal impulses), then the twist row with
 * that limit times the patch radius.  The non-penetration rows run forwards in even iterations and backwards in odd
 * ones: with a fixed order the last point of every manifold always ends exact and the first one always carries the
 * residual, and that bias turns a tall stack's residuals into a slow whirl that never dies (tests/test_oracle.py,
 * test_kicked_column_comes_to_rest). */
static void solve_velocity(orc_world *w, manifold_t *m, uint32_t it)
{
	body_t *A = &w->bodies[m->a];
	body_t *B = m->b < ORC_STATIC_BODY_BASE ? &w->bodies[m->b] : NULL;
	const rows_t *R = &m->rows;
	if (m->np == 0) return;
	vel_t u = load_vel(A, B);
	const float maxf = m->friction * (((m->ln[0] + m->ln[1]) + m->ln[2]) + m->ln[3]);
	/* nothing to hold with and nothing held (speculative points that do not touch): the rows stay at zero */
	if (!(maxf == 0.0f && m->cf[0] == 0.0f && m->cf[1] == 0.0f))
	{
		float l1 = m->cf[0] + (R->t[0].em * row_jv(&R->t[0], m->t1, &u));
		float l2 = m->cf[1] + (R->t[1].em * row_jv(&R->t[1], m->t2, &u));
		float sq = (l1 * l1) + (l2 * l2);
		if (sq > (maxf * maxf))
		{
			/* no normal impulse: 0 / sqrt(sq) is that zero */
			float sc = maxf == 0.0f ? maxf : maxf / sqrtf(sq);
			l1 = l1 * sc;
			l2 = l2 * sc;
		}
		const float d1 = l1 - m->cf[0], d2 = l2 - m->cf[1];
		m->cf[0] = l1;
		m->cf[1] = l2;
		row_apply(&R->t[0], R->tA[0], R->tB[0], d1, &u);
		row_apply(&R->t[1], R->tA[1], R->tB[1], d2, &u);
	}
	if (m->np >= 2)
	{
		const float lim = maxf * R->rp;
		if (!(lim == 0.0f && m->cf[2] == 0.0f))
		{
			float l = m->cf[2] + (R->w.em * twist_jv(m->n, &u));
			l = fminf(fmaxf(l, -lim), lim);
			const float d = l - m->cf[2];
			m->cf[2] = l;
			twist_apply(&R->w, d, &u);
		}
	}
	for (int i = 0; i < 4; i++)
	{
		const int k = (it & 1u) ? 3 - i : i;
		float lambda = R->n[k].em * (row_jv(&R->n[k], m->n, &u) - R->bias[k]);
		float nt = fmaxf(0.0f, m->ln[k] + lambda);
		lambda = nt - m->ln[k];
		m->ln[k] = nt;
		row_apply(&R->n[k], R->nA, R->nB, lambda, &u);
	}
	store_vel(A, B, &u);
}

static void solve_position(orc_world *w, manifold_t *m)
{
	body_t *A = &w->bodies[m->a];
	body_t *B = m->b < ORC_STATIC_BODY_BASE ? &w->bodies[m->b] : NULL;
	body_world_inertia(A);
	if (B) body_world_inertia(B);
	const float imb = B ? B->im : 0.0f;
	const float *MB = B ? B->M : ZERO_M;
	for (int k = 0; k < m->np; k++)
	{
		v3 p1 = vadd(A->x, qrot(A->q, m->p1l[k]));
		v3 p2 = B ? vadd(B->x, qrot(B->q, m->p2l[k])) : m->p2l[k];
		float sep = vdot(vsub(p2, p1), m->n) + PENETRATION_SLOP;
		if (sep >= 0.0f) continue;
		v3 mid = vscale(vadd(p1, p2), 0.5f);
		v3 r1 = vsub(mid, A->x), r2 = B ? vsub(mid, B->x) : V(0, 0, 0);
		float e = eff_mass(A->im, A->M, imb, MB, r1, r2, m->n);
		float c = fmaxf(sep, -MAX_PENETRATION_DISTANCE);
		float lambda = (-e * BAUMGARTE) * c;
		v3 P = vscale(m->n, lambda);
		if (is_dyn(A))
		{
			A->x = vsub(A->x, mask_lin(A->dofs, vscale(P, A->im)));
			A->q = qstep(A->q, vneg(sym_mul(A->M, vcross(r1, P))));
		}
		if (B && is_dyn(B))
		{
			B->x = vadd(B->x, mask_lin(B->dofs, vscale(P, B->im)));
			B->q = qstep(B->q, sym_mul(B->M, vcross(r2, P)));
		}
	}
}

/* Contact events of the tick (the ContactListener / CharacterContactListener analogue, PlayerPhysics.c:89-152): the set
 * of touching pairs after the last sub-step — solver manifolds and sensor overlaps — against the set of the previous
 * tick.  Canonical order: added/persisted pairs sorted by (a, b), then removed pairs sorted by (a, b). */
static int cmp_u64(const void *x, const void *y)
{
	uint64_t a = *(const uint64_t *)x, b = *(const uint64_t *)y;
	return a < b ? -1 : (a > b ? 1 : 0);
}

static void make_events(orc_world *w)
{
	uint32_t n = 0;
	for (uint32_t i = 0; i < w->nprev; i++) w->ev_cur[n++] = ((uint64_t)w->prev[i].a << 32) | w->prev[i].b;
	for (uint32_t i = 0; i < w->nsens; i++) w->ev_cur[n++] = w->sens[i];
	for (uint32_t i = 0; i < w->ch_nkeys; i++) w->ev_cur[n++] = w->ch_keys[i];
	qsort(w->ev_cur, n, sizeof(uint64_t), cmp_u64);
	uint32_t m = 0;
	for (uint32_t i = 0; i < n; i++)
		if (m == 0 || w->ev_cur[m - 1] != w->ev_cur[i]) w->ev_cur[m++] = w->ev_cur[i];
	w->nev_cur = m;
	uint32_t e = 0, i = 0, j = 0;
	/* added / persisted */
	for (i = 0; i < w->nev_cur; i++)
	{
		while (j < w->nev_prev && w->ev_prev[j] < w->ev_cur[i]) j++;
		int persisted = j < w->nev_prev && w->ev_prev[j] == w->ev_cur[i];
		w->events[3 * e + 0] = (uint32_t)(w->ev_cur[i] >> 32);
		w->events[3 * e + 1] = (uint32_t)(w->ev_cur[i] & 0xFFFFFFFFu);
		w->events[3 * e + 2] = persisted ? 2u : 1u;
		e++;
	}
	/* removed */
	i = 0;
	for (j = 0; j < w->nev_prev; j++)
	{
		while (i < w->nev_cur && w->ev_cur[i] < w->ev_prev[j]) i++;
		if (i < w->nev_cur && w->ev_cur[i] == w->ev_prev[j]) continue;
		w->events[3 * e + 0] = (uint32_t)(w->ev_prev[j] >> 32);
		w->events[3 * e + 1] = (uint32_t)(w->ev_prev[j] & 0xFFFFFFFFu);
		w->events[3 * e + 2] = 3u;
		e++;
	}
	w->nevents = e;
	uint64_t *t = w->ev_prev;
	w->ev_prev = w->ev_cur;
	w->ev_cur = t;
	w->nev_prev = w->nev_cur;
}

static v3 clamp_len(v3 v, float maxl)
{
	float l2 = vlen2(v);
	if (l2 > (maxl * maxl)) return vscale(v, maxl / sqrtf(l2));
	return v;
}


/* ------------------------------------------------------------------------------------------ sleeping
 * Jolt's sleep test [upstream, restated from its documented behaviour; SURVEY §8 row a2: "0.03 m/s for 0.5 s"]: three
 * test points per body (centre of mass and the box extents along its two larger local axes) are each kept inside a
 * growing sphere; a sphere radius above 0.03 m/s * 0.5 s = 15 mm restarts the test, 0.5 s without a restart makes the
 * body a sleep candidate, and an island (bodies connected by contacts) goes to sleep when all its bodies are candidates:
 * velocities are zeroed and the bodies behave as static until an active body touches them or the host wakes them.
 * Evaluated once per tick (Jolt: once per collision step), on the contacts of the tick's last sub-step.  Sensors and
 * bodies created with allow_sleeping = 0 never sleep. */
#define SLEEP_POINT_VELOCITY 0.03f
#define SLEEP_TIME 0.5f

static void sleep_points(const body_t *b, v3 *pts)
{
	const v3 e = b->shape == ORC_SHAPE_SPHERE ? V(b->he.x, b->he.x, b->he.x) : b->he;
	const int lowest = e.x < e.y ? (e.z < e.x ? 2 : 0) : (e.z < e.y ? 2 : 1);
	const v3 ax = qrot(b->q, V(1.0f, 0.0f, 0.0f)), ay = qrot(b->q, V(0.0f, 1.0f, 0.0f)), az = qrot(b->q, V(0.0f, 0.0f, 1.0f));
	pts[0] = b->x;
	if (lowest == 0)
	{
		pts[1] = vmadd(b->x, ay, e.y);
		pts[2] = vmadd(b->x, az, e.z);
	}
	else if (lowest == 1)
	{
		pts[1] = vmadd(b->x, ax, e.x);
		pts[2] = vmadd(b->x, az, e.z);
	}
	else
	{
		pts[1] = vmadd(b->x, ax, e.x);
		pts[2] = vmadd(b->x, ay, e.y);
	}
}

static int sleep_candidate(body_t *b, float dt)
{
	if (!b->allow_sleep || b->sensor)
	{
		b->sleep_t = -1.0f;
		return 0;
	}
	v3 pts[3];
	sleep_points(b, pts);
	int restart = b->sleep_t < 0.0f;
	for (int i = 0; i < 3 && !restart; i++)
	{
		/* grow the sphere just enough to hold the point */
		const v3 d = vsub(pts[i], b->sleep_c[i]);
		const float d2 = vlen2(d), r = b->sleep_r[i];
		if (d2 > (r * r))
		{
			const float dist = sqrtf(d2), nr = 0.5f * (r + dist);
			b->sleep_c[i] = vmadd(b->sleep_c[i], d, (nr - r) / dist);
			b->sleep_r[i] = nr;
		}
		if (b->sleep_r[i] > (SLEEP_POINT_VELOCITY * SLEEP_TIME)) restart = 1;
	}
	if (restart)
	{
		for (int i = 0; i < 3; i++)
		{
			b->sleep_c[i] = pts[i];
			b->sleep_r[i] = 0.0f;
		}
		b->sleep_t = 0.0f;
		return 0;
	}
	b->sleep_t += dt;
	return b->sleep_t >= SLEEP_TIME;
}

static uint32_t uf_find(uint32_t *parent, uint32_t x)
{
	while (parent[x] != x)
	{
		parent[x] = parent[parent[x]];
		x = parent[x];
	}
	return x;
}

static void sleep_pass(orc_world *w, float dt)
{
	const uint32_t nb = w->max_bodies;
	uint32_t *parent = (uint32_t *)malloc(sizeof(uint32_t) * nb);
	unsigned char *can = (unsigned char *)malloc(nb);
	for (uint32_t i = 0; i < nb; i++)
	{
		parent[i] = i;
		can[i] = 1;
	}
	/* islands: awake dynamic bodies joined by the contacts of the last sub-step; the smaller index becomes the root */
	for (uint32_t k = 0; k < w->nprev; k++)
	{
		const manifold_t *m = &w->prev[k];
		if (m->b >= ORC_STATIC_BODY_BASE || !is_dyn(&w->bodies[m->a]) || !is_dyn(&w->bodies[m->b])) continue;
		uint32_t ra = uf_find(parent, m->a), rb = uf_find(parent, m->b);
		if (ra == rb) continue;
		if (ra < rb) parent[rb] = ra;
		else parent[ra] = rb;
	}
	for (uint32_t i = 0; i < nb; i++)
	{
		body_t *b = &w->bodies[i];
		if (!b->alive || !is_dyn(b)) continue;
		if (!sleep_candidate(b, dt)) can[uf_find(parent, i)] = 0;
	}
	for (uint32_t i = 0; i < nb; i++)
	{
		body_t *b = &w->bodies[i];
		if (!b->alive || !is_dyn(b) || !can[uf_find(parent, i)]) continue;
		b->asleep = 1;
		b->v = V(0.0f, 0.0f, 0.0f);
		b->w = V(0.0f, 0.0f, 0.0f);
	}
	free(parent);
	free(can);
}

/* gravity, damping, velocity clamps and the world-space inverse inertia of bodies [lo, hi) */
static void apply_forces(orc_world *w, float h, uint32_t lo, uint32_t hi)
{
	for (uint32_t i = lo; i < hi; i++)
	{
		body_t *b = &w->bodies[i];
		if (!b->alive) continue;
		if (is_dyn(b))
		{
			b->v = vadd(b->v, vscale(w->gravity, h * b->grav_factor));
			b->v = vscale(b->v, fmaxf(0.0f, 1.0f - (b->lin_damp * h)));
			b->w = vscale(b->w, fmaxf(0.0f, 1.0f - (b->ang_damp * h)));
			b->v = clamp_len(mask_lin(b->dofs, b->v), MAX_LINEAR_VELOCITY);
			v3 ww = b->w;
			if (!(b->dofs & 8u)) ww.x = 0.0f;
			if (!(b->dofs & 16u)) ww.y = 0.0f;
			if (!(b->dofs & 32u)) ww.z = 0.0f;
			b->w = clamp_len(ww, MAX_ANGULAR_VELOCITY);
		}
		body_world_inertia(b);
	}
}

static void integrate_positions(orc_world *w, float h, uint32_t lo, uint32_t hi)
{
	for (uint32_t i = lo; i < hi; i++)
	{
		body_t *b = &w->bodies[i];
		if (!b->alive || b->motion == ORC_MOTION_STATIC || b->asleep) continue;
		b->x = vadd(b->x, vscale(b->v, h));
		b->q = qstep(b->q, vscale(b->w, h));
	}
}

int orc_step(orc_world *w, float dt, int collision_steps)
{
	int err = 0;
	if (collision_steps < 1) collision_steps = 1;
	const float h = dt / (float)collision_steps;
	for (int s = 0; s < collision_steps; s++)
	{
		apply_forces(w, h, 0, w->max_bodies);
		find_contacts(w, &err);
		warm_start_match(w);
		if (w->mode == 1) colour_manifolds_jp(w);
		else colour_manifolds(w);
		for (uint32_t k = 0; k < w->nman; k++) setup_manifold(w, &w->man[k], h);
		for (uint32_t k = 0; k < w->nman; k++) warm_start(w, &w->man[w->order[k]]);
		for (uint32_t it = 0; it < w->vel_steps; it++)
			for (uint32_t k = 0; k < w->nman; k++) solve_velocity(w, &w->man[w->order[k]], it);
		integrate_positions(w, h, 0, w->max_bodies);
		for (uint32_t it = 0; it < w->pos_steps; it++)
			for (uint32_t k = 0; k < w->nman; k++) solve_position(w, &w->man[w->order[k]]);
		manifold_t *t = w->prev;
		w->prev = w->man;
		w->man = t;
		w->nprev = w->nman;
		/* sleepers touched by something active in this sub-step are awake from the next one on */
		for (uint32_t i = 0; i < w->max_bodies; i++)
			if (w->bodies[i].alive && w->bodies[i].wake_mark) wake_body(&w->bodies[i]);
	}
	sleep_pass(w, dt);
	make_events(w);
	return err;
}

/* ---- orc_step over host threads: the CPU baseline bench.py quotes for ONE large world (mode 1).  Same stages, same
 * arithmetic; what runs side by side is what cannot interact: bodies, the contacts of disjoint ranges of first bodies
 * (gathered in range order, so the manifold list is the one orc_step builds), manifolds of one colour (they share no
 * dynamic body).  tests/test_oracle.py holds it to orc_step's bits. */
typedef struct
{
	orc_world *w;
	float h;
	uint32_t it;
	int stage;
	uint32_t lo, hi;            /* slice of `order` for the colour stages */
	const uint64_t *cand;
	const uint32_t *cand_first;
	orc_world *shadow;          /* per slot: a copy of the world header with its own manifold / sensor buffers */
	int *slot_err;
	int nslots;
	const uint32_t *run;        /* warm-start match: first old manifold per body */
} mtjob_t;
enum { ST_FORCES, ST_CONTACTS, ST_MATCH, ST_SETUP, ST_WARM, ST_VEL, ST_INTEGRATE, ST_POS };

static void mt_range(void *ctx, int64_t lo, int64_t hi)
{
	mtjob_t *j = (mtjob_t *)ctx;
	orc_world *w = j->w;
	switch (j->stage)
	{
	case ST_FORCES: apply_forces(w, j->h, (uint32_t)lo, (uint32_t)hi); break;
	case ST_INTEGRATE: integrate_positions(w, j->h, (uint32_t)lo, (uint32_t)hi); break;
	case ST_CONTACTS:
		/* [lo, hi) are slots: slot k owns the bodies [k, k + 1) * max_bodies / nslots */
		for (int64_t k = lo; k < hi; k++)
		{
			orc_world *t = &j->shadow[k];
			t->nman = 0;
			t->nsens = 0;
			j->slot_err[k] = 0;
			find_contacts_among(t, &j->slot_err[k], j->cand, j->cand_first, (uint32_t)((uint64_t)w->max_bodies * k / j->nslots),
								(uint32_t)((uint64_t)w->max_bodies * (k + 1) / j->nslots));
		}
		break;
	case ST_MATCH: warm_start_match_range(w, j->run, (uint32_t)lo, (uint32_t)hi); break;
	case ST_SETUP:
		for (int64_t k = lo; k < hi; k++) setup_manifold(w, &w->man[k], j->h);
		break;
	case ST_WARM:
		for (int64_t k = lo; k < hi; k++) warm_start(w, &w->man[w->order[j->lo + k]]);
		break;
	case ST_VEL:
		for (int64_t k = lo; k < hi; k++) solve_velocity(w, &w->man[w->order[j->lo + k]], j->it);
		break;
	case ST_POS:
		for (int64_t k = lo; k < hi; k++) solve_position(w, &w->man[w->order[j->lo + k]]);
		break;
	}
}

/* a persistent pool: parallel_for's thread-per-call start-up would dominate the hundred colour phases of a tick */
typedef struct
{
	pthread_t th[256];
	int nt, started;
	pthread_mutex_t mu;
	pthread_cond_t go, done;
	uint64_t gen;
	int pending;
	void (*fn)(void *, int64_t, int64_t);
	void *ctx;
	int64_t n;
} pool_t;
static pool_t g_pool = {.mu = PTHREAD_MUTEX_INITIALIZER, .go = PTHREAD_COND_INITIALIZER, .done = PTHREAD_COND_INITIALIZER};

static void *pool_main(void *arg)
{
	const int t = (int)(intptr_t)arg;
	uint64_t seen = 0;
	for (;;)
	{
		pthread_mutex_lock(&g_pool.mu);
		while (g_pool.gen == seen) pthread_cond_wait(&g_pool.go, &g_pool.mu);
		seen = g_pool.gen;
		void (*fn)(void *, int64_t, int64_t) = g_pool.fn;
		void *ctx = g_pool.ctx;
		const int64_t n = g_pool.n;
		pthread_mutex_unlock(&g_pool.mu);
		const int64_t lo = n * t / g_pool.nt, hi = n * (t + 1) / g_pool.nt;
		if (hi > lo) fn(ctx, lo, hi);
		pthread_mutex_lock(&g_pool.mu);
		if (--g_pool.pending == 0) pthread_cond_signal(&g_pool.done);
		pthread_mutex_unlock(&g_pool.mu);
	}
	return NULL;
}

/* fn over [0, n) in one slice per thread; fewer than `grain` items are not worth a wake-up */
static void pool_for(int64_t n, void (*fn)(void *, int64_t, int64_t), void *ctx, int64_t grain)
{
	if (!g_pool.started)
	{
		g_pool.nt = orc_max_threads();
		if (g_pool.nt > 256) g_pool.nt = 256;
		for (int t = 1; t < g_pool.nt; t++) pthread_create(&g_pool.th[t], NULL, pool_main, (void *)(intptr_t)t);
		g_pool.started = 1;
	}
	if (n < grain || g_pool.nt == 1)
	{
		if (n > 0) fn(ctx, 0, n);
		return;
	}
	pthread_mutex_lock(&g_pool.mu);
	g_pool.fn = fn;
	g_pool.ctx = ctx;
	g_pool.n = n;
	g_pool.pending = g_pool.nt - 1;
	g_pool.gen++;
	pthread_cond_broadcast(&g_pool.go);
	pthread_mutex_unlock(&g_pool.mu);
	const int64_t hi = n / g_pool.nt;  /* slice 0 is the caller's */
	if (hi > 0) fn(ctx, 0, hi);
	pthread_mutex_lock(&g_pool.mu);
	while (g_pool.pending) pthread_cond_wait(&g_pool.done, &g_pool.mu);
	pthread_mutex_unlock(&g_pool.mu);
}

int orc_step_mt(orc_world *w, float dt, int collision_steps)
{
	if (w->mode != 1) return orc_step(w, dt, collision_steps);
	int err = 0;
	if (collision_steps < 1) collision_steps = 1;
	const float h = dt / (float)collision_steps;
	const int nslots = 4 * orc_max_threads() < 1024 ? 4 * orc_max_threads() : 1024;
	orc_world *shadow = (orc_world *)malloc(sizeof(orc_world) * nslots);
	const uint32_t slot_cap = 4u * (w->max_manifolds / (uint32_t)nslots) + 1024u;
	int *slot_err = (int *)calloc((size_t)nslots, sizeof(int));
	for (int k = 0; k < nslots; k++)
	{
		shadow[k] = *w;
		shadow[k].max_manifolds = slot_cap;
		shadow[k].man = (manifold_t *)malloc(sizeof(manifold_t) * slot_cap);
		shadow[k].sens = (uint64_t *)malloc(sizeof(uint64_t) * slot_cap);
	}
	mtjob_t j;
	memset(&j, 0, sizeof(j));
	j.w = w;
	j.h = h;
	j.shadow = shadow;
	j.slot_err = slot_err;
	j.nslots = nslots;
	for (int s = 0; s < collision_steps; s++)
	{
		j.stage = ST_FORCES;
		pool_for(w->max_bodies, mt_range, &j, 2048);
		/* contacts: the candidate sweep in slices of sorted positions, the narrowphase per slot of first bodies */
		uint32_t *cand_first = NULL;
		uint64_t *cand = w->max_bodies > SWEEP_MIN_BODIES ? sweep_candidates(w, &cand_first, 1) : NULL;
		for (int k = 0; k < nslots; k++)
		{
			shadow[k].bodies = w->bodies;
			shadow[k].prev = w->prev;
			shadow[k].nprev = w->nprev;
		}
		j.cand = cand;
		j.cand_first = cand_first;
		j.stage = ST_CONTACTS;
		pool_for(nslots, mt_range, &j, 2);
		free(cand);
		free(cand_first);
		w->nman = 0;
		w->nsens = 0;
		int crowded = 0;
		for (int k = 0; k < nslots; k++) crowded |= slot_err[k] != 0;
		for (int k = 0; k < nslots && !crowded; k++)
		{
			if (w->nman + shadow[k].nman > w->max_manifolds)
			{
				err = 4;
				break;
			}
			memcpy(&w->man[w->nman], shadow[k].man, sizeof(manifold_t) * shadow[k].nman);
			w->nman += shadow[k].nman;
			for (uint32_t q = 0; q < shadow[k].nsens && w->nsens < w->max_manifolds; q++) w->sens[w->nsens++] = shadow[k].sens[q];
		}
		if (crowded) find_contacts(w, &err);  /* a slot's share of the manifold buffer was too small: the serial search */
		/* warm-start match over manifolds */
		uint32_t *run = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)w->max_bodies + 1));
		for (uint32_t a = 0, q = 0; a <= w->max_bodies; a++)
		{
			while (q < w->nprev && w->prev[q].a < a) q++;
			run[a] = q;
		}
		j.run = run;
		j.stage = ST_MATCH;
		pool_for(w->nman, mt_range, &j, 2048);
		free(run);
		colour_manifolds_jp(w);
		j.stage = ST_SETUP;
		pool_for(w->nman, mt_range, &j, 2048);
		/* colour ranges of `order` (colour-major); the last colour may hold manifolds that share a body: serial */
		uint32_t start[66];
		int ncol = 0;
		for (uint32_t k = 0; k < w->nman; k++)
		{
			const int c = w->man[w->order[k]].colour;
			while (ncol <= c) start[ncol++] = k;
		}
		start[ncol] = w->nman;
		for (int pass = 0; pass < 2 + (int)w->vel_steps + (int)w->pos_steps; pass++)
		{
			if (pass == 1 + (int)w->vel_steps)
			{
				j.stage = ST_INTEGRATE;
				pool_for(w->max_bodies, mt_range, &j, 2048);
				continue;
			}
			j.stage = pass == 0 ? ST_WARM : (pass <= (int)w->vel_steps ? ST_VEL : ST_POS);
			j.it = pass >= 1 && pass <= (int)w->vel_steps ? (uint32_t)(pass - 1) : 0u;
			for (int c = 0; c < ncol; c++)
			{
				j.lo = start[c];
				j.hi = start[c + 1];
				if (c == 63)
					mt_range(&j, 0, (int64_t)(j.hi - j.lo));
				else
					pool_for((int64_t)(j.hi - j.lo), mt_range, &j, 1024);
			}
		}
		manifold_t *t = w->prev;
		w->prev = w->man;
		w->man = t;
		w->nprev = w->nman;
		for (uint32_t i = 0; i < w->max_bodies; i++)
			if (w->bodies[i].alive && w->bodies[i].wake_mark) wake_body(&w->bodies[i]);
	}
	for (int k = 0; k < nslots; k++)
	{
		free(shadow[k].man);
		free(shadow[k].sens);
	}
	free(shadow);
	free(slot_err);
	sleep_pass(w, dt);
	make_events(w);
	return err;
}

typedef struct { orc_world **ws; float dt; int steps, ticks; int err; } stepjob_t;
static void step_range(void *ctx, int64_t lo, int64_t hi)
{
	stepjob_t *j = (stepjob_t *)ctx;
	int err = 0;
	for (int64_t i = lo; i < hi; i++)
		for (int t = 0; t < j->ticks; t++) err |= orc_step(j->ws[i], j->dt, j->steps);
	if (err) __atomic_fetch_or(&j->err, err, __ATOMIC_RELAXED);
}

int orc_step_many(orc_world **ws, uint32_t n, float dt, int collision_steps, int ticks)
{
	stepjob_t j = {ws, dt, collision_steps, ticks, 0};
	parallel_for((int64_t)n, step_range, &j);
	return j.err;
}


/* ------------------------------------------------------------------------------------------ player character
 *
 * The engine's player is a JPH_CharacterVirtual: a capsule (half height 0.2, radius 0.25) that is not a body of the
 * physics system; every tick MovePlayer sets its velocity and UpdatePlayer calls JPH_CharacterVirtual_ExtendedUpdate
 * BEFORE the physics update (engine/src/physics/PlayerPhysics.c:173-194,203-295,439-453; MapPhysics.c:66-77), and a
 * contact listener turns its contacts into actor callbacks (PlayerPhysics.c:89-152).  Jolt's CharacterVirtual is not
 * available here (PARITY UNPINNED, see orc.h); this restates the behaviour the engine relies on with a discrete
 * collide-and-slide:
 *   move by v * dt; up to 8 times: find the deepest penetration of the capsule against the static triangles and the
 *   solid bodies, push the capsule out along that contact normal and remove the velocity component into it;
 *   ground state from the contact normals (max slope 50 degrees) and, failing that, from a 5 cm probe below;
 *   contacts (for the callbacks): every body or static mesh within the contact margin of the capsule, sensors included.
 * The stick-to-floor step of ExtendedUpdate is restated as closing gaps of up to 5 cm over walkable ground; stair
 * stepping is not. */

#define CH_MAX_ITERS 8
#define CH_CONTACT_MARGIN 0.02f
#define CH_GROUND_PROBE 0.05f
#define CH_ID 0x3FFFFFu

static float clamp01(float t) { return t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t); }

/* closest points of segments p1-q1 and p2-q2 (Ericson, Real-Time Collision Detection 5.1.9) */
static void seg_seg(v3 p1, v3 q1, v3 p2, v3 q2, v3 *c1, v3 *c2)
{
	v3 d1 = vsub(q1, p1), d2 = vsub(q2, p2), r = vsub(p1, p2);
	float a = vdot(d1, d1), e = vdot(d2, d2), f = vdot(d2, r);
	float s, t;
	const float EPS = 1.0e-12f;
	if (a <= EPS && e <= EPS)
	{
		s = t = 0.0f;
	}
	else if (a <= EPS)
	{
		s = 0.0f;
		t = clamp01(f / e);
	}
	else
	{
		float c = vdot(d1, r);
		if (e <= EPS)
		{
			t = 0.0f;
			s = clamp01(-c / a);
		}
		else
		{
			float b = vdot(d1, d2);
			float denom = (a * e) - (b * b);
			s = denom != 0.0f ? clamp01(((b * f) - (c * e)) / denom) : 0.0f;
			t = ((b * s) + f) / e;
			if (t < 0.0f)
			{
				t = 0.0f;
				s = clamp01(-c / a);
			}
			else if (t > 1.0f)
			{
				t = 1.0f;
				s = clamp01((b - c) / a);
			}
		}
	}
	*c1 = vadd(p1, vscale(d1, s));
	*c2 = vadd(p2, vscale(d2, t));
}

/* closest points between segment p0-p1 and a triangle; returns squared distance (0 when the segment pierces it) */
static float seg_tri(v3 p0, v3 p1, v3 a, v3 b, v3 c, v3 *cs, v3 *ct)
{
	/* piercing: the segment as a ray of length 1 */
	{
		v3 d = vsub(p1, p0), e1 = vsub(b, a), e2 = vsub(c, a);
		v3 pv = vcross(d, e2);
		float det = vdot(e1, pv);
		if (fabsf(det) >= 1.0e-12f)
		{
			float inv = 1.0f / det;
			v3 tv = vsub(p0, a);
			float u = vdot(tv, pv) * inv;
			if (u >= 0.0f && u <= 1.0f)
			{
				v3 q = vcross(tv, e1);
				float v = vdot(d, q) * inv;
				if (v >= 0.0f && (u + v) <= 1.0f)
				{
					float t = vdot(e2, q) * inv;
					if (t >= 0.0f && t <= 1.0f)
					{
						*cs = *ct = vadd(p0, vscale(d, t));
						return 0.0f;
					}
				}
			}
		}
	}
	float best = 3.0e38f;
	const v3 tv[3] = {a, b, c};
	for (int i = 0; i < 3; i++)
	{
		v3 x, y;
		seg_seg(p0, p1, tv[i], tv[(i + 1) % 3], &x, &y);
		float d2 = vlen2(vsub(x, y));
		if (d2 < best) { best = d2; *cs = x; *ct = y; }
	}
	const v3 ends[2] = {p0, p1};
	for (int i = 0; i < 2; i++)
	{
		v3 y = closest_on_tri(ends[i], a, b, c);
		float d2 = vlen2(vsub(ends[i], y));
		if (d2 < best) { best = d2; *cs = ends[i]; *ct = y; }
	}
	return best;
}

/* closest points between segment p0-p1 and an oriented box; returns squared distance (0 when the segment enters it) */
static float seg_box(v3 p0, v3 p1, const body_t *B, v3 *cs, v3 *cb)
{
	m33 R = qmat(B->q);
	v3 l0 = mtmul(&R, vsub(p0, B->x)), l1 = mtmul(&R, vsub(p1, B->x));
	const v3 he = B->he;
	/* slab test of the segment against the box */
	{
		v3 d = vsub(l1, l0);
		float tn = 0.0f, tf = 1.0f;
		int hit = 1;
		for (int k = 0; k < 3 && hit; k++)
		{
			float ok = vget(l0, k), dk = vget(d, k), hk = vget(he, k);
			if (dk == 0.0f)
			{
				if (ok < -hk || ok > hk) hit = 0;
				continue;
			}
			float inv = 1.0f / dk;
			float t1 = (-hk - ok) * inv, t2 = (hk - ok) * inv;
			if (t1 > t2) { float tt = t1; t1 = t2; t2 = tt; }
			if (t1 > tn) tn = t1;
			if (t2 < tf) tf = t2;
			if (tn > tf) hit = 0;
		}
		if (hit)
		{
			v3 lp = vadd(l0, vscale(d, tn));
			*cs = *cb = vadd(B->x, mmul(&R, lp));
			return 0.0f;
		}
	}
	float best = 3.0e38f;
	v3 bs = l0, bb = l0;
	const v3 ends[2] = {l0, l1};
	for (int i = 0; i < 2; i++)
	{
		v3 l = ends[i];
		v3 q = V(fminf(fmaxf(l.x, -he.x), he.x), fminf(fmaxf(l.y, -he.y), he.y), fminf(fmaxf(l.z, -he.z), he.z));
		float d2 = vlen2(vsub(l, q));
		if (d2 < best) { best = d2; bs = l; bb = q; }
	}
	for (int k = 0; k < 3; k++)
	{
		const int u = (k + 1) % 3, v = (k + 2) % 3;
		for (int sgn = 0; sgn < 4; sgn++)
		{
			float e0[3], e1[3];
			const float su = (sgn & 1) ? 1.0f : -1.0f, sv = (sgn & 2) ? 1.0f : -1.0f;
			e0[k] = -vget(he, k); e1[k] = vget(he, k);
			e0[u] = e1[u] = su * vget(he, u);
			e0[v] = e1[v] = sv * vget(he, v);
			v3 x, y;
			seg_seg(l0, l1, V(e0[0], e0[1], e0[2]), V(e1[0], e1[1], e1[2]), &x, &y);
			float d2 = vlen2(vsub(x, y));
			if (d2 < best) { best = d2; bs = x; bb = y; }
		}
	}
	*cs = vadd(B->x, mmul(&R, bs));
	*cb = vadd(B->x, mmul(&R, bb));
	return best;
}

static float seg_point(v3 p0, v3 p1, v3 c, v3 *cs)
{
	v3 d = vsub(p1, p0);
	float a = vdot(d, d);
	float t = a > 1.0e-12f ? clamp01(vdot(vsub(c, p0), d) / a) : 0.0f;
	*cs = vadd(p0, vscale(d, t));
	return vlen2(vsub(*cs, c));
}

void orc_character_create(orc_world *w, const float pos[3], float half_height, float radius, float max_slope_deg)
{
	w->ch_alive = 1;
	w->ch_x = V(pos[0], pos[1], pos[2]);
	w->ch_v = V(0, 0, 0);
	w->ch_hh = half_height;
	w->ch_r = radius;
	w->ch_cos_slope = cosf(max_slope_deg * 0.0174532925f);
	w->ch_ground = 3;
	w->ch_ground_body = ORC_INVALID;
	w->ch_ground_n = V(0, 1, 0);
	w->ch_nkeys = 0;
}
void orc_character_destroy(orc_world *w) { w->ch_alive = 0; w->ch_nkeys = 0; }

uint32_t orc_character_contacts(const orc_world *w, uint32_t *others, uint32_t cap)
{
	uint64_t keys[64];
	uint32_t n = w->ch_nkeys;
	memcpy(keys, w->ch_keys, sizeof(uint64_t) * n);
	qsort(keys, n, sizeof(uint64_t), cmp_u64);
	uint32_t m = 0;
	for (uint32_t i = 0; i < n; i++)
		if (i == 0 || keys[i] != keys[i - 1])
		{
			const uint32_t a = (uint32_t)(keys[i] >> 32), b = (uint32_t)(keys[i] & 0xFFFFFFFFu);
			if (m < cap) others[m] = a == CH_ID ? b : a;
			m++;
		}
	return m;
}
void orc_character_set_velocity(orc_world *w, const float v[3]) { w->ch_v = V(v[0], v[1], v[2]); }
void orc_character_set_position(orc_world *w, const float p[3]) { w->ch_x = V(p[0], p[1], p[2]); }
void orc_character_get(const orc_world *w, float pos[3], float vel[3], uint32_t *ground, uint32_t *ground_body)
{
	pos[0] = w->ch_x.x; pos[1] = w->ch_x.y; pos[2] = w->ch_x.z;
	vel[0] = w->ch_v.x; vel[1] = w->ch_v.y; vel[2] = w->ch_v.z;
	*ground = w->ch_ground;
	*ground_body = w->ch_ground_body;
}

/* does the character's layer (PLAYER) collide with this body as a solid obstacle? */
static int ch_solid(const body_t *b)
{
	return b->alive && b->shape != ORC_SHAPE_EMPTY && !b->sensor && (b->layer == 0 || b->layer == 1);
}

/* deepest penetration of the capsule at x; id = triangle index or ORC_STATIC_BODY_BASE-less body id + 0x80000000 */
static float capsule_deepest(const orc_world *w, v3 x, float hh, float r, v3 *n_out, uint32_t *hit_body, v3 *cp_out)
{
	const v3 p0 = V(x.x, x.y - hh, x.z), p1 = V(x.x, x.y + hh, x.z);
	float best = 0.0f;
	uint32_t best_id = 0xFFFFFFFFu;
	for (uint32_t t = 0; t < w->ntris; t++)
	{
		const tri_t *T = &w->tris[t];
		if (p0.x - r > T->hi.x || p1.x + r < T->lo.x || p0.y - r > T->hi.y || p1.y + r < T->lo.y || p0.z - r > T->hi.z ||
			p1.z + r < T->lo.z)
			continue;
		v3 cs, ct;
		float d2 = seg_tri(p0, p1, T->v0, T->vb, T->vc, &cs, &ct);
		float dist = sqrtf(d2);
		float pen = r - dist;
		if (pen > best || (pen == best && pen > 0.0f && t < best_id))
		{
			v3 n;
			if (dist > 1.0e-6f) n = vscale(vsub(cs, ct), 1.0f / dist);
			else n = vdot(vsub(x, T->v0), T->n) >= 0.0f ? T->n : vneg(T->n);
			best = pen;
			best_id = t;
			*n_out = n;
			*hit_body = ORC_STATIC_BODY_BASE + T->body;
			*cp_out = ct;
		}
	}
	for (uint32_t i = 0; i < w->max_bodies; i++)
	{
		const body_t *B = &w->bodies[i];
		if (!ch_solid(B)) continue;
		v3 cs, cb;
		float d2, rr = r;
		if (B->shape == ORC_SHAPE_BOX) d2 = seg_box(p0, p1, B, &cs, &cb);
		else
		{
			d2 = seg_point(p0, p1, B->x, &cs);
			cb = B->x;
			rr = r + B->he.x;
		}
		float dist = sqrtf(d2);
		float pen = rr - dist;
		const uint32_t id = 0x80000000u + i;
		if (pen > best || (pen == best && pen > 0.0f && id < best_id))
		{
			v3 n;
			if (dist > 1.0e-6f) n = vscale(vsub(cs, cb), 1.0f / dist);
			else
			{
				v3 d = vsub(x, B->x);
				float l = vlen(d);
				n = l > 1.0e-6f ? vscale(d, 1.0f / l) : V(0, 1, 0);
			}
			best = pen;
			best_id = id;
			*n_out = n;
			*hit_body = i;
			*cp_out = cb;
		}
	}
	return best;
}

static float ch_deepest(const orc_world *w, v3 x, v3 *n_out, uint32_t *hit_body)
{
	v3 cp;
	return capsule_deepest(w, x, w->ch_hh, w->ch_r, n_out, hit_body, &cp);
}

/* The character pushes the dynamic bodies it runs into (CharacterVirtual's contact impulse [upstream, restated from its
 * documented behaviour]): it wants the body to move away at 0.9 of the closing speed plus 0.4 of the penetration per
 * update, through the body's effective mass at the contact point, capped by the character's strength (100 N) times
 * dt; no push along gravity.  `n` points from the body to the character, `v` is the character's velocity. */
#define CH_PUSH_DAMPING 0.9f
#define CH_PUSH_PENETRATION 0.4f
#define CH_MAX_STRENGTH 100.0f
static void ch_push_body(body_t *B, v3 n, float pen, v3 cp, v3 v, float dt)
{
	if (B->motion != ORC_MOTION_DYNAMIC || B->sensor) return;
	const v3 rB = vsub(cp, B->x);
	const v3 vB = vadd(B->v, vcross(B->w, rB));
	const float dv = (-(vdot(vsub(v, vB), n)) * CH_PUSH_DAMPING) + ((pen * CH_PUSH_PENETRATION) / dt);
	if (!(dv > 0.0f)) return;
	if (B->asleep) wake_body(B); /* AddImpulse activates a sleeping body; an awake one keeps its sleep timer */
	body_world_inertia(B);
	const v3 jac = vcross(rB, n);
	const float inv_eff = vdot(sym_mul(B->M, jac), jac) + B->im;
	if (!(inv_eff > 0.0f)) return;
	const float impulse = fminf(dv / inv_eff, CH_MAX_STRENGTH * dt);
	v3 P = vscale(n, -impulse);
	if (P.y < 0.0f) P.y = 0.0f;
	B->v = vadd(B->v, mask_lin(B->dofs, vscale(P, B->im)));
	B->w = vadd(B->w, sym_mul(B->M, vcross(rB, P)));
}

/* the same query for a caller's own upright capsule (gpx_overlap_capsule_batch): depth 0 and ORC_INVALID when free */
float orc_overlap_capsule(const orc_world *w, const float center[3], float half_height, float radius, float normal[3],
						  uint32_t *body)
{
	v3 n = V(0, 1, 0);
	uint32_t hb = ORC_INVALID;
	v3 cp;
	float pen = capsule_deepest(w, V(center[0], center[1], center[2]), half_height, radius, &n, &hb, &cp);
	if (!(pen > 0.0f))
	{
		pen = 0.0f;
		n = V(0, 1, 0);
		hb = ORC_INVALID;
	}
	normal[0] = n.x; normal[1] = n.y; normal[2] = n.z;
	*body = hb;
	return pen;
}

/* One collide-and-slide pass: up to CH_MAX_ITERS times find the deepest penetration, push the capsule out along that
 * normal and remove the velocity component into it; ground state from the contact normals.  `push`: dynamic bodies in
 * the way get the character's contact impulse.  `blocked` reports a contact too steep to walk on that faces the motion
 * direction `dir` (unit, horizontal) within the angle whose cosine is cos_fwd — what lets ExtendedUpdate try a stair
 * step. */
typedef struct { v3 x, v, ground_n; uint32_t ground, ground_body; int blocked; } slide_t;
#define CH_MAX_PIECES 16 /* pieces of one tick's move: 2 m per tick, 120 m/s */

static slide_t ch_slide(orc_world *w, v3 x, v3 v, float dt, int push, v3 dir, float cos_fwd)
{
	slide_t s;
	s.ground = 3;
	s.ground_body = ORC_INVALID;
	s.ground_n = V(0, 1, 0);
	s.blocked = 0;
	for (int it = 0; it < CH_MAX_ITERS; it++)
	{
		v3 n, cp;
		uint32_t hb;
		float pen = capsule_deepest(w, x, w->ch_hh, w->ch_r, &n, &hb, &cp);
		if (!(pen > 0.0f)) break;
		if (push && hb < ORC_STATIC_BODY_BASE) ch_push_body(&w->bodies[hb], n, pen, cp, v, dt);
		x = vadd(x, vscale(n, pen));
		float vn = vdot(v, n);
		if (vn < 0.0f) v = vsub(v, vscale(n, vn));
		if (n.y >= w->ch_cos_slope)
		{
			s.ground = 0;
			s.ground_body = hb;
			s.ground_n = n;
		}
		else
		{
			if (n.y > 0.0f && s.ground != 0)
			{
				s.ground = 1;
				s.ground_body = hb;
				s.ground_n = n;
			}
			/* too steep to walk on: does it face the motion? */
			const float hl = sqrtf((n.x * n.x) + (n.z * n.z));
			if (hl > 1.0e-6f && (-((n.x * dir.x) + (n.z * dir.z))) >= (cos_fwd * hl)) s.blocked = 1;
		}
	}
	s.x = x;
	s.v = v;
	return s;
}

/* Floor below x within `reach`, by probes 5 cm apart: the first probe depth at which the capsule touches something that
 * faces up.  Returns that depth (0: nothing), the penetration found there, its normal and body. */
static float ch_probe_down(const orc_world *w, v3 x, float reach, float *pen_out, v3 *n_out, uint32_t *hb_out)
{
	for (float s = CH_GROUND_PROBE; s <= reach + 1.0e-6f; s += CH_GROUND_PROBE)
	{
		v3 n;
		uint32_t hb;
		float pen = ch_deepest(w, V(x.x, x.y - s, x.z), &n, &hb);
		if (pen > 0.0f && n.y > 0.0f)
		{
			*pen_out = pen;
			*n_out = n;
			*hb_out = hb;
			return s;
		}
	}
	return 0.0f;
}

void orc_character_update(orc_world *w, float dt)
{
	const orc_character_settings none = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
	orc_character_update_ex(w, dt, &none);
}

/* JPH_CharacterVirtual_ExtendedUpdate as the engine calls it (PlayerPhysics.c:439-453), restated on the discrete
 * collide-and-slide: the move, then stick-to-floor (a character that stood on walkable ground and is now in the air without
 * moving up is set down on a floor found within stick_to_floor_step_down), then walk-stairs (a character on the ground
 * whose horizontal move was cut short by something too steep tries the same move lifted by walk_stairs_step_up and, if
 * that makes headway and there is walkable floor within the step height below, stands there). */
void orc_character_update_ex(orc_world *w, float dt, const orc_character_settings *cfg)
{
	if (!w->ch_alive) return;
	const v3 x_old = w->ch_x;
	const int was_on_ground = w->ch_ground == 0 || w->ch_ground == 1; /* Jolt's IsSupported(): on ground or on steep ground */
	/* what the host asked for, horizontally */
	const v3 want = V(w->ch_v.x * dt, 0.0f, w->ch_v.z * dt);
	const float want_len = sqrtf((want.x * want.x) + (want.z * want.z));
	const v3 dir = want_len > 0.0f ? V(want.x / want_len, 0.0f, want.z / want_len) : V(0, 0, 0);
	/* swept motion, restated on the discrete test: the move in pieces no longer than half the capsule's radius, each
	 * collided and slid before the next, so that no floor or wall is stepped over (one piece up to 7.5 m/s) */
	const float travel = vlen(vscale(w->ch_v, dt));
	int pieces = (int)ceilf(travel / (0.5f * w->ch_r));
	pieces = pieces < 1 ? 1 : (pieces > CH_MAX_PIECES ? CH_MAX_PIECES : pieces);
	const float pdt = dt / (float)pieces;
	slide_t m = ch_slide(w, vadd(w->ch_x, vscale(w->ch_v, pdt)), w->ch_v, pdt, 1, dir, cfg->walk_stairs_cos_angle_forward_contact);
	for (int k = 1; k < pieces; k++)
	{
		const int blocked = m.blocked;
		m = ch_slide(w, vadd(m.x, vscale(m.v, pdt)), m.v, pdt, 1, dir, cfg->walk_stairs_cos_angle_forward_contact);
		m.blocked |= blocked;
	}
	v3 x = m.x, v = m.v;
	uint32_t ground = m.ground, ground_body = m.ground_body;
	v3 ground_n = m.ground_n;
	if (ground == 3)
	{
		/* ground within 5 cm counts as ground; beyond that only stick-to-floor reaches, and only from a standing start */
		const float reach = (was_on_ground && v.y <= 0.0f && cfg->stick_to_floor_step_down > CH_GROUND_PROBE)
								? cfg->stick_to_floor_step_down : CH_GROUND_PROBE;
		v3 n;
		uint32_t hb;
		float pen;
		const float s = ch_probe_down(w, x, reach, &pen, &n, &hb);
		if (s > 0.0f)
		{
			ground = n.y >= w->ch_cos_slope ? 0u : 1u;
			ground_body = hb;
			ground_n = n;
			/* stick to the floor (the stickToFloorStepDown of ExtendedUpdate, PlayerPhysics.c:439-446): close the gap
			 * when standing on walkable ground and not moving up */
			if (ground == 0u && v.y <= 0.0f) x.y = x.y - fmaxf(0.0f, s - pen);
		}
	}
	if (cfg->walk_stairs_step_up > 0.0f && want_len > 0.0f && (ground == 0u || ground == 1u || was_on_ground) && m.blocked)
	{
		const v3 got = vsub(x, x_old);
		const float got_len = fmaxf(0.0f, (got.x * dir.x) + (got.z * dir.z));
		if ((got_len + 1.0e-4f) < want_len)
		{
			const float fwd = fmaxf(cfg->walk_stairs_min_step_forward, want_len - got_len);
			const v3 up = V(x.x, x.y + cfg->walk_stairs_step_up, x.z);
			v3 n;
			uint32_t hb;
			if (!(ch_deepest(w, up, &n, &hb) > 0.0f)) /* head room */
			{
				/* the lifted move forward, swept like the move itself */
				int fp = (int)ceilf(fwd / (0.5f * w->ch_r));
				fp = fp < 1 ? 1 : (fp > CH_MAX_PIECES ? CH_MAX_PIECES : fp);
				const float step = fwd / (float)fp;
				slide_t f = ch_slide(w, V(up.x + (dir.x * step), up.y, up.z + (dir.z * step)), v, dt, 0, dir, 2.0f);
				for (int k = 1; k < fp; k++)
					f = ch_slide(w, V(f.x.x + (dir.x * step), f.x.y, f.x.z + (dir.z * step)), f.v, dt, 0, dir, 2.0f);
				const v3 adv = vsub(f.x, up);
				/* headway, and on the level: a push-out that lifted the capsule means the step is higher than step_up */
				if (((adv.x * dir.x) + (adv.z * dir.z)) > 1.0e-4f && fabsf(adv.y) <= 1.0e-3f)
				{
					float pen;
					const float s = ch_probe_down(w, f.x, cfg->walk_stairs_step_up + CH_GROUND_PROBE, &pen, &n, &hb);
					int ok = s > 0.0f && n.y >= w->ch_cos_slope;
					if (s > 0.0f && !ok && cfg->walk_stairs_step_forward_test > 0.0f)
					{
						/* landed on the edge of the step: is there walkable floor a little further on? */
						const float t = cfg->walk_stairs_step_forward_test;
						v3 n2;
						uint32_t hb2;
						float pen2;
						const v3 ahead = V(up.x + (dir.x * t), up.y, up.z + (dir.z * t));
						/* (the lifted capsule must fit there: a step higher than step_up is in the way) */
						if (!(ch_deepest(w, ahead, &n2, &hb2) > 0.0f))
						{
							const float s2 = ch_probe_down(w, ahead, cfg->walk_stairs_step_up + CH_GROUND_PROBE, &pen2, &n2, &hb2);
							ok = s2 > 0.0f && n2.y >= w->ch_cos_slope;
						}
					}
					if (ok)
					{
						x = V(f.x.x, f.x.y - fmaxf(0.0f, s - pen), f.x.z);
						ground = n.y >= w->ch_cos_slope ? 0u : 1u;
						ground_body = hb;
						ground_n = n;
					}
				}
			}
		}
	}
	w->ch_x = x;
	w->ch_v = v;
	w->ch_ground = ground;
	w->ch_ground_body = ground_body;
	w->ch_ground_n = ground_n;
	/* contacts for the callbacks: bodies (sensors included) and static meshes within the contact margin */
	const v3 p0 = V(x.x, x.y - w->ch_hh, x.z), p1 = V(x.x, x.y + w->ch_hh, x.z);
	const float reach = w->ch_r + CH_CONTACT_MARGIN;
	w->ch_nkeys = 0;
	for (uint32_t i = 0; i < w->max_bodies && w->ch_nkeys < 64; i++)
	{
		const body_t *B = &w->bodies[i];
		if (!B->alive || B->shape == ORC_SHAPE_EMPTY || !(B->layer == 0 || B->layer == 1 || B->layer == 3)) continue;
		v3 cs, cb;
		float d2, rr = reach;
		if (B->shape == ORC_SHAPE_BOX) d2 = seg_box(p0, p1, B, &cs, &cb);
		else
		{
			d2 = seg_point(p0, p1, B->x, &cs);
			rr = reach + B->he.x;
		}
		if (d2 <= rr * rr) w->ch_keys[w->ch_nkeys++] = ((uint64_t)i << 32) | CH_ID;
	}
	uint32_t last_sb = ORC_INVALID;
	for (uint32_t t = 0; t < w->ntris && w->ch_nkeys < 64; t++)
	{
		const tri_t *T = &w->tris[t];
		if (T->body == last_sb) continue; /* one contact per static mesh; triangles of a mesh are contiguous */
		if (p0.x - reach > T->hi.x || p1.x + reach < T->lo.x || p0.y - reach > T->hi.y || p1.y + reach < T->lo.y ||
			p0.z - reach > T->hi.z || p1.z + reach < T->lo.z)
			continue;
		v3 cs, ct;
		float d2 = seg_tri(p0, p1, T->v0, T->vb, T->vc, &cs, &ct);
		if (d2 <= reach * reach)
		{
			w->ch_keys[w->ch_nkeys++] = ((uint64_t)CH_ID << 32) | (ORC_STATIC_BODY_BASE + T->body);
			last_sb = T->body;
		}
	}
}
