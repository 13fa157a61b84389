// gpx_solver.cuh — the device-side building blocks of the tick, shared by the ensemble kernel (gpx_tick.cu: one tile
// of lanes per small world, state in shared memory) and the wide-world kernels (gpx_wide.cu: one thread per body /
// pair / manifold, state in global memory).  Every function works through generic pointers, so the SAME fp32
// expressions run in both layouts and both stay bit-comparable with the CPU restatement of the tick
// (JPH_PhysicsSystem_Update as called at engine/src/physics/MapPhysics.c:105-108).
#pragma once
#include "gpx_internal.h"
#include "gpx_narrow.cuh"

namespace gpx {

struct SBody  // 39 words: odd stride, conflict-free when lane = body
{
	v3 x;
	q4 q;
	v3 v, w;
	float im;    // inverse mass as seen by the solver (0 unless dynamic)
	float M[6];  // world inverse inertia xx xy xz yy yz zz
	v3 he;
	float friction, restitution, lin_damp, ang_damp, grav;
	uint32_t flags;
	v3 inv_i;
	v3 lo, hi;
	float inv_mass;  // as stored in the body store
};

struct SMan  // 49 words (odd stride): what outlives one manifold's register-resident solve
{
	uint32_t a, b;
	int np;      // 0 = empty slot (pair that did not touch)
	int colour;
	v3 n;
	float friction, restitution;
	v3 p1l[4], p2l[4];
	float ln[4], lt1[4], lt2[4];
	float bias[4];
};

// One manifold's constraint rows while its lane iterates (phase 7).  The per-manifold part lives in registers; the
// per-point part (lever arms, effective masses) in a small per-lane record in shared memory so that the four points
// run through ONE rolled loop body — fully unrolled, the velocity iteration alone was ~18 KB of SASS and the warps
// spent a third of their time waiting for instruction fetch.  Accumulated impulses and biases are read and written
// in place in the manifold record.
struct Con
{
	uint32_t ia, ib;
	bool has_b, a_dyn, b_dyn;
	uint32_t a_dofs, b_dofs;
	float ima, imb;
	float MA[6], MB[6];
	int np;
	float friction;
	v3 n, t1, t2;
};

struct ConPts  // 36 words
{
	v3 r1[4], r2[4];
	float em[4][3];
};


// "dynamic" for everything the tick does: a dynamic body that is awake.  A sleeping body is static until woken.
__device__ __forceinline__ bool is_dynamic(uint32_t f)
{
	return (f & ((3u << BF_MOTION_SHIFT) | BF_ASLEEP)) == ((uint32_t)GPX_MOTION_DYNAMIC << BF_MOTION_SHIFT);
}
// bodies that find contacts and wake sleepers: awake dynamic ones and kinematic ones that move
__device__ __forceinline__ bool is_active_body(uint32_t f) { return is_dynamic(f) || (f & BF_KIN_MOVING); }
__device__ __forceinline__ uint32_t shape_of(uint32_t f) { return (f >> BF_SHAPE_SHIFT) & 7u; }
__device__ __forceinline__ uint32_t layer_of(uint32_t f) { return (f >> BF_LAYER_SHIFT) & 3u; }
__device__ __forceinline__ uint32_t dofs_of(uint32_t f) { return (f >> BF_DOF_SHIFT) & 63u; }

// ---- sleeping: one body's sleep test (Jolt's: three test points, each in a sphere that grows to hold it; a radius above
// 0.03 m/s x 0.5 s restarts the test, 0.5 s without a restart makes the body a candidate).  Updates the 13 floats of test
// state in `bs`; returns true for a candidate.  `g` = global body index, `f` = its flags (alive, awake, dynamic).
constexpr float SLEEP_POINT_VELOCITY = 0.03f, SLEEP_TIME = 0.5f;

__device__ __forceinline__ bool sleep_test_body(const BodyStore &bs, size_t g, uint32_t f, float dt)
{
	if (!(f & BF_ALLOW_SLEEP) || (f & BF_SENSOR))
	{
		bs.sleep_t[g] = -1.0f;
		return false;
	}
	const v3 x = V(bs.pos[g]);
	const q4 q = Q(bs.quat[g]);
	const float4 p1 = bs.prop1[g];
	const v3 e = shape_of(f) == GPX_SHAPE_SPHERE ? V(p1.x, p1.x, p1.x) : V(p1);
	const int lowest = e.x < e.y ? (e.z < e.x ? 2 : 0) : (e.z < e.y ? 2 : 1);
	const v3 ax = qrot(q, V(1.0f, 0.0f, 0.0f)), ay = qrot(q, V(0.0f, 1.0f, 0.0f)), az = qrot(q, V(0.0f, 0.0f, 1.0f));
	v3 pts[3];
	pts[0] = x;
	pts[1] = lowest == 0 ? madd(x, ay, e.y) : madd(x, ax, e.x);
	pts[2] = lowest == 2 ? madd(x, ay, e.y) : madd(x, az, e.z);
	float t = bs.sleep_t[g];
	bool restart = t < 0.0f, candidate = false;
	float4 sp[3];
#pragma unroll
	for (int k = 0; k < 3; k++)
	{
		sp[k] = bs.sleep_c[3ull * g + k];
		if (restart) continue;
		// grow the sphere just enough to hold the point
		const v3 d = pts[k] - V(sp[k]);
		const float d2 = len2(d), r = sp[k].w;
		if (d2 > (r * r))
		{
			const float dist = sqrtf(d2), nr = 0.5f * (r + dist);
			sp[k] = F4(madd(V(sp[k]), d, (nr - r) / dist), nr);
		}
		if (sp[k].w > (SLEEP_POINT_VELOCITY * SLEEP_TIME)) restart = true;
	}
	if (restart)
	{
#pragma unroll
		for (int k = 0; k < 3; k++) sp[k] = F4(pts[k], 0.0f);
		t = 0.0f;
	}
	else
	{
		t += dt;
		candidate = t >= SLEEP_TIME;
	}
#pragma unroll
	for (int k = 0; k < 3; k++) bs.sleep_c[3ull * g + k] = sp[k];
	bs.sleep_t[g] = t;
	return candidate;
}

__device__ __forceinline__ v3 mask_lin(uint32_t dofs, v3 a)
{
	if (!(dofs & 1u)) a.x = 0.0f;
	if (!(dofs & 2u)) a.y = 0.0f;
	if (!(dofs & 4u)) a.z = 0.0f;
	return a;
}
__device__ __forceinline__ v3 clamp_len(v3 v, float maxl)
{
	float l2 = len2(v);
	if (l2 > (maxl * maxl)) return v * (maxl / sqrtf(l2));
	return v;
}

// world inverse inertia R diag(inv_i) R^T, locked rotation axes zeroed; im = inverse mass (dynamic only)
__device__ __forceinline__ void body_world_inertia(SBody &b)
{
	if (!is_dynamic(b.flags))
	{
#pragma unroll
		for (int k = 0; k < 6; k++) b.M[k] = 0.0f;
		b.im = 0.0f;
		return;
	}
	m33 R = qmat(b.q);
	v3 s0 = R.c0 * b.inv_i.x, s1 = R.c1 * b.inv_i.y, s2 = R.c2 * b.inv_i.z;
	float xx = ((s0.x * R.c0.x) + (s1.x * R.c1.x)) + (s2.x * R.c2.x);
	float xy = ((s0.x * R.c0.y) + (s1.x * R.c1.y)) + (s2.x * R.c2.y);
	float xz = ((s0.x * R.c0.z) + (s1.x * R.c1.z)) + (s2.x * R.c2.z);
	float yy = ((s0.y * R.c0.y) + (s1.y * R.c1.y)) + (s2.y * R.c2.y);
	float yz = ((s0.y * R.c0.z) + (s1.y * R.c1.z)) + (s2.y * R.c2.z);
	float zz = ((s0.z * R.c0.z) + (s1.z * R.c1.z)) + (s2.z * R.c2.z);
	const uint32_t dofs = dofs_of(b.flags);
	const bool lx = !(dofs & 8u), ly = !(dofs & 16u), lz = !(dofs & 32u);
	b.M[0] = lx ? 0.0f : xx;
	b.M[1] = (lx || ly) ? 0.0f : xy;
	b.M[2] = (lx || lz) ? 0.0f : xz;
	b.M[3] = ly ? 0.0f : yy;
	b.M[4] = (ly || lz) ? 0.0f : yz;
	b.M[5] = lz ? 0.0f : zz;
	b.im = b.inv_mass;
}

__device__ __forceinline__ void body_aabb(SBody &b)
{
	v3 e;
	if (shape_of(b.flags) == GPX_SHAPE_BOX)
	{
		m33 R = qmat(b.q);
		e.x = ((fabsf(R.c0.x) * b.he.x) + (fabsf(R.c1.x) * b.he.y)) + (fabsf(R.c2.x) * b.he.z);
		e.y = ((fabsf(R.c0.y) * b.he.x) + (fabsf(R.c1.y) * b.he.y)) + (fabsf(R.c2.y) * b.he.z);
		e.z = ((fabsf(R.c0.z) * b.he.x) + (fabsf(R.c1.z) * b.he.y)) + (fabsf(R.c2.z) * b.he.z);
	}
	else
		e = V(b.he.x, b.he.x, b.he.x);
	b.lo = b.x - e;
	b.hi = b.x + e;
}

__device__ __forceinline__ bool aabb_overlap(v3 alo, v3 ahi, v3 blo, v3 bhi, float m)
{
	return (alo.x - m) <= bhi.x && blo.x <= (ahi.x + m) && (alo.y - m) <= bhi.y && blo.y <= (ahi.y + m) &&
		   (alo.z - m) <= bhi.z && blo.z <= (ahi.z + m);
}

__device__ __forceinline__ bool layers_collide(uint32_t la, uint32_t lb)
{
	// ObjectLayerShouldCollide in both orders (engine/src/physics/Physics.c:35-52)
	bool a_init = la == 1 || la == 2, b_init = lb == 1 || lb == 2;
	bool a_tgt = la == 0 || la == 1 || la == 3, b_tgt = lb == 0 || lb == 1 || lb == 3;
	return (a_init && b_tgt) || (b_init && a_tgt);
}

// Box query of the static LBVH: leaves whose exact triangle box overlaps [lo-m, hi+m], sorted by triangle index.
static __device__ __noinline__ int query_static(const float4 *__restrict__ nodes, const float4 *__restrict__ tris,
										 uint32_t n_nodes, v3 lo, v3 hi, float m, int *cand_orig, int *cand_leaf,
										 bool &overflow)
{
	int nc = 0;
	if (n_nodes == 0) return 0;
	int stack[64];
	int sp = 0, node = 0;
	const float qlx = lo.x - m, qly = lo.y - m, qlz = lo.z - m, qhx = hi.x + m, qhy = hi.y + m, qhz = hi.z + m;
	while (true)
	{
		if (node >= 0)
		{
			const float4 n0 = __ldg(&nodes[4 * node + 0]), n1 = __ldg(&nodes[4 * node + 1]),
						 n2 = __ldg(&nodes[4 * node + 2]), n3 = __ldg(&nodes[4 * node + 3]);
			bool h0 = qlx <= n0.y && n0.x <= qhx && qly <= n0.w && n0.z <= qhy && qlz <= n2.y && n2.x <= qhz;
			bool h1 = qlx <= n1.y && n1.x <= qhx && qly <= n1.w && n1.z <= qhy && qlz <= n2.w && n2.z <= qhz;
			int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
			if (h0 && h1)
			{
				stack[sp++] = c1;
				node = c0;
				continue;
			}
			if (h0) { node = c0; continue; }
			if (h1) { node = c1; continue; }
		}
		else
		{
			const int leaf = ~node;
			const float4 A = __ldg(&tris[4 * leaf + 0]), B = __ldg(&tris[4 * leaf + 1]), C = __ldg(&tris[4 * leaf + 2]);
			v3 tlo = V(fminf(A.x, fminf(B.x, C.x)), fminf(A.y, fminf(B.y, C.y)), fminf(A.z, fminf(B.z, C.z)));
			v3 thi = V(fmaxf(A.x, fmaxf(B.x, C.x)), fmaxf(A.y, fmaxf(B.y, C.y)), fmaxf(A.z, fmaxf(B.z, C.z)));
			if (aabb_overlap(lo, hi, tlo, thi, m))
			{
				if (nc < MAX_TRI_CANDIDATES)
				{
					int orig = (int)__float_as_uint(A.w);
					int k = nc++;
					while (k > 0 && cand_orig[k - 1] > orig)
					{
						cand_orig[k] = cand_orig[k - 1];
						cand_leaf[k] = cand_leaf[k - 1];
						k--;
					}
					cand_orig[k] = orig;
					cand_leaf[k] = leaf;
				}
				else
					overflow = true;
			}
		}
		if (sp == 0) break;
		node = stack[--sp];
	}
	return nc;
}

// Candidate triangles of one body.  A box query of the LBVH costs a chain of dependent L2 round trips, and a body
// that rests or creeps asks the same question every sub-step, so the answer is cached per body in global memory: the
// leaves whose boxes touch a FAT box around the body.  While the body's query box stays inside the fat box the
// candidates are re-derived from that list with the same exact leaf test the traversal applies, so the result (set
// and order) is identical to a fresh query.  Layout per body: 8 x uint4 = {count | 0xFFFFFFFF, fat lo.xyz},
// {fat hi.xyz, 0}, 6 x 4 leaf indices.
constexpr float CAND_FAT = 0.05f;
constexpr uint32_t CAND_INVALID = 0xFFFFFFFFu;

struct StaticView
{
	const float4 *nodes;
	const float4 *tris;
	uint32_t n_nodes;
};

static __device__ __noinline__ int static_candidates(const StaticView &a, uint4 *rec, v3 lo, v3 hi, int *cand_orig, int *cand_leaf,
											  bool &overflow)
{
	const float m = SPECULATIVE_DISTANCE;
	if (a.n_nodes == 0) return 0;
	const uint4 h0 = __ldcg(&rec[0]), h1 = __ldcg(&rec[1]);
	const v3 flo = V(__uint_as_float(h0.y), __uint_as_float(h0.z), __uint_as_float(h0.w));
	const v3 fhi = V(__uint_as_float(h1.x), __uint_as_float(h1.y), __uint_as_float(h1.z));
	const bool inside = h0.x != CAND_INVALID && (lo.x - m) >= flo.x && (lo.y - m) >= flo.y && (lo.z - m) >= flo.z &&
						(hi.x + m) <= fhi.x && (hi.y + m) <= fhi.y && (hi.z + m) <= fhi.z;
	if (!inside)
	{
		// refill: query the fat box; more leaves than the record holds -> leave it invalid and query exactly each time
		const v3 qlo = V((lo.x - m) - CAND_FAT, (lo.y - m) - CAND_FAT, (lo.z - m) - CAND_FAT);
		const v3 qhi = V((hi.x + m) + CAND_FAT, (hi.y + m) + CAND_FAT, (hi.z + m) + CAND_FAT);
		bool fat_overflow = false;
		int nf = query_static(a.nodes, a.tris, a.n_nodes, qlo, qhi, 0.0f, cand_orig, cand_leaf, fat_overflow);
		if (!fat_overflow)
		{
			// A triangle whose PLANE stays clear of the fat box cannot touch the body while the body stays inside it
			// (the box-triangle test starts with exactly this axis): a sloping wall's box covers half a sector, its plane
			// does not.  1 mm of slack keeps the filter on the safe side of the exact test's rounding.
			const v3 fc = V(0.5f * (qlo.x + qhi.x), 0.5f * (qlo.y + qhi.y), 0.5f * (qlo.z + qhi.z));
			const v3 fe = V(0.5f * (qhi.x - qlo.x), 0.5f * (qhi.y - qlo.y), 0.5f * (qhi.z - qlo.z));
			int kept = 0;
			for (int c = 0; c < nf; c++)
			{
				const int leaf = cand_leaf[c];
				const float4 A = __ldg(&a.tris[4 * leaf + 0]), N = __ldg(&a.tris[4 * leaf + 3]);
				const float s = ((fc.x - A.x) * N.x) + ((fc.y - A.y) * N.y) + ((fc.z - A.z) * N.z);
				const float r = (fabsf(N.x) * fe.x) + (fabsf(N.y) * fe.y) + (fabsf(N.z) * fe.z);
				if (fabsf(s) - r > m + 1.0e-3f) continue;
				cand_orig[kept] = cand_orig[c];
				cand_leaf[kept] = leaf;
				kept++;
			}
			nf = kept;
		}
		if (fat_overflow)
		{
			__stcg(&rec[0], make_uint4(CAND_INVALID, 0u, 0u, 0u));
			return query_static(a.nodes, a.tris, a.n_nodes, lo, hi, m, cand_orig, cand_leaf, overflow);
		}
		__stcg(&rec[0], make_uint4((uint32_t)nf, __float_as_uint(qlo.x), __float_as_uint(qlo.y), __float_as_uint(qlo.z)));
		__stcg(&rec[1], make_uint4(__float_as_uint(qhi.x), __float_as_uint(qhi.y), __float_as_uint(qhi.z), 0u));
		for (int c = 0; c < nf; c += 4)
			__stcg(&rec[2 + c / 4], make_uint4((uint32_t)cand_leaf[c], c + 1 < nf ? (uint32_t)cand_leaf[c + 1] : 0u,
											   c + 2 < nf ? (uint32_t)cand_leaf[c + 2] : 0u,
											   c + 3 < nf ? (uint32_t)cand_leaf[c + 3] : 0u));
		// fall through: filter the fresh list exactly like a cached one (it is already in registers/local)
		int nc = 0;
		for (int c = 0; c < nf; c++)
		{
			const int leaf = cand_leaf[c];
			const float4 A = __ldg(&a.tris[4 * leaf + 0]), B = __ldg(&a.tris[4 * leaf + 1]), C = __ldg(&a.tris[4 * leaf + 2]);
			v3 tlo = V(fminf(A.x, fminf(B.x, C.x)), fminf(A.y, fminf(B.y, C.y)), fminf(A.z, fminf(B.z, C.z)));
			v3 thi = V(fmaxf(A.x, fmaxf(B.x, C.x)), fmaxf(A.y, fmaxf(B.y, C.y)), fmaxf(A.z, fmaxf(B.z, C.z)));
			if (!aabb_overlap(lo, hi, tlo, thi, m)) continue;
			cand_orig[nc] = cand_orig[c];
			cand_leaf[nc] = leaf;
			nc++;
		}
		return nc;
	}
	const int nf = (int)h0.x;
	int nc = 0;
	for (int c0 = 0; c0 < nf; c0 += 4)
	{
		const uint4 ids = __ldcg(&rec[2 + c0 / 4]);
		const uint32_t id4[4] = {ids.x, ids.y, ids.z, ids.w};
#pragma unroll
		for (int k = 0; k < 4; k++)
		{
			if (c0 + k >= nf) break;
			const int leaf = (int)id4[k];
			const float4 A = __ldg(&a.tris[4 * leaf + 0]), B = __ldg(&a.tris[4 * leaf + 1]), C = __ldg(&a.tris[4 * leaf + 2]);
			v3 tlo = V(fminf(A.x, fminf(B.x, C.x)), fminf(A.y, fminf(B.y, C.y)), fminf(A.z, fminf(B.z, C.z)));
			v3 thi = V(fmaxf(A.x, fmaxf(B.x, C.x)), fmaxf(A.y, fmaxf(B.y, C.y)), fmaxf(A.z, fmaxf(B.z, C.z)));
			if (!aabb_overlap(lo, hi, tlo, thi, m)) continue;
			cand_orig[nc] = (int)__float_as_uint(A.w);
			cand_leaf[nc] = leaf;
			nc++;
		}
	}
	return nc;
}

// ---- solver pieces (same formulas for every manifold; b == static geometry has zero mass and velocity)

__device__ __forceinline__ float eff_mass(float ima, const float *MA, float imb, const float *MB, v3 r1, v3 r2, v3 axis)
{
	v3 r1xa = cross(r1, axis), r2xa = cross(r2, axis);
	float k = ((ima + imb) + dot(r1xa, sym_mul(MA, r1xa))) + dot(r2xa, sym_mul(MB, r2xa));
	return k > 0.0f ? 1.0f / k : 0.0f;
}

// The two bodies' velocities of one manifold, in registers for the duration of one colour phase.
struct Vel
{
	v3 va, wa, vb, wb;
};

__device__ __forceinline__ v3 rel_vel(const Con &c, const Vel &u, v3 r1, v3 r2)
{
	v3 ua = u.va + cross(u.wa, r1);
	if (!c.has_b) return ua;
	return ua - (u.vb + cross(u.wb, r2));
}

__device__ __forceinline__ void apply_impulse(const Con &c, Vel &u, v3 r1, v3 r2, v3 P)
{
	if (c.a_dyn)
	{
		u.va = u.va - mask_lin(c.a_dofs, P * c.ima);
		u.wa = u.wa - sym_mul(c.MA, cross(r1, P));
	}
	if (c.has_b && c.b_dyn)
	{
		u.vb = u.vb + mask_lin(c.b_dofs, P * c.imb);
		u.wb = u.wb + sym_mul(c.MB, cross(r2, P));
	}
}

__device__ __forceinline__ void load_vel(const Con &c, const SBody *bodies, Vel &u)
{
	u.va = bodies[c.ia].v;
	u.wa = bodies[c.ia].w;
	if (c.has_b)
	{
		u.vb = bodies[c.ib].v;
		u.wb = bodies[c.ib].w;
	}
	else
		u.vb = u.wb = V(0.0f, 0.0f, 0.0f);
}

__device__ __forceinline__ void store_vel(const Con &c, SBody *bodies, const Vel &u)
{
	if (c.a_dyn)
	{
		bodies[c.ia].v = u.va;
		bodies[c.ia].w = u.wa;
	}
	if (c.has_b && c.b_dyn)
	{
		bodies[c.ib].v = u.vb;
		bodies[c.ib].w = u.wb;
	}
}

// Per-manifold part of the constraint set-up (reads body state only)
__device__ __forceinline__ void con_header(Con &c, const SMan &m, const SBody *bodies)
{
	const SBody &A = bodies[m.a];
	c.ia = m.a;
	c.has_b = m.b < STATIC_BODY_BASE;
	c.ib = c.has_b ? m.b : m.a;
	const SBody &B = bodies[c.ib];
	c.a_dyn = is_dynamic(A.flags);
	c.b_dyn = c.has_b && is_dynamic(B.flags);
	c.a_dofs = dofs_of(A.flags);
	c.b_dofs = dofs_of(B.flags);
	c.ima = A.im;
	c.imb = c.has_b ? B.im : 0.0f;
#pragma unroll
	for (int k = 0; k < 6; k++)
	{
		c.MA[k] = A.M[k];
		c.MB[k] = c.has_b ? B.M[k] : 0.0f;
	}
	c.np = m.np;
	c.friction = m.friction;
	c.n = m.n;
	c.t1 = vperp(c.n);
	c.t2 = cross(c.n, c.t1);
}

// Full set-up of one manifold: header + per-point lever arms / effective masses into `pt`, and the speculative /
// restitution bias into the manifold record.
__device__ __forceinline__ void build_con(Con &c, ConPts &pt, SMan &m, const SBody *bodies, float h)
{
	con_header(c, m, bodies);
	const SBody &A = bodies[c.ia], &B = bodies[c.ib];
	const v3 ax = A.x, bx = B.x;
	const q4 aq = A.q, bq = B.q;
	Vel u;
	load_vel(c, bodies, u);
#pragma unroll 1
	for (int k = 0; k < c.np; k++)
	{
		v3 p1 = ax + qrot(aq, m.p1l[k]);
		v3 p2 = c.has_b ? bx + qrot(bq, m.p2l[k]) : m.p2l[k];
		v3 mid = (p1 + p2) * 0.5f;
		const v3 r1 = mid - ax;
		const v3 r2 = c.has_b ? mid - bx : V(0.0f, 0.0f, 0.0f);
		pt.r1[k] = r1;
		pt.r2[k] = r2;
		pt.em[k][0] = eff_mass(c.ima, c.MA, c.imb, c.MB, r1, r2, c.n);
		pt.em[k][1] = eff_mass(c.ima, c.MA, c.imb, c.MB, r1, r2, c.t1);
		pt.em[k][2] = eff_mass(c.ima, c.MA, c.imb, c.MB, r1, r2, c.t2);
		float pen = dot(p1 - p2, c.n);
		float bias = fmaxf(0.0f, -pen / h);
		if (m.restitution > 0.0f)
		{
			float nv = -dot(c.n, rel_vel(c, u, r1, r2));
			if (nv < -MIN_VELOCITY_FOR_RESTITUTION) bias = m.restitution * nv;
		}
		m.bias[k] = bias;
	}
}

// ConPts <-> 9 float4 in global memory (worlds with more manifolds than lanes; the wide-world kernels)
__device__ __forceinline__ void park_con(const ConPts &pt, float4 *g)
{
	const float *f = reinterpret_cast<const float *>(&pt);
#pragma unroll
	for (int i = 0; i < 9; i++) __stcg(&g[i], make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]));
}

__device__ __forceinline__ void unpark_con(ConPts &pt, const float4 *g)
{
	float *f = reinterpret_cast<float *>(&pt);
#pragma unroll
	for (int i = 0; i < 9; i++)
	{
		const float4 v = __ldcg(&g[i]);
		f[4 * i] = v.x; f[4 * i + 1] = v.y; f[4 * i + 2] = v.z; f[4 * i + 3] = v.w;
	}
}

// `M` is any record with ln / lt1 / lt2 (and, for solve_velocity, bias) arrays: the manifold itself or a solver record
template <typename M>
__device__ __forceinline__ void warm_start(const Con &c, const ConPts &pt, const M &m, Vel &u)
{
#pragma unroll 1
	for (int k = 0; k < c.np; k++)
	{
		const float ln = m.ln[k], lt1 = m.lt1[k], lt2 = m.lt2[k];
		if (ln == 0.0f && lt1 == 0.0f && lt2 == 0.0f) continue;
		v3 P = ((c.n * ln) + (c.t1 * lt1)) + (c.t2 * lt2);
		apply_impulse(c, u, pt.r1[k], pt.r2[k], P);
	}
}

template <typename M>
__device__ __forceinline__ void solve_velocity(const Con &c, const ConPts &pt, M &m, Vel &u)
{
	// friction first: non-penetration is more important, so it goes last
#pragma unroll 1
	for (int k = 0; k < c.np; k++)
	{
		const float o1 = m.lt1[k], o2 = m.lt2[k];
		const float maxf = c.friction * m.ln[k];
		// nothing to hold with and nothing held (a speculative point that does not touch): the row stays at zero
		if (maxf == 0.0f && o1 == 0.0f && o2 == 0.0f) continue;
		const v3 r1 = pt.r1[k], r2 = pt.r2[k];
		v3 rv = rel_vel(c, u, r1, r2);
		float l1 = o1 + (pt.em[k][1] * dot(c.t1, rv));
		float l2 = o2 + (pt.em[k][2] * dot(c.t2, rv));
		float sq = (l1 * l1) + (l2 * l2);
		if (sq > (maxf * maxf))
		{
			// no normal impulse yet (a speculative point): 0 / sqrt(sq) is that zero, skip the division and the root
			float s = maxf == 0.0f ? maxf : maxf / sqrtf(sq);
			l1 = l1 * s;
			l2 = l2 * s;
		}
		const float d1 = l1 - o1, d2 = l2 - o2;
		m.lt1[k] = l1;
		m.lt2[k] = l2;
		// an impulse is applied only when it is not zero (Jolt's AxisConstraintPart::ApplyVelocityStep)
		if (d1 != 0.0f || d2 != 0.0f) apply_impulse(c, u, r1, r2, (c.t1 * d1) + (c.t2 * d2));
	}
#pragma unroll 1
	for (int k = 0; k < c.np; k++)
	{
		const v3 r1 = pt.r1[k], r2 = pt.r2[k];
		const float old = m.ln[k];
		v3 rv = rel_vel(c, u, r1, r2);
		float lambda = pt.em[k][0] * (dot(c.n, rv) - m.bias[k]);
		float nt = fmaxf(0.0f, old + lambda);
		lambda = nt - old;
		m.ln[k] = nt;
		if (lambda != 0.0f) apply_impulse(c, u, r1, r2, c.n * lambda);
	}
}

static __device__ __noinline__ void solve_position(SMan &m, SBody *bodies)
{
	SBody &A = bodies[m.a];
	SBody *B = m.b < STATIC_BODY_BASE ? &bodies[m.b] : nullptr;
	const float zero_m[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
	// The world inertia of the two bodies is refreshed when the first point needs a correction — at the orientation the
	// manifold's pass starts with, as the restatement does up front; a manifold inside the slop (the usual case at
	// rest) costs four separations and nothing else.  Nothing reads M between here and the next sub-step's refresh.
	bool ready = false;
	float imb = 0.0f;
	const float *MB = zero_m;
	for (int k = 0; k < m.np; k++)
	{
		v3 p1 = A.x + qrot(A.q, m.p1l[k]);
		v3 p2 = B ? B->x + qrot(B->q, m.p2l[k]) : m.p2l[k];
		float sep = dot(p2 - p1, m.n) + PENETRATION_SLOP;
		if (sep >= 0.0f) continue;
		if (!ready)
		{
			body_world_inertia(A);
			if (B)
			{
				body_world_inertia(*B);
				imb = B->im;
				MB = B->M;
			}
			ready = true;
		}
		v3 mid = (p1 + p2) * 0.5f;
		v3 r1 = mid - A.x, r2 = B ? mid - B->x : V(0.0f, 0.0f, 0.0f);
		float e = eff_mass(A.im, A.M, imb, MB, r1, r2, m.n);
		float c = fmaxf(sep, -MAX_PENETRATION_DISTANCE);
		float lambda = (-e * BAUMGARTE) * c;
		v3 P = m.n * lambda;
		if (is_dynamic(A.flags))
		{
			A.x = A.x - mask_lin(dofs_of(A.flags), P * A.im);
			A.q = qstep(A.q, -sym_mul(A.M, cross(r1, P)));
		}
		if (B && is_dynamic(B->flags))
		{
			B->x = B->x + mask_lin(dofs_of(B->flags), P * B->im);
			B->q = qstep(B->q, sym_mul(B->M, cross(r2, P)));
		}
	}
}

// manifold points: world -> local frames (b static: world)
__device__ __forceinline__ void store_points(SMan &m, const SBody &A, const SBody *B, int np, const v3 *p1, const v3 *p2)
{
	m33 RA = qmat(A.q);
	m.np = np;
	m33 RB;
	if (B) RB = qmat(B->q);
	for (int i = 0; i < np; i++)
	{
		m.p1l[i] = mtmul(RA, p1[i] - A.x);
		m.p2l[i] = B ? mtmul(RB, p2[i] - B->x) : p2[i];
	}
	for (int i = 0; i < 4; i++) m.ln[i] = m.lt1[i] = m.lt2[i] = 0.0f;
}

struct StaticSlot
{
	v3 n;
	float depth, friction;
	uint32_t sbody;
	int np;
	v3 p1[8], p2[8];
};

// Contacts of one body against the static map: candidate triangles (cached LBVH box query), box/sphere-vs-triangle
// manifolds, grouped per static body into <= MAX_SLOTS manifolds by normal and reduced to four points each.
// Returns the number of slots filled; `err` collects capacity errors.
__device__ __forceinline__ int body_static_contacts(const StaticView &sv, uint4 *cand_rec, const SBody &A, Scratch &scratch,
													StaticSlot *slots, uint32_t &err)
{
	const uint32_t fa = A.flags;
	int nslots = 0;
	int cand_orig[MAX_TRI_CANDIDATES], cand_leaf[MAX_TRI_CANDIDATES];
	bool overflow = false;
	const int nc = static_candidates(sv, cand_rec, A.lo, A.hi, cand_orig, cand_leaf, overflow);
	if (overflow) err |= GPX_ERR_BODY_PAIR_CACHE_FULL;
	Box bx;
	bx.x = A.x;
	bx.R = qmat(A.q);
	bx.he = A.he;
	int group_start = 0;
	uint32_t cur_body = 0xFFFFFFFFu;
	for (int c = 0; c < nc; c++)
	{
		const int leaf = cand_leaf[c];
		const float4 TA = __ldg(&sv.tris[4 * leaf + 0]), TB = __ldg(&sv.tris[4 * leaf + 1]),
					 TC = __ldg(&sv.tris[4 * leaf + 2]), TN = __ldg(&sv.tris[4 * leaf + 3]);
		const uint32_t sbody = __float_as_uint(TB.w);
		if (sbody != cur_body)
		{
			cur_body = sbody;
			group_start = nslots;
		}
		Tri T;
		T.a = V(TA); T.b = V(TB); T.c = V(TC); T.n = V(TN);
		Hit hit;
		hit_bind(hit, scratch);
		bool ok = shape_of(fa) == GPX_SHAPE_BOX ? collide_box_tri(bx, T, SPECULATIVE_DISTANCE, scratch, hit)
												: collide_sphere_tri(A.x, A.he.x, T, SPECULATIVE_DISTANCE, hit);
		if (!ok) continue;
		prune_points(A.x, hit.n, hit.np, hit.p1, hit.p2);
		int s = -1;
		for (int k = group_start; k < nslots; k++)
			if (dot(slots[k].n, hit.n) >= NORMAL_COS_MAX_DELTA)
			{
				s = k;
				break;
			}
		if (s < 0)
		{
			if (nslots - group_start == MAX_SLOTS) continue;
			if (nslots == MAX_STATIC_PER_BODY)
			{
				err |= GPX_ERR_MANIFOLD_CACHE_FULL;
				continue;
			}
			s = nslots++;
			slots[s].n = hit.n;
			slots[s].depth = hit.depth;
			slots[s].friction = sqrtf(A.friction * TC.w);
			slots[s].sbody = sbody;
			slots[s].np = 0;
		}
		else if (hit.depth > slots[s].depth)
		{
			slots[s].depth = hit.depth;
			slots[s].n = hit.n;
		}
		for (int k = 0; k < hit.np; k++)
		{
			slots[s].p1[slots[s].np] = hit.p1[k];
			slots[s].p2[slots[s].np] = hit.p2[k];
			slots[s].np++;
		}
		prune_points(A.x, slots[s].n, slots[s].np, slots[s].p1, slots[s].p2);
	}
					return nslots;
}

// Contact manifold of one body pair (one thread per pair): fills normal, material and <= 4 local points of `m`;
// m.np stays 0 when the shapes are farther apart than the speculative distance.
__device__ __forceinline__ void pair_contact(const SBody &A, const SBody &B, Scratch &scratch, SMan &m)
{
	Hit hit;
	hit_bind(hit, scratch);
	bool ok;
	const uint32_t sa = shape_of(A.flags), sb = shape_of(B.flags);
	if (sa == GPX_SHAPE_BOX && sb == GPX_SHAPE_BOX)
	{
		Box ba, bb;
		ba.x = A.x; ba.R = qmat(A.q); ba.he = A.he;
		bb.x = B.x; bb.R = qmat(B.q); bb.he = B.he;
		ok = collide_box_box(ba, bb, SPECULATIVE_DISTANCE, scratch, hit);
	}
	else if (sa == GPX_SHAPE_SPHERE && sb == GPX_SHAPE_SPHERE)
		ok = collide_sphere_sphere(A.x, A.he.x, B.x, B.he.x, SPECULATIVE_DISTANCE, hit);
	else if (sa == GPX_SHAPE_SPHERE)
	{
		Box bb;
		bb.x = B.x; bb.R = qmat(B.q); bb.he = B.he;
		ok = collide_sphere_box(A.x, A.he.x, bb, SPECULATIVE_DISTANCE, hit);
	}
	else
	{
		Box ba;
		ba.x = A.x; ba.R = qmat(A.q); ba.he = A.he;
		ok = collide_sphere_box(B.x, B.he.x, ba, SPECULATIVE_DISTANCE, hit);
		if (ok)
		{
			hit.n = -hit.n;
			v3 t = hit.p1[0];
			hit.p1[0] = hit.p2[0];
			hit.p2[0] = t;
		}
	}
	if (!ok) return;
	prune_points(A.x, hit.n, hit.np, hit.p1, hit.p2);
	m.n = hit.n;
	m.friction = sqrtf(A.friction * B.friction);
	m.restitution = fmaxf(A.restitution, B.restitution);
	store_points(m, A, &B, hit.np, hit.p1, hit.p2);
}

}  // namespace gpx
