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

struct SMan  // 41 words (odd stride): a contact manifold between narrowphase, solver and warm-start cache
{
	uint32_t a, b;
	int np;      // 0 = empty slot (pair that did not touch)
	int colour;
	v3 n;
	float friction, restitution;
	v3 p1l[4], p2l[4];
	float ln[4];  // accumulated non-penetration impulse per point
	float cf[3];  // accumulated friction impulse of the manifold: tangent 1, tangent 2 (through the centroid), twist about n
	uint32_t tri;  // against a static body: the triangle that opened this manifold's slot (part of the warm-start key); else 0
};

// One constraint row: the velocity it measures is Jv = (axis . va + a1 . wa) - (axis . vb + a2 . wb); an impulse d along it
// changes va by -lA d, wa by -I1 d, vb by +lB d, wb by +I2 d (lA / lB = axis times the body's inverse mass, kept once per
// axis in Rows).  em = 1 / (J M^-1 J^T).  Everything an iteration needs is precomputed: a row costs 12 FMAs to measure
// and 12 to apply.
struct Row
{
	v3 a1, a2;  // r1 x axis, r2 x axis
	v3 I1, I2;  // world inverse inertia times a1 / a2
	float em;
};

// The friction rows of a manifold: two tangent rows through the centroid of the contact points plus a twist row about
// the normal, limited by friction x (sum of the normal impulses) (x the patch radius for the twist).  52 words = 13
// float4: with that stride the 16-byte shared-memory loads of a quarter-warp fall on different banks.
struct __align__(16) FricRows
{
	Row t[2];
	v3 wI1, wI2;  // twist row: inverse inertias times the normal
	float wem;
	float rp;     // patch radius: RMS distance of the contact points from their centroid
	v3 tA[2], tB[2];  // tangent times inverse mass, locked translation axes zeroed
	float pad[6];
};
static_assert(sizeof(FricRows) == 52 * sizeof(float), "FricRows is staged as 13 float4");

// The rows of one manifold, rebuilt every sub-step from the bodies' poses: a non-penetration row per contact point
// (rows of points the manifold does not have are zero: they measure nothing and apply nothing) and the friction rows.
// In the ensemble kernel a lane keeps the non-penetration rows in REGISTERS for the whole velocity solve (every index
// is a compile-time constant after unrolling) and the friction rows, which an iteration visits once, in its slice of
// shared memory; the wide-world kernels keep everything in their solver records.
struct __align__(16) Rows
{
	Row n[4];
	float bias[4];
	float ln[4];  // words 56..59 (float4 14 of the parked record)
	float cf[3];  // words 60..62 (float4 15)
	float pad;
	v3 nA, nB;    // normal times inverse mass, locked translation axes zeroed
	float pad2[2];
	FricRows f;
};
static_assert(sizeof(Rows) == 124 * sizeof(float), "Rows is parked as 31 float4");

// What a manifold's lane needs besides the rows.
struct Con
{
	uint32_t ia, ib;
	bool has_b, a_dyn, b_dyn;
	int np;
	float friction;
	v3 n, t1, t2;
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

// v - t * s
__device__ __forceinline__ v3 msub(v3 v, v3 t, float s) { return V(fmaf(-t.x, s, v.x), fmaf(-t.y, s, v.y), fmaf(-t.z, s, v.z)); }

// The two bodies' velocities of one manifold, in registers for the duration of one colour phase.
struct Vel
{
	v3 va, wa, vb, wb;
};

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

__device__ __forceinline__ void row_setup(Row &r, float ima, const float *MA, float imb, const float *MB, v3 r1, v3 r2, v3 axis)
{
	r.a1 = cross(r1, axis);
	r.a2 = cross(r2, axis);
	r.I1 = sym_mul(MA, r.a1);
	r.I2 = sym_mul(MB, r.a2);
	const float k = ((ima + imb) + dot(r.a1, r.I1)) + dot(r.a2, r.I2);
	r.em = k > 0.0f ? 1.0f / k : 0.0f;
}

__device__ __forceinline__ float row_jv(const Row &r, v3 axis, const Vel &u)
{
	return (dot(axis, u.va) + dot(r.a1, u.wa)) - (dot(axis, u.vb) + dot(r.a2, u.wb));
}

__device__ __forceinline__ void row_apply(const Row &r, v3 lA, v3 lB, float d, Vel &u)
{
	u.va = msub(u.va, lA, d);
	u.wa = msub(u.wa, r.I1, d);
	u.vb = madd(u.vb, lB, d);
	u.wb = madd(u.wb, r.I2, d);
}

// Set-up of one manifold (reads body state only): header, rows, speculative / restitution bias; the accumulated impulses
// move from the manifold record into the rows.  Rows of points the manifold does not have are zero: they measure
// nothing and apply nothing.
__device__ __forceinline__ void build_rows(Con &c, Rows &R, const SMan &m, const SBody *bodies, float h)
{
	const SBody &A = bodies[m.a];
	c.ia = m.a;
	c.has_b = m.b < STATIC_BODY_BASE;
	c.ib = c.has_b ? m.b : m.a;
	const SBody &B = bodies[c.ib];
	c.a_dyn = is_dynamic(A.flags);
	c.b_dyn = c.has_b && is_dynamic(B.flags);
	const uint32_t a_dofs = dofs_of(A.flags), b_dofs = dofs_of(B.flags);
	const float ima = A.im, imb = c.has_b ? B.im : 0.0f;
	float MA[6], MB[6];
#pragma unroll
	for (int k = 0; k < 6; k++)
	{
		MA[k] = A.M[k];
		MB[k] = c.has_b ? B.M[k] : 0.0f;
	}
	c.np = m.np;
	c.friction = m.friction;
	c.n = m.n;
	c.t1 = vperp(c.n);
	c.t2 = cross(c.n, c.t1);
	const v3 zero = V(0.0f, 0.0f, 0.0f);
	R.nA = mask_lin(a_dofs, c.n * ima);
	R.f.tA[0] = mask_lin(a_dofs, c.t1 * ima);
	R.f.tA[1] = mask_lin(a_dofs, c.t2 * ima);
	R.nB = c.has_b ? mask_lin(b_dofs, c.n * imb) : zero;
	R.f.tB[0] = c.has_b ? mask_lin(b_dofs, c.t1 * imb) : zero;
	R.f.tB[1] = c.has_b ? mask_lin(b_dofs, c.t2 * imb) : zero;
	const v3 ax = A.x, bx = B.x;
	const q4 aq = A.q, bq = B.q;
	Vel u;
	load_vel(c, bodies, u);
	v3 mid[4];
	v3 csum = zero;
#pragma unroll
	for (int k = 0; k < 4; k++)
	{
		if (k < c.np)
		{
			const v3 p1 = ax + qrot(aq, m.p1l[k]);
			const v3 p2 = c.has_b ? bx + qrot(bq, m.p2l[k]) : m.p2l[k];
			mid[k] = (p1 + p2) * 0.5f;
			csum = csum + mid[k];
			const v3 r1 = mid[k] - ax;
			const v3 r2 = c.has_b ? mid[k] - bx : zero;
			row_setup(R.n[k], ima, MA, imb, MB, r1, r2, c.n);
			const float pen = dot(p1 - p2, c.n);
			float bias = fmaxf(0.0f, -pen / h);
			if (m.restitution > 0.0f)
			{
				const float nv = -row_jv(R.n[k], c.n, u);
				if (nv < -MIN_VELOCITY_FOR_RESTITUTION) bias = m.restitution * nv;
			}
			R.bias[k] = bias;
			R.ln[k] = m.ln[k];
		}
		else
		{
			mid[k] = zero;
			R.n[k].a1 = R.n[k].a2 = R.n[k].I1 = R.n[k].I2 = zero;
			R.n[k].em = 0.0f;
			R.bias[k] = 0.0f;
			R.ln[k] = 0.0f;
		}
	}
	R.cf[0] = m.cf[0];
	R.cf[1] = m.cf[1];
	R.cf[2] = m.cf[2];
	R.pad = R.pad2[0] = R.pad2[1] = 0.0f;
#pragma unroll
	for (int k = 0; k < 6; k++) R.f.pad[k] = 0.0f;
	const float inv_np = 1.0f / (float)c.np;
	const v3 cen = csum * inv_np;
	float s = 0.0f;
#pragma unroll
	for (int k = 0; k < 4; k++)
		if (k < c.np) s = s + len2(mid[k] - cen);
	R.f.rp = sqrtf(s * inv_np);
	const v3 rc1 = cen - ax, rc2 = c.has_b ? cen - bx : zero;
	row_setup(R.f.t[0], ima, MA, imb, MB, rc1, rc2, c.t1);
	row_setup(R.f.t[1], ima, MA, imb, MB, rc1, rc2, c.t2);
	R.f.wI1 = sym_mul(MA, c.n);
	R.f.wI2 = sym_mul(MB, c.n);
	const float kw = dot(c.n, R.f.wI1) + dot(c.n, R.f.wI2);
	R.f.wem = kw > 0.0f ? 1.0f / kw : 0.0f;
}

// accumulated impulses back into the manifold record (warm-start cache, position pass keep using SMan)
__device__ __forceinline__ void save_impulses(SMan &m, const Rows &R)
{
#pragma unroll
	for (int k = 0; k < 4; k++) m.ln[k] = R.ln[k];
	m.cf[0] = R.cf[0];
	m.cf[1] = R.cf[1];
	m.cf[2] = R.cf[2];
}

// Rows <-> 31 float4 in global memory (worlds with more manifolds than lanes)
__device__ __forceinline__ void park_rows(const Rows &R, float4 *g)
{
	const float *f = reinterpret_cast<const float *>(&R);
#pragma unroll
	for (int i = 0; i < 31; i++) __stcg(&g[i], make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]));
}

__device__ __forceinline__ void unpark_rows(Rows &R, const float4 *g)
{
	float *f = reinterpret_cast<float *>(&R);
#pragma unroll
	for (int i = 0; i < 31; i++)
	{
		const float4 v = __ldcg(&g[i]);
		f[4 * i] = v.x; f[4 * i + 1] = v.y; f[4 * i + 2] = v.z; f[4 * i + 3] = v.w;
	}
}

// the manifold header alone, for a lane that pulls parked rows
__device__ __forceinline__ void con_header(Con &c, const SMan &m, const SBody *bodies)
{
	c.ia = m.a;
	c.has_b = m.b < STATIC_BODY_BASE;
	c.ib = c.has_b ? m.b : m.a;
	c.a_dyn = is_dynamic(bodies[c.ia].flags);
	c.b_dyn = c.has_b && is_dynamic(bodies[c.ib].flags);
	c.np = m.np;
	c.friction = m.friction;
	c.n = m.n;
	c.t1 = vperp(c.n);
	c.t2 = cross(c.n, c.t1);
}

// re-apply the impulses carried over from the previous sub-step.  `F` are the manifold's friction rows: R.f or a copy.
__device__ __forceinline__ void warm_start(const Con &c, const Rows &R, const FricRows &F, Vel &u)
{
#pragma unroll
	for (int k = 0; k < 4; k++) row_apply(R.n[k], R.nA, R.nB, R.ln[k], u);
	row_apply(F.t[0], F.tA[0], F.tB[0], R.cf[0], u);
	row_apply(F.t[1], F.tA[1], F.tB[1], R.cf[1], u);
	u.wa = msub(u.wa, F.wI1, R.cf[2]);
	u.wb = madd(u.wb, F.wI2, R.cf[2]);
}

__device__ __forceinline__ void normal_row(const Con &c, Rows &R, int k, Vel &u)
{
	float lambda = R.n[k].em * (row_jv(R.n[k], c.n, u) - R.bias[k]);
	const float nt = fmaxf(0.0f, R.ln[k] + lambda);
	lambda = nt - R.ln[k];
	R.ln[k] = nt;
	row_apply(R.n[k], R.nA, R.nB, lambda, u);
}

// One velocity iteration of a manifold: friction first (non-penetration is more important, so it goes last) — the two
// tangent rows through the centroid share one limit, then the twist row — then the four non-penetration rows, forwards
// in even iterations and backwards in odd ones (a fixed order leaves the residual on the same point every time, and a
// tall stack turns that bias into a whirl that never dies).
__device__ __forceinline__ void solve_velocity(const Con &c, Rows &R, const FricRows &F, Vel &u, uint32_t it)
{
	const float maxf = c.friction * (((R.ln[0] + R.ln[1]) + R.ln[2]) + R.ln[3]);
	// nothing to hold with and nothing held (speculative points that do not touch): the rows stay at zero
	if (!(maxf == 0.0f && R.cf[0] == 0.0f && R.cf[1] == 0.0f))
	{
		float l1 = R.cf[0] + (F.t[0].em * row_jv(F.t[0], c.t1, u));
		float l2 = R.cf[1] + (F.t[1].em * row_jv(F.t[1], c.t2, u));
		const float sq = (l1 * l1) + (l2 * l2);
		if (sq > (maxf * maxf))
		{
			// no normal impulse: 0 / sqrt(sq) is that zero, skip the division and the root
			const float sc = maxf == 0.0f ? maxf : maxf / sqrtf(sq);
			l1 = l1 * sc;
			l2 = l2 * sc;
		}
		const float d1 = l1 - R.cf[0], d2 = l2 - R.cf[1];
		R.cf[0] = l1;
		R.cf[1] = l2;
		row_apply(F.t[0], F.tA[0], F.tB[0], d1, u);
		row_apply(F.t[1], F.tA[1], F.tB[1], d2, u);
	}
	if (c.np >= 2)
	{
		const float lim = maxf * F.rp;
		if (!(lim == 0.0f && R.cf[2] == 0.0f))
		{
			float l = R.cf[2] + (F.wem * (dot(c.n, u.wa) - dot(c.n, u.wb)));
			l = fminf(fmaxf(l, -lim), lim);
			const float d = l - R.cf[2];
			R.cf[2] = l;
			u.wa = msub(u.wa, F.wI1, d);
			u.wb = madd(u.wb, F.wI2, d);
		}
	}
	if (it & 1u)
	{
#pragma unroll
		for (int k = 3; k >= 0; k--) normal_row(c, R, k, u);
	}
	else
	{
#pragma unroll
		for (int k = 0; k < 4; k++) normal_row(c, R, k, u);
	}
}

// returns whether any point of the manifold was outside the slop (i.e. whether anything moved)
static __device__ __noinline__ bool solve_position(SMan &m, SBody *bodies)
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
	return ready;
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
	for (int i = 0; i < 4; i++) m.ln[i] = 0.0f;
	m.cf[0] = m.cf[1] = m.cf[2] = 0.0f;
}

struct StaticSlot
{
	v3 n;
	float depth, friction;
	uint32_t sbody;
	uint32_t tri;  // the triangle that opened the slot
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
			slots[s].tri = (uint32_t)cand_orig[c];
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
