// gpx_rays.cu — batched closest-hit ray queries against the static LBVH and the world's bodies.
//
// Replaces JPH_NarrowPhaseQuery_CastRay_GAME / CastRay2_GAME as called by the player crosshair
// (engine/src/physics/PlayerPhysics.c:297-315) and by lasers (game/src/actor/prop/Laser.c:127-158).  The filter
// callbacks of the reference are pure functions of the object layer, so they arrive here as a layer bit mask.
//
// Kernel shape: persistent CTAs, one per SM slot; each CTA copies the whole tree (nodes + triangle records) into
// shared memory once with a bulk async copy, then its warps pull batches of 32 rays.  A ray is 32 B in, 16 B out;
// everything else stays on chip.
#include <atomic>
#include <cstdlib>
#include "gpx_internal.h"
#include "gpx_math.cuh"
#include "gpx_narrow.cuh"

namespace gpx {

constexpr int RAY_THREADS = 256;       // CTA size when several CTAs share an SM
constexpr int RAY_THREADS_MAX = 1024;  // one CTA per SM (a large tree in shared memory): as many warps as registers allow
constexpr int STACK_DEPTH = 64;  // a radix tree over 64-bit keys is at most 64 levels deep
constexpr int SSTACK_DEPTH = 32;  // levels of the shared-memory traversal stack (16-bit entries, one column per thread)

struct RayArgs
{
	const float4 *nodes;
	const float4 *tris;
	uint32_t n_nodes, n_tris;
	const float4 *rays;  // 2 per ray
	float4 *hits;        // 1 per ray
	unsigned long long n;
	// bodies (for masks that include non-static layers)
	const float4 *pos, *quat, *prop1;
	const uint32_t *flags;
	uint32_t worlds, cap;
};

__device__ __forceinline__ bool ray_tri(v3 o, v3 d, float tmax, v3 a, v3 b, v3 c, float &tout)
{
	v3 e1 = b - a, e2 = c - a;
	v3 p = cross(d, e2);
	float det = dot(e1, p);
	if (fabsf(det) < 1.0e-12f) return false;
	float inv = 1.0f / det;
	v3 tv = o - a;
	float u = dot(tv, p) * inv;
	if (u < 0.0f || u > 1.0f) return false;
	v3 q = cross(tv, e1);
	float v = dot(d, q) * inv;
	if (v < 0.0f || (u + v) > 1.0f) return false;
	float t = dot(e2, q) * inv;
	if (t < 0.0f || t > tmax) return false;
	tout = t;
	return true;
}

__device__ __forceinline__ bool ray_box(v3 o, v3 d, float tmax, v3 x, q4 q, v3 he, float &tout, uint32_t &face)
{
	m33 R = qmat(q);
	v3 lo = mtmul(R, o - x);
	v3 ld = mtmul(R, d);
	float tn = -3.0e38f, tf = 3.0e38f;
	uint32_t fn = 0;
#pragma unroll
	for (int k = 0; k < 3; k++)
	{
		float ok = get(lo, k), dk = get(ld, k), hk = get(he, k);
		if (dk == 0.0f)
		{
			if (ok < -hk || ok > hk) return false;
			continue;
		}
		float inv = 1.0f / dk;
		float t1 = (-hk - ok) * inv, t2 = (hk - ok) * inv;
		uint32_t f1 = 2u * k, f2 = 2u * k + 1u;
		if (t1 > t2)
		{
			float tt = t1; t1 = t2; t2 = tt;
			f1 = f2;
		}
		if (t1 > tn) { tn = t1; fn = f1; }
		if (t2 < tf) tf = t2;
	}
	if (tn > tf || tf < 0.0f) return false;
	float t = tn < 0.0f ? 0.0f : tn;
	if (t > tmax) return false;
	tout = t;
	face = fn;
	return true;
}

__device__ __forceinline__ bool ray_sphere(v3 o, v3 d, float tmax, v3 x, float r, float &tout)
{
	v3 m = o - x;
	float bq = dot(m, d);
	float c = dot(m, m) - (r * r);
	if (c > 0.0f && bq > 0.0f) return false;
	float disc = (bq * bq) - c;
	if (disc < 0.0f) return false;
	float t = -bq - sqrtf(disc);
	if (t < 0.0f) t = 0.0f;
	if (t > tmax) return false;
	tout = t;
	return true;
}

// One ray through the tree.  NODES/TRIS point at shared or global memory.  SSTACK: the traversal stack is this thread's
// column of a shared-memory array of 16-bit node codes (entry k at sstack[k * stride]: a warp's accesses fall on
// consecutive half-words) — the per-lane array it replaces lived in local memory and cost 2.7 M local loads and stores
// per 2^20 rays; used when the tree is known to be shallow enough and its node codes fit 16 bits.
template <bool SSTACK>
__device__ __forceinline__ void trace_static(const float4 *__restrict__ NODES, const float4 *__restrict__ TRIS,
											 uint32_t n_nodes, v3 o, v3 d, float tmax, bool need_flag, float &best,
											 uint32_t &bbody, uint32_t &bface, short *sstack, uint32_t stride)
{
	if (n_nodes == 0) return;
	// reciprocal direction with zeros nudged so that slabs never produce 0 * inf
	const float tiny = 1.0e-20f;
	float idx = 1.0f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x));
	float idy = 1.0f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y));
	float idz = 1.0f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z));
	// the slab arithmetic only decides which nodes are visited (boxes are padded by BVH_PAD, far more than its rounding),
	// so it may use fused multiply-adds; the triangle test below is the exact one
	const float ox = -(o.x * idx), oy = -(o.y * idy), oz = -(o.z * idz);
	int stack[SSTACK ? 1 : STACK_DEPTH];
	int sp = 0;
	int node = 0;
	float limit = fminf(best, tmax);
	while (true)
	{
		if (node >= 0)
		{
			const float4 n0 = NODES[4 * node + 0], n1 = NODES[4 * node + 1], n2 = NODES[4 * node + 2],
						 n3 = NODES[4 * node + 3];
			// child 0
			float ax0 = fmaf(n0.x, idx, ox), ax1 = fmaf(n0.y, idx, ox), ay0 = fmaf(n0.z, idy, oy), ay1 = fmaf(n0.w, idy, oy);
			float az0 = fmaf(n2.x, idz, oz), az1 = fmaf(n2.y, idz, oz);
			float tn0 = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fmaxf(fminf(az0, az1), 0.0f));
			float tf0 = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fminf(fmaxf(az0, az1), limit));
			// child 1
			float bx0 = fmaf(n1.x, idx, ox), bx1 = fmaf(n1.y, idx, ox), by0 = fmaf(n1.z, idy, oy), by1 = fmaf(n1.w, idy, oy);
			float bz0 = fmaf(n2.z, idz, oz), bz1 = fmaf(n2.w, idz, oz);
			float tn1 = fmaxf(fmaxf(fminf(bx0, bx1), fminf(by0, by1)), fmaxf(fminf(bz0, bz1), 0.0f));
			float tf1 = fminf(fminf(fmaxf(bx0, bx1), fmaxf(by0, by1)), fminf(fmaxf(bz0, bz1), limit));
			// the boxes are padded by BVH_PAD, far more than the rounding of these slabs
			bool h0 = tn0 <= tf0, h1 = tn1 <= tf1;
			int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
			if (h0 && h1)
			{
				bool swap = tn1 < tn0;
				if (SSTACK)
					sstack[(uint32_t)(sp++) * stride] = (short)(swap ? c0 : c1);
				else
					stack[sp++] = swap ? c0 : c1;
				node = swap ? c1 : c0;
				continue;
			}
			if (h0) { node = c0; continue; }
			if (h1) { node = c1; continue; }
		}
		else
		{
			const int leaf = ~node;
			const float4 A = TRIS[4 * leaf + 0], B = TRIS[4 * leaf + 1], C = TRIS[4 * leaf + 2];
			float t;
			bool ok = true;
			if (need_flag) ok = (__float_as_uint(TRIS[4 * leaf + 3].w) & 1u) != 0;
			if (ok && ray_tri(o, d, tmax, V(A), V(B), V(C), t))
			{
				uint32_t orig = __float_as_uint(A.w);
				// closest hit; equal distances resolve to the lower triangle index, independent of traversal order
				if (t < best || (t == best && orig < bface))
				{
					best = t;
					limit = t;
					bface = orig;
					bbody = STATIC_BODY_BASE + __float_as_uint(B.w);
				}
			}
		}
		if (sp == 0) break;
		node = SSTACK ? (int)sstack[(uint32_t)(--sp) * stride] : stack[--sp];
	}
}

__device__ __forceinline__ void trace_bodies(const RayArgs &a, uint32_t world, uint32_t layers, bool need_flag, v3 o,
											 v3 d, float tmax, float &best, uint32_t &bbody, uint32_t &bface)
{
	if (world >= a.worlds) return;
	const uint32_t base = world * a.cap;
	for (uint32_t i = 0; i < a.cap; i++)
	{
		uint32_t f = a.flags[base + i];
		if (!(f & BF_ALIVE)) continue;
		uint32_t shape = (f >> BF_SHAPE_SHIFT) & 7u;
		if (shape == GPX_SHAPE_EMPTY) continue;
		if (!((layers >> ((f >> BF_LAYER_SHIFT) & 3u)) & 1u)) continue;
		if (need_flag && !((f >> BF_RAYFLAG_SHIFT) & 1u)) continue;
		float4 p = a.pos[base + i], p1 = a.prop1[base + i];
		float t;
		uint32_t face = 0;
		bool hit;
		if (shape == GPX_SHAPE_BOX)
			hit = ray_box(o, d, tmax, V(p), Q(a.quat[base + i]), V(p1), t, face);
		else
			hit = ray_sphere(o, d, tmax, V(p), p1.x, t);
		if (hit && t < best)
		{
			best = t;
			bbody = i;
			bface = face;
		}
	}
}

__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
	uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
	uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
				 "l"(gmem_src), "r"(bytes), "r"(b)
				 : "memory");
}

// SMEM = tree staged in shared memory (TMA bulk copy); otherwise read through L1/L2.
template <bool SMEM, bool SSTACK>
__global__ void __launch_bounds__(RAY_THREADS_MAX) k_raycast(RayArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const float4 *NODES = a.nodes;
	const float4 *TRIS = a.tris;
	short *sstack = nullptr;
	if (SSTACK) sstack = reinterpret_cast<short *>(smem_raw + (size_t)a.n_nodes * 64u + (size_t)a.n_tris * 64u) + threadIdx.x;
	if (SMEM)
	{
		float4 *s_nodes = reinterpret_cast<float4 *>(smem_raw);
		float4 *s_tris = s_nodes + 4ull * a.n_nodes;
		__shared__ __align__(8) uint64_t bar;
		const uint32_t node_bytes = a.n_nodes * 64u, tri_bytes = a.n_tris * 64u;
		uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
		if (threadIdx.x == 0)
		{
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();
		if (threadIdx.x == 0)
		{
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr),
						 "r"(node_bytes + tri_bytes)
						 : "memory");
			bulk_copy_g2s(s_nodes, a.nodes, node_bytes, &bar);
			bulk_copy_g2s(s_tris, a.tris, tri_bytes, &bar);
		}
		// everyone waits for phase 0 of the barrier
		uint32_t done = 0;
		while (!done)
		{
			asm volatile(
				"{\n"
				".reg .pred p;\n"
				"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
				"selp.u32 %0, 1, 0, p;\n"
				"}\n"
				: "=r"(done)
				: "r"(bar_addr)
				: "memory");
		}
		NODES = s_nodes;
		TRIS = s_tris;
	}
	const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride)
	{
		const float4 r0 = __ldg(&a.rays[2 * i]), r1 = __ldg(&a.rays[2 * i + 1]);
		const v3 o = V(r0), d = V(r1);
		const float tmax = r0.w;
		const uint32_t mask = __float_as_uint(r1.w);
		const uint32_t layers = mask & 0xFu;
		const bool need_flag = (mask & GPX_RAYMASK_REQUIRE_BLOCKS_LASERS) != 0;
		const uint32_t world = mask >> 16;
		float best = 3.0e38f;
		uint32_t bbody = GPX_INVALID_BODY, bface = GPX_INVALID_FACE;
		if (layers & 1u) trace_static<SSTACK>(NODES, TRIS, a.n_nodes, o, d, tmax, need_flag, best, bbody, bface, sstack, blockDim.x);
		if (layers & ~1u) trace_bodies(a, world, layers, need_flag, o, d, tmax, best, bbody, bface);
		float4 h;
		h.x = bbody == GPX_INVALID_BODY ? RAY_MISS_FRACTION : best / tmax;
		h.y = __uint_as_float(bbody);
		h.z = __uint_as_float(bface);
		h.w = __uint_as_float(world);
		a.hits[i] = h;
	}
}

// ---------------------------------------------------------------------------------------------------- sphere casts
// First contact of a sphere swept along a ray with a triangle = the earliest of: the sphere's lowest point reaching the
// triangle's plane inside the triangle; the centre's ray entering a cylinder of the sphere's radius around an edge; the
// centre's ray entering a sphere of that radius around a vertex.  Boxes: six faces, twelve edges, eight corners in the
// box's frame.  Same formulas, same order as the CPU restatement.

// the centre's ray (o, unit d) against the cylinder of radius r around the segment p0 -> p1 / the sphere around a vertex.
// Both quadratics are solved from an origin advanced to just outside the feature's bounding sphere: with the cast's own
// origin, tens of metres away, the terms that cancel are a million times the radius squared and fp32 leaves nothing of the
// answer.  A solution whose contact point is not at the radius is refused.
constexpr float SWEEP_RADIUS_TOL = 0.02f;
__device__ __forceinline__ bool sweep_edge(v3 o, v3 d, float tmax, float r, v3 p0, v3 p1, float &best, v3 &n)
{
	const v3 ed = p1 - p0;
	const float ee = dot(ed, ed);
	const float t0 = fmaxf(0.0f, dot(madd(p0, ed, 0.5f) - o, d) - ((0.5f * sqrtf(ee)) + r));
	const v3 o2 = madd(o, d, t0), m = o2 - p0;
	const float md = dot(m, ed), dd = dot(d, ed);
	const float a = ee - (dd * dd);
	if (!(a > (1.0e-5f * ee))) return false;  // along the edge: the spheres around its ends cover it
	const float k = dot(m, m) - (r * r);
	const float c = (ee * k) - (md * md);
	const float b = (ee * dot(m, d)) - (dd * md);
	const float disc = (b * b) - (a * c);
	if (disc < 0.0f) return false;
	const float t = (-b - sqrtf(disc)) / a, tt = t0 + t;
	if (!(t >= 0.0f && tt <= tmax && tt < best)) return false;
	const float s = md + (t * dd);
	if (s < 0.0f || s > ee) return false;
	const v3 q = madd(o2, d, t) - madd(p0, ed, s / ee);
	if (fabsf(len2(q) - (r * r)) > (SWEEP_RADIUS_TOL * (r * r))) return false;
	best = tt;
	n = q * (1.0f / r);
	return true;
}

__device__ __forceinline__ bool sweep_vertex(v3 o, v3 d, float tmax, float r, v3 p, float &best, v3 &n)
{
	const float t0 = fmaxf(0.0f, dot(p - o, d) - r);
	const v3 o2 = madd(o, d, t0), m = o2 - p;
	const float b = dot(m, d), c = dot(m, m) - (r * r);
	const float disc = (b * b) - c;
	if (disc < 0.0f) return false;
	const float t = -b - sqrtf(disc), tt = t0 + t;
	if (!(t >= 0.0f && tt <= tmax && tt < best)) return false;
	const v3 q = madd(o2, d, t) - p;
	if (fabsf(len2(q) - (r * r)) > (SWEEP_RADIUS_TOL * (r * r))) return false;
	best = tt;
	n = q * (1.0f / r);
	return true;
}

static __device__ __noinline__ bool sweep_sphere_tri(v3 o, v3 d, float tmax, float r, v3 a, v3 b, v3 c, float &tout, v3 &nout)
{
	// overlapping at the start
	const v3 cp = closest_on_tri(o, a, b, c);
	const v3 dv = o - cp;
	const float d2 = len2(dv);
	if (d2 <= (r * r))
	{
		tout = 0.0f;
		nout = d2 > 1.0e-12f ? dv * (1.0f / sqrtf(d2)) : -d;
		return true;
	}
	const v3 e1 = b - a, e2 = c - a;
	const v3 nn = cross(e1, e2);
	const float l2 = len2(nn);
	float best = 3.0e38f;
	v3 bn = V(0.0f, 0.0f, 0.0f);
	if (l2 > 1.0e-20f)
	{
		v3 n = nn * (1.0f / sqrtf(l2));
		float s0 = dot(n, o - a), nd = dot(n, d);
		if (s0 < 0.0f)
		{
			n = -n;
			s0 = -s0;
			nd = -nd;
		}
		if (nd < 0.0f && s0 > r)
		{
			const float t = (r - s0) / nd;
			if (t <= tmax)
			{
				// where the sphere touches the plane; inside the triangle?
				const v3 p = madd(o, d, t) - (n * r);
				const v3 ca = cross(b - a, p - a), cb = cross(c - b, p - b), cc = cross(a - c, p - c);
				const float sa = dot(ca, nn), sb = dot(cb, nn), sc = dot(cc, nn);
				if (sa >= 0.0f && sb >= 0.0f && sc >= 0.0f)
				{
					tout = t;
					nout = n;
					return true;
				}
			}
		}
	}
	if (!(r > 0.0f)) return false;  // a ray only meets the face
	bool hit = false;
	hit |= sweep_edge(o, d, tmax, r, a, b, best, bn);
	hit |= sweep_edge(o, d, tmax, r, b, c, best, bn);
	hit |= sweep_edge(o, d, tmax, r, c, a, best, bn);
	hit |= sweep_vertex(o, d, tmax, r, a, best, bn);
	hit |= sweep_vertex(o, d, tmax, r, b, best, bn);
	hit |= sweep_vertex(o, d, tmax, r, c, best, bn);
	if (!hit) return false;
	tout = best;
	nout = bn;
	return true;
}

__device__ __forceinline__ bool sweep_sphere_sphere(v3 o, v3 d, float tmax, float r, v3 x, float R, float &tout, v3 &nout)
{
	const float rr = r + R;
	const v3 m = o - x;
	const float mm = dot(m, m);
	if (mm <= (rr * rr))
	{
		tout = 0.0f;
		nout = mm > 1.0e-12f ? m * (1.0f / sqrtf(mm)) : -d;
		return true;
	}
	float best = 3.0e38f;
	v3 n;
	if (!sweep_vertex(o, d, tmax, rr, x, best, n)) return false;
	tout = best;
	nout = n;
	return true;
}

// in the box's frame; the normal goes back to the world; face = the box face the normal leans to (2k: -axis, 2k+1: +axis)
static __device__ __noinline__ bool sweep_sphere_box(v3 o, v3 d, float tmax, float r, v3 x, q4 q, v3 he, float &tout, v3 &nout,
													 uint32_t &face)
{
	const m33 R = qmat(q);
	const v3 lo = mtmul(R, o - x), ld = mtmul(R, d);
	float best = 3.0e38f;
	v3 bn = V(0.0f, 0.0f, 0.0f);
	bool hit = false;
	const v3 cl = V(fminf(fmaxf(lo.x, -he.x), he.x), fminf(fmaxf(lo.y, -he.y), he.y), fminf(fmaxf(lo.z, -he.z), he.z));
	const v3 dv = lo - cl;
	const float d2 = len2(dv);
	if (d2 <= (r * r))
	{
		best = 0.0f;
		bn = d2 > 1.0e-12f ? dv * (1.0f / sqrtf(d2)) : -ld;
		hit = true;
	}
	else
	{
		// faces: the plane he_k + r on the side the centre comes from, hit inside the face's rectangle
		for (int k = 0; k < 3; k++)
		{
			const float ok = get(lo, k), dk = get(ld, k), hk = get(he, k);
			const float side = ok >= 0.0f ? 1.0f : -1.0f;
			if ((dk * side) >= 0.0f || (ok * side) <= (hk + r)) continue;  // moving away, or not outside this slab
			const float t = (((hk + r) * side) - ok) / dk;
			if (!(t >= 0.0f && t <= tmax && t < best)) continue;
			const v3 p = madd(lo, ld, t);
			const int u = (k + 1) % 3, v = (k + 2) % 3;
			if (fabsf(get(p, u)) <= get(he, u) && fabsf(get(p, v)) <= get(he, v))
			{
				best = t;
				bn = V(k == 0 ? side : 0.0f, k == 1 ? side : 0.0f, k == 2 ? side : 0.0f);
				hit = true;
			}
		}
		// twelve edges (four along each axis), eight corners (a ray, radius 0, only meets the faces)
		for (int k = 0; k < 3 && r > 0.0f; k++)
			for (int su = -1; su <= 1; su += 2)
				for (int sv = -1; sv <= 1; sv += 2)
				{
					const int u = (k + 1) % 3, v = (k + 2) % 3;
					float p0[3], p1[3];
					p0[k] = -get(he, k); p1[k] = get(he, k);
					p0[u] = p1[u] = (float)su * get(he, u);
					p0[v] = p1[v] = (float)sv * get(he, v);
					hit |= sweep_edge(lo, ld, tmax, r, V(p0[0], p0[1], p0[2]), V(p1[0], p1[1], p1[2]), best, bn);
				}
		for (int sx = -1; sx <= 1 && r > 0.0f; sx += 2)
			for (int sy = -1; sy <= 1; sy += 2)
				for (int sz = -1; sz <= 1; sz += 2)
					hit |= sweep_vertex(lo, ld, tmax, r, V((float)sx * he.x, (float)sy * he.y, (float)sz * he.z), best, bn);
	}
	if (!hit) return false;
	const float ax = fabsf(bn.x), ay = fabsf(bn.y), az = fabsf(bn.z);
	int k = 0;
	if (ay > ax) k = 1;
	if (az > (k == 0 ? ax : ay)) k = 2;
	face = (uint32_t)(2 * k + (get(bn, k) > 0.0f ? 1 : 0));
	tout = best;
	nout = mmul(R, bn);
	return true;
}

struct CastArgs
{
	const float4 *nodes;
	const float4 *tris;
	uint32_t n_nodes;
	const float4 *casts;  // 3 per cast: origin|tmax, dir|mask, radius|-
	float4 *hits;         // 2 per cast: fraction body face world | normal -
	unsigned long long n;
	const float4 *pos, *quat, *prop1;
	const uint32_t *flags;
	uint32_t worlds, cap;
};

// One thread per cast: the rays' tree with every box grown by the radius, the exact sweep at the leaves (a triangle cut
// into several leaves is simply tested again), then the world's bodies.
__global__ void __launch_bounds__(128) k_spherecast(CastArgs a)
{
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.n) return;
	const float4 c0 = __ldg(&a.casts[3 * i]), c1 = __ldg(&a.casts[3 * i + 1]), c2 = __ldg(&a.casts[3 * i + 2]);
	const v3 o = V(c0), d = V(c1);
	const float tmax = c0.w, r = c2.x;
	const uint32_t mask = __float_as_uint(c1.w), layers = mask & 0xFu, world = mask >> 16;
	const bool need_flag = (mask & GPX_RAYMASK_REQUIRE_BLOCKS_LASERS) != 0;
	float best = 3.0e38f;
	uint32_t bbody = GPX_INVALID_BODY, bface = GPX_INVALID_FACE;
	v3 bn = V(0.0f, 0.0f, 0.0f);
	if ((layers & 1u) && a.n_nodes > 0)
	{
		const float tiny = 1.0e-20f;
		const float idx = 1.0f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x));
		const float idy = 1.0f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y));
		const float idz = 1.0f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z));
		const float pad = r + 1.0e-3f;  // the swept sphere reaches `r` beyond the centre's ray
		int stack[STACK_DEPTH];
		int sp = 0, node = 0;
		while (true)
		{
			if (node >= 0)
			{
				const float4 n0 = __ldg(&a.nodes[4 * node + 0]), n1 = __ldg(&a.nodes[4 * node + 1]), n2 = __ldg(&a.nodes[4 * node + 2]),
							 n3 = __ldg(&a.nodes[4 * node + 3]);
				const float limit = fminf(best, tmax);
				bool h[2];
#pragma unroll
				for (int ch = 0; ch < 2; ch++)
				{
					const float4 bx = ch ? n1 : n0;
					const float zl = ch ? n2.z : n2.x, zh = ch ? n2.w : n2.y;
					const float ax0 = ((bx.x - pad) - o.x) * idx, ax1 = ((bx.y + pad) - o.x) * idx;
					const float ay0 = ((bx.z - pad) - o.y) * idy, ay1 = ((bx.w + pad) - o.y) * idy;
					const float az0 = ((zl - pad) - o.z) * idz, az1 = ((zh + pad) - o.z) * idz;
					const float tn = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fmaxf(fminf(az0, az1), 0.0f));
					const float tf = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fminf(fmaxf(az0, az1), limit));
					h[ch] = tn <= tf;
				}
				const int k0 = __float_as_int(n3.x), k1 = __float_as_int(n3.y);
				if (h[0] && h[1])
				{
					stack[sp++] = k1;
					node = k0;
					continue;
				}
				if (h[0]) { node = k0; continue; }
				if (h[1]) { node = k1; continue; }
			}
			else
			{
				const int leaf = ~node;
				const float4 A = __ldg(&a.tris[4 * leaf + 0]), B = __ldg(&a.tris[4 * leaf + 1]), C = __ldg(&a.tris[4 * leaf + 2]);
				bool ok = true;
				if (need_flag) ok = (__float_as_uint(__ldg(&a.tris[4 * leaf + 3]).w) & 1u) != 0;
				float t;
				v3 n;
				if (ok && sweep_sphere_tri(o, d, tmax, r, V(A), V(B), V(C), t, n))
				{
					const uint32_t orig = __float_as_uint(A.w);
					// first contact; equal times resolve to the lower triangle index, independent of traversal order
					if (t < best || (t == best && orig < bface))
					{
						best = t;
						bn = n;
						bface = orig;
						bbody = STATIC_BODY_BASE + __float_as_uint(B.w);
					}
				}
			}
			if (sp == 0) break;
			node = stack[--sp];
		}
	}
	if ((layers & ~1u) && world < a.worlds)
	{
		const uint32_t base = world * a.cap;
		for (uint32_t b = 0; b < a.cap; b++)
		{
			const uint32_t f = a.flags[base + b];
			if (!(f & BF_ALIVE)) continue;
			const uint32_t shape = (f >> BF_SHAPE_SHIFT) & 7u;
			if (shape == GPX_SHAPE_EMPTY) continue;
			if (!((layers >> ((f >> BF_LAYER_SHIFT) & 3u)) & 1u)) continue;
			if (need_flag && !((f >> BF_RAYFLAG_SHIFT) & 1u)) continue;
			const float4 p = a.pos[base + b], p1 = a.prop1[base + b];
			float t;
			v3 n;
			uint32_t face = 0;
			const bool hit = shape == GPX_SHAPE_BOX ? sweep_sphere_box(o, d, tmax, r, V(p), Q(a.quat[base + b]), V(p1), t, n, face)
													: sweep_sphere_sphere(o, d, tmax, r, V(p), p1.x, t, n);
			if (hit && t < best)
			{
				best = t;
				bn = n;
				bbody = b;
				bface = face;
			}
		}
	}
	const bool miss = bbody == GPX_INVALID_BODY;
	a.hits[2 * i] = make_float4(miss ? RAY_MISS_FRACTION : best / tmax, __uint_as_float(bbody), __uint_as_float(bface), __uint_as_float(world));
	a.hits[2 * i + 1] = miss ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : make_float4(bn.x, bn.y, bn.z, 0.0f);
}

int launch_spherecast(gpx_world *w, const void *d_casts, uint64_t n, void *d_hits)
{
	if (n == 0) return GPX_OK;
	CastArgs a;
	a.nodes = w->sd.ray_nodes;
	a.tris = w->sd.ray_tri;
	a.n_nodes = w->sd.n_ray_nodes;
	a.casts = (const float4 *)d_casts;
	a.hits = (float4 *)d_hits;
	a.n = n;
	a.pos = w->bs.pos;
	a.quat = w->bs.quat;
	a.prop1 = w->bs.prop1;
	a.flags = w->bs.flags;
	a.worlds = w->W;
	a.cap = w->cap;
	k_spherecast<<<(unsigned)((n + 127) / 128), 128, 0, w->stream>>>(a);
	count_launch();
	GPX_CUDA(cudaGetLastError());
	return GPX_OK;
}

int launch_raycast(gpx_world *w, const void *d_rays, uint64_t n, void *d_hits)
{
	if (n == 0) return GPX_OK;
	RayArgs a;

	a.nodes = w->sd.ray_nodes;
	a.tris = w->sd.ray_tri;
	a.n_nodes = w->sd.n_ray_nodes;
	a.n_tris = w->sd.n_ray_leaves;
	a.rays = (const float4 *)d_rays;
	a.hits = (float4 *)d_hits;
	a.n = n;
	a.pos = w->bs.pos;
	a.quat = w->bs.quat;
	a.prop1 = w->bs.prop1;
	a.flags = w->bs.flags;
	a.worlds = w->W;
	a.cap = w->cap;
	int dev = w->device, sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const size_t tree_bytes = (size_t)a.n_nodes * 64u + (size_t)a.n_tris * 64u;
	const bool smem = a.n_nodes > 0 && tree_bytes <= 200u * 1024u;
	unsigned long long want = (n + RAY_THREADS - 1) / RAY_THREADS;
	if (smem)
	{
		// function attributes are per device, and one process may hold worlds on several: remember per device
		static std::atomic<bool> attr_set[GPX_MAX_DEVICES];
		if (dev < 0 || dev >= GPX_MAX_DEVICES || !attr_set[dev].load(std::memory_order_acquire))
		{
			GPX_CUDA(cudaFuncSetAttribute(k_raycast<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
			GPX_CUDA(cudaFuncSetAttribute(k_raycast<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
			if (dev >= 0 && dev < GPX_MAX_DEVICES) attr_set[dev].store(true, std::memory_order_release);
		}
		// the traversal stack joins the tree in shared memory when the tree is shallow enough (the SAH build knows its depth)
		// and its node codes fit 16 bits: 64 bytes per thread
		const bool sstack = w->sd.ray_depth < (uint32_t)SSTACK_DEPTH && a.n_nodes < 32768u && a.n_tris < 32768u;
		const size_t per_thread = sstack ? sizeof(short) * SSTACK_DEPTH : 0;
		int per_sm = 4, threads = 256;
		// about 1024 threads per SM whatever the number of tree copies that fit
		for (;; per_sm--)
		{
			threads = per_sm >= 4 ? 256 : (per_sm == 3 ? 320 : (per_sm == 2 ? 512 : RAY_THREADS_MAX));
			if (per_sm == 1 || (tree_bytes + per_thread * threads + 1024u) * per_sm <= 227u * 1024u) break;
		}
		size_t smem_bytes = tree_bytes + per_thread * threads;
		const bool use_sstack = sstack && smem_bytes <= 226u * 1024u;
		if (!use_sstack) smem_bytes = tree_bytes;
		want = (n + threads - 1) / threads;
		unsigned long long grid = (unsigned long long)sms * per_sm;
		if (grid > want) grid = want;
		if (use_sstack)
			k_raycast<true, true><<<(unsigned)grid, threads, smem_bytes, w->stream>>>(a);
		else
			k_raycast<true, false><<<(unsigned)grid, threads, smem_bytes, w->stream>>>(a);
	}
	else
	{
		unsigned long long grid = (unsigned long long)sms * 8;
		if (grid > want) grid = want;
		k_raycast<false, false><<<(unsigned)grid, RAY_THREADS, 0, w->stream>>>(a);
	}
	count_launch();
	GPX_CUDA(cudaGetLastError());
	return GPX_OK;
}

}  // namespace gpx
