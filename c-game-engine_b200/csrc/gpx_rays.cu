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
