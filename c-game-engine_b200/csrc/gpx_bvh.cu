// gpx_bvh.cu — LBVH over the map's static collision triangles, built on the device.
//
// Replaces what JPH_PhysicsSystem_OptimizeBroadPhase + MeshShape tree construction do for the reference at map
// load (engine/src/assets/MapLoader.c:200-273): the zlib-decompressed triangles are uploaded once, Morton-sorted,
// linked into a binary radix tree (Karras 2012) and refitted bottom-up.  Output is two flat float4 arrays laid out
// for the traversal kernels: 64-byte nodes that carry BOTH children's boxes, and 64-byte triangle records in leaf
// order.  The whole tree of a shipped map (<= 2k triangles) is <= 250 KB and is read through shared memory / L2.
#include <cooperative_groups.h>
#include <cfloat>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "gpx_internal.h"
#include "gpx_math.cuh"

namespace cg = cooperative_groups;

namespace gpx {

__device__ __forceinline__ uint32_t expand_bits10(uint32_t v)
{
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}

// key = 30-bit Morton code of the centroid (high word) | original triangle index (low word): unique, so the radix
// tree never has to break ties.
// `refb` (6 floats per primitive: lo xyz, hi xyz) is given when the primitives are split references to triangles (the
// ray tree); the key then comes from the centre of the reference's box instead of the triangle's centroid.
__global__ void k_morton_keys(const float *__restrict__ tris, const float *__restrict__ refb, uint32_t n, uint32_t n_pad, float3 lo,
							  float3 inv_ext, unsigned long long *__restrict__ keys)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_pad) return;
	if (i >= n)
	{
		keys[i] = ~0ull;
		return;
	}
	const float *t = tris + 9ull * i;
	float cx, cy, cz;
	if (refb)
	{
		const float *b = refb + 6ull * i;
		cx = 0.5f * (b[0] + b[3]);
		cy = 0.5f * (b[1] + b[4]);
		cz = 0.5f * (b[2] + b[5]);
	}
	else
	{
		cx = (t[0] + t[3] + t[6]) * (1.0f / 3.0f);
		cy = (t[1] + t[4] + t[7]) * (1.0f / 3.0f);
		cz = (t[2] + t[5] + t[8]) * (1.0f / 3.0f);
	}
	float fx = fminf(fmaxf((cx - lo.x) * inv_ext.x * 1024.0f, 0.0f), 1023.0f);
	float fy = fminf(fmaxf((cy - lo.y) * inv_ext.y * 1024.0f, 0.0f), 1023.0f);
	float fz = fminf(fmaxf((cz - lo.z) * inv_ext.z * 1024.0f, 0.0f), 1023.0f);
	uint32_t m = (expand_bits10((uint32_t)fx) << 2) | (expand_bits10((uint32_t)fy) << 1) | expand_bits10((uint32_t)fz);
	keys[i] = ((unsigned long long)m << 32) | i;
}

// One compare-exchange pass of a bitonic network over a power-of-two key array.
__global__ void k_bitonic_pass(unsigned long long *keys, uint32_t n_pad, uint32_t j, uint32_t k)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_pad) return;
	uint32_t p = i ^ j;
	if (p > i)
	{
		unsigned long long a = keys[i], b = keys[p];
		bool up = (i & k) == 0;
		if ((a > b) == up)
		{
			keys[i] = b;
			keys[p] = a;
		}
	}
}

// All passes with stride < TILE of one merge stage run inside a CTA's shared memory (TILE keys per CTA).
template <uint32_t TILE>
__global__ void __launch_bounds__(1024) k_bitonic_smem(unsigned long long *keys, uint32_t n_pad, uint32_t k_first, uint32_t k_last,
														uint32_t j_start_for_first)
{
	extern __shared__ __align__(8) unsigned char sort_raw[];
	unsigned long long *s = reinterpret_cast<unsigned long long *>(sort_raw);
	const uint32_t base = blockIdx.x * TILE;
	for (uint32_t t = threadIdx.x; t < TILE; t += blockDim.x) s[t] = (base + t < n_pad) ? keys[base + t] : ~0ull;
	__syncthreads();
	for (uint32_t k = k_first; k <= k_last; k <<= 1)
	{
		uint32_t j0 = (k == k_first) ? j_start_for_first : (k >> 1);
		for (uint32_t j = j0; j > 0; j >>= 1)
		{
			for (uint32_t t = threadIdx.x; t < TILE; t += blockDim.x)
			{
				uint32_t i = t, p = t ^ j;
				if (p > i)
				{
					unsigned long long a = s[i], b = s[p];
					bool up = ((base + i) & k) == 0;
					if ((a > b) == up)
					{
						s[i] = b;
						s[p] = a;
					}
				}
			}
			__syncthreads();
		}
	}
	for (uint32_t t = threadIdx.x; t < TILE; t += blockDim.x)
		if (base + t < n_pad) keys[base + t] = s[t];
}

template <uint32_t TILE>
static void bitonic_sort_tiles(unsigned long long *d_keys, uint32_t n_pad, cudaStream_t st)
{
	const uint32_t tb = 256;
	const size_t smem = (size_t)TILE * sizeof(unsigned long long);
	static bool configured = false;
	if (!configured)
	{
		cudaFuncSetAttribute(k_bitonic_smem<TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		configured = true;
	}
	if (n_pad <= TILE)
	{
		k_bitonic_smem<TILE><<<1, 1024, smem, st>>>(d_keys, n_pad, 2, n_pad, 1);
		count_launch();
		return;
	}
	k_bitonic_smem<TILE><<<n_pad / TILE, 1024, smem, st>>>(d_keys, n_pad, 2, TILE, 1);
	count_launch();
	for (uint32_t k = 2u * TILE; k <= n_pad; k <<= 1)
	{
		for (uint32_t j = k >> 1; j >= TILE; j >>= 1)
		{
			k_bitonic_pass<<<(n_pad + tb - 1) / tb, tb, 0, st>>>(d_keys, n_pad, j, k);
			count_launch();
		}
		k_bitonic_smem<TILE><<<n_pad / TILE, 1024, smem, st>>>(d_keys, n_pad, k, k, TILE / 2);
		count_launch();
	}
}

// Ascending sort of a power-of-two array of 64-bit keys: strides >= the tile go through global memory, the rest of each
// merge stage runs inside one CTA's shared memory.  Large arrays (the wide-world broadphase, ~10^5 keys every sub-step)
// use 8192-key tiles: 15 launches instead of 31 for 131 072 keys.
void bitonic_sort_u64(unsigned long long *d_keys, uint32_t n_pad, cudaStream_t st)
{
	if (n_pad >= 65536u) bitonic_sort_tiles<8192u>(d_keys, n_pad, st);
	else bitonic_sort_tiles<2048u>(d_keys, n_pad, st);
}

// ---- LSD radix sort, 8 bits per pass, stable.  Used where a key array is re-sorted every sub-step (the wide-world
// broadphase): three launches per pass — per-block digit histograms, one scan over (digit, block), a scatter that
// ranks keys inside each warp with match_any so equal digits keep their order.
constexpr uint32_t RADIX_BLOCK = 256, RADIX_ITEMS = 4, RADIX_TILE = RADIX_BLOCK * RADIX_ITEMS;  // 1024 keys per block

// `fused` (a few hundred blocks at most): counters are stored block-major and every scatter block sums the table itself —
// 256 threads read it coalesced out of L2 in a microsecond, where the one-block scan kernel in between cost ten.
__device__ __forceinline__ void radix_hist_block(const unsigned long long *__restrict__ keys, uint32_t shift,
												 uint32_t *__restrict__ ghist, bool fused)
{
	__shared__ uint32_t hist[256];
	hist[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t base = blockIdx.x * RADIX_TILE;
#pragma unroll
	for (uint32_t i = 0; i < RADIX_ITEMS; i++)
		atomicAdd(&hist[(uint32_t)(keys[base + i * RADIX_BLOCK + threadIdx.x] >> shift) & 255u], 1u);
	__syncthreads();
	if (fused)
		ghist[blockIdx.x * 256u + threadIdx.x] = hist[threadIdx.x];
	else
		ghist[threadIdx.x * gridDim.x + blockIdx.x] = hist[threadIdx.x];  // digit-major: the scan order of the scatter
}
__global__ void __launch_bounds__(RADIX_BLOCK) k_radix_hist(const unsigned long long *__restrict__ keys, uint32_t shift,
															uint32_t *__restrict__ ghist, bool fused,
															const uint32_t *__restrict__ only_if)
{
	if (only_if && *only_if == 0u) return;  // the caller found the array sorted already
	radix_hist_block(keys, shift, ghist, fused);
}

// exclusive scan of `n` counters by one block (n = 256 digits x blocks, a few 10^4): each warp owns a contiguous
// chunk and walks it 32 entries at a time (coalesced); the loads of the first pass are independent, the second pass
// re-reads from cache
__global__ void __launch_bounds__(1024) k_radix_scan(uint32_t *__restrict__ ghist, uint32_t n, const uint32_t *__restrict__ only_if)
{
	__shared__ uint32_t warp_sum[32];
	if (only_if && *only_if == 0u) return;
	const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	const uint32_t chunk = (((n + 31u) / 32u) + 31u) & ~31u;
	const uint32_t lo = wid * chunk, hi = min(lo + chunk, n);
	uint32_t sum = 0;
#pragma unroll 8
	for (uint32_t i = lo + lane; i < hi; i += 32u) sum += ghist[i];
	for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
	if (lane == 0) warp_sum[wid] = sum;
	__syncthreads();
	if (wid == 0)
	{
		const uint32_t w = warp_sum[lane];
		uint32_t wi = w;
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, o);
			if (lane >= (uint32_t)o) wi += v;
		}
		warp_sum[lane] = wi - w;  // exclusive prefix of the warp totals
	}
	__syncthreads();
	uint32_t run = warp_sum[wid];
	for (uint32_t base = lo; base < hi; base += 32u)  // uniform per warp
	{
		const uint32_t i = base + lane;
		const uint32_t c = i < hi ? ghist[i] : 0u;
		uint32_t incl = c;
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
			if (lane >= (uint32_t)o) incl += v;
		}
		if (i < hi) ghist[i] = run + (incl - c);
		run += __shfl_sync(0xFFFFFFFFu, incl, 31);
	}
}

__device__ __forceinline__ void radix_scatter_block(const unsigned long long *__restrict__ keys,
													unsigned long long *__restrict__ out, uint32_t shift,
													const uint32_t *__restrict__ ghist, bool fused)
{
	__shared__ uint32_t wcnt[RADIX_BLOCK / 32][256];
	__shared__ uint32_t wtot[RADIX_BLOCK / 32];
	const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
	for (uint32_t k = threadIdx.x; k < (RADIX_BLOCK / 32) * 256; k += RADIX_BLOCK) (&wcnt[0][0])[k] = 0;
	__syncthreads();
	// a warp owns 32 * RADIX_ITEMS consecutive keys; key order inside it is (item, lane)
	const uint32_t base = blockIdx.x * RADIX_TILE + w * (32u * RADIX_ITEMS);
	unsigned long long key[RADIX_ITEMS];
	uint32_t local[RADIX_ITEMS];
#pragma unroll
	for (uint32_t i = 0; i < RADIX_ITEMS; i++)
	{
		key[i] = keys[base + i * 32u + lane];
		const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
		const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
		const uint32_t before = wcnt[w][d];
		__syncwarp();
		if (lane == (uint32_t)__ffs((int)peers) - 1u) wcnt[w][d] = before + (uint32_t)__popc(peers);
		__syncwarp();
		local[i] = before + (uint32_t)__popc(peers & ((1u << lane) - 1u));
	}
	// thread = digit: where this block's keys with that digit start in the output
	uint32_t run;
	if (fused)
	{
		// keys with a smaller digit anywhere + keys with this digit in earlier blocks
		uint32_t below = 0, all = 0;
#pragma unroll 8
		for (uint32_t b = 0; b < gridDim.x; b++)
		{
			const uint32_t c = ghist[b * 256u + threadIdx.x];
			all += c;
			if (b < blockIdx.x) below += c;
		}
		uint32_t incl = all;
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
			if (lane >= (uint32_t)o) incl += v;
		}
		if (lane == 31u) wtot[w] = incl;
		__syncthreads();
		uint32_t before = 0;
		for (uint32_t ww = 0; ww < w; ww++) before += wtot[ww];
		run = before + (incl - all) + below;
	}
	else
	{
		run = ghist[threadIdx.x * gridDim.x + blockIdx.x];
		__syncthreads();
	}
	{
		// turn the per-warp counts into global start positions, warp after warp
		const uint32_t d = threadIdx.x;
		for (uint32_t ww = 0; ww < RADIX_BLOCK / 32; ww++)
		{
			const uint32_t c = wcnt[ww][d];
			wcnt[ww][d] = run;
			run += c;
		}
	}
	__syncthreads();
#pragma unroll
	for (uint32_t i = 0; i < RADIX_ITEMS; i++) out[wcnt[w][(uint32_t)(key[i] >> shift) & 255u] + local[i]] = key[i];
}
__global__ void __launch_bounds__(RADIX_BLOCK) k_radix_scatter(const unsigned long long *__restrict__ keys,
															   unsigned long long *__restrict__ out, uint32_t shift,
															   const uint32_t *__restrict__ ghist, bool fused,
															   const uint32_t *__restrict__ only_if)
{
	if (only_if && *only_if == 0u) return;
	radix_scatter_block(keys, out, shift, ghist, fused);
}

// The whole sort in ONE cooperative launch (at most 256 blocks, all resident): used where the sort is a rarely needed
// fall-back behind `only_if`, so that not needing it costs one empty launch instead of two per pass.
__global__ void __launch_bounds__(RADIX_BLOCK) k_radix_sort_coop(unsigned long long *keys, unsigned long long *tmp,
																 uint32_t *ghist, uint32_t first_bit, const uint32_t *only_if)
{
	if (only_if && *only_if == 0u) return;  // the same answer in every block
	cg::grid_group grid = cg::this_grid();
	unsigned long long *src = keys, *dst = tmp;
	int passes = 0;
	for (uint32_t sh = first_bit; sh < 64u; sh += 8u) passes++;
	for (int p = 0; p < passes + (passes & 1); p++)
	{
		const uint32_t sh = min(first_bit + 8u * (uint32_t)p, 56u);
		radix_hist_block(src, sh, ghist, true);
		grid.sync();
		radix_scatter_block(src, dst, sh, ghist, true);
		grid.sync();
		unsigned long long *t = src;
		src = dst;
		dst = t;
	}
}

// In-place exclusive prefix sum of `n` counters (one block; the radix sort's scan).
void exclusive_scan_u32(uint32_t *d, uint32_t n, cudaStream_t st)
{
	k_radix_scan<<<1, 1024, 0, st>>>(d, n, nullptr);
	count_launch();
}

// Ascending stable sort of bits [first_bit, 64) of `n` keys (n a multiple of 1024); bits below first_bit keep their
// input order.  `tmp` holds n keys, `ghist` 256 * n / 1024 counters.  The result ends in `d_keys`.
// `only_if` (optional): a device word; the passes do nothing when it is zero at the time they run.
void radix_sort_u64(unsigned long long *d_keys, unsigned long long *tmp, uint32_t *ghist, uint32_t n, uint32_t first_bit,
					cudaStream_t st, const uint32_t *only_if)
{
	const uint32_t blocks = n / RADIX_TILE;
	if (only_if && blocks <= 256u)
	{
		void *params[] = {&d_keys, &tmp, &ghist, &first_bit, &only_if};
		if (cudaLaunchCooperativeKernel((const void *)k_radix_sort_coop, dim3(blocks), dim3(RADIX_BLOCK), params, 0, st) == cudaSuccess)
		{
			count_launch();
			return;
		}
		cudaGetLastError();  // not resident on this device: the per-pass launches below
	}
	uint32_t shifts[9];
	int passes = 0;
	for (uint32_t sh = first_bit; sh < 64u; sh += 8u) shifts[passes++] = sh > 56u ? 56u : sh;
	if (passes & 1) shifts[passes++] = 56u;  // an even number of passes leaves the result in d_keys (a repeat is harmless)
	unsigned long long *src = d_keys, *dst = tmp;
	const bool fused = blocks <= 256u;
	for (int p = 0; p < passes; p++)
	{
		k_radix_hist<<<blocks, RADIX_BLOCK, 0, st>>>(src, shifts[p], ghist, fused, only_if);
		if (!fused) k_radix_scan<<<1, 1024, 0, st>>>(ghist, 256u * blocks, only_if);
		k_radix_scatter<<<blocks, RADIX_BLOCK, 0, st>>>(src, dst, shifts[p], ghist, fused, only_if);
		count_launch(fused ? 2 : 3);
		unsigned long long *t = src;
		src = dst;
		dst = t;
	}
}

__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j)
{
	if (j < 0 || j >= n) return -1;
	return __clzll(keys[i] ^ keys[j]);
}

// Karras 2012: internal node i covers a key range found by binary search on common-prefix length.
// child >= 0: internal node; child < 0: leaf ~child.
__global__ void k_hierarchy(const unsigned long long *__restrict__ keys, int n, int2 *__restrict__ children,
							int *__restrict__ parent_internal, int *__restrict__ parent_leaf)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n - 1) return;
	int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
	int dmin = delta(keys, n, i, i - d);
	int lmax = 2;
	while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
	int l = 0;
	for (int t = lmax >> 1; t >= 1; t >>= 1)
		if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
	int j = i + l * d;
	int dnode = delta(keys, n, i, j);
	int s = 0;
	for (int t = (l + 1) >> 1;; t = (t + 1) >> 1)
	{
		if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
		if (t <= 1) break;
	}
	int gamma = i + s * d + min(d, 0);
	int lo = min(i, j), hi = max(i, j);
	int2 c;
	if (lo == gamma)
	{
		c.x = ~gamma;
		parent_leaf[gamma] = i;
	}
	else
	{
		c.x = gamma;
		parent_internal[gamma] = i;
	}
	if (hi == gamma + 1)
	{
		c.y = ~(gamma + 1);
		parent_leaf[gamma + 1] = i;
	}
	else
	{
		c.y = gamma + 1;
		parent_internal[gamma + 1] = i;
	}
	children[i] = c;
	if (i == 0) parent_internal[0] = -1;
}

__device__ __forceinline__ void tri_bounds(const float *t, float3 &lo, float3 &hi)
{
	lo.x = fminf(t[0], fminf(t[3], t[6]));
	lo.y = fminf(t[1], fminf(t[4], t[7]));
	lo.z = fminf(t[2], fminf(t[5], t[8]));
	hi.x = fmaxf(t[0], fmaxf(t[3], t[6]));
	hi.y = fmaxf(t[1], fmaxf(t[4], t[7]));
	hi.z = fmaxf(t[2], fmaxf(t[5], t[8]));
}

// bounds of primitive `p`: its reference box when the tree is built over split references, else the triangle's
__device__ __forceinline__ void prim_bounds(const float *tris, const float *refb, uint32_t p, float3 &lo, float3 &hi)
{
	if (refb)
	{
		const float *b = refb + 6ull * p;
		lo = make_float3(b[0], b[1], b[2]);
		hi = make_float3(b[3], b[4], b[5]);
	}
	else
		tri_bounds(tris + 9ull * p, lo, hi);
}

// Bottom-up refit: the second thread to reach a node owns it (its sibling subtree is complete and fenced).
__global__ void k_refit(const float *__restrict__ tris, const float *__restrict__ refb, const unsigned long long *__restrict__ keys, int n,
						const int2 *__restrict__ children, const int *__restrict__ parent_internal,
						const int *__restrict__ parent_leaf, float *lo_out, float *hi_out, int *visit)
{
	int leaf = blockIdx.x * blockDim.x + threadIdx.x;
	if (leaf >= n) return;
	int node = parent_leaf[leaf];
	while (node >= 0)
	{
		__threadfence();
		if (atomicAdd(&visit[node], 1) == 0) return;
		float3 lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
		int2 c = children[node];
		int cc[2] = {c.x, c.y};
#pragma unroll
		for (int k = 0; k < 2; k++)
		{
			float3 l, h;
			if (cc[k] < 0)
				prim_bounds(tris, refb, (uint32_t)(keys[~cc[k]] & 0xFFFFFFFFull), l, h);
			else
			{
				volatile float *vl = lo_out + 3 * cc[k], *vh = hi_out + 3 * cc[k];
				l = make_float3(vl[0], vl[1], vl[2]);
				h = make_float3(vh[0], vh[1], vh[2]);
			}
			lo.x = fminf(lo.x, l.x); lo.y = fminf(lo.y, l.y); lo.z = fminf(lo.z, l.z);
			hi.x = fmaxf(hi.x, h.x); hi.y = fmaxf(hi.y, h.y); hi.z = fmaxf(hi.z, h.z);
		}
		lo_out[3 * node + 0] = lo.x; lo_out[3 * node + 1] = lo.y; lo_out[3 * node + 2] = lo.z;
		hi_out[3 * node + 0] = hi.x; hi_out[3 * node + 1] = hi.y; hi_out[3 * node + 2] = hi.z;
		node = parent_internal[node];
	}
}

// Emit traversal nodes (both children's padded boxes per node) and leaf-ordered triangle records.
__global__ void k_pack(const float *__restrict__ tris, const float *__restrict__ refb, const uint32_t *__restrict__ ref_orig,
					   const uint32_t *__restrict__ tri_body, const float *__restrict__ body_friction,
					   const uint32_t *__restrict__ body_rayflags,
					   const unsigned long long *__restrict__ keys, int n, const int2 *__restrict__ children,
					   const float *__restrict__ lo_in, const float *__restrict__ hi_in, float4 *__restrict__ nodes,
					   float4 *__restrict__ tri_out)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
	{
		const uint32_t prim = (uint32_t)(keys[i] & 0xFFFFFFFFull);
		const uint32_t orig = ref_orig ? ref_orig[prim] : prim;  // the face id a hit reports
		const float *t = tris + 9ull * orig;
		v3 a = V(t[0], t[1], t[2]), b = V(t[3], t[4], t[5]), c = V(t[6], t[7], t[8]);
		v3 nn = cross(b - a, c - a);
		float l = len(nn);
		v3 nu = l > 0.0f ? nn * (1.0f / l) : V(0.0f, 1.0f, 0.0f);
		uint32_t body = tri_body[orig];
		tri_out[4 * i + 0] = F4(a, __uint_as_float(orig));
		tri_out[4 * i + 1] = F4(b, __uint_as_float(body));
		tri_out[4 * i + 2] = F4(c, body_friction[body]);
		tri_out[4 * i + 3] = F4(nu, __uint_as_float(body_rayflags[body]));
	}
	int n_nodes = n > 1 ? n - 1 : 1;
	if (i < n_nodes)
	{
		int2 c = n > 1 ? children[i] : make_int2(~0, ~0);
		int cc[2] = {c.x, c.y};
		float3 lo[2], hi[2];
#pragma unroll
		for (int k = 0; k < 2; k++)
		{
			if (n == 1 && k == 1)
			{
				// single-triangle map: the second child is an empty box that no query can enter
				lo[k] = make_float3(FLT_MAX, FLT_MAX, FLT_MAX);
				hi[k] = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
				continue;
			}
			if (cc[k] < 0)
				prim_bounds(tris, refb, (uint32_t)(keys[~cc[k]] & 0xFFFFFFFFull), lo[k], hi[k]);
			else
			{
				lo[k] = make_float3(lo_in[3 * cc[k]], lo_in[3 * cc[k] + 1], lo_in[3 * cc[k] + 2]);
				hi[k] = make_float3(hi_in[3 * cc[k]], hi_in[3 * cc[k] + 1], hi_in[3 * cc[k] + 2]);
			}
			lo[k].x -= BVH_PAD; lo[k].y -= BVH_PAD; lo[k].z -= BVH_PAD;
			hi[k].x += BVH_PAD; hi[k].y += BVH_PAD; hi[k].z += BVH_PAD;
		}
		nodes[4 * i + 0] = make_float4(lo[0].x, hi[0].x, lo[0].y, hi[0].y);
		nodes[4 * i + 1] = make_float4(lo[1].x, hi[1].x, lo[1].y, hi[1].y);
		nodes[4 * i + 2] = make_float4(lo[0].z, hi[0].z, lo[1].z, hi[1].z);
		nodes[4 * i + 3] = make_float4(__int_as_float(cc[0]), __int_as_float(cc[1]), 0.0f, 0.0f);
	}
}

uint32_t next_pow2(uint32_t v)
{
	uint32_t p = 1;
	while (p < v) p <<= 1;
	return p;
}

// ---- split references for the ray tree -----------------------------------------------------------------------------
// A shipped map is a few hundred triangles of which some span a whole sector; their boxes make every ray test them.
// For rays the tree is therefore built over REFERENCES: a large triangle is cut along the longest axis of its box,
// again and again (largest box first), and every piece becomes a leaf with the box of that piece — pointing at the
// original triangle, which is what the leaf test intersects and what the hit reports.  The pieces cover the triangle,
// so the set of triangles a ray can reach is unchanged and so is the closest hit.
struct Ref
{
	uint32_t orig;
	int nv;
	double v[10][3];  // the piece: a convex polygon (a triangle cut by axis-aligned planes)
	double lo[3], hi[3], area;
};

static void ref_bounds(Ref &r)
{
	for (int k = 0; k < 3; k++)
	{
		r.lo[k] = 1e300;
		r.hi[k] = -1e300;
	}
	for (int i = 0; i < r.nv; i++)
		for (int k = 0; k < 3; k++)
		{
			r.lo[k] = r.v[i][k] < r.lo[k] ? r.v[i][k] : r.lo[k];
			r.hi[k] = r.v[i][k] > r.hi[k] ? r.v[i][k] : r.hi[k];
		}
	const double ex = r.hi[0] - r.lo[0], ey = r.hi[1] - r.lo[1], ez = r.hi[2] - r.lo[2];
	r.area = 2.0 * (ex * ey + ey * ez + ez * ex);
}

// Sutherland-Hodgman against one axis-aligned half space: keep = below (sign < 0) or above (sign > 0) the plane
static int clip_axis(const double (*in)[3], int n, int axis, double pos, double sign, double (*out)[3])
{
	int m = 0;
	for (int i = 0; i < n; i++)
	{
		const double *a = in[i], *b = in[(i + 1) % n];
		const double da = sign * (a[axis] - pos), db = sign * (b[axis] - pos);
		if (da >= 0.0)
		{
			if (m < 10) memcpy(out[m++], a, sizeof(double) * 3);
		}
		if ((da >= 0.0) != (db >= 0.0))
		{
			const double t = da / (da - db);
			if (m < 10)
			{
				for (int k = 0; k < 3; k++) out[m][k] = a[k] + t * (b[k] - a[k]);
				out[m][axis] = pos;
				m++;
			}
		}
	}
	return m;
}

static void split_references(const std::vector<float> &tris, uint32_t n, uint32_t budget, std::vector<uint32_t> &ref_orig,
							 std::vector<float> &ref_box)
{
	std::vector<Ref> refs(n);
	for (uint32_t i = 0; i < n; i++)
	{
		Ref &r = refs[i];
		r.orig = i;
		r.nv = 3;
		for (int v = 0; v < 3; v++)
			for (int k = 0; k < 3; k++) r.v[v][k] = tris[9ull * i + 3 * v + k];
		ref_bounds(r);
	}
	// largest box first; stop when the budget is spent or the largest box is no bigger than four average ones
	double total = 0.0;
	for (const Ref &r : refs) total += r.area;
	const double floor_area = 4.0 * total / (double)(n ? n : 1) * ((double)n / (double)budget);
	auto cmp = [&refs](uint32_t a, uint32_t b) { return refs[a].area < refs[b].area || (refs[a].area == refs[b].area && a > b); };
	std::vector<uint32_t> heap(n);
	for (uint32_t i = 0; i < n; i++) heap[i] = i;
	std::make_heap(heap.begin(), heap.end(), cmp);
	while (refs.size() < budget && !heap.empty())
	{
		std::pop_heap(heap.begin(), heap.end(), cmp);
		const uint32_t top = heap.back();
		heap.pop_back();
		if (refs[top].area <= floor_area) break;
		const Ref r = refs[top];
		int axis = 0;
		for (int k = 1; k < 3; k++)
			if (r.hi[k] - r.lo[k] > r.hi[axis] - r.lo[axis]) axis = k;
		const double pos = 0.5 * (r.lo[axis] + r.hi[axis]);
		Ref a = r, b = r;
		a.nv = clip_axis(r.v, r.nv, axis, pos, -1.0, a.v);
		b.nv = clip_axis(r.v, r.nv, axis, pos, +1.0, b.v);
		if (a.nv < 3 || b.nv < 3 || a.nv > 9 || b.nv > 9) continue;  // does not split cleanly: stays one leaf
		ref_bounds(a);
		ref_bounds(b);
		refs[top] = a;
		refs.push_back(b);
		heap.push_back(top);
		std::push_heap(heap.begin(), heap.end(), cmp);
		heap.push_back((uint32_t)refs.size() - 1u);
		std::push_heap(heap.begin(), heap.end(), cmp);
	}
	ref_orig.resize(refs.size());
	ref_box.resize(6 * refs.size());
	for (size_t i = 0; i < refs.size(); i++)
	{
		ref_orig[i] = refs[i].orig;
		for (int k = 0; k < 3; k++)
		{
			// outward, past the rounding of the cut positions and of the conversion to float
			const double pad = 1.0e-5 + 1.0e-6 * (fabs(refs[i].lo[k]) + fabs(refs[i].hi[k]));
			ref_box[6 * i + k] = nextafterf((float)(refs[i].lo[k] - pad), -FLT_MAX);
			ref_box[6 * i + 3 + k] = nextafterf((float)(refs[i].hi[k] + pad), FLT_MAX);
		}
	}
}

// ---- topology of the rays' tree: binned SAH, top-down, on the host.  The tree has at most RAY_TREE_MAX_LEAVES leaves and
// is built once per map, so a surface-area-heuristic build costs nothing and saves traversal steps on every ray; the
// device kernels (refit, pack) then run on this topology exactly as they do on Karras' radix tree.
struct SahTopology
{
	std::vector<unsigned long long> keys;  // leaf -> primitive (reference) index
	std::vector<int2> children;            // >= 0 internal node, < 0 ~leaf
	std::vector<int> parent_internal, parent_leaf;
	int depth = 0;  // deepest leaf; the traversal keeps a 64-entry stack per lane
};

static int sah_build(const std::vector<float> &box, std::vector<uint32_t> &idx, uint32_t lo, uint32_t hi, int parent, SahTopology &t,
					 int level = 0)
{
	if (level > t.depth) t.depth = level;
	// returns the child code of the subtree over idx[lo, hi)
	if (hi - lo == 1)
	{
		const int leaf = (int)t.keys.size();
		t.keys.push_back((unsigned long long)idx[lo]);
		t.parent_leaf.push_back(parent);
		return ~leaf;
	}
	const int node = (int)t.children.size();
	t.children.push_back(make_int2(0, 0));
	t.parent_internal.push_back(parent);
	// centroid bounds
	double cl[3] = {1e300, 1e300, 1e300}, ch[3] = {-1e300, -1e300, -1e300};
	for (uint32_t i = lo; i < hi; i++)
		for (int k = 0; k < 3; k++)
		{
			const double c = 0.5 * ((double)box[6ull * idx[i] + k] + (double)box[6ull * idx[i] + 3 + k]);
			cl[k] = c < cl[k] ? c : cl[k];
			ch[k] = c > ch[k] ? c : ch[k];
		}
	constexpr int BINS = 16;
	double best_cost = 1e300;
	int best_axis = -1, best_bin = 0;
	for (int axis = 0; axis < 3; axis++)
	{
		const double ext = ch[axis] - cl[axis];
		if (!(ext > 1e-12)) continue;
		double blo[BINS][3], bhi[BINS][3];
		int cnt[BINS];
		for (int b = 0; b < BINS; b++)
		{
			cnt[b] = 0;
			for (int k = 0; k < 3; k++)
			{
				blo[b][k] = 1e300;
				bhi[b][k] = -1e300;
			}
		}
		for (uint32_t i = lo; i < hi; i++)
		{
			const float *bx = &box[6ull * idx[i]];
			const double c = 0.5 * ((double)bx[axis] + (double)bx[3 + axis]);
			int b = (int)((c - cl[axis]) / ext * BINS);
			b = b < 0 ? 0 : (b >= BINS ? BINS - 1 : b);
			cnt[b]++;
			for (int k = 0; k < 3; k++)
			{
				blo[b][k] = bx[k] < blo[b][k] ? bx[k] : blo[b][k];
				bhi[b][k] = bx[3 + k] > bhi[b][k] ? bx[3 + k] : bhi[b][k];
			}
		}
		// sweep: cost of splitting after bin s = area(left) * n(left) + area(right) * n(right)
		double rl[BINS][3], rh[BINS][3];
		int rc[BINS];
		double al[3] = {1e300, 1e300, 1e300}, ah[3] = {-1e300, -1e300, -1e300};
		int ac = 0;
		for (int b = BINS - 1; b >= 0; b--)
		{
			ac += cnt[b];
			for (int k = 0; k < 3; k++)
			{
				al[k] = blo[b][k] < al[k] ? blo[b][k] : al[k];
				ah[k] = bhi[b][k] > ah[k] ? bhi[b][k] : ah[k];
				rl[b][k] = al[k];
				rh[b][k] = ah[k];
			}
			rc[b] = ac;
		}
		double ll[3] = {1e300, 1e300, 1e300}, lh[3] = {-1e300, -1e300, -1e300};
		int lc = 0;
		for (int sbin = 0; sbin < BINS - 1; sbin++)
		{
			lc += cnt[sbin];
			for (int k = 0; k < 3; k++)
			{
				ll[k] = blo[sbin][k] < ll[k] ? blo[sbin][k] : ll[k];
				lh[k] = bhi[sbin][k] > lh[k] ? bhi[sbin][k] : lh[k];
			}
			if (lc == 0 || rc[sbin + 1] == 0) continue;
			auto area = [](const double *a, const double *b) {
				const double x = b[0] - a[0], y = b[1] - a[1], z = b[2] - a[2];
				return 2.0 * (x * y + y * z + z * x);
			};
			const double cost = area(ll, lh) * lc + area(rl[sbin + 1], rh[sbin + 1]) * rc[sbin + 1];
			if (cost < best_cost)
			{
				best_cost = cost;
				best_axis = axis;
				best_bin = sbin;
			}
		}
	}
	uint32_t mid;
	if (best_axis < 0)
		mid = lo + (hi - lo) / 2;  // all centroids coincide: any balanced split
	else
	{
		const double ext = ch[best_axis] - cl[best_axis];
		auto left = [&](uint32_t p) {
			const float *bx = &box[6ull * p];
			const double c = 0.5 * ((double)bx[best_axis] + (double)bx[3 + best_axis]);
			int b = (int)((c - cl[best_axis]) / ext * BINS);
			b = b < 0 ? 0 : (b >= BINS ? BINS - 1 : b);
			return b <= best_bin;
		};
		mid = (uint32_t)(std::stable_partition(idx.begin() + lo, idx.begin() + hi, left) - idx.begin());
		if (mid == lo || mid == hi) mid = lo + (hi - lo) / 2;
	}
	const int c0 = sah_build(box, idx, lo, mid, node, t, level + 1);
	const int c1 = sah_build(box, idx, mid, hi, node, t, level + 1);
	t.children[node] = make_int2(c0, c1);
	return node;
}

// Device LBVH over `n` primitives (triangles, or references when d_refb/d_ref_orig are given): Morton keys, sort,
// Karras hierarchy, refit, pack.  Allocates *nodes_out / *tri_out.
static int build_tree(gpx_world *w, uint32_t n, const float *d_tris, const float *d_refb, const uint32_t *d_ref_orig,
					  const uint32_t *d_body, const float *d_fr, const uint32_t *d_rf, float3 flo, float3 inv, float4 **nodes_out,
					  float4 **tri_out, const SahTopology *topo = nullptr)
{
	const uint32_t n_nodes = n > 1 ? n - 1 : 1;
	const uint32_t n_pad = next_pow2(n);
	float *d_lo = nullptr, *d_hi = nullptr;
	unsigned long long *d_keys = nullptr;
	int2 *d_children = nullptr;
	int *d_pi = nullptr, *d_pl = nullptr, *d_visit = nullptr;
	cudaStream_t st = w->stream;
	GPX_CUDA(cudaMalloc(&d_keys, sizeof(unsigned long long) * n_pad));
	GPX_CUDA(cudaMalloc(&d_children, sizeof(int2) * (n + 1)));
	GPX_CUDA(cudaMalloc(&d_pi, sizeof(int) * (n + 1)));
	GPX_CUDA(cudaMalloc(&d_pl, sizeof(int) * (n + 1)));
	GPX_CUDA(cudaMalloc(&d_visit, sizeof(int) * (n + 1)));
	GPX_CUDA(cudaMalloc(&d_lo, sizeof(float) * 3ull * (n + 1)));
	GPX_CUDA(cudaMalloc(&d_hi, sizeof(float) * 3ull * (n + 1)));
	GPX_CUDA(cudaMalloc(tri_out, sizeof(float4) * 4ull * n));
	GPX_CUDA(cudaMalloc(nodes_out, sizeof(float4) * 4ull * n_nodes));
	GPX_CUDA(cudaMemsetAsync(d_visit, 0, sizeof(int) * (n + 1), st));
	GPX_CUDA(cudaMemsetAsync(d_pl, 0xFF, sizeof(int) * (n + 1), st));
	const uint32_t tb = 256;
	if (topo)
	{
		// topology from the host (SAH): leaf order, children, parents
		GPX_CUDA(cudaMemcpyAsync(d_keys, topo->keys.data(), sizeof(unsigned long long) * n, cudaMemcpyHostToDevice, st));
		GPX_CUDA(cudaMemcpyAsync(d_children, topo->children.data(), sizeof(int2) * (n - 1), cudaMemcpyHostToDevice, st));
		GPX_CUDA(cudaMemcpyAsync(d_pi, topo->parent_internal.data(), sizeof(int) * (n - 1), cudaMemcpyHostToDevice, st));
		GPX_CUDA(cudaMemcpyAsync(d_pl, topo->parent_leaf.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
	}
	else
	{
		k_morton_keys<<<(n_pad + tb - 1) / tb, tb, 0, st>>>(d_tris, d_refb, n, n_pad, flo, inv, d_keys);
		count_launch();
		bitonic_sort_u64(d_keys, n_pad, st);
	}
	if (n > 1)
	{
		if (!topo)
		{
			k_hierarchy<<<(n + tb - 1) / tb, tb, 0, st>>>(d_keys, (int)n, d_children, d_pi, d_pl);
			count_launch();
		}
		k_refit<<<(n + tb - 1) / tb, tb, 0, st>>>(d_tris, d_refb, d_keys, (int)n, d_children, d_pi, d_pl, d_lo, d_hi, d_visit);
		count_launch();
	}
	k_pack<<<(n + tb - 1) / tb, tb, 0, st>>>(d_tris, d_refb, d_ref_orig, d_body, d_fr, d_rf, d_keys, (int)n, d_children, d_lo, d_hi,
											 *nodes_out, *tri_out);
	count_launch();
	GPX_CUDA(cudaGetLastError());
	GPX_CUDA(cudaStreamSynchronize(st));
	cudaFree(d_keys); cudaFree(d_children); cudaFree(d_pi); cudaFree(d_pl); cudaFree(d_visit); cudaFree(d_lo); cudaFree(d_hi);
	return GPX_OK;
}

int build_static(gpx_world *w)
{
	StaticDevice &sd = w->sd;
	if (sd.ray_tri && sd.ray_tri != sd.tri) cudaFree(sd.ray_tri);
	if (sd.ray_nodes && sd.ray_nodes != sd.nodes) cudaFree(sd.ray_nodes);
	if (sd.tri) cudaFree(sd.tri);
	if (sd.nodes) cudaFree(sd.nodes);
	sd.tri = sd.nodes = sd.ray_tri = sd.ray_nodes = nullptr;
	const uint32_t n = (uint32_t)w->h_tri_body.size();
	sd.n_tris = n;
	sd.n_nodes = n == 0 ? 0 : (n > 1 ? n - 1 : 1);
	sd.n_ray_leaves = sd.n_ray_nodes = 0;
	w->static_dirty = false;
	// leaf indices change with every rebuild: forget the per-body candidate lists
	if (w->d_cand) GPX_CUDA(cudaMemsetAsync(w->d_cand, 0xFF, sizeof(uint4) * 8 * (size_t)w->W * w->cap, w->stream));
	if (n == 0) return GPX_OK;

	// scene bounds on the host (it already walks every vertex while appending)
	float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
	for (size_t i = 0; i < w->h_tris.size(); i++)
	{
		float v = w->h_tris[i];
		int a = (int)(i % 3);
		lo[a] = v < lo[a] ? v : lo[a];
		hi[a] = v > hi[a] ? v : hi[a];
	}
	for (int k = 0; k < 3; k++)
	{
		sd.lo[k] = lo[k];
		sd.hi[k] = hi[k];
	}
	float3 flo = make_float3(lo[0], lo[1], lo[2]);
	float3 inv = make_float3(hi[0] > lo[0] ? 1.0f / (hi[0] - lo[0]) : 0.0f, hi[1] > lo[1] ? 1.0f / (hi[1] - lo[1]) : 0.0f,
							 hi[2] > lo[2] ? 1.0f / (hi[2] - lo[2]) : 0.0f);

	std::vector<float> fr(w->sbodies.size());
	std::vector<uint32_t> rf(w->sbodies.size());
	for (size_t i = 0; i < w->sbodies.size(); i++)
	{
		fr[i] = w->sbodies[i].friction;
		rf[i] = w->sbodies[i].ray_flags;
	}

	float *d_tris = nullptr, *d_fr = nullptr, *d_refb = nullptr;
	uint32_t *d_body = nullptr, *d_rf = nullptr, *d_ref_orig = nullptr;
	cudaStream_t st = w->stream;
	GPX_CUDA(cudaMalloc(&d_tris, sizeof(float) * 9ull * n));
	GPX_CUDA(cudaMalloc(&d_body, sizeof(uint32_t) * n));
	GPX_CUDA(cudaMalloc(&d_fr, sizeof(float) * fr.size()));
	GPX_CUDA(cudaMalloc(&d_rf, sizeof(uint32_t) * rf.size()));
	GPX_CUDA(cudaMemcpyAsync(d_tris, w->h_tris.data(), sizeof(float) * 9ull * n, cudaMemcpyHostToDevice, st));
	GPX_CUDA(cudaMemcpyAsync(d_body, w->h_tri_body.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
	GPX_CUDA(cudaMemcpyAsync(d_fr, fr.data(), sizeof(float) * fr.size(), cudaMemcpyHostToDevice, st));
	GPX_CUDA(cudaMemcpyAsync(d_rf, rf.data(), sizeof(uint32_t) * rf.size(), cudaMemcpyHostToDevice, st));

	// the tick's tree: one leaf per triangle (body-vs-map candidates must name each triangle once)
	int rc = build_tree(w, n, d_tris, nullptr, nullptr, d_body, d_fr, d_rf, flo, inv, &sd.nodes, &sd.tri);
	sd.ray_nodes = sd.nodes;
	sd.ray_tri = sd.tri;
	sd.n_ray_leaves = n;
	sd.n_ray_nodes = sd.n_nodes;
	sd.ray_depth = 64;
	// the rays' tree: split references, as many as still let the whole tree sit in one SM's shared memory, under a SAH
	// topology built on the host (maps too large for shared memory keep one leaf per triangle but still get the SAH)
	uint32_t budget = RAY_TREE_MAX_LEAVES;
	if (const char *e = getenv("GPX_RAY_LEAVES")) budget = (uint32_t)atoi(e);
	if (rc == GPX_OK && n >= 2 && n <= RAY_TREE_MAX_HOST_BUILD && getenv("GPX_RAY_LBVH") == nullptr)
	{
		std::vector<uint32_t> ref_orig;
		std::vector<float> ref_box;
		split_references(w->h_tris, n, budget > n ? budget : n, ref_orig, ref_box);
		const uint32_t nr = (uint32_t)ref_orig.size();
		GPX_CUDA(cudaMalloc(&d_refb, sizeof(float) * 6ull * nr));
		GPX_CUDA(cudaMalloc(&d_ref_orig, sizeof(uint32_t) * nr));
		GPX_CUDA(cudaMemcpyAsync(d_refb, ref_box.data(), sizeof(float) * 6ull * nr, cudaMemcpyHostToDevice, st));
		GPX_CUDA(cudaMemcpyAsync(d_ref_orig, ref_orig.data(), sizeof(uint32_t) * nr, cudaMemcpyHostToDevice, st));
		SahTopology topo;
		std::vector<uint32_t> idx(nr);
		for (uint32_t i = 0; i < nr; i++) idx[i] = i;
		sah_build(ref_box, idx, 0, nr, -1, topo);
		// a SAH tree deeper than the traversal stack (pathological input) falls back to the radix tree, which never is
		rc = build_tree(w, nr, d_tris, d_refb, d_ref_orig, d_body, d_fr, d_rf, flo, inv, &sd.ray_nodes, &sd.ray_tri,
						topo.depth <= 56 ? &topo : nullptr);
		sd.n_ray_leaves = nr;
		sd.n_ray_nodes = nr - 1;
		sd.ray_depth = topo.depth <= 56 ? (uint32_t)topo.depth : 64u;
	}
	cudaFree(d_tris); cudaFree(d_body); cudaFree(d_fr); cudaFree(d_rf); cudaFree(d_refb); cudaFree(d_ref_orig);
	return rc;
}

}  // namespace gpx
