// gpx_char.cu — the player character: a capsule moved by collide-and-slide against the map and the solid bodies.
//
// Replaces JPH_CharacterVirtual as the engine uses it (engine/src/physics/PlayerPhysics.c): created per map
// (:173-194, capsule half height 0.2, radius 0.25, max slope 50 degrees), given a velocity by MovePlayer (:203-295),
// advanced by JPH_CharacterVirtual_ExtendedUpdate BEFORE the physics update (:439-453, MapPhysics.c:66-77), read back
// with GetPosition / GetLinearVelocity / GetGroundState, and its contacts forwarded to actor callbacks (:89-152).
// One warp per world: the lanes share the candidate triangles (LBVH box query) and the world's bodies, each computes
// segment-vs-shape closest points, and a shuffle reduction picks the deepest penetration (ties: lowest id) — up to
// eight push-outs per update, then the ground probe / stick-to-floor step and the contact list that the tick's event
// pass merges with the body contacts.
#include <cstring>
#include "gpx_solver.cuh"

namespace gpx {

constexpr int CH_MAX_ITERS = 8;
constexpr float CH_CONTACT_MARGIN = 0.02f;
constexpr float CH_GROUND_PROBE = 0.05f;

__device__ __forceinline__ float clamp01(float t) { return t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t); }

// closest points of segments p1-q1 and p2-q2 (Ericson, Real-Time Collision Detection 5.1.9)
__device__ __forceinline__ void seg_seg(v3 p1, v3 q1, v3 p2, v3 q2, v3 &c1, v3 &c2)
{
	v3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
	float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
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
		float c = dot(d1, r);
		if (e <= EPS)
		{
			t = 0.0f;
			s = clamp01(-c / a);
		}
		else
		{
			float b = dot(d1, d2);
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
	c1 = p1 + (d1 * s);
	c2 = p2 + (d2 * t);
}

// closest points between segment p0-p1 and a triangle; returns squared distance (0 when the segment pierces it)
static __device__ __noinline__ float seg_tri(v3 p0, v3 p1, v3 a, v3 b, v3 c, v3 &cs, v3 &ct)
{
	{
		v3 d = p1 - p0, e1 = b - a, e2 = c - a;
		v3 pv = cross(d, e2);
		float det = dot(e1, pv);
		if (fabsf(det) >= 1.0e-12f)
		{
			float inv = 1.0f / det;
			v3 tv = p0 - a;
			float u = dot(tv, pv) * inv;
			if (u >= 0.0f && u <= 1.0f)
			{
				v3 q = cross(tv, e1);
				float v = dot(d, q) * inv;
				if (v >= 0.0f && (u + v) <= 1.0f)
				{
					float t = dot(e2, q) * inv;
					if (t >= 0.0f && t <= 1.0f)
					{
						cs = ct = p0 + (d * t);
						return 0.0f;
					}
				}
			}
		}
	}
	float best = 3.0e38f;
#pragma unroll
	for (int i = 0; i < 3; i++)
	{
		const v3 ea = i == 0 ? a : (i == 1 ? b : c), eb = i == 0 ? b : (i == 1 ? c : a);
		v3 x, y;
		seg_seg(p0, p1, ea, eb, x, y);
		float d2 = len2(x - y);
		if (d2 < best) { best = d2; cs = x; ct = y; }
	}
#pragma unroll
	for (int i = 0; i < 2; i++)
	{
		const v3 e = i == 0 ? p0 : p1;
		v3 y = closest_on_tri(e, a, b, c);
		float d2 = len2(e - y);
		if (d2 < best) { best = d2; cs = e; ct = y; }
	}
	return best;
}

// closest points between segment p0-p1 and an oriented box; returns squared distance (0 when the segment enters it)
static __device__ __noinline__ float seg_box(v3 p0, v3 p1, v3 bx, q4 bq, v3 he, v3 &cs, v3 &cb)
{
	m33 R = qmat(bq);
	v3 l0 = mtmul(R, p0 - bx), l1 = mtmul(R, p1 - bx);
	{
		v3 d = l1 - l0;
		float tn = 0.0f, tf = 1.0f;
		bool hit = true;
#pragma unroll
		for (int k = 0; k < 3; k++)
		{
			if (!hit) continue;
			float ok = get(l0, k), dk = get(d, k), hk = get(he, k);
			if (dk == 0.0f)
			{
				if (ok < -hk || ok > hk) hit = false;
				continue;
			}
			float inv = 1.0f / dk;
			float t1 = (-hk - ok) * inv, t2 = (hk - ok) * inv;
			if (t1 > t2) { float tt = t1; t1 = t2; t2 = tt; }
			if (t1 > tn) tn = t1;
			if (t2 < tf) tf = t2;
			if (tn > tf) hit = false;
		}
		if (hit)
		{
			v3 lp = l0 + (d * tn);
			cs = cb = bx + mmul(R, lp);
			return 0.0f;
		}
	}
	float best = 3.0e38f;
	v3 bs = l0, bb = l0;
#pragma unroll
	for (int i = 0; i < 2; i++)
	{
		v3 l = i == 0 ? l0 : l1;
		v3 q = V(fminf(fmaxf(l.x, -he.x), he.x), fminf(fmaxf(l.y, -he.y), he.y), fminf(fmaxf(l.z, -he.z), he.z));
		float d2 = len2(l - q);
		if (d2 < best) { best = d2; bs = l; bb = q; }
	}
	for (int k = 0; k < 3; k++)
	{
		const int u = (k + 1) % 3, v = (k + 2) % 3;
		for (int sgn = 0; sgn < 4; sgn++)
		{
			const float su = (sgn & 1) ? 1.0f : -1.0f, sv = (sgn & 2) ? 1.0f : -1.0f;
			float e0[3], e1[3];
			e0[k] = -get(he, k); e1[k] = get(he, k);
			e0[u] = e1[u] = su * get(he, u);
			e0[v] = e1[v] = sv * get(he, v);
			v3 x, y;
			seg_seg(l0, l1, V(e0[0], e0[1], e0[2]), V(e1[0], e1[1], e1[2]), x, y);
			float d2 = len2(x - y);
			if (d2 < best) { best = d2; bs = x; bb = y; }
		}
	}
	cs = bx + mmul(R, bs);
	cb = bx + mmul(R, bb);
	return best;
}

__device__ __forceinline__ float seg_point(v3 p0, v3 p1, v3 c, v3 &cs)
{
	v3 d = p1 - p0;
	float a = dot(d, d);
	float t = a > 1.0e-12f ? clamp01(dot(c - p0, d) / a) : 0.0f;
	cs = p0 + (d * t);
	return len2(cs - c);
}

struct CharArgs
{
	CharDev *ch;
	unsigned long long *keys;  // 64 per world
	uint32_t *nkeys;
	BodyStore bs;
	StaticView sv;
	uint32_t worlds, cap;
	uint32_t *err;
	float dt;
	gpx_character_update_settings cfg;  // JPH_ExtendedUpdateSettings; all zero = plain update
};

struct Deepest
{
	float pen;
	uint32_t id;
	v3 n;
	uint32_t body;
	v3 cp;  // contact point on the thing hit
};

// Deepest penetration of the capsule centred at x (warp-cooperative; every lane returns the same result).
__device__ __forceinline__ Deepest ch_deepest(const CharArgs &a, uint32_t world, v3 x, float hh, float r, int *cand_orig,
											 int *cand_leaf, int *s_nc, uint32_t &err)
{
	const int lane = threadIdx.x & 31;
	const v3 p0 = V(x.x, x.y - hh, x.z), p1 = V(x.x, x.y + hh, x.z);
	if (lane == 0)
	{
		bool overflow = false;
		*s_nc = query_static(a.sv.nodes, a.sv.tris, a.sv.n_nodes, V(p0.x - r, p0.y - r, p0.z - r), V(p1.x + r, p1.y + r, p1.z + r),
							 0.0f, cand_orig, cand_leaf, overflow);
		if (overflow) err |= GPX_ERR_BODY_PAIR_CACHE_FULL;
	}
	__syncwarp();
	const int nc = *s_nc;
	Deepest best;
	best.pen = 0.0f;
	best.id = 0xFFFFFFFFu;
	best.n = V(0.0f, 1.0f, 0.0f);
	best.body = GPX_INVALID_BODY;
	best.cp = V(0.0f, 0.0f, 0.0f);
	for (int c = lane; c < nc; c += 32)
	{
		const int leaf = cand_leaf[c];
		const float4 TA = __ldg(&a.sv.tris[4 * leaf + 0]), TB = __ldg(&a.sv.tris[4 * leaf + 1]),
					 TC = __ldg(&a.sv.tris[4 * leaf + 2]), TN = __ldg(&a.sv.tris[4 * leaf + 3]);
		v3 cs, ct;
		const float d2 = seg_tri(p0, p1, V(TA), V(TB), V(TC), cs, ct);
		const float dist = sqrtf(d2);
		const float pen = r - dist;
		const uint32_t id = (uint32_t)cand_orig[c];
		if (pen > best.pen || (pen == best.pen && pen > 0.0f && id < best.id))
		{
			v3 n;
			if (dist > 1.0e-6f) n = (cs - ct) * (1.0f / dist);
			else n = dot(x - V(TA), V(TN)) >= 0.0f ? V(TN) : -V(TN);
			best.pen = pen;
			best.id = id;
			best.n = n;
			best.body = STATIC_BODY_BASE + __float_as_uint(TB.w);
			best.cp = ct;
		}
	}
	const uint32_t g0 = world * a.cap;
	for (uint32_t i = lane; i < a.cap; i += 32)
	{
		const uint32_t f = a.bs.flags[g0 + i];
		const uint32_t shape = shape_of(f), layer = layer_of(f);
		if (!(f & BF_ALIVE) || shape == GPX_SHAPE_EMPTY || (f & BF_SENSOR) || !(layer == 0 || layer == 1)) continue;
		const v3 bx = V(a.bs.pos[g0 + i]);
		const float4 p1h = a.bs.prop1[g0 + i];
		v3 cs, cb;
		float d2, rr = r;
		if (shape == GPX_SHAPE_BOX) d2 = seg_box(p0, p1, bx, Q(a.bs.quat[g0 + i]), V(p1h), cs, cb);
		else
		{
			d2 = seg_point(p0, p1, bx, cs);
			cb = bx;
			rr = r + p1h.x;
		}
		const float dist = sqrtf(d2);
		const float pen = rr - dist;
		const uint32_t id = 0x80000000u + i;
		if (pen > best.pen || (pen == best.pen && pen > 0.0f && id < best.id))
		{
			v3 n;
			if (dist > 1.0e-6f) n = (cs - cb) * (1.0f / dist);
			else
			{
				v3 d = x - bx;
				float l = len(d);
				n = l > 1.0e-6f ? d * (1.0f / l) : V(0.0f, 1.0f, 0.0f);
			}
			best.pen = pen;
			best.id = id;
			best.n = n;
			best.body = i;
			best.cp = cb;
		}
	}
	// warp reduction: largest penetration, ties to the lowest id
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		const float open = __shfl_xor_sync(0xFFFFFFFFu, best.pen, o);
		const uint32_t oid = __shfl_xor_sync(0xFFFFFFFFu, best.id, o);
		const float onx = __shfl_xor_sync(0xFFFFFFFFu, best.n.x, o), ony = __shfl_xor_sync(0xFFFFFFFFu, best.n.y, o),
					onz = __shfl_xor_sync(0xFFFFFFFFu, best.n.z, o);
		const uint32_t ob = __shfl_xor_sync(0xFFFFFFFFu, best.body, o);
		const float ocx = __shfl_xor_sync(0xFFFFFFFFu, best.cp.x, o), ocy = __shfl_xor_sync(0xFFFFFFFFu, best.cp.y, o),
					ocz = __shfl_xor_sync(0xFFFFFFFFu, best.cp.z, o);
		if (open > best.pen || (open == best.pen && oid < best.id))
		{
			best.pen = open;
			best.id = oid;
			best.n = V(onx, ony, onz);
			best.body = ob;
			best.cp = V(ocx, ocy, ocz);
		}
	}
	__syncwarp();
	return best;
}

// The character pushes the dynamic bodies it runs into (CharacterVirtual's contact impulse): the body should move away
// at 0.9 of the closing speed plus 0.4 of the penetration per update, through its effective mass at the contact point,
// capped by the character's strength (100 N) times dt; no push along gravity; the body wakes.  Every lane computes the
// same values, lane 0 stores.  `n` points from the body to the character, `v` is the character's velocity.
constexpr float CH_PUSH_DAMPING = 0.9f, CH_PUSH_PENETRATION = 0.4f, CH_MAX_STRENGTH = 100.0f;

__device__ __forceinline__ void ch_push_body(const CharArgs &a, uint32_t g, v3 n, float pen, v3 cp, v3 v, float dt, int lane)
{
	uint32_t f = a.bs.flags[g];
	if (((f >> BF_MOTION_SHIFT) & 3u) != GPX_MOTION_DYNAMIC || (f & BF_SENSOR)) return;
	SBody B;
	B.x = V(a.bs.pos[g]);
	B.q = Q(a.bs.quat[g]);
	B.v = V(a.bs.lin[g]);
	B.w = V(a.bs.ang[g]);
	const float4 p0 = a.bs.prop0[g];
	B.inv_mass = p0.x;
	B.inv_i = V(p0.y, p0.z, p0.w);
	const v3 rB = cp - B.x;
	const v3 vB = B.v + cross(B.w, rB);
	const float dv = (-(dot(v - vB, n)) * CH_PUSH_DAMPING) + ((pen * CH_PUSH_PENETRATION) / dt);
	if (!(dv > 0.0f)) return;
	const bool woke = (f & BF_ASLEEP) != 0;
	f &= ~BF_ASLEEP;  // AddImpulse activates
	B.flags = f;
	body_world_inertia(B);
	const v3 jac = cross(rB, n);
	const float inv_eff = dot(sym_mul(B.M, jac), jac) + B.im;
	if (lane == 0 && woke)
	{
		a.bs.flags[g] = f;
		a.bs.sleep_t[g] = -1.0f;
	}
	if (!(inv_eff > 0.0f)) return;
	const float impulse = fminf(dv / inv_eff, CH_MAX_STRENGTH * dt);
	v3 P = n * (-impulse);
	if (P.y < 0.0f) P.y = 0.0f;
	const v3 nv = B.v + mask_lin(dofs_of(f), P * B.im);
	const v3 nw = B.w + sym_mul(B.M, cross(rB, P));
	if (lane == 0)
	{
		a.bs.lin[g] = F4(nv, 0.0f);
		a.bs.ang[g] = F4(nw, 0.0f);
	}
}

// One collide-and-slide pass (warp-cooperative; every lane carries the same state): up to CH_MAX_ITERS times find the
// deepest penetration, push the capsule out along that normal and remove the velocity component into it; ground state from
// the contact normals.  `push`: dynamic bodies in the way get the character's contact impulse.  `blocked`: some contact
// too steep to walk on faces the motion direction `dir` within the angle whose cosine is cos_fwd.
constexpr int CH_MAX_PIECES = 16;  // pieces of one tick's move: 2 m per tick, 120 m/s
struct Slide
{
	v3 x, v, ground_n;
	uint32_t ground, ground_body;
	bool blocked;
};

__device__ __forceinline__ Slide ch_slide(const CharArgs &a, uint32_t world, const CharDev &ch, v3 x, v3 v, float dt, bool push, v3 dir, float cos_fwd,
										  int *cand_orig, int *cand_leaf, int *s_nc, uint32_t &err, int lane)
{
	Slide s;
	s.ground = 3u;
	s.ground_body = GPX_INVALID_BODY;
	s.ground_n = V(0.0f, 1.0f, 0.0f);
	s.blocked = false;
	for (int it = 0; it < CH_MAX_ITERS; it++)
	{
		const Deepest d = ch_deepest(a, world, x, ch.hh, ch.r, cand_orig, cand_leaf, s_nc, err);
		if (!(d.pen > 0.0f)) break;
		if (push && d.body < STATIC_BODY_BASE)
		{
			ch_push_body(a, world * a.cap + d.body, d.n, d.pen, d.cp, v, dt, lane);
			__syncwarp();  // the next round reads the body's new velocity
		}
		x = x + (d.n * d.pen);
		const float vn = dot(v, d.n);
		if (vn < 0.0f) v = v - (d.n * vn);
		if (d.n.y >= ch.cos_slope)
		{
			s.ground = 0u;
			s.ground_body = d.body;
			s.ground_n = d.n;
		}
		else
		{
			if (d.n.y > 0.0f && s.ground != 0u)
			{
				s.ground = 1u;
				s.ground_body = d.body;
				s.ground_n = d.n;
			}
			// too steep to walk on: does it face the motion?
			const float hl = sqrtf((d.n.x * d.n.x) + (d.n.z * d.n.z));
			if (hl > 1.0e-6f && (-((d.n.x * dir.x) + (d.n.z * dir.z))) >= (cos_fwd * hl)) s.blocked = true;
		}
	}
	s.x = x;
	s.v = v;
	return s;
}

// Floor below x within `reach`, by probes 5 cm apart: the first probe depth at which the capsule touches something that
// faces up (0: nothing); `d` = what was found there.
__device__ __forceinline__ float ch_probe_down(const CharArgs &a, uint32_t world, const CharDev &ch, v3 x, float reach, Deepest &d,
											   int *cand_orig, int *cand_leaf, int *s_nc, uint32_t &err)
{
	for (float s = CH_GROUND_PROBE; s <= reach + 1.0e-6f; s += CH_GROUND_PROBE)
	{
		d = ch_deepest(a, world, V(x.x, x.y - s, x.z), ch.hh, ch.r, cand_orig, cand_leaf, s_nc, err);
		if (d.pen > 0.0f && d.n.y > 0.0f) return s;
	}
	return 0.0f;
}

// JPH_CharacterVirtual_ExtendedUpdate as the engine calls it (PlayerPhysics.c:439-453), restated on the discrete
// collide-and-slide: the move, then stick-to-floor, then walk-stairs.
__global__ void __launch_bounds__(128) k_character(CharArgs a)
{
	__shared__ int s_orig[4][MAX_TRI_CANDIDATES], s_leaf[4][MAX_TRI_CANDIDATES], s_nc[4];
	__shared__ unsigned char s_touch[4][MAX_TRI_CANDIDATES];
	const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t world = blockIdx.x * 4 + wib;
	if (world >= a.worlds) return;
	CharDev ch = a.ch[world];
	if (!ch.alive) return;
	uint32_t err = 0;
	const float hh = ch.hh, r = ch.r;
	const gpx_character_update_settings &cfg = a.cfg;
	const v3 x_old = V(ch.px, ch.py, ch.pz), v_in = V(ch.vx, ch.vy, ch.vz);
	const bool was_on_ground = ch.ground == 0u || ch.ground == 1u;  // Jolt's IsSupported(): on ground or on steep ground
	// what the host asked for, horizontally
	const v3 want = V(v_in.x * a.dt, 0.0f, v_in.z * a.dt);
	const float want_len = sqrtf((want.x * want.x) + (want.z * want.z));
	const v3 dir = want_len > 0.0f ? V(want.x / want_len, 0.0f, want.z / want_len) : V(0.0f, 0.0f, 0.0f);
	// swept motion, restated on the discrete test: the move in pieces no longer than half the capsule's radius, each
	// collided and slid before the next, so that no floor or wall is stepped over (one piece up to 7.5 m/s)
	const float travel = len(v_in * a.dt);
	int pieces = (int)ceilf(travel / (0.5f * r));
	pieces = pieces < 1 ? 1 : (pieces > CH_MAX_PIECES ? CH_MAX_PIECES : pieces);
	const float pdt = a.dt / (float)pieces;
	Slide m = ch_slide(a, world, ch, x_old + (v_in * pdt), v_in, pdt, true, dir, cfg.walk_stairs_cos_angle_forward_contact, s_orig[wib],
					   s_leaf[wib], &s_nc[wib], err, lane);
	for (int k = 1; k < pieces; k++)
	{
		const bool blocked = m.blocked;
		m = ch_slide(a, world, ch, m.x + (m.v * pdt), m.v, pdt, true, dir, cfg.walk_stairs_cos_angle_forward_contact, s_orig[wib],
					 s_leaf[wib], &s_nc[wib], err, lane);
		m.blocked |= blocked;
	}
	v3 x = m.x, v = m.v;
	uint32_t ground = m.ground, ground_body = m.ground_body;
	v3 ground_n = m.ground_n;
	if (ground == 3u)
	{
		// ground within 5 cm counts as ground; beyond that only stick-to-floor reaches, and only from a standing start
		const float reach = (was_on_ground && v.y <= 0.0f && cfg.stick_to_floor_step_down > CH_GROUND_PROBE) ? cfg.stick_to_floor_step_down
																										   : CH_GROUND_PROBE;
		Deepest d;
		const float s = ch_probe_down(a, world, ch, x, reach, d, s_orig[wib], s_leaf[wib], &s_nc[wib], err);
		if (s > 0.0f)
		{
			ground = d.n.y >= ch.cos_slope ? 0u : 1u;
			ground_body = d.body;
			ground_n = d.n;
			// stick to the floor: close the gap over walkable ground when not moving up
			if (ground == 0u && v.y <= 0.0f) x.y = x.y - fmaxf(0.0f, s - d.pen);
		}
	}
	if (cfg.walk_stairs_step_up > 0.0f && want_len > 0.0f && (ground == 0u || ground == 1u || was_on_ground) && m.blocked)
	{
		const v3 got = x - x_old;
		const float got_len = fmaxf(0.0f, (got.x * dir.x) + (got.z * dir.z));
		if ((got_len + 1.0e-4f) < want_len)
		{
			const float fwd = fmaxf(cfg.walk_stairs_min_step_forward, want_len - got_len);
			const v3 up = V(x.x, x.y + cfg.walk_stairs_step_up, x.z);
			const Deepest head = ch_deepest(a, world, up, hh, r, s_orig[wib], s_leaf[wib], &s_nc[wib], err);
			if (!(head.pen > 0.0f))  // head room
			{
				// the lifted move forward, swept like the move itself
				int fp = (int)ceilf(fwd / (0.5f * r));
				fp = fp < 1 ? 1 : (fp > CH_MAX_PIECES ? CH_MAX_PIECES : fp);
				const float step = fwd / (float)fp;
				Slide f = ch_slide(a, world, ch, V(up.x + (dir.x * step), up.y, up.z + (dir.z * step)), v, a.dt, false, dir, 2.0f, s_orig[wib],
								   s_leaf[wib], &s_nc[wib], err, lane);
				for (int k = 1; k < fp; k++)
					f = ch_slide(a, world, ch, V(f.x.x + (dir.x * step), f.x.y, f.x.z + (dir.z * step)), f.v, a.dt, false, dir, 2.0f,
								 s_orig[wib], s_leaf[wib], &s_nc[wib], err, lane);
				const v3 adv = f.x - up;
				// headway, and on the level: a push-out that lifted the capsule means the step is higher than step_up
				if (((adv.x * dir.x) + (adv.z * dir.z)) > 1.0e-4f && fabsf(adv.y) <= 1.0e-3f)
				{
					Deepest d;
					const float s = ch_probe_down(a, world, ch, f.x, cfg.walk_stairs_step_up + CH_GROUND_PROBE, d, s_orig[wib], s_leaf[wib],
												  &s_nc[wib], err);
					bool ok = s > 0.0f && d.n.y >= ch.cos_slope;
					if (s > 0.0f && !ok && cfg.walk_stairs_step_forward_test > 0.0f)
					{
						// landed on the edge of the step: is there walkable floor a little further on?
						const float t = cfg.walk_stairs_step_forward_test;
						const v3 ahead = V(up.x + (dir.x * t), up.y, up.z + (dir.z * t));
						// (the lifted capsule must fit there: a step higher than step_up is in the way)
						const Deepest fit = ch_deepest(a, world, ahead, hh, r, s_orig[wib], s_leaf[wib], &s_nc[wib], err);
						if (!(fit.pen > 0.0f))
						{
							Deepest d2;
							const float s2 = ch_probe_down(a, world, ch, ahead, cfg.walk_stairs_step_up + CH_GROUND_PROBE, d2, s_orig[wib],
														   s_leaf[wib], &s_nc[wib], err);
							ok = s2 > 0.0f && d2.n.y >= ch.cos_slope;
						}
					}
					if (ok)
					{
						x = V(f.x.x, f.x.y - fmaxf(0.0f, s - d.pen), f.x.z);
						ground = d.n.y >= ch.cos_slope ? 0u : 1u;
						ground_body = d.body;
						ground_n = d.n;
					}
				}
			}
		}
	}
	// ---- contacts for the callbacks: bodies (sensors included) and static meshes within the contact margin
	const v3 p0 = V(x.x, x.y - hh, x.z), p1 = V(x.x, x.y + hh, x.z);
	const float reach = r + CH_CONTACT_MARGIN;
	unsigned long long *keys = a.keys + 64ull * world;
	uint32_t nkeys = 0;
	const uint32_t g0 = world * a.cap;
	for (uint32_t i0 = 0; i0 < a.cap; i0 += 32)
	{
		const uint32_t i = i0 + lane;
		bool touch = false;
		if (i < a.cap)
		{
			const uint32_t f = a.bs.flags[g0 + i];
			const uint32_t shape = shape_of(f), layer = layer_of(f);
			if ((f & BF_ALIVE) && shape != GPX_SHAPE_EMPTY && (layer == 0 || layer == 1 || layer == 3))
			{
				const v3 bx = V(a.bs.pos[g0 + i]);
				const float4 p1h = a.bs.prop1[g0 + i];
				v3 cs, cb;
				float d2, rr = reach;
				if (shape == GPX_SHAPE_BOX) d2 = seg_box(p0, p1, bx, Q(a.bs.quat[g0 + i]), V(p1h), cs, cb);
				else
				{
					d2 = seg_point(p0, p1, bx, cs);
					rr = reach + p1h.x;
				}
				touch = d2 <= rr * rr;
			}
		}
		const uint32_t m = __ballot_sync(0xFFFFFFFFu, touch);
		if (touch)
		{
			const uint32_t k = nkeys + __popc(m & ((1u << lane) - 1u));
			if (k < 64u) keys[k] = ((unsigned long long)i << 32) | (unsigned long long)CHARACTER_BODY_ID;
		}
		nkeys = min(nkeys + __popc(m), 64u);
	}
	if (lane == 0)
	{
		bool overflow = false;
		s_nc[wib] = query_static(a.sv.nodes, a.sv.tris, a.sv.n_nodes, V(p0.x - reach, p0.y - reach, p0.z - reach),
								 V(p1.x + reach, p1.y + reach, p1.z + reach), 0.0f, s_orig[wib], s_leaf[wib], overflow);
		if (overflow) err |= GPX_ERR_BODY_PAIR_CACHE_FULL;
	}
	__syncwarp();
	const int nc = s_nc[wib];
	for (int c = lane; c < nc; c += 32)
	{
		const int leaf = s_leaf[wib][c];
		const float4 TA = __ldg(&a.sv.tris[4 * leaf + 0]), TB = __ldg(&a.sv.tris[4 * leaf + 1]), TC = __ldg(&a.sv.tris[4 * leaf + 2]);
		v3 cs, ct;
		const float d2 = seg_tri(p0, p1, V(TA), V(TB), V(TC), cs, ct);
		s_touch[wib][c] = d2 <= reach * reach ? 1 : 0;
	}
	__syncwarp();
	if (lane == 0)
	{
		// candidates are in upload order and the triangles of one mesh are contiguous: one contact per static mesh
		uint32_t last = 0xFFFFFFFFu;
		for (int c = 0; c < nc && nkeys < 64u; c++)
		{
			if (!s_touch[wib][c]) continue;
			const uint32_t sb = __float_as_uint(__ldg(&a.sv.tris[4 * s_leaf[wib][c] + 1]).w);
			if (sb == last) continue;
			keys[nkeys++] = ((unsigned long long)CHARACTER_BODY_ID << 32) | (unsigned long long)(STATIC_BODY_BASE + sb);
			last = sb;
		}
		a.nkeys[world] = nkeys;
		ch.px = x.x; ch.py = x.y; ch.pz = x.z;
		ch.vx = v.x; ch.vy = v.y; ch.vz = v.z;
		ch.gnx = ground_n.x; ch.gny = ground_n.y; ch.gnz = ground_n.z;
		ch.ground = ground;
		ch.ground_body = ground_body;
		a.ch[world] = ch;
		if (err)
		{
			atomicOr(&a.err[1 + world], err);
			atomicOr(&a.err[0], err);
		}
	}
}

// ---- batched capsule overlap queries: the collide-shape query the character controller is made of, for callers that
// bring their own capsules (one warp per query; same routine, same tie-breaks as the character's push-out)
__global__ void __launch_bounds__(128) k_overlap_capsules(CharArgs a, const float4 *__restrict__ q, unsigned long long n, float4 *__restrict__ out)
{
	__shared__ int s_orig[4][MAX_TRI_CANDIDATES], s_leaf[4][MAX_TRI_CANDIDATES], s_nc[4];
	const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const unsigned long long i = (unsigned long long)blockIdx.x * 4ull + wib;
	if (i >= n) return;
	const float4 q0 = __ldg(&q[2 * i]), q1 = __ldg(&q[2 * i + 1]);  // centre xyz, half height | radius, world, -, -
	const uint32_t world = __float_as_uint(q1.y);
	uint32_t err = 0;
	Deepest d;
	d.pen = 0.0f;
	d.n = V(0.0f, 1.0f, 0.0f);
	d.body = GPX_INVALID_BODY;
	if (world < a.worlds) d = ch_deepest(a, world, V(q0), q0.w, q1.x, s_orig[wib], s_leaf[wib], &s_nc[wib], err);
	if (lane == 0)
	{
		const bool hit = d.pen > 0.0f;
		out[2 * i] = make_float4(hit ? d.pen : 0.0f, d.n.x, d.n.y, d.n.z);
		out[2 * i + 1] = make_float4(__uint_as_float(hit ? d.body : GPX_INVALID_BODY), __uint_as_float(world), 0.0f, 0.0f);
		if (err) atomicOr(&a.err[0], err);
	}
}

int launch_overlap_capsules(gpx_world *w, const void *d_queries, uint64_t n, void *d_out)
{
	if (n == 0) return GPX_OK;
	CharArgs a;
	a.ch = nullptr;
	a.keys = nullptr;
	a.nkeys = nullptr;
	a.bs = w->bs;
	a.sv.nodes = w->sd.nodes;
	a.sv.tris = w->sd.tri;
	a.sv.n_nodes = w->sd.n_nodes;
	a.worlds = w->W;
	a.cap = w->cap;
	a.err = w->d_err;
	a.dt = 0.0f;
	memset(&a.cfg, 0, sizeof(a.cfg));
	k_overlap_capsules<<<(unsigned)((n + 3) / 4), 128, 0, w->stream>>>(a, (const float4 *)d_queries, n, (float4 *)d_out);
	count_launch();
	GPX_CUDA(cudaGetLastError());
	return GPX_OK;
}

int launch_character(gpx_world *w, float dt, const gpx_character_update_settings *cfg)
{
	CharArgs a;
	memset(&a.cfg, 0, sizeof(a.cfg));
	if (cfg) a.cfg = *cfg;
	a.ch = w->d_ch;
	a.keys = w->d_ch_keys;
	a.nkeys = w->d_ch_nkeys;
	a.bs = w->bs;
	a.sv.nodes = w->sd.nodes;
	a.sv.tris = w->sd.tri;
	a.sv.n_nodes = w->sd.n_nodes;
	a.worlds = w->W;
	a.cap = w->cap;
	a.err = w->d_err;
	a.dt = dt;
	k_character<<<(w->W + 3) / 4, 128, 0, w->stream>>>(a);
	count_launch();
	GPX_CUDA(cudaGetLastError());
	return GPX_OK;
}

}  // namespace gpx
