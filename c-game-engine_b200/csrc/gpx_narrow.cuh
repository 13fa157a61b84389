// gpx_narrow.cuh — per-pair contact generation (one thread per pair).
//
// Stands in for Jolt's narrow phase as the reference reaches it through JPH_PhysicsSystem_Update
// (engine/src/physics/MapPhysics.c:105-108): find the axis of least penetration (SAT for boxes / closest features
// for spheres), take the supporting face of each shape along it, clip one face against the other, keep the points
// within the speculative contact distance, and reduce to four points.
#pragma once
#include "gpx_math.cuh"

namespace gpx {

constexpr float SPECULATIVE_DISTANCE = 0.02f;
constexpr float PENETRATION_SLOP = 0.02f;
constexpr float BAUMGARTE = 0.2f;
constexpr float MAX_PENETRATION_DISTANCE = 0.2f;
constexpr float MAX_LINEAR_VELOCITY = 500.0f;
constexpr float MAX_ANGULAR_VELOCITY = 47.1238898f;  // 0.25 * pi * 60
constexpr float MIN_VELOCITY_FOR_RESTITUTION = 1.0f;
constexpr float NORMAL_COS_MAX_DELTA = 0.99619470f;  // cos 5 deg
constexpr float PRESERVE_LAMBDA_MAX_DIST_SQ = 1.0e-4f;
constexpr float FACE_AXIS_TOL = 1.0e-4f;
constexpr float EDGE_AXIS_TOL = 2.0e-3f;
constexpr int MAX_POLY = 8;   // a quad or triangle clipped by four planes has at most 8 vertices
constexpr int MAX_SLOTS = 4;                 // manifolds per (body, static body) pair
constexpr int MAX_STATIC_PER_BODY = 8;       // manifolds per body against all static geometry
constexpr int MAX_TRI_CANDIDATES = 24;       // triangles whose box overlaps one body's box

// Per-lane polygon scratch (shared memory in the tick kernel): two ping-pong buffers for face clipping; afterwards
// they hold the contact points on a (buf[0]) and on b (buf[1]).  Thread-local arrays would live in local memory and
// every clip step would pay an L1/L2 round trip.
struct Scratch
{
	v3 buf[2][MAX_POLY];
	float pad;  // 49 words: an odd stride keeps lanes (lane = record) on distinct shared-memory banks
};

struct Hit
{
	v3 n;         // from a to b
	float depth;  // -separation along n
	int np;
	v3 *p1, *p2;  // world points on a and on b (in the caller's Scratch)
};

__device__ __forceinline__ void hit_bind(Hit &h, Scratch &s)
{
	h.p1 = s.buf[0];
	h.p2 = s.buf[1];
}

struct Box
{
	v3 x;
	m33 R;
	v3 he;
};

__device__ __forceinline__ float box_radius(const Box &b, v3 L)
{
	return ((b.he.x * fabsf(dot(b.R.c0, L))) + (b.he.y * fabsf(dot(b.R.c1, L)))) + (b.he.z * fabsf(dot(b.R.c2, L)));
}

__device__ __forceinline__ v3 box_support(const Box &b, v3 dir)
{
	v3 l = mtmul(b.R, dir);
	v3 s = V(l.x < 0.0f ? -b.he.x : b.he.x, l.y < 0.0f ? -b.he.y : b.he.y, l.z < 0.0f ? -b.he.z : b.he.z);
	return b.x + mmul(b.R, s);
}

// supporting face of a box along dir (4 world vertices) and its outward unit normal
__device__ __forceinline__ void box_face(const Box &b, v3 dir, v3 *out4, v3 &nout)
{
	v3 l = mtmul(b.R, dir);
	float ax = fabsf(l.x), ay = fabsf(l.y), az = fabsf(l.z);
	int k = 0;
	if (ay > ax) k = 1;
	if (az > (k == 0 ? ax : ay)) k = 2;
	float s = get(l, k) < 0.0f ? -1.0f : 1.0f;
	int u = (k + 1) % 3, v = (k + 2) % 3;
	v3 ck = col(b.R, k) * (s * get(b.he, k));
	v3 cu = col(b.R, u) * get(b.he, u);
	v3 cv = col(b.R, v) * get(b.he, v);
	v3 c = b.x + ck;
	out4[0] = (c + cu) + cv;
	out4[1] = (c - cu) + cv;
	out4[2] = (c - cu) - cv;
	out4[3] = (c + cu) - cv;
	nout = col(b.R, k) * s;
}

// Sutherland-Hodgman against one plane; keeps (p - origin).normal >= 0.  Not inlined: a face test calls it four times,
// and one copy that stays in the instruction cache beats eight inlined ones (ncu: the inlined clip code was 9 % of the
// tick's stall samples, 60 % of them "no instruction").
static __device__ __noinline__ int clip_plane(const v3 *in, int n, v3 origin, v3 normal, v3 *out)
{
	int m = 0;
	if (n == 0) return 0;
	v3 e1 = in[n - 1];
	float prev = dot(origin - e1, normal);
	bool prev_in = prev < 0.0f;
	for (int i = 0; i < n; i++)
	{
		v3 e2 = in[i];
		float num = dot(origin - e2, normal);
		bool cur_in = num < 0.0f;
		if (cur_in != prev_in)
		{
			v3 e12 = e2 - e1;
			float den = dot(e12, normal);
			if (den != 0.0f)
			{
				if (m < MAX_POLY) out[m++] = e1 + (e12 * (prev / den));
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

// Clip face2 by the side planes of face1 (through face1's edges, parallel to axis); keep points within max_sep of
// face1's plane and project them onto it.  N1, N2: vertex counts (compile time, so the faces stay in registers).
template <int N1, int N2>
__device__ __forceinline__ void manifold_between_faces(const v3 *f1, v3 n1, const v3 *f2, v3 axis, float max_sep,
														   Scratch &sc, Hit &h)
{
	v3 *src = sc.buf[0], *dst = sc.buf[1];
	int n = N2;
#pragma unroll
	for (int i = 0; i < N2; i++) src[i] = f2[i];
	v3 cen = f1[0];
#pragma unroll
	for (int i = 1; i < N1; i++) cen = cen + f1[i];
	cen = cen * (1.0f / (float)N1);
#pragma unroll
	for (int i = 0; i < N1; i++)
	{
		if (n > 0)
		{
			v3 a = f1[i], b = f1[(i + 1) % N1];
			v3 pn = cross(axis, b - a);
			if (dot(cen - a, pn) < 0.0f) pn = -pn;
			n = clip_plane(src, n, a, pn, dst);
			v3 *t = src; src = dst; dst = t;
		}
	}
	// contact points: on b = the clipped vertex, on a = its projection onto face1's plane.  Points on b are compacted
	// in place; points on a go to the other buffer, which clipping no longer needs.
	int np = 0;
	for (int i = 0; i < n; i++)
	{
		v3 p = src[i];
		float dist = dot(p - f1[0], n1);
		if (dist <= max_sep)
		{
			src[np] = p;
			dst[np] = p - (n1 * dist);
			np++;
		}
	}
	h.np = np;
	h.p2 = src;
	h.p1 = dst;
}

// closest points of segments p1 + s d1 (|s| <= h1) and p2 + t d2 (|t| <= h2); unit directions
__device__ __forceinline__ void closest_on_edges(v3 p1, v3 d1, float h1, v3 p2, v3 d2, float h2, v3 &c1, v3 &c2)
{
	v3 r = p1 - p2;
	float b = dot(d1, d2);
	float c = dot(d1, r);
	float f = dot(d2, r);
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
	c1 = p1 + (d1 * s);
	c2 = p2 + (d2 * t);
}

static __device__ __noinline__ bool collide_box_box(const Box &A, const Box &B, float max_sep, Scratch &sc, Hit &h)
{
	v3 d = B.x - A.x;
	float best = -3.0e38f;
	v3 bn = V(0.0f, 1.0f, 0.0f);
	int kind = 0, ei = 0, ej = 0;  // 0: face of A, 1: face of B, 2: edge pair
	for (int i = 0; i < 3; i++)
	{
		v3 L = col(A.R, i);
		float dl = dot(d, L);
		float sep = fabsf(dl) - (get(A.he, i) + box_radius(B, L));
		if (sep > max_sep) return false;
		if (i == 0 || sep > best + FACE_AXIS_TOL)
		{
			best = sep;
			bn = dl < 0.0f ? -L : L;
			kind = 0;
		}
	}
	for (int i = 0; i < 3; i++)
	{
		v3 L = col(B.R, i);
		float dl = dot(d, L);
		float sep = fabsf(dl) - (box_radius(A, L) + get(B.he, i));
		if (sep > max_sep) return false;
		if (sep > best + FACE_AXIS_TOL)
		{
			best = sep;
			bn = dl < 0.0f ? -L : L;
			kind = 1;
		}
	}
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++)
		{
			v3 L = cross(col(A.R, i), col(B.R, j));
			float l2 = len2(L);
			if (l2 < 1.0e-6f) continue;
			L = L * (1.0f / sqrtf(l2));
			float dl = dot(d, L);
			float sep = fabsf(dl) - (box_radius(A, L) + box_radius(B, L));
			if (sep > max_sep) return false;
			if (sep > best + EDGE_AXIS_TOL)
			{
				best = sep;
				bn = dl < 0.0f ? -L : L;
				kind = 2;
				ei = i;
				ej = j;
			}
		}
	h.n = bn;
	h.depth = -best;
	v3 fa[4], fb[4], na, nb;
	box_face(A, bn, fa, na);
	box_face(B, -bn, fb, nb);
	manifold_between_faces<4, 4>(fa, na, fb, bn, max_sep, sc, h);
	if (h.np == 0)
	{
		if (kind == 2)
		{
			v3 pa = A.x, pb = B.x;
			for (int k = 0; k < 3; k++)
			{
				if (k != ei) pa = pa + (col(A.R, k) * (dot(bn, col(A.R, k)) < 0.0f ? -get(A.he, k) : get(A.he, k)));
				if (k != ej) pb = pb - (col(B.R, k) * (dot(bn, col(B.R, k)) < 0.0f ? -get(B.he, k) : get(B.he, k)));
			}
			closest_on_edges(pa, col(A.R, ei), get(A.he, ei), pb, col(B.R, ej), get(B.he, ej), h.p1[0], h.p2[0]);
		}
		else if (kind == 0)
		{
			h.p2[0] = box_support(B, -bn);
			h.p1[0] = h.p2[0] + (bn * (-best));
		}
		else
		{
			h.p1[0] = box_support(A, bn);
			h.p2[0] = h.p1[0] + (bn * best);
		}
		h.np = 1;
	}
	return true;
}

struct Tri
{
	v3 a, b, c, n;
};

static __device__ __noinline__ bool collide_box_tri(const Box &A, const Tri &T, float max_sep, Scratch &sc, Hit &h)
{
	v3 tv[3] = {T.a, T.b, T.c};
	float best;
	v3 bn;
	int kind = 0, ei = 0, ej = 0;
	{
		float r = box_radius(A, T.n);
		float cp = dot(A.x, T.n), tp = dot(T.a, T.n);
		float sp = tp - (cp + r), sm = (cp - r) - tp;
		if (sp > sm) { best = sp; bn = T.n; }
		else { best = sm; bn = -T.n; }
		if (best > max_sep) return false;
	}
	for (int i = 0; i < 3; i++)
	{
		v3 L = col(A.R, i);
		float p0 = dot(tv[0], L), p1 = dot(tv[1], L), p2 = dot(tv[2], L);
		float tmin = fminf(p0, fminf(p1, p2)), tmax = fmaxf(p0, fmaxf(p1, p2));
		float cp = dot(A.x, L), r = get(A.he, i);
		float sp = tmin - (cp + r), sm = (cp - r) - tmax;
		float sep = sp > sm ? sp : sm;
		if (sep > max_sep) return false;
		if (sep > best + FACE_AXIS_TOL)
		{
			best = sep;
			bn = sp > sm ? L : -L;
			kind = 1;
		}
	}
	v3 te[3];
	te[0] = T.b - T.a;
	te[1] = T.c - T.b;
	te[2] = T.a - T.c;
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++)
		{
			v3 L = cross(col(A.R, i), te[j]);
			float l2 = len2(L);
			if (l2 < 1.0e-8f * len2(te[j])) continue;
			L = L * (1.0f / sqrtf(l2));
			float p0 = dot(tv[0], L), p1 = dot(tv[1], L), p2 = dot(tv[2], L);
			float tmin = fminf(p0, fminf(p1, p2)), tmax = fmaxf(p0, fmaxf(p1, p2));
			float cp = dot(A.x, L), r = box_radius(A, L);
			float sp = tmin - (cp + r), sm = (cp - r) - tmax;
			float sep = sp > sm ? sp : sm;
			if (sep > max_sep) return false;
			if (sep > best + EDGE_AXIS_TOL)
			{
				best = sep;
				bn = sp > sm ? L : -L;
				kind = 2;
				ei = i;
				ej = j;
			}
		}
	h.n = bn;
	h.depth = -best;
	v3 fa[4], na;
	box_face(A, bn, fa, na);
	manifold_between_faces<4, 3>(fa, na, tv, bn, max_sep, sc, h);
	if (h.np == 0)
	{
		if (kind == 2)
		{
			v3 pa = A.x;
			for (int k = 0; k < 3; k++)
				if (k != ei) pa = pa + (col(A.R, k) * (dot(bn, col(A.R, k)) < 0.0f ? -get(A.he, k) : get(A.he, k)));
			float el = len(te[ej]);
			v3 ed = te[ej] * (1.0f / el);
			v3 mid = tv[ej] + (te[ej] * 0.5f);
			closest_on_edges(pa, col(A.R, ei), get(A.he, ei), mid, ed, 0.5f * el, h.p1[0], h.p2[0]);
		}
		else
		{
			h.p1[0] = box_support(A, bn);
			h.p2[0] = h.p1[0] + (bn * best);
		}
		h.np = 1;
	}
	return true;
}

// closest point on a triangle (Ericson, Real-Time Collision Detection 5.1.5)
__device__ __forceinline__ v3 closest_on_tri(v3 p, v3 a, v3 b, v3 c)
{
	v3 ab = b - a, ac = c - a, ap = p - a;
	float d1 = dot(ab, ap), d2 = dot(ac, ap);
	if (d1 <= 0.0f && d2 <= 0.0f) return a;
	v3 bp = p - b;
	float d3 = dot(ab, bp), d4 = dot(ac, bp);
	if (d3 >= 0.0f && d4 <= d3) return b;
	float vc = (d1 * d4) - (d3 * d2);
	if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) return a + (ab * (d1 / (d1 - d3)));
	v3 cp = p - c;
	float d5 = dot(ab, cp), d6 = dot(ac, cp);
	if (d6 >= 0.0f && d5 <= d6) return c;
	float vb = (d5 * d2) - (d1 * d6);
	if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) return a + (ac * (d2 / (d2 - d6)));
	float va = (d3 * d6) - (d5 * d4);
	if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) return b + ((c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6))));
	float den = 1.0f / ((va + vb) + vc);
	float v = vb * den, w = vc * den;
	return (a + (ab * v)) + (ac * w);
}

__device__ __forceinline__ bool collide_sphere_tri(v3 x, float r, const Tri &T, float max_sep, Hit &h)
{
	v3 c = closest_on_tri(x, T.a, T.b, T.c);
	v3 d = c - x;
	float dist = len(d);
	if ((dist - r) > max_sep) return false;
	h.n = dist > 1.0e-9f ? d * (1.0f / dist) : -T.n;
	h.depth = r - dist;
	h.np = 1;
	h.p1[0] = x + (h.n * r);
	h.p2[0] = c;
	return true;
}

__device__ __forceinline__ bool collide_sphere_sphere(v3 xa, float ra, v3 xb, float rb, float max_sep, Hit &h)
{
	v3 d = xb - xa;
	float dist = len(d);
	if ((dist - (ra + rb)) > max_sep) return false;
	h.n = dist > 1.0e-9f ? d * (1.0f / dist) : V(0.0f, 1.0f, 0.0f);
	h.depth = (ra + rb) - dist;
	h.np = 1;
	h.p1[0] = xa + (h.n * ra);
	h.p2[0] = xb - (h.n * rb);
	return true;
}

// sphere (sx, r) against box X; normal from the sphere to the box
__device__ __forceinline__ bool collide_sphere_box(v3 sx, float r, const Box &X, float max_sep, Hit &h)
{
	v3 l = mtmul(X.R, sx - X.x);
	v3 cl = V(fminf(fmaxf(l.x, -X.he.x), X.he.x), fminf(fmaxf(l.y, -X.he.y), X.he.y), fminf(fmaxf(l.z, -X.he.z), X.he.z));
	v3 dl = cl - l;
	float dist = len(dl);
	v3 nl, pb;
	float depth;
	if (dist > 1.0e-9f)
	{
		if ((dist - r) > max_sep) return false;
		nl = dl * (1.0f / dist);
		pb = cl;
		depth = r - dist;
	}
	else
	{
		float dx = X.he.x - fabsf(l.x), dy = X.he.y - fabsf(l.y), dz = X.he.z - fabsf(l.z);
		int k = 0;
		float dm = dx;
		if (dy < dm) { dm = dy; k = 1; }
		if (dz < dm) { dm = dz; k = 2; }
		float s = get(l, k) < 0.0f ? -1.0f : 1.0f;
		nl = V(k == 0 ? -s : 0.0f, k == 1 ? -s : 0.0f, k == 2 ? -s : 0.0f);
		pb = l;
		if (k == 0) pb.x = s * X.he.x;
		if (k == 1) pb.y = s * X.he.y;
		if (k == 2) pb.z = s * X.he.z;
		depth = r + dm;
	}
	h.n = mmul(X.R, nl);
	h.depth = depth;
	h.np = 1;
	h.p1[0] = sx + (h.n * r);
	h.p2[0] = X.x + mmul(X.R, pb);
	return true;
}

// Keep <= 4 points: the one with most leverage x depth, the farthest from it, and the extremes on both sides of
// that segment (the reduction Jolt documents for PruneContactPoints).  Projections are re-derived per pass instead of
// being parked in thread-local arrays; p1/p2 may point at shared or local memory.
__device__ __forceinline__ v3 prune_proj(v3 xa, v3 axis, v3 p1)
{
	v3 v1 = p1 - xa;
	return v1 - (axis * dot(v1, axis));
}
__device__ __forceinline__ float prune_dsq(v3 p1, v3 p2) { return fmaxf(1.0e-6f, len2(p2 - p1)); }

__device__ __forceinline__ void prune_points(v3 xa, v3 axis, int &np, v3 *p1, v3 *p2)
{
	const int n = np;
	if (n <= 4) return;
	int i1 = 0;
	float best = -1.0f;
	for (int i = 0; i < n; i++)
	{
		v3 a = p1[i];
		float v = fmaxf(1.0e-6f, len2(prune_proj(xa, axis, a))) * prune_dsq(a, p2[i]);
		if (v > best) { best = v; i1 = i; }
	}
	const v3 proj1 = prune_proj(xa, axis, p1[i1]);
	int i2 = -1;
	best = -1.0f;
	for (int i = 0; i < n; i++)
		if (i != i1)
		{
			v3 a = p1[i];
			float v = fmaxf(1.0e-6f, len2(prune_proj(xa, axis, a) - proj1)) * prune_dsq(a, p2[i]);
			if (v > best) { best = v; i2 = i; }
		}
	int i3 = -1, i4 = -1;
	float mn = 0.0f, mx = 0.0f;
	v3 perp = cross(prune_proj(xa, axis, p1[i2]) - proj1, axis);
	for (int i = 0; i < n; i++)
		if (i != i1 && i != i2)
		{
			float v = dot(perp, prune_proj(xa, axis, p1[i]) - proj1);
			if (v < mn) { mn = v; i3 = i; }
			else if (v > mx) { mx = v; i4 = i; }
		}
	// output order: i1, (i3), i2, (i4)
	const int e0 = i1, e1 = i3 >= 0 ? i3 : i2, e2 = i3 >= 0 ? i2 : i4, e3 = i3 >= 0 ? i4 : -1;
	const int m = (i3 >= 0 ? 3 : 2) + (i4 >= 0 ? 1 : 0);
	const v3 a0 = p1[e0], b0 = p2[e0], a1 = p1[e1], b1 = p2[e1];
	const v3 a2 = p1[e2 >= 0 ? e2 : 0], b2 = p2[e2 >= 0 ? e2 : 0], a3 = p1[e3 >= 0 ? e3 : 0], b3 = p2[e3 >= 0 ? e3 : 0];
	p1[0] = a0; p2[0] = b0;
	p1[1] = a1; p2[1] = b1;
	if (m > 2) { p1[2] = a2; p2[2] = b2; }
	if (m > 3) { p1[3] = a3; p2[3] = b3; }
	np = m;
}

}  // namespace gpx
