/* orc_math.h — scalar fp32 vector helpers for the CPU oracle (test infrastructure only; see orc.h).
 * Every expression is fully parenthesised: with -ffp-contract=off the compiler never contracts on its own, and the
 * fused multiply-adds of dot / cross / matrix-vector / quaternion products are spelled fmaf(), so the rounding
 * sequence is fixed (fmaf is exact-then-round on every conforming libm / FMA unit). */
#ifndef ORC_MATH_H
#define ORC_MATH_H
#include <math.h>

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } q4;
typedef struct { v3 c0, c1, c2; } m33; /* columns = rotated basis vectors */

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float vdot(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline v3 vcross(v3 a, v3 b)
{
	return V(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
static inline float vlen2(v3 a) { return vdot(a, a); }
static inline float vlen(v3 a) { return sqrtf(vdot(a, a)); }
static inline float vget(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
/* v + t * s */
static inline v3 vmadd(v3 v, v3 t, float s) { return V(fmaf(t.x, s, v.x), fmaf(t.y, s, v.y), fmaf(t.z, s, v.z)); }
static inline v3 vmulc(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }

static inline v3 qrot(q4 q, v3 v)
{
	v3 u = V(q.x, q.y, q.z);
	v3 t = vscale(vcross(u, v), 2.0f);
	return vadd(vmadd(v, t, q.w), vcross(u, t));
}
static inline q4 qconj(q4 q) { q4 r = {-q.x, -q.y, -q.z, q.w}; return r; }
static inline q4 qmul(q4 a, q4 b)
{
	q4 r;
	r.x = fmaf(-a.z, b.y, fmaf(a.y, b.z, fmaf(a.x, b.w, a.w * b.x)));
	r.y = fmaf(a.z, b.x, fmaf(a.y, b.w, fmaf(-a.x, b.z, a.w * b.y)));
	r.z = fmaf(a.z, b.w, fmaf(-a.y, b.x, fmaf(a.x, b.y, a.w * b.z)));
	r.w = fmaf(-a.z, b.z, fmaf(-a.y, b.y, fmaf(-a.x, b.x, a.w * b.w)));
	return r;
}
static inline q4 qnormalize(q4 q)
{
	float l = sqrtf((((q.x * q.x) + (q.y * q.y)) + (q.z * q.z)) + (q.w * q.w));
	float inv = 1.0f / l;
	q4 r = {q.x * inv, q.y * inv, q.z * inv, q.w * inv};
	return r;
}
static inline m33 qmat(q4 q)
{
	m33 m;
	m.c0 = qrot(q, V(1, 0, 0));
	m.c1 = qrot(q, V(0, 1, 0));
	m.c2 = qrot(q, V(0, 0, 1));
	return m;
}
/* world = M * local */
static inline v3 mmul(const m33 *m, v3 v)
{
	return vmadd(vmadd(vscale(m->c0, v.x), m->c1, v.y), m->c2, v.z);
}
/* local = M^T * world */
static inline v3 mtmul(const m33 *m, v3 v) { return V(vdot(m->c0, v), vdot(m->c1, v), vdot(m->c2, v)); }

/* sin and cos of a small angle (|a| <= ~0.8) by fixed-order polynomials: identical on CPU and GPU.
 * Rotation steps are at most max_angular_velocity * h = 47.1/120 rad, half angle <= 0.2. */
static inline void small_sincos(float a, float *s, float *c)
{
	float a2 = a * a;
	float ps = 1.0f + (a2 * (-1.0f / 6.0f + (a2 * (1.0f / 120.0f + (a2 * (-1.0f / 5040.0f + (a2 * (1.0f / 362880.0f))))))));
	float pc = 1.0f + (a2 * (-0.5f + (a2 * (1.0f / 24.0f + (a2 * (-1.0f / 720.0f + (a2 * (1.0f / 40320.0f))))))));
	*s = a * ps;
	*c = pc;
}
/* q' = normalize(rotation(axis = d/|d|, angle = |d|) * q); no-op when |d| <= 1e-6 */
static inline q4 qstep(q4 q, v3 d)
{
	float len = vlen(d);
	if (len > 1.0e-6f)
	{
		float s, c;
		small_sincos(0.5f * len, &s, &c);
		float k = s / len;
		q4 r = {d.x * k, d.y * k, d.z * k, c};
		return qnormalize(qmul(r, q));
	}
	return q;
}
/* Perpendicular used for friction tangents */
static inline v3 vperp(v3 n)
{
	if (fabsf(n.x) > fabsf(n.y))
	{
		float l = sqrtf((n.x * n.x) + (n.z * n.z));
		return V(n.z / l, 0.0f, -n.x / l);
	}
	float l = sqrtf((n.y * n.y) + (n.z * n.z));
	return V(0.0f, n.z / l, -n.y / l);
}
#endif
