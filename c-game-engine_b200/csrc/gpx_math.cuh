// gpx_math.cuh — fp32 vector helpers for the sm_100a kernels.
// Compiled with -fmad=false: the compiler never contracts on its own, every expression is parenthesised, and the
// fused multiply-adds that matter (dot, cross, matrix-vector, quaternion products: about a third of the tick's fp32
// instructions) are written out as fmaf().  The rounding sequence is therefore fixed and the kernels are reproducible
// run to run and comparable bit-for-bit with a scalar evaluation of the same formulas.
#pragma once
#include <cuda_runtime.h>

namespace gpx {

struct v3 { float x, y, z; };
struct q4 { float x, y, z, w; };
struct m33 { v3 c0, c1, c2; };  // columns: images of the basis vectors

#define GPX_HD __host__ __device__ __forceinline__

GPX_HD v3 V(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
GPX_HD v3 V(const float4 &f) { return V(f.x, f.y, f.z); }
GPX_HD v3 operator+(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
GPX_HD v3 operator-(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
GPX_HD v3 operator-(v3 a) { return V(-a.x, -a.y, -a.z); }
GPX_HD v3 operator*(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
GPX_HD float dot(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
GPX_HD v3 cross(v3 a, v3 b)
{
	return V(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
GPX_HD float len2(v3 a) { return dot(a, a); }
GPX_HD float len(v3 a) { return sqrtf(dot(a, a)); }
GPX_HD float get(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
GPX_HD v3 mulc(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
// v + t * s
GPX_HD v3 madd(v3 v, v3 t, float s) { return V(fmaf(t.x, s, v.x), fmaf(t.y, s, v.y), fmaf(t.z, s, v.z)); }
GPX_HD float4 F4(v3 a, float w) { return make_float4(a.x, a.y, a.z, w); }

GPX_HD v3 qrot(q4 q, v3 v)
{
	v3 u = V(q.x, q.y, q.z);
	v3 t = cross(u, v) * 2.0f;
	return madd(v, t, q.w) + cross(u, t);
}
GPX_HD q4 Q(const float4 &f) { q4 q; q.x = f.x; q.y = f.y; q.z = f.z; q.w = f.w; return q; }
GPX_HD q4 qmul(q4 a, q4 b)
{
	q4 r;
	r.x = fmaf(-a.z, b.y, fmaf(a.y, b.z, fmaf(a.x, b.w, a.w * b.x)));
	r.y = fmaf(a.z, b.x, fmaf(a.y, b.w, fmaf(-a.x, b.z, a.w * b.y)));
	r.z = fmaf(a.z, b.w, fmaf(-a.y, b.x, fmaf(a.x, b.y, a.w * b.z)));
	r.w = fmaf(-a.z, b.z, fmaf(-a.y, b.y, fmaf(-a.x, b.x, a.w * b.w)));
	return r;
}
GPX_HD q4 qnormalize(q4 q)
{
	float l = sqrtf((((q.x * q.x) + (q.y * q.y)) + (q.z * q.z)) + (q.w * q.w));
	float inv = 1.0f / l;
	q4 r; r.x = q.x * inv; r.y = q.y * inv; r.z = q.z * inv; r.w = q.w * inv;
	return r;
}
GPX_HD m33 qmat(q4 q)
{
	m33 m;
	m.c0 = qrot(q, V(1.0f, 0.0f, 0.0f));
	m.c1 = qrot(q, V(0.0f, 1.0f, 0.0f));
	m.c2 = qrot(q, V(0.0f, 0.0f, 1.0f));
	return m;
}
GPX_HD v3 mmul(const m33 &m, v3 v) { return madd(madd(m.c0 * v.x, m.c1, v.y), m.c2, v.z); }
GPX_HD v3 mtmul(const m33 &m, v3 v) { return V(dot(m.c0, v), dot(m.c1, v), dot(m.c2, v)); }
GPX_HD v3 col(const m33 &m, int i) { return i == 0 ? m.c0 : (i == 1 ? m.c1 : m.c2); }

// sin/cos of a small angle by fixed polynomials (rotation steps are <= 47.1/120 rad, half angle <= 0.2)
GPX_HD void small_sincos(float a, float *s, float *c)
{
	float a2 = a * a;
	float ps = 1.0f + (a2 * (-1.0f / 6.0f + (a2 * (1.0f / 120.0f + (a2 * (-1.0f / 5040.0f + (a2 * (1.0f / 362880.0f))))))));
	float pc = 1.0f + (a2 * (-0.5f + (a2 * (1.0f / 24.0f + (a2 * (-1.0f / 720.0f + (a2 * (1.0f / 40320.0f))))))));
	*s = a * ps;
	*c = pc;
}
// q' = normalize(rotation(d/|d|, |d|) * q)
GPX_HD q4 qstep(q4 q, v3 d)
{
	float l = len(d);
	if (l > 1.0e-6f)
	{
		float s, c;
		small_sincos(0.5f * l, &s, &c);
		float k = s / l;
		q4 r; r.x = d.x * k; r.y = d.y * k; r.z = d.z * k; r.w = c;
		return qnormalize(qmul(r, q));
	}
	return q;
}
GPX_HD v3 vperp(v3 n)
{
	if (fabsf(n.x) > fabsf(n.y))
	{
		float l = sqrtf((n.x * n.x) + (n.z * n.z));
		return V(n.z / l, 0.0f, -n.x / l);
	}
	float l = sqrtf((n.y * n.y) + (n.z * n.z));
	return V(0.0f, n.z / l, -n.y / l);
}
// symmetric 3x3 (xx xy xz yy yz zz) times vector
GPX_HD v3 sym_mul(const float *M, v3 v)
{
	return V(fmaf(M[2], v.z, fmaf(M[1], v.y, M[0] * v.x)), fmaf(M[4], v.z, fmaf(M[3], v.y, M[1] * v.x)),
			 fmaf(M[5], v.z, fmaf(M[4], v.y, M[2] * v.x)));
}

}  // namespace gpx
