// glm-compat shim for building the ORACLE only (oracle/_ref): glm is a third-party
// dependency of the reference (CMakeLists.txt:34 `find_package(glm REQUIRED)`, un-pinned;
// README.md:18,23 installs distro libglm-dev = 0.9.9.x) and is neither vendored under
// /root/reference nor installed in this image.  This header restates the handful of glm
// entities the reference's cuda_rasterizer uses, following glm 0.9.9's published formulae:
//   * mat3 is column-major, m[col][row]; the 9-scalar constructor fills columns.
//   * (A*B)[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2]   (type_mat3x3.inl)
//   * dot(a,b) = a.x*b.x + a.y*b.y + a.z*b.z ; length(v) = sqrt(dot(v,v))   (func_geometric.inl)
//   * v / s divides per component; max(v, s) is per-component std::max, i.e. (v < s ? s : v).
// Test infrastructure: never included by the product (omnigs-fork_b200/csrc).
#pragma once
#include <cmath>
#if defined(__CUDACC__)
#define OGS_GLM_Q __host__ __device__ inline
#else
#define OGS_GLM_Q inline
#endif
namespace glm {

struct vec3 {
	float x, y, z;
	OGS_GLM_Q vec3() : x(0.f), y(0.f), z(0.f) {}
	OGS_GLM_Q vec3(float a, float b, float c) : x(a), y(b), z(c) {}
	OGS_GLM_Q explicit vec3(float s) : x(s), y(s), z(s) {}
	OGS_GLM_Q float& operator[](int i) { return (&x)[i]; }
	OGS_GLM_Q const float& operator[](int i) const { return (&x)[i]; }
	OGS_GLM_Q vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
	OGS_GLM_Q vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
	OGS_GLM_Q vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4 {
	float x, y, z, w;
	OGS_GLM_Q vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
	OGS_GLM_Q vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};

OGS_GLM_Q vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
OGS_GLM_Q vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
OGS_GLM_Q vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
OGS_GLM_Q vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
OGS_GLM_Q vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
OGS_GLM_Q float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
OGS_GLM_Q float length(const vec3& a) { return sqrtf(dot(a, a)); }
OGS_GLM_Q vec3 max(const vec3& a, float s) { return vec3(a.x < s ? s : a.x, a.y < s ? s : a.y, a.z < s ? s : a.z); }

struct mat3 {
	vec3 col[3];
	OGS_GLM_Q mat3() {}
	OGS_GLM_Q explicit mat3(float d) { col[0] = vec3(d, 0.f, 0.f); col[1] = vec3(0.f, d, 0.f); col[2] = vec3(0.f, 0.f, d); }
	OGS_GLM_Q mat3(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2)
	{ col[0] = vec3(x0, y0, z0); col[1] = vec3(x1, y1, z1); col[2] = vec3(x2, y2, z2); }
	OGS_GLM_Q vec3& operator[](int i) { return col[i]; }
	OGS_GLM_Q const vec3& operator[](int i) const { return col[i]; }
};

OGS_GLM_Q mat3 transpose(const mat3& m)
{
	return mat3(m[0][0], m[1][0], m[2][0],
	            m[0][1], m[1][1], m[2][1],
	            m[0][2], m[1][2], m[2][2]);
}

OGS_GLM_Q mat3 operator*(const mat3& A, const mat3& B)
{
	mat3 R;
	for (int c = 0; c < 3; c++)
		for (int r = 0; r < 3; r++)
			R[c][r] = A[0][r] * B[c][0] + A[1][r] * B[c][1] + A[2][r] * B[c][2];
	return R;
}

OGS_GLM_Q mat3 operator*(float s, const mat3& A)
{
	mat3 R;
	R[0] = A[0] * s; R[1] = A[1] * s; R[2] = A[2] * s;
	return R;
}

} // namespace glm
