// glm.hpp — the sliver of GLM (g-truc/glm 1.0.1, pinned by the reference's CMakeLists.txt:17, absent from this image)
// that the reference's raster path touches, written from the library's documented semantics. TEST INFRASTRUCTURE: it only
// exists so that oracle/ref_build.py can compile the reference's own translation units with g++ (see simd_gxx.h).
// Column-major matrices (m[col][row]); products are evaluated in GLM's scalar order:
//   mat4 * mat4 : column c = ((a[0]*b[c].x + a[1]*b[c].y) + a[2]*b[c].z) + a[3]*b[c].w
//   mat4 * vec4 : (m[0]*v.x + m[1]*v.y) + (m[2]*v.z + m[3]*v.w)
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>

namespace glm {

template<int L, typename T> struct vec;

template<typename T> struct vec<2, T> {
    T x, y;
    constexpr vec() : x(0), y(0) {}
    constexpr explicit vec(T s) : x(s), y(s) {}
    template<typename A, typename B> constexpr vec(A x_, B y_) : x(static_cast<T>(x_)), y(static_cast<T>(y_)) {}
    template<typename U> constexpr vec(const vec<2, U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)) {}
    template<typename U> constexpr explicit vec(const vec<3, U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)) {}
    template<typename U> constexpr explicit vec(const vec<4, U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)) {}
    constexpr T& operator[](size_t i) { return i == 0 ? x : y; }
    constexpr const T& operator[](size_t i) const { return i == 0 ? x : y; }
};
template<typename T> struct vec<3, T> {
    T x, y, z;
    constexpr vec() : x(0), y(0), z(0) {}
    constexpr explicit vec(T s) : x(s), y(s), z(s) {}
    template<typename A, typename B, typename C> constexpr vec(A x_, B y_, C z_) : x(static_cast<T>(x_)), y(static_cast<T>(y_)), z(static_cast<T>(z_)) {}
    template<typename U, typename C> constexpr vec(const vec<2, U>& xy, C z_) : x(static_cast<T>(xy.x)), y(static_cast<T>(xy.y)), z(static_cast<T>(z_)) {}
    template<typename U> constexpr vec(const vec<3, U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)), z(static_cast<T>(o.z)) {}
    template<typename U> constexpr explicit vec(const vec<4, U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)), z(static_cast<T>(o.z)) {}
    constexpr T& operator[](size_t i) { return i == 0 ? x : (i == 1 ? y : z); }
    constexpr const T& operator[](size_t i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template<typename T> struct vec<4, T> {
    T x, y, z, w;
    constexpr vec() : x(0), y(0), z(0), w(0) {}
    constexpr explicit vec(T s) : x(s), y(s), z(s), w(s) {}
    template<typename A, typename B, typename C, typename D>
    constexpr vec(A x_, B y_, C z_, D w_) : x(static_cast<T>(x_)), y(static_cast<T>(y_)), z(static_cast<T>(z_)), w(static_cast<T>(w_)) {}
    template<typename U, typename D> constexpr vec(const vec<3, U>& xyz, D w_) : x(static_cast<T>(xyz.x)), y(static_cast<T>(xyz.y)), z(static_cast<T>(xyz.z)), w(static_cast<T>(w_)) {}
    template<typename U> constexpr vec(const vec<4, U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)), z(static_cast<T>(o.z)), w(static_cast<T>(o.w)) {}
    constexpr T& operator[](size_t i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    constexpr const T& operator[](size_t i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};

// component-wise operators, vector (op) vector / vector (op) scalar / scalar (op) vector
#define GLM_STUB_OP(sym)                                                                                                      \
    template<int L, typename T> constexpr vec<L, T> operator sym(const vec<L, T>& a, const vec<L, T>& b) {                      \
        vec<L, T> r; for (int i = 0; i < L; i++) r[i] = static_cast<T>(a[i] sym b[i]); return r; }                              \
    template<int L, typename T, typename S> requires std::is_arithmetic_v<S>                                                    \
    constexpr vec<L, T> operator sym(const vec<L, T>& a, S b) { vec<L, T> r; for (int i = 0; i < L; i++) r[i] = static_cast<T>(a[i] sym static_cast<T>(b)); return r; } \
    template<int L, typename T, typename S> requires std::is_arithmetic_v<S>                                                    \
    constexpr vec<L, T> operator sym(S a, const vec<L, T>& b) { vec<L, T> r; for (int i = 0; i < L; i++) r[i] = static_cast<T>(static_cast<T>(a) sym b[i]); return r; } \
    template<int L, typename T> constexpr vec<L, T>& operator sym##=(vec<L, T>& a, const vec<L, T>& b) { return a = a sym b; }  \
    template<int L, typename T, typename S> requires std::is_arithmetic_v<S> constexpr vec<L, T>& operator sym##=(vec<L, T>& a, S b) { return a = a sym b; }
}  // namespace glm
#include <type_traits>
namespace glm {
GLM_STUB_OP(+) GLM_STUB_OP(-) GLM_STUB_OP(*) GLM_STUB_OP(/)
#undef GLM_STUB_OP
#define GLM_STUB_IOP(sym)                                                                                                     \
    template<int L, typename T> requires std::is_integral_v<T> constexpr vec<L, T> operator sym(const vec<L, T>& a, const vec<L, T>& b) { \
        vec<L, T> r; for (int i = 0; i < L; i++) r[i] = static_cast<T>(a[i] sym b[i]); return r; }                              \
    template<int L, typename T, typename S> requires(std::is_integral_v<T> && std::is_integral_v<S>)                            \
    constexpr vec<L, T> operator sym(const vec<L, T>& a, S b) { vec<L, T> r; for (int i = 0; i < L; i++) r[i] = static_cast<T>(a[i] sym static_cast<T>(b)); return r; }
GLM_STUB_IOP(&) GLM_STUB_IOP(|) GLM_STUB_IOP(^) GLM_STUB_IOP(<<) GLM_STUB_IOP(>>)
#undef GLM_STUB_IOP
template<int L, typename T> constexpr vec<L, T> operator-(const vec<L, T>& a) { vec<L, T> r; for (int i = 0; i < L; i++) r[i] = -a[i]; return r; }
template<int L, typename T> constexpr vec<L, T> operator~(const vec<L, T>& a) { vec<L, T> r; for (int i = 0; i < L; i++) r[i] = ~a[i]; return r; }
template<int L, typename T> constexpr bool operator==(const vec<L, T>& a, const vec<L, T>& b) { for (int i = 0; i < L; i++) if (a[i] != b[i]) return false; return true; }

template<typename T> requires std::is_arithmetic_v<T> constexpr T abs(T x) { return x < 0 ? -x : x; }
template<typename T> requires std::is_arithmetic_v<T> constexpr T min(T a, T b) { return b < a ? b : a; }
template<typename T> requires std::is_arithmetic_v<T> constexpr T max(T a, T b) { return a < b ? b : a; }
template<int L, typename T> constexpr vec<L, T> min(const vec<L, T>& a, const vec<L, T>& b) { vec<L, T> r; for (int i = 0; i < L; i++) r[i] = min(a[i], b[i]); return r; }
template<int L, typename T> constexpr vec<L, T> max(const vec<L, T>& a, const vec<L, T>& b) { vec<L, T> r; for (int i = 0; i < L; i++) r[i] = max(a[i], b[i]); return r; }
template<int L, typename T, typename S> requires std::is_arithmetic_v<S> constexpr vec<L, T> min(const vec<L, T>& a, S b) { return min(a, vec<L, T>(static_cast<T>(b))); }
template<int L, typename T, typename S> requires std::is_arithmetic_v<S> constexpr vec<L, T> max(const vec<L, T>& a, S b) { return max(a, vec<L, T>(static_cast<T>(b))); }
template<int L, typename T> T dot(const vec<L, T>& a, const vec<L, T>& b) { T s = a[0] * b[0]; for (int i = 1; i < L; i++) s = s + a[i] * b[i]; return s; }
template<int L, typename T> T length(const vec<L, T>& a) { return std::sqrt(dot(a, a)); }
template<int L, typename T> vec<L, T> normalize(const vec<L, T>& a) { return a * (T(1) / std::sqrt(dot(a, a))); }
template<typename T> vec<3, T> cross(const vec<3, T>& a, const vec<3, T>& b) { return { a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y }; }

// ---- matrices: C columns of R rows -------------------------------------------------------------------------------------------
template<int C, int R, typename T> struct mat {
    using col_type = vec<R, T>;
    col_type col[C];
    constexpr mat() : col{} {}
    constexpr explicit mat(T diag) : col{} { for (int i = 0; i < (C < R ? C : R); i++) col[i][i] = diag; }
    template<int C2, int R2> requires(C2 != C || R2 != R)
    constexpr mat(const mat<C2, R2, T>& o) : col{} {       // crop / pad with identity; implicit like GLM without GLM_FORCE_EXPLICIT_CTOR
        for (int c = 0; c < C; c++) for (int r = 0; r < R; r++) col[c][r] = (c < C2 && r < R2) ? o.col[c][r] : (c == r ? T(1) : T(0));
    }
    constexpr col_type& operator[](size_t c) { return col[c]; }
    constexpr const col_type& operator[](size_t c) const { return col[c]; }
};
template<typename T> mat<4, 4, T> operator*(const mat<4, 4, T>& a, const mat<4, 4, T>& b) {
    mat<4, 4, T> r;
    for (int c = 0; c < 4; c++) r[c] = ((a[0] * b[c][0] + a[1] * b[c][1]) + a[2] * b[c][2]) + a[3] * b[c][3];
    return r;
}
template<typename T> vec<4, T> operator*(const mat<4, 4, T>& m, const vec<4, T>& v) {
    return (m[0] * v.x + m[1] * v.y) + (m[2] * v.z + m[3] * v.w);
}
template<typename T> vec<3, T> operator*(const mat<3, 3, T>& m, const vec<3, T>& v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z; }
template<int C, int R, typename T> mat<R, C, T> transpose(const mat<C, R, T>& m) {
    mat<R, C, T> r;
    for (int c = 0; c < C; c++) for (int rr = 0; rr < R; rr++) r[rr][c] = m[c][rr];
    return r;
}

// affine post-multiplications (ext/matrix_transform.hpp): m * T(v), m * S(v)
template<typename T> mat<4, 4, T> translate(const mat<4, 4, T>& m, const vec<3, T>& v) {
    mat<4, 4, T> r = m;
    r[3] = ((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3];
    return r;
}
template<typename T> mat<4, 4, T> scale(const mat<4, 4, T>& m, const vec<3, T>& v) {
    mat<4, 4, T> r = m;
    r[0] = m[0] * v.x; r[1] = m[1] * v.y; r[2] = m[2] * v.z;
    return r;
}
// inverse by the adjugate: inv[c][r] = (-1)^(r+c) * minor(c, r) / det, every 3x3 minor expanded along its first row.
// (GLM evaluates the same cofactors from shared 2x2 sub-determinants; the two agree to rounding, not bit for bit.)
template<typename T> mat<4, 4, T> inverse(const mat<4, 4, T>& m) {
    auto minor3 = [&](int skipCol, int skipRow) {
        T a[3][3];
        for (int c = 0, cc = 0; c < 4; c++) {
            if (c == skipCol) continue;
            for (int r = 0, rr = 0; r < 4; r++) { if (r == skipRow) continue; a[cc][rr++] = m[c][r]; }
            cc++;
        }
        return a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2]) - a[1][0] * (a[0][1] * a[2][2] - a[2][1] * a[0][2]) +
               a[2][0] * (a[0][1] * a[1][2] - a[1][1] * a[0][2]);
    };
    mat<4, 4, T> adj;
    for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) adj[c][r] = (((r + c) & 1) ? T(-1) : T(1)) * minor3(r, c);
    T det = (m[0][0] * adj[0][0] + m[1][0] * adj[0][1]) + (m[2][0] * adj[0][2] + m[3][0] * adj[0][3]);
    T inv = T(1) / det;
    for (int c = 0; c < 4; c++) adj[c] = adj[c] * inv;
    return adj;
}

using vec2 = vec<2, float>; using vec3 = vec<3, float>; using vec4 = vec<4, float>;
using ivec2 = vec<2, int32_t>; using ivec3 = vec<3, int32_t>; using ivec4 = vec<4, int32_t>;
using uvec2 = vec<2, uint32_t>; using uvec3 = vec<3, uint32_t>; using uvec4 = vec<4, uint32_t>;
using mat3 = mat<3, 3, float>; using mat4 = mat<4, 4, float>; using mat4x4 = mat4; using mat3x3 = mat3;
using mat4x3 = mat<4, 3, float>;

}  // namespace glm
