// simd_gxx.h — g++ stand-in for the reference's SIMD layer, so that the reference's OWN translation units
// (src/SwRast/Rasterizer.cpp, Shading.cpp, ImageHelpers.cpp + their headers) compile with the g++ of this image.
//
// TEST INFRASTRUCTURE (oracle/): the reference is written against Clang vector extensions (`E [[clang::ext_vector_type(N)]]`,
// bool vectors, __builtin_elementwise_*, _BitInt) which g++ does not have. oracle/ref_build.py compiles the reference
// sources where they lie under /root/reference with this header standing in for src/SwRast/SIMD.h: same names, same
// call signatures, same lane semantics — re-implemented from scratch on a plain array-of-lanes class. What that buys:
//   * every floating-point operator is evaluated per lane in IEEE binary32 (no -ffast-math, -ffp-contract=off), i.e. the
//     canonical arithmetic of SURVEY.md App. A; simd::fma is std::fma per lane — fused exactly where the source fuses;
//   * the integer / permute / saturating intrinsics the sources call directly (`_mm512_*`) still run as the real AVX-512
//     instructions (vectors convert to and from __m512i / __m512 / __m128i), so their bit-level behaviour is the hardware's;
//   * approx_rcp / approx_rsqrt are IEEE 1/x and 1/sqrt(x) by default (the canonical choice of the oracle) and the real
//     vrcp14ps / vrsqrt14ps with -DSWR_COMPAT_RCP14 (what an upstream binary computes), for the sensitivity comparison.
// Nothing here is used by the product.
#pragma once

#include <algorithm>
#include <bit>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <type_traits>
#include <utility>

#include <glm/glm.hpp>

#include <immintrin.h>

// Clang declares the streaming stores with `void*` destinations, g++ with `__m512i*` / `__m128i*`; the sources pass uint32_t*.
// (A function-like macro is not re-expanded inside its own expansion, so these forward to the real intrinsics.)
#define _mm512_stream_si512(p, v) _mm512_stream_si512(reinterpret_cast<__m512i*>(p), (v))
#define _mm_stream_si128(p, v) _mm_stream_si128(reinterpret_cast<__m128i*>(p), (v))
inline int _mm_tzcnt_32(unsigned int x) { return static_cast<int>(__tzcnt_u32(x)); }   // Clang's spelling of _tzcnt_u32

#define SIMD_INLINE inline

namespace simd {

constexpr int vec_width = 16;

namespace detail {
template<int Bytes> struct sint;
template<> struct sint<1> { using type = int8_t; };
template<> struct sint<2> { using type = int16_t; };
template<> struct sint<4> { using type = int32_t; };
template<> struct sint<8> { using type = int64_t; };
template<typename E> using mask_elem = typename sint<sizeof(E)>::type;
template<typename E> using uint_of = std::make_unsigned_t<mask_elem<E>>;
constexpr size_t vec_align(size_t bytes) { return bytes >= 64 ? 64 : (bytes >= 32 ? 32 : (bytes >= 16 ? 16 : bytes)); }
}  // namespace detail

// ---- the vector type: N lanes of E --------------------------------------------------------------------------------------
template<typename E, int N = vec_width>
struct alignas(detail::vec_align(sizeof(E) * N)) vec {
    using elem = E;
    using mask_t = vec<detail::mask_elem<E>, N>;
    static constexpr int width = N;
    static constexpr bool is_simd_vec = true;

    E lane[N];

    constexpr vec() : lane{} {}
    // scalar -> every lane (Clang converts scalars to vectors implicitly)
    template<typename S> requires std::is_arithmetic_v<S>
    constexpr vec(S s) : lane{} { for (int i = 0; i < N; i++) lane[i] = static_cast<E>(s); }
    // one value per lane
    template<typename... A> requires(sizeof...(A) == N && N > 1 && (std::is_arithmetic_v<A> && ...))
    constexpr vec(A... a) : lane{ static_cast<E>(a)... } {}
    // C-style / functional cast between vectors of the same lane size and count: a bit cast, like Clang's
    template<typename F> requires(!std::is_same_v<E, F> && sizeof(F) == sizeof(E))
    constexpr explicit vec(const vec<F, N>& o) : lane{} { for (int i = 0; i < N; i++) lane[i] = std::bit_cast<E>(o.lane[i]); }
    // lane mask -> bool vector (`v_bool(a < b)` in Clipper::ComputeClipCodes): a byte mask stands in for Clang's bool vector
    template<typename F> requires(std::is_same_v<E, int8_t> && std::is_signed_v<F> && std::is_integral_v<F> && sizeof(F) > 1)
    constexpr explicit vec(const vec<F, N>& o) : lane{} { for (int i = 0; i < N; i++) lane[i] = o.lane[i] != 0 ? -1 : 0; }

    // raw registers
    vec(__m512i r) requires(sizeof(E) * N == 64) { std::memcpy(lane, &r, 64); }
    vec(__m512 r) requires(sizeof(E) * N == 64) { std::memcpy(lane, &r, 64); }
    vec(__m256i r) requires(sizeof(E) * N == 32) { std::memcpy(lane, &r, 32); }
    vec(__m128i r) requires(sizeof(E) * N == 16) { std::memcpy(lane, &r, 16); }
    operator __m512i() const requires(sizeof(E) * N == 64) { __m512i r; std::memcpy(&r, lane, 64); return r; }
    operator __m512() const requires(sizeof(E) * N == 64) { __m512 r; std::memcpy(&r, lane, 64); return r; }
    operator __m256i() const requires(sizeof(E) * N == 32) { __m256i r; std::memcpy(&r, lane, 32); return r; }
    operator __m128i() const requires(sizeof(E) * N == 16) { __m128i r; std::memcpy(&r, lane, 16); return r; }

    constexpr E& operator[](size_t i) { return lane[i]; }
    constexpr const E& operator[](size_t i) const { return lane[i]; }

#define SIMD_GXX_ARITH(sym)                                                                                     \
    friend constexpr vec operator sym(const vec& a, const vec& b) {                                               \
        vec r;                                                                                                    \
        for (int i = 0; i < N; i++) r.lane[i] = static_cast<E>(a.lane[i] sym b.lane[i]);                          \
        return r;                                                                                                 \
    }                                                                                                             \
    friend constexpr vec& operator sym##=(vec& a, const vec& b) { return a = a sym b; }
    SIMD_GXX_ARITH(+) SIMD_GXX_ARITH(-) SIMD_GXX_ARITH(*) SIMD_GXX_ARITH(/)
#undef SIMD_GXX_ARITH
#define SIMD_GXX_BITS(sym)                                                                                      \
    friend constexpr vec operator sym(const vec& a, const vec& b) requires std::is_integral_v<E> {                \
        vec r;                                                                                                    \
        for (int i = 0; i < N; i++) r.lane[i] = static_cast<E>(a.lane[i] sym b.lane[i]);                          \
        return r;                                                                                                 \
    }                                                                                                             \
    friend constexpr vec& operator sym##=(vec& a, const vec& b) requires std::is_integral_v<E> { return a = a sym b; }
    SIMD_GXX_BITS(|) SIMD_GXX_BITS(&) SIMD_GXX_BITS(^)
#undef SIMD_GXX_BITS
    // shifts: left shifts go through the unsigned type (wrapping), right shifts are arithmetic for signed lanes
    friend constexpr vec operator<<(const vec& a, const vec& b) requires std::is_integral_v<E> {
        vec r;
        for (int i = 0; i < N; i++) r.lane[i] = static_cast<E>(static_cast<detail::uint_of<E>>(a.lane[i]) << (b.lane[i] & (sizeof(E) * 8 - 1)));
        return r;
    }
    friend constexpr vec operator>>(const vec& a, const vec& b) requires std::is_integral_v<E> {
        vec r;
        for (int i = 0; i < N; i++) r.lane[i] = static_cast<E>(a.lane[i] >> (b.lane[i] & (sizeof(E) * 8 - 1)));
        return r;
    }
    friend constexpr vec& operator<<=(vec& a, const vec& b) requires std::is_integral_v<E> { return a = a << b; }
    friend constexpr vec& operator>>=(vec& a, const vec& b) requires std::is_integral_v<E> { return a = a >> b; }
    friend constexpr vec operator-(const vec& a) {
        vec r;
        for (int i = 0; i < N; i++) r.lane[i] = static_cast<E>(-a.lane[i]);
        return r;
    }
    friend constexpr vec operator+(const vec& a) { return a; }
    friend constexpr vec operator~(const vec& a) requires std::is_integral_v<E> {
        vec r;
        for (int i = 0; i < N; i++) r.lane[i] = static_cast<E>(~a.lane[i]);
        return r;
    }
#define SIMD_GXX_CMP(sym)                                                                                       \
    friend constexpr mask_t operator sym(const vec& a, const vec& b) {                                            \
        mask_t r;                                                                                                 \
        for (int i = 0; i < N; i++) r.lane[i] = (a.lane[i] sym b.lane[i]) ? -1 : 0;                               \
        return r;                                                                                                 \
    }
    SIMD_GXX_CMP(<) SIMD_GXX_CMP(>) SIMD_GXX_CMP(<=) SIMD_GXX_CMP(>=) SIMD_GXX_CMP(==) SIMD_GXX_CMP(!=)
#undef SIMD_GXX_CMP
    friend constexpr mask_t operator&&(const vec& a, const vec& b) {
        mask_t r;
        for (int i = 0; i < N; i++) r.lane[i] = (a.lane[i] != 0 && b.lane[i] != 0) ? -1 : 0;
        return r;
    }
    friend constexpr mask_t operator||(const vec& a, const vec& b) {
        mask_t r;
        for (int i = 0; i < N; i++) r.lane[i] = (a.lane[i] != 0 || b.lane[i] != 0) ? -1 : 0;
        return r;
    }
    friend constexpr mask_t operator!(const vec& a) {
        mask_t r;
        for (int i = 0; i < N; i++) r.lane[i] = a.lane[i] == 0 ? -1 : 0;
        return r;
    }
};

template<typename T> concept is_vector = requires { requires std::remove_cvref_t<T>::is_simd_vec; };
template<typename T> using elem_type = typename std::remove_cvref_t<T>::elem;
template<is_vector T> constexpr int width_of = std::remove_cvref_t<T>::width;
template<is_vector T> using mask_vec = typename std::remove_cvref_t<T>::mask_t;
template<is_vector T>
using bitmask = std::conditional_t<width_of<T> <= 8, uint8_t, std::conditional_t<width_of<T> <= 16, uint16_t,
                std::conditional_t<width_of<T> <= 32, uint32_t, uint64_t>>>;

template<typename S> concept scalar = std::is_arithmetic_v<S>;

// ---- small fixed-size tuples of vectors (x, y[, z[, w]]) ------------------------------------------------------------------
template<typename T, int N> struct mdvec;

#define SIMD_GXX_MD_OPS(MD, EXPR_VV, EXPR_VS, EXPR_SV)                                                          \
    SIMD_GXX_MD_OP(+, MD, EXPR_VV, EXPR_VS, EXPR_SV) SIMD_GXX_MD_OP(-, MD, EXPR_VV, EXPR_VS, EXPR_SV)             \
    SIMD_GXX_MD_OP(*, MD, EXPR_VV, EXPR_VS, EXPR_SV) SIMD_GXX_MD_OP(/, MD, EXPR_VV, EXPR_VS, EXPR_SV)             \
    SIMD_GXX_MD_OP(|, MD, EXPR_VV, EXPR_VS, EXPR_SV) SIMD_GXX_MD_OP(&, MD, EXPR_VV, EXPR_VS, EXPR_SV)             \
    SIMD_GXX_MD_OP(^, MD, EXPR_VV, EXPR_VS, EXPR_SV) SIMD_GXX_MD_OP(<<, MD, EXPR_VV, EXPR_VS, EXPR_SV)            \
    SIMD_GXX_MD_OP(>>, MD, EXPR_VV, EXPR_VS, EXPR_SV)

template<typename T>
struct mdvec<T, 2> {
    T x, y;
    constexpr mdvec() = default;
    constexpr mdvec(T v) : x(v), y(v) {}
    template<scalar S> constexpr mdvec(S s) : x(s), y(s) {}
    constexpr mdvec(T x_, T y_) : x(x_), y(y_) {}
    constexpr mdvec(const glm::vec<2, elem_type<T>>& g) : x(g.x), y(g.y) {}
    explicit constexpr mdvec(const mdvec<T, 3>& o) : x(o.x), y(o.y) {}
    explicit constexpr mdvec(const mdvec<T, 4>& o) : x(o.x), y(o.y) {}
    constexpr T& operator[](size_t i) { assert(i < 2); return i == 0 ? x : y; }
    constexpr const T& operator[](size_t i) const { return i == 0 ? x : y; }
    friend constexpr mdvec operator-(const mdvec& a) { return { -a.x, -a.y }; }
#define SIMD_GXX_MD_OP(sym, MD, VV, VS, SV)                                                                     \
    friend constexpr mdvec operator sym(const mdvec& a, const mdvec& b) { return { a.x sym b.x, a.y sym b.y }; }  \
    template<scalar S> friend constexpr mdvec operator sym(const mdvec& a, S b) { return { a.x sym T(b), a.y sym T(b) }; } \
    template<scalar S> friend constexpr mdvec operator sym(S a, const mdvec& b) { return { T(a) sym b.x, T(a) sym b.y }; } \
    friend constexpr mdvec& operator sym##=(mdvec& a, const mdvec& b) { return a = a sym b; }                     \
    template<scalar S> friend constexpr mdvec& operator sym##=(mdvec& a, S b) { return a = a sym b; }
    SIMD_GXX_MD_OPS(mdvec, 0, 0, 0)
#undef SIMD_GXX_MD_OP
};

template<typename T>
struct mdvec<T, 3> {
    T x, y, z;
    constexpr mdvec() = default;
    constexpr mdvec(T v) : x(v), y(v), z(v) {}
    template<scalar S> constexpr mdvec(S s) : x(s), y(s), z(s) {}
    constexpr mdvec(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
    constexpr mdvec(const mdvec<T, 2>& xy, T z_) : x(xy.x), y(xy.y), z(z_) {}
    constexpr mdvec(const glm::vec<3, elem_type<T>>& g) : x(g.x), y(g.y), z(g.z) {}
    explicit constexpr mdvec(const mdvec<T, 4>& o) : x(o.x), y(o.y), z(o.z) {}
    constexpr T& operator[](size_t i) { assert(i < 3); return i == 0 ? x : (i == 1 ? y : z); }
    constexpr const T& operator[](size_t i) const { return i == 0 ? x : (i == 1 ? y : z); }
    friend constexpr mdvec operator-(const mdvec& a) { return { -a.x, -a.y, -a.z }; }
#define SIMD_GXX_MD_OP(sym, MD, VV, VS, SV)                                                                     \
    friend constexpr mdvec operator sym(const mdvec& a, const mdvec& b) { return { a.x sym b.x, a.y sym b.y, a.z sym b.z }; } \
    template<scalar S> friend constexpr mdvec operator sym(const mdvec& a, S b) { return { a.x sym T(b), a.y sym T(b), a.z sym T(b) }; } \
    template<scalar S> friend constexpr mdvec operator sym(S a, const mdvec& b) { return { T(a) sym b.x, T(a) sym b.y, T(a) sym b.z }; } \
    friend constexpr mdvec& operator sym##=(mdvec& a, const mdvec& b) { return a = a sym b; }                     \
    template<scalar S> friend constexpr mdvec& operator sym##=(mdvec& a, S b) { return a = a sym b; }
    SIMD_GXX_MD_OPS(mdvec, 0, 0, 0)
#undef SIMD_GXX_MD_OP
};

template<typename T>
struct mdvec<T, 4> {
    T x, y, z, w;
    constexpr mdvec() = default;
    constexpr mdvec(T v) : x(v), y(v), z(v), w(v) {}
    template<scalar S> constexpr mdvec(S s) : x(s), y(s), z(s), w(s) {}
    constexpr mdvec(T x_, T y_, T z_, T w_) : x(x_), y(y_), z(z_), w(w_) {}
    constexpr mdvec(const mdvec<T, 2>& xy, T z_, T w_) : x(xy.x), y(xy.y), z(z_), w(w_) {}
    constexpr mdvec(const mdvec<T, 3>& xyz, T w_) : x(xyz.x), y(xyz.y), z(xyz.z), w(w_) {}
    constexpr mdvec(const glm::vec<4, elem_type<T>>& g) : x(g.x), y(g.y), z(g.z), w(g.w) {}
    constexpr T& operator[](size_t i) { assert(i < 4); return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    constexpr const T& operator[](size_t i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    friend constexpr mdvec operator-(const mdvec& a) { return { -a.x, -a.y, -a.z, -a.w }; }
#define SIMD_GXX_MD_OP(sym, MD, VV, VS, SV)                                                                     \
    friend constexpr mdvec operator sym(const mdvec& a, const mdvec& b) { return { a.x sym b.x, a.y sym b.y, a.z sym b.z, a.w sym b.w }; } \
    template<scalar S> friend constexpr mdvec operator sym(const mdvec& a, S b) { return { a.x sym T(b), a.y sym T(b), a.z sym T(b), a.w sym T(b) }; } \
    template<scalar S> friend constexpr mdvec operator sym(S a, const mdvec& b) { return { T(a) sym b.x, T(a) sym b.y, T(a) sym b.z, T(a) sym b.w }; } \
    friend constexpr mdvec& operator sym##=(mdvec& a, const mdvec& b) { return a = a sym b; }                     \
    template<scalar S> friend constexpr mdvec& operator sym##=(mdvec& a, S b) { return a = a sym b; }
    SIMD_GXX_MD_OPS(mdvec, 0, 0, 0)
#undef SIMD_GXX_MD_OP
};
#undef SIMD_GXX_MD_OPS

}  // namespace simd

using v_int = simd::vec<int32_t>;
using v_uint = simd::vec<uint32_t>;
using v_float = simd::vec<float>;
using v_mask = simd::bitmask<v_int>;

using v_float2 = simd::mdvec<v_float, 2>;
using v_float3 = simd::mdvec<v_float, 3>;
using v_float4 = simd::mdvec<v_float, 4>;
using v_int2 = simd::mdvec<v_int, 2>;
using v_int3 = simd::mdvec<v_int, 3>;
using v_int4 = simd::mdvec<v_int, 4>;
using v_uint2 = simd::mdvec<v_uint, 2>;
using v_uint3 = simd::mdvec<v_uint, 3>;
using v_uint4 = simd::mdvec<v_uint, 4>;

using float2 = glm::vec2;
using float3 = glm::vec3;
using float4 = glm::vec4;
using float4x4 = glm::mat4x4;
using int2 = glm::ivec2;
using int3 = glm::ivec3;
using int4 = glm::ivec4;
using uint2 = glm::uvec2;
using uint3 = glm::uvec3;
using uint4 = glm::uvec4;

namespace simd {

template<is_vector T>
constexpr T lane_idx = [] { T r; for (int i = 0; i < width_of<T>; i++) r.lane[i] = static_cast<elem_type<T>>(i); return r; }();

// ---- memory ---------------------------------------------------------------------------------------------------------------
template<typename T>
SIMD_INLINE T load(const void* ptr) { T v; std::memcpy(&v, ptr, sizeof(T)); return v; }
template<typename E, int N = 64 / sizeof(E)>
SIMD_INLINE vec<E, N> load(const E* ptr) { return load<vec<E, N>>(static_cast<const void*>(ptr)); }
template<typename T>
SIMD_INLINE void store(void* ptr, const T& value) { std::memcpy(ptr, &value, sizeof(T)); }

template<is_vector T>
SIMD_INLINE T gather(const void* ptr, mask_vec<T> idx, mask_vec<T> mask = -1) {
    T r;
    for (int i = 0; i < width_of<T>; i++) if (mask.lane[i] < 0) r.lane[i] = static_cast<const elem_type<T>*>(ptr)[idx.lane[i]];
    return r;
}
template<typename E, int N = 64 / sizeof(E)>
SIMD_INLINE vec<E, N> gather(const E* ptr, mask_vec<vec<E, N>> idx, mask_vec<vec<E, N>> mask = -1) {
    return gather<vec<E, N>>(static_cast<const void*>(ptr), idx, mask);
}
template<typename T> requires(sizeof(T) == 4)
SIMD_INLINE mdvec<vec<T>, 2> gather2(const T* ptr, v_int idx, v_int mask = -1) {
    mdvec<vec<T>, 2> r;
    for (int i = 0; i < vec_width; i++) if (mask.lane[i] < 0) { r.x.lane[i] = ptr[idx.lane[i]]; r.y.lane[i] = ptr[idx.lane[i] + 1]; }
    return r;
}
template<typename T> requires(sizeof(T) == 4)
SIMD_INLINE mdvec<vec<T>, 4> gather4(const T* ptr, v_int idx, v_int mask = -1) {
    auto a = gather2(ptr + 0, idx, mask);
    auto b = gather2(ptr + 2, idx, mask);
    return { a.x, a.y, b.x, b.y };
}
// 64-entry table lookups: the reference builds them from two vpermt2d and a select on idx < 32; kept as the instructions
SIMD_INLINE v_uint gather_preload64(const uint32_t data[64], v_uint idx) {
    __m512i v0 = _mm512_loadu_si512(&data[0]), v1 = _mm512_loadu_si512(&data[16]);
    __m512i v2 = _mm512_loadu_si512(&data[32]), v3 = _mm512_loadu_si512(&data[48]);
    __m512i i = idx;
    __m512i p01 = _mm512_permutex2var_epi32(v0, i, v1), p23 = _mm512_permutex2var_epi32(v2, i, v3);
    return _mm512_mask_blend_epi32(_mm512_cmplt_epu32_mask(i, _mm512_set1_epi32(32)), p23, p01);
}
SIMD_INLINE v_float gather_preload64(const float data[64], v_uint idx) {
    return std::bit_cast<v_float>(gather_preload64(reinterpret_cast<const uint32_t*>(data), idx));
}
SIMD_INLINE v_uint gather_preload128(const uint8_t data[128], v_uint idx) {
    return _mm512_permutex2var_epi8(_mm512_loadu_si512(&data[0]), idx, _mm512_loadu_si512(&data[64]));
}

template<auto indices, is_vector T>
SIMD_INLINE constexpr T shuffle(T values) {
    T r;
    for (int i = 0; i < width_of<T>; i++) r.lane[i] = values.lane[indices.lane[i]];
    return r;
}

// ---- masks ----------------------------------------------------------------------------------------------------------------
template<is_vector T> requires std::is_signed_v<elem_type<T>>
SIMD_INLINE constexpr bitmask<T> movemask(T mask) {
    bitmask<T> r = 0;
    for (int i = 0; i < width_of<T>; i++) if (mask.lane[i] < 0) r |= static_cast<bitmask<T>>(bitmask<T>(1) << i);
    return r;
}
template<is_vector T> requires std::is_signed_v<elem_type<T>>
SIMD_INLINE constexpr bool any(T mask) { return movemask(mask) != 0; }
template<is_vector T> requires std::is_signed_v<elem_type<T>>
SIMD_INLINE constexpr bool all(T mask) {
    for (int i = 0; i < width_of<T>; i++) if (!(mask.lane[i] < 0)) return false;
    return true;
}

// `cond ? a : b` on vectors (a Clang extension) — the patched reference sources call this instead
template<is_vector T>
SIMD_INLINE constexpr T select(mask_vec<T> cond, T ifTrue, T ifFalse) {
    T r;
    for (int i = 0; i < width_of<T>; i++) r.lane[i] = cond.lane[i] != 0 ? ifTrue.lane[i] : ifFalse.lane[i];
    return r;
}
template<is_vector T, scalar S> SIMD_INLINE constexpr T select(mask_vec<T> cond, T a, S b) { return select<T>(cond, a, T(b)); }
template<is_vector T, scalar S> SIMD_INLINE constexpr T select(mask_vec<T> cond, S a, T b) { return select<T>(cond, T(a), b); }
template<typename T, int N>
SIMD_INLINE constexpr mdvec<T, N> select(mask_vec<T> cond, mdvec<T, N> ifTrue, mdvec<T, N> ifFalse) {
    mdvec<T, N> r;
    for (int i = 0; i < N; i++) r[i] = select<T>(cond, ifTrue[i], ifFalse[i]);
    return r;
}
template<typename T, int N>
SIMD_INLINE constexpr mdvec<T, N> select(mdvec<mask_vec<T>, N> cond, mdvec<T, N> ifTrue, mdvec<T, N> ifFalse) {
    mdvec<T, N> r;
    for (int i = 0; i < N; i++) r[i] = select<T>(cond[i], ifTrue[i], ifFalse[i]);
    return r;
}
template<is_vector T>
SIMD_INLINE void cmov(T& dest, T ifTrue, mask_vec<T> cond) { dest = select<T>(cond, ifTrue, dest); }
template<typename T, int N>
SIMD_INLINE void cmov(mdvec<T, N>& dest, mdvec<T, N> ifTrue, mask_vec<T> cond) { dest = select(cond, ifTrue, dest); }

// ---- fundamentals -----------------------------------------------------------------------------------------------------------
#define SIMD_GXX_MAP1(name, expr)                                                                               \
    template<is_vector T> SIMD_INLINE T name(T x) {                                                               \
        T r;                                                                                                      \
        for (int i = 0; i < width_of<T>; i++) { auto v = x.lane[i]; r.lane[i] = static_cast<elem_type<T>>(expr); } \
        return r;                                                                                                 \
    }
template<is_vector T> requires std::is_floating_point_v<elem_type<T>>
SIMD_INLINE T fma(T a, T b, T c) {
    T r;
    for (int i = 0; i < width_of<T>; i++) r.lane[i] = std::fma(a.lane[i], b.lane[i], c.lane[i]);
    return r;
}
template<is_vector T> requires std::is_floating_point_v<elem_type<T>> SIMD_INLINE T fma(T a, elem_type<T> b, T c) { return fma(a, T(b), c); }
template<is_vector T> requires std::is_floating_point_v<elem_type<T>> SIMD_INLINE T fma(T a, T b, elem_type<T> c) { return fma(a, b, T(c)); }
template<is_vector T> requires std::is_floating_point_v<elem_type<T>> SIMD_INLINE T fma(T a, elem_type<T> b, elem_type<T> c) { return fma(a, T(b), T(c)); }

SIMD_GXX_MAP1(sqrt, std::sqrt(v))
SIMD_GXX_MAP1(floor, std::floor(v))
SIMD_GXX_MAP1(ceil, std::ceil(v))
SIMD_GXX_MAP1(round, std::round(v))
#ifdef SWR_COMPAT_RCP14
SIMD_INLINE v_float approx_rsqrt(v_float x) { return _mm512_rsqrt14_ps(x); }
SIMD_INLINE v_float approx_rcp(v_float x) { return _mm512_rcp14_ps(x); }
#else
SIMD_INLINE v_float approx_rsqrt(v_float x) { v_float r; for (int i = 0; i < 16; i++) r.lane[i] = 1.0f / std::sqrt(x.lane[i]); return r; }
SIMD_INLINE v_float approx_rcp(v_float x) { v_float r; for (int i = 0; i < 16; i++) r.lane[i] = 1.0f / x.lane[i]; return r; }
#endif
SIMD_INLINE v_float approx_sqrt(v_float x) { return approx_rsqrt(x) * x; }
SIMD_INLINE v_float fract(v_float x) { return _mm512_reduce_ps(x, _MM_FROUND_TO_NEG_INF); }
SIMD_INLINE v_int floor2i(v_float x) { return _mm512_cvt_roundps_epi32(x, _MM_FROUND_TO_NEG_INF | _MM_FROUND_NO_EXC); }
SIMD_INLINE v_int round2i(v_float x) { return _mm512_cvtps_epi32(x); }

template<is_vector T> SIMD_INLINE T abs(T x) {
    T r;
    for (int i = 0; i < width_of<T>; i++) {
        if constexpr (std::is_floating_point_v<elem_type<T>>) r.lane[i] = std::fabs(x.lane[i]);
        else r.lane[i] = x.lane[i] < 0 ? static_cast<elem_type<T>>(0 - x.lane[i]) : x.lane[i];
    }
    return r;
}
template<is_vector T> SIMD_INLINE T min(T x, T y) {
    T r;
    for (int i = 0; i < width_of<T>; i++) {
        if constexpr (std::is_floating_point_v<elem_type<T>>) r.lane[i] = std::fmin(x.lane[i], y.lane[i]);
        else r.lane[i] = y.lane[i] < x.lane[i] ? y.lane[i] : x.lane[i];
    }
    return r;
}
template<is_vector T> SIMD_INLINE T max(T x, T y) {
    T r;
    for (int i = 0; i < width_of<T>; i++) {
        if constexpr (std::is_floating_point_v<elem_type<T>>) r.lane[i] = std::fmax(x.lane[i], y.lane[i]);
        else r.lane[i] = y.lane[i] > x.lane[i] ? y.lane[i] : x.lane[i];
    }
    return r;
}
template<is_vector T, scalar S> SIMD_INLINE T min(T x, S y) { return min(x, T(y)); }
template<is_vector T, scalar S> SIMD_INLINE T min(S x, T y) { return min(T(x), y); }
template<is_vector T, scalar S> SIMD_INLINE T max(T x, S y) { return max(x, T(y)); }
template<is_vector T, scalar S> SIMD_INLINE T max(S x, T y) { return max(T(x), y); }
template<is_vector T> SIMD_INLINE T clamp(T x, T a, T b) { return min(max(x, a), b); }
template<is_vector T, scalar A, scalar B> SIMD_INLINE T clamp(T x, A a, B b) { return min(max(x, T(a)), T(b)); }

SIMD_INLINE v_float min_abs(v_float x, v_float y) { return _mm512_range_ps(x, y, 0b1010); }
SIMD_INLINE v_float max_abs(v_float x, v_float y) { return _mm512_range_ps(x, y, 0b1011); }

template<is_vector T> requires std::is_integral_v<elem_type<T>> SIMD_INLINE T add_sat(T x, T y) {
    T r;
    for (int i = 0; i < width_of<T>; i++) {
        elem_type<T> o;
        if (__builtin_add_overflow(x.lane[i], y.lane[i], &o)) o = (std::is_signed_v<elem_type<T>> && y.lane[i] < 0) ? std::numeric_limits<elem_type<T>>::min() : std::numeric_limits<elem_type<T>>::max();
        r.lane[i] = o;
    }
    return r;
}
template<is_vector T> requires std::is_integral_v<elem_type<T>> SIMD_INLINE T sub_sat(T x, T y) {
    T r;
    for (int i = 0; i < width_of<T>; i++) {
        elem_type<T> o;
        if (__builtin_sub_overflow(x.lane[i], y.lane[i], &o)) o = (std::is_signed_v<elem_type<T>> && y.lane[i] < 0) ? std::numeric_limits<elem_type<T>>::max() : std::numeric_limits<elem_type<T>>::min();
        r.lane[i] = o;
    }
    return r;
}

SIMD_INLINE v_float mulsign(v_float x, v_float y) {      // y < 0 ? -x : x
    v_float r;
    for (int i = 0; i < 16; i++) r.lane[i] = std::bit_cast<float>(std::bit_cast<uint32_t>(x.lane[i]) ^ (std::bit_cast<uint32_t>(y.lane[i]) & 0x80000000u));
    return r;
}
SIMD_INLINE v_float copysign(v_float x, v_float s) {
    v_float r;
    for (int i = 0; i < 16; i++) r.lane[i] = std::bit_cast<float>((std::bit_cast<uint32_t>(x.lane[i]) & 0x7FFFFFFFu) | (std::bit_cast<uint32_t>(s.lane[i]) & 0x80000000u));
    return r;
}

template<typename R, is_vector T>
SIMD_INLINE constexpr vec<R, width_of<T>> conv(T x) {
    vec<R, width_of<T>> r;
    for (int i = 0; i < width_of<T>; i++) r.lane[i] = static_cast<R>(x.lane[i]);
    return r;
}
template<typename R, is_vector T, int N>
SIMD_INLINE constexpr mdvec<vec<R, width_of<T>>, N> conv(mdvec<T, N> x) {
    mdvec<vec<R, width_of<T>>, N> r;
    for (int i = 0; i < N; i++) r[i] = conv<R>(x[i]);
    return r;
}
template<std::integral R, is_vector T> requires(sizeof(R) < sizeof(elem_type<T>))
SIMD_INLINE constexpr vec<R, width_of<T>> conv_sat(T x) {
    vec<R, width_of<T>> r;
    for (int i = 0; i < width_of<T>; i++) {
        auto v = x.lane[i];
        auto lo = static_cast<elem_type<T>>(std::is_unsigned_v<elem_type<T>> ? 0 : std::numeric_limits<R>::min());
        auto hi = static_cast<elem_type<T>>(std::numeric_limits<R>::max());
        r.lane[i] = static_cast<R>(v < lo ? lo : (v > hi ? hi : v));
    }
    return r;
}

template<typename Dst, typename Src>
SIMD_INLINE constexpr Dst as(Src x) { return std::bit_cast<Dst>(x); }
template<typename Dst, typename Src, int N> requires std::is_arithmetic_v<Dst>
SIMD_INLINE vec<Dst, int(sizeof(Src) * N / sizeof(Dst))> as(vec<Src, N> x) { return std::bit_cast<vec<Dst, int(sizeof(Src) * N / sizeof(Dst))>>(x); }

// ---- math -----------------------------------------------------------------------------------------------------------------
constexpr float pi = 3.141592653589793f;
constexpr float tau = 6.283185307179586f;
constexpr float inv_pi = 0.3183098861837907f;

SIMD_INLINE v_float2 sincos_2pi(v_float x) {               // the reference's polynomial (sin 9e-7, cos 7e-6 max abs error)
    x = _mm512_reduce_ps(x + 0.25f, _MM_FROUND_TO_NEAREST_INT);
    v_float x1 = abs(x) - 0.25f, x2 = x1 * x1;
    v_float s = fma(x2, fma(x2, fma(x2, -70.993433272f, 81.340768887f), -41.337142371f), 6.283164044f) * x1;
    v_float c = fma(x2, fma(x2, fma(x2, -78.216131988f, 64.660541218f), -19.735752060f), 0.999993295f);
    return v_float2(s, mulsign(c, x));
}
SIMD_INLINE v_float2 sincos(v_float a) { return sincos_2pi(a * 0.15915494309189535f); }
SIMD_INLINE v_float sin(v_float a) { return sincos(a).x; }
SIMD_INLINE v_float cos(v_float a) { return sincos(a).y; }
SIMD_INLINE v_float approx_log2(v_float x) { return fma(conv<float>(as<uint32_t>(x)), 1.1920928955078125e-7f, -126.94269504f); }
SIMD_INLINE v_float approx_exp2(v_float x) { x = max(x, -126.0f); return as<float>(round2i((1 << 23) * (x + 126.94269504f))); }
SIMD_INLINE v_float approx_pow(v_float x, v_float y) { return approx_exp2(approx_log2(x) * y); }
SIMD_INLINE v_int ilog2(v_float x) { return (as<v_int>(x) - (127 << 23)) >> 23; }

SIMD_INLINE v_float dot(v_float2 a, v_float2 b) { return fma(a.x, b.x, a.y * b.y); }
SIMD_INLINE v_float dot(v_float3 a, v_float3 b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }
SIMD_INLINE v_float3 cross(v_float3 a, v_float3 b) {
    return { fma(a.y, b.z, -a.z * b.y), fma(a.z, b.x, -a.x * b.z), fma(a.x, b.y, -a.y * b.x) };
}
SIMD_INLINE v_float3 normalize(v_float3 a) { return a * approx_rsqrt(dot(a, a)); }
SIMD_INLINE v_float length(v_float3 p) { return approx_sqrt(dot(p, p)); }
SIMD_INLINE v_float3 reflect(v_float3 i, v_float3 n) { return i - 2.0f * dot(n, i) * n; }
template<is_vector T> requires std::is_floating_point_v<elem_type<T>>
SIMD_INLINE T lerp(T a, T b, T t) { return fma(t, b, fma(-t, a, a)); }
SIMD_INLINE v_uint lerp16(v_uint a, v_uint b, v_uint t) { return _mm512_add_epi16(a, _mm512_mulhrs_epi16(_mm512_sub_epi16(b, a), t)); }
SIMD_INLINE v_float smoothstep(v_float a, v_float b, v_float t) {
    t = clamp((t - a) / (b - a), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
SIMD_INLINE v_float4 mul(const glm::mat4& m, const v_float4& v) {
    return { fma(v.x, m[0][0], fma(v.y, m[1][0], fma(v.z, m[2][0], v.w * m[3][0]))),
             fma(v.x, m[0][1], fma(v.y, m[1][1], fma(v.z, m[2][1], v.w * m[3][1]))),
             fma(v.x, m[0][2], fma(v.y, m[1][2], fma(v.z, m[2][2], v.w * m[3][2]))),
             fma(v.x, m[0][3], fma(v.y, m[1][3], fma(v.z, m[2][3], v.w * m[3][3]))) };
}
SIMD_INLINE v_float3 mul(const glm::mat3& m, const v_float3& n) {
    return { fma(n.x, m[0][0], fma(n.y, m[1][0], n.z * m[2][0])),
             fma(n.x, m[0][1], fma(n.y, m[1][1], n.z * m[2][1])),
             fma(n.x, m[0][2], fma(n.y, m[1][2], n.z * m[2][2])) };
}
SIMD_INLINE v_float4 perspective_div(const v_float4& v) {
    v_float rw = 1.0f / v.w;
    return { v.x * rw, v.y * rw, v.z * rw, rw };
}

// ---- bitwise ----------------------------------------------------------------------------------------------------------------
template<is_vector T> requires std::is_integral_v<elem_type<T>> SIMD_INLINE T popcnt(T x) {
    T r; for (int i = 0; i < width_of<T>; i++) r.lane[i] = static_cast<elem_type<T>>(std::popcount(static_cast<detail::uint_of<elem_type<T>>>(x.lane[i]))); return r;
}
template<is_vector T> requires std::is_integral_v<elem_type<T>> SIMD_INLINE T lzcnt(T x) {
    T r; for (int i = 0; i < width_of<T>; i++) r.lane[i] = static_cast<elem_type<T>>(std::countl_zero(static_cast<detail::uint_of<elem_type<T>>>(x.lane[i]))); return r;
}
template<is_vector T> requires std::is_integral_v<elem_type<T>> SIMD_INLINE T tzcnt(T x) {
    T r; for (int i = 0; i < width_of<T>; i++) r.lane[i] = static_cast<elem_type<T>>(std::countr_zero(static_cast<detail::uint_of<elem_type<T>>>(x.lane[i]))); return r;
}
template<std::integral T> SIMD_INLINE constexpr uint32_t popcnt(T value) { return (uint32_t)std::popcount(static_cast<std::make_unsigned_t<T>>(value)); }
template<std::integral T> SIMD_INLINE constexpr uint32_t tzcnt(T value) { return (uint32_t)std::countr_zero(static_cast<std::make_unsigned_t<T>>(value)); }
template<std::integral T> SIMD_INLINE constexpr uint32_t lzcnt(T value) { return (uint32_t)std::countl_zero(static_cast<std::make_unsigned_t<T>>(value)); }
SIMD_INLINE v_uint rotl(v_uint a, int b) { return (a << b) | (a >> (32 - b)); }
SIMD_INLINE v_uint rotr(v_uint a, int b) { return (a >> b) | (a << (32 - b)); }

// ---- misc -----------------------------------------------------------------------------------------------------------------
struct AlignedDeleter {
    void operator()(void* data) const { _mm_free(data); }
};
template<typename T> using AlignedBuffer = std::unique_ptr<T[], AlignedDeleter>;
template<typename T>
AlignedBuffer<T> alloc_buffer(size_t count, size_t align = 64) { return AlignedBuffer<T>((T*)_mm_malloc(count * sizeof(T), align)); }

class BitIter {           // for (uint32_t i : BitIter(mask)) visits the set bits in ascending order
    uint64_t mask;
public:
    BitIter(uint64_t m) : mask(m) {}
    BitIter& operator++() { mask &= mask - 1; return *this; }
    uint32_t operator*() const { return (uint32_t)std::countr_zero(mask); }
    friend bool operator!=(const BitIter& a, const BitIter& b) { return a.mask != b.mask; }
    BitIter begin() const { return *this; }
    BitIter end() const { return BitIter(0); }
};

}  // namespace simd
