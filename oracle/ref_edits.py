"""The edit table of oracle/ref_build.py: what has to change in the reference's sources before g++ accepts them.

TEST INFRASTRUCTURE. Each entry names a file, a 1-based line of the reference snapshot, the tokens on that line that
are replaced and their replacement. The recipe refuses to build when a line no longer holds the tokens (i.e. when
/root/reference is not the snapshot this table was written against). No entry changes an arithmetic operation, an
operand order, a constant or control flow; the kinds are

  T  `mask ? a : b` with a VECTOR condition (Clang's vector ternary) -> simd::select(mask, a, b): the same lane-wise pick
  V  a Clang-only vector type spelling -> the stand-in's spelling of the same lanes
  A  inline asm with a vector-register constraint on a class type -> the intrinsic the source itself names on the next line
  X  `#if 0` around code outside the raster path that needs libraries / stand-in functions this image lacks

Everything else the sources need from Clang (implicit scalar->vector conversion, vector compares yielding lane masks,
`&&` / `||` / `!` on vectors, C-style bit casts between vectors, intrinsics taking void*) is provided by
oracle/compat/simd_gxx.h without touching the sources.
"""

R = "Rasterizer.cpp"
S = "Shading.cpp"
T = "Texture.h"

EDITS = [
    # ---- Rasterizer.cpp ------------------------------------------------------------------------------------------------
    # TrianglePacket::Setup: det = flip ? -det : det                                                               (T)
    (R, 266, "flip ? -det : det", "simd::select(flip, -det, det)"),
    # ComputeEdge: top-left bias                                                                                   (T)
    (R, 293, "(a > 0 || (a == 0 && b > 0)) ? 0 : -1", "simd::select(a > 0 || (a == 0 && b > 0), v_int(0), v_int(-1))"),
    # TriangleEdgeVars::Setup: sign flip of the edge deltas for clockwise triangles                                (T)
    (R, 308, "A01 = flip ? -A01 : A01, B01 = flip ? -B01 : B01", "A01 = simd::select(flip, -A01, A01), B01 = simd::select(flip, -B01, B01)"),
    (R, 309, "A12 = flip ? -A12 : A12, B12 = flip ? -B12 : B12", "A12 = simd::select(flip, -A12, A12), B12 = simd::select(flip, -B12, B12)"),
    (R, 310, "A20 = flip ? -A20 : A20, B20 = flip ? -B20 : B20", "A20 = simd::select(flip, -A20, A20), B20 = simd::select(flip, -B20, B20)"),
    (R, 311, "flip ? -det : det", "simd::select(flip, -det, det)"),
    # Clipper::ComputeClipCodes: 16 x u8 / 16 x bool vectors                                                       (V)
    (R, 354, "uint8_t [[clang::ext_vector_type(16)]]", "simd::vec<uint8_t, 16>"),
    (R, 355, "bool [[clang::ext_vector_type(16)]]", "simd::vec<int8_t, 16>"),
    #   outcode |= v_bool(cmp) ? v_byte(bit) : 0                                                                   (T)
    (R, 376, "v_bool(vi.x < -vi.w) ? v_byte(1 << (int)ClipPlane::Left) : 0", "simd::select(v_bool(vi.x < -vi.w), v_byte(1 << (int)ClipPlane::Left), v_byte(0))"),
    (R, 377, "v_bool(vi.x > +vi.w) ? v_byte(1 << (int)ClipPlane::Right) : 0", "simd::select(v_bool(vi.x > +vi.w), v_byte(1 << (int)ClipPlane::Right), v_byte(0))"),
    (R, 378, "v_bool(vi.y < -vi.w) ? v_byte(1 << (int)ClipPlane::Top) : 0", "simd::select(v_bool(vi.y < -vi.w), v_byte(1 << (int)ClipPlane::Top), v_byte(0))"),
    (R, 379, "v_bool(vi.y > +vi.w) ? v_byte(1 << (int)ClipPlane::Bottom) : 0", "simd::select(v_bool(vi.y > +vi.w), v_byte(1 << (int)ClipPlane::Bottom), v_byte(0))"),
    (R, 380, "v_bool(vi.z < -vi.w) ? v_byte(1 << (int)ClipPlane::Near) : 0", "simd::select(v_bool(vi.z < -vi.w), v_byte(1 << (int)ClipPlane::Near), v_byte(0))"),
    (R, 381, "v_bool(vi.z > +vi.w) ? v_byte(1 << (int)ClipPlane::Far) : 0", "simd::select(v_bool(vi.z > +vi.w), v_byte(1 << (int)ClipPlane::Far), v_byte(0))"),

    # ---- Texture.h -----------------------------------------------------------------------------------------------------
    # pixfmt::RG16f::Unpack: `asm("vpermb ...")` works around a Clang bug; the comment below it gives the intrinsic      (A)
    (T, 119, 'asm("vpermb %2, %1, %0" : "=v"(split) : "v"(shuf), "v"(packed));', "split = _mm512_permutexvar_epi8(shuf, packed);"),
    # mirrored-repeat wrap                                                                                         (T)
    (T, 426, "((ix & (MaskLerpU + 1)) != 0 ? MaskLerpU : 0)", "simd::select((ix & (MaskLerpU + 1)) != 0, v_int(MaskLerpU), v_int(0))"),
    (T, 427, "((iy & (MaskLerpV + 1)) != 0 ? MaskLerpV : 0)", "simd::select((iy & (MaskLerpV + 1)) != 0, v_int(MaskLerpV), v_int(0))"),
    # bilinear footprint: second row / column only when in bounds                                                  (T)
    (T, 520, "(inboundY ? (1 << stride) : 0)", "simd::select(inboundY, v_int(1) << stride, v_int(0))"),
    (T, 529, "(inboundY ? 1 : 0)", "simd::select(inboundY, v_int(1), v_int(0))"),
    (T, 545, "inboundX ? fx : 0", "simd::select(inboundX, fx, v_uint(0))"),
    (T, 563, "inboundX ? fx : 0", "simd::select(inboundX, fx, v_float(0))"),

    # ---- Shading.cpp ---------------------------------------------------------------------------------------------------
    # FS_Overdraw: +1 pixel for covered lanes, +1 helper for the rest                                              (T)
    (S, 337, "isActive ? v_uint(0x0001'0000) : v_uint(0x0000'0001)", "simd::select(isActive, v_uint(0x0001'0000), v_uint(0x0000'0001))"),
    # Resolve: sky lanes take depth 1                                                                              (T)
    (S, 666, "skyMask ? 1.0f : tileDepth", "simd::select(skyMask, v_float(1.0f), tileDepth)"),
    # Resolve: light markers behind geometry keep the background                                                   (T)
    (S, 728, "depthMask ? finalColor : bgColor", "simd::select(depthMask, finalColor, bgColor)"),
    # ResolveDebug: sky lanes show the checkerboard                                                                (T)
    (S, 769, "skyMask ? backgroundRGB : swr::pixfmt::RGBA8u::Pack({ finalColor, 1.0f })", "simd::select(skyMask, v_uint(backgroundRGB), swr::pixfmt::RGBA8u::Pack({ finalColor, 1.0f }))"),
    # CullMeshlets: inactive lanes fetch texel (0, 0)                                                              (T)
    (S, 834, "activeMask ? x : 0, activeMask ? y : 0", "simd::select(activeMask, x, v_int(0)), simd::select(activeMask, y, v_int(0))"),
]

# (file, first line, last line) wrapped in `#if 0 ... #endif`                                                       (X)
DISABLED = [
    # Image-based-lighting precomputation (importance sampling, irradiance / radiance maps, the BRDF LUT): runs once at
    # start-up in the Playground app, nothing in Draw / Resolve / CullMeshlets calls it (SURVEY.md §2 "out of scope").
    (S, 35, 219),
]
