// Tracy.hpp — the reference marks its phases with Tracy zones (Rasterizer.cpp:494,518,599,612) and names its workers (:872); the profiler client is not in
// this image, the macros expand to nothing. TEST INFRASTRUCTURE (see simd_gxx.h).
#pragma once
#define ZoneScoped
#define ZoneScopedN(name)
#define FrameMark
namespace tracy { inline void SetThreadName(const char*) {} }
