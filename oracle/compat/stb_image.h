// stb_image.h — declarations only. TEST INFRASTRUCTURE (see simd_gxx.h): the reference's ImageHelpers.cpp includes the stb image
// decoder (nothings/stb, unpinned master, absent from this image) for its file loaders; the raster path never decodes a
// file, so the loaders link against stand-ins that report failure.
#pragma once
typedef unsigned char stbi_uc;
#ifdef STB_IMAGE_IMPLEMENTATION
extern "C" {
inline stbi_uc* stbi_load(const char*, int*, int*, int*, int) { return nullptr; }
inline float* stbi_loadf(const char*, int*, int*, int*, int) { return nullptr; }
inline void stbi_image_free(void*) {}
}
#endif
