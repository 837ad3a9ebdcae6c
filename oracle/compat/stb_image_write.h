// stb_image_write.h — see stb_image.h. TEST INFRASTRUCTURE.
#pragma once
#ifdef STB_IMAGE_WRITE_IMPLEMENTATION
extern "C" {
inline int stbi_write_png(const char*, int, int, int, const void*, int) { return 0; }
inline int stbi_write_hdr(const char*, int, int, int, const float*) { return 0; }
}
#endif
