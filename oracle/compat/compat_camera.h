// compat_camera.h — stands in for src/SwRast/Camera.h when the reference's Shading.cpp is compiled by oracle/ref_build.py.
// TEST INFRASTRUCTURE (see simd_gxx.h). Camera.h is the Playground's interactive camera (ImGui input, quaternions); the one
// thing the raster path takes from it is the free function GetInverseScreenProjMatrix (Camera.h:139-146), called by
// ShadingContext::Resolve (Shading.cpp:659). It is restated here on the GLM stand-in: inverse, then the three affine
// post-multiplications in the order the source applies them. GLM's own inverse is not available for comparison, so the
// wrapper (oracle/ref_api.cpp) can also hand Resolve the matrix the caller computed: comparisons against the restatement
// then cover everything downstream of this one host-side matrix, bit for bit.
#pragma once

#include <cstring>

#include <glm/glm.hpp>

namespace swr_compat { inline const float* g_invScreenProj = nullptr; }

static glm::mat4 GetInverseScreenProjMatrix(const glm::mat4& mat, glm::ivec2 viewSize, glm::vec2 subpixelOffset = glm::vec2(0.5f)) {
    if (swr_compat::g_invScreenProj != nullptr) {
        glm::mat4 given;
        std::memcpy(&given, swr_compat::g_invScreenProj, 64);
        return given;
    }
    glm::mat4 m = glm::inverse(mat);
    m = glm::translate(m, glm::vec3(-1.0f, -1.0f, 0.0f));
    m = glm::scale(m, glm::vec3(2.0f / glm::vec2(viewSize), 1.0f));
    m = glm::translate(m, glm::vec3(subpixelOffset, 0.0f));
    return m;
}
