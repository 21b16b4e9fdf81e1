// TEST INFRASTRUCTURE ONLY.  The reference's own HOST code for what the scene layer hands to the kernels, compiled unmodified with g++
// (oracle/Makefile target `refcpu`): MeshInstance::ToDevice (Nexus/src/Scene/MeshInstance.h:36-66: T * Rz * Ry * Rx * S with Euler
// degrees, its inverse, the world box of the eight transformed corners of the mesh box) and Camera::ToDevice (Nexus/src/Scene/Camera.cpp:
// 130-156: right = cross(forward, +Y), up, viewport vectors, lower-left corner, lens radius).  Used to pin nx_scene_export_instances /
// nx_scene_export_camera, i.e. the matrices and the camera frame the product's kernels consume, against the reference's conventions.
#include <cstring>
#include "Scene/MeshInstance.h"
#include "Scene/Camera.h"
#include "Input.h"

// Camera.cpp references the UI input layer from Camera::OnUpdate (never called here); Input.cpp needs a GLFW window.
GLFWwindow* Input::m_Window = nullptr;
float2 Input::GetMousePosition() { return make_float2(0.0f, 0.0f); }
bool Input::IsKeyDown(int) { return false; }
bool Input::IsMouseButtonDown(int) { return false; }
void Input::SetCursorMode(int) {}

static_assert(sizeof(D_MeshInstance) == 160 && sizeof(D_Camera) == 88, "reference device layouts");

extern "C" void ref_host_instance(const float pos[3], const float rotDeg[3], const float scale[3], const float meshBounds[6],
                                   uint32_t meshIdx, uint32_t materialIdx, void* out160)
{
    MeshInstance mi;
    mi.SetTransform(make_float3(pos[0], pos[1], pos[2]), make_float3(rotDeg[0], rotDeg[1], rotDeg[2]), make_float3(scale[0], scale[1], scale[2]));
    mi.meshBounds.bMin = make_float3(meshBounds[0], meshBounds[1], meshBounds[2]);
    mi.meshBounds.bMax = make_float3(meshBounds[3], meshBounds[4], meshBounds[5]);
    mi.meshIdx = meshIdx; mi.materialIdx = materialIdx;
    const D_MeshInstance d = MeshInstance::ToDevice(mi);
    std::memcpy(out160, &d, sizeof(d));
}

extern "C" void ref_host_camera(const float pos[3], const float forward[3], float horizontalFov, float focusDist, float defocusAngle,
                                 uint32_t width, uint32_t height, void* out88)
{
    Camera cam(make_float3(pos[0], pos[1], pos[2]), make_float3(forward[0], forward[1], forward[2]), horizontalFov, make_uint2(width, height), focusDist, defocusAngle);
    D_Camera d;
    std::memset(&d, 0, sizeof(d));
    d = Camera::ToDevice(cam);
    std::memcpy(out88, &d, sizeof(d));
}
