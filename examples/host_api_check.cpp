// Exercises the reference-shaped host API of include/nexus_b200.hpp beyond a plain render: host objects edited in place and pushed
// by Scene::Update after Invalidate* (Scene.h:19-49, MeshInstance.h:22-53, Camera.h:19-35, AssetManager.h:18-44), light list edits,
// pixel query (PathTracer.h:23-27) and the pipelined display read-back.  Prints "host api ok" and exits 0 when every check holds.
//
//   g++ -std=c++17 -Iinclude examples/host_api_check.cpp -Lnexus_b200 -lnexus_b200 -Wl,-rpath,$PWD/nexus_b200 -o host_api_check
#include <cmath>
#include <cstdio>
#include <string>
#include "nexus_b200_import.hpp"

using namespace nexus;

#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "check failed at line %d: %s\n", __LINE__, #cond); return 2; } } while (0)

static std::vector<NXB::Triangle> quadXZ(float half, float y)
{
    return {NXB::Triangle{{-half, y, half}, {half, y, half}, {half, y, -half}}, NXB::Triangle{{-half, y, half}, {half, y, -half}, {-half, y, -half}}};
}
static double meanOf(const std::vector<float>& v) { double m = 0; for (float x : v) m += x; return m / (double)v.size(); }

int main(int argc, char** argv)
{
    try {
        const uint2 res{64, 48};
        Context ctx(0);
        Scene scene(ctx, res);
        AssetManager& am = scene.GetAssetManager();
        Material grey; grey.baseColor = {0.6f, 0.6f, 0.6f}; grey.roughness = 0.9f; grey.specularWeight = 0.0f;
        Material blue = grey; blue.baseColor = {0.1f, 0.2f, 0.9f};
        const uint32_t mGrey = am.AddMaterial(grey), mBlue = am.AddMaterial(blue);
        const uint32_t floorMesh = am.AddMesh("floor", mGrey, quadXZ(4.0f, 0.0f));
        const uint32_t plateMesh = am.AddMesh("plate", mGrey, quadXZ(0.5f, 0.0f));
        scene.CreateMeshInstance(floorMesh);
        MeshInstance& plate = scene.CreateMeshInstance(plateMesh, {0.0f, 1.0f, 0.0f});
        CHECK(scene.GetMeshInstances().size() == 2 && !scene.IsEmpty() && scene.GetMaterials().size() == 2);

        // camera above the plate looking straight down; right given explicitly because forward is parallel to +Y
        std::shared_ptr<Camera> cam = scene.GetCamera();
        cam->SetPosition({0.0f, 6.0f, 0.0f}); cam->SetForwardDirection({0.0f, -1.0f, 0.0f}); cam->SetRightDirection({1.0f, 0.0f, 0.0f});
        cam->SetHorizontalFOV(50.0f); cam->Invalidate();
        scene.GetRenderSettings().pathLength = 3;
        scene.GetRenderSettings().backgroundColor = {0.5f, 0.5f, 0.5f};
        const size_t li = scene.AddLight(Light{Light::Type::POINT, {0.0f, 4.0f, 0.0f}, {0, -1, 0}, {1, 1, 1}, 40.0f, 0});
        CHECK(li == 0 && scene.GetLights().size() == 1 && scene.IsInvalid());
        scene.Update();
        CHECK(!scene.IsInvalid());

        PathTracer pt(ctx, res);
        pt.UpdateDeviceScene(scene);
        // pixel query: the image centre sees the plate (instance 1), a corner sees the floor (instance 0)
        pt.SetPixelQuery(res.x / 2, res.y / 2); CHECK(pt.PixelQueryPending());
        pt.Render(scene);
        CHECK(pt.SynchronizePixelQuery() == 1 && pt.GetSelectedInstance() == 1 && !pt.PixelQueryPending());
        pt.SetPixelQuery(1, 1); pt.Render(scene); CHECK(pt.SynchronizePixelQuery() == 0);

        // move the plate away through the host object: SetPosition + InvalidateMeshInstance + Update, as the reference's UI does
        const NXB::AABB before = plate.GetBounds();
        CHECK(std::fabs(before.bmin[1] - 1.0f) < 1e-6f && std::fabs(before.bmax[0] - 0.5f) < 1e-6f);
        plate.SetPosition({3.0f, 1.0f, 3.0f}); plate.SetScale(0.5f); plate.SetRotationY(90.0f);
        scene.InvalidateMeshInstance(plate.index());
        CHECK(scene.IsInvalid());
        scene.Update();
        const std::array<float, 16> m = plate.GetTransfromationMatrix();
        CHECK(std::fabs(m[3] - 3.0f) < 1e-6f && std::fabs(m[7] - 1.0f) < 1e-6f && std::fabs(m[11] - 3.0f) < 1e-6f);     // translation column, row-major
        CHECK(std::fabs(m[0]) < 1e-6f && std::fabs(std::fabs(m[2]) - 0.5f) < 1e-6f);                                       // 90 degrees about Y, scale 0.5
        const NXB::AABB after = plate.GetBounds();
        CHECK(std::fabs(after.bmin[0] - 2.75f) < 1e-5f && std::fabs(after.bmax[2] - 3.25f) < 1e-5f);
        pt.Reset();
        pt.SetPixelQuery(res.x / 2, res.y / 2); pt.Render(scene); CHECK(pt.SynchronizePixelQuery() == 0);                 // the floor now

        // material edits: recolour the floor via GetMaterials + InvalidateMaterial, give the plate its own material
        pt.Reset(); pt.Render(scene, 16);
        const std::vector<float> greyImg = pt.ReadAccumulation();
        scene.GetMaterials()[mGrey].baseColor = {0.9f, 0.1f, 0.1f};
        am.InvalidateMaterial(mGrey);
        plate.AssignMaterial((int)mBlue);
        CHECK(scene.IsInvalid());
        scene.Update();
        pt.Reset(); pt.Render(scene, 16);
        const std::vector<float> redImg = pt.ReadAccumulation();
        double r = 0, g = 0; for (size_t i = 0; i < redImg.size(); i += 3) { r += redImg[i]; g += redImg[i + 1]; }
        CHECK(r > 2.0 * g && meanOf(greyImg) > 0.0);

        // lights: brighten through GetLights + InvalidateLight, then remove
        const double lit = meanOf(redImg);
        scene.GetLights()[0].intensity = 160.0f; scene.InvalidateLight(0); scene.Update();
        pt.Reset(); pt.Render(scene, 16);
        const double brighter = meanOf(pt.ReadAccumulation());
        scene.RemoveLight(0); CHECK(scene.GetLights().empty()); scene.Update();
        pt.Reset(); pt.Render(scene, 16);
        const double unlit = meanOf(pt.ReadAccumulation());
        CHECK(brighter > 1.2 * lit && unlit < lit && unlit > 0.0);

        // pipelined read-back delivers the same image as the blocking call
        std::vector<uint32_t> a((size_t)res.x * res.y), b;
        const int ticket = pt.Present(scene, a.data());
        const nx_frame_stats st = pt.PresentWait(ticket);
        b = pt.ReadRGBA8(scene);
        CHECK(a == b && st.frames == 16 && st.extension_rays >= 16ull * res.x * res.y);

        std::vector<nx_bvh2_node> hostNodes(4); NXB::FreeHostBVH(hostNodes); CHECK(hostNodes.empty());
        // optional: assets named on the command line (.obj / .glb) imported into a fresh scene through Scene::CreateMeshInstanceFromFile
        for (int a = 1; a < argc; a++) {
            Scene imported(ctx, res);
            imported.GetRenderSettings().backgroundColor = {0.6f, 0.7f, 0.8f};
            imported.GetRenderSettings().pathLength = 2;
            const size_t before = imported.GetMaterials().size();
            const std::vector<uint32_t> ids = CreateMeshInstanceFromFile(imported, "", argv[a]);
            CHECK(!ids.empty() && imported.GetMeshInstances().size() == ids.size() && imported.GetMaterials().size() > before);
            const NXB::AABB box = imported.GetMeshInstances()[0].GetBounds();
            std::shared_ptr<Camera> c = imported.GetCamera();
            c->SetPosition({0.5f * (box.bmin[0] + box.bmax[0]), 0.5f * (box.bmin[1] + box.bmax[1]), box.bmax[2] + 6.0f}); c->SetForwardDirection({0.0f, 0.0f, -1.0f}); c->Invalidate();
            imported.Update();
            PathTracer ip(ctx, res);
            ip.SetPixelQuery(res.x / 2, res.y / 2);
            ip.Render(imported, 4);
            const int32_t picked = ip.SynchronizePixelQuery();
            CHECK(meanOf(ip.ReadAccumulation()) > 0.0 && picked >= -1 && picked < (int32_t)ids.size());
            const std::vector<float> acc = ip.ReadAccumulation();
            double rgb[3] = {0, 0, 0};
            for (size_t k = 0; k + 2 < acc.size(); k += 3) { rgb[0] += acc[k]; rgb[1] += acc[k + 1]; rgb[2] += acc[k + 2]; }
            std::printf("imported %s: %zu instance(s), centre pixel sees instance %d, %zu texture(s), mean rgb %.5f %.5f %.5f\n", argv[a], ids.size(), picked,
                        imported.GetAssetManager().GetTextureCount(), 3.0 * rgb[0] / acc.size(), 3.0 * rgb[1] / acc.size(), 3.0 * rgb[2] / acc.size());
        }
        std::printf("host api ok\n");
    } catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
    return 0;
}
