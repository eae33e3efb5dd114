// oracle/ref_patch/iile_hook.cpp — TEST INFRASTRUCTURE demonstrating the drop-in INSIDE the reference
// ("if the maintainer wants to keep the reference's parser"; VERDICT r1 missing #7).
//
// The reference's own parser (parser.cpp) and api.cpp state machine read the .pbrt file and build their
// objects exactly as always; pbrtWorldEnd(), patched by patch_api.sh, then calls iileWorldEndHook() with what
// RenderOptions holds.  This file walks those objects — GeometricPrimitive / Triangle / TriangleMesh / Sphere,
// Matte / Plastic / Mirror / Glass materials over ConstantTextures, PointLight / DiffuseAreaLight — into an
// IileSceneDesc, hands it to iile_scene_create() (include/iile_b200.h), renders with iile_render() and writes
// the image with the reference's own WriteImage.  Anything it does not flatten makes it decline, and the
// reference's CPU path renders as before: a drop-in, not a fork.
//
// The reference keeps these members private; a maintainer would add accessors or friend declarations.  Here
// the class definitions are read with their access specifiers opened (as oracle/ref_harness does), after
// every header they pull in has been included normally.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "pbrt.h"
#include "api.h"
#include "geometry.h"
#include "imageio.h"
#include "interaction.h"
#include "light.h"
#include "material.h"
#include "paramset.h"
#include "shape.h"
#include "spectrum.h"
#include "texture.h"
#include "transform.h"
#define private public
#define protected public
#include "lights/diffuse.h"
#include "lights/point.h"
#include "materials/glass.h"
#include "materials/matte.h"
#include "materials/mirror.h"
#include "materials/plastic.h"
#include "primitive.h"
#include "shapes/sphere.h"
#include "shapes/triangle.h"
#include "textures/constant.h"
#undef private
#undef protected

#include "../../include/iile_b200.h"
#include "iile_hook.h"

namespace pbrt {

namespace {

struct Decline {
    std::string why;
};

void matrixOut(const Matrix4x4 &m, float *out) {
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) out[4 * r + c] = m.m[r][c];
}

void spectrumOf(const std::shared_ptr<Texture<Spectrum>> &t, float out[3], const char *what) {
    auto *c = dynamic_cast<const ConstantTexture<Spectrum> *>(t.get());
    if (!c) throw Decline{std::string("non-constant spectrum texture for ") + what};
    Float rgb[3];
    c->value.ToRGB(rgb);
    out[0] = rgb[0], out[1] = rgb[1], out[2] = rgb[2];
}
float floatOf(const std::shared_ptr<Texture<Float>> &t, const char *what) {
    auto *c = dynamic_cast<const ConstantTexture<Float> *>(t.get());
    if (!c) throw Decline{std::string("non-constant float texture for ") + what};
    return c->value;
}

// the option block: what MakeCamera / MakeFilm / MakeFilter / MakeSampler / MakeIntegrator / MakeAccelerator read
void fillSetup(const IileHookArgs &a, IileRenderSetup *S) {
    std::memset(S, 0, sizeof(*S));
    if (a.CameraName != "perspective") throw Decline{"Camera \"" + a.CameraName + "\""};
    if (a.FilmName != "image") throw Decline{"Film \"" + a.FilmName + "\""};
    if (a.IntegratorName != "path") throw Decline{"Integrator \"" + a.IntegratorName + "\""};
    if (a.AcceleratorName != "bvh") throw Decline{"Accelerator \"" + a.AcceleratorName + "\""};
    // film.cpp:259-303
    S->xRes = a.FilmParams.FindOneInt("xresolution", 1280);
    S->yRes = a.FilmParams.FindOneInt("yresolution", 720);
    if (PbrtOptions.quickRender) throw Decline{"--quick"};
    std::string fn = a.FilmParams.FindOneString("filename", "");
    if (PbrtOptions.imageFile != "") fn = PbrtOptions.imageFile;
    if (fn == "") fn = "pbrt.exr";
    std::snprintf(S->filename, sizeof(S->filename), "%s", fn.c_str());
    S->cropWindow[0] = 0, S->cropWindow[1] = 1, S->cropWindow[2] = 0, S->cropWindow[3] = 1;
    int cwi;
    if (const Float *cr = a.FilmParams.FindFloat("cropwindow", &cwi)) {
        if (cwi != 4) throw Decline{"cropwindow"};
        S->cropWindow[0] = Clamp(std::min(cr[0], cr[1]), 0.f, 1.f);
        S->cropWindow[1] = Clamp(std::max(cr[0], cr[1]), 0.f, 1.f);
        S->cropWindow[2] = Clamp(std::min(cr[2], cr[3]), 0.f, 1.f);
        S->cropWindow[3] = Clamp(std::max(cr[2], cr[3]), 0.f, 1.f);
    }
    S->filmScale = a.FilmParams.FindOneFloat("scale", 1.f);
    S->diagonal = a.FilmParams.FindOneFloat("diagonal", 35.f);
    S->maxSampleLuminance = a.FilmParams.FindOneFloat("maxsampleluminance", Infinity);
    // filters/*.cpp
    const ParamSet &fp = a.FilterParams;
    if (a.FilterName == "box") {
        S->filterType = IILE_FILTER_BOX;
        S->filterRadius[0] = fp.FindOneFloat("xwidth", 0.5f), S->filterRadius[1] = fp.FindOneFloat("ywidth", 0.5f);
    } else if (a.FilterName == "gaussian") {
        S->filterType = IILE_FILTER_GAUSSIAN;
        S->filterRadius[0] = fp.FindOneFloat("xwidth", 2.f), S->filterRadius[1] = fp.FindOneFloat("ywidth", 2.f);
        S->filterParam[0] = fp.FindOneFloat("alpha", 2.f);
    } else if (a.FilterName == "mitchell") {
        S->filterType = IILE_FILTER_MITCHELL;
        S->filterRadius[0] = fp.FindOneFloat("xwidth", 2.f), S->filterRadius[1] = fp.FindOneFloat("ywidth", 2.f);
        S->filterParam[0] = fp.FindOneFloat("B", 1.f / 3.f), S->filterParam[1] = fp.FindOneFloat("C", 1.f / 3.f);
    } else if (a.FilterName == "sinc") {
        S->filterType = IILE_FILTER_SINC;
        S->filterRadius[0] = fp.FindOneFloat("xwidth", 4.f), S->filterRadius[1] = fp.FindOneFloat("ywidth", 4.f);
        S->filterParam[0] = fp.FindOneFloat("tau", 3.f);
    } else if (a.FilterName == "triangle") {
        S->filterType = IILE_FILTER_TRIANGLE;
        S->filterRadius[0] = fp.FindOneFloat("xwidth", 2.f), S->filterRadius[1] = fp.FindOneFloat("ywidth", 2.f);
    } else
        throw Decline{"Filter \"" + a.FilterName + "\""};
    // perspective.cpp:293-329
    matrixOut(a.CameraToWorld.GetMatrix(), S->cameraToWorld);
    matrixOut(a.CameraToWorld.GetInverseMatrix(), S->worldToCamera);
    const ParamSet &cp = a.CameraParams;
    S->shutterOpen = cp.FindOneFloat("shutteropen", 0.f);
    S->shutterClose = cp.FindOneFloat("shutterclose", 1.f);
    if (S->shutterClose < S->shutterOpen) std::swap(S->shutterClose, S->shutterOpen);
    S->lensRadius = cp.FindOneFloat("lensradius", 0.f);
    S->focalDistance = cp.FindOneFloat("focaldistance", 1e6f);
    const Float frame = cp.FindOneFloat("frameaspectratio", Float(S->xRes) / Float(S->yRes));
    if (frame > 1.f) {
        S->screenWindow[0] = -frame, S->screenWindow[1] = frame, S->screenWindow[2] = -1.f, S->screenWindow[3] = 1.f;
    } else {
        S->screenWindow[0] = -1.f, S->screenWindow[1] = 1.f, S->screenWindow[2] = -1.f / frame, S->screenWindow[3] = 1.f / frame;
    }
    int swi;
    if (const Float *sw = cp.FindFloat("screenwindow", &swi)) {
        if (swi != 4) throw Decline{"screenwindow"};
        for (int i = 0; i < 4; ++i) S->screenWindow[i] = sw[i];
    }
    S->fov = cp.FindOneFloat("fov", 90.f);
    const Float halffov = cp.FindOneFloat("halffov", -1.f);
    if (halffov > 0.f) S->fov = 2.f * halffov;
    S->cameraType = IILE_CAMERA_PERSPECTIVE;
    // samplers
    const ParamSet &sp = a.SamplerParams;
    if (a.SamplerName == "halton") {
        S->samplerType = IILE_SAMPLER_HALTON;
        S->pixelSamples = sp.FindOneInt("pixelsamples", 16);
        S->samplePixelCenter = sp.FindOneBool("samplepixelcenter", false);
    } else if (a.SamplerName == "sobol") {
        S->samplerType = IILE_SAMPLER_SOBOL;
        S->pixelSamples = RoundUpPow2(sp.FindOneInt("pixelsamples", 16));
    } else if (a.SamplerName == "random") {
        S->samplerType = IILE_SAMPLER_RANDOM;
        S->pixelSamples = sp.FindOneInt("pixelsamples", 4);
    } else
        throw Decline{"Sampler \"" + a.SamplerName + "\""};
    if (std::getenv("IILE_PATH_SAMPLES_OVERRIDE")) throw Decline{"IILE_PATH_SAMPLES_OVERRIDE"};
    // path.cpp:196-231
    const ParamSet &ip = a.IntegratorParams;
    S->integratorType = IILE_INTEGRATOR_PATH;
    S->maxDepth = ip.FindOneInt("maxdepth", 5);
    S->rrThreshold = ip.FindOneFloat("rrthreshold", 1.f);
    int np;
    if (const int *pb = ip.FindInt("pixelbounds", &np)) {
        if (np != 4) throw Decline{"pixelbounds"};
        S->hasPixelBounds = 1;
        for (int i = 0; i < 4; ++i) S->pixelBounds[i] = pb[i];
    }
    const std::string ls = ip.FindOneString("lightsamplestrategy", "spatial");
    S->lightSampleStrategy = ls == "uniform" ? IILE_LIGHTSTRATEGY_UNIFORM
                             : ls == "power" ? IILE_LIGHTSTRATEGY_POWER : IILE_LIGHTSTRATEGY_SPATIAL;
    // bvh.cpp:740-760
    const std::string sm = a.AcceleratorParams.FindOneString("splitmethod", "sah");
    S->splitMethod = sm == "hlbvh" ? IILE_SPLIT_HLBVH : sm == "middle" ? IILE_SPLIT_MIDDLE
                     : sm == "equal" ? IILE_SPLIT_EQUAL : IILE_SPLIT_SAH;
    S->maxNodePrims = a.AcceleratorParams.FindOneInt("maxnodeprims", 4);
}

}  // namespace

bool iileWorldEndHook(const IileHookArgs &a) {
    if (std::getenv("IILE_B200") == nullptr) return false;  // the switch a maintainer would put behind --b200
    try {
        IileRenderSetup setup;
        fillSetup(a, &setup);
        std::vector<float> P, N, UV;
        std::vector<IilePrim> prims;
        std::vector<IileSphere> spheres;
        std::vector<IileMaterial> materials;
        std::vector<IileLight> lights;
        std::map<const TriangleMesh *, int> meshBase;
        std::map<const Material *, int> materialIndex;
        std::map<const Light *, int> lightOfArea;
        bool anyN = false, anyUV = false;
        // lights first, in RenderOptions' order (the order Scene::lights and the light distributions see)
        for (const std::shared_ptr<Light> &l : a.lights) {
            IileLight L;
            std::memset(&L, 0, sizeof(L));
            L.prim = -1;
            Float rgb[3];
            if (auto *pl = dynamic_cast<const PointLight *>(l.get())) {
                L.type = IILE_LIGHT_POINT;
                L.pos[0] = pl->pLight.x, L.pos[1] = pl->pLight.y, L.pos[2] = pl->pLight.z;
                pl->I.ToRGB(rgb);
            } else if (auto *al = dynamic_cast<const DiffuseAreaLight *>(l.get())) {
                L.type = IILE_LIGHT_DIFFUSE_AREA;
                al->Lemit.ToRGB(rgb);
                L.twoSided = (al->twoSided ? 1 : 0) | ((std::max(al->nSamples, 1) - 1) << 8);
                L.area = al->area;
                lightOfArea[l.get()] = (int)lights.size();
            } else
                throw Decline{"a light that is neither a point nor a diffuse area light"};
            L.L[0] = rgb[0], L.L[1] = rgb[1], L.L[2] = rgb[2];
            lights.push_back(L);
        }
        for (const std::shared_ptr<Primitive> &p : a.primitives) {
            auto *gp = dynamic_cast<const GeometricPrimitive *>(p.get());
            if (!gp) throw Decline{"a primitive that is not a GeometricPrimitive (instancing / animation)"};
            if (gp->mediumInterface.inside || gp->mediumInterface.outside) throw Decline{"participating media"};
            IilePrim pr;
            std::memset(&pr, 0, sizeof(pr));
            // material
            pr.material = -1;
            if (const Material *m = gp->material.get()) {
                auto it = materialIndex.find(m);
                if (it == materialIndex.end()) {
                    IileMaterial M;
                    std::memset(&M, 0, sizeof(M));
                    if (auto *mm = dynamic_cast<const MatteMaterial *>(m)) {
                        if (mm->bumpMap) throw Decline{"bump map"};
                        M.type = IILE_MAT_MATTE;
                        spectrumOf(mm->Kd, M.Kd, "matte Kd");
                        M.sigma = floatOf(mm->sigma, "matte sigma");
                    } else if (auto *pm = dynamic_cast<const PlasticMaterial *>(m)) {
                        if (pm->bumpMap) throw Decline{"bump map"};
                        M.type = IILE_MAT_PLASTIC;
                        spectrumOf(pm->Kd, M.Kd, "plastic Kd");
                        spectrumOf(pm->Ks, M.Ks, "plastic Ks");
                        M.roughness = floatOf(pm->roughness, "plastic roughness");
                        M.remapRoughness = pm->remapRoughness;
                    } else if (auto *rm = dynamic_cast<const MirrorMaterial *>(m)) {
                        if (rm->bumpMap) throw Decline{"bump map"};
                        M.type = IILE_MAT_MIRROR;
                        spectrumOf(rm->Kr, M.Ks, "mirror Kr");
                    } else if (auto *gm = dynamic_cast<const GlassMaterial *>(m)) {
                        if (gm->bumpMap) throw Decline{"bump map"};
                        M.type = IILE_MAT_GLASS;
                        spectrumOf(gm->Kr, M.Ks, "glass Kr");
                        spectrumOf(gm->Kt, M.Kt, "glass Kt");
                        M.eta = floatOf(gm->index, "glass index");
                        M.uRoughness = floatOf(gm->uRoughness, "glass uroughness");
                        M.vRoughness = floatOf(gm->vRoughness, "glass vroughness");
                        M.remapRoughness = gm->remapRoughness;
                    } else
                        throw Decline{"a material other than matte / plastic / mirror / glass"};
                    materials.push_back(M);
                    it = materialIndex.emplace(m, (int)materials.size() - 1).first;
                }
                pr.material = it->second;
            }
            // shape
            const Shape *shape = gp->shape.get();
            if (auto *tri = dynamic_cast<const Triangle *>(shape)) {
                const TriangleMesh *mesh = tri->mesh.get();
                if (mesh->alphaMask || mesh->shadowAlphaMask) throw Decline{"alpha mask"};
                if (mesh->s) throw Decline{"per-vertex tangents"};
                auto it = meshBase.find(mesh);
                if (it == meshBase.end()) {
                    it = meshBase.emplace(mesh, (int)(P.size() / 3)).first;
                    for (int i = 0; i < mesh->nVertices; ++i) {
                        P.push_back(mesh->p[i].x), P.push_back(mesh->p[i].y), P.push_back(mesh->p[i].z);
                        const Normal3f n = mesh->n ? mesh->n[i] : Normal3f(0, 0, 0);
                        N.push_back(n.x), N.push_back(n.y), N.push_back(n.z);
                        const Point2f uv = mesh->uv ? mesh->uv[i] : Point2f(0, 0);
                        UV.push_back(uv.x), UV.push_back(uv.y);
                    }
                    anyN = anyN || mesh->n;
                    anyUV = anyUV || mesh->uv;
                }
                pr.type = IILE_PRIM_TRIANGLE;
                for (int k = 0; k < 3; ++k) pr.v[k] = it->second + tri->v[k];
                if (mesh->n) pr.flags |= IILE_PRIM_HAS_N;
                if (mesh->uv) pr.flags |= IILE_PRIM_HAS_UV;
            } else if (auto *sph = dynamic_cast<const Sphere *>(shape)) {
                IileSphere s;
                std::memset(&s, 0, sizeof(s));
                matrixOut(sph->ObjectToWorld->GetMatrix(), s.objectToWorld);
                matrixOut(sph->WorldToObject->GetMatrix(), s.worldToObject);
                s.radius = sph->radius, s.zMin = sph->zMin, s.zMax = sph->zMax;
                s.thetaMin = sph->thetaMin, s.thetaMax = sph->thetaMax, s.phiMax = sph->phiMax;
                s.phiMaxDegrees = Degrees(sph->phiMax);
                s.flags = (sph->reverseOrientation ? 1u : 0u) | (sph->transformSwapsHandedness ? 2u : 0u);
                pr.type = IILE_PRIM_SPHERE;
                pr.v[0] = (int)spheres.size();
                spheres.push_back(s);
            } else
                throw Decline{"a shape other than a triangle or a sphere"};
            if (shape->reverseOrientation ^ shape->transformSwapsHandedness) pr.flags |= IILE_PRIM_FLIP_NORMAL;
            pr.light = -1;
            if (gp->areaLight) {
                auto it = lightOfArea.find(gp->areaLight.get());
                if (it == lightOfArea.end()) throw Decline{"an area light that is not in the scene's light list"};
                pr.light = it->second;
                lights[it->second].prim = (int)prims.size();
            }
            prims.push_back(pr);
        }
        if (prims.empty()) throw Decline{"an empty scene"};
        IileSceneDesc d;
        std::memset(&d, 0, sizeof(d));
        d.nVerts = (uint32_t)(P.size() / 3);
        d.P = P.data();
        d.N = anyN ? N.data() : nullptr;
        d.UV = anyUV ? UV.data() : nullptr;
        d.nPrims = (uint32_t)prims.size();
        d.prims = prims.data();
        d.nSpheres = (uint32_t)spheres.size();
        d.spheres = spheres.data();
        d.nMaterials = (uint32_t)materials.size();
        d.materials = materials.data();
        d.nLights = (uint32_t)lights.size();
        d.lights = lights.data();
        d.setup = setup;
        IileScene *scene = nullptr;
        if (const char *flat = std::getenv("IILE_B200_FLAT")) {
            // test aid: only flatten (no device needed) and save what iile_scene_create received
            if (iile_scene_create(&d, /*upload=*/0, &scene) != IILE_OK || iile_scene_save_flat(scene, flat) != IILE_OK)
                throw Decline{std::string("iile_scene_create / save_flat: ") + iile_last_error()};
            iile_scene_destroy(scene);
            std::fprintf(stderr, "iile hook: flattened scene saved to %s (IILE_B200_FLAT); nothing rendered\n", flat);
            return true;
        }
        if (iile_scene_create(&d, /*upload=*/1, &scene) != IILE_OK) throw Decline{std::string("iile_scene_create: ") + iile_last_error()};
        IileSceneInfo info;
        iile_scene_info(scene, &info);
        const int w = info.croppedBounds[2] - info.croppedBounds[0], h = info.croppedBounds[3] - info.croppedBounds[1];
        std::vector<float> film((size_t)w * h * 4), rgb((size_t)w * h * 3);
        IileRenderParams rp;
        std::memset(&rp, 0, sizeof(rp));
        rp.clearFilm = 1;
        IileRenderStats st;
        const bool ok = iile_render(scene, &rp, film.data(), &st) == IILE_OK &&
                        iile_film_resolve(scene, film.data(), rgb.data()) == IILE_OK;
        const std::string err = ok ? "" : iile_last_error();
        iile_scene_destroy(scene);
        if (!ok) throw Decline{"render: " + err};
        // the writer Film::WriteImage ends in (film.cpp:187-235 -> imageio.cpp:81-122)
        WriteImage(setup.filename, rgb.data(),
                   Bounds2i(Point2i(info.croppedBounds[0], info.croppedBounds[1]),
                            Point2i(info.croppedBounds[2], info.croppedBounds[3])),
                   Point2i(setup.xRes, setup.yRes));
        std::fprintf(stderr, "iile hook: %u primitives (%u triangles' vertices, %zu spheres), %zu materials, %zu lights rendered by the "
                             "B200 library: %llu camera samples, %llu rays, %llu launches -> %s\n",
                     d.nPrims, d.nVerts, spheres.size(), materials.size(), lights.size(), (unsigned long long)st.cameraRays,
                     (unsigned long long)(st.closestRays + st.shadowRays), (unsigned long long)st.launches, setup.filename);
        return true;
    } catch (const Decline &dcl) {
        std::fprintf(stderr, "iile hook: not taken (%s); the reference's CPU path renders\n", dcl.why.c_str());
        return false;
    }
}

}  // namespace pbrt
