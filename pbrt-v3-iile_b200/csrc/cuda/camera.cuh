// Camera rays shared by the wavefront's generate kernel and the IISPT task kernels: the camera-space ray
// of EnvironmentCamera / OrthographicCamera / PerspectiveCamera::GenerateRay (environment.cpp:44-57,
// orthographic.cpp:44-67, perspective.cpp:100-149) and CameraToWorld(ray) (transform.h:251-263).
#pragma once

#include "shading.cuh"
#include "wavefront.h"

namespace iile {

// (fx, fy) = CameraSample::pFilm, (lu, lv) = CameraSample::pLens
__device__ __forceinline__ void cameraSpaceRay(const RenderConsts &rc, float fx, float fy, float lu, float lv, V3f *oOut,
                                               V3f *dOut) {
    V3f o(0, 0, 0), d;
    if (rc.cameraType == IILE_CAMERA_ENVIRONMENT) {  // EnvironmentCamera::GenerateRay, environment.cpp:44-57
        const float theta = IILE_PI * fy / rc.fullRes[1];
        const float phi = 2 * IILE_PI * fx / rc.fullRes[0];
        float st, ct, sp, cp;  // std::sin / std::cos of a float: the host's sinf / cosf (sincosLibm)
        sincosLibm(theta, &st, &ct);
        sincosLibm(phi, &sp, &cp);
        d = V3f(st * cp, ct, st * sp);
    } else if (rc.cameraType == IILE_CAMERA_ORTHOGRAPHIC) {  // OrthographicCamera::GenerateRay, orthographic.cpp:44-67
        o = xfPoint(rc.rasterToCamera, V3f(fx, fy, 0));
        d = V3f(0, 0, 1);
        if (rc.lensRadius > 0) {
            float lx, ly;
            concentricSampleDisk(lu, lv, &lx, &ly);
            lx *= rc.lensRadius;
            ly *= rc.lensRadius;
            float ft = rc.focalDistance / d.z;
            V3f pFocus = o + d * ft;
            o = V3f(lx, ly, 0);
            d = normalize(pFocus - o);
        }
    } else {
        const float *m = rc.rasterToCamera;
        V3f pc = xfPoint(m, V3f(fx, fy, 0));
        d = normalize(pc);
        if (rc.lensRadius > 0) {
            float lx, ly;
            concentricSampleDisk(lu, lv, &lx, &ly);
            lx *= rc.lensRadius;
            ly *= rc.lensRadius;
            float ft = rc.focalDistance / d.z;
            V3f pFocus = o + d * ft;
            o = V3f(lx, ly, 0);
            d = normalize(pFocus - o);
        }
    }
    *oOut = o;
    *dOut = d;
}

// HemisphericCamera::GenerateRay's camera-space direction (hemispheric.cpp:24-28): the film spans half of the
// sphere of directions (phi in [0, pi))
__device__ __forceinline__ V3f hemiCameraSpaceDir(int hemiSize, float fx, float fy) {
    const float theta = IILE_PI * fy / hemiSize;
    const float phi = IILE_PI * fx / hemiSize;
    // std::sin / std::cos of a float: the host's sinf / cosf, restated (libm_exact.cuh) — CUDA's float versions are
    // 1-2 ulp off, and even double-then-round misses glibc's value on 1-2 % of the arguments
    float sT, cT, sP, cP;
    sincosLibm(theta, &sT, &cT);
    sincosLibm(phi, &sP, &cP);
    return V3f(sT * cP, cT, sT * sP);
}

// CameraToWorld(ray): Transform::operator()(const Ray&), transform.h:251-263
__device__ __forceinline__ void rayToWorld(const float *c2w, V3f *oInOut, V3f *dInOut) {
    const V3f o = *oInOut, d = *dInOut;
    const float *m = c2w;
    const float g3 = gammaF(3);
    V3f ow = xfPoint(m, o);
    V3f oe(g3 * (fabsf(m[0] * o.x) + fabsf(m[1] * o.y) + fabsf(m[2] * o.z) + fabsf(m[3])),
           g3 * (fabsf(m[4] * o.x) + fabsf(m[5] * o.y) + fabsf(m[6] * o.z) + fabsf(m[7])),
           g3 * (fabsf(m[8] * o.x) + fabsf(m[9] * o.y) + fabsf(m[10] * o.z) + fabsf(m[11])));
    V3f dw(m[0] * d.x + m[1] * d.y + m[2] * d.z, m[4] * d.x + m[5] * d.y + m[6] * d.z,
           m[8] * d.x + m[9] * d.y + m[10] * d.z);
    float l2 = lengthSq(dw);
    if (l2 > 0) {
        float dt = dot(absV(dw), oe) / l2;
        ow = ow + dw * dt;
    }
    *oInOut = ow;
    *dInOut = dw;
}

}  // namespace iile
