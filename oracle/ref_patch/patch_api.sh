#!/bin/bash
# oracle/ref_patch/patch_api.sh <reference api.cpp> <out api.cpp>
# Produces the patched copy of src/core/api.cpp that INTEGRATION.md B describes: pbrtWorldEnd() offers the
# parsed scene (RenderOptions' strings, ParamSets and built objects) to iileWorldEndHook() first and runs
# MakeIntegrator / MakeScene / Render only when the hook declines.  The reference tree is not touched; the
# copy lives under oracle/_ref/ (git-ignored).
set -e
src=$1; out=$2
grep -q 'std::unique_ptr<Integrator> integrator(renderOptions->MakeIntegrator());' "$src"
grep -q 'if (scene && integrator) integrator->Render(\*scene);' "$src"
sed -e 's|^#include "api.h"|#include "api.h"\n#include "iile_hook.h"|' \
    -e 's|^        std::unique_ptr<Integrator> integrator(renderOptions->MakeIntegrator());|        IileHookArgs iileArgs{renderOptions->FilterName, renderOptions->FilterParams, renderOptions->FilmName,\n                              renderOptions->FilmParams, renderOptions->SamplerName, renderOptions->SamplerParams,\n                              renderOptions->AcceleratorName, renderOptions->AcceleratorParams,\n                              renderOptions->IntegratorName, renderOptions->IntegratorParams,\n                              renderOptions->CameraName, renderOptions->CameraParams,\n                              renderOptions->CameraToWorld[0], renderOptions->lights, renderOptions->primitives};\n        const bool iileRendered = iileWorldEndHook(iileArgs);\n        std::unique_ptr<Integrator> integrator(iileRendered ? nullptr : renderOptions->MakeIntegrator());|' \
    -e 's|^        std::unique_ptr<Scene> scene(renderOptions->MakeScene());|        std::unique_ptr<Scene> scene(iileRendered ? nullptr : renderOptions->MakeScene());|' \
    "$src" > "$out"
grep -q iileWorldEndHook "$out"
