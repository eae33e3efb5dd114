// C ABI: render entry points (stub until wavefront.cu lands).
#include <cstring>

#include "scene_handle.h"

using namespace iile;

IileScene::~IileScene() { freeRenderState(this); }

namespace iile {
void freeRenderState(IileScene *) {}
}  // namespace iile

extern "C" {

int iile_render_dev(IileScene *, const IileRenderParams *, float *, IileRenderStats *, void *) {
    return setError(IILE_E_UNSUPPORTED, "iile_render_dev: not built yet");
}
int iile_render(IileScene *, const IileRenderParams *, float *, IileRenderStats *) {
    return setError(IILE_E_UNSUPPORTED, "iile_render: not built yet");
}
int iile_film_resolve(const IileScene *scene, const float *filmRGBW, float *rgb) {
    if (!scene || !filmRGBW || !rgb) return setError(IILE_E_INVALID, "iile_film_resolve: null argument");
    FilmGeometry g = filmGeometry(scene->host.setup);
    resolveFilm(scene->host.setup, filmRGBW, (size_t)g.width() * g.height(), rgb);
    return IILE_OK;
}
int iile_write_image(const char *path, const float *rgb, int xRes, int yRes, int fullX, int fullY,
                     int offX, int offY) {
    if (!path || !rgb) return setError(IILE_E_INVALID, "iile_write_image: null argument");
    try {
        writeImage(path, rgb, xRes, yRes, fullX, fullY, offX, offY);
    } catch (const std::exception &ex) {
        return setError(IILE_E_IO, ex.what());
    }
    return IILE_OK;
}
int iile_render_hemispheres(IileScene *, size_t, const float *, const float *, int, int, float *,
                            float *, float *, IileRenderStats *) {
    return setError(IILE_E_UNSUPPORTED, "iile_render_hemispheres: not built yet");
}
}
