// Host-side film geometry, filter table and image output.
#pragma once

#include <string>
#include <vector>

#include "../../../include/iile_b200.h"

namespace iile {

constexpr int kFilterTableWidth = 16;  // film.h:129

struct FilmGeometry {
    int cropped[4];       // x0 y0 x1 y1   croppedPixelBounds, film.cpp:55-59
    int sampleBounds[4];  // x0 y0 x1 y1   Film::GetSampleBounds, film.cpp:77-83
    int width() const { return cropped[2] - cropped[0]; }
    int height() const { return cropped[3] - cropped[1]; }
};
FilmGeometry filmGeometry(const IileRenderSetup &s);

// filterTable[y*16+x] = filter.Evaluate(((x+.5)rx/16, (y+.5)ry/16)), film.cpp:65-74
void buildFilterTable(const IileRenderSetup &s, float table[kFilterTableWidth * kFilterTableWidth]);

// Film::MergeFilmTile + to_rgb_array, film.cpp:135-148, 187-222
void resolveFilm(const IileRenderSetup &s, const float *filmRGBW, size_t nPixels, float *rgb);

// pbrt::WriteImage, imageio.cpp:81-122 (.pfm and .exr)
void writeImage(const std::string &path, const float *rgb, int xRes, int yRes, int fullX, int fullY,
                int offX, int offY);

}  // namespace iile
