// The map transforms IisptRenderRunner applies around the network
// (src/integrators/iisptrenderrunner.cpp:1041-1136, ImageFilm ops src/film/imagefilm.cpp:203-379),
// batched over hemispheres and kept in HBM: RenderView's film / normal / distance maps go in, the
// network's input tensor comes out (and back) without the per-hemisphere stdio round trip of
// IisptNnConnector (iisptnnconnector.cpp:36-148).
//
// One block per hemisphere.  Tensors are laid out as main_stdio_net.read_input builds them:
// [7][size][size] planes (R, G, B, nx, ny, nz, distance), row Y of a plane = ImageFilm row Y =
// camera pixel row size-1-Y (Film::to_intensity_film and set_camera_coord both flip y,
// film.cpp:238-256, imagefilm.cpp:26-31).
//
// The reference's means are serial double-precision sums over the film in array order
// (imagefilm.cpp:203-254); one thread per hemisphere walks the same order so the means — and
// everything scaled by them — carry the same bits.  std::log(1.0 + v) is evaluated in double and
// rounded once, std::exp(float) is glibc's (correctly rounded in practice) expf: exp in double,
// rounded to float.
#include "wavefront.h"

namespace iile {

namespace {

constexpr int kMapBlock = 256;

// Film::to_rgb_array for one pixel (film.cpp:187-226 with splatScale 1; RGBToXYZ / XYZToRGB,
// spectrum.h:56-66) — the arithmetic of host/film_host.cpp resolveFilm
__device__ __forceinline__ void resolvePixel(float4 p, float scale, float *o) {
    const float r = p.x, g = p.y, b = p.z, w = p.w;
    const float X = 0.412453f * r + 0.357580f * g + 0.180423f * b;
    const float Y = 0.212671f * r + 0.715160f * g + 0.072169f * b;
    const float Z = 0.019334f * r + 0.119193f * g + 0.950227f * b;
    o[0] = 3.240479f * X - 1.537150f * Y - 0.498535f * Z;
    o[1] = -0.969256f * X + 1.875991f * Y + 0.041556f * Z;
    o[2] = 0.055648f * X - 0.204043f * Y + 1.057311f * Z;
    if (w != 0) {
        const float inv = 1.f / w;
        for (int c = 0; c < 3; ++c) o[c] = fmaxf(0.f, o[c] * inv);
    }
    for (int c = 0; c < 3; ++c) {
        o[c] += 1.f * 0.f;
        o[c] = o[c] * scale;
    }
}

// ImageFilm::positiveLog (imagefilm.cpp:324-333)
__device__ __forceinline__ float positiveLog(float v) {
    float vv = v;
    if (vv <= 0.f) vv = 0.f;
    return (float)log(1.0 + (double)vv);
}

// IisptRenderRunner::normalizeMapsDownstream (iisptrenderrunner.cpp:1041-1095).  kStage: the four
// planes the serial sums walk (R, G, B, distance) are staged in shared memory (16 B per pixel; the
// 32x32 maps of the reference take 16 KB), so the two summing threads read at shared-memory latency.
template <bool kStage>
__global__ void __launch_bounds__(kMapBlock) hemiToNNKernel(const float4 *__restrict__ film,
                                                            const float *__restrict__ normals,
                                                            const float *__restrict__ distance, int size,
                                                            float filmScale, float *__restrict__ nn,
                                                            float *__restrict__ means) {
    const size_t h = blockIdx.x;
    const int px = size * size;
    const float4 *f = film + h * px;
    const float *nr = normals + h * 3 * (size_t)px, *ds = distance + h * (size_t)px;
    float *out = nn + h * 7 * (size_t)px;
    __shared__ float sRatio, sZRatio;
    extern __shared__ float stage[];  // kStage: [4][px] = R, G, B, distance
    // raw maps in ImageFilm order
    for (int i = threadIdx.x; i < px; i += blockDim.x) {
        const int Y = i / size, x = i - Y * size;
        const int src = (size - 1 - Y) * size + x;
        float rgb[3];
        resolvePixel(f[src], filmScale, rgb);
        for (int c = 0; c < 3; ++c) {
            if (kStage) stage[c * px + i] = rgb[c];
            else out[c * px + i] = rgb[c];
            // NormalFilm: ImageFilm::normalize(-1, 1) -> mid 0, r 1 (imagefilm.cpp:359-379)
            float v = nr[3 * src + c];
            v = v - 0.f;
            v = v / 1.f;
            out[(3 + c) * px + i] = v < -1.f ? -1.f : (v > 1.f ? 1.f : v);
        }
        if (kStage) stage[3 * px + i] = ds[src];
        else out[6 * px + i] = ds[src];
    }
    __syncthreads();
    const float *pr = kStage ? stage : out, *pg = kStage ? stage + px : out + px;
    const float *pb = kStage ? stage + 2 * px : out + 2 * px, *pz = kStage ? stage + 3 * px : out + 6 * px;
    if (threadIdx.x == 0) {
        // computeMeanChannels + computeMean (imagefilm.cpp:203-254): double sums in array order
        double rs = 0, gs = 0, bs = 0, all = 0;
#pragma unroll 8
        for (int i = 0; i < px; ++i) {
            const float r = pr[i], g = pg[i], b = pb[i];
            rs += r;
            gs += g;
            bs += b;
            all += (r + g + b);
        }
        means[3 * h] = (float)(rs / px);
        means[3 * h + 1] = (float)(gs / px);
        means[3 * h + 2] = (float)(bs / px);
        const float intensityMean = (float)(all / (3 * px));
        sRatio = intensityMean == 0.0 ? (float)0.0 : (float)(1.0 / (10.0 * intensityMean));
    } else if (threadIdx.x == 32) {
        double zs = 0;
#pragma unroll 8
        for (int i = 0; i < px; ++i) zs += pz[i];
        const float zMean = (float)(zs / px);
        float distanceDiv = (float)(10.0 * (zMean + 1.0));
        if (distanceDiv == 0) distanceDiv = 1.0f;
        sZRatio = (float)(1.0 / distanceDiv);
    }
    __syncthreads();
    const float ratio = sRatio, zRatio = sZRatio, minusTenth = (float)-0.1, one = (float)1.0;
    for (int i = threadIdx.x; i < px; i += blockDim.x) {
        for (int c = 0; c < 3; ++c) {
            float v = (kStage ? stage[c * px + i] : out[c * px + i]) * ratio;
            v = positiveLog(v);
            out[c * px + i] = v + minusTenth;
        }
        float z = pz[i] + one;
        z = z * zRatio;
        z = positiveLog(z);
        out[6 * px + i] = z + minusTenth;
    }
}

// IisptRenderRunner::transformMapsUpstream (iisptrenderrunner.cpp:1098-1136): network output
// [3][size][size] -> intensity [size][size][3] (what IntensityFilm::populate_from_float_array reads)
__global__ void __launch_bounds__(kMapBlock) hemiFromNNKernel(const float *__restrict__ nnOut,
                                                              const float *__restrict__ means, int size,
                                                              float *__restrict__ intensity) {
    const size_t h = blockIdx.x;
    const int px = size * size;
    const float *in = nnOut + h * 3 * (size_t)px;
    float *out = intensity + h * 3 * (size_t)px;
    __shared__ float sMul[3];
    // ImageFilm::positiveLogInverse (imagefilm.cpp:337-346)
    for (int i = threadIdx.x; i < px; i += blockDim.x)
        for (int c = 0; c < 3; ++c) {
            float y = in[c * px + i];
            if (y < 0.f) y = 0.f;
            const float e = (float)exp((double)y);
            out[3 * i + c] = (float)((double)e - 1.0);
        }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        double s = 0;
        for (int i = 0; i < px; ++i) s += out[3 * i + c];
        const float actual = (float)(s / px);
        sMul[c] = actual > 1e-10 ? means[3 * h + c] / actual : (float)0.0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < px; i += blockDim.x)
        for (int c = 0; c < 3; ++c) out[3 * i + c] = sMul[c] * out[3 * i + c];
}

// Film::to_rgb_array over a batch of hemisphere films: {R,G,B,w} -> RGB (what WriteImage / to_intensity_film read)
__global__ void __launch_bounds__(kMapBlock) hemiResolveKernel(const float4 *__restrict__ film, size_t nPix,
                                                               float filmScale, float *__restrict__ rgb) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nPix; i += (size_t)gridDim.x * blockDim.x) {
        float o[3];
        resolvePixel(film[i], filmScale, o);
        rgb[3 * i] = o[0];
        rgb[3 * i + 1] = o[1];
        rgb[3 * i + 2] = o[2];
    }
}

}  // namespace

cudaError_t hemiResolve(const float4 *film, size_t nPix, float filmScale, float *rgb, cudaStream_t stream) {
    if (nPix == 0) return cudaSuccess;
    hemiResolveKernel<<<148 * 8, kMapBlock, 0, stream>>>(film, nPix, filmScale, rgb);
    return cudaGetLastError();
}

cudaError_t hemiToNN(const float4 *film, const float *normals, const float *distance, size_t n, int size,
                     float filmScale, float *nn, float *means, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const size_t stageBytes = (size_t)size * size * 4 * sizeof(float);
    if (stageBytes <= 48 * 1024)
        hemiToNNKernel<true><<<(unsigned)n, kMapBlock, stageBytes, stream>>>(film, normals, distance, size, filmScale, nn, means);
    else
        hemiToNNKernel<false><<<(unsigned)n, kMapBlock, 0, stream>>>(film, normals, distance, size, filmScale, nn, means);
    return cudaGetLastError();
}

cudaError_t hemiFromNN(const float *nnOut, const float *means, size_t n, int size, float *intensity,
                       cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    hemiFromNNKernel<<<(unsigned)n, kMapBlock, 0, stream>>>(nnOut, means, size, intensity);
    return cudaGetLastError();
}

}  // namespace iile
