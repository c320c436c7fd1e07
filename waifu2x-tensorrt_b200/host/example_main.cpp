// Minimal C++ consumer of the shim, mirroring the render branch of /root/reference/src/main.cpp:211-274 without the
// ffmpeg/CLI11/spdlog plumbing: build (if needed), load, render one synthetic frame.  Build:
//   g++ -std=c++17 example_main.cpp -I../../include -L../lib -lw2x -Wl,-rpath,'$ORIGIN/../lib' -o example_main
#include <cstdio>
#include <vector>

#include "img2img.hpp"

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s models/cunet/art/noise3_scale2x.onnx [tile=256] [batch=8]\n", argv[0]); return 2; }
    const int tile = argc > 2 ? std::atoi(argv[2]) : 256, batch = argc > 3 ? std::atoi(argv[3]) : 8, scale = 2;
    trt::Img2Img engine;
    engine.setMessageCallback([](trt::Severity s, const std::string& m) { std::fprintf(stderr, "[%d] %s\n", (int)s, m.c_str()); });
    engine.setProgressCallback([](int cur, int total, double its) { std::fprintf(stderr, "batch %d/%d @ %.2f it/s\n", cur, total, its); });
    trt::BuildConfig bc;
    bc.minBatchSize = bc.optBatchSize = bc.maxBatchSize = batch;
    bc.minWidth = bc.optWidth = bc.maxWidth = bc.minHeight = bc.optHeight = bc.maxHeight = tile;
    trt::RenderConfig rc;
    rc.batchSize = batch; rc.height = rc.width = tile; rc.scaling = scale;
    if (!engine.load(argv[1], rc)) {              // no engine file yet: build it like `waifu2x-tensorrt build`
        if (!engine.build(argv[1], bc) || !engine.load(argv[1], rc)) return -1;
    }
    const int w = 640, h = 360;
    std::vector<unsigned char> src((size_t)w * h * 3), dst((size_t)w * scale * h * scale * 3);
    for (size_t i = 0; i < src.size(); ++i) src[i] = (unsigned char)((i * 2654435761u) >> 24);
    if (!engine.render(src.data(), w, h, (size_t)w * 3, dst.data(), (size_t)w * scale * 3)) return -1;
    std::printf("rendered %dx%d -> %dx%d, first pixel BGR = %d %d %d\n", w, h, w * scale, h * scale, dst[0], dst[1], dst[2]);
    return 0;
}
