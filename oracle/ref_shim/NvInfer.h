// TEST INFRASTRUCTURE ONLY -- part of oracle/_ref.  Stand-in for the TensorRT and ONNX-parser interfaces the reference's
// src/tensorrt/*.cpp calls (img2img_build.cpp:68-148, img2img_load.cpp:158-243, img2img_infer.cpp:80), so that
// Img2Img::build / load / render / infer compile unmodified and run on the CPU with the network replaced by a callback:
//   * the "engine" is a one-line text blob `W2XSHIMENGINE <scale> <outMinus>`: output tile = scale * input tile - outMinus
//     (UpCUNet 2x: 2T - 72, CUNet 1x: T - 56, SwinUNet 4x: 4T - 64), written by the fake builder, read by the fake runtime;
//   * IExecutionContext::enqueueV3 hands the bound input tensor [B,3,T,T] f32 to the registered model function and expects
//     [B,3,outT,outT] f32 back (w2xshim::model()).
// Nothing here computes anything the parity tests check: tile grid, padding, weights, augmentation, batching, accumulation
// and packing all execute in the reference's own code.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "w2x_cudashim.hpp"

namespace w2xshim {
typedef void (*ModelFn)(const float* in, int n, int c, int h, int w, float* out, int oh, int ow, void* user);
struct Model {
    int scale = 2, outMinus = 72;
    ModelFn fn = nullptr;
    void* user = nullptr;
    long calls = 0;
};
inline Model& model() {
    static Model m;
    return m;
}
}  // namespace w2xshim

namespace nvinfer1 {

class ILogger {
public:
    enum class Severity : int32_t { kINTERNAL_ERROR = 0, kERROR = 1, kWARNING = 2, kINFO = 3, kVERBOSE = 4 };
    virtual void log(Severity severity, const char* msg) noexcept = 0;
    virtual ~ILogger() = default;
};

struct Dims32 {
    static constexpr int32_t MAX_DIMS = 8;
    int32_t nbDims = 0;
    int32_t d[MAX_DIMS] = {0, 0, 0, 0, 0, 0, 0, 0};
};
typedef Dims32 Dims;
struct Dims4 : public Dims32 {
    Dims4() { nbDims = 4; }
    Dims4(int32_t a, int32_t b, int32_t c, int32_t e) {
        nbDims = 4;
        d[0] = a; d[1] = b; d[2] = c; d[3] = e;
    }
};

enum class NetworkDefinitionCreationFlag : int32_t { kEXPLICIT_BATCH = 0 };
enum class BuilderFlag : int32_t { kFP16 = 0, kTF32 = 7 };
enum class OptProfileSelector : int32_t { kMIN = 0, kOPT = 1, kMAX = 2 };

class IHostMemory {
public:
    std::string blob;
    void* data() const { return (void*)blob.data(); }
    size_t size() const { return blob.size(); }
};

class ITensor {
public:
    const char* getName() const { return "x"; }
    Dims getDimensions() const { return Dims4(-1, 3, -1, -1); }
};

class INetworkDefinition {
public:
    ITensor in;
    bool parsed = false;
    int32_t getNbInputs() const { return 1; }
    ITensor* getInput(int32_t) { return &in; }
};

class IOptimizationProfile {
public:
    Dims dims[3];
    bool setDimensions(const char*, OptProfileSelector s, const Dims& d) {
        dims[(int)s] = d;
        return true;
    }
};

class IBuilderConfig {
public:
    std::vector<IOptimizationProfile*> profiles;
    std::vector<BuilderFlag> flagsSet;
    int32_t addOptimizationProfile(IOptimizationProfile* p) {
        profiles.push_back(p);
        return (int32_t)profiles.size() - 1;
    }
    void setFlag(BuilderFlag f) { flagsSet.push_back(f); }
    void setProfileStream(cudaStream_t) {}
};

class IBuilder {
public:
    std::vector<IOptimizationProfile*> owned;
    ~IBuilder() {
        for (auto* p : owned) delete p;
    }
    INetworkDefinition* createNetworkV2(uint32_t) { return new INetworkDefinition(); }
    IBuilderConfig* createBuilderConfig() { return new IBuilderConfig(); }
    IOptimizationProfile* createOptimizationProfile() {
        owned.push_back(new IOptimizationProfile());
        return owned.back();
    }
    bool platformHasFastFp16() const { return true; }
    bool platformHasTf32() const { return true; }
    IHostMemory* buildSerializedNetwork(INetworkDefinition& n, IBuilderConfig&) {
        if (!n.parsed) return nullptr;
        auto* m = new IHostMemory();
        m->blob = "W2XSHIMENGINE " + std::to_string(w2xshim::model().scale) + " " + std::to_string(w2xshim::model().outMinus) + "\n";
        return m;
    }
};
inline IBuilder* createInferBuilder(ILogger&) { return new IBuilder(); }

class ICudaEngine;

class IExecutionContext {
public:
    int scale = 1, outMinus = 0;
    Dims in, out;
    std::map<std::string, void*> addr;
    bool setInputShape(const char*, const Dims& d) {
        if (d.nbDims != 4 || d.d[0] < 1 || d.d[1] != 3) return false;
        in = d;
        out = Dims4(d.d[0], d.d[1], d.d[2] * scale - outMinus, d.d[3] * scale - outMinus);
        return out.d[2] > 0 && out.d[3] > 0;
    }
    Dims getTensorShape(const char* name) const { return std::strcmp(name, "x") == 0 ? in : out; }
    bool setTensorAddress(const char* name, void* p) {
        addr[name] = p;
        return true;
    }
    bool enqueueV3(cudaStream_t) {
        auto& m = w2xshim::model();
        if (!m.fn || !addr.count("x") || !addr.count("y")) return false;
        ++m.calls;
        m.fn((const float*)addr["x"], in.d[0], in.d[1], in.d[2], in.d[3], (float*)addr["y"], out.d[2], out.d[3], m.user);
        return true;
    }
};

class ICudaEngine {
public:
    int scale = 1, outMinus = 0;
    int32_t getNbIOTensors() const { return 2; }
    const char* getIOTensorName(int32_t i) const { return i == 0 ? "x" : "y"; }
    Dims getTensorShape(const char*) const { return Dims4(-1, 3, -1, -1); }
    IExecutionContext* createExecutionContext() {
        auto* c = new IExecutionContext();
        c->scale = scale;
        c->outMinus = outMinus;
        return c;
    }
};

class IRuntime {
public:
    ICudaEngine* deserializeCudaEngine(const void* data, size_t size) {
        const std::string s((const char*)data, size);
        int sc = 0, om = 0;
        if (std::sscanf(s.c_str(), "W2XSHIMENGINE %d %d", &sc, &om) != 2 || sc < 1) return nullptr;
        auto* e = new ICudaEngine();
        e->scale = sc;
        e->outMinus = om;
        return e;
    }
};
inline IRuntime* createInferRuntime(ILogger&) { return new IRuntime(); }

}  // namespace nvinfer1

namespace nvonnxparser {
class IParser {
public:
    nvinfer1::INetworkDefinition* net;
    explicit IParser(nvinfer1::INetworkDefinition& n) : net(&n) {}
    bool parseFromFile(const char* path, int) {
        std::FILE* f = std::fopen(path, "rb");
        if (!f) return false;
        std::fclose(f);
        net->parsed = true;
        return true;
    }
};
inline IParser* createParser(nvinfer1::INetworkDefinition& n, nvinfer1::ILogger&) { return new IParser(n); }
}  // namespace nvonnxparser
