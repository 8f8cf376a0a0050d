// LPIPS-AlexNet scorer (see lpips.cu).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>

namespace ddpm {

class Lpips {
   public:
    Lpips();
    ~Lpips();
    int init();
    // names follow lpips.LPIPS.state_dict(): net.slice{1..5}.{0,3,6,8,10}.{weight,bias}, lin{0..4}.model.1.weight
    int set_param(const char* name, const float* data, long long numel, cudaStream_t stream);
    int finalize();
    size_t workspace_bytes(int B, int H, int W) const;
    // in0, in1: fp32 [B, C, H, W] with C in {1, 3}; out: fp32 [B]
    int forward(const float* in0, const float* in1, float* out, int B, int C, int H, int W, bool normalize, void* ws,
                size_t ws_bytes, cudaStream_t stream);
    long long launches() const { return launches_; }

   private:
    struct Slot { float* dst; long long numel; bool set; };
    float* arena_ = nullptr;
    float *w_[5], *b_[5], *lin_[5];
    // tensor-core path of the three 3x3 convs (large maps: 3-D volumes scored slice by slice): split-precision weights
    // [Cout][hi | hi | lo] fp16 (see lpips.cu), packed on the first forward that needs them
    void* wsplit_ = nullptr;
    void* wsplit_k_[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float* bias1_pad_ = nullptr;  // conv2's bias padded from 192 to 256 output channels (two 128-wide N tiles)
    bool wsplit_ready_ = false;
    std::map<std::string, Slot> slots_;
    bool ready_ = false;
    long long launches_ = 0;
};

}  // namespace ddpm
