// Geometry of one trunk layer seen as a convolution, and the FP32 SIMT engine entry points.
#pragma once
#include "common.cuh"

namespace drl {

struct ConvGeom {
  int H, W, Cin;     // input map (NHWC)
  int OH, OW, Cout;  // output map
  int KH, KW, S;     // kernel and stride (no padding anywhere in NatureCNN / the RND nets)
};

// NatureCNN (drl/agents/architectures/stateless/nature_cnn.py:27-36); layer 3 is the fc layer.
inline ConvGeom nature_layer(int l) {
  switch (l) {
    case 0: return ConvGeom{84, 84, 4, 20, 20, 32, 8, 8, 4};
    case 1: return ConvGeom{20, 20, 32, 9, 9, 64, 4, 4, 2};
    case 2: return ConvGeom{9, 9, 64, 7, 7, 64, 3, 3, 1};
    default: return ConvGeom{7, 7, 64, 1, 1, 512, 7, 7, 1};
  }
}

template <typename TIn>
int simt_conv_forward(const TIn* in, const int64_t* row_idx, int nb, const ConvGeom& g, const float* w,
                      const float* bias, float* out, cudaStream_t st, int act = 1);
int simt_conv_dgrad(const float* dy, int nb, const ConvGeom& g, const float* w, const float* x_act,
                    float* dx, cudaStream_t st, int act = 1);
template <typename TIn>
int simt_conv_wgrad(const float* dy, const TIn* in, const int64_t* row_idx, int nb, const ConvGeom& g,
                    float* gw, float* gb, cudaStream_t st);

// RND first conv (normalise + clip of the last frame channel fused into the gather)
int simt_rnd_conv1_forward(const uint8_t* obs, const int64_t* row_idx, int nb, const float* m1, const float* m2,
                           const ConvGeom& g, const float* w, const float* bias, float* out, cudaStream_t st);
int simt_rnd_conv1_wgrad(const float* dy, const uint8_t* obs, const int64_t* row_idx, int nb, const float* m1,
                         const float* m2, const ConvGeom& g, float* gw, float* gb, cudaStream_t st);

}  // namespace drl
