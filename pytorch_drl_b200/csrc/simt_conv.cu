// FP32 CUDA-core implicit-GEMM convolution engine: the strict-parity mode of the NatureCNN trunk
// (DRL_PRECISION_FP32).  Forward, data-gradient and weight-gradient of every trunk layer
// (nature_cnn.py:27-36; the fc layer is a 7x7 "conv" over the 7x7x64 map so that the NCHW flatten
// order c*49+h*7+w of nature_cnn.py:34 is honoured) run through ONE tiled SIMT GEMM whose
// operands are fetched by gather functors, directly on the reference's parameter layouts (OIHW).
// Activations are NHWC fp32.  conv1 reads uint8 frames and applies x*(1/255) exactly like
// scale_observations.py:35, with the minibatch gather (nested.py:6-12) folded into the row index.
#include "common.cuh"
#include "simt_conv.cuh"

namespace drl {

constexpr int BM = 64, BK = 16;     // BN (64 / 32 / 16) is a template parameter: the RND layers have 16 / 32 columns

// Operand functors are evaluated in two levels so that the per-element cost is two shared-memory reads, an add
// and the load: `d0(i)` describes a row of the operand (computed once per tile for A rows / B columns), `d1(j)` a
// position along the reduction (computed once per 16-wide k-step), `at(d0, d1)` fetches the element.  The index
// arithmetic (divisions by the map geometry) therefore runs once per row / per k instead of once per element.
struct Desc { long long a; int b; int c; };

template <bool A_M_FAST, int BN, class AF, class BF, class EF>
__global__ void __launch_bounds__(256) simt_gemm_kernel(int M, int N, int K, int k_per_split, AF af,
                                                        BF bf, EF ef) {
  constexpr int TN = BN / 16;                        // columns per thread
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ Desc dAm[BM], dBn[BN], dAk[BK], dBk[BK];
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;   // M tiles on x: up to 2^31-1 blocks
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  if (threadIdx.x < BM) {
    if (m0 + (int)threadIdx.x < M) dAm[threadIdx.x] = af.d0(m0 + threadIdx.x);
  } else if (threadIdx.x < BM + BN) {
    const int nn = threadIdx.x - BM;
    if (n0 + nn < N) dBn[nn] = bf.d1(n0 + nn);
  }
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    if (threadIdx.x < BK) {
      if (k0 + (int)threadIdx.x < kend) dAk[threadIdx.x] = af.d1(k0 + threadIdx.x);
    } else if (threadIdx.x >= 32 && threadIdx.x < 32 + BK) {
      const int kk = threadIdx.x - 32;
      if (k0 + kk < kend) dBk[kk] = bf.d0(k0 + kk);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + 256 * i;
      const int kk = A_M_FAST ? e / BM : e % BK;
      const int mm = A_M_FAST ? e % BM : e / BK;
      As[kk][mm] = (m0 + mm < M && k0 + kk < kend) ? af.at(dAm[mm], dAk[kk]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < BK * BN / 256; ++i) {
      const int e = threadIdx.x + 256 * i;
      const int kk = e / BN, nn = e % BN;
      Bs[kk][nn] = (n0 + nn < N && k0 + kk < kend) ? bf.at(dBk[kk], dBn[nn]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[TN];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < N) ef(m, n, acc[i][j]);
    }
  }
}

// ---- gather functors ----------------------------------------------------------------------------
template <typename TIn> __device__ __forceinline__ float load_in(const TIn* p, size_t i);
template <> __device__ __forceinline__ float load_in<float>(const float* p, size_t i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float load_in<uint8_t>(const uint8_t* p, size_t i) {
  return __fmul_rn((float)__ldg(p + i), __uint_as_float(0x3b808081u));   // float32(1/255)
}

// im2col element: row r = (b, oy, ox) of the output grid, column c = (ky, kx, ci)
template <typename TIn>
struct PatchGather {
  const TIn* in; const int64_t* row_idx; ConvGeom g;
  __device__ __forceinline__ Desc d0(int r) const {          // offset of the patch's first element
    const int pos = g.OH * g.OW;
    const int b = r / pos, rem = r - b * pos;
    const int oy = rem / g.OW, ox = rem - oy * g.OW;
    const long long row = row_idx ? (long long)row_idx[b] : (long long)b;
    return Desc{((row * g.H + oy * g.S) * g.W + ox * g.S) * g.Cin, 0, 0};
  }
  __device__ __forceinline__ Desc d1(int c) const {          // offset of tap (ky, kx, ci) inside the patch
    const int kc = g.KW * g.Cin;
    const int ky = c / kc, r2 = c - ky * kc;
    return Desc{0, ky * g.W * g.Cin + r2, 0};
  }
  __device__ __forceinline__ float at(const Desc& r, const Desc& c) const { return load_in<TIn>(in, (size_t)(r.a + c.b)); }
};

struct WeightFwd {            // B(k=(ky,kx,ci), n=co) = W[co, ci, ky, kx]
  const float* w; ConvGeom g;
  __device__ __forceinline__ Desc d0(int k) const {
    const int kc = g.KW * g.Cin;
    const int ky = k / kc, r2 = k - ky * kc;
    const int kx = r2 / g.Cin, ci = r2 - kx * g.Cin;
    return Desc{0, (ci * g.KH + ky) * g.KW + kx, 0};
  }
  __device__ __forceinline__ Desc d1(int n) const { return Desc{(long long)n * g.Cin * g.KH * g.KW, 0, 0}; }
  __device__ __forceinline__ float at(const Desc& k, const Desc& n) const { return __ldg(w + n.a + k.b); }
};

// act: 0 identity, 1 ReLU, 2 LeakyReLU(0.2)
template <typename TOut>
struct EpiBiasRelu {
  const float* bias; TOut* out; int N; int act;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    const float v = acc + __ldg(bias + n);
    out[(size_t)m * N + n] = act == 1 ? fmaxf(v, 0.f) : (act == 2 ? (v > 0.f ? v : 0.2f * v) : v);
  }
};

// dgrad: A(m=(b,iy,ix), k=(ky,kx,co)) = dY[b,(iy-ky)/S,(ix-kx)/S,co] when that lands on the grid
struct DyGather {
  const float* dy; ConvGeom g;
  __device__ __forceinline__ Desc d0(int m) const {
    const int pos = g.H * g.W;
    const int b = m / pos, rem = m - b * pos;
    const int iy = rem / g.W, ix = rem - iy * g.W;
    return Desc{(long long)b * g.OH * g.OW * g.Cout, iy, ix};
  }
  __device__ __forceinline__ Desc d1(int k) const {
    const int kc = g.KW * g.Cout;
    const int ky = k / kc, r2 = k - ky * kc;
    const int kx = r2 / g.Cout, co = r2 - kx * g.Cout;
    return Desc{co, ky, kx};
  }
  __device__ __forceinline__ float at(const Desc& r, const Desc& c) const {
    const int ty = r.b - c.b, tx = r.c - c.c;
    if (ty < 0 || tx < 0) return 0.f;
    const int oy = ty / g.S, ox = tx / g.S;
    if (oy * g.S != ty || ox * g.S != tx || oy >= g.OH || ox >= g.OW) return 0.f;
    return __ldg(dy + r.a + (size_t)(oy * g.OW + ox) * g.Cout + c.a);
  }
};

struct WeightDgrad {          // B(k=(ky,kx,co), n=ci) = W[co, ci, ky, kx]
  const float* w; ConvGeom g;
  __device__ __forceinline__ Desc d0(int k) const {
    const int kc = g.KW * g.Cout;
    const int ky = k / kc, r2 = k - ky * kc;
    const int kx = r2 / g.Cout, co = r2 - kx * g.Cout;
    return Desc{(long long)co * g.Cin * g.KH * g.KW + ky * g.KW + kx, 0, 0};
  }
  __device__ __forceinline__ Desc d1(int n) const { return Desc{0, n * g.KH * g.KW, 0}; }
  __device__ __forceinline__ float at(const Desc& k, const Desc& n) const { return __ldg(w + k.a + n.b); }
};

struct EpiReluMask {          // dX = acc * act'(X): activation of the producing layer (1 ReLU, 2 LeakyReLU 0.2)
  const float* x_act; float* dx; int N; int act;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    const size_t i = (size_t)m * N + n;
    dx[i] = __ldg(x_act + i) > 0.f ? acc : (act == 2 ? 0.2f * acc : 0.f);
  }
};

// ---- strided data gradient by output-parity classes ---------------------------------------------------------
// For stride S > 1 only the taps ky = py + S*a, kx = px + S*b reach input pixel (iy, ix) = (S*u + py, S*v + px), so
// each of the S*S classes is a dense stride-1 correlation over dY with a (KH/S x KW/S) sub-kernel: no tap is
// fetched to be multiplied by zero.  m' = (b, u, v) over the class's Hc x Wc pixels, k' = (a, b', co).
struct DyGatherCls {
  const float* dy; ConvGeom g; int py, px, Hc, Wc;
  __device__ __forceinline__ Desc d0(int m) const {
    const int pos = Hc * Wc;
    const int b = m / pos, rem = m - b * pos;
    const int u = rem / Wc, v = rem - u * Wc;
    return Desc{(long long)b * g.OH * g.OW * g.Cout, u, v};
  }
  __device__ __forceinline__ Desc d1(int k) const {
    const int kc = (g.KW / g.S) * g.Cout;
    const int a = k / kc, r2 = k - a * kc;
    const int bb = r2 / g.Cout, co = r2 - bb * g.Cout;
    return Desc{co, a, bb};
  }
  __device__ __forceinline__ float at(const Desc& r, const Desc& c) const {
    const int oy = r.b - c.b, ox = r.c - c.c;
    if (oy < 0 || ox < 0 || oy >= g.OH || ox >= g.OW) return 0.f;
    return __ldg(dy + r.a + (size_t)(oy * g.OW + ox) * g.Cout + c.a);
  }
};
struct WeightDgradCls {       // B(k'=(a,b',co), n=ci) = W[co, ci, py + S*a, px + S*b']
  const float* w; ConvGeom g; int py, px;
  __device__ __forceinline__ Desc d0(int k) const {
    const int kc = (g.KW / g.S) * g.Cout;
    const int a = k / kc, r2 = k - a * kc;
    const int bb = r2 / g.Cout, co = r2 - bb * g.Cout;
    return Desc{(long long)co * g.Cin * g.KH * g.KW + (py + g.S * a) * g.KW + (px + g.S * bb), 0, 0};
  }
  __device__ __forceinline__ Desc d1(int n) const { return Desc{0, n * g.KH * g.KW, 0}; }
  __device__ __forceinline__ float at(const Desc& k, const Desc& n) const { return __ldg(w + k.a + n.b); }
};
struct EpiReluMaskCls {       // class row m' = (b, u, v) -> input pixel (S*u + py, S*v + px)
  const float* x_act; float* dx; ConvGeom g; int py, px, Hc, Wc, act;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    const int pos = Hc * Wc;
    const int b = m / pos, rem = m - b * pos;
    const int u = rem / Wc, v = rem - u * Wc;
    const size_t i = (((size_t)b * g.H + (g.S * u + py)) * g.W + (g.S * v + px)) * g.Cin + n;
    dx[i] = __ldg(x_act + i) > 0.f ? acc : (act == 2 ? 0.2f * acc : 0.f);
  }
};

struct DyT {                  // wgrad A'(m=co, k=r) = dY[r, co]
  const float* dy; int Cout;
  __device__ __forceinline__ Desc d0(int m) const { return Desc{0, m, 0}; }
  __device__ __forceinline__ Desc d1(int k) const { return Desc{(long long)k * Cout, 0, 0}; }
  __device__ __forceinline__ float at(const Desc& m, const Desc& k) const { return __ldg(dy + k.a + m.b); }
};

template <typename TIn>
struct PatchGatherT {         // wgrad B'(k=r, n=(ky,kx,ci)) plus a ones column (n == Kp) for the bias
  PatchGather<TIn> pg; int Kp;
  __device__ __forceinline__ Desc d0(int k) const { return pg.d0(k); }
  __device__ __forceinline__ Desc d1(int n) const { Desc d = pg.d1(n < Kp ? n : 0); d.c = n < Kp ? 0 : 1; return d; }
  __device__ __forceinline__ float at(const Desc& k, const Desc& n) const { return n.c ? 1.f : pg.at(k, n); }
};

struct EpiWgrad {             // scatter-add into the OIHW gradient (+ bias gradient)
  float* gw; float* gb; ConvGeom g; int Kp;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    if (n < Kp) {
      const int kc = g.KW * g.Cin;
      const int ky = n / kc, r2 = n - ky * kc;
      const int kx = r2 / g.Cin, ci = r2 - kx * g.Cin;
      atomicAdd(gw + ((size_t)(m * g.Cin + ci) * g.KH + ky) * g.KW + kx, acc);
    } else {
      atomicAdd(gb + m, acc);
    }
  }
};

// RND first conv: the LAST channel of the uint8 frame, scaled by 1/255, normalised with the running
// moments and clipped to [-5, 5] (random_network_distillation.py:228-229; normalize.py:149-162),
// fused into the im2col gather.  Cin = 1.
struct NormPatchGather {
  const uint8_t* obs; const int64_t* row_idx; const float* m1; const float* m2; ConvGeom g;
  __device__ __forceinline__ Desc d0(int r) const {          // frame offset and pixel index of the patch origin
    const int pos = g.OH * g.OW;
    const int b = r / pos, rem = r - b * pos;
    const int oy = rem / g.OW, ox = rem - oy * g.OW;
    const long long row = row_idx ? (long long)row_idx[b] : (long long)b;
    const int pix = (oy * g.S) * 84 + ox * g.S;
    return Desc{(row * 7056 + pix) * 4 + 3, pix, 0};
  }
  __device__ __forceinline__ Desc d1(int c) const {
    const int ky = c / g.KW, kx = c - ky * g.KW;
    return Desc{0, ky * 84 + kx, 0};
  }
  __device__ __forceinline__ float at(const Desc& r, const Desc& c) const {
    const int pix = r.b + c.b;
    const float v = __fmul_rn((float)__ldg(obs + r.a + (size_t)c.b * 4), __uint_as_float(0x3b808081u));
    const float mu = __ldg(m1 + pix);
    const float var = __fsub_rn(__ldg(m2 + pix), __fmul_rn(mu, mu));
    const float z = __fmul_rn(__fsub_rn(v, mu), rsqrtf(__fadd_rn(var, 1e-8f)));
    return fminf(fmaxf(z, -5.f), 5.f);
  }
};
struct NormPatchGatherT {
  NormPatchGather pg; int Kp;
  __device__ __forceinline__ Desc d0(int k) const { return pg.d0(k); }
  __device__ __forceinline__ Desc d1(int n) const { Desc d = pg.d1(n < Kp ? n : 0); d.c = n < Kp ? 0 : 1; return d; }
  __device__ __forceinline__ float at(const Desc& k, const Desc& n) const { return n.c ? 1.f : pg.at(k, n); }
};

static inline int pick_bn(int N) { return N <= 16 ? 16 : N <= 32 ? 32 : 64; }
static inline dim3 tiles(int M, int N, int z = 1) {
  const int bn = pick_bn(N);
  return dim3((M + BM - 1) / BM, (N + bn - 1) / bn, z);
}
// launch with the column-tile width that fits N
#define SIMT_LAUNCH(A_M_FAST, grid, st, M, N, K, kper, af, bf, ef)                                              \
  do {                                                                                                          \
    const int bn_ = pick_bn(N);                                                                                 \
    if (bn_ == 16) simt_gemm_kernel<A_M_FAST, 16><<<grid, 256, 0, st>>>(M, N, K, kper, af, bf, ef);             \
    else if (bn_ == 32) simt_gemm_kernel<A_M_FAST, 32><<<grid, 256, 0, st>>>(M, N, K, kper, af, bf, ef);        \
    else simt_gemm_kernel<A_M_FAST, 64><<<grid, 256, 0, st>>>(M, N, K, kper, af, bf, ef);                       \
  } while (0)

template <typename TIn>
int simt_conv_forward(const TIn* in, const int64_t* row_idx, int nb, const ConvGeom& g, const float* w,
                      const float* bias, float* out, cudaStream_t st, int act) {
  const int M = nb * g.OH * g.OW, N = g.Cout, K = g.KH * g.KW * g.Cin;
  PatchGather<TIn> af{in, row_idx, g};
  WeightFwd bf{w, g};
  EpiBiasRelu<float> ef{bias, out, N, act};
  ProfileScope prof("simt_fwd", st);
  SIMT_LAUNCH(false, tiles(M, N), st, M, N, K, K, af, bf, ef);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}
template int simt_conv_forward<float>(const float*, const int64_t*, int, const ConvGeom&, const float*,
                                      const float*, float*, cudaStream_t, int);
template int simt_conv_forward<uint8_t>(const uint8_t*, const int64_t*, int, const ConvGeom&,
                                        const float*, const float*, float*, cudaStream_t, int);

int simt_conv_dgrad(const float* dy, int nb, const ConvGeom& g, const float* w, const float* x_act,
                    float* dx, cudaStream_t st, int act) {
  const int N = g.Cin;
  ProfileScope prof("simt_dgrad", st);
  if (g.S > 1 && g.KH % g.S == 0 && g.KW % g.S == 0) {
    // one dense stride-1 problem per output-parity class; pixels no tap reaches (none in NatureCNN / RND) keep the
    // zero the class sum would give them because every input pixel belongs to exactly one class
    for (int py = 0; py < g.S; ++py)
      for (int px = 0; px < g.S; ++px) {
        const int Hc = (g.H - py + g.S - 1) / g.S, Wc = (g.W - px + g.S - 1) / g.S;
        const int M = nb * Hc * Wc, K = (g.KH / g.S) * (g.KW / g.S) * g.Cout;
        DyGatherCls af{dy, g, py, px, Hc, Wc};
        WeightDgradCls bf{w, g, py, px};
        EpiReluMaskCls ef{x_act, dx, g, py, px, Hc, Wc, act};
        SIMT_LAUNCH(false, tiles(M, N), st, M, N, K, K, af, bf, ef);
        DRL_LAUNCH_CHECK();
      }
    return DRL_OK;
  }
  const int M = nb * g.H * g.W, K = g.KH * g.KW * g.Cout;
  DyGather af{dy, g};
  WeightDgrad bf{w, g};
  EpiReluMask ef{x_act, dx, N, act};
  SIMT_LAUNCH(false, tiles(M, N), st, M, N, K, K, af, bf, ef);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

template <typename TIn>
int simt_conv_wgrad(const float* dy, const TIn* in, const int64_t* row_idx, int nb, const ConvGeom& g,
                    float* gw, float* gb, cudaStream_t st) {
  const int Kp = g.KH * g.KW * g.Cin;
  const int M = g.Cout, N = Kp + 1, K = nb * g.OH * g.OW;
  // split the (long) reduction over the rows so that the grid fills the machine
  const int base = tiles(M, N).x * tiles(M, N).y;
  int splits = (sm_count() * 4 + base - 1) / base;
  int k_per = (K + splits - 1) / splits;
  k_per = ((k_per + BK - 1) / BK) * BK;
  splits = (K + k_per - 1) / k_per;
  DyT af{dy, g.Cout};
  PatchGatherT<TIn> bf{PatchGather<TIn>{in, row_idx, g}, Kp};
  EpiWgrad ef{gw, gb, g, Kp};
  ProfileScope prof("simt_wgrad", st);
  SIMT_LAUNCH(true, tiles(M, N, splits), st, M, N, K, k_per, af, bf, ef);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}
template int simt_conv_wgrad<float>(const float*, const float*, const int64_t*, int, const ConvGeom&,
                                    float*, float*, cudaStream_t);
template int simt_conv_wgrad<uint8_t>(const float*, const uint8_t*, const int64_t*, int,
                                      const ConvGeom&, float*, float*, cudaStream_t);

}  // namespace drl

namespace drl {

int simt_rnd_conv1_forward(const uint8_t* obs, const int64_t* row_idx, int nb, const float* m1, const float* m2,
                           const ConvGeom& g, const float* w, const float* bias, float* out, cudaStream_t st) {
  const int M = nb * g.OH * g.OW, N = g.Cout, K = g.KH * g.KW;
  NormPatchGather af{obs, row_idx, m1, m2, g};
  WeightFwd bf{w, g};
  EpiBiasRelu<float> ef{bias, out, N, 2};
  ProfileScope prof("rnd_simt_fwd", st);
  SIMT_LAUNCH(false, tiles(M, N), st, M, N, K, K, af, bf, ef);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

int simt_rnd_conv1_wgrad(const float* dy, const uint8_t* obs, const int64_t* row_idx, int nb, const float* m1,
                         const float* m2, const ConvGeom& g, float* gw, float* gb, cudaStream_t st) {
  const int Kp = g.KH * g.KW;
  const int M = g.Cout, N = Kp + 1, K = nb * g.OH * g.OW;
  const int base = tiles(M, N).x * tiles(M, N).y;
  int splits = (sm_count() * 4 + base - 1) / base;
  int k_per = (K + splits - 1) / splits;
  k_per = ((k_per + BK - 1) / BK) * BK;
  splits = (K + k_per - 1) / k_per;
  DyT af{dy, g.Cout};
  NormPatchGatherT bf{NormPatchGather{obs, row_idx, m1, m2, g}, Kp};
  EpiWgrad ef{gw, gb, g, Kp};
  ProfileScope prof("rnd_simt_wgrad", st);
  SIMT_LAUNCH(true, tiles(M, N, splits), st, M, N, K, k_per, af, bf, ef);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

}  // namespace drl
