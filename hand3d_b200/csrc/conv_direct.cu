// fp32 CUDA-core kernels: generic NHWC direct convolution (tf.nn.conv2d 'SAME' + bias + leaky ReLU,
// utils/general.py:36-59) used for the layers tensor cores cannot help (Cin = 3 first layers, K = 27;
// Cout = 2 / 21 score-map heads; the tiny stride-2 lifting pyramids), the fully connected layers
// (utils/general.py:113-136) and as the fp32 yard-stick path (H3D_PREC_FP32_FFMA).
#include "common.cuh"

namespace h3d {

namespace {

constexpr int TM = 64;   // output pixels per CTA
constexpr int TN = 64;   // output channels per CTA
constexpr int KC = 16;   // reduction chunk

template <bool FP16>
__device__ __forceinline__ uint16_t to_h16(float v) {
    if (FP16) return __half_as_ushort(__float2half_rn(v));
    return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
template <bool FP16>
__device__ __forceinline__ float from_h16(uint16_t v) {
    if (FP16) return __half2float(__ushort_as_half(v));
    return __uint_as_float((uint32_t)v << 16);
}

struct ConvGeom {
    int B, H, W, Ho, Wo, Cin, Cout, k, stride, pad_t, pad_l;
    int Cin_total, cin_off, Cout_total, cout_off, Cs_total, cs_off;
    int leaky;
};

// VEC: Cin % 16 == 0 and 16-byte aligned input channels -> one tap per K chunk, float4 gathers.
template <bool VEC, bool FP16>
__global__ void __launch_bounds__(256)
conv_direct_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ y, uint16_t* __restrict__ yhi, uint16_t* __restrict__ ylo, ConvGeom g) {
    __shared__ __align__(16) float As[KC][TM + 4];
    __shared__ __align__(16) float Bs[KC][TN + 4];
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int64_t M = (int64_t)g.B * g.Ho * g.Wo;
    const int64_t m0 = (int64_t)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    const int Ktot = g.k * g.k * g.Cin;

    // the pixel this thread gathers for (fixed across the K loop)
    const int a_pix = VEC ? (t >> 2) : (t & 63);
    const int64_t am = m0 + a_pix;
    const bool a_valid = am < M;
    int ab = 0, aoy = 0, aox = 0;
    if (a_valid) {
        aox = (int)(am % g.Wo);
        aoy = (int)((am / g.Wo) % g.Ho);
        ab = (int)(am / ((int64_t)g.Wo * g.Ho));
    }
    const int iy0 = aoy * g.stride - g.pad_t, ix0 = aox * g.stride - g.pad_l;
    const float* xb = x + (int64_t)ab * g.H * g.W * g.Cin_total + g.cin_off;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int kk0 = 0; kk0 < Ktot; kk0 += KC) {
        // ---- gather A chunk
        if (VEC) {
            const int tap = kk0 / g.Cin, c0 = kk0 - tap * g.Cin;
            const int kh = tap / g.k, kw = tap - kh * g.k;
            const int iy = iy0 + kh, ix = ix0 + kw;
            const int cs = (t & 3) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_valid && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
                v = __ldg(reinterpret_cast<const float4*>(xb + ((int64_t)iy * g.W + ix) * g.Cin_total + c0 + cs));
            As[cs + 0][a_pix] = v.x; As[cs + 1][a_pix] = v.y; As[cs + 2][a_pix] = v.z; As[cs + 3][a_pix] = v.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int ks = (t >> 6) + q * 4;
                const int kk = kk0 + ks;
                float v = 0.f;
                if (a_valid && kk < Ktot) {
                    const int tap = kk / g.Cin, ci = kk - tap * g.Cin;
                    const int kh = tap / g.k, kw = tap - kh * g.k;
                    const int iy = iy0 + kh, ix = ix0 + kw;
                    if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) v = __ldg(xb + ((int64_t)iy * g.W + ix) * g.Cin_total + ci);
                }
                As[ks][a_pix] = v;
            }
        }
        // ---- load B chunk (HWIO: row kk, Cout contiguous)
        {
            const int r = t >> 4, c4 = (t & 15) * 4;
            const int kk = kk0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kk < Ktot) {
                const float* wr = w + (int64_t)kk * g.Cout + n0 + c4;
                if ((g.Cout & 3) == 0 && n0 + c4 + 3 < g.Cout) {
                    v = __ldg(reinterpret_cast<const float4*>(wr));
                } else {
                    if (n0 + c4 + 0 < g.Cout) v.x = __ldg(wr + 0);
                    if (n0 + c4 + 1 < g.Cout) v.y = __ldg(wr + 1);
                    if (n0 + c4 + 2 < g.Cout) v.z = __ldg(wr + 2);
                    if (n0 + c4 + 3 < g.Cout) v.w = __ldg(wr + 3);
                }
            }
            *reinterpret_cast<float4*>(&Bs[r][c4]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][tx * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][ty * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue: bias + leaky ReLU, fp32 and / or split store
    const int nb = n0 + ty * 4;
    float bvals[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bvals[j] = (nb + j < g.Cout) ? __ldg(bias + nb + j) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + tx * 4 + i;
        if (m >= M) continue;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = acc[i][j] + bvals[j];
            if (g.leaky) v = fmaxf(v, kNegSlope * v);
            o[j] = v;
        }
        if (y) {
            float* dst = y + m * g.Cout_total + g.cout_off + nb;
            if (nb + 3 < g.Cout && ((g.Cout_total | g.cout_off) & 3) == 0) {
                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (nb + j < g.Cout) dst[j] = o[j];
            }
        }
        if (yhi) {
            const int64_t off = m * g.Cs_total + g.cs_off + nb;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (nb + j < g.Cout) {
                    const uint16_t h = to_h16<FP16>(o[j]);
                    yhi[off + j] = h;
                    if (ylo) ylo[off + j] = to_h16<FP16>(o[j] - from_h16<FP16>(h));
                }
            }
        }
    }
}

}  // namespace

int launch_conv_direct(const DirectConvArgs& a, cudaStream_t s) {
    H3D_REQUIRE(a.k >= 1 && a.stride >= 1 && a.Cin >= 1 && a.Cout >= 1, "conv_direct: bad geometry");
    ConvGeom g;
    g.B = a.B; g.H = a.H; g.W = a.W; g.Cin = a.Cin; g.Cout = a.Cout; g.k = a.k; g.stride = a.stride;
    g.Ho = ceil_div(a.H, a.stride); g.Wo = ceil_div(a.W, a.stride);
    const int tot_h = std::max((g.Ho - 1) * a.stride + a.k - a.H, 0), tot_w = std::max((g.Wo - 1) * a.stride + a.k - a.W, 0);
    g.pad_t = tot_h / 2; g.pad_l = tot_w / 2;   // TF 'SAME': the odd pixel goes to the bottom / right
    g.Cin_total = a.Cin_total; g.cin_off = a.cin_off; g.Cout_total = a.Cout_total; g.cout_off = a.cout_off;
    g.Cs_total = a.Cs_total; g.cs_off = a.cs_off; g.leaky = a.leaky;
    const int64_t M = (int64_t)g.B * g.Ho * g.Wo;
    dim3 grid((unsigned)ceil_div64(M, TM), (unsigned)ceil_div(a.Cout, TN));
    const bool vec = (a.Cin % KC == 0) && (a.Cin_total % 4 == 0) && (a.cin_off % 4 == 0) && (((uintptr_t)a.x & 15) == 0);
    const bool fp16 = a.half == Half16::FP16;
#define LAUNCH(V, F) conv_direct_kernel<V, F><<<grid, 256, 0, s>>>(a.x, a.w, a.bias, a.y, a.ys.hi, a.ys.lo, g)
    if (vec) { if (fp16) LAUNCH(true, true); else LAUNCH(true, false); }
    else     { if (fp16) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
    H3D_CHECK_LAUNCH();
    return H3D_OK;
}

// ---------------------------------------------------------------------------------------------
// NetworkOps.fully_connected(_relu): y = x @ W[in,out] + b (+ leaky).  16 batch rows x 64 outputs per
// CTA; W is streamed once per batch tile with coalesced reads, x is staged in shared memory.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fc_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
          int B, int in_f, int out_f, int leaky, int x_stride) {
    constexpr int FB = 16, FK = 64;
    __shared__ float xs[FB][FK];
    const int t = threadIdx.x;
    const int n = blockIdx.x * 64 + (t & 63);
    const int rb = (t >> 6) * 4;          // 4 batch rows per thread
    const int b0 = blockIdx.y * FB;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < in_f; k0 += FK) {
        for (int i = t; i < FB * FK; i += 256) {
            const int r = i / FK, c = i - r * FK;
            xs[r][c] = (b0 + r < B && k0 + c < in_f) ? x[(int64_t)(b0 + r) * x_stride + k0 + c] : 0.f;
        }
        __syncthreads();
        if (n < out_f) {
            const int kmax = min(FK, in_f - k0);
            for (int k = 0; k < kmax; ++k) {
                const float wv = __ldg(w + (int64_t)(k0 + k) * out_f + n);
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] = fmaf(xs[rb + r][k], wv, acc[r]);
            }
        }
        __syncthreads();
    }
    if (n < out_f) {
        const float bv = __ldg(bias + n);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int b = b0 + rb + r;
            if (b < B) {
                float v = acc[r] + bv;
                if (leaky) v = fmaxf(v, kNegSlope * v);
                y[(int64_t)b * out_f + n] = v;
            }
        }
    }
}

int launch_fc(const float* x, const float* w, const float* bias, float* y, int B, int in_f, int out_f, int leaky, int x_stride,
              cudaStream_t s) {
    dim3 grid(ceil_div(out_f, 64), ceil_div(B, 16));
    fc_kernel<<<grid, 256, 0, s>>>(x, w, bias, y, B, in_f, out_f, leaky, x_stride);
    H3D_CHECK_LAUNCH();
    return H3D_OK;
}

// tf.reshape([B,-1]) of the NHWC feature map + tf.concat([x, hand_side], 1)  (nets/...:262-263,297-298)
__global__ void concat_handside_kernel(const float* __restrict__ feat, const float* __restrict__ hs, float* __restrict__ out,
                                       int B, int n) {
    const int64_t total = (int64_t)B * (n + 2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % (n + 2));
        const int b = (int)(i / (n + 2));
        out[i] = c < n ? feat[(int64_t)b * n + c] : hs[2 * b + (c - n)];
    }
}
int launch_concat_handside(const float* feat, const float* hand_side, float* out, int B, int feat_n, cudaStream_t s) {
    const int64_t total = (int64_t)B * (feat_n + 2);
    concat_handside_kernel<<<(int)std::min<int64_t>(ceil_div64(total, 256), 1024), 256, 0, s>>>(feat, hand_side, out, B, feat_n);
    H3D_CHECK_LAUNCH();
    return H3D_OK;
}

}  // namespace h3d
