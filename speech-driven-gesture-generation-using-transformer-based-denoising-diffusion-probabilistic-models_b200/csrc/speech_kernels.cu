// Row kernels of the once-per-clip speech encoder (ResNetSE-34 trunk of HA2GSpeechEncoder) around the tensor-core
// convolutions (gemm_tcgen05.cu, MODE_CONV): the 1-channel stem, the squeeze-excite gate, the block tail and the pixel
// shuffles of the pyramid heads.  Feature maps are channel-last bf16 pixel rows on zero-bordered grids (gd_b200.h);
// every kernel here is HBM/L2-bound, 16-byte vectorised along the channels, and none of them writes a border pixel.
#include "common.cuh"
#include "host_util.h"

namespace gd {

static inline int grid_for(size_t total, int block) { return (int)((total + block - 1) / block); }

__device__ __forceinline__ void unpack_bf16x8(const uint4& q, float (&f)[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// A feature-map row holds `c` channels as one bf16 plane (split = 0) or as two planes [hi(c) | lo(c)], value = hi + lo.
__device__ __forceinline__ void load_ch8(const __nv_bfloat16* row, int c, int g, int split, float (&f)[8]) {
    unpack_bf16x8(__ldcg(reinterpret_cast<const uint4*>(row) + g), f);  // activations: L2-coherent (common.cuh)
    if (split) {
        float l[8];
        unpack_bf16x8(__ldcg(reinterpret_cast<const uint4*>(row + c) + g), l);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += l[j];
    }
}
__device__ __forceinline__ void store_ch8(__nv_bfloat16* row, int c, int g, int split, const float (&v)[8]) {
    const uint4 hi = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    reinterpret_cast<uint4*>(row)[g] = hi;
    if (split) {
        float h[8];
        unpack_bf16x8(hi, h);
        reinterpret_cast<uint4*>(row + c)[g] = make_uint4(pack_bf16x2(v[0] - h[0], v[1] - h[1]), pack_bf16x2(v[2] - h[2], v[3] - h[3]),
                                                          pack_bf16x2(v[4] - h[4], v[5] - h[5]), pack_bf16x2(v[6] - h[6], v[7] - h[7]));
    }
}

// ------------------------------------------------------------------------------------------ stem
// conv1 (1 -> c_real, 3x3, zero padding, bias) + ReLU + folded BatchNorm.  One thread = one pixel x 8 channels.
__global__ void __launch_bounds__(256) speech_stem_kernel(const float* __restrict__ mel, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, __nv_bfloat16* __restrict__ out,
                                                          int n_images, int H, int W, int c_real, int c_pad, int split) {
    extern __shared__ float s_par[];  // [c_real*9 | c_real | c_real | c_real]
    pdl_launch_dependents();
    pdl_wait();
    float* s_w = s_par;
    float* s_b = s_w + c_real * 9;
    float* s_sc = s_b + c_real;
    float* s_sh = s_sc + c_real;
    for (int i = threadIdx.x; i < c_real * 9; i += blockDim.x) s_w[i] = w[i];
    for (int i = threadIdx.x; i < c_real; i += blockDim.x) s_b[i] = bias[i], s_sc[i] = scale[i], s_sh[i] = shift[i];
    __syncthreads();
    const int groups = c_pad >> 3;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t pix = idx / groups;
    const int g = (int)(idx - pix * groups);
    if (pix >= (size_t)n_images * H * W) return;
    const int img = (int)(pix / ((size_t)H * W));
    const int rem = (int)(pix - (size_t)img * H * W);
    const int y = rem / W, x = rem - y * W;
    float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (g * 8 < c_real) {
        float m[9];
        const float* base = mel + (size_t)img * H * W;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int yy = y + ky - 1, xx = x + kx - 1;
                m[ky * 3 + kx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldcg(base + yy * W + xx) : 0.f;
            }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g * 8 + j;
            float acc = 0.f;
            if (c < c_real) {
#pragma unroll
                for (int t = 0; t < 9; ++t) acc = fmaf(s_w[c * 9 + t], m[t], acc);
                acc = fmaxf(acc + s_b[c], 0.f) * s_sc[c] + s_sh[c];
            }
            r[j] = acc;
        }
    }
    const size_t orow = (size_t)img * (H + 2) * (W + 2) + (size_t)(y + 1) * (W + 2) + (x + 1);
    store_ch8(out + orow * (split ? 2 * c_pad : c_pad), c_pad, g, split, r);
}

// ------------------------------------------------------------------------------------------ squeeze-excite gate
// grid (slices, images): a CTA sums the channels of SE_SLICE pixel rows of one image (the border is zero) in a fixed
// order and writes its partial; the last CTA of an image to finish adds the partials in slice order and applies the two
// tiny Linear layers.  The slicing depends on the image size only, so a clip's gate does not depend on its batch.
constexpr int SE_SLICE = 256;

__global__ void __launch_bounds__(256) se_gate_kernel(const __nv_bfloat16* __restrict__ y, int grid_px, int interior_px, int c,
                                                      int split, int c_real, int c_hidden, const float* __restrict__ w1,
                                                      const float* __restrict__ b1, const float* __restrict__ w2,
                                                      const float* __restrict__ b2, float* __restrict__ gate,
                                                      float* __restrict__ partial, int* __restrict__ counters) {
    __shared__ float s_part[256 * 8];
    __shared__ float s_mean[256];
    __shared__ float s_hid[32];
    __shared__ int s_last;
    pdl_launch_dependents();
    pdl_wait();
    const int img = blockIdx.y, slice = blockIdx.x, slices = gridDim.x;
    const int groups = c >> 3;            // 8-channel groups per pixel row
    const int lanes = 256 / groups;       // pixel rows read concurrently
    const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int ld = split ? 2 * c : c;
    const __nv_bfloat16* base = y + (size_t)img * grid_px * ld;
    const int p_end = min(grid_px, (slice + 1) * SE_SLICE);
    if (pl < lanes) {
        for (int p = slice * SE_SLICE + pl; p < p_end; p += lanes) {
            float f[8];
            load_ch8(base + (size_t)p * ld, c, g, split, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_part[threadIdx.x * 8 + j] = acc[j];
    __syncthreads();
    float* my_partials = partial + (size_t)img * slices * c;
    if (threadIdx.x < c) {
        const int ch = threadIdx.x, cg = ch >> 3, cj = ch & 7;
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += s_part[(l * groups + cg) * 8 + cj];
        my_partials[slice * c + ch] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(&counters[img], 1);
        s_last = (done == slices - 1);
        if (s_last) counters[img] = 0;  // ready for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < c) {
        float s = 0.f;
        for (int sl = 0; sl < slices; ++sl) s += __ldcg(my_partials + sl * c + threadIdx.x);
        s_mean[threadIdx.x] = s / (float)interior_px;
    }
    __syncthreads();
    if (threadIdx.x < c_hidden) {
        float h = b1[threadIdx.x];
        for (int i = 0; i < c_real; ++i) h = fmaf(w1[threadIdx.x * c_real + i], s_mean[i], h);
        s_hid[threadIdx.x] = fmaxf(h, 0.f);
    }
    __syncthreads();
    if (threadIdx.x < c) {
        float gv = 0.f;
        if (threadIdx.x < c_real) {
            float a = b2[threadIdx.x];
            for (int j = 0; j < c_hidden; ++j) a = fmaf(w2[threadIdx.x * c_hidden + j], s_hid[j], a);
            gv = 1.0f / (1.0f + expf(-a));
        }
        gate[(size_t)img * c + threadIdx.x] = gv;
    }
}

// ------------------------------------------------------------------------------------------ block tail
// out = relu(gate * y + residual) on the interior pixels.  One thread = one pixel x 8 channels.
__global__ void __launch_bounds__(256) se_residual_relu_kernel(const __nv_bfloat16* __restrict__ y,
                                                               const __nv_bfloat16* __restrict__ res,
                                                               const float* __restrict__ gate, __nv_bfloat16* __restrict__ out,
                                                               int n_images, int grid_h, int grid_w, int c, int split) {
    pdl_launch_dependents();
    pdl_wait();
    const int groups = c >> 3;
    const int ld = split ? 2 * c : c;
    const int H = grid_h - 2, W = grid_w - 2;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t pix = idx / groups;
    const int g = (int)(idx - pix * groups);
    if (pix >= (size_t)n_images * H * W) return;
    const int img = (int)(pix / ((size_t)H * W));
    const int rem = (int)(pix - (size_t)img * H * W);
    const int py = rem / W, px = rem - py * W;
    const size_t row = (size_t)img * grid_h * grid_w + (size_t)(py + 1) * grid_w + (px + 1);
    float a[8], r[8];
    load_ch8(y + row * ld, c, g, split, a);
    load_ch8(res + row * ld, c, g, split, r);
    const float4 g0 = __ldcg(reinterpret_cast<const float4*>(gate + (size_t)img * c) + 2 * g);
    const float4 g1 = __ldcg(reinterpret_cast<const float4*>(gate + (size_t)img * c) + 2 * g + 1);
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaf(gv[j], a[j], r[j]), 0.f);
    store_ch8(out + row * ld, c, g, split, o);
}

// ------------------------------------------------------------------------------------------ pixel shuffle
// out[img, y*r+i, x*r+j, ch] = in[img, y+1, x+1, ch*r*r + i*r + j]; ch >= c_in/r^2 written as 0.  Split rows: both planes.
__global__ void __launch_bounds__(256) pixel_shuffle_rows_kernel(const __nv_bfloat16* __restrict__ in,
                                                                 __nv_bfloat16* __restrict__ out, int n_images, int H, int W,
                                                                 int c_in, int r, int c_out_pad, int split) {
    pdl_launch_dependents();
    pdl_wait();
    const int groups = c_out_pad >> 3;
    const int planes = split ? 2 : 1;
    const int Ho = H * r, Wo = W * r, c_real = c_in / (r * r);
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t pix = idx / groups;
    const int g = (int)(idx - pix * groups);
    if (pix >= (size_t)n_images * Ho * Wo) return;
    const int img = (int)(pix / ((size_t)Ho * Wo));
    const int rem = (int)(pix - (size_t)img * Ho * Wo);
    const int oy = rem / Wo, ox = rem - oy * Wo;
    const int y = oy / r, i = oy - y * r, x = ox / r, j = ox - x * r;
    for (int pl = 0; pl < planes; ++pl) {
        const unsigned short* src = reinterpret_cast<const unsigned short*>(in) +
                                    ((size_t)img * (H + 2) * (W + 2) + (size_t)(y + 1) * (W + 2) + (x + 1)) * (planes * c_in) +
                                    pl * c_in + i * r + j;
        unsigned short v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ch = g * 8 + k;
            v[k] = ch < c_real ? __ldcg(src + ch * r * r) : (unsigned short)0;
        }
        uint4 o;
        o.x = v[0] | ((uint32_t)v[1] << 16), o.y = v[2] | ((uint32_t)v[3] << 16);
        o.z = v[4] | ((uint32_t)v[5] << 16), o.w = v[6] | ((uint32_t)v[7] << 16);
        reinterpret_cast<uint4*>(out + pix * (planes * c_out_pad) + pl * c_out_pad)[g] = o;
    }
}

}  // namespace gd

using namespace gd;

extern "C" int gd_speech_stem(const float* mel, const float* w, const float* bias, const float* scale, const float* shift,
                              void* out_bf16, int32_t n_images, int32_t H, int32_t W, int32_t c_real, int32_t c_pad,
                              int32_t split, void* stream) {
    if (!mel || !w || !bias || !scale || !shift || !out_bf16) return set_error(GD_ERR_INVALID, "gd_speech_stem: null pointer");
    if (n_images <= 0 || H <= 0 || W <= 0 || c_real <= 0 || c_real > c_pad || c_pad % 8 || c_real > 256)
        return set_error(GD_ERR_INVALID, "gd_speech_stem: bad shape");
    const size_t total = (size_t)n_images * H * W * (c_pad / 8);
    GD_CUDA_CHECK(launch_k(speech_stem_kernel, grid_for(total, 256), 256, (size_t)c_real * 12 * sizeof(float),
                           reinterpret_cast<cudaStream_t>(stream), 1, mel, w, bias, scale, shift,
                           reinterpret_cast<__nv_bfloat16*>(out_bf16), n_images, H, W, c_real, c_pad, split));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

// scratch = [SE_MAX_IMAGES arrival counters | per-image partial sums].  The counter block has a fixed size so that calls with
// different image counts on the same scratch agree on where the (self-resetting) counters live.
constexpr int SE_MAX_IMAGES = 65536;

extern "C" int64_t gd_se_gate_scratch_bytes(int32_t n_images, int32_t grid_h, int32_t grid_w, int32_t c) {
    const int64_t slices = ((int64_t)grid_h * grid_w + SE_SLICE - 1) / SE_SLICE;
    return 4 * ((int64_t)SE_MAX_IMAGES + (int64_t)n_images * slices * c);
}

extern "C" int gd_se_gate(const void* y_bf16, int32_t n_images, int32_t grid_h, int32_t grid_w, int32_t c, int32_t split,
                          int32_t c_real, int32_t c_hidden, const float* w1, const float* b1, const float* w2, const float* b2, float* gate,
                          void* scratch, int64_t scratch_bytes, void* stream) {
    if (!y_bf16 || !w1 || !b1 || !w2 || !b2 || !gate || !scratch) return set_error(GD_ERR_INVALID, "gd_se_gate: null pointer");
    if (n_images <= 0 || grid_h < 3 || grid_w < 3 || c < 32 || c > 256 || (c & (c - 1)) || c_real <= 0 || c_real > c ||
        c_hidden <= 0 || c_hidden > 32)
        return set_error(GD_ERR_INVALID, "gd_se_gate: bad shape (c in {32,64,128,256}, c_hidden <= 32)");
    const int slices = (grid_h * grid_w + SE_SLICE - 1) / SE_SLICE;
    if (scratch_bytes < gd_se_gate_scratch_bytes(n_images, grid_h, grid_w, c))
        return set_error(GD_ERR_INVALID, "gd_se_gate: scratch too small (see gd_se_gate_scratch_bytes)");
    if (n_images > 65535) return set_error(GD_ERR_INVALID, "gd_se_gate: at most 65535 images per launch");
    int* counters = reinterpret_cast<int*>(scratch);
    float* partial = reinterpret_cast<float*>(scratch) + SE_MAX_IMAGES;
    GD_CUDA_CHECK(launch_k(se_gate_kernel, dim3(slices, n_images), 256, 0, reinterpret_cast<cudaStream_t>(stream), 1,
                           reinterpret_cast<const __nv_bfloat16*>(y_bf16), grid_h * grid_w, (grid_h - 2) * (grid_w - 2), c,
                           split, c_real, c_hidden, w1, b1, w2, b2, gate, partial, counters));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

extern "C" int gd_se_residual_relu(const void* y_bf16, const void* residual_bf16, const float* gate, void* out_bf16,
                                   int32_t n_images, int32_t grid_h, int32_t grid_w, int32_t c, int32_t split, void* stream) {
    if (!y_bf16 || !residual_bf16 || !gate || !out_bf16) return set_error(GD_ERR_INVALID, "gd_se_residual_relu: null pointer");
    if (n_images <= 0 || grid_h < 3 || grid_w < 3 || c <= 0 || c % 8) return set_error(GD_ERR_INVALID, "gd_se_residual_relu: bad shape");
    const size_t total = (size_t)n_images * (grid_h - 2) * (grid_w - 2) * (c / 8);
    GD_CUDA_CHECK(launch_k(se_residual_relu_kernel, grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 1,
                           reinterpret_cast<const __nv_bfloat16*>(y_bf16), reinterpret_cast<const __nv_bfloat16*>(residual_bf16),
                           gate, reinterpret_cast<__nv_bfloat16*>(out_bf16), n_images, grid_h, grid_w, c, split));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

extern "C" int gd_pixel_shuffle_rows(const void* in_bf16, void* out_bf16, int32_t n_images, int32_t H, int32_t W,
                                     int32_t c_in, int32_t r, int32_t c_out_pad, int32_t split, void* stream) {
    if (!in_bf16 || !out_bf16) return set_error(GD_ERR_INVALID, "gd_pixel_shuffle_rows: null pointer");
    if (n_images <= 0 || H <= 0 || W <= 0 || r < 1 || c_in % (r * r) || c_out_pad % 8 || c_in / (r * r) > c_out_pad)
        return set_error(GD_ERR_INVALID, "gd_pixel_shuffle_rows: bad shape");
    const size_t total = (size_t)n_images * H * r * W * r * (c_out_pad / 8);
    GD_CUDA_CHECK(launch_k(pixel_shuffle_rows_kernel, grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 1,
                           reinterpret_cast<const __nv_bfloat16*>(in_bf16), reinterpret_cast<__nv_bfloat16*>(out_bf16), n_images,
                           H, W, c_in, r, c_out_pad, split));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

// ------------------------------------------------------------------------------------------ mel front end
// HA2GSpeechEncoder's wav2spec (speech_encoder.py:18-27,50): pre-emphasis (reflect pad 1), STFT n_fft 1024 / hop 512 /
// periodic Hann / centre reflect padding, power spectrum, 128-bin HTK mel filterbank, + eps.  One CTA per (clip, frame):
// the 1024 windowed samples go through an in-place radix-2 FFT in shared memory (fp32, host-computed float64 twiddles),
// then every mel bin sums its own (contiguous) band of the power spectrum.
namespace gd {

constexpr int MEL_NFFT = 1024, MEL_HOP = 512, MEL_BINS = 128, MEL_FREQS = MEL_NFFT / 2 + 1;

__global__ void __launch_bounds__(256) mel_power_kernel(const float* __restrict__ wav, int wav_len, int frames,
                                                        const float* __restrict__ window, const float2* __restrict__ twiddle,
                                                        const float* __restrict__ fb, const int2* __restrict__ fb_range,
                                                        float preemph, float add_eps, float* __restrict__ mel) {
    __shared__ float2 s[MEL_NFFT];
    __shared__ float2 s_tw[MEL_NFFT / 2];
    __shared__ float s_pow[MEL_FREQS];
    pdl_launch_dependents();
    pdl_wait();
    const int clip = blockIdx.x / frames, frame = blockIdx.x - clip * frames;
    const float* x = wav + (size_t)clip * wav_len;
    for (int i = threadIdx.x; i < MEL_NFFT / 2; i += 256) s_tw[i] = __ldg(twiddle + i);
    for (int i = threadIdx.x; i < MEL_NFFT; i += 256) {
        int j = frame * MEL_HOP + i - MEL_NFFT / 2;      // sample index of the pre-emphasised signal, reflect-padded
        if (j < 0) j = -j;
        if (j >= wav_len) j = 2 * (wav_len - 1) - j;
        const float cur = __ldg(x + j), prev = __ldg(x + (j > 0 ? j - 1 : 1));
        const float y = cur - preemph * prev;
        s[__brev((unsigned)i) >> 22] = make_float2(y * __ldg(window + i), 0.f);
    }
    __syncthreads();
#pragma unroll 1
    for (int half = 1; half < MEL_NFFT; half <<= 1) {
        const int tw_step = (MEL_NFFT / 2) / half;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int b = threadIdx.x + r * 256;
            const int j = b & (half - 1);
            const int i0 = ((b - j) << 1) + j, i1 = i0 + half;
            const float2 w = s_tw[j * tw_step], a = s[i0], c = s[i1];
            const float tr = w.x * c.x - w.y * c.y, ti = w.x * c.y + w.y * c.x;
            s[i0] = make_float2(a.x + tr, a.y + ti);
            s[i1] = make_float2(a.x - tr, a.y - ti);
        }
        __syncthreads();
    }
    for (int k = threadIdx.x; k < MEL_FREQS; k += 256) s_pow[k] = s[k].x * s[k].x + s[k].y * s[k].y;
    __syncthreads();
    if (threadIdx.x < MEL_BINS) {
        const int m = threadIdx.x;
        const int2 rg = __ldg(fb_range + m);  // [first, last] frequency bin with a non-zero weight
        float acc = 0.f;
        for (int k = rg.x; k <= rg.y; ++k) acc = fmaf(s_pow[k], __ldg(fb + k * MEL_BINS + m), acc);
        mel[((size_t)clip * MEL_BINS + m) * frames + frame] = acc + add_eps;
    }
}

// nn.InstanceNorm1d (no affine, biased variance): one warp per row, two passes over the row (L1-resident).
__global__ void __launch_bounds__(256) instance_norm_rows_kernel(float* __restrict__ x, int rows, int len, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* p = x + (size_t)row * len;
    float s = 0.f;
    for (int i = lane; i < len; i += 32) s += p[i];
    const float mean = warp_sum(s) / (float)len;
    float q = 0.f;
    for (int i = lane; i < len; i += 32) {
        const float d = p[i] - mean;
        q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)len + eps);
    for (int i = lane; i < len; i += 32) p[i] = (p[i] - mean) * rstd;
}

}  // namespace gd

extern "C" int gd_mel_power(const float* wav, int32_t n_clips, int32_t wav_len, const float* window, const float* twiddle,
                            const float* fb, const int32_t* fb_range, float preemph, float add_eps, float* mel,
                            void* stream) {
    if (!wav || !window || !twiddle || !fb || !fb_range || !mel) return set_error(GD_ERR_INVALID, "gd_mel_power: null pointer");
    if (n_clips <= 0 || wav_len <= MEL_NFFT / 2) return set_error(GD_ERR_INVALID, "gd_mel_power: wav_len must exceed 512 samples");
    const int frames = wav_len / MEL_HOP + 1;
    if ((int64_t)n_clips * frames >= ((int64_t)1 << 31)) return set_error(GD_ERR_INVALID, "gd_mel_power: too many frames");
    GD_CUDA_CHECK(launch_k(mel_power_kernel, n_clips * frames, 256, 0, reinterpret_cast<cudaStream_t>(stream), 1, wav, wav_len,
                           frames, window, reinterpret_cast<const float2*>(twiddle), fb, reinterpret_cast<const int2*>(fb_range),
                           preemph, add_eps, mel));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

extern "C" int gd_instance_norm_rows(float* x, int32_t rows, int32_t len, float eps, void* stream) {
    if (!x || rows <= 0 || len <= 0) return set_error(GD_ERR_INVALID, "gd_instance_norm_rows: bad argument");
    GD_CUDA_CHECK(launch_k(instance_norm_rows_kernel, (rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream), 1, x, rows,
                           len, eps));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}
