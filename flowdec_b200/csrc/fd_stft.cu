// flowdec_b200 — waveform <-> compressed complex spectrogram.
//
// Reference: flowdec/util/other.py:55-82 (normalize_noisy), flowdec/data/feature_extractors.py
// :86-109 (torch.stft / torch.istft, n_fft = 1534 = 2*13*59, hop 384, symmetric Hann, center /
// reflect, onesided -> 768 bins), :118-139 (|X|^alpha e^{j angle X} * beta and inverse),
// flowdec/util/other.py:25-52 (zero pad the time axis to a multiple of 64).
//
// n_fft is not a power of two, so the transform is evaluated as a direct DFT against a
// 1534-entry twiddle table held in shared memory (0.30 GMAC per audio-second, < 0.05 % of one
// backbone evaluation).  fp32 accumulation; table generated in fp64.
#include "fd_common.cuh"

namespace fd {

constexpr int kNfft = 1534;
constexpr int kHop = 384;
constexpr int kBins = 768;
constexpr int kPad = kNfft / 2;  // 767

// per-sample max |y| -> normfac (<= 1e-8 -> 1, torch.isclose(normfac, 0) with default atol)
// `lengths` (optional, int32 [B]): per-clip sample counts of a ragged batch stored with row pitch L
__global__ void __launch_bounds__(1024) normfac_kernel(const float* __restrict__ y, int L,
                                                        const int* __restrict__ lengths, int mode,
                                                        float* __restrict__ normfac) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int Lb = lengths ? min(lengths[b], L) : L;
  float m = 0.f;
  if (mode == 1)
    for (int i = threadIdx.x; i < Lb; i += blockDim.x) m = fmaxf(m, fabsf(y[static_cast<size_t>(b) * L + i]));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (threadIdx.x == 0) normfac[b] = (mode == 1) ? ((m <= 1e-8f) ? 1.0f : m) : 1.0f;
  }
}

// twiddle table: tw[j] = (cos(2 pi j / 1534), sin(2 pi j / 1534))
__global__ void twiddle_kernel(float2* __restrict__ tw) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < kNfft) {
    const double a = 2.0 * 3.14159265358979323846 * j / kNfft;
    tw[j] = make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
  }
}

// STFT + amplitude compression + zero padding of the frame axis.
//   block: 128 frames (4 per lane: lane + 32 f) x 8 bin-groups (warps), 8 bins per thread -> 64 bins per block;
//          one twiddle fetch (warp-uniform address -> broadcast) feeds 4 frames
//   grid : (ceil(Tp/128), 768/64, B)
// out[b][k][m] = beta * |X|^(alpha-1) * X,  X[k,m] = sum_n w[n] ypad[m*hop+n] e^{-2 pi i k n / N}
constexpr int kStftFpt = 4;                 // frames per thread
constexpr int kStftFrames = 32 * kStftFpt;  // frames per block
__global__ void __launch_bounds__(256) stft_compress_kernel(const float* __restrict__ y, int L,
                                                            const int* __restrict__ lengths,
                                                            const float* __restrict__ normfac,
                                                            const float* __restrict__ window,
                                                            const float2* __restrict__ tw_g,
                                                            float alpha, float beta, int frames_all, int Tp,
                                                            float2* __restrict__ out) {
  constexpr int CH = 59;  // samples per smem chunk; 1534 = 26 * 59
  __shared__ float2 tw[kNfft];
  __shared__ float sx[CH][kStftFrames + 1];
  const int lane = threadIdx.x & 31, wg = threadIdx.x >> 5;
  const int m0 = blockIdx.x * kStftFrames, k0 = blockIdx.y * 64 + wg * 8, b = blockIdx.z;
  for (int i = threadIdx.x; i < kNfft; i += 256) tw[i] = tw_g[i];
  const float inv_nf = 1.0f / normfac[b];
  const float* yb = y + static_cast<size_t>(b) * L;
  const int Lb = lengths ? min(lengths[b], L) : L;       // ragged batch: this clip's own length / frame count
  const int frames = lengths ? min(1 + Lb / kHop, Tp) : frames_all;
  float re[kStftFpt][8], im[kStftFpt][8];
  int idx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    idx[j] = 0;
#pragma unroll
    for (int f = 0; f < kStftFpt; ++f) re[f][j] = im[f][j] = 0.f;
  }
  for (int n0 = 0; n0 < kNfft; n0 += CH) {
    __syncthreads();
    for (int i = threadIdx.x; i < CH * kStftFrames; i += 256) {
      const int nn = i % CH, f = i / CH;
      const int m = m0 + f;
      float v = 0.f;
      if (m < frames) {
        int p = m * kHop + n0 + nn - kPad;  // index into the unpadded signal
        if (p < 0) p = -p;                  // reflect (no edge repeat)
        if (p >= Lb) p = 2 * (Lb - 1) - p;
        v = yb[p] * inv_nf * window[n0 + nn];
      }
      sx[nn][f] = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int nn = 0; nn < CH; ++nn) {
      float sv[kStftFpt];
#pragma unroll
      for (int f = 0; f < kStftFpt; ++f) sv[f] = sx[nn][lane + 32 * f];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 t = tw[idx[j]];        // (k_j * n) mod N, same for the whole warp
#pragma unroll
        for (int f = 0; f < kStftFpt; ++f) {
          re[f][j] = fmaf(sv[f], t.x, re[f][j]);
          im[f][j] = fmaf(-sv[f], t.y, im[f][j]);
        }
        idx[j] += k0 + j;
        if (idx[j] >= kNfft) idx[j] -= kNfft;
      }
    }
  }
#pragma unroll
  for (int f = 0; f < kStftFpt; ++f) {
    const int m = m0 + lane + 32 * f;
    if (m >= Tp) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float2 o = make_float2(0.f, 0.f);
      if (m < frames) {
        const float mag = sqrtf(re[f][j] * re[f][j] + im[f][j] * im[f][j]);
        if (mag > 0.f) {
          const float sc = beta * powf(mag, alpha - 1.0f);
          o = make_float2(re[f][j] * sc, im[f][j] * sc);
        }
      }
      out[(static_cast<size_t>(b) * kBins + (k0 + j)) * Tp + m] = o;
    }
  }
}

// decompression + inverse STFT + de-normalisation.
//   block = one hop (384 output samples) of one clip; the <= 4 frames overlapping it are
//   decompressed into smem, every thread accumulates its sample over frames x bins.
__global__ void __launch_bounds__(384) istft_decompress_kernel(const float2* __restrict__ X, int Tp,
                                                               int frames_all,
                                                               const int* __restrict__ lengths,
                                                               const float* __restrict__ window,
                                                               const float2* __restrict__ tw_g,
                                                               const float* __restrict__ normfac,
                                                               float alpha, float beta, int L,
                                                               float* __restrict__ out) {
  __shared__ float2 tw[kNfft];
  __shared__ float2 sX[4][kBins];
  const int b = blockIdx.y;
  const int hblk = blockIdx.x;  // output samples [hblk*384, hblk*384+384)
  const int Lb = lengths ? min(lengths[b], L) : L;
  const int frames = lengths ? min(1 + Lb / kHop, Tp) : frames_all;
  if (hblk * kHop >= Lb) {      // ragged batch: this hop lies beyond the clip -> zeros (uniform per block)
    const int n = hblk * kHop + threadIdx.x;
    if (n < L) out[static_cast<size_t>(b) * L + n] = 0.f;
    return;
  }
  for (int i = threadIdx.x; i < kNfft; i += 384) tw[i] = tw_g[i];
  // padded position p = n + 767 lies in [hblk*384 + 767, ...): frames m with 0 <= p - m*hop < N
  // for the block: m in [m_hi - 3, m_hi], m_hi = floor((hblk*384 + 767 + 383) / 384)
  const int m_hi = (hblk * kHop + kPad + kHop - 1) / kHop;
  const float inv_alpha_m1 = 1.0f / alpha - 1.0f;
  for (int i = threadIdx.x; i < 4 * kBins; i += 384) {
    const int fi = i / kBins, k = i % kBins;
    const int m = m_hi - 3 + fi;
    float2 v = make_float2(0.f, 0.f);
    if (m >= 0 && m < frames) {
      float2 x = X[(static_cast<size_t>(b) * kBins + k) * Tp + m];
      x.x /= beta;
      x.y /= beta;
      const float mag = sqrtf(x.x * x.x + x.y * x.y);
      if (mag > 0.f) {
        const float sc = powf(mag, inv_alpha_m1);
        v = make_float2(x.x * sc, x.y * sc);
      }
      // hermitian weights of the one-sided inverse: DC and Nyquist count once (imag ignored)
      if (k == 0 || k == kBins - 1) v.y = 0.f; else { v.x *= 2.f; v.y *= 2.f; }
    }
    sX[fi][k] = v;
  }
  __syncthreads();
  const int n = hblk * kHop + threadIdx.x;
  if (n >= L) return;
  if (n >= Lb) {
    out[static_cast<size_t>(b) * L + n] = 0.f;
    return;
  }
  const int p = n + kPad;
  float acc = 0.f, env = 0.f;
#pragma unroll 1
  for (int fi = 0; fi < 4; ++fi) {
    const int m = m_hi - 3 + fi;
    const int j = p - m * kHop;
    if (m < 0 || m >= frames || j < 0 || j >= kNfft) continue;
    const float wj = window[j];
    env = fmaf(wj, wj, env);
    // sum_k Re(X_k e^{+2 pi i j k / N}): the phasor advances by e^{2 pi i j / N} per bin; it is advanced by a
    // complex multiply and re-seeded from the table every 16 bins (fp32 drift <= ~1e-6), which replaces a
    // per-bin table lookup whose addresses (j*k mod N, j = lane-dependent) collide in the shared-memory banks
    const float2 rot = tw[j];
    const int inc16 = (16 * j) % kNfft;
    float s0 = 0.f, s1 = 0.f;
    int idx = 0;
#pragma unroll 1
    for (int kb = 0; kb < kBins; kb += 16) {
      float2 ph = tw[idx];
#pragma unroll
      for (int kk = 0; kk < 16; kk += 2) {
        const float2 x0 = sX[fi][kb + kk];
        s0 = fmaf(x0.x, ph.x, s0);
        s0 = fmaf(-x0.y, ph.y, s0);
        const float2 p1 = make_float2(fmaf(ph.x, rot.x, -ph.y * rot.y), fmaf(ph.x, rot.y, ph.y * rot.x));
        const float2 x1 = sX[fi][kb + kk + 1];
        s1 = fmaf(x1.x, p1.x, s1);
        s1 = fmaf(-x1.y, p1.y, s1);
        ph = make_float2(fmaf(p1.x, rot.x, -p1.y * rot.y), fmaf(p1.x, rot.y, p1.y * rot.x));
      }
      idx += inc16;
      if (idx >= kNfft) idx -= kNfft;
    }
    const float s = s0 + s1;
    acc = fmaf(wj, s * (1.0f / kNfft), acc);
  }
  out[static_cast<size_t>(b) * L + n] = (env > 1e-11f ? acc / env : 0.f) * normfac[b];
}

}  // namespace fd

namespace fd {   // fd_stft_pfa.cu
int stft_pfa_enabled();
int stft_pfa_launch(const float* y, int B, int L, const int* lengths, const float* normfac, const float* window,
                    const void* tw, float alpha, float beta, int frames, int Tp, void* out, cudaStream_t stream);
}  // namespace fd

using namespace fd;

extern "C" int fd_twiddles1534(void* tw, cudaStream_t stream) {
  twiddle_kernel<<<(kNfft + 255) / 256, 256, 0, stream>>>(static_cast<float2*>(tw));
  return check_launch("fd_twiddles1534");
}

extern "C" int fd_normfac(const float* y, int B, int L, int mode, float* normfac, cudaStream_t stream) {
  normfac_kernel<<<B, 1024, 0, stream>>>(y, L, nullptr, mode, normfac);
  return check_launch("fd_normfac");
}

extern "C" int fd_normfac_ragged(const float* y, int B, int L, const int* lengths, int mode, float* normfac,
                                 cudaStream_t stream) {
  FD_REQUIRE(lengths != nullptr, "fd_normfac_ragged: lengths is NULL");
  normfac_kernel<<<B, 1024, 0, stream>>>(y, L, lengths, mode, normfac);
  return check_launch("fd_normfac_ragged");
}

extern "C" int fd_stft1534_compress(const float* y, int B, int L, const float* normfac,
                                    const float* window, const void* tw, float alpha, float beta,
                                    int Tp, void* out, cudaStream_t stream) {
  FD_REQUIRE(L > kPad, "fd_stft1534_compress: L=%d must exceed the reflect pad %d", L, kPad);
  const int frames = 1 + L / kHop;
  FD_REQUIRE(Tp >= frames, "fd_stft1534_compress: Tp=%d < frames=%d", Tp, frames);
  if (stft_pfa_enabled() && Tp % 4 == 0)      // prime-factor FFT (fd_stft_pfa.cu); the direct DFT below is the A/B path
    return stft_pfa_launch(y, B, L, nullptr, normfac, window, tw, alpha, beta, frames, Tp, out, stream);
  dim3 grid((Tp + kStftFrames - 1) / kStftFrames, kBins / 64, B);
  stft_compress_kernel<<<grid, 256, 0, stream>>>(y, L, nullptr, normfac, window,
                                                 static_cast<const float2*>(tw), alpha, beta, frames, Tp,
                                                 static_cast<float2*>(out));
  return check_launch("fd_stft1534_compress");
}

extern "C" int fd_stft1534_compress_ragged(const float* y, int B, int L, const int* lengths,
                                           const float* normfac, const float* window, const void* tw,
                                           float alpha, float beta, int Tp, void* out, cudaStream_t stream) {
  FD_REQUIRE(lengths != nullptr && L > kPad, "fd_stft1534_compress_ragged: lengths NULL or pitch %d <= %d", L, kPad);
  FD_REQUIRE(Tp >= 1, "fd_stft1534_compress_ragged: Tp=%d", Tp);
  if (stft_pfa_enabled() && Tp % 4 == 0)
    return stft_pfa_launch(y, B, L, lengths, normfac, window, tw, alpha, beta, 0, Tp, out, stream);
  dim3 grid((Tp + kStftFrames - 1) / kStftFrames, kBins / 64, B);
  stft_compress_kernel<<<grid, 256, 0, stream>>>(y, L, lengths, normfac, window,
                                                 static_cast<const float2*>(tw), alpha, beta, 0, Tp,
                                                 static_cast<float2*>(out));
  return check_launch("fd_stft1534_compress_ragged");
}

extern "C" int fd_istft1534_decompress(const void* X, int B, int Tp, int L, const float* window,
                                       const void* tw, const float* normfac, float alpha, float beta,
                                       float* out, cudaStream_t stream) {
  const int frames = 1 + L / kHop;
  FD_REQUIRE(Tp >= frames, "fd_istft1534_decompress: Tp=%d < frames=%d", Tp, frames);
  dim3 grid((L + kHop - 1) / kHop, B);
  istft_decompress_kernel<<<grid, 384, 0, stream>>>(static_cast<const float2*>(X), Tp, frames, nullptr, window,
                                                    static_cast<const float2*>(tw), normfac, alpha, beta,
                                                    L, out);
  return check_launch("fd_istft1534_decompress");
}

extern "C" int fd_istft1534_decompress_ragged(const void* X, int B, int Tp, int L, const int* lengths,
                                              const float* window, const void* tw, const float* normfac,
                                              float alpha, float beta, float* out, cudaStream_t stream) {
  FD_REQUIRE(lengths != nullptr && L >= 1, "fd_istft1534_decompress_ragged: lengths NULL or empty pitch");
  dim3 grid((L + kHop - 1) / kHop, B);
  istft_decompress_kernel<<<grid, 384, 0, stream>>>(static_cast<const float2*>(X), Tp, 0, lengths, window,
                                                    static_cast<const float2*>(tw), normfac, alpha, beta,
                                                    L, out);
  return check_launch("fd_istft1534_decompress_ragged");
}
