// flowdec_b200 — 1534-point STFT / iSTFT as a prime-factor (Good-Thomas) FFT, fused with normalisation, amplitude
// (de)compression, frame padding and overlap-add.
//
// Replaces torch.stft / torch.istft (cuFFT) + CompressAmplitudesAndScale + pad_spec of the reference
// (/root/reference/flowdec/data/feature_extractors.py:86-139, util/other.py:25-52, model.py:150-152,182-187).
//
// 1534 = 26 x 59 with gcd(26, 59) = 1, so with n = (59 n1 + 26 n2) mod 1534 and (k1, k2) = (k mod 26, k mod 59)
//     W_1534^{n k} = W_26^{n1 k1} * W_59^{n2 k2}
// and the transform factors into 26 DFTs of length 59 followed by 59 DFTs of length 26 WITHOUT twiddle factors
// between the stages:
//     forward:  S[n1][k2] = sum_n2 x[n(n1,n2)] W_59^{n2 k2}         (x real  =>  S[n1][59-k2] = conj S[n1][k2]: 30 columns)
//               X[k]      = sum_n1 S[n1][k mod 59] W_26^{n1 (k mod 26)}                        k = 0 .. 767
//     inverse:  T[n1][k2] = sum_k1 Xfull[(885 k1 + 650 k2) mod 1534] W_26^{-n1 k1}             (CRT; 30 columns)
//               x[n(n1,n2)] = (T[n1][0] + 2 Re sum_{k2=1..29} T[n1][k2] W_59^{-n2 k2}) / 1534
// 172 K real FMAs per frame each way instead of 2.36 M for the direct DFT of fd_stft.cu (kept as the A/B path:
// fd_stft_use_pfa(0)).  fp32 throughout; the small DFTs are evaluated directly from 59- / 26-entry tables.
// A block transforms 4 consecutive frames of one clip, so every table fetch feeds 4 frames and the spectrogram
// (layout [B, 768, Tp], frames innermost) is read / written in 32-byte runs.
#include "fd_common.cuh"

namespace fd {

constexpr int kPN = 1534, kPHop = 384, kPBins = 768, kPPad = 767;
constexpr int kN1 = 26, kN2 = 59, kK2h = 30;
constexpr int kFr = 4;                      // frames per block
constexpr int kItems = kN1 * kK2h;          // 780 (n1, k2) pairs
constexpr int kFramePitch = 1536;           // floats per frame in the iSTFT workspace
// forward: sx 24544 + sS 24960 + tables 680 = 50184 B; inverse: sX 24576 + sT 24960 + tables 680 = 50216 B
constexpr int kPfaSmem = 50432;

__device__ __forceinline__ void load_tables(const float2* __restrict__ tw_g, float2* tw59, float2* tw26) {
  for (int i = threadIdx.x; i < kN2; i += blockDim.x) tw59[i] = tw_g[kN1 * i];   // W_1534^{26 i} = W_59^i  (cos, sin)
  for (int i = threadIdx.x; i < kN1; i += blockDim.x) tw26[i] = tw_g[kN2 * i];   // W_1534^{59 i} = W_26^i
}

// ------------------------------------------------------------------------------------------------
// forward: grid (Tp / 4, B), 256 threads
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stft_pfa_kernel(const float* __restrict__ y, int L, const int* __restrict__ lengths,
                                                       const float* __restrict__ normfac, const float* __restrict__ window,
                                                       const float2* __restrict__ tw_g, float alpha, float beta,
                                                       int frames_all, int Tp, float2* __restrict__ out) {
  extern __shared__ float smem[];
  float* sx = smem;                                              // [kFr][1534]
  float2* sS = reinterpret_cast<float2*>(sx + kFr * kPN);        // [780][kFr]
  float2* tw59 = sS + kItems * kFr;
  float2* tw26 = tw59 + kN2;
  const int b = blockIdx.y, m0 = blockIdx.x * kFr;
  const int Lb = lengths ? min(lengths[b], L) : L;
  const int frames = lengths ? min(1 + Lb / kPHop, Tp) : frames_all;
  if (m0 >= frames) {                      // padded frames (pad_spec 'zero'): uniform per block
    for (int i = threadIdx.x; i < kPBins; i += 256) {
      float4* o = reinterpret_cast<float4*>(out + (static_cast<size_t>(b) * kPBins + i) * Tp + m0);
      o[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  load_tables(tw_g, tw59, tw26);
  const float inv_nf = 1.0f / normfac[b];
  const float* yb = y + static_cast<size_t>(b) * L;
  for (int i = threadIdx.x; i < kFr * kPN; i += 256) {
    const int f = i / kPN, n = i - f * kPN;
    const int m = m0 + f;
    float v = 0.f;
    if (m < frames) {
      int p = m * kPHop + n - kPPad;       // index into the unpadded signal, reflected at both ends
      if (p < 0) p = -p;
      if (p >= Lb) p = 2 * (Lb - 1) - p;
      v = yb[p] * inv_nf * window[n];
    }
    sx[i] = v;
  }
  __syncthreads();
  // stage 1: 59-point DFTs over n2
  for (int it = threadIdx.x; it < kItems; it += 256) {
    const int n1 = it / kK2h, k2 = it - n1 * kK2h;
    int a = (kN2 * n1) % kPN, idx = 0;
    float re[kFr], im[kFr];
#pragma unroll
    for (int f = 0; f < kFr; ++f) re[f] = im[f] = 0.f;
#pragma unroll 1
    for (int n2 = 0; n2 < kN2; ++n2) {
      const float2 t = tw59[idx];
#pragma unroll
      for (int f = 0; f < kFr; ++f) {
        const float xv = sx[f * kPN + a];
        re[f] = fmaf(xv, t.x, re[f]);
        im[f] = fmaf(-xv, t.y, im[f]);
      }
      a += kN1;
      if (a >= kPN) a -= kPN;
      idx += k2;
      if (idx >= kN2) idx -= kN2;
    }
#pragma unroll
    for (int f = 0; f < kFr; ++f) sS[it * kFr + f] = make_float2(re[f], im[f]);
  }
  __syncthreads();
  // stage 2: 26-point DFTs over n1, compression, store
  for (int k = threadIdx.x; k < kPBins; k += 256) {
    const int k1 = k % kN1;
    int k2 = k % kN2;
    const bool cj = k2 >= kK2h;
    if (cj) k2 = kN2 - k2;
    float re[kFr], im[kFr];
#pragma unroll
    for (int f = 0; f < kFr; ++f) re[f] = im[f] = 0.f;
    int idx = 0;
#pragma unroll 1
    for (int n1 = 0; n1 < kN1; ++n1) {
      const float2 t = tw26[idx];                                  // W_26^{n1 k1} = t.x - i t.y
      const float4* sp = reinterpret_cast<const float4*>(sS + (n1 * kK2h + k2) * kFr);
      const float4 s01 = sp[0], s23 = sp[1];
      const float sr[kFr] = {s01.x, s01.z, s23.x, s23.z};
      const float si[kFr] = {s01.y, s01.w, s23.y, s23.w};
#pragma unroll
      for (int f = 0; f < kFr; ++f) {
        const float a = sr[f], bq = cj ? -si[f] : si[f];
        re[f] = fmaf(a, t.x, fmaf(bq, t.y, re[f]));
        im[f] = fmaf(bq, t.x, fmaf(-a, t.y, im[f]));
      }
      idx += k1;
      if (idx >= kN1) idx -= kN1;
    }
    float o[2 * kFr];
#pragma unroll
    for (int f = 0; f < kFr; ++f) {
      float2 v = make_float2(0.f, 0.f);
      if (m0 + f < frames) {
        const float mag = sqrtf(re[f] * re[f] + im[f] * im[f]);
        if (mag > 0.f) {
          const float sc = beta * powf(mag, alpha - 1.0f);
          v = make_float2(re[f] * sc, im[f] * sc);
        }
      }
      o[2 * f] = v.x;
      o[2 * f + 1] = v.y;
    }
    float4* op = reinterpret_cast<float4*>(out + (static_cast<size_t>(b) * kPBins + k) * Tp + m0);
    op[0] = make_float4(o[0], o[1], o[2], o[3]);
    op[1] = make_float4(o[4], o[5], o[6], o[7]);
  }
}

// ------------------------------------------------------------------------------------------------
// inverse, part 1: decompress + inverse transform + synthesis window of 4 frames -> fr[B][Tp][1536]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) istft_pfa_frames_kernel(const float2* __restrict__ X, int Tp, int frames_all,
                                                               const int* __restrict__ lengths, int L,
                                                               const float* __restrict__ window,
                                                               const float2* __restrict__ tw_g, float alpha, float beta,
                                                               float* __restrict__ fr) {
  extern __shared__ float smem[];
  float2* sX = reinterpret_cast<float2*>(smem);                  // [768][kFr] one-sided spectrum, decompressed
  float2* sT = sX + kPBins * kFr;                                // [780][kFr]
  float2* tw59 = sT + kItems * kFr;
  float2* tw26 = tw59 + kN2;
  const int b = blockIdx.y, m0 = blockIdx.x * kFr;
  const int Lb = lengths ? min(lengths[b], L) : L;
  const int frames = lengths ? min(1 + Lb / kPHop, Tp) : frames_all;
  if (m0 >= frames) return;                // frames beyond the clip are never read by the overlap-add
  load_tables(tw_g, tw59, tw26);
  const float inv_alpha_m1 = 1.0f / alpha - 1.0f, inv_beta = 1.0f / beta;
  for (int k = threadIdx.x; k < kPBins; k += 256) {
    const float4* xp = reinterpret_cast<const float4*>(X + (static_cast<size_t>(b) * kPBins + k) * Tp + m0);
    const float4 x01 = xp[0], x23 = xp[1];
    const float xr[kFr] = {x01.x, x01.z, x23.x, x23.z};
    const float xi[kFr] = {x01.y, x01.w, x23.y, x23.w};
#pragma unroll
    for (int f = 0; f < kFr; ++f) {
      float2 v = make_float2(0.f, 0.f);
      if (m0 + f < frames) {
        const float a = xr[f] * inv_beta, c = xi[f] * inv_beta;
        const float mag = sqrtf(a * a + c * c);
        if (mag > 0.f) {
          const float sc = powf(mag, inv_alpha_m1);
          v = make_float2(a * sc, c * sc);
        }
        if (k == 0 || k == kPBins - 1) v.y = 0.f;     // irfft ignores the imaginary part of DC and Nyquist
      }
      sX[k * kFr + f] = v;
    }
  }
  __syncthreads();
  // stage A: 26-point inverse DFTs over k1 at the CRT positions k = (885 k1 + 650 k2) mod 1534
  for (int it = threadIdx.x; it < kItems; it += 256) {
    const int n1 = it / kK2h, k2 = it - n1 * kK2h;
    int kk = (650 * k2) % kPN, idx = 0;
    float re[kFr], im[kFr];
#pragma unroll
    for (int f = 0; f < kFr; ++f) re[f] = im[f] = 0.f;
#pragma unroll 1
    for (int k1 = 0; k1 < kN1; ++k1) {
      const float2 t = tw26[idx];                                  // W_26^{-n1 k1} = t.x + i t.y
      const bool cj = kk >= kPBins;                                // Xfull[k] = conj X[1534 - k] for k > 767
      const float4* sp = reinterpret_cast<const float4*>(sX + (cj ? kPN - kk : kk) * kFr);
      const float4 s01 = sp[0], s23 = sp[1];
      const float sr[kFr] = {s01.x, s01.z, s23.x, s23.z};
      const float si[kFr] = {s01.y, s01.w, s23.y, s23.w};
#pragma unroll
      for (int f = 0; f < kFr; ++f) {
        const float a = sr[f], bq = cj ? -si[f] : si[f];
        re[f] = fmaf(a, t.x, fmaf(-bq, t.y, re[f]));
        im[f] = fmaf(a, t.y, fmaf(bq, t.x, im[f]));
      }
      kk += 885;
      if (kk >= kPN) kk -= kPN;
      idx += n1;
      if (idx >= kN1) idx -= kN1;
    }
#pragma unroll
    for (int f = 0; f < kFr; ++f) sT[it * kFr + f] = make_float2(re[f], im[f]);
  }
  __syncthreads();
  // stage B: 59-point inverse DFTs over k2 (hermitian: 30 columns), synthesis window, 1 / N
  for (int n = threadIdx.x; n < kPN; n += 256) {
    const int n1 = (15 * (n % kN1)) % kN1;       // n = 59 n1 + 26 n2 (mod 1534): 59^-1 = 15 (mod 26), 26^-1 = 25 (mod 59)
    const int n2 = (25 * (n % kN2)) % kN2;
    float acc[kFr];
    {
      const float4* sp = reinterpret_cast<const float4*>(sT + (n1 * kK2h) * kFr);
      const float4 s01 = sp[0], s23 = sp[1];
      acc[0] = 0.5f * s01.x; acc[1] = 0.5f * s01.z; acc[2] = 0.5f * s23.x; acc[3] = 0.5f * s23.z;
    }
    int idx = n2;
#pragma unroll 1
    for (int k2 = 1; k2 < kK2h; ++k2) {
      const float2 t = tw59[idx];                                  // W_59^{-n2 k2} = t.x + i t.y
      const float4* sp = reinterpret_cast<const float4*>(sT + (n1 * kK2h + k2) * kFr);
      const float4 s01 = sp[0], s23 = sp[1];
      acc[0] = fmaf(s01.x, t.x, fmaf(-s01.y, t.y, acc[0]));
      acc[1] = fmaf(s01.z, t.x, fmaf(-s01.w, t.y, acc[1]));
      acc[2] = fmaf(s23.x, t.x, fmaf(-s23.y, t.y, acc[2]));
      acc[3] = fmaf(s23.z, t.x, fmaf(-s23.w, t.y, acc[3]));
      idx += n2;
      if (idx >= kN2) idx -= kN2;
    }
    const float wn = window[n] * (2.0f / kPN);
#pragma unroll
    for (int f = 0; f < kFr; ++f)
      if (m0 + f < frames) fr[(static_cast<size_t>(b) * Tp + m0 + f) * kFramePitch + n] = acc[f] * wn;
  }
}

// inverse, part 2: overlap-add of the (already windowed) frames, window-envelope normalisation, de-normalisation
__global__ void __launch_bounds__(256) istft_pfa_ola_kernel(const float* __restrict__ fr, int Tp, int frames_all,
                                                            const int* __restrict__ lengths, const float* __restrict__ window,
                                                            const float* __restrict__ normfac, int L, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= L) return;
  const int Lb = lengths ? min(lengths[b], L) : L;
  const int frames = lengths ? min(1 + Lb / kPHop, Tp) : frames_all;
  float r = 0.f;
  if (n < Lb) {
    const int p = n + kPPad;
    const int m_hi = min(p / kPHop, frames - 1);
    float acc = 0.f, env = 0.f;
    for (int m = m_hi; m >= 0; --m) {
      const int j = p - m * kPHop;
      if (j >= kPN) break;
      const float wj = window[j];
      env = fmaf(wj, wj, env);
      acc += fr[(static_cast<size_t>(b) * Tp + m) * kFramePitch + j];
    }
    r = (env > 1e-11f ? acc / env : 0.f) * normfac[b];
  }
  out[static_cast<size_t>(b) * L + n] = r;
}

static int g_use_pfa = 1;

int stft_pfa_enabled() { return g_use_pfa; }

int stft_pfa_launch(const float* y, int B, int L, const int* lengths, const float* normfac, const float* window,
                    const void* tw, float alpha, float beta, int frames, int Tp, void* out, cudaStream_t stream) {
  FD_REQUIRE(Tp % kFr == 0, "fd_stft1534_compress: Tp=%d must be a multiple of %d for the prime-factor kernel", Tp, kFr);
  static bool attr_set[kMaxDevices] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev = (dev >= 0 && dev < kMaxDevices) ? dev : 0;
  if (!attr_set[dev]) {
    cudaFuncSetAttribute(stft_pfa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfaSmem);
    cudaFuncSetAttribute(istft_pfa_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfaSmem);
    attr_set[dev] = true;
  }
  stft_pfa_kernel<<<dim3(Tp / kFr, B), 256, kPfaSmem, stream>>>(y, L, lengths, normfac, window,
                                                                 static_cast<const float2*>(tw), alpha, beta, frames, Tp,
                                                                 static_cast<float2*>(out));
  return check_launch("fd_stft1534_compress(pfa)");
}

}  // namespace fd

using namespace fd;

// 1: prime-factor FFT kernels (default); 0: the direct-DFT kernels of fd_stft.cu.  Returns the previous setting.
extern "C" int fd_stft_use_pfa(int on) {
  const int prev = g_use_pfa;
  g_use_pfa = on;
  return prev;
}

// iSTFT through the prime-factor kernels.  frames_ws: caller-owned workspace of B * Tp * 1536 floats (the windowed
// time-domain frames before overlap-add).  lengths may be NULL (every clip has L samples).
extern "C" int fd_istft1534_decompress_pfa(const void* X, int B, int Tp, int L, const int* lengths, const float* window,
                                           const void* tw, const float* normfac, float alpha, float beta,
                                           float* frames_ws, float* out, cudaStream_t stream) {
  FD_REQUIRE(X != nullptr && frames_ws != nullptr && out != nullptr, "fd_istft1534_decompress_pfa: NULL pointer");
  FD_REQUIRE(Tp % kFr == 0 && Tp >= kFr, "fd_istft1534_decompress_pfa: Tp=%d must be a multiple of %d", Tp, kFr);
  const int frames = 1 + L / kPHop;
  FD_REQUIRE(lengths != nullptr || Tp >= frames, "fd_istft1534_decompress_pfa: Tp=%d < frames=%d", Tp, frames);
  static bool attr_set[kMaxDevices] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev = (dev >= 0 && dev < kMaxDevices) ? dev : 0;
  if (!attr_set[dev]) {
    cudaFuncSetAttribute(stft_pfa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfaSmem);
    cudaFuncSetAttribute(istft_pfa_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfaSmem);
    attr_set[dev] = true;
  }
  istft_pfa_frames_kernel<<<dim3(Tp / kFr, B), 256, kPfaSmem, stream>>>(
      static_cast<const float2*>(X), Tp, lengths ? 0 : frames, lengths, L, window, static_cast<const float2*>(tw), alpha,
      beta, frames_ws);
  if (check_launch("fd_istft1534_decompress_pfa(frames)")) return 2;
  istft_pfa_ola_kernel<<<dim3((L + 255) / 256, B), 256, 0, stream>>>(frames_ws, Tp, lengths ? 0 : frames, lengths, window,
                                                                     normfac, L, out);
  return check_launch("fd_istft1534_decompress_pfa(ola)");
}
