/* flowdec_b200 — C ABI of libflowdec_b200.so (sm_100a kernels for the FlowDec postfilter path).
 *
 * Conventions
 *   - every function returns 0 on success; non-zero = failure, message via fd_last_error()
 *     (thread-local).  Kernel launch errors are checked (the reference's pybind op does not,
 *     op/upfirdn2d_kernel.cu:220-380).
 *   - all pointers are raw DEVICE pointers owned by the caller unless marked "host"; nothing
 *     is allocated, freed or synchronised inside; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*), re-entrant per stream, CUDA-graph capturable.
 *   - layouts: activations bf16 NHWC [B,H,W,C] (H = frequency bins, W = STFT frames);
 *     4-channel pyramids fp32 [B,H,W,4]; spectrograms / ODE state float2 [B,F,T]
 *     (bit-identical to torch.complex64 [B,1,F,T]); waveforms fp32 [B,L].
 *
 * The reference's native boundary for this path is a pybind11 module built at import
 * (flowdec/backbones/ncsnpp_utils/op/upfirdn2d.cpp:38-48, fused_bias_act.cpp:37-46); the rest
 * of its hot path is ATen/cuDNN/cuFFT calls from Python.  Each entry below names the reference
 * code it replaces.  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 */
#ifndef FLOWDEC_B200_H
#define FLOWDEC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fd_stream_t; /* cudaStream_t */

const char* fd_last_error(void);
int fd_abi_version(void);

/* ---- NCSN++ convolutions: tcgen05 implicit GEMM ---------------------------------------------
 * replaces nn.Conv2d -> cuDNN (flowdec/backbones/ncsnpp_utils/layers.py:110-134 as used by
 * layerspp.py:235,243,245 and ncsnpp.py:218,230) including the res-block's skip path and
 * (x + h)/sqrt(2) (layerspp.py:278-284), which are folded into extra K segments. */
struct fd_conv_src {
  const void* ptr;          /* bf16 NHWC [B,H,W,C] */
  int C;                    /* channel pitch of the tensor */
  int c_begin;              /* first channel consumed (multiple of 8) */
  int c_count;              /* channels consumed (multiple of 64) */
  int taps;                 /* 9 = 3x3 zero-padded, 1 = 1x1 */
  const float* scale_shift; /* NULL, or GroupNorm scale/shift fp32 [B][ss_pitch][2] positioned at this
                               source's first consumed channel: the kernel feeds SiLU(x*scale+shift)
                               (zero outside the image) to the MMA instead of x (layerspp.py:253,274) */
  int ss_pitch;             /* channels per sample in that table (width of the virtual concat) */
};
/* out[b,h,w,o] = bias[o] + sum_seg sum_tap sum_c src_seg[b,h+dh,w+dw,c] * wpacked[o, k(seg,tap,c)]
 * wpacked: bf16 [npad, ktot], K-major, k ordered segment-major, then tap (kh-major), then channel.
 * out_is_f32 = 0: out bf16 NHWC [B,H,W,cout], cout == npad in {128,256}
 * out_is_f32 = 1: out fp32 NHWC [B,H,W,cout], cout % 4 == 0, cout <= npad, npad in {16, 48}
 *                 (the 4-channel pyramid convs, directly or as 36 per-tap partial products)
 * bias: fp32 [npad] or NULL.  Requires W % 8 == 0 and H % (128 / min(128, pow2 divisor of W)) == 0.
 * stats (bf16 output only, may be NULL): GroupNorm partial sums of the OUTPUT, fp32
 * [B, S, cout, 2] with S = H*W/128 slabs (one per 128-pixel tile; the four epilogue warps are combined in shared
 * memory in a fixed order), consumed by
 * fd_gn_finalize — the statistics pass of the next GroupNorm fused into this conv's epilogue.
 * max_ctas: 0 = one persistent CTA per SM.  flags bit 0: CTA pairs (cta_group::2, 256-row MMAs, weight
 * tile split across the pair) when the tile count is even; bit 1: 16x8-pixel "halo" tiles (ONE 18x10-pixel
 * A box per k-slice serves all nine taps through row-shifted swizzled descriptors) when W % 8 == 0 and
 * H % 16 == 0 — required for sources with scale_shift != NULL (GroupNorm+SiLU fused into the operand). */
int fd_conv2d_igemm(const struct fd_conv_src* srcs, int nsrc, const void* wpacked, int ktot,
                    const float* bias, void* out, int out_is_f32, int cout, int npad, int B, int H,
                    int W, float* stats, int max_ctas, int flags, fd_stream_t stream);

/* 1: the bf16 halo convolutions run in clusters of FOUR CTAs — two MMA pairs sharing one weight stream (each weight half
 * is loaded once and TMA-multicast to the same-parity CTA of both pairs) — whenever the tile count is a multiple of 4;
 * 0 (default): CTA pairs only (measured faster on B200: clusters of 4 leave 16 of 148 SMs idle).  Returns the previous
 * setting. */
int fd_conv_cluster4(int on);

/* ---- GroupNorm + SiLU + FIR resampling ------------------------------------------------------
 * replace nn.GroupNorm / nn.SiLU (layerspp.py:229,241,253,274; ncsnpp.py:216,228) and
 * upsample_2d / downsample_2d -> upfirdn2d (up_or_down_sampling.py:220-282,
 * op/upfirdn2d.cpp:38-48, op/upfirdn2d_kernel.cu:118-218). */
/* partial[b][s][c][0..1] = sum / sum of squares of x over slab s of the HW pixels (S slabs) */
int fd_chan_stats(const void* x_bf16, int B, int HW, int C, float* partial, int S, fd_stream_t stream);
/* out[b][chunk][c][2] = sum over the chunk's slabs of in[b][s][c][2] (optional compaction of many slabs; the backbone
 * no longer needs it: fd_gn_finalize walks whole slabs) */
int fd_slab_reduce(const float* in, int B, int S, int C, float* out, int chunks, fd_stream_t stream);
/* group statistics over the virtual channel concat [part1 [B,S1,C1,2], part2 [B,S2,C2,2]] (fp64) ->
 * scale_shift fp32 [B, C1+C2, 2]:  y = x * scale + shift == GroupNorm(x) with gamma/beta */
int fd_gn_finalize(const float* part1, int C1, int S1, const float* part2, int C2, int S2, int B,
                   double count, const float* gamma, const float* beta, int groups, float eps,
                   float* scale_shift, fd_stream_t stream);
/* act = FIR_mode(SiLU(x*scale+shift)) -> out, and/or raw = FIR_mode(x) -> out_raw (the res-block's
 * resampled skip input, layerspp.py:257-268), over the virtual concat [src1, src2]; either output
 * may be NULL.  mode 0 none (activated output only), 1 down x2 ([1,3,3,1]/8 per axis, pad (1,1)),
 * 2 up x2 ([1,3,3,1]/4, pad (2,1)); outputs bf16 NHWC [B,H',W',C1+C2] */
int fd_gn_act_resample(const void* src1, int C1, const void* src2, int C2, const float* scale_shift,
                       void* out, void* out_raw, int B, int H, int W, int mode, fd_stream_t stream);
/* the up / down variants with both outputs run as TMA-tiled kernels (fd_fir_tiles.cu) when C1, C2 % 64 == 0,
 * H % 8 == 0, W % 16 == 0; fd_fir_tiles_enable(0) forces the register kernels (returns the previous setting) */
int fd_fir_tiles_enable(int on);

/* ---- 4-channel paths of NCSN++ ---------------------------------------------------------------*/
/* ncsnpp.py:261,401-404: (re x, im x, re y, im y) -> fp32 [npix,4] */
int fd_pack4(const void* x_f2, const void* y_f2, void* out4, size_t npix, fd_stream_t stream);
/* ncsnpp.py:300 pyramid_downsample on the 4-channel input pyramid */
int fd_fir_down4(const void* in4, void* out4, int B, int H, int W, fd_stream_t stream);
/* ncsnpp.py:355,360: out = FIR_up(lo[B,H,W,4]) + add[B,2H,2W,4] (out may alias add) */
int fd_pyramid_up_add(const void* lo4, const void* add4, void* out4, int B, int H, int W, fd_stream_t stream);
/* ncsnpp.py:218,230,355-360 pyramid conv 3x3 C->4 finished "GEMM first, shift after": part fp32
 * [B,H,W,part_channels] holds the 36 per-tap products W_tap . a (fd_conv2d_igemm, npad 48, 1 tap);
 * out[p] = bias + sum_tap part[p + delta_tap][tap*4 + co] (+ FIR_up(lo4[B,H/2,W/2,4]) if lo4 != NULL) */
int fd_pyramid_gather(const float* part, int part_channels, const float* bias, const void* lo4, void* out4,
                      int B, int H, int W, fd_stream_t stream);
/* ncsnpp.py:284 input conv 3x3 4->64 (w fp32 OIHW [64,4,3,3]) -> bf16 NHWC [B,H,W,64] */
int fd_conv_in(const void* in4, const float* w, const float* bias, void* out, int B, int H, int W,
               fd_stream_t stream);
/* layerspp.py:62-69 Combine('sum'): out = h + Conv1x1_{4->C}(pyr) + bias   (w fp32 [C,4]) */
int fd_combine(const void* pyr4, const float* w, const float* bias, const void* h, void* out,
               size_t npix, int C, fd_stream_t stream);
/* ncsnpp.py:398 output 1x1 conv (w_out_host8 = HOST pointer to the 2x4 weights) fused with one
 * sampler stage: torchdyn Euler/Midpoint and Heun2 (flowdec/sampling/solvers.py:15-57), the
 * reverse-diffusion predictor (sampling/predictors.py:61-71, sdes.py:118-123) and the ALD corrector
 * (sampling/correctors.py:54-66) are all affine in (x, y, noise, v):
 *   v = W_out * pyr ; out = c1*base1 + c2*base2 + c3*base3 + coef*v  (NULL bases / out / v_out allowed) */
int fd_output_axpy(const void* pyr4, const float* w_out_host8, const void* base1, float c1,
                   const void* base2, float c2, const void* base3, float c3, float coef, void* out,
                   void* v_out, size_t npix, fd_stream_t stream);
/* model.py:512,530-536: out = Y + fac * (float)(sigma[f] (f64) * eps) */
int fd_x0(const void* Y, const double* sigma, const void* eps, float fac, void* out, int B, int F, int T,
          fd_stream_t stream);

/* ---- time embedding (ncsnpp.py:263-274, layerspp.py:49-51,270-272) --------------------------*/
int fd_fourier_embed(float t, const float* Wf, int nf, float* out /* [2*nf] */, fd_stream_t stream);
/* out[m] = (add ? add[m] : 0) + out_scale * (b[m] + sum_k W[m,k] * (silu_in ? SiLU(in[k]) : in[k])) */
int fd_matvec(const float* in, int K, int silu_in, const float* W, const float* b, const float* add,
              float out_scale, float* out, int M, fd_stream_t stream);

/* ---- waveform <-> compressed spectrogram (util/other.py:25-82, feature_extractors.py:86-139) -*/
int fd_twiddles1534(void* tw /* float2 [1534] */, fd_stream_t stream);
/* mode 1: normfac[b] = max|y[b,:]| with values <= 1e-8 replaced by 1 ; mode 0: 1 */
int fd_normfac(const float* y, int B, int L, int mode, float* normfac, fd_stream_t stream);
/* out float2 [B,768,Tp]: beta*|X|^alpha e^{j angle X} of the n_fft=1534/hop=384 centred STFT of
 * y/normfac; frames >= 1+L/384 are zero (pad_spec 'zero') */
int fd_stft1534_compress(const float* y, int B, int L, const float* normfac, const float* window,
                         const void* tw, float alpha, float beta, int Tp, void* out, fd_stream_t stream);
/* inverse: decompress, overlap-add iSTFT (torch.istft semantics incl. length=L), times normfac */
int fd_istft1534_decompress(const void* X, int B, int Tp, int L, const float* window, const void* tw,
                            const float* normfac, float alpha, float beta, float* out, fd_stream_t stream);
/* STFT / iSTFT algorithm: 1 (default) = prime-factor FFT (1534 = 26 x 59, fd_stft_pfa.cu; needs Tp % 4 == 0, which
 * pad_spec's multiple of 64 always gives), 0 = direct DFT.  Returns the previous setting. */
int fd_stft_use_pfa(int on);
/* iSTFT through the prime-factor kernels: same contract as fd_istft1534_decompress(_ragged) (lengths may be NULL) plus
 * a caller-owned workspace frames_ws of B * Tp * 1536 floats (windowed time-domain frames before overlap-add). */
int fd_istft1534_decompress_pfa(const void* X, int B, int Tp, int L, const int* lengths, const float* window,
                                const void* tw, const float* normfac, float alpha, float beta, float* frames_ws,
                                float* out, fd_stream_t stream);
/* ragged batches (length-bucketed batching across files): rows of pitch L hold clips of
 * lengths[b] <= L samples (int32 [B], device; every lengths[b] > 767 and 1 + lengths[b]/384 <= Tp, checked by
 * the caller).  Each clip gets exactly the frames, reflect padding, normalisation and istft(length=) it would
 * get when processed alone (enhance.py:113-131 runs one file at a time); samples / frames beyond a clip are 0. */
int fd_normfac_ragged(const float* y, int B, int L, const int* lengths, int mode, float* normfac,
                      fd_stream_t stream);
int fd_stft1534_compress_ragged(const float* y, int B, int L, const int* lengths, const float* normfac,
                                const float* window, const void* tw, float alpha, float beta, int Tp, void* out,
                                fd_stream_t stream);
int fd_istft1534_decompress_ragged(const void* X, int B, int Tp, int L, const int* lengths, const float* window,
                                   const void* tw, const float* normfac, float alpha, float beta, float* out,
                                   fd_stream_t stream);

/* ---- upstream NDAC (descript-audio-codec 1.0.0; call sites demo.ipynb:101-105) ----------------
 * layout [B, C, T] fp32 as upstream; weight-norm already folded into w. */
/* dac.nn.quantize.ResidualVectorQuantize.from_codes: z[b,d,t] = sum_i out_proj_i(codebook_i[codes[b,i,t]])
 * codes int64 [B,nq,T]; codebooks [nq_total,csize,cdim]; out_proj_w [nq_total,D,cdim]; out_proj_b [nq_total,D] */
int fd_rvq_from_codes(const long long* codes, const float* codebooks, const float* out_proj_w,
                      const float* out_proj_b, float* z, int B, int nq, int T, int D, int codebook_dim,
                      int codebook_size, fd_stream_t stream);
/* WNConv1d with optional Snake1d on the input (alpha [Cin] or NULL), optional residual add and tanh:
 * out[b,co,t] = epi(bias[co] + sum w[co,ci,k] * snake(x[b,ci,t + k*dilation - pad])) (+ residual) */
int fd_dac_conv1d(const float* x, const float* w, const float* bias, const float* snake_alpha,
                  const float* residual, float* out, int B, int Cin, int Cout, int Tin, int K, int dilation,
                  int pad, int do_tanh, fd_stream_t stream);
/* encoder down-sampling conv (dac/model/dac.py EncoderBlock: Snake1d -> WNConv1d(k = 2*stride, stride,
 * padding ceil(stride/2))): out[b,co,t] = bias[co] + sum w[co,ci,k] * snake(x[b,ci,t*stride + k - pad]) */
int fd_dac_conv1d_strided(const float* x, const float* w, const float* bias, const float* snake_alpha,
                          float* out, int B, int Cin, int Cout, int Tin, int K, int stride, int pad,
                          fd_stream_t stream);
/* dac.nn.quantize.ResidualVectorQuantize.forward in eval mode with n_quantizers = nq (demo.ipynb:102 via
 * DAC.encode): per quantizer z_e = in_proj_i(residual); nearest code between the L2-normalised z_e and the
 * L2-normalised codebook (VectorQuantize.decode_latents; first index wins ties); z_q_i = out_proj_i(code).
 * z, zq [B,D,T]; in_proj_w [nq_total,cdim,D]; in_proj_b [nq_total,cdim]; codebooks / codebooks_l2n
 * [nq_total,csize,cdim] (raw / row-normalised); codebooks_l2n_sq [nq_total,csize] = |row|^2 of the normalised
 * table; out_proj_w [nq_total,D,cdim]; out_proj_b [nq_total,D]; codes int64 [B,nq,T]; latents [B,nq*cdim,T];
 * loss [1] or NULL = commitment (= codebook) loss; sqerr_ws workspace [nq*B*T] floats */
int fd_rvq_encode(const float* z, const float* in_proj_w, const float* in_proj_b, const float* codebooks,
                  const float* codebooks_l2n, const float* codebooks_l2n_sq, const float* out_proj_w,
                  const float* out_proj_b, long long* codes, float* zq, float* latents, float* loss,
                  float* sqerr_ws, int B, int nq, int T, int D, int codebook_dim, int codebook_size,
                  fd_stream_t stream);
/* Snake1d -> WNConvTranspose1d(kernel 2*stride, stride, padding pad); w [Cin,Cout,2*stride] */
int fd_dac_conv_transpose1d(const float* x, const float* w, const float* bias, const float* snake_alpha,
                            float* out, int B, int Cin, int Cout, int Tin, int stride, int pad,
                            fd_stream_t stream);

/* ---- NDAC decoder on tensor cores (fd_dac_tc.cu): activations fp32 time-major [B, T, C] ------------------
 * One decoder layer as a tf32 implicit GEMM: out[b,t,n] = bias[n] + sum_tap sum_c x[b, t + tap_offsets[tap], c] *
 * wpacked[n, tap*Cin + c] (+ residual[b,t,n]); x rows outside [0, Tin) read as zero.  wpacked fp32 [Ntot, ntaps*Cin],
 * tf32-rounded; Cin % 32 == 0, Ntot % 32 == 0, max - min of tap_offsets (HOST int array) <= 56.
 *   Conv1d(k=7, dilation d, pad 3d): offsets (j-3)*d.  Conv1d(k=1): {0}.
 *   ConvTranspose1d(k=2s, stride s, pad p): offsets {0,-1}, Ntot = s*Cout, wpacked[r*Cout+co, tap*Cin+ci] =
 *   w[ci,co,r+tap*s], Tout_rows = Tin + 1; the result [B, Tin+1, s*Cout] read as [B, (Tin+1)*s, Cout] holds the
 *   transposed conv's output at rows p .. p + Tout - 1.
 * out_raw and/or out_act (either may be NULL): fp32 [B, Tout_rows, Ntot] with batch stride out_bstride (elements);
 * out_act = Snake1d(out_raw) with alpha_next[n % alpha_mod] (x + sin^2(a x)/(a + 1e-9)), rounded to tf32 — the
 * NEXT layer's activation applied in this layer's epilogue.  x / residual batch strides in elements. */
int fd_dac_conv_tc(const float* x, int B, int Tin, long long x_bstride, int Cin, const float* wpacked, int Ntot,
                   int ntaps, const int* tap_offsets, const float* bias, const float* residual, long long res_bstride,
                   const float* alpha_next, int alpha_mod, float* out_raw, float* out_act, int Tout_rows,
                   long long out_bstride, fd_stream_t stream);
/* last decoder layer: out[b,t] = tanh(bias[0] + sum_{k<7,c} x_act[b, t+k-3, c] * w[c,k]);  w fp32 [C,7] */
int fd_dac_final_conv(const float* x_act, long long x_bstride, const float* w, const float* bias, float* out, int B,
                      int T, int C, fd_stream_t stream);
/* [B,C,T] -> [B,T,C]; round_tf32_out = 1 rounds to tf32 (the latent from fd_rvq_from_codes entering the tensor-core
 * decoder); with C and T swapped it is the way back (the encoder's latent for fd_rvq_encode) */
int fd_dac_nct_to_ntc(const float* in, float* out, int B, int C, int T, int round_tf32_out, fd_stream_t stream);
/* first encoder layer Conv1d(1 -> C, k7, pad 3) on x [B,T] -> raw and/or Snake-activated [B,T,C] (w fp32 [C,7]) */
int fd_dac_first_conv(const float* x, const float* w, const float* bias, const float* alpha, float* raw, float* act,
                      int B, int T, int C, fd_stream_t stream);
/* Encoder down-sampling conv (k = 2s, stride s, pad p) through fd_dac_conv_tc: read the activated input [B,T,C] as
 * [B, T/s, s*C] (a free view), offsets {-1,0,1}, wpacked[co, tap*(s*C) + j*C + ci] = w[co,ci,k] for k = (tap-1)*s + j + p
 * inside [0, 2s) and 0 elsewhere. */

/* ---- the reference's native op, as a C entry point ------------------------------------------------
 * upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1) of
 * flowdec/backbones/ncsnpp_utils/op/upfirdn2d.cpp:38-48 (kernel op/upfirdn2d_kernel.cu:118-218, semantics
 * op/upfirdn2d.py:182-224): fp32 NCHW input [N,C,in_h,in_w] passed as planes = N*C contiguous images, fp32
 * kernel [kh,kw]; zero-insertion upsampling, padding (negative = cropping), correlation with the flipped
 * kernel, decimation.  out fp32 [planes, out_h, out_w] with
 *   out_h = (in_h*up_y + pad_y0 + pad_y1 - kh) / down_y + 1,  out_w likewise — allocated by the caller
 * (the reference op allocates it itself, upfirdn2d_kernel.cu:253-254; see INTEGRATION.md §2 for the binding).
 * The fused backbone path does not call this (it uses fd_gn_act_resample on bf16 NHWC); it is the drop-in for
 * `upfirdn2d_op.upfirdn2d` so op/upfirdn2d.py:169-180 can bind this library unchanged. */
int fd_upfirdn2d_f32(const float* input, int planes, int in_h, int in_w, const float* kernel, int kh, int kw,
                     int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                     float* out, fd_stream_t stream);

/* ---- tf32 "precise" mode of the backbone ------------------------------------------------------------
 * fd_conv2d_igemm with flags bit 2 (value 4): sources, weights and output are fp32 (weights rounded to tf32 at
 * pack time), MMAs are kind::tf32 with fp32 accumulation — the precision class of the reference's own GPU convs
 * (cuDNN with allow_tf32, layers.py:110-134).  Needs the halo tiles (flags 7): c_count % 32 == 0, W % 8 == 0,
 * H % 16 == 0, even tile count; out_is_f32 = 1 with cout == npad in {128,256}, or the 4-channel pyramid form
 * (npad 16, cout 4) which also exists for bf16 sources.  The `_f32` entry points below are the same kernels as
 * their namesakes on fp32 NHWC activations. */
int fd_chan_stats_f32(const void* x_f32, int B, int HW, int C, float* partial, int S, fd_stream_t stream);
int fd_gn_act_resample_f32(const void* src1, int C1, const void* src2, int C2, const float* scale_shift,
                           void* out, void* out_raw, int B, int H, int W, int mode, fd_stream_t stream);
int fd_conv_in_f32(const void* in4, const float* w, const float* bias, void* out, int B, int H, int W,
                   fd_stream_t stream);
int fd_combine_f32(const void* pyr4, const float* w, const float* bias, const void* h, void* out, size_t npix, int C,
                   fd_stream_t stream);

/* ---- shape-generic kernels (7-level / bottleneck-attention NCSN++) ------------------
 * fd_conv2d_direct: same contract and packed weights as fd_conv2d_igemm, for shapes the tcgen05 tiles do not
 * take: any H, W; segment channel counts multiples of 8; cout rows of wpacked; out bf16 or fp32 NHWC with
 * channel pitch out_pitch >= cout.  flags bit 0: sources with scale_shift get the affine WITHOUT SiLU
 * (AttnBlockpp's GroupNorm, layerspp.py:86).  CUDA-core fp32 accumulation of bf16 operands. */
int fd_conv2d_direct(const struct fd_conv_src* srcs, int nsrc, const void* wpacked, int ktot, const float* bias,
                     void* out, int out_is_f32, int cout, int out_pitch, int B, int H, int W, int flags,
                     fd_stream_t stream);
/* AttnBlockpp core (layerspp.py:91-96): qkv fp32 [B,T,3C] (q | k | v per token, T = H*W) ->
 * out bf16 [B,T,C] = softmax_j(q_i . k_j * scale) v_j */
int fd_attention(const float* qkv, int B, int T, int C, float scale, void* out, fd_stream_t stream);
/* fd_gn_act_resample mode 1 for any even H, W (the tiled / patch kernels need multiples of 4) */
int fd_gn_act_down_any(const void* src1, int C1, const void* src2, int C2, const float* scale_shift, void* out,
                       void* out_raw, int B, int H, int W, fd_stream_t stream);
/* fd_conv_in for cout != 64 (w fp32 OIHW [cout,4,3,3]) */
int fd_conv_in_any(const void* in4, const float* w, const float* bias, void* out, int B, int H, int W, int cout,
                   fd_stream_t stream);
/* fd_output_axpy with a 3x3 output layer (ncsnpp.py:100, output_layer_kwargs.kernel_size = 3):
 * w72 = DEVICE fp32 [2,4,3,3]; v = Conv3x3(pyr), out = c1*base1 + c2*base2 + c3*base3 + coef*v */
int fd_output_conv3_axpy(const void* pyr4, const float* w72, const void* base1, float c1, const void* base2,
                         float c2, const void* base3, float c3, float coef, void* out, void* v_out, int B, int H,
                         int W, fd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOWDEC_B200_H */
