"""TEST INFRASTRUCTURE — parity metrics, restated from the reference's evaluation code
(/root/reference/flowdec/eval/metrics.py).  Only tests/ and bench.py's checker legs import this."""
import torch


def snr_db(x_hat, x):
    """waveform SNR 10 log10(|x|^2 / |x - x_hat|^2)"""
    x_hat, x = x_hat.double().reshape(-1), x.double().reshape(-1)
    return float(10 * torch.log10(x.pow(2).sum() / (x - x_hat).pow(2).sum()))


def si_sdr_db(x_hat, x):
    """SI-SDR as eval/metrics.py:256-270 + 554-563 computes it with a zero noise reference
    (s_target = <x_hat, x>/|x|^2 x ; SI-SDR = |s_target|^2 / |x_hat - s_target|^2)"""
    x_hat, x = x_hat.double().reshape(-1), x.double().reshape(-1)
    s_target = torch.dot(x_hat, x) / x.pow(2).sum() * x
    return float(10 * torch.log10(s_target.pow(2).sum() / (x_hat - s_target).pow(2).sum()))


def logspec_mse(x_hat, x, sr=48000, win_dur=32e-3, hop_dur=8e-3, eps=1e-8):
    """LogSpecMSE.forward (eval/metrics.py:333-372): power spectrograms (Hann window of 32 ms, hop 8 ms, centred,
    torchaudio Spectrogram(power=2) == |stft|^2 with reflect padding), 10 log10 clamp(., 1e-8), mean squared difference"""
    n_fft, hop = int(win_dur * sr), int(hop_dur * sr)
    win = torch.signal.windows.hann(n_fft)       # the reference passes torch.signal.windows.hann (symmetric)

    def spec(v):
        return torch.stft(v.float().reshape(-1), n_fft, hop_length=hop, win_length=n_fft, window=win, center=True,
                          pad_mode="reflect", return_complex=True).abs().pow(2)
    a = 10 * torch.log10(torch.clamp(spec(x_hat), min=eps))
    b = 10 * torch.log10(torch.clamp(spec(x), min=eps))
    return float(torch.mean(torch.square(a - b)))
