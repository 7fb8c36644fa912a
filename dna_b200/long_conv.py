"""Drop-in for the reference's `src/models/sequence/long_conv.py` + `long_conv_kernel.py` (registry name "long-conv",
src/utils/registry.py:49): an explicit (learned, soft-thresholded) long convolution kernel per channel, the long
convolution with skip on the hand-written sm_100a kernels, then activation / dropout / position-wise output Linear.

Same constructor keywords, parameter names (`D`, `kernel.kernel`, `output_linear.0.{weight,bias}`) and forward
semantics as the reference (long_conv.py:19-175):
    k = soft_threshold(kernel)                                   long_conv_kernel.py:68-81
    bidirectional:  k = pad(k0, (0, L)) + pad(flip(k1), (L, 0))  long_conv.py:129-132   (circular over L_kernel + L)
    y = irfft(rfft(u, n) * rfft(k, n))[..., :L],  n = L_kernel + L     long_conv.py:143-146 -> the causal kernels
    y = y + u * D ; flatten channels ; activation ; dropout ; output_linear        long_conv.py:148-161
CUDA tensors only (no CPU fallback).  `block_fft_conv=True` (the Monarch BlockFFT variant with learnable DFT matrices,
block_fft.py) is not implemented."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .fftconv import FFTConvFunc, circular_conv
from .hyena import Activation, OptimModule


class DropoutNd(nn.Module):
    """Reference: src/models/nn/components.py:68-92 (mask tied across the sequence axes when tie=True)."""

    def __init__(self, p: float = 0.5, tie=True, transposed=True):
        super().__init__()
        if p < 0 or p >= 1:
            raise ValueError("dropout probability has to be in [0, 1), but got {}".format(p))
        self.p, self.tie, self.transposed = p, tie, transposed

    def forward(self, X):
        if not self.training:
            return X
        if not self.transposed:
            X = X.movedim(1, -1)
        mask_shape = X.shape[:2] + (1,) * (X.ndim - 2) if self.tie else X.shape
        mask = torch.rand(*mask_shape, device=X.device) < 1.0 - self.p
        X = X * mask * (1.0 / (1 - self.p))
        if not self.transposed:
            X = X.movedim(-1, 1)
        return X


class TransposedLinear(nn.Module):
    """Linear on the second-to-last axis of [B, D, L] (src/models/nn/components.py:200-223)."""

    def __init__(self, d_input, d_output, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(d_output, d_input))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            self.bias = nn.Parameter(torch.empty(d_output))
            bound = 1 / math.sqrt(d_input)
            nn.init.uniform_(self.bias, -bound, bound)
            setattr(self.bias, "_optim", {"weight_decay": 0.0})
        else:
            self.bias = 0.0

    def forward(self, x):
        y = torch.matmul(self.weight.to(x.dtype), x)
        if isinstance(self.bias, torch.Tensor):
            y = y + self.bias.to(x.dtype).view(-1, *[1] * (x.dim() - 2))
        return y


def LinearActivation(d_input, d_output, bias=True, zero_bias_init=False, transposed=False, initializer=None,
                     activation=None, activate=False, weight_norm=False, **kwargs):
    """src/models/nn/components.py:146-179."""
    linear_cls = TransposedLinear if transposed else nn.Linear
    if activation == "glu":
        d_output *= 2
    linear = linear_cls(d_input, d_output, bias=bias, **kwargs)
    if initializer is not None:
        gain = {"uniform": nn.init.kaiming_uniform_, "normal": nn.init.kaiming_normal_}.get(initializer)
        if gain is not None:
            nl = "linear" if activation in (None, "id", "identity", "linear", "glu") else (
                activation if activation in ("relu", "tanh", "sigmoid") else "relu")
            gain(linear.weight, nonlinearity=nl)
        elif initializer == "xavier":
            nn.init.xavier_normal_(linear.weight)
        elif initializer == "zero":
            nn.init.constant_(linear.weight, 0)
        elif initializer == "one":
            nn.init.constant_(linear.weight, 1)
        else:
            raise NotImplementedError(f"get_initializer: initializer type {initializer} not supported")
    if bias and zero_bias_init:
        nn.init.zeros_(linear.bias)
    if weight_norm:
        linear = nn.utils.weight_norm(linear)
    if activate and activation is not None:
        linear = nn.Sequential(linear, Activation(activation, d_output, dim=1 if transposed else -1))
    return linear


class LongConvKernel(OptimModule):
    """Reference: src/models/sequence/long_conv_kernel.py:8-85 — the kernel is a plain [channels, H, L] parameter,
    squashed by a soft threshold `relu(|k| - lam) * sign(k)` every forward (optionally moving-average smoothed)."""

    def __init__(self, H, L, channels=1, learning_rate=None, lam=0.1, causal=True, kernel_dropout=0, weight_init="random",
                 use_ma_smoothing=False, ma_window_len=7, smooth_freq=False, **kwargs):
        super().__init__()
        self.drop = nn.Dropout(p=kernel_dropout)
        self.H, self.weight_init, self.causal = H, weight_init, causal
        self.L = L * 2 if not causal else L
        self.channels, self.lam = channels, lam
        self.register("kernel", self._parameter_initialization(), learning_rate)
        self.use_ma_smoothing, self.smooth_freq, self.ma_window_len = use_ma_smoothing, smooth_freq, ma_window_len
        if use_ma_smoothing:
            if smooth_freq:
                raise NotImplementedError("LongConvKernel(smooth_freq=True) is not implemented")
            assert ma_window_len % 2 != 0, "window size must be odd"
            self.smooth = nn.AvgPool1d(kernel_size=ma_window_len, stride=1, padding=ma_window_len // 2)

    def _parameter_initialization(self):
        if self.weight_init == "random":
            return torch.randn(self.channels, self.H, self.L) * 0.002
        if self.weight_init == "double_exp":
            K = torch.randn(self.channels, self.H, self.L, dtype=torch.float32) * 0.02
            i = torch.arange(self.H, dtype=torch.float32)[:, None] / self.H
            j = torch.arange(self.L, dtype=torch.float32)[None, :] / self.L
            double_exp = torch.exp(-j * torch.pow(torch.tensor(float(int(self.H / 2))), i))
            return K * double_exp[None]
        raise NotImplementedError(f"{self.weight_init} is not valid")

    def forward(self, **kwargs):
        k = self.kernel
        if self.use_ma_smoothing:
            k = self.smooth(k)
        k = F.relu(torch.abs(k) - self.lam) * torch.sign(k)
        return self.drop(k), None

    @property
    def d_output(self):
        return self.H


class LongConv(nn.Module):
    """Reference: src/models/sequence/long_conv.py:19-175."""

    def __init__(self, d_model, l_max=1024, channels=1, bidirectional=False, activation="gelu", postact="glu",
                 initializer=None, weight_norm=False, dropout=0.0, tie_dropout=False, transposed=True, verbose=False,
                 block_fft_conv=False, block_fft_conv_args={}, **kernel_args):
        super().__init__()
        if block_fft_conv:
            raise NotImplementedError("hyena-b200 LongConv: block_fft_conv=True (learnable Monarch BlockFFT) is not implemented")
        self.d_model = self.H = d_model
        self.L = l_max
        self.bidirectional, self.channels, self.transposed = bidirectional, channels, transposed
        self.D = nn.Parameter(torch.randn(channels, self.H))
        if bidirectional:
            channels *= 2
        kernel_args = {k: v for k, v in kernel_args.items() if k not in ("layer_idx", "device", "dtype")}
        self.kernel = LongConvKernel(self.H, L=self.L, channels=channels, verbose=verbose, **kernel_args)
        self.activation = Activation(activation)
        dropout_fn = DropoutNd if tie_dropout else nn.Dropout
        self.dropout = dropout_fn(dropout) if dropout > 0.0 else nn.Identity()
        if postact is None:
            self.output_linear = nn.Identity()
        else:
            self.output_linear = LinearActivation(self.d_model * self.channels, self.d_model, transposed=self.transposed,
                                                  initializer=initializer, activation=postact, activate=True,
                                                  weight_norm=weight_norm)

    def forward(self, u, state=None, rate=1.0, lengths=None, **kwargs):
        """u: [B, H, L] if transposed else [B, L, H]; returns (y of the same layout, None)."""
        if not self.transposed:
            u = u.transpose(-1, -2)
        L = u.size(-1)
        if isinstance(lengths, int):
            lengths = torch.tensor(lengths, dtype=torch.long, device=u.device) if lengths != L else None
        if lengths is not None:
            assert isinstance(lengths, torch.Tensor) and lengths.ndim == 1 and lengths.size(0) in [1, u.size(0)]
            mask = torch.where(torch.arange(L, device=lengths.device) < lengths[:, None, None], 1.0, 0.0)
            u = u * mask.to(u.dtype)
        if rate != 1.0:
            raise NotImplementedError("hyena-b200 LongConv: rate != 1 is not implemented")
        k, _ = self.kernel(L=L if self.L is None else min(L, round(self.L / rate)), rate=rate, state=state)   # [C, H, Lk]
        Lk = k.shape[-1]
        B, H = u.shape[0], u.shape[1]
        C = self.channels
        u = u.contiguous()
        uc = u.unsqueeze(1).expand(B, C, H, L).reshape(B, C * H, L)
        if Lk <= L:
            # transform length L_kernel + L = Lk + L: nothing wraps
            if self.bidirectional:
                k0, k1 = k[:C], k[C:]
                # causal part + the flipped kernel placed at the END of the circular kernel: tap k1[j] reads u[t + 1 + j]
                # == the causal kernel pad(k1, (1, 0)) applied to the time-reversed sequence
                y = self._causal(uc, k0.reshape(C * H, Lk), L)
                y = y + self._causal(uc.flip(-1), F.pad(k1.reshape(C * H, Lk), (1, 0)), L).flip(-1)
            else:
                y = self._causal(uc, k.reshape(C * H, Lk), L)
        else:
            # L < l_max: the reference still transforms at n = L_kernel + L = 2L (long_conv.py:123,143-146) while the kernel
            # module returns all l_max taps -> rfft(k, n) crops the kernel to n taps and the convolution is circular
            n = 2 * L
            if self.bidirectional:
                k0, k1 = k[:C], k[C:]
                kk = F.pad(k0, (0, L)) + F.pad(k1.flip(-1), (L, 0))
            else:
                kk = k
            kk = kk[..., :n]
            if kk.shape[-1] < n:
                kk = F.pad(kk, (0, n - kk.shape[-1]))
            y = circular_conv(uc, kk.reshape(C * H, n), L)
        y = y + uc.float() * self.D.reshape(1, C * H, 1).float()
        y = y.to(u.dtype)                                          # '... c h l -> ... (c h) l'
        if not self.transposed:
            y = y.transpose(-1, -2)
        y = self.activation(y)
        y = self.dropout(y)
        y = self.output_linear(y)
        return y, None

    @staticmethod
    def _causal(u, k, L):
        """first L outputs of the linear convolution of u [B, R, L] with k [R, Lk] — the kernels run at length
        max(L, Lk) and the (zero-padded) tail is dropped"""
        Lk = k.shape[-1]
        n = max(L, Lk)
        up = F.pad(u, (0, n - L)) if n > L else u
        kp = F.pad(k, (0, n - Lk)) if n > Lk else k
        zero = torch.zeros(kp.shape[0], dtype=torch.float32, device=kp.device)
        return FFTConvFunc.apply(up, kp, zero, None, False)[..., :L].float()

    @property
    def d_state(self):
        return self.H

    @property
    def d_output(self):
        return self.d_model
