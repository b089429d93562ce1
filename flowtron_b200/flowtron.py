"""Host-side mirror of the reference's model API for the AR-flow hot path.

Same class names, constructor signatures, attribute names and ``state_dict`` layout as
/root/reference/flowtron.py (Flowtron :831-961, AR_Step :645-828, AR_Back_Step :595-642,
Attention :528-592, DenseLayer :453-464, LinearNorm :278-288, Encoder :467-525, FlowtronLoss :185-275), so a
reference checkpoint loads with ``strict=True`` and callers switch by changing the import.

The flow steps, the NLL/gate loss and inference run on hand-written sm_100a kernels through the C ABI
(include/flowtron_b200.h).  There is NO torch/CPU fallback for that path: CPU tensors raise.  The Encoder and
the embeddings (run once per utterance over L tokens, <2% of FLOPs; SURVEY.md §8f "next") stay ordinary
PyTorch modules, written device-agnostically.
"""
from __future__ import annotations

import os
import warnings
from typing import Optional

import torch
from torch import nn, Tensor
from torch.nn import functional as F

from . import _lib
from ._lib import FlowtronB200Error


def get_mask_from_lengths(lengths: Tensor) -> Tensor:
    """flowtron.py:39-50 (device-agnostic: the reference hard-codes torch.cuda.LongTensor)."""
    max_len = int(torch.max(lengths).item())
    ids = torch.arange(0, max_len, device=lengths.device, dtype=lengths.dtype)
    return (ids < lengths.unsqueeze(1)).bool()


get_gate_mask_from_lengths = get_mask_from_lengths


# --------------------------------------------------------------------------------------------- encoder (torch)
def masked_instance_norm(input, mask, weight, bias, eps: float = 1e-5):
    """flowtron.py:53-92 with use_input_stats=True, no running stats."""
    lengths = mask.sum((-1,))
    mean = (input * mask).sum((-1,)) / lengths
    var = (((input - mean[..., None]) * mask) ** 2).sum((-1,)) / lengths
    out = (input - mean[..., None]) / torch.sqrt(var[..., None] + eps)
    return out * weight[None, :, None] + bias[None, :, None]


class MaskedInstanceNorm1d(nn.InstanceNorm1d):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=False, track_running_stats=False):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)

    def forward(self, input, mask=None):
        if mask is None:
            return F.instance_norm(input, self.running_mean, self.running_var, self.weight, self.bias,
                                   self.training or not self.track_running_stats, self.momentum, self.eps)
        return masked_instance_norm(input, mask.to(input.dtype), self.weight, self.bias, self.eps)


class LinearNorm(nn.Module):
    def __init__(self, in_dim, out_dim, bias=True, w_init_gain='linear'):
        super().__init__()
        self.linear_layer = nn.Linear(in_dim, out_dim, bias=bias)
        nn.init.xavier_uniform_(self.linear_layer.weight, gain=nn.init.calculate_gain(w_init_gain))

    def forward(self, x):
        return self.linear_layer(x)


class ConvNorm(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=None, dilation=1, bias=True,
                 w_init_gain='linear'):
        super().__init__()
        if padding is None:
            assert kernel_size % 2 == 1
            padding = int(dilation * (kernel_size - 1) / 2)
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, bias=bias)
        nn.init.xavier_uniform_(self.conv.weight, gain=nn.init.calculate_gain(w_init_gain))

    def forward(self, signal):
        return self.conv(signal)


class DenseLayer(nn.Module):
    def __init__(self, in_dim=1024, sizes=[1024, 1024]):
        super().__init__()
        in_sizes = [in_dim] + sizes[:-1]
        self.layers = nn.ModuleList([LinearNorm(i, o, bias=True) for (i, o) in zip(in_sizes, sizes)])

    def forward(self, x):
        for linear in self.layers:
            x = torch.tanh(linear(x))
        return x


_enc_streams = {}


def _encoder_overlap_stream(device):
    key = ("overlap", device.type, device.index)
    if key not in _enc_streams:
        _enc_streams[key] = torch.cuda.Stream(device=device)
    return _enc_streams[key]


def _encoder_side_stream(device):
    key = (device.type, device.index)
    if key not in _enc_streams:
        _enc_streams[key] = torch.cuda.Stream(device=device)
    return _enc_streams[key]


class _EncoderFn(torch.autograd.Function):
    """Encoder.forward / Encoder.infer on the CUDA kernels (ft_encoder_fwd / ft_encoder_bwd)."""

    @staticmethod
    def forward(ctx, enc, x, in_lens, masked, dropout_p, *flat_params):
        B, C, L = x.shape
        names = [n for n, _ in _lib.ENC_PARAM_FIELDS]
        counts = [k for _, k in _lib.ENC_PARAM_FIELDS]
        params, i = {}, 0
        for n, k in zip(names, counts):
            params[n] = [t.detach().contiguous() for t in flat_params[i:i + k]]
            i += k
        desc = _lib.encoder_desc(B, L, masked, dropout_p, enc.convolutions[0][1].eps)
        lens32 = in_lens.to(torch.int32) if masked else None
        x = x.detach().contiguous().float()
        out = torch.empty(B, L, C, dtype=torch.float32, device=x.device)
        rng = enc._rng_state(x.device) if dropout_p > 0 else None
        saved = _lib.encoder_fwd(desc, params, x, lens32, rng, out, out.stride(0), out.stride(1))
        ctx.desc, ctx.params, ctx.saved = desc, params, saved
        ctx.need_dx = x.requires_grad or ctx.needs_input_grad[1]
        return out

    @staticmethod
    def backward(ctx, d_out):
        if ctx.saved is None:
            raise RuntimeError("Encoder backward called twice (the saved activations were released)")
        d_x, grads = _lib.encoder_bwd(ctx.desc, ctx.params, d_out.contiguous().float(), ctx.saved, need_dx=ctx.need_dx)
        ctx.saved = None
        flat = []
        for n, _ in _lib.ENC_PARAM_FIELDS:
            flat += grads[n]
        return (None, d_x, None, None, None, *flat)


class _EncoderTokensFn(torch.autograd.Function):
    """embedding(text) + Encoder.forward / .infer on the CUDA kernels with the row gather fused into the first kernel and the
    embedding gradient scattered by the last one (ft_encoder_fwd_tokens / ft_encoder_bwd_tokens)."""

    @staticmethod
    def forward(ctx, enc, tokens, emb_weight, in_lens, masked, dropout_p, *flat_params):
        B, L = tokens.shape
        params, i = {}, 0
        for n, k in _lib.ENC_PARAM_FIELDS:
            params[n] = [t.detach().contiguous() for t in flat_params[i:i + k]]
            i += k
        desc = _lib.encoder_desc(B, L, masked, dropout_p, enc.convolutions[0][1].eps)
        lens32 = in_lens.to(torch.int32) if masked else None
        tokens = tokens.contiguous()
        embw = emb_weight.detach().contiguous()
        out = torch.empty(B, L, embw.size(1), dtype=torch.float32, device=tokens.device)
        rng = enc._rng_state(tokens.device) if dropout_p > 0 else None
        saved = _lib.encoder_fwd(desc, params, None, lens32, rng, out, out.stride(0), out.stride(1), tokens=tokens, emb_weight=embw)
        ctx.desc, ctx.params, ctx.saved, ctx.tokens, ctx.embw = desc, params, saved, tokens, embw
        return out

    @staticmethod
    def backward(ctx, d_out):
        if ctx.saved is None:
            raise RuntimeError("Encoder backward called twice (the saved activations were released)")
        d_emb, grads = _lib.encoder_bwd(ctx.desc, ctx.params, d_out.contiguous().float(), ctx.saved, tokens=ctx.tokens, emb_weight=ctx.embw)
        ctx.saved = None
        flat = []
        for n, _ in _lib.ENC_PARAM_FIELDS:
            flat += grads[n]
        return (None, None, d_emb, None, None, None, *flat)


class Encoder(nn.Module):
    """flowtron.py:467-525 (3 x conv+masked instance norm+relu+dropout, packed BiLSTM).  CUDA tensors run on the encoder kernels
    (csrc/encoder.cu); the torch code below is the module's CPU mirror (host-logic tests) and the FT_ENC_KERNELS=0 A/B path."""

    def __init__(self, encoder_n_convolutions=3, encoder_embedding_dim=512, encoder_kernel_size=5, norm_fn=nn.BatchNorm1d):
        super().__init__()
        convolutions = []
        for _ in range(encoder_n_convolutions):
            convolutions.append(nn.Sequential(
                ConvNorm(encoder_embedding_dim, encoder_embedding_dim, kernel_size=encoder_kernel_size, stride=1,
                         padding=int((encoder_kernel_size - 1) / 2), dilation=1, w_init_gain='relu'),
                norm_fn(encoder_embedding_dim, affine=True)))
        self.convolutions = nn.ModuleList(convolutions)
        self.lstm = nn.LSTM(encoder_embedding_dim, int(encoder_embedding_dim / 2), 1, batch_first=True, bidirectional=True)
        self.p_dropout = 0.5        # flowtron.py:502 hard-codes 0.5; exposed so tests can run deterministic train-mode steps
        # FT_ENC_STREAMS=1: run the two LSTM directions concurrently (parity-tested; measured r1: only -0.4 ms/step, so off)
        self.two_streams = os.environ.get("FT_ENC_STREAMS", "0") != "0"
        self.use_kernels = os.environ.get("FT_ENC_KERNELS", "1") != "0"
        self._rng = None

    def _rng_state(self, device):
        """device uint64[2] = {seed, call counter} of the dropout masks (advanced on the device by every training forward)."""
        if self._rng is None or self._rng.device != device:
            self._rng = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=device)
        return self._rng

    def _kernel_params(self):
        cs = self.convolutions
        flat = [c[0].conv.weight for c in cs] + [c[0].conv.bias for c in cs] + [c[1].weight for c in cs] + [c[1].bias for c in cs]
        for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
            flat += [getattr(self.lstm, n), getattr(self.lstm, n + "_reverse")]
        return flat

    def forward_tokens(self, tokens, embedding, in_lens):
        """Encoder.forward(embedding(tokens).transpose(1, 2), in_lens) with the lookup fused into the encoder kernels."""
        return _EncoderTokensFn.apply(self, tokens, embedding.weight, in_lens, tokens.size(0) > 1,
                                      self.p_dropout if self.training else 0.0, *self._kernel_params())

    def infer_tokens(self, tokens, embedding):
        """Encoder.infer(embedding(tokens).transpose(1, 2)) likewise."""
        return _EncoderTokensFn.apply(self, tokens, embedding.weight, None, False, self.p_dropout if self.training else 0.0,
                                      *self._kernel_params())

    def _tokens_ok(self, tokens, embedding):
        return (self.use_kernels and tokens.is_cuda and embedding.weight.size(1) == 512 and len(self.convolutions) == 3
                and tokens.size(0) <= 64 and embedding.padding_idx is None and embedding.max_norm is None
                and self.convolutions[0][0].conv.kernel_size[0] == 5 and isinstance(self.convolutions[0][1], MaskedInstanceNorm1d))

    def _kernels_ok(self, x):
        return (self.use_kernels and x.is_cuda and x.size(1) == 512 and len(self.convolutions) == 3 and x.size(0) <= 64
                and self.convolutions[0][0].conv.kernel_size[0] == 5 and isinstance(self.convolutions[0][1], MaskedInstanceNorm1d))

    def _lstm_dir(self, x, sfx):
        """One direction of the BiLSTM on a padded [B, L, C] batch."""
        w = [getattr(self.lstm, n + sfx) for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
        h0 = x.new_zeros(1, x.size(0), self.lstm.hidden_size)
        with warnings.catch_warnings():           # "weights are not part of single contiguous chunk": one direction at a time
            warnings.simplefilter("ignore", UserWarning)
            out, _, _ = torch._VF.lstm(x, (h0, h0), w, True, 1, 0.0, self.training, False, True)
        return out

    def _reverse_dir(self, x, rev):
        gidx = rev[..., None].expand(-1, -1, x.size(2))
        out = self._lstm_dir(torch.gather(x, 1, gidx), "_reverse")
        return torch.gather(out, 1, rev[..., None].expand(-1, -1, out.size(2)))

    def forward(self, x, in_lens):
        """Same result as the reference's packed BiLSTM (flowtron.py:505-512) without packing: the forward direction is
        causal, so running it over the padded batch and zeroing t >= len is identical; the reverse direction runs on
        each utterance reversed inside its own length (one gather in, one gather out).  No pack/unpack and no host sync
        (`in_lens.cpu()`); cuDNN still launches its per-time-step fp32 kernels (measured: same GPU time as packed), and
        they are latency-bound, so with ``two_streams`` the two directions run concurrently on two CUDA streams (autograd
        replays each direction's backward on the stream its forward ran on).  Requires the text batch to be padded to
        max(in_lens), which DataCollate guarantees (data.py:200-208)."""
        if self._kernels_ok(x):
            return _EncoderFn.apply(self, x, in_lens, x.size(0) > 1, self.p_dropout if self.training else 0.0, *self._kernel_params())
        Bn, _, L = x.shape
        ar = torch.arange(L, device=x.device)
        valid = ar[None, :] < in_lens[:, None]                                   # [B, L]
        mask = valid.unsqueeze(1) if Bn > 1 else None
        for conv, norm in self.convolutions:
            if mask is not None:
                x = x.masked_fill(~mask, 0.)
            x = F.dropout(F.relu(norm(conv(x), mask=mask)), self.p_dropout, self.training)
        x = x.transpose(1, 2).contiguous()                                        # [B, L, C]
        rev = torch.where(valid, in_lens[:, None] - 1 - ar[None, :], ar[None, :])   # per-utterance time reversal (involution)
        if self.two_streams and x.is_cuda:
            cur = torch.cuda.current_stream(x.device)
            side = _encoder_side_stream(x.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                bwd = self._reverse_dir(x, rev)
            fwd = self._lstm_dir(x, "")
            cur.wait_stream(side)
            bwd.record_stream(cur)                # allocated from the side stream's pool, consumed on the caller's stream
        else:
            fwd = self._lstm_dir(x, "")
            bwd = self._reverse_dir(x, rev)
        return torch.cat([fwd, bwd], -1) * valid[..., None].to(x.dtype)

    def infer(self, x):
        if self._kernels_ok(x):
            return _EncoderFn.apply(self, x, None, False, self.p_dropout if self.training else 0.0, *self._kernel_params())
        for conv in self.convolutions:
            x = F.dropout(F.relu(conv(x)), self.p_dropout, self.training)
        x = x.transpose(1, 2)
        # no self.lstm.flatten_parameters() (flowtron.py:521): on cuDNN it moves the weights into a new buffer (param.set_),
        # which would detach them from a flat-buffer optimizer / all-reduce bucket built earlier
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)
            outputs, _ = self.lstm(x)
        return outputs


# --------------------------------------------------------------------------------------------- attention params
class Attention(nn.Module):
    """Parameter container + mutable ``temperature`` of flowtron.py:528-592.  The math (Q/K/V projections,
    tanh score, masking, softmax, prior posterior, context) runs inside the fused CUDA attention kernels invoked by
    AR_Step; this module is never called on its own by the model."""

    def __init__(self, n_mel_channels=80, n_speaker_dim=128, n_text_channels=512, n_att_channels=128, temperature=1.0):
        super().__init__()
        self.temperature = temperature
        self.query = LinearNorm(n_mel_channels, n_att_channels, bias=False, w_init_gain='tanh')
        self.key = LinearNorm(n_text_channels + n_speaker_dim, n_att_channels, bias=False, w_init_gain='tanh')
        self.value = LinearNorm(n_text_channels + n_speaker_dim, n_att_channels, bias=False, w_init_gain='tanh')
        self.v = LinearNorm(n_att_channels, 1, bias=False, w_init_gain='tanh')
        self.score_mask_value = -float("inf")

    def forward(self, *args, **kwargs):
        raise FlowtronB200Error("Attention runs fused inside AR_Step on the CUDA path; call AR_Step / Flowtron instead")


# --------------------------------------------------------------------------------------------- AR step
def _lens_i32(lens: Optional[Tensor], device) -> Optional[Tensor]:
    if lens is None:
        return None
    return lens.to(device=device, dtype=torch.int32).contiguous()


# Parameters whose gradients are produced by the attention-LSTM half of the backward (ft_ar_step_bwd_attn_lstm: the
# attention LSTM itself, plus lstm layer 0 and the attention query/key/value projections, whose weight-gradient GEMMs are
# deferred so they run underneath that BPTT kernel); indices into AR_Step._param_list() / _lib.AR_WEIGHT_FIELDS.
_ATTN_HALF = (0, 1, 2, 3, 4, 5, 6, 7, 12, 13, 14)
_MAIN_HALF = tuple(i for i in range(24) if i not in _ATTN_HALF)


class _FlowState:
    """What the two autograd nodes of one flow step share between forward and the two backward calls."""
    __slots__ = ("desc", "saved", "plist", "mel_c", "in_lens", "out_lens", "attn", "carry", "grads", "text_shape", "has_gate",
                 "fuse_bwd", "d_mel")

    def clear(self):
        for k in self.__slots__:
            setattr(self, k, None)


class _ArStepAttnLstmFn(torch.autograd.Function):
    """Autograd node A of a flow step: owns `mel` and the _ATTN_HALF parameters.  Its forward does nothing (the whole
    forward pass is ONE C call, issued by node B); it returns a token that node B consumes, so that A.backward runs after
    B.backward.  Splitting the backward in two nodes lets the autograd engine run everything that depends only on d_text
    (the text encoder's backward, on its own stream) concurrently with the attention LSTM's BPTT of the first flow."""

    @staticmethod
    def forward(ctx, state, mel, *params_a):
        ctx.state = state
        return mel.new_zeros(1)

    @staticmethod
    def backward(ctx, g_token):
        st = ctx.state
        if st is not None and st.d_mel is not None:          # node B ran the whole backward in one call (ft_ar_step_bwd)
            d_mel, grads_a = st.d_mel, [st.grads[i] for i in _ATTN_HALF]
            st.clear()
            ctx.state = None
            return (None, d_mel, *grads_a)
        if st is None or st.saved is None or st.carry is None:
            raise FlowtronB200Error("AR_Step backward ran twice: the flow's saved activations are released after the first "
                                    "backward (retain_graph is not supported on the CUDA path)")
        dev = st.mel_c.device
        d_mel = torch.empty_like(st.mel_c)
        _, scratch_bytes = _lib.ar_step_sizes(st.desc)
        scratch = _lib.scratch_buffer(scratch_bytes, dev)
        _lib.ar_step_bwd_attn_lstm(st.desc, _lib.make_weights(st.plist), st.out_lens, d_mel, _lib.make_weights(st.grads),
                                   st.saved, scratch, st.carry)
        grads_a = [st.grads[i] for i in _ATTN_HALF]
        st.clear()                                   # saved activations, carry, the output `attn` -> ctx reference cycle
        ctx.state = None
        return (None, d_mel, *grads_a)


class _ArStepFn(torch.autograd.Function):
    """Autograd node B: ft_ar_step_fwd (the whole forward) / ft_ar_step_bwd_main."""

    @staticmethod
    def forward(ctx, state, step, reversed_flag, token, mel, text, in_lens, out_lens, attn_prior, *params_b):
        if not mel.is_cuda:
            raise FlowtronB200Error("AR_Step needs CUDA tensors: the sm_100a kernels are the only implementation")
        T, B, M = mel.shape
        L, _, E = text.shape
        dev = mel.device
        has_gate = hasattr(step, 'gate_layer')
        desc = _lib.FtArStepDesc(T, B, L, M, step.lstm.hidden_size, step.attention_layer.query.linear_layer.out_features,
                                 E, int(reversed_flag), int(has_gate), int(attn_prior is not None),
                                 float(step.attention_layer.temperature))
        mel_c = mel.detach().float().contiguous()
        text_c = text.detach().float().contiguous()
        prior_c = None if attn_prior is None else attn_prior.detach().float().contiguous()
        plist = [p.detach() if p is not None else None for p in step._param_list()]
        plist = [p if (p is None or p.is_contiguous()) else p.contiguous() for p in plist]
        weights = _lib.make_weights(plist)
        saved_bytes, scratch_bytes = _lib.ar_step_sizes(desc)
        saved = torch.empty(saved_bytes, dtype=torch.uint8, device=dev)
        scratch = _lib.scratch_buffer(scratch_bytes, dev)
        mel_out = torch.empty(T, B, M, device=dev)
        log_s = torch.empty(T, B, M, device=dev)
        gates = torch.empty(T, B, 1, device=dev) if has_gate else None
        attn = torch.empty(B, T, L, device=dev)
        logprob = torch.empty(B, T, L, device=dev)
        ev = getattr(step, "_text_event", None)
        if ev is not None:                        # `text` was produced on another stream: wait inside the flow, not here
            step._text_event = None
            _lib.set_text_ready_event(ev)
        _lib.ar_step_fwd(desc, weights, mel_c, text_c, in_lens, out_lens, prior_c, mel_out, log_s, gates, attn, logprob,
                         saved, scratch)
        state.desc, state.saved, state.plist = desc, saved, plist
        state.mel_c, state.in_lens, state.out_lens, state.attn = mel_c, in_lens, out_lens, attn
        state.text_shape, state.has_gate = text.shape, has_gate
        state.fuse_bwd = bool(getattr(step, "_fuse_bwd", False))
        ctx.state = state
        ctx.has_gate = has_gate
        if has_gate:
            return mel_out, log_s, gates, attn, logprob
        return mel_out, log_s, attn, logprob

    @staticmethod
    def backward(ctx, *grads):
        if ctx.has_gate:
            d_mel_out, d_log_s, d_gates, d_attn, d_lp = grads
        else:
            d_mel_out, d_log_s, d_attn, d_lp = grads
            d_gates = None
        c = lambda g: None if g is None else g.float().contiguous()
        d_mel_out, d_log_s, d_gates, d_attn, d_lp = map(c, (d_mel_out, d_log_s, d_gates, d_attn, d_lp))
        st = ctx.state
        if st is None or st.saved is None:
            raise FlowtronB200Error("AR_Step backward ran twice: the flow's saved activations are released after the first "
                                    "backward (retain_graph is not supported on the CUDA path)")
        desc = st.desc
        dev = st.mel_c.device
        d_text = torch.empty(st.text_shape, device=dev, dtype=torch.float32)
        st.grads = [None if p is None else torch.empty_like(p) for p in st.plist]
        _, scratch_bytes = _lib.ar_step_sizes(desc)
        scratch = _lib.scratch_buffer(scratch_bytes, dev)
        if st.fuse_bwd:
            # nobody waits for this flow's d_text alone (the encoder's backward starts after the FIRST flow's): run the whole
            # backward in one call, which overlaps the attention backward with the attention LSTM's BPTT chunk by chunk
            st.d_mel = torch.empty_like(st.mel_c)
            _lib.ar_step_bwd(desc, _lib.make_weights(st.plist), st.mel_c, st.in_lens, st.out_lens, st.attn, d_mel_out, d_log_s,
                             d_gates, d_attn, d_lp, st.d_mel, d_text, _lib.make_weights(st.grads), st.saved, scratch)
        else:
            st.carry = torch.empty(_lib.ar_step_bwd_carry_bytes(desc), dtype=torch.uint8, device=dev)
            _lib.ar_step_bwd_main(desc, _lib.make_weights(st.plist), st.mel_c, st.in_lens, st.out_lens, st.attn, d_mel_out, d_log_s,
                                  d_gates, d_attn, d_lp, d_text, _lib.make_weights(st.grads), st.saved, scratch, st.carry)
        grads_b = [st.grads[i] for i in _MAIN_HALF]
        ctx.state = None
        # (state, step, reversed, token, mel, text, in_lens, out_lens, prior, *params_b): mel's gradient comes out of node A
        return (None, None, None, torch.zeros(1, device=dev), None, d_text, None, None, None, *grads_b)


class AR_Step(nn.Module):
    """flowtron.py:645-828.  Same parameters / state_dict keys; forward/infer run on the sm_100a kernels."""

    def __init__(self, n_mel_channels, n_speaker_dim, n_text_channels, n_in_channels, n_hidden, n_attn_channels,
                 n_lstm_layers, add_gate, use_cumm_attention):
        super().__init__()
        if use_cumm_attention:
            raise NotImplementedError("use_cumm_attention is outside the B200 hot-path scope (off in every shipped config)")
        if n_lstm_layers != 2 or n_hidden != 1024:
            raise NotImplementedError("the sm_100a kernels are built for n_lstm_layers=2, n_hidden=1024 (config.json:56-57)")
        self.use_cumm_attention = use_cumm_attention
        self.conv = nn.Conv1d(n_hidden, 2 * n_mel_channels, 1)
        self.conv.weight.data = 0.0 * self.conv.weight.data
        self.conv.bias.data = 0.0 * self.conv.bias.data
        self.lstm = nn.LSTM(n_hidden + n_attn_channels, n_hidden, n_lstm_layers)
        self.attention_lstm = nn.LSTM(n_mel_channels, n_hidden)
        self.attention_layer = Attention(n_hidden, n_speaker_dim, n_text_channels, n_attn_channels)
        self.dense_layer = DenseLayer(in_dim=n_hidden, sizes=[n_hidden, n_hidden])
        if add_gate:
            self.gate_threshold = 0.5
            self.gate_layer = LinearNorm(n_hidden + n_attn_channels, 1, bias=True, w_init_gain='sigmoid')

    def _param_list(self):
        """Order of _lib.AR_WEIGHT_FIELDS."""
        al, ls, at, dl = self.attention_lstm, self.lstm, self.attention_layer, self.dense_layer
        gate = getattr(self, 'gate_layer', None)
        return [al.weight_ih_l0, al.weight_hh_l0, al.bias_ih_l0, al.bias_hh_l0,
                ls.weight_ih_l0, ls.weight_hh_l0, ls.bias_ih_l0, ls.bias_hh_l0,
                ls.weight_ih_l1, ls.weight_hh_l1, ls.bias_ih_l1, ls.bias_hh_l1,
                at.query.linear_layer.weight, at.key.linear_layer.weight, at.value.linear_layer.weight,
                at.v.linear_layer.weight,
                dl.layers[0].linear_layer.weight, dl.layers[0].linear_layer.bias,
                dl.layers[1].linear_layer.weight, dl.layers[1].linear_layer.bias,
                self.conv.weight, self.conv.bias,
                None if gate is None else gate.linear_layer.weight, None if gate is None else gate.linear_layer.bias]

    def _run(self, mel, text, mask, out_lens, attn_prior, reversed_flag):
        dev = mel.device
        in_lens = None
        if mask is not None:
            in_lens = (~mask[..., 0]).sum(1).to(torch.int32).contiguous()
        plist = self._param_list()
        state = _FlowState()
        state.clear()
        token = _ArStepAttnLstmFn.apply(state, mel, *[plist[i] for i in _ATTN_HALF])
        out = _ArStepFn.apply(state, self, reversed_flag, token, mel, text, in_lens, _lens_i32(out_lens, dev), attn_prior,
                              *[plist[i] for i in _MAIN_HALF])
        if hasattr(self, 'gate_layer'):
            mel_out, log_s, gates, attn, lp = out
        else:
            (mel_out, log_s, attn, lp), gates = out, None
        return mel_out, log_s, gates, attn, lp

    def forward(self, mel, text, mask, out_lens, attn_prior=None):
        return self._run(mel, text, mask, out_lens, attn_prior, False)

    def infer(self, residual, text, attns, attn_prior=None):
        from .inference import ar_step_infer
        return ar_step_infer(self, residual, text, attns, attn_prior, reversed_flag=False)


class AR_Back_Step(nn.Module):
    """flowtron.py:595-642.  The flip/roll of the reference is an index map inside the kernels (no copies, no host syncs)."""

    def __init__(self, n_mel_channels, n_speaker_dim, n_text_dim, n_in_channels, n_hidden, n_attn_channels, n_lstm_layers,
                 add_gate, use_cumm_attention):
        super().__init__()
        self.ar_step = AR_Step(n_mel_channels, n_speaker_dim, n_text_dim, n_mel_channels + n_speaker_dim, n_hidden,
                               n_attn_channels, n_lstm_layers, add_gate, use_cumm_attention)

    def forward(self, mel, text, mask, out_lens, attn_prior=None):
        if out_lens is None:
            out_lens = torch.full((mel.size(1),), mel.size(0), dtype=torch.long, device=mel.device)
        return self.ar_step._run(mel, text, mask, out_lens, attn_prior, True)

    def infer(self, residual, text, attns, attn_prior=None):
        from .inference import ar_step_infer
        return ar_step_infer(self.ar_step, residual, text, attns, attn_prior, reversed_flag=True)


# --------------------------------------------------------------------------------------------- model
class Flowtron(nn.Module):
    """flowtron.py:831-961 (n_components=0 / no cumulative attention: the variants every shipped config uses)."""

    def __init__(self, n_speakers, n_speaker_dim, n_text, n_text_dim, n_flows, n_mel_channels, n_hidden,
                 n_attn_channels, n_lstm_layers, use_gate_layer, mel_encoder_n_hidden, n_components, fixed_gaussian,
                 mean_scale, dummy_speaker_embedding, use_cumm_attention):
        super().__init__()
        if n_components > 1:
            raise NotImplementedError("the GMM prior (n_components > 1) is outside the B200 hot-path scope")
        norm_fn = MaskedInstanceNorm1d
        self.speaker_embedding = nn.Embedding(n_speakers, n_speaker_dim)
        self.embedding = nn.Embedding(n_text, n_text_dim)
        self.flows = nn.ModuleList()
        self.encoder = Encoder(norm_fn=norm_fn, encoder_embedding_dim=n_text_dim)
        self.dummy_speaker_embedding = dummy_speaker_embedding
        for i in range(n_flows):
            add_gate = True if (i == (n_flows - 1) and use_gate_layer) else False
            cls = AR_Step if i % 2 == 0 else AR_Back_Step
            self.flows.append(cls(n_mel_channels, n_speaker_dim, n_text_dim, n_mel_channels + n_speaker_dim, n_hidden,
                                  n_attn_channels, n_lstm_layers, add_gate, use_cumm_attention))

    # Optional: run two half batches as independent pipelines on two CUDA streams (each recurrence then uses a 64-SM
    # persistent kernel, so both are co-resident).  This does NOT shorten the recurrence chain (a half batch takes as
    # many dependent steps as the full batch); it only lets one half's GEMMs / attention / launch gaps overlap the other
    # half's recurrences.  Measured on B200 (B=32, T=1000): 91.5 -> 89.1 ms/step (+2.7 %).  Off by default.
    n_streams = 1
    # run the Encoder underneath the first flow's attention LSTM (measured r1: 85.5 -> 82.4 ms/step); FT_ENC_OVERLAP=0 disables
    overlap_encoder = os.environ.get("FT_ENC_OVERLAP", "1") != "0"
    min_split_batch = 8

    # Utterances per kernel launch.  The kernels take up to 64, but the fast paths (cluster BPTT, layer pipeline, folded input
    # projection) are B <= 32 paths: a B = 64 batch (configs[2]) runs faster as two consecutive 32-utterance launches than as
    # one 64-utterance launch on the fallback kernels (r2: 118.8 ms/step as one launch).  FT_MAX_KERNEL_BATCH overrides.
    max_kernel_batch = int(os.environ.get("FT_MAX_KERNEL_BATCH", "32"))

    def _run_flows(self, mel, encoder_outputs, mask, out_lens, attn_prior):
        B = mel.size(1)
        if B > self.max_kernel_batch:          # larger batches run as consecutive chunks (utterances are independent)
            ech = getattr(self, "_enc_chunks", None) or {}      # contiguous text slices made on the encoder's stream (forward())
            outs = [self._run_flows(mel[:, b0:b0 + self.max_kernel_batch], ech.get(b0, encoder_outputs[:, b0:b0 + self.max_kernel_batch]),
                                    mask[b0:b0 + self.max_kernel_batch],
                                    None if out_lens is None else out_lens[b0:b0 + self.max_kernel_batch],
                                    None if attn_prior is None else attn_prior[b0:b0 + self.max_kernel_batch])
                    for b0 in range(0, B, self.max_kernel_batch)]
            n_flows = len(self.flows)
            return (torch.cat([o[0] for o in outs], 1), [torch.cat([o[1][i] for o in outs], 1) for i in range(n_flows)],
                    None if outs[0][2] is None else torch.cat([o[2] for o in outs], 1),
                    [torch.cat([o[3][i] for o in outs], 0) for i in range(n_flows)],
                    [torch.cat([o[4][i] for o in outs], 0) for i in range(n_flows)])
        log_s_list, attns_list, attns_logprob_list, gate = [], [], [], None
        for i, flow in enumerate(self.flows):
            # the module whose kernels run first gets the text-ready event (AR_Back_Step wraps its AR_Step as .ar_step)
            getattr(flow, "ar_step", flow)._text_event = getattr(self, "_text_event", None) if i == 0 else None
            # flows after the first: the encoder's backward cannot start before the first flow's d_text exists, so their backward
            # may run as one call (attention backward overlapped with the attention-LSTM BPTT); FT_FUSE_BWD=0 disables
            getattr(flow, "ar_step", flow)._fuse_bwd = (i != 0) and os.environ.get("FT_FUSE_BWD", "0") != "0"
            mel, log_s, gate, attn_out, attn_logprob_out = flow(mel, encoder_outputs, mask, out_lens, attn_prior)
            log_s_list.append(log_s)
            attns_list.append(attn_out)
            attns_logprob_list.append(attn_logprob_out)
        return mel, log_s_list, gate, attns_list, attns_logprob_list

    def _encode(self, speaker_ids, text, in_lens):
        """flowtron.py:871-880: embeddings + Encoder + speaker vector concat -> [L, B, n_text + n_speaker]."""
        speaker_ids = speaker_ids * 0 if self.dummy_speaker_embedding else speaker_ids
        speaker_vecs = self.speaker_embedding(speaker_ids)
        if self.encoder._tokens_ok(text, self.embedding):
            text = self.encoder.forward_tokens(text, self.embedding, in_lens)     # embedding gather fused into the encoder kernels
        else:
            text = self.encoder(self.embedding(text).transpose(1, 2), in_lens)
        text = text.transpose(0, 1)
        return torch.cat([text, speaker_vecs.expand(text.size(0), -1, -1)], 2)

    def forward(self, mel, speaker_ids, text, in_lens, out_lens, attn_prior=None):
        mean, log_var, prob = None, None, None
        n_text = text.size(1)
        B = mel.size(0)
        # The first flow needs the text encoding only after its attention LSTM (T dependent steps that read mel alone), so
        # the encoder's ~700 small latency-bound kernels can run on a second stream underneath it; the flow's C entry
        # point waits for `text_event` right before it first reads the encoding (ft_ar_step_set_text_ready_event).
        overlap = self.overlap_encoder and mel.is_cuda and self.n_streams <= 1 and B <= 64
        self._text_event = None
        if overlap:
            cur = torch.cuda.current_stream(mel.device)
            es = _encoder_overlap_stream(mel.device)
            es.wait_stream(cur)
            self._enc_chunks = None
            with torch.cuda.stream(es):
                encoder_outputs = self._encode(speaker_ids, text, in_lens)
                if B > self.max_kernel_batch:
                    # the flows will run on slices of the batch: make them contiguous HERE, on the encoder's stream, in front of
                    # the event (a .contiguous() issued later on the caller's stream would read the encoding before it exists)
                    mkb = self.max_kernel_batch
                    self._enc_chunks = {b0: encoder_outputs[:, b0:b0 + mkb].contiguous() for b0 in range(0, B, mkb)}
                    for t in self._enc_chunks.values():
                        t.record_stream(cur)
                ev = torch.cuda.Event()
                ev.record(es)
            encoder_outputs.record_stream(cur)
            self._text_event = ev
        else:
            self._enc_chunks = None
            encoder_outputs = self._encode(speaker_ids, text, in_lens)
        mel = mel.permute(2, 0, 1)
        # key-padding mask from the padded text length (== max(in_lens) by the collate contract): no .item() host sync
        mask = ~(torch.arange(n_text, device=in_lens.device)[None, :] < in_lens[:, None])[..., None]
        split = self.n_streams > 1 and mel.is_cuda and B >= self.min_split_batch and out_lens is not None
        if mel.is_cuda:
            _lib.set_lstm_half_sm(split)
        if not split:
            mel, log_s_list, gate, attns_list, attns_logprob_list = self._run_flows(mel, encoder_outputs, mask, out_lens, attn_prior)
            return (mel, log_s_list, gate, attns_list, attns_logprob_list, mean, log_var, prob)

        cur = torch.cuda.current_stream()
        if not hasattr(self, "_side_streams") or self._side_streams[0].device != mel.device:
            self._side_streams = [torch.cuda.Stream(device=mel.device) for _ in range(2)]
        bounds = [(0, B // 2), (B // 2, B)]
        parts = []
        for (b0, b1), s in zip(bounds, self._side_streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                part = self._run_flows(mel[:, b0:b1], encoder_outputs[:, b0:b1], mask[b0:b1], out_lens[b0:b1],
                                       None if attn_prior is None else attn_prior[b0:b1])
            parts.append(part)
        for s in self._side_streams:
            cur.wait_stream(s)

        def cat(ts, dim):
            for t in ts:
                t.record_stream(cur)
            return torch.cat(ts, dim)
        n_flows = len(self.flows)
        mel = cat([p[0] for p in parts], 1)
        log_s_list = [cat([p[1][i] for p in parts], 1) for i in range(n_flows)]
        gate = None if parts[0][2] is None else cat([p[2] for p in parts], 1)
        attns_list = [cat([p[3][i] for p in parts], 0) for i in range(n_flows)]
        attns_logprob_list = [cat([p[4][i] for p in parts], 0) for i in range(n_flows)]
        return (mel, log_s_list, gate, attns_list, attns_logprob_list, mean, log_var, prob)

    def infer(self, residual, speaker_ids, text, temperature=1.0, gate_threshold=0.5, attns=None, attn_prior=None):
        speaker_ids = speaker_ids * 0 if self.dummy_speaker_embedding else speaker_ids
        speaker_vecs = self.speaker_embedding(speaker_ids)
        if self.encoder._tokens_ok(text, self.embedding):
            text = self.encoder.infer_tokens(text, self.embedding)
        else:
            text = self.encoder.infer(self.embedding(text).transpose(1, 2))
        text = text.transpose(0, 1)
        encoder_outputs = torch.cat([text, speaker_vecs.expand(text.size(0), -1, -1)], 2)
        residual = residual.permute(2, 0, 1)
        attention_weights = []
        for i, flow in enumerate(reversed(self.flows)):
            attn = None if attns is None else list(reversed(attns))[i]
            self.set_temperature_and_gate(flow, temperature, gate_threshold)
            residual, attention_weight = flow.infer(residual, encoder_outputs, attn, attn_prior=attn_prior)
            attention_weights.append(attention_weight)
        return residual.permute(1, 2, 0), attention_weights

    @staticmethod
    def set_temperature_and_gate(flow, temperature, gate_threshold):
        flow = flow.ar_step if hasattr(flow, "ar_step") else flow
        flow.attention_layer.temperature = temperature
        if hasattr(flow, 'gate_layer'):
            flow.gate_threshold = gate_threshold


# --------------------------------------------------------------------------------------------- loss
class _NllGateFn(torch.autograd.Function):
    """Masked Gaussian NLL + log-det + gate BCE (flowtron.py:205-243) as two CUDA reductions."""

    @staticmethod
    def forward(ctx, sigma, z, gate_pred, gate_target, out_lens, *log_s_list):
        if not z.is_cuda:
            raise FlowtronB200Error("FlowtronLoss needs CUDA tensors")
        T, B, M = z.shape
        dev = z.device
        zc = z.detach().float().contiguous()
        lsl = [ls.detach().float().contiguous() for ls in log_s_list]
        gp = None if gate_pred is None else gate_pred.detach().float().contiguous()
        gt = None if gate_pred is None else gate_target.detach().float().contiguous()
        lens = out_lens.to(device=dev, dtype=torch.int32).contiguous()
        sums = torch.empty(4, device=dev)
        _lib.nll_reduce(zc, lsl, gp, gt, lens, sums)
        n = sums[3]
        nll = (sums[0] / (2 * sigma * sigma) - sums[1]) / (n * M)
        gate_loss = (sums[2] / n).reshape(1)
        ctx.sigma, ctx.n_flows = sigma, len(lsl)
        ctx.save = (zc, gp, gt, lens, sums, lsl)
        return nll, gate_loss

    @staticmethod
    def backward(ctx, g_nll, g_gate):
        zc, gp, gt, lens, sums, lsl = ctx.save
        dz = torch.empty_like(zc)
        dls = torch.empty_like(zc)
        dgate = None if gp is None else torch.empty_like(gp)
        gn = g_nll.reshape(1).float().contiguous()
        gg = None if g_gate is None else g_gate.reshape(1).float().contiguous()
        _lib.nll_grad(zc, gp, gt, lens, ctx.sigma, sums, gn, gg, dz, dls, dgate)
        return (None, dz, dgate, None, None, *([dls] * ctx.n_flows))


class _AttnCtcFn(torch.autograd.Function):
    """Batched attention-CTC cost per utterance and its gradient in one launch (ft_attn_ctc_loss, csrc/ctc.cu)."""

    @staticmethod
    def forward(ctx, attn_logprob, in_lens, out_lens, time_reversed, blank_logprob):
        if not attn_logprob.is_cuda:
            raise FlowtronB200Error("AttentionCTCLoss needs CUDA tensors: the sm_100a kernel is the only implementation")
        lp = attn_logprob.detach().float().contiguous()
        dev = lp.device
        cost = torch.empty(lp.size(0), device=dev)
        dlp = torch.empty_like(lp)
        _lib.attn_ctc_loss(lp, _lens_i32(in_lens, dev), _lens_i32(out_lens, dev), time_reversed, blank_logprob, cost, dlp)
        ctx.save_for_backward(dlp)
        return cost

    @staticmethod
    def backward(ctx, g_cost):
        (dlp,) = ctx.saved_tensors
        return dlp * g_cost.reshape(-1, 1, 1), None, None, None, None


class AttentionCTCLoss(nn.Module):
    """flowtron.py:155-182.  Same constructor and call signature; the per-utterance Python loop (blank padding, slicing,
    log_softmax, nn.CTCLoss, `.item()` host syncs) is ONE kernel launch for the whole batch.
    ``time_reversed=True`` consumes an AR_Back_Step's attn_logprob in its flipped time directly, replacing the
    reference's in-place roll + flip (flowtron.py:250-256, 266-270)."""

    def __init__(self, blank_logprob=-1):
        super().__init__()
        self.log_softmax = nn.LogSoftmax(dim=3)
        self.blank_logprob = blank_logprob
        self.CTCLoss = nn.CTCLoss(zero_infinity=True)

    def forward(self, attn, in_lens, out_lens, attn_logprob, time_reversed=False):
        assert attn_logprob is not None
        lp = attn_logprob[:, 0] if attn_logprob.dim() == 4 else attn_logprob          # [B, 1, T, L] like the reference
        cost = _AttnCtcFn.apply(lp, in_lens, out_lens, bool(time_reversed), float(self.blank_logprob))
        return cost.sum() / lp.shape[0]


class FlowtronLoss(nn.Module):
    def __init__(self, sigma=1.0, gm_loss=False, gate_loss=True, use_ctc_loss=False, ctc_loss_weight=0.0, blank_logprob=-1):
        super().__init__()
        if gm_loss:
            raise NotImplementedError("gm_loss (GMM prior) is outside the B200 hot-path scope")
        self.sigma = sigma
        self.gm_loss = gm_loss
        self.gate_loss = gate_loss
        self.use_ctc_loss = use_ctc_loss
        self.ctc_loss_weight = ctc_loss_weight
        self.blank_logprob = blank_logprob
        self.attention_loss = AttentionCTCLoss(blank_logprob=self.blank_logprob)

    def forward(self, model_output, gate_target, in_lengths, out_lengths, is_validation=False):
        z, log_s_list, gate_pred, attn_list, attn_logprob_list, mean, log_var, prob = model_output
        use_gate = self.gate_loss > 0 and gate_pred is not None
        loss, gate_loss = _NllGateFn.apply(self.sigma, z, gate_pred if use_gate else None, gate_target, out_lengths,
                                           *log_s_list)
        if not use_gate:
            gate_loss = torch.zeros(1, device=z.device)
        loss_ctc = torch.zeros_like(gate_loss)
        if self.use_ctc_loss:
            for cur_flow_idx, flow_attn in enumerate(attn_list):
                cur = attn_logprob_list[cur_flow_idx]
                # back steps return flipped time; the reference un-rolls / un-flips in place (flowtron.py:250-256) and
                # restores afterwards (:266-270) -- here the kernel indexes the flipped tensor directly
                loss_ctc = loss_ctc + self.attention_loss(flow_attn.unsqueeze(1), in_lengths, out_lengths,
                                                          attn_logprob=cur.unsqueeze(1),
                                                          time_reversed=(cur_flow_idx % 2 != 0))
            loss_ctc = loss_ctc / float(len(attn_list))
        return loss, gate_loss, loss_ctc
