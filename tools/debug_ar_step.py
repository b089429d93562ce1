#!/usr/bin/env python
"""Stage-by-stage error report of one CUDA AR step against the CPU oracle (debug aid, GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import flowtron_oracle as O
from flowtron_b200 import synth
from flowtron_b200 import _lib
from flowtron_b200.flowtron import Flowtron


def main():
    n_flows = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    B, T, L = 3, 40, 14
    cfg = dict(synth.DEFAULT_MODEL_CONFIG, n_flows=n_flows)
    p = synth.synth_params(cfg, 1)
    model = Flowtron(**cfg); model.load_state_dict(p, strict=True); model = model.cuda().eval()
    batch = synth.synth_batch(B, T, L, cfg, 1, out_lens=[40, 13, 29], with_prior=True)
    cu = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    with torch.no_grad():
        spk = model.speaker_embedding(cu["speaker_ids"])
        t = model.encoder(model.embedding(cu["text"]).transpose(1, 2), cu["in_lens"]).transpose(0, 1)
        enc = torch.cat([t, spk.expand(t.size(0), -1, -1)], 2).contiguous()
    enc_cpu = enc.cpu()
    mel = cu["mel"].permute(2, 0, 1).contiguous()
    mask = ~O.get_mask_from_lengths(batch["in_lens"], L)[..., None]
    # oracle intermediates for flow 0
    pre = "flows.0"
    mel_cpu = mel.cpu()
    mel0 = torch.cat([torch.zeros_like(mel_cpu[:1]), mel_cpu[:-1]], 0)
    hA = O._zero_after_len(O._lstm(mel0, p, f"{pre}.attention_lstm", 0, False), batch["out_lens"])
    ctx, attn, lp = O.attention_forward(p, f"{pre}.attention_layer", hA, enc_cpu, mask, batch["attn_prior"])
    ctx = ctx.permute(2, 0, 1)
    d = torch.cat((hA, ctx), -1)
    h0 = O._zero_after_len(O._lstm(d, p, f"{pre}.lstm", 0, False), batch["out_lens"])
    h1 = O._zero_after_len(O._lstm(h0, p, f"{pre}.lstm", 1, False), batch["out_lens"])
    o = O._dense_conv(p, pre, h1)
    Q = hA @ p[f"{pre}.attention_layer.query.linear_layer.weight"].t()
    Kp = enc_cpu @ p[f"{pre}.attention_layer.key.linear_layer.weight"].t()
    flow = model.flows[0]
    from flowtron_b200.flowtron import _ArStepFn, _lens_i32
    in_lens = (~mask[..., 0]).sum(1).to(torch.int32).cuda()
    out_lens = _lens_i32(cu["out_lens"], mel.device)
    with torch.no_grad():
        outs = flow(mel, enc, mask.cuda(), cu["out_lens"], cu["attn_prior"])
    torch.cuda.synchronize()
    print("status", _lib.device_status())
    # re-run through the Function to get at the saved buffer
    class Ctx: pass
    desc = _lib.FtArStepDesc(T, B, L, 80, 1024, 640, 640, 0, int(n_flows == 1), 1, 1.0)
    saved_b, scr_b = _lib.ar_step_sizes(desc)
    saved = torch.zeros(saved_b, dtype=torch.uint8, device="cuda")
    scratch = _lib.scratch_buffer(scr_b, mel.device)
    w = _lib.make_weights([x.detach() if x is not None else None for x in flow._param_list()])
    mo, ls = torch.empty(T, B, 80, device="cuda"), torch.empty(T, B, 80, device="cuda")
    g = torch.empty(T, B, 1, device="cuda") if n_flows == 1 else None
    at, lpp = torch.empty(B, T, L, device="cuda"), torch.empty(B, T, L, device="cuda")
    _lib.ar_step_fwd(desc, w, mel, enc, in_lens, out_lens, cu["attn_prior"].contiguous(), mo, ls, g, at, lpp, saved, scratch)
    torch.cuda.synchronize()
    vm = (torch.arange(T)[:, None] < batch["out_lens"][None, :])
    def rep(name, mine, ref, m=None):
        mine, ref = mine.float().cpu(), ref.float()
        if m is not None:
            mine, ref = mine[m], ref[m]
        e = (mine - ref).abs().max().item(); s = ref.abs().max().item()
        print(f"{name:10s} abs {e:.3e}  scale {s:.3e}  rel {e / max(s, 1e-12):.3e}")
    d16 = _lib.ar_step_saved_view(desc, saved, "d16", torch.float16).view(T, B, 1664)
    rep("hA", d16[:, :, :1024], hA, vm)
    rep("Q", _lib.ar_step_saved_view(desc, saved, "Q", torch.float32).view(T, B, 640), Q, vm)
    rep("Kp", _lib.ar_step_saved_view(desc, saved, "Kp", torch.float32).view(L, B, 640), Kp)
    rep("attn", at, attn, vm.t())
    rep("logprob", lpp, lp, vm.t())
    rep("ctx", d16[:, :, 1024:], ctx, vm)
    rep("h0", _lib.ar_step_saved_view(desc, saved, "h0_16", torch.float16).view(T, B, 1024), h0, vm)
    rep("h1", _lib.ar_step_saved_view(desc, saved, "h1_16", torch.float16).view(T, B, 1024), h1, vm)
    rep("o", _lib.ar_step_saved_view(desc, saved, "o32", torch.float32).view(T, B, 160), o, vm)
    rep("log_s", ls, o[:, :, :80], vm)
    ref_z = torch.exp(o[:, :, :80]) * mel_cpu + o[:, :, 80:]
    rep("z", mo, ref_z, vm)

if __name__ == "__main__":
    main()
