"""ORACLE (test infrastructure, NOT product code) -- CPU / fp32 restatement of the reference training hot path
`tri_mbt_vsltcls` (`--vslt-type TIE --imgtxt-time 1`, swin image encoder, biobert text).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this module,
and only as the checker / CPU baseline -- never as the thing shipped. The product path
(medical_tri_modal_pilot_b200/) calls hand-written sm_100a kernels and has no fallback to this file.

Parity pin: the reference has no tests or golden vectors of its own. This restatement is pinned
against the *reference itself*, imported in the build container from /root/reference by tools/make_golden.py,
which stores inputs/outputs under tests/golden/*.npz; tests/test_oracle_golden.py replays them through this file.

Every function cites the reference code it restates (paths relative to the reference root).
Plain PyTorch fp32 tensor algebra is used (the path is floating point); autograd provides the backward.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class OracleConfig:
    n_layers: int = 6                # --transformer-num-layers (control/config.py:98)
    n_head: int = 4                  # --transformer-num-head (:99)
    d_model: int = 256               # --transformer-dim (:97)
    multiimages: int = 1             # --multiimages (:32); 1 => 3 images per sample (tri_mbt_vsltcls.py:207,227)
    vsltonly: int = 1                # --mbt-only-vslt (:122)
    fusion_startidx: int = 0         # --mbt-fusion-startIdx (:121)
    bottlenecks_n: int = 4           # tri_mbt_vsltcls.py:38
    training: bool = True            # BatchNorm1d batch statistics (fc_list.1)


# ----------------------------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------------------------
def ie_branch(sd, prefix, s):
    """nn.Sequential(Linear(k,256), nn.LayerNorm(256), ReLU)  -- tri_mbt_vsltcls.py:61-76. s: [..., k]."""
    z = F.linear(s, sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"])
    z = F.layer_norm(z, (z.shape[-1],), sd[f"{prefix}.1.weight"], sd[f"{prefix}.1.bias"], 1e-5)
    return torch.relu(z)


def custom_layernorm(z, gamma, beta, eps=1e-6):
    """builder/models/src/transformer/module.py:138-144 -- unbiased std, eps added to std."""
    mean = z.mean(dim=-1, keepdim=True)
    std = z.std(dim=-1, keepdim=True)
    return gamma * ((z - mean) / (std + eps)) + beta


def positional_encoding(d_model, length):
    """module.py:21-32 (PositionalEncoding buffer `pe`)."""
    pe = torch.zeros(length, d_model)
    position = torch.arange(0, length, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def attn_pad_mask(lengths, T, expand_T):
    """utils.py:79-94,116-125 -- mask[b,q,k] = True where key k is padding (k >= lengths[b])."""
    ar = torch.arange(T)
    pad = ar[None, :] >= lengths.to(torch.long)[:, None]          # non_pad_mask[i, len:] = 0 ; .lt(1)
    return pad[:, None, :].expand(-1, expand_T, -1)


def multi_head_attention(sd, prefix, x, mask, n_head):
    """attention.py:65-84 (MultiHeadAttention) + :24-49 (ScaledDotProductAttention). No output projection."""
    B, T, Dm = x.shape
    d_head = Dm // n_head
    q = F.linear(x, sd[f"{prefix}.query_proj.linear.weight"], sd[f"{prefix}.query_proj.linear.bias"])
    k = F.linear(x, sd[f"{prefix}.key_proj.linear.weight"], sd[f"{prefix}.key_proj.linear.bias"])
    v = F.linear(x, sd[f"{prefix}.value_proj.linear.weight"], sd[f"{prefix}.value_proj.linear.bias"])
    # head-major batch [H*B, T, d] (attention.py:72-74)
    q = q.view(B, T, n_head, d_head).permute(2, 0, 1, 3).reshape(n_head * B, T, d_head)
    k = k.view(B, T, n_head, d_head).permute(2, 0, 1, 3).reshape(n_head * B, T, d_head)
    v = v.view(B, T, n_head, d_head).permute(2, 0, 1, 3).reshape(n_head * B, T, d_head)
    score = torch.bmm(q, k.transpose(1, 2)) / math.sqrt(d_head)                       # :35
    if mask is not None:
        score = score.masked_fill(mask.repeat(n_head, 1, 1), -65504.0)                # :38, :77
    attn = torch.softmax(score, -1)                                                   # :41
    ctx = torch.bmm(attn, v)                                                          # :48
    ctx = ctx.view(n_head, B, T, d_head).permute(1, 2, 0, 3).reshape(B, T, Dm)        # :81-82
    return ctx


def feed_forward_conv(sd, prefix, x):
    """module.py:74-80 FeedForwardUseConv: Conv1d(k=1) pair == position-wise linear; dropout p=0 in parity runs."""
    w1 = sd[f"{prefix}.w_1.weight"][:, :, 0]
    w2 = sd[f"{prefix}.w_2.weight"][:, :, 0]
    h = torch.relu(F.linear(x, w1, sd[f"{prefix}.w_1.bias"]))
    return F.linear(h, w2, sd[f"{prefix}.w_2.bias"])


def encoder_layer(sd, prefix, x, mask, n_head):
    """encoder.py:23-34 TransformerEncoderLayer.forward (pre-norm, residuals)."""
    h = custom_layernorm(x, sd[f"{prefix}.attention_prenorm.gamma"], sd[f"{prefix}.attention_prenorm.beta"])
    out = multi_head_attention(sd, f"{prefix}.self_attention", h, mask, n_head) + x
    h2 = custom_layernorm(out, sd[f"{prefix}.feed_forward_prenorm.gamma"], sd[f"{prefix}.feed_forward_prenorm.beta"])
    return feed_forward_conv(sd, f"{prefix}.feed_forward", h2) + out


# ----------------------------------------------------------------------------------------------------------------
# lengths (integer part: must be bit-exact)
# ----------------------------------------------------------------------------------------------------------------
def stream_lengths(input_lengths, txt_lengths, img_time, cfg: OracleConfig):
    """Valid (CLS-inclusive) lengths per stream as the encoder computes them.
    tri_mbt_vsltcls.py:226-237 (img length from the `10` sentinel, txt +2) and mbt_encoder.py:703-707 (+1, ==3 -> 0).
    Returns (len_v [B], len_i [B] or None when the img stream is unmasked, len_t [B])."""
    len_v = input_lengths.to(torch.long) + 1
    if cfg.multiimages == 1:
        it = img_time.reshape(-1, 3) - 10
        len_i = torch.count_nonzero(it, dim=1) * 49
        len_i = len_i.to(torch.int32).to(torch.long) + 1           # .type(torch.IntTensor) :232, += 1
    else:
        len_i = None                                               # mask=[True, False, True] :124-127,144
    len_t = txt_lengths.to(torch.long) + 2 + 1
    len_t = torch.where(len_t == 3, torch.zeros_like(len_t), len_t)
    return len_v, len_i, len_t


def fused_kv_lengths(input_lengths, txt_lengths, img_time, cfg: OracleConfig, T_i, T_t):
    """kv_len[3,B] as used by the fused layers (mbt_encoder.py:748: varying_lengths + bottlenecks_n), clipped to the
    stream length exactly as slicing `non_pad_mask[i, len:] = 0` clips."""
    len_v, len_i, len_t = stream_lengths(input_lengths, txt_lengths, img_time, cfg)
    B = input_lengths.numel()
    kv_v = len_v + cfg.bottlenecks_n
    kv_i = (len_i + cfg.bottlenecks_n) if len_i is not None else torch.full((B,), T_i, dtype=torch.long)
    kv_t = torch.clamp(len_t + cfg.bottlenecks_n, max=T_t)
    return torch.stack([kv_v, torch.clamp(kv_i, max=T_i), kv_t]).to(torch.int32)


def missing_to_num(missing):
    """trainer.py:68-84: rows of missing[B,3] = [0, img_missing, txt_missing] ranked against the 4 canonical rows."""
    sample = torch.tensor([[0., 0., 0.], [0., 0., 1.], [0., 1., 0.], [0., 1., 1.]])
    cat = torch.cat([sample, missing.float()], 0)
    _, inv = torch.unique(cat, dim=0, sorted=True, return_inverse=True)
    return inv[4:].to(torch.long)


# ----------------------------------------------------------------------------------------------------------------
# the model forward (tri_mbt_vsltcls.py:167-263 + mbt_encoder.py:696-784)
# ----------------------------------------------------------------------------------------------------------------
def umse_vslt_embedding(sd, x):
    """tri_mbt_vsltcls.py:183-189 TIE branch. x[B,L,3] = (time, value, feature id as float)."""
    value_embedding = ie_branch(sd, "ie_vslt", x[:, :, 1].unsqueeze(2))
    time_embedding = ie_branch(sd, "ie_time", x[:, :, 0].unsqueeze(2))
    feat = x[:, :, 2].to(torch.int32).long()                                        # .type(torch.IntTensor) :187
    return value_embedding + time_embedding + sd["ie_feat.weight"][feat]


def forward(sd, batch, cfg: OracleConfig, return_aux=False):
    """batch keys: x[B,L,3], age[B], gen[B], input_lengths[B], txts[B,128,768], txt_lengths[B],
    img_feats[B*n_img,49,768] (output of the frozen image encoder, flattened), img_time[B,n_img] or [B], txt_time[B],
    missing[B] int64 codes.  Returns logits [B,1]."""
    x = batch["x"].float()
    B = x.shape[0]
    nb = cfg.bottlenecks_n
    aux = {}
    demographic = torch.stack([batch["age"].float(), batch["gen"].float()], dim=1)      # :176
    demo_embedding = ie_branch(sd, "ie_demo", demographic)                              # :177
    vslt_embedding = umse_vslt_embedding(sd, x)
    aux["vslt_embedding"] = vslt_embedding
    txt_embedding = F.linear(batch["txts"].float(), sd["txt_embedding.weight"], sd["txt_embedding.bias"])   # :200
    img_embedding = F.linear(batch["img_feats"].float(), sd["linear.weight"], sd["linear.bias"])           # :210-211
    img_time = batch["img_time"].float().reshape(-1)                                    # :212
    txt_time = batch["txt_time"].float()
    img_embedding = img_embedding + ie_branch(sd, "ie_time", img_time.unsqueeze(1)).unsqueeze(1) + sd["ie_feat.weight"][18]
    txt_embedding = txt_embedding + ie_branch(sd, "ie_time", txt_time.unsqueeze(1)).unsqueeze(1) + sd["ie_feat.weight"][19]
    if cfg.multiimages == 1:
        img_embedding = img_embedding.reshape(-1, 3, 49, 256).reshape(-1, 147, 256)     # :227-228
    len_v, len_i, len_t = stream_lengths(batch["input_lengths"], batch["txt_lengths"], batch["img_time"].float(), cfg)

    # ---- TrimodalTransformerEncoder_MBT.forward -----------------------------------------------------------
    P = "fusion_transformer"
    streams = [vslt_embedding, img_embedding, txt_embedding]
    enc_inputs = [torch.cat([sd[f"{P}.cls_token_per_modality.{m}"].expand(B, -1, -1), s], 1) for m, s in enumerate(streams)]
    lens = [len_v, len_i, len_t]
    use_pe = [False, False, True]                                                       # tri_mbt_vsltcls.py:60,143
    enc_outputs = []
    for m in range(3):                                                                  # mbt_encoder.py:719-729
        y = F.layer_norm(enc_inputs[m], (256,), sd[f"{P}.layer_norms_in.{m}.weight"], sd[f"{P}.layer_norms_in.{m}.bias"], 1e-5)
        if use_pe[m]:
            y = y + positional_encoding(cfg.d_model, enc_inputs[m].shape[1])
        enc_outputs.append(y)
    bottlenecks = sd[f"{P}.bottlenecks"].expand(B, -1, -1)
    b_masks = []
    for m in range(3):                                                                  # :746-751
        Tm = enc_inputs[m].shape[1] + nb
        b_masks.append(None if lens[m] is None else attn_pad_mask(lens[m] + nb, Tm, Tm))
    aux["b_masks"] = b_masks
    idx_order = torch.arange(B)
    missing = batch["missing"].to(torch.long)
    for l in range(cfg.n_layers):                                                       # :731-779 (fusion_idx = 0)
        enc_in = list(enc_outputs)
        enc_outputs = []
        bott_out = []
        last = cfg.vsltonly == 1 and cfg.n_layers == l + 1
        for m in range(3):
            xb = torch.cat([bottlenecks, enc_in[m]], 1)                                  # :745
            y = encoder_layer(sd, f"{P}.layer_stacks.{l}.{m}", xb, b_masks[m], cfg.n_head)
            bott_out.append(y[:, :nb])
            enc_outputs.append(y[:, nb:])
            if last:
                break
        if last:
            break
        st = torch.stack(bott_out)                                                      # :764
        tri = st.mean(0)
        vt = torch.stack([st[0], st[2]]).mean(0)
        vi = st[:2].mean(0)
        allb = torch.stack([tri, vi, vt, st[0]])                                        # :768
        bottlenecks = allb[missing, idx_order]                                          # :776
    aux["vslt_out"] = enc_outputs[0]

    logits = classifier_head(sd, enc_outputs[0][:, 0, :], demo_embedding, cfg)
    return (logits, aux) if return_aux else logits


def classifier_head(sd, cls, demo_embedding, cfg: OracleConfig):
    """tri_mbt_vsltcls.py:248-255: nn.LayerNorm on the vslt CLS output, concat the demographic embedding, fc_list =
    Linear(512,256) -> BatchNorm1d (batch statistics in train mode, running statistics in eval mode) -> ReLU -> Linear."""
    c = F.layer_norm(cls, (256,), sd["layer_norms_after_concat.weight"], sd["layer_norms_after_concat.bias"], 1e-5)
    c = torch.cat([c, demo_embedding], dim=1)
    hdn = F.linear(c, sd["fc_list.0.weight"], sd["fc_list.0.bias"])
    if cfg.training:
        hdn = F.batch_norm(hdn, None, None, sd["fc_list.1.weight"], sd["fc_list.1.bias"], True, 0.1, 1e-5)
    else:
        hdn = F.batch_norm(hdn, sd["fc_list.1.running_mean"], sd["fc_list.1.running_var"], sd["fc_list.1.weight"],
                           sd["fc_list.1.bias"], False, 0.1, 1e-5)
    return F.linear(torch.relu(hdn), sd["fc_list.3.weight"], sd["fc_list.3.bias"])


def loss_fn(logits, y):
    """2_train.py:76 BCEWithLogitsLoss(mean) on output.squeeze() (trainer.py:128,176)."""
    return F.binary_cross_entropy_with_logits(logits.squeeze(-1), y.float())


def train_step_grads(sd, batch, cfg: OracleConfig):
    """loss + gradient of every parameter that receives one (the reference's 'live' set)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point
              and not k.endswith(("running_mean", "running_var")) and "positional_encoding" not in k}
    full = dict(sd)
    full.update(leaves)
    logits = forward(full, batch, cfg)
    loss = loss_fn(logits, batch["y"])
    loss.backward()
    grads = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    return logits.detach(), loss.detach(), grads
