"""`TRI_MBT_VSLTCLS` -- B200-native drop-in for the reference model class of the same name
(reference builder/models/8_missing_models/tri_mbt_vsltcls.py:17-263, TIE + --imgtxt-time 1 + swin + biobert).

Same constructor (`args` Namespace from control/config.py), same 18-positional-argument forward returning
`(logits[B,1], None, None)`, same state_dict keys and shapes, so `builder.models.get_model`,
`builder.trainer.missing_trainer`, `model.parameters()`, `.train()/.eval()` and checkpoints work unchanged.

What is different is everything underneath: the UMSE/TIE embedding, the encoder prologue, every LayerNorm, the
QKV / FFN / projection GEMMs (tcgen05), the modality-aware attention (tcgen05 flash kernel, forward and backward),
the bottleneck exchange and all weight/bias gradients run as hand-written sm_100a kernels from libtmp_b200.so
(see runtime.py). Only the frozen image encoder (stock torchvision Swin-T), the 2-feature demographic branch, the
classifier head and the optimizer stay PyTorch, as the task statement allows. There is no CPU fallback: calling
forward on CPU tensors raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.init as init

from .runtime import FusedPath

D = 256


class _XavierLinear(nn.Module):
    """state_dict shape of reference module.py:113-127 (`Linear` wrapper: xavier weight, zero bias)."""

    def __init__(self, i, o):
        super().__init__()
        self.linear = nn.Linear(i, o)
        init.xavier_uniform_(self.linear.weight)
        init.zeros_(self.linear.bias)


class _RefLayerNormParams(nn.Module):
    """Parameters of the reference's hand-written LayerNorm (module.py:130-136): `gamma`, `beta`."""

    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))
        self.beta = nn.Parameter(torch.zeros(dim))


class _MHAParams(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.query_proj = _XavierLinear(dim, dim)
        self.key_proj = _XavierLinear(dim, dim)
        self.value_proj = _XavierLinear(dim, dim)


class _FFNParams(nn.Module):
    def __init__(self, d_in, d_hid):
        super().__init__()
        self.w_1 = nn.Conv1d(d_in, d_hid, 1)   # reference module.py:60-61 (weights [out,in,1])
        self.w_2 = nn.Conv1d(d_hid, d_in, 1)


class _EncoderLayerParams(nn.Module):
    """Parameter container with the names of reference encoder.py:8-21 (compute is in the fused kernels)."""

    def __init__(self, d_model, d_ff):
        super().__init__()
        self.attention_prenorm = _RefLayerNormParams(d_model)
        self.feed_forward_prenorm = _RefLayerNormParams(d_model)
        self.self_attention = _MHAParams(d_model)
        self.feed_forward = _FFNParams(d_model, d_ff)


class _PositionalEncoding(nn.Module):
    def __init__(self, d_model, max_len):
        super().__init__()
        import math
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))


class _FusionTransformerParams(nn.Module):
    """Names of TrimodalTransformerEncoder_MBT.__init__ (reference mbt_encoder.py:643-694)."""

    def __init__(self, n_modality, bottlenecks_n, n_layers, d_model, d_ff, pe_maxlen):
        super().__init__()
        self.layer_norms_after_concat = nn.LayerNorm(d_model)   # present (unused) in the reference as well
        self.cls_token_per_modality = nn.ParameterList(
            [nn.Parameter(torch.randn(1, 1, d_model)) for _ in range(n_modality)])
        self.bottlenecks = nn.Parameter(torch.randn(1, bottlenecks_n, d_model))
        self.layer_norms_in = nn.ModuleList([nn.LayerNorm(d_model) for _ in range(n_modality)])
        self.positional_encoding = _PositionalEncoding(d_model, pe_maxlen)
        self.layer_stacks = nn.ModuleList(
            nn.ModuleList([_EncoderLayerParams(d_model, d_ff) for _ in range(n_modality)]) for _ in range(n_layers))


def build_swin_t_m():
    """Stock torchvision Swin-T patched like the reference's copy (swin_transformer.py:611-618, 646): 1-channel
    patch conv, forward returns the normalised [N,7,7,768] feature map (no pooling / head)."""
    from torchvision.models.swin_transformer import SwinTransformer

    class SwinFeatures(SwinTransformer):
        def forward(self, x):
            return self.norm(self.features(x))

    m = SwinFeatures(patch_size=[4, 4], embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
                     window_size=[7, 7], stochastic_depth_prob=0.2)
    m.features[0][0] = nn.Conv2d(1, 96, kernel_size=(4, 4), stride=(4, 4))
    return m


class TRI_MBT_VSLTCLS(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        if getattr(args, "vslt_type", "TIE") != "TIE":
            raise NotImplementedError("B200 path implements --vslt-type TIE (the north-star configuration)")
        if getattr(args, "berttype", "biobert") != "biobert":
            raise NotImplementedError("B200 path implements --berttype biobert (768-d token embeddings)")
        if getattr(args, "img_model_type", "swin") != "swin":
            raise NotImplementedError("B200 path implements --img-model-type swin")
        if int(getattr(args, "mbt_fusion_startIdx", 0)) != 0:
            raise NotImplementedError("--mbt-fusion-startIdx > 0 (non-fused prefix layers) is out of scope")
        if int(getattr(args, "residual_bottlenecks", 0)) != 0:
            raise NotImplementedError("--residual-bottlenecks 1 is out of scope")
        self.num_layers = int(args.transformer_num_layers)
        self.num_heads = int(getattr(args, "transformer_num_head", 4))
        self.model_dim = int(getattr(args, "transformer_dim", D))
        if self.model_dim != D or self.num_heads != 4:
            raise NotImplementedError("kernels are specialised for --transformer-dim 256 / --transformer-num-head 4")
        self.dropout = float(getattr(args, "dropout", 0.1))
        self.multiimages = int(getattr(args, "multiimages", 0))
        self.vsltonly = int(getattr(args, "mbt_only_vslt", 1))
        # The reference only runs with vslt_img_txt; the 1- and 2-modal input types are served
        # by the exactly equivalent constant `missing` code.
        self.input_types = getattr(args, "input_types", "vslt_img_txt")
        if self.input_types not in ("vslt", "vslt_txt", "vslt_img", "vslt_img_txt"):
            raise ValueError(f"unknown --input-types {self.input_types}")
        self.bottlenecks_n = 4

        mk_ie = lambda k: nn.Sequential(nn.Linear(k, D), nn.LayerNorm(D), nn.ReLU(inplace=True))
        self.activations = nn.ModuleDict([["lrelu", nn.LeakyReLU()], ["prelu", nn.PReLU()], ["relu", nn.ReLU(inplace=True)],
                                          ["tanh", nn.Tanh()], ["sigmoid", nn.Sigmoid()],
                                          ["leaky_relu", nn.LeakyReLU(0.2)], ["elu", nn.ELU()]])
        self.ie_vslt = mk_ie(1)
        self.ie_time = mk_ie(1)
        self.ie_feat = nn.Embedding(20, D)
        self.ie_demo = mk_ie(2)
        self.txt_embedding = nn.Linear(768, D)
        self.img_encoder = build_swin_t_m()
        self.img_encoder.eval()
        self.linear = nn.Linear(768, D)
        self.flatten = nn.Flatten(1, 2)
        self.fusion_transformer = _FusionTransformerParams(3, self.bottlenecks_n, self.num_layers, D, 4 * D, 2500)
        self.rmse_layer = nn.Linear(2 * D, 1)
        self.layer_norms_after_concat = nn.LayerNorm(D)
        self.fc_list = nn.Sequential(nn.Linear(2 * D, D), nn.BatchNorm1d(D), self.activations["relu"], nn.Linear(D, 1))
        self._fused = FusedPath(self)
        self.native_swin = True      # frozen image encoder through swin_feed.SwinFeed (False: stock torchvision forward)
        self.swin_skip_dead = True   # skip the encoder work of images that no key of the img stream can see
        self.img_autocast = True     # run the frozen Swin in bf16 (its output feeds an fp16 tensor-core GEMM anyway)
        # precision mode: "fp16" (tensor-core plan, the reference's autocast dtype) or "fp32" (north-star FP32 parity mode:
        # fp32 storage, bf16x3-split tensor-core GEMMs, fp32 attention; runtime.FusedPath.precision). Not a reference flag:
        # `args.precision` if the caller sets it, else env TMP_B200_PRECISION.
        import os
        self.set_precision(getattr(args, "precision", None) or os.environ.get("TMP_B200_PRECISION", "fp16"))

    def set_precision(self, precision: str):
        if precision not in ("fp16", "fp32"):
            raise ValueError(f"precision must be 'fp16' or 'fp32', got {precision!r}")
        self._fused.precision = precision
        if precision == "fp32":      # the frozen image encoder runs as the stock fp32 torchvision module in this mode
            self.native_swin = False
            self.img_autocast = False
        return self

    # -- reference forward contract (tri_mbt_vsltcls.py:167) ----------------------------------------------------
    def forward(self, x, h, m, d, x_m, age, gen, input_lengths, txts, txt_lengths, img, missing, f_indices, img_time,
                txt_time, flow_type, reports_tokens, reports_lengths):
        if not x.is_cuda:
            raise RuntimeError("TRI_MBT_VSLTCLS (B200) needs CUDA tensors: the fused path has no CPU fallback")
        B = x.shape[0]
        missing = self.tri_missing_code(missing, B, x.device)
        # the frozen image encoder runs INSIDE the fused path, on the img modality's CUDA stream (FusedPath.forward calls
        # encode_images there), so the vslt / txt streams of layer 0 overlap it
        cls_out = self._fused(x, input_lengths, txts, txt_lengths, img, img_time, txt_time, missing)
        from . import head
        if head.usable(self, cls_out):
            # training mode: LayerNorm + demographic branch + Linear / BatchNorm1d / ReLU / Linear in three launches (head.py)
            return head.fused_head(self, cls_out, age, gen), None, None
        # the demographic branch (reference tri_mbt_vsltcls.py:176-177, in front of the encoder there) is evaluated AFTER the fused path: its
        # autograd nodes are then younger than the fused function's, so the engine runs their backward first and every head
        # gradient exists when dL/dCLS reaches the fused backward (trainer.GradSync averages them during that backward)
        demographic = torch.cat([age.unsqueeze(1), gen.unsqueeze(1)], dim=1).float()
        demo_embedding = self.ie_demo(demographic)
        classInput = self.layer_norms_after_concat(cls_out)
        classInput = torch.cat([classInput, demo_embedding], dim=1)
        if "rmse" in getattr(self.args, "auxiliary_loss_type", "none"):
            output2 = self.rmse_layer(classInput).squeeze()
        else:
            output2 = None
        output1 = self.fc_list(classInput)
        return output1, output2, None

    def tri_missing_code(self, missing, B, device):
        """Map the trainer's `missing_num` of a 1-/2-modal run onto the tri-modal code (0 all present, 1 txt missing,
        2 img missing, 3 both; reference trainer.py:68-84 and its remap :99-105, inverted -- SURVEY.md 8c):
        vslt -> 3; vslt_txt {0,1} -> {2,3}; vslt_img {0,1} -> {1,3}."""
        if self.input_types == "vslt_img_txt":
            return missing.to(torch.long)
        if self.input_types == "vslt" or missing is None:
            code = {"vslt": 3, "vslt_txt": 2, "vslt_img": 1}[self.input_types]
            return torch.full((B,), code, dtype=torch.long, device=device)
        m = (missing.to(torch.long) != 0).to(torch.long)
        return 2 + m if self.input_types == "vslt_txt" else 1 + 2 * m

    def live_images(self, img_time, missing):
        """uint8 [B*n_img]: which images have a consumer. With --multiimages 1 the img stream's key length is 49 * #(img_time
        != 10) (tri_mbt_vsltcls.py:229-232), i.e. the FIRST count_b slots of a sample are visible, whatever their position
        in time; the rest are masked keys. Samples whose `missing` code de-selects the img stream (2, 3) have no live image.
        With --multiimages 0 the stream is unmasked (:144,:234): only the missing code decides."""
        B = missing.shape[0]
        present = ~((missing == 2) | (missing == 3))
        if self.multiimages == 1:
            cnt = (img_time.reshape(B, -1) != 10).sum(1, keepdim=True)
            live = (torch.arange(img_time.numel() // B, device=missing.device)[None, :] < cnt) & present[:, None]
        else:
            live = present[:, None]
        return live.reshape(-1).to(torch.uint8).contiguous()

    def encode_images(self, img, missing=None, ready=None, img_time=None):
        """Frozen image encoder (reference tri_mbt_vsltcls.py:205-209: reshape(-1,1,224,224), torch.no_grad).
        Returns [B*n_img, 49, 768] fp16 (the A operand of the 768->256 projection GEMM). `ready`: optional per-chunk CUDA
        events of a staged upload (swin_feed.SwinFeed.__call__). With `missing` and `img_time` given (the fused path does),
        images without a consumer are skipped and come out as zero rows (`live_images`; exact: SURVEY Appendix A)."""
        f32 = self._fused.precision == "fp32"

        def wait_all():
            for ev in (ready or ()):
                torch.cuda.current_stream().wait_event(ev)

        if img.dim() == 3 and img.shape[-1] == 768:      # test hook: pre-computed Swin features
            wait_all()
            return img.to(torch.float32 if f32 else torch.float16).contiguous()
        if self.multiimages == 1:
            img = img.reshape(-1, 1, 224, 224)
        with torch.no_grad():
            if self.native_swin:
                live = None
                if self.swin_skip_dead and missing is not None and img_time is not None:
                    live = self.live_images(img_time, missing)
                return self._swin_feed()(img, ready=ready, live=live)     # sm_100a kernels (swin_feed.py), fp16 [N,49,768]
            wait_all()
            if self.img_autocast:
                f = self._img_encoder_bf16()(img.to(torch.bfloat16))
            else:
                f = self.img_encoder(img)
        return f.reshape(f.shape[0], 49, 768).to(torch.float32 if f32 else torch.float16).contiguous()

    def _swin_sig(self):
        first = self.img_encoder.features[0][0].weight
        last = self.img_encoder.norm.weight
        return (first.data_ptr(), first._version, last.data_ptr(), last._version, first.device)

    def _swin_feed(self):
        """B200-native forward of the frozen image encoder; fp16 weight copies rebuilt when the master weights change."""
        sig = self._swin_sig()
        cached = self.__dict__.get("_swin_native")
        if cached is None or cached[0] != sig:
            from .swin_feed import SwinFeed
            cached = (sig, SwinFeed(self.img_encoder))
            self.__dict__["_swin_native"] = cached
        return cached[1]

    def _img_encoder_bf16(self):
        """bf16 shadow of the frozen image encoder (the fp32 module stays the state_dict master). Pure-bf16 weights
        avoid autocast's per-op casts and fp32 LayerNorm traffic; rebuilt when the master weights change."""
        first = self.img_encoder.features[0][0].weight
        last = self.img_encoder.norm.weight
        sig = (first.data_ptr(), first._version, last.data_ptr(), last._version, first.device)
        cached = self.__dict__.get("_img_lp")
        if cached is None or cached[0] != sig:
            import copy
            lp = copy.deepcopy(self.img_encoder).to(device=first.device, dtype=torch.bfloat16).eval()
            for p in lp.parameters():
                p.requires_grad_(False)
            cached = (sig, lp)
            self.__dict__["_img_lp"] = cached
        return cached[1]
