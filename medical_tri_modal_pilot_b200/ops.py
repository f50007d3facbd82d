"""Tensor-level wrappers over the C ABI (include/tmp_b200.h). Every function launches hand-written sm_100a
kernels on the current CUDA stream; nothing here computes with PyTorch ops."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr, ptr_array, stream_ptr

D = 256
H = 4
ACT = torch.float16     # forward activations + 16-bit weight copies
GRD = torch.float16     # gradient tensors (carry runtime.GRAD_SCALE)
_FMT = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}   # 2 (fp32): gate / residual of the fp32 mode only
F32 = torch.float32


def _f32(t):
    """True when `t` is an fp32-stored activation / gradient tensor, i.e. the call belongs to the fp32 ("precise") mode."""
    return t is not None and t.dtype == torch.float32


def _fmt(t):
    if t is None:
        return 0
    try:
        return _FMT[t.dtype]
    except KeyError:
        raise RuntimeError(f"fp16 / bf16 (or fp32 in the fp32 mode) tensor required, got {t.dtype}") from None



def _cuda_contig(t, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: CUDA tensor required (no CPU fallback on this path)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: must be contiguous")
    return t


def set_reserved_sms(n: int) -> None:
    """Leave n SMs to communication kernels: every persistent / one-wave grid is sized from (SMs - n). Call before a CUDA
    graph of the step is captured (include/tmp_b200.h, tmp_set_reserved_sms)."""
    lib = _lib.load()
    if lib.tmp_set_reserved_sms(int(n)) != 0:
        raise RuntimeError(f"tmp_set_reserved_sms failed: {_lib.last_error()}")


def num_sms() -> int:
    return int(_lib.load().tmp_num_sms())


def lse_len(T: int) -> int:
    return (T + 127) // 128 * 128


def build_lengths(input_lengths, txt_lengths, img_time, n_img, multiimages, missing, skip_missing, T_v, T_i, T_t):
    """kv_len[3,B] int32. img_time: fp32 [B,n_img] or None."""
    B = input_lengths.numel()
    _cuda_contig(input_lengths, torch.int64, "input_lengths")
    _cuda_contig(txt_lengths, torch.int64, "txt_lengths")
    if img_time is not None:
        _cuda_contig(img_time, torch.float32, "img_time")
    if missing is not None:
        _cuda_contig(missing, torch.int64, "missing")
    out = torch.empty(3, B, dtype=torch.int32, device=input_lengths.device)
    check(_lib.load().tmp_build_lengths(ptr(input_lengths), ptr(txt_lengths), ptr(img_time), int(n_img),
                                        int(multiimages), ptr(missing), int(skip_missing), B, T_v, T_i, T_t, ptr(out),
                                        stream_ptr()), "tmp_build_lengths")
    return out


def materialize_mask(kv_len, T):
    B = kv_len.numel()
    _cuda_contig(kv_len, torch.int32, "kv_len")
    out = torch.empty(B, T, T, dtype=torch.uint8, device=kv_len.device)
    check(_lib.load().tmp_debug_materialize_mask(ptr(kv_len), B, T, ptr(out), stream_ptr()), "tmp_debug_materialize_mask")
    return out.bool()


def umse_embed(x, val4, tim4, Wfeat, out_dtype=torch.float32):
    """x [..., 3] fp32 -> E [..., 256]. val4/tim4: (Linear.weight, Linear.bias, LN.weight, LN.bias) fp32 tensors."""
    _cuda_contig(x, torch.float32, "x")
    n_tok = x.numel() // 3
    out = torch.empty(*x.shape[:-1], D, dtype=out_dtype, device=x.device)
    check(_lib.load().tmp_umse_embed_fwd(ptr(x), n_tok, ptr_array(val4), ptr_array(tim4), ptr(Wfeat), ptr(out),
                                         int(out_dtype == ACT), stream_ptr()), "tmp_umse_embed_fwd")
    return out


def stream_prologue_fwd(kind, B, n, x, val4, proj, times, n_slots, feat_id, tim4, Wfeat, cls, bottlenecks, ln_g, ln_b,
                        pe, drop_p, seed, salt, X0, seed_dev=None):
    fn = _lib.load().tmp_stream_prologue_fwd_f32 if _f32(X0) else _lib.load().tmp_stream_prologue_fwd
    check(fn(kind, B, n, ptr(x), ptr_array(val4) if val4 else None, ptr(proj),
                                              ptr(times), n_slots, feat_id, ptr_array(tim4), ptr(Wfeat), ptr(cls),
                                              ptr(bottlenecks), ptr(ln_g), ptr(ln_b), ptr(pe), float(drop_p), seed,
                                              salt, ptr(seed_dev), ptr(X0), stream_ptr()), "tmp_stream_prologue_fwd")



def stream_prologue_bwd(kind, B, n, x, val4, proj, times, n_slots, feat_id, tim4, Wfeat, cls, bottlenecks, ln_g, ln_b,
                        pe, drop_p, seed, salt, dX0, g_val, g_tim, g_feat, g_cls, g_bott, g_ln, dproj, seed_dev=None):
    fn = _lib.load().tmp_stream_prologue_bwd_f32 if _f32(dX0) else _lib.load().tmp_stream_prologue_bwd
    check(fn(kind, B, n, ptr(x), ptr_array(val4) if val4 else None, ptr(proj),
                                              ptr(times), n_slots, feat_id, ptr_array(tim4), ptr(Wfeat), ptr(cls),
                                              ptr(bottlenecks), ptr(ln_g), ptr(ln_b), ptr(pe), float(drop_p), seed,
                                              salt, ptr(seed_dev), ptr(dX0), ptr(g_val), ptr(g_tim), ptr(g_feat), ptr(g_cls),
                                              ptr(g_bott), ptr(g_ln), ptr(dproj), stream_ptr()),
          "tmp_stream_prologue_bwd")


def layernorm_fwd(x, gamma, beta, y, add=None, sum_out=None):
    rows = x.numel() // D
    fn = _lib.load().tmp_layernorm_fwd_f32 if _f32(x) else _lib.load().tmp_layernorm_fwd
    check(fn(ptr(x), ptr(add), ptr(gamma), ptr(beta), rows, ptr(sum_out), ptr(y), stream_ptr()), "tmp_layernorm_fwd")


def layernorm_bwd(dy, x, dres, gamma, dx, dgamma, dbeta, dx_drop=None, drop_p=0.0, seed=0, salt=0, seed_dev=None):
    rows = x.numel() // D
    fn = _lib.load().tmp_layernorm_bwd_f32 if _f32(x) else _lib.load().tmp_layernorm_bwd
    check(fn(ptr(dy), ptr(x), ptr(dres), ptr(gamma), rows, ptr(dx), ptr(dx_drop), float(drop_p), seed, salt,
             ptr(seed_dev), ptr(dgamma), ptr(dbeta), stream_ptr()), "tmp_layernorm_bwd")


def split_bf16x3(src, side_b=False, stack_rows=False):
    """fp32 [R,C] (last-dim contiguous, row stride a multiple of 4) -> bf16 [R,6C] (K-side blocks) or [6R,C] (row-stacked
    blocks): the bf16x3 operand split of the fp32 mode (csrc/precise.cu)."""
    C = src.shape[-1]
    R = src.numel() // C if src.is_contiguous() else src.shape[0]
    ld = C if src.is_contiguous() else src.stride(0)
    out = torch.empty((6 * R, C) if stack_rows else (R, 6 * C), dtype=torch.bfloat16, device=src.device)
    check(_lib.load().tmp_split_bf16x3(ptr(src), ld, R, C, int(side_b), int(stack_rows), ptr(out), stream_ptr()),
          "tmp_split_bf16x3")
    return out


def layernorm_bwd_attn(dy, x, dres, gamma, dx, dgamma, dbeta, attn_O, T, delta, dQKV):
    """layernorm_bwd whose dx is the attention's dO: also writes delta [B,4,T_lse] and zeroes dQKV[:, :256] (the fused
    protocol of attn_bwd, dQ_acc=None)."""
    rows = x.numel() // D
    check(_lib.load().tmp_layernorm_bwd_attn(ptr(dy), ptr(x), ptr(dres), ptr(gamma), rows, ptr(dx), ptr(dgamma),
                                             ptr(dbeta), ptr(attn_O), T, delta.shape[-1], ptr(delta), ptr(dQKV),
                                             stream_ptr()), "tmp_layernorm_bwd_attn")


def gemm(A, Bw, out=None, out_f32=None, bias=None, relu=False, gate=None, residual=None, alpha=1.0, drop_p=0.0, seed=0,
         salt=0, M=None, seed_dev=None, mask_out=None, row_live=None, rows_per_group=0):
    """out[M,N] = residual + dropout(gate>0 ? relu?(alpha*A@Bw^T + bias) : 0). A [M,K], Bw [N,K]: fp16 or bf16.
    mask_out (int32 [M, N/32]): receives the (result > 0) bit pattern; an int32 `gate` is read as such a bit mask.
    row_live (uint8 per group of rows_per_group rows): output tiles that only cover dead groups are skipped.
    fp32 mode (A and Bw fp32): both operands are split into bf16x3 along K and the same tcgen05 kernel accumulates the six
    partial products in fp32; `out` is then an fp32 tensor, gate / residual are fp32."""
    if _f32(A):
        if not _f32(Bw) or (out is not None and not _f32(out)):
            raise RuntimeError("gemm (fp32 mode): A, Bw and out must all be fp32")
        K = A.shape[-1]
        M = A.numel() // K if M is None else M
        A2 = A.reshape(-1, K)[:M] if A.is_contiguous() else A
        return gemm(split_bf16x3(A2, side_b=False), split_bf16x3(Bw, side_b=True), out_f32=out if out is not None else out_f32,
                    bias=bias, relu=relu, gate=gate, residual=residual, alpha=alpha, drop_p=drop_p, seed=seed, salt=salt,
                    M=M, seed_dev=seed_dev)
    K = A.shape[-1]
    M = A.numel() // K if M is None else M
    N = Bw.shape[0]
    ld_out = (out if out is not None else out_f32).shape[-1]
    gate_fmt = 3 if (gate is not None and gate.dtype == torch.int32) else _fmt(gate)
    check(_lib.load().tmp_gemm_bias_act_fwd(ptr(A), _fmt(A), A.stride(-2) if A.dim() > 1 else K, ptr(Bw), _fmt(Bw),
                                            Bw.stride(0), M, N, K, float(alpha), ptr(bias), int(relu), ptr(gate),
                                            gate_fmt, gate.shape[-1] if gate is not None else 0, ptr(residual),
                                            _fmt(residual), residual.shape[-1] if residual is not None else 0,
                                            float(drop_p), seed, salt, ptr(seed_dev), ptr(out), _fmt(out), ptr(out_f32),
                                            ld_out, ptr(mask_out), ptr(row_live), int(rows_per_group), stream_ptr()),
          "tmp_gemm_bias_act_fwd")


def gemm_wgrad(dY, X, dW, M=None, dbias=None):
    """dW[N,K] fp32 += dY[M,N]^T @ X[M,K]; dbias[N] fp32 += column sums of dY when given (fused bias gradient).
    fp32 mode (dY, X fp32): bf16x3 blocks stacked along the token dimension (6M rows); the bias gradient is a separate
    column-sum pass over the fp32 dY."""
    N, K = dY.shape[-1], X.shape[-1]
    M = dY.numel() // N if M is None else M
    if _f32(dY):
        if dbias is not None:
            colsum(dY, dbias, M=M)
        dY6 = split_bf16x3(dY.reshape(-1, N)[:M], side_b=False, stack_rows=True)
        X6 = split_bf16x3(X.reshape(-1, K)[:M], side_b=True, stack_rows=True)
        return gemm_wgrad(dY6, X6, dW, M=6 * M)
    check(_lib.load().tmp_gemm_wgrad(ptr(dY), _fmt(dY), N, ptr(X), _fmt(X), K, M, N, K, ptr(dW),
                                     ptr(dbias) if dbias is not None else None, stream_ptr()), "tmp_gemm_wgrad")


def colsum(dY, out, M=None):
    N = dY.shape[-1]
    M = dY.numel() // N if M is None else M
    fn = _lib.load().tmp_colsum_f32 if _f32(dY) else _lib.load().tmp_colsum
    check(fn(ptr(dY), N, M, N, ptr(out), stream_ptr()), "tmp_colsum")


def attn_fwd(qkv, kv_len, B, T, O, lse2, q_rows=None):
    if _f32(qkv):      # fp32 mode: CUDA-core fp32 attention (csrc/precise.cu)
        return check(_lib.load().tmp_attn_fwd_f32(ptr(qkv), ptr(kv_len), B, T, H, ptr(O), O.shape[-1], ptr(lse2),
                                                  lse2.shape[-1], stream_ptr()), "tmp_attn_fwd_f32")
    check(_lib.load().tmp_mma_attn_fwd(ptr(qkv), ptr(kv_len), B, T, H, ptr(O), O.shape[-1], ptr(lse2), lse2.shape[-1],
                                       T if q_rows is None else q_rows, stream_ptr()), "tmp_mma_attn_fwd")


def attn_bwd_single_query(qkv, dO_row, O_row, kv_len, B, T, q_row, lse2, dQKV):
    """Attention backward when only query row `q_row` of every sample carries a gradient (dO_row / O_row: [B,256] fp16).
    Writes all of dQKV [B*T,768]."""
    for t_, nm in ((qkv, "qkv"), (dO_row, "dO_row"), (O_row, "O_row"), (dQKV, "dQKV")):
        _cuda_contig(t_, torch.float16, nm)
    check(_lib.load().tmp_attn_bwd_single_query(ptr(qkv), ptr(dO_row), ptr(O_row), ptr(kv_len), B, T, H, int(q_row),
                                                ptr(lse2), lse2.shape[-1], ptr(dQKV), stream_ptr()),
          "tmp_attn_bwd_single_query")


def attn_bwd(qkv, O, dO, kv_len, B, T, lse2, delta, dQ_acc, dQKV, q_rows=None):
    if _f32(qkv):
        return check(_lib.load().tmp_attn_bwd_f32(ptr(qkv), ptr(O), ptr(dO), O.shape[-1], ptr(kv_len), B, T, H, ptr(lse2),
                                                  lse2.shape[-1], ptr(delta), ptr(dQKV), stream_ptr()), "tmp_attn_bwd_f32")
    check(_lib.load().tmp_mma_attn_bwd(ptr(qkv), ptr(O), ptr(dO), O.shape[-1], ptr(kv_len), B, T, H, ptr(lse2),
                                       lse2.shape[-1], ptr(delta), ptr(dQ_acc), ptr(dQKV),
                                       T if q_rows is None else q_rows, stream_ptr()),
          "tmp_mma_attn_bwd" if dQ_acc is None else "tmp_mma_attn_bwd(standalone)")


def bottleneck_mix_fwd(Yv, Yi, Yt, missing):
    B = Yv.shape[0]
    fn = _lib.load().tmp_bottleneck_mix_fwd_f32 if _f32(Yv) else _lib.load().tmp_bottleneck_mix_fwd
    check(fn(ptr(Yv), ptr(Yi), ptr(Yt), Yv.shape[1], Yi.shape[1], Yt.shape[1],
                                             ptr(missing), B, stream_ptr()), "tmp_bottleneck_mix_fwd")


def bottleneck_mix_bwd(dYv, dYi, dYt, upper_has_img_txt, missing, dropped=(None, None, None), drop_p=0.0, seed=0,
                       salts=(0, 0, 0), seed_dev=None):
    """dropped[m] (optional, same shape as dY_m): receives rows 0..3 of dY_m after dropout(drop_p, seed, salts[m]) -- see
    include/tmp_b200.h."""
    B = dYv.shape[0]
    fn = _lib.load().tmp_bottleneck_mix_bwd_f32 if _f32(dYv) else _lib.load().tmp_bottleneck_mix_bwd
    for d, y in zip(dropped, (dYv, dYi, dYt)):
        if d is not None and (d.shape != y.shape or d.dtype != y.dtype or not d.is_contiguous()):
            raise ValueError("bottleneck_mix_bwd: a dropped output must match its gradient tensor")
    check(fn(ptr(dYv), ptr(dYi), ptr(dYt), dYv.shape[1], dYi.shape[1], dYt.shape[1],
             int(upper_has_img_txt), ptr(missing), B, ptr(dropped[0]), ptr(dropped[1]), ptr(dropped[2]), float(drop_p),
             seed, ptr(seed_dev), int(salts[0]), int(salts[1]), int(salts[2]), stream_ptr()),
          "tmp_bottleneck_mix_bwd")


def dropout_apply(inp, out, drop_p, seed, salt, seed_dev=None):
    fn = _lib.load().tmp_dropout_apply_f32 if _f32(inp) else _lib.load().tmp_dropout_apply
    check(fn(ptr(inp), ptr(out), inp.numel(), float(drop_p), seed, salt, ptr(seed_dev), stream_ptr()), "tmp_dropout_apply")


def cast_weights(descs_dev, n_desc, max_R, max_C):
    check(_lib.load().tmp_cast_weights(ptr(descs_dev), n_desc, max_R, max_C, stream_ptr()), "tmp_cast_weights")


def adamw_step(w, g, m, v, lr, beta1, beta2, eps, weight_decay, step):
    """torch.optim.AdamW update of the flat fp32 buffers w (in place), with moments m, v (in place)."""
    for t, nm in ((w, "w"), (g, "g"), (m, "m"), (v, "v")):
        _cuda_contig(t, torch.float32, nm)
    check(_lib.load().tmp_adamw_step(ptr(w), ptr(g), ptr(m), ptr(v), w.numel(), float(lr), float(beta1), float(beta2),
                                     float(eps), float(weight_decay), int(step), stream_ptr()), "tmp_adamw_step")


def grad_nonfinite(g, state):
    """state[2] <- state[0] when g (fp32, numel % 4 == 0) holds an inf / NaN (see adamw_step_dev)."""
    _cuda_contig(g, torch.float32, "g")
    _cuda_contig(state, torch.int32, "state")
    if state.numel() < 3:
        raise ValueError("grad_nonfinite: state must hold 3 int32 words")
    check(_lib.load().tmp_grad_nonfinite(ptr(g), g.numel(), ptr(state), stream_ptr()), "tmp_grad_nonfinite")


def adamw_step_dev(w, g, m, v, lr_dev, beta1, beta2, eps, weight_decay, step_dev, count_skip=True):
    """Same update with lr (fp32 [1]) and the step words (int32 [3]: calls, skipped, last call with a non-finite
    gradient -- a flagged call is skipped and not counted) read from device memory: CUDA-graph replayable."""
    for t, nm in ((w, "w"), (g, "g"), (m, "m"), (v, "v"), (lr_dev, "lr_dev")):
        _cuda_contig(t, torch.float32, nm)
    _cuda_contig(step_dev, torch.int32, "step_dev")
    if step_dev.numel() < 3:
        raise ValueError("adamw_step_dev: step_dev must hold 3 int32 words (calls, skipped, last bad call)")
    check(_lib.load().tmp_adamw_step_dev(ptr(w), ptr(g), ptr(m), ptr(v), w.numel(), ptr(lr_dev), float(beta1),
                                         float(beta2), float(eps), float(weight_decay), ptr(step_dev), int(count_skip),
                                         stream_ptr()),
          "tmp_adamw_step_dev")


# ---- image-encoder feed (csrc/swin.cu) -------------------------------------------------------------------------
# `live`: optional uint8 [n_img] device tensor, 0 = skip the image (its features have no consumer)
def swin_patch_embed_ln(img, Wt, bconv, g, b, out, Cp, live=None):
    n_img = img.numel() // (224 * 224)
    _cuda_contig(img, torch.float32, "img")
    check(_lib.load().tmp_swin_patch_embed_ln(ptr(img), n_img, ptr(Wt), ptr(bconv), ptr(g), ptr(b), ptr(out), Cp,
                                              ptr(live), stream_ptr()), "tmp_swin_patch_embed_ln")


def swin_ln_window(x, g, b, n_img, H, C, Cp, shift, out, live=None, zero_dead=False):
    check(_lib.load().tmp_swin_ln_window(ptr(x), ptr(g), ptr(b), n_img, H, H, C, Cp, shift, ptr(out), ptr(live),
                                         int(zero_dead), stream_ptr()), "tmp_swin_ln_window")


def swin_window_attn(qkv, rel_bias, n_img, H, C, heads, shift, out, live=None):
    check(_lib.load().tmp_swin_window_attn(ptr(qkv), qkv.shape[-1], ptr(rel_bias), n_img, H, H, C, heads, shift, ptr(out),
                                           out.shape[-1], ptr(live), stream_ptr()), "tmp_swin_window_attn")


def swin_unwindow_add_ln(y, x, g, b, n_img, H, C, Cp, shift, hn, live=None):
    check(_lib.load().tmp_swin_unwindow_add_ln(ptr(y), ptr(x), ptr(g), ptr(b), n_img, H, H, C, Cp, shift, ptr(hn),
                                               ptr(live), stream_ptr()), "tmp_swin_unwindow_add_ln")


def swin_merge_ln(x, g, b, n_img, H, C, Cp, out, live=None):
    check(_lib.load().tmp_swin_merge_ln(ptr(x), ptr(g), ptr(b), n_img, H, H, C, Cp, ptr(out), ptr(live), stream_ptr()),
          "tmp_swin_merge_ln")


# ---- classifier head (training mode): csrc/head.cu ---------------------------------------------------------------------
HEAD_PARAM_ORDER = ("layer_norms_after_concat.weight", "layer_norms_after_concat.bias", "ie_demo.0.weight", "ie_demo.0.bias",
                    "ie_demo.1.weight", "ie_demo.1.bias", "fc_list.0.weight", "fc_list.0.bias", "fc_list.1.weight",
                    "fc_list.1.bias", "fc_list.3.weight", "fc_list.3.bias")


def head_scratch_floats(B: int) -> int:
    return max(32 * B, 64 * 7 * 256)


def head_fwd(cls, age, gen, params, run_mean, run_var, nbt, momentum, eps, saved, scratch, counter, logits):
    """params: 12 fp32 tensors in HEAD_PARAM_ORDER; saved: (Z [B,512], XC, XD [B,256], rstd_c, rstd_d [B], XH [B,256],
    invstd [256]) written here; logits [B]."""
    B = cls.shape[0]
    check(_lib.load().tmp_head_fwd(ptr(cls), ptr(age), ptr(gen), B, ptr_array(params), ptr(run_mean), ptr(run_var), ptr(nbt),
                                   float(momentum), float(eps), ptr_array(saved), ptr(scratch), ptr(counter), ptr(logits),
                                   stream_ptr()), "tmp_head_fwd")


def head_bwd(dlogit, age, gen, params, saved, grads, dcls, DH, scratch, counter):
    B = dlogit.shape[0]
    check(_lib.load().tmp_head_bwd(ptr(dlogit), ptr(age), ptr(gen), B, ptr_array(params), ptr_array(saved), ptr_array(grads),
                                   ptr(dcls), ptr(DH), ptr(scratch), ptr(counter), stream_ptr()), "tmp_head_bwd")
