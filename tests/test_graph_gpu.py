"""CUDA-graph replay of the training step (trainer.GraphedStep) == the eager step, and the device-resident per-step
scalars (dropout step counter, AdamW step / learning rate) behave like their host-scalar forms."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _trainer_inputs(batch):
    miss = batch["missing"]
    miss3 = torch.stack([torch.zeros_like(miss), (miss >= 2).long(), (miss % 2).long()], 1).float()
    static = torch.stack([batch["gen"], batch["age"]], 1)
    return static, miss3


def _run_steps(cuda_graph, n_steps, dropout=0.0, lrs=None, probe_at=None):
    from golden_util import fixture_inputs, fixture_names, load_fixture
    from builder.trainer import get_trainer
    from medical_tri_modal_pilot_b200.optim import FlatAdamW
    from test_model_parity_gpu import build_model
    fx = load_fixture(fixture_names()[0])
    sd, batch, cfg = fixture_inputs(fx)
    B = batch["x"].shape[0]
    torch.manual_seed(0)
    model = build_model(cfg, sd, B, dropout=dropout).train()
    model.args.cuda_graph = cuda_graph
    # eps = 1e-3: with the default 1e-8 AdamW turns the ~1e-7 run-to-run jitter of the fp32-atomic weight gradients
    # into full-size +-lr updates of the near-zero-gradient parameters, and two EAGER runs already drift apart
    opt = FlatAdamW(model, lr=1e-3, weight_decay=1e-2, eps=1e-3)
    crit = torch.nn.BCEWithLogitsLoss()
    static, miss3 = _trainer_inputs(batch)
    losses = []
    for it in range(n_steps):
        if lrs is not None:
            opt.param_groups[0]["lr"] = lrs[it]
        _, loss = get_trainer(model.args, it, batch["x"], static, batch["input_lengths"], batch["y"], None, model, None,
                              torch.device("cuda"), None, opt, crit, x_txt=batch["txts"], x_img=batch["img_feats"],
                              txt_lengths=batch["txt_lengths"], imgtxt_time=(batch["img_time"], batch["txt_time"]),
                              missing=miss3, flow_type="train")
        losses.append(loss)
        if probe_at is not None and it == probe_at:
            # an evaluation batch of ANOTHER shape between two train steps (validation inside the epoch loop): the fused
            # path switches workspaces; the captured train graph must keep replaying into its own
            nb, nl = B // 2, batch["x"].shape[1] - 7
            model.eval()
            with torch.no_grad():
                get_trainer(model.args, it, batch["x"][:nb, :nl].contiguous(), static[:nb], batch["input_lengths"][:nb].clamp(max=nl),
                            batch["y"][:nb], None, model, _Log(), torch.device("cuda"), None, opt, crit,
                            x_txt=batch["txts"][:nb], x_img=batch["img_feats"][:nb * (3 if cfg.multiimages else 1)],
                            txt_lengths=batch["txt_lengths"][:nb], imgtxt_time=(batch["img_time"][:nb], batch["txt_time"][:nb]),
                            missing=miss3[:nb], flow_type="test")
            model.train()
    params = {k: p.detach().clone() for k, p in model.named_parameters() if not k.startswith("img_encoder.")}
    return losses, params, model


class _Log:
    class evaluator:
        @staticmethod
        def add_batch(y, p):
            pass


def test_graph_survives_a_workspace_switch():
    """a captured graph bakes raw workspace pointers; an eval batch of a different shape used to free
    and reallocate those workspaces. Same training run with and without an eval probe of another shape after step 3
    (after the capture): identical losses within the eager run-to-run jitter."""
    l_a, p_a, _ = _run_steps(True, 6)
    l_b, p_b, model = _run_steps(True, 6, probe_at=3)
    assert len(model._fused._ws_cache) == 2
    for a, b in zip(l_a, l_b):
        assert abs(a - b) < max(1e-3, 0.03 * a), (l_a, l_b)
    # and the same probe in eager mode
    l_c, _, _ = _run_steps(False, 6, probe_at=3)
    for a, c in zip(l_a, l_c):
        assert abs(a - c) < max(1e-3, 0.03 * a), (l_a, l_c)


def test_backward_after_an_interleaved_forward_raises():
    from golden_util import fixture_inputs, fixture_names, load_fixture
    from test_model_parity_gpu import build_model, run_model
    fx = load_fixture(fixture_names()[0])
    sd, batch, cfg = fixture_inputs(fx)
    model = build_model(cfg, sd, batch["x"].shape[0]).train()
    out, b = run_model(model, batch)
    with torch.no_grad():
        run_model(model, batch)                      # e.g. an eval probe before backward
    with pytest.raises(RuntimeError, match="another forward"):
        out.sum().backward()


def test_gradient_accumulation_adds():
    """Two backward calls without zero_grad accumulate (used to overwrite silently)."""
    from golden_util import fixture_inputs, fixture_names, load_fixture
    from test_model_parity_gpu import build_model, run_model
    fx = load_fixture(fixture_names()[0])
    sd, batch, cfg = fixture_inputs(fx)
    model = build_model(cfg, sd, batch["x"].shape[0]).train()
    out, b = run_model(model, batch)
    out.sum().backward()
    g1 = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    out, b = run_model(model, batch)
    out.sum().backward()
    for k, p in model.named_parameters():
        if p.grad is None or k.startswith("img_encoder."):
            continue
        ref = 2 * g1[k]
        assert (p.grad - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-6, k


def test_flat_adamw_state_dict_round_trip():
    """reference logger.py:167 writes optimizer.state_dict() into every checkpoint; resuming must restore the
    flat moments and the step count."""
    import io
    from medical_tri_modal_pilot_b200.optim import FlatAdamW
    l1, p1, model = _run_steps(False, 3)
    # continue two more steps from a checkpoint in a fresh model/optimizer vs in place
    from golden_util import fixture_inputs, fixture_names, load_fixture
    from builder.trainer import get_trainer
    from test_model_parity_gpu import build_model
    fx = load_fixture(fixture_names()[0])
    sd, batch, cfg = fixture_inputs(fx)
    B = batch["x"].shape[0]
    static, miss3 = _trainer_inputs(batch)
    crit = torch.nn.BCEWithLogitsLoss()

    def steps(model, opt, n):
        out = []
        for it in range(n):
            _, loss = get_trainer(model.args, it, batch["x"], static, batch["input_lengths"], batch["y"], None, model, None,
                                  torch.device("cuda"), None, opt, crit, x_txt=batch["txts"], x_img=batch["img_feats"],
                                  txt_lengths=batch["txt_lengths"], imgtxt_time=(batch["img_time"], batch["txt_time"]),
                                  missing=miss3, flow_type="train")
            out.append(loss)
        return out

    torch.manual_seed(0)
    m_a = build_model(cfg, sd, B).train(); m_a.args.cuda_graph = False
    o_a = FlatAdamW(m_a, lr=1e-3, weight_decay=1e-2, eps=1e-3)
    steps(m_a, o_a, 3)
    buf = io.BytesIO()
    torch.save({"model": m_a.state_dict(), "optimizer": o_a.state_dict()}, buf)     # logger.py:166-177 layout
    buf.seek(0)
    ck = torch.load(buf, map_location="cuda", weights_only=False)
    m_b = build_model(cfg, sd, B).train(); m_b.args.cuda_graph = False
    m_b.load_state_dict(ck["model"])
    o_b = FlatAdamW(m_b, lr=1e-3, weight_decay=1e-2, eps=1e-3)
    o_b.load_state_dict(ck["optimizer"])
    assert o_b.t == 3 and o_b.steps_taken() == (3, 0) and o_b.m.abs().sum().item() > 0
    la, lb = steps(m_a, o_a, 2), steps(m_b, o_b, 2)
    for a, b in zip(la, lb):
        assert abs(a - b) < max(1e-3, 0.03 * a), (la, lb)


def test_graph_replay_matches_eager_steps():
    """6 optimisation steps (2 eager warm-up calls, capture, 4 replays) with a changing learning rate give the same
    loss curve and parameters as 6 eager steps. fp32 atomics in the weight-gradient kernels give ~1e-6 jitter per step
    which BatchNorm over the 24-sample fixture batch amplifies step by step (two EAGER runs drift apart by ~6e-4 in the
    loss after 6 steps, measured; the last step follows the lr = 2e-3 update and has shown 2.5e-3 between a replayed
    and an eager run whose first five losses agreed to 5e-5). The bar per step is 5x the drift between two eager runs,
    floor 1e-3, or 3 % of that step's loss -- a stale captured learning rate or step count moves the loss by > 15 %."""
    lrs = [1e-3, 1e-3, 5e-4, 5e-4, 2e-3, 1e-3]
    l_e, p_e, _ = _run_steps(False, 6, lrs=lrs)
    l_e2, _, _ = _run_steps(False, 6, lrs=lrs)
    l_g, p_g, model = _run_steps(True, 6, lrs=lrs)
    tol = max(1e-3, 5 * max(abs(a - b) for a, b in zip(l_e, l_e2)))
    gs = next(iter(model.__dict__["_graphed_steps"].values()))
    assert gs.graph is not None and gs.launches_per_replay > 50
    for a, b in zip(l_e, l_g):
        assert abs(a - b) < max(tol, 0.03 * a), (l_e, l_e2, l_g)
    assert l_e[-1] < l_e[0]                 # it trains
    for k in p_e:
        d = (p_e[k] - p_g[k]).abs().max().item()
        assert d <= 2e-4 + 1e-3 * p_e[k].abs().max().item() + 10 * max(tol, 0.03 * l_e[-1]), (k, d)


def test_graph_replays_draw_fresh_dropout_masks():
    """With dropout on, two replays of the same graph on the same batch must not reuse the mask: the device step counter
    feeds the hash. Checked on the X0 activations (prologue dropout) saved by consecutive replays."""
    from medical_tri_modal_pilot_b200 import trainer
    losses, _, model = _run_steps(True, 4, dropout=0.3)
    fp = model._fused
    x_a = fp.ws[0]["X"][0].clone()
    gs = next(iter(model.__dict__["_graphed_steps"].values()))
    assert gs.graph is not None
    c0 = int(fp.step_dev.item())
    gs.step()
    torch.cuda.synchronize()
    assert int(fp.step_dev.item()) == c0 + 1
    x_b = fp.ws[0]["X"][0]
    za, zb = (x_a[:, 5:] == 0), (x_b[:, 5:] == 0)
    assert 0.2 < za.float().mean().item() < 0.4 and 0.2 < zb.float().mean().item() < 0.4
    assert (za != zb).float().mean().item() > 0.2       # independent masks differ on 2 p (1-p) = 42 % of the elements


def test_seed_dev_equals_scalar_seed():
    from medical_tri_modal_pilot_b200 import ops
    g = torch.randn(64, 256, device="cuda").half()
    o1, o2 = torch.empty_like(g), torch.empty_like(g)
    ops.dropout_apply(g, o1, 0.25, 1234 + 7, 5)
    ops.dropout_apply(g, o2, 0.25, 1234, 5, seed_dev=torch.tensor([7], dtype=torch.int32, device="cuda"))
    assert torch.equal(o1, o2)
    A = torch.randn(256, 256, device="cuda").half()
    W = torch.randn(256, 256, device="cuda").half() / 16
    y1, y2 = torch.empty(256, 256, device="cuda", dtype=torch.float16), torch.empty(256, 256, device="cuda", dtype=torch.float16)
    ops.gemm(A, W, out=y1, drop_p=0.25, seed=99 + 3, salt=11)
    ops.gemm(A, W, out=y2, drop_p=0.25, seed=99, salt=11, seed_dev=torch.tensor([3], dtype=torch.int32, device="cuda"))
    assert torch.equal(y1, y2) and 0.15 < (y1 == 0).float().mean().item() < 0.35


def test_adamw_dev_equals_host_scalars():
    from medical_tri_modal_pilot_b200 import ops
    torch.manual_seed(1)
    n = 1 << 18
    w1 = torch.randn(n, device="cuda"); w2 = w1.clone()
    m1, v1, m2, v2 = (torch.zeros(n, device="cuda") for _ in range(4))
    lr_dev = torch.zeros(1, device="cuda")
    t_dev = torch.zeros(4, dtype=torch.int32, device="cuda")
    for t, lr in ((1, 3e-3), (2, 1e-3), (3, 2e-3)):
        g = torch.randn(n, device="cuda")
        ops.adamw_step(w1, g, m1, v1, lr, 0.9, 0.999, 1e-8, 1e-2, t)
        lr_dev.fill_(lr); t_dev.zero_(); t_dev[:1].fill_(t)
        ops.adamw_step_dev(w2, g, m2, v2, lr_dev, 0.9, 0.999, 1e-8, 1e-2, t_dev)
    assert (w1 - w2).abs().max().item() < 1e-6


def test_nonfinite_gradient_skips_the_optimizer_step():
    """Overflow guard of the 16-bit plan (gradients pass through fp16 scratch with a static scale): a call whose flat
    gradient holds an inf / NaN leaves w, m, v untouched and is not counted; the next clean call continues with the bias
    corrections of the steps actually taken. All on the device (graph-replayable): {calls, skipped, last bad call}."""
    from medical_tri_modal_pilot_b200 import ops
    torch.manual_seed(2)
    n = 1 << 16
    w = torch.randn(n, device="cuda"); w_ref = w.clone()
    m, v, m_ref, v_ref = (torch.zeros(n, device="cuda") for _ in range(4))
    lr_dev = torch.full((1,), 1e-3, device="cuda")
    state = torch.zeros(4, dtype=torch.int32, device="cuda")
    taken = 0
    for call, poison in enumerate((None, float("inf"), None, float("nan"), float("-inf"), None), start=1):
        g = torch.randn(n, device="cuda")
        if poison is not None:
            g[12345] = poison
        state[:1].add_(1)
        ops.grad_nonfinite(g, state)
        before = (w.clone(), m.clone(), v.clone())
        ops.adamw_step_dev(w, g, m, v, lr_dev, 0.9, 0.999, 1e-8, 1e-2, state)
        if poison is None:
            taken += 1
            ops.adamw_step(w_ref, g, m_ref, v_ref, 1e-3, 0.9, 0.999, 1e-8, 1e-2, taken)
        else:
            assert torch.equal(w, before[0]) and torch.equal(m, before[1]) and torch.equal(v, before[2])
        assert state[:2].tolist() == [call, call - taken]
    assert taken == 3 and (w - w_ref).abs().max().item() < 1e-6 and torch.isfinite(w).all()


def test_flat_adamw_survives_an_overflowing_backward():
    """End to end: a step whose fused-path gradient overflowed (injected inf in flat_g) changes no fused-path weight and
    FlatAdamW reports it; training continues on the next batch."""
    from builder.models import get_model
    from medical_tri_modal_pilot_b200 import synth, trainer
    from medical_tri_modal_pilot_b200.config import make_args
    from medical_tri_modal_pilot_b200.optim import FlatAdamW
    dev = torch.device("cuda", 0)
    args = make_args(transformer_num_layers=2, multiimages=1, mbt_only_vslt=1, input_types="vslt_img_txt", imgtxt_time=1,
                     dropout=0.1, batch_size=8, img_pretrain="No", TIE_len=40)
    args.device = dev
    torch.manual_seed(0)
    model = get_model(args)(args).to(dev).train()
    opt = FlatAdamW(model, lr=1e-3, weight_decay=1e-6)
    crit = torch.nn.BCEWithLogitsLoss()
    host = synth.make_batch(8, 40, n_img=3, seed=5, full_length=False, missing_mode="mixed", with_pixels=True, feats=False)
    miss = host["missing"]
    missing3 = torch.stack([torch.zeros_like(miss), (miss >= 2).long(), (miss % 2).long()], 1).float()
    r = {k: v.to(dev) for k, v in host.items()}
    b = trainer.prepare_batch(args, dev, r["x"], torch.stack([r["gen"], r["age"]], 1), r["input_lengths"], r["y"], r["img"],
                              r["txts"], r["txt_lengths"], (r["img_time"], r["txt_time"]), missing3.to(dev))
    fp = model._fused
    trainer.train_step(args, model, opt, crit, b, None, 0, None)
    assert opt.steps_taken() == (1, 0)
    w0 = fp.flat_w.clone()
    # second step by hand: forward / backward, then poison one gradient element before the optimizer runs
    opt.zero_grad()
    _, loss = trainer.forward_loss(args, model, crit, b, "train")
    loss.backward()
    fp.flat_g[7] = float("inf")
    opt.step()
    assert opt.steps_taken() == (1, 1) and torch.equal(fp.flat_w, w0)
    trainer.train_step(args, model, opt, crit, b, None, 2, None)
    assert opt.steps_taken() == (2, 1) and not torch.equal(fp.flat_w, w0) and torch.isfinite(fp.flat_w).all()


def test_staged_upload_feeds_every_replay_with_the_new_batch():
    """trainer.GraphedStep uploads each batch on a copy stream in stages (pixel chunk 0, small tensors, pixel chunks 1..,
    text) and the captured graph waits on EXTERNAL events exactly where it first touches each group. The test poisons the
    static input buffers with NaN before every upload (in the compute stream, which the copy stream waits for): a kernel
    that read a chunk before its upload had landed would read NaN and the loss would be NaN. Two different pinned-host
    batches alternate through the user call for 8 steps (2 eager warm-ups, capture, 5 replays); the losses must also equal
    those of the plain eager step (copies issued in the compute stream)."""
    import math
    from builder.models import get_model
    from builder.trainer import get_trainer
    from medical_tri_modal_pilot_b200 import synth, trainer
    from medical_tri_modal_pilot_b200.config import make_args
    from medical_tri_modal_pilot_b200.optim import FlatAdamW
    B, L, NL = 32, 200, 2

    def run(cuda_graph):
        torch.manual_seed(0)
        args = make_args(transformer_num_layers=NL, multiimages=1, mbt_only_vslt=1, input_types="vslt_img_txt", imgtxt_time=1,
                         dropout=0.0, batch_size=B, img_pretrain="No")
        args.device = torch.device("cuda")
        args.cuda_graph = cuda_graph
        model = get_model(args)(args).to(args.device).train()
        opt = FlatAdamW(model, lr=1e-4, weight_decay=1e-6, eps=1e-3)
        crit = torch.nn.BCEWithLogitsLoss()
        batches = []
        for seed in (5, 6):
            hb = synth.make_batch(B, L, n_img=3, seed=seed, missing_mode="none", with_pixels=True, feats=False)
            if seed == 6:
                hb["y"] = 1.0 - hb["y"]               # the two batches are distinguishable by their loss
            miss = hb["missing"]
            hb["missing3"] = torch.stack([torch.zeros_like(miss), (miss >= 2).long(), (miss % 2).long()], 1).float()
            hb["static"] = torch.stack([hb["gen"], hb["age"]], 1)
            batches.append({k: v.pin_memory() for k, v in hb.items()})
        orig_load = trainer.GraphedStep.load

        def poisoned_load(self, raw):
            for k in ("x_img", "x_txt", "train_x"):
                self.static[k].fill_(float("nan"))
            return orig_load(self, raw)
        trainer.GraphedStep.load = poisoned_load
        try:
            losses = []
            for it in range(8):
                src = batches[it % 2]
                _, loss = get_trainer(args, it, src["x"], src["static"], src["input_lengths"], src["y"], None, model, None,
                                      args.device, None, opt, crit, x_txt=src["txts"], x_img=src["img"],
                                      txt_lengths=src["txt_lengths"], imgtxt_time=(src["img_time"], src["txt_time"]),
                                      missing=src["missing3"], flow_type="train")
                losses.append(loss)
        finally:
            trainer.GraphedStep.load = orig_load
        return losses, model

    l_e, _ = run(False)
    l_g, model = run(True)
    gs = next(iter(model.__dict__["_graphed_steps"].values()))
    assert gs.graph is not None and len(gs.ready["img"]) == 3
    assert all(math.isfinite(v) for v in l_g), l_g
    assert abs(l_e[0] - l_e[1]) > 1e-2
    for a, b in zip(l_e, l_g):
        assert abs(a - b) < max(1e-3, 0.02 * abs(a)), (l_e, l_g)
