"""Run in a CHILD process by tests/test_eval_ckpt_gpu.py: builds the UNMODIFIED reference model (oracle/_ref) on the CPU,
writes a checkpoint in the reference's own format (builder/utils/logger.py:166-177: {'model', 'optimizer', 'best_step',
'last_step', 'score', 'epoch'}) and the reference's eval-mode logits on a seeded batch with real pixels (its own Swin-T).

    python tools/make_reference_ckpt.py <out.pth> <out_logits.pt> <n_layers> <B> <L>
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_ckpt, out_logits, nl, B, L = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    import torch
    from medical_tri_modal_pilot_b200 import synth
    from oracle import ref_loader
    args, mod, _ = ref_loader.load(nl, B, 1, 0.0, "cpu")
    torch.manual_seed(123)
    model = mod.TRI_MBT_VSLTCLS(args)
    # non-trivial BatchNorm running statistics (a trained checkpoint has them)
    with torch.no_grad():
        model.fc_list[1].running_mean.normal_(0, 0.3)
        model.fc_list[1].running_var.uniform_(0.5, 1.5)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-6)
    torch.save({"model": model.state_dict(), "optimizer": opt.state_dict(), "best_step": 7, "last_step": None,
                "score": 0.5, "epoch": 1}, out_ckpt)
    model.eval()
    hb = synth.make_batch(B, L, n_img=3, seed=77, missing_mode="mixed", with_pixels=True, feats=False)
    with torch.no_grad():
        out, _, _ = model(hb["x"], None, None, None, None, hb["age"], hb["gen"], hb["input_lengths"].clone(), hb["txts"],
                          hb["txt_lengths"].clone(), hb["img"], hb["missing"], None, hb["img_time"].clone(),
                          hb["txt_time"].clone(), "test", None, None)
    torch.save(out, out_logits)


if __name__ == "__main__":
    main()
