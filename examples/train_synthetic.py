"""The reference's training loop (train.py:84-151) on synthetic data, driving gcc_b200 through the same calls.

    python examples/train_synthetic.py --model pix2pix --iters 20 --batch_size 8
    python examples/train_synthetic.py --model cyclegan | srgan | sagan
    torchrun --nproc-per-node 2 examples/train_synthetic.py --model pix2pix          (data parallel)

Everything after ``get_model_class`` is the reference's own sequence of calls; only the dataset is replaced by
random tensors of the shapes its data loaders produce (SURVEY.md section 8b "input dict contract")."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gcc_b200 import options  # noqa: E402


def batch_for(model_name, b, device="cpu"):
    if model_name == "srgan":
        return {"lr": torch.randn(b, 3, 24, 24), "hr": torch.rand(b, 3, 96, 96) * 2 - 1, "lr_names": "", "hr_names": ""}
    if model_name == "sagan":
        return {"z": torch.randn(b, 128), "real_img": torch.rand(b, 3, 64, 64) * 2 - 1, "img_path": ""}
    return {"A": torch.rand(b, 3, 256, 256) * 2 - 1, "B": torch.rand(b, 3, 256, 256) * 2 - 1, "A_paths": "", "B_paths": ""}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="pix2pix", choices=["pix2pix", "cyclegan", "srgan", "sagan"])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--batch_size", type=int, default=4)
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    roots = {"pix2pix": "./database/cityscapes/", "cyclegan": "./database/horse2zebra", "srgan": "./database/sr",
             "sagan": "./database/celeb"}
    argv = ["--dataroot", roots[a.model], "--model", a.model, "--darts_discriminator", "--online_distillation",
            "--lambda_content", "50" if a.model == "pix2pix" else "1e-2", "--lambda_gram", "1e4" if a.model == "pix2pix" else "1e-1",
            "--gpu_ids", str(local), "--ngf", "32", "--teacher_ngf", "64"]
    opt = options.parse(argv)
    opt.batch_size = a.batch_size
    model_class = options.get_model_class(opt)
    model = model_class(opt)
    # train.py:92-105: the teacher is the same class at teacher widths with a plain discriminator
    import copy
    topt = copy.deepcopy(opt)
    topt.ngf, topt.ndf = opt.teacher_ngf, opt.teacher_ndf
    topt.darts_discriminator = topt.online_distillation = False
    teacher = model_class(topt)
    teacher.model_train()
    setattr(model, "teacher_model", teacher)
    model.init_distillation()
    teacher.init_distillation()
    model.model_train()
    t0 = time.time()
    for it in range(a.iters):
        model.set_input(batch_for(a.model, a.batch_size))
        model.optimize_parameters()
        model.set_input(batch_for(a.model, a.batch_size))       # a validation batch in the reference
        model.clipping_mask_alpha()
        model.optimizer_netD_arch()
        if it % 5 == 0 and int(os.environ.get("RANK", "0")) == 0:
            print("iter %d " % it + " ".join("%s: %.3f" % kv for kv in model.get_current_losses().items()), flush=True)
    torch.cuda.synchronize()
    if int(os.environ.get("RANK", "0")) == 0:
        print("%.1f iterations/s" % (a.iters / (time.time() - t0)))


if __name__ == "__main__":
    main()
