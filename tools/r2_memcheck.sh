#!/bin/bash
# compute-sanitizer memcheck over small cases of every kernel of the forward (edge resolutions, ViT-B chain, fused head,
# host path).  One tool per gpurun call (B200_PROFILING.md).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
cat > /tmp/memcheck_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from dino_b200 import DINOSeg, synthetic
def run(arch, nb, res, batch, head="mlp", ncls=7):
    cfg = synthetic.make_config(arch, nb, ncls, head=head)
    m = DINOSeg(head=head, n_blocks=nb, n_classes=ncls, arch=arch)
    m.load_state_dict(synthetic.init_state_dict(cfg, 1, "trained_like"), strict=True)
    m = m.to("cuda:0"); m.set_resolution(res)
    x = synthetic.make_frames(batch, res, seed=3)
    lp, low, lab = m.infer(x.cuda(), want_logprobs=True, want_lowres=True, want_labels=True)
    host = m.predict_batch(x.pin_memory(), output="labels")
    t = m.predict_batch_async(x.pin_memory()); host2 = m.predict_wait(t)
    torch.cuda.synchronize()
    assert (host == lab.cpu().numpy()).all() and (host2 == host).all() and torch.isfinite(lp).all()
    print("ok", arch, nb, res, batch, head, flush=True)
run("vit_small", 1, 8, 3)          # one patch per frame
run("vit_small", 2, 64, 5)         # a single key tile, ragged row blocks
run("vit_small", 1, 136, 2)        # 290 tokens: odd number of query tiles, dual attention items
run("vit_small", 1, 496, 1)        # output that is not 480 x 480 (p = 7, odd)
run("vit_small", 1, 240, 2, head="linear", ncls=5)
run("vit_base", 1, 64, 2)          # ViT-B: unfused MLP, kernel-chain head
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/memcheck_case.py > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "^ok|ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/memcheck.log | head -30; tail -3 gpurun_out/memcheck.log
