"""Launch the three fused attention kernels (csrc/attn_tc.cu) once each at BigGAN-128's generator shape (batch 64, 4096 queries
x 1024 keys, 24 / 96 channels) after a warm-up that ncu skips (cudaProfilerStart), for `ncu --set full`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from compare_gan_b200 import kernels as K, tape

K.init(0)
K.set_math_mode(1)
bsz, lq, lk, dk, dv = int(os.environ.get("ATTN_B", "64")), 4096, 1024, 24, 96
dev = K._RT["device"]
th, ph, g, gy = [tape.DT(torch.randn(*s, device=dev) * 0.3) for s in ((bsz, lq, dk), (bsz, lk, dk), (bsz, lk, dv), (bsz, lq, dv))]
out, lse = K.empty(bsz, lq, dv), K.empty(bsz, lq)
dq, dkk, dvv = K.empty(bsz, lq, dk), K.empty(bsz, lk, dk), K.empty(bsz, lk, dv)
for it in range(2):
  if it == 1:
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
  K._call("attention_fwd", th.ptr, ph.ptr, g.ptr, out.ptr, lse.ptr, bsz, lq, lk, dk, dv)
  K._call("attention_bwd", th.ptr, ph.ptr, g.ptr, out.ptr, lse.ptr, gy.ptr, dq.ptr, dkk.ptr, dvv.ptr, bsz, lq, lk, dk, dv)
  torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
